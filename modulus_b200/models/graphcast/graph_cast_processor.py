"""GraphCastProcessor (reference: physicsnemo/models/graphcast/graph_cast_processor.py:30-200).

L x (MeshEdgeBlock, MeshNodeBlock) on the multi-mesh with optional checkpoint segments; unlike
MeshGraphNetProcessor it returns BOTH (efeat, nfeat).  The graph-transformer variant of the
reference (:203-) is attention, not message passing, and is out of scope."""
from __future__ import annotations

from typing import Callable, Tuple

import torch.nn as nn
from torch import Tensor

from ..gnn_layers.mesh_edge_block import MeshEdgeBlock
from ..gnn_layers.mesh_node_block import MeshNodeBlock
from ..gnn_layers.utils import set_checkpoint_fn


class GraphCastProcessor(nn.Module):
    def __init__(
        self,
        aggregation: str = "sum",
        processor_layers: int = 16,
        input_dim_nodes: int = 512,
        input_dim_edges: int = 512,
        hidden_dim: int = 512,
        hidden_layers: int = 1,
        activation_fn: nn.Module = nn.SiLU(),
        norm_type: str = "LayerNorm",
        do_concat_trick: bool = False,
        recompute_activation: bool = False,
    ):
        super().__init__()
        layers = []
        for _ in range(processor_layers):
            layers.append(MeshEdgeBlock(input_dim_nodes, input_dim_edges, input_dim_edges, hidden_dim,
                                        hidden_layers, activation_fn, norm_type, do_concat_trick,
                                        recompute_activation))
            layers.append(MeshNodeBlock(aggregation, input_dim_nodes, input_dim_edges, input_dim_nodes,
                                        hidden_dim, hidden_layers, activation_fn, norm_type,
                                        recompute_activation))
        self.processor_layers = nn.ModuleList(layers)
        self.num_processor_layers = len(self.processor_layers)
        self.checkpoint_segments = [(0, self.num_processor_layers)]
        self.checkpoint_fn = set_checkpoint_fn(False)

    def set_checkpoint_segments(self, checkpoint_segments: int):
        """Reference: graph_cast_processor.py:107-134 (ValueError when the layer count is not a
        multiple of the segment count)."""
        if checkpoint_segments > 0:
            if self.num_processor_layers % checkpoint_segments != 0:
                raise ValueError("Processor layers must be a multiple of checkpoint_segments")
            size = self.num_processor_layers // checkpoint_segments
            self.checkpoint_segments = [(i, i + size) for i in range(0, self.num_processor_layers, size)]
            self.checkpoint_fn = set_checkpoint_fn(True)
        else:
            self.checkpoint_fn = set_checkpoint_fn(False)
            self.checkpoint_segments = [(0, self.num_processor_layers)]

    def run_function(self, segment_start: int, segment_end: int) -> Callable:
        segment = self.processor_layers[segment_start:segment_end]

        def custom_forward(efeat: Tensor, nfeat: Tensor, graph) -> Tuple[Tensor, Tensor]:
            for module in segment:
                efeat, nfeat = module(efeat, nfeat, graph)
            return efeat, nfeat

        return custom_forward

    def forward(self, efeat: Tensor, nfeat: Tensor, graph) -> Tuple[Tensor, Tensor]:
        for segment_start, segment_end in self.checkpoint_segments:
            efeat, nfeat = self.checkpoint_fn(
                self.run_function(segment_start, segment_end), efeat, nfeat, graph,
                use_reentrant=False, preserve_rng_state=False,
            )
        return efeat, nfeat
