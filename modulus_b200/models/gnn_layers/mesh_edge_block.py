"""MeshEdgeBlock (reference: physicsnemo/models/gnn_layers/mesh_edge_block.py:30-96)."""
from __future__ import annotations

from typing import Tuple

import torch.nn as nn
from torch import Tensor

from .mesh_graph_mlp import MeshGraphEdgeMLPConcat, MeshGraphEdgeMLPSum


class MeshEdgeBlock(nn.Module):
    """efeat' = edge_mlp(concat(efeat, nfeat[src], nfeat[dst])) + efeat ; returns (efeat', nfeat).

    Same constructor as the reference.  The residual add (mesh_edge_block.py:95) is fused into the
    LayerNorm epilogue of the MLP kernels."""

    def __init__(
        self,
        input_dim_nodes: int = 512,
        input_dim_edges: int = 512,
        output_dim: int = 512,
        hidden_dim: int = 512,
        hidden_layers: int = 1,
        activation_fn: nn.Module = nn.SiLU(),
        norm_type: str = "LayerNorm",
        do_concat_trick: bool = False,
        recompute_activation: bool = False,
    ):
        super().__init__()
        MLP = MeshGraphEdgeMLPSum if do_concat_trick else MeshGraphEdgeMLPConcat
        self.edge_mlp = MLP(
            efeat_dim=input_dim_edges,
            src_dim=input_dim_nodes,
            dst_dim=input_dim_nodes,
            output_dim=output_dim,
            hidden_dim=hidden_dim,
            hidden_layers=hidden_layers,
            activation_fn=activation_fn,
            norm_type=norm_type,
            recompute_activation=recompute_activation,
        )

    def forward(self, efeat: Tensor, nfeat: Tensor, graph) -> Tuple[Tensor, Tensor]:
        efeat_new = self.edge_mlp.edge_mlp(efeat, nfeat, graph, residual=efeat)
        return efeat_new, nfeat
