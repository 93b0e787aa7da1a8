"""MeshNodeBlock (reference: physicsnemo/models/gnn_layers/mesh_node_block.py:28-92)."""
from __future__ import annotations

from typing import Tuple

import torch.nn as nn
from torch import Tensor

from .mesh_graph_mlp import MeshGraphMLP, compute_dtype
from .utils import aggregate_and_concat


class MeshNodeBlock(nn.Module):
    """nfeat' = node_mlp(cat(aggregate(efeat by destination), nfeat)) + nfeat ; returns (efeat, nfeat').

    Same constructor as the reference; aggregation is "sum" or "mean"."""

    def __init__(
        self,
        aggregation: str = "sum",
        input_dim_nodes: int = 512,
        input_dim_edges: int = 512,
        output_dim: int = 512,
        hidden_dim: int = 512,
        hidden_layers: int = 1,
        activation_fn: nn.Module = nn.SiLU(),
        norm_type: str = "LayerNorm",
        recompute_activation: bool = False,
    ):
        super().__init__()
        self.aggregation = aggregation
        self.node_mlp = MeshGraphMLP(
            input_dim=input_dim_nodes + input_dim_edges,
            output_dim=output_dim,
            hidden_dim=hidden_dim,
            hidden_layers=hidden_layers,
            activation_fn=activation_fn,
            norm_type=norm_type,
            recompute_activation=recompute_activation,
        )

    def forward(self, efeat: Tensor, nfeat: Tensor, graph) -> Tuple[Tensor, Tensor]:
        dt = compute_dtype(efeat)
        cat_feat = aggregate_and_concat(efeat.to(dt), nfeat.to(dt), graph, self.aggregation)
        nfeat_new = self.node_mlp.mlp(cat_feat, residual=nfeat)
        return efeat, nfeat_new
