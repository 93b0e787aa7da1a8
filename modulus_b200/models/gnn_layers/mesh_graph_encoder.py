"""MeshGraphEncoder (reference: physicsnemo/models/gnn_layers/mesh_graph_encoder.py:29-137).

Bipartite grid -> mesh block of GraphCast: the tuple form `(src_feat, dst_feat)` of the operator
seam with N_src != N_dst.  Same constructor, parameter names (`edge_mlp`, `src_node_mlp`,
`dst_node_mlp`) and init order as the reference, so `state_dict`s are interchangeable."""
from __future__ import annotations

from typing import Tuple

import torch.nn as nn
from torch import Tensor

from .mesh_graph_mlp import MeshGraphEdgeMLPConcat, MeshGraphEdgeMLPSum, MeshGraphMLP, compute_dtype
from .utils import aggregate_and_concat


class MeshGraphEncoder(nn.Module):
    """efeat = edge_mlp(g2m_efeat, (grid, mesh)); mesh' = mesh + dst_node_mlp(cat(agg(efeat), mesh));
    grid' = grid + src_node_mlp(grid); returns (grid', mesh').  No residual on the edge features
    (mesh_graph_encoder.py:130-137)."""

    def __init__(
        self,
        aggregation: str = "sum",
        input_dim_src_nodes: int = 512,
        input_dim_dst_nodes: int = 512,
        input_dim_edges: int = 512,
        output_dim_src_nodes: int = 512,
        output_dim_dst_nodes: int = 512,
        output_dim_edges: int = 512,
        hidden_dim: int = 512,
        hidden_layers: int = 1,
        activation_fn: nn.Module = nn.SiLU(),
        norm_type: str = "LayerNorm",
        do_concat_trick: bool = False,
        recompute_activation: bool = False,
    ):
        super().__init__()
        self.aggregation = aggregation
        # construction order (edge MLP first) fixes the RNG stream, hence the initial weights, to the reference's
        shared = dict(hidden_dim=hidden_dim, hidden_layers=hidden_layers, activation_fn=activation_fn,
                      norm_type=norm_type, recompute_activation=recompute_activation)
        MLP = MeshGraphEdgeMLPSum if do_concat_trick else MeshGraphEdgeMLPConcat
        self.edge_mlp = MLP(efeat_dim=input_dim_edges, src_dim=input_dim_src_nodes, dst_dim=input_dim_dst_nodes,
                            output_dim=output_dim_edges, **shared)
        self.src_node_mlp = MeshGraphMLP(input_dim=input_dim_src_nodes, output_dim=output_dim_src_nodes, **shared)
        self.dst_node_mlp = MeshGraphMLP(input_dim=input_dim_dst_nodes + output_dim_edges,
                                         output_dim=output_dim_dst_nodes, **shared)

    def forward(self, g2m_efeat: Tensor, grid_nfeat: Tensor, mesh_nfeat: Tensor, graph) -> Tuple[Tensor, Tensor]:
        dt = compute_dtype(g2m_efeat)
        efeat = self.edge_mlp(g2m_efeat, (grid_nfeat, mesh_nfeat), graph)
        cat_feat = aggregate_and_concat(efeat, mesh_nfeat.to(dt), graph, self.aggregation)
        # residual adds fused into the MLP kernels' last epilogue
        mesh_new = self.dst_node_mlp.mlp(cat_feat, residual=mesh_nfeat)
        grid_new = self.src_node_mlp.mlp(grid_nfeat, residual=grid_nfeat)
        return grid_new, mesh_new
