"""MeshGraphDecoder (reference: physicsnemo/models/gnn_layers/mesh_graph_decoder.py:29-121).

Bipartite mesh -> grid block of GraphCast through the tuple form of the operator seam.  Same
constructor, parameter names (`edge_mlp`, `node_mlp`) and init order as the reference."""
from __future__ import annotations

import torch.nn as nn
from torch import Tensor

from .mesh_graph_mlp import MeshGraphEdgeMLPConcat, MeshGraphEdgeMLPSum, MeshGraphMLP, compute_dtype
from .utils import aggregate_and_concat


class MeshGraphDecoder(nn.Module):
    """efeat = edge_mlp(m2g_efeat, (mesh, grid)); grid' = node_mlp(cat(agg(efeat), grid)) + grid
    (mesh_graph_decoder.py:114-121)."""

    def __init__(
        self,
        aggregation: str = "sum",
        input_dim_src_nodes: int = 512,
        input_dim_dst_nodes: int = 512,
        input_dim_edges: int = 512,
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
        self.node_mlp = MeshGraphMLP(input_dim=input_dim_dst_nodes + output_dim_edges, output_dim=output_dim_dst_nodes, **shared)

    def forward(self, m2g_efeat: Tensor, grid_nfeat: Tensor, mesh_nfeat: Tensor, graph) -> Tensor:
        dt = compute_dtype(m2g_efeat)
        efeat = self.edge_mlp(m2g_efeat, (mesh_nfeat, grid_nfeat), graph)
        cat_feat = aggregate_and_concat(efeat, grid_nfeat.to(dt), graph, self.aggregation)
        return self.node_mlp.mlp(cat_feat, residual=grid_nfeat)
