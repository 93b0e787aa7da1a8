"""CuGraphCSC -- CSC graph container of the MeshGraphNet path.

Same constructor, attributes and method names as the reference class
(physicsnemo/models/gnn_layers/graph.py:48-479).  Differences are behind the interface:
instead of converting to cugraph-ops `StaticCSC/BipartiteCSC` or a DGL heterograph the
container lazily builds one `GraphPlan` (int32 CSC + CSR transpose + expanded destination
ids, built by CUDA kernels) that all modulus_b200 operators share.
"""
from __future__ import annotations

from typing import Any, List, Optional

import torch
from torch import Tensor

from ...ops import GraphPlan, require_cuda
from .distributed_graph import DistributedGraph, GraphPartition, partition_graph_by_coordinate_bbox


class CuGraphCSC:
    """Generic CSC graph wrapper (reference: graph.py:48-141).

    Parameters
    ----------
    offsets, indices : Tensor
        CSC offsets ``[num_dst_nodes+1]`` and source ids per in-edge ``[E]`` (int32 or int64).
    num_src_nodes, num_dst_nodes : int
    ef_indices : Optional[Tensor]
        maps CSC positions to rows of COO-ordered edge features (not with partitioning)
    reverse_graph_bwd, cache_graph : bool
        kept for API compatibility; the CSR transpose is always built once and cached unless
        ``cache_graph`` is False
    partition_size, partition_group_name, graph_partition
        distribute the graph over a process group (see ``DistributedGraph``)
    """

    def __init__(
        self,
        offsets: Tensor,
        indices: Tensor,
        num_src_nodes: int,
        num_dst_nodes: int,
        ef_indices: Optional[Tensor] = None,
        reverse_graph_bwd: bool = True,
        cache_graph: bool = True,
        partition_size: Optional[int] = -1,
        partition_group_name: Optional[str] = None,
        graph_partition: Optional[GraphPartition] = None,
    ) -> None:
        self.offsets = offsets
        self.indices = indices
        self.num_src_nodes = num_src_nodes
        self.num_dst_nodes = num_dst_nodes
        self.ef_indices = ef_indices
        self.reverse_graph_bwd = reverse_graph_bwd
        self.cache_graph = cache_graph

        # kept so reference-style code probing these attributes keeps working
        self.bipartite_csc = None
        self.static_csc = None
        self.dgl_graph = None
        self._plan: Optional[GraphPlan] = None

        self.is_distributed = False
        self.dist_csc = None

        if partition_size is None or partition_size <= 1:
            self.is_distributed = False
            return

        if self.ef_indices is not None:
            raise AssertionError("DistributedGraph does not support mapping CSC-indices to COO-indices.")

        self.dist_graph = DistributedGraph(
            self.offsets,
            self.indices,
            partition_size,
            partition_group_name,
            graph_partition=graph_partition,
        )

        # overwrite graph information with local graph after distribution (graph.py:137-141)
        self.offsets = self.dist_graph.graph_partition.local_offsets
        self.indices = self.dist_graph.graph_partition.local_indices
        self.num_src_nodes = self.dist_graph.graph_partition.num_local_src_nodes
        self.num_dst_nodes = self.dist_graph.graph_partition.num_local_dst_nodes
        self.is_distributed = True

    # ------------------------------------------------------------------ plan (B200 backend)
    def b200_plan(self) -> GraphPlan:
        """Index structures consumed by the CUDA operators (built once, cached)."""
        if self._plan is None or not self.cache_graph:
            require_cuda(self.offsets, self.indices)
            if self.ef_indices is None:
                plan = GraphPlan.from_csc(self.offsets, self.indices, self.num_src_nodes, self.num_dst_nodes)
            else:
                # edge features stay in COO order: row ef_indices[j] belongs to CSC position j
                base = GraphPlan.from_csc(self.offsets, self.indices, self.num_src_nodes, self.num_dst_nodes)
                ef = self.ef_indices.to(device=base.device, dtype=torch.int64)
                src = torch.empty_like(base.src)
                dst = torch.empty_like(base.dst)
                src[ef] = base.src
                dst[ef] = base.dst
                plan = GraphPlan.from_coo(src, dst, self.num_src_nodes, self.num_dst_nodes)
            self._plan = plan
        return self._plan

    @staticmethod
    def from_dgl(
        graph,
        partition_size: int = 1,
        partition_group_name: Optional[str] = None,
        partition_by_bbox: bool = False,
        src_coordinates: Optional[torch.Tensor] = None,
        dst_coordinates: Optional[torch.Tensor] = None,
        coordinate_separators_min: Optional[List[List[Optional[float]]]] = None,
        coordinate_separators_max: Optional[List[List[Optional[float]]]] = None,
    ):
        """Build from any object with the DGLGraph accessors used by the reference
        (graph.py:143-193): ``adj_tensors("csc")`` / ``adj_sparse("csc")`` or ``edges()``.
        Returns ``(CuGraphCSC, edge_perm)``; permute COO-ordered edge features by
        ``edge_perm`` (stable sort by destination)."""
        if hasattr(graph, "adj_tensors"):
            offsets, indices, edge_perm = graph.adj_tensors("csc")
        elif hasattr(graph, "adj_sparse"):
            offsets, indices, edge_perm = graph.adj_sparse("csc")
        elif hasattr(graph, "edges"):
            src, dst = graph.edges()
            n_dst = graph.num_dst_nodes()
            edge_perm = torch.argsort(dst.long(), stable=True)
            deg = torch.bincount(dst.long(), minlength=n_dst)
            offsets = torch.zeros(n_dst + 1, dtype=torch.int64, device=dst.device)
            offsets[1:] = torch.cumsum(deg, 0)
            indices = src.long()[edge_perm]
        else:
            raise ValueError("Passed graph object doesn't support conversion to CSC.")

        n_src_nodes, n_dst_nodes = (graph.num_src_nodes(), graph.num_dst_nodes())

        graph_partition = None
        if partition_by_bbox and partition_size > 1:
            from ...distributed import DistributedManager

            dist_manager = DistributedManager()
            partition_rank = dist_manager.group_rank(name=partition_group_name)
            graph_partition = partition_graph_by_coordinate_bbox(
                offsets.to(dtype=torch.int64),
                indices.to(dtype=torch.int64),
                src_coordinates=src_coordinates,
                dst_coordinates=dst_coordinates,
                coordinate_separators_min=coordinate_separators_min,
                coordinate_separators_max=coordinate_separators_max,
                partition_size=partition_size,
                partition_rank=partition_rank,
                device=dist_manager.device,
            )

        graph_csc = CuGraphCSC(
            offsets.to(dtype=torch.int64),
            indices.to(dtype=torch.int64),
            n_src_nodes,
            n_dst_nodes,
            partition_size=partition_size,
            partition_group_name=partition_group_name,
            graph_partition=graph_partition,
        )
        return graph_csc, edge_perm

    # ------------------------------------------------------------------ distributed facade
    def get_src_node_features_in_partition(self, global_src_feat, scatter_features: bool = False, src_rank: int = 0):
        if self.is_distributed:
            return self.dist_graph.get_src_node_features_in_partition(
                global_src_feat, scatter_features=scatter_features, src_rank=src_rank)
        return global_src_feat

    def get_src_node_features_in_local_graph(self, local_src_feat):
        """Halo exchange: all source rows the local graph references (graph.py:211-224)."""
        if self.is_distributed:
            return self.dist_graph.get_src_node_features_in_local_graph(local_src_feat)
        return local_src_feat

    def get_dst_node_features_in_partition(self, global_dst_feat, scatter_features: bool = False, src_rank: int = 0):
        if self.is_distributed:
            return self.dist_graph.get_dst_node_features_in_partition(
                global_dst_feat, scatter_features=scatter_features, src_rank=src_rank)
        return global_dst_feat

    def get_edge_features_in_partition(self, global_efeat, scatter_features: bool = False, src_rank: int = 0):
        if self.is_distributed:
            return self.dist_graph.get_edge_features_in_partition(
                global_efeat, scatter_features=scatter_features, src_rank=src_rank)
        return global_efeat

    def get_global_src_node_features(self, local_nfeat, get_on_all_ranks: bool = True, dst_rank: int = 0):
        if self.is_distributed:
            return self.dist_graph.get_global_src_node_features(local_nfeat, get_on_all_ranks, dst_rank=dst_rank)
        return local_nfeat

    def get_global_dst_node_features(self, local_nfeat, get_on_all_ranks: bool = True, dst_rank: int = 0):
        if self.is_distributed:
            return self.dist_graph.get_global_dst_node_features(local_nfeat, get_on_all_ranks, dst_rank=dst_rank)
        return local_nfeat

    def get_global_edge_features(self, local_efeat, get_on_all_ranks: bool = True, dst_rank: int = 0):
        if self.is_distributed:
            return self.dist_graph.get_global_edge_features(local_efeat, get_on_all_ranks, dst_rank=dst_rank)
        return local_efeat

    # ------------------------------------------------------------------ misc
    def to(self, *args: Any, **kwargs: Any) -> "CuGraphCSC":
        """Move offsets/indices(/ef_indices); only int32/int64 dtypes (graph.py:315-345)."""
        device, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        if dtype not in (None, torch.int32, torch.int64):
            raise TypeError(f"Invalid dtype, expected torch.int32 or torch.int64, got {dtype}.")
        self.offsets = self.offsets.to(device=device, dtype=dtype)
        self.indices = self.indices.to(device=device, dtype=dtype)
        if self.ef_indices is not None:
            self.ef_indices = self.ef_indices.to(device=device, dtype=dtype)
        self._plan = None
        return self

    def to_bipartite_csc(self, dtype=None):
        raise RuntimeError("Conversion failed, expected cugraph-ops to be installed. "
                           "(modulus_b200 uses CuGraphCSC.b200_plan() instead)")

    def to_static_csc(self, dtype=None):
        raise RuntimeError("Conversion failed, expected cugraph-ops to be installed. "
                           "(modulus_b200 uses CuGraphCSC.b200_plan() instead)")

    def to_dgl_graph(self):
        raise RuntimeError("modulus_b200 has no DGL dispatch; operators consume CuGraphCSC.b200_plan()")
