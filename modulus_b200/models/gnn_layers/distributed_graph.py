"""Graph partitioning and the distributed-graph facade of the MeshGraphNet path.

Public names, argument lists, the `GraphPartition` fields and every index map are those of
the reference (physicsnemo/models/gnn_layers/distributed_graph.py:35-1197); maps are
bit-exact (tests/test_partition.py compares against the reference's known-answer tests and
against partitions produced by the unmodified reference).  The construction is vectorised:
the reference expands every destination's edge range with a Python loop
(`for i in range(len(offset_start))`, :306-334), here one repeat_interleave does it, so an
8 M-node mesh partitions in seconds instead of minutes.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

from ...distributed import DistributedManager, all_gather_v, gather_v, indexed_all_to_all_v, scatter_v

logger = logging.getLogger(__name__)


@dataclass
class GraphPartition:
    """All buffers that define one rank's partition of a global CSC graph and the halo exchange
    it needs (reference: distributed_graph.py:35-151; same field names and meaning).

    Destination nodes and all their incoming edges live on exactly one rank.  Source rows needed
    by the local edges but owned elsewhere are fetched per layer with an indexed all-to-all-v:
    `scatter_indices[r]` = partition-local source rows this rank sends to rank r,
    `sizes[p][r]` = number of rows rank p sends to rank r.  The local source id space is the
    concatenation, in rank order, of the (sorted) global ids received from each rank.
    """

    partition_size: int
    partition_rank: int
    device: torch.device
    matrix_decomp: bool = False

    local_offsets: Optional[torch.Tensor] = None
    local_indices: Optional[torch.Tensor] = None
    num_local_src_nodes: int = -1
    num_local_dst_nodes: int = -1
    num_local_indices: int = -1
    map_partitioned_src_ids_to_global: Optional[torch.Tensor] = None
    map_concatenated_local_src_ids_to_global: Optional[torch.Tensor] = None
    map_partitioned_dst_ids_to_global: Optional[torch.Tensor] = None
    map_concatenated_local_dst_ids_to_global: Optional[torch.Tensor] = None
    map_partitioned_edge_ids_to_global: Optional[torch.Tensor] = None
    map_concatenated_local_edge_ids_to_global: Optional[torch.Tensor] = None
    map_global_src_ids_to_concatenated_local: Optional[torch.Tensor] = None
    map_global_dst_ids_to_concatenated_local: Optional[torch.Tensor] = None
    map_global_edge_ids_to_concatenated_local: Optional[torch.Tensor] = None

    sizes: Optional[List[List[int]]] = None
    scatter_indices: Optional[List[torch.Tensor]] = None
    num_src_nodes_in_each_partition: Optional[List[int]] = None
    num_dst_nodes_in_each_partition: Optional[List[int]] = None
    num_indices_in_each_partition: Optional[List[int]] = None

    def __post_init__(self):
        if self.partition_size <= 0:
            raise ValueError(f"Expected partition_size > 0, got {self.partition_size}")
        if not (0 <= self.partition_rank < self.partition_size):
            raise ValueError(
                f"Expected 0 <= partition_rank < {self.partition_size}, got {self.partition_rank}"
            )
        if self.sizes is None:
            self.sizes = [[None for _ in range(self.partition_size)] for _ in range(self.partition_size)]
        if self.scatter_indices is None:
            self.scatter_indices = [None] * self.partition_size
        if self.num_src_nodes_in_each_partition is None:
            self.num_src_nodes_in_each_partition = [None] * self.partition_size
        if self.num_dst_nodes_in_each_partition is None:
            self.num_dst_nodes_in_each_partition = [None] * self.partition_size
        if self.num_indices_in_each_partition is None:
            self.num_indices_in_each_partition = [None] * self.partition_size

    def to(self, *args, **kwargs):
        for attr in dir(self):
            attr_val = getattr(self, attr)
            if isinstance(attr_val, torch.Tensor):
                setattr(self, attr, attr_val.to(*args, **kwargs))
        self.scatter_indices = [idx.to(*args, **kwargs) for idx in self.scatter_indices]
        return self


def _expand_ranges(start: torch.Tensor, end: torch.Tensor, dtype, device) -> torch.Tensor:
    """cat([arange(start[i], end[i]) for i]) without the Python loop."""
    deg = (end - start).to(torch.int64)
    total = int(deg.sum().item())
    if total == 0:
        return torch.empty(0, dtype=dtype, device=device)
    seg_start = torch.cumsum(deg, 0) - deg
    base = torch.repeat_interleave(start.to(torch.int64) - seg_start, deg)
    return (torch.arange(total, dtype=torch.int64, device=device) + base).to(dtype)


def partition_graph_with_id_mapping(
    global_offsets: torch.Tensor,
    global_indices: torch.Tensor,
    mapping_src_ids_to_ranks: torch.Tensor,
    mapping_dst_ids_to_ranks: torch.Tensor,
    partition_size: int,
    partition_rank: int,
    device: torch.device,
) -> GraphPartition:
    """Partition a global CSC graph given id -> rank maps for the source and destination id spaces
    (reference: distributed_graph.py:154-398).  Every rank derives all P partitions' sizes (needed
    for the all-to-all) and keeps its own local graph."""
    graph_partition = GraphPartition(partition_size=partition_size, partition_rank=partition_rank, device=device)

    dst_nodes_in_each_partition = [None] * partition_size
    src_nodes_in_each_partition = [None] * partition_size
    num_dst_nodes_in_each_partition = [None] * partition_size
    num_src_nodes_in_each_partition = [None] * partition_size

    dtype = global_indices.dtype
    input_device = global_indices.device

    graph_partition.map_concatenated_local_src_ids_to_global = torch.empty_like(mapping_src_ids_to_ranks)
    graph_partition.map_concatenated_local_dst_ids_to_global = torch.empty_like(mapping_dst_ids_to_ranks)
    graph_partition.map_concatenated_local_edge_ids_to_global = torch.empty_like(global_indices)
    graph_partition.map_global_src_ids_to_concatenated_local = torch.empty_like(mapping_src_ids_to_ranks)
    graph_partition.map_global_dst_ids_to_concatenated_local = torch.empty_like(mapping_dst_ids_to_ranks)
    graph_partition.map_global_edge_ids_to_concatenated_local = torch.empty_like(global_indices)
    _map_global_src_ids_to_local = torch.empty_like(mapping_src_ids_to_ranks)

    _src_id_offset = 0
    _dst_id_offset = 0
    _edge_id_offset = 0

    for rank in range(partition_size):
        dst_nodes_in_each_partition[rank] = torch.nonzero(mapping_dst_ids_to_ranks == rank).view(-1)
        src_nodes_in_each_partition[rank] = torch.nonzero(mapping_src_ids_to_ranks == rank).view(-1)
        num_nodes = dst_nodes_in_each_partition[rank].numel()
        if num_nodes == 0:
            raise RuntimeError(f"Aborting partitioning, rank {rank} has 0 destination nodes to work on.")
        num_dst_nodes_in_each_partition[rank] = num_nodes

        num_nodes = src_nodes_in_each_partition[rank].numel()
        num_src_nodes_in_each_partition[rank] = num_nodes
        if num_nodes == 0:
            raise RuntimeError(f"Aborting partitioning, rank {rank} has 0 source nodes to work on.")

        ids = src_nodes_in_each_partition[rank]
        mapped_ids = torch.arange(start=_src_id_offset, end=_src_id_offset + ids.numel(), dtype=dtype,
                                  device=input_device)
        _map_global_src_ids_to_local[ids] = (mapped_ids - _src_id_offset).to(_map_global_src_ids_to_local.dtype)
        graph_partition.map_global_src_ids_to_concatenated_local[ids] = mapped_ids.to(
            graph_partition.map_global_src_ids_to_concatenated_local.dtype)
        graph_partition.map_concatenated_local_src_ids_to_global[mapped_ids] = ids.to(
            graph_partition.map_concatenated_local_src_ids_to_global.dtype)
        _src_id_offset += ids.numel()

        ids = dst_nodes_in_each_partition[rank]
        mapped_ids = torch.arange(start=_dst_id_offset, end=_dst_id_offset + ids.numel(), dtype=dtype,
                                  device=input_device)
        graph_partition.map_global_dst_ids_to_concatenated_local[ids] = mapped_ids.to(
            graph_partition.map_global_dst_ids_to_concatenated_local.dtype)
        graph_partition.map_concatenated_local_dst_ids_to_global[mapped_ids] = ids.to(
            graph_partition.map_concatenated_local_dst_ids_to_global.dtype)
        _dst_id_offset += ids.numel()

    graph_partition.num_src_nodes_in_each_partition = num_src_nodes_in_each_partition
    graph_partition.num_dst_nodes_in_each_partition = num_dst_nodes_in_each_partition

    for rank in range(partition_size):
        offset_start = global_offsets[dst_nodes_in_each_partition[rank]].view(-1)
        offset_end = global_offsets[dst_nodes_in_each_partition[rank] + 1].view(-1)
        degree = offset_end - offset_start
        local_offsets = degree.view(-1).cumsum(dim=0)
        local_offsets = torch.cat([torch.zeros(1, dtype=dtype, device=input_device), local_offsets])

        # all in-edges of the owned destinations, destination-major, in CSC order
        partitioned_edge_ids = _expand_ranges(offset_start, offset_end, dtype, input_device)

        ids = partitioned_edge_ids
        mapped_ids = torch.arange(_edge_id_offset, _edge_id_offset + ids.numel(), device=ids.device, dtype=ids.dtype)
        graph_partition.map_global_edge_ids_to_concatenated_local[ids] = mapped_ids
        graph_partition.map_concatenated_local_edge_ids_to_global[mapped_ids] = ids
        _edge_id_offset += ids.numel()

        partitioned_src_ids = global_indices[partitioned_edge_ids]

        global_src_ids_on_rank, inverse_mapping = partitioned_src_ids.unique(sorted=True, return_inverse=True)
        remote_local_src_ids_on_rank = _map_global_src_ids_to_local[global_src_ids_on_rank]

        # local source id = position in [ids owned by rank 0 | rank 1 | ...], each block sorted
        owner = mapping_src_ids_to_ranks[global_src_ids_on_rank]
        _num_local_indices = 0
        local_id_of_unique = torch.empty_like(global_src_ids_on_rank)
        for rank_offset in range(partition_size):
            mask = owner == rank_offset
            if partition_rank == rank_offset:
                graph_partition.scatter_indices[rank] = (
                    remote_local_src_ids_on_rank[mask].detach().clone().to(dtype=torch.int64)
                )
            numel_mask = int(mask.sum().item())
            graph_partition.sizes[rank_offset][rank] = numel_mask
            local_id_of_unique[mask] = torch.arange(
                _num_local_indices, _num_local_indices + numel_mask, device=input_device, dtype=dtype
            ).to(local_id_of_unique.dtype)
            _num_local_indices += numel_mask

        local_indices = local_id_of_unique[inverse_mapping].to(mapping_src_ids_to_ranks.dtype)
        graph_partition.num_indices_in_each_partition[rank] = local_indices.size(0)

        if rank == partition_rank:
            graph_partition.local_offsets = local_offsets
            graph_partition.local_indices = local_indices
            graph_partition.num_local_indices = graph_partition.local_indices.size(0)
            graph_partition.num_local_dst_nodes = num_dst_nodes_in_each_partition[rank]
            graph_partition.num_local_src_nodes = global_src_ids_on_rank.size(0)
            graph_partition.map_partitioned_src_ids_to_global = src_nodes_in_each_partition[rank]
            graph_partition.map_partitioned_dst_ids_to_global = dst_nodes_in_each_partition[rank]
            graph_partition.map_partitioned_edge_ids_to_global = partitioned_edge_ids

    for r in range(graph_partition.partition_size):
        err_msg = "error in graph partition: list containing sizes of exchanged indices does not match the tensor of indices to be exchanged"
        if graph_partition.sizes[graph_partition.partition_rank][r] != graph_partition.scatter_indices[r].numel():
            raise AssertionError(err_msg)

    graph_partition = graph_partition.to(device=device)
    return graph_partition


def partition_graph_with_matrix_decomposition(
    global_offsets: torch.Tensor,
    global_indices: torch.Tensor,
    num_nodes: int,
    partition_book: torch.Tensor,
    partition_size: int,
    partition_rank: int,
    device: torch.device,
) -> GraphPartition:
    """1-D row decomposition of a square adjacency matrix given contiguous node ranges
    `partition_book` (reference: distributed_graph.py:401-562)."""
    graph_partition = GraphPartition(partition_size=partition_size, partition_rank=partition_rank, device=device)
    dtype = global_indices.dtype
    num_edges = global_indices.size(0)
    node_offset = partition_book[partition_rank]
    num_local_nodes = partition_book[partition_rank + 1] - partition_book[partition_rank]
    edge_partition_offset = global_offsets[node_offset]
    if node_offset + num_local_nodes > num_nodes:
        raise ValueError("Invalid node offset and number of local nodes")

    local_offsets = global_offsets[node_offset: node_offset + num_local_nodes + 1].to(device=device, non_blocking=True)
    graph_partition.local_offsets = local_offsets - edge_partition_offset
    graph_partition.num_local_dst_nodes = num_local_nodes

    partition_book = partition_book.to(device=device)
    for to_partition in range(partition_size):
        local_indices = global_indices[
            global_offsets[partition_book[to_partition]]: global_offsets[partition_book[to_partition + 1]]
        ].to(device=device, non_blocking=True)
        global_src_node_at_partition, inverse_indices = local_indices.unique(sorted=True, return_inverse=True)
        global_src_node_at_partition_rank = (
            torch.bucketize(global_src_node_at_partition, partition_book, right=True) - 1
        )
        src_node_indices = torch.nonzero(global_src_node_at_partition_rank == partition_rank, as_tuple=False).squeeze(1)
        graph_partition.scatter_indices[to_partition] = global_src_node_at_partition[src_node_indices] - node_offset
        graph_partition.num_indices_in_each_partition[to_partition] = local_indices.size(0)
        graph_partition.num_dst_nodes_in_each_partition[to_partition] = (
            partition_book[to_partition + 1] - partition_book[to_partition]
        )
        graph_partition.num_src_nodes_in_each_partition[to_partition] = global_src_node_at_partition.size(0)

        if to_partition == partition_rank:
            graph_partition.local_indices = inverse_indices
            graph_partition.num_local_indices = graph_partition.local_indices.size(0)
            graph_partition.num_local_src_nodes = global_src_node_at_partition.size(0)
            graph_partition.map_partitioned_src_ids_to_global = global_src_node_at_partition

        for from_partition in range(partition_size):
            graph_partition.sizes[from_partition][to_partition] = torch.count_nonzero(
                global_src_node_at_partition_rank == from_partition
            )

    graph_partition.map_partitioned_dst_ids_to_global = torch.arange(
        node_offset, node_offset + num_local_nodes, dtype=dtype, device=device
    )
    graph_partition.map_partitioned_edge_ids_to_global = torch.arange(
        edge_partition_offset, edge_partition_offset + graph_partition.num_local_indices, dtype=dtype, device=device
    )
    graph_partition.map_concatenated_local_src_ids_to_global = torch.arange(num_nodes, dtype=dtype, device=device)
    graph_partition.map_concatenated_local_edge_ids_to_global = torch.arange(num_edges, dtype=dtype, device=device)
    graph_partition.map_concatenated_local_dst_ids_to_global = graph_partition.map_concatenated_local_src_ids_to_global
    graph_partition.map_global_src_ids_to_concatenated_local = graph_partition.map_concatenated_local_src_ids_to_global
    graph_partition.map_global_dst_ids_to_concatenated_local = graph_partition.map_concatenated_local_src_ids_to_global
    graph_partition.map_global_edge_ids_to_concatenated_local = graph_partition.map_concatenated_local_edge_ids_to_global
    graph_partition.matrix_decomp = True

    for r in range(graph_partition.partition_size):
        err_msg = "error in graph partition: list containing sizes of exchanged indices does not match the tensor of indices to be exchanged"
        if graph_partition.sizes[graph_partition.partition_rank][r] != graph_partition.scatter_indices[r].numel():
            raise AssertionError(err_msg)

    graph_partition = graph_partition.to(device=device)
    return graph_partition


def partition_graph_nodewise(
    global_offsets: torch.Tensor,
    global_indices: torch.Tensor,
    partition_size: int,
    partition_rank: int,
    device: torch.device,
    matrix_decomp: bool = False,
) -> GraphPartition:
    """Equal-size chunks of the source and destination id spaces: owner(v) = v // ceil(N / P)
    (reference: distributed_graph.py:565-666)."""
    num_global_src_nodes = global_indices.max().item() + 1
    num_global_dst_nodes = global_offsets.size(0) - 1
    num_dst_nodes_per_partition = (num_global_dst_nodes + partition_size - 1) // partition_size

    if matrix_decomp:
        if num_global_src_nodes != num_global_dst_nodes:
            raise ValueError("Must be square adj. matrix (num_src=num_dst) for matrix decomposition")
        partition_book = torch.arange(0, num_global_dst_nodes, num_dst_nodes_per_partition, dtype=global_indices.dtype)
        partition_book = torch.cat([partition_book, torch.tensor([num_global_dst_nodes], dtype=global_indices.dtype)])
        return partition_graph_with_matrix_decomposition(
            global_offsets, global_indices, num_global_dst_nodes, partition_book, partition_size, partition_rank,
            device,
        )

    num_src_nodes_per_partition = (num_global_src_nodes + partition_size - 1) // partition_size

    mapping_dst_ids_to_ranks = (
        torch.arange(num_global_dst_nodes, dtype=global_offsets.dtype, device=global_offsets.device)
        // num_dst_nodes_per_partition
    )
    mapping_src_ids_to_ranks = (
        torch.arange(num_global_src_nodes, dtype=global_offsets.dtype, device=global_offsets.device)
        // num_src_nodes_per_partition
    )
    return partition_graph_with_id_mapping(
        global_offsets, global_indices, mapping_src_ids_to_ranks, mapping_dst_ids_to_ranks, partition_size,
        partition_rank, device,
    )


def partition_graph_by_coordinate_bbox(
    global_offsets: torch.Tensor,
    global_indices: torch.Tensor,
    src_coordinates: torch.Tensor,
    dst_coordinates: torch.Tensor,
    coordinate_separators_min: List[List[Optional[float]]],
    coordinate_separators_max: List[List[Optional[float]]],
    partition_size: int,
    partition_rank: int,
    device: torch.device,
) -> GraphPartition:
    """Assign nodes to ranks by axis-aligned boxes ``min <= x < max`` (None = unbounded); boxes are
    applied in rank order, so the LAST matching box wins and unmatched points stay on rank 0
    (reference: distributed_graph.py:669-882)."""
    dim = src_coordinates.size(-1)
    if dst_coordinates.size(-1) != dim:
        raise ValueError()
    if len(coordinate_separators_min) != partition_size:
        a, b = len(coordinate_separators_min), partition_size
        raise ValueError(f"Expected len(coordinate_separators_min) == partition_size, but got {a} and {b} respectively")
    if len(coordinate_separators_max) != partition_size:
        a, b = len(coordinate_separators_max), partition_size
        raise ValueError(f"Expected len(coordinate_separators_max) == partition_size, but got {a} and {b} respectively")

    num_global_src_nodes = global_indices.max().item() + 1
    num_global_dst_nodes = global_offsets.size(0) - 1

    mapping_dst_ids_to_ranks = torch.zeros(num_global_dst_nodes, dtype=global_offsets.dtype, device=global_offsets.device)
    mapping_src_ids_to_ranks = torch.zeros(num_global_src_nodes, dtype=global_offsets.dtype, device=global_offsets.device)

    def _assign_ranks(mapping, coordinates):
        for p in range(partition_size):
            mask = torch.ones_like(mapping).to(dtype=torch.bool)
            for d in range(dim):
                min_val, max_val = coordinate_separators_min[p][d], coordinate_separators_max[p][d]
                if min_val is not None:
                    mask = mask & (coordinates[:, d] >= min_val)
                if max_val is not None:
                    mask = mask & (coordinates[:, d] < max_val)
            mapping[mask] = p

    _assign_ranks(mapping_src_ids_to_ranks, src_coordinates)
    _assign_ranks(mapping_dst_ids_to_ranks, dst_coordinates)

    return partition_graph_with_id_mapping(
        global_offsets, global_indices, mapping_src_ids_to_ranks, mapping_dst_ids_to_ranks, partition_size,
        partition_rank, device,
    )


class DistributedGraph:
    """Distributed graph over a process group: partition + the communication primitives that move
    node / edge features between the global, partitioned and local-graph index spaces
    (reference: distributed_graph.py:885-1197)."""

    def __init__(
        self,
        global_offsets: torch.Tensor,
        global_indices: torch.Tensor,
        partition_size: int,
        graph_partition_group_name: str = None,
        graph_partition: Optional[GraphPartition] = None,
    ):
        dist_manager = DistributedManager()
        self.device = dist_manager.device
        self.partition_rank = dist_manager.group_rank(name=graph_partition_group_name)
        self.partition_size = dist_manager.group_size(name=graph_partition_group_name)
        error_msg = f"Passed partition_size does not correspond to size of process_group, got {partition_size} and {self.partition_size} respectively."
        if self.partition_size != partition_size:
            raise AssertionError(error_msg)
        self.process_group = dist_manager.group(name=graph_partition_group_name)

        if graph_partition is None:
            self.graph_partition = partition_graph_nodewise(
                global_offsets, global_indices, self.partition_size, self.partition_rank, self.device,
            )
        else:
            error_msg = f"Passed graph_partition.partition_size does not correspond to size of process_group, got {graph_partition.partition_size} and {self.partition_size} respectively."
            if graph_partition.partition_size != self.partition_size:
                raise AssertionError(error_msg)
            error_msg = f"Passed graph_partition.device does not correspond to device of this rank, got {graph_partition.device} and {self.device} respectively."
            if torch.device(graph_partition.device) != torch.device(self.device):
                raise AssertionError(error_msg)
            self.graph_partition = graph_partition

        gp = self.graph_partition
        send_sizes = gp.sizes[gp.partition_rank]
        recv_sizes = [p[gp.partition_rank] for p in gp.sizes]
        logger.info(
            f"GraphPartition(rank={gp.partition_rank}, num_local_src_nodes={gp.num_local_src_nodes}, "
            f"num_local_dst_nodes={gp.num_local_dst_nodes}, "
            f"num_partitioned_src_nodes={gp.num_src_nodes_in_each_partition[gp.partition_rank]}, "
            f"num_partitioned_dst_nodes={gp.num_dst_nodes_in_each_partition[gp.partition_rank]}, "
            f"send_sizes={send_sizes}, recv_sizes={recv_sizes})"
        )
        if dist.is_available() and dist.is_initialized():
            dist.barrier(self.process_group)

    # ------------------------------------------------------------------ source nodes
    def get_src_node_features_in_partition(self, global_node_features, scatter_features: bool = False,
                                           src_rank: int = 0) -> torch.Tensor:
        if self.graph_partition.matrix_decomp:
            return self.get_dst_node_features_in_partition(global_node_features, scatter_features=scatter_features,
                                                           src_rank=src_rank)
        if scatter_features:
            global_node_features = global_node_features[self.graph_partition.map_concatenated_local_src_ids_to_global]
            return scatter_v(global_node_features, self.graph_partition.num_src_nodes_in_each_partition, dim=0,
                             src=src_rank, group=self.process_group)
        return global_node_features.to(device=self.device)[self.graph_partition.map_partitioned_src_ids_to_global, :]

    def get_src_node_features_in_local_graph(self, partitioned_src_node_features: torch.Tensor) -> torch.Tensor:
        """THE halo exchange: every source row the local edges reference, in local source id order
        (reference: distributed_graph.py:999-1011)."""
        return indexed_all_to_all_v(
            partitioned_src_node_features,
            indices=self.graph_partition.scatter_indices,
            sizes=self.graph_partition.sizes,
            use_fp32=True,
            dim=0,
            group=self.process_group,
        )

    # ------------------------------------------------------------------ destination nodes
    def get_dst_node_features_in_partition(self, global_node_features, scatter_features: bool = False,
                                           src_rank: int = 0) -> torch.Tensor:
        if scatter_features:
            global_node_features = global_node_features.to(device=self.device)[
                self.graph_partition.map_concatenated_local_dst_ids_to_global]
            return scatter_v(global_node_features, self.graph_partition.num_dst_nodes_in_each_partition, dim=0,
                             src=src_rank, group=self.process_group)
        return global_node_features.to(device=self.device)[self.graph_partition.map_partitioned_dst_ids_to_global, :]

    def get_dst_node_features_in_local_graph(self, partitioned_dst_node_features: torch.Tensor) -> torch.Tensor:
        return partitioned_dst_node_features

    # ------------------------------------------------------------------ edges
    def get_edge_features_in_partition(self, global_edge_features, scatter_features: bool = False,
                                       src_rank: int = 0) -> torch.Tensor:
        if scatter_features:
            global_edge_features = global_edge_features[self.graph_partition.map_concatenated_local_edge_ids_to_global]
            return scatter_v(global_edge_features, self.graph_partition.num_indices_in_each_partition, dim=0,
                             src=src_rank, group=self.process_group)
        return global_edge_features.to(device=self.device)[self.graph_partition.map_partitioned_edge_ids_to_global, :]

    def get_edge_features_in_local_graph(self, partitioned_edge_features: torch.Tensor) -> torch.Tensor:
        return partitioned_edge_features

    # ------------------------------------------------------------------ back to the global id space
    def _to_global(self, partitioned, sizes, inverse_map, get_on_all_ranks, dst_rank, what):
        if partitioned.device != torch.device(self.device):
            raise AssertionError(
                f"Passed partitioned_{what}_features.device does not correspond to device of this rank, got "
                f"{partitioned.device} and {self.device} respectively.")
        if not get_on_all_ranks:
            out = gather_v(partitioned, sizes, dim=0, dst=dst_rank, group=self.process_group)
            if self.graph_partition.partition_rank == dst_rank:
                out = out[inverse_map]
            return out
        out = all_gather_v(partitioned, sizes, dim=0, use_fp32=True, group=self.process_group)
        return out[inverse_map]

    def get_global_src_node_features(self, partitioned_node_features, get_on_all_ranks: bool = True,
                                     dst_rank: int = 0) -> torch.Tensor:
        if self.graph_partition.matrix_decomp:
            return self.get_global_dst_node_features(partitioned_node_features, get_on_all_ranks=get_on_all_ranks,
                                                     dst_rank=dst_rank)
        gp = self.graph_partition
        return self._to_global(partitioned_node_features, gp.num_src_nodes_in_each_partition,
                               gp.map_global_src_ids_to_concatenated_local, get_on_all_ranks, dst_rank, "node")

    def get_global_dst_node_features(self, partitioned_node_features, get_on_all_ranks: bool = True,
                                     dst_rank: int = 0) -> torch.Tensor:
        gp = self.graph_partition
        return self._to_global(partitioned_node_features, gp.num_dst_nodes_in_each_partition,
                               gp.map_global_dst_ids_to_concatenated_local, get_on_all_ranks, dst_rank, "node")

    def get_global_edge_features(self, partitioned_edge_features, get_on_all_ranks: bool = True,
                                 dst_rank: int = 0) -> torch.Tensor:
        gp = self.graph_partition
        return self._to_global(partitioned_edge_features, gp.num_indices_in_each_partition,
                               gp.map_global_edge_ids_to_concatenated_local, get_on_all_ranks, dst_rank, "edge")
