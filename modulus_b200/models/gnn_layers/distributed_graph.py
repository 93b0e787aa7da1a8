"""Graph partitioning and the distributed-graph facade of the MeshGraphNet path.

Public names, argument lists, the `GraphPartition` fields and every index map are those of
the reference (physicsnemo/models/gnn_layers/distributed_graph.py:35-1197); maps are
bit-exact (tests/test_partition.py compares against the reference's known-answer tests and
against partitions produced by the unmodified reference).  The construction is NOT the reference's
(loops over ranks and destination nodes, every rank building all P partitions): one keyed sort / unique
pass over the edge list yields every rank's halo sizes and this rank's own arrays (see `_HaloPairs`), on
the CPU or on the graph's device, so an 8 M-node mesh partitions in seconds instead of minutes.
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


# ----------------------------------------------------------------------------------------
# Partitioning as ONE keyed pass over the edge list
# ----------------------------------------------------------------------------------------
# Every partitioner below reduces to the same question: which (rank, source id) pairs occur, i.e. which
# source rows does each rank's edge set reference.  Edge e belongs to rank own(dst(e)); the pair is packed
# into one integer key  own(dst(e)) * n_src + src(e), and a single sorted `unique` over the E keys gives,
# for all ranks at once, the referenced sources in (rank, source id) order.  The reference's local source
# numbering "[rows owned by rank 0 | rank 1 | ...], each block ascending" (distributed_graph.py:336-368)
# is then the position after a stable regrouping of that list by (rank, owner(source)); the P x P `sizes`
# table is a bincount of (owner, rank); a rank's `scatter_indices` are slices of the same list.  Only the
# calling rank's own arrays are materialised.  All of it is sort / unique / bincount / searchsorted on
# whatever device the graph lives on (no per-node or per-rank Python loops over graph data).
_ERR_SIZES = ("error in graph partition: list containing sizes of exchanged indices does not match the tensor "
              "of indices to be exchanged")


class _HaloPairs:
    """The (rank, source) pairs of a partitioned edge list, grouped by (rank, owner(source), source id)."""

    def __init__(self, edge_rank: torch.Tensor, edge_src: torch.Tensor, src_owner: torch.Tensor, n_src: int, P: int):
        key = edge_rank.to(torch.int64) * n_src + edge_src.to(torch.int64)
        self.keys = torch.unique(key, sorted=True)                     # (rank, source) ascending
        self.rank = torch.div(self.keys, n_src, rounding_mode="floor")
        self.src = self.keys - self.rank * n_src
        self.owner = src_owner.to(torch.int64)[self.src]
        group = self.rank * P + self.owner
        # rows of the list regrouped by (rank, owner); within a group the source ids stay ascending
        self.order = torch.sort(group, stable=True).indices
        counts = torch.bincount(group, minlength=P * P).reshape(P, P)   # [rank, owner]
        self.counts = counts
        self.sizes = counts.t().tolist()                               # sizes[owner][rank]
        per_rank = counts.sum(dim=1)
        self.rank_start = torch.cumsum(per_rank, 0) - per_rank         # first list row of every rank
        self.n_src, self.P = n_src, P

    def local_ids(self, rank: int) -> torch.Tensor:
        """local source id of every list row of `rank`, indexed by its position in the (rank, source)-sorted list"""
        lo = int(self.rank_start[rank])
        hi = lo + int(self.counts[rank].sum())
        rows = self.order[lo:hi]                                        # list rows of this rank in (owner, source) order
        lid = torch.empty(hi - lo, dtype=torch.int64, device=rows.device)
        lid[rows - lo] = torch.arange(hi - lo, dtype=torch.int64, device=rows.device)
        return lid

    def edge_local_src(self, rank: int, edge_src: torch.Tensor) -> torch.Tensor:
        """local source id of each of `rank`'s edges (given their global source ids)"""
        lo = int(self.rank_start[rank])
        pos = torch.searchsorted(self.keys, rank * self.n_src + edge_src.to(torch.int64)) - lo
        return self.local_ids(rank)[pos]

    def sources_of(self, rank: int, owner: int) -> torch.Tensor:
        """global ids (ascending) of the sources owned by `owner` that `rank`'s edges reference"""
        c = self.counts[rank]
        lo = int(self.rank_start[rank]) + int(c[:owner].sum())
        return self.src[self.order[lo: lo + int(c[owner])]]


def _rank_major(mapping: torch.Tensor, P: int):
    """ids grouped by rank (ascending inside a rank), the per-rank counts and the inverse permutation"""
    order = torch.sort(mapping.to(torch.int64), stable=True).indices
    counts = torch.bincount(mapping.to(torch.int64), minlength=P)
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(order.numel(), dtype=torch.int64, device=order.device)
    return order, counts, inverse


def _edge_destinations(global_offsets: torch.Tensor) -> torch.Tensor:
    deg = (global_offsets[1:] - global_offsets[:-1]).to(torch.int64)
    return torch.repeat_interleave(torch.arange(deg.numel(), dtype=torch.int64, device=deg.device), deg)


def partition_graph_with_id_mapping(
    global_offsets: torch.Tensor,
    global_indices: torch.Tensor,
    mapping_src_ids_to_ranks: torch.Tensor,
    mapping_dst_ids_to_ranks: torch.Tensor,
    partition_size: int,
    partition_rank: int,
    device: torch.device,
) -> GraphPartition:
    """Partition a global CSC graph given id -> rank maps of the source and destination id spaces.  Same
    result, field by field and bit for bit, as the reference's distributed_graph.py:154-398 (checked by
    tests/test_partition.py against partitions the unmodified reference produced); built by the keyed pass
    described above instead of the reference's loops over ranks and destination nodes."""
    P, me = partition_size, partition_rank
    gp = GraphPartition(partition_size=P, partition_rank=me, device=device)
    idx_dtype, map_dtype = global_indices.dtype, mapping_src_ids_to_ranks.dtype
    dst_dtype = mapping_dst_ids_to_ranks.dtype

    dst_order, dst_counts, dst_inverse = _rank_major(mapping_dst_ids_to_ranks, P)
    src_order, src_counts, src_inverse = _rank_major(mapping_src_ids_to_ranks, P)
    for r in range(P):  # same complaints, in the same order, as the reference
        if int(dst_counts[r]) == 0:
            raise RuntimeError(f"Aborting partitioning, rank {r} has 0 destination nodes to work on.")
        if int(src_counts[r]) == 0:
            raise RuntimeError(f"Aborting partitioning, rank {r} has 0 source nodes to work on.")
    gp.num_dst_nodes_in_each_partition = dst_counts.tolist()
    gp.num_src_nodes_in_each_partition = src_counts.tolist()
    gp.map_concatenated_local_dst_ids_to_global = dst_order.to(dst_dtype)
    gp.map_global_dst_ids_to_concatenated_local = dst_inverse.to(dst_dtype)
    gp.map_concatenated_local_src_ids_to_global = src_order.to(map_dtype)
    gp.map_global_src_ids_to_concatenated_local = src_inverse.to(map_dtype)
    src_start = torch.cumsum(src_counts, 0) - src_counts
    dst_start = torch.cumsum(dst_counts, 0) - dst_counts
    # row of a source inside its owner's partitioned table
    src_row_in_partition = src_inverse - src_start[mapping_src_ids_to_ranks.to(torch.int64)]

    # edges: CSC order is destination-major, so a stable regrouping by the destination's rank is the
    # concatenation of every rank's CSC ranges
    edge_rank = mapping_dst_ids_to_ranks.to(torch.int64)[_edge_destinations(global_offsets)]
    edge_order, edge_counts, edge_inverse = _rank_major(edge_rank, P)
    gp.map_concatenated_local_edge_ids_to_global = edge_order.to(idx_dtype)
    gp.map_global_edge_ids_to_concatenated_local = edge_inverse.to(idx_dtype)
    gp.num_indices_in_each_partition = edge_counts.tolist()

    n_src = int(mapping_src_ids_to_ranks.numel())
    pairs = _HaloPairs(edge_rank, global_indices, mapping_src_ids_to_ranks, n_src, P)
    for q in range(P):
        for r in range(P):
            gp.sizes[q][r] = pairs.sizes[q][r]
    # rows this rank sends to each peer
    gp.scatter_indices = [src_row_in_partition[pairs.sources_of(r, me)].to(torch.int64).clone() for r in range(P)]

    # this rank's local graph
    e_lo = int(edge_counts[:me].sum())
    my_edges = edge_order[e_lo: e_lo + int(edge_counts[me])]
    my_dst = dst_order[int(dst_start[me]): int(dst_start[me]) + int(dst_counts[me])]
    degree = global_offsets[my_dst + 1] - global_offsets[my_dst]
    gp.local_offsets = torch.cat([torch.zeros(1, dtype=idx_dtype, device=global_indices.device), degree.cumsum(dim=0)])
    gp.local_indices = pairs.edge_local_src(me, global_indices[my_edges]).to(map_dtype)
    gp.num_local_indices = int(my_edges.numel())
    gp.num_local_dst_nodes = int(dst_counts[me])
    gp.num_local_src_nodes = int(pairs.counts[me].sum())
    gp.map_partitioned_src_ids_to_global = src_order[int(src_start[me]): int(src_start[me]) + int(src_counts[me])]
    gp.map_partitioned_dst_ids_to_global = my_dst
    gp.map_partitioned_edge_ids_to_global = my_edges.to(idx_dtype)

    for r in range(P):
        if gp.sizes[me][r] != gp.scatter_indices[r].numel():
            raise AssertionError(_ERR_SIZES)
    return gp.to(device=device)


def partition_graph_with_matrix_decomposition(
    global_offsets: torch.Tensor,
    global_indices: torch.Tensor,
    num_nodes: int,
    partition_book: torch.Tensor,
    partition_size: int,
    partition_rank: int,
    device: torch.device,
) -> GraphPartition:
    """1-D row decomposition of a square adjacency matrix over the contiguous node ranges of
    `partition_book` (reference: distributed_graph.py:401-562).  Here the local source space of a rank is
    simply its referenced sources in ascending global id (owners are contiguous ranges, so that IS the
    rank-grouped order) and the partitioned source table is that same list; the concatenated <-> global maps
    are identities.  Served by the same keyed pass."""
    P, me = partition_size, partition_rank
    gp = GraphPartition(partition_size=P, partition_rank=me, device=device, matrix_decomp=True)
    dtype = global_indices.dtype
    book = partition_book.to(device=global_indices.device, dtype=torch.int64)
    node_lo, node_hi = int(book[me]), int(book[me + 1])
    if node_hi > num_nodes:
        raise ValueError("Invalid node offset and number of local nodes")
    ids = torch.arange(num_nodes, dtype=torch.int64, device=global_indices.device)
    owner = torch.bucketize(ids, book, right=True) - 1
    edge_rank = owner[_edge_destinations(global_offsets)]
    pairs = _HaloPairs(edge_rank, global_indices, owner, num_nodes, P)

    e_lo, e_hi = int(global_offsets[node_lo]), int(global_offsets[node_hi])
    gp.local_offsets = global_offsets[node_lo: node_hi + 1] - global_offsets[node_lo]
    gp.local_indices = pairs.edge_local_src(me, global_indices[e_lo:e_hi])
    gp.num_local_indices = e_hi - e_lo
    gp.num_local_dst_nodes = node_hi - node_lo
    gp.num_local_src_nodes = int(pairs.counts[me].sum())
    gp.map_partitioned_src_ids_to_global = torch.cat([pairs.sources_of(me, q) for q in range(P)]).to(dtype)
    gp.scatter_indices = [(pairs.sources_of(r, me) - node_lo).to(dtype) for r in range(P)]
    edges_per_rank = global_offsets.to(torch.int64)[book[1:]] - global_offsets.to(torch.int64)[book[:-1]]
    gp.num_indices_in_each_partition = edges_per_rank.tolist()
    gp.num_dst_nodes_in_each_partition = (book[1:] - book[:-1]).tolist()
    gp.num_src_nodes_in_each_partition = pairs.counts.sum(dim=1).tolist()
    for q in range(P):
        for r in range(P):
            gp.sizes[q][r] = pairs.sizes[q][r]

    gp.map_partitioned_dst_ids_to_global = torch.arange(node_lo, node_hi, dtype=dtype, device=device)
    gp.map_partitioned_edge_ids_to_global = torch.arange(e_lo, e_hi, dtype=dtype, device=device)
    node_identity = torch.arange(num_nodes, dtype=dtype, device=device)
    edge_identity = torch.arange(global_indices.size(0), dtype=dtype, device=device)
    gp.map_concatenated_local_src_ids_to_global = node_identity
    gp.map_concatenated_local_dst_ids_to_global = node_identity
    gp.map_global_src_ids_to_concatenated_local = node_identity
    gp.map_global_dst_ids_to_concatenated_local = node_identity
    gp.map_concatenated_local_edge_ids_to_global = edge_identity
    gp.map_global_edge_ids_to_concatenated_local = edge_identity

    for r in range(P):
        if gp.sizes[me][r] != gp.scatter_indices[r].numel():
            raise AssertionError(_ERR_SIZES)
    return gp.to(device=device)


def _chunk_owner(n: int, P: int, like: torch.Tensor) -> torch.Tensor:
    per_rank = (n + P - 1) // P
    return torch.arange(n, dtype=like.dtype, device=like.device) // per_rank


def partition_graph_nodewise(
    global_offsets: torch.Tensor,
    global_indices: torch.Tensor,
    partition_size: int,
    partition_rank: int,
    device: torch.device,
    matrix_decomp: bool = False,
) -> GraphPartition:
    """Equal chunks of both id spaces: owner(v) = v // ceil(n / P) (reference: distributed_graph.py:565-666)."""
    n_src = int(global_indices.max()) + 1
    n_dst = int(global_offsets.numel()) - 1
    if matrix_decomp:
        if n_src != n_dst:
            raise ValueError("Must be square adj. matrix (num_src=num_dst) for matrix decomposition")
        per_rank = (n_dst + partition_size - 1) // partition_size
        book = torch.tensor(list(range(0, n_dst, per_rank)) + [n_dst], dtype=global_indices.dtype)
        return partition_graph_with_matrix_decomposition(global_offsets, global_indices, n_dst, book, partition_size,
                                                         partition_rank, device)
    return partition_graph_with_id_mapping(
        global_offsets, global_indices, _chunk_owner(n_src, partition_size, global_offsets),
        _chunk_owner(n_dst, partition_size, global_offsets), partition_size, partition_rank, device)


def _bbox_owner(coordinates: torch.Tensor, lo: List[List[Optional[float]]], hi: List[List[Optional[float]]],
                like: torch.Tensor) -> torch.Tensor:
    """rank of every point: the LAST box (in rank order) with lo <= x < hi on every bounded axis; 0 if none"""
    P, dim = len(lo), coordinates.size(-1)
    inf = float("inf")
    lo_t = torch.tensor([[-inf if v is None else v for v in row] for row in lo], dtype=coordinates.dtype,
                        device=coordinates.device)
    hi_t = torch.tensor([[inf if v is None else v for v in row] for row in hi], dtype=coordinates.dtype,
                        device=coordinates.device)
    x = coordinates[:, None, :dim]
    inside = ((x >= lo_t[None]) & (x < hi_t[None])).all(dim=-1)          # [n, P]
    ranks = torch.arange(1, P + 1, device=coordinates.device)
    return ((inside * ranks[None]).amax(dim=1) - 1).clamp_min(0).to(like.dtype)


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
    """Nodes go to ranks by axis-aligned boxes ``min <= x < max`` (None = unbounded).  The reference assigns
    box after box in rank order (distributed_graph.py:848-869), so overlapping boxes resolve to the last one
    and points in no box stay on rank 0; `_bbox_owner` states that rule directly."""
    dim = src_coordinates.size(-1)
    if dst_coordinates.size(-1) != dim:
        raise ValueError()
    for name, seps in (("coordinate_separators_min", coordinate_separators_min),
                       ("coordinate_separators_max", coordinate_separators_max)):
        if len(seps) != partition_size:
            raise ValueError(f"Expected len({name}) == partition_size, but got {len(seps)} and {partition_size} respectively")
    n_src = int(global_indices.max()) + 1
    n_dst = int(global_offsets.numel()) - 1
    return partition_graph_with_id_mapping(
        global_offsets, global_indices,
        _bbox_owner(src_coordinates[:n_src], coordinate_separators_min, coordinate_separators_max, global_offsets),
        _bbox_owner(dst_coordinates[:n_dst], coordinate_separators_min, coordinate_separators_max, global_offsets),
        partition_size, partition_rank, device)


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
