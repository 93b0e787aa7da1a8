"""Offline k-hop-halo partitions of one large mesh graph (the XAeroNet-S strategy).

The reference cuts a multi-million-node surface mesh into `num_partitions` pieces with
`dgl.metis_partition(graph, k, extra_cached_hops=halo_hops, reshuffle=True)`
(examples/cfd/external_aerodynamics/xaeronet/surface/preprocessor.py:129-155), trains on one piece at a time with
gradient accumulation and evaluates the loss on the piece's `inner_node`s only (train.py:183-208).  With
`halo_hops >= number of message-passing layers` the inner-node outputs of a piece are exactly those of the full graph.

METIS itself is a third-party library that is absent here; the ASSIGNMENT of nodes to pieces is therefore an input
(`part_id`, from METIS, from `partition_ids_by_slabs` below, or from any coordinate rule).  Everything after the
assignment -- the halo growth along in-edges, the local re-indexing, the per-piece CSC and the id maps back to the
global node / edge tables (DGL's `NID`, `EID`, `inner_node` fields) -- is built here with torch index arithmetic and
runs on whichever device the graph lives on (no host round trip on CUDA).

Halo rule (DGL `partition_graph_with_halo`): hop 0 is the inner set; hop h+1 adds every in-edge of the nodes of hop h
and the sources of those edges that are not in the piece yet.  Nodes of the last hop carry no in-edges.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch
from torch import Tensor


@dataclass
class HaloPartition:
    """One piece: local CSC over [inner nodes (ascending global id) ; halo nodes (ascending global id)]."""

    offsets: Tensor        # [n_local + 1] int64, in-edge segments of the local nodes
    indices: Tensor        # [e_local] int64, LOCAL source id per in-edge
    node_ids: Tensor       # [n_local] int64, global node id per local node           (dgl.NID)
    edge_ids: Tensor       # [e_local] int64, global CSC edge row per local edge      (dgl.EID)
    inner_node: Tensor     # [n_local] bool, True for the nodes this piece owns       (ndata["inner_node"])
    num_inner: int

    @property
    def num_nodes(self) -> int:
        return int(self.node_ids.numel())

    @property
    def num_edges(self) -> int:
        return int(self.edge_ids.numel())

    def to(self, *args, **kwargs) -> "HaloPartition":
        for f in ("offsets", "indices", "node_ids", "edge_ids", "inner_node"):
            setattr(self, f, getattr(self, f).to(*args, **kwargs))
        return self

    def graph(self):
        """The piece as a CuGraphCSC on the device its tensors live on."""
        from .graph import CuGraphCSC

        return CuGraphCSC(self.offsets, self.indices, self.num_nodes, self.num_nodes)


def _expand(start: Tensor, count: Tensor) -> Tensor:
    """Concatenation of the ranges [start[i], start[i] + count[i])."""
    total = int(count.sum())
    if total == 0:
        return torch.empty(0, dtype=torch.int64, device=start.device)
    seg_end = torch.cumsum(count, 0)
    seg = torch.repeat_interleave(torch.arange(count.numel(), device=start.device), count)
    within = torch.arange(total, device=start.device) - (seg_end - count)[seg]
    return start[seg] + within


def partition_with_halo(offsets: Tensor, indices: Tensor, part_id: Tensor, num_partitions: int,
                        halo_hops: int) -> List[HaloPartition]:
    """Cut the CSC graph (`offsets` [N+1], `indices` [E] = source id per in-edge) into `num_partitions` pieces by the
    node assignment `part_id` [N] and grow each by `halo_hops` hops along in-edges."""
    if halo_hops < 0:
        raise ValueError("halo_hops must be >= 0")
    n = offsets.numel() - 1
    if part_id.numel() != n:
        raise ValueError(f"part_id has {part_id.numel()} entries for a graph with {n} nodes")
    if n and (int(part_id.min()) < 0 or int(part_id.max()) >= num_partitions):
        raise ValueError("part_id entries must lie in [0, num_partitions)")
    dev = offsets.device
    offsets = offsets.to(torch.int64)
    indices = indices.to(torch.int64)
    degree = offsets[1:] - offsets[:-1]
    pieces: List[HaloPartition] = []
    for p in range(num_partitions):
        inner = torch.nonzero(part_id == p, as_tuple=False).flatten()
        if inner.numel() == 0:
            raise RuntimeError(f"partition {p} owns no node")  # as the reference's partitioners (distributed_graph.py:250)
        in_piece = torch.zeros(n, dtype=torch.bool, device=dev)
        in_piece[inner] = True
        has_edges = torch.zeros(n, dtype=torch.bool, device=dev)  # nodes whose in-edges belong to the piece
        frontier = inner
        for _ in range(halo_hops):
            if frontier.numel() == 0:
                break
            has_edges[frontier] = True
            srcs = indices[_expand(offsets[frontier], degree[frontier])]
            new = torch.unique(srcs[~in_piece[srcs]])
            in_piece[new] = True
            frontier = new
        halo_mask = in_piece.clone()
        halo_mask[inner] = False
        halo = torch.nonzero(halo_mask, as_tuple=False).flatten()
        node_ids = torch.cat([inner, halo])                    # both ascending
        local_of = torch.full((n,), -1, dtype=torch.int64, device=dev)
        local_of[node_ids] = torch.arange(node_ids.numel(), device=dev)
        local_deg = torch.where(has_edges[node_ids], degree[node_ids], torch.zeros_like(degree[node_ids]))
        edge_ids = _expand(offsets[node_ids], local_deg)       # local CSC order: by local dst, global row ascending
        loc_offsets = torch.zeros(node_ids.numel() + 1, dtype=torch.int64, device=dev)
        loc_offsets[1:] = torch.cumsum(local_deg, 0)
        loc_indices = local_of[indices[edge_ids]]
        inner_flag = torch.zeros(node_ids.numel(), dtype=torch.bool, device=dev)
        inner_flag[: inner.numel()] = True
        pieces.append(HaloPartition(loc_offsets, loc_indices, node_ids, edge_ids, inner_flag, int(inner.numel())))
    return pieces


def partition_ids_by_slabs(coordinates: Tensor, num_partitions: int, axis: int = 0) -> Tensor:
    """A METIS-free node assignment: equal-count slabs along one coordinate axis (ties broken by node id)."""
    n = coordinates.shape[0]
    order = torch.argsort(coordinates[:, axis], stable=True)
    part = torch.empty(n, dtype=torch.int64, device=coordinates.device)
    per = (n + num_partitions - 1) // num_partitions
    part[order] = torch.arange(n, device=coordinates.device) // per
    return part
