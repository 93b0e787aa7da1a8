"""Seeded synthetic meshes / graphs for tests and bench.py (SURVEY 8d workloads).

Everything is returned as a global CSC (offsets int64 [N+1], indices int64 [E] = source id per
in-edge, in-edges of a node sorted by source id) plus coordinates and the reference's edge
features (relative displacement and its norm, datapipes/gnn/vortex_shedding_dataset.py:324-331).
Pure index arithmetic with torch ops; runs on CPU or CUDA (`device=`).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def _csc_from_pairs(src: torch.Tensor, dst: torch.Tensor, n_dst: int, n_src: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sort directed edges by (dst, src), drop duplicates, return CSC."""
    key = dst.to(torch.int64) * n_src + src.to(torch.int64)
    key = torch.unique(key)  # sorted + de-duplicated (to_bidirected semantics)
    dst_s = torch.div(key, n_src, rounding_mode="floor")
    src_s = key - dst_s * n_src
    deg = torch.bincount(dst_s, minlength=n_dst)
    offsets = torch.zeros(n_dst + 1, dtype=torch.int64, device=key.device)
    offsets[1:] = torch.cumsum(deg, 0)
    return offsets, src_s


def _edge_features(coords: torch.Tensor, offsets: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    n = offsets.numel() - 1
    deg = offsets[1:] - offsets[:-1]
    dst = torch.repeat_interleave(torch.arange(n, device=offsets.device), deg)
    disp = coords[indices] - coords[dst]
    norm = disp.norm(dim=1, keepdim=True)
    return torch.cat([disp, norm], dim=1)


def triangle_grid_mesh(nx: int, ny: int, device="cpu") -> Dict:
    """Structured 2-D triangulation of an nx x ny point grid (row-major node ids).  Each quad is
    split along one diagonal; every interior node has degree 6, edges are bidirected: E ~ 6N."""
    i = torch.arange(nx, device=device).view(-1, 1).expand(nx, ny)
    j = torch.arange(ny, device=device).view(1, -1).expand(nx, ny)
    nid = (i * ny + j)
    pairs = []
    pairs.append((nid[:, :-1].reshape(-1), nid[:, 1:].reshape(-1)))      # (i,j)-(i,j+1)
    pairs.append((nid[:-1, :].reshape(-1), nid[1:, :].reshape(-1)))      # (i,j)-(i+1,j)
    pairs.append((nid[:-1, :-1].reshape(-1), nid[1:, 1:].reshape(-1)))   # (i,j)-(i+1,j+1)
    a = torch.cat([p[0] for p in pairs])
    b = torch.cat([p[1] for p in pairs])
    src = torch.cat([a, b])
    dst = torch.cat([b, a])
    n = nx * ny
    offsets, indices = _csc_from_pairs(src, dst, n, n)
    coords = torch.stack([i.reshape(-1).float() / max(nx - 1, 1), j.reshape(-1).float() / max(ny - 1, 1)], dim=1)
    return dict(num_nodes=n, offsets=offsets, indices=indices, coords=coords,
                edge_features=_edge_features(coords, offsets, indices))


def torus_surface_mesh(nu: int, nv: int, device="cpu", R: float = 2.0, r: float = 0.7) -> Dict:
    """Closed 3-D surface (torus) triangulated on a periodic nu x nv grid: every node has degree
    exactly 6 (E = 6N), row-major numbering keeps neighbours close in id space except for the
    wrap-around seam -- the Ahmed-body / DrivAer style workload of SURVEY 8d (C3, C4)."""
    i = torch.arange(nu, device=device).view(-1, 1).expand(nu, nv)
    j = torch.arange(nv, device=device).view(1, -1).expand(nu, nv)
    nid = (i * nv + j).reshape(-1)
    ip = (i + 1) % nu
    jp = (j + 1) % nv
    right = (i * nv + jp).reshape(-1)
    up = (ip * nv + j).reshape(-1)
    diag = (ip * nv + jp).reshape(-1)
    a = torch.cat([nid, nid, nid])
    b = torch.cat([right, up, diag])
    src = torch.cat([a, b])
    dst = torch.cat([b, a])
    n = nu * nv
    offsets, indices = _csc_from_pairs(src, dst, n, n)
    th = i.reshape(-1).float() * (2 * torch.pi / nu)
    ph = j.reshape(-1).float() * (2 * torch.pi / nv)
    coords = torch.stack([(R + r * torch.cos(ph)) * torch.cos(th), (R + r * torch.cos(ph)) * torch.sin(th),
                          r * torch.sin(ph)], dim=1)
    return dict(num_nodes=n, offsets=offsets, indices=indices, coords=coords,
                edge_features=_edge_features(coords, offsets, indices))


def random_graph_csc(num_src: int, num_dst: int, min_degree: int, max_degree: int, seed: int = 0,
                     device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Random bipartite CSC, the recipe of the reference's distributed tests
    (test/models/test_distributed_graph.py:26-51): in-degree uniform in [min,max], sources uniform,
    last index forced so every source id space is fully spanned."""
    g = torch.Generator().manual_seed(seed)
    degree = torch.randint(min_degree, max_degree + 1, (num_dst,), generator=g, dtype=torch.int64)
    offsets = torch.zeros(num_dst + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(degree, 0)
    indices = torch.randint(0, num_src, (int(offsets[-1]),), generator=g, dtype=torch.int64)
    if indices.numel() and int(indices.max()) != num_src - 1:
        indices[-1] = num_src - 1
    return offsets.to(device), indices.to(device)


def power_law_graph_csc(num_nodes: int, num_edges: int, alpha: float = 1.2, seed: int = 0,
                        device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Power-law (Zipf-like) in-degree graph with uniformly random sources: the skewed-degree
    case of the aggregate / gather microbench."""
    g = torch.Generator().manual_seed(seed)
    w = (torch.arange(1, num_nodes + 1, dtype=torch.float64)) ** (-alpha)
    w = w[torch.randperm(num_nodes, generator=g)]
    deg = torch.floor(w / w.sum() * num_edges).to(torch.int64)
    rem = num_edges - int(deg.sum())
    if rem > 0:
        deg[torch.topk(w, min(rem, num_nodes)).indices] += 1
    offsets = torch.zeros(num_nodes + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(deg, 0)
    indices = torch.randint(0, num_nodes, (int(offsets[-1]),), generator=g, dtype=torch.int64)
    return offsets.to(device), indices.to(device)


# ----------------------------------------------------------------------------------------------
# graph construction from mesh cells + edge features on the device
# ----------------------------------------------------------------------------------------------
def graph_from_cells(cells: torch.Tensor, num_nodes: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Triangle (or any polygon) cells [C, k] -> bidirected, de-duplicated graph as a CSC
    (offsets int64 [N+1], indices int64 [E], in-edges of a node sorted by source id).

    Same edge SET as the reference's `cell_to_adj` + `dgl.to_bidirected`
    (datapipes/gnn/vortex_shedding_dataset.py:307-322): edge (cells[i][j] -> cells[i][j+1 mod k]) for every
    cell side, plus the reverse of each, duplicates removed.  Runs wherever `cells` lives; on CUDA the sort /
    unique / scan are torch's device primitives, nothing touches the host."""
    cells = cells.to(torch.int64)
    a = cells.reshape(-1)
    b = torch.roll(cells, shifts=-1, dims=1).reshape(-1)
    src = torch.cat([a, b])
    dst = torch.cat([b, a])
    return _csc_from_pairs(src, dst, num_nodes, num_nodes)


def edge_features(pos: torch.Tensor, src: torch.Tensor, dst: torch.Tensor, mu: torch.Tensor = None,
                  std: torch.Tensor = None) -> torch.Tensor:
    """[E, dim+1] fp32 = (pos[src] - pos[dst], its norm), normalised per column by (x - mu) / std when given:
    `add_edge_features` + `normalize_edge` of the reference (vortex_shedding_dataset.py:324-349) as one CUDA
    pass (mgn_edge_features).  `src` / `dst` are the int32 endpoint arrays of a GraphPlan.  CUDA only."""
    from . import ops

    ops.require_cuda(pos, src, dst, mu, std)
    if pos.dim() != 2 or pos.shape[1] not in (2, 3):
        raise ValueError(f"pos must be [N, 2] or [N, 3], got {tuple(pos.shape)}")
    dim = pos.shape[1]
    for t in (mu, std):
        if t is not None and t.numel() != dim + 1:
            raise AssertionError("Graph edge data must be same size as stats.")  # vortex_shedding_dataset.py:343-348
    pos = pos.contiguous().float()
    src = src.to(torch.int32).contiguous()
    dst = dst.to(torch.int32).contiguous()
    mu = None if mu is None else mu.reshape(-1).contiguous().float()
    std = None if std is None else std.reshape(-1).contiguous().float()
    out = torch.empty((src.numel(), dim + 1), dtype=torch.float32, device=pos.device)
    ops.call("mgn_edge_features", pos.data_ptr(), dim, src.data_ptr(), dst.data_ptr(), src.numel(),
             None if mu is None else mu.data_ptr(), None if std is None else std.data_ptr(), out.data_ptr(),
             torch.cuda.current_stream().cuda_stream)
    return out
