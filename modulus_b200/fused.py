"""Fast path of the MeshGraphNet hot loop on the tcgen05 kernels (bf16 storage, hidden 128, ReLU,
LayerNorm, "sum" aggregation, CSC-ordered edges).

Two autograd Functions drive libmgn_b200.so directly, with hand-managed buffers:

  * `FusedProcessorFn`  the L x (MeshEdgeBlock, MeshNodeBlock) loop of MeshGraphNetProcessor
                         (reference: models/meshgraphnet/meshgraphnet.py:353-379)
  * `FusedMLPFn`        a stand-alone MeshGraphMLP: encoders / decoder (meshgraphnet.py:213-216)

Algebra.  The first Linear of both block MLPs is split by input block, the way the reference's own
"concat trick" does for the edge MLP (MeshGraphEdgeMLPSum, mesh_graph_mlp.py:278-458) -- here applied
internally, parameters keep the plain `Linear(3H, H)` / `Linear(2H, H)` layout:

    edge:  z1[e] = efeat[e] W1[:, :H]^T  + P[src[e], 0:H] + P[dst[e], H:2H] + b1
    node:  z1[v] = agg[v]   W1n[:, :H]^T + P[v, 2H:3H] + b1n
    P = nfeat [W1[:, H:2H]; W1[:, 2H:3H]; W1n[:, H:2H]]^T              one node-level GEMM per layer

so the per-edge tensor work is three 128-wide GEMMs, the gathered operands are per-node rows that
stay L2-resident, and in backward the per-edge gradient of the gathered rows is just g_z1, reduced to
nodes by the deterministic CSC / CSR segmented sums BEFORE it meets a weight:

    T = [ csr_sum(g_z1_edge) | csc_sum(g_z1_edge) | g_z1_node ]   [N, 3H]
    g_nfeat += T Wp,        g_Wp = T^T nfeat

Backward recomputes hidden activations inside the kernels; only layer inputs (efeat_l, nfeat_l), the
aggregate and P are kept from the forward pass.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, ops
from ._lib import ACT_IDS, MGN_BF16, call
from .ops import GraphPlan, TC_HIDDEN, _p, _stream

Tensor = torch.Tensor
ENABLED = True  # tests flip this to compare against the generic (unfused) kernels
FUSE_BWD_DST_SUM = True  # the h1 backward kernel also emits the destination sums of g_z1 (its reducer warps have the slack)
KEEP_H1 = True  # the edge forward stores relu(z1); the edge backward starts from it instead of recomputing GEMM1
H = TC_HIDDEN
BF16 = torch.bfloat16


# ----------------------------------------------------------------------------------------
# node-level dense helpers (plain GEMMs on [N, *] tables)
# ----------------------------------------------------------------------------------------
def _node_linear(x: Tensor, w: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """out[M, Nout] = x[M, K] w[Nout, K]^T   (bf16 rows, fp32 weights)."""
    return ops.linear_tc(x, w, out=out)


def _node_wgrad(g: Tensor, x: Tensor) -> Tensor:
    """[Ng, K] fp32 = g[M, Ng]^T x[M, K]."""
    return ops.wgrad_tc(g, x)


# ----------------------------------------------------------------------------------------
# halo exchange of the source-row projections (partitioned graphs)
# ----------------------------------------------------------------------------------------
def remote_only_index(src: Tensor, own_off: int, own_cnt: int, own_rows: Tensor, n_part: int, send_idx: Tensor,
                      send_splits: Sequence[int], rank: int):
    """Pure index maps of the remote-only halo exchange (see HaloContext).

    src        [E]  local source id of every edge, in the rank-ordered local source space of the reference
                    (distributed_graph.py:336-368): [rows owned by rank 0 | rank 1 | ...]
    own_rows   [own_cnt] partition row of the k-th own source (= scatter_indices[rank])
    send_idx   peer-major partition rows this rank sends (cat of scatter_indices), send_splits its split sizes

    Returns (src_ext [E]: own source -> its partition row, halo source -> n_part + slot among the received rows in
    rank order; send_idx_remote: send_idx without the own block)."""
    src = src.long()
    owned = (src >= own_off) & (src < own_off + own_cnt)
    slot = torch.where(src < own_off, src, src - own_cnt)
    own_of_src = own_rows.long()[(src - own_off).clamp(0, max(own_cnt - 1, 0))] if own_cnt > 0 else torch.zeros_like(src)
    src_ext = torch.where(owned, own_of_src, n_part + slot)
    so, sc = int(sum(send_splits[:rank])), int(send_splits[rank])
    send_idx_remote = torch.cat([send_idx[:so], send_idx[so + sc:]])
    return src_ext, send_idx_remote


class HaloContext:
    """Per-graph state of the fused path on a partitioned graph (reference: DistributedGraph
    .get_src_node_features_in_local_graph, distributed_graph.py:999-1011 -> indexed_all_to_all_v).

    What travels is the source projection P[:, 0:H] = nfeat W1[:, H:2H]^T (the reference's concat-trick order:
    per-node products BEFORE the exchange, mesh_graph_mlp.py:396-405).  Edges whose source row is owned by this
    rank ("interior") do not wait for the exchange: with a locality-preserving partition they form one long
    run of CSC edge rows, which is launched while the all-to-all is in flight; the boundary runs follow."""

    def __init__(self, graph, plan: GraphPlan):
        from .distributed import utils as du

        dg = graph.dist_graph
        gp = dg.graph_partition
        self.group = dg.process_group
        rank = gp.partition_rank
        self.xplan = du._halo_plan(gp.scatter_indices, gp.sizes, rank)
        self.n_part = int(gp.num_src_nodes_in_each_partition[rank])  # rows of the partitioned (owned) table
        self.n_src_local = int(gp.num_local_src_nodes)
        own_off = int(sum(gp.sizes[r][rank] for r in range(rank)))
        own_cnt = int(gp.sizes[rank][rank])
        src = plan.src.long()
        owned = (src >= own_off) & (src < own_off + own_cnt)
        E = int(src.numel())
        # longest run of interior edges
        bpos = torch.nonzero(~owned).flatten()
        if bpos.numel() == 0:
            e0, e1 = 0, E
        else:
            edges = torch.cat([bpos.new_tensor([-1]), bpos, bpos.new_tensor([E])])
            gaps = edges[1:] - edges[:-1] - 1
            k = int(torch.argmax(gaps))
            e0, e1 = int(edges[k]) + 1, int(edges[k + 1])
        if e1 - e0 < E // 2:
            e0 = e1 = 0  # no useful interior run: everything waits for the exchange
        self.e0, self.e1 = e0, e1
        # interior edges read their source projection straight from the local P table
        own_rows = gp.scatter_indices[rank].to(device=src.device, dtype=torch.int64)
        self.src_own = None
        if e1 > e0:
            self.src_own = own_rows[src[e0:e1] - own_off].to(torch.int32).contiguous()
        self.halo_rows = self.n_src_local - own_cnt
        # ---- remote-only exchange (used when the backward does not need the exchanged rows again, KEEP_H1): the
        # rank's own rows never travel.  Sources are re-indexed into an extended table [partition rows ; halo rows]:
        # own source k -> its partition row, halo source -> n_part + its slot in the received (rank-ordered) rows.
        xp = self.xplan
        src_ext, send_idx_remote = remote_only_index(src, own_off, own_cnt, own_rows, self.n_part, xp.send_idx,
                                                     xp.send_splits, rank)
        self.send_idx_remote = send_idx_remote.contiguous()
        self.send_splits_r = [0 if r == rank else int(v) for r, v in enumerate(xp.send_splits)]
        self.recv_splits_r = [0 if r == rank else int(v) for r, v in enumerate(xp.recv_splits)]
        self.src_ext = src_ext.to(torch.int32).contiguous()
        self.csrx_offsets, self.csrx_eids = ops._group_by_key(self.src_ext, self.n_part + self.halo_rows)
        # received halo gradients touch only the few partition rows this rank sends out: group them by those rows
        self.sent_rows, inv = torch.unique(self.send_idx_remote.long(), return_inverse=True)
        self.acc_remote = ops._group_by_key(inv.to(torch.int32).contiguous(), int(self.sent_rows.numel()))
        self.remote_only = True  # (the transport, NCCL or the point-to-point stand-in, is the same for both protocols)
        # optional transport of the remote-only protocol: rows stored straight into the peers' memory over NVLink
        # (distributed/peer_halo.py, MGN_HALO_P2P=1); None keeps the NCCL all-to-all
        from .distributed import peer_halo

        self.peer = peer_halo.try_create(self.group, rank, len(self.send_splits_r), self.send_splits_r, self.recv_splits_r, H,
                                         src.device)

    # forward: pack -> all-to-all (in flight) ; returns (work, recv buffer [n_src_local, H], packed keep-alive)
    def start_fwd(self, P: Tensor):
        from .distributed import utils as du

        xp = self.xplan
        packed = ops.gather_rows(P, 0, H, xp.send_idx, xp.send_idx.numel())
        work, recv = du.all_to_all_rows_async(packed, xp.send_splits, xp.recv_splits, group=self.group)
        return work, recv, packed

    # backward: gradients of the received rows travel back and are summed (fixed order) into out[:, col0:col0+H]
    def start_bwd(self, g_rows: Tensor):
        from .distributed import utils as du

        xp = self.xplan
        return du.all_to_all_rows_async(g_rows, xp.recv_splits, xp.send_splits, group=self.group)

    # remote-only variants: only halo rows are packed / sent / accumulated
    def start_fwd_remote(self, P: Tensor):
        from .distributed import utils as du

        if self.peer is not None:  # one launch: gather + peer stores + epoch flags
            from .distributed.peer_halo import PeerWork

            epoch, recv = self.peer.push_fwd(P, self.send_idx_remote)
            return PeerWork(self.peer.wait_fwd, epoch), recv, None
        n = int(self.send_idx_remote.numel())
        packed = ops.gather_rows(P, 0, H, self.send_idx_remote, n) if n > 0 else P.new_empty((0, H))
        work, recv = du.all_to_all_rows_async(packed, self.send_splits_r, self.recv_splits_r, group=self.group)
        return work, recv, packed

    def start_bwd_remote(self, g_halo: Tensor):
        from .distributed import utils as du

        if self.peer is not None:
            from .distributed.peer_halo import PeerWork

            epoch, recv = self.peer.push_bwd(g_halo.contiguous())
            return PeerWork(self.peer.wait_bwd, epoch), recv

        return du.all_to_all_rows_async(g_halo, self.recv_splits_r, self.send_splits_r, group=self.group)

    def finish_bwd_remote(self, work, recv: Tensor, out: Tensor, out_col0: int):
        if work is not None:
            work.wait()
        if recv.shape[0] > 0:  # fixed-order sums per sent row, then one add per (unique) row
            offsets, ids = self.acc_remote
            part = ops.segment_sum(recv, 0, H, offsets, ids, int(self.sent_rows.numel()))
            cols = out[:, out_col0:out_col0 + H]
            cols[self.sent_rows] = (cols[self.sent_rows].float() + part.float()).to(out.dtype)

    def finish_bwd(self, work, recv: Tensor, out: Tensor, out_col0: int):
        if work is not None:
            work.wait()
        offsets, ids = self.xplan.acc_structs(self.n_part)
        ops.segment_sum(recv, 0, H, offsets, ids, self.n_part, out=out, out_col0=out_col0)


# ----------------------------------------------------------------------------------------
# processor
# ----------------------------------------------------------------------------------------
def Ps_saved(halo, remote_only: bool) -> bool:
    """Whether the exchanged source-projection table of a layer is kept for the backward pass."""
    return halo is not None and not remote_only


class FusedProcessorFn(torch.autograd.Function):
    """args: nfeat [N,H] bf16, efeat [E,H] bf16, plan, L, then 16 parameters per layer:
    edge (w1 [H,3H], b1, w2, b2, w3, b3, gamma, beta), node (w1 [H,2H], b1, w2, b2, w3, b3, gamma, beta)."""

    @staticmethod
    def forward(ctx, nfeat: Tensor, efeat: Tensor, plan: GraphPlan, halo, L: int, eps: float, mean: bool, *params: Tensor):
        E, N = plan.n_edges, plan.n_dst
        src, dst = plan.src, plan.dst
        nfeat, efeat = nfeat.contiguous(), efeat.contiguous()
        saved: List[Tensor] = []
        # node-level projection weights of ALL layers in three launches: Wp[l] = [W1e[:, H:2H]; W1e[:, 2H:3H]; W1n[:, H:2H]]
        # ([3H, H]) and its transpose for the backward GEMM (instead of a cat + a transpose-copy per layer and direction)
        w1e_all = torch.stack([params[16 * l] for l in range(L)])          # [L, H, 3H]
        w1n_all = torch.stack([params[16 * l + 8] for l in range(L)])      # [L, H, 2H]
        wp_all = torch.cat([w1e_all[:, :, H:2 * H], w1e_all[:, :, 2 * H:], w1n_all[:, :, H:]], dim=1).contiguous()
        for l in range(L):
            ew, nw = params[16 * l: 16 * l + 8], params[16 * l + 8: 16 * l + 16]
            wp = wp_all[l]  # [3H, H]
            remote_only = halo is not None and halo.remote_only and KEEP_H1
            if remote_only:  # P with room for the halo rows of its source-projection columns behind the partition rows
                P_ext = torch.empty((N + halo.halo_rows, 3 * H), dtype=BF16, device=nfeat.device)
                P = _node_linear(nfeat, wp, out=P_ext[:N])
            else:
                P = _node_linear(nfeat, wp)  # [N, 3H]
            agg = None
            # relu(z1) of the edge MLP, kept for the backward pass (mgn_edge_block_bwd_tc starts from it)
            h1 = torch.empty((E, H), dtype=BF16, device=efeat.device) if KEEP_H1 else None
            if halo is None:  # edge update and destination sums in one pass over the edge rows
                efeat_new, agg = ops.edge_block_fwd_tc(efeat, P, src, dst, plan.csc_offsets, N, ew[0][:, :H], ew[1],
                                                       ew[2], ew[3], ew[4], ew[5], ew[6], ew[7], eps=eps, h1_out=h1)
            else:
                efeat_new = torch.empty_like(efeat)
                agg = torch.empty((N, H), dtype=BF16, device=efeat.device)
                if remote_only:
                    work, recv_r, keep = halo.start_fwd_remote(P)  # halo rows only; own rows never travel
                    Ps = None
                else:
                    work, Ps, keep = halo.start_fwd(P)  # all-to-all of the source projections, in flight
                # three launches over consecutive row ranges share one aggregation record array (row order)
                ranges = [(0, halo.e0), (halo.e0, halo.e1), (halo.e1, E)]
                tiles = [-(-(hi - lo) // 128) for lo, hi in ranges]
                ws = ops.agg_workspace(sum(tiles), efeat.device)

                def run(k, g1, g1_idx):
                    lo, hi = ranges[k]
                    if hi > lo:
                        ops.edge_block_fwd_part_tc(efeat[lo:hi], g1, g1_idx, P, dst[lo:hi], plan.csc_offsets, N,
                                                   ew[0][:, :H], ew[1], ew[2], ew[3], ew[4], ew[5], ew[6], ew[7], eps,
                                                   efeat_new[lo:hi], agg, ws, lo, sum(tiles), sum(tiles[:k]),
                                                   h1_out=None if h1 is None else h1[lo:hi])

                if remote_only:
                    run(1, P_ext, halo.src_ext[halo.e0:halo.e1])  # interior edges (own sources) overlap the exchange
                    if work is not None:
                        work.wait()
                    if halo.halo_rows > 0:
                        P_ext[N:, :H].copy_(recv_r)
                    run(0, P_ext, halo.src_ext[:halo.e0])
                    run(2, P_ext, halo.src_ext[halo.e1:])
                else:
                    run(1, P, halo.src_own)  # interior edges overlap the exchange
                    if work is not None:
                        work.wait()
                    run(0, Ps, src[:halo.e0])
                    run(2, Ps, src[halo.e1:])
                ops.agg_fixup(ws, sum(tiles), agg, N)
                del keep
            if agg is None:
                agg = ops.segment_sum(efeat_new, 0, H, plan.csc_offsets, None, N)
            if mean:  # aggregation="mean" (utils.py:372): rows of the destination sums scaled by 1 / max(in-degree, 1), in place
                ops.gather_rows(agg, 0, H, None, N, out=agg, inv_deg_offsets=plan.csc_offsets)
            h1n = torch.empty((N, H), dtype=BF16, device=nfeat.device) if KEEP_H1 else None
            nfeat_new = ops.node_block_fwd_tc(agg, P, 2 * H, nfeat, nw[0][:, :H], nw[1], nw[2], nw[3], nw[4], nw[5], nw[6],
                                              nw[7], eps=eps, h1_out=h1n)
            # (the [N, 3H] projection table is never kept: the backward from the stored h1 does not read it, and the memory-lean
            #  mode recomputes it from nfeat with one node-level GEMM -- 3 N H b bytes less per layer either way)
            saved += [efeat, nfeat, agg, P.new_empty(0)] + ([Ps] if Ps_saved(halo, remote_only) else []) + \
                     ([h1, h1n] if h1 is not None else [])
            efeat, nfeat = efeat_new, nfeat_new
        ctx.plan, ctx.L, ctx.eps, ctx.halo, ctx.keep_h1, ctx.mean = plan, L, eps, halo, KEEP_H1, mean
        ctx.remote_only = halo is not None and halo.remote_only and KEEP_H1
        ctx.save_for_backward(*saved, *params, wp_all)
        ctx.n_saved = len(saved)
        return nfeat

    @staticmethod
    def backward(ctx, g_n: Tensor):
        plan: GraphPlan = ctx.plan
        L, eps, halo = ctx.L, ctx.eps, ctx.halo
        ns = (5 if Ps_saved(halo, ctx.remote_only) else 4) + (2 if ctx.keep_h1 else 0)
        E, N = plan.n_edges, plan.n_dst
        src, dst = plan.src, plan.dst
        saved = ctx.saved_tensors[:ctx.n_saved]
        params = ctx.saved_tensors[ctx.n_saved:-1]
        wpt_all = ctx.saved_tensors[-1].transpose(1, 2).contiguous()  # [L, H, 3H]: g_n += T Wp as x W^T with W = Wp^T
        dev = g_n.device
        g_n = g_n.contiguous().to(BF16)
        g_e: Optional[Tensor] = None
        grads: List[Optional[Tensor]] = [None] * len(params)
        f32 = dict(dtype=torch.float32, device=dev)
        # (Fusing the destination sums of g_z1 into the edge backward kernel was built and measured: its reducer warps
        #  are already the busiest role and the kernel is at its register limit -- +1.1 ms per layer fused against
        #  0.38 ms for the stand-alone CSC sum, and +7 % on the kernel even with the option off.  Forward fuses them.)
        for l in range(L - 1, -1, -1):
            efeat, nfeat, agg, P = saved[ns * l: ns * l + 4]
            if not ctx.keep_h1:  # memory-lean mode: projections recomputed (bit-identical: same kernel, same inputs)
                P = _node_linear(nfeat, ctx.saved_tensors[-1][l])
            ew, nw = params[16 * l: 16 * l + 8], params[16 * l + 8: 16 * l + 16]
            gew1, gnw1 = torch.empty((H, 3 * H), **f32), torch.empty((H, 2 * H), **f32)
            ge = [gew1] + [torch.empty_like(t, dtype=torch.float32) for t in ew[1:]]
            gn = [gnw1] + [torch.empty_like(t, dtype=torch.float32) for t in nw[1:]]
            T = torch.empty((N, 3 * H), dtype=BF16, device=dev)
            # ---- node block: g_agg = dL/d agg, g_z1 (node) -> T[:, 2H:3H]
            if ctx.keep_h1:  # from the stored h1 of the node MLP (no residual over its layer-1 input: add_gout = 0)
                g_agg, _ = ops.edge_block_bwd_tc(agg, saved[ns * l + ns - 1], g_n, None, None, None, nw[0][:, :H], nw[2],
                                                 nw[3], nw[4], nw[5], nw[6], eps, gnw1[:, :H], gn[1], gn[2], gn[3], gn[4],
                                                 gn[5], gn[6], gn[7], add_gout=False, g_z1_out=T[:, 2 * H:])
            else:
                g_agg, _ = ops.mlp3_bwd_tc(agg, None, None, P, None, 2 * H, None, None, 0, g_n, None, None, N,
                                           nw[0][:, :H], nw[1], nw[2], nw[3], nw[4], nw[5], nw[6], H, eps,
                                           True, False, True, gnw1[:, :H], gn[1], gn[2], gn[3], gn[4], gn[5], gn[6],
                                           gn[7], g_z1_out=T[:, 2 * H:])
            if ctx.mean:  # the node block saw sum / deg: the gradient of the sums is g_agg / deg
                ops.gather_rows(g_agg, 0, H, None, N, out=g_agg, inv_deg_offsets=plan.csc_offsets)
            # ---- edge block: g_out = g_e + g_agg[dst]
            if g_e is None:
                go1, go1_idx, go2, go2_idx = g_agg, dst, None, None
            else:
                go1, go1_idx, go2, go2_idx = g_e, None, g_agg, dst
            g1 = saved[ns * l + 4] if Ps_saved(halo, ctx.remote_only) else P  # exchanged rows / local table
            if ctx.keep_h1:
                h1 = saved[ns * l + ns - 2]
                fuse_dst = FUSE_BWD_DST_SUM  # destination sums T[:, H:2H] out of the same kernel
                g_e, g_z1e = ops.edge_block_bwd_tc(efeat, h1, go1, go1_idx, go2, go2_idx, ew[0][:, :H], ew[2], ew[3], ew[4],
                                                   ew[5], ew[6], eps, gew1[:, :H], ge[1], ge[2], ge[3], ge[4], ge[5],
                                                   ge[6], ge[7], csc_offsets=plan.csc_offsets if fuse_dst else None,
                                                   dst=dst if fuse_dst else None,
                                                   dst_sum_out=T[:, H:2 * H] if fuse_dst else None)
            else:
                g_e, g_z1e = ops.mlp3_bwd_tc(efeat, None, None, g1, src, 0, P, dst, H, go1, go2, go2_idx, E,
                                             ew[0][:, :H], ew[1], ew[2], ew[3], ew[4], ew[5], ew[6], H, eps,
                                             True, True, True, gew1[:, :H], ge[1], ge[2], ge[3], ge[4], ge[5], ge[6],
                                             ge[7], go1_idx=go1_idx)
            # ---- per-node reductions of the gathered-row gradient, then the node-level GEMMs
            have_dst = ctx.keep_h1 and FUSE_BWD_DST_SUM
            if halo is None:
                ops.segment_sum(g_z1e, 0, H, plan.csr_offsets, plan.csr_eids, plan.n_src, out=T, out_col0=0)
                if not have_dst:
                    ops.segment_sum(g_z1e, 0, H, plan.csc_offsets, None, N, out=T, out_col0=H)
            elif ctx.remote_only:
                # own sources accumulate straight into their partition rows; only the halo rows' gradients travel
                ops.segment_sum(g_z1e, 0, H, halo.csrx_offsets[:N + 1], halo.csrx_eids, N, out=T, out_col0=0)
                g_halo = (ops.segment_sum(g_z1e, 0, H, halo.csrx_offsets[N:], halo.csrx_eids, halo.halo_rows)
                          if halo.halo_rows > 0 else g_z1e.new_empty((0, H)))
                work, recv = halo.start_bwd_remote(g_halo)
                if not have_dst:
                    ops.segment_sum(g_z1e, 0, H, plan.csc_offsets, None, N, out=T, out_col0=H)
                halo.finish_bwd_remote(work, recv, T, 0)
            else:
                # gradient of every referenced source row (incl. halo rows) goes back to its owner while the
                # destination-side sum runs; owners accumulate in a fixed order
                s_src = ops.segment_sum(g_z1e, 0, H, plan.csr_offsets, plan.csr_eids, plan.n_src)
                work, recv = halo.start_bwd(s_src)
                ops.segment_sum(g_z1e, 0, H, plan.csc_offsets, None, N, out=T, out_col0=H)
                halo.finish_bwd(work, recv, T, 0)
            g_n = ops.linear_tc(T, wpt_all[l], residual=g_n)  # g_n + T wp
            gwp = _node_wgrad(T, nfeat)  # [3H, H]
            gew1[:, H:2 * H] = gwp[:H]
            gew1[:, 2 * H:] = gwp[H:2 * H]
            gnw1[:, H:] = gwp[2 * H:]
            grads[16 * l: 16 * l + 8] = ge
            grads[16 * l + 8: 16 * l + 16] = gn
        return (g_n, g_e, None, None, None, None, None, *grads)


def _partition_eligible(graph, plan: GraphPlan) -> bool:
    """Partition-level half of `processor_eligible`, decided ONCE per graph and identically on every rank: the fused
    and the generic path issue different collectives, so a rank-local decision could hang the group.  Needs a square
    graph partitioned identically on both id spaces (one owned node table serves sources and destinations) and no
    rank without edges; the rank-local verdict is min-reduced over the partition group."""
    ok = plan.extra.get("fused_partition_ok")
    if ok is None:
        import torch.distributed as dist

        from .distributed import utils as du

        gp = graph.dist_graph.graph_partition
        ok = (not gp.matrix_decomp
              and list(map(int, gp.num_src_nodes_in_each_partition)) == list(map(int, gp.num_dst_nodes_in_each_partition))
              and all(int(n) > 0 for n in gp.num_indices_in_each_partition)
              and torch.equal(gp.map_partitioned_src_ids_to_global, gp.map_partitioned_dst_ids_to_global))
        group = graph.dist_graph.process_group
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group=group) > 1:
            flag = torch.tensor([1.0 if ok else 0.0], device=plan.src.device)
            if du._host_staged(group, flag):
                flag = flag.cpu()
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            ok = bool(flag.item() > 0.5)
        plan.extra["fused_partition_ok"] = ok
    return ok


def processor_eligible(proc, nfeat: Tensor, efeat: Tensor, graph, plan: GraphPlan, dt: torch.dtype) -> bool:
    """Conditions under which the fused tcgen05 path computes exactly what the generic path does."""
    from .models.gnn_layers.mesh_graph_mlp import MeshGraphEdgeMLPConcat, MeshGraphEdgeMLPSum
    from .models.layers.activations import activation_name

    if dt != BF16 or not plan.is_csc_ordered:
        return False
    if getattr(graph, "is_distributed", False):
        if not _partition_eligible(graph, plan):
            return False
    elif plan.n_src != plan.n_dst or plan.n_edges == 0:
        return False
    if nfeat.shape[1] != H or efeat.shape[1] != H or nfeat.shape[0] != plan.n_dst:
        return False
    aggs = set()
    for i, layer in enumerate(proc.processor_layers):
        mlp = layer.edge_mlp if i % 2 == 0 else layer.node_mlp
        if i % 2 == 0:
            # plain Linear(3H, H) first layer, or the reference's "concat trick" layout (lin_efeat / lin_src / lin_dst + bias):
            # the same algebra, re-assembled by MeshGraphEdgeMLPSum._flat_params
            if isinstance(mlp, MeshGraphEdgeMLPSum):
                if mlp.bias is None or not (mlp.efeat_dim == mlp.src_dim == mlp.dst_dim == H):
                    return False
            elif not isinstance(mlp, MeshGraphEdgeMLPConcat):
                return False
        else:
            if layer.aggregation not in ("sum", "mean"):
                return False
            aggs.add(layer.aggregation)
        if mlp.hidden_layers != 2 or mlp.norm_type is None or mlp.hidden_dim != H or mlp.output_dim != H:
            return False
        try:
            if activation_name(mlp.activation_fn) != "relu":
                return False
        except NotImplementedError:
            return False
        if any(q.dtype != torch.float32 or not q.is_cuda for q in mlp.parameters()):
            return False
    return len(aggs) == 1


def processor_forward(proc, nfeat: Tensor, efeat: Tensor, plan: GraphPlan, graph=None) -> Tensor:
    params: List[Tensor] = []
    eps = 1e-5
    for i, layer in enumerate(proc.processor_layers):
        mlp = layer.edge_mlp if i % 2 == 0 else layer.node_mlp
        params += mlp._flat_params()
        eps = mlp._norm().eps
    halo = None
    if graph is not None and getattr(graph, "is_distributed", False):
        halo = plan.extra.get("halo")
        if halo is None:
            halo = plan.extra["halo"] = HaloContext(graph, plan)
    mean = proc.processor_layers[1].aggregation == "mean"
    return FusedProcessorFn.apply(nfeat.to(BF16), efeat.to(BF16), plan, halo, proc.processor_size, eps, mean, *params)


# ----------------------------------------------------------------------------------------
# stand-alone MLP (encoders / decoder)
# ----------------------------------------------------------------------------------------
class FusedMLPFn(torch.autograd.Function):
    """x -> [LayerNorm]( W3 relu(W2 relu(W1 x + b1) + b2) + b3 ).  x is either raw features [M, d <= 64]
    (fp32 or bf16; encoders) or a bf16 [M, 128] table (decoder).  args: x, eps, w1, b1, w2, b2, w3, b3[, gamma, beta]."""

    @staticmethod
    def forward(ctx, x: Tensor, eps: float, *params: Tensor):
        w1, b1, w2, b2, w3, b3 = params[:6]
        gamma, beta = (params[6], params[7]) if len(params) == 8 else (None, None)
        M, d_in = x.shape
        n_out = w3.shape[0]
        small = d_in != H
        if small:
            x = x.contiguous()
            if x.dtype not in (torch.float32, BF16):
                x = x.float()
            out = ops.mlp3_fwd2_tc(None, None, x, None, None, 0, None, None, 0, M, w1, b1, w2, b2, w3, b3, gamma, beta,
                                   eps=eps, n_out=n_out)
        else:
            x = x.contiguous().to(BF16)
            out = ops.mlp3_fwd2_tc(x, None, None, None, None, 0, None, None, 0, M, w1, b1, w2, b2, w3, b3, gamma, beta,
                                   eps=eps, n_out=n_out)
        ctx.save_for_backward(x, *params)
        ctx.eps, ctx.small = eps, small
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        x, *params = ctx.saved_tensors
        w1, b1, w2, b2, w3, b3 = params[:6]
        gamma = params[6] if len(params) == 8 else None
        M = x.shape[0]
        n_out = w3.shape[0]
        dev = g.device
        g = g.contiguous().to(BF16)
        gs = [torch.empty_like(t, dtype=torch.float32) for t in params]
        if gamma is None:
            gg = gb = None
        else:
            gg, gb = gs[6], gs[7]
        need_gx = ctx.needs_input_grad[0]
        g_x = None
        if ctx.small:
            _, g_z1 = ops.mlp3_bwd_tc(None, None, x, None, None, 0, None, None, 0, g, None, None, M, w1, b1, w2, b2, w3,
                                      b3, gamma, n_out, ctx.eps, False, False, need_gx, gs[0], gs[1], gs[2], gs[3],
                                      gs[4], gs[5], gg, gb)
            if need_gx:
                g_x = ops._linear_bwd_data(g_z1, w1).to(x.dtype)
        else:
            g_x, _ = ops.mlp3_bwd_tc(x, None, None, None, None, 0, None, None, 0, g, None, None, M, w1, b1, w2, b2, w3,
                                     b3, gamma, n_out, ctx.eps, need_gx, False, False, gs[0], gs[1], gs[2], gs[3],
                                     gs[4], gs[5], gg, gb)
        return (g_x, None, *gs)


def mlp_eligible(mlp, x: Tensor, dt: torch.dtype) -> bool:
    from .models.layers.activations import activation_name

    if dt != BF16 or mlp.hidden_layers != 2 or mlp.hidden_dim != H or x.dim() != 2 or x.shape[0] == 0:
        return False
    d_in, d_out = mlp.input_dim, mlp.output_dim
    if not (d_in == H or d_in <= 64) or d_out > H or (mlp.norm_type is not None and d_out != H):
        return False
    try:
        if activation_name(mlp.activation_fn) != "relu":
            return False
    except NotImplementedError:
        return False
    return all(q.dtype == torch.float32 and q.is_cuda for q in mlp.parameters())


def mlp_forward(mlp, x: Tensor) -> Tensor:
    nrm = mlp._norm()
    return FusedMLPFn.apply(x, nrm.eps if nrm is not None else 1e-5, *mlp._flat_params())
