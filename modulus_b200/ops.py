"""Autograd-aware operators over libmgn_b200.so.

Three layers, bottom-up:
  * `GraphPlan`        int32 CSC/CSR structures precomputed once per graph on the device
  * raw op wrappers    tensors -> pointers -> C ABI (allocation of outputs/workspaces only)
  * autograd Functions concat_efeat / sum_efeat / aggregate_and_concat / fused MLP

Everything requires CUDA tensors; there is no CPU path (reference semantics are restated for
tests in oracle/, which this module never imports).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_IDS, MGN_BF16, MGN_F32, call

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return MGN_F32
    if t.dtype == torch.bfloat16:
        return MGN_BF16
    raise TypeError(f"modulus_b200 kernels support float32 and bfloat16 features, got {t.dtype}")


def _p(t: Optional[Tensor]):
    """Raw pointer of a tensor argument.  The library launches on the CURRENT device and on its current stream
    (`_stream()`), so a tensor that lives elsewhere would be touched from the wrong context with no stream ordering:
    that is refused here, at the one place every pointer passes through."""
    if t is None:
        return None
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise RuntimeError(
            f"modulus_b200: tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
            "run the model under `with torch.cuda.device(tensor.device):` or call torch.cuda.set_device first")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "modulus_b200: the MeshGraphNet operators run on CUDA (sm_100a) only; got a "
                f"{t.device} tensor and there is no CPU fallback"
            )


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _f32(t: Tensor) -> Tensor:
    """Parameters are consumed as fp32 in place; anything else is a usage error we surface."""
    if t.dtype != torch.float32:
        raise TypeError(f"modulus_b200: parameters must be float32 (got {t.dtype})")
    return _c(t)


# ----------------------------------------------------------------------------------------
# graph plan
# ----------------------------------------------------------------------------------------
def _group_by_key(keys: Tensor, n_keys: int) -> Tuple[Tensor, Tensor]:
    n = keys.numel()
    dev = keys.device
    offsets = torch.empty(n_keys + 1, dtype=torch.int32, device=dev)
    ids = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    ws_bytes = _lib.load().mgn_group_by_key_workspace_bytes(n_keys)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("mgn_group_by_key", _p(keys), n, n_keys, _p(offsets), _p(ids), _p(ws), ws_bytes, _stream())
    return offsets, ids[:n]


class GraphPlan:
    """Device-side index structures for one (bipartite) graph.

    Edge rows of every edge-feature table follow `src`/`dst` order.  For a CuGraphCSC that order
    is the CSC order (`csc_eids is None`, destination segments are contiguous row ranges); a
    graph that arrives as COO (DGL edge-id order) keeps its order and gets `csc_eids`.

      src, dst        [E] int32   endpoints per edge row          (gnn_layers/graph.py:462-471)
      csc_offsets     [n_dst+1]   in-edge segments by destination (CuGraphCSC.offsets)
      csr_offsets     [n_src+1]   out-edge segments by source
      csr_eids        [E]         edge rows grouped by source, ascending (backward scatter)
    """

    def __init__(self, src, dst, csc_offsets, csc_eids, csr_offsets, csr_eids, n_src, n_dst):
        self.src, self.dst = src, dst
        self.csc_offsets, self.csc_eids = csc_offsets, csc_eids
        self.csr_offsets, self.csr_eids = csr_offsets, csr_eids
        self.n_src, self.n_dst = int(n_src), int(n_dst)
        self.n_edges = int(src.numel())
        self.device = src.device
        self.extra = {}  # per-plan caches of the fused kernels (tile schedules, ...)

    @property
    def is_csc_ordered(self) -> bool:
        return self.csc_eids is None

    @staticmethod
    def from_csc(offsets: Tensor, indices: Tensor, n_src: int, n_dst: int) -> "GraphPlan":
        require_cuda(offsets, indices)
        if offsets.numel() != n_dst + 1:
            raise ValueError(f"offsets has {offsets.numel()} entries, expected num_dst_nodes+1 = {n_dst + 1}")
        E = int(indices.numel())
        if E >= 2**31:
            raise ValueError("graphs with >= 2^31 edges per rank are not supported")
        off32 = _c(offsets.to(torch.int32))
        src = _c(indices.to(torch.int32))
        dev = off32.device
        dst = torch.empty(E, dtype=torch.int32, device=dev)
        call("mgn_expand_offsets", _p(off32), n_dst, _p(dst), _stream())
        csr_offsets, csr_eids = _group_by_key(src, n_src)
        return GraphPlan(src, dst, off32, None, csr_offsets, csr_eids, n_src, n_dst)

    @staticmethod
    def from_coo(src: Tensor, dst: Tensor, n_src: int, n_dst: int) -> "GraphPlan":
        require_cuda(src, dst)
        src = _c(src.to(torch.int32))
        dst = _c(dst.to(torch.int32))
        csc_offsets, csc_eids = _group_by_key(dst, n_dst)
        csr_offsets, csr_eids = _group_by_key(src, n_src)
        return GraphPlan(src, dst, csc_offsets, csc_eids, csr_offsets, csr_eids, n_src, n_dst)


# ----------------------------------------------------------------------------------------
# raw wrappers
# ----------------------------------------------------------------------------------------
LONG_SEGMENT = 64  # rows; csrc/mgn_gather.cu: kLongSeg
_max_seg_cache: dict = {}


def _has_long_segments(offsets: Tensor) -> bool:
    """Whether an offsets array holds a segment longer than LONG_SEGMENT rows (hub nodes).  One host read-back per
    offsets tensor (graph plans are built once); a stale entry only costs speed, both kernels are exact."""
    key = (offsets.data_ptr(), offsets.numel())
    hit = _max_seg_cache.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            return True
        hit = bool(offsets.numel() > 1 and int((offsets[1:] - offsets[:-1]).max()) > LONG_SEGMENT)
        if len(_max_seg_cache) > 256:
            _max_seg_cache.clear()
        _max_seg_cache[key] = hit
    return hit


def segment_sum(inp: Tensor, in_col0: int, D: int, offsets: Tensor, eids: Optional[Tensor], n_seg: int,
                out: Optional[Tensor] = None, out_col0: int = 0, mean: bool = False,
                accumulate: bool = False) -> Tensor:
    if out is None:
        out = torch.empty((n_seg, D), dtype=inp.dtype, device=inp.device)
    ld_in = inp.stride(0) if inp.dim() == 2 else D
    if _has_long_segments(offsets):  # skewed degrees: hubs are split across CTAs (mgn_segment_sum_balanced)
        n_rows = inp.shape[0]
        nbytes = _lib.load().mgn_segment_sum_workspace_bytes(n_rows, D)
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=inp.device)
        call("mgn_segment_sum_balanced", _dt(inp), _p(inp), ld_in, in_col0, D, _p(offsets), _p(eids), n_seg, _p(out),
             out.stride(0), out_col0, int(mean), int(accumulate), n_rows, _p(ws), nbytes, _stream())
        return out
    call("mgn_segment_sum", _dt(inp), _p(inp), ld_in, in_col0, D,
         _p(offsets), _p(eids), n_seg, _p(out), out.stride(0), out_col0, int(mean), int(accumulate), _stream())
    return out


def gather_rows(inp: Tensor, in_col0: int, D: int, idx: Optional[Tensor], n_rows: int,
                out: Optional[Tensor] = None, out_col0: int = 0,
                inv_deg_offsets: Optional[Tensor] = None) -> Tensor:
    if out is None:
        out = torch.empty((n_rows, D), dtype=inp.dtype, device=inp.device)
    call("mgn_gather_rows", _dt(inp), _p(inp), inp.stride(0), in_col0, D, _p(idx), n_rows, _p(out),
         out.stride(0), out_col0, _p(inv_deg_offsets), _stream())
    return out


# ----------------------------------------------------------------------------------------
# operator seam
# ----------------------------------------------------------------------------------------
class ConcatEfeatFn(torch.autograd.Function):
    """concat_efeat (gnn_layers/utils.py:151-229): fwd two-sided gather + concat; bwd slice copy,
    in-segment CSC sum for the dst rows and CSR (transposed) sum for the src rows."""

    @staticmethod
    def forward(ctx, efeat, src_feat, dst_feat, plan: GraphPlan):
        require_cuda(efeat, src_feat, dst_feat)
        efeat, src_feat, dst_feat = _c(efeat), _c(src_feat), _c(dst_feat)
        E = plan.n_edges
        De, Ds, Dd = efeat.shape[1], src_feat.shape[1], dst_feat.shape[1]
        if efeat.shape[0] != E or src_feat.shape[0] < plan.n_src or dst_feat.shape[0] < plan.n_dst:
            raise ValueError(
                f"concat_efeat: feature rows ({efeat.shape[0]}, {src_feat.shape[0]}, {dst_feat.shape[0]}) do not "
                f"match the graph (E={E}, n_src={plan.n_src}, n_dst={plan.n_dst})")
        if not (efeat.dtype == src_feat.dtype == dst_feat.dtype):
            raise TypeError("concat_efeat: all feature tables must share one dtype")
        out = torch.empty((E, De + Ds + Dd), dtype=efeat.dtype, device=efeat.device)
        call("mgn_concat_efeat_fwd", _dt(efeat), _p(efeat), De, _p(src_feat), Ds, _p(dst_feat), Dd,
             _p(plan.src), _p(plan.dst), E, _p(out), _stream())
        ctx.plan = plan
        ctx.dims = (De, Ds, Dd, src_feat.shape[0], dst_feat.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        plan: GraphPlan = ctx.plan
        De, Ds, Dd, ns, nd = ctx.dims
        g = _c(g)
        ge = gs = gd = None
        if ctx.needs_input_grad[0]:
            ge = gather_rows(g, 0, De, None, plan.n_edges)
        if ctx.needs_input_grad[1]:
            gs = g.new_zeros((ns, Ds)) if ns != plan.n_src else None
            gs = segment_sum(g, De, Ds, plan.csr_offsets, plan.csr_eids, plan.n_src, out=gs)
        if ctx.needs_input_grad[2]:
            gd = g.new_zeros((nd, Dd)) if nd != plan.n_dst else None
            gd = segment_sum(g, De + Ds, Dd, plan.csc_offsets, plan.csc_eids, plan.n_dst, out=gd)
        return ge, gs, gd, None


class SumEfeatFn(torch.autograd.Function):
    """sum_efeat (gnn_layers/utils.py:232-334)."""

    @staticmethod
    def forward(ctx, efeat, src_feat, dst_feat, plan: GraphPlan):
        require_cuda(efeat, src_feat, dst_feat)
        efeat, src_feat, dst_feat = _c(efeat), _c(src_feat), _c(dst_feat)
        E, D = plan.n_edges, efeat.shape[1]
        if efeat.shape[0] != E or src_feat.shape[1] != D or dst_feat.shape[1] != D:
            raise ValueError("sum_efeat: shape mismatch")
        if src_feat.shape[0] < plan.n_src or dst_feat.shape[0] < plan.n_dst:
            raise ValueError(
                f"sum_efeat: node tables have ({src_feat.shape[0]}, {dst_feat.shape[0]}) rows, the graph needs "
                f"(n_src={plan.n_src}, n_dst={plan.n_dst})")
        if not (efeat.dtype == src_feat.dtype == dst_feat.dtype):
            raise TypeError("sum_efeat: all feature tables must share one dtype")
        out = torch.empty_like(efeat)
        call("mgn_sum_efeat_fwd", _dt(efeat), _p(efeat), _p(src_feat), _p(dst_feat), D, _p(plan.src), _p(plan.dst),
             E, _p(out), _stream())
        ctx.plan = plan
        ctx.dims = (D, src_feat.shape[0], dst_feat.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        plan: GraphPlan = ctx.plan
        D, ns, nd = ctx.dims
        g = _c(g)
        gs = gd = None
        if ctx.needs_input_grad[1]:
            gs = g.new_zeros((ns, D)) if ns != plan.n_src else None
            gs = segment_sum(g, 0, D, plan.csr_offsets, plan.csr_eids, plan.n_src, out=gs)
        if ctx.needs_input_grad[2]:
            gd = g.new_zeros((nd, D)) if nd != plan.n_dst else None
            gd = segment_sum(g, 0, D, plan.csc_offsets, plan.csc_eids, plan.n_dst, out=gd)
        return (g if ctx.needs_input_grad[0] else None), gs, gd, None


class AggConcatFn(torch.autograd.Function):
    """aggregate_and_concat (gnn_layers/utils.py:337-427): deterministic CSC segmented sum/mean of
    edge rows by destination, concatenated with the destination rows."""

    @staticmethod
    def forward(ctx, efeat, nfeat, plan: GraphPlan, mean: bool):
        require_cuda(efeat, nfeat)
        efeat, nfeat = _c(efeat), _c(nfeat)
        De, Dn = efeat.shape[1], nfeat.shape[1]
        if efeat.shape[0] != plan.n_edges or nfeat.shape[0] != plan.n_dst:
            raise ValueError(
                f"aggregate_and_concat: got {efeat.shape[0]} edge rows / {nfeat.shape[0]} node rows for a graph "
                f"with E={plan.n_edges}, n_dst={plan.n_dst}")
        if efeat.dtype != nfeat.dtype:
            raise TypeError("aggregate_and_concat: efeat and nfeat must share one dtype")
        out = torch.empty((plan.n_dst, De + Dn), dtype=efeat.dtype, device=efeat.device)
        segment_sum(efeat, 0, De, plan.csc_offsets, plan.csc_eids, plan.n_dst, out=out, out_col0=0, mean=mean)
        gather_rows(nfeat, 0, Dn, None, plan.n_dst, out=out, out_col0=De)
        ctx.plan, ctx.mean, ctx.dims = plan, mean, (De, Dn)
        return out

    @staticmethod
    def backward(ctx, g):
        plan: GraphPlan = ctx.plan
        De, Dn = ctx.dims
        g = _c(g)
        ge = gn = None
        if ctx.needs_input_grad[0]:
            ge = gather_rows(g, 0, De, plan.dst, plan.n_edges,
                             inv_deg_offsets=plan.csc_offsets if ctx.mean else None)
        if ctx.needs_input_grad[1]:
            gn = gather_rows(g, De, Dn, None, plan.n_dst)
        return ge, gn, None, None


# ----------------------------------------------------------------------------------------
# MeshGraphMLP on the fp32-accurate SIMT kernels (any width, fp32 or bf16 activations)
# ----------------------------------------------------------------------------------------
WIDE_TC = True  # bf16 MLPs whose widths are multiples of 128 (64 for the inner dimension) run their GEMMs on tcgen05


def _wide_ok(x: Tensor, inner: int, width: int) -> bool:
    return (WIDE_TC and x.dtype == torch.bfloat16 and inner % 64 == 0 and width % TC_HIDDEN == 0 and x.shape[0] > 0
            and x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0)


def gemm_bf16_tc(x: Tensor, w: Tensor, bias: Optional[Tensor], act: int = 0, transpose_w: bool = False) -> Tensor:
    """act(x W^T + bias) (transpose_w: x W) on the K-looped tensor-core GEMM (include/mgn_b200.h: mgn_gemm_bf16_tc); `w` is
    the fp32 nn.Linear weight, converted to a bf16 image per call."""
    rows, cols = w.shape
    N, K = (cols, rows) if transpose_w else (rows, cols)
    wb = torch.empty((N, K), dtype=torch.bfloat16, device=w.device)
    call("mgn_cast_weight_bf16", _p(w), rows, cols, w.stride(0), _p(wb), int(transpose_w), _stream())
    out = torch.empty((x.shape[0], N), dtype=torch.bfloat16, device=x.device)
    call("mgn_gemm_bf16_tc", _p(x), x.stride(0), x.shape[0], K, _p(wb), K, N, _p(bias), act, _p(out), N,
         _p(tc_status(x.device)), _stream())
    return out


def linear_f32_tc(x: Tensor, w: Tensor, bias: Optional[Tensor], act: int = 0, transpose_w: bool = False) -> Tensor:
    """fp32 act(x W^T + bias) (transpose_w: x W, the data gradient of the same Linear) on the tensor cores with fp32
    accuracy: 3 x TF32 operand split on `tcgen05.mma kind::tf32` (include/mgn_b200.h: mgn_split_weight_tf32 +
    mgn_linear_f32_tc; csrc/mgn_gemm_f32_tc.cu).  K % 32 == 0, N % 128 == 0, act none / relu.  A building block this round:
    the fp32 model path still calls the exact-fp32 SIMT kernels."""
    require_cuda(x, w, bias)
    if x.dtype != torch.float32 or w.dtype != torch.float32 or (bias is not None and bias.dtype != torch.float32):
        raise TypeError("linear_f32_tc: float32 tensors only")
    rows, cols = w.shape
    N, K = (cols, rows) if transpose_w else (rows, cols)
    if x.dim() != 2 or x.shape[1] != K:
        raise ValueError(f"linear_f32_tc: input has {tuple(x.shape)} but the weight expects K = {K}")
    x, w = _c(x), _c(w)
    ws = torch.empty((2 * N, K), dtype=torch.float32, device=w.device)
    call("mgn_split_weight_tf32", _p(w), rows, cols, w.stride(0), _p(ws), int(transpose_w), _stream())
    out = torch.empty((x.shape[0], N), dtype=torch.float32, device=x.device)
    call("mgn_linear_f32_tc", _p(x), x.stride(0), x.shape[0], K, _p(ws), N, _p(bias), act, _p(out), N,
         _p(tc_status(x.device)), _stream())
    return out


def _linear_fwd(x: Tensor, w: Tensor, b: Optional[Tensor], act: int, want_pre: bool):
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"linear: input has {K} features but the weight expects {w.shape[1]}")
    if _wide_ok(x, K, N):
        if not want_pre and act in (ACT_IDS[None], ACT_IDS["relu"]):
            return gemm_bf16_tc(x, w, b, act), None
        pre = gemm_bf16_tc(x, w, b, ACT_IDS[None])  # other activations: one elementwise pass over the stored pre-activation
        h = torch.empty_like(pre)
        call("mgn_act_fwd", _dt(pre), _p(pre), act, _p(h), pre.numel(), _stream())
        return h, (pre if want_pre else None)
    h = torch.empty((M, N), dtype=x.dtype, device=x.device)
    pre = torch.empty_like(h) if want_pre else None
    call("mgn_linear_fwd", _dt(x), _p(x), x.stride(0), M, K, _p(w), _p(b), N, act, _p(pre), _p(h), N, _stream())
    return h, pre


_ONES = {}


def _ones_rows(M: int, device) -> Tensor:
    """[>= M, 128] bf16 ones (bias gradient = g_y^T 1 on the tensor cores); grown on demand, one per device"""
    key = torch.device(device).index or 0
    t = _ONES.get(key)
    if t is None or t.shape[0] < M:
        t = _ONES[key] = torch.ones((max(M, 1), TC_HIDDEN), dtype=torch.bfloat16, device=device)
    return t


def _wgrad_wide(g_y: Tensor, x: Tensor, N: int, K: int, want_bias: bool):
    """g_w[N, K] = g_y^T x and g_b = g_y^T 1 from mgn_wgrad_tc blocks: 128 columns of x against up to 384 columns of g_y per
    launch, written straight into the [N, K] result (fp32, per-CTA partials summed in a fixed order)."""
    M = x.shape[0]
    dev = x.device
    g_w = torch.empty((N, K), dtype=torch.float32, device=dev)
    g_b = None
    lib = _lib.load()
    st = tc_status(dev)
    ones = _ones_rows(M, dev) if want_bias else None
    gb_blk = torch.empty((N, TC_HIDDEN), dtype=torch.float32, device=dev) if want_bias else None
    for n0 in range(0, N, 3 * TC_HIDDEN):
        jb = min(3, (N - n0) // TC_HIDDEN)
        nbytes = lib.mgn_wgrad_tc_workspace_bytes(M, jb)
        ws = _ws(nbytes, dev)
        gp = g_y.data_ptr() + 2 * n0
        for k0 in range(0, K, TC_HIDDEN):
            call("mgn_wgrad_tc", gp, g_y.stride(0), jb, x.data_ptr() + 2 * k0, x.stride(0), M,
                 g_w.data_ptr() + 4 * (n0 * K + k0), K, _p(ws), nbytes, _p(st), _stream())
        if want_bias:
            call("mgn_wgrad_tc", gp, g_y.stride(0), jb, _p(ones), ones.stride(0), M, gb_blk.data_ptr() + 4 * n0 * TC_HIDDEN,
                 TC_HIDDEN, _p(ws), nbytes, _p(st), _stream())
    if want_bias:
        g_b = gb_blk[:, 0].contiguous()
    return g_w, g_b


def _linear_bwd_weight(g_y: Tensor, x: Tensor, N: int, K: int, want_bias: bool):
    M = x.shape[0]
    if (WIDE_TC and g_y.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and N % TC_HIDDEN == 0 and K % TC_HIDDEN == 0
            and M > 0 and g_y.stride(1) == 1 and x.stride(1) == 1 and g_y.stride(0) % 8 == 0 and x.stride(0) % 8 == 0
            and g_y.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0):
        _p(g_y), _p(x)  # device guard
        return _wgrad_wide(g_y, x, N, K, want_bias)
    g_w = torch.empty((N, K), dtype=torch.float32, device=x.device)
    g_b = torch.empty((N,), dtype=torch.float32, device=x.device) if want_bias else None
    ws_bytes = _lib.load().mgn_linear_bwd_weight_workspace_bytes(M, N, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    call("mgn_linear_bwd_weight", _dt(x), _p(g_y), _p(x), x.stride(0), M, N, K, _p(g_w), _p(g_b), _p(ws), ws_bytes,
         _stream())
    return g_w, g_b


def _linear_bwd_data(g_y: Tensor, w: Tensor) -> Tensor:
    M, N = g_y.shape
    K = w.shape[1]
    if _wide_ok(g_y, N, K):
        return gemm_bf16_tc(g_y, w, None, 0, transpose_w=True)
    g_x = torch.empty((M, K), dtype=g_y.dtype, device=g_y.device)
    call("mgn_linear_bwd_data", _dt(g_y), _p(g_y), M, N, _p(w), K, _p(g_x), K, _stream())
    return g_x


class MLPFn(torch.autograd.Function):
    """MeshGraphMLP.forward (mesh_graph_mlp.py:142-203) + optional residual add of the blocks
    (mesh_edge_block.py:95, mesh_node_block.py:91):

        out = [LayerNorm]( Linear_L( act( ... act( Linear_0(x) ) ) ) ) [+ residual]

    args: x, residual|None, act_id, n_linear, has_norm, has_bias, eps, *params
          params = w0, b0, ..., w_{L}, b_{L} [, gamma, beta]   (fp32, read in place)
    """

    @staticmethod
    def forward(ctx, x, residual, act: int, n_linear: int, has_norm: bool, eps: float, *params):
        require_cuda(x, residual, *params)
        x = _c(x)
        ws = [_f32(params[2 * i]) for i in range(n_linear)]
        bs = [None if params[2 * i + 1] is None else _f32(params[2 * i + 1]) for i in range(n_linear)]
        gamma = beta = None
        if has_norm:
            gamma, beta = _f32(params[2 * n_linear]), _f32(params[2 * n_linear + 1])
        need_pre = act not in (ACT_IDS["relu"], ACT_IDS[None])
        inputs: List[Tensor] = [x]
        pres: List[Optional[Tensor]] = []
        h = x
        for i in range(n_linear - 1):
            h, pre = _linear_fwd(h, ws[i], bs[i], act, need_pre)
            inputs.append(h)
            pres.append(pre)
        y, _ = _linear_fwd(h, ws[-1], bs[-1], ACT_IDS[None], False)
        mean = rstd = None
        if has_norm:
            M, D = y.shape
            out = torch.empty_like(y)
            mean = torch.empty(M, dtype=torch.float32, device=y.device)
            rstd = torch.empty(M, dtype=torch.float32, device=y.device)
            res = None if residual is None else _c(residual)
            call("mgn_layernorm_fwd", _dt(y), _p(y), M, D, _p(gamma), _p(beta), eps, _p(res), _p(out), _p(mean),
                 _p(rstd), _stream())
        elif residual is not None:
            out = torch.empty_like(y)
            call("mgn_add", _dt(y), _p(y), _p(_c(residual)), _p(out), y.numel(), _stream())
        else:
            out = y
        ctx.cfg = (act, n_linear, has_norm, need_pre, residual is not None)
        # everything the backward reads goes through save_for_backward: `y` IS the output when there is neither a
        # LayerNorm nor a residual, and an output kept as a plain ctx attribute forms a reference cycle with its
        # own grad_fn (the iteration's autograd graph -- AccumulateGrad nodes included -- would outlive the step)
        saved: List[Tensor] = []

        def keep(t: Optional[Tensor]) -> int:
            if t is None:
                return -1
            saved.append(t)
            return len(saved) - 1

        ctx.slots = ([keep(t) for t in ws], [keep(t) for t in bs], keep(gamma), [keep(t) for t in inputs],
                     [keep(t) for t in pres], keep(y), keep(mean), keep(rstd))
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    def backward(ctx, g):
        act, n_linear, has_norm, need_pre, has_res = ctx.cfg
        saved = ctx.saved_tensors

        def get(i: int) -> Optional[Tensor]:
            return None if i < 0 else saved[i]

        s_ws, s_bs, s_gamma, s_inputs, s_pres, s_y, s_mean, s_rstd = ctx.slots
        ws, bs, gamma = [get(i) for i in s_ws], [get(i) for i in s_bs], get(s_gamma)
        inputs, pres = [get(i) for i in s_inputs], [get(i) for i in s_pres]
        y, mean, rstd = get(s_y), get(s_mean), get(s_rstd)
        g = _c(g)
        grads: List[Optional[Tensor]] = [None] * (2 * n_linear + (2 if has_norm else 0))
        g_res = g if (has_res and ctx.needs_input_grad[1]) else None
        if has_norm:
            M, D = y.shape
            g_y = torch.empty_like(y)
            g_gamma = torch.empty(D, dtype=torch.float32, device=y.device)
            g_beta = torch.empty(D, dtype=torch.float32, device=y.device)
            ws_bytes = _lib.load().mgn_layernorm_bwd_workspace_bytes(M, D)
            wsb = torch.empty(ws_bytes, dtype=torch.uint8, device=y.device)
            call("mgn_layernorm_bwd", _dt(y), _p(g), _p(y), _p(mean), _p(rstd), _p(gamma), M, D, _p(g_y),
                 _p(g_gamma), _p(g_beta), _p(wsb), ws_bytes, _stream())
            grads[2 * n_linear], grads[2 * n_linear + 1] = g_gamma, g_beta
        else:
            g_y = g
        g_x = None
        for i in range(n_linear - 1, -1, -1):
            N, K = ws[i].shape
            g_w, g_b = _linear_bwd_weight(g_y, inputs[i], N, K, bs[i] is not None)
            grads[2 * i], grads[2 * i + 1] = g_w, g_b
            if i > 0 or ctx.needs_input_grad[0]:
                g_in = _linear_bwd_data(g_y, ws[i])
                if i > 0:
                    ref = pres[i - 1] if need_pre else inputs[i]
                    if act != ACT_IDS[None]:
                        g_y = torch.empty_like(g_in)
                        call("mgn_act_bwd", _dt(g_in), _p(g_in), _p(ref), act, _p(g_y), g_in.numel(), _stream())
                    else:
                        g_y = g_in
                else:
                    g_x = g_in
        return (g_x, g_res, None, None, None, None, *grads)


def mlp_forward(x: Tensor, params: Sequence[Optional[Tensor]], n_linear: int, act: str, has_norm: bool,
                residual: Optional[Tensor] = None, eps: float = 1e-5) -> Tensor:
    if act not in ACT_IDS:
        raise NotImplementedError(f"activation '{act}' has no modulus_b200 kernel")
    return MLPFn.apply(x, residual, ACT_IDS[act], n_linear, has_norm, eps, *params)


class ActFn(torch.autograd.Function):
    """Standalone activation (leading activation of MeshGraphEdgeMLPSum, mesh_graph_mlp.py:352)."""

    @staticmethod
    def forward(ctx, x, act: int):
        require_cuda(x)
        x = _c(x)
        y = torch.empty_like(x)
        call("mgn_act_fwd", _dt(x), _p(x), act, _p(y), x.numel(), _stream())
        ctx.act = act
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _c(g)
        gx = torch.empty_like(g)
        call("mgn_act_bwd", _dt(g), _p(g), _p(x), ctx.act, _p(gx), g.numel(), _stream())
        return gx, None


def activation(x: Tensor, act: str) -> Tensor:
    if act not in ACT_IDS:
        raise NotImplementedError(f"activation '{act}' has no modulus_b200 kernel")
    if ACT_IDS[act] == 0:
        return x
    return ActFn.apply(x, ACT_IDS[act])


# ----------------------------------------------------------------------------------------
# tensor-core (tcgen05) path: bf16 storage, hidden width 128, ReLU
# ----------------------------------------------------------------------------------------
TC_HIDDEN = 128
_STATUS = {}


TC_CHECK_EVERY = 50  # fused forward passes between two non-blocking looks at the status word (see tc_poll)
_POLL = {}


def tc_status(device) -> Tensor:
    """Device int the fused kernels OR an error code into (bounded mbarrier wait ran out = a pipeline stall that left
    partial outputs).  Never read synchronously on the hot path; surfaced by `tc_poll` (every model forward),
    `tc_found_inf` (skips the optimizer step on the device) and `tc_check` (explicit, synchronising)."""
    key = torch.device(device).index or 0
    if key not in _STATUS:
        _STATUS[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return _STATUS[key]


def tc_check(device="cuda") -> None:
    code = int(tc_status(device).item())
    if code != 0:
        tc_status(device).zero_()
        raise _lib.MGNError(f"libmgn_b200: fused tensor-core kernel reported internal status {code}")


def tc_poll(device) -> None:
    """Production-path check without a per-step synchronisation: every TC_CHECK_EVERY calls the status word is copied
    to pinned host memory asynchronously; the copy started at the PREVIOUS poll is inspected (its event has long
    completed), so a stalled kernel is reported at most 2 x TC_CHECK_EVERY steps later instead of never.  Not taken
    while a CUDA graph is being captured."""
    dev = torch.device(device)
    key = dev.index or 0
    st = _POLL.get(key)
    if st is None:
        st = _POLL[key] = {"n": 0, "host": torch.zeros(1, dtype=torch.int32).pin_memory(), "event": None}
    st["n"] += 1
    if st["n"] % TC_CHECK_EVERY != 0 or torch.cuda.is_current_stream_capturing():
        return
    if st["event"] is not None and st["event"].query():
        code = int(st["host"][0])
        if code != 0:
            st["event"] = None
            tc_status(dev).zero_()
            raise _lib.MGNError(f"libmgn_b200: fused tensor-core kernel reported internal status {code} "
                                "(pipeline stall: outputs of a recent step are incomplete)")
    st["host"].copy_(tc_status(dev), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    st["event"] = ev


def tc_found_inf(device) -> Tensor:
    """fp32 [1] device flag, nonzero when a fused kernel has reported a stall since the last check: pass it as
    `found_inf` to FusedAdam.step (optim.py does so by default) and the update is skipped on the device."""
    return (tc_status(device) != 0).to(torch.float32)


def _ws(nbytes: int, dev) -> Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


def mlp3_fwd2_tc(a: Optional[Tensor], a_idx: Optional[Tensor], small_x: Optional[Tensor],
                 g1: Optional[Tensor], g1_idx: Optional[Tensor], g1_col0: int,
                 g2: Optional[Tensor], g2_idx: Optional[Tensor], g2_col0: int, M: int,
                 w1: Tensor, b1, w2, b2, w3, b3, gamma=None, beta=None, eps: float = 1e-5,
                 residual: Optional[Tensor] = None, res_is_a: bool = False, n_out: int = TC_HIDDEN,
                 out: Optional[Tensor] = None) -> Tensor:
    """Fused forward of the node / encoder / decoder forms (include/mgn_b200.h: mgn_mlp3_fwd2_tc)."""
    dev = (small_x if small_x is not None else a).device
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.bfloat16, device=dev)
    small_in = 0 if small_x is None else small_x.shape[1]
    small_f32 = int(small_x is not None and small_x.dtype == torch.float32)
    call("mgn_mlp3_fwd2_tc", _p(a), _p(a_idx), _p(small_x), small_in, small_f32,
         _p(g1), _p(g1_idx), 0 if g1 is None else g1.stride(0), g1_col0,
         _p(g2), _p(g2_idx), 0 if g2 is None else g2.stride(0), g2_col0,
         _p(residual), int(res_is_a), M, _p(w1), w1.stride(0), _p(b1), _p(w2), _p(b2), _p(w3), _p(b3),
         _p(gamma), _p(beta), n_out, eps, _p(out), out.stride(0), _p(tc_status(dev)), _stream())
    return out


def edge_block_fwd_tc(efeat: Tensor, P: Tensor, src: Tensor, dst: Tensor, csc_offsets: Tensor, n_dst: int,
                      w1a: Tensor, b1, w2, b2, w3, b3, gamma, beta, eps: float = 1e-5, h1_out: Optional[Tensor] = None):
    """MeshEdgeBlock forward fused with the sum aggregation of the next MeshNodeBlock
    (include/mgn_b200.h: mgn_edge_block_fwd_tc).  P [N, >=2H]: source projections in columns [0,H), destination
    projections in [H,2H).  Returns (efeat_new [E,H], agg [n_dst,H]) bf16."""
    E = efeat.shape[0]
    dev = efeat.device
    out = torch.empty((E, TC_HIDDEN), dtype=torch.bfloat16, device=dev)
    agg = torch.empty((n_dst, TC_HIDDEN), dtype=torch.bfloat16, device=dev)
    nbytes = _lib.load().mgn_mlp3_fwd2_agg_workspace_bytes(E)
    ws = _ws(nbytes, dev)
    call("mgn_edge_block_fwd_tc", _p(efeat), _p(P), _p(src), P.stride(0), 0, _p(P), _p(dst), P.stride(0), TC_HIDDEN, E,
         _p(w1a), w1a.stride(0), _p(b1), _p(w2), _p(b2), _p(w3), _p(b3), _p(gamma), _p(beta), eps, _p(out), _p(h1_out),
         _p(csc_offsets), n_dst, _p(agg), agg.stride(0), _p(ws), nbytes, _p(tc_status(dev)), _stream())
    return out, agg


def node_block_fwd_tc(agg: Tensor, P: Tensor, p_col0: int, nfeat: Tensor, w1a: Tensor, b1, w2, b2, w3, b3, gamma, beta,
                      eps: float = 1e-5, h1_out: Optional[Tensor] = None) -> Tensor:
    """MeshNodeBlock forward (include/mgn_b200.h: mgn_node_block_fwd_tc); optionally keeps relu(z1) for the backward."""
    N = agg.shape[0]
    out = torch.empty((N, TC_HIDDEN), dtype=torch.bfloat16, device=agg.device)
    call("mgn_node_block_fwd_tc", _p(agg), _p(P), P.stride(0), p_col0, _p(nfeat), N, _p(w1a), w1a.stride(0), _p(b1), _p(w2),
         _p(b2), _p(w3), _p(b3), _p(gamma), _p(beta), eps, _p(out), _p(h1_out), _p(tc_status(agg.device)), _stream())
    return out


def agg_workspace(total_tiles: int, dev) -> Tensor:
    """Record array shared by the launches of one partitioned edge forward (2 records of 128 fp32 + id per tile)."""
    return _ws(2 * int(total_tiles) * (TC_HIDDEN * 4 + 4), dev)


def edge_block_fwd_part_tc(efeat: Tensor, g1: Tensor, g1_idx: Tensor, P: Tensor, dst: Tensor, csc_offsets: Tensor,
                           n_dst: int, w1a: Tensor, b1, w2, b2, w3, b3, gamma, beta, eps: float, out: Tensor,
                           agg: Tensor, ws: Tensor, row_base: int, total_tiles: int, rec_base: int,
                           h1_out: Optional[Tensor] = None) -> None:
    """One row range of a partitioned edge forward with fused aggregation (mgn_edge_block_fwd_part_tc): source
    projections from `g1` (local table or exchanged rows) via g1_idx, destination projections from P[:, H:2H]."""
    call("mgn_edge_block_fwd_part_tc", _p(efeat), _p(g1), _p(g1_idx), g1.stride(0), 0, _p(P), _p(dst), P.stride(0),
         TC_HIDDEN, efeat.shape[0], _p(w1a), w1a.stride(0), _p(b1), _p(w2), _p(b2), _p(w3), _p(b3), _p(gamma), _p(beta),
         eps, _p(out), _p(h1_out), _p(csc_offsets), n_dst, _p(agg), agg.stride(0), _p(ws), ws.numel(), row_base, total_tiles,
         rec_base,
         _p(tc_status(efeat.device)), _stream())


def agg_fixup(ws: Tensor, total_tiles: int, agg: Tensor, n_dst: int) -> None:
    call("mgn_agg_fixup", _p(ws), total_tiles, _p(agg), agg.stride(0), n_dst, _stream())


def edge_block_bwd_tc(efeat: Tensor, h1: Tensor, go1: Tensor, go1_idx: Optional[Tensor], go2: Optional[Tensor],
                      go2_idx: Optional[Tensor], w1a: Tensor, w2, b2, w3, b3, gamma, eps: float,
                      g_w1a: Tensor, g_b1, g_w2, g_b2, g_w3, g_b3, g_gamma, g_beta, add_gout: bool = True,
                      g_z1_out: Optional[Tensor] = None, csc_offsets: Optional[Tensor] = None,
                      dst: Optional[Tensor] = None, dst_sum_out: Optional[Tensor] = None):
    """MeshEdgeBlock backward from the stored h1 (include/mgn_b200.h: mgn_edge_block_bwd_tc).  Returns
    (g_efeat, g_z1) bf16 [E,128]; parameter gradients go to the caller-allocated fp32 tensors."""
    E = efeat.shape[0]
    dev = efeat.device
    g_e = torch.empty((E, TC_HIDDEN), dtype=torch.bfloat16, device=dev)
    g_z1 = g_z1_out if g_z1_out is not None else torch.empty((E, TC_HIDDEN), dtype=torch.bfloat16, device=dev)
    nbytes = _lib.load().mgn_edge_block_bwd_tc_workspace_bytes(E)
    ws = _ws(nbytes, dev)
    abytes = _lib.load().mgn_mlp3_fwd2_agg_workspace_bytes(E) if csc_offsets is not None else 0
    aws = _ws(abytes, dev) if csc_offsets is not None else None
    call("mgn_edge_block_bwd_tc", _p(efeat), _p(h1), _p(go1), _p(go1_idx), _p(go2), _p(go2_idx), E, _p(w1a), w1a.stride(0),
         _p(w2), _p(b2), _p(w3), _p(b3), _p(gamma), eps, int(add_gout), _p(g_e), _p(g_z1), g_z1.stride(0), _p(g_w1a),
         g_w1a.stride(0),
         _p(g_b1), _p(g_w2), _p(g_b2), _p(g_w3), _p(g_b3), _p(g_gamma), _p(g_beta), _p(ws), nbytes,
         _p(csc_offsets), _p(dst), 0 if dst_sum_out is None else dst_sum_out.shape[0], _p(dst_sum_out),
         0 if dst_sum_out is None else dst_sum_out.stride(0), _p(aws), abytes, _p(tc_status(dev)), _stream())
    return g_e, g_z1


def mlp3_bwd_tc(a: Optional[Tensor], a_idx: Optional[Tensor], small_x: Optional[Tensor],
                g1: Optional[Tensor], g1_idx: Optional[Tensor], g1_col0: int,
                g2: Optional[Tensor], g2_idx: Optional[Tensor], g2_col0: int,
                go1: Tensor, go2: Optional[Tensor], go2_idx: Optional[Tensor], M: int,
                w1: Tensor, b1, w2, b2, w3, b3, gamma, n_out: int, eps: float,
                want_ga: bool, add_gout: bool, want_gz1: bool,
                g_w1: Tensor, g_b1, g_w2, g_b2, g_w3, g_b3, g_gamma, g_beta,
                go1_idx: Optional[Tensor] = None, g_z1_out: Optional[Tensor] = None):
    """Fused backward (include/mgn_b200.h: mgn_mlp3_bwd_tc).  Gradient tensors are caller-allocated fp32
    (g_w1 may be a column-block view); returns (g_a, g_z1) bf16 [M,128] or None."""
    dev = go1.device
    g_a = torch.empty((M, TC_HIDDEN), dtype=torch.bfloat16, device=dev) if want_ga else None
    g_z1 = g_z1_out
    if want_gz1 and g_z1 is None:
        g_z1 = torch.empty((M, TC_HIDDEN), dtype=torch.bfloat16, device=dev)
    nbytes = _lib.load().mgn_mlp3_bwd_tc_workspace_bytes(M)
    ws = _ws(nbytes, dev)
    small_in = 0 if small_x is None else small_x.shape[1]
    small_f32 = int(small_x is not None and small_x.dtype == torch.float32)
    common = (_p(a), _p(a_idx), _p(small_x), small_in, small_f32,
              _p(g1), _p(g1_idx), 0 if g1 is None else g1.stride(0), g1_col0,
              _p(g2), _p(g2_idx), 0 if g2 is None else g2.stride(0), g2_col0,
              _p(go1), _p(go1_idx), _p(go2), _p(go2_idx), M, _p(w1), w1.stride(0), _p(b1), _p(w2), _p(b2), _p(w3), _p(b3),
              _p(gamma), n_out, eps, _p(g_a), int(add_gout), _p(g_z1), 0 if g_z1 is None else g_z1.stride(0), _p(g_w1),
              g_w1.stride(0), _p(g_b1), _p(g_w2), _p(g_b2), _p(g_w3), _p(g_b3), _p(g_gamma), _p(g_beta), _p(ws), nbytes,
              _p(tc_status(dev)))
    call("mgn_mlp3_bwd_tc", *common, _stream())
    return g_a, g_z1


# ----------------------------------------------------------------------------------------
# node-level plain GEMMs of the fused path (bf16 rows, fp32 weights)
# ----------------------------------------------------------------------------------------
def linear_tc(x: Tensor, w: Tensor, out: Optional[Tensor] = None, residual: Optional[Tensor] = None) -> Tensor:
    """out[M, Nout] = x[M, K] w[Nout, K]^T (+ residual); K, Nout multiples of 128 (K <= 384).  tcgen05 GEMM
    (include/mgn_b200.h: mgn_node_gemm_tc; mgn_linear128_tc per 128-column block for the wider products)."""
    M, K = x.shape
    n_out = w.shape[0]
    if K % TC_HIDDEN or n_out % TC_HIDDEN or K > 3 * TC_HIDDEN or w.shape[1] != K or x.dtype != torch.bfloat16:
        raise ValueError(f"linear_tc: unsupported shapes x{tuple(x.shape)} w{tuple(w.shape)} {x.dtype}")
    if residual is not None and n_out != TC_HIDDEN:
        raise ValueError("linear_tc: residual needs a 128-wide output")
    w = _f32(w)
    if x.stride(1) != 1:
        x = x.contiguous()
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.bfloat16, device=x.device)
    st = tc_status(x.device)
    kb = K // TC_HIDDEN
    res = None if residual is None else _c(residual)
    nb = n_out // TC_HIDDEN
    if kb * nb <= 3 and out.stride(1) == 1 and (res is None or res.data_ptr() != out.data_ptr()):
        # one TMA-staged pass over x and out (include/mgn_b200.h: mgn_node_gemm_tc)
        call("mgn_node_gemm_tc", _p(x), x.stride(0), kb, M, _p(w), w.stride(0), nb, _p(res), _p(out), out.stride(0),
             _p(st), _stream())
        return out
    for j in range(n_out // TC_HIDDEN):
        oj = out[:, TC_HIDDEN * j: TC_HIDDEN * (j + 1)]
        for k in range(kb):  # K blocks accumulate through the residual input
            xk = x[:, TC_HIDDEN * k: TC_HIDDEN * (k + 1)]
            wjk = w[TC_HIDDEN * j: TC_HIDDEN * (j + 1), TC_HIDDEN * k: TC_HIDDEN * (k + 1)]
            # in place is safe: a tile's residual rows are staged in shared memory before the same tile's output
            # rows are written, and tiles own disjoint rows
            r = res if k == 0 else oj
            call("mgn_linear128_tc", _p(xk), x.stride(0), M, _p(wjk), w.stride(0), None, _p(r),
                 _p(oj), out.stride(0), _p(st), _stream())
    return out


def wgrad_tc(g: Tensor, x: Tensor) -> Tensor:
    """[Ng, 128] fp32 = g[M, Ng]^T x[M, 128], Ng in {128, 256, 384} (include/mgn_b200.h: mgn_wgrad_tc)."""
    M, ng = g.shape
    if ng % TC_HIDDEN or ng > 3 * TC_HIDDEN or x.shape != (M, TC_HIDDEN) or g.dtype != torch.bfloat16:
        raise ValueError(f"wgrad_tc: unsupported shapes g{tuple(g.shape)} x{tuple(x.shape)}")
    if g.stride(1) != 1:
        g = g.contiguous()
    x = _c(x)
    out = torch.empty((ng, TC_HIDDEN), dtype=torch.float32, device=g.device)
    if M == 0:
        return out.zero_()
    jb = ng // TC_HIDDEN
    nbytes = _lib.load().mgn_wgrad_tc_workspace_bytes(M, jb)
    ws = _ws(nbytes, g.device)
    call("mgn_wgrad_tc", _p(g), g.stride(0), jb, _p(x), x.stride(0), M, _p(out), out.stride(0), _p(ws), nbytes,
         _p(tc_status(g.device)), _stream())
    return out
