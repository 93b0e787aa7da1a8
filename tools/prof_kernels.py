"""Stand-alone launches of the fused kernels (for ncu captures and timing): python tools/prof_kernels.py NX NY REPS.
Phase breakdowns need a library built with MGN_NVCC_EXTRA=-DMGN_DEBUG_HOOKS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulus_b200 import ops
from modulus_b200.mesh import triangle_grid_mesh

DEV = "cuda:0"
nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (316, 317)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mesh = triangle_grid_mesh(nx, ny, device=DEV)
N, E = mesh["num_nodes"], int(mesh["indices"].numel())
plan = ops.GraphPlan.from_csc(mesh["offsets"], mesh["indices"], N, N)
g = torch.Generator(device=DEV).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device=DEV)
efeat, P = r(E, 128).bfloat16(), (r(N, 384) * 0.5).bfloat16()
g_e, g_agg = r(E, 128).bfloat16(), r(N, 128).bfloat16()
w1, w2, w3 = r(128, 384) / 20, r(128, 128) / 11, r(128, 128) / 11
b1, b2, b3, gamma, beta = r(128) * .1, r(128) * .1, r(128) * .1, 1 + .1 * r(128), .1 * r(128)
gw1 = torch.empty(128, 384, device=DEV); gw2 = torch.empty(128, 128, device=DEV); gw3 = torch.empty(128, 128, device=DEV)
gb = [torch.empty(128, device=DEV) for _ in range(5)]

def fwd2():
    return ops.mlp3_fwd2_tc(efeat, None, None, P, plan.src, 0, P, plan.dst, 128, E, w1[:, :128], b1, w2, b2, w3, b3,
                            gamma, beta, res_is_a=True)

def eblk():
    return ops.edge_block_fwd_tc(efeat, P, plan.src, plan.dst, plan.csc_offsets, N, w1[:, :128], b1, w2, b2, w3, b3, gamma, beta)

def bwd():
    return ops.mlp3_bwd_tc(efeat, None, None, P, plan.src, 0, P, plan.dst, 128, g_e, g_agg, plan.dst, E,
                           w1[:, :128], b1, w2, b2, w3, b3, gamma, 128, 1e-5, True, True, True,
                           gw1[:, :128], gb[0], gw2, gb[1], gw3, gb[2], gb[3], gb[4])

h1s = torch.empty(E, 128, dtype=torch.bfloat16, device=DEV)
ops.edge_block_fwd_tc(efeat, P, plan.src, plan.dst, plan.csc_offsets, N, w1[:, :128], b1, w2, b2, w3, b3, gamma, beta, h1_out=h1s)

def eblk_h1():
    return ops.edge_block_fwd_tc(efeat, P, plan.src, plan.dst, plan.csc_offsets, N, w1[:, :128], b1, w2, b2, w3, b3, gamma, beta, h1_out=h1s)

def bwd2():
    return ops.edge_block_bwd_tc(efeat, h1s, g_e, None, g_agg, plan.dst, w1[:, :128], w2, b2, w3, b3, gamma, 1e-5,
                                 gw1[:, :128], gb[0], gw2, gb[1], gw3, gb[2], gb[3], gb[4])

Tz = torch.empty(N, 384, dtype=torch.bfloat16, device=DEV)
nfeat_n = r(N, 128).bfloat16()

def bwd2_dst():
    return ops.edge_block_bwd_tc(efeat, h1s, g_e, None, g_agg, plan.dst, w1[:, :128], w2, b2, w3, b3, gamma, 1e-5,
                                 gw1[:, :128], gb[0], gw2, gb[1], gw3, gb[2], gb[3], gb[4], csc_offsets=plan.csc_offsets,
                                 dst=plan.dst, dst_sum_out=Tz[:, 128:256])

agg_n = r(N, 128).bfloat16(); h1n = torch.empty(N, 128, dtype=torch.bfloat16, device=DEV)

def nodefwd():
    return ops.node_block_fwd_tc(agg_n, P, 256, nfeat_n, w1[:, :128], b1, w2, b2, w3, b3, gamma, beta, h1_out=h1n)

def nodebwd():
    return ops.edge_block_bwd_tc(agg_n, h1n, g_agg, None, None, None, w1[:, :128], w2, b2, w3, b3, gamma, 1e-5, gw1[:, :128],
                                 gb[0], gw2, gb[1], gw3, gb[2], gb[3], gb[4], add_gout=False, g_z1_out=Tz[:, 256:])

def agg():
    return ops.segment_sum(efeat, 0, 128, plan.csc_offsets, None, N)

def csr():
    return ops.segment_sum(efeat, 0, 128, plan.csr_offsets, plan.csr_eids, N)

nfeat = r(N, 128).bfloat16(); T3 = r(N, 384).bfloat16(); wp = r(384, 128) / 11; wpt = wp.t().contiguous()

def lin_p():
    return ops.linear_tc(nfeat, wp)

def lin_t():
    return ops.linear_tc(T3, wpt, residual=nfeat)

def wgrad():
    return ops.wgrad_tc(T3, nfeat)

ONLY = os.environ.get("MGN_PROF_ONLY")  # run just one of the functions above (ncu captures): e.g. MGN_PROF_ONLY=bwd2_dst
if ONLY:
    fn = globals()[ONLY]
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    ops.tc_check(DEV)
    sys.exit(0)

_ALL = (("fwd2 edge", fwd2), ("eblk fwd3+agg", eblk), ("eblk fwd3+agg+h1", eblk_h1), ("bwd edge (recompute)", bwd), ("bwd edge (from h1)", bwd2), ("bwd edge (from h1) + dst sums", bwd2_dst), ("node fwd (+h1)", nodefwd), ("node bwd (from h1)", nodebwd), ("segsum csc", agg), ("segsum csr", csr),
                 ("P=nfeat Wp^T", lin_p), ("g_n+T Wp", lin_t), ("T^T nfeat", wgrad))
if os.environ.get("MGN_PROF_ONLY2"):  # quick A/B runs: the two edge-forward launches only
    _ALL = _ALL[4:6] if os.environ["MGN_PROF_ONLY2"] == "bwd" else _ALL[1:3]
for name, fn in _ALL:
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in evs:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    ms = ts[len(ts) // 2]
    print(f"{name:12s} N={N} E={E}: {ms:.3f} ms median, {ts[0]:.3f} min of {reps}  ({E / ms / 1e3:.1f} M edges/s)", flush=True)
ops.tc_check(DEV)
if os.environ.get("MGN_PROF_ONLY2"):
    sys.exit(0)

# per-phase cycle breakdown of CTA 0 (only in a library built with -DMGN_DEBUG_HOOKS, see include/mgn_b200_debug.h)
from modulus_b200 import _lib
if not _lib.has_debug_hooks():
    print("(no phase breakdown: build with MGN_NVCC_EXTRA=-DMGN_DEBUG_HOOKS for it)")
    sys.exit(0)
tbuf = torch.zeros(96 + 4 * 148, dtype=torch.int64, device=DEV)

def per_cta_clock(t, label):
    # (cycles, ns, tiles) of every CTA's tile loop -> effective SM clock and spread across CTAs
    c = t[96:].view(-1, 4).double()
    c = c[c[:, 2] > 0]
    if c.numel() == 0:
        return
    cyc_tile = c[:, 0] / c[:, 2]
    ghz = c[:, 0] / c[:, 1].clamp(min=1)
    if os.environ.get("MGN_PROF_DUMP"):
        order = torch.argsort(c[:, 3])
        print(f"{label}: cycles/tile by SM id:", " ".join(f"{int(c[i, 3])}:{int(cyc_tile[i])}" for i in order.tolist()))
    print(f"{label}: per-CTA cycles/tile min {cyc_tile.min():.0f} median {cyc_tile.median():.0f} max {cyc_tile.max():.0f}; "
          f"loop time us min {c[:, 1].min() / 1e3:.0f} max {c[:, 1].max() / 1e3:.0f}; SM clock GHz min {ghz.min():.3f} median {ghz.median():.3f} max {ghz.max():.3f}")

_lib.call("mgn_debug_set_bwd_timing", tbuf.data_ptr())
bwd(); torch.cuda.synchronize()
_lib.call("mgn_debug_set_bwd_timing", None)
t = tbuf.cpu()[:96].view(3, 32)
n_tiles = (E + 127) // 128
per_cta = -(-n_tiles // 148)
names = {0: "MMA  : wA+E6prev | g1 | wE1 | g2 | wE2 | g3 | wE3 | L3 | wE4 | L2 | wE5 | dgrad1+wA2 | wgrad1",
         1: "LOADER: top | wE1 | issue go1 | wW3+CS0 | issue A2 | wE5+st gz1 | wMMA7 | - | issue A' | wE6 | st gA",
         2: "EPI  : wMMA1+G | E1 | wMMA2 | E2 | wMMA3 | wGO | E3 | wMMA4 | E4 | wMMA5 | E5 | wMMA6 | E6"}
for r in range(3):
    print(names[r])
    print("   cycles/tile:", [int(v) // per_cta for v in t[r, :16].tolist()], " total/tile:", int(t[r].sum()) // per_cta)

tbuf.zero_()
_lib.call("mgn_debug_set_fwd2_timing", tbuf.data_ptr())
fwd2(); torch.cuda.synchronize()
_lib.call("mgn_debug_set_fwd2_timing", None)
t = tbuf.cpu()[:96].view(3, 32)
print("FWD2 kernel, CTA 0, cycles per tile")
print(" MMA   : wait IN+OUT | issue1 | wait H1 | wait H2 (incl issue2) | issue3 :", [int(v) // per_cta for v in t[0, :5].tolist()], "total", int(t[0].sum()) // per_cta)
print(" MOVER : issue next A | wait H1 | wait OUT (incl issue G) | store+ids | wait cp+sync :", [int(v) // per_cta for v in t[1, :5].tolist()], "total", int(t[1].sum()) // per_cta)
print(" EPI   : wait M1+IN | E1 | wait M2 | E2 | wait M3 | E3 pass2 | E3 stats | E3 exchange :", [int(v) // per_cta for v in t[2, :8].tolist()], "total", int(t[2].sum()) // per_cta)

tbuf.zero_()
_lib.call("mgn_debug_set_fwd2_timing", tbuf.data_ptr())
nodefwd(); torch.cuda.synchronize()
_lib.call("mgn_debug_set_fwd2_timing", None)
t = tbuf.cpu()[:96].view(3, 32)
pc_n = -(-((N + 127) // 128) // 148)
print("FWD2 kernel NODE form, CTA 0, cycles per tile")
print(" MMA   : wait IN+OUT | issue1 | wait H1 | wait H2 (incl issue2) | issue3 :", [int(v) // pc_n for v in t[0, :5].tolist()], "total", int(t[0].sum()) // pc_n)
print(" MOVER : issue next A | wait H1 | wait OUT (incl issue G) | store+ids | wait cp+sync :", [int(v) // pc_n for v in t[1, :5].tolist()], "total", int(t[1].sum()) // pc_n)
print(" EPI   : wait M1+IN | E1 | wait M2 | E2 | wait M3 | E3 pass2 | E3 stats | E3 exchange :", [int(v) // pc_n for v in t[2, :8].tolist()], "total", int(t[2].sum()) // pc_n)

tbuf.zero_()
_lib.call("mgn_debug_set_fwd2_timing", tbuf.data_ptr())
eblk(); torch.cuda.synchronize()
_lib.call("mgn_debug_set_fwd2_timing", None)
t = tbuf.cpu()
print("FWD3 edge-block kernel, CTA 0, epilogue cycles per tile: wait M2 | E2 | wait M1+G | E1 | wait M3 | E3 :",
      [int(v) // per_cta for v in t[:6].tolist()], "total", int(t[:6].sum()) // per_cta)
print("FWD3 MMA thread: wait E2 | issue M3 | wait E1' | issue M2' | wait E3prev+A'' | issue M1'' :",
      [int(v) // per_cta for v in t[8:14].tolist()], "total", int(t[8:14].sum()) // per_cta)

tbuf.zero_()
_lib.call("mgn_debug_set_edge_bwd2_timing", tbuf.data_ptr())
bwd2(); torch.cuda.synchronize()
_lib.call("mgn_debug_set_edge_bwd2_timing", None)
t = tbuf.cpu()[:96].view(3, 32)
print("BWD2 (from h1) kernel, CTA 0, cycles per tile")
print(" MMA   : wH1+E6prev | g2 | wE2 | g3 | wE3 | L3 | wE4 | L2 | wE5 | dgrad1 | wgrad1 :", [int(v) // per_cta for v in t[0, :11].tolist()], "total", int(t[0].sum()) // per_cta)
print(" LOADER: top | wW3+CS0 | wW2+CS1 | wE5 | st gz1 | wMMA7+CS2 | wE6 | st gA :", [int(v) // per_cta for v in t[1, :8].tolist()], "total", int(t[1].sum()) // per_cta)
print(" EPI   : wMMA2+XF | E2 | wMMA3 | wGO | E3 | E4 | E5 | E6 :", [int(v) // per_cta for v in t[2, :8].tolist()], "total", int(t[2].sum()) // per_cta)
per_cta_clock(tbuf.cpu(), "BWD2 (from h1), single launch")
tbuf.zero_()
_lib.call("mgn_debug_set_edge_bwd2_timing", tbuf.data_ptr())
for _ in range(30):
    bwd2()
torch.cuda.synchronize()
_lib.call("mgn_debug_set_edge_bwd2_timing", None)
per_cta_clock(tbuf.cpu(), "BWD2 (from h1), 30th back-to-back launch")

tbuf.zero_()
_lib.call("mgn_debug_set_fwd2_timing", tbuf.data_ptr())
eblk_h1(); torch.cuda.synchronize()
_lib.call("mgn_debug_set_fwd2_timing", None)
t = tbuf.cpu()
print("FWD3 + h1 store, epilogue: wait M2 | E2 | wait M1+G | E1 | wait M3 | E3 :", [int(v) // per_cta for v in t[:6].tolist()], "total", int(t[:6].sum()) // per_cta)
print("FWD3 + h1 store, MMA thread: wait E2 | issue M3 | wait E1' | issue M2' | wait E3prev+A'' | issue M1'' :", [int(v) // per_cta for v in t[8:14].tolist()])
per_cta_clock(t, "FWD3 + h1 store")
print("FWD3 + h1 store, mover warp 1: idx loads | wait E1' | h1 store | gather + publish | wait E3 | dst sums :", [int(v) // per_cta for v in t[16:22].tolist()], "total", int(t[16:22].sum()) // per_cta)

# experiment (debug builds only): the same launch without the fused destination sums
os.environ["MGN_FWD3_NO_AGG"] = "1"
for name, fn in (("eblk fwd3 (no agg)", eblk), ("eblk fwd3+h1 (no agg)", eblk_h1)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in evs:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    print(f"{name:12s} N={N} E={E}: {ts[len(ts) // 2]:.3f} ms median, {ts[0]:.3f} min of {reps}", flush=True)
tbuf.zero_()
_lib.call("mgn_debug_set_fwd2_timing", tbuf.data_ptr())
eblk_h1(); torch.cuda.synchronize()
_lib.call("mgn_debug_set_fwd2_timing", None)
t = tbuf.cpu()
per_cta_clock(t, "FWD3 + h1 store, NO dst sums")
print("FWD3 + h1 store, NO dst sums, epilogue: wait M2 | E2 | wait M1+G | E1 | wait M3 | E3 :", [int(v) // per_cta for v in t[:6].tolist()], "total", int(t[:6].sum()) // per_cta)
print("FWD3 + h1 store, NO dst sums, mover warp 1: idx loads | wait E1' | h1 store | gather + publish | wait E3 | dst sums :", [int(v) // per_cta for v in t[16:22].tolist()], "total", int(t[16:22].sum()) // per_cta)
del os.environ["MGN_FWD3_NO_AGG"]
