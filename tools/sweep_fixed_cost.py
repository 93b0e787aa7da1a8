"""Fixed cost per launch of the fused kernels: time vs rows (k tiles per CTA), fit a + b*k.  python tools/sweep_fixed_cost.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulus_b200 import ops
DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device=DEV)
w1, w2, w3 = r(128, 384) / 20, r(128, 128) / 11, r(128, 128) / 11
b1, b2, b3, gamma, beta = r(128) * .1, r(128) * .1, r(128) * .1, 1 + .1 * r(128), .1 * r(128)
gw1 = torch.empty(128, 384, device=DEV); gw2 = torch.empty(128, 128, device=DEV); gw3 = torch.empty(128, 128, device=DEV)
gb = [torch.empty(128, device=DEV) for _ in range(5)]
wp = r(384, 128) / 11

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2] * 1e3  # us

print("k tiles/CTA | rows | edge bwd(from h1) | edge fwd(+agg,h1) | node gemm P | wgrad T^T x | csr segsum   [us]")
for k in (1, 2, 4, 8, 16, 32):
    E = 148 * 128 * k
    N = max(E // 6, 128)
    ef, h1, ge = r(E, 128).bfloat16(), r(E, 128).abs().bfloat16(), r(E, 128).bfloat16()
    P = r(N, 384).bfloat16(); nf = r(N, 128).bfloat16(); T3 = r(N, 384).bfloat16()
    dst = torch.sort(torch.randint(0, N, (E,), device=DEV, generator=g)).values.int()
    src = torch.randint(0, N, (E,), device=DEV, generator=g).int()
    off = torch.zeros(N + 1, dtype=torch.int32, device=DEV); off[1:] = torch.cumsum(torch.bincount(dst.long(), minlength=N), 0).int()
    plan_csr = ops._group_by_key(src, N)
    h1o = torch.empty(E, 128, dtype=torch.bfloat16, device=DEV)
    t_bwd = timeit(lambda: ops.edge_block_bwd_tc(ef, h1, ge, None, None, None, w1[:, :128], w2, b2, w3, b3, gamma, 1e-5, gw1[:, :128], gb[0], gw2, gb[1], gw3, gb[2], gb[3], gb[4]))
    t_fwd = timeit(lambda: ops.edge_block_fwd_tc(ef, P, src, dst, off, N, w1[:, :128], b1, w2, b2, w3, b3, gamma, beta, h1_out=h1o))
    t_g = timeit(lambda: ops.linear_tc(nf, wp))
    t_w = timeit(lambda: ops.wgrad_tc(T3, nf))
    t_s = timeit(lambda: ops.segment_sum(ge, 0, 128, plan_csr[0], plan_csr[1], N))
    print(f"{k:3d} | {E:8d} | {t_bwd:8.1f} | {t_fwd:8.1f} | {t_g:8.1f} | {t_w:8.1f} | {t_s:8.1f}")
ops.tc_check(DEV)
