"""Aggregate forward on the power-law graph of the C5 microbench (tools/bench_c5.py), one width: wall time per call and,
under `ncu --metrics gpu__time_duration.sum`, the split over its kernels.  python tools/prof_powerlaw.py [H] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from modulus_b200 import ops
from modulus_b200.mesh import power_law_graph_csc, triangle_grid_mesh

DEV = "cuda:0"
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N, E = 1000000, 5992002
off, idx = power_law_graph_csc(N, E, alpha=1.2, seed=0, device=DEV)
deg = off[1:] - off[:-1]
print(f"power-law: max degree {int(deg.max())}, segments > {ops.LONG_SEGMENT} rows: {int((deg > ops.LONG_SEGMENT).sum())} "
      f"holding {int(deg[deg > ops.LONG_SEGMENT].sum())} rows; 64..{ops.LONG_SEGMENT}: {int(((deg > 64) & (deg <= ops.LONG_SEGMENT)).sum())} "
      f"holding {int(deg[(deg > 64) & (deg <= ops.LONG_SEGMENT)].sum())}; empty {int((deg == 0).sum())}")
plans = {"power-law": ops.GraphPlan.from_csc(off.int(), idx.int(), N, N)}
mesh = triangle_grid_mesh(1000, 1000, device=DEV)
plans["regular"] = ops.GraphPlan.from_csc(mesh["offsets"], mesh["indices"], N, N)
for dt in (torch.bfloat16, torch.float32):
    ef = torch.randn(E, H, device=DEV, dtype=dt)
    nf = torch.randn(N, H, device=DEV, dtype=dt)
    for name, plan in plans.items():
        for _ in range(2):
            ops.AggConcatFn.apply(ef, nf, plan, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.AggConcatFn.apply(ef, nf, plan, False)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:10s} H={H} {str(dt)[6:]:9s} aggregate_and_concat fwd {e0.elapsed_time(e1) / reps:.3f} ms")
