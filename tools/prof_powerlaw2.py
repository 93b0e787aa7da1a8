"""Which degree class of the power-law graph costs the segmented sum its time: the same 1 M segments with the hubs /
mid-length / short segments emptied in turn (rows re-packed), H = 128 bf16.  python tools/prof_powerlaw2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from modulus_b200 import ops
from modulus_b200.mesh import power_law_graph_csc

DEV = "cuda:0"
N, E, H = 1000000, 5992002, 128
off, _ = power_law_graph_csc(N, E, alpha=1.2, seed=0, device=DEV)
deg0 = (off[1:] - off[:-1])
ef = torch.randn(E, H, device=DEV, dtype=torch.bfloat16)


def run(name, deg):
    o = torch.zeros(N + 1, dtype=torch.int32, device=DEV)
    o[1:] = torch.cumsum(deg, 0).int()
    rows = int(o[-1])
    x = ef[:max(rows, 1)]
    for _ in range(3):
        ops.segment_sum(x, 0, H, o, None, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.segment_sum(x, 0, H, o, None, N)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:46s} rows {rows:8d} non-empty {int((deg > 0).sum()):7d} max {int(deg.max()):8d}: {e0.elapsed_time(e1) / 10 * 1e3:7.1f} us")


L = ops.LONG_SEGMENT
run("all", deg0)
run(f"hubs only (> {L})", torch.where(deg0 > L, deg0, 0))
run(f"no hubs (<= {L})", torch.where(deg0 <= L, deg0, 0))
run("<= 64", torch.where(deg0 <= 64, deg0, 0))
run("<= 6", torch.where(deg0 <= 6, deg0, 0))
run("all empty", torch.zeros_like(deg0))
run("65..256 only", torch.where((deg0 > 64) & (deg0 <= L), deg0, 0))
run("7..64 only", torch.where((deg0 > 6) & (deg0 <= 64), deg0, 0))
run("uniform 6", torch.full_like(deg0, 6)[: E // 6].new_full((N,), 0).index_fill_(0, torch.arange(E // 6, device=DEV), 6))
run("top hub only", torch.where(deg0 == deg0.max(), deg0, 0))
run("hubs 257..2048", torch.where((deg0 > L) & (deg0 <= 2048), deg0, 0))
