"""C5 microbench (BASELINE.json configs[4]): aggregate_and_concat and concat_efeat, forward and backward,
through the operator seam (ops.AggConcatFn / ops.ConcatEfeatFn -> libmgn_b200.so) at hidden 128 / 256 / 512 on a
regular degree-6 mesh and a power-law in-degree graph with the same edge count, bf16 and fp32, against the measured
HBM copy rate.  Algorithmic bytes: aggregate fwd (E+N)Hb + 2NHb + 4(N+1), bwd 2NHb + (E+N)Hb;
concat fwd (E+2N)Hb + 3EHb + 8E, bwd 3EHb + (E+2N)Hb + 2EHb (the CSR side re-reads its slice).

    python tools/bench_c5.py [n_side=1000] [reps=5]   ->  markdown table on stdout
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulus_b200 import ops
from modulus_b200.mesh import triangle_grid_mesh, power_law_graph_csc

DEV = "cuda:0"
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
PEAK = peaks["hbm_gbs"]


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


mesh = triangle_grid_mesh(side, side, device=DEV)
N = mesh["num_nodes"]
E = int(mesh["indices"].numel())
graphs = {"regular mesh (deg 6)": ops.GraphPlan.from_csc(mesh["offsets"], mesh["indices"], N, N)}
off, idx = power_law_graph_csc(N, E, alpha=1.2, seed=0, device=DEV)
graphs["power-law in-degree"] = ops.GraphPlan.from_csc(off.int(), idx.int(), N, N)
del mesh
print(f"# C5 microbench: N={N}, E={E}, {reps} reps after 2 warm-ups, peak = measured copy rate {PEAK:.0f} GB/s\n")
print("| graph | op | H | dtype | ms | algorithmic GB | GB/s | frac of HBM peak |")
print("|---|---|---:|---|---:|---:|---:|---:|")
for gname, plan in graphs.items():
    for H in (128, 256, 512):
        for dt, b in ((torch.bfloat16, 2), (torch.float32, 4)):
            torch.manual_seed(0)
            ef = torch.randn(E, H, device=DEV, dtype=dt)
            nf = torch.randn(N, H, device=DEV, dtype=dt)
            rows = []
            # aggregate_and_concat
            out = ops.AggConcatFn.apply(ef, nf, plan, False)
            g = torch.randn_like(out)
            t_f = timed(lambda: ops.AggConcatFn.apply(ef, nf, plan, False))
            efr, nfr = ef.clone().requires_grad_(True), nf.clone().requires_grad_(True)
            o = ops.AggConcatFn.apply(efr, nfr, plan, False)
            t_b = timed(lambda: torch.autograd.grad(o, (efr, nfr), g, retain_graph=True))
            rows.append(("aggregate_and_concat fwd", t_f, (E + N) * H * b + 2 * N * H * b + 4 * (N + 1)))
            rows.append(("aggregate_and_concat bwd", t_b, 2 * N * H * b + (E + N) * H * b + 4 * E))
            del out, g, o, efr, nfr
            # concat_efeat
            out = ops.ConcatEfeatFn.apply(ef, nf, nf, plan)
            g = torch.randn_like(out)
            t_f = timed(lambda: ops.ConcatEfeatFn.apply(ef, nf, nf, plan))
            efr, nfr = ef.clone().requires_grad_(True), nf.clone().requires_grad_(True)
            o = ops.ConcatEfeatFn.apply(efr, nfr, nfr, plan)
            t_b = timed(lambda: torch.autograd.grad(o, (efr, nfr), g, retain_graph=True))
            rows.append(("concat_efeat fwd", t_f, (E + 2 * N) * H * b + 3 * E * H * b + 8 * E))
            rows.append(("concat_efeat bwd", t_b, 3 * E * H * b + (E + 2 * N) * H * b + 2 * E * H * b + 8 * E))
            del out, g, o, efr, nfr, ef, nf
            torch.cuda.empty_cache()
            for name, ms, byt in rows:
                gbs = byt / (ms * 1e-3) / 1e9
                print(f"| {gname} | {name} | {H} | {'bf16' if b == 2 else 'fp32'} | {ms:.3f} | {byt / 1e9:.2f} | {gbs:.0f} | {gbs / PEAK:.2f} |", flush=True)
