#!/usr/bin/env python
"""bench.py -- MeshGraphNet fwd+bwd edges/s on B200 (BASELINE.json metric) with roofline evidence.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3] [--impl b200|reference]

One "step" = zero_grad -> forward -> MSE loss -> backward of a 15-layer, hidden-128 MeshGraphNet
(examples/cfd/vortex_shedding_mgn/train.py:151-166 of the reference; optimizer excluded)
on a synthetic mesh.  N=1 default workload is c3 = BASELINE.json configs[2] (3-D surface mesh, 1M nodes /
~6M edges, bf16): the configuration north_star's target is quoted on, and the per-GPU share of the 8M-node
configs[3] at 8 GPUs; c2 (configs[1], 100k nodes) and c1 remain selectable.  With N>1 ranks (torchrun) the mesh
grows N-fold and is partitioned
with DistributedGraph (weak scaling, halo exchange = NCCL all-to-all); value = global edges / max
over ranks of the device time.

The JSON line carries: value (inputs resident in HBM), e2e (host buffers, H2D + loss D2H inside
the timed region), roofline (dominant kernel, CUDA-event durations recorded inside the timed
region against algorithmic bytes/flops), cpu_baseline (the staged reference on the host cores, bounded
sample), clocks, gpu_launches.  `config` holds only what identifies the workload (mesh, sizes, partition scheme, the L2
note) and is the same dict in both arms for the same command line; what was measured about the partition (halo rows per
rank, halo fraction, transport) is under `halo`.

--impl reference times the reference's own CPU implementation of the same path on the host cores: the UNMODIFIED
physicsnemo MeshGraphNet staged under oracle/_ref by oracle/stage_reference.py (cpu_baseline.kind "reference"; absent
third-party imports served by oracle/ref_shim), or, where that was not staged, the oracle port oracle/mgn_oracle.py
(kind "port"; pinned to the reference's golden vectors in tests/test_oracle.py).  Each step is a bounded sample of the
SAME workload (same mesh generator, same MeshGraphNet dimensions, fewer mesh rows), plus one full-size step of c2.

The line also carries `parity`: an in-run check of the CUDA path against the oracle on a 10k-node mesh.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MGN fwd+bwd edges/s"
UNIT = "edges/s"
H = 128
L = 15

WORKLOADS = {
    # name: (generator, args, d_node, d_edge, d_out, dtype)
    "c1": ("triangle_grid_mesh", (42, 45), 6, 3, 3, "f32"),     # ~1.9k nodes (vortex_shedding_mgn size)
    "c2": ("triangle_grid_mesh", (316, 317), 6, 3, 3, "bf16"),   # 100k nodes / ~600k edges
    "c3": ("torus_surface_mesh", (1000, 1000), 11, 4, 4, "bf16"),  # 1M nodes / 6M edges
    # BASELINE configs[3]: ONE 8M-node / 48M-edge mesh partitioned over the ranks (strong scaling; >= 2 GPUs)
    "c4": ("torus_surface_mesh", (2000, 4000), 11, 4, 4, "bf16"),
}
STRONG = {"c4"}


def load_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/r01_ncu_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(workload, {}).get(kernel)
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------
# algorithmic work (SURVEY 8d / BASELINE.md section 3)
# ------------------------------------------------------------------------------------------
def step_flops(N, E, d_n, d_e, d_out):
    f_l = 10 * H * H * E + 8 * H * H * N
    f_encdec = 2 * (d_e * H + 2 * H * H) * E + 2 * (d_n * H + 2 * H * H) * N + 2 * (2 * H * H + H * d_out) * N
    return 3 * (L * f_l + f_encdec)


def step_bytes(N, E, b):
    return L * (9 * E + 13 * N) * H * b + L * (24 * E + 12 * N)


class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU arm -- the reference's own implementation (staged, unmodified) or the oracle port, on the host cores
# ------------------------------------------------------------------------------------------
# bounded sample of a workload: same generator and model dimensions, fewer mesh rows (about 40k nodes / 240k edges,
# 2-3 s of CPU work per step); c1 is small enough to run whole
SAMPLE_ROWS = {"c1": 42, "c2": 126, "c3": 40, "c4": 10}


def cpu_inputs(workload: str, rows=None):
    import torch
    from modulus_b200 import mesh as meshgen

    gen, gargs, d_n, d_e, d_out, dtype = WORKLOADS[workload]
    rows = gargs[0] if rows is None else rows
    mesh = getattr(meshgen, gen)(rows, gargs[1])
    n, E = mesh["num_nodes"], int(mesh["indices"].numel())
    g = torch.Generator().manual_seed(1)
    nf, tgt = torch.randn(n, d_n, generator=g), torch.randn(n, d_out, generator=g)
    ef = mesh["edge_features"]
    ef = (ef[:, :d_e] if ef.shape[1] >= d_e else torch.cat([ef, ef[:, :1].expand(-1, d_e - ef.shape[1])], 1)).contiguous()
    return mesh, n, E, nf, ef, tgt, f"{gen}({rows},{gargs[1]})"


CPU_BUDGET_S = 150.0  # the CPU legs stop adding timed repetitions once this much wall time is spent


def cpu_step_rate(workload: str, rows, reps: int, warmup: int, threads: int):
    """(edges/s, s per step, nodes, edges, kind, mesh name, reps done) of the CPU implementation on `rows` mesh rows of
    `workload`; at most `reps` timed steps, fewer when CPU_BUDGET_S runs out (slow hosts)."""
    import torch
    from oracle import mgn_oracle as O  # CPU baseline leg only
    from oracle import ref_runner as R

    torch.set_num_threads(threads)
    gen, gargs, d_n, d_e, d_out, dtype = WORKLOADS[workload]
    mesh, n, E, nf, ef, tgt, name = cpu_inputs(workload, rows)
    if R.available():
        model, graph = R.build(d_n, d_e, d_out, mesh["offsets"], mesh["indices"], processor_size=L, seed=0)
        kind, one = "reference", (lambda: R.step(model, graph, nf, ef, tgt))
    else:
        src, dst = O.coo_from_csc(mesh["offsets"], mesh["indices"])
        torch.manual_seed(0)
        sd = O.make_state_dict(d_n, d_e, d_out, processor_size=L, hidden=H)
        kind, one = "port", (lambda: O.step_fwd_bwd(sd, nf, ef, src, dst, tgt, processor_size=L))
    times, t_start = [], time.perf_counter()
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        one()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > CPU_BUDGET_S:
                break
    t = sum(times) / len(times)
    return E / t, t, n, E, kind, name, len(times)


def _cpu_desc(kind, name, n, E, t, reps, warmup):
    what = ("unmodified physicsnemo MeshGraphNet (oracle/_ref + oracle/ref_shim)" if kind == "reference"
            else "oracle port (oracle/mgn_oracle.py)")
    return (f"{what} on {name}: {n} nodes / {E} edges, 15 layers, hidden 128, fp32, torch CPU, {reps} reps after "
            f"{warmup} warm-up, {t:.2f} s per step")


def mesh_size(gen, gargs):
    """(nodes, edges) of a synthetic workload mesh without a GPU: the torus has degree 6 everywhere (E = 6N, modulus_b200/mesh.py),
    the small triangle grids are built on the host."""
    if gen == "torus_surface_mesh":
        n = gargs[0] * gargs[1]
        return n, 6 * n
    from modulus_b200 import mesh as meshgen

    m = getattr(meshgen, gen)(*gargs)
    return m["num_nodes"], int(m["indices"].numel())


def workload_config(args, world, n_glob, E_glob):
    """`config` of the JSON line: what identifies the workload, and nothing measured -- both arms print the SAME dict for the same
    command line (the reference arm times a bounded sample of this workload and says so in `cpu_baseline.sample`)."""
    _, _, _, _, _, dtype = WORKLOADS[args.workload]
    b = 2 if dtype == "bf16" else 4
    E1 = E_glob // world
    return {"workload": workload_name(args.workload, world), "nodes": n_glob, "edges": E_glob,
            "partition": (args.partition + (f" of {args.stripe_rows} mesh rows" if args.partition == "stripes" else
                                            " slabs (partition_graph_nodewise)")) if world > 1 else "none",
            "l2": "per-step working set (edge table alone %.0f MB) exceeds the 126 MB L2; no explicit flush"
                  % (E1 * 128 * b / 1e6) if E1 * 128 * b > 126e6 else
                  "working set smaller than L2: numbers are L2-warm (c1 is launch-bound by construction)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = SAMPLE_ROWS[args.workload]
    reps, warm = max(args.steps, 1), min(args.warmup, 1)
    rate, t, n, E, kind, name, reps = cpu_step_rate(args.workload, rows, reps, warm, threads)
    sample = _cpu_desc(kind, name, n, E, t, reps, warm)
    gen, gargs = WORKLOADS[args.workload][:2]
    n_glob, E_glob = mesh_size(gen, gargs if args.workload in STRONG else (gargs[0] * world, gargs[1]))
    full = None
    if not args.no_full_size and args.workload != "c1":
        # like-for-like at a BASELINE size: ONE step of the whole c2 mesh (100k nodes / 600k edges), no warm-up
        r2, t2, n2, E2, _, name2, _ = cpu_step_rate("c2", None, 1, 0, threads)
        full = {"workload": workload_name("c2", 1), "value": r2, "unit": UNIT, "s_per_step": t2, "nodes": n2, "edges": E2,
                "note": "whole configs[1] mesh, one step"}
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world, n_glob, E_glob), "sample": sample,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "full_size_step": full,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(w, world):
    gen, gargs, d_n, d_e, d_out, dtype = WORKLOADS[w]
    how = "" if world == 1 else (f" partitioned over {world} ranks (DistributedGraph)" if w in STRONG else
                                 f" x{world} ranks (rows scaled, DistributedGraph)")
    return f"{w}: MeshGraphNet({d_n},{d_e},{d_out}) 15 layers hidden 128 on {gen}{gargs}" + how


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def parity_check(dev):
    """In-run parity (last row): the CUDA path against the CPU oracle on the same seeded inputs and weights, a
    10k-node / 59k-edge triangle mesh through all 15 layers.  fp32 path: output, input gradients and every weight gradient
    within 1e-3 (north_star's 15-layer bar); fused bf16 path: output within 2e-2.  The oracle is the checker only."""
    import torch
    from modulus_b200.mesh import triangle_grid_mesh
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from oracle import mgn_oracle as O  # checker

    mesh = triangle_grid_mesh(100, 100)
    n = mesh["num_nodes"]
    torch.manual_seed(0)
    model = MeshGraphNet(6, 3, 3).to(dev)
    g = torch.Generator().manual_seed(2)
    nf, tgt, ef = torch.randn(n, 6, generator=g), torch.randn(n, 3, generator=g), mesh["edge_features"].clone()
    graph = CuGraphCSC(mesh["offsets"].to(dev), mesh["indices"].to(dev), n, n)
    src, dst = O.coo_from_csc(mesh["offsets"], mesh["indices"])
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref_out, _, ref_g = O.step_fwd_bwd(sd, nf, ef, src, dst, tgt, processor_size=L)

    def rel(a, b):  # outputs: max norm
        return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

    def l2(a, b):  # gradients: relative L2 (a handful of fp32 ReLU-mask flips dominate the max norm at this size, between
        # ANY two summation orders -- tests/test_gpu_baseline_sizes.py measures that floor)
        return float((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    res = {"mesh": f"triangle_grid_mesh(100,100): {n} nodes / {int(src.numel())} edges, 15 layers, hidden 128",
           "oracle": "oracle/mgn_oracle.py fp32 (pinned to the reference's goldens)"}
    for name, bf16 in (("fp32", False), ("bf16", True)):
        model.zero_grad(set_to_none=True)
        x = nf.to(dev).requires_grad_(True)
        e = ef.to(dev).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            out = model(x, e, graph)
        torch.nn.functional.mse_loss(out.float(), tgt.to(dev)).backward()
        torch.cuda.synchronize()
        res[name + "_out"] = rel(out.float(), ref_out)
        if not bf16:
            res["fp32_grad_inputs_l2"] = max(l2(x.grad, ref_g["__node_features"]), l2(e.grad, ref_g["__edge_features"]))
            res["fp32_grad_weights_l2_max"] = max(l2(p.grad, ref_g[k]) for k, p in model.named_parameters())
        else:
            res["bf16_out_l2"] = l2(out.float(), ref_out)
    res["tolerances"] = {"fp32_out_maxnorm": 1e-3, "fp32_grads_l2": 1e-3, "bf16_out_l2": 2e-2}
    res["ok"] = bool(res["fp32_out"] < 1e-3 and res["fp32_grad_inputs_l2"] < 1e-3 and res["fp32_grad_weights_l2_max"] < 1e-3
                     and res["bf16_out_l2"] < 2e-2)
    return res


def side_rate(workload, dev, peaks, steps=10, warmup=3):
    """fwd+bwd rate of another BASELINE configuration on this GPU (inputs resident), for the record next to the headline"""
    import torch
    from modulus_b200 import mesh as meshgen
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    gen, gargs, d_n, d_e, d_out, dtype = WORKLOADS[workload]
    mesh = getattr(meshgen, gen)(*gargs, device=dev)
    n, E = mesh["num_nodes"], int(mesh["indices"].numel())
    torch.manual_seed(0)
    model = MeshGraphNet(d_n, d_e, d_out).to(dev)
    graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n, n)
    g = torch.Generator().manual_seed(1)
    nf, tgt = torch.randn(n, d_n, generator=g).to(dev), torch.randn(n, d_out, generator=g).to(dev)
    ef = mesh["edge_features"]
    ef = (ef[:, :d_e] if ef.shape[1] >= d_e else torch.cat([ef, ef[:, :1].expand(-1, d_e - ef.shape[1])], 1)).float().contiguous()

    def step():
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == "bf16"):
            pred = model(nf, ef, graph)
        torch.nn.functional.mse_loss(pred.float(), tgt).backward()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for _ in range(warmup):
        step()
    ms = timed(step)
    # the same step as ONE CUDA-graph launch (every kernel on the path is capturable: no host sync, no allocation inside;
    # modulus_b200/capture.py does this for training loops): at this size the eager loop is partly bound by the host
    ms_graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        model.zero_grad(set_to_none=True)
        graph_obj = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_obj):
            step()
        for _ in range(2):
            graph_obj.replay()
        ms_graph = timed(graph_obj.replay)
    except Exception as exc:  # pragma: no cover - reported, not fatal
        ms_graph = None
        graph_err = repr(exc)[:200]
    b = 2 if dtype == "bf16" else 4
    return {"workload": workload_name(workload, 1), "value": E / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "dtype": dtype,
            "cuda_graph": None if ms_graph is None else {"ms_per_step": ms_graph, "value": E / (ms_graph * 1e-3), "unit": UNIT,
                                                         "hbm_frac": step_bytes(n, E, b) / (ms_graph * 1e-3) / 1e9 / peaks["hbm"]},
            "hbm_frac": step_bytes(n, E, b) / (ms * 1e-3) / 1e9 / peaks["hbm"],
            "tensor_frac": step_flops(n, E, d_n, d_e, d_out) / (ms * 1e-3) / 1e12 / peaks["tc_sust"],
            "l2": "edge table %.0f MB: %s the 126 MB L2" % (E * H * b / 1e6, "exceeds" if E * H * b > 126e6 else "fits")}


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from modulus_b200 import _lib, ops
    from modulus_b200 import mesh as meshgen
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    peaks = load_peaks()

    gen, gargs, d_n, d_e, d_out, dtype = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    if strong and world < 2:
        raise SystemExit(f"bench.py: workload {args.workload} (8M nodes) needs >= 2 GPUs")
    if not strong:
        gargs = (gargs[0] * world, gargs[1])  # weak scaling: more rows, same row length
    mesh = getattr(meshgen, gen)(*gargs, device=dev)
    n_glob, E_glob = mesh["num_nodes"], int(mesh["indices"].numel())

    torch.manual_seed(0)
    model = MeshGraphNet(d_n, d_e, d_out).to(dev)
    g = torch.Generator().manual_seed(1)
    if world > 1:
        from modulus_b200.distributed import DistributedManager, mark_module_as_shared
        DistributedManager.initialize()
        dm = DistributedManager()
        dm.create_process_subgroup("graph_partition", world)
        gpart = None
        if args.partition == "stripes":
            # a partition with a REAL halo: mesh rows dealt to the ranks in stripes of --stripe-rows rows (every stripe has
            # two boundary rows whose sources live on the neighbouring ranks); nodewise slabs have only two such rows per rank
            from modulus_b200.models.gnn_layers import partition_graph_with_id_mapping
            row_of = torch.arange(n_glob, device=dev) // gargs[1]
            owner = (row_of // args.stripe_rows) % world
            gpart = partition_graph_with_id_mapping(mesh["offsets"], mesh["indices"], owner, owner, world, rank, dev)
        if n_glob // world * 6 > 12_500_000:  # memory-lean mode above 12.5M edges per rank: no stored h1
            from modulus_b200 import fused
            fused.KEEP_H1 = False
        graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n_glob, n_glob, partition_size=world,
                           partition_group_name="graph_partition", graph_partition=gpart)
        mark_module_as_shared(model, "graph_partition")
        gp = graph.dist_graph.graph_partition
        n_loc, E_loc = gp.num_local_dst_nodes, gp.num_local_indices
        halo_rows = int(gp.num_local_src_nodes - gp.sizes[rank][rank])
        hr = torch.tensor([halo_rows, n_loc], device=dev, dtype=torch.int64)
        hr_all = [torch.zeros_like(hr) for _ in range(world)]
        dist.all_gather(hr_all, hr)
        halo_all = [(int(t[0]), int(t[1])) for t in hr_all]
        ef_all = mesh["edge_features"][:, :d_e]
        ef_host = graph.get_edge_features_in_partition(ef_all).float().cpu().contiguous()
    else:
        graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n_glob, n_glob)
        n_loc, E_loc, halo_rows = n_glob, E_glob, 0
        halo_all = [(0, n_glob)]
        ef_src = mesh["edge_features"]
        ef_host = (ef_src[:, :d_e] if ef_src.shape[1] >= d_e else
                   torch.cat([ef_src, ef_src[:, :1].expand(-1, d_e - ef_src.shape[1])], 1)).float().cpu().contiguous()
    nf_host = torch.randn(n_loc, d_n, generator=g)
    tgt_host = torch.randn(n_loc, d_out, generator=g)
    nf_host, ef_host, tgt_host = nf_host.pin_memory(), ef_host.pin_memory(), tgt_host.pin_memory()
    del mesh
    graph.b200_plan()

    use_bf16 = dtype == "bf16"

    def step(nf, ef, tgt):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_bf16):
            pred = model(nf, ef, graph)
        loss = torch.nn.functional.mse_loss(pred.float(), tgt)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nf_d, ef_d, tgt_d = nf_host.to(dev), ef_host.to(dev), tgt_host.to(dev)

    # ---- warm-up (also the profiling pass that finds the dominant C-ABI entry point)
    for _ in range(max(args.warmup - 1, 2)):
        step(nf_d, ef_d, tgt_d)
    torch.cuda.synchronize()
    _lib.PROFILE.start(all_symbols=True, n_edges=E_loc)
    step(nf_d, ef_d, tgt_d)
    torch.cuda.synchronize()
    shares = _lib.PROFILE.stop()
    top = max(shares.items(), key=lambda kv: kv[1]["ms"])[0] if shares else None

    # ---- timed region A: inputs resident in HBM
    clocks = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else
                          int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]) if
                          os.environ["CUDA_VISIBLE_DEVICES"].replace(",", "").isdigit() else local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.start()
    l0 = lib.mgn_launch_count()
    _lib.PROFILE.start(all_symbols=False, only=top, n_edges=E_loc)
    ev0.record()
    for _ in range(args.steps):
        step(nf_d, ef_d, tgt_d)
    ev1.record()
    barrier()
    top_prof = _lib.PROFILE.stop()
    launches = (lib.mgn_launch_count() - l0) // max(args.steps, 1)
    t_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)  # ms
    clk = clocks.stop()
    ops.tc_check(dev)

    # ---- timed region B: end to end from pinned host buffers, loss read back every step
    from modulus_b200.prefetch import DevicePrefetcher
    pf = DevicePrefetcher(dev)
    pf.reserve(nf_host, ef_host, tgt_host)  # static device slots: allocation is not part of a step
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = 0.0
    # every step's inputs are copied from pinned host memory inside the timed region; the copy of step i+1 runs on
    # the prefetcher's stream while step i computes (modulus_b200/prefetch.py)
    pf.stage(nf_host, ef_host, tgt_host)
    for i in range(args.steps):
        nf, ef, tg = pf.take()
        if i + 1 < args.steps:
            pf.stage(nf_host, ef_host, tgt_host)
        loss = step(nf, ef, tg)
        pf.release()
        # the loss of every step lands in pinned host memory by an asynchronous copy that the host waits for one
        # step later: it queues step i+1 while the device runs step i instead of idling the device after a .item()
        loss_host[i % 2].copy_(loss.detach(), non_blocking=True)
        loss_ev[i % 2].record()
        if i > 0:
            loss_ev[(i - 1) % 2].synchronize()
            last = float(loss_host[(i - 1) % 2])
            if last != last:
                raise RuntimeError("bench.py: loss is NaN")
    loss_ev[(args.steps - 1) % 2].synchronize()
    last = float(loss_host[(args.steps - 1) % 2])
    e1.record()
    assert pf.h2d_bytes == args.steps * (nf_host.numel() + ef_host.numel() + tgt_host.numel()) * 4
    barrier()
    t_e2e = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    h2d = (nf_host.numel() + ef_host.numel() + tgt_host.numel()) * 4
    if last != last:
        raise RuntimeError("bench.py: loss is NaN")

    # ---- optimizer step, reported separately (the metric is forward + backward; train.py:111-123)
    opt_info = None
    if world == 1:
        from modulus_b200.optim import FusedAdam
        opt = FusedAdam(model.parameters(), lr=0.0)  # lr 0: the measured weights stay what the timed regions used
        for _ in range(3):
            opt.step()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l_opt = lib.mgn_launch_count()
        o0.record()
        for _ in range(10):
            opt.step()
        o1.record()
        torch.cuda.synchronize()
        n_par = sum(p.numel() for p in model.parameters() if p.requires_grad)
        opt_info = {"impl": "modulus_b200.optim.FusedAdam (mgn_adam_multi_step)", "ms_per_step": o0.elapsed_time(o1) / 10,
                    "launches_per_step": int((lib.mgn_launch_count() - l_opt) // 10), "parameters": n_par,
                    "tensors": sum(1 for p in model.parameters() if p.requires_grad),
                    "algorithmic_bytes": 28 * n_par}

    if rank != 0:
        return
    parity = side = None
    if not args.no_extra:
        del nf_d, ef_d, tgt_d
        parity = parity_check(dev)
        if world == 1 and args.workload == "c3":
            side = {"c2": side_rate("c2", dev, peaks)}
    b = 2 if use_bf16 else 4
    N1, E1 = n_glob // world, E_glob // world  # per-rank work (weak scaling)
    flops, bytes_ = step_flops(N1, E1, d_n, d_e, d_out), step_bytes(N1, E1, b)
    roof = None
    if top is not None and top in top_prof and top_prof[top]["calls"]:
        if top_prof[top]["bound"] is not None:
            # one entry point can serve several launch forms (mgn_edge_block_bwd_tc: E edge rows and, for the node block, N
            # node rows): the roofline is taken over the launches of the LARGEST form only (same algorithmic work per call)
            per = top_prof[top]["per_call"]
            wmax = max(w for _, w in per)
            big = [(ms, w) for ms, w in per if w >= 0.999 * wmax]
            avg_ms = sum(ms for ms, _ in big) / len(big)
            kind, amount = top_prof[top]["bound"], wmax
            if kind == "hbm":
                ach = amount / (avg_ms * 1e-3) / 1e9
                roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": ach / peaks["hbm"], "traffic": None}
            else:
                ach = amount / (avg_ms * 1e-3) / 1e12
                roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peaks["tc_sust"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tc_sust"], "traffic": None}
            roof["traffic"] = load_traffic(top, args.workload) if world == 1 else None
            if roof["traffic"] is not None:
                roof["traffic_launch"] = "one launch of the same form under ncu --set full (profiles/r02_ncu_traffic.json)"
            # the same launch against the OTHER roofline (the fused edge kernels sit at the ridge point): rows per
            # launch from the flop count, bytes per row = tensors crossing the kernel boundary
            per_row = {"mgn_edge_block_bwd_tc": (20.0 * H * H, 5 * H * 2), "mgn_edge_block_fwd_tc": (10.0 * H * H, 3 * H * 2)}
            if kind == "tensor" and top in per_row:
                rows_ = amount / per_row[top][0]
                gbs = rows_ * per_row[top][1] / (avg_ms * 1e-3) / 1e9
                roof["hbm_view"] = {"algorithmic_bytes": rows_ * per_row[top][1], "achieved": gbs, "unit": "GB/s",
                                    "peak": peaks["hbm"], "frac": gbs / peaks["hbm"]}
            roof["launches_timed"] = len(big)
            roof["algorithmic_per_launch"] = amount
            roof["avg_launch_ms"] = avg_ms
            roof["share_of_step"] = shares[top]["ms"] / max(sum(v["ms"] for v in shares.values()), 1e-9)
            roof["peak_source"] = peaks["src"]
    # whole-step fractions (report both)
    whole = {
        "hbm_frac": bytes_ / (t_step * 1e-3) / 1e9 / peaks["hbm"],
        "tensor_frac": flops / (t_step * 1e-3) / 1e12 / peaks["tc_sust"],
        "algorithmic_GB_per_step": bytes_ / 1e9, "algorithmic_TFLOP_per_step": flops / 1e12,
    }
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, t, n_s, E_s, kind, name, reps = cpu_step_rate(args.workload, SAMPLE_ROWS[args.workload], 4, 1, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": _cpu_desc(kind, name, n_s, E_s, t, reps, 1) + f" ({(reps + 1) * t:.0f} s of CPU work)"}
    line = {
        "metric": METRIC, "value": E_glob / (t_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": workload_config(args, world, n_glob, E_glob),
        "halo": {"rows_rank0": halo_rows, "rows_per_rank": [h for h, _ in halo_all],
                 "fraction_max": max(h / max(n, 1) for h, n in halo_all),
                 "transport": _halo_transport(graph) if world > 1 else "none"},
        "e2e": {"value": E_glob / (t_e2e * 1e-3), "unit": UNIT, "ms_per_step": t_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": roof, "whole_step": whole, "cpu_baseline": cpu, "clocks": clk, "optimizer_step": opt_info,
        "parity": parity, "other_configs": side,
        "kernel_shares": {k: round(v["ms"], 3) for k, v in sorted(shares.items(), key=lambda kv: -kv[1]["ms"])[:8]},
    }
    print(json.dumps(line), flush=True)


def _halo_transport(graph) -> str:
    """which transport the partitioned fused path used (after the first step has built its HaloContext)"""
    try:
        h = graph.b200_plan().extra.get("halo")
    except Exception:  # noqa: BLE001
        h = None
    if h is None:
        return "generic path (indexed_all_to_all_v over NCCL)"
    return ("peer memory (mgn_halo_push over NVLink, MGN_HALO_P2P=1)" if getattr(h, "peer", None) is not None
            else "NCCL all-to-all")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--partition", default="nodewise", choices=["nodewise", "stripes"],
                    help="N>1: contiguous slabs (reference default) or striped rows (a partition with a large halo)")
    ap.add_argument("--stripe-rows", type=int, default=25)
    ap.add_argument("--no-full-size", action="store_true", help="reference arm: skip the one full-size c2 step")
    ap.add_argument("--no-extra", action="store_true", help="skip the c2 side measurement and the in-run parity check")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"bench.py: --gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    if world > 1:
        import torch.distributed as dist
        import torch
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
