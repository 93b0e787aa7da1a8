"""Optimizer step and whole training step around the message-passing path.

  (a) optimizer step over the default MeshGraphNet's 263 parameter tensors (2.33 M parameters):
      torch.optim.Adam (foreach), torch.optim.Adam(fused=True) and modulus_b200.optim.FusedAdam (one launch)
  (b) c1 (1.9 k nodes, launch-bound) and c2 training step = zero_grad + forward + MSE + backward + FusedAdam.step,
      eager vs modulus_b200.capture.StaticCaptureTraining (one CUDA graph launch per step)

    python tools/bench_train_step.py [reps=20]      ->  markdown on stdout
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulus_b200 import _lib
from modulus_b200.mesh import triangle_grid_mesh
from modulus_b200.models.gnn_layers import CuGraphCSC
from modulus_b200.models.meshgraphnet import MeshGraphNet
from modulus_b200.optim import FusedAdam
from modulus_b200.capture import StaticCaptureTraining

DEV = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20


def timed(fn, n=reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("## (a) optimizer step, default MeshGraphNet(11,4,4): 15 layers, hidden 128\n")
print("| optimizer | tensors | parameters | ms per step | launches per step |")
print("|---|---:|---:|---:|---:|")
for name, make in (("torch.optim.Adam (foreach)", lambda ps: torch.optim.Adam(ps, lr=1e-3)),
                   ("torch.optim.Adam(fused=True)", lambda ps: torch.optim.Adam(ps, lr=1e-3, fused=True)),
                   ("modulus_b200.optim.FusedAdam", lambda ps: FusedAdam(ps, lr=1e-3))):
    torch.manual_seed(0)
    model = MeshGraphNet(11, 4, 4).to(DEV)
    ps = [p for p in model.parameters() if p.requires_grad]
    for p in ps:
        p.grad = torch.randn_like(p)
    opt = make(ps)
    ms = timed(opt.step)
    l0 = _lib.load().mgn_launch_count()
    opt.step()
    launches = _lib.load().mgn_launch_count() - l0 if "modulus" in name else "n/a"
    print(f"| {name} | {len(ps)} | {sum(p.numel() for p in ps)} | {ms:.3f} | {launches} |")

print("\n## (b) training step (zero_grad + forward + MSE + backward + FusedAdam.step), eager vs one CUDA graph\n")
print("| workload | nodes | edges | dtype | eager ms | graph ms | speed-up |")
print("|---|---:|---:|---|---:|---:|---:|")
for wname, (nx, ny), bf16 in (("c1", (42, 45), False), ("c2", (316, 317), True)):
    mesh = triangle_grid_mesh(nx, ny, device=DEV)
    n, E = mesh["num_nodes"], int(mesh["indices"].numel())
    graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n, n)
    torch.manual_seed(0)
    model = MeshGraphNet(6, 3, 3).to(DEV)
    opt = FusedAdam(model.parameters(), lr=1e-4)
    nf, ef, tgt = torch.randn(n, 6, device=DEV), mesh["edge_features"], torch.randn(n, 3, device=DEV)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            out = model(nf, ef, graph)
        loss = torch.nn.functional.mse_loss(out.float(), tgt)
        loss.backward()
        opt.step()
        return loss

    t_eager = timed(step)

    @StaticCaptureTraining(model=model, optim=opt, use_amp=bf16, cuda_graph_warmup=2)
    def captured(nf, tgt):
        return torch.nn.functional.mse_loss(model(nf, ef, graph).float(), tgt)

    t_graph = timed(lambda: captured(nf, tgt))
    print(f"| {wname} | {n} | {E} | {'bf16' if bf16 else 'f32'} | {t_eager:.3f} | {t_graph:.3f} | {t_eager / t_graph:.2f}x |")

print("\n## (c) step before the path at the c3 size: edge features of a 1 M-node / 6 M-edge surface mesh (fp32)\n")
from modulus_b200.mesh import edge_features, graph_from_cells, torus_surface_mesh
from modulus_b200.ops import GraphPlan
import json
peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
PEAK = json.load(open(peaks_path)).get("hbm_gbs", 6550.0) if os.path.exists(peaks_path) else 6550.0
mesh = torus_surface_mesh(1000, 1000, device=DEV)
n, E = mesh["num_nodes"], int(mesh["indices"].numel())
plan = GraphPlan.from_csc(mesh["offsets"], mesh["indices"], n, n)
pos = mesh["coords"]
mu, sd = mesh["edge_features"].mean(0), mesh["edge_features"].std(0)


def torch_expr():
    d = pos[plan.src.long()] - pos[plan.dst.long()]
    return (torch.cat((d, torch.linalg.norm(d, dim=-1, keepdim=True)), 1) - mu) / sd


t_k = timed(lambda: edge_features(pos, plan.src, plan.dst, mu, sd))
t_t = timed(torch_expr)
alg = E * (2 * 4 + 4 * 4) + n * 12  # endpoint ids + output row per edge, coordinate table once
print("| implementation | ms | algorithmic MB | GB/s | frac of HBM peak |")
print("|---|---:|---:|---:|---:|")
print(f"| mgn_edge_features (one launch) | {t_k:.4f} | {alg / 1e6:.1f} | {alg / t_k / 1e6:.0f} | {alg / t_k / 1e6 / PEAK:.2f} |")
print(f"| torch expression (gather, sub, norm, cat, normalise) | {t_t:.4f} | {alg / 1e6:.1f} | {alg / t_t / 1e6:.0f} | {alg / t_t / 1e6 / PEAK:.2f} |")
# cells -> graph (sort + unique + scan on the device)
i = torch.arange(1000, device=DEV).view(-1, 1).expand(1000, 1000)
j = torch.arange(1000, device=DEV).view(1, -1).expand(1000, 1000)
a, b, c, d = (i * 1000 + j), (i * 1000 + (j + 1) % 1000), (((i + 1) % 1000) * 1000 + j), (((i + 1) % 1000) * 1000 + (j + 1) % 1000)
cells = torch.cat([torch.stack([a, b, d], -1).reshape(-1, 3), torch.stack([a, d, c], -1).reshape(-1, 3)])
t_g = timed(lambda: graph_from_cells(cells, n), n=5, warm=1)
off, idx = graph_from_cells(cells, n)
print(f"\ngraph_from_cells: {cells.shape[0]} triangles -> {idx.numel()} directed edges in {t_g:.2f} ms "
      f"(same CSC as the generator: {bool(torch.equal(off, mesh['offsets']) and torch.equal(idx, mesh['indices']))})")
