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
