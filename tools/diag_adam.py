"""Diagnostic for the FusedAdam-vs-torch.optim.Adam comparison: repeats the first step of
tests/test_train_step.py::test_fused_adam_matches_torch_adam and, where the two differ by more than the test's
tolerance, prints the element, its gradient and both results next to a float64 evaluation of the same update."""
import sys

import torch

from modulus_b200.optim import FusedAdam

DEV = "cuda"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
bad = 0
for rep in range(reps):
    torch.manual_seed(1)
    shapes = [(128, 384), (128,), (3, 5), (1,), (4099,), (128, 128)]
    ours = [torch.nn.Parameter(torch.randn(*s, device=DEV)) for s in shapes]
    p0 = [p.detach().cpu().clone() for p in ours]
    ref = [torch.nn.Parameter(p.clone()) for p in p0]
    opt = FusedAdam(ours, lr=3e-3)
    ropt = torch.optim.Adam(ref, lr=3e-3, foreach=False)
    grads = []
    for p, r in zip(ours, ref):
        gr = torch.randn(*r.shape)
        grads.append(gr)
        r.grad = gr
        p.grad = gr.to(DEV)
    opt.step()
    ropt.step()
    for i, (p, r) in enumerate(zip(ours, ref)):
        got, want = p.detach().cpu(), r.detach()
        g64, p64 = grads[i].double(), p0[i].double()
        m = 0.1 * g64
        v = 0.001 * g64 * g64
        truth = p64 - (3e-3 / 0.1) * m / (v.sqrt() / (0.001 ** 0.5) + 1e-8)
        tol = 2e-7 + 2e-6 * want.abs()
        mask = (got - want).abs() > tol
        if mask.any():
            bad += 1
            idx = mask.flatten().nonzero().flatten()[:4].tolist()
            for k in idx:
                print(f"rep {rep} tensor {i} elem {k}: p0 {p0[i].flatten()[k]:.9g} g {grads[i].flatten()[k]:.9g} "
                      f"ours {got.flatten()[k]:.9g} torch {want.flatten()[k]:.9g} fp64 {truth.flatten()[k]:.12g}")
            print(f"  rep {rep} tensor {i}: {int(mask.sum())} elements out; ours-vs-fp64 max {float((got.double() - truth).abs().max()):.3g}, "
                  f"torch-vs-fp64 max {float((want.double() - truth).abs().max()):.3g}")
print(f"{bad} tensor comparisons out of tolerance in {reps} repetitions; torch threads {torch.get_num_threads()}")
