"""Diagnostics for mgn_mlp3_bwd_tc: per-tensor and per-row errors against an fp64 autograd reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_tc import _trick_case, bf, dev_params, DEV, st_round
from modulus_b200 import ops

def run(M, use_go2=True, add_gout=True):
    A, P, src, dst, go1, go2, p = _trick_case(M, seed=100 + M)
    leaves = {k: bf(v).double().requires_grad_(True) if k in ("w1", "w2", "w3") else v.double().requires_grad_(True)
              for k, v in p.items()}
    A64 = A.double().requires_grad_(True)
    w1e = leaves["w1"][:, :128]
    Gs = (P[src][:, :128].double() + P[dst][:, 128:256].double()).requires_grad_(True)
    z1 = A64 @ w1e.T + Gs + leaves["b1"]
    h1 = st_round(F.relu(z1)); h1.retain_grad()
    z2 = h1 @ leaves["w2"].T + leaves["b2"]
    h2 = st_round(F.relu(z2)); h2.retain_grad()
    y = h2 @ leaves["w3"].T + leaves["b3"]; y.retain_grad()
    out = F.layer_norm(y, (128,), leaves["gamma"], leaves["beta"], 1e-5) + (A64 if add_gout else 0)
    gout = bf(go1 + go2[dst]).double() if use_go2 else go1.double()
    (out * gout).sum().backward()
    d = dev_params(p)
    Ad, Pd = A.to(DEV).bfloat16(), P.to(DEV).bfloat16()
    gw1 = torch.zeros(128, 384, device=DEV)
    gw2, gw3 = torch.empty(128, 128, device=DEV), torch.empty(128, 128, device=DEV)
    gb1, gb2, gb3, gga, gbe = (torch.empty(128, device=DEV) for _ in range(5))
    g_a, g_z1 = ops.mlp3_bwd_tc(Ad, None, None, Pd, src.to(DEV).int(), 0, Pd, dst.to(DEV).int(), 128,
                                go1.to(DEV).bfloat16(), go2.to(DEV).bfloat16() if use_go2 else None,
                                dst.to(DEV).int() if use_go2 else None, M,
                                d["w1"][:, :128], d["b1"], d["w2"], d["b2"], d["w3"], d["b3"], d["gamma"], 128, 1e-5,
                                True, add_gout, True, gw1[:, :128], gb1, gw2, gb2, gw3, gb3, gga, gbe)
    torch.cuda.synchronize()
    print(f"--- M={M} go2={use_go2} add_gout={add_gout} status={int(ops.tc_status(DEV).item())}")
    ops.tc_status(DEV).zero_()
    def rep(name, got, ref):
        got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
        err = (got - ref).abs()
        print(f"{name:8s} rel={float(err.max() / ref.abs().max().clamp_min(1e-30)):.3e} maxabs={float(err.max()):.3e}", end="")
        if got.dim() == 2 and got.shape[0] == M:
            rows = err.max(dim=1).values
            bad = (rows > 0.05 * ref.abs().max()).nonzero().flatten()
            print(f"  bad rows: {bad.numel()} first {bad[:12].tolist()} last {bad[-4:].tolist()}", end="")
        print()
    rep("g_a", g_a.float(), A64.grad)
    rep("g_z1", g_z1.float(), Gs.grad)
    rep("gw1", gw1[:, :128], leaves["w1"].grad[:, :128])
    rep("gw2", gw2, leaves["w2"].grad)
    rep("gw3", gw3, leaves["w3"].grad)
    rep("gb1", gb1, leaves["b1"].grad); rep("gb2", gb2, leaves["b2"].grad); rep("gb3", gb3, leaves["b3"].grad)
    rep("ggamma", gga, leaves["gamma"].grad); rep("gbeta", gbe, leaves["beta"].grad)

for M in [int(a) for a in sys.argv[1:]] or [1, 2, 128, 129, 130, 256, 1000]:
    run(M)
run(130, use_go2=False)
run(130, use_go2=False, add_gout=False)
