"""GraphCast / AeroGraphNet regime: one bf16 MeshGraphMLP (3H -> H -> H, SiLU, LayerNorm) forward + backward at hidden 512 /
256, generic bf16 path with the K-looped tcgen05 GEMM (mgn_gemm_bf16_tc + mgn_wgrad_tc blocks) vs the fp32-accurate SIMT
kernels (ops.WIDE_TC = False).  python tools/bench_wide_mlp.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulus_b200 import ops
from modulus_b200.models.gnn_layers import MeshGraphMLP
DEV = "cuda:0"
torch.manual_seed(0)
print("| hidden | rows | path | ms fwd+bwd | useful TFLOP/s |\n|---:|---:|---|---:|---:|")
for hidden, M in ((512, 327660), (256, 400000)):  # GraphCast level-6 mesh edges; AeroGraphNet-size encoder
    mlp = MeshGraphMLP(3 * hidden, hidden, hidden, 1, activation_fn=torch.nn.SiLU()).to(DEV)
    x0 = torch.randn(M, 3 * hidden, device=DEV).bfloat16()
    for wide in (True, False):
        ops.WIDE_TC = wide

        def step():
            mlp.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = mlp(x)
            y.float().sum().backward()

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 3 * 2 * M * (3 * hidden * hidden + hidden * hidden)
        print(f"| {hidden} | {M} | {'tcgen05 K-looped GEMM' if wide else 'SIMT (fp32-accurate)'} | {ms:.2f} | {fl / ms / 1e9:.0f} |")
ops.WIDE_TC = True
ops.tc_check(DEV)
