"""Diagnostic: capture one fp32 training step in a CUDA graph and print the FIRST exception raised inside the capture."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulus_b200.mesh import triangle_grid_mesh
from modulus_b200.models.gnn_layers import CuGraphCSC
from modulus_b200.models.meshgraphnet import MeshGraphNet
from modulus_b200.optim import FusedAdam

DEV = "cuda:0"
bf16 = len(sys.argv) > 1 and sys.argv[1] == "bf16"
mesh = triangle_grid_mesh(42, 45, device=DEV)
n = mesh["num_nodes"]
graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n, n)
torch.manual_seed(0)
model = MeshGraphNet(6, 3, 3, processor_size=2).to(DEV)
opt = FusedAdam(model.parameters(), lr=1e-4)
nf, ef, tgt = torch.randn(n, 6, device=DEV), mesh["edge_features"], torch.randn(n, 3, device=DEV)


def step(stage):
    try:
        opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            out = model(nf, ef, graph)
        print(stage, "forward ok", flush=True)
        loss = torch.nn.functional.mse_loss(out.float(), tgt)
        loss.backward()
        print(stage, "backward ok", flush=True)
        opt.step()
        print(stage, "optimizer ok", flush=True)
    except Exception:
        traceback.print_exc()
        raise


for _ in range(2):
    step("eager")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step("side-stream")
torch.cuda.current_stream().wait_stream(s)
cg = torch.cuda.CUDAGraph()
with torch.cuda.graph(cg):
    step("capture")
cg.replay()
torch.cuda.synchronize()
print("capture + replay ok")
