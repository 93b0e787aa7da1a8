"""When do the AccumulateGrad node of a parameter and the root node of an eager step die?  (probe objects in the nodes'
metadata dicts print from __del__).  Run once plainly and once under compute-sanitizer."""
import gc

import torch

from modulus_b200.mesh import triangle_grid_mesh
from modulus_b200.models.gnn_layers import CuGraphCSC
from modulus_b200.models.meshgraphnet import MeshGraphNet
from modulus_b200.optim import FusedAdam

PHASE = ["start"]


class Probe:
    def __init__(self, name):
        self.name = name

    def __del__(self):
        print(f"    [{self.name}] destroyed during phase: {PHASE[0]}", flush=True)


def phase(s):
    PHASE[0] = s
    print("phase:", s, flush=True)


DEV = "cuda"
mesh = triangle_grid_mesh(20, 21, device=DEV)
n = mesh["num_nodes"]
graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n, n)
ef = mesh["edge_features"]
g = torch.Generator().manual_seed(4)
nf, tgt = torch.randn(n, 6, generator=g).to(DEV), torch.randn(n, 3, generator=g).to(DEV)
torch.manual_seed(7)
model = MeshGraphNet(6, 3, 3, processor_size=2).to(DEV)
opt = FusedAdam(model.parameters(), lr=1e-3)
params = dict(model.named_parameters())
watch = [k for k in params if k.endswith("model.0.weight")][:3] + [k for k in params if "node_decoder" in k][:2]
gc.collect()
gc.disable()
for it in range(2):
    phase(f"eager {it}: forward")
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model(nf, ef, graph).float(), tgt)
    loss.grad_fn.metadata["probe"] = Probe(f"root of eager step {it}")
    for k in watch:
        acc = params[k].view_as(params[k]).grad_fn.next_functions[0][0]
        if "probe" not in acc.metadata:
            acc.metadata["probe"] = Probe(f"AccumulateGrad {k} (seen in eager step {it})")
        del acc
    phase(f"eager {it}: backward")
    loss.backward()
    phase(f"eager {it}: optimizer")
    opt.step()
phase("del loss")
del loss
phase("synchronize")
torch.cuda.synchronize()
phase("gc.collect")
print("  gc freed", gc.collect())
phase("side-stream forward")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    loss = torch.nn.functional.mse_loss(model(nf, ef, graph).float(), tgt)
    phase("side-stream backward")
    loss.backward()
phase("end")
torch.cuda.synchronize()
