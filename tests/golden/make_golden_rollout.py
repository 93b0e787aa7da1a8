"""Golden vectors of the autoregressive rollout: produced by the reference's UNMODIFIED
`MGNRollout.predict` (examples/cfd/vortex_shedding_mgn/inference.py:90-150) driving the unmodified reference
MeshGraphNet and `VortexSheddingDataset.denormalize / normalize_node` (datapipes/gnn/vortex_shedding_dataset.py:334-355).

The script module imports hydra / omegaconf / matplotlib / dgl.dataloading at its top; none of them is used by
`predict`, so empty stand-in modules satisfy the imports (build container only):

    python tests/golden/make_golden_rollout.py

`predict` runs on a stand-in `self` that carries exactly the attributes it reads: `dataset` (node_stats + the two
static methods of the real dataset class), `dataloader` (a list of (graph, cells, mask) like GraphDataLoader yields),
`model`, `device`, `num_test_time_steps`."""
import importlib.util
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

import dgl  # noqa: E402  (the stand-in)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_stub("hydra", main=lambda **k: (lambda f: f), utils=_stub("hydra.utils", to_absolute_path=lambda p: p))
_stub("omegaconf", DictConfig=dict)
mpl = _stub("matplotlib")
for sub in ("pyplot", "animation", "tri", "patches"):
    setattr(mpl, sub, _stub(f"matplotlib.{sub}", Rectangle=object))
if not hasattr(dgl, "dataloading"):
    dgl.dataloading = _stub("dgl.dataloading", GraphDataLoader=object)
try:
    import physicsnemo.launch.logging  # noqa: F401
    import physicsnemo.launch.utils  # noqa: F401
except Exception:  # optional logging back ends missing: predict() uses neither
    _stub("physicsnemo.launch.logging", PythonLogger=object)
    _stub("physicsnemo.launch.utils", load_checkpoint=lambda *a, **k: None)

from physicsnemo.datapipes.gnn.vortex_shedding_dataset import VortexSheddingDataset  # noqa: E402
from physicsnemo.models.meshgraphnet import MeshGraphNet  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_inference", "/root/reference/examples/cfd/vortex_shedding_mgn/inference.py")
ref_inference = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_inference)

from modulus_b200.mesh import triangle_grid_mesh  # noqa: E402

torch.set_num_threads(8)
T, SEED = 6, 41
mesh = triangle_grid_mesh(9, 10)
n = mesh["num_nodes"]
offsets, indices = mesh["offsets"], mesh["indices"]
dst = torch.repeat_interleave(torch.arange(n), offsets[1:] - offsets[:-1])
gen = torch.Generator().manual_seed(SEED)
node_type = torch.nn.functional.one_hot(torch.randint(0, 4, (n,), generator=gen), 4).float()
frames = [torch.cat([torch.randn(n, 2, generator=gen), node_type], 1) for _ in range(T)]      # normalised node features
targets = [torch.randn(n, 3, generator=gen) for _ in range(T)]                                # graph.ndata["y"] (unused by pred)
mask = (node_type[:, 0] + node_type[:, 1] > 0).reshape(-1, 1)
stats = dict(velocity_mean=torch.tensor([0.3, -0.1]), velocity_std=torch.tensor([1.5, 0.7]),
             velocity_diff_mean=torch.tensor([0.01, 0.02]), velocity_diff_std=torch.tensor([0.05, 0.03]),
             pressure_mean=torch.tensor([0.2]), pressure_std=torch.tensor([2.0]))
ef = mesh["edge_features"].clone()

torch.manual_seed(SEED)
HIDDEN = dict(hidden_dim_processor=32, hidden_dim_node_encoder=32, hidden_dim_edge_encoder=32, hidden_dim_node_decoder=32)
model = MeshGraphNet(6, 3, 3, processor_size=2, **HIDDEN).eval()


def _graph(i):
    g = dgl.graph((indices.clone(), dst), num_nodes=n)
    g.ndata["x"] = frames[i].clone()
    g.ndata["y"] = targets[i].clone()
    g.edata["x"] = ef.clone()
    return g


me = types.SimpleNamespace(
    dataset=types.SimpleNamespace(node_stats={k: v.clone() for k, v in stats.items()},
                                  denormalize=VortexSheddingDataset.denormalize,
                                  normalize_node=VortexSheddingDataset.normalize_node),
    dataloader=[(_graph(i), mesh["cells"].unsqueeze(0) if "cells" in mesh else torch.zeros(1, 1, 3, dtype=torch.long), mask.clone())
                for i in range(T)],
    model=model, device="cpu", num_test_time_steps=T + 1)  # the dataset yields num_steps - 1 frames per trajectory
with torch.no_grad():
    ref_inference.MGNRollout.predict(me)
pred = torch.stack(me.pred)
exact = torch.stack(me.exact)
path = os.path.join(HERE, "ref_rollout.pt")
torch.save(dict(seed=SEED, steps=T, kwargs=dict(input_dim_nodes=6, input_dim_edges=3, output_dim=3, processor_size=2, **HIDDEN),
                offsets=offsets, indices=indices, n_nodes=n, frames=torch.stack(frames), edge_features=ef, mask=mask,
                stats=stats, pred=pred, exact=exact,
                state_dict={k: v.clone() for k, v in model.state_dict().items()}), path)
print(f"wrote {os.path.basename(path)}: {os.path.getsize(path) / 1024:.1f} KiB; pred {tuple(pred.shape)}, |pred| {pred.abs().mean():.4f}")
