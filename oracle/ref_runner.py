"""Run the staged, UNMODIFIED reference MeshGraphNet (oracle/_ref/physicsnemo + the import stand-ins of oracle/ref_shim)
on the host cores.  TEST / MEASUREMENT INFRASTRUCTURE: used by bench.py's `--impl reference` arm and cpu_baseline leg and
by tests; never by the product.  `available()` is False where oracle/_ref has not been staged (oracle/stage_reference.py).
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "MANIFEST.json"))


def _import():
    for p in (REF, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings

    warnings.filterwarnings("ignore")
    import dgl  # the stand-in
    from physicsnemo.models.meshgraphnet import MeshGraphNet  # the reference's own class

    return dgl, MeshGraphNet


def build(d_n: int, d_e: int, d_out: int, offsets, indices, processor_size: int = 15, seed: int = 0):
    """(model, graph): reference MeshGraphNet under `seed` and the DGL-style graph of a CSC (edge ids = CSC positions,
    as CuGraphCSC.to_dgl_graph builds it, gnn_layers/graph.py:459-477)."""
    import torch

    dgl, MeshGraphNet = _import()
    torch.manual_seed(seed)
    model = MeshGraphNet(d_n, d_e, d_out, processor_size=processor_size)
    deg = offsets[1:] - offsets[:-1]
    dst = torch.repeat_interleave(torch.arange(offsets.numel() - 1), deg)
    graph = dgl.graph((indices.long(), dst), num_nodes=int(offsets.numel() - 1))
    return model, graph


def step(model, graph, nf, ef, tgt):
    """zero_grad -> forward -> MSE -> backward (examples/cfd/vortex_shedding_mgn/train.py:151-166)"""
    import torch

    model.zero_grad(set_to_none=True)
    pred = model(nf, ef, graph)
    loss = torch.nn.functional.mse_loss(pred, tgt)
    loss.backward()
    return pred.detach(), loss.detach()


def time_steps(model, graph, nf, ef, tgt, reps: int, warmup: int):
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        step(model, graph, nf, ef, tgt)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)
