"""Autoregressive rollout of a trained MeshGraphNet on a stationary mesh, kept on the device.

The reference's inference loop (examples/cfd/vortex_shedding_mgn/inference.py:90-150) feeds each prediction back as
the next input: normalise the velocity columns, run the model, de-normalise the predicted velocity difference and
pressure, zero the update on wall / outflow nodes, integrate -- and moves every frame to the host (`.cpu()`), which
synchronises once per time step.  Here the whole trajectory stays in HBM (`[steps, N, 3]`), the per-step work is one
CUDA-graph launch (`capture.StaticCaptureEvaluateNoGrad` around normalise -> model -> de-normalise -> mask ->
integrate) plus two small device copies, and nothing synchronises until the caller reads the result.

`model` is any callable `(node_features, edge_features, graph) -> [N, 3]`; with a `modulus_b200` MeshGraphNet on CUDA
`use_graphs=True` records the step once and replays it.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch
from torch import Tensor

REQUIRED_STATS = ("velocity_mean", "velocity_std", "velocity_diff_mean", "velocity_diff_std", "pressure_mean",
                  "pressure_std")


def rollout(model: Callable, graph, node_features0: Tensor, edge_features: Tensor, update_mask: Tensor,
            stats: Dict[str, Tensor], steps: int, use_graphs: bool = False, use_amp: bool = False,
            cuda_graph_warmup: int = 1) -> Tensor:
    """Trajectory `[steps, N, 3]` = (u, v, p) per node and step, de-normalised.

    node_features0  [N, d]  first frame as the dataset stores it: columns 0:2 NORMALISED velocity, 2: static columns
                            (one-hot node type)                                   (inference.py:113-121)
    update_mask     [N, 1] or [N]  True where the velocity is integrated; wall-boundary and outflow nodes keep
                            their value                                            (inference.py:135-139)
    stats           the dataset's node statistics (`VortexSheddingDataset.node_stats`), each `[1, k]` or `[k]`
    """
    for k in REQUIRED_STATS:
        if k not in stats:
            raise KeyError(f"rollout: missing statistic '{k}'")
    if steps < 0:
        raise ValueError("steps must be >= 0")
    dev = node_features0.device
    st = {k: stats[k].to(dev).reshape(1, -1).float() for k in REQUIRED_STATS}
    mask = update_mask.to(dev).reshape(-1, 1).bool()
    static_cols = node_features0[:, 2:].float()
    n = node_features0.shape[0]
    # de-normalised velocity of the current frame: the state that is integrated (inference.py:96-98)
    vel = (node_features0[:, 0:2].float() * st["velocity_std"] + st["velocity_mean"]).contiguous()
    out = torch.empty((steps, n, 3), dtype=torch.float32, device=dev)

    def step(vel_now: Tensor) -> Tensor:
        invar = torch.cat(((vel_now - st["velocity_mean"]) / st["velocity_std"], static_cols), dim=1)
        pred = model(invar, edge_features, graph).float()
        dv = pred[:, 0:2] * st["velocity_diff_std"] + st["velocity_diff_mean"]
        p = pred[:, 2:3] * st["pressure_std"] + st["pressure_mean"]
        dv = torch.where(mask, dv, torch.zeros_like(dv))
        return torch.cat((vel_now + dv, p), dim=1)

    if use_graphs or use_amp:
        from .capture import StaticCaptureEvaluateNoGrad

        if not isinstance(model, torch.nn.Module):
            raise ValueError("rollout: use_graphs / use_amp need the model as a torch.nn.Module")
        step = StaticCaptureEvaluateNoGrad(model=model, use_graphs=use_graphs, use_amp=use_amp,
                                           cuda_graph_warmup=cuda_graph_warmup)(step)
    with torch.no_grad():
        for i in range(steps):
            frame = step(vel)              # static input `vel`; the result buffer is rewritten by every replay
            out[i].copy_(frame)
            vel.copy_(frame[:, 0:2])
    return out
