"""StaticCaptureTraining / StaticCaptureEvaluateNoGrad: the training (inference) step around the message-passing
path as ONE CUDA-graph launch.

Same decorator interface as the reference (physicsnemo/utils/capture.py:341-436 training, :437-515 evaluation;
used by the recipes as `@StaticCaptureTraining(model=..., optim=..., logger=...)`): the decorated function computes
the forward pass and returns the loss; calling it runs zero-grad, forward, backward and the optimizer step.  The
first `cuda_graph_warmup` calls run eagerly on a side stream, the next call records the graph, every later call
replays it -- inputs must therefore be static tensors the caller copies new data into, exactly as with the
reference.

B200-first differences
  * every kernel on the path is a C-ABI call that takes the stream explicitly, allocates nothing and never
    synchronises the host (include/mgn_b200.h), so the whole step is capturable by construction; no model
    metadata flag is consulted
  * with `modulus_b200.optim.FusedAdam` the optimizer step is recorded INSIDE the graph (device-resident step
    counter, learning rate read from a device scalar that follows the host-side scheduler), so a training step is
    one `cudaGraphLaunch`; any other optimizer steps after the replay, like the reference
  * AMP: bfloat16 needs no GradScaler.  `amp_type=torch.float16` (the reference's default) keeps the reference's
    protocol -- float16 autocast region, `GradScaler.scale(loss).backward()`, `scaler.step(optim)`, `scaler.update()`
    (capture.py:246-288) -- while the kernels compute in bf16 storage (models/gnn_layers/mesh_graph_mlp.compute_dtype);
    with FusedAdam the unscale and the skip-on-overflow run inside the optimizer kernel from the scaler's device
    tensors, so the scaled step stays ONE graph launch with no host read-back.  `compile=True` is refused (no tracing
    compiler here)
"""
from __future__ import annotations

import functools
import logging
from contextlib import nullcontext
from typing import Any, Callable, Optional

import torch

from .optim import FusedAdam


class _StaticCapture:
    _logger = logging.getLogger("capture")

    def __init__(self, model: torch.nn.Module, optim: Optional[torch.optim.Optimizer], logger, use_graphs: bool,
                 use_autocast: bool, compile: bool, cuda_graph_warmup: int, amp_type: torch.dtype,
                 gradient_clip_norm: Optional[float], label: Optional[str], eval_mode: bool):
        self.logger = logger if logger else self._logger
        if hasattr(model, "module") and isinstance(model.module, torch.nn.Module):  # DDP wrapper
            model = model.module
        if not isinstance(model, torch.nn.Module):
            self.logger.error("Model not a torch.nn.Module!")
            raise ValueError("Model not a torch.nn.Module!")
        if compile:
            raise ValueError("modulus_b200: compile=True is not supported (the hot path is hand-written CUDA)")
        if amp_type not in (torch.float16, torch.bfloat16):
            raise ValueError("AMP type must be torch.float16 or torch.bfloat16")  # capture.py:93-94
        self.model, self.optim = model, optim
        self.eval, self.no_grad = eval_mode, eval_mode
        self.gradient_clip_norm = gradient_clip_norm
        self.label = label
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("modulus_b200: the MeshGraphNet path runs on CUDA only; there is no CPU fallback")
        self.device = dev
        self.cuda_graphs_enabled = bool(use_graphs)
        self.use_autocast, self.amp_dtype = bool(use_autocast), amp_type
        # loss scaling exactly when the reference does it (capture.py:112-121): float16 autocast, training
        self.scaler = torch.amp.GradScaler("cuda", enabled=bool(use_autocast) and amp_type == torch.float16 and not eval_mode)
        self.optimizer_in_graph = (not eval_mode) and isinstance(optim, FusedAdam)
        self.replay_stream = torch.cuda.Stream(dev)
        self.graph = torch.cuda.CUDAGraph() if self.cuda_graphs_enabled else None
        self.output = None
        self.iteration = 0
        self.cuda_graph_warmup = max(int(cuda_graph_warmup), 1)  # >= 1: optimizer state must exist before recording

    # ------------------------------------------------------------------ pieces of one step
    def _zero_grads(self) -> None:
        """set_to_none, as the reference (capture.py:221-244): gradients re-created inside the graph's memory pool are
        static across replays and are overwritten, not accumulated, by each replay."""
        if self.no_grad:
            return
        if self.optim is not None:
            self.optim.zero_grad(set_to_none=True)
        self.model.zero_grad(set_to_none=True)

    def _amp_forward(self, *args: Any, **kwargs: Any):
        with torch.autocast("cuda", enabled=self.use_autocast, dtype=self.amp_dtype):
            output = self.function(*args, **kwargs)
        if not self.eval:
            self.scaler.scale(output).backward()
            if self.gradient_clip_norm is not None:
                self.scaler.unscale_(self.optim)
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.gradient_clip_norm)
        return output

    def _optim_step(self) -> None:
        self.scaler.step(self.optim)  # plain optim.step() when the scaler is disabled
        self.scaler.update()

    def _step(self, *args: Any, **kwargs: Any):
        output = self._amp_forward(*args, **kwargs)
        if self.optimizer_in_graph:
            self._optim_step()
        return output.detach()

    def _cuda_graph_step(self, *args: Any, **kwargs: Any) -> None:
        if self.iteration < self.cuda_graph_warmup:
            self.replay_stream.wait_stream(torch.cuda.current_stream(self.device))
            self._zero_grads()
            with torch.cuda.stream(self.replay_stream):
                self.output = self._step(*args, **kwargs)
            torch.cuda.current_stream(self.device).wait_stream(self.replay_stream)
        else:
            if self.optimizer_in_graph:
                self.optim.prepare_replay()
            if self.iteration == self.cuda_graph_warmup:
                self.logger.warning(f"Recording graph of '{self.function.__name__}'")
                self._zero_grads()
                torch.cuda.synchronize(self.device)
                with torch.cuda.graph(self.graph, stream=self.replay_stream):
                    self.output = self._step(*args, **kwargs)
                if self.optimizer_in_graph:
                    self.optim.prepare_replay()  # gradient addresses first seen while recording
            self.graph.replay()
        self.iteration += 1

    # ------------------------------------------------------------------ decorator
    def __call__(self, fn: Callable) -> Callable:
        self.function = fn

        @functools.wraps(fn)
        def decorated(*args: Any, **kwds: Any) -> Any:
            with torch.no_grad() if self.no_grad else nullcontext():
                if self.cuda_graphs_enabled:
                    self._cuda_graph_step(*args, **kwds)
                else:
                    self._zero_grads()
                    self.output = self._step(*args, **kwds)
                if not self.eval and not self.optimizer_in_graph and self.optim is not None:
                    self._optim_step()
            return self.output

        return decorated


class StaticCaptureTraining(_StaticCapture):
    """Decorator for a training step function that returns the loss (reference: capture.py:341-436)."""

    def __init__(self, model: torch.nn.Module, optim: torch.optim.Optimizer, logger=None, use_graphs: bool = True,
                 use_amp: bool = True, compile: bool = False, cuda_graph_warmup: int = 11,
                 amp_type: torch.dtype = torch.bfloat16, gradient_clip_norm: Optional[float] = None,
                 label: Optional[str] = None):
        super().__init__(model, optim, logger, use_graphs, use_amp, compile, cuda_graph_warmup, amp_type,
                         gradient_clip_norm, label, eval_mode=False)


class StaticCaptureEvaluateNoGrad(_StaticCapture):
    """Decorator for an inference step function that returns the prediction (reference: capture.py:437-515);
    the forward runs under `torch.no_grad()`."""

    def __init__(self, model: torch.nn.Module, logger=None, use_graphs: bool = True, use_amp: bool = True,
                 compile: bool = False, cuda_graph_warmup: int = 11, amp_type: torch.dtype = torch.bfloat16,
                 label: Optional[str] = None):
        super().__init__(model, None, logger, use_graphs, use_amp, compile, cuda_graph_warmup, amp_type, None, label,
                         eval_mode=True)
