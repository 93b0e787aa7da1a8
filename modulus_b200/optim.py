"""FusedAdam: the optimizer step after the message-passing path as ONE kernel launch.

The reference's recipe uses `apex.optimizers.FusedAdam` when installed and `torch.optim.Adam` otherwise
(examples/cfd/vortex_shedding_mgn/train.py:111-123), with `GradScaler.step` deciding whether the step is
skipped under AMP (:161-163).  A default MeshGraphNet has 262 parameter tensors (263 state_dict entries with its
one buffer); torch's foreach Adam
issues ~10 launches over them, the single-tensor path ~2 000.  Here a device-resident pointer table
(parameters, gradients, both moments) and a chunk map are built once and `mgn_adam_multi_step` updates
every tensor in one launch (include/mgn_b200.h).  The step counter lives on the device, so the call never
synchronises and can sit inside a captured CUDA graph.

Constructor arguments follow `torch.optim.Adam` (plus apex's `adam_w_mode`); `state_dict()` has the layout of
`torch.optim.Adam` (`step`, `exp_avg`, `exp_avg_sq` per parameter), so checkpoints move both ways.
There is no CPU path: parameters must be fp32 CUDA tensors.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib

CHUNK_ELEMS = 4096  # elements per CTA (256 threads x 4 float4)


class _Table:
    """Device pointer table of one parameter group."""

    def __init__(self, params: List[Tensor], state, device):
        self.params = params
        n = len(params)
        self.numel_host = [p.numel() for p in params]
        chunk_tensor, chunk_start = [], []
        for i, ne in enumerate(self.numel_host):
            for s in range(0, ne, CHUNK_ELEMS):
                chunk_tensor.append(i)
                chunk_start.append(s)
        self.n_chunks = len(chunk_tensor)
        self.numel = torch.tensor(self.numel_host, dtype=torch.int64, device=device)
        self.chunk_tensor = torch.tensor(chunk_tensor, dtype=torch.int32, device=device)
        self.chunk_start = torch.tensor(chunk_start, dtype=torch.int64, device=device)
        # rows: 0 params, 1 grads, 2 exp_avg, 3 exp_avg_sq
        # ring of pinned staging buffers: the upload is asynchronous, so a buffer is rewritten only after the copy
        # that read it has run (its event), which lets the host stay several steps ahead of the device
        self.ring = [(torch.zeros((4, max(n, 1)), dtype=torch.int64).pin_memory(), torch.cuda.Event())
                     for _ in range(4)]
        self.ring_pos = 0
        self.pending_rows = None  # pointer rows seen during a graph capture, uploaded by prepare_replay()
        self.lr_dev = torch.zeros((), dtype=torch.float32, device=device)  # learning rate a captured step reads
        self.lr_host: Optional[float] = None
        self.ptrs = torch.zeros((4, max(n, 1)), dtype=torch.int64, device=device)
        self.key: Optional[Tuple[int, ...]] = None
        self.state = state

    def refresh(self) -> None:
        """Re-upload the pointer rows when any tensor moved (e.g. `zero_grad(set_to_none=True)` makes autograd
        allocate fresh gradient tensors every step).  One 8 KB async copy from a pinned ring buffer; skipped when
        nothing changed, which is what a captured graph needs (static gradients)."""
        rows = ([p.data_ptr() for p in self.params],
                [0 if p.grad is None else p.grad.data_ptr() for p in self.params],
                [self.state[p]["exp_avg"].data_ptr() for p in self.params],
                [self.state[p]["exp_avg_sq"].data_ptr() for p in self.params])
        key = tuple(rows[0] + rows[1] + rows[2] + rows[3])
        if key == self.key:
            return
        if torch.cuda.is_current_stream_capturing():
            # gradients freed before the capture are re-created inside the graph's memory pool, so their addresses
            # are only known now.  Nothing is recorded for them: the captured kernel reads the device table at replay
            # time, and prepare_replay() uploads these rows eagerly after the capture and before the first replay.
            self.pending_rows = rows
            self.key = key
            return
        self.upload(rows)
        self.key = key

    def upload(self, rows) -> None:
        host, event = self.ring[self.ring_pos % len(self.ring)]
        if self.ring_pos >= len(self.ring):
            event.synchronize()
        self.ring_pos += 1
        host.copy_(torch.tensor(rows, dtype=torch.int64))
        self.ptrs.copy_(host, non_blocking=True)
        event.record()
        self.pending_rows = None


class FusedAdam(torch.optim.Optimizer):
    # torch.amp.GradScaler.step() hands optimizers that declare this their `grad_scale` / `found_inf` device tensors and
    # does not read anything back: unscaling and the skip-on-overflow then happen inside mgn_adam_multi_step
    _step_supports_amp_scaling = True

    def __init__(
        self,
        params: Iterable,
        lr: float = 1e-3,
        betas: Tuple[float, float] = (0.9, 0.999),
        eps: float = 1e-8,
        weight_decay: float = 0.0,
        amsgrad: bool = False,
        adam_w_mode: bool = False,
    ):
        if amsgrad:
            raise RuntimeError("FusedAdam does not support the AMSGrad variant.")  # as apex
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adam_w_mode=adam_w_mode)
        super().__init__(params, defaults)
        self._tables: dict = {}

    # ------------------------------------------------------------------ state
    def _init_group(self, gi: int, group) -> _Table:
        params = [p for p in group["params"] if p.requires_grad]
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("modulus_b200.optim.FusedAdam: parameters must be contiguous float32 CUDA tensors "
                                   f"(got {p.dtype} on {p.device}); there is no CPU fallback")
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        dev = params[0].device if params else torch.device("cuda")
        tab = _Table(params, self.state, dev)
        # one device counter per group; every state[p]["step"] aliases it so state_dict() keeps torch's layout
        step0 = self.state[params[0]]["step"] if params else torch.zeros((), device=dev)
        tab.step = torch.as_tensor(step0).detach().to(device=dev, dtype=torch.float32).clone().reshape(())
        for p in params:
            self.state[p]["step"] = tab.step
        self._tables[gi] = tab
        return tab

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._tables.clear()  # moments were replaced: rebuild the tables (and re-alias the step counters)

    def add_param_group(self, param_group) -> None:
        super().add_param_group(param_group)
        if hasattr(self, "_tables"):
            self._tables.clear()

    def prepare_replay(self) -> None:
        """Host-side bookkeeping around a captured step; call it before capturing and before every replay.  Mirrors
        each group's host-side learning rate (what LR schedulers update) into the device scalar a captured step
        reads -- one tiny fill per group and only when the value changed -- and uploads pointer rows that were first
        seen during the capture (gradients re-created inside the graph's memory pool)."""
        for gi, group in enumerate(self.param_groups):
            tab = self._tables.get(gi)
            if tab is None:
                tab = self._init_group(gi, group)
            lr = group["lr"]
            if not isinstance(lr, Tensor) and tab.lr_host != float(lr):
                tab.lr_dev.fill_(float(lr))
                tab.lr_host = float(lr)
            if tab.pending_rows is not None and not torch.cuda.is_current_stream_capturing():
                tab.upload(tab.pending_rows)

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, found_inf: Optional[Tensor] = None, inv_scale: Optional[Tensor] = None):
        """One Adam update of every group.  `found_inf` / `inv_scale` are optional 1-element fp32 CUDA tensors:
        the step is skipped on the device when found_inf != 0 and gradients are multiplied by inv_scale
        (what `GradScaler.unscale_` + `GradScaler.step` do in two passes and one host read-back).  Called
        through `GradScaler.step(optimizer)` the scaler supplies both as attributes (`_step_supports_amp_scaling`):
        no host synchronisation, CUDA-graph capturable (the reference recipe's AMP switch: float16 autocast +
        GradScaler, examples/cfd/vortex_shedding_mgn/train.py:153-166)."""
        gs, fi = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        if gs is not None and inv_scale is None:
            inv_scale = gs.to(torch.float32).reshape(1).reciprocal()
        if fi is not None:
            fi = fi.to(torch.float32).reshape(1)
            found_inf = fi if found_inf is None else torch.maximum(found_inf.to(torch.float32).reshape(1), fi)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:  # flags a torch.optim.Adam checkpoint may carry into the groups
            if group.get("amsgrad", False) or group.get("maximize", False):
                raise RuntimeError("FusedAdam does not support amsgrad / maximize (found in a parameter group)")
        tables = [self._tables.get(gi) or self._init_group(gi, group) for gi, group in enumerate(self.param_groups)]
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        for tab, group in zip(tables, self.param_groups):
            if not tab.params:
                continue
            # a fused tensor-core kernel that reported a pipeline stall left partial gradients behind: such a step is
            # skipped ON THE DEVICE (same mechanism as an inf/nan under GradScaler), and ops.tc_poll raises soon after
            from . import ops

            stalled = ops.tc_found_inf(tab.params[0].device)
            found_inf = stalled if found_inf is None else torch.maximum(found_inf.to(torch.float32).reshape(1), stalled)
            tab.refresh()
            lr = group["lr"]
            lr_dev = lr if isinstance(lr, Tensor) else None
            if lr_dev is None and torch.cuda.is_current_stream_capturing():
                # a by-value learning rate would be frozen into the graph: read it from the device instead;
                # `prepare_replay()` (called by StaticCaptureTraining before each replay) follows the scheduler
                if tab.lr_host != float(lr):
                    raise RuntimeError("FusedAdam: call prepare_replay() before capturing a step")
                lr_dev = tab.lr_dev
            b1, b2 = group["betas"]
            rc = lib.mgn_adam_multi_step(
                tab.ptrs[0].data_ptr(), tab.ptrs[1].data_ptr(), tab.ptrs[2].data_ptr(), tab.ptrs[3].data_ptr(),
                tab.numel.data_ptr(), tab.chunk_tensor.data_ptr(), tab.chunk_start.data_ptr(), tab.n_chunks,
                CHUNK_ELEMS, 0.0 if lr_dev is not None else float(lr), float(b1), float(b2), float(group["eps"]),
                float(group["weight_decay"]), int(bool(group.get("adam_w_mode", group.get("decoupled_weight_decay", False)))), tab.step.data_ptr(),
                None if lr_dev is None else lr_dev.data_ptr(),
                None if inv_scale is None else inv_scale.data_ptr(),
                None if found_inf is None else found_inf.data_ptr(), stream)
            if rc != 0:
                raise _lib.MGNError(f"mgn_adam_multi_step failed: {lib.mgn_error_string(rc).decode()} ({rc})")
        return loss
