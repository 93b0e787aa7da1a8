"""DevicePrefetcher: double-buffered staging of a step's host inputs into HBM on a dedicated copy stream.

The training step of the reference moves each sample with `graph.to(device)` right before the forward
(examples/cfd/vortex_shedding_mgn/train.py:144-146), i.e. the copy sits on the compute stream in front of every step.
Here the copy of step i+1 is issued from pinned host memory on its own stream while step i computes; the compute
stream only waits on the event of the slot it is about to read.  Device buffers are allocated once per slot, so the
inputs are static tensors (what a captured CUDA graph needs, see capture.py).

    pf = DevicePrefetcher(device)
    pf.stage(nf_host, ef_host, target_host)              # pinned host tensors
    for i in range(steps):
        nf, ef, target = pf.take()                       # device tensors of this step
        if i + 1 < steps:
            pf.stage(*next_host_batch)                   # overlaps with this step's kernels
        loss = training_step(nf, ef, target)
        pf.release()                                     # the slot may be overwritten once this step has run
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor


class DevicePrefetcher:
    def __init__(self, device, slots: int = 2):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("modulus_b200.prefetch: staging targets a CUDA device; there is no CPU fallback")
        if slots < 2:
            raise ValueError("DevicePrefetcher needs at least two slots")
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.buffers: List[Optional[Tuple[Tensor, ...]]] = [None] * slots
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free: List[Optional[torch.cuda.Event]] = [None] * slots
        self.staged = 0   # batches staged so far
        self.taken = 0    # batches handed out so far
        self.h2d_bytes = 0

    def reserve(self, *host: Tensor) -> None:
        """Allocate the device buffers of every slot for batches shaped like `host` (no copy), so that no step pays
        for a cudaMalloc."""
        for slot in range(len(self.buffers)):
            self.buffers[slot] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)

    def stage(self, *host: Tensor) -> None:
        if self.staged - self.taken >= len(self.buffers):
            raise RuntimeError("DevicePrefetcher: every slot holds a batch that has not been taken yet")
        slot = self.staged % len(self.buffers)
        bufs = self.buffers[slot]
        if bufs is None or len(bufs) != len(host) or any(b.shape != h.shape or b.dtype != h.dtype
                                                         for b, h in zip(bufs, host)):
            bufs = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            self.buffers[slot] = bufs
        if self.free[slot] is not None:
            self.stream.wait_event(self.free[slot])  # the step that last read this slot has finished
        else:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))  # fresh buffers: order after their allocation
        with torch.cuda.stream(self.stream):
            for b, h in zip(bufs, host):
                b.copy_(h, non_blocking=True)
                self.h2d_bytes += h.numel() * h.element_size()
            self.ready[slot].record(self.stream)
        self.staged += 1

    def take(self) -> Tuple[Tensor, ...]:
        if self.taken >= self.staged:
            raise RuntimeError("DevicePrefetcher: nothing staged")
        slot = self.taken % len(self.buffers)
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        self.taken += 1
        return self.buffers[slot]

    def release(self) -> None:
        """Mark the most recently taken slot as reusable once the work queued so far on the current stream has run."""
        slot = (self.taken - 1) % len(self.buffers)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev
