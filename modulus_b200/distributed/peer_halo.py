"""Halo exchange over peer memory (NVLink / NVSwitch) for the partitioned fused path.

The reference moves halo rows with `indexed_all_to_all_v` (physicsnemo/distributed/utils.py:541-765): an index gather into a
send buffer, an NCCL all-to-all, and a copy of the received rows into place.  On one NVSwitch node every GPU can store into
every other GPU's memory, so here the gather kernel writes each row straight into its final place in the RECEIVER's buffer
(`mgn_halo_push`, csrc/mgn_halo.cu) and publishes an epoch flag; the receiver waits for the flags right before the launches
that read the rows (`mgn_halo_wait`).  One launch per rank, layer and direction instead of pack + collective + copy.

Buffers come from `torch.distributed._symmetric_memory` (one allocation per graph: two parities x two directions + flags).
Double buffering by call parity is safe because the exchange pattern is symmetric (checked): a rank waits for its peer's
flag of call e + 1 -- which the peer publishes only after the launches that read call e's rows -- before its own push of call
e + 2 can overwrite them.  Opt-in (`MGN_HALO_P2P=1`); anything that does not fit (another backend, ranks on several nodes,
an asymmetric pattern, no symmetric memory) falls back to the NCCL protocol of `HaloContext`."""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

_FLAG_SLOTS = 16  # = csrc/mgn_halo.cu kMaxPeers


def requested() -> bool:
    return os.environ.get("MGN_HALO_P2P", "0") == "1"


def _i64(vals: Sequence[int]):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


class PeerHalo:
    """Forward: rows `table[send_idx]` -> the peers' forward buffers.  Backward: this rank's halo-row gradients (already
    ordered by owner) -> the owners' backward buffers (rows in the order the owner sent them)."""

    def __init__(self, group, rank: int, world: int, send_splits: Sequence[int], recv_splits: Sequence[int], row_elems: int,
                 device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem

        if world > _FLAG_SLOTS:
            raise RuntimeError("peer halo exchange: more ranks than flag slots")
        self.rank, self.world, self.device = rank, world, device
        t = torch.tensor(list(send_splits) + list(recv_splits), dtype=torch.int64, device=device)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t, group=group)
        S = [a[:world].tolist() for a in allt]  # S[r][d]: rows r sends to d in the forward pass
        R = [a[world:].tolist() for a in allt]  # R[r][s]: rows r receives from s
        for r in range(world):
            for d in range(world):
                if S[r][d] != R[d][r]:
                    raise RuntimeError("peer halo exchange: inconsistent split tables")
                if (S[r][d] > 0) != (S[d][r] > 0):
                    raise RuntimeError("peer halo exchange: asymmetric pattern (double buffering needs flags both ways)")
        self.n_send, self.n_recv = int(sum(S[rank])), int(sum(R[rank]))
        self.row_b = row_b = row_elems * 2  # bf16 rows
        cap_rows = max(max(sum(S[r]), sum(R[r])) for r in range(world))
        region = ((max(cap_rows, 1) * row_b + 255) // 256) * 256
        self.region = region
        flags_off = 4 * region
        total = flags_off + 2 * _FLAG_SLOTS * 4
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        ptrs: List[int] = [int(p) for p in self.hdl.buffer_ptrs]
        base = int(self.buf.data_ptr())
        if ptrs[rank] != base:
            raise RuntimeError("peer halo exchange: unexpected local mapping of the symmetric buffer")
        pre = lambda xs, k: int(sum(xs[:k]))
        # forward: my rows for d land behind the rows d receives from the ranks before me
        self.f_seg = _i64([pre(S[rank], d) for d in range(world + 1)])
        self.f_dst = [_i64([ptrs[d] + par * region + pre(R[d], rank) * row_b for d in range(world)]) for par in (0, 1)]
        self.f_flag = _i64([ptrs[d] + flags_off + rank * 4 if S[rank][d] > 0 else 0 for d in range(world)])
        self.f_mask = sum(1 << s for s in range(world) if R[rank][s] > 0)
        # backward: the gradient rows of what I received from s go back to s, behind what s sent to the ranks before me
        self.b_seg = _i64([pre(R[rank], s) for s in range(world + 1)])
        self.b_dst = [_i64([ptrs[s] + (2 + par) * region + pre(S[s], rank) * row_b for s in range(world)]) for par in (0, 1)]
        self.b_flag = _i64([ptrs[s] + flags_off + (_FLAG_SLOTS + rank) * 4 if R[rank][s] > 0 else 0 for s in range(world)])
        self.b_mask = sum(1 << d for d in range(world) if S[rank][d] > 0)
        self.flags_f = base + flags_off
        self.flags_b = base + flags_off + _FLAG_SLOTS * 4
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.e_f = 0
        self.e_b = 0
        bf = torch.bfloat16
        self._recv_f = [self.buf[par * region: par * region + self.n_recv * row_b].view(bf).view(self.n_recv, row_elems)
                        for par in (0, 1)]
        self._recv_b = [self.buf[(2 + par) * region: (2 + par) * region + self.n_send * row_b].view(bf).view(self.n_send, row_elems)
                        for par in (0, 1)]

    # ---- forward: table[send_idx, :row_elems] -> peers; returns (epoch, view of this rank's received rows)
    def push_fwd(self, table: torch.Tensor, send_idx: Optional[torch.Tensor]):
        from .. import ops

        self.e_f += 1
        par = self.e_f & 1
        ops.call("mgn_halo_push", ops._p(table), table.stride(0) * 2, 0, self.row_b, ops._p(send_idx), self.n_send, self.world,
                 self.f_seg, self.f_dst[par], self.row_b, self.f_flag, self.e_f, ops._p(self.counter), ops._stream())
        return self.e_f, self._recv_f[par]

    def wait_fwd(self, epoch: int) -> None:
        from .. import ops

        ops.call("mgn_halo_wait", self.flags_f, self.f_mask, epoch, ops._p(ops.tc_status(self.device)), ops._stream())

    # ---- backward: g_halo [n_recv, row_elems] (owner-ordered) -> owners; returns (epoch, view of the gradients of my sent rows)
    def push_bwd(self, g_halo: torch.Tensor):
        from .. import ops

        self.e_b += 1
        par = self.e_b & 1
        ops.call("mgn_halo_push", ops._p(g_halo), g_halo.stride(0) * 2 if g_halo.numel() else self.row_b, 0, self.row_b, None,
                 self.n_recv, self.world, self.b_seg, self.b_dst[par], self.row_b, self.b_flag, self.e_b, ops._p(self.counter),
                 ops._stream())
        return self.e_b, self._recv_b[par]

    def wait_bwd(self, epoch: int) -> None:
        from .. import ops

        ops.call("mgn_halo_wait", self.flags_b, self.b_mask, epoch, ops._p(ops.tc_status(self.device)), ops._stream())


class PeerWork:
    """`work.wait()` of the NCCL protocol: here it enqueues the flag wait on the current stream."""

    def __init__(self, fn, epoch: int):
        self._fn, self._epoch = fn, epoch

    def wait(self) -> None:
        self._fn(self._epoch)


def try_create(group, rank: int, world: int, send_splits, recv_splits, row_elems: int, device) -> Optional[PeerHalo]:
    """A PeerHalo when asked for (MGN_HALO_P2P=1) and possible; None otherwise (the caller keeps the NCCL protocol).
    The decision is collective: every rank of the group must succeed, or none uses it."""
    if not requested():
        return None
    ok, ph, why = 1, None, ""
    try:
        if dist.get_backend(group) != "nccl":
            raise RuntimeError("backend is not nccl")
        ph = PeerHalo(group, rank, world, send_splits, recv_splits, row_elems, device)
    except Exception as ex:  # noqa: BLE001  (any failure means: use NCCL)
        ok, why = 0, f"{type(ex).__name__}: {ex}"
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        if why and rank == 0:
            import warnings

            warnings.warn(f"modulus_b200: MGN_HALO_P2P=1 but the peer-memory halo exchange is unavailable ({why}); using NCCL")
        return None
    return ph
