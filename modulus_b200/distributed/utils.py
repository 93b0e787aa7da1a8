"""Variable-size collectives of the partitioned-graph path.

Function names, argument meaning and error behaviour follow
physicsnemo/distributed/utils.py:216-765.  What differs is the realisation of the hot one:

  indexed_all_to_all_v (utils.py:541-603) -- the per-layer HALO EXCHANGE -- is executed as
    one pack kernel (a single indexed row gather for ALL peers, mgn_gather_rows) into a
    contiguous send buffer  ->  ONE all_to_all_single (NCCL over NVLink)  ->  the receive
    buffer IS the concatenated result (no per-peer gathers, no torch.cat);
  its backward (utils.py:606-707) is the reverse all_to_all_single followed by a
    deterministic segmented sum over rows grouped by destination row (mgn_segment_sum, fp32
    accumulation) instead of `index_add_` atomics.

gather_v / scatter_v / all_gather_v move whole tensors once per step at most (model
boundary) and stay thin wrappers over torch.distributed.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops
from .manager import DistributedManager


# ----------------------------------------------------------------------------------------
# transport.  NCCL is the product transport (one all_to_all_single over NVLink).  Any other backend
# (gloo: the CPU test harness, and the single-device multi-process emulation of tests/test_gpu_dist.py, where
# several ranks share one GPU, which NCCL refuses) gets a point-to-point realisation; device tensors are staged
# through host memory there, because gloo moves host buffers only.  Nothing else differs between the two.
# ----------------------------------------------------------------------------------------
def _backend_has_alltoall(group) -> bool:
    try:
        return dist.get_backend(group) == "nccl"
    except Exception:
        return False


def _host_staged(group, *tensors) -> bool:
    return (not _backend_has_alltoall(group)) and any(t.is_cuda for t in tensors if t is not None)


def all_to_all_list(x_recv: List[torch.Tensor], x_send: List[torch.Tensor], group=None) -> None:
    """dist.all_to_all with a point-to-point realisation for backends without alltoall (gloo)."""
    if _backend_has_alltoall(group):
        dist.all_to_all(x_recv, x_send, group=group)
        return
    rank = dist.get_rank(group=group)
    size = dist.get_world_size(group=group)
    ranks = dist.get_process_group_ranks(group if group is not None else dist.group.WORLD)
    staged = _host_staged(group, *x_recv, *x_send)
    recv_bufs = [torch.empty(t.shape, dtype=t.dtype) for t in x_recv] if staged else x_recv
    x_recv[rank].copy_(x_send[rank])
    ops_ = []
    for r in range(size):
        if r == rank:
            continue
        if x_send[r].numel() > 0:
            ops_.append(dist.P2POp(dist.isend, x_send[r].contiguous().cpu() if staged else x_send[r].contiguous(),
                                   ranks[r], group=group))
        if x_recv[r].numel() > 0:
            ops_.append(dist.P2POp(dist.irecv, recv_bufs[r], ranks[r], group=group))
    if ops_:
        for req in dist.batch_isend_irecv(ops_):
            req.wait()
    if staged:
        for r in range(size):
            if r != rank and x_recv[r].numel() > 0:
                x_recv[r].copy_(recv_bufs[r])


def all_to_all_rows(send: torch.Tensor, send_splits: Sequence[int], recv_splits: Sequence[int],
                    group=None) -> torch.Tensor:
    """Row-wise all_to_all_single: `send` is [sum(send_splits), ...], returns [sum(recv_splits), ...]."""
    recv = send.new_empty((int(sum(recv_splits)),) + tuple(send.shape[1:]))
    if _backend_has_alltoall(group):
        dist.all_to_all_single(recv, send, list(recv_splits), list(send_splits), group=group)
    else:
        all_to_all_list(list(torch.split(recv, list(recv_splits), dim=0)),
                        list(torch.split(send, list(send_splits), dim=0)), group=group)
    return recv


def all_to_all_rows_async(send: torch.Tensor, send_splits: Sequence[int], recv_splits: Sequence[int], group=None):
    """(work, recv): the exchange is in flight on NCCL's stream until work.wait(); `work` is None when the
    transport completed it synchronously."""
    if _backend_has_alltoall(group):
        recv = send.new_empty((int(sum(recv_splits)),) + tuple(send.shape[1:]))
        work = dist.all_to_all_single(recv, send, list(recv_splits), list(send_splits), group=group, async_op=True)
        return work, recv
    return None, all_to_all_rows(send, send_splits, recv_splits, group=group)


def _all_reduce_sum(t: torch.Tensor, group=None) -> None:
    if _host_staged(group, t):
        h = t.cpu()
        dist.all_reduce(h, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, group=group)


# ----------------------------------------------------------------------------------------
# halo exchange plan
# ----------------------------------------------------------------------------------------
class HaloExchangePlan:
    """Precomputed index structures of one indexed all-to-all-v.

      send_idx    [S] int32  = cat(indices[r] for r)  rows of the local table to pack, peer-major
      send_splits [P]        = sizes[rank][r]
      recv_splits [P]        = sizes[r][rank]
      acc_offsets / acc_ids  rows of the received-gradient buffer grouped by local row
                             (stable) -> deterministic backward accumulation
    """

    def __init__(self, indices: Sequence[torch.Tensor], sizes: List[List[int]], rank: int):
        P = len(indices)
        self.send_splits = [int(sizes[rank][r]) for r in range(P)]
        self.recv_splits = [int(sizes[r][rank]) for r in range(P)]
        for r in range(P):
            if int(indices[r].numel()) != self.send_splits[r]:
                raise ValueError("sizes and indices of the indexed all-to-all do not match")
        self.send_idx64 = torch.cat([i.reshape(-1).to(torch.int64) for i in indices], dim=0)
        self.send_idx = self.send_idx64.to(torch.int32)
        self.device = self.send_idx.device
        self._acc = {}

    def acc_structs(self, n_rows: int):
        if n_rows not in self._acc:
            if self.device.type == "cuda":
                self._acc[n_rows] = ops._group_by_key(self.send_idx, n_rows)
            else:
                self._acc[n_rows] = None
        return self._acc[n_rows]


_PLAN_CACHE = {}


def _halo_plan(indices: Sequence[torch.Tensor], sizes: List[List[int]], rank: int) -> HaloExchangePlan:
    key = (tuple((int(i.data_ptr()), int(i.numel())) for i in indices), rank)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) > 64:
            _PLAN_CACHE.clear()
        plan = HaloExchangePlan(indices, sizes, rank)
        plan._keepalive = list(indices)  # the cache key holds data_ptrs: keep the tensors alive
        _PLAN_CACHE[key] = plan
    return plan


def _check_a2a_args(tensor, indices, sizes, dim, group):
    comm_size = dist.get_world_size(group=group)
    rank = dist.get_rank(group=group)
    if len(sizes) != comm_size:
        raise ValueError()
    if dim >= tensor.dim():
        raise ValueError()
    if len(sizes[rank]) != comm_size:
        raise ValueError()
    if len(indices) != comm_size:
        raise ValueError()
    if dim != 0:
        raise NotImplementedError("modulus_b200 exchanges rows (dim=0) only")
    return comm_size, rank


def indexed_all_to_all_v_wrapper(
    tensor: torch.Tensor,
    indices: List[torch.Tensor],
    sizes: List[List[int]],
    dim: int = 0,
    group: Optional[dist.ProcessGroup] = None,
) -> torch.Tensor:
    """Forward halo exchange (reference: utils.py:541-603): rank p receives
    cat_r( tensor_r[indices_r[p]] ) in rank order."""
    _, rank = _check_a2a_args(tensor, indices, sizes, dim, group)
    plan = _halo_plan(indices, sizes, rank)
    tensor = tensor.contiguous()
    if tensor.is_cuda:
        flat = tensor.reshape(tensor.shape[0], -1)
        packed = ops.gather_rows(flat, 0, flat.shape[1], plan.send_idx, plan.send_idx.numel())
        packed = packed.reshape((packed.shape[0],) + tuple(tensor.shape[1:]))
    else:  # host tensors only occur with gloo (CPU test harness of the exchange protocol)
        packed = tensor[plan.send_idx64]
    return all_to_all_rows(packed, plan.send_splits, plan.recv_splits, group=group)


def indexed_all_to_all_v_wrapper_bwd(
    tensor: torch.Tensor,
    indices: List[torch.Tensor],
    sizes: List[List[int]],
    tensor_size_along_dim: int,
    use_fp32: bool = True,
    dim: int = 0,
    group: Optional[dist.ProcessGroup] = None,
) -> torch.Tensor:
    """Backward halo exchange (reference: utils.py:606-707): gradients travel back and are summed
    into the rows they were gathered from; fp32 accumulation for sub-fp32 dtypes."""
    _, rank = _check_a2a_args(tensor, indices, sizes, dim, group)
    plan = _halo_plan(indices, sizes, rank)
    tensor = tensor.contiguous()
    recv = all_to_all_rows(tensor, plan.recv_splits, plan.send_splits, group=group)
    if recv.is_cuda:
        offsets, ids = plan.acc_structs(tensor_size_along_dim)
        flat = recv.reshape(recv.shape[0], -1)
        # the kernel accumulates in fp32 registers and rounds once (>= the reference's use_fp32 path)
        out = ops.segment_sum(flat, 0, flat.shape[1], offsets, ids, tensor_size_along_dim)
        return out.reshape((tensor_size_along_dim,) + tuple(tensor.shape[1:]))
    acc_dtype = torch.float32 if (use_fp32 and recv.dtype.itemsize < 4 and recv.dtype.is_floating_point) else recv.dtype
    out = torch.zeros((tensor_size_along_dim,) + tuple(tensor.shape[1:]), dtype=acc_dtype)
    out.index_add_(0, plan.send_idx64, recv.to(acc_dtype))
    return out.to(tensor.dtype)


# ----------------------------------------------------------------------------------------
# whole-tensor primitives (model boundary only)
# ----------------------------------------------------------------------------------------
def all_gather_v_wrapper(tensor, sizes: Optional[List[int]] = None, dim: int = 0, group=None) -> torch.Tensor:
    """Reference: utils.py:216-297."""
    comm_size = dist.get_world_size(group=group)
    if (sizes is not None) and (len(sizes) != comm_size):
        raise ValueError(f"Mismatch in sizes {len(sizes)} and comm_size {comm_size}")
    if dim >= tensor.dim():
        raise ValueError()
    if comm_size == 1:
        return tensor
    shape = list(tensor.shape)
    tensor_list = []
    for r in range(comm_size):
        if sizes is not None:
            shape[dim] = sizes[r]
        tensor_list.append(torch.empty(shape, dtype=tensor.dtype, device=tensor.device))
    if _backend_has_alltoall(group):
        dist.all_gather(tensor_list, tensor.contiguous(), group=group)
    else:  # gloo's all_gather wants equal shapes: every rank sends its block to every peer instead
        all_to_all_list(tensor_list, [tensor.contiguous()] * comm_size, group=group)
    return torch.cat(tensor_list, dim=dim).contiguous()


def all_gather_v_bwd_wrapper(tensor, sizes: List[int], dim: int = 0, use_fp32: bool = True, group=None):
    """All-reduce-v = backward of all_gather_v (reference: utils.py:299-372)."""
    comm_size = dist.get_world_size(group=group)
    rank = dist.get_rank(group=group)
    if len(sizes) != comm_size:
        raise ValueError()
    if dim >= tensor.dim():
        raise ValueError()
    shape = list(tensor.shape)
    shape[dim] = sizes[rank]
    tmp = [torch.empty(shape, dtype=tensor.dtype, device=tensor.device) for _ in range(comm_size)]
    scatter_list = [t.contiguous() for t in torch.split(tensor, sizes, dim=dim)]
    all_to_all_list(tmp, scatter_list, group=group)
    stacked = torch.stack(tmp, dim=tensor.dim())
    if use_fp32 and (stacked.dtype.itemsize < 4) and stacked.dtype.is_floating_point:
        return stacked.sum(dim=tensor.dim(), dtype=torch.float32).to(dtype=tensor.dtype)
    return stacked.sum(dim=tensor.dim())


def gather_v_wrapper(tensor, sizes: List[int], dim: int = 0, dst: int = 0, group=None) -> torch.Tensor:
    """Reference: utils.py:374-451."""
    comm_size = dist.get_world_size(group=group)
    rank = dist.get_rank(group=group)
    if len(sizes) != comm_size:
        raise ValueError()
    if dim >= tensor.dim():
        raise ValueError()
    if not (0 <= dst < comm_size):
        raise ValueError()
    if tensor.size(dim) != sizes[rank]:
        raise ValueError()
    if comm_size == 1:
        return tensor
    shape = list(tensor.shape)
    x_recv, x_send = [None] * comm_size, [None] * comm_size
    for r in range(comm_size):
        shape[dim] = sizes[r] if rank == dst else 0
        x_recv[r] = torch.empty(shape, dtype=tensor.dtype, device=tensor.device)
        if r == dst:
            x_send[r] = tensor.contiguous()
        else:
            shape[dim] = 0
            x_send[r] = torch.empty(shape, dtype=tensor.dtype, device=tensor.device)
    all_to_all_list(x_recv, x_send, group=group)
    if rank != dst:
        for r in range(comm_size):
            shape[dim] = sizes[r]
            x_recv[r] = torch.empty(shape, dtype=tensor.dtype, device=tensor.device)
    return torch.cat(x_recv, dim=dim)


def scatter_v_wrapper(tensor, sizes: List[int], dim: int = 0, src: int = 0, group=None) -> torch.Tensor:
    """Reference: utils.py:453-539."""
    comm_size = dist.get_world_size(group=group)
    rank = dist.get_rank(group=group)
    if len(sizes) != comm_size:
        raise ValueError()
    if dist.get_rank(group=group) == 0 and dim >= tensor.dim():
        raise ValueError()
    if not (0 <= src < comm_size):
        raise ValueError()
    shape = list(tensor.shape)
    x_send, x_recv = [None] * comm_size, [None] * comm_size
    if rank == src:
        x_send = [t.contiguous() for t in torch.split(tensor, sizes, dim=dim)]
    else:
        for r in range(comm_size):
            shape[dim] = 0
            x_send[r] = torch.empty(shape, device=tensor.device, dtype=tensor.dtype)
    for r in range(comm_size):
        shape[dim] = sizes[rank] if r == src else 0
        x_recv[r] = torch.empty(shape, device=tensor.device, dtype=tensor.dtype)
    all_to_all_list(x_recv, x_send, group=group)
    return x_recv[src]


# ----------------------------------------------------------------------------------------
# shared-weight gradient reduction
# ----------------------------------------------------------------------------------------
def _reduce(input_, use_fp32=True, group=None):
    """All-reduce across the model-parallel group (reference: utils.py:176-193)."""
    if dist.get_world_size(group=group) == 1:
        return input_
    if use_fp32 and (input_.dtype.itemsize < 4) and input_.dtype.is_floating_point:
        dtype = input_.dtype
        inputf_ = input_.float()
        _all_reduce_sum(inputf_, group=group)
        input_ = inputf_.to(dtype)
    else:
        _all_reduce_sum(input_, group=group)
    return input_


def mark_module_as_shared(module: nn.Module, process_group: Optional[str], recurse: bool = True,
                          use_fp32_reduction: bool = True) -> nn.Module:
    """Attach gradient hooks that sum parameter gradients over the partition group
    (reference: utils.py:710-765; one all-reduce per parameter, post-accumulate hook)."""
    group = DistributedManager().group(process_group)
    handle_key = "_shared_weight_dist_hook"

    # Same result as the reference's one-all-reduce-per-parameter hooks, one NCCL call per backward pass: every
    # hook files its parameter, the first one queues an end-of-backward callback that all-reduces ONE flat fp32
    # bucket (263 tensors / 9.3 MB for the default MeshGraphNet) and scatters it back.
    pending: List[torch.Tensor] = []

    def flush() -> None:
        params = list(pending)
        pending.clear()
        if not params or dist.get_world_size(group=group) == 1:
            return
        grads = [p.grad for p in params]
        acc = torch.float32 if use_fp32_reduction else None
        flat = torch.cat([g.reshape(-1).to(acc or g.dtype) for g in grads])
        _all_reduce_sum(flat, group=group)
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n

    def hook_post_accum(param: torch.Tensor) -> None:
        if not pending:
            try:
                torch.autograd.Variable._execution_engine.queue_callback(flush)
            except Exception:  # not inside a backward pass driven by the engine: reduce right away
                param.grad = _reduce(param.grad, group=group, use_fp32=use_fp32_reduction)
                return
        pending.append(param)

    for name, param in module.named_parameters(recurse=recurse):
        if hasattr(param, handle_key):
            raise RuntimeError(
                f"Parameter {name} already marked as having shared weights, can't mark it again!")
        handle = param.register_post_accumulate_grad_hook(hook_post_accum)
        setattr(param, handle_key, handle)
    return module


def unmark_module_as_shared(module: nn.Module, recurse: bool = True) -> nn.Module:
    """Reference: utils.py:768-796."""
    handle_key = "_shared_weight_dist_hook"
    for name, param in module.named_parameters(recurse=recurse):
        if not hasattr(param, handle_key):
            raise RuntimeError(f"Parameter {name} NOT marked as having shared weights, can't unmark it!")
        getattr(param, handle_key).remove()
        delattr(param, handle_key)
    return module


def reduce_shared_gradients(module: nn.Module, process_group: Optional[str] = None) -> None:
    """B200 realisation of the same reduction: ONE flat fp32 all-reduce over all parameter
    gradients of `module` (263 tensors for the default MeshGraphNet) instead of one NCCL call per
    tensor.  Call after backward() when the module was NOT marked with mark_module_as_shared."""
    group = DistributedManager().group(process_group)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group=group) == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    _all_reduce_sum(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


@torch.no_grad()
def reduce_loss(loss: float, dst_rank: int = 0, mean: bool = True):
    """Reference: utils.py:113-150."""
    if not DistributedManager.is_initialized():
        raise Exception("Distributed manager should be initialized when using reduce_loss")
    distmng = DistributedManager()
    loss = torch.Tensor([loss]).to(distmng.device)
    if distmng.world_size == 1:
        return float(loss)
    dist.reduce(loss, dst_rank, dist.ReduceOp.SUM, group=None)
    if mean:
        loss = loss / distmng.world_size
    return float(loss.cpu()) if distmng.rank == dst_rank else None
