"""Autograd wrappers of the variable-size collectives
(reference: physicsnemo/distributed/autograd.py:33-430; same public functions)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .utils import (
    all_gather_v_bwd_wrapper,
    all_gather_v_wrapper,
    gather_v_wrapper,
    indexed_all_to_all_v_wrapper,
    indexed_all_to_all_v_wrapper_bwd,
    scatter_v_wrapper,
)


class AllGatherVAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, sizes, dim=0, use_fp32=True, group=None):
        out = all_gather_v_wrapper(tensor, sizes, dim=dim, group=group)
        ctx.sizes, ctx.group, ctx.dim, ctx.use_fp32 = sizes, group, dim, use_fp32
        return out

    @staticmethod
    def backward(ctx, grad_output):
        g = all_gather_v_bwd_wrapper(grad_output, ctx.sizes, dim=ctx.dim, use_fp32=ctx.use_fp32, group=ctx.group)
        return (g if ctx.needs_input_grad[0] else None), None, None, None, None


class GatherVAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, sizes, dim=0, dst=0, group=None):
        out = gather_v_wrapper(tensor, sizes, dim=dim, dst=dst, group=group)
        ctx.sizes, ctx.dim, ctx.dst, ctx.group = sizes, dim, dst, group
        return out

    @staticmethod
    def backward(ctx, grad_output):
        g = scatter_v_wrapper(grad_output, ctx.sizes, dim=ctx.dim, src=ctx.dst, group=ctx.group)
        return (g if ctx.needs_input_grad[0] else None), None, None, None, None


class ScatterVAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, sizes, dim=0, src=0, group=None):
        out = scatter_v_wrapper(tensor, sizes, dim=dim, src=src, group=group)
        ctx.sizes, ctx.dim, ctx.src, ctx.group = sizes, dim, src, group
        return out

    @staticmethod
    def backward(ctx, grad_output):
        g = gather_v_wrapper(grad_output, ctx.sizes, dim=ctx.dim, dst=ctx.src, group=ctx.group)
        return (g if ctx.needs_input_grad[0] else None), None, None, None, None


class IndexedAllToAllVAutograd(torch.autograd.Function):
    """The halo exchange with its transposed backward (reference: autograd.py:184-247)."""

    @staticmethod
    def forward(ctx, tensor, indices, sizes, use_fp32=True, dim=0, group=None):
        out = indexed_all_to_all_v_wrapper(tensor, indices, sizes, dim=dim, group=group)
        ctx.sizes, ctx.use_fp32, ctx.group = sizes, use_fp32, group
        ctx.tensor_size_along_dim = tensor.size(dim)
        ctx.indices, ctx.dim = indices, dim
        return out

    @staticmethod
    def backward(ctx, grad_output):
        g = indexed_all_to_all_v_wrapper_bwd(
            grad_output, ctx.indices, ctx.sizes, tensor_size_along_dim=ctx.tensor_size_along_dim,
            use_fp32=ctx.use_fp32, dim=ctx.dim, group=ctx.group)
        return (g if ctx.needs_input_grad[0] else None), None, None, None, None, None


def all_gather_v(tensor: torch.Tensor, sizes: Optional[List[int]] = None, dim: int = 0, use_fp32: bool = True,
                 group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    return AllGatherVAutograd.apply(tensor, sizes, dim, use_fp32, group)


def gather_v(tensor: torch.Tensor, sizes: List[int], dim: int = 0, dst: int = 0,
             group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    return GatherVAutograd.apply(tensor, sizes, dim, dst, group)


def scatter_v(tensor: torch.Tensor, sizes: List[int], dim: int = 0, src: int = 0,
              group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    return ScatterVAutograd.apply(tensor, sizes, dim, src, group)


def indexed_all_to_all_v(tensor: torch.Tensor, indices: List[torch.Tensor], sizes: List[List[int]],
                         use_fp32: bool = True, dim: int = 0,
                         group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    return IndexedAllToAllVAutograd.apply(tensor, indices, sizes, use_fp32, dim, group)
