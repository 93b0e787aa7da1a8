"""MeshGraphMLP / MeshGraphEdgeMLPConcat / MeshGraphEdgeMLPSum on the B200 kernels.

Constructor signatures, parameter names / shapes / init order (hence `state_dict` layout and
the RNG stream under a fixed seed) are those of the reference
(physicsnemo/models/gnn_layers/mesh_graph_mlp.py:103-458).  The `nn.Sequential` named
`model` only HOLDS the parameters; `forward` hands them, in place and in fp32, to the CUDA
kernels (ops.mlp_forward or the fused tcgen05 path) -- nothing here calls nn.Linear.forward.
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from ... import fused, ops
from ..layers.activations import activation_name
from .graph import CuGraphCSC
from .utils import _split, graph_plan, sum_efeat


def compute_dtype(x: Tensor) -> torch.dtype:
    """fp32 tensors run the fp32 kernels; under `torch.autocast(dtype=bfloat16)` (the reference's
    AMP switch, examples/cfd/vortex_shedding_mgn/train.py:153) or for bf16 inputs the bf16 kernels
    (bf16 storage, fp32 accumulate, fp32 LayerNorm statistics)."""
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        if dt == torch.float16:
            # The reference recipe's default AMP is float16 + GradScaler (train.py:153-166).  The kernels' reduced
            # precision storage type is bfloat16 (fp32 range: no overflow, loss scaling is harmless); a float16 autocast
            # region therefore computes in bf16 and MeshGraphNet.forward hands back a float16 tensor, as the caller expects.
            return torch.bfloat16
        if dt != torch.bfloat16:
            raise NotImplementedError(f"modulus_b200 supports bfloat16 / float16 autocast only (got {dt})")
        return dt
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"modulus_b200 supports float32 / bfloat16 features, got {x.dtype}")
    return x.dtype


def autocast_result_dtype(x: Tensor) -> torch.dtype:
    """dtype a model output gets: float16 inside a float16 autocast region (see compute_dtype), else unchanged"""
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.float16:
        return torch.float16
    return x.dtype


def _check_norm(norm_type):
    if norm_type is not None and norm_type not in ["LayerNorm", "TELayerNorm"]:
        raise ValueError(
            f"Invalid norm type {norm_type}. Supported types are LayerNorm and TELayerNorm."
        )


class MeshGraphMLP(nn.Module):
    """Linear(in,hid), act, [Linear(hid,hid), act] x (hidden_layers-1), Linear(hid,out), [LayerNorm(out)]
    (reference: mesh_graph_mlp.py:103-203).  `hidden_layers=None` collapses to the identity.
    "TELayerNorm" is accepted and served by the same fused LayerNorm kernels."""

    def __init__(
        self,
        input_dim: int,
        output_dim: int = 512,
        hidden_dim: int = 512,
        hidden_layers: Union[int, None] = 1,
        activation_fn: nn.Module = nn.SiLU(),
        norm_type: str = "LayerNorm",
        recompute_activation: bool = False,
    ):
        super().__init__()
        self.input_dim, self.output_dim, self.hidden_dim = input_dim, output_dim, hidden_dim
        self.activation_fn = activation_fn
        if hidden_layers is not None:
            layers = [nn.Linear(input_dim, hidden_dim), activation_fn]
            self.hidden_layers = hidden_layers
            for _ in range(hidden_layers - 1):
                layers += [nn.Linear(hidden_dim, hidden_dim), activation_fn]
            layers.append(nn.Linear(hidden_dim, output_dim))

            self.norm_type = norm_type
            _check_norm(norm_type)
            if norm_type is not None:
                layers.append(nn.LayerNorm(output_dim))
            self.model = nn.Sequential(*layers)
        else:
            self.hidden_layers = None
            self.norm_type = None
            self.model = nn.Identity()

        if recompute_activation:
            if not isinstance(activation_fn, nn.SiLU):
                raise ValueError(activation_fn)
            self.recompute_activation = True
        else:
            self.recompute_activation = False

    # -------------------------------------------------------------- parameter views
    def _linears(self) -> List[nn.Linear]:
        return [self.model[2 * i] for i in range(self.hidden_layers + 1)]

    def _norm(self) -> Optional[nn.LayerNorm]:
        return self.model[2 * self.hidden_layers + 1] if self.norm_type is not None else None

    def _flat_params(self) -> List[Optional[Tensor]]:
        ps: List[Optional[Tensor]] = []
        for lin in self._linears():
            ps += [lin.weight, lin.bias]
        nrm = self._norm()
        if nrm is not None:
            ps += [nrm.weight, nrm.bias]
        return ps

    # -------------------------------------------------------------- forward
    def mlp(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        """out = MLP(x) [+ residual]; residual add is fused into the LayerNorm epilogue."""
        if self.hidden_layers is None:
            return x if residual is None else x + residual
        dt = compute_dtype(x)
        if residual is None and fused.ENABLED and fused.mlp_eligible(self, x, dt):
            return fused.mlp_forward(self, x)  # tcgen05 path: encoders / decoder
        x = x.to(dt)
        if residual is not None:
            residual = residual.to(dt)
        nrm = self._norm()
        return ops.mlp_forward(
            x, self._flat_params(), self.hidden_layers + 1, activation_name(self.activation_fn),
            nrm is not None, residual=residual, eps=(nrm.eps if nrm is not None else 1e-5),
        )

    def default_forward(self, x: Tensor) -> Tensor:
        return self.mlp(x)

    def custom_silu_linear_forward(self, x: Tensor) -> Tensor:
        """The reference recomputes SiLU in backward here (mesh_graph_mlp.py:184-197); the kernels
        already keep only what their backward needs, so this is the same computation."""
        return self.mlp(x)

    def forward(self, x: Tensor) -> Tensor:
        return self.mlp(x)


class MeshGraphEdgeMLPConcat(MeshGraphMLP):
    """concat_efeat followed by the MLP (reference: mesh_graph_mlp.py:206-275).  W1 columns are
    ordered [efeat | src | dst] (utils.py:108)."""

    def __init__(
        self,
        efeat_dim: int = 512,
        src_dim: int = 512,
        dst_dim: int = 512,
        output_dim: int = 512,
        hidden_dim: int = 512,
        hidden_layers: int = 2,
        activation_fn: nn.Module = nn.SiLU(),
        norm_type: str = "LayerNorm",
        bias: bool = True,
        recompute_activation: bool = False,
    ):
        cat_dim = efeat_dim + src_dim + dst_dim
        super().__init__(cat_dim, output_dim, hidden_dim, hidden_layers, activation_fn, norm_type,
                         recompute_activation)
        self.efeat_dim, self.src_dim, self.dst_dim = efeat_dim, src_dim, dst_dim

    def edge_mlp(self, efeat: Tensor, nfeat, graph, residual: Optional[Tensor] = None) -> Tensor:
        dt = compute_dtype(efeat)
        efeat = efeat.to(dt)
        src_feat, dst_feat = _split(nfeat, graph)
        src_feat, dst_feat = src_feat.to(dt), dst_feat.to(dt)
        plan = graph_plan(graph, efeat.device)
        cat = ops.ConcatEfeatFn.apply(efeat, src_feat, dst_feat, plan)
        return self.mlp(cat, residual=residual)

    def forward(self, efeat: Tensor, nfeat: Union[Tensor, Tuple[Tensor]], graph) -> Tensor:
        return self.edge_mlp(efeat, nfeat, graph)


class MeshGraphEdgeMLPSum(nn.Module):
    """"Concat trick" edge MLP (reference: mesh_graph_mlp.py:278-458): the first Linear is split into
    three bias-free per-source matmuls applied before the gather, then summed per edge.  Parameter
    names (`lin_efeat`, `lin_src`, `lin_dst`, `bias`, `model.*`) and the init RNG stream match."""

    def __init__(
        self,
        efeat_dim: int,
        src_dim: int,
        dst_dim: int,
        output_dim: int = 512,
        hidden_dim: int = 512,
        hidden_layers: int = 1,
        activation_fn: nn.Module = nn.SiLU(),
        norm_type: str = "LayerNorm",
        bias: bool = True,
        recompute_activation: bool = False,
    ):
        super().__init__()
        self.efeat_dim, self.src_dim, self.dst_dim = efeat_dim, src_dim, dst_dim
        self.hidden_dim, self.output_dim = hidden_dim, output_dim
        self.activation_fn = activation_fn

        tmp_lin = nn.Linear(efeat_dim + src_dim + dst_dim, hidden_dim, bias=bias)
        w_efeat, w_src, w_dst = torch.split(tmp_lin.weight, [efeat_dim, src_dim, dst_dim], dim=1)
        self.lin_efeat = nn.Parameter(w_efeat)
        self.lin_src = nn.Parameter(w_src)
        self.lin_dst = nn.Parameter(w_dst)
        self.bias = tmp_lin.bias if bias else None

        layers = [activation_fn]
        self.hidden_layers = hidden_layers
        for _ in range(hidden_layers - 1):
            layers += [nn.Linear(hidden_dim, hidden_dim), activation_fn]
        layers.append(nn.Linear(hidden_dim, output_dim))

        self.norm_type = norm_type
        _check_norm(norm_type)
        if norm_type is not None:
            layers.append(nn.LayerNorm(output_dim))
        self.model = nn.Sequential(*layers)

        if recompute_activation:
            if not isinstance(activation_fn, nn.SiLU):
                raise ValueError(activation_fn)
            self.recompute_activation = True
        else:
            self.recompute_activation = False

    # ---- the fused tensor-core path reads this module through the same two accessors as MeshGraphMLP: the first Linear
    # re-assembled as ONE [hidden, efeat | src | dst] matrix (what it was split from, mesh_graph_mlp.py:335-350); autograd
    # routes the gradient of the concatenation back to lin_efeat / lin_src / lin_dst
    def _norm(self) -> Optional[nn.LayerNorm]:
        return self.model[2 * self.hidden_layers] if self.norm_type is not None else None

    def _flat_params(self) -> List[Optional[Tensor]]:
        ps: List[Optional[Tensor]] = [torch.cat([self.lin_efeat, self.lin_src, self.lin_dst], dim=1), self.bias]
        for i in range(self.hidden_layers):
            lin = self.model[2 * i + 1]
            ps += [lin.weight, lin.bias]
        nrm = self._norm()
        if nrm is not None:
            ps += [nrm.weight, nrm.bias]
        return ps

    def forward_truncated_sum(self, efeat: Tensor, nfeat, graph) -> Tensor:
        dt = compute_dtype(efeat)
        efeat = efeat.to(dt)
        if isinstance(nfeat, Tensor):
            src_feat, dst_feat = nfeat, nfeat
        else:
            src_feat, dst_feat = nfeat
        src_feat, dst_feat = src_feat.to(dt), dst_feat.to(dt)
        # node-level matmuls run BEFORE the halo exchange, exactly like the reference (:396-405)
        mlp_efeat = ops.mlp_forward(efeat, [self.lin_efeat, None], 1, "identity", False)
        mlp_src = ops.mlp_forward(src_feat, [self.lin_src, None], 1, "identity", False)
        mlp_dst = ops.mlp_forward(dst_feat, [self.lin_dst, self.bias], 1, "identity", False)
        return sum_efeat(mlp_efeat, (mlp_src, mlp_dst), graph)

    def edge_mlp(self, efeat: Tensor, nfeat, graph, residual: Optional[Tensor] = None) -> Tensor:
        act = activation_name(self.activation_fn)
        x = ops.activation(self.forward_truncated_sum(efeat, nfeat, graph), act)
        params: List[Optional[Tensor]] = []
        for i in range(self.hidden_layers):
            lin = self.model[2 * i + 1]
            params += [lin.weight, lin.bias]
        nrm = self.model[2 * self.hidden_layers] if self.norm_type is not None else None
        if nrm is not None:
            params += [nrm.weight, nrm.bias]
        if residual is not None:
            residual = residual.to(x.dtype)
        return ops.mlp_forward(x, params, self.hidden_layers, act, nrm is not None, residual=residual,
                               eps=(nrm.eps if nrm is not None else 1e-5))

    def default_forward(self, efeat, nfeat, graph) -> Tensor:
        return self.edge_mlp(efeat, nfeat, graph)

    def custom_silu_linear_forward(self, efeat, nfeat, graph) -> Tensor:
        return self.edge_mlp(efeat, nfeat, graph)

    def forward(self, efeat: Tensor, nfeat: Union[Tensor, Tuple[Tensor]], graph) -> Tensor:
        return self.edge_mlp(efeat, nfeat, graph)
