"""The operator seam of the GNN layers: concat_efeat, sum_efeat, aggregate_and_concat.

Same names, arguments and error behaviour as the reference
(physicsnemo/models/gnn_layers/utils.py:151-229, :260-334, :381-427); the DGL /
cugraph-ops backends are replaced by one backend: the CUDA kernels of libmgn_b200.so
driven through a per-graph `GraphPlan`.
"""
from __future__ import annotations

from typing import Any, Callable, Tuple, Union

import torch
from torch import Tensor
from torch.utils.checkpoint import checkpoint

from ...ops import AggConcatFn, ConcatEfeatFn, GraphPlan, SumEfeatFn
from .graph import CuGraphCSC


def checkpoint_identity(layer: Callable, *args: Any, **kwargs: Any) -> Any:
    """Identity stand-in for `torch.utils.checkpoint` (reference: utils.py:43-64)."""
    return layer(*args)


def set_checkpoint_fn(do_checkpointing: bool) -> Callable:
    """Reference: utils.py:67-91."""
    return checkpoint if do_checkpointing else checkpoint_identity


def graph_plan(graph, device=None) -> GraphPlan:
    """GraphPlan of a CuGraphCSC, a GraphPlan, or a DGL-like homogeneous/bipartite graph object
    (anything exposing ``edges()``, ``num_src_nodes()``, ``num_dst_nodes()``; edge rows are in
    edge-id order, as with DGLGraph in the reference)."""
    if isinstance(graph, GraphPlan):
        return graph
    if isinstance(graph, CuGraphCSC):
        if device is not None and graph.offsets.device != torch.device(device):
            graph.to(device)
        return graph.b200_plan()
    if hasattr(graph, "edges"):
        plan = getattr(graph, "_b200_plan", None)
        if plan is None or (device is not None and plan.device != torch.device(device)):
            src, dst = graph.edges()
            dev = device if device is not None else src.device
            plan = GraphPlan.from_coo(src.to(dev), dst.to(dev), graph.num_src_nodes(), graph.num_dst_nodes())
            try:
                graph._b200_plan = plan
            except Exception:  # objects with __slots__: just do not cache
                pass
        return plan
    raise TypeError(f"unsupported graph type {type(graph)!r}; expected CuGraphCSC or a DGL-like graph")


def _split(nfeat: Union[Tensor, Tuple[Tensor, Tensor]], graph) -> Tuple[Tensor, Tensor]:
    if isinstance(nfeat, Tensor):
        src_feat, dst_feat = nfeat, nfeat
    else:
        src_feat, dst_feat = nfeat
    if isinstance(graph, CuGraphCSC) and graph.is_distributed:
        # halo exchange of the source rows owned by other ranks (utils.py:177-178, 212-213)
        src_feat = graph.get_src_node_features_in_local_graph(src_feat)
    return src_feat, dst_feat


def concat_efeat(
    efeat: Tensor,
    nfeat: Union[Tensor, Tuple[Tensor, Tensor]],
    graph,
) -> Tensor:
    """cat(efeat, src_feat[src], dst_feat[dst]) per edge (reference: utils.py:151-229)."""
    src_feat, dst_feat = _split(nfeat, graph)
    return ConcatEfeatFn.apply(efeat, src_feat, dst_feat, graph_plan(graph, efeat.device))


def sum_efeat(
    efeat: Tensor,
    nfeat: Union[Tensor, Tuple[Tensor, Tensor]],
    graph,
) -> Tensor:
    """efeat + src_feat[src] + dst_feat[dst] per edge (reference: utils.py:260-334)."""
    src_feat, dst_feat = _split(nfeat, graph)
    return SumEfeatFn.apply(efeat, src_feat, dst_feat, graph_plan(graph, efeat.device))


def aggregate_and_concat(
    efeat: Tensor,
    nfeat: Tensor,
    graph,
    aggregation: str,
) -> Tensor:
    """cat(sum|mean of efeat over incoming edges, nfeat) per destination node
    (reference: utils.py:381-427).  No communication even when distributed (:412-416)."""
    if aggregation not in ("sum", "mean"):
        raise RuntimeError("Not a valid aggregation!")
    return AggConcatFn.apply(efeat, nfeat, graph_plan(graph, efeat.device), aggregation == "mean")
