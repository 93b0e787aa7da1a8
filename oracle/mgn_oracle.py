"""CPU oracle for the MeshGraphNet message-passing hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product path (modulus_b200) never does
and fails loudly when its CUDA library is missing.

This is a plain-torch (CPU, fp32 or fp64) restatement of the reference algorithm, written
functionally over a state_dict that uses the reference's own parameter names, so the same
weights drive the reference, the oracle and the CUDA path.  Each function cites the
reference file:line it follows (paths relative to /root/reference/physicsnemo).

Parity pin: tests/test_oracle.py checks this file against
  * the reference's own golden vector test/models/data/meshgraphnet_output.pth
    (stored as tests/golden/kat1_meshgraphnet_output.pt), and
  * outputs + gradients produced by the unmodified reference imported in the build
    container (tests/golden/make_golden.py, fixtures tests/golden/ref_*.pt).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

_ACT = {
    "relu": F.relu,
    "silu": F.silu,
    "gelu": F.gelu,
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
    "leaky_relu": F.leaky_relu,
    "elu": F.elu,
    "selu": F.selu,
    "identity": lambda x: x,
}


# ----------------------------------------------------------------------------------------
# graph helpers
# ----------------------------------------------------------------------------------------
def csc_from_coo(src: Tensor, dst: Tensor, num_dst: int) -> Tuple[Tensor, Tensor, Tensor]:
    """COO -> CSC with a STABLE sort by destination (models/gnn_layers/graph.py:143-193;
    DGL's adj_tensors("csc") edge order is not pinned by the reference).
    Returns (offsets[int64, num_dst+1], indices = src ids in CSC order, edge_perm)."""
    perm = torch.argsort(dst.long(), stable=True)
    deg = torch.bincount(dst.long(), minlength=num_dst)
    offsets = torch.zeros(num_dst + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(deg, 0)
    return offsets, src.long()[perm], perm


def coo_from_csc(offsets: Tensor, indices: Tensor) -> Tuple[Tensor, Tensor]:
    """CSC -> (src, dst) per edge in CSC order (models/gnn_layers/graph.py:459-471)."""
    deg = offsets[1:] - offsets[:-1]
    dst = torch.repeat_interleave(torch.arange(offsets.numel() - 1, dtype=torch.int64), deg.long())
    return indices.long(), dst


def csr_from_csc(offsets: Tensor, indices: Tensor, num_src: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Transpose of the CSC structure: for every source node the list of CSC edge
    positions leaving it, in ascending edge position (stable).  Returns
    (csr_offsets[num_src+1], csr_eids[E] = CSC edge position, csr_dst[E])."""
    src, dst = coo_from_csc(offsets, indices)
    perm = torch.argsort(src, stable=True)
    deg = torch.bincount(src, minlength=num_src)
    csr_offsets = torch.zeros(num_src + 1, dtype=torch.int64)
    csr_offsets[1:] = torch.cumsum(deg, 0)
    return csr_offsets, perm, dst[perm]


# ----------------------------------------------------------------------------------------
# operator seam (models/gnn_layers/utils.py)
# ----------------------------------------------------------------------------------------
def concat_efeat(efeat: Tensor, src_feat: Tensor, dst_feat: Tensor, src: Tensor, dst: Tensor) -> Tensor:
    """cat(efeat, src_feat[src], dst_feat[dst]) -- utils.py:94-109 (concat order e, src, dst)."""
    return torch.cat((efeat, src_feat[src], dst_feat[dst]), dim=1)


def sum_efeat(efeat: Tensor, src_feat: Tensor, dst_feat: Tensor, src: Tensor, dst: Tensor) -> Tensor:
    """efeat + src_feat[src] + dst_feat[dst] -- utils.py:232-257."""
    return efeat + src_feat[src] + dst_feat[dst]


def aggregate_and_concat(efeat: Tensor, dst_feat: Tensor, dst: Tensor, aggregation: str = "sum") -> Tensor:
    """cat(segment-reduce of efeat by destination, dst_feat) -- utils.py:337-378.
    mean divides by the in-degree; zero in-degree nodes aggregate to 0."""
    n = dst_feat.shape[0]
    h = torch.zeros((n, efeat.shape[1]), dtype=efeat.dtype).index_add(0, dst, efeat)
    if aggregation == "mean":
        deg = torch.bincount(dst, minlength=n).clamp(min=1).to(efeat.dtype)
        h = h / deg[:, None]
    elif aggregation != "sum":
        raise RuntimeError("Not a valid aggregation!")
    return torch.cat((h, dst_feat), dim=-1)


# ----------------------------------------------------------------------------------------
# MLPs and blocks
# ----------------------------------------------------------------------------------------
def mesh_graph_mlp(sd: Dict[str, Tensor], prefix: str, x: Tensor, hidden_layers: Optional[int] = 2,
                   norm: bool = True, act: str = "relu") -> Tensor:
    """MeshGraphMLP.forward -- models/gnn_layers/mesh_graph_mlp.py:142-168,200-203.
    Linear(in,hid), act, [Linear(hid,hid), act] x (hidden_layers-1), Linear(hid,out), [LayerNorm]."""
    if hidden_layers is None:
        return x
    a = _ACT[act]
    for i in range(hidden_layers):
        x = a(F.linear(x, sd[f"{prefix}.model.{2 * i}.weight"], sd[f"{prefix}.model.{2 * i}.bias"]))
    j = 2 * hidden_layers
    x = F.linear(x, sd[f"{prefix}.model.{j}.weight"], sd[f"{prefix}.model.{j}.bias"])
    if norm:
        w, b = sd[f"{prefix}.model.{j + 1}.weight"], sd[f"{prefix}.model.{j + 1}.bias"]
        x = F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)
    return x


def edge_mlp_sum(sd: Dict[str, Tensor], prefix: str, efeat: Tensor, src_feat: Tensor, dst_feat: Tensor,
                 src: Tensor, dst: Tensor, hidden_layers: int = 2, norm: bool = True, act: str = "relu") -> Tensor:
    """MeshGraphEdgeMLPSum.forward ("concat trick") -- mesh_graph_mlp.py:390-430.
    Three bias-free per-source matmuls at node/edge level, bias folded into the dst one,
    gather-add, then the remaining layers (model = [act, (Linear, act)*, Linear, LN])."""
    a = _ACT[act]
    m_e = F.linear(efeat, sd[f"{prefix}.lin_efeat"], None)
    m_s = F.linear(src_feat, sd[f"{prefix}.lin_src"], None)
    m_d = F.linear(dst_feat, sd[f"{prefix}.lin_dst"], sd.get(f"{prefix}.bias"))
    x = a(sum_efeat(m_e, m_s, m_d, src, dst))
    idx = 1
    for _ in range(hidden_layers - 1):
        x = a(F.linear(x, sd[f"{prefix}.model.{idx}.weight"], sd[f"{prefix}.model.{idx}.bias"]))
        idx += 2
    x = F.linear(x, sd[f"{prefix}.model.{idx}.weight"], sd[f"{prefix}.model.{idx}.bias"])
    if norm:
        x = F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.model.{idx + 1}.weight"],
                         sd[f"{prefix}.model.{idx + 1}.bias"], 1e-5)
    return x


def mesh_edge_block(sd, prefix, efeat, nfeat, src, dst, hidden_layers=2, act="relu", concat_trick=False,
                    src_feat: Optional[Tensor] = None) -> Tensor:
    """MeshEdgeBlock.forward -- mesh_edge_block.py:88-96: efeat' = edge_mlp(...) + efeat.
    `src_feat` overrides the source-node table (distributed: rows after the halo exchange)."""
    s = nfeat if src_feat is None else src_feat
    if concat_trick:
        y = edge_mlp_sum(sd, f"{prefix}.edge_mlp", efeat, s, nfeat, src, dst, hidden_layers, True, act)
    else:
        y = mesh_graph_mlp(sd, f"{prefix}.edge_mlp", concat_efeat(efeat, s, nfeat, src, dst),
                           hidden_layers, True, act)
    return y + efeat


def mesh_node_block(sd, prefix, efeat, nfeat, dst, hidden_layers=2, act="relu", aggregation="sum") -> Tensor:
    """MeshNodeBlock.forward -- mesh_node_block.py:82-92: nfeat' = node_mlp(agg||nfeat) + nfeat."""
    cat = aggregate_and_concat(efeat, nfeat, dst, aggregation)
    return mesh_graph_mlp(sd, f"{prefix}.node_mlp", cat, hidden_layers, True, act) + nfeat


def meshgraphnet_forward(sd: Dict[str, Tensor], node_features: Tensor, edge_features: Tensor,
                         src: Tensor, dst: Tensor, processor_size: int = 15, act: str = "relu",
                         num_layers: int = 2, aggregation: str = "sum", concat_trick: bool = False,
                         halo=None) -> Tensor:
    """MeshGraphNet.forward -- models/meshgraphnet/meshgraphnet.py:206-217 and the processor
    loop :353-379 (edge block then node block, node update sees the updated edges).
    `halo(nfeat) -> src rows` emulates CuGraphCSC.get_src_node_features_in_local_graph."""
    e = mesh_graph_mlp(sd, "edge_encoder", edge_features, num_layers, True, act)
    n = mesh_graph_mlp(sd, "node_encoder", node_features, num_layers, True, act)
    for i in range(processor_size):
        s = None if halo is None else halo(n)
        e = mesh_edge_block(sd, f"processor.processor_layers.{2 * i}", e, n, src, dst, num_layers, act,
                            concat_trick, src_feat=s)
        n = mesh_node_block(sd, f"processor.processor_layers.{2 * i + 1}", e, n, dst, num_layers, act,
                            aggregation)
    return mesh_graph_mlp(sd, "node_decoder", n, num_layers, False, act)


# ----------------------------------------------------------------------------------------
# parameter construction that reproduces the reference's RNG stream
# ----------------------------------------------------------------------------------------
def _mlp_params(sd, prefix, in_dim, out_dim, hid, hidden_layers, norm):
    """Same construction order as MeshGraphMLP.__init__ (mesh_graph_mlp.py:142-168):
    nn.Linear default init consumes the global RNG weight-then-bias, layer by layer."""
    dims = [in_dim] + [hid] * hidden_layers + [out_dim]
    for i in range(hidden_layers + 1):
        lin = torch.nn.Linear(dims[i], dims[i + 1])
        sd[f"{prefix}.model.{2 * i}.weight"] = lin.weight.detach()
        sd[f"{prefix}.model.{2 * i}.bias"] = lin.bias.detach()
    if norm:
        j = 2 * hidden_layers + 1
        sd[f"{prefix}.model.{j}.weight"] = torch.ones(out_dim)
        sd[f"{prefix}.model.{j}.bias"] = torch.zeros(out_dim)


def make_state_dict(input_dim_nodes: int, input_dim_edges: int, output_dim: int, processor_size: int = 15,
                    hidden: int = 128, num_layers: int = 2) -> Dict[str, Tensor]:
    """Random-init state_dict drawn in the reference's construction order
    (meshgraphnet.py:162-203: edge_encoder, node_encoder, node_decoder, then ALL edge
    blocks, then ALL node blocks, :267-273) so `torch.manual_seed(s)` gives the very same
    weights as `MeshGraphNet(...)` built under the same seed."""
    sd: Dict[str, Tensor] = {}
    _mlp_params(sd, "edge_encoder", input_dim_edges, hidden, hidden, num_layers, True)
    _mlp_params(sd, "node_encoder", input_dim_nodes, hidden, hidden, num_layers, True)
    _mlp_params(sd, "node_decoder", hidden, output_dim, hidden, num_layers, False)
    for i in range(processor_size):
        _mlp_params(sd, f"processor.processor_layers.{2 * i}.edge_mlp", 3 * hidden, hidden, hidden, num_layers, True)
    for i in range(processor_size):
        _mlp_params(sd, f"processor.processor_layers.{2 * i + 1}.node_mlp", 2 * hidden, hidden, hidden, num_layers, True)
    return sd


# ----------------------------------------------------------------------------------------
# training step used as the CPU baseline (bench.py) and by the gradient parity tests
# ----------------------------------------------------------------------------------------
def step_fwd_bwd(sd: Dict[str, Tensor], node_features: Tensor, edge_features: Tensor, src: Tensor, dst: Tensor,
                 target: Tensor, **kw) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """zero_grad -> forward -> MSE -> backward (examples/cfd/vortex_shedding_mgn/train.py:151-166).
    Returns (prediction, loss, {name: grad}) incl. grads of the two input feature tables."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    nf = node_features.detach().clone().requires_grad_(True)
    ef = edge_features.detach().clone().requires_grad_(True)
    pred = meshgraphnet_forward(leaves, nf, ef, src, dst, **kw)
    loss = F.mse_loss(pred, target)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items()}
    grads["__node_features"] = nf.grad
    grads["__edge_features"] = ef.grad
    return pred.detach(), loss.detach(), grads


# ----------------------------------------------------------------------------------------
# halo exchange emulated in one process (distributed/utils.py:541-707)
# ----------------------------------------------------------------------------------------
def indexed_all_to_all_v_emulated(tensors: Sequence[Tensor], scatter_indices: Sequence[Sequence[Tensor]],
                                  sizes: List[List[int]]) -> List[Tensor]:
    """What every rank receives from indexed_all_to_all_v (utils.py:590-603):
    rank r gets cat_p( tensors[p][scatter_indices[p][r]] ) in rank order p = 0..P-1."""
    P = len(tensors)
    out = []
    for r in range(P):
        parts = [tensors[p][scatter_indices[p][r]] for p in range(P)]
        for p in range(P):
            assert parts[p].shape[0] == sizes[p][r]
        out.append(torch.cat(parts, dim=0))
    return out


# ----------------------------------------------------------------------------------------
# the steps either side of the path
# ----------------------------------------------------------------------------------------
def cells_to_bidirected_coo(cells: Tensor, num_nodes: int) -> Tuple[Tensor, Tensor]:
    """`cell_to_adj` + `create_graph` (datapipes/gnn/vortex_shedding_dataset.py:307-322): one directed edge per
    cell side (vertex j -> vertex j+1 mod 3), then `dgl.to_bidirected` (DGL v2.4, unpinned third party: adds
    every reverse edge and removes duplicates; the coalesced result is ordered by (src, dst))."""
    cells = cells.to(torch.int64)
    src = cells[:, [0, 1, 2]].reshape(-1)
    dst = cells[:, [1, 2, 0]].reshape(-1)
    s = torch.cat([src, dst])
    d = torch.cat([dst, src])
    key = torch.unique(s * num_nodes + d)
    return torch.div(key, num_nodes, rounding_mode="floor"), key % num_nodes


def edge_features(pos: Tensor, src: Tensor, dst: Tensor, mu: Optional[Tensor] = None,
                  std: Optional[Tensor] = None) -> Tensor:
    """`add_edge_features` (:324-333): cat(pos[src] - pos[dst], ||.||) and `normalize_edge` (:342-349):
    (x - mu) / std."""
    disp = pos[src.long()] - pos[dst.long()]
    x = torch.cat((disp, torch.linalg.norm(disp, dim=-1, keepdim=True)), dim=1)
    if mu is not None:
        x = (x - mu) / std
    return x


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-3, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0, adamw: bool = False) -> None:
    """In-place Adam update, the arithmetic of torch.optim.Adam's single-tensor path (torch 2.x
    optim/adam.py `_single_tensor_adam`, the optimizer of examples/cfd/vortex_shedding_mgn/train.py:122);
    `step` is the 1-based count AFTER this update.  Pinned against torch.optim.Adam / AdamW in
    tests/test_train_step.py."""
    if weight_decay != 0.0:
        if adamw:
            p.mul_(1.0 - lr * weight_decay)
        else:
            g = g + weight_decay * p
    m.lerp_(g, 1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def khop_halo_partition(offsets: Tensor, indices: Tensor, part_id: Tensor, num_partitions: int, halo_hops: int):
    """Pure-Python restatement of the k-hop-halo partitions the reference obtains from
    `dgl.metis_partition(graph, k, extra_cached_hops=halo_hops, reshuffle=True)`
    (examples/cfd/external_aerodynamics/xaeronet/surface/preprocessor.py:129-155; DGL v2.4
    `partition_graph_with_halo`, unpinned third party): starting from the nodes a piece owns, each hop adds all
    in-edges of the previous hop's nodes and their not-yet-present sources.  The node assignment (METIS in the
    reference) is an input.  Returns, per piece, python lists: node_ids (inner ascending, then halo ascending),
    inner flags, global edge rows and local (src, dst) pairs in local-CSC order.  Small graphs only."""
    off = offsets.tolist()
    idx = indices.tolist()
    pid = part_id.tolist()
    out = []
    for p in range(num_partitions):
        inner = [v for v in range(len(pid)) if pid[v] == p]
        members = set(inner)
        with_edges = set()
        frontier = list(inner)
        for _ in range(halo_hops):
            nxt = []
            for v in frontier:
                with_edges.add(v)
                for e in range(off[v], off[v + 1]):
                    u = idx[e]
                    if u not in members:
                        members.add(u)
                        nxt.append(u)
            frontier = nxt
        halo = sorted(members - set(inner))
        nodes = inner + halo
        local = {v: i for i, v in enumerate(nodes)}
        edge_rows, src_l, dst_l = [], [], []
        for v in nodes:
            if v in with_edges:
                for e in range(off[v], off[v + 1]):
                    edge_rows.append(e)
                    src_l.append(local[idx[e]])
                    dst_l.append(local[v])
        out.append(dict(node_ids=nodes, inner=[True] * len(inner) + [False] * len(halo), edge_ids=edge_rows,
                        src=src_l, dst=dst_l))
    return out


def rollout_reference_loop(model_fn, frames_x: Sequence[Tensor], edge_features: Tensor, mask: Tensor,
                           stats: Dict[str, Tensor]) -> List[Tensor]:
    """The prediction loop of examples/cfd/vortex_shedding_mgn/inference.py:90-150 restated statement by statement
    for ONE trajectory: `frames_x[i]` is the dataset's (normalised) node-feature frame i, `model_fn(x, efeat)` the
    network.  Returns the list of de-normalised predictions (u, v, p) the reference stores in `self.pred`."""
    def denorm(x, mu, std):  # VortexSheddingDataset.denormalize (vortex_shedding_dataset.py:351-355)
        return x * std + mu

    def norm(x, mu, std):    # normalize_node (:335-340)
        return (x - mu.expand(x.size())) / std.expand(x.size())

    preds: List[Tensor] = []
    for i, x in enumerate(frames_x):
        x = x.clone()
        x[:, 0:2] = denorm(x[:, 0:2], stats["velocity_mean"], stats["velocity_std"])
        invar = x.clone()
        if i != 0:
            invar[:, 0:2] = preds[i - 1][:, 0:2].clone()
        invar[:, 0:2] = norm(invar[:, 0:2], stats["velocity_mean"], stats["velocity_std"])
        pred_i = model_fn(invar, edge_features).detach().clone()
        pred_i[:, 0:2] = denorm(pred_i[:, 0:2], stats["velocity_diff_mean"], stats["velocity_diff_std"])
        pred_i[:, 2] = denorm(pred_i[:, 2], stats["pressure_mean"], stats["pressure_std"])
        invar[:, 0:2] = denorm(invar[:, 0:2], stats["velocity_mean"], stats["velocity_std"])
        m = torch.cat((mask, mask), dim=-1)
        pred_i[:, 0:2] = torch.where(m, pred_i[:, 0:2], torch.zeros_like(pred_i[:, 0:2]))
        preds.append(torch.cat(((pred_i[:, 0:2] + invar[:, 0:2]), pred_i[:, [2]]), dim=-1))
    return preds
