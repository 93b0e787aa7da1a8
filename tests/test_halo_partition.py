"""Offline k-hop-halo partitions (reference: xaeronet/surface/preprocessor.py:129-155,
train.py:183-208).  CPU: index maps bit-exact against the pure-Python oracle, and the two properties the strategy
rests on, checked with the CPU oracle in float64: (1) with halo_hops >= layers the inner-node outputs of a piece equal
those of the full graph, (2) gradient accumulation over the pieces reproduces the full-graph gradient.
GPU: the same two properties through the product model and the C ABI."""
import pytest
import torch

from oracle import mgn_oracle as O

DEV = "cuda"


def _graphs():
    from modulus_b200.mesh import random_graph_csc, triangle_grid_mesh

    mesh = triangle_grid_mesh(9, 11)
    yield "mesh", mesh["offsets"], mesh["indices"], mesh["coords"]
    off, idx = random_graph_csc(57, 57, 0, 5, seed=3)
    yield "random", off, idx, torch.rand(57, 2, generator=torch.Generator().manual_seed(0))


@pytest.mark.parametrize("hops", [0, 1, 2, 4])
@pytest.mark.parametrize("P", [1, 3, 4])
def test_halo_partitions_bit_exact_against_oracle(P, hops):
    from modulus_b200.models.gnn_layers import partition_ids_by_slabs, partition_with_halo

    for name, off, idx, coords in _graphs():
        part = partition_ids_by_slabs(coords, P)
        assert int(part.max()) == P - 1 and torch.bincount(part, minlength=P).min() > 0
        pieces = partition_with_halo(off, idx, part, P, hops)
        ref = O.khop_halo_partition(off, idx, part, P, hops)
        owned = torch.zeros(off.numel() - 1, dtype=torch.int64)
        for pc, r in zip(pieces, ref):
            assert pc.node_ids.tolist() == r["node_ids"], name
            assert pc.inner_node.tolist() == r["inner"] and pc.num_inner == sum(r["inner"])
            assert pc.edge_ids.tolist() == r["edge_ids"]
            src, dst = O.coo_from_csc(pc.offsets, pc.indices)
            assert src.tolist() == r["src"] and dst.tolist() == r["dst"]
            # id maps lead back to the global tables
            assert torch.equal(pc.node_ids[pc.indices], idx[pc.edge_ids])
            owned[pc.node_ids[pc.inner_node]] += 1
        assert torch.equal(owned, torch.ones_like(owned))  # every node is inner in exactly one piece


def test_halo_partition_argument_errors():
    from modulus_b200.models.gnn_layers import partition_with_halo

    off = torch.tensor([0, 1, 2, 3])
    idx = torch.tensor([1, 2, 0])
    with pytest.raises(ValueError):
        partition_with_halo(off, idx, torch.tensor([0, 1]), 2, 1)
    with pytest.raises(ValueError):
        partition_with_halo(off, idx, torch.tensor([0, 1, 2]), 2, 1)
    with pytest.raises(ValueError):
        partition_with_halo(off, idx, torch.tensor([0, 1, 1]), 2, -1)
    with pytest.raises(RuntimeError):
        partition_with_halo(off, idx, torch.tensor([0, 0, 0]), 2, 1)


def _case(L=2):
    from modulus_b200.mesh import triangle_grid_mesh
    from modulus_b200.models.gnn_layers import partition_ids_by_slabs, partition_with_halo

    mesh = triangle_grid_mesh(12, 13)
    n = mesh["num_nodes"]
    torch.manual_seed(3)
    nf, ef, tgt = torch.randn(n, 5), mesh["edge_features"].clone(), torch.randn(n, 2)
    pieces = partition_with_halo(mesh["offsets"], mesh["indices"], partition_ids_by_slabs(mesh["coords"], 3), 3, L)
    return mesh, n, nf, ef, tgt, pieces


def test_oracle_piecewise_training_equals_full_graph_training():
    L = 2
    mesh, n, nf, ef, tgt, pieces = _case(L)
    torch.manual_seed(4)
    sd = {k: v.double() for k, v in O.make_state_dict(5, 3, 2, processor_size=L, hidden=16).items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    full_sd = {**sd, **params}
    src, dst = O.coo_from_csc(mesh["offsets"], mesh["indices"])
    out = O.meshgraphnet_forward(full_sd, nf.double(), ef.double(), src, dst, processor_size=L)
    loss = ((out - tgt.double()) ** 2).sum() / n
    g_full = torch.autograd.grad(loss, list(params.values()))
    g_acc = [torch.zeros_like(g) for g in g_full]
    for pc in pieces:
        s, d = O.coo_from_csc(pc.offsets, pc.indices)
        o = O.meshgraphnet_forward(full_sd, nf.double()[pc.node_ids], ef.double()[pc.edge_ids], s, d, processor_size=L)
        inner = pc.inner_node
        assert torch.allclose(o[inner], out[pc.node_ids[inner]], rtol=1e-10, atol=1e-12)  # property (1)
        part_loss = ((o[inner] - tgt.double()[pc.node_ids[inner]]) ** 2).sum() / n
        for a, g in zip(g_acc, torch.autograd.grad(part_loss, list(params.values()))):
            a += g
    for a, g in zip(g_acc, g_full):
        assert torch.allclose(a, g, rtol=1e-9, atol=1e-12)                                # property (2)


def test_oracle_too_small_halo_changes_the_inner_outputs():
    """Negative control: one hop of halo is not enough for two message-passing layers."""
    from modulus_b200.models.gnn_layers import partition_ids_by_slabs, partition_with_halo

    mesh, n, nf, ef, tgt, _ = _case(2)
    pieces = partition_with_halo(mesh["offsets"], mesh["indices"], partition_ids_by_slabs(mesh["coords"], 3), 3, 1)
    torch.manual_seed(4)
    sd = {k: v.double() for k, v in O.make_state_dict(5, 3, 2, processor_size=2, hidden=16).items()}
    src, dst = O.coo_from_csc(mesh["offsets"], mesh["indices"])
    out = O.meshgraphnet_forward(sd, nf.double(), ef.double(), src, dst, processor_size=2)
    pc = pieces[1]
    s, d = O.coo_from_csc(pc.offsets, pc.indices)
    o = O.meshgraphnet_forward(sd, nf.double()[pc.node_ids], ef.double()[pc.edge_ids], s, d, processor_size=2)
    assert not torch.allclose(o[pc.inner_node], out[pc.node_ids[pc.inner_node]], rtol=1e-6, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("use_bf16", [False, True])
def test_piecewise_training_equals_full_graph_training_on_the_product_path(use_bf16):
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    L = 2
    mesh, n, nf, ef, tgt, pieces = _case(L)
    torch.manual_seed(5)
    model = MeshGraphNet(5, 3, 2, processor_size=L).to(DEV)
    nf, ef, tgt = nf.to(DEV), ef.to(DEV), tgt.to(DEV)
    graph = CuGraphCSC(mesh["offsets"].to(DEV), mesh["indices"].to(DEV), n, n)

    def run(g, x, e):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_bf16):
            return model(x, e, g).float()

    model.zero_grad(set_to_none=True)
    out = run(graph, nf, ef)
    (((out - tgt) ** 2).sum() / n).backward()
    g_full = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    for pc in pieces:
        pc = pc.to(DEV)
        o = run(pc.graph(), nf[pc.node_ids], ef[pc.edge_ids])
        inner = pc.inner_node
        tol = dict(rtol=2e-2, atol=2e-2) if use_bf16 else dict(rtol=1e-4, atol=1e-5)
        assert torch.allclose(o[inner], out[pc.node_ids[inner]].detach(), **tol)
        (((o[inner] - tgt[pc.node_ids[inner]]) ** 2).sum() / n).backward()   # accumulates into .grad
    for k, p in model.named_parameters():
        err = float((p.grad - g_full[k]).norm() / g_full[k].norm().clamp_min(1e-12))
        assert err < (1e-1 if use_bf16 else 1e-4), (k, err)
