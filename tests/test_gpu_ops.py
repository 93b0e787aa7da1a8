"""GPU parity of the operator seam and the fp32 (SIMT) MeshGraphNet path against the CPU oracle
and the reference goldens.  All calls go through the C ABI (libmgn_b200.so)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def rel_err(a, b):
    """max-norm relative error: max|a-b| / max|b| (the metric of every tolerance below)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def O():
    from oracle import mgn_oracle

    return mgn_oracle


def _plan_from_csc(offsets, indices, n_src, n_dst):
    from modulus_b200.ops import GraphPlan

    return GraphPlan.from_csc(offsets.to(DEV), indices.to(DEV), n_src, n_dst)


@pytest.mark.parametrize("kind", ["mesh", "random", "powerlaw", "empty_rows"])
def test_plan_csr_bit_exact(O, kind):
    from modulus_b200.mesh import power_law_graph_csc, random_graph_csc, triangle_grid_mesh

    if kind == "mesh":
        m = triangle_grid_mesh(40, 37)
        off, idx, ns, nd = m["offsets"], m["indices"], m["num_nodes"], m["num_nodes"]
    elif kind == "random":
        off, idx = random_graph_csc(4321, 1234, 2, 8, seed=42)
        ns, nd = 4321, 1234
    elif kind == "powerlaw":  # long segments exercise the block-level sort
        off, idx = power_law_graph_csc(3000, 60000, seed=3)
        idx = (idx % 50)  # few sources -> CSR segments of >1000 edges
        ns, nd = 50, 3000
    else:
        off = torch.tensor([0, 0, 3, 3, 3, 5], dtype=torch.int64)
        idx = torch.tensor([4, 0, 4, 2, 0], dtype=torch.int64)
        ns, nd = 6, 5
    plan = _plan_from_csc(off, idx, ns, nd)
    src, dst = O.coo_from_csc(off, idx)
    co, ce, _ = O.csr_from_csc(off, idx, ns)
    assert torch.equal(plan.src.cpu().long(), src) and torch.equal(plan.dst.cpu().long(), dst)
    assert torch.equal(plan.csr_offsets.cpu().long(), co)
    assert torch.equal(plan.csr_eids.cpu().long(), ce)


def test_plan_from_coo_matches_stable_sort(O):
    from modulus_b200.ops import GraphPlan

    g = torch.Generator().manual_seed(5)
    src = torch.randint(0, 300, (2000,), generator=g)
    dst = torch.randint(0, 200, (2000,), generator=g)
    plan = GraphPlan.from_coo(src.to(DEV), dst.to(DEV), 300, 200)
    perm = torch.argsort(dst, stable=True)
    assert torch.equal(plan.csc_eids.cpu().long(), perm)
    assert torch.equal(plan.csc_offsets.cpu().long()[1:], torch.cumsum(torch.bincount(dst, minlength=200), 0))
    assert torch.equal(plan.csr_eids.cpu().long(), torch.argsort(src, stable=True))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("D", [128, 5])
@pytest.mark.parametrize("agg", ["sum", "mean"])
def test_seam_ops_fwd_bwd(O, dtype, tol, D, agg):
    from modulus_b200.mesh import random_graph_csc
    from modulus_b200.models.gnn_layers import CuGraphCSC, aggregate_and_concat, concat_efeat, sum_efeat

    ns, nd = 777, 513
    off, idx = random_graph_csc(ns, nd, 0, 9, seed=1)
    E = idx.numel()
    src, dst = O.coo_from_csc(off, idx)
    torch.manual_seed(0)
    ef, sf, df = torch.randn(E, D), torch.randn(ns, D), torch.randn(nd, D)
    graph = CuGraphCSC(off.to(DEV), idx.to(DEV), ns, nd)

    def leaves(*ts):
        return [t.detach().clone().to(DEV).to(dtype).requires_grad_(True) for t in ts]

    def cpu_leaves(*ts):
        return [t.detach().clone().to(dtype).float().requires_grad_(True) for t in ts]

    # concat_efeat (bipartite tuple form)
    e, s, d = leaves(ef, sf, df)
    out = concat_efeat(e, (s, d), graph)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    out.backward(w.to(DEV).to(dtype))
    ce, cs, cd = cpu_leaves(ef, sf, df)
    ref = O.concat_efeat(ce, cs, cd, src, dst)
    ref.backward(w.to(dtype).float())
    assert torch.equal(out.float().cpu(), ref.detach())  # pure copy: bit exact
    assert rel_err(e.grad, ce.grad) == 0
    assert rel_err(s.grad, cs.grad) < tol and rel_err(d.grad, cd.grad) < tol

    # sum_efeat
    e, s, d = leaves(ef, sf, df)
    out = sum_efeat(e, (s, d), graph)
    out.backward(w[:, :D].to(DEV).to(dtype))
    ce, cs, cd = cpu_leaves(ef, sf, df)
    ref = O.sum_efeat(ce, cs, cd, src, dst)
    ref.backward(w[:, :D].to(dtype).float())
    assert rel_err(out, ref) < tol
    assert rel_err(s.grad, cs.grad) < tol and rel_err(d.grad, cd.grad) < tol and rel_err(e.grad, ce.grad) < tol

    # aggregate_and_concat
    e, d = leaves(ef, df)
    out = aggregate_and_concat(e, d, graph, agg)
    w2 = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    out.backward(w2.to(DEV).to(dtype))
    ce, cd = cpu_leaves(ef, df)
    ref = O.aggregate_and_concat(ce, cd, dst, agg)
    ref.backward(w2.to(dtype).float())
    assert rel_err(out, ref) < tol
    assert rel_err(e.grad, ce.grad) < tol and rel_err(d.grad, cd.grad) < tol
    with pytest.raises(RuntimeError, match="Not a valid aggregation"):
        aggregate_and_concat(e, d, graph, "max")


def test_segment_sum_is_deterministic():
    from modulus_b200 import ops
    from modulus_b200.mesh import random_graph_csc

    off, idx = random_graph_csc(500, 400, 0, 40, seed=9)
    plan = _plan_from_csc(off, idx, 500, 400)
    x = torch.randn(idx.numel(), 128, device=DEV)
    a = ops.segment_sum(x, 0, 128, plan.csr_offsets, plan.csr_eids, 500)
    for _ in range(3):
        assert torch.equal(a, ops.segment_sum(x, 0, 128, plan.csr_offsets, plan.csr_eids, 500))


def _run_model(model, g, graph, dtype=torch.float32):
    nf = g["node_features"].to(DEV).requires_grad_(True)
    ef = g["edge_features"].to(DEV).requires_grad_(True)
    if dtype == torch.bfloat16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(nf, ef, graph)
    else:
        out = model(nf, ef, graph)
    loss = torch.nn.functional.mse_loss(out.float(), g["target"].to(DEV))
    loss.backward()
    return out, loss, nf.grad, ef.grad


@pytest.mark.parametrize("case", ["relu_sum", "relu_mean", "silu_sum", "concat_trick"])
def test_model_fp32_matches_reference_small(case):
    """fp32 path vs the unmodified reference: outputs and ALL gradients, 1e-4 relative."""
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    g = load_golden(f"ref_mgn_{case}.pt")
    model = MeshGraphNet(**g["kwargs"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    graph = CuGraphCSC(g["offsets"].to(DEV), g["indices"].to(DEV), g["n_nodes"], g["n_nodes"])
    out, loss, gnf, gef = _run_model(model, g, graph)
    assert rel_err(out, g["output"]) < 1e-4
    assert rel_err(gnf, g["grad_node_features"]) < 1e-4 and rel_err(gef, g["grad_edge_features"]) < 1e-4
    for k, p in model.named_parameters():
        assert rel_err(p.grad, g["grads"][k]) < 1e-4, k


@pytest.mark.parametrize("L,tol", [(1, 1e-4), (15, 1e-3)])
def test_model_fp32_default_size(L, tol):
    """north_star bar: 1e-4 after one message-passing layer, 1e-3 after 15 (hidden 128)."""
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    g = load_golden(f"ref_mgn_h128_L{L}.pt")
    torch.manual_seed(g["seed"])
    model = MeshGraphNet(6, 3, 3, processor_size=L).to(DEV)  # same RNG stream as the reference
    chk = float(sum(v.double().sum() for k, v in model.state_dict().items()))
    assert abs(chk - g["weight_checksum"]) < 1e-6
    graph = CuGraphCSC(g["offsets"].to(DEV), g["indices"].to(DEV), g["n_nodes"], g["n_nodes"])
    out, loss, gnf, gef = _run_model(model, g, graph)
    assert rel_err(out, g["output"]) < tol
    assert rel_err(gnf, g["grad_node_features"]) < tol and rel_err(gef, g["grad_edge_features"]) < tol
    named = dict(model.named_parameters())
    for k, v in g["grads_selected"].items():
        assert rel_err(named[k].grad, v) < tol, k
    for k, nrm in g["grad_norms"].items():
        assert abs(float(named[k].grad.double().norm()) - nrm) <= tol * max(nrm, 1e-8), k


def test_kat1_with_dgl_like_graph_in_edge_id_order():
    """The reference's own golden vector, graph passed as a DGL-like COO object (edge-id order)."""
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    g = load_golden("kat1_meshgraphnet_output.pt")

    class COOGraph:
        def __init__(self, src, dst, n):
            self._s, self._d, self._n = src, dst, n

        def edges(self):
            return self._s, self._d

        def num_src_nodes(self):
            return self._n

        def num_dst_nodes(self):
            return self._n

    torch.manual_seed(0)
    np.random.seed(0)
    model = MeshGraphNet(4, 3, 2).to(DEV)
    nf, ef = torch.randn(40, 4), torch.randn(20, 3)
    with torch.no_grad():
        out = model(nf.to(DEV), ef.to(DEV), COOGraph(g["src"], g["dst"], 40))
    assert torch.allclose(out.cpu(), g["output"], rtol=1e-3, atol=1e-3)  # the reference's tolerance
    assert rel_err(out, g["output"]) < 1e-4


def test_model_bf16_simt_path_within_2e2():
    """bf16 storage / fp32 accumulate vs the fp32 reference: 2e-2 (north_star)."""
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    g = load_golden("ref_mgn_relu_sum.pt")
    model = MeshGraphNet(**g["kwargs"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    graph = CuGraphCSC(g["offsets"].to(DEV), g["indices"].to(DEV), g["n_nodes"], g["n_nodes"])
    out, loss, gnf, gef = _run_model(model, g, graph, torch.bfloat16)
    assert out.dtype == torch.bfloat16
    assert rel_err(out, g["output"]) < 2e-2
    # Gradients of a random-init MGN are ill-conditioned in bf16 (ReLU mask flips): the REFERENCE's own
    # bf16 autocast path deviates from its fp32 path by ~1e-1 in relative L2 (measured by the oracle
    # under CPU autocast on the same inputs).  Bar: no worse than 1.5x that deviation.
    from oracle import mgn_oracle as O

    src, dst = O.coo_from_csc(g["offsets"], g["indices"])
    with torch.autocast("cpu", dtype=torch.bfloat16):
        _, _, ref16 = O.step_fwd_bwd(g["state_dict"], g["node_features"], g["edge_features"], src, dst, g["target"],
                                     processor_size=g["kwargs"]["processor_size"])
    ref_dev = l2_err(ref16["__node_features"], g["grad_node_features"])
    assert l2_err(gnf, g["grad_node_features"]) < 1.5 * ref_dev + 1e-3


def test_checkpoint_segments_give_same_result():
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    g = load_golden("ref_mgn_relu_sum.pt")
    kw = dict(g["kwargs"])
    m1 = MeshGraphNet(**kw).to(DEV)
    m1.load_state_dict(g["state_dict"])
    m2 = MeshGraphNet(**kw, num_processor_checkpoint_segments=2).to(DEV)
    m2.load_state_dict(g["state_dict"])
    graph = CuGraphCSC(g["offsets"].to(DEV), g["indices"].to(DEV), g["n_nodes"], g["n_nodes"])
    o1, _, a1, b1 = _run_model(m1, g, graph)
    o2, _, a2, b2 = _run_model(m2, g, graph)
    assert torch.equal(o1, o2) and torch.equal(a1, a2) and torch.equal(b1, b2)


@pytest.mark.parametrize("dtype,D", [(torch.float32, 128), (torch.bfloat16, 128), (torch.bfloat16, 512), (torch.float32, 40)])
def test_segment_sum_with_hub_segments(dtype, D):
    """Skewed degrees (BASELINE configs[4]): segments far longer than a warp's share go through the chunked worklist
    (mgn_segment_sum_balanced); same values as a float64 index_add, mean and accumulate included, run to run identical."""
    from modulus_b200 import ops
    from modulus_b200.mesh import power_law_graph_csc

    off, idx = power_law_graph_csc(3000, 60000, alpha=1.2, seed=1)
    plan = _plan_from_csc(off, idx, 3000, 3000)
    deg = (plan.csc_offsets[1:] - plan.csc_offsets[:-1]).long()
    assert int(deg.max()) > 4 * ops.LONG_SEGMENT and ops._has_long_segments(plan.csc_offsets)
    x = torch.randn(idx.numel(), D, device=DEV).to(dtype)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for offsets, eids, key in ((plan.csc_offsets, None, plan.dst), (plan.csr_offsets, plan.csr_eids, plan.src)):
        ref = torch.zeros(3000, D, dtype=torch.float64, device=DEV).index_add_(0, key.long(), x.double())
        out = ops.segment_sum(x, 0, D, offsets, eids, 3000)
        assert rel_err(out.float(), ref) < tol
        assert torch.equal(out, ops.segment_sum(x, 0, D, offsets, eids, 3000))
        d = (offsets[1:] - offsets[:-1]).clamp_min(1).double()[:, None]
        base = torch.ones(3000, D, device=DEV, dtype=dtype)
        out_m = ops.segment_sum(x, 0, D, offsets, eids, 3000, out=base.clone(), mean=True, accumulate=True)
        assert rel_err(out_m.float(), 1.0 + ref / d) < tol
