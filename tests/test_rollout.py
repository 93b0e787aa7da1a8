"""Autoregressive rollout (reference: examples/cfd/vortex_shedding_mgn/inference.py:90-150).
CPU: the device-resident loop of modulus_b200/rollout.py against a statement-by-statement restatement of the
reference loop, with the CPU oracle network standing in for the model.  GPU: product model, eager vs one CUDA graph
per step."""
import pytest
import torch

from oracle import mgn_oracle as O

DEV = "cuda"


def _setup(seed=0):
    from modulus_b200.mesh import triangle_grid_mesh

    mesh = triangle_grid_mesh(9, 10)
    n = mesh["num_nodes"]
    g = torch.Generator().manual_seed(seed)
    node_type = torch.nn.functional.one_hot(torch.randint(0, 4, (n,), generator=g), 4).float()
    x0 = torch.cat([torch.randn(n, 2, generator=g), node_type], 1)
    mask = (node_type[:, 0] + node_type[:, 1] > 0).reshape(-1, 1)       # "normal" and "inflow"-like nodes move
    # 1-D statistics, the shapes VortexSheddingDataset._get_node_stats produces (vortex_shedding_dataset.py:231-280)
    stats = dict(velocity_mean=torch.tensor([0.3, -0.1]), velocity_std=torch.tensor([1.5, 0.7]),
                 velocity_diff_mean=torch.tensor([0.01, 0.02]), velocity_diff_std=torch.tensor([0.05, 0.03]),
                 pressure_mean=torch.tensor([0.2]), pressure_std=torch.tensor([2.0]))
    return mesh, n, x0, mask, stats


def test_rollout_matches_reference_loop_on_cpu():
    from modulus_b200.rollout import rollout

    mesh, n, x0, mask, stats = _setup()
    torch.manual_seed(1)
    sd = O.make_state_dict(6, 3, 3, processor_size=2, hidden=16)
    src, dst = O.coo_from_csc(mesh["offsets"], mesh["indices"])

    def net(x, ef, graph=None):
        return O.meshgraphnet_forward(sd, x, ef, src, dst, processor_size=2)

    T = 5
    ours = rollout(net, None, x0, mesh["edge_features"], mask, stats, T)
    # the dataset would hand the reference one stored frame per step; only columns 2: of frames 1.. are used
    frames = [x0] + [torch.cat([torch.full((n, 2), float("nan")), x0[:, 2:]], 1) for _ in range(T - 1)]
    ref = O.rollout_reference_loop(lambda x, ef: net(x, ef), frames, mesh["edge_features"], mask, stats)
    assert ours.shape == (T, n, 3)
    for i in range(T):
        assert torch.allclose(ours[i], ref[i], rtol=1e-5, atol=1e-6), i
    frozen = ~mask.flatten()
    v0 = x0[:, :2] * stats["velocity_std"] + stats["velocity_mean"]
    assert torch.allclose(ours[-1][frozen, :2], v0[frozen], rtol=1e-6, atol=1e-6)   # wall / outflow nodes never move


def test_rollout_argument_errors():
    from modulus_b200.rollout import rollout

    mesh, n, x0, mask, stats = _setup()
    bad = dict(stats)
    del bad["pressure_std"]
    with pytest.raises(KeyError):
        rollout(lambda *a: None, None, x0, mesh["edge_features"], mask, bad, 1)
    with pytest.raises(ValueError):
        rollout(lambda *a: None, None, x0, mesh["edge_features"], mask, stats, -1)
    with pytest.raises(ValueError):
        rollout(lambda *a: None, None, x0, mesh["edge_features"], mask, stats, 1, use_graphs=True)
    assert rollout(lambda *a: None, None, x0, mesh["edge_features"], mask, stats, 0).shape == (0, n, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("use_amp", [False, True])
def test_rollout_one_graph_launch_per_step_matches_eager(use_amp):
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from modulus_b200.rollout import rollout

    mesh, n, x0, mask, stats = _setup()
    torch.manual_seed(2)
    model = MeshGraphNet(6, 3, 3, processor_size=2).to(DEV).eval()
    graph = CuGraphCSC(mesh["offsets"].to(DEV), mesh["indices"].to(DEV), n, n)
    ef = mesh["edge_features"].to(DEV)
    T = 6
    eager = rollout(model, graph, x0.to(DEV), ef, mask, stats, T, use_graphs=False, use_amp=use_amp)
    graphed = rollout(model, graph, x0.to(DEV), ef, mask, stats, T, use_graphs=True, use_amp=use_amp)
    torch.cuda.synchronize()
    assert torch.isfinite(eager).all() and torch.equal(eager, graphed)
    if not use_amp:   # fp32 product path against the CPU oracle network with the same weights
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        src, dst = O.coo_from_csc(mesh["offsets"], mesh["indices"])
        ref = rollout(lambda x, e, g: O.meshgraphnet_forward(sd, x, e, src, dst, processor_size=2), None, x0,
                      mesh["edge_features"], mask, stats, T)
        assert torch.allclose(eager.cpu(), ref, rtol=1e-3, atol=1e-4)


def _golden():
    import os

    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "ref_rollout.pt"), weights_only=True)


def test_rollout_golden_of_the_unmodified_reference_loop_on_cpu():
    """`ref_rollout.pt` is what the reference's own `MGNRollout.predict` stored in `self.pred` (tests/golden/
    make_golden_rollout.py: unmodified script class, unmodified reference network and dataset statics).  Pins (1) the
    oracle's restatement of the loop and (2) the product loop, both driving the oracle network with the golden's weights."""
    from modulus_b200.rollout import rollout

    g = _golden()
    sd, T = g["state_dict"], g["steps"]
    src, dst = O.coo_from_csc(g["offsets"], g["indices"])
    net = lambda x, ef, graph=None: O.meshgraphnet_forward(sd, x, ef, src, dst, processor_size=2)  # noqa: E731
    frames = [g["frames"][i] for i in range(T)]
    ref = O.rollout_reference_loop(lambda x, ef: net(x, ef), frames, g["edge_features"], g["mask"], g["stats"])
    for i in range(T):
        assert torch.allclose(ref[i], g["pred"][i], rtol=1e-5, atol=1e-6), i
    ours = rollout(net, None, frames[0], g["edge_features"], g["mask"], g["stats"], T)
    assert torch.allclose(ours, g["pred"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_rollout_golden_of_the_unmodified_reference_loop_on_gpu():
    """Product model (reference state_dict loaded as is) + device-resident loop against the same golden: six
    autoregressive steps, fp32."""
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from modulus_b200.rollout import rollout

    g = _golden()
    model = MeshGraphNet(**g["kwargs"])
    model.load_state_dict(g["state_dict"])
    model = model.to(DEV).eval()
    n = g["n_nodes"]
    graph = CuGraphCSC(g["offsets"].to(DEV), g["indices"].to(DEV), n, n)
    for use_graphs in (False, True):
        out = rollout(model, graph, g["frames"][0].to(DEV), g["edge_features"].to(DEV), g["mask"], g["stats"], g["steps"],
                      use_graphs=use_graphs)
        torch.cuda.synchronize()
        assert torch.allclose(out.cpu(), g["pred"], rtol=1e-4, atol=1e-4), (use_graphs, float((out.cpu() - g["pred"]).abs().max()))
