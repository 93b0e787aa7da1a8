"""The steps either side of the message-passing path:
  * mesh cells -> bidirected graph -> edge features -> normalisation
    (reference: datapipes/gnn/vortex_shedding_dataset.py:307-349; golden tests/golden/ref_datapipe_graph.pt)
  * FusedAdam, one multi-tensor launch (reference: examples/cfd/vortex_shedding_mgn/train.py:111-123 =
    torch.optim.Adam / apex FusedAdam).
CPU part pins the oracle; GPU part checks the CUDA path through the C ABI."""
import copy
import os

import pytest
import torch

from conftest import load_golden
from oracle import mgn_oracle as O

DEV = "cuda"


# ------------------------------------------------------------------------------ oracle pins (CPU)
@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_datapipe_matches_reference(dim):
    g = load_golden("ref_datapipe_graph.pt")
    c = g[f"dim{dim}"]
    n = c["pos"].shape[0]
    src, dst = O.cells_to_bidirected_coo(g["cells"], n)
    assert torch.equal(src, c["src"].long()) and torch.equal(dst, c["dst"].long())  # integer work: bit exact
    assert torch.equal(O.edge_features(c["pos"], src, dst), c["edge_features"])
    assert torch.equal(O.edge_features(c["pos"], src, dst, c["mu"], c["std"]), c["edge_features_normalized"])


def test_graph_from_cells_same_edge_set_as_reference():
    """Product host function (pure index arithmetic, runs on CPU too): same edge set, CSC order."""
    from modulus_b200.mesh import graph_from_cells

    g = load_golden("ref_datapipe_graph.pt")
    c = g["dim2"]
    n = c["pos"].shape[0]
    offsets, indices = graph_from_cells(g["cells"], n)
    src, dst = O.coo_from_csc(offsets, indices)
    ours = torch.unique(dst * n + src)
    ref = torch.unique(c["dst"].long() * n + c["src"].long())
    assert indices.numel() == c["src"].numel() and torch.equal(ours, ref)
    # in-edges of a node are sorted by source id: CSC order == sort by (dst, src)
    assert torch.equal(dst * n + src, ours)


@pytest.mark.parametrize("wd,adamw", [(0.0, False), (0.01, False), (0.01, True)])
def test_oracle_adam_matches_torch(wd, adamw):
    torch.manual_seed(0)
    p0 = torch.randn(257)
    ref_p = p0.clone().requires_grad_(True)
    opt = (torch.optim.AdamW if adamw else torch.optim.Adam)([ref_p], lr=3e-3, weight_decay=wd, foreach=False)
    p, m, v = p0.clone(), torch.zeros(257), torch.zeros(257)
    for step in range(1, 6):
        g = torch.randn(257)
        ref_p.grad = g.clone()
        opt.step()
        O.adam_step(p, g, m, v, step, lr=3e-3, weight_decay=wd, adamw=adamw)
        assert torch.allclose(p, ref_p.detach(), rtol=1e-6, atol=1e-7)


def test_fused_adam_argument_errors():
    from modulus_b200.optim import FusedAdam

    w = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(RuntimeError):
        FusedAdam([w], amsgrad=True)
    with pytest.raises(ValueError):
        FusedAdam([w], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([w], betas=(1.0, 0.9))
    opt = FusedAdam([w])
    w.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    opt.param_groups[0]["maximize"] = True      # e.g. carried in by load_state_dict from torch.optim.Adam(maximize=True)
    with pytest.raises(RuntimeError, match="maximize"):
        opt.step()


# ------------------------------------------------------------------------------ CUDA path (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
def test_edge_features_kernel_matches_reference(dim):
    from modulus_b200.mesh import edge_features, graph_from_cells
    from modulus_b200.ops import GraphPlan

    g = load_golden("ref_datapipe_graph.pt")
    c = g[f"dim{dim}"]
    n = c["pos"].shape[0]
    # reference edge order
    out = edge_features(c["pos"].to(DEV), c["src"].to(DEV), c["dst"].to(DEV))
    assert torch.allclose(out.cpu(), c["edge_features"], rtol=1e-6, atol=1e-7)
    outn = edge_features(c["pos"].to(DEV), c["src"].to(DEV), c["dst"].to(DEV), c["mu"].to(DEV), c["std"].to(DEV))
    assert torch.allclose(outn.cpu(), c["edge_features_normalized"], rtol=1e-5, atol=1e-6)
    # whole device pipeline: cells -> CSC -> plan -> features, against the oracle on the same edges
    offsets, indices = graph_from_cells(g["cells"].to(DEV), n)
    plan = GraphPlan.from_csc(offsets, indices, n, n)
    out2 = edge_features(c["pos"].to(DEV), plan.src, plan.dst, c["mu"].to(DEV), c["std"].to(DEV))
    ref2 = O.edge_features(c["pos"], plan.src.cpu(), plan.dst.cpu(), c["mu"], c["std"])
    assert torch.allclose(out2.cpu(), ref2, rtol=1e-5, atol=1e-6)
    with pytest.raises(AssertionError):
        edge_features(c["pos"].to(DEV), plan.src, plan.dst, c["mu"][:2].to(DEV), c["std"].to(DEV))


@pytest.mark.gpu
def test_edge_features_large_roundtrip_property():
    """1 M-node torus: features of edge (u->v) are minus those of (v->u) in the displacement columns and equal in
    the norm column (size-independent property at the BASELINE size)."""
    from modulus_b200.mesh import edge_features, torus_surface_mesh
    from modulus_b200.ops import GraphPlan

    mesh = torus_surface_mesh(1000, 1000, device=DEV)
    n = mesh["num_nodes"]
    plan = GraphPlan.from_csc(mesh["offsets"], mesh["indices"], n, n)
    ef = edge_features(mesh["coords"], plan.src, plan.dst)
    assert torch.allclose(ef, mesh["edge_features"], rtol=1e-5, atol=1e-6)
    key = plan.dst.long() * n + plan.src.long()      # sorted (CSC order)
    rkey = plan.src.long() * n + plan.dst.long()     # key of the reverse edge
    pos = torch.searchsorted(key, rkey)
    assert torch.equal(key[pos], rkey)
    assert torch.equal(ef[pos, :3], -ef[:, :3]) and torch.equal(ef[pos, 3], ef[:, 3])


@pytest.mark.gpu
@pytest.mark.parametrize("wd,adamw", [(0.0, False), (0.01, False), (0.01, True)])
def test_fused_adam_matches_torch_adam(wd, adamw):
    """Six steps of the one-launch kernel against torch.optim.Adam / AdamW on the same parameters and gradients.  The
    per-step check is against the float64 trajectory of torch's optimizer (the more accurate oracle: a float32 CPU
    trajectory carries its own rounding); the float32 CPU trajectory is kept for the moments and the state_dict layout."""
    from modulus_b200.optim import FusedAdam

    torch.manual_seed(1)
    shapes = [(128, 384), (128,), (3, 5), (1,), (4099,), (128, 128)]
    ours = [torch.nn.Parameter(torch.randn(*s, device=DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ours]
    ref64 = [torch.nn.Parameter(p.detach().cpu().double()) for p in ours]
    opt = FusedAdam(ours, lr=3e-3, weight_decay=wd, adam_w_mode=adamw)
    cls = torch.optim.AdamW if adamw else torch.optim.Adam
    ropt = cls(ref, lr=3e-3, weight_decay=wd, foreach=False)
    ropt64 = cls(ref64, lr=3e-3, weight_decay=wd, foreach=False)
    for step in range(6):
        for i, (p, r, r64) in enumerate(zip(ours, ref, ref64)):
            if step == 5 and i == 2:
                p.grad, r.grad, r64.grad = None, None, None  # a parameter without a gradient is skipped (last step: torch
                # would not advance this parameter's own step count, the fused counter is per group)
                continue
            gr = torch.randn(*r.shape)
            r.grad, r64.grad = gr, gr.double()
            p.grad = gr.to(DEV)                      # fresh tensor every step: the pointer table is refreshed
        opt.step()
        ropt.step()
        ropt64.step()
        for i, (p, r, r64) in enumerate(zip(ours, ref, ref64)):
            got, want = p.detach().cpu().double(), r64.detach()
            ok = torch.allclose(got, want, rtol=2e-6, atol=2e-7)
            if os.environ.get("MGN_ADAM_DUMP") and not ok:
                torch.save({"step": step, "i": i, "grad": r.grad, "got": got, "want64": want.clone(), "want32": r.detach().clone(),
                            "exp_avg": opt.state[p]["exp_avg"].cpu(), "exp_avg_sq": opt.state[p]["exp_avg_sq"].cpu(),
                            "ref_exp_avg": ropt64.state[r64]["exp_avg"].clone(),
                            "ref_exp_avg_sq": ropt64.state[r64]["exp_avg_sq"].clone()},
                           os.path.join(os.environ["MGN_ADAM_DUMP"], f"adam_fail_{os.getpid()}.pt"))
            assert ok, (step, i, tuple(p.shape), float((got - want).abs().max()),
                        float((r.detach().double() - want).abs().max()), float(opt.state[p]["step"]))
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sd["state"][0]["step"]) == 6.0
    assert torch.allclose(sd["state"][0]["exp_avg"].cpu(), ropt.state_dict()["state"][0]["exp_avg"], rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_fused_adam_state_dict_roundtrip_with_torch_adam():
    """A torch.optim.Adam checkpoint continues under FusedAdam with the same trajectory."""
    from modulus_b200.optim import FusedAdam

    torch.manual_seed(2)
    w = torch.randn(300, device=DEV)
    a = torch.nn.Parameter(w.clone())
    b = torch.nn.Parameter(w.clone())
    ta = torch.optim.Adam([a], lr=1e-2)
    grads = [torch.randn(300, device=DEV) for _ in range(6)]
    for g in grads[:3]:
        a.grad = g.clone()
        ta.step()
    fb = FusedAdam([b], lr=1e-2)
    b.data.copy_(a.data)
    fb.load_state_dict(copy.deepcopy(ta.state_dict()))  # load_state_dict aliases same-device tensors
    for g in grads[3:]:
        a.grad = g.clone()
        ta.step()
        b.grad = g.clone()
        fb.step()
    assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)


@pytest.mark.gpu
def test_fused_adam_skip_on_found_inf_and_inv_scale():
    from modulus_b200.optim import FusedAdam

    torch.manual_seed(3)
    p = torch.nn.Parameter(torch.randn(1000, device=DEV))
    r = torch.nn.Parameter(p.detach().clone())
    opt, ropt = FusedAdam([p], lr=1e-2), torch.optim.Adam([r], lr=1e-2)
    g = torch.randn(1000, device=DEV)
    p.grad = g * 8.0
    before = p.detach().clone()
    opt.step(found_inf=torch.ones(1, device=DEV), inv_scale=torch.full((1,), 0.125, device=DEV))
    assert torch.equal(p.detach(), before) and float(opt.state[p]["step"]) == 0.0
    opt.step(found_inf=torch.zeros(1, device=DEV), inv_scale=torch.full((1,), 0.125, device=DEV))
    r.grad = g
    ropt.step()
    assert torch.allclose(p, r, rtol=2e-6, atol=2e-7)


@pytest.mark.gpu
def test_training_step_with_fused_adam_in_a_cuda_graph():
    """zero_grad -> forward -> loss -> backward -> FusedAdam.step captured by hand as ONE CUDA graph with static
    gradients (zero_grad(set_to_none=False)) and replayed: same parameters as the eager loop."""
    from modulus_b200.mesh import triangle_grid_mesh
    from modulus_b200.models.gnn_layers import CuGraphCSC
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from modulus_b200.optim import FusedAdam

    mesh = triangle_grid_mesh(20, 21, device=DEV)
    n = mesh["num_nodes"]
    graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n, n)
    torch.manual_seed(4)
    nf, ef, tgt = torch.randn(n, 6, device=DEV), mesh["edge_features"], torch.randn(n, 3, device=DEV)

    def make():
        torch.manual_seed(5)
        model = MeshGraphNet(6, 3, 3, processor_size=3).to(DEV)
        return model, FusedAdam(model.parameters(), lr=1e-3)

    def step(model, opt):
        opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(nf, ef, graph)
        loss = torch.nn.functional.mse_loss(out.float(), tgt)
        loss.backward()
        opt.step()
        return loss

    eager, eopt = make()
    for _ in range(5):
        step(eager, eopt)

    model, opt = make()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):                      # warm-up on a side stream: plans, workspaces, static gradients
            step(model, opt)
    torch.cuda.current_stream().wait_stream(s)
    opt.prepare_replay()                        # learning rate -> device scalar the captured step reads
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        loss = step(model, opt)
    for _ in range(3):                          # capture itself executes nothing: 2 warm-up + 3 replays = 5 steps
        cg.replay()
    torch.cuda.synchronize()
    assert float(opt.state[next(model.parameters())]["step"]) == 5.0
    assert torch.isfinite(loss).all()
    for (k, a), b in zip(eager.named_parameters(), model.parameters()):
        # bf16 forward/backward: replays follow the same kernels, so only atomics-free summation order could differ
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5), k


# ------------------------------------------------------------------------------ StaticCaptureTraining
def test_static_capture_argument_errors():
    from modulus_b200.capture import StaticCaptureTraining

    lin = torch.nn.Linear(2, 2)
    opt = torch.optim.Adam(lin.parameters())
    with pytest.raises(ValueError):
        StaticCaptureTraining(model=lin, optim=opt, compile=True)
    with pytest.raises(ValueError):
        StaticCaptureTraining(model=lin, optim=opt, amp_type=torch.float32)
    with pytest.raises(ValueError):
        StaticCaptureTraining(model="not a module", optim=opt)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        StaticCaptureTraining(model=lin, optim=opt)


def _capture_case(n_layers=3):
    from modulus_b200.mesh import triangle_grid_mesh
    from modulus_b200.models.gnn_layers import CuGraphCSC

    mesh = triangle_grid_mesh(20, 21, device=DEV)
    n = mesh["num_nodes"]
    graph = CuGraphCSC(mesh["offsets"], mesh["indices"], n, n)
    g = torch.Generator().manual_seed(4)
    data = [(torch.randn(n, 6, generator=g).to(DEV), torch.randn(n, 3, generator=g).to(DEV)) for _ in range(7)]
    return graph, mesh["edge_features"], data


@pytest.mark.gpu
@pytest.mark.parametrize("use_amp,fused_opt", [(True, True), (False, True), (True, False)])
def test_static_capture_training_matches_eager_loop(use_amp, fused_opt):
    """The reference's training-step decorator (utils/capture.py:341): 2 eager warm-up calls, one recording call,
    then replays, with a LambdaLR scheduler stepping between calls and new data copied into static inputs.  Same
    parameters as the plain loop; with FusedAdam the optimizer step is inside the graph."""
    from modulus_b200.capture import StaticCaptureTraining
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from modulus_b200.optim import FusedAdam

    graph, ef, data = _capture_case()

    def make():
        torch.manual_seed(5)
        model = MeshGraphNet(6, 3, 3, processor_size=3).to(DEV)
        opt = FusedAdam(model.parameters(), lr=1e-3) if fused_opt else torch.optim.Adam(model.parameters(), lr=1e-3)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda e: 0.8 ** e)
        return model, opt, sched

    eager, eopt, esched = make()
    for nf, tgt in data:
        eopt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_amp):
            loss_e = torch.nn.functional.mse_loss(eager(nf, ef, graph).float(), tgt)
        loss_e.backward()
        eopt.step()
        esched.step()

    model, opt, sched = make()
    s_nf, s_tgt = torch.empty_like(data[0][0]), torch.empty_like(data[0][1])

    @StaticCaptureTraining(model=model, optim=opt, use_amp=use_amp, cuda_graph_warmup=2)
    def training_step(nf, tgt):
        return torch.nn.functional.mse_loss(model(nf, ef, graph).float(), tgt)

    for nf, tgt in data:
        s_nf.copy_(nf)
        s_tgt.copy_(tgt)
        loss = training_step(s_nf, s_tgt)
        sched.step()
    torch.cuda.synchronize()
    assert torch.allclose(loss, loss_e.detach(), rtol=1e-3, atol=1e-6)
    tol = dict(rtol=2e-3, atol=2e-5) if use_amp else dict(rtol=1e-4, atol=1e-6)
    for (k, a), b in zip(eager.named_parameters(), model.parameters()):
        assert torch.allclose(a, b, **tol), k
    if fused_opt:
        assert float(opt.state[next(model.parameters())]["step"]) == float(len(data))


@pytest.mark.gpu
@pytest.mark.parametrize("fused_opt", [True, False])
def test_float16_autocast_with_gradscaler_recipe(fused_opt):
    """The reference recipe's AMP switch (examples/cfd/vortex_shedding_mgn/train.py:153-166): float16 autocast +
    GradScaler.  The model accepts it (computing in bf16 storage, returning float16), the loss scale -- a power of two --
    is applied and removed exactly as scaling / unscaling by hand does; with FusedAdam the scaler's `grad_scale` /
    `found_inf` tensors are consumed on the device (`_step_supports_amp_scaling`), and an overflowing step is skipped."""
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from modulus_b200.optim import FusedAdam

    graph, ef, data = _capture_case()

    def make():
        torch.manual_seed(5)
        model = MeshGraphNet(6, 3, 3, processor_size=3).to(DEV)
        opt = FusedAdam(model.parameters(), lr=1e-3) if fused_opt else torch.optim.Adam(model.parameters(), lr=1e-3)
        return model, opt

    ref, ropt = make()  # same float16 autocast region, loss scaled and gradients unscaled by hand
    for nf, tgt in data[:4]:
        ropt.zero_grad(set_to_none=True)
        with torch.autocast("cuda"):
            loss_r = torch.nn.functional.mse_loss(ref(nf, ef, graph).float(), tgt)
        (loss_r * 2.0 ** 12).backward()
        for p in ref.parameters():
            p.grad.mul_(2.0 ** -12)
        ropt.step()

    model, opt = make()
    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 12)
    for nf, tgt in data[:4]:
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda"):  # float16, the CUDA default
            out = model(nf, ef, graph)
            assert out.dtype == torch.float16
            loss = torch.nn.functional.mse_loss(out.float(), tgt)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
    torch.cuda.synchronize()
    assert float(scaler.get_scale()) == 2.0 ** 12  # no step was skipped
    # the scaler's protocol (device-side unscale + overflow check inside the optimizer) == scaling and unscaling by hand
    for (k, a), b in zip(ref.named_parameters(), model.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), k
    # an overflowing backward pass: the step is skipped, the scale halves
    before = [p.detach().clone() for p in model.parameters()]
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda"):
        loss = torch.nn.functional.mse_loss(model(data[4][0], ef, graph).float(), data[4][1])
    scaler.scale(loss * float("inf")).backward()
    scaler.step(opt)
    scaler.update()
    torch.cuda.synchronize()
    assert float(scaler.get_scale()) == 2.0 ** 11
    for a, b in zip(before, model.parameters()):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_static_capture_training_float16_amp_is_one_graph_with_the_scaler_inside():
    """StaticCaptureTraining(amp_type=torch.float16) -- the reference's default AMP type (utils/capture.py:341-436): the
    GradScaler protocol runs inside the recorded step (FusedAdam consumes the scale on the device) and the result follows
    the eager float16 + GradScaler loop."""
    from modulus_b200.capture import StaticCaptureTraining
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from modulus_b200.optim import FusedAdam

    graph, ef, data = _capture_case()

    def make():
        torch.manual_seed(5)
        model = MeshGraphNet(6, 3, 3, processor_size=3).to(DEV)
        return model, FusedAdam(model.parameters(), lr=1e-3)

    eager, eopt = make()
    scaler = torch.amp.GradScaler("cuda")
    for nf, tgt in data:
        eopt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            loss_e = torch.nn.functional.mse_loss(eager(nf, ef, graph).float(), tgt)
        scaler.scale(loss_e).backward()
        scaler.step(eopt)
        scaler.update()

    model, opt = make()
    s_nf, s_tgt = torch.empty_like(data[0][0]), torch.empty_like(data[0][1])

    @StaticCaptureTraining(model=model, optim=opt, use_amp=True, amp_type=torch.float16, cuda_graph_warmup=2)
    def training_step(nf, tgt):
        return torch.nn.functional.mse_loss(model(nf, ef, graph).float(), tgt)

    for nf, tgt in data:
        s_nf.copy_(nf)
        s_tgt.copy_(tgt)
        loss = training_step(s_nf, s_tgt)
    torch.cuda.synchronize()
    assert torch.allclose(loss, loss_e.detach(), rtol=1e-3, atol=1e-6)
    for (k, a), b in zip(eager.named_parameters(), model.parameters()):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-5), k
    assert float(opt.state[next(model.parameters())]["step"]) == float(len(data))


@pytest.mark.gpu
def test_static_capture_evaluate_no_grad():
    """Inference caller of the same forward (examples/cfd/vortex_shedding_mgn/inference.py): replays follow the
    static input."""
    from modulus_b200.capture import StaticCaptureEvaluateNoGrad
    from modulus_b200.models.meshgraphnet import MeshGraphNet

    graph, ef, data = _capture_case()
    torch.manual_seed(6)
    model = MeshGraphNet(6, 3, 3, processor_size=3).to(DEV)
    s_nf = torch.empty_like(data[0][0])

    @StaticCaptureEvaluateNoGrad(model=model, cuda_graph_warmup=1)
    def predict(nf):
        return model(nf, ef, graph)

    for nf, _ in data[:4]:
        s_nf.copy_(nf)
        out = predict(s_nf)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            ref = model(nf, ef, graph)
        assert torch.equal(out, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("use_amp", [False, True])
def test_capture_after_eager_steps_on_the_default_stream(use_amp):
    """Eager steps on the default stream followed by a capture on a side stream: the autograd graph of an eager step
    (and with it the AccumulateGrad nodes bound to the default stream) must be gone once its loss is dropped --
    without waiting for the garbage collector -- or the recording fails with cudaErrorStreamCaptureImplicit."""
    import gc

    from modulus_b200.capture import StaticCaptureTraining
    from modulus_b200.models.meshgraphnet import MeshGraphNet
    from modulus_b200.optim import FusedAdam

    graph, ef, data = _capture_case()
    torch.manual_seed(7)
    model = MeshGraphNet(6, 3, 3, processor_size=2).to(DEV)
    opt = FusedAdam(model.parameters(), lr=1e-3)
    nf, tgt = data[0]
    gc.collect()
    gc.disable()
    try:
        for _ in range(2):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_amp):
                loss = torch.nn.functional.mse_loss(model(nf, ef, graph).float(), tgt)
            loss.backward()
            opt.step()
        del loss

        @StaticCaptureTraining(model=model, optim=opt, use_amp=use_amp, cuda_graph_warmup=1)
        def training_step(nf, tgt):
            return torch.nn.functional.mse_loss(model(nf, ef, graph).float(), tgt)

        for _ in range(3):
            out = training_step(nf, tgt)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
    finally:
        gc.enable()


# ------------------------------------------------------------------------------ DevicePrefetcher
def test_prefetcher_has_no_cpu_path():
    from modulus_b200.prefetch import DevicePrefetcher

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DevicePrefetcher("cpu")


@pytest.mark.gpu
def test_prefetcher_stages_batches_in_order_while_compute_runs():
    from modulus_b200.prefetch import DevicePrefetcher

    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(300_000, 7, generator=g).pin_memory(), torch.randn(50_000, 3, generator=g).pin_memory())
               for _ in range(5)]
    pf = DevicePrefetcher(DEV)
    with pytest.raises(RuntimeError, match="nothing staged"):
        pf.take()
    pf.stage(*batches[0])
    sums = []
    ptrs = set()
    for i in range(len(batches)):
        a, b = pf.take()
        if i + 1 < len(batches):
            pf.stage(*batches[i + 1])
        x = a
        for _ in range(20):                       # keep the compute stream busy while the next copy runs
            x = x * 1.0001 + 0.0
        sums.append((a.double().sum() + b.double().sum() + 0 * x.double().sum()))
        ptrs.add(a.data_ptr())
        pf.release()
    torch.cuda.synchronize()
    for s, (ha, hb) in zip(sums, batches):
        assert abs(float(s) - float(ha.double().sum() + hb.double().sum())) < 1e-6 * float(ha.double().abs().sum())
    assert len(ptrs) == 2                         # two static slots, reused
    assert pf.h2d_bytes == sum(a.numel() * 4 + b.numel() * 4 for a, b in batches)
    pf.stage(*batches[0])
    pf.stage(*batches[1])
    with pytest.raises(RuntimeError, match="every slot"):
        pf.stage(*batches[2])
