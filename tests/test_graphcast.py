"""GraphCast blocks on the MeshGraphNet operator seam: MeshGraphEncoder,
MeshGraphDecoder (bipartite, N_src != N_dst, tuple form of concat_efeat / sum_efeat) and
GraphCastProcessor, against goldens of the unmodified reference
(tests/golden/make_golden_graphcast.py).  CPU part: constructor / state_dict compatibility;
GPU part: outputs and ALL gradients through the C ABI, fp32 1e-4 relative, bf16 2e-2."""
import pytest
import torch

from conftest import load_golden

DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _build(name, case):
    from modulus_b200.models.gnn_layers import MeshGraphDecoder, MeshGraphEncoder
    from modulus_b200.models.graphcast import GraphCastProcessor

    cls = {"encoder": MeshGraphEncoder, "decoder": MeshGraphDecoder, "processor": GraphCastProcessor}[name.split("_")[0]]
    model = cls(**case["kwargs"])
    model.load_state_dict(case["state_dict"], strict=True)
    return model


CASES = [f"{b}_{m}_{a}" for b in ("encoder", "decoder") for m in ("concat", "trick") for a in ("sum", "mean")] + \
        ["processor_concat", "processor_trick"]


@pytest.mark.parametrize("name", CASES)
def test_state_dict_layout_matches_reference(name):
    g = load_golden("ref_graphcast_blocks.pt")
    model = _build(name, g[name])
    assert list(model.state_dict().keys()) == list(g[name]["state_dict"].keys())
    for k, v in model.state_dict().items():
        assert v.shape == g[name]["state_dict"][k].shape, k


def test_same_rng_stream_as_reference():
    """Constructing under the same seed consumes the RNG like the reference: identical weights."""
    from modulus_b200.models.gnn_layers import MeshGraphEncoder

    g = load_golden("ref_graphcast_blocks.pt")
    for trick in (False, True):
        case = g[f"encoder_{'trick' if trick else 'concat'}_sum"]
        torch.manual_seed(31 + trick)
        model = MeshGraphEncoder(**case["kwargs"])
        for k, v in model.state_dict().items():
            assert torch.equal(v, case["state_dict"][k]), k


def test_processor_checkpoint_segments_error():
    from modulus_b200.models.graphcast import GraphCastProcessor

    p = GraphCastProcessor(processor_layers=3, input_dim_nodes=8, input_dim_edges=8, hidden_dim=8)
    with pytest.raises(ValueError):
        p.set_checkpoint_segments(4)
    p.set_checkpoint_segments(2)
    assert p.checkpoint_segments == [(0, 3), (3, 6)]
    p.set_checkpoint_segments(-1)
    assert p.checkpoint_segments == [(0, 6)]


def _graph(g, name):
    from modulus_b200.models.gnn_layers import CuGraphCSC

    if name.startswith("processor"):
        s = g["square_graph"]
        return CuGraphCSC(s["offsets"].to(DEV), s["indices"].to(DEV), s["n"], s["n"])
    b = g["bipartite_graph"]
    return CuGraphCSC(b["offsets"].to(DEV), b["indices"].to(DEV), b["n_src"], b["n_dst"])


def _loss(outs):
    return sum((o.float() * torch.linspace(-1, 1, o.numel(), device=o.device).view_as(o)).sum() for o in outs)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_blocks_fp32_match_reference(name):
    g = load_golden("ref_graphcast_blocks.pt")
    case = g[name]
    model = _build(name, case).to(DEV)
    graph = _graph(g, name)
    inputs = [x.to(DEV).requires_grad_(True) for x in case["inputs"]]
    outs = model(*inputs, graph)
    outs = list(outs) if isinstance(outs, tuple) else [outs]
    _loss(outs).backward()
    for o, ref in zip(outs, case["outputs"]):
        assert rel_err(o, ref) < 1e-4
    for x, ref in zip(inputs, case["input_grads"]):
        assert rel_err(x.grad, ref) < 1e-4
    for k, p in model.named_parameters():
        assert rel_err(p.grad, case["param_grads"][k]) < 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["encoder_concat_sum", "decoder_trick_mean", "processor_concat"])
def test_blocks_bf16_within_2e2(name):
    g = load_golden("ref_graphcast_blocks.pt")
    case = g[name]
    model = _build(name, case).to(DEV)
    graph = _graph(g, name)
    inputs = [x.to(DEV) for x in case["inputs"]]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(*inputs, graph)
    outs = list(outs) if isinstance(outs, tuple) else [outs]
    for o, ref in zip(outs, case["outputs"]):
        a, b = o.detach().float().cpu(), ref
        assert float((a - b).norm() / b.norm()) < 2e-2


@pytest.mark.gpu
def test_processor_checkpointing_same_result():
    g = load_golden("ref_graphcast_blocks.pt")
    case = g["processor_concat"]
    model = _build("processor_concat", case).to(DEV)
    graph = _graph(g, "processor_concat")
    res = []
    for seg in (-1, 3):
        model.set_checkpoint_segments(seg)
        model.zero_grad()
        inputs = [x.to(DEV).requires_grad_(True) for x in case["inputs"]]
        ef, nf = model(*inputs, graph)
        _loss([ef, nf]).backward()
        res.append((ef.detach(), nf.detach(), inputs[0].grad, inputs[1].grad))
    for a, b in zip(*res):
        assert torch.equal(a, b)
