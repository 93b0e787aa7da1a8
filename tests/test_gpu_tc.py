"""Parity of the fused tcgen05 kernels (bf16) against an fp32 CPU restatement that applies the
same roundings (bf16 inputs / weights / hidden activations, fp32 accumulate)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bf(x):
    return x.to(torch.bfloat16).float()


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def ref_mlp(A, w1, b1, w2, b2, w3, b3, gamma, beta, residual):
    """fp32 math on bf16-rounded operands, hidden activations rounded to bf16 like the kernel."""
    h1 = bf(F.relu(A @ bf(w1).T + b1))
    h2 = bf(F.relu(h1 @ bf(w2).T + b2))
    y = h2 @ bf(w3).T + b3
    if gamma is not None:
        y = F.layer_norm(y, (y.shape[1],), gamma, beta, 1e-5)
    if residual is not None:
        y = y + residual
    return y, h1, h2


def make_params(k1, n_out=128, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(w1=r(128, k1) / k1 ** 0.5, b1=r(128) * 0.1, w2=r(128, 128) / 128 ** 0.5, b2=r(128) * 0.1,
                w3=r(n_out, 128) / 128 ** 0.5, b3=r(n_out) * 0.1, gamma=1 + 0.1 * r(128), beta=0.1 * r(128))


def dev_params(p):
    return {k: v.to(DEV).contiguous() for k, v in p.items()}


@pytest.mark.parametrize("M", [1, 100, 128, 1000, 70001])
def test_edge_block_fwd_tc(M):
    from modulus_b200 import ops

    g = torch.Generator().manual_seed(M)
    N = max(M // 5, 3)
    e = bf(torch.randn(M, 128, generator=g))
    n = bf(torch.randn(N, 128, generator=g))
    src = torch.randint(0, N, (M,), generator=g)
    dst = torch.randint(0, N, (M,), generator=g)
    p = make_params(384)
    A = torch.cat([e, n[src], n[dst]], 1)
    ref, h1, h2 = ref_mlp(A, p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["gamma"], p["beta"], e)
    d = dev_params(p)
    eb, nb = e.to(DEV).bfloat16(), n.to(DEV).bfloat16()
    out, s1, s2 = ops.mlp3_fwd_tc([eb, nb, nb], [None, src.to(DEV).int(), dst.to(DEV).int()], M,
                                  d["w1"], d["b1"], d["w2"], d["b2"], d["w3"], d["b3"], d["gamma"], d["beta"],
                                  residual=eb, save_hidden=True)
    torch.cuda.synchronize()
    ops.tc_check(DEV)
    assert rel_err(s1, h1) < 1e-2 and rel_err(s2, h2) < 1e-2
    assert rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("M", [77, 30000])
def test_node_and_plain_and_decoder_fwd_tc(M):
    from modulus_b200 import ops

    g = torch.Generator().manual_seed(M + 1)
    a = bf(torch.randn(M, 128, generator=g))
    n = bf(torch.randn(M, 128, generator=g))
    # node block: tables (agg, nfeat), K1 = 256, residual = nfeat
    p = make_params(256, seed=1)
    ref, _, _ = ref_mlp(torch.cat([a, n], 1), p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["gamma"],
                        p["beta"], n)
    d = dev_params(p)
    ab, nb = a.to(DEV).bfloat16(), n.to(DEV).bfloat16()
    out, _, _ = ops.mlp3_fwd_tc([ab, nb], [None, None], M, d["w1"], d["b1"], d["w2"], d["b2"], d["w3"], d["b3"],
                                d["gamma"], d["beta"], residual=nb)
    assert rel_err(out, ref) < 1e-2
    # decoder: one table, no LayerNorm, 3 outputs
    p = make_params(128, n_out=3, seed=2)
    ref, _, _ = ref_mlp(n, p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], None, None, None)
    d = dev_params(p)
    out, _, _ = ops.mlp3_fwd_tc([nb], [None], M, d["w1"], d["b1"], d["w2"], d["b2"], d["w3"], d["b3"], n_out=3)
    assert out.shape == (M, 3) and rel_err(out, ref) < 1e-2
    # encoder: raw fp32 features with 6 columns
    x = torch.randn(M, 6, generator=g)
    p = make_params(6, seed=3)
    ref, _, _ = ref_mlp(bf(x), p["w1"], p["b1"], p["w2"], p["b2"], p["w3"], p["b3"], p["gamma"], p["beta"], None)
    d = dev_params(p)
    out, _, _ = ops.mlp3_fwd_tc([], [], M, d["w1"], d["b1"], d["w2"], d["b2"], d["w3"], d["b3"], d["gamma"],
                                d["beta"], small_x=x.to(DEV))
    torch.cuda.synchronize()
    ops.tc_check(DEV)
    assert rel_err(out, ref) < 1e-2
