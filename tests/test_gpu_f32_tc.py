"""fp32 Linear products on the tensor cores (3 x TF32 split, include/mgn_b200.h: mgn_linear_f32_tc) against float64:
the bar is the one the exact-fp32 SIMT kernel meets on the same inputs (the reference's fp32 default, cuBLAS SGEMM with
TF32 off: test/models/meshgraphnet/test_meshgraphnet_snmg.py:31).  Through the C ABI."""
import pytest
import torch

DEV = "cuda"


def _case(M, K, N, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(M, K, generator=g) * 2 - 1).to(DEV)
    w = ((torch.rand(N, K, generator=g) * 2 - 1) * 0.1).to(DEV)
    b = (torch.rand(N, generator=g) * 2 - 1).to(DEV)
    return x, w, b


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,N", [(1000, 384, 128), (5, 32, 128), (128, 128, 256), (20011, 256, 128)])
def test_forward_matches_float64_as_closely_as_fp32_does(M, K, N):
    from modulus_b200 import ops
    from modulus_b200._lib import ACT_IDS

    x, w, b = _case(M, K, N, 3)
    out = ops.linear_f32_tc(x, w, b, ACT_IDS["relu"])
    ops.tc_check(DEV)
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    scale = (x.double().abs() @ w.double().abs().t()).max()      # sum |a b|: the error scale of a dot product
    assert out.shape == (M, N) and out.dtype == torch.float32
    assert float((out.double() - ref).abs().max()) < 1e-6 * float(scale)   # ~2^-20; one TF32 product alone: ~2^-12
    pre = ops.linear_f32_tc(x, w, None, ACT_IDS[None])
    assert float((pre.double() - x.double() @ w.double().t()).abs().max()) < 1e-6 * float(scale)


@pytest.mark.gpu
def test_data_gradient_through_the_transposed_image():
    from modulus_b200 import ops

    M, K, N = 3001, 384, 128
    _, w, _ = _case(M, K, N, 4)
    g_y = (torch.rand(M, N, generator=torch.Generator().manual_seed(5)) * 2 - 1).to(DEV)
    g_x = ops.linear_f32_tc(g_y, w, None, 0, transpose_w=True)     # g_x = g_y W: N = 384 in three 128-column launches
    ops.tc_check(DEV)
    ref = g_y.double() @ w.double()
    scale = (g_y.double().abs() @ w.double().abs()).max()
    assert g_x.shape == (M, K)
    assert float((g_x.double() - ref).abs().max()) < 1e-6 * float(scale)


@pytest.mark.gpu
def test_shapes_the_kernel_does_not_cover_are_refused():
    from modulus_b200 import ops

    x, w, b = _case(64, 48, 128, 6)                                # K % 32 != 0
    with pytest.raises(ValueError):
        ops.linear_f32_tc(x, w, b)
    x, w, b = _case(64, 64, 96, 7)                                 # N % 128 != 0
    with pytest.raises(ValueError):
        ops.linear_f32_tc(x, w, b)
    with pytest.raises(TypeError):
        ops.linear_f32_tc(x.bfloat16(), w, b)
