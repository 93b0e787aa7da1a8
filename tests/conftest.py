import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.fixture(autouse=True)
def _stall_flag_is_clear_after_every_gpu_test(request):
    """The fused kernels report a stalled pipeline (bounded mbarrier waits) through a device flag instead of hanging; the
    optimizer skips its update while the flag is set.  A GPU test that leaves it set has hit a real stall: fail THAT test,
    and do not let the flag leak into the tests that follow."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    import torch

    if not torch.cuda.is_available():
        return
    from modulus_b200 import ops

    for st in list(ops._STATUS.values()):  # one flag per device the tests touched
        code = int(st.item())
        if code != 0:
            st.zero_()
            pytest.fail(f"a fused kernel reported a stall during this test (status {code})")
