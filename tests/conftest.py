import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    """Make sure the CPU checkers exist (the oracle always; oracle/_ref only
    where /root/reference is available or a prebuilt copy travelled along)."""
    import oracle
    if not oracle.have_oracle():
        oracle.build("oracle")
    if not oracle.have_ref() and os.path.isdir("/root/reference/src"):
        oracle.build("ref")
    yield


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# The TMA-staged kernels (k_ct_tma, k_edge_efield_tma) only take blocks whose
# tiles fill the chip (about 128^3 cells and up). The small blocks of the
# pass-structure, parity and domain tests therefore run twice on the GPU: with
# the library's defaults, and with the TMA-staged kernels forced on (option
# "pair_kernels" bit 5, through the environment variable a new handle reads).
TMA_FORCED_MODULES = {"test_gpu_parts", "test_gpu_fused_timestep", "test_gpu_domain",
                      "test_gpu_parity"}


@pytest.fixture(autouse=True)
def vlct_kernel_variant(request):
    variant = getattr(request, "param", "default")
    old = os.environ.get("VLCT_PAIR_MASK")
    if variant == "tma":
        os.environ["VLCT_PAIR_MASK"] = "62"
    yield variant
    if variant == "tma":
        if old is None:
            os.environ.pop("VLCT_PAIR_MASK", None)
        else:
            os.environ["VLCT_PAIR_MASK"] = old


def pytest_generate_tests(metafunc):
    if ("vlct_kernel_variant" in metafunc.fixturenames and
            metafunc.definition.get_closest_marker("gpu") is not None and
            metafunc.module.__name__ in TMA_FORCED_MODULES):
        metafunc.parametrize("vlct_kernel_variant", ["default", "tma"], indirect=True)
