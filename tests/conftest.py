import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """Build (or reuse) libbocf_b200.so; nvcc cross-compiles sm_100a without a GPU."""
    import __graft_entry__ as ge
    return ge.build_library()


# Every GPU parity test runs twice: with the fp64 DMMA contractions and with the library default ("auto": tcgen05
# int8 digit planes, falling back to fp64 for ill-conditioned factors).  tests.helpers.tol() widens the fp64-tight
# tolerances to the north-star bars in the second mode.
@pytest.fixture(params=["fp64", "auto"])
def cuda_device(request, built_library):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    old = os.environ.get("BOCF_PRECISION")
    os.environ["BOCF_PRECISION"] = request.param
    yield "cuda:0"
    if old is None:
        os.environ.pop("BOCF_PRECISION", None)
    else:
        os.environ["BOCF_PRECISION"] = old
