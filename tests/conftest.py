import os
import sys

import pytest

os.environ.setdefault("BACON_IVP_SENTINEL", "1")  # status arrays start at -1: a trajectory the kernels lost would show

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a runaway kernel (a trajectory that never retires) must fail one test, not eat the GPU call:
    # the "thread" method ends the process even while a C call is blocking
    for it in items:
        if "gpu" in it.keywords and not any(m.name == "timeout" for m in it.iter_markers()):
            it.add_marker(pytest.mark.timeout(300, method="thread"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so), compiled on demand. Test infrastructure only."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engine():
    """The product: bacon_b200 over libbacon_ivp.so. Fails loudly if the library is missing."""
    import bacon_b200
    from bacon_b200._lib import lib
    lib()
    return bacon_b200


@pytest.fixture(scope="session")
def cuda():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    return torch
