import os
import sys

import pytest

# LocalSlabGroup puts several slab engines (2 streams each, kernels that spin on their neighbours' flags) on ONE device:
# more hardware queues than the default 8 keep their streams from aliasing (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    from oracle import oracle as O
    O.build()
