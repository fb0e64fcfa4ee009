import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own sources compiled for CPU (oracle/_ref); only where it was built."""
    from oracle.oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def rtc():
    import rtc_b200
    rtc_b200.load_library()
    return rtc_b200


@pytest.fixture(scope="session")
def ctx(rtc):
    """A GPU context; fails loudly (no CPU fallback) if the CUDA library cannot run."""
    c = rtc.Context(0)
    yield c
    c.close()
