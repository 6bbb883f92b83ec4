import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "suffix-array-searching_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def sst():
    """The product package. On a GPU box the CUDA library must be present and usable: fail loudly."""
    import sst_b200

    sst_b200.lib()
    return sst_b200


@pytest.fixture(scope="session")
def gpu(sst):
    if sst.device_count() < 1:
        pytest.fail("GPU test selected but no sm_100 device is usable (no CPU fallback exists)")
    return sst


@pytest.fixture(autouse=True)
def _restore_library_options():
    """Tests change library options with sst.set_option(); every test starts from the load-time values."""
    yield
    mod = sys.modules.get("sst_b200")
    if mod is not None and getattr(mod, "_lib", None) is not None:
        mod.reset_options()
