import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    try:
        from fingering_dynamics_b200 import _native as nat
        return int(nat.lib().fdlbm_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """plain `pytest` on a machine without a CUDA device: gpu-marked tests are SKIPPED (not failed), so CPU-side
    regressions stay visible.  With `-m gpu` nothing is skipped: on a GPU box a missing device or library must fail
    loudly (tests/test_host.py::test_no_cpu_fallback_without_a_device is the explicit contract test)."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device here (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load
