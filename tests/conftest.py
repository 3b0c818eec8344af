import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gps-gaussian_b200", "dropin")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and return the path of libgpsg_sm100.so; nvcc cross-compiles without a GPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_gpsg_build", os.path.join(ROOT, "gps-gaussian_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()
