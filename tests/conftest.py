import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "meg-decoding_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Builds libmegb200.so if it is missing (nvcc cross-compiles without a GPU)."""
    from megb200 import lib
    if not os.path.exists(lib.LIB_PATH):
        import subprocess
        subprocess.check_call(["bash", os.path.join(PKG, "csrc", "build.sh")])
    return lib.load()
