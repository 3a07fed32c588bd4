import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build everything once per session (no-ops when up to date; on the GPU box, where the
    reference sources are absent, only the prebuilt files are checked)."""
    from meep_b200 import build
    build.build_all()
    return True


def have_gpu():
    try:
        from meep_b200 import capi
        return capi.load().mb200_device_count() > 0
    except Exception:
        return False
