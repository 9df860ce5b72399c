import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENE_DIR = os.environ.get("MTB_SCENE_DIR", "/tmp/mtb_scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def scene_dir():
    os.makedirs(SCENE_DIR, exist_ok=True)
    return SCENE_DIR


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def product_lib():
    """Builds (if stale) and loads the product library; GPU tests must run the CUDA path or fail."""
    from mythtracer_b200 import build as mtb_build
    mtb_build.build()
    from mythtracer_b200 import api
    return api.load_library()
