import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def R():
    """The product package; builds librtb.so if it is missing (nvcc cross-compiles without a GPU)."""
    from rust_raytrace_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(os.path.join(os.path.dirname(_lib.LIB_PATH), "raytrace_b200")):
        _lib.build()
    import rust_raytrace_b200

    return rust_raytrace_b200


@pytest.fixture(scope="session")
def teapot_mesh(R):
    return R.obj_parser.load_mesh_bin()


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    g = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(g, "main_scene_stats.json")) as fh:
        stats = json.load(fh)
    return stats, np.load(os.path.join(g, "main_scene_64.npz"))
