import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Several tests exchange a few hundred MB with helper binaries through tmp_path.  In the build container /tmp sits on a copy-on-write
# layer that writes at ~15-75 MB/s while the repo's own volume does > 1 GB/s, so the temporary root goes under build/ (git-ignored).
_TMPROOT = os.path.join(ROOT, "build", "pytest-tmp")
try:
    os.makedirs(_TMPROOT, exist_ok=True)
    os.environ.setdefault("PYTEST_DEBUG_TEMPROOT", _TMPROOT)
except OSError:
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle_api import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def dg():
    import discregrid_b200
    return discregrid_b200


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def box_mesh(dg):
    return dg.TriangleMesh(os.path.join(GOLDEN, "box.obj"))


@pytest.fixture(scope="session")
def torus_small(dg):
    return dg.bumpy_torus(48, 48, 1.0, 0.4, 0.05, 7, 5)      # the mesh of tests/golden/ref_torus_*.npz


def ref_resource(name):
    """Reference mesh staged by `make -C oracle ref` (travels to the GPU box via oracle/_ref); None if absent."""
    p = os.path.join(ROOT, "oracle", "_ref", "resources", name)
    return p if os.path.exists(p) else None


def grid_for(orc, V, res):
    """(oracle grid arrays, dg_grid_desc kwargs) for the GenerateSDF-padded domain of vertices V."""
    mn, mx = orc.generate_sdf_domain(V)
    gd, r = orc.grid_desc(mn, mx, res)
    return mn, mx, gd, r


def bits_equal(a, b):
    """bit-exact comparison of float arrays (distinguishes -0.0 / NaN payloads)."""
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
