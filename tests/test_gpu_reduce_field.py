"""SURVEY 8(f) N3 on the GPU: dg_reduce_field with its index passes in k4_reduce.cu (cell flags, cell map + row compaction, node marks,
Z-curve keys, renumbering, coefficient gather) must leave nodes / cells / cell map exactly as the reference tool and the reference class do --
the same checks as tests/test_reduce_field.py (which forces the host passes), with the default flags."""
import pytest

import test_reduce_field as T

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def gpu_passes(monkeypatch):
    monkeypatch.setattr(T, "HOST", 0)


test_sdf_field_reduced_like_the_reference_tool = T.test_sdf_field_reduced_like_the_reference_tool
test_density_field_reduced_like_the_reference_tool = T.test_density_field_reduced_like_the_reference_tool
test_reduce_field_edge_cases = T.test_reduce_field_edge_cases
test_tied_morton_keys_follow_the_reference_sort = T.test_tied_morton_keys_follow_the_reference_sort
test_against_the_reference_library = T.test_against_the_reference_library


def test_gpu_and_host_passes_agree_on_a_large_field(dg):
    """a 40^3 blob field, ~half of the cells removed: GPU passes == host passes (nodes, cells, cell map), ties included"""
    import numpy as np
    rng = np.random.default_rng(3)
    res = np.array([40, 33, 27], np.uint32)
    mn, mx = np.array([-1.0, -0.7, -0.4]), np.array([1.3, 0.9, 0.8])
    d = dg.grid_desc(mn, mx, res)
    g = {"mn": mn, "mx": mx, "res": res, "cell": np.array(d.cell_size[:]), "inv": np.array(d.inv_cell_size[:])}
    grid = dg.CubicLagrangeDiscreteGrid(mn, mx, res)
    x = grid.nodePositions()
    v = np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1]) + 0.3 * x[:, 2] + 0.05 * rng.standard_normal(len(x))
    from discregrid_b200 import _capi as capi
    import ctypes as C
    cells = np.empty((int(np.prod(res.astype(np.uint64))), 32), np.uint32)
    capi.check(capi.lib.dg_build_cells(d.resolution, 0, len(cells), capi.ptr(cells, capi.U32P)))
    keep = (np.abs(v) < 0.35).astype(np.uint8)
    a = T.reduce_field(dg, g, v, keep, cells, flags=0)
    b = T.reduce_field(dg, g, v, keep, cells, flags=2)
    assert len(a[0]) > 1000 and len(a[1]) < len(cells)
    assert T.bits_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
