"""K2 parity (GPU): batched interpolate / gradient through the C-ABI vs the CPU oracle.  Bar: bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal

pytestmark = pytest.mark.gpu
DBL_MAX = np.finfo(np.float64).max


def _desc_arrays(orc, g):
    return orc.grid_desc(g.m_domain[0], g.m_domain[1], g.m_resolution, g.m_cell_size, g.m_inv_cell_size)


def _queries(g, n, seed, pad=0.02):
    rng = np.random.default_rng(seed)
    lo, hi = g.m_domain
    span = hi - lo
    return lo - pad * span + rng.random((n, 3)) * (1 + 2 * pad) * span       # a few percent fall outside the domain


def test_shape_functions_bit_exact(dg, orc):
    import ctypes as C
    from discregrid_b200 import _capi as capi
    rng = np.random.default_rng(3)
    xi = np.concatenate([rng.uniform(-1, 1, (5000, 3)), np.array([[0, 0, 0], [1, 1, 1], [-1, -1, -1], [1, -1, 1 / 3]])])
    N = np.empty((len(xi), 32)); dN = np.empty((len(xi), 32, 3))
    capi.check(capi.lib.dg_shape_functions(capi.ptr(xi, capi.F64P), len(xi), capi.ptr(N, capi.F64P), capi.ptr(dN, capi.F64P)))
    No, dNo = orc.shape_functions(xi)
    assert bits_equal(N, No) and bits_equal(dN, dNo)


def test_box_cdf_million_queries(dg, orc):
    g = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    x = _queries(g, 1_000_000, 0x5EED)
    gd, res = _desc_arrays(orc, g)
    phi, grad = g.interpolate(0, x, gradient=True)
    po, go = orc.interpolate(gd, res, g.m_nodes[0], x, grad=True, cells=g.m_cells[0], cell_map=g.m_cell_map[0])
    assert bits_equal(phi, po) and bits_equal(grad, go)
    assert (phi == DBL_MAX).sum() > 1000                      # the out-of-domain sentinel path was exercised
    pv = g.interpolate(x)                                     # value-only overload, field 0
    assert bits_equal(pv, orc.interpolate(gd, res, g.m_nodes[0], x, grad=False)[0])


def test_nodal_property_and_boundaries(dg, orc):
    g = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    xn = g.nodePositions()
    phi = g.interpolate(0, xn)
    inside = phi != DBL_MAX
    assert inside.mean() > 0.99
    assert np.max(np.abs(phi[inside] - g.m_nodes[0][inside])) < 1e-14
    lo, hi = g.m_domain
    x = np.array([lo, hi, [hi[0], lo[1], hi[2]], np.nextafter(hi, np.inf), np.nextafter(lo, -np.inf), [np.nan, 0, 0]])
    gd, res = _desc_arrays(orc, g)
    phi, grad = g.interpolate(0, x, gradient=True)
    po, go = orc.interpolate(gd, res, g.m_nodes[0], x, grad=True)
    assert bits_equal(phi, po) and bits_equal(grad, go)
    assert phi[3] == DBL_MAX and phi[4] == DBL_MAX and phi[5] == DBL_MAX and not grad[3:].any()


@pytest.mark.parametrize("res", [(9, 4, 6), (1, 1, 1), (33, 2, 17)])
def test_random_field_anisotropic(dg, orc, res):
    rng = np.random.default_rng(11)
    g = dg.CubicLagrangeDiscreteGrid([-0.3, 1.0, -2.5], [0.9, 1.7, 0.25], res)
    g.addSampledFunction(rng.normal(size=g.nNodes()))
    x = _queries(g, 200_003, 5)                                # not a multiple of 32: tail warp
    gd, r = _desc_arrays(orc, g)
    phi, grad = g.interpolate(0, x, gradient=True)
    po, go = orc.interpolate(gd, r, g.m_nodes[0], x, grad=True)
    assert bits_equal(phi, po) and bits_equal(grad, go)


def test_reduced_field_and_sentinels(dg, orc):
    """a field as reduceField leaves it: cells removed (cell_map == UINT_MAX), nodes renumbered, plus DBL_MAX coefficients"""
    rng = np.random.default_rng(2)
    g = dg.CubicLagrangeDiscreteGrid([0, 0, 0], [1, 2, 1.5], (6, 5, 4))
    vals = rng.normal(size=g.nNodes())
    vals[rng.random(len(vals)) < 0.01] = DBL_MAX               # missing coefficients (addFunction with a predicate)
    g.addSampledFunction(vals)
    keep = rng.random(g.nCells()) < 0.6
    cells = g.m_cells[0][keep]
    used, inv = np.unique(cells, return_inverse=True)
    perm = rng.permutation(len(used))                           # arbitrary renumbering of the surviving nodes
    g.m_nodes[0] = vals[used][np.argsort(perm)]
    g.m_cells[0] = perm[inv.reshape(cells.shape)].astype(np.uint32)
    cmap = np.full(g.nCells(), 0xFFFFFFFF, np.uint32); cmap[keep] = np.arange(keep.sum(), dtype=np.uint32)
    g.m_cell_map[0] = cmap
    g._invalidate()
    x = _queries(g, 100_000, 9)
    gd, r = _desc_arrays(orc, g)
    phi, grad = g.interpolate(0, x, gradient=True)
    po, go = orc.interpolate(gd, r, g.m_nodes[0], x, grad=True, cells=g.m_cells[0], cell_map=g.m_cell_map[0])
    assert bits_equal(phi, po) and bits_equal(grad, go)
    assert 0.2 < (phi == DBL_MAX).mean() < 0.8


def test_sdf_field_round_trip(dg, orc, torus_small):
    """addFunction (K1) -> interpolate (K2) on the same grid, vs the oracle end to end"""
    mn, mx = dg.generate_sdf_domain(torus_small.vertices)
    g = dg.CubicLagrangeDiscreteGrid(mn, mx, (20, 20, 10))
    g.addFunction(dg.MeshSignedDistance(dg.TriangleMeshDistance(torus_small)))
    x = _queries(g, 50_000, 4, pad=0.0)
    gd, r = _desc_arrays(orc, g)
    phi, grad = g.interpolate(0, x, gradient=True)
    po, go = orc.interpolate(gd, r, g.m_nodes[0], x, grad=True)
    assert bits_equal(phi, po) and bits_equal(grad, go)
    # the interpolant approximates the true distance (sanity of the whole pipeline, loose)
    true = dg.TriangleMeshDistance(torus_small).signed_distance(x).distance
    assert np.median(np.abs(phi - true)) < 5e-3


def test_host_pipeline_chunks_pinned_and_pageable(dg, orc):
    """dg_interpolate_batch's three-slot pipeline: 2.3 M queries = five 512k-query chunks, so every slot is reused; pageable caller buffers
    (staged through the pinned pool), page-locked ones (direct DMA), value-only and ragged sizes all equal the single-launch device result"""
    import ctypes as C
    from discregrid_b200 import _capi as capi
    torch = pytest.importorskip("torch")
    g = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    n = 2_300_017
    x = _queries(g, n, 77)
    fh = g._device_field(0)
    xd = torch.from_numpy(x).cuda(); pd = torch.empty(n, dtype=torch.float64, device="cuda"); gd_ = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    capi.check(capi.lib.dg_interpolate_batch_device(fh, C.c_void_p(xd.data_ptr()), n, C.c_void_p(pd.data_ptr()), C.c_void_p(gd_.data_ptr()), None))
    torch.cuda.synchronize()
    want_p, want_g = pd.cpu().numpy(), gd_.cpu().numpy()
    # pageable
    p1 = np.full(n, np.nan); g1 = np.full((n, 3), np.nan)
    capi.check(capi.lib.dg_interpolate_batch(fh, capi.ptr(x, capi.F64P), n, capi.ptr(p1, capi.F64P), capi.ptr(g1, capi.F64P)))
    assert bits_equal(p1, want_p) and bits_equal(g1, want_g)
    # page-locked
    xp = torch.from_numpy(x).pin_memory(); pp = torch.empty(n, dtype=torch.float64).pin_memory(); gp = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    capi.check(capi.lib.dg_interpolate_batch(fh, C.cast(xp.data_ptr(), capi.F64P), n, C.cast(pp.data_ptr(), capi.F64P), C.cast(gp.data_ptr(), capi.F64P)))
    assert bits_equal(pp.numpy(), want_p) and bits_equal(gp.numpy(), want_g)
    # value only, a size that is not a multiple of anything, mixed: pinned in, pageable out
    m = 1_048_577
    p2 = np.full(m, np.nan)
    capi.check(capi.lib.dg_interpolate_batch(fh, C.cast(xp.data_ptr(), capi.F64P), m, capi.ptr(p2, capi.F64P), None))
    assert bits_equal(p2, want_p[:m])
    # a single query still works after the pool has grown
    p3 = np.empty(1); g3 = np.empty((1, 3))
    capi.check(capi.lib.dg_interpolate_batch(fh, capi.ptr(x[5:6].copy(), capi.F64P), 1, capi.ptr(p3, capi.F64P), capi.ptr(g3, capi.F64P)))
    assert bits_equal(p3, want_p[5:6]) and bits_equal(g3, want_g[5:6])
