"""Multi-GPU through the C-ABI from ONE process (dg_mesh_group_*): dg_add_function_sdf_multi and the NCCL device form reproduce the
single-launch result bit for bit.  n_gpus = 1 exercises the same code (2 interleaved parts on one GPU) on a 1-GPU box; n_gpus = 2..
runs when the box has the devices."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _group(dg, capi, md, n):
    h = C.c_void_p()
    capi.check(capi.lib.dg_mesh_group_create(md.handle, n, None, C.byref(h)))
    assert capi.lib.dg_mesh_group_size(h) == n
    return h


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
@pytest.mark.parametrize("res", [(11, 9, 6), (16, 16, 16)])
def test_add_function_multi_equals_single(dg, orc, torus_small, n_gpus, res):
    from discregrid_b200 import _capi as capi
    if dg.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    md = dg.TriangleMeshDistance(torus_small)
    mn, mx = dg.generate_sdf_domain(torus_small.vertices)
    desc = dg.grid_desc(mn, mx, res)
    n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
    n_cells = res[0] * res[1] * res[2]
    single = np.empty(n); cells1 = np.empty((n_cells, 32), np.uint32); map1 = np.empty(n_cells, np.uint32)
    capi.check(capi.lib.dg_add_function_sdf(md.handle, C.byref(desc), -1.0, capi.ptr(single, capi.F64P), capi.ptr(cells1, capi.U32P), capi.ptr(map1, capi.U32P), None))
    gd, r = orc.grid_desc(mn, mx, res)
    want = orc.mesh(torus_small.vertices, torus_small.faces).sample_sdf(gd, r, sign=-1.0)
    assert np.array_equal(single.view(np.uint64), want.view(np.uint64))
    assert np.array_equal(cells1, orc.build_cells(r)) and np.array_equal(map1, np.arange(n_cells, dtype=np.uint32))
    grp = _group(dg, capi, md, n_gpus)
    try:
        multi = np.full(n, np.nan); cells2 = np.zeros((n_cells, 32), np.uint32); map2 = np.zeros(n_cells, np.uint32); tm = np.zeros(6)
        capi.check(capi.lib.dg_add_function_sdf_multi(grp, C.byref(desc), -1.0, capi.ptr(multi, capi.F64P), capi.ptr(cells2, capi.U32P), capi.ptr(map2, capi.U32P),
                                                      capi.ptr(tm, capi.F64P)))
        assert np.array_equal(multi.view(np.uint64), single.view(np.uint64))
        assert np.array_equal(cells2, cells1) and np.array_equal(map2, map1) and tm[5] == n_gpus
        # tables optional
        multi2 = np.full(n, np.nan)
        capi.check(capi.lib.dg_add_function_sdf_multi(grp, C.byref(desc), -1.0, capi.ptr(multi2, capi.F64P), None, None, None))
        assert np.array_equal(multi2.view(np.uint64), single.view(np.uint64))
    finally:
        capi.lib.dg_mesh_group_destroy(grp)


@pytest.mark.parametrize("n_gpus", [1, 2, 8])
def test_device_form_allgather_equals_single(dg, torus_small, n_gpus):
    """dg_sample_sdf_multi_device: every GPU of the group ends with the full coefficient array (ncclAllGather + unpack)"""
    torch = pytest.importorskip("torch")
    from discregrid_b200 import _capi as capi
    if dg.device_count() < n_gpus or not torch.cuda.is_available():
        pytest.skip(f"needs {n_gpus} GPUs")
    md = dg.TriangleMeshDistance(torus_small)
    mn, mx = dg.generate_sdf_domain(torus_small.vertices)
    desc = dg.grid_desc(mn, mx, (20, 18, 10))
    n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
    single = np.empty(n)
    capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(desc), 1.0, 0, n, capi.ptr(single, capi.F64P)))
    grp = _group(dg, capi, md, n_gpus)
    try:
        outs = [torch.full((n,), float("nan"), dtype=torch.float64, device=f"cuda:{i}") for i in range(n_gpus)]
        ptrs = (C.c_void_p * n_gpus)(*[C.c_void_p(o.data_ptr()) for o in outs])
        for _ in range(2):                                   # second call: communicators and buffers reused
            capi.check(capi.lib.dg_sample_sdf_multi_device(grp, C.byref(desc), 1.0, ptrs))
        for i, o in enumerate(outs):
            torch.cuda.synchronize(i)
            assert np.array_equal(o.cpu().numpy().view(np.uint64), single.view(np.uint64)), f"device {i}"
    finally:
        capi.lib.dg_mesh_group_destroy(grp)


def test_group_argument_errors(dg, torus_small):
    from discregrid_b200 import _capi as capi
    md = dg.TriangleMeshDistance(torus_small)
    h = C.c_void_p()
    assert capi.lib.dg_mesh_group_create(md.handle, 0, None, C.byref(h)) == capi.DG_ERR_INVALID
    assert capi.lib.dg_mesh_group_create(md.handle, 17, None, C.byref(h)) == capi.DG_ERR_INVALID
    assert capi.lib.dg_mesh_group_create(None, 1, None, C.byref(h)) == capi.DG_ERR_INVALID
    assert capi.lib.dg_mesh_group_create(md.handle, dg.device_count() + 1, None, C.byref(h)) == capi.DG_ERR_INVALID
    assert capi.lib.dg_mesh_group_destroy(None) == capi.DG_OK


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_large_grid_uses_more_parts_per_gpu(dg, torus_small, n_gpus):
    """12.4 M nodes: dg_add_function_sdf_multi deals more than two parts to a GPU (three with one GPU) -- same coefficients as the single-GPU call"""
    from discregrid_b200 import _capi as capi
    if dg.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    md = dg.TriangleMeshDistance(torus_small)
    mn, mx = dg.generate_sdf_domain(torus_small.vertices)
    desc = dg.grid_desc(mn, mx, (120, 120, 120))
    n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
    single = np.empty(n)
    capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(desc), 1.0, 0, n, capi.ptr(single, capi.F64P)))
    grp = _group(dg, capi, md, n_gpus)
    try:
        multi = np.full(n, np.nan)
        capi.check(capi.lib.dg_add_function_sdf_multi(grp, C.byref(desc), 1.0, capi.ptr(multi, capi.F64P), None, None, None))
        assert np.array_equal(multi.view(np.uint64), single.view(np.uint64))
    finally:
        capi.lib.dg_mesh_group_destroy(grp)
