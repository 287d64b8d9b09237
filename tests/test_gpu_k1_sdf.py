"""K1 parity (GPU): mesh BVH construction, TriangleMeshDistance queries and the addFunction node loop, through the
C-ABI, against the committed golden vectors (reference-generated) and the CPU oracle.  Bar: bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal, grid_for, ref_resource

pytestmark = pytest.mark.gpu


def test_selftest_and_device(dg):
    assert dg.device_count() >= 1
    dg.selftest()


def test_box_cdf_golden_bit_exact(dg, box_mesh):
    """GenerateSDF -r "5 5 5" box.obj == cmd/generate_sdf/resources/box.cdf: domain, cell sizes, all 1296 coefficients,
    125x32 connectivity, identity cell map -- and the saved file is byte-identical."""
    gold = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    mn, mx = dg.generate_sdf_domain(box_mesh.vertices)
    assert bits_equal(mn, gold.m_domain[0]) and bits_equal(mx, gold.m_domain[1])
    md = dg.TriangleMeshDistance(box_mesh)
    sdf = dg.CubicLagrangeDiscreteGrid(mn, mx, [5, 5, 5])
    assert bits_equal(sdf.m_cell_size, gold.m_cell_size) and bits_equal(sdf.m_inv_cell_size, gold.m_inv_cell_size)
    fid = sdf.addFunction(dg.MeshSignedDistance(md), True)
    assert fid == 0 and sdf.nFields() == 1
    assert bits_equal(sdf.m_nodes[0], gold.m_nodes[0])
    assert np.array_equal(sdf.m_cells[0], gold.m_cells[0])
    assert np.array_equal(sdf.m_cell_map[0], gold.m_cell_map[0])
    out = "/tmp/_dg_box_test.cdf"
    sdf.save(out)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, "box.cdf"), "rb").read()


def test_tree_and_pseudonormals_match_reference_golden(dg, torus_small):
    g = np.load(os.path.join(GOLDEN, "ref_torus_tree.npz"))
    md = dg.TriangleMeshDistance(torus_small)
    sph, kids = md.tree()
    assert np.array_equal(kids, g["kids"])
    assert bits_equal(sph[kids[:, 0] != -1], g["spheres_internal"])
    pt, pe, pv = md.pseudonormals()
    assert bits_equal(pt, g["pn_tri"]) and bits_equal(pe, g["pn_edge"]) and bits_equal(pv, g["pn_vert"])


def test_queries_match_reference_golden(dg, torus_small):
    g = np.load(os.path.join(GOLDEN, "ref_torus_queries.npz"))
    md = dg.TriangleMeshDistance(torus_small)
    r = md.signed_distance(g["x"])
    assert bits_equal(r.distance, g["distance"])
    assert bits_equal(r.nearest_point, g["nearest"])
    assert np.array_equal(r.nearest_entity, g["entity"]) and np.array_equal(r.triangle_id, g["triangle"])
    assert bits_equal(md.unsigned_distance(g["x"]).distance, g["unsigned"])


def test_surface_points_match_reference_golden(dg):
    """points exactly on vertices / edges / faces and 1e-9 off the surface: ties and sign at zero distance"""
    g = np.load(os.path.join(GOLDEN, "ref_sphere_surface.npz"))
    a = g["sphere_args"]
    s = dg.uv_sphere(int(a[0]), int(a[1]), a[2], tuple(a[3:6]))
    md = dg.TriangleMeshDistance(s)
    r = md.signed_distance(g["x"])
    assert bits_equal(r.distance, g["distance"]) and bits_equal(r.nearest_point, g["nearest"])
    assert np.array_equal(r.nearest_entity, g["entity"]) and np.array_equal(r.triangle_id, g["triangle"])


@pytest.mark.parametrize("res", [(16, 16, 16), (7, 3, 5)])
def test_box_grid_vs_oracle_all_nodes(dg, orc, box_mesh, res):
    """config 1 of BASELINE.json: box.obj, 16^3 (plus an anisotropic grid for the index algebra)"""
    mn, mx, gd, r = grid_for(orc, box_mesh.vertices, res)
    want = orc.mesh(box_mesh.vertices, box_mesh.faces).sample_sdf(gd, r)
    sdf = dg.CubicLagrangeDiscreteGrid(mn, mx, res)
    assert bits_equal(sdf.nodePositions(), orc.node_positions(gd, r, 0, len(want)))
    sdf.addFunction(dg.MeshSignedDistance(dg.TriangleMeshDistance(box_mesh)))
    assert bits_equal(sdf.m_nodes[0], want)
    assert np.array_equal(sdf.m_cells[0], orc.build_cells(r))


def test_invert_and_subranges(dg, orc, torus_small):
    import ctypes as C
    from discregrid_b200 import _capi as capi
    res = (12, 10, 9)
    mn, mx, gd, r = grid_for(orc, torus_small.vertices, res)
    om = orc.mesh(torus_small.vertices, torus_small.faces)
    want = om.sample_sdf(gd, r, sign=-1.0)
    md = dg.TriangleMeshDistance(torus_small)
    sdf = dg.CubicLagrangeDiscreteGrid(mn, mx, res)
    sdf.addFunction(dg.MeshSignedDistance(md, invert=True))
    assert bits_equal(sdf.m_nodes[0], want)
    # a rank's shard: arbitrary contiguous node ranges concatenate to the full result
    n = len(want)
    cuts = [0, 1, 777, n // 3, n // 3, n - 5, n]
    parts = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        out = np.empty(e - b)
        capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(sdf._desc), -1.0, b, e, capi.ptr(out, capi.F64P)))
        parts.append(out)
    assert bits_equal(np.concatenate(parts), want)


def test_torus_grid_vs_oracle(dg, orc, torus_small):
    res = (24, 24, 12)
    mn, mx, gd, r = grid_for(orc, torus_small.vertices, res)
    want = orc.mesh(torus_small.vertices, torus_small.faces).sample_sdf(gd, r)
    sdf = dg.CubicLagrangeDiscreteGrid(mn, mx, res)
    sdf.addFunction(dg.MeshSignedDistance(dg.TriangleMeshDistance(torus_small)))
    assert bits_equal(sdf.m_nodes[0], want)


@pytest.mark.parametrize("name,res", [("bunny.obj", 24), ("dragon.obj", 20), ("happy_buddha.obj", 12)])
def test_reference_meshes_vs_oracle(dg, orc, name, res):
    """the reference's own meshes (staged under oracle/_ref/resources by `make -C oracle ref`)"""
    path = ref_resource(name)
    if path is None:
        pytest.skip(f"{name} not staged (oracle/_ref/resources)")
    mesh = dg.TriangleMesh(path)
    mn, mx, gd, r = grid_for(orc, mesh.vertices, (res, res, res))
    om = orc.mesh(mesh.vertices, mesh.faces)
    want = om.sample_sdf(gd, r)
    md = dg.TriangleMeshDistance(mesh)
    assert md.info()["watertight_flags"] == om.flags()
    sph, kids = md.tree()
    so, ko = om.tree()
    assert np.array_equal(kids, ko) and bits_equal(sph[ko[:, 0] != -1], so[ko[:, 0] != -1])
    sdf = dg.CubicLagrangeDiscreteGrid(mn, mx, (res, res, res))
    sdf.addFunction(dg.MeshSignedDistance(md))
    assert bits_equal(sdf.m_nodes[0], want)


def test_edge_cases(dg, orc):
    # single triangle: the root is a leaf
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float); F = np.array([[0, 1, 2]], np.uint32)
    md = dg.TriangleMeshDistance(V, F)
    x = np.array([[0.2, 0.2, 1.0], [2, 2, 0], [-1, -1, -1], [0.25, 0.25, 0.0]])
    got = md.signed_distance(x); want = orc.mesh(V, F).distance(x)
    assert bits_equal(got.distance, want[0]) and np.array_equal(got.nearest_entity, want[2])
    # two triangles, degenerate (zero-area) triangle included
    V2 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 0]], float); F2 = np.array([[0, 1, 2], [0, 1, 3]], np.uint32)
    got = dg.TriangleMeshDistance(V2, F2).unsigned_distance(x); want = orc.mesh(V2, F2).distance(x, signed=False)
    assert bits_equal(got.distance, want[0])
    # empty inputs
    with pytest.raises(dg.DiscregridError):
        dg.TriangleMeshDistance(np.zeros((0, 3)), np.zeros((0, 3), np.uint32))
    with pytest.raises(dg.DiscregridError):
        dg.TriangleMeshDistance(V, np.array([[0, 1, 7]], np.uint32))          # index out of range
    with pytest.raises(dg.DiscregridError):
        dg.TriangleMeshDistance().signed_distance(x)                           # not constructed
    assert len(md.signed_distance(np.zeros((0, 3))).distance) == 0             # empty batch


def test_full_size_config1_properties(dg, orc):
    """BASELINE.json configs[1] size (128^3 = 14,926,977 nodes, ~70k triangles) through the C-ABI: size-independent properties
    + oracle parity on a strided 60k-node sample (the oracle needs seconds for that, minutes for all nodes)."""
    import ctypes as C
    from discregrid_b200 import _capi as capi
    mesh = dg.bumpy_torus(186, 187)                                # the bench workload
    md = dg.TriangleMeshDistance(mesh)
    mn, mx = dg.generate_sdf_domain(mesh.vertices)
    res = (128, 128, 128)
    g = dg.CubicLagrangeDiscreteGrid(mn, mx, res)
    n = g.nNodes()
    assert n == 14926977
    g.addFunction(dg.MeshSignedDistance(md))
    full = g.m_nodes[0]
    assert np.isfinite(full).all() and (full < 0).any() and (full > 0).any()
    # invert: -1.0 * d is exact, so the inverted field is the bitwise negation
    inv = np.empty(n)
    capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(g._desc), -1.0, 0, n, capi.ptr(inv, capi.F64P)))
    assert bits_equal(inv, -full)
    # determinism + sharding: 7 uneven ranges concatenate to the same bits
    cuts = np.linspace(0, n, 8).astype(np.int64); cuts[3] += 12345
    parts = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        out = np.empty(e - b)
        capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(g._desc), 1.0, int(b), int(e), capi.ptr(out, capi.F64P)))
        parts.append(out)
    assert bits_equal(np.concatenate(parts), full)
    # |grad| of a distance field is 1 a.e.: vertex nodes one cell apart differ by at most the cell size (Lipschitz-1)
    nx = 129
    v = full[:nx ** 3].reshape(nx, nx, nx)                        # [k][j][i]
    assert np.abs(np.diff(v, axis=2)).max() <= g.m_cell_size[0] * (1 + 1e-12)
    assert np.abs(np.diff(v, axis=1)).max() <= g.m_cell_size[1] * (1 + 1e-12)
    assert np.abs(np.diff(v, axis=0)).max() <= g.m_cell_size[2] * (1 + 1e-12)
    # oracle parity on a strided sample through the point-query API and the node positions
    ids = np.linspace(0, n - 1, 60000).astype(np.int64)
    gd, r = orc.grid_desc(mn, mx, res)
    x = np.concatenate([orc.node_positions(gd, r, int(l), int(l) + 1) for l in ids[:200]])
    assert bits_equal(x, g.nodePositions()[ids[:200]])
    xs = g.nodePositions()[ids]
    want = orc.mesh(mesh.vertices, mesh.faces).distance(xs)[0]
    assert bits_equal(full[ids], want)


def test_meshes_of_different_depth_coexist(dg, orc, box_mesh, torus_small):
    """two live meshes with different tree depths (the shared-memory stack size is a per-kernel attribute)"""
    big = dg.TriangleMeshDistance(torus_small)          # deep tree first
    small = dg.TriangleMeshDistance(box_mesh)           # then a shallow one
    x = np.random.default_rng(3).uniform(-1.4, 1.4, (5000, 3))
    assert bits_equal(big.signed_distance(x).distance, orc.mesh(torus_small.vertices, torus_small.faces).distance(x)[0])
    assert bits_equal(small.signed_distance(x).distance, orc.mesh(box_mesh.vertices, box_mesh.faces).distance(x)[0])
    assert bits_equal(big.signed_distance(x).distance, orc.mesh(torus_small.vertices, torus_small.faces).distance(x)[0])


@pytest.mark.parametrize("res,parts", [((12, 10, 9), 3), ((16, 16, 16), 8), ((5, 4, 3), 4)])
def test_slab_parts_cover_the_grid(dg, orc, torus_small, res, parts):
    """dg_sample_sdf_slab_device: the parts of an n-way slab split, run one after another on one GPU, fill the full array with the
    single-launch result (what each rank of a multi-GPU job computes; the exchange is tested in test_gpu_multi / gloo)"""
    import ctypes as C
    import torch
    from discregrid_b200 import _capi as capi
    mn, mx, gd, r = grid_for(orc, torus_small.vertices, res)
    want = orc.mesh(torus_small.vertices, torus_small.faces).sample_sdf(gd, r)
    md = dg.TriangleMeshDistance(torus_small)
    desc = dg.grid_desc(mn, mx, res)
    full = torch.full((len(want),), float("nan"), dtype=torch.float64, device="cuda")
    for p in range(parts):
        capi.check(capi.lib.dg_sample_sdf_slab_device(md.handle, C.byref(desc), 1.0, p, parts, C.c_void_p(full.data_ptr()), None))
    torch.cuda.synchronize()
    assert bits_equal(full.cpu().numpy(), want)


@pytest.mark.parametrize("res,parts", [((12, 10, 9), 3), ((16, 16, 16), 8), ((5, 4, 3), 4), ((7, 3, 5), 16), ((9, 9, 9), 1)])
def test_interleaved_parts_cover_the_grid(dg, orc, torus_small, res, parts):
    """dg_sample_sdf_interleaved_device + dg_interleaved_unpack_device: plane pairs dealt round-robin over `parts`, every part run on
    this one GPU into its slot, unpacked == the single-launch result"""
    import ctypes as C
    import torch
    from discregrid_b200 import _capi as capi
    mn, mx, gd, r = grid_for(orc, torus_small.vertices, res)
    want = orc.mesh(torus_small.vertices, torus_small.faces).sample_sdf(gd, r)
    md = dg.TriangleMeshDistance(torus_small)
    desc = dg.grid_desc(mn, mx, res)
    se = C.c_uint64(); capi.check(capi.lib.dg_interleaved_slot_elems(C.byref(desc), parts, C.byref(se)))
    slots = torch.full((parts * se.value,), float("nan"), dtype=torch.float64, device="cuda")
    for p in range(parts):
        capi.check(capi.lib.dg_sample_sdf_interleaved_device(md.handle, C.byref(desc), 1.0, p, parts, C.c_void_p(slots.data_ptr() + 8 * p * se.value), None))
    full = torch.full((len(want),), float("nan"), dtype=torch.float64, device="cuda")
    capi.check(capi.lib.dg_interleaved_unpack_device(C.byref(desc), parts, C.c_void_p(slots.data_ptr()), C.c_void_p(full.data_ptr()), None))
    torch.cuda.synchronize()
    assert bits_equal(full.cpu().numpy(), want)
    # ... and the slots hold exactly what the host statement of the layout (dg_interleaved_node_slots) says
    from discregrid_b200.distributed import interleaved_node_slots
    part, pos = interleaved_node_slots(desc, parts, 0, len(want))
    assert bits_equal(slots.cpu().numpy()[part.astype(np.int64) * se.value + pos.astype(np.int64)], want)
    assert capi.lib.dg_interleaved_slot_elems(C.byref(desc), 17, C.byref(se)) == capi.DG_ERR_INVALID
