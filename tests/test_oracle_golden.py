"""Pins the CPU oracle (oracle/dg_oracle.cpp) -- runs WITHOUT a GPU.
  * against the reference's only golden vector, cmd/generate_sdf/resources/box.cdf (committed copy in tests/golden/);
  * against golden vectors generated from the reference's own, unmodified TriangleMeshDistance.h
    (tests/golden/make_golden.py, oracle/_ref/libdgref.so);
  * and, when oracle/_ref is built (the build container), live against that library on the reference's meshes."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal, ref_resource
from oracle_api import RefMesh, have_ref


def read_cdf(path):
    """stand-alone reader of the reference's .cdf layout (cubic_lagrange_discrete_grid.cpp:678-719) -- deliberately not
    the product's loader"""
    raw = open(path, "rb").read()
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, raw, off); off += struct.calcsize(fmt); return v
    mn = np.array(take("<3d")); mx = np.array(take("<3d")); res = np.array(take("<3I"), np.uint32)
    cell = np.array(take("<3d")); inv = np.array(take("<3d")); n_cells, n_fields = take("<QQ")

    def nested(dtype, width):
        nonlocal off
        (outer,) = take("<Q"); out = []
        for _ in range(outer):
            (n,) = take("<Q")
            a = np.frombuffer(raw, dtype, n * width, off).copy(); off += a.nbytes
            out.append(a.reshape(n, width) if width > 1 else a)
        return out
    nodes, cells, cmap = nested(np.float64, 1), nested(np.uint32, 32), nested(np.uint32, 1)
    assert off == len(raw)
    return dict(mn=mn, mx=mx, res=res, cell=cell, inv=inv, n_cells=n_cells, n_fields=n_fields, nodes=nodes, cells=cells, cmap=cmap)


def read_obj(path):
    V, F = [], []
    for line in open(path):
        if line.startswith("v "):
            V.append([float(t) for t in line.split()[1:4]])
        elif line.startswith("f "):
            F.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
    return np.array(V), np.array(F, np.uint32)


def test_box_cdf_reproduced_bit_exactly(orc):
    g = read_cdf(os.path.join(GOLDEN, "box.cdf"))
    V, F = read_obj(os.path.join(GOLDEN, "box.obj"))
    assert len(V) == 8 and len(F) == 12 and g["n_cells"] == 125 and g["n_fields"] == 1 and len(g["nodes"][0]) == 1296
    mn, mx = orc.generate_sdf_domain(V)                       # asymmetric padding, cmd/generate_sdf/main.cpp:89-90
    assert bits_equal(mn, g["mn"]) and bits_equal(mx, g["mx"])
    gd, res = orc.grid_desc(mn, mx, g["res"])
    assert bits_equal(gd[6:9], g["cell"]) and bits_equal(gd[9:12], g["inv"])     # discrete_grid.hpp:22-29
    assert orc.num_nodes(res) == 1296
    coeffs = orc.mesh(V, F).sample_sdf(gd, res)
    assert bits_equal(coeffs, g["nodes"][0])                  # all 1296 signed distances
    assert np.array_equal(orc.build_cells(res), g["cells"][0])
    assert np.array_equal(g["cmap"][0], np.arange(125, dtype=np.uint32))


def test_reference_header_golden_queries(orc, torus_small):
    g = np.load(os.path.join(GOLDEN, "ref_torus_queries.npz"))
    m = orc.mesh(torus_small.vertices, torus_small.faces)
    d, near, ent, tri = m.distance(g["x"], signed=True)
    assert bits_equal(d, g["distance"]) and bits_equal(near, g["nearest"])
    assert np.array_equal(ent, g["entity"]) and np.array_equal(tri, g["triangle"])
    assert bits_equal(m.distance(g["x"], signed=False)[0], g["unsigned"])


def test_reference_header_golden_tree(orc, torus_small):
    g = np.load(os.path.join(GOLDEN, "ref_torus_tree.npz"))
    m = orc.mesh(torus_small.vertices, torus_small.faces)
    sph, kids = m.tree()
    assert np.array_equal(kids, g["kids"]) and bits_equal(sph[kids[:, 0] != -1], g["spheres_internal"])
    pt, pe, pv = m.pseudonormals()
    assert bits_equal(pt, g["pn_tri"]) and bits_equal(pe, g["pn_edge"]) and bits_equal(pv, g["pn_vert"])
    assert m.flags() == 0                                      # watertight


def test_reference_header_golden_surface_points(orc):
    from discregrid_b200.mesh import uv_sphere
    g = np.load(os.path.join(GOLDEN, "ref_sphere_surface.npz"))
    a = g["sphere_args"]
    s = uv_sphere(int(a[0]), int(a[1]), a[2], tuple(a[3:6]))
    d, near, ent, tri = orc.mesh(s.vertices, s.faces).distance(g["x"], signed=True)
    assert bits_equal(d, g["distance"]) and bits_equal(near, g["nearest"])
    assert np.array_equal(ent, g["entity"]) and np.array_equal(tri, g["triangle"])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libdgref.so not built (needs /root/reference)")
@pytest.mark.parametrize("name,n", [("bunny.obj", 20000), ("dragon.obj", 20000), ("happy_buddha.obj", 3000)])
def test_live_against_reference_header(orc, name, n):
    path = ref_resource(name)
    if path is None:
        pytest.skip("mesh not staged")
    V, F = read_obj(path)
    om, rm = orc.mesh(V, F), RefMesh(V, F)
    so, ko = om.tree(); sr, kr = rm.tree()
    assert np.array_equal(ko, kr) and bits_equal(so[kr[:, 0] != -1], sr[kr[:, 0] != -1])
    for a, b in zip(om.pseudonormals(), rm.pseudonormals()):
        assert np.array_equal(a, b, equal_nan=True)
    mn, mx = orc.generate_sdf_domain(V)
    x = mn + np.random.default_rng(7).random((n, 3)) * (mx - mn)
    for signed in (True, False):
        for a, b in zip(om.distance(x, signed), rm.distance(x, signed)):
            assert np.array_equal(a, b)


def test_shape_function_identities(orc):
    """the reference stores no interpolate() output; A8/A9 are pinned by identities:
    nodal property at the 32 abscissae (cubic_lagrange_discrete_grid.cpp:58-94), partition of unity, FD Jacobian (:1028-1042)"""
    t = 1.0 / 3.0
    absc = [[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [1, 1, -1], [-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]]
    absc += [[s * t, y, z] for (y, z) in ((-1, -1), (-1, 1), (1, -1), (1, 1)) for s in (-1, 1)]
    absc += [[x, s * t, z] for (x, z) in ((-1, -1), (1, -1), (-1, 1), (1, 1)) for s in (-1, 1)]
    absc += [[x, y, s * t] for (x, y) in ((-1, -1), (-1, 1), (1, -1), (1, 1)) for s in (-1, 1)]
    N, _ = orc.shape_functions(np.array(absc, float))
    assert np.max(np.abs(N - np.eye(32))) < 1e-15
    xi = np.random.default_rng(0).uniform(-1, 1, (2000, 3))
    N, dN = orc.shape_functions(xi)
    assert np.max(np.abs(N.sum(1) - 1)) < 1e-14 and np.max(np.abs(dN.sum(1))) < 1e-13
    eps = 1e-6
    for d in range(3):
        xp, xm = xi.copy(), xi.copy(); xp[:, d] += eps; xm[:, d] -= eps
        fd = (orc.shape_functions(xp, False)[0] - orc.shape_functions(xm, False)[0]) / (2 * eps)
        assert np.max(np.abs(fd - dN[:, :, d])) < 1e-8


def test_interpolation_reproduces_nodes_of_box_cdf(orc):
    g = read_cdf(os.path.join(GOLDEN, "box.cdf"))
    gd, res = orc.grid_desc(g["mn"], g["mx"], g["res"], g["cell"], g["inv"])
    x = orc.node_positions(gd, res, 0, 1296)
    phi, grad = orc.interpolate(gd, res, g["nodes"][0], x, grad=True, cells=g["cells"][0], cell_map=g["cmap"][0])
    inside = phi != np.finfo(float).max
    assert inside.mean() > 0.99 and np.max(np.abs(phi[inside] - g["nodes"][0][inside])) < 1e-14
    # the box SDF has |grad| = 1 away from the medial axis / surface kinks: sanity of the Jacobian path
    far = np.abs(g["nodes"][0]) > 0.2
    assert np.median(np.linalg.norm(grad[inside & far], axis=1)) == pytest.approx(1.0, abs=0.05)
    # closed-form connectivity == stored table
    p2, g2 = orc.interpolate(gd, res, g["nodes"][0], x, grad=True)
    assert bits_equal(phi, p2) and bits_equal(grad, g2)
