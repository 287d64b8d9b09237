"""The PRODUCT's host-side BVH / pseudonormal builder (discregrid_b200/csrc/bvh_build.cpp), checked WITHOUT a GPU against the
oracle and the reference-generated golden tree: same tree (children + internal spheres), same pseudonormals, same flags; the
precomputed triangle records and the fp32 filter shadows are consistent with the fp64 data they shadow."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, bits_equal, ref_resource

EXE = os.path.join(ROOT, "build", "bin", "bvh_host_check")


def run_builder(tmp_path, V, F):
    if not os.path.exists(EXE):
        pytest.skip("build/bin/bvh_host_check not built (make cpp)")
    V = np.ascontiguousarray(V, np.float64); F = np.ascontiguousarray(F, np.uint32)
    V.tofile(tmp_path / "V.bin"); F.tofile(tmp_path / "F.bin")
    r = subprocess.run([EXE, str(tmp_path / "V.bin"), str(tmp_path / "F.bin"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(tmp_path / "o.bin", "rb").read()
    nT, nV = len(F), len(V)
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype, count, off); off += a.nbytes; return a
    hdr = take(np.float64, 4); nn = int(hdr[0])
    out = dict(depth=int(hdr[1]), flags=int(hdr[2]), half_extent=hdr[3])
    out["spheres"] = take(np.float64, 8 * nn).reshape(nn, 8); out["kids"] = take(np.int32, 2 * nn).reshape(nn, 2)
    out["pn_tri"] = take(np.float64, 3 * nT).reshape(nT, 3); out["pn_edge"] = take(np.float64, 9 * nT).reshape(nT, 3, 3)
    out["pn_vert"] = take(np.float64, 3 * nV).reshape(nV, 3)
    out["leaves"] = take(np.float64, 16 * nT).reshape(nT, 16)
    out["spheres_f"] = take(np.float32, 8 * nT).reshape(nT, 8); out["boxes_f"] = take(np.float32, 12 * nT).reshape(nT, 12)
    out["center"] = take(np.float64, 3)
    assert off == len(raw)
    return out


def check_against_oracle(orc, got, V, F):
    om = orc.mesh(V, F)
    so, ko = om.tree()
    assert np.array_equal(got["kids"], ko)
    internal = ko[:, 0] != -1
    assert bits_equal(got["spheres"][internal], so[internal])
    pt, pe, pv = om.pseudonormals()
    assert np.array_equal(got["pn_tri"], pt, equal_nan=True) and np.array_equal(got["pn_edge"], pe, equal_nan=True)
    assert np.array_equal(got["pn_vert"], pv, equal_nan=True)
    assert got["flags"] == om.flags()
    # -0.0 vs +0.0 must also agree wherever the values are finite
    for a, b in ((got["pn_tri"], pt), (got["pn_edge"], pe), (got["pn_vert"], pv)):
        fin = np.isfinite(a) & np.isfinite(b)
        assert np.array_equal(np.signbit(a[fin]), np.signbit(b[fin]))


def test_golden_reference_tree(orc, torus_small, tmp_path):
    g = np.load(os.path.join(GOLDEN, "ref_torus_tree.npz"))
    got = run_builder(tmp_path, torus_small.vertices, torus_small.faces)
    assert np.array_equal(got["kids"], g["kids"]) and bits_equal(got["spheres"][g["kids"][:, 0] != -1], g["spheres_internal"])
    assert bits_equal(got["pn_tri"], g["pn_tri"]) and bits_equal(got["pn_edge"], g["pn_edge"]) and bits_equal(got["pn_vert"], g["pn_vert"])
    assert got["flags"] == 0


def test_box_and_open_and_nonmanifold_meshes(orc, box_mesh, tmp_path):
    check_against_oracle(orc, run_builder(tmp_path, box_mesh.vertices, box_mesh.faces), box_mesh.vertices, box_mesh.faces)
    # open mesh (one face removed) -> flag bit0; fin attached to an edge (3 triangles on one edge) -> flag bit1; degenerate triangle
    V = np.vstack([box_mesh.vertices, [[3.0, 3.0, 3.0]]])
    F_open = box_mesh.faces[:-1]
    got = run_builder(tmp_path, V, F_open); check_against_oracle(orc, got, V, F_open); assert got["flags"] & 1
    F_fin = np.vstack([box_mesh.faces, [[box_mesh.faces[0][0], box_mesh.faces[0][1], 8]]]).astype(np.uint32)
    got = run_builder(tmp_path, V, F_fin); check_against_oracle(orc, got, V, F_fin); assert got["flags"] & 2
    F_deg = np.vstack([box_mesh.faces, [[0, 0, 1]]]).astype(np.uint32)           # repeated vertex: NaN normal, edge (0,1) counted twice
    got = run_builder(tmp_path, V, F_deg); check_against_oracle(orc, got, V, F_deg)


def test_records_and_fp32_shadows(orc, torus_small, tmp_path):
    V, F = torus_small.vertices, torus_small.faces
    got = run_builder(tmp_path, V, F)
    kids = got["kids"]
    # leaf order = order of the leaves in the tree; record = point-independent terms of point_triangle_sq_unsigned
    leaf_tris = kids[kids[:, 0] == -1][:, 1]
    L = got["leaves"]
    ids = np.ascontiguousarray(L[:, 15]).view(np.int32)[::2]
    assert np.array_equal(ids, leaf_tris)
    v0, v1, v2 = V[F[ids, 0]], V[F[ids, 1]], V[F[ids, 2]]
    e0, e1 = v1 - v0, v2 - v0
    dot = lambda a, b: a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]
    a00, a01, a11 = dot(e0, e0), dot(e0, e1), dot(e1, e1)
    assert bits_equal(L[:, 0:3], v0) and bits_equal(L[:, 3:6], e0) and bits_equal(L[:, 6:9], e1)
    assert bits_equal(L[:, 9], a00) and bits_equal(L[:, 10], a01) and bits_equal(L[:, 11], a11)
    det = np.abs(a00 * a11 - a01 * a01)
    assert bits_equal(L[:, 12], det) and bits_equal(L[:, 13], 1 / det) and bits_equal(L[:, 14], a00 - 2 * a01 + a11)
    # fp32 spheres shadow the fp64 ones within rounding; fp32 boxes CONTAIN the vertices of their subtree
    ctr = got["center"]
    internal = np.nonzero(kids[:, 0] != -1)[0]
    sph64 = got["spheres"][internal]
    # implicit index m of an internal node = number of leaves left of its split = leaves in the left subtree + leaves before the node
    n_leaves = np.zeros(len(kids), np.int64)
    for i in range(len(kids) - 1, -1, -1):
        n_leaves[i] = 1 if kids[i, 0] == -1 else n_leaves[kids[i, 0]] + n_leaves[kids[i, 1]]
    begin = np.zeros(len(kids), np.int64)
    for i in range(len(kids)):
        if kids[i, 0] != -1:
            begin[kids[i, 0]] = begin[i]; begin[kids[i, 1]] = begin[i] + n_leaves[kids[i, 0]]
    m = begin[internal] + n_leaves[kids[internal, 0]]
    sf = got["spheres_f"][m].astype(np.float64)
    rel = sph64.copy(); rel[:, 0:3] -= ctr; rel[:, 4:7] -= ctr
    assert np.max(np.abs(sf - rel)) <= 2.0 ** -23 * max(1.0, np.abs(rel).max())
    assert got["half_extent"] == pytest.approx(np.abs(V - ctr).max(), rel=1e-12)
    bx = got["boxes_f"][m].astype(np.float64)
    for side, child in ((0, kids[internal, 0]), (6, kids[internal, 1])):
        for k in np.random.default_rng(0).choice(len(internal), 300, replace=False):
            lo, cnt = begin[child[k]], n_leaves[child[k]]
            pts = V[F[ids[lo:lo + cnt]].ravel()] - ctr
            assert (pts >= bx[k, side:side + 3] - 0).all() and (pts <= bx[k, side + 3:side + 6] + 0).all()


@pytest.mark.parametrize("name", ["bunny.obj", "happy_buddha.obj"])
def test_reference_meshes(dg, orc, name, tmp_path):
    path = ref_resource(name)
    if path is None:
        pytest.skip("mesh not staged")
    mesh = dg.TriangleMesh(path)
    check_against_oracle(orc, run_builder(tmp_path, mesh.vertices, mesh.faces), mesh.vertices, mesh.faces)


def test_division_by_known_reciprocal_is_the_ieee_quotient():
    """tests/cpp/fast_div_check.cpp: discregrid_b200/csrc/fast_div.h (K3's division by the smoothing length through its reciprocal; round 2 measured the same trick in K1's leaf test slower and removed it there)
    gives num / den bit for bit on 3e7 operand pairs, including significands next to 1 and 2 and exactly divisible ones"""
    exe = os.path.join(ROOT, "build", "bin", "fast_div_check")
    if not os.path.exists(exe):
        pytest.skip("build/bin/fast_div_check not built (make cpp)")
    r = subprocess.run([exe, "30000000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
