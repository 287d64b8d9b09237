"""The C++ facade (cpp/include/Discregrid: the reference's class API over the C-ABI) and the two reference tools rebuilt on
it, run as a downstream C++ caller would, checked against the golden box.cdf and the oracle.  Bar: bit-exact."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, bits_equal

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "build", "bin")
DBL_MAX = np.finfo(np.float64).max


def _need(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (make cpp)")
    return p


def test_generate_sdf_tool_reproduces_box_cdf(tmp_path):
    exe = _need("GenerateSDF")
    out = tmp_path / "box.cdf"
    r = subprocess.run([exe, "-r", "5 5 5", "-o", str(out), os.path.join(GOLDEN, "box.obj")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert out.read_bytes() == open(os.path.join(GOLDEN, "box.cdf"), "rb").read()      # byte-identical to the reference's output


@pytest.mark.parametrize("n_gpus", [2, 8])
def test_generate_sdf_tool_multi_gpu_same_file(dg, tmp_path, n_gpus):
    """GenerateSDF --gpus N (dg_add_function_sdf_multi through the facade) writes the same bytes as the single-GPU run"""
    exe = _need("GenerateSDF")
    if dg.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    a, b = tmp_path / "one.cdf", tmp_path / "many.cdf"
    for out, extra in ((a, []), (b, ["--gpus", str(n_gpus)])):
        r = subprocess.run([exe, "-r", "21 17 12", "-i", "-o", str(out)] + extra + [os.path.join(GOLDEN, "sphere.obj")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    assert a.read_bytes() == b.read_bytes()


def test_facade_api_vs_oracle(dg, orc, tmp_path):
    exe = _need("facade_check")
    g = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    lo, hi = g.m_domain
    rng = np.random.default_rng(5)
    x = lo - 0.05 + rng.random((3000, 3)) * (hi - lo + 0.1)
    x[:3] = [lo, hi, 0.5 * (lo + hi)]
    pts = tmp_path / "p.bin"; x.tofile(pts)
    out = tmp_path / "o.bin"
    r = subprocess.run([exe, os.path.join(GOLDEN, "box.cdf"), os.path.join(GOLDEN, "box.obj"), str(pts), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    o = np.fromfile(out); n = len(x); k = 0
    gd, res = orc.grid_desc(lo, hi, g.m_resolution, g.m_cell_size, g.m_inv_cell_size)
    po, go = orc.interpolate(gd, res, g.m_nodes[0], x, grad=True, cells=g.m_cells[0], cell_map=g.m_cell_map[0])
    outside = ~np.all((lo <= x) & (x <= hi), axis=1)
    go_ref = go.copy(); go_ref[outside] = 7.0                                # the reference leaves *gradient untouched there
    a = o[k:k + 4 * n].reshape(n, 4); k += 4 * n
    assert bits_equal(a[:, 0], po) and bits_equal(a[:, 1:], go_ref)
    b = o[k:k + 5 * n].reshape(n, 5); k += 5 * n
    assert np.array_equal(b[:, 0] == 1.0, ~outside)
    assert bits_equal(b[~outside, 1], po[~outside]) and bits_equal(b[~outside, 2:], go[~outside])
    assert bits_equal(o[k:k + n], po); k += n
    box = dg.TriangleMesh(os.path.join(GOLDEN, "box.obj"))
    om = orc.mesh(box.vertices, box.faces)
    d, near, ent, tri = om.distance(x[:64])
    c = o[k:k + 6 * 64].reshape(64, 6); k += 6 * 64
    assert bits_equal(c[:, 0], d) and bits_equal(c[:, 1:4], near) and np.array_equal(c[:, 4], ent) and np.array_equal(c[:, 5], tri)
    assert bits_equal(o[k:k + n], om.distance(x, signed=False)[0]); k += n
    assert o[k] == 0.0; k += 1
    gd2, res2 = orc.grid_desc(lo, hi, (3, 4, 2))
    want = om.sample_sdf(gd2, res2, sign=-1.0)
    assert bits_equal(o[k:k + len(want)], want); k += len(want)
    assert o[k] == 1.0 and k + 1 == len(o)                                   # opaque lambda rejected, nothing left over


def test_generate_density_map_tool(dg, orc, tmp_path):
    sdf_exe, dm_exe = _need("GenerateSDF"), _need("GenerateDensityMap")
    cdf = tmp_path / "s.cdf"; cdm = tmp_path / "s.cdm"
    sph = dg.uv_sphere(10, 16, 0.5); obj = tmp_path / "s.obj"; sph.exportOBJ(str(obj))
    assert subprocess.run([sdf_exe, "-r", "10 10 10", "-d", "-2 -2 -2 2 2 2", "-o", str(cdf), str(obj)], capture_output=True).returncode == 0
    r = subprocess.run([dm_exe, "-s", "0.03", "-r", "1000", "--no-reduction", "-o", str(cdm), str(cdf)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    g = dg.CubicLagrangeDiscreteGrid(str(cdm))
    assert g.nFields() == 2
    gd, res = orc.grid_desc(g.m_domain[0], g.m_domain[1], g.m_resolution, g.m_cell_size, g.m_inv_cell_size)
    assert bits_equal(g.m_nodes[1], orc.density_map(gd, res, g.m_nodes[0], 0.03, 1000.0, True, 0, len(g.m_nodes[0])))
    # with reduction: the reduced file must load, interpolate consistently with the unreduced field wherever cells survive
    cdm2 = tmp_path / "r.cdm"
    r = subprocess.run([dm_exe, "-s", "0.03", "-r", "1000", "-o", str(cdm2), str(cdf)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    gr = dg.CubicLagrangeDiscreteGrid(str(cdm2))
    assert len(gr.m_nodes[0]) < len(g.m_nodes[0]) and (gr.m_cell_map[0] == 0xFFFFFFFF).any()
    x = g.m_domain[0] + np.random.default_rng(1).random((20000, 3)) * (g.m_domain[1] - g.m_domain[0])
    full0 = g.interpolate(0, x); red0 = gr.interpolate(0, x)
    alive = red0 != DBL_MAX
    assert alive.any() and (~alive).any() and bits_equal(red0[alive], full0[alive])     # same coefficients, same arithmetic
    # z-sorted node order: positions' Morton keys are non-decreasing is an internal detail; coefficient multiset is preserved
    assert set(np.unique(gr.m_nodes[0])) <= set(np.unique(g.m_nodes[0]))


def test_discrete_field_to_bitmap_tool(dg, orc, tmp_path):
    """N4: the reference's batched-interpolate consumer on the batch API: pixel values == oracle interpolation at the same samples"""
    exe = _need("DiscreteFieldToBitmap")
    out = tmp_path / "box.bmp"
    r = subprocess.run([exe, "-s", "96", "-p", "xz", "-d", "0.25", "-o", str(out), os.path.join(GOLDEN, "box.cdf")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = out.read_bytes()
    assert raw[:2] == b"BM" and int.from_bytes(raw[18:22], "little") == 96 and int.from_bytes(raw[22:26], "little") == 96
    g = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    lo, hi = g.m_domain; diag = hi - lo
    k = np.arange(96 * 96); i, j = k % 96, k // 96
    x = np.empty((len(k), 3))
    x[:, 0] = lo[0] + i / 96 * diag[0] + 0.5 * diag[0] / 96
    x[:, 2] = lo[2] + j / 96 * diag[2] + 0.5 * diag[2] / 96
    x[:, 1] = lo[1] + 0.5 * 1.25 * diag[1]
    gd, res = orc.grid_desc(lo, hi, g.m_resolution, g.m_cell_size, g.m_inv_cell_size)
    v = orc.interpolate(gd, res, g.m_nodes[0], x, grad=False)[0]
    v[v == DBL_MAX] = 0.0
    vn = np.where(v >= 0, v / abs(v.max()), v / abs(v.min()))
    green = np.where(vn >= 0, np.clip(255.0 * (1 - vn), 0, 255), 0).astype(np.uint8)
    blue = np.where(vn < 0, np.clip(255.0 * (1 + vn), 0, 255), 0).astype(np.uint8)
    px = np.frombuffer(raw, np.uint8, 96 * 96 * 3, 54).reshape(-1, 3)          # BGR, 96*3 is a multiple of 4: no padding
    assert np.array_equal(px[:, 1], green) and np.array_equal(px[:, 0], blue) and not px[:, 2].any()
