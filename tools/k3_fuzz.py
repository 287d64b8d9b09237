#!/usr/bin/env python
"""Parity fuzzer (CPU) for the density map: the product's K3 kernel code run by tests/emu (build/bin/libk23emu*.so) against the oracle's
restatement of GenerateDensityMap's node function (itself pinned to the reference tool's .cdm output) on random small grids: smooth and rough
SDF-like fields, fields with DBL_MAX sentinels, both predicate modes, support radii from a fraction of a cell to several cells, node
sub-ranges.  usage: tools/k3_fuzz.py [rounds=40] [seed=0] [lib]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_api import Oracle
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
lib = C.CDLL(sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "build", "bin", "libk23emu.so"))
dp, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
lib.emu_field_create.restype = C.c_void_p
lib.emu_field_create.argtypes = [dp, u32p, dp, u32p, C.c_uint64, u32p]
lib.emu_field_destroy.argtypes = [C.c_void_p]
lib.emu_density_map.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_uint64, dp]
orc = Oracle()
DBL_MAX = np.finfo(np.float64).max
bad = total = 0
for k in range(rounds):
    res = tuple(int(v) for v in rng.integers(1, 5, 3))
    mn = rng.standard_normal(3); mx = mn + 0.5 + rng.random(3) * 2.0
    gd, r = orc.grid_desc(mn, mx, res)
    cells = orc.build_cells(r)
    n = int(cells.max()) + 1
    x = orc.node_positions(gd, r, 0, n)
    ctr = 0.5 * (mn + mx) + 0.2 * rng.standard_normal(3)
    v = np.linalg.norm(x - ctr, axis=1) - (0.3 + 0.5 * rng.random())                 # signed distance to a sphere
    if k % 4 == 1: v += 0.05 * rng.standard_normal(n)
    if k % 4 == 2: v[rng.random(n) < 0.03] = DBL_MAX
    h = float((0.15 + rng.random() * 2.0) * gd[6:9].min())
    rho0 = float(10.0 ** rng.integers(0, 4))
    no_red = int(k % 2)
    l0 = int(rng.integers(0, n // 2)) if k % 3 == 0 else 0
    l1 = int(rng.integers(l0 + 1, n + 1)) if k % 3 == 0 else n
    want = orc.density_map(gd, r, v, h, rho0, bool(no_red), l0, l1)
    nodes = np.ascontiguousarray(v)
    hf = lib.emu_field_create(gd.ctypes.data_as(dp), r.ctypes.data_as(u32p), nodes.ctypes.data_as(dp), None, len(cells), None)
    got = np.full(l1 - l0, np.nan)
    lib.emu_density_map(hf, h, rho0, no_red, l0, l1, got.ctypes.data_as(dp))
    lib.emu_field_destroy(hf)
    ok = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
    total += len(got)
    if not ok.all():
        i = int(np.nonzero(~ok)[0][0]); bad += int((~ok).sum())
        print(f"MISMATCH case {k} res {res} h {h} node {l0 + i}: {got[i]!r} vs {want[i]!r}")
print(f"{rounds} fields, {total} nodes, {bad} mismatches")
sys.exit(1 if bad else 0)
