#!/usr/bin/env python
"""Event counts of the REAL K1 kernel code run by tests/emu (build/bin/libk1emu.so; `make cpp`): node steps, how many of them fall back to
the fp64 sphere test, leaf tests / acceptances, deferred re-tests and how they are resolved -- per query, on the bench mesh or an OBJ.
usage: tools/k1_emu_profile.py [resolution=24] [mesh.obj]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import discregrid_b200 as dg
import bench
from oracle_api import Oracle
res = int(sys.argv[1]) if len(sys.argv) > 1 else 24
mesh = dg.TriangleMesh(sys.argv[2]) if len(sys.argv) > 2 else dg.bumpy_torus(*bench.WORKLOAD["torus"])
lib = C.CDLL(os.environ.get("DG_K1_EMU_LIB", os.path.join(ROOT, "build", "bin", "libk1emu.so")))
lib.emu_mesh_create.restype = C.c_void_p
dp, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
lib.emu_mesh_create.argtypes = [dp, C.c_uint64, u32p, C.c_uint64]
lib.emu_sample_sdf.argtypes = [C.c_void_p, dp, u32p, C.c_double, C.c_uint64, C.c_uint64, dp]
orc = Oracle()
V = np.ascontiguousarray(mesh.vertices, np.float64); F = np.ascontiguousarray(mesh.faces, np.uint32)
mn, mx = orc.generate_sdf_domain(V)
gd, r = orc.grid_desc(mn, mx, (res,) * 3)
n = 4 * (res + 1) ** 3 - 3 * (res + 1) ** 2 * 1 if False else None
nn = (res + 1) ** 3 + 2 * 3 * res * (res + 1) ** 2
h = lib.emu_mesh_create(V.ctypes.data_as(dp), len(V), F.ctypes.data_as(u32p), len(F))
out = np.empty(nn)
cnt = (C.c_ulonglong * 32)()
lib.emu_counters(cnt, 1)
assert lib.emu_sample_sdf(h, gd.ctypes.data_as(dp), r.ctypes.data_as(u32p), 1.0, 0, nn, out.ctypes.data_as(dp)) == 0
lib.emu_counters(cnt, 1)
c = [int(x) for x in cnt]
q = float(nn)
print(f"{len(F)} triangles, {res}^3 grid, {nn} nodes (emulated kernel code)")
print(f"warp iterations per brick      {c[0] / 32 / (q / 32):8.1f}   (lane-loop trips / 32, per 32 queries)")
print(f"node steps per query           {c[1] / q:8.1f}   fp64 fallback {100.0 * c[2] / max(1, c[1]):5.2f} %")
print(f"leaf tests per query           {c[3] / q:8.1f}   accepted {c[4] / q:5.2f} per query")
print(f"deferred pushes per query      {c[9] / q:8.1f}")
print(f"deferred re-tests per query    {c[5] / q:8.1f}   fp64 fallback {100.0 * c[6] / max(1, c[5]):5.2f} %, dropped by the box test {100.0 * c[7] / max(1, c[5]):5.2f} %")
