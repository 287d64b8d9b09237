#!/usr/bin/env python
"""Where does dg_add_function_sdf's time go?  device single launch vs dg_sample_sdf (host buffer, no tables) vs dg_add_function_sdf, at
the headline (bunny 128^3) and target (256^3 / 100k triangles) configurations.  usage: DG_HOST_THREADS=n python tools/e2e_probe.py [target]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import discregrid_b200 as dg
from discregrid_b200 import _capi as capi
import bench
which = sys.argv[1] if len(sys.argv) > 1 else "bunny"
mesh = bench.workload_mesh(dg, "target" if which == "target" else "bunny")[0]
res = [256] * 3 if which == "target" else [128] * 3
md = dg.TriangleMeshDistance(mesh)
mn, mx = dg.generate_sdf_domain(mesh.vertices)
desc = dg.grid_desc(mn, mx, res)
n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
nc = res[0] ** 3
dev = torch.empty(n, dtype=torch.float64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t_dev():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, 0, n, C.c_void_p(dev.data_ptr()), st)); torch.cuda.synchronize()
    return time.perf_counter() - t0
def t_host(fresh):
    out = np.empty(n) if fresh else t_host.buf
    t0 = time.perf_counter()
    capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(desc), 1.0, 0, n, capi.ptr(out, capi.F64P)))
    return time.perf_counter() - t0
t_host.buf = np.zeros(n)
def t_add(tables, fresh=True):
    t0 = time.perf_counter()
    nodes = np.empty(n); cells = np.empty((nc, 32), np.uint32) if tables else None; cm = np.empty(nc, np.uint32) if tables else None; tm = np.zeros(6)
    capi.check(capi.lib.dg_add_function_sdf(md.handle, C.byref(desc), 1.0, capi.ptr(nodes, capi.F64P), capi.ptr(cells, capi.U32P), capi.ptr(cm, capi.U32P), capi.ptr(tm, capi.F64P)))
    return time.perf_counter() - t0
for name, fn in (("device single launch", t_dev), ("dg_sample_sdf warm buffer", lambda: t_host(False)), ("dg_sample_sdf fresh buffer", lambda: t_host(True)),
                 ("dg_add_function_sdf nodes only", lambda: t_add(False)), ("dg_add_function_sdf + tables", lambda: t_add(True))):
    fn(); ts = [fn() for _ in range(3)]
    print(f"{which} DG_HOST_THREADS={os.environ.get('DG_HOST_THREADS','-')} {name}: {min(ts)*1e3:.1f} ms (mean {np.mean(ts)*1e3:.1f})", flush=True)
