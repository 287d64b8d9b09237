#!/usr/bin/env python
"""dg_add_function_sdf_multi on all visible GPUs: whole addFunction from one process, pageable vs page-locked node array, with and without the
index tables; against the single-GPU call.  usage: python tools/multi_probe.py [bunny|target]"""
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
ng = dg.device_count()
def single(tables):
    nodes = np.empty(n); cells = np.empty((nc, 32), np.uint32) if tables else None; cm = np.empty(nc, np.uint32) if tables else None
    t0 = time.perf_counter()
    capi.check(capi.lib.dg_add_function_sdf(md.handle, C.byref(desc), 1.0, capi.ptr(nodes, capi.F64P), capi.ptr(cells, capi.U32P), capi.ptr(cm, capi.U32P), None))
    return time.perf_counter() - t0, nodes
ref = single(False)[1]
for g in sorted({1, 2, ng} & set(range(1, ng + 1))):
    grp = C.c_void_p(); capi.check(capi.lib.dg_mesh_group_create(md.handle, g, None, C.byref(grp)))
    for pinned in (False, True):
        for tables in (False, True):
            ts = []
            for it in range(4):
                if pinned:
                    nt = torch.empty(n, dtype=torch.float64).pin_memory(); nodes = nt.numpy()
                else:
                    nodes = np.empty(n)
                cells = np.empty((nc, 32), np.uint32) if tables else None; cm = np.empty(nc, np.uint32) if tables else None; tm = np.zeros(6)
                t0 = time.perf_counter()
                capi.check(capi.lib.dg_add_function_sdf_multi(grp, C.byref(desc), 1.0, capi.ptr(nodes, capi.F64P), capi.ptr(cells, capi.U32P), capi.ptr(cm, capi.U32P), capi.ptr(tm, capi.F64P)))
                if it: ts.append(time.perf_counter() - t0)
            ok = np.array_equal(nodes.view(np.uint64), ref.view(np.uint64))
            print(f"{which} gpus={g} pinned={int(pinned)} tables={int(tables)}: {min(ts)*1e3:.1f} ms (node pipeline {tm[1]:.1f} ms, workers {int(tm[4])}) equal={ok}", flush=True)
    capi.lib.dg_mesh_group_destroy(grp)
for tables in (False, True):
    ts = [single(tables)[0] for _ in range(3)]
    print(f"{which} single-GPU dg_add_function_sdf tables={int(tables)}: {min(ts)*1e3:.1f} ms", flush=True)
