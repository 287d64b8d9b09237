"""diagnostic (1 GPU): sha256 of the K1 node-loop output on the bench workloads, to compare two builds of the library bit for bit at full
size (run once per build with DISCREGRID_B200_LIB set).  usage: k1_out_hash.py [bunny|torus|target ...]"""
import ctypes as C, hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import discregrid_b200 as dg
from discregrid_b200 import _capi as capi
import bench
for which in (sys.argv[1:] or ["bunny", "torus", "target"]):
    if which == "bunny": mesh, res = bench.workload_mesh(dg, "bunny")[0], 128
    elif which == "torus": mesh, res = dg.bumpy_torus(*bench.WORKLOAD["torus"]), 128
    else: mesh, res = dg.bumpy_torus(), 256
    md = dg.TriangleMeshDistance(mesh)
    mn, mx = dg.generate_sdf_domain(mesh.vertices); desc = dg.grid_desc(mn, mx, [res] * 3)
    n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, 0, n, C.c_void_p(out.data_ptr()), sp)); torch.cuda.synchronize()
    e0.record(); capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, 0, n, C.c_void_p(out.data_ptr()), sp)); e1.record(); e1.synchronize()
    h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"{which} {res}^3 {len(mesh.faces)} tris {n} nodes: {e0.elapsed_time(e1):.2f} ms sha {h}")
