import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import discregrid_b200 as dg
from discregrid_b200 import _capi as capi
import bench
mesh = dg.bumpy_torus(*bench.WORKLOAD["torus"]); md = dg.TriangleMeshDistance(mesh)
mn, mx = dg.generate_sdf_domain(mesh.vertices); desc = dg.grid_desc(mn, mx, [128] * 3)
n = 14926977
full = torch.empty(n, dtype=torch.float64, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(b, e):
    capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, b, e, C.c_void_p(full.data_ptr() + 8 * b), sp))
def t(b, e, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(b, e); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
run(0, n); torch.cuda.synchronize()
for size in (64, 2048, 16641, 33282, 66564, 266256, 1000000):
    for off in (0, 1000000, 5000000, 9000000, 13000000):
        print(f"size {size:8d} off {off:9d}: {t(off, off + size):7.3f} ms")
