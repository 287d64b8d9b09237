"""diagnostic: per-chunk K1 time for a given world size layout (1 GPU)"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import discregrid_b200 as dg
from discregrid_b200 import _capi as capi
from discregrid_b200.distributed import make_sharding
import bench
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mesh = dg.bumpy_torus(*bench.WORKLOAD["torus"]); md = dg.TriangleMeshDistance(mesh)
mn, mx = dg.generate_sdf_domain(mesh.vertices); desc = dg.grid_desc(mn, mx, [128] * 3)
n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
sh = make_sharding(n, world, rows)
full = torch.empty(sh.padded, dtype=torch.float64, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(b, e):
    capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, b, e, C.c_void_p(full.data_ptr() + 8 * b), sp))
run(0, n); torch.cuda.synchronize()
tot = {r: 0.0 for r in range(world)}
for r in range(world):
    for (j, b, e) in sh.chunks_of(r):
        if e <= b: continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(b, e); e1.record(); e1.synchronize()
        t = e0.elapsed_time(e1); tot[r] += t
        print(f"rank {r} row {j} [{b},{e}) {t:7.2f} ms  {(e-b)/t/1e3:7.1f} Mnodes/s")
print(tot)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(0, n); e1.record(); e1.synchronize(); print("full", e0.elapsed_time(e1))
# experiment: the chunks of each rank launched round-robin on S streams
for S in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(S)]
    for r in range(world):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams: st.wait_stream(torch.cuda.current_stream())
        for k, (j, b, e) in enumerate(sh.chunks_of(r)):
            if e <= b: continue
            st = streams[k % S]
            capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, b, e, C.c_void_p(full.data_ptr() + 8 * b), C.c_void_p(st.cuda_stream)))
        for st in streams: torch.cuda.current_stream().wait_stream(st)
        e1.record(); e1.synchronize()
        print(f"streams {S} rank {r}: {e0.elapsed_time(e1):.2f} ms")
