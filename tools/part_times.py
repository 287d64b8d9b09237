"""diagnostic (1 GPU): per-part K1 kernel time of the three multi-GPU shardings for a simulated world size -- max over parts is what a
real N-GPU step would wait for (collective excluded).  usage: part_times.py [world] [resolution] [mesh: torus|target|bunny]"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import discregrid_b200 as dg
from discregrid_b200 import _capi as capi
from discregrid_b200.distributed import make_sharding, SlabSharding
import bench
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
res = int(sys.argv[2]) if len(sys.argv) > 2 else 128
which = sys.argv[3] if len(sys.argv) > 3 else "torus"
mesh = bench.workload_mesh(dg, "bunny")[0] if which == "bunny" else (dg.bumpy_torus(*bench.WORKLOAD["torus"]) if which == "torus" else dg.bumpy_torus())
md = dg.TriangleMeshDistance(mesh)
mn, mx = dg.generate_sdf_domain(mesh.vertices); desc = dg.grid_desc(mn, mx, [res] * 3)
n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
full = torch.empty(n + 4096, dtype=torch.float64, device="cuda")

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if os.environ.get("DG_FLUSH_L2") == "1" else None

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None: flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

t_full = timed(lambda: capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, 0, n, C.c_void_p(full.data_ptr()), sp)))
print(f"world {world} res {res} mesh {which} tris {len(mesh.faces)}: single launch {t_full:.2f} ms -> ideal {t_full / world:.2f} ms per part")

# chunks (rows = 2), both chunks of a rank on two streams as the sampler does
sh = make_sharding(n, world)
fullc = torch.empty(sh.padded, dtype=torch.float64, device="cuda")
streams = [torch.cuda.Stream() for _ in range(2)]
def chunks_of(r):
    def f():
        cur = torch.cuda.current_stream()
        for st in streams: st.wait_stream(cur)
        for k, (j, b, e) in enumerate(sh.chunks_of(r)):
            if e <= b: continue
            capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, b, e, C.c_void_p(fullc.data_ptr() + 8 * b), C.c_void_p(streams[k % 2].cuda_stream)))
        for st in streams: cur.wait_stream(st)
    return f
tc = [timed(chunks_of(r)) for r in range(world)]
print("chunks      ", " ".join(f"{t:6.2f}" for t in tc), f"| max {max(tc):.2f} mean {np.mean(tc):.2f}")

# slabs
ss = SlabSharding(desc, world)
fulls = torch.empty(ss.padded, dtype=torch.float64, device="cuda")
def slab_of(r):
    return lambda: capi.check(capi.lib.dg_sample_sdf_slab_device(md.handle, C.byref(desc), 1.0, r, world, C.c_void_p(fulls.data_ptr()), sp))
try:
    ts_ = [timed(slab_of(r)) for r in range(world)]
    print("slab        ", " ".join(f"{t:6.2f}" for t in ts_), f"| max {max(ts_):.2f} mean {np.mean(ts_):.2f}")
except Exception as ex:
    print("slab: skipped", ex)

# interleaved
se = C.c_uint64(); capi.check(capi.lib.dg_interleaved_slot_elems(C.byref(desc), world, C.byref(se))); se = se.value
slots = torch.empty(world * se, dtype=torch.float64, device="cuda")
def inter_of(r):
    return lambda: capi.check(capi.lib.dg_sample_sdf_interleaved_device(md.handle, C.byref(desc), 1.0, r, world, C.c_void_p(slots.data_ptr() + 8 * r * se), sp))
ti = [timed(inter_of(r)) for r in range(world)]
print("interleaved ", " ".join(f"{t:6.2f}" for t in ti), f"| max {max(ti):.2f} mean {np.mean(ti):.2f}")
# interleaved over 2 * world parts, two launches per rank on two streams (InterleavedSdfSampler(splits=2))
if 2 * world <= 16:
    se2 = C.c_uint64(); capi.check(capi.lib.dg_interleaved_slot_elems(C.byref(desc), 2 * world, C.byref(se2))); se2 = se2.value
    slots2 = torch.empty(2 * world * se2, dtype=torch.float64, device="cuda")
    def inter2_of(r):
        def f():
            cur = torch.cuda.current_stream()
            for st in streams: st.wait_stream(cur)
            for k in range(2):
                part = 2 * r + k
                capi.check(capi.lib.dg_sample_sdf_interleaved_device(md.handle, C.byref(desc), 1.0, part, 2 * world, C.c_void_p(slots2.data_ptr() + 8 * part * se2),
                                                                     C.c_void_p(streams[k].cuda_stream)))
            for st in streams: cur.wait_stream(st)
        return f
    t2 = [timed(inter2_of(r)) for r in range(world)]
    print("interleaved2", " ".join(f"{t:6.2f}" for t in t2), f"| max {max(t2):.2f} mean {np.mean(t2):.2f}")
tu = timed(lambda: capi.check(capi.lib.dg_interleaved_unpack_device(C.byref(desc), world, C.c_void_p(slots.data_ptr()), C.c_void_p(full.data_ptr()), sp)))
print(f"unpack {tu:.3f} ms; slot {se} elems ({8 * se * world / 1e6:.1f} MB gathered vs {8 * n / 1e6:.1f} MB of nodes)")
