#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Discregrid hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): SDF grid nodes/s of CubicLagrangeDiscreteGrid::addFunction with the GenerateSDF functor,
on configs[1] "Stanford-bunny-class mesh (~70k triangles), 128^3 grid, fp64" -- synthetic closed mesh of that size
(BASELINE.json north_star: "synthetic meshes/grids of the named shape"; /root/reference does not exist on the GPU box).
A "step" = one full pass of the node loop over the 14,926,977 nodes.  N > 1: the node range is dealt in chunks to the
ranks (strong scaling), followed by the all-gather of the coefficient array (discregrid_b200/distributed.py).
The JSON line also carries the second half of the metric, interpolate()+gradient Mqueries/s (config 4: 10 M uniform
random queries on a 256^3 SDF), under "interpolate".

--impl reference: the reference's own CPU path (oracle/_ref = its unmodified TriangleMeshDistance.h compiled here, else
the oracle port) on the host cores, same workload, each step a bounded sample of the node loop.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"mesh": "synthetic bumpy torus, 186x187 quads = 69,564 triangles / 34,782 vertices (bunny-class: bunny.obj has 69,630)",
            "resolution": [128, 128, 128], "torus": (186, 187, 1.0, 0.4, 0.05, 7, 5)}
INTERP = {"resolution": [256, 256, 256], "queries": 10_000_000, "seed": 0x5EED}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--resolution", type=int, default=0, help="override the SDF grid resolution (diagnostics)")
    ap.add_argument("--interp-resolution", type=int, default=0, help="override the interpolate grid resolution (diagnostics)")
    ap.add_argument("--target-resolution", type=int, default=256, help="grid resolution of the target-config leg (diagnostics / rehearsal; the config is 256)")
    ap.add_argument("--real-resolution", type=int, default=0, help="override the grid resolution of the reference-mesh leg (diagnostics / rehearsal)")
    ap.add_argument("--no-interp", action="store_true", help="skip the interpolate half (diagnostics / profiling)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (diagnostics / profiling)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg (diagnostics / profiling)")
    ap.add_argument("--no-target", action="store_true", help="skip the 256^3 / 100k-triangle target-config leg")
    ap.add_argument("--sharding", default="auto", choices=["auto", "slab", "chunks", "interleaved"],
                    help="N>1: round-robin node-id chunks, whole-plane slabs, or plane pairs dealt round-robin (one launch + one all-gather + unpack); "
                         "auto = interleaved from 4 ranks up (measured at N=8: equal at 128^3, 58.2 vs 64.2 ms at 256^3), chunks below")
    ap.add_argument("--no-real", action="store_true", help="skip the leg on the reference meshes staged under oracle/_ref/resources")
    ap.add_argument("--no-density", action="store_true", help="skip the density-map (K3) leg (diagnostics / profiling)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    args = ap.parse_args()
    if args.sharding == "auto":
        args.sharding = "interleaved" if int(os.environ.get("WORLD_SIZE", "1")) >= 4 else "chunks"
    return args


# ------------------------------------------------------------------------------------------------ helpers
def scratch_dir():
    """for the few hundred MB the CPU legs exchange with the reference class through files: the repo's own volume (build/ is git-ignored);
    /tmp can be a slow copy-on-write layer"""
    d = os.path.join(ROOT, "build")
    try:
        os.makedirs(d, exist_ok=True)
        return d
    except OSError:
        return "/tmp"


def reduce_field_leg(capi, desc, values, cells, lo, hi, with_reference, tmp_dir=None):
    """reduceField(field, lo <= v <= hi) (cmd/generate_density_map/main.cpp:141-144) on a sampled field -- host code on
    both sides: dg_reduce_field (index passes, multithreaded) vs the reference class's own reduceField (oracle/_ref, when built)."""
    if tmp_dir is None:
        tmp_dir = scratch_dir()
    keep = np.ascontiguousarray((lo <= values) & (values <= hi) & (values != np.finfo(np.float64).max), np.uint8)
    n_grid_cells = int(desc.resolution[0]) * int(desc.resolution[1]) * int(desc.resolution[2])
    best, out = None, None
    for _ in range(3):
        nodes, cc = values.copy(), cells.copy()
        cmap = np.empty(n_grid_cells, np.uint32); n1, n2 = C.c_uint64(), C.c_uint64(); tm = np.zeros(5)
        t0 = time.perf_counter()
        capi.check(capi.lib.dg_reduce_field(C.byref(desc), capi.ptr(nodes, capi.F64P), len(nodes), keep.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            capi.ptr(cc, capi.U32P), len(cc), capi.ptr(cmap, capi.U32P), 0, C.byref(n1), C.byref(n2), capi.ptr(tm, capi.F64P)))
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, out = dt, (nodes[:n1.value], cc[:n2.value], cmap, tm.copy())
    leg = {"what": "reduceField of the density field with the tool's predicate 0 <= v <= 3 rho0: host index passes of dg_reduce_field",
           "nodes_in": int(len(values)), "nodes_out": int(len(out[0])), "cells_in": int(len(cells)), "cells_out": int(len(out[1])),
           "ms": best * 1e3, "ms_cells_nodes_sort_write": [float(t) for t in out[3][:4]], "morton_keys_tied": bool(out[3][4])}
    if with_reference:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_api import REF_GRID_SO, RefGrid
        if os.path.exists(REF_GRID_SO):
            import struct
            src, dst = os.path.join(tmp_dir, f"dg_reduce_in_{os.getpid()}.cdf"), os.path.join(tmp_dir, f"dg_reduce_out_{os.getpid()}.cdf")
            try:
                with open(src, "wb") as f:                              # the reference's one-field file layout (:678-719)
                    f.write(struct.pack("<3d", *desc.domain_min)); f.write(struct.pack("<3d", *desc.domain_max)); f.write(struct.pack("<3I", *desc.resolution))
                    f.write(struct.pack("<3d", *desc.cell_size)); f.write(struct.pack("<3d", *desc.inv_cell_size)); f.write(struct.pack("<QQ", n_grid_cells, 1))
                    for arr in (values, cells, np.arange(len(cells), dtype=np.uint32)):
                        f.write(struct.pack("<QQ", 1, len(arr))); f.write(np.ascontiguousarray(arr).tobytes())
                ref = RefGrid(src); t_ref = ref.reduce_window(0, lo, hi); ref.save(dst); ref.close()
                raw = np.fromfile(dst, np.uint8)
                off = 24 * 4 + 12 + 16                                   # header: 4 x 3 doubles, 3 uint32, n_cells, n_fields
                def nested(dtype, width):
                    nonlocal off
                    n = int(raw[off + 8:off + 16].view(np.uint64)[0]); off += 16
                    a = raw[off:off + n * width * np.dtype(dtype).itemsize].view(dtype); off += a.nbytes
                    return a.reshape(n, width) if width > 1 else a
                rn, rc, rm = nested(np.float64, 1), nested(np.uint32, 32), nested(np.uint32, 1)
                same = bool(np.array_equal(rn.view(np.uint64), out[0].view(np.uint64)) and np.array_equal(rc, out[1]) and np.array_equal(rm, out[2]))
                leg["reference"] = {"ms": t_ref * 1e3, "kind": "reference", "impl": "CubicLagrangeDiscreteGrid::reduceField of oracle/_ref (serial, one std::set per node)",
                                    "identical_nodes_cells_cell_map": same, "speedup": t_ref / best}
            finally:
                for p_ in (src, dst):
                    if os.path.exists(p_):
                        os.remove(p_)
    return leg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


def splitmix_points(n, seed, lo, hi):
    """u = (splitmix64(seed, counter = 3q + d) >> 11) * 2^-53, x_d = lo_d + u * (hi_d - lo_d)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (np.arange(3 * n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return lo + u.reshape(n, 3) * (hi - lo)


def host_cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample_rate(mesh, mn, mx, res, seconds, steps=1, warmup=0):
    """Times the reference's CPU path on a strided sample of the node loop.  Returns (nodes/s list per step, info)."""
    # all host threads this process may use -- set before the OpenMP runtime starts, because torchrun exports OMP_NUM_THREADS=1
    os.environ["OMP_NUM_THREADS"] = os.environ.get("DG_CPU_THREADS") or str(len(os.sched_getaffinity(0)))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_api import Oracle, RefMesh, have_ref
    orc = Oracle()
    gd, r = orc.grid_desc(mn, mx, res)
    n_nodes = orc.num_nodes(r)
    threads = orc.max_threads()
    kind = "reference" if have_ref() else "port"
    m = RefMesh(mesh.vertices, mesh.faces) if kind == "reference" else orc.mesh(mesh.vertices, mesh.faces)

    def run(ids):
        # node positions by the oracle's indexToNodePosition (untimed), then the node-loop body on all host threads
        x = _positions(orc, gd, r, ids)
        t0 = time.perf_counter()
        if kind == "reference":
            m.sample_points(x, nthreads=threads)
        else:
            m.distance(x)
        return time.perf_counter() - t0

    # calibrate on 20k strided nodes, then size the sample for ~`seconds`
    probe = np.linspace(0, n_nodes - 1, 20_000).astype(np.int64)
    dt = run(probe)
    n_sample = int(min(n_nodes, max(20_000, seconds / max(dt, 1e-6) * len(probe))))
    ids = np.linspace(0, n_nodes - 1, n_sample).astype(np.int64)
    rates = []
    for it in range(warmup + steps):
        dt = run(ids)
        if it >= warmup:
            rates.append(n_sample / dt)
    model, ncpu = host_cpu_info()
    info = {"kind": kind, "cores": threads, "cpu_model": model, "logical_cpus": ncpu,
            "sample": f"{n_sample} of {n_nodes} nodes, evenly strided over the node index space, OpenMP schedule(static), "
                      f"{'reference TriangleMeshDistance.h (oracle/_ref)' if kind == 'reference' else 'oracle port'}"}
    return rates, info


def _positions(orc, gd, r, ids):
    """positions of arbitrary node ids via the oracle's indexToNodePosition"""
    return orc.node_positions_at(gd, r, ids)


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    res = [args.resolution] * 3 if args.resolution else WORKLOAD["resolution"]

    import discregrid_b200 as dg           # fails loudly if the CUDA library is not built
    from discregrid_b200 import _capi as capi
    from discregrid_b200.distributed import make_sharding, allgather_rows, ShardedSdfSampler, SlabSdfSampler, InterleavedSdfSampler

    mesh = dg.bumpy_torus(*WORKLOAD["torus"])
    mn, mx = dg.generate_sdf_domain(mesh.vertices)
    desc = dg.grid_desc(mn, mx, res)
    n_nodes = C.c_uint64()
    capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n_nodes)))
    n_nodes = n_nodes.value
    config = {"workload": f"GenerateSDF addFunction: {WORKLOAD['mesh']}; {res[0]}x{res[1]}x{res[2]} grid = {n_nodes} nodes; "
                          "GenerateSDF-padded domain; fp64 bit-exact with the reference",
              "mesh_triangles": int(mesh.nFaces()), "grid": res, "nodes": n_nodes,
              "l2": "flushed between timed iterations (256 MiB write)", "parallelism": f"x{world}" + ("" if world == 1 else (", slabs of whole plane pairs of the four node arrays: one launch per rank + one NCCL all-gather per node array" if args.sharding == "slab" else ", plane pairs dealt round-robin: one launch per rank + ONE NCCL all-gather + unpack kernel" if args.sharding == "interleaved" else ", 2 round-robin node-id chunks per rank on 2 streams + one in-place NCCL all-gather per row"))}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        rates, info = cpu_sample_rate(mesh, mn, mx, res, args.cpu_seconds / 2, steps=args.steps, warmup=min(args.warmup, 1))
        v = float(np.mean(rates))
        n_sample = int(info["sample"].split()[0])
        line = {"impl": "reference", "metric": "SDF grid nodes/sec (addFunction)", "value": v, "unit": "nodes/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_sample / v, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": dict(info, value=v, unit="nodes/s"),
                "e2e": {"value": v, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm (GPU)
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    capi.check(capi.lib.dg_set_device(local_rank))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    md = dg.TriangleMeshDistance(mesh)
    mesh_info = md.info()
    sh = make_sharding(n_nodes, world)
    full = torch.empty(sh.padded, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    my_chunks = [(j, b, e) for (j, b, e) in sh.chunks_of(rank)]
    stream = torch.cuda.current_stream()

    def make_sampler(md_, desc_, n_):
        if args.sharding == "slab":
            return SlabSdfSampler(md_, desc_, rank, world)
        if args.sharding == "interleaved":
            return InterleavedSdfSampler(md_, desc_, rank, world)
        return ShardedSdfSampler(md_, desc_, make_sharding(n_, world), rank)

    if args.sharding != "chunks":
        sdf_sampler = make_sampler(md, desc, n_nodes)
        full = torch.empty(sdf_sampler.sh.padded, dtype=torch.float64, device=dev)
    else:
        sdf_sampler = ShardedSdfSampler(md, desc, sh, rank)

    def sdf_step():
        sdf_sampler.step(full)

    def same_as_single_launch(md_, desc_, n_, sharded):
        """every rank: the assembled array of the sharded step vs ONE dg_sample_sdf_device launch over all nodes on this GPU (bit-for-bit)"""
        if world == 1:
            return None
        single = torch.empty(n_, dtype=torch.float64, device=dev)
        capi.check(capi.lib.dg_sample_sdf_device(md_.handle, C.byref(desc_), 1.0, 0, n_, C.c_void_p(single.data_ptr()), C.c_void_p(stream.cuda_stream)))
        ok = bool(torch.equal(single.view(torch.int64), sharded[:n_].view(torch.int64)))
        del single
        return bool(max_over_ranks(0.0 if ok else 1.0) == 0.0)

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        times = []
        barrier()
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)                                  # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step_fn(); e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        barrier()
        wall = time.perf_counter() - wall0
        return [max_over_ranks(t) for t in times], wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = dg.kernel_launch_count()
    sdf_ms, sdf_wall = timed(sdf_step, args.steps, args.warmup)
    launches = (dg.kernel_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    sharded_ok = same_as_single_launch(md, desc, n_nodes, full)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = float(np.mean(sdf_ms))
    value = n_nodes / (ms_step * 1e-3)

    # kernel-only time of K1 on this rank (no collective): what the roofline object refers to
    def k1_only():
        sdf_sampler.launch(full)
    k1_ms, _ = timed(k1_only, max(3, args.steps // 2), 1)
    k1_ms = float(np.mean(k1_ms))
    if args.sharding == "slab":
        my_nodes = sum(e - b for (b, e) in sdf_sampler.sh.ranges[rank]); n_launch = 1
    elif args.sharding == "interleaved":
        my_nodes = n_nodes // world; n_launch = 1
    else:
        my_nodes = sum(e - b for (_j, b, e) in my_chunks); n_launch = sum(1 for (_j, b, e) in my_chunks if e > b)
    peaks, peak_src = measured_peaks()
    k1_alg_bytes = 8.0 * my_nodes + mesh_info["device_bytes"]          # 8 B/node written + mesh records read once
    k1_gbs = k1_alg_bytes / (k1_ms * 1e-3) / 1e9
    roofline = {"kernel": "sdf_sample_nodes_kernel (K1)", "bound": "hbm", "achieved": k1_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": k1_gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                "traffic_ncu_other_config": {"bytes_per_launch": 27.4e6, "capture": "profiles/r1d_ncu_summary.csv: ncu --set full at 64^3 (the bench runs 128^3): dram read "
                                                                                    "27.0 MB = the mesh records once, 0.4 MB written back inside the launch"},
                "algorithmic_bytes_per_launch": k1_alg_bytes / max(1, n_launch), "launches_per_step": n_launch,
                "avg_launch_ms": k1_ms / max(1, n_launch),
                "note": "K1 is NOT HBM-bound: BVH + triangle records are L2-resident and compulsory HBM traffic is 8 B/node "
                        "(SURVEY 8d); it is bound by divergent fp64 ALU work and L1/L2 latency -- grade it on nodes/s"}

    # ---------------------------------------------------------------- e2e: C-ABI with HOST buffers
    e2e = None
    if not args.no_e2e:
        if world == 1:
            my_ranges = [(0, n_nodes)]
        elif args.sharding == "slab":
            my_ranges = [(b, e) for (b, e) in sdf_sampler.sh.ranges[rank] if e > b]
        else:                                                  # the host API takes node-id ranges: chunks (also in interleaved mode)
            my_ranges = [(b, e) for (_j, b, e) in my_chunks if e > b]
        outs = [np.empty(e - b) for (b, e) in my_ranges]

        def e2e_step():
            for (b, e), o in zip(my_ranges, outs):
                capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(desc), 1.0, b, e, capi.ptr(o, capi.F64P)))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, args.steps // 2)
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / n_e2e)
        e2e = {"value": n_nodes / dt, "unit": "nodes/s", "h2d_bytes_per_step": C.sizeof(capi.GridDesc) * len(my_ranges),
               "d2h_bytes_per_step": 8 * n_nodes, "ms_per_step": dt * 1e3,
               "api": "dg_sample_sdf(mesh, grid, sign, l_begin, l_end, out_host): kernel + D2H of the coefficient array into a "
                      "pageable host buffer; the mesh/BVH was uploaded once by dg_mesh_create (as TriangleMeshDistance is built "
                      "once, outside addFunction's timer, in the reference)",
               "mesh_upload": {"host_build_ms": mesh_info["build_us"] / 1e3, "h2d_ms": mesh_info["upload_us"] / 1e3,
                               "h2d_bytes": mesh_info["device_bytes"]}}

    # ---------------------------------------------------------------- interpolate half of the metric (config 4)
    interp = None
    if not args.no_interp:
        ires = [args.interp_resolution] * 3 if args.interp_resolution else INTERP["resolution"]
        idesc = dg.grid_desc(mn, mx, ires)
        nn = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(idesc.resolution, C.byref(nn))); nn = nn.value
        coeffs = torch.empty(nn, dtype=torch.float64, device=dev)
        sp = C.c_void_p(stream.cuda_stream)
        t0 = time.perf_counter()
        capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(idesc), 1.0, 0, nn, C.c_void_p(coeffs.data_ptr()), sp))
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        fh = C.c_void_p()
        capi.check(capi.lib.dg_field_create_device(C.byref(idesc), C.c_void_p(coeffs.data_ptr()), nn, sp, C.byref(fh)))
        torch.cuda.synchronize()
        del coeffs
        nq = INTERP["queries"]
        q_lo = (nq * rank) // world; q_hi = (nq * (rank + 1)) // world
        xh = splitmix_points(nq, INTERP["seed"], mn, mx)[q_lo:q_hi]
        xh_t = torch.from_numpy(np.ascontiguousarray(xh)).pin_memory()
        xd = xh_t.to(dev)
        phi = torch.empty(q_hi - q_lo, dtype=torch.float64, device=dev)
        grad = torch.empty((q_hi - q_lo, 3), dtype=torch.float64, device=dev)

        def interp_step(with_grad=True):
            capi.check(capi.lib.dg_interpolate_batch_device(fh, C.c_void_p(xd.data_ptr()), q_hi - q_lo, C.c_void_p(phi.data_ptr()),
                                                            C.c_void_p(grad.data_ptr()) if with_grad else None, sp))
        ig_ms, _ = timed(lambda: interp_step(True), 20, 3)
        iv_ms, _ = timed(lambda: interp_step(False), 20, 3)
        ig_ms, iv_ms = float(np.mean(ig_ms)), float(np.mean(iv_ms))
        # the same points sorted by cell (z-major cell index) separate gather locality from arithmetic
        cell = ((xh - mn) / (mx - mn) * ires[0]).astype(np.int64).clip(0, ires[0] - 1)
        order = np.argsort((cell[:, 2] * ires[1] + cell[:, 1]) * ires[0] + cell[:, 0], kind="stable")
        xd_sorted = torch.from_numpy(np.ascontiguousarray(xh[order])).to(dev)
        xd_keep = xd
        xd = xd_sorted
        is_ms, _ = timed(lambda: interp_step(True), 20, 3)
        is_ms = float(np.mean(is_ms))
        xd = xd_keep
        alg = 312.0 * (q_hi - q_lo)
        gbs = alg / (ig_ms * 1e-3) / 1e9
        # e2e through the host API
        xq = np.ascontiguousarray(xh); ph = np.empty(len(xq)); gh = np.empty((len(xq), 3))
        for _ in range(2):
            capi.check(capi.lib.dg_interpolate_batch(fh, capi.ptr(xq, capi.F64P), len(xq), capi.ptr(ph, capi.F64P), capi.ptr(gh, capi.F64P)))
        barrier(); t0 = time.perf_counter()
        for _ in range(5):
            capi.check(capi.lib.dg_interpolate_batch(fh, capi.ptr(xq, capi.F64P), len(xq), capi.ptr(ph, capi.F64P), capi.ptr(gh, capi.F64P)))
        dt = max_over_ranks((time.perf_counter() - t0) / 5)
        interp = {"metric": "interpolate()+gradient Mqueries/s", "value": nq / (ig_ms * 1e-3) / 1e6, "unit": "Mqueries/s",
                  "value_only_mqps": nq / (iv_ms * 1e-3) / 1e6, "cell_sorted_mqps": nq / (is_ms * 1e-3) / 1e6, "ms_per_launch": ig_ms, "queries": nq,
                  "config": {"workload": f"10M splitmix64 uniform queries (seed 0x5EED) on the {ires[0]}^3 SDF of the same mesh "
                                         f"({nn} nodes; packed cell blocks {16 * ires[0] * ires[1] * ires[2] * 16 / 1e9:.2f} GB >> L2)",
                             "field_build_s": build_s},
                  "roofline": {"kernel": "interpolate_kernel<true> (K2)", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                               "algorithmic_bytes_per_query": 312,
                               "traffic_ncu_other_config": {"bytes_per_launch": 2.78e9, "algorithmic_bytes_per_launch": 3.12e9,
                                                            "capture": "profiles/r1e_k2k3_ncu_summary.csv: ncu --set full on the 128^3 field (the bench runs 256^3), "
                                                                       "dram read 2.46 GB + write 0.314 GB"}},
                  "e2e": {"value": nq / dt / 1e6, "unit": "Mqueries/s", "h2d_bytes_per_step": 24 * nq, "d2h_bytes_per_step": 32 * nq,
                          "api": "dg_interpolate_batch(field, x_host, n, phi_host, grad_host)"}}
        capi.lib.dg_field_destroy(fh)
        # CPU baseline of interpolate: the reference's own CubicLagrangeDiscreteGrid::interpolate (oracle/_ref/libdiscregrid_ref.so =
        # its unmodified sources against the Eigen stand-in) in the OpenMP loop of cmd/discrete_field_to_bitmap/main.cpp:118-135
        if rank == 0 and world == 1 and not args.no_cpu:
            os.environ["OMP_NUM_THREADS"] = os.environ.get("DG_CPU_THREADS") or str(len(os.sched_getaffinity(0)))
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle_api import Oracle, RefGrid, have_ref_grid
            cres = [min(128, r_) for r_ in ires]                 # a 256^3 .cdf is 3 GB on disk: the CPU arm reads the 128^3 field of the same mesh
            cgrid = dg.CubicLagrangeDiscreteGrid(mn, mx, cres)
            cgrid.addFunction(dg.MeshSignedDistance(md))
            nq_cpu = 2_000_000
            xc = np.ascontiguousarray(splitmix_points(nq_cpu, INTERP["seed"], mn, mx))
            if have_ref_grid():
                tmpf = os.path.join(scratch_dir(), f"_dg_bench_field_{os.getpid()}.cdf")
                cgrid.save(tmpf)
                rg = RefGrid(tmpf)
                rg.interpolate(0, xc[:100000], grad=True)
                t0 = time.perf_counter(); pr, gr = rg.interpolate(0, xc, grad=True); t_cpu = time.perf_counter() - t0
                kind = "reference"
                rg.close(); os.remove(tmpf)
            else:
                orc = Oracle(); gd_c, r_c = orc.grid_desc(mn, mx, cres)
                t0 = time.perf_counter(); pr, gr = orc.interpolate(gd_c, r_c, cgrid.m_nodes[0], xc, grad=True); t_cpu = time.perf_counter() - t0
                kind = "port"
            pg, gg = cgrid.interpolate(0, xc, gradient=True)
            interp["cpu_baseline"] = {"value": nq_cpu / t_cpu / 1e6, "unit": "Mqueries/s", "kind": kind, "cores": int(os.environ["OMP_NUM_THREADS"]),
                                      "sample": f"{nq_cpu} of the 10M queries on the {cres[0]}^3 field of the same mesh (OpenMP parallel for, value+gradient)",
                                      "bit_exact_vs_gpu": bool(np.array_equal(pr.view(np.uint64), pg.view(np.uint64)) and np.array_equal(gr.view(np.uint64), gg.view(np.uint64)))}
            del cgrid

    # ---------------------------------------------------------------- north-star target config: 256^3 grid, 100,000-triangle mesh
    target = None
    if not args.no_target:
        tmesh = dg.bumpy_torus()                               # 250 x 200 quads = exactly 100,000 triangles
        tmd = dg.TriangleMeshDistance(tmesh)
        tmn, tmx = dg.generate_sdf_domain(tmesh.vertices)
        tres = [args.target_resolution] * 3
        tdesc = dg.grid_desc(tmn, tmx, tres)
        tn = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(tdesc.resolution, C.byref(tn))); tn = tn.value
        tsampler = make_sampler(tmd, tdesc, tn)
        tfull = torch.empty(tsampler.sh.padded, dtype=torch.float64, device=dev)
        t_ms, _ = timed(lambda: tsampler.step(tfull), 3, 1)
        t_ms = float(np.mean(t_ms))
        target = {"workload": f"north_star target: {tres[0]}^3 grid ({tn} nodes), synthetic bumpy torus with exactly 100,000 triangles, "
                              "strong scaling, sharding as config.parallelism", "ms_per_step": t_ms, "value": tn / (t_ms * 1e-3), "unit": "nodes/s",
                  "n_gpus": world, "sharded_equals_single_launch": same_as_single_launch(tmd, tdesc, tn, tfull)}
        if rank == 0 and world == 1 and not args.no_cpu:
            rates, info = cpu_sample_rate(tmesh, tmn, tmx, tres, args.cpu_seconds)
            target["cpu_baseline"] = dict(info, value=float(np.mean(rates)), unit="nodes/s")
        del tfull, tsampler, tmd

    # ---------------------------------------------------------------- the reference's own meshes (configs 2/3/5), when staged
    real = None
    res_dir = os.path.join(ROOT, "oracle", "_ref", "resources")      # copied there by `make -C oracle ref`; travels with the repo
    if not args.no_real and os.path.isdir(res_dir):
        real = []
        for name, r3 in (("bunny.obj", 128), ("dragon.obj", 256), ("happy_buddha.obj", 256)):
            path = os.path.join(res_dir, name)
            if not os.path.exists(path):
                continue
            rmesh = dg.TriangleMesh(path)
            t0 = time.perf_counter(); rmd = dg.TriangleMeshDistance(rmesh); t_create = time.perf_counter() - t0
            rmn, rmx = dg.generate_sdf_domain(rmesh.vertices)
            r3 = args.real_resolution or r3
            rdesc = dg.grid_desc(rmn, rmx, [r3] * 3)
            rn = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(rdesc.resolution, C.byref(rn))); rn = rn.value
            rs = make_sampler(rmd, rdesc, rn)
            rfull = torch.empty(rs.sh.padded, dtype=torch.float64, device=dev)
            r_ms, _ = timed(lambda: rs.step(rfull), 2, 1)
            r_ms = float(np.mean(r_ms))
            entry = {"mesh": name, "triangles": int(rmesh.nFaces()), "grid": r3, "nodes": rn, "ms_per_step": r_ms, "value": rn / (r_ms * 1e-3),
                     "unit": "nodes/s", "mesh_create_s": t_create, "watertight_flags": rmd.info()["watertight_flags"]}
            if name == "dragon.obj" and world == 1 and not args.no_density:      # config 5: GenerateDensityMap on the dragon SDF, h = 0.1, rho0 = 1000
                fh = C.c_void_p(); sp = C.c_void_p(stream.cuda_stream)
                capi.check(capi.lib.dg_field_create_device(C.byref(rdesc), C.c_void_p(rfull.data_ptr()), rn, sp, C.byref(fh)))
                dens = torch.empty(rn, dtype=torch.float64, device=dev)
                k3_ms, _ = timed(lambda: capi.check(capi.lib.dg_density_map_device(fh, 0.1, 1000.0, 0, 0, rn, C.c_void_p(dens.data_ptr()), sp)), 1, 1)
                entry["density_map"] = {"h": 0.1, "rho0": 1000.0, "ms": float(k3_ms[0]), "value": rn / (k3_ms[0] * 1e-3), "unit": "nodes/s",
                                        "nodes_in_quadrature_branch": int(((dens > 0) & (dens < 1e300)).sum().item())}
                capi.lib.dg_field_destroy(fh); del dens
            real.append(entry)
            del rfull, rs, rmd

    # ---------------------------------------------------------------- density map (config 5 kernel), N = 1 only
    density = None
    if world == 1 and not args.no_density:
        sdf_sampler.launch(full); torch.cuda.synchronize()
        fh = C.c_void_p()
        sp = C.c_void_p(stream.cuda_stream)
        capi.check(capi.lib.dg_field_create_device(C.byref(desc), C.c_void_p(full.data_ptr()), n_nodes, sp, C.byref(fh)))
        dens = torch.empty(n_nodes, dtype=torch.float64, device=dev)
        h_dm = 0.1 * float(np.max(mx - mn)) / 2.5            # the reference default h = 0.1 is for a ~2.5-unit dragon
        def dm_step():
            capi.check(capi.lib.dg_density_map_device(fh, h_dm, 1000.0, 0, 0, n_nodes, C.c_void_p(dens.data_ptr()), sp))
        dm_ms, _ = timed(dm_step, 2, 1)
        dm_ms = float(np.mean(dm_ms))
        dens_h = dens.cpu().numpy()
        active = int(((dens_h > 0) & (dens_h < 1e300)).sum())
        density = {"metric": "GenerateDensityMap nodes/s (K3)", "value": n_nodes / (dm_ms * 1e-3), "unit": "nodes/s", "ms": dm_ms,
                   "config": {"workload": f"density_func + predicate over the {res[0]}^3 SDF above, h = {h_dm:.4f}, rho0 = 1000, 16^3 Gauss points",
                              "nodes_in_quadrature_branch": active}}
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle_api import Oracle
            orc = Oracle()
            gd, r = orc.grid_desc(mn, mx, res)
            coeff_h = full[:n_nodes].cpu().numpy()
            t_cpu, n_cpu, ok = 0.0, 0, True
            for k in range(8):                                 # eight 192-node windows spread over the node index space
                l0 = int((k + 0.5) * n_nodes / 8); l1 = min(l0 + 192, n_nodes)
                t0 = time.perf_counter()
                ref = orc.density_map(gd, r, coeff_h, h_dm, 1000.0, False, l0, l1)
                t_cpu += time.perf_counter() - t0; n_cpu += l1 - l0
                ok = ok and np.array_equal(ref.view(np.uint64), dens_h[l0:l1].view(np.uint64))
            density["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": "nodes/s", "kind": "port", "cores": orc.max_threads(),
                                       "sample": f"{n_cpu} nodes in 8 windows, oracle/dg_oracle.cpp (the reference tool needs Eigen)",
                                       "bit_exact_vs_gpu": bool(ok)}
        try:                                                    # N3: the step after the density map in GenerateDensityMap (host code)
            cells_h = np.empty((int(np.prod(np.array(res, np.uint64))), 32), np.uint32)
            capi.check(capi.lib.dg_build_cells(desc.resolution, 0, len(cells_h), capi.ptr(cells_h, capi.U32P)))
            density["reduce_field"] = reduce_field_leg(capi, desc, dens_h, cells_h, 0.0, 3000.0, with_reference=not args.no_cpu)
            del cells_h
        except Exception as ex:                                 # an auxiliary leg must not take the bench line down
            density["reduce_field"] = {"error": repr(ex)}
        capi.lib.dg_field_destroy(fh)
        del dens

    if world > 1 and not args.no_density:                      # N > 1: the same node function over node-id chunks + the SDF's all-gather
        from discregrid_b200.distributed import ShardedDensityMap
        sdf_sampler.step(full); torch.cuda.synchronize()
        fh, dens, dmap, setup_err = C.c_void_p(), None, None, None
        sp = C.c_void_p(stream.cuda_stream)
        try:                                                    # anything that can fail on ONE rank only happens before the collective part
            capi.check(capi.lib.dg_field_create_device(C.byref(desc), C.c_void_p(full.data_ptr()), n_nodes, sp, C.byref(fh)))
            dmap = ShardedDensityMap(fh, n_nodes, rank, world)
            dens = torch.empty(dmap.sh.padded, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
        except Exception as ex:
            setup_err = repr(ex)
        if max_over_ranks(0.0 if setup_err is None else 1.0) == 0.0:
            h_dm = 0.1 * float(np.max(mx - mn)) / 2.5
            dm_ms, _ = timed(lambda: dmap.step(dens, h_dm, 1000.0), 2, 1)
            dm_ms = float(np.mean(dm_ms))
            density = {"metric": "GenerateDensityMap nodes/s (K3)", "value": n_nodes / (dm_ms * 1e-3), "unit": "nodes/s", "ms": dm_ms, "n_gpus": world,
                       "config": {"workload": f"density_func + predicate over the {res[0]}^3 SDF above, h = {h_dm:.4f}, rho0 = 1000, 16^3 Gauss points; "
                                              f"8 round-robin node-id chunks per rank + one in-place all-gather per row"}}
        else:
            density = {"error": setup_err or "setup failed on another rank"}
        if fh:
            capi.lib.dg_field_destroy(fh)
        del dens

    # ---------------------------------------------------------------- CPU baseline, rank 0, N = 1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rates, info = cpu_sample_rate(mesh, mn, mx, res, args.cpu_seconds)
        cpu = dict(info, value=float(np.mean(rates)), unit="nodes/s")

    if rank == 0:
        line = {"metric": "SDF grid nodes/sec (addFunction)", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "interpolate": interp, "target_config": target, "reference_meshes": real, "density_map": density,
                "timing": {"per_step_ms": sdf_ms, "wall_s_timed_region": sdf_wall, "k1_only_ms_per_step": k1_ms},
                "sharded_equals_single_launch": sharded_ok}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
