"""Multi-GPU parity (GPU box with >= 2 devices; skipped otherwise): the sharded node loop + NCCL all-gather of the
coefficient array reproduces the single-GPU result bit for bit on every rank (the kernel is deterministic per node)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ROOT)
import discregrid_b200 as dg
from discregrid_b200 import _capi as capi
from discregrid_b200.distributed import make_sharding, allgather_rows
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
# DG_REHEARSAL=1 (tests/test_gpu_rehearsal.py): the same script on host tensors over gloo against the emulated library -- a smaller grid
REHEARSAL = os.environ.get("DG_REHEARSAL") == "1"
DEV = "cpu" if REHEARSAL else "cuda"
def sync():
    if not REHEARSAL: torch.cuda.synchronize()
if REHEARSAL:
    dist.init_process_group("gloo")
else:
    torch.cuda.set_device(lr); capi.check(capi.lib.dg_set_device(lr))
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mesh = dg.bumpy_torus(60, 50) if not REHEARSAL else dg.bumpy_torus(16, 12)
md = dg.TriangleMeshDistance(mesh)
mn, mx = dg.generate_sdf_domain(mesh.vertices)
desc = dg.grid_desc(mn, mx, (40, 36, 20) if not REHEARSAL else (9, 8, 5))
n = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(n))); n = n.value
sh = make_sharding(n, world, rows=5, align=256)
full = torch.full((sh.padded,), float("nan"), dtype=torch.float64, device=DEV)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream) if not REHEARSAL else C.c_void_p(0)
for (_j, b, e) in sh.chunks_of(rank):
    if e > b:
        capi.check(capi.lib.dg_sample_sdf_device(md.handle, C.byref(desc), 1.0, b, e, C.c_void_p(full.data_ptr() + 8 * b), sp))
allgather_rows(full, sh)
sync()
single = np.empty(n)
capi.check(capi.lib.dg_sample_sdf(md.handle, C.byref(desc), 1.0, 0, n, capi.ptr(single, capi.F64P)))
got = full[:n].cpu().numpy()
assert np.array_equal(got.view(np.uint64), single.view(np.uint64)), f"rank {rank}: sharded != single-GPU"
# slab sharding: one launch per rank, whole plane pairs, four all-gathers
from discregrid_b200.distributed import SlabSdfSampler
ss = SlabSdfSampler(md, desc, rank, world)
full2 = torch.full((ss.sh.padded,), float("nan"), dtype=torch.float64, device=DEV)
ss.step(full2)
sync()
got2 = full2[:n].cpu().numpy()
assert np.array_equal(got2.view(np.uint64), single.view(np.uint64)), f"rank {rank}: slab-sharded != single-GPU"
# interleaved slab sharding: plane pairs dealt round-robin, one all-gather, unpack
from discregrid_b200.distributed import InterleavedSdfSampler
it = InterleavedSdfSampler(md, desc, rank, world)
full3 = torch.full((n,), float("nan"), dtype=torch.float64, device=DEV)
it.step(full3)
sync()
got3 = full3.cpu().numpy()
assert np.array_equal(got3.view(np.uint64), single.view(np.uint64)), f"rank {rank}: interleaved != single-GPU"
# density map (K3) over node-id chunks of the replicated SDF field + the same all-gather
from discregrid_b200.distributed import ShardedDensityMap
fh = C.c_void_p()
capi.check(capi.lib.dg_field_create_device(C.byref(desc), C.c_void_p(full3.data_ptr()), n, sp, C.byref(fh)))
dm = ShardedDensityMap(fh, n, rank, world, rows=3)
dens = torch.full((dm.sh.padded,), float("nan"), dtype=torch.float64, device=DEV)
H = 0.08 if not REHEARSAL else 0.2
dm.step(dens, H, 1000.0)
one = torch.empty(n, dtype=torch.float64, device=DEV)
capi.check(capi.lib.dg_density_map_device(fh, H, 1000.0, 0, 0, n, C.c_void_p(one.data_ptr()), sp))
sync()
assert torch.equal(dens[:n].view(torch.int64), one.view(torch.int64)), f"rank {rank}: sharded density map != single launch"
assert bool(((one > 0) & (one < 1e300)).any()) and bool((one == 0).any())
capi.lib.dg_field_destroy(fh)
dist.barrier()
if rank == 0: print("MULTI_OK", world)
dist.destroy_process_group()
'''


def test_sharded_equals_single_gpu(dg, tmp_path):
    n = dg.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    script = tmp_path / "w.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", "29617", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
