"""Rehearsal of the `-m gpu` suite on the CPU.  build/bin/libdgemu.so is the whole product library with every kernel run by tests/emu
(lanes as fibers) and the CUDA runtime stubbed on host memory.  With DISCREGRID_B200_LIB pointing at it -- and LD_PRELOAD, so that the
rebuilt C++ tools resolve the C-ABI there too -- the GPU tests that need neither torch nor a full-size grid run unchanged: kernels,
launchers, C-ABI layer, Python mirror, C++ facade and tools against the same goldens as on the B200.  A failure here means the GPU
run at the end of the round would fail for a reason that has nothing to do with the GPU.  (Not a CPU fallback: test infrastructure.)"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EMU = os.path.join(ROOT, "build", "bin", "libdgemu.so")
FILES = ["tests/test_gpu_k1_sdf.py", "tests/test_gpu_k2_interp.py", "tests/test_gpu_k3_density.py", "tests/test_gpu_reference_tools.py",
         "tests/test_gpu_cpp_facade.py"]
# left to the real GPU: device tensors through torch, the full-size configuration, the 69k-855k-triangle meshes (minutes when emulated)
SKIP = "not full_size and not slab_parts and not interleaved_parts and not reference_meshes_vs_oracle"


def test_gpu_suite_passes_on_the_emulated_library():
    if not os.path.exists(EMU):
        pytest.skip("build/bin/libdgemu.so not built (make cpp)")
    env = dict(os.environ, DISCREGRID_B200_LIB=EMU, LD_PRELOAD=EMU)
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-k", SKIP, "-p", "no:cacheprovider"] + FILES,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


def test_multi_gpu_worker_on_two_gloo_ranks(tmp_path):
    """the worker script of tests/test_gpu_multi.py (node-id chunks, slabs, interleaved plane pairs, sharded density map, each compared
    with a single launch) on two gloo ranks with host tensors against the emulated library: the N > 1 plumbing end to end"""
    if not os.path.exists(EMU):
        pytest.skip("build/bin/libdgemu.so not built (make cpp)")
    from test_gpu_multi import WORKER
    script = tmp_path / "w.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    env = dict(os.environ, DISCREGRID_B200_LIB=EMU, DG_REHEARSAL="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", str(script)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "MULTI_OK 2" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
