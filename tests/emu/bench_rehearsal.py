"""tests/emu/bench_rehearsal.py -- TEST INFRASTRUCTURE: runs the repo's bench.py END TO END without a GPU, to catch Python-level breakage
in any of its legs before the driver runs it on a B200.  torch's CUDA surface is replaced by host stand-ins (tensors on the CPU, events =
perf_counter, NCCL -> gloo), the library is build/bin/libdgemu.so (every kernel emulated), and the workloads are shrunk to toy sizes.
The printed numbers mean nothing; only "every leg ran and the JSON line has its keys" is checked (tests/test_gpu_rehearsal.py).
usage: DISCREGRID_B200_LIB=.../libdgemu.so python tests/emu/bench_rehearsal.py [bench.py arguments]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch                         # noqa: E402
import torch.distributed as dist     # noqa: E402

_real_device = torch.device


def _device(kind, index=None):
    return _real_device("cpu") if str(kind).startswith("cuda") else (_real_device(kind) if index is None else _real_device(kind, index))


class _Stream:
    cuda_stream = 0

    def wait_stream(self, _other):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


_real_init = dist.init_process_group


def _init_process_group(backend=None, **kw):
    kw.pop("device_id", None)
    return _real_init("gloo", **kw)


torch.device = _device
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *_a, **_k: None
torch.cuda.synchronize = lambda *_a, **_k: None
torch.cuda.current_stream = lambda *_a, **_k: _Stream()
torch.cuda.Stream = _Stream
torch.cuda.Event = _Event
torch.Tensor.pin_memory = lambda self, *_a, **_k: self
dist.init_process_group = _init_process_group

import bench                         # noqa: E402

bench.INTERP["queries"] = 3000                                   # instead of 10 M
bench.WORKLOAD["torus"] = (24, 20, 1.0, 0.4, 0.05, 7, 5)        # 960 triangles instead of 69,564
bench.WORKLOAD["source"] = "torus"                              # not the 69,630-triangle bunny.obj

if __name__ == "__main__":
    bench.main()
