#!/bin/bash
# final 8-GPU record: the multi-GPU tests and the default bench line under torchrun N = 8 (as the driver's scaling run launches it)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi_capi.py tests/test_gpu_multi.py -m gpu -q > $O/r2r_pytest.txt 2>&1; tail -2 $O/r2r_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2r_bench_n8.out 2> $O/r2r_bench_n8.err
python - <<'PY'
import json
lines=[l for l in open("gpurun_out/r2r_bench_n8.out").read().splitlines() if l.startswith('{"metric"')]
print("stdout lines:", len(open("gpurun_out/r2r_bench_n8.out").read().splitlines()))
d=json.loads(lines[-1])
print("N=8", round(d["ms_per_step"],2), "ms k1_only", round(d["timing"]["k1_only_ms_per_step"],2), "e2e", d["e2e"].get("ms_per_step"), "c-abi", (d["e2e"].get("single_process_c_abi") or {}).get("ms_per_step"), "target", d["target_config"]["ms_per_step"], d["sharded_equals_single_launch"], "interp", d["interpolate"]["value"], "K3", d["density_map"].get("ms"))
for m in d.get("reference_meshes") or []: print(m["mesh"], m["ms_per_step"])
PY
