#!/bin/bash
# the final build on 2 GPUs: the whole -m gpu suite (1-GPU tests + the 2-GPU ones) and the default bench line under torchrun N = 2
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2x_pytest_2gpu.txt 2>&1; tail -2 $O/r2x_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-real > $O/r2x_bench_n2.out 2> $O/r2x_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2x_bench_n2.out").read().splitlines() if l.startswith('{"metric"')][-1])
print("N=2", round(d["ms_per_step"],2), "ms k1_only", round(d["timing"]["k1_only_ms_per_step"],2), "e2e", d["e2e"].get("ms_per_step"), "c-abi", (d["e2e"].get("single_process_c_abi") or {}).get("ms_per_step"), "target", d["target_config"]["ms_per_step"], d["sharded_equals_single_launch"], "interp", d["interpolate"]["value"], "K3", d["density_map"].get("ms"))
PY
