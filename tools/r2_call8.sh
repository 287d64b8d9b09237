#!/bin/bash
# 8-GPU call: the C-ABI multi-GPU group on 8 devices (tests), GenerateSDF --gpus 8, bench.py under torchrun with N = 8
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l > $O/r2h_gpus.txt; cat /sys/fs/cgroup/cpu.max >> $O/r2h_gpus.txt 2>/dev/null
timeout 600 python -m pytest tests/test_gpu_multi_capi.py tests/test_gpu_cpp_facade.py tests/test_gpu_multi.py -m gpu -q > $O/r2h_pytest.txt 2>&1; tail -3 $O/r2h_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --no-real > $O/r2h_bench_n8.json 2> $O/r2h_bench_n8.err
tail -c 300 $O/r2h_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_bench_n8.json").read().strip().splitlines()[-1])
print("N=8", round(d["ms_per_step"],2), "ms k1_only", round(d["timing"]["k1_only_ms_per_step"],2), "e2e", d["e2e"].get("ms_per_step"), "c-abi multi", d["e2e"].get("single_process_c_abi"), "target", d["target_config"]["ms_per_step"], d["sharded_equals_single_launch"])
PY
