#!/bin/bash
# 2-GPU call: the multi-GPU C-ABI (dg_mesh_group: peer replicas, dg_add_function_sdf_multi, NCCL device form), GenerateSDF --gpus 2,
# the torch.distributed shardings, and bench.py under torchrun with N = 2.
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r2e_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi_capi.py tests/test_gpu_multi.py tests/test_gpu_cpp_facade.py -m gpu -q > $O/r2e_pytest.txt 2>&1; tail -5 $O/r2e_pytest.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-real > $O/r2e_bench_n2.json 2> $O/r2e_bench_n2.err
tail -c 400 $O/r2e_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e_bench_n2.json").read().strip().splitlines()[-1])
print("N=2", round(d["ms_per_step"],2), "ms", "e2e", d["e2e"].get("ms_per_step"), "c-abi multi", d["e2e"].get("single_process_c_abi"), "target", d["target_config"]["ms_per_step"], d["sharded_equals_single_launch"])
PY
