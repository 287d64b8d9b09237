#!/bin/bash
# final numbers of the round on one GPU: both bench arms as the driver runs them, plus the multi-GPU C-ABI test with one GPU
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_k1_sdf.py tests/test_gpu_multi_capi.py -m gpu -x -q > $O/r2n_pytest.txt 2>&1; tail -2 $O/r2n_pytest.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2n_bench_ref.json 2> $O/r2n_bench_ref.err
timeout 1800 python bench.py --steps 20 --warmup 5 > $O/r2n_bench.json 2> $O/r2n_bench.err
python - <<'PY'
import json
for f in ("r2n_bench_ref","r2n_bench"):
    d=json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith('{')][-1])
    print(f, round(d["value"]/1e6,2),"Mnodes/s", round(d["ms_per_step"],2),"ms", "e2e", (d.get("e2e") or {}).get("ms_per_step"))
PY
