#!/bin/bash
# Round-2 GPU call 3: everything new on the real device -- the whole -m gpu suite (C-ABI multi-GPU group with n = 1, pipelined
# interpolate, dg_add_function_sdf, facade), smoke(), the rewritten bench.py end to end (both arms), the wave kernel after its fixes.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.txt 2>&1; tail -3 $O/r2c_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.txt 2>&1; tail -2 $O/r2c_smoke.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c_bench_ref.json 2> $O/r2c_bench_ref.err
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/r2c_bench.json 2> $O/r2c_bench.err
tail -c 600 $O/r2c_bench.err
DISCREGRID_B200_LIB=$PWD/build/variants/wave2.so timeout 300 python bench.py --steps 5 --warmup 3 --no-interp --no-cpu --no-e2e --no-real --no-density --no-target > $O/r2c_bench_wave2.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2c_bench_ref","r2c_bench","r2c_bench_wave2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2),"Mnodes/s", round(d["ms_per_step"],2),"ms", "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as ex: print(f, "ERR", ex)
PY
