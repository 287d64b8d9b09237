#!/bin/bash
# last check of the round on one GPU, the final build as the driver will run it: -m gpu suite, smoke(), both bench arms
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2y_pytest.txt 2>&1; tail -2 $O/r2y_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2y_smoke.txt 2>&1; tail -1 $O/r2y_smoke.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2y_bench_ref.json 2> $O/r2y_bench_ref.err
timeout 1800 python bench.py --steps 20 --warmup 5 > $O/r2y_bench.json 2> $O/r2y_bench.err
python - <<'PY'
import json
for f in ("r2y_bench_ref","r2y_bench"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith('{')][-1])
        print(f, round(d["value"]/1e6,2),"Mnodes/s", round(d["ms_per_step"],2),"ms", "e2e", (d.get("e2e") or {}).get("ms_per_step"), "parity", d.get("parity_full"))
    except Exception as ex: print(f, "ERR", ex)
PY
