#!/bin/bash
# final verification of the round on one GPU, as the driver will run it: -m gpu suite, smoke(), both bench arms; ncu of the shipped K1 and K3
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2k_pytest.txt 2>&1; tail -3 $O/r2k_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.txt 2>&1; tail -1 $O/r2k_smoke.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2k_bench_ref.json 2> $O/r2k_bench_ref.err
timeout 1800 python bench.py --steps 20 --warmup 5 > $O/r2k_bench.json 2> $O/r2k_bench.err
tail -c 400 $O/r2k_bench.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2k_k1_bunny128 python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2k_ncu.log 2>&1
timeout 400 $NCU -k regex:density_map -c 1 -f -o $O/r2k_k3_bunny128 python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target >> $O/r2k_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-real > /dev/null 2>&1
python - <<'PY'
import json
for f in ("r2k_bench_ref","r2k_bench"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2),"Mnodes/s", round(d["ms_per_step"],2),"ms", "e2e", (d.get("e2e") or {}).get("ms_per_step"), "parity", d.get("parity_full"), "tools", d.get("tools_e2e"))
    except Exception as ex: print(f, "ERR", ex)
PY
