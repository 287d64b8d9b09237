#!/bin/bash
# the packet walk as the shipped default (K1_PKT_MIN_BLOCKS = 14): occupancy variants, the whole -m gpu suite, smoke(), the default bench arm, ncu of K1 + launch list
O=gpurun_out; mkdir -p $O
{
echo "product: $(timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
for n in pkC13 pkC15 pkC16; do
  echo "$n: $(DISCREGRID_B200_LIB=$PWD/build/variants/$n.so timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
done
} > $O/r2w_variants.txt 2>&1
cat $O/r2w_variants.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2w_pytest.txt 2>&1; tail -3 $O/r2w_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2w_smoke.txt 2>&1; tail -1 $O/r2w_smoke.txt
timeout 1800 python bench.py --steps 20 --warmup 5 > $O/r2w_bench.json 2> $O/r2w_bench.err
tail -c 300 $O/r2w_bench.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2w_k1_bunny128 python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2w_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-real > /dev/null 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2w_bench.json").read().splitlines() if l.startswith('{"metric"')][-1])
print("bench", round(d["value"]/1e6,2),"Mnodes/s", round(d["ms_per_step"],2),"ms", "e2e", (d.get("e2e") or {}).get("ms_per_step"), "parity", d.get("parity_full"), "roofline", round(d["roofline"]["frac"],3), "target", d["target_config"].get("ms_per_step"), (d["target_config"].get("e2e") or {}).get("ms_per_step"), "tools", d.get("tools_e2e"))
for m in d.get("reference_meshes") or []: print(m["mesh"], m["ms_per_step"], m.get("parity"))
PY
