#!/bin/bash
# 8 GPUs: K1-only bench lines (128^3 bunny + 256^3 target) for the rotated interleaved deal with 1 and 2 launches per rank
O=gpurun_out; mkdir -p $O
i=0
for cfg in "--sharding interleaved --splits 1" "--sharding interleaved --splits 2"; do
  i=$((i+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus 8 --steps 10 --warmup 3 --no-real --no-interp --no-density --no-e2e $cfg > $O/r2j_n8_$i.out 2> $O/r2j_n8_$i.err
  python - "$cfg" $O/r2j_n8_$i.out <<'PY'
import json,sys
lines=[l for l in open(sys.argv[2]).read().splitlines() if l.startswith('{"metric"')]
if not lines: print(sys.argv[1], "NO JSON LINE"); sys.exit(0)
d=json.loads(lines[-1]); t=d['target_config']
print(sys.argv[1], '| 128^3', round(d['ms_per_step'],2),'ms k1_only', round(d['timing']['k1_only_ms_per_step'],2), '| 256^3 target', round(t['ms_per_step'],2), 'ms', d['sharded_equals_single_launch'], t['sharded_equals_single_launch'])
PY
done > $O/r2j_n8_shardings.txt 2>&1
cat $O/r2j_n8_shardings.txt; tail -c 300 $O/r2j_n8_1.err
