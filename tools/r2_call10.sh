#!/bin/bash
# 8 GPUs: which sharding for the node loop?  K1-only bench lines (128^3 bunny + 256^3 target) for the rotated interleaved deal with 1 and 2
# launches per rank and for the node-id chunks.
O=gpurun_out; mkdir -p $O
for cfg in "--sharding interleaved --splits 1" "--sharding interleaved --splits 2" "--sharding chunks"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-real --no-interp --no-density --no-e2e $cfg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d['target_config']
print('$cfg', '| 128^3', round(d['ms_per_step'],2),'ms k1_only', round(d['timing']['k1_only_ms_per_step'],2), '| 256^3 target', round(t['ms_per_step'],2), 'ms', d['sharded_equals_single_launch'], t['sharded_equals_single_launch'])"
done > $O/r2j_n8_shardings.txt 2>&1
cat $O/r2j_n8_shardings.txt
