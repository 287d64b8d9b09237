#!/bin/bash
# heaviest-class-first launch order (K1_TAIL_CLASSES): parity, N = 1 times, and the per-part times of an 8-way split on one GPU
O=gpurun_out; mkdir -p $O
for so in build/variants/tail*.so; do
  n=$(basename $so .so)
  ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 400 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1)
  echo "$n parity: $ok"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 8 --warmup 3 --mesh bunny --no-interp --no-cpu --no-e2e --no-real --no-density 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric\"')][-1]); print('$n bunny', round(d['ms_per_step'],2),'ms | target', round(d['target_config']['ms_per_step'],1))"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python tools/part_times.py 8 128 bunny 2>&1 | grep -E "single|interleaved |chunks"
done > $O/r2p_tail_order.txt 2>&1
cat $O/r2p_tail_order.txt
