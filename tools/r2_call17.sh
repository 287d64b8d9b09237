#!/bin/bash
# super-tile block enumeration (K1_SUPERTILE): parity, N = 1 times, per-part times of an 8-way split on one GPU
O=gpurun_out; mkdir -p $O
for so in build/variants/st*.so; do
  n=$(basename $so .so)
  ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 400 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1)
  echo "$n parity: $ok"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 8 --warmup 3 --mesh bunny --no-interp --no-cpu --no-e2e --no-real --no-density 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric\"')][-1]); print('$n bunny', round(d['ms_per_step'],2),'ms | target', round(d['target_config']['ms_per_step'],1))"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 8 --warmup 3 --mesh torus --no-interp --no-cpu --no-e2e --no-real --no-density --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric\"')][-1]); print('$n torus', round(d['ms_per_step'],2),'ms')"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python tools/part_times.py 8 128 bunny 2>&1 | grep -E "interleaved |chunks|slab"
done > $O/r2q_supertile.txt 2>&1
cat $O/r2q_supertile.txt
