#!/bin/bash
# last batch of existing K1 knobs against the new default build (redux vote + per-array bricks), parity first
O=gpurun_out; mkdir -p $O
for mesh in bunny torus; do timeout 300 python bench.py --steps 5 --warmup 3 --mesh $mesh --no-interp --no-cpu --no-e2e --no-real --no-density --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default $mesh K1 128^3', round(d['ms_per_step'],2),'ms')"; done > $O/r2i_sweep.txt 2>&1
for so in build/variants/*.so; do
  n=$(basename $so .so)
  ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 400 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1)
  echo "$n parity: $ok"
  for mesh in bunny torus; do DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --mesh $mesh --no-interp --no-cpu --no-e2e --no-real --no-density --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$n $mesh K1 128^3', round(d['ms_per_step'],2),'ms')"; done
done >> $O/r2i_sweep.txt 2>&1
cat $O/r2i_sweep.txt
