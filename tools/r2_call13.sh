#!/bin/bash
# A/B of the outside-in plane-group order at N = 1 (bunny 128^3, torus 128^3, target 256^3), two passes each
O=gpurun_out; mkdir -p $O
for pass in 1 2; do for so in build/variants/oi0.so build/variants/oi1.so; do
  n=$(basename $so .so)
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 8 --warmup 3 --mesh bunny --no-interp --no-cpu --no-e2e --no-real --no-density 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric\"')][-1]); print('$n bunny', round(d['ms_per_step'],2),'ms | target', round(d['target_config']['ms_per_step'],1))"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 8 --warmup 3 --mesh torus --no-interp --no-cpu --no-e2e --no-real --no-density --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric\"')][-1]); print('$n torus', round(d['ms_per_step'],2),'ms')"
done; done > $O/r2m_outside_in.txt 2>&1
cat $O/r2m_outside_in.txt
