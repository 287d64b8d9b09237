#!/bin/bash
# Round-2 GPU call 2: the wavefront node-loop kernel (K1_WAVE=1) -- parity against the oracle / goldens first, then the K1-only bench on
# the headline workload (bunny.obj 128^3), the round-1 torus and the 256^3 target; ncu --set full of the winner.
O=gpurun_out
mkdir -p $O
for so in build/variants/*.so; do
  n=$(basename $so .so)
  ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 400 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1)
  echo "$n parity: $ok"
  for mesh in bunny torus; do
    extra="--no-target"; [ "$mesh" = bunny ] && [ "$n" = wave -o "$n" = base ] && extra=""
    DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --mesh $mesh --no-interp --no-cpu --no-e2e --no-real --no-density $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('target_config') or {}
print('$n $mesh', 'K1 128^3', round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1),'Mnodes/s | target', round(t.get('ms_per_step',0),1),'ms', 'fp64 frac', round(d['roofline']['frac'],4), 'peak', round(d['roofline']['peak'],2))"
  done
done > $O/r2b_sweep.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
DISCREGRID_B200_LIB=$PWD/build/variants/wave.so timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2b_k1wave_bunny128 python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2b_ncu_k1wave.log 2>&1
cat $O/r2b_sweep.txt
