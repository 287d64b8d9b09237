#!/bin/bash
# quantised node records (K1_QBOX) for both node-loop kernels, and the wavefront kernel with the fp32 leaf-filter phase
O=gpurun_out; mkdir -p $O
for so in build/variants/*.so; do
  n=$(basename $so .so)
  ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 400 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1)
  echo "$n parity: $ok"
  for mesh in bunny torus; do
    DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --mesh $mesh --no-interp --no-cpu --no-e2e --no-real --no-density --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$n $mesh', 'K1 128^3', round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1),'Mnodes/s')"
  done
done > $O/r2f_sweep.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
DISCREGRID_B200_LIB=$PWD/build/variants/wave_qbox_lf.so timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2f_k1_wave_qbox_lf python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2f_ncu.log 2>&1
DISCREGRID_B200_LIB=$PWD/build/variants/pl_qbox.so timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2f_k1_pl_qbox python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density >> $O/r2f_ncu.log 2>&1
cat $O/r2f_sweep.txt
