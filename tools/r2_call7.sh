#!/bin/bash
# last knob checks (redux vote, per-array brick shape, K3 occupancy) + the new host-pipeline / reduce-field GPU tests
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_k2_interp.py tests/test_gpu_reduce_field.py tests/test_gpu_multi_capi.py -m gpu -q > $O/r2g_pytest.txt 2>&1; tail -3 $O/r2g_pytest.txt
for so in build/variants/*.so; do
  n=$(basename $so .so)
  case $n in
    k3*) DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 3 --warmup 3 --no-interp --no-cpu --no-e2e --no-real --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$n K3 128^3 bunny field', round(d['density_map']['ms'],1),'ms')";;
    *) for mesh in bunny torus; do DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --mesh $mesh --no-interp --no-cpu --no-e2e --no-real --no-density --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$n $mesh K1 128^3', round(d['ms_per_step'],2),'ms')"; done;;
  esac
done > $O/r2g_sweep.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-interp --no-cpu --no-e2e --no-real --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default K3 128^3 bunny field', round(d['density_map']['ms'],1),'ms', 'K1', round(d['ms_per_step'],2))" >> $O/r2g_sweep.txt
cat $O/r2g_sweep.txt
