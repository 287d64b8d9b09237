#!/bin/bash
# GPU box: for every build/variants/*.so (tools/build_variants.py) -- parity first (the K1 GPU tests against that library), then the K1-only
# bench on the 128^3 workload and on the 256^3 / 100k-triangle target.  Typical round-2 opening:
#   python tools/build_variants.py base: fastdiv:-DK1_FAST_DIV=1 redux:-DK1_VOTE_REDUX=1 auto:-DK1_BRICK_AUTO=1 \
#          all:-DK1_FAST_DIV=1,-DK1_VOTE_REDUX=1,-DK1_BRICK_AUTO=1 cost:-DK1_COST_ORDER=1 k3div:-DK3_FAST_DIV=1          (here, no GPU needed)
#   gpurun --timeout 900 -- 'bash tools/k1_sweep.sh > gpurun_out/sweep.txt 2>&1'
for so in build/variants/*.so; do
  n=$(basename $so .so)
  ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 300 python -m pytest tests/test_gpu_k1_sdf.py tests/test_gpu_k3_density.py -m gpu -q -x -k "not full_size" 2>&1 | tail -1)
  echo "$n parity: $ok"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --no-interp --no-cpu --no-e2e --no-real "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('target_config') or {}; k=(d.get('density_map') or {})
print('$n', 'K1 128^3', round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1),'Mnodes/s | target', round(t.get('ms_per_step',0),1),'ms | K3', round(k.get('ms',0),1),'ms')"
done
