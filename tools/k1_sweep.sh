#!/bin/bash
# runs the K1-only bench for every build/variants/*.so (GPU box)
for so in build/variants/*.so; do
  n=$(basename $so .so)
  DISCREGRID_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 1 --no-interp --no-cpu --no-e2e --no-density --no-target --no-real "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$n', round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1),'Mnodes/s')"
done
