#!/bin/bash
O=gpurun_out; mkdir -p $O
for w in bunny target; do for t in "" 4 8 16 32; do
  if [ -z "$t" ]; then timeout 300 python tools/e2e_probe.py $w; else DG_HOST_THREADS=$t timeout 300 python tools/e2e_probe.py $w; fi
done; done > $O/r2d_e2e_probe.txt 2>&1
cat $O/r2d_e2e_probe.txt
