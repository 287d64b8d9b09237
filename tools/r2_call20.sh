#!/bin/bash
# ncu --set full of the packet-traversal node-loop kernel (K1_PACKET, 64 registers) on the bench workload (bunny.obj 128^3)
O=gpurun_out; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
DISCREGRID_B200_LIB=$PWD/build/variants/pk16.so timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2t_k1_packet python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2t_ncu.log 2>&1
tail -3 $O/r2t_ncu.log
