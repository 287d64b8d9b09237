#!/bin/bash
# packet traversal, second batch: warp-uniform range stack (13 -> 16 KB of shared memory per block), occupancy targets, K1_PKT_K = 6, child prefetch;
# fallback reasons of a sample of blocks; ncu --set full of the 12-blocks build
O=gpurun_out; mkdir -p $O
{
DISCREGRID_B200_LIB=$PWD/build/variants/pkB12dbg.so timeout 300 python tools/k1_out_hash.py bunny 2>&1 | grep PKTFB | head -600 > $O/r2v_pktfb.txt
echo "debug lines: $(wc -l < $O/r2v_pktfb.txt)"
for n in pkB12 pkB10 pkB14 pkB12k6 pkB12pf; do
  echo "$n: $(DISCREGRID_B200_LIB=$PWD/build/variants/$n.so timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
done
ok=$(DISCREGRID_B200_LIB=$PWD/build/variants/pkB12.so timeout 400 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1); echo "pkB12 parity: $ok"
} > $O/r2v_packet.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
DISCREGRID_B200_LIB=$PWD/build/variants/pkB12.so timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2v_k1_packet python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2v_ncu.log 2>&1
cat $O/r2v_packet.txt; head -3 $O/r2v_pktfb.txt
