#!/bin/bash
# packet traversal: why do bricks fall back on the GPU (debug prints of a sample of blocks), and the stored pop key
O=gpurun_out; mkdir -p $O
{
DISCREGRID_B200_LIB=$PWD/build/variants/pkAdbg.so timeout 300 python tools/k1_out_hash.py bunny 2>&1 | grep PKTFB | head -400 > $O/r2u_pktfb.txt
wc -l $O/r2u_pktfb.txt
for n in pkA pkA12; do
  echo "$n: $(DISCREGRID_B200_LIB=$PWD/build/variants/$n.so timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
done
} > $O/r2u_packet.txt 2>&1
cat $O/r2u_packet.txt; head -5 $O/r2u_pktfb.txt
