#!/bin/bash
# packet traversal (K1_PACKET): parity, full-size output hashes against the per-lane build, times at several occupancy targets
O=gpurun_out; mkdir -p $O
{
echo "default: $(timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
for so in build/variants/pk10.so build/variants/pk16.so build/variants/pk12.so build/variants/pk8.so build/variants/pk10t128.so; do
  n=$(basename $so .so)
  if [ $n = pk10 ]; then
    ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 400 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1); echo "$n parity: $ok"
  fi
  echo "$n: $(DISCREGRID_B200_LIB=$PWD/$so timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
done
} > $O/r2s_packet.txt 2>&1
cat $O/r2s_packet.txt
