#!/bin/bash
# packet walk with one warp per block (finer block granularity): hashes + times against the shipped 64-thread build
O=gpurun_out; mkdir -p $O
{
echo "product: $(timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
for n in t32mb28 t32mb26; do
  echo "$n: $(DISCREGRID_B200_LIB=$PWD/build/variants/$n.so timeout 300 python tools/k1_out_hash.py 2>&1 | tr '\n' '|')"
done
} > $O/r2z_t32.txt 2>&1
cat $O/r2z_t32.txt
