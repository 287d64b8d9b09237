#!/bin/bash
# dg_sample_sdf with pre-fault / copy workers for large pageable ranges: A/B against the single-threaded copy, 128^3 and 256^3; K1 parity tests on the rebuilt library
O=gpurun_out; mkdir -p $O
{
timeout 200 python tools/e2e_probe.py bunny 2>&1 | grep -E "ms \(" | sed 's/^/helpers on:  /'
DG_HOST_HELPERS_MIN_BYTES=1000000000000 timeout 200 python tools/e2e_probe.py bunny 2>&1 | grep -E "dg_sample_sdf" | sed 's/^/helpers off: /'
timeout 200 python tools/e2e_probe.py target 2>&1 | grep -E "ms \(" | sed 's/^/helpers on:  /'
DG_HOST_HELPERS_MIN_BYTES=1000000000000 timeout 200 python tools/e2e_probe.py target 2>&1 | grep -E "dg_sample_sdf" | sed 's/^/helpers off: /'
timeout 300 python -m pytest tests/test_gpu_k1_sdf.py -m gpu -q -x 2>&1 | tail -1
} > $O/r2z_sample_sdf_helpers.txt 2>&1
cat $O/r2z_sample_sdf_helpers.txt
