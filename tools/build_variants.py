#!/usr/bin/env python
"""Builds tuning variants of libdiscregrid_b200.so (same sources, different -D knobs) into build/variants/<name>.so.
Usage: python tools/build_variants.py name1:-DK1_RUN=16,-DK1_FETCH_MIN=4 name2:...   (select one with DISCREGRID_B200_LIB)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "discregrid_b200", "csrc")
NV = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-std=c++17", "-O3", "-lineinfo",
      "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden"]
out_dir = os.path.join(ROOT, "build", "variants"); os.makedirs(out_dir, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    srcs = [os.path.join(SRC, f) for f in ("k1_sdf.cu", "k2_interp.cu", "k3_density.cu", "k4_reduce.cu", "dg_api.cu", "bvh_build.cpp", "host_threads.cpp", "reduce_field.cpp", "obj_reader.cpp", "sort_replay.cpp")]
    cmd = NV + defs + ["-shared", "-o", os.path.join(out_dir, name + ".so")] + srcs
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    assert p.wait() == 0, name
    print("built", name)
