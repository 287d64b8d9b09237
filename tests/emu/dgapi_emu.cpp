// tests/emu/dgapi_emu.cpp -- TEST INFRASTRUCTURE: the product's C-ABI translation unit (dg_api.cu) compiled for the CPU; together
// with k1_emu.cpp / k2_emu.cpp / k3_emu.cpp, the host sources and cudart_stub.cpp it forms build/bin/libdgemu.so, the whole library
// with every kernel emulated (tests/emu/cuda_emu.h).  Pointing DISCREGRID_B200_LIB at it lets the Python-level `-m gpu` tests that do
// not need torch be rehearsed without a GPU (tests/test_gpu_rehearsal.py).  It is not a CPU fallback: the product never loads it.
#define DG_EMU 1
#include "cuda_emu.h"
#include "../../discregrid_b200/csrc/dg_api.cu"
