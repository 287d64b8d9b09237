// tests/emu/k4_emu.cpp -- TEST INFRASTRUCTURE: the product's reduceField kernels (discregrid_b200/csrc/k4_reduce.cu) compiled for the CPU
// through tests/emu/cuda_emu.h (part of build/bin/libdgemu.so).
#define DG_EMU 1
#include "cuda_emu.h"
#include "../../discregrid_b200/csrc/k4_reduce.cu"
