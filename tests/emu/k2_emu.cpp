// tests/emu/k2_emu.cpp -- TEST INFRASTRUCTURE: the product's K2 translation unit (k2_interp.cu: pack / axis tables / interpolate / shape
// functions + launchers) compiled for the CPU through tests/emu/cuda_emu.h.  The TMA bulk copy + mbarrier staging is replaced by a
// plain copy there (cuda_emu.h cannot run PTX): what is checked is the cell lookup, the packed layout and the arithmetic.
#define DG_EMU 1
#include "cuda_emu.h"
#include "../../discregrid_b200/csrc/k2_interp.cu"
