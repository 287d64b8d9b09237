// Kernel launch spelling shared by the .cu files.  tests/emu compiles these translation units for the CPU (DG_EMU): the device code
// runs unchanged with the lanes of a warp as fibers (tests/emu/cuda_emu.h), so kernels AND launchers can be checked against the
// oracle without a GPU; only the launch syntax, the inline-PTX helpers and a few runtime calls differ there.
#pragma once
#ifdef DG_EMU
#define DG_KERNEL_LAUNCH(kernel, grid, block, smem, stream, ...) dg_emu::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define DG_AFTER_LAUNCH() cudaSuccess
#else
#define DG_KERNEL_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DG_AFTER_LAUNCH() cudaGetLastError()
#endif
