// Launch interface of the K3 kernel (k3_density.cu): GenerateDensityMap's per-node function.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "k2_interp.h"

namespace dgb {
cudaError_t k3_launch_density(const FieldDev& f, double h, double rho0, int no_reduction, uint64_t l_begin, uint64_t count,
                              double* d_out, cudaStream_t stream);
}
