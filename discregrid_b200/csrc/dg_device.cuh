// Shared device-side definitions for the sm_100a kernels.
//
// NUMERICAL CONTRACT: every translation unit that includes this header is compiled with -fmad=false.
// All fp64 expressions are written in the reference's operation order and rely on
// nvcc NOT contracting a*b+c into an FMA; fp64 '/' and sqrt() are IEEE-correct on the device.  dg_selftest()
// probes for contraction at run time and refuses to run if it is on.
#pragma once
#include <cstdint>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define DG_HD __host__ __device__ __forceinline__
#else                              // host-only translation units (reduce_field.cpp) share the index algebra; built with -ffp-contract=off
#define DG_HD inline
#endif

namespace dgb {

struct GridDev {                 // dg_grid_desc, device copy (passed by value as a kernel parameter)
    double mn[3], mx[3];
    double cell[3], inv[3];
    unsigned n[3];
    unsigned nv;                 // (nx+1)(ny+1)(nz+1)
    unsigned ne_x, ne_y, ne_z;   // edge counts per axis (cubic_lagrange_discrete_grid.cpp:611-614)
};

// indexToNodePosition, cubic_lagrange_discrete_grid.cpp:604-665.  Unsigned 32-bit index algebra exactly as
// the reference; position = min + cell*ijk, then the 1/3 or 2/3 offset ((1.0 + b) / 3.0) * cell[d].
DG_HD void node_position(const GridDev& g, unsigned l, double& x, double& y, double& z)
{
    const unsigned nx = g.n[0], ny = g.n[1], nz = g.n[2];
    unsigned i, j, k;
    int axis = -1;
    unsigned par = 0;
    if (l < g.nv) {
        const unsigned s = (ny + 1) * (nx + 1);
        k = l / s; const unsigned t = l - k * s;
        j = t / (nx + 1); i = t - j * (nx + 1);
    } else if (l < g.nv + 2 * g.ne_x) {
        l -= g.nv; const unsigned e = l >> 1; par = l & 1u; axis = 0;
        const unsigned s = (ny + 1) * nx;
        k = e / s; const unsigned t = e - k * s;
        j = t / nx; i = t - j * nx;
    } else if (l < g.nv + 2 * (g.ne_x + g.ne_y)) {
        l -= (g.nv + 2 * g.ne_x); const unsigned e = l >> 1; par = l & 1u; axis = 1;
        const unsigned s = (nz + 1) * ny;
        i = e / s; const unsigned t = e - i * s;
        k = t / ny; j = t - k * ny;
    } else {
        l -= (g.nv + 2 * (g.ne_x + g.ne_y)); const unsigned e = l >> 1; par = l & 1u; axis = 2;
        const unsigned s = (nx + 1) * nz;
        j = e / s; const unsigned t = e - j * s;
        i = t / nz; k = t - i * nz;
    }
    x = g.mn[0] + g.cell[0] * (double)i;
    y = g.mn[1] + g.cell[1] * (double)j;
    z = g.mn[2] + g.cell[2] * (double)k;
    const double f = (1.0 + (double)par) / 3.0;      // exactly 1.0/3.0 or 2.0/3.0
    if (axis == 0) x = x + f * g.cell[0];
    else if (axis == 1) y = y + f * g.cell[1];
    else if (axis == 2) z = z + f * g.cell[2];
}

// closed-form connectivity of cell (i,j,k): node id of local node jn (cubic_lagrange_discrete_grid.cpp:848-885)
DG_HD unsigned cell_node_id(const GridDev& g, unsigned i, unsigned j, unsigned k, unsigned jn)
{
    const unsigned nx = g.n[0], ny = g.n[1], nz = g.n[2];
    if (jn < 8) {
        const unsigned di = jn & 1u, dj = (jn >> 1) & 1u, dk = (jn >> 2) & 1u;
        return (nx + 1) * (ny + 1) * (k + dk) + (nx + 1) * (j + dj) + i + di;
    }
    const unsigned par = jn & 1u;
    const unsigned q = (jn - 8) >> 1;          // 0..11
    const unsigned grp = q >> 2, w = q & 3u;   // group 0: x edges, 1: y edges, 2: z edges
    if (grp == 0) {        // cell[8..15]: (j,k),(j,k+1),(j+1,k),(j+1,k+1)
        const unsigned dk = w & 1u, dj = w >> 1;
        return g.nv + 2 * (nx * (ny + 1) * (k + dk) + nx * (j + dj) + i) + par;
    } else if (grp == 1) { // cell[16..23]: (i,k),(i+1,k),(i,k+1),(i+1,k+1)
        const unsigned di = w & 1u, dk = w >> 1;
        return g.nv + 2 * g.ne_x + 2 * (ny * (nz + 1) * (i + di) + ny * (k + dk) + j) + par;
    } else {               // cell[24..31]: (i,j),(i,j+1),(i+1,j),(i+1,j+1)
        const unsigned dj = w & 1u, di = w >> 1;
        return g.nv + 2 * (g.ne_x + g.ne_y) + 2 * (nz * (nx + 1) * (j + dj) + nz * (i + di) + k) + par;
    }
}

}  // namespace dgb
