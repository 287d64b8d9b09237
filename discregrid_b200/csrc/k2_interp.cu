// K2: batched CubicLagrangeDiscreteGrid::interpolate (+ gradient), one query per lane.
//
// Replaces interpolate (cubic_lagrange_discrete_grid.cpp:977-1063), determineShapeFunctions + split interpolate
// (:901-975), shape_function_ (:339-580) and the index helpers discrete_grid.cpp:8-38.
//
// Device layout (built once by dg_field_create): the 32 coefficients of every kept cell are gathered into one
// contiguous, 256-byte-aligned block packed[cell][32] (the reference keeps a shared node array + a 32-index
// table per cell: 16 scattered 16-byte pairs per query, i.e. >= 2x the algorithmic DRAM traffic).  A query's
// compulsory traffic is then exactly 24 B (x) + 256 B (block) + 8 B (phi) [+ 24 B (grad)].
//
// Kernel: each lane locates its cell, then issues ONE bulk async copy (cp.async.bulk, the TMA engine's 1-D
// path) of its 256-byte block into the warp's shared-memory staging area and signals the warp's mbarrier.
// While the copies are in flight the lane evaluates everything that does not need coefficients (xi and all
// shape-function temporaries).  After the mbarrier flips, the 32 terms are accumulated in the reference's
// order j = 0..31, each shape function / Jacobian row evaluated on the fly (no N[32]/dN[32][3] arrays in
// registers).  Staging slots are 272 bytes apart so that the 16-byte shared loads of a quarter-warp hit
// distinct banks.  No registers are held by in-flight data, so memory-level parallelism is bounded only by
// shared memory (8.5 KB per warp).
//
// The per-axis reference-cell maps c0 = 2/denom, c1 = (hi+lo)/denom (:1000-1002) depend only on the cell
// index along that axis; they are tabulated once per field with the reference's exact operation order, which
// removes six fp64 divisions per query without changing a bit.
#include "dg_device.cuh"
#include "k2_interp.h"
#include "dg_launch.h"

#include <cfloat>
#include <climits>
#include <cstdint>
#include <cstring>

namespace dgb {

namespace {

constexpr int K2_SLOT_BYTES = 272;          // 256-byte block + 16 bytes of padding (bank spreading)
constexpr int K2_WARPS = K2_THREADS / 32;

#ifdef DG_EMU
// CPU emulation: "shared" addresses are plain pointers, the bulk copy is a memcpy that has completed when it returns, the mbarrier
// has nothing left to wait for (each lane only reads the slot it copied itself)
typedef uintptr_t smem_addr_t;
inline smem_addr_t smem_u32(const void* p) { return (smem_addr_t)p; }
inline void mbar_init(smem_addr_t, unsigned) {}
inline void mbar_expect_tx(smem_addr_t, unsigned) {}
inline void bulk_g2s(smem_addr_t dst, const void* src, unsigned bytes, smem_addr_t) { std::memcpy((void*)dst, src, bytes); }
inline void mbar_wait(smem_addr_t, unsigned) {}
inline void __syncwarp() {}
#else
typedef unsigned smem_addr_t;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
#endif

// Gathers the packed per-cell blocks.  One thread per (kept cell, local node).
__global__ void pack_cells_kernel(GridDev g, const double* __restrict__ nodes, const unsigned* __restrict__ cells /*nullable*/,
                                  unsigned long long n32, double* __restrict__ packed)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n32) return;
    unsigned v;
    if (cells) v = cells[idx];
    else {
        const unsigned c = (unsigned)(idx >> 5), jn = (unsigned)(idx & 31u);
        const unsigned nxy = g.n[0] * g.n[1];
        const unsigned k = c / nxy, t = c - k * nxy, j = t / g.n[0], i = t - j * g.n[0];
        v = cell_node_id(g, i, j, k, jn);
    }
    packed[idx] = nodes[v];
}

// Per-axis reference-cell maps: tab[off_d + m] = (c0, c1) for cell index m along axis d.
// subdomain(): origin = min + double(m)*cell, max = origin + cell (discrete_grid.cpp:26-32);
// denom = max - min, c0 = 2/denom, c1 = (max+min)/denom (cubic_lagrange_discrete_grid.cpp:1000-1002).
__global__ void axis_tables_kernel(GridDev g, double2* __restrict__ tab)
{
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    if (idx >= n0 + n1 + n2) return;
    int d; unsigned m;
    if (idx < n0) { d = 0; m = idx; } else if (idx < n0 + n1) { d = 1; m = idx - n0; } else { d = 2; m = idx - n0 - n1; }
    const double lo = g.mn[d] + (double)m * g.cell[d];
    const double hi = lo + g.cell[d];
    const double denom = hi - lo;
    tab[idx] = make_double2(2.0 / denom, (hi + lo) / denom);
}

template <bool GRAD>
__global__ void __launch_bounds__(K2_THREADS)
interpolate_kernel(GridDev g, const double* __restrict__ packed, const unsigned* __restrict__ cell_map /*nullable*/,
                   const double2* __restrict__ tab, const double* __restrict__ xq, unsigned long long n,
                   double* __restrict__ phi_out, double* __restrict__ grad_out)
{
    __shared__ __align__(16) unsigned char stage[K2_WARPS][32 * K2_SLOT_BYTES];
    __shared__ __align__(8) unsigned long long bars[K2_WARPS];

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned long long q0 = (unsigned long long)blockIdx.x * K2_THREADS + threadIdx.x;
    if (q0 - lane >= n) return;                             // whole warp past the end
    const bool live = q0 < n;
    const unsigned long long q = live ? q0 : n - 1;         // tail lanes shadow the last query (no store)

    const smem_addr_t bar = smem_u32(&bars[warp]);
    if (lane == 0) { mbar_init(bar, 1); mbar_expect_tx(bar, 32u * 256u); }
    __syncwarp();

    const double x = xq[3 * q], y = xq[3 * q + 1], z = xq[3 * q + 2];

    // ---- cell location (:981-994).  NaN fails `contains` like Eigen's AlignedBox.
    bool valid = (g.mn[0] <= x) && (x <= g.mx[0]) && (g.mn[1] <= y) && (y <= g.mx[1]) && (g.mn[2] <= z) && (z <= g.mx[2]);
    unsigned mi0 = 0, mi1 = 0, mi2 = 0, cell = 0;
    if (valid) {
        mi0 = (unsigned)((x - g.mn[0]) * g.inv[0]);
        mi1 = (unsigned)((y - g.mn[1]) * g.inv[1]);
        mi2 = (unsigned)((z - g.mn[2]) * g.inv[2]);
        if (mi0 >= g.n[0]) mi0 = g.n[0] - 1;
        if (mi1 >= g.n[1]) mi1 = g.n[1] - 1;
        if (mi2 >= g.n[2]) mi2 = g.n[2] - 1;
        const unsigned i = g.n[1] * g.n[0] * mi2 + g.n[0] * mi1 + mi0;     // multiToSingleIndex
        cell = cell_map ? __ldg(cell_map + i) : i;
        if (cell == UINT_MAX) { valid = false; cell = 0; }                  // removed by reduceField (:993)
    }

    // ---- one bulk copy per lane: packed[cell][0..31] -> stage[warp][lane]
    const smem_addr_t slot = smem_u32(&stage[warp][lane * K2_SLOT_BYTES]);
    bulk_g2s(slot, packed + (size_t)cell * 32, 256u, bar);

    // ---- reference-cell coordinates (:996-1003) and every coefficient-independent temporary (:344-385, :440-460)
    const double2 tx = __ldg(tab + mi0), ty = __ldg(tab + g.n[0] + mi1), tz = __ldg(tab + g.n[0] + g.n[1] + mi2);
    const double X = tx.x * x - tx.y, Y = ty.x * y - ty.y, Z = tz.x * z - tz.y;     // xi = c0*x - c1
    const double x2 = X * X, y2 = Y * Y, z2 = Z * Z;
    const double _1mx = 1.0 - X, _1my = 1.0 - Y, _1mz = 1.0 - Z;
    const double _1px = 1.0 + X, _1py = 1.0 + Y, _1pz = 1.0 + Z;
    const double _1m3x = 1.0 - 3.0 * X, _1m3y = 1.0 - 3.0 * Y, _1m3z = 1.0 - 3.0 * Z;
    const double _1p3x = 1.0 + 3.0 * X, _1p3y = 1.0 + 3.0 * Y, _1p3z = 1.0 + 3.0 * Z;
    const double mxmy = _1mx * _1my, mxpy = _1mx * _1py, pxmy = _1px * _1my, pxpy = _1px * _1py;
    const double mxmz = _1mx * _1mz, mxpz = _1mx * _1pz, pxmz = _1px * _1mz, pxpz = _1px * _1pz;
    const double mymz = _1my * _1mz, mypz = _1my * _1pz, pymz = _1py * _1mz, pypz = _1py * _1pz;
    const double _1mx2 = 1.0 - x2, _1my2 = 1.0 - y2, _1mz2 = 1.0 - z2;
    const double facc = 1.0 / 64.0 * (9.0 * (x2 + y2 + z2) - 19.0);                   // :388
    const double fx = 9.0 / 64.0 * _1mx2, fy = 9.0 / 64.0 * _1my2, fz = 9.0 / 64.0 * _1mz2;   // :400,412,424
    const double fxm = fx * _1m3x, fxp = fx * _1p3x, fym = fy * _1m3y, fyp = fy * _1p3y, fzm = fz * _1m3z, fzp = fz * _1p3z;

    double xmA = 0, xpA = 0, ymB = 0, ypB = 0, zmC = 0, zpC = 0, mX = 0, pX = 0, Xm = 0, Xp = 0, mY = 0, pY = 0, Ym = 0, Yp = 0,
           mZ = 0, pZ = 0, Zm = 0, Zp = 0;
    if (GRAD) {
        const double A = 9.0 * (3.0 * x2 + y2 + z2) - 19.0;                           // :440-442
        const double B = 9.0 * (x2 + 3.0 * y2 + z2) - 19.0;
        const double C = 9.0 * (x2 + y2 + 3.0 * z2) - 19.0;
        const double _18x = 18.0 * X, _18y = 18.0 * Y, _18z = 18.0 * Z;
        xmA = _18x - A; xpA = _18x + A; ymB = _18y - B; ypB = _18y + B; zmC = _18z - C; zpC = _18z + C;   // :455-460
        const double _3m9x2 = 3.0 - 9.0 * x2, _3m9y2 = 3.0 - 9.0 * y2, _3m9z2 = 3.0 - 9.0 * z2;
        const double _2x = 2.0 * X, _2y = 2.0 * Y, _2z = 2.0 * Z;
        mX = -_3m9x2 - _2x; pX = _3m9x2 - _2x; Xm = _1mx2 * _1m3x; Xp = _1mx2 * _1p3x;                    // :489-492
        mY = -_3m9y2 - _2y; pY = _3m9y2 - _2y; Ym = _1my2 * _1m3y; Yp = _1my2 * _1p3y;                    // :518-521
        mZ = -_3m9z2 - _2z; pZ = _3m9z2 - _2z; Zm = _1mz2 * _1m3z; Zp = _1mz2 * _1p3z;                    // :547-550
    }

    // ---- wait for the warp's 32 blocks
    mbar_wait(bar, 0);
    const double2* cb = reinterpret_cast<const double2*>(&stage[warp][lane * K2_SLOT_BYTES]);

    double phi = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
    bool missing = false;
    // one term of the sequential sums (:1044-1060); corner rows are scaled by /64, edge rows by *(9/64) (:487, :576)
#define DG_CORNER(cj, n, dx, dy, dz)                                         \
    {                                                                        \
        const double c_ = (cj);                                              \
        missing |= (c_ == DBL_MAX);                                          \
        phi = phi + c_ * (n);                                                \
        if (GRAD) { g0 = g0 + c_ * ((dx) * 0.015625); g1 = g1 + c_ * ((dy) * 0.015625); g2 = g2 + c_ * ((dz) * 0.015625); } \
    }
#define DG_EDGE(cj, n, dx, dy, dz)                                           \
    {                                                                        \
        const double c_ = (cj);                                              \
        missing |= (c_ == DBL_MAX);                                          \
        phi = phi + c_ * (n);                                                \
        if (GRAD) { g0 = g0 + c_ * ((dx) * (9.0 / 64.0)); g1 = g1 + c_ * ((dy) * (9.0 / 64.0)); g2 = g2 + c_ * ((dz) * (9.0 / 64.0)); } \
    }
    double2 c;
    // corners (:389-396, :462-485)
    c = cb[0];
    DG_CORNER(c.x, facc * mxmy * _1mz, xmA * mymz, mxmz * ymB, mxmy * zmC)
    DG_CORNER(c.y, facc * pxmy * _1mz, xpA * mymz, pxmz * ymB, pxmy * zmC)
    c = cb[1];
    DG_CORNER(c.x, facc * mxpy * _1mz, xmA * pymz, mxmz * ypB, mxpy * zmC)
    DG_CORNER(c.y, facc * pxpy * _1mz, xpA * pymz, pxmz * ypB, pxpy * zmC)
    c = cb[2];
    DG_CORNER(c.x, facc * mxmy * _1pz, xmA * mypz, mxpz * ymB, mxmy * zpC)
    DG_CORNER(c.y, facc * pxmy * _1pz, xpA * mypz, pxpz * ymB, pxmy * zpC)
    c = cb[3];
    DG_CORNER(c.x, facc * mxpy * _1pz, xmA * pypz, mxpz * ypB, mxpy * zpC)
    DG_CORNER(c.y, facc * pxpy * _1pz, xpA * pypz, pxpz * ypB, pxpy * zpC)
    // x edges (:403-410, :493-516)
    c = cb[4];
    DG_EDGE(c.x, fxm * mymz, mX * mymz, -Xm * _1mz, -Xm * _1my)
    DG_EDGE(c.y, fxp * mymz, pX * mymz, -Xp * _1mz, -Xp * _1my)
    c = cb[5];
    DG_EDGE(c.x, fxm * mypz, mX * mypz, -Xm * _1pz, Xm * _1my)
    DG_EDGE(c.y, fxp * mypz, pX * mypz, -Xp * _1pz, Xp * _1my)
    c = cb[6];
    DG_EDGE(c.x, fxm * pymz, mX * pymz, Xm * _1mz, -Xm * _1py)
    DG_EDGE(c.y, fxp * pymz, pX * pymz, Xp * _1mz, -Xp * _1py)
    c = cb[7];
    DG_EDGE(c.x, fxm * pypz, mX * pypz, Xm * _1pz, Xm * _1py)
    DG_EDGE(c.y, fxp * pypz, pX * pypz, Xp * _1pz, Xp * _1py)
    // y edges (:415-422, :522-545)
    c = cb[8];
    DG_EDGE(c.x, fym * mxmz, -Ym * _1mz, mY * mxmz, -Ym * _1mx)
    DG_EDGE(c.y, fyp * mxmz, -Yp * _1mz, pY * mxmz, -Yp * _1mx)
    c = cb[9];
    DG_EDGE(c.x, fym * pxmz, Ym * _1mz, mY * pxmz, -Ym * _1px)
    DG_EDGE(c.y, fyp * pxmz, Yp * _1mz, pY * pxmz, -Yp * _1px)
    c = cb[10];
    DG_EDGE(c.x, fym * mxpz, -Ym * _1pz, mY * mxpz, Ym * _1mx)
    DG_EDGE(c.y, fyp * mxpz, -Yp * _1pz, pY * mxpz, Yp * _1mx)
    c = cb[11];
    DG_EDGE(c.x, fym * pxpz, Ym * _1pz, mY * pxpz, Ym * _1px)
    DG_EDGE(c.y, fyp * pxpz, Yp * _1pz, pY * pxpz, Yp * _1px)
    // z edges (:427-434, :551-574)
    c = cb[12];
    DG_EDGE(c.x, fzm * mxmy, -Zm * _1my, -Zm * _1mx, mZ * mxmy)
    DG_EDGE(c.y, fzp * mxmy, -Zp * _1my, -Zp * _1mx, pZ * mxmy)
    c = cb[13];
    DG_EDGE(c.x, fzm * mxpy, -Zm * _1py, Zm * _1mx, mZ * mxpy)
    DG_EDGE(c.y, fzp * mxpy, -Zp * _1py, Zp * _1mx, pZ * mxpy)
    c = cb[14];
    DG_EDGE(c.x, fzm * pxmy, Zm * _1my, -Zm * _1px, mZ * pxmy)
    DG_EDGE(c.y, fzp * pxmy, Zp * _1my, -Zp * _1px, pZ * pxmy)
    c = cb[15];
    DG_EDGE(c.x, fzm * pxpy, Zm * _1py, Zm * _1px, mZ * pxpy)
    DG_EDGE(c.y, fzp * pxpy, Zp * _1py, Zp * _1px, pZ * pxpy)
#undef DG_CORNER
#undef DG_EDGE

    if (!live) return;
    if (!valid || missing) {                        // sentinel paths (:981-982, :993-994, :1015-1018, :1050-1054)
        phi_out[q] = DBL_MAX;
        if (GRAD) { grad_out[3 * q] = 0.0; grad_out[3 * q + 1] = 0.0; grad_out[3 * q + 2] = 0.0; }
        return;
    }
    phi_out[q] = phi;
    if (GRAD) {                                     // gradient->array() *= c0.array()  (:1060)
        grad_out[3 * q] = g0 * tx.x;
        grad_out[3 * q + 1] = g1 * ty.x;
        grad_out[3 * q + 2] = g2 * tz.x;
    }
}

// shape_function_ alone (diagnostics / dg_shape_functions): evaluates through the same code path by
// interpolating 32 unit fields would be wasteful; instead a direct restatement, one thread per point.
__global__ void shape_functions_kernel(const double* __restrict__ xi, unsigned long long n, double* __restrict__ N, double* __restrict__ dN)
{
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const double X = xi[3 * q], Y = xi[3 * q + 1], Z = xi[3 * q + 2];
    const double x2 = X * X, y2 = Y * Y, z2 = Z * Z;
    const double _1mx = 1.0 - X, _1my = 1.0 - Y, _1mz = 1.0 - Z;
    const double _1px = 1.0 + X, _1py = 1.0 + Y, _1pz = 1.0 + Z;
    const double _1m3x = 1.0 - 3.0 * X, _1m3y = 1.0 - 3.0 * Y, _1m3z = 1.0 - 3.0 * Z;
    const double _1p3x = 1.0 + 3.0 * X, _1p3y = 1.0 + 3.0 * Y, _1p3z = 1.0 + 3.0 * Z;
    const double mxmy = _1mx * _1my, mxpy = _1mx * _1py, pxmy = _1px * _1my, pxpy = _1px * _1py;
    const double mxmz = _1mx * _1mz, mxpz = _1mx * _1pz, pxmz = _1px * _1mz, pxpz = _1px * _1pz;
    const double mymz = _1my * _1mz, mypz = _1my * _1pz, pymz = _1py * _1mz, pypz = _1py * _1pz;
    const double _1mx2 = 1.0 - x2, _1my2 = 1.0 - y2, _1mz2 = 1.0 - z2;
    const double facc = 1.0 / 64.0 * (9.0 * (x2 + y2 + z2) - 19.0);
    const double fx = 9.0 / 64.0 * _1mx2, fy = 9.0 / 64.0 * _1my2, fz = 9.0 / 64.0 * _1mz2;
    const double fxm = fx * _1m3x, fxp = fx * _1p3x, fym = fy * _1m3y, fyp = fy * _1p3y, fzm = fz * _1m3z, fzp = fz * _1p3z;
    const double A = 9.0 * (3.0 * x2 + y2 + z2) - 19.0, B = 9.0 * (x2 + 3.0 * y2 + z2) - 19.0, C = 9.0 * (x2 + y2 + 3.0 * z2) - 19.0;
    const double _18x = 18.0 * X, _18y = 18.0 * Y, _18z = 18.0 * Z;
    const double xmA = _18x - A, xpA = _18x + A, ymB = _18y - B, ypB = _18y + B, zmC = _18z - C, zpC = _18z + C;
    const double _3m9x2 = 3.0 - 9.0 * x2, _3m9y2 = 3.0 - 9.0 * y2, _3m9z2 = 3.0 - 9.0 * z2;
    const double _2x = 2.0 * X, _2y = 2.0 * Y, _2z = 2.0 * Z;
    const double mX = -_3m9x2 - _2x, pX = _3m9x2 - _2x, Xm = _1mx2 * _1m3x, Xp = _1mx2 * _1p3x;
    const double mY = -_3m9y2 - _2y, pY = _3m9y2 - _2y, Ym = _1my2 * _1m3y, Yp = _1my2 * _1p3y;
    const double mZ = -_3m9z2 - _2z, pZ = _3m9z2 - _2z, Zm = _1mz2 * _1m3z, Zp = _1mz2 * _1p3z;
    double* Nq = N + 32 * q;
    double* Dq = dN ? dN + 96 * q : nullptr;
    int j = 0;
#define DG_ROW(n, dx, dy, dz, SC)                                                            \
    {                                                                                        \
        Nq[j] = (n);                                                                         \
        if (Dq) { Dq[3 * j] = SC(dx); Dq[3 * j + 1] = SC(dy); Dq[3 * j + 2] = SC(dz); }      \
        j++;                                                                                 \
    }
#define DG_SC_C(v) ((v) * 0.015625)   /* == /64.0 exactly */
#define DG_SC_E(v) ((v) * (9.0 / 64.0))
    DG_ROW(facc * mxmy * _1mz, xmA * mymz, mxmz * ymB, mxmy * zmC, DG_SC_C)
    DG_ROW(facc * pxmy * _1mz, xpA * mymz, pxmz * ymB, pxmy * zmC, DG_SC_C)
    DG_ROW(facc * mxpy * _1mz, xmA * pymz, mxmz * ypB, mxpy * zmC, DG_SC_C)
    DG_ROW(facc * pxpy * _1mz, xpA * pymz, pxmz * ypB, pxpy * zmC, DG_SC_C)
    DG_ROW(facc * mxmy * _1pz, xmA * mypz, mxpz * ymB, mxmy * zpC, DG_SC_C)
    DG_ROW(facc * pxmy * _1pz, xpA * mypz, pxpz * ymB, pxmy * zpC, DG_SC_C)
    DG_ROW(facc * mxpy * _1pz, xmA * pypz, mxpz * ypB, mxpy * zpC, DG_SC_C)
    DG_ROW(facc * pxpy * _1pz, xpA * pypz, pxpz * ypB, pxpy * zpC, DG_SC_C)
    DG_ROW(fxm * mymz, mX * mymz, -Xm * _1mz, -Xm * _1my, DG_SC_E)
    DG_ROW(fxp * mymz, pX * mymz, -Xp * _1mz, -Xp * _1my, DG_SC_E)
    DG_ROW(fxm * mypz, mX * mypz, -Xm * _1pz, Xm * _1my, DG_SC_E)
    DG_ROW(fxp * mypz, pX * mypz, -Xp * _1pz, Xp * _1my, DG_SC_E)
    DG_ROW(fxm * pymz, mX * pymz, Xm * _1mz, -Xm * _1py, DG_SC_E)
    DG_ROW(fxp * pymz, pX * pymz, Xp * _1mz, -Xp * _1py, DG_SC_E)
    DG_ROW(fxm * pypz, mX * pypz, Xm * _1pz, Xm * _1py, DG_SC_E)
    DG_ROW(fxp * pypz, pX * pypz, Xp * _1pz, Xp * _1py, DG_SC_E)
    DG_ROW(fym * mxmz, -Ym * _1mz, mY * mxmz, -Ym * _1mx, DG_SC_E)
    DG_ROW(fyp * mxmz, -Yp * _1mz, pY * mxmz, -Yp * _1mx, DG_SC_E)
    DG_ROW(fym * pxmz, Ym * _1mz, mY * pxmz, -Ym * _1px, DG_SC_E)
    DG_ROW(fyp * pxmz, Yp * _1mz, pY * pxmz, -Yp * _1px, DG_SC_E)
    DG_ROW(fym * mxpz, -Ym * _1pz, mY * mxpz, Ym * _1mx, DG_SC_E)
    DG_ROW(fyp * mxpz, -Yp * _1pz, pY * mxpz, Yp * _1mx, DG_SC_E)
    DG_ROW(fym * pxpz, Ym * _1pz, mY * pxpz, Ym * _1px, DG_SC_E)
    DG_ROW(fyp * pxpz, Yp * _1pz, pY * pxpz, Yp * _1px, DG_SC_E)
    DG_ROW(fzm * mxmy, -Zm * _1my, -Zm * _1mx, mZ * mxmy, DG_SC_E)
    DG_ROW(fzp * mxmy, -Zp * _1my, -Zp * _1mx, pZ * mxmy, DG_SC_E)
    DG_ROW(fzm * mxpy, -Zm * _1py, Zm * _1mx, mZ * mxpy, DG_SC_E)
    DG_ROW(fzp * mxpy, -Zp * _1py, Zp * _1mx, pZ * mxpy, DG_SC_E)
    DG_ROW(fzm * pxmy, Zm * _1my, -Zm * _1px, mZ * pxmy, DG_SC_E)
    DG_ROW(fzp * pxmy, Zp * _1my, -Zp * _1px, pZ * pxmy, DG_SC_E)
    DG_ROW(fzm * pxpy, Zm * _1py, Zm * _1px, mZ * pxpy, DG_SC_E)
    DG_ROW(fzp * pxpy, Zp * _1py, Zp * _1px, pZ * pxpy, DG_SC_E)
#undef DG_ROW
#undef DG_SC_C
#undef DG_SC_E
}

}  // namespace

cudaError_t k2_launch_pack(const GridDev& g, const double* d_nodes, const unsigned* d_cells, uint64_t n_cells_kept,
                           double* d_packed, cudaStream_t stream)
{
    if (n_cells_kept == 0) return cudaSuccess;
    const unsigned long long n32 = (unsigned long long)n_cells_kept * 32ull;
    DG_KERNEL_LAUNCH(pack_cells_kernel, (unsigned)((n32 + 255) / 256), 256, 0, stream, g, d_nodes, d_cells, n32, d_packed);
    return DG_AFTER_LAUNCH();
}

cudaError_t k2_launch_axis_tables(const GridDev& g, double2* d_tab, cudaStream_t stream)
{
    const unsigned n = g.n[0] + g.n[1] + g.n[2];
    DG_KERNEL_LAUNCH(axis_tables_kernel, (n + 127) / 128, 128, 0, stream, g, d_tab);
    return DG_AFTER_LAUNCH();
}

cudaError_t k2_launch_interpolate(const FieldDev& f, const double* d_x, uint64_t n, double* d_phi, double* d_grad,
                                  cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + K2_THREADS - 1) / K2_THREADS);
    if (d_grad)
        DG_KERNEL_LAUNCH((interpolate_kernel<true>), blocks, K2_THREADS, 0, stream, f.g, f.packed, f.cell_map, f.tab, d_x, (unsigned long long)n, d_phi, d_grad);
    else
        DG_KERNEL_LAUNCH((interpolate_kernel<false>), blocks, K2_THREADS, 0, stream, f.g, f.packed, f.cell_map, f.tab, d_x, (unsigned long long)n, d_phi, nullptr);
    return DG_AFTER_LAUNCH();
}

cudaError_t k2_launch_shape_functions(const double* d_xi, uint64_t n, double* d_N, double* d_dN, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    DG_KERNEL_LAUNCH(shape_functions_kernel, (unsigned)((n + 127) / 128), 128, 0, stream, d_xi, (unsigned long long)n, d_N, d_dN);
    return DG_AFTER_LAUNCH();
}

}  // namespace dgb
