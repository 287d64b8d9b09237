// K1: nearest-triangle query + pseudonormal sign, one query per lane, exact reference traversal order.
//
// Replaces TriangleMeshDistance::signed_distance / unsigned_distance / _query / point_triangle_sq_unsigned
// (TriangleMeshDistance.h:269-328, 514-562, 564-820) and, fused in front of it, indexToNodePosition
// (cubic_lagrange_discrete_grid.cpp:604-665) for the addFunction node loop (:806-817).
//
// Why per-lane depth-first order and not a warp-shared traversal: _query keeps the FIRST strictly smaller
// triangle in nearer-child-first order and prunes with the running best (TriangleMeshDistance.h:528, 542-560),
// so both the winner among equidistant triangles (hence the sign) and, through floating-point cancellation in
// d2, the last bits of the distance depend on each query's own visit order.  Bit-exact parity therefore needs
// each lane to walk its own order.  What the warp shares instead is the code path (one loop, leaf tests
// batched behind a ballot) and the caches: the 32 lanes of a warp are 32 adjacent grid nodes, so they fetch
// mostly the same 64-byte sphere pairs and 128-byte triangle records.
//
// Data (see bvh_build.h): implicit tree over leaf ranges [b,e), sphere pair at spheres[(b+e)>>1], one
// 128-byte record per triangle, one 7x3 pseudonormal block per triangle.  The per-lane stack of deferred
// siblings (range + its sphere distance) lives in shared memory, laid out [depth][lane] so it is
// bank-conflict-free whatever depth each lane is at.
#include "dg_device.cuh"
#include "bvh_build.h"
#include "k1_sdf.h"

#include <cfloat>

namespace dgb {

namespace {

struct QueryResult {
    double dist;        // unsigned distance (sqrt of best d2)
    double s, t;        // barycentric parameters of the nearest point on the winning triangle
    int pos;            // leaf position of the winning triangle (-1 if none)
    int entity;         // dg_nearest_entity
};

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// point_triangle_sq_unsigned (TriangleMeshDistance.h:564-820) on a precomputed record.  Returns d2 and the
// (s, t, entity) the reference would use for nearest_point (:818); the point itself is only needed for the
// final winner, so it is reconstructed after the traversal.
__device__ __forceinline__ double tri_dist2(const LeafRecord* __restrict__ rec, double px, double py, double pz,
                                            double& s_out, double& t_out, int& ent_out)
{
    const double* r = reinterpret_cast<const double*>(rec);
    const double2 q0 = ldg2(r + 0), q1 = ldg2(r + 2), q2 = ldg2(r + 4), q3 = ldg2(r + 6);
    const double2 q4 = ldg2(r + 8), q5 = ldg2(r + 10), q6 = ldg2(r + 12), q7 = ldg2(r + 14);
    const double v0x = q0.x, v0y = q0.y, v0z = q1.x;
    const double e0x = q1.y, e0y = q2.x, e0z = q2.y;
    const double e1x = q3.x, e1y = q3.y, e1z = q4.x;
    const double a00 = q4.y, a01 = q5.x, a11 = q5.y;
    const double det = q6.x, inv_det = q6.y, denom = q7.x;

    const double dx = v0x - px, dy = v0y - py, dz = v0z - pz;          // diff = v0 - point  (:566)
    const double b0 = dx * e0x + dy * e0y + dz * e0z;                  // :572
    const double b1 = dx * e1x + dy * e1y + dz * e1z;                  // :573
    const double c = dx * dx + dy * dy + dz * dz;                      // :574
    double s = a01 * b1 - a11 * b0;                                    // :576
    double t = a01 * b0 - a00 * b1;                                    // :577
    double d2;
    int ent;
    // the three vertex / three edge outcomes, written once (each appears several times in the reference)
#define DG_V0() { ent = 0; s = 0; t = 0; d2 = c; }
#define DG_V1() { ent = 1; s = 1; t = 0; d2 = a00 + 2 * b0 + c; }
#define DG_V2() { ent = 2; s = 0; t = 1; d2 = a11 + 2 * b1 + c; }
#define DG_E01() { ent = 3; t = 0; s = -b0 / a00; d2 = b0 * s + c; }
#define DG_E02() { ent = 5; s = 0; t = -b1 / a11; d2 = b1 * t + c; }
#define DG_QUAD() (s * (a00 * s + a01 * t + 2 * b0) + t * (a01 * s + a11 * t + 2 * b1) + c)
    if (s + t <= det) {
        if (s < 0) {
            if (t < 0) {                                // region 4 (:585-625)
                if (b0 < 0) { if (-b0 >= a00) DG_V1() else DG_E01() }
                else { if (b1 >= 0) DG_V0() else if (-b1 >= a11) DG_V2() else DG_E02() }
            } else {                                    // region 3 (:626-647)
                if (b1 >= 0) DG_V0() else if (-b1 >= a11) DG_V2() else DG_E02()
            }
        } else if (t < 0) {                             // region 5 (:649-670)
            if (b0 >= 0) DG_V0() else if (-b0 >= a00) DG_V1() else DG_E01()
        } else {                                        // region 0 (:671-680)
            ent = 6;
            s *= inv_det; t *= inv_det;
            d2 = DG_QUAD();
        }
    } else {
        if (s < 0) {                                    // region 2 (:686-732)
            const double tmp0 = a01 + b0, tmp1 = a11 + b1;
            if (tmp1 > tmp0) {
                const double numer = tmp1 - tmp0;
                if (numer >= denom) DG_V1()
                else { ent = 4; s = numer / denom; t = 1 - s; d2 = DG_QUAD(); }
            } else {
                if (tmp1 <= 0) DG_V2() else if (b1 >= 0) DG_V0() else DG_E02()
            }
        } else if (t < 0) {                             // region 6 (:733-779)
            const double tmp0 = a01 + b1, tmp1 = a00 + b0;
            if (tmp1 > tmp0) {
                const double numer = tmp1 - tmp0;
                if (numer >= denom) DG_V2()
                else { ent = 4; t = numer / denom; s = 1 - t; d2 = DG_QUAD(); }
            } else {
                if (tmp1 <= 0) DG_V1() else if (b0 >= 0) DG_V0() else DG_E01()
            }
        } else {                                        // region 1 (:780-809)
            const double numer = a11 + b1 - a01 - b0;
            if (numer <= 0) DG_V2()
            else if (numer >= denom) DG_V1()
            else { ent = 4; s = numer / denom; t = 1 - s; d2 = DG_QUAD(); }
        }
    }
#undef DG_V0
#undef DG_V1
#undef DG_V2
#undef DG_E01
#undef DG_E02
#undef DG_QUAD
    if (d2 < 0) d2 = 0;                                 // :812-816
    s_out = s; t_out = t; ent_out = ent;
    return d2;
}

// _query (TriangleMeshDistance.h:514-562), iterative and WARP-SYNCHRONOUS: all 32 lanes of the warp call this
// together (`alive` = lane has a query).  Every lane walks its own reference order with its own stack, but the warp
// executes one phase at a time, chosen by ballot: an "internal" phase (two sphere tests, push / descend / prune) for
// the lanes sitting at an internal node, or a "leaf" phase (point-triangle test, accept, pop) for the lanes sitting at
// a leaf.  The phase with more (cost-weighted) lanes runs; the others idle for that iteration.  This keeps the
// expensive straight-line blocks converged instead of letting 32 private loops drift apart.
// stack_rng/stack_d point at this lane's column of the block's shared-memory stack ([depth][lane], conflict-free).
__device__ __forceinline__ QueryResult nearest_triangle(const SpherePair* __restrict__ spheres,
                                                        const LeafRecord* __restrict__ leaves, int n_tri, bool alive,
                                                        double px, double py, double pz,
                                                        uint2* stack_rng, double* stack_d, int stride)
{
    QueryResult res;
    res.dist = DBL_MAX; res.s = 0; res.t = 0; res.pos = -1; res.entity = 0;
    double best = DBL_MAX;             // result.distance
    double best_sq = best * best;      // result.distance * result.distance (= +inf initially), :528
    int b = 0, e = n_tri, sp = 0;
    // pops deferred siblings until one passes the reference's second test `d < result.distance` (:549, :557)
    auto pop = [&]() {
        for (;;) {
            if (sp == 0) { alive = false; return; }
            sp--;
            if (stack_d[sp * stride] < best) {
                const uint2 r = stack_rng[sp * stride];
                b = (int)r.x; e = (int)r.y;
                return;
            }
        }
    };
    for (;;) {
        const bool at_leaf = alive && (e - b == 1);
        const bool at_node = alive && (e - b > 1);
        const unsigned m_leaf = __ballot_sync(0xffffffffu, at_leaf);
        const unsigned m_node = __ballot_sync(0xffffffffu, at_node);
        if ((m_leaf | m_node) == 0u) break;
        if (K1_NODE_WEIGHT * __popc(m_node) >= K1_LEAF_WEIGHT * __popc(m_leaf)) {
            if (at_node) {                                                      // internal (:537-561)
                const int m = (b + e) >> 1;
                const double* sp8 = reinterpret_cast<const double*>(spheres + m);
                const double2 a0 = ldg2(sp8), a1 = ldg2(sp8 + 2), a2 = ldg2(sp8 + 4), a3 = ldg2(sp8 + 6);
                const double lx = px - a0.x, ly = py - a0.y, lz = pz - a1.x;
                const double rx = px - a2.x, ry = py - a2.y, rz = pz - a3.x;
                const double d_left = sqrt(lx * lx + ly * ly + lz * lz) - a1.y;     // :539
                const double d_right = sqrt(rx * rx + ry * ry + rz * rz) - a3.y;    // :540
                const bool left_first = d_left < d_right;                           // :542
                const double d_first = left_first ? d_left : d_right;
                const double d_second = left_first ? d_right : d_left;
                const int fb = left_first ? b : m, fe = left_first ? m : e;
                const int sb = left_first ? m : b, se = left_first ? e : m;
                if (d_first < best) {                  // visit first now; second is re-tested when popped (:545-551)
                    stack_rng[sp * stride] = make_uint2((unsigned)sb, (unsigned)se);
                    stack_d[sp * stride] = d_second;
                    sp++;
                    b = fb; e = fe;
                } else if (d_second < best) {          // only reachable through NaNs; kept for fidelity
                    b = sb; e = se;
                } else {
                    pop();
                }
            }
        } else {
            if (at_leaf) {                                                      // leaf (:517-534)
                double s, t; int ent;
                const double d2 = tri_dist2(leaves + b, px, py, pz, s, t, ent);
                if (d2 < best_sq) {
                    best = sqrt(d2);
                    best_sq = best * best;
                    res.s = s; res.t = t; res.pos = b; res.entity = ent;
                }
                pop();
            }
        }
    }
    res.dist = best;
    return res;
}

// nearest_point (TriangleMeshDistance.h:818) and the pseudonormal sign (:274-305)
__device__ __forceinline__ void finish_query(const LeafRecord* __restrict__ leaves, const PseudoNormals* __restrict__ normals,
                                             const QueryResult& r, double px, double py, double pz, bool is_signed,
                                             double& dist, double& qx, double& qy, double& qz, int& tri_id)
{
    if (r.pos < 0) {            // nothing accepted (NaN input); the reference would index triangles[-1]
        dist = r.dist; qx = qy = qz = 0.0; tri_id = -1;
        return;
    }
    const double* rec = reinterpret_cast<const double*>(leaves + r.pos);
    const double v0x = __ldg(rec + 0), v0y = __ldg(rec + 1), v0z = __ldg(rec + 2);
    const double e0x = __ldg(rec + 3), e0y = __ldg(rec + 4), e0z = __ldg(rec + 5);
    const double e1x = __ldg(rec + 6), e1y = __ldg(rec + 7), e1z = __ldg(rec + 8);
    tri_id = __ldg(&leaves[r.pos].tri_id);
    qx = v0x + r.s * e0x + r.t * e1x;                  // v0 + s*edge0 + t*edge1
    qy = v0y + r.s * e0y + r.t * e1y;
    qz = v0z + r.s * e0z + r.t * e1z;
    dist = r.dist;
    if (is_signed) {
        const double* n = normals[r.pos].n[r.entity];
        const double nx = __ldg(n), ny = __ldg(n + 1), nz = __ldg(n + 2);
        const double ux = px - qx, uy = py - qy, uz = pz - qz;
        const double d = ux * nx + uy * ny + uz * nz;
        dist = dist * ((d >= 0.0) ? 1.0 : -1.0);       // :305
    }
}

extern __shared__ __align__(16) unsigned char k1_smem[];

// addFunction node loop: out[l - l_begin] = sign * signed_distance(indexToNodePosition(l)).distance.
// Thread mapping: the node array is four row-major 3-D arrays (vertices, x-, y-, z-edge nodes; K1Segment).  A warp
// owns a 4 x 4 x 2 brick (fast x mid x slow) of one of them and a 128-thread block four bricks side by side along the
// fast axis, so the 32 queries of a warp are spatial neighbours: they visit nearly the same tree nodes (L1 hits) and
// take similar numbers of steps.  Whole slow-planes are covered; nodes outside [l_begin, l_end) are masked.
__global__ void __launch_bounds__(K1_THREADS)
sdf_sample_nodes_kernel(const SpherePair* __restrict__ spheres, const LeafRecord* __restrict__ leaves,
                        const PseudoNormals* __restrict__ normals, int n_tri, int stack_depth, GridDev g, K1Work w, double sign,
                        double* __restrict__ out)
{
    double* stack_d = reinterpret_cast<double*>(k1_smem);
    uint2* stack_rng = reinterpret_cast<uint2*>(k1_smem + (size_t)stack_depth * K1_THREADS * sizeof(double));
    // which segment does this block belong to (<= 4, uniform per block)
    int sg = 0;
#pragma unroll
    for (int k = 1; k < 4; k++) if (k < w.nseg && blockIdx.x >= w.seg[k].block_begin) sg = k;
    const K1Segment& S = w.seg[sg];
    unsigned t = blockIdx.x - S.block_begin;
    const unsigned tf = t % S.tiles_f; t /= S.tiles_f;
    const unsigned tm = t % S.tiles_m; const unsigned ts = t / S.tiles_m;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned f = tf * 16u + warp * 4u + (lane & 3u);
    const unsigned m = tm * 4u + ((lane >> 2) & 3u);
    const unsigned sl = S.s0 + ts * 2u + (lane >> 4);
    const unsigned l = S.l_base + (sl * S.Dm + m) * S.Df + f;
    const bool alive = (f < S.Df) && (m < S.Dm) && (sl < S.s1) && (l >= w.l_begin) && (l < w.l_end);

    // indexToNodePosition (cubic_lagrange_discrete_grid.cpp:604-665) from the array coordinates
    unsigned i, j, k;
    int axis = S.kind - 1;                  // -1 vertex, 0/1/2 edge axis
    const unsigned par = f & 1u, fh = f >> 1;
    if (S.kind == 0) { i = f; j = m; k = sl; }
    else if (S.kind == 1) { i = fh; j = m; k = sl; }
    else if (S.kind == 2) { i = sl; k = m; j = fh; }
    else { j = sl; i = m; k = fh; }
    double px = g.mn[0] + g.cell[0] * (double)i;
    double py = g.mn[1] + g.cell[1] * (double)j;
    double pz = g.mn[2] + g.cell[2] * (double)k;
    const double fr = (1.0 + (double)par) / 3.0;
    if (axis == 0) px = px + fr * g.cell[0];
    else if (axis == 1) py = py + fr * g.cell[1];
    else if (axis == 2) pz = pz + fr * g.cell[2];

    const QueryResult r = nearest_triangle(spheres, leaves, n_tri, alive, px, py, pz, stack_rng + threadIdx.x,
                                           stack_d + threadIdx.x, K1_THREADS);
    if (!alive) return;
    double dist, qx, qy, qz; int tri;
    finish_query(leaves, normals, r, px, py, pz, true, dist, qx, qy, qz, tri);
    out[l - w.l_begin] = (sign == 1.0) ? dist : sign * dist;     // cmd/generate_sdf/main.cpp:97 (-1.0 * d) / :101
}

// batched TriangleMeshDistance::{signed,unsigned}_distance on arbitrary points
__global__ void __launch_bounds__(K1_THREADS)
mesh_distance_kernel(const SpherePair* __restrict__ spheres, const LeafRecord* __restrict__ leaves,
                     const PseudoNormals* __restrict__ normals, int n_tri, int stack_depth,
                     const double* __restrict__ pts, unsigned long long count, int is_signed,
                     double* __restrict__ dist_out, double* __restrict__ near_out, int* __restrict__ ent_out,
                     int* __restrict__ tri_out)
{
    double* stack_d = reinterpret_cast<double*>(k1_smem);
    uint2* stack_rng = reinterpret_cast<uint2*>(k1_smem + (size_t)stack_depth * K1_THREADS * sizeof(double));
    const unsigned long long idx = (unsigned long long)blockIdx.x * K1_THREADS + threadIdx.x;
    const bool alive = idx < count;
    const unsigned long long ix = alive ? idx : 0ull;
    const double px = pts[3 * ix], py = pts[3 * ix + 1], pz = pts[3 * ix + 2];
    const QueryResult r = nearest_triangle(spheres, leaves, n_tri, alive, px, py, pz, stack_rng + threadIdx.x,
                                           stack_d + threadIdx.x, K1_THREADS);
    if (!alive) return;
    double dist, qx, qy, qz; int tri;
    finish_query(leaves, normals, r, px, py, pz, is_signed != 0, dist, qx, qy, qz, tri);
    if (dist_out) dist_out[idx] = dist;
    if (near_out) { near_out[3 * idx] = qx; near_out[3 * idx + 1] = qy; near_out[3 * idx + 2] = qz; }
    if (ent_out) ent_out[idx] = r.entity;
    if (tri_out) tri_out[idx] = tri;
}

__global__ void node_positions_kernel(GridDev g, unsigned l_begin, unsigned long long count, double* __restrict__ x)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    double px, py, pz;
    node_position(g, l_begin + (unsigned)idx, px, py, pz);
    x[3 * idx] = px; x[3 * idx + 1] = py; x[3 * idx + 2] = pz;
}

// closed-form cell table (cubic_lagrange_discrete_grid.cpp:833-886): one thread per (cell, local node)
__global__ void build_cells_kernel(GridDev g, unsigned c_begin, unsigned long long count32, unsigned* __restrict__ cells)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count32) return;
    const unsigned c = c_begin + (unsigned)(idx >> 5), jn = (unsigned)(idx & 31u);
    const unsigned nxy = g.n[0] * g.n[1];
    const unsigned k = c / nxy, t = c - k * nxy, j = t / g.n[0], i = t - j * g.n[0];
    cells[idx] = cell_node_id(g, i, j, k, jn);
}

// FMA-contraction probe: a*b+c with operands for which the fused and unfused roundings differ.
__global__ void fma_probe_kernel(double a, double b, double c, double* out) { out[0] = a * b + c; }

}  // namespace

static inline size_t k1_smem_bytes(int stack_depth) { return (size_t)stack_depth * K1_THREADS * (sizeof(double) + sizeof(uint2)); }

cudaError_t k1_configure(int stack_depth)
{
    const size_t bytes = k1_smem_bytes(stack_depth);
    cudaError_t e = cudaFuncSetAttribute(sdf_sample_nodes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(mesh_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Splits [l_begin, l_begin+count) into the (at most four) node arrays it touches and tiles whole slow-planes of each.
cudaError_t k1_launch_sample_nodes(const DeviceBvh& m, const GridDev& g, double sign, uint64_t l_begin, uint64_t count,
                                   double* d_out, cudaStream_t stream)
{
    if (count == 0) return cudaSuccess;
    const uint64_t l_end = l_begin + count;
    const unsigned nx = g.n[0], ny = g.n[1], nz = g.n[2];
    // {base, Ds, Dm, Df} of the four row-major node arrays (cubic_lagrange_discrete_grid.cpp:618-662)
    const uint64_t base[5] = {0, g.nv, (uint64_t)g.nv + 2ull * g.ne_x, (uint64_t)g.nv + 2ull * (g.ne_x + (uint64_t)g.ne_y),
                              (uint64_t)g.nv + 2ull * ((uint64_t)g.ne_x + g.ne_y + g.ne_z)};
    const unsigned dims[4][3] = {{nz + 1, ny + 1, nx + 1}, {nz + 1, ny + 1, 2 * nx}, {nx + 1, nz + 1, 2 * ny}, {ny + 1, nx + 1, 2 * nz}};
    K1Work w;
    w.nseg = 0; w.l_begin = (unsigned)l_begin; w.l_end = (unsigned)l_end;
    unsigned blocks = 0;
    for (int k = 0; k < 4; k++) {
        const uint64_t a = l_begin > base[k] ? l_begin : base[k];
        const uint64_t b = l_end < base[k + 1] ? l_end : base[k + 1];
        if (a >= b) continue;
        K1Segment& S = w.seg[w.nseg++];
        S.kind = k; S.l_base = (unsigned)base[k];
        S.Ds = dims[k][0]; S.Dm = dims[k][1]; S.Df = dims[k][2];
        const uint64_t plane = (uint64_t)S.Dm * S.Df;
        S.s0 = (unsigned)((a - base[k]) / plane);
        S.s1 = (unsigned)((b - 1 - base[k]) / plane) + 1;
        S.tiles_f = (S.Df + 15) / 16; S.tiles_m = (S.Dm + 3) / 4;
        const unsigned tiles_s = (S.s1 - S.s0 + 1) / 2;
        S.block_begin = blocks;
        blocks += S.tiles_f * S.tiles_m * tiles_s;
    }
    for (int k = w.nseg; k < 4; k++) { w.seg[k] = w.seg[0]; w.seg[k].block_begin = 0xffffffffu; }
    sdf_sample_nodes_kernel<<<blocks, K1_THREADS, k1_smem_bytes(m.stack_depth), stream>>>(
        m.spheres, m.leaves, m.normals, m.n_tri, m.stack_depth, g, w, sign, d_out);
    return cudaGetLastError();
}

cudaError_t k1_launch_distance(const DeviceBvh& m, const double* d_pts, uint64_t count, int is_signed, double* d_dist,
                               double* d_near, int* d_ent, int* d_tri, cudaStream_t stream)
{
    if (count == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((count + K1_THREADS - 1) / K1_THREADS);
    mesh_distance_kernel<<<blocks, K1_THREADS, k1_smem_bytes(m.stack_depth), stream>>>(
        m.spheres, m.leaves, m.normals, m.n_tri, m.stack_depth, d_pts, (unsigned long long)count, is_signed, d_dist, d_near,
        d_ent, d_tri);
    return cudaGetLastError();
}

cudaError_t k1_launch_node_positions(const GridDev& g, uint64_t l_begin, uint64_t count, double* d_x, cudaStream_t stream)
{
    if (count == 0) return cudaSuccess;
    node_positions_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(g, (unsigned)l_begin, (unsigned long long)count, d_x);
    return cudaGetLastError();
}

cudaError_t k1_launch_build_cells(const GridDev& g, uint64_t c_begin, uint64_t count, unsigned* d_cells, cudaStream_t stream)
{
    if (count == 0) return cudaSuccess;
    const unsigned long long n32 = (unsigned long long)count * 32ull;
    build_cells_kernel<<<(unsigned)((n32 + 255) / 256), 256, 0, stream>>>(g, (unsigned)c_begin, n32, d_cells);
    return cudaGetLastError();
}

cudaError_t k1_launch_fma_probe(double a, double b, double c, double* d_out, cudaStream_t stream)
{
    fma_probe_kernel<<<1, 1, 0, stream>>>(a, b, c, d_out);
    return cudaGetLastError();
}

}  // namespace dgb
