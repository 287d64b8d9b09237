// K4: the index passes of CubicLagrangeDiscreteGrid::reduceField (cubic_lagrange_discrete_grid.cpp:1065-1174) on the GPU -- SURVEY 8(f) N3.
//
// reduceField keeps the cells that own at least one node the predicate accepted, keeps the nodes those cells refer to, renumbers the
// nodes in Z-curve order and rewrites the connectivity.  The reference does it with one std::set of back-references per node; what it
// computes is a renumbering, and everything except two steps is data-parallel:
//   kernels here : keep flag per cell (:1081-1098) . cell map + compaction of the surviving rows . marks of the surviving nodes (:1117-1134)
//                  . Morton key of every survivor (zValue, :583-601 / z_sort_table.hpp:119-134) . new node ids, renumbered rows, gathered
//                  coefficients (:1158-1173)
//   host (reduce_field.cpp): the reference's back-to-front swap compaction replayed on a permutation array (:1135-1156, O(n), inherently
//                  serial) and the sort of the survivors by key, whose order among EQUAL keys is libstdc++'s introsort's (reduce_field.h).
// All kernels are warp-synchronous (ballot / popc), so tests/emu runs them unchanged.  Integer work, bit-exact by construction; the one
// floating-point step (the key's cast of inv * x) uses the reference's operations in the reference's order (no FMA: -fmad=false).
#include "k4_reduce.h"
#include "dg_launch.h"

#include <climits>

namespace dgb {
namespace {

constexpr unsigned TILE = 1024;                       // cells per warp in the counting / compaction passes

// keep flag per cell: any of its 32 nodes kept?  One warp per cell, lane j reads entry j.  bad[0] is set when an entry is not a node id.
__global__ void cell_keep_kernel(const unsigned* __restrict__ cells, unsigned long long n_cells, const unsigned char* __restrict__ keep_node,
                                 unsigned n_nodes, unsigned char* __restrict__ cell_keep, unsigned* __restrict__ bad)
{
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    const bool live = warp < n_cells;                 // whole warps: uniform
    unsigned v = 0;
    if (live) v = __ldg(cells + warp * 32ull + lane);
    const bool oob = live && v >= n_nodes;
    const unsigned any_bad = __ballot_sync(0xffffffffu, oob);
    const unsigned any_keep = __ballot_sync(0xffffffffu, live && !oob && keep_node[v] != 0);
    if (live && lane == 0) {
        if (any_bad) atomicAdd(bad, 1u);
        cell_keep[warp] = (any_keep && !any_bad) ? 1 : 0;
    }
}

// number of set flags per tile of TILE entries (one warp per tile)
__global__ void tile_count_kernel(const unsigned char* __restrict__ flag, unsigned long long n, unsigned* __restrict__ tile_count)
{
    const unsigned long long tile = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long base = tile * TILE;
    if (base >= n) return;                            // whole warps
    unsigned cnt = 0;
    for (unsigned k = 0; k < TILE; k += 32) {
        const unsigned long long i = base + k + lane;
        cnt += (unsigned)__popc(__ballot_sync(0xffffffffu, i < n && flag[i] != 0));
    }
    if (lane == 0) tile_count[tile] = cnt;
}

// cell_map (:1076-1097) and the surviving rows moved to the front, order kept: one warp per tile, tile_offset = exclusive scan of the counts
__global__ void cell_compact_kernel(const unsigned* __restrict__ cells, unsigned long long n_cells, const unsigned char* __restrict__ cell_keep,
                                    const unsigned* __restrict__ tile_offset, unsigned* __restrict__ cells_out, unsigned* __restrict__ cell_map)
{
    const unsigned long long tile = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long base = tile * TILE;
    if (base >= n_cells) return;
    unsigned run = tile_offset[tile];
    const unsigned lt = (1u << lane) - 1u;
    for (unsigned k = 0; k < TILE; k += 32) {
        const unsigned long long c = base + k + lane;
        const bool kept = c < n_cells && cell_keep[c] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, kept);
        if (c < n_cells) cell_map[c] = kept ? run + (unsigned)__popc(m & lt) : 0xffffffffu;
        // the warp copies the kept rows of this chunk, 128 bytes at a time
        unsigned rest = m, r = run;
        while (rest) {
            const unsigned src_lane = (unsigned)__ffs((int)rest) - 1u;
            rest &= rest - 1u;
            const unsigned long long src = base + k + src_lane;
            cells_out[(unsigned long long)r * 32ull + lane] = __ldg(cells + src * 32ull + lane);
            r++;
        }
        run += (unsigned)__popc(m);
    }
}

__global__ void mark_used_kernel(const unsigned* __restrict__ cells_out, unsigned long long n_entries, unsigned char* __restrict__ used)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_entries) used[cells_out[i]] = 1;
}

// morton_lut (z_sort_table.hpp:119-134): only the LOW 16 bits of each coordinate reach the key (x at bit 0, y at bit 1, z at bit 2)
__device__ __forceinline__ unsigned long long spread16(unsigned long long v)
{
    v &= 0xffffull;
    v = (v | v << 16) & 0x0000ff0000ffull;
    v = (v | v << 8) & 0x00f00f00f00full;
    v = (v | v << 4) & 0x0c30c30c30c3ull;
    v = (v | v << 2) & 0x249249249249ull;
    return v;
}

// key[i] = zValue(indexToNodePosition(perm[i]), 4 * min(inv_cell_size))   (:1113-1114, :583-601)
__global__ void morton_key_kernel(GridDev g, const unsigned* __restrict__ perm, unsigned long long m, double inv, unsigned long long* __restrict__ key)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double x[3];
    node_position(g, perm[i], x[0], x[1], x[2]);
    unsigned p[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int c = (x[k] >= 0.0) ? static_cast<int>(inv * x[k]) : static_cast<int>(inv * x[k]) - 1;             // :589-592
        p[k] = static_cast<unsigned>(static_cast<long long>(c) - (static_cast<long long>(INT_MIN) + 1));          // :595-598
    }
    key[i] = spread16(p[0]) | (spread16(p[1]) << 1) | (spread16(p[2]) << 2);
}

// rank r holds the node that sat at position order[r]: new id of that node = r; its coefficient moves to r (:1158-1173)
__global__ void renumber_nodes_kernel(const unsigned* __restrict__ perm, const unsigned* __restrict__ order, unsigned long long m,
                                      const double* __restrict__ nodes, unsigned* __restrict__ new_id, double* __restrict__ nodes_out)
{
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const unsigned v = perm[order[r]];
    new_id[v] = (unsigned)r;
    nodes_out[r] = nodes[v];
}

__global__ void renumber_cells_kernel(unsigned* __restrict__ cells_out, unsigned long long n_entries, const unsigned* __restrict__ new_id)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_entries) cells_out[i] = new_id[cells_out[i]];
}

inline unsigned blocks_for(unsigned long long threads, unsigned block) { return (unsigned)((threads + block - 1) / block); }

}  // namespace

unsigned k4_tile() { return TILE; }

cudaError_t k4_launch_cell_keep(const unsigned* d_cells, uint64_t n_cells, const unsigned char* d_keep_node, unsigned n_nodes, unsigned char* d_cell_keep,
                                unsigned* d_bad, cudaStream_t stream)
{
    if (n_cells == 0) return cudaSuccess;
    DG_KERNEL_LAUNCH(cell_keep_kernel, blocks_for(n_cells * 32ull, 128), 128, 0, stream, d_cells, (unsigned long long)n_cells, d_keep_node, n_nodes, d_cell_keep, d_bad);
    return DG_AFTER_LAUNCH();
}

cudaError_t k4_launch_tile_count(const unsigned char* d_flag, uint64_t n, unsigned* d_tile_count, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const uint64_t tiles = (n + TILE - 1) / TILE;
    DG_KERNEL_LAUNCH(tile_count_kernel, blocks_for(tiles * 32ull, 128), 128, 0, stream, d_flag, (unsigned long long)n, d_tile_count);
    return DG_AFTER_LAUNCH();
}

cudaError_t k4_launch_cell_compact(const unsigned* d_cells, uint64_t n_cells, const unsigned char* d_cell_keep, const unsigned* d_tile_offset,
                                   unsigned* d_cells_out, unsigned* d_cell_map, cudaStream_t stream)
{
    if (n_cells == 0) return cudaSuccess;
    const uint64_t tiles = (n_cells + TILE - 1) / TILE;
    DG_KERNEL_LAUNCH(cell_compact_kernel, blocks_for(tiles * 32ull, 128), 128, 0, stream, d_cells, (unsigned long long)n_cells, d_cell_keep, d_tile_offset, d_cells_out, d_cell_map);
    return DG_AFTER_LAUNCH();
}

cudaError_t k4_launch_mark_used(const unsigned* d_cells_out, uint64_t n_entries, unsigned char* d_used, cudaStream_t stream)
{
    if (n_entries == 0) return cudaSuccess;
    DG_KERNEL_LAUNCH(mark_used_kernel, blocks_for(n_entries, 256), 256, 0, stream, d_cells_out, (unsigned long long)n_entries, d_used);
    return DG_AFTER_LAUNCH();
}

cudaError_t k4_launch_morton_keys(const GridDev& g, const unsigned* d_perm, uint64_t m, unsigned long long* d_key, cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    double inv = g.inv[0]; if (g.inv[1] < inv) inv = g.inv[1]; if (g.inv[2] < inv) inv = g.inv[2];
    inv = 4.0 * inv;                                                                                    // 4.0 * m_inv_cell_size.minCoeff(), :1114
    DG_KERNEL_LAUNCH(morton_key_kernel, blocks_for(m, 256), 256, 0, stream, g, d_perm, (unsigned long long)m, inv, d_key);
    return DG_AFTER_LAUNCH();
}

cudaError_t k4_launch_renumber(const unsigned* d_perm, const unsigned* d_order, uint64_t m, const double* d_nodes, unsigned* d_new_id, double* d_nodes_out,
                               unsigned* d_cells_out, uint64_t n_entries, cudaStream_t stream)
{
    if (m) { DG_KERNEL_LAUNCH(renumber_nodes_kernel, blocks_for(m, 256), 256, 0, stream, d_perm, d_order, (unsigned long long)m, d_nodes, d_new_id, d_nodes_out); }
    if (n_entries) { DG_KERNEL_LAUNCH(renumber_cells_kernel, blocks_for(n_entries, 256), 256, 0, stream, d_cells_out, (unsigned long long)n_entries, d_new_id); }
    return DG_AFTER_LAUNCH();
}

}  // namespace dgb
