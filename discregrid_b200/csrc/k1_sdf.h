// Launch interface of the K1 kernels (k1_sdf.cu).  Host-callable, no kernel code here.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "bvh_build.h"
#include "dg_device.cuh"

namespace dgb {

#ifndef K1_THREADS_PER_BLOCK
#define K1_THREADS_PER_BLOCK 64
#endif
constexpr int K1_THREADS = K1_THREADS_PER_BLOCK;
// phase choice of the warp-synchronous traversal: the phase with the largest weight * #lanes runs
#ifndef K1_NODE_WEIGHT
#define K1_NODE_WEIGHT 2
#endif
#ifndef K1_LEAF_WEIGHT
#define K1_LEAF_WEIGHT 3
#endif
#ifndef K1_POP_WEIGHT
#define K1_POP_WEIGHT 4
#endif
#ifndef K1_POP_TRIES
#define K1_POP_TRIES 4          // deferred siblings re-tested per POP phase and lane
#endif
#ifndef K1_POP_MERGED
#define K1_POP_MERGED 0         // 1: POP lanes re-test siblings at the end of every iteration; 0: POP is a phase of its own
#endif
#ifndef K1_LEAF_MODE
#define K1_LEAF_MODE 0          // leaf arithmetic: 0 compiler's branches, 1 fully straight-line, 2 one guarded division + one guarded quadratic form
#endif
#ifndef K1_FILTER
#define K1_FILTER 1             // 1: fp32 interval filter for the sphere decisions (exact fp64 fallback); 0: all fp64
#endif
#ifndef K1_LEAF_FILTER
#define K1_LEAF_FILTER 0         // 1: reject leaf tests whose certified fp32 lower bound exceeds the best, as a phase of its own.
                                // Exact (all parity tests pass) and rejects 83-90 % of the leaf tests, but measured SLOWER
                                // (67.3 vs 54.6 ms): the extra phase costs more iterations than the fp64 tests it saves.  Off.
#endif
#ifndef K1_LEAFF_WEIGHT
#define K1_LEAFF_WEIGHT 2
#endif
// brick of nodes owned by one warp: fast x mid x slow = 32
#ifndef K1_BRICK_F
#define K1_BRICK_F 4
#endif
#ifndef K1_BRICK_M
#define K1_BRICK_M 4
#endif
#ifndef K1_BRICK_S
#define K1_BRICK_S 2
#endif
static_assert(K1_BRICK_F * K1_BRICK_M * K1_BRICK_S == 32, "a brick is one warp");
#ifndef K1_NODE_REPEAT
#define K1_NODE_REPEAT 0         // >0: up to this many consecutive node steps per phase vote (while >= half of the lanes stay at internal nodes)
#endif
#ifndef K1_PREFETCH
#define K1_PREFETCH 0            // prefetch.global.L1 of both children's records during a node step
#endif
#ifndef K1_NODEF_STRIDE
#define K1_NODEF_STRIDE 5         // float4s per fp32 node record: 5 = packed 80 B, 6 = padded to 96 B
#endif
#ifndef K1_EARLY_BOX
#define K1_EARLY_BOX 1           // issue the box loads together with the sphere loads (latency) instead of after the sphere decision
#endif
#ifndef K1_BOX_SKIP
#define K1_BOX_SKIP 1           // skip subtrees whose box is certainly farther than the best (they cannot change the result)
#endif
#ifndef K1_SKIP_HOPELESS
#define K1_SKIP_HOPELESS 1      // do not stack a sibling whose sphere is already certainly farther than the best
#endif
#ifndef K1_MIN_BLOCKS
#define K1_MIN_BLOCKS (1024 / K1_THREADS_PER_BLOCK)   // blocks per SM the register allocation must allow (32 warps)
#endif

// K1_PACKET 1 (default since round 2, profiles/r2s - r2v): the 32 queries of a brick walk the tree TOGETHER (one shared order, certified fp32
// pruning) and the reference's order-dependent accept rule is replayed afterwards on the few near-minimum triangles; lanes whose checks fail
// are walked again per lane (k1_sdf.cu, nearest_triangle_packet).  0: one query per lane in the reference's own order for the whole walk
// (the round-1 kernel; still what mesh_distance_kernel -- arbitrary points, no bricks -- and the fallback run).
//   K1_PKT_K          near-minimum candidates kept per query (8: K = 6 falls back on bunny.obj's valence-7/8 vertices, 50.3 vs 37.3 ms)
//   K1_PKT_MIN_BLOCKS blocks per SM the register allocation of the node-loop kernel must allow
#ifndef K1_PACKET
#define K1_PACKET 1
#endif
#ifndef K1_PKT_MIN_BLOCKS
#define K1_PKT_MIN_BLOCKS 14
#endif
#define K1_NEEDS_LEAF_SHADOW (K1_LEAF_FILTER || K1_PACKET)     // the fp32 triangle shadows (LeafF) are built and uploaded
// K1_WAVE 1: the node-loop kernel is the WAVEFRONT variant (k1_sdf.cu): persistent warps, a pool of K1_WAVE_SLOTS query slots per warp in
// shared memory, per iteration the fullest phase is compacted onto the lanes by ballot.  0: the per-lane kernel (one query per lane).
#ifndef K1_WAVE
#define K1_WAVE 0
#endif
#ifndef K1_WAVE_SLOTS
#define K1_WAVE_SLOTS 64           // query slots per warp (two home slots per lane)
#endif
#ifndef K1_WAVE_HOME
#define K1_WAVE_HOME 0             // 1: no compaction across lanes -- a lane serves only its two home slots (conflict-free shared memory)
#endif
#ifndef K1_WAVE_REFILL
#define K1_WAVE_REFILL 16          // refill from the next brick as soon as this many slots are free
#endif
#ifndef K1_WAVE_MIN_BLOCKS
#define K1_WAVE_MIN_BLOCKS 10      // blocks of K1_THREADS per SM the register allocation must allow (shared memory allows about as many)
#endif

// One of the four row-major 3-D node arrays of the grid (vertex nodes, x-/y-/z-edge nodes), restricted to the
// slow-planes [s0, s1) that a node range touches.  l = l_base + (s*Dm + m)*Df + f.
// K1_BRICK_AUTO 1: the warp's brick is chosen per node array at launch time -- the power-of-two shape (f x m x s, 32 nodes) whose PHYSICAL
// diagonal is shortest -- instead of the fixed K1_BRICK_F x K1_BRICK_M x K1_BRICK_S.  On cubic cells that is the default 4 x 4 x 2
// (x-edge arrays: their node spacing along f is half a cell); on strongly anisotropic cells (a flat bounding box sampled with a cubic
// resolution) a flatter brick keeps the 32 queries closer together (CPU schedule model: -6 % issued instructions on the bench torus).
// Checked against the oracle by the emulated test builds (on and off).
#ifndef K1_BRICK_AUTO
#define K1_BRICK_AUTO 1          // measured with K1_VOTE_REDUX (profiles/r2g_sweep.txt): bunny.obj 128^3 77.0 -> 75.0 ms, anisotropic torus grid 54.9 -> 51.9 ms
#endif

struct K1Segment {
    unsigned l_base, Ds, Dm, Df, s0, s1, tiles_f, tiles_m, tiles_s, block_begin;
#if K1_BRICK_AUTO
    unsigned lf, lm;               // log2 of the brick's extents along f and m (s extent = 32 >> (lf + lm))
#endif
    unsigned pl_stride;            // plane groups between consecutive brick layers (1 = contiguous slab, n_parts = interleaved deal)
    unsigned out_base;             // interleaved mode: element offset of this array inside the part's exchange slot
    int kind;                      // 0 vertex (s,m,f)=(k,j,i); 1 x-edge (k,j,2i+b); 2 y-edge (i,k,2j+b); 3 z-edge (j,i,2k+b)
};
struct K1Work {
    K1Segment seg[4];
    int nseg;
    unsigned l_begin, l_end;
    int compact;                   // 0: out[l - l_begin]; 1: interleaved exchange slot (see k1_launch_sample_interleaved)
};

// K1_VOTE_REDUX 1: the phase vote reads the three lane counts from ONE warp-wide integer sum (redux.sync) instead of two ballots +
// three popcounts (the loop head is ~19 % of the issued instructions).
#ifndef K1_VOTE_REDUX
#define K1_VOTE_REDUX 1          // measured (profiles/r2a_sweep.txt, r2g_sweep.txt): -1.5 %
#endif

struct DeviceBvh {                 // device mirrors of HostBvh, uploaded once by dg_mesh_create
    const SpherePair* spheres = nullptr;
    const LeafRecord* leaves = nullptr;
    const PseudoNormals* normals = nullptr;
    const LeafF* leaves_f = nullptr;            // fp32 triangle shadows (leaf filter)
    const float4* nodes_f = nullptr;          // fp32 record per internal node, K1_NODEF_STRIDE float4s: SpherePairF (2) + BoxPairF (3) [+ pad]
    double ctr[3] = {0, 0, 0};
    float half_extent = 0.f;
    int n_tri = 0;
    int stack_depth = 1;           // deferred-sibling slots per lane (= tree levels - 1, at least 1)
};

cudaError_t k1_configure(int stack_depth);
cudaError_t k1_launch_sample_nodes(const DeviceBvh& m, const GridDev& g, double sign, uint64_t l_begin, uint64_t count,
                                   double* d_out, cudaStream_t stream);
// slab form: planes [plane_begin[a], plane_end[a]) of node array a (a = 0..3), written at d_full[l]
cudaError_t k1_launch_sample_slab(const DeviceBvh& m, const GridDev& g, double sign, const unsigned plane_begin[4], const unsigned plane_end[4],
                                  double* d_full, cudaStream_t stream);
// interleaved deal of plane pairs: part `part` owns pairs part, part + n_parts, ... of every node array and writes
// them compactly into its exchange slot; k1_launch_unpack_interleaved scatters the gathered slots into node order.
struct InterleavedLayout {
    unsigned n_parts = 1;
    unsigned rot[4];               // plane group p of array a belongs to part (p + rot[a]) % n_parts: the deal is rotated per array so that the
                                   // parts that get one plane group more differ from array to array (129 planes = 65 pairs over 8 parts: 9:8 in
                                   // every array without the rotation, 33:32 over the four arrays with it)
    unsigned off[4][16];           // off[a][part]: element offset of array a in part's slot
    unsigned pairs[4];             // plane pairs of array a
    unsigned plane[4];             // elements per plane of array a
    uint64_t slot_elems = 0;       // elements per slot (max over parts)
#if K1_BRICK_AUTO
    unsigned lf[4], lm[4];         // brick shape of array a (its plane groups are 32 >> (lf + lm) planes thick)
#endif
};
bool k1_interleaved_layout(const GridDev& g, unsigned n_parts, InterleavedLayout& L);
cudaError_t k1_launch_sample_interleaved(const DeviceBvh& m, const GridDev& g, double sign, const InterleavedLayout& L, unsigned part,
                                         double* d_slot, cudaStream_t stream);
void k1_interleaved_node_slots(const GridDev& g, const InterleavedLayout& L, uint64_t l_begin, uint64_t count, uint32_t* part_out, uint64_t* pos_out);
struct K1Run { uint64_t slot_pos, node_begin, count; };      // slot[slot_pos .. +count) == nodes[node_begin .. +count)
void k1_interleaved_runs(const GridDev& g, const InterleavedLayout& L, unsigned part, std::vector<K1Run>& runs);
cudaError_t k1_launch_unpack_interleaved(const GridDev& g, const InterleavedLayout& L, const double* d_slots, double* d_nodes, cudaStream_t stream);
cudaError_t k1_launch_distance(const DeviceBvh& m, const double* d_pts, uint64_t count, int is_signed, double* d_dist,
                               double* d_near, int* d_ent, int* d_tri, cudaStream_t stream);
cudaError_t k1_launch_node_positions(const GridDev& g, uint64_t l_begin, uint64_t count, double* d_x, cudaStream_t stream);
cudaError_t k1_launch_build_cells(const GridDev& g, uint64_t c_begin, uint64_t count, unsigned* d_cells, cudaStream_t stream);
cudaError_t k1_launch_fp64_rate_probe(int blocks, int iters, double* d_out, cudaStream_t stream);
cudaError_t k1_launch_fma_probe(double a, double b, double c, double* d_out, cudaStream_t stream);

}  // namespace dgb
