// reduceField (cubic_lagrange_discrete_grid.cpp:1065-1174) as index arithmetic: SURVEY 8(f) row N3.
//
// The reference keeps, per node, a std::set of (cell, slot) back-references and moves nodes around one at a time; what it
// computes is a renumbering.  Here the same renumbering is produced from three flat passes:
//   1. cells: a cell survives if any of its 32 nodes is marked keep (:1081-1098); survivors keep their order; cell_map as :1076-1097;
//   2. nodes: those referenced by a surviving cell survive.  The reference compacts them by "swap the dead node with the last
//      live slot, walking from the back" (:1136-1156) -- replayed on a permutation array, O(n), no sets;
//   3. the survivors are ordered by the Morton key of their position (zValue, :583-601, with the LUT's 16-bit truncation,
//      z_sort_table.hpp:119-134).  std::sort is not stable, so when two survivors share a key -- the normal case on the anisotropic
//      cells of a bounding-box-fitted domain -- the reference's order is whatever libstdc++'s introsort leaves.  That sort is
//      replayed on all host threads (replay_std_sort below: same comparisons, same swaps, same result); should the process run on
//      a standard library that sorts differently, the plain std::sort call is made instead (any sort while keys are distinct).
// Host code: the reference's reduceField is host bookkeeping around the sampled fields; nothing here is a kernel's fallback.
#pragma once
#include <cstdint>
#include <vector>
#include "dg_device.cuh"

namespace dgb {

struct ReduceStats {
    uint64_t nodes_out = 0, cells_out = 0;
    int tie_path = 0;              // 1 = duplicate Morton keys among the survivors: their order is std::sort's, not the keys'
    double ms_cells = 0, ms_nodes = 0, ms_sort = 0, ms_write = 0;
};

// nodes[n_nodes] / cells[n_cells_in][32] are rewritten in place (first nodes_out / cells_out entries valid afterwards);
// cell_map[n_cells_grid] is written as the reference does (:1076-1097).  keep_node[l] != 0 <=> pred(x_l, c_l) && c_l != DBL_MAX.
// force_std_sort: always take the reference's own sort (tests).
bool reduce_field_host(const GridDev& g, double* nodes, uint64_t n_nodes, const uint8_t* keep_node, uint32_t* cells, uint64_t n_cells_in,
                       uint32_t* cell_map, uint64_t n_cells_grid, bool force_std_sort, ReduceStats& st, const char** err);

// The tie path: std::sort on (key, position) records ordered by key only, replayed on all threads (sort_replay.h).
struct KeyPos { uint64_t key; uint32_t pos; };
void replay_std_sort(KeyPos* a, uint64_t n, unsigned n_threads);
bool replay_std_sort_matches();
uint64_t replay_std_sort_heap_fallbacks();      // number of depth-exhausted ranges so far (tests)

// the two stages that stay on the host when the index passes run on the GPU (k4_reduce.cu / dg_reduce_field): the reference's
// back-to-front swap compaction replayed on a permutation, and the (tie-order-preserving) sort of the survivors
uint64_t reduce_swap_walk(const uint8_t* used, uint64_t n_nodes, std::vector<uint32_t>& perm);
void reduce_order_survivors(std::vector<KeyPos>& kp, bool force_std_sort, std::vector<uint32_t>& order, int& tie_path);

// zValue(indexToNodePosition(l), 4 * min(inv_cell_size)) for one node (exposed for tests)
uint64_t reduce_field_morton_key(const GridDev& g, uint32_t l);

}  // namespace dgb
