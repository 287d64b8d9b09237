// reduceField (cubic_lagrange_discrete_grid.cpp:1065-1174) as index arithmetic: SURVEY 8(f) row N3.
//
// The reference keeps, per node, a std::set of (cell, slot) back-references and moves nodes around one at a time; what it
// computes is a renumbering.  Here the same renumbering is produced from three flat passes:
//   1. cells: a cell survives if any of its 32 nodes is marked keep (:1081-1098); survivors keep their order; cell_map as :1076-1097;
//   2. nodes: those referenced by a surviving cell survive.  The reference compacts them by "swap the dead node with the last
//      live slot, walking from the back" (:1136-1156) -- replayed on a permutation array, O(n), no sets;
//   3. the survivors are ordered by the Morton key of their position (zValue, :583-601, with the LUT's 16-bit truncation,
//      z_sort_table.hpp:119-134).  std::sort is not stable, so when two survivors share a key the reference's order is whatever
//      libstdc++'s introsort leaves: in that case the very same std::sort call is made (same input sequence, same comparisons =>
//      same result); when all keys are distinct -- the normal case -- the order is unique and a multithreaded sample sort is used.
// Host code: the reference's reduceField is host bookkeeping around the sampled fields; nothing here is a kernel's fallback.
#pragma once
#include <cstdint>
#include "dg_device.cuh"

namespace dgb {

struct ReduceStats {
    uint64_t nodes_out = 0, cells_out = 0;
    int tie_path = 0;              // 1 = duplicate Morton keys among the survivors: the reference's std::sort was replayed
    double ms_cells = 0, ms_nodes = 0, ms_sort = 0, ms_write = 0;
};

// nodes[n_nodes] / cells[n_cells_in][32] are rewritten in place (first nodes_out / cells_out entries valid afterwards);
// cell_map[n_cells_grid] is written as the reference does (:1076-1097).  keep_node[l] != 0 <=> pred(x_l, c_l) && c_l != DBL_MAX.
// force_std_sort: always take the reference's own sort (tests).
bool reduce_field_host(const GridDev& g, double* nodes, uint64_t n_nodes, const uint8_t* keep_node, uint32_t* cells, uint64_t n_cells_in,
                       uint32_t* cell_map, uint64_t n_cells_grid, bool force_std_sort, ReduceStats& st, const char** err);

// zValue(indexToNodePosition(l), 4 * min(inv_cell_size)) for one node (exposed for tests)
uint64_t reduce_field_morton_key(const GridDev& g, uint32_t l);

}  // namespace dgb
