// Launch interface of the reduceField index kernels (k4_reduce.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "dg_device.cuh"

namespace dgb {

unsigned k4_tile();          // entries per tile of the counting / compaction passes
cudaError_t k4_launch_cell_keep(const unsigned* d_cells, uint64_t n_cells, const unsigned char* d_keep_node, unsigned n_nodes, unsigned char* d_cell_keep,
                                unsigned* d_bad, cudaStream_t stream);
cudaError_t k4_launch_tile_count(const unsigned char* d_flag, uint64_t n, unsigned* d_tile_count, cudaStream_t stream);
cudaError_t k4_launch_cell_compact(const unsigned* d_cells, uint64_t n_cells, const unsigned char* d_cell_keep, const unsigned* d_tile_offset,
                                   unsigned* d_cells_out, unsigned* d_cell_map, cudaStream_t stream);
cudaError_t k4_launch_mark_used(const unsigned* d_cells_out, uint64_t n_entries, unsigned char* d_used, cudaStream_t stream);
cudaError_t k4_launch_morton_keys(const GridDev& g, const unsigned* d_perm, uint64_t m, unsigned long long* d_key, cudaStream_t stream);
cudaError_t k4_launch_renumber(const unsigned* d_perm, const unsigned* d_order, uint64_t m, const double* d_nodes, unsigned* d_new_id, double* d_nodes_out,
                               unsigned* d_cells_out, uint64_t n_entries, cudaStream_t stream);

}  // namespace dgb
