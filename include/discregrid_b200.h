/* =====================================================================================
 * discregrid_b200.h -- C-ABI of the B200-native (sm_100a) Discregrid hot path.
 *
 * Plain C, plain pointers and sizes, no torch / Eigen / STL types.  This is the drop-in
 * boundary: the reference (InteractiveComputerGraphics/Discregrid @ ddf20dc) has no FFI of its
 * own -- its boundary is the C++ class API -- so each entry point names the reference
 * function(s) it replaces (paths relative to the reference root):
 *
 *   geometry/TriangleMeshDistance.h = discregrid/include/Discregrid/geometry/TriangleMeshDistance.h
 *   cubic_lagrange_discrete_grid.cpp = discregrid/src/cubic_lagrange_discrete_grid.cpp
 *   discrete_grid.{hpp,cpp}          = discregrid/include/Discregrid/discrete_grid.hpp, discregrid/src/discrete_grid.cpp
 *
 * Conventions
 *   - every function returns DG_OK (0) or a negative dg_status; dg_last_error() gives the message of
 *     the calling thread's last failure.  The library never calls exit() (the reference does:
 *     TriangleMeshDistance.h:318-321, 338-341).
 *   - "host" entry points take host pointers and include the H2D/D2H copies; "_device" entry points
 *     take device pointers of the CURRENT CUDA device plus a cudaStream_t (passed as void*, NULL =
 *     default stream), are asynchronous on that stream, and are what bench.py times with CUDA events.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     DG_ERR_NO_DEVICE.
 *   - arithmetic is IEEE fp64, round-to-nearest, no FMA contraction, in the reference's operation
 *     order, so results are bit-identical to the reference's x86-64 build.
 *   - the sentinel for "no value" is DBL_MAX, as in the reference
 *     (cubic_lagrange_discrete_grid.cpp:817, 981-982, 993-994, 1015-1018, 1050-1054).
 * ===================================================================================== */
#ifndef DISCREGRID_B200_H
#define DISCREGRID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_ABI_VERSION 1

#if defined(__GNUC__)
#define DG_API __attribute__((visibility("default")))
#else
#define DG_API
#endif

typedef enum dg_status {
    DG_OK = 0,
    DG_ERR_INVALID = -1,     /* bad argument (NULL, empty mesh, zero resolution, range out of bounds ...) */
    DG_ERR_NO_DEVICE = -2,   /* no usable CUDA device: there is no CPU fallback */
    DG_ERR_CUDA = -3,        /* a CUDA runtime call failed; dg_last_error() has the cudaError string */
    DG_ERR_NOMEM = -4,       /* host or device allocation failed */
    DG_ERR_SELFTEST = -5,    /* device self-test failed (e.g. library built with FMA contraction on) */
    DG_ERR_IO = -6           /* a file could not be opened or read */
} dg_status;

/* TriangleMeshDistance.h:75 -- same numbering */
typedef enum dg_nearest_entity { DG_V0 = 0, DG_V1, DG_V2, DG_E01, DG_E12, DG_E02, DG_F } dg_nearest_entity;

/* Grid geometry = the DiscreteGrid members (discrete_grid.hpp:93-98).  cell_size / inv_cell_size are passed
 * verbatim from the host object (never recomputed on the device); dg_grid_init fills them the way the
 * DiscreteGrid constructor does (discrete_grid.hpp:22-29). */
typedef struct dg_grid_desc {
    double   domain_min[3];
    double   domain_max[3];
    uint32_t resolution[3];
    uint32_t _pad;
    double   cell_size[3];
    double   inv_cell_size[3];
} dg_grid_desc;

typedef struct dg_mesh  dg_mesh;    /* device-resident TriangleMeshDistance (BVH + triangles + pseudonormals) */
typedef struct dg_field dg_field;   /* device-resident field of a CubicLagrangeDiscreteGrid (one m_nodes/m_cells/m_cell_map entry) */

/* ---- library ------------------------------------------------------------------------------- */
DG_API int         dg_abi_version(void);
DG_API const char* dg_last_error(void);
/* Number of CUDA devices visible (0 if none / no driver). */
DG_API int         dg_device_count(void);
/* cudaSetDevice for this library's CUDA runtime instance (call it with LOCAL_RANK in a one-process-per-GPU job). */
DG_API int         dg_set_device(int device);
/* Runs the device self-test (FMA-contraction probe + trivial kernel) on the current device. */
DG_API int         dg_selftest(void);
/* Measures the fp64 issue rate of the current device for the library's instruction mix (DMUL + DADD, no FMA contraction -- the
 * numerical contract forbids fused operations): Tflop/s.  bench.py uses it as the measured denominator of the fp64 rooflines. */
DG_API int         dg_fp64_rate_probe(double* tflops);
/* Kernels launched by this library in this process since load / last reset (all streams). */
DG_API uint64_t    dg_kernel_launch_count(void);
DG_API void        dg_kernel_launch_count_reset(void);

/* ---- grid helpers (host, no GPU) ----------------------------------------------------------------
 * replaces: DiscreteGrid ctor discrete_grid.hpp:22-29; node count cubic_lagrange_discrete_grid.cpp:790-796;
 * GenerateSDF's asymmetric domain padding cmd/generate_sdf/main.cpp:83-91. */
DG_API int dg_grid_init(const double domain_min[3], const double domain_max[3], const uint32_t resolution[3], dg_grid_desc* out);
DG_API int dg_grid_num_nodes(const uint32_t resolution[3], uint64_t* n_nodes);
DG_API int dg_generate_sdf_domain(const double* vertices, uint64_t n_vertices, double domain_min[3], double domain_max[3]);

/* ---- OBJ reader (host, no GPU) ---------------------------------------------------------
 * replaces the parser of Discregrid::TriangleMesh(path) (src/mesh/triangle_mesh.cpp:90-124): "v x y z" and "f a[/..] b[/..] c[/..]"
 * lines, everything else ignored; same doubles as the stream extraction (correctly rounded), indices made 0-based.  The arrays
 * are allocated by the library and released with dg_obj_free.  DG_ERR_IO when the file cannot be read, DG_ERR_INVALID for a face
 * index the reference's std::stoi would throw on. */
DG_API int dg_obj_read(const char* path, double** vertices, uint64_t* n_vertices, uint32_t** triangles, uint64_t* n_triangles);
DG_API void dg_obj_free(double* vertices, uint32_t* triangles);

/* ---- mesh / distance ------------------------------------------------------------------------------
 * dg_mesh_create replaces TriangleMeshDistance::construct / _construct / _build_tree
 * (TriangleMeshDistance.h:251-267, 336-441, 443-512): builds the reference's bounding-sphere tree (same
 * split, same std::sort order, same spheres) and the angle-weighted pseudonormals on the host, re-lays
 * them out for the GPU and uploads them ONCE.  vertices: n_vertices x 3 doubles (xyzxyz), triangles:
 * n_triangles x 3 uint32 (0-based).  Fails with DG_ERR_INVALID on an empty triangle list or an index
 * out of range (the reference exits / reads out of bounds). */
DG_API int dg_mesh_create(const double* vertices, uint64_t n_vertices, const uint32_t* triangles, uint64_t n_triangles, dg_mesh** out);
DG_API int dg_mesh_destroy(dg_mesh* mesh);
/* info[0]=n_vertices, [1]=n_triangles, [2]=max stack depth, [3]=watertight flags (bit0: edge with one
 * triangle, bit1: edge with >2 triangles -- the two warnings of TriangleMeshDistance.h:422-438),
 * [4]=device bytes, [5]=host build microseconds, [6]=upload microseconds */
DG_API int dg_mesh_info(const dg_mesh* mesh, uint64_t info[8]);
/* Test/diagnostic: copies the host-built tree in the reference's node numbering (pre-order; left == -1 marks
 * a leaf whose `right` is the triangle id).  spheres: n_nodes x 8 (cl.xyz, rl, cr.xyz, rr), kids: n_nodes x 2.
 * n_nodes = 2*n_triangles-1.  Any pointer may be NULL. */
DG_API int dg_mesh_tree(const dg_mesh* mesh, double* spheres, int32_t* kids);
/* Test/diagnostic: pseudonormals in the reference's arrays (TriangleMeshDistance.h:122-124). */
DG_API int dg_mesh_pseudonormals(const dg_mesh* mesh, double* tri /*nT x3*/, double* edge /*nT x3x3*/, double* vert /*nV x3*/);

/* Batched TriangleMeshDistance::signed_distance / unsigned_distance (TriangleMeshDistance.h:269-328) for n
 * points (n x 3 doubles).  Outputs (each nullable): distance[n], nearest_point[n x 3], nearest_entity[n]
 * (dg_nearest_entity), triangle_id[n] -- the fields of Discregrid::Result (TriangleMeshDistance.h:80-86). */
DG_API int dg_mesh_distance(const dg_mesh* mesh, const double* points, uint64_t n, int is_signed,
                     double* distance, double* nearest_point, int32_t* nearest_entity, int32_t* triangle_id);
DG_API int dg_mesh_distance_device(const dg_mesh* mesh, const double* d_points, uint64_t n, int is_signed,
                            double* d_distance, double* d_nearest_point, int32_t* d_nearest_entity,
                            int32_t* d_triangle_id, void* stream);

/* ---- K1: addFunction node loop with the GenerateSDF functor -------------------------------------------
 * replaces the hot loop cubic_lagrange_discrete_grid.cpp:806-817 with
 * func = sign * md.signed_distance(x).distance (cmd/generate_sdf/main.cpp:97,101; sign = -1 for --invert):
 *     out[l - l_begin] = sign * signed_distance(indexToNodePosition(l)).distance,  l in [l_begin, l_end)
 * indexToNodePosition = cubic_lagrange_discrete_grid.cpp:604-665.  Node ids are the reference's (uint32 in
 * the file format); 64-bit here only so ranges can be expressed without overflow.
 * A sub-range is what a rank computes when the grid is sharded across GPUs.
 * The host form overlaps kernel chunks, D2H DMA (pooled pinned double buffer) and the copy into `out_host`; for ranges of 32 MiB and more
 * (environment: DG_HOST_HELPERS_MIN_BYTES) a few worker threads (a quarter of the usable CPUs, DG_HOST_THREADS) pre-fault `out_host` and
 * share that copy, as in dg_add_function_sdf; calls are serialised per process.  The _device form is asynchronous on `stream` and may be
 * issued concurrently on several streams. */
DG_API int dg_sample_sdf(const dg_mesh* mesh, const dg_grid_desc* grid, double sign, uint64_t l_begin, uint64_t l_end, double* out_host);
DG_API int dg_sample_sdf_device(const dg_mesh* mesh, const dg_grid_desc* grid, double sign, uint64_t l_begin, uint64_t l_end,
                         double* d_out, void* stream);
/* The WHOLE of CubicLagrangeDiscreteGrid::addFunction with the GenerateSDF functor (cubic_lagrange_discrete_grid.cpp:780-899) into
 * the three arrays the reference appends to m_nodes / m_cells / m_cell_map: nodes_host[n_nodes] (the node loop :806-817, on the GPU),
 * cells_host[n_cells x 32] (connectivity :833-886) and cell_map_host[n_cells] (identity :888-891) -- the two index tables are written
 * by host threads straight into the caller's memory WHILE the GPU samples the nodes, so the call ends one D2H piece after the last
 * kernel chunk.  cells_host / cell_map_host may be NULL (skipped).  timings_ms: NULL or 6 doubles = total, node pipeline done,
 * index tables done, coefficient array pre-faulted (all ms since entry), worker threads used, 0.
 * This is what the C++ facade's addFunction(MeshSignedDistanceFunction) calls and what bench.py times as `e2e`. */
DG_API int dg_add_function_sdf(const dg_mesh* mesh, const dg_grid_desc* grid, double sign, double* nodes_host, uint32_t* cells_host,
                        uint32_t* cell_map_host, double* timings_ms);
/* ---- multi-GPU from ONE process (`dg_sample_sdf(..., n_gpus, ...)`; 8e) -------------------------------------------
 * A dg_mesh_group replicates the device records of `mesh` on n_gpus - 1 further GPUs (peer copies from the mesh's device; the host
 * build is not repeated).  devices: NULL = the mesh's device first, then the lowest other ids; else n_gpus ids with devices[0] == the
 * mesh's device.  The group borrows `mesh` (destroy the group first).  n_gpus = 1 is valid (same code path on one GPU).
 *   dg_add_function_sdf_multi : dg_add_function_sdf across the group -- plane pairs of the four node arrays dealt round-robin to
 *       2 x n_gpus parts, one host thread and two streams per GPU, every part DMA'd to its final place in nodes_host (directly if that
 *       memory is page-locked, else through pinned staging); index tables by host threads meanwhile.  No collective: the exchange
 *       target is host memory.  This is what `GenerateSDF --gpus N` calls.
 *   dg_sample_sdf_multi_device: the coefficient array device-resident on EVERY GPU of the group (d_out[i] = n_nodes doubles in
 *       device i's memory): one launch per GPU into its slot of an exchange buffer, ONE in-place ncclAllGather per device
 *       (ncclCommInitAll, grouped calls; libnccl.so.2 is loaded at first use), one unpack kernel into the reference's node order. */
typedef struct dg_mesh_group dg_mesh_group;
DG_API int dg_mesh_group_create(const dg_mesh* mesh, int n_gpus, const int* devices, dg_mesh_group** out);
DG_API int dg_mesh_group_destroy(dg_mesh_group* group);
DG_API int dg_mesh_group_size(const dg_mesh_group* group);
DG_API int dg_add_function_sdf_multi(dg_mesh_group* group, const dg_grid_desc* grid, double sign, double* nodes_host, uint32_t* cells_host,
                              uint32_t* cell_map_host, double* timings_ms);
DG_API int dg_sample_sdf_multi_device(dg_mesh_group* group, const dg_grid_desc* grid, double sign, double* const* d_out);
/* Slab sharding of the same loop for one-process-per-GPU jobs ("P1"): part `part` of `n_parts` samples, in ONE launch,
 * whole slow-plane PAIRS of each of the four row-major node arrays (vertex / x-edge nodes: z-slabs; y-edge nodes: x-slabs; z-edge
 * nodes: y-slabs), i.e. no partially filled bricks and a single launch tail per rank.  Results are written at their final positions
 * of the full coefficient array d_full (n_nodes doubles).  ranges[8] = the four [l_begin, l_end) node ranges of this part (some may
 * be empty); dg_slab_ranges computes them on the host (no GPU) for the exchange step. */
DG_API int dg_slab_ranges(const dg_grid_desc* grid, uint32_t part, uint32_t n_parts, uint64_t ranges[8]);
DG_API int dg_sample_sdf_slab_device(const dg_mesh* mesh, const dg_grid_desc* grid, double sign, uint32_t part, uint32_t n_parts,
                              double* d_full, void* stream);
/* Interleaved slab sharding: the slow-plane pairs of every node array are dealt round-robin -- pair p belongs to
 * part p % n_parts -- which balances the spatially varying cost almost perfectly and still needs ONE launch per rank and no masked
 * bricks.  A part writes its pairs compactly into its slot of an exchange buffer of n_parts x slot_elems doubles
 * (dg_interleaved_slot_elems); after ONE all-gather of the slots, dg_interleaved_unpack_device scatters them into the
 * reference's node order.  n_parts <= 16. */
DG_API int dg_interleaved_slot_elems(const dg_grid_desc* grid, uint32_t n_parts, uint64_t* slot_elems);
DG_API int dg_sample_sdf_interleaved_device(const dg_mesh* mesh, const dg_grid_desc* grid, double sign, uint32_t part, uint32_t n_parts,
                                     double* d_slot, void* stream);
DG_API int dg_interleaved_unpack_device(const dg_grid_desc* grid, uint32_t n_parts, const double* d_slots, double* d_nodes, void* stream);
/* the same layout on the host (no GPU): for node ids l in [l_begin, l_end), the part that samples node l and its position inside
 * that part's slot (either output may be NULL) -- for hosts that consume or exchange slots themselves */
DG_API int dg_interleaved_node_slots(const dg_grid_desc* grid, uint32_t n_parts, uint64_t l_begin, uint64_t l_end, uint32_t* part_out,
                              uint64_t* pos_out);
/* indexToNodePosition for l in [l_begin, l_end) -> x[(l-l_begin)*3 ..] (cubic_lagrange_discrete_grid.cpp:604-665) */
DG_API int dg_node_positions(const dg_grid_desc* grid, uint64_t l_begin, uint64_t l_end, double* x_host);
/* Cell connectivity table of addFunction (cubic_lagrange_discrete_grid.cpp:833-886) for cells
 * [c_begin, c_end): cells_host[(c-c_begin)*32 + j]. */
DG_API int dg_build_cells(const uint32_t resolution[3], uint64_t c_begin, uint64_t c_end, uint32_t* cells_host);

/* ---- fields / K2: interpolate ------------------------------------------------------------------------
 * dg_field_create uploads one field (m_nodes[f], m_cells[f], m_cell_map[f] of
 * cubic_lagrange_discrete_grid.hpp:69-71) and builds the device layout: one contiguous 256-byte block of the
 * 32 coefficients per kept cell.  cells == NULL means the closed-form table addFunction builds
 * (cubic_lagrange_discrete_grid.cpp:833-886, n_cells_kept must then equal nx*ny*nz); cell_map == NULL means identity
 * (:888-891).  After reduceField() pass the reduced arrays (cell_map entries UINT32_MAX = removed cell). */
DG_API int dg_field_create(const dg_grid_desc* grid, const double* nodes, uint64_t n_nodes, const uint32_t* cells,
                    uint64_t n_cells_kept, const uint32_t* cell_map, dg_field** out);
/* same, coefficient array already on the device (e.g. straight out of dg_sample_sdf_device); unreduced only */
DG_API int dg_field_create_device(const dg_grid_desc* grid, const double* d_nodes, uint64_t n_nodes, void* stream, dg_field** out);
DG_API int dg_field_destroy(dg_field* field);
/* info[0]=n_nodes, [1]=n_cells_kept, [2]=device bytes */
DG_API int dg_field_info(const dg_field* field, uint64_t info[4]);

/* Batched CubicLagrangeDiscreteGrid::interpolate(field_id, x, gradient*) (cubic_lagrange_discrete_grid.cpp:977-1063;
 * identical arithmetic to determineShapeFunctions + split interpolate, :901-975) for n points x[n x 3].
 * phi[n]; grad[n x 3] nullable (value-only path :1006-1023).  Out-of-domain / removed cell / missing coefficient
 * -> phi = DBL_MAX and grad = 0 exactly as the reference. */
DG_API int dg_interpolate_batch(const dg_field* field, const double* x, uint64_t n, double* phi, double* grad);
DG_API int dg_interpolate_batch_device(const dg_field* field, const double* d_x, uint64_t n, double* d_phi, double* d_grad, void* stream);
/* shape_function_ (cubic_lagrange_discrete_grid.cpp:339-580) for n reference-cell points xi[n x 3]:
 * N[n x 32], dN[n x 32 x 3] (nullable). */
DG_API int dg_shape_functions(const double* xi, uint64_t n, double* N, double* dN);

/* ---- K3: GenerateDensityMap node function ---------------------------------------------------------
 * replaces addFunction(density_func, verbose, pred) of cmd/generate_density_map/main.cpp:96-133 with
 * GaussQuadrature::integrate(p=30 -> 16^3 points, gauss_quadrature.cpp:5927-5960) and CubicKernel::W
 * (sph_kernel.hpp:22-42) over nodes [l_begin, l_end) of the grid of `sdf` (its field 0):
 * out = DBL_MAX where the predicate rejects (unless no_reduction), 0 where dist > 2h, else rho0 * integral. */
DG_API int dg_density_map(const dg_field* sdf, double h, double rho0, int no_reduction, uint64_t l_begin, uint64_t l_end, double* out_host);
DG_API int dg_density_map_device(const dg_field* sdf, double h, double rho0, int no_reduction, uint64_t l_begin, uint64_t l_end,
                          double* d_out, void* stream);

/* ---- reduceField -------------------------------------------------------------------
 * replaces the body of CubicLagrangeDiscreteGrid::reduceField (cubic_lagrange_discrete_grid.cpp:1065-1174) after the predicate has
 * been evaluated: keep_node[l] != 0 <=> pred(x_l, c_l) && c_l != DBL_MAX (:1069-1074).  Rewrites IN PLACE, exactly as the
 * reference leaves its members: cells (n_cells_in x 32; first *n_cells_out rows = surviving cells, renumbered), nodes (first
 * *n_nodes_out coefficients = surviving nodes in the reference's Z-curve order) and cell_map (resolution product entries; 0xffffffff
 * for removed cells).
 * flags: DG_REDUCE_REFERENCE_SORT makes the plain single-threaded std::sort call instead of its multithreaded replay (same result).
 * timings_ms: NULL or 5 doubles (ms spent on cells, node compaction, sort, write-back; 1.0 if surviving nodes shared Morton keys,
 * i.e. the node order depended on how std::sort leaves equal keys). */
#define DG_REDUCE_REFERENCE_SORT 1u
/* DG_REDUCE_HOST_PASSES (or the environment variable DG_REDUCE_FIELD_HOST=1): every pass on the host threads.  Default: the index passes run
 * on the GPU (k4_reduce.cu: cell flags, cell map + row compaction, node marks, Z-curve keys, renumbering and coefficient gather); the host
 * keeps the reference's serial swap-walk compaction and the sort whose order among equal keys is std::sort's.  Without a CUDA device the
 * default fails with DG_ERR_NO_DEVICE -- nothing switches silently. */
#define DG_REDUCE_HOST_PASSES 2u
DG_API int dg_reduce_field(const dg_grid_desc* grid, double* nodes, uint64_t n_nodes, const uint8_t* keep_node, uint32_t* cells,
                    uint64_t n_cells_in, uint32_t* cell_map, uint32_t flags, uint64_t* n_nodes_out, uint64_t* n_cells_out,
                    double* timings_ms);

#ifdef __cplusplus
}
#endif
#endif /* DISCREGRID_B200_H */
