// C-ABI of libdiscregrid_b200.so (include/discregrid_b200.h): handles, error reporting, host<->device plumbing.
// All numerics live in the kernels (k1_sdf.cu, k2_interp.cu, k3_density.cu) and in bvh_build.cpp.
#include "../../include/discregrid_b200.h"

#include <atomic>
#include <condition_variable>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#if defined(__linux__)
#include <sys/mman.h>
#endif

// NCCL is used through dlopen (multi-GPU section): only its types are needed at build time
#if __has_include(<nccl.h>) && !defined(DG_EMU)
#include <nccl.h>
#include <dlfcn.h>
#define DG_HAVE_NCCL 1
#else
#define DG_HAVE_NCCL 0
#endif

#include "bvh_build.h"
#include "dg_device.cuh"
#include "k1_sdf.h"
#include "k2_interp.h"
#include "k3_density.h"
#include "k4_reduce.h"
#include "reduce_field.h"
#include "obj_reader.h"

using namespace dgb;

// ------------------------------------------------------------------------------------------------ state
namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// C++ exceptions (allocation failures in the host-side containers and threads) must not cross the C boundary
template <class F>
int guarded(const char* who, F&& body)
{
    try { return body(); }
    catch (const std::bad_alloc&) { return fail(DG_ERR_NOMEM, "%s: out of host memory", who); }
    catch (const std::exception& ex) { return fail(DG_ERR_INVALID, "%s: %s", who, ex.what()); }
}

#define DG_CUDA(call)                                                                                    \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            const int code_ = (e_ == cudaErrorMemoryAllocation) ? DG_ERR_NOMEM                           \
                              : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? DG_ERR_NO_DEVICE \
                                                                                                : DG_ERR_CUDA; \
            return fail(code_, "%s failed: %s", #call, cudaGetErrorString(e_));                         \
        }                                                                                                \
    } while (0)

#define DG_LAUNCH(call)          \
    do {                         \
        DG_CUDA(call);           \
        g_launches.fetch_add(1); \
    } while (0)

int require_device()
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(DG_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return DG_OK;
}

// a handle's device arrays (and K1's dynamic shared-memory attribute) belong to the device it was created on
int check_handle_device(int handle_device, const char* who)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(DG_ERR_NO_DEVICE, "%s: no current CUDA device", who); }
    if (dev != handle_device) return fail(DG_ERR_INVALID, "%s: the handle lives on device %d but the current device is %d (dg_set_device)", who, handle_device, dev);
    return DG_OK;
}

// page-locked host memory (cudaMallocHost / cudaHostRegister, by any CUDA runtime instance of the process) can be DMA'd directly
bool is_pinned_host(const void* p)
{
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// simple owning device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        return cudaMalloc(reinterpret_cast<void**>(&p), (count ? count : 1) * sizeof(T));
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

bool grid_to_dev(const dg_grid_desc* d, GridDev& g, const char** why)
{
    if (!d) { *why = "grid descriptor is NULL"; return false; }
    for (int k = 0; k < 3; k++) {
        if (d->resolution[k] == 0) { *why = "resolution must be >= 1 in every dimension"; return false; }
        g.mn[k] = d->domain_min[k]; g.mx[k] = d->domain_max[k];
        g.cell[k] = d->cell_size[k]; g.inv[k] = d->inv_cell_size[k];
        g.n[k] = d->resolution[k];
    }
    uint64_t nn = 0;
    if (dg_grid_num_nodes(d->resolution, &nn) != DG_OK) { *why = "grid too large: node ids must fit uint32 (reference file format)"; return false; }
    const unsigned nx = g.n[0], ny = g.n[1], nz = g.n[2];
    g.nv = (nx + 1) * (ny + 1) * (nz + 1);
    g.ne_x = nx * (ny + 1) * (nz + 1);
    g.ne_y = (nx + 1) * ny * (nz + 1);
    g.ne_z = (nx + 1) * (ny + 1) * nz;
    return true;
}

std::atomic<int> g_selftest_state{0};   // 0 = not run, 1 = ok, -1 = failed

}  // namespace

// ------------------------------------------------------------------------------------------------ handles
struct dg_mesh {
    HostBvh host;
    DevBuf<SpherePair> d_spheres;
    DevBuf<LeafRecord> d_leaves;
    DevBuf<PseudoNormals> d_normals;
    DevBuf<float4> d_nodes_f;
    DevBuf<LeafF> d_leaves_f;
    DeviceBvh dev;
    int device = 0;
    uint64_t build_us = 0, upload_us = 0;
};

struct dg_field {
    dg_grid_desc desc;
    FieldDev dev;
    DevBuf<double> d_packed;
    DevBuf<unsigned> d_cell_map;
    DevBuf<double2> d_tab;
    uint64_t n_nodes = 0, n_cells_kept = 0;
    int device = 0;
};

// ------------------------------------------------------------------------------------------------ library
extern "C" {

int dg_abi_version(void) { return DG_ABI_VERSION; }
const char* dg_last_error(void) { return g_err.c_str(); }
uint64_t dg_kernel_launch_count(void) { return g_launches.load(); }
void dg_kernel_launch_count_reset(void) { g_launches.store(0); }

int dg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int dg_set_device(int device)
{
    if (int rc = require_device()) return rc;
    DG_CUDA(cudaSetDevice(device));
    return DG_OK;
}

int dg_selftest(void)
{
    if (int rc = require_device()) return rc;
    // a*b + c where the unfused result (two roundings) differs from the fused one:
    // a = b = 1 + 2^-27  ->  a*b = 1 + 2^-26 + 2^-54 (rounds to 1 + 2^-26); c = -(1 + 2^-26).
    const double a = 1.0 + std::ldexp(1.0, -27), c = -(1.0 + std::ldexp(1.0, -26));
    DevBuf<double> out;
    DG_CUDA(out.alloc(1));
    DG_LAUNCH(k1_launch_fma_probe(a, a, c, out.p, nullptr));
    double r = -1.0;
    DG_CUDA(cudaMemcpy(&r, out.p, sizeof r, cudaMemcpyDeviceToHost));
    if (r != 0.0) {
        g_selftest_state = -1;
        return fail(DG_ERR_SELFTEST, "device code was compiled with FMA contraction (a*b+c probe returned %.17g, expected 0): "
                    "build with -fmad=false", r);
    }
    g_selftest_state = 1;
    return DG_OK;
}

// Measured fp64 issue rate of the current device for the library's own instruction mix (DMUL + DADD, no FMA): Tflop/s.
int dg_fp64_rate_probe(double* tflops)
{
    if (int rc = require_device()) return rc;
    if (!tflops) return fail(DG_ERR_INVALID, "dg_fp64_rate_probe: NULL argument");
    int dev = 0, n_sm = 0;
    DG_CUDA(cudaGetDevice(&dev));
    DG_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = n_sm * 8, iters = 1 << 14;
    DevBuf<double> out;
    DG_CUDA(out.alloc(1));
    cudaEvent_t e0, e1;
    DG_CUDA(cudaEventCreate(&e0)); DG_CUDA(cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 4; rep++) {                       // first repetition warms up
        DG_CUDA(cudaEventRecord(e0, nullptr));
        DG_LAUNCH(k1_launch_fp64_rate_probe(blocks, iters, out.p, nullptr));
        DG_CUDA(cudaEventRecord(e1, nullptr));
        DG_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        DG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double flops = (double)blocks * 256.0 * 8.0 * 2.0 * (double)iters;
    *tflops = best > 0.f ? flops / (best * 1e-3) / 1e12 : 0.0;
    return DG_OK;
}

static int ensure_selftest()
{
    const int s = g_selftest_state.load();
    if (s == 1) return DG_OK;
    if (s == -1) return fail(DG_ERR_SELFTEST, "device self-test failed earlier (FMA contraction on)");
    return dg_selftest();
}

// ------------------------------------------------------------------------------------------------ grid helpers
int dg_grid_init(const double mn[3], const double mx[3], const uint32_t res[3], dg_grid_desc* out)
{
    if (!mn || !mx || !res || !out) return fail(DG_ERR_INVALID, "dg_grid_init: NULL argument");
    uint64_t nn;
    if (int rc = dg_grid_num_nodes(res, &nn)) return rc;
    std::memset(out, 0, sizeof *out);
    for (int d = 0; d < 3; d++) {
        out->domain_min[d] = mn[d]; out->domain_max[d] = mx[d]; out->resolution[d] = res[d];
        out->cell_size[d] = (mx[d] - mn[d]) / (double)res[d];      // diagonal().cwiseQuotient(n.cast<double>())
        out->inv_cell_size[d] = 1.0 / out->cell_size[d];           // cwiseInverse()
    }
    return DG_OK;
}

int dg_grid_num_nodes(const uint32_t res[3], uint64_t* n_nodes)
{
    if (!res || !n_nodes) return fail(DG_ERR_INVALID, "dg_grid_num_nodes: NULL argument");
    if (res[0] == 0 || res[1] == 0 || res[2] == 0) return fail(DG_ERR_INVALID, "resolution must be >= 1 in every dimension");
    const uint64_t nx = res[0], ny = res[1], nz = res[2];
    const uint64_t nv = (nx + 1) * (ny + 1) * (nz + 1);
    const uint64_t ne = nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz;
    const uint64_t total = nv + 2 * ne;
    // the reference indexes nodes with `int l` (cubic_lagrange_discrete_grid.cpp:809) and stores uint32 ids
    if (nx > 65535 || ny > 65535 || nz > 65535 || total > 0x7fffffffull || nx * ny * nz > 0x7fffffffull)
        return fail(DG_ERR_INVALID, "grid too large: %llu nodes exceed the reference's int/uint32 node ids", (unsigned long long)total);
    *n_nodes = total;
    return DG_OK;
}

int dg_generate_sdf_domain(const double* V, uint64_t nV, double mn[3], double mx[3])
{
    if (!V || !mn || !mx || nV == 0) return fail(DG_ERR_INVALID, "dg_generate_sdf_domain: NULL / empty argument");
    for (int d = 0; d < 3; d++) { mn[d] = DBL_MAX; mx[d] = -DBL_MAX; }     // AlignedBox::setEmpty + extend
    for (uint64_t i = 0; i < nV; i++)
        for (int d = 0; d < 3; d++) { mn[d] = std::fmin(mn[d], V[3 * i + d]); mx[d] = std::fmax(mx[d], V[3 * i + d]); }
    // cmd/generate_sdf/main.cpp:89-90: max is padded first, then min with the grown diagonal.  Eigen >= 3.3 reduces a
    // Vector3d as (a0 + a1) + a2.
    for (int pass = 0; pass < 2; pass++) {
        const double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        const double nrm = std::sqrt((dx * dx + dy * dy) + dz * dz);
        const double pad = 1.0e-3 * nrm * 1.0;
        for (int d = 0; d < 3; d++) { if (pass == 0) mx[d] += pad; else mn[d] -= pad; }
    }
    return DG_OK;
}

// ------------------------------------------------------------------------------------------------ mesh
int dg_mesh_create(const double* V, uint64_t nV, const uint32_t* F, uint64_t nT, dg_mesh** out)
{
    if (!out) return fail(DG_ERR_INVALID, "dg_mesh_create: out is NULL");
    *out = nullptr;
    if (!V || !F || nT == 0 || nV == 0) return fail(DG_ERR_INVALID, "dg_mesh_create: empty triangle list");   // ref: exit(-1)
    if (int rc = require_device()) return rc;
    if (int rc = ensure_selftest()) return rc;
    dg_mesh* m = new (std::nothrow) dg_mesh();
    if (!m) return fail(DG_ERR_NOMEM, "dg_mesh_create: out of host memory");
    const auto t0 = std::chrono::steady_clock::now();
    const char* why = "";
    try {
        if (!build_host_bvh(V, nV, F, nT, m->host, &why, K1_NEEDS_LEAF_SHADOW != 0)) { delete m; return fail(DG_ERR_INVALID, "dg_mesh_create: %s", why); }
    } catch (const std::bad_alloc&) { delete m; return fail(DG_ERR_NOMEM, "dg_mesh_create: out of host memory"); }
    const auto t1 = std::chrono::steady_clock::now();
    m->build_us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();

    auto cleanup = [&](int rc) { delete m; return rc; };
#define DG_CUDA_M(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cleanup(fail(e_ == cudaErrorMemoryAllocation ? DG_ERR_NOMEM : DG_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
    DG_CUDA_M(cudaGetDevice(&m->device));
    DG_CUDA_M(m->d_spheres.alloc(nT));
    DG_CUDA_M(m->d_leaves.alloc(nT));
    DG_CUDA_M(m->d_normals.alloc(nT));
    DG_CUDA_M(m->d_nodes_f.alloc(nT * K1_NODEF_STRIDE));
#if K1_NEEDS_LEAF_SHADOW
    DG_CUDA_M(m->d_leaves_f.alloc(nT));
    DG_CUDA_M(cudaMemcpy(m->d_leaves_f.p, m->host.leaves_f.data(), nT * sizeof(LeafF), cudaMemcpyHostToDevice));
    m->dev.leaves_f = m->d_leaves_f.p;
#endif
    DG_CUDA_M(cudaMemcpy(m->d_spheres.p, m->host.spheres.data(), nT * sizeof(SpherePair), cudaMemcpyHostToDevice));
    DG_CUDA_M(cudaMemcpy(m->d_leaves.p, m->host.leaves.data(), nT * sizeof(LeafRecord), cudaMemcpyHostToDevice));
    DG_CUDA_M(cudaMemcpy(m->d_normals.p, m->host.normals.data(), nT * sizeof(PseudoNormals), cudaMemcpyHostToDevice));
    {   // the fp32 record of every internal node: sphere pair + child boxes, bvh_build.h
        static_assert(K1_NODEF_STRIDE * sizeof(float4) >= sizeof(SpherePairF) + sizeof(BoxPairF), "node record too small");
        RawVec<float4> rec(nT * K1_NODEF_STRIDE);
        pack_node_records(m->host, K1_NODEF_STRIDE, reinterpret_cast<float*>(rec.data()));
        DG_CUDA_M(cudaMemcpy(m->d_nodes_f.p, rec.data(), rec.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    m->dev.nodes_f = m->d_nodes_f.p;
    for (int d = 0; d < 3; d++) m->dev.ctr[d] = m->host.center[d];
    m->dev.half_extent = (float)m->host.half_extent * 1.0000002f;   // rounded up
    m->dev.spheres = m->d_spheres.p; m->dev.leaves = m->d_leaves.p; m->dev.normals = m->d_normals.p;
    m->dev.n_tri = (int)nT;
    m->dev.stack_depth = m->host.max_depth > 1 ? m->host.max_depth - 1 : 1;
    DG_CUDA_M(k1_configure(m->dev.stack_depth));
#undef DG_CUDA_M
    const auto t2 = std::chrono::steady_clock::now();
    m->upload_us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
    // the device records are the product; the host copies of the big arrays are no longer needed
    RawVec<SpherePair>().swap(m->host.spheres);
    RawVec<SpherePairF>().swap(m->host.spheres_f);
    RawVec<BoxPairF>().swap(m->host.boxes_f);
    RawVec<LeafF>().swap(m->host.leaves_f);        // keep `order`, V, F, pseudonormals for diagnostics
    RawVec<LeafRecord>().swap(m->host.leaves);
    RawVec<PseudoNormals>().swap(m->host.normals);
    *out = m;
    return DG_OK;
}

int dg_mesh_destroy(dg_mesh* m)
{
    if (!m) return DG_OK;
    delete m;
    return DG_OK;
}

int dg_mesh_info(const dg_mesh* m, uint64_t info[8])
{
    if (!m || !info) return fail(DG_ERR_INVALID, "dg_mesh_info: NULL argument");
    info[0] = m->host.n_vertices; info[1] = m->host.n_triangles; info[2] = (uint64_t)m->dev.stack_depth;
    info[3] = (uint64_t)m->host.flags;
    info[4] = m->d_spheres.bytes() + m->d_leaves.bytes() + m->d_normals.bytes() + m->d_nodes_f.bytes() + m->d_leaves_f.bytes();
    info[5] = m->build_us; info[6] = m->upload_us; info[7] = 0;
    return DG_OK;
}

int dg_mesh_tree(const dg_mesh* m, double* spheres, int32_t* kids)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_mesh_tree: mesh is NULL");
    // the sphere pairs were released after upload: read them back from the device
    HostBvh tmp;
    tmp.n_triangles = m->host.n_triangles;
    tmp.order = m->host.order;
    tmp.spheres.resize(m->host.n_triangles);
    DG_CUDA(cudaMemcpy(tmp.spheres.data(), m->d_spheres.p, m->host.n_triangles * sizeof(SpherePair), cudaMemcpyDeviceToHost));
    export_reference_tree(tmp, spheres, kids);
    return DG_OK;
}

int dg_mesh_pseudonormals(const dg_mesh* m, double* tri, double* edge, double* vert)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_mesh_pseudonormals: mesh is NULL");
    if (tri) std::memcpy(tri, m->host.pn_tri.data(), m->host.pn_tri.size() * sizeof(double));
    if (edge) std::memcpy(edge, m->host.pn_edge.data(), m->host.pn_edge.size() * sizeof(double));
    if (vert) std::memcpy(vert, m->host.pn_vert.data(), m->host.pn_vert.size() * sizeof(double));
    return DG_OK;
}

int dg_mesh_distance_device(const dg_mesh* m, const double* d_pts, uint64_t n, int is_signed, double* d_dist, double* d_near,
                            int32_t* d_ent, int32_t* d_tri, void* stream)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_mesh_distance: mesh is NULL (not constructed)");
    if (n && !d_pts) return fail(DG_ERR_INVALID, "dg_mesh_distance: points is NULL");
    if (int rc = check_handle_device(m->device, "dg_mesh_distance")) return rc;
    DG_LAUNCH(k1_launch_distance(m->dev, d_pts, n, is_signed, d_dist, d_near, d_ent, d_tri, (cudaStream_t)stream));
    return DG_OK;
}

int dg_mesh_distance(const dg_mesh* m, const double* pts, uint64_t n, int is_signed, double* dist, double* nearp, int32_t* ent,
                     int32_t* tri)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_mesh_distance: mesh is NULL (not constructed)");
    if (n == 0) return DG_OK;
    if (!pts) return fail(DG_ERR_INVALID, "dg_mesh_distance: points is NULL");
    DevBuf<double> d_pts, d_dist, d_near;
    DevBuf<int> d_ent, d_tri;
    DG_CUDA(d_pts.alloc(3 * n));
    DG_CUDA(cudaMemcpy(d_pts.p, pts, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    if (dist) DG_CUDA(d_dist.alloc(n));
    if (nearp) DG_CUDA(d_near.alloc(3 * n));
    if (ent) DG_CUDA(d_ent.alloc(n));
    if (tri) DG_CUDA(d_tri.alloc(n));
    if (int rc = dg_mesh_distance_device(m, d_pts.p, n, is_signed, d_dist.p, d_near.p, d_ent.p, d_tri.p, nullptr)) return rc;
    if (dist) DG_CUDA(cudaMemcpy(dist, d_dist.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (nearp) DG_CUDA(cudaMemcpy(nearp, d_near.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (ent) DG_CUDA(cudaMemcpy(ent, d_ent.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (tri) DG_CUDA(cudaMemcpy(tri, d_tri.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    DG_CUDA(cudaDeviceSynchronize());
    return DG_OK;
}

// ------------------------------------------------------------------------------------------------ K1
static int check_range(const dg_grid_desc* grid, uint64_t b, uint64_t e, GridDev& g, const char* who)
{
    const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "%s: %s", who, why);
    uint64_t nn = 0;
    if (int rc = dg_grid_num_nodes(grid->resolution, &nn)) return rc;
    if (b > e || e > nn) return fail(DG_ERR_INVALID, "%s: node range [%llu, %llu) outside [0, %llu]", who, (unsigned long long)b,
                                      (unsigned long long)e, (unsigned long long)nn);
    return DG_OK;
}

int dg_sample_sdf_device(const dg_mesh* m, const dg_grid_desc* grid, double sign, uint64_t l_begin, uint64_t l_end, double* d_out,
                         void* stream)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_sample_sdf: mesh is NULL (not constructed)");
    GridDev g;
    if (int rc = check_range(grid, l_begin, l_end, g, "dg_sample_sdf")) return rc;
    if (l_end > l_begin && !d_out) return fail(DG_ERR_INVALID, "dg_sample_sdf: output is NULL");
    if (int rc = check_handle_device(m->device, "dg_sample_sdf")) return rc;
    DG_LAUNCH(k1_launch_sample_nodes(m->dev, g, sign, l_begin, l_end - l_begin, d_out, (cudaStream_t)stream));
    return DG_OK;
}

// Host-buffer path of K1.  Resources that are expensive to create per call (device scratch, two pinned staging buffers,
// streams, events) live in a per-process pool.  The range is cut into a few kernel launches dealt to two streams (a K1
// launch ends with a ~2 ms tail of long-running warps; the next launch fills the machine meanwhile); a third stream DMAs
// finished chunks into the pinned double buffer while the CPU copies the previous piece into the caller's (pageable) memory.
namespace {
struct HostPathPool {
    std::mutex mu;
    int device = -1;
    double* d_out = nullptr; size_t d_cap = 0;
    double* stage[2] = {nullptr, nullptr};
    cudaStream_t s_k[2] = {nullptr, nullptr}, s_c = nullptr;
    cudaEvent_t dma_ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> chunk_ev;
    static constexpr size_t kPiece = 4u << 20;          // doubles per staging buffer (32 MiB)
    cudaError_t prepare(size_t n, size_t n_chunks)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev != device) { release(); device = dev; }
        if (!s_c) {
            for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&s_k[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_c, cudaStreamNonBlocking);
            for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&dma_ev[i], cudaEventDisableTiming);
            for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaMallocHost(reinterpret_cast<void**>(&stage[i]), kPiece * sizeof(double));
            if (e != cudaSuccess) { release(); return e; }          // never leave a half-built pool behind
        }
        if (n > d_cap) {
            if (d_out) cudaFree(d_out);
            d_out = nullptr; d_cap = 0;
            e = cudaMalloc(reinterpret_cast<void**>(&d_out), n * sizeof(double));
            if (e != cudaSuccess) return e;
            d_cap = n;
        }
        while (chunk_ev.size() < n_chunks) {
            cudaEvent_t ev;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            chunk_ev.push_back(ev);
        }
        return cudaSuccess;
    }
    void release()
    {
        if (d_out) cudaFree(d_out);
        d_out = nullptr; d_cap = 0;
        for (int i = 0; i < 2; i++) { if (stage[i]) cudaFreeHost(stage[i]); stage[i] = nullptr; if (s_k[i]) cudaStreamDestroy(s_k[i]); s_k[i] = nullptr;
                                      if (dma_ev[i]) cudaEventDestroy(dma_ev[i]); dma_ev[i] = nullptr; }
        if (s_c) cudaStreamDestroy(s_c);
        s_c = nullptr;
        for (auto ev : chunk_ev) cudaEventDestroy(ev);
        chunk_ev.clear();
    }
};
HostPathPool g_pool;
}  // namespace

// the pipeline of the host-buffer path; the caller holds g_pool.mu.  copier: the table workers of dg_add_function_sdf, which share the copies
// from the pinned piece into the caller's memory with this thread (fresh pageable memory: page faults and remote NUMA nodes make a single
// thread's copy the bottleneck on a loaded host); NULL: this thread copies alone
namespace { struct HostTablesJob; }
static void tables_job_copy(HostTablesJob* job, void* dst, const void* src, size_t bytes);
static int sample_sdf_host_locked(const dg_mesh* m, const GridDev& g, double sign, uint64_t l_begin, uint64_t n, double* out_host, HostTablesJob* copier = nullptr)
{
    // kernel chunks: a handful, each a multiple of the staging piece so that pieces never straddle chunks
    const uint64_t piece = HostPathPool::kPiece;
    const uint64_t n_pieces = (n + piece - 1) / piece;
    const uint64_t pieces_per_chunk = n_pieces <= 4 ? 1 : (n_pieces + 7) / 8;
    const uint64_t chunk = pieces_per_chunk * piece;
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    DG_CUDA(g_pool.prepare(n, n_chunks));
    for (uint64_t c = 0; c < n_chunks; c++) {
        const uint64_t b = c * chunk, cnt = (b + chunk <= n) ? chunk : n - b;
        DG_LAUNCH(k1_launch_sample_nodes(m->dev, g, sign, l_begin + b, cnt, g_pool.d_out + b, g_pool.s_k[c & 1]));
        DG_CUDA(cudaEventRecord(g_pool.chunk_ev[c], g_pool.s_k[c & 1]));
    }
    for (uint64_t i = 0; i <= n_pieces; i++) {
        if (i < n_pieces) {
            const uint64_t off = i * piece, cnt = (off + piece <= n) ? piece : n - off;
            DG_CUDA(cudaStreamWaitEvent(g_pool.s_c, g_pool.chunk_ev[off / chunk], 0));
            DG_CUDA(cudaMemcpyAsync(g_pool.stage[i & 1], g_pool.d_out + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, g_pool.s_c));
            DG_CUDA(cudaEventRecord(g_pool.dma_ev[i & 1], g_pool.s_c));
        }
        if (i >= 1) {
            const uint64_t off = (i - 1) * piece, cnt = (off + piece <= n) ? piece : n - off;
            DG_CUDA(cudaEventSynchronize(g_pool.dma_ev[(i - 1) & 1]));
            const double* src = g_pool.stage[(i - 1) & 1];
            if (copier) tables_job_copy(copier, out_host + off, src, cnt * sizeof(double));
            else std::memcpy(out_host + off, src, cnt * sizeof(double));
        }
    }
    return DG_OK;
}

// Makes the pages of [p, p + bytes) present and writable WITHOUT changing their content (safe against a concurrent writer).
static void prefault_preserving(char* p, uint64_t bytes)
{
    if (bytes == 0) return;
#if defined(__linux__) && !defined(DG_EMU)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(4095), z = (reinterpret_cast<uintptr_t>(p) + bytes + 4095) & ~uintptr_t(4095);
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
    if (madvise(reinterpret_cast<void*>(a), z - a, MADV_POPULATE_WRITE) == 0) return;
#endif
    for (uint64_t o = 0; o < bytes; o += 4096) __atomic_fetch_or(reinterpret_cast<unsigned char*>(p + o), 0, __ATOMIC_RELAXED);
    __atomic_fetch_or(reinterpret_cast<unsigned char*>(p + bytes - 1), 0, __ATOMIC_RELAXED);
}

// closed-form connectivity of cells [c0, c1) (cubic_lagrange_discrete_grid.cpp:833-886), host side: plain index algebra
static void fill_cells_host(const GridDev& g, uint64_t c0, uint64_t c1, uint32_t* out /* (c1 - c0) x 32 */)
{
    const unsigned nx = g.n[0], ny = g.n[1], nz = g.n[2];
    const unsigned nv = g.nv, off_x = nv, off_y = nv + 2u * g.ne_x, off_z = off_y + 2u * g.ne_y;
    unsigned k = (unsigned)(c0 / ((uint64_t)nx * ny)), rem = (unsigned)(c0 % ((uint64_t)nx * ny)), j = rem / nx, i = rem % nx;
    for (uint64_t c = c0; c < c1; c++) {
        uint32_t* cell = out + 32 * (c - c0);
        const unsigned v00 = (nx + 1) * (ny + 1) * k + (nx + 1) * j + i, v10 = v00 + (nx + 1), v01 = v00 + (nx + 1) * (ny + 1), v11 = v01 + (nx + 1);
        cell[0] = v00; cell[1] = v00 + 1; cell[2] = v10; cell[3] = v10 + 1; cell[4] = v01; cell[5] = v01 + 1; cell[6] = v11; cell[7] = v11 + 1;
        const unsigned x0 = off_x + 2 * (nx * (ny + 1) * k + nx * j + i), x1 = off_x + 2 * (nx * (ny + 1) * (k + 1) + nx * j + i);
        const unsigned x2 = off_x + 2 * (nx * (ny + 1) * k + nx * (j + 1) + i), x3 = off_x + 2 * (nx * (ny + 1) * (k + 1) + nx * (j + 1) + i);
        cell[8] = x0; cell[9] = x0 + 1; cell[10] = x1; cell[11] = x1 + 1; cell[12] = x2; cell[13] = x2 + 1; cell[14] = x3; cell[15] = x3 + 1;
        const unsigned y0 = off_y + 2 * (ny * (nz + 1) * i + ny * k + j), y1 = off_y + 2 * (ny * (nz + 1) * (i + 1) + ny * k + j);
        const unsigned y2 = off_y + 2 * (ny * (nz + 1) * i + ny * (k + 1) + j), y3 = off_y + 2 * (ny * (nz + 1) * (i + 1) + ny * (k + 1) + j);
        cell[16] = y0; cell[17] = y0 + 1; cell[18] = y1; cell[19] = y1 + 1; cell[20] = y2; cell[21] = y2 + 1; cell[22] = y3; cell[23] = y3 + 1;
        const unsigned z0 = off_z + 2 * (nz * (nx + 1) * j + nz * i + k), z1 = off_z + 2 * (nz * (nx + 1) * (j + 1) + nz * i + k);
        const unsigned z2 = off_z + 2 * (nz * (nx + 1) * j + nz * (i + 1) + k), z3 = off_z + 2 * (nz * (nx + 1) * (j + 1) + nz * (i + 1) + k);
        cell[24] = z0; cell[25] = z0 + 1; cell[26] = z1; cell[27] = z1 + 1; cell[28] = z2; cell[29] = z2 + 1; cell[30] = z3; cell[31] = z3 + 1;
        if (++i == nx) { i = 0; if (++j == ny) { j = 0; ++k; } }
    }
}

// Host side of addFunction while the GPU samples the nodes: worker threads pre-fault the coefficient array (content-preserving: the
// D2H pipeline may already be writing there), then write the connectivity table (:833-886) and the identity cell map (:888-891)
// straight into the caller's memory -- 4 MiB blocks dealt dynamically.
namespace {
struct HostTablesJob {
    GridDev g; uint64_t n_nodes = 0, n_cells = 0;
    double* nodes = nullptr; uint32_t* cells = nullptr; uint32_t* cell_map = nullptr;
    std::atomic<uint64_t> next_task{0};
    std::atomic<int> prefault_left{0};
    uint64_t nb_nodes = 0, nb_cells = 0, nb_map = 0, n_tasks = 0;
    static constexpr uint64_t blk = 4u << 20;
    std::chrono::steady_clock::time_point t0;
    double ms_prefault = 0.0;
    unsigned n_workers = 0;
    std::vector<std::thread> th;
    // copy service: once the tables are written the workers wait here and share the copies staging buffer -> caller memory with the
    // thread that drives the pipeline (one thread copies ~2-4 GB/s into pages another thread faulted in; the pipeline needs a 32 MiB
    // piece every ~14 ms, which a loaded host does not sustain single-threaded: profiles/r2l_multi_probe.txt)
    struct CopyJob { char* dst; const char* src; size_t bytes; std::atomic<size_t> next{0}, done{0}; };
    std::mutex c_mu; std::condition_variable c_cv;
    std::shared_ptr<CopyJob> c_job;                              // guarded by c_mu; a late worker only ever touches the job it picked up
    uint64_t c_gen = 0; bool c_quit = false;
    static constexpr size_t c_blk = 1u << 20;
    static void copy_blocks(CopyJob& j)
    {
        for (;;) {
            const size_t b = j.next.fetch_add(c_blk);
            if (b >= j.bytes) break;
            const size_t n = std::min(c_blk, j.bytes - b);
            std::memcpy(j.dst + b, j.src + b, n);
            j.done.fetch_add(n);
        }
    }
    void copy(void* dst, const void* src, size_t bytes)          // called by the pipeline's thread
    {
        if (bytes < 4 * c_blk || th.empty()) { std::memcpy(dst, src, bytes); return; }
        auto job = std::make_shared<CopyJob>();
        job->dst = static_cast<char*>(dst); job->src = static_cast<const char*>(src); job->bytes = bytes;
        { std::lock_guard<std::mutex> lk(c_mu); c_job = job; c_gen++; }
        c_cv.notify_all();
        copy_blocks(*job);
        while (job->done.load() < bytes) std::this_thread::yield();
    }
    void serve_copies()
    {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<CopyJob> job;
            {
                std::unique_lock<std::mutex> lk(c_mu);
                c_cv.wait(lk, [&] { return c_quit || c_gen != seen; });
                if (c_quit) return;
                seen = c_gen; job = c_job;
            }
            if (job) copy_blocks(*job);
        }
    }
    void work()
    {
        for (;;) {
            const uint64_t t = next_task.fetch_add(1);
            if (t >= n_tasks) break;
            if (t < nb_nodes) {
                char* p = reinterpret_cast<char*>(nodes);
                const uint64_t b = t * blk, e = std::min<uint64_t>(n_nodes * 8, b + blk);
                prefault_preserving(p + b, e - b);
                if (prefault_left.fetch_sub(1) == 1) ms_prefault = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            } else if (t < nb_nodes + nb_cells) {
                const uint64_t per = blk / 128, c0 = (t - nb_nodes) * per, c1 = std::min<uint64_t>(n_cells, c0 + per);
                fill_cells_host(g, c0, c1, cells + 32 * c0);
            } else {
                const uint64_t per = blk / 4, c0 = (t - nb_nodes - nb_cells) * per, c1 = std::min<uint64_t>(n_cells, c0 + per);
                for (uint64_t c = c0; c < c1; c++) cell_map[c] = (uint32_t)c;
            }
        }
    }
    void start(const GridDev& g_, uint64_t n_nodes_, double* nodes_, uint32_t* cells_, uint32_t* cell_map_)
    {
        g = g_; n_nodes = n_nodes_; n_cells = (uint64_t)g.n[0] * g.n[1] * g.n[2];
        nodes = nodes_; cells = cells_; cell_map = cell_map_;
        nb_nodes = (n_nodes * 8 + blk - 1) / blk;
        nb_cells = cells ? (n_cells * 128 + blk - 1) / blk : 0;
        nb_map = cell_map ? (n_cells * 4 + blk - 1) / blk : 0;
        n_tasks = nb_nodes + nb_cells + nb_map;
        prefault_left = (int)std::min<uint64_t>(nb_nodes, 0x7fffffff);
        t0 = std::chrono::steady_clock::now();
        // a quarter of the usable CPUs, at most 24.  Measured (profiles/r2d_e2e_probe.txt, 1-GPU lease with a 16-CPU cgroup quota, 256^3):
        // 2 workers -> the call ends 8 ms after the last kernel; 11 workers + 4 copy threads (= every CPU of the quota busy) -> 95 ms
        // later, because the bandwidth controller then throttles the thread that drives the GPU pipeline along with the rest.
        const unsigned hw = host_threads();
        n_workers = std::max(1u, std::min(hw / 4u, 24u));
        try { for (unsigned k = 0; k < n_workers; k++) th.emplace_back([this]() { work(); serve_copies(); }); } catch (...) { /* fewer workers: finish() does the rest */ }
    }
    void finish()
    {
        work();                                         // whatever is left
        { std::lock_guard<std::mutex> lk(c_mu); c_quit = true; }
        c_cv.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
    ~HostTablesJob() { if (!th.empty()) finish(); }      // an early return of the caller must not leave joinable threads behind
};
}  // namespace

static void tables_job_copy(HostTablesJob* job, void* dst, const void* src, size_t bytes) { job->copy(dst, src, bytes); }

int dg_sample_sdf(const dg_mesh* m, const dg_grid_desc* grid, double sign, uint64_t l_begin, uint64_t l_end, double* out_host)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_sample_sdf: mesh is NULL (not constructed)");
    GridDev g;
    if (int rc = check_range(grid, l_begin, l_end, g, "dg_sample_sdf")) return rc;
    const uint64_t n = l_end - l_begin;
    if (n == 0) return DG_OK;
    if (!out_host) return fail(DG_ERR_INVALID, "dg_sample_sdf: output is NULL");
    if (int rc = check_handle_device(m->device, "dg_sample_sdf")) return rc;
    // A large range into pageable memory: one thread copying staging buffer -> caller memory (2 - 4 GB/s into pages it has to fault in itself)
    // is slower than the packet walk produces coefficients (5.5 GB/s at 256^3), so a few workers pre-fault the range and share the copies,
    // exactly as in dg_add_function_sdf (profiles/r2x: per-rank e2e 26.1 ms for 17.8 ms of kernel before this).
    static const uint64_t helpers_min_bytes = []() -> uint64_t {
        const char* e = std::getenv("DG_HOST_HELPERS_MIN_BYTES");
        return (e && *e) ? (uint64_t)std::strtoull(e, nullptr, 10) : (uint64_t)(32u << 20);
    }();
    if (n * sizeof(double) >= helpers_min_bytes) {
        HostTablesJob job;
        job.start(g, n, out_host, nullptr, nullptr);
        int rc;
        {
            std::lock_guard<std::mutex> lock(g_pool.mu);
            rc = sample_sdf_host_locked(m, g, sign, l_begin, n, out_host, &job);
        }
        job.finish();
        return rc;
    }
    std::lock_guard<std::mutex> lock(g_pool.mu);
    return sample_sdf_host_locked(m, g, sign, l_begin, n, out_host);
}


// The whole of CubicLagrangeDiscreteGrid::addFunction(GenerateSDF functor) into the caller's three arrays
// (cubic_lagrange_discrete_grid.cpp:780-899): node loop on the GPU (K1 chunks on two streams, D2H through the pinned double
// buffer), and -- while the GPU works -- the host threads write the 32-index connectivity table (:833-886) and the identity cell
// map (:888-891) straight into the caller's memory and pre-fault the coefficient array, so that the call ends one D2H piece after
// the last kernel chunk.
static int dg_add_function_sdf_impl(const dg_mesh* m, const dg_grid_desc* grid, double sign, double* nodes_host, uint32_t* cells_host, uint32_t* cell_map_host,
                        double* timings_ms)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_add_function_sdf: mesh is NULL (not constructed)");
    GridDev g;
    uint64_t n_nodes = 0;
    if (grid && dg_grid_num_nodes(grid->resolution, &n_nodes)) return DG_ERR_INVALID;
    if (int rc = check_range(grid, 0, n_nodes, g, "dg_add_function_sdf")) return rc;
    if (!nodes_host) return fail(DG_ERR_INVALID, "dg_add_function_sdf: nodes output is NULL");
    if (int rc = check_handle_device(m->device, "dg_add_function_sdf")) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    HostTablesJob job;
    job.start(g, n_nodes, nodes_host, cells_host, cell_map_host);
    int rc;
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        rc = sample_sdf_host_locked(m, g, sign, 0, n_nodes, nodes_host, &job);
    }
    const double ms_nodes = ms_since(t0);
    job.finish();
    if (timings_ms) { timings_ms[0] = ms_since(t0); timings_ms[1] = ms_nodes; timings_ms[2] = timings_ms[0]; timings_ms[3] = job.ms_prefault; timings_ms[4] = (double)job.n_workers; timings_ms[5] = 1.0; }
    return rc;
}

// ------------------------------------------------------------------------------------------------ multi-GPU from ONE process
// `dg_sample_sdf(..., n_gpus, ...)`.  A dg_mesh_group replicates a mesh's device records on further GPUs (peer
// copies over NVLink from the first device, no second host build) and keeps per-device streams, buffers and -- for the device-resident
// form -- one NCCL communicator per device (ncclCommInitAll; libnccl is loaded at first use, the library does not link it).
// Work split: plane pairs of the four node arrays dealt round-robin (the interleaved layout of k1_sdf.cu), every part one launch.

struct dg_mesh_group {
    struct PerDev {
        cudaStream_t st[2] = {nullptr, nullptr};
        cudaEvent_t ev[16] = {};                            // one per part of this device (at most 16 parts in all)
        double* d_buf = nullptr; size_t d_cap = 0;          // host form: 2 slots; device form: n exchange slots
        double* h_stage = nullptr; size_t h_cap = 0;        // pinned staging (pageable destinations only)
        void* comm = nullptr;                               // ncclComm_t
    };
    int n = 0;
    std::vector<int> devices;
    std::vector<dg_mesh*> parts;                            // parts[0] is the caller's mesh (not owned); the others are replicas
    std::vector<PerDev> dev;
    bool comms_ready = false;
    std::mutex mu;
};

namespace {
#if DG_HAVE_NCCL
struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (so) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { so = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (so) break; }
        if (!so) return false;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(so, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(so, "ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(so, "ncclAllGather"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(so, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(so, "ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(so, "ncclGetErrorString"));
        return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString;
    }
};
NcclApi g_nccl;
#endif

struct DeviceGuard {                                  // restores the caller's current device
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int group_grow(dg_mesh_group::PerDev& d, size_t dev_elems, size_t host_elems)
{
    if (dev_elems > d.d_cap) {
        if (d.d_buf) cudaFree(d.d_buf);
        d.d_buf = nullptr; d.d_cap = 0;
        DG_CUDA(cudaMalloc(reinterpret_cast<void**>(&d.d_buf), dev_elems * sizeof(double)));
        d.d_cap = dev_elems;
    }
    if (host_elems > d.h_cap) {
        if (d.h_stage) cudaFreeHost(d.h_stage);
        d.h_stage = nullptr; d.h_cap = 0;
        DG_CUDA(cudaMallocHost(reinterpret_cast<void**>(&d.h_stage), host_elems * sizeof(double)));
        d.h_cap = host_elems;
    }
    return DG_OK;
}
}  // namespace

int dg_mesh_group_destroy(dg_mesh_group* grp)
{
    if (!grp) return DG_OK;
    DeviceGuard guard;
    for (int i = 0; i < grp->n; i++) {
        if (cudaSetDevice(grp->devices[i]) != cudaSuccess) { cudaGetLastError(); continue; }
        auto& d = grp->dev[i];
#if DG_HAVE_NCCL
        if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(d.comm));
#endif
        for (int k = 0; k < 2; k++) if (d.st[k]) cudaStreamDestroy(d.st[k]);
        for (int k = 0; k < 16; k++) if (d.ev[k]) cudaEventDestroy(d.ev[k]);
        if (d.d_buf) cudaFree(d.d_buf);
        if (d.h_stage) cudaFreeHost(d.h_stage);
        if (i > 0 && (size_t)i < grp->parts.size()) delete grp->parts[i];
    }
    delete grp;
    return DG_OK;
}

int dg_mesh_group_create(const dg_mesh* mesh, int n_gpus, const int* devices, dg_mesh_group** out)
{
    if (!out) return fail(DG_ERR_INVALID, "dg_mesh_group_create: out is NULL");
    *out = nullptr;
    if (!mesh) return fail(DG_ERR_INVALID, "dg_mesh_group_create: mesh is NULL (not constructed)");
    if (int rc = require_device()) return rc;
    const int have = dg_device_count();
    if (n_gpus < 1 || n_gpus > 16 || n_gpus > have) return fail(DG_ERR_INVALID, "dg_mesh_group_create: n_gpus = %d, but 1..min(16, %d visible devices) are possible", n_gpus, have);
    std::unique_ptr<dg_mesh_group, int (*)(dg_mesh_group*)> grp(new (std::nothrow) dg_mesh_group(), dg_mesh_group_destroy);
    if (!grp) return fail(DG_ERR_NOMEM, "dg_mesh_group_create: out of host memory");
    grp->devices.push_back(mesh->device);
    for (int k = 0, next = 0; k < n_gpus - 1; k++) {
        int dv;
        if (devices) { dv = devices[k + 1]; if (devices[0] != mesh->device) return fail(DG_ERR_INVALID, "dg_mesh_group_create: devices[0] must be the mesh's device (%d)", mesh->device); }
        else { while (next == mesh->device) next++; dv = next++; }
        if (dv < 0 || dv >= have || std::find(grp->devices.begin(), grp->devices.end(), dv) != grp->devices.end())
            return fail(DG_ERR_INVALID, "dg_mesh_group_create: bad or repeated device id %d", dv);
        grp->devices.push_back(dv);
    }
    grp->dev.resize(n_gpus);
    grp->parts.push_back(const_cast<dg_mesh*>(mesh));
    DeviceGuard guard;
    const size_t nT = mesh->host.n_triangles;
    for (int i = 0; i < n_gpus; i++) {
        DG_CUDA(cudaSetDevice(grp->devices[i]));
        grp->n = i + 1;                                             // destroy() only visits initialised entries
        auto& d = grp->dev[i];
        for (int k = 0; k < 2; k++) DG_CUDA(cudaStreamCreateWithFlags(&d.st[k], cudaStreamNonBlocking));
        for (int k = 0; k < 16; k++) DG_CUDA(cudaEventCreateWithFlags(&d.ev[k], cudaEventDisableTiming));
        if (i == 0) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, grp->devices[i], mesh->device) == cudaSuccess && can) { if (cudaDeviceEnablePeerAccess(mesh->device, 0) != cudaSuccess) cudaGetLastError(); }
        dg_mesh* r = new (std::nothrow) dg_mesh();
        if (!r) return fail(DG_ERR_NOMEM, "dg_mesh_group_create: out of host memory");
        grp->parts.push_back(r);
        r->device = grp->devices[i];
        r->host.n_triangles = mesh->host.n_triangles; r->host.n_vertices = mesh->host.n_vertices; r->host.flags = mesh->host.flags;
        DG_CUDA(r->d_spheres.alloc(nT)); DG_CUDA(r->d_leaves.alloc(nT)); DG_CUDA(r->d_normals.alloc(nT)); DG_CUDA(r->d_nodes_f.alloc(nT * K1_NODEF_STRIDE));
        DG_CUDA(cudaMemcpyPeer(r->d_spheres.p, r->device, mesh->d_spheres.p, mesh->device, nT * sizeof(SpherePair)));
        DG_CUDA(cudaMemcpyPeer(r->d_leaves.p, r->device, mesh->d_leaves.p, mesh->device, nT * sizeof(LeafRecord)));
        DG_CUDA(cudaMemcpyPeer(r->d_normals.p, r->device, mesh->d_normals.p, mesh->device, nT * sizeof(PseudoNormals)));
        DG_CUDA(cudaMemcpyPeer(r->d_nodes_f.p, r->device, mesh->d_nodes_f.p, mesh->device, nT * K1_NODEF_STRIDE * sizeof(float4)));
        r->dev = mesh->dev;
        r->dev.spheres = r->d_spheres.p; r->dev.leaves = r->d_leaves.p; r->dev.normals = r->d_normals.p; r->dev.nodes_f = r->d_nodes_f.p;
#if K1_NEEDS_LEAF_SHADOW
        DG_CUDA(r->d_leaves_f.alloc(nT));
        DG_CUDA(cudaMemcpyPeer(r->d_leaves_f.p, r->device, mesh->d_leaves_f.p, mesh->device, nT * sizeof(LeafF)));
        r->dev.leaves_f = r->d_leaves_f.p;
#endif
        DG_CUDA(k1_configure(r->dev.stack_depth));
        DG_CUDA(cudaDeviceSynchronize());
    }
    *out = grp.release();
    return DG_OK;
}

int dg_mesh_group_size(const dg_mesh_group* grp) { return grp ? grp->n : 0; }

// addFunction across the group's GPUs, host arrays out.  n x 2 parts (two launches per GPU on two streams, so that the D2H of a
// GPU's first part runs under its second launch); each GPU is driven by its own host thread and DMAs the contiguous plane-pair runs
// of its parts to their final positions (directly when nodes_host is page-locked, else through pinned staging).  No collective is
// needed: the exchange target is host memory.  cells_host / cell_map_host as in dg_add_function_sdf (nullable).
static int dg_add_function_sdf_multi_impl(dg_mesh_group* grp, const dg_grid_desc* grid, double sign, double* nodes_host, uint32_t* cells_host,
                              uint32_t* cell_map_host, double* timings_ms)
{
    if (!grp) return fail(DG_ERR_INVALID, "dg_add_function_sdf_multi: group is NULL");
    GridDev g;
    uint64_t n_nodes = 0;
    if (grid && dg_grid_num_nodes(grid->resolution, &n_nodes)) return DG_ERR_INVALID;
    if (int rc = check_range(grid, 0, n_nodes, g, "dg_add_function_sdf_multi")) return rc;
    if (!nodes_host) return fail(DG_ERR_INVALID, "dg_add_function_sdf_multi: nodes output is NULL");
    std::lock_guard<std::mutex> lock(grp->mu);
    const int n = grp->n;
    // parts per GPU: at least two (the D2H and the host copy of one part run under the next launch), more when the grid is large -- what
    // stays exposed at the end is the copy of ONE part into the caller's memory (profiles/r2l_multi_probe.txt) -- but never below ~4 M
    // nodes per launch (a launch ends with a ~2 ms tail) and never more than 16 parts in all
    unsigned splits = (2 * n <= 16) ? 2u : 1u;
    {
        const uint64_t by_size = n_nodes / (4000000ull * (uint64_t)n);
        const unsigned cap = 16u / (unsigned)n;
        if (by_size > splits) splits = (unsigned)std::min<uint64_t>(by_size, cap);
    }
    InterleavedLayout L;
    if (!k1_interleaved_layout(g, (unsigned)n * splits, L)) return fail(DG_ERR_INVALID, "dg_add_function_sdf_multi: grid too large for the exchange layout");
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    const bool direct = is_pinned_host(nodes_host);
    HostTablesJob job;
    job.start(g, n_nodes, nodes_host, cells_host, cell_map_host);
    std::vector<int> rcs(n, DG_OK);
    std::vector<std::string> errs(n);
    std::vector<uint64_t> launches(n, 0);
    auto drive = [&](int i) -> int {
        DG_CUDA(cudaSetDevice(grp->devices[i]));
        auto& d = grp->dev[i];
        if (int rc = group_grow(d, (size_t)splits * L.slot_elems, direct ? 0 : (size_t)splits * L.slot_elems)) return rc;
        std::vector<K1Run> runs[16];
        for (unsigned sp = 0; sp < splits; sp++) {
            const unsigned part = (unsigned)i + sp * (unsigned)n;
            double* d_slot = d.d_buf + (size_t)sp * L.slot_elems;
            cudaStream_t stream = d.st[sp & 1u];
            DG_CUDA(k1_launch_sample_interleaved(grp->parts[i]->dev, g, sign, L, part, d_slot, stream));
            launches[i]++;
            k1_interleaved_runs(g, L, part, runs[sp]);
            uint64_t used = 0;
            for (const auto& r : runs[sp]) used = std::max<uint64_t>(used, r.slot_pos + r.count);
            if (direct) { for (const auto& r : runs[sp]) DG_CUDA(cudaMemcpyAsync(nodes_host + r.node_begin, d_slot + r.slot_pos, r.count * sizeof(double), cudaMemcpyDeviceToHost, stream)); }
            else if (used) DG_CUDA(cudaMemcpyAsync(d.h_stage + (size_t)sp * L.slot_elems, d_slot, used * sizeof(double), cudaMemcpyDeviceToHost, stream));
            DG_CUDA(cudaEventRecord(d.ev[sp], stream));
        }
        for (unsigned sp = 0; sp < splits; sp++) {
            DG_CUDA(cudaEventSynchronize(d.ev[sp]));
            if (!direct) for (const auto& r : runs[sp]) std::memcpy(nodes_host + r.node_begin, d.h_stage + (size_t)sp * L.slot_elems + r.slot_pos, r.count * sizeof(double));
        }
        return DG_OK;
    };
    {
        DeviceGuard guard;
        std::vector<std::thread> th;
        auto guarded = [&](int i) {
            try { rcs[i] = drive(i); if (rcs[i] != DG_OK) errs[i] = g_err; }
            catch (const std::exception& ex) { rcs[i] = DG_ERR_NOMEM; errs[i] = ex.what(); }
        };
        for (int i = 1; i < n; i++) th.emplace_back([&, i]() { DeviceGuard tg; guarded(i); });
        guarded(0);
        for (auto& t : th) t.join();
    }
    const double ms_nodes = ms_since(t0);
    job.finish();
    for (int i = 0; i < n; i++) g_launches.fetch_add(launches[i]);
    if (timings_ms) { timings_ms[0] = ms_since(t0); timings_ms[1] = ms_nodes; timings_ms[2] = timings_ms[0]; timings_ms[3] = job.ms_prefault; timings_ms[4] = (double)job.n_workers; timings_ms[5] = (double)n; }
    for (int i = 0; i < n; i++) if (rcs[i] != DG_OK) return fail(rcs[i], "dg_add_function_sdf_multi (device %d): %s", grp->devices[i], errs[i].c_str());
    return DG_OK;
}

// Device-resident form (the coefficient array is needed on every GPU, e.g. for K2/K3 replicas): GPU i samples part i of n into its
// slot of an n x slot_elems exchange buffer, ONE in-place ncclAllGather per device (grouped) moves the slots over NVLink, and an
// unpack kernel scatters them into the reference's node order in d_out[i] (device i's memory, n_nodes doubles each).
int dg_sample_sdf_multi_device(dg_mesh_group* grp, const dg_grid_desc* grid, double sign, double* const* d_out)
{
    if (!grp || !d_out) return fail(DG_ERR_INVALID, "dg_sample_sdf_multi_device: NULL argument");
    GridDev g;
    uint64_t n_nodes = 0;
    if (grid && dg_grid_num_nodes(grid->resolution, &n_nodes)) return DG_ERR_INVALID;
    if (int rc = check_range(grid, 0, n_nodes, g, "dg_sample_sdf_multi_device")) return rc;
    std::lock_guard<std::mutex> lock(grp->mu);
    const int n = grp->n;
    for (int i = 0; i < n; i++) if (!d_out[i]) return fail(DG_ERR_INVALID, "dg_sample_sdf_multi_device: d_out[%d] is NULL", i);
    InterleavedLayout L;
    if (!k1_interleaved_layout(g, (unsigned)n, L)) return fail(DG_ERR_INVALID, "dg_sample_sdf_multi_device: grid too large for the exchange layout");
    DeviceGuard guard;
#if DG_HAVE_NCCL
    if (n > 1 && !grp->comms_ready) {
        if (!g_nccl.load()) return fail(DG_ERR_CUDA, "dg_sample_sdf_multi_device: libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbol");
        std::vector<ncclComm_t> comms(n);
        const ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, grp->devices.data());
        if (r != ncclSuccess) return fail(DG_ERR_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
        for (int i = 0; i < n; i++) grp->dev[i].comm = comms[i];
        grp->comms_ready = true;
    }
#else
    if (n > 1) return fail(DG_ERR_CUDA, "dg_sample_sdf_multi_device: built without NCCL");
#endif
    for (int i = 0; i < n; i++) {
        DG_CUDA(cudaSetDevice(grp->devices[i]));
        if (int rc = group_grow(grp->dev[i], (size_t)n * L.slot_elems, 0)) return rc;
        DG_LAUNCH(k1_launch_sample_interleaved(grp->parts[i]->dev, g, sign, L, (unsigned)i, grp->dev[i].d_buf + (size_t)i * L.slot_elems, grp->dev[i].st[0]));
    }
#if DG_HAVE_NCCL
    if (n > 1) {
        ncclResult_t r = g_nccl.GroupStart();
        for (int i = 0; i < n && r == ncclSuccess; i++)
            r = g_nccl.AllGather(grp->dev[i].d_buf + (size_t)i * L.slot_elems, grp->dev[i].d_buf, L.slot_elems, ncclDouble, static_cast<ncclComm_t>(grp->dev[i].comm), grp->dev[i].st[0]);
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) return fail(DG_ERR_CUDA, "ncclAllGather failed: %s", g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
    }
#endif
    for (int i = 0; i < n; i++) {
        DG_CUDA(cudaSetDevice(grp->devices[i]));
        DG_LAUNCH(k1_launch_unpack_interleaved(g, L, grp->dev[i].d_buf, d_out[i], grp->dev[i].st[0]));
    }
    for (int i = 0; i < n; i++) {
        DG_CUDA(cudaSetDevice(grp->devices[i]));
        DG_CUDA(cudaStreamSynchronize(grp->dev[i].st[0]));
    }
    return DG_OK;
}

// plane ranges of part `part` of `n_parts`: boundaries fall on even plane indices (a brick spans K1_BRICK_S = 2 planes)
static void slab_planes(const GridDev& g, uint32_t part, uint32_t n_parts, unsigned pb[4], unsigned pe[4])
{
    const unsigned Ds[4] = {g.n[2] + 1, g.n[2] + 1, g.n[0] + 1, g.n[1] + 1};
    for (int a = 0; a < 4; a++) {
        const uint64_t groups = (Ds[a] + K1_BRICK_S - 1) / K1_BRICK_S;
        const uint64_t q0 = groups * part / n_parts, q1 = groups * (part + 1) / n_parts;
        pb[a] = (unsigned)std::min<uint64_t>(q0 * K1_BRICK_S, Ds[a]);
        pe[a] = (unsigned)std::min<uint64_t>(q1 * K1_BRICK_S, Ds[a]);
    }
}

int dg_slab_ranges(const dg_grid_desc* grid, uint32_t part, uint32_t n_parts, uint64_t ranges[8])
{
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_slab_ranges: %s", why);
    if (!ranges || n_parts == 0 || part >= n_parts) return fail(DG_ERR_INVALID, "dg_slab_ranges: bad part / n_parts");
    unsigned pb[4], pe[4];
    slab_planes(g, part, n_parts, pb, pe);
    const uint64_t base[4] = {0, g.nv, (uint64_t)g.nv + 2ull * g.ne_x, (uint64_t)g.nv + 2ull * (g.ne_x + (uint64_t)g.ne_y)};
    const uint64_t plane[4] = {(uint64_t)(g.n[1] + 1) * (g.n[0] + 1), (uint64_t)(g.n[1] + 1) * 2 * g.n[0], (uint64_t)(g.n[2] + 1) * 2 * g.n[1],
                               (uint64_t)(g.n[0] + 1) * 2 * g.n[2]};
    for (int a = 0; a < 4; a++) { ranges[2 * a] = base[a] + plane[a] * pb[a]; ranges[2 * a + 1] = base[a] + plane[a] * pe[a]; }
    return DG_OK;
}

int dg_sample_sdf_slab_device(const dg_mesh* m, const dg_grid_desc* grid, double sign, uint32_t part, uint32_t n_parts, double* d_full, void* stream)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_sample_sdf_slab: mesh is NULL (not constructed)");
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_sample_sdf_slab: %s", why);
    if (n_parts == 0 || part >= n_parts || !d_full) return fail(DG_ERR_INVALID, "dg_sample_sdf_slab: bad part / n_parts / output");
    if (int rc = check_handle_device(m->device, "dg_sample_sdf_slab")) return rc;
    unsigned pb[4], pe[4];
    slab_planes(g, part, n_parts, pb, pe);
    DG_LAUNCH(k1_launch_sample_slab(m->dev, g, sign, pb, pe, d_full, (cudaStream_t)stream));
    return DG_OK;
}

int dg_interleaved_slot_elems(const dg_grid_desc* grid, uint32_t n_parts, uint64_t* slot_elems)
{
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_interleaved_slot_elems: %s", why);
    InterleavedLayout L;
    if (!slot_elems || !k1_interleaved_layout(g, n_parts, L)) return fail(DG_ERR_INVALID, "dg_interleaved_slot_elems: n_parts must be 1..16");
    *slot_elems = L.slot_elems;
    return DG_OK;
}

int dg_interleaved_node_slots(const dg_grid_desc* grid, uint32_t n_parts, uint64_t l_begin, uint64_t l_end, uint32_t* part_out, uint64_t* pos_out)
{
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_interleaved_node_slots: %s", why);
    InterleavedLayout L;
    if (!k1_interleaved_layout(g, n_parts, L)) return fail(DG_ERR_INVALID, "dg_interleaved_node_slots: n_parts must be 1..16");
    const uint64_t n = (uint64_t)g.nv + 2ull * ((uint64_t)g.ne_x + g.ne_y + g.ne_z);
    if (l_begin > l_end || l_end > n) return fail(DG_ERR_INVALID, "dg_interleaved_node_slots: node range outside [0, %llu]", (unsigned long long)n);
    k1_interleaved_node_slots(g, L, l_begin, l_end - l_begin, part_out, pos_out);
    return DG_OK;
}

int dg_sample_sdf_interleaved_device(const dg_mesh* m, const dg_grid_desc* grid, double sign, uint32_t part, uint32_t n_parts, double* d_slot, void* stream)
{
    if (!m) return fail(DG_ERR_INVALID, "dg_sample_sdf_interleaved: mesh is NULL (not constructed)");
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_sample_sdf_interleaved: %s", why);
    InterleavedLayout L;
    if (!k1_interleaved_layout(g, n_parts, L) || part >= n_parts || !d_slot) return fail(DG_ERR_INVALID, "dg_sample_sdf_interleaved: bad part / n_parts (1..16) / output");
    if (int rc = check_handle_device(m->device, "dg_sample_sdf_interleaved")) return rc;
    DG_LAUNCH(k1_launch_sample_interleaved(m->dev, g, sign, L, part, d_slot, (cudaStream_t)stream));
    return DG_OK;
}

int dg_interleaved_unpack_device(const dg_grid_desc* grid, uint32_t n_parts, const double* d_slots, double* d_nodes, void* stream)
{
    if (int rc = require_device()) return rc;
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_interleaved_unpack: %s", why);
    InterleavedLayout L;
    if (!k1_interleaved_layout(g, n_parts, L) || !d_slots || !d_nodes) return fail(DG_ERR_INVALID, "dg_interleaved_unpack: bad n_parts (1..16) / NULL buffer");
    DG_LAUNCH(k1_launch_unpack_interleaved(g, L, d_slots, d_nodes, (cudaStream_t)stream));
    return DG_OK;
}

int dg_node_positions(const dg_grid_desc* grid, uint64_t l_begin, uint64_t l_end, double* x_host)
{
    if (int rc = require_device()) return rc;
    if (int rc = ensure_selftest()) return rc;
    GridDev g;
    if (int rc = check_range(grid, l_begin, l_end, g, "dg_node_positions")) return rc;
    const uint64_t n = l_end - l_begin;
    if (n == 0) return DG_OK;
    if (!x_host) return fail(DG_ERR_INVALID, "dg_node_positions: output is NULL");
    DevBuf<double> d_x;
    DG_CUDA(d_x.alloc(3 * n));
    DG_LAUNCH(k1_launch_node_positions(g, l_begin, n, d_x.p, nullptr));
    DG_CUDA(cudaMemcpy(x_host, d_x.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return DG_OK;
}

int dg_build_cells(const uint32_t res[3], uint64_t c_begin, uint64_t c_end, uint32_t* cells_host)
{
    if (int rc = require_device()) return rc;
    if (!res) return fail(DG_ERR_INVALID, "dg_build_cells: NULL argument");
    dg_grid_desc d;
    const double mn[3] = {0, 0, 0}, mx[3] = {1, 1, 1};
    if (int rc = dg_grid_init(mn, mx, res, &d)) return rc;
    GridDev g; const char* why = "";
    if (!grid_to_dev(&d, g, &why)) return fail(DG_ERR_INVALID, "dg_build_cells: %s", why);
    const uint64_t nc = (uint64_t)res[0] * res[1] * res[2];
    if (c_begin > c_end || c_end > nc) return fail(DG_ERR_INVALID, "dg_build_cells: cell range out of bounds");
    const uint64_t n = c_end - c_begin;
    if (n == 0) return DG_OK;
    if (!cells_host) return fail(DG_ERR_INVALID, "dg_build_cells: output is NULL");
    // same pooled double-buffer pipeline as dg_sample_sdf: the kernel fills one 32 MiB piece (256 Ki cells) of device scratch,
    // the copy stream DMAs it into pinned memory while the CPU moves the previous piece into the caller's array
    std::lock_guard<std::mutex> lock(g_pool.mu);
    const uint64_t piece_cells = HostPathPool::kPiece * sizeof(double) / (32 * sizeof(uint32_t));    // cells per staging buffer
    const uint64_t n_pieces = (n + piece_cells - 1) / piece_cells;
    const uint64_t piece_doubles = std::min<uint64_t>(HostPathPool::kPiece, n * 16);      // 32 uint32 = 16 doubles per cell
    DG_CUDA(g_pool.prepare(2 * piece_doubles, 2));                        // device scratch: two pieces, no larger than the request
    uint32_t* d_piece[2] = {reinterpret_cast<uint32_t*>(g_pool.d_out), reinterpret_cast<uint32_t*>(g_pool.d_out + piece_doubles)};
    for (uint64_t i = 0; i <= n_pieces; i++) {
        if (i < n_pieces) {
            const uint64_t off = i * piece_cells, cnt = (off + piece_cells <= n) ? piece_cells : n - off;
            // piece i reuses device/pinned buffer i&1: its previous user (piece i-2) was fully consumed in iteration i-1
            DG_LAUNCH(k1_launch_build_cells(g, c_begin + off, cnt, d_piece[i & 1], g_pool.s_c));
            DG_CUDA(cudaMemcpyAsync(g_pool.stage[i & 1], d_piece[i & 1], cnt * 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_pool.s_c));
            DG_CUDA(cudaEventRecord(g_pool.dma_ev[i & 1], g_pool.s_c));
        }
        if (i >= 1) {
            const uint64_t off = (i - 1) * piece_cells, cnt = (off + piece_cells <= n) ? piece_cells : n - off;
            DG_CUDA(cudaEventSynchronize(g_pool.dma_ev[(i - 1) & 1]));
            std::memcpy(cells_host + 32 * off, g_pool.stage[(i - 1) & 1], cnt * 32 * sizeof(uint32_t));
        }
    }
    return DG_OK;
}

// ------------------------------------------------------------------------------------------------ fields / K2
static int field_common(const dg_grid_desc* grid, uint64_t n_nodes, dg_field** out, dg_field*& f)
{
    if (!out) return fail(DG_ERR_INVALID, "dg_field_create: out is NULL");
    *out = nullptr;
    if (int rc = require_device()) return rc;
    if (int rc = ensure_selftest()) return rc;
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_field_create: %s", why);
    uint64_t nn = 0;
    if (int rc = dg_grid_num_nodes(grid->resolution, &nn)) return rc;
    (void)n_nodes;
    f = new (std::nothrow) dg_field();
    if (!f) return fail(DG_ERR_NOMEM, "dg_field_create: out of host memory");
    f->desc = *grid; f->dev.g = g;
    return DG_OK;
}

static int field_finish(dg_field* f, const double* d_nodes, const unsigned* d_cells, cudaStream_t stream)
{
    cudaError_t e = cudaGetDevice(&f->device);
    if (e == cudaSuccess) e = f->d_packed.alloc(32 * (f->n_cells_kept ? f->n_cells_kept : 1));
    if (e == cudaSuccess && f->n_cells_kept == 0) e = cudaMemsetAsync(f->d_packed.p, 0, 32 * sizeof(double), stream);
    if (e == cudaSuccess) e = f->d_tab.alloc(f->dev.g.n[0] + f->dev.g.n[1] + f->dev.g.n[2]);
    if (e == cudaSuccess) { e = k2_launch_pack(f->dev.g, d_nodes, d_cells, f->n_cells_kept, f->d_packed.p, stream); g_launches.fetch_add(1); }
    if (e == cudaSuccess) { e = k2_launch_axis_tables(f->dev.g, f->d_tab.p, stream); g_launches.fetch_add(1); }
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? DG_ERR_NOMEM : DG_ERR_CUDA, "dg_field_create: %s", cudaGetErrorString(e));
    f->dev.packed = f->d_packed.p;
    f->dev.cell_map = f->d_cell_map.p;
    f->dev.tab = f->d_tab.p;
    return DG_OK;
}

int dg_field_create(const dg_grid_desc* grid, const double* nodes, uint64_t n_nodes, const uint32_t* cells, uint64_t n_cells_kept,
                    const uint32_t* cell_map, dg_field** out)
{
    dg_field* f = nullptr;
    if (int rc = field_common(grid, n_nodes, out, f)) return rc;
    auto bail = [&](int rc) { delete f; return rc; };
    const uint64_t n_cells = (uint64_t)grid->resolution[0] * grid->resolution[1] * grid->resolution[2];
    if (!nodes && n_nodes) return bail(fail(DG_ERR_INVALID, "dg_field_create: nodes is NULL"));
    if (!cells && n_cells_kept != n_cells) return bail(fail(DG_ERR_INVALID, "dg_field_create: cells == NULL requires n_cells_kept == nx*ny*nz"));
    if (!cell_map && n_cells_kept != n_cells)      // K2/K3 read a NULL cell map as the identity: it must then cover every grid cell
        return bail(fail(DG_ERR_INVALID, "dg_field_create: cell_map == NULL requires n_cells_kept == nx*ny*nz (pass the reduced field's cell map)"));
    if (!cells) {
        uint64_t nn = 0; dg_grid_num_nodes(grid->resolution, &nn);
        if (n_nodes != nn) return bail(fail(DG_ERR_INVALID, "dg_field_create: closed-form cells need all %llu nodes (got %llu)",
                                            (unsigned long long)nn, (unsigned long long)n_nodes));
    } else {
        for (uint64_t i = 0; i < 32 * n_cells_kept; i++)
            if (cells[i] >= n_nodes) return bail(fail(DG_ERR_INVALID, "dg_field_create: cell table references node %u >= n_nodes", cells[i]));
    }
    if (cell_map)
        for (uint64_t i = 0; i < n_cells; i++)
            if (cell_map[i] != UINT32_MAX && cell_map[i] >= n_cells_kept)
                return bail(fail(DG_ERR_INVALID, "dg_field_create: cell_map entry %u >= n_cells_kept", cell_map[i]));
    f->n_nodes = n_nodes; f->n_cells_kept = n_cells_kept;
    DevBuf<double> d_nodes;
    DevBuf<unsigned> d_cells;
    cudaError_t e = d_nodes.alloc(n_nodes);
    if (e == cudaSuccess) e = cudaMemcpy(d_nodes.p, nodes, n_nodes * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && cells) {
        e = d_cells.alloc(32 * n_cells_kept);
        if (e == cudaSuccess) e = cudaMemcpy(d_cells.p, cells, 32 * n_cells_kept * sizeof(unsigned), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && cell_map) {
        e = f->d_cell_map.alloc(n_cells);
        if (e == cudaSuccess) e = cudaMemcpy(f->d_cell_map.p, cell_map, n_cells * sizeof(unsigned), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) return bail(fail(e == cudaErrorMemoryAllocation ? DG_ERR_NOMEM : DG_ERR_CUDA, "dg_field_create: %s", cudaGetErrorString(e)));
    if (int rc = field_finish(f, d_nodes.p, cells ? d_cells.p : nullptr, nullptr)) return bail(rc);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return bail(fail(DG_ERR_CUDA, "dg_field_create: %s", cudaGetErrorString(e)));
    *out = f;
    return DG_OK;
}

int dg_field_create_device(const dg_grid_desc* grid, const double* d_nodes, uint64_t n_nodes, void* stream, dg_field** out)
{
    dg_field* f = nullptr;
    if (int rc = field_common(grid, n_nodes, out, f)) return rc;
    uint64_t nn = 0; dg_grid_num_nodes(grid->resolution, &nn);
    if (!d_nodes || n_nodes != nn) { delete f; return fail(DG_ERR_INVALID, "dg_field_create_device: need all %llu node coefficients", (unsigned long long)nn); }
    f->n_nodes = n_nodes;
    f->n_cells_kept = (uint64_t)grid->resolution[0] * grid->resolution[1] * grid->resolution[2];
    if (int rc = field_finish(f, d_nodes, nullptr, (cudaStream_t)stream)) { delete f; return rc; }
    *out = f;
    return DG_OK;
}

int dg_field_destroy(dg_field* f)
{
    delete f;
    return DG_OK;
}

int dg_field_info(const dg_field* f, uint64_t info[4])
{
    if (!f || !info) return fail(DG_ERR_INVALID, "dg_field_info: NULL argument");
    info[0] = f->n_nodes; info[1] = f->n_cells_kept;
    info[2] = f->d_packed.bytes() + f->d_cell_map.bytes() + f->d_tab.bytes(); info[3] = 0;
    return DG_OK;
}

int dg_interpolate_batch_device(const dg_field* f, const double* d_x, uint64_t n, double* d_phi, double* d_grad, void* stream)
{
    if (!f) return fail(DG_ERR_INVALID, "dg_interpolate_batch: field is NULL");
    if (n && (!d_x || !d_phi)) return fail(DG_ERR_INVALID, "dg_interpolate_batch: x / phi is NULL");
    if (int rc = check_handle_device(f->device, "dg_interpolate_batch")) return rc;
    DG_LAUNCH(k2_launch_interpolate(f->dev, d_x, n, d_phi, d_grad, (cudaStream_t)stream));
    return DG_OK;
}

// Host-buffer path of K2: a three-slot pipeline.  Each slot owns device buffers for one chunk of queries, a stream and an event;
// chunk c runs H2D -> kernel -> D2H on slot c % 3, so the uploads, kernels and downloads of neighbouring chunks overlap on the two
// DMA directions of the link.  Caller memory that is already page-locked (cudaMallocHost / cudaHostRegister / torch pin_memory) is
// DMA'd directly; pageable memory goes through pooled pinned staging buffers, copied in and out by a few host threads while the
// other slots are in flight.  Calls are serialised per process (pool mutex).
namespace {
struct InterpPool {
    static constexpr int kSlots = 3;
    static constexpr uint64_t kChunk = 1u << 19;                 // queries per chunk: 12 MiB up, 16 MiB down
    std::mutex mu;
    int device = -1;
    uint64_t cap = 0;                                            // queries per slot currently allocated
    double* d_x[kSlots] = {}; double* d_phi[kSlots] = {}; double* d_grad[kSlots] = {};
    double* h_in[kSlots] = {}; double* h_out[kSlots] = {};       // pinned staging (allocated on first pageable call)
    cudaStream_t st[kSlots] = {}; cudaEvent_t done[kSlots] = {};
    void release()
    {
        for (int i = 0; i < kSlots; i++) {
            if (d_x[i]) cudaFree(d_x[i]); if (d_phi[i]) cudaFree(d_phi[i]); if (d_grad[i]) cudaFree(d_grad[i]);
            if (h_in[i]) cudaFreeHost(h_in[i]); if (h_out[i]) cudaFreeHost(h_out[i]);
            if (st[i]) cudaStreamDestroy(st[i]); if (done[i]) cudaEventDestroy(done[i]);
            d_x[i] = d_phi[i] = d_grad[i] = h_in[i] = h_out[i] = nullptr; st[i] = nullptr; done[i] = nullptr;
        }
        cap = 0;
    }
    cudaError_t prepare(uint64_t chunk, bool need_stage_in, bool need_stage_out)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev != device) { release(); device = dev; }
        if (chunk > cap) { release(); device = dev; }
        if (cap == 0) {
            for (int i = 0; i < kSlots && e == cudaSuccess; i++) {
                e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_x[i]), chunk * 3 * sizeof(double));
                if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_phi[i]), chunk * sizeof(double));
                if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_grad[i]), chunk * 3 * sizeof(double));
            }
            if (e != cudaSuccess) { release(); return e; }
            cap = chunk;
        }
        for (int i = 0; i < kSlots && e == cudaSuccess; i++) {
            if (need_stage_in && !h_in[i]) e = cudaMallocHost(reinterpret_cast<void**>(&h_in[i]), cap * 3 * sizeof(double));
            if (e == cudaSuccess && need_stage_out && !h_out[i]) e = cudaMallocHost(reinterpret_cast<void**>(&h_out[i]), cap * 4 * sizeof(double));
        }
        if (e != cudaSuccess) release();
        return e;
    }
};
InterpPool g_ipool;

void threaded_copy(void* dst, const void* src, size_t bytes, unsigned nt)
{
    if (nt <= 1 || bytes < (1u << 20)) { std::memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes + nt - 1) / nt + 63) & ~size_t(63);
    for (unsigned k = 1; k < nt; k++) {
        const size_t b = k * per, e = std::min(bytes, b + per);
        if (b < e) th.emplace_back([=]() { std::memcpy(static_cast<char*>(dst) + b, static_cast<const char*>(src) + b, e - b); });
    }
    std::memcpy(dst, src, std::min(bytes, per));
    for (auto& t : th) t.join();
}
}  // namespace

static int dg_interpolate_batch_impl(const dg_field* f, const double* x, uint64_t n, double* phi, double* grad)
{
    if (!f) return fail(DG_ERR_INVALID, "dg_interpolate_batch: field is NULL");
    if (n == 0) return DG_OK;
    if (!x || !phi) return fail(DG_ERR_INVALID, "dg_interpolate_batch: x / phi is NULL");
    if (int rc = check_handle_device(f->device, "dg_interpolate_batch")) return rc;
    std::lock_guard<std::mutex> lock(g_ipool.mu);
    constexpr int S = InterpPool::kSlots;
    const uint64_t chunk = std::min<uint64_t>(InterpPool::kChunk, n);
    const bool x_direct = is_pinned_host(x), out_direct = is_pinned_host(phi) && is_pinned_host(grad);
    DG_CUDA(g_ipool.prepare(std::max<uint64_t>(chunk, g_ipool.cap), !x_direct, !out_direct));
    const unsigned nt = std::max(1u, std::min(4u, host_threads() / 4));      // copy threads of the staged (pageable) path, started per chunk
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    auto unstage = [&](uint64_t c) {                      // results of chunk c: pinned staging -> caller memory
        const uint64_t off = c * chunk, cnt = std::min(chunk, n - off);
        const double* src = g_ipool.h_out[c % S];
        threaded_copy(phi + off, src, cnt * sizeof(double), nt);
        if (grad) threaded_copy(grad + 3 * off, src + g_ipool.cap, 3 * cnt * sizeof(double), nt);
    };
    for (uint64_t c = 0; c < n_chunks; c++) {
        const int s = (int)(c % S);
        const uint64_t off = c * chunk, cnt = std::min(chunk, n - off);
        if (c >= (uint64_t)S) {                           // the slot's previous chunk must have left the device buffers (and staging)
            DG_CUDA(cudaEventSynchronize(g_ipool.done[s]));
            if (!out_direct) unstage(c - S);
        }
        const double* src = x + 3 * off;
        if (!x_direct) { threaded_copy(g_ipool.h_in[s], src, 3 * cnt * sizeof(double), nt); src = g_ipool.h_in[s]; }
        DG_CUDA(cudaMemcpyAsync(g_ipool.d_x[s], src, 3 * cnt * sizeof(double), cudaMemcpyHostToDevice, g_ipool.st[s]));
        DG_LAUNCH(k2_launch_interpolate(f->dev, g_ipool.d_x[s], cnt, g_ipool.d_phi[s], grad ? g_ipool.d_grad[s] : nullptr, g_ipool.st[s]));
        double* dst_phi = out_direct ? phi + off : g_ipool.h_out[s];
        double* dst_grad = out_direct ? grad + 3 * off : g_ipool.h_out[s] + g_ipool.cap;
        DG_CUDA(cudaMemcpyAsync(dst_phi, g_ipool.d_phi[s], cnt * sizeof(double), cudaMemcpyDeviceToHost, g_ipool.st[s]));
        if (grad) DG_CUDA(cudaMemcpyAsync(dst_grad, g_ipool.d_grad[s], 3 * cnt * sizeof(double), cudaMemcpyDeviceToHost, g_ipool.st[s]));
        DG_CUDA(cudaEventRecord(g_ipool.done[s], g_ipool.st[s]));
    }
    for (uint64_t c = (n_chunks > (uint64_t)S ? n_chunks - S : 0); c < n_chunks; c++) {
        DG_CUDA(cudaEventSynchronize(g_ipool.done[c % S]));
        if (!out_direct) unstage(c);
    }
    return DG_OK;
}

int dg_shape_functions(const double* xi, uint64_t n, double* N, double* dN)
{
    if (int rc = require_device()) return rc;
    if (int rc = ensure_selftest()) return rc;
    if (n == 0) return DG_OK;
    if (!xi || !N) return fail(DG_ERR_INVALID, "dg_shape_functions: NULL argument");
    DevBuf<double> d_xi, d_N, d_dN;
    DG_CUDA(d_xi.alloc(3 * n));
    DG_CUDA(d_N.alloc(32 * n));
    if (dN) DG_CUDA(d_dN.alloc(96 * n));
    DG_CUDA(cudaMemcpy(d_xi.p, xi, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    DG_LAUNCH(k2_launch_shape_functions(d_xi.p, n, d_N.p, dN ? d_dN.p : nullptr, nullptr));
    DG_CUDA(cudaMemcpy(N, d_N.p, 32 * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (dN) DG_CUDA(cudaMemcpy(dN, d_dN.p, 96 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return DG_OK;
}

// ------------------------------------------------------------------------------------------------ K3
int dg_density_map_device(const dg_field* f, double h, double rho0, int no_reduction, uint64_t l_begin, uint64_t l_end, double* d_out,
                          void* stream)
{
    if (!f) return fail(DG_ERR_INVALID, "dg_density_map: field is NULL");
    GridDev g;
    if (int rc = check_range(&f->desc, l_begin, l_end, g, "dg_density_map")) return rc;
    if (!(h > 0.0)) return fail(DG_ERR_INVALID, "dg_density_map: smoothing length must be > 0");
    if (l_end > l_begin && !d_out) return fail(DG_ERR_INVALID, "dg_density_map: output is NULL");
    if (int rc = check_handle_device(f->device, "dg_density_map")) return rc;
    DG_LAUNCH(k3_launch_density(f->dev, h, rho0, no_reduction, l_begin, l_end - l_begin, d_out, (cudaStream_t)stream));
    return DG_OK;
}

int dg_density_map(const dg_field* f, double h, double rho0, int no_reduction, uint64_t l_begin, uint64_t l_end, double* out_host)
{
    if (!f) return fail(DG_ERR_INVALID, "dg_density_map: field is NULL");
    if (l_end < l_begin) return fail(DG_ERR_INVALID, "dg_density_map: bad range");
    const uint64_t n = l_end - l_begin;
    if (n && !out_host) return fail(DG_ERR_INVALID, "dg_density_map: output is NULL");
    DevBuf<double> d_out;
    DG_CUDA(d_out.alloc(n));
    if (int rc = dg_density_map_device(f, h, rho0, no_reduction, l_begin, l_end, d_out.p, nullptr)) return rc;
    if (n) DG_CUDA(cudaMemcpy(out_host, d_out.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    return DG_OK;
}

// reduceField with its index passes on the GPU (k4_reduce.cu); the swap-walk compaction and the tie-order-preserving sort stay on the host
static int reduce_field_gpu(const GridDev& g, double* nodes, uint64_t n_nodes, const uint8_t* keep_node, uint32_t* cells, uint64_t n_cells_in,
                            uint32_t* cell_map, uint64_t n_cells_grid, bool force_std_sort, ReduceStats& st)
{
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&]() { const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); t0 = std::chrono::steady_clock::now(); return ms; };
    const uint64_t tile = k4_tile(), n_tiles = (n_cells_in + tile - 1) / tile;
    DevBuf<double> d_nodes, d_nodes_out;
    DevBuf<unsigned char> d_keep, d_cell_keep, d_used;
    DevBuf<unsigned> d_cells, d_cells_out, d_cell_map, d_tile, d_bad, d_perm, d_order, d_new_id;
    DevBuf<unsigned long long> d_key;
    // ---- 1. surviving cells, cell map, compacted rows (:1076-1098)
    DG_CUDA(d_nodes.alloc(n_nodes)); DG_CUDA(d_keep.alloc(n_nodes)); DG_CUDA(d_cells.alloc(32 * n_cells_in)); DG_CUDA(d_cells_out.alloc(32 * n_cells_in));
    DG_CUDA(d_cell_keep.alloc(n_cells_in)); DG_CUDA(d_cell_map.alloc(n_cells_in)); DG_CUDA(d_tile.alloc(n_tiles)); DG_CUDA(d_bad.alloc(1)); DG_CUDA(d_used.alloc(n_nodes));
    DG_CUDA(cudaMemcpyAsync(d_keep.p, keep_node, n_nodes, cudaMemcpyHostToDevice, nullptr));
    DG_CUDA(cudaMemcpyAsync(d_cells.p, cells, 32 * n_cells_in * sizeof(unsigned), cudaMemcpyHostToDevice, nullptr));
    DG_CUDA(cudaMemcpyAsync(d_nodes.p, nodes, n_nodes * sizeof(double), cudaMemcpyHostToDevice, nullptr));
    DG_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(unsigned), nullptr));
    DG_CUDA(cudaMemsetAsync(d_used.p, 0, n_nodes, nullptr));
    DG_LAUNCH(k4_launch_cell_keep(d_cells.p, n_cells_in, d_keep.p, (unsigned)n_nodes, d_cell_keep.p, d_bad.p, nullptr));
    DG_LAUNCH(k4_launch_tile_count(d_cell_keep.p, n_cells_in, d_tile.p, nullptr));
    std::vector<unsigned> tile_cnt(n_tiles);
    unsigned bad = 0;
    if (n_tiles) DG_CUDA(cudaMemcpy(tile_cnt.data(), d_tile.p, n_tiles * sizeof(unsigned), cudaMemcpyDeviceToHost));
    DG_CUDA(cudaMemcpy(&bad, d_bad.p, sizeof bad, cudaMemcpyDeviceToHost));
    if (bad) return fail(DG_ERR_INVALID, "dg_reduce_field: a cell refers to a node id >= n_nodes");
    uint64_t kept_cells = 0;
    for (uint64_t t = 0; t < n_tiles; t++) { const unsigned c = tile_cnt[t]; tile_cnt[t] = (unsigned)kept_cells; kept_cells += c; }     // exclusive scan of ~n/1024 counts
    if (n_tiles) DG_CUDA(cudaMemcpyAsync(d_tile.p, tile_cnt.data(), n_tiles * sizeof(unsigned), cudaMemcpyHostToDevice, nullptr));
    DG_LAUNCH(k4_launch_cell_compact(d_cells.p, n_cells_in, d_cell_keep.p, d_tile.p, d_cells_out.p, d_cell_map.p, nullptr));
    DG_LAUNCH(k4_launch_mark_used(d_cells_out.p, 32 * kept_cells, d_used.p, nullptr));
    if (n_cells_in) DG_CUDA(cudaMemcpyAsync(cell_map, d_cell_map.p, n_cells_in * sizeof(unsigned), cudaMemcpyDeviceToHost, nullptr));
    std::vector<uint8_t> used(n_nodes);
    DG_CUDA(cudaMemcpy(used.data(), d_used.p, n_nodes, cudaMemcpyDeviceToHost));
    for (uint64_t i = n_cells_in; i < n_cells_grid; i++) cell_map[i] = (uint32_t)i;                     // resize + iota (:1076-1078) for cells the field never had
    st.cells_out = kept_cells;
    st.ms_cells = lap();
    // ---- 2. surviving nodes: the reference's swap walk on a permutation (host, serial by nature)
    std::vector<uint32_t> perm;
    const uint64_t m = reduce_swap_walk(used.data(), n_nodes, perm);
    st.nodes_out = m;
    st.ms_nodes = lap();
    // ---- 3. Z-curve keys on the GPU, the order among them on the host (std::sort's tie order)
    DG_CUDA(d_perm.alloc(m)); DG_CUDA(d_key.alloc(m)); DG_CUDA(d_order.alloc(m)); DG_CUDA(d_new_id.alloc(n_nodes)); DG_CUDA(d_nodes_out.alloc(m));
    if (m) DG_CUDA(cudaMemcpyAsync(d_perm.p, perm.data(), m * sizeof(unsigned), cudaMemcpyHostToDevice, nullptr));
    DG_LAUNCH(k4_launch_morton_keys(g, d_perm.p, m, d_key.p, nullptr));
    std::vector<unsigned long long> key(m);
    if (m) DG_CUDA(cudaMemcpy(key.data(), d_key.p, m * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    std::vector<KeyPos> kp(m);
    parallel_for(m, [&](uint64_t b0, uint64_t e0) { for (uint64_t i = b0; i < e0; i++) { kp[i].key = key[i]; kp[i].pos = (uint32_t)i; } });
    std::vector<uint32_t> order;
    reduce_order_survivors(kp, force_std_sort, order, st.tie_path);
    st.ms_sort = lap();
    // ---- 4. new ids, renumbered rows, gathered coefficients (:1158-1173)
    if (m) DG_CUDA(cudaMemcpyAsync(d_order.p, order.data(), m * sizeof(unsigned), cudaMemcpyHostToDevice, nullptr));
    DG_LAUNCH(k4_launch_renumber(d_perm.p, d_order.p, m, d_nodes.p, d_new_id.p, d_nodes_out.p, d_cells_out.p, 32 * kept_cells, nullptr));
    if (kept_cells) DG_CUDA(cudaMemcpyAsync(cells, d_cells_out.p, 32 * kept_cells * sizeof(unsigned), cudaMemcpyDeviceToHost, nullptr));
    if (m) DG_CUDA(cudaMemcpyAsync(nodes, d_nodes_out.p, m * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
    DG_CUDA(cudaDeviceSynchronize());
    st.ms_write = lap();
    return DG_OK;
}

static int dg_reduce_field_impl(const dg_grid_desc* grid, double* nodes, uint64_t n_nodes, const uint8_t* keep_node, uint32_t* cells, uint64_t n_cells_in,
                    uint32_t* cell_map, uint32_t flags, uint64_t* n_nodes_out, uint64_t* n_cells_out, double* timings_ms)
{
    GridDev g; const char* why = "";
    if (!grid_to_dev(grid, g, &why)) return fail(DG_ERR_INVALID, "dg_reduce_field: %s", why);
    const uint64_t n_grid_nodes = (uint64_t)g.nv + 2ull * ((uint64_t)g.ne_x + g.ne_y + g.ne_z);
    const uint64_t n_grid_cells = (uint64_t)g.n[0] * g.n[1] * g.n[2];
    if (!nodes || !keep_node || !cell_map || (n_cells_in && !cells) || !n_nodes_out || !n_cells_out)
        return fail(DG_ERR_INVALID, "dg_reduce_field: NULL argument");
    if (n_nodes != n_grid_nodes)          // node ids must be the grid's own numbering: their positions feed the Z-curve key (:1113)
        return fail(DG_ERR_INVALID, "dg_reduce_field: the field has %llu nodes, the grid %llu (already reduced?)", (unsigned long long)n_nodes,
                    (unsigned long long)n_grid_nodes);
    if (n_cells_in > n_grid_cells) return fail(DG_ERR_INVALID, "dg_reduce_field: more cells than the grid has");
    ReduceStats st; const char* err = "";
    const char* env = std::getenv("DG_REDUCE_FIELD_HOST");
    const bool host_passes = (flags & DG_REDUCE_HOST_PASSES) != 0 || (env && env[0] == '1');
    if (host_passes) {
        if (!reduce_field_host(g, nodes, n_nodes, keep_node, cells, n_cells_in, cell_map, n_grid_cells, (flags & DG_REDUCE_REFERENCE_SORT) != 0, st, &err))
            return fail(DG_ERR_INVALID, "dg_reduce_field: %s", err);
    } else {
        if (int rc = require_device()) return rc;
        if (int rc = reduce_field_gpu(g, nodes, n_nodes, keep_node, cells, n_cells_in, cell_map, n_grid_cells, (flags & DG_REDUCE_REFERENCE_SORT) != 0, st)) return rc;
    }
    *n_nodes_out = st.nodes_out; *n_cells_out = st.cells_out;
    if (timings_ms) { timings_ms[0] = st.ms_cells; timings_ms[1] = st.ms_nodes; timings_ms[2] = st.ms_sort; timings_ms[3] = st.ms_write; timings_ms[4] = (double)st.tie_path; }
    return DG_OK;
}

int dg_obj_read(const char* path, double** vertices, uint64_t* n_vertices, uint32_t** triangles, uint64_t* n_triangles)
{
    if (!path || !vertices || !n_vertices || !triangles || !n_triangles) return fail(DG_ERR_INVALID, "dg_obj_read: NULL argument");
    *vertices = nullptr; *triangles = nullptr; *n_vertices = 0; *n_triangles = 0;
    ObjData d; std::string err;
    if (!read_obj(path, d, err)) return fail(err.rfind("Cannot open", 0) == 0 || err.rfind("short read", 0) == 0 ? DG_ERR_IO : DG_ERR_INVALID, "dg_obj_read: %s", err.c_str());
    double* v = static_cast<double*>(std::malloc(std::max<size_t>(1, d.vertices.size() * sizeof(double))));
    uint32_t* f = static_cast<uint32_t*>(std::malloc(std::max<size_t>(1, d.faces.size() * sizeof(uint32_t))));
    if (!v || !f) { std::free(v); std::free(f); return fail(DG_ERR_NOMEM, "dg_obj_read: out of host memory"); }
    if (!d.vertices.empty()) std::memcpy(v, d.vertices.data(), d.vertices.size() * sizeof(double));
    if (!d.faces.empty()) std::memcpy(f, d.faces.data(), d.faces.size() * sizeof(uint32_t));
    *vertices = v; *triangles = f; *n_vertices = d.vertices.size() / 3; *n_triangles = d.faces.size() / 3;
    return DG_OK;
}

void dg_obj_free(double* vertices, uint32_t* triangles)
{
    std::free(vertices); std::free(triangles);
}


// ---- entry points whose bodies use host containers / threads: exceptions are turned into status codes
int dg_add_function_sdf(const dg_mesh* m, const dg_grid_desc* grid, double sign, double* nodes_host, uint32_t* cells_host, uint32_t* cell_map_host,
                        double* timings_ms)
{
    return guarded("dg_add_function_sdf", [&]() { return dg_add_function_sdf_impl(m, grid, sign, nodes_host, cells_host, cell_map_host, timings_ms); });
}

int dg_add_function_sdf_multi(dg_mesh_group* grp, const dg_grid_desc* grid, double sign, double* nodes_host, uint32_t* cells_host,
                              uint32_t* cell_map_host, double* timings_ms)
{
    return guarded("dg_add_function_sdf_multi", [&]() { return dg_add_function_sdf_multi_impl(grp, grid, sign, nodes_host, cells_host, cell_map_host, timings_ms); });
}

int dg_interpolate_batch(const dg_field* f, const double* x, uint64_t n, double* phi, double* grad)
{
    return guarded("dg_interpolate_batch", [&]() { return dg_interpolate_batch_impl(f, x, n, phi, grad); });
}

int dg_reduce_field(const dg_grid_desc* grid, double* nodes, uint64_t n_nodes, const uint8_t* keep_node, uint32_t* cells, uint64_t n_cells_in,
                    uint32_t* cell_map, uint32_t flags, uint64_t* n_nodes_out, uint64_t* n_cells_out, double* timings_ms)
{
    return guarded("dg_reduce_field", [&]() { return dg_reduce_field_impl(grid, nodes, n_nodes, keep_node, cells, n_cells_in, cell_map, flags, n_nodes_out, n_cells_out, timings_ms); });
}

}  // extern "C"
