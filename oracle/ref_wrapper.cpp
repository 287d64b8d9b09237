// oracle/ref_wrapper.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" wrapper around the reference's own, unmodified
// discregrid/include/Discregrid/geometry/TriangleMeshDistance.h (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libdgref.so).  Used to pin oracle/dg_oracle.cpp and, when present,
// as the "reference" CPU baseline of bench.py.  `private` is opened only to dump the tree and the
// pseudonormals for comparison; no reference code is modified.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <cmath>
#define private public
#include "geometry/TriangleMeshDistance.h"
#undef private
#ifdef _OPENMP
#include <omp.h>
#endif

using Discregrid::TriangleMeshDistance;

extern "C" {

void* ref_mesh_create(const double* V, uint64_t nV, const uint32_t* F, uint64_t nT)
{
    if (nT == 0) return nullptr;  // the reference would exit(-1) (TriangleMeshDistance.h:338-341)
    Discregrid::TriangleMesh mesh;
    mesh.m_v.resize(nV);
    for (uint64_t i = 0; i < nV; i++) mesh.m_v[i] = {V[3 * i], V[3 * i + 1], V[3 * i + 2]};
    mesh.m_f.resize(nT);
    for (uint64_t i = 0; i < nT; i++) mesh.m_f[i] = {F[3 * i], F[3 * i + 1], F[3 * i + 2]};
    return new TriangleMeshDistance(mesh);     // the ctor GenerateSDF uses, cmd/generate_sdf/main.cpp:74
}
void ref_mesh_destroy(void* h) { delete (TriangleMeshDistance*)h; }
uint64_t ref_mesh_num_nodes(void* h) { return ((TriangleMeshDistance*)h)->nodes.size(); }
void ref_mesh_tree(void* h, double* spheres, int32_t* kids)
{
    auto* m = (TriangleMeshDistance*)h;
    for (size_t i = 0; i < m->nodes.size(); i++) {
        const auto& n = m->nodes[i];
        double* s = spheres + 8 * i;
        for (int d = 0; d < 3; d++) { s[d] = n.bv_left.center[d]; s[4 + d] = n.bv_right.center[d]; }
        s[3] = n.bv_left.radius; s[7] = n.bv_right.radius;
        kids[2 * i] = n.left; kids[2 * i + 1] = n.right;
    }
}
void ref_mesh_pseudonormals(void* h, double* tri, double* edge, double* vert)
{
    auto* m = (TriangleMeshDistance*)h;
    for (size_t i = 0; i < m->pseudonormals_triangles.size(); i++)
        for (int d = 0; d < 3; d++) tri[3 * i + d] = m->pseudonormals_triangles[i][d];
    for (size_t i = 0; i < m->pseudonormals_edges.size(); i++)
        for (int e = 0; e < 3; e++) for (int d = 0; d < 3; d++) edge[9 * i + 3 * e + d] = m->pseudonormals_edges[i][e][d];
    for (size_t i = 0; i < m->pseudonormals_vertices.size(); i++)
        for (int d = 0; d < 3; d++) vert[3 * i + d] = m->pseudonormals_vertices[i][d];
}
void ref_mesh_distance(void* h, const double* x, uint64_t n, int is_signed, double* dist, double* nearest,
                       int32_t* entity, int32_t* tri)
{
    const auto* m = (const TriangleMeshDistance*)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (long long q = 0; q < (long long)n; q++) {
        const std::array<double, 3> p = {x[3 * q], x[3 * q + 1], x[3 * q + 2]};
        const Discregrid::Result r = is_signed ? m->signed_distance(p) : m->unsigned_distance(p);
        dist[q] = r.distance;
        if (nearest) for (int d = 0; d < 3; d++) nearest[3 * q + d] = r.nearest_point[d];
        if (entity) entity[q] = (int)r.nearest_entity;
        if (tri) tri[q] = r.triangle_id;
    }
}
// The addFunction node loop (cubic_lagrange_discrete_grid.cpp:806-817, schedule(static)) with the
// GenerateSDF functor (cmd/generate_sdf/main.cpp:97,101) over pre-computed node positions x[n][3]
// (positions come from the oracle's indexToNodePosition restatement; Eigen blocks the grid .cpp).
void ref_sample_points(void* h, const double* x, uint64_t n, double sign, double* out, int nthreads)
{
    const auto* m = (const TriangleMeshDistance*)h;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)n; q++) {
        const std::array<double, 3> p = {x[3 * q], x[3 * q + 1], x[3 * q + 2]};
        const double d = m->signed_distance(p).distance;
        out[q] = (sign == 1.0) ? d : sign * d;
    }
}
}
