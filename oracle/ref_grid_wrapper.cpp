// oracle/ref_grid_wrapper.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" access to the reference's own Discregrid::CubicLagrangeDiscreteGrid (its unmodified discrete_grid.cpp and
// cubic_lagrange_discrete_grid.cpp, compiled by oracle/Makefile against the Eigen stand-in oracle/ref_eigen).  Used to pin the
// oracle's interpolate / shape functions to the reference's code, to generate golden vectors, and as the "reference" CPU
// baseline of interpolate (the OpenMP pixel-loop pattern of cmd/discrete_field_to_bitmap/main.cpp:118-135).
// every standard / Eigen header the reference headers pull in is included BEFORE `private` is opened, so only the reference's own
// class bodies are affected
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <cmath>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <Eigen/Dense>
#define private public          // only to copy m_nodes / m_cells / m_cell_map out after the reference's own addFunction; no reference code is modified
#include <Discregrid/All>
#undef private
#include <chrono>
#include <cstring>
#include <cstdint>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace Eigen;
using Discregrid::CubicLagrangeDiscreteGrid;

extern "C" {

void* refg_load(const char* path) { return new CubicLagrangeDiscreteGrid(std::string(path)); }
void refg_destroy(void* h) { delete (CubicLagrangeDiscreteGrid*)h; }
void refg_info(void* h, double* dom6, uint32_t* res3, double* cell3, double* inv3, uint64_t* n_cells)
{
    auto* g = (CubicLagrangeDiscreteGrid*)h;
    for (int d = 0; d < 3; d++) { dom6[d] = g->domain().min()[d]; dom6[3 + d] = g->domain().max()[d]; res3[d] = g->resolution()[d];
                                  cell3[d] = g->cellSize()[d]; inv3[d] = g->invCellSize()[d]; }
    *n_cells = g->nCells();
}
// batched interpolate(field, x, grad*): grad nullable; the gradient buffer is pre-zeroed by the caller (the reference leaves it
// untouched when it returns before the loop, :981-982 / :993-994)
void refg_interpolate(void* h, unsigned field, const double* x, uint64_t n, double* phi, double* grad, int nthreads)
{
    auto* g = (CubicLagrangeDiscreteGrid*)h;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)n; q++) {
        const Vector3d p(x[3 * q], x[3 * q + 1], x[3 * q + 2]);
        if (grad) {
            Vector3d gq(grad[3 * q], grad[3 * q + 1], grad[3 * q + 2]);
            phi[q] = g->interpolate(field, p, &gq);
            for (int d = 0; d < 3; d++) grad[3 * q + d] = gq[d];
        } else {
            phi[q] = g->interpolate(field, p);
        }
    }
}
// determineShapeFunctions + split interpolate (:901-975) for n points: ok[n], N[n x 32], dN[n x 32 x 3], c0[n x 3], cell[n x 32], phi2[n]
void refg_split(void* h, unsigned field, const double* x, uint64_t n, int32_t* ok, double* N, double* dN, double* c0, uint32_t* cell, double* phi2,
                double* grad2)
{
    auto* g = (CubicLagrangeDiscreteGrid*)h;
    for (uint64_t q = 0; q < n; q++) {
        const Vector3d p(x[3 * q], x[3 * q + 1], x[3 * q + 2]);
        std::array<unsigned int, 32> c; Vector3d cc; Matrix<double, 32, 1> Nq; Matrix<double, 32, 3> dNq;
        ok[q] = g->determineShapeFunctions(field, p, c, cc, Nq, &dNq) ? 1 : 0;
        if (!ok[q]) continue;
        for (int j = 0; j < 32; j++) { N[32 * q + j] = Nq[j]; cell[32 * q + j] = c[j]; for (int d = 0; d < 3; d++) dN[96 * q + 3 * j + d] = dNq(j, d); }
        for (int d = 0; d < 3; d++) c0[3 * q + d] = cc[d];
        Vector3d gq = Vector3d::Zero();
        phi2[q] = g->interpolate(field, p, c, cc, Nq, &gq, &dNq);
        for (int d = 0; d < 3; d++) grad2[3 * q + d] = gq[d];
    }
}
// reduceField (:1065-1174) with the value-window predicate of cmd/generate_density_map/main.cpp:141-144 (lo <= v <= hi), timed;
// the result is read back through save() (the members have no accessors).
double refg_reduce_window(void* h, unsigned field, double lo, double hi)
{
    auto* g = (CubicLagrangeDiscreteGrid*)h;
    const auto t0 = std::chrono::steady_clock::now();
    g->reduceField(field, [lo, hi](Vector3d const&, double v) { return lo <= v && v <= hi; });
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// The reference's REAL addFunction with the GenerateSDF functor, exactly as cmd/generate_sdf/main.cpp:74,92-105 drives it:
// TriangleMeshDistance md(vertices, triangles) (the std::vector overload, TriangleMeshDistance.h:251-267 -- the one the
// TriangleMesh ctor forwards to, :227-230), CubicLagrangeDiscreteGrid sdf(domain, resolution), func = [&md](x){ return md.signed_distance(x).distance; }
// (or -1.0 * for --invert), sdf.addFunction(func, verbose=false).  Only the addFunction call is timed (steady_clock), i.e. the
// OpenMP node loop :806-831, the serial connectivity loop :833-886 and the cell map :888-891 -- what SURVEY 8(d) calls the
// addFunction wall-clock.  The mesh-distance object is built once and reused (it is outside addFunction's timer in the reference too).
void* refg_md_create(const double* V, uint64_t nV, const uint32_t* F, uint64_t nT)
{
    std::vector<std::array<double, 3>> v(nV);
    std::vector<std::array<int, 3>> f(nT);
    for (uint64_t i = 0; i < nV; i++) v[i] = {V[3 * i], V[3 * i + 1], V[3 * i + 2]};
    for (uint64_t i = 0; i < nT; i++) f[i] = {(int)F[3 * i], (int)F[3 * i + 1], (int)F[3 * i + 2]};
    return new Discregrid::TriangleMeshDistance(v, f);
}
void refg_md_destroy(void* md) { delete (Discregrid::TriangleMeshDistance*)md; }
// returns the seconds addFunction took; nodes_out (nullable) receives m_nodes[0], cells_out (nullable) m_cells[0], n_nodes_out the count
double refg_add_function_sdf(void* md_, const double* dom_min, const double* dom_max, const uint32_t* res, int invert, int nthreads,
                             double* nodes_out, uint32_t* cells_out, uint64_t* n_nodes_out)
{
    auto& md = *(Discregrid::TriangleMeshDistance*)md_;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    Eigen::AlignedBox3d domain(Vector3d(dom_min[0], dom_min[1], dom_min[2]), Vector3d(dom_max[0], dom_max[1], dom_max[2]));
    std::array<unsigned int, 3> resolution = {{res[0], res[1], res[2]}};
    CubicLagrangeDiscreteGrid sdf(domain, resolution);
    auto func = Discregrid::DiscreteGrid::ContinuousFunction{};
    if (invert) func = [&md](Vector3d const& xi) { return -1.0 * md.signed_distance(xi).distance; };
    else        func = [&md](Vector3d const& xi) { return md.signed_distance(xi).distance; };
    const auto t0 = std::chrono::steady_clock::now();
    sdf.addFunction(func, false);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (n_nodes_out) *n_nodes_out = sdf.m_nodes[0].size();
    if (nodes_out) std::memcpy(nodes_out, sdf.m_nodes[0].data(), sdf.m_nodes[0].size() * sizeof(double));
    if (cells_out) std::memcpy(cells_out, sdf.m_cells[0].data(), sdf.m_cells[0].size() * 32 * sizeof(uint32_t));
    return dt;
}
// A reference grid object holding a given coefficient field, built in memory (a 256^3 .cdf is 3 GB on disk): the reference's own
// addFunction runs with a constant functor -- that creates m_nodes[0], the connectivity m_cells[0] (:833-886) and m_cell_map[0] (:888-891)
// exactly as GenerateSDF would -- and the coefficients are then overwritten with `nodes`.  Used as the interpolate CPU baseline on the
// same field the GPU leg uses.
void* refg_grid_from_nodes(const double* dom_min, const double* dom_max, const uint32_t* res, const double* nodes, uint64_t n_nodes, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    Eigen::AlignedBox3d domain(Vector3d(dom_min[0], dom_min[1], dom_min[2]), Vector3d(dom_max[0], dom_max[1], dom_max[2]));
    std::array<unsigned int, 3> resolution = {{res[0], res[1], res[2]}};
    auto* g = new CubicLagrangeDiscreteGrid(domain, resolution);
    g->addFunction([](Vector3d const&) { return 0.0; }, false);
    if (g->m_nodes[0].size() != n_nodes) { delete g; return nullptr; }
    std::memcpy(g->m_nodes[0].data(), nodes, n_nodes * sizeof(double));
    return g;
}
int refg_omp_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void refg_save(void* h, const char* path) { ((CubicLagrangeDiscreteGrid*)h)->save(std::string(path)); }
// the reference's OBJ loader, Discregrid::TriangleMesh(path) (src/mesh/triangle_mesh.cpp:90-124), timed
void* refm_open(const char* path, double* seconds)
{
    const auto t0 = std::chrono::steady_clock::now();
    auto* m = new Discregrid::TriangleMesh(std::string(path));
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return m;
}
void refm_sizes(void* h, uint64_t* nv, uint64_t* nf) { auto* m = (Discregrid::TriangleMesh*)h; *nv = m->nVertices(); *nf = m->nFaces(); }
void refm_copy(void* h, double* V, uint32_t* F)
{
    auto* m = (Discregrid::TriangleMesh*)h;
    for (std::size_t i = 0; i < m->nVertices(); i++) for (int d = 0; d < 3; d++) V[3 * i + d] = m->vertex_data()[i][d];
    for (std::size_t i = 0; i < m->nFaces(); i++) for (int d = 0; d < 3; d++) F[3 * i + d] = m->face_data()[i][d];
}
void refm_close(void* h) { delete (Discregrid::TriangleMesh*)h; }
}
