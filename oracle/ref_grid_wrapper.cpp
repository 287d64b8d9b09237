// oracle/ref_grid_wrapper.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" access to the reference's own Discregrid::CubicLagrangeDiscreteGrid (its unmodified discrete_grid.cpp and
// cubic_lagrange_discrete_grid.cpp, compiled by oracle/Makefile against the Eigen stand-in oracle/ref_eigen).  Used to pin the
// oracle's interpolate / shape functions to the reference's code, to generate golden vectors, and as the "reference" CPU
// baseline of interpolate (the OpenMP pixel-loop pattern of cmd/discrete_field_to_bitmap/main.cpp:118-135).
#include <Discregrid/All>
#include <chrono>
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
