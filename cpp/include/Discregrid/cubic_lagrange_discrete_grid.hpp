// Discregrid::CubicLagrangeDiscreteGrid -- the reference's public surface and members
// (discregrid/include/Discregrid/cubic_lagrange_discrete_grid.hpp:8-72) with the hot path on the B200:
//
//   addFunction(MeshSignedDistanceFunction)  -> dg_sample_sdf   (K1)   replaces the OpenMP node loop :806-831
//   addFunction(DensityMapFunction)          -> dg_density_map  (K3)   (GenerateDensityMap's functor + predicate)
//   interpolate(field, x, grad*)             -> dg_interpolate_batch (K2), also batched: interpolate(field, n, x, phi, grad)
//   connectivity table of addFunction        -> dg_build_cells  (GPU)  replaces the serial loop :833-886
//
// m_nodes / m_cells / m_cell_map hold exactly what the reference would have produced, so save()/load() stay
// byte-compatible (.cdf/.cdm, :678-778) and reduceField() (host bookkeeping, restated from :1065-1174) keeps working.
// Device copies of fields are caches owned by this object, dropped by addFunction / load / reduceField.
//
// addFunction takes the reference's opaque std::function.  Functors of the two types the reference's own tools build are
// recognised with std::function::target<>() and run on the GPU.  Any other callable cannot run on a GPU: it throws
// std::invalid_argument unless DISCREGRID_B200_ALLOW_HOST_CALLBACK is defined, in which case YOUR callback is invoked
// once per node on the host (node positions still come from the GPU); that path is never part of any reported number.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <stdexcept>
#include "discrete_grid.hpp"
#include "geometry/TriangleMeshDistance.h"
#include "discregrid_b200.h"

#if defined(__GNUC__) && !defined(__clang__)
#define DG_NO_CONTRACT __attribute__((optimize("fp-contract=off")))
#else
#define DG_NO_CONTRACT
#endif

namespace Discregrid {

class CubicLagrangeDiscreteGrid;

// std::vector whose resize() leaves trivially-constructible elements uninitialised.  The field arrays (947 MB of coefficients and a
// 2 GiB connectivity table at 256^3) are written exactly once, by the library: value-initialising them first would touch every page
// serially before the real fill does.  Same interface as std::vector; the members are private in the reference too
// (cubic_lagrange_discrete_grid.hpp:69-71), so no caller sees the allocator.
template <class T>
struct FieldAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = FieldAllocator<U>; };
    using std::allocator<T>::allocator;
    FieldAllocator() = default;
    template <class U> FieldAllocator(FieldAllocator<U> const&) noexcept {}
    template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using FieldVector = std::vector<T, FieldAllocator<T>>;

// GenerateDensityMap's density functor + sample predicate (cmd/generate_density_map/main.cpp:96-133) as a recognisable type
struct DensityMapFunction {
    const CubicLagrangeDiscreteGrid* grid;
    unsigned int sdf_field_id;
    double smoothing_length, rest_density;
    bool no_reduction;
    double operator()(Eigen::Vector3d const&) const { throw std::logic_error("DensityMapFunction is evaluated by addFunction on the GPU"); }
};

class CubicLagrangeDiscreteGrid : public DiscreteGrid {
public:
    CubicLagrangeDiscreteGrid(std::string const& filename) { load(filename); }
    CubicLagrangeDiscreteGrid(Eigen::AlignedBox3d const& domain, std::array<unsigned int, 3> const& resolution) : DiscreteGrid(domain, resolution) {}
    CubicLagrangeDiscreteGrid(const CubicLagrangeDiscreteGrid&) = delete;
    CubicLagrangeDiscreteGrid& operator=(const CubicLagrangeDiscreteGrid&) = delete;
    ~CubicLagrangeDiscreteGrid() override { invalidate(); }

    // ---- save / load: byte layout of cubic_lagrange_discrete_grid.cpp:678-719 / 721-778 (serialize.hpp:12-25)
    void save(std::string const& filename) const override
    {
        std::ofstream out(filename, std::ios::binary);
        auto put = [&](const void* p, std::size_t n) { out.write(static_cast<const char*>(p), (std::streamsize)n); };
        put(m_domain.min().data(), 24); put(m_domain.max().data(), 24);
        put(m_resolution.data(), 12); put(m_cell_size.data(), 24); put(m_inv_cell_size.data(), 24);
        put(&m_n_cells, 8); put(&m_n_fields, 8);
        std::size_t n = m_nodes.size(); put(&n, 8);
        for (auto const& v : m_nodes) { n = v.size(); put(&n, 8); put(v.data(), 8 * n); }
        n = m_cells.size(); put(&n, 8);
        for (auto const& v : m_cells) { n = v.size(); put(&n, 8); put(v.data(), 128 * n); }
        n = m_cell_map.size(); put(&n, 8);
        for (auto const& v : m_cell_map) { n = v.size(); put(&n, 8); put(v.data(), 4 * n); }
    }
    void load(std::string const& filename) override
    {
        invalidate();
        std::ifstream in(filename, std::ios::binary);
        if (!in.good()) { std::cerr << "ERROR: Discrete grid can not be loaded. Input file does not exist!" << std::endl; return; }
        auto get = [&](void* p, std::size_t n) { in.read(static_cast<char*>(p), (std::streamsize)n); };
        get(m_domain.min().data(), 24); get(m_domain.max().data(), 24);
        get(m_resolution.data(), 12); get(m_cell_size.data(), 24); get(m_inv_cell_size.data(), 24);
        get(&m_n_cells, 8); get(&m_n_fields, 8);
        std::size_t n = 0; get(&n, 8); m_nodes.resize(n);
        for (auto& v : m_nodes) { get(&n, 8); v.resize(n); get(v.data(), 8 * n); }
        get(&n, 8); m_cells.resize(n);
        for (auto& v : m_cells) { get(&n, 8); v.resize(n); get(v.data(), 128 * n); }
        get(&n, 8); m_cell_map.resize(n);
        for (auto& v : m_cell_map) { get(&n, 8); v.resize(n); get(v.data(), 4 * n); }
    }

    // ---- addFunction (cubic_lagrange_discrete_grid.cpp:780-899)
    unsigned int addFunction(ContinuousFunction const& func, bool verbose = false, SamplePredicate const& pred = nullptr) override
    {
        const dg_grid_desc d = desc();
        std::uint64_t n_nodes = 0;
        check(dg_grid_num_nodes(d.resolution, &n_nodes));
        if (auto const* sdf = func.target<MeshSignedDistanceFunction>()) {
            if (pred) throw std::invalid_argument("addFunction: a sample predicate is only supported with DensityMapFunction");
            // the whole of :780-899 in one library call: node loop on the GPU; connectivity (:833-886) and cell map (:888-891) written by
            // the library's host threads into these arrays while the GPU works
            FieldVector<double> coeffs(n_nodes);
            FieldVector<std::array<unsigned int, 32>> cells(m_n_cells);
            FieldVector<unsigned int> cell_map(m_n_cells);
            if (sdf->md->group())             // TriangleMeshDistance::useGpus(n): the same call spread over n GPUs
                check(dg_add_function_sdf_multi(sdf->md->group(), &d, sdf->sign, coeffs.data(), reinterpret_cast<std::uint32_t*>(cells.data()), cell_map.data(),
                                                m_last_add_function_ms));
            else
                check(dg_add_function_sdf(sdf->md->handle(), &d, sdf->sign, coeffs.data(), reinterpret_cast<std::uint32_t*>(cells.data()), cell_map.data(),
                                          m_last_add_function_ms));
            if (verbose) std::cout << "Construction: " << n_nodes << " nodes sampled on the GPU in " << m_last_add_function_ms[0] << " ms" << std::endl;
            invalidate();
            m_nodes.push_back(std::move(coeffs)); m_cells.push_back(std::move(cells)); m_cell_map.push_back(std::move(cell_map));
            return static_cast<unsigned int>(m_n_fields++);
        }
        FieldVector<double> coeffs(n_nodes);
        if (auto const* dm = func.target<DensityMapFunction>()) {
            check(dg_density_map(dm->grid->deviceField(dm->sdf_field_id), dm->smoothing_length, dm->rest_density, dm->no_reduction ? 1 : 0,
                                 0, n_nodes, coeffs.data()));
        } else {
#ifdef DISCREGRID_B200_ALLOW_HOST_CALLBACK
            FieldVector<double> x(3 * n_nodes);
            check(dg_node_positions(&d, 0, n_nodes, x.data()));
            for (std::uint64_t l = 0; l < n_nodes; l++) {
                const Eigen::Vector3d p(x[3 * l], x[3 * l + 1], x[3 * l + 2]);
                coeffs[l] = (!pred || pred(p)) ? func(p) : std::numeric_limits<double>::max();
            }
#else
            (void)pred;
            throw std::invalid_argument("CubicLagrangeDiscreteGrid::addFunction: only MeshSignedDistanceFunction / DensityMapFunction run on the GPU "
                                        "(define DISCREGRID_B200_ALLOW_HOST_CALLBACK to have other callables invoked per node on the host)");
#endif
        }
        if (verbose) std::cout << "Construction: " << n_nodes << " nodes sampled on the GPU" << std::endl;
        return addSampledFunction(std::move(coeffs));
    }
    // appends a field from node values + the closed-form connectivity (:833-886, built on the GPU) + identity cell map (:888-891)
    unsigned int addSampledFunction(FieldVector<double> coeffs)
    {
        invalidate();
        m_nodes.push_back(std::move(coeffs));
        m_cells.emplace_back(m_n_cells);
        check(dg_build_cells(m_resolution.data(), 0, m_n_cells, reinterpret_cast<std::uint32_t*>(m_cells.back().data())));
        m_cell_map.emplace_back(m_n_cells);
        std::iota(m_cell_map.back().begin(), m_cell_map.back().end(), 0u);
        return static_cast<unsigned int>(m_n_fields++);
    }

    std::size_t nCells() const { return m_n_cells; }

    // ---- interpolate (cubic_lagrange_discrete_grid.cpp:977-1063), scalar form kept for source compatibility
    double interpolate(unsigned int field_id, Eigen::Vector3d const& xi, Eigen::Vector3d* gradient = nullptr) const override
    {
        double phi = 0.0, g[3] = {0, 0, 0};
        // the reference leaves *gradient untouched when it returns DBL_MAX before the loop (:981-982, :993-994)
        check(dg_interpolate_batch(deviceField(field_id), xi.data(), 1, &phi, gradient ? g : nullptr));
        if (gradient && !(phi == std::numeric_limits<double>::max() && !in_kept_cell(field_id, xi))) *gradient = Eigen::Vector3d(g[0], g[1], g[2]);
        return phi;
    }
    // batched form (new; the one to use): x = n x 3, phi = n, grad = n x 3 or nullptr
    void interpolate(unsigned int field_id, std::size_t n, const double* x, double* phi, double* grad = nullptr) const
    {
        check(dg_interpolate_batch(deviceField(field_id), x, n, phi, grad));
    }

    // ---- split API (:901-975): shape functions on the GPU (dg_shape_functions), index algebra on the host
    DG_NO_CONTRACT bool determineShapeFunctions(unsigned int field_id, Eigen::Vector3d const& x, std::array<unsigned int, 32>& cell, Eigen::Vector3d& c0,
                                                Eigen::Matrix<double, 32, 1>& N, Eigen::Matrix<double, 32, 3>* dN = nullptr) const override
    {
        if (!m_domain.contains(x)) return false;
        unsigned mi[3];
        for (int d = 0; d < 3; d++) {
            mi[d] = static_cast<unsigned int>((x[d] - m_domain.min()[d]) * m_inv_cell_size[d]);
            if (mi[d] >= m_resolution[d]) mi[d] = m_resolution[d] - 1;
        }
        const unsigned i = multiToSingleIndex({{mi[0], mi[1], mi[2]}});
        const unsigned i_ = m_cell_map[field_id][i];
        if (i_ == std::numeric_limits<unsigned int>::max()) return false;
        double xi[3];
        for (int d = 0; d < 3; d++) {
            const double lo = m_domain.min()[d] + static_cast<double>(mi[d]) * m_cell_size[d];
            const double hi = lo + m_cell_size[d];
            const double denom = hi - lo;
            c0[d] = 2.0 / denom;
            const double c1 = (hi + lo) / denom;
            xi[d] = c0[d] * x[d] - c1;
        }
        cell = m_cells[field_id][i_];
        double Nb[32], dNb[96];
        check(dg_shape_functions(xi, 1, Nb, dN ? dNb : nullptr));
        for (int j = 0; j < 32; j++) N[j] = Nb[j];
        if (dN) for (int j = 0; j < 32; j++) for (int d = 0; d < 3; d++) (*dN)(j, d) = dNb[3 * j + d];
        return true;
    }
    DG_NO_CONTRACT double interpolate(unsigned int field_id, Eigen::Vector3d const&, const std::array<unsigned int, 32>& cell, const Eigen::Vector3d& c0,
                                      const Eigen::Matrix<double, 32, 1>& N, Eigen::Vector3d* gradient = nullptr, Eigen::Matrix<double, 32, 3>* dN = nullptr) const override
    {
        double phi = 0.0;
        if (gradient) gradient->setZero();
        for (unsigned j = 0; j < 32u; ++j) {
            const double c = m_nodes[field_id][cell[j]];
            if (c == std::numeric_limits<double>::max()) { if (gradient) gradient->setZero(); return std::numeric_limits<double>::max(); }
            phi = phi + c * N[j];
            if (gradient) for (int d = 0; d < 3; d++) (*gradient)[d] = (*gradient)[d] + c * (*dN)(j, d);
        }
        if (gradient) for (int d = 0; d < 3; d++) (*gradient)[d] = (*gradient)[d] * c0[d];
        return phi;
    }

    // ---- reduceField (:1065-1174): the predicate is evaluated here; cells / nodes / cell map are rewritten by dg_reduce_field
    // and end up exactly as the reference leaves them (kept cells in order, surviving nodes in its Z-curve order)
    void reduceField(unsigned int field_id, Predicate pred) override;

    void forEachCell(unsigned int, std::function<void(unsigned int, Eigen::AlignedBox3d const&, unsigned int)> const& cb) const
    {
        const unsigned n = m_resolution[0] * m_resolution[1] * m_resolution[2];
        for (unsigned i = 0; i < n; ++i) cb(i, subdomain(i), 0);
    }

    // ---- accessors used by tools / tests
    FieldVector<double> const& nodeData(unsigned int f) const { return m_nodes[f]; }
    FieldVector<std::array<unsigned int, 32>> const& cellData(unsigned int f) const { return m_cells[f]; }
    FieldVector<unsigned int> const& cellMap(unsigned int f) const { return m_cell_map[f]; }
    // ms since entry of the last addFunction(MeshSignedDistanceFunction): total, node pipeline, index tables, pre-fault; workers
    const double* lastAddFunctionTimings() const { return m_last_add_function_ms; }
    std::size_t nFields() const { return m_n_fields; }
    dg_grid_desc desc() const
    {
        dg_grid_desc d; std::memset(&d, 0, sizeof d);
        for (int k = 0; k < 3; k++) { d.domain_min[k] = m_domain.min()[k]; d.domain_max[k] = m_domain.max()[k]; d.resolution[k] = m_resolution[k];
                                      d.cell_size[k] = m_cell_size[k]; d.inv_cell_size[k] = m_inv_cell_size[k]; }
        return d;
    }
    const dg_field* deviceField(unsigned int field_id) const
    {
        auto it = m_dev.find(field_id);
        if (it != m_dev.end()) return it->second;
        const dg_grid_desc d = desc();
        dg_field* f = nullptr;
        check(dg_field_create(&d, m_nodes.at(field_id).data(), m_nodes[field_id].size(), reinterpret_cast<const std::uint32_t*>(m_cells[field_id].data()),
                              m_cells[field_id].size(), m_cell_map[field_id].data(), &f));
        m_dev[field_id] = f;
        return f;
    }

private:
    std::vector<FieldVector<double>> m_nodes;
    std::vector<FieldVector<std::array<unsigned int, 32>>> m_cells;
    std::vector<FieldVector<unsigned int>> m_cell_map;
    double m_last_add_function_ms[6] = {0, 0, 0, 0, 0, 0};
    mutable std::map<unsigned int, dg_field*> m_dev;

    static void check(int rc) { if (rc != DG_OK) throw std::runtime_error(std::string("discregrid_b200: ") + dg_last_error()); }
    void invalidate() { for (auto& kv : m_dev) dg_field_destroy(kv.second); m_dev.clear(); }
    bool in_kept_cell(unsigned int field_id, Eigen::Vector3d const& x) const
    {
        if (!m_domain.contains(x)) return false;
        unsigned mi[3];
        for (int d = 0; d < 3; d++) { mi[d] = static_cast<unsigned int>((x[d] - m_domain.min()[d]) * m_inv_cell_size[d]); if (mi[d] >= m_resolution[d]) mi[d] = m_resolution[d] - 1; }
        return m_cell_map[field_id][multiToSingleIndex({{mi[0], mi[1], mi[2]}})] != std::numeric_limits<unsigned int>::max();
    }
};

inline void CubicLagrangeDiscreteGrid::reduceField(unsigned int field_id, Predicate pred)
{
    invalidate();
    auto& coeffs = m_nodes[field_id];
    auto& cells = m_cells[field_id];
    const dg_grid_desc d = desc();
    std::vector<double> pos(3 * coeffs.size());
    check(dg_node_positions(&d, 0, coeffs.size(), pos.data()));              // indexToNodePosition for every node (:604-665)
    const double dbl_max = std::numeric_limits<double>::max();
    std::vector<std::uint8_t> keep(coeffs.size());                            // the predicate is the caller's: evaluated here, serially (:1069-1074)
    for (std::size_t l = 0; l < coeffs.size(); ++l)
        keep[l] = (pred(Eigen::Vector3d(pos[3 * l], pos[3 * l + 1], pos[3 * l + 2]), coeffs[l]) && coeffs[l] != dbl_max) ? 1 : 0;
    std::vector<double>().swap(pos);
    auto& cell_map = m_cell_map[field_id];
    cell_map.resize(m_n_cells);
    // surviving cells, surviving nodes in Z-curve order, renumbered connectivity, cell map (:1076-1173) -- the library's index passes
    std::uint64_t n_nodes = 0, n_cells = 0;
    check(dg_reduce_field(&d, coeffs.data(), coeffs.size(), keep.data(), reinterpret_cast<std::uint32_t*>(cells.data()), cells.size(),
                          cell_map.data(), 0u, &n_nodes, &n_cells, nullptr));
    coeffs.resize(n_nodes);
    cells.resize(n_cells);
}

}  // namespace Discregrid
