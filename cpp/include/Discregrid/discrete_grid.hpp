// Discregrid::DiscreteGrid for the B200-native library.
//
// Interface-compatible with the reference's abstract base class (discregrid/include/Discregrid/discrete_grid.hpp:11-99 and
// discregrid/src/discrete_grid.cpp:8-38): same public type aliases, the same pure-virtual methods with the same parameter lists
// and defaults, the same protected members -- that is what "drop-in" means for code written against it.  Everything here is
// header-only; the regular-grid index algebra is written out per component instead of through Eigen expression templates.
#pragma once
#include <array>
#include <cstddef>
#include <fstream>
#include <functional>
#include <string>
#include <vector>
#include <Eigen/Dense>

namespace Discregrid {

class DiscreteGrid {
public:
    // ---- the aliases user code spells out (discrete_grid.hpp:15-19)
    typedef std::array<unsigned int, 3> MultiIndex;
    typedef Eigen::Matrix<double, 32, 1> CoefficientVector;
    typedef std::function<double(Eigen::Vector3d const&)> ContinuousFunction;
    typedef std::function<bool(Eigen::Vector3d const&)> SamplePredicate;
    typedef std::function<bool(Eigen::Vector3d const&, double)> Predicate;

    DiscreteGrid() = default;
    virtual ~DiscreteGrid() = default;

    // cell size = diagonal / resolution, inverse = 1 / cell size, cell count as an unsigned product (discrete_grid.hpp:22-29)
    DiscreteGrid(Eigen::AlignedBox3d const& box, std::array<unsigned int, 3> const& res) : m_domain(box), m_resolution(res), m_n_fields(0u)
    {
        for (int axis = 0; axis < 3; ++axis) {
            m_cell_size[axis] = (box.max()[axis] - box.min()[axis]) / static_cast<double>(res[axis]);
            m_inv_cell_size[axis] = 1.0 / m_cell_size[axis];
        }
        const unsigned int n_cells = res[0] * res[1] * res[2];
        m_n_cells = n_cells;
    }

    // ---- geometry accessors
    Eigen::AlignedBox3d const& domain() const { return m_domain; }
    std::array<unsigned int, 3> const& resolution() const { return m_resolution; }
    Eigen::Vector3d const& cellSize() const { return m_cell_size; }
    Eigen::Vector3d const& invCellSize() const { return m_inv_cell_size; }

    // ---- cell index algebra: cell (i, j, k) <-> i + nx * (j + ny * k)
    unsigned int multiToSingleIndex(MultiIndex const& c) const { return m_resolution[1] * m_resolution[0] * c[2] + m_resolution[0] * c[1] + c[0]; }
    MultiIndex singleToMultiIndex(unsigned int cell) const
    {
        const unsigned int per_layer = m_resolution[0] * m_resolution[1];
        const unsigned int k = cell / per_layer, in_layer = cell % per_layer;
        MultiIndex c = {{in_layer % m_resolution[0], in_layer / m_resolution[0], k}};
        return c;
    }
    Eigen::AlignedBox3d subdomain(MultiIndex const& c) const
    {
        Eigen::Vector3d lo, hi;
        for (int axis = 0; axis < 3; ++axis) {
            lo[axis] = m_domain.min()[axis] + static_cast<double>(c[axis]) * m_cell_size[axis];
            hi[axis] = lo[axis] + m_cell_size[axis];
        }
        return Eigen::AlignedBox3d(lo, hi);
    }
    Eigen::AlignedBox3d subdomain(unsigned int cell) const { return subdomain(singleToMultiIndex(cell)); }

    // ---- persistence and sampling, implemented by CubicLagrangeDiscreteGrid
    virtual void load(std::string const& filename) = 0;
    virtual void save(std::string const& filename) const = 0;
    virtual unsigned int addFunction(ContinuousFunction const& func, bool verbose = false, SamplePredicate const& pred = nullptr) = 0;
    virtual void reduceField(unsigned int /*field_id*/, Predicate /*pred*/) {}

    // ---- evaluation: value (and gradient) of field `field_id` at x; the one-argument form reads field 0
    virtual double interpolate(unsigned int field_id, Eigen::Vector3d const& x, Eigen::Vector3d* grad = nullptr) const = 0;
    double interpolate(Eigen::Vector3d const& x, Eigen::Vector3d* grad = nullptr) const { return interpolate(0u, x, grad); }

    // split form: cell, reference-cell scale c0 and shape functions N (and Jacobian dN) once, then any number of fields
    virtual bool determineShapeFunctions(unsigned int field_id, Eigen::Vector3d const& x, std::array<unsigned int, 32>& cell, Eigen::Vector3d& c0,
                                         Eigen::Matrix<double, 32, 1>& N, Eigen::Matrix<double, 32, 3>* dN = nullptr) const = 0;
    virtual double interpolate(unsigned int field_id, Eigen::Vector3d const& x, const std::array<unsigned int, 32>& cell, const Eigen::Vector3d& c0,
                               const Eigen::Matrix<double, 32, 1>& N, Eigen::Vector3d* grad = nullptr, Eigen::Matrix<double, 32, 3>* dN = nullptr) const = 0;

protected:
    // same members, same order and types as the reference: save()/load() write them verbatim
    Eigen::AlignedBox3d m_domain;
    std::array<unsigned int, 3> m_resolution;
    Eigen::Vector3d m_cell_size;
    Eigen::Vector3d m_inv_cell_size;
    std::size_t m_n_cells;
    std::size_t m_n_fields;
};

}  // namespace Discregrid
