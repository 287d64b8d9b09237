// Discregrid::DiscreteGrid -- same abstract interface, typedefs and members as the reference
// (discregrid/include/Discregrid/discrete_grid.hpp:11-99, discregrid/src/discrete_grid.cpp:8-38).
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
    using CoefficientVector = Eigen::Matrix<double, 32, 1>;
    using ContinuousFunction = std::function<double(Eigen::Vector3d const&)>;
    using MultiIndex = std::array<unsigned int, 3>;
    using Predicate = std::function<bool(Eigen::Vector3d const&, double)>;
    using SamplePredicate = std::function<bool(Eigen::Vector3d const&)>;

    DiscreteGrid() = default;
    DiscreteGrid(Eigen::AlignedBox3d const& domain, std::array<unsigned int, 3> const& resolution)
        : m_domain(domain), m_resolution(resolution), m_n_fields(0u)
    {
        for (int d = 0; d < 3; d++) {                                    // discrete_grid.hpp:25-27
            m_cell_size[d] = (domain.max()[d] - domain.min()[d]) / static_cast<double>(resolution[d]);
            m_inv_cell_size[d] = 1.0 / m_cell_size[d];
        }
        m_n_cells = resolution[0] * resolution[1] * resolution[2];       // unsigned product, as n.prod()
    }
    virtual ~DiscreteGrid() = default;

    virtual void save(std::string const& filename) const = 0;
    virtual void load(std::string const& filename) = 0;
    virtual unsigned int addFunction(ContinuousFunction const& func, bool verbose = false, SamplePredicate const& pred = nullptr) = 0;

    double interpolate(Eigen::Vector3d const& xi, Eigen::Vector3d* gradient = nullptr) const { return interpolate(0u, xi, gradient); }
    virtual double interpolate(unsigned int field_id, Eigen::Vector3d const& xi, Eigen::Vector3d* gradient = nullptr) const = 0;
    virtual bool determineShapeFunctions(unsigned int field_id, Eigen::Vector3d const& x, std::array<unsigned int, 32>& cell, Eigen::Vector3d& c0,
                                         Eigen::Matrix<double, 32, 1>& N, Eigen::Matrix<double, 32, 3>* dN = nullptr) const = 0;
    virtual double interpolate(unsigned int field_id, Eigen::Vector3d const& xi, const std::array<unsigned int, 32>& cell, const Eigen::Vector3d& c0,
                               const Eigen::Matrix<double, 32, 1>& N, Eigen::Vector3d* gradient = nullptr, Eigen::Matrix<double, 32, 3>* dN = nullptr) const = 0;
    virtual void reduceField(unsigned int, Predicate) {}

    MultiIndex singleToMultiIndex(unsigned int l) const
    {
        const unsigned n01 = m_resolution[0] * m_resolution[1];
        const unsigned k = l / n01, temp = l % n01;
        return {{temp % m_resolution[0], temp / m_resolution[0], k}};
    }
    unsigned int multiToSingleIndex(MultiIndex const& ijk) const { return m_resolution[1] * m_resolution[0] * ijk[2] + m_resolution[0] * ijk[1] + ijk[0]; }
    Eigen::AlignedBox3d subdomain(MultiIndex const& ijk) const
    {
        Eigen::Vector3d origin;
        for (int d = 0; d < 3; d++) origin[d] = m_domain.min()[d] + static_cast<double>(ijk[d]) * m_cell_size[d];
        Eigen::Vector3d top;
        for (int d = 0; d < 3; d++) top[d] = origin[d] + m_cell_size[d];
        return Eigen::AlignedBox3d(origin, top);
    }
    Eigen::AlignedBox3d subdomain(unsigned int l) const { return subdomain(singleToMultiIndex(l)); }

    Eigen::AlignedBox3d const& domain() const { return m_domain; }
    std::array<unsigned int, 3> const& resolution() const { return m_resolution; }
    Eigen::Vector3d const& cellSize() const { return m_cell_size; }
    Eigen::Vector3d const& invCellSize() const { return m_inv_cell_size; }

protected:
    Eigen::AlignedBox3d m_domain;
    std::array<unsigned int, 3> m_resolution;
    Eigen::Vector3d m_cell_size;
    Eigen::Vector3d m_inv_cell_size;
    std::size_t m_n_cells;
    std::size_t m_n_fields;
};
}  // namespace Discregrid
