// CPU-only check of the facade's reduceField glue (cpp/include/Discregrid/cubic_lagrange_discrete_grid.hpp): loads a two-field
// .cdm, applies GenerateDensityMap's two reductions (cmd/generate_density_map/main.cpp:137-144) through the facade class and saves.
// The facade asks the library for node positions (a GPU kernel); this TEST binary interposes that one entry point with the oracle's
// indexToNodePosition so the check runs without a GPU.  Test infrastructure only.
//   reduce_facade_check in.cdm h rho0 out.cdm
#include <Discregrid/All>
#include <cstdlib>
#include <iostream>

extern "C" void orc_node_positions(const double* gd, const uint32_t* res, uint64_t l_begin, uint64_t l_end, double* x);

extern "C" int dg_node_positions(const dg_grid_desc* g, uint64_t l_begin, uint64_t l_end, double* x_host)
{
    double gd[12];
    for (int k = 0; k < 3; k++) { gd[k] = g->domain_min[k]; gd[3 + k] = g->domain_max[k]; gd[6 + k] = g->cell_size[k]; gd[9 + k] = g->inv_cell_size[k]; }
    orc_node_positions(gd, g->resolution, l_begin, l_end, x_host);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 5) { std::cerr << "usage: reduce_facade_check in.cdm h rho0 out.cdm" << std::endl; return 2; }
    const double h = std::atof(argv[2]), rho0 = std::atof(argv[3]);
    Discregrid::CubicLagrangeDiscreteGrid grid{std::string(argv[1])};
    if (grid.nFields() != 2) { std::cerr << "expected two fields" << std::endl; return 1; }
    const double cell_diag = grid.cellSize().norm();
    long n_pos_checked = 0;
    grid.reduceField(0u, [&](const Eigen::Vector3d& x, double v) { if (grid.domain().contains(x)) n_pos_checked++; return -6.0 * h < v + cell_diag && v - cell_diag < 2.0 * h; });
    grid.reduceField(1u, [&](const Eigen::Vector3d&, double v) { return 0.0 <= v && v <= 3.0 * rho0; });
    grid.save(argv[4]);
    std::cout << "nodes " << grid.nodeData(0).size() << " " << grid.nodeData(1).size() << " cells " << grid.cellData(0).size() << " " << grid.cellData(1).size()
              << " positions_in_domain " << n_pos_checked << std::endl;
    return 0;
}
