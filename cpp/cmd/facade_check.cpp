// facade_check -- exercises the C++ facade exactly as a downstream caller of the reference would (scalar interpolate,
// determineShapeFunctions + split interpolate, batched interpolate, TriangleMeshDistance queries, addFunction) and dumps the
// results as raw doubles for tests/test_gpu_cpp_facade.py to compare with the oracle.
//   facade_check <grid.cdf> <mesh.obj> <points.bin (n x 3 doubles)> <out.bin>
#include <Discregrid/All>
#include <cstdio>
#include <iostream>
#include <vector>

using namespace Eigen;

int main(int argc, char** argv)
{
    if (argc < 5) { std::cerr << "usage: facade_check grid.cdf mesh.obj points.bin out.bin" << std::endl; return 2; }
    try {
        Discregrid::CubicLagrangeDiscreteGrid grid(argv[1]);
        std::vector<double> pts;
        { FILE* f = std::fopen(argv[3], "rb"); if (!f) return 2; double b[3]; while (std::fread(b, 8, 3, f) == 3) pts.insert(pts.end(), b, b + 3); std::fclose(f); }
        const std::size_t n = pts.size() / 3;
        std::vector<double> out;
        // 1. scalar interpolate with gradient (gradient pre-set to a marker: untouched outside the domain, as in the reference)
        for (std::size_t q = 0; q < n; q++) {
            Vector3d x(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]), g(7.0, 7.0, 7.0);
            const double phi = grid.interpolate(0u, x, &g);
            out.push_back(phi); out.push_back(g[0]); out.push_back(g[1]); out.push_back(g[2]);
        }
        // 2. split API
        for (std::size_t q = 0; q < n; q++) {
            Vector3d x(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]), c0, g(7.0, 7.0, 7.0);
            std::array<unsigned int, 32> cell; Matrix<double, 32, 1> N; Matrix<double, 32, 3> dN;
            double phi = -1.0;
            const bool ok = grid.determineShapeFunctions(0u, x, cell, c0, N, &dN);
            if (ok) phi = grid.interpolate(0u, x, cell, c0, N, &g, &dN);
            out.push_back(ok ? 1.0 : 0.0); out.push_back(phi); out.push_back(g[0]); out.push_back(g[1]); out.push_back(g[2]);
        }
        // 3. batched, value only
        { std::vector<double> phi(n); grid.interpolate(0u, n, pts.data(), phi.data(), nullptr); out.insert(out.end(), phi.begin(), phi.end()); }
        // 4. TriangleMeshDistance: scalar queries on the first 64 points, batch on all
        Discregrid::TriangleMesh mesh(argv[2]);
        Discregrid::TriangleMeshDistance md(mesh);
        for (std::size_t q = 0; q < n && q < 64; q++) {
            const Discregrid::Result r = md.signed_distance(Vector3d(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]));
            out.push_back(r.distance); out.push_back(r.nearest_point[0]); out.push_back(r.nearest_point[1]); out.push_back(r.nearest_point[2]);
            out.push_back((double)(int)r.nearest_entity); out.push_back((double)r.triangle_id);
        }
        for (auto const& r : md.unsigned_distance_batch(pts.data(), n)) out.push_back(r.distance);
        // 5. addFunction with the inverted functor on a fresh grid with the same domain
        Discregrid::CubicLagrangeDiscreteGrid g2(grid.domain(), {{3, 4, 2}});
        Discregrid::DiscreteGrid::ContinuousFunction f = Discregrid::MeshSignedDistanceFunction(md, true);
        const unsigned id = g2.addFunction(f);
        out.push_back((double)id);
        out.insert(out.end(), g2.nodeData(0).begin(), g2.nodeData(0).end());
        bool threw = false;
        try { g2.addFunction([](Vector3d const&) { return 0.0; }); } catch (std::invalid_argument const&) { threw = true; }
        out.push_back(threw ? 1.0 : 0.0);
        FILE* f2 = std::fopen(argv[4], "wb"); std::fwrite(out.data(), 8, out.size(), f2); std::fclose(f2);
        std::cout << "facade_check wrote " << out.size() << " doubles" << std::endl;
    } catch (std::exception const& e) { std::cerr << "error: " << e.what() << std::endl; return 1; }
    return 0;
}
