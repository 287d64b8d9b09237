// Test helper (no GPU): runs the PRODUCT's host-side builder (discregrid_b200/csrc/bvh_build.cpp) on an OBJ-less raw mesh and
// dumps the tree in the reference's numbering, the pseudonormals and the device records, for tests/test_host_bvh.py.
//   bvh_host_check <V.bin (nV x 3 f64)> <F.bin (nT x 3 u32)> <out.bin>
#include "../../discregrid_b200/csrc/bvh_build.h"
#include <cstdio>
#include <vector>
template <class T> static std::vector<T> slurp(const char* p) { std::vector<T> v; FILE* f = std::fopen(p, "rb"); if (!f) return v; std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET); v.resize(n / sizeof(T)); if (std::fread(v.data(), sizeof(T), v.size(), f) != v.size()) v.clear(); std::fclose(f); return v; }
int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    const auto V = slurp<double>(argv[1]); const auto F = slurp<uint32_t>(argv[2]);
    dgb::HostBvh h; const char* err = "";
    if (!dgb::build_host_bvh(V.data(), V.size() / 3, F.data(), F.size() / 3, h, &err)) { std::fprintf(stderr, "build failed: %s\n", err); return 1; }
    const size_t nT = F.size() / 3, nn = 2 * nT - 1;
    std::vector<double> sph(8 * nn); std::vector<int32_t> kids(2 * nn);
    dgb::export_reference_tree(h, sph.data(), kids.data());
    FILE* f = std::fopen(argv[3], "wb");
    const double hdr[4] = {(double)nn, (double)h.max_depth, (double)h.flags, h.half_extent};
    std::fwrite(hdr, 8, 4, f);
    std::fwrite(sph.data(), 8, sph.size(), f); std::fwrite(kids.data(), 4, kids.size(), f);
    std::fwrite(h.pn_tri.data(), 8, h.pn_tri.size(), f); std::fwrite(h.pn_edge.data(), 8, h.pn_edge.size(), f); std::fwrite(h.pn_vert.data(), 8, h.pn_vert.size(), f);
    std::fwrite(h.leaves.data(), sizeof(dgb::LeafRecord), nT, f);
    std::fwrite(h.spheres_f.data(), sizeof(dgb::SpherePairF), nT, f);
    std::fwrite(h.boxes_f.data(), sizeof(dgb::BoxPairF), nT, f);
    std::fwrite(h.center, 8, 3, f);
    std::fclose(f);
    return 0;
}
