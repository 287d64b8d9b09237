// DiscreteFieldToBitmap -- the reference's only batched-interpolate consumer (cmd/discrete_field_to_bitmap/main.cpp:106-177),
// on the batch API: the pixel loop `#pragma omp parallel for ... interpolate(field_id, sample)` (:118-140) becomes ONE
// dg_interpolate_batch call.  Same options, same sample positions, same normalisation and colour maps; 24-bit BMP output.
//   DiscreteFieldToBitmap [-f field_id] [-s samples] [-p xy|xz|yz|...] [-d depth] [-c gb|rs] [-o out.bmp] file.cdf|.cdm
#include <Discregrid/All>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

using namespace Eigen;

static void put_u32(std::vector<unsigned char>& v, std::uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((unsigned char)(x >> (8 * i))); }
static void put_u16(std::vector<unsigned char>& v, std::uint16_t x) { v.push_back((unsigned char)x); v.push_back((unsigned char)(x >> 8)); }

static bool write_bmp(const std::string& path, unsigned w, unsigned h, const std::vector<unsigned char>& rgb)
{
    const unsigned line = ((w * 3u + 3u) >> 2) << 2;                        // rows padded to 4 bytes, stored in the given row order
    std::vector<unsigned char> f;
    f.push_back('B'); f.push_back('M'); put_u32(f, 54u + line * h); put_u16(f, 0); put_u16(f, 0); put_u32(f, 54u);
    put_u32(f, 40u); put_u32(f, w); put_u32(f, h); put_u16(f, 1); put_u16(f, 24); put_u32(f, 0); put_u32(f, line * h);
    put_u32(f, 4000); put_u32(f, 4000); put_u32(f, 0); put_u32(f, 0);
    for (unsigned j = 0; j < h; j++) {
        for (unsigned i = 0; i < w; i++) { const unsigned char* p = &rgb[3 * (j * w + i)]; f.push_back(p[2]); f.push_back(p[1]); f.push_back(p[0]); }
        for (unsigned k = 3 * w; k < line; k++) f.push_back(0);
    }
    FILE* fp = std::fopen(path.c_str(), "wb");
    if (!fp) return false;
    const bool ok = std::fwrite(f.data(), 1, f.size(), fp) == f.size();
    std::fclose(fp);
    return ok;
}

int main(int argc, char* argv[])
{
    unsigned field_id = 0, xsamples = 1024;
    std::string plane = "xy", out_file, cm = "gb", filename;
    double depth = 0.0;
    for (int a = 1; a < argc; a++) {
        const std::string s = argv[a];
        auto next = [&]() -> std::string { if (a + 1 >= argc) { std::cerr << "missing value for " << s << std::endl; std::exit(1); } return argv[++a]; };
        if (s == "-f" || s == "--field_id") field_id = (unsigned)std::stoul(next());
        else if (s == "-s" || s == "--samples") xsamples = (unsigned)std::stoul(next());
        else if (s == "-p" || s == "--plane") plane = next();
        else if (s == "-d" || s == "--depth") depth = std::stod(next());
        else if (s == "-o" || s == "--output") out_file = next();
        else if (s == "-c" || s == "--colormap") cm = next();
        else if (s == "-h" || s == "--help") { std::cout << "Transforms a slice of a discrete SDF to a bitmap image.\nExample: DiscreteFieldToBitmap -p xz file.cdf" << std::endl; return 0; }
        else filename = s;
    }
    if (filename.empty()) { std::cout << "ERROR: No input file given." << std::endl; return 1; }
    try {
        std::cout << "Load SDF...";
        Discregrid::CubicLagrangeDiscreteGrid sdf(filename);
        std::cout << "DONE" << std::endl;
        if (sdf.nFields() <= field_id) { std::cerr << "ERROR: field " << field_id << " does not exist" << std::endl; return 1; }
        auto const& domain = sdf.domain();
        const Vector3d diag = domain.diagonal();
        if (plane.length() != 2 || plane[0] == plane[1]) { std::cerr << "ERROR: Invalid option for plane provided. Should be one of the following options: xy, xz, yz, yx" << std::endl; return 1; }
        int dir[3] = {0, 0, 0};                                             // main.cpp:87-100
        if (plane[0] == 'y') dir[0] = 1; else if (plane[0] == 'z') dir[0] = 2;
        if (plane[1] == 'y') dir[1] = 1; else if (plane[1] == 'z') dir[1] = 2;
        if (dir[0] != 1 && dir[1] != 1) dir[2] = 1;
        if (dir[0] != 2 && dir[1] != 2) dir[2] = 2;
        const unsigned ysamples = (unsigned)std::round(diag[dir[1]] / diag[dir[0]] * (double)xsamples);
        const double xwidth = diag[dir[0]] / xsamples, ywidth = diag[dir[1]] / ysamples;
        const std::size_t n = (std::size_t)xsamples * ysamples;
        std::vector<double> x(3 * n), data(n);
        std::cout << "Sample field...";
        for (std::size_t k = 0; k < n; k++) {                               // sample positions of main.cpp:120-134
            const unsigned i = (unsigned)(k % xsamples), j = (unsigned)(k / xsamples);
            const double xr = (double)i / (double)xsamples, yr = (double)j / (double)ysamples;
            x[3 * k + dir[0]] = domain.min()[dir[0]] + xr * diag[dir[0]] + 0.5 * xwidth;
            x[3 * k + dir[1]] = domain.min()[dir[1]] + yr * diag[dir[1]] + 0.5 * ywidth;
            x[3 * k + dir[2]] = domain.min()[dir[2]] + 0.5 * (1.0 + depth) * diag[dir[2]];
        }
        sdf.interpolate(field_id, n, x.data(), data.data(), nullptr);      // one batched launch
        for (auto& v : data) if (v == std::numeric_limits<double>::max()) v = 0.0;       // :136-139
        std::cout << "DONE" << std::endl;
        const double min_v = *std::min_element(data.begin(), data.end()), max_v = *std::max_element(data.begin(), data.end());
        if (out_file.empty()) {
            out_file = filename;
            if (out_file.find(".") != std::string::npos) out_file = out_file.substr(0, out_file.find_last_of("."));
            out_file += ".bmp";
        }
        std::cout << "Ouput file: " << out_file << std::endl << "Export BMP...";
        if (cm != "gb" && cm != "rs") { std::cerr << "WARNING: Unknown color map option. Fallback to mode 'gb'." << std::endl; cm = "gb"; }
        std::vector<unsigned char> rgb(3 * n, 0);
        auto clamp255 = [](double v) { return (unsigned char)std::min(std::max(v, 0.0), 255.0); };
        for (std::size_t k = 0; k < n; k++) {
            const double v = data[k] >= 0.0 ? data[k] / std::abs(max_v) : data[k] / std::abs(min_v);      // :159
            if (cm == "rs") rgb[3 * k] = clamp255(255.0 * v);                                            // red sequential (:25-28)
            else if (v >= 0.0) rgb[3 * k + 1] = clamp255(255.0 * (1.0 - v));                              // green / blue inverse diverging (:15-23)
            else rgb[3 * k + 2] = clamp255(255.0 * (1.0 + v));
        }
        if (!write_bmp(out_file, xsamples, ysamples, rgb)) { std::cerr << "ERROR: cannot write " << out_file << std::endl; return 1; }
        std::cout << "DONE" << std::endl << std::endl << "Statistics:" << std::endl;
        std::cout << "\tmin value      = " << min_v << std::endl << "\tmax value      = " << max_v << std::endl;
        std::cout << "\tbmp resolution = " << xsamples << " x " << ysamples << std::endl;
    } catch (std::exception const& e) { std::cerr << "error: " << e.what() << std::endl; return 1; }
    return 0;
}
