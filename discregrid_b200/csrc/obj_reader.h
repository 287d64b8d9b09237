// Wavefront OBJ reader for the input side of the path: what Discregrid::TriangleMesh(path) accepts
// (src/mesh/triangle_mesh.cpp:90-124) -- "v x y z" and "f a[/..] b[/..] c[/..]" lines, everything else ignored -- parsed from one
// in-memory copy of the file by all host threads (std::from_chars, correctly rounded like the stream extraction it replaces).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dgb {

struct ObjData {
    std::vector<double> vertices;     // nV x 3
    std::vector<uint32_t> faces;      // nT x 3, zero-based
};

// false + message on an unreadable file or a face line the reference's std::stoi would throw on
bool read_obj(const char* path, ObjData& out, std::string& err);

}  // namespace dgb
