// Discregrid::TriangleMesh -- the subset of the reference mesh class the hot path consumes
// (reference: discregrid/include/Discregrid/mesh/triangle_mesh.hpp, discregrid/src/mesh/triangle_mesh.cpp:90-147):
// OBJ reader (`v`/`f` records, `a/b/c` tolerated, 1-based, triangles), vertex/face accessors, exportOBJ.
// The half-edge adjacency of the reference class is not needed by TriangleMeshDistance (TriangleMeshDistance.h:227-230).
#pragma once
#include <array>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>
#include <Eigen/Dense>
#include "discregrid_b200.h"

namespace Discregrid {
class TriangleMesh {
public:
    using Face = std::array<unsigned int, 3>;
    TriangleMesh() = default;
    TriangleMesh(std::vector<Eigen::Vector3d> const& vertices, std::vector<Face> const& faces) : m_vertices(vertices), m_faces(faces) {}
    TriangleMesh(double const* vertices, unsigned int const* faces, std::size_t nv, std::size_t nf)
    {
        m_vertices.resize(nv); m_faces.resize(nf);
        for (std::size_t i = 0; i < nv; i++) m_vertices[i] = Eigen::Vector3d(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
        for (std::size_t i = 0; i < nf; i++) m_faces[i] = {{faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]}};
    }
    // OBJ file: the library's reader (dg_obj_read) -- same records, same values as the reference's stream parser (triangle_mesh.cpp:90-124)
    explicit TriangleMesh(std::string const& path)
    {
        double* v = nullptr; std::uint32_t* f = nullptr;
        std::uint64_t nv = 0, nf = 0;
        const int rc = dg_obj_read(path.c_str(), &v, &nv, &f, &nf);
        if (rc == DG_ERR_IO) { std::cerr << "Cannot open " << path << std::endl; return; }              // :93-97: message, object left empty
        if (rc != DG_OK) throw std::invalid_argument(std::string("stoi: ") + dg_last_error());            // what std::stoi does at :113
        m_vertices.resize(nv); m_faces.resize(nf);
        for (std::size_t i = 0; i < nv; i++) m_vertices[i] = Eigen::Vector3d(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        for (std::size_t i = 0; i < nf; i++) m_faces[i] = {{f[3 * i], f[3 * i + 1], f[3 * i + 2]}};
        dg_obj_free(v, f);
    }
    void exportOBJ(std::string const& filename) const
    {
        std::ofstream out(filename.c_str());
        out << "g default" << std::endl;
        for (auto const& p : m_vertices) out << "v " << p[0] << " " << p[1] << " " << p[2] << "\n";
        for (auto const& f : m_faces) out << "f " << f[0] + 1 << " " << f[1] + 1 << " " << f[2] + 1 << std::endl;
    }
    std::vector<Eigen::Vector3d> const& vertices() const { return m_vertices; }
    std::vector<Eigen::Vector3d> const& vertex_data() const { return m_vertices; }
    std::vector<Eigen::Vector3d>& vertex_data() { return m_vertices; }
    std::vector<Face> const& faces() const { return m_faces; }
    std::vector<Face> const& face_data() const { return m_faces; }
    std::vector<Face>& face_data() { return m_faces; }
    Eigen::Vector3d const& vertex(unsigned int i) const { return m_vertices[i]; }
    std::size_t nFaces() const { return m_faces.size(); }
    std::size_t nVertices() const { return m_vertices.size(); }
private:
    std::vector<Eigen::Vector3d> m_vertices;
    std::vector<Face> m_faces;
};
}  // namespace Discregrid
