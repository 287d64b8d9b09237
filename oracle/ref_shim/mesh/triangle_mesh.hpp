// oracle/ref_shim/mesh/triangle_mesh.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's Eigen-dependent discregrid/include/Discregrid/mesh/triangle_mesh.hpp
// (Eigen3 is not installed in this image).  The reference's TriangleMeshDistance.h only needs
// vertex_data() / face_data() (TriangleMeshDistance.h:227-230, 251-266).  oracle/Makefile copies this
// file next to a build-time copy of the UNMODIFIED reference header under oracle/_ref/src/.
#pragma once
#include <array>
#include <vector>
namespace Discregrid {
class TriangleMesh {
public:
    std::vector<std::array<double, 3>> m_v;
    std::vector<std::array<unsigned int, 3>> m_f;
    std::vector<std::array<double, 3>> const& vertex_data() const { return m_v; }
    std::vector<std::array<unsigned int, 3>> const& face_data() const { return m_f; }
};
}
