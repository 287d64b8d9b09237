// Discregrid::TriangleMeshDistance with the reference's public surface (geometry/TriangleMeshDistance.h:75-208),
// implemented on the B200 through the C-ABI (include/discregrid_b200.h).  construct() builds the reference's
// sphere tree + pseudonormals on the host and uploads them once (dg_mesh_create); signed_distance / unsigned_distance run
// in the sm_100a traversal kernel -- one point per call for source compatibility, or batched with the *_batch overloads
// (the form to use: a launch per point is dominated by latency).
// Differences, all deliberate: unconstructed / empty mesh throws std::runtime_error instead of exit(-1)
// (:318-321, :338-341); the raw-pointer construct() sizes its arrays correctly (the reference over-allocates 3x and
// injects degenerate triangles, :232-249).
#pragma once
#include <array>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#include "../mesh/triangle_mesh.hpp"
#include "discregrid_b200.h"

namespace Discregrid {

enum class NearestEntity { V0, V1, V2, E01, E12, E02, F };

template <typename FLOAT>
class Vec3r {
public:
    std::array<FLOAT, 3> v;
    Vec3r() {}
    template <typename FLOAT_I> Vec3r(const FLOAT_I& x, const FLOAT_I& y, const FLOAT_I& z) { v[0] = static_cast<FLOAT>(x); v[1] = static_cast<FLOAT>(y); v[2] = static_cast<FLOAT>(z); }
    template <typename SIZE_T> const FLOAT& operator[](const SIZE_T& i) const { return v[i]; }
    template <typename SIZE_T> FLOAT& operator[](const SIZE_T& i) { return v[i]; }
};
using Vec3d = Vec3r<double>;

struct Result {
    double distance = std::numeric_limits<double>::max();
    Vec3d nearest_point;
    Discregrid::NearestEntity nearest_entity;
    int triangle_id = -1;
};

class TriangleMeshDistance {
public:
    TriangleMeshDistance() = default;
    TriangleMeshDistance(const TriangleMeshDistance&) = delete;
    TriangleMeshDistance& operator=(const TriangleMeshDistance&) = delete;
    ~TriangleMeshDistance() { if (m_group) dg_mesh_group_destroy(m_group); if (m_mesh) dg_mesh_destroy(m_mesh); }

    template <typename FLOAT, typename INT, typename SIZE_T>
    TriangleMeshDistance(const FLOAT* vertices, const SIZE_T n_vertices, const INT* triangles, const SIZE_T n_triangles) { construct(vertices, n_vertices, triangles, n_triangles); }
    template <typename IndexableVector3double, typename IndexableVector3int>
    TriangleMeshDistance(const std::vector<IndexableVector3double>& vertices, const std::vector<IndexableVector3int>& triangles) { construct(vertices, triangles); }
    TriangleMeshDistance(const TriangleMesh& mesh) { construct(mesh.vertex_data(), mesh.face_data()); }

    template <typename FLOAT, typename INT, typename SIZE_T>
    void construct(const FLOAT* vertices, const SIZE_T n_vertices, const INT* triangles, const SIZE_T n_triangles)
    {
        std::vector<double> V(3 * (size_t)n_vertices); std::vector<uint32_t> F(3 * (size_t)n_triangles);
        for (size_t i = 0; i < V.size(); i++) V[i] = (double)vertices[i];
        for (size_t i = 0; i < F.size(); i++) F[i] = (uint32_t)triangles[i];
        upload(V, F);
    }
    template <typename IndexableVector3double, typename IndexableVector3int>
    void construct(const std::vector<IndexableVector3double>& vertices, const std::vector<IndexableVector3int>& triangles)
    {
        std::vector<double> V(3 * vertices.size()); std::vector<uint32_t> F(3 * triangles.size());
        for (size_t i = 0; i < vertices.size(); i++) for (int d = 0; d < 3; d++) V[3 * i + d] = (double)vertices[i][d];
        for (size_t i = 0; i < triangles.size(); i++) for (int d = 0; d < 3; d++) F[3 * i + d] = (uint32_t)triangles[i][d];
        upload(V, F);
    }

    template <typename IndexableVector3double> Result unsigned_distance(const IndexableVector3double& p) const { return unsigned_distance(std::array<double, 3>{{(double)p[0], (double)p[1], (double)p[2]}}); }
    Result unsigned_distance(const std::array<double, 3>& p) const { Result r; query(p.data(), 1, 0, &r); return r; }
    template <typename IndexableVector3double> Result signed_distance(const IndexableVector3double& p) const { return signed_distance(std::array<double, 3>{{(double)p[0], (double)p[1], (double)p[2]}}); }
    Result signed_distance(const std::array<double, 3>& p) const { Result r; query(p.data(), 1, 1, &r); return r; }

    // batched forms (new): points = n x 3 doubles, xyzxyz
    std::vector<Result> signed_distance_batch(const double* points, size_t n) const { std::vector<Result> r(n); query(points, n, 1, r.data()); return r; }
    std::vector<Result> unsigned_distance_batch(const double* points, size_t n) const { std::vector<Result> r(n); query(points, n, 0, r.data()); return r; }

    bool is_constructed() const { return m_mesh != nullptr; }
    const dg_mesh* handle() const { require(); return m_mesh; }
    // new: spread addFunction over n GPUs of this node (`GenerateSDF --gpus N`): the device records are replicated once, by peer copies
    void useGpus(int n_gpus)
    {
        require();
        if (m_group) { dg_mesh_group_destroy(m_group); m_group = nullptr; }
        if (n_gpus > 1 && dg_mesh_group_create(m_mesh, n_gpus, nullptr, &m_group) != DG_OK) throw std::runtime_error(std::string("TriangleMeshDistance: ") + dg_last_error());
    }
    dg_mesh_group* group() const { return m_group; }

private:
    dg_mesh* m_mesh = nullptr;
    dg_mesh_group* m_group = nullptr;
    void require() const { if (!m_mesh) throw std::runtime_error("DistanceTriangleMesh error: not constructed."); }
    void upload(const std::vector<double>& V, const std::vector<uint32_t>& F)
    {
        if (m_group) { dg_mesh_group_destroy(m_group); m_group = nullptr; }
        if (m_mesh) { dg_mesh_destroy(m_mesh); m_mesh = nullptr; }
        if (F.empty()) throw std::runtime_error("DistanceTriangleMesh error: Empty triangle list.");
        if (dg_mesh_create(V.data(), V.size() / 3, F.data(), F.size() / 3, &m_mesh) != DG_OK) throw std::runtime_error(std::string("TriangleMeshDistance: ") + dg_last_error());
    }
    void query(const double* pts, size_t n, int is_signed, Result* out) const
    {
        require();
        std::vector<double> d(n), q(3 * n); std::vector<int32_t> e(n), t(n);
        if (dg_mesh_distance(m_mesh, pts, n, is_signed, d.data(), q.data(), e.data(), t.data()) != DG_OK) throw std::runtime_error(std::string("TriangleMeshDistance: ") + dg_last_error());
        for (size_t i = 0; i < n; i++) { out[i].distance = d[i]; out[i].nearest_point = Vec3d(q[3 * i], q[3 * i + 1], q[3 * i + 2]); out[i].nearest_entity = (NearestEntity)e[i]; out[i].triangle_id = t[i]; }
    }
};

// The functor GenerateSDF passes to addFunction (cmd/generate_sdf/main.cpp:94-102), as a TYPE the grid can recognise through
// std::function::target<>() (the reference uses an anonymous lambda, which a GPU cannot run -- SURVEY F6/H1).  Calling it
// evaluates one point through the batch API, like the reference lambda.
struct MeshSignedDistanceFunction {
    const TriangleMeshDistance* md;
    double sign;                                    // -1.0 for --invert
    explicit MeshSignedDistanceFunction(const TriangleMeshDistance& m, bool invert = false) : md(&m), sign(invert ? -1.0 : 1.0) {}
    double operator()(Eigen::Vector3d const& xi) const { const double d = md->signed_distance(xi).distance; return sign == 1.0 ? d : sign * d; }
};

}  // namespace Discregrid
