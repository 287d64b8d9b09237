// =====================================================================================
// oracle/dg_oracle.cpp  --  TEST INFRASTRUCTURE ONLY.  NOT part of the product path.
//
// CPU restatement of the Discregrid hot path (reference commit ddf20dc), used ONLY by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
// the checker.  Nothing under discregrid_b200/ may include, link or call this file.
//
// Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py) against
//   * the reference's golden fixture cmd/generate_sdf/resources/box.cdf (1296 node
//     coefficients bit-exact, 125x32 connectivity, grid constants, domain padding);
//   * oracle/_ref/libdgref.so = the reference's own, unmodified TriangleMeshDistance.h
//     compiled by oracle/Makefile (tree, pseudonormals and signed distances bit-exact);
//   * the reference's own tools and grid class -- its UNMODIFIED discregrid/src/*.cpp and
//     cmd/*/*.cpp compiled by oracle/Makefile against the Eigen stand-in oracle/ref_eigen
//     (Eigen3 is not installed): GenerateSDF / GenerateDensityMap / DiscreteFieldToBitmap outputs
//     and CubicLagrangeDiscreteGrid::interpolate / determineShapeFunctions results are committed
//     as fixtures (tests/golden/make_golden.py) and reproduced bit for bit: interpolate incl.
//     reduced fields, shape_function_ + Jacobian, the density map on every node.
//     Caveat: the stand-in fixes Eigen's 3-term norm() order to (a0+a1)+a2 (Eigen >= 3.3); real
//     Eigen could not be run here.  It touches only cell_diag, W(r) and the GenerateSDF padding;
//   * mathematical identities of shape_function_ (nodal property, partition of unity,
//     finite-difference Jacobian).
//
// Language note: this is C++ rather than plain C for one reason only -- the reference
// BVH build (TriangleMeshDistance.h:494-499) orders triangles with the *unstable*
// std::sort, and triangles that share their first vertex tie in the sort key, so the tree
// shape depends on libstdc++'s introsort.  Re-using std::sort with the same comparator on
// the same sequence is the only faithful restatement.  Everything else is plain structs
// and loops.  All arithmetic is IEEE fp64 without contraction (-ffp-contract=off, no
// -march), as in the reference build (CMakeLists.txt:9-11 sets only C++11).
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/).
// =====================================================================================
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 { double x, y, z; };
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
// discregrid/include/Discregrid/geometry/TriangleMeshDistance.h:53  (left-to-right sum)
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// TriangleMeshDistance.h:54
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, -a.x * b.z + a.z * b.x, a.x * b.y - a.y * b.x}; }
static inline V3 scale(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }   // :59,70
static inline V3 divs(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }    // :61
static inline double norm(V3 a) { return std::sqrt(dot(a, a)); }                 // :64-65
static inline V3 normalized(V3 a) { return divs(a, norm(a)); }                   // :66
static inline double comp(const V3& a, int d) { return d == 0 ? a.x : (d == 1 ? a.y : a.z); }

struct Sphere { V3 c; double r; };
struct Node { Sphere bl, br; int left = -1, right = -1; };   // TriangleMeshDistance.h:103-109
struct BuildTri { V3 v[3]; int id; };                        // :111-115

enum Entity { V0 = 0, V1, V2, E01, E12, E02, F };            // :75

struct Result { double distance; V3 nearest; int entity; int tri; };  // :80-86

struct Mesh {
    std::vector<V3> V;
    std::vector<int> T;                 // 3 per triangle
    std::vector<Node> nodes;
    std::vector<V3> pn_tri;             // [nT]
    std::vector<V3> pn_edge;            // [nT][3]
    std::vector<V3> pn_vert;            // [nV]
    int flags = 0;                      // bit0: single edge found, bit1: >2 triangles per edge
    long long visits = 0, leaves = 0;   // instrumentation (single-threaded stats call only)
};

// TriangleMeshDistance.h:443-512
void build_tree(Mesh& m, int node_id, Sphere& bs, std::vector<BuildTri>& tris, int begin, int end)
{
    const int n = end - begin;
    if (n == 1) {                                                         // :451-462
        m.nodes[node_id].left = -1;
        m.nodes[node_id].right = tris[begin].id;
        const BuildTri& t = tris[begin];
        const V3 center = divs(add(add(t.v[0], t.v[1]), t.v[2]), 3.0);
        const double radius = std::max(std::max(norm(sub(t.v[0], center)), norm(sub(t.v[1], center))), norm(sub(t.v[2], center)));
        bs.c = center; bs.r = radius;
        return;
    }
    double top[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};                        // :465-478
    double bot[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
    V3 center = {0, 0, 0};
    for (int i = begin; i < end; i++)
        for (int k = 0; k < 3; k++) {
            const V3& p = tris[i].v[k];
            center = add(center, p);
            for (int d = 0; d < 3; d++) {
                top[d] = std::max(top[d], comp(p, d));
                bot[d] = std::min(bot[d], comp(p, d));
            }
        }
    center = divs(center, (double)(3 * n));                               // :479 (int -> double)
    const double diag[3] = {top[0] - bot[0], top[1] - bot[1], top[2] - bot[2]};
    const int split_dim = (int)(std::max_element(diag, diag + 3) - diag);  // :481 (first max)
    double radius_sq = 0.0;                                               // :484-491
    for (int i = begin; i < end; i++)
        for (int k = 0; k < 3; k++) {
            const V3 d = sub(center, tris[i].v[k]);
            radius_sq = std::max(radius_sq, dot(d, d));
        }
    bs.c = center; bs.r = std::sqrt(radius_sq);
    std::sort(tris.begin() + begin, tris.begin() + end,                    // :494-499
              [split_dim](const BuildTri& a, const BuildTri& b) { return comp(a.v[0], split_dim) < comp(b.v[0], split_dim); });
    const int mid = (int)(0.5 * (begin + end));                           // :502
    const int l = (int)m.nodes.size();                                    // :504-506
    m.nodes[node_id].left = l;
    m.nodes.push_back(Node());
    { Sphere s; build_tree(m, l, s, tris, begin, mid); m.nodes[node_id].bl = s; }
    const int r = (int)m.nodes.size();                                    // :508-510
    m.nodes[node_id].right = r;
    m.nodes.push_back(Node());
    { Sphere s; build_tree(m, r, s, tris, mid, end); m.nodes[node_id].br = s; }
}

// TriangleMeshDistance.h:336-441
void construct(Mesh& m)
{
    const int nT = (int)(m.T.size() / 3);
    std::vector<BuildTri> tris(nT);
    for (int i = 0; i < nT; i++) {
        tris[i].id = i;
        for (int k = 0; k < 3; k++) tris[i].v[k] = m.V[m.T[3 * i + k]];
    }
    m.nodes.clear();
    m.nodes.reserve(2 * (size_t)nT);
    m.nodes.push_back(Node());
    Sphere root;
    build_tree(m, 0, root, tris, 0, nT);

    // pseudonormals :359-420
    const uint64_t nV = (uint64_t)m.V.size();
    std::unordered_map<uint64_t, V3> edge_normals;
    std::unordered_map<uint64_t, int> edge_count;
    auto key_of = [&](int i, int j) { return (uint64_t)std::min(i, j) * nV + (uint64_t)std::max(i, j); };
    auto add_edge = [&](int i, int j, V3 n) {
        const uint64_t k = key_of(i, j);
        auto it = edge_normals.find(k);
        if (it == edge_normals.end()) { edge_normals[k] = n; edge_count[k] = 1; }
        else { it->second = add(it->second, n); edge_count[k] += 1; }
    };
    m.pn_tri.assign(nT, V3{0, 0, 0});
    m.pn_edge.assign(3 * (size_t)nT, V3{0, 0, 0});
    m.pn_vert.assign(m.V.size(), V3{0, 0, 0});
    for (int i = 0; i < nT; i++) {
        const int* t = &m.T[3 * i];
        const V3 a = m.V[t[0]], b = m.V[t[1]], c = m.V[t[2]];
        const V3 n = normalized(cross(sub(b, a), sub(c, a)));                                        // :394
        m.pn_tri[i] = n;
        const double a0 = std::acos(std::abs(dot(normalized(sub(b, a)), normalized(sub(c, a)))));   // :398-400
        const double a1 = std::acos(std::abs(dot(normalized(sub(a, b)), normalized(sub(c, b)))));
        const double a2 = std::acos(std::abs(dot(normalized(sub(b, c)), normalized(sub(a, c)))));
        m.pn_vert[t[0]] = add(m.pn_vert[t[0]], scale(a0, n));                                       // :401-403
        m.pn_vert[t[1]] = add(m.pn_vert[t[1]], scale(a1, n));
        m.pn_vert[t[2]] = add(m.pn_vert[t[2]], scale(a2, n));
        add_edge(t[0], t[1], n); add_edge(t[1], t[2], n); add_edge(t[0], t[2], n);                  // :406-408
    }
    for (V3& n : m.pn_vert) { const double l = norm(n); n.x /= l; n.y /= l; n.z /= l; }              // :411-413, :67
    for (int i = 0; i < nT; i++) {                                                                   // :415-420
        const int* t = &m.T[3 * i];
        m.pn_edge[3 * i + 0] = normalized(edge_normals[key_of(t[0], t[1])]);
        m.pn_edge[3 * i + 1] = normalized(edge_normals[key_of(t[1], t[2])]);
        m.pn_edge[3 * i + 2] = normalized(edge_normals[key_of(t[0], t[2])]);
    }
    m.flags = 0;                                                                                     // :422-438
    for (const auto& ec : edge_count) { if (ec.second == 1) m.flags |= 1; else if (ec.second > 2) m.flags |= 2; }
}

// TriangleMeshDistance.h:564-820
double point_triangle_sq_unsigned(int& entity, V3& nearest, V3 p, V3 v0, V3 v1, V3 v2)
{
    const V3 diff = sub(v0, p), e0 = sub(v1, v0), e1 = sub(v2, v0);
    const double a00 = dot(e0, e0), a01 = dot(e0, e1), a11 = dot(e1, e1);
    const double b0 = dot(diff, e0), b1 = dot(diff, e1), c = dot(diff, diff);
    const double det = std::abs(a00 * a11 - a01 * a01);
    double s = a01 * b1 - a11 * b0;
    double t = a01 * b0 - a00 * b1;
    double d2 = -1.0;
    if (s + t <= det) {
        if (s < 0) {
            if (t < 0) {                                   // region 4  :585-625
                if (b0 < 0) {
                    t = 0;
                    if (-b0 >= a00) { entity = V1; s = 1; d2 = a00 + 2 * b0 + c; }
                    else { entity = E01; s = -b0 / a00; d2 = b0 * s + c; }
                } else {
                    s = 0;
                    if (b1 >= 0) { entity = V0; t = 0; d2 = c; }
                    else if (-b1 >= a11) { entity = V2; t = 1; d2 = a11 + 2 * b1 + c; }
                    else { entity = E02; t = -b1 / a11; d2 = b1 * t + c; }
                }
            } else {                                       // region 3  :626-647
                s = 0;
                if (b1 >= 0) { entity = V0; t = 0; d2 = c; }
                else if (-b1 >= a11) { entity = V2; t = 1; d2 = a11 + 2 * b1 + c; }
                else { entity = E02; t = -b1 / a11; d2 = b1 * t + c; }
            }
        } else if (t < 0) {                                // region 5  :649-670
            t = 0;
            if (b0 >= 0) { entity = V0; s = 0; d2 = c; }
            else if (-b0 >= a00) { entity = V1; s = 1; d2 = a00 + 2 * b0 + c; }
            else { entity = E01; s = -b0 / a00; d2 = b0 * s + c; }
        } else {                                           // region 0  :671-680
            entity = F;
            const double invDet = 1 / det;
            s *= invDet; t *= invDet;
            d2 = s * (a00 * s + a01 * t + 2 * b0) + t * (a01 * s + a11 * t + 2 * b1) + c;
        }
    } else {
        double tmp0, tmp1, numer, denom;
        if (s < 0) {                                       // region 2  :686-732
            tmp0 = a01 + b0; tmp1 = a11 + b1;
            if (tmp1 > tmp0) {
                numer = tmp1 - tmp0; denom = a00 - 2 * a01 + a11;
                if (numer >= denom) { entity = V1; s = 1; t = 0; d2 = a00 + 2 * b0 + c; }
                else { entity = E12; s = numer / denom; t = 1 - s;
                       d2 = s * (a00 * s + a01 * t + 2 * b0) + t * (a01 * s + a11 * t + 2 * b1) + c; }
            } else {
                s = 0;
                if (tmp1 <= 0) { entity = V2; t = 1; d2 = a11 + 2 * b1 + c; }
                else if (b1 >= 0) { entity = V0; t = 0; d2 = c; }
                else { entity = E02; t = -b1 / a11; d2 = b1 * t + c; }
            }
        } else if (t < 0) {                                // region 6  :733-779
            tmp0 = a01 + b1; tmp1 = a00 + b0;
            if (tmp1 > tmp0) {
                numer = tmp1 - tmp0; denom = a00 - 2 * a01 + a11;
                if (numer >= denom) { entity = V2; t = 1; s = 0; d2 = a11 + 2 * b1 + c; }
                else { entity = E12; t = numer / denom; s = 1 - t;
                       d2 = s * (a00 * s + a01 * t + 2 * b0) + t * (a01 * s + a11 * t + 2 * b1) + c; }
            } else {
                t = 0;
                if (tmp1 <= 0) { entity = V1; s = 1; d2 = a00 + 2 * b0 + c; }
                else if (b0 >= 0) { entity = V0; s = 0; d2 = c; }
                else { entity = E01; s = -b0 / a00; d2 = b0 * s + c; }
            }
        } else {                                           // region 1  :780-809
            numer = a11 + b1 - a01 - b0;
            if (numer <= 0) { entity = V2; s = 0; t = 1; d2 = a11 + 2 * b1 + c; }
            else {
                denom = a00 - 2 * a01 + a11;
                if (numer >= denom) { entity = V1; s = 1; t = 0; d2 = a00 + 2 * b0 + c; }
                else { entity = E12; s = numer / denom; t = 1 - s;
                       d2 = s * (a00 * s + a01 * t + 2 * b0) + t * (a01 * s + a11 * t + 2 * b1) + c; }
            }
        }
    }
    if (d2 < 0) d2 = 0;                                    // :812-816
    nearest = add(add(v0, scale(s, e0)), scale(t, e1));    // :818
    return d2;
}

// TriangleMeshDistance.h:514-562
void query(const Mesh& m, Result& res, const Node& node, V3 p, long long* stats)
{
    if (stats) stats[0]++;
    if (node.left == -1) {
        if (stats) stats[1]++;
        const int tri = node.right;
        const int* t = &m.T[3 * tri];
        V3 np; int ent;
        const double d2 = point_triangle_sq_unsigned(ent, np, p, m.V[t[0]], m.V[t[1]], m.V[t[2]]);
        if (d2 < res.distance * res.distance) {
            res.nearest = np; res.entity = ent; res.distance = std::sqrt(d2); res.tri = tri;
        }
    } else {
        const double dl = norm(sub(p, node.bl.c)) - node.bl.r;
        const double dr = norm(sub(p, node.br.c)) - node.br.r;
        if (dl < dr) {
            if (dl < res.distance) query(m, res, m.nodes[node.left], p, stats);
            if (dr < res.distance) query(m, res, m.nodes[node.right], p, stats);
        } else {
            if (dr < res.distance) query(m, res, m.nodes[node.right], p, stats);
            if (dl < res.distance) query(m, res, m.nodes[node.left], p, stats);
        }
    }
}

// TriangleMeshDistance.h:316-328 and :269-308
Result unsigned_distance(const Mesh& m, V3 p, long long* stats = nullptr)
{
    Result r; r.distance = DBL_MAX; r.nearest = {0, 0, 0}; r.entity = 0; r.tri = -1;
    query(m, r, m.nodes[0], p, stats);
    return r;
}
Result signed_distance(const Mesh& m, V3 p, long long* stats = nullptr)
{
    Result r = unsigned_distance(m, p, stats);
    const int* t = &m.T[3 * r.tri];
    V3 n = {0, 0, 0};
    switch (r.entity) {
        case V0: n = m.pn_vert[t[0]]; break;
        case V1: n = m.pn_vert[t[1]]; break;
        case V2: n = m.pn_vert[t[2]]; break;
        case E01: n = m.pn_edge[3 * r.tri + 0]; break;
        case E12: n = m.pn_edge[3 * r.tri + 1]; break;
        case E02: n = m.pn_edge[3 * r.tri + 2]; break;
        case F: n = m.pn_tri[r.tri]; break;
    }
    const V3 u = sub(p, r.nearest);
    r.distance *= (dot(u, n) >= 0.0) ? 1.0 : -1.0;
    return r;
}

// ------------------------------------------------------------------ grid half
struct Grid {
    double mn[3], mx[3];
    unsigned n[3];
    double cell[3], inv[3];
};

// discregrid/src/cubic_lagrange_discrete_grid.cpp:604-665
void index_to_node_position(const Grid& g, unsigned l, double x[3])
{
    const unsigned* n = g.n;
    const unsigned nv = (n[0] + 1) * (n[1] + 1) * (n[2] + 1);
    const unsigned ne_x = (n[0] + 0) * (n[1] + 1) * (n[2] + 1);
    const unsigned ne_y = (n[0] + 1) * (n[1] + 0) * (n[2] + 1);
    unsigned ijk[3];
    int axis = -1; unsigned par = 0;
    if (l < nv) {
        ijk[2] = l / ((n[1] + 1) * (n[0] + 1));
        const unsigned temp = l % ((n[1] + 1) * (n[0] + 1));
        ijk[1] = temp / (n[0] + 1);
        ijk[0] = temp % (n[0] + 1);
    } else if (l < nv + 2 * ne_x) {
        l -= nv; const unsigned e = l / 2;
        ijk[2] = e / ((n[1] + 1) * n[0]);
        const unsigned temp = e % ((n[1] + 1) * n[0]);
        ijk[1] = temp / n[0]; ijk[0] = temp % n[0];
        axis = 0; par = l % 2;
    } else if (l < nv + 2 * (ne_x + ne_y)) {
        l -= (nv + 2 * ne_x); const unsigned e = l / 2;
        ijk[0] = e / ((n[2] + 1) * n[1]);
        const unsigned temp = e % ((n[2] + 1) * n[1]);
        ijk[2] = temp / n[1]; ijk[1] = temp % n[1];
        axis = 1; par = l % 2;
    } else {
        l -= (nv + 2 * (ne_x + ne_y)); const unsigned e = l / 2;
        ijk[1] = e / ((n[0] + 1) * n[2]);
        const unsigned temp = e % ((n[0] + 1) * n[2]);
        ijk[0] = temp / n[2]; ijk[2] = temp % n[2];
        axis = 2; par = l % 2;
    }
    for (int d = 0; d < 3; d++) x[d] = g.mn[d] + g.cell[d] * (double)ijk[d];       // :625,636,648,660
    if (axis >= 0) x[axis] += (1.0 + (double)par) / 3.0 * g.cell[axis];             // :637,649,661
}

// cubic_lagrange_discrete_grid.cpp:836-886
void build_cell(const unsigned n[3], unsigned l, unsigned cell[32])
{
    const unsigned nx = n[0], ny = n[1], nz = n[2];
    const unsigned k = l / (ny * nx), temp = l % (ny * nx), j = temp / nx, i = temp % nx;
    const unsigned nv = (nx + 1) * (ny + 1) * (nz + 1);
    const unsigned ne_x = nx * (ny + 1) * (nz + 1), ne_y = (nx + 1) * ny * (nz + 1);
    cell[0] = (nx + 1) * (ny + 1) * k + (nx + 1) * j + i;
    cell[1] = (nx + 1) * (ny + 1) * k + (nx + 1) * j + i + 1;
    cell[2] = (nx + 1) * (ny + 1) * k + (nx + 1) * (j + 1) + i;
    cell[3] = (nx + 1) * (ny + 1) * k + (nx + 1) * (j + 1) + i + 1;
    cell[4] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * j + i;
    cell[5] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * j + i + 1;
    cell[6] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * (j + 1) + i;
    cell[7] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * (j + 1) + i + 1;
    unsigned off = nv;
    cell[8] = off + 2 * (nx * (ny + 1) * k + nx * j + i);               cell[9] = cell[8] + 1;
    cell[10] = off + 2 * (nx * (ny + 1) * (k + 1) + nx * j + i);        cell[11] = cell[10] + 1;
    cell[12] = off + 2 * (nx * (ny + 1) * k + nx * (j + 1) + i);        cell[13] = cell[12] + 1;
    cell[14] = off + 2 * (nx * (ny + 1) * (k + 1) + nx * (j + 1) + i);  cell[15] = cell[14] + 1;
    off += 2 * ne_x;
    cell[16] = off + 2 * (ny * (nz + 1) * i + ny * k + j);              cell[17] = cell[16] + 1;
    cell[18] = off + 2 * (ny * (nz + 1) * (i + 1) + ny * k + j);        cell[19] = cell[18] + 1;
    cell[20] = off + 2 * (ny * (nz + 1) * i + ny * (k + 1) + j);        cell[21] = cell[20] + 1;
    cell[22] = off + 2 * (ny * (nz + 1) * (i + 1) + ny * (k + 1) + j);  cell[23] = cell[22] + 1;
    off += 2 * ne_y;
    cell[24] = off + 2 * (nz * (nx + 1) * j + nz * i + k);              cell[25] = cell[24] + 1;
    cell[26] = off + 2 * (nz * (nx + 1) * (j + 1) + nz * i + k);        cell[27] = cell[26] + 1;
    cell[28] = off + 2 * (nz * (nx + 1) * j + nz * (i + 1) + k);        cell[29] = cell[28] + 1;
    cell[30] = off + 2 * (nz * (nx + 1) * (j + 1) + nz * (i + 1) + k);  cell[31] = cell[30] + 1;
}

// cubic_lagrange_discrete_grid.cpp:339-580 (shape_function_, the live one)
void shape_function_(const double xi[3], double N[32], double (*dN)[3])
{
    const double x = xi[0], y = xi[1], z = xi[2];
    const double x2 = x * x, y2 = y * y, z2 = z * z;
    const double _1mx = 1.0 - x, _1my = 1.0 - y, _1mz = 1.0 - z;
    const double _1px = 1.0 + x, _1py = 1.0 + y, _1pz = 1.0 + z;
    const double _1m3x = 1.0 - 3.0 * x, _1m3y = 1.0 - 3.0 * y, _1m3z = 1.0 - 3.0 * z;
    const double _1p3x = 1.0 + 3.0 * x, _1p3y = 1.0 + 3.0 * y, _1p3z = 1.0 + 3.0 * z;
    const double _1mxt1my = _1mx * _1my, _1mxt1py = _1mx * _1py, _1pxt1my = _1px * _1my, _1pxt1py = _1px * _1py;
    const double _1mxt1mz = _1mx * _1mz, _1mxt1pz = _1mx * _1pz, _1pxt1mz = _1px * _1mz, _1pxt1pz = _1px * _1pz;
    const double _1myt1mz = _1my * _1mz, _1myt1pz = _1my * _1pz, _1pyt1mz = _1py * _1mz, _1pyt1pz = _1py * _1pz;
    const double _1mx2 = 1.0 - x2, _1my2 = 1.0 - y2, _1mz2 = 1.0 - z2;

    double fac = 1.0 / 64.0 * (9.0 * (x2 + y2 + z2) - 19.0);              // :388
    N[0] = fac * _1mxt1my * _1mz; N[1] = fac * _1pxt1my * _1mz; N[2] = fac * _1mxt1py * _1mz; N[3] = fac * _1pxt1py * _1mz;
    N[4] = fac * _1mxt1my * _1pz; N[5] = fac * _1pxt1my * _1pz; N[6] = fac * _1mxt1py * _1pz; N[7] = fac * _1pxt1py * _1pz;

    fac = 9.0 / 64.0 * _1mx2;                                             // :400
    const double fact1m3x = fac * _1m3x, fact1p3x = fac * _1p3x;
    N[8] = fact1m3x * _1myt1mz;  N[9] = fact1p3x * _1myt1mz;  N[10] = fact1m3x * _1myt1pz; N[11] = fact1p3x * _1myt1pz;
    N[12] = fact1m3x * _1pyt1mz; N[13] = fact1p3x * _1pyt1mz; N[14] = fact1m3x * _1pyt1pz; N[15] = fact1p3x * _1pyt1pz;

    fac = 9.0 / 64.0 * _1my2;                                             // :412
    const double fact1m3y = fac * _1m3y, fact1p3y = fac * _1p3y;
    N[16] = fact1m3y * _1mxt1mz; N[17] = fact1p3y * _1mxt1mz; N[18] = fact1m3y * _1pxt1mz; N[19] = fact1p3y * _1pxt1mz;
    N[20] = fact1m3y * _1mxt1pz; N[21] = fact1p3y * _1mxt1pz; N[22] = fact1m3y * _1pxt1pz; N[23] = fact1p3y * _1pxt1pz;

    fac = 9.0 / 64.0 * _1mz2;                                             // :424
    const double fact1m3z = fac * _1m3z, fact1p3z = fac * _1p3z;
    N[24] = fact1m3z * _1mxt1my; N[25] = fact1p3z * _1mxt1my; N[26] = fact1m3z * _1mxt1py; N[27] = fact1p3z * _1mxt1py;
    N[28] = fact1m3z * _1pxt1my; N[29] = fact1p3z * _1pxt1my; N[30] = fact1m3z * _1pxt1py; N[31] = fact1p3z * _1pxt1py;

    if (!dN) return;
    const double A = 9.0 * (3.0 * x2 + y2 + z2) - 19.0;                   // :440-442
    const double B = 9.0 * (x2 + 3.0 * y2 + z2) - 19.0;
    const double C = 9.0 * (x2 + y2 + 3.0 * z2) - 19.0;
    const double _18x = 18.0 * x, _18y = 18.0 * y, _18z = 18.0 * z;
    const double _3m9x2 = 3.0 - 9.0 * x2, _3m9y2 = 3.0 - 9.0 * y2, _3m9z2 = 3.0 - 9.0 * z2;
    const double _2x = 2.0 * x, _2y = 2.0 * y, _2z = 2.0 * z;
    const double xmA = _18x - A, xpA = _18x + A, ymB = _18y - B, ypB = _18y + B, zmC = _18z - C, zpC = _18z + C;
    dN[0][0] = xmA * _1myt1mz; dN[0][1] = _1mxt1mz * ymB; dN[0][2] = _1mxt1my * zmC;   // :462-485
    dN[1][0] = xpA * _1myt1mz; dN[1][1] = _1pxt1mz * ymB; dN[1][2] = _1pxt1my * zmC;
    dN[2][0] = xmA * _1pyt1mz; dN[2][1] = _1mxt1mz * ypB; dN[2][2] = _1mxt1py * zmC;
    dN[3][0] = xpA * _1pyt1mz; dN[3][1] = _1pxt1mz * ypB; dN[3][2] = _1pxt1py * zmC;
    dN[4][0] = xmA * _1myt1pz; dN[4][1] = _1mxt1pz * ymB; dN[4][2] = _1mxt1my * zpC;
    dN[5][0] = xpA * _1myt1pz; dN[5][1] = _1pxt1pz * ymB; dN[5][2] = _1pxt1my * zpC;
    dN[6][0] = xmA * _1pyt1pz; dN[6][1] = _1mxt1pz * ypB; dN[6][2] = _1mxt1py * zpC;
    dN[7][0] = xpA * _1pyt1pz; dN[7][1] = _1pxt1pz * ypB; dN[7][2] = _1pxt1py * zpC;
    for (int r = 0; r < 8; r++) for (int c = 0; c < 3; c++) dN[r][c] /= 64.0;          // :487

    const double mX = -_3m9x2 - _2x, pX = _3m9x2 - _2x;                                // :489-516
    const double Xm = _1mx2 * _1m3x, Xp = _1mx2 * _1p3x;
    dN[8][0] = mX * _1myt1mz;  dN[8][1] = -Xm * _1mz;  dN[8][2] = -Xm * _1my;
    dN[9][0] = pX * _1myt1mz;  dN[9][1] = -Xp * _1mz;  dN[9][2] = -Xp * _1my;
    dN[10][0] = mX * _1myt1pz; dN[10][1] = -Xm * _1pz; dN[10][2] = Xm * _1my;
    dN[11][0] = pX * _1myt1pz; dN[11][1] = -Xp * _1pz; dN[11][2] = Xp * _1my;
    dN[12][0] = mX * _1pyt1mz; dN[12][1] = Xm * _1mz;  dN[12][2] = -Xm * _1py;
    dN[13][0] = pX * _1pyt1mz; dN[13][1] = Xp * _1mz;  dN[13][2] = -Xp * _1py;
    dN[14][0] = mX * _1pyt1pz; dN[14][1] = Xm * _1pz;  dN[14][2] = Xm * _1py;
    dN[15][0] = pX * _1pyt1pz; dN[15][1] = Xp * _1pz;  dN[15][2] = Xp * _1py;

    const double mY = -_3m9y2 - _2y, pY = _3m9y2 - _2y;                                // :518-545
    const double Ym = _1my2 * _1m3y, Yp = _1my2 * _1p3y;
    dN[16][0] = -Ym * _1mz; dN[16][1] = mY * _1mxt1mz; dN[16][2] = -Ym * _1mx;
    dN[17][0] = -Yp * _1mz; dN[17][1] = pY * _1mxt1mz; dN[17][2] = -Yp * _1mx;
    dN[18][0] = Ym * _1mz;  dN[18][1] = mY * _1pxt1mz; dN[18][2] = -Ym * _1px;
    dN[19][0] = Yp * _1mz;  dN[19][1] = pY * _1pxt1mz; dN[19][2] = -Yp * _1px;
    dN[20][0] = -Ym * _1pz; dN[20][1] = mY * _1mxt1pz; dN[20][2] = Ym * _1mx;
    dN[21][0] = -Yp * _1pz; dN[21][1] = pY * _1mxt1pz; dN[21][2] = Yp * _1mx;
    dN[22][0] = Ym * _1pz;  dN[22][1] = mY * _1pxt1pz; dN[22][2] = Ym * _1px;
    dN[23][0] = Yp * _1pz;  dN[23][1] = pY * _1pxt1pz; dN[23][2] = Yp * _1px;

    const double mZ = -_3m9z2 - _2z, pZ = _3m9z2 - _2z;                                // :547-574
    const double Zm = _1mz2 * _1m3z, Zp = _1mz2 * _1p3z;
    dN[24][0] = -Zm * _1my; dN[24][1] = -Zm * _1mx; dN[24][2] = mZ * _1mxt1my;
    dN[25][0] = -Zp * _1my; dN[25][1] = -Zp * _1mx; dN[25][2] = pZ * _1mxt1my;
    dN[26][0] = -Zm * _1py; dN[26][1] = Zm * _1mx;  dN[26][2] = mZ * _1mxt1py;
    dN[27][0] = -Zp * _1py; dN[27][1] = Zp * _1mx;  dN[27][2] = pZ * _1mxt1py;
    dN[28][0] = Zm * _1my;  dN[28][1] = -Zm * _1px; dN[28][2] = mZ * _1pxt1my;
    dN[29][0] = Zp * _1my;  dN[29][1] = -Zp * _1px; dN[29][2] = pZ * _1pxt1my;
    dN[30][0] = Zm * _1py;  dN[30][1] = Zm * _1px;  dN[30][2] = mZ * _1pxt1py;
    dN[31][0] = Zp * _1py;  dN[31][1] = Zp * _1px;  dN[31][2] = pZ * _1pxt1py;
    for (int r = 8; r < 32; r++) for (int c = 0; c < 3; c++) dN[r][c] *= 9.0 / 64.0;    // :576
}

// cubic_lagrange_discrete_grid.cpp:977-1063 (+ discrete_grid.cpp:8-38).  cells==NULL => closed form
// (identical to the table addFunction builds, :833-886); cell_map==NULL => identity (:888-891).
double interpolate(const Grid& g, const double* nodes, const unsigned* cells, const unsigned* cell_map,
                   const double x[3], double* grad)
{
    for (int d = 0; d < 3; d++) if (!(g.mn[d] <= x[d] && x[d] <= g.mx[d])) return DBL_MAX;   // :981 (AlignedBox::contains)
    unsigned mi[3];
    for (int d = 0; d < 3; d++) {
        mi[d] = (unsigned)((x[d] - g.mn[d]) * g.inv[d]);                                      // :984
        if (mi[d] >= g.n[d]) mi[d] = g.n[d] - 1;                                              // :985-990
    }
    unsigned i = g.n[1] * g.n[0] * mi[2] + g.n[0] * mi[1] + mi[0];                            // discrete_grid.cpp:20-24
    const unsigned i_ = cell_map ? cell_map[i] : i;
    if (i_ == UINT_MAX) return DBL_MAX;                                                       // :993
    // subdomain(i): singleToMultiIndex then origin (discrete_grid.cpp:8-38) -- same ijk as mi
    double c0[3], xi[3];
    for (int d = 0; d < 3; d++) {
        const double lo = g.mn[d] + (double)mi[d] * g.cell[d];
        const double hi = lo + g.cell[d];
        const double denom = hi - lo;                                                         // :1000
        c0[d] = 2.0 / denom;                                                                  // :1001
        const double c1 = (hi + lo) / denom;                                                  // :1002
        xi[d] = c0[d] * x[d] - c1;                                                            // :1003
    }
    unsigned closed[32];
    const unsigned* cell;
    if (cells) cell = cells + 32 * (size_t)i_; else { build_cell(g.n, i_, closed); cell = closed; }
    double N[32];
    if (!grad) {                                                                              // :1006-1023
        shape_function_(xi, N, nullptr);
        double phi = 0.0;
        for (unsigned j = 0; j < 32; j++) {
            const double c = nodes[cell[j]];
            if (c == DBL_MAX) return DBL_MAX;
            phi += c * N[j];
        }
        return phi;
    }
    double dN[32][3];                                                                         // :1025-1062
    shape_function_(xi, N, dN);
    double phi = 0.0;
    grad[0] = grad[1] = grad[2] = 0.0;
    for (unsigned j = 0; j < 32; j++) {
        const double c = nodes[cell[j]];
        if (c == DBL_MAX) { grad[0] = grad[1] = grad[2] = 0.0; return DBL_MAX; }
        phi += c * N[j];
        grad[0] += c * dN[j][0]; grad[1] += c * dN[j][1]; grad[2] += c * dN[j][2];
    }
    grad[0] *= c0[0]; grad[1] *= c0[1]; grad[2] *= c0[2];
    return phi;
}

// cmd/generate_density_map/gauss_quadrature.cpp:616-632 / :3422-3438 (p = 30 -> n = 16, :100-101)
const double GA16[16] = {
    -0.989400934991649938510249739920, -0.944575023073232600268056557979, -0.865631202387831755196145877562,
    -0.755404408355002998654015300417, -0.617876244402643770570193737512, -0.458016777657227369680015272024,
    -0.281603550779258915426339626720, -0.095012509837637426635126303154, 0.095012509837637426635126303154,
    0.281603550779258915426339626720, 0.458016777657227369680015272024, 0.617876244402643770570193737512,
    0.755404408355002998654015300417, 0.865631202387831755196145877562, 0.944575023073232600268056557979,
    0.989400934991649938510249739920};
const double GW16[16] = {
    0.027152459411758110563450685504, 0.062253523938649010793788818319, 0.095158511682492036287683845330,
    0.124628971255533488315947465708, 0.149595988816575764523975067277, 0.169156519395001675443168664970,
    0.182603415044922529064663763165, 0.189450610455067447457366824892, 0.189450610455067447457366824892,
    0.182603415044922529064663763165, 0.169156519395001675443168664970, 0.149595988816575764523975067277,
    0.124628971255533488315947465708, 0.095158511682492036287683845330, 0.062253523938649010793788818319,
    0.027152459411758110563450685504};

// Eigen's 3-vector norm() = sqrt(cwiseAbs2().sum()).  Eigen >= 3.3 (what the reference CI apt-installs,
// .github/workflows/build-linux.yml:12-13) reduces a Vector3d with one SSE2 packet + scalar tail:
// (a0 + a1) + a2.  Eigen 3.2 used the scalar unroller a0 + (a1 + a2).  Eigen is not in this container so
// this cannot be run here; box.cdf cannot discriminate (its diagonal has three equal components).  The
// order only touches cell_diag, W(r) and the GenerateSDF padding -- never K1/K2 given the same domain.
// Default: modern Eigen.  -DORC_EIGEN32_NORM selects the 3.2 order.
static inline double norm3_eigen(double a, double b, double c)
{
#ifdef ORC_EIGEN32_NORM
    return std::sqrt(a * a + (b * b + c * c));
#else
    return std::sqrt((a * a + b * b) + c * c);
#endif
}

// cmd/generate_density_map/sph_kernel.hpp:10-42
struct CubicKernel {
    double h, k;
    void set(double r) { h = r; const double pi = M_PI; const double h3 = h * h * h; k = 8.0 / (pi * h3); }
    double W(double rx, double ry, double rz) const {
        double res = 0.0;
        const double rl = norm3_eigen(rx, ry, rz);
        const double q = rl / h;
        if (q <= 1.0) {
            if (q <= 0.5) { const double q2 = q * q, q3 = q2 * q; res = k * (6.0 * q3 - 6.0 * q2 + 1.0); }
            else { const double m = 1.0 - q; res = k * (2.0 * m * m * m); }
        }
        return res;
    }
};

}  // namespace

// =============================================================== extern "C" surface
extern "C" {

void* orc_mesh_create(const double* V, uint64_t nV, const uint32_t* F, uint64_t nT)
{
    if (nT == 0) return nullptr;   // reference: prints + exit(-1), TriangleMeshDistance.h:338-341
    Mesh* m = new Mesh();
    m->V.resize(nV);
    for (uint64_t i = 0; i < nV; i++) m->V[i] = {V[3 * i], V[3 * i + 1], V[3 * i + 2]};
    m->T.resize(3 * nT);
    for (uint64_t i = 0; i < 3 * nT; i++) m->T[i] = (int)F[i];
    construct(*m);
    return m;
}
void orc_mesh_destroy(void* h) { delete (Mesh*)h; }
int orc_mesh_flags(void* h) { return ((Mesh*)h)->flags; }
uint64_t orc_mesh_num_nodes(void* h) { return ((Mesh*)h)->nodes.size(); }

// dump tree: spheres[n][8] = (cl, rl, cr, rr), kids[n][2] = (left, right)
void orc_mesh_tree(void* h, double* spheres, int32_t* kids)
{
    Mesh* m = (Mesh*)h;
    for (size_t i = 0; i < m->nodes.size(); i++) {
        const Node& n = m->nodes[i];
        double* s = spheres + 8 * i;
        s[0] = n.bl.c.x; s[1] = n.bl.c.y; s[2] = n.bl.c.z; s[3] = n.bl.r;
        s[4] = n.br.c.x; s[5] = n.br.c.y; s[6] = n.br.c.z; s[7] = n.br.r;
        kids[2 * i] = n.left; kids[2 * i + 1] = n.right;
    }
}
// pseudonormals: tri[nT][3], edge[nT][3][3], vert[nV][3]
void orc_mesh_pseudonormals(void* h, double* tri, double* edge, double* vert)
{
    Mesh* m = (Mesh*)h;
    memcpy(tri, m->pn_tri.data(), sizeof(V3) * m->pn_tri.size());
    memcpy(edge, m->pn_edge.data(), sizeof(V3) * m->pn_edge.size());
    memcpy(vert, m->pn_vert.data(), sizeof(V3) * m->pn_vert.size());
}

// batch point query; is_signed=0 -> unsigned_distance.  nearest/entity/tri nullable.
void orc_mesh_distance(void* h, const double* x, uint64_t n, int is_signed, double* dist, double* nearest,
                       int32_t* entity, int32_t* tri)
{
    const Mesh& m = *(Mesh*)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (long long q = 0; q < (long long)n; q++) {
        const V3 p = {x[3 * q], x[3 * q + 1], x[3 * q + 2]};
        const Result r = is_signed ? signed_distance(m, p) : unsigned_distance(m, p);
        dist[q] = r.distance;
        if (nearest) { nearest[3 * q] = r.nearest.x; nearest[3 * q + 1] = r.nearest.y; nearest[3 * q + 2] = r.nearest.z; }
        if (entity) entity[q] = r.entity;
        if (tri) tri[q] = r.tri;
    }
}
// traversal statistics (visits, leaf tests) summed over n points, single-threaded
void orc_mesh_stats(void* h, const double* x, uint64_t n, int64_t* visits, int64_t* leaves)
{
    const Mesh& m = *(Mesh*)h;
    long long st[2] = {0, 0};
    for (uint64_t q = 0; q < n; q++) signed_distance(m, {x[3 * q], x[3 * q + 1], x[3 * q + 2]}, st);
    *visits = st[0]; *leaves = st[1];
}

// grid: g = {min[3], max[3], cell[3], inv[3]} as 12 doubles; res[3]
static Grid mkgrid(const double* gd, const uint32_t* res)
{
    Grid g;
    for (int d = 0; d < 3; d++) { g.mn[d] = gd[d]; g.mx[d] = gd[3 + d]; g.cell[d] = gd[6 + d]; g.inv[d] = gd[9 + d]; g.n[d] = res[d]; }
    return g;
}
// discrete_grid.hpp:22-29
void orc_grid_constants(const double mn[3], const double mx[3], const uint32_t res[3], double cell[3], double inv[3])
{
    for (int d = 0; d < 3; d++) { cell[d] = (mx[d] - mn[d]) / (double)res[d]; inv[d] = 1.0 / cell[d]; }
}
// cmd/generate_sdf/main.cpp:83-91 (asymmetric padding; Eigen norm order, see norm3_eigen)
void orc_generate_sdf_domain(const double* V, uint64_t nV, double mn[3], double mx[3])
{
    for (int d = 0; d < 3; d++) { mn[d] = DBL_MAX; mx[d] = -DBL_MAX; }
    for (uint64_t i = 0; i < nV; i++) for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], V[3 * i + d]); mx[d] = std::max(mx[d], V[3 * i + d]); }
    { const double nrm = norm3_eigen(mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]); for (int d = 0; d < 3; d++) mx[d] += 1.0e-3 * nrm * 1.0; }
    { const double nrm = norm3_eigen(mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]); for (int d = 0; d < 3; d++) mn[d] -= 1.0e-3 * nrm * 1.0; }
}
uint64_t orc_num_nodes(const uint32_t n[3])
{   // cubic_lagrange_discrete_grid.cpp:790-796 (unsigned arithmetic in the reference; 64-bit here only for the return)
    const unsigned nv = (n[0] + 1) * (n[1] + 1) * (n[2] + 1);
    const unsigned ne = n[0] * (n[1] + 1) * (n[2] + 1) + (n[0] + 1) * n[1] * (n[2] + 1) + (n[0] + 1) * (n[1] + 1) * n[2];
    return (uint64_t)(nv + 2 * ne);
}
void orc_node_positions(const double* gd, const uint32_t* res, uint64_t l_begin, uint64_t l_end, double* x)
{
    const Grid g = mkgrid(gd, res);
#pragma omp parallel for schedule(static)
    for (long long l = (long long)l_begin; l < (long long)l_end; l++) index_to_node_position(g, (unsigned)l, x + 3 * (l - l_begin));
}
// indexToNodePosition for an arbitrary list of node ids (bench.py's strided CPU sample)
void orc_node_positions_at(const double* gd, const uint32_t* res, const uint64_t* ids, uint64_t n, double* x)
{
    const Grid g = mkgrid(gd, res);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)n; q++) index_to_node_position(g, (unsigned)ids[q], x + 3 * q);
}
// addFunction node loop (cubic_lagrange_discrete_grid.cpp:806-817) with func = sign * md.signed_distance(x).distance
// (cmd/generate_sdf/main.cpp:97,101).  schedule(static) like the reference.  nthreads<=0 -> OpenMP default.
void orc_sample_sdf(void* h, const double* gd, const uint32_t* res, double sign, uint64_t l_begin, uint64_t l_end,
                    double* out, int nthreads)
{
    const Mesh& m = *(Mesh*)h;
    const Grid g = mkgrid(gd, res);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (long long l = (long long)l_begin; l < (long long)l_end; l++) {
        double x[3];
        index_to_node_position(g, (unsigned)l, x);
        const double d = signed_distance(m, {x[0], x[1], x[2]}).distance;
        out[l - l_begin] = (sign == 1.0) ? d : sign * d;
    }
}
void orc_build_cells(const uint32_t* res, uint64_t c_begin, uint64_t c_end, uint32_t* cells)
{
#pragma omp parallel for schedule(static)
    for (long long l = (long long)c_begin; l < (long long)c_end; l++) build_cell(res, (unsigned)l, cells + 32 * (l - c_begin));
}
void orc_shape_functions(const double* xi, uint64_t n, double* N, double* dN /*nullable, n*32*3*/)
{
    for (uint64_t q = 0; q < n; q++) shape_function_(xi + 3 * q, N + 32 * q, dN ? (double (*)[3])(dN + 96 * q) : nullptr);
}
// batched interpolate in the pattern of cmd/discrete_field_to_bitmap/main.cpp:118-135
void orc_interpolate(const double* gd, const uint32_t* res, const double* nodes, const uint32_t* cells,
                     const uint32_t* cell_map, const double* x, uint64_t n, double* phi, double* grad, int nthreads)
{
    const Grid g = mkgrid(gd, res);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)n; q++) phi[q] = interpolate(g, nodes, cells, cell_map, x + 3 * q, grad ? grad + 3 * q : nullptr);
}

// GenerateDensityMap per-node function (cmd/generate_density_map/main.cpp:86-133,
// gauss_quadrature.cpp:5927-5960, sph_kernel.hpp:22-42), sampled on nodes [l_begin, l_end).
void orc_density_map(const double* gd, const uint32_t* res, const double* nodes, const uint32_t* cells,
                     const uint32_t* cell_map, double h, double rho0, int no_reduction, uint64_t l_begin,
                     uint64_t l_end, double* out, int nthreads)
{
    const Grid g = mkgrid(gd, res);
    CubicKernel ker; ker.set(h);
    const double cell_diag = norm3_eigen(g.cell[0], g.cell[1], g.cell[2]);          // main.cpp:117
    // GaussQuadrature::integrate: c0 = 0.5*diag, c1 = 0.5*(min+max) of [-h,h]^3
    const double c0 = 0.5 * (h - (-h)), c1 = 0.5 * ((-h) + h);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (long long l = (long long)l_begin; l < (long long)l_end; l++) {
        double x[3];
        index_to_node_position(g, (unsigned)l, x);
        bool keep = true;
        if (!no_reduction) {                                                          // main.cpp:119-133
            double xc[3];
            for (int d = 0; d < 3; d++) xc[d] = std::min(std::max(x[d], g.mn[d]), g.mx[d]);
            const double dist = interpolate(g, nodes, cells, cell_map, xc, nullptr);
            if (dist == DBL_MAX) keep = false;
            else keep = (-6.0 * h < dist + cell_diag) && (dist - cell_diag < 2.0 * h);
        }
        if (!keep) { out[l - l_begin] = DBL_MAX; continue; }                          // cubic_lagrange_discrete_grid.cpp:814-817
        const double dist = interpolate(g, nodes, cells, cell_map, x, nullptr);       // main.cpp:98
        if (dist > 2.0 * h) { out[l - l_begin] = 0.0; continue; }
        double res = 0.0;
        for (int i = 0; i < 16; i++) {                                                // gauss_quadrature.cpp:5942-5956
            const double wi = GW16[i];
            const double yx = c0 * GA16[i] + c1;
            for (int j = 0; j < 16; j++) {
                const double wij = wi * GW16[j];
                const double yy = c0 * GA16[j] + c1;
                for (int k = 0; k < 16; k++) {
                    const double wijk = wij * GW16[k];
                    const double yz = c0 * GA16[k] + c1;
                    const double xs[3] = {x[0] + yx, x[1] + yy, x[2] + yz};
                    const double d = interpolate(g, nodes, cells, cell_map, xs, nullptr);  // gamma, main.cpp:86-93
                    const double gam = (d > h) ? 0.0 : 1.0 - d / h;
                    res += wijk * (gam * ker.W(yx, yy, yz));
                }
            }
        }
        res *= c0 * c0 * c0;                                                          // :5958 (Eigen prod(): (c0*c0)*c0)
        out[l - l_begin] = rho0 * res;
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
