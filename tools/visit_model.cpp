// CPU model of the K1 traversal's WORK (node steps, leaf tests) under different pruning rules -- an experiment bench, not
// product code and not a parity oracle: distances use a plain closest-point routine, ties are irrelevant for counting.
//   A  reference rule: sphere lower bounds only (TriangleMeshDistance.h:514-562)
//   B  the kernel's rule: spheres + "hopeless" box skip against the running best (k1_sdf.cu)
//   C  B + an external pruning bound U = d(brick centre) + |x - centre| (one extra query per 4x4x2 brick)
//   D  B + the ideal bound U = d(x) (1 + 1e-6): the least work this tree allows
// build: g++ -O2 -fopenmp -ffp-contract=off -I discregrid_b200/csrc tools/visit_model.cpp discregrid_b200/csrc/bvh_build.cpp discregrid_b200/csrc/sort_replay.cpp -o visit_model -lpthread
// usage: visit_model mesh.V mesh.F resolution [brick_stride]
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "bvh_build.h"

using namespace dgb;

struct P3 { double x, y, z; };
static inline P3 sub(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// squared distance point-triangle (Ericson, Real-Time Collision Detection 5.1.5)
static double tri_d2(const LeafRecord& L, P3 p)
{
    const P3 a = {L.v0[0], L.v0[1], L.v0[2]}, ab = {L.e0[0], L.e0[1], L.e0[2]}, ac = {L.e1[0], L.e1[1], L.e1[2]};
    const P3 ap = sub(p, a);
    const double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return dot(ap, ap);
    const P3 b = {a.x + ab.x, a.y + ab.y, a.z + ab.z}, bp = sub(p, b);
    const double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return dot(bp, bp);
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { const double v = d1 / (d1 - d3); const P3 q = {ap.x - v * ab.x, ap.y - v * ab.y, ap.z - v * ab.z}; return dot(q, q); }
    const P3 c = {a.x + ac.x, a.y + ac.y, a.z + ac.z}, cp = sub(p, c);
    const double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return dot(cp, cp);
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { const double w = d2 / (d2 - d6); const P3 q = {ap.x - w * ac.x, ap.y - w * ac.y, ap.z - w * ac.z}; return dot(q, q); }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        const P3 bc = sub(c, b); const P3 q = {bp.x - w * bc.x, bp.y - w * bc.y, bp.z - w * bc.z}; return dot(q, q);
    }
    const double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
    const P3 q = {ap.x - v * ab.x - w * ac.x, ap.y - v * ab.y - w * ac.y, ap.z - v * ab.z - w * ac.z};
    return dot(q, q);
}

struct Counts { long long nodes = 0, leaves = 0, pops = 0; };

struct Model {
    const HostBvh& H;
    int T;
    explicit Model(const HostBvh& h) : H(h), T((int)h.n_triangles) {}
    static double sph(const double* c, double r, P3 p) { const double x = p.x - c[0], y = p.y - c[1], z = p.z - c[2]; return std::sqrt(x * x + y * y + z * z) - r; }
    // distance to child box `which` (0 left, 1 right) of the internal node with split m; boxes are relative to nothing here (fp64 absolute)
    double boxd(int m, int which, P3 p) const
    {
        const BoxPairF& Bx = H.boxes_f[m];
        const float* lo = which ? Bx.r_lo : Bx.l_lo; const float* hi = which ? Bx.r_hi : Bx.l_hi;
        const double qx = p.x - H.center[0], qy = p.y - H.center[1], qz = p.z - H.center[2];
        const double gx = std::max(std::max(lo[0] - qx, qx - hi[0]), 0.0), gy = std::max(std::max(lo[1] - qy, qy - hi[1]), 0.0),
                     gz = std::max(std::max(lo[2] - qz, qz - hi[2]), 0.0);
        return std::sqrt(gx * gx + gy * gy + gz * gz);
    }
    struct Item { int b, e; double d; int m; int which; };
    // use_box: the kernel's hopeless test; U: external pruning bound (DBL_MAX = none)
    double query(P3 p, bool use_box, double U, Counts& c) const
    {
        double best = DBL_MAX;
        std::vector<Item> st; st.reserve(64);
        int b = 0, e = T;
        bool have = true;
        for (;;) {
            if (!have) {
                bool found = false;
                while (!st.empty()) {
                    const Item it = st.back(); st.pop_back(); c.pops++;
                    if (!(it.d < best)) continue;
                    if (it.d > U) continue;
                    if (use_box || U < DBL_MAX) { const double bd = boxd(it.m, it.which, p); if ((use_box && bd > best) || bd > U) continue; }
                    b = it.b; e = it.e; found = true; break;
                }
                if (!found) break;
            }
            have = false;
            if (e - b == 1) {
                c.leaves++;
                const double d2 = tri_d2(H.leaves[b], p);
                if (d2 < best * best) best = std::sqrt(d2);
                continue;
            }
            c.nodes++;
            const int m = (b + e) >> 1;
            const SpherePair& S = H.spheres[m];
            const double dl = sph(S.lc, S.lr, p), dr = sph(S.rc, S.rr, p);
            const bool lf = dl < dr;
            const double d1 = lf ? dl : dr, d2 = lf ? dr : dl;
            bool go1 = d1 < best && !(d1 > U), def2 = d2 < best && !(d2 > U);
            if (go1 && (use_box || U < DBL_MAX)) {
                const double b1 = boxd(m, lf ? 0 : 1, p), b2 = boxd(m, lf ? 1 : 0, p);
                if ((use_box && b1 > best) || b1 > U) go1 = false;
                if ((use_box && b2 > best) || b2 > U) def2 = false;
            }
            const int fb = lf ? b : m, fe = lf ? m : e, sb = lf ? m : b, se = lf ? e : m;
            if (go1) {
                if (def2) st.push_back({sb, se, d2, m, lf ? 1 : 0});
                b = fb; e = fe; have = true;
            } else if (def2) {       // first child skipped: the second is what the walk turns to next
                b = sb; e = se; have = true;
            }
        }
        return best;
    }
};

static std::vector<unsigned char> slurp(const char* path)
{
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) exit(1); fclose(f); return v;
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s mesh.V mesh.F resolution [brick_stride]\n", argv[0]); return 1; }
    auto vb = slurp(argv[1]), fb = slurp(argv[2]);
    const int res = atoi(argv[3]), stride = argc > 4 ? atoi(argv[4]) : 3;
    const uint64_t nV = vb.size() / 24, nT = fb.size() / 12;
    const double* V = (const double*)vb.data(); const uint32_t* F = (const uint32_t*)fb.data();
    HostBvh H; const char* err = nullptr;
    if (!build_host_bvh(V, nV, F, nT, H, &err)) { fprintf(stderr, "build: %s\n", err); return 1; }
    // GenerateSDF domain rule (cmd/generate_sdf/main.cpp:83-91)
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint64_t i = 0; i < nV; i++) for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], V[3 * i + d]); mx[d] = std::max(mx[d], V[3 * i + d]); }
    const double diag = std::sqrt((mx[0] - mn[0]) * (mx[0] - mn[0]) + (mx[1] - mn[1]) * (mx[1] - mn[1]) + (mx[2] - mn[2]) * (mx[2] - mn[2]));
    for (int d = 0; d < 3; d++) { mn[d] -= 1e-3 * diag; mx[d] += 1e-3 * diag; }
    double cell[3]; for (int d = 0; d < 3; d++) cell[d] = (mx[d] - mn[d]) / res;
    Model M(H);
    const int nb[3] = {(res + 1 + 3) / 4, (res + 1 + 3) / 4, (res + 1 + 1) / 2};
    Counts A, B, C, Cc, D; long long nq = 0, nbricks = 0;
    // histogram of per-query node steps under B (to see the spread that makes launch tails)
#pragma omp parallel
    {
        Counts a, b, c, cc, d; long long q = 0, bricks = 0;
#pragma omp for schedule(dynamic, 4) nowait
        for (int bz = 0; bz < nb[2]; bz++)
            for (int by = 0; by < nb[1]; by++)
                for (int bx = 0; bx < nb[0]; bx++) {
                    if ((bx + 3 * by + 7 * bz) % stride) continue;
                    bricks++;
                    const P3 ctr = {mn[0] + cell[0] * (4 * bx + 1.5), mn[1] + cell[1] * (4 * by + 1.5), mn[2] + cell[2] * (2 * bz + 0.5)};
                    const double dc = M.query(ctr, true, DBL_MAX, cc);
                    for (int k = 0; k < 2; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
                        const int gi = 4 * bx + i, gj = 4 * by + j, gk = 2 * bz + k;
                        if (gi > res || gj > res || gk > res) continue;
                        const P3 p = {mn[0] + cell[0] * gi, mn[1] + cell[1] * gj, mn[2] + cell[2] * gk};
                        q++;
                        const double dA = M.query(p, false, DBL_MAX, a);
                        const double dB = M.query(p, true, DBL_MAX, b);
                        const P3 off = sub(p, ctr);
                        const double U = (dc + std::sqrt(dot(off, off))) * (1 + 1e-6);
                        const double dC = M.query(p, true, U, c);
                        const double dD = M.query(p, true, dA * (1 + 1e-6), d);
                        if (dA != dB || dA != dC || dA != dD) { fprintf(stderr, "MISMATCH %g %g %g %g\n", dA, dB, dC, dD); }
                    }
                }
#pragma omp critical
        {
            A.nodes += a.nodes; A.leaves += a.leaves; B.nodes += b.nodes; B.leaves += b.leaves; C.nodes += c.nodes; C.leaves += c.leaves;
            Cc.nodes += cc.nodes; Cc.leaves += cc.leaves; D.nodes += d.nodes; D.leaves += d.leaves; nq += q; nbricks += bricks;
            A.pops += a.pops; B.pops += b.pops; C.pops += c.pops; D.pops += d.pops;
        }
    }
    printf("triangles %llu, %d^3 vertex lattice, %lld queries in %lld bricks\n", (unsigned long long)nT, res, nq, nbricks);
    auto row = [&](const char* name, const Counts& c, double extra_nodes, double extra_leaves) {
        printf("%-44s node steps %8.1f  leaf tests %7.1f  pops %7.1f per query\n", name, (c.nodes + extra_nodes) / (double)nq, (c.leaves + extra_leaves) / (double)nq, c.pops / (double)nq);
    };
    row("A reference (spheres only)", A, 0, 0);
    row("B kernel rule (spheres + box skip)", B, 0, 0);
    row("C B + bound from the brick-centre query", C, 0, 0);
    row("C including the centre queries", C, (double)Cc.nodes, (double)Cc.leaves);
    row("D B + ideal bound", D, 0, 0);
    return 0;
}
