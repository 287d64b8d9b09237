// CPU model of a PACKET traversal for K1: the 32 queries of a brick walk the tree TOGETHER (one shared stack, one node per
// step for all lanes, order by majority), each lane pruning with its own best.  Counts what the warp would execute so it can
// be weighed against the per-lane kernel's instruction count (tools/warp_model.cpp) BEFORE any device code is written.
// An experiment bench, not product code and not a parity oracle.
// build: g++ -O2 -fopenmp -ffp-contract=off -I discregrid_b200/csrc tools/packet_model.cpp discregrid_b200/csrc/bvh_build.cpp discregrid_b200/csrc/sort_replay.cpp discregrid_b200/csrc/host_threads.cpp -o packet_model -lpthread
// usage: packet_model mesh.V mesh.F resolution [brick_stride]
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


struct PCounts { long long nodes_even = 0, nodes_odd = 0, nodes = 0, filters = 0, exacts = 0, exact_lanes = 0, pops = 0, pops_hit = 0, bricks = 0, queries = 0, maxstack = 0; };

struct Packet {
    const HostBvh& H; int T;
    explicit Packet(const HostBvh& h) : H(h), T((int)h.n_triangles) {}
    static double sph(const double* c, double r, P3 p) { const double x = p.x - c[0], y = p.y - c[1], z = p.z - c[2]; return std::sqrt(x * x + y * y + z * z) - r; }
    double boxd(int m, int which, P3 p) const
    {
        const BoxPairF& Bx = H.boxes_f[m];
        const float* lo = which ? Bx.r_lo : Bx.l_lo; const float* hi = which ? Bx.r_hi : Bx.l_hi;
        const double qx = p.x - H.center[0], qy = p.y - H.center[1], qz = p.z - H.center[2];
        const double gx = std::max(std::max(lo[0] - qx, qx - hi[0]), 0.0), gy = std::max(std::max(lo[1] - qy, qy - hi[1]), 0.0),
                     gz = std::max(std::max(lo[2] - qz, qz - hi[2]), 0.0);
        return std::sqrt(gx * gx + gy * gy + gz * gz);
    }
    // lower bound of child `which` of the node split at m for point p (sphere and box, as the kernel's rule B)
    double lb(int m, int which, P3 p) const
    {
        const SpherePair& S = H.spheres[m];
        const double s = which ? sph(S.rc, S.rr, p) : sph(S.lc, S.lr, p);
        return std::max(s, boxd(m, which, p));
    }
    struct Item { int b, e, m, which, depth; };
    void brick(const P3* p, int n, double* best, PCounts& c, double slack, int order_mode) const
    {
        for (int i = 0; i < n; i++) best[i] = DBL_MAX;
        std::vector<Item> st; st.reserve(128);
        int b = 0, e = T, depth = 0; bool have = true;
        for (;;) {
            if (!have) {
                bool found = false;
                while (!st.empty()) {
                    const Item it = st.back(); st.pop_back(); c.pops++;
                    bool any = false;
                    for (int i = 0; i < n && !any; i++) any = lb(it.m, it.which, p[i]) < best[i] * slack;
                    if (!any) continue;
                    c.pops_hit++; b = it.b; e = it.e; depth = it.depth; found = true; break;
                }
                if (!found) break;
            }
            have = false;
            if (e - b == 1) {
                c.filters++;
                int pass = 0;
                double d2[32];
                for (int i = 0; i < n; i++) { d2[i] = tri_d2(H.leaves[b], p[i]); if (std::sqrt(d2[i]) < best[i] * slack) pass++; }
                if (pass) { c.exacts++; c.exact_lanes += pass; for (int i = 0; i < n; i++) if (d2[i] < best[i] * best[i]) best[i] = std::sqrt(d2[i]); }
                continue;
            }
            c.nodes++; if (depth & 1) c.nodes_odd++; else c.nodes_even++;
            const int m = (b + e) >> 1;
            int wantL = 0, wantR = 0, prefL = 0, prefR = 0; double sumL = 0, sumR = 0;
            for (int i = 0; i < n; i++) {
                const double l = lb(m, 0, p[i]), r = lb(m, 1, p[i]);
                const bool wl = l < best[i] * slack, wr = r < best[i] * slack;
                wantL += wl; wantR += wr;
                if (wl || wr) { if (l < r) prefL++; else prefR++; }
                sumL += l; sumR += r;
            }
            bool leftFirst = order_mode == 0 ? (prefL >= prefR) : (sumL <= sumR);
            const int fb = leftFirst ? b : m, fe = leftFirst ? m : e, sb = leftFirst ? m : b, se = leftFirst ? e : m;
            const int w1 = leftFirst ? wantL : wantR, w2 = leftFirst ? wantR : wantL;
            if (w1) {
                if (w2) { st.push_back({sb, se, m, leftFirst ? 1 : 0, depth + 1}); c.maxstack = std::max<long long>(c.maxstack, (long long)st.size()); }
                b = fb; e = fe; depth++; have = true;
            } else if (w2) { b = sb; e = se; depth++; have = true; }
        }
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
    const int res = atoi(argv[3]), stride = argc > 4 ? atoi(argv[4]) : 5;
    const uint64_t nV = vb.size() / 24, nT = fb.size() / 12;
    const double* V = (const double*)vb.data(); const uint32_t* F = (const uint32_t*)fb.data();
    HostBvh H; const char* err = nullptr;
    if (!build_host_bvh(V, nV, F, nT, H, &err)) { fprintf(stderr, "build: %s\n", err); return 1; }
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint64_t i = 0; i < nV; i++) for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], V[3 * i + d]); mx[d] = std::max(mx[d], V[3 * i + d]); }
    const double diag = std::sqrt((mx[0] - mn[0]) * (mx[0] - mn[0]) + (mx[1] - mn[1]) * (mx[1] - mn[1]) + (mx[2] - mn[2]) * (mx[2] - mn[2]));
    for (int d = 0; d < 3; d++) { mn[d] -= 1e-3 * diag; mx[d] += 1e-3 * diag; }
    double cell[3]; for (int d = 0; d < 3; d++) cell[d] = (mx[d] - mn[d]) / res;
    Packet M(H);
    const int nb[3] = {(res + 1 + 3) / 4, (res + 1 + 3) / 4, (res + 1 + 1) / 2};
    for (int mode = 0; mode < 2; mode++) {
        PCounts tot;
#pragma omp parallel
        {
            PCounts c;
#pragma omp for schedule(dynamic, 4) nowait
            for (int bz = 0; bz < nb[2]; bz++)
                for (int by = 0; by < nb[1]; by++)
                    for (int bx = 0; bx < nb[0]; bx++) {
                        if ((bx + 3 * by + 7 * bz) % stride) continue;
                        P3 p[32]; double best[32]; int n = 0;
                        for (int k = 0; k < 2; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
                            const int gi = 4 * bx + i, gj = 4 * by + j, gk = 2 * bz + k;
                            if (gi > res || gj > res || gk > res) continue;
                            p[n++] = {mn[0] + cell[0] * gi, mn[1] + cell[1] * gj, mn[2] + cell[2] * gk};
                        }
                        if (!n) continue;
                        c.bricks++; c.queries += n;
                        M.brick(p, n, best, c, 1.0 + 1e-5, mode);
                    }
#pragma omp critical
            {
                tot.nodes += c.nodes; tot.nodes_even += c.nodes_even; tot.nodes_odd += c.nodes_odd; tot.filters += c.filters; tot.exacts += c.exacts; tot.exact_lanes += c.exact_lanes; tot.pops += c.pops; tot.pops_hit += c.pops_hit;
                tot.bricks += c.bricks; tot.queries += c.queries; tot.maxstack = std::max(tot.maxstack, c.maxstack);
            }
        }
        const double B = (double)tot.bricks;
        printf("order %s: per brick: node steps %.1f  leaf filter steps %.1f  exact blocks %.1f (%.1f lanes each)  pops %.1f (%.1f hit)  max stack %lld\n",
               mode == 0 ? "majority" : "mean-lb", tot.nodes / B, tot.filters / B, tot.exacts / B, tot.exact_lanes / (double)std::max(1LL, tot.exacts), tot.pops / B, tot.pops_hit / B, tot.maxstack);
        printf("   node steps at even / odd depth: %.1f / %.1f per brick (a 4-wide node = an even-depth node with its two children inlined: %.1f steps instead of %.1f)\n", tot.nodes_even / B, tot.nodes_odd / B, tot.nodes_even / B, tot.nodes / B);
        const double instr = tot.nodes / B * 60 + tot.filters / B * 90 + tot.exacts / B * 270 + tot.pops / B * 40;
        printf("   modelled warp instructions per brick %.0f = %.0f per query (node 60, filter 90, exact 270, pop 40)\n", instr, instr * B / tot.queries);
    }
    return 0;
}
