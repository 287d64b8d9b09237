// CPU model of K1's WARP-SYNCHRONOUS schedule (k1_sdf.cu: nearest_triangle): 32 lanes = one brick of grid nodes, each lane a small
// state machine (NODE / LEAF / POP / DONE); per iteration the warp runs ONE phase, chosen by a weighted vote, and only the lanes in
// that state advance.  Counts iterations per phase and the lanes active in them, for different brick shapes and voting rules --
// an experiment bench to find schedules worth trying on the GPU, not product code (distances: plain closest-point routine).
// build: g++ -O2 -fopenmp -ffp-contract=off -I discregrid_b200/csrc tools/warp_model.cpp discregrid_b200/csrc/bvh_build.cpp discregrid_b200/csrc/sort_replay.cpp -o warp_model -lpthread
// usage: warp_model mesh.V mesh.F resolution [brick_stride]
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "bvh_build.h"

using namespace dgb;

struct P3 { double x, y, z; };
static inline P3 sub(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

static double tri_d2(const LeafRecord& L, P3 p)          // Ericson, Real-Time Collision Detection 5.1.5
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

enum State { NODE = 0, LEAF = 1, POP = 2, DONE = 3 };

struct Item { int b, e; double d; int m; int which; };

struct Lane {
    P3 p; double best; int b, e; State st; std::vector<Item> stack;
    void start(P3 q, int T) { p = q; best = DBL_MAX; b = 0; e = T; st = (T == 1) ? LEAF : NODE; stack.clear(); }
};

struct Policy {
    const char* name;
    int w_node, w_leaf, w_pop;      // vote weights (k1_sdf.h: 2 / 3 / 4)
    int pop_tries;                  // deferred entries re-tested per POP iteration (4)
    bool merge_node_pop;            // NODE and POP lanes advance in the same iteration (a fused phase)
    int leaf_patience;              // run LEAF only when it wins the vote OR has waited this many iterations (0 = plain vote)
    bool chain_pop_node = false;    // a lane whose POP finds an internal node takes that node step in the same iteration
};

struct Tally { long long it_chain = 0, act_chain = 0; long long lane_steps = 0, max_lane_steps = 0; long long it_node = 0, it_leaf = 0, it_pop = 0, it_fused = 0, act_node = 0, act_leaf = 0, act_pop = 0, act_fused = 0, bricks = 0, queries = 0; };

struct Sim {
    const HostBvh& H; int T;
    explicit Sim(const HostBvh& h) : H(h), T((int)h.n_triangles) {}
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
    void node_step(Lane& L) const
    {
        const int b = L.b, e = L.e, m = (b + e) >> 1;
        const SpherePair& S = H.spheres[m];
        const double dl = sph(S.lc, S.lr, L.p), dr = sph(S.rc, S.rr, L.p);
        const bool lf = dl < dr;
        const double d1 = lf ? dl : dr, d2 = lf ? dr : dl;
        bool go1 = d1 < L.best, def2 = d2 < L.best, go2 = false;
        if (go1) {
            const double b1 = boxd(m, lf ? 0 : 1, L.p), b2 = boxd(m, lf ? 1 : 0, L.p);
            if (b2 > L.best) def2 = false;
            if (b1 > L.best) { go1 = false; go2 = def2; }          // hopeless first child: turn to the second right away
        }
        const int fb = lf ? b : m, fe = lf ? m : e, sb = lf ? m : b, se = lf ? e : m;
        if (go1) { if (def2) L.stack.push_back({sb, se, d2, m, lf ? 1 : 0}); L.b = fb; L.e = fe; L.st = (fe - fb == 1) ? LEAF : NODE; }
        else if (go2) { L.b = sb; L.e = se; L.st = (se - sb == 1) ? LEAF : NODE; }
        else L.st = POP;
    }
    void leaf_step(Lane& L) const
    {
        const double d2 = tri_d2(H.leaves[L.b], L.p);
        if (d2 < L.best * L.best) L.best = std::sqrt(d2);
        L.st = POP;
    }
    void pop_step(Lane& L, int tries) const
    {
        for (int a = 0; a < tries; a++) {
            if (L.stack.empty()) { L.st = DONE; return; }
            const Item it = L.stack.back(); L.stack.pop_back();
            if (!(it.d < L.best)) continue;
            if (boxd(it.m, it.which, L.p) > L.best) continue;
            L.b = it.b; L.e = it.e; L.st = (it.e - it.b == 1) ? LEAF : NODE;
            return;
        }
    }
    void run_brick(Lane* lanes, int n, const Policy& P, Tally& t) const
    {
        int waited = 0;
        {   // how long each lane's own walk is (node steps + leaf tests + pop iterations), independent of the schedule
            long long mx = 0;
            for (int i = 0; i < n; i++) {
                Lane L = lanes[i]; long long steps = 0;
                while (L.st != DONE) { if (L.st == NODE) node_step(L); else if (L.st == LEAF) leaf_step(L); else pop_step(L, P.pop_tries); steps++; }
                t.lane_steps += steps; mx = std::max(mx, steps);
            }
            t.max_lane_steps += mx;
        }
        for (;;) {
            int c[4] = {0, 0, 0, 0};
            for (int i = 0; i < n; i++) c[lanes[i].st]++;
            if (c[DONE] == n) break;
            if (P.merge_node_pop) {
                const int wf = P.w_node * (c[NODE] + c[POP]), wl = P.w_leaf * c[LEAF];
                const bool leaf = (c[NODE] + c[POP] == 0) || (wl > wf) || (P.leaf_patience && c[LEAF] && waited >= P.leaf_patience);
                if (leaf) { t.it_leaf++; t.act_leaf += c[LEAF]; waited = 0; for (int i = 0; i < n; i++) if (lanes[i].st == LEAF) leaf_step(lanes[i]); }
                else {
                    t.it_fused++; t.act_fused += c[NODE] + c[POP]; if (c[LEAF]) waited++;
                    for (int i = 0; i < n; i++) { if (lanes[i].st == NODE) node_step(lanes[i]); else if (lanes[i].st == POP) pop_step(lanes[i], P.pop_tries); }
                }
                continue;
            }
            const int wn = P.w_node * c[NODE], wl = P.w_leaf * c[LEAF], wp = P.w_pop * c[POP];
            const int wmax = std::max(wn, std::max(wl, wp));
            // the kernel's order of preference on equal weights: POP, NODE, LEAF
            if (wp == wmax && !(P.leaf_patience && c[LEAF] && waited >= P.leaf_patience)) {
                t.it_pop++; t.act_pop += c[POP]; if (c[LEAF]) waited++;
                int chained = 0;
                for (int i = 0; i < n; i++) if (lanes[i].st == POP) { pop_step(lanes[i], P.pop_tries); if (P.chain_pop_node && lanes[i].st == NODE) { node_step(lanes[i]); chained++; } }
                if (chained) { t.it_chain++; t.act_chain += chained; }
            } else if (wn == wmax && !(P.leaf_patience && c[LEAF] && waited >= P.leaf_patience)) {
                t.it_node++; t.act_node += c[NODE]; if (c[LEAF]) waited++;
                for (int i = 0; i < n; i++) if (lanes[i].st == NODE) node_step(lanes[i]);
            } else {
                t.it_leaf++; t.act_leaf += c[LEAF]; waited = 0;
                for (int i = 0; i < n; i++) if (lanes[i].st == LEAF) leaf_step(lanes[i]);
            }
        }
        t.bricks++; t.queries += n;
    }
};

// K queries per lane: the warp owns K bricks; in the chosen phase every lane advances ONE of its K queries that is in that state.
// More lanes have something to do per iteration, at the price of K times the per-lane state.
static void run_multi(const Sim& S, Lane* q /*[K][32]*/, int K, const int* n, const Policy& P, Tally& t)
{
    for (;;) {
        int c[4] = {0, 0, 0, 0}, total = 0;
        for (int k = 0; k < K; k++) for (int i = 0; i < n[k]; i++) { c[q[k * 32 + i].st]++; total++; }
        if (c[DONE] == total) break;
        // lanes that could act in each phase (what the vote should weigh: a lane advances one query per iteration)
        int can[3] = {0, 0, 0};
        for (int i = 0; i < 32; i++) for (int ph = 0; ph < 3; ph++) { bool any = false; for (int k = 0; k < K; k++) if (i < n[k] && q[k * 32 + i].st == ph) any = true; can[ph] += any; }
        const int wn = P.w_node * can[NODE], wl = P.w_leaf * can[LEAF], wp = P.w_pop * can[POP];
        const int wmax = std::max(wn, std::max(wl, wp));
        const int ph = (wp == wmax) ? POP : ((wn == wmax) ? NODE : LEAF);
        for (int i = 0; i < 32; i++)
            for (int k = 0; k < K; k++) {
                if (i >= n[k]) continue;
                Lane& L = q[k * 32 + i];
                if (L.st != ph) continue;
                if (ph == NODE) S.node_step(L); else if (ph == LEAF) S.leaf_step(L); else S.pop_step(L, P.pop_tries);
                break;
            }
        if (ph == NODE) { t.it_node++; t.act_node += can[NODE]; } else if (ph == LEAF) { t.it_leaf++; t.act_leaf += can[LEAF]; } else { t.it_pop++; t.act_pop += can[POP]; }
    }
    for (int k = 0; k < K; k++) { t.bricks++; t.queries += n[k]; }
}

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
    const int res = atoi(argv[3]), stride = argc > 4 ? atoi(argv[4]) : 7;
    const uint64_t nV = vb.size() / 24, nT = fb.size() / 12;
    const double* V = (const double*)vb.data(); const uint32_t* F = (const uint32_t*)fb.data();
    HostBvh H; const char* err = nullptr;
    if (!build_host_bvh(V, nV, F, nT, H, &err)) { fprintf(stderr, "build: %s\n", err); return 1; }
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint64_t i = 0; i < nV; i++) for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], V[3 * i + d]); mx[d] = std::max(mx[d], V[3 * i + d]); }
    const double diag = std::sqrt((mx[0] - mn[0]) * (mx[0] - mn[0]) + (mx[1] - mn[1]) * (mx[1] - mn[1]) + (mx[2] - mn[2]) * (mx[2] - mn[2]));
    for (int d = 0; d < 3; d++) { mn[d] -= 1e-3 * diag; mx[d] += 1e-3 * diag; }
    double cell[3]; for (int d = 0; d < 3; d++) cell[d] = (mx[d] - mn[d]) / res;
    Sim S(H);

    struct Shape { int f, m, s; };
    const Shape shapes[] = {{4, 4, 2}, {8, 2, 2}, {2, 4, 4}, {4, 2, 4}, {2, 2, 8}, {1, 4, 8}, {2, 1, 16}, {8, 4, 1}, {32, 1, 1}};
    const Policy policies[] = {
        {"kernel: vote 2/3/4, 4 pop tries", 2, 3, 4, 4, false, 0},
        {"vote 1/1/1", 1, 1, 1, 4, false, 0},
        {"vote 2/6/4 (leaf-eager)", 2, 6, 4, 4, false, 0},
        {"vote 2/1/4 (leaf-lazy)", 2, 1, 4, 4, false, 0},
        {"vote 2/3/4, 1 pop try", 2, 3, 4, 1, false, 0},
        {"vote 2/3/4, 16 pop tries", 2, 3, 4, 16, false, 0},
        {"fused node+pop, leaf when heavier (2:3)", 2, 3, 0, 4, true, 0},
        {"fused node+pop, leaf weight 1 (lazy)", 2, 1, 0, 4, true, 0},
        {"fused node+pop, leaf lazy, patience 8", 2, 1, 0, 4, true, 8},
        {"vote 2/3/4 + pop chains into the node step", 2, 3, 4, 4, false, 0, true},
        {"vote 2/3/3 + pop chains into the node step", 2, 3, 3, 4, false, 0, true},
    };
    // instruction cost per iteration of each phase, from the SASS/ncu breakdown in profiles/README.md (node 70; the others solved from
    // the issued-instruction shares 21 : 40 : 14 : 18 of node : leaf : pop : loop head with the kernel policy's iteration counts below)
    printf("triangles %llu, %d^3 vertex lattice, every %dth brick\n", (unsigned long long)nT, res, stride);
    double cost_leaf = 0, cost_pop = 0, cost_head = 0;
    for (const Shape& sh : shapes)
        for (const Policy& P : policies) {
            if (&sh != &shapes[0] && &P != &policies[0]) continue;                          // other shapes: kernel policy only
            const int nb[3] = {(res + sh.f) / sh.f, (res + sh.m) / sh.m, (res + sh.s) / sh.s};
            Tally tot;
#pragma omp parallel
            {
                Tally t; std::vector<Lane> lanes(32);
#pragma omp for schedule(dynamic, 4) nowait
                for (int bz = 0; bz < nb[2]; bz++)
                    for (int by = 0; by < nb[1]; by++)
                        for (int bx = 0; bx < nb[0]; bx++) {
                            if ((bx + 3 * by + 7 * bz) % stride) continue;
                            int n = 0;
                            for (int k = 0; k < sh.s; k++) for (int j = 0; j < sh.m; j++) for (int i = 0; i < sh.f; i++) {
                                const int gi = sh.f * bx + i, gj = sh.m * by + j, gk = sh.s * bz + k;
                                if (gi > res || gj > res || gk > res) continue;
                                lanes[n++].start({mn[0] + cell[0] * gi, mn[1] + cell[1] * gj, mn[2] + cell[2] * gk}, S.T);
                            }
                            if (n) S.run_brick(lanes.data(), n, P, t);
                        }
#pragma omp critical
                { tot.it_node += t.it_node; tot.it_leaf += t.it_leaf; tot.it_pop += t.it_pop; tot.it_fused += t.it_fused; tot.act_node += t.act_node; tot.act_leaf += t.act_leaf;
                  tot.act_pop += t.act_pop; tot.act_fused += t.act_fused; tot.bricks += t.bricks; tot.queries += t.queries; tot.lane_steps += t.lane_steps; tot.max_lane_steps += t.max_lane_steps; tot.it_chain += t.it_chain; tot.act_chain += t.act_chain; }
            }
            const double B = (double)tot.bricks;
            const long long its = tot.it_node + tot.it_leaf + tot.it_pop + tot.it_fused;
            if (&sh == &shapes[0] && &P == &policies[0]) {
                const double node_instr = 70.0 * tot.it_node;
                cost_leaf = node_instr * (40.0 / 21.0) / tot.it_leaf; cost_pop = node_instr * (14.0 / 21.0) / tot.it_pop; cost_head = node_instr * (18.0 / 21.0) / its;
                printf("calibrated instruction cost per iteration: node 70, leaf %.0f, pop %.0f, loop head %.0f\n", cost_leaf, cost_pop, cost_head);
                printf("steps of a lane's own walk: mean %.1f, mean over bricks of the longest lane %.1f; warp iterations per brick %.1f\n",
                       (double)tot.lane_steps / tot.queries, (double)tot.max_lane_steps / B, (double)its / B);
            }
            const double fused_cost = 70.0 + cost_pop * 0.8;           // a fused phase issues both bodies (predicated), sharing some setup
            const double instr = 70.0 * tot.it_node + cost_leaf * tot.it_leaf + cost_pop * tot.it_pop + fused_cost * tot.it_fused + cost_head * its + 70.0 * tot.it_chain;
            printf("brick %2dx%dx%d  %-44s it/brick node %6.1f leaf %6.1f pop %6.1f fused %6.1f | lanes node %4.1f leaf %4.1f pop %4.1f fused %4.1f | instr/query %7.0f\n",
                   sh.f, sh.m, sh.s, P.name, tot.it_node / B, tot.it_leaf / B, tot.it_pop / B, tot.it_fused / B,
                   tot.it_node ? (double)tot.act_node / tot.it_node : 0.0, tot.it_leaf ? (double)tot.act_leaf / tot.it_leaf : 0.0,
                   tot.it_pop ? (double)tot.act_pop / tot.it_pop : 0.0, tot.it_fused ? (double)tot.act_fused / tot.it_fused : 0.0, instr / tot.queries);
        }
    // K bricks per warp (neighbours along the fast axis), kernel vote
    for (int K : {1, 2, 4}) {
        const Shape sh = shapes[0]; const Policy& P = policies[0];
        const int nb[3] = {(res + sh.f) / sh.f, (res + sh.m) / sh.m, (res + sh.s) / sh.s};
        Tally tot;
#pragma omp parallel
        {
            Tally t; std::vector<Lane> lanes(32 * K);
#pragma omp for schedule(dynamic, 4) nowait
            for (int bz = 0; bz < nb[2]; bz++)
                for (int by = 0; by < nb[1]; by++)
                    for (int bx0 = 0; bx0 < nb[0]; bx0 += K) {
                        if ((bx0 / K + 3 * by + 7 * bz) % stride) continue;
                        int n[8] = {0};
                        for (int kk = 0; kk < K; kk++) {
                            const int bx = bx0 + kk;
                            for (int k = 0; k < sh.s; k++) for (int j = 0; j < sh.m; j++) for (int i = 0; i < sh.f; i++) {
                                const int gi = sh.f * bx + i, gj = sh.m * by + j, gk = sh.s * bz + k;
                                if (gi > res || gj > res || gk > res) continue;
                                lanes[kk * 32 + n[kk]++].start({mn[0] + cell[0] * gi, mn[1] + cell[1] * gj, mn[2] + cell[2] * gk}, S.T);
                            }
                        }
                        run_multi(S, lanes.data(), K, n, P, t);
                    }
#pragma omp critical
            { tot.it_node += t.it_node; tot.it_leaf += t.it_leaf; tot.it_pop += t.it_pop; tot.act_node += t.act_node; tot.act_leaf += t.act_leaf; tot.act_pop += t.act_pop;
              tot.bricks += t.bricks; tot.queries += t.queries; }
        }
        const long long its = tot.it_node + tot.it_leaf + tot.it_pop;
        const double instr = 70.0 * tot.it_node + cost_leaf * tot.it_leaf + cost_pop * tot.it_pop + cost_head * its;
        printf("%d queries per lane (kernel vote): lanes node %4.1f leaf %4.1f pop %4.1f | instr/query %7.0f (+ per-iteration selection overhead not counted)\n", K,
               (double)tot.act_node / tot.it_node, (double)tot.act_leaf / tot.it_leaf, (double)tot.act_pop / tot.it_pop, instr / tot.queries);
    }
    return 0;
}
