// Launch-tail model built on the warp-schedule model (tools/warp_model.cpp): list scheduling of the blocks' modelled costs on the resident
// block slots of a B200, in launch order vs heaviest-class-first (K1_COST_ORDER) vs true cost order.
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

static std::vector<unsigned char> slurp(const char* path)
{
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) exit(1); fclose(f); return v;
}

// list scheduling of `cost` (in launch order `ord`) on P equal slots: returns the makespan
static double makespan(const std::vector<double>& cost, const std::vector<int>& ord, int P)
{
    std::vector<double> heap(P, 0.0);                              // min-heap of slot finish times
    std::make_heap(heap.begin(), heap.end(), std::greater<double>());
    double end = 0;
    for (int id : ord) {
        std::pop_heap(heap.begin(), heap.end(), std::greater<double>());
        heap.back() += cost[id]; end = std::max(end, heap.back());
        std::push_heap(heap.begin(), heap.end(), std::greater<double>());
    }
    return end;
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s mesh.V mesh.F resolution [parts=1]\n", argv[0]); return 1; }
    auto vb = slurp(argv[1]), fb = slurp(argv[2]);
    const int res = atoi(argv[3]), parts = argc > 4 ? atoi(argv[4]) : 1;
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
    const Policy P = {"kernel", 2, 3, 4, 4, false, 0};
    // blocks of the vertex array: 8 x 4 x 2 nodes (two 4 x 4 x 2 bricks side by side), plane pairs dealt to `parts` (part 0 is simulated)
    const int nbx = (res + 8) / 8, nby = (res + 4) / 4, nbz = (res + 2) / 2;
    std::vector<double> cost; std::vector<double> cdist;
    std::vector<int> bxs, bys, bzs;
    for (int bz = 0; bz < nbz; bz++) { if (bz % parts) continue; for (int by = 0; by < nby; by++) for (int bx = 0; bx < nbx; bx++) { bxs.push_back(bx); bys.push_back(by); bzs.push_back(bz); } }
    const int nblk = (int)bxs.size();
    cost.assign(nblk, 0); cdist.assign(nblk, 0);
#pragma omp parallel
    {
        std::vector<Lane> lanes(32);
#pragma omp for schedule(dynamic, 16)
        for (int b = 0; b < nblk; b++) {
            double worst = 0;
            for (int half = 0; half < 2; half++) {
                int n = 0;
                for (int k = 0; k < 2; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
                    const int gi = 8 * bxs[b] + 4 * half + i, gj = 4 * bys[b] + j, gk = 2 * bzs[b] + k;
                    if (gi > res || gj > res || gk > res) continue;
                    lanes[n++].start({mn[0] + cell[0] * gi, mn[1] + cell[1] * gj, mn[2] + cell[2] * gk}, S.T);
                }
                if (!n) continue;
                Tally t; S.run_brick(lanes.data(), n, P, t);
                const double instr = 70.0 * t.it_node + 253.0 * t.it_leaf + 65.0 * t.it_pop + 27.0 * (t.it_node + t.it_leaf + t.it_pop);
                worst = std::max(worst, instr);
            }
            cost[b] = worst;                                      // a block lives as long as its slower warp
            // lattice proxy: unsigned distance at the centre of the 16^3 lattice cell that holds the block's centre
            const double c[3] = {mn[0] + cell[0] * (8 * bxs[b] + 4), mn[1] + cell[1] * (4 * bys[b] + 2), mn[2] + cell[2] * (2 * bzs[b] + 1)};
            double q[3];
            for (int d = 0; d < 3; d++) { int li = (int)((c[d] - mn[d]) / (mx[d] - mn[d]) * 16); li = std::min(15, std::max(0, li)); q[d] = mn[d] + (mx[d] - mn[d]) * ((li + 0.5) / 16); }
            Lane L; L.start({q[0], q[1], q[2]}, S.T);
            while (L.st != DONE) { if (L.st == NODE) S.node_step(L); else if (L.st == LEAF) S.leaf_step(L); else S.pop_step(L, 4); }
            cdist[b] = L.best;
        }
    }
    double total = 0, cmax = 0, dmax = 0;
    for (int b = 0; b < nblk; b++) { total += cost[b]; cmax = std::max(cmax, cost[b]); dmax = std::max(dmax, cdist[b]); }
    const int slots = 148 * 16;                                   // resident 64-thread blocks on a B200 at 64 registers
    std::vector<int> in_order(nblk), by_class(nblk), ideal(nblk);
    for (int b = 0; b < nblk; b++) in_order[b] = by_class[b] = ideal[b] = b;
    std::stable_sort(by_class.begin(), by_class.end(), [&](int a, int b) { return std::min(31, (int)(cdist[a] / dmax * 32)) > std::min(31, (int)(cdist[b] / dmax * 32)); });
    std::stable_sort(ideal.begin(), ideal.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    const double lb = total / slots;
    printf("%d blocks (part 0 of %d), cost per block: mean %.0f max %.0f instr; %d slots; perfect balance %.0f\n", nblk, parts, total / nblk, cmax, slots, lb);
    printf("makespan / perfect balance: launch order %.3f   heaviest class first (16^3 lattice, 32 classes) %.3f   true cost order %.3f\n",
           makespan(cost, in_order, slots) / lb, makespan(cost, by_class, slots) / lb, makespan(cost, ideal, slots) / lb);
    return 0;
}
