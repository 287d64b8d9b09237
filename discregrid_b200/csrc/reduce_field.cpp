// See reduce_field.h.  Host C++ (threads), compiled with -ffp-contract=off like every translation unit of the library.
#include "reduce_field.h"
#include "sort_replay.h"

#include <algorithm>
namespace dgb { unsigned host_threads(); }   // host_threads.cpp: affinity mask capped by the cgroup CPU quota
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace dgb {
namespace {

unsigned n_threads()
{
    unsigned nt = dgb::host_threads();
    if (nt == 0) nt = 1;
    return nt > 32 ? 32 : nt;
}

// static partition of [0, n) over the host threads; fn(thread, begin, end)
template <class Fn>
void parallel_chunks(uint64_t n, unsigned nt, Fn fn)
{
    if (n < 65536 || nt <= 1) { fn(0u, (uint64_t)0, n); return; }
    std::vector<std::thread> th;
    const uint64_t per = (n + nt - 1) / nt;
    for (unsigned k = 0; k < nt; k++) {
        const uint64_t b = std::min(n, k * per), e = std::min(n, b + per);
        th.emplace_back([=, &fn]() { fn(k, b, e); });
    }
    for (auto& t : th) t.join();
}

double ms_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// morton_lut (z_sort_table.hpp:119-134): the first stage is shifted by 48 and then by 24 bits, so only the LOW 16 bits of each
// coordinate reach the 64-bit key (x at bit 0, y at bit 1, z at bit 2)
inline uint64_t spread16(uint64_t v)
{
    v &= 0xffffull;
    v = (v | v << 16) & 0x0000ff0000ffull;
    v = (v | v << 8) & 0x00f00f00f00full;
    v = (v | v << 4) & 0x0c30c30c30c3ull;
    v = (v | v << 2) & 0x249249249249ull;
    return v;
}

// Sorts `a` by key with a multithreaded sample sort.  Only called when the caller goes on to verify that all keys are distinct
// (otherwise the order among equal keys would be ours, not the reference's).
void sample_sort(std::vector<KeyPos>& a, unsigned nt)
{
    const uint64_t m = a.size();
    auto less = [](const KeyPos& x, const KeyPos& y) { return x.key < y.key; };
    if (m < (1u << 16) || nt <= 1) { std::sort(a.begin(), a.end(), less); return; }
    const unsigned nb = nt * 8;                                   // buckets
    std::vector<uint64_t> sample;
    const uint64_t n_sample = (uint64_t)nb * 64;
    sample.reserve(n_sample);
    for (uint64_t i = 0; i < n_sample; i++) sample.push_back(a[(i * m) / n_sample].key);
    std::sort(sample.begin(), sample.end());
    std::vector<uint64_t> split(nb - 1);
    for (unsigned b = 1; b < nb; b++) split[b - 1] = sample[(uint64_t)b * 64];
    auto bucket_of = [&](uint64_t key) { return (unsigned)(std::upper_bound(split.begin(), split.end(), key) - split.begin()); };
    std::vector<std::vector<uint64_t>> cnt(nt, std::vector<uint64_t>(nb, 0));
    parallel_chunks(m, nt, [&](unsigned t, uint64_t b, uint64_t e) { for (uint64_t i = b; i < e; i++) cnt[t][bucket_of(a[i].key)]++; });
    std::vector<uint64_t> start(nb + 1, 0);
    std::vector<std::vector<uint64_t>> off(nt, std::vector<uint64_t>(nb, 0));
    {
        uint64_t run = 0;
        for (unsigned b = 0; b < nb; b++) {
            start[b] = run;
            for (unsigned t = 0; t < nt; t++) { off[t][b] = run; run += cnt[t][b]; }
        }
        start[nb] = run;
    }
    std::vector<KeyPos> out(m);
    parallel_chunks(m, nt, [&](unsigned t, uint64_t b, uint64_t e) { for (uint64_t i = b; i < e; i++) out[off[t][bucket_of(a[i].key)]++] = a[i]; });
    std::atomic<unsigned> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&]() {
            for (;;) {
                const unsigned b = next.fetch_add(1);
                if (b >= nb) break;
                std::sort(out.begin() + start[b], out.begin() + start[b + 1], less);
            }
        });
    for (auto& t : th) t.join();
    a.swap(out);
}

}  // namespace

uint64_t reduce_field_morton_key(const GridDev& g, uint32_t l)
{
    double x[3];
    node_position(g, l, x[0], x[1], x[2]);                                                    // :1113
    const double inv = 4.0 * std::min(std::min(g.inv[0], g.inv[1]), g.inv[2]);                // :1114
    uint32_t p[3];
    for (int k = 0; k < 3; k++) {
        const int key = (x[k] >= 0.0) ? static_cast<int>(inv * x[k]) : static_cast<int>(inv * x[k]) - 1;                 // :589-592
        p[k] = static_cast<uint32_t>(static_cast<int64_t>(key) - (std::numeric_limits<int>::lowest() + 1));              // :595-598
    }
    return spread16(p[0]) | (spread16(p[1]) << 1) | (spread16(p[2]) << 2);
}

void replay_std_sort(KeyPos* a, uint64_t n, unsigned nt)
{
    replay_sort(a, n, [](const KeyPos& x, const KeyPos& y) { return x.key < y.key; }, nt);
}

bool replay_std_sort_matches() { return sort_replay_matches_std_sort(); }
uint64_t replay_std_sort_heap_fallbacks() { return g_replay_heap_sorts.load(); }

// The reference compacts the surviving nodes by "swap the dead node with the last live slot, walking from the back" (:1135-1156):
// replayed on a permutation array.  Returns the number of survivors m; perm[0..m) = original ids of the survivors in that order.
uint64_t reduce_swap_walk(const uint8_t* used, uint64_t n_nodes, std::vector<uint32_t>& perm)
{
    perm.resize(n_nodes);
    std::iota(perm.begin(), perm.end(), 0u);
    uint32_t last = (uint32_t)(n_nodes - 1);
    for (int64_t i = (int64_t)n_nodes - 1; i >= 0; --i) {
        if (used[i]) continue;                                   // position i still holds node i when the walk reaches it
        std::swap(perm[i], perm[last]);
        last--;                                                  // wraps to 0xffffffff when nothing survives: last + 1 == 0 below
    }
    return (uint64_t)(uint32_t)(last + 1u);
}

// Order of the survivors: positions 0..m-1 sorted by the Morton key of the node sitting there (:1160-1165), equal keys in the order
// libstdc++'s std::sort leaves them (see reduce_field.h).  kp[i] = {key of the node at position i, i}; kp is consumed.
void reduce_order_survivors(std::vector<KeyPos>& kp, bool force_std_sort, std::vector<uint32_t>& order, int& tie_path)
{
    const uint64_t m = kp.size();
    const unsigned nt = n_threads();
    order.resize(m);
    tie_path = 0;
    auto has_ties = [&](const std::vector<KeyPos>& sorted) {
        std::atomic<int> dup{0};
        parallel_chunks(m, nt, [&](unsigned, uint64_t b, uint64_t e) {
            for (uint64_t i = std::max<uint64_t>(b, 1); i < e; i++) if (sorted[i].key == sorted[i - 1].key) { dup.store(1); break; }
        });
        return dup.load() != 0;
    };
    if (replay_std_sort_matches() && !force_std_sort) {
        // the reference's own sort on the reference's own input sequence, replayed on all threads (reduce_field.h): right with and
        // without ties
        replay_std_sort(kp.data(), m, nt);
        tie_path = has_ties(kp) ? 1 : 0;
        parallel_chunks(m, nt, [&](unsigned, uint64_t b, uint64_t e) { for (uint64_t i = b; i < e; i++) order[i] = kp[i].pos; });
    } else {
        // another standard library (or force_std_sort): any correct sort will do while the keys are distinct; with ties only the
        // reference's very call reproduces its order
        bool ties = force_std_sort;
        if (!ties) {
            std::vector<KeyPos> sorted(kp);
            sample_sort(sorted, nt);
            ties = has_ties(sorted);
            if (!ties) parallel_chunks(m, nt, [&](unsigned, uint64_t b, uint64_t e) { for (uint64_t i = b; i < e; i++) order[i] = sorted[i].pos; });
        }
        if (ties) {
            std::iota(order.begin(), order.end(), 0u);
            std::sort(order.begin(), order.end(), [&](unsigned int i, unsigned int j) { return kp[i].key < kp[j].key; });
            tie_path = 1;
        }
    }
}

bool reduce_field_host(const GridDev& g, double* nodes, uint64_t n_nodes, const uint8_t* keep_node, uint32_t* cells, uint64_t n_cells_in,
                       uint32_t* cell_map, uint64_t n_cells_grid, bool force_std_sort, ReduceStats& st, const char** err)
{
    const uint32_t NONE = std::numeric_limits<uint32_t>::max();
    if (n_nodes == 0 || n_nodes > 0x7fffffffull) { *err = "node count must be in [1, 2^31) (the reference indexes nodes with int/unsigned)"; return false; }
    if (n_cells_in > n_cells_grid) { *err = "more cells than the grid has"; return false; }
    const unsigned nt = n_threads();
    auto t0 = std::chrono::steady_clock::now();

    // ---- 1. surviving cells (:1076-1098).  Every cell entry must name a node.
    std::vector<uint8_t> cell_keep(n_cells_in);
    std::atomic<int> bad{0};
    parallel_chunks(n_cells_in, nt, [&](unsigned, uint64_t b, uint64_t e) {
        for (uint64_t c = b; c < e; c++) {
            const uint32_t* row = cells + 32 * c;
            uint8_t any = 0;
            for (int j = 0; j < 32; j++) {
                if (row[j] >= n_nodes) { bad.store(1); any = 0; break; }
                any |= keep_node[row[j]];
            }
            cell_keep[c] = any ? 1 : 0;
        }
    });
    if (bad.load()) { *err = "a cell refers to a node id >= n_nodes"; return false; }
    for (uint64_t i = 0; i < n_cells_grid; i++) cell_map[i] = (uint32_t)i;                    // resize + iota (:1076-1078)
    uint64_t kept_cells = 0;
    for (uint64_t c = 0; c < n_cells_in; c++) {                                               // serial: a prefix sum and a forward move
        if (cell_keep[c]) {
            if (kept_cells != c) std::memmove(cells + 32 * kept_cells, cells + 32 * c, 32 * sizeof(uint32_t));
            cell_map[c] = (uint32_t)kept_cells++;
        } else {
            cell_map[c] = NONE;
        }
    }
    st.cells_out = kept_cells;
    st.ms_cells = ms_since(t0); t0 = std::chrono::steady_clock::now();

    // ---- 2. surviving nodes = nodes of surviving cells (:1117-1134), compacted by the reference's swap walk (:1135-1156)
    std::vector<uint8_t> used(n_nodes, 0);
    parallel_chunks(kept_cells * 32, nt, [&](unsigned, uint64_t b, uint64_t e) {
        for (uint64_t k = b; k < e; k++) __atomic_store_n(&used[cells[k]], (uint8_t)1, __ATOMIC_RELAXED);
    });
    std::vector<uint32_t> perm;                                  // perm[position] = original node id
    const uint64_t m = reduce_swap_walk(used.data(), n_nodes, perm);
    st.nodes_out = m;
    st.ms_nodes = ms_since(t0); t0 = std::chrono::steady_clock::now();

    // ---- 3. order of the survivors: positions 0..m-1 sorted by the Morton key of the node sitting there (:1160-1165)
    std::vector<KeyPos> kp(m);
    parallel_chunks(m, nt, [&](unsigned, uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; i++) { kp[i].key = reduce_field_morton_key(g, perm[i]); kp[i].pos = (uint32_t)i; }
    });
    std::vector<uint32_t> order;                                 // sort_pattern
    reduce_order_survivors(kp, force_std_sort, order, st.tie_path);
    st.ms_sort = ms_since(t0); t0 = std::chrono::steady_clock::now();

    // ---- 4. write back: new id of an original node = rank of its position; coefficients in rank order (:1167-1173)
    std::vector<uint32_t> new_id(n_nodes, NONE);
    parallel_chunks(m, nt, [&](unsigned, uint64_t b, uint64_t e) { for (uint64_t r = b; r < e; r++) new_id[perm[order[r]]] = (uint32_t)r; });
    parallel_chunks(kept_cells * 32, nt, [&](unsigned, uint64_t b, uint64_t e) { for (uint64_t k = b; k < e; k++) cells[k] = new_id[cells[k]]; });
    std::vector<double> out(m);
    parallel_chunks(m, nt, [&](unsigned, uint64_t b, uint64_t e) { for (uint64_t r = b; r < e; r++) out[r] = nodes[perm[order[r]]]; });
    parallel_chunks(m, nt, [&](unsigned, uint64_t b, uint64_t e) { if (e > b) std::memcpy(nodes + b, out.data() + b, (e - b) * sizeof(double)); });
    st.ms_write = ms_since(t0);
    return true;
}

}  // namespace dgb
