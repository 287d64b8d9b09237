// std::sort, replayed on several threads.
//
// Two places of the host code must order records exactly as the reference's std::sort calls do, INCLUDING where records with equal
// keys end up (std::sort is not stable): the median split of the sphere tree (TriangleMeshDistance.h:494-499 -- triangles sharing a
// first vertex have equal keys, and the tree decides which of several equidistant triangles a query reports) and the Z-curve order
// of reduceField (cubic_lagrange_discrete_grid.cpp:1160-1165).  libstdc++'s std::sort is introsort: median-of-3 quicksort down to
// 16-element blocks with a 2*floor(log2 n) depth budget (heapsort beyond it), then one insertion-sort pass.  Every quicksort split
// leaves two independent ranges, and the insertion pass never moves a record across a split (left part <= pivot <= right part), so
// the same sequence of comparisons and swaps can run on many threads and ends in the same arrangement.
//
// replay_sort<Rec, Less>(a, n, less, threads) does that.  sort_replay_matches_std_sort() (sort_replay.cpp) checks the scheme
// against std::sort itself once per process on tie-heavy, presorted and adversarial inputs; callers fall back to std::sort when it
// does not hold (a different standard library may sort differently).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace dgb {

extern std::atomic<uint64_t> g_replay_heap_sorts;            // how often a depth budget ran out (tests)
bool sort_replay_matches_std_sort();

namespace sort_replay_detail {

// sift `v` down from `hole` in the max-heap a[0, len), then up again (the usual "move the hole to a leaf, then push" form)
template <class Rec, class Less>
void heap_adjust(Rec* a, int64_t hole, int64_t len, Rec v, Less lt)
{
    const int64_t top = hole;
    int64_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lt(a[child], a[child - 1])) child--;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1];
        hole = child - 1;
    }
    int64_t parent = (hole - 1) / 2;
    while (hole > top && lt(a[parent], v)) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = v;
}

template <class Rec, class Less>
void heap_sort(Rec* a, int64_t len, Less lt)                 // make_heap + sort_heap: the depth-exhausted branch
{
    g_replay_heap_sorts.fetch_add(1, std::memory_order_relaxed);
    if (len < 2) return;
    for (int64_t parent = (len - 2) / 2;; parent--) {
        heap_adjust(a, parent, len, a[parent], lt);
        if (parent == 0) break;
    }
    for (int64_t last = len - 1; last >= 1; last--) {
        const Rec v = a[last];
        a[last] = a[0];
        heap_adjust(a, (int64_t)0, last, v, lt);
    }
}

// median of a[1], a[len/2], a[len-1] to a[0]; Hoare partition of a[1, len) around it; returns the cut
template <class Rec, class Less>
int64_t split(Rec* a, int64_t len, Less lt)
{
    Rec *x = a + 1, *y = a + len / 2, *z = a + len - 1, *m;
    if (lt(*x, *y)) m = lt(*y, *z) ? y : (lt(*x, *z) ? z : x);
    else m = lt(*x, *z) ? x : (lt(*y, *z) ? z : y);
    std::swap(*a, *m);
    Rec* lo = a + 1;
    Rec* hi = a + len;
    for (;;) {
        while (lt(*lo, *a)) ++lo;
        --hi;
        while (lt(*a, *hi)) --hi;
        if (!(lo < hi)) return lo - a;
        std::swap(*lo, *hi);
        ++lo;
    }
}

template <class Rec, class Less>
void insertion_block(Rec* a, int64_t len, Less lt)           // what the final insertion pass does to one finished block
{
    for (int64_t i = 1; i < len; i++) {
        const Rec v = a[i];
        int64_t j = i;
        while (j > 0 && lt(v, a[j - 1])) { a[j] = a[j - 1]; j--; }
        a[j] = v;
    }
}

template <class Rec, class Less>
class Pool {
public:
    Pool(unsigned nt, int64_t grain, Less lt) : nt_(nt < 1 ? 1 : nt), grain_(grain), lt_(lt) {}
    void run(Rec* a, int64_t n, int depth)
    {
        push({a, n, depth});
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt_; t++) th.emplace_back([this]() { work(); });
        work();
        for (auto& t : th) t.join();
    }
private:
    struct Job { Rec* a; int64_t len; int depth; };
    void push(Job j)
    {
        { std::lock_guard<std::mutex> g(mu_); jobs_.push_back(j); pending_++; }
        cv_.notify_one();
    }
    void work()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [this]() { return !jobs_.empty() || pending_ == 0; });
                if (jobs_.empty()) return;
                j = jobs_.back(); jobs_.pop_back();
            }
            loop(j.a, j.len, j.depth);
            bool done;
            { std::lock_guard<std::mutex> g(mu_); done = (--pending_ == 0); }
            if (done) cv_.notify_all();
        }
    }
    // the introsort loop on a[0, len): the right side of each split is handed on (or recursed into when small), the left side continues
    void loop(Rec* a, int64_t len, int depth)
    {
        while (len > 16) {
            if (depth == 0) { heap_sort(a, len, lt_); return; }
            --depth;
            const int64_t cut = split(a, len, lt_);
            if (len - cut >= grain_ && nt_ > 1) push({a + cut, len - cut, depth});
            else loop(a + cut, len - cut, depth);
            len = cut;
        }
        insertion_block(a, len, lt_);
    }
    unsigned nt_;
    int64_t grain_;
    Less lt_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Job> jobs_;
    int64_t pending_ = 0;
};

}  // namespace sort_replay_detail

// a[0, n) ends up exactly as std::sort(a, a + n, lt) leaves it (libstdc++), using up to n_threads threads
template <class Rec, class Less>
void replay_sort(Rec* a, uint64_t n, Less lt, unsigned n_threads)
{
    if (n < 2) return;
    int lg = 0;
    for (uint64_t v = n; v > 1; v >>= 1) lg++;                // floor(log2 n)
    sort_replay_detail::Pool<Rec, Less> pool(n < (1u << 15) ? 1u : n_threads, 1 << 13, lt);
    pool.run(a, (int64_t)n, 2 * lg);
}

}  // namespace dgb
