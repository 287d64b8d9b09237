// See sort_replay.h: the once-per-process check of the replay against the standard library's own std::sort.
#include "sort_replay.h"

namespace dgb {

std::atomic<uint64_t> g_replay_heap_sorts{0};

namespace {
struct KeyId { uint64_t key; uint32_t id; };
struct DKeyId { double key; int32_t id; };

template <class Rec, class Less>
bool same(std::vector<Rec> a, Less lt)
{
    std::vector<Rec> b(a);
    std::sort(b.begin(), b.end(), lt);
    replay_sort(a.data(), a.size(), lt, 4);
    for (size_t i = 0; i < a.size(); i++) if (a[i].id != b[i].id) return false;
    return true;
}
}  // namespace

bool sort_replay_matches_std_sort()
{
    static const bool ok = []() {
        // tie-heavy, ordered, organ-pipe and few-distinct inputs; sizes around the block threshold and large enough to use the threads
        // (small sizes: this runs in every process that sorts; threads change the schedule, not the algorithm -- the threaded
        // runs are compared with std::sort by tests/cpp/sort_replay_check.cpp)
        const uint64_t sizes[] = {0, 1, 2, 3, 15, 16, 17, 33, 100, 1000, 3000};
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        auto next = [&rng]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
        for (uint64_t n : sizes)
            for (int kind = 0; kind < 6; kind++) {
                std::vector<KeyId> a(n); std::vector<DKeyId> d(n);
                for (uint64_t i = 0; i < n; i++) {
                    uint64_t k;
                    switch (kind) {
                        case 0: k = next() % (n / 3 + 1); break;              // about three records per key
                        case 1: k = next() & 7; break;                       // eight distinct keys
                        case 2: k = i / 2; break;                            // sorted, pairs
                        case 3: k = (n - i) / 2; break;                      // reversed, pairs
                        case 4: k = i < n / 2 ? i : n - i; break;            // organ pipe
                        default: k = next(); break;                          // distinct
                    }
                    a[i] = {k, (uint32_t)i};
                    d[i] = {(double)(k % 1000003) * 0.125 - 7.0, (int32_t)i};
                }
                if (!same(a, [](const KeyId& x, const KeyId& y) { return x.key < y.key; })) return false;
                if (!same(d, [](const DKeyId& x, const DKeyId& y) { return x.key < y.key; })) return false;
            }
        return true;
    }();
    return ok;
}

}  // namespace dgb
