// Checks the product's threaded replay of std::sort (discregrid_b200/csrc/reduce_field.cpp: replay_std_sort) against std::sort
// itself: same final arrangement INCLUDING the order of records with equal keys -- that order is what makes a reduced field's node
// numbering match the reference's when Morton keys tie.  Inputs: tie-heavy random, few distinct keys, presorted, and an input built
// by an adversary against this very std::sort (McIlroy, "A killer adversary for quicksort") so that the depth budget runs out and
// the heapsort branch is exercised too.  Test infrastructure; prints "OK" or the first difference.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <vector>
#include "reduce_field.h"

using dgb::KeyPos;

static bool same_as_std_sort(std::vector<KeyPos> a, unsigned nt, const char* what)
{
    std::vector<KeyPos> b(a);
    std::sort(b.begin(), b.end(), [](const KeyPos& x, const KeyPos& y) { return x.key < y.key; });
    dgb::replay_std_sort(a.data(), a.size(), nt);
    for (size_t i = 0; i < a.size(); i++)
        if (a[i].pos != b[i].pos || a[i].key != b[i].key) { std::printf("MISMATCH %s n=%zu threads=%u at %zu\n", what, a.size(), nt, i); return false; }
    return true;
}

// adversary: decides the keys while std::sort runs so that every pivot turns out to be (nearly) the smallest remaining value
struct Adversary {
    std::vector<int> val; int gas, nsolid = 0, candidate = 0;
    explicit Adversary(int n) : val(n, n), gas(n) {}
    bool less(int x, int y)
    {
        if (val[x] == gas && val[y] == gas) { if (x == candidate) val[x] = nsolid++; else val[y] = nsolid++; }
        if (val[x] == gas) candidate = x; else if (val[y] == gas) candidate = y;
        return val[x] < val[y];
    }
};

int main()
{
    bool ok = dgb::replay_std_sort_matches();
    if (!ok) std::printf("MISMATCH in the library's own self-check\n");
    uint64_t rng = 12345;
    auto next = [&rng]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    for (size_t n : {(size_t)5, (size_t)1 << 12, (size_t)1 << 17, (size_t)3000000})
        for (unsigned nt : {1u, 3u, 8u, 32u}) {
            std::vector<KeyPos> a(n);
            for (size_t i = 0; i < n; i++) a[i] = {next() % (n / 2 + 1), (uint32_t)i};
            ok = same_as_std_sort(a, nt, "random, ~2 records per key") && ok;
            for (size_t i = 0; i < n; i++) a[i].key = next() % 5;
            ok = same_as_std_sort(a, nt, "five distinct keys") && ok;
            for (size_t i = 0; i < n; i++) a[i].key = i / 3;
            ok = same_as_std_sort(a, nt, "sorted triples") && ok;
        }
    // killer input for this standard library's std::sort
    for (int n : {2000, 200000}) {
        Adversary adv(n);
        std::vector<int> idx(n); std::iota(idx.begin(), idx.end(), 0);
        std::sort(idx.begin(), idx.end(), [&](int x, int y) { return adv.less(x, y); });
        const uint64_t before = dgb::replay_std_sort_heap_fallbacks();
        std::vector<KeyPos> a(n);
        for (int i = 0; i < n; i++) a[i] = {(uint64_t)adv.val[i], (uint32_t)i};
        ok = same_as_std_sort(a, 4, "adversarial, distinct keys") && ok;
        const uint64_t mid = dgb::replay_std_sort_heap_fallbacks();
        for (int i = 0; i < n; i++) a[i].key = (uint64_t)adv.val[i] / 2;                 // the same shape with every key doubled up
        ok = same_as_std_sort(a, 4, "adversarial, paired keys") && ok;
        std::printf("n=%d heap fallbacks: %llu (distinct) %llu (paired)\n", n, (unsigned long long)(mid - before),
                    (unsigned long long)(dgb::replay_std_sort_heap_fallbacks() - mid));
        if (n >= 200000 && mid == before) { std::printf("the adversarial input did not exhaust the depth budget\n"); ok = false; }
    }
    std::printf(ok ? "OK\n" : "FAILED\n");
    return ok ? 0 : 1;
}
