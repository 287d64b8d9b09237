// Brute-force check (CPU) of discregrid_b200/csrc/fast_div.h: div_by_known_reciprocal(num, den, RN(1/den)) == num / den, bit for bit,
// for random operands and for the significand patterns where a reciprocal-based quotient is most likely to be off by one ulp
// (den next to 1 and next to 2, num likewise, short significands with exact or nearly exact quotients).  The same function body is
// what the K1_FAST_DIV build of the leaf test runs on the device (fma.rn.f64 is the IEEE fused operation on both sides).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "fast_div.h"

static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double mk(uint64_t mant, int e) { uint64_t u = ((uint64_t)(1023 + e) << 52) | (mant & 0xfffffffffffffull); double d; std::memcpy(&d, &u, 8); return d; }

int main(int argc, char** argv)
{
    const long long n = argc > 1 ? std::atoll(argv[1]) : 100000000LL;
    long long bad = 0;
    for (long long it = 0; it < n; it++) {
        uint64_t ma = rnd(), mb = rnd();
        const int mode = (int)(it & 7);
        if (mode == 1) mb = 0xfffffffffffffull - (rnd() & 0xff);
        if (mode == 2) mb = rnd() & 0xff;
        if (mode == 3) ma = 0xfffffffffffffull - (rnd() & 0xff);
        if (mode == 4) ma = rnd() & 0xff;
        if (mode == 5) { mb = rnd() & 0xfffff; ma = rnd() & 0xfffff; }
        const int ea = (mode == 6) ? (int)(rnd() % 600) - 300 : (int)(rnd() % 40) - 20, eb = (mode == 6) ? (int)(rnd() % 600) - 300 : (int)(rnd() % 40) - 20;
        double a = mk(ma, ea);
        const double b = mk(mb, eb);
        if (rnd() & 1) a = -a;
        if (!dgb::in_fast_div_range(a) || !dgb::in_fast_div_range(b)) continue;
        const double want = a / b, got = dgb::div_by_known_reciprocal(a, b, 1.0 / b);
        if (std::memcmp(&want, &got, 8) != 0) { if (bad++ < 5) std::printf("MISMATCH %a / %a: want %a got %a\n", a, b, want, got); }
    }
    std::printf(bad ? "FAILED (%lld mismatches)\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
