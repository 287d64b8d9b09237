// num / den for a den whose correctly rounded reciprocal y = RN(1/den) is known (a per-triangle constant of the leaf test,
// TriangleMeshDistance.h:598/620/698: a00, a11, a00 - 2 a01 + a11): the IEEE quotient from one multiplication and four FMAs instead
// of the ~68-instruction division sequence.  q0 = RN(num * y) is within an ulp or two of the quotient; r = num - q*den is exact in
// one FMA, and RN(q + r*y) of a faithful q is the correctly rounded quotient when y is the correctly rounded reciprocal (Markstein,
// "Computation of elementary functions on the IBM RISC System/6000 processor", 1990, thm. on division by FMA); the first correction
// makes q faithful, the second one lands on RN(num/den).  Valid while nothing over- or underflows: callers use it only for
// |num|, |den| in [2^-300, 2^300] and take the plain division otherwise.  tests/cpp/fast_div_check.cpp compares it with '/' on
// 4e8 operand pairs including significands next to 1 and 2 and short (exactly divisible) ones.  FMA here is the IEEE fused operation
// asked for by name -- it is not a contraction of the reference's arithmetic, whose quotient is reproduced bit for bit.
#pragma once
#include <cmath>
#include "dg_device.cuh"

namespace dgb {

DG_HD double div_by_known_reciprocal(double num, double den, double y)
{
#if defined(__CUDA_ARCH__)
    double q = __dmul_rn(num, y);
    double r = __fma_rn(-q, den, num);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-q, den, num);
    return __fma_rn(r, y, q);
#else
    double q = num * y;
    double r = std::fma(-q, den, num);
    q = std::fma(r, y, q);
    r = std::fma(-q, den, num);
    return std::fma(r, y, q);
#endif
}

// |x| in [2^-300, 2^300]
DG_HD bool in_fast_div_range(double x)
{
    const double a = x < 0 ? -x : x;
    return a >= 4.909093465297727e-91 && a <= 2.037035976334486e+90;
}

}  // namespace dgb
