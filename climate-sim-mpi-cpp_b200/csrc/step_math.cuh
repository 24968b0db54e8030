// step_math.cuh — per-cell arithmetic of the timestep, shared by every kernel variant.
//
// One place fixes the operation order so that all kernels (reference-shaped, fused one-step,
// temporally blocked) produce the same bits as the reference CPU build:
//   src/diffusion.cpp:12-14 and src/advection.cpp:16-31 (see kernels.cu header).
// Only round-to-nearest intrinsics are used; the compiler cannot contract them into FMA.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace csim {

struct StepK {
    double dtD;                    // dt*D
    double ndt;                    // -dt
    double vx, vy;
    double dx2, dy2, dx, dy;       // divisors (IEEE-division mode)
    double rdx2, rdy2, rdx, rdy;   // reciprocals (exact when the divisors are powers of two)
    int vx_pos, vy_pos;            // vx >= 0, vy >= 0
};

// a / d, correctly rounded, for a divisor that is known in advance: d > 0 finite and normal, y = RN(1/d)
// computed once on the host by an IEEE division.
//   q0 = RN(a*y), r0 = RN(a - q0*d) [FMA], q1 = RN(q0 + r0*y) [FMA]        (Markstein's correction step)
// q1 is almost always RN(a/d) already; instead of relying on a theorem about which (a, d) are exceptions, the
// result is CHECKED: r1 = a - q1*d is exact whenever q1 is within an ulp of a/d (and otherwise at least an ulp
// of q1 times d in magnitude), so  2|r1| < d * ulp(q1)  proves that q1 is the unique double nearest to a/d.
// Ties, results next to a power of two, tiny/huge/non-finite operands fail the test on purpose and take
// the IEEE division.  Six FP64-pipe instructions instead of the ~20 of a full division; the accepted
// result is identical to it by construction (tests/test_div_exact.py drives the host build of this very
// function against true division).
#ifdef __CUDACC__
#define CSIM_MATH_HD __host__ __device__ __forceinline__
#else
#define CSIM_MATH_HD inline
#endif
// The candidate quotient and whether the check proves it (no branch: callers decide what to do with !ok).
CSIM_MATH_HD double div_by_const_try(double a, double d, double y, bool& ok) {
#ifdef __CUDA_ARCH__
    const double q0 = __dmul_rn(a, y);
    const double r0 = __fma_rn(-q0, d, a);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, d, a);
    const unsigned long long qb = static_cast<unsigned long long>(__double_as_longlong(q1));
    const unsigned long long ab = static_cast<unsigned long long>(__double_as_longlong(a));
#else
    const double q0 = a * y;
    const double r0 = std::fma(-q0, d, a);
    const double q1 = std::fma(r0, y, q0);
    const double r1 = std::fma(-q1, d, a);
    unsigned long long qb, ab;
    std::memcpy(&qb, &q1, 8);
    std::memcpy(&ab, &a, 8);
#endif
    const unsigned qe = static_cast<unsigned>(qb >> 52) & 0x7ffu, ae = static_cast<unsigned>(ab >> 52) & 0x7ffu;
    // exponents far from the ends: no underflow in r1 or in d*ulp(q1), no overflow anywhere
    // (unsigned wrap-around folds the two-sided range tests into one compare each)
    const bool in_range = (qe - 120u) <= 1780u && (ae - 120u) <= 1780u;
    // ulp(q1) = 2^(e-52); one binade lower when q1 is a power of two (the neighbour below is closer)
    const unsigned long long ue = static_cast<unsigned long long>(qe - ((qb & 0xfffffffffffffull) ? 52u : 53u)) << 52;
#ifdef __CUDA_ARCH__
    const double ulp = __longlong_as_double(static_cast<long long>(ue));
    ok = in_range && __dmul_rn(2.0, fabs(r1)) < __dmul_rn(d, ulp);
#else
    double ulp;
    std::memcpy(&ulp, &ue, 8);
    ok = in_range && 2.0 * std::fabs(r1) < d * ulp;
#endif
    return q1;
}
CSIM_MATH_HD double div_by_const(double a, double d, double y) {
    bool ok;
    const double q = div_by_const_try(a, d, y, ok);
    if (ok) return q;
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, d);
#else
    return a / d;
#endif
}
// 1 if div_by_const(a, d, 1/d) took the checked fast path, 0 if it fell back (host only; tests)
inline int div_by_const_took_fast_path(double a, double d) {
    const double y = 1.0 / d;
    const double q0 = a * y, r0 = std::fma(-q0, d, a), q1 = std::fma(r0, y, q0), r1 = std::fma(-q1, d, a);
    unsigned long long qb, ab;
    std::memcpy(&qb, &q1, 8);
    std::memcpy(&ab, &a, 8);
    const unsigned qe = static_cast<unsigned>(qb >> 52) & 0x7ffu, ae = static_cast<unsigned>(ab >> 52) & 0x7ffu;
    if (!(qe >= 120u && qe <= 1900u && ae >= 120u && ae <= 1900u)) return 0;
    const unsigned long long ue = static_cast<unsigned long long>(qe - ((qb & 0xfffffffffffffull) ? 52u : 53u)) << 52;
    double ulp;
    std::memcpy(&ulp, &ue, 8);
    return 2.0 * std::fabs(r1) < d * ulp ? 1 : 0;
}

template <bool kDiv>
__device__ __forceinline__ double scale(double v, double divisor, double recip) {
    if (kDiv) return div_by_const(v, divisor, recip);
    return __dmul_rn(v, recip);
}

// c + (dt*D) * ( ((e-2c)+w)/dx² + ((n-2c)+s)/dy² )
template <bool kDiv>
__device__ __forceinline__ double diffusion_update(double c, double w, double e, double s, double n,
                                                   const StepK& k) {
    const double c2 = __dmul_rn(2.0, c);
    const double lx = scale<kDiv>(__dadd_rn(__dsub_rn(e, c2), w), k.dx2, k.rdx2);
    const double ly = scale<kDiv>(__dadd_rn(__dsub_rn(n, c2), s), k.dy2, k.rdy2);
    return __dadd_rn(c, __dmul_rn(k.dtD, __dadd_rn(lx, ly)));
}

// (-dt) * ( vx*dudx + vy*dudy ), first-order upwind
template <bool kDiv>
__device__ __forceinline__ double advection_increment(double c, double w, double e, double s, double n,
                                                      const StepK& k) {
    const double ddx = scale<kDiv>(k.vx_pos ? __dsub_rn(c, w) : __dsub_rn(e, c), k.dx, k.rdx);
    const double ddy = scale<kDiv>(k.vy_pos ? __dsub_rn(c, s) : __dsub_rn(n, c), k.dy, k.rdy);
    const double adv = __dadd_rn(__dmul_rn(k.vx, ddx), __dmul_rn(k.vy, ddy));
    return __dmul_rn(k.ndt, adv);
}

// diffusion result, then += advection increment: the two reference passes in one
template <bool kDiv>
__device__ __forceinline__ double fused_update(double c, double w, double e, double s, double n,
                                               const StepK& k) {
    return __dadd_rn(diffusion_update<kDiv>(c, w, e, s, n, k), advection_increment<kDiv>(c, w, e, s, n, k));
}

}  // namespace csim
