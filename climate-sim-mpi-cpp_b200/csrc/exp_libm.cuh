// exp_libm.cuh — the host libm's exp(), restated so that a device kernel returns the same bits.
//
// Why: the initial condition (reference src/init.cpp:12-33) is A * std::exp(-r2 / (2 sig^2)) per cell,
// and parity is defined bit for bit, so a device-side initial condition must reproduce the HOST's exp,
// not CUDA's.  glibc >= 2.28 computes exp(x) with a 128-entry table of 2^(k/128) and a degree-5
// polynomial (S. Nagy's algorithm; the table is re-derived by tools/gen_exp_table.py, the constants
// below are the published minimax coefficients and ln2 splits):
//     k  = round(x * 128/ln2),  r = x - k*ln2/128  (two-step, hi/lo),
//     exp(x) = 2^(k/128) * exp(r) = scale * (1 + tail + r + r^2 (C2 + r C3) + r^4 (C4 + r C5)).
// The result is NOT correctly rounded (< 0.511 ulp), so the exact sequence of roundings matters, and
// that sequence depends on the build of libm that runs on the host:
//     FMA = true   the x86-64 ifunc variant for CPUs with FMA (compiled with -mfma: the compiler
//                  contracted every a*b+c of the source; contraction pattern read off the
//                  disassembly of glibc 2.39's __exp_fma),
//     FMA = false  the source order without contraction (baseline x86-64 / any non-FMA build).
// csim_exp_variant() (host_misc.cpp) probes the host's exp() against both restatements and the device
// kernel uses the one that matches; if neither does, the device path is refused (CSIM_ERR_UNSUPPORTED)
// and callers keep the host initial condition.  tests/test_exp_restatement.py pins the restatement
// against the host libm on millions of inputs, tests/test_gpu_parity.py the device tile against the
// host tile.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define CSIM_HD __host__ __device__ __forceinline__
#else
#define CSIM_HD inline
#endif

namespace csim {

namespace expd {
// one rounding per call; never contracted (device: intrinsics; host: -ffp-contract=off, std::fma)
CSIM_HD double mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
CSIM_HD double add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
CSIM_HD double sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
CSIM_HD double fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
CSIM_HD uint64_t bits(double v) {
#ifdef __CUDA_ARCH__
    return static_cast<uint64_t>(__double_as_longlong(v));
#else
    uint64_t b;
    std::memcpy(&b, &v, sizeof b);
    return b;
#endif
}
CSIM_HD double from_bits(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(static_cast<long long>(b));
#else
    double v;
    std::memcpy(&v, &b, sizeof v);
    return v;
#endif
}
}  // namespace expd

// Result for |x| >= 512: 2^(k/128) alone may overflow or underflow although the result does not, so the
// scale is moved by 2^-1009 (k > 0) or 2^+1022 (k < 0) first; below 2^-1022 the sum is rounded to the
// subnormal grid in two exact steps to avoid a double rounding.
template <bool FMA>
CSIM_HD double exp_libm_special(double tmp, uint64_t sbits, uint64_t ki) {
    using namespace expd;
    if ((ki & 0x80000000ull) == 0) {
        sbits -= 1009ull << 52;
        const double scale = from_bits(sbits);
        const double y = FMA ? fma(scale, tmp, scale) : add(scale, mul(scale, tmp));
        return mul(0x1p1009, y);
    }
    sbits += 1022ull << 52;
    const double scale = from_bits(sbits);
    const double prod = mul(scale, tmp);  // used twice, hence not contracted in either build
    double y = add(scale, prod);
    if (y < 1.0) {
        double lo = add(sub(scale, y), prod);
        const double hi = add(1.0, y);
        lo = add(add(sub(1.0, hi), y), lo);
        y = sub(add(hi, lo), 1.0);
        if (y == 0.0) y = 0.0;  // no -0.0
    }
    return mul(0x1p-1022, y);
}

// tab: the 256 words of exp_table.inc
template <bool FMA>
CSIM_HD double exp_libm(double x, const uint64_t* tab) {
    using namespace expd;
    const double InvLn2N = 0x1.71547652b82fep0 * 128, Shift = 0x1.8p52;
    const double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
    const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5,
                 C5 = 0x1.1111167a4d017p-7;
    const uint64_t ix = bits(x);
    uint32_t abstop = static_cast<uint32_t>(ix >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu) {                             // |x| < 2^-54 or |x| >= 512 or NaN
        if (abstop - 0x3c9u >= 0x80000000u) return add(1.0, x);  // tiny: exp(x) rounds like 1 + x
        if (abstop >= 0x409u) {                                  // |x| >= 1024
            if (ix == 0xfff0000000000000ull) return 0.0;         // exp(-inf)
            if (abstop >= 0x7ffu) return add(1.0, x);            // +inf, NaN
            return (ix >> 63) ? 0.0 : from_bits(0x7ff0000000000000ull);  // underflow to +0 / overflow
        }
        abstop = 0;  // 512 <= |x| < 1024: through the main path, finished by exp_libm_special
    }
    double kd = FMA ? fma(InvLn2N, x, Shift) : add(mul(InvLn2N, x), Shift);
    const uint64_t ki = bits(kd);
    kd = sub(kd, Shift);
    double r;
    if (FMA) {
        r = fma(kd, NegLn2hiN, x);
        r = fma(kd, NegLn2loN, r);
    } else {
        r = add(add(x, mul(kd, NegLn2hiN)), mul(kd, NegLn2loN));
    }
    const uint64_t idx = 2 * (ki % 128);
    const uint64_t top = ki << 45;
    const double tail = from_bits(tab[idx]);
    const uint64_t sbits = tab[idx + 1] + top;
    const double r2 = mul(r, r);
    double tmp;
    if (FMA) {
        const double p23 = fma(r, C3, C2);
        const double tr = add(tail, r);
        const double p45 = fma(r, C5, C4);
        const double q = fma(p23, r2, tr);
        tmp = fma(mul(r2, r2), p45, q);
    } else {
        tmp = add(add(add(tail, r), mul(r2, add(C2, mul(r, C3)))), mul(mul(r2, r2), add(C4, mul(r, C5))));
    }
    if (abstop == 0) return exp_libm_special<FMA>(tmp, sbits, ki);
    const double scale = from_bits(sbits);
    return FMA ? fma(scale, tmp, scale) : add(scale, mul(scale, tmp));
}

}  // namespace csim
