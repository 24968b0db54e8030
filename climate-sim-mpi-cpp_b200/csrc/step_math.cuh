// step_math.cuh — per-cell arithmetic of the timestep, shared by every kernel variant.
//
// One place fixes the operation order so that all kernels (reference-shaped, fused one-step,
// temporally blocked) produce the same bits as the reference CPU build:
//   src/diffusion.cpp:12-14 and src/advection.cpp:16-31 (see kernels.cu header).
// Only round-to-nearest intrinsics are used; the compiler cannot contract them into FMA.
#pragma once
#include <cstdint>

namespace csim {

struct StepK {
    double dtD;                    // dt*D
    double ndt;                    // -dt
    double vx, vy;
    double dx2, dy2, dx, dy;       // divisors (IEEE-division mode)
    double rdx2, rdy2, rdx, rdy;   // reciprocals (exact when the divisors are powers of two)
    int vx_pos, vy_pos;            // vx >= 0, vy >= 0
};

template <bool kDiv>
__device__ __forceinline__ double scale(double v, double divisor, double recip) {
    if (kDiv) return __ddiv_rn(v, divisor);
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
