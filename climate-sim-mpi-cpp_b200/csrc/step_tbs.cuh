// step_tbs.cuh — the staged variant of the blocked sweep: level-0 rows land in shared memory by TMA.
//
// Why (profiles/r02_tuning.md): at 16384^2 the register-only sweep k_step_tb<T=3> spends 97 % of its
// time in the fast loop, moves 4.34 GB per launch at 5.6 TB/s (87 % of the measured copy bandwidth) and
// loses 38 % of its issue slots to long-scoreboard stalls on the first instructions that touch a freshly
// loaded row: it is bound by HBM, and a row requested one tick ahead does not arrive in time.  Both ends
// of that are attacked here:
//   * the level-0 rows never occupy registers while in flight.  One elected lane per warp issues
//     `cp.async.bulk` (TMA, 1-D bulk copy: one 1 KB row segment per instruction) for the row pair that is
//     due kTbsSlots-1 ticks later into a per-warp ring in shared memory, completion is signalled on an
//     mbarrier (complete_tx), and the warp reads the four level-0 rows of a tick with LDS.128 when it
//     needs them.  Warps stay independent: the ring and its barriers are per warp, no __syncthreads.
//   * the 32 registers that held the level-0 rows pay for a fourth time level: T = 4 with the register
//     budget of T = 3 (168, 12 warps per SM), i.e. one read and one write of the field per FOUR steps.
// Everything else — strips, chunks, tick flavours, boundary rules, stores, arithmetic — is step_tb.cuh's.
#pragma once
#include "step_tb.cuh"

namespace csim {

constexpr int kTbsSlots = 4;                    // ring of row pairs per warp (a power of two)
constexpr int kTbsRowBytes = kTbWidth * 8;      // 1 KB: one strip row
constexpr int kTbsPairBytes = 2 * kTbsRowBytes;

__device__ __forceinline__ uint32_t tbs_smem(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void tbs_bar_init(uint32_t bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tbs_bar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tbs_bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "TBS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra TBS_DONE;\n\t"
        "bra TBS_WAIT;\n\t"
        "TBS_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// 1-D bulk copy global → shared through the TMA unit; `bytes` arrive on the mbarrier as transaction count
__device__ __forceinline__ void tbs_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tbs_lds4(uint32_t addr, double (&v)[4]) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(addr + 16));
}

// Request row pair p of the work item (rows r0+2p, r0+2p+1 of the strip) into its ring slot.
__device__ __forceinline__ void tbs_issue(uint32_t ring0, uint32_t bar0, const double* g, long long pitch, int p) {
    const int slot = p & (kTbsSlots - 1);
    const uint32_t bar = bar0 + 8u * slot, dst = ring0 + static_cast<uint32_t>(kTbsPairBytes) * slot;
    const double* src = g + 2LL * p * pitch;
    tbs_bar_expect(bar, kTbsPairBytes);
    tbs_bulk_load(dst, src, kTbsRowBytes, bar);
    tbs_bulk_load(dst + kTbsRowBytes, src + pitch, kTbsRowBytes, bar);
}

// One tick (see tb_tick): rows r, r+1 of level 0 become available, rows r-T, r-T+1 of level T leave.
// st holds levels 1 … T-1 only; level 0 is read from the ring: pair t-1 (rows r-2, r-1) and pair t (r, r+1).
template <int T, int MODE, int VXS, int VYS, int PH, int KIND>
__device__ __forceinline__ void tbs_tick(const TbArgs& a, const TbLane& ln, int lane, bool lane_store_all, int r, int q,
                                         int h, int ya, int yb, int npairs, uint32_t ring0, uint32_t bar0,
                                         const double* g, double*& dst, double (&st)[T - 1][2][2][4]) {
    double fin[2][4];
    const int t = (q >> 1) + T;  // tick index of the work item: q = 2t - 2T
    {
        const uint32_t la = ring0 + static_cast<uint32_t>(kTbsPairBytes) * ((t - 1) & (kTbsSlots - 1)) + 32u * lane;
        const uint32_t lc = ring0 + static_cast<uint32_t>(kTbsPairBytes) * (t & (kTbsSlots - 1)) + 32u * lane;
        // pairs past the rows the item needs are neither requested nor awaited: stale finite rows feed
        // rows nobody stores
        if (t < npairs) tbs_bar_wait(bar0 + 8u * (t & (kTbsSlots - 1)), (t / kTbsSlots) & 1);
        double A[4], B[4], C[4], D[4];
        tbs_lds4(la, A);
        tbs_lds4(la + kTbsRowBytes, B);
        tbs_lds4(lc, C);
        tb_row<MODE, VXS, VYS, KIND>(a, ln, r - 1, A, B, C, T > 1 ? st[0][1 - PH][0] : fin[0]);
        tbs_lds4(lc + kTbsRowBytes, D);
        tb_row<MODE, VXS, VYS, KIND>(a, ln, r, B, C, D, T > 1 ? st[0][1 - PH][1] : fin[1]);
        // pair t-1 has been read for the last time: its slot takes the pair that is due kTbsSlots-1 ticks on
        __syncwarp();
        if (lane == 0 && t + kTbsSlots - 1 < npairs) tbs_issue(ring0, bar0, g, a.pitch, t + kTbsSlots - 1);
    }
#pragma unroll
    for (int k = 1; k < T; ++k) {
        double(&A)[4] = st[k - 1][PH][0];
        double(&B)[4] = st[k - 1][PH][1];
        double(&C)[4] = st[k - 1][1 - PH][0];
        double(&D)[4] = st[k - 1][1 - PH][1];
        if (k + 1 < T) {
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k - 1, A, B, C, st[k + 1 < T ? k : k - 1][1 - PH][0]);
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k, B, C, D, st[k + 1 < T ? k : k - 1][1 - PH][1]);
        } else {
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k - 1, A, B, C, fin[0]);
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k, B, C, D, fin[1]);
        }
    }
    if (KIND == TICK_FAST) {
        if (lane_store_all && q >= 0 && q < h) tb_store4(dst, fin[0]);
        if (lane_store_all && q + 1 >= 0 && q + 1 < h) tb_store4(dst + a.pitch, fin[1]);
    } else {
        tb_store_row(a, ln, lane, lane_store_all, r - T, ya, yb, fin[0]);
        tb_store_row(a, ln, lane, lane_store_all, r - T + 1, ya, yb, fin[1]);
    }
    dst += 2 * a.pitch;
}

template <int T, int MODE, int VXS, int VYS>
__global__ void __launch_bounds__(32 * kTbWarpsPerBlock, kTbBlocksPerSM) k_step_tbs(const __grid_constant__ TbArgs a) {
    static_assert(T >= 2 && T <= kTbMaxT, "T out of range");
    __shared__ __align__(128) double ring[kTbWarpsPerBlock][kTbsSlots][2][kTbWidth];
    __shared__ __align__(8) unsigned long long bars[kTbWarpsPerBlock][kTbsSlots];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kTbWarpsPerBlock + warp;
    if (item >= a.n_items) return;  // warp-uniform; warps never synchronise with each other

    int strip, ya, yb;
    if (!tb_item_map(a, item, strip, ya, yb)) return;
    tb_frame_enter(a, item, lane);  // coupled launches: no ghost line is requested before its halo has landed
    const int xb = strip * kTbWout - kTbHX;
    TbLane ln;
    ln.x0 = xb + lane * kTbCells;
    const bool lane_store_all = lane >= 1 && lane <= 30 && ln.x0 >= a.sx0 && ln.x0 + 3 < a.sx1;
    const bool lane_partial = !lane_store_all && lane >= 1 && lane <= 30 && ln.x0 + 3 >= a.sx0 && ln.x0 < a.sx1;
    const bool strip_fast =
        xb >= a.fx0 && xb + kTbWidth <= a.fx1 && __ballot_sync(0xffffffffu, lane_partial) == 0u;
    const bool strip_xmask = !(((a.phys & 1) && a.bcL != 2 && xb <= 0) ||
                               ((a.phys & 2) && a.bcR != 2 && xb + kTbWidth >= a.nx));
    const int mid = strip_fast ? TICK_FAST : (strip_xmask ? TICK_XMASK : TICK_GEN);
    {
        const bool physL = a.phys & 1, physR = a.phys & 2;
        ln.inx = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (ln.x0 + i >= a.xlo && ln.x0 + i < a.xhi) ln.inx |= 1 << i;
        ln.at_l = physL && ln.x0 == 0;
        ln.ghost_l = physL && ln.x0 + 3 == -1;
        const int dr = a.nx - 1 - ln.x0, dg = a.nx - ln.x0;
        ln.at_r = (physR && dr >= 0 && dr < 4) ? dr : -1;
        ln.ghost_r = (physR && dg >= 0 && dg < 4) ? dg : -1;
    }

    // this warp's ring and barriers; the ring starts as zeros ("rows" before the first pair)
    const uint32_t ring0 = tbs_smem(&ring[warp][0][0][0]);
    const uint32_t bar0 = tbs_smem(&bars[warp][0]);
    for (int i = lane; i < kTbsSlots * kTbsPairBytes / 16; i += 32)
        asm volatile("st.shared.v2.f64 [%0], {%1,%1};" ::"r"(ring0 + 16u * i), "d"(0.0) : "memory");
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTbsSlots; ++s) tbs_bar_init(bar0 + 8u * s);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic zero stores before TMA writes
    __syncwarp();

    const int h = yb - ya;
    int r = ya - T;
    const int npairs = (h + 2 * T + 1) >> 1;
    // row r of the strip, from its first column: a 1 KB aligned segment per row (xb*8 is a multiple of 32;
    // a strip that overhangs the tile reads on into the padding of the next row, inside the allocation
    // because no item reads beyond row ny+T+1)
    const double* g = a.u + static_cast<long long>(r) * a.pitch + xb;
    if (lane == 0) {
#pragma unroll
        for (int p = 0; p < kTbsSlots - 1; ++p)
            if (p < npairs) tbs_issue(ring0, bar0, g, a.pitch, p);
    }

    double st[T - 1][2][2][4];
#pragma unroll
    for (int k = 0; k < T - 1; ++k)
#pragma unroll
        for (int e = 0; e < 16; ++e) st[k][e >> 3][(e >> 2) & 1][e & 3] = 0.0;
    double* dst = a.out + static_cast<long long>(r - T) * a.pitch + ln.x0;  // first tick finishes rows r-T, r-T+1

    const int r_end = yb + T;
    while (r < r_end) {
        if (mid != TICK_GEN && r - T >= a.fy0 && r + 2 < a.fy1) {
            int n_it = (min(a.fy1 - 2, r_end) - r + 3) >> 2;
            if (mid == TICK_FAST) {
                int q = r - T - ya;
                r += 4 * n_it;
#pragma unroll 1
                for (; n_it > 0; --n_it, q += 4) {
                    tbs_tick<T, MODE, VXS, VYS, 0, TICK_FAST>(a, ln, lane, lane_store_all, 0, q, h, 0, 0, npairs, ring0,
                                                              bar0, g, dst, st);
                    tbs_tick<T, MODE, VXS, VYS, 1, TICK_FAST>(a, ln, lane, lane_store_all, 0, q + 2, h, 0, 0, npairs,
                                                              ring0, bar0, g, dst, st);
                }
            } else {
#pragma unroll 1
                for (; n_it > 0; --n_it, r += 4) {
                    tbs_tick<T, MODE, VXS, VYS, 0, TICK_XMASK>(a, ln, lane, lane_store_all, r, r - T - ya, h, ya, yb,
                                                               npairs, ring0, bar0, g, dst, st);
                    tbs_tick<T, MODE, VXS, VYS, 1, TICK_XMASK>(a, ln, lane, lane_store_all, r + 2, r + 2 - T - ya, h, ya,
                                                               yb, npairs, ring0, bar0, g, dst, st);
                }
            }
        } else {
            tbs_tick<T, MODE, VXS, VYS, 0, TICK_GEN>(a, ln, lane, lane_store_all, r, r - T - ya, h, ya, yb, npairs, ring0,
                                                     bar0, g, dst, st);
            tbs_tick<T, MODE, VXS, VYS, 1, TICK_GEN>(a, ln, lane, lane_store_all, r + 2, r + 2 - T - ya, h, ya, yb, npairs,
                                                     ring0, bar0, g, dst, st);
            r += 4;
        }
    }
    tb_frame_leave(a, item, lane);
}

}  // namespace csim
