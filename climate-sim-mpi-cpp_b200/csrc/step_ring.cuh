// step_ring.cuh — the second-generation sweep (CSIM_TB_KERNEL=ring): same strip/chunk decomposition,
// same row routines and arithmetic as step_tb.cuh, but ONE row per tick, a rotating ring of row
// slots in registers, and level-0 rows that land in shared memory (cp.async) several ticks before a
// register ever depends on them.
//
// Why: k_step_tb (two rows per tick) keeps four rows per time level in registers (96 registers of
// state at T = 3, 162-168 in total → 12 warps per SM) and its FP64 pipe sits at 74-79 % of the active
// cycles with "wait" (fixed-latency dependency) as the top stall (profiles/r01b_*).  Here a warp
// holds 2T + 2 row slots:
//     L            the level-0 row read back from shared memory during the previous tick
//     c0           the newest level-0 row in use
//     b_k, a_k     the two older rows of level k  (k = 0 … T-1)
// A tick at level-0 row r computes, for k = 0 … T-1, row r-1-k of level k+1 from (a_k, b_k, c_k) and
// writes it INTO a_k's registers — a_k (the "south" row) is dead once the update has read it — where
// it serves as c_{k+1} for the next level of the same tick and as b_{k+1}, a_{k+1} in the two ticks
// after.  The last level's result, row r-T of level T, is stored and its slot receives row r+2 from
// shared memory.  From one tick to the next every role moves to the neighbouring slot — a pure
// rotation of the ring — so U consecutive ticks are unrolled with compile-time slot numbers and the
// ring is turned back by U slots with register moves (U = N would need no moves but is 30 KB of
// code, which lost 14 % of the issue slots to instruction fetch; U = 3 keeps the loop near 13 KB).
// State: 8(2T+2) registers — 64 at T = 3 — under 128 registers per thread: 16 warps per SM.
//
// Level-0 rows travel HBM → shared memory with cp.async (each lane copies and later reads back its
// own 32 bytes, so no barrier is involved): kRingStages rows per warp are in flight, i.e. a row is
// requested kRingStages-2 ticks (≈ 5 µs) before the LDS that brings it into the ring.  The first
// register-landing version of this kernel lost 17 % of its samples to one rotation MOV that touched a
// row requested a single tick earlier (profiles/r01b_tuning.md).
//
// Ticks that touch a boundary (edge strips, first/last rows of the tile, chunk ends that do not fill
// a group of U) run a single generic tick (GEN = true rows of step_tb.cuh) at phase 0 and then
// rotate the ring by one slot; only the unrolled groups are hot.
#pragma once
#include "step_tb.cuh"

namespace csim {

#ifndef CSIM_RING_WARPS
#define CSIM_RING_WARPS 4
#endif
#ifndef CSIM_RING_BLOCKS
#define CSIM_RING_BLOCKS 4
#endif
// 4 warps x 4 CTAs = 16 warps per SM at 128 registers per thread.  Anything above 12 warps per SM
// means 4 warps on some SM sub-partition, i.e. at most 16384/4/32 = 128 registers per thread
// (2 x 7 CTAs was tried to get 146: ptxas still caps at 128).
constexpr int kRingWarpsPerBlock = CSIM_RING_WARPS;
constexpr int kRingBlocksPerSM = CSIM_RING_BLOCKS;
constexpr int kRingStages = 8;       // level-0 rows per warp in shared memory (1 KB each)
constexpr int kRingSmemBytes = kRingWarpsPerBlock * kRingStages * 1024;

#ifndef CSIM_RING_U
#define CSIM_RING_U 3
#endif

__host__ __device__ constexpr int ring_slots(int T) { return 2 * T + 2; }

// slot of a role at phase P of the unrolled group
// roles: 0 = L, 1 = c0, 2+2k = b_k, 3+2k = a_k; c_k (k > 0) sits in a_{k-1}'s slot
template <int N>
__host__ __device__ constexpr int ring_slot(int role, int P) {
    return ((role - P) % N + N) % N;
}

// ---- level-0 rows through shared memory ----------------------------------------------------------
// Stage s of a warp is the 1 KB row segment as it lies in memory.  A request is two cp.async
// instructions in which the 32 lanes cover 512 contiguous bytes each (whole 32-byte sectors per
// instruction: a first version had every lane copy its own two 16-byte halves, i.e. half a sector per
// lane and instruction, fetched every sector twice from L2 and ran at 6.0e11).  The lane that later
// reads 32 bytes back is therefore not the lane that copied them: the read is ordered after the
// copies of the whole warp by cp.async.wait_group + __syncwarp().
__device__ __forceinline__ void ring_request(unsigned req_addr, const double* src, bool ok0, bool ok1) {
    const unsigned b0 = ok0 ? 16u : 0u, b1 = ok1 ? 16u : 0u;  // 0: zero-fill, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(req_addr), "l"(src), "r"(b0) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(req_addr + 512u), "l"(src + 64), "r"(b1)
                 : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void ring_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}
__device__ __forceinline__ void ring_fetch(unsigned get_addr, double (&v)[4]) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(get_addr) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(get_addr + 16u) : "memory");
}

// The T level-rows of one tick (see the file header); leaves row r-T of level T in role N-1's slot.
template <int T, int MODE, int VXS, int VYS, int P, bool GEN>
__device__ __forceinline__ void ring_levels(const TbArgs& a, const TbLane& ln, int r, double (&S)[2 * T + 2][4]) {
    constexpr int N = ring_slots(T);
#pragma unroll
    for (int k = 0; k < T; ++k) {
        double(&s)[4] = S[ring_slot<N>(3 + 2 * k, P)];  // a_k : row r-2-k of level k
        double(&c)[4] = S[ring_slot<N>(2 + 2 * k, P)];  // b_k : row r-1-k
        double(&n)[4] = S[ring_slot<N>(1 + 2 * k, P)];  // c_k : row r-k
        double res[4];
        tb_row<MODE, VXS, VYS, GEN>(a, ln, r - 1 - k, s, c, n, res);
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = res[i];      // c_{k+1} replaces a_k
    }
}

// Per-warp cursor of the shared-memory stages: row `row` lives in stage row mod kRingStages.
struct RingFeed {
    unsigned base;      // shared address of this warp's stage 0
    unsigned lane;
    int req_stage;      // stage of the next row to request
    int get_stage;      // stage of the next row to read back
    const double* src;  // global address of this lane's first 16 bytes of the next row to request
    bool ok0, ok1;      // this lane's two 16-byte pieces lie inside the row allocation
};
__device__ __forceinline__ void feed_request(RingFeed& f, long long pitch, bool wanted) {
    ring_request(f.base + static_cast<unsigned>(f.req_stage) * 1024u + f.lane * 16u, f.src, wanted && f.ok0,
                 wanted && f.ok1);
    f.req_stage = (f.req_stage + 1) & (kRingStages - 1);
    f.src += pitch;
}
// wait until at most PENDING of this warp's requests are outstanding, then read this lane's 4 cells
template <int PENDING>
__device__ __forceinline__ void feed_get(RingFeed& f, double (&v)[4]) {
    ring_wait<PENDING>();
    __syncwarp();
    ring_fetch(f.base + static_cast<unsigned>(f.get_stage) * 1024u + f.lane * 32u, v);
    f.get_stage = (f.get_stage + 1) & (kRingStages - 1);
}

// U consecutive hot ticks (phases 0 … U-1) starting at level-0 row r, then the rotation by U slots.
// Hot ticks run only where every lane stores all four cells or none and every row involved is an
// interior row inside the allocation: one predicated 256-bit store per tick, one unconditional row
// request, one read-back of the row requested kRingStages-2 ticks ago.
template <int T, int MODE, int VXS, int VYS, int U, int P>
__device__ __forceinline__ void ring_group(const TbArgs& a, const TbLane& ln, bool lane_store_all, int r, int ya,
                                           int yb, int req_stop, RingFeed& feed, double*& dst,
                                           double (&S)[2 * T + 2][4]) {
    constexpr int N = ring_slots(T);
    if constexpr (P < U) {
        ring_levels<T, MODE, VXS, VYS, P, false>(a, ln, r + P, S);
        double(&fin)[4] = S[ring_slot<N>(N - 1, P)];
        if (lane_store_all && r + P - T >= ya && r + P - T < yb) tb_store4(dst, fin);
        dst += a.pitch;
        feed_request(feed, a.pitch, r + P + kRingStages < req_stop);  // row r+P+kRingStages, if anyone needs it
        if (P == U - 1 && U < N) {
            // rotate by U: the slot of role i at phase U is (i - U) mod N; move it back to slot i
            double t[N][4];
#pragma unroll
            for (int q = 0; q < N; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) t[q][i] = S[ring_slot<N>(q, U)][i];
#pragma unroll
            for (int q = 0; q < N; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) S[q][i] = t[q][i];
            feed_get<kRingStages - 2>(feed, S[0]);  // row r+P+2 has landed
        } else {
            feed_get<kRingStages - 2>(feed, fin);   // row r+P+2; becomes L of the next tick
        }
        ring_group<T, MODE, VXS, VYS, U, P + 1>(a, ln, lane_store_all, r, ya, yb, req_stop, feed, dst, S);
    }
}

template <int T, int MODE, int VXS, int VYS>
__global__ void __launch_bounds__(32 * kRingWarpsPerBlock, kRingBlocksPerSM) k_step_ring(const __grid_constant__ TbArgs a) {
    static_assert(T >= 1 && T <= kTbMaxT, "T out of range");
    static_assert((kRingStages & (kRingStages - 1)) == 0 && kRingStages >= 4, "stages: power of two, >= 4");
    constexpr int N = ring_slots(T);
    extern __shared__ __align__(16) unsigned char ring_smem[];
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kRingWarpsPerBlock + (threadIdx.x >> 5);
    if (item >= a.n_items) return;  // warp-uniform

    int strip, ya, yb;
    if (!tb_item_map(a, item, strip, ya, yb)) return;
    const int xb = strip * kTbWout - kTbHX;
    TbLane ln;
    ln.x0 = xb + lane * kTbCells;
    const bool lane_store_all = lane >= 1 && lane <= 30 && ln.x0 >= a.sx0 && ln.x0 + 3 < a.sx1;
    const bool lane_partial = !lane_store_all && lane >= 1 && lane <= 30 && ln.x0 + 3 >= a.sx0 && ln.x0 < a.sx1;
    const bool strip_fast =
        xb >= a.fx0 && xb + kTbWidth <= a.fx1 && __ballot_sync(0xffffffffu, lane_partial) == 0u;
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

    double S[N][4];
#pragma unroll
    for (int q = 0; q < N; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) S[q][i] = 0.0;

    // rows ya-T … yb+T-1 of level 0 enter, rows ya … yb-1 of level T leave
    int r = ya - T;
    const int r_end = yb + T;
    double* dst = a.out + static_cast<long long>(r - T) * a.pitch + ln.x0;
    RingFeed feed;
    feed.base = static_cast<unsigned>(__cvta_generic_to_shared(ring_smem)) +
                static_cast<unsigned>(threadIdx.x >> 5) * (kRingStages * 1024u);
    feed.lane = static_cast<unsigned>(lane);
    feed.req_stage = 0;
    feed.get_stage = 0;
    feed.src = a.u + static_cast<long long>(r) * a.pitch + xb + lane * 2;
    feed.ok0 = xb + lane * 2 + 1 < a.xmax_load;
    feed.ok1 = xb + 64 + lane * 2 + 1 < a.xmax_load;
    // The last row any tick of this chunk consumes is r_end + 1 (<= ny + T + 2 < ny + kLeadY, inside the
    // allocation); the look-ahead would run kRingStages rows further, so requests stop there (a
    // request with size 0 reads nothing and zero-fills its stage).
    const int req_stop = r_end + 2;
    int next_row = r;
#pragma unroll
    for (int q = 0; q < kRingStages; ++q) {
        feed_request(feed, a.pitch, next_row < req_stop);
        ++next_row;
    }
    feed_get<kRingStages - 1>(feed, S[1]);  // c0 = row r
    feed_get<kRingStages - 2>(feed, S[0]);  // L  = row r+1

    // a tick at row r produces rows r-T … r-1; a group of U ticks is hot when all of them are interior
    // rows of a strip without boundary columns and the group fits into the chunk
    constexpr int U = CSIM_RING_U < N ? CSIM_RING_U : N;
    const int hot_lo = a.fy0 + T;
    const int hot_hi = min(a.fy1 - (U - 1), r_end - U);
    while (r < r_end) {
        while (strip_fast && r >= hot_lo && r <= hot_hi) {
            ring_group<T, MODE, VXS, VYS, U, 0>(a, ln, lane_store_all, r, ya, yb, req_stop, feed, dst, S);
            r += U;
            next_row += U;
        }
        if (r >= r_end) break;
        // one generic tick at phase 0, then the ring is rotated by one slot with register moves so
        // that the roles are back at phase 0
        ring_levels<T, MODE, VXS, VYS, 0, true>(a, ln, r, S);
        tb_store_row(a, ln, lane, lane_store_all, r - T, ya, yb, S[N - 1]);
        dst += a.pitch;
        feed_request(feed, a.pitch, next_row < req_stop);  // row r+kRingStages
        ++next_row;
#pragma unroll
        for (int q = N - 1; q > 0; --q)
#pragma unroll
            for (int i = 0; i < 4; ++i) S[q][i] = S[q - 1][i];
        feed_get<kRingStages - 2>(feed, S[0]);  // row r+2 has landed
        r += 1;
    }
    ring_wait<0>();  // nothing of this warp is in flight when its shared memory is handed on
}

}  // namespace csim
