// step_ring.cuh — the hot kernel, second generation: same strip/chunk decomposition, same row
// routines and arithmetic as step_tb.cuh, but ONE row per tick and a rotating ring of row slots.
//
// Why: k_step_tb (two rows per tick) keeps four rows per time level in registers (96 registers of
// state at T = 3, 162-168 in total → 12 warps per SM) and its FP64 pipe sat at 79 % of the active
// cycles with "wait" (fixed-latency dependency) as the top stall (profiles/r01b_*).  Here a warp
// holds 2T + 3 row slots in all:
//     L1, L0       the two level-0 rows in flight from HBM (requested one and two ticks ago)
//     c0           the newest level-0 row
//     b_k, a_k     the two older rows of level k  (k = 0 … T-1)
// A tick at level-0 row r computes, for k = 0 … T-1, row r-1-k of level k+1 from (a_k, b_k, c_k) and
// writes it INTO a_k's registers — a_k (the "south" row) is dead once the update has read it — where
// it serves as c_{k+1} for the next level of the same tick and as b_{k+1}, a_{k+1} in the two ticks
// after.  The last level's result, row r-T of level T, is stored and its slot receives the load of
// row r+3.  Net effect: from one tick to the next every role moves to the neighbouring slot, a pure
// rotation of the ring, so N = 2T+3 consecutive ticks are unrolled with compile-time slot numbers and
// no register is ever moved.  State: 8(2T+3) registers — 72 at T = 3 — under 128 registers per
// thread: 16 warps per SM instead of 12.
//
// Ticks that touch a boundary (edge strips, first/last rows of the tile, chunk ends that do not fill
// a group of N) run a single generic tick (GEN = true rows of step_tb.cuh) at phase 0 and then
// rotate the ring with register moves; only the unrolled groups are hot.
#pragma once
#include "step_tb.cuh"

namespace csim {

constexpr int kRingBlocksPerSM = 4;  // 128 registers per thread, 16 warps per SM

__host__ __device__ constexpr int ring_slots(int T) { return 2 * T + 3; }

// slot of a role at phase P of the unrolled group
// roles: 0 = L1, 1 = L0 (rows in flight), 2 = c0, 3+2k = b_k, 4+2k = a_k; c_k (k > 0) sits in a_{k-1}'s slot
template <int N>
__host__ __device__ constexpr int ring_slot(int role, int P) {
    return ((role - P) % N + N) % N;
}

// The T level-rows of one tick (see the file header); leaves row r-T of level T in role N-1's slot.
template <int T, int MODE, int VXS, int VYS, int P, bool GEN>
__device__ __forceinline__ void ring_levels(const TbArgs& a, const TbLane& ln, int r, double (&S)[2 * T + 3][4]) {
    constexpr int N = ring_slots(T);
#pragma unroll
    for (int k = 0; k < T; ++k) {
        double(&s)[4] = S[ring_slot<N>(4 + 2 * k, P)];  // a_k : row r-2-k of level k
        double(&c)[4] = S[ring_slot<N>(3 + 2 * k, P)];  // b_k : row r-1-k
        double(&n)[4] = S[ring_slot<N>(2 + 2 * k, P)];  // c_k : row r-k
        double res[4];
        tb_row<MODE, VXS, VYS, GEN>(a, ln, r - 1 - k, s, c, n, res);
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = res[i];      // c_{k+1} replaces a_k
    }
}

// U consecutive hot ticks (phases 0 … U-1) starting at level-0 row r, then — if U < N — a rotation of
// the ring by U slots with register moves, which puts the roles back at phase 0.  U = N needs no
// moves but unrolls N ticks (30 KB of code at T = 3, which thrashed the instruction cache:
// stall_no_instructions 14 % of samples); U = 3 costs 8N/U moves per tick on the integer pipes and
// keeps the loop at a third of that size.  Hot ticks run only where every lane stores all four
// cells or none and every row involved is an interior row inside the allocation: one predicated
// 256-bit store and one unconditional 256-bit load per tick.  The load of row r+3 goes into the slot
// the store just freed and is first read two ticks later; in the last tick of a group it is issued
// after the rotation so that the moves only touch rows requested at least one tick earlier.
template <int T, int MODE, int VXS, int VYS, int U, int P>
__device__ __forceinline__ void ring_group(const TbArgs& a, const TbLane& ln, bool lane_store_all, int r, int ya,
                                           int yb, const double*& src, double*& dst, double (&S)[2 * T + 3][4]) {
    constexpr int N = ring_slots(T);
    if constexpr (P < U) {
        ring_levels<T, MODE, VXS, VYS, P, false>(a, ln, r + P, S);
        double(&fin)[4] = S[ring_slot<N>(N - 1, P)];
        if (lane_store_all && r + P - T >= ya && r + P - T < yb) tb_store4(dst, fin);
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
            tb_load4(src, true, S[0]);
        } else {
            tb_load4(src, true, fin);
        }
        if (r + P + 3 + a.pf_rows < a.row_limit) tb_prefetch_l2(src + a.pf_off);
        src += a.pitch;
        dst += a.pitch;
        ring_group<T, MODE, VXS, VYS, U, P + 1>(a, ln, lane_store_all, r, ya, yb, src, dst, S);
    }
}

#ifndef CSIM_RING_U
#define CSIM_RING_U 3
#endif

template <int T, int MODE, int VXS, int VYS>
__global__ void __launch_bounds__(32 * kTbWarpsPerBlock, kRingBlocksPerSM) k_step_ring(const __grid_constant__ TbArgs a) {
    static_assert(T >= 1 && T <= kTbMaxT, "T out of range");
    constexpr int N = ring_slots(T);
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kTbWarpsPerBlock + (threadIdx.x >> 5);
    if (item >= a.n_items) return;  // warp-uniform

    // work item → (strip, first row, row count): identical to k_step_tb (the slower edge-strip items
    // come first so that they never form the tail of the launch)
    int strip, ya, h;
    {
        const int n_edge = a.nstrips >= 2 ? 2 : 1;
        const int n_int = a.nstrips - n_edge;
        if (item < a.n_edge_items) {
            strip = (item % n_edge) ? a.nstrips - 1 : 0;
            h = (a.chunk_h + a.edge_split - 1) / a.edge_split;
            ya = a.sy0 + (item / n_edge) * h;
        } else {
            const int e = item - a.n_edge_items;
            strip = 1 + e % n_int;
            const int ci = e / n_int;
            const int chunk = a.frame_pair ? (ci ? a.nchunks - 1 : 0) : a.int_chunk0 + ci;
            // the first n_main chunks are chunk_h rows tall, the rest (the tail of the launch) chunk_h2
            const bool tail = chunk >= a.n_main;
            h = tail ? a.chunk_h2 : a.chunk_h;
            ya = a.sy0 + (tail ? a.n_main * a.chunk_h + (chunk - a.n_main) * a.chunk_h2 : chunk * a.chunk_h);
        }
    }
    if (ya >= a.sy1) return;
    const int yb = min(ya + h, a.sy1);
    const int xb = strip * kTbWout - kTbHX;
    TbLane ln;
    ln.x0 = xb + lane * kTbCells;
    const bool can_load = ln.x0 + 3 < a.xmax_load;
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
    const double* src = a.u + static_cast<long long>(r) * a.pitch + ln.x0;
    double* dst = a.out + static_cast<long long>(r - T) * a.pitch + ln.x0;
    tb_load4(src, can_load, S[2]);                // c0 = row r
    tb_load4(src + a.pitch, can_load, S[1]);      // L0 = row r+1
    tb_load4(src + 2 * a.pitch, can_load, S[0]);  // L1 = row r+2
    src += 3 * a.pitch;

    // a tick at row r produces rows r-T … r-1; a group of U ticks is hot when all of them are interior
    // rows of a strip without boundary columns and the group fits into the chunk
    constexpr int U = CSIM_RING_U < N ? CSIM_RING_U : N;
    const int hot_lo = a.fy0 + T;
    const int hot_hi = min(a.fy1 - (U - 1), r_end - U);
    while (r < r_end) {
        while (strip_fast && r >= hot_lo && r <= hot_hi) {
            ring_group<T, MODE, VXS, VYS, U, 0>(a, ln, lane_store_all, r, ya, yb, src, dst, S);
            r += U;
        }
        if (r >= r_end) break;
        // one generic tick at phase 0, then the ring is rotated by one slot with register moves so
        // that the roles are back at phase 0.  The load is issued after the rotation: the moves then
        // touch only rows that were requested at least one whole tick ago.
        ring_levels<T, MODE, VXS, VYS, 0, true>(a, ln, r, S);
        tb_store_row(a, ln, lane, lane_store_all, r - T, ya, yb, S[N - 1]);
#pragma unroll
        for (int q = N - 1; q > 0; --q)
#pragma unroll
            for (int i = 0; i < 4; ++i) S[q][i] = S[q - 1][i];
        tb_load4(src, can_load, S[0]);  // row r+3
        if (can_load && r + 3 + a.pf_rows < a.row_limit) tb_prefetch_l2(src + a.pf_off);
        src += a.pitch;
        dst += a.pitch;
        r += 1;
    }
}

}  // namespace csim
