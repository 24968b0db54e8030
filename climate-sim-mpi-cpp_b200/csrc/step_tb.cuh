// step_tb.cuh — the blocked sweep: fused diffusion+advection, T time steps per pass over HBM.
// Holds the row routines, the work-item map and the register-only kernel k_step_tb; the default kernel
// k_step_tbs (step_tbs.cuh) shares all of it and differs only in how the level-0 rows reach the warp
// (TMA bulk copies into a shared-memory ring instead of register loads), which buys a fourth time level.
//
// Design: one WARP owns a strip of 128 columns (4 cells per lane, one 256-bit
// load per lane per row) and streams down a chunk of rows.  For every time level k < T the warp
// keeps the two most recent rows of that level in registers; when row r of level 0 arrives from
// HBM it produces row r-1 of level 1, from that row r-2 of level 2, … and finally stores row r-T of
// level T.  x-neighbours cross lanes with two 64-bit shuffles per level-row; nothing goes through
// shared memory and warps never synchronise with each other.  Each level loses one column of
// validity on both strip edges, so a strip yields 120 finished columns out of 128 (T <= 4) and a
// chunk re-computes T rows above and below itself.  Every cell is read from HBM once and written
// once per T steps: 16/T bytes per cell update instead of 16.
//
// Boundaries are folded into the sweep (no separate apply_boundary pass, no std::copy):
//   - cells outside [xlo,xhi) x [ylo,yhi) keep their value from level to level (that is what
//     "Periodic" means in the reference: ghosts are frozen), except that ghost
//     cells of a physical Dirichlet/Neumann side take the value apply_boundary (boundary.cpp:23-53)
//     would have written before the step, so the ring of `out` is what diffusion.cpp:18-25 copies;
//   - cells on the first/last interior line of a physical side read `value` (Dirichlet) or their
//     own value (Neumann mirror) instead of the neighbour.
// These fix-ups run only in the two edge strips and in the first/last row of the tile; interior
// level-rows take a branch-free fast path, and interior rows of strips that merely touch a frozen-ghost
// or neighbour side take the fast path plus a per-cell keep/advance select (TICK_XMASK).
//
// MODE_DIV (a spacing that is not a power of two) keeps the reference's divisions; the divisors are
// constants of a run, so each quotient is formed from the reciprocal and proved equal to the IEEE quotient
// (div_by_const_try, step_math.cuh), and a level-row with an unproved quotient is redone with __ddiv_rn.
//
// Arithmetic: tb_update issues exactly the reference's operations in the reference's order with
// non-contractible round-to-nearest intrinsics (see step_math.cuh); MODE_UNIT drops the four
// multiplications by 1.0 (dx = dy = 1), which are exact identities.
#pragma once
#include <cstdint>

#include "step_math.cuh"

namespace csim {

constexpr int kTbCells = 4;                        // cells per lane
constexpr int kTbWidth = 32 * kTbCells;            // 128 columns per strip
constexpr int kTbHX = 4;                           // columns discarded on each strip edge
constexpr int kTbWout = kTbWidth - 2 * kTbHX;      // 120 finished columns per strip
constexpr int kTbMaxT = 4;                         // T <= kTbHX
constexpr int kTbWarpsPerBlock = 4;
constexpr int kTbBlocksPerSM = 3;                  // 168 registers per thread, 12 warps per SM

enum { MODE_UNIT = 0, MODE_RECIP = 1, MODE_DIV = 2 };
// internal to the row routines: MODE_DIV's divisions by the checked reciprocal sequence (step_math.cuh), the
// verdict of the checks AND-ed into tb_update's `ok`; a row whose checks did not all pass is redone in MODE_DIV_IEEE
enum { MODE_DIV_IEEE = 3 };

struct TbArgs {
    const double* u;   // level-0 field, pointer to interior cell (0,0)
    double* out;       // level-T field, pointer to interior cell (0,0)
    long long out_minus_u;  // byte distance from a cell of u to the same cell of out
    long long pitch;   // doubles per row
    int nx, ny;        // interior size of the tile
    int xlo, xhi, ylo, yhi;  // cells advanced by the stencil: xlo<=x<xhi, ylo<=y<yhi
    int sx0, sx1, sy0, sy1;  // cells stored (ghost lines of physical sides included)
    int fx0, fx1, fy0, fy1;  // cells that need no boundary fix-up (fast path)
    int nstrips;       // strips across x
    int nchunks;       // chunks down y for interior strips
    int chunk_h;       // rows per chunk (interior strips); edge strips use chunk_h / edge_split
    int n_main;        // interior strips: chunks 0 … n_main-1 are chunk_h rows tall,
    int chunk_h2;      // chunks n_main … nchunks-1 are chunk_h2 rows tall (finer grain for the last round)
    int edge_split;
    int n_items;       // total (strip, chunk) work items of this launch
    int n_edge_items;  // of which the first n_edge_items belong to the edge strips (0: none in this launch)
    int int_chunk0;    // interior strips: first chunk enumerated ...
    int frame_pair;    // ... or, if set, exactly the first and the last chunk of every interior strip
    int xmax_load;     // a lane may load its 4 cells iff x0+3 < xmax_load (row allocation bound)
    int pf_rows;       // L2 prefetch distance in rows (off: a huge distance, which no row guard passes)
    long long pf_off;  // pf_rows * pitch
    // coupled launch of the multi-GPU loop (part TB_COUPLED): the frame items come first, [0, n_frame_items),
    // wait for the halo of this block before they read a ghost line and count a ticket down when they are done
    int n_frame_items;      // 0: no coupling (every other part)
    int n_frame_tickets;    // frame items that have rows (the ones that will take a ticket)
    int n_frame_edge_items; // of the frame items, how many belong to the edge strips
    unsigned seq;           // this block's sequence number
    unsigned* halo_flag;    // >= seq once the exchange of this block has landed in the ghost lines
    unsigned* done_flag;    // set to seq by the last frame item to finish: the bands of the next exchange are final
    unsigned* ticket;
    unsigned* err;
    unsigned long long timeout_ns;
    int phys;          // bit s: side s (left,right,bottom,top) is a physical boundary
    int bcL, bcR, bcB, bcT;
    double value;      // Dirichlet value
    StepK k;
};

// Work item → (strip, rows [ya, yb)).  Shared by both sweep kernels and by the host-side plan
// (csim_sweep_plan), which is how the geometry is tested without a GPU.  Interior strips get chunks of
// chunk_h rows (chunk_h2 in the tail of the launch); the slower edge strips get edge_split times as
// many chunks of chunk_h/edge_split rows and are enumerated first so that they never form the tail.
// Returns false for an item without rows.
__host__ __device__ __forceinline__ bool tb_item_decode(const TbArgs& a, int item, int n_edge_items, int int_chunk0,
                                                        int frame_pair, int& strip, int& ya, int& yb) {
    const int n_edge = a.nstrips >= 2 ? 2 : 1;
    const int n_int = a.nstrips - n_edge;
    int h;
    if (item < n_edge_items) {
        strip = (item % n_edge) ? a.nstrips - 1 : 0;
        h = (a.chunk_h + a.edge_split - 1) / a.edge_split;
        ya = a.sy0 + (item / n_edge) * h;
    } else {
        const int e = item - n_edge_items;
        strip = 1 + e % n_int;
        const int ci = e / n_int;
        const int chunk = frame_pair ? (ci ? a.nchunks - 1 : 0) : int_chunk0 + ci;
        const bool tail = chunk >= a.n_main;
        h = tail ? a.chunk_h2 : a.chunk_h;
        ya = a.sy0 + (tail ? a.n_main * a.chunk_h + (chunk - a.n_main) * a.chunk_h2 : chunk * a.chunk_h);
    }
    yb = ya + h < a.sy1 ? ya + h : a.sy1;
    return ya < a.sy1;
}
__host__ __device__ __forceinline__ bool tb_item_map(const TbArgs& a, int item, int& strip, int& ya, int& yb) {
    if (a.n_frame_items > 0) {  // coupled launch: the frame's items in the frame's order, then the interior's
        if (item < a.n_frame_items)
            return tb_item_decode(a, item, a.n_frame_edge_items, 0, a.frame_pair, strip, ya, yb);
        return tb_item_decode(a, item - a.n_frame_items, 0, 1, 0, strip, ya, yb);
    }
    return tb_item_decode(a, item, a.n_edge_items, a.int_chunk0, a.frame_pair, strip, ya, yb);
}

// Coupled launches: a frame item may not read a ghost line before the exchange of this block has landed
// (flag written on the exchange stream, normally long before the launch), and the last frame item to finish
// releases the next exchange.  Both are warp-uniform and outside every loop.
__device__ __forceinline__ void tb_frame_enter(const TbArgs& a, int item, int lane) {
    if (item >= a.n_frame_items) return;
    if (lane == 0) {
        const volatile unsigned* f = a.halo_flag;
        if (static_cast<int>(*f - a.seq) < 0) {
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (static_cast<int>(*f - a.seq) < 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > a.timeout_ns) {
                    *reinterpret_cast<volatile unsigned*>(a.err) = 201u;
                    __threadfence_system();
                    __trap();
                }
                __nanosleep(200);
            }
        }
        __threadfence();
    }
    __syncwarp();
}
__device__ __forceinline__ void tb_frame_leave(const TbArgs& a, int item, int lane) {
    if (item >= a.n_frame_items) return;
    __threadfence();  // this item's cells are visible device-wide before its ticket
    __syncwarp();
    if (lane == 0 && atomicAdd(a.ticket, 1u) == static_cast<unsigned>(a.n_frame_tickets) - 1u) {
        *a.ticket = 0;  // re-armed for the next block (launches of one stream do not overlap)
        __threadfence();
        *reinterpret_cast<volatile unsigned*>(a.done_flag) = a.seq;
    }
}

// One cell update.  VXS / VYS select the upwind side: +1 for v >= 0 (backward difference), -1 for
// v < 0 (forward difference), 0 for a velocity component that is exactly +0.0, whose whole term is
// dropped.  Dropping is exact for finite fields without negative zeros: (+0)*d is a signed zero, q + (±0)
// equals q unless q is itself a zero, and the sign of a zero `adv` can only change the result when
// o == -0, which needs c == -0 (x + y is -0 only for (-0) + (-0)); by the same argument an update never
// creates a -0 where there was none.  The dispatcher (kernels.cu, zero_terms_allowed) uses the 0
// variants only for fields it has scanned (finite, no -0, on every rank) and a stable time step.
template <int MODE, int VXS, int VYS>
__device__ __forceinline__ double tb_update(double c, double w, double e, double s, double n, const StepK& k,
                                            bool& ok) {
    // e - 2.0*c and n - 2.0*c as one FMA each: 2.0*c is exact (a power-of-two scaling), so
    // fma(-2, c, e) rounds the same real number as the reference's (e - 2.0*c) — one rounding either
    // way — as long as 2|c| does not overflow (|c| < 2^1023; csim_field_health reports max|u|).
    // Every other product of the update is rounded separately in the reference and stays unfused.
    double lx = __dadd_rn(__fma_rn(-2.0, c, e), w);
    double ly = __dadd_rn(__fma_rn(-2.0, c, n), s);
    if (MODE == MODE_RECIP) {
        lx = __dmul_rn(lx, k.rdx2);
        ly = __dmul_rn(ly, k.rdy2);
    } else if (MODE == MODE_DIV) {
        bool o1, o2;
        lx = div_by_const_try(lx, k.dx2, k.rdx2, o1);
        ly = div_by_const_try(ly, k.dy2, k.rdy2, o2);
        ok = ok && o1 && o2;
    } else if (MODE == MODE_DIV_IEEE) {
        lx = __ddiv_rn(lx, k.dx2);
        ly = __ddiv_rn(ly, k.dy2);
    }
    const double o = __dadd_rn(c, __dmul_rn(k.dtD, __dadd_rn(lx, ly)));
    if (VXS == 0 && VYS == 0) return o;
    double px = 0.0, py = 0.0;
    if (VXS != 0) {
        double ddx = VXS > 0 ? __dsub_rn(c, w) : __dsub_rn(e, c);
        if (MODE == MODE_RECIP) ddx = __dmul_rn(ddx, k.rdx);
        if (MODE == MODE_DIV) {
            bool o;
            ddx = div_by_const_try(ddx, k.dx, k.rdx, o);
            ok = ok && o;
        }
        if (MODE == MODE_DIV_IEEE) ddx = __ddiv_rn(ddx, k.dx);
        px = __dmul_rn(k.vx, ddx);
    }
    if (VYS != 0) {
        double ddy = VYS > 0 ? __dsub_rn(c, s) : __dsub_rn(n, c);
        if (MODE == MODE_RECIP) ddy = __dmul_rn(ddy, k.rdy);
        if (MODE == MODE_DIV) {
            bool o;
            ddy = div_by_const_try(ddy, k.dy, k.rdy, o);
            ok = ok && o;
        }
        if (MODE == MODE_DIV_IEEE) ddy = __ddiv_rn(ddy, k.dy);
        py = __dmul_rn(k.vy, ddy);
    }
    const double adv = VXS == 0 ? py : (VYS == 0 ? px : __dadd_rn(px, py));
    return __dadd_rn(o, __dmul_rn(k.ndt, adv));
}

// Dirichlet → value, Neumann → mirror, Periodic → keep
__device__ __forceinline__ double bc_pick(int bc, double value, double mirror, double keep) {
    return bc == 0 ? value : (bc == 1 ? mirror : keep);
}

// Per-lane, per-work-item constants of the boundary fix-ups (only read on the general path).
struct TbLane {
    int x0;        // x of the lane's cell 0 (always a multiple of 4)
    int inx;       // bit i: cell i lies in [xlo, xhi) and is advanced by the stencil
    bool at_l;     // cell 0 is x == 0 on a physical left side      → its west neighbour is the BC
    bool ghost_l;  // cell 3 is x == -1 on a physical left side     → ghost rule
    int at_r;      // index of the cell with x == nx-1 on a physical right side, or -1
    int ghost_r;   // index of the cell with x == nx   on a physical right side, or -1
};

// The three flavours of a tick.  TICK_FAST: the plain stencil on all four cells (interior strips,
// interior rows).  TICK_XMASK: interior rows of a strip that touches a side whose ghost cells are
// merely frozen or advanced like interior cells (a "periodic" physical side, SURVEY.md Q1/Q2, or a
// side with a neighbour): the plain stencil, then cells outside [xlo, xhi) keep their value.
// TICK_GEN: every boundary rule of the reference (see the file header); all conditions on j are
// warp-uniform, all conditions on x are per-lane selects, so the four stencils still interleave.
enum { TICK_FAST = 0, TICK_XMASK = 1, TICK_GEN = 2 };

// One row of one level.  s/c/n are rows j-1, j, j+1 of the current level; the result is row j of
// the next level.
template <int MODE, int VXS, int VYS, int KIND>
__device__ __forceinline__ void tb_row(const TbArgs& a, const TbLane& ln, int j, const double (&s)[4],
                                       const double (&c)[4], const double (&n)[4], double (&res)[4]) {
    const double w0 = __shfl_up_sync(0xffffffffu, c[3], 1);
    const double e3 = __shfl_down_sync(0xffffffffu, c[0], 1);
    if (KIND != TICK_GEN) {
        double r[4];
        bool ok = true;
        r[0] = tb_update<MODE, VXS, VYS>(c[0], w0, c[1], s[0], n[0], a.k, ok);
        r[1] = tb_update<MODE, VXS, VYS>(c[1], c[0], c[2], s[1], n[1], a.k, ok);
        r[2] = tb_update<MODE, VXS, VYS>(c[2], c[1], c[3], s[2], n[2], a.k, ok);
        r[3] = tb_update<MODE, VXS, VYS>(c[3], c[2], e3, s[3], n[3], a.k, ok);
        if (MODE == MODE_DIV && __any_sync(0xffffffffu, !ok)) {  // rare: a quotient the check could not prove
            r[0] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[0], w0, c[1], s[0], n[0], a.k, ok);
            r[1] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[1], c[0], c[2], s[1], n[1], a.k, ok);
            r[2] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[2], c[1], c[3], s[2], n[2], a.k, ok);
            r[3] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[3], c[2], e3, s[3], n[3], a.k, ok);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) res[i] = (KIND == TICK_FAST || ((ln.inx >> i) & 1)) ? r[i] : c[i];
        return;
    }
    const bool physB = a.phys & 4, physT = a.phys & 8;
    if (j < a.ylo || j >= a.yhi) {  // warp-uniform: the row is not advanced (ghost row, padding)
        const bool gb = physB && j == -1, gt = physT && j == a.ny;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double r = c[i];
            const bool in = (ln.inx >> i) & 1;
            // ghost rows of physical Dirichlet/Neumann sides take what apply_boundary wrote into u
            // before this step (diffusion_step then copies it into out's ring)
            if (gb && in) r = bc_pick(a.bcB, a.value, n[i], r);
            if (gt && in) r = bc_pick(a.bcT, a.value, s[i], r);
            res[i] = r;
        }
        return;
    }
    // first/last interior row of a physical side: the neighbour row is the boundary value
    double ss[4], nn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ss[i] = s[i];
        nn[i] = n[i];
    }
    if (physB && j == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) ss[i] = bc_pick(a.bcB, a.value, c[i], s[i]);
    }
    if (physT && j == a.ny - 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) nn[i] = bc_pick(a.bcT, a.value, c[i], n[i]);
    }
    // x direction: only a few (lane, cell) pairs differ from the plain stencil
    double ww0 = w0;
    double ee[4] = {c[1], c[2], c[3], e3};
    double keep[4] = {c[0], c[1], c[2], c[3]};  // value of cells that are not advanced (frozen)
    if (ln.at_l) ww0 = bc_pick(a.bcL, a.value, c[0], w0);         // x == 0: west neighbour is the BC
    if (ln.ghost_l) keep[3] = bc_pick(a.bcL, a.value, e3, c[3]);  // x == -1: ghost rule
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i == ln.at_r) ee[i] = bc_pick(a.bcR, a.value, c[i], ee[i]);  // x == nx-1: east neighbour is the BC
        if (i == ln.ghost_r) keep[i] = bc_pick(a.bcR, a.value, i == 0 ? w0 : c[i == 0 ? 0 : i - 1], c[i]);  // x == nx
    }
    double r[4];
    bool ok = true;
    r[0] = tb_update<MODE, VXS, VYS>(c[0], ww0, ee[0], ss[0], nn[0], a.k, ok);
    r[1] = tb_update<MODE, VXS, VYS>(c[1], c[0], ee[1], ss[1], nn[1], a.k, ok);
    r[2] = tb_update<MODE, VXS, VYS>(c[2], c[1], ee[2], ss[2], nn[2], a.k, ok);
    r[3] = tb_update<MODE, VXS, VYS>(c[3], c[2], ee[3], ss[3], nn[3], a.k, ok);
    if (MODE == MODE_DIV && __any_sync(0xffffffffu, !ok)) {
        r[0] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[0], ww0, ee[0], ss[0], nn[0], a.k, ok);
        r[1] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[1], c[0], ee[1], ss[1], nn[1], a.k, ok);
        r[2] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[2], c[1], ee[2], ss[2], nn[2], a.k, ok);
        r[3] = tb_update<MODE_DIV_IEEE, VXS, VYS>(c[3], c[2], ee[3], ss[3], nn[3], a.k, ok);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) res[i] = ((ln.inx >> i) & 1) ? r[i] : keep[i];
}

__device__ __forceinline__ void tb_load4(const double* p, bool ok, double (&v)[4]) {
    if (ok) {
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                     : "l"(p));
    } else {
        v[0] = v[1] = v[2] = v[3] = 0.0;
    }
}
__device__ __forceinline__ void tb_prefetch_l2(const double* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void tb_store4(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                 : "memory");
}

// Store row `jo` of level T (held in v) if it belongs to this work item.
__device__ __forceinline__ void tb_store_row(const TbArgs& a, const TbLane& ln, int lane, bool lane_store_all,
                                             int jo, int ya, int yb, const double (&v)[4]) {
    if (jo < ya || jo >= yb) return;  // warp-uniform
    double* dst = a.out + static_cast<long long>(jo) * a.pitch + ln.x0;
    const bool ring_row = jo < 0 || jo >= a.ny;
    if (lane_store_all && !ring_row) {
        tb_store4(dst, v);
    } else {
        const int lo = ring_row ? max(a.sx0, 0) : a.sx0;  // ring rows: no corners
        const int hi = ring_row ? min(a.sx1, a.nx) : a.sx1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = ln.x0 + i, pos = lane * kTbCells + i;
            const bool zone = (pos >= kTbHX && pos < kTbWidth - kTbHX) || x == -1 || x == a.nx;
            if (zone && x >= lo && x < hi) dst[i] = v[i];
        }
    }
}

// One tick: two new level-0 rows (r, r+1) enter, two finished rows (r-T, r-T+1) leave.
// st[k][slot][row][cell]: per level two slots of two rows.  In phase PH, slot PH of every level is its
// state (rows r-k-2, r-k-1) and slot 1-PH its incoming pair (rows r-k, r-k+1); afterwards the
// incoming pair is the state and slot PH is free for the next pair, so consecutive ticks alternate PH
// and no register is ever moved.
// Rows are counted twice: `r` is the absolute row (boundary rules of the general ticks), `q` = r-T-ya
// is the first finished row relative to the work item, which has `h` rows.  Fast ticks use only q and
// h (store window 0 <= q < h, request window q+2 < h) and derive the store address from the load
// pointer, so their loop carries neither r, ya, yb nor a second pointer.
template <int T, int MODE, int VXS, int VYS, int PH, int KIND>
__device__ __forceinline__ void tb_tick(const TbArgs& a, const TbLane& ln, int lane, bool lane_store_all,
                                        bool can_load, int r, int q, int h, int ya, int yb,
                                        const double*& src, double (&st)[T][2][2][4]) {
    double fin[2][4];
#pragma unroll
    for (int k = 0; k < T; ++k) {
        double(&A)[4] = st[k][PH][0];
        double(&B)[4] = st[k][PH][1];
        double(&C)[4] = st[k][1 - PH][0];
        double(&D)[4] = st[k][1 - PH][1];
        if (k + 1 < T) {
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k - 1, A, B, C, st[k + 1 < T ? k + 1 : k][1 - PH][0]);
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k, B, C, D, st[k + 1 < T ? k + 1 : k][1 - PH][1]);
        } else {
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k - 1, A, B, C, fin[0]);
            tb_row<MODE, VXS, VYS, KIND>(a, ln, r - k, B, C, D, fin[1]);
        }
        if (k == 0) {
            // rows r-2, r-1 of level 0 are dead now: request rows r+2, r+3 into their registers.  The
            // last level-0 row a work item needs is yb+T-1 <= ny+T (r+2 < yb+T  <=>  q+2 < h); past it the
            // fast ticks re-read rows 0 and 1 of their own columns instead — an address that is always
            // inside the allocation and resident in L2 — so that the load itself stays unconditional:
            // a predicated or branched-around load makes ptxas land the rows in temporaries and move them
            // within the same tick, which stalls the warp for a DRAM round trip (35 % of all stall
            // samples in profiles/r02a).  The values feed rows nobody stores.  Fast ticks run only in
            // strips that lie inside the row allocation with all 32 lanes (strip_fast implies
            // xb + 128 <= nx + T < xmax_load).
            const bool rows_needed = q + 2 < h;
            if (KIND != TICK_FAST) {
                tb_load4(src, can_load && rows_needed, A);
                tb_load4(src + a.pitch, can_load && rows_needed, B);
            } else {
                const double* ls = rows_needed ? src : a.u + ln.x0;
                tb_load4(ls, true, A);
                tb_load4(ls + a.pitch, true, B);
            }
            // pull the rows pf_rows ahead into L2 (costs no registers; never past the rows the item reads)
            if ((KIND != TICK_FAST ? can_load : true) && q + 3 + a.pf_rows < h) {
                tb_prefetch_l2(src + a.pf_off);
                tb_prefetch_l2(src + a.pf_off + a.pitch);
            }
            src += 2 * a.pitch;
        }
    }
    if (KIND == TICK_FAST) {
        // fast ticks run only in strips where every lane stores all four cells or none, and the rows
        // they finish are interior rows: one predicated 256-bit store per row, no per-cell tests.
        // src points at row r+4 of `u` by now; row r-T of `out` lies a fixed (warp-uniform) distance away.
        double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(const_cast<double*>(src)) + a.out_minus_u) -
                      (T + 4) * a.pitch;
        if (lane_store_all && q >= 0 && q < h) tb_store4(dst, fin[0]);
        if (lane_store_all && q + 1 >= 0 && q + 1 < h) tb_store4(dst + a.pitch, fin[1]);
    } else {
        tb_store_row(a, ln, lane, lane_store_all, r - T, ya, yb, fin[0]);
        tb_store_row(a, ln, lane, lane_store_all, r - T + 1, ya, yb, fin[1]);
    }
}

template <int T, int MODE, int VXS, int VYS>
__global__ void __launch_bounds__(32 * kTbWarpsPerBlock, kTbBlocksPerSM) k_step_tb(const __grid_constant__ TbArgs a) {
    static_assert(T >= 1 && T <= kTbMaxT, "T out of range");
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kTbWarpsPerBlock + (threadIdx.x >> 5);
    if (item >= a.n_items) return;  // warp-uniform

    int strip, ya, yb;
    if (!tb_item_map(a, item, strip, ya, yb)) return;
    tb_frame_enter(a, item, lane);
    const int xb = strip * kTbWout - kTbHX;
    TbLane ln;
    ln.x0 = xb + lane * kTbCells;
    const bool can_load = ln.x0 + 3 < a.xmax_load;
    const bool lane_store_all = lane >= 1 && lane <= 30 && ln.x0 >= a.sx0 && ln.x0 + 3 < a.sx1;
    // a lane that stores some but not all of its cells (the store range ends inside its group) needs
    // the per-cell store of the general path; fast strips have none
    const bool lane_partial = !lane_store_all && lane >= 1 && lane <= 30 && ln.x0 + 3 >= a.sx0 && ln.x0 < a.sx1;
    const bool strip_fast =
        xb >= a.fx0 && xb + kTbWidth <= a.fx1 && __ballot_sync(0xffffffffu, lane_partial) == 0u;
    // the strip touches no Dirichlet/Neumann column: its interior rows need the cell mask only
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

    double st[T][2][2][4];
#pragma unroll
    for (int k = 0; k < T; ++k)
#pragma unroll
        for (int q = 0; q < 16; ++q) st[k][q >> 3][(q >> 2) & 1][q & 3] = 0.0;

    int r = ya - T;
    const double* src = a.u + static_cast<long long>(r) * a.pitch + ln.x0;
    tb_load4(src, can_load, st[0][1][0]);
    tb_load4(src + a.pitch, can_load, st[0][1][1]);
    src += 2 * a.pitch;

    // A loop iteration is two ticks (phases 0 and 1) and touches rows r-T .. r+3.  Iterations whose rows
    // are all plain interior rows run as an inner loop of the strip's flavour (TICK_FAST or TICK_XMASK);
    // the fast one carries nothing but a row counter, the item height and the load pointer, and
    // everything the general ticks need (boundary flags, per-lane masks) stays out of its registers.
    const int h = yb - ya;
    const int r_end = yb + T;
    while (r < r_end) {
        if (mid != TICK_GEN && r - T >= a.fy0 && r + 2 < a.fy1) {
            // iterations until the first one that would touch row fy1 or lie beyond the chunk
            int n_it = (min(a.fy1 - 2, r_end) - r + 3) >> 2;
            if (mid == TICK_FAST) {
                int q = r - T - ya;
                r += 4 * n_it;
#pragma unroll 1
                for (; n_it > 0; --n_it, q += 4) {
                    tb_tick<T, MODE, VXS, VYS, 0, TICK_FAST>(a, ln, lane, lane_store_all, true, 0, q, h, 0, 0, src, st);
                    tb_tick<T, MODE, VXS, VYS, 1, TICK_FAST>(a, ln, lane, lane_store_all, true, 0, q + 2, h, 0, 0, src, st);
                }
            } else {
#pragma unroll 1
                for (; n_it > 0; --n_it, r += 4) {
                    tb_tick<T, MODE, VXS, VYS, 0, TICK_XMASK>(a, ln, lane, lane_store_all, can_load, r, r - T - ya, h, ya, yb, src, st);
                    tb_tick<T, MODE, VXS, VYS, 1, TICK_XMASK>(a, ln, lane, lane_store_all, can_load, r + 2, r + 2 - T - ya, h, ya, yb, src, st);
                }
            }
        } else {
            tb_tick<T, MODE, VXS, VYS, 0, TICK_GEN>(a, ln, lane, lane_store_all, can_load, r, r - T - ya, h, ya, yb, src, st);
            tb_tick<T, MODE, VXS, VYS, 1, TICK_GEN>(a, ln, lane, lane_store_all, can_load, r + 2, r + 2 - T - ya, h, ya, yb, src, st);
            r += 4;
        }
    }
    tb_frame_leave(a, item, lane);
}

}  // namespace csim
