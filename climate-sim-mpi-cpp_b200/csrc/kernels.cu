// kernels.cu — the device kernels of the timestep path and their C-ABI launchers.
//
// Arithmetic contract (bit parity with the reference CPU build): every FP64
// operation is issued through __dadd_rn/__dsub_rn/__dmul_rn/__ddiv_rn, which round to nearest-even
// and are never contracted into FMA, in exactly the reference's order (the one deliberate FMA of the
// blocked sweep, e - 2.0*c, rounds the same real number once either way: step_tb.cuh, tb_update):
//   lap  = ((e - 2.0*c) + w) / (dx*dx) + ((n - 2.0*c) + s) / (dy*dy)      src/diffusion.cpp:12-13
//   o    = c + (dt*D) * lap                                               src/diffusion.cpp:14
//   dudx = vx>=0 ? (c - w)/dx : (e - c)/dx                                src/advection.cpp:16-20
//   dudy = vy>=0 ? (c - s)/dy : (n - c)/dy                                src/advection.cpp:23-27
//   o    = o + (-dt) * (vx*dudx + vy*dudy)                                src/advection.cpp:29-31
// Division by a power of two is replaced by multiplication with its exact reciprocal (same real
// value, same rounding); any other spacing keeps IEEE division unless CSIM_STEP_FAST_RECIP is set.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "csim_internal.hpp"
#include "step_math.cuh"
#include "step_tb_inst.cuh"

namespace csim {

// ---- reference-shaped single-function kernels (API parity, not the hot path) ------------------

// diffusion_step interior, src/diffusion.cpp:9-16.  One thread per cell; u/out are interior
// pointers (cell (0,0) = Field::at(h,h)).
template <bool kDiv>
__global__ void __launch_bounds__(256) k_diffusion(const double* __restrict__ u, double* __restrict__ out,
                                                   int nx, int ny, int64_t pitch, StepK k) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    const double* p = u + static_cast<int64_t>(y) * pitch + x;
    const double c = p[0];
    out[static_cast<int64_t>(y) * pitch + x] = diffusion_update<kDiv>(c, p[-1], p[1], p[-pitch], p[pitch], k);
}

// advection_step, src/advection.cpp:13-33: out += (-dt)*adv
template <bool kDiv>
__global__ void __launch_bounds__(256) k_advection(const double* __restrict__ u, double* __restrict__ out,
                                                   int nx, int ny, int64_t pitch, StepK k) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    const double* p = u + static_cast<int64_t>(y) * pitch + x;
    const double c = p[0];
    double* q = out + static_cast<int64_t>(y) * pitch + x;
    *q = __dadd_rn(*q, advection_increment<kDiv>(c, p[-1], p[1], p[-pitch], p[pitch], k));
}

// Outermost ring of out := that of u, src/diffusion.cpp:18-25.  Pointers address padded cell
// (0,0); nxt/nyt are the padded sizes.
__global__ void k_ring_copy(const double* __restrict__ u, double* __restrict__ out, int nxt, int nyt,
                            int64_t pitch) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nxt) {
        out[t] = u[t];
        out[static_cast<int64_t>(nyt - 1) * pitch + t] = u[static_cast<int64_t>(nyt - 1) * pitch + t];
    }
    if (t < nyt) {
        out[static_cast<int64_t>(t) * pitch] = u[static_cast<int64_t>(t) * pitch];
        out[static_cast<int64_t>(t) * pitch + nxt - 1] = u[static_cast<int64_t>(t) * pitch + nxt - 1];
    }
}

// apply_boundary, left and right columns: src/boundary.cpp:23-37.  f addresses padded cell (0,0).
// mode: -1 skip (side has a neighbour, or Periodic), 0 Dirichlet, 1 Neumann.
__global__ void k_bc_columns(double* __restrict__ f, int nx, int ny, int h, int64_t pitch, int mode_l,
                             int mode_r, double value) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > h + ny) return;  // rows jB..jT = 0..h+ny
    double* row = f + static_cast<int64_t>(j) * pitch;
    if (mode_l == 0)
        row[0] = value;
    else if (mode_l == 1)
        row[0] = row[h];
    if (mode_r == 0)
        row[h + nx] = value;
    else if (mode_r == 1)
        row[h + nx] = row[h + nx - 1];
}
// bottom and top rows: src/boundary.cpp:39-53 (runs after the columns, so corners end up with the
// row rule and Neumann rows read the freshly written column ghosts, as in the reference).
__global__ void k_bc_rows(double* __restrict__ f, int nx, int ny, int h, int64_t pitch, int mode_b,
                          int mode_t, double value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx + 2 * h) return;  // i0..i1 = 0..nx_tot-1
    if (mode_b == 0)
        f[i] = value;
    else if (mode_b == 1)
        f[i] = f[static_cast<int64_t>(h) * pitch + i];
    double* top = f + static_cast<int64_t>(h + ny) * pitch;
    if (mode_t == 0)
        top[i] = value;
    else if (mode_t == 1)
        top[i] = f[static_cast<int64_t>(h + ny - 1) * pitch + i];
}

// ---- fused step, version 1: one sweep, one step ------------------------------------------------
// Each thread owns two x-adjacent cells (one 16-byte load) and walks kRowsV1 rows with the
// south/centre/north rows in registers, so every interior cell is loaded from HBM once; the two
// x-neighbours outside the pair are scalar loads that hit L1 (the adjacent thread's line).
// Writes the interior of `out` and the ghost ring of `out` := ghost ring of `u` (diffusion.cpp:18-25),
// which replaces the reference's std::copy(u → tmp) (main.cpp:104).
constexpr int kRowsV1 = 16;

template <bool kDiv>
__global__ void __launch_bounds__(256) k_step_v1(const double* __restrict__ u, double* __restrict__ out,
                                                 int nx, int ny, int64_t pitch, StepK k) {
    const int x = (blockIdx.x * 32 + threadIdx.x) * 2;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * kRowsV1;
    if (x >= nx || y0 >= ny) return;
    const int y1 = min(y0 + kRowsV1, ny);
    const bool has2 = x + 1 < nx;

    const double* p = u + static_cast<int64_t>(y0 - 1) * pitch + x;
    double2 s = *reinterpret_cast<const double2*>(p);
    if (y0 == 0 && x == 0) out[-pitch - 1] = p[-1];  // corner (-1,-1)
    if (y0 == 0 && (x + 2 == nx || !has2)) out[-pitch + nx] = has2 ? p[2] : s.y;  // corner (nx,-1)
    p += pitch;
    double2 c = *reinterpret_cast<const double2*>(p);
    double w = p[-1], e = p[2];
#pragma unroll 4
    for (int y = y0; y < y1; ++y) {
        const double* pn = p + pitch;
        const double2 n = *reinterpret_cast<const double2*>(pn);
        const double wn = pn[-1], en = pn[2];
        const double o0 = fused_update<kDiv>(c.x, w, c.y, s.x, n.x, k);
        double* q = out + static_cast<int64_t>(y) * pitch + x;
        if (has2) {
            const double o1 = fused_update<kDiv>(c.y, c.x, e, s.y, n.y, k);
            *reinterpret_cast<double2*>(q) = make_double2(o0, o1);
            if (x + 2 == nx) q[2] = e;  // right ghost column
        } else {
            q[0] = o0;
            q[1] = c.y;  // x == nx-1: the pair's second cell IS the right ghost
        }
        if (x == 0) q[-1] = w;  // left ghost column
        if (y == 0) {           // bottom ghost row
            if (has2)
                *reinterpret_cast<double2*>(q - pitch) = s;
            else
                q[-pitch] = s.x;
        }
        if (y == ny - 1) {  // top ghost row (+ its two corners)
            if (has2)
                *reinterpret_cast<double2*>(q + pitch) = n;
            else
                q[pitch] = n.x;
            if (x == 0) q[pitch - 1] = wn;
            if (has2 && x + 2 == nx) q[pitch + 2] = en;
            if (!has2) q[pitch + 1] = n.y;
        }
        s = c;
        c = n;
        w = wn;
        e = en;
        p = pn;
    }
}

// ---- reductions -----------------------------------------------------------------------------

__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// min/max over a w x hgt window starting at `f` (row pitch `pitch`); per-block partials to
// part[2*b], part[2*b+1].  min and max are exact and order-independent, so the tree shape does not
// matter for parity with std::min_element/std::max_element (main.cpp:74-75).
__global__ void __launch_bounds__(256) k_minmax_partial(const double* __restrict__ f, int w, int hgt,
                                                        int64_t pitch, double* __restrict__ part) {
    double lo = DBL_MAX, hi = -DBL_MAX;
    for (int j = blockIdx.x; j < hgt; j += gridDim.x) {
        const double* row = f + static_cast<int64_t>(j) * pitch;
        for (int i = threadIdx.x; i < w; i += blockDim.x) {
            const double v = row[i];
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
    }
    __shared__ double slo[8], shi[8];
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) {
        slo[threadIdx.x >> 5] = lo;
        shi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = threadIdx.x < (blockDim.x >> 5) ? slo[threadIdx.x] : DBL_MAX;
        hi = threadIdx.x < (blockDim.x >> 5) ? shi[threadIdx.x] : -DBL_MAX;
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (threadIdx.x == 0) {
            part[2 * blockIdx.x] = lo;
            part[2 * blockIdx.x + 1] = hi;
        }
    }
}
__global__ void __launch_bounds__(256) k_minmax_final(const double* __restrict__ part, int nblocks,
                                                      double* __restrict__ result) {
    double lo = DBL_MAX, hi = -DBL_MAX;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        lo = fmin(lo, part[2 * b]);
        hi = fmax(hi, part[2 * b + 1]);
    }
    __shared__ double slo[8], shi[8];
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) {
        slo[threadIdx.x >> 5] = lo;
        shi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = threadIdx.x < (blockDim.x >> 5) ? slo[threadIdx.x] : DBL_MAX;
        hi = threadIdx.x < (blockDim.x >> 5) ? shi[threadIdx.x] : -DBL_MAX;
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (threadIdx.x == 0) {
            result[0] = lo;
            result[1] = hi;
        }
    }
}

// max |u| and number of non-finite cells over the interior (stability diagnostic)
__global__ void __launch_bounds__(256) k_health(const double* __restrict__ f, int nx, int ny, int64_t pitch,
                                                double* __restrict__ max_abs_bits,
                                                unsigned long long* __restrict__ nonfinite) {
    double hi = 0.0;
    unsigned long long bad = 0;
    for (int j = blockIdx.x; j < ny; j += gridDim.x) {
        const double* row = f + static_cast<int64_t>(j) * pitch;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) {
            const double v = row[i];
            if (isfinite(v))
                hi = fmax(hi, fabs(v));
            else
                ++bad;
        }
    }
    hi = warp_max(hi);
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0) {
        // non-negative doubles order like their bit patterns → integer atomicMax is exact
        atomicMax(reinterpret_cast<unsigned long long*>(max_abs_bits),
                  static_cast<unsigned long long>(__double_as_longlong(hi)));
        if (bad) atomicAdd(nonfinite, bad);
    }
}

// Value scan behind resolve_zero_terms: counts the cells of the padded tile that are not finite, are
// 2^1000 or larger in magnitude, or are -0.0.  f addresses padded cell (0,0).
__global__ void __launch_bounds__(256) k_scan_values(const double* __restrict__ f, int nxt, int nyt, int64_t pitch,
                                                     unsigned long long* __restrict__ bad_out) {
    unsigned long long bad = 0;
    for (int j = blockIdx.x; j < nyt; j += gridDim.x) {
        const double* row = f + static_cast<int64_t>(j) * pitch;
        for (int i = threadIdx.x; i < nxt; i += blockDim.x) {
            const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(row[i]));
            // exponent field >= 1000 + 1023 covers inf and NaN as well
            if (((b >> 52) & 0x7ffull) >= 2023ull || b == 0x8000000000000000ull) ++bad;
        }
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(bad_out, bad);
}

StepK make_consts(double dx, double dy, double D, double vx, double vy, double dt, int flags, bool* use_div) {
    StepK k;
    k.dtD = dt * D;  // "dt * D * lap" groups as (dt*D)*lap, src/diffusion.cpp:14
    k.ndt = -dt;     // src/advection.cpp:31
    k.vx = vx;
    k.vy = vy;
    k.dx2 = dx * dx;  // src/diffusion.cpp:12
    k.dy2 = dy * dy;  // src/diffusion.cpp:13
    k.dx = dx;
    k.dy = dy;
    k.rdx2 = 1.0 / k.dx2;
    k.rdy2 = 1.0 / k.dy2;
    k.rdx = 1.0 / dx;
    k.rdy = 1.0 / dy;
    k.vx_pos = vx >= 0.0 ? 1 : 0;  // src/advection.cpp:16
    k.vy_pos = vy >= 0.0 ? 1 : 0;  // src/advection.cpp:23
    const bool exact = is_pow2(k.dx2) && is_pow2(k.dy2) && is_pow2(dx) && is_pow2(dy);
    *use_div = !(exact || (flags & CSIM_STEP_FAST_RECIP));
    return k;
}

int launch_boundary(csim_field* f, const int nbr[4], const int bc[4], double value) {
    csim_ctx* c = f->ctx;
    int mode[4];
    for (int s = 0; s < 4; ++s) {
        CSIM_REQUIRE(bc[s] >= 0 && bc[s] <= 2, CSIM_ERR_INVALID, "apply_boundary: unknown BC type");
        mode[s] = (nbr[s] == CSIM_PROC_NULL && bc[s] != CSIM_BC_PERIODIC) ? bc[s] : -1;
    }
    double* f00 = f->at(0, 0);
    if (mode[0] >= 0 || mode[1] >= 0) {
        const int n = f->h + f->ny + 1;
        CSIM_LAUNCH(c, k_bc_columns, (n + 255) / 256, 256, 0, f00, f->nx, f->ny, f->h, f->pitch, mode[0],
                    mode[1], value);
    }
    if (mode[2] >= 0 || mode[3] >= 0) {
        const int n = f->nxt();
        CSIM_LAUNCH(c, k_bc_rows, (n + 255) / 256, 256, 0, f00, f->nx, f->ny, f->h, f->pitch, mode[2],
                    mode[3], value);
    }
    return CSIM_OK;
}

int launch_step_v1(const csim_field* u, csim_field* out, const StepK& k, bool use_div) {
    csim_ctx* c = u->ctx;
    const dim3 block(32, 8);
    const dim3 grid((u->nx + 63) / 64, (u->ny + 8 * kRowsV1 - 1) / (8 * kRowsV1));
    if (use_div)
        CSIM_LAUNCH(c, k_step_v1<true>, grid, block, 0, u->interior(), out->interior(), u->nx, u->ny, u->pitch, k);
    else
        CSIM_LAUNCH(c, k_step_v1<false>, grid, block, 0, u->interior(), out->interior(), u->nx, u->ny, u->pitch,
                    k);
    return CSIM_OK;
}

// ---- temporally blocked sweep (step_tb.cuh): host-side geometry and dispatch ---------------------

// With fewer than 3 chunks (or no interior strip) there is nothing to overlap with the exchange.
bool tb_split_pointless(int nchunks, int n_int) { return nchunks < 3 || n_int < 1; }

int tb_max_T() {
    static int cached = -1;
    if (cached < 0) {
        // the staged sweep (step_tbs.cuh) runs four levels in the registers the register-only sweep needs
        // for three; the latter spills at T = 4 and measures slower (profiles/r01_tb_tuning.md)
        const char* k = std::getenv("CSIM_TB_KERNEL");
        cached = (k && std::strcmp(k, "reg") == 0) ? 3 : 4;
        if (const char* e = std::getenv("CSIM_TB_MAXT")) {
            const int v = std::atoi(e);
            if (v >= 1 && v <= kTbMaxT) cached = v;
        }
    }
    return cached;
}
static int tb_env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
int tb_carveout_env() {
    static const int v = tb_env_int("CSIM_CARVEOUT", -1);
    return v;
}
// Blocking depth with IEEE division (non-power-of-two spacing).  The divisions make the sweep
// compute-bound already at one step per sweep, so blocking in time only adds the re-computed halo
// cells; measured in profiles/r02_tuning.md.  CSIM_TB_DIV_MAXT = 2 or 3 opts in.
int tb_max_T_div() {
    static const int v = [] {
        const int e = tb_env_int("CSIM_TB_DIV_MAXT", 1);
        return e >= 1 && e <= 3 ? e : 1;
    }();
    return v < tb_max_T() ? v : tb_max_T();
}

// Advance `u` by T steps into `out` in one sweep.  nbr/bc as in csim_step_params.  Sides with a
// neighbour must hold T valid ghost lines (csim::wide_exchange; T == 1: csim_halo_exchange).
// part: TB_ALL, or TB_INTERIOR / TB_FRAME to split the sweep into the work items that do not / do
// read ghost lines, so the former can run while the exchange is in flight.  Returns CSIM_OK and
// sets *launched = false when the requested part is empty.
static bool is_pos_zero(double v) {
    uint64_t b;
    std::memcpy(&b, &v, sizeof b);
    return b == 0;
}
bool tb_has_zero_variant(int T, int mode) { return (T == 3 || T == 4) && (mode == MODE_UNIT || mode == MODE_RECIP); }
bool tb_has_staged_variant(int T, int mode) { return (T == 3 || T == 4) && (mode == MODE_UNIT || mode == MODE_RECIP); }
// CSIM_TB_KERNEL=reg selects the register-only sweep for every depth (A/B timing); default: the staged
// sweep wherever it exists
static bool tb_staged_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("CSIM_TB_KERNEL");
        return !(e && std::strcmp(e, "reg") == 0);
    }();
    return on;
}

// See csim_internal.hpp.  The dropped-term kernels are bit-identical to the full ones when (tb_update)
//   (1) every cell the sweep reads is finite and no cell is -0.0              → scanned here, once
//       per upload; kept by the step kernels: an update only yields -0 from c == -0, and
//   (2) the field stays finite                                                → the step is a convex
//       combination of the five cells when dt*(2D(1/dx²+1/dy²) + |vx|/dx + |vy|/dy) <= 1 (max
//       principle; the scan bounds |u| by 2^1000, so rounding cannot carry it to overflow), and
// Anything else — a tainted tile, a non-monotone step, a -0.0 velocity — runs the full arithmetic.
// Multi-rank runs decide per rank, without a collective (a host-synchronous reduction per call would
// drain the launch queue: measured 7.1 → 8.0 ms per 100 steps on 2 GPUs).  That is sound for (1)'s
// -0.0 half: the halo cells a rank advances redundantly are only ever NEIGHBOURS of its own cells,
// and the sign of a zero neighbour cannot change a result whose centre is not -0.0.  It leaves one
// gap, stated in DESIGN.md: a non-finite value that enters a clean rank's halo from a tainted
// neighbour meets 0*inf = NaN in the reference but not here, so WHICH cells turn NaN next to a
// blow-up front may differ across a rank boundary.  Finite fields are bit-identical.
int resolve_zero_terms(csim_field* u, const csim_step_params* p, const StepK& k, int mode, int maxT,
                       bool* allowed) {
    *allowed = false;
    static const bool off = tb_env_int("CSIM_ZERO_TERMS", 1) == 0;
    const bool candidate = !off && tb_has_zero_variant(maxT, mode) && (is_pos_zero(k.vx) || is_pos_zero(k.vy));
    if (!candidate) return CSIM_OK;  // decided from the parameters alone: identical on every rank
    bool ok = std::isfinite(p->bc_value) && std::fabs(p->bc_value) < 0x1p1000 && p->dt > 0.0 && p->D >= 0.0;
    if (ok) {
        const double w = p->dt * (2.0 * p->D * (1.0 / (u->dx * u->dx) + 1.0 / (u->dy * u->dy)) +
                                  std::fabs(p->vx) / u->dx + std::fabs(p->vy) / u->dy);
        ok = w <= 1.0;  // NaN compares false
    }
    csim_ctx* c = u->ctx;
    if (ok && u->values == csim_field::kUnknown) {
        auto* d_bad = reinterpret_cast<unsigned long long*>(c->d_scratch);
        CSIM_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
        if (u->nxt() > 0 && u->nyt() > 0) {
            int nblocks = c->sm_count * 8;
            if (nblocks > u->nyt()) nblocks = u->nyt();
            CSIM_LAUNCH(c, k_scan_values, nblocks, 256, 0, u->at(0, 0), u->nxt(), u->nyt(), u->pitch, d_bad);
        }
        CSIM_CUDA(cudaMemcpyAsync(c->h_scratch, d_bad, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  c->stream));
        CSIM_CUDA(cudaStreamSynchronize(c->stream));
        unsigned long long bad = 0;
        std::memcpy(&bad, c->h_scratch, sizeof bad);
        u->values = bad ? csim_field::kTainted : csim_field::kClean;
    }
    *allowed = ok && u->values == csim_field::kClean;
    return CSIM_OK;
}

cudaError_t tb_launch(int vxs, int vys, bool staged, int T, int mode, const TbArgs& a, cudaStream_t stream) {
    switch ((vxs + 1) * 3 + (vys + 1)) {
        case 0: return tb_launch_nn(staged, T, mode, a, stream);
        case 1: return tb_launch_nz(staged, T, mode, a, stream);
        case 2: return tb_launch_np(staged, T, mode, a, stream);
        case 3: return tb_launch_zn(staged, T, mode, a, stream);
        case 4: return tb_launch_zz(staged, T, mode, a, stream);
        case 5: return tb_launch_zp(staged, T, mode, a, stream);
        case 6: return tb_launch_pn(staged, T, mode, a, stream);
        case 7: return tb_launch_pz(staged, T, mode, a, stream);
        default: return tb_launch_pp(staged, T, mode, a, stream);
    }
}

// Geometry of one sweep: which cells are advanced / stored / need no boundary fix-up, and how the
// tile is cut into (strip, chunk) work items.  Pure host arithmetic (csim_sweep_plan exposes it to the
// CPU tests).  phys: bit s set = side s is a physical boundary; slots: resident warps of the machine.
// Returns false when the requested part has no work item.
static bool tb_geometry(int nx, int ny, long long pitch, int T, int phys, int slots, int part, TbArgs& a) {
    a.pitch = pitch;
    a.nx = nx;
    a.ny = ny;
    a.phys = phys;
    const bool pl = a.phys & 1, pr = a.phys & 2, pb = a.phys & 4, pt = a.phys & 8;
    a.xlo = pl ? 0 : -T;  // ghost lines of a neighbour side are advanced too (their validity shrinks
    a.xhi = pr ? nx : nx + T;  // by one line per level, which is exactly what T lines allow)
    a.ylo = pb ? 0 : -T;
    a.yhi = pt ? ny : ny + T;
    a.sx0 = pl ? -1 : 0;
    a.sx1 = pr ? nx + 1 : nx;
    a.sy0 = pb ? -1 : 0;
    a.sy1 = pt ? ny + 1 : ny;
    a.fx0 = pl ? 1 : a.xlo;
    a.fx1 = pr ? nx - 1 : a.xhi;
    a.fy0 = pb ? 1 : a.ylo;
    a.fy1 = pt ? ny - 1 : a.yhi;
    a.xmax_load = static_cast<int>(pitch) - kLeadX;
    a.pf_rows = tb_env_int("CSIM_TB_PF", 4);
    a.pf_off = static_cast<long long>(a.pf_rows > 0 ? a.pf_rows : 0) * pitch;
    if (a.pf_rows <= 0) a.pf_rows = 1 << 29;  // off: r + 3 + pf_rows < r_end never holds
    a.nstrips = (nx + kTbWout - 1) / kTbWout;
    if (a.nstrips < 1) a.nstrips = 1;
    a.edge_split = tb_env_int("CSIM_TB_EDGE_SPLIT", 2);
    if (a.edge_split < 1) a.edge_split = 1;
    // chunk height: fill k whole rounds of the resident warp slots of the machine
    const int rows = a.sy1 - a.sy0;
    const int n_edge = a.nstrips >= 2 ? 2 : 1, n_int = a.nstrips - n_edge;
    const int weight = n_int + n_edge * a.edge_split;
    // Chunk height.  A launch runs as several rounds of resident warps; short chunks keep the last
    // round from idling the machine, tall chunks amortise the 2T rows each chunk re-computes.
    // Measured on B200: at 8192^2 64-128 rows is the flat optimum (profiles/r01_tb_tuning.md); at
    // 16384^2, with four times the items, 192-384 rows gain 3-4 % over 96 (profiles/r02_tuning.md).
    // So: at least four rounds of resident warps, at most 384 rows.
    int ch = tb_env_int("CSIM_TB_CHUNK", 0);
    if (ch <= 0) {
        static const int ch_max = tb_env_int("CSIM_TB_CHUNK_MAX", 384);
        static const int min_rounds = tb_env_int("CSIM_TB_ROUNDS", 4);
        const long long want = static_cast<long long>(rows) * weight / (static_cast<long long>(min_rounds > 0 ? min_rounds : 1) * slots);
        ch = static_cast<int>(want < 32 ? 32 : (want > ch_max ? ch_max : want));
    }
    if (ch > rows) ch = rows;
    if (ch < 1) ch = 1;
    // a last chunk shorter than T rows would put the ghost-line reads of the bottom T rows into the
    // chunk before it, which the multi-GPU loop runs while the halos are still travelling
    for (int tries = 0; tries < 2 * kTbMaxT && ch < rows && rows % ch != 0 && rows % ch < T; ++tries) ++ch;
    a.chunk_h = ch;
    a.nchunks = (rows + ch - 1) / ch;
    const int eh = (ch + a.edge_split - 1) / a.edge_split;
    const int nch_edge = (rows + eh - 1) / eh;
    // Tail of the launch.  With uniform chunks the last, partly filled round of resident warps costs
    // a whole chunk time (3.44 rounds at 8192^2: SMs active 93 % of the launch).  Chunks that fill
    // whole rounds keep the full height; the rows left over are cut into shorter chunks, which are
    // enumerated last.  The geometry does not depend on `part`: interior and frame launches of one
    // block must cover every chunk exactly once.
    a.n_main = a.nchunks;
    a.chunk_h2 = ch;
    static const int tail_on = tb_env_int("CSIM_TB_TAIL", 1);
    if (tail_on && n_int > 0) {
        const long long edge_items = static_cast<long long>(n_edge) * nch_edge;
        const long long items_uniform = edge_items + static_cast<long long>(a.nchunks) * n_int;
        const long long full_rounds = items_uniform / slots;
        long long best = ((items_uniform + slots - 1) / slots) * (ch + 2 * T);
        long long n_main = (full_rounds * slots - edge_items) / n_int;
        if (n_main > a.nchunks) n_main = a.nchunks;
        if (full_rounds >= 1 && n_main >= 1 && n_main * ch < rows) {
            const long long rest = rows - n_main * ch;
            for (int div = 2; div <= 4; ++div) {
                int h2 = ch / div < 16 ? 16 : ch / div;
                for (int tries = 0; tries < 2 * kTbMaxT && rest % h2 != 0 && rest % h2 < T; ++tries) ++h2;
                const long long n_small = (rest + h2 - 1) / h2;
                const long long cost = full_rounds * (ch + 2 * T) + ((n_small * n_int + slots - 1) / slots) * (h2 + 2 * T);
                if (cost < best) {
                    best = cost;
                    a.n_main = static_cast<int>(n_main);
                    a.chunk_h2 = h2;
                    a.nchunks = static_cast<int>(n_main + n_small);
                }
            }
        }
    }
    a.int_chunk0 = 0;
    a.frame_pair = 0;
    a.n_frame_items = a.n_frame_tickets = a.n_frame_edge_items = 0;
    int int_chunks = a.nchunks;
    a.n_edge_items = n_edge * nch_edge;
    // The interior part may run while the halos travel, so none of its items may read a ghost line:
    // an item reads T rows above and below its chunk and T columns beside its strip's finished
    // columns.  Chunks 1 … nchunks-2 of strips 1 … nstrips-2 qualify when the first and the last chunk
    // are at least T rows tall and the last strip is at least T columns wide; otherwise (tiny or
    // awkward tiles) the sweep is not split and everything runs as the frame.
    const int first_h = a.n_main >= 1 ? a.chunk_h : a.chunk_h2;
    const int last_h = a.nchunks > a.n_main
                           ? rows - a.n_main * a.chunk_h - (a.nchunks - a.n_main - 1) * a.chunk_h2
                           : rows - (a.nchunks - 1) * a.chunk_h;
    const int last_w = nx - (a.nstrips - 1) * kTbWout;
    const bool split_ok = !tb_split_pointless(a.nchunks, n_int) && first_h >= T && last_h >= T && last_w >= T;
    if (part == TB_INTERIOR) {  // interior strips, all chunks but the first and the last
        a.n_edge_items = 0;
        a.int_chunk0 = 1;
        int_chunks = split_ok ? a.nchunks - 2 : 0;
    } else if (part == TB_FRAME) {  // both edge strips + first and last chunk of every interior strip
        a.frame_pair = 1;
        int_chunks = 2;
        if (!split_ok) {  // the interior part is empty: the frame is everything
            a.frame_pair = 0;
            int_chunks = a.nchunks;
        }
    }
    else if (part == TB_COUPLED) {  // the frame's items first, then the interior's, in one launch
        a.frame_pair = split_ok ? 1 : 0;
        a.n_frame_edge_items = a.n_edge_items;
        const int frame_int = split_ok ? 2 : a.nchunks;
        a.n_frame_items = a.n_edge_items + n_int * frame_int;
        const int interior_chunks = split_ok ? a.nchunks - 2 : 0;
        a.n_items = a.n_frame_items + n_int * interior_chunks;
        int tickets = 0;
        for (int item = 0; item < a.n_frame_items; ++item) {
            int strip, ya, yb;
            if (tb_item_map(a, item, strip, ya, yb)) ++tickets;
        }
        a.n_frame_tickets = tickets;
        return a.n_items > 0;
    }
    if (int_chunks < 0) int_chunks = 0;
    a.n_items = n_int * int_chunks + a.n_edge_items;
    return a.n_items > 0;
}

int launch_step_tb(const csim_field* u, csim_field* out, const csim_step_params* p, const StepK& k, int mode,
                   int T, int part, cudaStream_t stream, bool* launched, bool zero_terms, const TbCoupling* coupling) {
    csim_ctx* c = u->ctx;
    TbArgs a;
    a.u = u->interior();
    a.out = out->interior();
    a.out_minus_u = reinterpret_cast<const char*>(a.out) - reinterpret_cast<const char*>(a.u);
    int phys = 0;
    for (int s = 0; s < 4; ++s)
        if (p->nbr[s] == CSIM_PROC_NULL) phys |= 1 << s;
    if (launched) *launched = false;
    const int slots = c->sm_count * kTbBlocksPerSM * kTbWarpsPerBlock;
    if (!tb_geometry(u->nx, u->ny, u->pitch, T, phys, slots, part, a)) return CSIM_OK;
    if (part == TB_COUPLED) {
        CSIM_REQUIRE(coupling != nullptr, CSIM_ERR_INVALID, "launch_step_tb: coupled launch without coupling");
        a.seq = coupling->seq;
        a.halo_flag = coupling->halo_flag;
        a.done_flag = coupling->done_flag;
        a.ticket = coupling->ticket;
        a.err = coupling->err;
        a.timeout_ns = coupling->timeout_ns;
    } else {
        a.seq = 0;
        a.halo_flag = a.done_flag = a.ticket = a.err = nullptr;
        a.timeout_ns = 0;
    }
    a.bcL = p->bc[0];
    a.bcR = p->bc[1];
    a.bcB = p->bc[2];
    a.bcT = p->bc[3];
    a.value = p->bc_value;
    a.k = k;
    if (launched) *launched = true;
    // upwind selectors; a component that is exactly +0.0 may drop its term (tb_update) when the
    // caller established the conditions (zero_terms) and the variant is instantiated
    int vxs = k.vx_pos ? 1 : -1, vys = k.vy_pos ? 1 : -1;
    if (zero_terms && tb_has_zero_variant(T, mode)) {
        if (is_pos_zero(k.vx)) vxs = 0;
        if (is_pos_zero(k.vy)) vys = 0;
    }
    const bool staged = tb_staged_enabled() && tb_has_staged_variant(T, mode);
    const cudaError_t e = tb_launch(vxs, vys, staged, T, mode, a, stream);
    ++c->launches;
    if (e != cudaSuccess) return cuda_fail(e, "k_step_tb", __FILE__, __LINE__);
    return CSIM_OK;
}

// physics constants and arithmetic mode of a step on tile `u`
int step_setup(const csim_field* u, const csim_step_params* p, StepK* k, int* mode) {
    for (int s = 0; s < 4; ++s)
        CSIM_REQUIRE(p->bc[s] >= 0 && p->bc[s] <= 2, CSIM_ERR_INVALID, "step: unknown BC type");
    bool use_div = false;
    *k = make_consts(u->dx, u->dy, p->D, p->vx, p->vy, p->dt, p->flags, &use_div);
    const bool unit = k->rdx2 == 1.0 && k->rdy2 == 1.0 && k->rdx == 1.0 && k->rdy == 1.0;
    *mode = use_div ? MODE_DIV : (unit ? MODE_UNIT : MODE_RECIP);
    return CSIM_OK;
}

static int check_pair(const csim_field* u, const csim_field* out, const char* who) {
    CSIM_REQUIRE(u != nullptr && out != nullptr, CSIM_ERR_INVALID, std::string(who) + ": null field");
    CSIM_REQUIRE(u->ctx == out->ctx, CSIM_ERR_INVALID, std::string(who) + ": fields belong to different contexts");
    CSIM_REQUIRE(u->nx == out->nx && u->ny == out->ny && u->h == out->h, CSIM_ERR_INVALID,
                 std::string(who) + ": fields differ in geometry");
    CSIM_REQUIRE(u->base != out->base, CSIM_ERR_INVALID, std::string(who) + ": u and out alias");
    // with h == 0 the reference reads at(i-1,…) out of range and throws std::out_of_range
    CSIM_REQUIRE(u->h >= 1 || u->nx == 0 || u->ny == 0, CSIM_ERR_RANGE, "Field index out of range");
    return CSIM_OK;
}

}  // namespace csim

using namespace csim;

extern "C" {

int csim_diffusion_step(const csim_field* u, csim_field* out, double D, double dt) {
    if (int rc = check_pair(u, out, "csim_diffusion_step")) return rc;
    out->values = csim_field::kUnknown;
    csim_ctx* c = u->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    bool use_div = false;
    const StepK k = make_consts(u->dx, u->dy, D, 0.0, 0.0, dt, 0, &use_div);
    if (u->nx > 0 && u->ny > 0) {
        const dim3 block(64, 4);
        const dim3 grid((u->nx + 63) / 64, (u->ny + 3) / 4);
        if (use_div)
            CSIM_LAUNCH(c, k_diffusion<true>, grid, block, 0, u->interior(), out->interior(), u->nx, u->ny,
                        u->pitch, k);
        else
            CSIM_LAUNCH(c, k_diffusion<false>, grid, block, 0, u->interior(), out->interior(), u->nx, u->ny,
                        u->pitch, k);
    }
    const int n = u->nxt() > u->nyt() ? u->nxt() : u->nyt();
    if (n > 0 && u->nxt() > 0 && u->nyt() > 0)
        CSIM_LAUNCH(c, k_ring_copy, (n + 255) / 256, 256, 0, u->at(0, 0), out->at(0, 0), u->nxt(), u->nyt(),
                    u->pitch);
    return CSIM_OK;
}

int csim_advection_step(const csim_field* u, csim_field* out, double vx, double vy, double dt) {
    if (int rc = check_pair(u, out, "csim_advection_step")) return rc;
    out->values = csim_field::kUnknown;
    csim_ctx* c = u->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    bool use_div = false;
    const StepK k = make_consts(u->dx, u->dy, 0.0, vx, vy, dt, 0, &use_div);
    if (u->nx > 0 && u->ny > 0) {
        const dim3 block(64, 4);
        const dim3 grid((u->nx + 63) / 64, (u->ny + 3) / 4);
        if (use_div)
            CSIM_LAUNCH(c, k_advection<true>, grid, block, 0, u->interior(), out->interior(), u->nx, u->ny,
                        u->pitch, k);
        else
            CSIM_LAUNCH(c, k_advection<false>, grid, block, 0, u->interior(), out->interior(), u->nx, u->ny,
                        u->pitch, k);
    }
    return CSIM_OK;
}

int csim_apply_boundary(csim_field* f, const int nbr[4], const int bc[4], double value) {
    CSIM_REQUIRE(f != nullptr && nbr != nullptr && bc != nullptr, CSIM_ERR_INVALID,
                 "csim_apply_boundary: null argument");
    CSIM_CUDA(cudaSetDevice(f->ctx->device));
    f->values = csim_field::kUnknown;
    return launch_boundary(f, nbr, bc, value);
}

int csim_step_fused(csim_field* u, csim_field* tmp, const csim_step_params* p, int nsteps) {
    if (int rc = check_pair(u, tmp, "csim_step_fused")) return rc;
    CSIM_REQUIRE(p != nullptr && nsteps >= 0, CSIM_ERR_INVALID, "csim_step_fused: bad arguments");
    CSIM_REQUIRE(u->h == 1, CSIM_ERR_UNSUPPORTED, "csim_step_fused: the fused path needs halo == 1 (main.cpp:65)");
    csim_ctx* c = u->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    StepK k;
    int mode = 0;
    if (int rc = step_setup(u, p, &k, &mode)) return rc;
    if (u->nx == 0 || u->ny == 0) {  // nothing to advance; keep the reference's swap parity
        if (nsteps & 1) csim_field_swap(u, tmp);
        return CSIM_OK;
    }
    if (p->flags & CSIM_STEP_NO_TEMPORAL) {
        // plain path: boundary kernels + one-step sweep (independent implementation, kept as a
        // cross-check of the blocked kernel)
        for (int n = 0; n < nsteps; ++n) {
            if (int rc = launch_boundary(u, p->nbr, p->bc, p->bc_value)) return rc;  // main.cpp:102
            if (int rc = launch_step_v1(u, tmp, k, mode == MODE_DIV)) return rc;     // main.cpp:104-107
            u->values = tmp->values = csim_field::kUnknown;
            csim_field_swap(u, tmp);                                                 // main.cpp:109
        }
        return CSIM_OK;
    }
    bool all_phys = true;
    for (int s = 0; s < 4; ++s) all_phys = all_phys && p->nbr[s] == CSIM_PROC_NULL;
    // IEEE division is compute-bound: blocking in time buys nothing there.  Tiles with neighbours
    // carry one ghost line per csim_halo_exchange, so here they advance one step per sweep
    // (csim_run_steps exchanges T lines and blocks them too).
    const int maxT = !all_phys ? 1 : (mode == MODE_DIV ? tb_max_T_div() : tb_max_T());
    // CSIM_DEBUG_SPLIT=1 (measurement aid): run the single-GPU sweep as the interior + frame pair of
    // launches on two streams exactly as the multi-GPU loop does, to price the split by itself.
    static const bool debug_split = tb_env_int("CSIM_DEBUG_SPLIT", 0) != 0;
    bool zero_terms = false;
    if (nsteps >= maxT)
        if (int rc = resolve_zero_terms(u, p, k, mode, maxT, &zero_terms)) return rc;
    const int values_after = zero_terms ? csim_field::kClean
                                        : (u->values == csim_field::kTainted ? csim_field::kTainted : csim_field::kUnknown);
    int left = nsteps;
    while (left > 0) {
        const int T = left < maxT ? left : maxT;
        if (debug_split) {
            CSIM_CUDA(cudaEventRecord(c->ev_fork, c->stream));
            CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));
            if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_INTERIOR, c->stream, nullptr, zero_terms)) return rc;
            if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_FRAME, c->stream_x, nullptr, zero_terms)) return rc;
            CSIM_CUDA(cudaEventRecord(c->ev_join, c->stream_x));
            CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        } else {
            if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_ALL, c->stream, nullptr, zero_terms)) return rc;
        }
        tmp->values = values_after;  // tmp is rewritten from u; a clean u under a monotone step stays clean
        csim_field_swap(u, tmp);
        left -= T;
    }
    return CSIM_OK;
}

// The kernels' division by a constant divisor, run on the host (same source, step_math.cuh) — for tests.
double csim_div_by_const(double a, double d) { return div_by_const(a, d, 1.0 / d); }
int csim_div_by_const_fast(double a, double d) { return div_by_const_took_fast_path(a, d); }

int csim_steps_per_sweep(void) { return tb_max_T(); }
const char* csim_sweep_kernel(void) {
    return tb_staged_enabled() && tb_has_staged_variant(tb_max_T(), MODE_UNIT) ? "k_step_tbs" : "k_step_tb";
}

int csim_sweep_plan(int nx, int ny, int T, const int nbr[4], int resident_warps, int part,
                    csim_sweep_item* items, int capacity, int* count) {
    CSIM_REQUIRE(nbr != nullptr && count != nullptr && (items != nullptr || capacity == 0), CSIM_ERR_INVALID,
                 "csim_sweep_plan: null argument");
    CSIM_REQUIRE(nx >= 1 && ny >= 1 && T >= 1 && T <= kTbMaxT && part >= TB_ALL && part <= TB_COUPLED, CSIM_ERR_INVALID,
                 "csim_sweep_plan: bad size, depth or part");
    int phys = 0;
    for (int s = 0; s < 4; ++s)
        if (nbr[s] == CSIM_PROC_NULL) phys |= 1 << s;
    const int slots = resident_warps > 0 ? resident_warps : 148 * kTbBlocksPerSM * kTbWarpsPerBlock;
    const long long pitch = (static_cast<long long>(kLeadX) + nx + kTailX + 15) / 16 * 16;
    TbArgs a;
    *count = 0;
    if (!tb_geometry(nx, ny, pitch, T, phys, slots, part, a)) return CSIM_OK;
    int n = 0;
    for (int item = 0; item < a.n_items; ++item) {
        int strip, ya, yb;
        if (!tb_item_map(a, item, strip, ya, yb)) continue;
        if (n < capacity) {
            csim_sweep_item& it = items[n];
            it.strip = strip;
            it.y0 = ya;
            it.y1 = yb;
            // finished columns of a strip (step_tb.cuh, tb_store_row): positions 4 … 123 of its 128,
            // plus the ghost column of a physical side in the first / last strip
            it.x0 = strip == 0 ? a.sx0 : strip * kTbWout;
            it.x1 = strip == a.nstrips - 1 ? a.sx1 : (strip + 1) * kTbWout;
        }
        ++n;
    }
    *count = n;
    return CSIM_OK;
}

int csim_minmax(const csim_field* f, double* mn, double* mx) {
    CSIM_REQUIRE(f != nullptr && mn != nullptr && mx != nullptr, CSIM_ERR_INVALID, "csim_minmax: null argument");
    CSIM_REQUIRE(f->nxt() > 0 && f->nyt() > 0, CSIM_ERR_INVALID, "csim_minmax: empty field");
    csim_ctx* c = f->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    int nblocks = c->sm_count * 4;
    if (nblocks > f->nyt()) nblocks = f->nyt();
    if (static_cast<size_t>(2 * nblocks + 2) > c->scratch_doubles) nblocks = static_cast<int>(c->scratch_doubles / 2 - 1);
    CSIM_LAUNCH(c, k_minmax_partial, nblocks, 256, 0, f->at(0, 0), f->nxt(), f->nyt(), f->pitch, c->d_scratch + 2);
    CSIM_LAUNCH(c, k_minmax_final, 1, 256, 0, c->d_scratch + 2, nblocks, c->d_scratch);
    CSIM_CUDA(cudaMemcpyAsync(c->h_scratch, c->d_scratch, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    *mn = c->h_scratch[0];
    *mx = c->h_scratch[1];
    return CSIM_OK;
}

int csim_field_value_state(const csim_field* f) { return f ? f->values : -1; }

int csim_field_health(const csim_field* f, double* max_abs, uint64_t* nonfinite) {
    CSIM_REQUIRE(f != nullptr && max_abs != nullptr && nonfinite != nullptr, CSIM_ERR_INVALID,
                 "csim_field_health: null argument");
    csim_ctx* c = f->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    CSIM_CUDA(cudaMemsetAsync(c->d_scratch, 0, 2 * sizeof(double), c->stream));
    if (f->nx > 0 && f->ny > 0) {
        int nblocks = c->sm_count * 4;
        if (nblocks > f->ny) nblocks = f->ny;
        CSIM_LAUNCH(c, k_health, nblocks, 256, 0, f->interior(), f->nx, f->ny, f->pitch, c->d_scratch,
                    reinterpret_cast<unsigned long long*>(c->d_scratch + 1));
    }
    CSIM_CUDA(cudaMemcpyAsync(c->h_scratch, c->d_scratch, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    *max_abs = c->h_scratch[0];
    std::memcpy(nonfinite, &c->h_scratch[1], sizeof(uint64_t));
    return CSIM_OK;
}

}  // extern "C"
