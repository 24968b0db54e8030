/*
 * csim.h — C ABI of the B200-native timestep hot path (libcsim_b200.so).
 *
 * This is the drop-in boundary for the hot path of climate-sim-mpi-cpp:
 *     exchange_halos → apply_boundary → diffusion_step → advection_step → swap
 *     (reference src/main.cpp:93-118).
 * The reference has no FFI; its boundary is the set of free functions in include/*.hpp.  Each entry
 * point below names the reference interface it stands behind (file:line, relative to the reference
 * repository).  The C++ wrappers with the reference's own names and signatures live in
 * climate-sim-mpi-cpp_b200/host/ (see INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns CSIM_OK (0) or a csim_status code; no exception crosses the ABI;
 *     csim_last_error() gives the message of the calling thread's last failure;
 *   - compute calls are ASYNCHRONOUS on the context's CUDA stream; csim_sync() or any call that
 *     moves data to host memory orders them;
 *   - one context per GPU (per rank); a context is not thread-safe;
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with
 *     CSIM_ERR_CUDA.
 *
 * Host-side field layout = the reference's (include/field.hpp:5-21, src/field.cpp:20-25):
 * row-major (ny+2h) x (nx+2h) doubles, idx(i,j) = j*(nx+2h)+i, interior at [h, h+n).
 * Device-side layout: see csim_field_info and DESIGN.md ("data layout in HBM").
 */
#ifndef CSIM_H
#define CSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSIM_ABI_VERSION 1

typedef enum csim_status {
    CSIM_OK = 0,
    CSIM_ERR_INVALID = 1,     /* bad argument (null pointer, size mismatch, unknown enum)      */
    CSIM_ERR_CUDA = 2,        /* CUDA runtime/driver failure, or no device                     */
    CSIM_ERR_NOMEM = 3,       /* host or device allocation failed                              */
    CSIM_ERR_RANGE = 4,       /* index outside the padded tile (Field::at → std::out_of_range) */
    CSIM_ERR_UNSUPPORTED = 5, /* valid in the reference, not on this path (e.g. halo != 1)     */
    CSIM_ERR_COMM = 6,        /* NCCL failure                                                  */
    CSIM_ERR_TIMEOUT = 7      /* a neighbour's halo did not arrive within the bounded wait     */
} csim_status;

/* enum class BCType { Dirichlet, Neumann, Periodic } — reference include/boundary.hpp:5 */
typedef enum csim_bc { CSIM_BC_DIRICHLET = 0, CSIM_BC_NEUMANN = 1, CSIM_BC_PERIODIC = 2 } csim_bc;

/* side order used by every 4-array: BCConfig{left,right,bottom,top} (include/boundary.hpp:7-12)
 * and Decomp2D{nbr_lr[0],nbr_lr[1],nbr_du[0],nbr_du[1]} (include/decomp.hpp:8-9) */
enum { CSIM_LEFT = 0, CSIM_RIGHT = 1, CSIM_BOTTOM = 2, CSIM_TOP = 3 };

#define CSIM_PROC_NULL (-1) /* MPI_PROC_NULL: no neighbour, the side is a physical boundary */

typedef struct csim_ctx csim_ctx;     /* one GPU, one stream, scratch buffers             */
typedef struct csim_field csim_field; /* one halo-padded, pitch-aligned device tile       */

/* ---- context -------------------------------------------------------------------------------- */

/* Bind a context to CUDA device `device` and create its stream. */
int csim_ctx_create(int device, csim_ctx** out);
int csim_ctx_destroy(csim_ctx* ctx);
/* Number of CUDA devices visible to the process (0 with CSIM_ERR_CUDA when there is none). */
int csim_device_count(int* count);
/* MPI_Barrier / end-of-loop ordering of src/main.cpp:82,120: wait for all queued work. */
int csim_sync(csim_ctx* ctx);
/* cudaStream_t of the context as an opaque pointer (for CUDA-event timing by the caller). */
void* csim_ctx_stream(csim_ctx* ctx);
int csim_ctx_device(const csim_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t csim_ctx_launch_count(const csim_ctx* ctx);
const char* csim_last_error(void);
int csim_abi_version(void);

/* ---- Field: reference include/field.hpp:5-21, src/field.cpp:6-31 -------------------------- */

typedef struct csim_field_info {
    int nx, ny, halo;  /* Field::nx_local, ny_local, halo                                       */
    double dx, dy;     /* Field::dx, dy                                                          */
    int64_t pitch;     /* doubles between consecutive device rows (multiple of 16 = 128 B)       */
    int lead_x;        /* doubles before interior x=0 in a device row (16: interior 128-B aligned) */
    int lead_y;        /* device rows before interior y=0                                        */
    int64_t rows;      /* device rows allocated                                                  */
    void* base;        /* device pointer of the allocation                                       */
    void* interior;    /* device pointer of interior cell (0,0) = Field::at(h,h)                 */
} csim_field_info;

/* Field::Field(nx,ny,h,dx,dy) — src/field.cpp:6-12; zero-initialised like the std::vector.
 * Sizes are 64-bit internally (the reference overflows int above ~46339^2, field.cpp:12). */
int csim_field_create(csim_ctx* ctx, int nx, int ny, int halo, double dx, double dy,
                      csim_field** out);
int csim_field_destroy(csim_field* f);
int csim_field_get_info(const csim_field* f, csim_field_info* out);
/* Field::fill — src/field.cpp:31 (whole padded tile) */
int csim_field_fill(csim_field* f, double value);
/* host padded tile [(ny+2h)*(nx+2h)] → device / device → host; synchronous w.r.t. the host buffer */
int csim_field_upload(csim_field* f, const double* host_padded);
int csim_field_download(const csim_field* f, double* host_padded);
/* The de-haloed tile, ny*nx doubles — the buffer src/io.cpp:411-416 assembles before its put.
 * `host_dense` should be pinned for full PCIe speed; csim_host_alloc provides that. */
int csim_field_download_interior(const csim_field* f, double* host_dense);
/* A w x h window of the tile, first cell (x0, y0) in INTERIOR coordinates (ghost cells are -1 and nx / ny),
 * into a dense host array of h rows of w doubles; synchronous.  For spot checks of tiles too large to
 * bring back whole.  CSIM_ERR_RANGE if the window leaves the padded tile. */
int csim_field_download_window(const csim_field* f, int x0, int y0, int w, int h, double* host_dense);
/* Same, asynchronous on the context stream (host buffer must be pinned; order with csim_sync). */
int csim_field_upload_async(csim_field* f, const double* host_padded_pinned);
int csim_field_download_interior_async(const csim_field* f, double* host_dense_pinned);
/* Snapshot path (src/io.cpp:411-418): the de-haloed tile packed dense AND byte-swapped to big-endian
 * (the NetCDF wire order) by a kernel, then copied to pinned host memory, asynchronously on the
 * context stream.  The host bytes can be written to a CDF-5 file as they are. */
int csim_field_download_interior_be_async(const csim_field* f, void* host_dense_pinned);
/* Snapshot hand-off that does not hold up the time loop (src/main.cpp:96-99 → src/io.cpp:411-418): the
 * de-haloed tile is packed into one of two dense device staging buffers on the context stream (≈ 1 ms at
 * 16384^2), optionally byte-swapped to big-endian, and a separate copy stream moves it to pinned host
 * memory while the time steps queued after this call run.  *event completes when the host buffer is
 * filled: pass it to csim_event_wait (which also releases it).  A third snapshot waits (on the device)
 * for the first one's copy. */
int csim_field_snapshot_async(const csim_field* f, void* host_dense_pinned, int big_endian, void** event);
/* Record an event after everything queued on the context stream so far; csim_event_wait blocks the
 * calling host thread until that point is reached (and releases the event).  Lets a writer thread
 * wait for one snapshot copy without waiting for the time steps queued behind it. */
int csim_event_record(csim_ctx* ctx, void** event);
int csim_event_wait(csim_ctx* ctx, void* event);

/* Field::at(i,j) read / write of one cell in padded coordinates — src/field.cpp:14-29.
 * Out-of-range indices return CSIM_ERR_RANGE (the reference throws std::out_of_range). */
int csim_field_get(const csim_field* f, int i, int j, double* value);
int csim_field_set(csim_field* f, int i, int j, double value);
/* std::swap(u.data, tmp.data) — src/main.cpp:109: exchange the device buffers of two tiles of
 * identical geometry. */
int csim_field_swap(csim_field* a, csim_field* b);
/* std::copy(u.data → tmp.data) — src/main.cpp:104 (whole padded tile, device to device). */
int csim_field_copy(const csim_field* src, csim_field* dst);
/* Best effort: restrict the calling thread to the CPUs of the NUMA node the GPU `device` hangs off
 * (sysfs numa_node of its PCI function), so that pinned buffers it allocates next — first touch — and its
 * copies stay node-local when all GPUs of a box move tiles at once.  *node receives the node or -1 when
 * the platform reports none (then nothing is changed).  Never fails on a missing sysfs entry. */
int csim_bind_thread_to_device_numa(int device, int* node);
/* pinned host memory for upload/download buffers */
int csim_host_alloc(size_t bytes, void** out);
int csim_host_free(void* p);

/* ---- the kernels of the path, one reference function each ---------------------------------- */

/* diffusion_step(const Field& u, Field& out, double D, double dt) — include/diffusion.hpp:4,
 * src/diffusion.cpp:3-26.  Interior: out = c + (dt*D)*lap5(u) in the reference's operation order,
 * no FMA; then out's outermost ring := u's. */
int csim_diffusion_step(const csim_field* u, csim_field* out, double D, double dt);

/* advection_step(const Field& u, Field& out, double vx, double vy, double dt) —
 * include/advection.hpp:4, src/advection.cpp:5-34.  ACCUMULATES: out += (-dt)*(vx*dudx+vy*dudy),
 * upwind side chosen by vx>=0 / vy>=0. */
int csim_advection_step(const csim_field* u, csim_field* out, double vx, double vy, double dt);

/* apply_boundary(Field&, const Decomp2D&, const BCConfig&, double value) — include/boundary.hpp:14,
 * src/boundary.cpp:12-54.  nbr[s]==CSIM_PROC_NULL marks side s physical; Dirichlet writes `value`,
 * Neumann mirrors the adjacent interior line, Periodic writes nothing.  Order left, right,
 * bottom, top, corners included, exactly as the reference. */
int csim_apply_boundary(csim_field* f, const int nbr[4], const int bc[4], double value);

/* Physics + boundary description of one rank's tile for the fused path. */
typedef struct csim_step_params {
    double D, vx, vy, dt; /* SimConfig::D, vx, vy, dt — include/io.hpp:21-39 */
    int bc[4];            /* csim_bc per side (left,right,bottom,top)         */
    int nbr[4];           /* neighbour rank per side or CSIM_PROC_NULL        */
    double bc_value;      /* Dirichlet value; src/main.cpp:102 passes 0.0     */
    int flags;            /* CSIM_STEP_* below, 0 = defaults                  */
} csim_step_params;

/* Division mode.  The reference divides by dx*dx, dy*dy, dx, dy (diffusion.cpp:12-13,
 * advection.cpp:17-26).  When those divisors are powers of two the kernels multiply by the exact
 * reciprocal, which is bit-identical; otherwise they use IEEE division.  CSIM_STEP_FAST_RECIP
 * forces reciprocal multiplication for any spacing (NOT bit-exact then: L-inf rel <= 1e-12). */
#define CSIM_STEP_FAST_RECIP 0x1
/* Force the plain one-step-per-sweep kernel even where temporal blocking is possible. */
#define CSIM_STEP_NO_TEMPORAL 0x2

/* The fused time step — the body of the loop at src/main.cpp:101-109 without the exchange:
 *   apply_boundary(u) ; tmp := u ; diffusion_step(u,tmp) ; advection_step(u,tmp) ; swap(u,tmp)
 * repeated `nsteps` times in ONE pass per step over HBM (no separate copy, no second sweep).
 * On return (stream order) `u` holds the newest state, as after the reference's swap; `tmp` is
 * scratch (the reference overwrites it at the top of every step, main.cpp:104).  The library may
 * advance several steps per sweep over HBM (temporal blocking) where that leaves the result
 * bit-identical.  Interior cells are bit-identical to the reference's; edge ghost cells hold
 * what the reference's would hold; corner ghosts are unspecified.
 * Sides with a neighbour (nbr != PROC_NULL) read their ghost line as the exchange left it, so
 * for multi-rank runs call csim_halo_exchange before each step (or use csim_run_steps). */
int csim_step_fused(csim_field* u, csim_field* tmp, const csim_step_params* p, int nsteps);
/* Largest number of time steps one sweep advances (the temporal blocking depth T; env CSIM_TB_MAXT). */
int csim_steps_per_sweep(void);
/* Name of the sweep kernel that depth runs with unit spacing: "k_step_tbs" (level-0 rows staged through
 * shared memory by TMA; the default) or "k_step_tb" (register-only; CSIM_TB_KERNEL=reg).  For reports. */
const char* csim_sweep_kernel(void);

/* min / max over the whole padded tile, ghosts included — the "IC min/max" reduction of
 * src/main.cpp:73-77 (std::min_element / std::max_element over Field::data). */
int csim_minmax(const csim_field* f, double* mn, double* mx);

/* safe_dt — include/stability.hpp:5-16 (host scalar; no field is involved in the reference). */
double csim_safe_dt(double dx, double dy, double vx, double vy, double D);
/* Device-side stability diagnostic (not in the reference; north_star's "CFL check as a
 * warp-shuffle reduction"): max |u| over the interior and count of non-finite cells. */
int csim_field_health(const csim_field* f, double* max_abs, uint64_t* nonfinite);
/* What the library currently knows about the values of the tile: 0 unknown (written from outside the
 * fused step since the last scan), 1 clean (every cell of the padded tile finite, below 2^1000 in
 * magnitude and not -0.0), 2 tainted (a scan found such a cell).  When a velocity component is exactly
 * +0.0, csim_step_fused / csim_run_steps scan an unknown tile once and, on clean tiles with a monotone
 * time step (dt*(2D(1/dx^2+1/dy^2)+|vx|/dx+|vy|/dy) <= 1), skip that component's advection term —
 * bit-identical there (DESIGN.md "dropped zero-velocity terms"), 3 of 14 FP64 operations per cell
 * fewer.  Any other tile runs the full arithmetic.  CSIM_ZERO_TERMS=0 in the environment disables it. */
int csim_field_value_state(const csim_field* f);

/* ---- decomposition: include/decomp.hpp:4-17, src/decomp.cpp:5-39 --------------------------- */

typedef struct csim_decomp {
    int dims[2];   /* MPI_Dims_create(size,2): most square, non-increasing; dims[0] splits x */
    int coords[2]; /* row-major: (rank / dims[1], rank % dims[1])                            */
    int nbr[4];    /* left, right, down, up (nbr_lr[0], nbr_lr[1], nbr_du[0], nbr_du[1])     */
    int nx_global, ny_global;
    int nx_local, ny_local; /* last rank of a dimension absorbs the remainder                */
    int x_offset, y_offset; /* coords * base size                                            */
} csim_decomp;

/* Decomp2D::init(comm, nx_global, ny_global) for `rank` of `size` ranks; rank r runs on GPU r. */
int csim_decomp_init(int size, int rank, int nx_global, int ny_global, csim_decomp* out);

/* ---- halo exchange: include/halo.hpp:7, src/halo.cpp:6-50 ---------------------------------- */

#define CSIM_UNIQUE_ID_BYTES 128
/* NCCL bootstrap.  Rank 0 calls csim_comm_unique_id and ships the bytes to the other ranks by any
 * means (bench.py uses torch.distributed); every rank then calls csim_comm_init. */
int csim_comm_unique_id(char id[CSIM_UNIQUE_ID_BYTES]);
int csim_comm_init(csim_ctx* ctx, int size, int rank, const char id[CSIM_UNIQUE_ID_BYTES]);
int csim_comm_destroy(csim_ctx* ctx);
/* MPI_Reduce(..., MPI_DOUBLE, MPI_MAX, 0, comm) of src/main.cpp:127-128 (every rank gets the
 * result) and, with n == 0, MPI_Barrier (main.cpp:82): element-wise max of `n` host doubles over
 * all ranks of the context's communicator; synchronous.  Without a communicator it is the identity. */
int csim_comm_allreduce_max(csim_ctx* ctx, double* inout, int n);

/* exchange_halos(Field&, const Decomp2D&, MPI_Comm): fill the ghost lines of `f` from the up to
 * four neighbours in `dec->nbr`: columns over j in [h, h+ny), rows over all nx+2h cells
 * (src/halo.cpp:28-43).  Edge lines are packed by a kernel, moved with grouped ncclSend/ncclRecv
 * over NVLink, and unpacked by a kernel; all on the context stream. */
int csim_halo_exchange(csim_field* f, const csim_decomp* dec);

/* Which halo path csim_run_steps used last on this context: "none" (no neighbours yet), "nccl" (the default:
 * T-line bands packed by a kernel, one grouped ncclSend/ncclRecv per block, unpack kernel) or "peer"
 * (CSIM_HALO=peer: the bands are stored straight into the neighbours' ghost lines over NVLink by single-warp
 * CTAs that co-reside with the interior sweep, one flag per neighbour; tiles mapped with CUDA IPC between
 * processes, peer access inside one; falls back to "nccl" when a peer cannot be mapped).  The peer path is
 * set up on first use — collectively: every rank's first csim_run_steps with a given pair of tiles must be
 * the same call.  Measurements of both: profiles/r02_multigpu.md. */
const char* csim_halo_path(const csim_ctx* ctx);

/* One region of the wide (T-line, 8-neighbour) exchange csim_run_steps performs per T-step block:
 * interior coordinates of its first cell, extent, and the rank on the other side (-1: no such
 * neighbour).  Index k enumerates directions (dx,dy) row by row from (-1,-1) to (1,1) without (0,0). */
typedef struct csim_xregion {
    int x0, y0, w, h;
    int peer;
} csim_xregion;
/* Host-only: the regions a rank with decomposition `dec` packs (send[k], out of its own tile) and
 * fills (recv[k], in its ghost area) when exchanging T lines.  Bands span the ghost line of a
 * perpendicular physical side (frozen "periodic" ghosts travel with them); corners are T x T. */
int csim_wide_exchange_plan(const csim_decomp* dec, int T, csim_xregion send[8], csim_xregion recv[8]);

/* Host-only: the work items one fused sweep of T steps is cut into on a tile of nx x ny cells whose
 * sides with nbr[s] == CSIM_PROC_NULL are physical, for a machine with `resident_warps` warp slots
 * (0: 148 SMs x 12).  part: 0 the whole sweep, 3 the same items with the frame's first (one coupled launch), 1 the items that read no ghost line (overlapped with the
 * halo exchange), 2 the others (the frame).  Each item is one warp's job: the rows [y0, y1) of the
 * finished columns [x0, x1) of strip `strip` (interior coordinates; ghost lines of physical sides are
 * -1 and nx / ny).  Writes at most `capacity` items and returns the total number in *count.  Lets the
 * geometry be tested without a GPU: the items of part 0 tile the stored region exactly once, and parts
 * 1 and 2 partition them. */
typedef struct csim_sweep_item {
    int strip, x0, x1, y0, y1;
} csim_sweep_item;
int csim_sweep_plan(int nx, int ny, int T, const int nbr[4], int resident_warps, int part,
                    csim_sweep_item* items, int capacity, int* count);

/* ---- the loop ------------------------------------------------------------------------------ */

/* `nsteps` iterations of src/main.cpp:101-109 on this rank: exchange (if the context has a
 * communicator and the tile has neighbours), then the fused step.  Equivalent to calling
 * csim_halo_exchange + csim_step_fused(…,1) nsteps times; the library is free to overlap the
 * exchange with the interior update and to block several steps per sweep where that leaves the
 * fields bit-identical.  Multi-rank runs advance in blocks of T steps (csim_steps_per_sweep): T ghost lines
 * from all eight neighbours before a block, the sweep split into the work items that read ghost lines (frame)
 * and the others (interior), the exchange of the next block travelling while the interior items run.
 * Environment: CSIM_LOOP=coupled runs frame and interior items as one launch coupled to the exchange stream by
 * device flags; CSIM_GRAPH=1 replays the (split) block loop as a CUDA graph; CSIM_HALO=peer, see csim_halo_path. */
int csim_run_steps(csim_field* u, csim_field* tmp, const csim_step_params* p,
                   const csim_decomp* dec, int nsteps);

/* Timeline of the multi-rank block loop (diagnostic; bench.py's halo report).  After
 * csim_halo_profile(ctx, 1) the NEXT csim_run_steps call on a tile with neighbours runs eagerly (no graph
 * replay), with CUDA-event timestamps around every exchange, frame sweep and interior sweep, waits for
 * completion and leaves the figures here; profiling switches itself off again. */
typedef struct csim_halo_stats {
    int blocks;                /* T-step blocks of the profiled call                                    */
    size_t bytes_per_exchange; /* bytes this rank SENDS per exchange (8 regions of T lines; as many come in) */
    double first_exchange_us;  /* exchange(0): pack + NCCL group + unpack, nothing to hide behind       */
    double exchange_us;        /* mean duration of the later exchanges (each runs beside an interior sweep) */
    double overlap_fraction;   /* share of that time that lies inside the concurrent interior sweep     */
    double frame_us;           /* mean frame sweep (edge strips + first/last chunks: the ghost-line readers) */
    double interior_us;        /* mean interior sweep                                                   */
    double total_ms;           /* first to last timestamp of the call                                   */
    double push_us;            /* peer path: this rank's own store kernel, start of the exchange to its end (0: NCCL path) */
    double wait_for_interior_us; /* mean time between the end of an exchange and the start of its frame sweep:
                                  the exchange stream waiting for the previous block's interior sweep  */
} csim_halo_stats;
int csim_halo_profile(csim_ctx* ctx, int enable);
int csim_halo_stats_get(csim_ctx* ctx, csim_halo_stats* out);

/* gaussian_hotspot / constant_zero initial condition on the host tile — src/init.cpp:12-47.
 * Computed with the host libm's exp() so it is bit-identical to the reference's; writes the
 * interior of `host_padded` (ghosts untouched).  preset: 0 gaussian_hotspot, 1 constant_zero. */
int csim_initial_condition_host(double* host_padded, const csim_decomp* dec, int halo,
                                int nx_global, int ny_global, double dx, double dy, int preset,
                                double A, double sigma_frac, double xc_frac, double yc_frac);

/* The same initial condition generated ON THE DEVICE, into the interior of `f` (ghost cells untouched),
 * asynchronously on the context stream — src/init.cpp:12-47 without the host loop and the upload.
 * Bit-identical to csim_initial_condition_host: same operation order, and exp() is the host libm's
 * table-driven algorithm restated for the device (csrc/exp_libm.cuh).  Which rounding sequence the
 * host's exp() uses depends on its build (FMA-contracted or not); csim_exp_variant() probes it once:
 * 1 = FMA variant, 0 = plain variant, -1 = neither matches, in which case this call returns
 * CSIM_ERR_UNSUPPORTED and the caller keeps the host path.  `dec` supplies the offsets, `f` dx and dy. */
int csim_initial_condition_device(csim_field* f, const csim_decomp* dec, int nx_global, int ny_global,
                                  int preset, double A, double sigma_frac, double xc_frac, double yc_frac);
int csim_exp_variant(void);
/* The kernels' division by a divisor known in advance (csrc/step_math.cuh: reciprocal, one correction step,
 * an exact check, IEEE division as the fallback) executed on the host, and whether (a, d) passes the check —
 * for tests against true division.  d must be positive, finite and normal. */
double csim_div_by_const(double a, double d);
int csim_div_by_const_fast(double a, double d);
/* The restated exp() itself on the host (variant 1 or 0), for tests against the host libm. */
double csim_exp_restated(double x, int variant);

#ifdef __cplusplus
}
#endif
#endif /* CSIM_H */
