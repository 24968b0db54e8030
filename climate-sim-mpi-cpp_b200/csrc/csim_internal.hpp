// csim_internal.hpp — private declarations shared by the translation units of libcsim_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "csim.h"

// Device tile geometry (DESIGN.md "data layout in HBM").  A row holds kLeadX doubles of padding,
// then the interior; the interior origin of every row is 128-byte aligned, the pitch is a multiple
// of 128 bytes, and kLeadY rows precede interior row 0.  Ghost cells of a Field with halo h live at
// x in [-h,0) and [nx,nx+h), i.e. inside the padding; the extra padding is what the wide-halo
// (temporally blocked, multi-GPU) path uses.
constexpr int kLeadX = 16;
constexpr int kTailX = 16;
constexpr int kLeadY = 8;
constexpr int kMaxHalo = 8;

struct csim_field;

struct csim_ctx {
    int device = 0;
    std::vector<csim_field*> fields;  // live tiles of this context (orphaned by csim_ctx_destroy)
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    // scratch for reductions: device partials + pinned host landing zone
    double* d_scratch = nullptr;
    double* h_scratch = nullptr;
    size_t scratch_doubles = 0;
    // NCCL state (halo.cu)
    void* comm = nullptr;
    int comm_size = 1, comm_rank = 0;
    double* d_pack = nullptr;  // 4 column buffers: send L, send R, recv L, recv R
    size_t pack_doubles = 0;
    double* d_snap = nullptr;  // dense big-endian snapshot staging (csim_field_download_interior_be_async)
    size_t snap_doubles = 0;
    double* d_stage = nullptr;  // dense staging of large asynchronous host transfers (context.cu)
    size_t stage_doubles = 0;
    double* d_wide = nullptr;  // wide-halo exchange staging: 8 send + 8 recv regions
    size_t wide_doubles = 0;
    // peer-memory halo (halo.cu): the up to eight neighbours' tiles and flag words mapped into this process
    struct PeerLink {
        int rank = -1;
        double* tile[2] = {nullptr, nullptr};  // the neighbour's two tile allocations (its u, tmp at setup)
        long long pitch = 0;
        int nx = 0, ny = 0;
        unsigned* flags = nullptr;             // the neighbour's flag words
        bool ipc = false;                      // mapped with cudaIpcOpenMemHandle (else same-process pointers)
    };
    PeerLink peer[8];
    bool peer_ready = false;
    bool peer_failed = false;                   // mapping was refused once: stay on the NCCL path
    double* peer_tile[2] = {nullptr, nullptr};  // this rank's two allocations, in the order given at setup
    unsigned* d_flags = nullptr;                // [0..7] written by the neighbours, [8] ticket, [9] pushes, [10] waits
    unsigned* h_err = nullptr;                  // pinned, mapped: set by the wait kernel before it traps
    unsigned* d_err = nullptr;
    // snapshot hand-off (context.cu, csim_field_snapshot_async): two dense staging buffers, a copy stream
    cudaStream_t stream_copy = nullptr;
    double* d_snapbuf[2] = {nullptr, nullptr};
    size_t snapbuf_doubles[2] = {0, 0};
    cudaEvent_t ev_snap_packed[2] = {nullptr, nullptr}, ev_snap_done[2] = {nullptr, nullptr};
    int snap_next = 0;
    bool exp_table_loaded = false;    // initcond.cu: 2^(k/128) table copied to this device
    unsigned* d_couple = nullptr;     // halo.cu: [0] halo landed, [1] frame done, [2] ticket (coupled block loop)
    unsigned* h_couple_err = nullptr; // pinned, mapped: set before a coupled wait traps
    unsigned* d_couple_err = nullptr;
    unsigned couple_seq = 0;
    void* run_state = nullptr;        // halo.cu: captured block loops (CUDA graphs) and the halo timeline
    cudaStream_t stream_x = nullptr;  // exchange + frame sweep, overlapped with the interior sweep
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_go = nullptr;
    int sm_count = 148;
};

struct csim_field {
    csim_ctx* ctx = nullptr;
    int nx = 0, ny = 0, h = 0;
    double dx = 1.0, dy = 1.0;
    int64_t pitch = 0, rows = 0;
    double* base = nullptr;
    // What the library knows about the VALUES of the padded tile (kernels.cu, resolve_zero_terms):
    // kUnknown after any write from outside the step kernels, kClean once a scan found every cell
    // finite, below 2^1000 in magnitude and not -0.0, kTainted once a scan found otherwise.  The fused
    // step preserves kClean under a monotone time step, so a run scans once per upload.
    enum { kUnknown = 0, kClean = 1, kTainted = 2 };
    int values = kUnknown;
    double* interior() const { return base + static_cast<int64_t>(kLeadY) * pitch + kLeadX; }
    int nxt() const { return nx + 2 * h; }
    int nyt() const { return ny + 2 * h; }
    // pointer to padded-coordinate cell (i,j) = Field::at(i,j)
    double* at(int i, int j) const { return interior() + static_cast<int64_t>(j - h) * pitch + (i - h); }
};

namespace csim {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// 1/x is exact and x*(1/x)==1 iff x is a (normal) power of two
bool is_pow2(double x);

void run_state_destroy(csim_ctx* c);  // halo.cu
int peer_teardown(csim_ctx* c);        // halo.cu
struct StepK;
enum { TB_ALL = 0, TB_INTERIOR = 1, TB_FRAME = 2, TB_COUPLED = 3 };
int tb_max_T();
int tb_max_T_div();
bool tb_split_pointless(int nchunks, int n_int);
// kernels.cu
int step_setup(const csim_field* u, const csim_step_params* p, StepK* k, int* mode);
// May the sweep drop the advection term of a velocity component that is exactly +0.0 (tb_update)?
// Scans `u` if its state is unknown (one pass + host sync, once per upload).
int resolve_zero_terms(csim_field* u, const csim_step_params* p, const StepK& k, int mode, int maxT,
                       bool* allowed);
// Coupling of a TB_COUPLED launch with the exchange stream (step_tb.cuh: tb_frame_enter / tb_frame_leave)
struct TbCoupling {
    unsigned seq = 0;
    unsigned* halo_flag = nullptr;
    unsigned* done_flag = nullptr;
    unsigned* ticket = nullptr;
    unsigned* err = nullptr;
    unsigned long long timeout_ns = 0;
};
int launch_step_tb(const csim_field* u, csim_field* out, const csim_step_params* p, const StepK& k, int mode,
                   int T, int part, cudaStream_t stream, bool* launched, bool zero_terms = false,
                   const TbCoupling* coupling = nullptr);

}  // namespace csim

#define CSIM_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return csim::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define CSIM_REQUIRE(cond, code, msg)                 \
    do {                                              \
        if (!(cond)) return csim::fail((code), (msg)); \
    } while (0)

// launch bookkeeping: every kernel launch of the library goes through this
#define CSIM_LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
    do {                                                                         \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);         \
        ++(ctx)->launches;                                                       \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) return csim::cuda_fail(e__, #kernel, __FILE__, __LINE__); \
    } while (0)
