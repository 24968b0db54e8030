// halo.cu — halo exchange between ranks (one rank per GPU) over NVLink.
//
// Stands behind exchange_halos(Field&, const Decomp2D&, MPI_Comm) — include/halo.hpp:7,
// src/halo.cpp:6-50.  The reference posts Irecv/Isend pairs with a strided column datatype and a
// contiguous row datatype.  Here: the two strided edge columns are packed by a kernel, rows are
// sent straight out of (and received straight into) the pitched tile, all transfers go out as one
// grouped ncclSend/ncclRecv batch on the context stream, and a kernel unpacks the two ghost
// columns.  Row payloads are nx+2h cells wide, ghost-column cells included (halo.cpp:16-18), and
// carry those cells' pre-exchange values; corner ghosts are unspecified as in the reference.
//
// NCCL is resolved with dlopen at first use, not at link time: a process that also hosts PyTorch
// must share PyTorch's bundled libnccl.so.2 (loading the system copy first breaks `import torch`),
// while the stand-alone C++ driver picks up the system library.  CSIM_NCCL_LIB overrides the name.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "csim_internal.hpp"
#include "step_tb.cuh"

namespace csim {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.ok) return CSIM_OK;
    const char* env = std::getenv("CSIM_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(CSIM_ERR_COMM, std::string("cannot load NCCL: ") + dlerror());
#define CSIM_SYM(field, name)                                                      \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));       \
    if (!g_nccl.field) return fail(CSIM_ERR_COMM, std::string("NCCL symbol missing: ") + name)
    CSIM_SYM(GetUniqueId, "ncclGetUniqueId");
    CSIM_SYM(CommInitRank, "ncclCommInitRank");
    CSIM_SYM(CommDestroy, "ncclCommDestroy");
    CSIM_SYM(Send, "ncclSend");
    CSIM_SYM(Recv, "ncclRecv");
    CSIM_SYM(AllReduce, "ncclAllReduce");
    CSIM_SYM(AllGather, "ncclAllGather");
    CSIM_SYM(GroupStart, "ncclGroupStart");
    CSIM_SYM(GroupEnd, "ncclGroupEnd");
    CSIM_SYM(GetErrorString, "ncclGetErrorString");
#undef CSIM_SYM
    g_nccl.ok = true;
    return CSIM_OK;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    return fail(CSIM_ERR_COMM, std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
}
#define CSIM_NCCL(call)                                         \
    do {                                                        \
        ncclResult_t r__ = (call);                              \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);   \
    } while (0)

// send buffers: [0,ny) left interior column x=0, [ny,2ny) right interior column x=nx-1
__global__ void k_pack_columns(const double* __restrict__ u, int nx, int ny, int64_t pitch,
                               double* __restrict__ buf) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ny) return;
    const double* row = u + static_cast<int64_t>(j) * pitch;
    buf[j] = row[0];
    buf[ny + j] = row[nx - 1];
}
// recv buffers: [2ny,3ny) → ghost column x=-1, [3ny,4ny) → ghost column x=nx
__global__ void k_unpack_columns(double* __restrict__ u, int nx, int ny, int64_t pitch,
                                 const double* __restrict__ buf, int has_left, int has_right) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ny) return;
    double* row = u + static_cast<int64_t>(j) * pitch;
    if (has_left) row[-1] = buf[2 * ny + j];
    if (has_right) row[nx] = buf[3 * ny + j];
}

// ---- wide exchange: T ghost lines from all eight neighbours, for the temporally blocked sweep ------
// A sweep of T steps needs T valid lines around the tile, corners included.  Each rank packs eight
// regions of its own tile (four bands of T lines, four T x T corners) into one staging buffer, all
// sixteen transfers go out as one NCCL group, and one kernel scatters the received regions into the
// ghost area.  Bands span the physical ghost line of a perpendicular physical side as well, because
// "periodic" ghosts are frozen values that the neighbour's halo cells depend on.
struct XRegion {
    int x0, y0, w, h;  // interior coordinates of the region's first cell, extent
    long long off;     // offset of the region in the staging buffer (doubles)
    int peer;          // rank on the other side, -1: unused
};
struct XTable {
    XRegion r[8];
};

__global__ void __launch_bounds__(256) k_pack_regions(const double* __restrict__ u, long long pitch, XTable t,
                                                      double* __restrict__ buf) {
    const XRegion g = t.r[blockIdx.y];
    if (g.peer < 0) return;
    const int n = g.w * g.h;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int yy = e / g.w, xx = e - yy * g.w;
        buf[g.off + e] = u[static_cast<long long>(g.y0 + yy) * pitch + g.x0 + xx];
    }
}
__global__ void __launch_bounds__(256) k_unpack_regions(double* __restrict__ u, long long pitch, XTable t,
                                                        const double* __restrict__ buf) {
    const XRegion g = t.r[blockIdx.y];
    if (g.peer < 0) return;
    const int n = g.w * g.h;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int yy = e / g.w, xx = e - yy * g.w;
        u[static_cast<long long>(g.y0 + yy) * pitch + g.x0 + xx] = buf[g.off + e];
    }
}

// Staging of the wide exchange: 8 send + 8 receive regions of up to T lines.  Sized before the block
// loop (never inside a stream capture).
static int ensure_wide(csim_ctx* c, const csim_field* f, const csim_decomp* dec, int T) {
    csim_decomp d = *dec;
    d.nx_local = f->nx;
    d.ny_local = f->ny;
    csim_xregion ps[8], pr8[8];
    if (int rc = csim_wide_exchange_plan(&d, T, ps, pr8)) return rc;
    size_t total = 0;
    for (int q = 0; q < 8; ++q) total += static_cast<size_t>(ps[q].w) * ps[q].h;
    if (c->wide_doubles >= 2 * total) return CSIM_OK;
    if (c->d_wide) {
        CSIM_CUDA(cudaDeviceSynchronize());
        CSIM_CUDA(cudaFree(c->d_wide));
        c->d_wide = nullptr;
        c->wide_doubles = 0;
    }
    CSIM_CUDA(cudaMalloc(&c->d_wide, 2 * total * sizeof(double)));
    c->wide_doubles = 2 * total;
    return CSIM_OK;
}

// Fill T ghost lines of `f` on every side that has a neighbour, on `stream`.  *bytes_sent (optional)
// receives what this rank puts on the wire.
static int wide_exchange(csim_field* f, const csim_decomp* dec, int T, cudaStream_t stream, size_t* bytes_sent) {
    csim_ctx* c = f->ctx;
    csim_decomp d = *dec;  // the plan is a function of the decomposition; the tile fixes the local size
    d.nx_local = f->nx;
    d.ny_local = f->ny;
    csim_xregion ps[8], pr8[8];
    if (int rc = csim_wide_exchange_plan(&d, T, ps, pr8)) return rc;
    XTable snd, rcv;
    long long off = 0, wire = 0;
    for (int q = 0; q < 8; ++q) {
        snd.r[q] = XRegion{ps[q].x0, ps[q].y0, ps[q].w, ps[q].h, off, ps[q].peer};
        rcv.r[q] = XRegion{pr8[q].x0, pr8[q].y0, pr8[q].w, pr8[q].h, 0, pr8[q].peer};
        off += static_cast<long long>(ps[q].w) * ps[q].h;
        if (ps[q].peer >= 0) wire += static_cast<long long>(ps[q].w) * ps[q].h;
    }
    const long long send_total = off;
    for (int q = 0; q < 8; ++q) rcv.r[q].off = send_total + snd.r[q].off;
    CSIM_REQUIRE(c->wide_doubles >= static_cast<size_t>(2 * send_total), CSIM_ERR_INVALID,
                 "wide_exchange: staging buffer not sized (ensure_wide)");
    if (bytes_sent) *bytes_sent = static_cast<size_t>(wire) * sizeof(double);
    double* buf = c->d_wide;
    const dim3 grid(32, 8);
    k_pack_regions<<<grid, 256, 0, stream>>>(f->interior(), f->pitch, snd, buf);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
    ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
    CSIM_NCCL(g_nccl.GroupStart());
    for (int q = 0; q < 8; ++q) {
        if (snd.r[q].peer < 0) continue;
        const size_t n = static_cast<size_t>(snd.r[q].w) * snd.r[q].h;
        CSIM_NCCL(g_nccl.Recv(buf + rcv.r[q].off, n, ncclDouble, rcv.r[q].peer, comm, stream));
        CSIM_NCCL(g_nccl.Send(buf + snd.r[q].off, n, ncclDouble, snd.r[q].peer, comm, stream));
    }
    CSIM_NCCL(g_nccl.GroupEnd());
    ++c->launches;
    k_unpack_regions<<<grid, 256, 0, stream>>>(f->interior(), f->pitch, rcv, buf);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
    return CSIM_OK;
}

// ---- the block loop of csim_run_steps, its CUDA-graph replay and its timeline -----------------------

// Everything a captured block loop depends on; two calls with equal keys enqueue identical work.
struct RunKey {
    const double* u;
    const double* tmp;
    int nsteps, maxT, mode, zero_terms;
    csim_step_params p;
    csim_decomp d;
};
static bool same_key(const RunKey& a, const RunKey& b) { return std::memcmp(&a, &b, sizeof(RunKey)) == 0; }

struct RunGraph {
    RunKey key;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;  // kernels one replay launches
    int swaps = 0;          // buffer swaps one replay stands for
    uint64_t last_use = 0;
};

// Timeline of one profiled call (csim_halo_profile): per block, timestamps around the exchange and the
// frame sweep on the exchange stream and around the interior sweep on the main stream.
struct RunProfile {
    bool on = false;
    std::vector<cudaEvent_t> ev;  // base, then 6 per block: x0 x1 f1 i0 i1 (f0 == x1) + spare
    size_t used = 0;
    int blocks = 0;
    size_t bytes_per_exchange = 0;
    cudaEvent_t next(csim_ctx* c) {
        if (used == ev.size()) {
            cudaEvent_t e;
            cudaSetDevice(c->device);
            cudaEventCreate(&e);
            ev.push_back(e);
        }
        return ev[used++];
    }
};

struct RunState {
    std::vector<RunGraph> graphs;
    uint64_t tick = 0;
    bool comm_warm = false;  // one eager pass has set up NCCL's connections to every neighbour
    RunProfile prof;
    csim_halo_stats last{};
};

static RunState* run_state(csim_ctx* c) {
    if (!c->run_state) c->run_state = new RunState();
    return static_cast<RunState*>(c->run_state);
}

void run_state_destroy(csim_ctx* c) {
    RunState* rs = static_cast<RunState*>(c->run_state);
    if (!rs) return;
    for (RunGraph& g : rs->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (cudaEvent_t e : rs->prof.ev) cudaEventDestroy(e);
    delete rs;
    c->run_state = nullptr;
}

// Software pipeline over blocks of T steps, two streams:
//   exchange stream (high priority): go(n) → frame(n) → exchange(n+1)
//   main stream                    : wait go(n) → interior(n)
// frame(n) = the work items that read ghost lines (edge strips, first/last chunk of every strip);
// interior(n) = all the others.  The bands exchange(n+1) packs are all produced by frame(n), so
// the next block's halos travel while interior(n) runs and are in place when block n+1 starts.
//   frame(n)    needs exchange(n) (same stream) and interior(n-1) (event ev_fork)
//   interior(n) needs frame(n-1) and interior(n-1); it is released by the event go(n), recorded
//               on the exchange stream right before frame(n), so that both kernels become
//               eligible together and the high-priority frame blocks are placed first.  Released
//               by stream order alone, the interior blocks fill every SM a few microseconds
//               before the frame's event arrives and the frame waits a whole round for slots:
//               measured chain frame-wait 88 + frame 88 + exchange 132 us = 308 us per block
//               against 285 us of work (profiles/r01_multigpu_phases.md).
// Swaps u and tmp once per block (host bookkeeping); *swaps returns how often.
static int enqueue_blocks(csim_field* u, csim_field* tmp, const csim_step_params* p, const csim_decomp* dec,
                          const StepK& k, int mode, int maxT, int nsteps, bool zero_terms, int values_after,
                          RunProfile* prof, int* swaps) {
    csim_ctx* c = u->ctx;
    int left = nsteps;
    int T = left < maxT ? left : maxT;
    *swaps = 0;
    auto stamp = [&](cudaStream_t st) -> int {
        if (!prof) return CSIM_OK;
        CSIM_CUDA(cudaEventRecord(prof->next(c), st));
        return CSIM_OK;
    };
    CSIM_CUDA(cudaEventRecord(c->ev_fork, c->stream));  // everything queued so far = "interior(-1)"
    CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));
    if (int rc = stamp(c->stream_x)) return rc;  // x0(0)
    size_t wire = 0;
    if (int rc = wide_exchange(u, dec, T, c->stream_x, &wire)) return rc;  // exchange(0): the only one not hidden
    if (prof) prof->bytes_per_exchange = wire;
    bool first = true;
    while (left > 0) {
        bool launched = false;
        if (!first) CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));  // interior(n-1) done
        if (int rc = stamp(c->stream_x)) return rc;                              // x1(n) = f0(n)
        CSIM_CUDA(cudaEventRecord(c->ev_go, c->stream_x));                       // go(n)
        if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_FRAME, c->stream_x, &launched, zero_terms)) return rc;
        CSIM_CUDA(cudaEventRecord(c->ev_join, c->stream_x));                     // frame(n) done
        if (int rc = stamp(c->stream_x)) return rc;                              // f1(n)
        CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_go, 0));
        if (int rc = stamp(c->stream)) return rc;                                // i0(n)
        if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_INTERIOR, c->stream, &launched, zero_terms)) return rc;
        if (int rc = stamp(c->stream)) return rc;                                // i1(n)
        CSIM_CUDA(cudaEventRecord(c->ev_fork, c->stream));                       // interior(n) done
        tmp->values = values_after;
        csim_field_swap(u, tmp);
        ++*swaps;
        if (prof) ++prof->blocks;
        left -= T;
        first = false;
        if (left > 0) {
            T = left < maxT ? left : maxT;
            if (int rc = stamp(c->stream_x)) return rc;                          // x0(n+1)
            if (int rc = wide_exchange(u, dec, T, c->stream_x, nullptr)) return rc;  // exchange(n+1): reads frame(n)'s cells
        }
    }
    CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));  // the main stream orders everything again
    return CSIM_OK;
}

// Turn the recorded timeline into csim_halo_stats (host-synchronous; profiled calls only).
static int finish_profile(csim_ctx* c, RunState* rs) {
    RunProfile& pr = rs->prof;
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream_x));
    csim_halo_stats st{};
    st.blocks = pr.blocks;
    st.bytes_per_exchange = pr.bytes_per_exchange;
    // event layout: base, x0(0), then per block n: x1(n) f1(n) i0(n) i1(n) [x0(n+1) unless last]
    std::vector<double> t(pr.used, 0.0);
    for (size_t i = 1; i < pr.used; ++i) {
        float ms = 0.f;
        CSIM_CUDA(cudaEventElapsedTime(&ms, pr.ev[0], pr.ev[i]));
        t[i] = ms;
    }
    double ex = 0.0, hidden = 0.0, frame = 0.0, interior = 0.0;
    int n_hidden = 0;
    size_t i = 1;  // t[0] is the time base
    double x0 = t[i++];
    double prev_i0 = 0.0, prev_i1 = 0.0;
    for (int n = 0; n < pr.blocks && i + 3 < pr.used; ++n) {
        const double x1 = t[i], f1 = t[i + 1], i0 = t[i + 2], i1 = t[i + 3];
        i += 4;
        const double dur = x1 - x0;
        if (n == 0) {
            st.first_exchange_us = 1e3 * dur;
        } else {  // exchange(n) ran beside interior(n-1): how much of it lies inside that interval
            ex += dur;
            const double lo = x0 > prev_i0 ? x0 : prev_i0, hi = x1 < prev_i1 ? x1 : prev_i1;
            if (hi > lo) hidden += hi - lo;
            ++n_hidden;
        }
        frame += f1 - x1;
        interior += i1 - i0;
        prev_i0 = i0;
        prev_i1 = i1;
        if (n + 1 < pr.blocks && i < pr.used) x0 = t[i++];
    }
    st.exchange_us = n_hidden ? 1e3 * ex / n_hidden : st.first_exchange_us;
    st.overlap_fraction = ex > 0.0 ? hidden / ex : 0.0;
    st.frame_us = pr.blocks ? 1e3 * frame / pr.blocks : 0.0;
    st.interior_us = pr.blocks ? 1e3 * interior / pr.blocks : 0.0;
    st.total_ms = pr.used ? t[pr.used - 1] : 0.0;
    rs->last = st;
    pr.on = false;
    return CSIM_OK;
}

}  // namespace csim

using namespace csim;

extern "C" {

int csim_comm_unique_id(char id[CSIM_UNIQUE_ID_BYTES]) {
    CSIM_REQUIRE(id != nullptr, CSIM_ERR_INVALID, "csim_comm_unique_id: null argument");
    static_assert(sizeof(ncclUniqueId) == CSIM_UNIQUE_ID_BYTES, "ncclUniqueId size changed");
    if (int rc = load_nccl()) return rc;
    ncclUniqueId uid;
    CSIM_NCCL(g_nccl.GetUniqueId(&uid));
    std::memcpy(id, &uid, sizeof uid);
    return CSIM_OK;
}

int csim_comm_init(csim_ctx* c, int size, int rank, const char id[CSIM_UNIQUE_ID_BYTES]) {
    CSIM_REQUIRE(c != nullptr && id != nullptr, CSIM_ERR_INVALID, "csim_comm_init: null argument");
    CSIM_REQUIRE(size >= 1 && rank >= 0 && rank < size, CSIM_ERR_INVALID, "csim_comm_init: bad size/rank");
    CSIM_REQUIRE(c->comm == nullptr, CSIM_ERR_INVALID, "csim_comm_init: communicator already initialised");
    if (int rc = load_nccl()) return rc;
    CSIM_CUDA(cudaSetDevice(c->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    ncclComm_t comm;
    CSIM_NCCL(g_nccl.CommInitRank(&comm, size, uid, rank));
    c->comm = comm;
    c->comm_size = size;
    c->comm_rank = rank;
    return CSIM_OK;
}

int csim_comm_destroy(csim_ctx* c) {
    CSIM_REQUIRE(c != nullptr, CSIM_ERR_INVALID, "csim_comm_destroy: ctx is null");
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(static_cast<ncclComm_t>(c->comm));
        c->comm = nullptr;
    }
    return CSIM_OK;
}

int csim_comm_allreduce_max(csim_ctx* c, double* inout, int n) {
    CSIM_REQUIRE(c != nullptr && n >= 0 && (n == 0 || inout != nullptr), CSIM_ERR_INVALID,
                 "csim_comm_allreduce_max: bad arguments");
    CSIM_REQUIRE(static_cast<size_t>(n) + 1 <= c->scratch_doubles, CSIM_ERR_INVALID,
                 "csim_comm_allreduce_max: too many elements");
    CSIM_CUDA(cudaSetDevice(c->device));
    if (!c->comm) {
        CSIM_CUDA(cudaStreamSynchronize(c->stream));
        return CSIM_OK;
    }
    const size_t cnt = static_cast<size_t>(n) + 1;  // one dummy element so that n == 0 is a barrier
    for (int k = 0; k < n; ++k) c->h_scratch[k] = inout[k];
    c->h_scratch[n] = 0.0;
    CSIM_CUDA(cudaMemcpyAsync(c->d_scratch, c->h_scratch, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CSIM_NCCL(g_nccl.AllReduce(c->d_scratch, c->d_scratch, cnt, ncclDouble, ncclMax, static_cast<ncclComm_t>(c->comm),
                               c->stream));
    ++c->launches;
    CSIM_CUDA(cudaMemcpyAsync(c->h_scratch, c->d_scratch, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n; ++k) inout[k] = c->h_scratch[k];
    return CSIM_OK;
}

int csim_halo_exchange(csim_field* f, const csim_decomp* dec) {
    CSIM_REQUIRE(f != nullptr && dec != nullptr, CSIM_ERR_INVALID, "csim_halo_exchange: null argument");
    const int left = dec->nbr[CSIM_LEFT], right = dec->nbr[CSIM_RIGHT];
    const int down = dec->nbr[CSIM_BOTTOM], up = dec->nbr[CSIM_TOP];
    if (left == CSIM_PROC_NULL && right == CSIM_PROC_NULL && down == CSIM_PROC_NULL && up == CSIM_PROC_NULL)
        return CSIM_OK;  // rcount == 0, src/halo.cpp:45
    CSIM_REQUIRE(f->h == 1, CSIM_ERR_UNSUPPORTED, "csim_halo_exchange: needs halo == 1 (main.cpp:65)");
    csim_ctx* c = f->ctx;
    CSIM_REQUIRE(c->comm != nullptr, CSIM_ERR_COMM, "csim_halo_exchange: tile has neighbours but no communicator");
    f->values = csim_field::kUnknown;  // ghost lines now hold a neighbour's cells
    for (int s = 0; s < 4; ++s)
        CSIM_REQUIRE(dec->nbr[s] == CSIM_PROC_NULL || (dec->nbr[s] >= 0 && dec->nbr[s] < c->comm_size),
                     CSIM_ERR_INVALID, "csim_halo_exchange: neighbour rank outside the communicator");
    CSIM_CUDA(cudaSetDevice(c->device));
    const int nx = f->nx, ny = f->ny, nxt = f->nxt();
    if (c->pack_doubles < static_cast<size_t>(4) * ny) {
        if (c->d_pack) {
            CSIM_CUDA(cudaStreamSynchronize(c->stream));
            CSIM_CUDA(cudaFree(c->d_pack));
            c->d_pack = nullptr;
        }
        CSIM_CUDA(cudaMalloc(&c->d_pack, static_cast<size_t>(4) * ny * sizeof(double)));
        c->pack_doubles = static_cast<size_t>(4) * ny;
    }
    double* buf = c->d_pack;
    double* in = f->interior();
    const bool cols = (left != CSIM_PROC_NULL || right != CSIM_PROC_NULL) && ny > 0;
    if (cols) CSIM_LAUNCH(c, k_pack_columns, (ny + 255) / 256, 256, 0, in, nx, ny, f->pitch, buf);

    ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
    const size_t n_col = static_cast<size_t>(ny), n_row = static_cast<size_t>(nxt);
    CSIM_NCCL(g_nccl.GroupStart());
    if (left != CSIM_PROC_NULL && n_col) {  // halo.cpp:28-31
        CSIM_NCCL(g_nccl.Recv(buf + 2 * n_col, n_col, ncclDouble, left, comm, c->stream));
        CSIM_NCCL(g_nccl.Send(buf, n_col, ncclDouble, left, comm, c->stream));
    }
    if (right != CSIM_PROC_NULL && n_col) {  // halo.cpp:32-35
        CSIM_NCCL(g_nccl.Recv(buf + 3 * n_col, n_col, ncclDouble, right, comm, c->stream));
        CSIM_NCCL(g_nccl.Send(buf + n_col, n_col, ncclDouble, right, comm, c->stream));
    }
    if (down != CSIM_PROC_NULL) {  // halo.cpp:36-39: ghost row j=0 ← down; send interior row j=h
        CSIM_NCCL(g_nccl.Recv(f->at(0, 0), n_row, ncclDouble, down, comm, c->stream));
        CSIM_NCCL(g_nccl.Send(f->at(0, 1), n_row, ncclDouble, down, comm, c->stream));
    }
    if (up != CSIM_PROC_NULL) {  // halo.cpp:40-43: ghost row j=h+ny ← up; send interior row j=h+ny-1
        CSIM_NCCL(g_nccl.Recv(f->at(0, 1 + ny), n_row, ncclDouble, up, comm, c->stream));
        CSIM_NCCL(g_nccl.Send(f->at(0, ny), n_row, ncclDouble, up, comm, c->stream));
    }
    CSIM_NCCL(g_nccl.GroupEnd());
    c->launches += 1;  // the grouped NCCL transfer is one fused device kernel
    if (cols)
        CSIM_LAUNCH(c, k_unpack_columns, (ny + 255) / 256, 256, 0, in, nx, ny, f->pitch, buf,
                    left != CSIM_PROC_NULL, right != CSIM_PROC_NULL);
    return CSIM_OK;
}

int csim_run_steps(csim_field* u, csim_field* tmp, const csim_step_params* p, const csim_decomp* dec,
                   int nsteps) {
    CSIM_REQUIRE(u != nullptr && tmp != nullptr && p != nullptr, CSIM_ERR_INVALID, "csim_run_steps: null argument");
    CSIM_REQUIRE(nsteps >= 0, CSIM_ERR_INVALID, "csim_run_steps: negative step count");
    bool has_nbr = false;
    if (dec)
        for (int s = 0; s < 4; ++s) has_nbr = has_nbr || dec->nbr[s] != CSIM_PROC_NULL;
    if (!has_nbr) return csim_step_fused(u, tmp, p, nsteps);
    for (int s = 0; s < 4; ++s)
        CSIM_REQUIRE(p->nbr[s] == dec->nbr[s], CSIM_ERR_INVALID,
                     "csim_run_steps: step params and decomposition disagree on neighbours");
    csim_ctx* c = u->ctx;
    CSIM_REQUIRE(c == tmp->ctx && u->nx == tmp->nx && u->ny == tmp->ny && u->h == tmp->h && u->base != tmp->base,
                 CSIM_ERR_INVALID, "csim_run_steps: fields differ in geometry or alias");
    CSIM_REQUIRE(u->h == 1, CSIM_ERR_UNSUPPORTED, "csim_run_steps: needs halo == 1 (main.cpp:65)");
    CSIM_REQUIRE(c->comm != nullptr, CSIM_ERR_COMM, "csim_run_steps: tile has neighbours but no communicator");
    CSIM_CUDA(cudaSetDevice(c->device));
    StepK k;
    int mode = 0;
    if (int rc = step_setup(u, p, &k, &mode)) return rc;
    // Blocking T steps needs T lines from every neighbour and a tile at least T cells wide (every
    // rank's tile: the last rank of a dimension is never the smallest, decomp.cpp:29-30).
    const int min_nx = dec->nx_global / dec->dims[0], min_ny = dec->ny_global / dec->dims[1];
    int maxT = (p->flags & CSIM_STEP_NO_TEMPORAL) ? 1 : (mode == MODE_DIV ? tb_max_T_div() : tb_max_T());
    if (maxT > min_nx) maxT = min_nx;
    if (maxT > min_ny) maxT = min_ny;
    if (maxT < 1 || (p->flags & CSIM_STEP_NO_TEMPORAL)) {
        // reference-shaped path: one-line exchange, then one step, every step
        for (int n = 0; n < nsteps; ++n) {
            if (int rc = csim_halo_exchange(u, dec)) return rc;     // main.cpp:101
            if (int rc = csim_step_fused(u, tmp, p, 1)) return rc;  // main.cpp:102-109
        }
        return CSIM_OK;
    }
    // one decision for the whole call, taken per rank (see resolve_zero_terms)
    bool zero_terms = false;
    if (nsteps >= maxT)
        if (int rc = resolve_zero_terms(u, p, k, mode, maxT, &zero_terms)) return rc;
    const int values_after = zero_terms ? csim_field::kClean
                                        : (u->values == csim_field::kTainted ? csim_field::kTainted : csim_field::kUnknown);
    if (nsteps == 0) return CSIM_OK;
    if (int rc = ensure_wide(c, u, dec, maxT)) return rc;
    RunState* rs = run_state(c);
    int swaps = 0;
    if (rs->prof.on) {  // profiled call: eager, with timestamps (csim_halo_profile)
        rs->prof.used = 0;
        rs->prof.blocks = 0;
        CSIM_CUDA(cudaEventRecord(rs->prof.next(c), c->stream));  // time base
        if (int rc = enqueue_blocks(u, tmp, p, dec, k, mode, maxT, nsteps, zero_terms, values_after, &rs->prof, &swaps))
            return rc;
        rs->comm_warm = true;
        return finish_profile(c, rs);
    }
    // The block loop is ~8 launches per block (pack, NCCL group, unpack, frame, interior + events): at
    // 34 blocks per 100 steps the host spent as long enqueueing them as the GPUs spent running them
    // (round 1: 7.4 ms against 7.9 ms at 8192^2 per GPU).  So the loop is captured into a CUDA graph the
    // first time a (tiles, parameters, step count) combination is seen and replayed afterwards; the
    // very first call of a communicator runs eagerly, because NCCL sets up its connections to the
    // neighbours on first use, which must not happen inside a capture.  CSIM_GRAPH=0 disables it.
    static const bool graphs_on = [] {
        const char* e = std::getenv("CSIM_GRAPH");
        return !(e && std::strcmp(e, "0") == 0);
    }();
    if (!graphs_on || !rs->comm_warm) {
        if (int rc = enqueue_blocks(u, tmp, p, dec, k, mode, maxT, nsteps, zero_terms, values_after, nullptr, &swaps))
            return rc;
        rs->comm_warm = true;
        return CSIM_OK;
    }
    RunKey key;
    std::memset(&key, 0, sizeof key);  // padding bytes take part in the comparison
    key.u = u->base;
    key.tmp = tmp->base;
    key.nsteps = nsteps;
    key.maxT = maxT;
    key.mode = mode;
    key.zero_terms = zero_terms ? 1 : 0;
    std::memcpy(&key.p, p, sizeof key.p);
    std::memcpy(&key.d, dec, sizeof key.d);
    ++rs->tick;
    for (RunGraph& g : rs->graphs)
        if (same_key(g.key, key)) {
            CSIM_CUDA(cudaGraphLaunch(g.exec, c->stream));
            g.last_use = rs->tick;
            c->launches += g.launches;
            if (g.swaps & 1) csim_field_swap(u, tmp);
            u->values = values_after;  // the newest state was written by the sweep
            return CSIM_OK;
        }
    const uint64_t launches0 = c->launches;
    cudaGraph_t graph = nullptr;
    CSIM_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_blocks(u, tmp, p, dec, k, mode, maxT, nsteps, zero_terms, values_after, nullptr, &swaps);
    const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc != CSIM_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture(csim_run_steps)", __FILE__, __LINE__);
    RunGraph g;
    g.key = key;
    g.launches = c->launches - launches0;
    g.swaps = swaps;
    g.last_use = rs->tick;
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return cuda_fail(ie, "cudaGraphInstantiate(csim_run_steps)", __FILE__, __LINE__);
    if (rs->graphs.size() >= 16) {  // keep the cache small: drop the least recently used graph
        size_t victim = 0;
        for (size_t i = 1; i < rs->graphs.size(); ++i)
            if (rs->graphs[i].last_use < rs->graphs[victim].last_use) victim = i;
        cudaGraphExecDestroy(rs->graphs[victim].exec);
        rs->graphs.erase(rs->graphs.begin() + static_cast<long>(victim));
    }
    rs->graphs.push_back(g);
    CSIM_CUDA(cudaGraphLaunch(g.exec, c->stream));  // the capture recorded the work; this runs it
    return CSIM_OK;
}

int csim_halo_profile(csim_ctx* c, int enable) {
    CSIM_REQUIRE(c != nullptr, CSIM_ERR_INVALID, "csim_halo_profile: ctx is null");
    run_state(c)->prof.on = enable != 0;
    return CSIM_OK;
}

int csim_halo_stats_get(csim_ctx* c, csim_halo_stats* out) {
    CSIM_REQUIRE(c != nullptr && out != nullptr, CSIM_ERR_INVALID, "csim_halo_stats_get: null argument");
    *out = run_state(c)->last;
    return CSIM_OK;
}

}  // extern "C"
