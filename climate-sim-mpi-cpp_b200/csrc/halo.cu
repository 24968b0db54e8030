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

#include <unistd.h>

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

// ---- peer-memory exchange (CSIM_HALO=peer): the bands go straight into the neighbours' ghost lines ----
// Why these kernels are "light".  The interior sweep holds 3 CTAs x 128 threads x 168 registers = 64 512 of
// an SM's 65 536 registers for its whole duration, and with 320-row chunks a CTA lives ~250 us.  A helper
// kernel that needs more than the 1 024 registers left over has to wait for interior CTAs to exit; a CTA of
// ONE warp with at most 32 registers per thread is exactly 1 024 registers and can start beside a full sweep.
// What was measured (profiles/r02_multigpu.md): on 2 GPUs at 8192^2 per GPU this path beats the NCCL path
// (1.905e12 against 1.845e12); its own store kernel takes 34 us there but 497 us at 16384^2 per GPU, where a
// column band spans 2.1 GB of strided rows, and on 8 GPUs (five neighbours, 1.6 MB per exchange) the exchange no
// longer fits behind the interior sweep and the path collapses — which is why NCCL is the default.
struct PushRegion {
    int x0, y0, w, h;     // source region in this rank's tile (interior coordinates)
    double* dst;          // neighbour's cell that receives the region's first cell (mapped pointer)
    long long dst_pitch;  // neighbour's row pitch
    unsigned* flag;       // neighbour's flag word for this direction, nullptr: no neighbour
};
struct PushTable {
    PushRegion r[8];
};
// words of csim_ctx::d_flags: [0..7] "halo arrived" flags written by the neighbours, then this rank's own
// counters, then [16..23] "ready to receive" flags written by the neighbours
enum { kCtlTicket = 8, kCtlPushSeq = 9, kCtlWaitSeq = 10, kCtlReadySeq = 11, kCtlReadyFlags = 16, kCtlWords = 32 };

// Store all regions into the neighbours' tiles; the last CTA to finish (ticket) publishes the sequence
// number of this push in every neighbour's flag word.  The sequence number lives in device memory so
// that a CUDA-graph replay pushes the right one.
// The work is cut into chunks of 32 x kPushBatch cells across all eight regions, and a warp keeps a whole
// chunk in flight: kPushBatch independent loads per lane, then as many remote stores.  With one load and
// one store at a time (the first version) every cell paid a full round trip through a memory system that
// the interior sweep keeps saturated: 1.4 GB/s, 1.1 ms per exchange on 8 GPUs (profiles/r02_multigpu.md).
constexpr int kPushBatch = 8;
__global__ void __maxnreg__(32) k_push_light(const double* __restrict__ u, long long pitch, PushTable t,
                                             unsigned* __restrict__ ctl) {
    const int lane = threadIdx.x;
    constexpr int kChunk = 32 * kPushBatch;
    int chunk0 = 0;  // first global chunk id of region q
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const PushRegion g = t.r[q];
        if (!g.flag) continue;
        const int n = g.w * g.h;
        const int nchunks = (n + kChunk - 1) / kChunk;
        // chunks are dealt round-robin over the CTAs, continuing across regions
        int first = static_cast<int>(blockIdx.x) - chunk0 % static_cast<int>(gridDim.x);
        if (first < 0) first += gridDim.x;
#pragma unroll 1
        for (int ch = first; ch < nchunks; ch += gridDim.x) {
            const int e0 = ch * kChunk + lane;
            double v[kPushBatch];
#pragma unroll
            for (int k = 0; k < kPushBatch; ++k) {
                const int e = e0 + 32 * k;
                const int yy = e / g.w, xx = e - yy * g.w;
                v[k] = e < n ? u[static_cast<long long>(g.y0 + yy) * pitch + g.x0 + xx] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kPushBatch; ++k) {
                const int e = e0 + 32 * k;
                const int yy = e / g.w, xx = e - yy * g.w;
                if (e < n) g.dst[static_cast<long long>(yy) * g.dst_pitch + xx] = v[k];
            }
        }
        chunk0 += nchunks;
    }
    __threadfence_system();  // this thread's peer stores are visible system-wide before the ticket
    __syncwarp();
    unsigned last = 0, seq = 0;
    if (lane == 0) {
        last = atomicAdd(&ctl[kCtlTicket], 1u) == gridDim.x - 1 ? 1u : 0u;
        if (last) {
            ctl[kCtlTicket] = 0;  // re-arm for the next push (stream order protects it)
            seq = ctl[kCtlPushSeq] + 1u;
            ctl[kCtlPushSeq] = seq;
            __threadfence_system();
        }
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    seq = __shfl_sync(0xffffffffu, seq, 0);
    if (last && lane < 8 && t.r[lane].flag) *reinterpret_cast<volatile unsigned*>(t.r[lane].flag) = seq;
}

// Gate of the frame sweep: wait until every neighbour in `mask` has published this exchange's sequence
// number.  Bounded: after `timeout_ns` the kernel records which neighbour is missing and traps, which
// fails every later call on the context instead of sweeping over stale ghost lines.
__global__ void __maxnreg__(32) k_wait_light(const unsigned* flags, unsigned mask,
                                                                   unsigned* __restrict__ ctl,
                                                                   unsigned long long timeout_ns, unsigned* err) {
    const int k = threadIdx.x;
    unsigned seq = 0;
    if (k == 0) {
        seq = ctl[kCtlWaitSeq] + 1u;
        ctl[kCtlWaitSeq] = seq;
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    if (k < 8 && ((mask >> k) & 1)) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const volatile unsigned* f = flags + k;
        // seq wraps after 2^32 exchanges; compare as a signed distance
        while (static_cast<int>(*f - seq) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                *reinterpret_cast<volatile unsigned*>(err) = 1u + static_cast<unsigned>(k);
                __threadfence_system();
                __trap();
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

static void drop_graphs(csim_ctx* c);  // below, with the graph cache

// Handshake at the start of every csim_run_steps call.  Inside a call the double buffering orders the remote
// stores (see peer_exchange), but the FIRST push of a call may reach a neighbour that has not entered the
// call yet and is still writing the same tile from its side — an upload, a fill, a stand-alone kernel — which
// would overwrite the ghost lines just stored (bench.py's parity windows caught exactly that at 16384^2).
// So every rank first tells its neighbours "everything I queued on this tile before the call is done"
// (stream order) and waits for the same from them: a barrier among neighbours, one light kernel per call.
struct ReadyTable {
    unsigned* flag[8];  // neighbour's ready word for my direction, nullptr: no neighbour
};
__global__ void __maxnreg__(32) k_ready_light(ReadyTable t, unsigned* __restrict__ ctl, unsigned long long timeout_ns,
                                              unsigned* err) {
    const int k = threadIdx.x;
    unsigned seq = 0;
    if (k == 0) {
        seq = ctl[kCtlReadySeq] + 1u;
        ctl[kCtlReadySeq] = seq;
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    __threadfence_system();
    if (k < 8 && t.flag[k]) {
        *reinterpret_cast<volatile unsigned*>(t.flag[k]) = seq;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const volatile unsigned* f = ctl + kCtlReadyFlags + k;
        while (static_cast<int>(*f - seq) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                *reinterpret_cast<volatile unsigned*>(err) = 101u + static_cast<unsigned>(k);
                __threadfence_system();
                __trap();
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

static unsigned long long peer_timeout_ns() {
    static const unsigned long long v = [] {
        const char* e = std::getenv("CSIM_HALO_TIMEOUT_S");
        const double s = e ? std::atof(e) : 60.0;
        return static_cast<unsigned long long>((s > 0.0 ? s : 60.0) * 1e9);
    }();
    return v;
}

// Queue the neighbour barrier on `stream` (peer path only).
static int peer_ready_barrier(csim_ctx* c, cudaStream_t stream) {
    ReadyTable t;
    for (int q = 0; q < 8; ++q) {
        const csim_ctx::PeerLink& L = c->peer[q];
        t.flag[q] = (L.rank >= 0 && L.flags) ? L.flags + kCtlReadyFlags + (7 - q) : nullptr;
    }
    k_ready_light<<<1, 32, 0, stream>>>(t, c->d_flags, peer_timeout_ns(), c->d_err);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
    return CSIM_OK;
}

// CSIM_HALO=peer opts into the peer-store path; the default is the NCCL path, which holds ~100 % weak-scaling
// efficiency at 8 GPUs where the peer path, better at 2 GPUs, collapses (profiles/r02_multigpu.md).
static bool peer_path_wanted() {
    static const bool peer = [] {
        const char* e = std::getenv("CSIM_HALO");
        return e && (std::strcmp(e, "peer") == 0 || std::strcmp(e, "p2p") == 0);
    }();
    return peer;
}
// CSIM_CARVEOUT=<percent>: pin the shared-memory carve-out of the light kernels (and, in kernels.cu, of the
// staged sweep) to the same value — a measurement aid for the question whether kernels with different
// carve-outs can share an SM.
static int light_carveout() {
    static const int v = [] {
        const char* e = std::getenv("CSIM_CARVEOUT");
        return e ? std::atoi(e) : -1;
    }();
    return v;
}
static void light_prepare() {
    static const bool done = [] {
        const int pct = light_carveout();
        if (pct >= 0) {
            cudaFuncSetAttribute(k_push_light, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(k_wait_light, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(k_ready_light, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        return true;
    }();
    (void)done;
}
static bool peer_tiles_match(const csim_ctx* c, const csim_field* u, const csim_field* tmp) {
    return c->peer_ready && ((u->base == c->peer_tile[0] && tmp->base == c->peer_tile[1]) ||
                             (u->base == c->peer_tile[1] && tmp->base == c->peer_tile[0]));
}

// Peer version of wide_exchange: store this rank's bands of tile `f` into the neighbours' copy of the
// same tile (their u when f is our u: every rank swaps in step), then gate `stream` on the neighbours'
// stores into ours.  Double buffering makes the remote stores safe: exchange(n+1) writes the buffer whose
// ghost lines the neighbour last read in frame(n-1), and we only get here after the neighbour's push(n),
// which it queued behind its frame(n-1).
static int peer_exchange(csim_field* f, const csim_decomp* dec, int T, cudaStream_t stream, size_t* bytes_sent,
                         cudaEvent_t after_push = nullptr) {
    csim_ctx* c = f->ctx;
    light_prepare();
    csim_decomp d = *dec;
    d.nx_local = f->nx;
    d.ny_local = f->ny;
    csim_xregion ps[8], pr8[8];
    if (int rc = csim_wide_exchange_plan(&d, T, ps, pr8)) return rc;
    const int slot = f->base == c->peer_tile[0] ? 0 : 1;
    const bool pl = dec->nbr[CSIM_LEFT] == CSIM_PROC_NULL, pb = dec->nbr[CSIM_BOTTOM] == CSIM_PROC_NULL;
    PushTable t;
    unsigned mask = 0;
    long long wire = 0;
    int q = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            PushRegion& g = t.r[q];
            g.x0 = ps[q].x0;
            g.y0 = ps[q].y0;
            g.w = ps[q].w;
            g.h = ps[q].h;
            g.dst = nullptr;
            g.dst_pitch = 0;
            g.flag = nullptr;
            const csim_ctx::PeerLink& L = c->peer[q];
            if (ps[q].peer >= 0) {
                CSIM_REQUIRE(L.rank == ps[q].peer && L.tile[slot] != nullptr, CSIM_ERR_COMM,
                             "peer_exchange: neighbour not mapped");
                // my region lands in the neighbour's ghost area on ITS side (-dx,-dy): its own receive rule
                // with its own tile size (bands keep their along-side origin: the neighbour shares that
                // physical side with me)
                const int rx = dx > 0 ? -T : (dx < 0 ? L.nx : (pl ? -1 : 0));
                const int ry = dy > 0 ? -T : (dy < 0 ? L.ny : (pb ? -1 : 0));
                double* interior = L.tile[slot] + static_cast<long long>(kLeadY) * L.pitch + kLeadX;
                g.dst = interior + static_cast<long long>(ry) * L.pitch + rx;
                g.dst_pitch = L.pitch;
                g.flag = L.flags + (7 - q);  // the slot of direction (-dx,-dy) in the neighbour's array
                mask |= 1u << q;
                wire += static_cast<long long>(g.w) * g.h;
            }
            ++q;
        }
    if (bytes_sent) *bytes_sent = static_cast<size_t>(wire) * sizeof(double);
    const unsigned long long timeout_ns = peer_timeout_ns();
    k_push_light<<<c->sm_count, 32, 0, stream>>>(f->interior(), f->pitch, t, c->d_flags);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
    if (after_push) CSIM_CUDA(cudaEventRecord(after_push, stream));
    k_wait_light<<<1, 32, 0, stream>>>(c->d_flags, mask, c->d_flags, timeout_ns, c->d_err);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
    return CSIM_OK;
}

// Bootstrap record every rank contributes to the all-gather in peer_setup.
struct PeerInfo {
    cudaIpcMemHandle_t tile[2];
    cudaIpcMemHandle_t flags;
    unsigned long long raw_tile[2], raw_flags;  // same-process pointers
    long long pid;
    long long pitch;
    int nx, ny, device, pad;
};

int peer_teardown(csim_ctx* c) {
    if (!c->peer_ready && !c->d_flags) return CSIM_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream_x);
    for (auto& L : c->peer) {
        if (L.ipc) {
            for (double*& t : L.tile)
                if (t) cudaIpcCloseMemHandle(t);
            if (L.flags) cudaIpcCloseMemHandle(L.flags);
        }
        L = csim_ctx::PeerLink();
    }
    c->peer_ready = false;
    if (c->d_flags) cudaFree(c->d_flags);
    c->d_flags = nullptr;
    if (c->h_err) cudaFreeHost(c->h_err);
    c->h_err = nullptr;
    c->d_err = nullptr;
    return CSIM_OK;
}

// Collective over the communicator: every rank maps its neighbours' copies of (u, tmp) and their flag words.
// Returns CSIM_OK with c->peer_ready == false (and peer_failed set) when a peer cannot be mapped; the
// decision is agreed across ranks, so either all use the peer path or none does.
static int peer_setup(csim_field* u, csim_field* tmp, const csim_decomp* dec) {
    csim_ctx* c = u->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream_x));
    if (int rc = peer_teardown(c)) return rc;
    drop_graphs(c);  // captured block loops hold pointers into the old mappings
    const int size = c->comm_size, me = c->comm_rank;
    CSIM_CUDA(cudaMalloc(&c->d_flags, kCtlWords * sizeof(unsigned)));
    CSIM_CUDA(cudaMemset(c->d_flags, 0, kCtlWords * sizeof(unsigned)));
    CSIM_CUDA(cudaHostAlloc(&c->h_err, sizeof(unsigned), cudaHostAllocMapped));
    *c->h_err = 0;
    CSIM_CUDA(cudaHostGetDevicePointer(&c->d_err, c->h_err, 0));
    c->peer_tile[0] = u->base;
    c->peer_tile[1] = tmp->base;

    PeerInfo mine;
    std::memset(&mine, 0, sizeof mine);
    CSIM_CUDA(cudaIpcGetMemHandle(&mine.tile[0], u->base));
    CSIM_CUDA(cudaIpcGetMemHandle(&mine.tile[1], tmp->base));
    CSIM_CUDA(cudaIpcGetMemHandle(&mine.flags, c->d_flags));
    mine.raw_tile[0] = reinterpret_cast<unsigned long long>(u->base);
    mine.raw_tile[1] = reinterpret_cast<unsigned long long>(tmp->base);
    mine.raw_flags = reinterpret_cast<unsigned long long>(c->d_flags);
    mine.pid = static_cast<long long>(getpid());
    mine.pitch = u->pitch;
    mine.nx = u->nx;
    mine.ny = u->ny;
    mine.device = c->device;

    // all-gather of the records over the communicator that is already there
    static_assert(sizeof(PeerInfo) % 8 == 0, "PeerInfo must be a whole number of doubles");
    const size_t words = sizeof(PeerInfo) / 8;
    double* d_all = nullptr;
    CSIM_CUDA(cudaMalloc(&d_all, sizeof(PeerInfo) * static_cast<size_t>(size + 1)));
    CSIM_CUDA(cudaMemcpyAsync(d_all + words * size, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    CSIM_NCCL(g_nccl.AllGather(d_all + words * size, d_all, words, ncclDouble, static_cast<ncclComm_t>(c->comm),
                               c->stream));
    std::vector<PeerInfo> all(static_cast<size_t>(size));
    CSIM_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(PeerInfo) * static_cast<size_t>(size), cudaMemcpyDeviceToHost,
                              c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    CSIM_CUDA(cudaFree(d_all));

    const int cx = dec->coords[0], cy = dec->coords[1];
    double failed = 0.0;
    int q = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            csim_ctx::PeerLink& L = c->peer[q++];
            const int x = cx + dx, y = cy + dy;
            if (x < 0 || y < 0 || x >= dec->dims[0] || y >= dec->dims[1]) continue;
            const int r = x * dec->dims[1] + y;
            CSIM_REQUIRE(r != me && r < size, CSIM_ERR_INVALID, "peer_setup: bad neighbour rank");
            const PeerInfo& o = all[static_cast<size_t>(r)];
            L.rank = r;
            L.pitch = o.pitch;
            L.nx = o.nx;
            L.ny = o.ny;
            if (o.pid == mine.pid) {  // ranks are threads of one process: plain peer access
                if (o.device != c->device) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(o.device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled)
                        cudaGetLastError();
                    else if (e != cudaSuccess) {
                        cudaGetLastError();
                        failed = 1.0;
                        continue;
                    }
                }
                L.tile[0] = reinterpret_cast<double*>(o.raw_tile[0]);
                L.tile[1] = reinterpret_cast<double*>(o.raw_tile[1]);
                L.flags = reinterpret_cast<unsigned*>(o.raw_flags);
                L.ipc = false;
            } else {
                void* p[3] = {nullptr, nullptr, nullptr};
                const cudaIpcMemHandle_t* hs[3] = {&o.tile[0], &o.tile[1], &o.flags};
                bool ok = true;
                for (int k = 0; k < 3 && ok; ++k)
                    if (cudaIpcOpenMemHandle(&p[k], *hs[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                        cudaGetLastError();
                        ok = false;
                    }
                L.tile[0] = static_cast<double*>(p[0]);
                L.tile[1] = static_cast<double*>(p[1]);
                L.flags = static_cast<unsigned*>(p[2]);
                L.ipc = true;
                if (!ok) failed = 1.0;
            }
        }
    // agree: nobody pushes before everyone has mapped everyone (and zeroed its flags), and if any rank
    // could not map a neighbour all ranks stay on the NCCL path
    if (int rc = csim_comm_allreduce_max(c, &failed, 1)) return rc;
    if (failed != 0.0) {
        peer_teardown(c);
        c->peer_failed = true;
        return CSIM_OK;
    }
    c->peer_ready = true;
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
    std::vector<cudaEvent_t> ev;  // base, x0(0) x1(0), then per block f0 f1 i0 i1 [x0 x1]
    std::vector<cudaEvent_t> mid; // peer path: one per exchange, recorded between the push and the wait kernel
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
    int last_path = 0;       // 0 none yet, 1 peer stores, 2 NCCL
    RunProfile prof;
    csim_halo_stats last{};
};

static void drop_graphs(csim_ctx* c) {
    if (!c->run_state) return;
    RunState* rs = static_cast<RunState*>(c->run_state);
    for (RunGraph& g : rs->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    rs->graphs.clear();
}

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

// ---- the coupled block loop: ONE sweep launch per block, flags instead of stream events -------------
// Main stream   : sweep(0), sweep(1), … back to back.  In sweep(n) the frame items (the ones that read ghost
//                 lines) come first; each checks `halo landed >= n` before it requests a row (normally true
//                 long before the launch), and the last one to finish sets `frame done = n`.
// Exchange stream: wait(frame done >= n) → exchange(n+1) → set(halo landed = n+1), all while sweep(n)'s
//                 interior items run.  The wait and set kernels are single-warp CTAs that fit beside a full
//                 sweep; the exchange itself has a whole sweep of slack (≈ 1 ms at 16384^2).
// Against the split loop below (frame and interior as two launches on two streams, ordered by events) this
// removes the two cross-stream hops every block start paid for — interior(n) done → frame(n+1) may start →
// interior(n+1) may start — which cost 3–4.5 % at 2 and 4 GPUs (profiles/r02_multigpu.md), and it no longer
// matters when NCCL's kernel finds room on the SMs.
__global__ void __maxnreg__(32) k_flag_wait_light(const unsigned* flag, unsigned seq, unsigned long long timeout_ns,
                                                  unsigned* err) {
    if (threadIdx.x == 0) {
        const volatile unsigned* f = flag;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (static_cast<int>(*f - seq) < 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                *reinterpret_cast<volatile unsigned*>(err) = 202u;
                __threadfence_system();
                __trap();
            }
            __nanosleep(200);
        }
        __threadfence();
    }
}
__global__ void __maxnreg__(32) k_flag_set_light(unsigned* flag, unsigned seq) {
    if (threadIdx.x == 0) {
        __threadfence();
        *reinterpret_cast<volatile unsigned*>(flag) = seq;
    }
}

static int ensure_coupling(csim_ctx* c) {
    if (c->d_couple) return CSIM_OK;
    CSIM_CUDA(cudaMalloc(&c->d_couple, 8 * sizeof(unsigned)));
    CSIM_CUDA(cudaMemset(c->d_couple, 0, 8 * sizeof(unsigned)));
    CSIM_CUDA(cudaHostAlloc(&c->h_couple_err, sizeof(unsigned), cudaHostAllocMapped));
    *c->h_couple_err = 0;
    CSIM_CUDA(cudaHostGetDevicePointer(&c->d_couple_err, c->h_couple_err, 0));
    c->couple_seq = 0;
    // Load the two flag kernels NOW.  With lazy module loading the first launch of a function can wait for
    // the device to go idle; a sweep whose frame items spin until k_flag_wait_light → exchange → k_flag_set_light
    // have run would then wait for a kernel that cannot be loaded while it spins (observed: both ranks
    // trapped on the 60 s bound the first time the wait kernel was launched behind a dependent sweep).
    cudaFuncAttributes fa;
    CSIM_CUDA(cudaFuncGetAttributes(&fa, k_flag_wait_light));
    CSIM_CUDA(cudaFuncGetAttributes(&fa, k_flag_set_light));
    k_flag_set_light<<<1, 32, 0, c->stream_x>>>(c->d_couple + 1, 0u);
    k_flag_wait_light<<<1, 32, 0, c->stream_x>>>(c->d_couple + 1, 0u, 1000000000ull, c->d_couple_err);
    CSIM_CUDA(cudaStreamSynchronize(c->stream_x));
    return CSIM_OK;
}

static int enqueue_blocks_coupled(csim_field* u, csim_field* tmp, const csim_step_params* p, const csim_decomp* dec,
                                  const StepK& k, int mode, int maxT, int nsteps, bool zero_terms, int values_after,
                                  RunProfile* prof, int* swaps) {
    csim_ctx* c = u->ctx;
    if (int rc = ensure_coupling(c)) return rc;
    const bool peer = peer_tiles_match(c, u, tmp);
    auto exchange = [&](csim_field* f, int lines, size_t* wire_out) {
        return peer ? peer_exchange(f, dec, lines, c->stream_x, wire_out)
                    : wide_exchange(f, dec, lines, c->stream_x, wire_out);
    };
    auto stamp = [&](cudaStream_t st) -> int {
        if (!prof) return CSIM_OK;
        CSIM_CUDA(cudaEventRecord(prof->next(c), st));
        return CSIM_OK;
    };
    TbCoupling cp;
    cp.halo_flag = c->d_couple + 0;
    cp.done_flag = c->d_couple + 1;
    cp.ticket = c->d_couple + 2;
    cp.err = c->d_couple_err;
    cp.timeout_ns = peer_timeout_ns();
    int left = nsteps;
    int T = left < maxT ? left : maxT;
    *swaps = 0;
    CSIM_CUDA(cudaEventRecord(c->ev_fork, c->stream));  // everything queued on the tiles so far
    CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));
    if (peer)
        if (int rc = peer_ready_barrier(c, c->stream_x)) return rc;
    if (int rc = stamp(c->stream_x)) return rc;  // x0(0)
    size_t wire = 0;
    if (int rc = exchange(u, T, &wire)) return rc;  // exchange(0): the only one not hidden
    if (prof) prof->bytes_per_exchange = wire;
    unsigned seq = ++c->couple_seq;
    k_flag_set_light<<<1, 32, 0, c->stream_x>>>(cp.halo_flag, seq);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
    if (int rc = stamp(c->stream_x)) return rc;  // x1(0)
    // Host order: sweep(n+1) is queued on the main stream BEFORE the exchange stream gets the work of
    // exchange(n+1).  The two only meet through the flags, so the order is free on the host — and NCCL's
    // enqueue can block the host until the device has taken earlier operations of the communicator, which
    // here wait for sweep(n)'s frame items: queued the other way round, every sweep launch arrived late
    // (measured: host enqueue time = device time, 65 us of idle GPU per block at 16384^2).
    bool launched = false;
    cp.seq = seq;
    if (int rc = stamp(c->stream)) return rc;  // s0(0)
    if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_COUPLED, c->stream, &launched, zero_terms, &cp)) return rc;
    if (int rc = stamp(c->stream)) return rc;  // s1(0)
    tmp->values = values_after;
    csim_field_swap(u, tmp);  // u: the state after block 0 (being computed)
    ++*swaps;
    if (prof) ++prof->blocks;
    left -= T;
    while (left > 0) {
        const unsigned seq_done = seq;  // frame items of the sweep in flight publish this
        T = left < maxT ? left : maxT;
        seq = ++c->couple_seq;
        cp.seq = seq;
        if (int rc = stamp(c->stream)) return rc;  // s0(n+1)
        if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_COUPLED, c->stream, &launched, zero_terms, &cp)) return rc;
        if (int rc = stamp(c->stream)) return rc;  // s1(n+1)
        // exchange(n+1) on the state sweep(n) is writing (u): its bands are final once sweep(n)'s frame items are done
        k_flag_wait_light<<<1, 32, 0, c->stream_x>>>(cp.done_flag, seq_done, cp.timeout_ns, cp.err);
        ++c->launches;
        CSIM_CUDA(cudaGetLastError());
        if (int rc = stamp(c->stream_x)) return rc;  // x0(n+1)
        if (int rc = exchange(u, T, nullptr)) return rc;
        k_flag_set_light<<<1, 32, 0, c->stream_x>>>(cp.halo_flag, seq);
        ++c->launches;
        CSIM_CUDA(cudaGetLastError());
        if (int rc = stamp(c->stream_x)) return rc;  // x1(n+1)
        tmp->values = values_after;
        csim_field_swap(u, tmp);
        ++*swaps;
        if (prof) ++prof->blocks;
        left -= T;
    }
    // whatever follows on the main stream is ordered after the exchange stream as well
    CSIM_CUDA(cudaEventRecord(c->ev_join, c->stream_x));
    CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    return CSIM_OK;
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
    const bool peer = peer_tiles_match(c, u, tmp);
    auto exchange = [&](csim_field* f, int lines, size_t* wire_out) {
        if (!peer) return wide_exchange(f, dec, lines, c->stream_x, wire_out);
        cudaEvent_t mid = nullptr;
        if (prof) {  // push | wait split of the exchange, kept apart from the main timeline
            cudaSetDevice(c->device);
            cudaEventCreate(&mid);
            prof->mid.push_back(mid);
        }
        return peer_exchange(f, dec, lines, c->stream_x, wire_out, mid);
    };
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
    if (peer)
        if (int rc = peer_ready_barrier(c, c->stream_x)) return rc;  // the neighbours' tiles may be written now
    if (int rc = stamp(c->stream_x)) return rc;  // x0(0)
    size_t wire = 0;
    if (int rc = exchange(u, T, &wire)) return rc;  // exchange(0): the only one not hidden
    if (int rc = stamp(c->stream_x)) return rc;  // x1(0)
    if (prof) prof->bytes_per_exchange = wire;
    bool first = true;
    while (left > 0) {
        bool launched = false;
        if (!first) CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));  // interior(n-1) done
        if (int rc = stamp(c->stream_x)) return rc;                              // f0(n)
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
            if (int rc = exchange(u, T, nullptr)) return rc;  // exchange(n+1): reads frame(n)'s cells
            if (int rc = stamp(c->stream_x)) return rc;                          // x1(n+1)
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
    // event layout: base, x0(0), x1(0), then per block n: f0(n) f1(n) i0(n) i1(n) [x0(n+1) x1(n+1) unless last]
    std::vector<double> t(pr.used, 0.0);
    for (size_t i = 1; i < pr.used; ++i) {
        float ms = 0.f;
        CSIM_CUDA(cudaEventElapsedTime(&ms, pr.ev[0], pr.ev[i]));
        t[i] = ms;
    }
    double ex = 0.0, hidden = 0.0, frame = 0.0, interior = 0.0, gap = 0.0;
    int n_hidden = 0;
    size_t i = 1;  // t[0] is the time base
    double x0 = 0.0, x1 = 0.0;
    if (pr.used >= 3) {
        x0 = t[i];
        x1 = t[i + 1];
        i += 2;
    }
    double prev_i0 = 0.0, prev_i1 = 0.0;
    for (int n = 0; n < pr.blocks && i + 3 < pr.used; ++n) {
        const double f0 = t[i], f1 = t[i + 1], i0 = t[i + 2], i1 = t[i + 3];
        i += 4;
        const double dur = x1 - x0;
        if (n == 0) {
            st.first_exchange_us = 1e3 * dur;
        } else {  // exchange(n) ran beside interior(n-1): how much of it lies inside that interval
            ex += dur;
            const double lo = x0 > prev_i0 ? x0 : prev_i0, hi = x1 < prev_i1 ? x1 : prev_i1;
            if (hi > lo) hidden += hi - lo;
            gap += f0 - x1;  // the frame sweep waits for the interior sweep of the previous block
            ++n_hidden;
        }
        frame += f1 - f0;
        interior += i1 - i0;
        prev_i0 = i0;
        prev_i1 = i1;
        if (n + 1 < pr.blocks && i + 1 < pr.used) {
            x0 = t[i];
            x1 = t[i + 1];
            i += 2;
        }
    }
    st.wait_for_interior_us = n_hidden ? 1e3 * gap / n_hidden : 0.0;
    // peer path: how much of an exchange is this rank's own store kernel (the rest is waiting for the neighbours')
    st.push_us = 0.0;
    if (pr.mid.size() > 1) {
        // exchange k (k >= 1) starts at the x0 recorded right before it: events 1 (k = 0), then 7 + 6 (k - 1)
        double sum = 0.0;
        int cnt = 0;
        for (size_t k = 1; k < pr.mid.size(); ++k) {
            const size_t ix0 = 7 + 6 * (k - 1);
            if (ix0 >= pr.used) break;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.ev[ix0], pr.mid[k]) == cudaSuccess) {
                sum += ms;
                ++cnt;
            }
        }
        st.push_us = cnt ? 1e3 * sum / cnt : 0.0;
    }
    for (cudaEvent_t e : pr.mid) cudaEventDestroy(e);
    pr.mid.clear();
    st.exchange_us = n_hidden ? 1e3 * ex / n_hidden : st.first_exchange_us;
    st.overlap_fraction = ex > 0.0 ? hidden / ex : 0.0;
    st.frame_us = pr.blocks ? 1e3 * frame / pr.blocks : 0.0;
    st.interior_us = pr.blocks ? 1e3 * interior / pr.blocks : 0.0;
    st.total_ms = pr.used ? t[pr.used - 1] : 0.0;
    rs->last = st;
    pr.on = false;
    return CSIM_OK;
}

// Timeline of a profiled coupled call: base, x0(0), x1(0), then per block s0(n) s1(n) [x0(n+1) x1(n+1)].
static int finish_profile_coupled(csim_ctx* c, RunState* rs) {
    RunProfile& pr = rs->prof;
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream_x));
    csim_halo_stats st{};
    st.blocks = pr.blocks;
    st.bytes_per_exchange = pr.bytes_per_exchange;
    std::vector<double> t(pr.used, 0.0);
    for (size_t i = 1; i < pr.used; ++i) {
        float ms = 0.f;
        CSIM_CUDA(cudaEventElapsedTime(&ms, pr.ev[0], pr.ev[i]));
        t[i] = ms;
    }
    double ex = 0.0, hidden = 0.0, sweep = 0.0;
    int n_ex = 0;
    if (pr.used >= 3) st.first_exchange_us = 1e3 * (t[2] - t[1]);
    // events: base, x0(0), x1(0), s0(0), s1(0), then for every further block m: s0(m) s1(m) x0(m) x1(m);
    // exchange(m) runs beside sweep(m-1)
    size_t i = 3;
    double prev_s0 = 0.0, prev_s1 = 0.0;
    for (int m = 0; m < pr.blocks && i + 1 < pr.used; ++m) {
        const double s0 = t[i], s1 = t[i + 1];
        i += 2;
        sweep += s1 - s0;
        if (m > 0 && i + 1 < pr.used) {
            const double x0 = t[i], x1 = t[i + 1];
            i += 2;
            ex += x1 - x0;
            const double lo = x0 > prev_s0 ? x0 : prev_s0, hi = x1 < prev_s1 ? x1 : prev_s1;
            if (hi > lo) hidden += hi - lo;
            ++n_ex;
        }
        prev_s0 = s0;
        prev_s1 = s1;
    }
    st.exchange_us = n_ex ? 1e3 * ex / n_ex : st.first_exchange_us;
    st.overlap_fraction = ex > 0.0 ? hidden / ex : 0.0;
    st.frame_us = 0.0;  // frame and interior items share one launch
    st.interior_us = pr.blocks ? 1e3 * sweep / pr.blocks : 0.0;
    st.total_ms = pr.used ? t[pr.used - 1] : 0.0;
    st.push_us = 0.0;
    st.wait_for_interior_us = 0.0;
    for (cudaEvent_t e : pr.mid) cudaEventDestroy(e);
    pr.mid.clear();
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
    // halo path: peer stores unless refused; mapped on first use and again when the tiles change
    if (peer_path_wanted() && !c->peer_failed && !peer_tiles_match(c, u, tmp))
        if (int rc = peer_setup(u, tmp, dec)) return rc;
    if (c->h_err && *c->h_err)
        return fail(CSIM_ERR_TIMEOUT, "csim_run_steps: a neighbour's halo did not arrive within the bounded wait");
    const bool peer = peer_tiles_match(c, u, tmp);
    if (!peer)
        if (int rc = ensure_wide(c, u, dec, maxT)) return rc;
    RunState* rs = run_state(c);
    rs->last_path = peer ? 1 : 2;
    // CSIM_LOOP=coupled: one sweep launch per block, coupled to the exchange stream by flags; default
    // (CSIM_LOOP=split): frame and interior as two launches on two streams, ordered by events
    static const bool coupled = [] {
        const char* e = std::getenv("CSIM_LOOP");
        return e && std::strcmp(e, "coupled") == 0;
    }();
    if (c->h_couple_err && *c->h_couple_err)
        return fail(CSIM_ERR_TIMEOUT, "csim_run_steps: a halo or frame flag did not arrive within the bounded wait");
    if (coupled) {
        int swaps_c = 0;
        RunProfile* pf = rs->prof.on ? &rs->prof : nullptr;
        if (pf) {
            pf->used = 0;
            pf->blocks = 0;
            CSIM_CUDA(cudaEventRecord(pf->next(c), c->stream));  // time base
        }
        if (int rc = enqueue_blocks_coupled(u, tmp, p, dec, k, mode, maxT, nsteps, zero_terms, values_after, pf, &swaps_c))
            return rc;
        rs->comm_warm = true;
        return pf ? finish_profile_coupled(c, rs) : CSIM_OK;
    }
    int swaps = 0;
    if (rs->prof.on) {  // profiled call: eager, with timestamps (csim_halo_profile)
        rs->prof.used = 0;
        rs->prof.blocks = 0;
        for (cudaEvent_t e : rs->prof.mid) cudaEventDestroy(e);
        rs->prof.mid.clear();
        CSIM_CUDA(cudaEventRecord(rs->prof.next(c), c->stream));  // time base
        if (int rc = enqueue_blocks(u, tmp, p, dec, k, mode, maxT, nsteps, zero_terms, values_after, &rs->prof, &swaps))
            return rc;
        rs->comm_warm = true;
        return finish_profile(c, rs);
    }
    // CUDA-graph replay of the block loop (CSIM_GRAPH=1; off by default).  With the NCCL halo path the loop
    // is ~8 launches per block and ncclGroupEnd alone costs ~200 us of host time, so at 8192^2 per GPU the
    // host spent as long enqueueing a window as the GPUs spent running it (round 1: 7.4 ms against 7.9 ms).
    // Capturing the loop takes the host out (0.03 ms per window) — but the replay measured SLOWER on the
    // device (2 x B200, 8192^2 per GPU: 8.46 ms per window against 7.42 ms eager; 16384^2: 26.9 against
    // 26.6, profiles/r02_multigpu.md): in the graph the frame sweep and the exchange lose the stream
    // priority that lets them overtake the interior sweep.  The peer halo path needs four small launches per
    // block and no NCCL call, which removes the host cost without a graph; the replay stays as an option.
    // The first call of a communicator always runs eagerly (NCCL sets up its connections on first use,
    // which must not happen inside a capture).
    static const bool graphs_on = [] {
        const char* e = std::getenv("CSIM_GRAPH");
        return e && std::strcmp(e, "1") == 0;
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
    key.zero_terms = (zero_terms ? 1 : 0) | (peer ? 2 : 0);
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

const char* csim_halo_path(const csim_ctx* c) {
    if (!c || !c->run_state) return "none";
    const int p = static_cast<const RunState*>(c->run_state)->last_path;
    return p == 1 ? "peer" : (p == 2 ? "nccl" : "none");
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
