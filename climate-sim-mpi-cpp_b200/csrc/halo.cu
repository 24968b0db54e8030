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

// Fill T ghost lines of `f` on every side that has a neighbour, on `stream`.
static int wide_exchange(csim_field* f, const csim_decomp* dec, int T, cudaStream_t stream) {
    csim_ctx* c = f->ctx;
    csim_decomp d = *dec;  // the plan is a function of the decomposition; the tile fixes the local size
    d.nx_local = f->nx;
    d.ny_local = f->ny;
    csim_xregion ps[8], pr8[8];
    if (int rc = csim_wide_exchange_plan(&d, T, ps, pr8)) return rc;
    XTable snd, rcv;
    long long off = 0;
    for (int q = 0; q < 8; ++q) {
        snd.r[q] = XRegion{ps[q].x0, ps[q].y0, ps[q].w, ps[q].h, off, ps[q].peer};
        rcv.r[q] = XRegion{pr8[q].x0, pr8[q].y0, pr8[q].w, pr8[q].h, 0, pr8[q].peer};
        off += static_cast<long long>(ps[q].w) * ps[q].h;
    }
    const long long send_total = off;
    for (int q = 0; q < 8; ++q) rcv.r[q].off = send_total + snd.r[q].off;
    if (c->wide_doubles < static_cast<size_t>(2 * send_total)) {
        if (c->d_wide) {
            CSIM_CUDA(cudaDeviceSynchronize());
            CSIM_CUDA(cudaFree(c->d_wide));
            c->d_wide = nullptr;
        }
        CSIM_CUDA(cudaMalloc(&c->d_wide, static_cast<size_t>(2 * send_total) * sizeof(double)));
        c->wide_doubles = static_cast<size_t>(2 * send_total);
    }
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

// ---- peer-memory exchange: store the bands straight into the neighbours' ghost lines ---------------
struct PushRegion {
    int x0, y0, w, h;     // source region in this rank's tile (interior coordinates)
    double* dst;          // neighbour's cell that receives the region's first cell (mapped pointer)
    long long dst_pitch;  // neighbour's row pitch
    unsigned* flag;       // neighbour's flag word for this direction, nullptr: no neighbour
};
struct PushTable {
    PushRegion r[8];
};

// Copy all regions, then (last CTA only, after a system-scope fence) publish `seq` in every
// neighbour's flag word.  Few small CTAs on purpose: the kernel has to find SM slots while the
// interior sweep fills the machine.
__global__ void __launch_bounds__(128) k_push_regions(const double* __restrict__ u, long long pitch, PushTable t,
                                                      unsigned seq, unsigned* __restrict__ ticket) {
    for (int q = 0; q < 8; ++q) {
        const PushRegion g = t.r[q];
        if (!g.flag) continue;
        const int n = g.w * g.h;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
            const int yy = e / g.w, xx = e - yy * g.w;
            g.dst[static_cast<long long>(yy) * g.dst_pitch + xx] =
                u[static_cast<long long>(g.y0 + yy) * pitch + g.x0 + xx];
        }
    }
    __threadfence_system();  // this thread's peer stores are visible system-wide before the ticket
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < 8) {
        if (threadIdx.x == 0) *ticket = 0;  // re-arm for the next push (stream order protects it)
        unsigned* f = t.r[threadIdx.x].flag;
        if (f) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned*>(f) = seq;
        }
    }
}

// Gate of the frame sweep: wait until every neighbour in `mask` has published >= seq.  Bounded: after
// `timeout_ns` the kernel records the failure and returns, so a lost neighbour ends in an error code.
__global__ void k_wait_flags(const unsigned* flags, unsigned mask, unsigned seq, unsigned long long timeout_ns,
                             unsigned* err) {
    const int k = threadIdx.x;
    if (k >= 8 || !((mask >> k) & 1)) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    const volatile unsigned* f = flags + k;
    // seq wraps after 2^32 exchanges; compare as a signed distance
    while (static_cast<int>(*f - seq) < 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            atomicExch(err, 1u + static_cast<unsigned>(k));
            return;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

static bool peer_path_enabled(const csim_ctx* c, const csim_field* u, const csim_field* tmp) {
    if (!c->peer_ready) return false;
    static const bool force_nccl = [] {
        const char* e = std::getenv("CSIM_HALO");
        return e && std::strcmp(e, "nccl") == 0;
    }();
    if (force_nccl) return false;
    return (u->base == c->peer_tile[0] && tmp->base == c->peer_tile[1]) ||
           (u->base == c->peer_tile[1] && tmp->base == c->peer_tile[0]);
}

// Peer version of wide_exchange: push this rank's bands of tile `f` into the neighbours' copy of
// the same tile, then gate `stream` on the neighbours' pushes into ours.
static int peer_exchange(csim_field* f, const csim_decomp* dec, int T, cudaStream_t stream) {
    csim_ctx* c = f->ctx;
    csim_decomp d = *dec;
    d.nx_local = f->nx;
    d.ny_local = f->ny;
    csim_xregion ps[8], pr8[8];
    if (int rc = csim_wide_exchange_plan(&d, T, ps, pr8)) return rc;
    const int slot = f->base == c->peer_tile[0] ? 0 : 1;
    const bool pl = dec->nbr[CSIM_LEFT] == CSIM_PROC_NULL, pb = dec->nbr[CSIM_BOTTOM] == CSIM_PROC_NULL;
    PushTable t;
    unsigned mask = 0;
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
                             "peer_exchange: neighbour not mapped (csim_peer_setup with other tiles?)");
                // my region lands in the neighbour's ghost area on ITS side (-dx,-dy): its own
                // receive rule with its own tile size (bands keep their along-side origin: the
                // neighbour shares that physical side with me)
                const int rx = dx > 0 ? -T : (dx < 0 ? L.nx : (pl ? -1 : 0));
                const int ry = dy > 0 ? -T : (dy < 0 ? L.ny : (pb ? -1 : 0));
                double* interior = L.tile[slot] + static_cast<long long>(kLeadY) * L.pitch + kLeadX;
                g.dst = interior + static_cast<long long>(ry) * L.pitch + rx;
                g.dst_pitch = L.pitch;
                g.flag = L.flags + (7 - q);  // the slot of direction (-dx,-dy) in the neighbour's array
                mask |= 1u << q;
            }
            ++q;
        }
    const unsigned seq = ++c->push_seq;
    k_push_regions<<<8, 128, 0, stream>>>(f->interior(), f->pitch, t, seq, c->d_flags + 8);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
    k_wait_flags<<<1, 32, 0, stream>>>(c->d_flags, mask, seq, 5000000000ull, c->d_err);
    ++c->launches;
    CSIM_CUDA(cudaGetLastError());
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

// Bootstrap record every rank contributes to the all-gather in csim_peer_setup.
struct PeerInfo {
    cudaIpcMemHandle_t tile[2];
    cudaIpcMemHandle_t flags;
    unsigned long long raw_tile[2], raw_flags;  // same-process pointers
    long long pid;
    long long pitch;
    int nx, ny, device, pad;
};

int csim_peer_teardown(csim_ctx* c) {
    CSIM_REQUIRE(c != nullptr, CSIM_ERR_INVALID, "csim_peer_teardown: ctx is null");
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

int csim_peer_setup(csim_field* u, csim_field* tmp, const csim_decomp* dec) {
    CSIM_REQUIRE(u != nullptr && tmp != nullptr && dec != nullptr, CSIM_ERR_INVALID, "csim_peer_setup: null argument");
    csim_ctx* c = u->ctx;
    CSIM_REQUIRE(c == tmp->ctx && u->pitch == tmp->pitch && u->nx == tmp->nx && u->ny == tmp->ny, CSIM_ERR_INVALID,
                 "csim_peer_setup: tiles differ in geometry");
    CSIM_REQUIRE(c->comm != nullptr, CSIM_ERR_COMM, "csim_peer_setup: needs csim_comm_init first");
    CSIM_CUDA(cudaSetDevice(c->device));
    if (int rc = csim_peer_teardown(c)) return rc;
    const int size = c->comm_size, me = c->comm_rank;
    CSIM_CUDA(cudaMalloc(&c->d_flags, 16 * sizeof(unsigned)));
    CSIM_CUDA(cudaMemset(c->d_flags, 0, 16 * sizeof(unsigned)));
    CSIM_CUDA(cudaHostAlloc(&c->h_err, sizeof(unsigned), cudaHostAllocMapped));
    *c->h_err = 0;
    CSIM_CUDA(cudaHostGetDevicePointer(&c->d_err, c->h_err, 0));
    c->push_seq = 0;
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
    int q = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            csim_ctx::PeerLink& L = c->peer[q++];
            const int x = cx + dx, y = cy + dy;
            if (x < 0 || y < 0 || x >= dec->dims[0] || y >= dec->dims[1]) continue;
            const int r = x * dec->dims[1] + y;
            CSIM_REQUIRE(r != me && r < size, CSIM_ERR_INVALID, "csim_peer_setup: bad neighbour rank");
            const PeerInfo& o = all[static_cast<size_t>(r)];
            L.rank = r;
            L.pitch = o.pitch;
            L.nx = o.nx;
            L.ny = o.ny;
            if (o.pid == mine.pid) {  // ranks are threads of one process: plain peer access
                if (o.device != c->device) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(o.device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled)
                        cudaGetLastError();
                    else if (e != cudaSuccess)
                        return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                }
                L.tile[0] = reinterpret_cast<double*>(o.raw_tile[0]);
                L.tile[1] = reinterpret_cast<double*>(o.raw_tile[1]);
                L.flags = reinterpret_cast<unsigned*>(o.raw_flags);
                L.ipc = false;
            } else {
                void* p0 = nullptr;
                void* p1 = nullptr;
                void* pf = nullptr;
                CSIM_CUDA(cudaIpcOpenMemHandle(&p0, o.tile[0], cudaIpcMemLazyEnablePeerAccess));
                CSIM_CUDA(cudaIpcOpenMemHandle(&p1, o.tile[1], cudaIpcMemLazyEnablePeerAccess));
                CSIM_CUDA(cudaIpcOpenMemHandle(&pf, o.flags, cudaIpcMemLazyEnablePeerAccess));
                L.tile[0] = static_cast<double*>(p0);
                L.tile[1] = static_cast<double*>(p1);
                L.flags = static_cast<unsigned*>(pf);
                L.ipc = true;
            }
        }
    // nobody pushes before everyone has mapped everyone (and zeroed its flags)
    if (int rc = csim_comm_allreduce_max(c, nullptr, 0)) return rc;
    c->peer_ready = true;
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
    int maxT = (mode == MODE_DIV || (p->flags & CSIM_STEP_NO_TEMPORAL)) ? 1 : tb_max_T();
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
    // one decision for the whole call, taken per rank (see resolve_zero_terms)
    bool zero_terms = false;
    if (nsteps >= maxT)
        if (int rc = resolve_zero_terms(u, p, k, mode, maxT, &zero_terms)) return rc;
    const int values_after = zero_terms ? csim_field::kClean
                                        : (u->values == csim_field::kTainted ? csim_field::kTainted : csim_field::kUnknown);
    const bool p2p = peer_path_enabled(c, u, tmp);
    if (p2p && *c->h_err)
        return fail(CSIM_ERR_TIMEOUT, "csim_run_steps: a neighbour's halo did not arrive within the bounded wait");
    auto exchange = [&](csim_field* f, int lines) {
        return p2p ? peer_exchange(f, dec, lines, c->stream_x) : wide_exchange(f, dec, lines, c->stream_x);
    };
    int left = nsteps;
    int T = left < maxT ? left : maxT;
    CSIM_CUDA(cudaEventRecord(c->ev_fork, c->stream));  // everything queued so far = "interior(-1)"
    CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));
    if (int rc = exchange(u, T)) return rc;  // exchange(0): the only one not hidden
    bool first = true;
    while (left > 0) {
        bool launched = false;
        if (!first) CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));  // interior(n-1) done
        CSIM_CUDA(cudaEventRecord(c->ev_go, c->stream_x));                       // go(n)
        if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_FRAME, c->stream_x, &launched, zero_terms)) return rc;
        CSIM_CUDA(cudaEventRecord(c->ev_join, c->stream_x));                     // frame(n) done
        CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_go, 0));
        if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_INTERIOR, c->stream, &launched, zero_terms)) return rc;
        CSIM_CUDA(cudaEventRecord(c->ev_fork, c->stream));                       // interior(n) done
        tmp->values = values_after;
        csim_field_swap(u, tmp);
        left -= T;
        first = false;
        if (left > 0) {
            T = left < maxT ? left : maxT;
            if (int rc = exchange(u, T)) return rc;  // exchange(n+1): reads frame(n)'s cells
        }
    }
    CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));  // the main stream orders everything again
    return CSIM_OK;
}

}  // extern "C"
