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
    int left = nsteps;
    while (left > 0) {
        const int T = left < maxT ? left : maxT;
        // main stream: the work items that read no ghost line.  Exchange stream: pack → NCCL →
        // unpack, then the frame items (edge strips, first/last chunk), which do read them.
        CSIM_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        CSIM_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_fork, 0));
        if (int rc = wide_exchange(u, dec, T, c->stream_x)) return rc;
        bool launched = false;
        if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_INTERIOR, c->stream, &launched)) return rc;
        if (int rc = launch_step_tb(u, tmp, p, k, mode, T, TB_FRAME, c->stream_x, &launched)) return rc;
        CSIM_CUDA(cudaEventRecord(c->ev_join, c->stream_x));
        CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        csim_field_swap(u, tmp);
        left -= T;
    }
    return CSIM_OK;
}

}  // extern "C"
