// context.cu — contexts, device tiles and host<->device movement of libcsim_b200.so.
//
// Stands behind the reference's Field (include/field.hpp:5-21, src/field.cpp:6-31): a Field
// becomes a halo-padded, pitch-aligned device allocation (csim_field); the host keeps the
// reference layout and these calls convert between the two with pitched 2-D copies.
#include <cmath>
#include <cstring>

#include <algorithm>

#include "csim_internal.hpp"

namespace csim {

static thread_local std::string t_last_error;

void set_error(const std::string& msg) { t_last_error = msg; }
int fail(int code, const std::string& msg) {
    t_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", static_cast<int>(e),
                  cudaGetErrorString(e), file, line, what);
    t_last_error = buf;
    return e == cudaErrorMemoryAllocation ? CSIM_ERR_NOMEM : CSIM_ERR_CUDA;
}
bool is_pow2(double x) {
    if (!(x > 0.0) || !std::isfinite(x)) return false;
    int e = 0;
    const double m = std::frexp(x, &e);
    // keep both x and 1/x normal so the reciprocal is exact
    return m == 0.5 && e > -1000 && e < 1000;
}

__global__ void k_fill(double* __restrict__ p, int64_t n, double v) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride)
        p[k] = v;
}

// Row-wise copy between two device arrays of different pitch (doubles): the device half of the
// staged host transfers below (dense host layout <-> pitched tile).
__global__ void __launch_bounds__(256) k_repitch(const double* __restrict__ src, int64_t spitch,
                                                 double* __restrict__ dst, int64_t dpitch, int width, int height) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = blockIdx.y; y < height; y += gridDim.y)
        dst[static_cast<int64_t>(y) * dpitch + x] = src[static_cast<int64_t>(y) * spitch + x];
}

// Interior of the pitched tile → dense ny x nx array of big-endian doubles (NetCDF wire order).
__global__ void __launch_bounds__(256) k_pack_interior_be(const double* __restrict__ in, int nx, int ny,
                                                          int64_t pitch, unsigned long long* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx) return;
    for (int y = blockIdx.y; y < ny; y += gridDim.y) {
        const unsigned long long v =
            static_cast<unsigned long long>(__double_as_longlong(in[static_cast<int64_t>(y) * pitch + x]));
        const unsigned lo = static_cast<unsigned>(v), hi = static_cast<unsigned>(v >> 32);
        out[static_cast<int64_t>(y) * nx + x] =
            (static_cast<unsigned long long>(__byte_perm(lo, 0, 0x0123)) << 32) | __byte_perm(hi, 0, 0x0123);
    }
}

}  // namespace csim

using namespace csim;

extern "C" {

const char* csim_last_error(void) { return t_last_error.c_str(); }
int csim_abi_version(void) { return CSIM_ABI_VERSION; }

int csim_ctx_create(int device, csim_ctx** out) {
    CSIM_REQUIRE(out != nullptr, CSIM_ERR_INVALID, "csim_ctx_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CSIM_ERR_CUDA,
                    std::string("csim_ctx_create: no CUDA device (this library has no CPU path): ") +
                        cudaGetErrorString(e));
    CSIM_REQUIRE(device >= 0 && device < ndev, CSIM_ERR_INVALID, "csim_ctx_create: bad device index");
    CSIM_CUDA(cudaSetDevice(device));
    csim_ctx* c = new (std::nothrow) csim_ctx();
    CSIM_REQUIRE(c != nullptr, CSIM_ERR_NOMEM, "csim_ctx_create: host allocation failed");
    c->device = device;
    cudaDeviceProp prop;
    CSIM_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CSIM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
        // the exchange stream outranks the main stream so that its small kernels (pack, NCCL, unpack,
        // frame sweep) are not queued behind the interior sweep that fills every SM
        int lo = 0, hi = 0;
        CSIM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CSIM_CUDA(cudaStreamCreateWithPriority(&c->stream_x, cudaStreamNonBlocking, hi));
    }
    CSIM_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CSIM_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CSIM_CUDA(cudaEventCreateWithFlags(&c->ev_go, cudaEventDisableTiming));
    CSIM_CUDA(cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CSIM_CUDA(cudaEventCreateWithFlags(&c->ev_snap_packed[b], cudaEventDisableTiming));
        CSIM_CUDA(cudaEventCreateWithFlags(&c->ev_snap_done[b], cudaEventDisableTiming));
    }
    c->scratch_doubles = 4096;
    CSIM_CUDA(cudaMalloc(&c->d_scratch, c->scratch_doubles * sizeof(double)));
    CSIM_CUDA(cudaMallocHost(&c->h_scratch, c->scratch_doubles * sizeof(double)));
    *out = c;
    return CSIM_OK;
}

int csim_device_count(int* count) {
    CSIM_REQUIRE(count != nullptr, CSIM_ERR_INVALID, "csim_device_count: null argument");
    *count = 0;
    const cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(CSIM_ERR_CUDA, std::string("csim_device_count: ") + cudaGetErrorString(e));
    }
    return CSIM_OK;
}

int csim_ctx_destroy(csim_ctx* c) {
    if (!c) return CSIM_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    peer_teardown(c);
    run_state_destroy(c);  // graphs hold NCCL kernels: they go before the communicator
    if (c->comm) csim_comm_destroy(c);
    // tiles that outlive their context (garbage-collection order in a host language) are orphaned:
    // their device memory goes now, csim_field_destroy later only frees the handle
    cudaDeviceSynchronize();
    for (csim_field* f : c->fields) {
        if (f->base) cudaFree(f->base);
        f->base = nullptr;
        f->ctx = nullptr;
    }
    c->fields.clear();
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    if (c->stream_x) {
        cudaStreamSynchronize(c->stream_x);
        cudaStreamDestroy(c->stream_x);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_go) cudaEventDestroy(c->ev_go);
    if (c->stream_copy) {
        cudaStreamSynchronize(c->stream_copy);
        cudaStreamDestroy(c->stream_copy);
    }
    for (int b = 0; b < 2; ++b) {
        if (c->ev_snap_packed[b]) cudaEventDestroy(c->ev_snap_packed[b]);
        if (c->ev_snap_done[b]) cudaEventDestroy(c->ev_snap_done[b]);
        if (c->d_snapbuf[b]) cudaFree(c->d_snapbuf[b]);
    }
    if (c->d_couple) cudaFree(c->d_couple);
    if (c->h_couple_err) cudaFreeHost(c->h_couple_err);
    if (c->d_pack) cudaFree(c->d_pack);
    if (c->d_wide) cudaFree(c->d_wide);
    if (c->d_snap) cudaFree(c->d_snap);
    if (c->d_stage) cudaFree(c->d_stage);
    delete c;
    return CSIM_OK;
}

int csim_sync(csim_ctx* c) {
    CSIM_REQUIRE(c != nullptr, CSIM_ERR_INVALID, "csim_sync: ctx is null");
    CSIM_CUDA(cudaSetDevice(c->device));
    CSIM_CUDA(cudaStreamSynchronize(c->stream));
    CSIM_CUDA(cudaStreamSynchronize(c->stream_x));
    CSIM_CUDA(cudaStreamSynchronize(c->stream_copy));
    if ((c->h_err && *c->h_err) || (c->h_couple_err && *c->h_couple_err))
        return fail(CSIM_ERR_TIMEOUT, "csim_sync: a neighbour's halo did not arrive within the bounded wait");
    return CSIM_OK;
}

void* csim_ctx_stream(csim_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }
int csim_ctx_device(const csim_ctx* c) { return c ? c->device : -1; }
uint64_t csim_ctx_launch_count(const csim_ctx* c) { return c ? c->launches : 0; }

int csim_field_create(csim_ctx* c, int nx, int ny, int halo, double dx, double dy,
                      csim_field** out) {
    CSIM_REQUIRE(c != nullptr && out != nullptr, CSIM_ERR_INVALID, "csim_field_create: null argument");
    *out = nullptr;
    CSIM_REQUIRE(nx >= 0 && ny >= 0, CSIM_ERR_INVALID, "csim_field_create: negative size");
    CSIM_REQUIRE(halo >= 0 && halo <= kMaxHalo, CSIM_ERR_UNSUPPORTED,
                 "csim_field_create: halo must be in [0, 8]");
    CSIM_CUDA(cudaSetDevice(c->device));
    csim_field* f = new (std::nothrow) csim_field();
    CSIM_REQUIRE(f != nullptr, CSIM_ERR_NOMEM, "csim_field_create: host allocation failed");
    f->ctx = c;
    f->nx = nx;
    f->ny = ny;
    f->h = halo;
    f->dx = dx;
    f->dy = dy;
    f->pitch = (static_cast<int64_t>(kLeadX) + nx + kTailX + 15) / 16 * 16;
    f->rows = static_cast<int64_t>(ny) + 2 * kLeadY;
    const size_t bytes = static_cast<size_t>(f->pitch) * static_cast<size_t>(f->rows) * sizeof(double);
    cudaError_t e = cudaMalloc(&f->base, bytes);
    if (e != cudaSuccess) {
        delete f;
        return cuda_fail(e, "cudaMalloc(field)", __FILE__, __LINE__);
    }
    // std::vector<double>(n, 0.0) — src/field.cpp:12
    e = cudaMemsetAsync(f->base, 0, bytes, c->stream);
    if (e != cudaSuccess) {
        cudaFree(f->base);
        delete f;
        return cuda_fail(e, "cudaMemsetAsync(field)", __FILE__, __LINE__);
    }
    c->fields.push_back(f);
    *out = f;
    return CSIM_OK;
}

int csim_field_destroy(csim_field* f) {
    if (!f) return CSIM_OK;
    if (!f->ctx) {  // the context went first and already released the device memory
        delete f;
        return CSIM_OK;
    }
    auto& reg = f->ctx->fields;
    reg.erase(std::remove(reg.begin(), reg.end(), f), reg.end());
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    cudaStreamSynchronize(f->ctx->stream_x);
    // a tile the neighbours have mapped takes the peer links down with it: a later tile may get the same
    // address, and storing through stale mappings would corrupt the neighbours
    if (f->ctx->peer_ready && (f->base == f->ctx->peer_tile[0] || f->base == f->ctx->peer_tile[1]))
        peer_teardown(f->ctx);
    if (f->base) cudaFree(f->base);
    delete f;
    return CSIM_OK;
}

int csim_field_get_info(const csim_field* f, csim_field_info* o) {
    CSIM_REQUIRE(f != nullptr && o != nullptr, CSIM_ERR_INVALID, "csim_field_get_info: null argument");
    o->nx = f->nx;
    o->ny = f->ny;
    o->halo = f->h;
    o->dx = f->dx;
    o->dy = f->dy;
    o->pitch = f->pitch;
    o->lead_x = kLeadX;
    o->lead_y = kLeadY;
    o->rows = f->rows;
    o->base = f->base;
    o->interior = f->interior();
    return CSIM_OK;
}

int csim_field_fill(csim_field* f, double value) {
    CSIM_REQUIRE(f != nullptr, CSIM_ERR_INVALID, "csim_field_fill: field is null");
    csim_ctx* c = f->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    // Field::fill covers the padded tile (src/field.cpp:31); filling the whole allocation is a
    // superset and keeps the wide-halo padding defined.
    const int64_t n = f->pitch * f->rows;
    CSIM_LAUNCH(c, k_fill, c->sm_count * 8, 256, 0, f->base, n, value);
    f->values = csim_field::kUnknown;
    return CSIM_OK;
}

static int copy2d(const csim_field* f, void* dst, size_t dpitch, const void* src, size_t spitch,
                  size_t width_doubles, size_t height, cudaMemcpyKind kind, bool sync) {
    if (width_doubles == 0 || height == 0) return CSIM_OK;
    csim_ctx* c = f->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    CSIM_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_doubles * sizeof(double), height, kind,
                                c->stream));
    if (sync) CSIM_CUDA(cudaStreamSynchronize(c->stream));
    return CSIM_OK;
}

// Large asynchronous transfers go through a dense device staging buffer: ONE contiguous DMA plus a
// re-pitch kernel instead of cudaMemcpy2DAsync, whose per-row descriptors (8194 of them for an 8192²
// tile) kept the calling thread inside the driver for most of the copy and thereby held up every other
// host thread's CUDA calls (tools/e2e_timeline.py: a lane's H2D was issued only after another lane's
// D2H had finished).
static constexpr size_t kStagedMinBytes = size_t(4) << 20;
static int ensure_stage(csim_ctx* c, size_t doubles) {
    if (c->stage_doubles >= doubles) return CSIM_OK;
    if (c->d_stage) {
        CSIM_CUDA(cudaStreamSynchronize(c->stream));
        CSIM_CUDA(cudaFree(c->d_stage));
        c->d_stage = nullptr;
        c->stage_doubles = 0;
    }
    CSIM_CUDA(cudaMalloc(&c->d_stage, doubles * sizeof(double)));
    c->stage_doubles = doubles;
    return CSIM_OK;
}
static dim3 repitch_grid(int width, int height) { return dim3((width + 255) / 256, height < 2048 ? height : 2048); }

int csim_field_upload(csim_field* f, const double* host) {
    CSIM_REQUIRE(f != nullptr && host != nullptr, CSIM_ERR_INVALID, "csim_field_upload: null argument");
    f->values = csim_field::kUnknown;
    return copy2d(f, f->at(0, 0), f->pitch * sizeof(double), host, f->nxt() * sizeof(double), f->nxt(),
                  f->nyt(), cudaMemcpyHostToDevice, true);
}
int csim_field_upload_async(csim_field* f, const double* host) {
    CSIM_REQUIRE(f != nullptr && host != nullptr, CSIM_ERR_INVALID, "csim_field_upload_async: null argument");
    f->values = csim_field::kUnknown;
    const size_t n = static_cast<size_t>(f->nxt()) * static_cast<size_t>(f->nyt());
    if (n * sizeof(double) >= kStagedMinBytes) {
        csim_ctx* c = f->ctx;
        CSIM_CUDA(cudaSetDevice(c->device));
        if (int rc = ensure_stage(c, n)) return rc;
        CSIM_CUDA(cudaMemcpyAsync(c->d_stage, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CSIM_LAUNCH(c, k_repitch, repitch_grid(f->nxt(), f->nyt()), 256, 0, c->d_stage, f->nxt(), f->at(0, 0),
                    f->pitch, f->nxt(), f->nyt());
        return CSIM_OK;
    }
    return copy2d(f, f->at(0, 0), f->pitch * sizeof(double), host, f->nxt() * sizeof(double), f->nxt(),
                  f->nyt(), cudaMemcpyHostToDevice, false);
}
int csim_field_download(const csim_field* f, double* host) {
    CSIM_REQUIRE(f != nullptr && host != nullptr, CSIM_ERR_INVALID, "csim_field_download: null argument");
    return copy2d(f, host, f->nxt() * sizeof(double), f->at(0, 0), f->pitch * sizeof(double), f->nxt(),
                  f->nyt(), cudaMemcpyDeviceToHost, true);
}
int csim_field_download_interior(const csim_field* f, double* host) {
    CSIM_REQUIRE(f != nullptr && host != nullptr, CSIM_ERR_INVALID,
                 "csim_field_download_interior: null argument");
    return copy2d(f, host, f->nx * sizeof(double), f->interior(), f->pitch * sizeof(double), f->nx, f->ny,
                  cudaMemcpyDeviceToHost, true);
}
int csim_field_download_window(const csim_field* f, int x0, int y0, int w, int h, double* host) {
    CSIM_REQUIRE(f != nullptr && host != nullptr, CSIM_ERR_INVALID, "csim_field_download_window: null argument");
    CSIM_REQUIRE(w >= 0 && h >= 0, CSIM_ERR_INVALID, "csim_field_download_window: negative extent");
    CSIM_REQUIRE(x0 >= -f->h && y0 >= -f->h && static_cast<int64_t>(x0) + w <= static_cast<int64_t>(f->nx) + f->h &&
                     static_cast<int64_t>(y0) + h <= static_cast<int64_t>(f->ny) + f->h,
                 CSIM_ERR_RANGE, "Field index out of range");
    return copy2d(f, host, static_cast<size_t>(w) * sizeof(double),
                  f->interior() + static_cast<int64_t>(y0) * f->pitch + x0, f->pitch * sizeof(double),
                  static_cast<size_t>(w), static_cast<size_t>(h), cudaMemcpyDeviceToHost, true);
}
int csim_field_download_interior_async(const csim_field* f, double* host) {
    CSIM_REQUIRE(f != nullptr && host != nullptr, CSIM_ERR_INVALID,
                 "csim_field_download_interior_async: null argument");
    const size_t n = static_cast<size_t>(f->nx) * static_cast<size_t>(f->ny);
    if (n * sizeof(double) >= kStagedMinBytes) {
        csim_ctx* c = f->ctx;
        CSIM_CUDA(cudaSetDevice(c->device));
        if (int rc = ensure_stage(c, n)) return rc;
        CSIM_LAUNCH(c, k_repitch, repitch_grid(f->nx, f->ny), 256, 0, f->interior(), f->pitch, c->d_stage, f->nx,
                    f->nx, f->ny);
        CSIM_CUDA(cudaMemcpyAsync(host, c->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        return CSIM_OK;
    }
    return copy2d(f, host, f->nx * sizeof(double), f->interior(), f->pitch * sizeof(double), f->nx, f->ny,
                  cudaMemcpyDeviceToHost, false);
}

int csim_field_download_interior_be_async(const csim_field* f, void* host) {
    CSIM_REQUIRE(f != nullptr && host != nullptr, CSIM_ERR_INVALID,
                 "csim_field_download_interior_be_async: null argument");
    csim_ctx* c = f->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    const size_t n = static_cast<size_t>(f->nx) * static_cast<size_t>(f->ny);
    if (n == 0) return CSIM_OK;
    if (c->snap_doubles < n) {
        if (c->d_snap) {
            CSIM_CUDA(cudaStreamSynchronize(c->stream));
            CSIM_CUDA(cudaFree(c->d_snap));
            c->d_snap = nullptr;
        }
        CSIM_CUDA(cudaMalloc(&c->d_snap, n * sizeof(double)));
        c->snap_doubles = n;
    }
    const dim3 block(256), grid((f->nx + 255) / 256, f->ny < 1024 ? f->ny : 1024);
    CSIM_LAUNCH(c, k_pack_interior_be, grid, block, 0, f->interior(), f->nx, f->ny, f->pitch,
                reinterpret_cast<unsigned long long*>(c->d_snap));
    CSIM_CUDA(cudaMemcpyAsync(host, c->d_snap, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return CSIM_OK;
}

int csim_field_snapshot_async(const csim_field* f, void* host, int big_endian, void** event) {
    CSIM_REQUIRE(f != nullptr && host != nullptr && event != nullptr, CSIM_ERR_INVALID,
                 "csim_field_snapshot_async: null argument");
    *event = nullptr;
    csim_ctx* c = f->ctx;
    CSIM_CUDA(cudaSetDevice(c->device));
    const size_t n = static_cast<size_t>(f->nx) * static_cast<size_t>(f->ny);
    const int b = c->snap_next;
    c->snap_next ^= 1;
    if (c->snapbuf_doubles[b] < n) {
        if (c->d_snapbuf[b]) {
            CSIM_CUDA(cudaStreamSynchronize(c->stream_copy));
            CSIM_CUDA(cudaFree(c->d_snapbuf[b]));
            c->d_snapbuf[b] = nullptr;
            c->snapbuf_doubles[b] = 0;
        }
        CSIM_CUDA(cudaMalloc(&c->d_snapbuf[b], (n ? n : 1) * sizeof(double)));
        c->snapbuf_doubles[b] = n;
    }
    if (n) {
        // the copy that last read this staging buffer must be done before the pack overwrites it
        CSIM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_snap_done[b], 0));
        if (big_endian) {
            const dim3 grid((f->nx + 255) / 256, f->ny < 1024 ? f->ny : 1024);
            CSIM_LAUNCH(c, k_pack_interior_be, grid, 256, 0, f->interior(), f->nx, f->ny, f->pitch,
                        reinterpret_cast<unsigned long long*>(c->d_snapbuf[b]));
        } else {
            CSIM_LAUNCH(c, k_repitch, repitch_grid(f->nx, f->ny), 256, 0, f->interior(), f->pitch, c->d_snapbuf[b],
                        f->nx, f->nx, f->ny);
        }
        CSIM_CUDA(cudaEventRecord(c->ev_snap_packed[b], c->stream));
        CSIM_CUDA(cudaStreamWaitEvent(c->stream_copy, c->ev_snap_packed[b], 0));
        CSIM_CUDA(cudaMemcpyAsync(host, c->d_snapbuf[b], n * sizeof(double), cudaMemcpyDeviceToHost, c->stream_copy));
        CSIM_CUDA(cudaEventRecord(c->ev_snap_done[b], c->stream_copy));
    }
    cudaEvent_t e;
    CSIM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
    CSIM_CUDA(cudaEventRecord(e, n ? c->stream_copy : c->stream));
    *event = e;
    return CSIM_OK;
}

int csim_event_record(csim_ctx* c, void** event) {
    CSIM_REQUIRE(c != nullptr && event != nullptr, CSIM_ERR_INVALID, "csim_event_record: null argument");
    CSIM_CUDA(cudaSetDevice(c->device));
    cudaEvent_t e;
    CSIM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
    CSIM_CUDA(cudaEventRecord(e, c->stream));
    *event = e;
    return CSIM_OK;
}
int csim_event_wait(csim_ctx* c, void* event) {
    CSIM_REQUIRE(c != nullptr && event != nullptr, CSIM_ERR_INVALID, "csim_event_wait: null argument");
    CSIM_CUDA(cudaSetDevice(c->device));
    cudaEvent_t e = static_cast<cudaEvent_t>(event);
    CSIM_CUDA(cudaEventSynchronize(e));
    CSIM_CUDA(cudaEventDestroy(e));
    return CSIM_OK;
}

// check_bounds of src/field.cpp:14-18
static bool in_bounds(const csim_field* f, int i, int j) {
    return !(i < 0 || j < 0 || i >= f->nxt() || j >= f->nyt());
}
int csim_field_get(const csim_field* f, int i, int j, double* value) {
    CSIM_REQUIRE(f != nullptr && value != nullptr, CSIM_ERR_INVALID, "csim_field_get: null argument");
    CSIM_REQUIRE(in_bounds(f, i, j), CSIM_ERR_RANGE, "Field index out of range");
    CSIM_CUDA(cudaSetDevice(f->ctx->device));
    CSIM_CUDA(cudaMemcpyAsync(value, f->at(i, j), sizeof(double), cudaMemcpyDeviceToHost, f->ctx->stream));
    CSIM_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return CSIM_OK;
}
int csim_field_set(csim_field* f, int i, int j, double value) {
    CSIM_REQUIRE(f != nullptr, CSIM_ERR_INVALID, "csim_field_set: field is null");
    CSIM_REQUIRE(in_bounds(f, i, j), CSIM_ERR_RANGE, "Field index out of range");
    CSIM_CUDA(cudaSetDevice(f->ctx->device));
    CSIM_CUDA(cudaMemcpyAsync(f->at(i, j), &value, sizeof(double), cudaMemcpyHostToDevice, f->ctx->stream));
    CSIM_CUDA(cudaStreamSynchronize(f->ctx->stream));
    f->values = csim_field::kUnknown;
    return CSIM_OK;
}

static bool same_geometry(const csim_field* a, const csim_field* b) {
    return a->ctx == b->ctx && a->nx == b->nx && a->ny == b->ny && a->h == b->h && a->pitch == b->pitch &&
           a->rows == b->rows;
}

int csim_field_swap(csim_field* a, csim_field* b) {
    CSIM_REQUIRE(a != nullptr && b != nullptr, CSIM_ERR_INVALID, "csim_field_swap: null argument");
    CSIM_REQUIRE(same_geometry(a, b), CSIM_ERR_INVALID, "csim_field_swap: tiles differ in geometry");
    double* t = a->base;
    a->base = b->base;
    b->base = t;
    const int v = a->values;
    a->values = b->values;
    b->values = v;
    return CSIM_OK;
}

int csim_field_copy(const csim_field* src, csim_field* dst) {
    CSIM_REQUIRE(src != nullptr && dst != nullptr, CSIM_ERR_INVALID, "csim_field_copy: null argument");
    CSIM_REQUIRE(same_geometry(src, dst), CSIM_ERR_INVALID, "csim_field_copy: tiles differ in geometry");
    CSIM_CUDA(cudaSetDevice(src->ctx->device));
    CSIM_CUDA(cudaMemcpyAsync(dst->base, src->base,
                              static_cast<size_t>(src->pitch) * src->rows * sizeof(double),
                              cudaMemcpyDeviceToDevice, src->ctx->stream));
    dst->values = src->values;
    return CSIM_OK;
}

int csim_host_alloc(size_t bytes, void** out) {
    CSIM_REQUIRE(out != nullptr, CSIM_ERR_INVALID, "csim_host_alloc: out is null");
    CSIM_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return CSIM_OK;
}
int csim_host_free(void* p) {
    if (p) CSIM_CUDA(cudaFreeHost(p));
    return CSIM_OK;
}

}  // extern "C"
