// dropin.cpp — the reference's C++ functions for the timestep path, implemented on the C ABI.
//
// Everything here is glue: argument checks and the host/device coherence of Field::data.  The
// arithmetic lives in libcsim_b200.so (csrc/).  Reference interfaces by file:line are listed in
// csim_dropin.hpp.
#include "csim_dropin.hpp"

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <thread>

namespace csim_host {

// ---- MPI shim state ----------------------------------------------------------------------------
namespace {
int env_int(const char* a, const char* b, int dflt) {
    for (const char* n : {a, b}) {
        if (!n) continue;
        if (const char* v = std::getenv(n))
            if (*v) return std::atoi(v);
    }
    return dflt;
}
struct World {
    int rank = env_int("CSIM_RANK", "RANK", 0);
    int size = env_int("CSIM_WORLD_SIZE", "WORLD_SIZE", 1);
    bool initialized = false, finalized = false;
};
World& world() {
    static World w;
    return w;
}
csim_ctx* g_ctx = nullptr;
std::once_flag g_ctx_once;
}  // namespace

void check(int rc) {
    if (rc == CSIM_OK) return;
    const std::string msg = csim_last_error();
    if (rc == CSIM_ERR_RANGE) throw std::out_of_range(msg);  // src/field.cpp:16
    throw std::runtime_error(msg);
}

csim_ctx* default_context() {
    std::call_once(g_ctx_once, [] {
        int dev = env_int("CSIM_DEVICE", nullptr, -1);
        if (dev < 0) {
            const int local = env_int("CSIM_LOCAL_RANK", "LOCAL_RANK", world().rank);
            dev = local;  // one rank per GPU of the box, GPU id = rank
        }
        check(csim_ctx_create(dev, &g_ctx));
    });
    return g_ctx;
}

// ---- MirroredData -------------------------------------------------------------------------------
MirroredData::MirroredData(int nx, int ny, int h, double dx, double dy) : st_(new State) {
    st_->nx = nx;
    st_->ny = ny;
    st_->h = h;
    st_->dx = dx;
    st_->dy = dy;
    // 64-bit size: the reference multiplies in int (src/field.cpp:12) and overflows above ~46339^2
    st_->host.assign(static_cast<std::size_t>(nx + 2 * h) * static_cast<std::size_t>(ny + 2 * h), 0.0);
}
MirroredData::MirroredData(const MirroredData& o) {
    if (!o.st_) return;
    o.pull();
    st_.reset(new State(*o.st_));
    st_->dev = nullptr;
    st_->dev_newer = false;
    st_->host_newer = true;
}
MirroredData& MirroredData::operator=(const MirroredData& o) {
    if (this != &o) {
        MirroredData t(o);
        swap(t);
    }
    return *this;
}
MirroredData& MirroredData::operator=(MirroredData&& o) noexcept {
    if (this != &o) {
        if (st_ && st_->dev) csim_field_destroy(st_->dev);
        st_ = std::move(o.st_);
    }
    return *this;
}
MirroredData::~MirroredData() {
    if (st_ && st_->dev) csim_field_destroy(st_->dev);
}

void MirroredData::pull() const {
    State& s = *st_;
    if (s.dev && s.dev_newer) {
        check(csim_field_download(s.dev, s.host.data()));
        s.dev_newer = false;
    }
}
void MirroredData::push() const {
    State& s = *st_;
    // a new tile is zero-filled like the host vector (src/field.cpp:12): nothing to upload unless the
    // host copy has been handed out for writing (host_rw) or was copy-constructed (both set host_newer)
    if (!s.dev) check(csim_field_create(default_context(), s.nx, s.ny, s.h, s.dx, s.dy, &s.dev));
    if (s.host_newer) {
        if (!s.host.empty()) check(csim_field_upload(s.dev, s.host.data()));
        s.host_newer = false;
    }
}
double* MirroredData::host_rw() {
    pull();
    st_->host_newer = true;
    return st_->host.data();
}
const double* MirroredData::host_ro() const {
    pull();
    return st_->host.data();
}
double& MirroredData::at(std::size_t k) {
    if (k >= size()) throw std::out_of_range("vector::_M_range_check");
    return host_rw()[k];
}
const double& MirroredData::at(std::size_t k) const {
    if (k >= size()) throw std::out_of_range("vector::_M_range_check");
    return host_ro()[k];
}
csim_field* MirroredData::device_ro() const {
    push();
    return st_->dev;
}
csim_field* MirroredData::device_rw() {
    push();
    st_->dev_newer = true;
    return st_->dev;
}
void MirroredData::fill_device(double v) {
    State& s = *st_;
    if (!s.dev) check(csim_field_create(default_context(), s.nx, s.ny, s.h, s.dx, s.dy, &s.dev));
    check(csim_field_fill(s.dev, v));
    s.host_newer = false;
    s.dev_newer = true;
}

}  // namespace csim_host

using csim_host::check;

// ---- Field: src/field.cpp:6-31 -----------------------------------------------------------------
Field::Field(int nx, int ny, int h, double dx_, double dy_)
    : nx_local(nx), ny_local(ny), halo(h), dx(dx_), dy(dy_), data(nx, ny, h, dx_, dy_) {}

std::size_t Field::idx(int i, int j) const {
    const int nxt = nx_total(), nyt = ny_total();
    if (i < 0 || j < 0 || i >= nxt || j >= nyt) throw std::out_of_range("Field index out of range");
    return static_cast<std::size_t>(j) * static_cast<std::size_t>(nxt) + static_cast<std::size_t>(i);
}
double& Field::at(int i, int j) { return data.at(idx(i, j)); }
const double& Field::at(int i, int j) const { return data.at(idx(i, j)); }
void Field::fill(double value) { data.fill_device(value); }

// ---- Decomp2D: src/decomp.cpp:5-39 ---------------------------------------------------------------
static csim_decomp to_c(const Decomp2D& d) {
    csim_decomp c;
    for (int k = 0; k < 2; ++k) {
        c.dims[k] = d.dims[k];
        c.coords[k] = d.coords[k];
    }
    c.nbr[CSIM_LEFT] = d.nbr_lr[0];
    c.nbr[CSIM_RIGHT] = d.nbr_lr[1];
    c.nbr[CSIM_BOTTOM] = d.nbr_du[0];
    c.nbr[CSIM_TOP] = d.nbr_du[1];
    c.nx_global = d.nx_global;
    c.ny_global = d.ny_global;
    c.nx_local = d.nx_local;
    c.ny_local = d.ny_local;
    c.x_offset = d.x_offset;
    c.y_offset = d.y_offset;
    return c;
}

void Decomp2D::init(MPI_Comm comm_world, int nxg, int nyg) {
    int size = 1, rank = 0;
    MPI_Comm_size(comm_world, &size);
    MPI_Comm_rank(comm_world, &rank);
    csim_decomp c;
    check(csim_decomp_init(size, rank, nxg, nyg, &c));
    cart_comm = comm_world;
    for (int k = 0; k < 2; ++k) {
        dims[k] = c.dims[k];
        coords[k] = c.coords[k];
    }
    nbr_lr[0] = c.nbr[CSIM_LEFT];
    nbr_lr[1] = c.nbr[CSIM_RIGHT];
    nbr_du[0] = c.nbr[CSIM_BOTTOM];
    nbr_du[1] = c.nbr[CSIM_TOP];
    nx_global = nxg;
    ny_global = nyg;
    nx_local = c.nx_local;
    ny_local = c.ny_local;
    x_offset = c.x_offset;
    y_offset = c.y_offset;
}
void Decomp2D::finalize() { cart_comm = MPI_COMM_NULL; }

// ---- step functions ------------------------------------------------------------------------------
void diffusion_step(const Field& u, Field& out, double D, double dt) {
    check(csim_diffusion_step(u.data.device_ro(), out.data.device_rw(), D, dt));
}
void advection_step(const Field& u, Field& out, double vx, double vy, double dt) {
    check(csim_advection_step(u.data.device_ro(), out.data.device_rw(), vx, vy, dt));
}
static void bc_ints(const BCConfig& bc, int out[4]) {
    out[0] = static_cast<int>(bc.left);
    out[1] = static_cast<int>(bc.right);
    out[2] = static_cast<int>(bc.bottom);
    out[3] = static_cast<int>(bc.top);
}
void apply_boundary(Field& f, const Decomp2D& dec, const BCConfig& bc, double value) {
    const int nbr[4] = {dec.nbr_lr[0], dec.nbr_lr[1], dec.nbr_du[0], dec.nbr_du[1]};
    int b[4];
    bc_ints(bc, b);
    check(csim_apply_boundary(f.data.device_rw(), nbr, b, value));
}
void exchange_halos(Field& f, const Decomp2D& dec, MPI_Comm) {
    const csim_decomp c = to_c(dec);
    check(csim_halo_exchange(f.data.device_rw(), &c));
}
double safe_dt(double dx, double dy, double vx, double vy, double D) { return csim_safe_dt(dx, dy, vx, vy, D); }

void run_timesteps(Field& u, Field& tmp, const Decomp2D& dec, const BCConfig& bc, double D, double vx, double vy,
                   double dt, int nsteps) {
    csim_step_params p;
    p.D = D;
    p.vx = vx;
    p.vy = vy;
    p.dt = dt;
    bc_ints(bc, p.bc);
    const csim_decomp c = to_c(dec);
    for (int s = 0; s < 4; ++s) p.nbr[s] = c.nbr[s];
    p.bc_value = 0.0;  // src/main.cpp:102
    p.flags = 0;
    csim_field* du = u.data.device_rw();
    csim_field* dt_ = tmp.data.device_rw();
    // more than one rank: halos travel as packed T-line bands over grouped NCCL send/recv
    // csim_run_steps swaps the device buffers of the two tiles an odd or even number of times; the
    // newest state ends up in u's handle either way.
    check(csim_run_steps(du, dt_, &p, &c, nsteps));
}

// ---- MPI shim ------------------------------------------------------------------------------------
extern "C" {

int MPI_Comm_rank(MPI_Comm, int* rank) {
    *rank = csim_host::world().rank;
    return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm, int* size) {
    *size = csim_host::world().size;
    return MPI_SUCCESS;
}
int MPI_Initialized(int* flag) {
    *flag = csim_host::world().initialized ? 1 : 0;
    return MPI_SUCCESS;
}
int MPI_Finalized(int* flag) {
    *flag = csim_host::world().finalized ? 1 : 0;
    return MPI_SUCCESS;
}
double MPI_Wtime(void) {
    using clk = std::chrono::steady_clock;
    static const clk::time_point t0 = clk::now();
    return std::chrono::duration<double>(clk::now() - t0).count();
}

// Rendezvous of the NCCL id through a file: rank 0 writes it atomically, the others poll.
static std::string rendezvous_path() {
    if (const char* p = std::getenv("CSIM_RENDEZVOUS")) return p;
    const char* port = std::getenv("MASTER_PORT");
    return std::string("/tmp/csim_rendezvous_") + (port ? port : "default");
}

int MPI_Init(int*, char***) {
    auto& w = csim_host::world();
    if (w.initialized) return MPI_SUCCESS;
    w.initialized = true;
    if (w.size <= 1) return MPI_SUCCESS;
    csim_ctx* ctx = csim_host::default_context();
    char id[CSIM_UNIQUE_ID_BYTES];
    const std::string path = rendezvous_path();
    if (w.rank == 0) {
        check(csim_comm_unique_id(id));
        const std::string tmp = path + ".tmp";
        {
            std::ofstream o(tmp, std::ios::binary | std::ios::trunc);
            o.write(id, sizeof id);
        }
        std::rename(tmp.c_str(), path.c_str());
    } else {
        bool got = false;
        for (int tries = 0; tries < 1200 && !got; ++tries) {  // up to 60 s
            std::ifstream in(path, std::ios::binary);
            if (in && in.read(id, sizeof id)) got = true;
            if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
        if (!got) throw std::runtime_error("MPI_Init (csim shim): no rendezvous file " + path);
    }
    check(csim_comm_init(ctx, w.size, w.rank, id));
    check(csim_comm_allreduce_max(ctx, nullptr, 0));  // barrier: everyone has read the id
    if (w.rank == 0) std::remove(path.c_str());
    return MPI_SUCCESS;
}
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided) {
    if (provided) *provided = required;
    return MPI_Init(argc, argv);
}
int MPI_Finalize(void) {
    auto& w = csim_host::world();
    if (csim_host::g_ctx) csim_sync(csim_host::g_ctx);
    w.finalized = true;
    return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm) {
    if (csim_host::world().size > 1 || csim_host::g_ctx) check(csim_comm_allreduce_max(csim_host::default_context(), nullptr, 0));
    return MPI_SUCCESS;
}
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int, MPI_Comm) {
    if (datatype != MPI_DOUBLE || op != MPI_MAX) return 1;
    if (recvbuf != sendbuf) std::memcpy(recvbuf, sendbuf, sizeof(double) * static_cast<std::size_t>(count));
    if (csim_host::world().size > 1)
        check(csim_comm_allreduce_max(csim_host::default_context(), static_cast<double*>(recvbuf), count));
    return MPI_SUCCESS;
}

}  // extern "C"

std::pair<double, double> field_minmax(const Field& f) {
    double mn = 0.0, mx = 0.0;
    check(csim_minmax(f.data.device_ro(), &mn, &mx));
    return {mn, mx};
}
