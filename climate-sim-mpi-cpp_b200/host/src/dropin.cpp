// dropin.cpp — the reference's C++ functions for the timestep path, implemented on the C ABI.
//
// Everything here is glue: argument checks and the host/device coherence of Field::data.  The
// arithmetic lives in libcsim_b200.so (csrc/).  Reference interfaces by file:line are listed in
// csim_dropin.hpp.
#include "csim_dropin.hpp"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
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
            int ndev = 0;
            if (csim_device_count(&ndev) != CSIM_OK) ndev = 0;
            // one rank per GPU of the box, GPU id = rank; with more ranks than GPUs the
            // ranks share them round-robin
            dev = ndev > 0 ? local % ndev : local;
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
csim_field* MirroredData::device_overwrite() {
    State& s = *st_;
    if (!s.dev) check(csim_field_create(default_context(), s.nx, s.ny, s.h, s.dx, s.dy, &s.dev));
    s.host_newer = false;
    s.dev_newer = true;
    return s.dev;
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
    // with halo 1 the interior update plus the ring copy (diffusion.cpp:9-25) rewrite every cell of `out`,
    // so a newer host copy of `out` (main.cpp:104's std::copy leaves one) need not travel to the device
    const bool rewrites_all = out.halo == 1 && u.halo == 1 && out.nx_local == u.nx_local &&
                              out.ny_local == u.ny_local && u.nx_local > 0 && u.ny_local > 0 && &u != &out;
    check(csim_diffusion_step(u.data.device_ro(), rewrites_all ? out.data.device_overwrite() : out.data.device_rw(), D,
                              dt));
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
    if (nsteps <= 0) return;
    csim_field* du = u.data.device_rw();
    // tmp is scratch: every step rewrites it from u (main.cpp:104), so its host copy never needs uploading
    csim_field* dt_ = tmp.data.device_overwrite();
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
// The file lives under $XDG_RUNTIME_DIR when there is one (else /tmp), carries the user id and a job
// token in its name, and starts with a magic word and the same token, so that a file left behind by a
// crashed run or planted by someone else is not mistaken for this job's: CSIM_JOB_ID (any string the
// launcher gives every rank) or, without it, the parent process id (ranks started by one launcher share it).
static std::string job_token() {
    if (const char* j = std::getenv("CSIM_JOB_ID")) return j;
    if (const char* j = std::getenv("TORCHELASTIC_RUN_ID")) return j;
    return std::to_string(static_cast<long long>(getppid()));
}
static std::string rendezvous_path() {
    if (const char* p = std::getenv("CSIM_RENDEZVOUS")) return p;
    const char* dir = std::getenv("XDG_RUNTIME_DIR");
    const char* port = std::getenv("MASTER_PORT");
    std::string tok = job_token();
    for (char& ch : tok)
        if (!std::isalnum(static_cast<unsigned char>(ch))) ch = '_';
    return std::string(dir && *dir ? dir : "/tmp") + "/csim_rendezvous_" + std::to_string(static_cast<long long>(getuid())) +
           "_" + (port ? port : "default") + "_" + tok;
}
namespace {
struct RendezvousRecord {
    char magic[8];
    char token[56];
    char id[CSIM_UNIQUE_ID_BYTES];
};
}  // namespace

int MPI_Init(int*, char***) {
    auto& w = csim_host::world();
    if (w.initialized) return MPI_SUCCESS;
    w.initialized = true;
    if (w.size <= 1) return MPI_SUCCESS;
    csim_ctx* ctx = csim_host::default_context();
    RendezvousRecord rec;
    std::memset(&rec, 0, sizeof rec);
    std::memcpy(rec.magic, "CSIMRDV1", 8);
    std::strncpy(rec.token, job_token().c_str(), sizeof rec.token - 1);
    const std::string path = rendezvous_path();
    if (w.rank == 0) {
        check(csim_comm_unique_id(rec.id));
        const std::string tmp = path + ".tmp";
        ::unlink(tmp.c_str());
        ::unlink(path.c_str());  // a stale record of an earlier run must not be read as ours
        const int fd = ::open(tmp.c_str(), O_CREAT | O_EXCL | O_NOFOLLOW | O_WRONLY, 0600);
        if (fd < 0) throw std::runtime_error("MPI_Init (csim shim): cannot create " + tmp + ": " + std::strerror(errno));
        const bool ok = ::write(fd, &rec, sizeof rec) == static_cast<ssize_t>(sizeof rec);
        ::close(fd);
        if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0)
            throw std::runtime_error("MPI_Init (csim shim): cannot publish " + path);
    } else {
        bool got = false;
        RendezvousRecord in_rec;
        for (int tries = 0; tries < 1200 && !got; ++tries) {  // up to 60 s
            const int fd = ::open(path.c_str(), O_RDONLY | O_NOFOLLOW);
            if (fd >= 0) {
                const bool whole = ::read(fd, &in_rec, sizeof in_rec) == static_cast<ssize_t>(sizeof in_rec);
                ::close(fd);
                got = whole && std::memcmp(in_rec.magic, rec.magic, 8) == 0 &&
                      std::memcmp(in_rec.token, rec.token, sizeof rec.token) == 0;
            }
            if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
        if (!got) throw std::runtime_error("MPI_Init (csim shim): no rendezvous record for this job at " + path);
        std::memcpy(rec.id, in_rec.id, sizeof rec.id);
    }
    check(csim_comm_init(ctx, w.size, w.rank, rec.id));
    check(csim_comm_allreduce_max(ctx, nullptr, 0));  // barrier: everyone has read the id
    if (w.rank == 0) ::unlink(path.c_str());
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
