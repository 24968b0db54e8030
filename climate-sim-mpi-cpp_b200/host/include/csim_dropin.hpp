// csim_dropin.hpp — the reference's C++ interface for the timestep path, backed by libcsim_b200.so.
//
// A maintainer of climate-sim-mpi-cpp switches the hot path to the GPU by putting this directory
// first on the include path and linking libcsim_dropin + libcsim_b200 instead of the reference's
// field/diffusion/advection/boundary/halo/decomp objects.  The small headers next
// to this file (field.hpp, diffusion.hpp, …) carry the reference's file names and forward here, so
// `#include "diffusion.hpp"` keeps working unchanged.
//
// Same names, argument meaning and error behaviour as the reference:
//   Field                      include/field.hpp:5-21, src/field.cpp:6-31
//   diffusion_step             include/diffusion.hpp:4
//   advection_step             include/advection.hpp:4      (accumulates into `out`)
//   BCType, BCConfig, apply_boundary   include/boundary.hpp:5-14
//   Decomp2D                   include/decomp.hpp:4-17
//   exchange_halos             include/halo.hpp:7
//   safe_dt                    include/stability.hpp:5-16
// What differs: Field::data is not a std::vector but a host mirror of a device tile that knows
// which side is newer.  Reading or writing it from the host (begin()/end()/[]/at(), Field::at)
// pulls the tile back when the device copy is newer; the step functions push it when the host copy
// is newer.  std::swap(u.data, tmp.data) and std::copy over data work as in src/main.cpp:104-109.
// Code that pokes cells between kernels (the unit tests) therefore stays correct; code that only
// calls the step functions (the time loop) never leaves the device.
#pragma once
#include <mpi.h>

#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "csim.h"

namespace csim_host {

// Process-wide context: GPU `CSIM_DEVICE` (default: the MPI-shim rank modulo the device count).
csim_ctx* default_context();
// Status code → the exception the reference would have thrown (RANGE → std::out_of_range,
// anything else → std::runtime_error with csim_last_error()).
void check(int rc);

// Host mirror + device tile of one Field.
class MirroredData {
public:
    using value_type = double;
    using iterator = double*;
    using const_iterator = const double*;

    MirroredData() = default;
    MirroredData(int nx, int ny, int h, double dx, double dy);
    MirroredData(const MirroredData& o);
    MirroredData(MirroredData&& o) noexcept = default;
    MirroredData& operator=(const MirroredData& o);
    MirroredData& operator=(MirroredData&& o) noexcept;
    ~MirroredData();

    std::size_t size() const { return st_ ? st_->host.size() : 0; }
    bool empty() const { return size() == 0; }
    iterator begin() { return host_rw(); }
    iterator end() { return host_rw() + size(); }
    const_iterator begin() const { return host_ro(); }
    const_iterator end() const { return host_ro() + size(); }
    const_iterator cbegin() const { return host_ro(); }
    const_iterator cend() const { return host_ro() + size(); }
    double* data() { return host_rw(); }
    const double* data() const { return host_ro(); }
    double& operator[](std::size_t k) { return host_rw()[k]; }
    const double& operator[](std::size_t k) const { return host_ro()[k]; }
    double& at(std::size_t k);
    const double& at(std::size_t k) const;

    void swap(MirroredData& o) noexcept { st_.swap(o.st_); }

    // device side (used by the step functions)
    csim_field* device_ro() const;  // tile with current contents, for reading
    csim_field* device_rw();        // same, and marks the device copy as the newer one
    csim_field* device_overwrite(); // the tile WITHOUT its current contents (the caller rewrites every cell):
                                    // no upload of a newer host copy; marks the device copy as the newer one
    void fill_device(double v);     // Field::fill without touching the host copy

private:
    struct State {
        int nx = 0, ny = 0, h = 0;
        double dx = 1.0, dy = 1.0;
        std::vector<double> host;
        csim_field* dev = nullptr;
        bool host_newer = false;  // host copy has writes the device has not seen
        bool dev_newer = false;   // device copy has writes the host has not seen
    };
    double* host_rw();
    const double* host_ro() const;
    void pull() const;  // device → host if dev_newer
    void push() const;  // host → device if host_newer (creates the tile on first use)
    mutable std::unique_ptr<State> st_;
};
inline void swap(MirroredData& a, MirroredData& b) noexcept { a.swap(b); }

}  // namespace csim_host

// ---- Field -----------------------------------------------------------------------------------
struct Field {
    int nx_local, ny_local;
    int halo;
    double dx, dy;
    csim_host::MirroredData data;

    Field(int nx, int ny, int h, double dx_, double dy_);

    std::size_t idx(int i, int j) const;  // throws std::out_of_range like src/field.cpp:14-25
    double& at(int i, int j);
    const double& at(int i, int j) const;

    int nx_total() const { return nx_local + 2 * halo; }
    int ny_total() const { return ny_local + 2 * halo; }

    void fill(double value);
};

// ---- decomposition ---------------------------------------------------------------------------
struct Decomp2D {
    MPI_Comm cart_comm = MPI_COMM_NULL;
    int dims[2]{0, 0};
    int coords[2]{0, 0};
    int nbr_lr[2]{MPI_PROC_NULL, MPI_PROC_NULL};
    int nbr_du[2]{MPI_PROC_NULL, MPI_PROC_NULL};

    int nx_global = 0, ny_global = 0;
    int nx_local = 0, ny_local = 0;
    int x_offset = 0, y_offset = 0;

    void init(MPI_Comm comm_world, int nx_global_, int ny_global_);
    void finalize();
};

// ---- boundary conditions ---------------------------------------------------------------------
enum class BCType { Dirichlet, Neumann, Periodic };

struct BCConfig {
    BCType left = BCType::Dirichlet;
    BCType right = BCType::Dirichlet;
    BCType bottom = BCType::Dirichlet;
    BCType top = BCType::Dirichlet;
};

// ---- the step functions ------------------------------------------------------------------------
void diffusion_step(const Field& u, Field& out, double D, double dt);
void advection_step(const Field& u, Field& out, double vx, double vy, double dt);
void apply_boundary(Field& f, const Decomp2D& dec, const BCConfig& bc, double value);
void exchange_halos(Field& f, const Decomp2D& dec, MPI_Comm comm);
double safe_dt(double dx, double dy, double vx, double vy, double D);

// ---- beyond the reference: the fused loop body -------------------------------------------------
// `nsteps` iterations of src/main.cpp:101-109 (exchange, boundary, copy, diffusion, advection,
// swap) without leaving the device; on return `u` holds the newest state.  This is what the
// driver in host/src/main.cpp calls instead of the five separate statements.
void run_timesteps(Field& u, Field& tmp, const Decomp2D& dec, const BCConfig& bc, double D, double vx,
                   double vy, double dt, int nsteps);
// std::min_element / std::max_element over Field::data (src/main.cpp:74-75: the padded tile, ghosts
// included) as a device reduction: the tile does not travel to the host for two numbers.
std::pair<double, double> field_minmax(const Field& f);
