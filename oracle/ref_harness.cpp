// oracle/ref_harness.cpp — C entry points around the reference's OWN compiled objects.
//
// TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles this file together with the reference's
// unmodified src/{field,diffusion,advection,boundary,decomp,halo,init}.cpp (read in place from
// /root/reference; nothing is copied into this repository) and oracle/stub/mini_mpi.cpp into
// oracle/_ref/libcsim_ref.so.  What is restated here, because src/main.cpp cannot be built
// (io.cpp needs PnetCDF and yaml-cpp, which are absent): the time loop of src/main.cpp:62-118,
// with the snapshot writes replaced by copies of the de-haloed tile into a caller buffer
// (src/io.cpp:411-418 does the same de-halo before its collective put).
//
// Ranks are std::threads, one per emulated MPI rank (stub/mpi.h).  The same entry point is the
// "reference" CPU baseline of bench.py: reference compute code, no MPI launcher.
#include <mpi.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "advection.hpp"
#include "boundary.hpp"
#include "decomp.hpp"
#include "diffusion.hpp"
#include "field.hpp"
#include "halo.hpp"
#include "init.hpp"
#include "io.hpp"
#include "stability.hpp"

// boundary.hpp declares apply_boundary without the default argument (it sits on the definition,
// src/boundary.cpp:12), so callers outside that file always pass `value`.

extern "C" {

// Same member order as orc_config / orc_decomp in oracle_port.c so one ctypes layout serves both.
struct ref_config {
    int nx, ny;
    double dx, dy;
    double D, vx, vy;
    double dt;
    int steps, out_every;
    int bc[4];
    int ic_preset;
    double A, sigma_frac, xc_frac, yc_frac;
};
struct ref_decomp {
    int dims[2], coords[2], nbr_lr[2], nbr_du[2];
    int nx_global, ny_global, nx_local, ny_local, x_offset, y_offset;
};

static thread_local std::string t_err;
const char* ref_last_error(void) { return t_err.c_str(); }

static BCConfig to_bc(const int bc[4]) {
    BCConfig b;
    b.left = static_cast<BCType>(bc[0]);
    b.right = static_cast<BCType>(bc[1]);
    b.bottom = static_cast<BCType>(bc[2]);
    b.top = static_cast<BCType>(bc[3]);
    return b;
}
static SimConfig to_sim(const ref_config& c) {
    SimConfig s;
    s.nx = c.nx;
    s.ny = c.ny;
    s.dx = c.dx;
    s.dy = c.dy;
    s.D = c.D;
    s.vx = c.vx;
    s.vy = c.vy;
    s.dt = c.dt;
    s.steps = c.steps;
    s.out_every = c.out_every;
    s.bc = to_bc(c.bc);
    s.ic.preset = c.ic_preset == 0 ? "gaussian_hotspot" : "constant_zero";
    s.ic.A = c.A;
    s.ic.sigma_frac = c.sigma_frac;
    s.ic.xc_frac = c.xc_frac;
    s.ic.yc_frac = c.yc_frac;
    return s;
}

// --- single-function entry points on raw padded arrays (known-answer tests) -------------------

int ref_diffusion_step(const double* u, double* out, int nx, int ny, int h, double dx, double dy,
                       double D, double dt) {
    Field fu(nx, ny, h, dx, dy), fo(nx, ny, h, dx, dy);
    std::copy(u, u + fu.data.size(), fu.data.begin());
    std::copy(out, out + fo.data.size(), fo.data.begin());
    diffusion_step(fu, fo, D, dt);
    std::copy(fo.data.begin(), fo.data.end(), out);
    return 0;
}
int ref_advection_step(const double* u, double* out, int nx, int ny, int h, double dx, double dy,
                       double vx, double vy, double dt) {
    Field fu(nx, ny, h, dx, dy), fo(nx, ny, h, dx, dy);
    std::copy(u, u + fu.data.size(), fu.data.begin());
    std::copy(out, out + fo.data.size(), fo.data.begin());
    advection_step(fu, fo, vx, vy, dt);
    std::copy(fo.data.begin(), fo.data.end(), out);
    return 0;
}
int ref_apply_boundary(double* f, int nx, int ny, int h, const int nbr[4], const int bc[4],
                       double value) {
    Field ff(nx, ny, h, 1.0, 1.0);
    std::copy(f, f + ff.data.size(), ff.data.begin());
    Decomp2D dec;
    dec.nbr_lr[0] = nbr[0];
    dec.nbr_lr[1] = nbr[1];
    dec.nbr_du[0] = nbr[2];
    dec.nbr_du[1] = nbr[3];
    apply_boundary(ff, dec, to_bc(bc), value);
    std::copy(ff.data.begin(), ff.data.end(), f);
    return 0;
}
double ref_safe_dt(double dx, double dy, double vx, double vy, double D) {
    return safe_dt(dx, dy, vx, vy, D);
}
// Field::at bounds behaviour (src/field.cpp:14-29): 1 = throws std::out_of_range, 0 = no throw
int ref_field_at_throws(int nx, int ny, int h, int i, int j) {
    Field f(nx, ny, h, 1.0, 1.0);
    try {
        (void)f.at(i, j);
    } catch (const std::out_of_range&) {
        return 1;
    }
    return 0;
}
long ref_field_index(int nx, int ny, int h, int i, int j) {
    Field f(nx, ny, h, 1.0, 1.0);
    return static_cast<long>(&f.at(i, j) - f.data.data());
}

// Decomp2D::init of every rank (src/decomp.cpp:5-34), run on `size` threads.
int ref_decomp_all(int size, int nxg, int nyg, ref_decomp* out) {
    mini_mpi_set_world(size);
    std::vector<std::thread> th;
    for (int r = 0; r < size; ++r)
        th.emplace_back([=] {
            mini_mpi_set_rank(r);
            Decomp2D d;
            d.init(MPI_COMM_WORLD, nxg, nyg);
            ref_decomp& o = out[r];
            for (int k = 0; k < 2; ++k) {
                o.dims[k] = d.dims[k];
                o.coords[k] = d.coords[k];
                o.nbr_lr[k] = d.nbr_lr[k];
                o.nbr_du[k] = d.nbr_du[k];
            }
            o.nx_global = d.nx_global;
            o.ny_global = d.ny_global;
            o.nx_local = d.nx_local;
            o.ny_local = d.ny_local;
            o.x_offset = d.x_offset;
            o.y_offset = d.y_offset;
            mini_mpi_barrier();
            d.finalize();
        });
    for (auto& t : th) t.join();
    return 0;
}

// exchange_halos (src/halo.cpp:6-50) once on `size` ranks.  tiles[r] points at rank r's padded
// tile (sizes from Decomp2D), updated in place.
int ref_exchange_all(int size, int nxg, int nyg, int h, double** tiles) {
    mini_mpi_set_world(size);
    std::vector<std::thread> th;
    for (int r = 0; r < size; ++r)
        th.emplace_back([=] {
            mini_mpi_set_rank(r);
            Decomp2D d;
            d.init(MPI_COMM_WORLD, nxg, nyg);
            Field f(d.nx_local, d.ny_local, h, 1.0, 1.0);
            std::copy(tiles[r], tiles[r] + f.data.size(), f.data.begin());
            mini_mpi_barrier();
            exchange_halos(f, d, MPI_COMM_WORLD);
            mini_mpi_barrier();
            std::copy(f.data.begin(), f.data.end(), tiles[r]);
            d.finalize();
        });
    for (auto& t : th) t.join();
    return 0;
}

// --- the time loop, src/main.cpp:62-118 -------------------------------------------------------
// Arguments as orc_run in oracle_port.c, plus loop_seconds (max over ranks of the wall time of
// the step loop, the quantity main.cpp:120-128 reports as total_max) and `skip_frames`, which
// drops the per-frame de-halo copies from the loop so the CPU baseline times the step only.
int ref_run(const ref_config* cfg_in, int nranks, int clamp_dt, const double* u0_padded,
            double* frames, int max_frames, double* final_interior, double* final_padded_rank0,
            double* loop_seconds) {
    ref_config c = *cfg_in;
    if (clamp_dt) {
        const double lim = safe_dt(c.dx, c.dy, c.vx, c.vy, c.D);  // main.cpp:42-49
        if (c.dt > lim) c.dt = lim;
    }
    if (u0_padded && nranks != 1) {
        t_err = "u0_padded needs nranks == 1";
        return -2;
    }
    const SimConfig cfg = to_sim(c);
    mini_mpi_set_world(nranks);
    std::vector<double> secs(static_cast<size_t>(nranks), 0.0);
    std::vector<int> nfr(static_cast<size_t>(nranks), 0);
    std::vector<std::string> errs(static_cast<size_t>(nranks));
    const size_t gsz = static_cast<size_t>(cfg.nx) * static_cast<size_t>(cfg.ny);

    auto body = [&](int rank) {
        mini_mpi_set_rank(rank);
        Decomp2D dec;
        dec.init(MPI_COMM_WORLD, cfg.nx, cfg.ny);  // main.cpp:62-63
        const int halo = 1;                        // main.cpp:65
        Field u(dec.nx_local, dec.ny_local, halo, cfg.dx, cfg.dy);
        Field tmp(dec.nx_local, dec.ny_local, halo, cfg.dx, cfg.dy);
        u.fill(0.0);
        tmp.fill(0.0);
        if (u0_padded)
            std::copy(u0_padded, u0_padded + u.data.size(), u.data.begin());
        else
            apply_initial_condition(dec, u, cfg);  // main.cpp:71

        auto dehalo = [&](double* dst) {  // io.cpp:411-418 target region {y_off, x_off}
            for (int j = 0; j < dec.ny_local; ++j)
                std::memcpy(dst + static_cast<size_t>(dec.y_offset + j) * cfg.nx + dec.x_offset,
                            &u.data[static_cast<size_t>(j + halo) * u.nx_total() + halo],
                            sizeof(double) * static_cast<size_t>(dec.nx_local));
        };

        mini_mpi_barrier();  // main.cpp:82
        const auto t0 = std::chrono::steady_clock::now();
        int time_index = 0;
        for (int n = 0; n < cfg.steps; ++n) {
            if (n % cfg.out_every == 0 || n == 0) {  // main.cpp:96
                if (frames && time_index < max_frames) dehalo(frames + static_cast<size_t>(time_index) * gsz);
                time_index++;
            }
            exchange_halos(u, dec, MPI_COMM_WORLD);                   // main.cpp:101
            apply_boundary(u, dec, cfg.bc, 0.0);                      // main.cpp:102
            std::copy(u.data.begin(), u.data.end(), tmp.data.begin());  // main.cpp:104
            diffusion_step(u, tmp, cfg.D, cfg.dt);                    // main.cpp:106
            advection_step(u, tmp, cfg.vx, cfg.vy, cfg.dt);           // main.cpp:107
            std::swap(u.data, tmp.data);                              // main.cpp:109
        }
        mini_mpi_barrier();
        const auto t1 = std::chrono::steady_clock::now();
        secs[static_cast<size_t>(rank)] = std::chrono::duration<double>(t1 - t0).count();
        nfr[static_cast<size_t>(rank)] = std::min(time_index, max_frames);
        if (final_interior) dehalo(final_interior);
        if (final_padded_rank0 && rank == 0)
            std::copy(u.data.begin(), u.data.end(), final_padded_rank0);
        mini_mpi_barrier();
        dec.finalize();
    };

    std::vector<std::thread> th;
    for (int r = 0; r < nranks; ++r)
        th.emplace_back([&, r] {
            try {
                body(r);
            } catch (const std::exception& e) {
                errs[static_cast<size_t>(r)] = e.what();
            }
        });
    for (auto& t : th) t.join();
    for (auto& e : errs)
        if (!e.empty()) {
            t_err = e;
            return -1;
        }
    if (loop_seconds) *loop_seconds = *std::max_element(secs.begin(), secs.end());
    return frames ? nfr[0] : 0;
}

}  // extern "C"
