// test_dropin.cpp — the reference's unit tests for the timestep path, restated against the drop-in
// headers (no googletest in this image: plain checks, exit code 0 = all passed).  Needs a GPU.
//   Unit_Field      tests/simulation/unit/test_field.cpp:5-26
//   Unit_Diffusion  tests/simulation/unit/test_diffusion.cpp:17-34
//   Unit_Advection  tests/simulation/unit/test_advection.cpp:13-71
//   Unit_Boundary   tests/simulation/unit/test_boundary.cpp:9-69
//   Unit_Stability  tests/simulation/unit/test_stability.cpp:5-27
//   Unit_Decomp     tests/simulation/unit/test_decomp_mpi.cpp:5-36 (one rank)
// plus the loop body of src/main.cpp:101-109 written with the reference's five statements against
// run_timesteps (the fused path): both must give the same bits.
#include <mpi.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "advection.hpp"
#include "boundary.hpp"
#include "decomp.hpp"
#include "diffusion.hpp"
#include "field.hpp"
#include "halo.hpp"
#include "stability.hpp"

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                               \
    do {                                                                          \
        ++g_checks;                                                               \
        if (!(cond)) {                                                            \
            ++g_fail;                                                             \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);           \
        }                                                                         \
    } while (0)
#define CHECK_NEAR(a, b, tol) CHECK(std::fabs((a) - (b)) <= (tol))
template <class Ex, class F>
static bool throws(F f) {
    try {
        f();
    } catch (const Ex&) {
        return true;
    } catch (...) {
    }
    return false;
}

static void unit_field() {
    Field f(4, 3, 1, 1.0, 1.0);
    CHECK(f.data.size() == static_cast<size_t>(f.nx_total() * f.ny_total()));
    Field g(2, 2, 1, 1.0, 1.0);
    for (int j = 0; j < g.ny_total(); ++j)
        for (int i = 0; i < g.nx_total(); ++i) g.at(i, j) = 10 * j + i;
    CHECK(g.at(0, 0) == 0 && g.at(g.nx_total() - 1, 0) == 3 && g.at(0, 1) == 10 && g.at(3, 3) == 33);
    Field h(4, 4, 1, 1.0, 1.0);
    CHECK(throws<std::out_of_range>([&] { (void)h.at(-1, 0); }));
    CHECK(throws<std::out_of_range>([&] { (void)h.at(h.nx_total(), 0); }));
    CHECK(throws<std::out_of_range>([&] { (void)h.at(0, h.ny_total()); }));
    Field z(2, 2, 0, 1.0, 1.0);  // halo 0 constructs (test_io.cpp:174)
    CHECK(z.data.size() == 4);
}

static void unit_diffusion() {
    Field u(3, 3, 1, 1.0, 1.0), v(3, 3, 1, 1.0, 1.0);
    u.at(2, 2) = 1.0;
    for (int i = 0; i < u.nx_total(); ++i) u.at(i, 0) = u.at(i, u.ny_total() - 1) = 0.0;
    for (int j = 0; j < u.ny_total(); ++j) u.at(0, j) = u.at(u.nx_total() - 1, j) = 0.0;
    const double D = 0.1, dt = 0.1, alpha = D * dt / (u.dx * u.dx);
    diffusion_step(u, v, D, dt);
    CHECK_NEAR(v.at(2, 2), 1.0 - 4 * alpha, 1e-12);
    CHECK_NEAR(v.at(1, 2), alpha, 1e-12);
    CHECK_NEAR(v.at(3, 2), alpha, 1e-12);
    CHECK_NEAR(v.at(2, 1), alpha, 1e-12);
    CHECK_NEAR(v.at(2, 3), alpha, 1e-12);
}

static Field make_hotspot(int nx, int ny, int halo = 1) {
    Field f(nx, ny, halo, 1.0, 1.0);
    f.fill(0.0);
    f.at(nx / 2 + halo, ny / 2 + halo) = 1.0;
    return f;
}
static void unit_advection() {
    const int nx = 8, ny = 8;
    Field u = make_hotspot(nx, ny);
    {
        Field out(nx, ny, 1, 1.0, 1.0);
        out.fill(0.0);
        advection_step(u, out, 0.0, 0.0, 0.1);
        bool all_zero = true;
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) all_zero = all_zero && out.at(i, j) == 0.0;
        CHECK(all_zero);
    }
    const double v[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
    for (auto& w : v) {
        Field out(nx, ny, 1, 1.0, 1.0);
        out.fill(0.0);
        advection_step(u, out, w[0], w[1], 0.1);
        CHECK(out.at(nx / 2 + 1, ny / 2 + 1) != 0.0);
    }
}

static void unit_boundary() {
    int init = 0;
    MPI_Initialized(&init);
    if (!init) {
        int prov = 0;
        MPI_Init_thread(nullptr, nullptr, MPI_THREAD_FUNNELED, &prov);
    }
    int size = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    CHECK(size == 1);
    const int NX = 4, NY = 3, h = 1;
    Decomp2D dec;
    dec.init(MPI_COMM_WORLD, NX, NY);
    Field f(NX, NY, h, 1.0, 1.0);
    f.fill(-1.0);
    for (int j = h; j < h + NY; ++j)
        for (int i = h; i < h + NX; ++i) f.at(i, j) = 10.0;
    BCConfig bc_dir;
    apply_boundary(f, dec, bc_dir, 5.0);
    bool ok = true;
    for (int j = 0; j < f.ny_total(); ++j) ok = ok && f.at(0, j) == 5.0 && f.at(h + NX, j) == 5.0;
    for (int i = 0; i < f.nx_total(); ++i) ok = ok && f.at(i, 0) == 5.0 && f.at(i, h + NY) == 5.0;
    CHECK(ok);
    f.fill(-1.0);
    for (int j = h; j < h + NY; ++j)
        for (int i = h; i < h + NX; ++i) f.at(i, j) = static_cast<double>(j);
    BCConfig bc_neu;
    bc_neu.left = bc_neu.right = bc_neu.bottom = bc_neu.top = BCType::Neumann;
    apply_boundary(f, dec, bc_neu, 0.0);
    ok = true;
    for (int j = 0; j < f.ny_total(); ++j)
        ok = ok && f.at(0, j) == f.at(h, j) && f.at(h + NX, j) == f.at(h + NX - 1, j);
    for (int i = 0; i < f.nx_total(); ++i)
        ok = ok && f.at(i, 0) == f.at(i, h) && f.at(i, h + NY) == f.at(i, h + NY - 1);
    CHECK(ok);
    dec.finalize();
}

static void unit_stability() {
    CHECK(safe_dt(1, 1, 0.5, 0.5, 0.1) > 0.0);
    CHECK(safe_dt(1, 1, 5, 5, 0.1) < safe_dt(1, 1, 0.5, 0.5, 0.1));
    CHECK(safe_dt(1, 1, 0.5, 0.5, 1.0) < safe_dt(1, 1, 0.5, 0.5, 0.1));
    CHECK(safe_dt(1, 1, 0.5, 0.0, 0.05) == 2.0);
}

static void unit_decomp() {
    Decomp2D d;
    d.init(MPI_COMM_WORLD, 16, 12);
    CHECK(d.dims[0] * d.dims[1] == 1 && d.coords[0] == 0 && d.coords[1] == 0);
    CHECK(d.nbr_lr[0] == MPI_PROC_NULL && d.nbr_lr[1] == MPI_PROC_NULL);
    CHECK(d.nx_local == 16 && d.ny_local == 12 && d.x_offset == 0 && d.y_offset == 0);
    d.finalize();
    CHECK(d.cart_comm == MPI_COMM_NULL);
}

// src/main.cpp:101-109 written exactly as the reference has it, against the fused entry point
static void loop_body_equals_fused() {
    const int nx = 301, ny = 217, steps = 13;
    Decomp2D dec;
    dec.init(MPI_COMM_WORLD, nx, ny);
    BCConfig bc;
    bc.left = BCType::Dirichlet;
    bc.right = BCType::Neumann;
    bc.bottom = BCType::Periodic;
    bc.top = BCType::Dirichlet;
    const double D = 0.05, vx = -0.5, vy = 0.25, dt = 0.1;
    Field u(nx, ny, 1, 1.0, 1.0), tmp(nx, ny, 1, 1.0, 1.0), a(nx, ny, 1, 1.0, 1.0), b(nx, ny, 1, 1.0, 1.0);
    unsigned s = 12345u;
    for (auto& x : u.data) {
        s = s * 1664525u + 1013904223u;
        x = (static_cast<double>(s >> 8) / (1u << 24)) - 0.5;
    }
    std::copy(u.data.begin(), u.data.end(), a.data.begin());
    for (int n = 0; n < steps; ++n) {
        exchange_halos(u, dec, MPI_COMM_WORLD);
        apply_boundary(u, dec, bc, 0.0);
        std::copy(u.data.begin(), u.data.end(), tmp.data.begin());
        diffusion_step(u, tmp, D, dt);
        advection_step(u, tmp, vx, vy, dt);
        std::swap(u.data, tmp.data);
    }
    run_timesteps(a, b, dec, bc, D, vx, vy, dt, steps);
    bool same = true;
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            const double p = u.at(i, j), q = a.at(i, j);
            same = same && std::memcmp(&p, &q, sizeof p) == 0;
        }
    CHECK(same);
    const double mn = *std::min_element(a.data.begin(), a.data.end());  // main.cpp:74
    CHECK(std::isfinite(mn));
}

int main() {
    unit_field();
    unit_diffusion();
    unit_advection();
    unit_boundary();
    unit_stability();
    unit_decomp();
    loop_body_equals_fused();
    int fin = 0;
    MPI_Finalized(&fin);
    if (!fin) MPI_Finalize();
    std::printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "ALL PASS", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
