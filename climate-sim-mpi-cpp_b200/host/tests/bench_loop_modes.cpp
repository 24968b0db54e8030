// bench_loop_modes.cpp — what the three ways of driving the time loop through the drop-in headers cost
// (INTEGRATION.md §1 quotes the output):
//   (a) src/main.cpp:101-109 verbatim — five calls per step plus std::copy(u.data → tmp.data), which goes
//       through the host mirror: the tile crosses PCIe every step;
//   (b) the same with the copy statement removed (the GPU diffusion_step seeds `out` itself): five GPU calls
//       per step, nothing crosses PCIe, but one step per sweep and two passes;
//   (c) run_timesteps(): the fused, temporally blocked sweep.
// usage: bench_loop_modes [n = 4096] [steps = 20]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "csim_driver.hpp"

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 4096, steps = argc > 2 ? std::atoi(argv[2]) : 20;
    MPI_Init(&argc, &argv);
    SimConfig cfg;
    cfg.nx = cfg.ny = n;
    cfg.D = 0.05;
    cfg.vx = 0.5;
    cfg.vy = 0.25;
    cfg.dt = 0.1;
    Decomp2D dec;
    dec.init(MPI_COMM_WORLD, n, n);
    double t[3] = {0, 0, 0};
    double sum[3] = {0, 0, 0};
    for (int mode = 0; mode < 3; ++mode) {
        Field u(n, n, 1, 1.0, 1.0), tmp(n, n, 1, 1.0, 1.0);
        apply_initial_condition(dec, u, cfg);
        csim_sync(csim_host::default_context());
        const int warm = 2;
        double t0 = 0;
        for (int s = 0; s < steps + warm; ++s) {
            if (s == warm) {
                csim_sync(csim_host::default_context());
                t0 = now();
            }
            if (mode == 2) {
                run_timesteps(u, tmp, dec, cfg.bc, cfg.D, cfg.vx, cfg.vy, cfg.dt, 1);
                continue;
            }
            exchange_halos(u, dec, MPI_COMM_WORLD);                                // main.cpp:101
            apply_boundary(u, dec, cfg.bc, 0.0);                                   // :102
            if (mode == 0) std::copy(u.data.begin(), u.data.end(), tmp.data.begin());  // :104
            diffusion_step(u, tmp, cfg.D, cfg.dt);                                 // :106
            advection_step(u, tmp, cfg.vx, cfg.vy, cfg.dt);                        // :107
            std::swap(u.data, tmp.data);                                           // :109
        }
        csim_sync(csim_host::default_context());
        t[mode] = (now() - t0) / steps;
        const auto mm = field_minmax(u);
        sum[mode] = mm.second;
    }
    // mode 2 with the steps in one call (what the driver does): blocked in time
    {
        Field u(n, n, 1, 1.0, 1.0), tmp(n, n, 1, 1.0, 1.0);
        apply_initial_condition(dec, u, cfg);
        run_timesteps(u, tmp, dec, cfg.bc, cfg.D, cfg.vx, cfg.vy, cfg.dt, 4);
        csim_sync(csim_host::default_context());
        const double t0 = now();
        run_timesteps(u, tmp, dec, cfg.bc, cfg.D, cfg.vx, cfg.vy, cfg.dt, 100);
        csim_sync(csim_host::default_context());
        const double per = (now() - t0) / 100;
        std::printf("loop modes at %dx%d (ms per time step): main.cpp:101-109 verbatim %.3f | without the std::copy "
                    "statement %.3f | run_timesteps one step per call %.3f | run_timesteps 100 steps per call %.4f\n",
                    n, n, 1e3 * t[0], 1e3 * t[1], 1e3 * t[2], 1e3 * per);
    }
    std::printf("max after %d steps: %.17g %.17g %.17g (must agree)\n", steps + 2, sum[0], sum[1], sum[2]);
    MPI_Finalize();
    return (sum[0] == sum[1] && sum[1] == sum[2]) ? 0 : 1;
}
