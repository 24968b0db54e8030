// host_misc.cpp — host-only parts of the path: decomposition, stability limit, initial condition.
// No device work happens here; these mirror scalar/host logic of the reference.
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <fstream>
#include <string>
#include <limits>
#include <thread>
#include <vector>

#include "csim_internal.hpp"

using namespace csim;

extern "C" {

// safe_dt — include/stability.hpp:5-16
double csim_safe_dt(double dx, double dy, double vx, double vy, double D) {
    const double inf = std::numeric_limits<double>::infinity();
    const double ax = std::abs(vx), ay = std::abs(vy);
    const double denom_adv = (ax > 0 ? ax / dx : 0.0) + (ay > 0 ? ay / dy : 0.0);
    const double dt_adv = denom_adv > 0 ? 1.0 / denom_adv : inf;
    const double denom_diff = 1.0 / (dx * dx) + 1.0 / (dy * dy);
    const double dt_diff = D > 0 ? 1.0 / (2.0 * D * denom_diff) : inf;
    return std::min(dt_adv, dt_diff);
}

// Decomp2D::init — src/decomp.cpp:5-34.  MPI_Dims_create(size, 2, dims) yields the most square
// factorisation in non-increasing order; MPI_Cart_create(reorder=0, periods={0,0}) keeps row-major
// rank order, so coords = (rank / dims[1], rank % dims[1]) and off-grid neighbours are PROC_NULL.
int csim_decomp_init(int size, int rank, int nxg, int nyg, csim_decomp* d) {
    CSIM_REQUIRE(d != nullptr, CSIM_ERR_INVALID, "csim_decomp_init: out is null");
    CSIM_REQUIRE(size >= 1 && rank >= 0 && rank < size, CSIM_ERR_INVALID, "csim_decomp_init: bad size/rank");
    int small = 1;
    for (int f = 1; static_cast<long long>(f) * f <= size; ++f)
        if (size % f == 0) small = f;
    d->dims[0] = size / small;
    d->dims[1] = small;
    const int cx = rank / d->dims[1], cy = rank % d->dims[1];
    d->coords[0] = cx;
    d->coords[1] = cy;
    auto rank_of = [&](int x, int y) {
        return (x < 0 || y < 0 || x >= d->dims[0] || y >= d->dims[1]) ? CSIM_PROC_NULL : x * d->dims[1] + y;
    };
    d->nbr[CSIM_LEFT] = rank_of(cx - 1, cy);    // nbr_lr[0], decomp.cpp:21
    d->nbr[CSIM_RIGHT] = rank_of(cx + 1, cy);   // nbr_lr[1]
    d->nbr[CSIM_BOTTOM] = rank_of(cx, cy - 1);  // nbr_du[0], decomp.cpp:22
    d->nbr[CSIM_TOP] = rank_of(cx, cy + 1);     // nbr_du[1]
    d->nx_global = nxg;
    d->ny_global = nyg;
    const int bx = nxg / d->dims[0], by = nyg / d->dims[1];
    d->nx_local = bx + (cx == d->dims[0] - 1 ? nxg % d->dims[0] : 0);  // decomp.cpp:29
    d->ny_local = by + (cy == d->dims[1] - 1 ? nyg % d->dims[1] : 0);  // decomp.cpp:30
    d->x_offset = cx * bx;                                             // decomp.cpp:32
    d->y_offset = cy * by;                                             // decomp.cpp:33
    return CSIM_OK;
}

// Geometry of the wide exchange (used by halo.cu and, on the CPU, by the gloo tests).
int csim_wide_exchange_plan(const csim_decomp* dec, int T, csim_xregion snd[8], csim_xregion rcv[8]) {
    CSIM_REQUIRE(dec != nullptr && snd != nullptr && rcv != nullptr, CSIM_ERR_INVALID,
                 "csim_wide_exchange_plan: null argument");
    CSIM_REQUIRE(T >= 1 && T <= kMaxHalo, CSIM_ERR_INVALID, "csim_wide_exchange_plan: T out of range");
    const int nx = dec->nx_local, ny = dec->ny_local;
    const int cx = dec->coords[0], cy = dec->coords[1];
    auto rank_of = [&](int x, int y) {
        return (x < 0 || y < 0 || x >= dec->dims[0] || y >= dec->dims[1]) ? -1 : x * dec->dims[1] + y;
    };
    const bool pl = dec->nbr[CSIM_LEFT] == CSIM_PROC_NULL, pr = dec->nbr[CSIM_RIGHT] == CSIM_PROC_NULL;
    const bool pb = dec->nbr[CSIM_BOTTOM] == CSIM_PROC_NULL, pt = dec->nbr[CSIM_TOP] == CSIM_PROC_NULL;
    // extent of a band along its own side: interior plus the ghost line of a physical end
    const int bx0 = pl ? -1 : 0, bx1 = pr ? nx + 1 : nx;
    const int by0 = pb ? -1 : 0, by1 = pt ? ny + 1 : ny;
    int k = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            csim_xregion s, r;
            s.peer = r.peer = rank_of(cx + dx, cy + dy);
            // what goes towards (dx,dy): own cells next to that side
            s.x0 = dx < 0 ? 0 : (dx > 0 ? nx - T : bx0);
            s.w = dx != 0 ? T : bx1 - bx0;
            s.y0 = dy < 0 ? 0 : (dy > 0 ? ny - T : by0);
            s.h = dy != 0 ? T : by1 - by0;
            // where what comes from (dx,dy) lands: the ghost area on that side
            r.x0 = dx < 0 ? -T : (dx > 0 ? nx : bx0);
            r.w = s.w;
            r.y0 = dy < 0 ? -T : (dy > 0 ? ny : by0);
            r.h = s.h;
            snd[k] = s;
            rcv[k] = r;
            ++k;
        }
    return CSIM_OK;
}

// apply_initial_condition — src/init.cpp:12-47.  Host libm exp() keeps the tile bit-identical to
// the reference's; rows are split over host threads (each cell is independent).
int csim_initial_condition_host(double* host, const csim_decomp* dec, int halo, int nxg, int nyg, double dx,
                                double dy, int preset, double A, double sigma_frac, double xc_frac,
                                double yc_frac) {
    CSIM_REQUIRE(host != nullptr && dec != nullptr, CSIM_ERR_INVALID, "csim_initial_condition_host: null argument");
    CSIM_REQUIRE(preset == 0 || preset == 1, CSIM_ERR_INVALID, "Unknown IC preset");  // init.cpp:42
    if (preset == 1) return CSIM_OK;  // constant_zero: no-op, init.cpp:39-40
    const int nx = dec->nx_local, ny = dec->ny_local, h = halo;
    const int64_t nxt = nx + 2 * h;
    const double Lx = nxg * dx, Ly = nyg * dy;
    const double xc = xc_frac * Lx, yc = yc_frac * Ly;
    const double sig = sigma_frac * std::min(Lx, Ly);
    auto rows = [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            const int gj = dec->y_offset + j;
            const double y = (gj + 0.5) * dy;
            double* row = host + static_cast<int64_t>(j + h) * nxt + h;
            for (int i = 0; i < nx; ++i) {
                const int gi = dec->x_offset + i;
                const double x = (gi + 0.5) * dx;
                const double r2 = (x - xc) * (x - xc) + (y - yc) * (y - yc);
                row[i] = A * std::exp(-r2 / (2.0 * sig * sig));
            }
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 32) nt = 32;
    if (static_cast<int64_t>(nx) * ny < (1 << 16)) nt = 1;
    std::vector<std::thread> pool;
    const int chunk = (ny + static_cast<int>(nt) - 1) / static_cast<int>(nt);
    for (unsigned t = 0; t < nt; ++t) {
        const int j0 = static_cast<int>(t) * chunk, j1 = std::min(ny, j0 + chunk);
        if (j0 >= j1) break;
        pool.emplace_back(rows, j0, j1);
    }
    for (auto& th : pool) th.join();
    return CSIM_OK;
}

}  // extern "C"

extern "C" int csim_bind_thread_to_device_numa(int device, int* node) {
    if (node) *node = -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return csim::fail(CSIM_ERR_CUDA, "csim_bind_thread_to_device_numa: bad device index");
    }
    std::string id(bus);
    for (char& ch : id) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
    int n = -1;
    {
        std::ifstream in("/sys/bus/pci/devices/" + id + "/numa_node");
        if (!(in >> n)) n = -1;
    }
    if (n < 0) return CSIM_OK;  // single-node box or no affinity information
    std::ifstream in("/sys/devices/system/node/node" + std::to_string(n) + "/cpulist");
    std::string list;
    if (!std::getline(in, list) || list.empty()) return CSIM_OK;
    cpu_set_t set;
    CPU_ZERO(&set);
    int count = 0;
    size_t pos = 0;  // "0-31,64-95"
    while (pos < list.size()) {
        size_t end = list.find(',', pos);
        if (end == std::string::npos) end = list.size();
        const std::string part = list.substr(pos, end - pos);
        const size_t dash = part.find('-');
        try {
            const int lo = std::stoi(part.substr(0, dash));
            const int hi = dash == std::string::npos ? lo : std::stoi(part.substr(dash + 1));
            for (int cpu = lo; cpu <= hi && cpu < CPU_SETSIZE; ++cpu) {
                CPU_SET(cpu, &set);
                ++count;
            }
        } catch (...) {
            return CSIM_OK;
        }
        pos = end + 1;
    }
    // keep only CPUs this process may use at all (cgroup / taskset limits)
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0) {
        count = 0;
        for (int cpu = 0; cpu < CPU_SETSIZE; ++cpu) {
            if (CPU_ISSET(cpu, &set) && !CPU_ISSET(cpu, &allowed)) CPU_CLR(cpu, &set);
            if (CPU_ISSET(cpu, &set)) ++count;
        }
    }
    if (count == 0) return CSIM_OK;
    if (sched_setaffinity(0, sizeof set, &set) == 0 && node) *node = n;
    return CSIM_OK;
}

