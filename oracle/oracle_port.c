/*
 * oracle_port.c — CPU restatement of the climate-sim-mpi-cpp timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker.  The product path (climate-sim-mpi-cpp_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port bit-for-bit against
 *   (1) the reference's own field/diffusion/advection/boundary .cpp objects compiled unmodified
 *       into oracle/_ref/libcsim_ref.so (see oracle/Makefile, oracle/ref_harness.cpp), and
 *   (2) every exact known answer the reference's unit tests hold for this path
 *       (tests/simulation/unit/test_{field,diffusion,advection,boundary,halo,stability,decomp_mpi}.cpp),
 *   (3) golden vectors in tests/golden/ generated from (1).
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * repository root).  Build: gcc -O2 -ffp-contract=off, no -march (the reference build has no
 * arch flags, hence no FMA contraction; see CMakeLists.txt:1-18).
 *
 * Layout (include/field.hpp:5-21, src/field.cpp:20-25): a field tile is a row-major array of
 * (ny+2h) rows by (nx+2h) columns of doubles, idx(i,j) = j*(nx+2h) + i, i (x) fastest, interior
 * at [h, h+n) in both directions.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PROC_NULL (-1) /* stands for MPI_PROC_NULL */

enum { ORC_BC_DIRICHLET = 0, ORC_BC_NEUMANN = 1, ORC_BC_PERIODIC = 2 }; /* include/boundary.hpp:5 */

typedef struct {
    int nx, ny, h;
    double dx, dy;
    double* data; /* (ny+2h)*(nx+2h) */
} orc_field;

/* include/decomp.hpp:4-17 */
typedef struct {
    int dims[2];
    int coords[2];
    int nbr_lr[2];
    int nbr_du[2];
    int nx_global, ny_global;
    int nx_local, ny_local;
    int x_offset, y_offset;
} orc_decomp;

/* include/io.hpp:10-39 (the members the hot path reads) */
typedef struct {
    int nx, ny;
    double dx, dy;
    double D, vx, vy;
    double dt;
    int steps, out_every;
    int bc[4]; /* left, right, bottom, top */
    int ic_preset; /* 0 = gaussian_hotspot, 1 = constant_zero */
    double A, sigma_frac, xc_frac, yc_frac;
} orc_config;

static inline int nx_tot(const orc_field* f) { return f->nx + 2 * f->h; } /* field.hpp:17 */
static inline int ny_tot(const orc_field* f) { return f->ny + 2 * f->h; } /* field.hpp:18 */
static inline size_t fidx(const orc_field* f, int i, int j) {             /* field.cpp:20-25 */
    return (size_t)j * (size_t)nx_tot(f) + (size_t)i;
}

/* src/field.cpp:6-12 — allocation, zero-initialised. */
int orc_field_init(orc_field* f, int nx, int ny, int h, double dx, double dy) {
    f->nx = nx;
    f->ny = ny;
    f->h = h;
    f->dx = dx;
    f->dy = dy;
    f->data = (double*)calloc((size_t)(nx + 2 * h) * (size_t)(ny + 2 * h), sizeof(double));
    return f->data ? 0 : -1;
}
void orc_field_free(orc_field* f) {
    free(f->data);
    f->data = NULL;
}

/* src/diffusion.cpp:3-26.  Operation order exactly as diffusion.cpp:12-14:
 *   lap = ((e - 2.0*c) + w) / (dx*dx) + ((n - 2.0*c) + s) / (dy*dy);  out = c + (dt*D)*lap
 * then the outermost ring of `out` := that of `u` (diffusion.cpp:18-25). */
void orc_diffusion_step(const double* u, double* out, int nx, int ny, int h, double dx, double dy,
                        double D, double dt) {
    const int nxt = nx + 2 * h, nyt = ny + 2 * h;
    for (int j = h; j < ny + h; ++j) {
        for (int i = h; i < nx + h; ++i) {
            const size_t k = (size_t)j * nxt + i;
            const double uij = u[k];
            const double lap = (u[k + 1] - 2.0 * uij + u[k - 1]) / (dx * dx) +
                               (u[k + nxt] - 2.0 * uij + u[k - nxt]) / (dy * dy);
            out[k] = uij + dt * D * lap;
        }
    }
    for (int i = 0; i < nxt; ++i) {
        out[i] = u[i];
        out[(size_t)(nyt - 1) * nxt + i] = u[(size_t)(nyt - 1) * nxt + i];
    }
    for (int j = 0; j < nyt; ++j) {
        out[(size_t)j * nxt] = u[(size_t)j * nxt];
        out[(size_t)j * nxt + nxt - 1] = u[(size_t)j * nxt + nxt - 1];
    }
}

/* src/advection.cpp:5-34.  First-order upwind; the side is chosen by vx>=0 / vy>=0 (0 counts as
 * >=, advection.cpp:16,23); ACCUMULATES into out (advection.cpp:31). */
void orc_advection_step(const double* u, double* out, int nx, int ny, int h, double dx, double dy,
                        double vx, double vy, double dt) {
    const int nxt = nx + 2 * h;
    for (int j = h; j < h + ny; ++j) {
        for (int i = h; i < h + nx; ++i) {
            const size_t k = (size_t)j * nxt + i;
            double dudx, dudy;
            if (vx >= 0.0)
                dudx = (u[k] - u[k - 1]) / dx;
            else
                dudx = (u[k + 1] - u[k]) / dx;
            if (vy >= 0.0)
                dudy = (u[k] - u[k - nxt]) / dy;
            else
                dudy = (u[k + nxt] - u[k]) / dy;
            const double adv = vx * dudx + vy * dudy;
            out[k] += (-dt) * adv;
        }
    }
}

/* src/boundary.cpp:12-54.  nbr = {left,right,down,up} neighbour ranks; a side is physical when
 * its neighbour is PROC_NULL.  Order left, right, bottom, top; columns span j=0..h+ny, rows
 * i=0..nx_tot-1; Periodic does nothing (no branch exists for it). */
void orc_apply_boundary(double* f, int nx, int ny, int h, const int nbr[4], const int bc[4],
                        double value) {
    const int nxt = nx + 2 * h;
    const int iL = 0, iR = h + nx, jB = 0, jT = h + ny, i0 = 0, i1 = nxt - 1;
    if (nbr[0] == ORC_PROC_NULL) {
        if (bc[0] == ORC_BC_DIRICHLET)
            for (int j = jB; j <= jT; ++j) f[(size_t)j * nxt + iL] = value;
        else if (bc[0] == ORC_BC_NEUMANN)
            for (int j = jB; j <= jT; ++j) f[(size_t)j * nxt + iL] = f[(size_t)j * nxt + h];
    }
    if (nbr[1] == ORC_PROC_NULL) {
        if (bc[1] == ORC_BC_DIRICHLET)
            for (int j = jB; j <= jT; ++j) f[(size_t)j * nxt + iR] = value;
        else if (bc[1] == ORC_BC_NEUMANN)
            for (int j = jB; j <= jT; ++j) f[(size_t)j * nxt + iR] = f[(size_t)j * nxt + h + nx - 1];
    }
    if (nbr[2] == ORC_PROC_NULL) {
        if (bc[2] == ORC_BC_DIRICHLET)
            for (int i = i0; i <= i1; ++i) f[(size_t)jB * nxt + i] = value;
        else if (bc[2] == ORC_BC_NEUMANN)
            for (int i = i0; i <= i1; ++i) f[(size_t)jB * nxt + i] = f[(size_t)h * nxt + i];
    }
    if (nbr[3] == ORC_PROC_NULL) {
        if (bc[3] == ORC_BC_DIRICHLET)
            for (int i = i0; i <= i1; ++i) f[(size_t)jT * nxt + i] = value;
        else if (bc[3] == ORC_BC_NEUMANN)
            for (int i = i0; i <= i1; ++i) f[(size_t)jT * nxt + i] = f[(size_t)(h + ny - 1) * nxt + i];
    }
}

/* include/stability.hpp:5-16 */
double orc_safe_dt(double dx, double dy, double vx, double vy, double D) {
    const double denom_adv =
        (fabs(vx) > 0 ? fabs(vx) / dx : 0.0) + (fabs(vy) > 0 ? fabs(vy) / dy : 0.0);
    const double dt_adv = (denom_adv > 0) ? (1.0 / denom_adv) : INFINITY;
    const double denom_diff = (1.0 / (dx * dx)) + (1.0 / (dy * dy));
    const double dt_diff = (D > 0) ? (1.0 / (2.0 * D * denom_diff)) : INFINITY;
    return dt_adv < dt_diff ? dt_adv : dt_diff;
}

/* MPI_Dims_create(size, 2, dims) as src/decomp.cpp:13 uses it: the most-square factorisation in
 * non-increasing order (1→{1,1}, 2→{2,1}, 4→{2,2}, 6→{3,2}, 8→{4,2}). */
void orc_dims_create(int size, int dims[2]) {
    int b = 1;
    for (int d = 1; (long)d * d <= size; ++d)
        if (size % d == 0) b = d;
    dims[0] = size / b;
    dims[1] = b;
}

/* src/decomp.cpp:5-34 with MPI_Cart_create(periods={0,0}, reorder=0): row-major rank order,
 * coords = (r / dims[1], r % dims[1]); dims[0] splits x, dims[1] splits y; the last rank of a
 * dimension absorbs the remainder (decomp.cpp:29-30); offsets use the base size (:32-33). */
void orc_decomp_init(orc_decomp* d, int size, int rank, int nxg, int nyg) {
    d->nx_global = nxg;
    d->ny_global = nyg;
    orc_dims_create(size, d->dims);
    d->coords[0] = rank / d->dims[1];
    d->coords[1] = rank % d->dims[1];
    /* MPI_Cart_shift(dim 0, +1) → left/right; (dim 1, +1) → down/up (decomp.cpp:21-22) */
    d->nbr_lr[0] = d->coords[0] > 0 ? (d->coords[0] - 1) * d->dims[1] + d->coords[1] : ORC_PROC_NULL;
    d->nbr_lr[1] =
        d->coords[0] < d->dims[0] - 1 ? (d->coords[0] + 1) * d->dims[1] + d->coords[1] : ORC_PROC_NULL;
    d->nbr_du[0] = d->coords[1] > 0 ? d->coords[0] * d->dims[1] + d->coords[1] - 1 : ORC_PROC_NULL;
    d->nbr_du[1] =
        d->coords[1] < d->dims[1] - 1 ? d->coords[0] * d->dims[1] + d->coords[1] + 1 : ORC_PROC_NULL;
    const int base_nx = nxg / d->dims[0], base_ny = nyg / d->dims[1];
    const int rem_x = nxg % d->dims[0], rem_y = nyg % d->dims[1];
    d->nx_local = base_nx + (d->coords[0] == d->dims[0] - 1 ? rem_x : 0);
    d->ny_local = base_ny + (d->coords[1] == d->dims[1] - 1 ? rem_y : 0);
    d->x_offset = d->coords[0] * base_nx;
    d->y_offset = d->coords[1] * base_ny;
}

/* src/init.cpp:12-33 (gaussian_hotspot) and :35-47 (preset dispatch; constant_zero is a no-op). */
void orc_apply_ic(const orc_decomp* dec, orc_field* u, const orc_config* cfg) {
    if (cfg->ic_preset != 0) return;
    const int h = u->h, nx = u->nx, ny = u->ny;
    const double Lx = cfg->nx * cfg->dx, Ly = cfg->ny * cfg->dy;
    const double xc = cfg->xc_frac * Lx, yc = cfg->yc_frac * Ly;
    const double sig = cfg->sigma_frac * (Lx < Ly ? Lx : Ly);
    for (int j = h; j < h + ny; ++j) {
        const int gj = dec->y_offset + (j - h);
        const double y = (gj + 0.5) * cfg->dy;
        for (int i = h; i < h + nx; ++i) {
            const int gi = dec->x_offset + (i - h);
            const double x = (gi + 0.5) * cfg->dx;
            const double r2 = (x - xc) * (x - xc) + (y - yc) * (y - yc);
            u->data[fidx(u, i, j)] = cfg->A * exp(-r2 / (2.0 * sig * sig));
        }
    }
}

/* src/halo.cpp:28-43, all ranks at once.  Each rank's ghost lines receive the neighbour's adjacent
 * interior lines: columns over j in [h, h+ny) only (colType, halo.cpp:12-14), rows over all nx_tot
 * cells including ghost columns (rowType, halo.cpp:16-18).  MPI delivers all eight transfers
 * concurrently, so the row payload's ghost-column cells (the corner ghosts) are indeterminate in
 * the reference; here every send buffer is snapshotted before any receive lands. */
void orc_exchange_halos(orc_field* f, const orc_decomp* dec, int nranks) {
    /* snapshot the four send lines of every rank */
    double** snap = (double**)malloc(sizeof(double*) * 4 * (size_t)nranks);
    for (int r = 0; r < nranks; ++r) {
        const orc_field* t = &f[r];
        const int h = t->h, nx = t->nx, ny = t->ny, nxt = nx_tot(t);
        double* L = (double*)malloc(sizeof(double) * (size_t)ny);
        double* R = (double*)malloc(sizeof(double) * (size_t)ny);
        double* B = (double*)malloc(sizeof(double) * (size_t)nxt);
        double* T = (double*)malloc(sizeof(double) * (size_t)nxt);
        for (int j = 0; j < ny; ++j) {
            L[j] = t->data[fidx(t, h, h + j)];          /* Isend &f.at(h,h)       halo.cpp:30 */
            R[j] = t->data[fidx(t, h + nx - 1, h + j)]; /* Isend &f.at(h+nx-1,h)  halo.cpp:34 */
        }
        memcpy(B, &t->data[fidx(t, 0, h)], sizeof(double) * (size_t)nxt);          /* halo.cpp:38 */
        memcpy(T, &t->data[fidx(t, 0, h + ny - 1)], sizeof(double) * (size_t)nxt); /* halo.cpp:42 */
        snap[4 * r + 0] = L;
        snap[4 * r + 1] = R;
        snap[4 * r + 2] = B;
        snap[4 * r + 3] = T;
    }
    for (int r = 0; r < nranks; ++r) {
        orc_field* t = &f[r];
        const int h = t->h, nx = t->nx, ny = t->ny, nxt = nx_tot(t);
        const int left = dec[r].nbr_lr[0], right = dec[r].nbr_lr[1];
        const int down = dec[r].nbr_du[0], up = dec[r].nbr_du[1];
        if (left != ORC_PROC_NULL) /* Irecv &f.at(0,h) ← left's right interior column  :29 */
            for (int j = 0; j < ny; ++j) t->data[fidx(t, 0, h + j)] = snap[4 * left + 1][j];
        if (right != ORC_PROC_NULL) /* Irecv &f.at(h+nx,h) ← right's left interior column  :33 */
            for (int j = 0; j < ny; ++j) t->data[fidx(t, h + nx, h + j)] = snap[4 * right + 0][j];
        if (down != ORC_PROC_NULL) /* Irecv &f.at(0,0) ← down's top interior row  :37 */
            memcpy(&t->data[fidx(t, 0, 0)], snap[4 * down + 3], sizeof(double) * (size_t)nxt);
        if (up != ORC_PROC_NULL) /* Irecv &f.at(0,h+ny) ← up's bottom interior row  :41 */
            memcpy(&t->data[fidx(t, 0, h + ny)], snap[4 * up + 2], sizeof(double) * (size_t)nxt);
    }
    for (int k = 0; k < 4 * nranks; ++k) free(snap[k]);
    free(snap);
}

/* One time step on one tile exactly as src/main.cpp:102-109 orders it (the halo exchange at :101
 * is the caller's): apply_boundary(u) → copy u→tmp → diffusion(u,tmp) → advection(u,tmp) → swap.
 * The swap is by pointer exchange, as std::swap(u.data,tmp.data) exchanges buffers. */
void orc_step_tile(orc_field* u, orc_field* tmp, const int nbr[4], const orc_config* cfg) {
    orc_apply_boundary(u->data, u->nx, u->ny, u->h, nbr, cfg->bc, 0.0);
    memcpy(tmp->data, u->data, sizeof(double) * (size_t)nx_tot(u) * (size_t)ny_tot(u));
    orc_diffusion_step(u->data, tmp->data, u->nx, u->ny, u->h, u->dx, u->dy, cfg->D, cfg->dt);
    orc_advection_step(u->data, tmp->data, u->nx, u->ny, u->h, u->dx, u->dy, cfg->vx, cfg->vy,
                       cfg->dt);
    double* t = u->data;
    u->data = tmp->data;
    tmp->data = t;
}

/* The driver loop of src/main.cpp:62-118 on `nranks` emulated ranks in one process.
 *   - dt is clamped to safe_dt first (main.cpp:42-49) unless clamp_dt == 0;
 *   - frames: the de-haloed GLOBAL field (io.cpp:411-418 assembles tiles at {y_off,x_off}) is
 *     written into frames[k*ny*nx ...] at the START of step n when n % out_every == 0
 *     (main.cpp:96-99); at most max_frames are stored;
 *   - final: global interior after the last step (the reference never writes it; parity uses it);
 *   - u0_padded (optional, single rank only): a full padded tile (ny+2)*(nx+2) that replaces the
 *     preset initial condition, ghosts included — used for seeded random-input parity tests.
 * Returns the number of frames stored, or <0 on error. */
int orc_run(const orc_config* cfg_in, int nranks, int clamp_dt, const double* u0_padded,
            double* frames, int max_frames, double* final_interior, double* final_padded_rank0) {
    orc_config cfg = *cfg_in;
    if (clamp_dt) {
        const double lim = orc_safe_dt(cfg.dx, cfg.dy, cfg.vx, cfg.vy, cfg.D);
        if (cfg.dt > lim) cfg.dt = lim;
    }
    if (u0_padded && nranks != 1) return -2;
    orc_decomp* dec = (orc_decomp*)malloc(sizeof(orc_decomp) * (size_t)nranks);
    orc_field* u = (orc_field*)malloc(sizeof(orc_field) * (size_t)nranks);
    orc_field* tmp = (orc_field*)malloc(sizeof(orc_field) * (size_t)nranks);
    const int halo = 1; /* main.cpp:65 */
    for (int r = 0; r < nranks; ++r) {
        orc_decomp_init(&dec[r], nranks, r, cfg.nx, cfg.ny);
        if (orc_field_init(&u[r], dec[r].nx_local, dec[r].ny_local, halo, cfg.dx, cfg.dy)) return -1;
        if (orc_field_init(&tmp[r], dec[r].nx_local, dec[r].ny_local, halo, cfg.dx, cfg.dy)) return -1;
        if (u0_padded)
            memcpy(u[r].data, u0_padded, sizeof(double) * (size_t)nx_tot(&u[r]) * (size_t)ny_tot(&u[r]));
        else
            orc_apply_ic(&dec[r], &u[r], &cfg);
    }
    int nframes = 0;
    const size_t gsz = (size_t)cfg.nx * (size_t)cfg.ny;
    for (int n = 0; n <= cfg.steps; ++n) {
        const int want_frame = (n < cfg.steps) && frames && (n % cfg.out_every == 0) && nframes < max_frames;
        const int want_final = (n == cfg.steps) && final_interior;
        if (want_frame || want_final) {
            double* dst = want_final ? final_interior : frames + (size_t)nframes * gsz;
            for (int r = 0; r < nranks; ++r) {
                const orc_field* t = &u[r];
                for (int j = 0; j < t->ny; ++j)
                    memcpy(dst + (size_t)(dec[r].y_offset + j) * cfg.nx + dec[r].x_offset,
                           &t->data[fidx(t, t->h, t->h + j)], sizeof(double) * (size_t)t->nx);
            }
            if (want_frame) ++nframes;
        }
        if (n == cfg.steps) break;
        orc_exchange_halos(u, dec, nranks);
        for (int r = 0; r < nranks; ++r) {
            const int nbr[4] = {dec[r].nbr_lr[0], dec[r].nbr_lr[1], dec[r].nbr_du[0], dec[r].nbr_du[1]};
            orc_step_tile(&u[r], &tmp[r], nbr, &cfg);
        }
    }
    if (final_padded_rank0)
        memcpy(final_padded_rank0, u[0].data,
               sizeof(double) * (size_t)nx_tot(&u[0]) * (size_t)ny_tot(&u[0]));
    for (int r = 0; r < nranks; ++r) {
        orc_field_free(&u[r]);
        orc_field_free(&tmp[r]);
    }
    free(u);
    free(tmp);
    free(dec);
    return nframes;
}

/* main.cpp:73-77 — rank 0's "IC min/max" is taken over its whole padded tile, ghosts included. */
void orc_minmax(const double* data, size_t n, double* mn, double* mx) {
    double lo = data[0], hi = data[0];
    for (size_t k = 1; k < n; ++k) {
        if (data[k] < lo) lo = data[k];
        if (hi < data[k]) hi = data[k];
    }
    *mn = lo;
    *mx = hi;
}
