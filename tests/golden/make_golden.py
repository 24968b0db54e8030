"""Generate tests/golden/*.npz from the reference's OWN compiled objects (oracle/_ref/libcsim_ref.so).

Run in the build container, where /root/reference exists:
    python tests/golden/make_golden.py
The vectors are small on purpose (a few hundred KiB in total) and are committed; the GPU box has
no /root/reference, so the -m gpu tests and the port oracle are checked against these files.

Each case stores: the SimParams fields, the padded input tile `u0` (ghosts included, seeded RNG or
the preset initial condition), the global interior after `steps` steps (`final`), the frames the
reference driver would have written (`frames`, main.cpp:93-99) and rank 0's padded tile (`padded`,
single-rank cases only; its four corner cells are excluded from comparisons).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.cpu_oracle import DIRICHLET, NEUMANN, PERIODIC, Oracle, SimParams, build  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases():
    D, N, P = DIRICHLET, NEUMANN, PERIODIC
    rng = np.random.default_rng(20261018)
    out = []

    def add(name, p, random_input, nranks=1):
        u0 = None
        if random_input:
            # values spanning many binades, both signs, ghosts included (they matter for Periodic)
            u0 = rng.standard_normal((p.ny + 2, p.nx + 2)) * 10.0 ** rng.integers(-3, 3, (p.ny + 2, p.nx + 2))
        out.append((name, p, u0, nranks))

    # dev.yaml physics and BC mix on a small grid, preset IC, 4 emulated ranks (configs[0] in small)
    add("dev_small_4ranks", SimParams(nx=48, ny=40, D=0.05, vx=0.5, vy=0.0, dt=0.1, steps=30,
                                      out_every=10, bc=(D, N, P, D)), False, nranks=4)
    # all four upwind branches, random data, every BC on every side at least once
    k = 0
    for vx, vy in ((0.5, 0.25), (-0.5, 0.25), (0.5, -0.25), (-0.75, -0.5), (0.0, 0.0)):
        for bc in ((D, D, D, D), (N, N, N, N), (P, P, P, P), (D, N, P, N), (P, D, N, D)):
            add(f"rand_{k:02d}", SimParams(nx=37, ny=23, D=0.05, vx=vx, vy=vy, dt=0.1, steps=6,
                                           out_every=2, bc=bc), True)
            k += 1
    # non-power-of-two spacing → true IEEE division path
    add("nonpow2_spacing", SimParams(nx=33, ny=18, dx=0.3, dy=0.7, D=0.01, vx=-0.2, vy=0.4, dt=0.05,
                                     steps=5, out_every=5, bc=(N, D, D, P)), True)
    # power-of-two spacing other than 1 → exact reciprocal path with non-trivial factors
    add("pow2_spacing", SimParams(nx=40, ny=16, dx=0.5, dy=2.0, D=0.02, vx=0.3, vy=-0.6, dt=0.05,
                                  steps=5, out_every=5, bc=(D, P, N, N)), True)
    # ragged / degenerate shapes: single row, single column, 1x1, odd widths
    add("one_row", SimParams(nx=19, ny=1, D=0.05, vx=0.5, vy=-0.5, dt=0.1, steps=4, out_every=4,
                             bc=(N, D, P, N)), True)
    add("one_col", SimParams(nx=1, ny=21, D=0.05, vx=-0.5, vy=0.5, dt=0.1, steps=4, out_every=4,
                             bc=(D, N, N, P)), True)
    add("one_cell", SimParams(nx=1, ny=1, D=0.05, vx=0.5, vy=0.5, dt=0.1, steps=3, out_every=3,
                              bc=(N, N, N, N)), True)
    add("wide_odd", SimParams(nx=131, ny=9, D=0.1, vx=0.9, vy=0.1, dt=0.2, steps=7, out_every=3,
                              bc=(P, N, D, D)), True)
    add("tall_odd", SimParams(nx=7, ny=150, D=0.1, vx=-0.1, vy=-0.9, dt=0.2, steps=7, out_every=3,
                              bc=(N, P, D, N)), True)
    # dt above the stability limit → clamp path of main.cpp:42-49
    add("dt_clamped", SimParams(nx=24, ny=24, D=1.0, vx=2.0, vy=0.0, dt=5.0, steps=4, out_every=4,
                                bc=(D, D, N, N)), True)
    # decomposition with remainders, 6 ranks {3,2}
    add("ic_6ranks_remainder", SimParams(nx=50, ny=35, D=0.05, vx=0.5, vy=0.3, dt=0.1, steps=12,
                                         out_every=4, bc=(D, N, P, D)), False, nranks=6)
    # Signed zeros and zero velocity components (second session): flat patches (zero differences), +0.0
    # cells, -0.0 cells and a -0.0 velocity — the regime in which the GPU path may drop the advection
    # term of a +0.0 component, and the cases in which it must not.  Wide enough (3 strips) for the
    # branch-free hot path of the sweep.  A separate generator keeps the cases above unchanged.
    rz = np.random.default_rng(20261019)

    def add_zero(name, p, negzero):
        a = rz.standard_normal((p.ny + 2, p.nx + 2)) * 10.0 ** rz.integers(-3, 3, (p.ny + 2, p.nx + 2))
        for _ in range(10):
            y, x = rz.integers(0, p.ny), rz.integers(0, p.nx)
            a[y:y + rz.integers(2, 9), x:x + rz.integers(2, 9)] = rz.choice([0.0, 1.5, -2.25, 1e-300])
        a[rz.random(a.shape) < 0.03] = 0.0
        if negzero:
            a[rz.random(a.shape) < 0.03] = -0.0
            a[p.ny // 2:p.ny // 2 + 5, p.nx // 2:p.nx // 2 + 6] = -0.0
        out.append((name, p, a, 1))

    add_zero("zero_vy_flat_patches", SimParams(nx=260, ny=12, D=0.05, vx=0.5, vy=0.0, dt=0.1, steps=7,
                                               out_every=7, bc=(P, P, P, P)), False)
    add_zero("zero_vx_negzero_cells", SimParams(nx=260, ny=12, D=0.05, vx=0.0, vy=-0.3, dt=0.1, steps=7,
                                                out_every=7, bc=(D, N, P, N)), True)
    add_zero("zero_both_flat_patches", SimParams(nx=260, ny=12, D=0.05, vx=0.0, vy=0.0, dt=0.1, steps=7,
                                                 out_every=7, bc=(N, D, N, D)), False)
    add_zero("negzero_velocity", SimParams(nx=260, ny=12, D=0.05, vx=0.5, vy=-0.0, dt=0.1, steps=7,
                                           out_every=7, bc=(P, P, P, P)), False)
    return out


def main():
    build()
    ref = Oracle("ref")
    bundle = {}
    names = []
    for name, p, u0, nranks in cases():
        r = ref.run(p, nranks=nranks, u0_padded=u0, want_padded=(nranks == 1))
        names.append(name)
        meta = np.array([p.nx, p.ny, p.steps, p.out_every, *p.bc, p.ic_preset, nranks], dtype=np.int64)
        phys = np.array([p.dx, p.dy, p.D, p.vx, p.vy, p.dt, p.A, p.sigma_frac, p.xc_frac, p.yc_frac])
        bundle[name + "/meta"] = meta
        bundle[name + "/phys"] = phys
        if u0 is not None:
            bundle[name + "/u0"] = u0
        bundle[name + "/final"] = r["final"]
        bundle[name + "/frames"] = r["frames"]
        if r["padded"] is not None:
            bundle[name + "/padded"] = r["padded"]
    bundle["names"] = np.array(names)
    path = os.path.join(OUT, "timestep_golden.npz")
    np.savez_compressed(path, **bundle)
    print(f"wrote {path}: {len(names)} cases, {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
