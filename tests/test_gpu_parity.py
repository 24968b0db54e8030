"""Parity of the CUDA path (through the C ABI) against the oracle, bit for bit.  Needs a B200.

Tolerance: none.  Every comparison is on the uint64 bit patterns of the doubles; the only opt-out
is CSIM_STEP_FAST_RECIP with non-power-of-two spacing, tested with L-inf rel <= 1e-12.
"""
import numpy as np
import pytest

from conftest import bits_equal, mask_corners

pytestmark = pytest.mark.gpu

NONE = (-1, -1, -1, -1)


def make_fields(csim, ctx, u0, dx=1.0, dy=1.0, h=1):
    ny, nx = u0.shape[0] - 2 * h, u0.shape[1] - 2 * h
    u = csim.Field(ctx, nx, ny, h, dx, dy)
    t = csim.Field(ctx, nx, ny, h, dx, dy)
    u.upload(np.ascontiguousarray(u0))
    return u, t


def rand_tile(rng, ny, nx, h=1):
    return rng.standard_normal((ny + 2 * h, nx + 2 * h)) * 10.0 ** rng.integers(-3, 3, (ny + 2 * h, nx + 2 * h))


# ---- Field: tests/simulation/unit/test_field.cpp ------------------------------------------------

def test_field_allocation_layout_bounds(csim, ctx):
    f = csim.Field(ctx, 4, 3, 1, 1.0, 1.0)
    assert f.download().shape == (5, 6) and not f.download().any()  # zero-initialised, field.cpp:12
    f = csim.Field(ctx, 2, 2, 1, 1.0, 1.0)
    for j in range(4):
        for i in range(4):
            f.set(i, j, 10 * j + i)
    assert f.at(0, 0) == 0 and f.at(3, 0) == 3 and f.at(0, 1) == 10 and f.at(3, 3) == 33
    assert bits_equal(f.download(), np.add.outer(10.0 * np.arange(4), np.arange(4.0)))
    f = csim.Field(ctx, 4, 4, 1, 1.0, 1.0)
    for (i, j) in ((-1, 0), (6, 0), (0, 6)):
        with pytest.raises(IndexError, match="out of range"):  # std::out_of_range, field.cpp:16
            f.at(i, j)
    info = f.info
    assert info.pitch % 16 == 0 and info.interior % 128 == 0 and info.lead_x == 16


def test_field_roundtrip_fill_swap_copy_halo_widths(csim, ctx):
    rng = np.random.default_rng(0)
    for h in (0, 1, 2):  # h=0 must construct (reference test_io.cpp:129)
        a = rng.standard_normal((9 + 2 * h, 13 + 2 * h))
        f = csim.Field(ctx, 13, 9, h, 0.5, 2.0)
        f.upload(a)
        assert bits_equal(f.download(), a)
        assert bits_equal(f.download_interior(), a[h:h + 9, h:h + 13])
    g = csim.Field(ctx, 13, 9, 2, 0.5, 2.0)
    g.fill(7.25)
    assert np.all(g.download() == 7.25)
    f.swap(g)
    assert np.all(f.download() == 7.25) and bits_equal(g.download(), a)
    g.copy_to(f)
    assert bits_equal(f.download(), a)
    pin = ctx.pinned_empty((9, 13))
    g.download_interior_async(pin)
    ctx.sync()
    assert bits_equal(pin, a[2:-2, 2:-2])
    e = csim.Field(ctx, 0, 0, 1, 1.0, 1.0)  # empty interior
    assert e.download().shape == (2, 2)
    # large asynchronous transfers take the staged path (dense DMA + re-pitch kernel), twice in a row
    # through the same staging buffer, with odd sizes so that no row of the dense layout is aligned
    for (ny, nx) in ((700, 1025), (1031, 999)):
        big = rng.standard_normal((ny + 2, nx + 2))
        hin = ctx.pinned_empty(big.shape)
        hin[:] = big
        hout = ctx.pinned_empty((ny, nx))
        b = csim.Field(ctx, nx, ny, 1, 1.0, 1.0)
        b.upload_async(hin)
        b.download_interior_async(hout)
        ctx.sync()
        assert bits_equal(hout, big[1:-1, 1:-1])
        assert bits_equal(b.download(), big)  # ghost ring included


# ---- diffusion_step: test_diffusion.cpp + oracle ----------------------------------------------

def test_diffusion_single_impulse_one_step(csim, ctx):
    u0 = np.zeros((5, 5))
    u0[2, 2] = 1.0
    u, v = make_fields(csim, ctx, u0)
    D = dt = 0.1
    alpha = D * dt
    csim.diffusion_step(u, v, D, dt)
    assert abs(v.at(2, 2) - (1 - 4 * alpha)) <= 1e-12
    for (i, j) in ((1, 2), (3, 2), (2, 1), (2, 3)):
        assert abs(v.at(i, j) - alpha) <= 1e-12


@pytest.mark.parametrize("shape,dx,dy", [((23, 37), 1.0, 1.0), ((1, 19), 1.0, 1.0), ((21, 1), 1.0, 1.0),
                                         ((1, 1), 1.0, 1.0), ((64, 200), 0.5, 2.0), ((18, 33), 0.3, 0.7),
                                         ((130, 257), 1.0, 1.0)])
def test_diffusion_and_advection_match_oracle(csim, ctx, port, shape, dx, dy):
    rng = np.random.default_rng(sum(shape))
    ny, nx = shape
    u0 = rand_tile(rng, ny, nx)
    o0 = rand_tile(rng, ny, nx)
    u, out = make_fields(csim, ctx, u0, dx, dy)
    out.upload(o0)
    want = o0.copy()
    port.diffusion_step(u0, want, 1, dx, dy, 0.05, 0.1)
    csim.diffusion_step(u, out, 0.05, 0.1)
    got = out.download()
    assert bits_equal(got, want)  # interior AND the copied ring (diffusion.cpp:18-25)
    for vx, vy in ((0.5, 0.25), (-0.5, 0.25), (0.5, -0.25), (-0.75, -0.5), (0.0, 0.0)):
        port.advection_step(u0, want, 1, dx, dy, vx, vy, 0.1)
        csim.advection_step(u, out, vx, vy, 0.1)  # accumulates on top of the previous result
        assert bits_equal(out.download(), want), (vx, vy)


def test_advection_zero_velocity_signs_and_halo0_throws(csim, ctx):
    nx = ny = 8
    u0 = np.zeros((10, 10))
    u0[5, 5] = 1.0
    u, out = make_fields(csim, ctx, u0)
    csim.advection_step(u, out, 0.0, 0.0, 0.1)
    assert np.all(out.download_interior() == 0.0)
    for vx, vy in ((1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0)):
        out.fill(0.0)
        csim.advection_step(u, out, vx, vy, 0.1)
        assert out.at(5, 5) != 0.0
    a, b = csim.Field(ctx, 4, 4, 0, 1.0, 1.0), csim.Field(ctx, 4, 4, 0, 1.0, 1.0)
    with pytest.raises(IndexError):  # the reference's at(i-1,…) throws with halo 0
        csim.diffusion_step(a, b, 0.1, 0.1)
    c = csim.Field(ctx, 5, 4, 1, 1.0, 1.0)
    with pytest.raises(csim.CsimError):
        csim.diffusion_step(u, c, 0.1, 0.1)  # geometry mismatch


# ---- apply_boundary: test_boundary.cpp + oracle -------------------------------------------------

def test_boundary_known_answers(csim, ctx):
    NX, NY, h = 4, 3, 1
    f0 = np.full((NY + 2, NX + 2), -1.0)
    f0[1:-1, 1:-1] = 10.0
    f, _ = make_fields(csim, ctx, f0)
    csim.apply_boundary(f, None, csim.BCConfig(), 5.0)
    g = f.download()
    assert np.all(g[:, 0] == 5.0) and np.all(g[:, -1] == 5.0) and np.all(g[0] == 5.0) and np.all(g[-1] == 5.0)
    f0 = np.full((NY + 2, NX + 2), -1.0)
    for j in range(h, h + NY):
        f0[j, 1:-1] = float(j)
    f.upload(f0)
    N = csim.BCType.Neumann
    csim.apply_boundary(f, None, csim.BCConfig(N, N, N, N), 0.0)
    g = f.download()
    assert np.all(g[:, 0] == g[:, 1]) and np.all(g[:, -1] == g[:, -2])
    assert np.all(g[0] == g[1]) and np.all(g[-1] == g[-2])


def test_boundary_matches_oracle_all_combinations(csim, ctx, port):
    rng = np.random.default_rng(5)
    B = csim.BCType
    for (ny, nx) in ((3, 4), (1, 1), (17, 9), (40, 300)):
        f0 = rand_tile(rng, ny, nx)
        f, _ = make_fields(csim, ctx, f0)
        for bc in ((0, 0, 0, 0), (1, 1, 1, 1), (2, 2, 2, 2), (0, 1, 2, 1), (1, 2, 0, 0), (2, 0, 1, 2)):
            for nbr in (NONE, (3, -1, -1, 5), (-1, 2, 7, -1), (1, 2, 3, 4)):
                f.upload(f0)
                want = port.apply_boundary(f0.copy(), 1, nbr, bc, 2.5)
                csim.apply_boundary(f, nbr, csim.BCConfig(*[B(b) for b in bc]), 2.5)
                assert bits_equal(f.download(), want), (ny, nx, bc, nbr)  # corners included
    for h in (2, 3):  # the reference writes only the outermost layer whatever the halo (boundary.cpp:16-21)
        f0 = rand_tile(rng, 6, 7, h)
        f = csim.Field(ctx, 7, 6, h, 1.0, 1.0)
        f.upload(f0)
        want = port.apply_boundary(f0.copy(), h, NONE, (0, 1, 1, 0), -3.0)
        csim.apply_boundary(f, None, csim.BCConfig(B(0), B(1), B(1), B(0)), -3.0)
        assert bits_equal(f.download(), want)


# ---- the fused step vs golden vectors and the oracle -------------------------------------------

def run_case(csim, ctx, g, flags=0, chunk=None):
    """Run one golden case on one GPU; returns (frames, final interior, final padded)."""
    B = csim.BCType
    dec = csim.Decomp2D.single(g["nx"], g["ny"])
    if g["u0"] is not None:
        u0 = g["u0"]
    else:
        u0 = csim.initial_condition_host(dec, 1, g["dx"], g["dy"], "gaussian_hotspot", g["A"],
                                         g["sigma_frac"], g["xc_frac"], g["yc_frac"])
    u, tmp = make_fields(csim, ctx, u0, g["dx"], g["dy"])
    dt = min(g["dt"], csim.safe_dt(g["dx"], g["dy"], g["vx"], g["vy"], g["D"]))  # main.cpp:42-49
    p = csim.make_step_params(g["D"], g["vx"], g["vy"], dt, csim.BCConfig(*[B(b) for b in g["bc"]]),
                              dec, 0.0, flags)
    frames = []
    n = 0
    while n < g["steps"]:
        if n % g["out_every"] == 0:
            frames.append(u.download_interior())  # main.cpp:96-99
        k = g["out_every"] - n % g["out_every"]
        k = min(k, g["steps"] - n, chunk or k)
        csim.run_steps(u, tmp, p, dec, k)
        n += k
    return np.array(frames), u.download_interior(), u.download()


def test_fused_step_matches_every_golden_case(csim, ctx, golden):
    for name, g in golden.items():
        frames, final, padded = run_case(csim, ctx, g)
        assert bits_equal(final, g["final"]), name
        assert bits_equal(frames, g["frames"]), name
        if g["padded"] is not None:  # ghost ring as the reference leaves it; corners excluded (Q10)
            assert bits_equal(mask_corners(padded), mask_corners(g["padded"])), name


def test_fused_step_chunking_is_invisible(csim, ctx, golden):
    """Calling run_steps with 1, 2, 3, … steps at a time gives the same bits (temporal blocking,
    if any, must not leak)."""
    g = golden["rand_08"]
    for chunk in (1, 2, 3, 5):
        _, final, _ = run_case(csim, ctx, g, chunk=chunk)
        assert bits_equal(final, g["final"]), chunk
    _, final, _ = run_case(csim, ctx, g, flags=csim.STEP_NO_TEMPORAL)
    assert bits_equal(final, g["final"])


def test_fast_recip_mode_tolerance(csim, ctx, golden):
    g = golden["nonpow2_spacing"]
    _, final, _ = run_case(csim, ctx, g, flags=csim.STEP_FAST_RECIP)
    rel = np.max(np.abs(final - g["final"])) / np.max(np.abs(g["final"]))
    assert rel <= 1e-12  # stated tolerance of the opt-in reciprocal mode
    g = golden["pow2_spacing"]  # power-of-two spacing: the reciprocal IS exact
    _, final, _ = run_case(csim, ctx, g, flags=csim.STEP_FAST_RECIP)
    assert bits_equal(final, g["final"])


def test_dev_yaml_all_frames(csim, ctx, oracle_mod, port):
    """configs[0]: configs/dev.yaml, 512², 1000 steps, frames every 100, BC dirichlet/neumann/
    periodic/dirichlet.  The reference runs it on 4 ranks; by decomposition invariance (pinned in
    test_oracle.py) the single-rank oracle is the same field."""
    p = oracle_mod.SimParams(**oracle_mod.DEV_YAML)
    want = port.run(p)
    g = dict(nx=p.nx, ny=p.ny, dx=p.dx, dy=p.dy, D=p.D, vx=p.vx, vy=p.vy, dt=p.dt, steps=p.steps,
             out_every=p.out_every, bc=p.bc, u0=None, A=p.A, sigma_frac=p.sigma_frac, xc_frac=p.xc_frac,
             yc_frac=p.yc_frac)
    frames, final, _ = run_case(csim, ctx, g)
    assert frames.shape == (10, 512, 512)
    assert bits_equal(frames, want["frames"]) and bits_equal(final, want["final"])


def test_ragged_sizes_against_oracle(csim, ctx, oracle_mod, port):
    rng = np.random.default_rng(11)
    for (ny, nx) in ((2, 2), (3, 127), (129, 2), (33, 65), (64, 64), (100, 1000), (257, 513)):
        u0 = rand_tile(rng, ny, nx)
        p = oracle_mod.SimParams(nx=nx, ny=ny, D=0.05, vx=-0.4, vy=0.3, dt=0.1, steps=4, out_every=4,
                                 bc=(1, 2, 0, 1))
        want = port.run(p, u0_padded=u0)["final"]
        g = dict(nx=nx, ny=ny, dx=1.0, dy=1.0, D=p.D, vx=p.vx, vy=p.vy, dt=p.dt, steps=4, out_every=4,
                 bc=p.bc, u0=u0)
        _, final, _ = run_case(csim, ctx, g)
        assert bits_equal(final, want), (ny, nx)


# ---- full-size properties (BASELINE.json sizes; the oracle cannot run these whole) --------------

@pytest.mark.parametrize("n", [8192])
def test_full_size_windows_and_linearity(csim, ctx, oracle_mod, port, n):
    """configs[1]: 8192², all-periodic.  (a) windows of the field, with a 3-cell margin, are run
    through the oracle for 3 steps and must match the GPU field bit for bit in the window core;
    (b) step(2u) == 2·step(u) exactly (scaling by a power of two commutes with every rounding);
    (c) a constant field under all-Neumann boundaries stays constant."""
    steps = 3
    dec = csim.Decomp2D.single(n, n)
    u0 = csim.initial_condition_host(dec, 1, 1.0, 1.0)
    rng = np.random.default_rng(7)
    u0[1:-1, 1:-1] += 1e-3 * rng.standard_normal((n, n))  # break the Gaussian's symmetry
    P = csim.BCType.Periodic
    u, tmp = make_fields(csim, ctx, u0)
    p = csim.make_step_params(0.05, -0.5, 0.25, 0.1, csim.BCConfig(P, P, P, P), dec)
    csim.run_steps(u, tmp, p, dec, steps)
    got = u.download_interior()
    m, w = steps, 96
    for (y, x) in ((0, 0), (0, n - w), (n - w, 0), (n - w, n - w), (n // 2 - 48, n // 2 - 48),
                   (1234, 4321), (0, 5000), (7000, n - w)):
        y0, x0 = max(y - m, 0), max(x - m, 0)
        y1, x1 = min(y + w + m, n), min(x + w + m, n)
        sub = np.ascontiguousarray(u0[y0:y1 + 2, x0:x1 + 2])  # padded coords: +1 offset, +2 ghosts
        sp = oracle_mod.SimParams(nx=x1 - x0, ny=y1 - y0, D=0.05, vx=-0.5, vy=0.25, dt=0.1, steps=steps,
                                  out_every=steps, bc=(2, 2, 2, 2))
        want = port.run(sp, u0_padded=sub)["final"]
        assert bits_equal(got[y:y + w, x:x + w], want[y - y0:y - y0 + w, x - x0:x - x0 + w]), (y, x)
    # (b) linearity under power-of-two scaling
    u.upload(2.0 * u0)
    csim.run_steps(u, tmp, p, dec, steps)
    assert bits_equal(u.download_interior(), 2.0 * got)
    # (c) constants are fixed points under Neumann
    N = csim.BCType.Neumann
    u.fill(3.141592653589793)
    pn = csim.make_step_params(0.05, 0.5, -0.25, 0.1, csim.BCConfig(N, N, N, N), dec)
    csim.run_steps(u, tmp, pn, dec, 5)
    assert np.all(u.download_interior() == 3.141592653589793)


def _window_oracle(csim, oracle_mod, port, nxg, nyg, gx, gy, w, steps, phys, bc):
    """Oracle field of the w x w window at global (gx, gy) after `steps` time steps from the Gaussian initial
    condition: the sub-domain reaches `steps` cells beyond the window (its dependency cone) or to the physical
    boundary; cut sides carry the neighbouring cells' initial values as frozen ghosts."""
    y0, y1 = max(gy - steps, 0), min(gy + w + steps, nyg)
    x0, x1 = max(gx - steps, 0), min(gx + w + steps, nxg)
    sub = np.zeros((y1 - y0 + 2, x1 - x0 + 2))
    iy0, iy1, ix0, ix1 = max(y0 - 1, 0), min(y1 + 1, nyg), max(x0 - 1, 0), min(x1 + 1, nxg)
    sub[iy0 - (y0 - 1):iy1 - (y0 - 1), ix0 - (x0 - 1):ix1 - (x0 - 1)] = csim.initial_condition_host(
        csim.Decomp2D.window(nxg, nyg, ix0, iy0, ix1 - ix0, iy1 - iy0), 0, 1.0, 1.0)
    cut = (bc[0] if x0 == 0 else 2, bc[1] if x1 == nxg else 2, bc[2] if y0 == 0 else 2, bc[3] if y1 == nyg else 2)
    sp = oracle_mod.SimParams(nx=x1 - x0, ny=y1 - y0, steps=steps, out_every=steps, bc=cut, **phys)
    return port.run(sp, u0_padded=sub)["final"][gy - y0:gy - y0 + w, gx - x0:gx - x0 + w]


@pytest.mark.parametrize("n,bc", [(16384, (2, 2, 2, 2)), (32768, (0, 1, 0, 1))])
def test_north_star_tiles_100_steps_against_oracle_windows(csim, ctx, oracle_mod, port, n, bc):
    """configs[2]'s 16384^2 tile (all-periodic) and configs[3]'s 32768^2 grid with Dirichlet left/bottom and
    Neumann right/top on ONE GPU (2 x 8.6 GB), 100 time steps — a whole output window: 25 four-step sweeps,
    or 33 + a remainder sweep — for the headline physics (vy = 0: the dropped-term kernel) and the all-terms
    physics.  The initial condition is generated on the device where the host libm allows (else uploaded),
    the oracle runs on the sub-domain around each of 11 windows (four corners, four edge midpoints, centre,
    two interior points), and the windows come back with csim_field_download_window."""
    steps, w = 100, 48
    dec = csim.Decomp2D.single(n, n)
    B = csim.BCType
    u = csim.Field(ctx, n, n, 1, 1.0, 1.0)
    tmp = csim.Field(ctx, n, n, 1, 1.0, 1.0)
    pts = [(0, 0), (0, n - w), (n - w, 0), (n - w, n - w), (0, n // 2), (n - w, n // 2 - 7), (n // 2, 0),
           (n // 3, n - w), (n // 2 - w // 2, n // 2 - w // 2), (1234, 4321), (3 * n // 4, n // 3 + 5)]
    try:
        for phys in (dict(D=0.05, vx=0.5, vy=0.0, dt=0.1), dict(D=0.05, vx=-0.5, vy=0.25, dt=0.1)):
            if csim.exp_variant() >= 0:
                u.fill(0.0)
                csim.initial_condition_device(u, dec)
            else:
                u.upload(csim.initial_condition_host(dec, 1, 1.0, 1.0))
            p = csim.make_step_params(phys["D"], phys["vx"], phys["vy"], phys["dt"], csim.BCConfig(*[B(b) for b in bc]),
                                      dec)
            csim.run_steps(u, tmp, p, dec, steps)
            if phys["vy"] == 0.0:
                assert u.value_state == 1  # the scan found the tile clean: the dropped-term kernel ran
            for (y, x) in pts:
                want = _window_oracle(csim, oracle_mod, port, n, n, x, y, w, steps, phys, bc)
                assert bits_equal(u.download_window(x, y, w, w), want), (n, phys, y, x)
    finally:
        u.close()
        tmp.close()


# ---- dropped zero-velocity terms (csim_field_value_state) ------------------------------

def plateau_tile(rng, ny, nx, negzero=False):
    """Random tile with flat patches (zero differences → signed-zero arithmetic in the advection term),
    exact +0.0 cells and, optionally, -0.0 cells (which must send the run to the full arithmetic)."""
    a = rand_tile(rng, ny, nx)
    for _ in range(12):
        y, x = rng.integers(0, ny), rng.integers(0, nx)
        a[y:y + rng.integers(2, 9), x:x + rng.integers(2, 9)] = rng.choice([0.0, 1.5, -2.25, 1e-300])
    a[rng.random(a.shape) < 0.02] = 0.0
    if negzero:
        a[rng.random(a.shape) < 0.02] = -0.0
        y, x = ny // 2, nx // 2
        a[y:y + 6, x:x + 6] = -0.0  # a patch: c == -0 with o == -0 is the one case that differs
    return a


@pytest.mark.parametrize("vx,vy,D", [(0.5, 0.0, 0.05), (-0.4, 0.0, 0.05), (0.0, 0.3, 0.05), (0.0, -0.3, 0.0),
                                     (0.0, 0.0, 0.05), (0.5, 0.0, 0.0)])
def test_zero_velocity_terms_are_exact(csim, ctx, oracle_mod, port, vx, vy, D):
    rng = np.random.default_rng(int(1000 * abs(vx) + 100 * abs(vy)) + 3)
    T = csim.steps_per_sweep()
    steps = 2 * T + 1  # two full-depth sweeps (dropped-term kernels) + a remainder (full arithmetic)
    for (ny, nx), bc in (((130, 257), (2, 2, 2, 2)), ((64, 300), (0, 1, 2, 1)), ((257, 140), (1, 0, 1, 0))):
        for negzero in (False, True):
            u0 = plateau_tile(rng, ny, nx, negzero)
            sp = oracle_mod.SimParams(nx=nx, ny=ny, D=D, vx=vx, vy=vy, dt=0.1, steps=steps, out_every=steps, bc=bc)
            want = port.run(sp, u0_padded=u0)["final"]
            u, tmp = make_fields(csim, ctx, u0)
            dec = csim.Decomp2D.single(nx, ny)
            B = csim.BCType
            p = csim.make_step_params(D, vx, vy, 0.1, csim.BCConfig(*[B(b) for b in bc]), dec)
            assert u.value_state == 0
            csim.run_steps(u, tmp, p, dec, steps)
            assert bits_equal(u.download_interior(), want), (ny, nx, bc, negzero)
            if T >= 3:  # the scan ran and classified the tile; clean tiles took the dropped-term kernels
                assert u.value_state == (2 if negzero else 1)


def test_zero_velocity_terms_fallbacks(csim, ctx, oracle_mod, port):
    """Non-monotone time step, -0.0 velocity, non-finite cell: all must run the full arithmetic and
    match the reference bit for bit (NaN payloads aside, none arise here)."""
    rng = np.random.default_rng(77)
    ny, nx, steps = 96, 260, 7
    dec = csim.Decomp2D.single(nx, ny)
    P = csim.BCType.Periodic
    for (D, vx, vy, dt) in ((0.2, 0.5, 0.0, 1.0),     # dt*(2D*2+|vx|) = 1.3 > 1 although dt <= safe_dt
                            (0.05, 0.5, -0.0, 0.1)):  # -0.0 selects the backward difference AND flips the zero's sign
        u0 = plateau_tile(rng, ny, nx)
        sp = oracle_mod.SimParams(nx=nx, ny=ny, D=D, vx=vx, vy=vy, dt=dt, steps=steps, out_every=steps, bc=(2, 2, 2, 2))
        want = port.run(sp, u0_padded=u0)["final"]
        u, tmp = make_fields(csim, ctx, u0)
        p = csim.make_step_params(D, vx, vy, dt, csim.BCConfig(P, P, P, P), dec)
        csim.run_steps(u, tmp, p, dec, steps)
        assert bits_equal(u.download_interior(), want), (D, vx, vy, dt)
        assert u.value_state != 1 or vy != 0.0 or dt < 1.0
    u0 = plateau_tile(rng, ny, nx)
    u0[40, 100] = np.inf
    u, tmp = make_fields(csim, ctx, u0)
    p = csim.make_step_params(0.05, 0.5, 0.0, 0.1, csim.BCConfig(P, P, P, P), dec)
    T = csim.steps_per_sweep()
    csim.run_steps(u, tmp, p, dec, T)
    assert u.value_state == 2 or T < 3
    sp = oracle_mod.SimParams(nx=nx, ny=ny, D=0.05, vx=0.5, vy=0.0, dt=0.1, steps=T, out_every=T, bc=(2, 2, 2, 2))
    want = port.run(sp, u0_padded=u0)["final"]
    got = u.download_interior()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
    ok = np.isfinite(want)
    assert bits_equal(got[ok], want[ok])


# ---- reductions ----------------------------------------------------------------------------------

def test_minmax_and_health(csim, ctx):
    rng = np.random.default_rng(9)
    for (ny, nx) in ((1, 1), (7, 5), (300, 1000), (2049, 1025)):
        a = rng.standard_normal((ny + 2, nx + 2))
        f, _ = make_fields(csim, ctx, a)
        assert f.minmax() == (a.min(), a.max())  # over the padded tile, ghosts included (main.cpp:74-75)
        mx, bad = f.health()
        assert mx == np.abs(a[1:-1, 1:-1]).max() and bad == 0
    a[5 % (ny + 1) + 0, 3] = np.nan
    a[1, 1] = np.inf
    f.upload(a)
    assert f.health()[1] == 2


# ---- initial condition generated on the device (N4): src/init.cpp:12-47 ---------------------------

def _ic_pair(csim, ctx, size, rank, nxg, nyg, dx, dy, A, sf, xc, yc):
    dec = csim.Decomp2D.init(size, rank, nxg, nyg)
    want = csim.initial_condition_host(dec, 1, dx, dy, "gaussian_hotspot", A, sf, xc, yc)
    f = csim.Field(ctx, dec.nx_local, dec.ny_local, 1, dx, dy)
    csim.initial_condition_device(f, dec, "gaussian_hotspot", A, sf, xc, yc)
    got = f.download()
    f.close()
    return got, want


def test_initial_condition_device_matches_host(csim, ctx):
    """Tiles of 1-, 4- and 8-rank decompositions (offsets, remainders), non-unit spacing, narrow hotspots
    whose tails underflow through the subnormal range to zero, an off-centre hotspot: the device tile equals
    the host tile bit for bit, ghost ring (untouched zeros) included."""
    if csim.exp_variant() < 0:
        pytest.skip("host libm exp() is not one of the restated variants: device path disabled")
    cases = [(1, 64, 64, 1.0, 1.0, 1.0, 0.05, 0.5, 0.5), (4, 513, 300, 1.0, 1.0, 2.5, 0.05, 0.5, 0.5),
             (8, 1000, 777, 0.3, 0.7, 1.0, 0.05, 0.3, 0.8), (4, 2048, 2048, 1.0, 1.0, 1.0, 0.004, 0.5, 0.5),
             (2, 4096, 512, 0.5, 2.0, -3.0, 0.0125, 0.1, 0.9), (1, 7, 5, 1.0, 1.0, 1.0, 0.5, 0.5, 0.5)]
    for (size, nxg, nyg, dx, dy, A, sf, xc, yc) in cases:
        for rank in range(size):
            got, want = _ic_pair(csim, ctx, size, rank, nxg, nyg, dx, dy, A, sf, xc, yc)
            assert bits_equal(got, want), (size, rank, nxg, nyg, sf)
    f = csim.Field(ctx, 8, 8, 1, 1.0, 1.0)
    csim.initial_condition_device(f, csim.Decomp2D.single(8, 8), "constant_zero")  # no-op, init.cpp:39-40
    assert not f.download().any()
    with pytest.raises(RuntimeError, match="Unknown IC preset"):
        csim.initial_condition_device(f, csim.Decomp2D.single(8, 8), "checkerboard")


def test_initial_condition_device_full_size(csim, ctx):
    """configs[2]'s 16384^2 tile: all 2.7e8 cells, device against host."""
    if csim.exp_variant() < 0:
        pytest.skip("host libm exp() is not one of the restated variants: device path disabled")
    n = 16384
    dec = csim.Decomp2D.single(n, n)
    want = csim.initial_condition_host(dec, 1, 1.0, 1.0)
    f = csim.Field(ctx, n, n, 1, 1.0, 1.0)
    csim.initial_condition_device(f, dec)
    got = f.download()
    f.close()
    mismatches = int(np.count_nonzero(got.view(np.uint64) != want.view(np.uint64)))
    assert mismatches == 0, f"{mismatches} of {n * n} cells differ"


# ---- multi-GPU halo exchange (needs >= 2 devices; the 1-GPU box skips) ---------------------------

def _rank_worker(csim, size, rank, uid, nxg, nyg, steps, phys, bc, flags, results, errors):
    try:
        c = csim.Context(rank)
        c.comm_init(size, rank, uid)
        dec = csim.Decomp2D.init(size, rank, nxg, nyg)
        t0 = csim.initial_condition_host(dec, 1, 1.0, 1.0)
        u = csim.Field(c, dec.nx_local, dec.ny_local, 1, 1.0, 1.0)
        tmp = csim.Field(c, dec.nx_local, dec.ny_local, 1, 1.0, 1.0)
        u.upload(t0)
        p = csim.make_step_params(*phys, csim.BCConfig(*[csim.BCType(b) for b in bc]), dec, 0.0, flags)
        for k in steps:  # several calls: the block structure must not leak across calls
            csim.run_steps(u, tmp, p, dec, k)
        results[rank] = (dec, u.download_interior())
        c.sync()
    except Exception as e:  # noqa: BLE001
        errors.append((rank, repr(e)))


def _run_ranks(csim, size, nxg, nyg, steps, phys, bc, flags=0):
    import threading
    uid = csim.comm_unique_id()
    results, errors = {}, []
    th = [threading.Thread(target=_rank_worker, args=(csim, size, r, uid, nxg, nyg, steps, phys, bc, flags,
                                                      results, errors)) for r in range(size)]
    [t.start() for t in th]
    [t.join(180) for t in th]
    assert not errors, errors
    assert len(results) == size, "a rank did not finish"
    glob = np.zeros((nyg, nxg))
    for dec, tile in results.values():
        glob[dec.y_offset:dec.y_offset + dec.ny_local, dec.x_offset:dec.x_offset + dec.nx_local] = tile
    return glob


def test_multi_gpu_matches_single_rank_oracle(csim, oracle_mod, port):
    """Ranks = GPUs of one box (threads here, processes under torchrun).  By decomposition
    invariance the single-rank oracle pins every decomposition.  Covers: the blocked path with the
    wide (T-line, 8-neighbour) exchange overlapped with the interior sweep, the reference-shaped
    one-line exchange (CSIM_STEP_NO_TEMPORAL), tiles with remainders, tiles too small to split."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    cases = [  # (nxg, nyg, step chunks, physics, bc)
        (203, 150, (25,), (0.05, -0.5, 0.3, 0.1), (0, 1, 2, 0)),
        (1000, 700, (7, 3, 1, 5), (0.05, 0.5, -0.3, 0.1), (2, 2, 2, 2)),
        (777, 1301, (10,), (0.05, -0.4, -0.2, 0.1), (1, 0, 1, 2)),
    ]
    for size in [s for s in (2, 4, 8) if s <= ngpu]:
        for (nxg, nyg, steps, phys, bc) in cases:
            sp = oracle_mod.SimParams(nx=nxg, ny=nyg, D=phys[0], vx=phys[1], vy=phys[2], dt=phys[3],
                                      steps=sum(steps), out_every=sum(steps), bc=bc)
            want = port.run(sp)["final"]
            got = _run_ranks(csim, size, nxg, nyg, steps, phys, bc)
            assert bits_equal(got, want), ("blocked, NCCL exchange", size, nxg, nyg)
            got = _run_ranks(csim, size, nxg, nyg, steps, phys, bc, flags=csim.STEP_NO_TEMPORAL)
            assert bits_equal(got, want), ("one-line", size, nxg, nyg)


# ---- the C++ drop-in layer (host/): the reference's unit tests restated in C++ --------------------

def test_cpp_dropin_unit_tests():
    """host/tests/test_dropin.cpp: Field / diffusion / advection / boundary / stability / decomp tests
    of the reference, written against the drop-in headers, plus main.cpp's loop body vs the fused path."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "climate-sim-mpi-cpp_b200", "host", "build", "test_dropin")
    assert os.path.exists(exe), "host/build/test_dropin missing: run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("env,path", [({}, "nccl"), ({"CSIM_HALO": "peer"}, "peer"), ({"CSIM_LOOP": "coupled"}, "nccl"),
                                      ({"CSIM_GRAPH": "1"}, "nccl"), ({"CSIM_LOOP": "coupled", "CSIM_HALO": "peer"}, "peer"),
                                      ({"CSIM_HALO": "peer", "CSIM_GRAPH": "1"}, "peer")])
def test_multi_process_parity_under_torchrun(env, path):
    """One PROCESS per GPU under torchrun (how bench.py runs): the split block loop (frame and interior as two
    launches on two streams; the default) with the NCCL halo path and with the peer-store path
    (CSIM_HALO=peer: tiles mapped with CUDA IPC, which the threaded test above cannot exercise), eager and
    replayed as a CUDA graph, and the coupled loop (CSIM_LOOP=coupled: one sweep launch per block, flags between
    the sweep and the exchange stream) on both halo paths."""
    import os
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mp_parity_worker.py")]
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0 and "MP_PARITY_PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"halo={path}" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("env", [{"CSIM_TB_MAXT": "4"}, {"CSIM_TB_MAXT": "2"}, {"CSIM_TB_MAXT": "1"},
                                 {"CSIM_TB_MAXT": "3", "CSIM_TB_CHUNK": "7", "CSIM_TB_EDGE_SPLIT": "3"},
                                 {"CSIM_TB_MAXT": "4", "CSIM_TB_CHUNK": "500", "CSIM_TB_EDGE_SPLIT": "1",
                                  "CSIM_TB_PF": "0"},
                                 {"CSIM_TB_MAXT": "3", "CSIM_TB_KERNEL": "reg"},
                                 {"CSIM_TB_MAXT": "4", "CSIM_TB_KERNEL": "reg", "CSIM_TB_CHUNK": "33"},
                                 {"CSIM_TB_MAXT": "3", "CSIM_TB_DIV_MAXT": "3"},
                                 {"CSIM_TB_MAXT": "4", "CSIM_TB_DIV_MAXT": "2", "CSIM_TB_CHUNK": "21"}])
def test_every_blocking_depth_and_chunking_matches_golden(env):
    """The sweep is instantiated for T = 1..4 in two builds — level-0 rows staged through shared memory by
    TMA (the default for T >= 3) and register-only (CSIM_TB_KERNEL=reg; always for T <= 2) — and with
    IEEE division up to T = 3 (CSIM_TB_DIV_MAXT).  Each depth and build, odd chunk heights and edge splits
    must give the same bits (the knobs are read once per process → subprocess)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden_runner.py")], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"T={env['CSIM_TB_MAXT']} " in r.stdout
