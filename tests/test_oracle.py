"""The oracle pinned: port == reference objects == golden vectors == the reference's known answers.

CPU only.  Each known-answer test names the reference test it restates.
"""
import numpy as np
import pytest

from conftest import bits_equal


def params_of(oracle_mod, g, **over):
    keys = ("nx", "ny", "dx", "dy", "D", "vx", "vy", "dt", "steps", "out_every", "bc", "ic_preset", "A",
            "sigma_frac", "xc_frac", "yc_frac")
    d = {k: g[k] for k in keys}
    d.update(over)
    return oracle_mod.SimParams(**d)


def test_port_matches_golden(oracle_mod, port, golden):
    for name, g in golden.items():
        r = port.run(params_of(oracle_mod, g), nranks=g["nranks"], u0_padded=g["u0"], want_padded=True)
        assert bits_equal(r["final"], g["final"]), name
        assert bits_equal(r["frames"], g["frames"]), name
        if g["padded"] is not None:
            # whole padded tile, ghost ring included, corners too (single rank: deterministic)
            assert bits_equal(r["padded"], g["padded"]), name


def test_ref_objects_match_golden(oracle_mod, ref, golden):
    for name, g in golden.items():
        r = ref.run(params_of(oracle_mod, g), nranks=g["nranks"], u0_padded=g["u0"])
        assert bits_equal(r["final"], g["final"]), name


def test_port_is_decomposition_invariant(oracle_mod, port, golden):
    g = golden["ic_6ranks_remainder"]
    base = port.run(params_of(oracle_mod, g), nranks=1)
    for n in (2, 3, 4, 6, 8):
        r = port.run(params_of(oracle_mod, g), nranks=n)
        assert bits_equal(r["final"], base["final"]), n
    assert bits_equal(base["final"], g["final"])


def test_ref_O0_equals_O2(oracle_mod, golden):
    if not oracle_mod.available("ref_O0"):
        pytest.skip("flagless reference build absent")
    o0 = oracle_mod.Oracle("ref_O0")
    g = golden["rand_07"]
    r = o0.run(params_of(oracle_mod, g), u0_padded=g["u0"])
    assert bits_equal(r["final"], g["final"])


def test_dev_yaml_four_ranks_vs_single(oracle_mod, port, ref):
    """configs[0] (dev.yaml, mpirun -np 4) shortened to 200 steps: reference objects on 4 emulated
    ranks == port on 1 rank, all frames and the final state."""
    p = oracle_mod.SimParams(**oracle_mod.DEV_YAML)
    p.steps = 200
    a = ref.run(p, nranks=4)
    b = port.run(p, nranks=1)
    assert a["frames"].shape == (2, 512, 512)
    assert bits_equal(a["frames"], b["frames"]) and bits_equal(a["final"], b["final"])


# ---- the reference's own unit tests, restated against both oracles ----------------------------

@pytest.fixture(params=["port", "ref"])
def orc(request, port, ref):
    return port if request.param == "port" else ref


def test_diffusion_single_impulse_one_step(orc):
    """tests/simulation/unit/test_diffusion.cpp:17-34"""
    u = np.zeros((5, 5))
    v = np.zeros((5, 5))
    u[2, 2] = 1.0
    D = dt = 0.1
    alpha = D * dt / 1.0
    orc.diffusion_step(u, v, 1, 1.0, 1.0, D, dt)
    assert abs(v[2, 2] - (1 - 4 * alpha)) <= 1e-12
    for (j, i) in ((2, 1), (2, 3), (1, 2), (3, 2)):
        assert abs(v[j, i] - alpha) <= 1e-12


def test_advection_zero_velocity_and_signs(orc):
    """tests/simulation/unit/test_advection.cpp:13-71"""
    nx = ny = 8
    u = np.zeros((ny + 2, nx + 2))
    u[ny // 2 + 1, nx // 2 + 1] = 1.0
    out = np.zeros_like(u)
    orc.advection_step(u, out, 1, 1.0, 1.0, 0.0, 0.0, 0.1)
    assert np.all(out[1:-1, 1:-1] == 0.0)
    for vx, vy in ((1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0)):
        out = np.zeros_like(u)
        orc.advection_step(u, out, 1, 1.0, 1.0, vx, vy, 0.1)
        assert out[ny // 2 + 1, nx // 2 + 1] != 0.0


def test_advection_accumulates(orc):
    """src/advection.cpp:31 — out += …, not out = …"""
    rng = np.random.default_rng(1)
    u = rng.standard_normal((7, 9))
    a = np.zeros_like(u)
    orc.advection_step(u, a, 1, 1.0, 1.0, 0.5, -0.25, 0.1)
    b = np.full_like(u, 3.0)
    orc.advection_step(u, b, 1, 1.0, 1.0, 0.5, -0.25, 0.1)
    assert bits_equal(b[1:-1, 1:-1], 3.0 + a[1:-1, 1:-1])


def test_boundary_dirichlet_and_neumann_single_rank(orc, oracle_mod):
    """tests/simulation/unit/test_boundary.cpp:9-69"""
    NX, NY, h = 4, 3, 1
    f = np.full((NY + 2, NX + 2), -1.0)
    f[1:-1, 1:-1] = 10.0
    none = (oracle_mod.PROC_NULL,) * 4
    orc.apply_boundary(f, h, none, (0, 0, 0, 0), 5.0)
    assert np.all(f[:, 0] == 5.0) and np.all(f[:, h + NX] == 5.0)
    assert np.all(f[0, :] == 5.0) and np.all(f[h + NY, :] == 5.0)
    f = np.full((NY + 2, NX + 2), -1.0)
    for j in range(h, h + NY):
        f[j, 1:-1] = float(j)
    orc.apply_boundary(f, h, none, (1, 1, 1, 1), 0.0)
    assert np.all(f[:, 0] == f[:, h]) and np.all(f[:, h + NX] == f[:, h + NX - 1])
    assert np.all(f[0, :] == f[h, :]) and np.all(f[h + NY, :] == f[h + NY - 1, :])


def test_boundary_periodic_is_a_noop_and_neighbours_skip(orc, oracle_mod):
    """src/boundary.cpp:23-53 has no Periodic branch; sides with a neighbour are untouched."""
    rng = np.random.default_rng(2)
    f0 = rng.standard_normal((6, 7))
    f = f0.copy()
    orc.apply_boundary(f, 1, (oracle_mod.PROC_NULL,) * 4, (2, 2, 2, 2), 9.0)
    assert bits_equal(f, f0)
    f = f0.copy()
    orc.apply_boundary(f, 1, (3, 4, 5, 6), (0, 1, 0, 1), 9.0)
    assert bits_equal(f, f0)


def test_stability(orc):
    """tests/simulation/unit/test_stability.cpp:5-27 + include/stability.hpp:5-16 values"""
    assert orc.safe_dt(1, 1, 0.5, 0.5, 0.1) > 0
    assert orc.safe_dt(1, 1, 5, 5, 0.1) < orc.safe_dt(1, 1, 0.5, 0.5, 0.1)
    assert orc.safe_dt(1, 1, 0.5, 0.5, 1.0) < orc.safe_dt(1, 1, 0.5, 0.5, 0.1)
    assert orc.safe_dt(1, 1, 0.5, 0.0, 0.05) == 2.0  # dev.yaml: min(2, 5)
    assert orc.safe_dt(1, 1, 0, 0, 0) == float("inf")


def test_field_layout_and_bounds(ref):
    """tests/simulation/unit/test_field.cpp:5-26"""
    assert ref.lib.ref_field_index(2, 2, 1, 3, 3) == 3 * 4 + 3
    assert ref.lib.ref_field_index(4, 3, 1, 0, 1) == 6
    f = (4, 4, 1)
    assert ref.lib.ref_field_at_throws(*f, -1, 0) == 1
    assert ref.lib.ref_field_at_throws(*f, 6, 0) == 1
    assert ref.lib.ref_field_at_throws(*f, 0, 6) == 1
    assert ref.lib.ref_field_at_throws(*f, 5, 5) == 0


def test_decomp_dims_and_neighbours(port, ref):
    """tests/simulation/unit/test_decomp_mpi.cpp:5-36 + src/decomp.cpp:13-33, port == reference"""
    for size in (1, 2, 3, 4, 6, 8, 12):
        a, b = port.decomp(size, 16, 12), ref.decomp(size, 16, 12)
        assert a == b, size
        for d in a:
            assert d["dims"][0] * d["dims"][1] == size and d["dims"][0] >= d["dims"][1]
            cx, cy = d["coords"]
            assert (d["nbr_lr"][0] == -1) == (cx == 0)
            assert (d["nbr_lr"][1] == -1) == (cx == d["dims"][0] - 1)
            assert (d["nbr_du"][0] == -1) == (cy == 0)
            assert (d["nbr_du"][1] == -1) == (cy == d["dims"][1] - 1)
    assert [port.decomp(n, 8, 8)[0]["dims"] for n in (1, 2, 4, 8)] == [(1, 1), (2, 1), (2, 2), (4, 2)]
    d = port.decomp(8, 130, 67)
    assert sum(x["nx_local"] for x in d if x["coords"][1] == 0) == 130
    assert sum(x["ny_local"] for x in d if x["coords"][0] == 0) == 67


def test_halo_adaptive_faces(ref):
    """tests/simulation/unit/test_halo.cpp:9-63: interior = rank id, ghosts = -1, 8x8 grid."""
    for size in (2, 4, 8):
        decs = ref.decomp(size, 8, 8)
        tiles = []
        for r, d in enumerate(decs):
            t = np.full((d["ny_local"] + 2, d["nx_local"] + 2), -1.0)
            t[1:-1, 1:-1] = float(r)
            tiles.append(t)
        ref.exchange(size, 8, 8, 1, tiles)
        for r, d in enumerate(decs):
            t = tiles[r]
            if d["nbr_lr"][0] != -1:
                assert np.all(t[1:-1, 0] == d["nbr_lr"][0])
            if d["nbr_lr"][1] != -1:
                assert np.all(t[1:-1, -1] == d["nbr_lr"][1])
            if d["nbr_du"][0] != -1:
                assert np.all(t[0, 1:-1] == d["nbr_du"][0])
            if d["nbr_du"][1] != -1:
                assert np.all(t[-1, 1:-1] == d["nbr_du"][1])


def test_physical_properties(oracle_mod, port):
    """integration_diffusion.cpp:8-47 (peak decreases, field >= 0) and integration_advection.cpp:8-35
    (centre of mass moves 5±1 cells in 5 steps at vx=1, dt=1; mass within 5 %)."""
    p = oracle_mod.SimParams(nx=64, ny=64, D=1.0, dt=0.1, steps=10, out_every=10)
    r = port.run(p)
    assert r["final"].max() < r["frames"][0].max() and r["final"].min() >= 0.0
    p = oracle_mod.SimParams(nx=64, ny=64, vx=1.0, dt=1.0, steps=5, out_every=5)
    r = port.run(p)
    xs = np.arange(64) + 0.5
    com0 = (r["frames"][0].sum(0) * xs).sum() / r["frames"][0].sum()
    com1 = (r["final"].sum(0) * xs).sum() / r["final"].sum()
    assert abs((com1 - com0) - 5.0) <= 1.0
    assert abs(r["final"].sum() / r["frames"][0].sum() - 1.0) <= 0.05


def test_periodic_equals_dirichlet_zero(oracle_mod, port):
    """with zero ghosts, Periodic is bit-identical to Dirichlet(0)."""
    a = port.run(oracle_mod.SimParams(nx=40, ny=40, D=0.05, vx=0.5, steps=50, out_every=50, bc=(2, 2, 2, 2)))
    b = port.run(oracle_mod.SimParams(nx=40, ny=40, D=0.05, vx=0.5, steps=50, out_every=50, bc=(0, 0, 0, 0)))
    assert bits_equal(a["final"], b["final"])


# ---- signed zeros and zero velocities: the regime behind the dropped-term kernels -----------------

def _signed_zero_tile(rng, ny, nx, negzero):
    a = rng.standard_normal((ny + 2, nx + 2)) * 10.0 ** rng.integers(-3, 3, (ny + 2, nx + 2))
    for _ in range(10):
        y, x = rng.integers(0, ny), rng.integers(0, nx)
        a[y:y + rng.integers(2, 9), x:x + rng.integers(2, 9)] = rng.choice([0.0, 1.5, -2.25, 1e-300])
    a[rng.random(a.shape) < 0.03] = 0.0
    if negzero:
        a[rng.random(a.shape) < 0.03] = -0.0
        a[ny // 2:ny // 2 + 6, nx // 2:nx // 2 + 6] = -0.0
    return a


def test_port_equals_reference_objects_on_signed_zero_fields(oracle_mod, port, ref):
    """Pins the port where the GPU's zero-velocity tests use it: flat patches, +0.0 / -0.0 cells,
    velocity components that are +0.0 or -0.0."""
    rng = np.random.default_rng(2024)
    for (vx, vy, D) in ((0.5, 0.0, 0.05), (0.0, -0.3, 0.0), (0.0, 0.0, 0.05), (0.5, -0.0, 0.05)):
        for negzero in (False, True):
            u0 = _signed_zero_tile(rng, 40, 70, negzero)
            sp = oracle_mod.SimParams(nx=70, ny=40, D=D, vx=vx, vy=vy, dt=0.1, steps=7, out_every=7, bc=(0, 1, 2, 1))
            a = port.run(sp, u0_padded=u0)["final"]
            b = ref.run(sp, u0_padded=u0)["final"]
            assert bits_equal(a, b), (vx, vy, D, negzero)


def _np_update(c, w, e, s, n, dtD, ndt, vx, vy, drop_x=False, drop_y=False):
    """numpy restatement of tb_update (MODE_UNIT; numpy never contracts into FMA; e - 2c has one
    rounding either way)."""
    o = c + dtD * (((e - 2.0 * c) + w) + ((n - 2.0 * c) + s))
    if drop_x and drop_y:
        return o
    px = vx * ((c - w) if vx >= 0 else (e - c))
    py = vy * ((c - s) if vy >= 0 else (n - c))
    adv = py if drop_x else (px if drop_y else px + py)
    return o + ndt * adv


def test_dropped_zero_velocity_term_is_exact_without_negative_zero():
    """The claim the dispatcher (kernels.cu, resolve_zero_terms) relies on, checked with IEEE arithmetic
    on the host: for finite cells without -0.0, dropping the term of a +0.0 velocity component does
    not change a single bit; with -0.0 cells it can (which is why such tiles take the full kernels)."""
    rng = np.random.default_rng(5)
    differs_with_negzero = False
    for negzero in (False, True):
        a = _signed_zero_tile(rng, 200, 300, negzero)
        c, w, e, s, n = a[1:-1, 1:-1], a[1:-1, :-2], a[1:-1, 2:], a[:-2, 1:-1], a[2:, 1:-1]
        for (vx, vy, D) in ((0.5, 0.0, 0.05), (-0.5, 0.0, 0.05), (0.0, 0.25, 0.05), (0.0, -0.25, 0.0), (0.0, 0.0, 0.05)):
            full = _np_update(c, w, e, s, n, 0.1 * D, -0.1, vx, vy)
            drop = _np_update(c, w, e, s, n, 0.1 * D, -0.1, vx, vy, drop_x=(vx == 0.0), drop_y=(vy == 0.0))
            if negzero:
                differs_with_negzero |= not bits_equal(full, drop)
                ok = np.signbit(c) & (c == 0.0)  # differences may only sit on -0.0 centre cells
                assert bits_equal(full[~ok], drop[~ok])
            else:
                assert bits_equal(full, drop), (vx, vy, D)
                assert not np.any(np.signbit(full) & (full == 0.0))  # and no -0.0 is ever created
    assert differs_with_negzero
