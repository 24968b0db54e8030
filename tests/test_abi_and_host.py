"""CPU-only checks of the product library: the C ABI exports what include/csim.h declares, the
host-side functions (decomposition, stability limit, initial condition, BC names) agree with the
oracle, and compute entry points fail loudly without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, bits_equal


def declared_functions():
    text = open(os.path.join(ROOT, "include", "csim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csim_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(csim):
    names = declared_functions()
    assert len(names) >= 30
    L = ctypes.CDLL(csim.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert csim.lib().csim_abi_version() == 1


def test_decomp_matches_oracle(csim, port):
    for size in (1, 2, 3, 4, 6, 8, 12, 16):
        for (nxg, nyg) in ((16, 12), (130, 67), (8192, 8192)):
            o = port.decomp(size, nxg, nyg)
            for r in range(size):
                d = csim.Decomp2D.init(size, r, nxg, nyg)
                assert d.dims == o[r]["dims"] and d.coords == o[r]["coords"]
                assert d.nbr_lr == o[r]["nbr_lr"] and d.nbr_du == o[r]["nbr_du"]
                assert (d.nx_local, d.ny_local, d.x_offset, d.y_offset) == (
                    o[r]["nx_local"], o[r]["ny_local"], o[r]["x_offset"], o[r]["y_offset"])
    with pytest.raises(csim.CsimError):
        csim.Decomp2D.init(4, 4, 8, 8)


def test_safe_dt_matches_oracle(csim, port):
    rng = np.random.default_rng(3)
    for _ in range(200):
        dx, dy = rng.uniform(0.1, 3, 2)
        vx, vy = rng.uniform(-2, 2, 2) * (rng.random(2) > 0.2)
        D = rng.uniform(0, 2) * (rng.random() > 0.2)
        assert csim.safe_dt(dx, dy, vx, vy, D) == port.safe_dt(dx, dy, vx, vy, D)
    assert csim.safe_dt(1, 1, 0.5, 0.0, 0.05) == 2.0


def test_initial_condition_matches_reference_bits(csim, oracle_mod, port):
    """src/init.cpp:12-33 through the product's host IC == oracle frame 0, per rank tile."""
    for (nxg, nyg, size) in ((64, 64, 1), (50, 35, 6), (512, 512, 4)):
        p = oracle_mod.SimParams(nx=nxg, ny=nyg, steps=1, out_every=1, sigma_frac=0.07, xc_frac=0.4,
                                 yc_frac=0.6, A=1.5, dx=0.5, dy=2.0)
        frame0 = port.run(p, nranks=1)["frames"][0]
        for r in range(size):
            d = csim.Decomp2D.init(size, r, nxg, nyg)
            t = csim.initial_condition_host(d, 1, p.dx, p.dy, "gaussian_hotspot", p.A, p.sigma_frac,
                                            p.xc_frac, p.yc_frac)
            assert np.all(t[0] == 0) and np.all(t[:, 0] == 0)  # ghosts untouched
            want = frame0[d.y_offset:d.y_offset + d.ny_local, d.x_offset:d.x_offset + d.nx_local]
            assert bits_equal(t[1:-1, 1:-1], want)
    z = csim.initial_condition_host(csim.Decomp2D.single(8, 8), 1, 1.0, 1.0, "constant_zero")
    assert not z.any()
    with pytest.raises(RuntimeError, match="Unknown IC preset"):
        csim.initial_condition_host(csim.Decomp2D.single(8, 8), 1, 1.0, 1.0, "checkerboard")


def test_bc_strings(csim):
    """src/io.cpp:35-56"""
    B = csim.BCType
    assert csim.bc_from_string("Dirichlet") == B.Dirichlet and csim.bc_from_string("fixed") == B.Dirichlet
    assert csim.bc_from_string("NOFLUX") == B.Neumann and csim.bc_from_string("zero-flux") == B.Neumann
    assert csim.bc_from_string("period") == B.Periodic
    assert [csim.bc_to_string(b) for b in B] == ["dirichlet", "neumann", "periodic"]
    with pytest.raises(RuntimeError, match="Unknown BC type"):
        csim.bc_from_string("robin")


def test_no_cpu_fallback(csim):
    """Without a device the product refuses to run; it never routes through oracle/."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(csim.CsimError, match="no CUDA device"):
        csim.Context(0)


def test_product_never_imports_oracle():
    """No source file of the product package imports, includes, links or dlopens anything of oracle/."""
    pkg = os.path.join(ROOT, "climate-sim-mpi-cpp_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|cpu_oracle|liboracle|libcsim_ref|#include\s*[<\"].*oracle|oracle_port)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(text), f


# ---- geometry of the fused sweep (csim_sweep_plan, host only) -------------------------------------

def _cover(items, x_lo, x_hi, y_lo, y_hi):
    """How often each cell of [x_lo,x_hi) x [y_lo,y_hi) is covered by the items' rectangles."""
    cnt = np.zeros((y_hi - y_lo, x_hi - x_lo), dtype=np.int32)
    for (_, x0, x1, y0, y1) in items:
        assert x_lo <= x0 < x1 <= x_hi and y_lo <= y0 < y1 <= y_hi, (x0, x1, y0, y1)
        cnt[y0 - y_lo:y1 - y_lo, x0 - x_lo:x1 - x_lo] += 1
    return cnt


@pytest.mark.parametrize("nx,ny", [(1, 1), (7, 300), (120, 33), (121, 97), (241, 193), (242, 4099), (513, 257),
                                   (1000, 1000), (8192, 8192), (8192, 16384), (16384, 8191)])
def test_sweep_plan_tiles_the_stored_region_exactly_once(csim, nx, ny):
    """Every stored cell (interior plus the ghost line of each physical side) belongs to exactly one work
    item — also with the shorter chunks in the tail of the launch, for every blocking depth, for tiles
    with and without neighbours — and the interior / frame launches of the multi-GPU loop partition
    the items of the whole sweep."""
    big = nx * ny >= 4_000_000  # the full-size tiles: default depth and machine only (keeps the suite short)
    for T in ((3,) if big else (1, 2, 3, 4)):
        for nbr in ((-1, -1, -1, -1), (3, -1, -1, 5), (-1, 2, 7, -1), (1, 2, 3, 4)):
            for slots in ((1776,) if big else (0, 64, 1776)):
                x_lo, x_hi = (-1 if nbr[0] < 0 else 0), (nx + 1 if nbr[1] < 0 else nx)
                y_lo, y_hi = (-1 if nbr[2] < 0 else 0), (ny + 1 if nbr[3] < 0 else ny)
                whole = csim.sweep_plan(nx, ny, T, nbr, slots, 0)
                assert (_cover(whole, x_lo, x_hi, y_lo, y_hi) == 1).all(), (T, nbr, slots)
                inner = csim.sweep_plan(nx, ny, T, nbr, slots, 1)
                frame = csim.sweep_plan(nx, ny, T, nbr, slots, 2)
                assert sorted(inner + frame) == sorted(whole), (T, nbr, slots)
                # the coupled launch of the multi-GPU loop: the frame's items first, then the interior's
                assert csim.sweep_plan(nx, ny, T, nbr, slots, 3) == frame + inner, (T, nbr, slots)
                # interior items never touch the first/last chunk of a strip nor the edge strips: they
                # read no ghost line and may run while the halos travel
                nstrips = max(s for (s, *_r) in whole) + 1
                for (s, x0, x1, y0, y1) in inner:
                    assert 0 < s < nstrips - 1 and y0 > y_lo and y1 < y_hi
                    # its dependency cone (T rows above and below, T columns beside) stays clear of
                    # the ghost lines of every side that has a neighbour (those lines are in flight)
                    assert nbr[2] < 0 or y0 - T >= 0, (T, nbr, slots, y0)
                    assert nbr[3] < 0 or y1 + T <= ny, (T, nbr, slots, y1)
                    assert nbr[0] < 0 or x0 - T >= 0, (T, nbr, slots, x0)
                    assert nbr[1] < 0 or x1 + T <= nx, (T, nbr, slots, x1)
                if nx * ny >= 8192 * 8192 and slots:
                    assert len(whole) >= slots  # enough items to fill the machine at least once


def test_sweep_items_load_only_rows_inside_the_allocation(csim):
    """Load model of k_step_tb (step_tb.cuh): a work item [y0, y1) starts with level-0 rows y0-T and
    y0-T+1, and every tick at row r (two per loop iteration, r advancing by 2) requests rows r+2 and r+3
    iff r+2 < y1+T.  Every row it touches must lie inside the kLeadY = 8 rows of padding below and
    above the tile (csim_field_create allocates ny + 16 rows) — also at T = 4, with physical top and
    bottom sides (stored range -1 … ny) and with chunk heights that are 3 mod 4."""
    lead = 8
    rng = np.random.default_rng(2718)
    sizes = [(1, 1), (5, 3), (256, 256), (512, 512), (1024, 1024), (4096, 4096), (16352, 16384), (130, 4099)]
    sizes += [(int(rng.integers(1, 600)), int(rng.integers(1, 3000))) for _ in range(40)]
    for (nx, ny) in sizes:
        for T in (1, 2, 3, 4):
            for nbr in ((-1, -1, -1, -1), (1, 2, 3, 4), (-1, 2, -1, 4)):
                for (_s, _x0, _x1, y0, y1) in csim.sweep_plan(nx, ny, T, nbr, 1776, 0):
                    r_end = y1 + T
                    lo, hi = y0 - T, y0 - T + 1
                    r = y0 - T
                    while r < r_end:          # for (; r < r_end; r += 4) { tick(r); tick(r + 2); }
                        for rr in (r, r + 2):
                            if rr + 2 < r_end:
                                hi = max(hi, rr + 3)
                        r += 4
                    assert -lead <= lo and hi < ny + lead, (nx, ny, T, nbr, y0, y1, lo, hi)


def test_sweep_plan_rejects_bad_arguments(csim):
    for args in ((0, 5, 3), (5, 0, 3), (5, 5, 0), (5, 5, 5)):
        with pytest.raises(csim.CsimError):
            csim.sweep_plan(*args)


def test_sweep_plan_random_geometries(csim):
    """The same three properties (exact tiling, interior + frame = whole, interior items clear of in-flight
    ghost lines) over a few hundred random tile sizes, depths, neighbour sets and machine sizes."""
    rng = np.random.default_rng(314159)
    for _ in range(300):
        nx = int(rng.choice([rng.integers(1, 40), rng.integers(100, 400), rng.integers(400, 2600)]))
        ny = int(rng.choice([rng.integers(1, 40), rng.integers(30, 300), rng.integers(300, 2600)]))
        T = int(rng.integers(1, 5))
        nbr = tuple(int(v) for v in rng.choice([-1, 1], size=4))
        slots = int(rng.choice([0, 8, 96, 512, 1776, 2368]))
        x_lo, x_hi = (-1 if nbr[0] < 0 else 0), (nx + 1 if nbr[1] < 0 else nx)
        y_lo, y_hi = (-1 if nbr[2] < 0 else 0), (ny + 1 if nbr[3] < 0 else ny)
        whole = csim.sweep_plan(nx, ny, T, nbr, slots, 0)
        assert (_cover(whole, x_lo, x_hi, y_lo, y_hi) == 1).all(), (nx, ny, T, nbr, slots)
        inner = csim.sweep_plan(nx, ny, T, nbr, slots, 1)
        frame = csim.sweep_plan(nx, ny, T, nbr, slots, 2)
        assert sorted(inner + frame) == sorted(whole), (nx, ny, T, nbr, slots)
        assert csim.sweep_plan(nx, ny, T, nbr, slots, 3) == frame + inner, (nx, ny, T, nbr, slots)
        for (s, x0, x1, y0, y1) in inner:
            assert nbr[2] < 0 or y0 - T >= 0, (nx, ny, T, nbr, slots, y0)
            assert nbr[3] < 0 or y1 + T <= ny, (nx, ny, T, nbr, slots, y1)
            assert nbr[0] < 0 or x0 - T >= 0, (nx, ny, T, nbr, slots, x0)
            assert nbr[1] < 0 or x1 + T <= nx, (nx, ny, T, nbr, slots, x1)
