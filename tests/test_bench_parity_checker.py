"""bench.py's parity windows, checked on the CPU: the checker must pass on a correct field and must FAIL on a
wrong one.  The "GPU field" here is the oracle's own global field (decomposition-invariant, pinned in
test_oracle.py), cut into the tiles of a 1-, 2-, 4- and 8-rank decomposition exactly as bench.py sees them; then
single cells are corrupted — at a rank seam, at a decomposition corner, on the physical edge, in the interior — and
the checker has to notice each one.  No GPU: only host-side entry points of the library are used
(csim_decomp_init, csim_initial_condition_host)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("world,bc", [(1, (2, 2, 2, 2)), (2, (2, 2, 2, 2)), (4, (0, 1, 0, 1)), (8, (2, 2, 2, 2))])
def test_parity_windows_pass_on_the_oracle_and_catch_a_wrong_cell(csim, oracle_mod, port, bench, world, bc):
    tile, steps = 192, 9
    dims = csim.Decomp2D.init(world, 0, 1, 1).dims
    nxg, nyg = tile * dims[0], tile * dims[1]
    phys = dict(D=0.05, vx=-0.5, vy=0.25, dt=0.1)
    sp = oracle_mod.SimParams(nx=nxg, ny=nyg, steps=steps, out_every=steps, bc=bc, **phys)
    glob = port.run(sp)["final"]
    for rank in range(world):
        dec = csim.Decomp2D.init(world, rank, nxg, nyg)
        got = np.ascontiguousarray(glob[dec.y_offset:dec.y_offset + dec.ny_local, dec.x_offset:dec.x_offset + dec.nx_local])
        n, bad = bench.check_parity(csim, oracle_mod, port, got, dec, nxg, nyg, phys, bc, steps, w=32)
        assert n >= 8 and bad == 0, (world, rank, n, bad)
        # one wrong bit in a window at the tile's corner / edge / centre: the window holding it must differ
        pts, w = bench.parity_windows(dec, 32)
        for (y, x) in (pts[0], pts[len(pts) // 2], pts[-1]):
            broken = got.copy()
            cell = broken[y + w // 2:y + w // 2 + 1, x + w // 2:x + w // 2 + 1]
            cell.view(np.uint64)[...] ^= np.uint64(1)  # flip the last mantissa bit
            _, bad = bench.check_parity(csim, oracle_mod, port, broken, dec, nxg, nyg, phys, bc, steps, w=32)
            assert bad >= 1, (world, rank, y, x)


def test_parity_windows_cover_seams_and_corners(csim, bench):
    """Every rank checks its four tile corners (which on an inner rank are decomposition corners and on a
    boundary rank physical corners), the four edge midpoints and interior points: >= 8 distinct windows."""
    for world in (1, 2, 4, 8):
        dims = csim.Decomp2D.init(world, 0, 1, 1).dims
        for rank in range(world):
            dec = csim.Decomp2D.init(world, rank, 16384 * dims[0], 16384 * dims[1])
            pts, w = bench.parity_windows(dec, 64)
            assert len(pts) >= 8 and w == 64
            corners = {(0, 0), (0, dec.nx_local - w), (dec.ny_local - w, 0), (dec.ny_local - w, dec.nx_local - w)}
            assert corners <= set(pts)
            assert all(0 <= y <= dec.ny_local - w and 0 <= x <= dec.nx_local - w for (y, x) in pts)


def test_fp64_pipe_report_matches_the_profiled_kernel(csim, bench):
    """The FP64-pipe figure bench.py prints is a model (operations per update x re-computation factor of the
    library's own sweep plan); pin it to what is known from elsewhere: the operation counts of tb_update
    (11 with vy = 0 dropped, 14 with all terms, 7 with both dropped), the SASS count of the fast loop
    (profiles/r02_tuning.md: 704 FP64 instructions per two ticks = 4 rows x 4 levels x 4 cells x 11) and the
    ncu capture at 16384^2 (FP64 pipe 77.7 % of active cycles at ~1.06e12 cell-updates/s under the power cap)."""
    assert bench.fp64_ops_per_cell(0.5, 0.0, True) == 11
    assert bench.fp64_ops_per_cell(0.5, 0.0, False) == 14      # not dropped: the term is computed
    assert bench.fp64_ops_per_cell(-0.5, 0.25, False) == 14
    assert bench.fp64_ops_per_cell(0.0, 0.0, True) == 7
    assert 4 * 4 * 4 * bench.fp64_ops_per_cell(0.5, 0.0, True) == 704
    f = bench.computed_over_useful(csim, 16384, 16384, 4, (-1, -1, -1, -1))
    assert 1.09 < f < 1.105                                     # "9.6 % of the cell updates are re-computed"
    assert bench.computed_over_useful(csim, 8192, 8192, 4, (-1, -1, -1, -1)) > f   # shorter chunks, more overlap
    # with neighbours on every side the sweep stores no ghost line but computes the same strips
    assert abs(bench.computed_over_useful(csim, 16384, 16384, 4, (1, 2, 3, 4)) - f) < 0.01
    rep = bench.fp64_pipe_report(1.065e12, 11, f, 148, 1810.0, 1965)
    assert 0.70 < rep["frac_at_sampled_clock"] < 0.80           # ncu: 77.7 % of active cycles
    assert rep["frac_at_max_clock"] < rep["frac_at_sampled_clock"] <= 1.0
    assert bench.fp64_pipe_report(1e12, 11, f, 148, None, None)["frac_at_sampled_clock"] is None


def test_reference_arm_prints_the_same_config_keys(bench, capsys):
    """Both arms of bench.py print the same `config` key set (the driver compares the two lines); the GPU arm's
    keys are read off its source, the reference arm is run on a small tile."""
    import argparse
    import json
    import re
    args = argparse.Namespace(gpus=1, steps=1, warmup=0, tile=96, inner=100, ref_inner=1, bc="periodic")
    assert bench.run_reference(args) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
    src = open(os.path.join(ROOT, "bench.py")).read()
    ours = src[src.index("def run_ours"):]
    block = ours[ours.index("cfg.update({"):]
    block = block[:block.index("})")]
    gpu_keys = set(re.findall(r'^\s*"([a-z_0-9]+)":', block, flags=re.M)) | set(bench.base_config(args, (1, 1), 100))
    assert gpu_keys == set(line["config"]), (gpu_keys ^ set(line["config"]))
    assert line["config"]["workload"] == bench.workload_name(96, (1, 1))  # N = 1: the very same grid
