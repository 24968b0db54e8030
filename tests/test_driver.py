"""The callers either side of the path: config reader, driver, CDF-5 snapshots."""
import os
import re
import subprocess

import numpy as np
import pytest

from cdf5_reader import read_cdf5
from conftest import ROOT, bits_equal

BUILD = os.path.join(ROOT, "climate-sim-mpi-cpp_b200", "host", "build")


def _need(exe):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    assert os.path.exists(path), path
    return path


def test_config_reader_cpp(tmp_path):
    """host/tests/test_driver.cpp without --gpu: the reference's test_io.cpp config tests (YAML nested /
    flat / scalar bc, CLI precedence, space-separated values, aliases, validate() errors, quirks Q6/Q12)."""
    r = subprocess.run([_need("test_driver")], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_snapshot_writer_round_trip_cpp(tmp_path):
    r = subprocess.run([_need("test_driver"), "--gpu"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr


def _run_driver(tmp_path, args, ranks=1):
    exe = _need("climate_sim_b200")
    procs = []
    for r in range(ranks):
        env = dict(os.environ)
        if ranks > 1:
            env.update(RANK=str(r), WORLD_SIZE=str(ranks), LOCAL_RANK=str(r),
                       CSIM_RENDEZVOUS=str(tmp_path / "rendezvous"))
        procs.append(subprocess.Popen([exe] + args, cwd=tmp_path, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o + e
    return outs[0][0], outs[0][1]


@pytest.mark.gpu
def test_driver_dev_yaml_matches_oracle(tmp_path, oracle_mod, port):
    """configs[0] end to end: the driver binary with configs/dev.yaml (shortened to 300 steps) writes
    outputs/snapshots.nc; every frame equals the oracle's frame bit for bit; stdout keeps the
    reference's contract (banner, IC min/max, `timing: total_max=… s, worst_avg_step=… s`)."""
    out, err = _run_driver(tmp_path, [f"--config={ROOT}/configs/dev.yaml", "--steps=300"])
    assert out.startswith("climate-sim-mpi-cpp \n  grid: 512 x 512  dt: 0.1  steps: 300  D: 0.05  v=(0.5,0)\n"
                          "  bc: left=dirichlet right=neumann bottom=periodic top=dirichlet\n"), out
    assert "IC min/max: 0 / " in out and "Opening NetCDF file for parallel output\n" in out
    m = re.search(r"timing: total_max=([0-9.e+-]+) s, worst_avg_step=([0-9.e+-]+) s\n$", out)
    assert m and float(m.group(1)) > 0
    f = read_cdf5(tmp_path / "outputs" / "snapshots.nc")
    assert f["numrecs"] == 3 and [d[0] for d in f["dims"]] == ["time", "y", "x"] and f["var"] == "u"
    assert f["attrs"]["grid"] == "512 x 512" and f["attrs"]["dt"] == "0.100000" and f["attrs"]["steps"] == "300"
    assert f["attrs"]["boundary_conditions"] == "left=dirichlet right=neumann bottom=periodic top=dirichlet"
    p = oracle_mod.SimParams(**oracle_mod.DEV_YAML)
    p.steps = 300
    want = port.run(p)["frames"]
    assert bits_equal(f["data"], want)


@pytest.mark.gpu
def test_driver_clamps_dt_and_cli_overrides(tmp_path, oracle_mod, port):
    """dt above the stability limit is clamped with the reference's warning (main.cpp:42-49)."""
    out, err = _run_driver(tmp_path, ["--nx=96", "--ny=80", "--D=1.0", "--vx=2.0", "--dt=5.0", "--steps=7",
                                      "--out_every", "3", "--bc.left=neumann", "--bc.top", "periodic"])
    assert "[warn] dt=5 exceeds stability limit 0.25 -> clamping to dt=0.25" in err
    f = read_cdf5(tmp_path / "outputs" / "snapshots.nc")
    assert f["numrecs"] == 3  # n = 0, 3, 6
    p = oracle_mod.SimParams(nx=96, ny=80, D=1.0, vx=2.0, dt=5.0, steps=7, out_every=3, bc=(1, 0, 0, 2))
    assert bits_equal(f["data"], port.run(p)["frames"])


@pytest.mark.gpu
def test_driver_two_ranks_one_file(tmp_path, oracle_mod, port):
    """Two processes (RANK/WORLD_SIZE + rendezvous file, no MPI launcher), one GPU each, writing
    disjoint windows of the same CDF-5 file; halos over peer memory."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    args = ["--nx=301", "--ny=260", "--D=0.05", "--vx=-0.5", "--vy=0.25", "--steps=50", "--out_every=10",
            "--bc.right=neumann", "--bc.bottom=periodic"]
    _run_driver(tmp_path, args, ranks=2)
    f = read_cdf5(tmp_path / "outputs" / "snapshots.nc")
    p = oracle_mod.SimParams(nx=301, ny=260, D=0.05, vx=-0.5, vy=0.25, steps=50, out_every=10, bc=(0, 1, 2, 0))
    assert f["numrecs"] == 5 and bits_equal(f["data"], port.run(p)["frames"])
