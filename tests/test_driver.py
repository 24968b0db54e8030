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


def _spec_header(fmt, nx, ny, nrec, attrs):
    """The header of a classic NetCDF file with dims time(unlimited), y, x and one NC_DOUBLE record variable
    u(time,y,x), assembled here from the format grammar (netcdf classic format spec: NON_NEG and name
    lengths are 32-bit in CDF-1/2 and 64-bit in CDF-5, OFFSET is 64-bit in CDF-2 and CDF-5, everything is
    big-endian, names and attribute values are padded to 4 bytes)."""
    import struct
    cnt = (lambda v: struct.pack(">q", v)) if fmt == 5 else (lambda v: struct.pack(">I", v))
    pad = lambda b: b + b"\0" * (-len(b) % 4)  # noqa: E731
    name = lambda s: cnt(len(s)) + pad(s.encode())  # noqa: E731
    h = b"CDF" + bytes([fmt]) + cnt(nrec)
    h += struct.pack(">I", 0x0A) + cnt(3) + name("time") + cnt(0) + name("y") + cnt(ny) + name("x") + cnt(nx)
    h += struct.pack(">I", 0x0C) + cnt(len(attrs))
    for k, v in attrs:
        h += name(k) + struct.pack(">I", 2) + cnt(len(v)) + pad(v.encode())
    h += struct.pack(">I", 0x0B) + cnt(1) + name("u") + cnt(3) + cnt(0) + cnt(1) + cnt(2)
    h += struct.pack(">I", 0) + cnt(0) + struct.pack(">I", 6) + cnt(nx * ny * 8)
    return h  # + OFFSET begin (8 bytes)


@pytest.mark.parametrize("nx,ny,nrec", [(3, 2, 2), (37, 11, 3), (1, 1, 1), (64, 5, 0)])
def test_file_header_against_the_format_grammar_and_scipy(tmp_path, nx, ny, nrec):
    """No GPU: `test_driver --write-file` builds a whole file from the writer's own header builder.
    (a) CDF-5 and CDF-2 headers equal, byte for byte, a header assembled from the format grammar; the data
    offset that follows is 8 bytes, aligned, and inside the file.  (b) The CDF-2 flavour — the same builder,
    32-bit counts — is opened by scipy.io.netcdf_file, a reader this repo did not write: dimensions,
    unlimited record dimension, attributes, dtype, shape and every value agree.  (c) The in-repo CDF-5
    reader returns the same array from the CDF-5 flavour."""
    import struct
    from scipy.io import netcdf_file
    attrs = [("description", "climate-sim-mpi-cpp"), ("grid", f"{nx} x {ny}"), ("dt", "0.100000"), ("odd", "abcde")]
    want = (1e6 * np.arange(nrec)[:, None, None] + 1e3 * np.arange(ny)[None, :, None] + np.arange(nx)[None, None, :]
            + 0.25)
    for fmt in (5, 2):
        path = str(tmp_path / f"f{fmt}.nc")
        r = subprocess.run([_need("test_driver"), "--write-file", path, str(fmt), str(nx), str(ny), str(nrec)],
                           capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = open(path, "rb").read()
        head = _spec_header(fmt, nx, ny, nrec, attrs)
        assert raw[:len(head)] == head, fmt
        begin = struct.unpack(">q", raw[len(head):len(head) + 8])[0]
        assert begin % 4 == 0 and begin >= len(head) + 8 and len(raw) == begin + nrec * nx * ny * 8
        assert not any(raw[len(head) + 8:begin])  # padding up to the data is zero
    with netcdf_file(str(tmp_path / "f2.nc"), "r", mmap=False) as f:
        assert f.version_byte == 2
        assert list(f.dimensions.items()) == [("time", None), ("y", ny), ("x", nx)]
        assert {k: getattr(f, k).decode() for k, _ in attrs} == dict(attrs)
        u = f.variables["u"]
        assert u.dimensions == ("time", "y", "x") and u.isrec and u.data.dtype == np.dtype(">f8")
        assert u.shape == (nrec, ny, nx)
        assert bits_equal(np.array(u.data, dtype=np.float64), want)
    f5 = read_cdf5(str(tmp_path / "f5.nc"))
    assert f5["numrecs"] == nrec and bits_equal(f5["data"], want) and f5["attrs"] == dict(attrs)


@pytest.mark.gpu
def test_driver_cdf2_output_is_read_by_scipy(tmp_path, oracle_mod, port):
    """The driver with CSIM_NETCDF_FORMAT=cdf2: scipy.io.netcdf_file (not written here) reads the frames the
    GPU produced, and they equal the oracle's, bit for bit; the default CDF-5 file of the same run holds the
    same bytes after its own header."""
    from scipy.io import netcdf_file
    args = ["--nx=96", "--ny=80", "--D=0.05", "--vx=0.5", "--vy=-0.25", "--steps=12", "--out_every=4",
            "--bc.left=neumann", "--bc.top=periodic"]
    os.environ["CSIM_NETCDF_FORMAT"] = "cdf2"
    try:
        _run_driver(tmp_path, args)
    finally:
        del os.environ["CSIM_NETCDF_FORMAT"]
    p = oracle_mod.SimParams(nx=96, ny=80, D=0.05, vx=0.5, vy=-0.25, steps=12, out_every=4, bc=(1, 0, 0, 2))
    want = port.run(p)["frames"]
    with netcdf_file(str(tmp_path / "outputs" / "snapshots.nc"), "r", mmap=False) as f:
        assert f.version_byte == 2 and f.variables["u"].shape == (3, 80, 96)
        assert f.grid.decode() == "96 x 80" and f.steps.decode() == "12"
        got2 = np.array(f.variables["u"].data, dtype=np.float64)
    assert bits_equal(got2, want)
    _run_driver(tmp_path, args)
    assert bits_equal(read_cdf5(tmp_path / "outputs" / "snapshots.nc")["data"], want)


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
    disjoint windows of the same CDF-5 file; halos as packed bands over NCCL."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    args = ["--nx=301", "--ny=260", "--D=0.05", "--vx=-0.5", "--vy=0.25", "--steps=50", "--out_every=10",
            "--bc.right=neumann", "--bc.bottom=periodic"]
    _run_driver(tmp_path, args, ranks=2)
    f = read_cdf5(tmp_path / "outputs" / "snapshots.nc")
    p = oracle_mod.SimParams(nx=301, ny=260, D=0.05, vx=-0.5, vy=0.25, steps=50, out_every=10, bc=(0, 1, 2, 0))
    assert f["numrecs"] == 5 and bits_equal(f["data"], port.run(p)["frames"])
