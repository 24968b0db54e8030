"""bench.py's GPU arm, walked end to end on the CPU with stand-ins for the device.

Purpose: every statement of `run_ours` (timing loop, all-terms windows, the one-simulation e2e with its
double-buffered frames, parity windows, roofline and FP64-pipe report, the JSON line) is executed here, so a
slip in the reporting code cannot surface for the first time on the GPU box.  The stand-ins replace exactly what
needs a device: `Context`, `Field`, `run_steps` (backed by the CPU oracle — this is a test, the one place that
may do so) and the handful of `torch.cuda` calls; everything host-side (decomposition, initial condition, step
parameters, sweep plan) is the real library.  Nothing here measures anything: the numbers in the line are
meaningless, only its shape and the control flow are checked."""
import argparse
import importlib.util
import json
import os
import sys
import time
import types

import numpy as np
import pytest

from conftest import ROOT

PKG = "climate-sim-mpi-cpp_b200"


class FakeContext:
    def __init__(self, device=0):
        self.device, self.launch_count, self.stream_ptr = device, 0, 0

    def bind_numa(self):
        return -1

    def pinned_empty(self, shape):
        return np.empty(shape, dtype=np.float64)

    def sync(self):
        pass

    def event_wait(self, event):
        assert event is not None

    def close(self):
        pass


class FakeField:
    def __init__(self, ctx, nx, ny, h, dx, dy):
        assert h == 1
        self.ctx, self.nx, self.ny, self.dx, self.dy = ctx, nx, ny, dx, dy
        self.data = np.zeros((ny + 2, nx + 2))
        self.value_state = 1  # "clean": what the library reports for the Gaussian after its scan

    def upload(self, host):
        self.data[...] = host

    upload_async = upload

    def download_interior_async(self, out):
        out[...] = self.data[1:-1, 1:-1]

    def snapshot_async(self, out, big_endian=False):
        out[...] = self.data[1:-1, 1:-1]
        return object()

    def health(self):
        inner = self.data[1:-1, 1:-1]
        return float(np.abs(inner).max()), int((~np.isfinite(inner)).sum())

    def close(self):
        pass


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, 1e3 * (other.t - self.t))


@pytest.fixture()
def dry_bench(csim, oracle_mod, port, monkeypatch):
    """bench.py loaded as a module, with the device stood in for."""
    import torch

    fake = types.ModuleType(PKG)
    for name in dir(csim):
        if not name.startswith("__"):
            setattr(fake, name, getattr(csim, name))
    fake.Context, fake.Field = FakeContext, FakeField

    def run_steps(u, tmp, p, dec, nsteps):
        sp = oracle_mod.SimParams(nx=u.nx, ny=u.ny, dx=u.dx, dy=u.dy, D=p.D, vx=p.vx, vy=p.vy, dt=p.dt, steps=nsteps,
                                  out_every=nsteps, bc=tuple(int(b) for b in p.bc))
        u.data[1:-1, 1:-1] = port.run(sp, u0_padded=u.data)["final"]  # periodic sides: the ghost ring stays frozen
        T = csim.steps_per_sweep()
        u.ctx.launch_count += (nsteps + T - 1) // T

    fake.run_steps = run_steps
    monkeypatch.setitem(sys.modules, PKG, fake)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda *a, **k: object())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda d: types.SimpleNamespace(multi_processor_count=148))
    for var in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    spec = importlib.util.spec_from_file_location("bench_dry", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _args(**kw):
    base = dict(gpus=1, steps=4, warmup=1, impl="ours", tile=96, inner=8, ref_inner=1, no_cpu_baseline=True,
                no_e2e=False, no_parity=False, parity_steps=10, global_size=0, dx=1.0, dy=1.0, bc="periodic")
    base.update(kw)
    return argparse.Namespace(**base)


def test_gpu_arm_walks_through_on_stand_ins(dry_bench, capsys):
    mask = os.sched_getaffinity(0)
    assert dry_bench.run_ours(_args()) == 0
    assert os.sched_getaffinity(0) == mask  # the CPU mask is back after the NUMA-local allocations
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(out) == 1, out  # ONE JSON line
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "parity", "all_terms",
                "gpu_launches", "clocks"):
        assert key in line, key
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 4 and line["dtype"] == "f64"
    assert line["config"]["workload"].startswith("96x96 per GPU") and line["config"]["timesteps_per_step"] == 8
    T = 4
    assert line["gpu_launches"] == 4 * (8 // T)  # only the sweeps of the timed windows
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["steps_per_launch"] == T and rf["frac"] == rf["achieved"] / rf["peak"]
    assert rf["fp64_pipe"]["ops_per_cell_update"] == 11 and rf["fp64_pipe"]["computed_over_useful_cells"] > 1.0
    assert rf["fp64_pipe"]["sms"] == 148 and "error" not in rf["fp64_pipe"]
    assert line["all_terms"]["fp64_pipe"]["ops_per_cell_update"] == 14 and line["all_terms"]["steps"] == 4
    e2e = line["e2e"]
    assert e2e["steps"] == 4 and e2e["value"] > 0  # max(3, min(steps, 10)) windows
    assert e2e["d2h_bytes_per_step"] == 96 * 96 * 8 and e2e["h2d_bytes_per_step"] == 98 * 98 * 8 // 4
    assert e2e["per_window_copies"]["h2d_bytes_per_step"] == 98 * 98 * 8
    par = line["parity"]
    assert par["bit_identical"] is True and par["windows"] >= 16 and par["steps"] == 10
    assert par["parameter_sets"] == ["headline", "all_terms"]
    assert line["halo"] is None and line["shared_file"] is None and line["cpu_baseline"] is None


def test_gpu_arm_default_step_count_gives_ten_e2e_windows(dry_bench, capsys):
    """The driver runs --steps 20: the e2e simulation then has dev.yaml's ten output windows."""
    assert dry_bench.run_ours(_args(steps=20, inner=4, no_parity=True)) == 0
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][0])
    assert line["e2e"]["steps"] == 10 and line["parity"] is None
    assert line["e2e"]["h2d_bytes_per_step"] == 98 * 98 * 8 // 10


def test_gpu_arm_exits_non_zero_on_a_parity_mismatch(dry_bench, capsys, monkeypatch):
    """A field that differs from the oracle in one bit must end the run with a non-zero status."""
    fake = sys.modules[PKG]
    good = fake.run_steps

    def broken(u, tmp, p, dec, nsteps):
        good(u, tmp, p, dec, nsteps)
        if nsteps == 10:  # the parity run
            u.data[1:2, 1:2].view(np.uint64)[...] ^= np.uint64(1)  # the tile's corner cell: inside a window

    monkeypatch.setattr(fake, "run_steps", broken)
    with pytest.raises(SystemExit) as exc:
        dry_bench.run_ours(_args(no_e2e=True))
    assert exc.value.code == 3
    out = capsys.readouterr().out
    assert "GPU field differs from the oracle" in out and '"value"' not in out


def test_gpu_arm_refuses_to_run_without_a_device(csim, monkeypatch):
    """No CPU fallback: without CUDA the GPU arm stops (the real torch says no device here)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    spec = importlib.util.spec_from_file_location("bench_nodev", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for var in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    with pytest.raises(SystemExit) as exc:
        mod.run_ours(_args())
    assert "no CUDA device" in str(exc.value)
