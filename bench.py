#!/usr/bin/env python
"""bench.py — cell-updates/s of the fused diffusion+advection timestep on N B200s.

    python bench.py --gpus 1 --steps K --warmup W              # our CUDA path (default)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W   # one rank per GPU
    python bench.py --impl reference ...                          # the reference's CPU code

Workload (BASELINE.json configs[2], the tile north_star's target is stated on; per-GPU tile fixed as N
grows → weak scaling): a 16384x16384 tile per GPU, Gaussian hotspot, dx=dy=1, D=0.05, vx=0.5, vy=0,
dt=0.1, all-periodic boundaries (= frozen zero ghosts), 2-D Cartesian decomposition
{1,1},{2,1},{2,2},{4,2}.  `--tile 8192` gives configs[1].

A bench "step" is one output window of `--inner` (default 100, dev.yaml's out_every) time steps:
  value : cells * inner * K / device time, fields resident in HBM, timed with CUDA events on the
          library's stream, max over ranks.
  e2e   : ONE simulation through the C ABI with HOST buffers, as the reference's main() runs it
          (src/main.cpp:71,93-109): the initial tile goes up from pinned memory once, then every window
          is `inner` time steps followed by the device→host copy of the de-haloed tile (what
          write_field_netcdf hands to the file layer), the copy of window k overlapping the steps of
          window k+1.  The upload and every download are inside the timed region.
  parity: after the timing, K time steps from the initial condition are run again and >= 8 windows of
          every rank's tile (rank seams, tile corners, physical edges, deep interior) are compared bit for
          bit with the CPU oracle run on the sub-domain around each window; for the headline physics
          and for the all-terms physics.  A mismatch exits non-zero.
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PHYS = dict(D=0.05, vx=0.5, vy=0.0, dt=0.1)          # configs/dev.yaml physics
PHYS_ALL_TERMS = dict(D=0.05, vx=-0.5, vy=0.25, dt=0.1)  # both velocity components, forward difference in x
ALG_BYTES_PER_CELL = 16.0                              # one 8-byte read + one 8-byte write per cell update
NVLINK_GBS_PER_DIRECTION = 900.0                       # NVLink 5, per GPU and direction (B200_PROFILING.md)
METRIC = "cell updates/sec (diffusion+advection step)"


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(2.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def dims_for(n):
    """MPI_Dims_create(n, 2): most square factorisation, non-increasing (src/decomp.cpp:13)."""
    b = max(d for d in range(1, int(n ** 0.5) + 1) if n % d == 0)
    return (n // b, b)


FP64_LANES_PER_SM_PER_CLK = 64  # B200: one FP64 instruction per lane slot, 64 slots per SM and clock


def fp64_ops_per_cell(vx, vy, dropped):
    """FP64 instructions one cell update of the blocked sweep issues with unit spacing (csrc/step_tb.cuh,
    tb_update): 7 for diffusion (two FMA+add line sums, their sum, times dt*D, plus c); per velocity component
    that is not dropped a difference and a product; one add when both are present; times -dt and the final add."""
    nx_term = not (dropped and vx == 0.0)
    ny_term = not (dropped and vy == 0.0)
    n = 7 + 2 * nx_term + 2 * ny_term
    if nx_term and ny_term:
        n += 1
    if nx_term or ny_term:
        n += 2
    return n


def computed_over_useful(csim, nx, ny, T, nbr):
    """Cell updates the sweep computes per cell update it stores, from the library's own host-side plan
    (csim_sweep_plan): a work item of h rows runs ticks over h + 2T rows (rounded up to two ticks = 4 rows) of
    all 128 columns of its strip, for every one of the T levels; 120 columns and h rows of level T are stored."""
    items = csim.sweep_plan(nx, ny, T, tuple(int(v) for v in nbr))
    rows = sum(4 * ((y1 - y0 + 2 * T + 3) // 4) for (_, _, _, y0, y1) in items)
    return rows * 128.0 / (float(nx) * float(ny))


def fp64_pipe_report(rate_per_gpu, ops, factor, n_sm, sm_mhz, sm_max_mhz):
    """Share of the FP64 pipe's issue slots the sweep fills: lane-operations per second (cell updates/s x
    operations per update x re-computation factor) over SMs x 64 lanes x SM clock.  Against the clock sampled
    during the timed region (what the board ran at under its power cap) and against the maximum clock."""
    lane_ops = rate_per_gpu * ops * factor
    out = {"ops_per_cell_update": ops, "computed_over_useful_cells": factor, "lane_ops_per_s": lane_ops,
           "lanes_per_sm_per_clk": FP64_LANES_PER_SM_PER_CLK, "sms": n_sm, "sm_mhz_sampled": sm_mhz,
           "frac_at_sampled_clock": None, "frac_at_max_clock": None,
           "note": "share of the FP64 pipe's issue slots (SMs x 64 lanes x clock) filled by the sweep's FP64 "
                   "instructions, re-computed halo cells included: the kernel's other ceiling beside HBM"}
    if sm_mhz:
        out["frac_at_sampled_clock"] = lane_ops / (n_sm * FP64_LANES_PER_SM_PER_CLK * sm_mhz * 1e6)
    if sm_max_mhz:
        out["frac_at_max_clock"] = lane_ops / (n_sm * FP64_LANES_PER_SM_PER_CLK * sm_max_mhz * 1e6)
    return out


def workload_name(tile, dims):
    return (f"{tile}x{tile} per GPU, Gaussian hotspot, diffusion+advection, periodic BCs, "
            f"decomp {{{dims[0]},{dims[1]}}} (global {tile * dims[0]}x{tile * dims[1]})")


def base_config(args, dims, inner):
    """The `config` keys both arms print (the reference arm differs only in timesteps_per_step)."""
    return {"workload": workload_name(args.tile, dims), "timesteps_per_step": inner,
            "parallelism": f"cartesian {dims[0]}x{dims[1]}, 1 rank per GPU", "physics": dict(PHYS),
            "tile": args.tile, "bc": args.bc}


# -------------------------------------------------------------------------------------------------
def cpu_reference_rate(tile, timesteps, threads):
    """The reference's own compute objects (oracle/_ref) on `threads` emulated ranks: returns
    (cell-updates/s, loop seconds, kind).  Falls back to the C port if _ref is absent."""
    from oracle import cpu_oracle as co
    if not co.available("port"):
        co.build()
    kind = "reference" if co.available("ref") else "port"
    orc = co.Oracle("ref" if kind == "reference" else "port")
    p = co.SimParams(nx=tile, ny=tile, steps=timesteps, out_every=10 ** 9, bc=(2, 2, 2, 2), **PHYS)
    if kind == "reference":
        r = orc.run(p, nranks=threads, want_frames=False, want_final=False)
        secs = r["seconds"]
    else:
        threads = 1
        t0 = time.perf_counter()
        orc.run(p, nranks=1, want_frames=False, want_final=False)
        secs = time.perf_counter() - t0
    return tile * tile * timesteps / secs, secs, kind, threads


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # rank emulation wants a count the decomposition likes; cap at 64 threads
    threads = min(threads, 64)
    inner = args.ref_inner
    tile = args.tile
    rates, secs = [], []
    for i in range(args.warmup + args.steps):
        rate, s, kind, used = cpu_reference_rate(tile, inner, threads)
        if i >= args.warmup:
            rates.append(rate)
            secs.append(s)
    total_cells = tile * tile * inner * len(secs)
    value = total_cells / sum(secs)
    # the same objects as the reference's README builds them (no CMAKE_BUILD_TYPE → no optimisation flags):
    # one short sample, reported beside the -O2 figure, not used for the headline
    flagless = None
    try:
        from oracle import cpu_oracle as co
        if co.available("ref_O0"):
            p0 = co.SimParams(nx=tile, ny=tile, steps=1, out_every=10 ** 9, bc=(2, 2, 2, 2), **PHYS)
            r0 = co.Oracle("ref_O0").run(p0, nranks=threads, want_frames=False, want_final=False)
            flagless = {"value": tile * tile / r0["seconds"], "unit": "cell-updates/s",
                        "sample": f"1 time step of the {tile}x{tile} tile, flagless build (-O0) of the reference objects"}
    except Exception:  # noqa: BLE001
        flagless = None
    sample = (f"bounded sample: ONE {tile}x{tile} tile (the per-GPU tile of the workload, whatever --gpus says: the "
              f"CPU arm does not grow with N) split over {used} emulated ranks (threads), {inner} time steps per "
              f"bench step, loop time only (main.cpp:89-123 timing region without NetCDF writes); reference compute "
              f"objects, -O2, no MPI launcher (MPI not installed)")
    cfg = base_config(args, dims_for(args.gpus), inner)
    if args.gpus > 1:  # at N = 1 the two arms run the very same grid
        cfg["workload"] += "; reference arm: ONE tile on the host cores"
    # the keys our arm adds to `config`, with this arm's values (same key set in both lines)
    field_mib = ((tile + 2) * (tile + 2) * 8) >> 20
    cfg.update({
        "halo_exchange": f"in-process copies between {used} emulated ranks (threads) following src/halo.cpp:28-43",
        "arithmetic": "full, 15 FP64 ops per cell in the reference's order, no FMA (g++ -O2 -ffp-contract=off, no -march)",
        "l2_policy": f"host arm: two {field_mib} MiB fields in host memory, no cache flush between steps",
        "e2e_workload": "host-resident fields: nothing to copy, e2e = value",
        "numa_node_of_pinned_buffers": None})
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": used, "kind": kind, "sample": sample,
                         "flagless_build": flagless},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
def parity_windows(dec, w):
    """Top-left corners (local y, x) of the checked windows of one rank's tile: the four tile corners (a
    rank seam crossing, a four-corner point of the decomposition or a corner of the physical domain, depending
    on where the rank sits), the four edge midpoints, the centre and two off-centre interior points."""
    nx, ny = dec.nx_local, dec.ny_local
    w = min(w, nx, ny)
    pts = [(0, 0), (0, nx - w), (ny - w, 0), (ny - w, nx - w),
           (0, (nx - w) // 2), (ny - w, (nx - w) // 2), ((ny - w) // 2, 0), ((ny - w) // 2, nx - w),
           ((ny - w) // 2, (nx - w) // 2), (min(1234, ny - w), min(4321, nx - w)), (min(ny - w, 3 * ny // 4), nx // 3)]
    return sorted(set(pts)), w


def check_parity(csim, co, port, got, dec, nxg, nyg, phys, bc_codes, steps, w=64, dx=1.0, dy=1.0):
    """Bit-compare windows of `got` (this rank's interior after `steps` time steps from the initial condition)
    with the CPU oracle advanced on the sub-domain around each window.  The sub-domain reaches `steps` cells
    beyond the window (the dependency cone of `steps` 5-point updates) or to the physical boundary; its cut
    sides carry the neighbouring cells' initial values as frozen ghosts, whose error cannot reach the window
    in `steps` steps.  Returns (windows checked, windows that differ)."""
    pts, w = parity_windows(dec, w)
    bad = 0
    for (y, x) in pts:
        gy, gx = dec.y_offset + y, dec.x_offset + x
        y0, y1 = max(gy - steps, 0), min(gy + w + steps, nyg)
        x0, x1 = max(gx - steps, 0), min(gx + w + steps, nxg)
        sub = np.zeros((y1 - y0 + 2, x1 - x0 + 2))  # padded; ghosts outside the physical domain stay 0
        iy0, iy1, ix0, ix1 = max(y0 - 1, 0), min(y1 + 1, nyg), max(x0 - 1, 0), min(x1 + 1, nxg)
        ic = csim.initial_condition_host(csim.Decomp2D.window(nxg, nyg, ix0, iy0, ix1 - ix0, iy1 - iy0), 0, dx, dy)
        sub[iy0 - (y0 - 1):iy1 - (y0 - 1), ix0 - (x0 - 1):ix1 - (x0 - 1)] = ic
        bc = (bc_codes[0] if x0 == 0 else 2, bc_codes[1] if x1 == nxg else 2,
              bc_codes[2] if y0 == 0 else 2, bc_codes[3] if y1 == nyg else 2)
        sp = co.SimParams(nx=x1 - x0, ny=y1 - y0, dx=dx, dy=dy, steps=steps, out_every=steps, bc=bc, **phys)
        want = port.run(sp, u0_padded=sub)["final"]
        a = np.ascontiguousarray(got[y:y + w, x:x + w])
        b = np.ascontiguousarray(want[gy - y0:gy - y0 + w, gx - x0:gx - x0 + w])
        if not np.array_equal(a.view(np.uint64), b.view(np.uint64)):
            bad += 1
    return len(pts), bad


def shared_file_check(rank, world, local_rank, dist):
    """N > 1 only: the C++ driver (host/build/climate_sim_b200) run as one process per GPU of this box —
    RANK/WORLD_SIZE + rendezvous file, no MPI launcher — on a small grid with mixed boundaries; every rank
    writes its window of the SAME CDF-5 file (src/io.cpp:402-424).  Rank 0 reads the file back and compares
    every frame with the single-rank CPU oracle, bit for bit."""
    import shutil
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "climate-sim-mpi-cpp_b200", "host", "build", "climate_sim_b200")
    if not os.path.exists(exe):
        return None
    box = [tempfile.mkdtemp(prefix="csim_shared_file_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    work = box[0]
    nx, ny, steps, every = 301 + 64 * world, 260 + 32 * world, 50, 10
    cli = [f"--nx={nx}", f"--ny={ny}", "--D=0.05", "--vx=-0.5", "--vy=0.25", f"--steps={steps}", f"--out_every={every}",
           "--bc.right=neumann", "--bc.bottom=periodic"]
    env = dict(os.environ)
    env.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local_rank), CSIM_JOB_ID=os.path.basename(work),
               CSIM_RENDEZVOUS=os.path.join(work, "rendezvous"))
    r = subprocess.run([exe] + cli, cwd=work, env=env, capture_output=True, text=True, timeout=600)
    rcs = [None] * world
    dist.all_gather_object(rcs, (r.returncode, (r.stdout + r.stderr)[-400:] if r.returncode else ""))
    out = None
    if rank == 0:
        out = {"ranks": world, "grid": f"{nx}x{ny}", "frames": steps // every, "bit_identical": False,
               "what": "climate_sim_b200 as one process per GPU, all ranks writing disjoint windows of one CDF-5 "
                       "file; frames compared with the single-rank CPU oracle"}
        if any(rc for rc, _ in rcs):
            out["error"] = "; ".join(f"rank {i}: rc {rc} {msg}" for i, (rc, msg) in enumerate(rcs) if rc)
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from cdf5_reader import read_cdf5
            from oracle import cpu_oracle as co
            if not co.available("port"):
                co.build()
            f = read_cdf5(os.path.join(work, "outputs", "snapshots.nc"))
            p = co.SimParams(nx=nx, ny=ny, D=0.05, vx=-0.5, vy=0.25, steps=steps, out_every=every, bc=(0, 1, 2, 0))
            want = co.Oracle("port").run(p)["frames"]
            out["bit_identical"] = bool(f["numrecs"] == steps // every and f["data"].shape == want.shape and
                                        np.array_equal(f["data"].view(np.uint64), want.view(np.uint64)))
        shutil.rmtree(work, ignore_errors=True)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    csim = importlib.import_module("climate-sim-mpi-cpp_b200")
    ctx = csim.Context(local_rank)
    # pinned buffers are first touched on the GPU's own NUMA node; the thread's CPU mask is put back right after
    # the allocations (the CPU baseline below must see every core it would see without this)
    cpu_mask = os.sched_getaffinity(0)
    numa_node = ctx.bind_numa()
    tile, inner = args.tile, args.inner
    dims = csim.Decomp2D.init(world, 0, 1, 1).dims
    nxg, nyg = tile * dims[0], tile * dims[1]
    if args.global_size:  # strong scaling (configs[3]): the global grid is fixed, tiles shrink with N
        nxg = nyg = args.global_size
    dec = csim.Decomp2D.init(world, rank, nxg, nyg)
    if world > 1:
        box = [csim.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(world, rank, box[0])

    P = csim.BCType.Periodic
    bcs = csim.BCConfig(P, P, P, P)
    if args.bc == "dn":  # configs[3]: Dirichlet and Neumann sides
        Dn, Nn = csim.BCType.Dirichlet, csim.BCType.Neumann
        bcs = csim.BCConfig(Dn, Nn, Dn, Nn)
    params = csim.make_step_params(PHYS["D"], PHYS["vx"], PHYS["vy"], PHYS["dt"], bcs, dec)
    host_in = ctx.pinned_empty((dec.ny_local + 2, dec.nx_local + 2))
    host_in[:] = 0.0
    csim.initial_condition_host(dec, 1, args.dx, args.dy, out=host_in)
    host_out = [ctx.pinned_empty((dec.ny_local, dec.nx_local)) for _ in range(2)]
    for h in host_out:
        h[:1, :] = 0.0  # touch
    os.sched_setaffinity(0, cpu_mask)
    u = csim.Field(ctx, dec.nx_local, dec.ny_local, 1, args.dx, args.dy)
    tmp = csim.Field(ctx, dec.nx_local, dec.ny_local, 1, args.dx, args.dy)
    u.upload(host_in)
    halo_path = "none"
    if world > 1:
        halo_path = "decided on the first window (csim_halo_path)"

    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local_rank))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum_int(v):
        if world == 1:
            return int(v)
        t = torch.tensor([int(v)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count - l0
        barrier()
        per_rank.clear()
        if world > 1:
            t = torch.zeros(world, dtype=torch.float64, device="cuda")
            t[rank] = ms
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            per_rank.extend(float(v) for v in t.tolist())
        return reduce_max(ms), reduce_sum_int(launches)

    enqueue_s = []
    per_rank = []  # device time of the last timed() region on every rank

    def window_resident():
        t0 = time.perf_counter()
        csim.run_steps(u, tmp, params, dec, inner)
        enqueue_s.append(time.perf_counter() - t0)  # host time to enqueue one window (no sync inside)

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = timed(window_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_per_rank = [v / args.steps for v in per_rank]
    # dev.yaml has vy = 0, so on the (clean, monotone) benchmark field the library drops the y-advection
    # term: 11 instead of 14 FP64 operations per cell (csim_field_value_state).  Time the same window
    # with both velocity components non-zero and negative vx as well, so the full arithmetic and the
    # forward-difference branches are on record next to the headline, with the same number of windows.
    dropped = u.value_state == 1 and (PHYS["vx"] == 0.0 or PHYS["vy"] == 0.0) and csim.steps_per_sweep() >= 3
    gen_params = csim.make_step_params(*[PHYS_ALL_TERMS[k] for k in ("D", "vx", "vy", "dt")], bcs, dec)

    def window_general():
        csim.run_steps(u, tmp, gen_params, dec, inner)

    gen_steps = args.steps
    sampler_g = ClockSampler(local_rank)
    sampler_g.start()
    ms_gen, _ = timed(window_general, gen_steps, 2)
    clocks_gen = sampler_g.stop()

    # ---- halo timeline (N > 1): one window run eagerly with timestamps around every exchange -------------
    halo = None
    if world > 1:
        halo_path = {"peer": "T-line bands stored straight into the neighbours' ghost lines over NVLink (tiles mapped "
                             "with CUDA IPC) by single-warp CTAs that co-reside with the interior sweep, one flag per "
                             "neighbour; hidden behind the interior sweep; NCCL only for bootstrap",
                     "nccl": "T-line bands packed by a kernel, one grouped ncclSend/ncclRecv per block over NVLink, unpack "
                             "kernel; hidden behind the interior sweep"}.get(csim.halo_path(ctx), csim.halo_path(ctx))
        if os.environ.get("CSIM_GRAPH") == "1":
            halo_path += "; block loop replayed as a CUDA graph"
        barrier()
        csim.halo_profile(ctx, True)
        csim.run_steps(u, tmp, params, dec, inner)
        st = csim.halo_stats(ctx)
        barrier()
        wire = st["bytes_per_exchange"]
        us_max = reduce_max(st["exchange_us"])
        halo = {
            "bytes_sent_per_exchange_per_gpu": reduce_max(float(wire)),
            "formula": "8 B x sum over the up to 8 neighbours of (T lines x band length | T x T corner), T = "
                       f"{csim.steps_per_sweep()} (SURVEY.md 8d: 8*(edges_x*ny + edges_y*(nx+2))*T plus corners)",
            "exchanges_per_window": st["blocks"],
            "us_per_exchange": us_max,
            "us_first_exchange": reduce_max(st["first_exchange_us"]),
            "achieved_gbs_per_direction": (wire / (us_max * 1e-6) / 1e9) if us_max > 0 else None,
            "nvlink_peak_gbs_per_direction": NVLINK_GBS_PER_DIRECTION,
            "frac_of_nvlink": (wire / (us_max * 1e-6) / 1e9 / NVLINK_GBS_PER_DIRECTION) if us_max > 0 else None,
            "overlap_fraction": -reduce_max(-st["overlap_fraction"]),  # the worst rank
            "us_frame_sweep": reduce_max(st["frame_us"]), "us_interior_sweep": reduce_max(st["interior_us"]),
            "us_exchange_done_to_frame_start": reduce_max(st["wait_for_interior_us"]),
            "us_own_store_kernel": reduce_max(st["push_us"]) if csim.halo_path(ctx) == "peer" else None,
            "path": csim.halo_path(ctx),
            "note": "one exchange = everything between the end of the frame sweep and the start of the next one on the "
                    "exchange stream (peer: store kernel + flag wait; nccl: pack + grouped send/recv + unpack), CUDA "
                    "events, one eager window; max over ranks; overlap_fraction = share of the exchange time that "
                    "lies inside the concurrently running interior sweep (latency-bound: a 393 KB band is 0.4 us of "
                    "wire time at 900 GB/s)",
        }

    # ---- end to end: one simulation with host buffers ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 10))  # dev.yaml's shape: 10 output windows per run

        def simulation(nwin):
            pending = [None, None]
            u.upload_async(host_in)  # H2D of the padded initial tile from pinned memory (main.cpp:71)
            for k in range(nwin):
                csim.run_steps(u, tmp, params, dec, inner)
                if pending[k % 2] is not None:
                    ctx.event_wait(pending[k % 2])  # the host buffer is free again (its frame was consumed)
                pending[k % 2] = u.snapshot_async(host_out[k % 2])  # D2H of the de-haloed tile (io.cpp:411-418)
            for ev in pending:
                if ev is not None:
                    ctx.event_wait(ev)
            ctx.sync()

        simulation(2)  # warm-up: staging buffers, graph capture
        barrier()
        t0 = time.perf_counter()
        simulation(e2e_steps)
        secs = time.perf_counter() - t0
        barrier()
        secs = reduce_max(secs)
        # strict per-window variant: every window pays its own H2D and D2H, one after the other
        def window_copies():
            u.upload_async(host_in)
            csim.run_steps(u, tmp, params, dec, inner)
            u.download_interior_async(host_out[0])
            ctx.sync()

        ms_serial, _ = timed(window_copies, 2, 1)

        # what the box gives a bare device→host copy of one frame when every rank copies at once: the ceiling
        # of the per-window time above once the sweeps (shorter) are hidden behind the copy
        def bare_copy():
            ctx.event_wait(u.snapshot_async(host_out[0]))

        bare_copy()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            bare_copy()
        copy_s = reduce_max((time.perf_counter() - t0) / 3)
        barrier()
        cells_w = float(nxg) * float(nyg) * inner
        e2e = {"value": cells_w * e2e_steps / secs, "unit": "cell-updates/s",
               "h2d_bytes_per_step": int(host_in.nbytes // e2e_steps), "d2h_bytes_per_step": int(host_out[0].nbytes),
               "steps": e2e_steps, "ms_per_step": 1e3 * secs / e2e_steps,
               "mode": "ONE simulation per rank through the C ABI with pinned host buffers: H2D of the initial tile "
                       f"({host_in.nbytes} B, once, inside the timed region, amortised over the windows in "
                       "h2d_bytes_per_step), then per window 100 time steps + D2H of the de-haloed tile on the copy "
                       "stream, overlapping the next window's steps (what src/main.cpp:93-99 does at output steps); "
                       "host wall clock around a full drain, max over ranks",
               "d2h_copy_alone": {"ms_per_frame": 1e3 * copy_s, "gbs_per_rank": host_out[0].nbytes / copy_s / 1e9,
                                  "gbs_all_ranks": world * host_out[0].nbytes / copy_s / 1e9,
                                  "note": "pack + PCIe copy of one de-haloed frame with no time steps queued, all ranks "
                                          "at once, max over ranks: what a window costs when the copy is the bound"},
               "per_window_copies": {"value": cells_w * 2 / (ms_serial * 1e-3), "ms_per_step": ms_serial / 2,
                                     "h2d_bytes_per_step": int(host_in.nbytes), "d2h_bytes_per_step": int(host_out[0].nbytes),
                                     "mode": "every window: H2D of the padded tile, 100 steps, D2H, sync — no overlap"}}

    cells_per_window = float(nxg) * float(nyg) * inner
    value = cells_per_window * args.steps / (ms * 1e-3)
    gen_value = cells_per_window * gen_steps / (ms_gen * 1e-3)

    # sanity: the field must still be finite and must have moved (the work was really done)
    max_abs, bad = u.health()
    if bad or not (0.0 < max_abs <= 1.0):
        raise SystemExit(f"bench.py: field unhealthy after timing (max|u|={max_abs}, nonfinite={bad})")

    # ---- parity, outside every timed region ---------------------------------------------------------------
    parity = None
    if not args.no_parity:
        from oracle import cpu_oracle as co  # the checker, never the thing measured
        if not co.available("port"):
            co.build()
        port = co.Oracle("port")
        T = csim.steps_per_sweep()
        K = max(args.parity_steps, 2 * T + 1)  # at least two blocks and a remainder sweep
        n_win = n_bad = 0
        sets = []
        for name, phys in (("headline", PHYS), ("all_terms", PHYS_ALL_TERMS)):
            u.upload(host_in)
            pp = csim.make_step_params(phys["D"], phys["vx"], phys["vy"], phys["dt"], bcs, dec)
            csim.run_steps(u, tmp, pp, dec, K)
            u.download_interior_async(host_out[0])
            ctx.sync()
            w, b = check_parity(csim, co, port, host_out[0], dec, nxg, nyg, phys, bcs.as_tuple(), K, dx=args.dx,
                                dy=args.dy)
            n_win += w
            n_bad += b
            sets.append(name)
        n_win, n_bad = reduce_sum_int(n_win), reduce_sum_int(n_bad)
        parity = {"windows": n_win, "steps": K, "bit_identical": n_bad == 0, "windows_differing": n_bad,
                  "window": "64x64 cells", "per_rank": n_win // world, "parameter_sets": sets,
                  "against": "CPU oracle (oracle/oracle_port.c, pinned to the reference's objects by tests/test_oracle.py) "
                             "advanced on the sub-domain around each window: tile corners (rank seams / decomposition "
                             "corners / physical corners), edge midpoints, interior"}
        if n_bad:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "parity": parity, "error": "GPU field differs from the oracle"}),
                      flush=True)
            raise SystemExit(3)

    shared_file = None
    if world > 1 and not args.no_parity:
        barrier()
        shared_file = shared_file_check(rank, world, local_rank, dist)
        if rank == 0 and shared_file is not None and not shared_file["bit_identical"]:
            print(json.dumps({"metric": METRIC, "shared_file": shared_file,
                              "error": "the ranks' shared snapshot file differs from the oracle"}), flush=True)
        bad_file = [shared_file is not None and not shared_file["bit_identical"]] if rank == 0 else [None]
        dist.broadcast_object_list(bad_file, src=0)
        if bad_file[0]:
            raise SystemExit(4)

    # ---- roofline of the dominant kernel: the fused sweep k_step_tb, which advances T steps per launch ----
    # With all-periodic boundaries no boundary kernels run, so on one GPU the timed region is exactly
    # the sweeps; with N>1 the pack/NCCL/unpack and frame launches share the region (sweep count is
    # computed, not taken from the launch counter).
    peak, peak_src = measured_peak_gbs()
    T = csim.steps_per_sweep()
    sweeps_per_window = inner // T + (1 if inner % T else 0)
    if world == 1 and launches % args.steps == 0 and launches > 0:
        sweeps_per_window = launches // args.steps  # one GPU, frozen ghosts: every launch of the window is a sweep
        T = max(1, round(inner / sweeps_per_window))  # (the IEEE-division mode sweeps one step at a time)
    sweeps = args.steps * sweeps_per_window
    launch_ms = ms / sweeps
    steps_per_launch = inner / sweeps_per_window
    bytes_per_launch = float(dec.nx_local) * float(dec.ny_local) * ALG_BYTES_PER_CELL * steps_per_launch
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        try:
            ent = json.load(open(tpath)).get(f"{dec.nx_local}x{dec.ny_local}_T{T}", {})
            traffic = ent.get("bytes_per_launch")
            traffic_src = ent.get("source")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": f"csim::{csim.sweep_kernel()}<T={T}> (fused diffusion+advection sweep, {T} steps per launch"
                          + (", level-0 rows staged through shared memory by TMA bulk copies" if csim.sweep_kernel() == "k_step_tbs" else "")
                          + (", y-advection term dropped: vy == +0.0)" if dropped else ")"),
                "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "avg_launch_ms": launch_ms,
                "steps_per_launch": steps_per_launch, "frac_of_nominal_8TBs": achieved / 8000.0,
                "dram_frac_of_peak": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "note": "algorithmic bytes = 16 B per cell update; temporal blocking moves 16/T B per update "
                        "through HBM, so frac > 1 is expected; `traffic` = dram__bytes_read+write per launch of this "
                        "kernel on this tile from the committed ncu capture named in traffic_source (a one-GPU figure: "
                        "counters cannot be read outside a profiler); dram_frac_of_peak = traffic / live launch time / peak"}

    # the kernel's second ceiling: FP64 issue slots (unit spacing only; the division modes issue far more)
    fp64_general = None
    try:
        if (args.dx, args.dy) == (1.0, 1.0) and csim.steps_per_sweep() >= 2:
            n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
            factor = computed_over_useful(csim, dec.nx_local, dec.ny_local, csim.steps_per_sweep(), params.nbr)
            roofline["fp64_pipe"] = fp64_pipe_report(
                value / world, fp64_ops_per_cell(PHYS["vx"], PHYS["vy"], dropped), factor, n_sm,
                clocks.get("sm_mhz"), clocks.get("sm_max_mhz"))
            fp64_general = fp64_pipe_report(
                gen_value / world, fp64_ops_per_cell(PHYS_ALL_TERMS["vx"], PHYS_ALL_TERMS["vy"], False), factor, n_sm,
                clocks_gen.get("sm_mhz"), clocks_gen.get("sm_max_mhz"))
    except Exception as exc:  # noqa: BLE001 - a report, never a reason to lose the line
        roofline["fp64_pipe"] = {"error": repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = min(os.cpu_count() or 1, 64)
        ref_steps = max(4, int(args.ref_inner * 20 * (8192.0 / tile) ** 2))  # ≈ 10-20 s of CPU work
        rate, secs, kind, used = cpu_reference_rate(tile, ref_steps, threads)
        cpu = {"value": rate, "unit": "cell-updates/s", "cores": used, "kind": kind,
               "sample": f"{tile}x{tile}, {ref_steps} time steps of the same workload on {used} emulated ranks "
                         f"(threads), {secs:.1f} s loop time; reference compute objects -O2, no MPI launcher"}

    if rank == 0:
        cfg = base_config(args, dims, inner)
        if args.global_size:
            cfg["workload"] = (f"{nxg}x{nyg} global (strong scaling), Gaussian hotspot, diffusion+advection, BCs left/bottom "
                               f"Dirichlet, right/top Neumann, decomp {{{dims[0]},{dims[1]}}}, tile {dec.nx_local}x{dec.ny_local}"
                               if args.bc == "dn" else
                               f"{nxg}x{nyg} global (strong scaling), periodic BCs, decomp {{{dims[0]},{dims[1]}}}")
        if args.dx != 1.0 or args.dy != 1.0:
            cfg["workload"] += f"; dx={args.dx}, dy={args.dy}"
            cfg["spacing"] = {"dx": args.dx, "dy": args.dy,
                              "mode": "IEEE division (the spacing is not a power of two)" if not dropped and
                              csim.steps_per_sweep() and (args.dx, args.dy) != (1.0, 1.0) else "reciprocal"}
        cfg.update({
            "halo_exchange": halo_path,
            "arithmetic": ("vy == +0.0 on a scanned-clean field: y-advection term dropped, 11 FP64 ops per "
                           "cell, bit-identical (DESIGN.md 4.1)") if dropped else "full, 14 FP64 ops per cell",
            "l2_policy": f"inputs larger than L2 (two {host_in.nbytes >> 20} MiB fields per GPU vs 126 MB L2); no flush needed"
            if tile >= 4096 else "WARNING: fields fit in L2",
            "e2e_workload": "one simulation per rank: upload once, a de-haloed frame to the host every 100 steps",
            "numa_node_of_pinned_buffers": numa_node})
        line = {
            "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.global_size else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "halo": halo, "shared_file": shared_file,
            "all_terms": {"value": gen_value, "unit": "cell-updates/s", "vx": PHYS_ALL_TERMS["vx"],
                          "vy": PHYS_ALL_TERMS["vy"], "steps": gen_steps, "ms_per_step": ms_gen / gen_steps,
                          "clocks": clocks_gen, "fp64_pipe": fp64_general,
                          "note": "same windows with both velocity components non-zero (14 FP64 ops per cell): the "
                                  "rate of a run whose velocity has no exact zero component"},
            "gpu_launches": launches, "clocks": clocks,
            "ms_per_step_per_rank": ms_per_rank or None,
            "host_enqueue_ms_per_step": 1e3 * float(np.median(enqueue_s[-args.steps:])) if enqueue_s else None,
        }
        print(json.dumps(line), flush=True)
    u.close()
    tmp.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--tile", type=int, default=16384, help="per-GPU tile edge (16384 = configs[2], 8192 = configs[1])")
    ap.add_argument("--inner", type=int, default=100, help="time steps per bench step (output window)")
    ap.add_argument("--ref-inner", type=int, default=1, help="time steps per bench step for the CPU reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end simulation (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the bit-parity windows (profiling runs)")
    ap.add_argument("--parity-steps", type=int, default=10, help="time steps of the parity run (>= 2T+1)")
    ap.add_argument("--global-size", type=int, default=0,
                    help="strong scaling: fixed NxN global grid split over the ranks (32768 = configs[3])")
    ap.add_argument("--dx", type=float, default=1.0, help="grid spacing in x (a spacing that is not a power of two "
                    "makes the kernels divide: IEEE-division mode, src/diffusion.cpp:12-13)")
    ap.add_argument("--dy", type=float, default=1.0)
    ap.add_argument("--bc", choices=["periodic", "dn"], default="periodic",
                    help="dn: left/bottom Dirichlet, right/top Neumann (configs[3])")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"bench.py: note: warmup {args.warmup} < 3", file=sys.stderr)
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
