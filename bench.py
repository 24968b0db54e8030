#!/usr/bin/env python
"""bench.py — cell-updates/s of the fused diffusion+advection timestep on N B200s.

    python bench.py --gpus 1 --steps K --warmup W              # our CUDA path (default)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W   # one rank per GPU
    python bench.py --impl reference ...                          # the reference's CPU code

Workload (BASELINE.json configs[1]; per-GPU tile fixed as N grows → weak scaling): an 8192x8192
tile per GPU, Gaussian hotspot, dx=dy=1, D=0.05, vx=0.5, vy=0, dt=0.1, all-periodic boundaries
(= frozen zero ghosts), 2-D Cartesian decomposition {1,1},{2,1},{2,2},{4,2}.
`--tile 16384` gives configs[2].

A bench "step" is one output window of `--inner` (default 100, dev.yaml's out_every) time steps:
  value : cells * inner * K / device time, fields resident in HBM, timed with CUDA events on the
          library's stream, max over ranks.
  e2e   : same window through the C ABI with HOST buffers: H2D of the padded tile from pinned
          memory, `inner` steps, D2H of the de-haloed tile (what the reference hands to its NetCDF
          writer at an output step), every step, inside the timed region.
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

PHYS = dict(D=0.05, vx=0.5, vy=0.0, dt=0.1)  # configs/dev.yaml physics
ALG_BYTES_PER_CELL = 16.0                     # one 8-byte read + one 8-byte write per cell update
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


def workload_name(tile, dims):
    return (f"{tile}x{tile} per GPU, Gaussian hotspot, diffusion+advection, periodic BCs, "
            f"decomp {{{dims[0]},{dims[1]}}} (global {tile * dims[0]}x{tile * dims[1]})")


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
    sample = (f"bounded sample: ONE {tile}x{tile} tile (the per-GPU tile of the workload) split over {used} "
              f"emulated ranks (threads), {inner} time steps per bench step, loop time only (main.cpp:89-123 timing region without NetCDF writes); "
              f"reference compute objects, -O2, no MPI launcher (MPI not installed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(tile, dims_for(args.gpus)), "timesteps_per_step": inner},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": used, "kind": kind, "sample": sample,
                         "flagless_build": flagless},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
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
    csim.initial_condition_host(dec, 1, 1.0, 1.0, out=host_in)
    host_out = ctx.pinned_empty((dec.ny_local, dec.nx_local))
    u = csim.Field(ctx, dec.nx_local, dec.ny_local, 1, 1.0, 1.0)
    tmp = csim.Field(ctx, dec.nx_local, dec.ny_local, 1, 1.0, 1.0)
    u.upload(host_in)
    halo_path = "none"
    if world > 1:
        # default: T-line bands packed by a kernel, one grouped ncclSend/ncclRecv per block over NVLink,
        # hidden behind the interior sweep.  CSIM_HALO=p2p selects the peer-memory push instead, which
        # measured slower with the round-1b kernel (profiles/r01b_weak_scaling.md).
        halo_path = "pack + grouped NCCL send/recv over NVLink + unpack, overlapped with the interior sweep"
        if os.environ.get("CSIM_HALO", "nccl") == "p2p":
            csim.peer_setup(u, tmp, dec)  # neighbours' tiles mapped over CUDA IPC: direct NVLink stores
            halo_path = "peer-memory push (CUDA IPC over NVLink) + flag, NCCL only for bootstrap"

    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local_rank))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

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
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
        return ms, launches

    enqueue_s = []

    def window_resident():
        t0 = time.perf_counter()
        csim.run_steps(u, tmp, params, dec, inner)
        enqueue_s.append(time.perf_counter() - t0)  # host time to enqueue one window (no sync inside)

    def window_e2e():
        u.upload_async(host_in)
        csim.run_steps(u, tmp, params, dec, inner)
        u.download_interior_async(host_out)
        ctx.sync()  # the caller needs the frame on the host before the next window

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = timed(window_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    # dev.yaml has vy = 0, so on the (clean, monotone) benchmark field the library drops the y-advection
    # term: 11 instead of 14 FP64 operations per cell (csim_field_value_state).  Time the same window
    # with both velocity components non-zero and negative vx as well, so the full arithmetic and the
    # forward-difference branches are on record next to the headline.
    dropped = u.value_state == 1 and (PHYS["vx"] == 0.0 or PHYS["vy"] == 0.0) and csim.steps_per_sweep() >= 3
    gen_params = csim.make_step_params(PHYS["D"], -0.5, 0.25, PHYS["dt"], bcs, dec)

    def window_general():
        csim.run_steps(u, tmp, gen_params, dec, inner)

    gen_steps = max(2, min(args.steps, 5))
    ms_gen, _ = timed(window_general, gen_steps, 1)
    e2e_steps = max(2, min(args.steps, 3))
    ms_e2e = None
    if not args.no_e2e:
        ms_e2e, _ = timed(window_e2e, e2e_steps, 1)

    # The same end-to-end window with THREE windows in flight (one GPU only): extra contexts with their
    # own stream, tiles and pinned buffers, so that one window's PCIe copies (H2D before, D2H after its
    # 100 steps) overlap the other window's sweeps.  Every window still pays its own H2D and D2H inside
    # the timed region; only the overlap is new.  This is the throughput a caller with independent
    # members to advance (an ensemble) gets from the same C-ABI calls.
    ms_pipe, pipe_steps, n_lanes = None, 0, 3
    if world == 1 and not args.no_e2e:
        lanes = [(ctx, u, tmp, host_in, host_out)]
        for _ in range(n_lanes - 1):
            c2 = csim.Context(local_rank)
            hin2 = c2.pinned_empty(host_in.shape)
            hin2[:] = host_in
            lanes.append((c2, csim.Field(c2, dec.nx_local, dec.ny_local, 1, 1.0, 1.0),
                          csim.Field(c2, dec.nx_local, dec.ny_local, 1, 1.0, 1.0), hin2,
                          c2.pinned_empty(host_out.shape)))
        per_lane = max(2, min(args.steps, 4))
        pipe_steps = n_lanes * per_lane
        errors = []

        def lane_loop(lane, count):
            # one host thread per lane: a lane blocks in its own context (value scan after the upload,
            # final sync) without holding up the others; ctypes releases the GIL during the calls
            c, uu, tt, hin, hout = lane
            try:
                for _ in range(count):
                    uu.upload_async(hin)
                    csim.run_steps(uu, tt, params, dec, inner)
                    uu.download_interior_async(hout)
                    c.sync()  # the frame is on the host
            except Exception as exc:  # noqa: BLE001
                errors.append(exc)

        def run_lanes(count):
            ths = [threading.Thread(target=lane_loop, args=(lane, count)) for lane in lanes]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            if errors:
                raise errors[0]

        run_lanes(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_lanes(per_lane)
        torch.cuda.synchronize()
        ms_pipe = 1e3 * (time.perf_counter() - t0)  # several streams: wall clock around a full drain
        for lane in lanes[1:]:
            if not np.array_equal(host_out, lane[4]):
                raise SystemExit("bench.py: the pipelined lanes disagree")
            lane[1].close()
            lane[2].close()
            lane[0].close()

    cells_per_window = float(nxg) * float(nyg) * inner
    value = cells_per_window * args.steps / (ms * 1e-3)
    e2e_value = cells_per_window * e2e_steps / (ms_e2e * 1e-3) if ms_e2e else None
    gen_value = cells_per_window * gen_steps / (ms_gen * 1e-3)

    # sanity: the field must still be finite and must have moved (the work was really done)
    max_abs, bad = u.health()
    if bad or not (0.0 < max_abs <= 1.0):
        raise SystemExit(f"bench.py: field unhealthy after timing (max|u|={max_abs}, nonfinite={bad})")

    # roofline of the dominant kernel: the fused sweep k_step_tb, which advances T steps per launch.
    # With all-periodic boundaries no boundary kernels run, so on one GPU the timed region is exactly
    # the sweeps; with N>1 the pack/NCCL/unpack and frame launches share the region (sweep count is
    # computed, not taken from the launch counter).
    peak, peak_src = measured_peak_gbs()
    T = csim.steps_per_sweep()
    sweeps_per_window = inner // T + (1 if inner % T else 0)
    sweeps = args.steps * sweeps_per_window
    launch_ms = ms / sweeps
    steps_per_launch = inner / sweeps_per_window
    bytes_per_launch = float(dec.nx_local) * float(dec.ny_local) * ALG_BYTES_PER_CELL * steps_per_launch
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{tile}_T{T}", {}).get("bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": f"csim::k_step_tb<T={T}> (fused diffusion+advection sweep, {T} steps per launch"
                                              + (", y-advection term dropped: vy == +0.0)" if dropped else ")"),
                "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "avg_launch_ms": launch_ms,
                "steps_per_launch": steps_per_launch, "frac_of_nominal_8TBs": achieved / 8000.0,
                "note": "algorithmic bytes = 16 B per cell update; temporal blocking moves 16/T B per update "
                        "through HBM, so frac > 1 is expected; measured DRAM traffic per launch is in `traffic`; at "
                        "T = 3 the kernel runs the FP64 pipe at ~74 % and HBM at ~79 % of the measured peak "
                        "(DESIGN.md 4.1)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = min(os.cpu_count() or 1, 64)
        ref_steps = args.ref_inner * 20  # ≈ 10 s of CPU work at 8192² on 16 threads
        rate, secs, kind, used = cpu_reference_rate(tile, ref_steps, threads)
        cpu = {"value": rate, "unit": "cell-updates/s", "cores": used, "kind": kind,
               "sample": f"{tile}x{tile}, {ref_steps} time steps of the same workload on {used} emulated ranks "
                         f"(threads), {secs:.1f} s loop time; reference compute objects -O2, no MPI launcher"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.global_size else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(tile, dims) if not args.global_size else
                       (f"{nxg}x{nyg} global (strong scaling), Gaussian hotspot, diffusion+advection, BCs left/bottom "
                        f"Dirichlet, right/top Neumann, decomp {{{dims[0]},{dims[1]}}}, tile {dec.nx_local}x{dec.ny_local}"
                        if args.bc == "dn" else f"{nxg}x{nyg} global (strong scaling), periodic BCs, decomp {{{dims[0]},{dims[1]}}}"),
                       "timesteps_per_step": inner,
                       "parallelism": f"cartesian {dims[0]}x{dims[1]}, 1 rank per GPU", "halo_exchange": halo_path,
                       "physics": dict(PHYS),
                       "arithmetic": ("vy == +0.0 on a scanned-clean field: y-advection term dropped, 11 FP64 ops per "
                                      "cell, bit-identical (DESIGN.md 4.1)") if dropped else "full, 14 FP64 ops per cell",
                       "l2_policy": "inputs larger than L2 (two 537 MB fields per GPU vs 126 MB L2); no flush needed"
                       if tile >= 4096 else "WARNING: fields fit in L2"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": ({"value": cells_per_window * pipe_steps / (ms_pipe * 1e-3), "unit": "cell-updates/s",
                     "h2d_bytes_per_step": int(host_in.nbytes), "d2h_bytes_per_step": int(host_out.nbytes),
                     "steps": pipe_steps, "ms_per_step": ms_pipe / pipe_steps,
                     "mode": f"{n_lanes} windows in flight (one context each on the GPU): every window's H2D and D2H are "
                             "inside the timed region and overlap the other windows' sweeps; one host thread per window "
                             "lane; host wall clock around a full drain",
                     "serial": {"value": e2e_value, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                                "mode": "one window at a time: H2D, 100 steps, D2H, sync"}}
                    if ms_pipe else None if ms_e2e is None else
                    {"value": e2e_value, "unit": "cell-updates/s",
                     "h2d_bytes_per_step": int(host_in.nbytes), "d2h_bytes_per_step": int(host_out.nbytes),
                     "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                     "mode": "one window at a time per rank: H2D, 100 steps, D2H, sync"}),
            "all_terms": {"value": gen_value, "unit": "cell-updates/s", "vx": -0.5, "vy": 0.25, "steps": gen_steps,
                          "note": "same window with both velocity components non-zero (14 FP64 ops per cell)"},
            "gpu_launches": launches, "clocks": clocks,
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
    ap.add_argument("--tile", type=int, default=8192, help="per-GPU tile edge (8192 = configs[1], 16384 = configs[2])")
    ap.add_argument("--inner", type=int, default=100, help="time steps per bench step (output window)")
    ap.add_argument("--ref-inner", type=int, default=4, help="time steps per bench step for the CPU reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end windows (profiling runs)")
    ap.add_argument("--global-size", type=int, default=0,
                    help="strong scaling: fixed NxN global grid split over the ranks (32768 = configs[3])")
    ap.add_argument("--bc", choices=["periodic", "dn"], default="periodic",
                    help="dn: left/bottom Dirichlet, right/top Neumann (configs[3])")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"bench.py: note: warmup {args.warmup} < 3", file=sys.stderr)
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
