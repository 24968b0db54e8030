#!/usr/bin/env python
"""Timeline of the pipelined end-to-end windows (bench.py's e2e): per lane and window, when the H2D,
the sweeps and the D2H start and end on the device.  Run on the GPU box: python tools/e2e_timeline.py"""
import importlib
import os
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
csim = importlib.import_module("climate-sim-mpi-cpp_b200")
n, inner, n_lanes, per_lane = 8192, 100, int(os.environ.get("LANES", 3)), 3
dec = csim.Decomp2D.init(1, 0, n, n)
P = csim.BCType.Periodic
params = csim.make_step_params(0.05, 0.5, 0.0, 0.1, csim.BCConfig(P, P, P, P), dec)
lanes = []
for _ in range(n_lanes):
    c = csim.Context(0)
    hin = c.pinned_empty((n + 2, n + 2))
    hin[:] = 0.0
    csim.initial_condition_host(dec, 1, 1.0, 1.0, out=hin)
    lanes.append(dict(c=c, u=csim.Field(c, n, n, 1, 1.0, 1.0), t=csim.Field(c, n, n, 1, 1.0, 1.0), hin=hin,
                      hout=c.pinned_empty((n, n)), s=torch.cuda.ExternalStream(c.stream_ptr), ev=[]))
origin = torch.cuda.Event(enable_timing=True)


def loop(L, count):
    for _ in range(count):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(L["s"])
        L["u"].upload_async(L["hin"])
        e[1].record(L["s"])
        csim.run_steps(L["u"], L["t"], params, dec, inner)
        e[2].record(L["s"])
        L["u"].download_interior_async(L["hout"])
        e[3].record(L["s"])
        L["c"].sync()
        L["ev"].append(e)


for L in lanes:
    loop(L, 1)
    L["ev"].clear()
torch.cuda.synchronize()
origin.record(torch.cuda.current_stream())
torch.cuda.synchronize()
ths = [threading.Thread(target=loop, args=(L, per_lane)) for L in lanes]
for t in ths:
    t.start()
for t in ths:
    t.join()
torch.cuda.synchronize()
for i, L in enumerate(lanes):
    for w, e in enumerate(L["ev"]):
        ts = [origin.elapsed_time(x) for x in e]
        print(f"lane {i} window {w}: H2D {ts[0]:7.2f}-{ts[1]:7.2f}  steps -{ts[2]:7.2f}  D2H -{ts[3]:7.2f}   "
              f"(h2d {ts[1]-ts[0]:.2f}, steps {ts[2]-ts[1]:.2f}, d2h {ts[3]-ts[2]:.2f})")
