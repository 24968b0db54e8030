"""Run every single-rank golden case through the fused path and compare bit for bit.  Executed as a
subprocess by test_gpu_parity.py with CSIM_TB_MAXT / CSIM_TB_CHUNK / CSIM_TB_EDGE_SPLIT set, because the
library reads those knobs once per process."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    csim = importlib.import_module("climate-sim-mpi-cpp_b200")
    z = np.load(os.path.join(ROOT, "tests", "golden", "timestep_golden.npz"))
    ctx = csim.Context(0)
    bad = []
    for name in [str(n) for n in z["names"]]:
        meta, phys = z[name + "/meta"], z[name + "/phys"]
        nx, ny, steps = int(meta[0]), int(meta[1]), int(meta[2])
        bc = [int(b) for b in meta[4:8]]
        dx, dy, D, vx, vy, dt = (float(v) for v in phys[:6])
        dec = csim.Decomp2D.single(nx, ny)
        if name + "/u0" in z.files:
            u0 = np.ascontiguousarray(z[name + "/u0"])
        else:
            u0 = csim.initial_condition_host(dec, 1, dx, dy, "gaussian_hotspot", *[float(v) for v in phys[6:10]])
        u, tmp = csim.Field(ctx, nx, ny, 1, dx, dy), csim.Field(ctx, nx, ny, 1, dx, dy)
        u.upload(u0)
        dt = min(dt, csim.safe_dt(dx, dy, vx, vy, D))
        p = csim.make_step_params(D, vx, vy, dt, csim.BCConfig(*[csim.BCType(b) for b in bc]), dec)
        csim.run_steps(u, tmp, p, dec, steps)
        got = u.download_interior()
        if not np.array_equal(got.view(np.uint64), z[name + "/final"].view(np.uint64)):
            bad.append(name)
    print(f"T={csim.steps_per_sweep()} cases={len(z['names'])} bad={bad}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
