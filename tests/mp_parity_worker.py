"""One rank of a torchrun-launched parity run (used by tests/test_gpu_parity.py): every rank advances
its tile through csim_run_steps twice (with CSIM_GRAPH=1 the second pass replays the CUDA graph the first
one captured; CSIM_HALO=nccl selects the NCCL halo path instead of peer stores) and rank 0 gathers the tiles
and compares the global field with the single-rank CPU oracle, bit for bit."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    csim = importlib.import_module("climate-sim-mpi-cpp_b200")
    ctx = csim.Context(local)
    box = [csim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(world, rank, box[0])
    ok = True
    # the last case is large enough for ranks to drift apart: each of its passes re-uploads the initial tile
    # right before the call, which a neighbour's first halo store of that call must not overtake
    cases = [(1000, 700, (7, 3, 1, 5), (0.05, 0.5, -0.3, 0.1), (2, 2, 2, 2)),
             (777, 1301, (10, 4), (0.05, -0.4, -0.2, 0.1), (1, 0, 1, 2)),
             (6144, 4096, (9,), (0.05, 0.5, 0.0, 0.1), (2, 2, 2, 2))]
    for (nxg, nyg, steps, phys, bc) in cases:
        for path in ("captured", "replayed"):
            dec = csim.Decomp2D.init(world, rank, nxg, nyg)
            u = csim.Field(ctx, dec.nx_local, dec.ny_local, 1, 1.0, 1.0)
            tmp = csim.Field(ctx, dec.nx_local, dec.ny_local, 1, 1.0, 1.0)
            ic = csim.initial_condition_host(dec, 1, 1.0, 1.0)
            p = csim.make_step_params(*phys, csim.BCConfig(*[csim.BCType(b) for b in bc]), dec)
            for rep in range(3 if nxg > 4000 else 1):
                u.upload(ic)
                for k in steps:
                    csim.run_steps(u, tmp, p, dec, k)
            tile = u.download_interior()
            ctx.sync()
            tiles = [None] * world
            dist.all_gather_object(tiles, (dec.x_offset, dec.y_offset, tile))
            if rank == 0:
                from oracle import cpu_oracle as co
                glob = np.zeros((nyg, nxg))
                for (xo, yo, t) in tiles:
                    glob[yo:yo + t.shape[0], xo:xo + t.shape[1]] = t
                sp = co.SimParams(nx=nxg, ny=nyg, D=phys[0], vx=phys[1], vy=phys[2], dt=phys[3],
                                  steps=sum(steps), out_every=sum(steps), bc=bc)
                want = co.Oracle("port").run(sp)["final"]
                same = np.array_equal(glob.view(np.uint64), want.view(np.uint64))
                print(f"mp_parity {path} halo={csim.halo_path(ctx)} {nxg}x{nyg} world={world}: "
                      f"{'OK' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
            dist.barrier()
            u.close()
            tmp.close()
            dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MP_PARITY_PASS" if ok else "MP_PARITY_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
