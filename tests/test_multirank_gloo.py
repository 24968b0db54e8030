"""Host-side logic of the N>1 path on CPU: world_size 2, 4 and 8 over torch.distributed/gloo.

What runs here is the product's own decomposition (csim_decomp_init) and wide-exchange plan
(csim_wide_exchange_plan — the same table halo.cu feeds to its pack/NCCL/unpack sequence), with gloo
send/recv standing in for NCCL and numpy slicing for the pack/unpack kernels.  After one exchange
every rank's extended tile must equal the matching window of the global field: bands, corners,
remainder tiles, and the frozen ghost line of perpendicular physical sides included."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nxg, nyg, T, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        csim = importlib.import_module("climate-sim-mpi-cpp_b200")
        dec = csim.Decomp2D.init(world, rank, nxg, nyg)
        # the bootstrap bench.py uses: rank 0 makes an id, everyone receives it
        box = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == bytes(range(128))

        rng = np.random.default_rng(42)  # same global field on every rank
        G = rng.standard_normal((nyg + 2, nxg + 2))  # padded global field, ghost ring included
        nx, ny = dec.nx_local, dec.ny_local
        E = np.full((ny + 2 * T, nx + 2 * T), np.nan)  # extended tile, interior origin at (T, T)

        def win(x0, y0, w, h):  # view of E in interior coordinates
            return E[y0 + T:y0 + T + h, x0 + T:x0 + T + w]

        def gwin(x0, y0, w, h):  # same window of the global field
            gx, gy = dec.x_offset + x0 + 1, dec.y_offset + y0 + 1
            return G[gy:gy + h, gx:gx + w]

        # own cells: interior plus the ghost line of every physical side
        pl, pr, pb, pt = (n == csim.PROC_NULL for n in dec.nbr)
        ox0, ox1 = (-1 if pl else 0), (nx + 1 if pr else nx)
        oy0, oy1 = (-1 if pb else 0), (ny + 1 if pt else ny)
        win(ox0, oy0, ox1 - ox0, oy1 - oy0)[:] = gwin(ox0, oy0, ox1 - ox0, oy1 - oy0)

        snd, rcv = csim.wide_exchange_plan(dec, T)
        reqs, landing = [], []
        for s, r in zip(snd, rcv):
            if s.peer < 0:
                continue
            out = torch.from_numpy(np.ascontiguousarray(win(s.x0, s.y0, s.w, s.h)))
            buf = torch.empty((r.h, r.w), dtype=torch.float64)
            reqs.append(dist.isend(out, s.peer))
            reqs.append(dist.irecv(buf, r.peer))
            landing.append((r, buf))
        for rq in reqs:
            rq.wait()
        for r, buf in landing:
            win(r.x0, r.y0, r.w, r.h)[:] = buf.numpy()

        # every cell a T-step sweep may read must now hold the global value
        vx0, vx1 = (-1 if pl else -T), (nx + 1 if pr else nx + T)
        vy0, vy1 = (-1 if pb else -T), (ny + 1 if pt else ny + T)
        got = win(vx0, vy0, vx1 - vx0, vy1 - vy0)
        want = gwin(vx0, vy0, vx1 - vx0, vy1 - vy0)
        ok = bool(np.array_equal(got, want))
        # gather of tiles into the global array (what bench/tests do with download_interior)
        tiles = [None] * world
        dist.all_gather_object(tiles, (dec.x_offset, dec.y_offset, win(0, 0, nx, ny).copy()))
        glob = np.zeros((nyg, nxg))
        for (xo, yo, t) in tiles:
            glob[yo:yo + t.shape[0], xo:xo + t.shape[1]] = t
        ok = ok and bool(np.array_equal(glob, G[1:-1, 1:-1]))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, ""))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()))


# T = 4 is the blocking depth the library runs by default; {4,2} is the 8-GPU decomposition of the scaling runs
@pytest.mark.parametrize("world,nxg,nyg,T", [(2, 37, 20, 3), (2, 64, 48, 1), (4, 45, 38, 3), (4, 41, 33, 2),
                                             (2, 50, 31, 4), (4, 53, 47, 4), (8, 83, 41, 4)])
def test_wide_exchange_plan_over_gloo(world, nxg, nyg, T):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + world * 7 + T) % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, nxg, nyg, T, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(30) for p in procs]
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg or 'extended tile differs from the global field'}"
