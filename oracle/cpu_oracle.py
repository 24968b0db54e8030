"""ctypes front-end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package never does.

* ``Oracle("port")``  — ``liboracle_port.so``, the plain-C restatement (``oracle_port.c``).
* ``Oracle("ref")``   — ``_ref/libcsim_ref.so``, the reference's own objects behind
  ``ref_harness.cpp`` (built by ``oracle/Makefile`` where ``/root/reference`` exists; the prebuilt
  file travels to the GPU box).
* ``Oracle("ref_O0")`` — same at the flagless ``-O0`` build the reference README produces.

Both expose the same calls with the same struct layouts.  Arrays are C-contiguous float64 numpy
arrays in the reference layout (``include/field.hpp:5-21``): shape ``(ny+2h, nx+2h)``, x fastest.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

DIRICHLET, NEUMANN, PERIODIC = 0, 1, 2  # include/boundary.hpp:5
PROC_NULL = -1
BC_NAMES = {"dirichlet": 0, "fixed": 0, "neumann": 1, "noflux": 1, "zero-flux": 1, "periodic": 2,
            "period": 2}  # src/io.cpp:35-44


class _Config(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("D", C.c_double), ("vx", C.c_double), ("vy", C.c_double), ("dt", C.c_double),
                ("steps", C.c_int), ("out_every", C.c_int), ("bc", C.c_int * 4),
                ("ic_preset", C.c_int), ("A", C.c_double), ("sigma_frac", C.c_double),
                ("xc_frac", C.c_double), ("yc_frac", C.c_double)]


class _Decomp(C.Structure):
    _fields_ = [("dims", C.c_int * 2), ("coords", C.c_int * 2), ("nbr_lr", C.c_int * 2),
                ("nbr_du", C.c_int * 2), ("nx_global", C.c_int), ("ny_global", C.c_int),
                ("nx_local", C.c_int), ("ny_local", C.c_int), ("x_offset", C.c_int),
                ("y_offset", C.c_int)]


@dataclass
class SimParams:
    """The SimConfig members the hot path reads, defaults as include/io.hpp:10-39."""
    nx: int = 256
    ny: int = 256
    dx: float = 1.0
    dy: float = 1.0
    D: float = 0.0
    vx: float = 0.0
    vy: float = 0.0
    dt: float = 0.1
    steps: int = 100
    out_every: int = 50
    bc: tuple = (DIRICHLET, DIRICHLET, DIRICHLET, DIRICHLET)  # left, right, bottom, top
    ic_preset: int = 0  # 0 gaussian_hotspot, 1 constant_zero
    A: float = 1.0
    sigma_frac: float = 0.05
    xc_frac: float = 0.5
    yc_frac: float = 0.5

    def c(self) -> _Config:
        return _Config(self.nx, self.ny, self.dx, self.dy, self.D, self.vx, self.vy, self.dt,
                       self.steps, self.out_every, (C.c_int * 4)(*self.bc), self.ic_preset, self.A,
                       self.sigma_frac, self.xc_frac, self.yc_frac)


DEV_YAML = dict(nx=512, ny=512, dx=1.0, dy=1.0, D=0.05, vx=0.5, vy=0.0, dt=0.1, steps=1000,
                out_every=100, bc=(DIRICHLET, NEUMANN, PERIODIC, DIRICHLET))  # configs/dev.yaml


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _chk(a, shape=None):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need C-contiguous float64"
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def build(ref_root: str = "/root/reference") -> None:
    """Run oracle/Makefile (idempotent).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", HERE, f"REF={ref_root}"], check=True)


def available(kind: str) -> bool:
    return os.path.exists(_path(kind))


def _path(kind: str) -> str:
    return {"port": os.path.join(HERE, "liboracle_port.so"),
            "ref": os.path.join(HERE, "_ref", "libcsim_ref.so"),
            "ref_O0": os.path.join(HERE, "_ref", "libcsim_ref_O0.so")}[kind]


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        self.pfx = "orc_" if kind == "port" else "ref_"
        self.lib = C.CDLL(_path(kind))
        L = self.lib
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        f = self._f
        f("diffusion_step").argtypes = [dp, dp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4
        f("advection_step").argtypes = [dp, dp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 5
        f("apply_boundary").argtypes = [dp, C.c_int, C.c_int, C.c_int, ip, ip, C.c_double]
        f("safe_dt").argtypes = [C.c_double] * 5
        f("safe_dt").restype = C.c_double
        if kind == "port":
            L.orc_run.argtypes = [C.POINTER(_Config), C.c_int, C.c_int, dp, dp, C.c_int, dp, dp]
            L.orc_decomp_init.argtypes = [C.POINTER(_Decomp)] + [C.c_int] * 4
            L.orc_minmax.argtypes = [dp, C.c_size_t, dp, dp]
        else:
            L.ref_run.argtypes = [C.POINTER(_Config), C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, dp]
            L.ref_decomp_all.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(_Decomp)]
            L.ref_exchange_all.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(dp)]
            L.ref_field_at_throws.argtypes = [C.c_int] * 5
            L.ref_field_index.argtypes = [C.c_int] * 5
            L.ref_field_index.restype = C.c_long
            L.ref_last_error.restype = C.c_char_p

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    # -- single functions ---------------------------------------------------------------------
    def diffusion_step(self, u, out, h, dx, dy, D, dt):
        ny, nx = u.shape[0] - 2 * h, u.shape[1] - 2 * h
        _chk(u), _chk(out, u.shape)
        self._f("diffusion_step")(_dp(u), _dp(out), nx, ny, h, dx, dy, D, dt)
        return out

    def advection_step(self, u, out, h, dx, dy, vx, vy, dt):
        ny, nx = u.shape[0] - 2 * h, u.shape[1] - 2 * h
        _chk(u), _chk(out, u.shape)
        self._f("advection_step")(_dp(u), _dp(out), nx, ny, h, dx, dy, vx, vy, dt)
        return out

    def apply_boundary(self, f, h, nbr, bc, value):
        ny, nx = f.shape[0] - 2 * h, f.shape[1] - 2 * h
        _chk(f)
        self._f("apply_boundary")(_dp(f), nx, ny, h, (C.c_int * 4)(*nbr), (C.c_int * 4)(*bc), value)
        return f

    def safe_dt(self, dx, dy, vx, vy, D):
        return self._f("safe_dt")(dx, dy, vx, vy, D)

    def decomp(self, size, nxg, nyg):
        """List of per-rank dicts, src/decomp.cpp:5-34."""
        arr = (_Decomp * size)()
        if self.kind == "port":
            for r in range(size):
                self.lib.orc_decomp_init(C.byref(arr[r]), size, r, nxg, nyg)
        else:
            self.lib.ref_decomp_all(size, nxg, nyg, arr)
        return [dict(dims=tuple(d.dims), coords=tuple(d.coords), nbr_lr=tuple(d.nbr_lr),
                     nbr_du=tuple(d.nbr_du), nx_local=d.nx_local, ny_local=d.ny_local,
                     x_offset=d.x_offset, y_offset=d.y_offset) for d in arr]

    def exchange(self, size, nxg, nyg, h, tiles):
        """exchange_halos on all ranks (reference objects only)."""
        assert self.kind != "port"
        ptrs = (C.POINTER(C.c_double) * size)(*[_dp(_chk(t)) for t in tiles])
        self.lib.ref_exchange_all(size, nxg, nyg, h, ptrs)
        return tiles

    # -- the time loop ------------------------------------------------------------------------
    def run(self, p: SimParams, nranks=1, clamp_dt=True, u0_padded=None, want_frames=True,
            want_final=True, want_padded=False):
        """Returns dict(frames=(k,ny,nx) or None, final=(ny,nx) or None, padded=..., seconds=...).

        Frames follow src/main.cpp:93-99: one at the START of every step n with n % out_every == 0;
        ``final`` is the state after the last step (never written by the reference)."""
        nfr = (p.steps + p.out_every - 1) // p.out_every if want_frames else 0
        frames = np.zeros((max(nfr, 1), p.ny, p.nx)) if want_frames else None
        final = np.zeros((p.ny, p.nx)) if want_final else None
        padded = np.zeros((p.ny + 2, p.nx + 2)) if want_padded else None
        if u0_padded is not None:
            _chk(u0_padded, (p.ny + 2, p.nx + 2))
        cfg = p.c()
        secs = C.c_double(0.0)
        if self.kind == "port":
            rc = self.lib.orc_run(C.byref(cfg), nranks, int(clamp_dt), _dp(u0_padded), _dp(frames),
                                  nfr, _dp(final), _dp(padded))
        else:
            rc = self.lib.ref_run(C.byref(cfg), nranks, int(clamp_dt), _dp(u0_padded), _dp(frames),
                                  nfr, _dp(final), _dp(padded), C.byref(secs))
        if rc < 0:
            raise RuntimeError(f"oracle run failed rc={rc}")
        return dict(frames=frames[:nfr] if want_frames else None, final=final, padded=padded,
                    seconds=secs.value)

    def minmax(self, a):
        assert self.kind == "port"
        mn, mx = C.c_double(), C.c_double()
        a = np.ascontiguousarray(a)
        self.lib.orc_minmax(_dp(a), a.size, C.byref(mn), C.byref(mx))
        return mn.value, mx.value
