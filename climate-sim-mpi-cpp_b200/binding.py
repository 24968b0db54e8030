"""ctypes binding of libcsim_b200.so (include/csim.h) with the reference's own names on top.

This is the Python mirror of the reference's C++ interface for the timestep path — ``Field``,
``BCType``/``BCConfig``, ``Decomp2D``, ``diffusion_step``, ``advection_step``, ``apply_boundary``,
``exchange_halos``, ``safe_dt`` (reference ``include/*.hpp``) — used by ``tests/`` and ``bench.py``.
The C++ drop-in headers with the same names live in ``host/``.

There is no fallback: if the shared library is missing or no CUDA device is usable, every entry
point raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field as _dc_field
from enum import IntEnum

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSIM_LIB_PATH: A/B timing of two builds of the same library (tools/); never a different implementation
LIB_PATH = os.environ.get("CSIM_LIB_PATH") or os.path.join(_HERE, "libcsim_b200.so")

PROC_NULL = -1  # MPI_PROC_NULL
UNIQUE_ID_BYTES = 128


class CsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"csim error {code}: {msg}")
        self.code = code


class BCType(IntEnum):  # include/boundary.hpp:5
    Dirichlet = 0
    Neumann = 1
    Periodic = 2


_BC_ALIASES = {"dirichlet": BCType.Dirichlet, "fixed": BCType.Dirichlet, "neumann": BCType.Neumann,
               "noflux": BCType.Neumann, "zero-flux": BCType.Neumann, "periodic": BCType.Periodic,
               "period": BCType.Periodic}


def bc_from_string(s: str) -> BCType:
    """src/io.cpp:35-44 (case-insensitive, aliases fixed/noflux/zero-flux/period)."""
    try:
        return _BC_ALIASES[s.lower()]
    except KeyError:
        raise RuntimeError("Unknown BC type: " + s) from None


def bc_to_string(bc: BCType) -> str:
    return BCType(bc).name.lower()  # src/io.cpp:46-56


@dataclass
class BCConfig:  # include/boundary.hpp:7-12
    left: BCType = BCType.Dirichlet
    right: BCType = BCType.Dirichlet
    bottom: BCType = BCType.Dirichlet
    top: BCType = BCType.Dirichlet

    def as_tuple(self):
        return (int(self.left), int(self.right), int(self.bottom), int(self.top))


class _FieldInfo(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("halo", C.c_int), ("dx", C.c_double),
                ("dy", C.c_double), ("pitch", C.c_int64), ("lead_x", C.c_int), ("lead_y", C.c_int),
                ("rows", C.c_int64), ("base", C.c_void_p), ("interior", C.c_void_p)]


class HaloStats(C.Structure):
    """csim_halo_stats"""
    _fields_ = [("blocks", C.c_int), ("bytes_per_exchange", C.c_size_t), ("first_exchange_us", C.c_double),
                ("exchange_us", C.c_double), ("overlap_fraction", C.c_double), ("frame_us", C.c_double),
                ("interior_us", C.c_double), ("total_ms", C.c_double), ("push_us", C.c_double),
                ("wait_for_interior_us", C.c_double)]


class StepParams(C.Structure):
    """csim_step_params"""
    _fields_ = [("D", C.c_double), ("vx", C.c_double), ("vy", C.c_double), ("dt", C.c_double),
                ("bc", C.c_int * 4), ("nbr", C.c_int * 4), ("bc_value", C.c_double),
                ("flags", C.c_int)]


STEP_FAST_RECIP = 0x1
STEP_NO_TEMPORAL = 0x2


class SweepItem(C.Structure):
    _fields_ = [("strip", C.c_int), ("x0", C.c_int), ("x1", C.c_int), ("y0", C.c_int), ("y1", C.c_int)]


class XRegion(C.Structure):
    """csim_xregion"""
    _fields_ = [("x0", C.c_int), ("y0", C.c_int), ("w", C.c_int), ("h", C.c_int), ("peer", C.c_int)]


class _Decomp(C.Structure):
    _fields_ = [("dims", C.c_int * 2), ("coords", C.c_int * 2), ("nbr", C.c_int * 4),
                ("nx_global", C.c_int), ("ny_global", C.c_int), ("nx_local", C.c_int),
                ("ny_local", C.c_int), ("x_offset", C.c_int), ("y_offset", C.c_int)]


_lib = None


def lib():
    """Load the C-ABI library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (there is no CPU or PyTorch fallback for this path)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        sig = {
            "csim_ctx_create": [C.c_int, C.POINTER(vp)],
            "csim_ctx_destroy": [vp],
            "csim_sync": [vp],
            "csim_field_create": [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(vp)],
            "csim_field_destroy": [vp],
            "csim_field_get_info": [vp, C.POINTER(_FieldInfo)],
            "csim_field_fill": [vp, C.c_double],
            "csim_field_upload": [vp, vp],
            "csim_field_download": [vp, vp],
            "csim_field_download_interior": [vp, vp],
            "csim_field_upload_async": [vp, vp],
            "csim_field_download_interior_async": [vp, vp],
            "csim_field_download_interior_be_async": [vp, vp],
            "csim_event_record": [vp, C.POINTER(vp)],
            "csim_event_wait": [vp, vp],
            "csim_field_get": [vp, C.c_int, C.c_int, dp],
            "csim_field_set": [vp, C.c_int, C.c_int, C.c_double],
            "csim_field_swap": [vp, vp],
            "csim_field_copy": [vp, vp],
            "csim_host_alloc": [C.c_size_t, C.POINTER(vp)],
            "csim_host_free": [vp],
            "csim_diffusion_step": [vp, vp, C.c_double, C.c_double],
            "csim_advection_step": [vp, vp, C.c_double, C.c_double, C.c_double],
            "csim_apply_boundary": [vp, ip, ip, C.c_double],
            "csim_step_fused": [vp, vp, C.POINTER(StepParams), C.c_int],
            "csim_minmax": [vp, dp, dp],
            "csim_field_health": [vp, dp, C.POINTER(C.c_uint64)],
            "csim_field_value_state": [vp],
            "csim_decomp_init": [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Decomp)],
            "csim_comm_unique_id": [C.c_char_p],
            "csim_comm_init": [vp, C.c_int, C.c_int, C.c_char_p],
            "csim_comm_destroy": [vp],
            "csim_comm_allreduce_max": [vp, dp, C.c_int],
            "csim_halo_exchange": [vp, C.POINTER(_Decomp)],
            "csim_wide_exchange_plan": [C.POINTER(_Decomp), C.c_int, C.POINTER(XRegion), C.POINTER(XRegion)],
            "csim_run_steps": [vp, vp, C.POINTER(StepParams), C.POINTER(_Decomp), C.c_int],
            "csim_sweep_plan": [C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, C.POINTER(SweepItem), C.c_int, ip],
            "csim_initial_condition_host": [vp, C.POINTER(_Decomp), C.c_int, C.c_int, C.c_int,
                                            C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                            C.c_double, C.c_double],
            "csim_initial_condition_device": [vp, C.POINTER(_Decomp), C.c_int, C.c_int, C.c_int, C.c_double,
                                              C.c_double, C.c_double, C.c_double],
            "csim_device_count": [ip],
            "csim_field_download_window": [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp],
            "csim_halo_profile": [vp, C.c_int],
            "csim_field_snapshot_async": [vp, vp, C.c_int, C.POINTER(C.c_void_p)],
            "csim_bind_thread_to_device_numa": [C.c_int, ip],
            "csim_halo_stats_get": [vp, C.POINTER(HaloStats)],
        }
        for name, args in sig.items():
            fn = getattr(L, name, None)
            if fn is None:
                if os.environ.get("CSIM_LIB_PATH"):  # an older build under A/B timing lacks newer entry points
                    continue
                raise ImportError(f"{LIB_PATH} does not export {name}: rebuild it (__graft_entry__.build())")
            fn.argtypes = args
            fn.restype = C.c_int
        L.csim_last_error.restype = C.c_char_p
        L.csim_ctx_stream.argtypes = [vp]
        L.csim_ctx_stream.restype = vp
        L.csim_ctx_device.argtypes = [vp]
        L.csim_ctx_launch_count.argtypes = [vp]
        L.csim_ctx_launch_count.restype = C.c_uint64
        L.csim_safe_dt.argtypes = [C.c_double] * 5
        L.csim_safe_dt.restype = C.c_double
        L.csim_abi_version.restype = C.c_int
        L.csim_steps_per_sweep.restype = C.c_int
        if hasattr(L, "csim_div_by_const"):
            L.csim_div_by_const.argtypes = [C.c_double, C.c_double]
            L.csim_div_by_const.restype = C.c_double
            L.csim_div_by_const_fast.argtypes = [C.c_double, C.c_double]
            L.csim_div_by_const_fast.restype = C.c_int
        if hasattr(L, "csim_sweep_kernel"):
            L.csim_sweep_kernel.restype = C.c_char_p
        if hasattr(L, "csim_halo_path"):
            L.csim_halo_path.argtypes = [vp]
            L.csim_halo_path.restype = C.c_char_p
        if hasattr(L, "csim_exp_variant"):
            L.csim_exp_variant.restype = C.c_int
            L.csim_exp_restated.argtypes = [C.c_double, C.c_int]
            L.csim_exp_restated.restype = C.c_double
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise_for(rc)


def raise_for(rc):
    msg = lib().csim_last_error().decode(errors="replace")
    if rc == 4:  # CSIM_ERR_RANGE → the reference throws std::out_of_range (src/field.cpp:16)
        raise IndexError(msg)
    raise CsimError(rc, msg)


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need a C-contiguous float64 array"
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU, one stream (csim_ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(lib().csim_ctx_create(device, C.byref(self._h)))
        self.device = device
        self._pinned = []

    def sync(self):
        _check(lib().csim_sync(self._h))

    @property
    def stream_ptr(self) -> int:
        return lib().csim_ctx_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return int(lib().csim_ctx_launch_count(self._h))

    def pinned_empty(self, shape) -> np.ndarray:
        """Pinned host array (cudaMallocHost) for upload/download at full PCIe speed."""
        n = int(np.prod(shape))
        p = C.c_void_p()
        _check(lib().csim_host_alloc(n * 8, C.byref(p)))
        buf = (C.c_double * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float64).reshape(shape)
        self._pinned.append(p)
        return arr

    def bind_numa(self) -> int:
        """Pin the calling thread to the CPUs of this GPU's NUMA node (best effort); returns the node or -1."""
        node = C.c_int(-1)
        _check(lib().csim_bind_thread_to_device_numa(self.device, C.byref(node)))
        return node.value

    def event_wait(self, event):
        _check(lib().csim_event_wait(self._h, event))

    def comm_init(self, size: int, rank: int, unique_id: bytes):
        assert len(unique_id) == UNIQUE_ID_BYTES
        _check(lib().csim_comm_init(self._h, size, rank, unique_id))

    def allreduce_max(self, values):
        """MPI_Reduce(MAX) over the ranks of this context's communicator (all ranks get the result)."""
        arr = (C.c_double * len(values))(*values)
        _check(lib().csim_comm_allreduce_max(self._h, arr, len(values)))
        return list(arr)

    def close(self):
        if self._h:
            for p in self._pinned:
                lib().csim_host_free(p)
            self._pinned = []
            lib().csim_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(lib().csim_comm_unique_id(buf))
    return buf.raw


class Field:
    """Field(nx, ny, h, dx, dy) — include/field.hpp:5-21 — held in device memory."""

    def __init__(self, ctx: Context, nx: int, ny: int, h: int, dx: float, dy: float):
        self.ctx = ctx
        self.nx_local, self.ny_local, self.halo, self.dx, self.dy = nx, ny, h, dx, dy
        self._h = C.c_void_p()
        _check(lib().csim_field_create(ctx._h, nx, ny, h, dx, dy, C.byref(self._h)))

    def nx_total(self):
        return self.nx_local + 2 * self.halo

    def ny_total(self):
        return self.ny_local + 2 * self.halo

    @property
    def info(self) -> _FieldInfo:
        i = _FieldInfo()
        _check(lib().csim_field_get_info(self._h, C.byref(i)))
        return i

    def fill(self, value: float):
        _check(lib().csim_field_fill(self._h, value))

    def at(self, i: int, j: int) -> float:
        v = C.c_double()
        _check(lib().csim_field_get(self._h, i, j, C.byref(v)))
        return v.value

    def set(self, i: int, j: int, value: float):
        _check(lib().csim_field_set(self._h, i, j, value))

    def upload(self, host: np.ndarray):
        assert host.shape == (self.ny_total(), self.nx_total()), host.shape
        _check(lib().csim_field_upload(self._h, _ptr(host)))

    def upload_async(self, host: np.ndarray):
        assert host.shape == (self.ny_total(), self.nx_total()), host.shape
        _check(lib().csim_field_upload_async(self._h, _ptr(host)))

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        """Field::data as a (ny+2h, nx+2h) array."""
        if out is None:
            out = np.empty((self.ny_total(), self.nx_total()))
        _check(lib().csim_field_download(self._h, _ptr(out)))
        return out

    def download_interior(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.ny_local, self.nx_local))
        _check(lib().csim_field_download_interior(self._h, _ptr(out)))
        return out

    def download_interior_be_async(self, out: np.ndarray):
        """De-haloed tile as big-endian doubles (NetCDF wire order) into a pinned buffer, asynchronously."""
        assert out.size == self.nx_local * self.ny_local and out.flags["C_CONTIGUOUS"]
        _check(lib().csim_field_download_interior_be_async(self._h, out.ctypes.data_as(C.c_void_p)))

    def download_window(self, x0: int, y0: int, w: int, h: int) -> np.ndarray:
        """A w x h window, first cell (x0, y0) in interior coordinates (csim_field_download_window)."""
        out = np.empty((h, w))
        _check(lib().csim_field_download_window(self._h, x0, y0, w, h, _ptr(out)))
        return out

    def download_interior_async(self, out: np.ndarray):
        _check(lib().csim_field_download_interior_async(self._h, _ptr(out)))

    def snapshot_async(self, out: np.ndarray, big_endian=False):
        """De-haloed tile → pinned `out` on the copy stream, overlapped with the time steps queued next;
        returns the event to pass to Context.event_wait (csim_field_snapshot_async)."""
        assert out.size == self.nx_local * self.ny_local and out.flags["C_CONTIGUOUS"]
        ev = C.c_void_p()
        _check(lib().csim_field_snapshot_async(self._h, out.ctypes.data_as(C.c_void_p), 1 if big_endian else 0,
                                               C.byref(ev)))
        return ev

    def swap(self, other: "Field"):
        _check(lib().csim_field_swap(self._h, other._h))

    def copy_to(self, dst: "Field"):
        _check(lib().csim_field_copy(self._h, dst._h))

    def minmax(self):
        mn, mx = C.c_double(), C.c_double()
        _check(lib().csim_minmax(self._h, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def health(self):
        m, n = C.c_double(), C.c_uint64()
        _check(lib().csim_field_health(self._h, C.byref(m), C.byref(n)))
        return m.value, int(n.value)

    @property
    def value_state(self):
        """0 unknown, 1 clean, 2 tainted (csim_field_value_state)."""
        return int(lib().csim_field_value_state(self._h))

    def close(self):
        if self._h:
            lib().csim_field_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class Decomp2D:
    """include/decomp.hpp:4-17; ``init(size, rank, nx_global, ny_global)`` replaces the MPI communicator."""
    dims: tuple = (0, 0)
    coords: tuple = (0, 0)
    nbr_lr: tuple = (PROC_NULL, PROC_NULL)
    nbr_du: tuple = (PROC_NULL, PROC_NULL)
    nx_global: int = 0
    ny_global: int = 0
    nx_local: int = 0
    ny_local: int = 0
    x_offset: int = 0
    y_offset: int = 0
    _c: _Decomp = _dc_field(default_factory=_Decomp, repr=False)

    @classmethod
    def init(cls, size: int, rank: int, nx_global: int, ny_global: int) -> "Decomp2D":
        d = _Decomp()
        _check(lib().csim_decomp_init(size, rank, nx_global, ny_global, C.byref(d)))
        return cls(tuple(d.dims), tuple(d.coords), (d.nbr[0], d.nbr[1]), (d.nbr[2], d.nbr[3]),
                   d.nx_global, d.ny_global, d.nx_local, d.ny_local, d.x_offset, d.y_offset, d)

    @classmethod
    def single(cls, nx: int, ny: int) -> "Decomp2D":
        return cls.init(1, 0, nx, ny)

    @classmethod
    def window(cls, nx_global: int, ny_global: int, x_offset: int, y_offset: int, nx_local: int,
               ny_local: int) -> "Decomp2D":
        """A tile at an arbitrary position of the global grid, without neighbours (host-side helper: the
        initial condition of a sub-window, e.g. for window-wise parity checks)."""
        d = _Decomp()
        d.dims[0] = d.dims[1] = 1
        for s in range(4):
            d.nbr[s] = PROC_NULL
        d.nx_global, d.ny_global, d.nx_local, d.ny_local = nx_global, ny_global, nx_local, ny_local
        d.x_offset, d.y_offset = x_offset, y_offset
        return cls((1, 1), (0, 0), (PROC_NULL, PROC_NULL), (PROC_NULL, PROC_NULL), nx_global, ny_global, nx_local,
                   ny_local, x_offset, y_offset, d)

    @property
    def nbr(self):
        return (self.nbr_lr[0], self.nbr_lr[1], self.nbr_du[0], self.nbr_du[1])


def _nbr_of(dec) -> tuple:
    if dec is None:
        return (PROC_NULL,) * 4
    if isinstance(dec, Decomp2D):
        return dec.nbr
    return tuple(dec)


def diffusion_step(u: Field, out: Field, D: float, dt: float):
    """include/diffusion.hpp:4"""
    _check(lib().csim_diffusion_step(u._h, out._h, D, dt))


def advection_step(u: Field, out: Field, vx: float, vy: float, dt: float):
    """include/advection.hpp:4 (accumulates into out)"""
    _check(lib().csim_advection_step(u._h, out._h, vx, vy, dt))


def apply_boundary(f: Field, dec, bc: BCConfig, value: float = 0.0):
    """include/boundary.hpp:14"""
    nbr = (C.c_int * 4)(*_nbr_of(dec))
    bcs = (C.c_int * 4)(*bc.as_tuple())
    _check(lib().csim_apply_boundary(f._h, nbr, bcs, value))


def exchange_halos(f: Field, dec: Decomp2D):
    """include/halo.hpp:7 (the communicator is the one bound to the field's context)"""
    _check(lib().csim_halo_exchange(f._h, C.byref(dec._c)))


def sweep_plan(nx: int, ny: int, T: int, nbr=(-1, -1, -1, -1), resident_warps: int = 0, part: int = 0):
    """csim_sweep_plan (host only): list of (strip, x0, x1, y0, y1) work items of one fused sweep."""
    nb = (C.c_int * 4)(*nbr)
    n = C.c_int(0)
    _check(lib().csim_sweep_plan(nx, ny, T, nb, resident_warps, part, None, 0, C.byref(n)))
    items = (SweepItem * max(n.value, 1))()
    _check(lib().csim_sweep_plan(nx, ny, T, nb, resident_warps, part, items, n.value, C.byref(n)))
    return [(it.strip, it.x0, it.x1, it.y0, it.y1) for it in items[:n.value]]


def steps_per_sweep() -> int:
    """Temporal blocking depth T of the fused sweep."""
    return lib().csim_steps_per_sweep()


def wide_exchange_plan(dec: Decomp2D, T: int):
    """(send, recv) lists of 8 XRegion each: the geometry of the T-line exchange of csim_run_steps."""
    snd, rcv = (XRegion * 8)(), (XRegion * 8)()
    _check(lib().csim_wide_exchange_plan(C.byref(dec._c), T, snd, rcv))
    return list(snd), list(rcv)


def safe_dt(dx, dy, vx, vy, D) -> float:
    """include/stability.hpp:5-16"""
    return lib().csim_safe_dt(dx, dy, vx, vy, D)


def make_step_params(D, vx, vy, dt, bc: BCConfig, dec=None, bc_value=0.0, flags=0) -> StepParams:
    return StepParams(D, vx, vy, dt, (C.c_int * 4)(*bc.as_tuple()), (C.c_int * 4)(*_nbr_of(dec)),
                      bc_value, flags)


def step_fused(u: Field, tmp: Field, p: StepParams, nsteps: int = 1):
    """nsteps iterations of src/main.cpp:102-109 (no exchange); u holds the newest state after."""
    _check(lib().csim_step_fused(u._h, tmp._h, C.byref(p), nsteps))


def run_steps(u: Field, tmp: Field, p: StepParams, dec: Decomp2D | None, nsteps: int):
    """nsteps iterations of src/main.cpp:101-109 on this rank (exchange + fused step)."""
    _check(lib().csim_run_steps(u._h, tmp._h, C.byref(p), C.byref(dec._c) if dec is not None else None,
                                nsteps))


def initial_condition_host(dec: Decomp2D, halo, dx, dy, preset="gaussian_hotspot", A=1.0,
                           sigma_frac=0.05, xc_frac=0.5, yc_frac=0.5, out=None) -> np.ndarray:
    """src/init.cpp:12-47 on a host tile of the reference layout."""
    presets = {"gaussian_hotspot": 0, "constant_zero": 1}
    if preset not in presets:
        raise RuntimeError("Unknown IC preset: " + preset)  # init.cpp:42
    if out is None:
        out = np.zeros((dec.ny_local + 2 * halo, dec.nx_local + 2 * halo))
    _check(lib().csim_initial_condition_host(_ptr(out), C.byref(dec._c), halo, dec.nx_global,
                                             dec.ny_global, dx, dy, presets[preset], A, sigma_frac,
                                             xc_frac, yc_frac))
    return out


def initial_condition_device(f: "Field", dec: Decomp2D, preset="gaussian_hotspot", A=1.0, sigma_frac=0.05,
                             xc_frac=0.5, yc_frac=0.5):
    """src/init.cpp:12-47 generated on the device into the interior of `f` (bit-identical to
    initial_condition_host when exp_variant() is 0 or 1; raises CsimError(UNSUPPORTED) otherwise)."""
    presets = {"gaussian_hotspot": 0, "constant_zero": 1}
    if preset not in presets:
        raise RuntimeError("Unknown IC preset: " + preset)  # init.cpp:42
    _check(lib().csim_initial_condition_device(f._h, C.byref(dec._c), dec.nx_global, dec.ny_global,
                                               presets[preset], A, sigma_frac, xc_frac, yc_frac))


def div_by_const(a: float, d: float) -> float:
    """The kernels' constant-divisor division (step_math.cuh) on the host."""
    return float(lib().csim_div_by_const(a, d))


def div_by_const_fast(a: float, d: float) -> bool:
    return bool(lib().csim_div_by_const_fast(a, d))


def sweep_kernel() -> str:
    return lib().csim_sweep_kernel().decode() if hasattr(lib(), "csim_sweep_kernel") else "k_step_tb"


def exp_variant() -> int:
    """Which restated variant of exp() matches the host libm: 1 FMA, 0 plain, -1 neither."""
    return int(lib().csim_exp_variant())


def exp_restated(x: float, variant: int) -> float:
    return float(lib().csim_exp_restated(x, variant))


def halo_profile(ctx: "Context", enable=True):
    """The next run_steps on a tile with neighbours runs eagerly with timestamps (csim_halo_profile)."""
    _check(lib().csim_halo_profile(ctx._h, 1 if enable else 0))


def halo_path(ctx: "Context") -> str:
    """"peer", "nccl" or "none": the halo path the last run_steps on this context used."""
    return lib().csim_halo_path(ctx._h).decode() if hasattr(lib(), "csim_halo_path") else "nccl"


def halo_stats(ctx: "Context") -> dict:
    st = HaloStats()
    _check(lib().csim_halo_stats_get(ctx._h, C.byref(st)))
    return {name: getattr(st, name) for name, _ in HaloStats._fields_}
