"""pytest configuration.  `-m "not gpu"` runs on the CPU-only build box; `-m gpu` on a B200."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def csim():
    """The product package (ctypes binding of libcsim_b200.so).  Built by __graft_entry__.build()."""
    if not os.path.exists(os.path.join(ROOT, "climate-sim-mpi-cpp_b200", "libcsim_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("climate-sim-mpi-cpp_b200")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import cpu_oracle
    if not cpu_oracle.available("port"):
        cpu_oracle.build()
    return cpu_oracle


@pytest.fixture(scope="session")
def port(oracle_mod):
    return oracle_mod.Oracle("port")


@pytest.fixture(scope="session")
def ref(oracle_mod):
    """The reference's own objects.  Prebuilt .so travels to the GPU box; skip if absent."""
    if not oracle_mod.available("ref"):
        if os.path.exists("/root/reference/src/diffusion.cpp"):
            oracle_mod.build()
        else:
            pytest.skip("oracle/_ref/libcsim_ref.so not built and /root/reference absent")
    return oracle_mod.Oracle("ref")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "timestep_golden.npz")
    z = np.load(path)
    cases = {}
    for name in z["names"]:
        name = str(name)
        meta, phys = z[name + "/meta"], z[name + "/phys"]
        cases[name] = dict(
            nx=int(meta[0]), ny=int(meta[1]), steps=int(meta[2]), out_every=int(meta[3]),
            bc=tuple(int(b) for b in meta[4:8]), ic_preset=int(meta[8]), nranks=int(meta[9]),
            dx=float(phys[0]), dy=float(phys[1]), D=float(phys[2]), vx=float(phys[3]), vy=float(phys[4]),
            dt=float(phys[5]), A=float(phys[6]), sigma_frac=float(phys[7]), xc_frac=float(phys[8]),
            yc_frac=float(phys[9]),
            u0=z[name + "/u0"] if name + "/u0" in z.files else None,
            final=z[name + "/final"], frames=z[name + "/frames"],
            padded=z[name + "/padded"] if name + "/padded" in z.files else None)
    return cases


@pytest.fixture(scope="session")
def ctx(csim):
    c = csim.Context(0)
    yield c
    c.close()


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def mask_corners(a):
    """Copy with the four corner ghosts zeroed (indeterminate in the reference)."""
    a = np.array(a, copy=True)
    a[0, 0] = a[0, -1] = a[-1, 0] = a[-1, -1] = 0.0
    return a
