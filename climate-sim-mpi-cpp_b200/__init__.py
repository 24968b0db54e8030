"""climate-sim-mpi-cpp_b200 — the B200-native timestep hot path of climate-sim-mpi-cpp.

Contents: ``csrc/`` (sm_100a CUDA kernels + the C ABI of ``include/csim.h``), ``host/`` (C++ drop-in
headers with the reference's names, the driver) and ``binding.py`` (ctypes mirror for tests/bench).
Import with ``importlib.import_module("climate-sim-mpi-cpp_b200")`` (the name has hyphens).
"""
from .binding import *  # noqa: F401,F403
from .binding import lib, LIB_PATH  # noqa: F401
