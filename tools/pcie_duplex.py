#!/usr/bin/env python
"""PCIe probe for the e2e path: H2D alone, D2H alone, both at once (two streams), and a pitched 2-D
H2D like csim_field_upload's.  Prints GB/s.  Run on the GPU box: python tools/pcie_duplex.py"""
import ctypes
import time

import torch

n = 8194 * 8194
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda")
d_b = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 8 / 1e9


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


t = timed(h2d)
print(f"H2D alone      {gb / t:6.1f} GB/s  ({t * 1e3:.2f} ms)")
t = timed(d2h)
print(f"D2H alone      {gb / t:6.1f} GB/s  ({t * 1e3:.2f} ms)")
t = timed(both)
print(f"H2D+D2H duplex {2 * gb / t:6.1f} GB/s total ({t * 1e3:.2f} ms for both)")

# pitched 2-D copy as csim_field_upload does it (host rows of nx+2 doubles → device pitch)
rt = ctypes.CDLL("libcudart.so.12")
nx = 8194
pitch = (16 + 8192 + 16 + 15) // 16 * 16
d_p = torch.empty(pitch * (8192 + 16), dtype=torch.float64, device="cuda")


def h2d_2d():
    rt.cudaMemcpy2DAsync(ctypes.c_void_p(d_p.data_ptr()), ctypes.c_size_t(pitch * 8), ctypes.c_void_p(h_in.data_ptr()),
                         ctypes.c_size_t(nx * 8), ctypes.c_size_t(nx * 8), ctypes.c_size_t(8194), 1,
                         ctypes.c_void_p(s1.cuda_stream))


t = timed(h2d_2d)
print(f"H2D 2-D pitched {gb / t:6.1f} GB/s  ({t * 1e3:.2f} ms)")


def both_2d():
    h2d_2d()
    d2h()


t = timed(both_2d)
print(f"2-D H2D + D2H   {2 * gb / t:6.1f} GB/s total ({t * 1e3:.2f} ms for both)")
