"""Minimal reader for NetCDF CDF-5 files with one record variable, for round-trip tests of the
snapshot writer (netCDF4 / scipy's netcdf_file cannot be used here: absent / CDF-1,2 only)."""
import struct

import numpy as np


def read_cdf5(path):
    b = open(path, "rb").read()
    assert b[:4] == b"CDF\x05", b[:4]
    p = 4

    def i64():
        nonlocal p
        v = struct.unpack(">q", b[p:p + 8])[0]
        p += 8
        return v

    def u32():
        nonlocal p
        v = struct.unpack(">I", b[p:p + 4])[0]
        p += 4
        return v

    def name():
        nonlocal p
        n = i64()
        s = b[p:p + n].decode()
        p += (n + 3) // 4 * 4
        return s

    numrecs = i64()
    dims = []
    tag, n = u32(), i64()
    assert (tag, n) == (0, 0) or tag == 0x0A
    for _ in range(n):
        dims.append((name(), i64()))
    atts = {}
    tag, n = u32(), i64()
    assert (tag, n) == (0, 0) or tag == 0x0C
    for _ in range(n):
        k = name()
        assert u32() == 2  # NC_CHAR
        atts[k] = name()
    tag, nvars = u32(), i64()
    assert tag == 0x0B and nvars == 1
    vname = name()
    ndims = i64()
    dimids = [i64() for _ in range(ndims)]
    vt, vn = u32(), i64()
    assert (vt, vn) == (0, 0)
    assert u32() == 6  # NC_DOUBLE
    vsize, begin = i64(), i64()
    shape = [dims[d][1] for d in dimids]
    assert shape[0] == 0 and vsize == shape[1] * shape[2] * 8
    data = np.frombuffer(b, dtype=">f8", count=numrecs * shape[1] * shape[2], offset=begin)
    return dict(numrecs=numrecs, dims=dims, attrs=atts, var=vname,
                data=data.reshape(numrecs, shape[1], shape[2]).astype(np.float64), begin=begin)
