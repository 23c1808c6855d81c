"""CPU tier: the DEVICE header of the forcing preparation (vicres_b200/csrc/vic_atmos.cuh) compiled for the host by
tests/hostsim/hostsim_atmos.cpp, against the reference's own atmos[] records.  Checks the restructured arithmetic
(one thread per (cell, yearday) / (cell, day) / (cell, record) instead of one cell at a time) without a GPU: with
the host's libm it has to be bit-identical.  A test harness, not a product path."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from vicres_b200 import abi, atmos
from common import ATMOS_FIXTURES, ROOT, load_atmos_fixture

HS_DIR = os.path.join(ROOT, "tests", "hostsim")
HS_LIB = os.path.join(HS_DIR, "libhostsim_atmos.so")


@pytest.fixture(scope="module")
def hostsim():
    src = [os.path.join(HS_DIR, "hostsim_atmos.cpp"), os.path.join(ROOT, "vicres_b200", "csrc", "vic_atmos.cuh"),
           os.path.join(ROOT, "include", "vicres_atmos.h")]
    if not os.path.exists(HS_LIB) or any(os.path.getmtime(s) > os.path.getmtime(HS_LIB) for s in src):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c++17", "-Wno-unknown-pragmas",
                        "-o", HS_LIB, src[0], "-lm"], check=True)
    lib = C.CDLL(HS_LIB)
    lib.hs_atmos.argtypes = [C.c_void_p] * 5
    return lib


@pytest.mark.parametrize("name", ATMOS_FIXTURES)
def test_device_arithmetic_on_host_is_bit_identical(hostsim, name):
    fx = load_atmos_fixture(name)
    ps, pr = atmos.pack_site(fx["site"]), atmos.pack_raw(fx["raw"])
    out = np.zeros_like(fx["ref"])
    pd = C.POINTER(C.c_double)
    ptrs = (pd * abi.VR_NFORCING)(*[out[f].ctypes.data_as(pd) for f in range(abi.VR_NFORCING)])
    flag = np.full(fx["ref"].shape[1:2] + fx["ref"].shape[-1:], -1, dtype=np.int32)
    assert hostsim.hs_atmos(C.byref(fx["opt"]), ps.ref(), pr.ref(), ptrs, flag.ctypes.data) == 0
    if fx["snowflag"] is not None:       # SNOW_STEP < TIME_STEP fixtures: [9, nrecs, NF + 1, C] and atmos->snowflag[NR]
        assert out.ndim == 4 and out.shape[2] == atmos.n_windows(fx["opt"]) + 1
        assert np.array_equal(flag, fx["snowflag"]), name
    for f, fname in enumerate(abi.FORCING_FIELDS):
        bad = np.argwhere(out[f] != fx["ref"][f])
        assert bad.size == 0, "%s: %s differs at %d values, first (rec, cell) %s: %r vs %r" % (
            name, fname, len(bad), bad[0], out[f][tuple(bad[0])], fx["ref"][f][tuple(bad[0])])


def test_correctly_rounded_acos_cos_against_mpmath(hostsim):
    """vic_cr_trig.cuh (host build) gives the correctly rounded acos and cos; glibc's own differ from that in
    about 0.1 % of arguments, which bounds how often the device can still take another sunrise branch."""
    mp = pytest.importorskip("mpmath")
    rng = np.random.default_rng(1)
    n = 5000
    x, y = rng.uniform(-0.9999, 0.9999, n), rng.uniform(0.0, np.pi, n)
    a, c = np.zeros(n), np.zeros(n)
    hostsim.hs_cr.argtypes = [C.c_int64] + [C.c_void_p] * 4
    hostsim.hs_cr(n, x.ctypes.data, a.ctypes.data, y.ctypes.data, c.ctypes.data)
    mp.mp.dps = 60
    ra = np.array([float(mp.acos(mp.mpf(float(v)))) for v in x])
    rc = np.array([float(mp.cos(mp.mpf(float(v)))) for v in y])
    assert np.array_equal(a, ra) and np.array_equal(c, rc)
    libm = C.CDLL("libm.so.6")
    libm.acos.restype = libm.cos.restype = C.c_double
    libm.acos.argtypes = libm.cos.argtypes = [C.c_double]
    ga = np.array([libm.acos(float(v)) for v in x])
    gc = np.array([libm.cos(float(v)) for v in y])
    assert (ga != ra).mean() < 0.01 and (gc != rc).mean() < 0.01
