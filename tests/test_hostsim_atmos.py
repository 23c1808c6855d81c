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
    lib.hs_atmos.argtypes = [C.c_void_p] * 4
    return lib


@pytest.mark.parametrize("name", ATMOS_FIXTURES)
def test_device_arithmetic_on_host_is_bit_identical(hostsim, name):
    fx = load_atmos_fixture(name)
    ps, pr = atmos.pack_site(fx["site"]), atmos.pack_raw(fx["raw"])
    out = np.zeros_like(fx["ref"])
    pd = C.POINTER(C.c_double)
    ptrs = (pd * abi.VR_NFORCING)(*[out[f].ctypes.data_as(pd) for f in range(abi.VR_NFORCING)])
    assert hostsim.hs_atmos(C.byref(fx["opt"]), ps.ref(), pr.ref(), ptrs) == 0
    for f, fname in enumerate(abi.FORCING_FIELDS):
        bad = np.argwhere(out[f] != fx["ref"][f])
        assert bad.size == 0, "%s: %s differs at %d values, first (rec, cell) %s: %r vs %r" % (
            name, fname, len(bad), bad[0], out[f][tuple(bad[0])], fx["ref"][f][tuple(bad[0])])
