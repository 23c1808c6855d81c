"""CPU tier: the DEVICE physics headers (vicres_b200/csrc/vic_physics.cuh, vic_aggregate.cuh,
vic_host_prep.h) compiled for the host by tests/hostsim, against the golden fixtures.  This checks
the kernel's arithmetic (flattened step, hoisted constants, ordered aggregation) without a GPU; it
is a test harness, not a product path."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from vicres_b200 import abi
from common import FIXTURES, ROOT, load_fixture

HS_DIR = os.path.join(ROOT, "tests", "hostsim")
HS_LIB = os.path.join(HS_DIR, "libhostsim.so")


@pytest.fixture(scope="module")
def hostsim():
    src = [os.path.join(HS_DIR, "hostsim.cpp")] + [os.path.join(ROOT, "vicres_b200", "csrc", f)
                                                   for f in ("vic_physics.cuh", "vic_aggregate.cuh", "vic_host_prep.h")]
    if not os.path.exists(HS_LIB) or any(os.path.getmtime(s) > os.path.getmtime(HS_LIB) for s in src):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c++17", "-Wno-unknown-pragmas",
                        "-o", HS_LIB, src[0], "-lm"], check=True)
    lib = C.CDLL(HS_LIB)
    lib.hs_run.argtypes = [C.c_void_p] * 6
    return lib


@pytest.mark.parametrize("name", FIXTURES)
def test_device_physics_on_host_matches_golden(hostsim, name):
    fx = load_fixture(name)
    lib, cells = abi.pack_veglib(fx["lib"]), abi.pack_cells(fx["cells"])
    f = abi.pack_forcing(fx["month"], fx["doy"], fx["forcing"])
    out = np.zeros_like(fx["ref"])
    t0 = np.ascontiguousarray(fx["tair0"])
    assert hostsim.hs_run(C.byref(fx["opt"]), lib.ref(), cells.ref(), f.ref(), t0.ctypes.data, out.ctypes.data) == 0
    for i, n in enumerate(fx["names"]):
        assert np.array_equal(out[i], fx["ref"][i]), "%s differs from the reference (%s)" % (n, name)
