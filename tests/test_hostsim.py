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
    f = abi.pack_forcing(fx["month"], fx["doy"], fx["forcing"], fx["snowflag"])
    out = np.zeros_like(fx["ref"])
    t0 = np.ascontiguousarray(fx["tair0"])
    assert hostsim.hs_run(C.byref(fx["opt"]), lib.ref(), cells.ref(), f.ref(), t0.ctypes.data, out.ctypes.data) == 0
    for i, n in enumerate(fx["names"]):
        assert np.array_equal(out[i], fx["ref"][i]), "%s differs from the reference (%s)" % (n, name)


def test_pow_table_error_bound(tmp_path):
    """The arithmetic of the device power function (vr::pow_table_eval: table-driven log2 / exp2) compiled for the
    host, against powl() on 2e6 seeded arguments of the model's ranges: relative error <= 2e-16 * (2 + |y log2 x|),
    which is < 1e-13 for every argument the step produces."""
    exe = str(tmp_path / "pow_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-o", exe,
                    os.path.join(HS_DIR, "pow_table_check.cpp"), "-lm"], check=True)
    n, norm, rel, unhandled = subprocess.run([exe, "2000000"], check=True, capture_output=True, text=True).stdout.split()
    assert int(unhandled) == 0
    assert float(norm) <= 2e-16 and float(rel) < 1e-13
