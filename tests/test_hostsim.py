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


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver")),
                    reason="compiled reference (oracle/_ref) not built here")
def test_moistfract_against_the_live_reference(hostsim, tmp_path):
    """MOISTFRACT TRUE (soil moisture as volume fractions per tile before the area weighting, the water-balance storage back
    in mm: put_data.c:624-630, 906-911): the oracle and the device code compiled for the host, bit for bit against the
    compiled reference run on a seeded case"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    import refdump
    from vicres_b200 import synth
    g = synth.make_case(str(tmp_path / "case"), ncells=3, nlayer=3, nbands=2, ndays=150, cold=8.0, lat_range=(30, 60),
                        overstory_frac=0.5, seed=777, extra_global="MOISTFRACT TRUE")
    dump = str(tmp_path / "dump.bin")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver"), "-g", g, "-d", dump], check=True, capture_output=True)
    d = refdump.read_dump(dump)
    cs = refdump.case_from_dump(d)
    names = abi.OUT_NAMES
    opt = abi.default_options(cs["nlayer"], cs["nbands"], cs["dt"], out=names, max_snow_temp=cs["max_snow_temp"],
                              min_rain_temp=cs["min_rain_temp"], moistfract=True)
    ref = refdump.ref_outputs(d, names)
    assert ref[names.index("SOIL_LIQ0")].max() < 1.0          # fractions, not mm
    o = oracle_lib.Oracle(opt, cs["lib"], cs["cells"])
    o.init_state(cs["tair0"])
    rc, out, _ = o.step(cs["month"], cs["doy"], cs["forcing"])
    o.close()
    assert rc == 0 and np.array_equal(out, ref)
    lib, cells = abi.pack_veglib(cs["lib"]), abi.pack_cells(cs["cells"])
    f = abi.pack_forcing(cs["month"], cs["doy"], cs["forcing"], None)
    hout = np.zeros_like(ref)
    t0 = np.ascontiguousarray(cs["tair0"])
    assert hostsim.hs_run(C.byref(opt), lib.ref(), cells.ref(), f.ref(), t0.ctypes.data, hout.ctypes.data) == 0
    for i, n in enumerate(names):
        assert np.array_equal(hout[i], ref[i]), "%s differs from the reference under MOISTFRACT" % n


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver")),
                    reason="compiled reference (oracle/_ref) not built here")
def test_organic_fract_against_the_live_reference(hostsim, tmp_path):
    """ORGANIC_FRACT TRUE: the soil file's organic fraction and organic densities (read_soilparam.c:444-506, 647-655) through
    the parameter reader, and the step with organic soil layers (thermal conductivity and heat capacity of the solids,
    porosity and maximum moisture) -- reader, oracle and host build of the device code bit for bit against the compiled
    reference"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    import refdump
    from vicres_b200 import files, synth
    g = synth.make_case(str(tmp_path / "case"), ncells=3, nlayer=3, nbands=2, ndays=150, cold=8.0, lat_range=(30, 60),
                        overstory_frac=0.5, seed=778, organic=True)
    dump = str(tmp_path / "dump.bin")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver"), "-g", g, "-d", dump], check=True, capture_output=True)
    d = refdump.read_dump(dump)
    cs = refdump.case_from_dump(d)
    G = files.Global(g)
    assert G.unsupported is None and G.organic_fract == 1
    _, cells, _ = files.read_params(**G.param_files())
    assert cells["organic"][0].max() > 0.1
    for k in ("organic", "bulk_density", "soil_density", "porosity", "max_moist", "Wcr", "Wpwp", "bulk_dens_min", "soil_dens_min"):
        assert np.array_equal(cells[k], cs["cells"][k]), k
    names = abi.OUT_NAMES
    opt = abi.default_options(cs["nlayer"], cs["nbands"], cs["dt"], out=names, max_snow_temp=cs["max_snow_temp"],
                              min_rain_temp=cs["min_rain_temp"])
    ref = refdump.ref_outputs(d, names)
    o = oracle_lib.Oracle(opt, cs["lib"], cs["cells"])
    o.init_state(cs["tair0"])
    rc, out, _ = o.step(cs["month"], cs["doy"], cs["forcing"])
    o.close()
    assert rc == 0 and np.array_equal(out, ref)
    lib, pc = abi.pack_veglib(cs["lib"]), abi.pack_cells(cells)         # the cells as THIS library's reader delivers them
    f = abi.pack_forcing(cs["month"], cs["doy"], cs["forcing"], None)
    hout = np.zeros_like(ref)
    t0 = np.ascontiguousarray(cs["tair0"])
    assert hostsim.hs_run(C.byref(opt), lib.ref(), pc.ref(), f.ref(), t0.ctypes.data, hout.ctypes.data) == 0
    for i, n in enumerate(names):
        assert np.array_equal(hout[i], ref[i]), "%s differs from the reference under ORGANIC_FRACT" % n
