"""CPU tier: the C-ABI shared library loads and exports every symbol include/vicres.h declares, the
ctypes struct mirrors have the C layout, and the library refuses to run without a CUDA device (no
CPU fallback).  No compute calls."""
import ctypes as C
import numpy as np
import os
import re
import subprocess
import pytest
from vicres_b200 import abi, build, model
from common import ROOT


@pytest.fixture(scope="module")
def lib():
    build.build()
    return model.load_library()


def declared_symbols(header="vicres.h"):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vr_[a-z_0-9]+)\s*\(", txt)))


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), "libvicres.so does not export %s" % s
    assert sorted(model.EXPORTS) == syms


def test_exports_every_declared_file_symbol(lib):
    from vicres_b200 import files
    syms = declared_symbols("vicres_files.h")
    assert len(syms) == 27
    for s in syms:
        assert hasattr(lib, s), "libvicres.so does not export %s" % s
    assert sorted(files.FILES_EXPORTS) == syms


def test_exports_every_declared_routing_symbol(lib):
    from vicres_b200 import route
    syms = declared_symbols("vicres_route.h")
    for s in syms:
        assert hasattr(lib, s), "libvicres.so does not export %s" % s
    assert sorted(route.ROUTE_EXPORTS) == syms


def test_exports_every_declared_atmos_symbol(lib):
    from vicres_b200 import atmos
    syms = declared_symbols("vicres_atmos.h")
    assert len(syms) == 10
    for s in syms:
        assert hasattr(lib, s), "libvicres.so does not export %s" % s
    assert sorted(atmos.EXPORTS) == syms


def test_atmos_struct_layout_and_defaults(lib, tmp_path):
    from vicres_b200 import atmos
    src = tmp_path / "sza.c"
    src.write_text('#include <stdio.h>\n#include "vicres_atmos.h"\nint main(){printf("%zu %zu %zu\\n",'
                   "sizeof(vr_atmos_options),sizeof(vr_atmos_cells),sizeof(vr_atmos_raw));return 0;}\n")
    exe = tmp_path / "sza"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(atmos.vr_atmos_options), C.sizeof(atmos.vr_atmos_cells), C.sizeof(atmos.vr_atmos_raw)]
    o = atmos.vr_atmos_options()
    atmos._lib().vr_atmos_default_options(C.byref(o))
    mine = atmos.options(nrecs=0)
    for name, _ in atmos.vr_atmos_options._fields_:
        if not name.startswith("reserved"):
            assert getattr(o, name) == getattr(mine, name), name


def test_atmos_has_no_cpu_fallback_and_validates_without_a_device(lib):
    import torch
    from vicres_b200 import atmos
    L = atmos._lib()
    o = atmos.options(nrecs=10, layout=atmos.DAILY_TMAX_TMIN, force_dt=12)
    assert L.vr_atmos_prepare(0, C.byref(o), None, None, None) == abi.VR_ERR_ARG
    z = np.zeros(10)
    site = atmos.pack_site({"lat": z[:1], "lng": z[:1], "time_zone_lng": z[:1], "elevation": z[:1], "annual_prec": z[:1]})
    raw = atmos.pack_raw({"prec": z[:, None], "t_a": z[:, None], "t_b": z[:, None], "wind": z[:, None]})
    out = np.zeros((abi.VR_NFORCING, 10, 1))
    pd = C.POINTER(C.c_double)
    ptrs = (pd * abi.VR_NFORCING)(*[out[f].ctypes.data_as(pd) for f in range(abi.VR_NFORCING)])
    assert L.vr_atmos_prepare(0, C.byref(o), site.ref(), raw.ref(), ptrs) == abi.VR_ERR_UNSUPPORTED
    assert b"FORCE_DT" in L.vr_atmos_last_error()
    if not torch.cuda.is_available():
        o = atmos.options(nrecs=10)
        assert L.vr_atmos_prepare(0, C.byref(o), site.ref(), raw.ref(), ptrs) == abi.VR_ERR_CUDA
        assert b"no CPU path" in L.vr_atmos_last_error()


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "vicres.h"\nint main(){printf("%zu %zu %zu %zu %d %d\\n",'
                   "sizeof(vr_options),sizeof(vr_veglib),sizeof(vr_cells),sizeof(vr_forcing),VR_NFORCING,VR_NOUTVAR);return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(abi.vr_options), C.sizeof(abi.vr_veglib), C.sizeof(abi.vr_cells),
                   C.sizeof(abi.vr_forcing), abi.VR_NFORCING, abi.VR_NOUTVAR]


def test_default_options_match_reference_default_output_list(lib):
    o = abi.vr_options()
    lib.vr_default_options(C.byref(o), 3)
    mine = abi.default_options(3)
    assert o.n_out == mine.n_out == 26          # 19 + nlayer flux columns + 4 snow columns
    assert list(o.out_vars)[:26] == list(mine.out_vars)[:26]
    assert (o.nlayer, o.nbands, o.dt_hours, o.snow_step_hours) == (3, 1, 24, 24)


def test_argument_validation_needs_no_device(lib):
    h = C.c_void_p()
    bad = abi.default_options(3)
    bad.snow_step_hours = 5                     # SNOW_STEP must divide TIME_STEP (get_global_param.c:761-770)
    assert lib.vr_create(C.byref(bad), 0, C.byref(h)) == abi.VR_ERR_ARG
    bad = abi.default_options(3)
    bad.nlayer = 5
    assert lib.vr_create(C.byref(bad), 0, C.byref(h)) == abi.VR_ERR_ARG
    assert b"nlayer" in lib.vr_last_error(None)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.vr_create(C.byref(abi.default_options(3)), 0, C.byref(h))
    assert rc == abi.VR_ERR_CUDA and not h.value
    assert b"no CPU path" in lib.vr_last_error(None)
