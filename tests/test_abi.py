"""CPU tier: the C-ABI shared library loads and exports every symbol include/vicres.h declares, the
ctypes struct mirrors have the C layout, and the library refuses to run without a CUDA device (no
CPU fallback).  No compute calls."""
import ctypes as C
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
    assert len(syms) == 15
    for s in syms:
        assert hasattr(lib, s), "libvicres.so does not export %s" % s
    assert sorted(files.FILES_EXPORTS) == syms


def test_exports_every_declared_routing_symbol(lib):
    from vicres_b200 import route
    syms = declared_symbols("vicres_route.h")
    for s in syms:
        assert hasattr(lib, s), "libvicres.so does not export %s" % s
    assert sorted(route.ROUTE_EXPORTS) == syms


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
    bad.snow_step_hours = 1                     # NF > 1: outside the supported path
    assert lib.vr_create(C.byref(bad), 0, C.byref(h)) == abi.VR_ERR_UNSUPPORTED
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
