"""The whole `vicNl -g <global file>` command line on the GPU (vicres_b200/vicnl.py: global file -> parameter files
-> forcing files -> forcing preparation -> cell step -> result files) against the files the UNMODIFIED reference
wrote for the same case (tests/golden/vicnl_case, tools/make_vicnl_golden.py)."""
import os
import shutil
import numpy as np
import pytest
from common import GOLDEN
from vicres_b200 import files

CASES = os.path.join(GOLDEN, "vicnl_case")


def _stage(name, tmp_path):
    """copy the inputs, point the global file at the copy, drop the reference's results"""
    dst = tmp_path / name
    shutil.copytree(os.path.join(CASES, name), dst)
    ref = tmp_path / (name + "_ref")
    shutil.move(str(dst / "results"), ref)
    os.makedirs(dst / "results")
    g = dst / "global.txt"
    g.write_text(g.read_text().replace("@DIR@", str(dst)))
    return str(g), str(dst / "results"), str(ref)


@pytest.mark.parametrize("name", ["daily", "sub3h"])
def test_global_file_and_forcing_files_parse(name, tmp_path):
    """CPU tier: the host side of the driver up to the device hand-off."""
    from vicres_b200 import atmos
    gpath, _, ref = _stage(name, tmp_path)
    g = files.Global(gpath)
    assert g.unsupported is None
    lib, cells, meta = files.read_params(**g.param_files())
    ao = atmos.options_from_global(g)
    assert ao.layout == (atmos.DAILY_TMAX_TMIN if name == "daily" else atmos.SUBDAILY_AIR_TEMP)
    assert (ao.vp_interp, ao.vp_iter, ao.lw_type, ao.lw_cloud, ao.plapse) == (1, 1, 5, 1, 1)
    d, got = g.read_forcing(float(meta["lat"][0]), float(meta["lng"][0]))
    assert got == g.nrecs * g.time_step // g.force_dt and set(d) >= {"PREC", "WIND"}
    site = meta["site"]
    assert np.array_equal(site["lat"], meta["lat"].astype(np.float64)) and np.array_equal(site["lng"], meta["lng"].astype(np.float64))
    if name == "daily":
        assert np.all(site["time_zone_lng"] == 15.0) and np.all(site["annual_prec"] == 1000.0)
    assert len(os.listdir(ref)) == 2 * len(meta["gridcel"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["daily", "sub3h"])
def test_vicnl_command_line_reproduces_the_reference_result_files(name, tmp_path):
    from vicres_b200 import vicnl
    gpath, res, ref = _stage(name, tmp_path)
    assert vicnl.main(["-g", gpath]) == 0
    assert sorted(os.listdir(res)) == sorted(os.listdir(ref))
    nbad = ntot = 0
    for n in sorted(os.listdir(ref)):
        a, b = np.loadtxt(os.path.join(res, n)), np.loadtxt(os.path.join(ref, n))
        assert a.shape == b.shape, n
        ncal = 3 if name == "daily" else 4
        assert np.array_equal(a[:, :ncal], b[:, :ncal]), n
        assert np.all(np.abs(a - b) <= 1.0001e-4 + 1e-5 * np.abs(b)), n     # one unit in the 4th decimal
        nbad += int((a != b).sum())
        ntot += a.size
    print("VICNL %s: %d of %d printed values differ from the reference's files (by one unit of the last digit)" % (name, nbad, ntot))
    assert nbad <= 0.01 * ntot


def _build_c_driver(tmp_path):
    import subprocess
    from common import ROOT
    from vicres_b200 import build
    build.build()
    exe = str(tmp_path / "vicnl_gpu")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "host", "vicnl_gpu.c"), "-L", os.path.join(ROOT, "vicres_b200"), "-lvicres",
                    "-Wl,-rpath," + os.path.join(ROOT, "vicres_b200")], check=True)
    return exe


def test_c_host_driver_compiles_against_the_headers(tmp_path):
    """CPU tier: host/vicnl_gpu.c is plain C over include/*.h and links against libvicres.so; without a GPU it stops
    at the first device call with the library's message (no CPU path)."""
    import subprocess
    import torch
    exe = _build_c_driver(tmp_path)
    gpath, _, _ = _stage("daily", tmp_path)
    r = subprocess.run([exe, "-g", gpath], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CPU path" in r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["daily", "sub3h"])
def test_c_host_driver_reproduces_the_reference_result_files(name, tmp_path):
    """the same command line from the plain C host program (host/vicnl_gpu.c), records in blocks of 50"""
    import subprocess
    exe = _build_c_driver(tmp_path)
    gpath, res, ref = _stage(name, tmp_path)
    r = subprocess.run([exe, "-g", gpath, "--block", "50"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert sorted(os.listdir(res)) == sorted(os.listdir(ref))
    nbad = ntot = 0
    for n in sorted(os.listdir(ref)):
        a, b = np.loadtxt(os.path.join(res, n)), np.loadtxt(os.path.join(ref, n))
        assert a.shape == b.shape, n
        assert np.all(np.abs(a - b) <= 1.0001e-4 + 1e-5 * np.abs(b)), n
        nbad += int((a != b).sum())
        ntot += a.size
    print("VICNL_GPU (C host) %s: %d of %d printed values differ from the reference's files" % (name, nbad, ntot))
    assert nbad <= 0.01 * ntot
