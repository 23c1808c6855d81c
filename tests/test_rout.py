"""The `rout <configuration file>` command line on the GPU (vicres_b200/rout.py) against the reference's shipped
prebuilt binary: the case files are written from the golden fixtures (which hold the binary's inputs and outputs,
tools/make_route_golden.py), in today's configuration layout and in the binary's older one."""
import os
import numpy as np
import pytest
from common import GOLDEN
from vicres_b200 import rout, route_abi


def write_case(work, z, layout, sbs=None):
    files_ = list(z.files)
    z = {k: z[k] for k in files_}              # an NpzFile decompresses an array on every access
    nrow, ncol = z["dir"].shape
    hdr = {"xllcorner": repr(float(z["hdr"][0])), "yllcorner": repr(float(z["hdr"][1])), "cellsize": repr(float(z["hdr"][2]))}
    par = os.path.join(work, "parameters")
    for sub in ("", "unit_hydrograph", "reservoirs/res_par", "naturalized_flow"):
        os.makedirs(os.path.join(par, sub), exist_ok=True)
    for sub in ("input", "output", "model"):
        os.makedirs(os.path.join(work, sub), exist_ok=True)

    def esri(path, a, fmt, header=True):
        with open(path, "w") as f:
            if header:
                f.write("ncols %d\nnrows %d\nxllcorner %s\nyllcorner %s\ncellsize %s\nNODATA_value 0\n" %
                        (ncol, nrow, hdr["xllcorner"], hdr["yllcorner"], hdr["cellsize"]))
            for r in range(nrow):
                f.write(" ".join(fmt % x for x in a[r]) + "\n")
    esri(os.path.join(par, "flowdirection.txt"), z["dir"], "%d")
    esri(os.path.join(par, "reservoirlocation.txt"), z["reser"], "%d", header=False)
    esri(os.path.join(par, "velocity.txt"), z["velo"], "%.9g")
    esri(os.path.join(par, "diffusion.txt"), z["diff"], "%.9g")
    esri(os.path.join(par, "xmask.txt"), z["xmask"], "%.9g")
    r, c = [int(x) for x in z["station_rc"]]
    with open(os.path.join(par, "stations.txt"), "w") as f:
        f.write("1 OUTLT %d %d -9999\nNONE\n" % (c + 1, nrow - r))
    with open(os.path.join(par, "unit_hydrograph", "UH.all"), "w") as f:
        for k in range(12):
            f.write("%d %.9g\n" % (k, z["uh_box"][k]))
    for k in files_:
        if k.startswith("restxt_"):
            with open(os.path.join(par, "reservoirs", "res_par", "res%s.txt" % k[7:]), "w") as f:
                f.write(str(z[k]))
    ndays = int(z["ndays"])
    dates = [(2000, mo, dd) for mo, nd in zip(range(1, 6), (31, 29, 31, 30, 31)) for dd in range(1, nd + 1)][:ndays]
    for g in np.flatnonzero(z["present"]):
        rr, cc = divmod(int(g), ncol)
        lat, lon = route_abi.cell_latlon(hdr, nrow, ncol, rr, cc)
        with open(os.path.join(work, "input", "fluxes_%.4f_%.4f" % (lat, lon)), "w") as f:
            for day, (y, mo, dd) in enumerate(dates):
                ro, ba, evv, trv, evs = (z[k][g, day] for k in ("runoff", "baseflow", "ev_vege", "tr_vege", "ev_soil"))
                cols = [0.0, evv + trv + evs, ro, ba, 0.0, 10.0, 50.0, 55.0, -26.0, evv, trv, evs, 0.0, 0.0, 5.7, 15.0, 0.12,
                        z["rh"][g, day], 309.0, z["tair"][g, day], z["wind"][g, day]]
                f.write("%d\t%d\t%d\t" % (y, mo, dd) + "\t".join("%.4f" % x for x in cols) + "\n")
    cfg = ["# Routing main file", "# NAME OF FLOW DIRECTION FILE", "../parameters/flowdirection.txt",
           "# NAME OF VELOCITY FILE", ".true.", "../parameters/velocity.txt", "# NAME OF DIFF FILE", ".true.",
           "../parameters/diffusion.txt", "# NAME OF XMASK FILE", ".true.", "../parameters/xmask.txt",
           "# NAME OF FRACTION FILE", ".false.", "1.0", "# NAME OF STATION FILE", "../parameters/stations.txt"]
    if layout == "current":
        cfg += ["# NAME OF GRID LOCATION FILE", ".false.", "0.0", "0.0"]
    cfg += ["# PATH OF INPUT FILES AND PRECISION", "../input/fluxes_", "4"]
    if layout == "current":
        cfg += ["classic"]
    cfg += ["# PATH OF OUTPUT FILES", "../output/", "# RESERVOIR LOCATION", "../parameters/reservoirlocation.txt",
            "# RESERVOIR PARAMETERS", "../parameters/reservoirs/res_par/", "# NATURALIZED FLOW", ".false. ",
            "../parameters/naturalized_flow/ ", "# YEAR AND MONTH OF MODEL OUTPUT TO ROUTE & ROUTED OUTPUT TO WRITE         ",
            "2000 1 2000 5 ", "2000 1 2000 5 ", "# NAME OF UNIT HYDROGRAPH FILE         ", "../parameters/unit_hydrograph/",
            "# Simulation Mode: True: Run Step-by-Step Mode; False: Run all Steps", ".true." if sbs else ".false.",
            "# Start Day and Current Simulation Day (for STEPBYSTEP Version)",
            "2000 1 1 %d %d %d" % (sbs if sbs else (2000, 5, 31)), "# Coupler Iteration", "1",
            "# Write Outputs", ".true."]
    path = os.path.join(par, "configuration.txt")
    with open(path, "w") as f:
        f.write("\n".join(cfg) + "\n")
    return path


@pytest.mark.parametrize("layout", ["current", "binary"])
def test_configuration_file_parses_in_both_layouts(layout, tmp_path):
    z = np.load(os.path.join(GOLDEN, "route_toy.npz"))
    cfg = rout.read_configuration(write_case(str(tmp_path), z, layout))
    assert cfg["layout"] == layout and cfg["flowdirection"] == "../parameters/flowdirection.txt"
    assert cfg["velocity"] == "../parameters/velocity.txt" and cfg["fraction"] == 1.0 and cfg["dprec"] == 4
    assert (cfg["start_year"], cfg["start_month"], cfg["stop_year"], cfg["stop_month"]) == (2000, 1, 2000, 5)
    assert cfg["uh_path"] == "../parameters/unit_hydrograph/" and not cfg["stepbystep"] and not cfg["nf_exist"]
    assert rout.period(cfg)[0] == 152
    st = rout.read_stations(os.path.join(str(tmp_path), "parameters", "stations.txt"))
    assert len(st) == 1 and st[0]["name"] == "OUTLT"


def test_reference_bundled_configuration_parses():
    """the configuration.txt the reference ships (today's layout)"""
    p = "/root/reference/routing/parameters/configuration.txt"
    if not os.path.exists(p):
        pytest.skip("reference tree not present")
    cfg = rout.read_configuration(p)
    assert cfg["layout"] == "current" and cfg["vicdriver"] == "classic" and cfg["velocity"] == 1.2 and cfg["diffusion"] == 800.0
    assert cfg["grid_location"] == ("../parameters/grid_lat.txt", "../parameters/grid_lon.txt")
    assert cfg["inpath"] == "../input/fluxes_" and cfg["outpath"] == "../output/"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["route_toy", "route_rand12x10", "route_rand20x16"])
def test_rout_command_line_reproduces_the_binary_outputs(name, tmp_path):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfgp = write_case(str(tmp_path), z, "current")
    model_dir = os.path.join(str(tmp_path), "model")
    cwd = os.getcwd()
    os.chdir(model_dir)                       # the reference resolves the configuration's paths from the model directory
    try:
        assert rout.main(["../parameters/configuration.txt", "--legacy"]) == 0
    finally:
        os.chdir(cwd)
    out = os.path.join(str(tmp_path), "output")
    flow = np.loadtxt(os.path.join(out, "OUTLT.day"))
    assert flow.shape == (int(z["ndays"]), 4) and tuple(flow[0, :3]) == (2000, 1, 1) and tuple(flow[-1, :3]) == (2000, 5, 31)
    ref = z["station_flow"]
    err = np.abs(flow[:, 3] - ref) / np.maximum(np.abs(ref), 1e-2 * np.abs(ref).max())
    assert err.max() < 1e-4, "station discharge vs the binary: %.3e" % err.max()
    nres = 0
    for k in z.files:
        if not k.startswith("resday_"):
            continue
        a = np.loadtxt(os.path.join(out, "reservoir_%s_OUTLT.day" % k[7:]), skiprows=1)[:, 3:]
        b = z[k]
        assert a.shape == b.shape
        scale = np.maximum(np.abs(b), 1e-2 * np.abs(b).max(axis=0, keepdims=True) + 1e-2)
        err = np.abs(a - b) / scale
        assert err[:, :5].max() < 2e-5, "reservoir %s storage/level columns vs the binary: %.3e" % (k[7:], err[:, :5].max())
        # releases are differences of storages divided by 86.4: one fp32 ulp of a 1e5 storage is 1e-4 m3/s
        assert err[:, 5:].max() < 1e-3, "reservoir %s flow/energy columns vs the binary: %.3e" % (k[7:], err[:, 5:].max())
        nres += 1
    print("ROUT %s: station discharge max rel %.2e vs the binary, %d reservoir files" % (name, err.max(), nres))
    # WRITE_DATA's mm file and WRITE_MONTH's summaries (write_routines.f:64-70, 248-331) against the binary's
    for ext, col in (("day_mm", 3), ("month", 2), ("month_mm", 2), ("year", 1), ("year_mm", 1)):
        a = np.loadtxt(os.path.join(out, "OUTLT." + ext), ndmin=2)
        b = z["station_" + ext]
        if name == "route_toy" and ext != "day_mm":     # the binary's own monthly summaries of the toy basin are NaN /
            continue                                     # uninitialised (its FIRST/LAST months lie outside the 5 months run)
        # (the binary prints 9999 99 99 for the dates of its .day_mm: only the values are compared there)
        assert a.shape == b.shape and (ext == "day_mm" or np.array_equal(a[:, :col], b[:, :col])), ext
        e = np.abs(a[:, col] - b[:, col]) / np.maximum(np.abs(b[:, col]), 1e-2 * np.abs(b[:, col]).max() + 1e-30)
        assert e.max() < 1e-4, "%s vs the binary: %.3e" % (ext, e.max())
    assert os.path.exists(os.path.join(out, "OUTLT.end_of_month"))


@pytest.mark.gpu
def test_rout_step_by_step_mode_equals_all_steps(tmp_path):
    """STEPBYSTEP (rout.f:255-317, reservoir.f:1142-1150, 1596-1675): one simulated day per run of the command line,
    state in the reference's text files; the appended STEPBYSTEP/<station>.day must hold, line by line, what the
    all-steps run writes for the same days (chained reservoirs: rand20x16 has four)."""
    z = np.load(os.path.join(GOLDEN, "route_rand20x16.npz"))
    work = str(tmp_path)
    cfgp = write_case(work, z, "current")
    model_dir = os.path.join(work, "model")
    cwd = os.getcwd()
    os.chdir(model_dir)
    try:
        assert rout.main(["../parameters/configuration.txt"]) == 0
        whole = open(os.path.join(work, "output", "OUTLT.day")).read().splitlines()
        whole_mm = open(os.path.join(work, "output", "OUTLT.day_mm")).read().splitlines()
        days = [(2000, 1, d) for d in range(1, 32)] + [(2000, 2, d) for d in range(1, 10)]
        for sim in days:
            write_case(work, z, "current", sbs=sim)
            assert rout.main(["../parameters/configuration.txt"]) == 0
    finally:
        os.chdir(cwd)
    got = open(os.path.join(work, "output", "STEPBYSTEP", "OUTLT.day")).read().splitlines()
    got_mm = open(os.path.join(work, "output", "STEPBYSTEP", "OUTLT.day_mm")).read().splitlines()
    assert got == whole[:len(days)] and got_mm == whole_mm[:len(days)]
    # the carried files: this step's and the next step's storage (the one before is deleted, reservoir.f:1660-1668)
    left = sorted(os.listdir(os.path.join(work, "output", "stepbystep")))
    assert left == ["RESERVOIR_VOL_STATION1_STEP%d.txt" % len(days), "RESERVOIR_VOL_STATION1_STEP%d.txt" % (len(days) + 1)]
    # a step whose predecessor was not simulated is refused, as the reference does (label 9004)
    os.chdir(model_dir)
    try:
        write_case(work, z, "current", sbs=(2000, 3, 15))
        with pytest.raises(FileNotFoundError):
            rout.main(["../parameters/configuration.txt"])
    finally:
        os.chdir(cwd)


def _build_c_rout(tmp_path):
    import subprocess
    from common import ROOT
    from vicres_b200 import build
    build.build()
    exe = str(tmp_path / "rout_gpu")
    subprocess.run(["gcc", "-O2", "-Wall", "-Wno-format-truncation", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "host", "rout_gpu.c"), "-L", os.path.join(ROOT, "vicres_b200"), "-lvicres", "-lm",
                    "-Wl,-rpath," + os.path.join(ROOT, "vicres_b200")], check=True)
    return exe


def test_c_rout_host_compiles_against_the_headers(tmp_path):
    """CPU tier: host/rout_gpu.c is plain C over include/vicres_route.h; without a GPU it parses the case and stops at the
    first device call with the library's message (no CPU path)."""
    import subprocess
    import torch
    exe = _build_c_rout(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    z = np.load(os.path.join(GOLDEN, "route_rand12x10.npz"))
    write_case(str(tmp_path), z, "current")
    r = subprocess.run([exe, "../parameters/configuration.txt"], capture_output=True, text=True, cwd=os.path.join(str(tmp_path), "model"))
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "vr_route_create" in r.stderr, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,layout,legacy", [("route_rand20x16", "current", True), ("route_rand20x16", "binary", False),
                                                ("route_toy", "current", True)])
def test_c_rout_host_writes_the_same_files_as_the_python_command_line(name, layout, legacy, tmp_path):
    """host/rout_gpu.c (plain C: configuration, grids, stations, UH.all, flux files, resN.txt, writers) against
    python -m vicres_b200.rout on the same case: every file of the output directory, value for value (the two hosts do
    the same fp32 arithmetic for FACTOR and the summaries; the routing itself is the same library calls)."""
    import subprocess, shutil
    exe = _build_c_rout(tmp_path)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    work = str(tmp_path)
    write_case(work, z, layout)
    model_dir = os.path.join(work, "model")
    cwd = os.getcwd()
    os.chdir(model_dir)
    try:
        assert rout.main(["../parameters/configuration.txt"] + (["--legacy"] if legacy else [])) == 0
    finally:
        os.chdir(cwd)
    out = os.path.join(work, "output")
    shutil.move(out, os.path.join(work, "output_py"))
    os.makedirs(out)
    r = subprocess.run([exe, "../parameters/configuration.txt"] + (["--legacy"] if legacy else []), capture_output=True, text=True,
                       cwd=model_dir)
    assert r.returncode == 0, r.stderr[-2000:]
    names = sorted(n for n in os.listdir(os.path.join(work, "output_py")) if os.path.isfile(os.path.join(work, "output_py", n)))
    assert sorted(n for n in os.listdir(out) if os.path.isfile(os.path.join(out, n))) == names and len(names) >= 7
    same = 0
    for n in names:
        ta, tb = open(os.path.join(out, n)).read(), open(os.path.join(work, "output_py", n)).read()
        if ta == tb:
            same += 1
            continue
        skip = 1 if n.startswith("reservoir_") else 0
        a, b = np.loadtxt(os.path.join(out, n), skiprows=skip, ndmin=2), np.loadtxt(os.path.join(work, "output_py", n), skiprows=skip, ndmin=2)
        assert a.shape == b.shape, n
        # FACTOR holds |sin(lat - size/2) - sin(lat + size/2)| in fp32 (reservoir.f:357-364): a cancellation that turns the one
        # ulp between numpy's float32 sin and libm's sinf into ~5e-6 relative; it scales the flows of the --legacy variant
        # (FACTOR * 0.028) and the mm files.  Nothing else may differ.
        # (releases are storage differences divided by 86.4: an fp32 ulp of a storage is up to 2e-4 of the release column's range)
        bad = np.argwhere(~(np.abs(a - b) <= 2e-5 * np.abs(b) + 2e-4 * np.abs(b).max(axis=0, keepdims=True) * (1 if legacy else 0)))
        assert len(bad) == 0, "%s: %d values differ, first at %s: %r vs %r" % (n, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])])
    print("ROUT_GPU (C host) %s %s: %d of %d files byte-identical to the Python command line's" % (name, layout, same, len(names)))
    assert same >= 1


@pytest.mark.gpu
def test_c_rout_host_step_by_step(tmp_path):
    """the C host in STEPBYSTEP mode: 12 consecutive one-day runs append the first 12 lines of its all-steps run"""
    import subprocess
    exe = _build_c_rout(tmp_path)
    z = np.load(os.path.join(GOLDEN, "route_rand20x16.npz"))
    work = str(tmp_path)
    model_dir = os.path.join(work, "model")
    write_case(work, z, "current")
    assert subprocess.run([exe, "../parameters/configuration.txt"], cwd=model_dir).returncode == 0
    whole = open(os.path.join(work, "output", "OUTLT.day")).read().splitlines()
    for d in range(1, 13):
        write_case(work, z, "current", sbs=(2000, 1, d))
        assert subprocess.run([exe, "../parameters/configuration.txt"], cwd=model_dir).returncode == 0
    got = open(os.path.join(work, "output", "STEPBYSTEP", "OUTLT.day")).read().splitlines()
    assert got == whole[:12]
