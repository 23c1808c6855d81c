"""Golden fixtures for the routing pass, made by RUNNING THE REFERENCE'S SHIPPED PREBUILT BINARY
(routing/model/rout, x86-64 ELF; there is no Fortran compiler here).

    python tools/make_route_golden.py

Cases: the bundled 5x5 toy basin, and seeded random D8 trees with reservoirs.  For every case the
script writes the reference's input files (flowdirection, reservoirlocation, stations, UH.all,
resN.txt, fluxes_<lat>_<lon>, configuration in the layout the prebuilt binary expects), runs the
binary with scipy's bundled libgfortran, and stores the inputs plus the binary's .uh_s files (UH_S
per node and cell) and NF1.txt (naturalised inflow of every node = MAKE_CONVOLUTION) in
tests/golden/route_<case>.npz.  The binary predates two details of today's reservoir.f: it scales
runoff by FACTOR*0.028 (area based) where the source now hard-codes 0.18308 (reservoir.f:479-480);
the fixtures record which scale applies.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vicres_b200 import route_abi  # noqa: E402

REF = "/root/reference/routing"
SCIPY_LIBS = os.path.join(os.path.dirname(np.__file__), "..", "scipy.libs")

CODE = {(0, 1): 1, (-1, 1): 2, (-1, 0): 3, (-1, -1): 4, (0, -1): 5, (1, -1): 6, (1, 0): 7, (1, 1): 8}  # (dI, dJ)


def gfortran_env(tmp):
    d = os.path.join(tmp, "gf")
    os.makedirs(d, exist_ok=True)
    for pat, name in (("libgfortran-040039e1*", "libgfortran.so.5"), ("libquadmath-96973f99*", None)):
        for f in glob.glob(os.path.join(SCIPY_LIBS, pat)):
            dst = os.path.join(d, name or os.path.basename(f))
            if not os.path.exists(dst):
                os.symlink(os.path.abspath(f), dst)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = d
    return env


def random_tree(nrow, ncol, rng, fill=0.8):
    """D8 spanning tree grown from an outlet on the left edge; returns dir codes (file order), outlet (r, c)"""
    d = np.zeros((nrow, ncol), dtype=np.int32)
    out = (int(rng.integers(1, nrow - 1)), 0)
    seen = {out}
    frontier = [out]
    target = int(fill * nrow * ncol)
    while frontier and len(seen) < target:
        r, c = frontier.pop(int(rng.integers(0, len(frontier))))
        nb = [(r + dr, c + dc) for dr in (-1, 0, 1) for dc in (-1, 0, 1) if (dr or dc)]
        rng.shuffle(nb)
        for (rr, cc) in nb[:int(rng.integers(1, 4))]:
            if 0 <= rr < nrow and 0 <= cc < ncol and (rr, cc) not in seen:
                seen.add((rr, cc))
                dI, dJ = -(c - cc), -(r - rr)      # from (rr,cc) towards (r,c) in Fortran (I,J) = (ncol-c, nrow-r)
                d[rr, cc] = CODE[(dI, dJ)]
                frontier.append((rr, cc))
        if (r, c) not in frontier and rng.uniform() < 0.5:
            frontier.append((r, c))
    return d, out


def res_text(rng, rid):
    """A resN.txt in the layout the prebuilt binary reads (one HYDROPOWER value; today's reservoir.f reads two).
    Only what the binary provably shares with today's source is exercised: rules 1 and 2 and reservoirs that
    are commissioned after the simulated period (tests/test_route_oracle.py)."""
    hmax = float(np.round(rng.uniform(110, 160), 1))
    hmin = float(np.round(hmax - rng.uniform(15, 40), 1))
    vlive = float(np.round(rng.uniform(2000, 60000), 0))
    vdead = float(np.round(vlive * rng.uniform(0.1, 0.3), 0))
    q = float(np.round(rng.uniform(3, 40), 1))
    kind = rid % 4
    year = 2003 if kind == 3 else 1999
    vini = float(np.round(vdead + rng.uniform(0.2, 0.9) * vlive, 0))
    if kind == 2:
        rule, body = 2, " ".join("%.1f" % x for x in np.sort(rng.uniform(hmin, hmax, 12)))
    else:
        t1, t2 = sorted(int(x) for x in rng.choice(np.arange(20, 340), 2, replace=False))
        if rid % 2:
            t1, t2 = t2, t1
        rule, body = 1, "%.1f %.1f %d %d" % (hmax - 1.0, hmin + 2.0, t1, t2)
    return ("Hmax(M)\tHmin(M)\tSlive(1000M3)\tSdead(1000M3)\tHturbine(M)\tQdesign (M3/s)\tYear\tSinitial(1000M3)\tName\n"
            "%g\t%g\t%g\t%g\t%g\t%g\t%d\t%g\tR%d\nSEEPAGE\tINFILTRATION\n%.4f\t%.4f\nIRRIGATION\n0\nDUMMY\nHYDROPOWER\n1\n"
            "ENVIRONMENTAL FLOW\n%.2f\nOPERATION STRATEGY\n%d\n%s\nBATHYMETRY PARAMETERS\n0\n135.186458\t0.033764077\t-0.00041473\t1.11114436\n"
            % (hmax, hmin, vlive, vdead, float(np.round(0.6 * (hmax - hmin) + 20, 1)), q, year, vini, rid,
               rng.uniform(0, 0.05) if rid % 3 == 0 else 0.0, rng.uniform(0, 0.03) if rid % 3 == 0 else 0.0,
               rng.choice([0.05, 0.1]), rule, body))


def write_case(work, d, reser, station_rc, hdr, velo, diff, xmask, box, flux, ndays, res_rng=None):
    nrow, ncol = d.shape
    par = os.path.join(work, "parameters")
    for sub in ("", "unit_hydrograph", "reservoirs/res_par", "naturalized_flow/search_catchment"):
        os.makedirs(os.path.join(par, sub), exist_ok=True)
    os.makedirs(os.path.join(work, "input"), exist_ok=True)
    os.makedirs(os.path.join(work, "output"), exist_ok=True)
    os.makedirs(os.path.join(work, "model"), exist_ok=True)
    shutil.copy(os.path.join(REF, "model", "rout"), os.path.join(work, "model", "rout"))
    os.chmod(os.path.join(work, "model", "rout"), 0o755)

    def esri(path, a, fmt, header=True):
        with open(path, "w") as f:
            if header:
                f.write("ncols %d\nnrows %d\nxllcorner %s\nyllcorner %s\ncellsize %s\nNODATA_value 0\n" %
                        (ncol, nrow, hdr["xllcorner"], hdr["yllcorner"], hdr["cellsize"]))
            for r in range(nrow):
                f.write(" ".join(fmt % x for x in a[r]) + "\n")
    esri(os.path.join(par, "flowdirection.txt"), d, "%d")
    esri(os.path.join(par, "reservoirlocation.txt"), reser, "%d", header=False)
    esri(os.path.join(par, "velocity.txt"), velo, "%.4f")
    esri(os.path.join(par, "diffusion.txt"), diff, "%.2f")
    esri(os.path.join(par, "xmask.txt"), xmask, "%.1f")
    r, c = station_rc
    with open(os.path.join(par, "stations.txt"), "w") as f:
        f.write("1 OUTLT %d %d -9999\nNONE\n" % (c + 1, nrow - r))
    with open(os.path.join(par, "unit_hydrograph", "UH.all"), "w") as f:
        for k in range(12):
            f.write("%d %.4f\n" % (k, box[k]))
    res_txt = open(os.path.join(REF, "parameters", "reservoirs", "res_par", "res1.txt")).read()
    for rid in sorted(set(int(x) for x in reser.ravel() if 0 < x < 9999)):
        with open(os.path.join(par, "reservoirs", "res_par", "res%d.txt" % rid), "w") as f:
            f.write(res_text(res_rng, rid) if res_rng is not None else
                    open(os.path.join(REF, "parameters", "reservoirs", "res_par", "res%d.txt" % min(rid, 2))).read())
    for (rr, cc), a in flux.items():
        lat, lon = route_abi.cell_latlon(hdr, nrow, ncol, rr, cc)
        with open(os.path.join(work, "input", "fluxes_%.4f_%.4f" % (lat, lon)), "w") as f:
            day = 0
            for mo, nd in zip(range(1, 6), (31, 29, 31, 30, 31)):
                for dd in range(1, nd + 1):
                    ro, ba, evv, trv, evs = a[day][:5]
                    rh, ta, wi = a[day][5:8] if a.shape[1] >= 8 else (49.0, 15.0, 3.0)
                    cols = [0.0, evv + trv + evs, ro, ba, 0.0, 10.0, 50.0, 55.0, -26.0, evv, trv, evs, 0.0, 0.0, 5.7, 15.0,
                            0.12, rh, 309.0, ta, wi]
                    f.write("2000\t%d\t%d\t" % (mo, dd) + "\t".join("%.4f" % x for x in cols) + "\n")
                    day += 1
    cfg = ["# Routing main file", "# NAME OF FLOW DIRECTION FILE", "../parameters/flowdirection.txt",
           "# NAME OF VELOCITY FILE", ".true.", "../parameters/velocity.txt", "# NAME OF DIFF FILE", ".true.",
           "../parameters/diffusion.txt", "# NAME OF XMASK FILE", ".true.", "../parameters/xmask.txt",
           "# NAME OF FRACTION FILE", ".false.", "1.0", "# NAME OF STATION FILE", "../parameters/stations.txt",
           "# PATH OF INPUT FILES AND PRECISION", "../input/fluxes_", "4", "# PATH OF OUTPUT FILES", "../output/",
           "# RESERVOIR LOCATION", "../parameters/reservoirlocation.txt", "# RESERVOIR PARAMETERS",
           "../parameters/reservoirs/res_par/", "# NATURALIZED FLOW", ".false. ", "../parameters/naturalized_flow/ ",
           "# YEAR AND MONTH OF MODEL OUTPUT TO ROUTE & ROUTED OUTPUT TO WRITE         ", "2000 1 2000 5 ", "2000 1 2000 5 ",
           "# NAME OF UNIT HYDROGRAPH FILE         ", "../parameters/unit_hydrograph/",
           "# Simulation Mode: True: Run Step-by-Step Mode; False: Run all Steps", ".false.",
           "# Start Day and Current Simulation Day (for STEPBYSTEP Version)", "2000 1 1 2000 5 31", "# Coupler Iteration", "1",
           "# Write Outputs", ".true."]
    with open(os.path.join(par, "config.txt"), "w") as f:
        f.write("\n".join(cfg) + "\n")
    return ndays


def run_binary(work, env):
    r = subprocess.run(["./rout", "../parameters/config.txt"], cwd=os.path.join(work, "model"), env=env,
                       capture_output=True, text=True, timeout=600)
    if not os.path.exists(os.path.join(work, "parameters", "naturalized_flow", "NF1.txt")):
        raise RuntimeError("rout failed:\n" + r.stdout[-3000:] + r.stderr[-1000:])
    return r.stdout


def collect(work, name, d, reser, station_rc, hdr, velo, diff, xmask, box, flux, ndays):
    nrow, ncol = d.shape
    nf = np.loadtxt(os.path.join(work, "parameters", "naturalized_flow", "NF1.txt"), ndmin=2).astype(np.float32)
    resdir = np.loadtxt(glob.glob(os.path.join(work, "parameters", "naturalized_flow", "search_catchment", "RES_DIR_*.txt"))[0], ndmin=2)
    uh = {}
    for f in glob.glob(os.path.join(work, "parameters", "unit_hydrograph", "*.uh_s")):
        uh[os.path.basename(f)[:-5]] = np.loadtxt(f, ndmin=2).astype(np.float32)
    arrays = {"dir": d, "reser": reser, "station_rc": np.array(station_rc), "velo": velo.astype(np.float32),
              "diff": diff.astype(np.float32), "xmask": xmask.astype(np.float32), "uh_box": np.asarray(box, dtype=np.float32),
              "hdr": np.array([float(hdr["xllcorner"]), float(hdr["yllcorner"]), float(hdr["cellsize"])]),
              "nf": nf, "res_order": resdir[:, 0].astype(np.int32), "ndays": np.array(ndays)}
    runoff = np.zeros((nrow * ncol, ndays), dtype=np.float32)
    base, evv, trv, evs, rh, tair, wind = (np.zeros_like(runoff) for _ in range(7))
    rh[:], tair[:], wind[:] = 49.0, 15.0, 3.0
    for (rr, cc), a in flux.items():
        k = rr * ncol + cc
        # what the binary reads back from the %.4f text
        a4 = np.array([[float("%.4f" % x) for x in row] for row in a], dtype=np.float32)
        runoff[k], base[k], evv[k], trv[k], evs[k] = a4.T[:5]
        if a4.shape[1] >= 8:
            rh[k], tair[k], wind[k] = a4.T[5:8]
    arrays.update(rh=rh, tair=tair, wind=wind)
    present = np.zeros(nrow * ncol, dtype=np.uint8)
    for (rr, cc) in flux:
        present[rr * ncol + cc] = 1
    arrays.update(runoff=runoff, baseflow=base, ev_vege=evv, tr_vege=trv, ev_soil=evs, present=present)
    for k, v in uh.items():
        arrays["uh_" + k] = v
    # reservoir pass: UH_R files, the resN.txt texts, the binary's reservoir_<id>_<station>.day and station flow
    for f in glob.glob(os.path.join(work, "parameters", "unit_hydrograph", "*.uh_r")):
        arrays["uhr_" + os.path.basename(f)[:-5]] = np.loadtxt(f).astype(np.float32)
    for rid in arrays["res_order"][:-1]:
        arrays["restxt_%d" % rid] = np.array(open(os.path.join(work, "parameters", "reservoirs", "res_par", "res%d.txt" % rid)).read())
        arrays["resday_%d" % rid] = np.loadtxt(os.path.join(work, "output", "reservoir_%d_OUTLT.day" % rid), skiprows=1)[:, 3:].astype(np.float32)
    arrays["station_flow"] = np.loadtxt(os.path.join(work, "output", "OUTLT.day"))[:, 3].astype(np.float32)
    # the summaries of WRITE_DATA / WRITE_MONTH (write_routines.f:64-70, 248-331): <station>.day_mm, .month, .month_mm,
    # .year, .year_mm as the binary wrote them (all columns, float64 of the printed text)
    for ext in ("day_mm", "month", "month_mm", "year", "year_mm"):
        f = os.path.join(work, "output", "OUTLT." + ext)
        if os.path.exists(f) and os.path.getsize(f) > 0:
            arrays["station_" + ext] = np.loadtxt(f, ndmin=2)
    out = os.path.join(ROOT, "tests", "golden", "route_%s.npz" % name)
    np.savez_compressed(out, **arrays)
    print("%-12s grid %dx%d nodes %d  uh files %s  %.0f kB" % (name, nrow, ncol, nf.shape[0], sorted(uh), os.path.getsize(out) / 1e3))


def toy_case(tmp, env):
    work = os.path.join(tmp, "toy")
    hdr0, d = route_abi.read_esri_int(os.path.join(REF, "parameters", "flowdirection.txt"))
    _, reser = route_abi.read_esri_int(os.path.join(REF, "parameters", "reservoirlocation.txt"), header=0)
    hdr = {k: hdr0[k] for k in ("xllcorner", "yllcorner", "cellsize")}
    nrow, ncol = d.shape
    ndays = 152
    flux = {}
    for r in range(nrow):
        for c in range(ncol):
            lat, lon = route_abi.cell_latlon(hdr, nrow, ncol, r, c)
            fn = os.path.join(REF, "input", "fluxes_%.4f_%.4f" % (lat, lon))
            if os.path.exists(fn):
                a = np.loadtxt(fn)[:ndays]
                flux[(r, c)] = np.stack([a[:, 5], a[:, 6], a[:, 12], a[:, 13], a[:, 14]], axis=1)
    box = [1.0] + [0.0] * 11
    ones = np.ones((nrow, ncol))
    write_case(work, d, reser, (nrow - 5, 0), hdr, 1.2 * ones, 800.0 * ones, 12500.0 * ones, box, flux, ndays)
    run_binary(work, env)
    collect(work, "toy", d, reser, (nrow - 5, 0), hdr, 1.2 * ones, 800.0 * ones, 12500.0 * ones, box, flux, ndays)


def random_case(tmp, env, name, nrow, ncol, seed, nres):
    rng = np.random.default_rng(seed)
    work = os.path.join(tmp, name)
    d, out = random_tree(nrow, ncol, rng)
    hdr = {"xllcorner": "100.0", "yllcorner": "10.0", "cellsize": "0.25"}
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    pick = rng.choice(len(active), size=min(nres, len(active)), replace=False)
    for i, k in enumerate(pick):
        reser[tuple(active[k])] = i + 1
    velo = rng.uniform(0.8, 2.5, (nrow, ncol))
    diff = rng.uniform(400.0, 1500.0, (nrow, ncol))
    xmask = rng.uniform(15000.0, 30000.0, (nrow, ncol))
    box = rng.dirichlet(np.ones(12) * 0.6)
    box = np.round(box, 4)
    ndays = 152
    flux = {}
    for (r, c) in [tuple(x) for x in active] + [out]:
        if rng.uniform() < 0.9:                  # some missing files -> zeros inserted by the reference
            ro = rng.gamma(0.5, 4.0, ndays) * (rng.uniform(size=ndays) < 0.4)
            ba = np.abs(rng.normal(1.0, 0.5, ndays))
            evv, trv, evs = rng.uniform(0, 1, ndays), rng.uniform(0, 2, ndays), rng.uniform(0, 1, ndays)
            trv[0] = -rng.uniform(5, 60)         # the first-day negative evaporation quirk
            flux[(r, c)] = np.stack([ro, ba, evv, trv, evs], axis=1)
    write_case(work, d, reser, out, hdr, velo, diff, xmask, box, flux, ndays, res_rng=rng)
    run_binary(work, env)
    collect(work, name, d, reser, out, hdr, np.round(velo, 4), np.round(diff, 2), np.round(xmask, 1), box, flux, ndays)


def evap_case(tmp, env, name="evap14x16", nrow=14, ncol=16, seed=31):
    """reservoir-surface cells (RESER == 9999) with varying RH / air temperature / wind: exercises the
    open-water evaporation of MAKE_CONVOLUTION (reservoir.f:410-475), incl. the 21.5-22 C table overrun"""
    rng = np.random.default_rng(seed)
    work = os.path.join(tmp, name)
    d, out = random_tree(nrow, ncol, rng)
    hdr = {"xllcorner": "100.0", "yllcorner": "10.0", "cellsize": "0.25"}
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    pick = rng.choice(len(active), size=10, replace=False)
    for i, k in enumerate(pick[:4]):
        reser[tuple(active[k])] = i + 1
    for k in pick[4:]:
        reser[tuple(active[k])] = 9999
    velo = rng.uniform(0.8, 2.5, (nrow, ncol))
    diff = rng.uniform(400.0, 1500.0, (nrow, ncol))
    xmask = rng.uniform(15000.0, 30000.0, (nrow, ncol))
    box = np.round(rng.dirichlet(np.ones(12) * 0.6), 4)
    ndays = 152
    flux = {}
    for (r, c) in [tuple(x) for x in active] + [out]:
        if rng.uniform() < 0.85:
            ro = rng.gamma(0.5, 4.0, ndays) * (rng.uniform(size=ndays) < 0.4)
            ba = np.abs(rng.normal(1.0, 0.5, ndays))
            flux[(r, c)] = np.stack([ro, ba, rng.uniform(0, 1, ndays), rng.uniform(0, 2, ndays), rng.uniform(0, 1, ndays),
                                     rng.uniform(20, 95, ndays), rng.uniform(-4, 27, ndays), rng.uniform(0, 8, ndays)], axis=1)
    write_case(work, d, reser, out, hdr, velo, diff, xmask, box, flux, ndays, res_rng=rng)
    run_binary(work, env)
    collect(work, name, d, reser, out, hdr, np.round(velo, 4), np.round(diff, 2), np.round(xmask, 1), box, flux, ndays)


def main():
    with tempfile.TemporaryDirectory() as tmp:
        env = gfortran_env(tmp)
        toy_case(tmp, env)
        random_case(tmp, env, "rand12x10", 10, 12, 7, 3)
        random_case(tmp, env, "rand20x16", 16, 20, 11, 4)
        evap_case(tmp, env)


if __name__ == "__main__":
    main()
