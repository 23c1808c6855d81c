"""Golden fixture for the WHOLE chain of the north-star path, made by running the reference's own two
programs back to back:

    parameter + forcing files -> vicNl (compiled in place, oracle/_ref/vic_ref_driver) -> fluxes_<lat>_<lng>
    -> routing/model/rout (the shipped prebuilt binary) -> OUTLT.day, reservoir_<id>_OUTLT.day

A seeded D8 tree on a 0.25-degree grid, one 2-layer VIC cell on the centre of every routed grid cell (the
classic flux-file reader of rout hard-codes the 2-layer column layout, reservoir.f:383-386), three reservoirs.
Stored in tests/golden/chain_case/: the parameter files (text), the routing grid and reservoir files, what
the step consumed (atmos records from the reference's initialize_atmos) and the binary's results.

    python tools/make_chain_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refdump  # noqa: E402
import make_route_golden as mrg  # noqa: E402
from vicres_b200 import route_abi, synth  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver")
DEST = os.path.join(ROOT, "tests", "golden", "chain_case")


def main():
    rng = np.random.default_rng(41)
    nrow, ncol, ndays = 9, 11, 152
    d, out = mrg.random_tree(nrow, ncol, rng)
    hdr = {"xllcorner": "100.0", "yllcorner": "40.0", "cellsize": "0.25"}
    active = [tuple(x) for x in np.argwhere(d > 0)] + [out]
    reser = np.zeros_like(d)
    for i, k in enumerate(rng.choice(len(active) - 1, size=3, replace=False)):
        reser[active[k]] = i + 1
    coords = [route_abi.cell_latlon(hdr, nrow, ncol, r, c) for (r, c) in active]
    with tempfile.TemporaryDirectory() as tmp:
        env = mrg.gfortran_env(tmp)
        vic = os.path.join(tmp, "vic")
        g = synth.make_case(vic, ncells=len(active), nlayer=2, nbands=1, ndays=ndays, startyear=2000, cold=6.0,
                            seed=43, coords=coords)
        dump = os.path.join(vic, "dump.bin")
        r = subprocess.run([DRIVER, "-g", g, "-d", dump], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-2000:])
        work = os.path.join(tmp, "rout")
        velo = rng.uniform(0.8, 2.5, (nrow, ncol)); diff = rng.uniform(400.0, 1500.0, (nrow, ncol))
        xmask = rng.uniform(15000.0, 30000.0, (nrow, ncol))
        box = np.round(rng.dirichlet(np.ones(12) * 0.6), 4)
        mrg.write_case(work, d, reser, out, hdr, velo, diff, xmask, box, {}, ndays, res_rng=rng)
        for fn in os.listdir(os.path.join(vic, "results")):           # the reference's own flux files feed rout
            if fn.startswith("fluxes_"):
                shutil.copy(os.path.join(vic, "results", fn), os.path.join(work, "input", fn))
        log = mrg.run_binary(work, env)
        assert "NOT FOUND" not in log, log[-2000:]
        shutil.rmtree(DEST, ignore_errors=True)
        os.makedirs(os.path.join(DEST, "res_par"))
        for fn in ["soil.txt", "veglib.txt", "veg.txt"]:
            shutil.copy(os.path.join(vic, fn), os.path.join(DEST, fn))
        dd = refdump.read_dump(dump)
        cs = refdump.case_from_dump(dd)
        arrays = {"forcing": cs["forcing"].astype(np.float64), "month": cs["month"], "doy": cs["doy"], "tair0": cs["tair0"],
                  "dmy": np.asarray(dd["dmy"]), "meta_f": np.array([cs["max_snow_temp"], cs["min_rain_temp"], cs["wind_h"]]),
                  "dir": d, "reser": reser, "station_rc": np.array(out), "velo": np.round(velo, 4).astype(np.float32),
                  "diff": np.round(diff, 2).astype(np.float32), "xmask": np.round(xmask, 1).astype(np.float32),
                  "uh_box": np.asarray(box, dtype=np.float32),
                  "hdr": np.array([float(hdr["xllcorner"]), float(hdr["yllcorner"]), float(hdr["cellsize"])]),
                  "station_flow": np.loadtxt(os.path.join(work, "output", "OUTLT.day"))[:, 3].astype(np.float32),
                  "nf": np.loadtxt(os.path.join(work, "parameters", "naturalized_flow", "NF1.txt"), ndmin=2).astype(np.float32)}
        ids = sorted(int(x) for x in reser.ravel() if 0 < x < 9999)
        for rid in ids:
            shutil.copy(os.path.join(work, "parameters", "reservoirs", "res_par", "res%d.txt" % rid),
                        os.path.join(DEST, "res_par", "res%d.txt" % rid))
            arrays["resday_%d" % rid] = np.loadtxt(os.path.join(work, "output", "reservoir_%d_OUTLT.day" % rid),
                                                   skiprows=1)[:, 3:].astype(np.float32)
        # one reference flux file kept for the byte comparison of the writer
        first = sorted(os.listdir(os.path.join(vic, "results")))[0]
        shutil.copy(os.path.join(vic, "results", first), os.path.join(DEST, first))
        np.savez_compressed(os.path.join(DEST, "chain.npz"), **arrays)
        print("cells", len(active), "reservoirs", ids, "station flow mean %.3f" % arrays["station_flow"].mean(),
              {k: round(os.path.getsize(os.path.join(DEST, k)) / 1e3) for k in os.listdir(DEST) if os.path.isfile(os.path.join(DEST, k))})


if __name__ == "__main__":
    main()
