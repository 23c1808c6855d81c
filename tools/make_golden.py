"""Generate the committed golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tools/make_golden.py

For each case it writes reference-format input files (vicres_b200/synth.py, or the reference's own
bundled basin files), runs oracle/_ref/vic_ref_driver (the unmodified reference objects behind a
dumping main(), oracle/ref_driver.c) and stores, compressed, (1) every number the step consumed
(derived soil/veg/band parameters, disaggregated atmos[] records, calendar) in the layout of
include/vicres.h and (2) the reference's per-record out_data[] values for the VR_OUT_* variables.
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refdump  # noqa: E402
from vicres_b200 import abi, synth  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver")
REF = "/root/reference/runoff"

CASES = {
    # name: synth kwargs
    "daily_snow_bands": dict(ncells=6, nlayer=3, nbands=3, ndays=500, cold=14.0, lat_range=(30, 62),
                             overstory_frac=0.5, seed=101),
    "daily_warm": dict(ncells=8, nlayer=3, nbands=1, ndays=400, cold=0.0, lat_range=(-40, 40), seed=102),
    "sub3h_snow": dict(ncells=4, nlayer=3, nbands=2, ndays=100, dt=3, cold=16.0, lat_range=(35, 65),
                       overstory_frac=0.5, seed=103),
    "zero_wind_l2": dict(ncells=4, nlayer=2, nbands=1, ndays=300, cold=10.0, wind_zero_frac=0.1, seed=104),
    # SNOW_STEP < TIME_STEP: the snow model sub-stepped (NF = 8 and NF = 24 passes per daily record; the reference
    # allows it only with a daily model step, get_global_param.c:771-774)
    "daily_snowstep3": dict(ncells=6, nlayer=3, nbands=2, ndays=400, cold=12.0, lat_range=(30, 62), overstory_frac=0.5,
                            snow_step=3, seed=105),
    "daily_snowstep1": dict(ncells=4, nlayer=3, nbands=1, ndays=200, cold=10.0, lat_range=(35, 60), overstory_frac=0.5,
                            snow_step=1, seed=106),
}


def run_driver(gfile, dump):
    r = subprocess.run([DRIVER, "-g", gfile, "-d", dump], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-2000:])


def bundled_case(work):
    """The reference's bundled Mekong cell: soil.txt line 14208 is the only
    cell whose forcing file is shipped; NLAYER 2, 1 veg tile, daily; first 1500 days kept."""
    with open(os.path.join(REF, "parameters", "soil.txt")) as f:
        line = f.readlines()[14207]
    with open(os.path.join(work, "soil1.txt"), "w") as f:
        f.write(line)
    os.makedirs(os.path.join(work, "out"), exist_ok=True)
    g = []
    with open(os.path.join(REF, "parameters", "globalparam.txt")) as f:
        for ln in f:
            key = ln.split()[0] if ln.split() else ""
            if key == "SOIL":
                ln = "SOIL %s/soil1.txt\n" % work
            elif key == "RESULT_DIR":
                ln = "RESULT_DIR %s/out/\n" % work
            elif key == "FORCING1":
                ln = "FORCING1 %s/input/data_\n" % REF
            elif key == "VEGLIB":
                ln = "VEGLIB %s/parameters/vegetlibrary.txt\n" % REF
            elif key == "VEGPARAM":
                ln = "VEGPARAM %s/parameters/land.txt\n" % REF
            elif key == "ENDYEAR":
                ln = "ENDYEAR 1985\n"
            elif key == "ENDMONTH":
                ln = "ENDMONTH 2\n"
            elif key == "ENDDAY":
                ln = "ENDDAY 8\n"
            g.append(ln)
    gpath = os.path.join(work, "global.txt")
    with open(gpath, "w") as f:
        f.writelines(g)
    return gpath


def save_fixture(name, dump):
    d = refdump.read_dump(dump)
    cs = refdump.case_from_dump(d)
    names = [n for n in abi.OUT_NAMES if not (cs["nlayer"] == 2 and n == "SOIL_LIQ2")]
    ref = refdump.ref_outputs(d, names)
    arrays = {"out_names": np.array(names), "ref_out": ref, "forcing": cs["forcing"], "month": cs["month"],
              "doy": cs["doy"], "tair0": cs["tair0"],
              "meta": np.array([cs["nlayer"], cs["nbands"], cs["dt"], cs["snow_step"]], dtype=np.int64),
              "meta_f": np.array([cs["max_snow_temp"], cs["min_rain_temp"], cs["wind_h"]])}
    if cs["snowflag"] is not None:
        arrays["snowflag"] = cs["snowflag"]
    for k, v in cs["lib"].items():
        arrays["lib_" + k] = np.asarray(v)
    for k, v in cs["cells"].items():
        arrays["cells_" + k] = np.asarray(v)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **arrays)
    print("%-20s cells %d recs %d  %.1f kB  maxSWE %.1f" %
          (name, cs["ncell"], cs["nrec"], os.path.getsize(out) / 1e3, ref[names.index("SWE")].max()))


def main():
    if not os.path.exists(DRIVER):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, kw in CASES.items():
            if only and name not in only:
                continue
            work = os.path.join(tmp, name)
            g = synth.make_case(work, **kw)
            dump = os.path.join(work, "dump.bin")
            run_driver(g, dump)
            save_fixture(name, dump)
        if only and "bundled_basin_cell" not in only:
            return
        work = os.path.join(tmp, "bundled")
        os.makedirs(work)
        g = bundled_case(work)
        dump = os.path.join(work, "dump.bin")
        run_driver(g, dump)
        save_fixture("bundled_basin_cell", dump)


if __name__ == "__main__":
    main()
