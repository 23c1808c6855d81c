"""Golden fixture for the parameter-file readers and the flux writer (include/vicres_files.h).

Writes a small reference-format case (vicres_b200/synth.py) into tests/golden/files_case/, runs the
compiled reference on it (oracle/_ref/vic_ref_driver, which also writes the reference's own
fluxes_* files) and stores what the reference held after ITS readers plus its flux files:

    tests/golden/files_case/{soil.txt,veglib.txt,veg.txt,snowband.txt}   inputs (text)
    tests/golden/files_case/expected.npz                                  reference structs + forcing
    tests/golden/files_case/fluxes/fluxes_<lat>_<lng>                     reference output files

Run in the build container:  python tools/make_files_golden.py
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
import refdump  # noqa: E402
from vicres_b200 import synth  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver")
DEST = os.path.join(ROOT, "tests", "golden", "files_case")


def main():
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "case")
        g = synth.make_case(work, ncells=5, nlayer=3, nbands=3, ndays=90, cold=9.0, lat_range=(25, 60),
                            overstory_frac=0.5, seed=311)
        dump = os.path.join(work, "dump.bin")
        r = subprocess.run([DRIVER, "-g", g, "-d", dump], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-2000:])
        d = refdump.read_dump(dump)
        cs = refdump.case_from_dump(d)
        shutil.rmtree(DEST, ignore_errors=True)
        os.makedirs(os.path.join(DEST, "fluxes"))
        for fn in ["soil.txt", "veglib.txt", "veg.txt", "snowband.txt"]:
            shutil.copy(os.path.join(work, fn), os.path.join(DEST, fn))
        for fn in sorted(os.listdir(os.path.join(work, "results"))):
            if fn.startswith("fluxes_"):
                shutil.copy(os.path.join(work, "results", fn), os.path.join(DEST, "fluxes", fn))
        arrays = {"forcing": cs["forcing"], "month": cs["month"], "doy": cs["doy"], "tair0": cs["tair0"],
                  "dmy": np.asarray(d["dmy"]),
                  "meta": np.array([cs["nlayer"], cs["nbands"], cs["dt"]], dtype=np.int64),
                  "meta_f": np.array([cs["max_snow_temp"], cs["min_rain_temp"], cs["wind_h"]])}
        for k, v in cs["lib"].items():
            arrays["lib_" + k] = np.asarray(v)
        for k, v in cs["cells"].items():
            arrays["cells_" + k] = np.asarray(v)
        np.savez_compressed(os.path.join(DEST, "expected.npz"), **arrays)
        print(sorted(os.listdir(DEST)), sorted(os.listdir(os.path.join(DEST, "fluxes"))))


if __name__ == "__main__":
    main()
