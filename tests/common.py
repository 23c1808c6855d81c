"""Shared helpers for the test tiers: fixture loading, error metrics, seeded synthetic inputs."""
import glob
import os
import numpy as np
from vicres_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))

# Parity tolerance of BASELINE.json's north_star: fp64 fluxes and states within 1e-9 relative per
# cell-timestep.  Values are compared as |a-b| <= RTOL * max(|a|,|b|,ABS_FLOOR); the floor stops
# quantities that are ~0 by cancellation (closure error ~1e-13 mm, fluxes that are exactly 0 on one
# side and 1e-17 on the other) from dominating a relative measure.
RTOL = 1e-9
ABS_FLOOR = 1e-3


def relerr(a, b, floor=ABS_FLOOR):
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    nlayer, nbands, dt = [int(x) for x in z["meta"]]
    mst, mrt, wind_h = [float(x) for x in z["meta_f"]]
    names = [str(x) for x in z["out_names"]]
    lib = {k[4:]: z[k] for k in z.files if k.startswith("lib_")}
    cells = {k[6:]: z[k] for k in z.files if k.startswith("cells_")}
    opt = abi.default_options(nlayer, nbands, dt, out=names, max_snow_temp=mst, min_rain_temp=mrt, wind_h=wind_h)
    return {"opt": opt, "lib": lib, "cells": cells, "month": z["month"], "doy": z["doy"],
            "forcing": z["forcing"], "tair0": z["tair0"], "names": names, "ref": z["ref_out"],
            "nlayer": nlayer, "nbands": nbands, "dt": dt}


def check_outputs(out, ref, names, rtol=RTOL, what=""):
    worst = []
    for i, n in enumerate(names):
        e = relerr(out[i], ref[i])
        if not np.all(np.isfinite(out[i])):
            raise AssertionError("%s %s: non-finite output" % (what, n))
        if e.max() > rtol:
            k = np.unravel_index(np.argmax(e), e.shape)
            worst.append("%s: rel %.3e at rec %d cell %d (%.15g vs %.15g), %d of %d beyond tol" %
                         (n, e.max(), k[0], k[1], out[i][k], ref[i][k], int((e > rtol).sum()), e.size))
    assert not worst, what + " parity failures:\n" + "\n".join(worst)
