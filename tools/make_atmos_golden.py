"""Generate tests/golden/atmos_*.npz by RUNNING THE REFERENCE's own initialize_atmos.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tools/make_atmos_golden.py [fixture names]

Each case is a set of reference-format files written by vicres_b200/synth.py (or the reference's bundled
Mekong cell); oracle/_ref/vic_ref_driver (the unmodified reference objects behind oracle/ref_driver.c) runs
initialize_atmos on every cell and dumps atmos[] in fp64.  A fixture keeps the raw forcing columns as the
reference read them, the site values of soil_con, the options, and the reference's nine atmos fields.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refdump  # noqa: E402
from vicres_b200 import synth  # noqa: E402
import make_golden  # noqa: E402

DRIVER = make_golden.DRIVER

# name: (synth kwargs, option overrides for vr_atmos_options)
CASES = {
    "atmos_daily": (dict(ncells=5, ndays=800, startyear=1983, lat_range=(-55, 64), seed=301), {}),
    "atmos_daily_offset": (dict(ncells=6, ndays=450, startyear=1984, lat_range=(-40, 60), off_gmt=0, seed=302), {}),
    "atmos_daily_cold_polar": (dict(ncells=4, ndays=420, startyear=1987, lat_range=(62, 82), cold=8.0, off_gmt=2, seed=303,
                                    extra_global="MTCLIM_SWE_CORR TRUE\nLW_CLOUD LW_CLOUD_BRAS\nLW_TYPE LW_TVA"),
                               dict(mtclim_swe_corr=True, lw_cloud="LW_CLOUD_BRAS", lw_type="LW_TVA")),
    "atmos_sub3h": (dict(ncells=4, ndays=200, dt=3, startyear=1988, lat_range=(-50, 65), cold=6.0, seed=304), {}),
    "atmos_sub6h_offset": (dict(ncells=4, ndays=120, dt=6, startyear=1990, lat_range=(20, 60), off_gmt=-3, seed=305), {}),
    "atmos_daily_to_3h": (dict(ncells=3, ndays=150, dt=3, force_daily=True, startyear=1991, lat_range=(-30, 50), off_gmt=5,
                               seed=306), {}),
    "atmos_converge_noplapse": (dict(ncells=3, ndays=380, startyear=1992, seed=307,
                                     extra_global="VP_ITER VP_ITER_CONVERGE\nPLAPSE FALSE\nLW_TYPE LW_BRUTSAERT"),
                                dict(vp_iter="VP_ITER_CONVERGE", plapse=False, lw_type="LW_BRUTSAERT")),
    # the reference's VP_ITER parser sets VP_INTERP to 1 for every value but VP_ITER_CONVERGE
    # (get_global_param.c:405-412), so "VP_INTERP FALSE" must come after VP_ITER to stay FALSE
    "atmos_annual_nointerp": (dict(ncells=3, ndays=300, startyear=1993, seed=308,
                                   extra_global="VP_ITER VP_ITER_ANNUAL\nVP_INTERP FALSE\nLW_TYPE LW_IDSO\nSW_PREC_THRESH 0.1"),
                              dict(vp_iter="VP_ITER_ANNUAL", vp_interp=False, lw_type="LW_IDSO", sw_prec_thresh=0.1)),
    # Not a fixture: STARTHOUR != 0 with sub-daily forcing.  The reference then indexes its forcing array up to
    # (Ndays_local*24 - starthour + hour_offset_int)/FORCE_DT - fstepspday, past the nrecs records it allocated
    # (initialize_atmos.c:511-513), so its last local day -- and through the precipitation totals all of MTCLIM's humidity --
    # depends on whatever follows the array in memory.  The library reads zeros there (what calloc gives inside the array).
    # observed columns: MTCLIM in its insw / invp modes, supplied longwave
    "atmos_daily_obs_sw_lw_vp": (dict(ncells=3, ndays=200, startyear=1997, lat_range=(-30, 58), off_gmt=1, cold=3.0, seed=312,
                                      extra_cols=("SHORTWAVE", "LONGWAVE", "VP")), {}),
    "atmos_sub3h_obs_sw_lw_vp": (dict(ncells=3, ndays=90, dt=3, startyear=1998, lat_range=(30, 60), cold=4.0, seed=313,
                                      extra_cols=("SHORTWAVE", "LONGWAVE", "VP")), {}),
    "atmos_daily_obs_sw": (dict(ncells=2, ndays=150, startyear=1999, lat_range=(10, 50), seed=314, extra_cols=("SHORTWAVE",)), {}),
    "atmos_sub3h_obs_vp": (dict(ncells=2, ndays=60, dt=3, startyear=1999, lat_range=(35, 55), off_gmt=-1, seed=315,
                                extra_cols=("VP",)), {}),
    "atmos_sub3h_obs_all": (dict(ncells=2, ndays=60, dt=3, startyear=2000, lat_range=(30, 55), seed=316,
                                 extra_cols=("SHORTWAVE", "LONGWAVE", "VP", "PRESSURE", "DENSITY")), {}),
    "atmos_daily_obs_pressure": (dict(ncells=2, ndays=100, startyear=2001, lat_range=(0, 45), off_gmt=3, seed=317,
                                      extra_cols=("PRESSURE",), extra_global="PLAPSE FALSE"), dict(plapse=False)),
    "atmos_daily_obs_density": (dict(ncells=2, ndays=100, startyear=2001, lat_range=(0, 45), seed=318, extra_cols=("DENSITY",)), {}),
    "atmos_daily_alma": (dict(ncells=2, ndays=120, startyear=2002, lat_range=(5, 50), off_gmt=-4, seed=319, alma=True,
                              extra_cols=("VP", "PRESSURE")), dict(alma_input=True)),
    "atmos_sub3h_alma": (dict(ncells=2, ndays=50, dt=3, startyear=2003, lat_range=(35, 60), seed=320, alma=True,
                              extra_cols=("SHORTWAVE", "LONGWAVE")), dict(alma_input=True)),
    "atmos_daily_no_wind": (dict(ncells=3, ndays=120, startyear=1996, lat_range=(-35, 55), off_gmt=-2, no_wind=True, seed=310),
                            dict(has_wind=False)),
    # SNOW_STEP < TIME_STEP (daily runs): every record carries NF windows + its own values, and the snow flags
    "atmos_daily_snow3": (dict(ncells=4, ndays=400, startyear=1985, lat_range=(30, 66), cold=7.0, nbands=3, snow_step=3, off_gmt=1,
                               seed=321), dict(snow_step=3)),
    "atmos_daily_snow1_lowwind": (dict(ncells=3, ndays=120, startyear=1994, lat_range=(-45, 55), cold=4.0, snow_step=1, wind_zero_frac=0.3,
                                       seed=322, extra_global="VP_INTERP FALSE\nMAX_SNOW_TEMP 1.5"),
                                  dict(snow_step=1, vp_interp=False, max_snow_temp=1.5)),
    "atmos_daily_snow6_obs": (dict(ncells=2, ndays=150, startyear=1996, lat_range=(35, 60), cold=5.0, nbands=2, snow_step=6, seed=323,
                                   extra_cols=("LONGWAVE", "VP", "PRESSURE")), dict(snow_step=6)),
    "atmos_daily_snow12_nowind": (dict(ncells=2, ndays=90, startyear=1997, lat_range=(40, 60), cold=6.0, snow_step=12, no_wind=True,
                                       off_gmt=-5, seed=324, extra_cols=("DENSITY",)), dict(snow_step=12, has_wind=False)),
    # humidity columns and the split precipitation / wind columns
    "atmos_daily_qair_pressure": (dict(ncells=3, ndays=150, startyear=2004, lat_range=(5, 55), off_gmt=2, seed=325,
                                       extra_cols=("QAIR", "PRESSURE")), {}),
    "atmos_daily_qair": (dict(ncells=3, ndays=150, startyear=2005, lat_range=(-40, 55), off_gmt=-3, cold=3.0, seed=326,
                              extra_cols=("QAIR",)), {}),
    "atmos_daily_rh": (dict(ncells=3, ndays=200, startyear=2006, lat_range=(10, 60), seed=327, extra_cols=("REL_HUMID",),
                            extra_global="VP_INTERP FALSE"), dict(vp_interp=False)),
    "atmos_daily_rh_to_3h": (dict(ncells=2, ndays=80, dt=3, force_daily=True, startyear=2007, lat_range=(20, 50), off_gmt=4, seed=328,
                                  extra_cols=("REL_HUMID",)), {}),
    "atmos_daily_rh_snow6": (dict(ncells=2, ndays=100, startyear=2008, lat_range=(35, 60), cold=5.0, snow_step=6, nbands=2, seed=329,
                                  extra_cols=("REL_HUMID",)), dict(snow_step=6)),
    "atmos_sub3h_rh": (dict(ncells=2, ndays=60, dt=3, startyear=2009, lat_range=(30, 55), off_gmt=1, seed=330,
                            extra_cols=("REL_HUMID",)), {}),
    "atmos_sub3h_qair_pressure_alma": (dict(ncells=2, ndays=50, dt=3, startyear=2010, lat_range=(35, 60), seed=331, alma=True,
                                            extra_cols=("QAIR", "PRESSURE")), dict(alma_input=True)),
    "atmos_daily_split_alma": (dict(ncells=2, ndays=120, startyear=2011, lat_range=(25, 60), cold=2.0, off_gmt=-2, seed=332, alma=True,
                                    extra_cols=("RAINF", "SNOWF", "WIND_E", "WIND_N")), dict(alma_input=True)),
    "atmos_sub3h_split": (dict(ncells=2, ndays=40, dt=3, startyear=2012, lat_range=(30, 60), cold=3.0, seed=333,
                               extra_cols=("RAINF", "SNOWF", "WIND_E", "WIND_N")), {}),
    "atmos_short_record": (dict(ncells=3, ndays=25, startyear=1995, seed=309,
                                extra_global="VP_ITER VP_ITER_NONE\nLW_TYPE LW_SATTERLUND"),
                           dict(vp_iter="VP_ITER_NONE", lw_type="LW_SATTERLUND")),
}


def read_raw(gfile, lat, lng, ncol):
    """The forcing columns as read_atmos_data parses them (ASCII), [ncol][nrec_f][C]."""
    g = dict(ln.split(None, 1) for ln in open(gfile) if ln.split() and ln.split()[0] == "FORCING1")
    pre = g["FORCING1"].strip()
    cols = []
    for la, lo in zip(lat, lng):
        a = np.loadtxt("%s%.4f_%.4f" % (pre, la, lo), ndmin=2)
        cols.append(a[:, :ncol])
    return np.ascontiguousarray(np.stack(cols, axis=2).transpose(1, 0, 2))


def save(name, gfile, dump, kw, ov, start=None):
    d = refdump.read_dump(dump)
    C = int(d["ncells"][0])
    sc = np.stack([d["c%d/soil_scalars" % c] for c in range(C)])
    site = {"lat": sc[:, 0], "lng": sc[:, 1], "elevation": sc[:, 2], "annual_prec": sc[:, 12], "time_zone_lng": sc[:, 17]}
    NF = int(d["NF"][0])
    if NF == 1:
        atm = np.stack([d["c%d/atmos" % c][:, :, 0] for c in range(C)], axis=2).transpose(1, 0, 2)   # [9, nrec, C]
    else:
        atm = np.stack([d["c%d/atmos" % c] for c in range(C)], axis=3).transpose(1, 0, 2, 3)         # [9, nrec, NF + 1, C]
    dt = int(d["dt"][0])
    daily = dt == 24 or kw.get("force_daily", False)
    extra = list(kw.get("extra_cols", ()))
    ncol = (4 if daily else 3) - (1 if kw.get("no_wind") else 0) + len(extra)
    raw = read_raw(gfile, site["lat"], site["lng"], ncol)
    nrecs = int(d["nrecs"][0])
    nrf = nrecs * dt // (24 if daily else dt)
    dmy = d["dmy"]
    skip = 0 if daily else int(dmy[0, 3]) // dt            # forcing records before the model's first hour (FORCEHOUR 0)
    arrays = {"raw": raw[:, skip:skip + nrf], "atmos": np.ascontiguousarray(atm),
              "meta": np.array([dt, nrecs, dmy[0, 0], dmy[0, 1], dmy[0, 2], dmy[0, 3], 0 if daily else 1, 24 if daily else dt],
                               dtype=np.int64),
              "ov_keys": np.array(list(ov.keys()) or [""]), "ov_vals": np.array([str(v) for v in ov.values()] or [""]),
              "extra_cols": np.array(extra or [""])}
    if NF > 1:
        arrays["snowflag"] = np.stack([d["c%d/snowflag" % c][:, NF] for c in range(C)], axis=1).astype(np.int32)   # [nrec, C], [NR]
        site["min_Tfactor"] = np.array([d["c%d/Tfactor" % c].min() for c in range(C)])
    for k, v in site.items():
        arrays["site_" + k] = v
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **arrays)
    print("%-26s cells %d recs %d  %.1f kB" % (name, C, nrecs, os.path.getsize(out) / 1e3))


def main():
    if not os.path.exists(DRIVER):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    with tempfile.TemporaryDirectory() as tmp:
        only = sys.argv[1:]                         # optional: the names to (re)generate
        for name, (kw, ov) in CASES.items():
            if only and name not in only:
                continue
            work = os.path.join(tmp, name)
            g = synth.make_case(work, max_veg=1, **kw)
            dump = os.path.join(work, "dump.bin")
            make_golden.run_driver(g, dump)
            save(name, g, dump, kw, ov)
        if only and "atmos_bundled_basin_cell" not in only:
            return
        work = os.path.join(tmp, "bundled")
        os.makedirs(work)
        g = make_golden.bundled_case(work)
        dump = os.path.join(work, "dump.bin")
        make_golden.run_driver(g, dump)
        save("atmos_bundled_basin_cell", g, dump, {}, {})


if __name__ == "__main__":
    main()
