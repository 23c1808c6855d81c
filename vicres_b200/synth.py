"""Seeded synthetic VIC-Res cases, written in the REFERENCE'S OWN FILE FORMATS.

One generator feeds both sides of every comparison: the reference build (oracle/_ref) reads the
files written here, and the B200 path reads the same files through vicres_b200.host parsers (or
the dump the reference driver makes from them), so both see identical numbers.

Formats follow the reference readers:
  global parameter file   runoff/model/get_global_param.c:210-727
  soil parameter file     runoff/model/read_soilparam.c:169-640   (53 columns for NLAYER=3)
  vegetation library      runoff/model/read_veglib.c:52-160
  vegetation parameters   runoff/model/read_vegparam.c:85-340
  snow band file          runoff/model/read_snowband.c:44-125
  forcing files           runoff/model/read_atmos_data.c:120-235  (ASCII, one row per record)

Parameter ranges are the ones SURVEY.md section 8(d) fixes for cfg2/cfg3.
"""
import os
import numpy as np

# A small synthetic vegetation library (class id, overstory, rarc, rmin, LAI[12], albedo[12],
# roughness[12], displacement[12], wind_h, RGL, rad_atten, wind_atten, trunk_ratio).
# Values are typical of the public UMD-class VIC libraries; they are NOT the reference's file.
_LAI_FOREST = [3.4, 3.4, 3.5, 3.7, 4.0, 4.4, 4.4, 4.3, 4.2, 3.7, 3.5, 3.4]
_LAI_DECID = [0.5, 0.5, 0.8, 1.6, 3.2, 4.6, 5.0, 4.8, 3.6, 1.8, 0.7, 0.5]
_LAI_SHRUB = [0.8, 0.8, 0.9, 1.1, 1.6, 2.2, 2.5, 2.4, 1.9, 1.3, 0.9, 0.8]
_LAI_GRASS = [0.4, 0.4, 0.5, 0.9, 1.6, 2.3, 2.6, 2.3, 1.6, 0.9, 0.5, 0.4]
_LAI_CROP = [0.1, 0.1, 0.2, 0.6, 1.8, 3.4, 4.2, 3.8, 2.0, 0.6, 0.2, 0.1]
VEGLIB = [
    # id ov  rarc  rmin  LAI          alb   rough  displ  wind_h RGL   rad_att wind_att trunk
    (1, 1, 60.0, 250.0, _LAI_FOREST, 0.12, 1.476, 8.04, 50.0, 30.0, 0.5, 0.5, 0.2),
    (2, 1, 60.0, 125.0, _LAI_DECID, 0.14, 1.230, 6.70, 50.0, 30.0, 0.5, 0.5, 0.2),
    (3, 1, 50.0, 150.0, _LAI_FOREST, 0.13, 0.984, 5.36, 40.0, 50.0, 0.5, 0.5, 0.2),
    (4, 0, 40.0, 135.0, _LAI_SHRUB, 0.18, 0.495, 1.00, 10.0, 75.0, 0.5, 0.5, 0.2),
    (5, 0, 25.0, 120.0, _LAI_GRASS, 0.20, 0.0738, 0.402, 10.0, 100.0, 0.5, 0.5, 0.2),
    (6, 0, 25.0, 120.0, _LAI_CROP, 0.19, 0.123, 0.67, 10.0, 100.0, 0.5, 0.5, 0.2),
]


def write_veglib(path):
    with open(path, "w") as f:
        f.write("#Class OvrStry Rarc Rmin LAI[12] ALB[12] ROU[12] DIS[12] WIND_H RGL rad_atten wind_atten trunk_ratio comment\n")
        for (cid, ov, rarc, rmin, lai, alb, rough, displ, wind_h, rgl, ra, wa, tr) in VEGLIB:
            cols = [str(cid), str(ov), "%.1f" % rarc, "%.1f" % rmin]
            cols += ["%.3f" % x for x in lai]
            cols += ["%.4f" % alb] * 12
            cols += ["%.4f" % rough] * 12
            cols += ["%.4f" % displ] * 12
            cols += ["%.1f" % wind_h, "%.1f" % rgl, "%.2f" % ra, "%.2f" % wa, "%.2f" % tr, "synthetic"]
            f.write(" ".join(cols) + "\n")


def _fmt(x, nd=6):
    return ("%." + str(nd) + "f") % x


def make_case(outdir, ncells=8, nlayer=3, nbands=1, ndays=365, dt=24, seed=20261019,
              startyear=1981, lat_range=(-40.0, 60.0), cold=0.0, overstory_frac=0.3,
              max_veg=3, extra_global="", result_dir=None, skip_output=False, snow_step=None,
              wind_zero_frac=0.0, coords=None, off_gmt=None, force_daily=False, no_wind=False, extra_cols=(), alma=False,
              organic=False):
    """Write a complete reference-format case under `outdir`; return the global file path.

    dt == 24: daily PREC/TMAX/TMIN/WIND forcing (as the bundled basin); dt < 24: sub-daily
    PREC/AIR_TEMP/WIND records at the model step.  `cold` lowers all temperatures (deg C) to
    force snow.  All randomness comes from `seed`.  `off_gmt` (hours) replaces the cells' own time zone
    (round(lng/15)), so that local time differs from the forcing's time; `force_daily` writes the daily
    four-column file even when dt < 24 (FORCE_DT 24 with a sub-daily TIME_STEP); `no_wind` leaves the
    WIND column out (the reference then uses DEFAULT_WIND_SPEED); `extra_cols` appends observed SHORTWAVE (W/m2),
    LONGWAVE (W/m2) and / or VP (kPa) columns; `alma` writes the file in ALMA units (PREC as a rate in mm/s,
    temperatures in K, VP and PRESSURE in Pa) with ALMA_INPUT TRUE.
    """
    rng = np.random.default_rng(seed)
    os.makedirs(outdir, exist_ok=True)
    forc_dir = os.path.join(outdir, "forcing")
    os.makedirs(forc_dir, exist_ok=True)
    result_dir = result_dir or os.path.join(outdir, "results")
    os.makedirs(result_dir, exist_ok=True)
    if snow_step is None:
        snow_step = dt
    L = nlayer

    write_veglib(os.path.join(outdir, "veglib.txt"))
    steps_per_day = 24 // dt
    nrec = ndays * steps_per_day

    soil_lines, veg_lines, band_lines = [], [], []
    for c in range(ncells):
        cellid = c + 1
        lat = round(float(rng.uniform(*lat_range)), 4)
        lng = round(float(rng.uniform(-120.0, 120.0)), 4)
        if coords is not None:                   # cells placed on given points (e.g. routing grid cell centres)
            lat, lng = round(float(coords[c][0]), 4), round(float(coords[c][1]), 4)
        elev = float(np.round(rng.uniform(0.0, 2000.0), 1))
        b_inf = rng.uniform(0.001, 0.4)
        Ds = rng.uniform(0.001, 0.9)
        Dsmax = rng.uniform(1.0, 30.0)
        Ws = rng.uniform(0.3, 0.95)
        cexp = 2.0
        expt = rng.uniform(4.0, 20.0, L)
        Ksat = np.exp(rng.uniform(np.log(50.0), np.log(5000.0), L))
        phi_s = np.full(L, -999.0)
        d1 = rng.uniform(0.05, 0.25)
        d2 = rng.uniform(0.3, 1.5)
        depth = [d1, d2] + [rng.uniform(0.3, 1.5) for _ in range(L - 2)]
        depth = np.array(depth[:L])
        bulk = rng.uniform(1300.0, 1600.0, L)
        sdens = rng.uniform(2650.0, 2685.0, L)
        poros = 1.0 - bulk / sdens
        max_moist = depth * poros * 1000.0
        init_moist = 0.6 * max_moist
        avg_T = rng.uniform(5.0, 27.0) - cold
        dp = 4.0
        bubble = rng.uniform(5.0, 40.0, L)
        quartz = rng.uniform(0.1, 0.8, L)
        off_gmt_c = round(lng / 15.0) if off_gmt is None else off_gmt
        wcr_fr = rng.uniform(0.55, 0.75, L)
        wpwp_fr = rng.uniform(0.2, 0.5, L)
        rough, snow_rough = 0.001, 0.0005
        annual_prec = 1000.0
        resid = rng.uniform(0.0, 0.05, L)
        cols = ["1", str(cellid), "%.4f" % lat, "%.4f" % lng, _fmt(b_inf), _fmt(Ds), _fmt(Dsmax),
                _fmt(Ws), _fmt(cexp)]
        cols += [_fmt(x, 4) for x in expt] + [_fmt(x, 3) for x in Ksat] + ["%.0f" % x for x in phi_s]
        cols += [_fmt(x, 4) for x in init_moist] + ["%.1f" % elev] + [_fmt(x, 4) for x in depth]
        cols += [_fmt(avg_T, 3), _fmt(dp, 1)] + [_fmt(x, 3) for x in bubble] + [_fmt(x, 4) for x in quartz]
        cols += [_fmt(x, 2) for x in bulk] + [_fmt(x, 2) for x in sdens]
        if organic:                      # ORGANIC_FRACT TRUE: organic fraction, organic bulk and soil densities per layer
            org = np.round(np.linspace(0.35, 0.0, L) * rng.uniform(0.5, 1.0), 3)      # organic top layer, mineral bottom layer
            cols += ["%.3f" % x for x in org] + ["%.1f" % x for x in rng.uniform(180.0, 320.0, L)]
            cols += ["%.1f" % x for x in rng.uniform(1250.0, 1350.0, L)]
        cols += ["%g" % off_gmt_c]
        cols += [_fmt(x, 4) for x in wcr_fr] + [_fmt(x, 4) for x in wpwp_fr]
        cols += [_fmt(rough, 4), _fmt(snow_rough, 4), _fmt(annual_prec, 1)] + [_fmt(x, 4) for x in resid] + ["0"]
        soil_lines.append("\t".join(cols))

        # vegetation tiles (0..max_veg) + implicit bare remainder
        nveg = int(rng.integers(0, max_veg + 1)) if max_veg > 0 else 0
        if c == 0:
            nveg = min(max_veg, 2)
        full_cover = (c % 5 == 3) and nveg > 0     # some cells without a bare tile
        cv = rng.dirichlet(np.ones(nveg + 1))[:nveg] if nveg else np.array([])
        if full_cover:
            cv = cv / cv.sum()
        veg_lines.append("%d %d" % (cellid, nveg))
        for v in range(nveg):
            if rng.uniform() < overstory_frac:
                cls = int(rng.choice([1, 2, 3]))
            else:
                cls = int(rng.choice([4, 5, 6]))
            zd = [0.10, 0.60, 0.80]
            zf = rng.dirichlet([2.0, 3.0, 1.0])
            zf = np.round(zf, 3)
            zf[2] = round(1.0 - zf[0] - zf[1], 3)
            cvv = round(float(cv[v]), 4)
            if full_cover and v == nveg - 1:
                cvv = round(1.0 - sum(round(float(x), 4) for x in cv[:-1]), 4)
            veg_lines.append("   %d %.4f %.2f %.3f %.2f %.3f %.2f %.3f" %
                             (cls, cvv, zd[0], zf[0], zd[1], zf[1], zd[2], zf[2]))

        if nbands > 1:
            af = np.round(rng.dirichlet(np.ones(nbands) * 2.0), 3)
            af[-1] = round(1.0 - af[:-1].sum(), 3)
            if c % 4 == 2:                       # an empty band now and then
                af[0] += af[-1]
                af[-1] = 0.0
            span = np.linspace(-600.0, 600.0, nbands)
            el = elev + span
            el = el - (el * af).sum() + elev     # area-weighted mean == cell elevation
            pf = af.copy()
            band_lines.append(" ".join([str(cellid)] + ["%.3f" % x for x in af] +
                                       ["%.2f" % max(x, 0.0) for x in el] + ["%.3f" % x for x in pf]))

        # forcing
        doy = (np.arange(ndays) % 365) + 1
        phase = 2.0 * np.pi * (doy - 200.0) / 365.0 if lat >= 0 else 2.0 * np.pi * (doy - 17.0) / 365.0
        tmean_ann = 27.0 - 0.45 * abs(lat) - 0.0065 * elev - cold
        amp = 2.0 + 0.30 * abs(lat)
        tmean = tmean_ann + amp * np.cos(phase) + rng.normal(0.0, 3.0, ndays)
        dtr = rng.uniform(5.0, 14.0, ndays)
        tmax = tmean + 0.5 * dtr
        tmin = tmean - 0.5 * dtr
        wet = np.zeros(ndays, bool)
        w = False
        u = rng.uniform(size=ndays)
        for i in range(ndays):
            p = 0.6 if w else 0.22           # ~0.35 stationary wet fraction
            w = u[i] < p
            wet[i] = w
        prec = np.where(wet, rng.gamma(0.8, 8.0, ndays), 0.0)
        wind = np.maximum(np.exp(rng.normal(1.0, 0.5, ndays)), 0.1)
        if wind_zero_frac > 0:
            wind = np.where(rng.uniform(size=ndays) < wind_zero_frac, 0.0, wind)
        # optional observed columns (values only need to be plausible)
        if set(extra_cols) - {"RAINF", "SNOWF", "WIND_E", "WIND_N"}:   # drawn only when asked for: the random stream of the other cases stays as it was
            sw_day = np.clip(180.0 + 120.0 * np.cos(phase) * (1 if lat >= 0 else -1) + rng.normal(0, 40, ndays), 5.0, 420.0)
            lw_day = np.clip(300.0 + 2.5 * tmean + rng.normal(0, 15, ndays), 120.0, 480.0)
            vp_day = np.clip(0.6108 * np.exp(17.27 * tmin / (237.3 + tmin)) * rng.uniform(0.7, 1.0, ndays), 0.02, 4.0)

        def extras(i, k=None):
            out = []
            for name in extra_cols:
                if name == "SHORTWAVE":
                    if k is None:
                        out.append(sw_day[i])
                    else:               # a daylight bell over the sub-daily records, zero at night, same daily mean
                        w = np.maximum(np.cos(2.0 * np.pi * (hrs - 12.0) / 24.0), 0.0)
                        out.append(sw_day[i] * w[k] * steps_per_day / max(w.sum(), 1e-9))
                elif name == "LONGWAVE":
                    out.append(lw_day[i] + (0.0 if k is None else 10.0 * np.cos(2.0 * np.pi * (hrs[k] - 15.0) / 24.0)))
                elif name == "PRESSURE":            # kPa
                    out.append(101.3 * np.exp(-elev / 8400.0) + 0.4 * np.sin(0.3 * i + (0 if k is None else 0.2 * k)))
                elif name == "DENSITY":
                    out.append(1.25 * np.exp(-elev / 9000.0) - 0.003 * tmean[i] + (0.0 if k is None else 0.004 * np.cos(hrs[k] / 3.8)))
                elif name == "VP":
                    out.append(vp_day[i] * (1.0 if k is None else 1.0 + 0.05 * np.cos(2.0 * np.pi * (hrs[k] - 15.0) / 24.0)))
                elif name == "QAIR":                # kg/kg
                    out.append(0.622 * vp_day[i] / (101.3 * np.exp(-elev / 8400.0)) * (1.0 if k is None else 1.0 + 0.04 * np.sin(hrs[k] / 3.0)))
                elif name == "REL_HUMID":           # %
                    out.append(float(np.clip(55.0 + 30.0 * np.sin(0.37 * i) + (0.0 if k is None else 8.0 * np.cos(2.0 * np.pi * (hrs[k] - 4.0) / 24.0)),
                                             8.0, 100.0)))
                elif name in ("RAINF", "SNOWF"):    # together they replace PREC (and differ from it: 90 % of it)
                    sf = 1.0 if tmean[i] < -1.0 else (0.5 if tmean[i] < 1.0 else 0.0)
                    amount = 0.9 * prec[i] * (sf if name == "SNOWF" else 1.0 - sf) / (1 if k is None else steps_per_day)
                    out.append(amount / ((24 if k is None else dt) * 3600.0) if alma else amount)
                elif name in ("WIND_E", "WIND_N"):  # together they replace WIND
                    ang = 0.9 * i + (0.0 if k is None else 0.3 * k)
                    out.append(1.1 * wind[i] * (np.cos(ang) if name == "WIND_E" else np.sin(ang)))
            if alma:
                out = [x * 1000.0 if name in ("VP", "PRESSURE") else x for x, name in zip(out, extra_cols)]
            fmt = {"QAIR": " %.7f", "RAINF": " %.8e" if alma else " %.4f", "SNOWF": " %.8e" if alma else " %.4f"}
            return "".join(fmt.get(name, " %.4f") % x for x, name in zip(out, extra_cols))
        hrs = (np.arange(steps_per_day) + 0.5) * dt if dt < 24 else None
        fn = os.path.join(forc_dir, "data_%.4f_%.4f" % (lat, lng))
        with open(fn, "w") as f:
            if dt == 24 or force_daily:
                for i in range(ndays):
                    if alma:
                        f.write("%.8e %.4f %.4f %.4f%s\n" % (prec[i] / 86400.0, tmax[i] + 273.15, tmin[i] + 273.15, wind[i], extras(i)))
                    elif no_wind:
                        f.write("%.4f %.4f %.4f%s\n" % (prec[i], tmax[i], tmin[i], extras(i)))
                    else:
                        f.write("%.4f %.4f %.4f %.4f%s\n" % (prec[i], tmax[i], tmin[i], wind[i], extras(i)))
            else:
                hrs = (np.arange(steps_per_day) + 0.5) * dt
                for i in range(ndays):
                    for k in range(steps_per_day):
                        ta = tmean[i] + 0.5 * dtr[i] * np.cos(2.0 * np.pi * (hrs[k] - 15.0) / 24.0)
                        if alma:
                            f.write("%.8e %.4f %.4f%s\n" % (prec[i] / 86400.0, ta + 273.15, wind[i], extras(i, k)))
                        else:
                            f.write("%.4f %.4f %.4f%s\n" % (prec[i] / steps_per_day, ta, wind[i], extras(i, k)))

    with open(os.path.join(outdir, "soil.txt"), "w") as f:
        f.write("\n".join(soil_lines) + "\n")
    with open(os.path.join(outdir, "veg.txt"), "w") as f:
        f.write("\n".join(veg_lines) + "\n")
    if nbands > 1:
        with open(os.path.join(outdir, "snowband.txt"), "w") as f:
            f.write("\n".join(band_lines) + "\n")

    g = []
    g += ["NLAYER %d" % L, "NODES 3", "TIME_STEP %d" % dt, "SNOW_STEP %d" % snow_step,
          "STARTYEAR %d" % startyear, "STARTMONTH 1", "STARTDAY 1", "STARTHOUR 0",
          "NRECS %d" % nrec,
          "FULL_ENERGY FALSE", "FROZEN_SOIL FALSE", "NO_FLUX FALSE", "CORRPREC FALSE",
          "MIN_WIND_SPEED 0.1", "MAX_SNOW_TEMP 0.5", "MIN_RAIN_TEMP -0.5",
          "FORCING1 %s/data_" % forc_dir, "FORCE_FORMAT ASCII", "FORCE_ENDIAN LITTLE"]
    if dt == 24 or force_daily:
        types = ["PREC", "TMAX", "TMIN"] + ([] if no_wind else ["WIND"]) + list(extra_cols)
        g += ["N_TYPES %d" % len(types)] + ["FORCE_TYPE %s" % t for t in types] + ["FORCE_DT 24"]
    else:
        types = ["PREC", "AIR_TEMP", "WIND"] + list(extra_cols)
        g += ["N_TYPES %d" % len(types)] + ["FORCE_TYPE %s" % t for t in types] + ["FORCE_DT %d" % dt]
    g += ["FORCEYEAR %d" % startyear, "FORCEMONTH 1", "FORCEDAY 1", "FORCEHOUR 0", "GRID_DECIMAL 4",
          "WIND_H 10.0", "MEASURE_H 2.0", "ALMA_INPUT %s" % ("TRUE" if alma else "FALSE"),
          "SOIL %s/soil.txt" % outdir, "BASEFLOW ARNO", "JULY_TAVG_SUPPLIED FALSE", "ORGANIC_FRACT %s" % ("TRUE" if organic else "FALSE"),
          "VEGLIB %s/veglib.txt" % outdir, "VEGPARAM %s/veg.txt" % outdir, "ROOT_ZONES 3",
          "VEGPARAM_LAI FALSE", "LAI_SRC FROM_VEGLIB"]
    if nbands > 1:
        g += ["SNOW_BAND %d %s/snowband.txt" % (nbands, outdir)]
    else:
        g += ["SNOW_BAND 1"]
    g += ["COMPUTE_TREELINE FALSE", "RESULT_DIR %s/" % result_dir, "OUT_STEP 0",
          "SKIPYEAR %d" % ((nrec * dt // 24) // 366 if skip_output else 0),   # make_dmy.c:163-169 indexes dmy[] per skipped year
          "COMPRESS FALSE", "BINARY_OUTPUT FALSE", "ALMA_OUTPUT FALSE", "MOISTFRACT FALSE",
          "PRT_HEADER FALSE", "PRT_SNOW_BAND FALSE"]
    if extra_global:
        g += [extra_global]
    gpath = os.path.join(outdir, "global.txt")
    with open(gpath, "w") as f:
        f.write("\n".join(g) + "\n")
    return gpath
