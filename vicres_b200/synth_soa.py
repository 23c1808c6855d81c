"""Seeded synthetic inputs generated DIRECTLY in the struct-of-arrays layout of include/vicres.h,
for sizes where writing one reference-format forcing file per cell is impractical (BASELINE.json
configs 2-4: 1e5..1e6 cells).  Parameter ranges are the ones SURVEY.md section 8(d) fixes; derived
quantities follow read_soilparam.c:640-900 (max_moist = depth*porosity*1000, Wcr/Wpwp fractions of
max_moist, mineral == bulk densities without organic matter).  The vegetation library is the one of
vicres_b200/synth.py.  Forcing holds the 9 atmos fields the step consumes (what initialize_atmos
would hand over), built from a seasonal/latitudinal climate with a wet-day Markov chain.
"""
import numpy as np
from . import synth

A_SVP, B_SVP, C_SVP = 0.61078, 17.269, 237.3


def veglib():
    n = len(synth.VEGLIB)
    lib = {k: np.zeros(n) for k in ["rarc", "rmin", "wind_h", "RGL", "rad_atten", "wind_atten", "trunk_ratio"]}
    lib["overstory"] = np.zeros(n, dtype=np.int32)
    lib["roughness"] = np.zeros((n, 12))
    lib["displacement"] = np.zeros((n, 12))
    lai = np.zeros((n, 12))
    alb = np.zeros((n, 12))
    for i, (cid, ov, rarc, rmin, l, a, rough, displ, wind_h, rgl, ra, wa, tr) in enumerate(synth.VEGLIB):
        lib["overstory"][i] = ov
        lib["rarc"][i], lib["rmin"][i], lib["wind_h"][i] = rarc, rmin, wind_h
        lib["RGL"][i] = np.float32(rgl)
        lib["rad_atten"][i], lib["wind_atten"][i], lib["trunk_ratio"][i] = ra, wa, tr
        lib["roughness"][i, :] = rough
        lib["displacement"][i, :] = displ
        lai[i, :] = l
        alb[i, :] = a
    return lib, lai, alb


def make_cells(ncell, nlayer=3, nbands=1, seed=20261019, lat_range=(-40.0, 60.0), max_veg=3,
               overstory_frac=0.3):
    rng = np.random.default_rng(seed)
    C, L = ncell, nlayer
    lib, lib_lai, lib_alb = veglib()
    cells = {}
    cells["lat"] = np.float32(rng.uniform(lat_range[0], lat_range[1], C)).astype(np.float64)
    cells["elevation"] = np.float32(rng.uniform(0.0, 2000.0, C)).astype(np.float64)
    cells["b_infilt"] = rng.uniform(0.001, 0.4, C)
    cells["Ds"] = rng.uniform(0.001, 0.9, C)
    cells["Dsmax"] = rng.uniform(1.0, 30.0, C)
    cells["Ws"] = rng.uniform(0.3, 0.95, C)
    cells["c_expt"] = np.full(C, 2.0)
    # soil_con.avg_temp is the mean annual air temperature: keep it consistent with the forcing climate
    cells["avg_temp"] = 27.0 - 0.45 * np.abs(cells["lat"]) - 0.0065 * cells["elevation"] + rng.normal(0.0, 1.0, C)
    cells["dp"] = np.full(C, 4.0)
    cells["rough"] = np.full(C, 0.001)
    cells["snow_rough"] = np.full(C, 0.0005)
    cells["expt"] = rng.uniform(4.0, 20.0, (L, C))
    cells["Ksat"] = np.exp(rng.uniform(np.log(50.0), np.log(5000.0), (L, C)))
    depth = np.empty((L, C))
    depth[0] = rng.uniform(0.05, 0.25, C)
    for l in range(1, L):
        depth[l] = rng.uniform(0.3, 1.5, C)
    depth = (np.float32(np.floor(depth * 1000.0 + 0.5)) / np.float32(1000.0)).astype(np.float64)  # read_soilparam.c:342
    cells["depth"] = depth
    bulk = rng.uniform(1300.0, 1600.0, (L, C))
    sdens = rng.uniform(2650.0, 2685.0, (L, C))
    poros = 1.0 - bulk / sdens
    cells["bulk_density"], cells["soil_density"] = bulk, sdens
    cells["bulk_dens_min"], cells["soil_dens_min"] = bulk.copy(), sdens.copy()
    cells["porosity"] = poros
    cells["organic"] = np.zeros((L, C))
    cells["quartz"] = rng.uniform(0.1, 0.8, (L, C))
    mm = depth * poros * 1000.0
    cells["max_moist"] = mm
    cells["Wcr"] = rng.uniform(0.55, 0.75, (L, C)) * mm
    cells["Wpwp"] = rng.uniform(0.2, 0.5, (L, C)) * mm
    cells["resid_moist"] = rng.uniform(0.0, 0.05, (L, C))
    cells["init_moist"] = 0.6 * mm
    # snow bands
    if nbands > 1:
        af = rng.dirichlet(np.ones(nbands) * 2.0, C).T            # [B, C]
        span = np.linspace(-600.0, 600.0, nbands)[:, None]
        el = cells["elevation"][None, :] + span
        el = el - (el * af).sum(0) + cells["elevation"][None, :]
        cells["AreaFract"] = af
        cells["Tfactor"] = (el - cells["elevation"][None, :]) * (-0.0065)      # read_snowband.c: T_LAPSE
        pf = af.copy()
        cells["Pfactor"] = pf / np.maximum(af, 1e-12) * 1.0 if False else np.ones_like(af)
    else:
        cells["AreaFract"] = np.ones((1, C))
        cells["Tfactor"] = np.zeros((1, C))
        cells["Pfactor"] = np.ones((1, C))
    # tiles: nveg in 0..max_veg real tiles + bare remainder (a fifth of the vegetated cells fully covered)
    nveg = rng.integers(0, max_veg + 1, C) if max_veg > 0 else np.zeros(C, dtype=np.int64)
    full = (np.arange(C) % 5 == 3) & (nveg > 0)
    ntile = nveg + (~full).astype(np.int64)
    tile_ptr = np.zeros(C + 1, dtype=np.int64)
    np.cumsum(ntile, out=tile_ptr[1:])
    T = int(tile_ptr[-1])
    cell_of = np.repeat(np.arange(C), ntile)
    k_in_cell = np.arange(T) - tile_ptr[cell_of]
    is_bare = (k_in_cell == nveg[cell_of]) & (~full[cell_of])
    ov = rng.uniform(size=T) < overstory_frac
    cls = np.where(ov, rng.integers(0, 3, T), rng.integers(3, 6, T)).astype(np.int32)
    ncls = len(synth.VEGLIB)
    cls[is_bare] = ncls
    w = rng.gamma(1.0, 1.0, T) + 0.05
    wsum = np.zeros(C)
    np.add.at(wsum, cell_of, w)
    cv = w / wsum[cell_of]
    zf = rng.dirichlet([2.0, 3.0, 1.0], T)                        # root fractions per layer
    root = np.zeros((T, L))
    root[:, :min(L, 3)] = zf[:, :min(L, 3)]
    if L == 2:
        root[:, 1] = zf[:, 1] + zf[:, 2]
    root = np.float32(root).astype(np.float64)
    root[is_bare] = 0.0
    safe_cls = np.where(is_bare, 0, cls)
    lai = lib_lai[safe_cls]
    alb = lib_alb[safe_cls]
    lai[is_bare] = 0.0
    alb[is_bare] = 0.0
    vc = np.ones((T, 12))
    vc[is_bare] = 0.0
    cells.update(tile_ptr=tile_ptr, tile_class=cls, tile_Cv=cv, tile_root=root, tile_LAI=lai,
                 tile_albedo=alb, tile_vegcover=vc)
    # library LAI / albedo (read by the NATVEG / VEGNOCR potential evaporation) and soil moisture v water table
    # curves (read by OUT_ZWT*).  The curves here are synthetic monotone tables of the right shape -- the real
    # ones (read_soilparam.c:822-921) are exercised by the golden fixtures and by csrc/vicres_files.cpp.
    lib["LAI"], lib["albedo"] = lib_lai, lib_alb
    frac = (np.arange(11) / 10.0)[None, :, None]                              # [1, 11, 1]
    zz = np.zeros((L + 2, 11, C))
    zm = np.zeros((L + 2, 11, C))
    top = np.zeros(C)
    for l in range(L):
        zz[l] = -(top[None, :] + depth[l][None, :] * frac[0]) * 100.0
        zm[l] = mm[l][None, :] * (1.0 - 0.4 * frac[0] ** 1.5)
        top = top + depth[l]
    zz[L] = -(depth[:L - 1].sum(0)[None, :] * frac[0]) * 100.0
    zm[L] = mm[:L - 1].sum(0)[None, :] * (1.0 - 0.4 * frac[0] ** 1.5)
    zz[L + 1] = -(depth.sum(0)[None, :] * frac[0]) * 100.0
    zm[L + 1] = mm.sum(0)[None, :] * (1.0 - 0.4 * frac[0] ** 1.5)
    cells["zwt_zwt"], cells["zwt_moist"] = zz, zm
    return lib, cells


def _svp(t):
    s = A_SVP * np.exp(B_SVP * t / (C_SVP + t))
    s = np.where(t < 0, s * (1.0 + .00972 * t + .000042 * t * t), s)
    return s * 1000.0


def make_forcing(cells, nrec, dt_hours=24, seed=7, start_doy=1, cold=0.0, xp=np, rng=None):
    """[9, nrec, C] float64 atmos fields + month[nrec], doy[nrec] (365-day calendar)."""
    lat, elev = cells["lat"], cells["elevation"]
    C = len(lat)
    rng = rng or np.random.default_rng(seed)
    per_day = 24 // dt_hours
    day = (np.arange(nrec) // per_day)
    doy = ((start_doy - 1 + day) % 365) + 1
    cum = np.cumsum([0, 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    month = np.searchsorted(cum, doy - 1, side="right").astype(np.int32)
    hour = (np.arange(nrec) % per_day + 0.5) * dt_hours
    phase = np.where(lat[None, :] >= 0, 2 * np.pi * (doy[:, None] - 200.0) / 365.0,
                     2 * np.pi * (doy[:, None] - 17.0) / 365.0)
    tmean = (27.0 - 0.45 * np.abs(lat) - 0.0065 * elev - cold)[None, :] + (2.0 + 0.30 * np.abs(lat))[None, :] * np.cos(phase)
    tmean = tmean + rng.normal(0.0, 3.0, (nrec, C))
    dtr = rng.uniform(5.0, 14.0, (nrec, C))
    if per_day > 1:
        tair = tmean + 0.5 * dtr * np.cos(2 * np.pi * (hour[:, None] - 15.0) / 24.0)
    else:
        tair = tmean
    wet = rng.uniform(size=(nrec, C)) < 0.35
    prec = np.where(wet, rng.gamma(0.8, 8.0, (nrec, C)), 0.0) / per_day
    wind = np.maximum(np.exp(rng.normal(1.0, 0.5, (nrec, C))), 0.1)
    rh = np.clip(rng.uniform(0.35, 0.95, (nrec, C)) + 0.1 * wet, 0.2, 1.0)
    es = _svp(tair)
    vp = rh * es
    vpd = es - vp
    pressure = 101325.0 * np.exp(-elev / 8434.5)[None, :] * np.ones((nrec, 1))
    density = pressure / (287.0 * (tair + 273.15))
    sw_season = 190.0 + 130.0 * np.cos(phase) * np.sign(1.0) * (np.abs(lat)[None, :] / 60.0 + 0.3)
    if per_day > 1:
        diurnal = np.maximum(np.cos(2 * np.pi * (hour[:, None] - 12.0) / 24.0), 0.0) * np.pi
    else:
        diurnal = 1.0
    shortwave = np.maximum(sw_season * (1.0 - 0.5 * wet) * diurnal, 0.0)
    longwave = (0.72 + 0.15 * wet) * 5.6696e-8 * (tair + 273.15) ** 4
    f = np.stack([tair, prec, wind, vp, vpd, pressure, density, shortwave, longwave]).astype(np.float64)
    return np.ascontiguousarray(f), month, doy.astype(np.int32)


def make_raw_forcing(lat, elevation, ndays, dt_hours=24, seed=11, cold=0.0, start_doy=1, decimals=4):
    """Seeded raw forcing COLUMNS (what the reference's forcing files hold): daily
    PREC/TMAX/TMIN/WIND for dt_hours == 24, else PREC/AIR_TEMP/WIND records of dt_hours.  Returns the dict
    vicres_b200.atmos.prepare takes; values rounded like a "%.4f" text column."""
    rng = np.random.default_rng(seed)
    C = len(lat)
    doy = ((np.arange(ndays) + start_doy - 1) % 365 + 1)[:, None]
    north = (lat >= 0)[None, :]
    phase = np.where(north, 2.0 * np.pi * (doy - 200.0) / 365.0, 2.0 * np.pi * (doy - 17.0) / 365.0)
    tmean = (27.0 - 0.45 * np.abs(lat) - 0.0065 * elevation - cold)[None, :] + (2.0 + 0.30 * np.abs(lat))[None, :] * np.cos(phase)
    tmean = tmean + rng.normal(0.0, 3.0, (ndays, C))
    dtr = rng.uniform(5.0, 14.0, (ndays, C))
    u = rng.uniform(size=(ndays, C))
    wet = np.zeros((ndays, C), bool)
    w = np.zeros(C, bool)
    for i in range(ndays):
        w = u[i] < np.where(w, 0.6, 0.22)
        wet[i] = w
    prec = np.where(wet, rng.gamma(0.8, 8.0, (ndays, C)), 0.0)
    wind = np.maximum(np.exp(rng.normal(1.0, 0.5, (ndays, C))), 0.1)
    r = lambda a: np.ascontiguousarray(np.round(a, decimals))
    if dt_hours == 24:
        return {"prec": r(prec), "t_a": r(tmean + 0.5 * dtr), "t_b": r(tmean - 0.5 * dtr), "wind": r(wind)}
    spd = 24 // dt_hours
    hrs = (np.arange(spd) + 0.5) * dt_hours
    ta = tmean[:, None, :] + 0.5 * dtr[:, None, :] * np.cos(2.0 * np.pi * (hrs - 15.0) / 24.0)[None, :, None]
    return {"prec": r(np.repeat(prec / spd, spd, axis=0)), "t_a": r(ta.reshape(ndays * spd, C)), "t_b": None,
            "wind": r(np.repeat(wind, spd, axis=0))}


def make_site(ncell, seed=3, lat_range=(-40.0, 60.0), own_time_zone=True):
    """Seeded site values as soil_con holds them (floats widened to double)."""
    rng = np.random.default_rng(seed)
    lat = np.float32(rng.uniform(lat_range[0], lat_range[1], ncell)).astype(np.float64)
    lng = np.float32(rng.uniform(-120.0, 120.0, ncell)).astype(np.float64)
    off_gmt = np.rint(lng / 15.0) if own_time_zone else np.zeros(ncell)
    return {"lat": lat, "lng": lng, "time_zone_lng": np.float32(off_gmt * 360.0 / 24.0).astype(np.float64),
            "elevation": np.float32(rng.uniform(0.0, 2000.0, ncell)).astype(np.float64),
            "annual_prec": np.full(ncell, 1000.0)}
