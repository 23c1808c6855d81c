"""Host-side mirror of the reference's forcing preparation over the C ABI of include/vicres_atmos.h.

The reference calls, once per cell, `initialize_atmos(atmos, dmy, filep.forcing, ...)`
(runoff/model/vicNl.c:239-241); `prepare()` does that for all cells at once on the GPU:

    o = atmos.options(time_step=24, nrecs=3653, startyear=1981)         # global-file values
    f = atmos.prepare(o, site, raw)         # [9, nrecs, C]: what VicRes.step() takes as forcing

ctypes declarations and marshalling only; every number is computed by the CUDA library
(vicres_b200/csrc/vicres_atmos.cu).  There is no CPU path behind it.
"""
import ctypes as C
import numpy as np
from . import abi, model

DAILY_TMAX_TMIN, SUBDAILY_AIR_TEMP = 0, 1
VP_ITER = {"VP_ITER_NONE": 0, "VP_ITER_ALWAYS": 1, "VP_ITER_ANNUAL": 2, "VP_ITER_CONVERGE": 3}
LW_TYPE = {"LW_TVA": 0, "LW_ANDERSON": 1, "LW_BRUTSAERT": 2, "LW_SATTERLUND": 3, "LW_IDSO": 4, "LW_PRATA": 5}
LW_CLOUD = {"LW_CLOUD_BRAS": 0, "LW_CLOUD_DEARDORFF": 1}

EXPORTS = ["vr_atmos_default_options", "vr_atmos_ndays", "vr_atmos_prepare", "vr_atmos_prepare_device",
           "vr_atmos_prepare_sub", "vr_atmos_prepare_sub_device",
           "vr_atmos_last_times", "vr_atmos_last_error", "vr_atmos_diag_cr", "vr_atmos_diag_solar"]

_pd = C.POINTER(C.c_double)


class vr_atmos_options(C.Structure):
    _fields_ = [("time_step", C.c_int32), ("nrecs", C.c_int32), ("startyear", C.c_int32), ("startmonth", C.c_int32),
                ("startday", C.c_int32), ("starthour", C.c_int32), ("layout", C.c_int32), ("force_dt", C.c_int32),
                ("has_wind", C.c_int32), ("vp_interp", C.c_int32), ("vp_iter", C.c_int32),
                ("mtclim_swe_corr", C.c_int32), ("lw_type", C.c_int32), ("lw_cloud", C.c_int32), ("plapse", C.c_int32),
                ("alma_input", C.c_int32), ("min_wind_speed", C.c_double), ("sw_prec_thresh", C.c_double),
                ("reserved", C.c_int32 * 8), ("snow_step", C.c_int32), ("pad_", C.c_int32), ("max_snow_temp", C.c_double)]


SITE_FIELDS = ["lat", "lng", "time_zone_lng", "elevation", "annual_prec", "slope", "aspect", "ehoriz", "whoriz", "min_Tfactor"]


class vr_atmos_cells(C.Structure):
    _fields_ = [("ncell", C.c_int64)] + [(n, _pd) for n in SITE_FIELDS]


class vr_atmos_raw(C.Structure):
    _fields_ = [("nrec_forcing", C.c_int64), ("prec", _pd), ("t_a", _pd), ("t_b", _pd), ("wind", _pd),
                ("shortwave", _pd), ("longwave", _pd), ("vp", _pd), ("pressure", _pd), ("density", _pd),
                ("qair", _pd), ("rel_humid", _pd), ("rainf", _pd), ("snowf", _pd), ("wind_e", _pd), ("wind_n", _pd)]


RAW_FIELDS = ("prec", "t_a", "t_b", "wind", "shortwave", "longwave", "vp", "pressure", "density", "qair", "rel_humid", "rainf",
              "snowf", "wind_e", "wind_n")


def options(time_step=24, nrecs=0, startyear=1981, startmonth=1, startday=1, starthour=0, layout=None, force_dt=None,
            has_wind=True, vp_interp=True, vp_iter="VP_ITER_ALWAYS", mtclim_swe_corr=False, lw_type="LW_PRATA",
            lw_cloud="LW_CLOUD_DEARDORFF", plapse=True, min_wind_speed=0.1, sw_prec_thresh=0.0, alma_input=False,
            snow_step=0, max_snow_temp=0.5):
    """vr_atmos_options with the reference's defaults (initialize_global.c:146-211)."""
    o = vr_atmos_options()
    if force_dt is None:
        force_dt = 24 if layout in (None, DAILY_TMAX_TMIN) else time_step
    if layout is None:
        layout = DAILY_TMAX_TMIN if force_dt == 24 else SUBDAILY_AIR_TEMP
    o.time_step, o.nrecs, o.startyear, o.startmonth, o.startday, o.starthour = time_step, nrecs, startyear, startmonth, startday, starthour
    o.layout, o.force_dt, o.has_wind = layout, force_dt, int(has_wind)
    o.vp_interp, o.vp_iter, o.mtclim_swe_corr = int(vp_interp), VP_ITER.get(vp_iter, vp_iter), int(mtclim_swe_corr)
    o.lw_type, o.lw_cloud, o.plapse = LW_TYPE.get(lw_type, lw_type), LW_CLOUD.get(lw_cloud, lw_cloud), int(plapse)
    o.min_wind_speed, o.sw_prec_thresh = min_wind_speed, sw_prec_thresh
    o.alma_input = int(alma_input)
    o.snow_step, o.max_snow_temp = int(snow_step), float(max_snow_temp)
    return o


def n_windows(opt):
    """NF = TIME_STEP / SNOW_STEP"""
    return opt.time_step // opt.snow_step if opt.snow_step > 0 else 1


class Packed:
    def __init__(self, struct, keep):
        self.struct, self._keep = struct, keep

    def ref(self):
        return C.byref(self.struct)


def pack_site(site):
    """site: dict of [C] arrays lat, lng, time_zone_lng, elevation, annual_prec (+ optional slope ...)."""
    s, keep = vr_atmos_cells(), []
    s.ncell = len(site["lat"])
    for n in SITE_FIELDS:
        if site.get(n) is not None:
            a = np.ascontiguousarray(site[n], dtype=np.float64)
            keep.append(a)
            setattr(s, n, a.ctypes.data_as(_pd))
    return Packed(s, keep)


def pack_raw(raw):
    """raw: dict prec, t_a, t_b (or None), wind (or None) and optionally shortwave, longwave, vp (kPa), each [nrec_forcing, C]."""
    s, keep = vr_atmos_raw(), []
    s.nrec_forcing = raw["t_a"].shape[0]
    for n in RAW_FIELDS:
        if raw.get(n) is not None:
            a = np.ascontiguousarray(raw[n], dtype=np.float64)
            keep.append(a)
            setattr(s, n, a.ctypes.data_as(_pd))
    return Packed(s, keep)


def _lib():
    L = model.load_library()
    if not getattr(L, "_atmos_declared", False):
        vp = C.c_void_p
        L.vr_atmos_default_options.argtypes = [vp]
        L.vr_atmos_default_options.restype = None
        L.vr_atmos_ndays.argtypes = [vp]
        L.vr_atmos_prepare.argtypes = [C.c_int, vp, vp, vp, vp]
        L.vr_atmos_prepare_device.argtypes = [C.c_int, vp, vp, vp, vp, vp]
        L.vr_atmos_prepare_sub.argtypes = [C.c_int, vp, vp, vp, vp, vp]
        L.vr_atmos_prepare_sub_device.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp]
        L.vr_atmos_last_times.argtypes = [vp, vp]
        L.vr_atmos_last_error.restype = C.c_char_p
        L.vr_atmos_diag_cr.argtypes = [C.c_int, C.c_int64, vp, vp, vp, vp]
        L.vr_atmos_diag_solar.argtypes = [C.c_int, C.c_int64, vp, vp, vp, vp, vp, C.c_int32]
        L._atmos_declared = True
    return L


def _check(L, rc):
    if rc != 0:
        raise model.VicResError(rc, (L.vr_atmos_last_error() or b"").decode())


def prepare(opt, site, raw, device=0):
    """initialize_atmos for all cells: returns the nine atmos fields [9, nrecs, C] (host arrays)."""
    L = _lib()
    ps, pr = pack_site(site), pack_raw(raw)
    Cn = ps.struct.ncell
    out = np.zeros((abi.VR_NFORCING, opt.nrecs, Cn))
    ptrs = (_pd * abi.VR_NFORCING)(*[out[f].ctypes.data_as(_pd) for f in range(abi.VR_NFORCING)])
    _check(L, L.vr_atmos_prepare(device, C.byref(opt), ps.ref(), pr.ref(), ptrs))
    return out


def prepare_sub(opt, site, raw, device=0):
    """initialize_atmos with the snow flags, and with the windows of a sub-stepped snow model when opt.snow_step <
    opt.time_step: returns (fields, snowflag); fields [9, nrecs, C] for NF == 1, else [9, nrecs, NF + 1, C] (the NF
    windows, then the record's own values: atmos-><field>[0..NF-1], [NR]); snowflag [nrecs, C] int32."""
    L = _lib()
    ps, pr = pack_site(site), pack_raw(raw)
    Cn = ps.struct.ncell
    NF = n_windows(opt)
    shape = (abi.VR_NFORCING, opt.nrecs, Cn) if NF == 1 else (abi.VR_NFORCING, opt.nrecs, NF + 1, Cn)
    out = np.zeros(shape)
    flag = np.zeros((opt.nrecs, Cn), dtype=np.int32)
    ptrs = (_pd * abi.VR_NFORCING)(*[out[f].ctypes.data_as(_pd) for f in range(abi.VR_NFORCING)])
    _check(L, L.vr_atmos_prepare_sub(device, C.byref(opt), ps.ref(), pr.ref(), ptrs, flag.ctypes.data))
    return out, flag


def prepare_device(opt, site_dev, raw_dev, out_dev, ncell, nrec_forcing, device=0, stream=None, snowflag_dev=None):
    """Device-resident variant: site_dev / raw_dev map names to device pointers (ints), out_dev is a list of
    nine device pointers, each [nrecs*C] doubles ([nrecs*(NF+1)*C] with opt.snow_step < opt.time_step); snowflag_dev:
    device pointer to [nrecs*C] int32, or None."""
    L = _lib()
    s = vr_atmos_cells()
    s.ncell = ncell
    for n in SITE_FIELDS:
        if site_dev.get(n):
            setattr(s, n, C.cast(site_dev[n], _pd))
    r = vr_atmos_raw()
    r.nrec_forcing = nrec_forcing
    for n in RAW_FIELDS:
        if raw_dev.get(n):
            setattr(r, n, C.cast(raw_dev[n], _pd))
    ptrs = (_pd * abi.VR_NFORCING)(*[C.cast(p, _pd) for p in out_dev])
    _check(L, L.vr_atmos_prepare_sub_device(device, C.byref(opt), C.byref(s), C.byref(r), ptrs, snowflag_dev, stream))


def last_times():
    """(solar-table ms, MTCLIM daily ms, disaggregation ms, launches) of the last prepare call."""
    L = _lib()
    ms = (C.c_double * 3)()
    n = C.c_int64()
    L.vr_atmos_last_times(ms, C.byref(n))
    return ms[0], ms[1], ms[2], n.value


def options_from_global(g):
    """vr_atmos_options from a parsed global file (vicres_b200.files.Global)."""
    g2 = g.second_forcing() if hasattr(g, "second_forcing") else None
    types = [t for t in g.force_types + (g2.force_types if g2 is not None else []) if t != "SKIP"]
    extra = [t for t in types if t not in ("PREC", "TMAX", "TMIN", "AIR_TEMP", "WIND", "SHORTWAVE", "LONGWAVE", "VP", "PRESSURE", "DENSITY",
                                           "QAIR", "REL_HUMID", "RAINF", "SNOWF", "WIND_E", "WIND_N")]
    if extra:
        raise model.VicResError(abi.VR_ERR_UNSUPPORTED, "forcing column(s) %s are outside the supported layouts" % ", ".join(extra))
    layout = SUBDAILY_AIR_TEMP if "AIR_TEMP" in types else DAILY_TMAX_TMIN
    return options(time_step=g.time_step, nrecs=g.nrecs, startyear=g.startyear, startmonth=g.startmonth,
                   startday=g.startday, starthour=g.starthour, layout=layout, force_dt=g.force_dt,
                   has_wind="WIND" in types or ("WIND_E" in types and "WIND_N" in types), min_wind_speed=g.min_wind_speed, vp_interp=g.vp_interp, vp_iter=g.vp_iter,
                   mtclim_swe_corr=g.mtclim_swe_corr, lw_type=g.lw_type, lw_cloud=g.lw_cloud, plapse=g.plapse,
                   sw_prec_thresh=g.sw_prec_thresh, alma_input=g.alma_input,
                   snow_step=g.snow_step if g.snow_step != g.time_step else 0, max_snow_temp=g.max_snow_temp)


def raw_from_columns(cols, layout):
    """{FORCE_TYPE: [nrec, C]} -> the dict prepare() takes."""
    return {"prec": cols.get("PREC"), "t_a": cols["TMAX" if layout == DAILY_TMAX_TMIN else "AIR_TEMP"],
            "t_b": cols.get("TMIN") if layout == DAILY_TMAX_TMIN else None, "wind": cols.get("WIND"),
            "shortwave": cols.get("SHORTWAVE"), "longwave": cols.get("LONGWAVE"), "vp": cols.get("VP"),
            "pressure": cols.get("PRESSURE"), "density": cols.get("DENSITY"), "qair": cols.get("QAIR"),
            "rel_humid": cols.get("REL_HUMID"), "rainf": cols.get("RAINF"), "snowf": cols.get("SNOWF"),
            "wind_e": cols.get("WIND_E"), "wind_n": cols.get("WIND_N")}


def diag_cr(x, y, device=0):
    """The device's correctly rounded acos(x), cos(y) (diagnostic for the tests)."""
    L = _lib()
    x, y = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
    a, c = np.zeros_like(x), np.zeros_like(y)
    _check(L, L.vr_atmos_diag_cr(device, len(x), x.ctypes.data, a.ctypes.data, y.ctypes.data, c.ctypes.data))
    return a, c


def diag_solar(lat, elevation, lng=None, time_zone_lng=None, device=0, exact_walk=False):
    """The device's solar-geometry tables of n cells: dict ttmax0/flat/slope/daylength [365, n], hf [365, 24, n]."""
    L = _lib()
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    lat, elevation, lng, time_zone_lng = f(lat), f(elevation), f(lng), f(time_zone_lng)
    n = len(lat)
    tab = np.zeros((365 * 28, n))
    p = lambda a: None if a is None else a.ctypes.data
    _check(L, L.vr_atmos_diag_solar(device, n, p(lat), p(elevation), p(lng), p(time_zone_lng), tab.ctypes.data, int(exact_walk)))
    return {"ttmax0": tab[0:365], "flat": tab[365:730], "slope": tab[730:1095], "daylength": tab[1095:1460],
            "hf": tab[1460:].reshape(365, 24, n)}
