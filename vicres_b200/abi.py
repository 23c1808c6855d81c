"""ctypes mirror of include/vicres.h (the C ABI of the B200 VIC-Res step).

Pure declarations: struct layouts, enum values and helpers that pack numpy arrays into the
plain-pointer structs.  No compute here.  The same packed structs are accepted by the product
library (vicres_b200/csrc -> libvicres.so) and, in tests only, by the oracle (oracle/_build).
"""
import ctypes as C
import numpy as np

VR_MAX_LAYERS = 3
VR_MAX_BANDS = 10
VR_MAX_OUT = 64

VR_OK, VR_ERR_ARG, VR_ERR_UNSUPPORTED, VR_ERR_CUDA, VR_ERR_STATE, VR_ERR_CELL = 0, -1, -2, -3, -4, -999

FORCING_FIELDS = ["air_temp", "prec", "wind", "vp", "vpd", "pressure", "density", "shortwave", "longwave"]
VR_NFORCING = len(FORCING_FIELDS)

OUT_NAMES = [
    "PREC", "EVAP", "RUNOFF", "BASEFLOW", "WDEW", "SOIL_LIQ0", "SOIL_LIQ1", "SOIL_LIQ2",
    "NET_SHORT", "R_NET", "EVAP_CANOP", "TRANSP_VEG", "EVAP_BARE", "SUB_CANOP", "SUB_SNOW",
    "AERO_RESIST", "SURF_TEMP", "ALBEDO", "REL_HUMID", "IN_LONG", "AIR_TEMP", "WIND",
    "SWE", "SNOW_DEPTH", "SNOW_CANOPY", "SNOW_COVER",
    "WATER_ERROR", "INFLOW", "SNOW_MELT", "RAINF", "SNOWF", "NET_LONG", "ASAT", "SOIL_WET",
    "ROOTMOIST", "GRND_FLUX", "DELTAH", "AERO_COND", "SHORTWAVE", "LONGWAVE", "DELSOILMOIST",
    "DELSWE", "DELINTERCEPT", "SNOW_SURF_TEMP", "SNOW_PACK_TEMP", "SALBEDO",
    "PET_SATSOIL", "PET_H2OSURF", "PET_SHORT", "PET_TALL", "PET_NATVEG", "PET_VEGNOCR", "ZWT", "ZWT_LUMPED",
    "DENSITY", "PRESSURE", "VP", "VPD", "QAIR",
]
OUT = {n: i for i, n in enumerate(OUT_NAMES)}
VR_NOUTVAR = len(OUT_NAMES)

_pd = C.POINTER(C.c_double)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)


class vr_options(C.Structure):
    _fields_ = [("nlayer", C.c_int32), ("nbands", C.c_int32), ("dt_hours", C.c_int32),
                ("snow_step_hours", C.c_int32), ("max_snow_temp", C.c_double),
                ("min_rain_temp", C.c_double), ("wind_h", C.c_double), ("n_out", C.c_int32),
                ("out_vars", C.c_int32 * VR_MAX_OUT), ("cells_per_block_hint", C.c_int32),
                ("moistfract", C.c_int32), ("reserved", C.c_int32 * 6)]


class vr_veglib(C.Structure):
    _fields_ = [("nclass", C.c_int32), ("overstory", _pi32), ("rarc", _pd), ("rmin", _pd),
                ("wind_h", _pd), ("RGL", _pd), ("rad_atten", _pd), ("wind_atten", _pd),
                ("trunk_ratio", _pd), ("roughness", _pd), ("displacement", _pd), ("LAI", _pd), ("albedo", _pd)]


CELL_SCALARS = ["lat", "elevation", "b_infilt", "Ds", "Dsmax", "Ws", "c_expt", "avg_temp", "dp",
                "rough", "snow_rough"]
CELL_LAYER = ["expt", "Ksat", "depth", "max_moist", "Wcr", "Wpwp", "resid_moist", "init_moist",
              "bulk_density", "soil_density", "bulk_dens_min", "soil_dens_min", "quartz", "organic",
              "porosity"]
CELL_BAND = ["AreaFract", "Tfactor", "Pfactor"]
TILE_F64 = ["tile_Cv", "tile_root", "tile_LAI", "tile_albedo", "tile_vegcover"]


class vr_cells(C.Structure):
    _fields_ = ([("ncell", C.c_int64)] + [(n, _pd) for n in CELL_SCALARS] +
                [(n, _pd) for n in CELL_LAYER] + [(n, _pd) for n in CELL_BAND] +
                [("tile_ptr", _pi64), ("tile_class", _pi32)] + [(n, _pd) for n in TILE_F64] +
                [("zwt_zwt", _pd), ("zwt_moist", _pd)])


class vr_forcing(C.Structure):
    _fields_ = [("nrec", C.c_int32), ("month", _pi32), ("day_in_year", _pi32),
                ("field", _pd * VR_NFORCING), ("snowflag", _pi32)]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, ty):
    return a.ctypes.data_as(ty)


def default_options(nlayer=3, nbands=1, dt_hours=24, out=None, max_snow_temp=0.5, min_rain_temp=-0.5,
                    wind_h=10.0, snow_step_hours=None, moistfract=False):
    """Reference defaults (initialize_global.c:146-211) and the default fluxes+snow output set
    (set_output_defaults.c:78-162): 19 + nlayer flux columns, 4 snow columns."""
    o = vr_options()
    o.nlayer, o.nbands, o.dt_hours, o.snow_step_hours = nlayer, nbands, dt_hours, snow_step_hours or dt_hours
    o.max_snow_temp, o.min_rain_temp, o.wind_h = max_snow_temp, min_rain_temp, wind_h
    o.moistfract = int(moistfract)
    if out is None:
        out = default_out_vars(nlayer)
    o.n_out = len(out)
    for i, v in enumerate(out):
        o.out_vars[i] = OUT[v] if isinstance(v, str) else int(v)
    return o


def default_out_vars(nlayer=3):
    names = ["PREC", "EVAP", "RUNOFF", "BASEFLOW", "WDEW"] + ["SOIL_LIQ%d" % l for l in range(nlayer)]
    names += ["NET_SHORT", "R_NET", "EVAP_CANOP", "TRANSP_VEG", "EVAP_BARE", "SUB_CANOP", "SUB_SNOW",
              "AERO_RESIST", "SURF_TEMP", "ALBEDO", "REL_HUMID", "IN_LONG", "AIR_TEMP", "WIND",
              "SWE", "SNOW_DEPTH", "SNOW_CANOPY", "SNOW_COVER"]
    return names


class Packed:
    """Holds a ctypes struct together with the numpy arrays its pointers refer to."""

    def __init__(self, struct, keep):
        self.struct = struct
        self.keep = keep

    def ref(self):
        return C.byref(self.struct)


def pack_veglib(lib):
    """lib: dict with nclass-length arrays (overstory int) and [nclass,12] roughness/displacement."""
    s = vr_veglib()
    keep = {}
    n = len(lib["overstory"])
    s.nclass = n
    keep["overstory"] = np.ascontiguousarray(lib["overstory"], dtype=np.int32)
    s.overstory = _ptr(keep["overstory"], _pi32)
    for k in ["rarc", "rmin", "wind_h", "RGL", "rad_atten", "wind_atten", "trunk_ratio", "roughness",
              "displacement"]:
        keep[k] = _f64(lib[k]).reshape(-1)
        setattr(s, k, _ptr(keep[k], _pd))
    for k in ["LAI", "albedo"]:                      # optional: only the PET outputs of the tile's own class read them
        if lib.get(k) is not None:
            keep[k] = _f64(lib[k]).reshape(-1)
            setattr(s, k, _ptr(keep[k], _pd))
    return Packed(s, keep)


def pack_cells(cells):
    """cells: dict of numpy arrays in the layout of include/vicres.h (per-layer [L,C], per-band
    [B,C], tiles CSR)."""
    s = vr_cells()
    keep = {}
    ncell = len(cells["lat"])
    s.ncell = ncell
    for k in CELL_SCALARS + CELL_LAYER + CELL_BAND + TILE_F64:
        keep[k] = _f64(cells[k]).reshape(-1)
        setattr(s, k, _ptr(keep[k], _pd))
    keep["tile_ptr"] = np.ascontiguousarray(cells["tile_ptr"], dtype=np.int64)
    keep["tile_class"] = np.ascontiguousarray(cells["tile_class"], dtype=np.int32)
    s.tile_ptr = _ptr(keep["tile_ptr"], _pi64)
    s.tile_class = _ptr(keep["tile_class"], _pi32)
    for k in ["zwt_zwt", "zwt_moist"]:               # optional [(L+2), 11, C] curves: only the ZWT outputs read them
        if cells.get(k) is not None:
            keep[k] = _f64(cells[k]).reshape(-1)
            setattr(s, k, _ptr(keep[k], _pd))
    return Packed(s, keep)


def pack_forcing(month, day_in_year, fields, snowflag=None):
    """fields: [9, nrec, C] float64 (or list of 9 [nrec, C] arrays) in FORCING_FIELDS order; with a sub-stepped snow
    model (NF > 1) [9, nrec * (NF + 1), C] and snowflag [nrec, C] int32."""
    s = vr_forcing()
    keep = {"month": np.ascontiguousarray(month, dtype=np.int32),
            "doy": np.ascontiguousarray(day_in_year, dtype=np.int32)}
    s.nrec = len(keep["month"])
    s.month = _ptr(keep["month"], _pi32)
    s.day_in_year = _ptr(keep["doy"], _pi32)
    for i in range(VR_NFORCING):
        keep["f%d" % i] = _f64(fields[i])
        s.field[i] = _ptr(keep["f%d" % i], _pd)
    if snowflag is not None:
        keep["snowflag"] = np.ascontiguousarray(snowflag, dtype=np.int32)
        s.snowflag = _ptr(keep["snowflag"], _pi32)
    return Packed(s, keep)
