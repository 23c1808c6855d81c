"""Reader for the named-array dumps written by oracle/ref_driver.c (TEST INFRASTRUCTURE ONLY).

Only tests/, tools/ (fixture generation), __graft_entry__.smoke() and bench.py's reference /
cpu_baseline legs may import this module.  Nothing here is on the product path.
"""
import struct
import numpy as np

_DT = {b"d": np.float64, b"i": np.int32, b"b": np.uint8}

# names of the per tile x band internals written with `-u` (ref_driver.c fill_unit order)
UNIT_FIELDS = (
    ["moist0", "moist1", "moist2", "evap0", "evap1", "evap2", "bef0", "bef1", "bef2",
     "runoff", "baseflow", "asat", "inflow", "aero_resist0", "aero_resist1",
     "Wdew", "canopyevap", "throughfall",
     "swq", "surf_temp", "pack_temp", "surf_water", "pack_water", "density", "depth", "albedo",
     "coldcontent", "coverage", "snow_canopy", "tmp_int_storage", "last_snow", "MELTING",
     "snow", "vapor_flux", "canopy_vapor_flux", "melt",
     "T0", "T1", "T2", "grnd_flux", "deltaH", "fusion", "snow_flux", "LongUnderOut", "Tfoliage",
     "AlbedoUnder", "NetShortAtmos", "NetLongAtmos"])
assert len(UNIT_FIELDS) == 48  # Tsurf is slot 48 -> dropped (NUNITF == 48)

ATMOS_FIELDS = ["air_temp", "prec", "wind", "vp", "vpd", "pressure", "density", "shortwave", "longwave"]
SOIL_SCALARS = ["lat", "lng", "elevation", "b_infilt", "Ds", "Dsmax", "Ws", "c", "avg_temp", "dp",
                "rough", "snow_rough", "annual_prec", "max_infil", "FS_ACTIVE", "frost_slope",
                "cell_area", "time_zone_lng"]
VEGLIB_SCALARS = ["overstory", "rarc", "rmin", "wind_h", "RGL", "rad_atten", "wind_atten"]


def read_dump(path):
    """Return {name: ndarray} for every record of a ref_driver dump."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    n = len(data)
    while off < n:
        name = data[off:off + 48].split(b"\0", 1)[0].decode()
        dt = data[off + 48:off + 49]
        (nd,) = struct.unpack_from("<i", data, off + 49)
        dims = struct.unpack_from("<%dq" % nd, data, off + 53)
        off += 53 + 8 * nd
        cnt = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(data, dtype=_DT[dt], count=cnt, offset=off).reshape(dims)
        off += arr.nbytes
        out[name] = arr
    return out


def out_index(dump):
    """Map OUT_* variable name -> (start column, nelem) in the per-cell `out` matrix."""
    names = [bytes(r).split(b"\0", 1)[0].decode() for r in dump["out_names"]]
    nelem = dump["out_nelem"]
    idx, col = {}, 0
    for nm, ne in zip(names, nelem):
        if nm:
            idx[nm] = (col, int(ne))
        col += int(ne)
    return idx


def cell_out(dump, icell, var):
    """[nrecs, nelem] array of one OUT_* variable for cell number `icell` of the dump."""
    col, ne = out_index(dump)[var]
    return dump["c%d/out" % icell][:, col:col + ne]


# ---------------------------------------------------------------------------------------------
# dump -> the input arrays of include/vicres.h (so the oracle port / the CUDA path can be handed
# exactly the numbers the reference consumed) and -> the reference's outputs in VR_OUT_* order
# ---------------------------------------------------------------------------------------------
_SOIL_IDX = {n: i for i, n in enumerate(SOIL_SCALARS)}

# VR_OUT_* name (vicres_b200/abi.py OUT_NAMES) -> (reference OUT_* name, element)
REF_OUT = {
    "PREC": ("OUT_PREC", 0), "EVAP": ("OUT_EVAP", 0), "RUNOFF": ("OUT_RUNOFF", 0),
    "BASEFLOW": ("OUT_BASEFLOW", 0), "WDEW": ("OUT_WDEW", 0), "SOIL_LIQ0": ("OUT_SOIL_LIQ", 0),
    "SOIL_LIQ1": ("OUT_SOIL_LIQ", 1), "SOIL_LIQ2": ("OUT_SOIL_LIQ", 2), "NET_SHORT": ("OUT_NET_SHORT", 0),
    "R_NET": ("OUT_R_NET", 0), "EVAP_CANOP": ("OUT_EVAP_CANOP", 0), "TRANSP_VEG": ("OUT_TRANSP_VEG", 0),
    "EVAP_BARE": ("OUT_EVAP_BARE", 0), "SUB_CANOP": ("OUT_SUB_CANOP", 0), "SUB_SNOW": ("OUT_SUB_SNOW", 0),
    "AERO_RESIST": ("OUT_AERO_RESIST", 0), "SURF_TEMP": ("OUT_SURF_TEMP", 0), "ALBEDO": ("OUT_ALBEDO", 0),
    "REL_HUMID": ("OUT_REL_HUMID", 0), "IN_LONG": ("OUT_IN_LONG", 0), "AIR_TEMP": ("OUT_AIR_TEMP", 0),
    "WIND": ("OUT_WIND", 0), "SWE": ("OUT_SWE", 0), "SNOW_DEPTH": ("OUT_SNOW_DEPTH", 0),
    "SNOW_CANOPY": ("OUT_SNOW_CANOPY", 0), "SNOW_COVER": ("OUT_SNOW_COVER", 0),
    "WATER_ERROR": ("OUT_WATER_ERROR", 0), "INFLOW": ("OUT_INFLOW", 0), "SNOW_MELT": ("OUT_SNOW_MELT", 0),
    "RAINF": ("OUT_RAINF", 0), "SNOWF": ("OUT_SNOWF", 0), "NET_LONG": ("OUT_NET_LONG", 0),
    "ASAT": ("OUT_ASAT", 0), "SOIL_WET": ("OUT_SOIL_WET", 0), "ROOTMOIST": ("OUT_ROOTMOIST", 0),
    "GRND_FLUX": ("OUT_GRND_FLUX", 0), "DELTAH": ("OUT_DELTAH", 0), "AERO_COND": ("OUT_AERO_COND", 0),
    "SHORTWAVE": ("OUT_SHORTWAVE", 0), "LONGWAVE": ("OUT_LONGWAVE", 0),
    "DELSOILMOIST": ("OUT_DELSOILMOIST", 0), "DELSWE": ("OUT_DELSWE", 0),
    "DELINTERCEPT": ("OUT_DELINTERCEPT", 0), "SNOW_SURF_TEMP": ("OUT_SNOW_SURF_TEMP", 0),
    "SNOW_PACK_TEMP": ("OUT_SNOW_PACK_TEMP", 0), "SALBEDO": ("OUT_SALBEDO", 0),
    "PET_SATSOIL": ("OUT_PET_SATSOIL", 0), "PET_H2OSURF": ("OUT_PET_H2OSURF", 0), "PET_SHORT": ("OUT_PET_SHORT", 0),
    "PET_TALL": ("OUT_PET_TALL", 0), "PET_NATVEG": ("OUT_PET_NATVEG", 0), "PET_VEGNOCR": ("OUT_PET_VEGNOCR", 0),
    "ZWT": ("OUT_ZWT", 0), "ZWT_LUMPED": ("OUT_ZWT_LUMPED", 0),
    "DENSITY": ("OUT_DENSITY", 0), "PRESSURE": ("OUT_PRESSURE", 0), "VP": ("OUT_VP", 0), "VPD": ("OUT_VPD", 0), "QAIR": ("OUT_QAIR", 0),
}


def case_from_dump(d):
    """Return dict(opt=..., lib=..., cells=..., month, doy, forcing [9,nrec,C], tair0 [C],
    ntile_max) built from a ref_driver dump `d`."""
    L, NB, nrec, C = int(d["Nlayer"][0]), int(d["Nbands"][0]), int(d["nrecs"][0]), int(d["ncells"][0])
    ncl = int(d["Nveglib"][0])
    vl = d["veglib"]
    lib = {
        "overstory": vl[:ncl, 0].astype(np.int32), "rarc": vl[:ncl, 1], "rmin": vl[:ncl, 2],
        "wind_h": vl[:ncl, 3], "RGL": vl[:ncl, 4], "rad_atten": vl[:ncl, 5], "wind_atten": vl[:ncl, 6],
        "trunk_ratio": d["veglib_trunk_ratio"][:ncl], "roughness": vl[:ncl, 31:43],
        "displacement": vl[:ncl, 43:55], "LAI": vl[:ncl, 7:19], "albedo": vl[:ncl, 19:31],
    }
    sc = np.stack([d["c%d/soil_scalars" % c] for c in range(C)])          # [C, 18]
    cells = {}
    for api, ref in [("lat", "lat"), ("elevation", "elevation"), ("b_infilt", "b_infilt"), ("Ds", "Ds"),
                     ("Dsmax", "Dsmax"), ("Ws", "Ws"), ("c_expt", "c"), ("avg_temp", "avg_temp"),
                     ("dp", "dp"), ("rough", "rough"), ("snow_rough", "snow_rough")]:
        cells[api] = sc[:, _SOIL_IDX[ref]].copy()
    for k in ["expt", "Ksat", "depth", "max_moist", "Wcr", "Wpwp", "resid_moist", "init_moist",
              "bulk_density", "soil_density", "bulk_dens_min", "soil_dens_min", "quartz", "organic",
              "porosity"]:
        cells[k] = np.stack([d["c%d/%s" % (c, k)] for c in range(C)], axis=1)      # [L, C]
    for k in ["AreaFract", "Tfactor", "Pfactor"]:
        cells[k] = np.stack([d["c%d/%s" % (c, k)] for c in range(C)], axis=1)      # [NB, C]
    # zwtvmoist curves: the dump holds [MAX_LAYERS+2, 11] per cell, curve k of the model is row k
    for k, nm in [("zwt_zwt", "zwtvmoist_zwt"), ("zwt_moist", "zwtvmoist_moist")]:
        cells[k] = np.stack([d["c%d/%s" % (c, nm)][:L + 2] for c in range(C)], axis=2)  # [L+2, 11, C]
    tile_ptr = [0]
    t_class, t_cv, t_root, t_lai, t_alb, t_vc = [], [], [], [], [], []
    ntile_max = 0
    for c in range(C):
        nveg = int(d["c%d/Nveg" % c][0])
        cls, cv = d["c%d/veg_class" % c], d["c%d/Cv" % c]
        nt = 0
        for v in range(nveg + 1):
            if v == nveg and not cv[v] > 0:
                continue                      # no bare-soil tile for this cell
            bare = v == nveg
            t_class.append(ncl if bare else int(cls[v]))
            t_cv.append(cv[v])
            t_root.append(np.zeros(L) if bare else d["c%d/root" % c][v])
            if bare:
                t_lai.append(np.zeros(12)); t_alb.append(np.zeros(12)); t_vc.append(np.zeros(12))
            else:
                vm = d["c%d/veg_monthly" % c][v]
                t_lai.append(vm[0]); t_alb.append(vm[1]); t_vc.append(vm[2])
            nt += 1
        ntile_max = max(ntile_max, nveg + 1)
        tile_ptr.append(tile_ptr[-1] + nt)
    cells["tile_ptr"] = np.array(tile_ptr, dtype=np.int64)
    cells["tile_class"] = np.array(t_class, dtype=np.int32)
    cells["tile_Cv"] = np.array(t_cv)
    cells["tile_root"] = np.array(t_root).reshape(-1, L)
    cells["tile_LAI"] = np.array(t_lai).reshape(-1, 12)
    cells["tile_albedo"] = np.array(t_alb).reshape(-1, 12)
    cells["tile_vegcover"] = np.array(t_vc).reshape(-1, 12)
    NF = int(d["NF"][0])
    if NF == 1:
        forcing = np.stack([d["c%d/atmos" % c][:, :, 0] for c in range(C)], axis=2)   # [nrec, 9, C]
        forcing = np.ascontiguousarray(forcing.transpose(1, 0, 2))                    # [9, nrec, C]
        snowflag = None
    else:
        # sub-stepped snow model: per record the NF sub-step values, then the record's own (index NR == NF):
        # [nrec, 9, NF+1] per cell -> [9, nrec * (NF+1), C], the layout of vr_forcing.field[]
        a = np.stack([d["c%d/atmos" % c] for c in range(C)], axis=3)                  # [nrec, 9, NF+1, C]
        forcing = np.ascontiguousarray(a.transpose(1, 0, 2, 3).reshape(9, nrec * (NF + 1), C))
        snowflag = np.ascontiguousarray(np.stack([d["c%d/snowflag" % c][:, NF] for c in range(C)], axis=1).astype(np.int32))
    dmy = d["dmy"]
    return {
        "nlayer": L, "nbands": NB, "nrec": nrec, "ncell": C, "dt": int(d["dt"][0]),
        "snow_step": int(d["snow_step"][0]), "NF": int(d["NF"][0]),
        "max_snow_temp": float(d["MAX_SNOW_TEMP"][0]), "min_rain_temp": float(d["MIN_RAIN_TEMP"][0]),
        "wind_h": float(d["wind_h"][0]),
        "lib": lib, "cells": cells, "month": dmy[:, 1].astype(np.int32), "doy": dmy[:, 4].astype(np.int32),
        "forcing": forcing, "snowflag": snowflag,
        "tair0": forcing[0, NF if NF > 1 else 0, :].copy(),       # atmos[0].air_temp[NR]
        "ntile_max": ntile_max,
    }


def ref_outputs(d, names):
    """[len(names), nrec, C] reference outputs (put_data's out_data[].data after each step)."""
    C = int(d["ncells"][0])
    idx = out_index(d)
    res = []
    for nm in names:
        rn, el = REF_OUT[nm]
        col, ne = idx[rn]
        res.append(np.stack([d["c%d/out" % c][:, col + el] for c in range(C)], axis=1))
    return np.stack(res)
