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
