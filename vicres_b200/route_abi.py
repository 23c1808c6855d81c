"""ctypes mirror of include/vicres_route.h and readers for the reference's routing input files
(host glue only; formats: routing/model/reservoir.f:149-211 READ_DIREC, init_routines.f:62-98
READ_RESE, rout.f:333 stations.txt, reservoir.f:53-77 UH.all)."""
import ctypes as C
import numpy as np

LE, UH_DAY, KE = 48, 96, 12
NS = KE + UH_DAY - 1


class vr_route_grid(C.Structure):
    _fields_ = [("ncol", C.c_int32), ("nrow", C.c_int32), ("dir", C.POINTER(C.c_int32)),
                ("reser", C.POINTER(C.c_int32)), ("velo", C.POINTER(C.c_float)),
                ("diff", C.POINTER(C.c_float)), ("xmask", C.POINTER(C.c_float))]


class PackedGrid:
    def __init__(self, s, keep):
        self.struct, self.keep = s, keep

    def ref(self):
        return C.byref(self.struct)


def pack_grid(grid):
    """grid: dict(dir [nrow,ncol] int, reser [nrow,ncol] int or None, velo/diff/xmask scalar or [nrow,ncol])"""
    d = np.ascontiguousarray(grid["dir"], dtype=np.int32)
    nrow, ncol = d.shape
    keep = {"dir": d}
    s = vr_route_grid()
    s.ncol, s.nrow = ncol, nrow
    s.dir = d.ctypes.data_as(C.POINTER(C.c_int32))
    if grid.get("reser") is not None:
        keep["reser"] = np.ascontiguousarray(grid["reser"], dtype=np.int32)
        s.reser = keep["reser"].ctypes.data_as(C.POINTER(C.c_int32))
    for k in ("velo", "diff", "xmask"):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(grid[k], dtype=np.float32), (nrow, ncol)))
        keep[k] = a
        setattr(s, k, a.ctypes.data_as(C.POINTER(C.c_float)))
    return PackedGrid(s, keep)


def read_esri_int(path, header=6):
    """flowdirection.txt: 6 header lines then nrow lines; reservoirlocation.txt: no header."""
    with open(path) as f:
        lines = [ln.split() for ln in f.read().splitlines() if ln.strip()]
    hdr = {ln[0].lower(): ln[1] for ln in lines[:header]}
    data = np.array([[int(float(x)) for x in ln] for ln in lines[header:]], dtype=np.int32)
    return hdr, data


def cell_latlon(hdr, nrow, ncol, r, c):
    """centre of file-order cell (r, c): the reference's JLOC/ILOC (reservoir.f:353-356)"""
    xc, yc, size = float(hdr["xllcorner"]), float(hdr["yllcorner"]), float(hdr["cellsize"])
    I, J = ncol - c, nrow - r
    return yc + J * size - size / 2.0, xc + (ncol - I) * size + size / 2.0


def cell_factors(hdr, nrow, ncol, fraction=1.0):
    """FACTOR of every file-order grid cell in fp32, as MAKE_CONVOLUTION computes it
    (reservoir.f:357-364): FRACTION*35.315*AREA/(86400*1000), AREA from the cell's latitude band."""
    f32 = np.float32
    size = f32(float(hdr["cellsize"]))
    PI = f32(np.arctan(f32(1.0)) * f32(4.0))
    RERD = f32(6371229.0)
    out = np.zeros(nrow * ncol, dtype=np.float32)
    frac = np.broadcast_to(np.asarray(fraction, dtype=np.float32), (nrow, ncol))
    for r in range(nrow):
        for c in range(ncol):
            jloc = f32(float(hdr["yllcorner"])) + f32(nrow - r) * size - size / f32(2.0)     # JLOC in REAL arithmetic
            area = RERD ** 2 * abs(size) * PI / f32(180) * abs(np.sin((jloc - size / f32(2.0)) * PI / f32(180)) -
                                                               np.sin((jloc + size / f32(2.0)) * PI / f32(180)))
            out[r * ncol + c] = frac[r, c] * f32(35.315) * f32(area) / (f32(86400.0) * f32(1000.0))
    return out
