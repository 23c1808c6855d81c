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


# ---- reservoirs (MAKE_CONVOLUTIONRS) ---------------------------------------------------------------
_pf = C.POINTER(C.c_float)


class vr_reservoir(C.Structure):
    _fields_ = [("hresermax", C.c_float), ("hresermin", C.c_float), ("v_live", C.c_float), ("v_dead", C.c_float),
                ("head", C.c_float), ("q_design", C.c_float), ("ope_year", C.c_int32), ("v_initial", C.c_float),
                ("seepage", C.c_float), ("infiltration", C.c_float), ("hydropower_type", C.c_int32),
                ("hydropower_generator", C.c_int32), ("env_flow", C.c_float), ("rule", C.c_int32),
                ("hmax", C.c_float), ("hmin", C.c_float), ("op1", C.c_float * 2), ("reslv", C.c_float * 12),
                ("demand", C.c_float), ("x1", C.c_float), ("x2", C.c_float), ("x3", C.c_float), ("x4", C.c_float),
                ("demand5", C.c_float * 12), ("op5x1", C.c_float * 12), ("op5x2", C.c_float * 12),
                ("op5x3", C.c_float * 12), ("op5x4", C.c_float * 12), ("bathymetry", C.c_int32),
                ("bth", C.c_float * 4), ("series", _pf), ("irrigation", _pf)]


class vr_res_options(C.Structure):
    _fields_ = [("start_year", C.c_int32), ("start_month", C.c_int32), ("legacy_rule_curve", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class vr_res_series(C.Structure):
    _fields_ = [(n, _pf) for n in ("vol", "hho", "flowin", "flowout", "turbine", "energy", "htk")]


RES_SERIES = [("vol", 1), ("hho", 1), ("flowin", 0), ("flowout", 0), ("turbine", 0), ("energy", 0), ("htk", 0)]


def reservoir(**kw):
    """A parameter dict with the reference's defaults for everything not given (zeros: the Fortran arrays
    live in static storage).  Keys are the vr_reservoir field names; series/irrigation are arrays."""
    d = dict(hresermax=0.0, hresermin=0.0, v_live=0.0, v_dead=0.0, head=0.0, q_design=0.0, ope_year=0, v_initial=0.0,
             seepage=0.0, infiltration=0.0, hydropower_type=1, hydropower_generator=1, env_flow=0.0, rule=0,
             hmax=0.0, hmin=0.0, op1=[0.0, 0.0], reslv=[0.0] * 12, demand=0.0, x1=0.0, x2=0.0, x3=0.0, x4=0.0,
             demand5=[0.0] * 12, op5x1=[0.0] * 12, op5x2=[0.0] * 12, op5x3=[0.0] * 12, op5x4=[0.0] * 12,
             bathymetry=0, bth=[0.0] * 4, series=None, irrigation=None)
    for k in kw:
        if k not in d:
            raise KeyError(k)
    d.update(kw)
    return d


def pack_reservoirs(sets):
    """sets: list (parameter sets) of lists (reservoirs in node order) of dicts -> (ctypes array, keepalive)"""
    nsets, nres = len(sets), len(sets[0])
    arr = (vr_reservoir * (nsets * nres))()
    keep = []
    for s, rs in enumerate(sets):
        assert len(rs) == nres
        for j, d in enumerate(rs):
            r = arr[s * nres + j]
            for name, ty in vr_reservoir._fields_:
                v = d[name]
                if name in ("series", "irrigation"):
                    if v is not None:
                        a = np.ascontiguousarray(v, dtype=np.float32)
                        keep.append(a)
                        setattr(r, name, a.ctypes.data_as(_pf))
                elif hasattr(ty, "_length_"):
                    setattr(r, name, ty(*[float(x) for x in v]))
                else:
                    setattr(r, name, v)
    return arr, keep


def read_reservoir_file(path, start_year=None, start_month=None, ndays=None):
    """resN.txt as MAKE_CONVOLUTIONRS reads it (reservoir.f:1021-1131).  Accepts today's layout (HYDROPOWER
    type + generator lines) and the older one the shipped res_par/*.txt use (a single HYDROPOWER value =
    generator switch).  Series files of rules 4/6 and irrigation are aligned to (start_year, start_month, 1)."""
    L = [ln.rstrip("\n") for ln in open(path)]
    it = iter(L)
    nxt = lambda: next(it)
    nxt()
    f = nxt().split()
    d = reservoir(hresermax=float(f[0]), hresermin=float(f[1]), v_live=float(f[2]), v_dead=float(f[3]), head=float(f[4]),
                  q_design=float(f[5]), ope_year=int(float(f[6])), v_initial=float(f[7]))
    nxt()
    f = nxt().split()
    d["seepage"], d["infiltration"] = float(f[0]), float(f[1])
    nxt()
    with_irr = int(nxt().split()[0])
    ipath = nxt().strip()
    nxt()                                   # HYDROPOWER header
    a = nxt().split()[0]
    b = nxt().strip()
    try:                                    # today's layout: type, generator, header, env flow
        d["hydropower_type"], d["hydropower_generator"] = int(a), int(b.split()[0])
        nxt()
    except (ValueError, IndexError):        # older layout: one value, b was the ENVIRONMENTAL FLOW header
        d["hydropower_type"], d["hydropower_generator"] = 1, int(a)
    d["env_flow"] = float(nxt().split()[0])
    nxt()
    d["rule"] = int(nxt().split()[0])

    def series(fn):
        rows = [r.split() for r in open(fn).read().splitlines()[1:] if r.strip()]
        vals = np.array([float(r[3]) for r in rows], dtype=np.float32)
        start = 0
        for k, r in enumerate(rows):
            if (int(r[0]), int(r[1]), int(r[2])) == (start_year, start_month, 1):
                start = k                    # STARTDAY = L-1 (reservoir.f:1087-1089)
        out = np.zeros(ndays, dtype=np.float32)
        n = max(0, min(ndays, len(vals) - start))
        out[:n] = vals[start:start + n]
        return out
    if d["rule"] == 0:
        nxt()
    elif d["rule"] == 1:
        f = [float(x) for x in nxt().split()[:4]]
        d["hmax"], d["hmin"], d["op1"] = f[0], f[1], [f[2], f[3]]
    elif d["rule"] == 2:
        d["reslv"] = [float(x) for x in nxt().split()[:12]]
    elif d["rule"] == 3:
        f = [float(x) for x in nxt().split()[:5]]
        d["demand"], d["x1"], d["x2"], d["x3"], d["x4"] = f
    elif d["rule"] in (4, 6):
        d["series"] = series(nxt().strip())
    elif d["rule"] == 5:
        for m in range(12):
            f = [float(x) for x in nxt().split()[:5]]
            d["demand5"][m], d["op5x1"][m], d["op5x2"][m], d["op5x3"][m], d["op5x4"][m] = f
    nxt()
    d["bathymetry"] = int(nxt().split()[0])
    if d["bathymetry"] == 1:
        d["bth"] = [float(x) for x in nxt().split()[:4]]
    if with_irr == 1:
        rows = open(ipath).read().splitlines()
        n = int(rows[1].split()[0])
        irr = np.zeros(ndays, dtype=np.float32)
        v = np.array([float(r.split()[0]) for r in rows[3:3 + n]], dtype=np.float32)
        irr[:min(ndays, n)] = v[:ndays]
        d["irrigation"] = irr
    return d


class vr_route_inputs(C.Structure):
    _fields_ = [("runoff", _pf), ("baseflow", _pf), ("scale", C.c_float), ("cell_scale", _pf), ("ev_vege", _pf),
                ("ev_soil", _pf), ("tr_vege", _pf), ("cell_factor", _pf), ("present", C.POINTER(C.c_ubyte)),
                ("rh", _pf), ("tair", _pf), ("wind", _pf), ("cell_lat", _pf), ("cell_lon", _pf),
                ("node_ope_year", C.POINTER(C.c_int32)), ("start_year", C.c_int32), ("start_month", C.c_int32),
                ("node_begin", C.c_int32), ("node_end", C.c_int32), ("reserved", C.c_int32 * 2)]


def pack_route_inputs(runoff, baseflow, scale=0.18308, cell_scale=None, evap=None, cell_factor=None, present=None,
                      met=None, cell_lat=None, cell_lon=None, node_ope_year=None, start_year=0, start_month=1,
                      nodes=None):
    """-> (vr_route_inputs, keepalive, ndays).  evap = (ev_vege, ev_soil, tr_vege); met = (rh, tair, wind)"""
    s = vr_route_inputs()
    keep = []

    def f32(a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a.ctypes.data_as(_pf)
    runoff = np.ascontiguousarray(runoff, dtype=np.float32)
    s.runoff, s.baseflow, s.scale, s.cell_scale = f32(runoff), f32(baseflow), scale, f32(cell_scale)
    if evap is not None:
        s.ev_vege, s.ev_soil, s.tr_vege = f32(evap[0]), f32(evap[1]), f32(evap[2])
    s.cell_factor = f32(cell_factor)
    if present is not None:
        p = np.ascontiguousarray(present, dtype=np.uint8)
        keep.append(p)
        s.present = p.ctypes.data_as(C.POINTER(C.c_ubyte))
    if met is not None:
        s.rh, s.tair, s.wind = f32(met[0]), f32(met[1]), f32(met[2])
    s.cell_lat, s.cell_lon = f32(cell_lat), f32(cell_lon)
    if node_ope_year is not None:
        y = np.ascontiguousarray(node_ope_year, dtype=np.int32)
        keep.append(y)
        s.node_ope_year = y.ctypes.data_as(C.POINTER(C.c_int32))
    s.start_year, s.start_month = start_year, start_month
    if nodes is not None:
        s.node_begin, s.node_end = int(nodes[0]), int(nodes[1])
    return s, keep, runoff.shape[1]


def cell_coords(hdr, nrow, ncol):
    """JLOC, ILOC of every file-order cell in REAL arithmetic (reservoir.f:353-356)"""
    f32 = np.float32
    xc, yc, size = f32(float(hdr["xllcorner"])), f32(float(hdr["yllcorner"])), f32(float(hdr["cellsize"]))
    lat = np.zeros(nrow * ncol, dtype=np.float32)
    lon = np.zeros(nrow * ncol, dtype=np.float32)
    for r in range(nrow):
        for c in range(ncol):
            I, J = ncol - c, nrow - r
            lon[r * ncol + c] = xc + f32(ncol - I) * size + size / f32(2.0)
            lat[r * ncol + c] = yc + f32(J) * size - size / f32(2.0)
    return lat, lon
