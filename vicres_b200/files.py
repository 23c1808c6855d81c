"""ctypes binding of include/vicres_files.h: the reference's parameter files -> the struct-of-arrays
hand-off of include/vicres.h, and the ASCII flux writer.  Host only (no GPU needed).

Mirrors read_soilparam / read_veglib / read_vegparam / read_snowband (runoff/model/*.c) and
write_data.c's ASCII branch; the parsing itself is in csrc/vicres_files.cpp.
"""
import ctypes as C
import numpy as np
from . import abi
from .model import load_library, VicResError


class ParamFilesC(C.Structure):
    _fields_ = [("nlayer", C.c_int32), ("nbands", C.c_int32), ("root_zones", C.c_int32),
                ("vegparam_lai", C.c_int32), ("lai_from_vegparam", C.c_int32),
                ("soil", C.c_char_p), ("veglib", C.c_char_p), ("vegparam", C.c_char_p),
                ("snowband", C.c_char_p), ("baseflow_nijssen", C.c_int32), ("organic_fract", C.c_int32)]


FILES_EXPORTS = ["vr_params_read", "vr_params_free", "vr_params_last_error", "vr_params_cells",
                 "vr_params_veglib", "vr_params_gridcel", "vr_params_lat", "vr_params_lng", "vr_params_site", "vr_write_results", "vr_round_decimals"]


def _bind(lib):
    lib.vr_params_read.argtypes = [C.POINTER(ParamFilesC), C.POINTER(C.c_void_p)]
    lib.vr_params_read.restype = C.c_int
    lib.vr_params_free.argtypes = [C.c_void_p]
    lib.vr_params_free.restype = None
    lib.vr_params_last_error.restype = C.c_char_p
    lib.vr_params_cells.argtypes = [C.c_void_p]
    lib.vr_params_cells.restype = C.POINTER(abi.vr_cells)
    lib.vr_params_veglib.argtypes = [C.c_void_p]
    lib.vr_params_veglib.restype = C.POINTER(abi.vr_veglib)
    for n, t in [("vr_params_gridcel", C.c_int32), ("vr_params_lat", C.c_float), ("vr_params_lng", C.c_float)]:
        getattr(lib, n).argtypes = [C.c_void_p]
        getattr(lib, n).restype = C.POINTER(t)
    lib.vr_params_site.argtypes = [C.c_void_p, C.c_int32]
    lib.vr_params_site.restype = C.POINTER(C.c_double)
    lib.vr_write_results.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int64, C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_double)]
    lib.vr_write_results.restype = C.c_int
    lib.vr_round_decimals.argtypes = [C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32]
    lib.vr_round_decimals.restype = C.c_int
    return lib


def _arr(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)


def read_params(soil, veglib, vegparam, snowband=None, nlayer=3, nbands=1, root_zones=3,
                vegparam_lai=False, lai_from_vegparam=False, lib=None, baseflow_nijssen=False, organic_fract=False):
    """Parse the four parameter files; returns (lib_dict, cells_dict, meta) in the layout
    abi.pack_veglib / abi.pack_cells take.  meta: gridcel, lat, lng (float32 as the reference stores)."""
    so = _bind(lib or load_library())
    pf = ParamFilesC(nlayer, nbands, root_zones, int(vegparam_lai), int(lai_from_vegparam),
                     soil.encode(), veglib.encode(), vegparam.encode(), snowband.encode() if snowband else None, int(baseflow_nijssen), int(organic_fract))
    h = C.c_void_p()
    rc = so.vr_params_read(C.byref(pf), C.byref(h))
    if rc != 0:
        raise VicResError(rc, "vr_params_read: %s" % so.vr_params_last_error().decode())
    try:
        cs, vl = so.vr_params_cells(h).contents, so.vr_params_veglib(h).contents
        Cn, L, NB, ncl = int(cs.ncell), nlayer, nbands, int(vl.nclass)
        cells = {k: _arr(getattr(cs, k), Cn, np.float64) for k in abi.CELL_SCALARS}
        for k in abi.CELL_LAYER:
            cells[k] = _arr(getattr(cs, k), L * Cn, np.float64).reshape(L, Cn)
        for k in abi.CELL_BAND:
            cells[k] = _arr(getattr(cs, k), NB * Cn, np.float64).reshape(NB, Cn)
        cells["tile_ptr"] = _arr(cs.tile_ptr, Cn + 1, np.int64)
        T = int(cells["tile_ptr"][-1])
        cells["tile_class"] = _arr(cs.tile_class, T, np.int32)
        cells["tile_Cv"] = _arr(cs.tile_Cv, T, np.float64)
        cells["tile_root"] = _arr(cs.tile_root, T * L, np.float64).reshape(T, L)
        for k in ["tile_LAI", "tile_albedo", "tile_vegcover"]:
            cells[k] = _arr(getattr(cs, k), T * 12, np.float64).reshape(T, 12)
        for k in ["zwt_zwt", "zwt_moist"]:
            cells[k] = _arr(getattr(cs, k), (L + 2) * 11 * Cn, np.float64).reshape(L + 2, 11, Cn)
        libd = {"overstory": _arr(vl.overstory, ncl, np.int32)}
        for k in ["rarc", "rmin", "wind_h", "RGL", "rad_atten", "wind_atten", "trunk_ratio"]:
            libd[k] = _arr(getattr(vl, k), ncl, np.float64)
        for k in ["roughness", "displacement", "LAI", "albedo"]:
            libd[k] = _arr(getattr(vl, k), ncl * 12, np.float64).reshape(ncl, 12)
        meta = {"gridcel": _arr(so.vr_params_gridcel(h), Cn, np.int32),
                "lat": _arr(so.vr_params_lat(h), Cn, np.float32), "lng": _arr(so.vr_params_lng(h), Cn, np.float32),
                # what initialize_atmos reads of soil_con: the vr_atmos_cells hand-off (vicres_b200/atmos.py)
                "site": {n: _arr(so.vr_params_site(h, i), Cn, np.float64)
                         for i, n in enumerate(["lat", "lng", "time_zone_lng", "elevation", "annual_prec", "min_Tfactor"])}}
    finally:
        so.vr_params_free(h)
    return libd, cells, meta


def write_results(result_dir, prefix, lat, lng, year, month, day, out, cols=None, hour=None, grid_decimal=4,
                  lib=None):
    """out: [n_out, nrec, C] float64 as VicRes.step returns; one `<prefix>_<lat>_<lng>` file per cell
    holding the rows `cols` of out (default all).  hour: per record when dt < 24, else None."""
    so = _bind(lib or load_library())
    out = np.ascontiguousarray(out, np.float64)
    n_out, nrec, Cn = out.shape
    cols = np.arange(n_out, dtype=np.int32) if cols is None else np.ascontiguousarray(cols, np.int32)
    lat, lng = np.ascontiguousarray(lat, np.float32), np.ascontiguousarray(lng, np.float32)
    y, m, d = (np.ascontiguousarray(a, np.int32) for a in (year, month, day))
    f32, i32 = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    hp = None
    if hour is not None:
        hour = np.ascontiguousarray(hour, np.int32)
        hp = hour.ctypes.data_as(i32)
    rc = so.vr_write_results(result_dir.encode(), prefix.encode(), grid_decimal, Cn, lat.ctypes.data_as(f32),
                             lng.ctypes.data_as(f32), nrec, y.ctypes.data_as(i32), m.ctypes.data_as(i32),
                             d.ctypes.data_as(i32), hp, len(cols), cols.ctypes.data_as(i32),
                             out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise VicResError(rc, "vr_write_results: %s" % so.vr_params_last_error().decode())


def round_decimals(a, decimals=4, lib=None):
    """values as they come back from a "%.<decimals>f" text column (printf rounding + strtod), without text"""
    so = _bind(lib or load_library())
    a = np.ascontiguousarray(a, np.float64)
    out = np.empty_like(a)
    pd = C.POINTER(C.c_double)
    rc = so.vr_round_decimals(a.size, a.ctypes.data_as(pd), out.ctypes.data_as(pd), decimals)
    if rc != 0:
        raise VicResError(rc, "vr_round_decimals")
    return out


# ---- global parameter file, calendar, forcing files ---------------------------------------------------------
VR_MAX_FORCE_TYPES = 16


class GlobalC(C.Structure):
    _fields_ = ([("nlayer", C.c_int32), ("time_step", C.c_int32), ("snow_step", C.c_int32),
                 ("max_snow_temp", C.c_double), ("min_rain_temp", C.c_double), ("wind_h", C.c_double),
                 ("min_wind_speed", C.c_double), ("measure_h", C.c_double)] +
                [(n, C.c_int32) for n in ("startyear", "startmonth", "startday", "starthour", "nrecs", "skipyear",
                                          "snow_band", "root_zones", "vegparam_lai", "lai_from_vegparam", "grid_decimal")] +
                [(n, C.c_char * 512) for n in ("soil", "veglib", "vegparam", "snowband", "result_dir", "forcing1")] +
                [(n, C.c_int32) for n in ("force_ascii", "force_little_endian", "force_dt", "n_force_types", "forceyear",
                                          "forcemonth", "forceday", "forcehour")] +
                [("force_type", (C.c_char * 16) * VR_MAX_FORCE_TYPES), ("force_signed", C.c_int32 * VR_MAX_FORCE_TYPES),
                 ("force_multiplier", C.c_double * VR_MAX_FORCE_TYPES)] +
                [(n, C.c_int32) for n in ("binary_output", "out_step", "compress", "full_energy", "frozen_soil", "lakes",
                                          "carbon", "blowing", "corrprec", "compute_treeline", "dist_prcp", "organic_fract",
                                          "july_tavg_supplied", "alma_input", "init_state",
                                          "vp_interp", "vp_iter", "mtclim_swe_corr", "lw_type", "lw_cloud", "plapse")] +
                [("sw_prec_thresh", C.c_double), ("init_state_file", C.c_char * 512), ("statename", C.c_char * 512)] +
                [(n, C.c_int32) for n in ("stateyear", "statemonth", "stateday", "binary_state_file",
                                          "forcing2", "n_outfile_lines", "n_outvar_lines")] +
                [("other_unsupported", C.c_char * 96), ("prt_header", C.c_int32), ("alma_output", C.c_int32),
                 ("output_force", C.c_int32), ("moistfract", C.c_int32),
                 ("baseflow_nijssen", C.c_int32)])


FILES_EXPORTS += ["vr_global_read", "vr_global_read_forcing2", "vr_global_unsupported", "vr_make_dmy", "vr_force_skip", "vr_output_skip",
                  "vr_read_forcing_file"]


def _bind_global(so):
    so.vr_global_read.argtypes = [C.c_char_p, C.POINTER(GlobalC)]
    so.vr_global_read_forcing2.argtypes = [C.c_char_p, C.POINTER(GlobalC)]
    so.vr_global_unsupported.argtypes = [C.POINTER(GlobalC)]
    so.vr_global_unsupported.restype = C.c_char_p
    so.vr_make_dmy.argtypes = [C.POINTER(GlobalC)] + [C.POINTER(C.c_int32)] * 5
    so.vr_force_skip.argtypes = [C.POINTER(GlobalC)]
    so.vr_output_skip.argtypes = [C.POINTER(GlobalC), C.POINTER(C.c_int32)]
    so.vr_read_forcing_file.argtypes = [C.POINTER(GlobalC), C.c_char_p, C.c_int64, C.POINTER(C.c_double)]
    so.vr_read_forcing_file.restype = C.c_int64
    return so


class Global:
    """a parsed global parameter file (get_global_param.c) with its calendar (make_dmy.c)"""

    def __init__(self, path, lib=None, second=False):
        """second: the description of FORCING2 in the forcing fields (vr_global_read_forcing2)"""
        self.so = _bind_global(_bind(lib or load_library()))
        self.c = GlobalC()
        self.path = path
        rc = (self.so.vr_global_read_forcing2 if second else self.so.vr_global_read)(path.encode(), C.byref(self.c))
        if rc != 0:
            raise VicResError(rc, "vr_global_read: %s" % self.so.vr_params_last_error().decode())

    def second_forcing(self):
        """the same global file seen from its second forcing file, or None"""
        if not self.c.forcing2:
            return None
        g2 = Global(self.path, second=True)
        if g2.force_dt != self.force_dt:
            raise VicResError(abi.VR_ERR_UNSUPPORTED, "FORCING2 with a FORCE_DT other than FORCING1's (%d, %d)" % (g2.force_dt, self.force_dt))
        return g2

    def __getattr__(self, k):
        v = getattr(self.c, k)
        return v.decode() if isinstance(v, bytes) else v

    @property
    def unsupported(self):
        r = self.so.vr_global_unsupported(C.byref(self.c))
        return r.decode() if r else None

    @property
    def force_types(self):
        return [self.c.force_type[i].value.decode() for i in range(self.c.n_force_types)]

    def dmy(self):
        """[nrecs, 5] year, month, day, hour, day_in_year"""
        n = self.c.nrecs
        a = [np.zeros(n, dtype=np.int32) for _ in range(5)]
        p = C.POINTER(C.c_int32)
        rc = self.so.vr_make_dmy(C.byref(self.c), *[x.ctypes.data_as(p) for x in a])
        if rc != 0:
            raise VicResError(rc, "vr_make_dmy")
        return np.stack(a, axis=1)

    def output_skip(self, dmy):
        """SKIPYEAR as a count of model records (make_dmy.c:162-169); dmy from self.dmy()"""
        y = np.ascontiguousarray(dmy[:, 0], np.int32)
        return int(self.so.vr_output_skip(C.byref(self.c), y.ctypes.data_as(C.POINTER(C.c_int32))))

    def param_files(self):
        """keyword arguments for read_params()"""
        return dict(soil=self.soil, veglib=self.veglib, vegparam=self.vegparam, snowband=self.snowband or None,
                    nlayer=self.nlayer, nbands=self.snow_band, root_zones=self.root_zones,
                    vegparam_lai=bool(self.vegparam_lai), lai_from_vegparam=bool(self.lai_from_vegparam),
                    baseflow_nijssen=bool(self.baseflow_nijssen), organic_fract=bool(self.organic_fract))

    def read_forcing(self, lat, lng):
        """the columns of the cell's forcing file, {FORCE_TYPE: [nrec_forcing]} (no disaggregation)"""
        nrec = self.c.nrecs * self.c.time_step // self.c.force_dt
        data = np.zeros((self.c.n_force_types, nrec), dtype=np.float64)
        path = "%s%.*f_%.*f" % (self.forcing1, self.grid_decimal, lat, self.grid_decimal, lng)
        got = self.so.vr_read_forcing_file(C.byref(self.c), path.encode(), nrec, data.ctypes.data_as(C.POINTER(C.c_double)))
        if got < 0:
            raise VicResError(int(got), "vr_read_forcing_file: %s" % self.so.vr_params_last_error().decode())
        return {t: data[i, :got] for i, t in enumerate(self.force_types)}, int(got)


# ---- output selection (N_OUTFILES / OUTFILE / OUTVAR), aggregation to OUT_STEP, ASCII / binary result files -----------
VR_MAX_OUTFILES, VR_MAX_OUTCOLS = 8, 64
OUT_TYPES = ["DEFAULT", "CHAR", "SINT", "USINT", "INT", "FLOAT", "DOUBLE"]
OUT_TYPE_DTYPE = {"CHAR": "i1", "SINT": "<i2", "USINT": "<u2", "INT": "<i4", "FLOAT": "<f4", "DOUBLE": "<f8"}


class OutSelC(C.Structure):
    _fields_ = [("nfiles", C.c_int32), ("prefix", (C.c_char * 64) * VR_MAX_OUTFILES), ("ncols", C.c_int32 * VR_MAX_OUTFILES),
                ("var", (C.c_int32 * VR_MAX_OUTCOLS) * VR_MAX_OUTFILES),
                ("format", ((C.c_char * 12) * VR_MAX_OUTCOLS) * VR_MAX_OUTFILES),
                ("type", (C.c_int32 * VR_MAX_OUTCOLS) * VR_MAX_OUTFILES),
                ("mult", (C.c_double * VR_MAX_OUTCOLS) * VR_MAX_OUTFILES),
                ("name", ((C.c_char * 32) * VR_MAX_OUTCOLS) * VR_MAX_OUTFILES), ("nvars", C.c_int32 * VR_MAX_OUTFILES),
                ("header", C.c_int32), ("hdr_nrecs", C.c_int32), ("hdr_out_dt", C.c_int32), ("hdr_start", C.c_int32 * 4),
                ("hdr_alma", C.c_int32), ("no_date", C.c_int32)]


FILES_EXPORTS += ["vr_outsel_read", "vr_outsel_step_vars", "vr_aggregate_device", "vr_aggregate", "vr_write_results_sel",
                  "vr_alma_output_device", "vr_alma_output", "vr_forcing_outputs_device", "vr_forcing_outputs"]


def _bind_outsel(so):
    i32, f32, pd = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)
    so.vr_outsel_read.argtypes = [C.c_char_p, C.c_int32, C.POINTER(OutSelC)]
    so.vr_outsel_step_vars.argtypes = [C.POINTER(OutSelC), C.c_int32, i32, i32]
    so.vr_aggregate_device.argtypes = [C.c_int, C.c_int32, i32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    so.vr_aggregate.argtypes = [C.c_int, C.c_int32, i32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, pd, pd]
    so.vr_forcing_outputs_device.argtypes = [C.c_int, C.c_int32, i32, C.c_int32, C.c_int32, C.c_int64, C.c_double, C.c_double,
                                             C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    so.vr_alma_output_device.argtypes = [C.c_int, C.c_int32, i32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    so.vr_write_results_sel.argtypes = [C.c_char_p, C.POINTER(OutSelC), C.c_int32, C.c_int32, C.c_int64, f32, f32, C.c_int32,
                                        i32, i32, i32, i32, C.c_int32, i32, pd]
    return so


class OutSel:
    """the output files of a run and their columns (parse_output_info.c / set_output_defaults.c)"""

    def __init__(self, global_path, nlayer=3, lib=None):
        self.so = _bind_outsel(_bind(lib or load_library()))
        self.c = OutSelC()
        rc = self.so.vr_outsel_read(global_path.encode(), nlayer, C.byref(self.c))
        if rc != 0:
            raise VicResError(rc, "vr_outsel_read: %s" % self.so.vr_params_last_error().decode())

    @property
    def files(self):
        """[(prefix, [(variable name, format, type name, multiplier)])]"""
        out = []
        for f in range(self.c.nfiles):
            cols = [(abi.OUT_NAMES[self.c.var[f][k]], self.c.format[f][k].value.decode(), OUT_TYPES[self.c.type[f][k]],
                     float(self.c.mult[f][k])) for k in range(self.c.ncols[f])]
            out.append((self.c.prefix[f].value.decode(), cols))
        return out

    def step_vars(self, aggregate=False, alma=False):
        """names of the outputs the step has to produce (each once); AERO_RESIST becomes AERO_COND when aggregating, and
        SUB_CANOP joins SUB_SNOW under ALMA_OUTPUT"""
        v = (C.c_int32 * abi.VR_MAX_OUT)()
        n = C.c_int32()
        rc = self.so.vr_outsel_step_vars(C.byref(self.c), int(bool(aggregate)) | (2 if alma else 0), v, C.byref(n))
        if rc != 0:
            raise VicResError(rc, "vr_outsel_step_vars: %s" % self.so.vr_params_last_error().decode())
        return [abi.OUT_NAMES[v[k]] for k in range(n.value)]

    def rows(self, f, names):
        """row of the output array (variables `names`) that holds each column of file f"""
        return np.asarray([names.index(abi.OUT_NAMES[self.c.var[f][k]]) for k in range(self.c.ncols[f])], dtype=np.int32)

    def wants(self, name):
        return any(abi.OUT_NAMES[self.c.var[f][k]] == name for f in range(self.c.nfiles) for k in range(self.c.ncols[f]))

    def set_header(self, nrecs, out_dt, start, alma=False):
        """PRT_HEADER: the run's record count, OUT_STEP (hours), the first model record's (year, month, day, hour)"""
        self.c.header, self.c.hdr_nrecs, self.c.hdr_out_dt, self.c.hdr_alma = 1, int(nrecs), int(out_dt), int(alma)
        for k in range(4):
            self.c.hdr_start[k] = int(start[k])

    def write(self, f, result_dir, lat, lng, year, month, day, hour, out, rows, binary=False, grid_decimal=4):
        out = np.ascontiguousarray(out, np.float64)
        _, nrec, Cn = out.shape
        lat, lng = np.ascontiguousarray(lat, np.float32), np.ascontiguousarray(lng, np.float32)
        y, m, d = (np.ascontiguousarray(a, np.int32) for a in (year, month, day))
        rows = np.ascontiguousarray(rows, np.int32)
        f32, i32 = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        hp = None
        if hour is not None:
            hour = np.ascontiguousarray(hour, np.int32)
            hp = hour.ctypes.data_as(i32)
        rc = self.so.vr_write_results_sel(result_dir.encode(), C.byref(self.c), f, grid_decimal, Cn, lat.ctypes.data_as(f32),
                                          lng.ctypes.data_as(f32), nrec, y.ctypes.data_as(i32), m.ctypes.data_as(i32),
                                          d.ctypes.data_as(i32), hp, int(binary), rows.ctypes.data_as(i32),
                                          out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise VicResError(rc, "vr_write_results_sel: %s" % self.so.vr_params_last_error().decode())


def aggregate(names, data, ratio, device=0, lib=None):
    """put_data.c:665-688 for host arrays: data [n, nrec, C] -> [n, nrec // ratio, C]; names: the variables of the rows, with
    "AERO_RESIST" meaning "this row holds AERO_COND, give back the resistance"."""
    so = _bind_outsel(_bind(lib or load_library()))
    data = np.ascontiguousarray(data, np.float64)
    n, nrec, Cn = data.shape
    nout = nrec // ratio
    data = np.ascontiguousarray(data[:, :nout * ratio])
    out = np.empty((n, nout, Cn))
    v = np.asarray([abi.OUT[x] for x in names], dtype=np.int32)
    pd = C.POINTER(C.c_double)
    rc = so.vr_aggregate(device, n, v.ctypes.data_as(C.POINTER(C.c_int32)), 1, nout * ratio, ratio, Cn, data.ctypes.data_as(pd),
                         out.ctypes.data_as(pd))
    if rc != 0:
        raise VicResError(rc, "vr_aggregate")
    return out


def aggregate_device(names, in_ptr, out_ptr, nrec, ratio, ncell, device=0, stream=0, lib=None):
    """the same on device buffers: in [n][nrec][ncell] -> out [n][nrec / ratio][ncell]; nrec must be a multiple of ratio"""
    so = _bind_outsel(_bind(lib or load_library()))
    v = np.asarray([abi.OUT[x] for x in names], dtype=np.int32)
    rc = so.vr_aggregate_device(device, len(names), v.ctypes.data_as(C.POINTER(C.c_int32)), 1, nrec, ratio, ncell, in_ptr, out_ptr,
                                stream or None)
    if rc != 0:
        raise VicResError(rc, "vr_aggregate_device")


def forcing_outputs_device(names, field_ptrs, out_ptr, nrecs, nf, ncell, max_snow_temp, min_rain_temp, device=0, stream=0, lib=None):
    """OUTPUT_FORCE rows (write_forcing_file.c): field_ptrs = the nine prepared forcing arrays on the device, out [n][nrecs * nf][ncell]"""
    so = _bind_outsel(_bind(lib or load_library()))
    v = np.asarray([abi.OUT[x] for x in names], dtype=np.int32)
    fp = (C.c_void_p * abi.VR_NFORCING)(*[int(p) for p in field_ptrs])
    rc = so.vr_forcing_outputs_device(device, len(names), v.ctypes.data_as(C.POINTER(C.c_int32)), nrecs, nf, ncell, max_snow_temp,
                                      min_rain_temp, fp, out_ptr, stream or None)
    if rc != 0:
        raise VicResError(rc, "vr_forcing_outputs_device: %s" % ("a variable that is not a forcing output" if rc == abi.VR_ERR_UNSUPPORTED
                                                                 else "failed"))


def alma_output_device(names, data_ptr, nrec, ncell, out_dt, device=0, stream=0, lib=None):
    """ALMA_OUTPUT units in place on device values [n][nrec][ncell] that are already at OUT_STEP = out_dt hours"""
    so = _bind_outsel(_bind(lib or load_library()))
    v = np.asarray([abi.OUT[x] for x in names], dtype=np.int32)
    rc = so.vr_alma_output_device(device, len(names), v.ctypes.data_as(C.POINTER(C.c_int32)), nrec, ncell, out_dt, data_ptr, stream or None)
    if rc != 0:
        raise VicResError(rc, "vr_alma_output_device")


def read_binary_results(path, cols, hourly, header=False, no_date=False):
    """a BINARY_OUTPUT result file -> (date [nrec, 3 or 4], values [ncol, nrec]); cols: [(name, format, type, mult)];
    header: the file starts with the PRT_HEADER block (its length is the fifth unsigned short); no_date: an OUTPUT_FORCE file"""
    dt = [("date", "<i4", 0 if no_date else (4 if hourly else 3))] + [("c%d" % k, OUT_TYPE_DTYPE[c[2]]) for k, c in enumerate(cols)]
    off = int(np.fromfile(path, dtype="<u2", count=5)[4]) if header else 0
    a = np.fromfile(path, dtype=np.dtype(dt), offset=off)
    return a["date"], np.stack([a["c%d" % k].astype(np.float64) / cols[k][3] for k in range(len(cols))])
