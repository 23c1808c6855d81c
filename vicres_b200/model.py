"""Host-side mirror of the reference call seam for the cell step, over the C ABI of libvicres.so.

The reference runs, per cell, `initialize_model_state(...)` then per record
`full_energy(...)` + `put_data(...)` (runoff/model/vicNl.c:260-306).  `VicRes` batches that:

    m = VicRes(opt, veglib, cells)        # read_veglib / read_soilparam / read_vegparam hand-off
    m.init_state(tair0)                   # initialize_model_state + put_data(rec = -nrecs)
    out, status = m.step(month, doy, forcing)   # full_energy + put_data for a block of records

Everything numerical happens in the CUDA library; this module only marshals numpy / torch buffers
into the plain-pointer structs of include/vicres.h.  There is no CPU implementation behind it: if
libvicres.so is missing or no CUDA device is present the constructor raises.
"""
import ctypes as C
import os
import numpy as np
from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("VICRES_LIB", os.path.join(HERE, "libvicres.so"))

EXPORTS = ["vr_create", "vr_destroy", "vr_last_error", "vr_default_options", "vr_set_veglib",
           "vr_set_cells", "vr_set_forcing_map", "vr_init_state", "vr_step_block",
           "vr_step_block_device", "vr_synchronize", "vr_state_size", "vr_get_state", "vr_set_state",
           "vr_write_state_file", "vr_read_state_file", "vr_num_units", "vr_num_launches", "vr_last_kernel_ms", "vr_kernel_times", "vr_diag_pow", "vr_diag_fp64_peak", "vr_fallback_counts", "vr_step_diag",
           "vr_full_energy_compat"]

_lib = None


class VicResError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vicres error %d: %s" % (code, msg))
        self.code = code


def load_library(path=LIBPATH):
    """dlopen libvicres.so and declare prototypes; raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not found: build it with `python -m vicres_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.vr_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.vr_destroy.argtypes = [vp]
    L.vr_destroy.restype = None
    L.vr_last_error.argtypes = [vp]
    L.vr_last_error.restype = C.c_char_p
    L.vr_default_options.argtypes = [vp, C.c_int]
    L.vr_default_options.restype = None
    L.vr_set_veglib.argtypes = [vp, vp]
    L.vr_set_cells.argtypes = [vp, vp]
    L.vr_set_forcing_map.argtypes = [vp, vp, C.c_int64]
    L.vr_init_state.argtypes = [vp, vp]
    L.vr_step_block.argtypes = [vp, vp, vp, vp]
    L.vr_step_block_device.argtypes = [vp, vp, vp, vp]
    L.vr_synchronize.argtypes = [vp]
    L.vr_write_state_file.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    L.vr_read_state_file.argtypes = [vp, C.c_char_p, C.c_int32, vp, vp]
    L.vr_state_size.argtypes = [vp]
    L.vr_state_size.restype = C.c_int64
    L.vr_get_state.argtypes = [vp, vp]
    L.vr_set_state.argtypes = [vp, vp]
    L.vr_num_units.argtypes = [vp]
    L.vr_num_units.restype = C.c_int64
    L.vr_num_launches.argtypes = [vp]
    L.vr_num_launches.restype = C.c_int64
    L.vr_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.vr_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.vr_diag_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.vr_diag_pow.argtypes = [C.c_int, C.c_int64, vp, vp, vp]
    L.vr_fallback_counts.argtypes = [vp, vp]
    L.vr_step_diag.argtypes = [vp, vp]
    L.vr_full_energy_compat.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp, vp]
    _lib = L
    return L


def diag_fp64_peak(device=0):
    """measured DFMA thread-instructions per second of the device"""
    v = C.c_double()
    rc = load_library().vr_diag_fp64_peak(device, C.byref(v))
    if rc != 0:
        raise VicResError(rc, "vr_diag_fp64_peak")
    return float(v.value)


def diag_pow(x, y, device=0):
    """the device power function on host arrays (diagnostic)"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = np.empty_like(x)
    rc = load_library().vr_diag_pow(device, x.size, x.ctypes.data, y.ctypes.data, z.ctypes.data)
    if rc != 0:
        raise VicResError(rc, "vr_diag_pow")
    return z


class VicRes:
    def __init__(self, opt, veglib, cells, device=0, forcing_map=None, n_forcing_cols=None):
        self.L = load_library()
        self.opt = opt
        self.h = C.c_void_p()
        rc = self.L.vr_create(C.byref(opt), device, C.byref(self.h))
        if rc != 0:
            raise VicResError(rc, self.L.vr_last_error(None).decode())
        self._lib = abi.pack_veglib(veglib)
        self._cells = abi.pack_cells(cells)
        self.ncell = int(self._cells.struct.ncell)
        self.ncol = self.ncell
        self._check(self.L.vr_set_veglib(self.h, self._lib.ref()))
        self._check(self.L.vr_set_cells(self.h, self._cells.ref()))
        if forcing_map is not None:
            fm = np.ascontiguousarray(forcing_map, dtype=np.int64)
            self.ncol = int(n_forcing_cols if n_forcing_cols is not None else fm.max() + 1)
            self._check(self.L.vr_set_forcing_map(self.h, fm.ctypes.data, self.ncol))

    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise VicResError(rc, self.L.vr_last_error(self.h).decode())
        return rc

    @property
    def n_out(self):
        return int(self.opt.n_out)

    @property
    def num_units(self):
        return int(self.L.vr_num_units(self.h))

    @property
    def num_launches(self):
        return int(self.L.vr_num_launches(self.h))

    def init_state(self, tair0):
        t = np.ascontiguousarray(tair0, dtype=np.float64)
        assert t.shape == (self.ncol,)
        self._check(self.L.vr_init_state(self.h, t.ctypes.data))

    def step(self, month, doy, forcing, out=None, snowflag=None):
        """forcing: [9, nrec, ncol] float64 host array (numpy, or a pinned CPU torch tensor's
        .numpy()); with a sub-stepped snow model (snow_step_hours < dt_hours, NF passes) [9, nrec * (NF + 1), ncol]
        and snowflag [nrec, ncol] int32.  Returns (out [n_out, nrec, C], cell_status [C])."""
        f = abi.pack_forcing(month, doy, forcing, snowflag)
        nrec = f.struct.nrec
        nf = self.opt.dt_hours // self.opt.snow_step_hours
        ns = nf + 1 if nf > 1 else 1
        for i in range(abi.VR_NFORCING):
            assert f.keep["f%d" % i].size == nrec * ns * self.ncol, "forcing field %d has the wrong size" % i
        if out is None:
            out = np.empty((self.n_out, nrec, self.ncell), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self.n_out * nrec * self.ncell
        status = np.zeros(self.ncell, dtype=np.int32)
        self._check(self.L.vr_step_block(self.h, f.ref(), out.ctypes.data, status.ctypes.data),
                    allow=(abi.VR_ERR_CELL,))
        return out, status

    def full_energy(self, rec, month, doy, atmos, snowflag=None):
        """one record of every cell, the reference's `full_energy(...); put_data(...)` pair (vr_full_energy_compat):
        atmos [9, C] (or [9, NF + 1, C]); returns (out [n_out, C], cell_status [C])"""
        a = np.ascontiguousarray(atmos, dtype=np.float64)
        out = np.empty((self.n_out, self.ncell), dtype=np.float64)
        status = np.zeros(self.ncell, dtype=np.int32)
        sf = None if snowflag is None else np.ascontiguousarray(snowflag, dtype=np.int32)
        self._check(self.L.vr_full_energy_compat(self.h, int(rec), int(month), int(doy), a.ctypes.data,
                                                 sf.ctypes.data if sf is not None else None, out.ctypes.data, status.ctypes.data),
                    allow=(abi.VR_ERR_CELL,))
        return out, status

    def step_device(self, month, doy, forcing_ptrs, out_ptr, stream=None, snowflag_ptr=None):
        """forcing_ptrs: 9 device pointers (int) to [nrec, ncol] float64 arrays ([nrec * (NF + 1), ncol] with a sub-stepped
        snow model, then with snowflag_ptr -> [nrec, ncol] int32 on the device); out_ptr: device pointer to
        [n_out, nrec, C].  Asynchronous on `stream` (int cudaStream_t) or the handle's."""
        s = abi.vr_forcing()
        mon = np.ascontiguousarray(month, dtype=np.int32)
        dy = np.ascontiguousarray(doy, dtype=np.int32)
        s.nrec = len(mon)
        s.month = mon.ctypes.data_as(C.POINTER(C.c_int32))
        s.day_in_year = dy.ctypes.data_as(C.POINTER(C.c_int32))
        for i in range(abi.VR_NFORCING):
            s.field[i] = C.cast(C.c_void_p(int(forcing_ptrs[i])), C.POINTER(C.c_double))
        if snowflag_ptr:
            s.snowflag = C.cast(C.c_void_p(int(snowflag_ptr)), C.POINTER(C.c_int32))
        self._keep = (mon, dy, s)
        self._check(self.L.vr_step_block_device(self.h, C.byref(s), C.c_void_p(int(out_ptr)),
                                                C.c_void_p(int(stream)) if stream else None))

    def write_state_file(self, path, year, month, day, gridcel, binary=False):
        """the reference's state file (write_model_state.c) of the current state"""
        g = np.ascontiguousarray(gridcel, dtype=np.int32)
        self._check(self.L.vr_write_state_file(self.h, path.encode(), year, month, day, int(binary), g.ctypes.data))

    def read_state_file(self, path, tair0, gridcel, binary=False):
        """restart from a state file of the reference (read_initial_model_state.c); replaces init_state()"""
        g = np.ascontiguousarray(gridcel, dtype=np.int32)
        t = np.ascontiguousarray(tair0, dtype=np.float64)
        self._check(self.L.vr_read_state_file(self.h, path.encode(), int(binary), t.ctypes.data, g.ctypes.data))

    def synchronize(self):
        self._check(self.L.vr_synchronize(self.h))

    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self.L.vr_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def kernel_times(self):
        """(step-kernel ms, k_cells ms, chunks, records) summed since the previous call; synchronises"""
        a, b, n, r = C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
        self._check(self.L.vr_kernel_times(self.h, C.byref(a), C.byref(b), C.byref(n), C.byref(r)))
        return float(a.value), float(b.value), int(n.value), int(r.value)

    def step_diag(self):
        """dict of the step-kernel diagnostics since the previous call (vr_step_diag); synchronises"""
        a = np.zeros(7, dtype=np.float64)
        self._check(self.L.vr_step_diag(self.h, a.ctypes.data))
        return {"front_ms": a[0], "snow_ms": a[1], "back_ms": a[2], "back_list_ms": a[6], "records_timed": int(a[3]),
                "snow_unit_records": int(a[4]), "unit_records": int(a[5])}

    def fallback_counts(self):
        a = np.zeros(2, dtype=np.int64)
        self._check(self.L.vr_fallback_counts(self.h, a.ctypes.data))
        return a

    def get_state(self):
        n = int(self.L.vr_state_size(self.h))
        blob = np.empty(n, dtype=np.uint8)
        self._check(self.L.vr_get_state(self.h, blob.ctypes.data))
        return blob

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        assert blob.size == int(self.L.vr_state_size(self.h))
        self._check(self.L.vr_set_state(self.h, blob.ctypes.data))

    def close(self):
        if self.h:
            self.L.vr_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
