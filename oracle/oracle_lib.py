"""ctypes loader for the oracle port (oracle/_build/libvicres_oracle.so).  TEST INFRASTRUCTURE:
only tests/, tools/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  It reuses the product's ctypes struct declarations (vicres_b200/abi.py), which
are data layouts only."""
import ctypes as C
import os
import subprocess
import numpy as np
from vicres_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
VO_NUNITF = 48


def build(quiet=True):
    subprocess.run(["make", "-C", HERE, "port"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)
    return os.path.join(HERE, "_build", "libvicres_oracle.so")


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "_build", "libvicres_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.vo_create.restype = C.c_void_p
        L.vo_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vo_destroy.argtypes = [C.c_void_p]
        L.vo_init_state.argtypes = [C.c_void_p, C.c_void_p]
        L.vo_step_block.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vo_step_block_range.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.vo_fallback_counts.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, opt, lib_dict, cells_dict):
        self.opt = opt
        self._lib = abi.pack_veglib(lib_dict)
        self._cells = abi.pack_cells(cells_dict)
        self.ncell = len(cells_dict["lat"])
        self.h = lib().vo_create(C.byref(opt), self._lib.ref(), self._cells.ref())

    def init_state(self, tair0):
        t = np.ascontiguousarray(tair0, dtype=np.float64)
        return lib().vo_init_state(self.h, t.ctypes.data)

    def step(self, month, doy, forcing, units_maxtile=0, snowflag=None):
        f = abi.pack_forcing(month, doy, forcing, snowflag)
        nrec = len(month)
        out = np.zeros((self.opt.n_out, nrec, self.ncell))
        units = None
        if units_maxtile:
            units = np.zeros((nrec, self.ncell, units_maxtile, self.opt.nbands, VO_NUNITF))
        rc = lib().vo_step_block(self.h, f.ref(), out.ctypes.data,
                                 units.ctypes.data if units is not None else None, units_maxtile)
        return rc, out, units

    def step_range(self, month, doy, forcing, out, c0, c1, snowflag=None):
        f = abi.pack_forcing(month, doy, forcing, snowflag)
        return lib().vo_step_block_range(self.h, f.ref(), out.ctypes.data, c0, c1)

    def fallback_counts(self):
        a = np.zeros(2, dtype=np.int64)
        lib().vo_fallback_counts(self.h, a.ctypes.data)
        return a

    def close(self):
        if self.h:
            lib().vo_destroy(self.h)
            self.h = None
