"""ctypes loader for the forcing-preparation oracle (oracle/_build/libatmos_oracle.so, atmos_oracle.c).
TEST INFRASTRUCTURE: only tests/, tools/, __graft_entry__.smoke() and bench.py's CPU legs may import
this.  It reuses the product's ctypes struct declarations (vicres_b200/atmos.py), data layouts only."""
import ctypes as C
import os
import subprocess
import numpy as np
from vicres_b200 import abi, atmos

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "_build", "libatmos_oracle.so")
        src = os.path.join(HERE, "atmos_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(src) > os.path.getmtime(path):
            subprocess.run(["make", "-C", HERE, "atmos"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.ao_ndays.argtypes = [C.c_void_p]
        L.ao_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        _lib = L
    return _lib


def prepare_sub(opt, site, raw, c_begin=0, c_end=None):
    """initialize_atmos restated, serially over cells [c_begin, c_end): returns (fields, snowflag) -- fields [9, nrecs, C], or
    [9, nrecs, NF + 1, C] when opt.snow_step < opt.time_step; snowflag [nrecs, C] int32 (atmos->snowflag[NR])."""
    L = lib()
    ps, pr = atmos.pack_site(site), atmos.pack_raw(raw)
    Cn = ps.struct.ncell
    NF = atmos.n_windows(opt)
    out = np.zeros((abi.VR_NFORCING, opt.nrecs, Cn) if NF == 1 else (abi.VR_NFORCING, opt.nrecs, NF + 1, Cn))
    flag = np.zeros((opt.nrecs, Cn), dtype=np.int32)
    pd = C.POINTER(C.c_double)
    ptrs = (pd * abi.VR_NFORCING)(*[out[f].ctypes.data_as(pd) for f in range(abi.VR_NFORCING)])
    rc = L.ao_prepare(C.byref(opt), ps.ref(), pr.ref(), ptrs, flag.ctypes.data, c_begin, Cn if c_end is None else c_end)
    if rc != 0:
        raise RuntimeError("atmos oracle: unsupported configuration (%d)" % rc)
    return out, flag


def prepare(opt, site, raw, c_begin=0, c_end=None):
    """the fields alone"""
    return prepare_sub(opt, site, raw, c_begin, c_end)[0]
