"""ctypes loader for the routing oracle (oracle/_build/libroute_oracle.so) -- TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess
import numpy as np
from vicres_b200 import route_abi

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "_build", "libroute_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "route_oracle.c")):
            subprocess.run(["make", "-C", HERE, "route"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.ro_create.restype = C.c_void_p
        L.ro_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ro_destroy.argtypes = [C.c_void_p]
        L.ro_num_nodes.argtypes = [C.c_void_p]
        L.ro_node_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ro_node_cells.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ro_get_uh.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ro_convolve.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        L.ro_convolve_in.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ro_res_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ro_get_uh_r.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ro_reservoirs.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
        _lib = L
    return _lib


class RouteOracle:
    def __init__(self, grid, station_col, station_row, uh_box):
        self.g = route_abi.pack_grid(grid)
        self.box = np.ascontiguousarray(uh_box, dtype=np.float32)
        self.h = lib().ro_create(self.g.ref(), station_col, station_row, self.box.ctypes.data)
        self.nnodes = lib().ro_num_nodes(self.h)

    def node(self, n):
        rid, nc = C.c_int(), C.c_longlong()
        lib().ro_node_info(self.h, n, C.byref(rid), C.byref(nc))
        cells = np.zeros(nc.value, dtype=np.int32)
        lib().ro_node_cells(self.h, n, cells.ctypes.data)
        uh = np.zeros((nc.value, route_abi.NS), dtype=np.float32)
        lib().ro_get_uh(self.h, n, uh.ctypes.data)
        return rid.value, cells, uh

    def convolve(self, runoff, baseflow, cell_scale=None, evap=None, cell_factor=None, present=None):
        """runoff, baseflow: [ncell, ndays]; evap = (ev_vege, ev_soil, tr_vege) or None"""
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        runoff, baseflow, cell_scale, cell_factor = f32(runoff), f32(baseflow), f32(cell_scale), f32(cell_factor)
        ev = [f32(a) for a in evap] if evap is not None else [None, None, None]
        ndays = runoff.shape[1]
        flows = np.zeros((self.nnodes, ndays), dtype=np.float32)
        p = lambda a: None if a is None else a.ctypes.data
        pres = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
        lib().ro_convolve(self.h, ndays, p(runoff), p(baseflow), p(cell_scale), p(ev[0]), p(ev[1]), p(ev[2]),
                          p(cell_factor), p(pres), p(flows))
        return flows

    def convolve_in(self, **kw):
        """MAKE_CONVOLUTION in full; keywords of route_abi.pack_route_inputs -> (flows, res_evaporation)"""
        s, keep, ndays = route_abi.pack_route_inputs(**kw)
        flows = np.zeros((self.nnodes, ndays), dtype=np.float32)
        resev = np.zeros((self.nnodes, ndays), dtype=np.float32)
        lib().ro_convolve_in(self.h, ndays, C.cast(C.pointer(s), C.c_void_p), flows.ctypes.data, resev.ctypes.data)
        return flows, resev

    def res_info(self, n):
        ds, tg = C.c_int(), C.c_int()
        lib().ro_res_info(self.h, n, C.byref(ds), C.byref(tg))
        uh = np.zeros(route_abi.NS, dtype=np.float32)
        lib().ro_get_uh_r(self.h, n, uh.ctypes.data)
        return ds.value, tg.value, uh

    def reservoirs_sets(self, natural, sets, start_year, start_month, legacy=False):
        """the Route.reservoirs surface: a list of parameter sets -> dict with a leading set axis"""
        outs = [self.reservoirs(natural, s, start_year, start_month, legacy=legacy) for s in sets]
        return {k: np.stack([o[k] for o in outs]) for k in outs[0]} if outs else {"flow": np.zeros((0, natural.shape[1]), np.float32)}

    def reservoirs(self, natural, res, start_year, start_month, legacy=False):
        """natural [nnodes, ndays]; res: list of route_abi.reservoir dicts (node order) -> dict of series"""
        natural = np.ascontiguousarray(natural, dtype=np.float32)
        ndays, nres = natural.shape[1], self.nnodes - 1
        arr, keep = route_abi.pack_reservoirs([res])
        opt = route_abi.vr_res_options(start_year, start_month, int(legacy))
        out = {"flow": np.zeros(ndays, dtype=np.float32)}
        for k, ext in route_abi.RES_SERIES:
            out[k] = np.zeros((nres, ndays + ext), dtype=np.float32)
        lib().ro_reservoirs(self.h, ndays, natural.ctypes.data, C.cast(arr, C.c_void_p), C.cast(C.pointer(opt), C.c_void_p),
                            out["flow"].ctypes.data, *[out[k].ctypes.data for k, _ in route_abi.RES_SERIES])
        return out

    def close(self):
        if self.h:
            lib().ro_destroy(self.h)
            self.h = None
