"""Host-side mirror of the routing pass over the C ABI of include/vicres_route.h (ctypes only).

    r = Route(grid, station_col, station_row, uh_box)     # catchments + unit hydrographs (GPU)
    rid, cells, uh_s = r.node(n)                           # what the reference writes to <name>.uh_s
    flows = r.convolve(runoff, baseflow)                   # [nnodes, ndays] node inflows (m3/s)
    out = r.reservoirs(flows, [res_list], 2000, 1)         # reservoir day loop for one or many rule sets
"""
import ctypes as C
import numpy as np
from . import model, route_abi

ROUTE_EXPORTS = ["vr_route_create", "vr_route_destroy", "vr_route_last_error", "vr_route_num_nodes",
                 "vr_route_node_info", "vr_route_node_cells", "vr_route_get_uh", "vr_route_convolve",
                 "vr_route_convolve_ex", "vr_route_convolve_in", "vr_route_last_kernel_ms", "vr_route_res_info", "vr_route_get_uh_r",
                 "vr_route_set_uh_r", "vr_route_reservoirs", "vr_route_reservoirs_range", "vr_objectives", "vr_route_fluxes_from_step"]


def _lib():
    L = model.load_library()
    if not getattr(L, "_route_ready", False):
        vp = C.c_void_p
        L.vr_route_create.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int, C.POINTER(vp)]
        L.vr_route_destroy.argtypes = [vp]
        L.vr_route_destroy.restype = None
        L.vr_route_last_error.argtypes = [vp]
        L.vr_route_last_error.restype = C.c_char_p
        L.vr_route_num_nodes.argtypes = [vp]
        L.vr_route_node_info.argtypes = [vp, C.c_int32, vp, vp]
        L.vr_route_node_cells.argtypes = [vp, C.c_int32, vp]
        L.vr_route_get_uh.argtypes = [vp, C.c_int32, vp]
        L.vr_route_convolve.argtypes = [vp, C.c_int32, vp, vp, C.c_float, vp]
        L.vr_route_convolve_ex.argtypes = [vp, C.c_int32, vp, vp, C.c_float] + [vp] * 7
        L.vr_route_convolve_in.argtypes = [vp, C.c_int32, vp, vp, vp]
        L.vr_route_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.vr_route_res_info.argtypes = [vp, C.c_int32, vp, vp]
        L.vr_route_get_uh_r.argtypes = [vp, C.c_int32, vp]
        L.vr_route_set_uh_r.argtypes = [vp, C.c_int32, vp]
        L.vr_route_reservoirs.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
        L.vr_route_reservoirs_range.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
        L.vr_objectives.argtypes = [C.c_int, C.c_int32, C.c_int32, vp, vp, C.c_int32, vp]
        L.vr_route_fluxes_from_step.argtypes = [C.c_int, vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, C.c_int64, vp, vp]
        L._route_ready = True
    return L


class Route:
    def __init__(self, grid, station_col, station_row, uh_box, device=0):
        self.L = _lib()
        self.g = route_abi.pack_grid(grid)
        self.box = np.ascontiguousarray(uh_box, dtype=np.float32)
        assert self.box.size == route_abi.KE
        self.h = C.c_void_p()
        rc = self.L.vr_route_create(self.g.ref(), station_col, station_row, self.box.ctypes.data, device, C.byref(self.h))
        if rc != 0:
            raise model.VicResError(rc, self.L.vr_route_last_error(None).decode())
        self.nnodes = int(self.L.vr_route_num_nodes(self.h))
        self.ncell = int(self.g.struct.ncol) * int(self.g.struct.nrow)

    def _check(self, rc):
        if rc != 0:
            raise model.VicResError(rc, self.L.vr_route_last_error(self.h).decode())

    def node(self, n):
        rid, nc = C.c_int32(), C.c_int64()
        self._check(self.L.vr_route_node_info(self.h, n, C.byref(rid), C.byref(nc)))
        cells = np.zeros(nc.value, dtype=np.int32)
        uh = np.zeros((nc.value, route_abi.NS), dtype=np.float32)
        self._check(self.L.vr_route_node_cells(self.h, n, cells.ctypes.data))
        self._check(self.L.vr_route_get_uh(self.h, n, uh.ctypes.data))
        return int(rid.value), cells, uh

    def convolve(self, runoff, baseflow, scale=0.18308, cell_scale=None, evap=None, cell_factor=None, present=None):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        runoff, baseflow, cell_scale, cell_factor = f32(runoff), f32(baseflow), f32(cell_scale), f32(cell_factor)
        assert runoff.shape == baseflow.shape and runoff.shape[0] == self.ncell
        ev = [f32(a) for a in evap] if evap is not None else [None, None, None]
        pres = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
        ndays = runoff.shape[1]
        flows = np.zeros((self.nnodes, ndays), dtype=np.float32)
        p = lambda a: None if a is None else a.ctypes.data
        self._check(self.L.vr_route_convolve_ex(self.h, ndays, p(runoff), p(baseflow), C.c_float(scale), p(cell_scale),
                                                p(ev[0]), p(ev[1]), p(ev[2]), p(cell_factor), p(pres), p(flows)))
        return flows

    def convolve_in(self, **kw):
        """MAKE_CONVOLUTION in full (open-water evaporation of reservoir-surface cells included); keywords of
        route_abi.pack_route_inputs -> (flows [nnodes, ndays], res_evaporation [nnodes, ndays])"""
        s, keep, ndays = route_abi.pack_route_inputs(**kw)
        flows = np.zeros((self.nnodes, ndays), dtype=np.float32)
        resev = np.zeros((self.nnodes, ndays), dtype=np.float32)
        self._check(self.L.vr_route_convolve_in(self.h, ndays, C.cast(C.pointer(s), C.c_void_p), flows.ctypes.data,
                                                resev.ctypes.data))
        return flows, resev

    def convolve_in_device(self, ndays, ptrs, cell_factor=None, cell_scale=None, scale=0.18308):
        """vr_route_convolve_in on columns that are already in DEVICE memory (fluxes_from_step_device): ptrs maps
        runoff, baseflow, ev_vege, ev_soil, tr_vege (and optionally present) to device pointers of [ncell][ndays] float32."""
        s = route_abi.vr_route_inputs()
        pf = C.POINTER(C.c_float)
        keep = []
        for k in ("runoff", "baseflow", "ev_vege", "ev_soil", "tr_vege"):
            if ptrs.get(k):
                setattr(s, k, C.cast(ptrs[k], pf))
        if ptrs.get("present"):
            s.present = C.cast(ptrs["present"], C.POINTER(C.c_ubyte))
        s.scale = scale
        for name, a in (("cell_factor", cell_factor), ("cell_scale", cell_scale)):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                keep.append(a)
                setattr(s, name, a.ctypes.data_as(pf))
        flows = np.zeros((self.nnodes, ndays), dtype=np.float32)
        self._check(self.L.vr_route_convolve_in(self.h, ndays, C.cast(C.pointer(s), C.c_void_p), flows.ctypes.data, None))
        return flows

    def res_info(self, n):
        """(downstream reservoir id, node the release is routed to, UH_R[107]) of reservoir node n"""
        ds, tg = C.c_int32(), C.c_int32()
        self._check(self.L.vr_route_res_info(self.h, n, C.byref(ds), C.byref(tg)))
        uh = np.zeros(route_abi.NS, dtype=np.float32)
        self._check(self.L.vr_route_get_uh_r(self.h, n, uh.ctypes.data))
        return int(ds.value), int(tg.value), uh

    def set_uh_r(self, n, uh_r):
        uh_r = np.ascontiguousarray(uh_r, dtype=np.float32)
        assert uh_r.size == route_abi.NS
        self._check(self.L.vr_route_set_uh_r(self.h, n, uh_r.ctypes.data))

    def reservoirs(self, natural, sets, start_year, start_month, legacy=False, days=None, state=None):
        """MAKE_CONVOLUTIONRS for len(sets) parameter sets at once.  natural: [nnodes, ndays] from convolve();
        sets: list of lists of route_abi.reservoir dicts (node order).  Returns dict: flow [nsets, ndays] and
        vol/hho [nsets, nres, ndays+1], flowin/flowout/turbine/energy/htk [nsets, nres, ndays].
        days = (d0, d1) operates only the days [d0, d1) (the reference's STEPBYSTEP mode, vr_route_reservoirs_range);
        with d0 > 0 `state` is the dict a previous call returned (its series are continued in place)."""
        natural = np.ascontiguousarray(natural, dtype=np.float32)
        assert natural.shape[0] == self.nnodes
        ndays, nres, nsets = natural.shape[1], self.nnodes - 1, len(sets)
        arr, keep = route_abi.pack_reservoirs(sets) if nres else (None, None)
        opt = route_abi.vr_res_options(start_year, start_month, int(legacy))
        d0, d1 = days if days is not None else (0, ndays)
        out = state if (state is not None and d0 > 0) else {"flow": np.zeros((nsets, ndays), dtype=np.float32)}
        ser = route_abi.vr_res_series()
        for k, ext in route_abi.RES_SERIES:
            if k not in out:
                out[k] = np.zeros((nsets, nres, ndays + ext), dtype=np.float32)
            assert out[k].shape == (nsets, nres, ndays + ext) and out[k].dtype == np.float32 and out[k].flags.c_contiguous
            setattr(ser, k, out[k].ctypes.data_as(C.POINTER(C.c_float)))
        self._check(self.L.vr_route_reservoirs_range(self.h, ndays, nsets, int(d0), int(d1), natural.ctypes.data,
                                                     C.cast(arr, C.c_void_p) if nres else None, C.cast(C.pointer(opt), C.c_void_p),
                                                     out["flow"].ctypes.data, C.cast(C.pointer(ser), C.c_void_p)))
        return out

    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self.L.vr_route_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def close(self):
        if self.h:
            self.L.vr_route_destroy(self.h)
            self.h = C.c_void_p()


def objectives(sim, obs, handle_negatives=True, device=0):
    """[nsets, 4] = NSE, TRMSE, MSDE, ROCE of sim [nsets, ndays] (float32 flows) against obs [ndays]"""
    L = _lib()
    sim = np.ascontiguousarray(sim, dtype=np.float32)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    assert sim.ndim == 2 and sim.shape[1] == obs.size
    out = np.zeros((sim.shape[0], 4), dtype=np.float64)
    rc = L.vr_objectives(device, sim.shape[0], sim.shape[1], sim.ctypes.data, obs.ctypes.data, int(handle_negatives),
                         out.ctypes.data)
    if rc != 0:
        raise model.VicResError(rc, L.vr_route_last_error(None).decode())
    return out


def fluxes_from_step_device(out_dev, names, grid_of_cell, ngrid, device=0):
    """The flux hand-off on the device.  out_dev: torch float64 CUDA tensor [n_out, ndays, C] (the step's output block),
    names: its VR_OUT names; grid_of_cell [C] (-1 = not on the routing grid).  Returns ({key: torch float32 CUDA tensor
    [ngrid, ndays]} for couple.KEYS, present uint8 CUDA tensor [ngrid]): what Route.convolve_in_device takes."""
    import torch
    from . import couple
    L = _lib()
    n_out, ndays, Cn = out_dev.shape
    rows = np.array([names.index(v) for v in couple.ROUTE_VARS], dtype=np.int32)
    gmap = np.ascontiguousarray(grid_of_cell, dtype=np.int64)
    dst = [torch.empty((ngrid, ndays), dtype=torch.float32, device=out_dev.device) for _ in couple.KEYS]
    present = torch.empty(ngrid, dtype=torch.uint8, device=out_dev.device)
    ptrs = (C.c_void_p * len(dst))(*[t.data_ptr() for t in dst])
    rc = L.vr_route_fluxes_from_step(device, out_dev.data_ptr(), Cn, ndays, len(dst), rows.ctypes.data, gmap.ctypes.data, ngrid,
                                     ptrs, present.data_ptr())
    if rc != 0:
        raise model.VicResError(rc, "vr_route_fluxes_from_step")
    return dict(zip(couple.KEYS, dst)), present
