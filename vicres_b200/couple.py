"""Hand-off from the cell step to the routing pass: what MAKE_CONVOLUTION reads from the fluxes_<lat>_<lng>
files (routing/model/reservoir.f:373-389), either from the files themselves or straight from the step's
output block in memory (no 10^6 files between the two GPU passes).

The classic reader takes, of the default flux file, columns RUNOFF, BASEFLOW, EVAP_CANOP, TRANSP_VEG,
EVAP_BARE, REL_HUMID, AIR_TEMP, WIND (2-layer layout hard-coded, reservoir.f:383-386)."""
import os
import numpy as np
from . import route_abi

ROUTE_VARS = ["RUNOFF", "BASEFLOW", "EVAP_CANOP", "TRANSP_VEG", "EVAP_BARE", "REL_HUMID", "AIR_TEMP", "WIND"]
KEYS = ["runoff", "baseflow", "ev_vege", "tr_vege", "ev_soil", "rh", "tair", "wind"]


def _grid_index(hdr, nrow, ncol, grid_decimal):
    """file name suffix -> file-order grid cell, as create_vic_names builds it from JLOC/ILOC"""
    idx = {}
    for r in range(nrow):
        for c in range(ncol):
            lat, lon = route_abi.cell_latlon(hdr, nrow, ncol, r, c)
            idx["%.*f_%.*f" % (grid_decimal, lat, grid_decimal, lon)] = r * ncol + c
    return idx


def inputs_from_outputs(out, names, lat, lng, hdr, nrow, ncol, grid_decimal=4, through_text=True):
    """out: [n_out, ndays, C] step outputs (daily), names: their VR_OUT names; lat/lng of the C cells.
    Returns dict(runoff, baseflow, ev_vege, tr_vege, ev_soil, rh, tair, wind: [nrow*ncol, ndays] float32,
    present [nrow*ncol] uint8).  through_text rounds every value through "%.4f" as the file round trip does
    (what makes the in-memory hand-off equal to the reference's files -> READ(20,*) path)."""
    sel = [names.index(v) for v in ROUTE_VARS]
    ndays, C = out.shape[1], out.shape[2]
    idx = _grid_index(hdr, nrow, ncol, grid_decimal)
    res = {k: np.zeros((nrow * ncol, ndays), dtype=np.float32) for k in KEYS}
    present = np.zeros(nrow * ncol, dtype=np.uint8)
    block = out[sel]                                            # [8, ndays, C]
    if through_text:
        from . import files
        block = files.round_decimals(block, 4)                   # %.4f, then list-directed READ into REAL
    for c in range(C):
        key = "%.*f_%.*f" % (grid_decimal, lat[c], grid_decimal, lng[c])
        g = idx.get(key)
        if g is None:
            continue                                             # a VIC cell outside the routing grid
        present[g] = 1
        for j, k in enumerate(KEYS):
            res[k][g] = block[j, :, c].astype(np.float32)
    res["present"] = present
    return res


def inputs_from_files(path_prefix, hdr, nrow, ncol, ndays, grid_decimal=4):
    """the same arrays read from <path_prefix><lat>_<lng> files in the 2-layer default layout"""
    res = {k: np.zeros((nrow * ncol, ndays), dtype=np.float32) for k in KEYS}
    present = np.zeros(nrow * ncol, dtype=np.uint8)
    cols = [5, 6, 12, 13, 14, 20, 22, 23]                        # 0-based, after year month day
    for name, g in _grid_index(hdr, nrow, ncol, grid_decimal).items():
        fn = path_prefix + name
        if not os.path.exists(fn):
            continue
        a = np.loadtxt(fn, ndmin=2)[:ndays]
        present[g] = 1
        for j, k in enumerate(KEYS):
            res[k][g] = a[:, cols[j]].astype(np.float32)
    res["present"] = present
    return res
