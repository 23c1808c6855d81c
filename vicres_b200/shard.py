"""Cell-block sharding for multi-GPU runs: the runoff step has no lateral
exchange (runoff/model/vicNl.c:197-356 loops over independent cells), so rank r of W simply owns the
contiguous cell block cell_block(C, r, W) -- no data-path collective.  The same helpers shard the
parameter-set axis of an ensemble (virtual cells = sets x basin cells)."""
import numpy as np
from . import abi


def cell_block(ncell, rank, world):
    """[c0, c1) of rank `rank`; block sizes differ by at most one."""
    base, rem = divmod(ncell, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def slice_cells(cells, c0, c1):
    """The vr_cells arrays (dict layout of abi.pack_cells) restricted to cells [c0, c1)."""
    out = {}
    for k in abi.CELL_SCALARS:
        out[k] = np.ascontiguousarray(cells[k][c0:c1])
    for k in abi.CELL_LAYER + abi.CELL_BAND:
        out[k] = np.ascontiguousarray(cells[k][:, c0:c1])
    tp = np.asarray(cells["tile_ptr"])
    t0, t1 = int(tp[c0]), int(tp[c1])
    out["tile_ptr"] = (tp[c0:c1 + 1] - t0).astype(np.int64)
    for k in ["tile_class", "tile_Cv", "tile_root", "tile_LAI", "tile_albedo", "tile_vegcover"]:
        out[k] = np.ascontiguousarray(np.asarray(cells[k])[t0:t1])
    for k in ["zwt_zwt", "zwt_moist"]:                 # optional [(L+2), 11, C] curves
        if cells.get(k) is not None:
            out[k] = np.ascontiguousarray(np.asarray(cells[k])[:, :, c0:c1])
    return out


def slice_atmos_inputs(site, raw, c0, c1):
    """The forcing-preparation inputs (vicres_b200/atmos.py: site values and raw forcing columns [nrec, C]) of cells
    [c0, c1): initialize_atmos is per cell too, so the ranks prepare their own blocks with no collective."""
    s = {k: np.ascontiguousarray(np.asarray(v)[c0:c1]) for k, v in site.items() if v is not None}
    r = {k: (np.ascontiguousarray(np.asarray(v)[:, c0:c1]) if v is not None else None) for k, v in raw.items()}
    return s, r
