"""Parameter-set axis for the toolbox's ensembles (calibration, eFAST, EET).

The toolbox evaluates one parameter set at a time: it rewrites columns of the soil parameter file and launches
`./vicNl` again (toolbox/calibration/basin_calibration_eNSGAII.py:177-229, `update_soil_parameters`: columns 5-9 =
b_infilt, Ds, Dsmax, Ws, c and the layer depths, each rounded to 6 decimals; toolbox/sensitivity/eFAST_analysis.py:97-114
scales the unit samples into the parameter ranges).  Here P parameter sets x N basin cells become P*N cells of ONE
handle that share N forcing columns (vr_set_forcing_map), so a whole population is one `vr_step_block` call.

`paramset_cells` goes through the same door the toolbox uses -- the soil parameter file -- so that everything
read_soilparam derives from the changed columns (max_moist, Wcr, Wpwp, the water-table curves ...) is derived by the
library's reader exactly as for a single run.  Host-side file handling and index arithmetic only.
"""
import os
import tempfile
import numpy as np
from . import files

# eFAST_analysis.py:97-114 / EET_analysis.py: unit sample -> parameter value
RANGES = {"Dsmax": (0.0, 30.0), "b_infilt": (0.0, 0.9), "c": (1.0, 3.0), "d1": (0.05, 0.25), "d2": (0.3, 1.5),
          "d3": (0.3, 1.5), "velocity": (0.5, 5.0), "diffusivity": (200.0, 4000.0), "Ds": (0.0, 0.9), "Ws": (0.0, 0.9)}
SOIL_PARAMS = ["b_infilt", "Ds", "Dsmax", "Ws", "c"]          # soil-file columns 5-9 (1-based), in file order


def scale_samples(X, names):
    """X: [P, len(names)] unit samples (what SAFE's AAT_sampling / OAT_sampling return); the toolbox's scaling block."""
    X = np.array(X, dtype=np.float64, copy=True)
    for j, n in enumerate(names):
        lo, hi = RANGES[n]
        X[:, j] = X[:, j] * hi if lo == 0.0 else X[:, j] * (hi - lo) + lo
    return X


def depth_columns(nlayer):
    """0-based soil-file columns of the layer depths: 4 + 5 + 4*L + 1 (read_soilparam.c:170-600)."""
    return [9 + 4 * nlayer + 1 + l for l in range(nlayer)]


def write_paramset_soil(soil_path, out_path, sets, names, nlayer, cell_mask=None):
    """Write a soil file holding, for every parameter set, a copy of the rows flagged to run with the named columns
    replaced (values rounded to 6 decimals as update_soil_parameters does).  names: from SOIL_PARAMS + d1..dL.
    cell_mask: optional [N] bool over the run-flagged rows (a calibration zone); other rows keep their values."""
    rows = [ln.split() for ln in open(soil_path) if ln.split() and ln.split()[0] != "0" and not ln.startswith("#")]
    col = {n: 4 + i for i, n in enumerate(SOIL_PARAMS)}
    for l, cidx in enumerate(depth_columns(nlayer)):
        col["d%d" % (l + 1)] = cidx
    sets = np.asarray(sets, dtype=np.float64)
    with open(out_path, "w") as f:
        for p in range(sets.shape[0]):
            for i, r in enumerate(rows):
                r = list(r)
                if cell_mask is None or cell_mask[i]:
                    for j, n in enumerate(names):
                        r[col[n]] = repr(round(float(sets[p, j]), 6))
                f.write("\t".join(r) + "\n")
    return len(rows)


def paramset_cells(param_files, sets, names, cell_mask=None):
    """param_files: keyword arguments of files.read_params (e.g. files.Global(...).param_files()).
    Returns (veglib, cells, meta, forcing_map, N): cells holds P*N cells, set-major; forcing_map[c] = c % N is what
    VicRes(..., forcing_map=..., n_forcing_cols=N) takes."""
    kw = dict(param_files)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "soil_paramsets.txt")
        N = write_paramset_soil(kw["soil"], path, sets, names, kw.get("nlayer", 3), cell_mask)
        kw["soil"] = path
        lib, cells, meta = files.read_params(**kw)
    P = len(sets)
    assert len(meta["gridcel"]) == P * N
    return lib, cells, meta, np.tile(np.arange(N, dtype=np.int64), P), N
