"""Parameter-set axis (vicres_b200/ensemble.py) against the toolbox's own way of evaluating a parameter set: rewrite
the soil file, read it again (toolbox/calibration/basin_calibration_eNSGAII.py:177-221), run."""
import os
import numpy as np
import pandas as pd
import pytest
from common import GOLDEN
from vicres_b200 import abi, ensemble, files

CASE = os.path.join(GOLDEN, "files_case")
_L, _NB, _DT = [int(x) for x in np.load(os.path.join(CASE, "expected.npz"))["meta"]]
KW = dict(soil=os.path.join(CASE, "soil.txt"), veglib=os.path.join(CASE, "veglib.txt"), vegparam=os.path.join(CASE, "veg.txt"),
          snowband=os.path.join(CASE, "snowband.txt"), nlayer=_L, nbands=_NB, root_zones=3)
NAMES = ["b_infilt", "Ds", "Dsmax", "Ws", "c", "d1", "d2", "d3"]


def _sets(P, seed=3):
    """unit samples scaled the toolbox's way; Ds is kept below Ws (above it the ARNO curve has a negative nonlinear
    branch, runoff.c:497-507, and the reference itself runs into NaN)"""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.02, 0.98, (P, len(NAMES)))
    X[:, NAMES.index("Ds")] *= 0.4
    X[:, NAMES.index("Ws")] = 0.5 + 0.5 * X[:, NAMES.index("Ws")]
    return ensemble.scale_samples(X, NAMES)


def _toolbox_rewrite(soil_in, soil_out, values):
    """update_soil_parameters for one zone covering every cell, with the toolbox's pandas calls (NLAYER 3: depth columns 23-25)"""
    soil_lines = pd.read_csv(soil_in, delimiter="\t", header=None)
    soil_lines.columns = [f"Col{i+1}" for i in range(soil_lines.shape[1])]
    for col, value in zip(["Col5", "Col6", "Col7", "Col8", "Col9", "Col23", "Col24", "Col25"], values):
        soil_lines.loc[:, col] = round(float(value), 6)
    soil_lines.to_csv(soil_out, sep="\t", index=False, header=False)


def test_scaling_block_matches_the_toolbox_ranges():
    X = np.array([[0.0] * len(NAMES), [1.0] * len(NAMES), [0.5] * len(NAMES)])
    S = ensemble.scale_samples(X, NAMES)
    assert np.allclose(S[0], [0, 0, 0, 0, 1, 0.05, 0.3, 0.3]) and np.allclose(S[1], [0.9, 0.9, 30, 0.9, 3, 0.25, 1.5, 1.5])
    assert np.allclose(S[2], [0.45, 0.45, 15, 0.45, 2, 0.15, 0.9, 0.9])


def test_paramset_cells_equal_the_toolbox_rewrite_read_again(tmp_path):
    P = 4
    sets = _sets(P)
    sets[:, 0] = np.maximum(sets[:, 0], 0.01)
    lib, cells, meta, fmap, N = ensemble.paramset_cells(KW, sets, NAMES)
    assert N == 5 and len(fmap) == P * N and list(fmap[:N]) == list(range(N))
    for p in range(P):
        one = str(tmp_path / ("soil_%d.txt" % p))
        _toolbox_rewrite(KW["soil"], one, sets[p])
        kw = dict(KW, soil=one)
        _, ref, _ = files.read_params(**kw)
        sl = slice(p * N, (p + 1) * N)
        for k in abi.CELL_SCALARS:
            assert np.array_equal(np.asarray(cells[k])[sl], ref[k]), k
        for k in abi.CELL_LAYER + abi.CELL_BAND:
            assert np.array_equal(np.asarray(cells[k])[:, sl], ref[k]), k
        assert np.array_equal(np.asarray(cells["zwt_moist"])[:, :, sl], ref["zwt_moist"])
        t0, t1 = cells["tile_ptr"][p * N], cells["tile_ptr"][(p + 1) * N]
        assert np.array_equal(np.asarray(cells["tile_Cv"])[t0:t1], ref["tile_Cv"])
    # the sets really differ where they should
    assert len(set(np.asarray(cells["Dsmax"])[::N])) == P and len(set(np.asarray(cells["max_moist"])[2, ::N])) == P


@pytest.mark.gpu
def test_population_in_one_call_equals_one_run_per_set():
    """P sets x N cells in one handle sharing N forcing columns = P separate handles, bit for bit."""
    from vicres_b200 import model
    z = np.load(os.path.join(CASE, "expected.npz"))
    P = 6
    sets = _sets(P, seed=9)
    sets[:, 0] = np.maximum(sets[:, 0], 0.01)
    lib, cells, meta, fmap, N = ensemble.paramset_cells(KW, sets, NAMES)
    L, NB, dt = [int(x) for x in z["meta"]]
    opt = abi.default_options(L, NB, dt, out=["RUNOFF", "BASEFLOW", "EVAP", "SWE"])
    m = model.VicRes(opt, lib, cells, forcing_map=fmap, n_forcing_cols=N)
    m.init_state(z["tair0"])
    out, status = m.step(z["month"], z["doy"], z["forcing"])
    assert not status.any() and np.isfinite(out).all()
    m.close()
    for p in (0, P - 1):
        one = {k: (np.asarray(v)[..., p * N:(p + 1) * N] if k in abi.CELL_SCALARS + abi.CELL_LAYER + abi.CELL_BAND + ["zwt_zwt", "zwt_moist"] else v)
               for k, v in cells.items()}
        t0, t1 = cells["tile_ptr"][p * N], cells["tile_ptr"][(p + 1) * N]
        one["tile_ptr"] = np.asarray(cells["tile_ptr"])[p * N:(p + 1) * N + 1] - t0
        for k in ["tile_class", "tile_Cv", "tile_LAI", "tile_albedo", "tile_vegcover"]:
            one[k] = np.asarray(cells[k])[t0:t1] if np.asarray(cells[k]).ndim == 1 else np.asarray(cells[k])[t0:t1]
        one["tile_root"] = np.asarray(cells["tile_root"])[t0:t1]
        m1 = model.VicRes(opt, lib, one)
        m1.init_state(z["tair0"])
        o1, _ = m1.step(z["month"], z["doy"], z["forcing"])
        m1.close()
        assert np.array_equal(out[:, :, p * N:(p + 1) * N], o1)
    assert np.abs(out[0, :, :N] - out[0, :, N:2 * N]).max() > 0
