"""ORACLE = TEST INFRASTRUCTURE.  numpy restatement of the toolbox's calibration objectives
(toolbox/calibration/basin_calibration_eNSGAII.py:86-175): NSE, TRMSE (Box-Cox lambda 0.3), MSDE, ROCE.
Pinned against the reference's own functions executed from its source (tests/golden/objectives.npz,
tools/make_objectives_golden.py)."""
import numpy as np


def objectives(obs, sim, handle_negatives=True):
    """sim: [nsets, ndays]; returns [nsets, 4]"""
    obs = np.asarray(obs, dtype=np.float64)
    out = np.zeros((sim.shape[0], 4))
    for s in range(sim.shape[0]):
        x = np.asarray(sim[s], dtype=np.float64)
        v = np.abs(x) if handle_negatives else x
        oa = np.abs(obs) if handle_negatives else obs
        out[s, 0] = 1 - np.sum((v - obs) ** 2) / np.sum((obs - np.mean(obs)) ** 2)
        with np.errstate(invalid="ignore"):
            zs = (np.power(v + 1, 0.3) - 1) / 0.3
            zo = (np.power(oa + 1, 0.3) - 1) / 0.3
        out[s, 1] = np.sqrt(np.mean((zs - zo) ** 2))
        out[s, 2] = np.mean((np.diff(obs) - np.diff(v)) ** 2) if obs.size > 1 else np.nan
        so = np.sum(obs)
        out[s, 3] = (np.inf if np.sum(v) != 0 else np.nan) if so == 0 else np.abs(np.sum(v) - so) / so
    return out
