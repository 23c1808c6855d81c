"""Calibration objectives against golden vectors made by the reference's own
functions (tests/golden/objectives.npz).  Tolerance 1e-10 relative: the reductions differ in summation order
(numpy pairwise vs a block tree) and pow() by an ulp."""
import os
import numpy as np
import pytest
import objectives_oracle
from common import GOLDEN

Z = np.load(os.path.join(GOLDEN, "objectives.npz"))


def close(a, b):
    both_nan = np.isnan(a) & np.isnan(b)
    return np.all(both_nan | (np.abs(a - b) <= 1e-10 * np.maximum(np.abs(b), 1e-30)))


@pytest.mark.parametrize("h, handle", [(0, True), (1, False)])
def test_objectives_oracle_matches_reference_functions(h, handle):
    got = objectives_oracle.objectives(Z["obs"], Z["sim"], handle)
    assert close(got, Z["ref"][h])
    assert np.isnan(Z["ref"][1]).any() and not np.isnan(Z["ref"][0]).any()      # raw negative flows -> NaN TRMSE, as the toolbox


@pytest.mark.gpu
@pytest.mark.parametrize("h, handle", [(0, True), (1, False)])
def test_cuda_objectives_match_reference_functions(h, handle):
    from vicres_b200 import build, route
    build.build()
    got = route.objectives(Z["sim"], Z["obs"], handle)
    assert close(got, Z["ref"][h]), np.abs(got - Z["ref"][h]).max()
