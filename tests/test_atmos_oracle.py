"""CPU tier: the plain-C restatement of the forcing preparation (oracle/atmos_oracle.c) against the
reference's own atmos[] records (tests/golden/atmos_*.npz, made by tools/make_atmos_golden.py from the
compiled reference).  The bar is bit equality: same libm, same operation order."""
import numpy as np
import pytest
import atmos_lib
from vicres_b200 import abi
from common import ATMOS_FIXTURES, load_atmos_fixture


@pytest.mark.parametrize("name", ATMOS_FIXTURES)
def test_restatement_is_bit_identical_to_the_reference(name):
    fx = load_atmos_fixture(name)
    out, flag = atmos_lib.prepare_sub(fx["opt"], fx["site"], fx["raw"])
    assert out.shape == fx["ref"].shape
    for f, fname in enumerate(abi.FORCING_FIELDS):
        bad = np.argwhere(out[f] != fx["ref"][f])
        assert bad.size == 0, "%s: %s differs at %d values, first index %s: %r vs %r" % (
            name, fname, len(bad), bad[0], out[f][tuple(bad[0])], fx["ref"][f][tuple(bad[0])])
    if fx["snowflag"] is not None:           # SNOW_STEP < TIME_STEP fixtures carry atmos->snowflag[NR]
        assert np.array_equal(flag, fx["snowflag"])


def test_fixture_set_covers_the_branches():
    """hour offsets of both signs, both layouts, a leap year, polar day, every VP_ITER value."""
    metas = {n: load_atmos_fixture(n) for n in ATMOS_FIXTURES}
    hoi = np.concatenate([np.rint((m["site"]["time_zone_lng"] - m["site"]["lng"]) / 15.0) for m in metas.values()])
    assert (hoi > 0).any() and (hoi < 0).any() and (hoi == 0).any()
    assert {m["opt"].layout for m in metas.values()} == {0, 1}
    assert {m["opt"].vp_iter for m in metas.values()} == {0, 1, 2, 3}
    assert any(np.abs(m["site"]["lat"]).max() > 67 for m in metas.values())
