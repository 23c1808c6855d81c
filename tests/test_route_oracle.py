"""CPU tier: the fp32 routing oracle (oracle/route_oracle.c) against fixtures produced by the
reference's shipped prebuilt `rout` binary (tools/make_route_golden.py): catchment lists / node order,
UH_S of every cell (what the binary wrote to its .uh_s files, 9 significant digits) and the
naturalised node inflows (NF1.txt = MAKE_CONVOLUTION)."""
import glob
import os
import numpy as np
import pytest
import route_lib
from vicres_b200 import route_abi
from common import ROOT

ROUTE_FIX = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "route_*.npz")))


def load_route_fixture(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "route_%s.npz" % name))
    nrow, ncol = z["dir"].shape
    hdr = {"xllcorner": str(z["hdr"][0]), "yllcorner": str(z["hdr"][1]), "cellsize": str(z["hdr"][2])}
    grid = dict(dir=z["dir"], reser=z["reser"], velo=z["velo"], diff=z["diff"], xmask=z["xmask"])
    r, c = [int(x) for x in z["station_rc"]]
    fac = route_abi.cell_factors(hdr, nrow, ncol)
    return z, grid, (c + 1, nrow - r), fac


def uh_names(z, res_ids):
    return ["uh_RES%d_OUTLT" % rid if rid > 0 else "uh_OUTLT" for rid in res_ids]


@pytest.mark.parametrize("name", ROUTE_FIX)
def test_route_oracle_matches_prebuilt_binary(name):
    z, grid, (col, row), fac = load_route_fixture(name)
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    assert o.nnodes == z["nf"].shape[0]
    res_ids = []
    for n in range(o.nnodes):
        rid, cells, uh = o.node(n)
        res_ids.append(rid)
        ref = z[uh_names(z, [rid])[0]]
        assert uh.shape == ref.shape, "catchment size of node %d" % n
        # the binary prints 9 significant digits: equal after the same rounding
        assert np.allclose(uh, ref, rtol=2e-7, atol=1e-30), "UH_S of node %d" % n
    assert res_ids[:-1] == [int(x) for x in z["res_order"][:-1]] and res_ids[-1] == 0
    # the shipped binary scales by FACTOR*0.028; today's source hard-codes 0.18308 (reservoir.f:479-480)
    flows = o.convolve(z["runoff"], z["baseflow"], cell_scale=fac * np.float32(0.028),
                       evap=(z["ev_vege"], z["ev_soil"], z["tr_vege"]), cell_factor=fac, present=z["present"])
    err = np.abs(flows - z["nf"]) / np.maximum(np.abs(z["nf"]), 1e-3)
    assert err.max() < 5e-6, "node inflow vs binary: %.3e" % err.max()
    o.close()


def test_unit_hydrographs_are_normalised():
    z, grid, (col, row), _ = load_route_fixture("rand20x16")
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    for n in range(o.nnodes):
        _, _, uh = o.node(n)
        assert np.allclose(uh.sum(axis=1), 1.0, atol=2e-6) and (uh >= 0).all()
    o.close()
