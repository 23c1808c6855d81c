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
    if (z["reser"] == 9999).any():                 # needs the evaporation inputs: see the open-water test below
        o.close()
        return
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


def fixture_reservoirs(z, res_ids, tmp_path, ndays):
    res = []
    for rid in res_ids:
        p = tmp_path / ("res%d.txt" % rid)
        p.write_text(str(z["restxt_%d" % rid]))
        res.append(route_abi.read_reservoir_file(str(p), 2000, 1, ndays))
    return res


def binary_columns(out, j, res):
    """the nine columns of reservoir_<id>_<station>.day (write_routines.f:172-200) from the day-loop arrays"""
    n = out["flowin"].shape[1]
    full = np.full(n, np.float32(res["v_live"]) + np.float32(res["v_dead"]), dtype=np.float32)
    full[2000 + np.arange(1, n + 1) // 365 < res["ope_year"]] = 0          # VRESER = 0 until commissioned (:1521)
    return np.stack([out["vol"][j][:-1], out["vol"][j][1:], full, out["hho"][j][:-1], out["hho"][j][1:],
                     out["flowin"][j], out["flowout"][j], out["turbine"][j], out["energy"][j]], axis=1)


@pytest.mark.parametrize("name", [n for n in ROUTE_FIX if n != "rand12x10"])
def test_reservoir_day_loop_matches_prebuilt_binary(name, tmp_path):
    """MAKE_RESERVOIR_UH + MAKE_CONVOLUTIONRS against the shipped binary, BIT FOR BIT: UH_R files, every
    column of every reservoir_*.day and the regulated station discharge.  The binary predates three details
    of today's rule-curve code (vr_res_options.legacy_rule_curve); it is run on what both share: rules 1, 2
    and not-yet-commissioned reservoirs, chained through UH_R."""
    z, grid, (col, row), fac = load_route_fixture(name)
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    ids = [o.node(n)[0] for n in range(o.nnodes - 1)]
    ndays = int(z["ndays"])
    for n, rid in enumerate(ids):
        ds, target, uh_r = o.res_info(n)
        key = "uhr_RES%d_RES%d" % (rid, ds) if target < o.nnodes - 1 else "uhr_RES%d_OUTLT" % rid
        assert np.array_equal(uh_r, z[key]), key
    res = fixture_reservoirs(z, ids, tmp_path, ndays)
    out = o.reservoirs(z["nf"], res, 2000, 1, legacy=True)
    for j, rid in enumerate(ids):
        ref = z["resday_%d" % rid]
        mine = binary_columns(out, j, res[j])
        # the writer clamps negatives to zero (write_routines.f:181-195)
        mine[:, [0, 3, 5, 8]] = np.maximum(mine[:, [0, 3, 5, 8]], 0)
        # the binary's writer prints min(FLOWOUT, QRESER) in the turbine column (seen on spill and on
        # environmental-flow days); FLOWOUT_TURB itself is pinned through the energy column it drives
        mine[:, 7] = np.where(mine[:, 2] > 0, np.minimum(mine[:, 6], np.float32(res[j]["q_design"])), mine[:, 7])
        assert np.array_equal(mine, ref), "reservoir %d: columns differing %s" % (rid, (mine != ref).sum(0))
    assert np.array_equal(out["flow"], z["station_flow"])
    o.close()


def test_current_rule_curve_differs_from_binary_only_where_documented(tmp_path):
    """today's reservoir.f (80 % bands, clamp-before-spill) is a different operating policy: same inflow to
    the head reservoir on day 1, different storage trajectory afterwards."""
    z, grid, (col, row), fac = load_route_fixture("toy")
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    ids = [o.node(n)[0] for n in range(o.nnodes - 1)]
    res = fixture_reservoirs(z, ids, tmp_path, int(z["ndays"]))
    new = o.reservoirs(z["nf"], res, 2000, 1, legacy=False)
    old = o.reservoirs(z["nf"], res, 2000, 1, legacy=True)
    assert np.array_equal(new["flowin"][0], old["flowin"][0])          # head reservoir: naturalised inflow only
    assert not np.array_equal(new["vol"][0], old["vol"][0])
    assert np.isfinite(new["flow"]).all() and (new["flow"] >= 0).all()
    # quirk kept: today's source clamps VOL before adding the spill, so FLOWOUT never exceeds the turbine flow
    assert np.array_equal(new["flowout"], new["turbine"])
    o.close()


def evap_inputs(z, grid, fac, nodes_res_ids, tmp_path):
    nrow, ncol = z["dir"].shape
    hdr = {"xllcorner": str(z["hdr"][0]), "yllcorner": str(z["hdr"][1]), "cellsize": str(z["hdr"][2])}
    lat, lon = route_abi.cell_coords(hdr, nrow, ncol)
    res = fixture_reservoirs(z, nodes_res_ids, tmp_path, int(z["ndays"]))
    return dict(runoff=z["runoff"], baseflow=z["baseflow"], cell_scale=fac * np.float32(0.028),
                evap=(z["ev_vege"], z["ev_soil"], z["tr_vege"]), cell_factor=fac, present=z["present"],
                met=(z["rh"], z["tair"], z["wind"]), cell_lat=lat, cell_lon=lon,
                node_ope_year=[r["ope_year"] for r in res] + [0], start_year=2000, start_month=1)


def test_open_water_evaporation_matches_prebuilt_binary(tmp_path):
    """reservoir-surface cells (9999): Penman table evaporation subtracted in MAKE_CONVOLUTION
    (reservoir.f:410-475,496-499) against the binary's NF1.txt; without it the node is 3 % off"""
    z, grid, (col, row), fac = load_route_fixture("evap14x16")
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    ids = [o.node(n)[0] for n in range(o.nnodes - 1)]
    n9 = [int((z["reser"].ravel()[o.node(n)[1]] == 9999).sum()) for n in range(o.nnodes)]
    assert sum(n9) >= 2
    kw = evap_inputs(z, grid, fac, ids, tmp_path)
    flows, resev = o.convolve_in(**kw)
    err = np.abs(flows - z["nf"]) / np.maximum(np.abs(z["nf"]), 1e-3)
    assert err.max() < 5e-6, "node inflow vs binary: %.3e" % err.max()
    kw.pop("met")
    dry, _ = o.convolve_in(**kw)
    err0 = np.abs(dry - z["nf"]) / np.maximum(np.abs(z["nf"]), 1e-3)
    assert err0.max() > 1e-3                       # the fixture does exercise the evaporation
    assert (resev[np.array(n9) > 0] > 0).any() and (resev[np.array(n9) == 0] == 0).all()
    o.close()


def test_memoised_catchment_search_equals_the_path_by_path_search(tmp_path):
    """host part of vr_route_create: one downstream walk per cell with memoisation gives the same cell lists, in the
    same order, as walking every cell's path again for every node (tests/hostsim/route_graph_check.cu; random
    direction grids incl. cycles)"""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "route_graph_check")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim", "route_graph_check.cu")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-w", "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "0 of" in r.stdout, r.stdout[-500:]
