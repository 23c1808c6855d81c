"""GPU tier: the CUDA routing pass through the C ABI against (1) the fixtures made by the reference's
prebuilt binary and (2) the fp32 oracle on a larger seeded network.  UH_S: the device powf/expf/sqrtf
differ from glibc's in the last bit, so 2e-6 relative; node inflows given the same UH_S would be
bit-identical (same fp32 accumulation order, -fmad=false), with the device UH_S 1e-5 relative."""
import numpy as np
import pytest
import route_lib
from vicres_b200 import build, route_abi
from test_route_oracle import ROUTE_FIX, load_route_fixture, uh_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Route():
    import torch
    assert torch.cuda.is_available()
    build.build()
    from vicres_b200 import route
    return route.Route


@pytest.mark.parametrize("name", ROUTE_FIX)
def test_cuda_route_matches_prebuilt_binary(Route, name):
    z, grid, (col, row), fac = load_route_fixture(name)
    r = Route(grid, col, row, z["uh_box"])
    assert r.nnodes == z["nf"].shape[0]
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    for n in range(r.nnodes):
        rid, cells, uh = r.node(n)
        orid, ocells, ouh = o.node(n)
        assert rid == orid and np.array_equal(cells, ocells)
        ref = z[uh_names(z, [rid])[0]]
        assert np.allclose(uh, ref, rtol=2e-6, atol=1e-12), "UH_S node %d: %.3e" % (n, np.abs(uh - ref).max())
    if (z["reser"] == 9999).any():                 # covered by test_cuda_open_water_evaporation
        with pytest.raises(Exception, match="9999"):
            r.convolve(z["runoff"], z["baseflow"])
        r.close(); o.close()
        return
    flows = r.convolve(z["runoff"], z["baseflow"], cell_scale=fac * np.float32(0.028),
                       evap=(z["ev_vege"], z["ev_soil"], z["tr_vege"]), cell_factor=fac, present=z["present"])
    err = np.abs(flows - z["nf"]) / np.maximum(np.abs(z["nf"]), 1e-3)
    assert err.max() < 1e-5, "node inflow vs binary %.3e" % err.max()
    # today's source: scalar 0.18308, no per-cell factor
    f2 = r.convolve(z["runoff"], z["baseflow"])
    o2 = o.convolve(z["runoff"], z["baseflow"])
    assert np.allclose(f2, o2, rtol=1e-5, atol=1e-6)
    r.close(); o.close()


def test_cuda_route_large_network_vs_oracle(Route):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from make_route_golden import random_tree
    rng = np.random.default_rng(5)
    nrow, ncol, ndays = 60, 80, 400
    d, out = random_tree(nrow, ncol, rng, fill=0.7)
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    for i, k in enumerate(rng.choice(len(active), 12, replace=False)):
        reser[tuple(active[k])] = i + 1
    grid = dict(dir=d, reser=reser, velo=rng.uniform(0.8, 2.5, (nrow, ncol)).astype(np.float32),
                diff=rng.uniform(400, 1500, (nrow, ncol)).astype(np.float32),
                xmask=rng.uniform(15000, 30000, (nrow, ncol)).astype(np.float32))
    box = rng.dirichlet(np.ones(12)).astype(np.float32)
    runoff = (rng.gamma(0.5, 4.0, (nrow * ncol, ndays)) * (rng.uniform(size=(nrow * ncol, ndays)) < 0.4)).astype(np.float32)
    base = np.abs(rng.normal(1.0, 0.5, (nrow * ncol, ndays))).astype(np.float32)
    r = Route(grid, out[1] + 1, nrow - out[0], box)
    o = route_lib.RouteOracle(grid, out[1] + 1, nrow - out[0], box)
    assert r.nnodes == o.nnodes and r.nnodes >= 2
    tot = 0
    for n in range(r.nnodes):
        rid, cells, uh = r.node(n)
        orid, ocells, ouh = o.node(n)
        assert rid == orid and np.array_equal(cells, ocells)
        assert np.allclose(uh, ouh, rtol=5e-6, atol=1e-12)
        tot += len(cells)
    assert tot > 2000
    f = r.convolve(runoff, base)
    fo = o.convolve(runoff, base)
    assert np.allclose(f, fo, rtol=2e-5, atol=1e-5)
    # mass: with normalised UH_S the routed volume equals the generated volume once the tail has drained
    print("routed cells %d, nodes %d, convolve kernel %.3f ms" % (tot, r.nnodes, r.last_kernel_ms()))
    r.close(); o.close()


import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tools"))
from test_gpu_route_helpers import random_reservoir  # noqa: E402


@pytest.mark.parametrize("legacy", [False, True])
def test_cuda_reservoir_day_loop_bit_exact_vs_oracle(Route, legacy):
    """all seven rules, chained reservoirs, 8 parameter sets in one launch; with the oracle's UH_R injected
    (the reference's read-an-existing-.uh_r path) every series must equal the serial restatement bit for bit"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from make_route_golden import random_tree
    rng = np.random.default_rng(17)
    nrow, ncol, ndays = 30, 40, 900
    d, out = random_tree(nrow, ncol, rng, fill=0.75)
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    for i, k in enumerate(rng.choice(len(active), 40, replace=False)):
        reser[tuple(active[k])] = i + 1
    grid = dict(dir=d, reser=reser, velo=rng.uniform(0.8, 2.5, (nrow, ncol)).astype(np.float32),
                diff=rng.uniform(400, 1500, (nrow, ncol)).astype(np.float32),
                xmask=rng.uniform(15000, 30000, (nrow, ncol)).astype(np.float32))
    box = rng.dirichlet(np.ones(12)).astype(np.float32)
    col, row = int(out[1]) + 1, int(nrow - out[0])
    r = Route(grid, col, row, box)
    o = route_lib.RouteOracle(grid, col, row, box)
    nres = r.nnodes - 1
    assert nres >= 10
    chained = 0
    for n in range(nres):
        ds, tg, uh = r.res_info(n)
        ods, otg, ouh = o.res_info(n)
        assert (ds, tg) == (ods, otg)
        assert np.allclose(uh, ouh, rtol=2e-6, atol=1e-12)
        chained += tg < nres
        r.set_uh_r(n, ouh)
    assert chained >= 2
    natural = (rng.gamma(0.6, 30.0, (r.nnodes, ndays)) * (rng.uniform(size=(r.nnodes, ndays)) < 0.7)).astype(np.float32)
    natural[:, 200:230] -= 5.0                      # negative naturalised inflow (evaporation larger than runoff)
    nsets = 8
    sets = [[random_reservoir(rng, int((n + s) % 7), ndays) for n in range(nres)] for s in range(nsets)]
    out_gpu = r.reservoirs(natural, sets, 2000, 3, legacy=legacy)
    for s in range(nsets):
        ref = o.reservoirs(natural, sets[s], 2000, 3, legacy=legacy)
        assert np.array_equal(out_gpu["flow"][s], ref["flow"], equal_nan=True), "station flow, set %d" % s
        for k, _ in route_abi.RES_SERIES:
            assert np.array_equal(out_gpu[k][s], ref[k], equal_nan=True), "%s, set %d: %d values differ" % (
                k, s, (out_gpu[k][s] != ref[k]).sum())
    assert r.last_kernel_ms() > 0
    r.close(); o.close()


def test_step_by_step_equals_one_run(Route):
    """The reference's STEPBYSTEP mode (reservoir.f:1142-1150, 1596-1675): the day loop run one day -- or one block of days --
    per call with VOL / FLOWOUT carried over must give, bit for bit, what the whole period gives in one call: every
    series of every reservoir (chained ones included: their inflow is the upstream releases routed through UH_R over up to
    107 earlier days) and the station discharge."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from make_route_golden import random_tree
    rng = np.random.default_rng(24)           # (a seed whose random tree covers the grid: 540 active cells)
    nrow, ncol, ndays = 24, 30, 260
    d, out = random_tree(nrow, ncol, rng, fill=0.75)
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    for i, k in enumerate(rng.choice(len(active), 25, replace=False)):
        reser[tuple(active[k])] = i + 1
    grid = dict(dir=d, reser=reser, velo=np.full((nrow, ncol), 1.5, np.float32), diff=np.full((nrow, ncol), 800.0, np.float32),
                xmask=np.full((nrow, ncol), 20000.0, np.float32))
    box = rng.dirichlet(np.ones(12)).astype(np.float32)
    r = Route(grid, int(out[1]) + 1, int(nrow - out[0]), box)
    nres = r.nnodes - 1
    natural = (rng.gamma(0.6, 30.0, (r.nnodes, ndays)) * (rng.uniform(size=(r.nnodes, ndays)) < 0.7)).astype(np.float32)
    sets = [[random_reservoir(rng, int((n + s) % 7), ndays) for n in range(nres)] for s in range(3)]
    whole = r.reservoirs(natural, sets, 2001, 2)
    # blocks of days, then day by day
    st = None
    for d0, d1 in [(0, 1), (1, 2), (2, 120), (120, 121), (121, ndays)]:
        st = r.reservoirs(natural, sets, 2001, 2, days=(d0, d1), state=st)
    for k in whole:
        assert np.array_equal(whole[k], st[k], equal_nan=True), k
    st = None
    for dd in range(40):
        st = r.reservoirs(natural, sets, 2001, 2, days=(dd, dd + 1), state=st)
    for k, ext in route_abi.RES_SERIES:
        assert np.array_equal(whole[k][:, :, :40 + ext], st[k][:, :, :40 + ext], equal_nan=True), k
    assert np.array_equal(whole["flow"][:, :40], st["flow"][:, :40], equal_nan=True)
    r.close()


def test_cuda_open_water_evaporation(Route, tmp_path):
    """reservoir-surface cells: device MAKE_CONVOLUTION incl. Penman table evaporation against the binary's
    NF1.txt (1e-5: device UH_S) and against the oracle (res_evaporation uses no UH: bit-exact)"""
    from test_route_oracle import evap_inputs
    z, grid, (col, row), fac = load_route_fixture("evap14x16")
    r = Route(grid, col, row, z["uh_box"])
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    ids = [r.node(n)[0] for n in range(r.nnodes - 1)]
    kw = evap_inputs(z, grid, fac, ids, tmp_path)
    flows, resev = r.convolve_in(**kw)
    oflows, oresev = o.convolve_in(**kw)
    err = np.abs(flows - z["nf"]) / np.maximum(np.abs(z["nf"]), 1e-3)
    assert err.max() < 1e-5, "node inflow vs binary %.3e" % err.max()
    assert np.array_equal(resev, oresev) and resev.max() > 0
    assert np.allclose(flows, oflows, rtol=1e-5, atol=1e-6)
    r.close(); o.close()


@pytest.mark.parametrize("name", ["toy", "rand20x16"])
def test_cuda_reservoirs_match_prebuilt_binary(Route, name, tmp_path):
    """end to end on the device (its own UH_S, UH_R and convolution) against the binary's reservoir files:
    the unit hydrographs differ from glibc's in the last bit, so tolerances (relative to max(|value|, 1 % of
    the column's range)) instead of bit equality; bit equality of the day loop itself is the test above"""
    from test_route_oracle import binary_columns, fixture_reservoirs
    z, grid, (col, row), fac = load_route_fixture(name)
    r = Route(grid, col, row, z["uh_box"])
    ids = [r.node(n)[0] for n in range(r.nnodes - 1)]
    ndays = int(z["ndays"])
    flows = r.convolve(z["runoff"], z["baseflow"], cell_scale=fac * np.float32(0.028),
                       evap=(z["ev_vege"], z["ev_soil"], z["tr_vege"]), cell_factor=fac, present=z["present"])
    res = fixture_reservoirs(z, ids, tmp_path, ndays)
    out = r.reservoirs(flows, [res], 2000, 1, legacy=True)
    one = {k: v[0] for k, v in out.items()}
    for j, rid in enumerate(ids):
        ref = z["resday_%d" % rid]
        mine = binary_columns(one, j, res[j])
        mine[:, [0, 3, 5, 8]] = np.maximum(mine[:, [0, 3, 5, 8]], 0)
        mine[:, 7] = np.where(mine[:, 2] > 0, np.minimum(mine[:, 6], np.float32(res[j]["q_design"])), mine[:, 7])
        scale = np.maximum(np.abs(ref), 1e-2 * np.abs(ref).max(axis=0, keepdims=True) + 1e-2)
        err = np.abs(mine - ref) / scale
        assert err[:, :5].max() < 2e-5, "reservoir %d storage/level: %.3e" % (rid, err[:, :5].max())
        # releases are differences of storages divided by 86.4: one fp32 ulp of a 1e5 storage is 1e-4 m3/s
        assert err[:, 5:].max() < 1e-3, "reservoir %d flows/energy: %.3e" % (rid, err[:, 5:].max())
    err = np.abs(one["flow"] - z["station_flow"]) / np.maximum(np.abs(z["station_flow"]), 1e-2)
    assert err.max() < 2e-5
    r.close()


def test_cuda_station_without_reservoirs(Route):
    """a catchment with no reservoir above the station: one node, the reservoir pass is the clamp of the
    naturalised flow (reservoir.f:1589-1593)"""
    z, grid, (col, row), fac = load_route_fixture("rand12x10")
    r = Route(grid, col, row, z["uh_box"])
    assert r.nnodes == 1
    nat = r.convolve(z["runoff"], z["baseflow"])
    nat[0, 5] = -3.0
    out = r.reservoirs(nat, [[]], 2000, 1)
    assert out["flow"].shape == (1, nat.shape[1]) and out["flow"][0, 5] == 0
    assert np.array_equal(np.maximum(nat[0], 0), out["flow"][0])
    r.close()


def test_device_flux_handoff_equals_the_text_round_trip(Route):
    """vr_route_fluxes_from_step (device) against the host path vr_round_decimals + conversion to float32: random values,
    exact decimal ties and their neighbours, values around zero; then the convolution fed from device memory against
    the one fed from the host arrays."""
    import torch
    from vicres_b200 import couple, files, route
    rng = np.random.default_rng(3)
    C, nd = 700, 130
    names = list(couple.ROUTE_VARS) + ["SWE"]
    x = rng.normal(0, 30, (len(names), nd, C))
    k = rng.integers(-200000, 200000, (nd, C))
    x[0] = (k + 0.5) / 1e4                                   # decimals that tie in the fifth place (as doubles: just off the tie)
    x[1] = np.nextafter((k + 0.5) / 1e4, np.inf)
    x[2] = np.nextafter((k + 0.5) / 1e4, -np.inf)
    x[3] = rng.uniform(-1, 1, (nd, C)) * 1e-4
    x[4] = k / 8.0 + 0.03125                                 # exactly representable binary fractions: true ties of the exact value
    dev = torch.device("cuda", 0)
    out_dev = torch.from_numpy(x).to(dev)
    ngrid = 900
    gmap = rng.permutation(ngrid)[:C].astype(np.int64)
    gmap[::50] = -1
    cols, present = route.fluxes_from_step_device(out_dev, names, gmap, ngrid)
    want = files.round_decimals(x[:8], 4).astype(np.float32)          # [8, nd, C]
    on = gmap >= 0
    for j, key in enumerate(couple.KEYS):
        got = cols[key].cpu().numpy()                        # [ngrid, nd]
        assert np.array_equal(got[gmap[on]], want[j][:, on].T), key
        off = np.ones(ngrid, bool); off[gmap[on]] = False
        assert not got[off].any()
    p = present.cpu().numpy()
    assert p.sum() == on.sum() and p[gmap[on]].all()


def test_convolution_from_device_columns(Route):
    import torch
    from vicres_b200 import route
    z, grid, (col, row), fac = load_route_fixture("rand20x16")
    r = Route(grid, col, row, z["uh_box"])
    ndays = int(z["ndays"])
    host = r.convolve_in(runoff=z["runoff"], baseflow=z["baseflow"], evap=(z["ev_vege"], z["ev_soil"], z["tr_vege"]),
                         cell_factor=fac, cell_scale=fac * np.float32(0.028), present=z["present"])[0]
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(z[k], dtype=np.float32)).to(dev) for k in ("runoff", "baseflow", "ev_vege", "ev_soil", "tr_vege")}
    pres = torch.from_numpy(np.ascontiguousarray(z["present"], dtype=np.uint8)).to(dev)
    ptrs = {k: v.data_ptr() for k, v in t.items()}
    ptrs["present"] = pres.data_ptr()
    got = r.convolve_in_device(ndays, ptrs, cell_factor=fac, cell_scale=fac * np.float32(0.028))
    assert np.array_equal(got, host)
    r.close()
