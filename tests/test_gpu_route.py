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
