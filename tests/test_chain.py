"""The whole north-star chain against the reference's own two programs run back to back (fixture
tests/golden/chain_case, tools/make_chain_golden.py: vicNl compiled in place -> fluxes files -> the shipped
prebuilt rout binary): parameter files -> readers -> cell step -> result writer / in-memory hand-off ->
unit hydrographs -> convolution -> reservoirs -> station discharge.

CPU tier: with the oracles (step oracle bit-exact, so the flux file is byte-identical; routing restatement at
fp32 rounding level).  GPU tier: the same chain entirely through the C ABI of libvicres.so."""
import filecmp
import glob
import os
import numpy as np
import pytest
import oracle_lib
import route_lib
from vicres_b200 import abi, build, couple, files, route_abi
from common import GOLDEN

CASE = os.path.join(GOLDEN, "chain_case")


def load_case():
    z = np.load(os.path.join(CASE, "chain.npz"))
    lib, cells, meta = files.read_params(os.path.join(CASE, "soil.txt"), os.path.join(CASE, "veglib.txt"),
                                         os.path.join(CASE, "veg.txt"), nlayer=2, nbands=1, root_zones=3)
    mst, mrt, wind_h = [float(x) for x in z["meta_f"]]
    opt = abi.default_options(2, 1, 24, max_snow_temp=mst, min_rain_temp=mrt, wind_h=wind_h)
    names = abi.default_out_vars(2)
    nrow, ncol = z["dir"].shape
    hdr = {"xllcorner": str(z["hdr"][0]), "yllcorner": str(z["hdr"][1]), "cellsize": str(z["hdr"][2])}
    grid = dict(dir=z["dir"], reser=z["reser"], velo=z["velo"], diff=z["diff"], xmask=z["xmask"])
    r, c = [int(x) for x in z["station_rc"]]
    return z, lib, cells, meta, opt, names, hdr, grid, (c + 1, nrow - r)


def route_and_check(z, R, inp, hdr, tol_nf, tol_vol, tol_flow):
    nrow, ncol = z["dir"].shape
    fac = route_abi.cell_factors(hdr, nrow, ncol)
    kw = dict(runoff=inp["runoff"], baseflow=inp["baseflow"], cell_scale=fac * np.float32(0.028),
              evap=(inp["ev_vege"], inp["ev_soil"], inp["tr_vege"]), cell_factor=fac, present=inp["present"])
    nat = R.convolve_in(**kw)[0]
    err = np.abs(nat - z["nf"]) / np.maximum(np.abs(z["nf"]), 1e-2 * np.abs(z["nf"]).max())
    assert err.max() < tol_nf, "naturalised inflows vs the binary: %.3e" % err.max()
    ids = [R.node(n)[0] for n in range(R.nnodes - 1)]
    res = [route_abi.read_reservoir_file(os.path.join(CASE, "res_par", "res%d.txt" % i), 2000, 1, nat.shape[1]) for i in ids]
    out = R.reservoirs(nat, res if isinstance(R, route_lib.RouteOracle) else [res], 2000, 1, legacy=True)
    if not isinstance(R, route_lib.RouteOracle):
        out = {k: v[0] for k, v in out.items()}
    for j, rid in enumerate(ids):
        ref = z["resday_%d" % rid]
        vol = out["vol"][j][:-1]
        scale = np.maximum(np.abs(ref[:, 0]), 1e-2 * np.abs(ref[:, 0]).max() + 1e-3)
        assert (np.abs(np.maximum(vol, 0) - ref[:, 0]) / scale).max() < tol_vol, "storage of reservoir %d" % rid
    ref = z["station_flow"]
    err = np.abs(out["flow"] - ref) / np.maximum(np.abs(ref), 1e-2 * np.abs(ref).max())
    assert err.max() < tol_flow, "regulated station discharge vs the binary: %.3e" % err.max()


def test_chain_with_oracles_matches_reference_programs(tmp_path):
    build.build()
    z, lib, cells, meta, opt, names, hdr, grid, (col, row) = load_case()
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(z["tair0"])
    rc, out, _ = o.step(z["month"], z["doy"], z["forcing"])
    assert rc == 0
    o.close()
    dmy = z["dmy"]
    files.write_results(str(tmp_path), "fluxes", meta["lat"], meta["lng"], dmy[:, 0], dmy[:, 1], dmy[:, 2], out,
                        cols=np.arange(21))
    kept = glob.glob(os.path.join(CASE, "fluxes_*"))[0]
    assert filecmp.cmp(os.path.join(str(tmp_path), os.path.basename(kept)), kept, shallow=False)
    nrow, ncol = z["dir"].shape
    a = couple.inputs_from_files(str(tmp_path) + "/fluxes_", hdr, nrow, ncol, out.shape[1])
    b = couple.inputs_from_outputs(out, names, meta["lat"], meta["lng"], hdr, nrow, ncol)
    for k in a:
        assert np.array_equal(a[k], b[k]), "in-memory hand-off differs from the file round trip: %s" % k
    assert int(a["present"].sum()) == cells["lat"].size
    R = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    route_and_check(z, R, b, hdr, tol_nf=5e-6, tol_vol=1e-5, tol_flow=2e-5)
    R.close()


@pytest.mark.gpu
def test_chain_on_gpu_matches_reference_programs():
    """everything through libvicres.so: CUDA cell step -> in-memory hand-off -> CUDA routing and reservoirs"""
    from vicres_b200 import model, route
    build.build()
    z, lib, cells, meta, opt, names, hdr, grid, (col, row) = load_case()
    m = model.VicRes(opt, lib, cells)
    m.init_state(z["tair0"])
    out, status = m.step(z["month"], z["doy"], z["forcing"])
    assert not status.any()
    m.close()
    nrow, ncol = z["dir"].shape
    inp = couple.inputs_from_outputs(out, names, meta["lat"], meta["lng"], hdr, nrow, ncol)
    R = route.Route(grid, col, row, z["uh_box"])
    # the step's libm-level differences move a few printed 4th decimals; the routing sums smooth them out
    route_and_check(z, R, inp, hdr, tol_nf=5e-5, tol_vol=1e-4, tol_flow=1e-4)
    R.close()
