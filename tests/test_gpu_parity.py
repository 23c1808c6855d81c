"""GPU tier (-m gpu): the CUDA path, called through the C ABI of libvicres.so, against the golden
fixtures (outputs of the reference itself) and against the oracle on seeded inputs; plus
size-independent properties at larger sizes.  Tolerance: 1e-9 relative per cell-timestep
(BASELINE.json north_star), see tests/common.py."""
import numpy as np
import pytest
import oracle_lib
from vicres_b200 import abi, build, synth_soa
from common import FIXTURES, RTOL, check_outputs, load_fixture, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def VicRes():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    build.build()
    from vicres_b200 import model
    return model.VicRes


@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_matches_reference_golden(VicRes, name):
    fx = load_fixture(name)
    m = VicRes(fx["opt"], fx["lib"], fx["cells"])
    m.init_state(fx["tair0"])
    out, status = m.step(fx["month"], fx["doy"], fx["forcing"], snowflag=fx["snowflag"])
    assert not status.any()
    check_outputs(out, fx["ref"], fx["names"], what="%s CUDA vs reference" % name)
    assert m.num_launches >= 2
    m.close()


def test_block_split_is_bit_identical(VicRes):
    fx = load_fixture("daily_snow_bands")
    m = VicRes(fx["opt"], fx["lib"], fx["cells"])
    m.init_state(fx["tair0"])
    whole, _ = m.step(fx["month"], fx["doy"], fx["forcing"])
    m.init_state(fx["tair0"])
    parts = []
    for a, b in [(0, 1), (1, 130), (130, 500)]:
        o, _ = m.step(fx["month"][a:b], fx["doy"][a:b], np.ascontiguousarray(fx["forcing"][:, a:b]))
        parts.append(o)
    assert np.array_equal(whole, np.concatenate(parts, axis=1))
    m.close()


@pytest.mark.parametrize("name", ["daily_snow_bands", "daily_snowstep3"])
def test_record_by_record_shim_equals_the_block_call(VicRes, name):
    """vr_full_energy_compat: the reference's per-record `full_energy(); put_data();` pair, bit-identical to vr_step_block;
    a record out of sequence is refused"""
    from vicres_b200 import abi
    fx = load_fixture(name)
    n = 60
    m = VicRes(fx["opt"], fx["lib"], fx["cells"])
    m.init_state(fx["tair0"])
    NS = fx["forcing"].shape[1] // len(fx["month"])
    sf = fx["snowflag"]
    whole, _ = m.step(fx["month"][:n], fx["doy"][:n], np.ascontiguousarray(fx["forcing"][:, :n * NS]),
                      snowflag=None if sf is None else np.ascontiguousarray(sf[:n]))
    m.init_state(fx["tair0"])
    for r in range(n):
        o, status = m.full_energy(r, fx["month"][r], fx["doy"][r], fx["forcing"][:, r * NS:(r + 1) * NS],
                                  None if sf is None else sf[r])
        assert not status.any() and np.array_equal(o, whole[:, r]), r
    with pytest.raises(Exception, match="next record"):
        m.full_energy(n + 3, 1, 1, fx["forcing"][:, :NS])
    m.close()


def test_state_roundtrip(VicRes):
    fx = load_fixture("sub3h_snow")
    m = VicRes(fx["opt"], fx["lib"], fx["cells"])
    m.init_state(fx["tair0"])
    h = 400
    m.step(fx["month"][:h], fx["doy"][:h], np.ascontiguousarray(fx["forcing"][:, :h]))
    blob = m.get_state()
    rest, _ = m.step(fx["month"][h:], fx["doy"][h:], np.ascontiguousarray(fx["forcing"][:, h:]))
    m2 = VicRes(fx["opt"], fx["lib"], fx["cells"])
    m2.set_state(blob)
    rest2, _ = m2.step(fx["month"][h:], fx["doy"][h:], np.ascontiguousarray(fx["forcing"][:, h:]))
    assert np.array_equal(rest, rest2)
    check_outputs(rest, fx["ref"][:, h:], fx["names"], what="restart")
    m.close(); m2.close()


def _tile(cells, rep):
    """replicate every cell `rep` times (cell-major: copies of cell c are adjacent)"""
    out = {}
    C = len(cells["lat"])
    idx = np.repeat(np.arange(C), rep)
    for k in abi.CELL_SCALARS:
        out[k] = cells[k][idx]
    for k in abi.CELL_LAYER + abi.CELL_BAND:
        out[k] = cells[k][:, idx]
    tp = cells["tile_ptr"]
    tiles = np.concatenate([np.arange(tp[c], tp[c + 1]) for c in idx])
    nt = (tp[1:] - tp[:-1])[idx]
    out["tile_ptr"] = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    for k in ["tile_class", "tile_Cv", "tile_root", "tile_LAI", "tile_albedo", "tile_vegcover"]:
        out[k] = cells[k][tiles]
    for k in ["zwt_zwt", "zwt_moist"]:
        if cells.get(k) is not None:
            out[k] = np.ascontiguousarray(cells[k][:, :, idx])
    return out, idx


def test_parameter_set_axis_shares_forcing(VicRes):
    """Ensemble layout (BASELINE config 4): P parameter sets x N cells with a forcing-column map."""
    fx = load_fixture("daily_warm")
    rep = 3
    cells, idx = _tile(fx["cells"], rep)
    rng = np.random.default_rng(1)
    cells["b_infilt"] = cells["b_infilt"] * rng.uniform(0.5, 1.5, len(idx))
    cells["Ds"] = np.clip(cells["Ds"] * rng.uniform(0.5, 1.5, len(idx)), 1e-3, 0.95)
    names = fx["names"]
    m = VicRes(fx["opt"], fx["lib"], cells, forcing_map=idx, n_forcing_cols=len(fx["tair0"]))
    m.init_state(fx["tair0"])
    out, _ = m.step(fx["month"], fx["doy"], fx["forcing"])
    # same thing with the forcing physically replicated
    m2 = VicRes(fx["opt"], fx["lib"], cells)
    m2.init_state(fx["tair0"][idx])
    out2, _ = m2.step(fx["month"], fx["doy"], np.ascontiguousarray(fx["forcing"][:, :, idx]))
    assert np.array_equal(out, out2)
    o = oracle_lib.Oracle(fx["opt"], fx["lib"], cells)
    o.init_state(fx["tair0"][idx])
    _, ref, _ = o.step(fx["month"], fx["doy"], np.ascontiguousarray(fx["forcing"][:, :, idx]))
    check_outputs(out, ref, names, what="ensemble vs oracle", mode="cells")
    m.close(); m2.close(); o.close()


@pytest.mark.parametrize("nbands,dt,cold", [(1, 24, 0.0), (5, 24, 12.0), (2, 3, 15.0)])
def test_cuda_matches_oracle_on_seeded_grid(VicRes, nbands, dt, cold):
    C, nrec = 600, (400 if dt == 24 else 480)
    lib, cells = synth_soa.make_cells(C, nbands=nbands, seed=11 + nbands, lat_range=(20.0, 65.0), overstory_frac=0.5)
    f, month, doy = synth_soa.make_forcing(cells, nrec, dt_hours=dt, seed=5, cold=cold)
    opt = abi.default_options(3, nbands, dt, out=abi.OUT_NAMES)
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(f[0, 0])
    rc, ref, _ = o.step(month, doy, f)
    assert rc == 0
    m = VicRes(opt, lib, cells)
    m.init_state(f[0, 0])
    out, status = m.step(month, doy, f)
    assert not status.any()
    check_outputs(out, ref, abi.OUT_NAMES, what="seeded grid nb=%d dt=%d" % (nbands, dt), mode="cells")
    assert abs(int(m.fallback_counts().sum()) - int(o.fallback_counts().sum())) <= 2
    assert ref[abi.OUT["SWE"]].max() > (10.0 if cold else -1)
    m.close(); o.close()


def test_closure_and_sampled_parity_at_scale(VicRes):
    """50k cells x 1 year: water-balance closure no worse than the reference's bar (1.35e-13 mm on
    the bundled cell scales with storage; we hold 5e-12 mm), all outputs finite, and a random sample
    of cells reproduces the oracle."""
    C, nrec = 50000, 365
    lib, cells = synth_soa.make_cells(C, nbands=1, seed=20261019)
    f, month, doy = synth_soa.make_forcing(cells, nrec, seed=7)
    names = abi.default_out_vars(3) + ["WATER_ERROR"]
    opt = abi.default_options(3, 1, 24, out=names)
    m = VicRes(opt, lib, cells)
    m.init_state(f[0, 0])
    out, status = m.step(month, doy, f)
    assert not status.any() and np.isfinite(out).all()
    assert np.abs(out[names.index("WATER_ERROR")]).max() < 5e-12
    # mass conservation over the year, cell by cell: sum(P - E - R - B) == change in storage
    i = {n: names.index(n) for n in names}
    stor = out[i["SOIL_LIQ0"]] + out[i["SOIL_LIQ1"]] + out[i["SOIL_LIQ2"]] + out[i["SWE"]] + out[i["SNOW_CANOPY"]] + out[i["WDEW"]]
    flux = (out[i["PREC"]] - out[i["EVAP"]] - out[i["RUNOFF"]] - out[i["BASEFLOW"]])[1:].sum(0)
    assert np.abs(flux - (stor[-1] - stor[0])).max() < 1e-8
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(C, 40, replace=False))
    sub, _ = _tile(cells, 1)
    sub = {k: v for k, v in sub.items()}
    # build the sub-grid of picked cells
    sel = {}
    for k in abi.CELL_SCALARS:
        sel[k] = cells[k][pick]
    for k in abi.CELL_LAYER + abi.CELL_BAND:
        sel[k] = cells[k][:, pick]
    tp = cells["tile_ptr"]
    tiles = np.concatenate([np.arange(tp[c], tp[c + 1]) for c in pick])
    sel["tile_ptr"] = np.concatenate([[0], np.cumsum((tp[1:] - tp[:-1])[pick])]).astype(np.int64)
    for k in ["tile_class", "tile_Cv", "tile_root", "tile_LAI", "tile_albedo", "tile_vegcover"]:
        sel[k] = cells[k][tiles]
    o = oracle_lib.Oracle(opt, lib, sel)
    fs = np.ascontiguousarray(f[:, :, pick])
    o.init_state(fs[0, 0])
    _, ref, _ = o.step(month, doy, fs)
    check_outputs(out[:, :, pick], ref, names, what="sampled cells at 50k", mode="cells")
    m.close(); o.close()


def _oracle_threads(opt, lib, cells, month, doy, f, nthreads=16):
    """the oracle over all cells, on threads (cells are independent; the C call releases the GIL)"""
    import threading
    C = f.shape[2]
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(f[0, 0])
    ref = np.zeros((opt.n_out, f.shape[1], C))
    edges = np.linspace(0, C, min(nthreads, C) + 1).astype(int)
    th = [threading.Thread(target=o.step_range, args=(month, doy, f, ref, int(a), int(b))) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    for t in th:
        t.start()
    for t in th:
        t.join()
    o.close()
    return ref


@pytest.mark.parametrize("config", ["cfg2", "cfg3"])
def test_parity_at_baseline_shapes(VicRes, config):
    """BASELINE.json configurations at their full record counts: 2 000 cells of configuration 2's distribution x 3 653
    daily records (10 years), and 500 cells of configuration 3's (5 snow bands, 3-hourly, cold, 30 % overstory) x 2 920
    records (1 year), every cell and record against the oracle."""
    if config == "cfg2":
        C, nrec, nb, dt, cold, kw = 2000, 3653, 1, 24, 0.0, {}
    else:
        C, nrec, nb, dt, cold, kw = 500, 2920, 5, 3, 5.0, {"lat_range": (35.0, 65.0), "overstory_frac": 0.3}
    lib, cells = synth_soa.make_cells(C, nlayer=3, nbands=nb, seed=20261019, **kw)
    f, month, doy = synth_soa.make_forcing(cells, nrec, dt_hours=dt, seed=7, cold=cold)
    names = abi.default_out_vars(3) + ["WATER_ERROR"]
    opt = abi.default_options(3, nb, dt, out=names)
    m = VicRes(opt, lib, cells)
    m.init_state(f[0, 0])
    out, status = m.step(month, doy, f)
    assert not status.any()
    ref = _oracle_threads(opt, lib, cells, month, doy, f)
    from common import MIN_FRAC_LONG
    check_outputs(out, ref, names, what="%s shape CUDA vs oracle (%d cells x %d records)" % (config, C, nrec), mode="cells",
                  min_frac=MIN_FRAC_LONG if config == "cfg2" else 0.995, max_bad_cells=0.06 if config == "cfg2" else 0.09)
    assert np.abs(out[names.index("WATER_ERROR")]).max() < 5e-11
    m.close()


def test_errors_are_reported_not_fatal(VicRes):
    fx = load_fixture("daily_warm")
    m = VicRes(fx["opt"], fx["lib"], fx["cells"])
    from vicres_b200.model import VicResError
    with pytest.raises(VicResError) as e:
        m.step(fx["month"][:5], fx["doy"][:5], np.ascontiguousarray(fx["forcing"][:, :5]))   # before init_state
    assert e.value.code == abi.VR_ERR_STATE
    m.close()


def test_failing_cell_is_reported_and_isolated(VicRes):
    """A cell whose top layer starts below zero moisture makes arno_evap return ERROR (arno_evap.c:112-151: ratio > 1).
    The library keeps going for the other cells, reports VR_ERR_CELL through the status array -- cell_status = first
    failing record + 1 -- and the neighbours' results are bit-identical to a run without the broken cell."""
    lib, cells = synth_soa.make_cells(96, nbands=1, seed=21)
    f, month, doy = synth_soa.make_forcing(cells, 12, seed=4)
    opt = abi.default_options(3, 1, 24)
    m = VicRes(opt, lib, cells)
    m.init_state(f[0, 0])
    good, st0 = m.step(month, doy, f)
    assert not st0.any()
    m.close()
    bad = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in cells.items()}
    bad["init_moist"][0, 17] = -5.0
    m = VicRes(opt, lib, bad)
    m.init_state(f[0, 0])
    out, st = m.step(month, doy, f)
    assert st[17] == 1, "first failing record + 1, got %d" % st[17]
    assert not np.delete(st, 17).any()
    keep = np.delete(np.arange(96), 17)
    assert np.array_equal(out[:, :, keep], good[:, :, keep])
    # the count continues over blocks: a second block of records reports record 12 + 1 .. for a cell failing later
    m.close()


def test_device_pow_error_bound():
    """vic_physics.cuh m_pow: relative error below 1e-13 over the model's argument ranges (Brooks-Corey
    drainage (0..1]^[3,40], ARNO curves, albedo decay 0.92^(t^0.58)), exact special cases"""
    from vicres_b200 import model
    rng = np.random.default_rng(3)
    n = 400000
    x = np.concatenate([rng.uniform(1e-6, 1.0, n), rng.uniform(0.0, 3.0, n), np.exp(rng.uniform(-30, 30, n))])
    y = np.concatenate([rng.uniform(3.0, 40.0, n), rng.uniform(0.0, 3.0, n), rng.uniform(-8.0, 8.0, n)])
    z = model.diag_pow(x, y)
    ref = np.power(x, y)
    ok = np.isfinite(ref) & (ref > 1e-300)
    rel = np.abs(z[ok] - ref[ok]) / ref[ok]
    assert rel.max() < 1e-13, "max relative error %.3e" % rel.max()
    sx = np.array([1.0, 0.0, 0.0, 2.0])
    sy = np.array([7.3, 2.0, 0.0, 0.0])
    assert np.array_equal(model.diag_pow(sx, sy), np.array([1.0, 0.0, 1.0, 1.0]))


def test_edge_shapes(VicRes):
    """one cell x one record; a grid whose last block is nearly empty; a cell without any tile (outputs 0,
    neighbours unaffected); a single output variable"""
    lib, cells = synth_soa.make_cells(1, nbands=1, seed=2)
    f, month, doy = synth_soa.make_forcing(cells, 1, seed=3)
    opt = abi.default_options(3, 1, 24)
    m = VicRes(opt, lib, cells)
    m.init_state(f[0, 0])
    out, status = m.step(month, doy, f)
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(f[0, 0])
    _, ref, _ = o.step(month, doy, f)
    assert out.shape == (opt.n_out, 1, 1) and relerr(out, ref).max() < RTOL
    m.close(); o.close()
    # 33 cells, cell 7 stripped of its tiles
    lib, cells = synth_soa.make_cells(33, nbands=2, seed=4)
    tp = cells["tile_ptr"].copy()
    a, b = int(tp[7]), int(tp[8])
    keep = np.r_[0:a, b:int(tp[-1])]
    for k in ["tile_class", "tile_Cv", "tile_root", "tile_LAI", "tile_albedo", "tile_vegcover"]:
        cells[k] = cells[k][keep]
    tp[8:] -= (b - a)
    cells["tile_ptr"] = tp
    f, month, doy = synth_soa.make_forcing(cells, 37, seed=6)
    opt1 = abi.default_options(3, 2, 24, out=["RUNOFF"])
    m = VicRes(opt1, lib, cells)
    m.init_state(f[0, 0])
    out, status = m.step(month, doy, f)
    assert out.shape == (1, 37, 33) and np.isfinite(out).all() and not status.any()
    assert (out[0, :, 7] == 0).all() and out[0].sum() > 0
    m.close()
