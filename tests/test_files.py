"""Parameter-file readers and result-file writer (include/vicres_files.h) against the reference.

CPU tier: (1) the committed text case tests/golden/files_case (tools/make_files_golden.py): the
readers must give exactly (bit for bit) the numbers the reference held after read_soilparam /
read_veglib / read_vegparam / read_snowband, and files -> readers -> oracle -> writer must reproduce
the reference's own fluxes_* files byte for byte; (2) when the compiled reference is here, fresh
seeded cases and the reference's bundled basin files.  GPU tier: the same chain with the CUDA path."""
import filecmp
import os
import subprocess
import numpy as np
import pytest
import oracle_lib
import refdump
from vicres_b200 import abi, build, files, synth
from common import GOLDEN, ROOT, load_fixture

CASE = os.path.join(GOLDEN, "files_case")
DRIVER = os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver")
REF = "/root/reference/runoff"


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build.build()


def _same(got, want, what):
    for k, v in want.items():
        a, v = np.asarray(got[k]), np.asarray(v)
        assert a.shape == v.shape or a.size == v.size, "%s %s: shape %s vs %s" % (what, k, a.shape, v.shape)
        assert np.array_equal(a.reshape(v.shape), v), "%s %s differs from the reference's value" % (what, k)


def _read_case(path, nlayer, nbands):
    return files.read_params(os.path.join(path, "soil.txt"), os.path.join(path, "veglib.txt"),
                             os.path.join(path, "veg.txt"),
                             os.path.join(path, "snowband.txt") if nbands > 1 else None,
                             nlayer=nlayer, nbands=nbands, root_zones=3)


def _expected():
    z = np.load(os.path.join(CASE, "expected.npz"))
    lib = {k[4:]: z[k] for k in z.files if k.startswith("lib_")}
    cells = {k[6:]: z[k] for k in z.files if k.startswith("cells_")}
    return z, lib, cells


def test_readers_match_reference_structs_exactly():
    z, lib_ref, cells_ref = _expected()
    L, NB, _ = [int(x) for x in z["meta"]]
    lib, cells, meta = _read_case(CASE, L, NB)
    _same(lib, lib_ref, "veglib")
    _same(cells, cells_ref, "cells")
    assert meta["gridcel"].tolist() == [1, 2, 3, 4, 5]
    assert np.array_equal(meta["lat"].astype(np.float64), cells["lat"])


def _fluxes_match(out_dir, ref_dir):
    names = sorted(os.listdir(ref_dir))
    assert names and sorted(n for n in os.listdir(out_dir) if n.startswith("fluxes_")) == names
    for n in names:
        assert filecmp.cmp(os.path.join(out_dir, n), os.path.join(ref_dir, n), shallow=False), n


def test_files_to_oracle_to_writer_reproduces_reference_flux_files(tmp_path):
    z, _, _ = _expected()
    L, NB, dt = [int(x) for x in z["meta"]]
    lib, cells, meta = _read_case(CASE, L, NB)
    mst, mrt, wind_h = [float(x) for x in z["meta_f"]]
    opt = abi.default_options(L, NB, dt, max_snow_temp=mst, min_rain_temp=mrt, wind_h=wind_h)
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(z["tair0"])
    rc, out, _ = o.step(z["month"], z["doy"], z["forcing"])
    assert rc == 0
    o.close()
    dmy = z["dmy"]
    nflux = 20 + L - 1                                       # set_output_defaults.c:78-84
    files.write_results(str(tmp_path), "fluxes", meta["lat"], meta["lng"], dmy[:, 0], dmy[:, 1], dmy[:, 2], out,
                        cols=np.arange(nflux))
    _fluxes_match(str(tmp_path), os.path.join(CASE, "fluxes"))


def test_reader_errors(tmp_path):
    with pytest.raises(Exception, match="cannot open"):
        files.read_params(str(tmp_path / "nope"), os.path.join(CASE, "veglib.txt"), os.path.join(CASE, "veg.txt"))
    with pytest.raises(Exception, match="too few columns"):
        # a 3-layer soil file read as NLAYER... the case file has 53 columns; ask for more layers than MAX
        short = tmp_path / "soil_short.txt"
        short.write_text("1 1 10.0 20.0 0.1 0.1\n")
        files.read_params(str(short), os.path.join(CASE, "veglib.txt"), os.path.join(CASE, "veg.txt"))
    with pytest.raises(Exception, match="snow band"):
        files.read_params(os.path.join(CASE, "soil.txt"), os.path.join(CASE, "veglib.txt"),
                          os.path.join(CASE, "veg.txt"), None, nbands=3)
    skip = tmp_path / "soil_skip.txt"
    rows = open(os.path.join(CASE, "soil.txt")).read().splitlines()
    rows[1] = "0" + rows[1][1:]                                # RUN flag off: the reference skips the row
    skip.write_text("\n".join(rows) + "\n")
    _, cells, meta = files.read_params(str(skip), os.path.join(CASE, "veglib.txt"), os.path.join(CASE, "veg.txt"),
                                       os.path.join(CASE, "snowband.txt"), nbands=3)
    assert meta["gridcel"].tolist() == [1, 3, 4, 5] and cells["lat"].shape == (4,)


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="compiled reference (oracle/_ref) not built here")
@pytest.mark.parametrize("kw", [dict(ncells=9, nlayer=3, nbands=1, seed=5), dict(ncells=7, nlayer=2, nbands=4, seed=6),
                                dict(ncells=6, nlayer=3, nbands=2, seed=8, dt=3)])
def test_readers_and_writer_against_live_reference(tmp_path, kw):
    work = str(tmp_path / "case")
    g = synth.make_case(work, ndays=15, cold=8.0, **kw)
    dump = os.path.join(work, "dump.bin")
    subprocess.run([DRIVER, "-g", g, "-d", dump], check=True, capture_output=True)
    d = refdump.read_dump(dump)
    cs = refdump.case_from_dump(d)
    lib, cells, meta = _read_case(work, kw["nlayer"], kw["nbands"])
    _same(lib, cs["lib"], "veglib")
    _same(cells, cs["cells"], "cells")
    opt = abi.default_options(cs["nlayer"], cs["nbands"], cs["dt"], max_snow_temp=cs["max_snow_temp"],
                              min_rain_temp=cs["min_rain_temp"], wind_h=cs["wind_h"])
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(cs["tair0"])
    rc, out, _ = o.step(cs["month"], cs["doy"], cs["forcing"])
    assert rc == 0
    o.close()
    dmy = np.asarray(d["dmy"])
    outdir = str(tmp_path / "mine")
    os.makedirs(outdir)
    nflux = 20 + cs["nlayer"] - 1
    hour = dmy[:, 3] if cs["dt"] < 24 else None
    files.write_results(outdir, "fluxes", meta["lat"], meta["lng"], dmy[:, 0], dmy[:, 1], dmy[:, 2], out,
                        cols=np.arange(nflux), hour=hour)
    files.write_results(outdir, "snow", meta["lat"], meta["lng"], dmy[:, 0], dmy[:, 1], dmy[:, 2], out,
                        cols=np.arange(nflux, nflux + 4), hour=hour)
    ref_dir = os.path.join(work, "results")
    for n in sorted(os.listdir(ref_dir)):
        assert filecmp.cmp(os.path.join(outdir, n), os.path.join(ref_dir, n), shallow=False), n
    assert len(os.listdir(ref_dir)) == 2 * cs["ncell"]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "parameters", "soil.txt")), reason="reference tree not present")
def test_bundled_basin_files_parse_like_the_reference():
    """The reference's own Mekong parameter files (NLAYER 2): every flagged row parses, and the one
    cell whose forcing ships (soil.txt row 14208, tools/make_golden.py) equals the golden fixture."""
    lib, cells, meta = files.read_params(os.path.join(REF, "parameters", "soil.txt"),
                                         os.path.join(REF, "parameters", "vegetlibrary.txt"),
                                         os.path.join(REF, "parameters", "land.txt"), nlayer=2, nbands=1, root_zones=3,
                                         vegparam_lai=True, lai_from_vegparam=True)   # globalparam.txt:45-56
    fx = load_fixture("bundled_basin_cell")
    with open(os.path.join(REF, "parameters", "soil.txt")) as f:
        row = f.readlines()[14207].split()
    i = int(np.nonzero(meta["gridcel"] == int(row[1]))[0][0])
    _same(lib, fx["lib"], "veglib")
    for k in abi.CELL_SCALARS:
        assert cells[k][i] == fx["cells"][k][0], k
    for k in abi.CELL_LAYER + abi.CELL_BAND:
        assert np.array_equal(cells[k][:, i], np.asarray(fx["cells"][k]).reshape(-1)), k
    a, b = cells["tile_ptr"][i], cells["tile_ptr"][i + 1]
    assert b - a == fx["cells"]["tile_ptr"][1]
    for k in ["tile_class", "tile_Cv", "tile_root", "tile_LAI", "tile_albedo", "tile_vegcover"]:
        assert np.array_equal(np.asarray(cells[k][a:b]).reshape(-1), np.asarray(fx["cells"][k]).reshape(-1)), k


@pytest.mark.gpu
def test_files_to_cuda_to_writer(tmp_path):
    """GPU tier: files -> readers -> CUDA path -> writer; the printed %.4f columns must equal the
    reference's flux files except where a value sits within the parity envelope of a rounding edge."""
    from vicres_b200 import model
    z, _, _ = _expected()
    L, NB, dt = [int(x) for x in z["meta"]]
    lib, cells, meta = _read_case(CASE, L, NB)
    mst, mrt, wind_h = [float(x) for x in z["meta_f"]]
    opt = abi.default_options(L, NB, dt, max_snow_temp=mst, min_rain_temp=mrt, wind_h=wind_h)
    m = model.VicRes(opt, lib, cells)
    m.init_state(z["tair0"])
    out, status = m.step(z["month"], z["doy"], z["forcing"])
    assert not status.any()
    m.close()
    dmy = z["dmy"]
    nflux = 20 + L - 1
    files.write_results(str(tmp_path), "fluxes", meta["lat"], meta["lng"], dmy[:, 0], dmy[:, 1], dmy[:, 2], out,
                        cols=np.arange(nflux))
    ref_dir = os.path.join(CASE, "fluxes")
    nbad = ntot = 0
    for n in sorted(os.listdir(ref_dir)):
        a = np.loadtxt(os.path.join(str(tmp_path), n))
        b = np.loadtxt(os.path.join(ref_dir, n))
        assert a.shape == b.shape
        assert np.array_equal(a[:, :3], b[:, :3])
        assert np.all(np.abs(a - b) <= 1.0001e-4 + 1e-5 * np.abs(b)), n     # one unit in the 4th decimal
        nbad += int((a != b).sum())
        ntot += a.size
    assert nbad <= 0.01 * ntot


def test_round_decimals_equals_text_round_trip():
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.normal(0, 50, 200000), rng.uniform(-1, 1, 100000) * 1e-3,
                        np.round(rng.normal(0, 5, 100000), 4) + 0.00005,          # decimal ties and near-ties
                        np.array([0.0, -0.0, 0.00005, 0.00015, 0.00025, 1e-9, 123456.78905, -2.5e-5])])
    want = np.array([float("%.4f" % v) for v in x])
    got = files.round_decimals(x, 4)
    assert np.array_equal(got, want)


def test_global_file_calendar_and_forcing_reader(tmp_path):
    """get_global_param / make_dmy / read_atmos_data on a written case: END* date -> NRECS across a leap year,
    option flags, ASCII and 2-byte binary forcing columns"""
    g = tmp_path / "global.txt"
    forc = tmp_path / "forc_"
    g.write_text("\n".join([
        "# comment", "NLAYER 2", "TIME_STEP 24", "SNOW_STEP 24", "STARTYEAR 1999", "STARTMONTH 12", "STARTDAY 30", "STARTHOUR 0",
        "ENDYEAR 2000", "ENDMONTH 3", "ENDDAY 1", "FULL_ENERGY FALSE", "FROZEN_SOIL FALSE", "MIN_WIND_SPEED 0.1",
        "MAX_SNOW_TEMP 0.7", "MIN_RAIN_TEMP -0.3", "FORCING1 %s" % forc, "FORCE_FORMAT ASCII", "N_TYPES 4", "FORCE_TYPE PREC",
        "FORCE_TYPE TMAX", "FORCE_TYPE TMIN", "FORCE_TYPE WIND", "FORCE_DT 24", "FORCEYEAR 1999", "FORCEMONTH 12",
        "FORCEDAY 28", "FORCEHOUR 0", "GRID_DECIMAL 4", "WIND_H 10.0", "SOIL a/soil.txt", "VEGLIB a/veglib.txt",
        "VEGPARAM a/veg.txt", "ROOT_ZONES 3", "VEGPARAM_LAI TRUE", "LAI_SRC FROM_VEGPARAM", "SNOW_BAND 5 a/bands.txt",
        "RESULT_DIR out/", "SKIPYEAR 0"]) + "\n")
    G = files.Global(str(g))
    assert (G.nlayer, G.time_step, G.snow_step, G.snow_band, G.root_zones) == (2, 24, 24, 5, 3)
    assert G.nrecs == 2 + 31 + 29 + 1 and G.unsupported is None            # 30 Dec 1999 .. 1 Mar 2000, leap February
    assert (G.max_snow_temp, G.min_rain_temp) == (0.7, -0.3) and G.snowband == "a/bands.txt"
    assert G.param_files()["lai_from_vegparam"] and G.force_types == ["PREC", "TMAX", "TMIN", "WIND"]
    dmy = G.dmy()
    assert dmy[0].tolist() == [1999, 12, 30, 0, 364] and dmy[2].tolist() == [2000, 1, 1, 0, 1]
    assert dmy[-1].tolist() == [2000, 3, 1, 0, 61]
    rng = np.random.default_rng(0)
    n = G.nrecs + 2
    vals = np.round(np.stack([rng.gamma(1, 5, n), rng.normal(20, 5, n), rng.normal(8, 5, n), rng.uniform(0, 9, n)], 1), 2)
    (tmp_path / "forc_12.3400_-56.7800").write_text("\n".join(" ".join("%.2f" % x for x in r) for r in vals) + "\n")
    cols, got = G.read_forcing(12.34, -56.78)
    assert got == G.nrecs
    for i, k in enumerate(G.force_types):                                  # two records skipped (FORCEDAY 28 -> STARTDAY 30)
        assert np.array_equal(cols[k], vals[2:, i])
    # the same data as 2-byte scaled binary with the VIC header
    g2 = tmp_path / "global_bin.txt"
    g2.write_text(g.read_text().replace("FORCE_FORMAT ASCII", "FORCE_FORMAT BINARY\nFORCE_ENDIAN LITTLE")
                  .replace("FORCE_TYPE PREC", "FORCE_TYPE PREC UNSIGNED 40").replace("FORCE_TYPE TMAX", "FORCE_TYPE TMAX SIGNED 100")
                  .replace("FORCE_TYPE TMIN", "FORCE_TYPE TMIN SIGNED 100").replace("FORCE_TYPE WIND", "FORCE_TYPE WIND SIGNED 100")
                  .replace(str(forc), str(tmp_path / "bin_")))
    mult = np.array([40, 100, 100, 100])
    raw = np.round(vals * mult).astype(np.int64)
    hdr = np.array([0xFFFF] * 4 + [10], dtype="<u2").tobytes()
    body = b"".join(np.array([r[0]], dtype="<u2").tobytes() + np.array(r[1:], dtype="<i2").tobytes() for r in raw)
    (tmp_path / "bin_12.3400_-56.7800").write_bytes(hdr + body)
    G2 = files.Global(str(g2))
    cols2, got2 = G2.read_forcing(12.34, -56.78)
    assert got2 == G.nrecs
    for i, k in enumerate(G2.force_types):
        assert np.array_equal(cols2[k], raw[2:, i] / mult[i])
    bad = tmp_path / "fe.txt"
    bad.write_text(g.read_text().replace("FULL_ENERGY FALSE", "FULL_ENERGY TRUE"))
    assert files.Global(str(bad)).unsupported == "FULL_ENERGY TRUE"
    # options that change the results and have a value outside what the library implements are named, never ignored
    for line, why in [("RC_MODE RC_PHOTO", "RC_MODE RC_PHOTO"), ("ARC_SOIL TRUE", "ARC_SOIL TRUE"),
                      ("SNOW_DENSITY DENS_SNTHRM", "SNOW_DENSITY DENS_SNTHRM"), ("PRT_SNOW_BAND TRUE", "PRT_SNOW_BAND TRUE"),
                      ("AERO_RESIST_CANSNOW AR_410", "AERO_RESIST_CANSNOW AR_410"), ("SPATIAL_SNOW TRUE", "SPATIAL_SNOW TRUE")]:
        bad.write_text(g.read_text() + line + "\n")
        assert files.Global(str(bad)).unsupported == why
    bad.write_text(g.read_text() + "SNOW_ALBEDO USACE\nGRND_FLUX_TYPE GF_410\nQUICK_FLUX TRUE\nEQUAL_AREA FALSE\nRESOLUTION 0.125\n"
                   "PRT_HEADER TRUE\nALMA_OUTPUT TRUE\nMOISTFRACT TRUE\n")
    G3 = files.Global(str(bad))
    assert G3.unsupported is None and G3.prt_header == 1 and G3.alma_output == 1 and G3.moistfract == 1
    sub = tmp_path / "sub.txt"
    sub.write_text(g.read_text().replace("SNOW_STEP 24", "SNOW_STEP 3"))
    assert files.Global(str(sub)).unsupported is None                       # the snow model sub-stepped: NF = 8
    sub.write_text(g.read_text().replace("SNOW_STEP 24", "SNOW_STEP 5"))
    assert "divide TIME_STEP evenly" in files.Global(str(sub)).unsupported      # get_global_param.c:764-765


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="compiled reference (oracle/_ref) not built here")
def test_global_file_against_live_reference(tmp_path):
    """the calendar and record count the reference derives from the same global file (dump of vic_ref_driver),
    and its raw forcing columns"""
    work = str(tmp_path / "case")
    g = synth.make_case(work, ncells=2, nlayer=3, nbands=2, ndays=430, startyear=1983, seed=12)
    dump = os.path.join(work, "dump.bin")
    subprocess.run([DRIVER, "-g", g, "-d", dump], check=True, capture_output=True)
    d = refdump.read_dump(dump)
    G = files.Global(g)
    assert G.unsupported is None and G.nrecs == int(d["nrecs"][0])
    assert np.array_equal(G.dmy(), np.asarray(d["dmy"]))
    lib, cells, meta = files.read_params(**G.param_files())
    cs = refdump.case_from_dump(d)
    _same(cells, cs["cells"], "cells via the global file")
    cols, got = G.read_forcing(float(meta["lat"][0]), float(meta["lng"][0]))
    assert got == G.nrecs and set(cols) == {"PREC", "TMAX", "TMIN", "WIND"}
    # the reference's step sees prec unchanged in daily mode (initialize_atmos only disaggregates sub-daily)
    assert np.allclose(cols["PREC"], cs["forcing"][1, :, 0], rtol=0, atol=1e-12)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "parameters", "globalparam.txt")), reason="reference tree not present")
def test_bundled_global_file():
    G = files.Global(os.path.join(REF, "parameters", "globalparam.txt"))
    assert G.unsupported is None and (G.nlayer, G.time_step, G.root_zones) == (2, 24, 3)
    assert G.nrecs == 15340                                  # 1981-01-01 .. 2022-12-31
    assert G.force_types == ["PREC", "TMAX", "TMIN", "WIND"] and G.vegparam_lai == 1 and G.lai_from_vegparam == 1


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="compiled reference (oracle/_ref) not built here")
def test_baseflow_nijssen2001_parameters_against_live_reference(tmp_path):
    """BASEFLOW NIJSSEN2001: soil columns 5-8 converted to the ARNO Ds, Dsmax, Ws (read_soilparam.c:719-726), bit for bit
    what the compiled reference holds in soil_con after reading the same file"""
    work = str(tmp_path / "case")
    g = synth.make_case(work, ncells=4, nlayer=3, nbands=1, ndays=30, startyear=1990, seed=91, extra_global="BASEFLOW NIJSSEN2001")
    dump = str(tmp_path / "dump.bin")
    subprocess.run([DRIVER, "-g", g, "-d", dump], check=True, capture_output=True)
    ref = refdump.case_from_dump(refdump.read_dump(dump))["cells"]
    G = files.Global(g)
    assert G.unsupported is None and G.baseflow_nijssen == 1
    _, cells, _ = files.read_params(**G.param_files())
    _, plain, _ = files.read_params(**dict(G.param_files(), baseflow_nijssen=False))
    for k in ("Ds", "Dsmax", "Ws", "c_expt", "b_infilt"):
        assert np.array_equal(cells[k], ref[k]), k
    assert not np.array_equal(cells["Dsmax"], plain["Dsmax"]) and np.array_equal(cells["c_expt"], plain["c_expt"])
