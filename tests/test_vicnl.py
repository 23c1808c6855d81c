"""The whole `vicNl -g <global file>` command line on the GPU (vicres_b200/vicnl.py: global file -> parameter files
-> forcing files -> forcing preparation -> cell step -> result files) against the files the UNMODIFIED reference
wrote for the same case (tests/golden/vicnl_case, tools/make_vicnl_golden.py)."""
import os
import shutil
import numpy as np
import pytest
from common import GOLDEN
from vicres_b200 import files

CASES = os.path.join(GOLDEN, "vicnl_case")


def _stage(name, tmp_path, variant=None):
    """copy the inputs, point the global file at the copy, drop the reference's results"""
    dst = tmp_path / name
    shutil.copytree(os.path.join(CASES, name), dst)
    rdir = "results" if variant is None else "results_" + variant
    ref = tmp_path / (name + "_ref")
    shutil.move(str(dst / rdir), ref)
    os.makedirs(dst / rdir)
    g = dst / ("global.txt" if variant is None else "global_%s.txt" % variant)
    g.write_text(g.read_text().replace("@DIR@", str(dst)))
    return str(g), str(dst / rdir), str(ref)


# variants of the two cases: output interval, output selection, binary files (tools/make_vicnl_golden.py: VARIANTS)
VARIANTS = [("sub3h", "out24"), ("sub3h", "out6sel"), ("daily", "selday"), ("sub3h", "bin"), ("sub3h", "bin24sel"),
            ("daily", "snow3"), ("skip", "snow1"), ("daily", "hdr"), ("sub3h", "hdrbin24"), ("sub3h", "alma6"), ("daily", "almadef"),
            ("sub3h", "met6"), ("daily", "metalma"), ("daily", "force"), ("sub3h", "forcebin"), ("daily", "force3"),
            ("daily", "mfrac"), ("sub3h", "mfrac6")]


def _quantum(tok):
    """one unit of the last printed digit of a number as the reference printed it"""
    t = tok.lower()
    if "e" in t:
        m, e = t.split("e")
        return 10.0 ** (int(e) - (len(m.split(".")[1]) if "." in m else 0))
    return 10.0 ** -(len(t.split(".")[1]) if "." in t else 0)


def _compare_ascii(res, ref, what):
    """every token of every result file: identical text, or one unit of the last printed digit apart (cancellation
    residues like OUT_WATER_ERROR ~ 1e-14 mm: both sides below 1e-9)"""
    assert sorted(os.listdir(res)) == sorted(os.listdir(ref))
    nbad = ntot = 0
    for n in sorted(os.listdir(ref)):
        la, lb = open(os.path.join(res, n)).read().split("\n"), open(os.path.join(ref, n)).read().split("\n")
        assert len(la) == len(lb), n
        for ra, rb in zip(la, lb):
            ta, tb = ra.split("\t"), rb.split("\t")
            assert len(ta) == len(tb), (n, ra, rb)
            for xa, xb in zip(ta, tb):
                ntot += 1
                if xa == xb:
                    continue
                a, b = float(xa), float(xb)
                nbad += int(max(abs(a), abs(b)) >= 1e-9)
                assert abs(a - b) <= max(1.0001 * _quantum(xb.strip()) + 1e-5 * abs(b), 1e-9), (n, rb, xa, xb)
    print("VICNL %s: %d of %d printed values differ from the reference's files (by one unit of the last digit)" % (what, nbad, ntot))
    assert nbad <= 0.01 * ntot


def _compare_binary(res, ref, gpath, nlayer, hourly, what, header=False, no_date=False):
    """BINARY_OUTPUT files: same size, same dates; integer columns within one count (the cast truncates), floats within
    fp32 rounding of the parity tolerance"""
    sel = files.OutSel(gpath, nlayer)
    assert sorted(os.listdir(res)) == sorted(os.listdir(ref))
    nbad = ntot = 0
    for prefix, cols in sel.files:
        for n in sorted(x for x in os.listdir(ref) if x.startswith(prefix + "_")):
            assert os.path.getsize(os.path.join(res, n)) == os.path.getsize(os.path.join(ref, n)), n
            da, a = files.read_binary_results(os.path.join(res, n), cols, hourly, header, no_date)
            db, b = files.read_binary_results(os.path.join(ref, n), cols, hourly, header, no_date)
            if header:       # PRT_HEADER: the block before the records, byte for byte
                nb = int(np.fromfile(os.path.join(ref, n), dtype="<u2", count=5)[4])
                assert open(os.path.join(res, n), "rb").read(nb) == open(os.path.join(ref, n), "rb").read(nb), n
            assert np.array_equal(da, db), n
            for k, c in enumerate(cols):
                tol = 1.0001 / c[3] if c[2] in ("SINT", "USINT", "INT", "CHAR") else 1e-4 + 1e-5 * np.abs(b[k])
                assert np.all(np.abs(a[k] - b[k]) <= tol), (n, c)
                # doubles are written in full: "differs" = beyond the parity tolerance, for the other types any difference
                nbad += int((np.abs(a[k] - b[k]) > 1e-9 * np.maximum(np.abs(b[k]), 1e-3)).sum() if c[2] == "DOUBLE" else (a[k] != b[k]).sum())
            ntot += a.size
    print("VICNL %s: %d of %d binary values differ from the reference's files" % (what, nbad, ntot))
    assert nbad <= 0.01 * ntot


def _compare_variant(case, variant, gpath, res, ref, what):
    if "bin" in variant:
        _compare_binary(res, ref, gpath, 3, variant == "bin", what, header=variant.startswith("hdr"), no_date=variant.startswith("force"))
    else:
        _compare_ascii(res, ref, what)


@pytest.mark.parametrize("case,variant", VARIANTS)
def test_output_selection_of_the_variants_parses(case, variant, tmp_path):
    """CPU tier: N_OUTFILES / OUTFILE / OUTVAR, OUT_STEP and BINARY_OUTPUT are read as parse_output_info.c reads them"""
    gpath, _, ref = _stage(case, tmp_path, variant)
    g = files.Global(gpath)
    assert g.unsupported is None
    sel = files.OutSel(gpath, g.nlayer)
    fl = dict(sel.files)
    assert sorted(set(n.rsplit("_", 2)[0] for n in os.listdir(ref))) == sorted(fl)       # <prefix>_<lat>_<lng>
    assert bool(g.binary_output) == ("bin" in variant)
    assert bool(g.prt_header) == (variant in ("hdr", "hdrbin24", "alma6", "force3"))
    assert bool(g.alma_output) == ("alma" in variant or variant == "force3")
    if variant == "alma6":
        assert "SUB_CANOP" not in sel.step_vars() and sel.step_vars(alma=True)[-1] == "SUB_CANOP"
        assert [c[0] for c in fl["al"]][4] == "SUB_SNOW"
    if variant.startswith("force"):
        assert g.output_force == 1 and sel.c.no_date == 1
        if variant != "force3":
            assert list(fl) == ["full_data"] and [c[0] for c in fl["full_data"]] == ["PREC", "AIR_TEMP", "SHORTWAVE", "LONGWAVE", "DENSITY",
                                                                                  "PRESSURE", "VP", "WIND"]
            assert [(c[2], c[3]) for c in fl["full_data"]][:3] == [("USINT", 40.0), ("SINT", 100.0), ("USINT", 50.0)]
    if variant.startswith("snow"):
        assert g.snow_step == int(variant[4:]) and g.time_step == 24
    if variant in ("out24", "bin"):
        assert [len(c) for c in fl.values()] == [22, 4] and g.out_step == (24 if variant == "out24" else 0)
        assert sel.step_vars(aggregate=True)[15] == "AERO_COND" and sel.step_vars()[15] == "AERO_RESIST"
    if variant == "out6sel":
        assert g.out_step == 6
        assert [c[0] for c in fl["wb"]] == ["PREC", "EVAP", "RUNOFF", "BASEFLOW", "SOIL_LIQ0", "SOIL_LIQ1", "SOIL_LIQ2", "DELSOILMOIST",
                                            "DELSWE", "DELINTERCEPT", "WATER_ERROR"]
        assert [c[1] for c in fl["wb"]][:4] == ["%.4f", "%.6f", "%.4f", "%.5e"] and fl["wb"][-1][1] == "%.3e"
        names = sel.step_vars(aggregate=True)
        assert names.count("AERO_COND") == 1 and "AERO_RESIST" not in names and sel.wants("AERO_RESIST") and sel.wants("AERO_COND")
    if variant == "bin24sel":
        assert [(c[0], c[2], c[3]) for c in fl["packed"]] == [
            ("PREC", "USINT", 40.0), ("AIR_TEMP", "SINT", 100.0), ("SWE", "FLOAT", 1.0), ("SOIL_LIQ0", "DOUBLE", 1.0),
            ("SOIL_LIQ1", "DOUBLE", 1.0), ("SOIL_LIQ2", "DOUBLE", 1.0), ("RUNOFF", "FLOAT", 1000.0), ("AERO_RESIST", "FLOAT", 1.0)]
        # the reference's own file: 3 date ints + 2 + 2 + 4 + 24 + 4 + 4 bytes per daily record
        n = os.listdir(ref)[0]
        assert os.path.getsize(os.path.join(ref, n)) == (g.nrecs // 8) * (12 + 40)
        d, v = files.read_binary_results(os.path.join(ref, n), fl["packed"], hourly=False)
        assert tuple(d[0]) == (1986, 1, 1) and v.shape == (8, g.nrecs // 8)


def test_unknown_output_variable_and_bad_out_step_are_refused(tmp_path):
    gpath, _, _ = _stage("sub3h", tmp_path, "out6sel")
    txt = open(gpath).read()
    open(gpath, "w").write(txt + "OUTVAR OUT_LAKE_DEPTH\n")
    with pytest.raises(Exception, match="OUT_LAKE_DEPTH"):
        files.OutSel(gpath, 3)
    open(gpath, "w").write(txt + "OUT_STEP 4\n")          # not a multiple of TIME_STEP 3 (get_global_param.c:754-760)
    assert "OUT_STEP" in files.Global(gpath).unsupported
    open(gpath, "w").write(txt.replace("N_OUTFILES 2", "N_OUTFILES 1"))
    with pytest.raises(Exception, match="N_OUTFILES"):
        files.OutSel(gpath, 3)


@pytest.mark.gpu
def test_aggregation_kernel_against_numpy():
    """vr_aggregate: END / SUM / AVG / inverted conductance in record order (put_data.c:665-688)"""
    rng = np.random.default_rng(5)
    names = ["SWE", "RUNOFF", "SURF_TEMP", "AERO_RESIST", "SOIL_LIQ1", "AERO_COND"]
    x = rng.uniform(0.1, 5.0, (len(names), 50, 301))
    for ratio in (1, 2, 8):
        got = files.aggregate(names, x, ratio)
        nout = 50 // ratio
        blk = x[:, :nout * ratio].reshape(len(names), nout, ratio, -1)
        avg = np.zeros((len(names), nout, 301))
        tot = np.zeros_like(avg)
        for r in range(ratio):
            avg += blk[:, :, r] / ratio
            tot += blk[:, :, r]
        want = np.stack([blk[0, :, -1], tot[1], avg[2], 1 / avg[3], blk[4, :, -1], avg[5]])
        assert np.array_equal(got, want), ratio


@pytest.mark.gpu
@pytest.mark.parametrize("case,variant", VARIANTS)
def test_vicnl_output_variants_reproduce_the_reference_result_files(case, variant, tmp_path):
    from vicres_b200 import vicnl
    gpath, res, ref = _stage(case, tmp_path, variant)
    assert vicnl.main(["-g", gpath]) == 0
    _compare_variant(case, variant, gpath, res, ref, "%s/%s" % (case, variant))


def test_skipyear_counts_records_like_make_dmy(tmp_path):
    """CPU tier: SKIPYEAR -> records (make_dmy.c:162-169), per year the length of the year the offset has reached"""
    gpath, _, ref = _stage("skip", tmp_path)
    g = files.Global(gpath)
    assert g.unsupported is None and g.skipyear == 1
    dmy = g.dmy()
    assert g.output_skip(dmy) == 366                        # 1984 is a leap year
    assert len(open(os.path.join(ref, sorted(os.listdir(ref))[0])).readlines()) == g.nrecs - 366
    txt = open(gpath).read()
    open(gpath, "w").write(txt.replace("STARTYEAR 1984", "STARTYEAR 1985").replace("SKIPYEAR 1", "SKIPYEAR 1\nENDYEAR 1986"))
    g = files.Global(gpath)
    assert g.output_skip(g.dmy()) == 365
    open(gpath, "w").write(txt.replace("SKIPYEAR 1", "SKIPYEAR 3"))      # beyond the run: nothing is written
    g = files.Global(gpath)
    assert g.output_skip(g.dmy()) == g.nrecs


@pytest.mark.parametrize("name", ["daily", "sub3h"])
def test_global_file_and_forcing_files_parse(name, tmp_path):
    """CPU tier: the host side of the driver up to the device hand-off."""
    from vicres_b200 import atmos
    gpath, _, ref = _stage(name, tmp_path)
    g = files.Global(gpath)
    assert g.unsupported is None
    lib, cells, meta = files.read_params(**g.param_files())
    ao = atmos.options_from_global(g)
    assert ao.layout == (atmos.DAILY_TMAX_TMIN if name == "daily" else atmos.SUBDAILY_AIR_TEMP)
    assert (ao.vp_interp, ao.vp_iter, ao.lw_type, ao.lw_cloud, ao.plapse) == (1, 1, 5, 1, 1)
    d, got = g.read_forcing(float(meta["lat"][0]), float(meta["lng"][0]))
    assert got == g.nrecs * g.time_step // g.force_dt and set(d) >= {"PREC", "WIND"}
    site = meta["site"]
    assert np.array_equal(site["lat"], meta["lat"].astype(np.float64)) and np.array_equal(site["lng"], meta["lng"].astype(np.float64))
    if name == "daily":
        assert np.all(site["time_zone_lng"] == 15.0) and np.all(site["annual_prec"] == 1000.0)
    assert len(os.listdir(ref)) == 2 * len(meta["gridcel"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["daily", "sub3h", "skip", "organic"])
def test_vicnl_command_line_reproduces_the_reference_result_files(name, tmp_path):
    from vicres_b200 import vicnl
    gpath, res, ref = _stage(name, tmp_path)
    assert vicnl.main(["-g", gpath]) == 0
    assert sorted(os.listdir(res)) == sorted(os.listdir(ref))
    nbad = ntot = 0
    for n in sorted(os.listdir(ref)):
        a, b = np.loadtxt(os.path.join(res, n)), np.loadtxt(os.path.join(ref, n))
        assert a.shape == b.shape, n
        ncal = 4 if name == "sub3h" else 3
        assert np.array_equal(a[:, :ncal], b[:, :ncal]), n
        assert np.all(np.abs(a - b) <= 1.0001e-4 + 1e-5 * np.abs(b)), n     # one unit in the 4th decimal
        nbad += int((a != b).sum())
        ntot += a.size
    print("VICNL %s: %d of %d printed values differ from the reference's files (by one unit of the last digit)" % (name, nbad, ntot))
    assert nbad <= 0.01 * ntot


def _build_c_driver(tmp_path):
    import subprocess
    from common import ROOT
    from vicres_b200 import build
    build.build()
    exe = str(tmp_path / "vicnl_gpu")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "host", "vicnl_gpu.c"), "-L", os.path.join(ROOT, "vicres_b200"), "-lvicres",
                    "-Wl,-rpath," + os.path.join(ROOT, "vicres_b200")], check=True)
    return exe


def test_c_host_driver_compiles_against_the_headers(tmp_path):
    """CPU tier: host/vicnl_gpu.c is plain C over include/*.h and links against libvicres.so; without a GPU it stops
    at the first device call with the library's message (no CPU path)."""
    import subprocess
    import torch
    exe = _build_c_driver(tmp_path)
    gpath, _, _ = _stage("daily", tmp_path)
    r = subprocess.run([exe, "-g", gpath], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CPU path" in r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["daily", "sub3h", "skip", "organic"])
def test_c_host_driver_reproduces_the_reference_result_files(name, tmp_path):
    """the same command line from the plain C host program (host/vicnl_gpu.c), records in blocks of 50"""
    import subprocess
    exe = _build_c_driver(tmp_path)
    gpath, res, ref = _stage(name, tmp_path)
    r = subprocess.run([exe, "-g", gpath, "--block", "50"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert sorted(os.listdir(res)) == sorted(os.listdir(ref))
    nbad = ntot = 0
    for n in sorted(os.listdir(ref)):
        a, b = np.loadtxt(os.path.join(res, n)), np.loadtxt(os.path.join(ref, n))
        assert a.shape == b.shape, n
        assert np.all(np.abs(a - b) <= 1.0001e-4 + 1e-5 * np.abs(b)), n
        nbad += int((a != b).sum())
        ntot += a.size
    print("VICNL_GPU (C host) %s: %d of %d printed values differ from the reference's files" % (name, nbad, ntot))
    assert nbad <= 0.01 * ntot


@pytest.mark.gpu
@pytest.mark.parametrize("case,variant", VARIANTS)
def test_c_host_driver_output_variants(case, variant, tmp_path):
    import subprocess
    exe = _build_c_driver(tmp_path)
    gpath, res, ref = _stage(case, tmp_path, variant)
    r = subprocess.run([exe, "-g", gpath, "--block", "64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    _compare_variant(case, variant, gpath, res, ref, "C host %s/%s" % (case, variant))


@pytest.mark.gpu
@pytest.mark.parametrize("host", ["python", "c"])
def test_compress_gzips_the_result_files_and_skipyear_beyond_the_run_leaves_them_empty(host, tmp_path):
    """COMPRESS TRUE (close_files.c:51, compress_files.c): every result file becomes <name>.gz with the same content;
    a SKIPYEAR past the end of the run leaves empty files like the reference's"""
    import gzip
    import subprocess
    from vicres_b200 import vicnl
    gpath, res, ref = _stage("skip", tmp_path)
    txt = open(gpath).read()
    run = (lambda: vicnl.main(["-g", gpath])) if host == "python" else \
          (lambda exe=_build_c_driver(tmp_path): subprocess.run([exe, "-g", gpath]).returncode)
    open(gpath, "w").write(txt.replace("COMPRESS FALSE", "COMPRESS TRUE"))
    assert run() == 0
    assert sorted(os.listdir(res)) == sorted(n + ".gz" for n in os.listdir(ref))
    for n in os.listdir(ref):
        a = np.loadtxt(gzip.open(os.path.join(res, n + ".gz"), "rt"))
        b = np.loadtxt(os.path.join(ref, n))
        assert a.shape == b.shape and np.all(np.abs(a - b) <= 1.0001e-4 + 1e-5 * np.abs(b)), n
    for n in os.listdir(res):
        os.remove(os.path.join(res, n))
    open(gpath, "w").write(txt.replace("SKIPYEAR 1", "SKIPYEAR 2"))
    assert run() == 0
    assert sorted(os.listdir(res)) == sorted(os.listdir(ref))
    assert all(os.path.getsize(os.path.join(res, n)) == 0 for n in os.listdir(res))


def _split_forcing(gpath, case_dir):
    """the `daily` case with its four forcing columns in TWO files: PREC, TMAX (+ a WIND column of zeros that the second file
    overrides) in FORCING1, TMIN, WIND in FORCING2 (get_global_param.c:441-470)"""
    f1, f2 = os.path.join(case_dir, "forcing"), os.path.join(case_dir, "forcing2")
    os.makedirs(f2)
    for n in os.listdir(f1):
        rows = [ln.split() for ln in open(os.path.join(f1, n)) if ln.strip()]
        open(os.path.join(f1, n), "w").write("".join("%s %s 0.0000\n" % (r[0], r[1]) for r in rows))
        open(os.path.join(f2, n), "w").write("".join("%s %s\n" % (r[2], r[3]) for r in rows))
    txt = open(gpath).read()
    i, j = txt.index("FORCING1"), txt.index("FORCEHOUR 0") + len("FORCEHOUR 0")
    date = "FORCEYEAR 1984\nFORCEMONTH 1\nFORCEDAY 1\nFORCEHOUR 0"
    desc = ("FORCING1 %s/data_\nFORCE_FORMAT ASCII\nFORCE_ENDIAN LITTLE\nN_TYPES 3\nFORCE_TYPE PREC\nFORCE_TYPE TMAX\nFORCE_TYPE WIND\n"
            "FORCE_DT 24\n%s\nFORCING2 %s/data_\nFORCE_FORMAT ASCII\nFORCE_ENDIAN LITTLE\nN_TYPES 2\nFORCE_TYPE TMIN\nFORCE_TYPE WIND\n"
            "FORCE_DT 24\n%s" % (f1, date, f2, date))
    open(gpath, "w").write(txt[:i] + desc + txt[j:])


def test_second_forcing_file_is_read_and_the_reference_agrees(tmp_path):
    """CPU tier: FORCING2 -- both descriptions parse, the merged columns are the unsplit file's (WIND from the second file),
    and the unmodified reference run on the split case writes the files of the unsplit golden case"""
    import subprocess
    from common import ROOT
    gpath, res, ref = _stage("daily", tmp_path)
    whole = {}
    g = files.Global(gpath)
    lib, cells, meta = files.read_params(**g.param_files())
    for c in range(len(meta["lat"])):
        whole[c] = g.read_forcing(float(meta["lat"][c]), float(meta["lng"][c]))[0]
    _split_forcing(gpath, os.path.dirname(gpath))
    g = files.Global(gpath)
    assert g.unsupported is None and g.forcing2 == 1 and g.force_types == ["PREC", "TMAX", "WIND"]
    g2 = g.second_forcing()
    assert g2.force_types == ["TMIN", "WIND"] and g2.forcing1.endswith("forcing2/data_") and g2.nrecs == g.nrecs
    for c in range(len(meta["lat"])):
        d, got = g.read_forcing(float(meta["lat"][c]), float(meta["lng"][c]))
        d2, got2 = g2.read_forcing(float(meta["lat"][c]), float(meta["lng"][c]))
        assert got == got2 == g.nrecs and not d["WIND"].any()
        d.update(d2)
        for k in ("PREC", "TMAX", "TMIN", "WIND"):
            assert np.array_equal(d[k], whole[c][k]), k
    vicnl_ref = os.path.join(ROOT, "oracle", "_ref", "vicNl")
    if os.path.exists(vicnl_ref):
        r = subprocess.run([vicnl_ref, "-g", gpath], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1500:]
        for n in sorted(os.listdir(ref)):
            assert open(os.path.join(res, n)).read() == open(os.path.join(ref, n)).read(), n


@pytest.mark.gpu
@pytest.mark.parametrize("host", ["python", "c"])
def test_vicnl_with_a_second_forcing_file(host, tmp_path):
    import subprocess
    from vicres_b200 import vicnl
    gpath, res, ref = _stage("daily", tmp_path)
    _split_forcing(gpath, os.path.dirname(gpath))
    if host == "python":
        assert vicnl.main(["-g", gpath]) == 0
    else:
        r = subprocess.run([_build_c_driver(tmp_path), "-g", gpath], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1500:]
    _compare_ascii(res, ref, "daily with FORCING2 (%s host)" % host)
