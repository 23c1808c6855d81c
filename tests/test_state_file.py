"""Checkpoint / restart in the reference's own state-file format (vr_write_state_file, vr_read_state_file) against the
UNMODIFIED reference: tests/golden/state_case holds a 200-day run of vicNl that saves its state after day 120
(STATENAME / STATEYEAR / STATEMONTH / STATEDAY) and a second run that restarts from that file (INIT_STATE) for the
remaining 80 days (tools/make_state_golden.py)."""
import os
import shutil
import numpy as np
import pytest
from common import GOLDEN
from vicres_b200 import files

CASE = os.path.join(GOLDEN, "state_case")


def _stage(tmp_path):
    dst = tmp_path / "case"
    shutil.copytree(CASE, dst)
    for g in ("global.txt", "global_restart.txt", "global_bin.txt", "global_restart_bin.txt"):
        p = dst / g
        p.write_text(p.read_text().replace("@DIR@", str(dst)))
    shutil.rmtree(dst / "results", ignore_errors=True)
    os.makedirs(dst / "results")
    return dst


def _rows(path):
    return [ln.split() for ln in open(path) if ln.strip()]


def test_global_file_state_options_parse(tmp_path):
    dst = _stage(tmp_path)
    g = files.Global(str(dst / "global.txt"))
    assert g.unsupported is None and not g.init_state
    assert g.statename == str(dst / "state") and (g.stateyear, g.statemonth, g.stateday, g.binary_state_file) == (1984, 4, 29, 0)
    r = files.Global(str(dst / "global_restart.txt"))
    assert r.unsupported is None and r.init_state == 1 and r.init_state_file == str(dst / "state_19840429")
    assert (r.startmonth, r.startday, r.nrecs) == (4, 30, 80)


def _compare_results(res, ref, ncal=3):
    nbad = ntot = 0
    for n in sorted(os.listdir(ref)):
        a, b = np.loadtxt(os.path.join(res, n)), np.loadtxt(os.path.join(ref, n))
        assert a.shape == b.shape, n
        assert np.array_equal(a[:, :ncal], b[:, :ncal]), n
        assert np.all(np.abs(a - b) <= 1.0001e-4 + 1e-5 * np.abs(b)), n     # one unit in the 4th decimal
        nbad += int((a != b).sum())
        ntot += a.size
    return nbad, ntot


@pytest.mark.gpu
def test_written_state_file_equals_the_reference_one(tmp_path):
    """the whole 200-day run on the GPU with STATENAME: the state file it leaves after day 120 against vicNl's"""
    from vicres_b200 import vicnl
    dst = _stage(tmp_path)
    ref_state = _rows(os.path.join(CASE, "state_19840429"))
    os.remove(dst / "state_19840429")
    assert vicnl.main(["-g", str(dst / "global.txt")]) == 0
    mine = _rows(str(dst / "state_19840429"))
    assert len(mine) == len(ref_state)
    nval = nbad = 0
    for a, b in zip(mine, ref_state):
        assert len(a) == len(b), (a, b)
        for x, y in zip(a, b):
            if "." not in y:
                assert x == y, (a, b)                      # cell numbers, tile / band indices, last_snow, MELTING
            else:
                nval += 1
                if x != y:
                    nbad += 1
                    assert abs(float(x) - float(y)) <= 2e-6 + 1e-6 * abs(float(y)), (a, b)
    print("STATE FILE: %d of %d printed values differ from the reference's file (last printed digit)" % (nbad, nval))
    assert nbad <= 0.01 * nval
    nb, nt = _compare_results(str(dst / "results"), os.path.join(CASE, "results_full"))
    assert nb <= 0.01 * nt


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["reference", "own"])
def test_restart_from_a_state_file_reproduces_the_reference_restart(tmp_path, source):
    """INIT_STATE on the GPU from the REFERENCE's state file (and from the file the GPU run wrote itself): the 80
    remaining days against the result files of vicNl restarted from its own state file"""
    from vicres_b200 import vicnl
    dst = _stage(tmp_path)
    if source == "own":
        os.remove(dst / "state_19840429")
        assert vicnl.main(["-g", str(dst / "global.txt")]) == 0
        shutil.rmtree(dst / "results")
        os.makedirs(dst / "results")
    assert vicnl.main(["-g", str(dst / "global_restart.txt")]) == 0
    nb, nt = _compare_results(str(dst / "results"), os.path.join(CASE, "results_restart"))
    print("RESTART (%s state file): %d of %d printed values differ from the reference's restarted run" % (source, nb, nt))
    assert nb <= 0.01 * nt


@pytest.mark.gpu
def test_binary_state_file_restart_and_layout(tmp_path):
    """BINARY_STATE_FILE TRUE: restart on the GPU from the reference's binary state file against vicNl restarted from it;
    and the binary file the GPU run writes has the reference's byte layout (same size, same integers, doubles to 1e-9)"""
    from vicres_b200 import vicnl
    dst = _stage(tmp_path)
    assert vicnl.main(["-g", str(dst / "global_restart_bin.txt")]) == 0
    nb, nt = _compare_results(str(dst / "results"), os.path.join(CASE, "results_restart_bin"))
    print("RESTART (reference binary state file): %d of %d printed values differ" % (nb, nt))
    assert nb <= 0.01 * nt
    ref = open(os.path.join(CASE, "statebin_19840429"), "rb").read()
    os.remove(dst / "statebin_19840429")
    shutil.rmtree(dst / "results")
    os.makedirs(dst / "results")
    assert vicnl.main(["-g", str(dst / "global_bin.txt")]) == 0
    mine = open(dst / "statebin_19840429", "rb").read()
    assert len(mine) == len(ref)
    # walk the reference's layout: header 5 ints; per cell 4 ints + 6 doubles; per (veg, band) 2 ints, 2L doubles,
    # [Wdew], int, char, 9 doubles, 3 doubles
    import struct
    off, L = 20, 3
    assert mine[:20] == ref[:20]
    ndbl = nbad = 0
    while off < len(ref):
        cell, nveg, nband, nbytes = struct.unpack_from("<4i", ref, off)
        assert struct.unpack_from("<4i", mine, off) == (cell, nveg, nband, nbytes)
        off += 16
        spans = [(off, 6)]
        off += 48
        for veg in range(nveg + 1):
            for b in range(nband):
                assert mine[off:off + 8] == ref[off:off + 8]
                off += 8
                n = 2 * L + (1 if veg < nveg else 0)
                spans.append((off, n)); off += 8 * n
                assert mine[off:off + 5] == ref[off:off + 5]          # last_snow, MELTING
                off += 5
                spans.append((off, 12)); off += 96
        for o, n in spans:
            a, r = np.frombuffer(mine, "<f8", n, o), np.frombuffer(ref, "<f8", n, o)
            ndbl += n
            nbad += int((np.abs(a - r) > 1e-9 * np.maximum(np.abs(r), 1e-3)).sum())
    print("BINARY STATE FILE: %d of %d doubles beyond 1e-9 of the reference's" % (nbad, ndbl))
    assert nbad <= 0.01 * ndbl
