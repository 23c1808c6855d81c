"""Generate tests/golden/state_case/: the reference's own checkpoint / restart on one case.

    python tools/make_state_golden.py

Run 1 (the unmodified vicNl): 200 days from 1984-01-01 with STATENAME/STATEYEAR/STATEMONTH/STATEDAY -> the state
file of 1984-04-29 (end of day 120).  Run 2: INIT_STATE <that file>, start 1984-04-30, the remaining 80 days ->
result files.  Kept: the inputs, the state file, the results of run 1 (whole period) and of run 2 (restart).
Global files are stored with @DIR@ for the case directory."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vicres_b200 import synth  # noqa: E402

VICNL = os.path.join(ROOT, "oracle", "_ref", "vicNl")


def main():
    dst = os.path.join(ROOT, "tests", "golden", "state_case")
    shutil.rmtree(dst, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "case")
        g1 = synth.make_case(work, ncells=4, nlayer=3, nbands=2, ndays=200, startyear=1984, cold=11.0, lat_range=(38, 60),
                             overstory_frac=0.5, seed=501,
                             extra_global="STATENAME %s/state\nSTATEYEAR 1984\nSTATEMONTH 4\nSTATEDAY 29\nBINARY_STATE_FILE FALSE" % work)
        r = subprocess.run([VICNL, "-g", g1], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        shutil.move(os.path.join(work, "results"), os.path.join(work, "results_full"))
        os.makedirs(os.path.join(work, "results"))
        txt = open(g1).read()
        txt2 = []
        for ln in txt.splitlines():
            k = ln.split()[0] if ln.split() else ""
            if k in ("STATENAME", "STATEYEAR", "STATEMONTH", "STATEDAY"):
                continue
            if k == "STARTMONTH":
                ln = "STARTMONTH 4"
            elif k == "STARTDAY":
                ln = "STARTDAY 30"
            elif k == "NRECS":
                ln = "NRECS 80"
            txt2.append(ln)
        txt2.append("INIT_STATE %s/state_19840429" % work)
        g2 = os.path.join(work, "global_restart.txt")
        open(g2, "w").write("\n".join(txt2) + "\n")
        r = subprocess.run([VICNL, "-g", g2], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        shutil.move(os.path.join(work, "results"), os.path.join(work, "results_restart"))
        # the same pair with BINARY_STATE_FILE TRUE (exact doubles instead of "%f" columns)
        g1b = os.path.join(work, "global_bin.txt")
        open(g1b, "w").write(txt.replace("BINARY_STATE_FILE FALSE", "BINARY_STATE_FILE TRUE").replace("%s/state\n" % work, "%s/statebin\n" % work))
        os.makedirs(os.path.join(work, "results"))
        r = subprocess.run([VICNL, "-g", g1b], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        shutil.rmtree(os.path.join(work, "results"))
        os.makedirs(os.path.join(work, "results"))
        g2b = os.path.join(work, "global_restart_bin.txt")
        open(g2b, "w").write("\n".join(txt2[:-1] + ["BINARY_STATE_FILE TRUE", "INIT_STATE %s/statebin_19840429" % work]) + "\n")
        r = subprocess.run([VICNL, "-g", g2b], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        shutil.move(os.path.join(work, "results"), os.path.join(work, "results_restart_bin"))
        shutil.copytree(work, dst)
        for g in ("global.txt", "global_restart.txt", "global_bin.txt", "global_restart_bin.txt"):
            p = os.path.join(dst, g)
            txt_g = open(p).read().replace(work, "@DIR@")
            open(p, "w").write(txt_g)
    sz = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(dst) for f in fs)
    print("state_case: %.0f kB" % (sz / 1e3))
    print(open(os.path.join(dst, "state_19840429")).read()[:900])


if __name__ == "__main__":
    main()
