"""Generate tests/golden/vicnl_case/: a complete reference-format case (global file, parameter files, forcing files)
with the result files the UNMODIFIED reference writes for it (oracle/_ref/vicNl -g global.txt).

    python tools/make_vicnl_golden.py

Two cases: `daily` (daily forcing, 2 snow bands, cells whose local time is not the forcing's) and `sub3h` (3-hourly
AIR_TEMP forcing).  The global files are stored with @DIR@ in place of the case directory."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vicres_b200 import synth  # noqa: E402

VICNL = os.path.join(ROOT, "oracle", "_ref", "vicNl")
CASES = {
    "daily": dict(ncells=4, nlayer=3, nbands=2, ndays=240, startyear=1984, cold=9.0, lat_range=(28, 58),
                  overstory_frac=0.5, off_gmt=1, seed=401),
    "sub3h": dict(ncells=3, nlayer=3, nbands=1, ndays=60, dt=3, startyear=1986, cold=12.0, lat_range=(40, 62), seed=402),
}


def main():
    if not os.path.exists(VICNL):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    dst_root = os.path.join(ROOT, "tests", "golden", "vicnl_case")
    shutil.rmtree(dst_root, ignore_errors=True)
    for name, kw in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            work = os.path.join(tmp, "case")
            g = synth.make_case(work, **kw)
            r = subprocess.run([VICNL, "-g", g], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stderr[-2000:])
            dst = os.path.join(dst_root, name)
            shutil.copytree(work, dst)
            txt = open(os.path.join(dst, "global.txt")).read().replace(work, "@DIR@")
            open(os.path.join(dst, "global.txt"), "w").write(txt)
            n = sum(len(fs) for _, _, fs in os.walk(dst))
            sz = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(dst) for f in fs)
            print("%s: %d files, %.0f kB" % (name, n, sz / 1e3))


if __name__ == "__main__":
    main()
