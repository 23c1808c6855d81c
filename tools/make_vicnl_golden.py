"""Generate tests/golden/vicnl_case/: a complete reference-format case (global file, parameter files, forcing files)
with the result files the UNMODIFIED reference writes for it (oracle/_ref/vicNl -g global.txt).

    python tools/make_vicnl_golden.py

Three cases: `skip` (SKIPYEAR 1 over a leap year), `daily` (daily forcing, 2 snow bands, cells whose local time is not the forcing's) and `sub3h` (3-hourly
AIR_TEMP forcing).  Each case also holds VARIANTS of its global file (`global_<variant>.txt`, results in
`results_<variant>/`): the output interval, the selection of output files / variables and the binary format.
The global files are stored with @DIR@ in place of the case directory."""
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
    # SKIPYEAR 1 from 1984-01-01: the 366 days of the leap year are computed but not written
    # ORGANIC_FRACT TRUE (organic top layers) with BASEFLOW NIJSSEN2001-free ARNO columns
    "organic": dict(ncells=2, nlayer=3, nbands=1, ndays=150, startyear=1988, cold=6.0, lat_range=(40, 60), overstory_frac=0.5, seed=404,
                    organic=True),
    "skip": dict(ncells=2, nlayer=3, nbands=1, ndays=480, startyear=1984, cold=6.0, lat_range=(35, 55), seed=403, skip_output=True),
}

# variant -> (case, lines appended to the case's global file; later lines win in get_global_param.c)
VARIANTS = {
    "out24": ("sub3h", ["OUT_STEP 24"]),
    "out6sel": ("sub3h", ["OUT_STEP 6", "N_OUTFILES 2",
                          "OUTFILE wb 9", "OUTVAR OUT_PREC", "OUTVAR OUT_EVAP %.6f", "OUTVAR OUT_RUNOFF", "OUTVAR OUT_BASEFLOW %.5e",
                          "OUTVAR OUT_SOIL_MOIST", "OUTVAR OUT_DELSOILMOIST", "OUTVAR OUT_DELSWE", "OUTVAR OUT_DELINTERCEPT",
                          "OUTVAR OUT_WATER_ERROR %.3e",
                          "OUTFILE eb 8", "OUTVAR OUT_AERO_COND %.6f", "OUTVAR OUT_AERO_RESIST", "OUTVAR OUT_SURF_TEMP",
                          "OUTVAR OUT_SNOW_MELT", "OUTVAR OUT_SWE %.3f", "OUTVAR OUT_PET_SHORT", "OUTVAR OUT_ROOTMOIST",
                          "OUTVAR OUT_SHORTWAVE %.2f"]),
    "selday": ("daily", ["N_OUTFILES 1", "OUTFILE cell 7", "OUTVAR OUT_RUNOFF", "OUTVAR OUT_BASEFLOW", "OUTVAR OUT_SWE %.5f",
                         "OUTVAR OUT_AERO_RESIST", "OUTVAR OUT_SOIL_WET", "OUTVAR OUT_SNOW_COVER %.0f", "OUTVAR OUT_RAINF"]),
    # the snow model sub-stepped (NF = 8 and NF = 24 passes per daily record): initialize_atmos's per-window values and snow flags
    "snow3": ("daily", ["SNOW_STEP 3"]),
    "snow1": ("skip", ["SNOW_STEP 1"]),
    # PRT_HEADER (write_header.c), ASCII and binary; ALMA_OUTPUT units (rates, Kelvin; OUT_SUB_SNOW includes OUT_SUB_CANOP)
    "hdr": ("daily", ["PRT_HEADER TRUE"]),
    "hdrbin24": ("sub3h", ["PRT_HEADER TRUE", "BINARY_OUTPUT TRUE", "OUT_STEP 24", "N_OUTFILES 1", "OUTFILE hb 5",
                           "OUTVAR OUT_PREC %.4f OUT_TYPE_USINT 40", "OUTVAR OUT_SOIL_MOIST %.4f OUT_TYPE_FLOAT 1",
                           "OUTVAR OUT_SWE %.4f OUT_TYPE_DOUBLE *", "OUTVAR OUT_AIR_TEMP %.4f OUT_TYPE_SINT 100", "OUTVAR OUT_RUNOFF"]),
    "alma6": ("sub3h", ["ALMA_OUTPUT TRUE", "PRT_HEADER TRUE", "OUT_STEP 6", "N_OUTFILES 1", "OUTFILE al 10",
                        "OUTVAR OUT_PREC %.6e", "OUTVAR OUT_EVAP %.6e", "OUTVAR OUT_RUNOFF %.6e", "OUTVAR OUT_BASEFLOW %.6e",
                        "OUTVAR OUT_SUB_SNOW %.6e", "OUTVAR OUT_SNOW_MELT %.6e", "OUTVAR OUT_SURF_TEMP", "OUTVAR OUT_AIR_TEMP",
                        "OUTVAR OUT_DELTAH %.5e", "OUTVAR OUT_SWE"]),
    "almadef": ("daily", ["ALMA_OUTPUT TRUE"]),
    # the record's own forcing values as outputs (put_data.c:271-292), plain and in ALMA units (Pa)
    "met6": ("sub3h", ["OUT_STEP 6", "N_OUTFILES 1", "OUTFILE met 8", "OUTVAR OUT_DENSITY %.6f", "OUTVAR OUT_PRESSURE %.5f",
                       "OUTVAR OUT_VP %.6f", "OUTVAR OUT_VPD %.6f", "OUTVAR OUT_QAIR %.8f", "OUTVAR OUT_REL_HUMID",
                       "OUTVAR OUT_SHORTWAVE", "OUTVAR OUT_LONGWAVE"]),
    "metalma": ("daily", ["ALMA_OUTPUT TRUE", "N_OUTFILES 1", "OUTFILE met 5", "OUTVAR OUT_PRESSURE %.3f", "OUTVAR OUT_VP %.4f",
                          "OUTVAR OUT_VPD %.4f", "OUTVAR OUT_QAIR %.8f", "OUTVAR OUT_AIR_TEMP"]),
    # OUTPUT_FORCE: the disaggregated forcing written out, one row per snow step, no date columns (write_forcing_file.c)
    "force": ("daily", ["OUTPUT_FORCE TRUE"]),
    "forcebin": ("sub3h", ["OUTPUT_FORCE TRUE", "BINARY_OUTPUT TRUE"]),
    "force3": ("daily", ["OUTPUT_FORCE TRUE", "SNOW_STEP 3", "PRT_HEADER TRUE", "ALMA_OUTPUT TRUE", "N_OUTFILES 1", "OUTFILE fd 8",
                         "OUTVAR OUT_PREC %.6e", "OUTVAR OUT_RAINF %.6e", "OUTVAR OUT_SNOWF %.6e", "OUTVAR OUT_AIR_TEMP",
                         "OUTVAR OUT_REL_HUMID %.2f", "OUTVAR OUT_QAIR %.8f", "OUTVAR OUT_PRESSURE %.2f", "OUTVAR OUT_LONGWAVE"]),
    # MOISTFRACT: soil moisture as volume fractions (per tile, before the area weighting), the water-balance check back in mm
    "mfrac": ("daily", ["MOISTFRACT TRUE"]),
    "mfrac6": ("sub3h", ["MOISTFRACT TRUE", "OUT_STEP 6", "N_OUTFILES 1", "OUTFILE mf 5", "OUTVAR OUT_SOIL_MOIST %.8f",
                         "OUTVAR OUT_SOIL_LIQ %.8f", "OUTVAR OUT_DELSOILMOIST %.8f", "OUTVAR OUT_WATER_ERROR %.3e", "OUTVAR OUT_ROOTMOIST"]),
    "bin": ("sub3h", ["BINARY_OUTPUT TRUE"]),
    "bin24sel": ("sub3h", ["BINARY_OUTPUT TRUE", "OUT_STEP 24", "N_OUTFILES 1", "OUTFILE packed 6",
                           "OUTVAR OUT_PREC %.4f OUT_TYPE_USINT 40", "OUTVAR OUT_AIR_TEMP %.4f OUT_TYPE_SINT 100",
                           "OUTVAR OUT_SWE %.4f OUT_TYPE_FLOAT 1", "OUTVAR OUT_SOIL_LIQ %.4f OUT_TYPE_DOUBLE *",
                           "OUTVAR OUT_RUNOFF %.4f OUT_TYPE_FLOAT 1000", "OUTVAR OUT_AERO_RESIST"]),
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
            base = open(g).read()
            gfiles = ["global.txt"]
            for v, (case, lines) in VARIANTS.items():
                if case != name:
                    continue
                os.makedirs(os.path.join(work, "results_" + v))
                gv = os.path.join(work, "global_%s.txt" % v)
                open(gv, "w").write(base.replace("RESULT_DIR %s/results/" % work, "RESULT_DIR %s/results_%s/" % (work, v)) +
                                    "\n".join(lines) + "\n")
                assert "results_" + v in open(gv).read()
                r = subprocess.run([VICNL, "-g", gv], capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(r.stderr[-2000:])
                gfiles.append("global_%s.txt" % v)
            dst = os.path.join(dst_root, name)
            shutil.copytree(work, dst)
            for gf in gfiles:
                txt = open(os.path.join(dst, gf)).read().replace(work, "@DIR@")
                open(os.path.join(dst, gf), "w").write(txt)
            n = sum(len(fs) for _, _, fs in os.walk(dst))
            sz = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(dst) for f in fs)
            print("%s: %d files, %.0f kB" % (name, n, sz / 1e3))


if __name__ == "__main__":
    main()
