"""`rout <configuration file>` on the GPU: the reference's routing command line (routing/model/rout.f:125-566)
over the C ABI of include/vicres_route.h.

    python -m vicres_b200.rout ../parameters/configuration.txt [--device 0] [--legacy]

Reads what rout.f reads -- the positional configuration file (rout.f:131-317; today's layout with the grid-location
block and the VIC driver line, or the older layout of the shipped prebuilt binary without them), the ESRI-style
flow direction / velocity / diffusion / xmask / fraction grids, stations.txt, reservoirlocation.txt, resN.txt,
UH.all and the `fluxes_<lat>_<lng>` files of the runoff model -- runs unit hydrographs, convolution and the
reservoir day loop on the device, and writes `<OUTPATH><station>.day` and `<OUTPATH>reservoir_<id>_<station>.day`
(write_routines.f:64, 200; list-directed records, so whitespace-separated columns as the toolbox parses them).
File handling only; every number comes from the CUDA library.  Not built: STEPBYSTEP mode, monthly/yearly summaries.
"""
import argparse
import os
import sys
import numpy as np
from . import couple, route, route_abi

DAYS_IN_MONTH = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def _isleap(y):
    return 1 if (y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)) else 0


def _torf(s):
    return s.strip().lower().startswith((".t", "t"))


def read_configuration(path):
    """rout.f:131-317, positional.  Comment lines are part of the layout (the Fortran skips them by position)."""
    L = [ln.rstrip("\n") for ln in open(path)]
    i = [0]

    def nxt():
        v = L[i[0]]
        i[0] += 1
        return v.strip()
    cfg = {}
    nxt(); nxt()                                   # READ(1,'(//A)'): two lines skipped
    cfg["flowdirection"] = nxt()
    for key in ("velocity", "diffusion", "xmask", "fraction"):
        nxt()                                      # comment
        flag = _torf(nxt())
        v = nxt()
        cfg[key] = v if flag else float(v.split()[0])
    nxt()
    cfg["stations"] = nxt()
    # today's source reads a grid-location block here (rout.f:181-193); the shipped binary's layout does not have it
    save = i[0]
    nxt()                                          # comment
    probe = nxt()
    if probe.lower().startswith((".t", ".f")):
        cfg["layout"] = "current"
        a, b = nxt(), nxt()
        cfg["grid_location"] = (a, b) if _torf(probe) else (float(a.split()[0]), float(b.split()[0]))
        nxt()                                      # comment
        cfg["inpath"] = nxt()
        cfg["dprec"] = int(nxt().split()[0])
        cfg["vicdriver"] = nxt()
    else:
        cfg["layout"] = "binary"
        i[0] = save
        nxt()
        cfg["inpath"] = nxt()
        cfg["dprec"] = int(nxt().split()[0])
    nxt(); cfg["outpath"] = nxt()
    nxt(); cfg["reservoir_location"] = nxt()
    nxt(); cfg["reservoir_path"] = nxt()
    nxt()
    cfg["nf_exist"] = _torf(nxt())
    cfg["nf_path"] = nxt()
    nxt()
    cfg["start_year"], cfg["start_month"], cfg["stop_year"], cfg["stop_month"] = [int(x) for x in nxt().split()[:4]]
    cfg["first_year"], cfg["first_month"], cfg["last_year"], cfg["last_month"] = [int(x) for x in nxt().split()[:4]]
    nxt(); cfg["uh_path"] = nxt()
    nxt(); cfg["stepbystep"] = _torf(nxt())
    return cfg


def period(cfg):
    """(ndays, [(year, month, day)]) of START_YEAR/START_MO .. STOP_YEAR/STOP_MO (rout.f:226-247)"""
    y, m = cfg["start_year"], cfg["start_month"]
    dates = []
    for _ in range(cfg["start_month"], 12 * (cfg["stop_year"] - cfg["start_year"]) + cfg["stop_month"] + 1):
        n = DAYS_IN_MONTH[m - 1] + (_isleap(y) if m == 2 else 0)
        dates += [(y, m, d) for d in range(1, n + 1)]
        m += 1
        if m > 12:
            m, y = 1, y + 1
    return len(dates), dates


def _grid_f32(cfg, key, hdr_rows, nrow, ncol, cwd):
    v = cfg[key]
    if isinstance(v, float):
        return np.full((nrow, ncol), v, dtype=np.float32)
    with open(os.path.join(cwd, v)) as f:
        rows = [ln.split() for ln in f.read().splitlines() if ln.strip()][hdr_rows:]
    return np.array([[float(x) for x in r] for r in rows], dtype=np.float32)


def read_stations(path):
    """stations.txt: `<active> <name> <col> <row> <area>` then the name of a UH_S file or NONE (rout.f:330-360)"""
    L = [ln.split() for ln in open(path) if ln.strip()]
    out = []
    for a, b in zip(L[0::2], L[1::2]):
        if int(a[0]) == 1:
            out.append({"name": a[1], "col": int(a[2]), "row": int(a[3]), "uh_s": b[0]})
    return out


def reservoir_day_columns(series, j, res, start_year, legacy=False):
    """The nine value columns of reservoir_<id>_<station>.day (write_routines.f:172-200) of reservoir j: storage at
    the start and end of the day, full supply volume (0 until the year of commissioning, reservoir.f:1154-1161), level
    at the start and end, inflow, release, turbine flow, energy.  legacy: as the shipped prebuilt binary prints them
    (negative storages / levels / inflows / energy as 0; the turbine column is the release capped at Qdesign)."""
    n = series["flowin"].shape[1]
    full = np.full(n, np.float32(res["v_live"]) + np.float32(res["v_dead"]), dtype=np.float32)
    full[start_year + np.arange(1, n + 1) // 365 < res["ope_year"]] = 0
    c = np.stack([series["vol"][j][:-1], series["vol"][j][1:], full, series["hho"][j][:-1], series["hho"][j][1:],
                  series["flowin"][j], series["flowout"][j], series["turbine"][j], series["energy"][j]], axis=1)
    if legacy:
        c[:, [0, 3, 5, 8]] = np.maximum(c[:, [0, 3, 5, 8]], 0)
        c[:, 7] = np.where(c[:, 2] > 0, np.minimum(c[:, 6], np.float32(res["q_design"])), c[:, 7])
    return c


def run(config_path, device=0, legacy=False, cwd=None, write=True):
    """legacy: the arithmetic variant of the shipped prebuilt binary (area-based runoff scale, older rule-curve code;
    DESIGN.md section 9) instead of today's reservoir.f.  Returns {station: {"flow", "dates", "reservoirs"}}."""
    cwd = cwd or os.getcwd()
    cfg = read_configuration(config_path)
    if cfg["stepbystep"]:
        raise NotImplementedError("STEPBYSTEP mode is not built")
    P = lambda p: os.path.join(cwd, p)
    hdr, d = route_abi.read_esri_int(P(cfg["flowdirection"]))
    nrow, ncol = d.shape
    _, reser = route_abi.read_esri_int(P(cfg["reservoir_location"]), header=0)
    grid = dict(dir=d, reser=reser, velo=_grid_f32(cfg, "velocity", 6, nrow, ncol, cwd),
                diff=_grid_f32(cfg, "diffusion", 6, nrow, ncol, cwd), xmask=_grid_f32(cfg, "xmask", 6, nrow, ncol, cwd))
    frac = _grid_f32(cfg, "fraction", 6, nrow, ncol, cwd)
    box = np.loadtxt(P(os.path.join(cfg["uh_path"], "UH.all")), ndmin=2)[:route_abi.KE, 1].astype(np.float32)
    ndays, dates = period(cfg)
    inp = couple.inputs_from_files(P(cfg["inpath"]), hdr, nrow, ncol, ndays, cfg["dprec"])
    fac = route_abi.cell_factors(hdr, nrow, ncol, frac)
    results = {}
    for st in read_stations(P(cfg["stations"])):
        R = route.Route(grid, st["col"], st["row"], box, device=device)
        kw = dict(runoff=inp["runoff"], baseflow=inp["baseflow"], evap=(inp["ev_vege"], inp["ev_soil"], inp["tr_vege"]),
                  cell_factor=fac, present=inp["present"])
        if legacy:
            kw["cell_scale"] = fac * np.float32(0.028)
        nat = R.convolve_in(**kw)[0]
        ids = [R.node(n)[0] for n in range(R.nnodes - 1)]
        res = [route_abi.read_reservoir_file(P(os.path.join(cfg["reservoir_path"], "res%d.txt" % i)), cfg["start_year"],
                                             cfg["start_month"], ndays) for i in ids]
        out = R.reservoirs(nat, [res], cfg["start_year"], cfg["start_month"], legacy=legacy)
        R.close()
        flow = out["flow"][0]
        results[st["name"]] = {"flow": flow, "dates": dates, "natural": nat, "ids": ids,
                               "reservoirs": {k: v[0] for k, v in out.items() if k != "flow"}}
        if write:
            os.makedirs(P(cfg["outpath"]), exist_ok=True)
            with open(P(os.path.join(cfg["outpath"], st["name"] + ".day")), "w") as f:
                for (y, m, dd), q in zip(dates, flow):
                    f.write("%12d%12d%12d   %.9g\n" % (y, m, dd, q))
            r = results[st["name"]]["reservoirs"]
            for j, rid in enumerate(ids):
                cols = reservoir_day_columns(r, j, res[j], cfg["start_year"], legacy)
                with open(P(os.path.join(cfg["outpath"], "reservoir_%d_%s.day" % (rid, st["name"]))), "w") as f:
                    f.write(" Year Month Day Volume_1000cm Volume_t+1_1000cm FullSupplyVolume WaterLevel_m WaterLevel_t+1_m Qinflow_cms \n")
                    for k, (y, m, dd) in enumerate(dates):
                        f.write("%12d%12d%12d" % (y, m, dd) + "".join("   %.9g" % v for v in cols[k]) + "\n")
    return results


def main(argv=None):
    ap = argparse.ArgumentParser(prog="vicres_b200.rout", description="rout <configuration file> on the GPU")
    ap.add_argument("config")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--legacy", action="store_true", help="arithmetic variant of the shipped prebuilt binary")
    a = ap.parse_args(argv)
    res = run(a.config, a.device, a.legacy)
    for name, r in res.items():
        print("vicres_b200.rout: station %s, %d days, %d reservoirs" % (name, len(r["flow"]), len(r["ids"])), file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
