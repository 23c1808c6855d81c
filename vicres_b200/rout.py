"""`rout <configuration file>` on the GPU: the reference's routing command line (routing/model/rout.f:125-566)
over the C ABI of include/vicres_route.h.

    python -m vicres_b200.rout ../parameters/configuration.txt [--device 0] [--legacy]

Reads what rout.f reads -- the positional configuration file (rout.f:131-317; today's layout with the grid-location
block and the VIC driver line, or the older layout of the shipped prebuilt binary without them), the ESRI-style
flow direction / velocity / diffusion / xmask / fraction grids, stations.txt, reservoirlocation.txt, resN.txt,
UH.all and the `fluxes_<lat>_<lng>` files of the runoff model -- runs unit hydrographs, convolution and the
reservoir day loop on the device, and writes `<OUTPATH><station>.day` and `<OUTPATH>reservoir_<id>_<station>.day`
(write_routines.f:64, 200; list-directed records, so whitespace-separated columns as the toolbox parses them).
File handling only; every number of the routing comes from the CUDA library (the monthly / yearly summaries of
WRITE_MONTH are fp32 sums of the daily discharge in the reference's order).  STEPBYSTEP mode: one simulated day per run,
the state carried in the reference's text files (vr_route_reservoirs_range).
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
    # rout.f:262-316: first and current simulation day of the STEPBYSTEP version, coupler iteration, write-outputs flag
    # (absent from the shipped binary's older layout)
    cfg["sbs"], cfg["coupler_iteration"], cfg["write_output"] = None, 1, True
    try:
        nxt(); cfg["sbs"] = tuple(int(x) for x in nxt().split()[:6])
        nxt(); cfg["coupler_iteration"] = int(nxt().split()[0])
        nxt(); cfg["write_output"] = _torf(nxt())
    except (IndexError, ValueError):
        pass
    return cfg


def nday_sim(cfg):
    """NDAY_SIM of the STEPBYSTEP version: 1-based index of the simulated day (rout.f:266-287)"""
    y0, m0, d0, ys, ms, ds = cfg["sbs"]
    n, m, y = 0, m0, y0
    if ys > y0 or ms > m0:
        for _ in range(m0, 12 * (ys - y0) + ms):
            n += DAYS_IN_MONTH[m - 1] + (_isleap(y) if m == 2 else 0)
            m += 1
            if m > 12:
                m, y = 1, y + 1
    return n + ds


def factor_sum(fac, cells):
    """FACTOR_SUM of MAKE_CONVOLUTION for the station's node (reservoir.f:307, 372): the cells' FACTORs added in list
    order in fp32"""
    s = np.float32(0.0)
    for g in cells:
        s = np.float32(s + np.float32(fac[g]))
    return s


def monthly_summaries(cfg, dates, flow, fsum):
    """WRITE_MONTH (write_routines.f:248-331) in fp32 and in its order: ([(year, month, mean flow)], [(.., mm)],
    [(month, mean of the monthly means or None)], [(month, .. mm)]).  The divisor of a February is 28 whatever the year
    (DAYS_IN_MONTH is the DATA statement of rout.f:109)."""
    f32 = np.float32
    nmonths = 12 * (cfg["stop_year"] - cfg["start_year"]) + cfg["stop_month"] - cfg["start_month"] + 1
    mo, yr, m, y = [], [], cfg["start_month"], cfg["start_year"]
    for _ in range(nmonths):
        mo.append(m); yr.append(y)
        m += 1
        if m > 12:
            m, y = 1, y + 1
    monthly = [f32(0.0)] * (12 * (cfg["stop_year"] - cfg["start_year"] + 1))
    monthly_mm = list(monthly)
    M, old = 0, mo[0]
    for (yy, mm, dd), q in zip(dates, flow):
        if mm != old:
            M += 1
            old = mm
        monthly[M] = f32(monthly[M] + f32(f32(q) / f32(DAYS_IN_MONTH[mm - 1])))
        monthly_mm[M] = f32(monthly_mm[M] + f32(f32(q) / fsum))
    skipto = (cfg["first_year"] - cfg["start_year"]) * 12 + (cfg["first_month"] - cfg["start_month"]) + 1
    stopat = nmonths - ((cfg["stop_year"] - cfg["last_year"]) * 12 + (cfg["stop_month"] - cfg["last_month"]))
    yearly, yearly_mm, cnt = [f32(0.0)] * 12, [f32(0.0)] * 12, [0] * 12
    rows, rows_mm = [], []
    for i in range(skipto, stopat + 1):
        rows.append((yr[i - 1], mo[i - 1], monthly[i - 1]))
        rows_mm.append((yr[i - 1], mo[i - 1], monthly_mm[i - 1]))
        k = (i - 1) % 12
        yearly[k] = f32(yearly[k] + monthly[i - 1])
        yearly_mm[k] = f32(yearly_mm[k] + monthly_mm[i - 1])
        cnt[mo[i - 1] - 1] += 1
    yrows = [(i + 1, f32(yearly[i] / f32(cnt[i])) if cnt[i] else None) for i in range(12)]
    yrows_mm = [(i + 1, f32(yearly_mm[i] / f32(cnt[i])) if cnt[i] else None) for i in range(12)]
    return rows, rows_mm, yrows, yrows_mm


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
    sbs = bool(cfg["stepbystep"])
    if sbs and not cfg["sbs"]:
        raise ValueError("STEPBYSTEP mode needs the start / current simulation day line of the configuration file")
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
        station_cells = R.node(R.nnodes - 1)[1]
        if not sbs:
            out = R.reservoirs(nat, [res], cfg["start_year"], cfg["start_month"], legacy=legacy)
        else:
            # one day per run (reservoir.f:1142-1150, 1596-1675): VOL of the day and the releases so far come from the
            # files the previous run left, this run's go to the files of the next
            k = nday_sim(cfg)
            W = len(results) + 1
            nres = R.nnodes - 1
            sdir_nf, sdir_out = P(os.path.join(cfg["nf_path"].strip(), "stepbystep")), P(os.path.join(cfg["outpath"], "stepbystep"))
            os.makedirs(sdir_nf, exist_ok=True); os.makedirs(sdir_out, exist_ok=True)
            state = None
            if k > 1:
                state = {"flow": np.zeros((1, ndays), dtype=np.float32)}
                for key, ext in route_abi.RES_SERIES:
                    state[key] = np.zeros((1, nres, ndays + ext), dtype=np.float32)
                fo = os.path.join(sdir_nf, "OUTFLOW%d_STEP%d.txt" % (W, k))
                vo = os.path.join(sdir_out, "RESERVOIR_VOL_STATION%d_STEP%d.txt" % (W, k))
                if not (os.path.exists(fo) and os.path.exists(vo)):
                    raise FileNotFoundError("ERROR in Reading Storage Volume for Current Time Step; Please Check if Previous "
                                            "Time Step is Simulated (%s, %s)" % (fo, vo))
                if nres:
                    state["flowout"][0] = np.loadtxt(fo, ndmin=2).astype(np.float32)[:nres, :ndays]
                    state["vol"][0, :, k - 1] = np.loadtxt(vo, ndmin=1).astype(np.float32)[:nres]
            out = R.reservoirs(nat, [res], cfg["start_year"], cfg["start_month"], legacy=legacy, days=(k - 1, k), state=state)
            fmt = lambda row: " ".join("%.9g" % x for x in row) + "\n"
            with open(os.path.join(sdir_nf, "OUTFLOW%d_STEP%d.txt" % (W, k + 1)), "w") as f:
                for j in range(nres):
                    f.write(fmt(out["flowout"][0, j]))
            with open(os.path.join(sdir_nf, "NF%d_STEP%d.txt" % (W, k + 1)), "w") as f:      # RESFLOWS: inflows seen so far
                for j in range(nres):
                    f.write(fmt(np.where(np.arange(ndays) < k, out["flowin"][0, j], nat[j])))
                f.write(fmt(np.where(np.arange(ndays) < k, out["flow"][0], nat[nres])))
            steps = [k, k + 1] if k == 1 else [k + 1]          # the first step keeps its own file for coupler iterations
            for kk in steps:
                with open(os.path.join(sdir_out, "RESERVOIR_VOL_STATION%d_STEP%d.txt" % (W, kk)), "w") as f:
                    for j in range(nres):
                        f.write("%.9g\n" % out["vol"][0, j, kk - 1])
            if k > 2:
                for pth in (os.path.join(sdir_nf, "NF%d_STEP%d.txt" % (W, k - 1)), os.path.join(sdir_nf, "OUTFLOW%d_STEP%d.txt" % (W, k - 1))):
                    if os.path.exists(pth):
                        os.remove(pth)
            if k > 1:
                pth = os.path.join(sdir_out, "RESERVOIR_VOL_STATION%d_STEP%d.txt" % (W, k - 1))
                if os.path.exists(pth):
                    os.remove(pth)
        R.close()
        flow = out["flow"][0]
        fsum = factor_sum(fac.reshape(-1), station_cells)
        results[st["name"]] = {"flow": flow, "dates": dates, "natural": nat, "ids": ids,
                               "reservoirs": {k: v[0] for k, v in out.items() if k != "flow"}}
        results[st["name"]]["factor_sum"] = fsum
        if write and cfg["write_output"] and sbs:
            # WRITE_DATA's step-by-step branch (write_routines.f:42-58): one line per run, appended
            k = nday_sim(cfg)
            sd = P(os.path.join(cfg["outpath"], "STEPBYSTEP"))
            os.makedirs(sd, exist_ok=True)
            mode = "w" if (k == 1 and cfg["coupler_iteration"] == 1) else "a"
            y, m, dd = cfg["sbs"][3:6]
            with open(os.path.join(sd, st["name"] + ".day"), mode) as f:
                f.write("%12d%12d%12d   %.9g\n" % (y, m, dd, flow[k - 1]))
            with open(os.path.join(sd, st["name"] + ".day_mm"), mode) as f:
                f.write("%12d%12d%12d   %.9g\n" % (y, m, dd, np.float32(flow[k - 1]) / fsum))
        elif write and cfg["write_output"]:
            os.makedirs(P(cfg["outpath"]), exist_ok=True)
            with open(P(os.path.join(cfg["outpath"], st["name"] + ".day")), "w") as f:
                for (y, m, dd), q in zip(dates, flow):
                    f.write("%12d%12d%12d   %.9g\n" % (y, m, dd, q))
            with open(P(os.path.join(cfg["outpath"], st["name"] + ".day_mm")), "w") as f:
                for (y, m, dd), q in zip(dates, flow):
                    f.write("%12d%12d%12d   %.9g\n" % (y, m, dd, np.float32(q) / fsum))
            rows, rows_mm, yrows, yrows_mm = monthly_summaries(cfg, dates, flow, fsum)
            for ext, rr in ((".month", rows), (".month_mm", rows_mm)):
                with open(P(os.path.join(cfg["outpath"], st["name"] + ext)), "w") as f:
                    for (y, m, v) in rr:
                        f.write("%12d%12d   %.9g\n" % (y, m, v))
            for ext, rr in ((".year", yrows), (".year_mm", yrows_mm)):
                with open(P(os.path.join(cfg["outpath"], st["name"] + ext)), "w") as f:
                    for (m, v) in rr:
                        f.write("%12d   %s\n" % (m, "0" if v is None else "%.9g" % v))
            open(P(os.path.join(cfg["outpath"], st["name"] + ".end_of_month")), "w").close()
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
