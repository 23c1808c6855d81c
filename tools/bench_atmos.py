"""Time the forcing preparation on one GPU: raw daily columns of C cells x D days resident in HBM ->
nine atmos fields in HBM (vr_atmos_prepare_device).  Prints one JSON line.
    python tools/bench_atmos.py [--cells 100000] [--days 3653] [--dt 24] [--reps 3] [--cpu-cells 16]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def measure(cells=100000, days=3653, dt=24, reps=3, cpu_cells=16, batch=0, device=0):
    """Run vr_atmos_prepare_device `reps` times (after one warm-up) on raw columns resident in HBM; returns a dict."""
    import torch
    from vicres_b200 import atmos, synth_soa
    C, nd = cells, days
    spd = 24 // dt
    site = synth_soa.make_site(C, seed=41)
    # raw columns: one year generated, tiled over the period (values do not matter for the timing)
    yr = synth_soa.make_raw_forcing(site["lat"], site["elevation"], 365, dt_hours=dt, seed=42)
    tiles = -(-nd // 365)
    dev = torch.device("cuda", device)
    raw_t = {k: (torch.from_numpy(v).to(dev).repeat(tiles, 1)[:nd * spd].contiguous() if v is not None else None) for k, v in yr.items()}
    site_t = {k: torch.from_numpy(v).to(dev) for k, v in site.items()}
    nrecs = nd * spd
    out = torch.empty((9, nrecs, C), dtype=torch.float64, device=dev)
    opt = atmos.options(time_step=dt, nrecs=nrecs, startyear=1981, layout=atmos.DAILY_TMAX_TMIN if dt == 24 else atmos.SUBDAILY_AIR_TEMP,
                        force_dt=dt)
    opt.reserved[0] = batch
    ptr = lambda t: t.data_ptr() if t is not None else 0
    ms_all = []
    for r in range(reps + 1):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        atmos.prepare_device(opt, {k: ptr(v) for k, v in site_t.items()}, {k: ptr(v) for k, v in raw_t.items()},
                             [out[f].data_ptr() for f in range(9)], C, nrecs * dt // opt.force_dt, device=device)
        torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) * 1e3
        k = atmos.last_times()
        if r > 0 or reps == 0:
            ms_all.append((wall, k[0], k[1], k[2], k[3]))
    ms = np.array(ms_all).mean(axis=0)
    finite = bool(torch.isfinite(out).all().item())
    res = {"workload": "forcing preparation (initialize_atmos + MTCLIM), %d cells x %d days, dt %d h, raw columns and results resident in HBM" % (C, nd, dt),
           "cells": C, "days": nd, "records": nrecs, "wall_ms": float(ms[0]), "k_solar_ms": float(ms[1]), "mtclim_daily_ms": float(ms[2]),
           "hours_and_records_ms": float(ms[3]), "launches": int(ms[4]), "cell_days_per_s": C * nd / (ms[0] * 1e-3),
           "cell_records_per_s": C * nrecs / (ms[0] * 1e-3), "outputs_finite": finite,
           "out_bytes": 9 * nrecs * C * 8, "in_bytes": int(sum(v.numel() * 8 for v in raw_t.values() if v is not None))}
    if cpu_cells > 0:
        import atmos_lib
        n = cpu_cells
        s = {k: v[:n] for k, v in site.items()}
        rw = {k: (np.ascontiguousarray(v[:, :n].cpu().numpy()) if v is not None else None) for k, v in raw_t.items()}
        t0 = time.perf_counter()
        ref = atmos_lib.prepare(opt, s, rw)
        cpu_s = time.perf_counter() - t0
        got = out[:, :, :n].cpu().numpy()
        e = np.abs(got - ref) / np.maximum(np.maximum(np.abs(got), np.abs(ref)), 1e-3)
        res["cpu_port"] = {"cells": n, "seconds": cpu_s, "cell_days_per_s_one_core": n * nd / cpu_s,
                           "frac_within_1e-9_vs_gpu": float((e <= 1e-9).mean()), "max_rel": float(e.max())}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--days", type=int, default=3653)
    ap.add_argument("--dt", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-cells", type=int, default=16)
    ap.add_argument("--batch", type=int, default=0)
    a = ap.parse_args()
    print(json.dumps(measure(a.cells, a.days, a.dt, a.reps, a.cpu_cells, a.batch)))


if __name__ == "__main__":
    main()
