"""BASELINE config 5 on one GPU, end to end: raw forcing columns -> forcing preparation -> cell step -> in-memory flux
hand-off -> unit-hydrograph convolution -> reservoir day loop -> station discharge.  Seeded dendritic D8 network on a
rows x cols grid with one VIC cell per grid node, reservoirs on random stream cells.  Prints one JSON line with the time
of every stage (host wall clock around the public calls, CUDA-event kernel times where the library reports them).

    python tools/bench_pipeline.py [--rows 200 --cols 250 --reservoirs 200 --days 365 --sets 8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def run(argv=None):
    """runs the chain once; returns the line's dict"""
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200)
    ap.add_argument("--cols", type=int, default=250)
    ap.add_argument("--reservoirs", type=int, default=200)
    ap.add_argument("--days", type=int, default=365)
    ap.add_argument("--sets", type=int, default=8)
    ap.add_argument("--vic-cells", type=int, default=0,
                    help="cells of the runoff grid (BASELINE config 5: 1 000 000); the first rows x cols of them are the routed "
                         "network's cells, the others lie outside the basin.  0 = one VIC cell per grid node")
    ap.add_argument("--block", type=int, default=73, help="records per vr_step_block_device call when --vic-cells is set")
    a = ap.parse_args(argv)
    import torch
    from make_route_golden import random_tree
    from test_gpu_route_helpers import random_reservoir
    from vicres_b200 import abi, atmos, couple, files, model, route, route_abi, synth_soa
    nrow, ncol, nd = a.rows, a.cols, a.days
    for seed in range(2026, 2126):
        rng = np.random.default_rng(seed)
        d, out = random_tree(nrow, ncol, rng, fill=0.98)
        if (d > 0).sum() > 0.9 * nrow * ncol:
            break
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    for i, k in enumerate(rng.choice(len(active), min(a.reservoirs, len(active)), replace=False)):
        reser[tuple(active[k])] = i + 1
    grid = dict(dir=d, reser=reser, velo=np.full((nrow, ncol), 1.5, np.float32), diff=np.full((nrow, ncol), 800.0, np.float32),
                xmask=np.full((nrow, ncol), 12500.0, np.float32))
    box = np.zeros(12, np.float32); box[0] = 1.0
    hdr = {"xllcorner": "100.0", "yllcorner": "10.0", "cellsize": "0.125"}
    CR = nrow * ncol                                      # routed cells: one VIC cell per grid node, file order
    C = max(a.vic_cells, CR)
    lib, cells = synth_soa.make_cells(C, nlayer=3, nbands=1, seed=11, lat_range=(10.0, 35.0))
    site = synth_soa.make_site(C, seed=12, lat_range=(10.0, 35.0))
    site["lat"], site["elevation"] = cells["lat"], cells["elevation"]
    raw = synth_soa.make_raw_forcing(site["lat"], site["elevation"], nd, seed=13)
    names = abi.default_out_vars(3)
    opt = abi.default_options(3, 1, 24, out=names)
    dev = torch.device("cuda", 0)
    T = {}
    t_all = time.perf_counter()

    # 1. forcing preparation (host columns in, device fields out)
    t0 = time.perf_counter()
    ao = atmos.options(time_step=24, nrecs=nd, startyear=1990)
    site_t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in site.items()}
    raw_t = {k: (torch.from_numpy(v).to(dev) if v is not None else None) for k, v in raw.items()}
    f_dev = torch.empty((abi.VR_NFORCING, nd, C), dtype=torch.float64, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else 0
    atmos.prepare_device(ao, {k: ptr(v) for k, v in site_t.items()}, {k: ptr(v) for k, v in raw_t.items()},
                         [f_dev[f].data_ptr() for f in range(abi.VR_NFORCING)], C, nd)
    torch.cuda.synchronize(dev)
    T["forcing_prep_s"] = time.perf_counter() - t0
    T["forcing_prep_kernel_ms"] = sum(atmos.last_times()[:3])
    T["forcing_prep_launches"] = int(atmos.last_times()[3])

    # 2. cell step, device resident
    t0 = time.perf_counter()
    m = model.VicRes(opt, lib, cells, device=0)
    m.init_state(f_dev[0, 0].cpu().numpy())
    T["model_setup_s"] = time.perf_counter() - t0
    doy = (np.arange(nd) % 365 + 1).astype(np.int32)
    month = np.minimum((doy - 1) // 31 + 1, 12).astype(np.int32)
    idx = np.full(C, -1, dtype=np.int64)                  # VIC cell -> routing grid node (-1: outside the basin)
    idx[:CR] = np.arange(CR)
    if C == CR:
        t0 = time.perf_counter()
        out_dev = torch.empty((opt.n_out, nd, C), dtype=torch.float64, device=dev)
        m.step_device(month, doy, [f_dev[f].data_ptr() for f in range(abi.VR_NFORCING)], out_dev.data_ptr())
        m.synchronize()
        T["step_s"] = time.perf_counter() - t0
        # 3. hand-off on the device: the eight routing columns through the "%.4f" round trip the flux files impose,
        #    transposed to [cell][day] float32 (vr_route_fluxes_from_step); one VIC cell per grid node, same order
        t0 = time.perf_counter()
        cols, present_dev = route.fluxes_from_step_device(out_dev, names, idx, CR)
        torch.cuda.synchronize(dev)
        T["handoff_s"] = time.perf_counter() - t0
    else:
        # the runoff grid is larger than the routed basin: records in blocks, the basin's columns handed off per block
        T["step_s"], T["handoff_s"] = 0.0, 0.0
        parts, present_dev = [], None
        out_dev = torch.empty((opt.n_out, a.block, C), dtype=torch.float64, device=dev)
        for r0 in range(0, nd, a.block):
            nr = min(a.block, nd - r0)
            t0 = time.perf_counter()
            ob = out_dev[:, :nr] if nr == a.block else torch.empty((opt.n_out, nr, C), dtype=torch.float64, device=dev)
            m.step_device(month[r0:r0 + nr], doy[r0:r0 + nr], [f_dev[f, r0:r0 + nr].data_ptr() for f in range(abi.VR_NFORCING)],
                          ob.data_ptr())
            m.synchronize()
            T["step_s"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            cb, present_dev = route.fluxes_from_step_device(ob, names, idx, CR)
            parts.append(cb)
            torch.cuda.synchronize(dev)
            T["handoff_s"] += time.perf_counter() - t0
        cols = {k: torch.cat([p[k] for p in parts], dim=1).contiguous() for k in parts[0]}
    T["step_cell_steps_per_s"] = C * nd / T["step_s"]
    T["step_launches"] = int(m.num_launches)
    m.close()

    # 4. routing: unit hydrographs, convolution, reservoirs
    t0 = time.perf_counter()
    r = route.Route(grid, int(out[1]) + 1, int(nrow - out[0]), box)
    T["uh_build_s"] = time.perf_counter() - t0
    fac = route_abi.cell_factors(hdr, nrow, ncol)
    t0 = time.perf_counter()
    ptrs = {k: cols[k].data_ptr() for k in ("runoff", "baseflow", "ev_vege", "ev_soil", "tr_vege")}
    ptrs["present"] = present_dev.data_ptr()
    nat = r.convolve_in_device(nd, ptrs, cell_factor=fac)
    T["convolve_s"] = time.perf_counter() - t0
    T["convolve_kernel_ms"] = r.last_kernel_ms()
    nres = r.nnodes - 1
    sets = [[random_reservoir(rng, int((n + s) % 7), nd) for n in range(nres)] for s in range(a.sets)]
    t0 = time.perf_counter()
    res = r.reservoirs(nat, sets, 1990, 1)
    T["reservoirs_s"] = time.perf_counter() - t0
    T["reservoirs_kernel_ms"] = r.last_kernel_ms()
    routed = sum(len(r.node(n)[1]) for n in range(r.nnodes))
    r.close()
    T["total_s"] = time.perf_counter() - t_all
    line = {"workload": "config 5 chain: %d VIC cells, the first %d of them on a %dx%d D8 network (%d routed cells in the catchment lists), %d reservoirs "
                        "x %d rule sets, %d days" % (C, CR, nrow, ncol, routed, nres, a.sets, nd),
            "stages": T, "station_flow_finite": bool(np.isfinite(res["flow"]).all()),
            "mean_station_discharge_m3s": float(res["flow"].mean()),
            "cells": int(C), "days": int(nd)}
    return line


def main():
    print(json.dumps(run()))


if __name__ == "__main__":
    main()
