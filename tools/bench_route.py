"""Routing-pass measurements ("routing reported separately as routed-cell-days/s").
Seeded dendritic D8 network (random spanning tree), reservoirs on random stream cells, BASELINE config-5-like
sizes by default.  Prints one JSON line; CUDA-event kernel times from the library, the fp32 oracle restatement
(oracle/route_oracle.c, one host thread) beside them on a bounded sample.

    python tools/bench_route.py [--rows 200 --cols 250 --reservoirs 200 --days 3653 --sets 64]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200)
    ap.add_argument("--cols", type=int, default=250)
    ap.add_argument("--reservoirs", type=int, default=200)
    ap.add_argument("--days", type=int, default=3653)
    ap.add_argument("--sets", type=int, default=64)
    ap.add_argument("--cpu-days", type=int, default=365)
    args = ap.parse_args()
    from make_route_golden import random_tree
    from test_gpu_route_helpers import random_reservoir
    from vicres_b200 import route
    import route_lib
    nrow, ncol, nd = args.rows, args.cols, args.days
    for seed in range(2026, 2126):                 # the tree generator dies early for some seeds
        rng = np.random.default_rng(seed)
        d, out = random_tree(nrow, ncol, rng, fill=0.98)
        if (d > 0).sum() > 0.9 * nrow * ncol:
            break
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    for i, k in enumerate(rng.choice(len(active), min(args.reservoirs, len(active)), replace=False)):
        reser[tuple(active[k])] = i + 1
    grid = dict(dir=d, reser=reser, velo=np.full((nrow, ncol), 1.5, np.float32), diff=np.full((nrow, ncol), 800.0, np.float32),
                xmask=np.full((nrow, ncol), 12500.0, np.float32))
    box = np.zeros(12, np.float32); box[0] = 1.0
    col, row = int(out[1]) + 1, int(nrow - out[0])
    t0 = time.perf_counter()
    r = route.Route(grid, col, row, box)
    t_build = time.perf_counter() - t0
    ncell_routed = sum(len(r.node(n)[1]) for n in range(r.nnodes))
    runoff = (rng.gamma(0.5, 4.0, (nrow * ncol, nd)) * (rng.uniform(size=(nrow * ncol, nd)) < 0.4)).astype(np.float32)
    base = np.abs(rng.normal(1.0, 0.5, (nrow * ncol, nd))).astype(np.float32)
    r.convolve(runoff, base)                                     # warm-up
    t0 = time.perf_counter()
    nat = r.convolve(runoff, base)
    t_conv_e2e = time.perf_counter() - t0
    conv_ms = r.last_kernel_ms()
    nres = r.nnodes - 1
    targets = [r.res_info(n)[1] for n in range(nres)]
    n_up = np.bincount(np.array(targets), minlength=r.nnodes)
    sets = [[random_reservoir(rng, int((n + s) % 7), nd) for n in range(nres)] for s in range(args.sets)]
    r.reservoirs(nat, sets[:1], 2000, 1)
    t0 = time.perf_counter()
    r.reservoirs(nat, sets, 2000, 1)
    t_res_e2e = time.perf_counter() - t0
    res_ms = r.last_kernel_ms()
    r.reservoirs(nat, sets[:1], 2000, 1)
    res_ms_one = r.last_kernel_ms()
    # CPU restatement (one host thread) on a bounded sample: a 40x50 network of the same kind, fewer days, one
    # parameter set; rates per routed cell-day / reservoir-day are size independent
    for seed in range(7, 107):
        rng2 = np.random.default_rng(seed)
        d2, out2 = random_tree(40, 50, rng2, fill=0.98)
        if (d2 > 0).sum() > 0.9 * 2000:
            break
    act2 = np.argwhere(d2 > 0)
    res2 = np.zeros_like(d2)
    for i, k in enumerate(rng2.choice(len(act2), 16, replace=False)):
        res2[tuple(act2[k])] = i + 1
    g2 = dict(dir=d2, reser=res2, velo=np.full((40, 50), 1.5, np.float32), diff=np.full((40, 50), 800.0, np.float32),
              xmask=np.full((40, 50), 12500.0, np.float32))
    t0 = time.perf_counter()
    o = route_lib.RouteOracle(g2, int(out2[1]) + 1, int(40 - out2[0]), box)
    t_cpu_build = time.perf_counter() - t0
    n2 = sum(len(o.node(n)[1]) for n in range(o.nnodes))
    cd = min(args.cpu_days, nd)
    ro2 = runoff[:2000, :cd].copy(); ba2 = base[:2000, :cd].copy()
    t0 = time.perf_counter()
    onat = o.convolve(ro2, ba2)
    t_cpu_conv = time.perf_counter() - t0
    s2 = [random_reservoir(rng2, int(n % 7), cd) for n in range(o.nnodes - 1)]
    t0 = time.perf_counter()
    o.reservoirs(onat, s2, 2000, 1)
    t_cpu_res = time.perf_counter() - t0
    cpu_nres = o.nnodes - 1
    line = {
        "grid": "%dx%d" % (nrow, ncol), "routed_cells_in_catchment_lists": ncell_routed, "nodes": r.nnodes, "days": nd,
        "uh_build_s (graph walk + k_uhm + k_grid_uh for every cell and reservoir)": t_build,
        "convolve": {"kernel_ms": conv_ms, "routed_cell_days_per_s_kernel": ncell_routed * nd / (conv_ms * 1e-3),
                     "e2e_s_with_host_copies": t_conv_e2e,
                     "cpu_restatement_routed_cell_days_per_s": n2 * cd / t_cpu_conv},
        "reservoirs": {"sets": args.sets, "kernel_ms": res_ms, "kernel_ms_one_set": res_ms_one,
                       "upstream_reservoirs_per_node": {"station": int(n_up[-1]), "max_reservoir": int(n_up[:-1].max()),
                                                        "reservoirs_with_upstream": int((n_up[:-1] > 0).sum())}, "reservoir_days_per_s_kernel": args.sets * nres * nd / (res_ms * 1e-3),
                       "e2e_s_with_host_copies": t_res_e2e, "cpu_restatement_reservoir_days_per_s": cpu_nres * cd / t_cpu_res},
        "cpu_sample": "40x50 network, %d routed cells, %d reservoirs, %d days, one thread; UH build %.1f s" % (n2, cpu_nres, cd, t_cpu_build),
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
