"""Multi-GPU routing check: run under torchrun (one rank per GPU, NCCL).  Every rank builds the routing model on
its own device, convolves its node range, takes its block of reservoir parameter sets; two all-gathers rebuild the
full results, which must equal the single-GPU results bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/route_multi_gpu.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import torch.distributed as dist
    from make_route_golden import random_tree
    from test_gpu_route_helpers import random_reservoir
    from vicres_b200 import route, route_shard
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    nrow, ncol, nd, nsets = 80, 100, 1461, 32
    for seed in range(2026, 2126):
        rng = np.random.default_rng(seed)
        d, out = random_tree(nrow, ncol, rng, fill=0.98)
        if (d > 0).sum() > 0.9 * nrow * ncol:
            break
    active = np.argwhere(d > 0)
    reser = np.zeros_like(d)
    for i, k in enumerate(rng.choice(len(active), 60, replace=False)):
        reser[tuple(active[k])] = i + 1
    grid = dict(dir=d, reser=reser, velo=np.full((nrow, ncol), 1.5, np.float32), diff=np.full((nrow, ncol), 800.0, np.float32),
                xmask=np.full((nrow, ncol), 12500.0, np.float32))
    box = np.zeros(12, np.float32); box[0] = 1.0
    r = route.Route(grid, int(out[1]) + 1, int(nrow - out[0]), box, device=local)
    runoff = (rng.gamma(0.5, 4.0, (nrow * ncol, nd)) * (rng.uniform(size=(nrow * ncol, nd)) < 0.4)).astype(np.float32)
    base = np.abs(rng.normal(1.0, 0.5, (nrow * ncol, nd))).astype(np.float32)
    sets = [[random_reservoir(rng, int((n + s) % 7), nd) for n in range(r.nnodes - 1)] for s in range(nsets)]
    cells = [len(r.node(n)[1]) for n in range(r.nnodes)]
    dev = torch.device("cuda", local)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    nat = route_shard.convolve_sharded(r, dist, cells, device=dev, runoff=runoff, baseflow=base, scale=0.18308)
    flow = route_shard.reservoirs_sharded(r, dist, nat, sets, 2000, 1, device=dev)
    dist.barrier(); torch.cuda.synchronize()
    t_shard = time.perf_counter() - t0
    ok = True
    if rank == 0:
        t0 = time.perf_counter()
        nat1, _ = r.convolve_in(runoff=runoff, baseflow=base, scale=0.18308)
        flow1 = r.reservoirs(nat1, sets, 2000, 1)["flow"]
        t_one = time.perf_counter() - t0
        ok = bool(np.array_equal(nat, nat1) and np.array_equal(flow, flow1, equal_nan=True))
        print(json.dumps({"world": world, "nodes": r.nnodes, "routed_cells": int(sum(cells)), "days": nd, "sets": nsets,
                          "node_ranges": route_shard.node_ranges(cells, world), "bit_identical_to_one_gpu": ok,
                          "sharded_s": t_shard, "one_gpu_s": t_one}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
