"""CPU tier, world_size 2 over gloo: the multi-GPU routing layout (vicres_b200/route_shard.py) -- node ranges for
the convolution, parameter-set blocks for the reservoir loop, one all-gather each -- gives bit for bit the
single-process result.  The per-rank engine here is the fp32 oracle (same methods as route.Route)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from vicres_b200 import route_shard
from common import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


class OracleEngine:
    def __init__(self, o):
        self.o = o

    def convolve_in(self, **kw):
        return self.o.convolve_in(**kw)

    def reservoirs(self, natural, sets, y, m, legacy=False):
        return self.o.reservoirs_sets(natural, sets, y, m, legacy=legacy)


def _case():
    import route_lib
    from test_route_oracle import load_route_fixture
    from test_gpu_route_helpers import random_reservoir
    z, grid, (col, row), fac = load_route_fixture("rand20x16")
    o = route_lib.RouteOracle(grid, col, row, z["uh_box"])
    kw = dict(runoff=z["runoff"], baseflow=z["baseflow"], scale=0.18308)
    rng = np.random.default_rng(3)
    nd = z["runoff"].shape[1]
    sets = [[random_reservoir(rng, int((n + s) % 7), nd) for n in range(o.nnodes - 1)] for s in range(5)]
    return o, kw, sets


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o, kw, sets = _case()
    eng = OracleEngine(o)
    cells = [len(o.node(n)[1]) for n in range(o.nnodes)]
    nat = route_shard.convolve_sharded(eng, dist, cells, **kw)
    flow = route_shard.reservoirs_sharded(eng, dist, nat, sets, 2000, 1)
    if rank == 0:
        q.put((nat, flow))
    dist.destroy_process_group()


def test_node_ranges_cover_and_balance():
    r = route_shard.node_ranges([5, 1, 1, 100, 3, 90, 2], 3)
    assert r[0][0] == 0 and r[-1][1] == 7 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert route_shard.node_ranges([4], 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])
    assert [route_shard.set_block(5, k, 2) for k in range(2)] == [(0, 3), (3, 5)]


def test_two_rank_routing_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29641
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    nat, flow = q.get(timeout=300)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    o, kw, sets = _case()
    nat1, _ = o.convolve_in(**kw)
    flow1 = o.reservoirs_sets(nat1, sets, 2000, 1)["flow"]
    assert np.array_equal(nat, nat1) and np.array_equal(flow, flow1, equal_nan=True)
