"""CPU tier: the multi-GPU layout of bench.py / a sharded run -- contiguous cell blocks per rank, no
data-path collective, one MAX all-reduce for timing -- exercised with world_size 2 over gloo.  Each
rank runs the ORACLE on its own cell block (the CUDA path needs a GPU); the union must equal the
single-process result bit for bit, which is what makes cell sharding exact (cells never interact:
runoff/model/vicNl.c:197-356)."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from vicres_b200 import abi, shard, synth_soa


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path):
    import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, nrec = 40, 60
    lib, cells = synth_soa.make_cells(C, nbands=2, seed=3, lat_range=(30.0, 60.0))
    f, month, doy = synth_soa.make_forcing(cells, nrec, seed=4, cold=8.0)
    c0, c1 = shard.cell_block(C, rank, world)
    sub = shard.slice_cells(cells, c0, c1)
    opt = abi.default_options(3, 2, 24)
    o = oracle_lib.Oracle(opt, lib, sub)
    fs = np.ascontiguousarray(f[:, :, c0:c1])
    o.init_state(fs[0, 0])
    rc, out, _ = o.step(month, doy, fs)
    assert rc == 0
    # the only collective of a sharded run: max over ranks of the elapsed time (here a dummy value)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    gathered = [None] * world
    dist.all_gather_object(gathered, (c0, c1, out))
    if rank == 0:
        full = np.zeros((opt.n_out, nrec, C))
        for a, b, part in gathered:
            full[:, :, a:b] = part
        np.save(path, full)
    dist.barrier()
    dist.destroy_process_group()


def test_cell_block_partition_covers_all_cells():
    for C in (1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            blocks = [shard.cell_block(C, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == C
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_run_equals_single_process(tmp_path):
    import oracle_lib
    path = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(2, _free_port(), path), nprocs=2, join=True)
    full = np.load(path)
    C, nrec = 40, 60
    lib, cells = synth_soa.make_cells(C, nbands=2, seed=3, lat_range=(30.0, 60.0))
    f, month, doy = synth_soa.make_forcing(cells, nrec, seed=4, cold=8.0)
    opt = abi.default_options(3, 2, 24)
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(f[0, 0])
    _, ref, _ = o.step(month, doy, f)
    assert np.array_equal(full, ref)


def _atmos_worker(rank, world, port, path):
    import atmos_lib
    from vicres_b200 import atmos
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, nd = 9, 150
    site = synth_soa.make_site(C, seed=5, lat_range=(-45.0, 62.0), own_time_zone=False)
    raw = synth_soa.make_raw_forcing(site["lat"], site["elevation"], nd, seed=6, cold=4.0)
    opt = atmos.options(time_step=24, nrecs=nd, startyear=1988)
    c0, c1 = shard.cell_block(C, rank, world)
    s, r = shard.slice_atmos_inputs(site, raw, c0, c1)
    part = atmos_lib.prepare(opt, s, r)
    gathered = [None] * world
    dist.all_gather_object(gathered, (c0, c1, part))
    if rank == 0:
        full = np.zeros((abi.VR_NFORCING, nd, C))
        for a, b, p_ in gathered:
            full[:, :, a:b] = p_
        np.save(path, full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_forcing_preparation_equals_single_process(tmp_path):
    """the forcing preparation shards like the step: per-rank cell blocks (ragged: 9 cells over 2 ranks), no collective"""
    import atmos_lib
    from vicres_b200 import atmos
    path = str(tmp_path / "atmos.npy")
    mp.spawn(_atmos_worker, args=(2, _free_port(), path), nprocs=2, join=True)
    C, nd = 9, 150
    site = synth_soa.make_site(C, seed=5, lat_range=(-45.0, 62.0), own_time_zone=False)
    raw = synth_soa.make_raw_forcing(site["lat"], site["elevation"], nd, seed=6, cold=4.0)
    ref = atmos_lib.prepare(atmos.options(time_step=24, nrecs=nd, startyear=1988), site, raw)
    assert np.array_equal(np.load(path), ref)
