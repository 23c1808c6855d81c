"""Multi-GPU routing.  MAKE_CONVOLUTION treats every node (reservoir or station) on its
own -- a node's inflow is the sum over ITS catchment cells (reservoir.f:340-505) -- so the ranks take node
ranges, balanced by catchment size, and one all-gather of the [node, day] rows rebuilds the naturalised
inflows: the fp32 summation order of every node is untouched and the result is bit-identical to one GPU
(the partial-sum reduce sketched in the survey would change it).  The reservoir day loop is sequential in
time; what shards is its parameter-set axis (optimisation candidates), again with one all-gather of the
station discharges.  `engine` is a vicres_b200.route.Route (GPU) -- or, in the CPU tests, the oracle wrapper,
which has the same two methods."""
import numpy as np


def node_ranges(cells_per_node, world):
    """contiguous node ranges [(n0, n1)] * world with about equal numbers of catchment cells"""
    w = np.asarray(cells_per_node, dtype=np.float64) + 1.0
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(cum, cum[-1] * r / world, side="left"))
        cuts.append(min(max(k, cuts[-1]), len(w)))
    cuts.append(len(w))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def set_block(nsets, rank, world):
    base, rem = divmod(nsets, world)
    s0 = rank * base + min(rank, rem)
    return s0, s0 + base + (1 if rank < rem else 0)


def _all_gather_rows(dist, part, counts, device=None):
    """concatenate per-rank row blocks (rank r holds counts[r] rows) on every rank"""
    import torch
    world = dist.get_world_size()
    width = part.shape[1]
    pad = max(counts)
    buf = torch.zeros((pad, width), dtype=torch.float32, device=device)
    if part.shape[0]:
        buf[:part.shape[0]] = torch.from_numpy(np.ascontiguousarray(part)).to(buf.device)
    got = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(got, buf)
    return np.concatenate([g[:counts[r]].cpu().numpy() for r, g in enumerate(got)], axis=0)


def convolve_sharded(engine, dist, cells_per_node, device=None, **inputs):
    """every rank returns the full naturalised inflows [nnodes, ndays]"""
    rank, world = dist.get_rank(), dist.get_world_size()
    ranges = node_ranges(cells_per_node, world)
    n0, n1 = ranges[rank]
    if n1 > n0:
        flows, _ = engine.convolve_in(nodes=(n0, n1), **inputs)
        part = flows[n0:n1]
    else:
        part = np.zeros((0, np.asarray(inputs["runoff"]).shape[1]), dtype=np.float32)
    return _all_gather_rows(dist, part, [b - a for a, b in ranges], device)


def reservoirs_sharded(engine, dist, natural, sets, start_year, start_month, legacy=False, device=None):
    """station discharge of every parameter set [nsets, ndays] on every rank; the sets are split over the ranks"""
    rank, world = dist.get_rank(), dist.get_world_size()
    blocks = [set_block(len(sets), r, world) for r in range(world)]
    s0, s1 = blocks[rank]
    ndays = natural.shape[1]
    if s1 > s0:
        part = engine.reservoirs(natural, sets[s0:s1], start_year, start_month, legacy=legacy)["flow"]
    else:
        part = np.zeros((0, ndays), dtype=np.float32)
    return _all_gather_rows(dist, part, [b - a for a, b in blocks], device)
