#!/usr/bin/env python
"""bench.py -- cell-timesteps/s of the water-balance cell step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref)

Workload (config.workload): BASELINE.json configs[1] -- synthetic 100k-cell grid per GPU, 3 soil
layers, water-balance mode, daily step; one "step" of the benchmark is one simulated YEAR (365
records) of every cell, so the default K=10 timed steps are the 10 years the config names.
  value   whole-job cell-timesteps/s with forcing and outputs resident in HBM (CUDA events)
  e2e     the same through the C-ABI host call vr_step_block(): pinned host forcing in, host outputs
          out, H2D + D2H copies inside the timed region
  roofline.achieved = 280 algorithmic bytes per cell-timestep (9 forcing + 26 output fp64 words,
          SURVEY.md 8d) x cells / mean duration of the step kernels of ONE record (k_front, k_back on the step
          stream beside k_snow, k_back_list on the snow stream -- together "the cell step"; CUDA events recorded
          by the library around every chunk of records on the launching stream)
  fp64    the same launches against the FP64 pipe: FP64 thread-instructions per cell-timestep (from
          the committed ncu capture, profiles/k_step_traffic.json) x rate / DFMA peak measured live
  cpu_baseline  the compiled reference (vic_ref_driver, step loop only) on a bounded sample on all
          host cores (kind "reference"), else the oracle port on threads (kind "port")
Multi-GPU: cells shard across ranks (no data-path collective) -> weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-timesteps/sec (water-balance, 3 layers)"
UNIT = "cell-timesteps/s"
B_ALG = 8 * (9 + 26)          # bytes per cell-timestep, SURVEY.md section 8(d)
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "vic_ref_driver")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# CPU arms (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def _make_ref_cases(tmp, cores, cells_per_core, ndays):
    from vicres_b200 import synth
    cases = []
    for k in range(cores):
        g = synth.make_case(os.path.join(tmp, "p%d" % k), ncells=cells_per_core, nlayer=3, nbands=1, ndays=ndays,
                            seed=20261019 + k, skip_output=True)
        cases.append(g)
    return cases


def _run_ref_once(cases):
    """one process per core, each over its own cells (cells are independent: vicNl.c:197-356);
    returns (cell_steps, max step-loop seconds, wall seconds)"""
    t0 = time.perf_counter()
    procs = [subprocess.Popen([REF_DRIVER, "-g", g, "-t"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for g in cases]
    steps, loop, prep = 0, 0.0, 0.0
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference driver failed")
        j = json.loads(out.strip().splitlines()[-1])
        steps += j["cell_steps"]
        loop = max(loop, j["step_loop_s"])
        prep = max(prep, j.get("forcing_prep_s", 0.0))
    _run_ref_once.forcing_prep_s = prep          # the reference's own initialize_atmos over the same cells
    return steps, loop, time.perf_counter() - t0


def _port_throughput(cores, cells_per_core, nrec):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_lib
    from vicres_b200 import abi, synth_soa
    C = cores * cells_per_core
    lib, cells = synth_soa.make_cells(C, nbands=1, seed=20261019)
    f, month, doy = synth_soa.make_forcing(cells, nrec, seed=7)
    opt = abi.default_options(3, 1, 24)
    o = oracle_lib.Oracle(opt, lib, cells)
    o.init_state(f[0, 0])
    out = np.zeros((opt.n_out, nrec, C))
    th = [threading.Thread(target=o.step_range, args=(month, doy, f, out, k * cells_per_core, (k + 1) * cells_per_core))
          for k in range(cores)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    o.close()
    return C * nrec / dt


def cpu_baseline(cells_per_core=24, ndays=3653):
    cores = host_cores()
    if os.path.exists(REF_DRIVER):
        with tempfile.TemporaryDirectory() as tmp:
            cases = _make_ref_cases(tmp, cores, cells_per_core, ndays)
            steps, loop, wall = _run_ref_once(cases)
        prep = getattr(_run_ref_once, "forcing_prep_s", 0.0)
        return {"value": steps / loop, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": "%d cells x %d daily records (%d per core, one vic_ref_driver process per core, "
                          "full_energy+put_data loop only; whole run incl. forcing disaggregation %.1f s)"
                          % (cores * cells_per_core, ndays, cells_per_core, wall),
                # the reference's initialize_atmos (forcing files -> atmos[]) over the same sample, all cores
                "forcing_prep_cell_days_per_s": (steps / prep) if prep > 0 else None}
    v = _port_throughput(cores, 64, 730)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d cells x 730 daily records, oracle port on %d threads" % (cores * 64, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    cells_per_core, ndays = 8, 3653
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "gpu_launches": 0}
    if os.path.exists(REF_DRIVER):
        with tempfile.TemporaryDirectory() as tmp:
            cases = _make_ref_cases(tmp, cores, cells_per_core, ndays)
            for _ in range(args.warmup):
                _run_ref_once(cases)
            steps = loop = wall = 0.0
            for _ in range(args.steps):
                s, l, w = _run_ref_once(cases)
                steps += s; loop += l; wall += w
        value = steps / loop
        kind = "reference"
        sample = ("each step = %d cells x %d daily records, one vic_ref_driver process per core (%d), "
                  "full_energy+put_data loop timed (max over processes); whole processes incl. forcing "
                  "disaggregation and parsing: %.3e cell-timesteps/s" % (cores * cells_per_core, ndays, cores, steps / wall))
        ms = 1e3 * loop / args.steps
    else:
        value = _port_throughput(cores, 64, 730)
        kind, sample, ms = "port", "oracle port, %d threads, %d cells x 730 records" % (cores, cores * 64), None
    line.update({"value": value, "ms_per_step": ms,
                 "config": {"workload": "synthetic 100k-cell-grid parameter/forcing distribution, 3 soil layers, "
                                        "water-balance mode, daily step, 10 years; bounded sample on host cores"},
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.p, self.rows = index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", os.environ.get("VICRES_BENCH_SAMPLER_MS", "200")],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, power, reasons = [], [], [], set()
        for ln in out.splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(s for s, p in zip(sm, power) if p >= 0.5 * max(power)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from vicres_b200 import abi, build, model, shard, synth_soa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    C, R, K, W = args.cells, args.block, args.steps, max(args.warmup, 0)
    if args.strong:            # the whole grid is args.cells: this rank takes its block of cells (vicres_b200/shard.py)
        c0, c1 = shard.cell_block(args.cells, rank, world)
        C = c1 - c0
    if args.ensemble and world > 1:     # parameter sets shard over the ranks, every rank holds the basin's cells and forcing
        p0, p1 = shard.cell_block(args.ensemble, rank, world)
        args.ensemble_local = p1 - p0
    else:
        args.ensemble_local = args.ensemble
    lib, cells = synth_soa.make_cells(C, nlayer=3, nbands=args.bands, seed=20261019 + rank,
                                      **({"lat_range": (35.0, 65.0), "overstory_frac": 0.3} if args.bands > 1 else {}))
    f_host, month, doy = synth_soa.make_forcing(cells, R, dt_hours=args.dt, seed=7 + rank, cold=args.cold)
    opt = abi.default_options(3, args.bands, args.dt, out=["RUNOFF", "BASEFLOW"] if args.ensemble else None)
    if args.chunk > 0:
        opt.cells_per_block_hint = args.chunk
    n_out = opt.n_out
    fmap = None
    if args.ensemble:       # P copies of the basin's cells with perturbed calibration parameters, one forcing column per basin cell
        P, N = args.ensemble_local, C
        rng = np.random.default_rng(99 + rank)
        idx = np.tile(np.arange(N), P)
        ens = shard.slice_cells(cells, 0, N)            # (copy)
        tp = np.asarray(cells["tile_ptr"]); nt = tp[1:] - tp[:-1]
        tiles = np.concatenate([np.arange(tp[c], tp[c + 1]) for c in range(N)])
        for k in abi.CELL_SCALARS:
            ens[k] = np.tile(cells[k], P)
        for k in abi.CELL_LAYER + abi.CELL_BAND:
            ens[k] = np.tile(cells[k], (1, P))
        for k in ["zwt_zwt", "zwt_moist"]:
            ens[k] = np.tile(cells[k], (1, 1, P))
        ens["tile_ptr"] = np.concatenate([[0], np.cumsum(np.tile(nt, P))]).astype(np.int64)
        for k in ["tile_class", "tile_Cv", "tile_root", "tile_LAI", "tile_albedo", "tile_vegcover"]:
            ens[k] = np.concatenate([np.asarray(cells[k])[tiles]] * P)
        ens["b_infilt"] = rng.uniform(0.001, 0.9, P * N); ens["Ds"] = rng.uniform(0.001, 0.9, P * N)
        ens["Dsmax"] = rng.uniform(1.0, 30.0, P * N); ens["Ws"] = rng.uniform(0.3, 0.9, P * N)
        ens["c_expt"] = rng.uniform(1.0, 3.0, P * N)
        cells, fmap, C = ens, idx, P * N
    m = model.VicRes(opt, lib, cells, device=local, forcing_map=fmap, n_forcing_cols=f_host.shape[2] if fmap is not None else None)
    m.init_state(f_host[0, 0])
    dev = torch.device("cuda", local)

    # ---- value: inputs/outputs resident in HBM ----
    f_dev = torch.from_numpy(f_host).to(dev)
    out_dev = torch.empty((n_out, R, C), dtype=torch.float64, device=dev)
    ptrs = [f_dev[i].data_ptr() for i in range(abi.VR_NFORCING)]
    stream = torch.cuda.Stream(dev)      # a real (non-default) stream: the kernel and the timing events share it

    def one_step():
        m.step_device(month, doy, ptrs, out_dev.data_ptr(), stream=stream.cuda_stream)

    torch.cuda.synchronize(dev)
    for _ in range(W):
        one_step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = m.num_launches
    m.kernel_times()                     # reset the library's per-launch event sums
    m.step_diag()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    torch.cuda.synchronize(dev)
    ev[0].record(stream)
    for k in range(K):
        one_step()
        ev[k + 1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    total_ms = ev[0].elapsed_time(ev[K])
    launches = m.num_launches - l0
    units_ms, cells_ms, pairs, krecs = m.kernel_times()      # every chunk of step kernels / k_cells launch of the timed region
    diag = m.step_diag()
    kernel_ms = units_ms / max(krecs, 1)                     # the step kernels of one record (both streams)
    recs_per_launch = 1.0
    finite = bool(torch.isfinite(out_dev).all().item())
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    cells_job = torch.tensor([float(C)], dtype=torch.float64, device=dev)     # cells (x sets) of all ranks
    if world > 1:
        dist.all_reduce(cells_job, op=dist.ReduceOp.SUM)
    cells_job = float(cells_job.item())
    value = cells_job * R * K / (total_ms_max * 1e-3)

    # ---- e2e: host buffers through vr_step_block (H2D + D2H inside the timed region) ----
    f_pin = torch.from_numpy(f_host).pin_memory()
    out_pin = torch.empty((n_out, R, C), dtype=torch.float64).pin_memory()
    f_np, out_np = f_pin.numpy(), out_pin.numpy()
    Ke = max(1, min(K, args.e2e_steps))
    for _ in range(min(W, 2)):
        m.step(month, doy, f_np, out=out_np)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    l1 = m.num_launches
    m.kernel_times()
    t0 = time.perf_counter()
    for _ in range(Ke):
        m.step(month, doy, f_np, out=out_np)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_units_ms, e2e_cells_ms, _, _ = m.kernel_times()      # the same kernels' durations with the copies running beside them
    launches += m.num_launches - l1
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = cells_job * R * Ke / float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    finite = finite and bool(np.isfinite(out_np).all())
    # the host-side ceiling of the end-to-end number: the same bytes (forcing in, outputs out) copied with no kernel at
    # all, H2D and D2H on two streams, all ranks at once (they share the host's memory system and PCIe root complexes)
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        with torch.cuda.stream(s_h2d):
            f_dev.copy_(f_pin, non_blocking=True)
        with torch.cuda.stream(s_d2h):
            out_pin.copy_(out_dev, non_blocking=True)
    torch.cuda.synchronize(dev)
    copy_s = time.perf_counter() - t0
    t = torch.tensor([copy_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    copy_ms_per_step = 1e3 * float(t.item()) / Ke

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, which = 6650.0, "fallback (B200_PROFILING.md)"
        b_alg = 8 * n_out if args.ensemble else B_ALG       # ensemble: the sets share the forcing columns
        achieved = b_alg * C * recs_per_launch / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k_step_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_cell_step", 0) * C * recs_per_launch or None
            fp64_per = tj.get("fp64_thread_inst_per_cell_step")
        else:
            fp64_per = None
        try:
            dfma_peak = model.diag_fp64_peak(local)
        except Exception:
            dfma_peak = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "strong" if (args.strong or args.ensemble) else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("synthetic 100k-cell grid, 3 soil layers, water-balance mode, daily step, "
                                    "10 years (one step = one simulated year of every cell)") if
                                   (C == 100000 and R == 365 and args.bands == 1 and args.dt == 24) else
                                   (("ensemble of %d parameter sets (%d on this rank) x %d basin cells sharing %d forcing columns, outputs RUNOFF+BASEFLOW, "
                                     % (args.ensemble, args.ensemble_local, C // max(args.ensemble_local, 1), C // max(args.ensemble_local, 1)) if args.ensemble else "") +
                                    (("BASELINE configuration %d: " % args.config) if args.config else "") +
                                    (("%d-cell grid sharded by cell blocks over %d rank(s), this rank " % (args.cells, world)) if args.strong else "") +
                                    "synthetic %d-cell grid, 3 soil layers, %d snow band(s), %d-hourly step, "
                                    "%d records per step" % (C, args.bands, args.dt, R)),
                       "cells_per_gpu": C, "records_per_step": R, "active_tile_band_units_per_gpu": m.num_units,
                       "parallelism": "cells sharded over %d GPU(s), no data-path collective" % world,
                       "l2": "inputs larger than L2: %.2f GB forcing + %.2f GB outputs per step" %
                             (f_host.nbytes / 1e9, out_np.nbytes / 1e9),
                       "outputs_finite": finite},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(f_host.nbytes),
                    "d2h_bytes_per_step": int(out_np.nbytes), "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
                    "copies_only_ms_per_step": copy_ms_per_step,      # no kernels: what the host and PCIe allow, all ranks at once
                    "copies_only_GBps_per_rank": (f_host.nbytes + out_np.nbytes) / 1e9 / (copy_ms_per_step * 1e-3),
                    "step_kernels_ms_per_step": e2e_units_ms / Ke, "k_cells_ms_per_step": e2e_cells_ms / Ke,
                    "step_kernels_ms_per_step_resident": units_ms / K, "k_cells_ms_per_step_resident": cells_ms / K},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": which,
                         "kernel": "step kernels of one record: k_front + k_back (step stream) || k_snow + k_back_list (snow stream)",
                         "kernel_ms": kernel_ms, "launches_timed": krecs, "records_per_launch": recs_per_launch,
                         "traffic_source": "profiles/k_step_traffic.json (ncu --set full at 100 000 cells) x cells / 100 000",
                         "kernel_share_of_gpu_time": units_ms / max(units_ms + cells_ms, 1e-9),   # compare with the ncu launch list
                         "kernel_share_of_step": units_ms / max(sum(step_ms), 1e-9),            # k_cells overlaps the next k_units
                         "k_cells_share_of_step": cells_ms / max(sum(step_ms), 1e-9),
                         "algorithmic_bytes_per_cell_step": b_alg,
                         "note": "the step is FP64 / dependent-issue bound, not HBM-bound (SURVEY.md 8d): see the "
                                 "fp64 object and DESIGN.md section 6"},
            "step_kernels": {"snow_unit_record_share": diag["snow_unit_records"] / max(diag["unit_records"], 1),
                             "us_per_record": None if not diag["records_timed"] else
                             {k: 1e3 * diag[k + "_ms"] / diag["records_timed"] for k in ("front", "snow", "back", "back_list")}},
            "fp64": None if not (fp64_per and dfma_peak) else {
                "fp64_thread_inst_per_cell_step": fp64_per,
                "achieved_inst_per_s": fp64_per * C * recs_per_launch / (kernel_ms * 1e-3),
                "peak_dfma_per_s": dfma_peak, "peak_source": "measured live (vr_diag_fp64_peak: independent DFMA chains, all SMs)",
                "frac": fp64_per * C * recs_per_launch / (kernel_ms * 1e-3) / dfma_peak},
        }
        if args.forcing_prep and world == 1 and not args.ensemble:
            # raw forcing columns -> the nine atmos fields, for the whole 10-year period of this grid
            # (outside the timed step; reported beside it)
            try:
                del f_dev, out_dev
                torch.cuda.empty_cache()
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_atmos
                line["forcing_prep"] = bench_atmos.measure(cells=args.cells, days=3653 if args.dt == 24 else 365, dt=args.dt,
                                                           reps=2, cpu_cells=8 if args.cpu_baseline else 0, device=local)
            except Exception as e:
                line["forcing_prep"] = {"error": str(e)}
        if args.cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline()
            except Exception as e:      # the baseline is a reported number; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "error", "sample": str(e)}
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--block", type=int, default=365, help="records per step (one simulated year)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-forcing-prep", dest="forcing_prep", action="store_false",
                    help="skip the forcing-preparation measurement (vr_atmos_prepare_device) appended to the line")
    ap.add_argument("--chunk", type=int, default=0, help="records per kernel-pair launch (0 = library default)")
    ap.add_argument("--bands", type=int, default=1, help="snow bands (BASELINE config 3 uses 5)")
    ap.add_argument("--dt", type=int, default=24, help="model step in hours (config 3: 3)")
    ap.add_argument("--cold", type=float, default=0.0, help="lower all temperatures (deg C) to force snow")
    ap.add_argument("--ensemble", type=int, default=0,
                    help="BASELINE config 4: P parameter sets x --cells basin cells sharing the basin's forcing columns "
                         "(outputs RUNOFF and BASEFLOW only)")
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5],
                    help="BASELINE.json configuration preset: 2 = the default (100k cells, daily, 365 records per step, weak scaling); "
                         "3 = 1M cells in total, 5 snow bands, 3-hourly, cold, 32 records per step, STRONG scaling over the ranks; "
                         "4 = 4096 parameter sets x 64 basin cells, sets sharded over the ranks; 5 = the chain forcing preparation -> step -> "
                         "hand-off -> routing with 200 reservoirs x 8 rule sets on 1M cells (tools/bench_pipeline.py), one GPU")
    ap.add_argument("--strong", action="store_true", help="--cells is the whole grid, sharded by cell blocks over the ranks")
    args = ap.parse_args()
    if args.config == 3:
        args.cells, args.bands, args.dt, args.cold, args.block, args.strong = 1000000, 5, 3, 5.0, 32, True
        args.forcing_prep = False
    elif args.config == 4:
        args.cells, args.ensemble, args.block = 64, 4096, 355
        args.forcing_prep = False
    if args.config == 5 and args.impl != "reference":
        return run_chain(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


def run_chain(args):
    """BASELINE configuration 5: the whole chain on ONE GPU (the routing network does not shard: replicas only; under
    torchrun the other ranks leave).  A step is one pass of the chain over a simulated year; value = VIC cell-timesteps per
    second of the chain's wall clock, host columns in and station discharge out (so `e2e` is the same number)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_pipeline
    import torch
    argv = ["--vic-cells", "1000000", "--rows", "200", "--cols", "250", "--reservoirs", "200", "--sets", "8", "--days", "365"]
    K, W = max(1, min(args.steps, 3)), min(args.warmup, 1)
    for _ in range(W):
        bench_pipeline.run(argv)
    sampler = ClockSampler(0)
    sampler.start()
    lines = [bench_pipeline.run(argv) for _ in range(K)]
    clocks = sampler.stop()
    tot = sum(l["stages"]["total_s"] for l in lines)
    C, nd = lines[-1]["cells"], lines[-1]["days"]
    v = C * nd * K / tot
    print(json.dumps({
        "metric": "cell-timesteps/sec (water-balance, 3 layers)", "value": v, "unit": "cell-timesteps/s", "n_gpus": 1, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * tot / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": "BASELINE configuration 5: " + lines[-1]["workload"],
                                        "parallelism": "one GPU (the routed network does not shard: replicas only)",
                                        "timing": "host wall clock around the public calls of the chain, model set-up included",
                                        "l2": "inputs larger than L2"},
        "clocks": clocks,
        "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": int(4 * 8 * C * nd), "d2h_bytes_per_step": 0,
                "note": "the chain starts from host forcing columns and ends with the station series on the host"},
        "gpu_launches": K * (lines[-1]["stages"]["forcing_prep_launches"] + lines[-1]["stages"]["step_launches"] + 2),
        "stages": lines[-1]["stages"], "station_flow_finite": lines[-1]["station_flow_finite"],
        "step_cell_steps_per_s": lines[-1]["stages"]["step_cell_steps_per_s"]}))


if __name__ == "__main__":
    main()
