"""`vicNl -g <global parameter file>` on the GPU: the reference's command line (runoff/model/vicNl.c:11-380,
cmd_proc.c) over the C ABI of libvicres.so.

    python -m vicres_b200.vicnl -g globalparam.txt [--device 0]

What the reference does per cell -- read_soilparam / read_vegparam / read_snowband, initialize_atmos (forcing file ->
atmos[]), initialize_model_state, the record loop of full_energy + put_data, write_data -- is done here for all cells
of the soil file at once:

    vr_global_read -> vr_params_read -> vr_read_forcing_file (per cell) -> vr_atmos_prepare_device
    -> vr_create / vr_set_cells / vr_init_state -> vr_step_block_device -> [vr_aggregate_device] -> vr_write_results_sel

and leaves the same result files in RESULT_DIR: `fluxes_<lat>_<lng>` / `snow_<lat>_<lng>` or the files and columns of the
N_OUTFILES / OUTFILE / OUTVAR lines (parse_output_info.c), aggregated to OUT_STEP (put_data.c:665-688), ASCII or
BINARY_OUTPUT (write_data.c).  Marshalling only: every number
comes from the CUDA library; a global file with options outside the library's path is refused with the reason
(vr_global_unsupported), never run differently.
"""
import argparse
import sys
import numpy as np
from . import abi, atmos, files, model


def run(global_path, device=0, write=True, records_per_block=0):
    """Returns (out [n_out, nrecs, C], meta, dmy); writes the result files when `write`."""
    import torch
    g = files.Global(global_path)
    why = g.unsupported
    if why:
        raise model.VicResError(abi.VR_ERR_UNSUPPORTED, why)
    pf = g.param_files()
    if g.output_force:                      # vicNl.c:226-234: read_snowband() is not called, so the cell keeps the soil file's
        pf["snowband"], pf["nbands"] = None, 1          # elevation (read_snowband.c replaces it by the bands' mean elevation)
    lib, cells, meta = files.read_params(**pf)
    Cn = len(meta["gridcel"])
    ao = atmos.options_from_global(g)
    # forcing files -> [type][record][cell] columns
    cols = None
    g2 = g.second_forcing()                 # FORCING2: its columns join (and, for a type in both, replace) those of FORCING1
    for c in range(Cn):
        d, got = g.read_forcing(float(meta["lat"][c]), float(meta["lng"][c]))
        if g2 is not None:
            d2, got2 = g2.read_forcing(float(meta["lat"][c]), float(meta["lng"][c]))
            d.update(d2)
            got = min(got, got2)
        need = g.nrecs * g.time_step // g.c.force_dt
        if got < need:                   # read_atmos_data.c:178-182, 258-262: "Not enough records in the forcing file"
            raise model.VicResError(abi.VR_ERR_ARG, "Not enough records in the forcing file of cell %d (%d < %d)"
                                    % (int(meta["gridcel"][c]), got, need))
        if cols is None:
            cols = {t: np.zeros((len(v), Cn)) for t, v in d.items() if t != "SKIP"}
        for t in cols:
            cols[t][:got, c] = d[t][:got]
    raw = atmos.raw_from_columns(cols, ao.layout)
    dev = torch.device("cuda", device)
    nrecs = g.nrecs
    site_t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in meta["site"].items()}
    raw_t = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if v is not None else None) for k, v in raw.items()}
    # SNOW_STEP < TIME_STEP (NF > 1): per record the NF windows' values, then the record's own, and the snow flags
    NF = atmos.n_windows(ao)
    NS = NF + 1 if NF > 1 else 1
    f_dev = torch.empty((abi.VR_NFORCING, nrecs * NS, Cn), dtype=torch.float64, device=dev)
    flag_dev = torch.empty((nrecs, Cn), dtype=torch.int32, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else 0
    atmos.prepare_device(ao, {k: ptr(v) for k, v in site_t.items()}, {k: ptr(v) for k, v in raw_t.items()},
                         [f_dev[f].data_ptr() for f in range(abi.VR_NFORCING)], Cn, raw["t_a"].shape[0], device=device,
                         snowflag_dev=flag_dev.data_ptr())
    torch.cuda.synchronize(dev)
    if g.output_force:                      # OUTPUT_FORCE TRUE: write the forcing and stop (vicNl.c:226-234, write_forcing_file.c)
        sel = files.OutSel(global_path, g.nlayer)
        names = sel.step_vars()
        rows = torch.empty((len(names), nrecs * NF, Cn), dtype=torch.float64, device=dev)
        files.forcing_outputs_device(names, [f_dev[f].data_ptr() for f in range(abi.VR_NFORCING)], rows.data_ptr(), nrecs, NF, Cn,
                                     g.max_snow_temp, g.min_rain_temp, device=device)
        if g.alma_output:
            files.alma_output_device(names, rows.data_ptr(), nrecs * NF, Cn, g.time_step, device=device)
        dmy = g.dmy()
        if g.prt_header:
            sel.set_header(nrecs, g.out_step if g.out_step > 0 else g.time_step, dmy[0, :4], alma=bool(g.alma_output))
        out = rows.cpu().numpy()
        if write:
            z = np.zeros(nrecs * NF, dtype=np.int32)
            for f in range(len(sel.files)):
                sel.write(f, g.result_dir, meta["lat"], meta["lng"], z, z, z, None, out, sel.rows(f, names), binary=bool(g.binary_output),
                          grid_decimal=g.grid_decimal)
        meta["out_names"] = names
        return out, meta, dmy
    # what to compute: the columns of the output files, each once (the conductance instead of the resistance when the
    # output interval spans several records: put_data.c:686 inverts the aggregated conductance)
    sel = files.OutSel(global_path, g.nlayer)
    ratio = g.out_step // g.time_step if g.out_step > 0 else 1
    names = sel.step_vars(aggregate=ratio > 1, alma=bool(g.alma_output))
    opt = abi.default_options(g.nlayer, g.snow_band, g.time_step, out=names, max_snow_temp=g.max_snow_temp,
                              min_rain_temp=g.min_rain_temp, wind_h=g.wind_h, snow_step_hours=g.snow_step,
                              moistfract=bool(g.moistfract))
    m = model.VicRes(opt, lib, cells, device=device)
    tair0 = f_dev[abi.FORCING_FIELDS.index("air_temp"), NS - 1].cpu().numpy()       # atmos[0].air_temp[NR]
    if g.init_state:                     # INIT_STATE <file>: read_initial_model_state instead of the cold start
        m.read_state_file(g.init_state_file, tair0, meta["gridcel"], binary=bool(g.binary_state_file))
    else:
        m.init_state(tair0)
    dmy = g.dmy()
    out_dev = torch.empty((opt.n_out, nrecs, Cn), dtype=torch.float64, device=dev)
    # STATENAME + STATEYEAR/MONTH/DAY: the state after the last record of that day (vicNl.c:311-318)
    save_after = -1
    if g.statename:
        hit = np.flatnonzero((dmy[:, 0] == g.stateyear) & (dmy[:, 1] == g.statemonth) & (dmy[:, 2] == g.stateday))
        save_after = int(hit[-1]) if len(hit) else -1
    B = records_per_block or nrecs
    cuts = sorted(set(list(range(0, nrecs, B)) + ([save_after + 1] if 0 <= save_after < nrecs - 1 else []) + [nrecs]))
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        blk = torch.empty((opt.n_out, r1 - r0, Cn), dtype=torch.float64, device=dev)
        fb = f_dev[:, r0 * NS:r1 * NS].contiguous()
        m.step_device(dmy[r0:r1, 1], dmy[r0:r1, 4], [fb[f].data_ptr() for f in range(abi.VR_NFORCING)], blk.data_ptr(),
                      snowflag_ptr=flag_dev[r0:r1].data_ptr() if NF > 1 else None)
        m.synchronize()
        out_dev[:, r0:r1] = blk
        if r1 - 1 == save_after:
            m.write_state_file("%s_%04d%02d%02d" % (g.statename, g.stateyear, g.statemonth, g.stateday), g.stateyear, g.statemonth,
                               g.stateday, meta["gridcel"], binary=bool(g.binary_state_file))
    m.close()
    if ratio > 1:                           # OUT_STEP > TIME_STEP: one output record per `ratio` model records, dated by the last
        nout = nrecs // ratio
        want_resist = sel.wants("AERO_RESIST")
        if want_resist and sel.wants("AERO_COND"):      # both columns from the one conductance row
            k = names.index("AERO_COND")
            out_dev = torch.cat([out_dev, out_dev[k:k + 1]])
            names = names + ["AERO_RESIST"]
        elif want_resist:
            names = ["AERO_RESIST" if n == "AERO_COND" else n for n in names]
        src = out_dev[:, :nout * ratio].contiguous()
        agg = torch.empty((len(names), nout, Cn), dtype=torch.float64, device=dev)
        files.aggregate_device(names, src.data_ptr(), agg.data_ptr(), nout * ratio, ratio, Cn, device=device)
        out_dev = agg
        dmy = dmy[ratio - 1::ratio][:nout]
    out_dt = g.out_step if g.out_step > 0 else g.time_step
    if g.alma_output:                       # put_data.c:697-760, after the aggregation
        out_dev = out_dev.contiguous()
        files.alma_output_device(names, out_dev.data_ptr(), out_dev.shape[1], Cn, out_dt, device=device)
    if g.prt_header:                        # write_header.c: the run's record count and the first MODEL record's date
        d0 = g.dmy()[0]
        sel.set_header(nrecs, out_dt, (d0[0], d0[1], d0[2], d0[3]), alma=bool(g.alma_output))
    # SKIPYEAR: an output record is written when its last model record is past the skipped years (put_data.c:765)
    skip = g.output_skip(g.dmy())
    first = 0
    while first < out_dev.shape[1] and first * ratio + ratio - 1 < skip:
        first += 1
    out = out_dev[:, first:].cpu().numpy()
    dmy = dmy[first:]
    if write:
        hour = dmy[:, 3] if out_dt < 24 else None
        for f in range(len(sel.files)):
            sel.write(f, g.result_dir, meta["lat"], meta["lng"], dmy[:, 0], dmy[:, 1], dmy[:, 2], hour, out, sel.rows(f, names),
                      binary=bool(g.binary_output), grid_decimal=g.grid_decimal)
        if g.compress:                      # close_files.c:51: gzip every result file (the forcing files are left alone)
            import subprocess
            for prefix, _ in sel.files:
                for c in range(Cn):
                    subprocess.run(["gzip", "-f", "%s/%s_%.*f_%.*f" % (g.result_dir, prefix, g.grid_decimal, float(meta["lat"][c]),
                                                                     g.grid_decimal, float(meta["lng"][c]))], check=True)
    meta["out_names"] = names
    return out, meta, dmy


def main(argv=None):
    ap = argparse.ArgumentParser(prog="vicres_b200.vicnl", description="vicNl -g on the GPU (water-balance mode)")
    ap.add_argument("-g", dest="global_file", required=True, help="global parameter file (get_global_param.c format)")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    out, meta, dmy = run(a.global_file, a.device)
    print("vicres_b200.vicnl: %d cells x %d records written" % (out.shape[2], out.shape[1]), file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
