/* vicnl_gpu.c -- `vicNl -g <global parameter file>` as a plain C host program over the C ABI of libvicres.so.
 *
 * What the reference's main() (runoff/model/vicNl.c:11-380) does cell by cell -- read_soilparam / read_vegparam /
 * read_snowband, initialize_atmos, initialize_model_state (or INIT_STATE), the record loop of full_energy + put_data,
 * write_data, write_model_state -- this program does for all cells of the soil file with the batched calls of
 * include/vicres.h, vicres_files.h and vicres_atmos.h:
 *
 *   vr_global_read -> vr_params_read -> vr_read_forcing_file (per cell) -> vr_atmos_prepare
 *   -> vr_create / vr_set_veglib / vr_set_cells / vr_init_state | vr_read_state_file
 *   -> vr_step_block (in blocks of records; the state stays on the GPU) [-> vr_write_state_file]
 *   -> [vr_aggregate when OUT_STEP > TIME_STEP] -> vr_write_results_sel (the default "fluxes_<lat>_<lng>" / "snow_<lat>_<lng>"
 *      or the files of the N_OUTFILES / OUTFILE / OUTVAR lines; ASCII or BINARY_OUTPUT)
 *
 * Build:  gcc -O2 -Iinclude -o host/vicnl_gpu host/vicnl_gpu.c -Lvicres_b200 -lvicres -Wl,-rpath,$PWD/vicres_b200
 * No numerics here: every number comes from the library, which has no CPU path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vicres.h"
#include "vicres_files.h"
#include "vicres_atmos.h"

static void die(const char *what, const char *why)
{
  fprintf(stderr, "vicnl_gpu: %s: %s\n", what, why ? why : "");
  exit(1);
}

int main(int argc, char **argv)
{
  const char *gfile = NULL;
  int device = 0, block = 365, i;
  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-g") && i + 1 < argc) gfile = argv[++i];
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--block") && i + 1 < argc) block = atoi(argv[++i]);
  }
  if (!gfile) die("usage", "vicnl_gpu -g <global parameter file> [--device N] [--block records]");

  static vr_global g;
  if (vr_global_read(gfile, &g) != VR_OK) die("global file", vr_params_last_error());
  if (vr_global_unsupported(&g)) die("option outside the water-balance path of this library", vr_global_unsupported(&g));

  /* parameter files */
  vr_param_files pf;
  memset(&pf, 0, sizeof pf);
  pf.nlayer = g.nlayer; pf.nbands = g.snow_band; pf.root_zones = g.root_zones; pf.vegparam_lai = g.vegparam_lai;
  pf.lai_from_vegparam = g.lai_from_vegparam; pf.soil = g.soil; pf.veglib = g.veglib; pf.vegparam = g.vegparam;
  pf.snowband = g.snowband[0] ? g.snowband : NULL;
  pf.baseflow_nijssen = g.baseflow_nijssen; pf.organic_fract = g.organic_fract;
  if (g.output_force) {      /* vicNl.c:226-234: read_snowband() is not called, so the cell keeps the soil file's elevation */
    pf.snowband = NULL;      /* (read_snowband.c replaces it by the bands' mean elevation)                                   */
    pf.nbands = 1;
  }
  vr_params *par = NULL;
  if (vr_params_read(&pf, &par) != VR_OK) die("parameter files", vr_params_last_error());
  const vr_cells *cells = vr_params_cells(par);
  const int64_t C = cells->ncell;
  const float *lat = vr_params_lat(par), *lng = vr_params_lng(par);
  const int32_t *gridcel = vr_params_gridcel(par);

  /* calendar */
  const int nrecs = g.nrecs;
  int32_t *year = malloc(sizeof(int32_t) * nrecs), *month = malloc(sizeof(int32_t) * nrecs), *day = malloc(sizeof(int32_t) * nrecs),
          *hour = malloc(sizeof(int32_t) * nrecs), *doy = malloc(sizeof(int32_t) * nrecs);
  if (vr_make_dmy(&g, year, month, day, hour, doy) != VR_OK) die("calendar", vr_params_last_error());

  /* forcing files -> raw columns [type][record][cell]; FORCING2's columns follow FORCING1's, and a type in both is taken from
   * the second file (the later entry wins below), as param_set.TYPE[].SUPPLIED ends up pointing at it */
  const int64_t nrf = (int64_t)nrecs * g.time_step / g.force_dt;
  static vr_global g2;
  const int nt1 = g.n_force_types;
  int nt2 = 0;
  if (g.forcing2) {
    if (vr_global_read_forcing2(gfile, &g2) != VR_OK) die("global file (FORCING2)", vr_params_last_error());
    if (g2.force_dt != g.force_dt) die("FORCING2", "a FORCE_DT other than FORCING1's is outside this library's forcing layouts");
    nt2 = g2.n_force_types;
  }
  const int nt = nt1 + nt2;
  double *one = malloc(sizeof(double) * (nt1 > nt2 ? nt1 : nt2) * nrf), *raw = calloc((size_t)nt * nrf * C, sizeof(double));
  int64_t c, r;
  int t, col_prec = -1, col_ta = -1, col_tb = -1, col_wind = -1, col_sw = -1, col_lw = -1, col_vp = -1, col_pr = -1, col_de = -1, subdaily = 0;
  int col_q = -1, col_rh = -1, col_rf = -1, col_sf = -1, col_we = -1, col_wn = -1;
  for (t = 0; t < nt; t++) {
    const char *ft = t < nt1 ? g.force_type[t] : g2.force_type[t - nt1];
    if (!strcmp(ft, "PREC")) col_prec = t;
    else if (!strcmp(ft, "TMAX")) col_ta = t;
    else if (!strcmp(ft, "TMIN")) col_tb = t;
    else if (!strcmp(ft, "AIR_TEMP")) { col_ta = t; subdaily = 1; }
    else if (!strcmp(ft, "WIND")) col_wind = t;
    else if (!strcmp(ft, "SHORTWAVE")) col_sw = t;
    else if (!strcmp(ft, "LONGWAVE")) col_lw = t;
    else if (!strcmp(ft, "VP")) col_vp = t;
    else if (!strcmp(ft, "PRESSURE")) col_pr = t;
    else if (!strcmp(ft, "DENSITY")) col_de = t;
    else if (!strcmp(ft, "QAIR")) col_q = t;
    else if (!strcmp(ft, "REL_HUMID")) col_rh = t;
    else if (!strcmp(ft, "RAINF")) col_rf = t;
    else if (!strcmp(ft, "SNOWF")) col_sf = t;
    else if (!strcmp(ft, "WIND_E")) col_we = t;
    else if (!strcmp(ft, "WIND_N")) col_wn = t;
    else if (strcmp(ft, "SKIP")) die("forcing column outside the supported layouts", ft);
  }
  if ((col_prec < 0 && (col_rf < 0 || col_sf < 0)) || col_ta < 0 || (!subdaily && col_tb < 0))
    die("forcing file", "needs PREC (or RAINF and SNOWF) and TMAX+TMIN or AIR_TEMP");
  for (c = 0; c < C; c++) {
    int file;
    for (file = 0; file < (g.forcing2 ? 2 : 1); file++) {
      const vr_global *gf = file ? &g2 : &g;
      const int ntf = file ? nt2 : nt1, t0 = file ? nt1 : 0;
      char path[1200];
      snprintf(path, sizeof path, "%s%.*f_%.*f", gf->forcing1, g.grid_decimal, lat[c], g.grid_decimal, lng[c]);
      int64_t got = vr_read_forcing_file(gf, path, nrf, one);
      if (got < 0) die(path, vr_params_last_error());
      if (got < nrf) die(path, "Not enough records in the forcing file");      /* read_atmos_data.c:178-182 */
      for (t = 0; t < ntf; t++)
        for (r = 0; r < got; r++) raw[((size_t)(t0 + t) * nrf + r) * C + c] = one[(size_t)t * nrf + r];
    }
  }

  /* initialize_atmos for all cells */
  vr_atmos_options ao;
  vr_atmos_default_options(&ao);
  ao.time_step = g.time_step; ao.nrecs = nrecs; ao.startyear = g.startyear; ao.startmonth = g.startmonth;
  ao.startday = g.startday; ao.starthour = g.starthour; ao.force_dt = g.force_dt;
  ao.layout = subdaily ? VR_ATMOS_SUBDAILY_AIR_TEMP : VR_ATMOS_DAILY_TMAX_TMIN;
  ao.has_wind = col_wind >= 0 || (col_we >= 0 && col_wn >= 0); ao.min_wind_speed = g.min_wind_speed; ao.vp_interp = g.vp_interp; ao.vp_iter = g.vp_iter;
  ao.mtclim_swe_corr = g.mtclim_swe_corr; ao.lw_type = g.lw_type; ao.lw_cloud = g.lw_cloud; ao.plapse = g.plapse;
  ao.sw_prec_thresh = g.sw_prec_thresh; ao.alma_input = g.alma_input;
  ao.snow_step = g.snow_step != g.time_step ? g.snow_step : 0; ao.max_snow_temp = g.max_snow_temp;
  /* SNOW_STEP < TIME_STEP: per record the NF windows' values, then the record's own (atmos-><field>[0..NF-1], [NR]) */
  const int NF = g.time_step / g.snow_step, NS = NF > 1 ? NF + 1 : 1;
  vr_atmos_cells site;
  memset(&site, 0, sizeof site);
  site.ncell = C; site.lat = vr_params_site(par, 0); site.lng = vr_params_site(par, 1);
  site.time_zone_lng = vr_params_site(par, 2); site.elevation = vr_params_site(par, 3); site.annual_prec = vr_params_site(par, 4);
  site.min_Tfactor = vr_params_site(par, 5);
  vr_atmos_raw rw;
  memset(&rw, 0, sizeof rw);
#define COL(k) ((k) >= 0 ? raw + (size_t)(k) * nrf * C : NULL)
  rw.qair = COL(col_q); rw.rel_humid = COL(col_rh); rw.rainf = COL(col_rf); rw.snowf = COL(col_sf); rw.wind_e = COL(col_we);
  rw.wind_n = COL(col_wn);
  rw.nrec_forcing = nrf; rw.prec = COL(col_prec); rw.t_a = raw + (size_t)col_ta * nrf * C;
  rw.t_b = col_tb >= 0 ? raw + (size_t)col_tb * nrf * C : NULL; rw.wind = col_wind >= 0 ? raw + (size_t)col_wind * nrf * C : NULL;
  rw.shortwave = col_sw >= 0 ? raw + (size_t)col_sw * nrf * C : NULL; rw.longwave = col_lw >= 0 ? raw + (size_t)col_lw * nrf * C : NULL;
  rw.vp = col_vp >= 0 ? raw + (size_t)col_vp * nrf * C : NULL;
  rw.pressure = col_pr >= 0 ? raw + (size_t)col_pr * nrf * C : NULL; rw.density = col_de >= 0 ? raw + (size_t)col_de * nrf * C : NULL;
  double *field[VR_NFORCING];
  for (i = 0; i < VR_NFORCING; i++) field[i] = malloc(sizeof(double) * (size_t)nrecs * NS * C);
  int32_t *snowflag = malloc(sizeof(int32_t) * (size_t)nrecs * C);
  if (vr_atmos_prepare_sub(device, &ao, &site, &rw, field, snowflag) != VR_OK) die("forcing preparation", vr_atmos_last_error());

  if (g.output_force) {      /* OUTPUT_FORCE TRUE: write the forcing and stop (vicNl.c:226-234, write_forcing_file.c) */
    static vr_outsel fs;
    int32_t fvars[VR_MAX_OUT], nfv = 0, fi, k;
    if (vr_outsel_read(gfile, g.nlayer, &fs) != VR_OK || vr_outsel_step_vars(&fs, 0, fvars, &nfv) != VR_OK)
      die("output selection", vr_params_last_error());
    const int nrows = nrecs * NF;
    double *rows_h = malloc(sizeof(double) * (size_t)nfv * nrows * C);
    if (vr_forcing_outputs(device, nfv, fvars, nrecs, NF, C, g.max_snow_temp, g.min_rain_temp, (const double *const *)field, rows_h) != VR_OK)
      die("forcing outputs", "a variable of the output selection is not a forcing variable");
    if (g.alma_output && vr_alma_output(device, nfv, fvars, nrows, C, g.time_step, rows_h) != VR_OK) die("vr_alma_output", "");
    if (g.prt_header) {
      fs.header = 1; fs.hdr_nrecs = nrecs; fs.hdr_out_dt = g.out_step > 0 ? g.out_step : g.time_step; fs.hdr_alma = g.alma_output;
      fs.hdr_start[0] = year[0]; fs.hdr_start[1] = month[0]; fs.hdr_start[2] = day[0]; fs.hdr_start[3] = hour[0];
    }
    int32_t *zero = calloc(nrows, sizeof(int32_t));
    for (fi = 0; fi < fs.nfiles; fi++) {
      int32_t rws[VR_MAX_OUTCOLS];
      for (k = 0; k < fs.ncols[fi]; k++) {
        rws[k] = 0;
        for (i = 0; i < nfv; i++) if (fvars[i] == fs.var[fi][k]) rws[k] = i;
      }
      if (vr_write_results_sel(g.result_dir, &fs, fi, g.grid_decimal, C, lat, lng, nrows, zero, zero, zero, NULL, g.binary_output, rws,
                               rows_h) != VR_OK) die(fs.prefix[fi], vr_params_last_error());
    }
    fprintf(stderr, "vicnl_gpu: forcing of %lld cells x %d rows written to %s\n", (long long)C, nrows, g.result_dir);
    vr_params_free(par);
    return 0;
  }

  /* the model */
  vr_options opt;
  vr_default_options(&opt, g.nlayer);
  /* the output files and what the step has to compute for them (parse_output_info.c; put_data.c:686 inverts the aggregated
   * conductance, so an aggregating run computes OUT_AERO_COND where a file asks for OUT_AERO_RESIST) */
  static vr_outsel sel;
  if (vr_outsel_read(gfile, g.nlayer, &sel) != VR_OK) die("output selection", vr_params_last_error());
  const int ratio = g.out_step > 0 ? g.out_step / g.time_step : 1;
  if (vr_outsel_step_vars(&sel, (ratio > 1 ? 1 : 0) | (g.alma_output ? 2 : 0), opt.out_vars, &opt.n_out) != VR_OK)
    die("output selection", vr_params_last_error());
  opt.nbands = g.snow_band; opt.dt_hours = g.time_step; opt.snow_step_hours = g.snow_step;
  opt.max_snow_temp = g.max_snow_temp; opt.min_rain_temp = g.min_rain_temp; opt.wind_h = g.wind_h;
  opt.moistfract = g.moistfract;
  vr_handle *h = NULL;
  if (vr_create(&opt, device, &h) != VR_OK) die("vr_create", vr_last_error(NULL));
  if (vr_set_veglib(h, vr_params_veglib(par)) != VR_OK || vr_set_cells(h, cells) != VR_OK) die("parameters", vr_last_error(h));
  const double *tair0 = field[VR_F_AIR_TEMP] + (size_t)(NS - 1) * C;      /* atmos[0].air_temp[NR] of every cell */
  if (g.init_state) {
    if (vr_read_state_file(h, g.init_state_file, g.binary_state_file, tair0, gridcel) != VR_OK) die("INIT_STATE", vr_last_error(h));
  }
  else if (vr_init_state(h, tair0) != VR_OK) die("vr_init_state", vr_last_error(h));

  /* record loop in blocks; a block ends at the day whose state is to be saved (vicNl.c:311-318) */
  const int n_out = opt.n_out;
  double *out = malloc(sizeof(double) * (size_t)n_out * nrecs * C), *blk = malloc(sizeof(double) * (size_t)n_out * block * C);
  int32_t *status = calloc(C, sizeof(int32_t));
  int save_after = -1, r0;
  if (g.statename[0])
    for (i = 0; i < nrecs; i++)
      if (year[i] == g.stateyear && month[i] == g.statemonth && day[i] == g.stateday) save_after = i;
  for (r0 = 0; r0 < nrecs;) {
    int nr = nrecs - r0 < block ? nrecs - r0 : block, v;
    if (save_after >= r0 && save_after < r0 + nr) nr = save_after - r0 + 1;
    vr_forcing f;
    f.nrec = nr; f.month = month + r0; f.day_in_year = doy + r0;
    for (i = 0; i < VR_NFORCING; i++) f.field[i] = field[i] + (size_t)r0 * NS * C;
    f.snowflag = NF > 1 ? snowflag + (size_t)r0 * C : NULL;
    int rc = vr_step_block(h, &f, blk, status);
    if (rc != VR_OK && rc != VR_ERR_CELL) die("vr_step_block", vr_last_error(h));
    for (v = 0; v < n_out; v++)
      memcpy(out + ((size_t)v * nrecs + r0) * C, blk + (size_t)v * nr * C, sizeof(double) * (size_t)nr * C);
    if (r0 + nr - 1 == save_after) {
      char path[1200];
      snprintf(path, sizeof path, "%s_%04d%02d%02d", g.statename, g.stateyear, g.statemonth, g.stateday);
      if (vr_write_state_file(h, path, g.stateyear, g.statemonth, g.stateday, g.binary_state_file, gridcel) != VR_OK)
        die("state file", vr_last_error(h));
    }
    r0 += nr;
  }

  /* put_data.c:665-688: one output record per `ratio` model records, dated by the interval's last record */
  int32_t agg_vars[VR_MAX_OUT + 1], n_rows = n_out, want_resist = 0, want_cond = 0, fi, k;
  for (fi = 0; fi < sel.nfiles; fi++)
    for (k = 0; k < sel.ncols[fi]; k++) {
      want_resist |= sel.var[fi][k] == VR_OUT_AERO_RESIST;
      want_cond |= sel.var[fi][k] == VR_OUT_AERO_COND;
    }
  for (i = 0; i < n_out; i++) agg_vars[i] = opt.out_vars[i];
  int nout = nrecs;
  if (ratio > 1) {
    nout = nrecs / ratio;
    int cond_row = -1;
    for (i = 0; i < n_out; i++) if (opt.out_vars[i] == VR_OUT_AERO_COND) cond_row = i;
    if (want_resist && want_cond) n_rows = n_out + 1;      /* both columns come from the one conductance row */
    double *src = malloc(sizeof(double) * (size_t)n_rows * nout * ratio * C), *agg = malloc(sizeof(double) * (size_t)n_rows * nout * C);
    for (i = 0; i < n_rows; i++)
      memcpy(src + (size_t)i * nout * ratio * C, out + (size_t)(i < n_out ? i : cond_row) * nrecs * C, sizeof(double) * (size_t)nout * ratio * C);
    if (want_resist) agg_vars[want_cond ? n_out : cond_row] = VR_OUT_AERO_RESIST;
    if (vr_aggregate(device, n_rows, agg_vars, 1, nout * ratio, ratio, C, src, agg) != VR_OK) die("vr_aggregate", vr_last_error(NULL));
    free(src); free(out);
    out = agg;
    for (i = 0; i < nout; i++) {
      year[i] = year[i * ratio + ratio - 1]; month[i] = month[i * ratio + ratio - 1];
      day[i] = day[i * ratio + ratio - 1]; hour[i] = hour[i * ratio + ratio - 1];
    }
  }
  const int out_dt = g.out_step > 0 ? g.out_step : g.time_step;
  if (g.alma_output && vr_alma_output(device, n_rows, agg_vars, nout, C, out_dt, out) != VR_OK)      /* put_data.c:697-760 */
    die("vr_alma_output", vr_last_error(NULL));
  if (g.prt_header) {      /* write_header.c: the run's record count and the date of the first MODEL record */
    int32_t y0, m0, d0, h0, j0;
    vr_global g1 = g;
    g1.nrecs = 1;
    vr_make_dmy(&g1, &y0, &m0, &d0, &h0, &j0);
    sel.header = 1; sel.hdr_nrecs = nrecs; sel.hdr_out_dt = out_dt; sel.hdr_alma = g.alma_output;
    sel.hdr_start[0] = y0; sel.hdr_start[1] = m0; sel.hdr_start[2] = d0; sel.hdr_start[3] = h0;
  }
  /* SKIPYEAR: an output record is written when its last model record is past the skipped years (put_data.c:765) */
  {
    int32_t *year_all = malloc(sizeof(int32_t) * nrecs), *t1 = malloc(sizeof(int32_t) * nrecs), *t2 = malloc(sizeof(int32_t) * nrecs),
            *t3 = malloc(sizeof(int32_t) * nrecs), *t4 = malloc(sizeof(int32_t) * nrecs);
    vr_make_dmy(&g, year_all, t1, t2, t3, t4);
    const int skip = vr_output_skip(&g, year_all);
    int first = 0;
    while (first < nout && first * ratio + ratio - 1 < skip) first++;
    if (first > 0) {
      for (i = 0; i < n_rows; i++)
        memmove(out + (size_t)i * (nout - first) * C, out + ((size_t)i * nout + first) * C, sizeof(double) * (size_t)(nout - first) * C);
      for (i = first; i < nout; i++) {
        year[i - first] = year[i]; month[i - first] = month[i]; day[i - first] = day[i]; hour[i - first] = hour[i];
      }
      nout -= first;
    }
    free(year_all); free(t1); free(t2); free(t3); free(t4);
  }
  /* write_data.c */
  for (fi = 0; fi < sel.nfiles; fi++) {
    int32_t rows[VR_MAX_OUTCOLS];
    for (k = 0; k < sel.ncols[fi]; k++) {
      rows[k] = -1;
      for (i = 0; i < n_rows; i++) if (agg_vars[i] == sel.var[fi][k]) rows[k] = i;
      if (rows[k] < 0) die("output selection", "column without a computed row");
    }
    if (vr_write_results_sel(g.result_dir, &sel, fi, g.grid_decimal, C, lat, lng, nout, year, month, day, out_dt < 24 ? hour : NULL,
                             g.binary_output, rows, out) != VR_OK) die(sel.prefix[fi], vr_params_last_error());
    if (g.compress) {      /* close_files.c:51: gzip every result file (the forcing files are left alone) */
      for (c = 0; c < C; c++) {
        char cmd[1400];
        snprintf(cmd, sizeof cmd, "gzip -f %s/%s_%.*f_%.*f", g.result_dir, sel.prefix[fi], g.grid_decimal, lat[c], g.grid_decimal, lng[c]);
        if (system(cmd) != 0) die("gzip", cmd);
      }
    }
  }
  fprintf(stderr, "vicnl_gpu: %lld cells x %d output records written to %s\n", (long long)C, nout, g.result_dir);
  vr_destroy(h);
  vr_params_free(par);
  return 0;
}
