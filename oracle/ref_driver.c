/* ref_driver.c -- TEST INFRASTRUCTURE ONLY (oracle side).  Never linked into the product.
 *
 * A stand-in for the reference's main() (runoff/model/vicNl.c:11-380) that links against the
 * UNMODIFIED reference objects (compiled in place from /root/reference/runoff/model by
 * oracle/Makefile) and drives exactly the same call sequence -- get_global_param,
 * read_soilparam, read_vegparam, initialize_atmos, initialize_model_state, then
 * full_energy() + put_data() per record (vicNl.c:293-306) -- but additionally
 *   (1) dumps, at full fp64 precision, everything the step consumes (derived soil/veg/band
 *       parameters, the disaggregated atmos[] records, the dmy calendar) and everything it
 *       produces (all out_data[] variables per record, optionally per tile x band state),
 *       as a stream of named binary arrays that tests/refdump.py reads; and
 *   (2) times the record loop alone (step-only CPU baseline for bench.py --impl reference).
 *
 * Usage: vic_ref_driver -g <global file> [-d <dump file>] [-u] [-q] [-t]
 *   -d FILE  write the named-array dump to FILE
 *   -u       also dump per tile x band internals after every step (large)
 *   -t       print one JSON line with the step-loop timing to stdout
 *
 * Dump record layout (little endian): char name[48]; char dtype ('d' f64,'i' i32,'b' u8);
 * int32 ndim; int64 dims[ndim]; payload.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <unistd.h>
#include <vicNl.h>
#include <global.h>

static FILE *g_dump = NULL;

static void dump_arr(const char *name, char dtype, int ndim, const int64_t *dims, const void *data)
{
  char nm[48];
  int32_t nd = ndim;
  size_t n = 1, esz;
  int i;
  if (!g_dump) return;
  memset(nm, 0, sizeof nm);
  strncpy(nm, name, sizeof nm - 1);
  fwrite(nm, 1, sizeof nm, g_dump);
  fwrite(&dtype, 1, 1, g_dump);
  fwrite(&nd, sizeof nd, 1, g_dump);
  fwrite(dims, sizeof(int64_t), ndim, g_dump);
  for (i = 0; i < ndim; i++) n *= (size_t)dims[i];
  esz = (dtype == 'd') ? 8 : (dtype == 'i') ? 4 : 1;
  fwrite(data, esz, n, g_dump);
}
static void dump_d1(const char *name, int64_t n, const double *p) { dump_arr(name, 'd', 1, &n, p); }
static void dump_i1(const char *name, int64_t n, const int *p)    { dump_arr(name, 'i', 1, &n, p); }
static void dump_d2(const char *name, int64_t a, int64_t b, const double *p)
{ int64_t d[2]; d[0]=a; d[1]=b; dump_arr(name, 'd', 2, d, p); }
static void dump_ds(const char *name, double v) { dump_d1(name, 1, &v); }
static void dump_is(const char *name, int v) { dump_i1(name, 1, &v); }

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9*ts.tv_nsec;
}

/* per tile x band internals recorded with -u (order fixed; tests/refdump.py names them) */
#define NUNITF 48
static int fill_unit(double *u, cell_data_struct *c, veg_var_struct *v, snow_data_struct *s,
                     energy_bal_struct *e, int HasVeg)
{
  int k = 0, l;
  for (l = 0; l < 3; l++) u[k++] = (l < options.Nlayer) ? c->layer[l].moist : 0.;
  for (l = 0; l < 3; l++) u[k++] = (l < options.Nlayer) ? c->layer[l].evap : 0.;
  for (l = 0; l < 3; l++) u[k++] = (l < options.Nlayer) ? c->layer[l].bare_evap_frac : 0.;
  u[k++] = c->runoff; u[k++] = c->baseflow; u[k++] = c->asat; u[k++] = c->inflow;
  u[k++] = c->aero_resist[0]; u[k++] = c->aero_resist[1];
  u[k++] = HasVeg ? v->Wdew : 0.; u[k++] = HasVeg ? v->canopyevap : 0.;
  u[k++] = HasVeg ? v->throughfall : 0.;
  u[k++] = s->swq; u[k++] = s->surf_temp; u[k++] = s->pack_temp; u[k++] = s->surf_water;
  u[k++] = s->pack_water; u[k++] = s->density; u[k++] = s->depth; u[k++] = s->albedo;
  u[k++] = s->coldcontent; u[k++] = s->coverage; u[k++] = s->snow_canopy;
  u[k++] = s->tmp_int_storage; u[k++] = (double)s->last_snow; u[k++] = (double)s->MELTING;
  u[k++] = (double)s->snow; u[k++] = s->vapor_flux; u[k++] = s->canopy_vapor_flux; u[k++] = s->melt;
  u[k++] = e->T[0]; u[k++] = e->T[1]; u[k++] = e->T[2];
  u[k++] = e->grnd_flux; u[k++] = e->deltaH; u[k++] = e->fusion; u[k++] = e->snow_flux;
  u[k++] = e->LongUnderOut; u[k++] = e->Tfoliage; u[k++] = e->AlbedoUnder;
  u[k++] = e->NetShortAtmos; u[k++] = e->NetLongAtmos; u[k++] = e->Tsurf;
  while (k < NUNITF) u[k++] = 0.;
  return k;
}

int main(int argc, char *argv[])
{
  char   RUN_MODEL, MODEL_DONE;
  char  *gargv[3];
  char   nm[64];
  const char *dumpname = NULL;
  int    want_units = 0, want_time = 0;
  int    rec, i, j, v, b, l, opt, Nveg_type, cellnum = -1, ErrorFlag, nrecs, nout_elem, ncells = 0;
  int    veg, band, nunits;
  long   n_cellsteps = 0;
  double t_step = 0., t_atmos = 0., t0;
  double *outbuf = NULL, *atmbuf = NULL, *unitbuf = NULL, *vhbuf = NULL;
  int   *dmybuf;
  dmy_struct          *dmy;
  atmos_data_struct   *atmos;
  veg_hist_struct    **veg_hist;
  veg_con_struct      *veg_con;
  soil_con_struct      soil_con;
  all_vars_struct      all_vars;
  filenames_struct     filenames;
  filep_struct         filep;
  lake_con_struct      lake_con;
  out_data_file_struct *out_data_files;
  out_data_struct     *out_data;
  save_data_struct     save_data;
  static char gopt[] = "-g";

  gargv[0] = argv[0]; gargv[1] = gopt; gargv[2] = NULL;
  while ((opt = getopt(argc, argv, "g:d:ut")) != -1) {
    if (opt == 'g') gargv[2] = optarg;
    else if (opt == 'd') dumpname = optarg;
    else if (opt == 'u') want_units = 1;
    else if (opt == 't') want_time = 1;
  }
  if (!gargv[2]) { fprintf(stderr, "usage: %s -g global [-d dump] [-u] [-t]\n", argv[0]); return 2; }
  if (dumpname && !(g_dump = fopen(dumpname, "wb"))) { perror(dumpname); return 2; }

  initialize_global();
  optind = 1;
  filenames = cmd_proc(3, gargv);
  filep.globalparam = open_file(filenames.global, "r");
  global_param = get_global_param(&filenames, filep.globalparam);
  out_data = create_output_list();
  out_data_files = set_output_defaults(out_data);
  fclose(filep.globalparam);
  filep.globalparam = open_file(filenames.global, "r");
  parse_output_info(&filenames, filep.globalparam, &out_data_files, out_data);
  check_files(&filep, &filenames);
  veg_lib = read_veglib(filep.veglib, &Nveg_type);
  dmy = make_dmy(&global_param);
  nrecs = global_param.nrecs;
  alloc_atmos(nrecs, &atmos);
  filep.statefile = NULL;

  /* ---- run-wide header ---- */
  dump_is("Nlayer", options.Nlayer); dump_is("Nbands", options.SNOW_BAND);
  dump_is("nrecs", nrecs); dump_is("dt", global_param.dt); dump_is("NF", NF); dump_is("NR", NR);
  dump_is("snow_step", options.SNOW_STEP); dump_is("Nveglib", Nveg_type);
  dump_ds("MAX_SNOW_TEMP", global_param.MAX_SNOW_TEMP);
  dump_ds("MIN_RAIN_TEMP", global_param.MIN_RAIN_TEMP);
  dump_ds("wind_h", global_param.wind_h);
  dmybuf = (int *)malloc(sizeof(int)*5*nrecs);
  for (rec = 0; rec < nrecs; rec++) {
    dmybuf[5*rec+0] = dmy[rec].year; dmybuf[5*rec+1] = dmy[rec].month; dmybuf[5*rec+2] = dmy[rec].day;
    dmybuf[5*rec+3] = dmy[rec].hour; dmybuf[5*rec+4] = dmy[rec].day_in_year;
  }
  { int64_t d[2]; d[0] = nrecs; d[1] = 5; dump_arr("dmy", 'i', 2, d, dmybuf); }
  {
    /* veg library incl. the 4 appended PET reference classes (read_veglib.c:168-187) */
    int ncl = Nveg_type + N_PET_TYPES_NON_NAT, nf = 7 + 6*12;
    double *vl = (double *)calloc((size_t)ncl*nf, sizeof(double));
    for (i = 0; i < ncl; i++) {
      double *p = vl + (size_t)i*nf;
      p[0] = veg_lib[i].overstory; p[1] = veg_lib[i].rarc; p[2] = veg_lib[i].rmin;
      p[3] = veg_lib[i].wind_h; p[4] = veg_lib[i].RGL; p[5] = veg_lib[i].rad_atten;
      p[6] = veg_lib[i].wind_atten;
      for (j = 0; j < 12; j++) {
        p[7+j] = veg_lib[i].LAI[j]; p[19+j] = veg_lib[i].albedo[j]; p[31+j] = veg_lib[i].roughness[j];
        p[43+j] = veg_lib[i].displacement[j]; p[55+j] = veg_lib[i].vegcover[j]; p[67+j] = veg_lib[i].Wdmax[j];
      }
    }
    dump_d2("veglib", ncl, nf, vl);
    { double *tr = (double *)calloc(ncl, sizeof(double)); int *cls = (int *)calloc(ncl, sizeof(int));
      for (i = 0; i < ncl; i++) { tr[i] = veg_lib[i].trunk_ratio; cls[i] = veg_lib[i].veg_class; }
      dump_d1("veglib_trunk_ratio", ncl, tr); dump_i1("veglib_class_id", ncl, cls); free(tr); free(cls); }
    free(vl);
  }
  nout_elem = 0;
  for (v = 0; v < N_OUTVAR_TYPES; v++) nout_elem += out_data[v].nelem;
  {
    int *ne = (int *)malloc(sizeof(int)*N_OUTVAR_TYPES);
    char *names = (char *)calloc((size_t)N_OUTVAR_TYPES*20, 1);
    int64_t d[2];
    for (v = 0; v < N_OUTVAR_TYPES; v++) { ne[v] = out_data[v].nelem; memcpy(names + 20*v, out_data[v].varname, 20); }
    dump_i1("out_nelem", N_OUTVAR_TYPES, ne);
    d[0] = N_OUTVAR_TYPES; d[1] = 20; dump_arr("out_names", 'b', 2, d, names);
    free(ne); free(names);
  }

  MODEL_DONE = FALSE;
  while (!MODEL_DONE) {
    soil_con = read_soilparam(filep.soilparam, &RUN_MODEL, &MODEL_DONE);
    if (!RUN_MODEL) continue;
    cellnum++;
    veg_con = read_vegparam(filep.vegparam, soil_con.gridcel, Nveg_type);
    calc_root_fractions(veg_con, &soil_con);
    make_in_and_outfiles(&filep, &filenames, &soil_con, out_data_files);
    read_snowband(filep.snowband, &soil_con);
    all_vars = make_all_vars(veg_con[0].vegetat_type_num);
    alloc_veg_hist(nrecs, veg_con[0].vegetat_type_num, &veg_hist);

    t0 = now_s();
    initialize_atmos(atmos, dmy, filep.forcing, veg_con, veg_hist, &soil_con, out_data_files, out_data);
    t_atmos += now_s() - t0;

    ErrorFlag = initialize_model_state(&all_vars, dmy[0], &global_param, filep, soil_con.gridcel,
                                       veg_con[0].vegetat_type_num, options.Nnode,
                                       atmos[0].air_temp[NR], &soil_con, veg_con, lake_con);
    if (ErrorFlag == ERROR) { fprintf(stderr, "cell %d: initialize_model_state failed\n", soil_con.gridcel); return 3; }

    if (g_dump) {
      int Nveg = veg_con[0].vegetat_type_num, L = options.Nlayer, NB = options.SNOW_BAND;
      double sc[64]; int k = 0;
      double tmp[16*64];
      sprintf(nm, "c%d/", cellnum);
#define NM(s) (sprintf(nm, "c%d/%s", cellnum, s), nm)
      dump_is(NM("gridcel"), soil_con.gridcel); dump_is(NM("Nveg"), Nveg);
      sc[k++] = soil_con.lat; sc[k++] = soil_con.lng; sc[k++] = soil_con.elevation;
      sc[k++] = soil_con.b_infilt; sc[k++] = soil_con.Ds; sc[k++] = soil_con.Dsmax; sc[k++] = soil_con.Ws;
      sc[k++] = soil_con.c; sc[k++] = soil_con.avg_temp; sc[k++] = soil_con.dp; sc[k++] = soil_con.rough;
      sc[k++] = soil_con.snow_rough; sc[k++] = soil_con.annual_prec; sc[k++] = soil_con.max_infil;
      sc[k++] = soil_con.FS_ACTIVE; sc[k++] = soil_con.frost_slope; sc[k++] = soil_con.cell_area;
      sc[k++] = soil_con.time_zone_lng;
      dump_d1(NM("soil_scalars"), k, sc);
#define DL(name, arr) dump_d1(NM(name), L, soil_con.arr)
      DL("expt", expt); DL("Ksat", Ksat); DL("phi_s", phi_s); DL("init_moist", init_moist); DL("depth", depth);
      DL("bubble", bubble); DL("quartz", quartz); DL("bulk_density", bulk_density); DL("soil_density", soil_density);
      DL("bulk_dens_min", bulk_dens_min); DL("soil_dens_min", soil_dens_min); DL("organic", organic);
      DL("porosity", porosity); DL("max_moist", max_moist); DL("Wcr", Wcr); DL("Wpwp", Wpwp);
      DL("resid_moist", resid_moist);
      dump_d2(NM("zwtvmoist_zwt"), MAX_LAYERS+2, MAX_ZWTVMOIST, &soil_con.zwtvmoist_zwt[0][0]);
      dump_d2(NM("zwtvmoist_moist"), MAX_LAYERS+2, MAX_ZWTVMOIST, &soil_con.zwtvmoist_moist[0][0]);
      dump_d1(NM("AreaFract"), NB, soil_con.AreaFract);
      dump_d1(NM("Tfactor"), NB, soil_con.Tfactor);
      dump_d1(NM("Pfactor"), NB, soil_con.Pfactor);
      for (b = 0; b < NB; b++) tmp[b] = soil_con.BandElev[b];
      dump_d1(NM("BandElev"), NB, tmp);
      for (b = 0; b < NB; b++) tmp[b] = soil_con.AboveTreeLine[b];
      dump_d1(NM("AboveTreeLine"), NB, tmp);
      /* tiles 0..Nveg (Nveg = bare soil; Cv may be 0) */
      { int cls[MAX_VEG+2]; double cv[MAX_VEG+2];
        for (v = 0; v <= Nveg; v++) { cls[v] = veg_con[v].veg_class; cv[v] = veg_con[v].Cv; }
        dump_i1(NM("veg_class"), Nveg+1, cls); dump_d1(NM("Cv"), Nveg+1, cv); }
      for (v = 0; v < Nveg; v++) for (l = 0; l < L; l++) tmp[v*L+l] = veg_con[v].root[l];
      dump_d2(NM("root"), Nveg, L, tmp);
      for (v = 0; v < Nveg; v++) for (j = 0; j < 12; j++) {
        tmp[(v*4+0)*12+j] = veg_con[v].LAI[j]; tmp[(v*4+1)*12+j] = veg_con[v].albedo[j];
        tmp[(v*4+2)*12+j] = veg_con[v].vegcover[j]; tmp[(v*4+3)*12+j] = veg_con[v].Wdmax[j];
      }
      { int64_t d[3]; d[0] = Nveg; d[1] = 4; d[2] = 12; dump_arr(NM("veg_monthly"), 'd', 3, d, tmp); }
      /* atmos: [rec][field 0..8][NF+1]; field order fixed */
      atmbuf = (double *)malloc(sizeof(double)*(size_t)nrecs*9*(NF+1));
      for (rec = 0; rec < nrecs; rec++) for (j = 0; j <= NF; j++) {
        int jj = (NF == 1) ? 0 : j;   /* NF==1: arrays have 1 element (NR==0) */
        double *p = atmbuf + ((size_t)rec*9)*(NF+1) + j;
        p[0*(NF+1)] = atmos[rec].air_temp[jj]; p[1*(NF+1)] = atmos[rec].prec[jj];
        p[2*(NF+1)] = atmos[rec].wind[jj];     p[3*(NF+1)] = atmos[rec].vp[jj];
        p[4*(NF+1)] = atmos[rec].vpd[jj];      p[5*(NF+1)] = atmos[rec].pressure[jj];
        p[6*(NF+1)] = atmos[rec].density[jj];  p[7*(NF+1)] = atmos[rec].shortwave[jj];
        p[8*(NF+1)] = atmos[rec].longwave[jj];
      }
      { int64_t d[3]; d[0] = nrecs; d[1] = 9; d[2] = NF+1; dump_arr(NM("atmos"), 'd', 3, d, atmbuf); }
      free(atmbuf);
      { unsigned char *sf = (unsigned char *)malloc((size_t)nrecs*(NF+1)); int64_t d[2];
        for (rec = 0; rec < nrecs; rec++) for (j = 0; j <= NF; j++)
          sf[rec*(NF+1)+j] = (unsigned char)atmos[rec].snowflag[(NF == 1) ? 0 : j];
        d[0] = nrecs; d[1] = NF+1; dump_arr(NM("snowflag"), 'b', 2, d, sf); free(sf); }
      if (Nveg > 0) {
        vhbuf = (double *)malloc(sizeof(double)*(size_t)nrecs*Nveg*3);
        for (rec = 0; rec < nrecs; rec++) for (v = 0; v < Nveg; v++) {
          vhbuf[((size_t)rec*Nveg+v)*3+0] = veg_hist[rec][v].albedo[0];
          vhbuf[((size_t)rec*Nveg+v)*3+1] = veg_hist[rec][v].LAI[0];
          vhbuf[((size_t)rec*Nveg+v)*3+2] = veg_hist[rec][v].vegcover[0];
        }
        { int64_t d[3]; d[0] = nrecs; d[1] = Nveg; d[2] = 3; dump_arr(NM("veg_hist"), 'd', 3, d, vhbuf); }
        free(vhbuf);
      }
      outbuf = (double *)malloc(sizeof(double)*(size_t)nrecs*nout_elem);
      nunits = (Nveg+1)*NB;
      if (want_units) unitbuf = (double *)calloc((size_t)nrecs*nunits*NUNITF, sizeof(double));
    }

    Error.filep = filep;
    Error.out_data_files = out_data_files;
    ErrorFlag = put_data(&all_vars, &atmos[0], &soil_con, veg_con, &lake_con, out_data_files, out_data,
                         &save_data, &dmy[0], -nrecs);

    t0 = now_s();
    for (rec = 0; rec < nrecs; rec++) {
      ErrorFlag = full_energy(cellnum, rec, &atmos[rec], &all_vars, dmy, &global_param, &lake_con,
                              &soil_con, veg_con, veg_hist);
      if (ErrorFlag != ERROR)
        ErrorFlag = put_data(&all_vars, &atmos[rec], &soil_con, veg_con, &lake_con, out_data_files,
                             out_data, &save_data, &dmy[rec], rec);
      if (ErrorFlag == ERROR) {
        fprintf(stderr, "ERROR: cell %d failed in record %d\n", soil_con.gridcel, rec);
        break;
      }
      if (g_dump) {
        double *o = outbuf + (size_t)rec*nout_elem;
        for (v = 0; v < N_OUTVAR_TYPES; v++)
          for (i = 0; i < out_data[v].nelem; i++) *o++ = out_data[v].data[i];
        if (want_units) {
          int Nveg = veg_con[0].vegetat_type_num;
          for (veg = 0; veg <= Nveg; veg++) for (band = 0; band < options.SNOW_BAND; band++)
            if (veg_con[veg].Cv > 0 && soil_con.AreaFract[band] > 0)
              fill_unit(unitbuf + (((size_t)rec*(Nveg+1)+veg)*options.SNOW_BAND+band)*NUNITF,
                        &all_vars.cell[veg][band], &all_vars.veg_var[veg][band],
                        &all_vars.snow[veg][band], &all_vars.energy[veg][band], veg < Nveg);
        }
      }
    }
    t_step += now_s() - t0;
    n_cellsteps += rec;
    ncells++;

    if (g_dump) {
      int Nveg = veg_con[0].vegetat_type_num;
      dump_is(NM("nrec_done"), rec);
      dump_d2(NM("out"), nrecs, nout_elem, outbuf);
      free(outbuf);
      if (want_units) {
        int64_t d[4]; d[0] = nrecs; d[1] = Nveg+1; d[2] = options.SNOW_BAND; d[3] = NUNITF;
        dump_arr(NM("units"), 'd', 4, d, unitbuf); free(unitbuf);
      }
    }
    close_files(&filep, out_data_files, &filenames);
    free_veg_hist(nrecs, veg_con[0].vegetat_type_num, &veg_hist);
    free_all_vars(&all_vars, veg_con[0].vegetat_type_num);
    free_vegcon(&veg_con);
    free((char *)soil_con.AreaFract); free((char *)soil_con.BandElev);
    free((char *)soil_con.Tfactor);   free((char *)soil_con.Pfactor);
    free((char *)soil_con.AboveTreeLine);
  }
  dump_is("ncells", ncells);
  if (g_dump) fclose(g_dump);
  if (want_time)
    printf("{\"cells\": %d, \"cell_steps\": %ld, \"step_loop_s\": %.6f, \"forcing_prep_s\": %.6f}\n",
           ncells, n_cellsteps, t_step, t_atmos);
  return 0;
}
