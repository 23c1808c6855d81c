/* rout_gpu.c -- `rout <configuration file>` as a plain C host program over the C ABI of libvicres.so
 * (include/vicres_route.h).
 *
 * What the reference's main program does (routing/model/rout.f:125-566) -- read the positional configuration file, the
 * ESRI grids (READ_DIREC reservoir.f:149-211, READ_VELO/DIFF/XMASK/FRACTION init_routines.f), stations.txt, UH.all, the
 * `fluxes_<lat>_<lng>` files (reservoir.f:373-389) and the resN.txt files (reservoir.f:1021-1131), run MAKE_UHM /
 * MAKE_GRID_UH / MAKE_CONVOLUTION / MAKE_CONVOLUTIONRS, write <station>.day, .day_mm, .month, .month_mm, .year,
 * .year_mm (write_routines.f:7-70, 248-331) and reservoir_<id>_<station>.day (:172-200), or one day of the STEPBYSTEP
 * version (rout.f:255-317, reservoir.f:1596-1675) -- this program does with
 *
 *   vr_route_create -> vr_route_convolve_in -> vr_route_reservoirs | vr_route_reservoirs_range
 *
 * File handling only: every routed number comes from the library, which has no CPU path.  The monthly / yearly
 * summaries are fp32 sums of the daily discharge in WRITE_MONTH's order.  The same program as
 * `python -m vicres_b200.rout` (tests/test_rout.py compares their files).
 *
 * Build:  gcc -O2 -Iinclude -o host/rout_gpu host/rout_gpu.c -Lvicres_b200 -lvicres -lm -Wl,-rpath,$PWD/vicres_b200
 * Usage:  rout_gpu <configuration file> [--legacy] [--device N]
 *         --legacy = the arithmetic variant of the shipped prebuilt binary
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include "vicres_route.h"

static const int DAYS_IN_MONTH[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};

static void die(const char *what, const char *why)
{
  fprintf(stderr, "rout_gpu: %s%s%s\n", what, why ? ": " : "", why ? why : "");
  exit(1);
}
static void *xmalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) die("out of memory", NULL); return p; }
static int isleap(int y) { return (y % 4 == 0 && y % 100 != 0) || y % 400 == 0; }
static int torf(const char *s) { while (*s == ' ' || *s == '\t') s++; return tolower((unsigned char)s[0]) == '.' && tolower((unsigned char)s[1]) == 't'; }
static int file_exists(const char *p) { struct stat st; return stat(p, &st) == 0; }
static void mkdirs(const char *p) { char b[1024]; snprintf(b, sizeof b, "%s", p); for (char *q = b + 1; *q; q++) if (*q == '/') { *q = 0; mkdir(b, 0777); *q = '/'; } mkdir(b, 0777); }

/* ---- text files as arrays of lines ---- */
typedef struct { char **l; int n; } Lines;
static Lines read_lines(const char *path)
{
  FILE *f = fopen(path, "r");
  if (!f) die("cannot open", path);
  Lines L = {NULL, 0};
  int cap = 0;
  char *buf = NULL;
  size_t bl = 0;
  ssize_t k;
  while ((k = getline(&buf, &bl, f)) >= 0) {
    while (k > 0 && (buf[k - 1] == '\n' || buf[k - 1] == '\r')) buf[--k] = 0;
    if (L.n == cap) { cap = cap ? 2 * cap : 64; L.l = realloc(L.l, cap * sizeof(char *)); }
    L.l[L.n++] = strdup(buf);
  }
  free(buf);
  fclose(f);
  return L;
}
static char *strip(char *s) { while (*s == ' ' || *s == '\t') s++; char *e = s + strlen(s); while (e > s && (e[-1] == ' ' || e[-1] == '\t')) *--e = 0; return s; }
static int blank(const char *s) { while (*s) if (!isspace((unsigned char)*s++)) return 0; return 1; }
/* whitespace-separated numbers of a line */
static int numbers(const char *s, double *v, int max)
{
  int n = 0;
  char *e;
  while (n < max) {
    double x = strtod(s, &e);
    if (e == s) break;
    v[n++] = x;
    s = e;
  }
  return n;
}

/* ---- configuration file (rout.f:131-317), positional ---- */
typedef struct {
  char flowdirection[512], stations[512], inpath[512], outpath[512], reservoir_location[512], reservoir_path[512], nf_path[512],
      uh_path[512];
  char gridfile[4][512];       /* velocity, diffusion, xmask, fraction: a file ... */
  int isfile[4];
  double gridconst[4];         /* ... or a constant */
  int dprec, start_year, start_month, stop_year, stop_month, first_year, first_month, last_year, last_month;
  int stepbystep, have_sbs, sbs[6], coupler_iteration, write_output;
} Config;

static Config read_configuration(const char *path)
{
  Lines L = read_lines(path);
  Config c;
  memset(&c, 0, sizeof c);
  int i = 0;
#define NXT() (i < L.n ? strip(L.l[i++]) : (die("configuration file too short", path), (char *)NULL))
  NXT(); NXT();
  snprintf(c.flowdirection, 512, "%s", NXT());
  for (int k = 0; k < 4; k++) {
    NXT();
    c.isfile[k] = torf(NXT());
    char *v = NXT();
    if (c.isfile[k]) snprintf(c.gridfile[k], 512, "%s", v);
    else c.gridconst[k] = strtod(v, NULL);
  }
  NXT();
  snprintf(c.stations, 512, "%s", NXT());
  /* today's source reads a grid-location block here (rout.f:181-193); the shipped binary's layout does not have it */
  int save = i;
  NXT();
  char *probe = NXT();
  if (probe[0] == '.' && (tolower((unsigned char)probe[1]) == 't' || tolower((unsigned char)probe[1]) == 'f')) {
    NXT(); NXT(); NXT();
    snprintf(c.inpath, 512, "%s", NXT());
    c.dprec = atoi(NXT());
    NXT();                                 /* VIC driver line */
  }
  else {
    i = save;
    NXT();
    snprintf(c.inpath, 512, "%s", NXT());
    c.dprec = atoi(NXT());
  }
  NXT(); snprintf(c.outpath, 512, "%s", NXT());
  NXT(); snprintf(c.reservoir_location, 512, "%s", NXT());
  NXT(); snprintf(c.reservoir_path, 512, "%s", NXT());
  NXT(); NXT();                            /* naturalised-flow flag (not used: the flows are always computed) */
  snprintf(c.nf_path, 512, "%s", NXT());
  NXT();
  double v[8];
  if (numbers(NXT(), v, 4) < 4) die("configuration file", "START/STOP year and month");
  c.start_year = (int)v[0]; c.start_month = (int)v[1]; c.stop_year = (int)v[2]; c.stop_month = (int)v[3];
  if (numbers(NXT(), v, 4) < 4) die("configuration file", "FIRST/LAST year and month");
  c.first_year = (int)v[0]; c.first_month = (int)v[1]; c.last_year = (int)v[2]; c.last_month = (int)v[3];
  NXT(); snprintf(c.uh_path, 512, "%s", NXT());
  NXT(); c.stepbystep = torf(NXT());
  c.coupler_iteration = 1; c.write_output = 1;
  if (i + 1 < L.n) {                       /* rout.f:262-316 (absent from the binary's older layout) */
    i++;
    if (numbers(L.l[i++], v, 6) == 6) { c.have_sbs = 1; for (int k = 0; k < 6; k++) c.sbs[k] = (int)v[k]; }
    if (i + 1 < L.n) { i++; c.coupler_iteration = atoi(L.l[i++]); }
    if (i + 1 < L.n) { i++; c.write_output = torf(L.l[i++]); }
  }
#undef NXT
  return c;
}

/* NDAY_SIM of the STEPBYSTEP version (rout.f:266-287) */
static int nday_sim(const Config *c)
{
  int n = 0, m = c->sbs[1], y = c->sbs[0];
  if (c->sbs[3] > c->sbs[0] || c->sbs[4] > c->sbs[1])
    for (int j = c->sbs[1]; j < 12 * (c->sbs[3] - c->sbs[0]) + c->sbs[4]; j++) {
      n += DAYS_IN_MONTH[m - 1] + (m == 2 ? isleap(y) : 0);
      if (++m > 12) { m = 1; y++; }
    }
  return n + c->sbs[5];
}

/* ---- grids ---- */
typedef struct { int nrow, ncol; double xll, yll, size; } Hdr;
static int *read_esri_int(const char *path, int header, Hdr *h)
{
  Lines L = read_lines(path);
  int first = 0, seen = 0;
  for (; first < L.n && seen < header; first++) {
    if (blank(L.l[first])) continue;
    char key[64];
    double val;
    if (sscanf(L.l[first], "%63s %lf", key, &val) == 2 && h) {
      for (char *p = key; *p; p++) *p = (char)tolower((unsigned char)*p);
      if (!strcmp(key, "ncols")) h->ncol = (int)val;
      else if (!strcmp(key, "nrows")) h->nrow = (int)val;
      else if (!strcmp(key, "xllcorner")) h->xll = val;
      else if (!strcmp(key, "yllcorner")) h->yll = val;
      else if (!strcmp(key, "cellsize")) h->size = val;
    }
    seen++;
  }
  int nrow = 0, ncol = 0;
  for (int k = first; k < L.n; k++) if (!blank(L.l[k])) nrow++;
  double *tmp = xmalloc(sizeof(double) * 1 << 20);
  int *a = NULL, r = 0;
  for (int k = first; k < L.n; k++) {
    if (blank(L.l[k])) continue;
    int n = numbers(L.l[k], tmp, 1 << 20);
    if (!a) { ncol = n; a = xmalloc(sizeof(int) * (size_t)nrow * ncol); }
    for (int c = 0; c < ncol && c < n; c++) a[(size_t)r * ncol + c] = (int)tmp[c];
    r++;
  }
  free(tmp);
  if (h && !h->nrow) { h->nrow = nrow; h->ncol = ncol; }
  if (h && (h->nrow != nrow || h->ncol != ncol)) { h->nrow = nrow; h->ncol = ncol; }
  return a;
}
static float *read_grid_f32(const Config *c, int k, int nrow, int ncol)
{
  float *a = xmalloc(sizeof(float) * (size_t)nrow * ncol);
  if (!c->isfile[k]) { for (size_t q = 0; q < (size_t)nrow * ncol; q++) a[q] = (float)c->gridconst[k]; return a; }
  Lines L = read_lines(c->gridfile[k]);
  double *tmp = xmalloc(sizeof(double) * 1 << 20);
  int seen = 0, r = 0;
  for (int q = 0; q < L.n && r < nrow; q++) {
    if (blank(L.l[q])) continue;
    if (seen++ < 6) continue;                       /* 6 header lines */
    int n = numbers(L.l[q], tmp, 1 << 20);
    for (int cc = 0; cc < ncol && cc < n; cc++) a[(size_t)r * ncol + cc] = (float)tmp[cc];
    r++;
  }
  free(tmp);
  return a;
}
/* centre of file-order cell (r, c): JLOC / ILOC (reservoir.f:353-356) */
static void cell_latlon(const Hdr *h, int r, int c, double *lat, double *lon)
{
  const int I = h->ncol - c, J = h->nrow - r;
  *lat = h->yll + J * h->size - h->size / 2.0;
  *lon = h->xll + (h->ncol - I) * h->size + h->size / 2.0;
}
/* FACTOR of every cell in REAL arithmetic (reservoir.f:357-364) */
static float *cell_factors(const Hdr *h, const float *frac)
{
  const float size = (float)h->size, PI = atanf(1.0f) * 4.0f, RERD = 6371229.0f;
  float *out = xmalloc(sizeof(float) * (size_t)h->nrow * h->ncol);
  for (int r = 0; r < h->nrow; r++)
    for (int c = 0; c < h->ncol; c++) {
      const float jloc = (float)h->yll + (float)(h->nrow - r) * size - size / 2.0f;
      const float area = RERD * RERD * fabsf(size) * PI / 180.0f *
                         fabsf(sinf((jloc - size / 2.0f) * PI / 180.0f) - sinf((jloc + size / 2.0f) * PI / 180.0f));
      out[(size_t)r * h->ncol + c] = frac[(size_t)r * h->ncol + c] * 35.315f * area / (86400.0f * 1000.0f);
    }
  return out;
}

/* ---- resN.txt (reservoir.f:1021-1131): today's layout and the older one of the shipped res_par files ---- */
static float *read_series(const char *fn, int start_year, int start_month, int ndays)
{
  Lines L = read_lines(fn);
  float *out = xmalloc(sizeof(float) * ndays);
  int nrows = 0, start = 0;
  float *vals = xmalloc(sizeof(float) * (L.n + 1));
  for (int k = 1; k < L.n; k++) {
    double v[4];
    if (blank(L.l[k]) || numbers(L.l[k], v, 4) < 4) continue;
    if ((int)v[0] == start_year && (int)v[1] == start_month && (int)v[2] == 1) start = nrows;     /* STARTDAY = L-1 (:1087-1089) */
    vals[nrows++] = (float)v[3];
  }
  for (int i = 0; i < ndays && start + i < nrows; i++) out[i] = vals[start + i];
  free(vals);
  return out;
}
static vr_reservoir read_reservoir_file(const char *path, int start_year, int start_month, int ndays)
{
  Lines L = read_lines(path);
  vr_reservoir d;
  memset(&d, 0, sizeof d);
  int i = 0;
  double f[16];
#define NXT() (i < L.n ? L.l[i++] : (die("reservoir file too short", path), (char *)NULL))
  NXT();
  if (numbers(NXT(), f, 8) < 8) die("reservoir file", path);
  d.hresermax = (float)f[0]; d.hresermin = (float)f[1]; d.v_live = (float)f[2]; d.v_dead = (float)f[3]; d.head = (float)f[4];
  d.q_design = (float)f[5]; d.ope_year = (int)f[6]; d.v_initial = (float)f[7];
  NXT();
  numbers(NXT(), f, 2);
  d.seepage = (float)f[0]; d.infiltration = (float)f[1];
  NXT();
  const int with_irr = atoi(NXT());
  char ipath[512];
  snprintf(ipath, sizeof ipath, "%s", strip(NXT()));
  NXT();                                     /* HYDROPOWER header */
  char *a = NXT(), *b = NXT();
  double fa = 0, fb = 0;
  if (numbers(a, &fa, 1) == 1 && numbers(b, &fb, 1) == 1) {       /* today's layout: type, generator, header, env flow */
    d.hydropower_type = (int)fa; d.hydropower_generator = (int)fb;
    NXT();
  }
  else { d.hydropower_type = 1; d.hydropower_generator = (int)fa; }   /* older layout: b was the ENVIRONMENTAL FLOW header */
  numbers(NXT(), f, 1); d.env_flow = (float)f[0];
  NXT();
  d.rule = atoi(NXT());
  if (d.rule == 0) NXT();
  else if (d.rule == 1) { numbers(NXT(), f, 4); d.hmax = (float)f[0]; d.hmin = (float)f[1]; d.op1[0] = (float)f[2]; d.op1[1] = (float)f[3]; }
  else if (d.rule == 2) { numbers(NXT(), f, 12); for (int m = 0; m < 12; m++) d.reslv[m] = (float)f[m]; }
  else if (d.rule == 3) { numbers(NXT(), f, 5); d.demand = (float)f[0]; d.x1 = (float)f[1]; d.x2 = (float)f[2]; d.x3 = (float)f[3]; d.x4 = (float)f[4]; }
  else if (d.rule == 4 || d.rule == 6) d.series = read_series(strip(NXT()), start_year, start_month, ndays);
  else if (d.rule == 5)
    for (int m = 0; m < 12; m++) {
      numbers(NXT(), f, 5);
      d.demand5[m] = (float)f[0]; d.op5x1[m] = (float)f[1]; d.op5x2[m] = (float)f[2]; d.op5x3[m] = (float)f[3]; d.op5x4[m] = (float)f[4];
    }
  NXT();
  d.bathymetry = atoi(NXT());
  if (d.bathymetry == 1) { numbers(NXT(), f, 4); for (int m = 0; m < 4; m++) d.bth[m] = (float)f[m]; }
  if (with_irr == 1) {
    Lines R = read_lines(ipath);
    const int n = R.n > 1 ? atoi(R.l[1]) : 0;
    float *irr = xmalloc(sizeof(float) * ndays);
    for (int k = 0; k < n && k < ndays && 3 + k < R.n; k++) irr[k] = (float)strtod(R.l[3 + k], NULL);
    d.irrigation = irr;
  }
#undef NXT
  return d;
}

/* the nine value columns of reservoir_<id>_<station>.day (write_routines.f:172-200); legacy: as the shipped binary prints them */
static void reservoir_day_columns(const vr_res_series *s, int j, int ndays, const vr_reservoir *r, int start_year, int legacy, int k,
                                  float c[9])
{
  const size_t b1 = (size_t)j * (ndays + 1), b0 = (size_t)j * ndays;
  const float full = (start_year + (k + 1) / 365 < r->ope_year) ? 0.0f : r->v_live + r->v_dead;
  c[0] = s->vol[b1 + k]; c[1] = s->vol[b1 + k + 1]; c[2] = full; c[3] = s->hho[b1 + k]; c[4] = s->hho[b1 + k + 1];
  c[5] = s->flowin[b0 + k]; c[6] = s->flowout[b0 + k]; c[7] = s->turbine[b0 + k]; c[8] = s->energy[b0 + k];
  if (legacy) {
    const int z[4] = {0, 3, 5, 8};
    for (int q = 0; q < 4; q++) if (c[z[q]] < 0) c[z[q]] = 0;
    if (c[2] > 0) c[7] = c[6] < r->q_design ? c[6] : r->q_design;
  }
}

static float *load_matrix(const char *path, int rows, int cols)      /* whitespace-separated rows of numbers */
{
  Lines L = read_lines(path);
  float *a = xmalloc(sizeof(float) * (size_t)(rows ? rows : 1) * cols);
  double *tmp = xmalloc(sizeof(double) * (cols + 8));
  int r = 0;
  for (int k = 0; k < L.n && r < rows; k++) {
    if (blank(L.l[k])) continue;
    int n = numbers(L.l[k], tmp, cols);
    for (int c = 0; c < n; c++) a[(size_t)r * cols + c] = (float)tmp[c];
    r++;
  }
  free(tmp);
  return a;
}

int main(int argc, char **argv)
{
  const char *cfgpath = NULL;
  int legacy = 0, device = 0;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--legacy")) legacy = 1;
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    else cfgpath = argv[i];
  }
  if (!cfgpath) die("usage", "rout_gpu <configuration file> [--legacy] [--device N]");
  Config cfg = read_configuration(cfgpath);
  const int sbs = cfg.stepbystep;
  if (sbs && !cfg.have_sbs) die("STEPBYSTEP mode needs the start / current simulation day line of the configuration file", NULL);

  Hdr hdr;
  memset(&hdr, 0, sizeof hdr);
  int *dir = read_esri_int(cfg.flowdirection, 6, &hdr);
  const int nrow = hdr.nrow, ncol = hdr.ncol;
  const size_t ncell = (size_t)nrow * ncol;
  int *reser = read_esri_int(cfg.reservoir_location, 0, NULL);
  float *velo = read_grid_f32(&cfg, 0, nrow, ncol), *diff = read_grid_f32(&cfg, 1, nrow, ncol), *xmask = read_grid_f32(&cfg, 2, nrow, ncol),
        *frac = read_grid_f32(&cfg, 3, nrow, ncol);
  float box[VR_UH_KE];
  {
    char p[1024];
    snprintf(p, sizeof p, "%sUH.all", cfg.uh_path);
    if (!file_exists(p)) snprintf(p, sizeof p, "%s/UH.all", cfg.uh_path);
    Lines L = read_lines(p);
    int k = 0;
    for (int q = 0; q < L.n && k < VR_UH_KE; q++) { double v[2]; if (numbers(L.l[q], v, 2) == 2) box[k++] = (float)v[1]; }
    if (k < VR_UH_KE) die("UH.all", "fewer than 12 taps");
  }
  /* the period START_YEAR/START_MO .. STOP_YEAR/STOP_MO (rout.f:226-247) */
  int ndays = 0;
  {
    int y = cfg.start_year, m = cfg.start_month;
    for (int j = cfg.start_month; j <= 12 * (cfg.stop_year - cfg.start_year) + cfg.stop_month; j++) {
      ndays += DAYS_IN_MONTH[m - 1] + (m == 2 ? isleap(y) : 0);
      if (++m > 12) { m = 1; y++; }
    }
  }
  int *dy = xmalloc(sizeof(int) * ndays), *dm = xmalloc(sizeof(int) * ndays), *dd = xmalloc(sizeof(int) * ndays);
  {
    int y = cfg.start_year, m = cfg.start_month, k = 0;
    while (k < ndays) {
      const int n = DAYS_IN_MONTH[m - 1] + (m == 2 ? isleap(y) : 0);
      for (int d = 1; d <= n && k < ndays; d++, k++) { dy[k] = y; dm[k] = m; dd[k] = d; }
      if (++m > 12) { m = 1; y++; }
    }
  }
  /* flux files: the 2-layer default layout's columns RUNOFF, BASEFLOW, EVAP_CANOP, TRANSP_VEG, EVAP_BARE (reservoir.f:383-386) */
  float *runoff = xmalloc(sizeof(float) * ncell * ndays), *baseflow = xmalloc(sizeof(float) * ncell * ndays),
        *ev_vege = xmalloc(sizeof(float) * ncell * ndays), *tr_vege = xmalloc(sizeof(float) * ncell * ndays),
        *ev_soil = xmalloc(sizeof(float) * ncell * ndays);
  unsigned char *present = xmalloc(ncell);
  {
    double *tmp = xmalloc(sizeof(double) * 64);
    for (int r = 0; r < nrow; r++)
      for (int c = 0; c < ncol; c++) {
        double lat, lon;
        cell_latlon(&hdr, r, c, &lat, &lon);
        char fn[1024];
        snprintf(fn, sizeof fn, "%s%.*f_%.*f", cfg.inpath, cfg.dprec, lat, cfg.dprec, lon);
        if (!file_exists(fn)) continue;
        const size_t g = (size_t)r * ncol + c;
        present[g] = 1;
        Lines L = read_lines(fn);
        int day = 0;
        for (int q = 0; q < L.n && day < ndays; q++) {
          if (blank(L.l[q])) continue;
          const int n = numbers(L.l[q], tmp, 64);
          if (n > 14) {
            runoff[g * ndays + day] = (float)tmp[5]; baseflow[g * ndays + day] = (float)tmp[6];
            ev_vege[g * ndays + day] = (float)tmp[12]; tr_vege[g * ndays + day] = (float)tmp[13]; ev_soil[g * ndays + day] = (float)tmp[14];
          }
          day++;
        }
        for (int q = 0; q < L.n; q++) free(L.l[q]);
        free(L.l);
      }
    free(tmp);
  }
  float *fac = cell_factors(&hdr, frac), *cell_scale = NULL;
  if (legacy) { cell_scale = xmalloc(sizeof(float) * ncell); for (size_t g = 0; g < ncell; g++) cell_scale[g] = fac[g] * 0.028f; }

  vr_route_grid grid;
  grid.ncol = ncol; grid.nrow = nrow; grid.dir = dir; grid.reser = reser; grid.velo = velo; grid.diff = diff; grid.xmask = xmask;

  /* stations.txt: `<active> <name> <col> <row> <area>` then the name of a UH_S file or NONE (rout.f:330-360) */
  Lines S = read_lines(cfg.stations);
  int W = 0;
  for (int q = 0; q + 1 < S.n; ) {
    while (q < S.n && blank(S.l[q])) q++;
    if (q >= S.n) break;
    int active, col, row;
    char name[64];
    if (sscanf(S.l[q], "%d %63s %d %d", &active, name, &col, &row) != 4) break;
    q++;
    while (q < S.n && blank(S.l[q])) q++;
    q++;                                            /* the UH_S line */
    if (active != 1) continue;
    W++;
    vr_route *R = NULL;
    if (vr_route_create(&grid, col, row, box, device, &R) != VR_OK) die("vr_route_create", vr_route_last_error(NULL));
    const int nnodes = vr_route_num_nodes(R), nres = nnodes - 1;
    vr_route_inputs in;
    memset(&in, 0, sizeof in);
    in.runoff = runoff; in.baseflow = baseflow; in.scale = 0.18308f; in.cell_scale = cell_scale;
    in.ev_vege = ev_vege; in.ev_soil = ev_soil; in.tr_vege = tr_vege; in.cell_factor = fac; in.present = present;
    in.start_year = cfg.start_year; in.start_month = cfg.start_month;
    float *nat = xmalloc(sizeof(float) * (size_t)nnodes * ndays), *resev = xmalloc(sizeof(float) * (size_t)nnodes * ndays);
    if (vr_route_convolve_in(R, ndays, &in, nat, resev) != VR_OK) die("vr_route_convolve_in", vr_route_last_error(R));
    /* reservoirs of the nodes, FACTOR_SUM of the station's node (reservoir.f:307, 372) */
    int *ids = xmalloc(sizeof(int) * (nres + 1));
    vr_reservoir *res = xmalloc(sizeof(vr_reservoir) * (nres + 1));
    for (int n = 0; n < nres; n++) {
      int64_t nc;
      vr_route_node_info(R, n, &ids[n], &nc);
      char p[1024];
      snprintf(p, sizeof p, "%sres%d.txt", cfg.reservoir_path, ids[n]);
      if (!file_exists(p)) snprintf(p, sizeof p, "%s/res%d.txt", cfg.reservoir_path, ids[n]);
      res[n] = read_reservoir_file(p, cfg.start_year, cfg.start_month, ndays);
    }
    float fsum = 0.0f;
    {
      int sid;
      int64_t nc;
      vr_route_node_info(R, nnodes - 1, &sid, &nc);
      int32_t *cells = xmalloc(sizeof(int32_t) * (size_t)nc);
      vr_route_node_cells(R, nnodes - 1, cells);
      for (int64_t k = 0; k < nc; k++) fsum = fsum + fac[cells[k]];
      free(cells);
    }
    vr_res_options opt;
    memset(&opt, 0, sizeof opt);
    opt.start_year = cfg.start_year; opt.start_month = cfg.start_month; opt.legacy_rule_curve = legacy;
    vr_res_series ser;
    const size_t n1 = (size_t)(nres ? nres : 1) * (ndays + 1), n0 = (size_t)(nres ? nres : 1) * ndays;
    ser.vol = xmalloc(sizeof(float) * n1); ser.hho = xmalloc(sizeof(float) * n1); ser.flowin = xmalloc(sizeof(float) * n0);
    ser.flowout = xmalloc(sizeof(float) * n0); ser.turbine = xmalloc(sizeof(float) * n0); ser.energy = xmalloc(sizeof(float) * n0);
    ser.htk = xmalloc(sizeof(float) * n0);
    float *flow = xmalloc(sizeof(float) * ndays);
    char p[1024], sd_nf[1024], sd_out[1024];
    if (!sbs) {
      if (vr_route_reservoirs(R, ndays, 1, nat, res, &opt, flow, &ser) != VR_OK) die("vr_route_reservoirs", vr_route_last_error(R));
    }
    else {
      /* one day per run (reservoir.f:1142-1150, 1596-1675): VOL of the day and the releases so far come from the files the
         previous run left, this run's go to the files of the next */
      const int k = nday_sim(&cfg);
      snprintf(sd_nf, sizeof sd_nf, "%s/stepbystep", strip(cfg.nf_path));
      snprintf(sd_out, sizeof sd_out, "%s/stepbystep", cfg.outpath);
      mkdirs(sd_nf); mkdirs(sd_out);
      if (k < 1 || k > ndays) die("STEPBYSTEP", "the simulation day lies outside the period");
      if (k > 1) {
        char fo[1200], vo[1200];
        snprintf(fo, sizeof fo, "%s/OUTFLOW%d_STEP%d.txt", sd_nf, W, k);
        snprintf(vo, sizeof vo, "%s/RESERVOIR_VOL_STATION%d_STEP%d.txt", sd_out, W, k);
        if (!file_exists(fo) || !file_exists(vo))
          die("ERROR in Reading Storage Volume for Current Time Step; Please Check if Previous Time Step is Simulated", fo);
        if (nres) {
          float *a = load_matrix(fo, nres, ndays), *v = load_matrix(vo, nres, 1);
          memcpy(ser.flowout, a, sizeof(float) * (size_t)nres * ndays);
          for (int j = 0; j < nres; j++) ser.vol[(size_t)j * (ndays + 1) + k - 1] = v[j];
          free(a); free(v);
        }
      }
      if (vr_route_reservoirs_range(R, ndays, 1, k - 1, k, nat, res, &opt, flow, &ser) != VR_OK)
        die("vr_route_reservoirs_range", vr_route_last_error(R));
      snprintf(p, sizeof p, "%s/OUTFLOW%d_STEP%d.txt", sd_nf, W, k + 1);
      FILE *f = fopen(p, "w");
      if (!f) die("cannot write", p);
      for (int j = 0; j < nres; j++) { for (int i = 0; i < ndays; i++) fprintf(f, "%s%.9g", i ? " " : "", ser.flowout[(size_t)j * ndays + i]); fprintf(f, "\n"); }
      fclose(f);
      snprintf(p, sizeof p, "%s/NF%d_STEP%d.txt", sd_nf, W, k + 1);          /* RESFLOWS: the inflows seen so far */
      f = fopen(p, "w");
      if (!f) die("cannot write", p);
      for (int j = 0; j <= nres; j++) {
        for (int i = 0; i < ndays; i++) {
          const float v = i < k ? (j < nres ? ser.flowin[(size_t)j * ndays + i] : flow[i]) : nat[(size_t)j * ndays + i];
          fprintf(f, "%s%.9g", i ? " " : "", v);
        }
        fprintf(f, "\n");
      }
      fclose(f);
      for (int kk = (k == 1 ? k : k + 1); kk <= k + 1; kk++) {
        snprintf(p, sizeof p, "%s/RESERVOIR_VOL_STATION%d_STEP%d.txt", sd_out, W, kk);
        f = fopen(p, "w");
        if (!f) die("cannot write", p);
        for (int j = 0; j < nres; j++) fprintf(f, "%.9g\n", ser.vol[(size_t)j * (ndays + 1) + kk - 1]);
        fclose(f);
      }
      if (k > 2) {
        snprintf(p, sizeof p, "%s/NF%d_STEP%d.txt", sd_nf, W, k - 1); unlink(p);
        snprintf(p, sizeof p, "%s/OUTFLOW%d_STEP%d.txt", sd_nf, W, k - 1); unlink(p);
      }
      if (k > 1) { snprintf(p, sizeof p, "%s/RESERVOIR_VOL_STATION%d_STEP%d.txt", sd_out, W, k - 1); unlink(p); }
    }
    vr_route_destroy(R);

    if (cfg.write_output && sbs) {
      /* WRITE_DATA's step-by-step branch (write_routines.f:42-58): one line per run, appended */
      const int k = nday_sim(&cfg);
      snprintf(p, sizeof p, "%s/STEPBYSTEP", cfg.outpath);
      mkdirs(p);
      const char *mode = (k == 1 && cfg.coupler_iteration == 1) ? "w" : "a";
      snprintf(p, sizeof p, "%s/STEPBYSTEP/%s.day", cfg.outpath, name);
      FILE *f = fopen(p, mode);
      if (!f) die("cannot write", p);
      fprintf(f, "%12d%12d%12d   %.9g\n", cfg.sbs[3], cfg.sbs[4], cfg.sbs[5], flow[k - 1]);
      fclose(f);
      snprintf(p, sizeof p, "%s/STEPBYSTEP/%s.day_mm", cfg.outpath, name);
      f = fopen(p, mode);
      if (!f) die("cannot write", p);
      fprintf(f, "%12d%12d%12d   %.9g\n", cfg.sbs[3], cfg.sbs[4], cfg.sbs[5], flow[k - 1] / fsum);
      fclose(f);
    }
    else if (cfg.write_output) {
      mkdirs(cfg.outpath);
      FILE *f, *g;
      snprintf(p, sizeof p, "%s%s.day", cfg.outpath, name); f = fopen(p, "w");
      snprintf(p, sizeof p, "%s%s.day_mm", cfg.outpath, name); g = fopen(p, "w");
      if (!f || !g) die("cannot write", p);
      for (int i = 0; i < ndays; i++) {
        fprintf(f, "%12d%12d%12d   %.9g\n", dy[i], dm[i], dd[i], flow[i]);
        fprintf(g, "%12d%12d%12d   %.9g\n", dy[i], dm[i], dd[i], flow[i] / fsum);
      }
      fclose(f); fclose(g);
      /* WRITE_MONTH (write_routines.f:248-331) in fp32 and in its order; a February's divisor is 28 whatever the year */
      const int nmonths = 12 * (cfg.stop_year - cfg.start_year) + cfg.stop_month - cfg.start_month + 1, nm = 12 * (cfg.stop_year - cfg.start_year + 1);
      int *mo = xmalloc(sizeof(int) * (nmonths + 1)), *yr = xmalloc(sizeof(int) * (nmonths + 1));
      for (int i = 0, m = cfg.start_month, y = cfg.start_year; i < nmonths; i++) { mo[i] = m; yr[i] = y; if (++m > 12) { m = 1; y++; } }
      float *monthly = xmalloc(sizeof(float) * (nm + 1)), *monthly_mm = xmalloc(sizeof(float) * (nm + 1));
      for (int i = 0, M = 0, old = mo[0]; i < ndays; i++) {
        if (dm[i] != old) { M++; old = dm[i]; }
        monthly[M] = monthly[M] + flow[i] / (float)DAYS_IN_MONTH[dm[i] - 1];
        monthly_mm[M] = monthly_mm[M] + flow[i] / fsum;
      }
      const int skipto = (cfg.first_year - cfg.start_year) * 12 + (cfg.first_month - cfg.start_month) + 1;
      const int stopat = nmonths - ((cfg.stop_year - cfg.last_year) * 12 + (cfg.stop_month - cfg.last_month));
      float yearly[12] = {0}, yearly_mm[12] = {0};
      int cnt[12] = {0};
      snprintf(p, sizeof p, "%s%s.month", cfg.outpath, name); f = fopen(p, "w");
      snprintf(p, sizeof p, "%s%s.month_mm", cfg.outpath, name); g = fopen(p, "w");
      if (!f || !g) die("cannot write", p);
      for (int i = skipto; i <= stopat; i++) {
        fprintf(f, "%12d%12d   %.9g\n", yr[i - 1], mo[i - 1], monthly[i - 1]);
        fprintf(g, "%12d%12d   %.9g\n", yr[i - 1], mo[i - 1], monthly_mm[i - 1]);
        yearly[(i - 1) % 12] = yearly[(i - 1) % 12] + monthly[i - 1];
        yearly_mm[(i - 1) % 12] = yearly_mm[(i - 1) % 12] + monthly_mm[i - 1];
        cnt[mo[i - 1] - 1]++;
      }
      fclose(f); fclose(g);
      snprintf(p, sizeof p, "%s%s.year", cfg.outpath, name); f = fopen(p, "w");
      snprintf(p, sizeof p, "%s%s.year_mm", cfg.outpath, name); g = fopen(p, "w");
      if (!f || !g) die("cannot write", p);
      for (int i = 0; i < 12; i++) {
        if (cnt[i]) { fprintf(f, "%12d   %.9g\n", i + 1, yearly[i] / (float)cnt[i]); fprintf(g, "%12d   %.9g\n", i + 1, yearly_mm[i] / (float)cnt[i]); }
        else { fprintf(f, "%12d   0\n", i + 1); fprintf(g, "%12d   0\n", i + 1); }
      }
      fclose(f); fclose(g);
      snprintf(p, sizeof p, "%s%s.end_of_month", cfg.outpath, name); f = fopen(p, "w"); if (f) fclose(f);
      for (int j = 0; j < nres; j++) {
        snprintf(p, sizeof p, "%sreservoir_%d_%s.day", cfg.outpath, ids[j], name);
        f = fopen(p, "w");
        if (!f) die("cannot write", p);
        fprintf(f, " Year Month Day Volume_1000cm Volume_t+1_1000cm FullSupplyVolume WaterLevel_m WaterLevel_t+1_m Qinflow_cms \n");
        for (int i = 0; i < ndays; i++) {
          float c9[9];
          reservoir_day_columns(&ser, j, ndays, &res[j], cfg.start_year, legacy, i, c9);
          fprintf(f, "%12d%12d%12d", dy[i], dm[i], dd[i]);
          for (int q = 0; q < 9; q++) fprintf(f, "   %.9g", c9[q]);
          fprintf(f, "\n");
        }
        fclose(f);
      }
    }
    fprintf(stderr, "rout_gpu: station %s, %d days, %d reservoirs\n", name, ndays, nres);
    free(nat); free(resev); free(flow);
  }
  return 0;
}
