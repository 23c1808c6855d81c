/* route_oracle.c -- ORACLE = TEST INFRASTRUCTURE (never linked into or called from vicres_b200/).
 *
 * Plain-C fp32 restatement of the reference's unit-hydrograph routing for one station:
 *   READ_DIREC              routing/model/reservoir.f:149-211
 *   SEARCH_WHOLECATCHMENT   routing/model/init_routines.f:103-168
 *   SEARCH_CATCHMENTRS      routing/model/reservoir.f:578-763   (cell lists only)
 *   MAKE_UHM                routing/model/unit_hyd_routines.f:7-50
 *   MAKE_GRID_UH            routing/model/unit_hyd_routines.f:55-178
 *   MAKE_CONVOLUTION        routing/model/reservoir.f:216-526   (without open-water evaporation)
 *   MAKE_RESERVOIR_UH       routing/model/unit_hyd_routines.f:218-346
 *   MAKE_CONVOLUTIONRS      routing/model/reservoir.f:1142-1684 (day loop; STEPBYSTEP = .FALSE.)
 * The Fortran cannot be compiled here (no Fortran compiler); this restatement is pinned against the
 * reference's shipped prebuilt `rout` binary run as a black box (tests/golden/route_*.npz made by
 * tools/make_route_golden.py: its .uh_s files and naturalised node inflows).  That binary predates
 * the current reservoir.f, so parity for routing is pinned to the binary, not to today's source.
 * Compile with -ffp-contract=off: the accumulations are plain fp32 multiply-then-add.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../include/vicres_route.h"

#define TMAX (VR_UH_DAY * 24)

typedef struct {
  int ncol, nrow;          /* ICOL, IROW */
  int *d1, *d2, *reser;    /* DIREC(I,J,1:2), RESER(I,J); index (J-1)*ncol + (I-1), I,J 1-based Fortran */
  float *velo, *diff, *xmask;
} fgrid;

#define IX(g, I, J) (((J)-1) * (g)->ncol + ((I)-1))

/* file order (r, c) -> Fortran (I, J): J = IROW - r, I = ICOL - c (READ(..) (H(I,J), I=ICOL,1,-1), J=IROW,1,-1) */
static void to_fortran(const vr_route_grid *g, fgrid *f)
{
  int n = g->ncol * g->nrow, r, c, I, J;
  f->ncol = g->ncol; f->nrow = g->nrow;
  f->d1 = calloc(n, sizeof(int)); f->d2 = calloc(n, sizeof(int)); f->reser = calloc(n, sizeof(int));
  f->velo = calloc(n, sizeof(float)); f->diff = calloc(n, sizeof(float)); f->xmask = calloc(n, sizeof(float));
  for (r = 0; r < g->nrow; r++) for (c = 0; c < g->ncol; c++) {
    int h = g->dir[r * g->ncol + c], k;
    I = g->ncol - c; J = g->nrow - r;
    k = IX(f, I, J);
    f->reser[k] = g->reser ? g->reser[r * g->ncol + c] : 0;
    f->velo[k] = g->velo[r * g->ncol + c]; f->diff[k] = g->diff[r * g->ncol + c]; f->xmask[k] = g->xmask[r * g->ncol + c];
    switch (h) {
      case 1: f->d1[k] = I;     f->d2[k] = J + 1; break;
      case 2: f->d1[k] = I - 1; f->d2[k] = J + 1; break;
      case 3: f->d1[k] = I - 1; f->d2[k] = J;     break;
      case 4: f->d1[k] = I - 1; f->d2[k] = J - 1; break;
      case 5: f->d1[k] = I;     f->d2[k] = J - 1; break;
      case 6: f->d1[k] = I + 1; f->d2[k] = J - 1; break;
      case 7: f->d1[k] = I + 1; f->d2[k] = J;     break;
      case 8: f->d1[k] = I + 1; f->d2[k] = J + 1; break;
      default: f->d1[k] = 0; f->d2[k] = 0;
    }
  }
}

static void free_fgrid(fgrid *f) { free(f->d1); free(f->d2); free(f->reser); free(f->velo); free(f->diff); free(f->xmask); }

/* walk from (I,J); returns 1 if it reaches (PI,PJ).  mode 0: whole catchment; 1: reservoir target
 * (FINAL = 0); 2: station remainder (FINAL = 1) */
static int reaches(const fgrid *f, int I, int J, int PI, int PJ, int mode)
{
  int II = I, JJ = J, guard = 0;
  for (;;) {
    if (II > f->ncol || II < 1 || JJ > f->nrow || JJ < 1) return 0;
    if (II == PI && JJ == PJ) return 1;
    if (mode == 2 && f->reser[IX(f, II, JJ)] > 0) return 0;
    if (mode == 1 && f->reser[IX(f, II, JJ)] > 0 && f->reser[IX(f, II, JJ)] != 9999) return 0;
    {
      int a = f->d1[IX(f, II, JJ)], b = f->d2[IX(f, II, JJ)];
      if (a != 0 && b != 0) { II = a; JJ = b; }
      else return 0;
    }
    if (++guard > f->ncol * f->nrow + 4) return 0;    /* a cycle in the direction grid */
  }
}

typedef struct { int res_id, ncells, PI, PJ; int *ci, *cj; int ds_id, target; float uh_r[VR_UH_NS]; } onode;

/* MAKE_UHM for one cell */
static void make_uhm_cell(float velo, float diff, float xmask, float *uh)
{
  float PI_ = 4.0f * atanf(1.0f), T = 0.0f, DT = 3600.0f, INTE = 0.0f, H, POT;
  int K;
  for (K = 0; K < VR_UH_LE; K++) {
    T = T + DT;
    if (velo > 0.0f) {
      POT = powf((velo * T - xmask), 2.0f) / (4.0f * diff * T);
      if (POT > 69.0f) H = 0.0f;
      else H = 1.0f / (2.0f * sqrtf(PI_ * diff)) * xmask / powf(T, 1.5f) * expf(-POT);
    }
    else H = 0.0f;
    uh[K] = H;
  }
  for (K = 0; K < VR_UH_LE; K++) INTE = INTE + uh[K];
  if (INTE > 0.0f) for (K = 0; K < VR_UH_LE; K++) uh[K] = uh[K] / INTE;
}

/* MAKE_GRID_UH for one cell (I,J) of the node (PI,PJ): uh_s[VR_UH_NS] */
static void grid_uh_cell(const fgrid *f, const float *uhm /* [ncell][LE] */, int I, int J, int PI, int PJ,
                         const float *uh_box, float *uh_s)
{
  static float FR1[TMAX + 1], FR2[TMAX + 1];
  float UH_DAILY[VR_UH_DAY + 1], SUM;
  int K, T, L, U;
  for (K = 1; K <= VR_UH_DAY; K++) UH_DAILY[K] = 0.0f;
  for (K = 1; K <= 24; K++) { FR1[K] = 1.0f / 24.0f; FR2[K] = 0.0f; }
  for (K = 25; K <= TMAX; K++) { FR1[K] = 0.0f; FR2[K] = 0.0f; }
  while (I != PI || J != PJ) {
    const float *u = uhm + (size_t)IX(f, I, J) * VR_UH_LE;
    int II, JJ;
    for (T = 1; T <= TMAX; T++)
      for (L = 1; L <= VR_UH_LE; L++)
        if ((T - L) > 0) FR2[T] = FR2[T] + FR1[T - L] * u[L - 1];
    II = f->d1[IX(f, I, J)]; JJ = f->d2[IX(f, I, J)];
    I = II; J = JJ;
    for (T = 1; T <= TMAX; T++) { FR1[T] = FR2[T]; FR2[T] = 0.0f; }
  }
  for (T = 1; T <= TMAX; T++) {
    int TT = (T + 23) / 24;
    UH_DAILY[TT] = UH_DAILY[TT] + FR1[T];
  }
  for (K = 0; K < VR_UH_NS; K++) uh_s[K] = 0.0f;
  for (K = 1; K <= VR_UH_KE; K++)
    for (U = 1; U <= VR_UH_DAY; U++)
      uh_s[K + U - 2] = uh_s[K + U - 2] + uh_box[K - 1] * UH_DAILY[U];
  SUM = 0;
  for (K = 0; K < VR_UH_NS; K++) SUM = SUM + uh_s[K];
  for (K = 0; K < VR_UH_NS; K++) uh_s[K] = uh_s[K] / SUM;
}

typedef struct ro_model {
  fgrid f;
  int nnodes;
  onode *node;
  float *uhm;
  float **uh_s;      /* per node [ncells][NS] */
  vr_route_grid g;
} ro_model;

static void list_cells(const fgrid *f, int PI, int PJ, int mode, onode *nd)
{
  int I, J, n = 0;
  nd->ci = malloc(sizeof(int) * f->ncol * f->nrow); nd->cj = malloc(sizeof(int) * f->ncol * f->nrow);
  for (I = 1; I <= f->ncol; I++) for (J = 1; J <= f->nrow; J++)
    if (reaches(f, I, J, PI, PJ, mode)) { nd->ci[n] = I; nd->cj[n] = J; n++; }
  nd->ncells = n; nd->PI = PI; nd->PJ = PJ;
}

ro_model *ro_create(const vr_route_grid *g, int station_col, int station_row, const float *uh_box)
{
  ro_model *m = calloc(1, sizeof *m);
  onode whole;
  int PI, PJ, n, k, I, J, ncell = g->ncol * g->nrow;
  to_fortran(g, &m->f);
  m->g = *g;
  PI = g->ncol + 1 - station_col;      /* rout.f:341 */
  PJ = station_row;
  /* SEARCH_WHOLECATCHMENT: reservoirs in scan order among the cells that reach the station */
  list_cells(&m->f, PI, PJ, 0, &whole);
  m->node = calloc(whole.ncells + 1, sizeof(onode));
  m->nnodes = 0;
  for (n = 0; n < whole.ncells; n++) {
    int r = m->f.reser[IX(&m->f, whole.ci[n], whole.cj[n])];
    if (r > 0 && r != 9999) {
      onode *nd = &m->node[m->nnodes++];
      list_cells(&m->f, whole.ci[n], whole.cj[n], 1, nd);
      nd->res_id = r;
    }
  }
  {
    onode *nd = &m->node[m->nnodes++];
    list_cells(&m->f, PI, PJ, 2, nd);
    nd->res_id = 0;
  }
  free(whole.ci); free(whole.cj);
  m->uhm = malloc(sizeof(float) * (size_t)ncell * VR_UH_LE);
  for (I = 1; I <= g->ncol; I++) for (J = 1; J <= g->nrow; J++) {
    k = IX(&m->f, I, J);
    make_uhm_cell(m->f.velo[k], m->f.diff[k], m->f.xmask[k], m->uhm + (size_t)k * VR_UH_LE);
  }
  m->uh_s = calloc(m->nnodes, sizeof(float *));
  for (n = 0; n < m->nnodes; n++) {
    onode *nd = &m->node[n];
    m->uh_s[n] = malloc(sizeof(float) * (size_t)(nd->ncells ? nd->ncells : 1) * VR_UH_NS);
    for (k = 0; k < nd->ncells; k++)
      grid_uh_cell(&m->f, m->uhm, nd->ci[k], nd->cj[k], nd->PI, nd->PJ, uh_box, m->uh_s[n] + (size_t)k * VR_UH_NS);
  }
  /* RES_DIRECT(:,2): walk downstream from each reservoir cell (reservoir.f:733-759); UH_R to the next
   * reservoir when it is one of this station's nodes, else to the station (unit_hyd_routines.f:246-270) */
  for (n = 0; n < m->nnodes - 1; n++) {
    onode *nd = &m->node[n];
    int II = nd->PI, JJ = nd->PJ, guard = 0, DI = PI, DJ = PJ;
    nd->ds_id = 0;
    for (;;) {
      int r, a, b;
      if (II > g->ncol || II < 1 || JJ > g->nrow || JJ < 1) { nd->ds_id = 0; break; }
      r = m->f.reser[IX(&m->f, II, JJ)];
      if (r > 0 && r != 9999 && (II != nd->PI || JJ != nd->PJ)) { nd->ds_id = r; break; }
      a = m->f.d1[IX(&m->f, II, JJ)]; b = m->f.d2[IX(&m->f, II, JJ)];
      if (a != 0 && b != 0) { II = a; JJ = b; } else break;
      if (++guard > ncell + 4) break;
    }
    nd->target = m->nnodes - 1;
    for (k = 0; k < m->nnodes - 1; k++)
      if (nd->ds_id > 0 && m->node[k].res_id == nd->ds_id) { nd->target = k; DI = m->node[k].PI; DJ = m->node[k].PJ; break; }
    grid_uh_cell(&m->f, m->uhm, nd->PI, nd->PJ, DI, DJ, uh_box, nd->uh_r);
  }
  m->node[m->nnodes - 1].ds_id = 0; m->node[m->nnodes - 1].target = -1;
  return m;
}

void ro_destroy(ro_model *m)
{
  int n;
  if (!m) return;
  for (n = 0; n < m->nnodes; n++) { free(m->node[n].ci); free(m->node[n].cj); free(m->uh_s[n]); }
  free(m->node); free(m->uh_s); free(m->uhm); free_fgrid(&m->f); free(m);
}

int ro_num_nodes(const ro_model *m) { return m->nnodes; }
int ro_node_info(const ro_model *m, int node, int *res_id, long long *ncells)
{
  *res_id = m->node[node].res_id; *ncells = m->node[node].ncells; return 0;
}
/* file-order grid index of the k-th cell: r = IROW - J, c = ICOL - I */
int ro_node_cells(const ro_model *m, int node, int *grid_index)
{
  int k;
  for (k = 0; k < m->node[node].ncells; k++)
    grid_index[k] = (m->f.nrow - m->node[node].cj[k]) * m->f.ncol + (m->f.ncol - m->node[node].ci[k]);
  return 0;
}
int ro_get_uh(const ro_model *m, int node, float *uh_s)
{
  memcpy(uh_s, m->uh_s[node], sizeof(float) * (size_t)m->node[node].ncells * VR_UH_NS);
  return 0;
}

/* MAKE_CONVOLUTION (reservoir.f:476-503) for every node.
 * cell_scale[gi] multiplies runoff/baseflow (NULL: the hard-coded 0.18308 of reservoir.f:479-480; the
 * shipped prebuilt binary still used FACTOR*0.028).  With ev_vege/ev_soil/tr_vege given, the
 * reference's correction on days whose VIC evaporation components sum below the (zero, outside
 * reservoir-surface cells) open-water evaporation is applied (reservoir.f:496-499).
 * present[gi] == 0 marks a cell whose flux file is missing: the reference inserts zeros for runoff and
 * baseflow but leaves vic_eva_vege/vic_trans_vege/vic_eva_soil holding the PREVIOUS cell's series
 * (reservoir.f:398-409 resets only runo/base/potentialevapo), which the correction then uses. */
int ro_convolve(const ro_model *m, int ndays, const float *runoff, const float *baseflow, const float *cell_scale,
                const float *ev_vege, const float *ev_soil, const float *tr_vege, const float *cell_factor,
                const unsigned char *present, float *flows)
{
  int n, k, I, J;
  long long last;          /* grid cell whose evaporation series is currently in the reference's arrays */
  float *RUNO = malloc(sizeof(float) * ndays), *BASE = malloc(sizeof(float) * ndays);
  for (n = 0; n < m->nnodes; n++) {
    const onode *nd = &m->node[n];
    float *FLOW = flows + (size_t)n * ndays;
    for (I = 0; I < ndays; I++) FLOW[I] = 0.0f;
    last = -1;   /* automatic arrays of MAKE_CONVOLUTION: fresh (read as zeros) at every call */
    for (k = 0; k < nd->ncells; k++) {
      int gi = (m->f.nrow - nd->cj[k]) * m->f.ncol + (m->f.ncol - nd->ci[k]);
      const float *us = m->uh_s[n] + (size_t)k * VR_UH_NS;
      const float sc = cell_scale ? cell_scale[gi] : 0.18308f;
      const int here = !present || present[gi];
      if (here) last = gi;
      for (I = 0; I < ndays; I++) {
        RUNO[I] = here ? runoff[(size_t)gi * ndays + I] * sc : 0.0f * sc;
        BASE[I] = here ? baseflow[(size_t)gi * ndays + I] * sc : 0.0f * sc;
      }
      for (I = 1; I <= ndays; I++) {
        for (J = 1; J <= VR_UH_NS; J++)
          if ((I - J + 1) >= 1) FLOW[I - 1] = FLOW[I - 1] + us[J - 1] * (BASE[I - J] + RUNO[I - J]);
        if (ev_vege && ev_soil && tr_vege && cell_factor && last >= 0) {
          size_t q = (size_t)last * ndays + (I - 1);
          float FACTOR = cell_factor[gi];
          if (0.0f > (ev_vege[q] + ev_soil[q] + tr_vege[q]))
            FLOW[I - 1] = FLOW[I - 1] - 0.0f + (ev_soil[q] + ev_vege[q] + tr_vege[q]) * FACTOR * 0.028f / 1000;
        }
      }
    }
  }
  free(RUNO); free(BASE);
  return 0;
}


/* Penman (1948) tables of MAKE_CONVOLUTION (reservoir.f:271-301): extraterrestrial radiation Ra (mm/day) by
 * month (rows) and 10-degree latitude band 60N..60S (13 columns), delta/gamma and saturation vapour
 * pressure by half degree 0..22 C, sigma*Ta^4 by degree -1..21 C */
static const float RA_TAB[13][12] = {
  {1.4f, 3.6f, 7.0f, 11.1f, 14.6f, 16.4f, 15.6f, 12.6f, 8.5f, 4.7f, 2.0f, 0.9f},
  {3.7f, 6.0f, 9.2f, 12.7f, 15.5f, 16.6f, 16.1f, 13.7f, 10.4f, 7.1f, 4.4f, 3.1f},
  {6.2f, 8.4f, 11.1f, 13.8f, 15.9f, 16.7f, 16.3f, 14.7f, 12.1f, 9.3f, 6.8f, 5.6f},
  {8.1f, 10.5f, 12.8f, 14.7f, 16.1f, 16.5f, 16.2f, 15.2f, 13.5f, 11.2f, 9.1f, 7.9f},
  {10.8f, 12.4f, 14.0f, 15.2f, 15.7f, 15.8f, 15.8f, 15.4f, 14.4f, 12.9f, 11.3f, 10.4f},
  {12.8f, 13.9f, 14.8f, 15.2f, 15.0f, 14.8f, 14.9f, 15.0f, 14.8f, 14.2f, 13.1f, 12.5f},
  {14.6f, 15.0f, 15.2f, 14.7f, 13.9f, 13.4f, 13.6f, 14.3f, 14.9f, 15.0f, 14.6f, 14.3f},
  {15.9f, 15.7f, 15.1f, 13.9f, 12.5f, 11.7f, 12.0f, 13.1f, 14.4f, 15.4f, 15.7f, 15.8f},
  {16.8f, 16.0f, 14.5f, 12.5f, 10.7f, 9.7f, 10.1f, 11.6f, 13.6f, 15.3f, 16.4f, 16.9f},
  {17.2f, 15.8f, 13.5f, 10.9f, 8.6f, 7.5f, 7.9f, 9.7f, 12.3f, 14.8f, 16.7f, 17.5f},
  {17.3f, 15.1f, 12.2f, 8.9f, 6.4f, 5.2f, 5.6f, 7.6f, 10.7f, 13.8f, 16.5f, 17.8f},
  {16.9f, 14.1f, 10.4f, 6.7f, 4.1f, 2.9f, 3.4f, 5.4f, 8.7f, 12.5f, 16.0f, 17.6f},
  {16.5f, 12.6f, 8.3f, 4.3f, 1.8f, 0.9f, 1.3f, 3.1f, 6.5f, 10.8f, 15.1f, 17.5f}};
static const float DG_TAB[45] = {0.69f, 0.71f, 0.73f, 0.76f, 0.78f, 0.80f, 0.83f, 0.85f, 0.88f, 0.91f, 0.94f, 0.97f, 1.00f, 1.03f, 1.06f,
  1.09f, 1.13f, 1.16f, 1.19f, 1.23f, 1.26f, 1.30f, 1.34f, 1.38f, 1.42f, 1.47f, 1.51f, 1.56f, 1.60f, 1.65f, 1.69f, 1.74f, 1.79f, 1.84f,
  1.89f, 1.94f, 1.99f, 2.04f, 2.11f, 2.17f, 2.23f, 2.28f, 2.35f, 2.41f, 2.48f};
static const float EA_TAB[45] = {4.58f, 4.75f, 4.93f, 5.11f, 5.30f, 5.49f, 5.69f, 5.89f, 6.10f, 6.32f, 6.54f, 6.77f, 7.02f, 7.26f, 7.52f,
  7.79f, 8.05f, 8.33f, 8.62f, 8.91f, 9.21f, 9.52f, 9.84f, 10.18f, 10.52f, 10.87f, 11.23f, 11.61f, 11.99f, 12.38f, 12.79f, 13.13f, 13.63f,
  14.08f, 14.53f, 15.00f, 15.49f, 15.97f, 16.47f, 17.53f, 18.68f, 19.80f, 21.10f, 0.0f, 0.0f};
static const float T4_TAB[23] = {11.0f, 11.2f, 11.4f, 11.5f, 11.7f, 11.9f, 12.0f, 12.2f, 12.3f, 12.5f, 12.7f, 12.9f, 13.1f, 13.3f, 13.5f,
  13.7f, 13.9f, 14.0f, 14.2f, 14.4f, 14.6f, 14.8f, 15.0f};

/* E0 (mm/day) of one reservoir-surface cell and day (reservoir.f:425-468).  The saturation-pressure table has
 * 43 entries but the Fortran indexes it up to 45 for 21.5 <= T <= 22 C (out of bounds).  The shipped binary
 * behaves there as if the two missing elements were 0 (that value reproduces its NF1.txt on the evaporation
 * fixture, tests/golden/route_evap14x16.npz), so the table is padded with two zeros. */
static float open_water_e0(float temp, float winds, float rhd, float Ra)
{
  const float nN = 0.6f;
  float wind = winds * 24 * 3600 / 1609, deltagamma, ea, deltaTa4, ed;
  if (temp < 0) deltagamma = 0.69f; else if (temp > 22) deltagamma = 2.48f; else deltagamma = DG_TAB[(int)(temp * 2)];
  if (temp < 0) ea = 4.4f; else if (temp > 22) ea = 21.1f;
  else ea = EA_TAB[(int)(temp * 2)];
  if (temp < -1) deltaTa4 = 11.0f; else if (temp > 21) deltaTa4 = 15.0f; else deltaTa4 = T4_TAB[(int)(temp) + 1];
  ed = ea * rhd / 100;
  return (deltagamma * (0.95f * (0.18f + 0.55f * nN) * Ra - deltaTa4 * (0.56f - 0.09f * sqrtf(ed)) * (0.10f + 0.90f * nN))
          + 0.35f * (0.5f + wind / 100) * (ea - ed)) / (deltagamma + 1);
}
/* column of Ra for the cell's latitude/longitude (reservoir.f:427-435), 0-based */
static int ra_column(float jloc, float iloc)
{
  const int b = (int)(jloc / 10);
  int c;
  if (b > 60) c = 1;
  else if (b <= 60 && (int)(iloc / 10) >= 0) c = 7 - b;
  else if (b < -60) c = 13;
  else c = 7 - b;
  if (c < 1) c = 1;
  if (c > 13) c = 13;
  return c - 1;
}

/* MAKE_CONVOLUTION in full (reservoir.f:216-526): as ro_convolve plus the open-water evaporation of
 * reservoir-surface cells.  potentialevapo(day) lives across the cells of one node: a surface cell met before
 * the node's commissioning test passes keeps the previous cell's values (the Fortran array is only reset for
 * non-surface cells and missing files). */
int ro_convolve_in(const ro_model *m, int ndays, const vr_route_inputs *in, float *flows, float *res_evap)
{
  int n, k, I, J;
  long long last;
  float *RUNO = malloc(sizeof(float) * ndays), *BASE = malloc(sizeof(float) * ndays), *pe = malloc(sizeof(float) * ndays);
  const int has_ev = in->ev_vege && in->ev_soil && in->tr_vege && in->cell_factor;
  const int has_met = in->rh && in->tair && in->wind && in->cell_lat && in->cell_lon;
  int MONTH_OF_YEAR = (int)((ndays % 365) / 30) + in->start_month;
  const int CURRENTYEAR = in->start_year + (int)(ndays % 365);
  if (MONTH_OF_YEAR > 12) MONTH_OF_YEAR = 12;
  for (n = (in->node_end > 0 ? in->node_begin : 0); n < (in->node_end > 0 ? in->node_end : m->nnodes); n++) {
    const onode *nd = &m->node[n];
    float *FLOW = flows + (size_t)n * ndays;
    const int OPEYEAR = in->node_ope_year ? in->node_ope_year[n] : 0;
    float EVA_FACTOR = 0.0f;
    for (I = 0; I < ndays; I++) { FLOW[I] = 0.0f; pe[I] = 0.0f; if (res_evap) res_evap[(size_t)n * ndays + I] = 0.0f; }
    last = -1;
    for (k = 0; k < nd->ncells; k++) {
      int gi = (m->f.nrow - nd->cj[k]) * m->f.ncol + (m->f.ncol - nd->ci[k]);
      const float *us = m->uh_s[n] + (size_t)k * VR_UH_NS;
      const float sc = in->cell_scale ? in->cell_scale[gi] : in->scale;
      const int here = !in->present || in->present[gi];
      if (here) last = gi;
      else for (I = 0; I < ndays; I++) pe[I] = 0.0f;
      for (I = 0; I < ndays; I++) {
        RUNO[I] = here ? in->runoff[(size_t)gi * ndays + I] * sc : 0.0f * sc;
        BASE[I] = here ? in->baseflow[(size_t)gi * ndays + I] * sc : 0.0f * sc;
      }
      if (m->f.reser[IX(&m->f, nd->ci[k], nd->cj[k])] == 9999 && has_met) {
        for (I = 0; I < ndays; I++) {
          if (CURRENTYEAR >= OPEYEAR) {
            const float Ra = RA_TAB[ra_column(in->cell_lat[gi], in->cell_lon[gi])][MONTH_OF_YEAR - 1];
            /* a missing file leaves AIRTEMP/WINDS/RHD holding the previous cell's series (zeros at first) */
            const float temp = last >= 0 ? in->tair[(size_t)last * ndays + I] : 0.0f;
            const float wnd = last >= 0 ? in->wind[(size_t)last * ndays + I] : 0.0f;
            const float rhd = last >= 0 ? in->rh[(size_t)last * ndays + I] : 0.0f;
            pe[I] = open_water_e0(temp, wnd, rhd, Ra);
            EVA_FACTOR = 0.9f;
            if (res_evap) res_evap[(size_t)n * ndays + I] = res_evap[(size_t)n * ndays + I] + pe[I] * EVA_FACTOR;
          }
        }
      }
      else for (I = 0; I < ndays; I++) pe[I] = 0.0f;
      for (I = 1; I <= ndays; I++) {
        for (J = 1; J <= VR_UH_NS; J++)
          if ((I - J + 1) >= 1) FLOW[I - 1] = FLOW[I - 1] + us[J - 1] * (BASE[I - J] + RUNO[I - J]);
        if (has_ev) {
          float evv = 0.0f, evs = 0.0f, trv = 0.0f, FACTOR = in->cell_factor[gi];
          if (last >= 0) { size_t q = (size_t)last * ndays + (I - 1); evv = in->ev_vege[q]; evs = in->ev_soil[q]; trv = in->tr_vege[q]; }
          if (pe[I - 1] * EVA_FACTOR > (evv + evs + trv))
            FLOW[I - 1] = FLOW[I - 1] - pe[I - 1] * FACTOR * 0.028f / 1000 * EVA_FACTOR + (evs + evv + trv) * FACTOR * 0.028f / 1000;
        }
        if (pe[I - 1] < 0.0001f) pe[I - 1] = 0;
      }
    }
  }
  free(RUNO); free(BASE); free(pe);
  return 0;
}

int ro_res_info(const ro_model *m, int node, int *ds_id, int *target)
{
  *ds_id = m->node[node].ds_id; *target = m->node[node].target; return 0;
}
int ro_get_uh_r(const ro_model *m, int node, float *uh_r)
{
  memcpy(uh_r, m->node[node].uh_r, sizeof(float) * VR_UH_NS); return 0;
}

static int cal_month(int CRTDATE)      /* reservoir.f:1688-1716 */
{
  static const int lim[11] = {31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
  int k;
  for (k = 0; k < 11; k++) if (CRTDATE <= lim[k]) return k + 1;
  return 12;
}

#define A2(a, J, I) (a)[(size_t)(J) * (ndays + 2) + (I)]     /* (reservoir J 0-based, day I 1-based) */

/* MAKE_CONVOLUTIONRS steps 3-6 for one station and one parameter set (reservoir.f:1007-1600, STEPBYSTEP
 * false so II == I).  natural: [nnodes][ndays]; res: [nnodes-1]; every local Fortran array starts at zero
 * (static storage).  Outputs as in include/vicres_route.h (any pointer may be NULL). */
int ro_reservoirs(const ro_model *m, int ndays, const float *natural, const vr_reservoir *res, const vr_res_options *opt,
                  float *flow, float *o_vol, float *o_hho, float *o_flowin, float *o_flowout, float *o_turb,
                  float *o_energy, float *o_htk)
{
  const int NORES = m->nnodes, NR = NORES - 1;
  const size_t W = (size_t)ndays + 2;
  float *RESFLOWS = calloc((size_t)NORES * W, sizeof(float)), *VOL = calloc((size_t)(NR + 1) * W, sizeof(float));
  float *FLOWIN = calloc((size_t)(NR + 1) * W, sizeof(float)), *FLOWOUT = calloc((size_t)(NR + 1) * W, sizeof(float));
  float *TURB = calloc((size_t)(NR + 1) * W, sizeof(float)), *HHO = calloc((size_t)(NR + 1) * W, sizeof(float));
  float *ENERGY = calloc((size_t)(NR + 1) * W, sizeof(float)), *HTK = calloc((size_t)(NR + 1) * W, sizeof(float));
  float *H0 = calloc(NR + 1, sizeof(float));
  int I, J, KK;
  for (J = 0; J < NORES; J++) for (I = 1; I <= ndays; I++) A2(RESFLOWS, J, I) = natural[(size_t)J * ndays + (I - 1)];
  for (J = 0; J < NR; J++) {
    const vr_reservoir *r = &res[J];
    if (r->bathymetry == 1) H0[J] = r->bth[0] / (1 + r->bth[3]);
    else {
      float KFACTOR = sqrtf(r->v_dead / r->v_live);
      H0[J] = (KFACTOR * r->hresermax - r->hresermin) / (KFACTOR - 1);
    }
    if (opt->start_year >= r->ope_year) A2(VOL, J, 1) = r->v_initial;
    else A2(VOL, J, 1) = 0.001f;
  }
  for (I = 1; I <= ndays; I++) {
    for (J = 0; J < NR; J++) {
      const vr_reservoir *r = &res[J];
      float HMAX = r->hmax, HMIN = r->hmin, VRESER = 0, REALHEAD = 0, QRESER = 0, Qmin = 0;
      int CURRENTYEAR = opt->start_year + (int)(I / 365), CRTDATE, operating;
      if (r->rule == 1 && HMAX < HMIN) { float t = HMAX; HMAX = HMIN; HMIN = t; }
      operating = CURRENTYEAR >= r->ope_year;
      if (operating) {
        float X1 = r->x1, X2 = r->x2, X3 = r->x3, X4 = r->x4, Demand = r->demand;
        VRESER = r->v_live + r->v_dead; REALHEAD = r->head; QRESER = r->q_design;
        A2(FLOWIN, J, I) = A2(RESFLOWS, J, I);
        Qmin = r->env_flow * r->q_design;
        CRTDATE = (int)(1.0f * (I % 365) + (opt->start_month - 1) * 30);
        if (r->rule == 1 || r->rule == 2) {
          float DESIGNWL, CURRENTWL;
          const float *OP1 = r->op1;
          if (r->rule == 1) {
            if (OP1[0] > OP1[1]) {
              if (CRTDATE > OP1[1] && CRTDATE < OP1[0]) DESIGNWL = (CRTDATE - OP1[1]) / (OP1[0] - OP1[1]) * (HMAX - HMIN);
              else if (CRTDATE >= OP1[0]) DESIGNWL = (HMAX - HMIN) - (CRTDATE - OP1[0]) / (365 - OP1[0] + OP1[1]) * (HMAX - HMIN);
              else DESIGNWL = (HMAX - HMIN) - (CRTDATE + 365 - OP1[0]) / (365 - OP1[0] + OP1[1]) * (HMAX - HMIN);
            }
            else {
              if (CRTDATE > OP1[0] && CRTDATE < OP1[1]) DESIGNWL = (HMAX - HMIN) - (CRTDATE - OP1[0]) / (OP1[1] - OP1[0]) * (HMAX - HMIN);
              else if (CRTDATE >= OP1[1]) DESIGNWL = (HMAX - HMIN) / (365 - OP1[1] + OP1[0]) * (CRTDATE - OP1[1]);
              else DESIGNWL = (HMAX - HMIN) / (365 - OP1[1] + OP1[0]) * (CRTDATE - OP1[0]) + (HMAX - HMIN);
            }
            DESIGNWL = DESIGNWL + (HMIN - H0[J]);
          }
          else {
            DESIGNWL = r->reslv[cal_month(CRTDATE) - 1];
            DESIGNWL = DESIGNWL - H0[J];
          }
          A2(HTK, J, I) = DESIGNWL + H0[J];
          if (r->bathymetry == 1)
            CURRENTWL = r->bth[0] / (expf(-1 * A2(VOL, J, I) / 1000 * r->bth[1]) + A2(VOL, J, I) / 1000 * r->bth[2] + r->bth[3]);
          else CURRENTWL = A2(VOL, J, I) * (r->hresermax - H0[J]) / VRESER;
          if (!opt->legacy_rule_curve) {
            float LRC_WL = HMIN + 0.80f * (A2(HTK, J, I) - HMIN), URC_WL = HMAX - 0.80f * (HMAX - A2(HTK, J, I));
            float LRC_VOL = (LRC_WL - H0[J]) * VRESER / (r->hresermax - H0[J]);
            float URC_VOL = (URC_WL - H0[J]) * VRESER / (r->hresermax - H0[J]);
            if (A2(VOL, J, I) > URC_VOL) {
              A2(VOL, J, I + 1) = URC_VOL;
              A2(FLOWOUT, J, I) = (A2(VOL, J, I) + A2(FLOWIN, J, I) * 24 * 3.6f - URC_VOL) / 24 / 3.6f;
              A2(TURB, J, I) = fminf(fmaxf(A2(FLOWOUT, J, I), Qmin), QRESER);
            }
            else if (A2(VOL, J, I) < LRC_VOL) {
              A2(FLOWOUT, J, I) = Qmin;
              A2(TURB, J, I) = Qmin;
              A2(VOL, J, I + 1) = A2(VOL, J, I) + (A2(FLOWIN, J, I) - A2(TURB, J, I)) * 24 * 3.6f;
            }
            else {
              float RC_SLOPE = (VRESER - r->v_dead) / (OP1[0] - OP1[1]);
              if (CRTDATE > OP1[1] && CRTDATE < OP1[0]) A2(VOL, J, I + 1) = A2(VOL, J, I) + RC_SLOPE;
              else A2(VOL, J, I + 1) = A2(VOL, J, I) - RC_SLOPE;
              A2(FLOWOUT, J, I) = (A2(VOL, J, I) - A2(VOL, J, I + 1)) / 24 / 3.6f + A2(FLOWIN, J, I);
              A2(TURB, J, I) = fminf(fmaxf(A2(FLOWOUT, J, I), Qmin), QRESER);
            }
          }
          else {       /* zone logic of the shipped binary (reservoir.f:1268-1302, commented out today) */
            float TARGET = (DESIGNWL * VRESER) / (r->hresermax - H0[J]);
            if (CURRENTWL >= DESIGNWL) {
              if ((A2(VOL, J, I) + A2(FLOWIN, J, I) * 24 * 3.6f - QRESER * 24 * 3.6f) > TARGET) {
                A2(VOL, J, I + 1) = A2(VOL, J, I) + A2(FLOWIN, J, I) * 24 * 3.6f - QRESER * 24 * 3.6f;
                A2(FLOWOUT, J, I) = QRESER;
                A2(TURB, J, I) = A2(FLOWOUT, J, I);
              }
              else {
                A2(VOL, J, I + 1) = TARGET;
                A2(FLOWOUT, J, I) = (A2(VOL, J, I) - A2(VOL, J, I + 1)) / 24 / 3.6f + A2(FLOWIN, J, I);
                A2(TURB, J, I) = A2(FLOWOUT, J, I);
              }
            }
            else {
              if ((A2(VOL, J, I) + A2(FLOWIN, J, I) * 24 * 3.6f) > TARGET) {
                A2(VOL, J, I + 1) = TARGET;
                A2(FLOWOUT, J, I) = A2(FLOWIN, J, I) - (A2(VOL, J, I + 1) - A2(VOL, J, I)) / 24 / 3.6f;
                A2(TURB, J, I) = A2(FLOWOUT, J, I);
                if (A2(TURB, J, I) > QRESER) {       /* the binary still stores the excess (:1289, commented out today) */
                  A2(VOL, J, I + 1) = A2(VOL, J, I + 1) + (A2(TURB, J, I) - QRESER) * 24 * 3.6f;
                  A2(TURB, J, I) = QRESER;
                }
              }
              else {
                A2(VOL, J, I + 1) = A2(VOL, J, I) + A2(FLOWIN, J, I) * 24 * 3.6f;
                A2(FLOWOUT, J, I) = Qmin;
                A2(TURB, J, I) = A2(FLOWOUT, J, I);
              }
            }
          }
        }
        else if (r->rule == 3 || r->rule == 5) {
          float V = A2(VOL, J, I), FIN = A2(FLOWIN, J, I), FO;
          if (r->rule == 5) {
            int mth = cal_month(CRTDATE) - 1;
            X1 = r->op5x1[mth]; X2 = r->op5x2[mth]; X3 = r->op5x3[mth]; X4 = r->op5x4[mth]; Demand = r->demand5[mth];
          }
          if (V < r->v_dead) FO = 0;
          else if (V <= X2) {
            if ((V - r->v_dead + FIN * 24 * 3.6f) <= (Demand + (V - X2) * tanf(X1) * 24 * 3.6f)) FO = (V - r->v_dead) / 24 / 3.6f + FIN;
            else FO = Demand + (V - X2) * tanf(X1);
          }
          else if (V >= X3) {
            if ((Demand + (V - X3) * tanf(X4)) * 24 * 3.6f > (V) - r->v_dead) FO = (V - r->v_dead) / 24 / 3.6f;
            else if ((Demand + (V - X3) * tanf(X4)) > QRESER) FO = QRESER;
            else FO = Demand + (V - X3) * tanf(X4);
          }
          else {
            if (Demand * 24 * 3.6f > (V - r->v_dead)) FO = (V - r->v_dead) / 24 / 3.6f;
            else FO = Demand;
          }
          A2(FLOWOUT, J, I) = FO;
          A2(TURB, J, I) = FO;
          if (A2(TURB, J, I) > QRESER) A2(TURB, J, I) = QRESER;
          A2(VOL, J, I + 1) = V + (FIN - FO) * 24 * 3.6f;
        }
        else if (r->rule == 4) {
          float OP4 = r->series ? r->series[I - 1] : 0.0f;
          if (OP4 > QRESER && A2(VOL, J, I) < VRESER) OP4 = QRESER;
          if (OP4 * 24 * 3.6f > ((A2(VOL, J, I) - r->v_dead)) + A2(FLOWIN, J, I) * 24 * 3.6f)
            A2(TURB, J, I) = (A2(VOL, J, I) - r->v_dead) / 24 / 3.6f + A2(FLOWIN, J, I);
          else A2(TURB, J, I) = OP4;
          if (A2(TURB, J, I) < 0) A2(TURB, J, I) = 0;
          A2(VOL, J, I + 1) = A2(VOL, J, I) + (A2(FLOWIN, J, I) - A2(TURB, J, I)) * 24 * 3.6f;
          if (A2(TURB, J, I) > QRESER) A2(TURB, J, I) = QRESER;
        }
        else if (r->rule == 6) {
          float OP6 = r->series ? r->series[I - 1] : 0.0f;
          if (A2(FLOWIN, J, I) - (OP6 - A2(VOL, J, I)) / 24 / 3.6f > 0) {
            A2(VOL, J, I + 1) = OP6;
            A2(FLOWOUT, J, I) = A2(FLOWIN, J, I) - (A2(VOL, J, I + 1) - A2(VOL, J, I)) / 24 / 3.6f;
          }
          else {
            A2(FLOWOUT, J, I) = Qmin;
            A2(VOL, J, I + 1) = A2(VOL, J, I) + (A2(FLOWIN, J, I) - A2(FLOWOUT, J, I)) * 24 * 3.6f;
          }
          A2(TURB, J, I) = A2(FLOWOUT, J, I);
          if (A2(TURB, J, I) > QRESER) A2(TURB, J, I) = QRESER;
        }
        /* STEP 5-7 (reservoir.f:1422-1508) */
        if (A2(ENERGY, J, I) < 0) A2(ENERGY, J, I) = 0;
        if (A2(VOL, J, I + 1) > VRESER) {
          if (opt->legacy_rule_curve) {     /* the binary spills the excess; today's source clamps first and adds 0 */
            A2(FLOWOUT, J, I) = A2(TURB, J, I) + (A2(VOL, J, I + 1) - VRESER) / 24 / 3.6f;
            A2(VOL, J, I + 1) = VRESER;
          }
          else {
            A2(VOL, J, I + 1) = VRESER;
            A2(FLOWOUT, J, I) = A2(TURB, J, I) + (A2(VOL, J, I + 1) - VRESER) / 24 / 3.6f;
          }
        }
        else A2(FLOWOUT, J, I) = A2(TURB, J, I);
        if (A2(VOL, J, I) < 0) A2(VOL, J, I) = 0;
        if (A2(FLOWOUT, J, I) < Qmin) A2(FLOWOUT, J, I) = Qmin;
        if (r->hydropower_generator == 0) A2(TURB, J, I) = 0.0f;
        {
          float IRR = r->irrigation ? r->irrigation[I - 1] : 0.0f;
          if (A2(FLOWOUT, J, I) >= IRR) A2(FLOWOUT, J, I) = A2(FLOWOUT, J, I) - IRR;
          else A2(FLOWOUT, J, I) = Qmin;
        }
        if (A2(VOL, J, I + 1) - (r->seepage + r->infiltration) * 24 * 3.6f > 0) {
          A2(VOL, J, I + 1) = A2(VOL, J, I + 1) - (r->seepage + r->infiltration) * 24 * 3.6f;
          A2(FLOWOUT, J, I) = A2(FLOWOUT, J, I) + r->seepage;
        }
        else A2(VOL, J, I + 1) = 0;
        A2(HHO, J, I) = A2(VOL, J, I) / VRESER * (r->hresermax - H0[J]) + H0[J];
        A2(HHO, J, I + 1) = A2(VOL, J, I + 1) / VRESER * (r->hresermax - H0[J]) + H0[J];
        A2(ENERGY, J, I) = 0.9f * 9.81f * A2(TURB, J, I) * ((A2(HHO, J, I) + A2(HHO, J, I + 1)) / 2 - (r->hresermax - REALHEAD)) / 1000;
        if (r->rule == 0) {
          A2(FLOWOUT, J, I) = A2(FLOWIN, J, I);
          if (A2(FLOWOUT, J, I) <= Qmin) A2(FLOWOUT, J, I) = Qmin;
          if (r->hydropower_generator == 1) A2(TURB, J, I) = A2(FLOWOUT, J, I);
          else A2(TURB, J, I) = 0.0f;
          if (A2(TURB, J, I) > QRESER) A2(TURB, J, I) = QRESER;
        }
        if (r->rule == 0 || r->hydropower_type == 0) {
          A2(VOL, J, I + 1) = VRESER;
          A2(ENERGY, J, I) = 0.9f * 9.81f * A2(TURB, J, I);
          A2(ENERGY, J, I) = A2(ENERGY, J, I) * (REALHEAD) / 1000;
          A2(HTK, J, I) = REALHEAD;
          A2(HHO, J, I) = r->hresermax;
          A2(HHO, J, I + 1) = r->hresermax;
        }
      }
      else {      /* label 125: not yet commissioned (reservoir.f:1514-1525) */
        A2(FLOWIN, J, I) = A2(RESFLOWS, J, I);
        A2(FLOWOUT, J, I) = A2(FLOWIN, J, I);
        A2(VOL, J, I) = 0; A2(ENERGY, J, I) = 0; A2(HTK, J, I) = 0; A2(HHO, J, I) = 0; A2(TURB, J, I) = 0.0f;
      }
      /* STEP 6: release through UH_R into the target node's inflow of day I+1 (reservoir.f:1569-1575) */
      {
        const int T = m->node[J].target;
        const float *UH_RR = m->node[J].uh_r;
        for (KK = 1; KK <= VR_UH_NS; KK++)
          if ((I - KK + 1) >= 1)
            A2(RESFLOWS, T, I + 1) = A2(RESFLOWS, T, I + 1) + UH_RR[KK - 1] * A2(FLOWOUT, J, I - KK + 1);
      }
    }
    if (A2(RESFLOWS, NORES - 1, I) < 0) A2(RESFLOWS, NORES - 1, I) = 0;
    flow[I - 1] = A2(RESFLOWS, NORES - 1, I);
  }
  for (J = 0; J < NR; J++) {
    for (I = 1; I <= ndays + 1; I++) {
      if (o_vol) o_vol[(size_t)J * (ndays + 1) + I - 1] = A2(VOL, J, I);
      if (o_hho) o_hho[(size_t)J * (ndays + 1) + I - 1] = A2(HHO, J, I);
    }
    for (I = 1; I <= ndays; I++) {
      size_t q = (size_t)J * ndays + I - 1;
      if (o_flowin) o_flowin[q] = A2(FLOWIN, J, I);
      if (o_flowout) o_flowout[q] = A2(FLOWOUT, J, I);
      if (o_turb) o_turb[q] = A2(TURB, J, I);
      if (o_energy) o_energy[q] = A2(ENERGY, J, I);
      if (o_htk) o_htk[q] = A2(HTK, J, I);
    }
  }
  free(RESFLOWS); free(VOL); free(FLOWIN); free(FLOWOUT); free(TURB); free(HHO); free(ENERGY); free(HTK); free(H0);
  return 0;
}
