/* route_oracle.c -- ORACLE = TEST INFRASTRUCTURE (never linked into or called from vicres_b200/).
 *
 * Plain-C fp32 restatement of the reference's unit-hydrograph routing for one station:
 *   READ_DIREC              routing/model/reservoir.f:149-211
 *   SEARCH_WHOLECATCHMENT   routing/model/init_routines.f:103-168
 *   SEARCH_CATCHMENTRS      routing/model/reservoir.f:578-763   (cell lists only)
 *   MAKE_UHM                routing/model/unit_hyd_routines.f:7-50
 *   MAKE_GRID_UH            routing/model/unit_hyd_routines.f:55-178
 *   MAKE_CONVOLUTION        routing/model/reservoir.f:216-526   (without open-water evaporation)
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

typedef struct { int res_id, ncells, PI, PJ; int *ci, *cj; } onode;

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
