// vicres_route.cu -- C ABI (include/vicres_route.h) and CUDA kernels of the routing pass:
// unit hydrographs (MAKE_UHM, MAKE_GRID_UH) and the node inflow convolution (MAKE_CONVOLUTION).
//
//   k_uhm        one thread per grid cell: the 48 hourly taps of the Lohmann Green's function
//   k_grid_uh    one thread block per (node, catchment cell): walks the D8 path to the node and
//                convolves the running hourly response FR(2304) with every path cell's UHM(48) in
//                shared memory -- thread t owns hour t and adds its 48 products in the reference's
//                tap order, so every fp32 sum rounds as the serial Fortran loop does; then the
//                24:1 daily fold, the UH_BOX convolution and the normalisation
//   k_convolve   one thread per (node, day): adds the cells of the node's catchment in list order,
//                taps ascending, i.e. exactly the accumulation order of reservoir.f:483-491, which is
//                what makes the fp32 result independent of the parallelisation
//   k_reservoirs one thread block per parameter set, one thread per reservoir: the MAKE_CONVOLUTIONRS day
//                loop (reservoir.f:1142-1600).  Days are sequential; inside a day the reservoirs are
//                independent (a release reaches the next node on day I+1), so phase A = every reservoir
//                operates, barrier, phase B = every node adds its upstream releases through UH_R in
//                reservoir order, taps ascending (the order of reservoir.f:1569-1575), barrier; the station's
//                inflow, which feeds nothing back, is summed after the loop with one thread per day.
//                Transcendentals (tan of the hedging angles, sqrt in H0) are evaluated on the host so
//                the device does only IEEE + - * / and compares: results equal the serial loop bit for bit
// Everything is fp32 like the Fortran REALs and compiled with -fmad=false.  No CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/vicres_route.h"

namespace {

constexpr int LE = VR_UH_LE, UH_DAY = VR_UH_DAY, KE = VR_UH_KE, NS = VR_UH_NS, TMAX = VR_UH_DAY * 24;

// ---- host-side graph preprocessing (READ_DIREC, SEARCH_WHOLECATCHMENT, SEARCH_CATCHMENTRS) ----
struct Graph {
  int ncol = 0, nrow = 0;
  std::vector<int> next;    // file-order index of the downstream cell, -1 = none / leaves the grid
  std::vector<int> reser;
};

// Fortran (I, J) <-> file order (r, c): I = ncol - c, J = nrow - r (arrays are read flipped, reservoir.f:170-173)
inline int gidx(const Graph &g, int I, int J) { return (g.nrow - J) * g.ncol + (g.ncol - I); }

void build_graph(const vr_route_grid *in, Graph &g)
{
  g.ncol = in->ncol; g.nrow = in->nrow;
  const int n = g.ncol * g.nrow;
  g.next.assign(n, -1);
  g.reser.assign(n, 0);
  static const int dI[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1}, dJ[9] = {0, 1, 1, 0, -1, -1, -1, 0, 1};
  for (int r = 0; r < g.nrow; r++)
    for (int c = 0; c < g.ncol; c++) {
      const int k = r * g.ncol + c, h = in->dir[k];
      g.reser[k] = in->reser ? in->reser[k] : 0;
      if (h < 1 || h > 8) continue;
      const int I = g.ncol - c + dI[h], J = g.nrow - r + dJ[h];
      // DIREC components equal to 0 mean "no downstream cell" in the reference's test (reservoir.f:617-618)
      if (I == 0 || J == 0) continue;
      if (I > g.ncol || J > g.nrow || I < 1 || J < 1) { g.next[k] = -2; continue; }   // leaves the grid
      g.next[k] = gidx(g, I, J);
    }
}

// mode 0: whole catchment; 1: reservoir target (FINAL = 0); 2: station remainder (FINAL = 1)
bool reaches(const Graph &g, int k, int target, int mode)
{
  for (int guard = 0; guard <= g.ncol * g.nrow + 4; guard++) {
    if (k < 0) return false;
    if (k == target) return true;
    if (mode == 2 && g.reser[k] > 0) return false;
    if (mode == 1 && g.reser[k] > 0 && g.reser[k] != 9999) return false;
    k = g.next[k];
  }
  return false;
}

struct Node { int res_id = 0, cell = -1, ds_id = 0, target = -1; std::vector<int> cells; };

// The three searches of reaches() for every cell at once, one downstream walk per cell with memoisation (the
// reference walks the path of every cell again for every node: SEARCH_WHOLECATCHMENT / SEARCH_CATCHMENTRS,
// reservoir.f:578-763):
//   whole[k]    mode 0: the path from k reaches the station
//   first_res[k] mode 1: the first real reservoir cell (reser > 0, != 9999) on the path from k, k included, or -1 --
//               k is in the list of the reservoir at cell T exactly when first_res[k] == T
//   direct[k]   mode 2: the path reaches the station before any cell with reser > 0 (k included; the station cell itself
//               counts as reached)
// A path that runs into a cycle or off the grid reaches nothing, like the guard of reaches().
struct CatchmentMaps { std::vector<char> whole, direct; std::vector<int> first_res; };
void catchment_maps(const Graph &g, int station, CatchmentMaps &m)
{
  const int n = g.ncol * g.nrow;
  m.whole.assign(n, 0); m.direct.assign(n, 0); m.first_res.assign(n, -1);
  std::vector<char> state(n, 0);                     // 0 new, 1 on the current path, 2 done
  std::vector<int> path;
  for (int s = 0; s < n; s++) {
    if (state[s]) continue;
    path.clear();
    int k = s;
    while (k >= 0 && state[k] == 0) { state[k] = 1; path.push_back(k); k = g.next[k]; }
    // k: off the grid (< 0), a finished cell, or a cell of the current path (the walk closed a cycle)
    char w = 0, d = 0;
    int fr = -1;
    if (k >= 0 && state[k] == 2) { w = m.whole[k]; d = m.direct[k]; fr = m.first_res[k]; }
    auto visit = [&](int c) {
      const bool real = g.reser[c] > 0 && g.reser[c] != 9999;
      if (c == station) { w = 1; d = 1; }
      else if (g.reser[c] > 0) d = 0;
      if (real) fr = c;
      m.whole[c] = w; m.direct[c] = d; m.first_res[c] = fr;
      state[c] = 2;
    };
    size_t tail = path.size();
    if (k >= 0 && state[k] == 1) {
      // a walk from a cell of a cycle goes round it until it meets its target (reaches() allows more steps than there are
      // cells): every cycle cell sees the nearest event ahead of it, wrapped -- two backward sweeps over the cycle
      size_t j = 0;
      while (path[j] != k) j++;
      for (int sweep = 0; sweep < 2; sweep++)
        for (size_t i = path.size(); i-- > j;) visit(path[i]);
      tail = j;
    }
    for (size_t i = tail; i-- > 0;) visit(path[i]);
  }
}

void list_cells(const Graph &g, int target, int mode, std::vector<int> &out)
{
  out.clear();
  for (int I = 1; I <= g.ncol; I++)          // the reference scans I outer, J inner
    for (int J = 1; J <= g.nrow; J++) {
      const int k = gidx(g, I, J);
      if (reaches(g, k, target, mode)) out.push_back(k);
    }
}

// ---- kernels ----
__global__ void k_uhm(int ncell, const float *__restrict__ velo, const float *__restrict__ diff,
                      const float *__restrict__ xmask, float *__restrict__ uhm)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncell) return;
  const float PI_ = 4.0f * atanf(1.0f), DT = 3600.0f, v = velo[k], d = diff[k], x = xmask[k];
  float T = 0.0f, inte = 0.0f, h[LE];
  for (int K = 0; K < LE; K++) {
    T = T + DT;
    float H = 0.0f;
    if (v > 0.0f) {
      const float POT = powf((v * T - x), 2.0f) / (4.0f * d * T);
      if (!(POT > 69.0f)) H = 1.0f / (2.0f * sqrtf(PI_ * d)) * x / powf(T, 1.5f) * expf(-POT);
    }
    h[K] = H;
  }
  for (int K = 0; K < LE; K++) inte = inte + h[K];
  for (int K = 0; K < LE; K++) uhm[(size_t)k * LE + K] = (inte > 0.0f) ? h[K] / inte : h[K];
}

// one block per catchment cell (of one node); uh_s[cell_in_list][NS]
__global__ void __launch_bounds__(256) k_grid_uh(const int *__restrict__ cells, int node_cell, const int *__restrict__ next,
                                                 const float *__restrict__ uhm, const float *__restrict__ uh_box,
                                                 float *__restrict__ uh_s)
{
  __shared__ float fr1[TMAX + 1], fr2[TMAX + 1], u[LE], daily[UH_DAY + 1], us[NS];
  const int tid = threadIdx.x;
  int k = cells[blockIdx.x];
  for (int t = tid + 1; t <= TMAX; t += blockDim.x) { fr1[t] = (t <= 24) ? 1.0f / 24.0f : 0.0f; fr2[t] = 0.0f; }
  __syncthreads();
  while (k != node_cell && k >= 0) {
    if (tid < LE) u[tid] = uhm[(size_t)k * LE + tid];
    __syncthreads();
    for (int t = tid + 1; t <= TMAX; t += blockDim.x) {
      float acc = fr2[t];                       // == 0
      const int lmax = (t - 1 < LE) ? t - 1 : LE;
      for (int L = 1; L <= lmax; L++) acc = acc + fr1[t - L] * u[L - 1];
      fr2[t] = acc;
    }
    __syncthreads();
    for (int t = tid + 1; t <= TMAX; t += blockDim.x) { fr1[t] = fr2[t]; fr2[t] = 0.0f; }
    k = next[k];
    __syncthreads();
  }
  if (tid < UH_DAY) {
    float s = 0.0f;
    for (int h = 1; h <= 24; h++) s = s + fr1[tid * 24 + h];      // TT = (T+23)/24, hours ascending
    daily[tid + 1] = s;
  }
  __syncthreads();
  if (tid < NS) {
    float s = 0.0f;                                               // UH_S(m) += UH_BOX(K)*UH_DAILY(U), K ascending
    const int m = tid + 1;
    for (int K = 1; K <= KE; K++) {
      const int U = m - K + 1;
      if (U >= 1 && U <= UH_DAY) s = s + uh_box[K - 1] * daily[U];
    }
    us[tid] = s;
  }
  __syncthreads();
  __shared__ float total;
  if (tid == 0) {
    float s = 0.0f;
    for (int K = 0; K < NS; K++) s = s + us[K];
    total = s;
  }
  __syncthreads();
  if (tid < NS) uh_s[(size_t)blockIdx.x * NS + tid] = us[tid] / total;
}

/* Penman (1948) tables of MAKE_CONVOLUTION (reservoir.f:271-301): extraterrestrial radiation Ra (mm/day) by
 * month (rows) and 10-degree latitude band 60N..60S (13 columns), delta/gamma and saturation vapour
 * pressure by half degree 0..22 C, sigma*Ta^4 by degree -1..21 C */
__constant__ float RA_TAB[13][12] = {
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
__constant__ float DG_TAB[45] = {0.69f, 0.71f, 0.73f, 0.76f, 0.78f, 0.80f, 0.83f, 0.85f, 0.88f, 0.91f, 0.94f, 0.97f, 1.00f, 1.03f, 1.06f,
  1.09f, 1.13f, 1.16f, 1.19f, 1.23f, 1.26f, 1.30f, 1.34f, 1.38f, 1.42f, 1.47f, 1.51f, 1.56f, 1.60f, 1.65f, 1.69f, 1.74f, 1.79f, 1.84f,
  1.89f, 1.94f, 1.99f, 2.04f, 2.11f, 2.17f, 2.23f, 2.28f, 2.35f, 2.41f, 2.48f};
__constant__ float EA_TAB[45] = {4.58f, 4.75f, 4.93f, 5.11f, 5.30f, 5.49f, 5.69f, 5.89f, 6.10f, 6.32f, 6.54f, 6.77f, 7.02f, 7.26f, 7.52f,
  7.79f, 8.05f, 8.33f, 8.62f, 8.91f, 9.21f, 9.52f, 9.84f, 10.18f, 10.52f, 10.87f, 11.23f, 11.61f, 11.99f, 12.38f, 12.79f, 13.13f, 13.63f,
  14.08f, 14.53f, 15.00f, 15.49f, 15.97f, 16.47f, 17.53f, 18.68f, 19.80f, 21.10f, 0.0f, 0.0f};
__constant__ float T4_TAB[23] = {11.0f, 11.2f, 11.4f, 11.5f, 11.7f, 11.9f, 12.0f, 12.2f, 12.3f, 12.5f, 12.7f, 12.9f, 13.1f, 13.3f, 13.5f,
  13.7f, 13.9f, 14.0f, 14.2f, 14.4f, 14.6f, 14.8f, 15.0f};

/* E0 (mm/day) of one reservoir-surface cell and day (reservoir.f:425-468).  The saturation-pressure table has
 * 43 entries but the Fortran indexes it up to 45 for 21.5 <= T <= 22 C (out of bounds).  The shipped binary
 * behaves there as if the two missing elements were 0 (that value reproduces its NF1.txt on the evaporation
 * fixture, tests/golden/route_evap14x16.npz), so the table is padded with two zeros.  sqrtf is IEEE on the device. */
__device__ float open_water_e0(float temp, float winds, float rhd, float Ra)
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
__device__ int ra_column(float jloc, float iloc)
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


struct ConvArgs {
  int nnodes, ndays;
  const int64_t *node_ptr;      // [nnodes+1] into cells / uh rows
  const int *cells;             // grid index per list entry
  const float *uh_s;            // [entries][NS]
  const float *runoff, *baseflow, *ev_vege, *ev_soil, *tr_vege;   // [ncell][ndays]
  const float *cell_scale, *cell_factor;                          // [ncell] or null
  const unsigned char *present;                                   // [ncell] or null
  float scale;
  float *flows;                 // [nnodes][ndays]
  // open-water evaporation of reservoir-surface cells (null rh: none)
  const float *rh, *tair, *wind, *cell_lat, *cell_lon;
  const int *reser, *node_ope_year;
  int month_of_year, current_year;
  int node_begin;               // first node of this launch (grid.y counts from it)
  float *res_evap;              // [nnodes][ndays] or null
};

// One thread per (node, day): cells in list order, taps ascending.  potentialevapo(day), EVA_FACTOR and the
// "arrays still hold the previous cell's file" state of the Fortran are per-day scalars here.
__global__ void __launch_bounds__(128) k_convolve(ConvArgs a)
{
  const int node = a.node_begin + blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;   // i = day - 1
  if (i >= a.ndays) return;
  float flow = 0.0f, pe = 0.0f, eva_factor = 0.0f, resev = 0.0f;
  long long last = -1;
  const bool evap = a.ev_vege && a.ev_soil && a.tr_vege && a.cell_factor;
  const int ope = a.node_ope_year ? a.node_ope_year[node] : 0;
  for (int64_t e = a.node_ptr[node]; e < a.node_ptr[node + 1]; e++) {
    const int gi = a.cells[e];
    const bool here = !a.present || a.present[gi];
    if (here) last = gi;
    if (here) {
      const float sc = a.cell_scale ? a.cell_scale[gi] : a.scale;
      const float *__restrict__ ro = a.runoff + (size_t)gi * a.ndays, *__restrict__ ba = a.baseflow + (size_t)gi * a.ndays;
      const float *__restrict__ us = a.uh_s + (size_t)e * NS;
      const int jmax = (i + 1 < NS) ? i + 1 : NS;
      for (int J = 1; J <= jmax; J++) flow = flow + us[J - 1] * (ba[i - J + 1] * sc + ro[i - J + 1] * sc);
    }
    else pe = 0.0f;
    // a missing cell contributes UH*(0*sc + 0*sc) = +0 to every day: nothing to add
    if (a.rh && a.reser[gi] == 9999) {
      if (a.current_year >= ope) {
        const size_t q = (size_t)(last >= 0 ? last : 0) * a.ndays + i;
        const float temp = last >= 0 ? a.tair[q] : 0.0f, wnd = last >= 0 ? a.wind[q] : 0.0f, rhd = last >= 0 ? a.rh[q] : 0.0f;
        pe = open_water_e0(temp, wnd, rhd, RA_TAB[ra_column(a.cell_lat[gi], a.cell_lon[gi])][a.month_of_year - 1]);
        eva_factor = 0.9f;
        resev = resev + pe * eva_factor;
      }
    }
    else pe = 0.0f;
    if (evap) {
      float evv = 0.0f, evs = 0.0f, trv = 0.0f;
      if (last >= 0) { const size_t q = (size_t)last * a.ndays + i; evv = a.ev_vege[q]; evs = a.ev_soil[q]; trv = a.tr_vege[q]; }
      const float FACTOR = a.cell_factor[gi];
      if (pe * eva_factor > (evv + evs + trv))
        flow = flow - pe * FACTOR * 0.028f / 1000 * eva_factor + (evs + evv + trv) * FACTOR * 0.028f / 1000;
    }
    if (pe < 0.0001f) pe = 0.0f;
  }
  a.flows[(size_t)node * a.ndays + i] = flow;
  if (a.res_evap) a.res_evap[(size_t)node * a.ndays + i] = resev;
}

// ---- reservoirs ------------------------------------------------------------------------------------
struct ResC {            // one reservoir of one parameter set, host-prepared
  float hresermax, vreser, v_dead, head, q_design, qmin, vol1, seepage, loss, H0, span;   // span = HRESERMAX - H0
  float hmax, hmin, op1a, op1b, reslv[12];
  float demand, x2, x3, tan1, tan4;
  float demand5[12], x2_5[12], x3_5[12], tan1_5[12], tan4_5[12];
  float bth[4];
  int ope_year, rule, hp_type, hp_gen, bathymetry;
  long long series, irrigation;        // offsets into the series pool, -1 = none
};

struct ResArgs {
  int nnodes, ndays, start_year, start_month, legacy;
  int day0, day1;                      // days [day0, day1] (1-based) are operated; earlier days only replay the routing of
                                       // their releases from the series already in the buffers (STEPBYSTEP restart)
  const ResC *res;                     // [nsets][nres]
  const float *pool;                   // rule 4/6 and irrigation series
  const float *natural;                // [nnodes][ndays]
  const float *uh_r;                   // [nres][NS]
  const int *target;                   // [nres]
  const int *up_ptr, *up_idx;          // CSR: upstream reservoirs of every node, ascending
  float *resflows;                     // [nsets][nnodes][ndays+2]   (index = day, 1-based)
  float *vol, *hho;                    // [nsets][nres][ndays+1]
  float *flowin, *flowout, *turb, *energy, *htk;   // [nsets][nres][ndays]
  float *flow;                         // [nsets][ndays]
};

__device__ __forceinline__ int cal_month(int d)     // reservoir.f:1688-1716
{
  const int lim[11] = {31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
  int m = 12;
#pragma unroll
  for (int k = 10; k >= 0; k--) if (d <= lim[k]) m = k + 1;
  return m;
}

__global__ void __launch_bounds__(1024) k_reservoirs(ResArgs a)
{
  const int set = blockIdx.x, nres = a.nnodes - 1, nd = a.ndays;
  const size_t W = (size_t)nd + 2;
  float *RF = a.resflows + (size_t)set * a.nnodes * W;
  for (size_t q = threadIdx.x; q < (size_t)a.nnodes * W; q += blockDim.x) {
    const int n = (int)(q / W), I = (int)(q % W);
    RF[q] = (I >= 1 && I <= nd) ? a.natural[(size_t)n * nd + I - 1] : 0.0f;
  }
  __syncthreads();
  for (int I = 1; I <= a.day1; I++) {
    // ---- phase A: operate every reservoir on day I ----
    for (int J = threadIdx.x; I >= a.day0 && J < nres; J += blockDim.x) {
      const ResC &r = a.res[(size_t)set * nres + J];
      const size_t b1 = ((size_t)set * nres + J) * (nd + 1), b0 = ((size_t)set * nres + J) * nd;
      float vol = (I == 1) ? r.vol1 : a.vol[b1 + I - 1];      // (a restart at day0 > 1 finds VOL(day0) in the buffer)
      float vnext = 0.0f, flowin, flowout = 0.0f, turb = 0.0f, energy = 0.0f, htk = 0.0f, hho0 = 0.0f, hho1 = 0.0f;
      const int year = a.start_year + I / 365;
      bool write_h1 = false;
      if (year >= r.ope_year) {
        const float VRESER = r.vreser, QRESER = r.q_design, Qmin = r.qmin;
        flowin = RF[(size_t)J * W + I];
        const int CRTDATE = (int)(1.0f * (float)(I % 365) + (float)((a.start_month - 1) * 30));
        if (r.rule == 1 || r.rule == 2) {
          float DESIGNWL;
          const float fd = (float)CRTDATE, dH = r.hmax - r.hmin;
          if (r.rule == 1) {
            if (r.op1a > r.op1b) {
              if (fd > r.op1b && fd < r.op1a) DESIGNWL = (fd - r.op1b) / (r.op1a - r.op1b) * dH;
              else if (fd >= r.op1a) DESIGNWL = dH - (fd - r.op1a) / (365.0f - r.op1a + r.op1b) * dH;
              else DESIGNWL = dH - ((float)(CRTDATE + 365) - r.op1a) / (365.0f - r.op1a + r.op1b) * dH;
            }
            else {
              if (fd > r.op1a && fd < r.op1b) DESIGNWL = dH - (fd - r.op1a) / (r.op1b - r.op1a) * dH;
              else if (fd >= r.op1b) DESIGNWL = dH / (365.0f - r.op1b + r.op1a) * (fd - r.op1b);
              else DESIGNWL = dH / (365.0f - r.op1b + r.op1a) * (fd - r.op1a) + dH;
            }
            DESIGNWL = DESIGNWL + (r.hmin - r.H0);
          }
          else DESIGNWL = r.reslv[cal_month(CRTDATE) - 1] - r.H0;
          htk = DESIGNWL + r.H0;
          if (!a.legacy) {
            const float LRC_WL = r.hmin + 0.80f * (htk - r.hmin), URC_WL = r.hmax - 0.80f * (r.hmax - htk);
            const float LRC_VOL = (LRC_WL - r.H0) * VRESER / r.span, URC_VOL = (URC_WL - r.H0) * VRESER / r.span;
            if (vol > URC_VOL) {
              vnext = URC_VOL;
              flowout = (vol + flowin * 24.0f * 3.6f - URC_VOL) / 24.0f / 3.6f;
              turb = fminf(fmaxf(flowout, Qmin), QRESER);
            }
            else if (vol < LRC_VOL) {
              flowout = Qmin; turb = Qmin;
              vnext = vol + (flowin - turb) * 24.0f * 3.6f;
            }
            else {
              const float RC_SLOPE = (VRESER - r.v_dead) / (r.op1a - r.op1b);
              vnext = (fd > r.op1b && fd < r.op1a) ? vol + RC_SLOPE : vol - RC_SLOPE;
              flowout = (vol - vnext) / 24.0f / 3.6f + flowin;
              turb = fminf(fmaxf(flowout, Qmin), QRESER);
            }
          }
          else {
            float CURRENTWL;
            if (r.bathymetry == 1) CURRENTWL = r.bth[0] / (expf(-1.0f * vol / 1000.0f * r.bth[1]) + vol / 1000.0f * r.bth[2] + r.bth[3]);
            else CURRENTWL = vol * r.span / VRESER;
            const float TARGET = (DESIGNWL * VRESER) / r.span;
            if (CURRENTWL >= DESIGNWL) {
              if ((vol + flowin * 24.0f * 3.6f - QRESER * 24.0f * 3.6f) > TARGET) {
                vnext = vol + flowin * 24.0f * 3.6f - QRESER * 24.0f * 3.6f;
                flowout = QRESER; turb = flowout;
              }
              else {
                vnext = TARGET;
                flowout = (vol - vnext) / 24.0f / 3.6f + flowin; turb = flowout;
              }
            }
            else {
              if ((vol + flowin * 24.0f * 3.6f) > TARGET) {
                vnext = TARGET;
                flowout = flowin - (vnext - vol) / 24.0f / 3.6f; turb = flowout;
                if (turb > QRESER) { vnext = vnext + (turb - QRESER) * 24.0f * 3.6f; turb = QRESER; }
              }
              else {
                vnext = vol + flowin * 24.0f * 3.6f;
                flowout = Qmin; turb = flowout;
              }
            }
          }
        }
        else if (r.rule == 3 || r.rule == 5) {
          float X2 = r.x2, X3 = r.x3, T1 = r.tan1, T4 = r.tan4, Demand = r.demand;
          if (r.rule == 5) {
            const int m = cal_month(CRTDATE) - 1;
            X2 = r.x2_5[m]; X3 = r.x3_5[m]; T1 = r.tan1_5[m]; T4 = r.tan4_5[m]; Demand = r.demand5[m];
          }
          if (vol < r.v_dead) flowout = 0.0f;
          else if (vol <= X2) {
            if ((vol - r.v_dead + flowin * 24.0f * 3.6f) <= (Demand + (vol - X2) * T1 * 24.0f * 3.6f)) flowout = (vol - r.v_dead) / 24.0f / 3.6f + flowin;
            else flowout = Demand + (vol - X2) * T1;
          }
          else if (vol >= X3) {
            if ((Demand + (vol - X3) * T4) * 24.0f * 3.6f > vol - r.v_dead) flowout = (vol - r.v_dead) / 24.0f / 3.6f;
            else if ((Demand + (vol - X3) * T4) > QRESER) flowout = QRESER;
            else flowout = Demand + (vol - X3) * T4;
          }
          else {
            if (Demand * 24.0f * 3.6f > (vol - r.v_dead)) flowout = (vol - r.v_dead) / 24.0f / 3.6f;
            else flowout = Demand;
          }
          turb = flowout;
          if (turb > QRESER) turb = QRESER;
          vnext = vol + (flowin - flowout) * 24.0f * 3.6f;
        }
        else if (r.rule == 4) {
          float OP4 = r.series >= 0 ? a.pool[r.series + I - 1] : 0.0f;
          if (OP4 > QRESER && vol < VRESER) OP4 = QRESER;
          if (OP4 * 24.0f * 3.6f > (vol - r.v_dead) + flowin * 24.0f * 3.6f) turb = (vol - r.v_dead) / 24.0f / 3.6f + flowin;
          else turb = OP4;
          if (turb < 0.0f) turb = 0.0f;
          vnext = vol + (flowin - turb) * 24.0f * 3.6f;
          if (turb > QRESER) turb = QRESER;
        }
        else if (r.rule == 6) {
          const float OP6 = r.series >= 0 ? a.pool[r.series + I - 1] : 0.0f;
          if (flowin - (OP6 - vol) / 24.0f / 3.6f > 0.0f) {
            vnext = OP6;
            flowout = flowin - (vnext - vol) / 24.0f / 3.6f;
          }
          else {
            flowout = Qmin;
            vnext = vol + (flowin - flowout) * 24.0f * 3.6f;
          }
          turb = flowout;
          if (turb > QRESER) turb = QRESER;
        }
        // step 5-7: spill, environmental flow, irrigation, seepage, head, energy (reservoir.f:1422-1508)
        if (vnext > VRESER) {
          if (a.legacy) { flowout = turb + (vnext - VRESER) / 24.0f / 3.6f; vnext = VRESER; }
          else { vnext = VRESER; flowout = turb + (vnext - VRESER) / 24.0f / 3.6f; }
        }
        else flowout = turb;
        if (vol < 0.0f) vol = 0.0f;
        if (flowout < Qmin) flowout = Qmin;
        if (r.hp_gen == 0) turb = 0.0f;
        {
          const float IRR = r.irrigation >= 0 ? a.pool[r.irrigation + I - 1] : 0.0f;
          if (flowout >= IRR) flowout = flowout - IRR;
          else flowout = Qmin;
        }
        if (vnext - r.loss * 24.0f * 3.6f > 0.0f) {
          vnext = vnext - r.loss * 24.0f * 3.6f;
          flowout = flowout + r.seepage;
        }
        else vnext = 0.0f;
        hho0 = vol / VRESER * r.span + r.H0;
        hho1 = vnext / VRESER * r.span + r.H0;
        energy = 0.9f * 9.81f * turb * ((hho0 + hho1) / 2.0f - (r.hresermax - r.head)) / 1000.0f;
        if (r.rule == 0) {
          flowout = flowin;
          if (flowout <= Qmin) flowout = Qmin;
          turb = (r.hp_gen == 1) ? flowout : 0.0f;
          if (turb > QRESER) turb = QRESER;
        }
        if (r.rule == 0 || r.hp_type == 0) {
          vnext = VRESER;
          energy = 0.9f * 9.81f * turb;
          energy = energy * r.head / 1000.0f;
          htk = r.head;
          hho0 = r.hresermax; hho1 = r.hresermax;
        }
        write_h1 = true;
      }
      else {                                            // label 125: not commissioned yet
        flowin = RF[(size_t)J * W + I];
        flowout = flowin;
        vol = 0.0f;
      }
      a.vol[b1 + I - 1] = vol;
      a.vol[b1 + I] = vnext;
      a.hho[b1 + I - 1] = hho0;
      if (write_h1 || I == nd) a.hho[b1 + I] = write_h1 ? hho1 : 0.0f;
      a.flowin[b0 + I - 1] = flowin; a.flowout[b0 + I - 1] = flowout; a.turb[b0 + I - 1] = turb;
      a.energy[b0 + I - 1] = energy; a.htk[b0 + I - 1] = htk;
    }
    __syncthreads();
    // ---- phase B: releases of day <= I reach the next RESERVOIR on day I+1 through UH_R (the station's
    //      inflow feeds nothing back, so it is summed after the day loop, all days in parallel) ----
    // One WARP per target: the lanes fetch and multiply 32 terms at a time, then every lane replays the serial
    // chain acc = acc + t_i through shuffles (same order, same roundings as the one-thread loop; the loads and
    // multiplies leave the dependent path).
    {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
      const int kmax = I < NS ? I : NS;
      for (int k = warp; k < nres; k += nwarps) {
        const int u0 = a.up_ptr[k], u1 = a.up_ptr[k + 1];
        if (u0 == u1) continue;
        float acc = RF[(size_t)k * W + I + 1];
        const int n = (u1 - u0) * kmax;
        // this lane's position in the chain, (upstream reservoir ju, tap kk), advanced by 32 terms per group
        int ju = lane / kmax, kk = lane % kmax + 1;
        auto term = [&]() -> float {
          float v = 0.0f;
          if (u0 + ju < u1) {
            const int J = a.up_idx[u0 + ju];
            v = a.uh_r[(size_t)J * NS + kk - 1] * a.flowout[((size_t)set * nres + J) * nd + I - kk];
          }
          kk += 32;
          while (kk > kmax) { kk -= kmax; ju++; }
          return v;
        };
        float t = term();
        for (int base = 0; base < n; base += 32) {
          const float tn = term();                   // next group's loads fly while this group's adds run
          float tv[32];
#pragma unroll
          for (int i = 0; i < 32; i++) tv[i] = __shfl_sync(0xffffffffu, t, i);     // 32 independent shuffles first
#pragma unroll
          for (int i = 0; i < 32; i++) acc = acc + tv[i];   // then the dependent chain of adds; past the end of the
          t = tn;                                            // chain the terms are +0.0f, which adds exactly nothing
        }
        if (lane == 0) RF[(size_t)k * W + I + 1] = acc;
      }
    }
    __syncthreads();
  }
  // ---- station: RESFLOWS(NORES, I) = naturalised inflow + the releases routed to it during day I-1, added in the
  //      reference's order (reservoirs ascending, taps ascending: reservoir.f:1569-1575), then clamped (:1589-1593).
  //      One thread per day: the serial chain of up to (reservoirs x 107) fp32 adds runs for all days at once.
  {
    const int k = nres, u0 = a.up_ptr[k], u1 = a.up_ptr[k + 1];
    for (int I = 1 + threadIdx.x; I <= a.day1; I += blockDim.x) {
      float acc = RF[(size_t)k * W + I];
      if (I >= 2)
        for (int u = u0; u < u1; u++) {
          const int J = a.up_idx[u];
          const float *uh = a.uh_r + (size_t)J * NS;
          const float *fo = a.flowout + ((size_t)set * nres + J) * nd;
          const int kmax = (I - 1) < NS ? (I - 1) : NS;
          for (int KK = 1; KK <= kmax; KK++) acc = acc + uh[KK - 1] * fo[I - 1 - KK];
        }
      if (acc < 0.0f) acc = 0.0f;
      a.flow[(size_t)set * nd + I - 1] = acc;
    }
  }
}

// ---- calibration objectives (toolbox/calibration/basin_calibration_eNSGAII.py:86-175) ----------------
// One block per parameter set; fp64 tree reductions over the days.  out[set][4] = NSE, TRMSE, MSDE, ROCE.
__device__ double block_sum(double v, double *sh)
{
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (t < s) sh[t] += sh[t + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) k_objectives(int ndays, const float *__restrict__ sim, const double *__restrict__ obs,
                                                    int handle_neg, double *__restrict__ out)
{
  __shared__ double sh[256];
  const float *x = sim + (size_t)blockIdx.x * ndays;
  double so = 0;
  for (int i = threadIdx.x; i < ndays; i += blockDim.x) so += obs[i];
  const double sum_obs = block_sum(so, sh), obs_mean = sum_obs / ndays;
  double num = 0, den = 0, tr = 0, ms = 0, ss = 0;
  for (int i = threadIdx.x; i < ndays; i += blockDim.x) {
    const double o = obs[i], v = handle_neg ? fabs((double)x[i]) : (double)x[i], oa = handle_neg ? fabs(o) : o;
    num += (v - o) * (v - o);
    den += (o - obs_mean) * (o - obs_mean);
    const double zs = (pow(v + 1, 0.3) - 1) / 0.3, zo = (pow(oa + 1, 0.3) - 1) / 0.3;
    tr += (zs - zo) * (zs - zo);
    ss += v;
    if (i + 1 < ndays) {
      const double v1 = handle_neg ? fabs((double)x[i + 1]) : (double)x[i + 1];
      const double dd = (obs[i + 1] - o) - (v1 - v);
      ms += dd * dd;
    }
  }
  num = block_sum(num, sh); den = block_sum(den, sh); tr = block_sum(tr, sh); ms = block_sum(ms, sh); ss = block_sum(ss, sh);
  if (threadIdx.x == 0) {
    double *r = out + (size_t)blockIdx.x * 4;
    r[0] = 1 - num / den;
    r[1] = sqrt(tr / ndays);
    r[2] = (ndays > 1) ? ms / (ndays - 1) : nan("");
    r[3] = (sum_obs == 0) ? ((ss != 0) ? INFINITY : nan("")) : fabs(ss - sum_obs) / sum_obs;
  }
}

template <class T>
struct DB {
  T *p = nullptr; size_t n = 0;
  cudaError_t put(const T *src, size_t count)
  {
    cudaError_t e = res(count);
    if (e != cudaSuccess || !count) return e;
    return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyDefault);     // host or device source (unified addressing)
  }
  cudaError_t res(size_t count)
  {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void rel() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct vr_route {
  int device = 0;
  std::string err;
  Graph g;
  std::vector<Node> nodes;
  std::vector<int64_t> node_ptr;
  std::vector<int> cells;
  DB<int> d_next, d_cells;
  DB<int64_t> d_node_ptr;
  DB<float> d_velo, d_diff, d_xmask, d_uhm, d_box, d_uh_s, d_ro, d_ba, d_evv, d_evs, d_trv, d_scale, d_factor, d_flows;
  DB<unsigned char> d_present;
  bool has_surface = false;            // reservoir-surface cells (9999) inside a catchment
  DB<float> d_rh, d_tair, d_wind, d_lat, d_lon, d_resev;
  DB<int> d_reser, d_ope;
  std::vector<float> uh_r;             // [nres][NS]
  DB<float> d_uh_r, d_pool, d_natural, d_resflows, d_vol, d_hho, d_flowin, d_flowout, d_turb, d_energy, d_htk, d_flow;
  DB<int> d_target, d_up_ptr, d_up_idx;
  DB<ResC> d_res;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
};

static std::string g_route_error;

#define RK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                   \
      return VR_ERR_CUDA;                                                                            \
    }                                                                                                \
  } while (0)

extern "C" {

const char *vr_route_last_error(const vr_route *h) { return h ? h->err.c_str() : g_route_error.c_str(); }

void vr_route_destroy(vr_route *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  h->d_next.rel(); h->d_cells.rel(); h->d_node_ptr.rel(); h->d_velo.rel(); h->d_diff.rel(); h->d_xmask.rel();
  h->d_uhm.rel(); h->d_box.rel(); h->d_uh_s.rel(); h->d_ro.rel(); h->d_ba.rel(); h->d_evv.rel(); h->d_evs.rel();
  h->d_trv.rel(); h->d_scale.rel(); h->d_factor.rel(); h->d_flows.rel(); h->d_present.rel();
  h->d_rh.rel(); h->d_tair.rel(); h->d_wind.rel(); h->d_lat.rel(); h->d_lon.rel(); h->d_resev.rel(); h->d_reser.rel(); h->d_ope.rel();
  h->d_uh_r.rel(); h->d_pool.rel(); h->d_natural.rel(); h->d_resflows.rel(); h->d_vol.rel(); h->d_hho.rel();
  h->d_flowin.rel(); h->d_flowout.rel(); h->d_turb.rel(); h->d_energy.rel(); h->d_htk.rel(); h->d_flow.rel();
  h->d_target.rel(); h->d_up_ptr.rel(); h->d_up_idx.rel(); h->d_res.rel();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
}

static int route_build(vr_route *h, const vr_route_grid *g, const float *uh_box)
{
  const int ncell = g->ncol * g->nrow;
  RK(h->d_next.put(h->g.next.data(), ncell));
  RK(h->d_velo.put(g->velo, ncell)); RK(h->d_diff.put(g->diff, ncell)); RK(h->d_xmask.put(g->xmask, ncell));
  RK(h->d_box.put(uh_box, KE));
  RK(h->d_uhm.res((size_t)ncell * LE));
  k_uhm<<<(ncell + 127) / 128, 128>>>(ncell, h->d_velo.p, h->d_diff.p, h->d_xmask.p, h->d_uhm.p);
  RK(cudaGetLastError());
  RK(h->d_cells.put(h->cells.data(), h->cells.size()));
  RK(h->d_node_ptr.put(h->node_ptr.data(), h->node_ptr.size()));
  RK(h->d_uh_s.res(h->cells.size() * NS));
  for (size_t n = 0; n < h->nodes.size(); n++) {
    const int64_t e0 = h->node_ptr[n], cnt = h->node_ptr[n + 1] - e0;
    if (cnt == 0) continue;
    k_grid_uh<<<(unsigned)cnt, 256>>>(h->d_cells.p + e0, h->nodes[n].cell, h->d_next.p, h->d_uhm.p, h->d_box.p,
                                      h->d_uh_s.p + (size_t)e0 * NS);
    RK(cudaGetLastError());
  }
  // MAKE_RESERVOIR_UH: the same path convolution from the reservoir cell to its target node's cell
  const int nres = (int)h->nodes.size() - 1;
  if (nres > 0) {
    std::vector<int> start(nres), target(nres), up_ptr(h->nodes.size() + 1, 0), up_idx;
    for (int j = 0; j < nres; j++) { start[j] = h->nodes[j].cell; target[j] = h->nodes[j].target; }
    for (size_t k = 0; k < h->nodes.size(); k++) {
      for (int j = 0; j < nres; j++) if (target[j] == (int)k) up_idx.push_back(j);
      up_ptr[k + 1] = (int)up_idx.size();
    }
    DB<int> d_start;
    RK(d_start.put(start.data(), nres));
    RK(h->d_uh_r.res((size_t)nres * NS));
    for (int j = 0; j < nres; j++) {
      k_grid_uh<<<1, 256>>>(d_start.p + j, h->nodes[target[j]].cell, h->d_next.p, h->d_uhm.p, h->d_box.p,
                            h->d_uh_r.p + (size_t)j * NS);
      RK(cudaGetLastError());
    }
    RK(cudaDeviceSynchronize());
    d_start.rel();
    h->uh_r.resize((size_t)nres * NS);
    RK(cudaMemcpy(h->uh_r.data(), h->d_uh_r.p, h->uh_r.size() * sizeof(float), cudaMemcpyDeviceToHost));
    RK(h->d_target.put(target.data(), nres));
    RK(h->d_up_ptr.put(up_ptr.data(), up_ptr.size()));
    RK(h->d_up_idx.put(up_idx.data(), up_idx.size()));
  }
  RK(cudaDeviceSynchronize());
  return VR_OK;
}

int vr_route_create(const vr_route_grid *g, int32_t station_col, int32_t station_row, const float *uh_box, int device,
                    vr_route **out)
{
  if (!g || !out || !uh_box || !g->dir || !g->velo || !g->diff || !g->xmask || g->ncol < 1 || g->nrow < 1) {
    g_route_error = "null or empty argument";
    return VR_ERR_ARG;
  }
  *out = nullptr;
  if (station_col < 1 || station_col > g->ncol || station_row < 1 || station_row > g->nrow) {
    g_route_error = "station outside the grid";
    return VR_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_route_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
    return VR_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_route_error = "device index out of range"; return VR_ERR_ARG; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_route_error = cudaGetErrorString(e); return VR_ERR_CUDA; }
  vr_route *h = new vr_route();
  h->device = device;
  build_graph(g, h->g);
  // station cell: PI = ICOL + 1 - col (rout.f:341), PJ = row
  const int PI = g->ncol + 1 - station_col, PJ = station_row, station = gidx(h->g, PI, PJ);
  {
    CatchmentMaps cm;
    catchment_maps(h->g, station, cm);
    std::vector<int> node_of(g->ncol * g->nrow, -1);
    for (int I = 1; I <= g->ncol; I++)                // the reference scans I outer, J inner
      for (int J = 1; J <= g->nrow; J++) {            // reservoirs in the reference's discovery (scan) order
        const int k = gidx(h->g, I, J), r = h->g.reser[k];
        if (cm.whole[k] && r > 0 && r != 9999) {
          Node nd;
          nd.res_id = r; nd.cell = k;
          node_of[k] = (int)h->nodes.size();
          h->nodes.push_back(std::move(nd));
        }
      }
    Node st;
    st.res_id = 0; st.cell = station;
    for (int I = 1; I <= g->ncol; I++)
      for (int J = 1; J <= g->nrow; J++) {
        const int k = gidx(h->g, I, J);
        const int fr = cm.first_res[k];
        if (fr >= 0 && node_of[fr] >= 0) h->nodes[node_of[fr]].cells.push_back(k);
        if (cm.direct[k]) st.cells.push_back(k);
      }
    h->nodes.push_back(std::move(st));
  }
  // RES_DIRECT(:,2): the next reservoir met walking downstream (reservoir.f:733-759); its release is routed to
  // that reservoir when it is one of this station's nodes, else to the station (reservoir.f:1529-1567)
  for (size_t n = 0; n + 1 < h->nodes.size(); n++) {
    Node &nd = h->nodes[n];
    int k = h->g.next[nd.cell];
    nd.ds_id = 0;
    for (int guard = 0; k >= 0 && guard <= g->ncol * g->nrow + 4; guard++) {
      const int r = h->g.reser[k];
      if (r > 0 && r != 9999) { nd.ds_id = r; break; }
      k = h->g.next[k];
    }
    nd.target = (int)h->nodes.size() - 1;
    for (size_t m = 0; m + 1 < h->nodes.size(); m++)
      if (nd.ds_id > 0 && h->nodes[m].res_id == nd.ds_id) { nd.target = (int)m; break; }
  }
  h->node_ptr.assign(1, 0);
  for (const Node &nd : h->nodes) {
    for (int k : nd.cells) {
      if (h->g.reser[k] == 9999) h->has_surface = true;
      h->cells.push_back(k);
    }
    h->node_ptr.push_back((int64_t)h->cells.size());
  }
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  int rc = route_build(h, g, uh_box);
  if (rc != VR_OK) { g_route_error = h->err; vr_route_destroy(h); return rc; }
  *out = h;
  return VR_OK;
}

int32_t vr_route_num_nodes(const vr_route *h) { return h ? (int32_t)h->nodes.size() : 0; }

int vr_route_node_info(const vr_route *h, int32_t node, int32_t *res_id, int64_t *ncells)
{
  if (!h || node < 0 || node >= (int32_t)h->nodes.size()) return VR_ERR_ARG;
  if (res_id) *res_id = h->nodes[node].res_id;
  if (ncells) *ncells = (int64_t)h->nodes[node].cells.size();
  return VR_OK;
}

int vr_route_node_cells(const vr_route *h, int32_t node, int32_t *grid_index)
{
  if (!h || !grid_index || node < 0 || node >= (int32_t)h->nodes.size()) return VR_ERR_ARG;
  for (size_t k = 0; k < h->nodes[node].cells.size(); k++) grid_index[k] = h->nodes[node].cells[k];
  return VR_OK;
}

int vr_route_get_uh(vr_route *h, int32_t node, float *uh_s)
{
  if (!h || !uh_s || node < 0 || node >= (int32_t)h->nodes.size()) return VR_ERR_ARG;
  RK(cudaSetDevice(h->device));
  const int64_t e0 = h->node_ptr[node], cnt = h->node_ptr[node + 1] - e0;
  if (cnt) RK(cudaMemcpy(uh_s, h->d_uh_s.p + (size_t)e0 * NS, (size_t)cnt * NS * sizeof(float), cudaMemcpyDeviceToHost));
  return VR_OK;
}

int vr_route_convolve_in(vr_route *h, int32_t ndays, const vr_route_inputs *in, float *flows, float *res_evaporation)
{
  if (!h || !in || !in->runoff || !in->baseflow || !flows || ndays < 1) return VR_ERR_ARG;
  RK(cudaSetDevice(h->device));
  const bool met = in->rh && in->tair && in->wind && in->cell_lat && in->cell_lon;
  if (h->has_surface && !met) {
    h->err = "reservoir-surface cells (9999) inside a catchment: MAKE_CONVOLUTION needs rh/tair/wind and the cell "
             "coordinates for their open-water evaporation (vr_route_convolve_in)";
    return VR_ERR_UNSUPPORTED;
  }
  const size_t ncell = (size_t)h->g.ncol * h->g.nrow, n = ncell * ndays;
  RK(h->d_ro.put(in->runoff, n)); RK(h->d_ba.put(in->baseflow, n));
  ConvArgs a;
  memset(&a, 0, sizeof a);
  a.nnodes = (int)h->nodes.size(); a.ndays = ndays; a.node_ptr = h->d_node_ptr.p; a.cells = h->d_cells.p;
  a.uh_s = h->d_uh_s.p; a.runoff = h->d_ro.p; a.baseflow = h->d_ba.p; a.scale = in->scale;
  if (in->cell_scale) { RK(h->d_scale.put(in->cell_scale, ncell)); a.cell_scale = h->d_scale.p; }
  if (in->ev_vege && in->ev_soil && in->tr_vege && in->cell_factor) {
    RK(h->d_evv.put(in->ev_vege, n)); RK(h->d_evs.put(in->ev_soil, n)); RK(h->d_trv.put(in->tr_vege, n));
    RK(h->d_factor.put(in->cell_factor, ncell));
    a.ev_vege = h->d_evv.p; a.ev_soil = h->d_evs.p; a.tr_vege = h->d_trv.p; a.cell_factor = h->d_factor.p;
  }
  if (in->present) { RK(h->d_present.put(in->present, ncell)); a.present = h->d_present.p; }
  if (met) {
    RK(h->d_rh.put(in->rh, n)); RK(h->d_tair.put(in->tair, n)); RK(h->d_wind.put(in->wind, n));
    RK(h->d_lat.put(in->cell_lat, ncell)); RK(h->d_lon.put(in->cell_lon, ncell));
    RK(h->d_reser.put(h->g.reser.data(), ncell));
    a.rh = h->d_rh.p; a.tair = h->d_tair.p; a.wind = h->d_wind.p; a.cell_lat = h->d_lat.p; a.cell_lon = h->d_lon.p;
    a.reser = h->d_reser.p;
    if (in->node_ope_year) { RK(h->d_ope.put(in->node_ope_year, a.nnodes)); a.node_ope_year = h->d_ope.p; }
    // both are computed from NDAY, not from the day (reservoir.f:420-421)
    a.month_of_year = (ndays % 365) / 30 + in->start_month;
    if (a.month_of_year > 12) a.month_of_year = 12;
    if (a.month_of_year < 1) { h->err = "start_month must be 1..12"; return VR_ERR_ARG; }
    a.current_year = in->start_year + ndays % 365;
    if (res_evaporation) { RK(h->d_resev.res((size_t)a.nnodes * ndays)); a.res_evap = h->d_resev.p; }
  }
  RK(h->d_flows.res((size_t)a.nnodes * ndays));
  a.flows = h->d_flows.p;
  int n0 = 0, n1 = a.nnodes;
  if (in->node_end > 0) {
    n0 = in->node_begin; n1 = in->node_end;
    if (n0 < 0 || n1 > a.nnodes || n0 >= n1) { h->err = "node range outside [0, nnodes)"; return VR_ERR_ARG; }
  }
  a.node_begin = n0;
  dim3 grid((ndays + 127) / 128, n1 - n0);
  cudaEventRecord(h->ev0);
  k_convolve<<<grid, 128>>>(a);
  cudaEventRecord(h->ev1);
  h->timed = true;
  RK(cudaGetLastError());
  const size_t off = (size_t)n0 * ndays, cnt = (size_t)(n1 - n0) * ndays;
  RK(cudaMemcpy(flows + off, h->d_flows.p + off, cnt * sizeof(float), cudaMemcpyDeviceToHost));
  if (res_evaporation) {
    if (a.res_evap) RK(cudaMemcpy(res_evaporation + off, h->d_resev.p + off, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    else memset(res_evaporation + off, 0, cnt * sizeof(float));
  }
  return VR_OK;
}

// ---- flux hand-off on the device -------------------------------------------------------------------------------
// A value as rout reads it back from vicNl's "%.4f" column (write_data.c -> READ(20,*), reservoir.f:373-389): the decimal
// with four places nearest to the binary double (ties of the exact value to even, like printf), converted to REAL.
// x*10^4 is split into its rounded product and the exact residual (fma), so the nearest integer is taken of the exact
// value; n/10^4 is the correctly rounded double of that decimal (what strtod returns), then one rounding to float --
// the same two steps as the host path (vr_round_decimals, then the float conversion).
__device__ __forceinline__ float through_text4(double x)
{
  const double p = x * 1e4, e = fma(x, 1e4, -p);
  double n = rint(p);
  // p - n is exact and, unless it is +-0.5, at least one ulp of p away from +-0.5 while |e| is at most half an ulp:
  // only a product that rounds onto a half decides by the sign of the residual (an exact tie keeps rint's even choice)
  const double d = p - n;
  if (d == 0.5 && e > 0.0) n += 1.0;
  else if (d == -0.5 && e < 0.0) n -= 1.0;
  return (float)(n / 1e4);
}

// tile of 32 records x 32 cells: coalesced reads along the cells, coalesced writes along the days of each grid cell
__global__ void __launch_bounds__(1024) k_fluxes_from_step(const double *__restrict__ out, int64_t C, int ndays, int nvar,
                                                           const int32_t *__restrict__ rows, const int64_t *__restrict__ grid_of_cell,
                                                           float *const *__restrict__ dst)
{
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
  for (int v = 0; v < nvar; v++) {
    const double *src = out + (size_t)rows[v] * ndays * C;
    const int64_t c = c0 + tx;
    const int d = d0 + ty;
    tile[ty][tx] = (c < C && d < ndays) ? through_text4(src[(size_t)d * C + c]) : 0.f;
    __syncthreads();
    const int64_t cw = c0 + ty;                       // this warp writes the days of cell cw
    const int dw = d0 + tx;
    if (cw < C && dw < ndays) {
      const int64_t g = grid_of_cell[cw];
      if (g >= 0) dst[v][(size_t)g * ndays + dw] = tile[tx][ty];
    }
    __syncthreads();
  }
}

int vr_route_fluxes_from_step(int device, const double *out_dev, int64_t ncell, int32_t ndays, int32_t nvar, const int32_t *rows,
                              const int64_t *grid_of_cell, int64_t ngrid, float *const dst_dev[], unsigned char *present_dev)
{
  if (!out_dev || ncell < 1 || ndays < 1 || nvar < 1 || nvar > 16 || !rows || !grid_of_cell || !dst_dev || ngrid < 1) return VR_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return VR_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  for (int64_t c = 0; c < ncell; c++)
    if (grid_of_cell[c] >= ngrid) return VR_ERR_ARG;
  DB<int32_t> d_rows;
  DB<int64_t> d_map;
  DB<float *> d_dst;
  std::vector<float *> dst(dst_dev, dst_dev + nvar);
  cudaError_t e = d_rows.put(rows, nvar);
  if (e == cudaSuccess) e = d_map.put(grid_of_cell, ncell);
  if (e == cudaSuccess) e = d_dst.put(dst.data(), nvar);
  for (int v = 0; v < nvar && e == cudaSuccess; v++) e = cudaMemset(dst_dev[v], 0, (size_t)ngrid * ndays * sizeof(float));
  if (e == cudaSuccess && present_dev) {
    std::vector<unsigned char> pres(ngrid, 0);
    for (int64_t c = 0; c < ncell; c++) if (grid_of_cell[c] >= 0) pres[grid_of_cell[c]] = 1;
    e = cudaMemcpy(present_dev, pres.data(), ngrid, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    dim3 grid((unsigned)((ncell + 31) / 32), (unsigned)((ndays + 31) / 32));
    k_fluxes_from_step<<<grid, 1024>>>(out_dev, ncell, ndays, nvar, d_rows.p, d_map.p, d_dst.p);
    e = cudaDeviceSynchronize();
  }
  if (d_rows.p) cudaFree(d_rows.p);
  if (d_map.p) cudaFree(d_map.p);
  if (d_dst.p) cudaFree(d_dst.p);
  return e == cudaSuccess ? VR_OK : VR_ERR_CUDA;
}

int vr_route_convolve_ex(vr_route *h, int32_t ndays, const float *runoff, const float *baseflow, float scale,
                         const float *cell_scale, const float *ev_vege, const float *ev_soil, const float *tr_vege,
                         const float *cell_factor, const unsigned char *present, float *flows)
{
  vr_route_inputs in;
  memset(&in, 0, sizeof in);
  in.runoff = runoff; in.baseflow = baseflow; in.scale = scale; in.cell_scale = cell_scale; in.ev_vege = ev_vege;
  in.ev_soil = ev_soil; in.tr_vege = tr_vege; in.cell_factor = cell_factor; in.present = present;
  return vr_route_convolve_in(h, ndays, &in, flows, nullptr);
}

int vr_route_convolve(vr_route *h, int32_t ndays, const float *runoff, const float *baseflow, float scale, float *flows)
{
  return vr_route_convolve_ex(h, ndays, runoff, baseflow, scale, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, flows);
}

int vr_route_last_kernel_ms(vr_route *h, float *ms)
{
  if (!h || !ms) return VR_ERR_ARG;
  if (!h->timed) { h->err = "no convolution launched yet"; return VR_ERR_STATE; }
  RK(cudaSetDevice(h->device));
  RK(cudaEventSynchronize(h->ev1));
  RK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return VR_OK;
}

int vr_objectives(int device, int32_t nsets, int32_t ndays, const float *sim, const double *obs, int32_t handle_negatives,
                  double *out)
{
  if (!sim || !obs || !out || nsets < 1 || ndays < 1) { g_route_error = "null or empty argument"; return VR_ERR_ARG; }
  if (cudaSetDevice(device) != cudaSuccess) { g_route_error = "no such CUDA device (this library has no CPU path)"; return VR_ERR_CUDA; }
  float *d_sim = nullptr; double *d_obs = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_sim, (size_t)nsets * ndays * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_obs, (size_t)ndays * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, (size_t)nsets * 4 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(d_sim, sim, (size_t)nsets * ndays * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_obs, obs, (size_t)ndays * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_objectives<<<nsets, 256>>>(ndays, d_sim, d_obs, handle_negatives, d_out);
    e = cudaMemcpy(out, d_out, (size_t)nsets * 4 * sizeof(double), cudaMemcpyDeviceToHost);
  }
  cudaFree(d_sim); cudaFree(d_obs); cudaFree(d_out);
  if (e != cudaSuccess) { g_route_error = cudaGetErrorString(e); return VR_ERR_CUDA; }
  return VR_OK;
}

int vr_route_res_info(const vr_route *h, int32_t node, int32_t *downstream_id, int32_t *target_node)
{
  if (!h || node < 0 || node >= (int32_t)h->nodes.size()) return VR_ERR_ARG;
  if (downstream_id) *downstream_id = h->nodes[node].ds_id;
  if (target_node) *target_node = h->nodes[node].target;
  return VR_OK;
}

int vr_route_get_uh_r(vr_route *h, int32_t node, float *uh_r)
{
  if (!h || !uh_r || node < 0 || node + 1 >= (int32_t)h->nodes.size()) return VR_ERR_ARG;
  memcpy(uh_r, h->uh_r.data() + (size_t)node * NS, NS * sizeof(float));
  return VR_OK;
}

int vr_route_set_uh_r(vr_route *h, int32_t node, const float *uh_r)
{
  if (!h || !uh_r || node < 0 || node + 1 >= (int32_t)h->nodes.size()) return VR_ERR_ARG;
  RK(cudaSetDevice(h->device));
  memcpy(h->uh_r.data() + (size_t)node * NS, uh_r, NS * sizeof(float));
  RK(cudaMemcpy(h->d_uh_r.p + (size_t)node * NS, uh_r, NS * sizeof(float), cudaMemcpyHostToDevice));
  return VR_OK;
}

int vr_route_reservoirs(vr_route *h, int32_t ndays, int32_t nsets, const float *natural, const vr_reservoir *res,
                        const vr_res_options *opt, float *flow, const vr_res_series *out)
{
  return vr_route_reservoirs_range(h, ndays, nsets, 0, ndays, natural, res, opt, flow, out);
}

// Days [day_begin, day_end) of the day loop (0-based).  day_begin > 0 continues a run: `out` then carries the state IN as
// well -- out->vol[.][day_begin] (VOL of the first day to operate) and out->flowout[.][0 .. day_begin) (the releases whose
// routing through UH_R still reaches later days) must hold what the previous call left; all series of `out` that are
// given are updated for the operated days, `flow` for every day < day_end.  This is the reference's STEPBYSTEP mode
// (reservoir.f:1142-1150, 1596-1675: one day per run of `rout`, VOL / FLOWOUT / RESFLOWS carried in files).
int vr_route_reservoirs_range(vr_route *h, int32_t ndays, int32_t nsets, int32_t day_begin, int32_t day_end, const float *natural,
                              const vr_reservoir *res, const vr_res_options *opt, float *flow, const vr_res_series *out)
{
  if (!h || !natural || !opt || !flow || ndays < 1 || nsets < 1) return VR_ERR_ARG;
  if (day_begin < 0 || day_end > ndays || day_begin >= day_end) return VR_ERR_ARG;
  if (day_begin > 0 && (!out || !out->vol || !out->flowout)) { h->err = "a continued day loop needs the vol and flowout series of the run so far"; return VR_ERR_ARG; }
  const int nnodes = (int)h->nodes.size(), nres = nnodes - 1;
  if (nres > 0 && !res) return VR_ERR_ARG;
  RK(cudaSetDevice(h->device));
  // host preparation: everything that needs libm (tan, sqrt) or is constant over the days
  std::vector<ResC> rc((size_t)nsets * nres);
  std::vector<float> pool;
  for (size_t q = 0; q < rc.size(); q++) {
    const vr_reservoir &r = res[q];
    ResC &c = rc[q];
    memset(&c, 0, sizeof c);
    if (r.rule < 0 || r.rule > 6) { h->err = "reservoir rule must be 0..6"; return VR_ERR_ARG; }
    c.hresermax = r.hresermax; c.vreser = r.v_live + r.v_dead; c.v_dead = r.v_dead; c.head = r.head; c.q_design = r.q_design;
    c.qmin = r.env_flow * r.q_design;
    c.vol1 = (opt->start_year >= r.ope_year) ? r.v_initial : 0.001f;
    c.seepage = r.seepage; c.loss = r.seepage + r.infiltration;
    if (r.bathymetry == 1) c.H0 = r.bth[0] / (1 + r.bth[3]);
    else {
      const float KFACTOR = sqrtf(r.v_dead / r.v_live);
      c.H0 = (KFACTOR * r.hresermax - r.hresermin) / (KFACTOR - 1);
    }
    c.span = r.hresermax - c.H0;
    c.hmax = r.hmax; c.hmin = r.hmin;
    if (r.rule == 1 && c.hmax < c.hmin) { const float t = c.hmax; c.hmax = c.hmin; c.hmin = t; }
    c.op1a = r.op1[0]; c.op1b = r.op1[1];
    c.demand = r.demand; c.x2 = r.x2; c.x3 = r.x3; c.tan1 = tanf(r.x1); c.tan4 = tanf(r.x4);
    for (int m = 0; m < 12; m++) {
      c.reslv[m] = r.reslv[m]; c.demand5[m] = r.demand5[m]; c.x2_5[m] = r.op5x2[m]; c.x3_5[m] = r.op5x3[m];
      c.tan1_5[m] = tanf(r.op5x1[m]); c.tan4_5[m] = tanf(r.op5x4[m]);
    }
    for (int m = 0; m < 4; m++) c.bth[m] = r.bth[m];
    c.ope_year = r.ope_year; c.rule = r.rule; c.hp_type = r.hydropower_type; c.hp_gen = r.hydropower_generator;
    c.bathymetry = r.bathymetry;
    c.series = -1; c.irrigation = -1;
    if ((r.rule == 4 || r.rule == 6) && r.series) { c.series = (long long)pool.size(); pool.insert(pool.end(), r.series, r.series + ndays); }
    if (r.irrigation) { c.irrigation = (long long)pool.size(); pool.insert(pool.end(), r.irrigation, r.irrigation + ndays); }
  }
  const size_t S = (size_t)nsets, R = (size_t)(nres > 0 ? nres : 1), D = (size_t)ndays;
  RK(h->d_res.put(rc.data(), rc.size()));
  RK(h->d_pool.put(pool.data(), pool.size()));
  RK(h->d_natural.put(natural, (size_t)nnodes * D));
  RK(h->d_resflows.res(S * nnodes * (D + 2)));
  RK(h->d_vol.res(S * R * (D + 1))); RK(h->d_hho.res(S * R * (D + 1)));
  RK(h->d_flowin.res(S * R * D)); RK(h->d_flowout.res(S * R * D)); RK(h->d_turb.res(S * R * D));
  RK(h->d_energy.res(S * R * D)); RK(h->d_htk.res(S * R * D)); RK(h->d_flow.res(S * D));
  RK(cudaMemset(h->d_hho.p, 0, S * R * (D + 1) * sizeof(float)));
  if (day_begin > 0 && nres > 0) {       // the state of the run so far
    RK(cudaMemcpy(h->d_vol.p, out->vol, S * R * (D + 1) * sizeof(float), cudaMemcpyHostToDevice));
    RK(cudaMemcpy(h->d_flowout.p, out->flowout, S * R * D * sizeof(float), cudaMemcpyHostToDevice));
    if (out->hho) RK(cudaMemcpy(h->d_hho.p, out->hho, S * R * (D + 1) * sizeof(float), cudaMemcpyHostToDevice));
    if (out->flowin) RK(cudaMemcpy(h->d_flowin.p, out->flowin, S * R * D * sizeof(float), cudaMemcpyHostToDevice));
    if (out->turbine) RK(cudaMemcpy(h->d_turb.p, out->turbine, S * R * D * sizeof(float), cudaMemcpyHostToDevice));
    if (out->energy) RK(cudaMemcpy(h->d_energy.p, out->energy, S * R * D * sizeof(float), cudaMemcpyHostToDevice));
    if (out->htk) RK(cudaMemcpy(h->d_htk.p, out->htk, S * R * D * sizeof(float), cudaMemcpyHostToDevice));
  }
  std::vector<int> up0(nnodes + 1, 0);
  if (nres == 0) RK(h->d_up_ptr.put(up0.data(), up0.size()));
  ResArgs a;
  a.nnodes = nnodes; a.ndays = ndays; a.start_year = opt->start_year; a.start_month = opt->start_month;
  a.day0 = day_begin + 1; a.day1 = day_end;
  a.legacy = opt->legacy_rule_curve; a.res = h->d_res.p; a.pool = h->d_pool.p; a.natural = h->d_natural.p;
  a.uh_r = h->d_uh_r.p; a.target = h->d_target.p; a.up_ptr = h->d_up_ptr.p; a.up_idx = h->d_up_idx.p;
  a.resflows = h->d_resflows.p; a.vol = h->d_vol.p; a.hho = h->d_hho.p; a.flowin = h->d_flowin.p;
  a.flowout = h->d_flowout.p; a.turb = h->d_turb.p; a.energy = h->d_energy.p; a.htk = h->d_htk.p; a.flow = h->d_flow.p;
  int threads = ((nnodes + 31) / 32) * 32;
  if (threads < 512) threads = 512;              // the day-parallel station pass wants threads, the day loop does not mind
  if (threads > 1024) threads = 1024;
  cudaEventRecord(h->ev0);
  k_reservoirs<<<nsets, threads>>>(a);
  cudaEventRecord(h->ev1);
  h->timed = true;
  RK(cudaGetLastError());
  RK(cudaMemcpy(flow, h->d_flow.p, S * D * sizeof(float), cudaMemcpyDeviceToHost));
  if (out && nres > 0) {
    const size_t n1 = S * R * (D + 1) * sizeof(float), n0 = S * R * D * sizeof(float);
    if (out->vol) RK(cudaMemcpy(out->vol, h->d_vol.p, n1, cudaMemcpyDeviceToHost));
    if (out->hho) RK(cudaMemcpy(out->hho, h->d_hho.p, n1, cudaMemcpyDeviceToHost));
    if (out->flowin) RK(cudaMemcpy(out->flowin, h->d_flowin.p, n0, cudaMemcpyDeviceToHost));
    if (out->flowout) RK(cudaMemcpy(out->flowout, h->d_flowout.p, n0, cudaMemcpyDeviceToHost));
    if (out->turbine) RK(cudaMemcpy(out->turbine, h->d_turb.p, n0, cudaMemcpyDeviceToHost));
    if (out->energy) RK(cudaMemcpy(out->energy, h->d_energy.p, n0, cudaMemcpyDeviceToHost));
    if (out->htk) RK(cudaMemcpy(out->htk, h->d_htk.p, n0, cudaMemcpyDeviceToHost));
  }
  return VR_OK;
}

}  // extern "C"
