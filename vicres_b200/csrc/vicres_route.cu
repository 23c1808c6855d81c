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

struct Node { int res_id = 0, cell = -1; std::vector<int> cells; };

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
};

__global__ void __launch_bounds__(128) k_convolve(ConvArgs a)
{
  const int node = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;   // i = day - 1
  if (i >= a.ndays) return;
  float flow = 0.0f;
  long long last = -1;
  const bool evap = a.ev_vege && a.ev_soil && a.tr_vege && a.cell_factor;
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
    // a missing cell contributes UH*(0*sc + 0*sc) = +0 to every day: nothing to add
    if (evap && last >= 0) {
      const size_t q = (size_t)last * a.ndays + i;
      if (0.0f > (a.ev_vege[q] + a.ev_soil[q] + a.tr_vege[q]))
        flow = flow - 0.0f + (a.ev_soil[q] + a.ev_vege[q] + a.tr_vege[q]) * a.cell_factor[gi] * 0.028f / 1000;
    }
  }
  a.flows[(size_t)node * a.ndays + i] = flow;
}

template <class T>
struct DB {
  T *p = nullptr; size_t n = 0;
  cudaError_t put(const T *src, size_t count)
  {
    cudaError_t e = res(count);
    if (e != cudaSuccess || !count) return e;
    return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
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
  std::vector<int> whole;
  list_cells(h->g, station, 0, whole);
  for (int k : whole) {                               // reservoirs in the reference's discovery (scan) order
    const int r = h->g.reser[k];
    if (r > 0 && r != 9999) {
      Node nd;
      nd.res_id = r; nd.cell = k;
      list_cells(h->g, k, 1, nd.cells);
      h->nodes.push_back(std::move(nd));
    }
  }
  {
    Node nd;
    nd.res_id = 0; nd.cell = station;
    list_cells(h->g, station, 2, nd.cells);
    h->nodes.push_back(std::move(nd));
  }
  h->node_ptr.assign(1, 0);
  for (const Node &nd : h->nodes) {
    for (int k : nd.cells) {
      if (h->g.reser[k] == 9999) {
        g_route_error = "reservoir-surface cells (9999) inside a catchment need the open-water evaporation of "
                        "reservoir.f:410-475, which is not part of this round";
        delete h;
        return VR_ERR_UNSUPPORTED;
      }
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

int vr_route_convolve_ex(vr_route *h, int32_t ndays, const float *runoff, const float *baseflow, float scale,
                         const float *cell_scale, const float *ev_vege, const float *ev_soil, const float *tr_vege,
                         const float *cell_factor, const unsigned char *present, float *flows)
{
  if (!h || !runoff || !baseflow || !flows || ndays < 1) return VR_ERR_ARG;
  RK(cudaSetDevice(h->device));
  const size_t ncell = (size_t)h->g.ncol * h->g.nrow, n = ncell * ndays;
  RK(h->d_ro.put(runoff, n)); RK(h->d_ba.put(baseflow, n));
  ConvArgs a;
  memset(&a, 0, sizeof a);
  a.nnodes = (int)h->nodes.size(); a.ndays = ndays; a.node_ptr = h->d_node_ptr.p; a.cells = h->d_cells.p;
  a.uh_s = h->d_uh_s.p; a.runoff = h->d_ro.p; a.baseflow = h->d_ba.p; a.scale = scale;
  if (cell_scale) { RK(h->d_scale.put(cell_scale, ncell)); a.cell_scale = h->d_scale.p; }
  if (ev_vege && ev_soil && tr_vege && cell_factor) {
    RK(h->d_evv.put(ev_vege, n)); RK(h->d_evs.put(ev_soil, n)); RK(h->d_trv.put(tr_vege, n));
    RK(h->d_factor.put(cell_factor, ncell));
    a.ev_vege = h->d_evv.p; a.ev_soil = h->d_evs.p; a.tr_vege = h->d_trv.p; a.cell_factor = h->d_factor.p;
  }
  if (present) { RK(h->d_present.put(present, ncell)); a.present = h->d_present.p; }
  RK(h->d_flows.res((size_t)a.nnodes * ndays));
  a.flows = h->d_flows.p;
  dim3 grid((ndays + 127) / 128, a.nnodes);
  cudaEventRecord(h->ev0);
  k_convolve<<<grid, 128>>>(a);
  cudaEventRecord(h->ev1);
  h->timed = true;
  RK(cudaGetLastError());
  RK(cudaMemcpy(flows, h->d_flows.p, (size_t)a.nnodes * ndays * sizeof(float), cudaMemcpyDeviceToHost));
  return VR_OK;
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

}  // extern "C"
