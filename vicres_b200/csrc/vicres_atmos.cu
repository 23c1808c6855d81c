// vicres_atmos.cu -- kernels and C ABI of the forcing preparation (include/vicres_atmos.h).
//
// Five kernels over one batch of cells at a time (scratch tables of a batch: 82 kB per cell for the solar
// geometry + 5 doubles and 2 bytes per cell-day), each a thin launch of a body of vic_atmos.cuh:
//   k_solar   one thread per (cell, yearday)   -- the 30-second hour-angle walk, >99 % of the arithmetic
//   k_prep    one thread per cell              -- recurrences over days (MTCLIM snowpack), daily totals
//   k_rad_tiled  one thread per (cell, local day), blocks of 32 cells x 32 days with the windows staged in shared
//                memory -- radiation / humidity chain of the day (k_rad: the same without the tile)
//   k_conv    one thread per cell              -- VP_ITER_CONVERGE only
//   k_hours   one thread per (cell, local day) -- hours of Tmin / Tmax
//   k_rec     one thread per (cell, record)    -- the nine atmos fields of the record
// Threads of a warp are 32 consecutive cells of the same row, so every scratch and output access is a
// coalesced 256-byte line.  No CPU path: without a device the entry points return VR_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "vic_atmos.cuh"

namespace {

constexpr int TPB = 128;
thread_local std::string g_err;
struct Timing { std::vector<cudaEvent_t> ev; int64_t launches = 0; cudaStream_t stream = nullptr; int device = 0; };
thread_local Timing g_t;

int fail(int code, const std::string &m) { g_err = m; return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(VR_ERR_CUDA, std::string(#call ": ") + cudaGetErrorString(e_)); } while (0)

// row = blockIdx.x / nbc, cells of the block = (blockIdx.x % nbc) * TPB ...
__device__ __forceinline__ bool locate(const va::Ctx &x, int64_t cb, int &row, int64_t &c)
{
  const int64_t nbc = (cb + TPB - 1) / TPB;
  row = (int)(blockIdx.x / nbc);
  const int64_t lc = (blockIdx.x % nbc) * TPB + threadIdx.x;
  c = x.c0 + lc;
  return lc < cb;
}
// The trip count of the hour-angle walk is the daylength, a monotonic function of latitude on a given yearday: the
// threads of a warp take cells that are neighbours in latitude (x.perm), not in the caller's order, so that they leave
// the walk together.  Only this kernel uses the order; it writes its tables at the cell's own position.
constexpr int LAT_BUCKETS = 1440;               // 1/8 degree
__global__ void k_lat_hist(const __grid_constant__ va::Ctx x, int64_t cb, int32_t *hist)
{
  const int64_t lc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= cb) return;
  int b = (int)((x.lat[x.c0 + lc] + 90.0) * 8.0);
  b = b < 0 ? 0 : (b >= LAT_BUCKETS ? LAT_BUCKETS - 1 : b);
  atomicAdd(&hist[b], 1);
}
__global__ void k_lat_scan(int32_t *hist)        // one thread: 1440 additions
{
  int32_t run = 0;
  for (int b = 0; b < LAT_BUCKETS; b++) { const int32_t n = hist[b]; hist[b] = run; run += n; }
}
__global__ void k_lat_scatter(const __grid_constant__ va::Ctx x, int64_t cb, int32_t *hist, int32_t *perm)
{
  const int64_t lc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= cb) return;
  int b = (int)((x.lat[x.c0 + lc] + 90.0) * 8.0);
  b = b < 0 ? 0 : (b >= LAT_BUCKETS ? LAT_BUCKETS - 1 : b);
  perm[atomicAdd(&hist[b], 1)] = (int32_t)lc;
}

__global__ void __launch_bounds__(TPB) k_solar(const __grid_constant__ va::Ctx x, int64_t cb)
{
  __shared__ double s_ck[20];
  __shared__ double s_t21[21 * TPB];            // transmittance of the 21 air-mass table entries, per thread
  const double ck[20] = VA_ZENITH_STEPS;
  if (threadIdx.x < 20) s_ck[threadIdx.x] = ck[threadIdx.x];
  __syncthreads();
  int yd; int64_t c;
  if (!locate(x, cb, yd, c)) return;
  if (x.perm) c = x.c0 + x.perm[c - x.c0];
  va::solar_day_t<true>(x, c, yd, s_ck, s_t21 + threadIdx.x, TPB);
}
// the walk exactly as the reference writes it (libm call per step): diagnostic twin of k_solar
__global__ void __launch_bounds__(TPB) k_solar_exact(const __grid_constant__ va::Ctx x, int64_t cb)
{
  int yd; int64_t c;
  if (locate(x, cb, yd, c)) va::solar_day_t<false>(x, c, yd, nullptr, nullptr, 0);
}
__global__ void __launch_bounds__(TPB) k_prep(const __grid_constant__ va::Ctx x, int64_t cb)
{
  int r; int64_t c;
  if (locate(x, cb, r, c)) va::daily_prep(x, c);
}
__global__ void __launch_bounds__(TPB) k_rad(const __grid_constant__ va::Ctx x, int64_t cb)
{
  int day; int64_t c;
  if (locate(x, cb, day, c)) va::daily_rad(x, c, day);
}
__global__ void __launch_bounds__(va::RT_C * va::RT_D) k_rad_tiled(const __grid_constant__ va::Ctx x, int64_t cb)
{
  const int64_t ntc = (cb + va::RT_C - 1) / va::RT_C;
  va::daily_rad_tile(x, cb, blockIdx.x % ntc, (int)(blockIdx.x / ntc));
}
__global__ void __launch_bounds__(TPB) k_conv(const __grid_constant__ va::Ctx x, int64_t cb)
{
  int r; int64_t c;
  if (locate(x, cb, r, c)) va::converge(x, c);
}
__global__ void __launch_bounds__(TPB) k_hours(const __grid_constant__ va::Ctx x, int64_t cb)
{
  int day; int64_t c;
  if (locate(x, cb, day, c)) va::minmax_hour(x, c, day);
}
__global__ void __launch_bounds__(TPB) k_humid(const __grid_constant__ va::Ctx x, int64_t cb)
{
  int day; int64_t c;
  if (locate(x, cb, day, c)) va::humid_day(x, c, day);
}
// columns derived element by element from others (vic_atmos.cuh: derive_*): mode 0 RAINF + SNOWF, 1 |(WIND_E, WIND_N)|,
// 2 QAIR * PRESSURE / EPS, 3 REL_HUMID * svp(AIR_TEMP) / 100
__global__ void k_derive(const __grid_constant__ va::Ctx x, int mode, int64_t n, const double *a, const double *b, double *out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = mode == 0 ? va::derive_prec(x, a[i], b[i]) : mode == 1 ? va::derive_wind(a[i], b[i])
         : mode == 2 ? va::derive_vp_qair(x, a[i], b[i]) : va::derive_vp_rh(x, a[i], b[i]);
}
__global__ void __launch_bounds__(TPB) k_rec(const __grid_constant__ va::Ctx x, int64_t cb)
{
  int rec; int64_t c;
  if (!locate(x, cb, rec, c)) return;
  // the plain four-column layout in traditional units gets a copy of the body in which the observed-column branches
  // and the unit factors are compile-time dead (the fields are constants of the local copy)
  if (!x.r_sw && !x.r_lw && !x.r_vp && !x.r_pr && !x.r_de && !x.alma && !x.prec_native) {
    va::Ctx y = x;
    y.r_sw = nullptr; y.r_lw = nullptr; y.r_vp = nullptr; y.r_pr = nullptr; y.r_de = nullptr; y.alma = 0;
    va::record_inl(y, c, rec);
  }
  else va::record_inl(x, c, rec);
}

__global__ void k_diag_cr(int64_t n, const double *x, double *a, const double *y, double *c)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = vcr::acos_cr(x[i]); c[i] = vcr::cos_cr(y[i]); }
}

void drop_events()
{
  for (cudaEvent_t e : g_t.ev) cudaEventDestroy(e);
  g_t.ev.clear();
  g_t.launches = 0;
}

}  // namespace

extern "C" {

void vr_atmos_default_options(vr_atmos_options *o)
{
  if (!o) return;
  *o = vr_atmos_options{};
  o->time_step = 24; o->startyear = 1981; o->startmonth = 1; o->startday = 1;
  o->layout = VR_ATMOS_DAILY_TMAX_TMIN; o->force_dt = 24; o->has_wind = 1;
  o->vp_interp = 1; o->vp_iter = VR_VP_ITER_ALWAYS; o->mtclim_swe_corr = 0; o->lw_type = VR_LW_PRATA;
  o->lw_cloud = VR_LW_CLOUD_DEARDORFF; o->plapse = 1; o->min_wind_speed = 0.1; o->sw_prec_thresh = 0.;
  o->snow_step = 0; o->max_snow_temp = 0.5;
}

int32_t vr_atmos_ndays(const vr_atmos_options *o) { return o ? va::ndays_of(*o) : 0; }
const char *vr_atmos_last_error(void) { return g_err.c_str(); }

int vr_atmos_prepare_device(int device, const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw,
                            double *const out[VR_NFORCING], void *cuda_stream)
{
  if (o && o->snow_step > 0 && o->snow_step != o->time_step)
    return fail(VR_ERR_ARG, "SNOW_STEP < TIME_STEP: the output is [nrecs][NF + 1][C], use vr_atmos_prepare_sub_device");
  return vr_atmos_prepare_sub_device(device, o, cells, raw, out, nullptr, cuda_stream);
}

int vr_atmos_prepare_sub_device(int device, const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw,
                                double *const out[VR_NFORCING], int32_t *snowflag, void *cuda_stream)
{
  if (!o || !cells || !raw || !out) return fail(VR_ERR_ARG, "null argument");
  if (const char *why = va::unsupported(*o)) return fail(VR_ERR_UNSUPPORTED, why);
  if (cells->ncell < 1 || !cells->lat || !cells->lng || !cells->time_zone_lng || !cells->elevation)
    return fail(VR_ERR_ARG, "cells: lat, lng, time_zone_lng and elevation are required");
  const bool split_prec = raw->rainf && raw->snowf, split_wind = raw->wind_e && raw->wind_n;
  if ((!raw->prec && !split_prec) || !raw->t_a || (o->layout == VR_ATMOS_DAILY_TMAX_TMIN && !raw->t_b) ||
      (o->has_wind && !raw->wind && !split_wind))
    return fail(VR_ERR_ARG, "raw: a forcing column of the declared layout is missing");
  if (raw->nrec_forcing < 1) return fail(VR_ERR_ARG, "raw: nrec_forcing < 1");
  if (const char *why = va::unsupported_raw(*o, *raw)) return fail(VR_ERR_UNSUPPORTED, why);
  for (int f = 0; f < VR_NFORCING; f++) if (!out[f]) return fail(VR_ERR_ARG, "out: null field");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev)
    return fail(VR_ERR_CUDA, "no CUDA device: the forcing preparation has no CPU path");
  CK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  drop_events();
  g_t.stream = st; g_t.device = device;

  va::Ctx x{};
  va::fill_options(*o, x);
  const int64_t C = cells->ncell;
  const int nd = x.Ndays + 1;                       // local days of a cell whose local time is not the forcing's
  int64_t CB = o->reserved[0] > 0 ? o->reserved[0] : 32768;
  if (CB > C) CB = C;
  x.C = C; x.CB = CB; x.nrf = raw->nrec_forcing;
  x.lat = cells->lat; x.lng = cells->lng; x.tz = cells->time_zone_lng; x.elev = cells->elevation;
  x.slope = cells->slope; x.aspect = cells->aspect; x.ehoriz = cells->ehoriz; x.whoriz = cells->whoriz;
  x.r_prec = raw->prec; x.r_ta = raw->t_a; x.r_tb = raw->t_b; x.r_wind = raw->wind;
  x.r_sw = raw->shortwave; x.r_lw = raw->longwave; x.r_vp = raw->vp; x.r_pr = raw->pressure; x.r_de = raw->density;
  for (int f = 0; f < VR_NFORCING; f++) x.out[f] = out[f];
  x.min_tf = cells->min_Tfactor; x.snowflag = snowflag;
  // derived columns (initialize_atmos.c:424-460, 773-866): which, and from what
  const bool vp_q = !raw->vp && raw->qair && raw->pressure;
  const bool vp_rh = !raw->vp && !vp_q && !raw->qair && raw->rel_humid && o->layout == VR_ATMOS_SUBDAILY_AIR_TEMP;
  if (!raw->vp && !vp_q && !vp_rh) {                 // the daily layout: after MTCLIM (humid_day)
    if (raw->qair) x.r_q = raw->qair;
    else if (raw->rel_humid) x.r_rh = raw->rel_humid;
  }
  if (split_wind) x.has_wind = 1;
  const int nder = (split_prec ? 1 : 0) + (split_wind ? 1 : 0) + ((vp_q || vp_rh) ? 1 : 0);
  const size_t n_col = (size_t)raw->nrec_forcing * (size_t)C;

  // scratch of one batch, one stream-ordered allocation
  const bool conv = o->vp_iter == VR_VP_ITER_CONVERGE;
  const int ndaily = 4 + (o->mtclim_swe_corr ? 1 : 0) + (conv ? 2 : 0);
  const size_t n_tab = (size_t)va::NYD * 28 * CB, n_day = (size_t)nd * CB;
  const size_t bytes = (n_tab + n_day * ndaily + 2 * (size_t)CB) * sizeof(double) + 2 * n_day + (size_t)(x.Ndays + 2) * sizeof(int16_t) + 64 +
                       ((size_t)CB + LAT_BUCKETS) * sizeof(int32_t) + 16 + (size_t)nder * n_col * sizeof(double) + 64;
  char *base = nullptr;
  {   // keep the scratch in the device's default pool between calls (the default threshold returns it at every sync)
    cudaMemPool_t pool;
    uint64_t keep = UINT64_MAX;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  CK(cudaMallocAsync((void **)&base, bytes, st));
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaFreeAsync(base, st); return fail(VR_ERR_CUDA, std::string(#call ": ") + cudaGetErrorString(e_)); } } while (0)
  double *d = (double *)base;
  x.ttmax0 = d; x.flat = d + (size_t)va::NYD * CB; x.slp = d + (size_t)2 * va::NYD * CB; x.dayl = d + (size_t)3 * va::NYD * CB;
  x.hf = d + (size_t)4 * va::NYD * CB;
  d += n_tab;
  x.prec_d = d; d += n_day;
  x.srad = d; d += n_day;
  x.pva = d; d += n_day;
  x.tskc = d; d += n_day;
  if (o->mtclim_swe_corr) { x.swe = d; d += n_day; }
  if (conv) { x.tdew = d; d += n_day; x.parr = d; d += n_day; }
  x.tcon = d; d += 2 * CB;
  x.tminh = (int8_t *)d;
  x.tmaxh = x.tminh + n_day;
  int16_t *cal_dev = (int16_t *)(((uintptr_t)(x.tmaxh + n_day) + 15) & ~(uintptr_t)15);
  int32_t *perm_dev = (int32_t *)(((uintptr_t)(cal_dev + x.Ndays + 2) + 15) & ~(uintptr_t)15);
  int32_t *hist_dev = perm_dev + CB;
  if (nder) {
    double *dcol = (double *)(((uintptr_t)(hist_dev + LAT_BUCKETS) + 63) & ~(uintptr_t)63);
    const unsigned nb = (unsigned)((n_col + 255) / 256);
    if (split_prec) { k_derive<<<nb, 256, 0, st>>>(x, 0, (int64_t)n_col, raw->rainf, raw->snowf, dcol); x.r_prec = dcol; x.prec_native = 1; dcol += n_col; }
    if (split_wind) { k_derive<<<nb, 256, 0, st>>>(x, 1, (int64_t)n_col, raw->wind_e, raw->wind_n, dcol); x.r_wind = dcol; dcol += n_col; }
    if (vp_q) { k_derive<<<nb, 256, 0, st>>>(x, 2, (int64_t)n_col, raw->qair, raw->pressure, dcol); x.r_vp = dcol; x.vp_native = 1; }
    if (vp_rh) { k_derive<<<nb, 256, 0, st>>>(x, 3, (int64_t)n_col, raw->rel_humid, raw->t_a, dcol); x.r_vp = dcol; x.vp_native = 1; }
    g_t.launches += nder;
  }
  std::vector<int16_t> cal(x.Ndays + 2);
  va::build_calendar(*o, x.Ndays, cal.data());
  CKF(cudaMemcpyAsync(cal_dev, cal.data(), cal.size() * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  CKF(cudaStreamSynchronize(st));                      // `cal` is pageable host memory that dies with this call
  x.cal = cal_dev;

  auto mark = [&]() { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); g_t.ev.push_back(e); };
  for (int64_t c0 = 0; c0 < C; c0 += CB) {
    const int64_t cb = (C - c0 < CB) ? C - c0 : CB;
    const int64_t nbc = (cb + TPB - 1) / TPB;
    x.c0 = c0;
    mark();
    if (o->reserved[2] != 1) {                       // latitude order of the batch (reserved[2] == 1: caller's order)
      CKF(cudaMemsetAsync(hist_dev, 0, LAT_BUCKETS * sizeof(int32_t), st));
      k_lat_hist<<<(unsigned)((cb + 255) / 256), 256, 0, st>>>(x, cb, hist_dev);
      k_lat_scan<<<1, 1, 0, st>>>(hist_dev);
      k_lat_scatter<<<(unsigned)((cb + 255) / 256), 256, 0, st>>>(x, cb, hist_dev, perm_dev);
      x.perm = perm_dev;
      g_t.launches += 3;
    }
    k_solar<<<(unsigned)(nbc * va::NYD), TPB, 0, st>>>(x, cb);
    x.perm = nullptr;
    mark();
    k_prep<<<(unsigned)nbc, TPB, 0, st>>>(x, cb);
    if (o->reserved[1] == 1) k_rad<<<(unsigned)(nbc * nd), TPB, 0, st>>>(x, cb);     // untiled twin, for the tests
    else k_rad_tiled<<<(unsigned)(((cb + va::RT_C - 1) / va::RT_C) * ((nd + va::RT_D - 1) / va::RT_D)), va::RT_C * va::RT_D, 0, st>>>(x, cb);
    g_t.launches += 3;
    if (conv) { k_conv<<<(unsigned)nbc, TPB, 0, st>>>(x, cb); g_t.launches++; }
    mark();
    k_hours<<<(unsigned)(nbc * nd), TPB, 0, st>>>(x, cb);
    if (x.r_q || x.r_rh) { k_humid<<<(unsigned)(nbc * nd), TPB, 0, st>>>(x, cb); g_t.launches++; }
    k_rec<<<(unsigned)(nbc * x.nrecs), TPB, 0, st>>>(x, cb);
    g_t.launches += 2;
    mark();
  }
  CKF(cudaGetLastError());
#undef CKF
  CK(cudaFreeAsync(base, st));
  return VR_OK;
}

int vr_atmos_prepare(int device, const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw,
                     double *const out[VR_NFORCING])
{
  if (o && o->snow_step > 0 && o->snow_step != o->time_step)
    return fail(VR_ERR_ARG, "SNOW_STEP < TIME_STEP: the output is [nrecs][NF + 1][C], use vr_atmos_prepare_sub");
  return vr_atmos_prepare_sub(device, o, cells, raw, out, nullptr);
}

int vr_atmos_prepare_sub(int device, const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw,
                         double *const out[VR_NFORCING], int32_t *snowflag)
{
  if (!o || !cells || !raw || !out) return fail(VR_ERR_ARG, "null argument");
  if (const char *why = va::unsupported(*o)) return fail(VR_ERR_UNSUPPORTED, why);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev)
    return fail(VR_ERR_CUDA, "no CUDA device: the forcing preparation has no CPU path");
  CK(cudaSetDevice(device));
  const int64_t C = cells->ncell, nrf = raw->nrec_forcing;
  if (C < 1 || nrf < 1) return fail(VR_ERR_ARG, "empty input");
  const int NSITE = 10;
  const double *site_h[NSITE] = {cells->lat, cells->lng, cells->time_zone_lng, cells->elevation, cells->annual_prec,
                                 cells->slope, cells->aspect, cells->ehoriz, cells->whoriz, cells->min_Tfactor};
  const int NRAW = 15;
  const double *raw_h[NRAW] = {raw->prec, raw->t_a, raw->t_b, raw->wind, raw->shortwave, raw->longwave, raw->vp, raw->pressure,
                               raw->density, raw->qair, raw->rel_humid, raw->rainf, raw->snowf, raw->wind_e, raw->wind_n};
  int nraw = 0;
  for (int i = 0; i < NRAW; i++) nraw += raw_h[i] ? 1 : 0;
  const int NF = o->snow_step > 0 ? o->time_step / o->snow_step : 1;
  const size_t n_flag = (size_t)o->nrecs * C, n_out = n_flag * (NF > 1 ? NF + 1 : 1);
  double *dev = nullptr;
  CK(cudaMalloc((void **)&dev, ((size_t)NSITE * C + (size_t)nraw * nrf * C + VR_NFORCING * n_out) * sizeof(double) + n_flag * sizeof(int32_t)));
  const double *site_d[NSITE] = {}, *raw_d[NRAW] = {};
  double *p = dev;
  int rc = VR_OK;
  for (int i = 0; i < NSITE && rc == VR_OK; i++, p += C)
    if (site_h[i]) { site_d[i] = p; if (cudaMemcpy(p, site_h[i], C * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) rc = VR_ERR_CUDA; }
  for (int i = 0; i < NRAW && rc == VR_OK; i++)
    if (raw_h[i]) {
      raw_d[i] = p;
      if (cudaMemcpy(p, raw_h[i], (size_t)nrf * C * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) rc = VR_ERR_CUDA;
      p += (size_t)nrf * C;
    }
  if (rc != VR_OK) { cudaFree(dev); return fail(rc, "host to device copy failed"); }
  vr_atmos_cells cd{C, site_d[0], site_d[1], site_d[2], site_d[3], site_d[4], site_d[5], site_d[6], site_d[7], site_d[8], site_d[9]};
  vr_atmos_raw rd{nrf, raw_d[0], raw_d[1], raw_d[2], raw_d[3], raw_d[4], raw_d[5], raw_d[6], raw_d[7], raw_d[8], raw_d[9], raw_d[10],
                  raw_d[11], raw_d[12], raw_d[13], raw_d[14]};
  double *out_d[VR_NFORCING];
  for (int f = 0; f < VR_NFORCING; f++) out_d[f] = p + (size_t)f * n_out;
  int32_t *flag_d = (int32_t *)(p + (size_t)VR_NFORCING * n_out);
  rc = vr_atmos_prepare_sub_device(device, o, &cd, &rd, out_d, snowflag ? flag_d : nullptr, nullptr);
  if (rc == VR_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(VR_ERR_CUDA, std::string("kernel: ") + cudaGetErrorString(cudaGetLastError()));
  for (int f = 0; f < VR_NFORCING && rc == VR_OK; f++)
    if (cudaMemcpy(out[f], out_d[f], n_out * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(VR_ERR_CUDA, "device to host copy failed");
  if (rc == VR_OK && snowflag && cudaMemcpy(snowflag, flag_d, n_flag * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = fail(VR_ERR_CUDA, "device to host copy failed");
  cudaFree(dev);
  return rc;
}

int vr_atmos_diag_cr(int device, int64_t n, const double *x, double *acos_out, const double *y, double *cos_out)
{
  if (n < 1 || !x || !acos_out || !y || !cos_out) return fail(VR_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev)
    return fail(VR_ERR_CUDA, "no CUDA device: the forcing preparation has no CPU path");
  CK(cudaSetDevice(device));
  double *d = nullptr;
  CK(cudaMalloc((void **)&d, 4 * n * sizeof(double)));
  cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d + 2 * n, y, n * sizeof(double), cudaMemcpyHostToDevice);
  k_diag_cr<<<(unsigned)((n + 255) / 256), 256>>>(n, d, d + n, d + 2 * n, d + 3 * n);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(acos_out, d + n, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(cos_out, d + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(VR_ERR_CUDA, cudaGetErrorString(e));
  return VR_OK;
}

int vr_atmos_diag_solar(int device, int64_t n, const double *lat, const double *elevation, const double *lng,
                        const double *time_zone_lng, double *tab, int32_t exact_walk)
{
  if (n < 1 || !lat || !elevation || !tab) return fail(VR_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev)
    return fail(VR_ERR_CUDA, "no CUDA device: the forcing preparation has no CPU path");
  CK(cudaSetDevice(device));
  const size_t n_tab = (size_t)va::NYD * 28 * n;
  double *d = nullptr;
  CK(cudaMalloc((void **)&d, (n_tab + 4 * n) * sizeof(double)));
  CK(cudaMemset(d, 0, (n_tab + 4 * n) * sizeof(double)));
  double *site = d + n_tab;
  cudaMemcpy(site, lat, n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(site + n, elevation, n * sizeof(double), cudaMemcpyHostToDevice);
  if (lng) cudaMemcpy(site + 2 * n, lng, n * sizeof(double), cudaMemcpyHostToDevice);
  if (time_zone_lng) cudaMemcpy(site + 3 * n, time_zone_lng, n * sizeof(double), cudaMemcpyHostToDevice);
  va::Ctx x{};
  x.C = n; x.CB = n; x.c0 = 0;
  x.lat = site; x.elev = site + n; x.lng = site + 2 * n; x.tz = site + 3 * n;
  x.ttmax0 = d; x.flat = d + (size_t)va::NYD * n; x.slp = d + (size_t)2 * va::NYD * n; x.dayl = d + (size_t)3 * va::NYD * n;
  x.hf = d + (size_t)4 * va::NYD * n;
  if (exact_walk) k_solar_exact<<<(unsigned)(((n + TPB - 1) / TPB) * va::NYD), TPB>>>(x, n);
  else k_solar<<<(unsigned)(((n + TPB - 1) / TPB) * va::NYD), TPB>>>(x, n);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(tab, d, n_tab * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(VR_ERR_CUDA, cudaGetErrorString(e));
  return VR_OK;
}

int vr_atmos_last_times(double ms[3], int64_t *launches)
{
  if (!ms) return fail(VR_ERR_ARG, "null argument");
  ms[0] = ms[1] = ms[2] = 0.;
  if (launches) *launches = g_t.launches;
  if (g_t.ev.empty()) return VR_OK;
  CK(cudaSetDevice(g_t.device));
  CK(cudaEventSynchronize(g_t.ev.back()));
  for (size_t b = 0; b + 3 < g_t.ev.size(); b += 4)
    for (int k = 0; k < 3; k++) {
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, g_t.ev[b + k], g_t.ev[b + k + 1]));
      ms[k] += t;
    }
  return VR_OK;
}

}  // extern "C"
