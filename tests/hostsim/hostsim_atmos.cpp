// hostsim_atmos.cpp -- TEST HARNESS ONLY.  Compiles the device header of the forcing preparation
// (vicres_b200/csrc/vic_atmos.cuh) for the HOST with g++ and runs the bodies serially in the order the
// kernels of vicres_atmos.cu run them, so that the non-GPU test tier can check the restructured
// arithmetic (hourly radiation fractions without the tiny_radfract array, per-day radiation chains,
// array-free Hermite spline) against the reference's own atmos[] records.  Not linked into
// libvicres.so, not a product code path.
#include <vector>
#include "../../vicres_b200/csrc/vic_atmos.cuh"

extern "C" int hs_atmos(const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw, double *const out[VR_NFORCING],
                        int32_t *snowflag)
{
  if (va::unsupported(*o)) return -2;
  va::Ctx x{};
  va::fill_options(*o, x);
  const int64_t C = cells->ncell;
  const int nd = x.Ndays + 1;
  std::vector<int16_t> cal(x.Ndays + 2);
  va::build_calendar(*o, x.Ndays, cal.data());
  x.cal = cal.data();
  x.C = C; x.c0 = 0; x.CB = C; x.nrf = raw->nrec_forcing;
  x.lat = cells->lat; x.lng = cells->lng; x.tz = cells->time_zone_lng; x.elev = cells->elevation;
  x.slope = cells->slope; x.aspect = cells->aspect; x.ehoriz = cells->ehoriz; x.whoriz = cells->whoriz;
  x.r_prec = raw->prec; x.r_ta = raw->t_a; x.r_tb = raw->t_b; x.r_wind = raw->wind;
  x.r_sw = raw->shortwave; x.r_lw = raw->longwave; x.r_vp = raw->vp; x.r_pr = raw->pressure; x.r_de = raw->density;
  if (va::unsupported_raw(*o, *raw)) return -2;
  // derived columns, as vr_atmos_prepare_sub_device builds them with k_derive
  const size_t n_col = (size_t)raw->nrec_forcing * C;
  std::vector<double> d_prec, d_wind, d_vp;
  if (raw->rainf && raw->snowf) {
    d_prec.resize(n_col);
    for (size_t i = 0; i < n_col; i++) d_prec[i] = va::derive_prec(x, raw->rainf[i], raw->snowf[i]);
    x.r_prec = d_prec.data(); x.prec_native = 1;
  }
  if (raw->wind_e && raw->wind_n) {
    d_wind.resize(n_col);
    for (size_t i = 0; i < n_col; i++) d_wind[i] = va::derive_wind(raw->wind_e[i], raw->wind_n[i]);
    x.r_wind = d_wind.data(); x.has_wind = 1;
  }
  const bool vp_q = !raw->vp && raw->qair && raw->pressure;
  const bool vp_rh = !raw->vp && !vp_q && !raw->qair && raw->rel_humid && o->layout == VR_ATMOS_SUBDAILY_AIR_TEMP;
  if (vp_q || vp_rh) {
    d_vp.resize(n_col);
    for (size_t i = 0; i < n_col; i++)
      d_vp[i] = vp_q ? va::derive_vp_qair(x, raw->qair[i], raw->pressure[i]) : va::derive_vp_rh(x, raw->rel_humid[i], raw->t_a[i]);
    x.r_vp = d_vp.data(); x.vp_native = 1;
  }
  else if (!raw->vp) { if (raw->qair) x.r_q = raw->qair; else if (raw->rel_humid) x.r_rh = raw->rel_humid; }
  std::vector<double> tab((size_t)va::NYD * 28 * C), dly((size_t)nd * 7 * C), tcon(2 * C);
  std::vector<int8_t> hrs((size_t)nd * 2 * C);
  x.ttmax0 = tab.data(); x.flat = x.ttmax0 + va::NYD * C; x.slp = x.flat + va::NYD * C; x.dayl = x.slp + va::NYD * C;
  x.hf = x.dayl + va::NYD * C;
  double *d = dly.data();
  x.prec_d = d; x.swe = d + (size_t)nd * C; x.srad = d + (size_t)2 * nd * C; x.pva = d + (size_t)3 * nd * C;
  x.tskc = d + (size_t)4 * nd * C; x.tdew = d + (size_t)5 * nd * C; x.parr = d + (size_t)6 * nd * C;
  x.tcon = tcon.data();
  x.tminh = hrs.data(); x.tmaxh = hrs.data() + (size_t)nd * C;
  for (int f = 0; f < VR_NFORCING; f++) x.out[f] = out[f];
  x.min_tf = cells->min_Tfactor; x.snowflag = snowflag;
  for (int64_t c = 0; c < C; c++)
    for (int yd = 0; yd < va::NYD; yd++) va::solar_day(x, c, yd);
  for (int64_t c = 0; c < C; c++) va::daily_prep(x, c);
  for (int64_t c = 0; c < C; c++)
    for (int day = 0; day < nd; day++) va::daily_rad(x, c, day);
  if (x.vp_iter == VR_VP_ITER_CONVERGE)
    for (int64_t c = 0; c < C; c++) va::converge(x, c);
  for (int64_t c = 0; c < C; c++)
    for (int day = 0; day < nd; day++) va::minmax_hour(x, c, day);
  for (int64_t c = 0; c < C; c++)
    for (int day = 0; day < nd; day++) va::humid_day(x, c, day);
  for (int64_t c = 0; c < C; c++)
    for (int rec = 0; rec < x.nrecs; rec++) va::record(x, c, rec);
  return 0;
}

#include "../../vicres_b200/csrc/vic_cr_trig.cuh"
// correctly rounded acos / cos of vic_cr_trig.cuh, for the test against mpmath
extern "C" void hs_cr(int64_t n, const double *x, double *acos_out, const double *y, double *cos_out)
{
  for (int64_t i = 0; i < n; i++) { acos_out[i] = vcr::acos_cr(x[i]); cos_out[i] = vcr::cos_cr(y[i]); }
}
extern "C" void hs_cr_sin(int64_t n, const double *y, double *sin_out)
{
  for (int64_t i = 0; i < n; i++) sin_out[i] = vcr::sin_cr(y[i]);
}

// solar-geometry tables of n cells, layout of vr_atmos_diag_solar
extern "C" void hs_solar(int64_t n, const double *lat, const double *elev, const double *lng, const double *tz, double *tab)
{
  std::vector<double> zero(n, 0.0);
  va::Ctx x{};
  x.C = n; x.CB = n; x.c0 = 0;
  x.lat = lat; x.elev = elev; x.lng = lng ? lng : zero.data(); x.tz = tz ? tz : zero.data();
  x.ttmax0 = tab; x.flat = tab + (size_t)va::NYD * n; x.slp = tab + (size_t)2 * va::NYD * n; x.dayl = tab + (size_t)3 * va::NYD * n;
  x.hf = tab + (size_t)4 * va::NYD * n;
  for (int64_t c = 0; c < n; c++)
    for (int yd = 0; yd < va::NYD; yd++) va::solar_day(x, c, yd);
}
