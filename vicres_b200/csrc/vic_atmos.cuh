// vic_atmos.cuh -- device arithmetic of the forcing preparation (include/vicres_atmos.h).
//
// The reference prepares one cell at a time (initialize_atmos.c:137-1818 -> mtclim_wrapper.c:69-278 ->
// mtclim_vic.c:404-1395, calc_air_temperature.c, calc_longwave.c, svp.c).  Here the same arithmetic is cut along
// its data dependences into five bodies, each a function of (cell, day) / (cell, record) indices so that a kernel
// launches one thread per index pair:
//
//   solar_day      (cell, yearday)   mtclim_vic.c:789-921   the 30-second hour-angle walk: ttmax0, flat/slope potential
//                                    radiation, daylength and -- instead of the reference's 366 x 2880 tiny_radfract
//                                    array per cell -- the 24 hourly sums mtclim_to_vic (mtclim_wrapper.c:257-269) forms
//   daily_prep     (cell)            the only true recurrences over days: MTCLIM's snowpack (:474-530); daily
//                                    precipitation totals; Tmax/Tmin of the sub-daily layout
//   daily_rad      (cell, day)       :649-741, :956-1066, compute_srad_humidity_onetime :1098-1223 -- every day's
//                                    radiation/humidity chain is independent once the windows are read from the inputs
//   converge       (cell)            VP_ITER_CONVERGE only: the rmse-controlled passes (:1041-1057)
//   minmax_hour    (cell, day)       set_max_min_hour, calc_air_temperature.c:139-201
//   record         (cell, record)    initialize_atmos.c:597-1388: hourly Hermite temperatures (HourlyT), VP interpolation,
//                                    aggregation to the model step, pressure, density, VPD, longwave
//
// Every formula keeps the reference's operation order (-fmad=false), every sum its term order.  The header also
// compiles for the host (tests/hostsim, test harness only) with libm throughout: that build is bit-identical
// to the reference.  On the device pow(trans1, am) is exp(am * log(trans1)) with the logarithm taken once per
// (cell, yearday), pow(x, 1.5) is x*sqrt(x) and pow(x, 0.5) is sqrt(x): differences of a few ulp; the two
// evaluations a discrete decision hangs on (the sunrise hour angle and its cosine) are correctly rounded.
#ifndef VIC_ATMOS_CUH
#define VIC_ATMOS_CUH
#include <math.h>
#include <stdint.h>
#include "../../include/vicres_atmos.h"
#include "vic_cr_trig.cuh"

#if defined(__CUDACC__)
#define VA_HD __host__ __device__ __forceinline__
#define VA_HDN __host__ __device__ __noinline__
#else
#define VA_HD inline
#define VA_HDN inline
#endif

namespace va {

// vicNl_def.h:275-333, mtclim_constants_vic.h, mtclim_parameters_vic.h
constexpr double KELVIN = 273.15, STEFAN_B = 5.6696e-8, GRAV = 9.81, T_LAPSE = -0.0065;
constexpr int RD = 287, PS_PM = 101300;
constexpr double A_SVP = 0.61078, B_SVP = 17.269, C_SVP = 237.3, EPS_MW = 0.62196351;
constexpr double SECPERRAD = 13750.9871, RADPERDAY = 0.017214, RADPERDEG = 0.01745329, MINDECL = -0.4092797;
constexpr double DAYSOFF = 11.25, SRADDT = 30.0, MA = 28.9644e-3, R_GAS = 8.3143, G_STD = 9.80665, P_STD = 101325.0;
constexpr double T_STD = 288.15, CP_AIR = 1010.0, LR_STD = 0.0065, MT_PI = 3.1415927, TDAYCOEF = 0.45;
constexpr double SNOW_TCRIT = -6.0, SNOW_TRATE = 0.042, TBASE = 0.870, ABASE = -6.1e-5, C_RAD = 1.5;
constexpr double B0 = 0.031, B1 = 0.201, B2 = 0.185, RAIN_SCALAR = 0.75, DIF_ALB = 0.6;
constexpr int TINY = 2880, TINY_HOUR = 120, NYD = 365;

// Everything a body needs.  Column arrays of the caller are indexed [r*C + c] with the GLOBAL cell index c;
// scratch arrays hold one batch of cells and are indexed [row*CB + (c - c0)].
struct Ctx {
  // options
  int dt, nrecs, starthour, layout, fdt, has_wind, vp_interp, vp_iter, swe_corr, lw_type, lw_cloud, plapse, alma;
  double min_wind, sw_thresh;
  int Ndays, fsteps;
  int NF;                      // TIME_STEP / SNOW_STEP: windows per record (1: the record itself)
  double max_snow_temp;        // MAX_SNOW_TEMP, for atmos->snowflag (:1736-1753)
  const int16_t *cal;          // [Ndays+2] day_in_year of the days from one day BEFORE the start date on
  // cells
  int64_t C, c0, CB, nrf;
  const double *lat, *lng, *tz, *elev, *slope, *aspect, *ehoriz, *whoriz;
  const double *r_prec, *r_ta, *r_tb, *r_wind;
  const double *r_sw, *r_lw, *r_vp;   // optional observed SHORTWAVE, LONGWAVE, VP (kPa) columns, or null
  const double *r_pr, *r_de;          // optional observed PRESSURE (kPa), DENSITY columns, or null
  const double *r_q, *r_rh;           // daily QAIR / REL_HUMID that replace MTCLIM's daily vapour pressure afterwards (:1176-1215), or null
  int prec_native, vp_native;         // r_prec / r_vp are columns derived on the device, already in mm per step / Pa
  // scratch: solar tables [365][CB] and hourly fractions [365*24][CB]
  double *ttmax0, *flat, *slp, *dayl, *hf;
  // scratch: daily series [ndl][CB]
  double *prec_d, *swe, *srad, *pva, *tskc, *tdew, *parr;
  double *tcon;                // [2][CB] Tmax, Tmin of the sub-daily layout
  int8_t *tminh, *tmaxh;       // [ndl][CB]
  const int32_t *perm;         // [CB] k_solar only: batch-local cell indices ordered by latitude (or null)
  const double *min_tf;        // [C] the lowest band Tfactor of the cell (snowflag), or null = 0
  double *out[VR_NFORCING];    // NF == 1: [nrecs][C]; NF > 1: [nrecs][NF + 1][C]
  int32_t *snowflag;           // [nrecs][C] atmos->snowflag[NR], or null
};

struct Geo {                   // initialize_atmos.c:217-225, 271-292
  double hour_offset;
  int hoi, shift, early, ndl, tiny_off;
};

VA_HD Geo cell_geo(const Ctx &x, int64_t c)
{
  Geo g;
  double ho = (x.tz[c] - x.lng[c]) * 24 / 360;
  g.hoi = (ho < 0) ? (int)(ho - 0.5) : (int)(ho + 0.5);
  ho -= g.hoi;
  g.hour_offset = ho;
  g.early = (x.starthour - g.hoi < 0) ? 1 : 0;          // the local calendar starts one day before the model's
  g.shift = g.early ? 24 : 0;
  g.ndl = x.Ndays + (g.hoi != 0 ? 1 : 0);
  g.tiny_off = (int)((float)TINY_HOUR * ho);            // mtclim_wrapper.c:239
  return g;
}
VA_HD int yday_of(const Ctx &x, const Geo &g, int day) { return x.cal[day + 1 - g.early]; }
VA_HD int ydi_of(const Ctx &x, const Geo &g, int day) { int y = yday_of(x, g, day) - 1; return y > 364 ? 364 : y; }   // day 366 = day 365, mtclim_vic.c:924-932

VA_HD double fetch(const double *col, const Ctx &x, int64_t c, int64_t i)
{
  return (col && i >= 0 && i < x.nrf) ? col[i * x.C + c] : 0.0;
}
// local_forcing_data[type][idx], initialize_atmos.c:488-539
VA_HD double local_daily(const double *col, const Ctx &x, const Geo &g, int64_t c, int idx)
{
  int i = idx;
  if (g.hoi > 0) i--;
  if (i < 0) i = 0;
  if (i >= x.Ndays) i = x.Ndays - 1;
  return fetch(col, x, c, i);
}
VA_HD double local_hourly(const double *col, const Ctx &x, const Geo &g, int64_t c, int idx)
{
  int i = (idx - x.starthour + g.hoi) / x.fdt;
  if (i < 0) i += x.fsteps;
  if (i >= x.Ndays * x.fsteps) i -= x.fsteps;
  return fetch(col, x, c, i);
}
// unit conversion of the columns as read (initialize_atmos.c:374-416): ALMA rates -> amounts per forcing step and K -> C,
// else kPa -> Pa.  (x * 1.0 and x - 0.0 are exact, so the traditional units go through unchanged.)
VA_HD double u_prec_raw(const Ctx &x, double v) { return v * (x.alma ? (double)(x.fdt * 3600) : 1.0); }
VA_HD double u_prec(const Ctx &x, double v) { return x.prec_native ? v : u_prec_raw(x, v); }
VA_HD double u_temp(const Ctx &x, double v) { return v - (x.alma ? KELVIN : 0.0); }
VA_HD double u_kpa(const Ctx &x, double v) { return v * (x.alma ? 1.0 : 1000.); }

VA_HD double day_tmax(const Ctx &x, const Geo &g, int64_t c, int day)
{
  return x.fdt == 24 ? u_temp(x, local_daily(x.r_ta, x, g, c, day)) : x.tcon[c - x.c0];
}
VA_HD double day_tmin(const Ctx &x, const Geo &g, int64_t c, int day)
{
  return x.fdt == 24 ? u_temp(x, local_daily(x.r_tb, x, g, c, day)) : x.tcon[x.CB + c - x.c0];
}

// observed columns in local time: hour idx of the local hourly axis (initialize_atmos.c:488-539; VP in Pa, :404-414)
VA_HD double obs_hour(const double *col, const Ctx &x, const Geo &g, int64_t c, int idx)
{
  return x.fdt == 24 ? local_daily(col, x, g, c, idx / 24) : local_hourly(col, x, g, c, idx);
}
VA_HD double obs_vp_hour(const Ctx &x, const Geo &g, int64_t c, int idx)
{
  const double v = obs_hour(x.r_vp, x, g, c, idx);
  return x.vp_native ? v : u_kpa(x, v);
}

#if defined(__CUDA_ARCH__)
#define VA_POW15(x) ((x) * sqrt(x))
#define VA_SQRT_POW(x) sqrt(x)
#else
#define VA_POW15(x) pow((x), C_RAD)
#define VA_SQRT_POW(x) pow((x), 0.5)
#endif

VA_HD double svp(double temp)                          // svp.c:7-22
{
  double v = A_SVP * exp((B_SVP * temp) / (C_SVP + temp));
  if (temp < 0) v *= 1.0 + .00972 * temp + .000042 * temp * temp;
  return v * 1000.;
}

VA_HD double calc_pet(double rad, double ta, double pa, double dayl)   // mtclim_vic.c:1255-1313
{
  double rnet = rad * 0.72;
  double lhvap = 2.5023e6 - 2430.54 * ta;
  double gamma = CP_AIR * pa / (lhvap * EPS_MW);
  double dt = 0.2, t1 = ta + dt, t2 = ta - dt;
  double s = (svp(t1) - svp(t2)) / (t1 - t2);
  double pet = (1.26 * (s / (s + gamma)) * rnet * dayl) / lhvap;
  return pet / 10.0;
}

VA_HD double atm_pres(double elev)                     // mtclim_vic.c:1316-1334
{
  double t1 = 1.0 - (LR_STD * elev) / T_STD;
  double t2 = G_STD / (LR_STD * (R_GAS / MA));
  return P_STD * pow(t1, t2);
}

VA_HD double longwave(const Ctx &x, double tskc, double air_temp, double vp)   // calc_longwave.c:7-76
{
  double ec = 0., e, t;
  air_temp += KELVIN;
  vp /= 100;
  if (x.lw_type == VR_LW_TVA) ec = 0.740 + 0.0049 * vp;
  else if (x.lw_type == VR_LW_ANDERSON) ec = 0.68 + 0.036 * VA_SQRT_POW(vp);
  else if (x.lw_type == VR_LW_BRUTSAERT) { t = vp / air_temp; ec = 1.24 * pow(t, 0.14285714); }
  else if (x.lw_type == VR_LW_SATTERLUND) ec = 1.08 * (1 - exp(-1 * pow(vp, (air_temp / 2016))));
  else if (x.lw_type == VR_LW_IDSO) ec = 0.7 + 5.95e-5 * vp * exp(1500 / air_temp);
  else if (x.lw_type == VR_LW_PRATA) { t = 46.5 * vp / air_temp; ec = 1 - (1 + t) * exp(-1 * VA_SQRT_POW((1.2 + 3 * t))); }
  if (x.lw_cloud == VR_LW_CLOUD_DEARDORFF) e = tskc * 1.0 + (1 - tskc) * ec;
  else e = (1.0 + 0.17 * tskc * tskc) * ec;
  return e * STEFAN_B * air_temp * air_temp * air_temp * air_temp;
}

// ---------------------------------------------------------------------------------------------------------------
// solar_day: one (cell, yearday) of the loop mtclim_vic.c:789-921.
// The reference fills tiny_radfract[yearday][2880] (last write wins where two hour angles fall into the same
// 30-second slot), divides every slot by the day's total and later (mtclim_wrapper.c:257-269) adds 120 consecutive
// slots, shifted by the cell's sub-hour offset, per clock hour.  The same quotients are added here in the same order
// without the array: the hour-angle walk is repeated once the total is known, slot values are committed when the
// slot index moves on, and the one clock hour that wraps around midnight gets its late slots first, then -- by a
// short third walk -- its early ones.
struct SolarGeom {
  double cosegeom, sinegeom, bsg1, bsg2, bsg3, hss, dir_beam_topa, dh, coszeh, coszwh;
};
#if defined(__CUDA_ARCH__)
// the same integer from a multiplication, falling back to the division when the quotient is within 1e-9 of an integer
__device__ __forceinline__ int tiny_slot_fast(double h)
{
  const double num = 12L * 3600L + h * SECPERRAD;
  double q = num * (1.0 / SRADDT);
  if (fabs(q - rint(q)) < 1e-9) q = num / SRADDT;
  int t = (int)q;
  if (t < 0) t = 0;
  if (t > TINY - 1) t = TINY - 1;
  return t;
}
#endif
VA_HD int tiny_slot(double h)
{
  int t = (int)((12L * 3600L + h * SECPERRAD) / SRADDT);
  if (t < 0) t = 0;
  if (t > TINY - 1) t = TINY - 1;
  return t;
}

// cos((70+k) * RADPERDEG), k = 0..19: the zenith angles at which the optical air mass table (mtclim_vic.c:851-859)
// steps to its next entry
#define VA_ZENITH_STEPS {0.34202030908371178, 0.32556832362552834, 0.30901716693075376, 0.29237188064059616, \
  0.27563753506904187, 0.25881922765838938, 0.24192208142651928, 0.22495124340637382, 0.20791188307812145, \
  0.19080919079448339, 0.17364837619970241, 0.1564346666426352, 0.13917330558445171, 0.12186955100142652, \
  0.10452867378330888, 0.087155956127758991, 0.069756689931339946, 0.052336175177555268, 0.03489971832242298, \
  0.017452630678078236}

// FAST (device only): same walk with three changes that stay far inside the 1e-9 bar and touch no decision --
//  * cos/sin of the hour angle by rotation through the constant step, re-anchored with sincos every 16 steps
//    (error < 1e-14) instead of two libm calls per step;
//  * the air-mass table entry from comparisons of cos(zenith) with the 20 table angles instead of acos per step,
//    and the transmittance of the 21 table entries from a per-thread table (`t21`, stride `ts21`) instead of pow;
//  * the slot values multiplied by 1/total instead of divided.
// ck: the 20 zenith steps (shared memory).
template <bool FAST>
VA_HDN void solar_day_t(const Ctx &x, int64_t c, int yd, const double *ck, double *t21, int ts21)
{
  const double optam[21] = {2.90, 3.05, 3.21, 3.39, 3.69, 3.82, 4.07, 4.37, 4.72, 5.12, 5.60,
                            6.18, 6.88, 7.77, 8.90, 10.39, 12.44, 15.36, 19.79, 26.96, 30.00};
  const int64_t lc = c - x.c0, CB = x.CB;
  const Geo g = cell_geo(x, c);
  const double site_slp = x.slope ? x.slope[c] : 0., site_asp = x.aspect ? x.aspect[c] : 0.;
  const double site_eh = x.ehoriz ? x.ehoriz[c] : 0., site_wh = x.whoriz ? x.whoriz[c] : 0.;
  // :754-786 (per cell; repeated per yearday thread, a few dozen operations against ~10^5 in the walk)
  double t1 = 1.0 - (LR_STD * x.elev[c]) / T_STD;
  double t2 = G_STD / (LR_STD * (R_GAS / MA));
  double pratio = pow(t1, t2);
  double trans1 = pow(TBASE, pratio);
  double lat = x.lat[c];
  lat *= RADPERDEG;
  if (lat > 1.5707) lat = 1.5707;
  if (lat < -1.5707) lat = -1.5707;
  // On the device every trigonometric value the sunrise hour angle is built from, the angle itself and the cosine
  // of the walk's first step are evaluated correctly rounded (vic_cr_trig.cuh): cos(zenith) at that first step is
  // zero up to rounding and its sign decides the hour of sunrise, so the device has to round like glibc does.
#if defined(__CUDA_ARCH__) || defined(VA_FORCE_CR)
#define VA_COS(x) vcr::cos_cr(x)
#define VA_SIN(x) vcr::sin_cr(x)
#define VA_ACOS(x) vcr::acos_cr(x)
#else
#define VA_COS(x) cos(x)
#define VA_SIN(x) sin(x)
#define VA_ACOS(x) acos(x)
#endif
  const double coslat = VA_COS(lat), sinlat = VA_SIN(lat);
  const double cosslp = cos(site_slp * RADPERDEG), sinslp = sin(site_slp * RADPERDEG);
  const double cosasp = cos(site_asp * RADPERDEG), sinasp = sin(site_asp * RADPERDEG);
  const double coszeh = cos(1.570796 - (site_eh * RADPERDEG)), coszwh = cos(1.570796 - (site_wh * RADPERDEG));
  const double dt = SRADDT, dh = dt / SECPERRAD;
#if defined(__CUDA_ARCH__)
  const double ln_trans1 = log(trans1);
#endif
  // :791-822
  double decl = MINDECL * VA_COS(((double)yd + DAYSOFF) * RADPERDAY);
  double cosdecl = VA_COS(decl), sindecl = VA_SIN(decl);
  double bsg1 = -sinslp * sinasp * cosdecl;
  double bsg2 = (-cosasp * sinslp * sinlat + cosslp * coslat) * cosdecl;
  double bsg3 = (cosasp * sinslp * coslat + cosslp * sinlat) * sindecl;
  double cosegeom = coslat * cosdecl, sinegeom = sinlat * sindecl;
  double coshss = -(sinegeom) / cosegeom;
  if (coshss < -1.0) coshss = -1.0;
  if (coshss > 1.0) coshss = 1.0;
  const double hss = VA_ACOS(coshss);
  const double cos_h0 = VA_COS(hss);
#define VA_COS_STEP(h) ((h) == -hss ? cos_h0 : cos(h))
  double daylength = 2.0 * hss * SECPERRAD;
  if (daylength > 86400) daylength = 86400;
  double sc = 1368.0 + 45.5 * sin((2.0 * MT_PI * (double)yd / 365.25) + 1.7);
  const double dir_beam_topa = sc * dt;
  double sum_trans = 0.0, sum_flat = 0.0, sum_slope = 0.0;

  // first walk, :828-899 without the tiny_radfract stores
#if defined(__CUDA_ARCH__)
  const double cd = cos(dh), sd = sin(dh);
  if (FAST) for (int k = 0; k < 21; k++) t21[k * ts21] = exp(optam[k] * ln_trans1);
#define VA_STEP_TRIG(n, h, cs, sn)                                                          \
  if (!FAST || ((n) & 15) == 0) { sincos((h), &sn, &cs); if ((n) == 0) cs = cos_h0; }        \
  else { double c2_ = __fma_rn(cs, cd, -(sn * sd)); sn = __fma_rn(sn, cd, cs * sd); cs = c2_; }
#else
#define VA_STEP_TRIG(n, h, cs, sn) { cs = VA_COS_STEP(h); sn = sin(h); }
#endif
  {
    double sinh_ = 0., cosh_ = 1.;
    int n = 0, ami = 0;
    for (double h = -hss; h < hss; h += dh, n++) {
      VA_STEP_TRIG(n, h, cosh_, sinh_)
      double cza = cosegeom * cosh_ + sinegeom;
      double cbsa = sinh_ * bsg1 + cosh_ * bsg2 + bsg3;
      if (cza > 0.0) {
        double dir_flat_topa = dir_beam_topa * cza;
        double trans2;
#if defined(__CUDA_ARCH__)
        if (FAST) {
          const double t = cza + 0.0000001;
          if (t * 2.9 < 1.0) {                             // am > 2.9; both branches give trans1^2.9 at the seam
            while (ami > 0 && cza > ck[ami - 1]) ami--;
            while (ami < 20 && cza <= ck[ami]) ami++;
            trans2 = t21[ami * ts21];
          }
          else trans2 = exp((1.0 / t) * ln_trans1);
        }
        else
#endif
        {
          double am = 1.0 / (cza + 0.0000001);
          if (am > 2.9) {
            int ai = (int)(acos(cza) / RADPERDEG) - 69;
            if (ai < 0) ai = 0;
            if (ai > 20) ai = 20;
            am = optam[ai];
          }
#if defined(__CUDA_ARCH__)
          trans2 = exp(am * ln_trans1);
#else
          trans2 = pow(trans1, am);
#endif
        }
        sum_trans += trans2 * dir_flat_topa;
        sum_flat += dir_flat_topa;
        if ((h < 0.0 && cza > coszeh && cbsa > 0.0) || (h >= 0.0 && cza > coszwh && cbsa > 0.0))
          sum_slope += dir_beam_topa * cbsa;
      }
    }
  }
  if (daylength) {                                        // :910-919
    x.ttmax0[yd * CB + lc] = sum_trans / sum_flat;
    x.flat[yd * CB + lc] = sum_flat / daylength;
    x.slp[yd * CB + lc] = sum_slope / daylength;
  }
  else { x.ttmax0[yd * CB + lc] = 0.0; x.flat[yd * CB + lc] = 0.0; x.slp[yd * CB + lc] = 0.0; }
  x.dayl[yd * CB + lc] = daylength;

  // second walk: hourly sums of the normalised slot values
  const bool norm = daylength && sum_flat > 0;            // :902-905
  const int off = g.tiny_off;
  const int E = ((TINY_HOUR - off) % TINY_HOUR + TINY_HOUR) % TINY_HOUR;   // early slots [0, E) belong to the wrapping hour
  const int jw = off > 0 ? 0 : 23;
  double *hf = x.hf + (int64_t)yd * 24 * CB + lc;
  for (int j = 0; j < 24; j++) hf[(int64_t)j * CB] = 0.0;
  int prev = -1, cur_j = -1, hour_end = -1;             // hour_end: first slot after the current clock hour
  double pend = 0.0, acc = 0.0;
#if defined(__CUDA_ARCH__)
  const double inv_flat = 1.0 / sum_flat;
  auto share = [&](double v) { return norm ? (FAST ? v * inv_flat : v / sum_flat) : v; };
#else
  auto share = [&](double v) { return norm ? v / sum_flat : v; };
#endif
  auto commit = [&](int ts, double v) {
    if (ts < E) return;
    if (ts >= hour_end) {
      const int w = ((ts + off) % TINY + TINY) % TINY;    // position in the clock day
      const int j = w / TINY_HOUR;
      if (cur_j >= 0) hf[(int64_t)cur_j * CB] = acc;
      cur_j = j;
      acc = 0.0;
      hour_end = ts + (TINY_HOUR - (w - j * TINY_HOUR));
    }
    acc += share(v);
  };
  double sinh_ = 0., cosh_ = 1.;
  int n = 0;
  for (double h = -hss; h < hss; h += dh, n++) {
    VA_STEP_TRIG(n, h, cosh_, sinh_)
    double cza = cosegeom * cosh_ + sinegeom;
    double v = 0.0;
    if (cza > 0.0) { double dft = dir_beam_topa * cza; if (dft > 0) v = dft; }
#if defined(__CUDA_ARCH__)
    int ts = FAST ? tiny_slot_fast(h) : tiny_slot(h);
#else
    int ts = tiny_slot(h);
#endif
    if (ts != prev) { if (prev >= 0) commit(prev, pend); prev = ts; }
    pend = v;
  }
  if (prev >= 0) commit(prev, pend);
  if (E > 0) {
    if (cur_j != jw) { if (cur_j >= 0) hf[(int64_t)cur_j * CB] = acc; cur_j = jw; acc = 0.0; }
    prev = -1; pend = 0.0; n = 0;
    for (double h = -hss; h < hss; h += dh, n++) {
      int ts = tiny_slot(h);
      if (ts != prev) {
        if (prev >= 0) acc += share(pend);
        prev = ts;
        if (ts >= E) { prev = -1; break; }
      }
      VA_STEP_TRIG(n, h, cosh_, sinh_)
      double cza = cosegeom * cosh_ + sinegeom;
      double v = 0.0;
      if (cza > 0.0) { double dft = dir_beam_topa * cza; if (dft > 0) v = dft; }
      pend = v;
    }
    if (prev >= 0 && prev < E) acc += share(pend);
  }
  if (cur_j >= 0) hf[(int64_t)cur_j * CB] = acc;
#undef VA_STEP_TRIG
#undef VA_COS_STEP
#undef VA_COS
#undef VA_SIN
#undef VA_ACOS
}
VA_HD void solar_day(const Ctx &x, int64_t c, int yd) { solar_day_t<false>(x, c, yd, nullptr, nullptr, 0); }

// ---------------------------------------------------------------------------------------------------------------
// daily_prep: per cell, sequential over the local days.
VA_HD double mt_prcp(const Ctx &x, int64_t lc, int day) { return x.prec_d[(int64_t)day * x.CB + lc] / 10.; }   // cm, mtclim_wrapper.c:206
// calc_tair with site_elev == base_elev (mtclim_vic.c:415-425): dz = 0, lapse rate 6.5
VA_HD double site_t(double t) { return t + (0.0 * 6.5); }

VA_HDN void daily_prep(const Ctx &x, int64_t c)
{
  const int64_t lc = c - x.c0, CB = x.CB;
  const Geo g = cell_geo(x, c);
  if (x.fdt != 24) {
    // initialize_atmos.c:758-766 -- the reference reads the FIRST local day's hours for every day
    double tmax = -9999, tmin = -9999;
    for (int hour = 0; hour < 24; hour++) {
      double t = u_temp(x, local_hourly(x.r_ta, x, g, c, hour));
      if (hour >= 9 && (tmax == -9999 || t > tmax)) tmax = t;
      if (hour < 12 && (tmin == -9999 || t < tmin)) tmin = t;
    }
    x.tcon[lc] = tmax; x.tcon[CB + lc] = tmin;
  }
  for (int day = 0; day < g.ndl; day++) {                  // :597-635
    double p;
    if (x.fdt == 24) p = u_prec(x, local_daily(x.r_prec, x, g, c, day));
    else {
      p = 0;
      for (int hour = 0; hour < 24; hour++) p += u_prec(x, local_hourly(x.r_prec, x, g, c, day * 24 + hour)) / x.fdt;
    }
    x.prec_d[(int64_t)day * CB + lc] = p;
  }
  if (x.r_sw)                                             // mtclim_wrapper.c:196-202: 24-hour mean of the observations
    for (int day = 0; day < g.ndl; day++) {
      double s = 0;
      for (int j = 0; j < 24; j++) s += obs_hour(x.r_sw, x, g, c, day * 24 + j);
      x.srad[(int64_t)day * CB + lc] = s / 24;
    }
  if (x.r_vp)                                             // initialize_atmos.c:873-898
    for (int day = 0; day < g.ndl; day++) {
      double v;
      if (x.fdt == 24) v = obs_vp_hour(x, g, c, day * 24);
      else {
        v = 0;
        for (int j = 0; j < 24; j++) v += obs_vp_hour(x, g, c, day * 24 + j);
        v /= 24;
      }
      x.pva[(int64_t)day * CB + lc] = v;
    }
  if (!x.swe_corr) return;
  // snowpack(), mtclim_vic.c:474-530; s_prcp = prcp * 1 (calc_prcp with site_isoh == base_isoh)
  double snowpack = 0.0, sum = 0.0;
  int count = 0;
  const int start_yday = yday_of(x, g, 0), prev_yday = (start_yday == 1) ? 365 : start_yday - 1;
  for (int pass = 0; pass < 2; pass++) {
    for (int i = 0; i < g.ndl; i++) {
      double newsnow = 0.0, snowmelt = 0.0, stmin = site_t(day_tmin(x, g, c, i));
      if (stmin <= SNOW_TCRIT) newsnow = mt_prcp(x, lc, i) * 1.0;
      else snowmelt = SNOW_TRATE * (stmin - SNOW_TCRIT);
      snowpack += newsnow - snowmelt;
      if (snowpack < 0.0) snowpack = 0.0;
      x.swe[(int64_t)i * CB + lc] = snowpack;
      if (pass == 0 && i >= 1) {
        int y = yday_of(x, g, i);
        if (y == start_yday || y == prev_yday) { count++; sum += snowpack; }
      }
    }
    if (!count) break;
    snowpack = sum / (double)count;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// compute_srad_humidity_onetime for one day, mtclim_vic.c:1117-1219
struct DayRad { double srad, tdew, pva, tfmax; };
// srad_obs: the observed 24-hour mean when SHORTWAVE is supplied (ctrl->insw), else unused
VA_HD DayRad rad_pass(const Ctx &x, int64_t lc, int ydi, double pva_in, double tfmax, double swe, double sky_prop,
                      double stmin, double stday, double parray, double pa, double dtr, double srad_obs = 0.)
{
  const int64_t k = (int64_t)ydi * x.CB + lc;
  const double daylength = x.dayl[k];
  double t_tmax = x.ttmax0[k] + ABASE * pva_in;
  if (t_tmax < 0.0001) t_tmax = 0.0001;
  double t_final = t_tmax * tfmax;
  double pdif = -1.25 * t_final + 1.25;
  if (pdif > 1.0) pdif = 1.0;
  if (pdif < 0.0) pdif = 0.0;
  double pdir = 1.0 - pdif;
  double srad1 = x.slp[k] * t_final * pdir;
  double srad2 = x.flat[k] * t_final * pdif * (sky_prop + DIF_ALB * (1.0 - sky_prop));
  double sc;
  if (x.swe_corr && swe > 0.0) {
    sc = (1.32 + 0.096 * swe) * 1e6;
    if (daylength > 0.0) sc /= daylength;
    else sc = 0.0;
    if (sc > 100.0) sc = 100.0;
  }
  else sc = 0.0;
  DayRad r;
  r.tfmax = tfmax;
  if (x.r_sw) {                                           // :1173-1182: cloud transmittance from the observation
    double potrad = (srad1 + srad2 + sc) * daylength / t_final / 86400;
    if (potrad > 0 && srad_obs > 0 && daylength > 0) {
      r.tfmax = (srad_obs) / (potrad * t_tmax);
      if (r.tfmax > 1.0) r.tfmax = 1.0;
    }
    else r.tfmax = 1.0;
    r.srad = srad_obs;
  }
  else r.srad = srad1 + srad2 + sc;
  double tmink = stmin + KELVIN;
  double pet = calc_pet(r.srad, stday, pa, daylength);
  double ratio = pet / parray;
  double ratio2 = ratio * ratio;
  double ratio3 = ratio2 * ratio;
  double tdewk = tmink * (-0.127 + 1.121 * (1.003 - 1.444 * ratio + 12.312 * ratio2 - 32.766 * ratio3) + 0.0006 * (dtr));
  r.tdew = tdewk - KELVIN;
  r.pva = svp(r.tdew);
  return r;
}

VA_HD double sky_proportion(const Ctx &x, int64_t c)          // mtclim_vic.c:938-951
{
  const double site_slp = x.slope ? x.slope[c] : 0., site_eh = x.ehoriz ? x.ehoriz[c] : 0., site_wh = x.whoriz ? x.whoriz[c] : 0.;
  double avg_horizon = (site_eh + site_wh) / 2.0;
  double horizon_scalar = 1.0 - sin(avg_horizon * RADPERDEG);
  double slope_excess = (site_slp > avg_horizon) ? site_slp - avg_horizon : 0.0;
  double slope_scalar;
  if (2.0 * avg_horizon > 180.0) slope_scalar = 0.0;
  else {
    slope_scalar = 1.0 - (slope_excess / (180.0 - 2.0 * avg_horizon));
    if (slope_scalar < 0.0) slope_scalar = 0.0;
  }
  return horizon_scalar * slope_scalar;
}

VA_HD double day_dtr(const Ctx &x, const Geo &g, int64_t c, int day)   // :649-655
{
  double tmax = day_tmax(x, g, c, day), tmin = day_tmin(x, g, c, day);
  if (tmax < tmin) tmax = tmin;
  return tmax - tmin;
}

// One (cell, day): the windows, t_fmax and the passes that do not need a reduction over days.
//   VP_ITER_NONE, VP_ITER_ANNUAL: one pass.  (ANNUAL either keeps the first pass or -- PET/P >= 2.5 -- resets the
//     humidity to its initial value and repeats that same pass, :1018-1057: the results are those of one pass.)
//   VP_ITER_ALWAYS: two passes.   VP_ITER_CONVERGE: one pass here, the rest in converge().
// window[k] of the 90-day precipitation window (:716-725) as an index into the daily series
VA_HD int prcp_window_src(int k, int ndays, int isloop) { return k < 90 ? (isloop ? ndays - 90 + k : k) : k - 90; }
VA_HD int prcp_isloop(const Ctx &x, const Geo &g)          // :706-713
{
  const int start_yday = yday_of(x, g, 0), end_yday = yday_of(x, g, g.ndl - 1);
  if (start_yday != 1) return (end_yday == start_yday - 1) ? 1 : 0;
  return (end_yday == 365 || end_yday == 366) ? 1 : 0;
}

// dtr_of(d): diurnal range of local day d; win_of(k): entry k of the precipitation window (cm).  The host build and
// k_rad read them from the inputs; k_rad_tiled reads them from the shared-memory tile of its 32 cells x 32 days.
template <class DTR, class WIN>
VA_HD void daily_rad_t(const Ctx &x, int64_t c, int day, const Geo &g, DTR dtr_of, WIN win_of)
{
  const int64_t lc = c - x.c0, CB = x.CB;
  const int ndays = g.ndl;
  // 30-day pulled boxcar of the diurnal range (:658-669, pulled_boxcar :1338-1394)
  const int w = ndays >= 30 ? 30 : ndays;
  double sum_wt = 0.0;
  for (int i = 0; i < w; i++) sum_wt += 1.0;
  const int e = day < w - 1 ? w - 1 : day;
  double total = 0.0;
  for (int j = 0; j < w; j++) total += dtr_of(e - w + j + 1) * 1.0;
  const double sm_dtr = total / sum_wt;
  const double dtr = dtr_of(day);
  // effective annual precipitation (:684-741)
  double parray;
  if (ndays < 90) {
    double s = 0.0;
    for (int i = 0; i < ndays; i++) s += mt_prcp(x, lc, i) * 1.0;
    parray = (s / (double)ndays) * 365.25;
    if (parray < 8.0) parray = 8.0;
  }
  else {
    double s = 0.0;
    for (int j = 0; j < 90; j++) s += win_of(day + j) * 1.0;
    s = (s / 90.0) * 365.25;
    parray = (s < 8.0) ? 8.0 : s;
  }
  // :956-967
  double b = B0 + B1 * exp(-B2 * sm_dtr);
  double tfmax = 1.0 - 0.9 * exp(-b * VA_POW15(dtr));
  if (mt_prcp(x, lc, day) > x.sw_thresh) tfmax *= RAIN_SCALAR;
  // :976-1005
  const double tmax = day_tmax(x, g, c, day), tmin = day_tmin(x, g, c, day);
  const double stmax = site_t(tmax), stmin = site_t(tmin);
  const double tmean = (stmax + stmin) / 2.0;
  const double stday = ((stmax - tmean) * TDAYCOEF) + tmean;
  const double pa = atm_pres(x.elev[c]);
  const double sky_prop = sky_proportion(x, c);
  const double swe = x.swe_corr ? x.swe[(int64_t)day * CB + lc] : 0.0;
  const int ydi = ydi_of(x, g, day);
  const int64_t k = (int64_t)day * CB + lc;
  const double srad_obs = x.r_sw ? x.srad[k] : 0., vp_obs = x.r_vp ? x.pva[k] : 0.;   // written by daily_prep
  DayRad r = rad_pass(x, lc, ydi, x.r_vp ? vp_obs : svp(stmin), tfmax, swe, sky_prop, stmin, stday, parray, pa, dtr, srad_obs);
  // observed humidity is restored after the first pass (:1018-1023); the second pass sees the transmittance the first
  // one derived from observed radiation
  if (x.vp_iter == VR_VP_ITER_ALWAYS)
    r = rad_pass(x, lc, ydi, x.r_vp ? vp_obs : r.pva, r.tfmax, swe, sky_prop, stmin, stday, parray, pa, dtr, srad_obs);
  x.srad[k] = r.srad;
  x.pva[k] = x.r_vp ? vp_obs : r.pva;                     // :1060-1066
  x.tskc[k] = r.tfmax;                                    // cloud transmittance; converted to tskc where it is read
  if (x.vp_iter == VR_VP_ITER_CONVERGE) { x.tdew[k] = r.tdew; x.parr[k] = parray; }
}

VA_HDN void daily_rad(const Ctx &x, int64_t c, int day)
{
  const int64_t lc = c - x.c0;
  const Geo g = cell_geo(x, c);
  if (day >= g.ndl) return;
  const int isloop = g.ndl >= 90 ? prcp_isloop(x, g) : 0;
  daily_rad_t(x, c, day, g, [&](int d) { return day_dtr(x, g, c, d); },
              [&](int k) { return mt_prcp(x, lc, prcp_window_src(k, g.ndl, isloop)); });
}

#if defined(__CUDACC__)
// One block = RT_C cells x RT_D consecutive local days.  Every thread needs the 90 window entries and the 30 ranges
// that end at its day; neighbouring days share all but one of them, so the block stages the RT_D+89 window entries and
// RT_D+29 ranges of each of its cells in shared memory once (global reads per thread 150 -> 8) and every thread adds
// its terms from there in the reference's order.
constexpr int RT_C = 32, RT_D = 32;
__device__ __forceinline__ void daily_rad_tile(const Ctx &x, int64_t cb, int64_t tile_c, int tile_d)
{
  __shared__ double s_win[RT_D + 89][RT_C];
  __shared__ double s_dtr[RT_D + 29][RT_C];
  const int tx = threadIdx.x % RT_C, ty = threadIdx.x / RT_C;
  const int64_t lc = tile_c * RT_C + tx;
  const bool valid = lc < cb;
  const int64_t c = x.c0 + (valid ? lc : 0);
  const Geo g = cell_geo(x, c);
  const int D0 = tile_d * RT_D, ndl = g.ndl;
  const int isloop = (valid && ndl >= 90) ? prcp_isloop(x, g) : 0;
  for (int kk = ty; kk < RT_D + 89; kk += RT_D) {
    const int src = prcp_window_src(D0 + kk, ndl, isloop);
    s_win[kk][tx] = (valid && ndl >= 90 && src >= 0 && src < ndl) ? mt_prcp(x, c - x.c0, src) : 0.0;
  }
  for (int rr = ty; rr < RT_D + 29; rr += RT_D) {
    const int r = D0 - 29 + rr;
    s_dtr[rr][tx] = (valid && r >= 0 && r < ndl) ? day_dtr(x, g, c, r) : 0.0;
  }
  __syncthreads();
  const int day = D0 + ty;
  if (!valid || day >= ndl) return;
  daily_rad_t(x, c, day, g, [&](int d) { return s_dtr[d - (D0 - 29)][tx]; }, [&](int k) { return s_win[k - D0][tx]; });
}
#endif

// VP_ITER_CONVERGE: :1041-1057 after the first pass; one cell, all days per pass, rmse in day order.
VA_HDN void converge(const Ctx &x, int64_t c)
{
  const int64_t lc = c - x.c0, CB = x.CB;
  const Geo g = cell_geo(x, c);
  const int ndays = g.ndl;
  const double pa = atm_pres(x.elev[c]), sky_prop = sky_proportion(x, c);
  const double tol = 0.01;
  double rmse = tol + 1;
  int iter = 1;
  while (rmse > tol && iter < 100) {
    rmse = 0;
    for (int i = 0; i < ndays; i++) {
      const int64_t k = (int64_t)i * CB + lc;
      const double tmax = day_tmax(x, g, c, i), tmin = day_tmin(x, g, c, i);
      const double stmax = site_t(tmax), stmin = site_t(tmin);
      const double tmean = (stmax + stmin) / 2.0;
      const double stday = ((stmax - tmean) * TDAYCOEF) + tmean;
      const double save = x.tdew[k];
      DayRad r = rad_pass(x, lc, ydi_of(x, g, i), x.pva[k], x.tskc[k], x.swe_corr ? x.swe[k] : 0.0, sky_prop, stmin, stday,
                          x.parr[k], pa, day_dtr(x, g, c, i));
      x.srad[k] = r.srad; x.pva[k] = r.pva; x.tdew[k] = r.tdew;
      rmse += (r.tdew - save) * (r.tdew - save);
    }
    rmse /= ndays;
    rmse = VA_SQRT_POW(rmse);
    iter++;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// hourlyrad[day*24 + j], mtclim_wrapper.c:252-269 (MTCLIM-estimated radiation: daylight average x daylength)
VA_HD double hourly_rad(const Ctx &x, const Geo &g, int64_t lc, int day, int j)
{
  if (x.r_sw && x.fdt < 24) return local_hourly(x.r_sw, x, g, x.c0 + lc, day * 24 + j);   // initialize_atmos.c:966-972
  const int ydi = ydi_of(x, g, day);
  const double srad = x.srad[(int64_t)day * x.CB + lc];
  double tmp_rad = x.r_sw ? srad * 24. : srad * x.dayl[(int64_t)ydi * x.CB + lc] / 3600.;   // mtclim_wrapper.c:252-257
  return x.hf[((int64_t)ydi * 24 + j) * x.CB + lc] * tmp_rad;
}

VA_HDN void minmax_hour(const Ctx &x, int64_t c, int day)      // set_max_min_hour, calc_air_temperature.c:139-201
{
  const int64_t lc = c - x.c0;
  const Geo g = cell_geo(x, c);
  if (day >= g.ndl) return;
  int risehour = -999, sethour = -999;
  double prev = day > 0 ? hourly_rad(x, g, lc, day - 1, 23) : 0.0;
  for (int hour = 0; hour < 24; hour++) {
    double r = hourly_rad(x, g, lc, day, hour);
    if (hour < 12) { if (r > 0 && ((day == 0 && hour == 0) || prev <= 0)) risehour = hour; }
    else if (r <= 0 && prev > 0) sethour = hour;
    prev = r;
  }
  if (day == g.ndl - 1 && sethour == -999) sethour = 23;
  int tmaxhour, tminhour;
  if (risehour >= 0 && sethour >= 0) {
    tmaxhour = (int)(0.67 * (sethour - risehour) + risehour);
    tminhour = risehour - 1;
  }
  else { tminhour = 2; tmaxhour = 14; }
  x.tminh[(int64_t)day * x.CB + lc] = (int8_t)tminhour;
  x.tmaxh[(int64_t)day * x.CB + lc] = (int8_t)tmaxhour;
}

// ---------------------------------------------------------------------------------------------------------------
// HourlyT's spline nodes (calc_air_temperature.c:91-117) without the arrays: node k of n = 2*ndays + 2
VA_HD double node_x(const Ctx &x, int64_t lc, int ndays, int k)
{
  const int n = 2 * ndays + 2;
  int tie = 0;                                            // the two tie-down nodes, :109-113
  if (k == 0) { k = 2; tie = -24; }
  else if (k == n - 1) { k = n - 3; tie = 24; }
  const int i = (k - 1) >> 1, second = (k - 1) & 1;
  const int tn = x.tminh[(int64_t)i * x.CB + lc], tx = x.tmaxh[(int64_t)i * x.CB + lc];
  const int first_is_min = tn < tx;
  const int hr = (first_is_min != second) ? tn : tx;
  return (double)(hr + i * 24) + tie;
}
VA_HD double node_y(const Ctx &x, const Geo &g, int64_t c, int ndays, int k)
{
  const int n = 2 * ndays + 2;
  if (k == 0) k = 2;
  else if (k == n - 1) k = n - 3;
  const int64_t lc = c - x.c0;
  const int i = (k - 1) >> 1, second = (k - 1) & 1;
  const int tn = x.tminh[(int64_t)i * x.CB + lc], tx = x.tmaxh[(int64_t)i * x.CB + lc];
  const int first_is_min = tn < tx;
  return (first_is_min != second) ? day_tmin(x, g, c, i) : day_tmax(x, g, c, i);
}
// hermint at xbar = hour, with hermite's coefficients of the one interval it lands in (yc2 == 0 everywhere)
VA_HD double hourly_tair(const Ctx &x, const Geo &g, int64_t c, int hour)
{
  const int64_t lc = c - x.c0;
  const int ndays = g.ndl, n = 2 * ndays + 2;
  const double xbar = hour;
  // hermint's bisection (calc_air_temperature.c:51-56) ends at the last node k <= n-2 with x[k] <= xbar.  The nodes
  // increase strictly (two per day, the later one before hour 23, the earlier one not before hour -1 of its day), so
  // that node is the last node of the day before unless one of this day's two nodes is already reached.
  const int d = hour / 24;
  int klo = 2 * d;
  if (node_x(x, lc, ndays, 2 * d + 1) <= xbar) klo = 2 * d + 1;
  if (node_x(x, lc, ndays, 2 * d + 2) <= xbar) klo = 2 * d + 2;
  if (klo > n - 2) klo = n - 2;
  const double x0 = node_x(x, lc, ndays, klo), x1 = node_x(x, lc, ndays, klo + 1);
  const double y0 = node_y(x, g, c, ndays, klo), y1 = node_y(x, g, c, ndays, klo + 1);
  double dx = x1 - x0;
  double divdf1 = (y1 - y0) / dx;
  double divdf3 = 0. + 0. - 2 * divdf1;
  double yc3 = (divdf1 - 0. - divdf3) / dx;
  double yc4 = divdf3 / (dx * dx);
  dx = xbar - x0;
  return y0 + dx * (0. + dx * (yc3 + dx * yc4));
}

// The same value with the interval's coefficients kept while consecutive hours stay in it (a record's hours change
// interval at most three times): identical arithmetic, evaluated once per interval instead of once per hour.
struct SplineSeg { int klo = -1; double x0 = 0., y0 = 0., yc3 = 0., yc4 = 0.; };
VA_HD double hourly_tair_cached(const Ctx &x, const Geo &g, int64_t c, int hour, SplineSeg &sg)
{
  const int64_t lc = c - x.c0;
  const int ndays = g.ndl, n = 2 * ndays + 2;
  const double xbar = hour;
  const int d = hour / 24;
  int klo = 2 * d;
  if (node_x(x, lc, ndays, 2 * d + 1) <= xbar) klo = 2 * d + 1;
  if (node_x(x, lc, ndays, 2 * d + 2) <= xbar) klo = 2 * d + 2;
  if (klo > n - 2) klo = n - 2;
  if (klo != sg.klo) {
    const double x0 = node_x(x, lc, ndays, klo), x1 = node_x(x, lc, ndays, klo + 1);
    const double y0 = node_y(x, g, c, ndays, klo), y1 = node_y(x, g, c, ndays, klo + 1);
    const double dx = x1 - x0;
    const double divdf1 = (y1 - y0) / dx;
    const double divdf3 = 0. + 0. - 2 * divdf1;
    sg.klo = klo; sg.x0 = x0; sg.y0 = y0;
    sg.yc3 = (divdf1 - 0. - divdf3) / dx;
    sg.yc4 = divdf3 / (dx * dx);
  }
  const double dx = xbar - sg.x0;
  return sg.y0 + dx * (0. + dx * (sg.yc3 + dx * sg.yc4));
}

// local VP of one hour, initialize_atmos.c:1229-1275
VA_HD double hourly_vp(const Ctx &x, const Geo &g, int64_t lc, int idx)
{
  if (x.r_vp && x.fdt < 24) return obs_vp_hour(x, g, x.c0 + lc, idx);    // sub-daily observations, :891-913
  const int day = idx / 24, hour = idx - day * 24, nd = g.ndl;
  const int64_t CB = x.CB;
  const double v = x.pva[(int64_t)day * CB + lc];
  if (!x.vp_interp) return v;
  const int tm = x.tminh[(int64_t)day * CB + lc];
  if (hour < tm) {
    if (day > 0) {
      const int tmm = x.tminh[(int64_t)(day - 1) * CB + lc];
      const double delta_t_minus = tm + 24 - tmm;         // day == 0 never gets here
      const double vm = x.pva[(int64_t)(day - 1) * CB + lc];
      return vm + (v - vm) * (hour + 24 - tmm) / delta_t_minus;
    }
    return v;
  }
  if (day < nd - 1) {
    const int tmp = x.tminh[(int64_t)(day + 1) * CB + lc];
    const double delta_t_plus = tmp + 24 - tm;            // day == nd-1 never gets here
    const double vp = x.pva[(int64_t)(day + 1) * CB + lc];
    return v + (vp - v) * (hour - tm) / delta_t_plus;
  }
  return v;
}

// Columns the reference derives from others before anything else, element by element of the forcing file's records
// (so in file time as well as in local time): :424-460 and :773-866.
VA_HD double derive_prec(const Ctx &x, double rainf, double snowf) { return u_prec_raw(x, rainf) + u_prec_raw(x, snowf); }
VA_HD double derive_wind(double e, double n) { return sqrt(e * e + n * n); }
VA_HD double derive_vp_qair(const Ctx &x, double q, double p_file) { return q * u_kpa(x, p_file) / EPS_MW; }
VA_HD double derive_vp_rh(const Ctx &x, double rh, double t_file) { return rh * svp(u_temp(x, t_file)) / 100; }

// What initialize_atmos stores for one window of `len` hours starting at local hour `hour` -- a sub-step of SNOW_STEP hours,
// which is the whole record when SNOW_STEP == TIME_STEP -- before the quantities derived from them: precipitation, wind,
// air temperature, shortwave, vapour pressure, and the observed density / pressure / longwave columns when the file has
// them.  nsub = windows per day (NF * stepspday), the divisor of a daily precipitation total (:601-606).
struct Win { double prec, wind, ta, sw, vp, dens, p, lw; int dayi; };
VA_HD Win window_vals(const Ctx &x, const Geo &g, int64_t c, int hour, int len, int nsub, bool with_vp = true)
{
  const int64_t lc = c - x.c0;
  Win w;
  w.dayi = (int)((float)hour / 24.0);
  const int dayi = w.dayi;
  if (x.fdt == 24) {                                       // :597-613, :641-659
    w.prec = u_prec(x, local_daily(x.r_prec, x, g, c, dayi)) / (float)nsub;
    w.wind = x.has_wind ? local_daily(x.r_wind, x, g, c, dayi) : 3.0;
    double s = 0;                                          // :1005-1019
    SplineSeg sg;
    for (int idx = hour; idx < hour + len; idx++) s += hourly_tair_cached(x, g, c, idx, sg);
    w.ta = s / len;
  }
  else {                                                   // :616-628, :662-678, :737-752
    double s = 0;
    for (int idx = hour; idx < hour + len; idx++) s += u_prec(x, local_hourly(x.r_prec, x, g, c, idx)) / x.fdt;
    w.prec = s;
    if (x.has_wind) {
      s = 0;
      for (int idx = hour; idx < hour + len; idx++) {
        double v = local_hourly(x.r_wind, x, g, c, idx);
        s += (v < x.min_wind) ? x.min_wind : v;
      }
      w.wind = s / len;
    }
    else w.wind = 3.0;
    s = 0;
    for (int idx = hour; idx < hour + len; idx++) s += u_temp(x, local_hourly(x.r_ta, x, g, c, idx));
    w.ta = s / len;
  }
  double sw = 0, vp = 0;                                   // :974-987, :1278-1291
  int rad_day = -1, rad_ydi = 0;
  double tmp_rad = 0.;
  for (int idx = hour; idx < hour + len; idx++) {
    const int d = idx / 24;
    if (x.r_sw) sw += hourly_rad(x, g, lc, d, idx - d * 24);
    else {
      if (d != rad_day) {                                  // hourly_rad()'s per-day factor, once per day
        rad_day = d;
        rad_ydi = ydi_of(x, g, d);
        tmp_rad = x.srad[(int64_t)d * x.CB + lc] * x.dayl[(int64_t)rad_ydi * x.CB + lc] / 3600.;
      }
      sw += x.hf[((int64_t)rad_ydi * 24 + (idx - d * 24)) * x.CB + lc] * tmp_rad;
    }
    if (with_vp) vp += hourly_vp(x, g, lc, idx);
  }
  w.sw = sw / len;
  w.vp = vp / len;
  w.dens = 0.; w.p = 0.; w.lw = 0.;                        // observed columns, :1032-1062, :1117-1145, :1389-1420
  if (x.r_de) {
    if (x.fdt == 24) w.dens = local_daily(x.r_de, x, g, c, dayi);
    else { double s = 0; for (int idx = hour; idx < hour + len; idx++) s += local_hourly(x.r_de, x, g, c, idx); w.dens = s / len; }
  }
  if (x.r_pr) {
    if (x.fdt == 24) w.p = u_kpa(x, local_daily(x.r_pr, x, g, c, dayi));
    else { double s = 0; for (int idx = hour; idx < hour + len; idx++) s += u_kpa(x, local_hourly(x.r_pr, x, g, c, idx)); w.p = s / len; }
  }
  if (x.r_lw) {
    if (x.fdt == 24) w.lw = local_daily(x.r_lw, x, g, c, dayi);
    else { double s = 0; for (int idx = hour; idx < hour + len; idx++) s += local_hourly(x.r_lw, x, g, c, idx); w.lw = s / len; }
  }
  return w;
}

// pressure and density of a window or of the record from what is observed and its air temperature, :1070-1170
VA_HD void pressure_density(const Ctx &x, double elevation, double ta, double &p, double &dens)
{
  if (x.r_pr) { }
  else if (x.r_de) p = x.plapse ? (KELVIN + ta) * dens * RD : (275.0 + ta) * dens / 0.003486;
  else if (x.plapse) p = PS_PM * exp(-elevation * GRAV / (RD * (KELVIN + ta + 0.5 * elevation * T_LAPSE)));
  else p = 95500.;
  if (!x.r_de) dens = x.plapse ? p / (RD * (KELVIN + ta)) : 0.003486 * p / (275.0 + ta);
}

VA_HD double cloud_tskc(const Ctx &x, int64_t lc, int dayi)     // :1188-1193, :1361-1388
{
  const double tf = x.tskc[(int64_t)dayi * x.CB + lc];
  return (x.lw_cloud == VR_LW_CLOUD_DEARDORFF) ? (1. - tf) : sqrt((1. - tf) / 0.65);
}

// Daily QAIR or REL_HUMID without the pressure / air temperature to turn them into VP beforehand (:1176-1215): after MTCLIM
// the reference walks the records and their windows in order and sets daily_vp[day of the window] = QAIR[day] *
// pressure[window] / EPS (or REL_HUMID[day] * svp(air_temp[window]) / 100); the last window that starts in a day leaves
// its value, days no record reaches keep MTCLIM's.  One (cell, local day).
VA_HD void humid_day(const Ctx &x, int64_t c, int day)
{
  if (!x.r_q && !x.r_rh) return;
  const Geo g = cell_geo(x, c);
  if (day >= g.ndl) return;
  const int NF = x.NF, ss = x.dt / NF, stepspday = 24 / x.dt;
  const int base = x.starthour - g.hoi + g.shift;
  const long last = (long)day * 24 + 23 - base;
  if (last < 0) return;
  long m = last / ss;
  if (m >= (long)x.nrecs * NF) m = (long)x.nrecs * NF - 1;
  const int hour = base + (int)m * ss;
  if ((int)((float)hour / 24.0) != day) return;
  const Win w = window_vals(x, g, c, hour, ss, NF * stepspday, false);
  double v;
  if (x.r_q) {
    double p = w.p, dens = w.dens;
    pressure_density(x, x.elev[c], w.ta, p, dens);
    v = local_daily(x.r_q, x, g, c, day) * p / EPS_MW;
  }
  else v = local_daily(x.r_rh, x, g, c, day) * svp(w.ta) / 100;
  x.pva[(int64_t)day * x.CB + (c - x.c0)] = v;
}

VA_HD void store_fields(const Ctx &x, int64_t o, double ta, double prec, double wind, double vp, double vpd, double p, double dens,
                        double sw, double lw)
{
  x.out[VR_F_AIR_TEMP][o] = ta;
  x.out[VR_F_PREC][o] = prec;
  x.out[VR_F_WIND][o] = wind;
  x.out[VR_F_VP][o] = vp;
  x.out[VR_F_VPD][o] = vpd;
  x.out[VR_F_PRESSURE][o] = p;
  x.out[VR_F_DENSITY][o] = dens;
  x.out[VR_F_SHORTWAVE][o] = sw;
  x.out[VR_F_LONGWAVE][o] = lw;
}

// One record of one cell.  NF == 1 (SNOW_STEP == TIME_STEP): the record is its only window, stored at [rec][c].  NF > 1: the
// NF windows of SNOW_STEP hours at [rec][0 .. NF-1][c], then the record's own values at [rec][NF][c] -- sums or means of
// the windows' values in the reference's order, NOT means over the record's hours -- as atmos-><field>[0..NF-1], [NR].
VA_HD void record_inl(const Ctx &x, int64_t c, int rec)
{
  const int64_t lc = c - x.c0;
  const Geo g = cell_geo(x, c);
  const int dt = x.dt, stepspday = 24 / dt, NF = x.NF;
  const int hour = rec * dt + x.starthour - g.hoi + g.shift;
  const double elevation = x.elev[c];
  const double min_tf = x.min_tf ? x.min_tf[c] : 0.;
  if (NF == 1) {
    const int64_t o = (int64_t)rec * x.C + c;
    Win w = window_vals(x, g, c, hour, dt, stepspday);
    // the MIN_WIND_SPEED clamp at :654-657 addresses wind[NF], one past the record's value: no effect
    double p = w.p, dens = w.dens, vp = w.vp;
    pressure_density(x, elevation, w.ta, p, dens);
    double vpd = svp(w.ta) - vp;                           // :1336-1355
    if (vpd < 0) { vpd = 0; vp = svp(w.ta); }
    const double lw = x.r_lw ? w.lw : longwave(x, cloud_tskc(x, lc, w.dayi), w.ta, vp);
    store_fields(x, o, w.ta, w.prec, w.wind, vp, vpd, p, dens, w.sw, lw);
    if (x.snowflag) x.snowflag[o] = ((w.ta + min_tf) < x.max_snow_temp && w.prec > 0) ? 1 : 0;      // :1736-1753
    return;
  }
  const int ss = dt / NF;
  const int64_t o0 = ((int64_t)rec * (NF + 1)) * x.C + c;
  double s_prec = 0, s_wind = 0, s_ta = 0, s_sw = 0, s_vp0 = 0, s_dens = 0, s_p = 0, s_vpd = 0, s_vp = 0, s_lw = 0;
  int flag = 0;
  for (int i = 0; i < NF; i++) {
    Win w = window_vals(x, g, c, hour + i * ss, ss, NF * stepspday);
    double p = w.p, dens = w.dens, vp = w.vp;
    pressure_density(x, elevation, w.ta, p, dens);
    double vpd = svp(w.ta) - vp;
    if (vpd < 0) { vpd = 0; vp = svp(w.ta); }
    const double lw = x.r_lw ? w.lw : longwave(x, cloud_tskc(x, lc, w.dayi), w.ta, vp);
    store_fields(x, o0 + (int64_t)i * x.C, w.ta, w.prec, w.wind, vp, vpd, p, dens, w.sw, lw);
    s_prec += w.prec; s_wind += w.wind; s_ta += w.ta; s_sw += w.sw; s_vp0 += w.vp; s_dens += w.dens; s_p += w.p;
    s_vpd += vpd; s_vp += vp; s_lw += lw;
    if ((w.ta + min_tf) < x.max_snow_temp && w.prec > 0) flag = 1;
  }
  const float fNF = (float)NF;
  const double ta = s_ta / fNF;
  double wind = x.has_wind ? s_wind / fNF : 3.0;
  // :654-657 clamps wind[j] with j == NF after the loop: with NF > 1 that IS the record's value (NR == NF)
  if (x.has_wind && x.fdt == 24 && dt == 24 && wind < x.min_wind) wind = x.min_wind;
  double p = s_p / fNF, dens = s_dens / fNF;               // observed columns: means of the windows (:1126, :1044)
  pressure_density(x, elevation, ta, p, dens);
  double vp, vpd;
  if (x.r_vp || x.vp_interp) { vpd = s_vpd / fNF; vp = s_vp / fNF; }          // :1345-1349
  else { vp = s_vp0 / fNF; vpd = svp(ta) - vp; }                              // :1350-1352 (no clamp)
  store_fields(x, o0 + (int64_t)NF * x.C, ta, s_prec, wind, vp, vpd, p, dens, s_sw / fNF, s_lw / fNF);
  if (x.snowflag) x.snowflag[(int64_t)rec * x.C + c] = flag;
}

VA_HDN void record(const Ctx &x, int64_t c, int rec) { record_inl(x, c, rec); }

// ---------------------------------------------------------------------------------------------------------------
// host helpers shared by the library and the test harness
inline int ndays_of(const vr_atmos_options &o)            // initialize_atmos.c:262-265
{
  int last_hour = (o.starthour + (o.nrecs - 1) * o.time_step) % 24;
  int tmp_nrecs = o.nrecs + o.starthour - 0 + (24 - o.time_step) - last_hour;
  return (tmp_nrecs * o.time_step) / 24;
}
// day_in_year of every local day, starting one day before the start date (initialize_atmos.c:274-338)
inline void build_calendar(const vr_atmos_options &o, int ndays, int16_t *cal /* [ndays+2] */)
{
  static const int month_days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int day = o.startday - 1, month = o.startmonth, year = o.startyear;
  if (day < 1) {
    month--;
    if (month < 1) { month = 12; year--; }
    day = month_days[month - 1];
    if (year % 4 == 0 && month == 2) day++;
  }
  int diy = day;
  for (int m = 1; m < month; m++) {
    int dim = month_days[m - 1];
    if (year % 4 == 0 && m == 2) dim++;
    diy += dim;
  }
  for (int i = 0; i < ndays + 2; i++) {
    cal[i] = (int16_t)diy;
    diy++; day++;
    int dim = month_days[month - 1];
    if (year % 4 == 0 && month == 2) dim++;
    if (day > dim) {
      day = 1; month++;
      if (month > 12) { diy = 1; month = 1; year++; }
    }
  }
}
inline const char *unsupported_raw(const vr_atmos_options &o, const vr_atmos_raw &r)
{
  const bool vp_before = r.vp || (r.qair && r.pressure) || (!r.qair && r.rel_humid && o.layout == VR_ATMOS_SUBDAILY_AIR_TEMP);
  if (o.vp_iter == VR_VP_ITER_CONVERGE && (r.shortwave || vp_before)) return "VP_ITER_CONVERGE with observed SHORTWAVE or VP columns";
  if (!r.vp && r.qair && !r.pressure && o.force_dt < 24)
    return "sub-daily QAIR without PRESSURE (the reference reads atmos->pressure at an index left over from an earlier loop, initialize_atmos.c:1302)";
  if ((r.rainf != nullptr) != (r.snowf != nullptr)) return "RAINF and SNOWF are used together";
  if ((r.wind_e != nullptr) != (r.wind_n != nullptr)) return "WIND_E and WIND_N are used together";
  return nullptr;
}
inline const char *unsupported(const vr_atmos_options &o)
{
  if (o.time_step < 1 || o.time_step > 24 || 24 % o.time_step) return "TIME_STEP must divide 24";
  if (o.nrecs < 1) return "nrecs < 1";
  if (o.snow_step != 0) {                                 // get_global_param.c:761-765
    if (o.time_step < 24 && o.snow_step != o.time_step) return "If the model step is smaller than daily, the snow model should run at the same time step as the rest of the model.";
    if (o.snow_step < 1 || o.snow_step > o.time_step || o.time_step % o.snow_step) return "SNOW_STEP should be <= TIME_STEP and divide TIME_STEP evenly";
  }
  if (o.layout == VR_ATMOS_DAILY_TMAX_TMIN) { if (o.force_dt != 24) return "daily TMAX/TMIN layout needs FORCE_DT 24"; }
  else if (o.layout == VR_ATMOS_SUBDAILY_AIR_TEMP) {
    if (o.force_dt < 1 || o.force_dt >= 24 || 24 % o.force_dt) return "sub-daily AIR_TEMP layout needs FORCE_DT below 24 dividing 24";
  }
  else return "unknown forcing layout";
  if (o.vp_iter < 0 || o.vp_iter > 3 || o.lw_type < 0 || o.lw_type > 5 || o.lw_cloud < 0 || o.lw_cloud > 1) return "option value out of range";
  return nullptr;
}
inline void fill_options(const vr_atmos_options &o, Ctx &x)
{
  x.dt = o.time_step; x.nrecs = o.nrecs; x.starthour = o.starthour; x.layout = o.layout; x.fdt = o.force_dt;
  x.has_wind = o.has_wind; x.vp_interp = o.vp_interp; x.vp_iter = o.vp_iter; x.swe_corr = o.mtclim_swe_corr;
  x.lw_type = o.lw_type; x.lw_cloud = o.lw_cloud; x.plapse = o.plapse; x.alma = o.alma_input;
  x.min_wind = (double)(float)o.min_wind_speed;       // option_struct.MIN_WIND_SPEED is a float (vicNl_def.h)
  x.sw_thresh = o.sw_prec_thresh; x.Ndays = ndays_of(o); x.fsteps = 24 / o.force_dt;
  x.NF = o.snow_step > 0 ? o.time_step / o.snow_step : 1; x.max_snow_temp = o.max_snow_temp;
}

}  // namespace va
#endif
