/* atmos_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle side).  Never linked into the product.
 *
 * Plain-C, one-cell-at-a-time restatement of the reference's forcing preparation for the two forcing
 * layouts of include/vicres_atmos.h, with or without the observed SHORTWAVE / LONGWAVE / VP / PRESSURE / DENSITY columns, QAIR /
 * REL_HUMID, the split RAINF + SNOWF and WIND_E + WIND_N columns, and the NF = TIME_STEP / SNOW_STEP windows per record of a
 * sub-stepped snow model with its snow flags:
 *   initialize_atmos            runoff/model/initialize_atmos.c:137-1818
 *   mtclim_wrapper/init/to_vic  runoff/model/mtclim_wrapper.c:69-278
 *   calc_tair, calc_prcp, snowpack, calc_srad_humidity_iterative, compute_srad_humidity_onetime,
 *   calc_pet, atm_pres, pulled_boxcar          runoff/model/mtclim_vic.c:404-1395
 *   HourlyT, hermite, hermint, set_max_min_hour runoff/model/calc_air_temperature.c:15-201
 *   calc_longwave               runoff/model/calc_longwave.c:7-76
 *   svp                         runoff/model/svp.c:7-22
 * Every function names the lines it follows.  Parity pinned: bit-identical to the compiled reference's
 * atmos[] records (oracle/_ref/vic_ref_driver dumps them in fp64) on the fixtures of
 * tools/make_atmos_golden.py (tests/test_atmos_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/vicres_atmos.h"

/* vicNl_def.h:275-333 and mtclim_constants_vic.h / mtclim_parameters_vic.h */
#define KELVIN 273.15
#define STEFAN_B 5.6696e-8
#define EPS_MW 0.62196351
#define GRAV 9.81
#define RD 287
#define PS_PM 101300
#define T_LAPSE -0.0065
#define A_SVP 0.61078
#define B_SVP 17.269
#define C_SVP 237.3
#define SECPERRAD 13750.9871
#define RADPERDAY 0.017214
#define RADPERDEG 0.01745329
#define MINDECL -0.4092797
#define DAYSOFF 11.25
#define SRADDT 30.0
#define MA 28.9644e-3
#define R_GAS 8.3143
#define G_STD 9.80665
#define P_STD 101325.0
#define T_STD 288.15
#define CP_AIR 1010.0
#define LR_STD 0.0065
#define MT_PI 3.1415927          /* vicNl_def.h:309 is seen first, so mtclim's #ifndef PI keeps this value */
#define TDAYCOEF 0.45
#define SNOW_TCRIT -6.0
#define SNOW_TRATE 0.042
#define TBASE 0.870
#define ABASE -6.1e-5
#define C_RAD 1.5
#define B0 0.031
#define B1 0.201
#define B2 0.185
#define RAIN_SCALAR 0.75
#define DIF_ALB 0.6
#define TINYSTEPS 2880           /* 86400 / SRADDT */

/* svp.c:7-22 */
static double svp(double temp)
{
  double v = A_SVP * exp((B_SVP * temp) / (C_SVP + temp));
  if (temp < 0) v *= 1.0 + .00972 * temp + .000042 * temp * temp;
  return v * 1000.;
}

/* mtclim_vic.c:1255-1313 */
static double calc_pet(double rad, double ta, double pa, double dayl)
{
  double rnet = rad * 0.72;
  double lhvap = 2.5023e6 - 2430.54 * ta;
  double gamma = CP_AIR * pa / (lhvap * EPS_MW);
  double dt = 0.2, t1 = ta + dt, t2 = ta - dt;
  double pvs1 = svp(t1), pvs2 = svp(t2);
  double s = (pvs1 - pvs2) / (t1 - t2);
  double pet = (1.26 * (s / (s + gamma)) * rnet * dayl) / lhvap;
  return pet / 10.0;
}

/* mtclim_vic.c:1316-1334 */
static double atm_pres(double elev)
{
  double t1 = 1.0 - (LR_STD * elev) / T_STD;
  double t2 = G_STD / (LR_STD * (R_GAS / MA));
  return P_STD * pow(t1, t2);
}

/* mtclim_vic.c:1338-1394, flat weights (w_flag == 0) */
static void pulled_boxcar(const double *in, double *out, int n, int w)
{
  int i, j;
  double sum_wt = 0.0;
  for (i = 0; i < w; i++) sum_wt += 1.0;
  for (i = w - 1; i < n; i++) {
    double total = 0.0;
    for (j = 0; j < w; j++) total += in[i - w + j + 1] * 1.0;
    out[i] = total / sum_wt;
  }
  for (i = 0; i < w - 1; i++) out[i] = out[w - 1];
}

/* calc_longwave.c:7-76 */
static double calc_longwave(const vr_atmos_options *o, double tskc, double air_temp, double vp)
{
  double emissivity, emissivity_clear = 0., x;
  air_temp += KELVIN;
  vp /= 100;
  if (o->lw_type == VR_LW_TVA) emissivity_clear = 0.740 + 0.0049 * vp;
  else if (o->lw_type == VR_LW_ANDERSON) emissivity_clear = 0.68 + 0.036 * pow(vp, 0.5);
  else if (o->lw_type == VR_LW_BRUTSAERT) { x = vp / air_temp; emissivity_clear = 1.24 * pow(x, 0.14285714); }
  else if (o->lw_type == VR_LW_SATTERLUND) emissivity_clear = 1.08 * (1 - exp(-1 * pow(vp, (air_temp / 2016))));
  else if (o->lw_type == VR_LW_IDSO) emissivity_clear = 0.7 + 5.95e-5 * vp * exp(1500 / air_temp);
  else if (o->lw_type == VR_LW_PRATA) { x = 46.5 * vp / air_temp; emissivity_clear = 1 - (1 + x) * exp(-1 * pow((1.2 + 3 * x), 0.5)); }
  if (o->lw_cloud == VR_LW_CLOUD_DEARDORFF) {
    double cloudfrac = tskc;
    emissivity = cloudfrac * 1.0 + (1 - cloudfrac) * emissivity_clear;
  }
  else {
    double cloudfactor = 1.0 + 0.17 * tskc * tskc;
    emissivity = cloudfactor * emissivity_clear;
  }
  return emissivity * STEFAN_B * air_temp * air_temp * air_temp * air_temp;
}

/* calc_air_temperature.c:139-201 */
static void set_max_min_hour(const double *hourlyrad, int ndays, int *tmaxhour, int *tminhour)
{
  int i, hour;
  for (i = 0; i < ndays; i++) {
    int risehour = -999, sethour = -999;
    for (hour = 0; hour < 12; hour++)
      if (hourlyrad[i * 24 + hour] > 0 && (i * 24 + hour == 0 || hourlyrad[i * 24 + hour - 1] <= 0)) risehour = hour;
    for (hour = 12; hour < 24; hour++)
      if (hourlyrad[i * 24 + hour] <= 0 && hourlyrad[i * 24 + hour - 1] > 0) sethour = hour;
    if (i == ndays - 1 && sethour == -999) sethour = 23;
    if (risehour >= 0 && sethour >= 0) {
      tmaxhour[i] = 0.67 * (sethour - risehour) + risehour;
      tminhour[i] = risehour - 1;
    }
    else { tminhour[i] = 2; tmaxhour[i] = 14; }
  }
}

/* calc_air_temperature.c:15-137 (hermite, hermint, HourlyT with Dt == 1) */
static void hourly_t(int ndays, const int *tmaxhour, const double *tmax, const int *tminhour, const double *tmin, double *tair)
{
  int n = ndays * 2 + 2, i, j, hour;
  double *x = calloc(n, sizeof(double)), *y = calloc(n, sizeof(double));
  double *yc2 = calloc(n, sizeof(double)), *yc3 = calloc(n, sizeof(double)), *yc4 = calloc(n, sizeof(double));
  for (i = 0, j = 1, hour = 0; i < ndays; i++, hour += 24) {
    if (tminhour[i] < tmaxhour[i]) {
      x[j] = tminhour[i] + hour; y[j++] = tmin[i];
      x[j] = tmaxhour[i] + hour; y[j++] = tmax[i];
    }
    else {
      x[j] = tmaxhour[i] + hour; y[j++] = tmax[i];
      x[j] = tminhour[i] + hour; y[j++] = tmin[i];
    }
  }
  x[0] = x[2] - 24; y[0] = y[2];
  x[n - 1] = x[n - 3] + 24; y[n - 1] = y[n - 3];
  for (i = 0; i < n - 1; i++) {
    double dx = x[i + 1] - x[i];
    double divdf1 = (y[i + 1] - y[i]) / dx;
    double divdf3 = yc2[i] + yc2[i + 1] - 2 * divdf1;
    yc3[i] = (divdf1 - yc2[i] - divdf3) / dx;
    yc4[i] = divdf3 / (dx * dx);
  }
  for (i = 0, hour = 0; i < ndays * 24; i++, hour += 1) {
    int klo = 0, khi = n - 1, k;
    double xbar = hour, dx;
    while (khi - klo > 1) { k = (khi + klo) >> 1; if (x[k] > xbar) khi = k; else klo = k; }
    dx = xbar - x[klo];
    tair[i] = y[klo] + dx * (yc2[klo] + dx * (yc3[klo] + dx * yc4[klo]));
  }
  free(x); free(y); free(yc2); free(yc3); free(yc4);
}

typedef struct mt_data {
  int ndays, insw, invp;
  int *yday;
  double *tmax, *tmin, *prcp, *s_tmax, *s_tmin, *s_tday, *s_prcp, *s_hum, *s_srad, *s_dayl, *s_swe, *s_fdir, *s_tskc,
         *s_tfmax;
} mt_data;

/* mtclim_vic.c:1098-1223 (no observed shortwave or vapour pressure: ctrl->insw == ctrl->invp == 0) */
static void srad_humidity_onetime(const vr_atmos_options *o, mt_data *d, double *tdew, double *pva, const double *ttmax0,
                                  const double *flat_potrad, const double *slope_potrad, double sky_prop,
                                  const double *daylength, double *pet, const double *parray, double pa, const double *dtr)
{
  int i, ndays = d->ndays;
  for (i = 0; i < ndays; i++) {
    int yday = d->yday[i] - 1;
    double t_tmax = ttmax0[yday] + ABASE * pva[i], t_final, pdif, pdir, srad1, srad2, sc;
    if (t_tmax < 0.0001) t_tmax = 0.0001;
    t_final = t_tmax * d->s_tfmax[i];
    pdif = -1.25 * t_final + 1.25;
    if (pdif > 1.0) pdif = 1.0;
    if (pdif < 0.0) pdif = 0.0;
    pdir = 1.0 - pdif;
    srad1 = slope_potrad[yday] * t_final * pdir;
    srad2 = flat_potrad[yday] * t_final * pdif * (sky_prop + DIF_ALB * (1.0 - sky_prop));
    if (o->mtclim_swe_corr && d->s_swe[i] > 0.0) {
      sc = (1.32 + 0.096 * d->s_swe[i]) * 1e6;
      if (daylength[yday] > 0.0) sc /= daylength[yday];
      else sc = 0.0;
      if (sc > 100.0) sc = 100.0;
    }
    else sc = 0.0;
    if (d->insw) {                                     /* :1173-1182 */
      double potrad = (srad1 + srad2 + sc) * daylength[yday] / t_final / 86400;
      if (potrad > 0 && d->s_srad[i] > 0 && daylength[yday] > 0) {
        d->s_tfmax[i] = (d->s_srad[i]) / (potrad * t_tmax);
        if (d->s_tfmax[i] > 1.0) d->s_tfmax[i] = 1.0;
      }
      else d->s_tfmax[i] = 1.0;
    }
    else d->s_srad[i] = srad1 + srad2 + sc;
    if (o->lw_cloud == VR_LW_CLOUD_DEARDORFF) d->s_tskc[i] = (1. - d->s_tfmax[i]);
    else d->s_tskc[i] = sqrt((1. - d->s_tfmax[i]) / 0.65);
    d->s_fdir[i] = pdir;
  }
  for (i = 0; i < ndays; i++) {
    double tmink = d->s_tmin[i] + KELVIN, ratio, ratio2, ratio3, tdewk;
    pet[i] = calc_pet(d->s_srad[i], d->s_tday[i], pa, d->s_dayl[i]);
    ratio = pet[i] / parray[i];
    ratio2 = ratio * ratio;
    ratio3 = ratio2 * ratio;
    tdewk = tmink * (-0.127 + 1.121 * (1.003 - 1.444 * ratio + 12.312 * ratio2 - 32.766 * ratio3) + 0.0006 * (dtr[i]));
    tdew[i] = tdewk - KELVIN;
    pva[i] = svp(tdew[i]);
  }
}

/* mtclim_vic.c:535-1095; radfract = the reference's tiny_radfract[366][2880], flat */
static void srad_humidity_iterative(const vr_atmos_options *o, mt_data *d, double site_lat, double site_elev, double site_slp,
                                    double site_asp, double site_ehoriz, double site_whoriz, double *radfract)
{
  static const double optam[21] = {2.90, 3.05, 3.21, 3.39, 3.69, 3.82, 4.07, 4.37, 4.72, 5.12, 5.60,
                                   6.18, 6.88, 7.77, 8.90, 10.39, 12.44, 15.36, 19.79, 26.96, 30.00};
  int ndays = d->ndays, i, j, iter, max_iter;
  double ttmax0[366], flat_potrad[366], slope_potrad[366], daylength[366];
  double *dtr = malloc(ndays * sizeof(double)), *sm_dtr = malloc(ndays * sizeof(double));
  double *parray = malloc(ndays * sizeof(double)), *window = malloc((ndays + 90) * sizeof(double));
  double *tdew = malloc(ndays * sizeof(double)), *pet = malloc(ndays * sizeof(double));
  double *pva = malloc(ndays * sizeof(double)), *tdew_save = malloc(ndays * sizeof(double));
  double *pva_save = malloc(ndays * sizeof(double));
  double sum_prcp, ann_prcp, sum_pet, ann_pet, t1, t2, pratio, trans1, lat, coslat, sinlat, cosslp, sinslp, cosasp, sinasp;
  double coszeh, coszwh, dt, dh, pa, sky_prop, avg_horizon, slope_excess, horizon_scalar, slope_scalar, tol, rmse_tdew;

  for (i = 0; i < ndays; i++) {                     /* :649-655 */
    double tmax = d->tmax[i], tmin = d->tmin[i];
    if (tmax < tmin) tmax = tmin;
    dtr[i] = tmax - tmin;
  }
  pulled_boxcar(dtr, sm_dtr, ndays, ndays >= 30 ? 30 : ndays);   /* :658-669 */
  sum_prcp = 0.0;                                   /* :672-677 */
  for (i = 0; i < ndays; i++) sum_prcp += d->s_prcp[i];
  ann_prcp = (sum_prcp / (double)ndays) * 365.25;
  if (ann_prcp == 0.0) ann_prcp = 1.0;
  if (ndays < 90) {                                 /* :684-699 */
    double eff;
    sum_prcp = 0.0;
    for (i = 0; i < ndays; i++) sum_prcp += d->s_prcp[i];
    eff = (sum_prcp / (double)ndays) * 365.25;
    if (eff < 8.0) eff = 8.0;
    for (i = 0; i < ndays; i++) parray[i] = eff;
  }
  else {                                            /* :700-741 */
    int start_yday = d->yday[0], end_yday = d->yday[ndays - 1], isloop;
    if (start_yday != 1) isloop = (end_yday == start_yday - 1) ? 1 : 0;
    else isloop = (end_yday == 365 || end_yday == 366) ? 1 : 0;
    for (i = 0; i < 90; i++) window[i] = isloop ? d->s_prcp[ndays - 90 + i] : d->s_prcp[i];
    for (i = 0; i < ndays; i++) window[i + 90] = d->s_prcp[i];
    for (i = 0; i < ndays; i++) {
      sum_prcp = 0.0;
      for (j = 0; j < 90; j++) sum_prcp += window[i + j];
      sum_prcp = (sum_prcp / 90.0) * 365.25;
      parray[i] = (sum_prcp < 8.0) ? 8.0 : sum_prcp;
    }
  }

  t1 = 1.0 - (LR_STD * site_elev) / T_STD;          /* :754-759 */
  t2 = G_STD / (LR_STD * (R_GAS / MA));
  pratio = pow(t1, t2);
  trans1 = pow(TBASE, pratio);
  lat = site_lat;                                   /* :764-786 */
  lat *= RADPERDEG;
  if (lat > 1.5707) lat = 1.5707;
  if (lat < -1.5707) lat = -1.5707;
  coslat = cos(lat); sinlat = sin(lat);
  cosslp = cos(site_slp * RADPERDEG); sinslp = sin(site_slp * RADPERDEG);
  cosasp = cos(site_asp * RADPERDEG); sinasp = sin(site_asp * RADPERDEG);
  coszeh = cos(1.570796 - (site_ehoriz * RADPERDEG));
  coszwh = cos(1.570796 - (site_whoriz * RADPERDEG));
  dt = SRADDT;
  dh = dt / SECPERRAD;
  memset(radfract, 0, sizeof(double) * 366 * TINYSTEPS);   /* mtclim_wrapper.c:212-217 */

  for (i = 0; i < 365; i++) {                       /* :789-921 */
    double decl = MINDECL * cos(((double)i + DAYSOFF) * RADPERDAY);
    double cosdecl = cos(decl), sindecl = sin(decl);
    double bsg1 = -sinslp * sinasp * cosdecl;
    double bsg2 = (-cosasp * sinslp * sinlat + cosslp * coslat) * cosdecl;
    double bsg3 = (cosasp * sinslp * coslat + cosslp * sinlat) * sindecl;
    double cosegeom = coslat * cosdecl, sinegeom = sinlat * sindecl;
    double coshss = -(sinegeom) / cosegeom, hss, sc, dir_beam_topa, sum_trans = 0.0, sum_flat_potrad = 0.0,
           sum_slope_potrad = 0.0, h;
    double *rf = radfract + (size_t)i * TINYSTEPS;
    if (coshss < -1.0) coshss = -1.0;
    if (coshss > 1.0) coshss = 1.0;
    hss = acos(coshss);
    daylength[i] = 2.0 * hss * SECPERRAD;
    if (daylength[i] > 86400) daylength[i] = 86400;
    sc = 1368.0 + 45.5 * sin((2.0 * MT_PI * (double)i / 365.25) + 1.7);
    dir_beam_topa = sc * dt;
    for (h = -hss; h < hss; h += dh) {
      double cosh_ = cos(h), sinh_ = sin(h);
      double cza = cosegeom * cosh_ + sinegeom;
      double cbsa = sinh_ * bsg1 + cosh_ * bsg2 + bsg3;
      double dir_flat_topa;
      int tinystep;
      if (cza > 0.0) {
        double am, trans2;
        dir_flat_topa = dir_beam_topa * cza;
        am = 1.0 / (cza + 0.0000001);
        if (am > 2.9) {
          int ami = (int)(acos(cza) / RADPERDEG) - 69;
          if (ami < 0) ami = 0;
          if (ami > 20) ami = 20;
          am = optam[ami];
        }
        trans2 = pow(trans1, am);
        sum_trans += trans2 * dir_flat_topa;
        sum_flat_potrad += dir_flat_topa;
        if ((h < 0.0 && cza > coszeh && cbsa > 0.0) || (h >= 0.0 && cza > coszwh && cbsa > 0.0))
          sum_slope_potrad += dir_beam_topa * cbsa;
      }
      else dir_flat_topa = -1;
      tinystep = (12L * 3600L + h * SECPERRAD) / SRADDT;
      if (tinystep < 0) tinystep = 0;
      if (tinystep > TINYSTEPS - 1) tinystep = TINYSTEPS - 1;
      rf[tinystep] = (dir_flat_topa > 0) ? dir_flat_topa : 0;
    }
    if (daylength[i] && sum_flat_potrad > 0)
      for (j = 0; j < TINYSTEPS; j++) rf[j] /= sum_flat_potrad;
    if (daylength[i]) {
      ttmax0[i] = sum_trans / sum_flat_potrad;
      flat_potrad[i] = sum_flat_potrad / daylength[i];
      slope_potrad[i] = sum_slope_potrad / daylength[i];
    }
    else { ttmax0[i] = 0.0; flat_potrad[i] = 0.0; slope_potrad[i] = 0.0; }
  }
  ttmax0[365] = ttmax0[364]; flat_potrad[365] = flat_potrad[364];   /* :924-932 */
  slope_potrad[365] = slope_potrad[364]; daylength[365] = daylength[364];
  memcpy(radfract + (size_t)365 * TINYSTEPS, radfract + (size_t)364 * TINYSTEPS, sizeof(double) * TINYSTEPS);

  avg_horizon = (site_ehoriz + site_whoriz) / 2.0;  /* :938-951 */
  horizon_scalar = 1.0 - sin(avg_horizon * RADPERDEG);
  slope_excess = (site_slp > avg_horizon) ? site_slp - avg_horizon : 0.0;
  if (2.0 * avg_horizon > 180.0) slope_scalar = 0.0;
  else {
    slope_scalar = 1.0 - (slope_excess / (180.0 - 2.0 * avg_horizon));
    if (slope_scalar < 0.0) slope_scalar = 0.0;
  }
  sky_prop = horizon_scalar * slope_scalar;

  for (i = 0; i < ndays; i++) {                     /* :956-967 */
    double b = B0 + B1 * exp(-B2 * sm_dtr[i]);
    double t_fmax = 1.0 - 0.9 * exp(-b * pow(dtr[i], C_RAD));
    if (d->prcp[i] > o->sw_prec_thresh) t_fmax *= RAIN_SCALAR;
    d->s_tfmax[i] = t_fmax;
  }
  for (i = 0; i < ndays; i++) { tdew[i] = d->s_tmin[i]; pva[i] = d->invp ? d->s_hum[i] : svp(tdew[i]); }   /* :976-996 */
  pa = atm_pres(site_elev);                         /* :999-1005 */
  for (i = 0; i < ndays; i++) {
    d->s_dayl[i] = daylength[d->yday[i] - 1];
    tdew_save[i] = tdew[i];
    pva_save[i] = pva[i];
  }
  srad_humidity_onetime(o, d, tdew, pva, ttmax0, flat_potrad, slope_potrad, sky_prop, daylength, pet, parray, pa, dtr);
  sum_pet = 0.0;                                    /* :1011-1015 */
  for (i = 0; i < ndays; i++) sum_pet += pet[i];
  ann_pet = (sum_pet / (double)ndays) * 365.25;
  if (d->invp || (o->vp_iter == VR_VP_ITER_ANNUAL && ann_pet / ann_prcp >= 2.5))   /* :1018-1023 */
    for (i = 0; i < ndays; i++) { tdew[i] = tdew_save[i]; pva[i] = pva_save[i]; }
  if (o->vp_iter == VR_VP_ITER_ALWAYS || (o->vp_iter == VR_VP_ITER_ANNUAL && ann_pet / ann_prcp >= 2.5) ||
      o->vp_iter == VR_VP_ITER_CONVERGE)            /* :1026-1038 */
    max_iter = (o->vp_iter == VR_VP_ITER_CONVERGE) ? 100 : 2;
  else max_iter = 1;
  tol = 0.01; iter = 1; rmse_tdew = tol + 1;        /* :1041-1057 */
  while (rmse_tdew > tol && iter < max_iter) {
    for (i = 0; i < ndays; i++) tdew_save[i] = tdew[i];
    srad_humidity_onetime(o, d, tdew, pva, ttmax0, flat_potrad, slope_potrad, sky_prop, daylength, pet, parray, pa, dtr);
    rmse_tdew = 0;
    for (i = 0; i < ndays; i++) rmse_tdew += (tdew[i] - tdew_save[i]) * (tdew[i] - tdew_save[i]);
    rmse_tdew /= ndays;
    rmse_tdew = pow(rmse_tdew, 0.5);
    iter++;
  }
  if (!d->invp) for (i = 0; i < ndays; i++) d->s_hum[i] = pva[i]; /* :1060-1066 */
  free(dtr); free(sm_dtr); free(parray); free(window); free(tdew); free(pet); free(pva); free(tdew_save); free(pva_save);
}

static double fetch(const double *col, int64_t stride, int64_t n, int64_t i)
{
  return (col && i >= 0 && i < n) ? col[i * stride] : 0.0;   /* beyond the file: read_forcing_data's calloc zeros */
}

int32_t ao_ndays(const vr_atmos_options *o)
{
  int last_hour = (o->starthour + (o->nrecs - 1) * o->time_step) % 24;
  int tmp_nrecs = o->nrecs + o->starthour - 0 + (24 - o->time_step) - last_hour;   /* initialize_atmos.c:262-265 */
  return (tmp_nrecs * o->time_step) / 24;
}

/* One cell.  Raw columns: record r at col[r*stride]; out[f]: record r at out[f][r*stride].
 * dbg (optional, [6][Ndays_local]): daily s_srad, s_dayl, tskc, vp, tminhour, tmaxhour. */
int ao_cell(const vr_atmos_options *o, double lat, double lng, double time_zone_lng, double elevation, double annual_prec,
            double slope, double aspect, double ehoriz, double whoriz, int64_t stride, int64_t nrec_forcing,
            const double *c_prec, const double *c_ta, const double *c_tb, const double *c_wind, const double *c_sw,
            const double *c_lw, const double *c_vp, const double *c_pr, const double *c_de, const double *c_q, const double *c_rh,
            const double *c_rainf, const double *c_snowf, const double *c_we, const double *c_wn, double min_Tfactor,
            double *const out[VR_NFORCING], int32_t *snowflag, double *dbg)
{
  static const int month_days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  const int dt = o->time_step, nrecs = o->nrecs, fdt = o->force_dt, daily = (o->layout == VR_ATMOS_DAILY_TMAX_TMIN);
  const int stepspday = 24 / dt;
  /* SNOW_STEP < TIME_STEP: NF windows of ss hours per record; a record's values follow its windows' (index NR = NF), and
   * when NF == 1 the single window IS the record (NR = 0): get_global_param.c:761-770 */
  const int NF = (o->snow_step > 0 && o->snow_step != dt) ? dt / o->snow_step : 1, ss = dt / NF, NS = NF > 1 ? NF + 1 : 1, NR = NF > 1 ? NF : 0;
  const double min_wind = (double)(float)o->min_wind_speed;    /* option_struct.MIN_WIND_SPEED is a float */
  double *l_q, *l_rh;
  int have_wind = o->has_wind || (c_we && c_wn), vp_supplied;
  double hour_offset = (time_zone_lng - lng) * 24 / 360;       /* :217-225 */
  int hour_offset_int = (hour_offset < 0) ? (int)(hour_offset - 0.5) : (int)(hour_offset + 0.5);
  int Ndays = ao_ndays(o), Ndays_local, local_starthour, local_startday, local_startmonth, local_startyear;
  int day_in_year, year, month, day, rec, hour, idx, i, j, k, shift;
  int *yday, *tmaxhour, *tminhour;
  double *l_sw, *l_lw, *l_pr, *l_de;
  double *l_prec, *l_ta, *l_tb, *l_wind, *l_vp, *hourlyrad, *prec, *tmax, *tmin, *tair, *tskc, *daily_vp, *radfract;
  mt_data d;
  hour_offset -= hour_offset_int;
  Ndays_local = Ndays + (hour_offset_int != 0 ? 1 : 0);       /* :271-272 */
  if ((daily && fdt != 24) || (!daily && (fdt >= 24 || 24 % fdt)) || 24 % dt) return -1;
  if (NF > 1 && (dt < 24 || dt % o->snow_step)) return -1;
  if (!c_vp && c_q && !c_pr && fdt < 24) return -1;            /* the reference reads a stale index there (:1302) */

  local_starthour = o->starthour - hour_offset_int;            /* :274-292 */
  local_startday = o->startday; local_startmonth = o->startmonth; local_startyear = o->startyear;
  if (local_starthour < 0) {
    local_starthour += 24;
    local_startday--;
    if (local_startday < 1) {
      local_startmonth--;
      if (local_startmonth < 1) { local_startmonth = 12; local_startyear--; }
      local_startday = month_days[local_startmonth - 1];
      if (local_startyear % 4 == 0 && local_startmonth == 2) local_startday++;
    }
  }
  yday = calloc(Ndays_local, sizeof(int));                     /* :299-338, one entry per local day */
  day_in_year = local_startday;
  for (month = 1; month < local_startmonth; month++) {
    int dim = month_days[month - 1];
    if (local_startyear % 4 == 0 && month == 2) dim++;
    day_in_year += dim;
  }
  year = local_startyear; month = local_startmonth; day = local_startday;
  for (i = 0; i < Ndays_local; i++) {
    int dim;
    yday[i] = day_in_year;
    day_in_year++; day++;
    dim = month_days[month - 1];
    if (year % 4 == 0 && month == 2) dim++;
    if (day > dim) {
      day = 1; month++;
      if (month > 12) { day_in_year = 1; month = 1; year++; }
    }
  }

  l_prec = calloc(Ndays_local * 24, sizeof(double)); l_ta = calloc(Ndays_local * 24, sizeof(double));
  l_tb = calloc(Ndays_local * 24, sizeof(double)); l_wind = calloc(Ndays_local * 24, sizeof(double));
  l_vp = calloc(Ndays_local * 24, sizeof(double)); hourlyrad = calloc(Ndays_local * 24, sizeof(double));
  l_sw = calloc(Ndays_local * 24, sizeof(double)); l_lw = calloc(Ndays_local * 24, sizeof(double));
  l_pr = calloc(Ndays_local * 24, sizeof(double)); l_de = calloc(Ndays_local * 24, sizeof(double));
  l_q = calloc(Ndays_local * 24, sizeof(double)); l_rh = calloc(Ndays_local * 24, sizeof(double));
  tair = calloc(Ndays_local * 24, sizeof(double));
  prec = calloc(Ndays_local, sizeof(double)); tmax = calloc(Ndays_local, sizeof(double));
  tmin = calloc(Ndays_local, sizeof(double)); tskc = calloc(Ndays_local, sizeof(double));
  daily_vp = calloc(Ndays_local, sizeof(double));
  tmaxhour = calloc(Ndays_local, sizeof(int)); tminhour = calloc(Ndays_local, sizeof(int));
  radfract = malloc(sizeof(double) * 366 * TINYSTEPS);

  /* unit conversion, :374-416: ALMA rates -> amounts per forcing step and K -> C, else kPa -> Pa */
  const double pscale = o->alma_input ? (double)(fdt * 3600) : 1.0, tshift = o->alma_input ? KELVIN : 0.0,
               kpa = o->alma_input ? 1.0 : 1000.;
  /* forcing arrays referenced to local time, :469-540 */
  if (fdt == 24) {
    for (idx = 0; idx < Ndays_local; idx++) {
      i = idx;
      if (hour_offset_int > 0) i--;
      if (i < 0) i = 0;
      if (i >= Ndays) i = Ndays - 1;
      /* :424-460: RAINF + SNOWF replace PREC, |(WIND_E, WIND_N)| replaces WIND (element by element, before the local shift) */
      l_prec[idx] = (c_rainf && c_snowf) ? fetch(c_rainf, stride, nrec_forcing, i) * pscale + fetch(c_snowf, stride, nrec_forcing, i) * pscale
                                         : fetch(c_prec, stride, nrec_forcing, i) * pscale;
      l_ta[idx] = fetch(c_ta, stride, nrec_forcing, i) - tshift;
      l_tb[idx] = fetch(c_tb, stride, nrec_forcing, i) - tshift;
      l_wind[idx] = (c_we && c_wn) ? sqrt(fetch(c_we, stride, nrec_forcing, i) * fetch(c_we, stride, nrec_forcing, i) +
                                          fetch(c_wn, stride, nrec_forcing, i) * fetch(c_wn, stride, nrec_forcing, i))
                                   : fetch(c_wind, stride, nrec_forcing, i);
      l_q[idx] = fetch(c_q, stride, nrec_forcing, i); l_rh[idx] = fetch(c_rh, stride, nrec_forcing, i);
      l_sw[idx] = fetch(c_sw, stride, nrec_forcing, i); l_lw[idx] = fetch(c_lw, stride, nrec_forcing, i);
      l_vp[idx] = fetch(c_vp, stride, nrec_forcing, i) * kpa;         /* kPa2Pa, :404-414 */
      l_pr[idx] = fetch(c_pr, stride, nrec_forcing, i) * kpa; l_de[idx] = fetch(c_de, stride, nrec_forcing, i);
    }
  }
  else {
    int fstepspday = 24 / fdt;
    for (idx = 0; idx < Ndays_local * 24; idx++) {
      i = (idx - o->starthour + hour_offset_int) / fdt;
      if (i < 0) i += fstepspday;
      if (i >= Ndays * fstepspday) i -= fstepspday;
      l_prec[idx] = ((c_rainf && c_snowf) ? fetch(c_rainf, stride, nrec_forcing, i) * pscale + fetch(c_snowf, stride, nrec_forcing, i) * pscale
                                          : fetch(c_prec, stride, nrec_forcing, i) * pscale) / fdt;
      l_ta[idx] = fetch(c_ta, stride, nrec_forcing, i) - tshift;
      l_wind[idx] = (c_we && c_wn) ? sqrt(fetch(c_we, stride, nrec_forcing, i) * fetch(c_we, stride, nrec_forcing, i) +
                                          fetch(c_wn, stride, nrec_forcing, i) * fetch(c_wn, stride, nrec_forcing, i))
                                   : fetch(c_wind, stride, nrec_forcing, i);
      l_q[idx] = fetch(c_q, stride, nrec_forcing, i); l_rh[idx] = fetch(c_rh, stride, nrec_forcing, i);
      l_sw[idx] = fetch(c_sw, stride, nrec_forcing, i); l_lw[idx] = fetch(c_lw, stride, nrec_forcing, i);
      l_vp[idx] = fetch(c_vp, stride, nrec_forcing, i) * kpa;
      l_pr[idx] = fetch(c_pr, stride, nrec_forcing, i) * kpa; l_de[idx] = fetch(c_de, stride, nrec_forcing, i);
    }
  }
  shift = (o->starthour - hour_offset_int < 0) ? 24 : 0;
#define WIN_HOUR(rec, i) ((rec) * dt + (i) * ss + o->starthour - hour_offset_int + shift)
#define DAY_OF(hour) ((int)((float)(hour) / 24.0))
#define O(rec, i) ((size_t)((rec) * NS + (i)) * stride)

  /* precipitation :597-635, wind :641-689, sub-daily air temperature :737-752 */
  for (rec = 0; rec < nrecs; rec++) {
    double sp = 0, sw_ = 0, st = 0;
    for (i = 0; i < NF; i++) {
      hour = WIN_HOUR(rec, i);
      if (fdt == 24) {
        out[VR_F_PREC][O(rec, i)] = l_prec[DAY_OF(hour)] / (float)(NF * stepspday);
        out[VR_F_WIND][O(rec, i)] = have_wind ? l_wind[DAY_OF(hour)] : 3.0;
      }
      else {
        double s = 0;
        for (idx = hour; idx < hour + ss; idx++) s += l_prec[idx];
        out[VR_F_PREC][O(rec, i)] = s;
        if (have_wind) {
          s = 0;
          for (idx = hour; idx < hour + ss; idx++) s += (l_wind[idx] < min_wind) ? min_wind : l_wind[idx];
          out[VR_F_WIND][O(rec, i)] = s / ss;
        }
        else out[VR_F_WIND][O(rec, i)] = 3.0;
        s = 0;
        for (idx = hour; idx < hour + ss; idx++) s += l_ta[idx];
        out[VR_F_AIR_TEMP][O(rec, i)] = s / ss;
        st += out[VR_F_AIR_TEMP][O(rec, i)];
      }
      sp += out[VR_F_PREC][O(rec, i)];
      sw_ += out[VR_F_WIND][O(rec, i)];
    }
    if (NF > 1) {
      out[VR_F_PREC][O(rec, NR)] = sp;
      out[VR_F_WIND][O(rec, NR)] = have_wind ? sw_ / (float)NF : 3.0;
      if (fdt < 24) out[VR_F_AIR_TEMP][O(rec, NR)] = st / (float)NF;
      /* the clamp to MIN_WIND_SPEED at :654-657 addresses wind[j] with j == NF after the loop: with NF > 1 that is the
       * record's value (NR == NF), with NF == 1 one past it (no effect) */
      if (have_wind && fdt == 24 && dt == 24 && out[VR_F_WIND][O(rec, NR)] < min_wind) out[VR_F_WIND][O(rec, NR)] = min_wind;
    }
  }
  for (day = 0; day < Ndays_local; day++) {
    if (fdt == 24) { prec[day] = l_prec[day]; tmax[day] = l_ta[day]; tmin[day] = l_tb[day]; }
    else {
      prec[day] = 0;
      for (hour = 0; hour < 24; hour++) prec[day] += l_prec[day * 24 + hour];
      /* :758-766 -- the reference indexes the FIRST local day's hours for every day */
      tmax[day] = tmin[day] = -9999;
      for (hour = 0; hour < 24; hour++) {
        if (hour >= 9 && (tmax[day] == -9999 || l_ta[hour] > tmax[day])) tmax[day] = l_ta[hour];
        if (hour < 12 && (tmin[day] == -9999 || l_ta[hour] < tmin[day])) tmin[day] = l_ta[hour];
      }
    }
  }

  /* vapour pressure, part 1: QAIR with PRESSURE, REL_HUMID with AIR_TEMP become the observed VP (:773-866; all columns share
   * FORCE_DT here, so the daily / sub-daily cases are the same element-wise products) */
  vp_supplied = c_vp != NULL;
  if (!vp_supplied && c_q && c_pr) {
    for (idx = 0; idx < Ndays_local * (fdt == 24 ? 1 : 24); idx++) l_vp[idx] = l_q[idx] * l_pr[idx] / 0.62196351;
    vp_supplied = 1;
  }
  else if (!vp_supplied && !c_q && c_rh && !daily) {
    for (idx = 0; idx < Ndays_local * 24; idx++) l_vp[idx] = l_rh[idx] * svp(l_ta[idx]) / 100;
    vp_supplied = 1;
  }
  /* (:869-915) and shortwave, part 1 (:925-940) */
  if (vp_supplied) {
    for (day = 0; day < Ndays_local; day++) {
      if (fdt == 24) daily_vp[day] = l_vp[day];
      else {
        daily_vp[day] = 0;
        for (hour = 0; hour < 24; hour++) daily_vp[day] += l_vp[day * 24 + hour];
        daily_vp[day] /= 24;
      }
    }
  }
  if (c_sw)
    for (day = 0; day < Ndays_local; day++)
      for (hour = 0; hour < 24; hour++) hourlyrad[day * 24 + hour] = (fdt == 24) ? l_sw[day] : l_sw[day * 24 + hour];

  /* mtclim_wrapper.c:84-218: init, calc_tair (mtclim_vic.c:405-433), calc_prcp (:437-469), snowpack (:474-530) */
  d.ndays = Ndays_local;
  d.insw = c_sw != NULL; d.invp = vp_supplied;
  d.yday = yday;
  d.tmax = tmax; d.tmin = tmin;
  d.prcp = calloc(Ndays_local, sizeof(double)); d.s_tmax = calloc(Ndays_local, sizeof(double));
  d.s_tmin = calloc(Ndays_local, sizeof(double)); d.s_tday = calloc(Ndays_local, sizeof(double));
  d.s_prcp = calloc(Ndays_local, sizeof(double)); d.s_hum = calloc(Ndays_local, sizeof(double));
  d.s_srad = calloc(Ndays_local, sizeof(double)); d.s_dayl = calloc(Ndays_local, sizeof(double));
  d.s_swe = calloc(Ndays_local, sizeof(double)); d.s_fdir = calloc(Ndays_local, sizeof(double));
  d.s_tskc = calloc(Ndays_local, sizeof(double)); d.s_tfmax = calloc(Ndays_local, sizeof(double));
  {
    double isoh = annual_prec / 10., ratio, dz = (elevation - elevation) / 1000.0, lr = -1 * T_LAPSE * 1000., snowpack, sum;
    int count, start_yday, prev_yday;
    for (i = 0; i < Ndays_local; i++) d.prcp[i] = prec[i] / 10.;
    for (i = 0; i < Ndays_local; i++) {               /* mtclim_wrapper.c:196-204 */
      if (d.insw) {
        d.s_srad[i] = 0;
        for (j = 0; j < 24; j++) d.s_srad[i] += hourlyrad[i * 24 + j];
        d.s_srad[i] /= 24;
      }
      if (d.invp) d.s_hum[i] = daily_vp[i];
    }
    for (i = 0; i < Ndays_local; i++) {
      double tx, tn, tmean;
      d.s_tmax[i] = tx = d.tmax[i] + (dz * lr);
      d.s_tmin[i] = tn = d.tmin[i] + (dz * lr);
      tmean = (tx + tn) / 2.0;
      d.s_tday[i] = ((tx - tmean) * TDAYCOEF) + tmean;
    }
    if (isoh < 1e-10 && isoh < 1e-10) ratio = 1.;
    else ratio = isoh / isoh;
    for (i = 0; i < Ndays_local; i++) d.s_prcp[i] = d.prcp[i] * ratio;
    snowpack = 0.0;
    for (i = 0; i < Ndays_local; i++) {
      double newsnow = 0.0, snowmelt = 0.0;
      if (d.s_tmin[i] <= SNOW_TCRIT) newsnow = d.s_prcp[i];
      else snowmelt = SNOW_TRATE * (d.s_tmin[i] - SNOW_TCRIT);
      snowpack += newsnow - snowmelt;
      if (snowpack < 0.0) snowpack = 0.0;
      d.s_swe[i] = snowpack;
    }
    start_yday = yday[0];
    prev_yday = (start_yday == 1) ? 365 : start_yday - 1;
    count = 0; sum = 0.0;
    for (i = 1; i < Ndays_local; i++)
      if (yday[i] == start_yday || yday[i] == prev_yday) { count++; sum += d.s_swe[i]; }
    if (count) {
      snowpack = sum / (double)count;
      for (i = 0; i < Ndays_local; i++) {
        double newsnow = 0.0, snowmelt = 0.0;
        if (d.s_tmin[i] <= SNOW_TCRIT) newsnow = d.s_prcp[i];
        else snowmelt = SNOW_TRATE * (d.s_tmin[i] - SNOW_TCRIT);
        snowpack += newsnow - snowmelt;
        if (snowpack < 0.0) snowpack = 0.0;
        d.s_swe[i] = snowpack;
      }
    }
  }
  srad_humidity_iterative(o, &d, lat, elevation, slope, aspect, ehoriz, whoriz, radfract);

  /* mtclim_to_vic, mtclim_wrapper.c:221-278 */
  {
    int tinystepsphour = 3600 / SRADDT;
    int tiny_offset = (int)((float)tinystepsphour * hour_offset);
    for (i = 0; i < Ndays_local; i++) {
      double tmp_rad = d.insw ? d.s_srad[i] * 24. : d.s_srad[i] * d.s_dayl[i] / 3600.;
      for (j = 0; j < 24; j++) {
        double s = 0;
        for (k = 0; k < tinystepsphour; k++) {
          int tinystep = j * tinystepsphour + k - tiny_offset;
          if (tinystep < 0) tinystep += 24 * tinystepsphour;
          if (tinystep > 24 * tinystepsphour - 1) tinystep -= 24 * tinystepsphour;
          s += radfract[(size_t)(yday[i] - 1) * TINYSTEPS + tinystep];
        }
        hourlyrad[i * 24 + j] = s * tmp_rad;
      }
      tskc[i] = d.s_tskc[i];
      daily_vp[i] = d.s_hum[i];
    }
  }

  if (c_sw && fdt < 24)                             /* :966-972: sub-daily observations replace MTCLIM's hours */
    for (i = 0; i < Ndays_local * 24; i++) hourlyrad[i] = l_sw[i];
  /* shortwave :974-987 */
  for (rec = 0; rec < nrecs; rec++) {
    double sum = 0;
    for (i = 0; i < NF; i++) {
      double s = 0;
      hour = WIN_HOUR(rec, i);
      for (idx = hour; idx < hour + ss; idx++) s += hourlyrad[idx];
      out[VR_F_SHORTWAVE][O(rec, i)] = s / ss;
      sum += out[VR_F_SHORTWAVE][O(rec, i)];
    }
    if (NF > 1) out[VR_F_SHORTWAVE][O(rec, NR)] = sum / (float)NF;
  }
  set_max_min_hour(hourlyrad, Ndays_local, tmaxhour, tminhour);   /* :998 */
  if (daily) {                                                      /* :1000-1021 */
    hourly_t(Ndays_local, tmaxhour, tmax, tminhour, tmin, tair);
    for (rec = 0; rec < nrecs; rec++) {
      double sum = 0;
      for (i = 0; i < NF; i++) {
        double s = 0;
        hour = WIN_HOUR(rec, i);
        for (idx = hour; idx < hour + ss; idx++) s += tair[idx];
        out[VR_F_AIR_TEMP][O(rec, i)] = s / ss;
        sum += out[VR_F_AIR_TEMP][O(rec, i)];
      }
      if (NF > 1) out[VR_F_AIR_TEMP][O(rec, NR)] = sum / (float)NF;
    }
  }
  /* density :1032-1064, pressure :1070-1146, density from pressure :1152-1170: observed columns per window and their mean for
   * the record; derived values from the window's or the record's own air temperature */
  for (rec = 0; rec < nrecs; rec++) {
    double sde = 0, spr = 0;
    for (i = 0; i <= NF; i++) {
      double ta, p = 0., de = 0.;
      const int k = (i < NF) ? i : NR;
      if (i == NF && NF == 1) break;
      ta = out[VR_F_AIR_TEMP][O(rec, k)];
      if (i < NF) {
        hour = WIN_HOUR(rec, i);
        if (c_de) {
          if (fdt == 24) de = l_de[DAY_OF(hour)];
          else { double s2 = 0; for (idx = hour; idx < hour + ss; idx++) s2 += l_de[idx]; de = s2 / ss; }
          sde += de;
        }
        if (c_pr) {
          if (fdt == 24) p = l_pr[DAY_OF(hour)];
          else { double s2 = 0; for (idx = hour; idx < hour + ss; idx++) s2 += l_pr[idx]; p = s2 / ss; }
          spr += p;
        }
      }
      else { de = sde / (float)NF; p = spr / (float)NF; }
      if (c_pr) { }
      else if (c_de) p = o->plapse ? (KELVIN + ta) * de * RD : (275.0 + ta) * de / 0.003486;
      else if (o->plapse) p = PS_PM * exp(-elevation * GRAV / (RD * (KELVIN + ta + 0.5 * elevation * T_LAPSE)));
      else p = 95500.;
      out[VR_F_PRESSURE][O(rec, k)] = p;
      if (c_de) out[VR_F_DENSITY][O(rec, k)] = de;
      else if (o->plapse) out[VR_F_DENSITY][O(rec, k)] = p / (RD * (KELVIN + ta));
      else out[VR_F_DENSITY][O(rec, k)] = 0.003486 * p / (275.0 + ta);
    }
  }
  /* vapour pressure, part 2 (:1176-1215): daily QAIR / REL_HUMID that could not become VP before MTCLIM replace its daily
   * value, window by window in record order (the last window of a day leaves its value) */
  if (!vp_supplied && fdt == 24 && (c_q || c_rh))
    for (rec = 0; rec < nrecs; rec++)
      for (i = 0; i < NF; i++) {
        idx = DAY_OF(WIN_HOUR(rec, i));
        if (c_q) daily_vp[idx] = l_q[idx] * out[VR_F_PRESSURE][O(rec, i)] / 0.62196351;
        else daily_vp[idx] = l_rh[idx] * svp(out[VR_F_AIR_TEMP][O(rec, i)]) / 100;
      }
  /* vapour pressure :1220-1291 (skipped when sub-daily VP was supplied: l_vp holds the observations) */
  if (vp_supplied && fdt < 24) { }
  else if (o->vp_interp) {
    for (day = 0; day < Ndays_local; day++) {
      double delta_t_minus, delta_t_plus;
      if (day == 0 && Ndays_local == 1) { delta_t_minus = 24; delta_t_plus = 24; }
      else if (day == 0) { delta_t_minus = 24; delta_t_plus = tminhour[day + 1] + 24 - tminhour[day]; }
      else if (day == Ndays_local - 1) { delta_t_minus = tminhour[day] + 24 - tminhour[day - 1]; delta_t_plus = 24; }
      else { delta_t_minus = tminhour[day] + 24 - tminhour[day - 1]; delta_t_plus = tminhour[day + 1] + 24 - tminhour[day]; }
      for (hour = 0; hour < 24; hour++) {
        if (hour < tminhour[day]) {
          if (day > 0)
            l_vp[day * 24 + hour] = daily_vp[day - 1] + (daily_vp[day] - daily_vp[day - 1]) * (hour + 24 - tminhour[day - 1]) / delta_t_minus;
          else l_vp[day * 24 + hour] = daily_vp[day];
        }
        else {
          if (day < Ndays_local - 1)
            l_vp[day * 24 + hour] = daily_vp[day] + (daily_vp[day + 1] - daily_vp[day]) * (hour - tminhour[day]) / delta_t_plus;
          else l_vp[day * 24 + hour] = daily_vp[day];
        }
      }
    }
  }
  else
    for (day = 0; day < Ndays_local; day++)
      for (hour = 0; hour < 24; hour++) l_vp[day * 24 + hour] = daily_vp[day];
  for (rec = 0; rec < nrecs; rec++) {
    double s_vp0 = 0, s_vp = 0, s_vpd = 0, s_lw = 0;
    int flag = 0;
    for (i = 0; i < NF; i++) {
      double s = 0, ta = out[VR_F_AIR_TEMP][O(rec, i)], vp, vpd, lw;
      hour = WIN_HOUR(rec, i);
      for (idx = hour; idx < hour + ss; idx++) s += l_vp[idx];            /* :1278-1291 */
      vp = s / ss;
      s_vp0 += vp;
      vpd = svp(ta) - vp;                                             /* :1336-1355 */
      if (vpd < 0) { vpd = 0; vp = svp(ta); }
      out[VR_F_VP][O(rec, i)] = vp;
      out[VR_F_VPD][O(rec, i)] = vpd;
      s_vp += vp; s_vpd += vpd;
      if (c_lw) {                                        /* :1389-1420 */
        if (fdt == 24) lw = l_lw[DAY_OF(hour)];
        else {
          double sl = 0;
          for (idx = hour; idx < hour + ss; idx++) sl += l_lw[idx];
          lw = sl / ss;
        }
      }
      else lw = calc_longwave(o, tskc[DAY_OF(hour)], ta, vp);   /* :1361-1388 */
      out[VR_F_LONGWAVE][O(rec, i)] = lw;
      s_lw += lw;
      if ((ta + min_Tfactor) < o->max_snow_temp && out[VR_F_PREC][O(rec, i)] > 0) flag = 1;      /* :1736-1753 */
    }
    if (NF > 1) {
      if (vp_supplied || o->vp_interp) {                 /* :1345-1349 */
        out[VR_F_VPD][O(rec, NR)] = s_vpd / (float)NF;
        out[VR_F_VP][O(rec, NR)] = s_vp / (float)NF;
      }
      else {                                             /* :1350-1352: vp[NR] as transferred, no clamp */
        out[VR_F_VP][O(rec, NR)] = s_vp0 / (float)NF;
        out[VR_F_VPD][O(rec, NR)] = svp(out[VR_F_AIR_TEMP][O(rec, NR)]) - out[VR_F_VP][O(rec, NR)];
      }
      out[VR_F_LONGWAVE][O(rec, NR)] = s_lw / (float)NF;
    }
    if (snowflag) snowflag[(size_t)rec * stride] = flag;
  }
  if (dbg) {
    for (i = 0; i < Ndays_local; i++) {
      dbg[0 * Ndays_local + i] = d.s_srad[i]; dbg[1 * Ndays_local + i] = d.s_dayl[i];
      dbg[2 * Ndays_local + i] = tskc[i]; dbg[3 * Ndays_local + i] = daily_vp[i];
      dbg[4 * Ndays_local + i] = tminhour[i]; dbg[5 * Ndays_local + i] = tmaxhour[i];
    }
  }
  free(l_sw); free(l_lw); free(l_pr); free(l_de); free(l_q); free(l_rh);
  free(yday); free(l_prec); free(l_ta); free(l_tb); free(l_wind); free(l_vp); free(hourlyrad); free(tair);
  free(prec); free(tmax); free(tmin); free(tskc); free(daily_vp); free(tmaxhour); free(tminhour); free(radfract);
  free(d.prcp); free(d.s_tmax); free(d.s_tmin); free(d.s_tday); free(d.s_prcp); free(d.s_hum); free(d.s_srad);
  free(d.s_dayl); free(d.s_swe); free(d.s_fdir); free(d.s_tskc); free(d.s_tfmax);
  return 0;
}

/* all cells, serially (cells are independent); arrays as in include/vicres_atmos.h */
int ao_prepare(const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw, double *const out[VR_NFORCING],
               int32_t *snowflag, int64_t c_begin, int64_t c_end)
{
  int64_t C = cells->ncell, c;
#define COL(p) ((p) ? (p) + c : NULL)
  for (c = c_begin; c < c_end; c++) {
    double *oc[VR_NFORCING];
    int f, rc;
    for (f = 0; f < VR_NFORCING; f++) oc[f] = out[f] + c;
    rc = ao_cell(o, cells->lat[c], cells->lng[c], cells->time_zone_lng[c], cells->elevation[c], cells->annual_prec[c],
                 cells->slope ? cells->slope[c] : 0., cells->aspect ? cells->aspect[c] : 0.,
                 cells->ehoriz ? cells->ehoriz[c] : 0., cells->whoriz ? cells->whoriz[c] : 0., C, raw->nrec_forcing,
                 COL(raw->prec), raw->t_a + c, COL(raw->t_b), COL(raw->wind), COL(raw->shortwave), COL(raw->longwave), COL(raw->vp),
                 COL(raw->pressure), COL(raw->density), COL(raw->qair), COL(raw->rel_humid), COL(raw->rainf), COL(raw->snowf),
                 COL(raw->wind_e), COL(raw->wind_n), cells->min_Tfactor ? cells->min_Tfactor[c] : 0., oc, snowflag ? snowflag + c : NULL,
                 NULL);
    if (rc) return rc;
  }
#undef COL
  return 0;
}
