/* vic_oracle.c -- ORACLE = TEST INFRASTRUCTURE.  Never linked into, imported by or executed
 * from the product path (vicres_b200/).  See vic_oracle.h.
 *
 * A plain-C, single-threaded restatement of VIC-Res's per-cell land-surface step in
 * frozen-soil-free water-balance mode (FULL_ENERGY FALSE, FROZEN_SOIL FALSE, QUICK_FLUX TRUE,
 * CLOSE_ENERGY FALSE, SNOW_STEP == TIME_STEP so NF == 1 and NR == 0), written from a reading of
 * /root/reference/runoff/model.  Each function names the reference file:line it follows and keeps
 * that code's operation order, so that with the same libm the results agree with the compiled
 * reference (oracle/_ref) to rounding (tests/test_oracle_vs_ref.py pins this; parity is PINNED
 * against the reference binary on seeded cases and on the bundled-basin cell).
 *
 * Not restated (outside the path): lakes, frozen soil / finite-difference soil
 * heat, carbon cycle, blowing snow, spatial snow/frost, treeline, precipitation correction,
 * the 6 PET reference evaporations (compute_pot_evap.c; they feed only OUT_PET_*), compute_zwt
 * (feeds only OUT_ZWT*).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vic_oracle.h"

/* ---- constants: runoff/model/vicNl_def.h:160-440, snow.h:40-90 ---- */
#define HUGE_RESIST 1.e20
#define SMALL 1.e-12
#define MISSING (-99999.)
#define VERROR (-999)
#define BARE_SOIL_ALBEDO 0.2
#define SECPHOUR 3600
#define SEC_PER_DAY 86400.
#define ice_density 917.
#define T_LAPSE (-0.0065)
#define von_K 0.40
#define KELVIN 273.15
#define STEFAN_B 5.6696e-8
#define Lf 3.337e5
#define RHO_W 999.842594
#define Cp 1013.0
#define CH_ICE 2100.0e3
#define CH_WATER 4186.8e3
#define K_SNOW 2.9302e-6
#define EPS 0.62196351
#define G 9.81
#define JOULESPCAL 4.1868
#define GRAMSPKG 1000.
#define A_SVP 0.61078
#define B_SVP 17.269
#define C_SVP 237.3
#define RATIO_RL_HEIGHT_VEG 0.123
#define RATIO_DH_HEIGHT_VEG 0.67
#define CLOSURE 4000
#define RSMAX 5000
#define VPDMINFACTOR 0.1
#define RARC_SOIL 100.0
#define WINDH_SOIL 10.0
#define CP_PM 1013
#define PS_PM 101300
#define LAI_WATER_FACTOR 0.1
#define MIN_VEGCOVER 0.0001
#define SNOW_DT 5.0
#define LIQUID_WATER_CAPACITY 0.035
#define LAI_SNOW_MULTIPLIER 0.0005
#define MIN_INTERCEPTION_STORAGE 0.005
#define MAX_SURFACE_SWE 0.125
#define NEW_SNOW_DENSITY 50.
#define SNDENS_ETA0 (3.6e6)
#define SNDENS_C5 0.08
#define SNDENS_C6 0.021
#define MIN_SWQ_EB_THRES 0.0010
#define NEW_SNOW_ALB 0.85
#define SNOW_ALB_ACCUM_A 0.94
#define SNOW_ALB_ACCUM_B 0.58
#define SNOW_ALB_THAW_A 0.82
#define SNOW_ALB_THAW_B 0.46
#define TraceSnow 0.03
#define MAXL 3

/* ---- trimmed state structs: only fields live in water-balance mode (vicNl_def.h:1120-1331) ---- */
typedef struct { double moist, evap, bare_evap_frac, kappa, Cs, T; } olayer;
typedef struct {
  olayer layer[MAXL];
  double runoff, baseflow, asat, inflow, aero_resist[2], rootmoist, wetness;
  double pot_evap[6], zwt, zwt_lumped, layer_zwt[MAXL];
} ocell;
typedef struct { double Wdew, canopyevap, throughfall, LAI, Wdmax, vegcover, albedo, rc; } ovegvar;
typedef struct {
  double swq, surf_temp, pack_temp, surf_water, pack_water, density, depth, albedo, coldcontent,
         coverage, snow_canopy, tmp_int_storage, vapor_flux, canopy_vapor_flux, melt,
         blowing_flux, surface_flux, mass_error;
  int last_snow, MELTING, snow, surf_temp_fbflag, surf_temp_fbcount;
} osnow;
typedef struct {
  double AlbedoOver, AlbedoUnder, AtmosLatent, AtmosLatentSub, AtmosSensible, LongOverIn,
         LongUnderIn, LongUnderOut, NetLongAtmos, NetLongOver, NetLongUnder, NetShortAtmos,
         NetShortGrnd, NetShortOver, NetShortUnder, ShortOverIn, ShortUnderIn, T[3], Tcanopy,
         Tfoliage, Tsurf, advected_sensible, advection, canopy_advection, canopy_latent,
         canopy_latent_sub, canopy_refreeze, canopy_sensible, deltaCC, deltaH, error, fusion,
         grnd_flux, kappa[2], Cs[2], latent, latent_sub, melt_energy, refreeze_energy, sensible,
         snow_flux;
  int Tfoliage_fbflag, Tfoliage_fbcount;
} oenergy;
typedef struct { ocell cell; ovegvar veg; osnow snow; oenergy energy; } ounit;

typedef struct {
  int overstory;
  double rarc, rmin, wind_h, RGL, rad_atten, wind_atten, trunk_ratio, roughness[12], displacement[12];
  double LAI[12], albedo[12];      /* library values: read by compute_pot_evap only */
} oveglib;

typedef struct {   /* soil_con_struct subset (vicNl_def.h:937-1008) for one cell */
  double lat, elevation, b_infilt, Ds, Dsmax, Ws, c, avg_temp, dp, rough, snow_rough;
  double expt[MAXL], Ksat[MAXL], depth[MAXL], max_moist[MAXL], Wcr[MAXL], Wpwp[MAXL],
         resid_moist[MAXL], init_moist[MAXL], bulk_density[MAXL], soil_density[MAXL],
         bulk_dens_min[MAXL], soil_dens_min[MAXL], quartz[MAXL], organic[MAXL], porosity[MAXL];
  double AreaFract[VR_MAX_BANDS], Tfactor[VR_MAX_BANDS], Pfactor[VR_MAX_BANDS];
  int have_zwt;
  double zwtvmoist_zwt[MAXL + 2][VR_ZWT_NPOINTS], zwtvmoist_moist[MAXL + 2][VR_ZWT_NPOINTS];
} osoil;

typedef struct {   /* one element of a record of atmos_data_struct: a sub-step [0..NF-1] or the record itself [NR] (vicNl_def.h:1090-1113) */
  double air_temp, prec, wind, vp, vpd, pressure, density, shortwave, longwave;
  int month, day_in_year, snowflag;
} oatmos;

struct vo_model {
  vr_options opt;
  int nclass;
  oveglib *lib;              /* [nclass+1]; entry nclass = bare-soil pseudo class (read_veglib.c:136-153) */
  int64_t C, T;
  osoil *soil;               /* [C] */
  int64_t *tile_ptr;         /* [C+1] */
  int32_t *tile_class;       /* [T] */
  double *tile_Cv, *tile_root, *tile_LAI, *tile_albedo, *tile_vegcover;
  ounit *unit;               /* [T * nbands] : unit (t, b) at t*nbands + b */
  /* save_data_struct + calc_water_balance_error statics, per cell (put_data.c:586-632) */
  double *save_soil, *save_swe, *save_wdew, *last_storage;
  int initialised;
  int want_pet;              /* an OUT_PET_* variable is requested: run the reference-cover passes */
};

static int NLAYER;   /* options.Nlayer; set per call from the model (single model per process use) */

/* ======================================================================================
 * L0 math
 * ====================================================================================== */

/* svp.c:7-34 */
static double svp(double temp)
{
  double SVP = A_SVP * exp((B_SVP * temp) / (C_SVP + temp));
  if (temp < 0) SVP *= 1.0 + .00972 * temp + .000042 * temp * temp;
  return SVP * 1000.;
}
static double svp_slope(double temp)
{
  return (B_SVP * C_SVP) / ((C_SVP + temp) * (C_SVP + temp)) * svp(temp);
}

/* calc_rainonly.c:8-58 */
static double calc_rainonly(double air_temp, double prec, double MAX_SNOW_TEMP, double MIN_RAIN_TEMP)
{
  double rainonly = 0.;
  if (air_temp < MAX_SNOW_TEMP && air_temp > MIN_RAIN_TEMP)
    rainonly = (air_temp - MIN_RAIN_TEMP) / (MAX_SNOW_TEMP - MIN_RAIN_TEMP) * prec;
  else if (air_temp >= MAX_SNOW_TEMP)
    rainonly = prec;
  return rainonly;
}

/* penman.c:44-91 (Jarvis branch only; ref_crop FALSE on this path) */
static double calc_rc(double rs, double net_short, float RGL, double tair, double vpd, double lai,
                      double gsm_inv)
{
  double f, DAYfactor, Tfactor, vpdfactor, rc;
  if (rs == 0) rc = 0;
  else if (lai == 0) rc = RSMAX;
  else {
    if (rs > 0.) {
      f = net_short / RGL;
      DAYfactor = (1. + f) / (f + rs / RSMAX);
    }
    else DAYfactor = 1.;
    Tfactor = .08 * tair - 0.0016 * tair * tair;
    Tfactor = (Tfactor <= 0.0) ? 1e-10 : Tfactor;
    vpdfactor = 1 - vpd / CLOSURE;
    vpdfactor = (vpdfactor < VPDMINFACTOR) ? VPDMINFACTOR : vpdfactor;
    rc = rs / (lai * gsm_inv * Tfactor * vpdfactor) * DAYfactor;
    rc = (rc > RSMAX) ? RSMAX : rc;
  }
  return rc;
}

/* penman.c:201-249 */
static double penman(double tair, double elevation, double rad, double vpd, double ra, double rc,
                     double rarc)
{
  double evap, slope, r_air, h, lv, pz, gamma;
  slope = svp_slope(tair);
  h = 287 / 9.81 * ((tair + 273.15) + 0.5 * (double)elevation * T_LAPSE);
  pz = PS_PM * exp(-(double)elevation / h);
  lv = 2501000 - 2361 * tair;
  gamma = 1628.6 * pz / lv;
  r_air = 0.003486 * pz / (275 + tair);
  evap = (slope * rad + r_air * CP_PM * vpd / ra) / (lv * (slope + gamma * (1 + (rc + rarc) / ra))) *
         SEC_PER_DAY;
  if (vpd >= 0.0 && evap < 0.0) evap = 0.0;
  return evap;
}

/* StabilityCorrection.c:44-81 */
static double StabilityCorrection(double Z, double d, double TSurf, double Tair, double Wind, double Z0)
{
  double Correction = 1.0, Ri, RiCr = 0.2, RiLimit;
  if (TSurf != Tair) {
    Ri = G * (Tair - TSurf) * (Z - d) / (((Tair + 273.15) + (TSurf + 273.15)) / 2.0 * Wind * Wind);
    RiLimit = (Tair + 273.15) / (((Tair + 273.15) + (TSurf + 273.15)) / 2.0 * (log((Z - d) / Z0) + 5));
    if (Ri > RiLimit) Ri = RiLimit;
    if (Ri > 0.0) Correction = (1 - Ri / RiCr) * (1 - Ri / RiCr);
    else {
      if (Ri < -0.5) Ri = -0.5;
      Correction = sqrt(1 - 16 * Ri);
    }
  }
  return Correction;
}

/* CalcAerodynamic.c:72-260.  U[0] holds the wind on entry. */
static int CalcAerodynamic(int OverStory, double Height, double Trunk, double Z0_SNOW, double Z0_SOIL,
                           double n, double *Ra, double *U, double *displacement,
                           double *ref_height, double *roughness)
{
  double d_Lower, d_Upper, K2, Uh, Ut, Uw, Z0_Lower, Z0_Upper, Zt, Zw, tmp_wind;
  tmp_wind = U[0];
  K2 = von_K * von_K;
  if (!OverStory) {
    Z0_Lower = roughness[0];
    d_Lower = displacement[0];
    U[0] = log((2. + Z0_Lower) / Z0_Lower) / log((ref_height[0] - d_Lower) / Z0_Lower);
    Ra[0] = log((2. + (1.0 / 0.63 - 1.0) * d_Lower) / Z0_Lower) *
            log((2. + (1.0 / 0.63 - 1.0) * d_Lower) / (0.1 * Z0_Lower)) / K2;
    U[1] = U[0];
    Ra[1] = Ra[0];
    ref_height[1] = ref_height[0];
    roughness[1] = roughness[0];
    displacement[1] = displacement[0];
    U[2] = log((2. + Z0_SNOW) / Z0_SNOW) / log(ref_height[0] / Z0_SNOW);
    Ra[2] = log((2. + Z0_SNOW) / Z0_SNOW) * log(ref_height[0] / Z0_SNOW) / K2;
    ref_height[2] = 2. + Z0_SNOW;
    roughness[2] = Z0_SNOW;
    displacement[2] = 0.;
  }
  else {
    Z0_Upper = roughness[0];
    d_Upper = displacement[0];
    Z0_Lower = Z0_SOIL;
    d_Lower = 0;
    Zw = 1.5 * Height - 0.5 * d_Upper;
    Zt = Trunk * Height;
    if (Zt < (Z0_Lower + d_Lower)) return VERROR;
    Ra[1] = log((ref_height[0] - d_Upper) / Z0_Upper) / K2 *
            (Height / (n * (Zw - d_Upper)) * (exp(n * (1 - (d_Upper + Z0_Upper) / Height)) - 1) +
             (Zw - Height) / (Zw - d_Upper) + log((ref_height[0] - d_Upper) / (Zw - d_Upper)));
    Uw = log((Zw - d_Upper) / Z0_Upper) / log((ref_height[0] - d_Upper) / Z0_Upper);
    Uh = Uw - (1 - (Height - d_Upper) / (Zw - d_Upper)) / log((ref_height[0] - d_Upper) / Z0_Upper);
    U[1] = Uh * exp(n * ((Z0_Upper + d_Upper) / Height - 1.));
    Ut = Uh * exp(n * (Zt / Height - 1.));
    U[0] = log((2. + Z0_Upper) / Z0_Upper) / log((ref_height[0] - d_Upper) / Z0_Upper);
    Ra[0] = log((2. + (1.0 / 0.63 - 1.0) * d_Upper) / Z0_Upper) *
            log((2. + (1.0 / 0.63 - 1.0) * d_Upper) / (0.1 * Z0_Upper)) / K2;
    if (Zt > (2. + Z0_SNOW)) {
      U[2] = Ut * log((2. + Z0_SNOW) / Z0_SNOW) / log(Zt / Z0_SNOW);
      Ra[2] = log((2. + Z0_SNOW) / Z0_SNOW) * log(Zt / Z0_SNOW) / (K2 * Ut);
    }
    else if (Height > (2. + Z0_SNOW)) {
      U[2] = Uh * exp(n * ((2. + Z0_SNOW) / Height - 1.));
      Ra[2] = log(Zt / Z0_SNOW) * log(Zt / Z0_SNOW) / (K2 * Ut) +
              Height * log((ref_height[0] - d_Upper) / Z0_Upper) / (n * K2 * (Zw - d_Upper)) *
              (exp(n * (1 - Zt / Height)) - exp(n * (1 - (Z0_SNOW + 2.) / Height)));
    }
    else {
      U[2] = Uh;
      Ra[2] = log(Zt / Z0_SNOW) * log(Zt / Z0_SNOW) / (K2 * Ut) +
              Height * log((ref_height[0] - d_Upper) / Z0_Upper) / (n * K2 * (Zw - d_Upper)) *
              (exp(n * (1 - Zt / Height)) - 1);
    }
    ref_height[1] = ref_height[0];
    roughness[1] = roughness[0];
    displacement[1] = displacement[0];
    ref_height[0] = 2.;
    roughness[0] = Z0_Lower;
    displacement[0] = d_Lower;
    ref_height[2] = 2. + Z0_SNOW;
    roughness[2] = Z0_SNOW;
    displacement[2] = 0.;
  }
  if (tmp_wind > 0.) {
    U[0] *= tmp_wind; Ra[0] /= tmp_wind;
    if (U[1] != -999) { U[1] *= tmp_wind; Ra[1] /= tmp_wind; }
    if (U[2] != -999) { U[2] *= tmp_wind; Ra[2] /= tmp_wind; }
  }
  else {
    U[0] *= tmp_wind; Ra[0] = HUGE_RESIST;
    if (U[1] != -999) U[1] *= tmp_wind;
    Ra[1] = HUGE_RESIST;
    if (U[2] != -999) U[2] *= tmp_wind;
    Ra[2] = HUGE_RESIST;
  }
  return 0;
}

/* estimate_T1.c:8-47 */
static double estimate_T1(double Ts, double T1_old, double T2, double D1, double D2, double kappa1,
                          double kappa2, double Cs1, double Cs2, double dp, double delta_t)
{
  double C1, C2, C3, T1;
  C1 = Cs2 * dp / D2 * (1. - exp(-D2 / dp));
  C2 = -(1. - exp(D1 / dp)) * exp(-D2 / dp);
  C3 = kappa1 / D1 - kappa2 / D1 + kappa2 / D1 * exp(-D1 / dp);
  T1 = (kappa1 / 2. / D1 / D2 * (Ts) + C1 / delta_t * T1_old +
        (2. * C2 - 1. + exp(-D1 / dp)) * kappa2 / 2. / D1 / D2 * T2) /
       (C1 / delta_t + kappa2 / D1 / D2 * C2 + C3 / 2. / D2);
  return T1;
}

/* soil_conduction.c:5-49 (ice == 0 on this path, so Wu == moist) */
static double soil_conductivity(double moist, double Wu, double soil_dens_min, double bulk_dens_min,
                                double quartz, double soil_density, double bulk_density, double organic)
{
  double Ke, Ki = 2.2, Kw = 0.57, Ksat, Kdry, Kdry_org = 0.05, Kdry_min, Ks, Ks_org = 0.25, Ks_min, Sr, K,
         porosity;
  Kdry_min = (0.135 * bulk_dens_min + 64.7) / (soil_dens_min - 0.947 * bulk_dens_min);
  Kdry = (1 - organic) * Kdry_min + organic * Kdry_org;
  if (moist > 0.) {
    porosity = 1.0 - bulk_density / soil_density;
    Sr = moist / porosity;
    if (quartz < .2) Ks_min = pow(7.7, quartz) * pow(3.0, 1.0 - quartz);
    else Ks_min = pow(7.7, quartz) * pow(2.2, 1.0 - quartz);
    Ks = (1 - organic) * Ks_min + organic * Ks_org;
    if (Wu == moist) {
      Ksat = pow(Ks, 1.0 - porosity) * pow(Kw, porosity);
      Ke = 0.7 * log10(Sr) + 1.0;
    }
    else {
      Ksat = pow(Ks, 1.0 - porosity) * pow(Ki, porosity - Wu) * pow(Kw, Wu);
      Ke = Sr;
    }
    K = (Ksat - Kdry) * Ke + Kdry;
    if (K < Kdry) K = Kdry;
  }
  else K = Kdry;
  return K;
}

/* soil_conduction.c:50-61 */
static double volumetric_heat_capacity(double soil_fract, double water_fract, double ice_fract,
                                       double organic_fract)
{
  double Cs;
  Cs = 2.0e6 * soil_fract * (1 - organic_fract);
  Cs += 2.7e6 * soil_fract * organic_fract;
  Cs += 4.2e6 * water_fract;
  Cs += 1.9e6 * ice_fract;
  Cs += 1.3e3 * (1. - (soil_fract + water_fract + ice_fract));
  return Cs;
}

/* soil_conduction.c:336-362 */
static void compute_soil_layer_thermal_properties(olayer *layer, const osoil *s, int Nlayers)
{
  int lidx;
  double moist, ice;
  for (lidx = 0; lidx < Nlayers; lidx++) {
    moist = layer[lidx].moist / s->depth[lidx] / 1000;
    ice = 0;
    ice += 0. / s->depth[lidx] / 1000 * 1.;
    layer[lidx].kappa = soil_conductivity(moist, moist - ice, s->soil_dens_min[lidx], s->bulk_dens_min[lidx],
                                          s->quartz[lidx], s->soil_density[lidx], s->bulk_density[lidx],
                                          s->organic[lidx]);
    layer[lidx].Cs = volumetric_heat_capacity(s->bulk_density[lidx] / s->soil_density[lidx], moist - ice, ice,
                                              s->organic[lidx]);
  }
}

/* soil_conduction.c:290-335, FROZEN_SOIL FALSE: layer temperatures only (ice stays 0) */
static void estimate_layer_ice_content_quick_flux(olayer *layer, const osoil *s, double Tsurf, double T1)
{
  int lidx;
  double Lsum[MAXL + 1], Dp = s->dp, Tp = s->avg_temp;
  Lsum[0] = 0;
  for (lidx = 1; lidx <= NLAYER; lidx++) Lsum[lidx] = s->depth[lidx - 1] + Lsum[lidx - 1];
  layer[0].T = 0.5 * (Tsurf + T1);
  for (lidx = 1; lidx < NLAYER; lidx++)
    layer[lidx].T = Tp - Dp / (s->depth[lidx]) * (T1 - Tp) *
                    (exp(-(Lsum[lidx + 1] - Lsum[1]) / Dp) - exp(-(Lsum[lidx] - Lsum[1]) / Dp));
}

/* ======================================================================================
 * root_brent.c:105-367 -- iteration-for-iteration (bracket expansion by TSTEP, then Brent)
 * ====================================================================================== */
typedef double (*brent_fn)(double x, void *ctx);
#define B_MAXTRIES 5
#define B_MAXITER 1000
#define B_MACHEPS 3e-8
#define B_TSTEP 10
#define B_T 1e-7
static double root_brent(double LowerBound, double UpperBound, brent_fn Function, void *ctx)
{
  double a, b, c = 0, d = 0, e = 0, fa, fb, fc, m, p, q, r, s, tol, last_bad = 0, last_good = 0;
  int which_err, i, j;
  a = LowerBound; b = UpperBound;
  fa = Function(a, ctx);
  fb = Function(b, ctx);
  which_err = 0;
  if (fa == VERROR && fb == VERROR) return VERROR;
  if (fa == VERROR || fb == VERROR) {
    if (fa == VERROR) { which_err = -1; last_bad = a; last_good = b; }
    else { which_err = 1; last_good = a; last_bad = b; }
    c = 0.5 * (last_bad + last_good);
    fc = Function(c, ctx);
    j = 0;
    while (fc == VERROR && j < B_MAXITER) {
      last_bad = c;
      c = 0.5 * (last_bad + last_good);
      fc = Function(c, ctx);
      j++;
    }
    if (fc == VERROR) return VERROR;
    if (which_err == -1) { a = c; fa = fc; }
    else { b = c; fb = fc; }
  }
  j = 0;
  while ((fa * fb) >= 0 && j < B_MAXTRIES) {
    if (which_err == 0) {
      a -= B_TSTEP; b += B_TSTEP;
      fa = Function(a, ctx);
      fb = Function(b, ctx);
    }
    else {
      if (which_err == -1) {
        b += B_TSTEP;
        fb = Function(b, ctx);
        if (fb == VERROR) return VERROR;
        last_good = a;
      }
      else {
        a -= B_TSTEP;
        fa = Function(a, ctx);
        if (fa == VERROR) return VERROR;
        last_good = b;
      }
      c = 0.5 * (last_good + last_bad);
      fc = Function(c, ctx);
      i = 0;
      while (fc == VERROR && i < B_MAXITER) {
        last_bad = c;
        c = 0.5 * (last_bad + last_good);
        fc = Function(c, ctx);
        i++;
      }
      if (fc == VERROR) return VERROR;
      if (which_err == -1) { a = c; fa = fc; }
      else { b = c; fb = fc; }
    }
    j++;
  }
  if ((fa * fb) >= 0) return VERROR;
  fc = fb;
  for (i = 0; i < B_MAXITER; i++) {
    if (fb * fc > 0) { c = a; fc = fa; d = b - a; e = d; }
    if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    tol = 2 * B_MACHEPS * fabs(b) + B_T;
    m = 0.5 * (c - b);
    if (fabs(m) <= tol || fb == 0) return b;
    if (fabs(e) < tol || fabs(fa) <= fabs(fb)) { d = m; e = d; }
    else {
      s = fb / fa;
      if (a == c) { p = 2 * m * s; q = 1 - s; }
      else {
        q = fa / fc; r = fb / fc;
        p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
        q = (q - 1) * (r - 1) * (s - 1);
      }
      if (p > 0) q = -q; else p = -p;
      s = e; e = d;
      if ((2 * p) < (3 * m * q - fabs(tol * q)) && p < fabs(0.5 * s * q)) d = p / q;
      else { d = m; e = d; }
    }
    a = b; fa = fb;
    b += (fabs(d) > tol) ? d : ((m > 0) ? tol : -tol);
    fb = Function(b, ctx);
    if (fb == VERROR) return VERROR;
  }
  return VERROR;
}

/* ======================================================================================
 * snow
 * ====================================================================================== */

/* snow_utility.c new_snow_density (DENS_BRAS branch, the default: initialize_global.c:189) */
static double new_snow_density(double air_temp)
{
  double density_new;
  air_temp = air_temp * 9. / 5. + 32.;
  if (air_temp > 0) density_new = (double)NEW_SNOW_DENSITY + 1000. * (air_temp / 100.) * (air_temp / 100.);
  else density_new = (double)NEW_SNOW_DENSITY;
  return density_new;
}

/* snow_utility.c snow_density, DENS_BRAS branch */
#define MAX_CHANGE 0.9
static double snow_density(const osnow *snow, double new_snow, double sswq, double Tgrnd, double Tair,
                           double dt)
{
  double density_new, density = 0.0, depth, swq, Tavg, overburden, viscosity, delta_depth, depth_new;
  (void)Tgrnd;
  if (new_snow > 0.) density_new = new_snow_density(Tair);
  else density_new = 0.0;
  Tavg = snow->surf_temp + KELVIN;
  depth = snow->depth;
  swq = sswq;
  if (new_snow > 0) {
    if (depth > 0.) {
      delta_depth = (((new_snow / 25.4) * (depth / 0.0254)) / (swq / 0.0254) *
                     pow((depth / 0.0254) / 10., 0.35)) * 0.0254;
      if (delta_depth > MAX_CHANGE * depth) delta_depth = MAX_CHANGE * depth;
      depth_new = new_snow / density_new;
      depth = depth - delta_depth + depth_new;
      swq += new_snow / 1000.;
      density = 1000. * swq / depth;
    }
    else {
      density = density_new;
      swq += new_snow / 1000.;
      depth = 1000. * swq / density;
    }
  }
  else density = 1000. * swq / snow->depth;
  if (depth > 0.) {
    overburden = 0.5 * G * RHO_W * swq;
    viscosity = SNDENS_ETA0 * exp(-SNDENS_C5 * (Tavg - KELVIN) + SNDENS_C6 * density);
    delta_depth = overburden / viscosity * depth * dt * SECPHOUR;
    if (delta_depth > MAX_CHANGE * depth) delta_depth = MAX_CHANGE * depth;
    depth -= delta_depth;
    density = 1000. * swq / depth;
  }
  return density;
}

/* snow_utility.c snow_albedo */
static double snow_albedo(double new_snow, double swq, double depth, double albedo, double cold_content,
                          double dt, int last_snow, int MELTING)
{
  (void)depth;
  if (new_snow > TraceSnow && cold_content < 0.0) albedo = NEW_SNOW_ALB;
  else if (swq > 0.0) {
    if (cold_content < 0.0 && !MELTING)
      albedo = NEW_SNOW_ALB * pow(SNOW_ALB_ACCUM_A, pow((double)last_snow * dt / 24., SNOW_ALB_ACCUM_B));
    else
      albedo = NEW_SNOW_ALB * pow(SNOW_ALB_THAW_A, pow((double)last_snow * dt / 24., SNOW_ALB_THAW_B));
  }
  else albedo = 0;
  return albedo;
}

/* latent_heat_from_snow.c:8-68 */
static void latent_heat_from_snow(double AirDens, double EactAir, double Lv, double Press, double Ra,
                                  double TMean, double Vpd, double *LatentHeat,
                                  double *LatentHeatSublimation, double *VaporMassFlux,
                                  double *BlowingMassFlux, double *SurfaceMassFlux)
{
  double EsSnow, Ls;
  EsSnow = svp(TMean);
  *SurfaceMassFlux = AirDens * (EPS / Press) * (EactAir - EsSnow) / Ra;
  if (Vpd == 0.0 && *SurfaceMassFlux < 0.0) *SurfaceMassFlux = 0.0;
  *VaporMassFlux = *SurfaceMassFlux + *BlowingMassFlux;
  if (TMean >= 0.0) {
    *LatentHeat = Lv * (*VaporMassFlux);
    *LatentHeatSublimation = 0;
  }
  else {
    Ls = (677. - 0.07 * TMean) * JOULESPCAL * GRAMSPKG;
    *LatentHeatSublimation = Ls * (*VaporMassFlux);
    *LatentHeat = 0;
  }
}

/* SnowPackEnergyBalance.c:88-328: the varargs list of the reference becomes this context */
typedef struct {
  double Dt, Ra, *Ra_used, Z, Z0_2, AirDens, EactAir, LongSnowIn, Lv, Press, Rain, NetShortUnder, Vpd,
         Wind, OldTSurf, SnowCoverFract, SnowDepth, SnowDensity, SurfaceLiquidWater, SweSurfaceLayer,
         Tair, TGrnd;
  double *AdvectedEnergy, *AdvectedSensibleHeat, *DeltaColdContent, *GroundFlux, *LatentHeat,
         *LatentHeatSub, *NetLongUnder, *RefreezeEnergy, *SensibleHeat, *vapor_flux, *blowing_flux,
         *surface_flux;
} spe_ctx;

static double SnowPackEnergyBalance(double TSurf, void *vctx)
{
  spe_ctx *x = (spe_ctx *)vctx;
  double Density, NetRad, RestTerm, TMean, Tmp, VaporMassFlux, BlowingMassFlux, SurfaceMassFlux;
  TMean = TSurf;
  Density = RHO_W;
  if (x->Wind > 0.0) x->Ra_used[0] = x->Ra / StabilityCorrection(x->Z, 0.f, TMean, x->Tair, x->Wind, x->Z0_2);
  else x->Ra_used[0] = HUGE_RESIST;
  Tmp = TMean + KELVIN;
  *x->NetLongUnder = x->LongSnowIn - STEFAN_B * Tmp * Tmp * Tmp * Tmp;
  NetRad = x->NetShortUnder + *x->NetLongUnder;
  *x->SensibleHeat = x->AirDens * Cp * (x->Tair - TMean) / x->Ra_used[0];
  *x->AdvectedSensibleHeat = 0;      /* SPATIAL_SNOW FALSE */
  VaporMassFlux = *x->vapor_flux * Density / x->Dt;
  BlowingMassFlux = *x->blowing_flux * Density / x->Dt;
  SurfaceMassFlux = *x->surface_flux * Density / x->Dt;
  latent_heat_from_snow(x->AirDens, x->EactAir, x->Lv, x->Press, x->Ra_used[0], TMean, x->Vpd,
                        x->LatentHeat, x->LatentHeatSub, &VaporMassFlux, &BlowingMassFlux,
                        &SurfaceMassFlux);
  *x->vapor_flux = VaporMassFlux * x->Dt / Density;
  *x->blowing_flux = BlowingMassFlux * x->Dt / Density;
  *x->surface_flux = SurfaceMassFlux * x->Dt / Density;
  if (TMean == 0) *x->AdvectedEnergy = (CH_WATER * (x->Tair) * x->Rain) / (x->Dt);
  else *x->AdvectedEnergy = 0.;
  *x->DeltaColdContent = CH_ICE * x->SweSurfaceLayer * (TSurf - x->OldTSurf) / (x->Dt);
  if (x->SnowDepth > 0.)
    *x->GroundFlux = 2.9302e-6 * x->SnowDensity * x->SnowDensity * (x->TGrnd - TMean) / x->SnowDepth / (x->Dt);
  else *x->GroundFlux = 0;
  *x->DeltaColdContent -= *x->GroundFlux;
  RestTerm = NetRad + *x->SensibleHeat + *x->LatentHeat + *x->LatentHeatSub + *x->AdvectedEnergy +
             *x->GroundFlux - *x->DeltaColdContent + *x->AdvectedSensibleHeat;
  *x->RefreezeEnergy = (x->SurfaceLiquidWater * Lf * Density) / (x->Dt);
  if (TSurf == 0.0 && RestTerm > -(*x->RefreezeEnergy)) {
    *x->RefreezeEnergy = -RestTerm;
    RestTerm = 0.0;
  }
  else RestTerm += *x->RefreezeEnergy;
  return RestTerm;
}

/* snow_melt.c:119-579 */
static int snow_melt(double Le, double NetShortSnow, double Tcanopy, double Tgrnd, double *Z0,
                     double aero_resist, double *aero_resist_used, double air_temp, double coverage,
                     double delta_t, double density, double grnd_flux_in, double LongSnowIn,
                     double pressure, double rainfall, double snowfall, double vp, double vpd, double wind,
                     double z2, double *NetLongSnow, double *OldTSurf, double *melt, double *save_Qnet,
                     double *save_advected_sensible, double *save_advection, double *save_deltaCC,
                     double *save_grnd_flux, double *save_latent, double *save_latent_sub,
                     double *save_refreeze_energy, double *save_sensible, int UNSTABLE_SNOW, osnow *snow)
{
  double DeltaPackCC, DeltaPackSwq, Ice, InitialSwq, MassBalanceError, MaxLiquidWater, PackCC, PackSwq,
         Qnet, RefreezeEnergy, RefrozenWater, SnowFallCC, SnowMelt, SurfaceCC, SurfaceSwq, SnowFall,
         RainFall, advection, deltaCC, grnd_flux, latent_heat, latent_heat_sub, sensible_heat,
         advected_sensible_heat, melt_energy = 0., PackRefreezeEnergy;
  spe_ctx x;
  (void)grnd_flux_in;
  SnowFall = snowfall / 1000.;
  RainFall = rainfall / 1000.;
  InitialSwq = snow->swq;
  (*OldTSurf) = snow->surf_temp;
  Ice = snow->swq - snow->pack_water - snow->surf_water;
  if (Ice > MAX_SURFACE_SWE) SurfaceSwq = MAX_SURFACE_SWE;
  else SurfaceSwq = Ice;
  PackSwq = Ice - SurfaceSwq;
  SurfaceCC = CH_ICE * SurfaceSwq * snow->surf_temp;
  PackCC = CH_ICE * PackSwq * snow->pack_temp;
  if (air_temp > 0.0) SnowFallCC = 0.0;
  else SnowFallCC = CH_ICE * SnowFall * air_temp;
  if (SnowFall > (MAX_SURFACE_SWE - SurfaceSwq) && (MAX_SURFACE_SWE - SurfaceSwq) > SMALL) {
    DeltaPackSwq = SurfaceSwq + SnowFall - MAX_SURFACE_SWE;
    if (DeltaPackSwq > SurfaceSwq)
      DeltaPackCC = SurfaceCC + (SnowFall - MAX_SURFACE_SWE) / SnowFall * SnowFallCC;
    else
      DeltaPackCC = DeltaPackSwq / SurfaceSwq * SurfaceCC;
    SurfaceSwq = MAX_SURFACE_SWE;
    SurfaceCC += SnowFallCC - DeltaPackCC;
    PackSwq += DeltaPackSwq;
    PackCC += DeltaPackCC;
  }
  else {
    SurfaceSwq += SnowFall;
    SurfaceCC += SnowFallCC;
  }
  if (SurfaceSwq > 0.0) snow->surf_temp = SurfaceCC / (CH_ICE * SurfaceSwq);
  else snow->surf_temp = 0.0;
  if (PackSwq > 0.0) snow->pack_temp = PackCC / (CH_ICE * PackSwq);
  else snow->pack_temp = 0.0;
  Ice += SnowFall;
  snow->surf_water += RainFall;

  x.Dt = delta_t; x.Ra = aero_resist; x.Ra_used = aero_resist_used; x.Z = z2; x.Z0_2 = Z0[2];
  x.AirDens = density; x.EactAir = vp; x.LongSnowIn = LongSnowIn; x.Lv = Le; x.Press = pressure;
  x.Rain = RainFall; x.NetShortUnder = NetShortSnow; x.Vpd = vpd; x.Wind = wind; x.OldTSurf = *OldTSurf;
  x.SnowCoverFract = coverage; x.SnowDepth = snow->depth; x.SnowDensity = snow->density;
  x.SurfaceLiquidWater = snow->surf_water; x.SweSurfaceLayer = SurfaceSwq; x.Tair = Tcanopy;
  x.TGrnd = Tgrnd;
  x.AdvectedEnergy = &advection; x.AdvectedSensibleHeat = &advected_sensible_heat;
  x.DeltaColdContent = &deltaCC; x.GroundFlux = &grnd_flux; x.LatentHeat = &latent_heat;
  x.LatentHeatSub = &latent_heat_sub; x.NetLongUnder = NetLongSnow; x.RefreezeEnergy = &RefreezeEnergy;
  x.SensibleHeat = &sensible_heat; x.vapor_flux = &snow->vapor_flux; x.blowing_flux = &snow->blowing_flux;
  x.surface_flux = &snow->surface_flux;

  Qnet = SnowPackEnergyBalance(0.0, &x);
  if (!UNSTABLE_SNOW) {
    if (Qnet == 0.0) {
      snow->surf_temp = 0.0;
      if (RefreezeEnergy >= 0.0) {
        RefrozenWater = RefreezeEnergy / (Lf * RHO_W) * delta_t;
        if (RefrozenWater > snow->surf_water) {
          RefrozenWater = snow->surf_water;
          RefreezeEnergy = RefrozenWater * Lf * RHO_W / (delta_t);
        }
        melt_energy += RefreezeEnergy;
        SurfaceSwq += RefrozenWater;
        Ice += RefrozenWater;
        snow->surf_water -= RefrozenWater;
        if (snow->surf_water < 0.0) snow->surf_water = 0.0;
        SnowMelt = 0.0;
      }
      else {
        SnowMelt = fabs(RefreezeEnergy) / (Lf * RHO_W) * delta_t;
        melt_energy += RefreezeEnergy;
      }
      if (snow->surf_water < -(snow->vapor_flux)) {
        snow->blowing_flux *= -(snow->surf_water / snow->vapor_flux);
        snow->vapor_flux = -(snow->surf_water);
        snow->surface_flux = -(snow->surf_water) - snow->blowing_flux;
        snow->surf_water = 0.0;
      }
      else snow->surf_water += snow->vapor_flux;
      if (SnowMelt < Ice) {
        if (SnowMelt <= PackSwq) {
          snow->surf_water += SnowMelt;
          PackSwq -= SnowMelt;
          Ice -= SnowMelt;
        }
        else {
          snow->surf_water += SnowMelt + snow->pack_water;
          snow->pack_water = 0.0;
          PackSwq = 0.0;
          Ice -= SnowMelt;
          SurfaceSwq = Ice;
        }
      }
      else {
        SnowMelt = Ice;
        snow->surf_water += Ice;
        SurfaceSwq = 0.0;
        snow->surf_temp = 0.0;
        PackSwq = 0.0;
        snow->pack_temp = 0.0;
        Ice = 0.0;
        melt_energy -= RefreezeEnergy;
        RefreezeEnergy = RefreezeEnergy / fabs(RefreezeEnergy) * SnowMelt * Lf * RHO_W / (delta_t);
        melt_energy += RefreezeEnergy;
      }
    }
    else {
      if (SurfaceSwq > MIN_SWQ_EB_THRES) {
        snow->surf_temp = root_brent((double)(snow->surf_temp - SNOW_DT), (double)(snow->surf_temp + SNOW_DT),
                                     SnowPackEnergyBalance, &x);
        if (snow->surf_temp <= -998) {
          /* options.TFALLBACK TRUE (initialize_global.c:194) */
          snow->surf_temp = *OldTSurf;
          snow->surf_temp_fbflag = 1;
          snow->surf_temp_fbcount++;
        }
      }
      else snow->surf_temp = 999;
      if (snow->surf_temp > -998 && snow->surf_temp < 999) {
        Qnet = SnowPackEnergyBalance(snow->surf_temp, &x);
        SnowMelt = 0.0;
        SurfaceSwq += snow->surf_water;
        Ice += snow->surf_water;
        snow->surf_water = 0.0;
        melt_energy += snow->surf_water * Lf * RHO_W / (delta_t);
        if (SurfaceSwq < -(snow->vapor_flux)) {
          snow->blowing_flux *= -(SurfaceSwq / snow->vapor_flux);
          snow->vapor_flux = -SurfaceSwq;
          snow->surface_flux = -SurfaceSwq - snow->blowing_flux;
          SurfaceSwq = 0.0;
          Ice = PackSwq;
        }
        else {
          SurfaceSwq += snow->vapor_flux;
          Ice += snow->vapor_flux;
        }
      }
    }
  }
  else snow->surf_temp = 999;

  MaxLiquidWater = LIQUID_WATER_CAPACITY * SurfaceSwq;
  if (snow->surf_water > MaxLiquidWater) {
    melt[0] = snow->surf_water - MaxLiquidWater;
    snow->surf_water = MaxLiquidWater;
  }
  else melt[0] = 0.0;
  snow->pack_water += melt[0];
  PackRefreezeEnergy = snow->pack_water * Lf * RHO_W;
  if (PackCC < -PackRefreezeEnergy) {
    PackSwq += snow->pack_water;
    Ice += snow->pack_water;
    snow->pack_water = 0.0;
    if (PackSwq > 0.0) {
      PackCC = PackSwq * CH_ICE * snow->pack_temp + PackRefreezeEnergy;
      snow->pack_temp = PackCC / (CH_ICE * PackSwq);
      if (snow->pack_temp > 0.) snow->pack_temp = 0.;
    }
    else snow->pack_temp = 0.0;
  }
  else {
    snow->pack_temp = 0.0;
    DeltaPackSwq = -PackCC / (Lf * RHO_W);
    snow->pack_water -= DeltaPackSwq;
    PackSwq += DeltaPackSwq;
    Ice += DeltaPackSwq;
  }
  MaxLiquidWater = LIQUID_WATER_CAPACITY * PackSwq;
  if (snow->pack_water > MaxLiquidWater) {
    melt[0] = snow->pack_water - MaxLiquidWater;
    snow->pack_water = MaxLiquidWater;
  }
  else melt[0] = 0.0;
  Ice = PackSwq + SurfaceSwq;
  if (Ice > MAX_SURFACE_SWE) {
    SurfaceCC = CH_ICE * snow->surf_temp * SurfaceSwq;
    PackCC = CH_ICE * snow->pack_temp * PackSwq;
    if (SurfaceSwq > MAX_SURFACE_SWE) {
      PackCC += SurfaceCC * (SurfaceSwq - MAX_SURFACE_SWE) / SurfaceSwq;
      SurfaceCC -= SurfaceCC * (SurfaceSwq - MAX_SURFACE_SWE) / SurfaceSwq;
      PackSwq += SurfaceSwq - MAX_SURFACE_SWE;
      SurfaceSwq -= SurfaceSwq - MAX_SURFACE_SWE;
    }
    else if (SurfaceSwq < MAX_SURFACE_SWE) {
      PackCC -= PackCC * (MAX_SURFACE_SWE - SurfaceSwq) / PackSwq;
      SurfaceCC += PackCC * (MAX_SURFACE_SWE - SurfaceSwq) / PackSwq;
      PackSwq -= MAX_SURFACE_SWE - SurfaceSwq;
      SurfaceSwq += MAX_SURFACE_SWE - SurfaceSwq;
    }
    snow->pack_temp = PackCC / (CH_ICE * PackSwq);
    snow->surf_temp = SurfaceCC / (CH_ICE * SurfaceSwq);
  }
  else {
    PackSwq = 0.0;
    PackCC = 0.0;
    snow->pack_temp = 0.0;
  }
  snow->swq = Ice + snow->pack_water + snow->surf_water;
  if (snow->swq == 0.0) {
    snow->surf_temp = 0.0;
    snow->pack_temp = 0.0;
  }
  MassBalanceError = (InitialSwq - snow->swq) + (RainFall + SnowFall) - melt[0] + snow->vapor_flux;
  melt[0] *= 1000.;
  snow->mass_error = MassBalanceError;
  snow->coldcontent = SurfaceCC;
  snow->vapor_flux *= -1.;
  *save_advection = advection;
  *save_deltaCC = deltaCC;
  *save_grnd_flux = grnd_flux;
  *save_latent = latent_heat;
  *save_latent_sub = latent_heat_sub;
  *save_sensible = sensible_heat;
  *save_advected_sensible = advected_sensible_heat;
  *save_refreeze_energy = RefreezeEnergy;
  *save_Qnet = Qnet;
  (void)melt_energy;
  return 0;
}

/* massrelease.c:35-93 (recursive in the reference; same recursion here) */
static void MassRelease(double *InterceptedSnow, double *TempInterceptionStorage, double *ReleasedMass,
                        double *Drip)
{
  double TempDrip, TempReleasedMass, Threshold, MaxRelease;
  if (*InterceptedSnow > MIN_INTERCEPTION_STORAGE) {
    Threshold = 0.10 * *InterceptedSnow;
    MaxRelease = 0.17 * *InterceptedSnow;
    if ((*TempInterceptionStorage) >= Threshold) {
      *Drip += Threshold;
      *InterceptedSnow -= Threshold;
      *TempInterceptionStorage -= Threshold;
      if (*InterceptedSnow < MIN_INTERCEPTION_STORAGE) TempReleasedMass = 0.0;
      else
        TempReleasedMass = ((*InterceptedSnow - MIN_INTERCEPTION_STORAGE) < MaxRelease)
                               ? (*InterceptedSnow - MIN_INTERCEPTION_STORAGE) : MaxRelease;
      *ReleasedMass += TempReleasedMass;
      *InterceptedSnow -= TempReleasedMass;
      MassRelease(InterceptedSnow, TempInterceptionStorage, ReleasedMass, Drip);
    }
    else {
      TempDrip = (*TempInterceptionStorage < *InterceptedSnow) ? *TempInterceptionStorage : *InterceptedSnow;
      *Drip += TempDrip;
      *InterceptedSnow -= TempDrip;
    }
  }
  else {
    TempDrip = (*TempInterceptionStorage < *InterceptedSnow) ? *TempInterceptionStorage : *InterceptedSnow;
    *Drip += TempDrip;
    *InterceptedSnow -= TempDrip;
    *TempInterceptionStorage = 0.0;
  }
}

/* ======================================================================================
 * evaporation
 * ====================================================================================== */

/* canopy_evap.c:235-529 transpiration (Nfrost == 1, ice == 0, CARBON FALSE, RC_JARVIS) */
static void transpiration(olayer *layer, ovegvar *veg_var, const oveglib *vl, double rad, double vpd,
                          double net_short, double air_temp, double ra, double dryFrac, double delta_t,
                          double elevation, const double *Wcr, const double *Wpwp, double *layerevap,
                          const float *root)
{
  int i;
  double gsm_inv, moist1, moist2, evap, Wcr1, root_sum, spare_evap, avail_moist[MAXL], gc;
  moist1 = 0.0;
  Wcr1 = 0.0;
  for (i = 0; i < NLAYER - 1; i++) {
    if (root[i] > 0.) {
      avail_moist[i] = 0;
      avail_moist[i] += ((layer[i].moist - 0.) * 1.);
      moist1 += avail_moist[i];
      Wcr1 += Wcr[i];
    }
    else avail_moist[i] = 0.;
  }
  i = NLAYER - 1;
  moist2 = 0;
  moist2 += ((layer[i].moist - 0.) * 1.);
  avail_moist[i] = moist2;
  /* options.SHARE_LAYER_MOIST TRUE (initialize_global.c:187) */
  if ((moist1 >= Wcr1 && moist2 >= Wcr[NLAYER - 1] && Wcr1 > 0.) ||
      (moist1 >= Wcr1 && (1 - root[NLAYER - 1]) >= 0.5) ||
      (moist2 >= Wcr[NLAYER - 1] && root[NLAYER - 1] >= 0.5)) {
    gsm_inv = 1.0;
    veg_var->rc = calc_rc(vl->rmin, net_short, (float)vl->RGL, air_temp, vpd, veg_var->LAI, gsm_inv);
    evap = penman(air_temp, elevation, rad, vpd, ra, veg_var->rc, vl->rarc) * delta_t / SEC_PER_DAY * dryFrac;
    root_sum = 1.0;
    spare_evap = 0.0;
    for (i = 0; i < NLAYER; i++) {
      if (avail_moist[i] >= Wcr[i]) layerevap[i] = evap * (double)root[i];
      else {
        if (avail_moist[i] >= Wpwp[i]) gsm_inv = (avail_moist[i] - Wpwp[i]) / (Wcr[i] - Wpwp[i]);
        else gsm_inv = 0.0;
        layerevap[i] = evap * gsm_inv * (double)root[i];
        root_sum -= root[i];
        spare_evap = evap * (double)root[i] * (1.0 - gsm_inv);
      }
    }
    if (spare_evap > 0.0) {
      for (i = 0; i < NLAYER; i++)
        if (avail_moist[i] >= Wcr[i]) layerevap[i] += (double)root[i] * spare_evap / root_sum;
    }
  }
  else {
    gc = 0;
    for (i = 0; i < NLAYER; i++) {
      if (avail_moist[i] >= Wcr[i]) gsm_inv = 1.0;
      else if (avail_moist[i] >= Wpwp[i] && avail_moist[i] < Wcr[i])
        gsm_inv = (avail_moist[i] - Wpwp[i]) / (Wcr[i] - Wpwp[i]);
      else gsm_inv = 0.0;
      if (gsm_inv > 0.0) {
        veg_var->rc = calc_rc(vl->rmin, net_short, (float)vl->RGL, air_temp, vpd, veg_var->LAI, gsm_inv);
        layerevap[i] = penman(air_temp, elevation, rad, vpd, ra, veg_var->rc, vl->rarc) * delta_t /
                       SEC_PER_DAY * dryFrac * (double)root[i];
        if (veg_var->rc > 0) gc += 1 / (veg_var->rc);
        else gc += HUGE_RESIST;
      }
      else {
        layerevap[i] = 0.0;
        gc += 0;
      }
    }
    if (gc > 0) veg_var->rc = 1 / gc;
    else veg_var->rc = HUGE_RESIST;
    if (veg_var->rc > RSMAX) veg_var->rc = RSMAX;
  }
  for (i = 0; i < NLAYER; i++) {
    if (layerevap[i] > layer[i].moist - Wpwp[i]) layerevap[i] = layer[i].moist - Wpwp[i];
    if (layerevap[i] < 0.0) layerevap[i] = 0.0;
  }
}

/* canopy_evap.c:46-233 */
static double canopy_evap(olayer *layer, ovegvar *veg_var, int CALC_EVAP, const oveglib *vl, double *Wdew,
                          double delta_t, double rad, double vpd, double net_short, double air_temp,
                          double ra, double elevation, double ppt, const double *Wcr, const double *Wpwp,
                          const float *root, double *dryFrac)
{
  int i;
  double Evap, tmp_Evap, canopyevap, tmp_Wdew, f, throughfall, rc, layerevap[MAXL];
  Evap = 0;
  for (i = 0; i < NLAYER; i++) layerevap[i] = 0;
  canopyevap = 0;
  throughfall = 0;
  tmp_Wdew = *Wdew;
  veg_var->Wdew = tmp_Wdew;
  if (tmp_Wdew > veg_var->Wdmax) {
    throughfall = tmp_Wdew - veg_var->Wdmax;
    tmp_Wdew = veg_var->Wdmax;
  }
  rc = calc_rc((double)0.0, net_short, (float)vl->RGL, air_temp, vpd, veg_var->LAI, (double)1.0);
  if (veg_var->LAI > 0)
    canopyevap = 240 * pow((tmp_Wdew / veg_var->Wdmax), (2.0 / 3.0)) *
                 penman(air_temp, elevation, rad, vpd, ra, rc, vl->rarc) * delta_t / SEC_PER_DAY;
  else canopyevap = 0;
  if (canopyevap > 0.0 && delta_t == SEC_PER_DAY)
    f = (1.0 < ((tmp_Wdew + ppt) / canopyevap)) ? 1.0 : ((tmp_Wdew + ppt) / canopyevap);
  else if (canopyevap > 0.0)
    f = (1.0 < ((tmp_Wdew) / canopyevap)) ? 1.0 : ((tmp_Wdew) / canopyevap);
  else f = 1.0;
  canopyevap *= f;
  if (veg_var->Wdmax > 0) *dryFrac = 1.0 - f * pow((tmp_Wdew / veg_var->Wdmax), (2.0 / 3.0));
  else *dryFrac = 0;
  tmp_Wdew += ppt - canopyevap;
  if (tmp_Wdew < 0.0) tmp_Wdew = 0.0;
  if (tmp_Wdew <= veg_var->Wdmax) throughfall += 0.0;
  else {
    throughfall += tmp_Wdew - veg_var->Wdmax;
    tmp_Wdew = veg_var->Wdmax;
  }
  if (CALC_EVAP)
    transpiration(layer, veg_var, vl, rad, vpd, net_short, air_temp, ra, *dryFrac, delta_t, elevation,
                  Wcr, Wpwp, layerevap, root);
  veg_var->canopyevap = canopyevap;
  veg_var->throughfall = throughfall;
  veg_var->Wdew = tmp_Wdew;
  tmp_Evap = canopyevap;
  for (i = 0; i < NLAYER; i++) {
    layer[i].evap = layerevap[i];
    tmp_Evap += layerevap[i];
  }
  Evap += tmp_Evap / (1000. * delta_t);
  return Evap;
}

/* arno_evap.c:62-203 (with VIC-Res's 240x factors, :100 and :176) */
static double arno_evap(olayer *layer, double rad, double air_temp, double vpd, double depth1,
                        double max_moist, double elevation, double b_infilt, double ra, double delta_t,
                        double moist_resid)
{
  int num_term, i;
  double tmp, beta_asp, dummy, ratio, as, Epot, moist, evap, max_infil, Evap, tmpsum;
  Evap = 0;
  moist = 0;
  moist += (layer[0].moist - 0.) * 1.;
  if (moist > max_moist) moist = max_moist;
  Epot = 240 * penman(air_temp, elevation, rad, vpd, ra, 0.0, RARC_SOIL) * delta_t / SEC_PER_DAY;
  max_infil = (1.0 + b_infilt) * max_moist;
  if (b_infilt == -1.0) tmp = max_infil;
  else {
    ratio = 1.0 - (moist) / (max_moist);
    if (ratio > 1.0) return VERROR;
    else {
      if (ratio < 0.0) return VERROR;
      else ratio = pow(ratio, (1.0 / (b_infilt + 1.0)));
    }
    tmp = max_infil * (1.0 - ratio);
  }
  if (tmp >= max_infil) evap = Epot;
  else {
    ratio = tmp / max_infil;
    ratio = 1.0 - ratio;
    if (ratio > 1.0) return VERROR;
    else {
      if (ratio < 0.0) return VERROR;
      else if (ratio != 0.0) ratio = pow(ratio, b_infilt);
    }
    as = 1 - ratio;
    ratio = pow(ratio, (1.0 / b_infilt));
    dummy = 1.0;
    for (num_term = 1; num_term <= 30; num_term++) {
      tmpsum = ratio;
      for (i = 1; i < num_term; i++) tmpsum *= ratio;
      dummy += b_infilt * tmpsum / (b_infilt + num_term);
    }
    beta_asp = as + (1.0 - as) * (1.0 - ratio) * dummy;
    evap = 240 * Epot * beta_asp;
  }
  if (evap > 0.0) {
    if (moist > moist_resid * depth1 * 1000.) {
      if (evap > moist - moist_resid * depth1 * 1000.) evap = moist - moist_resid * depth1 * 1000.;
    }
    else evap = 0.0;
  }
  layer[0].evap = evap;
  Evap += evap / 1000. / delta_t;
  return Evap;
}

/* ======================================================================================
 * canopy snow interception
 * ====================================================================================== */

/* func_canopy_energy_bal.c:29-290: varargs -> context (AERO_RESIST_CANSNOW == AR_406_FULL) */
typedef struct {
  double delta_t, elevation;
  const double *Wcr, *Wpwp;
  double AirDens, EactAir, Press, Le, Tcanopy, Vpd;
  double *dryFrac, *Evap, *Ra, *Ra_used;
  double Rainfall;
  const oveglib *vl;
  const float *root;
  double IntRain, IntSnow;
  double *Wdew;
  olayer *layer;
  ovegvar *veg_var;
  double LongOverIn, LongUnderOut, NetShortOver;
  double *AdvectedEnergy, *LatentHeat, *LatentHeatSub, *LongOverOut, *NetLongOver, *NetRadiation,
         *RefreezeEnergy, *SensibleHeat, *VaporMassFlux;
} fce_ctx;

static double func_canopy_energy_bal(double Tfoliage, void *vctx)
{
  fce_ctx *x = (fce_ctx *)vctx;
  double EsSnow, Ls, RestTerm, Tmp, prec;
  Tmp = Tfoliage + KELVIN;
  *x->LongOverOut = STEFAN_B * (Tmp * Tmp * Tmp * Tmp);
  *x->NetRadiation = x->NetShortOver + x->LongOverIn + x->LongUnderOut - 2 * (*x->LongOverOut);
  *x->NetLongOver = x->LongOverIn - (*x->LongOverOut);
  if (x->IntSnow > 0) {
    x->Ra_used[0] = x->Ra[0];
    x->Ra_used[1] = x->Ra[1];
    x->Ra_used[1] *= 10.;
    EsSnow = svp(Tfoliage);
    *x->VaporMassFlux = x->AirDens * (EPS / x->Press) * (x->EactAir - EsSnow) / x->Ra_used[1] / RHO_W;
    if (x->Vpd == 0.0 && *x->VaporMassFlux < 0.0) *x->VaporMassFlux = 0.0;
    Ls = (677. - 0.07 * Tfoliage) * JOULESPCAL * GRAMSPKG;
    *x->LatentHeatSub = Ls * *x->VaporMassFlux * RHO_W;
    *x->LatentHeat = 0;
    *x->Evap = 0;
    x->veg_var->throughfall = 0;
  }
  else {
    x->Ra_used[0] = x->Ra[0];
    x->Ra_used[1] = x->Ra[1];
    *x->Wdew = x->IntRain * 1000.;
    prec = x->Rainfall * 1000;
    *x->Evap = canopy_evap(x->layer, x->veg_var, 0, x->vl, x->Wdew, x->delta_t, *x->NetRadiation, x->Vpd,
                           x->NetShortOver, x->Tcanopy, x->Ra_used[1], x->elevation, prec, x->Wcr,
                           x->Wpwp, x->root, x->dryFrac);
    *x->Wdew /= 1000.;
    *x->LatentHeat = x->Le * *x->Evap * RHO_W;
    *x->LatentHeatSub = 0;
  }
  *x->SensibleHeat = x->AirDens * Cp * (x->Tcanopy - Tfoliage) / x->Ra_used[1];
  *x->AdvectedEnergy = (CH_WATER * x->Tcanopy * x->Rainfall) / (x->delta_t);
  RestTerm = *x->SensibleHeat + *x->LatentHeat + *x->LatentHeatSub + *x->NetRadiation + *x->AdvectedEnergy;
  if (x->IntSnow > 0) {
    *x->RefreezeEnergy = (x->IntRain * Lf * RHO_W) / (x->delta_t);
    if (Tfoliage == 0.0 && RestTerm > -(*x->RefreezeEnergy)) {
      *x->RefreezeEnergy = -RestTerm;
      RestTerm = 0.0;
    }
    else RestTerm += *x->RefreezeEnergy;
  }
  else *x->RefreezeEnergy = 0;
  return RestTerm;
}

/* snow_intercept.c:8-560 */
static int snow_intercept(double Dt, double F, double LAI, double Le, double LongOverIn, double LongUnderOut,
                          double MaxInt, double ShortOverIn, double ShortUnderIn, double Tcanopy,
                          double bare_albedo, double *AdvectedEnergy, double *AlbedoOver, double *IntRain,
                          double *IntSnow, double *LatentHeat, double *LatentHeatSub, double *LongOverOut,
                          double *MeltEnergy, double *NetLongOver, double *NetShortOver, double *Ra,
                          double *Ra_used, double *RainFall, double *SensibleHeat, double *SnowFall,
                          double *Tfoliage, int *Tfoliage_fbflag, int *Tfoliage_fbcount,
                          double *TempIntStorage, double *VaporMassFlux, double *Wind, const float *root,
                          const oveglib *vl, double *dryFrac, const oatmos *atmos, olayer *layer,
                          const osoil *soil_con, ovegvar *veg_var)
{
  double AirDens, EactAir, Press, Vpd, BlownSnow, DeltaSnowInt, Drip, ExcessSnowMelt, InitialSnowInt,
         InitialWaterInt, IntRainFract, IntSnowFract, MaxWaterInt, MaxSnowInt, NetRadiation, Overload,
         PotSnowMelt, RainThroughFall, RefreezeEnergy, ReleasedMass, SnowThroughFall, Imax1, IntRainOrg,
         Tupper, Tlower, Evap, OldTfoliage, Qnet, MassBalanceError;
  fce_ctx x;
  (void)ShortUnderIn;
  AirDens = atmos->density;
  EactAir = atmos->vp;
  Press = atmos->pressure;
  Vpd = atmos->vpd;
  *Tfoliage_fbflag = 0;
  *RainFall /= 1000.;
  *SnowFall /= 1000.;
  *IntRain /= 1000.;
  MaxInt /= 1000.;
  IntRainOrg = *IntRain;
  InitialWaterInt = *IntSnow + *IntRain;
  *IntSnow /= F;
  *IntRain /= F;
  InitialSnowInt = *IntSnow;
  Drip = 0.0;
  ReleasedMass = 0.0;
  OldTfoliage = *Tfoliage;
  Imax1 = 4.0 * LAI_SNOW_MULTIPLIER * LAI;
  if ((*Tfoliage) < -1.0 && (*Tfoliage) > -3.0) MaxSnowInt = ((*Tfoliage) * 3.0 / 2.0) + (11.0 / 2.0);
  else if ((*Tfoliage) > -1.0) MaxSnowInt = 4.0;
  else MaxSnowInt = 1.0;
  MaxSnowInt *= LAI_SNOW_MULTIPLIER * LAI;
  if (MaxSnowInt > 0) {
    DeltaSnowInt = (1 - *IntSnow / MaxSnowInt) * *SnowFall;
    if (DeltaSnowInt + *IntSnow > MaxSnowInt) DeltaSnowInt = MaxSnowInt - *IntSnow;
    if (DeltaSnowInt < 0.0) DeltaSnowInt = 0.0;
  }
  else DeltaSnowInt = 0.0;
  if ((*Tfoliage) < -3.0 && DeltaSnowInt > 0.0 && Wind[1] > 1.0) {
    BlownSnow = (0.2 * Wind[1] - 0.2) * DeltaSnowInt;
    if (BlownSnow >= DeltaSnowInt) BlownSnow = DeltaSnowInt;
    DeltaSnowInt -= BlownSnow;
  }
  if (*IntSnow + DeltaSnowInt > Imax1) DeltaSnowInt = 0.0;
  SnowThroughFall = (*SnowFall - DeltaSnowInt) * F + (*SnowFall) * (1 - F);
  if (*SnowFall == 0 && *IntSnow < MIN_SWQ_EB_THRES) {
    SnowThroughFall += *IntSnow;
    DeltaSnowInt -= *IntSnow;
  }
  *IntSnow += DeltaSnowInt;
  if (*IntSnow < SMALL) *IntSnow = 0.0;
  MaxWaterInt = LIQUID_WATER_CAPACITY * (*IntSnow) + MaxInt;
  if ((*IntRain + *RainFall) <= MaxWaterInt) {
    *IntRain += *RainFall;
    RainThroughFall = *RainFall * (1 - F);
  }
  else {
    RainThroughFall = (*IntRain + *RainFall - MaxWaterInt) * F + (*RainFall * (1 - F));
    *IntRain = MaxWaterInt;
  }
  if (*RainFall == 0 && *IntRain < MIN_SWQ_EB_THRES) {
    RainThroughFall += *IntRain;
    *IntRain = 0.0;
  }
  if (*IntRain + *IntSnow > Imax1) {
    Overload = (*IntSnow + *IntRain) - Imax1;
    IntRainFract = *IntRain / (*IntRain + *IntSnow);
    IntSnowFract = *IntSnow / (*IntRain + *IntSnow);
    *IntRain = *IntRain - Overload * IntRainFract;
    *IntSnow = *IntSnow - Overload * IntSnowFract;
    RainThroughFall = RainThroughFall + (Overload * IntRainFract) * F;
    SnowThroughFall = SnowThroughFall + (Overload * IntSnowFract) * F;
  }
  if (*IntRain + *IntSnow < SMALL) *Tfoliage = Tcanopy;

  x.delta_t = Dt; x.elevation = soil_con->elevation; x.Wcr = soil_con->Wcr; x.Wpwp = soil_con->Wpwp;
  x.AirDens = AirDens; x.EactAir = EactAir; x.Press = Press; x.Le = Le; x.Tcanopy = Tcanopy; x.Vpd = Vpd;
  x.dryFrac = dryFrac; x.Evap = &Evap; x.Ra = Ra; x.Ra_used = Ra_used; x.vl = vl; x.root = root;
  x.IntRain = IntRainOrg; x.Wdew = IntRain; x.layer = layer; x.veg_var = veg_var;
  x.LongOverIn = LongOverIn; x.LongUnderOut = LongUnderOut;
  x.AdvectedEnergy = AdvectedEnergy; x.LatentHeat = LatentHeat; x.LatentHeatSub = LatentHeatSub;
  x.LongOverOut = LongOverOut; x.NetLongOver = NetLongOver; x.NetRadiation = &NetRadiation;
  x.RefreezeEnergy = &RefreezeEnergy; x.SensibleHeat = SensibleHeat; x.VaporMassFlux = VaporMassFlux;

  Tupper = Tlower = MISSING;
  if (*IntSnow > 0 || *SnowFall > 0) {
    *AlbedoOver = NEW_SNOW_ALB;
    *NetShortOver = (1. - *AlbedoOver) * ShortOverIn;
    x.Rainfall = *RainFall; x.IntSnow = *IntSnow; x.NetShortOver = *NetShortOver;
    Qnet = func_canopy_energy_bal(0., &x);
    if (Qnet != 0) {
      Tupper = 0;
      if ((*Tfoliage) <= 0.) Tlower = (*Tfoliage) - SNOW_DT;
      else Tlower = -SNOW_DT;
    }
    else *Tfoliage = 0.;
  }
  else {
    *AlbedoOver = bare_albedo;
    *NetShortOver = (1. - *AlbedoOver) * ShortOverIn;
    Qnet = -9999;
    Tupper = (*Tfoliage) + SNOW_DT;
    Tlower = (*Tfoliage) - SNOW_DT;
  }
  if (Tupper != MISSING && Tlower != MISSING) {
    x.Rainfall = *RainFall; x.IntSnow = *IntSnow; x.NetShortOver = *NetShortOver;
    *Tfoliage = root_brent(Tlower, Tupper, func_canopy_energy_bal, &x);
    if (*Tfoliage <= -998) {
      *Tfoliage = OldTfoliage;      /* TFALLBACK */
      *Tfoliage_fbflag = 1;
      (*Tfoliage_fbcount)++;
    }
    Qnet = func_canopy_energy_bal(*Tfoliage, &x);
  }
  if (*IntSnow <= 0) RainThroughFall = veg_var->throughfall / 1000.;
  RefreezeEnergy *= Dt;
  MaxWaterInt = LIQUID_WATER_CAPACITY * (*IntSnow) + MaxInt;
  *VaporMassFlux *= Dt;
  if (*Tfoliage == 0) {
    if (-(*VaporMassFlux) > *IntRain) {
      *VaporMassFlux = -(*IntRain);
      *IntRain = 0.;
    }
    else *IntRain += *VaporMassFlux;
    if (RefreezeEnergy < 0) {
      PotSnowMelt = ((-RefreezeEnergy / Lf / RHO_W) < *IntSnow) ? (-RefreezeEnergy / Lf / RHO_W) : *IntSnow;
      *MeltEnergy -= (Lf * PotSnowMelt * RHO_W) / (Dt);
    }
    else {
      PotSnowMelt = 0;
      *MeltEnergy -= (Lf * PotSnowMelt * RHO_W) / (Dt);
    }
    if ((*IntRain + PotSnowMelt) <= MaxWaterInt) {
      *IntSnow -= PotSnowMelt;
      *IntRain += PotSnowMelt;
      PotSnowMelt = 0.0;
    }
    else {
      ExcessSnowMelt = PotSnowMelt + *IntRain - MaxWaterInt;
      *IntSnow -= MaxWaterInt - (*IntRain);
      *IntRain = MaxWaterInt;
      if (*IntSnow < 0.0) *IntSnow = 0.0;
      if (SnowThroughFall > 0.0 && InitialSnowInt <= MIN_INTERCEPTION_STORAGE) {
        Drip += ExcessSnowMelt;
        *IntSnow -= ExcessSnowMelt;
        if (*IntSnow < 0.0) *IntSnow = 0.0;
      }
      else *TempIntStorage += ExcessSnowMelt;
      MassRelease(IntSnow, TempIntStorage, &ReleasedMass, &Drip);
    }
    MaxWaterInt = LIQUID_WATER_CAPACITY * (*IntSnow) + MaxInt;
    if (*IntRain > MaxWaterInt) {
      Drip += *IntRain - MaxWaterInt;
      *IntRain = MaxWaterInt;
    }
  }
  else {
    *TempIntStorage = 0.0;
    if (-RefreezeEnergy > -(*IntRain) * Lf) {
      *IntSnow += fabs(RefreezeEnergy) / Lf;
      *IntRain -= fabs(RefreezeEnergy) / Lf;
      *MeltEnergy += (fabs(RefreezeEnergy) * RHO_W) / (Dt);
      RefreezeEnergy = 0.0;
    }
    else {
      *IntSnow += *IntRain;
      *MeltEnergy += (Lf * *IntRain * RHO_W) / (Dt);
      *IntRain = 0.0;
    }
    if (-(*VaporMassFlux) > *IntSnow) {
      *VaporMassFlux = -(*IntSnow);
      *IntSnow = 0.0;
    }
    else *IntSnow += *VaporMassFlux;
  }
  *IntSnow *= F;
  *IntRain *= F;
  *MeltEnergy *= F;
  *VaporMassFlux *= F;
  Drip *= F;
  ReleasedMass *= F;
  if (*IntSnow == 0 && *IntRain > MaxInt) {
    RainThroughFall += *IntRain - MaxInt;
    *IntRain = MaxInt;
  }
  MassBalanceError = (InitialWaterInt - (*IntSnow + *IntRain)) + (*SnowFall + *RainFall) -
                     (SnowThroughFall + RainThroughFall + Drip + ReleasedMass) + *VaporMassFlux;
  (void)MassBalanceError; (void)Qnet;
  *RainFall = RainThroughFall + Drip;
  *SnowFall = SnowThroughFall + ReleasedMass;
  *VaporMassFlux *= -1.;
  *RainFall *= 1000.;
  *SnowFall *= 1000.;
  *IntRain *= 1000.;
  *MeltEnergy = RefreezeEnergy / Dt;
  return 0;
}

/* ======================================================================================
 * solve_snow.c:7-577
 * ====================================================================================== */
static double solve_snow(int overstory, double BareAlbedo, double LongUnderOut, double MIN_RAIN_TEMP,
                         double MAX_SNOW_TEMP, double Tcanopy, double Tgrnd, double air_temp, double prec,
                         double snow_grnd_flux, double *AlbedoUnder, double *Le, double *LongUnderIn,
                         double *NetLongSnow, double *NetShortGrnd, double *NetShortSnow,
                         double *ShortUnderIn, double *Torg_snow, double *aero_resist,
                         double *aero_resist_used, double *coverage, double *delta_coverage,
                         double *displacement, double *melt_energy, double *out_prec, double *out_rain,
                         double *out_snow, double *ppt, double *rainfall, double *ref_height,
                         double *roughness, double *snow_inflow, double *snowfall, double *surf_atten,
                         double *wind, const float *root, int INCLUDE_SNOW, int Nveg, int iveg, int dt,
                         const oveglib *vl, int *UnderStory, double *dryFrac, const oatmos *atmos,
                         oenergy *energy, olayer *layer, osnow *snow, const osoil *soil_con,
                         ovegvar *veg_var)
{
  double ShortOverIn, melt, old_coverage, old_depth, old_swq, rainonly, tmp_grnd_flux, store_snowfall,
         density, longwave, pressure, shortwave, vp, vpd;
  int day_in_year = atmos->day_in_year;
  density = atmos->density;
  longwave = atmos->longwave;
  pressure = atmos->pressure;
  shortwave = atmos->shortwave;
  vp = atmos->vp;
  vpd = atmos->vpd;
  melt = 0.;
  *ppt = 0.;
  (*melt_energy) = 0.;
  rainonly = calc_rainonly(air_temp, prec, MAX_SNOW_TEMP, MIN_RAIN_TEMP);
  *snowfall = 1. * (prec - rainonly);          /* gauge_correction == 1 (CORRPREC FALSE) */
  *rainfall = 1. * rainonly;
  if (*snowfall < 1e-5) *snowfall = 0.;
  (*out_prec) = *snowfall + *rainfall;
  (*out_rain) = *rainfall;
  (*out_snow) = *snowfall;
  store_snowfall = *snowfall;
  (*Le) = (2.501e6 - 0.002361e6 * air_temp);
  if (*UnderStory == 999) {
    if (snow->swq > 0 || *snowfall > 0) *UnderStory = 2;
    else *UnderStory = 0;
  }
  (*ShortUnderIn) = shortwave;
  (*LongUnderIn) = longwave;
  if (snow->swq > 0 || *snowfall > 0. || (snow->snow_canopy > 0. && overstory)) {
    snow->snow = 1;
    if (!overstory) (*surf_atten) = 1.;
    old_coverage = snow->coverage;
    if (iveg != Nveg) {
      if (overstory) {
        (*ShortUnderIn) *= (*surf_atten);
        ShortOverIn = (1. - (*surf_atten)) * shortwave;
        ShortOverIn /= veg_var->vegcover;
        snow_intercept((double)dt * SECPHOUR, 1., veg_var->LAI, (*Le), longwave, LongUnderOut,
                       veg_var->Wdmax, ShortOverIn, *ShortUnderIn, Tcanopy, BareAlbedo,
                       &energy->canopy_advection, &energy->AlbedoOver, &veg_var->Wdew, &snow->snow_canopy,
                       &energy->canopy_latent, &energy->canopy_latent_sub, LongUnderIn,
                       &energy->canopy_refreeze, &energy->NetLongOver, &energy->NetShortOver, aero_resist,
                       aero_resist_used, rainfall, &energy->canopy_sensible, snowfall, &energy->Tfoliage,
                       &energy->Tfoliage_fbflag, &energy->Tfoliage_fbcount, &snow->tmp_int_storage,
                       &snow->canopy_vapor_flux, wind, root, vl, dryFrac, atmos, layer, soil_con, veg_var);
        veg_var->throughfall = *rainfall + *snowfall;
        energy->LongOverIn = longwave;
      }
      else if (*snowfall > 0. && veg_var->Wdew > 0.) {
        *rainfall += veg_var->Wdew;
        veg_var->throughfall = *rainfall + *snowfall;
        veg_var->Wdew = 0.;
        energy->NetLongOver = 0;
        energy->LongOverIn = 0;
        energy->Tfoliage = air_temp;
        energy->Tfoliage_fbflag = 0;
      }
      else {
        veg_var->throughfall = *rainfall + *snowfall;
        energy->NetLongOver = 0;
        energy->LongOverIn = 0;
        energy->Tfoliage = air_temp;
        energy->Tfoliage_fbflag = 0;
      }
      veg_var->throughfall = (1 - veg_var->vegcover) * (*out_prec) + veg_var->vegcover * veg_var->throughfall;
      *rainfall = (1 - veg_var->vegcover) * (*out_rain) + veg_var->vegcover * (*rainfall);
      *snowfall = (1 - veg_var->vegcover) * (*out_snow) + veg_var->vegcover * (*snowfall);
      snow->canopy_vapor_flux *= veg_var->vegcover;
      snow->snow_canopy *= veg_var->vegcover;
      veg_var->Wdew *= veg_var->vegcover;
      veg_var->canopyevap *= veg_var->vegcover;
      energy->canopy_advection *= veg_var->vegcover;
      energy->canopy_latent *= veg_var->vegcover;
      energy->canopy_latent_sub *= veg_var->vegcover;
      energy->canopy_sensible *= veg_var->vegcover;
      energy->canopy_refreeze *= veg_var->vegcover;
      energy->NetShortOver *= veg_var->vegcover;
      energy->NetLongOver *= veg_var->vegcover;
    }
    else {
      energy->NetLongOver = 0;
      energy->LongOverIn = 0;
    }
    if (snow->swq > 0.0 || *snowfall > 0) {
      (*NetShortGrnd) = 0.;
      (*snow_inflow) += *rainfall + *snowfall;
      old_swq = snow->swq;
      (*UnderStory) = 2;
      if (snow->swq > 0 && store_snowfall == 0) {
        snow->last_snow++;
        snow->albedo = snow_albedo(*snowfall, snow->swq, snow->depth, snow->albedo, snow->coldcontent,
                                   (double)dt, snow->last_snow, snow->MELTING);
        (*AlbedoUnder) = (*coverage * snow->albedo + (1. - *coverage) * BareAlbedo);
      }
      else {
        snow->last_snow = 0;
        snow->albedo = NEW_SNOW_ALB;
        (*AlbedoUnder) = snow->albedo;
      }
      (*NetShortSnow) = (1.0 - *AlbedoUnder) * (*ShortUnderIn);
      snow_melt((*Le), (*NetShortSnow), Tcanopy, Tgrnd, roughness, aero_resist[*UnderStory],
                aero_resist_used, air_temp, *coverage, (double)dt * SECPHOUR, density, snow_grnd_flux,
                *LongUnderIn, pressure, *rainfall, *snowfall, vp, vpd, wind[*UnderStory],
                ref_height[*UnderStory], NetLongSnow, Torg_snow, &melt, &energy->error,
                &energy->advected_sensible, &energy->advection, &energy->deltaCC, &tmp_grnd_flux,
                &energy->latent, &energy->latent_sub, &energy->refreeze_energy, &energy->sensible,
                INCLUDE_SNOW, snow);
      *ppt += melt;
      energy->AlbedoUnder = *AlbedoUnder;
      if (snow->swq > 0.) {
        if (snow->surf_temp <= 0)
          snow->density = snow_density(snow, *snowfall, old_swq, Tgrnd, air_temp, (double)dt);
        else if (snow->last_snow == 0)
          snow->density = new_snow_density(air_temp);
        old_depth = snow->depth;
        snow->depth = 1000. * snow->swq / snow->density;
        if (snow->coldcontent >= 0 &&
            ((soil_con->lat >= 0 && (day_in_year > 60 && day_in_year < 273)) ||
             (soil_con->lat < 0 && (day_in_year < 60 || day_in_year > 273))))
          snow->MELTING = 1;
        else if (snow->MELTING && *snowfall > TraceSnow)
          snow->MELTING = 0;
        if (snow->swq > 0) snow->coverage = 1.;
        else snow->coverage = 0.;
        (void)old_depth;
      }
      else snow->coverage = 0.;
      *delta_coverage = old_coverage - snow->coverage;
      if (*delta_coverage != 0) {
        if (old_coverage > snow->coverage) {
          *coverage = (old_coverage);
          (*AlbedoUnder) = (*coverage - snow->coverage) / (1. - snow->coverage) * snow->albedo;
          (*AlbedoUnder) += (1. - *coverage) / (1. - snow->coverage) * BareAlbedo;
          (*melt_energy) = (*delta_coverage) * (energy->advection - energy->deltaCC + energy->latent +
                                                energy->latent_sub + energy->sensible +
                                                energy->refreeze_energy + energy->advected_sensible);
        }
        else if (old_coverage < snow->coverage) {
          *coverage = snow->coverage;
          *delta_coverage = 0;
        }
        else {
          *coverage = snow->coverage;
          *delta_coverage = 0.;
        }
      }
      else if (old_coverage == 0 && snow->coverage == 0) {
        *delta_coverage = 1.;
        *coverage = 0.;
        (*melt_energy) = (energy->advection - energy->deltaCC + energy->latent + energy->latent_sub +
                          energy->sensible + energy->refreeze_energy + energy->advected_sensible);
      }
      (*NetLongSnow) *= (snow->coverage);
      (*NetShortSnow) *= (snow->coverage);
      (*NetShortGrnd) *= (snow->coverage);
      energy->latent *= (snow->coverage + *delta_coverage);
      energy->latent_sub *= (snow->coverage + *delta_coverage);
      energy->sensible *= (snow->coverage + *delta_coverage);
      if (snow->swq == 0) {
        snow->density = 0.;
        snow->depth = 0.;
        snow->surf_water = 0;
        snow->pack_water = 0;
        snow->surf_temp = 0;
        snow->pack_temp = 0;
        snow->coverage = 0;
        snow->MELTING = 0;
      }
      *snowfall = 0;
      *rainfall = 0;
    }
    else {
      *ppt += *rainfall;
      energy->AlbedoOver = 0.;
      (*AlbedoUnder) = BareAlbedo;
      (*NetLongSnow) = 0.;
      (*NetShortSnow) = 0.;
      (*NetShortGrnd) = 0.;
      (*delta_coverage) = 0.;
      energy->latent = 0.;
      energy->latent_sub = 0.;
      energy->sensible = 0.;
      snow->last_snow = (int)MISSING;
      snow->MELTING = 0;
    }
  }
  else {
    *UnderStory = 0;
    snow->snow = 0;
    energy->Tfoliage = air_temp;
    energy->AlbedoOver = 0.;
    (*AlbedoUnder) = BareAlbedo;
    energy->NetLongOver = 0.;
    energy->LongOverIn = 0.;
    energy->NetShortOver = 0.;
    energy->ShortOverIn = 0.;
    energy->latent = 0.;
    energy->latent_sub = 0.;
    energy->sensible = 0.;
    (*NetLongSnow) = 0.;
    (*NetShortSnow) = 0.;
    (*NetShortGrnd) = 0.;
    (*delta_coverage) = 0.;
    energy->Tfoliage = Tcanopy;
    snow->MELTING = 0;
    snow->last_snow = (int)MISSING;
    snow->albedo = NEW_SNOW_ALB;
  }
  energy->melt_energy *= -1.;
  return melt;
}

/* ======================================================================================
 * surface energy balance, water-balance branch
 * ====================================================================================== */

/* func_surf_energy_bal.c:8-875, QUICK_FLUX TRUE, GRND_FLUX_TYPE GF_410, FROZEN_SOIL FALSE.
 * The ~90 varargs become explicit parameters; `fusion` is never written on this path. */
static double func_surf_energy_bal(double Ts, int VEG, const oveglib *vl, double delta_t, double Cs1,
                                   double Cs2, double D1, double D2, double T1_old, double T2,
                                   double Ts_old, double dp, double kappa1, double kappa2,
                                   double max_moist, const float *root, int UnderStory, int overstory,
                                   double NetShortBare, double NetShortGrnd, double NetShortSnow,
                                   double Tair, double atmos_density, double atmos_pressure,
                                   double emissivity, double LongBareIn, double LongSnowIn,
                                   double surf_atten, double vp, double vpd, double *dryFrac, double *Wdew,
                                   double *displacement, double *ra, double *Ra_used, double rainfall,
                                   double *ref_height, double *roughness, double *wind, double Le,
                                   double Advection, double OldTSurf, double Tsnow_surf,
                                   double kappa_snow, double melt_energy, double snow_coverage,
                                   double snow_density_, double snow_swq, double snow_water,
                                   double *deltaCC, double *refreeze_energy, double *vapor_flux,
                                   double *blowing_flux, double *surface_flux, const osoil *soil_con,
                                   olayer *layer, ovegvar *veg_var, int INCLUDE_SNOW, int SNOWING,
                                   double *NetLongBare, double *NetLongSnow, double *T1, double *deltaH,
                                   double *fusion, double *grnd_flux, double *latent_heat,
                                   double *latent_heat_sub, double *sensible_heat, double *snow_flux,
                                   double *store_error)
{
  int i;
  double Evap, LongBareOut, NetBareRad, TMean, Tmp, error, VaporMassFlux, BlowingMassFlux, SurfaceMassFlux,
         temp_latent_heat, temp_latent_heat_sub, ga_veg, ga_bare, ga_average, tmp_height, transp[MAXL];
  double Ra_bare[3], Ra_veg[3], tmp_wind[3], tmp_displacement[3], tmp_roughness[3], tmp_ref_height[3];
  double elevation = soil_con->elevation, b_infilt = soil_con->b_infilt;
  const double *depth = soil_con->depth, *resid_moist = soil_con->resid_moist;
  (void)Ts_old; (void)Ra_veg;
  TMean = Ts;
  Tmp = TMean + KELVIN;
  for (i = 0; i < NLAYER; i++) transp[i] = 0;
  if (snow_coverage > 0 && !INCLUDE_SNOW) *snow_flux = (kappa_snow * (Tsnow_surf - TMean));
  else if (INCLUDE_SNOW) {
    *snow_flux = 0;
    Tsnow_surf = TMean;
  }
  else *snow_flux = 0;
  *T1 = estimate_T1(TMean, T1_old, T2, D1, D2, kappa1, kappa2, Cs1, Cs2, dp, delta_t);
  *grnd_flux = (snow_coverage + (1. - snow_coverage) * surf_atten) *
               (kappa1 / D1 * ((*T1) - TMean) + (kappa2 / D2 * (1. - exp(-D1 / dp)) * (T2 - (*T1)))) / 2.;
  *deltaH = (Cs1 * ((Ts_old + T1_old) - (TMean + *T1)) * D1 / delta_t / 2.);
  if (INCLUDE_SNOW) {
    if (TMean > 0) *deltaCC = CH_ICE * (snow_swq - snow_water) * (0 - OldTSurf) / delta_t;
    else *deltaCC = CH_ICE * (snow_swq - snow_water) * (TMean - OldTSurf) / delta_t;
    *refreeze_energy = (snow_water * Lf * snow_density_) / delta_t;
    *deltaCC *= snow_coverage;
    *refreeze_energy *= snow_coverage;
  }
  LongBareOut = STEFAN_B * Tmp * Tmp * Tmp * Tmp;
  if (INCLUDE_SNOW) (*NetLongSnow) = (LongSnowIn - snow_coverage * LongBareOut);
  (*NetLongBare) = (LongBareIn - (1. - snow_coverage) * LongBareOut);
  NetBareRad = (NetShortBare + (*NetLongBare) + *grnd_flux + *deltaH + *fusion);
  if (wind[UnderStory] > 0.0 && overstory && SNOWING)
    Ra_used[0] = ra[UnderStory] /
                 StabilityCorrection(ref_height[UnderStory], 0.f, TMean, Tair, wind[UnderStory], roughness[UnderStory]);
  else if (wind[UnderStory] > 0.0)
    Ra_used[0] = ra[UnderStory] / StabilityCorrection(ref_height[UnderStory], displacement[UnderStory], TMean,
                                                      Tair, wind[UnderStory], roughness[UnderStory]);
  else Ra_used[0] = HUGE_RESIST;
  if (veg_var->vegcover < 1) {
    Ra_veg[0] = Ra_used[0];
    if (Ra_veg[0] > 0) {
      ga_veg = 1 / Ra_veg[0];
      tmp_wind[0] = wind[0];
      tmp_wind[1] = VERROR;
      tmp_wind[2] = VERROR;
      tmp_height = soil_con->rough / RATIO_RL_HEIGHT_VEG;
      tmp_displacement[0] = RATIO_DH_HEIGHT_VEG * tmp_height;   /* calc_veg_displacement */
      tmp_roughness[0] = soil_con->rough;
      tmp_ref_height[0] = WINDH_SOIL;
      CalcAerodynamic(0, 0, 0, soil_con->snow_rough, soil_con->rough, 0, Ra_bare, tmp_wind, tmp_displacement,
                      tmp_ref_height, tmp_roughness);
      Ra_bare[0] /= StabilityCorrection(tmp_ref_height[0], tmp_displacement[0], TMean, Tair, tmp_wind[0],
                                        tmp_roughness[0]);
      if (Ra_bare[0] > 0) {
        ga_bare = 1 / Ra_bare[0];
        ga_average = veg_var->vegcover * ga_veg + (1 - veg_var->vegcover) * ga_bare;
        Ra_used[0] = 1 / ga_average;
      }
      else Ra_used[0] = 0;
    }
    else Ra_used[0] = 0;
  }
  if (VEG && !SNOWING && veg_var->vegcover > 0) {
    Evap = canopy_evap(layer, veg_var, 1, vl, Wdew, delta_t, NetBareRad, vpd, NetShortBare, Tair, Ra_used[1],
                       elevation, rainfall, soil_con->Wcr, soil_con->Wpwp, root, dryFrac);
    if (veg_var->vegcover < 1) {
      for (i = 0; i < NLAYER; i++) {
        transp[i] = layer[i].evap;
        layer[i].evap = 0;
      }
      Evap *= veg_var->vegcover;
      Evap += (1 - veg_var->vegcover) *
              arno_evap(layer, surf_atten * NetBareRad, Tair, vpd, depth[0], max_moist * depth[0] * 1000.,
                        elevation, b_infilt, Ra_used[0], delta_t, resid_moist[0]);
      for (i = 0; i < NLAYER; i++) {
        layer[i].evap = veg_var->vegcover * transp[i] + (1 - veg_var->vegcover) * layer[i].evap;
        if (layer[i].evap > 0)
          layer[i].bare_evap_frac = 1 - (veg_var->vegcover * transp[i]) / layer[i].evap;
        else layer[i].bare_evap_frac = 0;
      }
      veg_var->throughfall = (1 - veg_var->vegcover) * rainfall + veg_var->vegcover * veg_var->throughfall;
      veg_var->canopyevap *= veg_var->vegcover;
      veg_var->Wdew *= veg_var->vegcover;
    }
    else {
      for (i = 0; i < NLAYER; i++) layer[i].bare_evap_frac = 0;
    }
  }
  else if (!SNOWING) {
    Evap = arno_evap(layer, NetBareRad, Tair, vpd, depth[0], max_moist * depth[0] * 1000., elevation, b_infilt,
                     Ra_used[0], delta_t, resid_moist[0]);
    for (i = 0; i < NLAYER; i++) layer[i].bare_evap_frac = 1;
  }
  else Evap = 0.;
  *latent_heat = -RHO_W * Le * Evap;
  *latent_heat_sub = 0.;
  if (INCLUDE_SNOW) {
    VaporMassFlux = *vapor_flux * ice_density / delta_t;
    BlowingMassFlux = *blowing_flux * ice_density / delta_t;
    SurfaceMassFlux = *surface_flux * ice_density / delta_t;
    latent_heat_from_snow(atmos_density, vp, Le, atmos_pressure, Ra_used[0], TMean, vpd, &temp_latent_heat,
                          &temp_latent_heat_sub, &VaporMassFlux, &BlowingMassFlux, &SurfaceMassFlux);
    *latent_heat += temp_latent_heat * snow_coverage;
    *latent_heat_sub = temp_latent_heat_sub * snow_coverage;
    *vapor_flux = VaporMassFlux * delta_t / ice_density;
    *blowing_flux = BlowingMassFlux * delta_t / ice_density;
    *surface_flux = SurfaceMassFlux * delta_t / ice_density;
  }
  else *latent_heat *= (1. - snow_coverage);
  if (snow_coverage < 1 || INCLUDE_SNOW) {
    *sensible_heat = atmos_density * Cp * (Tair - (TMean)) / Ra_used[0];
    if (!INCLUDE_SNOW) (*sensible_heat) *= (1. - snow_coverage);
  }
  else *sensible_heat = 0.;
  error = (NetBareRad + NetShortGrnd + NetShortSnow + emissivity * (*NetLongSnow)) + *sensible_heat +
          (*latent_heat + *latent_heat_sub) + *snow_flux * snow_coverage + melt_energy + Advection - *deltaCC;
  if (INCLUDE_SNOW) {
    if (Tsnow_surf == 0.0 && error > -(*refreeze_energy)) {
      *refreeze_energy = -error;
      error = 0.0;
    }
    else error += *refreeze_energy;
  }
  *store_error = error;
  return error;
}

/* calc_surf_energy_bal.c:5-744, FULL_ENERGY FALSE branch (Tsurf = Tair, :537-542) */
static double calc_surf_energy_bal(double Le, double LongUnderIn, double NetLongSnow, double NetShortGrnd,
                                   double NetShortSnow, double OldTSurf, double ShortUnderIn,
                                   double SnowAlbedo, double SnowLatent, double SnowLatentSub,
                                   double SnowSensible, double Tair, double VPDcanopy, double VPcanopy,
                                   double delta_coverage, double dp, double melt_energy,
                                   double snow_coverage, double snow_depth, double BareAlbedo,
                                   double surf_atten, double *aero_resist, double *aero_resist_used,
                                   double *displacement, double *melt, double *ppt, double rainfall,
                                   double *ref_height, double *roughness, double *wind, const float *root,
                                   int INCLUDE_SNOW, int UnderStory, int Nveg, int dt, int iveg,
                                   int overstory, const oveglib *vl, double *dryFrac, const oatmos *atmos,
                                   oenergy *energy, olayer *layer, osnow *snow, const osoil *soil_con,
                                   ovegvar *veg_var)
{
  int VEG;
  double Cs1, Cs2, D1, D2, LongBareIn, NetLongBare, NetShortBare, T1, T1_old, T2, Ts_old, Tsnow_surf, Tsurf,
         delta_t, error, kappa1, kappa2, kappa_snow, max_moist, refrozen_water, Wdew, TmpNetLongSnow,
         TmpNetShortSnow, LongSnowIn, Tnew_node[3], atmos_density, atmos_pressure, emissivity;
  if (iveg != Nveg) {
    if (veg_var->vegcover > 0.0) VEG = 1;
    else VEG = 0;
  }
  else VEG = 0;
  T2 = soil_con->avg_temp;
  Ts_old = energy->T[0];
  T1_old = energy->T[1];
  D1 = soil_con->depth[0];
  D2 = soil_con->depth[0];
  kappa1 = energy->kappa[0];
  kappa2 = energy->kappa[1];
  Cs1 = energy->Cs[0];
  Cs2 = energy->Cs[1];
  atmos_density = atmos->density;
  atmos_pressure = atmos->pressure;
  emissivity = 1.;
  delta_t = (double)dt * 3600.;
  max_moist = soil_con->max_moist[0] / (soil_con->depth[0] * 1000.);
  Tsnow_surf = snow->surf_temp;
  Wdew = veg_var->Wdew;
  if (snow->depth > 0.) kappa_snow = K_SNOW * (snow->density) * (snow->density) / snow_depth;
  else kappa_snow = 0;
  NetShortBare = (ShortUnderIn * (1. - (snow_coverage + delta_coverage)) * (1. - BareAlbedo) +
                  ShortUnderIn * (delta_coverage) * (1. - SnowAlbedo));
  LongBareIn = (1. - snow_coverage) * LongUnderIn;
  if (INCLUDE_SNOW || snow->swq == 0) {
    TmpNetLongSnow = NetLongSnow;
    TmpNetShortSnow = NetShortSnow;
    LongSnowIn = snow_coverage * LongUnderIn;
  }
  else {
    TmpNetShortSnow = 0.;
    TmpNetLongSnow = 0.;
    LongSnowIn = 0.;
  }
  Tsurf = Tair;
  error = func_surf_energy_bal(Tsurf, VEG, vl, delta_t, Cs1, Cs2, D1, D2, T1_old, T2, Ts_old, dp, kappa1,
                               kappa2, max_moist, root, UnderStory, overstory, NetShortBare, NetShortGrnd,
                               TmpNetShortSnow, Tair, atmos_density, atmos_pressure, emissivity, LongBareIn,
                               LongSnowIn, surf_atten, VPcanopy, VPDcanopy, dryFrac, &Wdew, displacement,
                               aero_resist, aero_resist_used, rainfall, ref_height, roughness, wind, Le,
                               energy->advection, OldTSurf, Tsnow_surf, kappa_snow, melt_energy,
                               snow_coverage, snow->density, snow->swq, snow->surf_water, &energy->deltaCC,
                               &energy->refreeze_energy, &snow->vapor_flux, &snow->blowing_flux,
                               &snow->surface_flux, soil_con, layer, veg_var, INCLUDE_SNOW, snow->snow,
                               &NetLongBare, &TmpNetLongSnow, &T1, &energy->deltaH, &energy->fusion,
                               &energy->grnd_flux, &energy->latent, &energy->latent_sub, &energy->sensible,
                               &energy->snow_flux, &energy->error);
  energy->error = error;
  Tnew_node[0] = Tsurf;
  Tnew_node[1] = T1;
  Tnew_node[2] = soil_con->avg_temp + (T1 - soil_con->avg_temp) * exp(-(dp - D1) / dp);  /* Zsum_node[2] == dp */
  /* calc_layer_average_thermal_props (frozen_soil.c:8-42) */
  energy->T[0] = Tnew_node[0]; energy->T[1] = Tnew_node[1]; energy->T[2] = Tnew_node[2];
  estimate_layer_ice_content_quick_flux(layer, soil_con, energy->T[0], energy->T[1]);
  if (!snow->snow && !INCLUDE_SNOW) {
    if (iveg != Nveg) *ppt = veg_var->throughfall;
    else *ppt = rainfall;
  }
  energy->NetShortGrnd = NetShortGrnd;
  if (INCLUDE_SNOW) {
    energy->NetLongUnder = NetLongBare + TmpNetLongSnow;
    energy->NetShortUnder = NetShortBare + TmpNetShortSnow + NetShortGrnd;
  }
  else {
    energy->NetLongUnder = NetLongBare + NetLongSnow;
    energy->NetShortUnder = NetShortBare + NetShortSnow + NetShortGrnd;
    energy->latent = (SnowLatent + energy->latent);
    energy->latent_sub = (SnowLatentSub + energy->latent_sub);
    energy->sensible = (SnowSensible + energy->sensible);
  }
  energy->LongUnderOut = LongUnderIn - energy->NetLongUnder;
  energy->AlbedoUnder = ((1. - (snow_coverage + delta_coverage)) * BareAlbedo +
                         (snow_coverage + delta_coverage) * SnowAlbedo);
  energy->melt_energy = melt_energy;
  energy->Tsurf = (snow->coverage * snow->surf_temp + (1. - snow->coverage) * Tsurf);
  if (INCLUDE_SNOW) {
    if (-(snow->vapor_flux) > snow->swq) {
      snow->blowing_flux *= -(snow->swq / snow->vapor_flux);
      snow->vapor_flux = -(snow->swq);
      snow->surface_flux = snow->vapor_flux - snow->blowing_flux;
    }
    snow->swq += snow->vapor_flux;
    snow->surf_water += snow->vapor_flux;
    snow->surf_water = (snow->surf_water < 0) ? 0. : snow->surf_water;
    if (energy->refreeze_energy >= 0.0) {
      refrozen_water = energy->refreeze_energy / (Lf * RHO_W) * delta_t;
      if (refrozen_water > snow->surf_water) {
        refrozen_water = snow->surf_water;
        energy->refreeze_energy = refrozen_water * Lf * RHO_W / delta_t;
      }
      snow->surf_water -= refrozen_water;
      if (snow->surf_water < 0.0) snow->surf_water = 0.0;
      (*melt) = 0.0;
    }
    else {
      (*melt) = fabs(energy->refreeze_energy) / (Lf * RHO_W) * delta_t;
      snow->swq -= *melt;
      if (snow->swq < 0) {
        (*melt) += snow->swq;
        snow->swq = 0;
      }
    }
    if (snow->swq > 0) {
      snow->surf_temp = (Tsurf > 0) ? 0 : Tsurf;
      snow->coldcontent = CH_ICE * snow->surf_temp * snow->swq;
      snow->depth = 1000. * snow->swq / snow->density;
      if (snow->swq > 0) snow->coverage = 1.;
      else snow->coverage = 0.;
      if (snow->surf_temp > 0) energy->snow_flux = (energy->grnd_flux + energy->deltaH + energy->fusion);
    }
    else {
      snow->density = 0.;
      snow->depth = 0.;
      snow->surf_water = 0;
      snow->pack_water = 0;
      snow->surf_temp = 0;
      snow->pack_temp = 0;
      snow->coverage = 0;
    }
    snow->vapor_flux *= -1;
  }
  return Tsurf;
}

/* ======================================================================================
 * runoff.c:7-576 (Nfrost == 1, ice == 0)
 * ====================================================================================== */
/* compute_zwt.c:7-45 */
static double compute_zwt(const osoil *soil_con, int lindex, double moist)
{
  int i;
  double zwt = -99999.;
  i = VR_ZWT_NPOINTS - 1;
  while (i >= 1 && moist > soil_con->zwtvmoist_moist[lindex][i]) i--;
  if (i == VR_ZWT_NPOINTS - 1) {
    if (moist < soil_con->zwtvmoist_moist[lindex][i]) zwt = 999;
    else if (moist == soil_con->zwtvmoist_moist[lindex][i]) zwt = soil_con->zwtvmoist_zwt[lindex][i];
  }
  else
    zwt = soil_con->zwtvmoist_zwt[lindex][i + 1] +
          (soil_con->zwtvmoist_zwt[lindex][i] - soil_con->zwtvmoist_zwt[lindex][i + 1]) *
              (moist - soil_con->zwtvmoist_moist[lindex][i + 1]) /
              (soil_con->zwtvmoist_moist[lindex][i] - soil_con->zwtvmoist_moist[lindex][i + 1]);
  return zwt;
}

/* compute_zwt.c:48-113 */
static void wrap_compute_zwt(const osoil *soil_con, ocell *cell)
{
  int lindex;
  double total_depth = 0, tmp_depth, tmp_moist;
  for (lindex = 0; lindex < NLAYER; lindex++) total_depth += soil_con->depth[lindex];
  for (lindex = 0; lindex < NLAYER; lindex++) cell->layer_zwt[lindex] = compute_zwt(soil_con, lindex, cell->layer[lindex].moist);
  if (cell->layer_zwt[NLAYER - 1] == 999) cell->layer_zwt[NLAYER - 1] = -total_depth * 100;
  lindex = NLAYER - 1;
  tmp_depth = total_depth;
  while (lindex >= 0 && soil_con->max_moist[lindex] - cell->layer[lindex].moist <= SMALL) {
    tmp_depth -= soil_con->depth[lindex];
    lindex--;
  }
  if (lindex < 0) cell->zwt = 0;
  else if (lindex < NLAYER - 1) {
    if (cell->layer_zwt[lindex] != 999) cell->zwt = cell->layer_zwt[lindex];
    else cell->zwt = -tmp_depth * 100;
  }
  else cell->zwt = cell->layer_zwt[lindex];
  tmp_moist = 0;
  for (lindex = 0; lindex < NLAYER; lindex++) tmp_moist += cell->layer[lindex].moist;
  cell->zwt_lumped = compute_zwt(soil_con, NLAYER + 1, tmp_moist);
  if (cell->zwt_lumped == 999) cell->zwt_lumped = -total_depth * 100;
}

/* compute_pot_evap.c:8-80.  lib: the model's library incl. the 4 reference classes at [nclass .. nclass+3]; the
 * bare tile's class IS entry nclass.  net_short is read by calc_rc() before it is assigned for cover i (the
 * reference's order); it starts as 0 here, and only NATVEG (after TALL) can depend on it. */
static void compute_pot_evap(const oveglib *lib, int nclass, int veg_class, int mon, int dt, double shortwave,
                             double net_longwave, double tair, double vpd, double elevation,
                             double aero_resist[6][2], double *pot_evap)
{
  static const char ref_crop[6] = {0, 0, 1, 1, 0, 0};
  int i;
  double albedo, net_short = 0., net_rad, rs, rarc, RGL, lai, rc, ra;
  for (i = 0; i < 6; i++) {
    if (i < 4) {
      rs = lib[nclass + i].rmin; rarc = lib[nclass + i].rarc; RGL = lib[nclass + i].RGL;
      lai = lib[nclass + i].LAI[mon]; albedo = lib[nclass + i].albedo[mon];
    }
    else {
      rs = lib[veg_class].rmin;
      if (i == 5) rs = 0;
      rarc = lib[veg_class].rarc; RGL = lib[veg_class].RGL; lai = lib[veg_class].LAI[mon]; albedo = lib[veg_class].albedo[mon];
    }
    if (rs == 0) rc = 0;
    else if (lai == 0) rc = RSMAX;
    else if (ref_crop[i]) { rc = rs / (lai * 0.5); rc = (rc > RSMAX) ? RSMAX : rc; }
    else rc = calc_rc(rs, net_short, (float)RGL, tair, vpd, lai, 1.0);
    if (i < 4 || !lib[veg_class].overstory) ra = aero_resist[i][0];
    else ra = aero_resist[i][1];
    net_short = (1.0 - albedo) * shortwave;
    net_rad = net_short + net_longwave;
    pot_evap[i] = penman(tair, elevation, net_rad, vpd, ra, rc, rarc) * dt / 24.0;
  }
}

static void compute_runoff_and_asat(const osoil *soil_con, const double *moist, double inflow, double *A,
                                    double *runoff)
{
  double top_moist, top_max_moist, ex, max_infil, i_0, basis;
  int lindex;
  top_moist = 0.;
  top_max_moist = 0.;
  for (lindex = 0; lindex < NLAYER - 1; lindex++) {
    top_moist += moist[lindex];
    top_max_moist += soil_con->max_moist[lindex];
  }
  if (top_moist > top_max_moist) top_moist = top_max_moist;
  ex = soil_con->b_infilt / (1.0 + soil_con->b_infilt);
  *A = 1.0 - pow((1.0 - top_moist / top_max_moist), ex);
  max_infil = (1.0 + soil_con->b_infilt) * top_max_moist;
  i_0 = max_infil * (1.0 - pow((1.0 - *A), (1.0 / soil_con->b_infilt)));
  if (inflow == 0.0) *runoff = 0.0;
  else if (max_infil == 0.0) *runoff = inflow;
  else if ((i_0 + inflow) > max_infil) *runoff = inflow - top_max_moist + top_moist;
  else {
    basis = 1.0 - (i_0 + inflow) / max_infil;
    *runoff = (inflow - top_max_moist + top_moist + top_max_moist * pow(basis, 1.0 * (1.0 + soil_con->b_infilt)));
  }
  if (*runoff < 0.) *runoff = 0.;
}

static int runoff(ocell *cell, const osoil *soil_con, double ppt, int dt)
{
  int lindex, time_step, tmplayer;
  double A, frac, tmp_runoff, inflow, resid_moist[MAXL], org_moist[MAXL], avail_liq, liq[MAXL], ice[MAXL],
         max_moist[MAXL], Ksat[MAXL], Q12[MAXL - 1], tmp_moist_for_runoff[MAXL], runoff_, tmp_dt_runoff,
         baseflow, evap[MAXL], dt_inflow, dt_runoff, dt_baseflow, rel_moist, Dsmax, tmp_inflow, tmp_moist,
         tmp_liq, sum_liq, evap_fraction, evap_sum;
  olayer *layer = cell->layer;
  for (lindex = 0; lindex < NLAYER; lindex++)
    resid_moist[lindex] = soil_con->resid_moist[lindex] * soil_con->depth[lindex] * 1000.;
  cell->runoff = 0;
  cell->baseflow = 0;
  cell->asat = 0;
  baseflow = 0;
  for (lindex = 0; lindex < NLAYER; lindex++) {
    evap[lindex] = layer[lindex].evap / (double)dt;
    org_moist[lindex] = layer[lindex].moist;
    layer[lindex].moist = 0;
    if (evap[lindex] > 0) {
      sum_liq = 0;
      avail_liq = (org_moist[lindex] - 0. - resid_moist[lindex]);
      if (avail_liq < 0) avail_liq = 0;
      sum_liq += avail_liq * 1.;
      if (sum_liq > 0) evap_fraction = evap[lindex] / sum_liq;
      else evap_fraction = 1.0;
      evap_sum = evap[lindex];
      evap[lindex] = avail_liq * evap_fraction;
      avail_liq -= evap[lindex];
      evap_sum -= evap[lindex] * 1.;
      (void)evap_sum;
    }
  }
  inflow = ppt;
  for (lindex = 0; lindex < NLAYER; lindex++) {
    Ksat[lindex] = soil_con->Ksat[lindex] / 24.;
    liq[lindex] = org_moist[lindex] - 0.;
    ice[lindex] = 0.;
    max_moist[lindex] = soil_con->max_moist[lindex];
  }
  for (lindex = 0; lindex < NLAYER; lindex++) tmp_moist_for_runoff[lindex] = (liq[lindex] + ice[lindex]);
  compute_runoff_and_asat(soil_con, tmp_moist_for_runoff, inflow, &A, &runoff_);
  tmp_dt_runoff = runoff_ / (double)dt;
  dt_inflow = inflow / (double)dt;
  for (time_step = 0; time_step < dt; time_step++) {
    inflow = dt_inflow;
    for (lindex = 0; lindex < NLAYER - 1; lindex++) {
      if ((tmp_liq = liq[lindex] - evap[lindex]) < resid_moist[lindex]) tmp_liq = resid_moist[lindex];
      if (liq[lindex] > resid_moist[lindex])
        Q12[lindex] = Ksat[lindex] * pow(((tmp_liq - resid_moist[lindex]) /
                                          (soil_con->max_moist[lindex] - resid_moist[lindex])),
                                         soil_con->expt[lindex]);
      else Q12[lindex] = 0.;
    }
    for (lindex = 0; lindex < NLAYER - 1; lindex++) {
      if (lindex == 0) dt_runoff = tmp_dt_runoff;
      else dt_runoff = 0;
      tmp_inflow = 0.;
      liq[lindex] = liq[lindex] + (inflow - dt_runoff) - (Q12[lindex] + evap[lindex]);
      if ((liq[lindex] + ice[lindex]) > max_moist[lindex]) {
        tmp_inflow = (liq[lindex] + ice[lindex]) - max_moist[lindex];
        liq[lindex] = max_moist[lindex] - ice[lindex];
        if (lindex == 0) {
          Q12[lindex] += tmp_inflow;
          tmp_inflow = 0;
        }
        else {
          tmplayer = lindex;
          while (tmp_inflow > 0) {
            tmplayer--;
            if (tmplayer < 0) {
              runoff_ += tmp_inflow;
              tmp_inflow = 0;
            }
            else {
              liq[tmplayer] += tmp_inflow;
              if ((liq[tmplayer] + ice[tmplayer]) > max_moist[tmplayer]) {
                tmp_inflow = ((liq[tmplayer] + ice[tmplayer]) - max_moist[tmplayer]);
                liq[tmplayer] = max_moist[tmplayer] - ice[tmplayer];
              }
              else tmp_inflow = 0;
            }
          }
        }
      }
      if (liq[lindex] < 0) {
        Q12[lindex] += liq[lindex];
        liq[lindex] = 0;
      }
      if ((liq[lindex] + ice[lindex]) < resid_moist[lindex]) {
        Q12[lindex] += (liq[lindex] + ice[lindex]) - resid_moist[lindex];
        liq[lindex] = resid_moist[lindex] - ice[lindex];
      }
      inflow = (Q12[lindex] + tmp_inflow);
      Q12[lindex] += tmp_inflow;
    }
    lindex = NLAYER - 1;
    Dsmax = soil_con->Dsmax / 24.;
    rel_moist = (liq[lindex] - resid_moist[lindex]) / (soil_con->max_moist[lindex] - resid_moist[lindex]);
    frac = Dsmax * soil_con->Ds / soil_con->Ws;
    dt_baseflow = frac * rel_moist;
    if (rel_moist > soil_con->Ws) {
      frac = (rel_moist - soil_con->Ws) / (1 - soil_con->Ws);
      dt_baseflow += Dsmax * (1 - soil_con->Ds / soil_con->Ws) * pow(frac, soil_con->c);
    }
    if (dt_baseflow < 0) dt_baseflow = 0;
    liq[lindex] += Q12[lindex - 1] - (evap[lindex] + dt_baseflow);
    tmp_moist = 0;
    if ((liq[lindex] + ice[lindex]) < resid_moist[lindex]) {
      dt_baseflow += (liq[lindex] + ice[lindex]) - resid_moist[lindex];
      liq[lindex] = resid_moist[lindex] - ice[lindex];
    }
    if ((liq[lindex] + ice[lindex]) > max_moist[lindex]) {
      tmp_moist = ((liq[lindex] + ice[lindex]) - max_moist[lindex]);
      liq[lindex] = max_moist[lindex] - ice[lindex];
      tmplayer = lindex;
      while (tmp_moist > 0) {
        tmplayer--;
        if (tmplayer < 0) {
          runoff_ += tmp_moist;
          tmp_moist = 0;
        }
        else {
          liq[tmplayer] += tmp_moist;
          if ((liq[tmplayer] + ice[tmplayer]) > max_moist[tmplayer]) {
            tmp_moist = ((liq[tmplayer] + ice[tmplayer]) - max_moist[tmplayer]);
            liq[tmplayer] = max_moist[tmplayer] - ice[tmplayer];
          }
          else tmp_moist = 0;
        }
      }
    }
    baseflow += dt_baseflow;
  }
  lindex = NLAYER - 1;
  if (baseflow < 0) {
    layer[lindex].evap += baseflow;
    baseflow = 0;
  }
  for (lindex = 0; lindex < NLAYER; lindex++) tmp_moist_for_runoff[lindex] = (liq[lindex] + ice[lindex]);
  compute_runoff_and_asat(soil_con, tmp_moist_for_runoff, 0, &A, &tmp_runoff);
  for (lindex = 0; lindex < NLAYER; lindex++) layer[lindex].moist += ((liq[lindex] + ice[lindex]) * 1.);
  cell->asat += A * 1.;
  cell->runoff += runoff_ * 1.;
  cell->baseflow += baseflow * 1.;
  if (soil_con->have_zwt) wrap_compute_zwt(soil_con, cell);      /* runoff.c:503 */
  return 0;
}

/* ======================================================================================
 * surface_fluxes.c:11-1168: the sub-step loop (NF = TIME_STEP / SNOW_STEP passes when the tile has snow or the
 * record has snowfall, else one pass with the record's own forcing, :385-396) with
 * CLOSE_ENERGY FALSE (MAX_ITER_GRND_CANOPY == 0: each inner do-while body runs exactly once)
 * ====================================================================================== */
static int surface_fluxes(const vr_options *opt, int overstory, double BareAlbedo, double surf_atten,
                          double *aero_resist /* [3] of the current veg */, double *displacement,
                          double *ref_height, double *roughness, double *wind, const float *root, int Nveg,
                          int band, double dp, int iveg, const oveglib *vl, const oatmos *atm /* [NF + 1] or [1] */,
                          int NF, int NR, oenergy *energy, ocell *cell, osnow *snow, const osoil *soil_con,
                          ovegvar *veg_var, double *out_prec, double *out_rain, double *out_snow,
                          const oveglib *lib, int nclass, int veg_class, double pet_aero[7][3] /* NULL: PET not wanted */)
{
  int INCLUDE_SNOW = 0, UNSTABLE_SNOW = 0, UnderStory, lidx, N_steps = 1, step_dt = opt->dt_hours, hidx, endhidx;
  double store_ppt = 0, store_melt = 0, store_vapor_flux = 0, store_surface_flux = 0, store_blowing_flux = 0,
         store_pot_evap[6];
  double store_AlbedoOver = 0, store_AlbedoUnder = 0, store_AtmosLatent = 0, store_AtmosLatentSub = 0,
         store_AtmosSensible = 0, store_LongOverIn = 0, store_LongUnderIn = 0, store_LongUnderOut = 0,
         store_NetLongAtmos = 0, store_NetLongOver = 0, store_NetLongUnder = 0, store_NetShortAtmos = 0,
         store_NetShortGrnd = 0, store_NetShortOver = 0, store_NetShortUnder = 0, store_ShortOverIn = 0,
         store_ShortUnderIn = 0, store_canopy_advection = 0, store_canopy_latent = 0, store_canopy_latent_sub = 0,
         store_canopy_sensible = 0, store_canopy_refreeze = 0, store_deltaH = 0, store_fusion = 0, store_grnd_flux = 0,
         store_latent = 0, store_latent_sub = 0, store_melt_energy = 0, store_sensible = 0;
  double Le, LongUnderIn, LongUnderOut, NetLongSnow, NetShortSnow, NetShortGrnd, OldTSurf, ShortUnderIn,
         Tair, Tcanopy, Tgrnd, Tsurf, VPDcanopy, VPcanopy, coverage, delta_coverage, ppt, rainfall,
         snowfall, snow_flux, snow_grnd_flux, step_melt, step_melt_energy, step_out_prec, step_out_rain,
         step_out_snow, step_ppt, step_prec, step_Wdew, dryFrac, snow_inflow = 0, last_snow_coverage;
  double store_advected_sensible = 0, store_advection = 0, store_deltaCC = 0, store_snow_flux = 0,
         store_refreeze_energy = 0, store_aero_cond_used[2], store_layerevap[MAXL],
         store_throughfall = 0, store_canopyevap = 0, store_canopy_vapor_flux = 0;
  oenergy snow_energy, soil_energy, iter_snow_energy, iter_soil_energy;
  ovegvar snow_veg_var, soil_veg_var, iter_snow_veg_var, iter_soil_veg_var;
  osnow step_snow, iter_snow;
  olayer step_layer[MAXL], iter_layer[MAXL];
  double iter_aero_resist[3], iter_aero_resist_used[2];
  double *aero_resist_used = cell->aero_resist;
  olayer *layer = cell->layer;
  (void)Tsurf; (void)UNSTABLE_SNOW;

  energy->advection = 0;
  energy->deltaCC = 0;
  if (snow->swq > 0) snow_flux = energy->snow_flux;
  else snow_flux = -(energy->grnd_flux + energy->deltaH + energy->fusion);
  energy->refreeze_energy = 0;
  coverage = snow->coverage;
  snow_energy = (*energy);
  soil_energy = (*energy);
  snow_veg_var = (*veg_var);
  soil_veg_var = (*veg_var);
  step_snow = (*snow);
  for (lidx = 0; lidx < NLAYER; lidx++) step_layer[lidx] = layer[lidx];
  for (lidx = 0; lidx < NLAYER; lidx++) step_layer[lidx].evap = 0;
  soil_veg_var.canopyevap = 0;
  snow_veg_var.canopyevap = 0;
  soil_veg_var.throughfall = 0;
  snow_veg_var.throughfall = 0;
  last_snow_coverage = snow->coverage;
  step_Wdew = veg_var->Wdew;

  /* ---- sub-step controls (surface_fluxes.c:385-396) ---- */
  if (snow->swq > 0 || snow->snow_canopy > 0 || atm[NR].snowflag) {
    hidx = 0; endhidx = hidx + NF; step_dt = opt->snow_step_hours;
  }
  else {
    hidx = NR; endhidx = hidx + 1; step_dt = opt->dt_hours;
  }
  N_steps = 0;
  for (lidx = 0; lidx < MAXL; lidx++) store_layerevap[lidx] = 0;
  store_aero_cond_used[0] = store_aero_cond_used[1] = 0;
  for (lidx = 0; lidx < 6; lidx++) store_pot_evap[lidx] = 0;

  /* ---- the sub-steps (surface_fluxes.c:480-1025) ---- */
  do {
    const oatmos *atmos = &atm[hidx];
    Tair = atmos->air_temp + soil_con->Tfactor[band];
    step_prec = atmos->prec * soil_con->Pfactor[band];
    Tgrnd = energy->T[0];
    Tcanopy = Tair;
    VPcanopy = atmos->vp;
    VPDcanopy = atmos->vpd;
    step_snow.blowing_flux = 0.0;      /* BLOWING FALSE */
    UnderStory = 999;
    snow_grnd_flux = -snow_flux;
    iter_snow_energy = snow_energy;
    iter_soil_energy = soil_energy;
    iter_snow_veg_var = snow_veg_var;
    iter_soil_veg_var = soil_veg_var;
    iter_snow = step_snow;
    for (lidx = 0; lidx < NLAYER; lidx++) iter_layer[lidx] = step_layer[lidx];
    iter_snow_veg_var.Wdew = step_Wdew;
    iter_soil_veg_var.Wdew = step_Wdew;
    iter_snow_veg_var.canopyevap = 0;
    iter_soil_veg_var.canopyevap = 0;
    for (lidx = 0; lidx < NLAYER; lidx++) iter_layer[lidx].evap = 0;
    iter_aero_resist[0] = aero_resist[0]; iter_aero_resist[1] = aero_resist[1]; iter_aero_resist[2] = aero_resist[2];
    iter_aero_resist_used[0] = aero_resist_used[0];
    iter_aero_resist_used[1] = aero_resist_used[1];
    iter_snow.canopy_vapor_flux = 0;
    iter_snow.vapor_flux = 0;
    iter_snow.surface_flux = 0;
    LongUnderOut = iter_soil_energy.LongUnderOut;
    dryFrac = -1;
    step_melt = solve_snow(overstory, BareAlbedo, LongUnderOut, opt->min_rain_temp, opt->max_snow_temp, Tcanopy,
                           Tgrnd, Tair, step_prec, snow_grnd_flux, &energy->AlbedoUnder, &Le, &LongUnderIn,
                           &NetLongSnow, &NetShortGrnd, &NetShortSnow, &ShortUnderIn, &OldTSurf,
                           iter_aero_resist, iter_aero_resist_used, &coverage, &delta_coverage, displacement,
                           &step_melt_energy, &step_out_prec, &step_out_rain, &step_out_snow, &step_ppt,
                           &rainfall, ref_height, roughness, &snow_inflow, &snowfall, &surf_atten, wind, root,
                           UNSTABLE_SNOW, Nveg, iveg, step_dt, vl, &UnderStory, &dryFrac, atmos,
                           &iter_snow_energy, iter_layer, &iter_snow, soil_con, &iter_snow_veg_var);
    if ((iter_snow.surf_temp == 999 || UNSTABLE_SNOW) && iter_snow.swq > 0) {
      INCLUDE_SNOW = UnderStory + 1;
      iter_soil_energy.advection = iter_snow_energy.advection;
      iter_snow.surf_temp = step_snow.surf_temp;
      step_melt_energy = 0;
    }
    else INCLUDE_SNOW = 0;
    Tsurf = calc_surf_energy_bal(Le, LongUnderIn, NetLongSnow, NetShortGrnd, NetShortSnow, OldTSurf,
                                 ShortUnderIn, iter_snow.albedo, iter_snow_energy.latent,
                                 iter_snow_energy.latent_sub, iter_snow_energy.sensible, Tcanopy, VPDcanopy,
                                 VPcanopy, delta_coverage, dp, step_melt_energy, iter_snow.coverage,
                                 (step_snow.depth + iter_snow.depth) / 2., BareAlbedo, surf_atten,
                                 iter_aero_resist, iter_aero_resist_used, displacement, &step_melt, &step_ppt,
                                 rainfall, ref_height, roughness, wind, root, INCLUDE_SNOW, UnderStory, Nveg,
                                 step_dt, iveg, overstory, vl, &dryFrac, atmos, &iter_soil_energy, iter_layer,
                                 &iter_snow, soil_con, &iter_soil_veg_var);
    if (INCLUDE_SNOW) step_ppt += step_melt;
    if (iter_snow.snow && overstory) {
      /* calc_atmos_energy_bal.c with CLOSE_ENERGY FALSE: Tcanopy = Tair; F = 1 */
      double InOverSensible = iter_snow_energy.canopy_sensible, InUnderSensible = iter_soil_energy.sensible;
      double F = 1;
      iter_soil_energy.AtmosSensible = InOverSensible + InUnderSensible;
      iter_soil_energy.NetLongAtmos = (F * iter_snow_energy.NetLongOver + (1. - F) * iter_soil_energy.NetLongUnder);
      iter_soil_energy.NetShortAtmos = (iter_snow_energy.NetShortOver + iter_soil_energy.NetShortUnder);
      iter_soil_energy.AtmosLatent = (iter_snow_energy.canopy_latent + iter_soil_energy.latent);
      iter_soil_energy.AtmosLatentSub = (iter_snow_energy.canopy_latent_sub + iter_soil_energy.latent_sub);
      Tcanopy = Tair;
      /* func_atmos_energy_bal.c: SensibleHeat = density*Cp*(Tair-Tcanopy)/Ra */
      iter_soil_energy.AtmosSensible = atmos->density * Cp * (Tair - Tcanopy) / iter_aero_resist_used[1];
    }
    else {
      iter_soil_energy.AtmosLatent = iter_soil_energy.latent;
      iter_soil_energy.AtmosLatentSub = iter_soil_energy.latent_sub;
      iter_soil_energy.AtmosSensible = iter_soil_energy.sensible;
      iter_soil_energy.NetLongAtmos = iter_soil_energy.NetLongUnder;
      iter_soil_energy.NetShortAtmos = iter_soil_energy.NetShortUnder;
    }
    iter_soil_energy.Tcanopy = Tcanopy;
    iter_snow_energy.Tcanopy = Tcanopy;

    /* potential evaporation of the six reference covers (surface_fluxes.c:868-897, 1019, 1123) */
    if (pet_aero) {
      double stability_factor[2], step_aero_resist[6][2], iter_pot_evap[6];
      int p;
      if (iter_aero_resist_used[0] == HUGE_RESIST) stability_factor[0] = HUGE_RESIST;
      else stability_factor[0] = iter_aero_resist_used[0] / pet_aero[6][UnderStory];
      if (iter_aero_resist_used[1] == iter_aero_resist_used[0]) stability_factor[1] = stability_factor[0];
      else {
        if (iter_aero_resist_used[1] == HUGE_RESIST) stability_factor[1] = HUGE_RESIST;
        else stability_factor[1] = iter_aero_resist_used[1] / pet_aero[6][1];
      }
      for (p = 0; p < 6; p++) {
        if (stability_factor[0] == HUGE_RESIST) step_aero_resist[p][0] = HUGE_RESIST;
        else step_aero_resist[p][0] = pet_aero[p][UnderStory] * stability_factor[0];
        if (stability_factor[1] == HUGE_RESIST) step_aero_resist[p][1] = HUGE_RESIST;
        else step_aero_resist[p][1] = pet_aero[p][1] * stability_factor[1];
      }
      compute_pot_evap(lib, nclass, veg_class, atmos->month - 1, opt->dt_hours, atmos->shortwave, iter_soil_energy.NetLongAtmos,
                       Tair, VPDcanopy, soil_con->elevation, step_aero_resist, iter_pot_evap);
      for (p = 0; p < 6; p++) store_pot_evap[p] += iter_pot_evap[p];
    }

    /* ---- store the sub-step (surface_fluxes.c:905-1025) ---- */
    snow_energy = iter_snow_energy;
    soil_energy = iter_soil_energy;
    snow_veg_var = iter_snow_veg_var;
    soil_veg_var = iter_soil_veg_var;
    step_snow = iter_snow;
    for (lidx = 0; lidx < NLAYER; lidx++) step_layer[lidx] = iter_layer[lidx];
    if (iveg != Nveg) {
      if (step_snow.snow) {
        store_throughfall += snow_veg_var.throughfall;
        store_canopyevap += snow_veg_var.canopyevap;
        soil_veg_var.Wdew = snow_veg_var.Wdew;
      }
      else {
        store_throughfall += soil_veg_var.throughfall;
        store_canopyevap += soil_veg_var.canopyevap;
        snow_veg_var.Wdew = soil_veg_var.Wdew;
      }
      step_Wdew = soil_veg_var.Wdew;
    }
    for (lidx = 0; lidx < NLAYER; lidx++) store_layerevap[lidx] += step_layer[lidx].evap;
    store_ppt += step_ppt;
    if (iter_aero_resist_used[0] > 0) store_aero_cond_used[0] += 1 / iter_aero_resist_used[0];
    else store_aero_cond_used[0] += HUGE_RESIST;
    if (iter_aero_resist_used[1] > 0) store_aero_cond_used[1] += 1 / iter_aero_resist_used[1];
    else store_aero_cond_used[1] += HUGE_RESIST;
    if (iveg != Nveg) store_canopy_vapor_flux += step_snow.canopy_vapor_flux;
    store_melt += step_melt;
    store_vapor_flux += step_snow.vapor_flux;
    store_surface_flux += step_snow.surface_flux;
    store_blowing_flux += step_snow.blowing_flux;
    *out_prec += step_out_prec;
    *out_rain += step_out_rain;
    *out_snow += step_out_snow;
    if (INCLUDE_SNOW) {
      snow_energy.advected_sensible = soil_energy.advected_sensible;
      snow_energy.advection = soil_energy.advection;
      snow_energy.deltaCC = soil_energy.deltaCC;
      snow_energy.latent = soil_energy.latent;
      snow_energy.latent_sub = soil_energy.latent_sub;
      snow_energy.refreeze_energy = soil_energy.refreeze_energy;
      snow_energy.sensible = soil_energy.sensible;
      snow_energy.snow_flux = soil_energy.snow_flux;
    }
    store_AlbedoOver += snow_energy.AlbedoOver;
    store_AlbedoUnder += soil_energy.AlbedoUnder;
    store_AtmosLatent += soil_energy.AtmosLatent;
    store_AtmosLatentSub += soil_energy.AtmosLatentSub;
    store_AtmosSensible += soil_energy.AtmosSensible;
    store_LongOverIn += snow_energy.LongOverIn;
    store_LongUnderIn += LongUnderIn;
    store_LongUnderOut += soil_energy.LongUnderOut;
    store_NetLongAtmos += soil_energy.NetLongAtmos;
    store_NetLongOver += snow_energy.NetLongOver;
    store_NetLongUnder += soil_energy.NetLongUnder;
    store_NetShortAtmos += soil_energy.NetShortAtmos;
    store_NetShortGrnd += NetShortGrnd;
    store_NetShortOver += snow_energy.NetShortOver;
    store_NetShortUnder += soil_energy.NetShortUnder;
    store_ShortOverIn += snow_energy.ShortOverIn;
    store_ShortUnderIn += soil_energy.ShortUnderIn;
    store_canopy_advection += snow_energy.canopy_advection;
    store_canopy_latent += snow_energy.canopy_latent;
    store_canopy_latent_sub += snow_energy.canopy_latent_sub;
    store_canopy_sensible += snow_energy.canopy_sensible;
    store_canopy_refreeze += snow_energy.canopy_refreeze;
    store_deltaH += soil_energy.deltaH;
    store_fusion += soil_energy.fusion;
    store_grnd_flux += soil_energy.grnd_flux;
    store_latent += soil_energy.latent;
    store_latent_sub += soil_energy.latent_sub;
    store_melt_energy += step_melt_energy;
    store_sensible += soil_energy.sensible;
    if (step_snow.swq == 0 && INCLUDE_SNOW) {
      if (last_snow_coverage == 0 && step_prec > 0) last_snow_coverage = 1;
      store_advected_sensible += snow_energy.advected_sensible * last_snow_coverage;
      store_advection += snow_energy.advection * last_snow_coverage;
      store_deltaCC += snow_energy.deltaCC * last_snow_coverage;
      store_snow_flux += soil_energy.snow_flux * last_snow_coverage;
      store_refreeze_energy += snow_energy.refreeze_energy * last_snow_coverage;
    }
    else if (step_snow.snow || INCLUDE_SNOW) {
      store_advected_sensible += snow_energy.advected_sensible * (step_snow.coverage + delta_coverage);
      store_advection += snow_energy.advection * (step_snow.coverage + delta_coverage);
      store_deltaCC += snow_energy.deltaCC * (step_snow.coverage + delta_coverage);
      store_snow_flux += soil_energy.snow_flux * (step_snow.coverage + delta_coverage);
      store_refreeze_energy += snow_energy.refreeze_energy * (step_snow.coverage + delta_coverage);
    }
    N_steps++;
    hidx++;
  } while (hidx < endhidx);

  /* ---- fold the sub-steps back (surface_fluxes.c:1030-1160) ---- */
  (*snow) = step_snow;
  snow->vapor_flux = store_vapor_flux;
  snow->blowing_flux = store_blowing_flux;
  snow->surface_flux = store_surface_flux;
  snow->canopy_vapor_flux = store_canopy_vapor_flux;
  snow->melt = store_melt;
  ppt = store_ppt;
  (*energy) = soil_energy;
  energy->AlbedoOver = store_AlbedoOver / (double)N_steps;
  energy->AlbedoUnder = store_AlbedoUnder / (double)N_steps;
  energy->AtmosLatent = store_AtmosLatent / (double)N_steps;
  energy->AtmosLatentSub = store_AtmosLatentSub / (double)N_steps;
  energy->AtmosSensible = store_AtmosSensible / (double)N_steps;
  energy->LongOverIn = store_LongOverIn / (double)N_steps;
  energy->LongUnderIn = store_LongUnderIn / (double)N_steps;
  energy->LongUnderOut = store_LongUnderOut / (double)N_steps;
  energy->NetLongAtmos = store_NetLongAtmos / (double)N_steps;
  energy->NetLongOver = store_NetLongOver / (double)N_steps;
  energy->NetLongUnder = store_NetLongUnder / (double)N_steps;
  energy->NetShortAtmos = store_NetShortAtmos / (double)N_steps;
  energy->NetShortGrnd = store_NetShortGrnd / (double)N_steps;
  energy->NetShortOver = store_NetShortOver / (double)N_steps;
  energy->NetShortUnder = store_NetShortUnder / (double)N_steps;
  energy->ShortOverIn = store_ShortOverIn / (double)N_steps;
  energy->ShortUnderIn = store_ShortUnderIn / (double)N_steps;
  energy->advected_sensible = store_advected_sensible / (double)N_steps;
  energy->canopy_advection = store_canopy_advection / (double)N_steps;
  energy->canopy_latent = store_canopy_latent / (double)N_steps;
  energy->canopy_latent_sub = store_canopy_latent_sub / (double)N_steps;
  energy->canopy_refreeze = store_canopy_refreeze / (double)N_steps;
  energy->canopy_sensible = store_canopy_sensible / (double)N_steps;
  energy->deltaH = store_deltaH / (double)N_steps;
  energy->fusion = store_fusion / (double)N_steps;
  energy->grnd_flux = store_grnd_flux / (double)N_steps;
  energy->latent = store_latent / (double)N_steps;
  energy->latent_sub = store_latent_sub / (double)N_steps;
  energy->melt_energy = store_melt_energy / (double)N_steps;
  energy->sensible = store_sensible / (double)N_steps;
  if (snow->snow || INCLUDE_SNOW) {
    energy->advection = store_advection / (double)N_steps;
    energy->deltaCC = store_deltaCC / (double)N_steps;
    energy->refreeze_energy = store_refreeze_energy / (double)N_steps;
    energy->snow_flux = store_snow_flux / (double)N_steps;
  }
  energy->Tfoliage = snow_energy.Tfoliage;
  energy->Tfoliage_fbflag = snow_energy.Tfoliage_fbflag;
  energy->Tfoliage_fbcount = snow_energy.Tfoliage_fbcount;
  if (iveg != Nveg) {
    veg_var->throughfall = store_throughfall;
    veg_var->canopyevap = store_canopyevap;
    if (snow->snow) veg_var->Wdew = snow_veg_var.Wdew;
    else veg_var->Wdew = soil_veg_var.Wdew;
  }
  for (lidx = 0; lidx < NLAYER; lidx++) {
    layer[lidx] = step_layer[lidx];
    layer[lidx].evap = store_layerevap[lidx];
  }
  for (lidx = 0; lidx < 2; lidx++) {
    if (store_aero_cond_used[lidx] > 0 && store_aero_cond_used[lidx] < HUGE_RESIST)
      aero_resist_used[lidx] = 1 / (store_aero_cond_used[lidx] / (double)N_steps);
    else if (store_aero_cond_used[lidx] >= HUGE_RESIST) aero_resist_used[lidx] = 0;
    else aero_resist_used[lidx] = HUGE_RESIST;
  }
  if (pet_aero) for (lidx = 0; lidx < 6; lidx++) cell->pot_evap[lidx] = store_pot_evap[lidx] / (double)N_steps;
  cell->inflow = ppt;
  return runoff(cell, soil_con, ppt, opt->dt_hours);
}

/* ======================================================================================
 * full_energy.c:8-571 (one cell, one record) -- without lakes / carbon / PET reference covers
 * ====================================================================================== */
static int full_energy(vo_model *m, int64_t c, const oatmos *atm /* [NF + 1], or [1] when NF == 1 */, double *out_prec_sum,
                       double *out_rain_sum, double *out_snow_sum)
{
  const int NF = m->opt.dt_hours / m->opt.snow_step_hours, NR = NF > 1 ? NF : 0;
  const oatmos *atmos = &atm[NR];
  const vr_options *opt = &m->opt;
  const osoil *soil_con = &m->soil[c];
  int64_t t0 = m->tile_ptr[c], t1 = m->tile_ptr[c + 1], t;
  int NB = opt->nbands, band, lidx, mon = atmos->month - 1, Nveg, iveg, ntile = (int)(t1 - t0);
  int has_bare = (ntile > 0 && m->tile_class[t1 - 1] == m->nclass);
  float root0[MAXL], rootv[MAXL];
  double out_prec, out_rain, out_snow;
  Nveg = has_bare ? ntile - 1 : ntile;
  *out_prec_sum = *out_rain_sum = *out_snow_sum = 0;
  for (lidx = 0; lidx < NLAYER; lidx++) root0[lidx] = (Nveg > 0) ? (float)m->tile_root[t0 * NLAYER + lidx] : 0.f;

  /* veg_hist -> veg_var, and the "/= vegcover" of full_energy.c:210-227 */
  for (iveg = 0; iveg < Nveg; iveg++) {
    double vegcover, albedo, LAI;
    t = t0 + iveg;
    vegcover = m->tile_vegcover[t * 12 + mon];
    albedo = m->tile_albedo[t * 12 + mon];
    LAI = m->tile_LAI[t * 12 + mon];
    if (vegcover < MIN_VEGCOVER) vegcover = MIN_VEGCOVER;
    for (band = 0; band < NB; band++) {
      ounit *u = &m->unit[t * NB + band];
      u->veg.vegcover = vegcover;
      u->veg.albedo = albedo;
      u->veg.LAI = LAI;
      u->veg.LAI /= u->veg.vegcover;
      u->veg.Wdew /= u->veg.vegcover;
      u->veg.Wdmax = u->veg.LAI * LAI_WATER_FACTOR;
      u->snow.snow_canopy /= u->veg.vegcover;
    }
  }
  for (iveg = 0; iveg < ntile; iveg++) {
    double Cv, wind_h, surf_atten, bare_albedo, height, displacement[3], roughness[3], ref_height[3],
           tmp_wind[3], aero_resist[3], pet_aero[7][3];
    int overstory, veg_class, is_bare;
    oveglib vl;
    t = t0 + iveg;
    Cv = m->tile_Cv[t];
    if (!(Cv > 0.0)) continue;
    veg_class = m->tile_class[t];
    is_bare = (veg_class == m->nclass);
    vl = m->lib[veg_class];
    if (is_bare) {     /* read_soilparam.c:551-554 overwrites the bare entry per cell */
      int j;
      for (j = 0; j < 12; j++) {
        vl.roughness[j] = soil_con->rough;
        vl.displacement[j] = soil_con->rough * 0.667 / 0.123;
      }
    }
    for (band = 0; band < NB; band++) {
      if (soil_con->AreaFract[band] > 0) {
        ounit *u = &m->unit[t * NB + band];
        u->snow.vapor_flux = 0.;
        u->snow.canopy_vapor_flux = 0.;
        if (!is_bare) u->veg.rc = HUGE_RESIST;
      }
    }
    wind_h = vl.wind_h;
    {
      ounit *u0 = &m->unit[t * NB + 0];
      surf_atten = (1 - u0->veg.vegcover) * 1.0 + u0->veg.vegcover * exp(-vl.rad_atten * u0->veg.LAI);
      /* prepare_full_energy.c:55-110: layer thermal properties from the pre-step moisture */
      for (band = 0; band < NB; band++) {
        if (soil_con->AreaFract[band] > 0.0) {
          ounit *u = &m->unit[t * NB + band];
          olayer tmpl[MAXL];
          for (lidx = 0; lidx < NLAYER; lidx++) tmpl[lidx] = u->cell.layer[lidx];
          compute_soil_layer_thermal_properties(tmpl, soil_con, NLAYER);
          u->energy.kappa[0] = tmpl[0].kappa; u->energy.Cs[0] = tmpl[0].Cs;
          u->energy.kappa[1] = tmpl[1].kappa; u->energy.Cs[1] = tmpl[1].Cs;
        }
      }
      if (!is_bare) bare_albedo = u0->veg.albedo;
      else bare_albedo = BARE_SOIL_ALBEDO;
    }
    /* the p == N_PET_TYPES pass of full_energy.c:329-369 (current vegetation) */
    tmp_wind[0] = atmos->wind; tmp_wind[1] = VERROR; tmp_wind[2] = VERROR;
    displacement[0] = vl.displacement[mon];
    roughness[0] = vl.roughness[mon];
    overstory = vl.overstory;
    if (roughness[0] == 0) roughness[0] = soil_con->rough;
    height = displacement[0] / RATIO_DH_HEIGHT_VEG;   /* calc_veg_height */
    if (displacement[0] < wind_h) ref_height[0] = wind_h;
    else ref_height[0] = displacement[0] + wind_h + roughness[0];
    if (CalcAerodynamic(overstory, height, vl.trunk_ratio, soil_con->snow_rough, soil_con->rough, vl.wind_atten,
                        aero_resist, tmp_wind, displacement, ref_height, roughness) == VERROR)
      return VERROR;
    /* the p < N_PET_TYPES passes (full_energy.c:329-369): the reference covers, then twice the current one */
    if (m->want_pet) {
      int p;
      for (p = 0; p < 7; p++) {
        oveglib pv = (p < 4) ? m->lib[m->nclass + p] : vl;
        double pd[3], pr[3], ph[3], pw[3];
        if (p == 0) {       /* SATSOIL is the bare-soil entry read_soilparam.c:551-554 overwrote for this cell */
          pv.roughness[mon] = soil_con->rough;
          pv.displacement[mon] = soil_con->rough * 0.667 / 0.123;
        }
        pw[0] = atmos->wind; pw[1] = VERROR; pw[2] = VERROR;
        pd[0] = pv.displacement[mon]; pr[0] = pv.roughness[mon];
        if (p >= 4 && pr[0] == 0) pr[0] = soil_con->rough;
        if (pd[0] < wind_h) ph[0] = wind_h;
        else ph[0] = pd[0] + wind_h + pr[0];
        if (CalcAerodynamic(pv.overstory, pd[0] / RATIO_DH_HEIGHT_VEG, pv.trunk_ratio, soil_con->snow_rough, soil_con->rough,
                            pv.wind_atten, pet_aero[p], pw, pd, ph, pr) == VERROR)
          return VERROR;
      }
    }
    for (lidx = 0; lidx < NLAYER; lidx++) rootv[lidx] = is_bare ? 0.f : (float)m->tile_root[t * NLAYER + lidx];
    for (band = 0; band < NB; band++) {
      if (soil_con->AreaFract[band] > 0) {
        ounit *u = &m->unit[t * NB + band];
        double d3[3], rh3[3], ro3[3], w3[3], ar3[3];
        int k, err;
        u->cell.aero_resist[0] = aero_resist[0];
        u->cell.aero_resist[1] = aero_resist[1];
        for (k = 0; k < 3; k++) { d3[k] = displacement[k]; rh3[k] = ref_height[k]; ro3[k] = roughness[k]; w3[k] = tmp_wind[k]; ar3[k] = aero_resist[k]; }
        out_prec = out_rain = out_snow = 0;
        err = surface_fluxes(opt, overstory, bare_albedo, surf_atten, ar3, d3, rh3, ro3, w3, rootv,
                             is_bare ? iveg : -1 /* Nveg: iveg == Nveg only for the bare tile */, band,
                             soil_con->dp, iveg, &vl, atm, NF, NR, &u->energy, &u->cell, &u->snow, soil_con, &u->veg,
                             &out_prec, &out_rain, &out_snow, m->lib, m->nclass, veg_class, m->want_pet ? pet_aero : NULL);
        if (err == VERROR) return VERROR;
        *out_prec_sum += out_prec * Cv * soil_con->AreaFract[band];
        *out_rain_sum += out_rain * Cv * soil_con->AreaFract[band];
        *out_snow_sum += out_snow * Cv * soil_con->AreaFract[band];
        u->cell.rootmoist = 0;
        u->cell.wetness = 0;
        for (lidx = 0; lidx < NLAYER; lidx++) {
          if (root0[lidx] > 0) u->cell.rootmoist += u->cell.layer[lidx].moist;   /* veg_con->root: tile 0 */
          u->cell.wetness += (u->cell.layer[lidx].moist - soil_con->Wpwp[lidx]) /
                             (soil_con->porosity[lidx] * soil_con->depth[lidx] * 1000 - soil_con->Wpwp[lidx]);
        }
        u->cell.wetness /= NLAYER;
      }
    }
  }
  for (iveg = 0; iveg < Nveg; iveg++) {
    t = t0 + iveg;
    for (band = 0; band < NB; band++) {
      ounit *u = &m->unit[t * NB + band];
      u->veg.LAI *= u->veg.vegcover;
      u->veg.Wdmax *= u->veg.vegcover;
    }
  }
  return 0;
}

/* ======================================================================================
 * put_data.c:7-790 + collect_wb_terms/collect_eb_terms (:792-1221): tile x band -> cell sums
 * ====================================================================================== */
static void put_data(vo_model *m, int64_t c, const oatmos *atmos, double out_prec, double out_rain,
                     double out_snow, int init, double *o /* [VR_NOUTVAR] */)
{
  const vr_options *opt = &m->opt;
  const osoil *soil_con = &m->soil[c];
  int64_t t0 = m->tile_ptr[c], t1 = m->tile_ptr[c + 1], t;
  int NB = opt->nbands, band, index;
  double cv_snow = 0, aero_cond = 0, soil_moist_tot, inflow, outflow, storage;
  for (index = 0; index < VR_NOUTVAR; index++) o[index] = 0;
  o[VR_OUT_AIR_TEMP] = atmos->air_temp;
  o[VR_OUT_LONGWAVE] = atmos->longwave;
  o[VR_OUT_PREC] = out_prec;
  o[VR_OUT_RAINF] = out_rain;
  o[VR_OUT_REL_HUMID] = 100. * atmos->vp / (atmos->vp + atmos->vpd);
  o[VR_OUT_DENSITY] = atmos->density;                 /* put_data.c:275-291 */
  o[VR_OUT_PRESSURE] = atmos->pressure / 1000.;
  o[VR_OUT_VP] = atmos->vp / 1000.;
  o[VR_OUT_VPD] = atmos->vpd / 1000.;
  o[VR_OUT_QAIR] = 0.62196351 * atmos->vp / atmos->pressure;
  o[VR_OUT_SHORTWAVE] = atmos->shortwave;
  o[VR_OUT_SNOWF] = out_snow;
  o[VR_OUT_WIND] = atmos->wind;
  for (t = t0; t < t1; t++) {
    double Cv = m->tile_Cv[t];
    int HasVeg = (m->tile_class[t] != m->nclass);
    int overstory = m->lib[m->tile_class[t]].overstory;
    if (!(Cv > 0)) continue;
    for (band = 0; band < NB; band++) {
      double ThisAreaFract = soil_con->AreaFract[band], AreaFactor, tmp_evap, tmp_cond1, tmp_cond2 = 0;
      const ounit *u = &m->unit[t * NB + band];
      if (!(ThisAreaFract > 0.)) continue;
      if (u->snow.swq > 0.0) cv_snow += Cv * ThisAreaFract * 1.;
      /* collect_wb_terms */
      AreaFactor = Cv * ThisAreaFract * 1. * 1.;
      tmp_evap = 0.0;
      for (index = 0; index < NLAYER; index++) {
        tmp_evap += u->cell.layer[index].evap;
        if (HasVeg) {
          o[VR_OUT_EVAP_BARE] += u->cell.layer[index].evap * u->cell.layer[index].bare_evap_frac * AreaFactor;
          o[VR_OUT_TRANSP_VEG] += u->cell.layer[index].evap * (1 - u->cell.layer[index].bare_evap_frac) * AreaFactor;
        }
        else o[VR_OUT_EVAP_BARE] += u->cell.layer[index].evap * AreaFactor;
      }
      tmp_evap += u->snow.vapor_flux * 1000.;
      o[VR_OUT_SUB_SNOW] += u->snow.vapor_flux * 1000. * AreaFactor;
      if (HasVeg) {
        tmp_evap += u->snow.canopy_vapor_flux * 1000.;
        o[VR_OUT_SUB_CANOP] += u->snow.canopy_vapor_flux * 1000. * AreaFactor;
      }
      if (HasVeg) {
        tmp_evap += u->veg.canopyevap;
        o[VR_OUT_EVAP_CANOP] += u->veg.canopyevap * AreaFactor;
      }
      o[VR_OUT_EVAP] += tmp_evap * AreaFactor;
      o[VR_OUT_ASAT] += u->cell.asat * AreaFactor;
      o[VR_OUT_RUNOFF] += u->cell.runoff * AreaFactor;
      o[VR_OUT_BASEFLOW] += u->cell.baseflow * AreaFactor;
      o[VR_OUT_INFLOW] += (u->cell.inflow) * AreaFactor;
      if (HasVeg) o[VR_OUT_WDEW] += u->veg.Wdew * AreaFactor;
      if (u->cell.aero_resist[0] > SMALL) tmp_cond1 = (1 / u->cell.aero_resist[0]) * AreaFactor;
      else tmp_cond1 = HUGE_RESIST;
      if (overstory) {
        if (u->cell.aero_resist[1] > SMALL) tmp_cond2 = (1 / u->cell.aero_resist[1]) * AreaFactor;
        else tmp_cond2 = HUGE_RESIST;
      }
      if (overstory) aero_cond += tmp_cond2;
      else aero_cond += tmp_cond1;
      for (index = 0; index < NLAYER; index++) {
        double tmp_moist = u->cell.layer[index].moist, tmp_ice = 0;
        tmp_ice += (0. * 1.);
        tmp_moist -= tmp_ice;
        if (opt->moistfract) tmp_moist /= soil_con->depth[index] * 1000.;      /* put_data.c:906-909 */
        o[VR_OUT_SOIL_LIQ0 + index] += tmp_moist * AreaFactor;
      }
      o[VR_OUT_SOIL_WET] += u->cell.wetness * AreaFactor;
      o[VR_OUT_ROOTMOIST] += u->cell.rootmoist * AreaFactor;
      for (index = 0; index < 6; index++) o[VR_OUT_PET_SATSOIL + index] += u->cell.pot_evap[index] * AreaFactor;   /* put_data.c:846-851 */
      o[VR_OUT_ZWT] += u->cell.zwt * AreaFactor;                                                                   /* :917-918 */
      o[VR_OUT_ZWT_LUMPED] += u->cell.zwt_lumped * AreaFactor;
      o[VR_OUT_SWE] += u->snow.swq * AreaFactor * 1000.;
      o[VR_OUT_SNOW_DEPTH] += u->snow.depth * AreaFactor * 100.;
      if (u->snow.swq > 0.0) {
        o[VR_OUT_SALBEDO] += u->snow.albedo * AreaFactor;
        o[VR_OUT_SNOW_SURF_TEMP] += u->snow.surf_temp * AreaFactor;
        o[VR_OUT_SNOW_PACK_TEMP] += u->snow.pack_temp * AreaFactor;
      }
      if (HasVeg) o[VR_OUT_SNOW_CANOPY] += (u->snow.snow_canopy) * AreaFactor * 1000.;
      o[VR_OUT_SNOW_MELT] += u->snow.melt * AreaFactor;
      o[VR_OUT_SNOW_COVER] += u->snow.coverage * AreaFactor;
      /* collect_eb_terms */
      o[VR_OUT_SURF_TEMP] += u->energy.Tsurf * AreaFactor;
      o[VR_OUT_NET_SHORT] += u->energy.NetShortAtmos * AreaFactor;
      o[VR_OUT_NET_LONG] += u->energy.NetLongAtmos * AreaFactor;
      if (u->snow.snow && overstory) o[VR_OUT_IN_LONG] += u->energy.LongOverIn * AreaFactor;
      else o[VR_OUT_IN_LONG] += u->energy.LongUnderIn * AreaFactor;
      if (u->snow.snow && overstory) o[VR_OUT_ALBEDO] += u->energy.AlbedoOver * AreaFactor;
      else o[VR_OUT_ALBEDO] += u->energy.AlbedoUnder * AreaFactor;
      o[VR_OUT_GRND_FLUX] -= u->energy.grnd_flux * AreaFactor;
      o[VR_OUT_DELTAH] -= u->energy.deltaH * AreaFactor;
    }
  }
  if (cv_snow > 0) {
    o[VR_OUT_SALBEDO] /= cv_snow;
    o[VR_OUT_SNOW_SURF_TEMP] /= cv_snow;
    o[VR_OUT_SNOW_PACK_TEMP] /= cv_snow;
  }
  o[VR_OUT_AERO_COND] = aero_cond;
  if (aero_cond > SMALL) o[VR_OUT_AERO_RESIST] = 1 / aero_cond;
  else o[VR_OUT_AERO_RESIST] = HUGE_RESIST;
  soil_moist_tot = 0;
  o[VR_OUT_DELSOILMOIST] = 0;
  for (index = 0; index < NLAYER; index++) {
    double sm = o[VR_OUT_SOIL_LIQ0 + index] + 0.;   /* OUT_SOIL_MOIST = LIQ + ICE */
    o[VR_OUT_DELSOILMOIST] += sm;
    soil_moist_tot += sm;
  }
  if (!init) {
    o[VR_OUT_DELSOILMOIST] -= m->save_soil[c];
    o[VR_OUT_DELSWE] = o[VR_OUT_SWE] + o[VR_OUT_SNOW_CANOPY] - m->save_swe[c];
    o[VR_OUT_DELINTERCEPT] = o[VR_OUT_WDEW] - m->save_wdew[c];
  }
  o[VR_OUT_R_NET] = o[VR_OUT_NET_SHORT] + o[VR_OUT_NET_LONG];
  m->save_soil[c] = soil_moist_tot;
  m->save_swe[c] = o[VR_OUT_SWE] + o[VR_OUT_SNOW_CANOPY];
  m->save_wdew[c] = o[VR_OUT_WDEW];
  inflow = o[VR_OUT_PREC] + 0.;
  outflow = o[VR_OUT_EVAP] + o[VR_OUT_RUNOFF] + o[VR_OUT_BASEFLOW];
  storage = 0.;
  for (index = 0; index < NLAYER; index++) {      /* put_data.c:624-630 */
    if (opt->moistfract) storage += (o[VR_OUT_SOIL_LIQ0 + index] + 0.) * soil_con->depth[index] * 1000;
    else storage += o[VR_OUT_SOIL_LIQ0 + index] + 0.;
  }
  storage += o[VR_OUT_SWE] + o[VR_OUT_SNOW_CANOPY] + o[VR_OUT_WDEW] + 0.;
  /* calc_water_energy_balance_errors.c calc_water_balance_error */
  if (init) {
    m->last_storage[c] = storage;
    o[VR_OUT_WATER_ERROR] = 0.0;
  }
  else {
    o[VR_OUT_WATER_ERROR] = inflow - outflow - (storage - m->last_storage[c]);
    m->last_storage[c] = storage;
  }
}

/* ======================================================================================
 * model object
 * ====================================================================================== */
static double *dupd(const double *p, size_t n)
{
  double *q = (double *)malloc(sizeof(double) * (n ? n : 1));
  if (p) memcpy(q, p, sizeof(double) * n);
  else memset(q, 0, sizeof(double) * n);
  return q;
}

vo_model *vo_create(const vr_options *opt, const vr_veglib *lib, const vr_cells *cells)
{
  vo_model *m = (vo_model *)calloc(1, sizeof(vo_model));
  int64_t C = cells->ncell, T = cells->tile_ptr[C], c;
  int i, j, l, b, L = opt->nlayer, NB = opt->nbands;
  m->opt = *opt;
  m->nclass = lib->nclass;
  for (i = 0; i < opt->n_out; i++)
    if (opt->out_vars[i] >= VR_OUT_PET_SATSOIL && opt->out_vars[i] <= VR_OUT_PET_VEGNOCR) m->want_pet = 1;
  m->lib = (oveglib *)calloc(lib->nclass + 4, sizeof(oveglib));      /* + the 4 PET reference classes (read_veglib.c:168-187) */
  for (i = 0; i < lib->nclass; i++) {
    oveglib *v = &m->lib[i];
    v->overstory = lib->overstory[i]; v->rarc = lib->rarc[i]; v->rmin = lib->rmin[i];
    v->wind_h = lib->wind_h[i]; v->RGL = lib->RGL[i]; v->rad_atten = lib->rad_atten[i];
    v->wind_atten = lib->wind_atten[i]; v->trunk_ratio = lib->trunk_ratio[i];
    for (j = 0; j < 12; j++) { v->roughness[j] = lib->roughness[i * 12 + j]; v->displacement[j] = lib->displacement[i * 12 + j]; }
    for (j = 0; j < 12; j++) { v->LAI[j] = lib->LAI ? lib->LAI[i * 12 + j] : 0.0; v->albedo[j] = lib->albedo ? lib->albedo[i * 12 + j] : 0.0; }
  }
  { /* bare-soil pseudo class == PET_SATSOIL reference entry (global.h:9-21, read_veglib.c:136-153) */
    oveglib *v = &m->lib[lib->nclass];
    v->overstory = 0; v->rarc = 0.0; v->rmin = 0.0; v->wind_h = 10.0; v->RGL = 0.0; v->rad_atten = 0.0;
    v->wind_atten = 0.0; v->trunk_ratio = 0.0;
    for (j = 0; j < 12; j++) { v->roughness[j] = 0.001; v->displacement[j] = 0.0054; }
  }
  { /* global.h:53-65 */
    static const double ref_rarc[4] = {0.0, 0.0, 25, 25}, ref_rmin[4] = {0.0, 0.0, 100, 100}, ref_lai[4] = {1.0, 1.0, 2.88, 4.45},
                        ref_alb[4] = {BARE_SOIL_ALBEDO, 0.08, 0.23, 0.23}, ref_rough[4] = {0.001, 0.001, 0.0148, 0.0615},
                        ref_displ[4] = {0.0054, 0.0054, 0.08, 0.3333}, ref_RGL[4] = {0.0, 0.0, 100, 100};
    for (i = 0; i < 4; i++) {
      oveglib *v = &m->lib[lib->nclass + i];
      v->overstory = 0; v->rarc = ref_rarc[i]; v->rmin = ref_rmin[i]; v->wind_h = 10.0; v->RGL = ref_RGL[i];
      v->rad_atten = 0.0; v->wind_atten = 0.0; v->trunk_ratio = 0.0;
      for (j = 0; j < 12; j++) {
        v->roughness[j] = ref_rough[i]; v->displacement[j] = ref_displ[i]; v->LAI[j] = ref_lai[i]; v->albedo[j] = ref_alb[i];
      }
    }
  }
  m->C = C; m->T = T;
  m->soil = (osoil *)calloc(C ? C : 1, sizeof(osoil));
  for (c = 0; c < C; c++) {
    osoil *s = &m->soil[c];
    s->lat = cells->lat[c]; s->elevation = cells->elevation[c]; s->b_infilt = cells->b_infilt[c];
    s->Ds = cells->Ds[c]; s->Dsmax = cells->Dsmax[c]; s->Ws = cells->Ws[c]; s->c = cells->c_expt[c];
    s->avg_temp = cells->avg_temp[c]; s->dp = cells->dp[c]; s->rough = cells->rough[c];
    s->snow_rough = cells->snow_rough[c];
    for (l = 0; l < L; l++) {
      size_t k = (size_t)l * C + c;
      s->expt[l] = cells->expt[k]; s->Ksat[l] = cells->Ksat[k]; s->depth[l] = cells->depth[k];
      s->max_moist[l] = cells->max_moist[k]; s->Wcr[l] = cells->Wcr[k]; s->Wpwp[l] = cells->Wpwp[k];
      s->resid_moist[l] = cells->resid_moist[k]; s->init_moist[l] = cells->init_moist[k];
      s->bulk_density[l] = cells->bulk_density[k]; s->soil_density[l] = cells->soil_density[k];
      s->bulk_dens_min[l] = cells->bulk_dens_min[k]; s->soil_dens_min[l] = cells->soil_dens_min[k];
      s->quartz[l] = cells->quartz[k]; s->organic[l] = cells->organic[k]; s->porosity[l] = cells->porosity[k];
    }
    for (b = 0; b < NB; b++) {
      size_t k = (size_t)b * C + c;
      s->AreaFract[b] = cells->AreaFract[k]; s->Tfactor[b] = cells->Tfactor[k]; s->Pfactor[b] = cells->Pfactor[k];
    }
    s->have_zwt = (cells->zwt_zwt && cells->zwt_moist);
    if (s->have_zwt)
      for (l = 0; l < L + 2; l++)
        for (i = 0; i < VR_ZWT_NPOINTS; i++) {
          size_t k = ((size_t)l * VR_ZWT_NPOINTS + i) * C + c;
          s->zwtvmoist_zwt[l][i] = cells->zwt_zwt[k]; s->zwtvmoist_moist[l][i] = cells->zwt_moist[k];
        }
  }
  m->tile_ptr = (int64_t *)malloc(sizeof(int64_t) * (C + 1));
  memcpy(m->tile_ptr, cells->tile_ptr, sizeof(int64_t) * (C + 1));
  m->tile_class = (int32_t *)malloc(sizeof(int32_t) * (T ? T : 1));
  memcpy(m->tile_class, cells->tile_class, sizeof(int32_t) * T);
  m->tile_Cv = dupd(cells->tile_Cv, T);
  m->tile_root = dupd(cells->tile_root, (size_t)T * L);
  m->tile_LAI = dupd(cells->tile_LAI, (size_t)T * 12);
  m->tile_albedo = dupd(cells->tile_albedo, (size_t)T * 12);
  m->tile_vegcover = dupd(cells->tile_vegcover, (size_t)T * 12);
  m->unit = (ounit *)calloc((size_t)(T ? T : 1) * NB, sizeof(ounit));
  m->save_soil = dupd(NULL, C); m->save_swe = dupd(NULL, C); m->save_wdew = dupd(NULL, C);
  m->last_storage = dupd(NULL, C);
  return m;
}

void vo_destroy(vo_model *m)
{
  if (!m) return;
  free(m->lib); free(m->soil); free(m->tile_ptr); free(m->tile_class); free(m->tile_Cv); free(m->tile_root);
  free(m->tile_LAI); free(m->tile_albedo); free(m->tile_vegcover); free(m->unit); free(m->save_soil);
  free(m->save_swe); free(m->save_wdew); free(m->last_storage); free(m);
}

/* initialize_model_state.c:8-405 (QUICK_FLUX, no INIT_STATE) + initialize_snow/soil/veg +
 * the put_data(rec = -nrecs) call of vicNl.c:287 */
int vo_init_state(vo_model *m, const double *tair0)
{
  int64_t c, t;
  int NB = m->opt.nbands, band, l;
  NLAYER = m->opt.nlayer;
  for (c = 0; c < m->C; c++) {
    const osoil *s = &m->soil[c];
    double surf_temp = tair0[c], Tair = tair0[c], tmp, o[VR_NOUTVAR];
    oatmos a0;
    if (surf_temp < -1.) surf_temp = -1.;
    for (t = m->tile_ptr[c]; t < m->tile_ptr[c + 1]; t++) {
      for (band = 0; band < NB; band++) {
        ounit *u = &m->unit[t * NB + band];
        memset(u, 0, sizeof *u);
        u->snow.last_snow = (int)MISSING;
        for (l = 0; l < NLAYER; l++) {
          u->cell.layer[l].moist = s->init_moist[l];
          if (u->cell.layer[l].moist > s->max_moist[l]) u->cell.layer[l].moist = s->max_moist[l];
        }
        if (m->tile_Cv[t] > 0) {
          u->energy.T[0] = surf_temp;
          u->energy.T[1] = surf_temp;
          u->energy.T[2] = s->avg_temp;
          if (s->AreaFract[band] > 0.)
            estimate_layer_ice_content_quick_flux(u->cell.layer, s, u->energy.T[0], u->energy.T[1]);
        }
        tmp = u->energy.T[0] + KELVIN;
        u->energy.LongUnderOut = STEFAN_B * tmp * tmp * tmp * tmp;
        u->energy.Tfoliage = Tair + s->Tfactor[band];
      }
    }
    memset(&a0, 0, sizeof a0);
    a0.air_temp = tair0[c]; a0.vp = 1; a0.vpd = 1; a0.month = 1; a0.day_in_year = 1;
    put_data(m, c, &a0, 0., 0., 0., 1, o);
  }
  m->initialised = 1;
  return 0;
}

static void fill_unit_dbg(double *u, const ounit *x)
{
  int k = 0, l;
  for (l = 0; l < 3; l++) u[k++] = (l < NLAYER) ? x->cell.layer[l].moist : 0.;
  for (l = 0; l < 3; l++) u[k++] = (l < NLAYER) ? x->cell.layer[l].evap : 0.;
  for (l = 0; l < 3; l++) u[k++] = (l < NLAYER) ? x->cell.layer[l].bare_evap_frac : 0.;
  u[k++] = x->cell.runoff; u[k++] = x->cell.baseflow; u[k++] = x->cell.asat; u[k++] = x->cell.inflow;
  u[k++] = x->cell.aero_resist[0]; u[k++] = x->cell.aero_resist[1];
  u[k++] = x->veg.Wdew; u[k++] = x->veg.canopyevap; u[k++] = x->veg.throughfall;
  u[k++] = x->snow.swq; u[k++] = x->snow.surf_temp; u[k++] = x->snow.pack_temp; u[k++] = x->snow.surf_water;
  u[k++] = x->snow.pack_water; u[k++] = x->snow.density; u[k++] = x->snow.depth; u[k++] = x->snow.albedo;
  u[k++] = x->snow.coldcontent; u[k++] = x->snow.coverage; u[k++] = x->snow.snow_canopy;
  u[k++] = x->snow.tmp_int_storage; u[k++] = (double)x->snow.last_snow; u[k++] = (double)x->snow.MELTING;
  u[k++] = (double)x->snow.snow; u[k++] = x->snow.vapor_flux; u[k++] = x->snow.canopy_vapor_flux;
  u[k++] = x->snow.melt;
  u[k++] = x->energy.T[0]; u[k++] = x->energy.T[1]; u[k++] = x->energy.T[2];
  u[k++] = x->energy.grnd_flux; u[k++] = x->energy.deltaH; u[k++] = x->energy.fusion;
  u[k++] = x->energy.snow_flux; u[k++] = x->energy.LongUnderOut; u[k++] = x->energy.Tfoliage;
  u[k++] = x->energy.AlbedoUnder; u[k++] = x->energy.NetShortAtmos; u[k++] = x->energy.NetLongAtmos;
}

static int step_range(vo_model *m, const vr_forcing *f, double *out, double *units_dbg, int maxtile,
                      int64_t c0, int64_t c1)
{
  int64_t c, C = m->C, t;
  int rec, nrec = f->nrec, v, NB = m->opt.nbands, band, rc = 0;
  double o[VR_NOUTVAR], op, orn, osn;
  NLAYER = m->opt.nlayer;
  if (!m->initialised) return VR_ERR_STATE;
  for (c = c0; c < c1; c++) {
    for (rec = 0; rec < nrec; rec++) {
      /* the record's sub-steps [0..NF-1] and the record itself [NR] (vr_forcing.field: [nrec * NS][C]) */
      oatmos as[25];
      const int NF = m->opt.dt_hours / m->opt.snow_step_hours, NS = NF > 1 ? NF + 1 : 1;
      int j;
      for (j = 0; j < NS; j++) {
        oatmos *a = &as[j];
        size_t k = ((size_t)rec * NS + j) * C + c;
        a->air_temp = f->field[VR_F_AIR_TEMP][k]; a->prec = f->field[VR_F_PREC][k];
        a->wind = f->field[VR_F_WIND][k]; a->vp = f->field[VR_F_VP][k]; a->vpd = f->field[VR_F_VPD][k];
        a->pressure = f->field[VR_F_PRESSURE][k]; a->density = f->field[VR_F_DENSITY][k];
        a->shortwave = f->field[VR_F_SHORTWAVE][k]; a->longwave = f->field[VR_F_LONGWAVE][k];
        a->month = f->month[rec]; a->day_in_year = f->day_in_year[rec];
        a->snowflag = (NF > 1 && f->snowflag) ? f->snowflag[(size_t)rec * C + c] : 0;
      }
      if (full_energy(m, c, as, &op, &orn, &osn) == VERROR) { rc = VR_ERR_CELL; break; }
      put_data(m, c, &as[NS - 1], op, orn, osn, 0, o);
      for (v = 0; v < m->opt.n_out; v++) out[((size_t)v * nrec + rec) * C + c] = o[m->opt.out_vars[v]];
      if (units_dbg) {
        for (t = m->tile_ptr[c]; t < m->tile_ptr[c + 1]; t++) {
          int it = (int)(t - m->tile_ptr[c]);
          if (it >= maxtile) break;
          for (band = 0; band < NB; band++)
            if (m->tile_Cv[t] > 0 && m->soil[c].AreaFract[band] > 0)
              fill_unit_dbg(units_dbg + ((((size_t)rec * C + c) * maxtile + it) * NB + band) * VO_NUNITF,
                            &m->unit[t * NB + band]);
        }
      }
    }
  }
  return rc;
}

int vo_step_block(vo_model *m, const vr_forcing *f, double *out, double *units_dbg, int maxtile)
{
  return step_range(m, f, out, units_dbg, maxtile, 0, m->C);
}

int vo_step_block_range(vo_model *m, const vr_forcing *f, double *out, int64_t c0, int64_t c1)
{
  return step_range(m, f, out, NULL, 0, c0, c1);
}

void vo_fallback_counts(const vo_model *m, int64_t counts[2])
{
  int64_t i, n = m->T * m->opt.nbands;
  counts[0] = counts[1] = 0;
  for (i = 0; i < n; i++) {
    counts[0] += m->unit[i].snow.surf_temp_fbcount;
    counts[1] += m->unit[i].energy.Tfoliage_fbcount;
  }
}
