// vic_aggregate.cuh -- tile x band -> cell aggregation of the step's fluxes and states: the device
// restatement of put_data.c:298-632 (collect_wb_terms / collect_eb_terms, storage deltas and the
// water-balance closure error of calc_water_energy_balance_errors.c).
//
// Each unit thread turns its results into NTERM area-weighted terms (value * Cv * AreaFract, the
// reference's AreaFactor); the cell's leader thread then adds the terms of the cell's units IN
// REFERENCE ORDER (tiles ascending, bands ascending) so the fp64 sums round exactly as the serial
// reference does, finishes the derived outputs and keeps the per-cell save_data words.
#pragma once
#include "../../include/vicres.h"
#include "vic_physics.cuh"

namespace vr {

enum {
  TE_EB0 = 0, TE_EB1, TE_EB2, TE_TV0, TE_TV1, TE_TV2, TE_EVAP, TE_SUB_SNOW, TE_SUB_CANOP, TE_EVAP_CANOP,
  TE_ASAT, TE_RUNOFF, TE_BASEFLOW, TE_INFLOW, TE_WDEW, TE_COND, TE_LIQ0, TE_LIQ1, TE_LIQ2, TE_WET, TE_ROOTM,
  TE_SWE, TE_SDEPTH, TE_SALB, TE_SST, TE_SPT, TE_CVSNOW, TE_SCANOPY, TE_SMELT, TE_SCOVER, TE_SURFT,
  TE_NSHORT, TE_NLONG, TE_INLONG, TE_ALBEDO, TE_GFLUX, TE_DH, TE_PREC, TE_RAIN, TE_SNOWF, NTERM
};

// per-cell carried words of save_data_struct / calc_water_balance_error (put_data.c:586-632)
enum { SV_SOIL = 0, SV_SWE, SV_WDEW, SV_STORAGE, SV_N };

// terms[k * stride] for k in [0, NTERM)
VR_HDN void unit_terms(const UnitC &uc, double cv, double af, const UnitS &s, const UnitOut &o, int NL,
                      double *terms, int stride)
{
  const double AF = cv * af * 1. * 1.;
  const bool HasVeg = !uc.is_bare;
#define T_(k) terms[(k) * stride]
  double tmp_evap = 0.0;
#pragma unroll
  for (int l = 0; l < MAXL; l++) {
    if (l < NL) {
      tmp_evap += o.evap[l];
      if (HasVeg) {
        T_(TE_EB0 + l) = o.evap[l] * o.bef[l] * AF;
        T_(TE_TV0 + l) = o.evap[l] * (1 - o.bef[l]) * AF;
      }
      else {
        T_(TE_EB0 + l) = o.evap[l] * AF;
        T_(TE_TV0 + l) = 0.;
      }
      T_(TE_LIQ0 + l) = s.moist[l] * AF;
    }
    else { T_(TE_EB0 + l) = 0.; T_(TE_TV0 + l) = 0.; T_(TE_LIQ0 + l) = 0.; }
  }
  tmp_evap += o.vapor_flux * 1000.;
  T_(TE_SUB_SNOW) = o.vapor_flux * 1000. * AF;
  if (HasVeg) {
    tmp_evap += o.canopy_vapor_flux * 1000.;
    T_(TE_SUB_CANOP) = o.canopy_vapor_flux * 1000. * AF;
    tmp_evap += o.canopyevap;
    T_(TE_EVAP_CANOP) = o.canopyevap * AF;
    T_(TE_WDEW) = s.Wdew * AF;
    T_(TE_SCANOPY) = (s.snow_canopy) * AF * 1000.;
  }
  else { T_(TE_SUB_CANOP) = 0.; T_(TE_EVAP_CANOP) = 0.; T_(TE_WDEW) = 0.; T_(TE_SCANOPY) = 0.; }
  T_(TE_EVAP) = tmp_evap * AF;
  T_(TE_ASAT) = o.asat * AF;
  T_(TE_RUNOFF) = o.runoff * AF;
  T_(TE_BASEFLOW) = o.baseflow * AF;
  T_(TE_INFLOW) = (o.inflow) * AF;
  {
    const double ar = uc.overstory ? o.aero_resist[1] : o.aero_resist[0];
    T_(TE_COND) = (ar > SMALLV) ? (1 / ar) * AF : HUGE_RESIST;
  }
  T_(TE_WET) = o.wetness * AF;
  T_(TE_ROOTM) = o.rootmoist * AF;
  T_(TE_SWE) = s.swq * AF * 1000.;
  T_(TE_SDEPTH) = s.depth * AF * 100.;
  if (s.swq > 0.0) {
    T_(TE_SALB) = s.albedo * AF;
    T_(TE_SST) = s.surf_temp * AF;
    T_(TE_SPT) = s.pack_temp * AF;
    T_(TE_CVSNOW) = cv * af * 1.;
  }
  else { T_(TE_SALB) = 0.; T_(TE_SST) = 0.; T_(TE_SPT) = 0.; T_(TE_CVSNOW) = 0.; }
  T_(TE_SMELT) = o.melt * AF;
  T_(TE_SCOVER) = s.coverage * AF;
  T_(TE_SURFT) = o.Tsurf * AF;
  T_(TE_NSHORT) = o.NetShortAtmos * AF;
  T_(TE_NLONG) = o.NetLongAtmos * AF;
  T_(TE_INLONG) = o.in_long * AF;
  T_(TE_ALBEDO) = o.albedo * AF;
  T_(TE_GFLUX) = o.grnd_flux * AF;
  T_(TE_DH) = o.deltaH * AF;
  T_(TE_PREC) = o.out_prec * cv * af;
  T_(TE_RAIN) = o.out_rain * cv * af;
  T_(TE_SNOWF) = o.out_snow * cv * af;
#undef T_
}

// acc[NTERM] zero-initialised; add one unit's terms (same rounding sequence as the reference loop)
VR_HD void cell_accumulate(double *acc, const double *terms, int stride)
{
#pragma unroll
  for (int k = 0; k < NTERM; k++) {
    if (k == TE_GFLUX || k == TE_DH) acc[k] -= terms[k * stride];
    else acc[k] += terms[k * stride];
  }
}

// the per-unit order of the three per-layer EVAP_BARE / TRANSP_VEG adds is layer-major inside a unit
// (collect_wb_terms), which cell_accumulate cannot express with one slot per layer -> fold them here
VR_HD void cell_accumulate_ordered(double *acc, double *eb, double *tv, const double *terms, int stride, int NL)
{
  for (int l = 0; l < NL; l++) {
    *eb += terms[(TE_EB0 + l) * stride];
    *tv += terms[(TE_TV0 + l) * stride];
  }
  cell_accumulate(acc, terms, stride);
}

// finish one cell-record: o[VR_NOUTVAR]
VR_HDN void cell_finish(const double *acc, double eb, double tv, const Forc &f, int NL, bool init, double *sv, double *o)
{
#pragma unroll
  for (int k = 0; k < VR_NOUTVAR; k++) o[k] = 0;
  o[VR_OUT_AIR_TEMP] = f.air_temp;
  o[VR_OUT_LONGWAVE] = f.longwave;
  o[VR_OUT_PREC] = acc[TE_PREC];
  o[VR_OUT_RAINF] = acc[TE_RAIN];
  o[VR_OUT_REL_HUMID] = 100. * f.vp / (f.vp + f.vpd);
  o[VR_OUT_SHORTWAVE] = f.shortwave;
  o[VR_OUT_SNOWF] = acc[TE_SNOWF];
  o[VR_OUT_WIND] = f.wind;
  o[VR_OUT_EVAP_BARE] = eb;
  o[VR_OUT_TRANSP_VEG] = tv;
  o[VR_OUT_EVAP] = acc[TE_EVAP];
  o[VR_OUT_SUB_SNOW] = acc[TE_SUB_SNOW];
  o[VR_OUT_SUB_CANOP] = acc[TE_SUB_CANOP];
  o[VR_OUT_EVAP_CANOP] = acc[TE_EVAP_CANOP];
  o[VR_OUT_ASAT] = acc[TE_ASAT];
  o[VR_OUT_RUNOFF] = acc[TE_RUNOFF];
  o[VR_OUT_BASEFLOW] = acc[TE_BASEFLOW];
  o[VR_OUT_INFLOW] = acc[TE_INFLOW];
  o[VR_OUT_WDEW] = acc[TE_WDEW];
  o[VR_OUT_SOIL_LIQ0] = acc[TE_LIQ0];
  o[VR_OUT_SOIL_LIQ1] = acc[TE_LIQ1];
  o[VR_OUT_SOIL_LIQ2] = acc[TE_LIQ2];
  o[VR_OUT_SOIL_WET] = acc[TE_WET];
  o[VR_OUT_ROOTMOIST] = acc[TE_ROOTM];
  o[VR_OUT_SWE] = acc[TE_SWE];
  o[VR_OUT_SNOW_DEPTH] = acc[TE_SDEPTH];
  o[VR_OUT_SNOW_CANOPY] = acc[TE_SCANOPY];
  o[VR_OUT_SNOW_MELT] = acc[TE_SMELT];
  o[VR_OUT_SNOW_COVER] = acc[TE_SCOVER];
  o[VR_OUT_SURF_TEMP] = acc[TE_SURFT];
  o[VR_OUT_NET_SHORT] = acc[TE_NSHORT];
  o[VR_OUT_NET_LONG] = acc[TE_NLONG];
  o[VR_OUT_IN_LONG] = acc[TE_INLONG];
  o[VR_OUT_ALBEDO] = acc[TE_ALBEDO];
  o[VR_OUT_GRND_FLUX] = acc[TE_GFLUX];
  o[VR_OUT_DELTAH] = acc[TE_DH];
  const double cv_snow = acc[TE_CVSNOW];
  o[VR_OUT_SALBEDO] = acc[TE_SALB];
  o[VR_OUT_SNOW_SURF_TEMP] = acc[TE_SST];
  o[VR_OUT_SNOW_PACK_TEMP] = acc[TE_SPT];
  if (cv_snow > 0) {
    o[VR_OUT_SALBEDO] /= cv_snow;
    o[VR_OUT_SNOW_SURF_TEMP] /= cv_snow;
    o[VR_OUT_SNOW_PACK_TEMP] /= cv_snow;
  }
  o[VR_OUT_AERO_COND] = acc[TE_COND];
  o[VR_OUT_AERO_RESIST] = (acc[TE_COND] > SMALLV) ? 1 / acc[TE_COND] : HUGE_RESIST;
  double soil_tot = 0, dsm = 0;
  for (int l = 0; l < NL; l++) {
    double sm = o[VR_OUT_SOIL_LIQ0 + l] + 0.;
    dsm += sm;
    soil_tot += sm;
  }
  o[VR_OUT_DELSOILMOIST] = dsm;
  if (!init) {
    o[VR_OUT_DELSOILMOIST] -= sv[SV_SOIL];
    o[VR_OUT_DELSWE] = o[VR_OUT_SWE] + o[VR_OUT_SNOW_CANOPY] - sv[SV_SWE];
    o[VR_OUT_DELINTERCEPT] = o[VR_OUT_WDEW] - sv[SV_WDEW];
  }
  o[VR_OUT_R_NET] = o[VR_OUT_NET_SHORT] + o[VR_OUT_NET_LONG];
  sv[SV_SOIL] = soil_tot;
  sv[SV_SWE] = o[VR_OUT_SWE] + o[VR_OUT_SNOW_CANOPY];
  sv[SV_WDEW] = o[VR_OUT_WDEW];
  const double inflow = o[VR_OUT_PREC] + 0.;
  const double outflow = o[VR_OUT_EVAP] + o[VR_OUT_RUNOFF] + o[VR_OUT_BASEFLOW];
  double storage = 0.;
  for (int l = 0; l < NL; l++) storage += o[VR_OUT_SOIL_LIQ0 + l] + 0.;
  storage += o[VR_OUT_SWE] + o[VR_OUT_SNOW_CANOPY] + o[VR_OUT_WDEW] + 0.;
  if (init) o[VR_OUT_WATER_ERROR] = 0.0;
  else o[VR_OUT_WATER_ERROR] = inflow - outflow - (storage - sv[SV_STORAGE]);
  sv[SV_STORAGE] = storage;
}

}  // namespace vr
