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
  TE_NSHORT, TE_NLONG, TE_INLONG, TE_ALBEDO, TE_GFLUX, TE_DH, TE_PREC, TE_RAIN, TE_SNOWF,
  TE_PET0, TE_PET1, TE_PET2, TE_PET3, TE_PET4, TE_PET5, TE_ZWT, TE_ZWTL, NTERM
};

// per-cell carried words of save_data_struct / calc_water_balance_error (put_data.c:586-632)
enum { SV_SOIL = 0, SV_SWE, SV_WDEW, SV_STORAGE, SV_N };

// Which terms a set of requested outputs needs (the others are neither stored nor summed), and their
// rows in the shared-memory term buffer.  LIQ*, SWE, SCANOPY and WDEW are always kept: they feed
// the save_data words (storage deltas, closure error) that are carried from record to record.
struct TermMap {
  signed char row[NTERM];   // row in the term buffer, or -1
  int n;                    // rows in use
};

inline void build_term_map(const int *out_vars, int n_out, TermMap &tmap)
{
  bool need[NTERM];
  for (int k = 0; k < NTERM; k++) need[k] = false;
  auto N = [&](int k) { need[k] = true; };
  for (int k : {TE_LIQ0, TE_LIQ1, TE_LIQ2, TE_SWE, TE_SCANOPY, TE_WDEW}) N(k);
  for (int i = 0; i < n_out; i++) {
    switch (out_vars[i]) {
      case VR_OUT_PREC: N(TE_PREC); break;
      case VR_OUT_EVAP: N(TE_EVAP); break;
      case VR_OUT_RUNOFF: N(TE_RUNOFF); break;
      case VR_OUT_BASEFLOW: N(TE_BASEFLOW); break;
      case VR_OUT_NET_SHORT: N(TE_NSHORT); break;
      case VR_OUT_R_NET: N(TE_NSHORT); N(TE_NLONG); break;
      case VR_OUT_EVAP_CANOP: N(TE_EVAP_CANOP); break;
      case VR_OUT_TRANSP_VEG: N(TE_TV0); N(TE_TV1); N(TE_TV2); break;
      case VR_OUT_EVAP_BARE: N(TE_EB0); N(TE_EB1); N(TE_EB2); break;
      case VR_OUT_SUB_CANOP: N(TE_SUB_CANOP); break;
      case VR_OUT_SUB_SNOW: N(TE_SUB_SNOW); break;
      case VR_OUT_AERO_RESIST: case VR_OUT_AERO_COND: N(TE_COND); break;
      case VR_OUT_SURF_TEMP: N(TE_SURFT); break;
      case VR_OUT_ALBEDO: N(TE_ALBEDO); break;
      case VR_OUT_IN_LONG: N(TE_INLONG); break;
      case VR_OUT_SNOW_DEPTH: N(TE_SDEPTH); break;
      case VR_OUT_SNOW_COVER: N(TE_SCOVER); break;
      case VR_OUT_WATER_ERROR: N(TE_PREC); N(TE_EVAP); N(TE_RUNOFF); N(TE_BASEFLOW); break;
      case VR_OUT_INFLOW: N(TE_INFLOW); break;
      case VR_OUT_SNOW_MELT: N(TE_SMELT); break;
      case VR_OUT_RAINF: N(TE_RAIN); break;
      case VR_OUT_SNOWF: N(TE_SNOWF); break;
      case VR_OUT_NET_LONG: N(TE_NLONG); break;
      case VR_OUT_ASAT: N(TE_ASAT); break;
      case VR_OUT_SOIL_WET: N(TE_WET); break;
      case VR_OUT_ROOTMOIST: N(TE_ROOTM); break;
      case VR_OUT_GRND_FLUX: N(TE_GFLUX); break;
      case VR_OUT_DELTAH: N(TE_DH); break;
      case VR_OUT_SNOW_SURF_TEMP: N(TE_SST); N(TE_CVSNOW); break;
      case VR_OUT_SNOW_PACK_TEMP: N(TE_SPT); N(TE_CVSNOW); break;
      case VR_OUT_SALBEDO: N(TE_SALB); N(TE_CVSNOW); break;
      case VR_OUT_PET_SATSOIL: case VR_OUT_PET_H2OSURF: case VR_OUT_PET_SHORT: case VR_OUT_PET_TALL:
      case VR_OUT_PET_NATVEG: case VR_OUT_PET_VEGNOCR: N(TE_PET0 + (out_vars[i] - VR_OUT_PET_SATSOIL)); break;
      case VR_OUT_ZWT: N(TE_ZWT); break;
      case VR_OUT_ZWT_LUMPED: N(TE_ZWTL); break;
      default: break;
    }
  }
  tmap.n = 0;
  for (int k = 0; k < NTERM; k++) tmap.row[k] = need[k] ? (signed char)(tmap.n++) : (signed char)-1;
}

// One unit's area-weighted terms -> term buffer column `col` (rows of `stride` doubles).
template <int extras = EX_PET | EX_ZWT>
VR_HD void unit_terms(const UnitC &uc, double cv, double af, const UnitS &s, const UnitOut &o, int NL,
                       const TermMap &tmap, double *col, int stride, const UnitExtra *x = nullptr)
{
  const double AF = cv * af * 1. * 1.;
  const bool HasVeg = !uc.is_bare;
#define T_(k, expr) do { if (tmap.row[k] >= 0) col[(size_t)tmap.row[k] * stride] = (expr); } while (0)
  double tmp_evap = 0.0;
#pragma unroll
  for (int l = 0; l < MAXL; l++) {
    if (l < NL) {
      tmp_evap += o.evap[l];
      if (HasVeg) {
        T_(TE_EB0 + l, o.evap[l] * o.bef[l] * AF);
        T_(TE_TV0 + l, o.evap[l] * (1 - o.bef[l]) * AF);
      }
      else {
        T_(TE_EB0 + l, o.evap[l] * AF);
        T_(TE_TV0 + l, 0.);
      }
      T_(TE_LIQ0 + l, s.moist[l] * AF);
    }
    else { T_(TE_EB0 + l, 0.); T_(TE_TV0 + l, 0.); T_(TE_LIQ0 + l, 0.); }
  }
  tmp_evap += o.vapor_flux * 1000.;
  T_(TE_SUB_SNOW, o.vapor_flux * 1000. * AF);
  if (HasVeg) {
    tmp_evap += o.canopy_vapor_flux * 1000.;
    T_(TE_SUB_CANOP, o.canopy_vapor_flux * 1000. * AF);
    tmp_evap += o.canopyevap;
    T_(TE_EVAP_CANOP, o.canopyevap * AF);
    T_(TE_WDEW, s.Wdew * AF);
    T_(TE_SCANOPY, (s.snow_canopy) * AF * 1000.);
  }
  else { T_(TE_SUB_CANOP, 0.); T_(TE_EVAP_CANOP, 0.); T_(TE_WDEW, 0.); T_(TE_SCANOPY, 0.); }
  T_(TE_EVAP, tmp_evap * AF);
  T_(TE_ASAT, o.asat * AF);
  T_(TE_RUNOFF, o.runoff * AF);
  T_(TE_BASEFLOW, o.baseflow * AF);
  T_(TE_INFLOW, (o.inflow) * AF);
  {
    const double ar = uc.overstory ? o.aero_resist[1] : o.aero_resist[0];
    T_(TE_COND, (ar > SMALLV) ? (1 / ar) * AF : HUGE_RESIST);
  }
  T_(TE_WET, o.wetness * AF);
  T_(TE_ROOTM, o.rootmoist * AF);
  T_(TE_SWE, s.swq * AF * 1000.);
  T_(TE_SDEPTH, s.depth * AF * 100.);
  if (s.swq > 0.0) {
    T_(TE_SALB, s.albedo * AF);
    T_(TE_SST, s.surf_temp * AF);
    T_(TE_SPT, s.pack_temp * AF);
    T_(TE_CVSNOW, cv * af * 1.);
  }
  else { T_(TE_SALB, 0.); T_(TE_SST, 0.); T_(TE_SPT, 0.); T_(TE_CVSNOW, 0.); }
  T_(TE_SMELT, o.melt * AF);
  T_(TE_SCOVER, s.coverage * AF);
  T_(TE_SURFT, o.Tsurf * AF);
  T_(TE_NSHORT, o.NetShortAtmos * AF);
  T_(TE_NLONG, o.NetLongAtmos * AF);
  T_(TE_INLONG, o.in_long * AF);
  T_(TE_ALBEDO, o.albedo * AF);
  T_(TE_GFLUX, -(o.grnd_flux * AF));     // collect_eb_terms subtracts these two
  T_(TE_DH, -(o.deltaH * AF));
  T_(TE_PREC, o.out_prec * cv * af);
  T_(TE_RAIN, o.out_rain * cv * af);
  T_(TE_SNOWF, o.out_snow * cv * af);
  if (extras & EX_PET) {
#pragma unroll
    for (int i = 0; i < 6; i++) T_(TE_PET0 + i, (x ? x->pot_evap[i] : 0.) * AF);    // put_data.c:846-851
  }
  if (extras & EX_ZWT) {
    T_(TE_ZWT, (x ? x->zwt : 0.) * AF);                                  // put_data.c:917-918
    T_(TE_ZWTL, (x ? x->zwt_lumped : 0.) * AF);
  }
#undef T_
}

// One unit, one record, and its terms: the single out-of-line body the kernel calls per record.  Keeping
// unit_step() and unit_terms() in one function lets the ~30 per-unit results travel in registers
// instead of a UnitOut struct in local memory.
template <int extras>
VR_HDN void unit_step_terms(const CellC &cc, const UnitC &uc, const double *tm, const double *ae, const Forc &f, int NL,
                            int dt_hours, double max_snow_temp, double min_rain_temp, UnitS &s, double cv, double af,
                            const TermMap &tmap, double *col, int stride, bool write, const double *tl = nullptr,
                            const double *ap = nullptr)
{
  UnitOut uo;
  UnitExtra ux;
  unit_step<extras>(cc, uc, tm, ae, f, NL, dt_hours, max_snow_temp, min_rain_temp, s, uo, ux, tl, ap);
  if (write) unit_terms<extras>(uc, cv, af, s, uo, NL, tmap, col, stride, &ux);
}

// which optional results the requested outputs need
inline int term_extras(const TermMap &tmap)
{
  int extras = 0;
  for (int i = 0; i < 6; i++) if (tmap.row[TE_PET0 + i] >= 0) extras |= EX_PET;
  if (tmap.row[TE_ZWT] >= 0 || tmap.row[TE_ZWTL] >= 0) extras |= EX_ZWT;
  return extras;
}

// Sum the terms of one cell's units IN REFERENCE ORDER (tiles ascending, bands ascending) and leave
// the sums in the column of the cell's first unit.  cols[j] = buffer column of the cell's j-th unit.
// The three per-layer EVAP_BARE / TRANSP_VEG terms of a unit are added layer by layer into ONE
// accumulator each, exactly as collect_wb_terms does; the sums land in rows EB0 / TV0.
template <class ColOf>
VR_HD void cell_accumulate_inplace(double *buf, int stride, const TermMap &tmap, int nu, ColOf col_of, int NL)
{
  const int c0 = col_of(0);
#pragma unroll 1
  for (int k = 0; k < NTERM; k++) {
    const int r = tmap.row[k];
    if (r < 0 || (k >= TE_EB0 && k <= TE_TV2)) continue;
    double a = 0. + buf[(size_t)r * stride + c0];
    for (int j = 1; j < nu; j++) a += buf[(size_t)r * stride + col_of(j)];
    buf[(size_t)r * stride + c0] = a;
  }
  if (tmap.row[TE_EB0] >= 0) {
    double eb = 0;
    for (int j = 0; j < nu; j++)
      for (int l = 0; l < NL; l++) eb += buf[(size_t)tmap.row[TE_EB0 + l] * stride + col_of(j)];
    buf[(size_t)tmap.row[TE_EB0] * stride + c0] = eb;
  }
  if (tmap.row[TE_TV0] >= 0) {
    double tv = 0;
    for (int j = 0; j < nu; j++)
      for (int l = 0; l < NL; l++) tv += buf[(size_t)tmap.row[TE_TV0 + l] * stride + col_of(j)];
    buf[(size_t)tmap.row[TE_TV0] * stride + c0] = tv;
  }
}

// The same ordered sums, from a read-only term buffer whose columns u0 .. u0+nu-1 are the cell's units in
// reference order, into column `dcol` of a separate sum buffer (rows of `dstride` doubles).
VR_HD void cell_accumulate_to(const double *src, size_t sstride, int64_t u0, int nu, const TermMap &tmap, int NL,
                              double *dst, size_t dstride, int64_t dcol)
{
#pragma unroll 1
  for (int k = 0; k < NTERM; k++) {
    const int r = tmap.row[k];
    if (r < 0 || (k >= TE_EB0 && k <= TE_TV2)) continue;
    const double *p = src + (size_t)r * sstride + u0;
    double a = 0. + p[0];
    for (int j = 1; j < nu; j++) a += p[j];
    dst[(size_t)r * dstride + dcol] = a;
  }
  if (tmap.row[TE_EB0] >= 0) {
    double eb = 0;
    for (int j = 0; j < nu; j++)
      for (int l = 0; l < NL; l++) eb += src[(size_t)tmap.row[TE_EB0 + l] * sstride + u0 + j];
    dst[(size_t)tmap.row[TE_EB0] * dstride + dcol] = eb;
  }
  if (tmap.row[TE_TV0] >= 0) {
    double tv = 0;
    for (int j = 0; j < nu; j++)
      for (int l = 0; l < NL; l++) tv += src[(size_t)tmap.row[TE_TV0 + l] * sstride + u0 + j];
    dst[(size_t)tmap.row[TE_TV0] * dstride + dcol] = tv;
  }
}

// ordered sum of one term over the cell's units (the save_data words of the PREVIOUS record are rebuilt
// from these when the records of a chunk are aggregated in parallel)
VR_HD double cell_term_sum(const double *src, size_t sstride, int64_t u0, int nu, int row)
{
  const double *p = src + (size_t)row * sstride + u0;
  double a = 0. + p[0];
  for (int j = 1; j < nu; j++) a += p[j];
  return a;
}

// One output variable of a cell-record from the summed terms (put_data.c:540-632).  S(k) reads sum k.
// water_error / deltas use the carried save_data words sv[] BEFORE cell_carry() updates them.
template <class Sum>
VR_HD double cell_output(int var, Sum S, const Forc &f, int NL, bool init, const double *sv)
{
  switch (var) {
    case VR_OUT_AIR_TEMP: return f.air_temp;
    case VR_OUT_LONGWAVE: return f.longwave;
    case VR_OUT_SHORTWAVE: return f.shortwave;
    case VR_OUT_WIND: return f.wind;
    case VR_OUT_REL_HUMID: return 100. * f.vp / (f.vp + f.vpd);
    case VR_OUT_PREC: return S(TE_PREC);
    case VR_OUT_RAINF: return S(TE_RAIN);
    case VR_OUT_SNOWF: return S(TE_SNOWF);
    case VR_OUT_EVAP_BARE: return S(TE_EB0);
    case VR_OUT_TRANSP_VEG: return S(TE_TV0);
    case VR_OUT_EVAP: return S(TE_EVAP);
    case VR_OUT_SUB_SNOW: return S(TE_SUB_SNOW);
    case VR_OUT_SUB_CANOP: return S(TE_SUB_CANOP);
    case VR_OUT_EVAP_CANOP: return S(TE_EVAP_CANOP);
    case VR_OUT_ASAT: return S(TE_ASAT);
    case VR_OUT_RUNOFF: return S(TE_RUNOFF);
    case VR_OUT_BASEFLOW: return S(TE_BASEFLOW);
    case VR_OUT_INFLOW: return S(TE_INFLOW);
    case VR_OUT_WDEW: return S(TE_WDEW);
    case VR_OUT_SOIL_LIQ0: return S(TE_LIQ0);
    case VR_OUT_SOIL_LIQ1: return S(TE_LIQ1);
    case VR_OUT_SOIL_LIQ2: return (NL > 2) ? S(TE_LIQ2) : 0.;
    case VR_OUT_SOIL_WET: return S(TE_WET);
    case VR_OUT_ROOTMOIST: return S(TE_ROOTM);
    case VR_OUT_SWE: return S(TE_SWE);
    case VR_OUT_SNOW_DEPTH: return S(TE_SDEPTH);
    case VR_OUT_SNOW_CANOPY: return S(TE_SCANOPY);
    case VR_OUT_SNOW_MELT: return S(TE_SMELT);
    case VR_OUT_SNOW_COVER: return S(TE_SCOVER);
    case VR_OUT_SURF_TEMP: return S(TE_SURFT);
    case VR_OUT_NET_SHORT: return S(TE_NSHORT);
    case VR_OUT_NET_LONG: return S(TE_NLONG);
    case VR_OUT_R_NET: return S(TE_NSHORT) + S(TE_NLONG);
    case VR_OUT_IN_LONG: return S(TE_INLONG);
    case VR_OUT_ALBEDO: return S(TE_ALBEDO);
    case VR_OUT_GRND_FLUX: return S(TE_GFLUX);
    case VR_OUT_DELTAH: return S(TE_DH);
    case VR_OUT_SALBEDO: { double c = S(TE_CVSNOW), v = S(TE_SALB); return (c > 0) ? v / c : v; }
    case VR_OUT_SNOW_SURF_TEMP: { double c = S(TE_CVSNOW), v = S(TE_SST); return (c > 0) ? v / c : v; }
    case VR_OUT_SNOW_PACK_TEMP: { double c = S(TE_CVSNOW), v = S(TE_SPT); return (c > 0) ? v / c : v; }
    case VR_OUT_PET_SATSOIL: return S(TE_PET0);
    case VR_OUT_PET_H2OSURF: return S(TE_PET1);
    case VR_OUT_PET_SHORT: return S(TE_PET2);
    case VR_OUT_PET_TALL: return S(TE_PET3);
    case VR_OUT_PET_NATVEG: return S(TE_PET4);
    case VR_OUT_PET_VEGNOCR: return S(TE_PET5);
    case VR_OUT_ZWT: return S(TE_ZWT);
    case VR_OUT_ZWT_LUMPED: return S(TE_ZWTL);
    case VR_OUT_AERO_COND: return S(TE_COND);
    case VR_OUT_AERO_RESIST: return (S(TE_COND) > SMALLV) ? 1 / S(TE_COND) : HUGE_RESIST;
    case VR_OUT_DELSOILMOIST: {
      double d = 0;
      for (int l = 0; l < NL; l++) d += S(TE_LIQ0 + l) + 0.;
      return init ? d : d - sv[SV_SOIL];
    }
    case VR_OUT_DELSWE: return init ? 0. : S(TE_SWE) + S(TE_SCANOPY) - sv[SV_SWE];
    case VR_OUT_DELINTERCEPT: return init ? 0. : S(TE_WDEW) - sv[SV_WDEW];
    case VR_OUT_WATER_ERROR: {
      if (init) return 0.0;
      const double inflow = S(TE_PREC) + 0.;
      const double outflow = S(TE_EVAP) + S(TE_RUNOFF) + S(TE_BASEFLOW);
      double storage = 0.;
      for (int l = 0; l < NL; l++) storage += S(TE_LIQ0 + l) + 0.;
      storage += S(TE_SWE) + S(TE_SCANOPY) + S(TE_WDEW) + 0.;
      return inflow - outflow - (storage - sv[SV_STORAGE]);
    }
    default: return 0.;
  }
}

// carry the save_data words to the next record (put_data.c:612-620, calc_water_balance_error)
template <class Sum>
VR_HD void cell_carry(Sum S, int NL, double *sv)
{
  double soil_tot = 0, storage = 0.;
  for (int l = 0; l < NL; l++) soil_tot += S(TE_LIQ0 + l) + 0.;
  for (int l = 0; l < NL; l++) storage += S(TE_LIQ0 + l) + 0.;
  storage += S(TE_SWE) + S(TE_SCANOPY) + S(TE_WDEW) + 0.;
  sv[SV_SOIL] = soil_tot;
  sv[SV_SWE] = S(TE_SWE) + S(TE_SCANOPY);
  sv[SV_WDEW] = S(TE_WDEW);
  sv[SV_STORAGE] = storage;
}

}  // namespace vr
