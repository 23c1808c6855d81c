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

// The term buffer is written once by k_units and read once by k_cells, 50 MB per record of a 100k-cell grid.  With
// -DVR_STREAM_HINTS its stores and loads carry the evict-first hint; measured 2.6 % SLOWER on the 100k-cell benchmark
// (the L2 no longer merges the scattered 8-byte stores into full sectors before they leave; profiles/r2_tuning.md),
// so the hint is off by default.
#if defined(__CUDA_ARCH__) && defined(VR_STREAM_HINTS)
#define VR_STREAM_ST(p, v) __stcs((p), (v))
#define VR_STREAM_LD(p) __ldcs(p)
#else
#define VR_STREAM_ST(p, v) (*(p) = (v))
#define VR_STREAM_LD(p) (*(p))
#endif

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
// mf_depth: options.MOISTFRACT -- the cell's layer depths (m): the layer moisture is stored as a volume fraction,
// moist / (depth * 1000), before the area weighting (put_data.c:906-911); null = in mm.
VR_HD void unit_terms(const UnitC &uc, double cv, double af, const UnitS &s, const UnitOut &o, int NL,
                       const TermMap &tmap, double *col, int stride, const UnitExtra *x = nullptr, const double *mf_depth = nullptr)
{
  const double AF = cv * af * 1. * 1.;
  const bool HasVeg = !uc.is_bare;
#define T_(k, expr) do { if (tmap.row[k] >= 0) VR_STREAM_ST(col + (size_t)tmap.row[k] * stride, (expr)); } while (0)
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
      T_(TE_LIQ0 + l, (mf_depth ? s.moist[l] / (mf_depth[l] * 1000.) : s.moist[l]) * AF);
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

// One unit, one record, and its terms: full_energy.c:210-470 for one tile x band pair -- the canopy terms of the
// month, the passes of surface_fluxes' sub-step loop (one, or NF = TIME_STEP / SNOW_STEP when the snow model is
// sub-stepped and the unit has snow or the record has snowfall: surface_fluxes.c:385-396), their sums and averages
// (:1030-1123), runoff() once per record (:1130-1150), and the unit's area-weighted terms.  The record is cut at the call
// of runoff() into unit_front() and unit_back(): the product runs them as separate kernels (vicres_capi.cu), the host
// harness of the tests one after the other (unit_record()).
//   forc(k)  -> Forc of sub-step k = 0..NF-1, or of the record itself for k == NF (k == 0 when NF == 1)
//   sio      -> the unit's snow words: load() / store(SnowS) / fallback(kind); only touched on the snow path
//   h, snowy -> the hot state and "some snow word is non-zero"
//   tidy     -> the stored MELTING / last_snow / albedo are not yet what the reference's no-snow branch leaves behind
//               (solve_snow.c:565-577); the first record without snow stores them (sio.store_none())
// Terms that do not depend on runoff() are stored by unit_front(), the others by unit_back().  SUB == false is the instance
// for SNOW_STEP == TIME_STEP (NF == 1): one pass, no sums or averages (a sum over one pass and a division by 1.0 are
// exact, so the results are the same; the instance only sheds ~25 live accumulators and the divisions).

// what unit_front() hands to unit_back(): the record's water reaching the soil, the layer evaporation and its bare-soil
// fractions, and the three vapour sums that enter OUT_EVAP after the layer evaporation (put_data.c:318-350)
struct FrontOut { double ppt, vapor, cvapor, canopyevap; M3 evap, bef; int err; };
enum { FO_PPT = 0, FO_VAPOR, FO_CVAPOR, FO_CANOPYEVAP, FO_EVAP0, FO_EVAP1, FO_EVAP2, FO_BEF0, FO_BEF1, FO_BEF2, FO_N };

// does this unit run the pass WITH the snow model in this record: it carries snow, the record has snowfall anywhere in
// the cell (sub-stepping follows, surface_fluxes.c:385-396) or in this band
VR_HD bool unit_needs_snow(const Forc &fr, double Tfactor, double Pfactor, int NF, double max_snow_temp, double min_rain_temp,
                           bool snowy)
{
  if (snowy || (NF > 1 && fr.snowflag)) return true;
  double rf, sf;
  rain_snow_split(fr.air_temp + Tfactor, fr.prec * Pfactor, max_snow_temp, min_rain_temp, rf, sf);
  return sf > 0.;
}

// GENERAL: the instance with the snow model (surface_pass<.., true>); the caller decides with unit_needs_snow() -- per
// unit in the split kernels (the units that need it are compacted into their own launch), per warp in a fused loop.
template <int extras, bool SUB, bool GENERAL, class ForcOf, class SnowIO>
VR_HD FrontOut unit_front(CellC cc, const UnitC &uc, CanopyM cm, const double *ae, ForcOf forc, int NF, int NL,
                          int dt_hours, int snow_step_hours, double max_snow_temp, double min_rain_temp, HotS &h, bool &snowy,
                          bool &tidy, SnowIO &sio, double cv, double af, const TermMap &tmap, double *col, int stride,
                          bool write, bool sync_ok, const double *tl = nullptr, const double *ap = nullptr)
{
  const bool is_veg = !uc.is_bare, overstory = uc.overstory != 0;
  constexpr bool general = GENERAL;
  VegP vp;
  vp.rarc = uc.rarc; vp.rmin = uc.rmin; vp.RGL = uc.RGL;
#pragma unroll
  for (int l = 0; l < MAXL; l++) vp.root[l] = uc.root[l];
  if (!is_veg) { cm.vegcover = 0; cm.LAI = 0; cm.Wdmax = 0; cm.albedo = BARE_SOIL_ALBEDO; cm.surf_atten = 1.; }
  const Forc fr = forc(NF > 1 ? NF : 0);
  SnowS sn = snow_none();
  if (general) sn = sio.load();
  else if (tidy) { sio.store_none(); tidy = false; }
  // full_energy.c:221-224: storages per unit of vegetated area
  if (is_veg) {
    h.Wdew /= cm.vegcover;
    if (general) sn.snow_canopy /= cm.vegcover;
  }
  const bool substep = SUB && general && NF > 1 && (sn.swq > 0 || sn.snow_canopy > 0 || fr.snowflag);     // surface_fluxes.c:385-396
  const int npass = substep ? NF : 1, pdt = substep ? snow_step_hours : dt_hours;
#define ACC_(a, v) do { if (SUB) a += (v); else a = (v); } while (0)
#define AVG_(x) (SUB ? (x) / N : (x))
  const double Tgrnd = h.T0;
  double coverage = sn.coverage, surf_atten = cm.surf_atten;
  // sums over the passes (surface_fluxes.c:935-1025)
  double st_ppt = 0, st_canopyevap = 0, st_throughfall = 0, st_vapor = 0, st_cvapor = 0, st_melt = 0, st_prec = 0,
         st_rain = 0, st_snow = 0, st_nshort = 0, st_nlong = 0, st_lover = 0, st_lunder = 0, st_aover = 0, st_aunder = 0,
         st_gflux = 0, st_dh = 0, st_cond0 = 0, st_cond1 = 0, st_luo = 0, Tsurf = 0;
  M3 st_evap, bef;
  P6 st_pot;
#pragma unroll
  for (int l = 0; l < MAXL; l++) { st_evap.v[l] = 0; bef.v[l] = 0; }
#pragma unroll
  for (int i = 0; i < 6; i++) st_pot.v[i] = 0;
  int snow_flag = 0, err = 0;
#pragma unroll 1
  for (int p = 0; p < npass; p++) {
    Forc f = fr;
    if (substep) {
      f = forc(p);
      f.wind = fr.wind;           // the aerodynamic resistances of all passes come from wind[NR] (full_energy.c:332)
    }
    const SurfOut r = surface_pass<extras, GENERAL>(cc, vp, is_veg, overstory, uc.Tfactor, uc.Pfactor, cm, ae, f, NL, pdt,
                                                    max_snow_temp, min_rain_temp, Tgrnd, h, sn, coverage, surf_atten, tl, ap,
                                                    sync_ok && NF == 1, dt_hours);
    if (is_veg) { ACC_(st_throughfall, r.throughfall); ACC_(st_canopyevap, r.canopyevap); ACC_(st_cvapor, r.canopy_vapor_flux); }
#pragma unroll
    for (int l = 0; l < MAXL; l++) {
      if (l < NL) ACC_(st_evap.v[l], r.evap.v[l]);
      bef.v[l] = r.bef.v[l];
    }
    ACC_(st_ppt, r.ppt);
    ACC_(st_cond0, (r.ra_used0 > 0) ? 1 / r.ra_used0 : HUGE_RESIST);
    ACC_(st_cond1, (r.ra_used1 > 0) ? 1 / r.ra_used1 : HUGE_RESIST);
    ACC_(st_melt, r.melt); ACC_(st_vapor, r.vapor_flux);
    ACC_(st_prec, r.out_prec); ACC_(st_rain, r.out_rain); ACC_(st_snow, r.out_snow);
    ACC_(st_aover, r.AlbedoOver); ACC_(st_aunder, r.AlbedoUnder); ACC_(st_lover, r.LongOverIn); ACC_(st_lunder, r.LongUnderIn);
    ACC_(st_nlong, r.NetLongAtmos); ACC_(st_nshort, r.NetShortAtmos); ACC_(st_dh, r.deltaH); ACC_(st_gflux, r.grnd_flux);
    ACC_(st_luo, h.LongUnderOut);
    if (extras & EX_PET) {
#pragma unroll
      for (int i = 0; i < 6; i++) ACC_(st_pot.v[i], r.pot.v[i]);
    }
    Tsurf = r.Tsurf;
    snow_flag = r.snow_flag;
    err |= r.err;
    if (r.fb_snow) sio.fallback(0);
    if (r.fb_fol) sio.fallback(1);
  }
  if (general) {
    sio.store(sn);
    snowy = sn.swq > 0 || sn.snow_canopy > 0 || sn.coverage != 0;
    tidy = !snowy && (sn.last_snow != (int)MISSINGV || sn.MELTING != 0 || sn.albedo != NEW_SNOW_ALB);
  }
  const double N = (double)npass;
  h.LongUnderOut = AVG_(st_luo);     // the record leaves the AVERAGE of its passes behind (surface_fluxes.c:1041), the other
                                    // carried words (T[0], T[1], Tfoliage, snow, canopy storage) the last pass's values
  // surface_fluxes.c:1110-1123: resistance -> conductance -> resistance over the passes
  const double ar0 = (st_cond0 > 0 && st_cond0 < HUGE_RESIST) ? 1 / AVG_(st_cond0) : ((st_cond0 >= HUGE_RESIST) ? 0 : HUGE_RESIST);
  const double ar1 = (st_cond1 > 0 && st_cond1 < HUGE_RESIST) ? 1 / AVG_(st_cond1) : ((st_cond1 >= HUGE_RESIST) ? 0 : HUGE_RESIST);
  const double AF = cv * af * 1. * 1.;
#define T_(k, expr) do { if (write && tmap.row[k] >= 0) VR_STREAM_ST(col + (size_t)tmap.row[k] * stride, (expr)); } while (0)
  // ---- terms that are final before runoff() (collect_wb_terms / collect_eb_terms, put_data.c:298-632) ----
  T_(TE_SUB_SNOW, st_vapor * 1000. * AF);
  if (is_veg) {
    T_(TE_SUB_CANOP, st_cvapor * 1000. * AF);
    T_(TE_EVAP_CANOP, st_canopyevap * AF);
    T_(TE_WDEW, h.Wdew * AF);
    T_(TE_SCANOPY, (sn.snow_canopy) * AF * 1000.);
  }
  else { T_(TE_SUB_CANOP, 0.); T_(TE_EVAP_CANOP, 0.); T_(TE_WDEW, 0.); T_(TE_SCANOPY, 0.); }
  T_(TE_INFLOW, (st_ppt) * AF);
  {
    const double ar = overstory ? ar1 : ar0;
    T_(TE_COND, (ar > SMALLV) ? (1 / ar) * AF : HUGE_RESIST);
  }
  T_(TE_SWE, sn.swq * AF * 1000.);
  T_(TE_SDEPTH, sn.depth * AF * 100.);
  if (sn.swq > 0.0) {
    T_(TE_SALB, sn.albedo * AF);
    T_(TE_SST, sn.surf_temp * AF);
    T_(TE_SPT, sn.pack_temp * AF);
    T_(TE_CVSNOW, cv * af * 1.);
  }
  else { T_(TE_SALB, 0.); T_(TE_SST, 0.); T_(TE_SPT, 0.); T_(TE_CVSNOW, 0.); }
  T_(TE_SMELT, st_melt * AF);
  T_(TE_SCOVER, sn.coverage * AF);
  T_(TE_SURFT, Tsurf * AF);
  T_(TE_NSHORT, AVG_(st_nshort) * AF);
  T_(TE_NLONG, AVG_(st_nlong) * AF);
  T_(TE_INLONG, ((snow_flag && overstory) ? AVG_(st_lover) : AVG_(st_lunder)) * AF);
  T_(TE_ALBEDO, ((snow_flag && overstory) ? AVG_(st_aover) : AVG_(st_aunder)) * AF);
  T_(TE_GFLUX, -(AVG_(st_gflux) * AF));     // collect_eb_terms subtracts these two
  T_(TE_DH, -(AVG_(st_dh) * AF));
  T_(TE_PREC, st_prec * cv * af);
  T_(TE_RAIN, st_rain * cv * af);
  T_(TE_SNOWF, st_snow * cv * af);
  if (extras & EX_PET) {
#pragma unroll
    for (int i = 0; i < 6; i++) T_(TE_PET0 + i, AVG_(st_pot.v[i]) * AF);    // put_data.c:846-851
  }
#undef T_
#undef ACC_
#undef AVG_
  (void)st_throughfall; (void)N;
  FrontOut o;
  o.ppt = st_ppt; o.vapor = st_vapor; o.cvapor = st_cvapor; o.canopyevap = st_canopyevap;
  o.evap = st_evap; o.bef = bef; o.err = err;
  return o;
}

// runoff() of the record (surface_fluxes.c:1130-1150), full_energy.c:445-453, and the terms that depend on them
// run(ppt, moist, evap) -> RunoffOut: runoff() with the loop constants where the caller keeps them
template <int extras, class RunoffFn>
VR_HD void unit_back_fn(CellC cc, bool is_veg, int root0_mask, int NL, FrontOut fo, M3 &moist, double cv, double af,
                        const TermMap &tmap, double *col, int stride, bool write, RunoffFn run, bool moistfract = false)
{
  const double AF = cv * af * 1. * 1.;
#define T_(k, expr) do { if (write && tmap.row[k] >= 0) VR_STREAM_ST(col + (size_t)tmap.row[k] * stride, (expr)); } while (0)
  // ---- runoff.c ----
  const RunoffOut ro = run(fo.ppt, moist, fo.evap);
  moist = ro.moist;
  M3 st_evap = fo.evap;
  if (NL == 3) st_evap.v[2] = ro.evap_bottom; else st_evap.v[1] = ro.evap_bottom;
  // ---- full_energy.c:445-453 ----
  double wet = 0, rootm = 0, tmp_evap = 0.0;
#pragma unroll
  for (int l = 0; l < MAXL; l++) {
    if (l < NL) {
      if ((root0_mask >> l) & 1) rootm += moist.v[l];
      wet += (moist.v[l] - cc.Wpwp(l)) / cc.wet_den(l);
      tmp_evap += st_evap.v[l];
      if (is_veg) {
        T_(TE_EB0 + l, st_evap.v[l] * fo.bef.v[l] * AF);
        T_(TE_TV0 + l, st_evap.v[l] * (1 - fo.bef.v[l]) * AF);
      }
      else {
        T_(TE_EB0 + l, st_evap.v[l] * AF);
        T_(TE_TV0 + l, 0.);
      }
      T_(TE_LIQ0 + l, (moistfract ? moist.v[l] / (cc.depth(l) * 1000.) : moist.v[l]) * AF);
    }
    else { T_(TE_EB0 + l, 0.); T_(TE_TV0 + l, 0.); T_(TE_LIQ0 + l, 0.); }
  }
  wet /= NL;
  tmp_evap += fo.vapor * 1000.;
  if (is_veg) {
    tmp_evap += fo.cvapor * 1000.;
    tmp_evap += fo.canopyevap;
  }
  T_(TE_EVAP, tmp_evap * AF);
  T_(TE_ASAT, ro.asat * AF);
  T_(TE_RUNOFF, ro.runoff * AF);
  T_(TE_BASEFLOW, ro.baseflow * AF);
  T_(TE_WET, wet * AF);
  T_(TE_ROOTM, rootm * AF);
  if (extras & EX_ZWT) {
    const ZwtOut z = water_table(cc, NL, moist);
    T_(TE_ZWT, z.zwt * AF);                                  // put_data.c:917-918
    T_(TE_ZWTL, z.lumped * AF);
  }
#undef T_
}

template <int extras>
VR_HD void unit_back(CellC cc, bool is_veg, int root0_mask, int NL, int dt_hours, FrontOut fo, M3 &moist, double cv, double af,
                     const TermMap &tmap, double *col, int stride, bool write, bool moistfract = false)
{
  unit_back_fn<extras>(cc, is_veg, root0_mask, NL, fo, moist, cv, af, tmap, col, stride, write,
                       [&](double ppt, M3 m, M3 e) { return runoff_step(cc, NL, dt_hours, ppt, m, e); }, moistfract);
}

// the two halves one after the other (host harness of the tests).  Returns 1 when the reference would have returned ERROR
// (arno_evap.c:112-151).
template <int extras, bool SUB, class ForcOf, class SnowIO>
VR_HD int unit_record(CellC cc, const UnitC &uc, const double *tm, const double *ae, ForcOf forc, int NF, int NL,
                      int dt_hours, int snow_step_hours, double max_snow_temp, double min_rain_temp, HotS &h, bool &snowy,
                      bool &tidy, SnowIO &sio, double cv, double af, const TermMap &tmap, double *col, int stride, bool write,
                      const double *tl = nullptr, const double *ap = nullptr, bool moistfract = false)
{
  const Forc fr = forc(NF > 1 ? NF : 0);
  const bool general = unit_needs_snow(fr, uc.Tfactor, uc.Pfactor, NF, max_snow_temp, min_rain_temp, snowy);
  CanopyM cm;      // the month's canopy terms of the tile (full_energy.c:210-227,306-316)
  cm.vegcover = tm[TM_VEGCOVER]; cm.LAI = tm[TM_LAI]; cm.Wdmax = tm[TM_WDMAX]; cm.albedo = tm[TM_ALBEDO];
  cm.surf_atten = tm[TM_SURF_ATTEN];
  FrontOut fo;
  if (general)
    fo = unit_front<extras, SUB, true>(cc, uc, cm, ae, forc, NF, NL, dt_hours, snow_step_hours, max_snow_temp, min_rain_temp, h,
                                       snowy, tidy, sio, cv, af, tmap, col, stride, write, false, tl, ap);
  else
    fo = unit_front<extras, false, false>(cc, uc, cm, ae, forc, NF, NL, dt_hours, snow_step_hours, max_snow_temp, min_rain_temp,
                                          h, snowy, tidy, sio, cv, af, tmap, col, stride, write, false, tl, ap);
  unit_back<extras>(cc, !uc.is_bare, uc.root0_mask, NL, dt_hours, fo, h.moist, cv, af, tmap, col, stride, write, moistfract);
  return fo.err;
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
    double a = 0. + VR_STREAM_LD(p);
    for (int j = 1; j < nu; j++) a += VR_STREAM_LD(p + j);
    dst[(size_t)r * dstride + dcol] = a;
  }
  if (tmap.row[TE_EB0] >= 0) {
    double eb = 0;
    for (int j = 0; j < nu; j++)
      for (int l = 0; l < NL; l++) eb += VR_STREAM_LD(src + (size_t)tmap.row[TE_EB0 + l] * sstride + u0 + j);
    dst[(size_t)tmap.row[TE_EB0] * dstride + dcol] = eb;
  }
  if (tmap.row[TE_TV0] >= 0) {
    double tv = 0;
    for (int j = 0; j < nu; j++)
      for (int l = 0; l < NL; l++) tv += VR_STREAM_LD(src + (size_t)tmap.row[TE_TV0 + l] * sstride + u0 + j);
    dst[(size_t)tmap.row[TE_TV0] * dstride + dcol] = tv;
  }
}

// ordered sum of one term over the cell's units (the save_data words of the PREVIOUS record are rebuilt
// from these when the records of a chunk are aggregated in parallel)
VR_HD double cell_term_sum(const double *src, size_t sstride, int64_t u0, int nu, int row)
{
  const double *p = src + (size_t)row * sstride + u0;
  double a = 0. + VR_STREAM_LD(p);
  for (int j = 1; j < nu; j++) a += VR_STREAM_LD(p + j);
  return a;
}

// ordered sum of term k over the cell's units, straight from the term buffer (rows of sstride doubles, the cell's units
// in columns u0 .. u0+nu-1).  The three per-layer EVAP_BARE / TRANSP_VEG terms of a unit are added layer by layer into
// ONE accumulator, exactly as collect_wb_terms does; they are asked for as TE_EB0 / TE_TV0.
VR_HD double cell_sum(const double *src, size_t sstride, int64_t u0, int nu, const TermMap &tmap, int NL, int k)
{
  if (k == TE_EB0 || k == TE_TV0) {
    double a = 0;
    for (int j = 0; j < nu; j++)
      for (int l = 0; l < NL; l++) a += VR_STREAM_LD(src + (size_t)tmap.row[k + l] * sstride + u0 + j);
    return a;
  }
  return cell_term_sum(src, sstride, u0, nu, tmap.row[k]);
}

// One output variable of a cell-record from the summed terms (put_data.c:540-632).  S(k) reads sum k.
// water_error / deltas use the carried save_data words sv[] BEFORE cell_carry() updates them.
// mfd: options.MOISTFRACT -- the cell's layer depths (m); the soil terms are volume fractions then, and the storage of the
// water-balance check turns them back into mm (put_data.c:624-630); null = the terms are mm
template <class Sum>
VR_HD double cell_output(int var, Sum S, const Forc &f, int NL, bool init, const double *sv, const double *mfd = nullptr)
{
  switch (var) {
    case VR_OUT_AIR_TEMP: return f.air_temp;
    case VR_OUT_LONGWAVE: return f.longwave;
    case VR_OUT_SHORTWAVE: return f.shortwave;
    case VR_OUT_WIND: return f.wind;
    case VR_OUT_REL_HUMID: return 100. * f.vp / (f.vp + f.vpd);
    case VR_OUT_DENSITY: return f.density;                       // put_data.c:275-291
    case VR_OUT_PRESSURE: return f.pressure / 1000.;
    case VR_OUT_VP: return f.vp / 1000.;
    case VR_OUT_VPD: return f.vpd / 1000.;
    case VR_OUT_QAIR: return 0.62196351 * f.vp / f.pressure;
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
      for (int l = 0; l < NL; l++) storage += mfd ? (S(TE_LIQ0 + l) + 0.) * mfd[l] * 1000 : S(TE_LIQ0 + l) + 0.;
      storage += S(TE_SWE) + S(TE_SCANOPY) + S(TE_WDEW) + 0.;
      return inflow - outflow - (storage - sv[SV_STORAGE]);
    }
    default: return 0.;
  }
}

// carry the save_data words to the next record (put_data.c:612-620, calc_water_balance_error)
template <class Sum>
VR_HD void cell_carry(Sum S, int NL, double *sv, const double *mfd = nullptr)
{
  double soil_tot = 0, storage = 0.;
  for (int l = 0; l < NL; l++) soil_tot += S(TE_LIQ0 + l) + 0.;
  for (int l = 0; l < NL; l++) storage += mfd ? (S(TE_LIQ0 + l) + 0.) * mfd[l] * 1000 : S(TE_LIQ0 + l) + 0.;
  storage += S(TE_SWE) + S(TE_SCANOPY) + S(TE_WDEW) + 0.;
  sv[SV_SOIL] = soil_tot;
  sv[SV_SWE] = S(TE_SWE) + S(TE_SCANOPY);
  sv[SV_WDEW] = S(TE_WDEW);
  sv[SV_STORAGE] = storage;
}

}  // namespace vr
