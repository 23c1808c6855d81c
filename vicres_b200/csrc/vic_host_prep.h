// vic_host_prep.h -- host-side preparation for the B200 VIC-Res step: turns the vr_cells / vr_veglib
// hand-off (include/vicres.h) into the device layout
//   * per-cell constant block  cc[k][C]          (vr::CC_* rows; hoisted parameter arithmetic)
//   * active units (tile x band with Cv > 0 and AreaFract > 0) in reference order (tiles ascending,
//     bands ascending: full_energy.c:233,409), CSR over cells, SoA per-unit constants
//   * per (tile, month) canopy terms tm[T][12][TM_N]
//   * per (aero key, month) CalcAerodynamic coefficients ae[K][12][AE_N], deduplicated over
//     (class, soil roughness, snow roughness)
//   * thread slots: the units sorted by kind (overstory, other vegetation, bare soil) and, inside a kind,
//     by the climate of their cell and band, so that the lanes of a warp run the same branches; every
//     per-unit array is stored in SLOT order (thread k reads element k), and u_slot maps a unit in
//     reference order to its slot for the per-cell aggregation.
// All arithmetic here follows the reference expression order (read_soilparam-derived values are
// taken as given) so that it matches the oracle bit for bit on the same libm.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "../../include/vicres.h"
#include "vic_physics.cuh"

namespace vr {

struct Prepared {
  int NL = 0, NB = 0, cta_threads = 128;
  int ccn = 0;                       // rows of cc (VR_CC_N + the optional zwtvmoist tail)
  int64_t C = 0, T = 0, U = 0, K = 0, nCTA = 0;
  std::vector<double> cc;            // [CC_N][C]
  std::vector<int64_t> order;        // [C] sorted position -> caller's cell index (climate-sorted, see prepare())
  std::vector<int64_t> cell_uptr;    // [C+1] over SORTED positions
  std::vector<int32_t> u_cell, u_tile, u_flags, u_aero;   // [U]; u_cell = caller's cell index; flags: bit0 bare, bit1 overstory, bits 8.. root0 mask
  std::vector<int32_t> u_pos;        // [U] sorted position of the unit's cell
  std::vector<int32_t> u_slot;       // [U] reference-order unit (index into cell_uptr ranges) -> slot
  std::vector<int32_t> u_logical;    // [U] slot -> reference-order unit (column of the term buffer)
  std::vector<int64_t> pos_of;       // [C] caller's cell index -> sorted position
  std::vector<double> u_cv, u_af, u_Tfactor, u_Pfactor, u_rarc, u_rmin, u_RGL;   // [U]
  std::vector<float> u_root;         // [MAXL][U]
  std::vector<double> tm;            // [T][12][TM_N]
  std::vector<double> ae;            // [K][12][AE_N]
  std::vector<double> tl;            // [T][12][TL_N]  potential-evaporation tables, filled when an OUT_PET_* is requested
  std::vector<double> ap;            // [K][12][AP_N]
  std::vector<int32_t> cell_fail;    // [C] 1 = CalcAerodynamic rejects a tile of this cell (full_energy returns ERROR)
  std::string error;
};

// CalcAerodynamic.c:72-230 without the final wind scaling: Ra = RA/wind, U = UC*wind
inline int aero_coeffs(int OverStory, double Height, double Trunk, double Z0_SNOW, double Z0_SOIL, double n,
                       double *Ra, double *U, double *displacement, double *ref_height, double *roughness)
{
  const double von_K = 0.40;
  double K2 = von_K * von_K;
  if (!OverStory) {
    double Z0_Lower = roughness[0], d_Lower = displacement[0];
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
    double Z0_Upper = roughness[0], d_Upper = displacement[0], Z0_Lower = Z0_SOIL, d_Lower = 0;
    double Zw = 1.5 * Height - 0.5 * d_Upper, Zt = Trunk * Height;
    if (Zt < (Z0_Lower + d_Lower)) return -1;
    Ra[1] = log((ref_height[0] - d_Upper) / Z0_Upper) / K2 *
            (Height / (n * (Zw - d_Upper)) * (exp(n * (1 - (d_Upper + Z0_Upper) / Height)) - 1) +
             (Zw - Height) / (Zw - d_Upper) + log((ref_height[0] - d_Upper) / (Zw - d_Upper)));
    double Uw = log((Zw - d_Upper) / Z0_Upper) / log((ref_height[0] - d_Upper) / Z0_Upper);
    double Uh = Uw - (1 - (Height - d_Upper) / (Zw - d_Upper)) / log((ref_height[0] - d_Upper) / Z0_Upper);
    U[1] = Uh * exp(n * ((Z0_Upper + d_Upper) / Height - 1.));
    double Ut = Uh * exp(n * (Zt / Height - 1.));
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
  return 0;
}

inline uint64_t dbits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }

inline bool prepare(const vr_options &opt, const vr_veglib &lib, const vr_cells &cells, int cta_threads, Prepared &P)
{
  const int NL = opt.nlayer, NB = opt.nbands;
  const int64_t C = cells.ncell, T = cells.tile_ptr[C];
  P.NL = NL; P.NB = NB; P.C = C; P.T = T; P.cta_threads = cta_threads;
  if (NL < 2 || NL > VR_MAX_LAYERS || NB < 1 || NB > VR_MAX_BANDS) { P.error = "nlayer must be 2..3 and nbands 1..10"; return false; }
  // The cells keep the caller's order (P.order is the identity; the indirection stays in the kernels): the per-cell
  // aggregation then reads the forcing and writes the outputs coalesced.  What is sorted -- by kind and climate -- is the
  // UNITS, into thread slots (below).  -DVR_SORT_CELLS restores round 1's climate-sorted cells.
  P.order.resize(C);
  for (int64_t c = 0; c < C; c++) P.order[c] = c;
#if defined(VR_SORT_CELLS)
  std::stable_sort(P.order.begin(), P.order.end(), [&](int64_t a, int64_t b) {
    if (cells.avg_temp[a] != cells.avg_temp[b]) return cells.avg_temp[a] < cells.avg_temp[b];
    return cells.lat[a] > cells.lat[b];
  });
#endif
  // per-cell constants, stored BY SORTED POSITION so that neighbouring lanes read neighbouring words
  bool want_pet_nat = false, want_zwt = false, want_pet = false;
  for (int v = 0; v < opt.n_out; v++) {
    if (opt.out_vars[v] >= VR_OUT_PET_SATSOIL && opt.out_vars[v] <= VR_OUT_PET_VEGNOCR) want_pet = true;
    if (opt.out_vars[v] == VR_OUT_PET_NATVEG || opt.out_vars[v] == VR_OUT_PET_VEGNOCR) want_pet_nat = true;
    if (opt.out_vars[v] == VR_OUT_ZWT || opt.out_vars[v] == VR_OUT_ZWT_LUMPED) want_zwt = true;
  }
  if (want_pet_nat && (!lib.LAI || !lib.albedo)) { P.error = "OUT_PET_NATVEG / OUT_PET_VEGNOCR need vr_veglib.LAI and .albedo"; return false; }
  if (want_zwt && (!cells.zwt_zwt || !cells.zwt_moist)) { P.error = "OUT_ZWT / OUT_ZWT_LUMPED need vr_cells.zwt_zwt and .zwt_moist"; return false; }
  const int CCN = VR_CC_N(NL) + (want_zwt ? VR_CC_ZWT_N(NL) : 0);
  P.ccn = CCN;
  P.cc.assign((size_t)CCN * C, 0.0);
  for (int64_t pos = 0; pos < C; pos++) {
    const int64_t c = P.order[pos];
    auto CC = [&](int k, int64_t) -> double & { return P.cc[(size_t)k * C + pos]; };
    const double b = cells.b_infilt[c], elev = cells.elevation[c], dp = cells.dp[c];
    CC(CC_LAT, c) = cells.lat[c];
    CC(CC_ELEV, c) = elev;
    CC(CC_HALF_ELEV_LAPSE, c) = 0.5 * elev * (-0.0065);
    CC(CC_B, c) = b;
    CC(CC_INV_B, c) = (1.0 / b);
    CC(CC_INV_B1, c) = (1.0 / (b + 1.0));
    CC(CC_ONE_PLUS_B, c) = 1.0 * (1.0 + b);
    CC(CC_EX, c) = b / (1.0 + b);
    double top_max = 0.;
    for (int l = 0; l < NL - 1; l++) top_max += cells.max_moist[(size_t)l * C + c];
    CC(CC_TOP_MAX, c) = top_max;
    CC(CC_TOP_MAX_INFIL, c) = (1.0 + b) * top_max;
    const double d0 = cells.depth[c];
    const double mm_frac = cells.max_moist[c] / (d0 * 1000.);        // calc_surf_energy_bal.c max_moist
    const double arno_maxm = mm_frac * d0 * 1000.;                    // func_surf_energy_bal.c -> arno_evap max_moist
    CC(CC_ARNO_MAXM, c) = arno_maxm;
    CC(CC_ARNO_MAX_INFIL, c) = (1.0 + b) * arno_maxm;
    const double Dsmax24 = cells.Dsmax[c] / 24.;
    CC(CC_DSMAX24, c) = Dsmax24;
    CC(CC_BF_FRAC, c) = Dsmax24 * cells.Ds[c] / cells.Ws[c];
    CC(CC_BF_NL, c) = Dsmax24 * (1 - cells.Ds[c] / cells.Ws[c]);
    CC(CC_WS, c) = cells.Ws[c];
    CC(CC_ONE_M_WS, c) = (1 - cells.Ws[c]);
    CC(CC_INV_ONE_M_WS, c) = 1.0 / (1 - cells.Ws[c]);
    CC(CC_C, c) = cells.c_expt[c];
    CC(CC_AVG_T, c) = cells.avg_temp[c];
    CC(CC_DP, c) = dp;
    CC(CC_D1, c) = d0;
    CC(CC_EXP_M, c) = exp(-d0 / dp);
    CC(CC_EXP_P, c) = exp(d0 / dp);
    {   // bare-gap resistance coefficient: CalcAerodynamic(0,0,0,snow_rough,rough,0,...) with ref height
        // WINDH_SOIL and displacement 0.67*rough/0.123 (func_surf_energy_bal.c:717-727)
      double rough = cells.rough[c];
      double tmp_height = rough / 0.123;
      double dsp[3] = {0.67 * tmp_height, 0, 0}, rgh[3] = {rough, 0, 0}, rfh[3] = {10.0, 0, 0}, Ra[3], U[3];
      aero_coeffs(0, 0, 0, cells.snow_rough[c], rough, 0, Ra, U, dsp, rfh, rgh);
      CC(CC_KB_BARE, c) = Ra[0];
    }
    for (int l = 0; l < NL; l++) {
      const size_t k = (size_t)l * C + c;
      const int fb = CC_FL0 + l * CLF_N, bb = CC_BL0 + l * CLB_N;
      const double depth = cells.depth[k], resid = cells.resid_moist[k] * depth * 1000.;
      CC(bb + CLB_EXPT, c) = cells.expt[k];
      CC(bb + CLB_KSAT24, c) = cells.Ksat[k] / 24.;
      CC(bb + CLB_MAX_MOIST, c) = cells.max_moist[k];
      CC(fb + CLF_RESID, c) = resid;
      CC(bb + CLB_QDEN, c) = cells.max_moist[k] - resid;
      CC(bb + CLB_INV_QDEN, c) = 1.0 / (cells.max_moist[k] - resid);
      CC(fb + CLF_WCR, c) = cells.Wcr[k];
      CC(fb + CLF_WPWP, c) = cells.Wpwp[k];
      CC(fb + CLF_WCR_M_WPWP, c) = (cells.Wcr[k] - cells.Wpwp[k]);
      CC(bb + CLB_WET_DEN, c) = (cells.porosity[k] * depth * 1000 - cells.Wpwp[k]);
      CC(fb + CLF_DEPTH, c) = depth;
      {   // soil_conduction.c:5-49 constants (ice-free: Wu == moist)
        const double organic = cells.organic[k], quartz = cells.quartz[k];
        const double Kdry_min = (0.135 * cells.bulk_dens_min[k] + 64.7) / (cells.soil_dens_min[k] - 0.947 * cells.bulk_dens_min[k]);
        const double Kdry = (1 - organic) * Kdry_min + organic * 0.05;
        const double porosity = 1.0 - cells.bulk_density[k] / cells.soil_density[k];
        double Ks_min;
        if (quartz < .2) Ks_min = pow(7.7, quartz) * pow(3.0, 1.0 - quartz);
        else Ks_min = pow(7.7, quartz) * pow(2.2, 1.0 - quartz);
        const double Ks = (1 - organic) * Ks_min + organic * 0.25;
        const double Ksat = pow(Ks, 1.0 - porosity) * pow(0.57, porosity);
        const double soilfr = cells.bulk_density[k] / cells.soil_density[k];
        double cs = 2.0e6 * soilfr * (1 - organic);
        cs += 2.7e6 * soilfr * organic;
        CC(fb + CLF_KDRY, c) = Kdry;
        CC(fb + CLF_KSAT_M_KDRY, c) = (Ksat - Kdry);
        CC(fb + CLF_POROS, c) = porosity;
        CC(fb + CLF_SOILFR, c) = soilfr;
        CC(fb + CLF_CS_SOLID, c) = cs;
      }
    }
    if (want_zwt)
      for (int k = 0; k < NL + 2; k++)
        for (int i = 0; i < ZWT_NP; i++) {
          const size_t q = ((size_t)k * ZWT_NP + i) * C + c;
          CC(VR_CC_N(NL) + k * ZWT_NP + i, c) = cells.zwt_zwt[q];
          CC(VR_CC_N(NL) + (NL + 2 + k) * ZWT_NP + i, c) = cells.zwt_moist[q];
        }
  }

  // ---- tile x month canopy terms ----
  P.tm.assign((size_t)T * 12 * TM_N, 0.0);
  P.tl.assign(want_pet ? (size_t)T * 12 * TL_N : 1, 0.0);
  for (int64_t t = 0; t < T; t++) {
    int cls = cells.tile_class[t];
    if (cls < 0 || cls > lib.nclass) { P.error = "tile_class out of range"; return false; }
    for (int m = 0; m < 12 && want_pet; m++) {
      double *q = &P.tl[((size_t)t * 12 + m) * TL_N];
      if (cls == lib.nclass) { q[TL_LAI] = 1.0; q[TL_ALB] = BARE_SOIL_ALBEDO; }   // SATSOIL reference entry (global.h:56-57)
      else { q[TL_LAI] = lib.LAI ? lib.LAI[cls * 12 + m] : 0.0; q[TL_ALB] = lib.albedo ? lib.albedo[cls * 12 + m] : 0.0; }
    }
    if (cls == lib.nclass) continue;
    for (int m = 0; m < 12; m++) {
      double *q = &P.tm[((size_t)t * 12 + m) * TM_N];
      double vegcover = cells.tile_vegcover[t * 12 + m];
      if (vegcover < 0.0001) vegcover = 0.0001;          // MIN_VEGCOVER, full_energy.c:214-215
      double LAI = cells.tile_LAI[t * 12 + m];
      LAI /= vegcover;
      q[TM_VEGCOVER] = vegcover;
      q[TM_ALBEDO] = cells.tile_albedo[t * 12 + m];
      q[TM_LAI] = LAI;
      q[TM_WDMAX] = LAI * 0.1;                           // LAI_WATER_FACTOR
      q[TM_SURF_ATTEN] = (1 - vegcover) * 1.0 + vegcover * exp(-lib.rad_atten[cls] * LAI);
    }
  }

  // ---- units, aero keys ----
  P.cell_uptr.assign(C + 1, 0);
  P.cell_fail.assign(C, 0);
  std::map<std::tuple<int, uint64_t, uint64_t>, int> keys;
  std::vector<std::tuple<int, double, double>> keylist;
  std::vector<int32_t> t_key(T, 0);
  for (int64_t c = 0; c < C; c++) {
    for (int64_t t = cells.tile_ptr[c]; t < cells.tile_ptr[c + 1]; t++) {
      auto k = std::make_tuple((int)cells.tile_class[t], dbits(cells.rough[c]), dbits(cells.snow_rough[c]));
      auto it = keys.find(k);
      if (it == keys.end()) {
        int id = (int)keylist.size();
        keys[k] = id;
        keylist.emplace_back((int)cells.tile_class[t], cells.rough[c], cells.snow_rough[c]);
        t_key[t] = id;
      }
      else t_key[t] = it->second;
    }
  }
  P.K = (int64_t)keylist.size();
  P.ae.assign((size_t)(P.K ? P.K : 1) * 12 * AE_N, 0.0);
  P.ap.assign(want_pet ? (size_t)(P.K ? P.K : 1) * 12 * AP_N : 1, 0.0);
  std::vector<char> key_fail(P.K ? P.K : 1, 0);
  for (int64_t k = 0; k < P.K; k++) {
    int cls = std::get<0>(keylist[k]);
    double rough = std::get<1>(keylist[k]), snow_rough = std::get<2>(keylist[k]);
    bool bare = (cls == lib.nclass);
    for (int m = 0; m < 12; m++) {
      double dsp[3] = {0, 0, 0}, rgh[3] = {0, 0, 0}, rfh[3] = {0, 0, 0}, Ra[3] = {0, 0, 0}, U[3] = {0, 0, 0};
      double wind_h = bare ? 10.0 : lib.wind_h[cls];
      dsp[0] = bare ? rough * 0.667 / 0.123 : lib.displacement[cls * 12 + m];   // read_soilparam.c:551-554
      rgh[0] = bare ? rough : lib.roughness[cls * 12 + m];
      int overstory = bare ? 0 : lib.overstory[cls];
      if (rgh[0] == 0) rgh[0] = rough;                                           // full_energy.c:344-345
      double height = dsp[0] / 0.67;                                             // calc_veg_height
      if (dsp[0] < wind_h) rfh[0] = wind_h;
      else rfh[0] = dsp[0] + wind_h + rgh[0];
      int rc = aero_coeffs(overstory, height, bare ? 0.0 : lib.trunk_ratio[cls], snow_rough, rough,
                           bare ? 0.0 : lib.wind_atten[cls], Ra, U, dsp, rfh, rgh);
      if (rc != 0) key_fail[k] = 1;
      double *q = &P.ae[((size_t)k * 12 + m) * AE_N];
      for (int i = 0; i < 3; i++) {
        q[AE_RA0 + i] = Ra[i]; q[AE_U0 + i] = U[i]; q[AE_REF0 + i] = rfh[i]; q[AE_ROUGH0 + i] = rgh[i];
        q[AE_DISP0 + i] = dsp[i];
      }
      // the four PET reference covers seen from this tile (full_energy.c:329-369): their own displacement and
      // roughness (global.h:59-60; SATSOIL's were overwritten with the cell's soil values, read_soilparam.c:551-554),
      // no overstory, but the reference height test uses the TILE's wind_h
      for (int p = 0; p < 4 && want_pet; p++) {
        const double ref_rough[4] = {0.001, 0.001, 0.0148, 0.0615}, ref_displ[4] = {0.0054, 0.0054, 0.08, 0.3333};
        double d3[3] = {0, 0, 0}, r3[3] = {0, 0, 0}, h3[3] = {0, 0, 0}, Rp[3] = {0, 0, 0}, Up[3] = {0, 0, 0};
        d3[0] = (p == 0) ? rough * 0.667 / 0.123 : ref_displ[p];
        r3[0] = (p == 0) ? rough : ref_rough[p];
        if (d3[0] < wind_h) h3[0] = wind_h;
        else h3[0] = d3[0] + wind_h + r3[0];
        if (aero_coeffs(0, d3[0] / 0.67, 0.0, snow_rough, rough, 0.0, Rp, Up, d3, h3, r3) != 0) key_fail[k] = 1;
        for (int i = 0; i < 3; i++) P.ap[((size_t)k * 12 + m) * AP_N + 3 * p + i] = Rp[i];
      }
    }
  }
  for (int64_t cs = 0; cs < C; cs++) {
    const int64_t c = P.order[cs];
    int64_t n = 0;
    for (int64_t t = cells.tile_ptr[c]; t < cells.tile_ptr[c + 1]; t++) {
      if (!(cells.tile_Cv[t] > 0.0)) continue;
      if (key_fail[t_key[t]]) P.cell_fail[c] = 1;
      int cls = cells.tile_class[t];
      bool bare = (cls == lib.nclass);
      int64_t t0 = cells.tile_ptr[c];
      bool has_veg0 = (cells.tile_class[t0] != lib.nclass);
      int mask = 0;
      for (int l = 0; l < NL; l++)
        if (has_veg0 && (float)cells.tile_root[t0 * NL + l] > 0) mask |= (1 << l);   // veg_con->root of tile 0
      for (int b = 0; b < NB; b++) {
        if (!(cells.AreaFract[(size_t)b * C + c] > 0)) continue;
        P.u_cell.push_back((int32_t)c);
        P.u_pos.push_back((int32_t)cs);
        P.u_tile.push_back((int32_t)t);
        P.u_flags.push_back((bare ? 1 : 0) | ((!bare && lib.overstory[cls]) ? 2 : 0) | (mask << 8));
        P.u_aero.push_back(t_key[t]);
        P.u_cv.push_back(cells.tile_Cv[t]);
        P.u_af.push_back(cells.AreaFract[(size_t)b * C + c]);
        P.u_Tfactor.push_back(cells.Tfactor[(size_t)b * C + c]);
        P.u_Pfactor.push_back(cells.Pfactor[(size_t)b * C + c]);
        P.u_rarc.push_back(bare ? 0.0 : lib.rarc[cls]);
        P.u_rmin.push_back(bare ? 0.0 : lib.rmin[cls]);
        P.u_RGL.push_back(bare ? 0.0 : (double)(float)lib.RGL[cls]);
        n++;
      }
    }
    P.cell_uptr[cs + 1] = P.cell_uptr[cs] + n;
  }
  P.U = P.cell_uptr[C];
  P.u_root.assign((size_t)MAXL * (P.U ? P.U : 1), 0.f);
  for (int64_t u = 0; u < P.U; u++) {
    int64_t t = P.u_tile[u];
    bool bare = P.u_flags[u] & 1;
    for (int l = 0; l < NL; l++) P.u_root[(size_t)l * P.U + u] = bare ? 0.f : (float)cells.tile_root[t * NL + l];
  }
  // ---- thread slots: sort by kind, then by the unit's climate (cell mean temperature + band lapse offset) ----
  P.pos_of.assign(C, 0);
  for (int64_t cs = 0; cs < C; cs++) P.pos_of[P.order[cs]] = cs;
  const int64_t U = P.U;
  std::vector<int64_t> perm(U);
  for (int64_t u = 0; u < U; u++) perm[u] = u;
  auto kind_of = [&](int64_t u) { const int fl = P.u_flags[u]; return (fl & 2) ? 0 : ((fl & 1) ? 2 : 1); };
  std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) {
    const int ka = kind_of(a), kb = kind_of(b);
    if (ka != kb) return ka < kb;
#ifndef VR_SORT_NO_HEMISPHERE
    // the two hemispheres have opposite snow seasons: keep them in different warps
    const bool na = cells.lat[P.u_cell[a]] >= 0, nb = cells.lat[P.u_cell[b]] >= 0;
    if (na != nb) return na;
#endif
    // (a latitude-dependent seasonal-amplitude term in the key was tried and is slower: profiles/r1_tuning.md)
    const double ta = cells.avg_temp[P.u_cell[a]] + P.u_Tfactor[a], tb = cells.avg_temp[P.u_cell[b]] + P.u_Tfactor[b];
    if (ta != tb) return ta < tb;
    return a < b;
  });
  P.u_slot.assign(U ? U : 1, 0);
  P.u_logical.assign(U ? U : 1, 0);
  for (int64_t k = 0; k < U; k++) { P.u_slot[perm[k]] = (int32_t)k; P.u_logical[k] = (int32_t)perm[k]; }
  auto permute = [&](auto &v) {
    auto old = v;
    for (int64_t k = 0; k < U; k++) v[k] = old[perm[k]];
  };
  permute(P.u_cell); permute(P.u_pos); permute(P.u_tile); permute(P.u_flags); permute(P.u_aero); permute(P.u_cv);
  permute(P.u_af); permute(P.u_Tfactor); permute(P.u_Pfactor); permute(P.u_rarc); permute(P.u_rmin); permute(P.u_RGL);
  {
    std::vector<float> old = P.u_root;
    for (int l = 0; l < MAXL; l++)
      for (int64_t k = 0; k < U; k++) P.u_root[(size_t)l * U + k] = old[(size_t)l * U + perm[k]];
  }
  P.nCTA = (U + cta_threads - 1) / cta_threads;
  return true;
}

}  // namespace vr
