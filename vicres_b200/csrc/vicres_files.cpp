// vicres_files.cpp -- readers for the reference's parameter files and the ASCII flux writer
// (include/vicres_files.h).  Host C++ only; no CUDA here.  Each reader follows the conversions of
// the reference reader it replaces (sscanf %f -> strtof, %lf -> strtod, float fields kept float).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/vicres_files.h"

namespace {

std::string g_err;

struct VegClass {
  int id = 0, overstory = 0;
  double rarc = 0, rmin = 0, LAI[12], albedo[12], roughness[12], displacement[12], wind_h = 0, rad_atten = 0,
         wind_atten = 0, trunk_ratio = 0;
  float RGL = 0;
};

std::vector<std::string> tokens(const std::string &line)
{
  std::vector<std::string> t;
  size_t i = 0, n = line.size();
  while (i < n) {
    while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r' || line[i] == '\n')) i++;
    size_t j = i;
    while (j < n && !(line[j] == ' ' || line[j] == '\t' || line[j] == '\r' || line[j] == '\n')) j++;
    if (j > i) t.push_back(line.substr(i, j - i));
    i = j;
  }
  return t;
}

bool read_line(FILE *f, std::string &s)
{
  s.clear();
  int c;
  bool any = false;
  while ((c = fgetc(f)) != EOF) {
    any = true;
    if (c == '\n') break;
    s.push_back((char)c);
  }
  return any;
}

double linear_interp(double x, double lx, double ux, double ly, double uy) { return ((x - lx) / (ux - lx) * (uy - ly) + ly); }

}  // namespace

struct vr_params {
  int L = 0, NB = 0;
  std::vector<int32_t> gridcel;
  std::vector<float> lat, lng;
  std::vector<double> site_lng, site_tz, site_annual_prec;   // what initialize_atmos reads from soil_con
  std::vector<double> site_min_tf;                           // the lowest band Tfactor (initialize_atmos.c:1737-1741)
  // cells
  std::vector<double> c_lat, elevation, b_infilt, Ds, Dsmax, Ws, c_expt, avg_temp, dp, rough, snow_rough;
  std::vector<double> expt, Ksat, depth, max_moist, Wcr, Wpwp, resid_moist, init_moist, bulk_density, soil_density,
      bulk_dens_min, soil_dens_min, quartz, organic, porosity;           // [L][C]
  std::vector<double> AreaFract, Tfactor, Pfactor;                         // [NB][C]
  std::vector<int64_t> tile_ptr;
  std::vector<int32_t> tile_class;
  std::vector<double> tile_Cv, tile_root, tile_LAI, tile_albedo, tile_vegcover;
  // veglib
  std::vector<int32_t> l_over;
  std::vector<double> l_rarc, l_rmin, l_wind_h, l_RGL, l_rad, l_watt, l_trunk, l_rough, l_disp, l_LAI, l_alb;
  std::vector<double> zwt_zwt, zwt_moist;                                  // [(L+2)][11][C]
  vr_cells cells;
  vr_veglib lib;
};

extern "C" {

const char *vr_params_last_error(void) { return g_err.c_str(); }
void vr_params_free(vr_params *p) { delete p; }
const vr_cells *vr_params_cells(const vr_params *p) { return p ? &p->cells : nullptr; }
const vr_veglib *vr_params_veglib(const vr_params *p) { return p ? &p->lib : nullptr; }
const int32_t *vr_params_gridcel(const vr_params *p) { return p ? p->gridcel.data() : nullptr; }
const float *vr_params_lat(const vr_params *p) { return p ? p->lat.data() : nullptr; }
const float *vr_params_lng(const vr_params *p) { return p ? p->lng.data() : nullptr; }
const double *vr_params_site(const vr_params *p, int32_t which)
{
  if (!p) return nullptr;
  switch (which) {
    case 0: return p->cells.lat;
    case 1: return p->site_lng.data();
    case 2: return p->site_tz.data();
    case 3: return p->cells.elevation;
    case 4: return p->site_annual_prec.data();
    case 5: return p->site_min_tf.data();
    default: return nullptr;
  }
}

int vr_params_read(const vr_param_files *fo, vr_params **out)
{
  if (!fo || !out || !fo->soil || !fo->veglib || !fo->vegparam) { g_err = "null argument"; return VR_ERR_ARG; }
  *out = nullptr;
  const int L = fo->nlayer, NB = fo->nbands, RZ = fo->root_zones;
  if (L < 2 || L > VR_MAX_LAYERS || NB < 1 || NB > VR_MAX_BANDS || RZ < 1 || RZ > 16) { g_err = "nlayer/nbands/root_zones out of range"; return VR_ERR_ARG; }
  if (NB > 1 && !fo->snowband) { g_err = "SNOW_BAND > 1 needs a snow band file"; return VR_ERR_ARG; }

  // ---- vegetation library (read_veglib.c): class lines start with a digit ----
  std::vector<VegClass> lib;
  {
    FILE *f = fopen(fo->veglib, "r");
    if (!f) { g_err = std::string("cannot open ") + fo->veglib; return VR_ERR_ARG; }
    std::string line;
    while (read_line(f, line)) {
      size_t i = 0;
      while (i < line.size() && (line[i] == ' ' || line[i] == '\t')) i++;
      if (i >= line.size() || line[i] < '0' || line[i] > '9') continue;
      std::vector<std::string> t = tokens(line);
      const size_t need = 4 + 12 * 4 + 5;
      if (t.size() < need) { fclose(f); g_err = "vegetation library line has too few columns"; return VR_ERR_ARG; }
      VegClass v;
      size_t k = 0;
      v.id = atoi(t[k++].c_str());
      v.overstory = atoi(t[k++].c_str());
      v.rarc = strtod(t[k++].c_str(), nullptr);
      v.rmin = strtod(t[k++].c_str(), nullptr);
      for (int j = 0; j < 12; j++) v.LAI[j] = strtod(t[k++].c_str(), nullptr);
      for (int j = 0; j < 12; j++) v.albedo[j] = strtod(t[k++].c_str(), nullptr);
      for (int j = 0; j < 12; j++) v.roughness[j] = strtod(t[k++].c_str(), nullptr);
      for (int j = 0; j < 12; j++) v.displacement[j] = strtod(t[k++].c_str(), nullptr);
      v.wind_h = strtod(t[k++].c_str(), nullptr);
      v.RGL = strtof(t[k++].c_str(), nullptr);
      v.rad_atten = strtod(t[k++].c_str(), nullptr);
      v.wind_atten = strtod(t[k++].c_str(), nullptr);
      v.trunk_ratio = strtod(t[k++].c_str(), nullptr);
      lib.push_back(v);
    }
    fclose(f);
    if (lib.empty()) { g_err = "no classes in the vegetation library"; return VR_ERR_ARG; }
  }

  vr_params *P = new vr_params();
  P->L = L; P->NB = NB;
  struct Soil {
    int gridcel; float lat, lng, elevation; double off_gmt, annual_prec;
    double b, Ds, Dsmax, Ws, c, avg_temp, dp, rough, snow_rough;
    double expt[3], Ksat[3], init_moist[3], depth[3], quartz[3], bdm[3], sdm[3], wcr_fr[3], wpwp_fr[3], resid[3], bubble[3];
    double org[3], bdo[3], sdo[3];
  };
  std::vector<Soil> soils;
  {
    FILE *f = fopen(fo->soil, "r");
    if (!f) { delete P; g_err = std::string("cannot open ") + fo->soil; return VR_ERR_ARG; }
    std::string line;
    while (read_line(f, line)) {
      std::vector<std::string> t = tokens(line);
      if (t.empty()) continue;
      if (atoi(t[0].c_str()) == 0) continue;                  // RUN_MODEL flag
      const size_t need = 9 + 4 * L + 1 + L + 2 + 4 * L + 1 + 2 * L + 3 + L + 1 + (fo->organic_fract ? 3 * L : 0);
      if (t.size() < need) { fclose(f); delete P; g_err = "soil file row has too few columns for NLAYER"; return VR_ERR_ARG; }
      Soil s;
      size_t k = 1;
      s.gridcel = atoi(t[k++].c_str());
      s.lat = strtof(t[k++].c_str(), nullptr);
      s.lng = strtof(t[k++].c_str(), nullptr);
      s.b = strtod(t[k++].c_str(), nullptr);
      s.Ds = strtod(t[k++].c_str(), nullptr);
      s.Dsmax = strtod(t[k++].c_str(), nullptr);
      s.Ws = strtod(t[k++].c_str(), nullptr);
      s.c = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.expt[l] = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.Ksat[l] = strtod(t[k++].c_str(), nullptr);
      k += L;                                                  // phi_s
      for (int l = 0; l < L; l++) s.init_moist[l] = strtod(t[k++].c_str(), nullptr);
      s.elevation = strtof(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) {
        double d = strtod(t[k++].c_str(), nullptr);
        s.depth[l] = (float)(int)(d * 1000 + 0.5) / 1000;     // read_soilparam.c:342: float / int
      }
      s.avg_temp = strtod(t[k++].c_str(), nullptr);
      s.dp = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.bubble[l] = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.quartz[l] = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.bdm[l] = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.sdm[l] = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) { s.org[l] = 0.0; s.bdo[l] = -9999; s.sdo[l] = -9999; }      // read_soilparam.c:501-506
      if (fo->organic_fract) {                                                               // :444-499
        for (int l = 0; l < L; l++) s.org[l] = strtod(t[k++].c_str(), nullptr);
        for (int l = 0; l < L; l++) s.bdo[l] = strtod(t[k++].c_str(), nullptr);
        for (int l = 0; l < L; l++) s.sdo[l] = strtod(t[k++].c_str(), nullptr);
        for (int l = 0; l < L; l++) {
          if (s.org[l] > 1. || s.org[l] < 0) { fclose(f); delete P; g_err = "Need valid volumetric organic soil fraction when options.ORGANIC_FRACT is set to TRUE."; return VR_ERR_ARG; }
          if (s.bdo[l] <= 0 && s.org[l] > 0) s.bdo[l] = s.bdm[l];
          if (s.sdo[l] <= 0 && s.org[l] > 0) s.sdo[l] = s.sdm[l];
          if (s.org[l] > 0 && s.bdo[l] >= s.sdo[l]) { fclose(f); delete P; g_err = "organic bulk density must be less than organic soil density"; return VR_ERR_ARG; }
        }
      }
      s.off_gmt = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.wcr_fr[l] = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.wpwp_fr[l] = strtod(t[k++].c_str(), nullptr);
      s.rough = strtod(t[k++].c_str(), nullptr);
      s.snow_rough = strtod(t[k++].c_str(), nullptr);
      s.annual_prec = strtod(t[k++].c_str(), nullptr);
      for (int l = 0; l < L; l++) s.resid[l] = strtod(t[k++].c_str(), nullptr);
      if (s.b <= 0) { fclose(f); delete P; g_err = "b_infilt must be positive"; return VR_ERR_ARG; }
      soils.push_back(s);
    }
    fclose(f);
  }
  const int64_t C = (int64_t)soils.size();
  if (C == 0) { delete P; g_err = "no cells flagged to run in the soil file"; return VR_ERR_ARG; }
  auto LC = [&](std::vector<double> &v) { v.assign((size_t)L * C, 0.0); };
  P->c_lat.resize(C); P->elevation.resize(C); P->b_infilt.resize(C); P->Ds.resize(C); P->Dsmax.resize(C); P->Ws.resize(C);
  P->c_expt.resize(C); P->avg_temp.resize(C); P->dp.resize(C); P->rough.resize(C); P->snow_rough.resize(C);
  LC(P->expt); LC(P->Ksat); LC(P->depth); LC(P->max_moist); LC(P->Wcr); LC(P->Wpwp); LC(P->resid_moist); LC(P->init_moist);
  LC(P->bulk_density); LC(P->soil_density); LC(P->bulk_dens_min); LC(P->soil_dens_min); LC(P->quartz); LC(P->organic); LC(P->porosity);
  P->AreaFract.assign((size_t)NB * C, 0.0); P->Tfactor.assign((size_t)NB * C, 0.0); P->Pfactor.assign((size_t)NB * C, 1.0);
  P->gridcel.resize(C); P->lat.resize(C); P->lng.resize(C);
  P->site_lng.resize(C); P->site_tz.resize(C); P->site_annual_prec.resize(C);
  std::vector<float> elev_f(C);
  for (int64_t c = 0; c < C; c++) {
    const Soil &s = soils[c];
    P->gridcel[c] = s.gridcel; P->lat[c] = s.lat; P->lng[c] = s.lng; elev_f[c] = s.elevation;
    P->site_lng[c] = s.lng; P->site_tz[c] = (float)(s.off_gmt * 360. / 24.); P->site_annual_prec[c] = s.annual_prec;   // read_soilparam.c:934
    P->c_lat[c] = s.lat; P->b_infilt[c] = s.b; P->Ds[c] = s.Ds; P->Dsmax[c] = s.Dsmax; P->Ws[c] = s.Ws; P->c_expt[c] = s.c;
    P->avg_temp[c] = s.avg_temp; P->dp[c] = s.dp; P->rough[c] = s.rough; P->snow_rough[c] = s.snow_rough;
    for (int l = 0; l < L; l++) {
      const size_t k = (size_t)l * C + c;
      const double organic = s.org[l], bdo = s.bdo[l], sdo = s.sdo[l];      // read_soilparam.c:647-655
      const double bulk = (1 - organic) * s.bdm[l] + organic * bdo, sdens = (1 - organic) * s.sdm[l] + organic * sdo;
      const double poros = 1.0 - bulk / sdens, mm = s.depth[l] * poros * 1000.;
      P->expt[k] = s.expt[l]; P->Ksat[k] = s.Ksat[l]; P->depth[k] = s.depth[l]; P->max_moist[k] = mm;
      P->Wcr[k] = s.wcr_fr[l] * mm; P->Wpwp[k] = s.wpwp_fr[l] * mm;
      P->resid_moist[k] = (s.resid[l] == -99999.) ? 0.0 : s.resid[l];
      P->init_moist[k] = s.init_moist[l]; P->bulk_density[k] = bulk; P->soil_density[k] = sdens;
      P->bulk_dens_min[k] = s.bdm[l]; P->soil_dens_min[k] = s.sdm[l]; P->quartz[k] = s.quartz[l]; P->organic[k] = organic;
      P->porosity[k] = poros;
    }
    P->AreaFract[c] = 1.;                                       // band 0 (read_soilparam.c:478-486)
    if (fo->baseflow_nijssen) {                                 // read_soilparam.c:719-726, statement for statement
      const double mm = P->max_moist[(size_t)(L - 1) * C + c];
      P->Dsmax[c] = P->Dsmax[c] * pow((double)(1. / (mm - P->Ws[c])), -P->c_expt[c]) + P->Ds[c] * mm;
      P->Ds[c] = P->Ds[c] * P->Ws[c] / P->Dsmax[c];
      P->Ws[c] = P->Ws[c] / mm;
    }
  }
  // ---- soil moisture v water table curves (read_soilparam.c:822-921; arithmetic order kept) ----
  {
    const int NP = VR_ZWT_NPOINTS;
    P->zwt_zwt.assign((size_t)(L + 2) * NP * C, 0.0);
    P->zwt_moist.assign((size_t)(L + 2) * NP * C, 0.0);
    for (int64_t c = 0; c < C; c++) {
      const Soil &s = soils[c];
      double depth[3], max_moist[3], resid[3];
      for (int l = 0; l < L; l++) {
        depth[l] = s.depth[l]; max_moist[l] = P->max_moist[(size_t)l * C + c]; resid[l] = P->resid_moist[(size_t)l * C + c];
      }
      auto Z = [&](int k, int i) -> double & { return P->zwt_zwt[((size_t)k * NP + i) * C + c]; };
      auto M = [&](int k, int i) -> double & { return P->zwt_moist[((size_t)k * NP + i) * C + c]; };
      double tmp_depth = 0, b, bubble, tmp_resid_moist, zwt_prime, w_avg, tmp_max_moist, tmp_moist, tmp_depth2, b_save,
             bub_save, tmp_depth2_save, zwt_prime_eff;
      for (int layer = 0; layer < L; layer++) {
        b = 0.5 * (s.expt[layer] - 3);
        bubble = s.bubble[layer];
        tmp_resid_moist = resid[layer] * depth[layer] * 1000;
        zwt_prime = 0;
        for (int i = 0; i < NP; i++) {
          Z(layer, i) = -tmp_depth * 100 - zwt_prime;
          w_avg = (depth[layer] * 100 - zwt_prime - (b / (b - 1)) * bubble * (1 - pow((zwt_prime + bubble) / bubble, (b - 1) / b))) /
                  (depth[layer] * 100);
          if (w_avg < 0) w_avg = 0;
          if (w_avg > 1) w_avg = 1;
          M(layer, i) = w_avg * (max_moist[layer] - tmp_resid_moist) + tmp_resid_moist;
          zwt_prime += depth[layer] * 100 / (NP - 1);
        }
        tmp_depth += depth[layer];
      }
      tmp_depth = 0; b = 0; bubble = 0; tmp_max_moist = 0; tmp_resid_moist = 0;
      for (int layer = 0; layer < L - 1; layer++) {
        b += 0.5 * (s.expt[layer] - 3) * depth[layer];
        bubble += s.bubble[layer] * depth[layer];
        tmp_max_moist += max_moist[layer];
        tmp_resid_moist += resid[layer] * depth[layer] * 1000;
        tmp_depth += depth[layer];
      }
      b /= tmp_depth;
      bubble /= tmp_depth;
      zwt_prime = 0;
      for (int i = 0; i < NP; i++) {
        Z(L, i) = -zwt_prime;
        w_avg = (tmp_depth * 100 - zwt_prime - (b / (b - 1)) * bubble * (1 - pow((zwt_prime + bubble) / bubble, (b - 1) / b))) /
                (tmp_depth * 100);
        if (w_avg < 0) w_avg = 0;
        if (w_avg > 1) w_avg = 1;
        M(L, i) = w_avg * (tmp_max_moist - tmp_resid_moist) + tmp_resid_moist;
        zwt_prime += tmp_depth * 100 / (NP - 1);
      }
      tmp_depth = 0;
      for (int layer = 0; layer < L; layer++) tmp_depth += depth[layer];
      zwt_prime = 0;
      for (int i = 0; i < NP; i++) {
        Z(L + 1, i) = -zwt_prime;
        if (zwt_prime == 0) {
          tmp_moist = 0;
          for (int layer = 0; layer < L; layer++) tmp_moist += max_moist[layer];
          M(L + 1, i) = tmp_moist;
        }
        else {
          tmp_moist = 0;
          int layer = L - 1;
          tmp_depth2 = tmp_depth - depth[layer];
          while (layer > 0 && zwt_prime <= tmp_depth2 * 100) {
            tmp_moist += max_moist[layer];
            layer--;
            tmp_depth2 -= depth[layer];
          }
          w_avg = (tmp_depth2 * 100 + depth[layer] * 100 - zwt_prime) / (depth[layer] * 100);
          b = 0.5 * (s.expt[layer] - 3);
          bubble = s.bubble[layer];
          tmp_resid_moist = resid[layer] * depth[layer] * 1000;
          w_avg += -(b / (b - 1)) * bubble * (1 - pow((zwt_prime + bubble - tmp_depth2 * 100) / bubble, (b - 1) / b)) / (depth[layer] * 100);
          tmp_moist += w_avg * (max_moist[layer] - tmp_resid_moist) + tmp_resid_moist;
          b_save = b; bub_save = bubble; tmp_depth2_save = tmp_depth2;
          while (layer > 0) {
            layer--;
            tmp_depth2 -= depth[layer];
            b = 0.5 * (s.expt[layer] - 3);
            bubble = s.bubble[layer];
            tmp_resid_moist = resid[layer] * depth[layer] * 1000;
            zwt_prime_eff = tmp_depth2_save * 100 - bubble + bubble * pow((zwt_prime + bub_save - tmp_depth2_save * 100) / bub_save, b / b_save);
            w_avg = -(b / (b - 1)) * bubble * (1 - pow((zwt_prime_eff + bubble - tmp_depth2 * 100) / bubble, (b - 1) / b)) / (depth[layer] * 100);
            tmp_moist += w_avg * (max_moist[layer] - tmp_resid_moist) + tmp_resid_moist;
            b_save = b; bub_save = bubble; tmp_depth2_save = tmp_depth2;
          }
          M(L + 1, i) = tmp_moist;
        }
        zwt_prime += tmp_depth * 100 / (NP - 1);
      }
    }
  }

  // ---- snow bands (read_snowband.c) ----
  if (NB > 1) {
    FILE *f = fopen(fo->snowband, "r");
    if (!f) { delete P; g_err = std::string("cannot open ") + fo->snowband; return VR_ERR_ARG; }
    std::map<int, std::vector<std::string>> rows;
    std::string line;
    while (read_line(f, line)) {
      std::vector<std::string> t = tokens(line);
      if (t.size() >= (size_t)(1 + 3 * NB)) rows[atoi(t[0].c_str())] = t;
    }
    fclose(f);
    for (int64_t c = 0; c < C; c++) {
      auto it = rows.find(P->gridcel[c]);
      if (it == rows.end()) continue;                            // the reference warns and keeps one band
      const std::vector<std::string> &t = it->second;
      double total = 0., af[VR_MAX_BANDS], pf[VR_MAX_BANDS];
      float be[VR_MAX_BANDS], avg_elev = 0;
      for (int b = 0; b < NB; b++) { af[b] = strtod(t[1 + b].c_str(), nullptr); total += af[b]; }
      if (total != 1.) for (int b = 0; b < NB; b++) af[b] /= total;
      for (int b = 0; b < NB; b++) { be[b] = strtof(t[1 + NB + b].c_str(), nullptr); avg_elev += be[b] * af[b]; }
      if (fabs(avg_elev - elev_f[c]) > 1.0) elev_f[c] = (float)avg_elev;
      total = 0.;
      for (int b = 0; b < NB; b++) { pf[b] = strtod(t[1 + 2 * NB + b].c_str(), nullptr); total += pf[b]; }
      if (total != 1.) for (int b = 0; b < NB; b++) pf[b] /= total;
      for (int b = 0; b < NB; b++) {
        const size_t k = (size_t)b * C + c;
        P->AreaFract[k] = af[b];
        P->Tfactor[k] = (elev_f[c] - be[b]) * -1 * (-0.0065);
        P->Pfactor[k] = (af[b] > 0) ? pf[b] / af[b] : 0.;
      }
    }
  }
  for (int64_t c = 0; c < C; c++) P->elevation[c] = elev_f[c];

  // ---- vegetation parameters (read_vegparam.c) + root fractions (calc_root_fraction.c) ----
  {
    FILE *f = fopen(fo->vegparam, "r");
    if (!f) { delete P; g_err = std::string("cannot open ") + fo->vegparam; return VR_ERR_ARG; }
    struct Tile { int cls; double Cv; std::vector<float> zd, zf; double LAI[12]; bool has_lai; };
    std::map<int, std::vector<Tile>> blocks;
    std::string line;
    const int per_tile = 1 + (fo->vegparam_lai ? 1 : 0);
    while (read_line(f, line)) {
      std::vector<std::string> t = tokens(line);
      if (t.size() < 2) continue;
      const int cell = atoi(t[0].c_str()), nveg = atoi(t[1].c_str());
      std::vector<Tile> tiles;
      for (int i = 0; i < nveg; i++) {
        Tile tl;
        tl.has_lai = false;
        if (!read_line(f, line)) break;
        std::vector<std::string> v = tokens(line);
        if ((int)v.size() != 2 + 2 * RZ) { fclose(f); delete P; g_err = "vegetation parameter tile line: wrong number of fields for ROOT_ZONES"; return VR_ERR_ARG; }
        tl.cls = atoi(v[0].c_str());
        tl.Cv = atof(v[1].c_str());
        float sum = 0.f;
        for (int j = 0; j < RZ; j++) {
          tl.zd.push_back((float)atof(v[2 + 2 * j].c_str()));
          tl.zf.push_back((float)atof(v[3 + 2 * j].c_str()));
          sum += tl.zf.back();
        }
        if (sum != 1.f) for (int j = 0; j < RZ; j++) tl.zf[j] /= sum;
        if (per_tile == 2) {
          if (!read_line(f, line)) break;
          std::vector<std::string> m = tokens(line);
          if (m.size() != 12) { fclose(f); delete P; g_err = "vegetation parameter LAI line needs 12 values"; return VR_ERR_ARG; }
          for (int j = 0; j < 12; j++) tl.LAI[j] = atof(m[j].c_str());
          tl.has_lai = true;
        }
        tiles.push_back(tl);
      }
      blocks[cell] = tiles;
    }
    fclose(f);
    P->tile_ptr.assign(C + 1, 0);
    for (int64_t c = 0; c < C; c++) {
      auto it = blocks.find(P->gridcel[c]);
      if (it == blocks.end()) { delete P; g_err = "grid cell missing from the vegetation parameter file"; return VR_ERR_ARG; }
      std::vector<Tile> tiles = it->second;      // a copy: several soil rows may share a gridcel (parameter sets)
      const int Nveg = (int)tiles.size();
      double Cv_sum = 0;
      for (const Tile &tl : tiles) Cv_sum += tl.Cv;
      if (Cv_sum > 1.0 || (Cv_sum > 0.99 && Cv_sum < 1.0)) {
        for (Tile &tl : tiles) tl.Cv = tl.Cv / Cv_sum;
        Cv_sum = 1.;
      }
      for (int i = 0; i < Nveg; i++) {
        const Tile &tl = tiles[i];
        int cls = -1;
        for (size_t j = 0; j < lib.size(); j++) if (lib[j].id == tl.cls) cls = (int)j;
        if (cls < 0) { delete P; g_err = "vegetation class not in the library"; return VR_ERR_ARG; }
        // calc_root_fractions
        float root[VR_MAX_LAYERS] = {0, 0, 0}, sum_fract = 0, dum;
        int layer = 0, zone = 0;
        double Lstep = P->depth[(size_t)0 * C + c], Lsum = Lstep, Zsum = 0, Zstep, Zmin_fract, Zmin_depth, Zmax;
        while (zone < RZ) {
          Zstep = (double)tl.zd[zone];
          if ((Zsum + Zstep) <= Lsum && Zsum >= Lsum - Lstep) sum_fract += tl.zf[zone];
          else {
            if (Zsum < Lsum - Lstep) {
              Zmin_depth = Lsum - Lstep;
              Zmin_fract = linear_interp(Zmin_depth, Zsum, Zsum + Zstep, 0, tl.zf[zone]);
            }
            else { Zmin_depth = Zsum; Zmin_fract = 0.; }
            if (Zsum + Zstep <= Lsum) Zmax = Zsum + Zstep;
            else Zmax = Lsum;
            sum_fract += linear_interp(Zmax, Zsum, Zsum + Zstep, 0, tl.zf[zone]) - Zmin_fract;
            (void)Zmin_depth;
          }
          if (Zsum + Zstep < Lsum) { Zsum += Zstep; zone++; }
          else if (Zsum + Zstep == Lsum) {
            Zsum += Zstep; zone++;
            if (layer < L) { root[layer] = sum_fract; sum_fract = 0.; }
            layer++;
            if (layer < L) { Lstep = P->depth[(size_t)layer * C + c]; Lsum += Lstep; }
            else if (layer == L && zone < RZ) {
              Zstep = (double)tl.zd[zone];
              Lstep = Zsum + Zstep - Lsum;
              if (zone < RZ - 1) for (int i2 = zone + 1; i2 < RZ; i2++) Lstep += tl.zd[i2];
              Lsum += Lstep;
            }
          }
          else if (Zsum + Zstep > Lsum) {
            if (layer < L) { root[layer] = sum_fract; sum_fract = 0.; }
            layer++;
            if (layer < L) { Lstep = P->depth[(size_t)layer * C + c]; Lsum += Lstep; }
            else if (layer == L) {
              Lstep = Zsum + Zstep - Lsum;
              if (zone < RZ - 1) for (int i2 = zone + 1; i2 < RZ; i2++) Lstep += tl.zd[i2];
              Lsum += Lstep;
            }
          }
        }
        if (sum_fract > 0 && layer >= L) root[L - 1] += sum_fract;
        else if (sum_fract > 0) root[layer] += sum_fract;
        dum = 0.;
        for (int l = 0; l < L; l++) { if (root[l] < 1.e-4) root[l] = 0.; dum += root[l]; }
        if (dum == 0.0) { delete P; g_err = "root fractions sum to zero"; return VR_ERR_ARG; }
        for (int l = 0; l < L; l++) root[l] /= dum;
        P->tile_class.push_back(cls);
        P->tile_Cv.push_back(tl.Cv);
        for (int l = 0; l < L; l++) P->tile_root.push_back(root[l]);
        for (int j = 0; j < 12; j++) {
          double lai = lib[cls].LAI[j];
          if (tl.has_lai && fo->lai_from_vegparam && tl.LAI[j] != -1) lai = tl.LAI[j];    // NODATA_VH
          P->tile_LAI.push_back(lai);
          P->tile_albedo.push_back(lib[cls].albedo[j]);
          P->tile_vegcover.push_back(1.00);                      // VEGLIB_VEGCOVER FALSE
        }
      }
      if (Cv_sum < 1.) {                                         // bare-soil tile (read_vegparam.c:318-322)
        P->tile_class.push_back((int32_t)lib.size());
        P->tile_Cv.push_back(1.0 - Cv_sum);
        for (int l = 0; l < L; l++) P->tile_root.push_back(0.);
        for (int j = 0; j < 12; j++) { P->tile_LAI.push_back(0.); P->tile_albedo.push_back(0.); P->tile_vegcover.push_back(0.); }
      }
      P->tile_ptr[c + 1] = (int64_t)P->tile_class.size();
    }
  }

  // ---- veglib arrays ----
  for (const VegClass &v : lib) {
    P->l_over.push_back(v.overstory); P->l_rarc.push_back(v.rarc); P->l_rmin.push_back(v.rmin); P->l_wind_h.push_back(v.wind_h);
    P->l_RGL.push_back((double)v.RGL); P->l_rad.push_back(v.rad_atten); P->l_watt.push_back(v.wind_atten); P->l_trunk.push_back(v.trunk_ratio);
    for (int j = 0; j < 12; j++) {
      P->l_rough.push_back(v.roughness[j]); P->l_disp.push_back(v.displacement[j]);
      P->l_LAI.push_back(v.LAI[j]); P->l_alb.push_back(v.albedo[j]);
    }
  }
  vr_veglib &lb = P->lib;
  lb.nclass = (int32_t)lib.size(); lb.overstory = P->l_over.data(); lb.rarc = P->l_rarc.data(); lb.rmin = P->l_rmin.data();
  lb.wind_h = P->l_wind_h.data(); lb.RGL = P->l_RGL.data(); lb.rad_atten = P->l_rad.data(); lb.wind_atten = P->l_watt.data();
  lb.trunk_ratio = P->l_trunk.data(); lb.roughness = P->l_rough.data(); lb.displacement = P->l_disp.data();
  lb.LAI = P->l_LAI.data(); lb.albedo = P->l_alb.data();
  vr_cells &cl = P->cells;
  cl.ncell = C; cl.lat = P->c_lat.data(); cl.elevation = P->elevation.data(); cl.b_infilt = P->b_infilt.data(); cl.Ds = P->Ds.data();
  cl.Dsmax = P->Dsmax.data(); cl.Ws = P->Ws.data(); cl.c_expt = P->c_expt.data(); cl.avg_temp = P->avg_temp.data(); cl.dp = P->dp.data();
  cl.rough = P->rough.data(); cl.snow_rough = P->snow_rough.data(); cl.expt = P->expt.data(); cl.Ksat = P->Ksat.data();
  cl.depth = P->depth.data(); cl.max_moist = P->max_moist.data(); cl.Wcr = P->Wcr.data(); cl.Wpwp = P->Wpwp.data();
  cl.resid_moist = P->resid_moist.data(); cl.init_moist = P->init_moist.data(); cl.bulk_density = P->bulk_density.data();
  cl.soil_density = P->soil_density.data(); cl.bulk_dens_min = P->bulk_dens_min.data(); cl.soil_dens_min = P->soil_dens_min.data();
  cl.quartz = P->quartz.data(); cl.organic = P->organic.data(); cl.porosity = P->porosity.data();
  cl.AreaFract = P->AreaFract.data(); cl.Tfactor = P->Tfactor.data(); cl.Pfactor = P->Pfactor.data();
  P->site_min_tf.resize(C);
  for (int64_t c = 0; c < C; c++) {
    double m = P->Tfactor[c];
    for (int b = 1; b < NB; b++) if (P->Tfactor[(size_t)b * C + c] < m) m = P->Tfactor[(size_t)b * C + c];
    P->site_min_tf[c] = m;
  }
  cl.tile_ptr = P->tile_ptr.data(); cl.tile_class = P->tile_class.data(); cl.tile_Cv = P->tile_Cv.data();
  cl.tile_root = P->tile_root.data(); cl.tile_LAI = P->tile_LAI.data(); cl.tile_albedo = P->tile_albedo.data();
  cl.tile_vegcover = P->tile_vegcover.data();
  cl.zwt_zwt = P->zwt_zwt.data(); cl.zwt_moist = P->zwt_moist.data();
  *out = P;
  return VR_OK;
}

int vr_write_results(const char *result_dir, const char *prefix, int32_t grid_decimal, int64_t ncell, const float *lat,
                     const float *lng, int32_t nrec, const int32_t *year, const int32_t *month, const int32_t *day,
                     const int32_t *hour, int32_t n_cols, const int32_t *cols, const double *out)
{
  if (!result_dir || !prefix || !lat || !lng || !year || !month || !day || !out || !cols || ncell < 1 || nrec < 1 || n_cols < 1) {
    g_err = "null or empty argument";
    return VR_ERR_ARG;
  }
  char fmt[64], name[4096];
  snprintf(fmt, sizeof fmt, "%%s/%%s_%%.%df_%%.%df", grid_decimal, grid_decimal);      // make_in_and_outfiles.c:46-86
  for (int64_t c = 0; c < ncell; c++) {
    snprintf(name, sizeof name, fmt, result_dir, prefix, lat[c], lng[c]);
    FILE *f = fopen(name, "w");
    if (!f) { g_err = std::string("cannot create ") + name; return VR_ERR_ARG; }
    for (int r = 0; r < nrec; r++) {
      if (hour) fprintf(f, "%04i\t%02i\t%02i\t%02i\t", year[r], month[r], day[r], hour[r]);
      else fprintf(f, "%04i\t%02i\t%02i\t", year[r], month[r], day[r]);
      for (int v = 0; v < n_cols; v++) {
        if (v) fprintf(f, "\t ");
        fprintf(f, "%.4f", out[((size_t)cols[v] * nrec + r) * ncell + c]);
      }
      fprintf(f, "\n");
    }
    if (fclose(f) != 0) { g_err = std::string("write failed: ") + name; return VR_ERR_ARG; }
  }
  return VR_OK;
}

// ---- output selection (parse_output_info.c, set_output_defaults.c) and the selection-aware writer (write_data.c) ----
namespace {
struct OutName { const char *name; int id; int per_layer; };
// reference variable name -> first VR_OUT_* id; per_layer: the variable has nlayer elements (ids id .. id + nlayer - 1)
const OutName OUT_TABLE[] = {
  {"OUT_PREC", VR_OUT_PREC, 0}, {"OUT_EVAP", VR_OUT_EVAP, 0}, {"OUT_RUNOFF", VR_OUT_RUNOFF, 0}, {"OUT_BASEFLOW", VR_OUT_BASEFLOW, 0},
  {"OUT_WDEW", VR_OUT_WDEW, 0}, {"OUT_SOIL_LIQ", VR_OUT_SOIL_LIQ0, 1}, {"OUT_SOIL_MOIST", VR_OUT_SOIL_LIQ0, 1},
  {"OUT_NET_SHORT", VR_OUT_NET_SHORT, 0}, {"OUT_R_NET", VR_OUT_R_NET, 0}, {"OUT_EVAP_CANOP", VR_OUT_EVAP_CANOP, 0},
  {"OUT_TRANSP_VEG", VR_OUT_TRANSP_VEG, 0}, {"OUT_EVAP_BARE", VR_OUT_EVAP_BARE, 0}, {"OUT_SUB_CANOP", VR_OUT_SUB_CANOP, 0},
  {"OUT_SUB_SNOW", VR_OUT_SUB_SNOW, 0}, {"OUT_AERO_RESIST", VR_OUT_AERO_RESIST, 0}, {"OUT_SURF_TEMP", VR_OUT_SURF_TEMP, 0},
  {"OUT_ALBEDO", VR_OUT_ALBEDO, 0}, {"OUT_REL_HUMID", VR_OUT_REL_HUMID, 0}, {"OUT_IN_LONG", VR_OUT_IN_LONG, 0},
  {"OUT_AIR_TEMP", VR_OUT_AIR_TEMP, 0}, {"OUT_WIND", VR_OUT_WIND, 0}, {"OUT_SWE", VR_OUT_SWE, 0}, {"OUT_SNOW_DEPTH", VR_OUT_SNOW_DEPTH, 0},
  {"OUT_SNOW_CANOPY", VR_OUT_SNOW_CANOPY, 0}, {"OUT_SNOW_COVER", VR_OUT_SNOW_COVER, 0}, {"OUT_WATER_ERROR", VR_OUT_WATER_ERROR, 0},
  {"OUT_INFLOW", VR_OUT_INFLOW, 0}, {"OUT_SNOW_MELT", VR_OUT_SNOW_MELT, 0}, {"OUT_RAINF", VR_OUT_RAINF, 0}, {"OUT_SNOWF", VR_OUT_SNOWF, 0},
  {"OUT_NET_LONG", VR_OUT_NET_LONG, 0}, {"OUT_ASAT", VR_OUT_ASAT, 0}, {"OUT_SOIL_WET", VR_OUT_SOIL_WET, 0}, {"OUT_ROOTMOIST", VR_OUT_ROOTMOIST, 0},
  {"OUT_GRND_FLUX", VR_OUT_GRND_FLUX, 0}, {"OUT_DELTAH", VR_OUT_DELTAH, 0}, {"OUT_AERO_COND", VR_OUT_AERO_COND, 0},
  {"OUT_SHORTWAVE", VR_OUT_SHORTWAVE, 0}, {"OUT_LONGWAVE", VR_OUT_LONGWAVE, 0}, {"OUT_DELSOILMOIST", VR_OUT_DELSOILMOIST, 0},
  {"OUT_DELSWE", VR_OUT_DELSWE, 0}, {"OUT_DELINTERCEPT", VR_OUT_DELINTERCEPT, 0}, {"OUT_SNOW_SURF_TEMP", VR_OUT_SNOW_SURF_TEMP, 0},
  {"OUT_SNOW_PACK_TEMP", VR_OUT_SNOW_PACK_TEMP, 0}, {"OUT_SALBEDO", VR_OUT_SALBEDO, 0}, {"OUT_PET_SATSOIL", VR_OUT_PET_SATSOIL, 0},
  {"OUT_PET_H2OSURF", VR_OUT_PET_H2OSURF, 0}, {"OUT_PET_SHORT", VR_OUT_PET_SHORT, 0}, {"OUT_PET_TALL", VR_OUT_PET_TALL, 0},
  {"OUT_PET_NATVEG", VR_OUT_PET_NATVEG, 0}, {"OUT_PET_VEGNOCR", VR_OUT_PET_VEGNOCR, 0}, {"OUT_ZWT", VR_OUT_ZWT, 0},
  {"OUT_ZWT_LUMPED", VR_OUT_ZWT_LUMPED, 0}, {"OUT_DENSITY", VR_OUT_DENSITY, 0}, {"OUT_PRESSURE", VR_OUT_PRESSURE, 0},
  {"OUT_VP", VR_OUT_VP, 0}, {"OUT_VPD", VR_OUT_VPD, 0}, {"OUT_QAIR", VR_OUT_QAIR, 0},
};
const OutName *find_out(const char *name)
{
  for (const OutName &o : OUT_TABLE)
    if (!strcmp(o.name, name)) return &o;
  return nullptr;
}
bool outsel_add(vr_outsel *s, int f, const char *name, const char *format, int type, double mult, int nlayer)
{
  const OutName *o = find_out(name);
  if (!o) {
    g_err = std::string("set_output_var: \"") + name + "\" is not an output variable of the water-balance path of this library";
    return false;
  }
  const int ne = o->per_layer ? nlayer : 1;
  // format / type / mult belong to the VARIABLE (out_data[varid]): a later line for the same variable changes every column of it
  for (int ff = 0; ff < s->nfiles; ff++)
    for (int c = 0; c < s->ncols[ff]; c++)
      if (s->var[ff][c] >= o->id && s->var[ff][c] < o->id + ne) {
        if (strcmp(format, "*")) snprintf(s->format[ff][c], 12, "%s", format);
        if (type) s->type[ff][c] = type;
        if (mult != 0) s->mult[ff][c] = mult;
      }
  s->nvars[f]++;
  for (int e = 0; e < ne; e++) {
    int &n = s->ncols[f];
    if (n >= VR_MAX_OUTCOLS) { g_err = "too many output columns in one file"; return false; }
    if (ne > 1) snprintf(s->name[f][n], 32, "%s_%d", name, e);      // write_header.c:219-222
    else snprintf(s->name[f][n], 32, "%s", name);
    s->var[f][n] = o->id + e;
    snprintf(s->format[f][n], 12, "%s", strcmp(format, "*") ? format : "%.4f");
    s->type[f][n] = type ? type : VR_OUT_TYPE_FLOAT;
    s->mult[f][n] = mult != 0 ? mult : 1.0;
    n++;
  }
  return true;
}
}  // namespace

int vr_outsel_read(const char *global_path, int32_t nlayer, vr_outsel *s)
{
  if (!global_path || !s || nlayer < 1 || nlayer > 3) { g_err = "null or invalid argument"; return VR_ERR_ARG; }
  memset(s, 0, sizeof *s);
  FILE *gp = fopen(global_path, "r");
  if (!gp) { g_err = std::string("cannot open ") + global_path; return VR_ERR_ARG; }
  char line[4096], key[256], a1[256], a2[256], a3[256], a4[256];
  int declared = 0, f = -1, output_force = 0;
  while (fgets(line, sizeof line, gp)) {
    if (line[0] == '#' || line[0] == '\n' || line[0] == '\0') continue;
    if (sscanf(line, "%255s", key) != 1) continue;
    if (!strcasecmp(key, "OUTPUT_FORCE")) {
      a1[0] = 0;
      sscanf(line, "%*s %255s", a1);
      output_force = !strcasecmp(a1, "TRUE");
    }
    else if (!strcasecmp(key, "N_OUTFILES")) {
      if (sscanf(line, "%*s %d", &declared) != 1 || declared < 1 || declared > VR_MAX_OUTFILES) {
        fclose(gp); g_err = "N_OUTFILES out of range (1..8)"; return VR_ERR_ARG;
      }
      memset(s, 0, sizeof *s);
      f = -1;
    }
    else if (!strcasecmp(key, "OUTFILE")) {
      if (!declared) { fclose(gp); g_err = "\"N_OUTFILES\" must be specified before you can specify \"OUTFILE\""; return VR_ERR_ARG; }
      if (++f >= declared) { fclose(gp); g_err = "more OUTFILE lines than N_OUTFILES"; return VR_ERR_ARG; }
      int nv = 0;
      a1[0] = 0;
      sscanf(line, "%*s %63s %d", a1, &nv);
      snprintf(s->prefix[f], 64, "%s", a1);
      s->nfiles = f + 1;
    }
    else if (!strcasecmp(key, "OUTVAR")) {
      if (f < 0) { fclose(gp); g_err = "\"OUTFILE\" must be specified before you can specify \"OUTVAR\""; return VR_ERR_ARG; }
      a1[0] = a2[0] = a3[0] = a4[0] = 0;
      sscanf(line, "%*s %255s %255s %255s %255s", a1, a2, a3, a4);
      int type = VR_OUT_TYPE_DEFAULT;
      double mult = 0;
      const char *fmt = "*";
      if (a2[0]) {
        fmt = a2;
        if (!strcasecmp(a3, "OUT_TYPE_USINT")) type = VR_OUT_TYPE_USINT;
        else if (!strcasecmp(a3, "OUT_TYPE_SINT")) type = VR_OUT_TYPE_SINT;
        else if (!strcasecmp(a3, "OUT_TYPE_FLOAT")) type = VR_OUT_TYPE_FLOAT;
        else if (!strcasecmp(a3, "OUT_TYPE_DOUBLE")) type = VR_OUT_TYPE_DOUBLE;
        mult = !strcmp(a4, "*") ? 0 : (double)(float)atof(a4);
      }
      if (!outsel_add(s, f, a1, fmt, type, mult, nlayer)) { fclose(gp); return VR_ERR_ARG; }
    }
  }
  fclose(gp);
  if (!declared && output_force) {       // set_output_defaults.c:40-62: the forcing file
    s->nfiles = 1;
    snprintf(s->prefix[0], 64, "full_data");
    const struct { const char *n; int type; double mult; } fd[] = {
      {"OUT_PREC", VR_OUT_TYPE_USINT, 40}, {"OUT_AIR_TEMP", VR_OUT_TYPE_SINT, 100}, {"OUT_SHORTWAVE", VR_OUT_TYPE_USINT, 50},
      {"OUT_LONGWAVE", VR_OUT_TYPE_USINT, 80}, {"OUT_DENSITY", VR_OUT_TYPE_USINT, 100}, {"OUT_PRESSURE", VR_OUT_TYPE_USINT, 100},
      {"OUT_VP", VR_OUT_TYPE_SINT, 100}, {"OUT_WIND", VR_OUT_TYPE_USINT, 100}};
    for (const auto &v : fd) outsel_add(s, 0, v.n, "%.4f", v.type, v.mult, nlayer);
  }
  else if (!declared) {       // set_output_defaults.c:66-162, water-balance mode
    s->nfiles = 2;
    snprintf(s->prefix[0], 64, "fluxes");
    snprintf(s->prefix[1], 64, "snow");
    const char *flux[] = {"OUT_PREC", "OUT_EVAP", "OUT_RUNOFF", "OUT_BASEFLOW", "OUT_WDEW", "OUT_SOIL_LIQ", "OUT_NET_SHORT", "OUT_R_NET",
                          "OUT_EVAP_CANOP", "OUT_TRANSP_VEG", "OUT_EVAP_BARE", "OUT_SUB_CANOP", "OUT_SUB_SNOW", "OUT_AERO_RESIST",
                          "OUT_SURF_TEMP", "OUT_ALBEDO", "OUT_REL_HUMID", "OUT_IN_LONG", "OUT_AIR_TEMP", "OUT_WIND"};
    const char *snow[] = {"OUT_SWE", "OUT_SNOW_DEPTH", "OUT_SNOW_CANOPY", "OUT_SNOW_COVER"};
    for (const char *n : flux) outsel_add(s, 0, n, "%.4f", VR_OUT_TYPE_FLOAT, 1, nlayer);
    for (const char *n : snow) outsel_add(s, 1, n, "%.4f", VR_OUT_TYPE_FLOAT, 1, nlayer);
  }
  else if (s->nfiles != declared) { g_err = "fewer OUTFILE lines than N_OUTFILES"; return VR_ERR_ARG; }
  s->no_date = output_force;
  return VR_OK;
}

int vr_outsel_step_vars(const vr_outsel *s, int32_t mode, int32_t *vars, int32_t *n)
{
  if (!s || !vars || !n) { g_err = "null argument"; return VR_ERR_ARG; }
  const int aggregate = mode & 1, alma = mode & 2;
  *n = 0;
  for (int f = 0; f < s->nfiles; f++)
    for (int c = 0; c < s->ncols[f]; c++) {
      int v = s->var[f][c];
      if (aggregate && v == VR_OUT_AERO_RESIST) v = VR_OUT_AERO_COND;
      bool seen = false;
      for (int k = 0; k < *n; k++) seen = seen || vars[k] == v;
      if (seen) continue;
      if (*n >= VR_MAX_OUT) { g_err = "more than VR_MAX_OUT distinct output columns"; return VR_ERR_ARG; }
      vars[(*n)++] = v;
    }
  if (alma) {           // OUT_SUB_SNOW's ALMA value includes OUT_SUB_CANOP (put_data.c:742)
    bool snow = false, canop = false;
    for (int k = 0; k < *n; k++) { snow = snow || vars[k] == VR_OUT_SUB_SNOW; canop = canop || vars[k] == VR_OUT_SUB_CANOP; }
    if (snow && !canop) {
      if (*n >= VR_MAX_OUT) { g_err = "more than VR_MAX_OUT distinct output columns"; return VR_ERR_ARG; }
      vars[(*n)++] = VR_OUT_SUB_CANOP;
    }
  }
  return VR_OK;
}

int vr_write_results_sel(const char *result_dir, const vr_outsel *s, int32_t file, int32_t grid_decimal, int64_t ncell,
                         const float *lat, const float *lng, int32_t nrec, const int32_t *year, const int32_t *month,
                         const int32_t *day, const int32_t *hour, int32_t binary, const int32_t *rows, const double *out)
{
  if (!result_dir || !s || !lat || !lng || !year || !month || !day || !rows || !out || file < 0 || file >= s->nfiles || ncell < 1 ||
      nrec < 0) {            // nrec == 0 (everything inside SKIPYEAR): the files are created empty, as the reference leaves them
    g_err = "null or empty argument";
    return VR_ERR_ARG;
  }
  char fmt[64], name[4096];
  snprintf(fmt, sizeof fmt, "%%s/%%s_%%.%df_%%.%df", grid_decimal, grid_decimal);      // make_in_and_outfiles.c:46-86
  const int nc = s->ncols[file];
  for (int64_t c = 0; c < ncell; c++) {
    snprintf(name, sizeof name, fmt, result_dir, s->prefix[file], lat[c], lng[c]);
    FILE *f = fopen(name, binary ? "wb" : "w");
    if (!f) { g_err = std::string("cannot create ") + name; return VR_ERR_ARG; }
    if (s->header) {      // PRT_HEADER, write_header.c:46-307
      const int ndate = s->no_date ? 0 : (hour ? 4 : 3);
      if (binary) {
        static const char *DATE[4] = {"YEAR", "MONTH", "DAY", "HOUR"};
        unsigned short id = 0xFFFF, nb1 = sizeof(unsigned short) + 6 * sizeof(int) + 2, nb2 = sizeof(unsigned short);
        for (int k = 0; k < ndate; k++) nb2 += 1 + strlen(DATE[k]) + 1 + sizeof(float);
        for (int v = 0; v < nc; v++) nb2 += 1 + strlen(s->name[file][v]) + 1 + sizeof(float);
        const unsigned short nb = 4 * sizeof(unsigned short) + sizeof(unsigned short) + nb1 + nb2;
        for (int k = 0; k < 4; k++) fwrite(&id, sizeof id, 1, f);
        fwrite(&nb, sizeof nb, 1, f);
        fwrite(&nb1, sizeof nb1, 1, f);
        const int gl[6] = {s->hdr_nrecs, s->hdr_out_dt, s->hdr_start[0], s->hdr_start[1], s->hdr_start[2], s->hdr_start[3]};
        fwrite(gl, sizeof(int), 6, f);
        const char alma = s->hdr_alma ? 1 : 0, nvars = (char)(s->nvars[file] + ndate);
        fwrite(&alma, 1, 1, f);
        fwrite(&nvars, 1, 1, f);
        fwrite(&nb2, sizeof nb2, 1, f);
        auto var = [&](const char *nm, char type, float mult) {
          const char len = (char)strlen(nm);
          fwrite(&len, 1, 1, f); fwrite(nm, 1, len, f); fwrite(&type, 1, 1, f); fwrite(&mult, sizeof(float), 1, f);
        };
        for (int k = 0; k < ndate; k++) var(DATE[k], (char)VR_OUT_TYPE_INT, 1.f);
        for (int v = 0; v < nc; v++) var(s->name[file][v], (char)s->type[file][v], (float)s->mult[file][v]);
      }
      else {
        fprintf(f, "# NRECS: %d\n# DT: %d\n# STARTDATE: %04d-%02d-%02d %02d:00:00\n# ALMA_OUTPUT: %d\n# NVARS: %d\n# ", s->hdr_nrecs,
                s->hdr_out_dt, s->hdr_start[0], s->hdr_start[1], s->hdr_start[2], s->hdr_start[3], s->hdr_alma ? 1 : 0,
                s->nvars[file] + ndate);
        if (!s->no_date) fprintf(f, hour ? "YEAR\tMONTH\tDAY\tHOUR\t" : "YEAR\tMONTH\tDAY\t");
        for (int v = 0; v < nc; v++) fprintf(f, "%s%s", v ? "\t " : "", s->name[file][v]);
        fprintf(f, "\n");
      }
    }
    for (int r = 0; r < nrec; r++) {
      if (binary) {       // write_data.c:103-193
        const int date[4] = {year[r], month[r], day[r], hour ? hour[r] : 0};
        if (!s->no_date) fwrite(date, sizeof(int), hour ? 4 : 3, f);
        for (int v = 0; v < nc; v++) {
          const double x = out[((size_t)rows[v] * nrec + r) * ncell + c] * s->mult[file][v];      // put_data.c:766-772
          switch (s->type[file][v]) {
            case VR_OUT_TYPE_CHAR: { const char t = (char)x; fwrite(&t, sizeof t, 1, f); break; }
            case VR_OUT_TYPE_SINT: { const short t = (short)x; fwrite(&t, sizeof t, 1, f); break; }
            case VR_OUT_TYPE_USINT: { const unsigned short t = (unsigned short)x; fwrite(&t, sizeof t, 1, f); break; }
            case VR_OUT_TYPE_INT: { const int t = (int)x; fwrite(&t, sizeof t, 1, f); break; }
            case VR_OUT_TYPE_DOUBLE: fwrite(&x, sizeof x, 1, f); break;
            default: { const float t = (float)x; fwrite(&t, sizeof t, 1, f); break; }
          }
        }
      }
      else {              // write_data.c:196-226
        if (s->no_date) { }
        else if (hour) fprintf(f, "%04i\t%02i\t%02i\t%02i\t", year[r], month[r], day[r], hour[r]);
        else fprintf(f, "%04i\t%02i\t%02i\t", year[r], month[r], day[r]);
        for (int v = 0; v < nc; v++) {
          if (v) fprintf(f, "\t ");
          fprintf(f, s->format[file][v], out[((size_t)rows[v] * nrec + r) * ncell + c]);
        }
        fprintf(f, "\n");
      }
    }
    if (fclose(f) != 0) { g_err = std::string("write failed: ") + name; return VR_ERR_ARG; }
  }
  return VR_OK;
}

// ---- global parameter file (get_global_param.c) ----
static bool leapyr(int y) { return ((y % 4 == 0) && (y % 100 != 0)) || (y % 400 == 0); }
static void next_step(int &year, int &month, int &day, int &hr, int &jday, int dt)     // make_dmy.c:175-205
{
  int days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  hr += dt;
  if (hr >= 24) {
    hr = 0; day += 1; jday += 1;
    days[1] = leapyr(year) ? 29 : 28;
    if (day > days[month - 1]) {
      day = 1; month += 1;
      if (month == 13) { month = 1; jday = 1; year += 1; }
    }
  }
}

// Options that change the results (physics or file contents) and of which this library implements the reference's default
// only (initialize_global.c:146-211; parsed at get_global_param.c:230-720): key, the accepted values.
static bool other_option_unsupported(const char *key, const char *val)
{
  static const struct { const char *key, *ok1, *ok2; } T[] = {
    {"PRT_SNOW_BAND", "FALSE", nullptr},
    {"ARC_SOIL", "FALSE", nullptr}, {"AERO_RESIST_CANSNOW", "AR_406_FULL", nullptr}, {"SNOW_ALBEDO", "USACE", nullptr},
    {"SNOW_DENSITY", "DENS_BRAS", nullptr}, {"GRND_FLUX_TYPE", "GF_410", nullptr}, {"QUICK_FLUX", "TRUE", nullptr},
    {"TFALLBACK", "TRUE", nullptr}, {"SPATIAL_SNOW", "FALSE", nullptr}, {"SPATIAL_FROST", "FALSE", nullptr},
    {"RC_MODE", "RC_JARVIS", nullptr}, {"SHARE_LAYER_MOIST", "TRUE", nullptr}, {"VEGLIB_VEGCOVER", "FALSE", nullptr},
    {"VEGLIB_PHOTO", "FALSE", nullptr}, {"VEGPARAM_ALB", "FALSE", nullptr}, {"VEGPARAM_VEGCOVER", "FALSE", nullptr},
    {"ALB_SRC", "FROM_VEGLIB", "ALB_FROM_VEGLIB"}, {"VEGCOVER_SRC", "FROM_VEGLIB", "VEGCOVER_FROM_VEGLIB"},
    {"CLOSE_ENERGY", "FALSE", nullptr}, {"LAKE_PROFILE", "FALSE", nullptr}, {"CONTINUEONERROR", "TRUE", nullptr},
  };
  for (const auto &e : T)
    if (!strcasecmp(key, e.key)) return strcasecmp(val, e.ok1) != 0 && !(e.ok2 && strcasecmp(val, e.ok2) == 0);
  return false;
}

// which: 0 = the description of FORCING1 is kept (vr_global_read), 1 = that of FORCING2, in the same fields (vr_global_read_forcing2)
static int global_read_file(const char *path, vr_global *g, int which)
{
  if (!path || !g) { g_err = "null argument"; return VR_ERR_ARG; }
  FILE *f = fopen(path, "r");
  if (!f) { g_err = std::string("cannot open ") + path; return VR_ERR_ARG; }
  memset(g, 0, sizeof *g);
  // initialize_global.c:146-211
  g->nlayer = 3; g->time_step = -99999; g->snow_step = 1; g->max_snow_temp = 0.5; g->min_rain_temp = -0.5; g->wind_h = 10.0;
  g->min_wind_speed = 0.1; g->measure_h = 2.0; g->snow_band = 1; g->root_zones = -99999; g->grid_decimal = 2;
  g->vp_interp = 1; g->vp_iter = 1; g->mtclim_swe_corr = 0; g->lw_type = 5; g->lw_cloud = 1; g->plapse = 1; g->sw_prec_thresh = 0.;
  g->starthour = 0; g->nrecs = -99999; g->force_dt = -99999; g->force_ascii = 0; g->force_little_endian = 1;
  int endyear = -99999, endmonth = -99999, endday = -99999, file_num = -1, field = -1;
  auto flag = [](const std::vector<std::string> &t) { return t.size() > 1 && strcasecmp(t[1].c_str(), "TRUE") == 0; };
  auto copy = [](char *dst, const std::string &v) { snprintf(dst, 512, "%s", v.c_str()); };
  std::string line;
  while (read_line(f, line)) {
    std::vector<std::string> t = tokens(line);
    if (t.empty() || t[0][0] == '#') continue;
    const char *k = t[0].c_str();
    auto I = [&](int d = 0) { return t.size() > 1 ? atoi(t[1].c_str()) : d; };
    auto D = [&](double d = 0) { return t.size() > 1 ? strtod(t[1].c_str(), nullptr) : d; };
    if (!strcasecmp(k, "NLAYER")) g->nlayer = I();
    else if (!strcasecmp(k, "TIME_STEP")) g->time_step = I();
    else if (!strcasecmp(k, "SNOW_STEP")) g->snow_step = I();
    else if (!strcasecmp(k, "STARTYEAR")) g->startyear = I();
    else if (!strcasecmp(k, "STARTMONTH")) g->startmonth = I();
    else if (!strcasecmp(k, "STARTDAY")) g->startday = I();
    else if (!strcasecmp(k, "STARTHOUR")) g->starthour = I();
    else if (!strcasecmp(k, "NRECS")) g->nrecs = I();
    else if (!strcasecmp(k, "ENDYEAR")) endyear = I();
    else if (!strcasecmp(k, "ENDMONTH")) endmonth = I();
    else if (!strcasecmp(k, "ENDDAY")) endday = I();
    else if (!strcasecmp(k, "FULL_ENERGY")) g->full_energy = flag(t);
    else if (!strcasecmp(k, "FROZEN_SOIL")) g->frozen_soil = flag(t);
    else if (!strcasecmp(k, "BLOWING")) g->blowing = flag(t);
    else if (!strcasecmp(k, "CORRPREC")) g->corrprec = flag(t);
    else if (!strcasecmp(k, "DIST_PRCP")) g->dist_prcp = flag(t);
    else if (!strcasecmp(k, "CARBON")) g->carbon = flag(t);
    else if (!strcasecmp(k, "LAKES")) g->lakes = (t.size() > 1 && strcasecmp(t[1].c_str(), "FALSE") != 0);
    else if (!strcasecmp(k, "COMPUTE_TREELINE")) g->compute_treeline = (t.size() > 1 && strcasecmp(t[1].c_str(), "FALSE") != 0);
    else if (!strcasecmp(k, "ORGANIC_FRACT")) g->organic_fract = flag(t);
    else if (!strcasecmp(k, "JULY_TAVG_SUPPLIED")) g->july_tavg_supplied = flag(t);
    else if (!strcasecmp(k, "ALMA_INPUT")) g->alma_input = flag(t);
    else if (!strcasecmp(k, "INIT_STATE")) {
      g->init_state = (t.size() > 1 && strcasecmp(t[1].c_str(), "FALSE") != 0);
      if (g->init_state) copy(g->init_state_file, t[1]);
    }
    else if (!strcasecmp(k, "STATENAME") && t.size() > 1) copy(g->statename, t[1]);
    else if (!strcasecmp(k, "STATEYEAR")) g->stateyear = I();
    else if (!strcasecmp(k, "STATEMONTH")) g->statemonth = I();
    else if (!strcasecmp(k, "STATEDAY")) g->stateday = I();
    else if (!strcasecmp(k, "BINARY_STATE_FILE")) g->binary_state_file = !(t.size() > 1 && !strcasecmp(t[1].c_str(), "FALSE"));
    else if (!strcasecmp(k, "MIN_WIND_SPEED")) g->min_wind_speed = (double)strtof(t.size() > 1 ? t[1].c_str() : "0", nullptr);
    else if (!strcasecmp(k, "VP_INTERP")) g->vp_interp = flag(t);
    else if (!strcasecmp(k, "VP_ITER")) {
      // get_global_param.c:405-412: four ifs and one else -- every value but VP_ITER_CONVERGE also sets VP_INTERP to 1
      const char *v = t.size() > 1 ? t[1].c_str() : "";
      if (!strcasecmp(v, "VP_ITER_NONE")) g->vp_iter = 0;
      if (!strcasecmp(v, "VP_ITER_ALWAYS")) g->vp_iter = 1;
      if (!strcasecmp(v, "VP_ITER_ANNUAL")) g->vp_iter = 2;
      if (!strcasecmp(v, "VP_ITER_CONVERGE")) g->vp_iter = 3;
      else g->vp_interp = 1;
    }
    else if (!strcasecmp(k, "MTCLIM_SWE_CORR")) g->mtclim_swe_corr = flag(t);
    else if (!strcasecmp(k, "PLAPSE")) g->plapse = !(t.size() > 1 && !strcasecmp(t[1].c_str(), "FALSE"));
    else if (!strcasecmp(k, "SW_PREC_THRESH")) g->sw_prec_thresh = t.size() > 1 ? atof(t[1].c_str()) : 0.;
    else if (!strcasecmp(k, "LW_CLOUD")) g->lw_cloud = (t.size() > 1 && !strcasecmp(t[1].c_str(), "LW_CLOUD_DEARDORFF")) ? 1 : 0;
    else if (!strcasecmp(k, "LW_TYPE")) {
      static const char *names[6] = {"LW_TVA", "LW_ANDERSON", "LW_BRUTSAERT", "LW_SATTERLUND", "LW_IDSO", "LW_PRATA"};
      for (int i = 0; i < 6; i++) if (t.size() > 1 && !strcasecmp(t[1].c_str(), names[i])) g->lw_type = i;
    }
    else if (!strcasecmp(k, "MIN_RAIN_TEMP")) g->min_rain_temp = D();
    else if (!strcasecmp(k, "MAX_SNOW_TEMP")) g->max_snow_temp = D();
    else if (!strcasecmp(k, "WIND_H")) g->wind_h = D();
    else if (!strcasecmp(k, "MEASURE_H")) g->measure_h = D();
    else if (!strcasecmp(k, "GRID_DECIMAL")) g->grid_decimal = I();
    else if (!strcasecmp(k, "SOIL") && t.size() > 1) copy(g->soil, t[1]);
    else if (!strcasecmp(k, "VEGLIB") && t.size() > 1) copy(g->veglib, t[1]);
    else if (!strcasecmp(k, "VEGPARAM") && t.size() > 1) copy(g->vegparam, t[1]);
    else if (!strcasecmp(k, "ROOT_ZONES")) g->root_zones = I();
    else if (!strcasecmp(k, "VEGPARAM_LAI")) g->vegparam_lai = flag(t);
    else if (!strcasecmp(k, "LAI_SRC") || !strcasecmp(k, "GLOBAL_LAI"))
      g->lai_from_vegparam = (t.size() > 1 && (!strcasecmp(t[1].c_str(), "FROM_VEGPARAM") || !strcasecmp(t[1].c_str(), "LAI_FROM_VEGPARAM") ||
                                               (!strcasecmp(k, "GLOBAL_LAI") && flag(t))));
    else if (!strcasecmp(k, "SNOW_BAND")) { g->snow_band = I(1); if (t.size() > 2) copy(g->snowband, t[2]); }
    else if (!strcasecmp(k, "RESULT_DIR") && t.size() > 1) copy(g->result_dir, t[1]);
    else if (!strcasecmp(k, "SKIPYEAR")) g->skipyear = I();
    else if (!strcasecmp(k, "OUT_STEP")) g->out_step = I();
    else if (!strcasecmp(k, "BINARY_OUTPUT")) g->binary_output = flag(t);
    else if (!strcasecmp(k, "COMPRESS")) g->compress = flag(t);
    else if (!strcasecmp(k, "PRT_HEADER")) g->prt_header = flag(t);
    else if (!strcasecmp(k, "ALMA_OUTPUT")) g->alma_output = flag(t);
    else if (!strcasecmp(k, "OUTPUT_FORCE")) g->output_force = flag(t);
    else if (!strcasecmp(k, "MOISTFRACT")) g->moistfract = flag(t);
    else if (!strcasecmp(k, "BASEFLOW")) g->baseflow_nijssen = (t.size() > 1 && !strcasecmp(t[1].c_str(), "NIJSSEN2001"));
    else if (!strcasecmp(k, "OUTFILE")) g->n_outfile_lines++;
    else if (!strcasecmp(k, "OUTVAR")) g->n_outvar_lines++;
    else if (!strcasecmp(k, "FORCING1") && t.size() > 1) {
      file_num = 0;
      if (which == 0) { copy(g->forcing1, t[1]); field = -1; }
    }
    else if (other_option_unsupported(k, t.size() > 1 ? t[1].c_str() : "")) {
      if (!g->other_unsupported[0]) snprintf(g->other_unsupported, sizeof g->other_unsupported, "%s %s", k, t.size() > 1 ? t[1].c_str() : "");
    }
    else if (!strcasecmp(k, "FORCING2")) {     // the second forcing file: its description follows (get_global_param.c:441-470)
      file_num = 1;
      g->forcing2 = (t.size() > 1 && strcasecmp(t[1].c_str(), "FALSE") != 0 && strcasecmp(t[1].c_str(), "MISSING") != 0);
      if (which == 1 && g->forcing2) { copy(g->forcing1, t[1]); field = -1; }
    }
    else if (file_num == which) {      // description of the forcing file this call keeps (get_force_type.c)
      if (!strcasecmp(k, "FORCE_FORMAT")) g->force_ascii = (t.size() > 1 && !strcasecmp(t[1].c_str(), "ASCII"));
      else if (!strcasecmp(k, "FORCE_ENDIAN")) g->force_little_endian = !(t.size() > 1 && !strcasecmp(t[1].c_str(), "BIG"));
      else if (!strcasecmp(k, "N_TYPES")) g->n_force_types = I();
      else if (!strcasecmp(k, "FORCE_DT")) g->force_dt = I();
      else if (!strcasecmp(k, "FORCEYEAR")) g->forceyear = I();
      else if (!strcasecmp(k, "FORCEMONTH")) g->forcemonth = I();
      else if (!strcasecmp(k, "FORCEDAY")) g->forceday = I();
      else if (!strcasecmp(k, "FORCEHOUR")) g->forcehour = I();
      else if (!strcasecmp(k, "FORCE_TYPE") && t.size() > 1) {
        field++;
        if (field >= VR_MAX_FORCE_TYPES) { fclose(f); g_err = "too many FORCE_TYPE lines"; return VR_ERR_ARG; }
        snprintf(g->force_type[field], 16, "%s", t[1].c_str());
        g->force_signed[field] = 1; g->force_multiplier[field] = 1.0;
        if (t.size() > 3) {             // binary: FORCE_TYPE <name> SIGNED|UNSIGNED <multiplier>
          g->force_signed[field] = strcasecmp(t[2].c_str(), "UNSIGNED") != 0;
          g->force_multiplier[field] = strtod(t[3].c_str(), nullptr);
        }
      }
    }
  }
  fclose(f);
  if (g->time_step == -99999) { g_err = "TIME_STEP missing from the global parameter file"; return VR_ERR_ARG; }
  if (g->startyear <= 0 || g->startmonth < 1 || g->startmonth > 12 || g->startday < 1) { g_err = "simulation start date missing or invalid"; return VR_ERR_ARG; }
  if (g->nrecs == -99999) {            // make_dmy.c:52-90: count the steps up to the day after the end date
    if (endyear == -99999 || endmonth == -99999 || endday == -99999) { g_err = "neither NRECS nor ENDYEAR/ENDMONTH/ENDDAY given"; return VR_ERR_ARG; }
    int days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    days[1] = leapyr(endyear) ? 29 : 28;
    if (endday < days[endmonth - 1]) endday++;
    else { endday = 1; endmonth++; if (endmonth > 12) { endmonth = 1; endyear++; } }
    int y = g->startyear, mo = g->startmonth, d = g->startday, h = g->starthour, j = 0, n = 0;
    for (;;) {
      next_step(y, mo, d, h, j, g->time_step);
      n++;
      if (y == endyear && mo == endmonth && d == endday) break;
      if (n > 100000000) { g_err = "end date is not after the start date"; return VR_ERR_ARG; }
    }
    g->nrecs = n;
  }
  if (g->n_force_types != 0 && field + 1 != g->n_force_types) { g_err = "N_TYPES does not match the number of FORCE_TYPE lines"; return VR_ERR_ARG; }
  g->n_force_types = field + 1;
  return VR_OK;
}

int vr_global_read(const char *path, vr_global *g) { return global_read_file(path, g, 0); }

int vr_global_read_forcing2(const char *path, vr_global *g2)
{
  const int rc = global_read_file(path, g2, 1);
  if (rc == VR_OK && !g2->forcing2) { g_err = "the global parameter file names no second forcing file"; return VR_ERR_ARG; }
  return rc;
}

const char *vr_global_unsupported(const vr_global *g)
{
  if (!g) return "null";
  if (g->full_energy) return "FULL_ENERGY TRUE";
  if (g->frozen_soil) return "FROZEN_SOIL TRUE";
  if (g->lakes) return "LAKES";
  if (g->carbon) return "CARBON TRUE";
  if (g->blowing) return "BLOWING TRUE";
  if (g->corrprec) return "CORRPREC TRUE";
  if (g->compute_treeline) return "COMPUTE_TREELINE";
  if (g->dist_prcp) return "DIST_PRCP TRUE";
  if (g->time_step < 24 && g->snow_step != g->time_step)       // get_global_param.c:761-765
    return "If the model step is smaller than daily, the snow model should run at the same time step as the rest of the model.";
  if (g->snow_step < 1 || g->snow_step > g->time_step || g->time_step % g->snow_step)
    return "SNOW_STEP should be <= TIME_STEP and divide TIME_STEP evenly";
  if (g->nlayer < 2 || g->nlayer > VR_MAX_LAYERS) return "NLAYER outside 2..3";
  if (g->other_unsupported[0]) return g->other_unsupported;
  if (g->out_step > 0 && (g->out_step < g->time_step || g->out_step % g->time_step || g->out_step > 24))
    return "OUT_STEP must be a multiple of TIME_STEP and at most 24 (get_global_param.c:754-760)";
  return nullptr;
}

int vr_make_dmy(const vr_global *g, int32_t *year, int32_t *month, int32_t *day, int32_t *hour, int32_t *day_in_year)
{
  if (!g || !year || !month || !day || !hour || !day_in_year || g->nrecs < 1) { g_err = "null argument"; return VR_ERR_ARG; }
  int days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int y = g->startyear, mo = g->startmonth, d = g->startday, h = g->starthour, j = d;
  days[1] = leapyr(y) ? 29 : 28;
  for (int i = 0; i < mo - 1; i++) j += days[i];
  for (int i = 0; i < g->nrecs; i++) {
    hour[i] = h; day[i] = d; month[i] = mo; year[i] = y; day_in_year[i] = j;
    next_step(y, mo, d, h, j, g->time_step);
  }
  return VR_OK;
}

int32_t vr_force_skip(const vr_global *g)
{
  if (!g || g->forceyear <= 0) return 0;
  int days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int y = g->forceyear, mo = g->forcemonth, d = g->forceday, h = g->forcehour, j = d, skip = 0;
  days[1] = leapyr(y) ? 29 : 28;
  for (int i = 0; i < mo - 1; i++) j += days[i];
  int sj = g->startday;
  days[1] = leapyr(g->startyear) ? 29 : 28;
  for (int i = 0; i < g->startmonth - 1; i++) sj += days[i];
  while (y < g->startyear || (y == g->startyear && j < sj) || (y == g->startyear && j == sj && h < g->starthour)) {
    next_step(y, mo, d, h, j, g->time_step);          // the reference steps with the MODEL dt here (make_dmy.c:139-146)
    skip++;
  }
  return skip;
}

int32_t vr_output_skip(const vr_global *g, const int32_t *year)
{
  if (!g || !year || g->skipyear <= 0 || g->time_step <= 0) return 0;
  int64_t skip = 0;
  for (int i = 0; i < g->skipyear; i++) {
    const bool leap = skip < g->nrecs && leapyr(year[skip]);
    skip += (leap ? 366 : 365) * 24 / g->time_step;
    if (skip > (int64_t)g->nrecs) { skip = g->nrecs; break; }
  }
  return (int32_t)skip;
}

int64_t vr_read_forcing_file(const vr_global *g, const char *path, int64_t nrec, double *data)
{
  if (!g || !path || !data || nrec < 1 || g->n_force_types < 1) { g_err = "null or empty argument"; return VR_ERR_ARG; }
  const int nf = g->n_force_types;
  // read_atmos_data.c:91-98: vr_force_skip() counts MODEL steps; the file holds one line per FORCE_DT hours
  const int fdt = g->force_dt > 0 ? g->force_dt : g->time_step;
  const int64_t mskip = vr_force_skip(g);
  if ((g->time_step < 24 && ((int64_t)fdt * mskip) % g->time_step > 0) || (g->time_step == 24 && g->time_step % fdt > 0)) {
    g_err = "Currently unable to handle a model starting date that does not correspond to a line in the forcing file.";
    return VR_ERR_ARG;
  }
  const int64_t skip = (int64_t)((float)(g->time_step * mskip)) / fdt;
  int64_t rec = 0;
  if (g->force_ascii) {
    FILE *f = fopen(path, "r");
    if (!f) { g_err = std::string("cannot open ") + path; return VR_ERR_ARG; }
    std::string line;
    for (int64_t i = 0; i < skip; i++)
      if (!read_line(f, line)) { fclose(f); g_err = "no data for the specified period in the forcing file"; return VR_ERR_ARG; }
    while (rec < nrec) {
      bool ok = true;
      for (int i = 0; i < nf && ok; i++) ok = fscanf(f, "%lf", &data[(size_t)i * nrec + rec]) == 1;
      if (!ok) break;
      rec++;
    }
    fclose(f);
  }
  else {
    FILE *f = fopen(path, "rb");
    if (!f) { g_err = std::string("cannot open ") + path; return VR_ERR_ARG; }
    auto rd = [&](unsigned short &v) {
      if (fread(&v, 2, 1, f) != 1) return false;
      if (!g->force_little_endian) v = (unsigned short)(((v & 0xFF) << 8) | ((v >> 8) & 0xFF));   // host is little endian
      return true;
    };
    unsigned short id[4] = {0, 0, 0, 0}, nb = 0;
    long nbytes = 0;
    for (int i = 0; i < 4; i++) rd(id[i]);
    if (id[0] == 0xFFFF && id[1] == 0xFFFF && id[2] == 0xFFFF && id[3] == 0xFFFF && rd(nb)) nbytes = nb;
    fseek(f, nbytes + (long)(skip * nf * 2), SEEK_SET);
    while (rec < nrec) {
      bool ok = true;
      for (int i = 0; i < nf && ok; i++) {
        unsigned short u;
        ok = rd(u);
        if (ok) data[(size_t)i * nrec + rec] = (g->force_signed[i] ? (double)(short)u : (double)u) / g->force_multiplier[i];
      }
      if (!ok) break;
      rec++;
    }
    fclose(f);
  }
  return rec;
}

// What a value becomes on its way through a "%.<d>f" text file and back (printf rounds the exact binary value
// to d decimals, ties to even; strtod returns the double nearest to that decimal): the exact product x*10^d
// is rebuilt as t + r with an fma, so borderline cases round like printf does, without formatting a string.
int vr_round_decimals(int64_t n, const double *in, double *out, int32_t decimals)
{
  if (!in || !out || n < 0 || decimals < 0 || decimals > 9) { g_err = "bad argument"; return VR_ERR_ARG; }
  double p = 1.0;
  for (int i = 0; i < decimals; i++) p *= 10.0;
  for (int64_t i = 0; i < n; i++) {
    const double x = in[i];
    if (!(fabs(x) < 1e11)) { out[i] = x; continue; }          // beyond 2^53 / 10^d the product is no longer exact enough
    const double t = x * p, r = fma(x, p, -t);                 // x*p == t + r exactly
    double k = nearbyint(t);                                  // ties to even on the ROUNDED product
    const double d0 = t - k;                                   // exact, a multiple of ulp(t) in [-0.5, 0.5]
    if (d0 == 0.5) { if (r > 0.0) k += 1.0; }                  // exact product above the tie -> up
    else if (d0 == -0.5) { if (r < 0.0) k -= 1.0; }            // below the tie -> down
    // |d0| < 0.5: |r| <= ulp(t)/2 cannot carry the exact product across a tie
    out[i] = k / p;
  }
  return VR_OK;
}

}  // extern "C"
