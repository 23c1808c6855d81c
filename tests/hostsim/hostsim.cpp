// hostsim.cpp -- TEST HARNESS ONLY.  Compiles the device physics headers (vic_physics.cuh,
// vic_aggregate.cuh, vic_host_prep.h) for the HOST with g++ and runs the same per-unit step and
// per-cell aggregation serially, so that the non-GPU test tier can check the kernel's arithmetic
// against the oracle.  It is not linked into libvicres.so and is not a product code path: the
// product library only launches CUDA kernels and fails without a device.
#include <stdio.h>
#include <vector>
#include "../../vicres_b200/csrc/vic_host_prep.h"
#include "../../vicres_b200/csrc/vic_aggregate.cuh"

using namespace vr;

// the unit's snow words for unit_record(): here simply the UnitS of the harness
struct HostSnowIO {
  UnitS *s;
  SnowS load() const
  {
    SnowS x;
    x.swq = s->swq; x.surf_temp = s->surf_temp; x.pack_temp = s->pack_temp; x.surf_water = s->surf_water;
    x.pack_water = s->pack_water; x.density = s->density; x.depth = s->depth; x.albedo = s->albedo;
    x.coldcontent = s->coldcontent; x.coverage = s->coverage; x.snow_canopy = s->snow_canopy;
    x.tmp_int_storage = s->tmp_int_storage; x.last_snow = s->last_snow; x.MELTING = s->MELTING;
    return x;
  }
  void store(const SnowS &x) const
  {
    s->swq = x.swq; s->surf_temp = x.surf_temp; s->pack_temp = x.pack_temp; s->surf_water = x.surf_water;
    s->pack_water = x.pack_water; s->density = x.density; s->depth = x.depth; s->albedo = x.albedo;
    s->coldcontent = x.coldcontent; s->coverage = x.coverage; s->snow_canopy = x.snow_canopy;
    s->tmp_int_storage = x.tmp_int_storage; s->last_snow = x.last_snow; s->MELTING = x.MELTING;
  }
  void store_none() const { s->albedo = NEW_SNOW_ALB; s->last_snow = (int)MISSINGV; s->MELTING = 0; }
  void fallback(int kind) const { if (kind) s->fb_fol++; else s->fb_snow++; }
};

extern "C" int hs_run(const vr_options *opt, const vr_veglib *lib, const vr_cells *cells, const vr_forcing *f,
                      const double *tair0, double *out /* [n_out][nrec][C] */)
{
  Prepared P;
  if (!prepare(*opt, *lib, *cells, 128, P)) { fprintf(stderr, "hostsim: %s\n", P.error.c_str()); return -1; }
  const int NL = P.NL, nrec = f->nrec;
  const int64_t C = P.C, U = P.U;
  std::vector<UnitS> S_(U ? U : 1);
  std::vector<UnitC> UC(U ? U : 1);
  std::vector<double> sv((size_t)SV_N * C, 0.0);
  for (int64_t u = 0; u < U; u++) {
    int64_t c = P.u_cell[u];
    UnitC &uc = UC[u];
    uc.is_bare = P.u_flags[u] & 1; uc.overstory = (P.u_flags[u] >> 1) & 1; uc.root0_mask = P.u_flags[u] >> 8;
    uc.area = P.u_cv[u] * P.u_af[u]; uc.Tfactor = P.u_Tfactor[u]; uc.Pfactor = P.u_Pfactor[u];
    uc.rarc = P.u_rarc[u]; uc.rmin = P.u_rmin[u]; uc.RGL = P.u_RGL[u];
    for (int l = 0; l < MAXL; l++) uc.root[l] = P.u_root[(size_t)l * U + u];
    double im[MAXL], mm[MAXL];
    for (int l = 0; l < NL; l++) { im[l] = cells->init_moist[(size_t)l * C + c]; mm[l] = cells->max_moist[(size_t)l * C + c]; }
    unit_init(im, mm, NL, tair0[c], uc.Tfactor, S_[u]);
  }
  TermMap tmap;
  build_term_map(opt->out_vars, opt->n_out, tmap);
  const int MAXU = 512;
  std::vector<double> buf((size_t)NTERM * MAXU);
  for (int64_t cs = 0; cs < C; cs++) {
    const int64_t c = P.order[cs];
    const int64_t fu0 = P.cell_uptr[cs];
    const int nu = (int)(P.cell_uptr[cs + 1] - fu0);
    CellC cc;
    load_cell_consts(P.cc.data(), C, cs, NL, cc);
    auto col_of = [&](int j) { return j; };
    auto S = [&](int k) { return buf[(size_t)tmap.row[k] * MAXU + 0]; };
    double *svc = &sv[(size_t)c * SV_N];
    double depth_m[MAXL] = {0, 0, 0};                     // options.MOISTFRACT: the cell's layer depths
    for (int l = 0; l < NL; l++) depth_m[l] = cc.depth(l);
    const double *mfd = opt->moistfract ? depth_m : nullptr;
    {   // put_data(rec = -nrecs)
      UnitOut z;
      memset(&z, 0, sizeof z);
      for (int j = 0; j < nu; j++) {
        const int64_t u = P.u_slot[fu0 + j];       // per-unit arrays are in slot order
        unit_terms(UC[u], P.u_cv[u], P.u_af[u], S_[u], z, NL, tmap, &buf[j], MAXU, nullptr, mfd);
      }
      if (nu > 0) {
        cell_accumulate_inplace(buf.data(), MAXU, tmap, nu, col_of, NL);
        cell_carry(S, NL, svc, mfd);
      }
    }
    const int NF = opt->dt_hours / opt->snow_step_hours, NS = NF > 1 ? NF + 1 : 1;
    for (int rec = 0; rec < nrec; rec++) {
      // sub-record kk of this record ([field][rec * NS + kk][cell]; kk == NS - 1 is the record itself)
      auto forc = [&](int kk) {
        Forc fo;
        size_t k = ((size_t)rec * NS + kk) * C + c;
        fo.air_temp = f->field[VR_F_AIR_TEMP][k]; fo.prec = f->field[VR_F_PREC][k]; fo.wind = f->field[VR_F_WIND][k];
        fo.vp = f->field[VR_F_VP][k]; fo.vpd = f->field[VR_F_VPD][k]; fo.pressure = f->field[VR_F_PRESSURE][k];
        fo.density = f->field[VR_F_DENSITY][k]; fo.shortwave = f->field[VR_F_SHORTWAVE][k];
        fo.longwave = f->field[VR_F_LONGWAVE][k];
        fo.mon = f->month[rec] - 1; fo.doy = f->day_in_year[rec];
        fo.snowflag = (NF > 1 && f->snowflag) ? f->snowflag[(size_t)rec * C + c] : 0;
        return fo;
      };
      const Forc fo = forc(NS - 1);
      for (int j = 0; j < nu; j++) {
        const int64_t u = P.u_slot[fu0 + j];
        const double *tm = &P.tm[((size_t)P.u_tile[u] * 12 + fo.mon) * TM_N];
        const double *ae = &P.ae[((size_t)P.u_aero[u] * 12 + fo.mon) * AE_N];
        const double *tl = P.tl.size() > 1 ? &P.tl[((size_t)P.u_tile[u] * 12 + fo.mon) * TL_N] : nullptr;
        const double *ap = P.ap.size() > 1 ? &P.ap[((size_t)P.u_aero[u] * 12 + fo.mon) * AP_N] : nullptr;
        HostSnowIO sio{&S_[u]};
        UnitS &s = S_[u];
        HotS h;
        for (int l = 0; l < MAXL; l++) h.moist.v[l] = s.moist[l];
        h.Wdew = s.Wdew; h.T0 = s.T0; h.T1 = s.T1; h.LongUnderOut = s.LongUnderOut; h.Tfoliage = s.Tfoliage;
        bool snowy = s.swq > 0 || s.snow_canopy > 0 || s.coverage != 0;
        bool tidy = !snowy && (s.last_snow != (int)MISSINGV || s.MELTING != 0 || s.albedo != NEW_SNOW_ALB);
#define HS_ARGS cc, UC[u], tm, ae, forc, NF, NL, opt->dt_hours, opt->snow_step_hours, opt->max_snow_temp, opt->min_rain_temp, h, snowy, \
                tidy, sio, P.u_cv[u], P.u_af[u], tmap, &buf[j], MAXU, true, tl, ap, opt->moistfract != 0
        // the instance the kernel would launch: the one without sums / averages when NF == 1
#define HS_STEP(EX) do { if (NF > 1) unit_record<EX, true>(HS_ARGS); else unit_record<EX, false>(HS_ARGS); } while (0)
        switch (term_extras(tmap)) {
          case 0: HS_STEP(0); break;
          case EX_PET: HS_STEP(EX_PET); break;
          case EX_ZWT: HS_STEP(EX_ZWT); break;
          default: HS_STEP(EX_PET | EX_ZWT); break;
        }
#undef HS_STEP
#undef HS_ARGS
        if (tidy) sio.store_none();     // the harness keeps no flags between records
        for (int l = 0; l < MAXL; l++) s.moist[l] = h.moist.v[l];
        s.Wdew = h.Wdew; s.T0 = h.T0; s.T1 = h.T1; s.LongUnderOut = h.LongUnderOut; s.Tfoliage = h.Tfoliage;
      }
      if (nu > 0) {
        cell_accumulate_inplace(buf.data(), MAXU, tmap, nu, col_of, NL);
        for (int v = 0; v < opt->n_out; v++)
          out[((size_t)v * nrec + rec) * C + c] = cell_output(opt->out_vars[v], S, fo, NL, false, svc, mfd);
        cell_carry(S, NL, svc, mfd);
      }
    }
  }
  return 0;
}
