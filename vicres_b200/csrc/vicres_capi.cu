// vicres_capi.cu -- C ABI (include/vicres.h) and CUDA kernels of the B200-native VIC-Res cell step.
//
// Mapping (DESIGN.md sections 3 and 4).  One THREAD owns one active tile x band unit; the units are in slot order
// (sorted by kind and climate) and everything a thread touches -- constants, forcing, state, hand-off words -- is a
// struct-of-arrays column in slot order, so every request of a warp is 32 consecutive doubles.  A record of all units
// runs as three small kernels (the cut is the call of runoff(), surface_fluxes.c:1130):
//   k_front       surface_fluxes' pass WITHOUT the snow model for every unit that carries no snow and gets no snowfall
//                 in this record; the others are appended to a list (warp ballot + one atomic per warp)
//   k_snow  the pass WITH the snow model (solve_snow, Brent solves, sub-steps) for the listed units only: the
//                 snow lanes are compacted into full warps instead of idling beside snow-free lanes
//   k_back        runoff() -- the hourly drainage loop, 45 % of the step's instructions -- and the terms after it
// and a chunk of records ends with
//   k_cells       one thread per (cell, record): adds the terms of the cell's units in reference order (put_data.c),
//                 derives the requested outputs, carries the save_data words, writes out[var][rec][cell].
// Each kernel keeps only its own live values in registers (no spill stack shared across phases), is small enough for the
// instruction cache without block-wide lock step, and runs at the occupancy its own register count allows.  HBM
// bandwidth is what this path has to spare (the step is FP64 / latency bound), so the hand-off between the kernels
// (10 words per unit) and the term round trip (about 0.9 kB per cell-timestep) are cheap.
// There is NO CPU fallback: every entry point needs a CUDA device and reports VR_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/vicres.h"
#include "../../include/vicres_files.h"
#include "vic_physics.cuh"
#include "vic_aggregate.cuh"
#include "vic_host_prep.h"

// threads per block / blocks per SM the register allocation of each step kernel must allow (tuning knobs,
// profiles/r2_tuning.md)
#ifndef VR_FRONT_CTA
#define VR_FRONT_CTA 128
#endif
#ifndef VR_FRONT_MINB
#define VR_FRONT_MINB 3
#endif
#ifndef VR_SNOW_CTA
#define VR_SNOW_CTA 512
#endif
#ifndef VR_SNOW_MINB
#define VR_SNOW_MINB 1
#endif
#ifndef VR_BACK_CTA
#define VR_BACK_CTA 192
#endif
#ifndef VR_BACK_MINB
#define VR_BACK_MINB 6
#endif
#ifndef VR_FRONT_STAGE
#define VR_FRONT_STAGE 1     // k_front: constants staged in shared memory by bulk copies (0: read from global memory)
#endif
#ifndef VR_SNOW_LOCKSTEP
#define VR_SNOW_LOCKSTEP 1   // k_snow: phase barriers across the block (one 512-thread block per SM): +10 %, profiles/r2_tuning.md
#endif
#ifndef VR_A_STREAMS
#define VR_A_STREAMS 1       // slot ranges of the non-snow class on streams of their own (VICRES_A_STREAMS overrides)
#endif
#define VR_CTA 256     // the set-up kernels

using namespace vr;

// forcing is read once per unit and record (and once more by k_cells); -DVR_STREAM_HINTS marks it evict-first like the
// term stream (measured: 2.6 % slower on the 100k-cell benchmark, profiles/r2_tuning.md, so off by default)
#if defined(VR_STREAM_HINTS)
#define VR_FORC_LD(p) __ldcs(p)
#else
#define VR_FORC_LD(p) __ldg(p)
#endif

namespace {

struct DevModel {
  int NL, NB, dt_hours, snow_step_hours, NF, n_out;   // NF = dt_hours / snow_step_hours passes of the snow model
  double max_snow_temp, min_rain_temp;
  int64_t C, U, FC;                 // FC = number of forcing columns (== C unless a forcing map is set)
  const double *cc;                 // [ccn][C] per-cell constants by sorted position
  const double *ccu;                // [ccn][US] the same per unit in slot order: what the step kernels read (coalesced)
  const double *tmu;                // [12][TM_N][US] the month's canopy terms per unit in slot order
  int64_t US;                       // row stride of ccu / tmu: U rounded up to whole thread blocks (bulk copies read full rows)
  int ccn;
  int moistfract;                   // options.MOISTFRACT
  const int64_t *cell_uptr, *fcol;
  const int64_t *order;             // sorted position -> caller's cell index
  const int32_t *u_cell, *u_pos, *u_tile, *u_flags, *u_aero, *u_slot, *u_logical, *cell_fail;
  const double *u_cv, *u_af, *u_Tfactor, *u_Pfactor, *u_rarc, *u_rmin, *u_RGL, *tm, *ae, *init_moist;
  const double *tl, *ap;            // potential-evaporation tables (valid when an OUT_PET_* is requested)
  const float *u_root;
  double *state;                    // [ST_N][U]
  double *fo;                       // [FO_N][U] what k_front / k_snow hand to k_back (FrontOut)
  int32_t *uflag;                   // [U] bit 0: the unit carries snow, bit 1: its stored snow flags need tidying
  int32_t *snow_list;               // [U] slots of the chunk's snow-class units (k_snow runs them)
  int32_t *snow_cnt;                // [1] their number
  unsigned long long *snow_total;   // diagnostic: unit-records that went through k_snow
  double *sv;                       // [SV_N][C]
  int32_t *cell_status;             // [C]
  int32_t out_vars[VR_MAX_OUT];
  TermMap tmap;                     // rows of the term buffer the requested outputs need
};

struct DevForcing {
  int nrec;                         // records in this launch
  int rec0, nrec_total;             // position of the launch inside the caller's block (output indexing)
  const int32_t *month, *doy;       // device
  const double *field[VR_NFORCING]; // device, [nrec][FC]
  const double *slots;              // device, [nrec * NS][VR_NFORCING][U]: the same values per unit in slot order (k_forcing_slots)
  const int32_t *snowflag;          // device, [nrec][FC] or null: atmos->snowflag[NR] (read only when NF > 1)
  int rec_abs;                      // records the handle has stepped before this launch (cell_status = first failing record + 1)
};

__device__ __forceinline__ void load_unit_consts(const DevModel &M, int64_t u, UnitC &uc, double &cv, double &af)
{
  const int fl = M.u_flags[u];
  uc.is_bare = fl & 1; uc.overstory = (fl >> 1) & 1; uc.root0_mask = fl >> 8;
  cv = M.u_cv[u]; af = M.u_af[u];
  uc.area = cv * af; uc.Tfactor = M.u_Tfactor[u]; uc.Pfactor = M.u_Pfactor[u];
  uc.rarc = M.u_rarc[u]; uc.rmin = M.u_rmin[u]; uc.RGL = M.u_RGL[u];
#pragma unroll
  for (int l = 0; l < MAXL; l++) uc.root[l] = M.u_root[(size_t)l * M.U + u];
}

__device__ __forceinline__ void load_state(const DevModel &M, int64_t u, UnitS &s)
{
#define L_(k) M.state[(size_t)(k) * M.U + u]
  s.moist[0] = L_(ST_MOIST0); s.moist[1] = L_(ST_MOIST1); s.moist[2] = L_(ST_MOIST2); s.Wdew = L_(ST_WDEW);
  s.T0 = L_(ST_T0); s.T1 = L_(ST_T1); s.LongUnderOut = L_(ST_LONGUNDEROUT); s.Tfoliage = L_(ST_TFOLIAGE);
  s.swq = L_(ST_SWQ); s.surf_temp = L_(ST_SURF_TEMP); s.pack_temp = L_(ST_PACK_TEMP);
  s.surf_water = L_(ST_SURF_WATER); s.pack_water = L_(ST_PACK_WATER); s.density = L_(ST_DENSITY);
  s.depth = L_(ST_DEPTH); s.albedo = L_(ST_ALBEDO); s.coldcontent = L_(ST_COLDCONTENT);
  s.coverage = L_(ST_COVERAGE); s.snow_canopy = L_(ST_SNOW_CANOPY); s.tmp_int_storage = L_(ST_TMP_INT);
  s.last_snow = (int)L_(ST_LAST_SNOW); s.MELTING = (int)L_(ST_MELTING); s.fb_snow = (int)L_(ST_FB_SNOW);
  s.fb_fol = (int)L_(ST_FB_FOL);
#undef L_
}

__device__ __forceinline__ void store_state(const DevModel &M, int64_t u, const UnitS &s)
{
#define S_(k) M.state[(size_t)(k) * M.U + u]
  S_(ST_MOIST0) = s.moist[0]; S_(ST_MOIST1) = s.moist[1]; S_(ST_MOIST2) = s.moist[2]; S_(ST_WDEW) = s.Wdew;
  S_(ST_T0) = s.T0; S_(ST_T1) = s.T1; S_(ST_LONGUNDEROUT) = s.LongUnderOut; S_(ST_TFOLIAGE) = s.Tfoliage;
  S_(ST_SWQ) = s.swq; S_(ST_SURF_TEMP) = s.surf_temp; S_(ST_PACK_TEMP) = s.pack_temp;
  S_(ST_SURF_WATER) = s.surf_water; S_(ST_PACK_WATER) = s.pack_water; S_(ST_DENSITY) = s.density;
  S_(ST_DEPTH) = s.depth; S_(ST_ALBEDO) = s.albedo; S_(ST_COLDCONTENT) = s.coldcontent;
  S_(ST_COVERAGE) = s.coverage; S_(ST_SNOW_CANOPY) = s.snow_canopy; S_(ST_TMP_INT) = s.tmp_int_storage;
  S_(ST_LAST_SNOW) = (double)s.last_snow; S_(ST_MELTING) = (double)s.MELTING; S_(ST_FB_SNOW) = (double)s.fb_snow;
  S_(ST_FB_FOL) = (double)s.fb_fol;
#undef S_
}

// per-unit copies, in slot order, of the per-cell constants and of the per (tile, month) canopy terms: a warp reads 32
// consecutive doubles per constant instead of gathering 32 sectors (the units of a cell sit in different slots: slots are
// sorted by kind and climate), and a thread block's slice of a row is one contiguous bulk copy
__global__ void k_unit_consts(DevModel M, double *ccu, double *tmu)
{
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= M.U) return;
  const int64_t pos = M.u_pos[u];
  for (int k = 0; k < M.ccn; k++) ccu[(size_t)k * M.US + u] = M.cc[(size_t)k * M.C + pos];
  const double *tm = M.tm + (size_t)M.u_tile[u] * 12 * TM_N;
  for (int k = 0; k < 12 * TM_N; k++) tmu[(size_t)k * M.US + u] = tm[k];
}

// options.MOISTFRACT: the layer depths (m) of the cell at sorted position pos, or null when the option is off
__device__ __forceinline__ const double *moistfract_depths(const DevModel &M, int64_t pos, double *d)
{
  if (!M.moistfract) return nullptr;
  for (int l = 0; l < MAXL; l++) d[l] = l < M.NL ? M.cc[(size_t)(CC_FL0 + l * CLF_N + CLF_DEPTH) * M.C + pos] : 0.;
  return d;
}

// initialize_model_state() for every unit + its terms for put_data(rec = -nrecs)
__global__ void __launch_bounds__(VR_CTA) k_init_units(DevModel M, const double *tair0 /* [FC] */, double *terms)
{
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= M.U) return;
  const int64_t c = M.u_cell[u];
  UnitC uc; UnitS s; UnitOut z;
  double cv, af, im[MAXL], mm[MAXL];
  load_unit_consts(M, u, uc, cv, af);
  for (int l = 0; l < M.NL; l++) {
    im[l] = M.init_moist[(size_t)l * M.C + c];
    mm[l] = M.cc[(size_t)(CC_BL0 + l * CLB_N + CLB_MAX_MOIST) * M.C + M.u_pos[u]];
  }
  const int64_t fc = M.fcol ? M.fcol[c] : c;
  unit_init(im, mm, M.NL, tair0[fc], uc.Tfactor, s);
  store_state(M, u, s);
  memset(&z, 0, sizeof z);
  double dm[MAXL];
  unit_terms(uc, cv, af, s, z, M.NL, M.tmap, terms + M.u_logical[u], (int)M.U, nullptr, moistfract_depths(M, M.u_pos[u], dm));
}

// the terms of a state that is already in place (restart from a state file) for put_data(rec = -nrecs)
__global__ void __launch_bounds__(VR_CTA) k_restored_units(DevModel M, double *terms)
{
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= M.U) return;
  UnitC uc; UnitS s; UnitOut z;
  double cv, af;
  load_unit_consts(M, u, uc, cv, af);
  load_state(M, u, s);
  memset(&z, 0, sizeof z);
  double dm[MAXL];
  unit_terms(uc, cv, af, s, z, M.NL, M.tmap, terms + M.u_logical[u], (int)M.U, nullptr, moistfract_depths(M, M.u_pos[u], dm));
}

__global__ void __launch_bounds__(128) k_init_cells(DevModel M, const double *terms, double *sums, int svb)
{
  const int64_t cs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // sorted position
  if (cs >= M.C) return;
  const int64_t c = M.order[cs], fu0 = M.cell_uptr[cs];
  const int nu = (int)(M.cell_uptr[cs + 1] - fu0);
  double sv[SV_N] = {0, 0, 0, 0};
  if (nu > 0) {
    cell_accumulate_to(terms, (size_t)M.U, fu0, nu, M.tmap, M.NL, sums, (size_t)M.C, cs);
    auto S = [&](int k) { return sums[(size_t)M.tmap.row[k] * M.C + cs]; };
    double dm[MAXL];
    cell_carry(S, M.NL, sv, moistfract_depths(M, cs, dm));
  }
  for (int k = 0; k < SV_N; k++) M.sv[((size_t)svb * SV_N + k) * M.C + c] = sv[k];
  M.cell_status[c] = M.cell_fail[c] ? 1 : 0;
}

// Forcing of a chunk per unit in slot order: k_units then reads 32 consecutive doubles per field and warp (the units of
// a cell sit in different slots, so reading [field][rec][cell] directly is a gather of 32 sectors per request).
__global__ void __launch_bounds__(256) k_forcing_slots(DevModel M, DevForcing F, double *__restrict__ dst)
{
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int rec = blockIdx.y;             // sub-record: F.nrec * NS of them
  if (u >= M.U) return;
  const int64_t c = M.u_cell[u];
  const size_t q = (size_t)rec * M.FC + (M.fcol ? M.fcol[c] : c);
  double *d = dst + (size_t)rec * VR_NFORCING * M.U + u;
#pragma unroll
  for (int i = 0; i < VR_NFORCING; i++) d[(size_t)i * M.U] = VR_FORC_LD(F.field[i] + q);
}

// the snow words of unit u in the state block (slot order): read and written only on the snow path of a record
struct SnowIO {
  double *st;          // &state[u]
  size_t U;
  bool wr;             // false for the idle threads of the last block and for failed cells: they read, never write
  __device__ __forceinline__ double &w(int k) const { return st[(size_t)k * U]; }
  __device__ __forceinline__ SnowS load() const
  {
    SnowS s;
    s.swq = w(ST_SWQ); s.surf_temp = w(ST_SURF_TEMP); s.pack_temp = w(ST_PACK_TEMP); s.surf_water = w(ST_SURF_WATER);
    s.pack_water = w(ST_PACK_WATER); s.density = w(ST_DENSITY); s.depth = w(ST_DEPTH); s.albedo = w(ST_ALBEDO);
    s.coldcontent = w(ST_COLDCONTENT); s.coverage = w(ST_COVERAGE); s.snow_canopy = w(ST_SNOW_CANOPY);
    s.tmp_int_storage = w(ST_TMP_INT); s.last_snow = (int)w(ST_LAST_SNOW); s.MELTING = (int)w(ST_MELTING);
    return s;
  }
  __device__ __forceinline__ void store(const SnowS &s) const
  {
    if (!wr) return;
    w(ST_SWQ) = s.swq; w(ST_SURF_TEMP) = s.surf_temp; w(ST_PACK_TEMP) = s.pack_temp; w(ST_SURF_WATER) = s.surf_water;
    w(ST_PACK_WATER) = s.pack_water; w(ST_DENSITY) = s.density; w(ST_DEPTH) = s.depth; w(ST_ALBEDO) = s.albedo;
    w(ST_COLDCONTENT) = s.coldcontent; w(ST_COVERAGE) = s.coverage; w(ST_SNOW_CANOPY) = s.snow_canopy;
    w(ST_TMP_INT) = s.tmp_int_storage; w(ST_LAST_SNOW) = (double)s.last_snow; w(ST_MELTING) = (double)s.MELTING;
  }
  __device__ __forceinline__ void store_none() const
  {
    if (!wr) return;
    w(ST_ALBEDO) = NEW_SNOW_ALB; w(ST_LAST_SNOW) = (double)(int)MISSINGV; w(ST_MELTING) = 0.;
  }
  __device__ __forceinline__ void fallback(int kind) const
  {
    if (wr) w(kind ? ST_FB_FOL : ST_FB_SNOW) += 1.;
  }
};

// bits of M.uflag
enum { UF_SNOWY = 1, UF_TIDY = 2, UF_SCLASS = 4, UF_FAIL = 8 };     // SCLASS: snow class in the current chunk (k_classify);
                                                                   // FAIL: the unit's cell failed at set-up (never stepped)

// uflag from the state block (after vr_init_state / vr_set_state / a state file)
__global__ void k_unit_flags(DevModel M)
{
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= M.U) return;
  const double *st = M.state + u;
  const size_t U = (size_t)M.U;
  const bool snowy = st[ST_SWQ * U] > 0 || st[ST_SNOW_CANOPY * U] > 0 || st[ST_COVERAGE * U] != 0;
  const bool tidy = !snowy && ((int)st[ST_LAST_SNOW * U] != (int)MISSINGV || st[ST_MELTING * U] != 0. || st[ST_ALBEDO * U] != NEW_SNOW_ALB);
  M.uflag[u] = (snowy ? UF_SNOWY : 0) | (tidy ? UF_TIDY : 0) | (M.cell_fail[M.u_cell[u]] ? UF_FAIL : 0);
}

// Classes of a chunk of records.  A unit belongs to the SNOW CLASS of the chunk when it carries snow at the start of the
// chunk or any record of the chunk brings it snowfall (or, with a sub-stepped snow model, snowfall anywhere in its cell):
// only then can solve_snow() have anything to do in the chunk.  The snow class is compacted into M.snow_list (one atomic
// per warp, lanes keep their order: neighbours in the list are neighbours in kind and climate) and runs through k_snow on
// its own stream; the others run k_front / k_back with the snow model compiled out.  The two classes never exchange
// anything inside a chunk, so their kernel sequences overlap freely: the snow class is few warps with a long dependent
// chain (two Brent solves per unit), the rest is many warps with short ones.
__global__ void __launch_bounds__(256) k_classify(DevModel M, DevForcing F)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t u = k < M.U ? k : M.U - 1;
  const int64_t c = M.u_cell[u];
  const int flags = M.uflag[u];
  const bool valid = k < M.U && !(flags & UF_FAIL);
  bool s = false;
  if (valid) {
    s = (flags & UF_SNOWY) != 0;
    const int NS = M.NF > 1 ? M.NF + 1 : 1;
    const double Tf = M.u_Tfactor[u], Pf = M.u_Pfactor[u];
    for (int rec = 0; rec < F.nrec && !s; rec++) {
      const double *q = F.slots + ((size_t)rec * NS + (NS - 1)) * VR_NFORCING * M.U + u;
      Forc fr;
      fr.air_temp = __ldg(q + (size_t)VR_F_AIR_TEMP * M.U); fr.prec = __ldg(q + (size_t)VR_F_PREC * M.U);
      fr.snowflag = (M.NF > 1 && F.snowflag) ? (int)__ldg(F.snowflag + (size_t)rec * M.FC + (M.fcol ? M.fcol[c] : c)) : 0;
      s = unit_needs_snow(fr, Tf, Pf, M.NF, M.max_snow_temp, M.min_rain_temp, false);
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, s);
  if (ballot) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(M.snow_cnt, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (s) M.snow_list[base + __popc(ballot & ((1u << lane) - 1u))] = (int32_t)u;
  }
  if (valid) {
    const int nf = (flags & ~UF_SCLASS) | (s ? UF_SCLASS : 0);
    if (nf != flags) M.uflag[u] = nf;
  }
}

// sub-record kk of record rec of the chunk for unit u, from the slot-ordered copy of k_forcing_slots
__device__ __forceinline__ Forc load_forcing(const DevModel &M, const DevForcing &F, int rec, int kk, int NS, int64_t u, int64_t c,
                                             int mon, int doy)
{
  Forc f;
  const double *q = F.slots + ((size_t)rec * NS + kk) * VR_NFORCING * M.U + u;
  const size_t st = (size_t)M.U;
  f.air_temp = __ldg(q + VR_F_AIR_TEMP * st); f.prec = __ldg(q + VR_F_PREC * st); f.wind = __ldg(q + VR_F_WIND * st);
  f.vp = __ldg(q + VR_F_VP * st); f.vpd = __ldg(q + VR_F_VPD * st); f.pressure = __ldg(q + VR_F_PRESSURE * st);
  f.density = __ldg(q + VR_F_DENSITY * st); f.shortwave = __ldg(q + VR_F_SHORTWAVE * st); f.longwave = __ldg(q + VR_F_LONGWAVE * st);
  f.mon = mon; f.doy = doy;
  f.snowflag = (M.NF > 1 && F.snowflag) ? (int)__ldg(F.snowflag + (size_t)rec * M.FC + (M.fcol ? M.fcol[c] : c)) : 0;
  return f;
}

// ---- staging of a thread block's constants in shared memory -------------------------------------------------------------
// stage[row][thread]: rows 0 .. CC_FRONT_N-1 = the front block of the constants, then TM_N rows = the month's canopy terms.
// k_front owns VR_FRONT_CTA consecutive slots, so each row is one contiguous slice of ccu / tmu: ONE thread issues the
// rows as bulk asynchronous copies (TMA, cp.async.bulk -> SASS UBLKCP) that complete on an mbarrier while all threads
// classify their units; the physics then reads its constants from shared memory (~30 cycles) instead of stalling on one
// L2 / HBM round trip per constant (k_front was 75 % long-scoreboard stalls, profiles/r2_k_units_ncu.md).  While a block
// computes, the copies of the next blocks of the SM are in flight (4 resident blocks = a 4-deep ring of stages).
constexpr int STAGE_ROWS = CC_FRONT_N + TM_N;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the two tables of the device power function into vr_pow_tables, on the same mbarrier as the block's stage (instead of
// pow_tables_init()'s loads and block barrier)
constexpr uint32_t POW_TABLE_BYTES = (uint32_t)sizeof(PowTables);
__device__ __forceinline__ void pow_tables_bulk(uint64_t *bar)
{
  bulk_g2s(vr_pow_tables.logtab, VR_POW_LOGTAB_G, (uint32_t)sizeof(vr_pow_tables.logtab), bar);
  bulk_g2s(vr_pow_tables.exptab, VR_POW_EXPTAB_G, (uint32_t)sizeof(vr_pow_tables.exptab), bar);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  } while (!ok);
}

// surface_fluxes' passes of record `rec` of the chunk for unit u (everything before runoff()): hot state in, hot state
// and the hand-off words out, terms to the term buffer column of the unit.  cc: the unit's staged constants,
// tmrow: its staged canopy terms (row stride CTA)
template <int EXTRAS, bool SUB, bool SNOW>
__device__ __forceinline__ FrontOut front_unit(const DevModel &M, const DevForcing &F, int rec, int64_t u, int64_t c, int flags, int mon,
                                               CellC cc, const double *tmrow, int64_t CTA, double *__restrict__ terms,
                                               bool wr = true, bool lockstep = false)
{
  UnitC uc;
  double cv, af;
  load_unit_consts(M, u, uc, cv, af);
  SnowIO sio;
  sio.st = M.state + u;
  sio.U = (size_t)M.U;
  sio.wr = wr;
  double *st = M.state + u;
  const size_t U = (size_t)M.U;
  HotS h;
  h.moist.v[0] = st[ST_MOIST0 * U]; h.moist.v[1] = st[ST_MOIST1 * U]; h.moist.v[2] = st[ST_MOIST2 * U];
  h.Wdew = st[ST_WDEW * U]; h.T0 = st[ST_T0 * U]; h.T1 = st[ST_T1 * U]; h.LongUnderOut = st[ST_LONGUNDEROUT * U];
  h.Tfoliage = st[ST_TFOLIAGE * U];
  bool snowy = (flags & UF_SNOWY) != 0, tidy = (flags & UF_TIDY) != 0;
  const int doy = __ldg(F.doy + rec);
  const int NS = M.NF > 1 ? M.NF + 1 : 1;
  CanopyM cm;
  cm.vegcover = tmrow[TM_VEGCOVER * CTA]; cm.LAI = tmrow[TM_LAI * CTA]; cm.Wdmax = tmrow[TM_WDMAX * CTA];
  cm.albedo = tmrow[TM_ALBEDO * CTA]; cm.surf_atten = tmrow[TM_SURF_ATTEN * CTA];
  const double *ae = M.ae + ((size_t)M.u_aero[u] * 12 + mon) * AE_N;
  const double *tl = (EXTRAS & EX_PET) ? M.tl + ((size_t)M.u_tile[u] * 12 + mon) * TL_N : nullptr;
  const double *ap = (EXTRAS & EX_PET) ? M.ap + ((size_t)M.u_aero[u] * 12 + mon) * AP_N : nullptr;
  auto forc = [&](int kk) { return load_forcing(M, F, rec, kk, NS, u, c, mon, doy); };
  const FrontOut fo = unit_front<EXTRAS, SUB, SNOW>(cc, uc, cm, ae, forc, M.NF, M.NL, M.dt_hours, M.snow_step_hours, M.max_snow_temp,
                                                    M.min_rain_temp, h, snowy, tidy, sio, cv, af, M.tmap,
                                                    terms + (size_t)rec * M.tmap.n * M.U + M.u_logical[u], (int)M.U, wr, lockstep,
                                                    tl, ap);
  if (!wr) return fo;      // a thread that only keeps its block's barriers company (k_snow in lock step)
  st[ST_WDEW * U] = h.Wdew; st[ST_T0 * U] = h.T0; st[ST_T1 * U] = h.T1; st[ST_LONGUNDEROUT * U] = h.LongUnderOut;
  st[ST_TFOLIAGE * U] = h.Tfoliage;
  {     // hand-off to k_back
    double *o = M.fo + u;
    o[FO_PPT * U] = fo.ppt; o[FO_VAPOR * U] = fo.vapor; o[FO_CVAPOR * U] = fo.cvapor; o[FO_CANOPYEVAP * U] = fo.canopyevap;
    o[FO_EVAP0 * U] = fo.evap.v[0]; o[FO_EVAP1 * U] = fo.evap.v[1]; o[FO_EVAP2 * U] = fo.evap.v[2];
    o[FO_BEF0 * U] = fo.bef.v[0]; o[FO_BEF1 * U] = fo.bef.v[1]; o[FO_BEF2 * U] = fo.bef.v[2];
  }
  const int nf = (flags & (UF_SCLASS | UF_FAIL)) | (snowy ? UF_SNOWY : 0) | (tidy ? UF_TIDY : 0);
  if (nf != flags) M.uflag[u] = nf;
  if (fo.err) atomicCAS(M.cell_status + c, 0, F.rec_abs + rec + 1);     // first failing record + 1 (arno_evap.c:112-151)
  return fo;
}

// Record `rec` of the chunk, everything before runoff() for the units outside the chunk's snow class: one thread per unit
// in slot order (the lanes of snow-class units idle).
template <int EXTRAS>
__global__ void __launch_bounds__(VR_FRONT_CTA, VR_FRONT_MINB) k_front(DevModel M, DevForcing F, int rec, double *__restrict__ terms,
                                                                       int64_t ub, int64_t ue)
{
  extern __shared__ __align__(128) double stage[];     // [STAGE_ROWS][VR_FRONT_CTA]
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int64_t u0 = ub + (int64_t)blockIdx.x * VR_FRONT_CTA;     // this launch covers slots [ub, ue)
  const int mon = __ldg(F.month + rec) - 1;
  if (VR_FRONT_STAGE && tid == 0) mbar_init(&bar, 1);
  if (VR_FRONT_STAGE) __syncthreads();                 // the mbarrier is initialised for everyone
  else pow_tables_init();
  if (VR_FRONT_STAGE && tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)(STAGE_ROWS * VR_FRONT_CTA * sizeof(double)) + POW_TABLE_BYTES);
    pow_tables_bulk(&bar);
    for (int k = 0; k < CC_FRONT_N; k++)
      bulk_g2s(stage + k * VR_FRONT_CTA, M.ccu + (size_t)k * M.US + u0, VR_FRONT_CTA * sizeof(double), &bar);
    for (int k = 0; k < TM_N; k++)
      bulk_g2s(stage + (CC_FRONT_N + k) * VR_FRONT_CTA, M.tmu + ((size_t)mon * TM_N + k) * M.US + u0, VR_FRONT_CTA * sizeof(double), &bar);
  }
  const int64_t k = u0 + tid;
  const int64_t u = k < M.U ? k : M.U - 1;
  const int64_t c = M.u_cell[u];
  const int flags = M.uflag[u];
  const bool mine = k < ue && !(flags & (UF_SCLASS | UF_FAIL));
  if (VR_FRONT_STAGE) mbar_wait(&bar, 0);     // every thread: the block's shared memory must not be released under a copy in flight
  if (!mine) return;
  CellC cc;
  if (VR_FRONT_STAGE) {
    cc.p = stage + tid; cc.C = VR_FRONT_CTA;
    front_unit<EXTRAS, false, false>(M, F, rec, u, c, flags, mon, cc, stage + CC_FRONT_N * VR_FRONT_CTA + tid, VR_FRONT_CTA, terms);
  }
  else {
    cc.p = M.ccu + u; cc.C = M.US;
    front_unit<EXTRAS, false, false>(M, F, rec, u, c, flags, mon, cc, M.tmu + (size_t)mon * TM_N * M.US + u, M.US, terms);
  }
}

// Record `rec` of the chunk: runoff() and what follows it for unit u (k_back: the units outside the snow class, one
// thread per unit in slot order; k_back_list: the snow class).  lk: the unit's staged loop constants.
template <int EXTRAS>
__device__ __forceinline__ void back_unit(const DevModel &M, int rec, int64_t u, LoopSmem lk, double *__restrict__ terms)
{
  const size_t U = (size_t)M.U;
  CellC cc;
  cc.p = M.ccu + u; cc.C = M.US;
  // the constants read after the hourly loop: on their way to L1 while the loop runs
  for (int l = 0; l < M.NL; l++) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(cc.p + (size_t)(CC_FL0 + l * CLF_N + CLF_WPWP) * cc.C));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(cc.p + (size_t)(CC_BL0 + l * CLB_N + CLB_WET_DEN) * cc.C));
  }
  const double *o = M.fo + u;
  FrontOut fo;
  fo.ppt = o[FO_PPT * U]; fo.vapor = o[FO_VAPOR * U]; fo.cvapor = o[FO_CVAPOR * U]; fo.canopyevap = o[FO_CANOPYEVAP * U];
  fo.evap.v[0] = o[FO_EVAP0 * U]; fo.evap.v[1] = o[FO_EVAP1 * U]; fo.evap.v[2] = o[FO_EVAP2 * U];
  fo.bef.v[0] = o[FO_BEF0 * U]; fo.bef.v[1] = o[FO_BEF1 * U]; fo.bef.v[2] = o[FO_BEF2 * U];
  fo.err = 0;
  double *st = M.state + u;
  M3 moist;
  moist.v[0] = st[ST_MOIST0 * U]; moist.v[1] = st[ST_MOIST1 * U]; moist.v[2] = st[ST_MOIST2 * U];
  const int fl = M.u_flags[u];
  const int NL = M.NL, dt = M.dt_hours;
  unit_back_fn<EXTRAS>(cc, (fl & 1) == 0, fl >> 8, NL, fo, moist, M.u_cv[u], M.u_af[u], M.tmap,
                       terms + (size_t)rec * M.tmap.n * M.U + M.u_logical[u], (int)M.U, true,
                       [&](double ppt, M3 m, M3 e) { return runoff_step_staged(cc, lk, NL, dt, ppt, m, e); }, M.moistfract != 0);
  st[ST_MOIST0 * U] = moist.v[0]; st[ST_MOIST1 * U] = moist.v[1]; st[ST_MOIST2 * U] = moist.v[2];
}

// k_back owns VR_BACK_CTA consecutive slots: the 18 constants of the hourly loop are 18 contiguous row slices of ccu that
// one thread fetches as bulk copies (TMA) into the stage [LK_N][VR_BACK_CTA]; the loop re-reads them from shared memory
// every hour instead of holding 36 registers, which is what lets the kernel keep all its warps resident at once (one
// wave instead of two half-empty ones, profiles/r2_tuning.md).
template <int EXTRAS>
__global__ void __launch_bounds__(VR_BACK_CTA, VR_BACK_MINB) k_back(DevModel M, int rec, double *__restrict__ terms, int64_t ub,
                                                                     int64_t ue)
{
  __shared__ __align__(128) double stage[LK_N * VR_BACK_CTA];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int64_t u0 = ub + (int64_t)blockIdx.x * VR_BACK_CTA;      // this launch covers slots [ub, ue)
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();                                     // the mbarrier is initialised for everyone
  if (tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)(LK_N * VR_BACK_CTA * sizeof(double)) + POW_TABLE_BYTES);
    pow_tables_bulk(&bar);
    for (int k = 0; k < LK_N; k++)
      bulk_g2s(stage + k * VR_BACK_CTA, M.ccu + (size_t)loop_const_row(k, M.NL) * M.US + u0, VR_BACK_CTA * sizeof(double), &bar);
  }
  const int64_t k = u0 + tid;
  const int64_t u = k < M.U ? k : M.U - 1;
  const bool mine = k < ue && !(M.uflag[u] & (UF_SCLASS | UF_FAIL));
  mbar_wait(&bar, 0);
  if (!mine) return;
  LoopSmem lk;
  lk.a = smem_u32(stage + tid); lk.S = VR_BACK_CTA * (int)sizeof(double); lk.qd0 = lk.qd1 = lk.qdB = lk.one_m_Ws = 0;
  back_unit<EXTRAS>(M, rec, u, lk, terms);
}

// Record `rec` of the chunk for the snow class: solve_snow() and the rest of the pass (sub-stepped when SNOW_STEP <
// TIME_STEP: SUB) for the listed units, then k_back_list = runoff() for the same units at its own (smaller) register
// count.  Grid-stride over the list (its length is only known on the device).  Both run on the snow stream beside
// k_front / k_back of the other units.  (One launch per record, not a loop over the chunk's records inside the kernel:
// warps that drift apart in this large code thrash the instruction cache -- measured 495 us against 299 us per record,
// profiles/r2_tuning.md.)
// [lo, hi) of the snow list that part `part` of `nparts` handles (parts run on streams of their own so that the blocks
// of one part's next record can start while another part's slow blocks still run); boundaries on block multiples
__device__ __forceinline__ void list_part(int n, int part, int nparts, int align, int64_t &lo, int64_t &hi)
{
  const int64_t per = (((int64_t)n + nparts - 1) / nparts + align - 1) / align * align;
  lo = (int64_t)part * per;
  hi = lo + per;
  if (lo > n) lo = n;
  if (hi > n || part == nparts - 1) hi = n;
}

template <int EXTRAS, bool SUB>
__global__ void __launch_bounds__(VR_SNOW_CTA, VR_SNOW_MINB) k_snow(DevModel M, DevForcing F, int rec, double *__restrict__ terms,
                                                                     int part, int nparts)
{
  pow_tables_init();
  const int n = M.snow_cnt[0];
  const int mon = __ldg(F.month + rec) - 1;
  if (part == 0 && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(M.snow_total, (unsigned long long)n);
  int64_t lo, hi;
  list_part(n, part, nparts, VR_SNOW_CTA, lo, hi);
#if VR_SNOW_LOCKSTEP
  // one block per SM whose warps pass the phases of the pass together (block barriers in surface_pass): every
  // instruction-cache line is fetched once for all of them.  Threads past the end of the list shadow its last unit
  // without writing anything, so that all threads of a block reach the barriers.
  for (int64_t base = lo + (int64_t)blockIdx.x * blockDim.x; base < hi; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const bool wr = i < hi;
    const int64_t u = M.snow_list[wr ? i : hi - 1];
    CellC cc;
    cc.p = M.ccu + u; cc.C = M.US;
    front_unit<EXTRAS, SUB, true>(M, F, rec, u, M.u_cell[u], M.uflag[u], mon, cc, M.tmu + (size_t)mon * TM_N * M.US + u, M.US, terms,
                                  wr, true);
  }
#else
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = M.snow_list[i];
    CellC cc;
    cc.p = M.ccu + u; cc.C = M.US;
    front_unit<EXTRAS, SUB, true>(M, F, rec, u, M.u_cell[u], M.uflag[u], mon, cc, M.tmu + (size_t)mon * TM_N * M.US + u, M.US, terms);
  }
#endif
}

template <int EXTRAS>
__global__ void __launch_bounds__(VR_BACK_CTA, VR_BACK_MINB) k_back_list(DevModel M, int rec, double *__restrict__ terms, int part,
                                                                          int nparts)
{
  __shared__ double stage[LK_N * VR_BACK_CTA];
  pow_tables_init();
  const int n = M.snow_cnt[0];
  int64_t lo, hi;
  list_part(n, part, nparts, VR_SNOW_CTA, lo, hi);      // the same cut as k_snow's: a part's units stay on its stream
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = M.snow_list[i];
    double *col = stage + threadIdx.x;
#pragma unroll
    for (int k = 0; k < LK_N; k++) col[k * VR_BACK_CTA] = __ldg(M.ccu + (size_t)loop_const_row(k, M.NL) * M.US + u);
    LoopSmem lk;
    lk.a = smem_u32(col); lk.S = VR_BACK_CTA * (int)sizeof(double); lk.qd0 = lk.qd1 = lk.qdB = lk.one_m_Ws = 0;
    back_unit<EXTRAS>(M, rec, u, lk, terms);
  }
}

// put_data() for the records of a chunk: one thread per (cell, record).  A cell's units are contiguous columns of the
// term buffer and consecutive threads own consecutive cells, so the term reads, the forcing reads and the output
// writes out[var][rec][cell] are all coalesced (the cells keep the caller's order; only the units are sorted into
// slots).  Every requested output adds the terms it needs IN REFERENCE ORDER straight from the term buffer
// (collect_wb_terms' summation order, so the fp64 sums round like the reference's); nothing is staged in memory.  The
// save_data words a record needs from its predecessor (storage deltas, closure error) are rebuilt from the predecessor's
// terms, which makes the records of a chunk independent; record 0 uses the carried M.sv.  M.sv is double-buffered: the
// blocks of record 0 read buffer svb while the blocks of the last record write buffer svb ^ 1 (blocks of one grid have
// no order), the host flips svb per launch.
#ifndef VR_SNOW_PARTS
#define VR_SNOW_PARTS 1
#endif
#ifndef VR_CELLS_MINB
#define VR_CELLS_MINB 4     // resident blocks of k_cells the register allocation has to allow: 64 registers, no spill.  The
#endif                      // kernel waits on its loads (26 % of HBM); 32 warps per SM instead of 8 have more of them in flight
__global__ void __launch_bounds__(256, VR_CELLS_MINB) k_cells(DevModel M, DevForcing F, const double *__restrict__ terms,
                                               double *__restrict__ out, int svb)
{
  const int64_t cs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int rec = blockIdx.y;
  if (cs >= M.C) return;
  const int64_t c = M.order[cs], fu0 = M.cell_uptr[cs];
  const int nu = (int)(M.cell_uptr[cs + 1] - fu0);
  const int64_t fc = M.fcol ? M.fcol[c] : c;
  const bool ok = !M.cell_fail[c] && nu > 0;
  const int NL = M.NL;
  const size_t rec_stride = (size_t)M.tmap.n * M.U;
  const size_t orow = (size_t)(F.rec0 + rec) * M.C + c, ostride = (size_t)F.nrec_total * M.C;
  if (!ok) {
    for (int v = 0; v < M.n_out; v++) out[(size_t)v * ostride + orow] = 0.;
    return;
  }
  const double *T = terms + rec * rec_stride;
  auto S = [&](int k) { return cell_sum(T, (size_t)M.U, fu0, nu, M.tmap, NL, k); };
  double sv[SV_N], dm[MAXL];
  const double *mfd = moistfract_depths(M, cs, dm);
  if (rec == 0)
    for (int k = 0; k < SV_N; k++) sv[k] = M.sv[((size_t)svb * SV_N + k) * M.C + c];
  else {
    const double *P = terms + (rec - 1) * rec_stride;
    auto SP = [&](int k) { return cell_term_sum(P, (size_t)M.U, fu0, nu, M.tmap.row[k]); };
    cell_carry(SP, NL, sv, mfd);
  }
  Forc f;
  const int NS = M.NF > 1 ? M.NF + 1 : 1;
  const size_t q = ((size_t)rec * NS + (NS - 1)) * M.FC + fc;        // the record's own values (atmos-><field>[NR])
  f.air_temp = __ldg(F.field[VR_F_AIR_TEMP] + q); f.wind = __ldg(F.field[VR_F_WIND] + q);
  f.vp = __ldg(F.field[VR_F_VP] + q); f.vpd = __ldg(F.field[VR_F_VPD] + q);
  f.shortwave = __ldg(F.field[VR_F_SHORTWAVE] + q); f.longwave = __ldg(F.field[VR_F_LONGWAVE] + q);
  f.pressure = __ldg(F.field[VR_F_PRESSURE] + q); f.density = __ldg(F.field[VR_F_DENSITY] + q);
  f.prec = 0; f.mon = 0; f.doy = 0; f.snowflag = 0;
#pragma unroll 1
  for (int v = 0; v < M.n_out; v++) VR_STREAM_ST(out + (size_t)v * ostride + orow, cell_output(M.out_vars[v], S, f, NL, false, sv, mfd));
  if (rec == F.nrec - 1) {
    cell_carry(S, NL, sv, mfd);
    for (int k = 0; k < SV_N; k++) M.sv[((size_t)(svb ^ 1) * SV_N + k) * M.C + c] = sv[k];
  }
}

// Temporal aggregation of put_data.c:665-688: thread per (cell, output interval, variable).  kind: 0 = last record of the
// interval (AGG_TYPE_END: storages), 1 = sum (AGG_TYPE_SUM: fluxes), 2 = sum of x / ratio (AGG_TYPE_AVG), 3 = the
// conductance averaged like 2 and inverted (OUT_AERO_RESIST, :686), all in record order like the reference's += .
__global__ void k_aggregate(int n, const int *__restrict__ kind, int nrec, int ratio, int64_t C, const double *__restrict__ in,
                            double *__restrict__ out)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.z, nout = nrec / ratio;
  if (c >= C) return;
  const int k = kind[v];
  for (int o = blockIdx.y; o < nout; o += gridDim.y) {       // gridDim.y is capped at 65 535
    const double *p = in + ((size_t)v * nrec + (size_t)o * ratio) * C + c;
    double a = 0.0;
    if (k == 0) a = p[(size_t)(ratio - 1) * C];
    else
      for (int r = 0; r < ratio; r++) a += (k == 1) ? p[(size_t)r * C] : p[(size_t)r * C] / ratio;
    if (k == 3) a = 1 / a;
    out[((size_t)v * nout + o) * C + c] = a;
  }
}

// ALMA_OUTPUT (put_data.c:697-760) on aggregated values, in place: op 1 = / seconds, 2 = + KELVIN, 3 = * seconds, 5 = kPa -> Pa; op 4
// (OUT_SUB_SNOW: / seconds, then + the converted OUT_SUB_CANOP, :741-742) is left to k_alma_add, launched after this
// kernel so that it reads the companion row converted.  Thread per (cell, record, variable).
__global__ void k_alma(int n, const int *__restrict__ op, const int *__restrict__ aux, int nrec, int64_t C, double sec, double *__restrict__ d)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.z;
  if (c >= C) return;
  const int k = op[v];
  for (int r = blockIdx.y; r < nrec; r += gridDim.y) {
    const size_t i = ((size_t)v * nrec + r) * C + c;
    if (k == 1) d[i] = d[i] / sec;
    else if (k == 2) d[i] = d[i] + 273.15;
    else if (k == 3) d[i] = d[i] * sec;
    else if (k == 5) d[i] = d[i] * 1000;
  }
}
__global__ void k_alma_add(int v, int aux, int nrec, int64_t C, double sec, double *__restrict__ d)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int r = blockIdx.y; r < nrec; r += gridDim.y) {
    const size_t i = ((size_t)v * nrec + r) * C + c, j = ((size_t)aux * nrec + r) * C + c;
    d[i] = d[i] / sec;
    d[i] = d[i] + d[j];
  }
}

// OUTPUT_FORCE (write_forcing_file.c:46-85): thread per (cell, row = rec * nf + window, variable) of the forcing files
struct ForceFields { const double *f[VR_NFORCING]; };
__global__ void k_forcing_outputs(int n, const int *__restrict__ vars, int nrows, int nf, int64_t C, double max_snow, double min_rain,
                                  ForceFields F, double *__restrict__ out)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.z;
  if (c >= C) return;
  for (int row = blockIdx.y; row < nrows; row += gridDim.y) {
  const int rec = row / nf, j = row - rec * nf, NS = nf > 1 ? nf + 1 : 1;
  const size_t q = ((size_t)rec * NS + j) * C + c;
  const double ta = F.f[VR_F_AIR_TEMP][q], prec = F.f[VR_F_PREC][q], vp = F.f[VR_F_VP][q], vpd = F.f[VR_F_VPD][q],
               pr = F.f[VR_F_PRESSURE][q];
  double rain;                                         // the split of :66-78
  if (ta >= max_snow) rain = prec;
  else if (ta <= min_rain) rain = 0;
  else rain = ((ta - min_rain) / (max_snow - min_rain)) * prec;
  double x = 0.;
  switch (vars[v]) {
    case VR_OUT_AIR_TEMP: x = ta; break;
    case VR_OUT_DENSITY: x = F.f[VR_F_DENSITY][q]; break;
    case VR_OUT_LONGWAVE: x = F.f[VR_F_LONGWAVE][q]; break;
    case VR_OUT_PREC: x = prec; break;
    case VR_OUT_PRESSURE: x = pr / 1000.; break;
    case VR_OUT_QAIR: x = 0.62196351 * vp / pr; break;
    case VR_OUT_REL_HUMID: x = 100. * vp / (vp + vpd); break;
    case VR_OUT_SHORTWAVE: x = F.f[VR_F_SHORTWAVE][q]; break;
    case VR_OUT_VP: x = vp / 1000.; break;
    case VR_OUT_VPD: x = vpd / 1000.; break;
    case VR_OUT_WIND: x = F.f[VR_F_WIND][q]; break;
    case VR_OUT_RAINF: x = rain; break;
    case VR_OUT_SNOWF: x = (ta >= max_snow) ? 0 : ((ta <= min_rain) ? prec : prec - rain); break;
  }
  out[((size_t)v * nrows + row) * C + c] = x;
  }
}

// diagnostic: the device power function on arbitrary arguments (tests/test_gpu_parity.py checks its error bound)
__global__ void k_diag_pow(int64_t n, const double *x, const double *y, double *z)
{
  pow_tables_init();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = m_pow(x[i], y[i]);
}

// diagnostic: FP64 FMA issue peak of the device -- 8 independent DFMA chains per thread, all SMs
__global__ void __launch_bounds__(256) k_diag_dfma(int iters, double seed, double *sink)
{
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 12345.678) sink[0] = r;
}

template <class T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t upload(const std::vector<T> &v)
  {
    cudaError_t e = reserve(v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
  cudaError_t reserve(size_t count)
  {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct vr_handle {
  vr_options opt;
  int device = 0;
  std::string err;
  bool have_lib = false, have_cells = false, initialised = false;
  // veglib copy
  std::vector<int32_t> lib_overstory;
  std::vector<double> lib_rarc, lib_rmin, lib_wind_h, lib_RGL, lib_rad_atten, lib_wind_atten, lib_trunk, lib_rough, lib_disp,
      lib_LAI, lib_albedo;
  Prepared P;
  // what the reference's state file needs of the cells (write_model_state.c / read_initial_model_state.c)
  struct CellKeep {
    std::vector<int64_t> tile_ptr;
    std::vector<int32_t> tile_class;
    std::vector<double> tile_Cv, AreaFract, Tfactor, init_moist, max_moist, depth0, dp, avg_temp;
    int nclass = 0;
  } keep;
  std::vector<double> tair0_host;
  DevModel M;
  DBuf<double> d_tl, d_ap, d_ccu, d_tmu, d_fslots;
  DBuf<double> d_cc, d_u_cv, d_u_af, d_u_Tf, d_u_Pf, d_u_rarc, d_u_rmin, d_u_RGL, d_tm, d_ae, d_init_moist, d_state, d_sv,
      d_forc, d_out, d_tair0, d_terms, d_sums;
  DBuf<float> d_u_root;
  DBuf<int64_t> d_cell_uptr, d_fcol, d_order;
  DBuf<int32_t> d_u_logical, d_u_cell, d_u_pos, d_u_tile, d_u_flags, d_u_aero, d_u_slot, d_cell_fail, d_cell_status, d_month, d_doy;
  cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
  cudaStream_t s_snow = nullptr;     // k_snow of the chunk's snow class runs here, beside k_front / k_back on the step stream
  cudaEvent_t ev_cls = nullptr, ev_snow = nullptr;
  cudaStream_t s_cells = nullptr;    // k_cells of chunk k runs here, beside k_units of chunk k+1 on the step stream
  cudaEvent_t ev_fork = nullptr, ev_units[2] = {nullptr, nullptr}, ev_cells[2] = {nullptr, nullptr};
  int64_t pair_seq = 0;              // launch pairs so far: pair i uses term buffer i & 1
  int svb = 0;                       // which half of d_sv holds the carried save_data words
  int64_t rec_done = 0;              // records stepped since vr_init_state / vr_set_state (cell_status numbering)
  DBuf<int32_t> d_snowflag;
  struct EvTriple { cudaEvent_t u0, u1, c0, c1; int nrec; };
  std::vector<EvTriple> evs;         // around every k_units / k_cells launch since the last vr_kernel_times()
  size_t evs_used = 0;
  static constexpr int NBUF = 16;    // at most this many copy buffers per direction of the host pipeline (vr_step_block picks)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_in[NBUF] = {}, ev_comp[NBUF] = {}, ev_out[NBUF] = {};
  int extras = 0;                    // EX_PET | EX_ZWT when such outputs are requested: selects the kernel instances
  int cta = VR_CTA;                  // threads per block of the set-up kernels
  int64_t ncta = 0;
  int sms = 148;
  bool flags_dirty = true;           // M.uflag must be rebuilt from the state block before the next step
  DBuf<int32_t> d_uflag, d_snow_list, d_snow_cnt;
  DBuf<unsigned long long> d_snow_total;
  int na = 1;                        // slot ranges of the non-snow class that run side by side (VICRES_A_STREAMS, 1..4)
  cudaStream_t s_a[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_a[3] = {nullptr, nullptr, nullptr};
  int snow_parts = 1;                // VICRES_SNOW_PARTS: the snow list in this many parts on streams of their own (<= 4)
  cudaStream_t s_sp[3] = {};
  cudaEvent_t ev_sp[3] = {};
  bool snow_stream = true;           // k_snow on its own stream (VICRES_SNOW_STREAM=0: on the step stream, for A/B timing)
  bool detail = false;               // VICRES_TIME_KERNELS=1: CUDA events around every step kernel (vr_step_diag)
  std::vector<cudaEvent_t> dev;      // 6 per record: around k_front and k_back on the step stream, around k_snow on the snow stream
  size_t dev_used = 0;
  int64_t diag_unit_records = 0;
  int chunk = 8;                     // records per launch of k_cells (and per pipelined copy of vr_step_block)
  int64_t launches = 0;
  bool timed = false;
};

static std::string g_create_error;

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                   \
      return VR_ERR_CUDA;                                                                            \
    }                                                                                                \
  } while (0)

// the step kernels of record `rec` of a chunk: k_front, k_back on the step stream st, k_snow on the snow stream ss
template <int EX>
static void launch_record(vr_handle *h, const DevForcing &F, int rec, double *terms, cudaStream_t st, cudaStream_t ss)
{
  const int64_t U = h->M.U;
  const unsigned gf = (unsigned)((U + VR_FRONT_CTA - 1) / VR_FRONT_CTA), gb = (unsigned)((U + VR_BACK_CTA - 1) / VR_BACK_CTA);
  int64_t gs = (U + VR_SNOW_CTA - 1) / VR_SNOW_CTA;
  const int64_t gs_max = (int64_t)h->sms * VR_SNOW_MINB * 4;     // grid-stride: a few waves at most
  if (gs > gs_max) gs = gs_max;
  constexpr size_t smf = VR_FRONT_STAGE ? (size_t)STAGE_ROWS * VR_FRONT_CTA * sizeof(double) : 0;
  cudaEvent_t *e = nullptr;
  if (h->detail && h->dev_used + 6 <= 400000) {
    while (h->dev.size() < h->dev_used + 6) { cudaEvent_t x; cudaEventCreate(&x); h->dev.push_back(x); }
    e = &h->dev[h->dev_used];
    h->dev_used += 6;
  }
  if (e) cudaEventRecord(e[0], st);
  // the units outside the snow class, in h->na slot ranges that run their k_front -> k_back sequences side by side
  const int na = h->na;
  const int64_t per = ((U + na - 1) / na + 383) / 384 * 384;      // multiple of both block sizes
  for (int a = 0; a < na; a++) {
    const int64_t ub = a * per, ue = (ub + per < U) ? ub + per : U;
    if (ub >= U) break;
    cudaStream_t sa = a == 0 ? st : h->s_a[a - 1];
    k_front<EX><<<(unsigned)((ue - ub + VR_FRONT_CTA - 1) / VR_FRONT_CTA), VR_FRONT_CTA, smf, sa>>>(h->M, F, rec, terms, ub, ue);
    if (e && a == 0) cudaEventRecord(e[1], st);
    k_back<EX & EX_ZWT><<<(unsigned)((ue - ub + VR_BACK_CTA - 1) / VR_BACK_CTA), VR_BACK_CTA, 0, sa>>>(h->M, rec, terms, ub, ue);
  }
  if (e) cudaEventRecord(e[2], st);
  (void)gf;
  if (e) cudaEventRecord(e[3], ss);
  int64_t gl = gb;
  if (gl > (int64_t)h->sms * VR_BACK_MINB * 4) gl = (int64_t)h->sms * VR_BACK_MINB * 4;
  // the snow class in h->snow_parts parts of its list, each k_snow -> k_back_list sequence on a stream of its own
  const int np = h->snow_parts;
  const unsigned gsp = (unsigned)((gs + np - 1) / np), glp = (unsigned)((gl + np - 1) / np);
  for (int p = 0; p < np; p++) {
    cudaStream_t sp = p == 0 ? ss : h->s_sp[p - 1];
    if (h->M.NF > 1) k_snow<EX, true><<<gsp, VR_SNOW_CTA, 0, sp>>>(h->M, F, rec, terms, p, np);
    else k_snow<EX, false><<<gsp, VR_SNOW_CTA, 0, sp>>>(h->M, F, rec, terms, p, np);
    if (e && p == 0) cudaEventRecord(e[5], ss);
    k_back_list<EX & EX_ZWT><<<glp, VR_BACK_CTA, 0, sp>>>(h->M, rec, terms, p, np);
  }
  if (e) cudaEventRecord(e[4], ss);
  h->diag_unit_records += U;
}

extern "C" {

void vr_default_options(vr_options *opt, int nlayer)
{
  memset(opt, 0, sizeof *opt);
  opt->nlayer = nlayer; opt->nbands = 1; opt->dt_hours = 24; opt->snow_step_hours = 24;
  opt->max_snow_temp = 0.5; opt->min_rain_temp = -0.5; opt->wind_h = 10.0;
  int n = 0;
  const int head[] = {VR_OUT_PREC, VR_OUT_EVAP, VR_OUT_RUNOFF, VR_OUT_BASEFLOW, VR_OUT_WDEW};
  for (int v : head) opt->out_vars[n++] = v;
  for (int l = 0; l < nlayer; l++) opt->out_vars[n++] = VR_OUT_SOIL_LIQ0 + l;
  const int tail[] = {VR_OUT_NET_SHORT, VR_OUT_R_NET, VR_OUT_EVAP_CANOP, VR_OUT_TRANSP_VEG, VR_OUT_EVAP_BARE,
                      VR_OUT_SUB_CANOP, VR_OUT_SUB_SNOW, VR_OUT_AERO_RESIST, VR_OUT_SURF_TEMP, VR_OUT_ALBEDO,
                      VR_OUT_REL_HUMID, VR_OUT_IN_LONG, VR_OUT_AIR_TEMP, VR_OUT_WIND, VR_OUT_SWE,
                      VR_OUT_SNOW_DEPTH, VR_OUT_SNOW_CANOPY, VR_OUT_SNOW_COVER};
  for (int v : tail) opt->out_vars[n++] = v;
  opt->n_out = n;
}

int vr_create(const vr_options *opt, int device, vr_handle **out)
{
  if (!opt || !out) { g_create_error = "null argument"; return VR_ERR_ARG; }
  *out = nullptr;
  if (opt->nlayer < 2 || opt->nlayer > VR_MAX_LAYERS || opt->nbands < 1 || opt->nbands > VR_MAX_BANDS ||
      opt->n_out < 1 || opt->n_out > VR_MAX_OUT || opt->dt_hours < 1 || opt->dt_hours > 24 || 24 % opt->dt_hours) {
    g_create_error = "vr_options out of range (nlayer 2..3, nbands 1..10, dt_hours divides 24, 1 <= n_out <= 64)";
    return VR_ERR_ARG;
  }
  if (opt->snow_step_hours < 1 || opt->snow_step_hours > opt->dt_hours || opt->dt_hours % opt->snow_step_hours) {
    g_create_error = "SNOW_STEP must divide TIME_STEP (get_global_param.c:761-770)";
    return VR_ERR_ARG;
  }
  if (opt->dt_hours < 24 && opt->snow_step_hours != opt->dt_hours) {
    g_create_error = "If the model step is smaller than daily, the snow model should run at the same time step as the rest "
                     "of the model (get_global_param.c:762-763)";
    return VR_ERR_ARG;
  }
  if (!(opt->max_snow_temp > opt->min_rain_temp)) {
    g_create_error = "MAX_SNOW_TEMP must be greater than MIN_RAIN_TEMP (calc_rainonly.c)";
    return VR_ERR_ARG;
  }
  for (int v = 0; v < opt->n_out; v++)
    if (opt->out_vars[v] < 0 || opt->out_vars[v] >= VR_NOUTVAR) { g_create_error = "unknown output variable id"; return VR_ERR_ARG; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
    return VR_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return VR_ERR_ARG; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return VR_ERR_CUDA; }
  vr_handle *h = new vr_handle();
  h->opt = *opt;
  h->device = device;
  memset(&h->M, 0, sizeof h->M);
  bool ok = (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) == cudaSuccess &&
            (e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking)) == cudaSuccess &&
            (e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking)) == cudaSuccess &&
            (e = cudaStreamCreateWithFlags(&h->s_cells, cudaStreamNonBlocking)) == cudaSuccess &&
            (e = cudaStreamCreateWithFlags(&h->s_snow, cudaStreamNonBlocking)) == cudaSuccess &&
            (e = cudaEventCreateWithFlags(&h->ev_cls, cudaEventDisableTiming)) == cudaSuccess &&
            (e = cudaEventCreateWithFlags(&h->ev_snow, cudaEventDisableTiming)) == cudaSuccess &&
            (e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) == cudaSuccess &&
            (e = cudaEventCreate(&h->ev0)) == cudaSuccess && (e = cudaEventCreate(&h->ev1)) == cudaSuccess;
  for (int b = 0; ok && b < vr_handle::NBUF; b++)
    ok = (e = cudaEventCreateWithFlags(&h->ev_in[b], cudaEventDisableTiming)) == cudaSuccess &&
         (e = cudaEventCreateWithFlags(&h->ev_comp[b], cudaEventDisableTiming)) == cudaSuccess &&
         (e = cudaEventCreateWithFlags(&h->ev_out[b], cudaEventDisableTiming)) == cudaSuccess;
  for (int b = 0; ok && b < 2; b++)
    ok = (e = cudaEventCreateWithFlags(&h->ev_units[b], cudaEventDisableTiming)) == cudaSuccess &&
         (e = cudaEventCreateWithFlags(&h->ev_cells[b], cudaEventDisableTiming)) == cudaSuccess;
  if (!ok) {
    g_create_error = cudaGetErrorString(e);
    delete h;
    return VR_ERR_CUDA;
  }
  {   // the stage of k_front exceeds the 48 kB default of dynamic shared memory
    const int smf = (int)(STAGE_ROWS * VR_FRONT_CTA * sizeof(double));
    cudaFuncSetAttribute(k_front<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
    cudaFuncSetAttribute(k_front<EX_PET>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
    cudaFuncSetAttribute(k_front<EX_ZWT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
    cudaFuncSetAttribute(k_front<EX_PET | EX_ZWT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf);
  }
  if (opt->cells_per_block_hint > 0) h->chunk = opt->cells_per_block_hint;
  { const char *d = getenv("VICRES_TIME_KERNELS"); h->detail = d && d[0] == '1'; }
  { const char *d = getenv("VICRES_SNOW_STREAM"); h->snow_stream = !(d && d[0] == '0'); }
  { const char *d = getenv("VICRES_SNOW_PARTS"); int n = d ? atoi(d) : VR_SNOW_PARTS; h->snow_parts = n < 1 ? 1 : (n > 4 ? 4 : n); }
  if (!h->snow_stream) h->snow_parts = 1;
  for (int p = 1; p < h->snow_parts; p++) {
    cudaStreamCreateWithFlags(&h->s_sp[p - 1], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_sp[p - 1], cudaEventDisableTiming);
  }
  { const char *d = getenv("VICRES_A_STREAMS"); h->na = (d && d[0] >= '1' && d[0] <= '4') ? d[0] - '0' : VR_A_STREAMS; }
  for (int a = 0; a < 3; a++) {
    cudaStreamCreateWithFlags(&h->s_a[a], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_a[a], cudaEventDisableTiming);
  }
  *out = h;
  return VR_OK;
}

void vr_destroy(vr_handle *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  h->d_cc.release(); h->d_u_cv.release(); h->d_u_af.release(); h->d_u_Tf.release(); h->d_u_Pf.release();
  h->d_u_rarc.release(); h->d_u_rmin.release(); h->d_u_RGL.release(); h->d_tm.release(); h->d_ae.release();
  h->d_init_moist.release(); h->d_state.release(); h->d_sv.release(); h->d_forc.release(); h->d_out.release();
  h->d_tl.release(); h->d_ap.release(); h->d_ccu.release(); h->d_fslots.release(); h->d_snowflag.release();
  h->d_tair0.release(); h->d_u_root.release(); h->d_cell_uptr.release(); h->d_fcol.release(); h->d_terms.release(); h->d_sums.release(); h->d_u_logical.release();
  h->d_u_cell.release(); h->d_u_tile.release(); h->d_u_flags.release(); h->d_u_aero.release();
  h->d_cell_fail.release(); h->d_cell_status.release(); h->d_month.release(); h->d_doy.release();
  h->d_order.release(); h->d_u_slot.release(); h->d_u_pos.release();
  h->d_tmu.release(); h->d_snow_total.release();
  for (auto &x : h->dev) cudaEventDestroy(x); h->d_uflag.release(); h->d_snow_list.release(); h->d_snow_cnt.release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto &t : h->evs) { cudaEventDestroy(t.u0); cudaEventDestroy(t.u1); cudaEventDestroy(t.c0); cudaEventDestroy(t.c1); }
  if (h->s_cells) cudaStreamDestroy(h->s_cells);
  if (h->s_snow) cudaStreamDestroy(h->s_snow);
  for (int p = 0; p < 3; p++) {
    if (h->s_sp[p]) cudaStreamDestroy(h->s_sp[p]);
    if (h->ev_sp[p]) cudaEventDestroy(h->ev_sp[p]);
  }
  for (int a = 0; a < 3; a++) {
    if (h->s_a[a]) cudaStreamDestroy(h->s_a[a]);
    if (h->ev_a[a]) cudaEventDestroy(h->ev_a[a]);
  }
  if (h->ev_cls) cudaEventDestroy(h->ev_cls);
  if (h->ev_snow) cudaEventDestroy(h->ev_snow);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int b = 0; b < 2; b++) {
    if (h->ev_units[b]) cudaEventDestroy(h->ev_units[b]);
    if (h->ev_cells[b]) cudaEventDestroy(h->ev_cells[b]);
  }
  for (int b = 0; b < vr_handle::NBUF; b++) {
    if (h->ev_in[b]) cudaEventDestroy(h->ev_in[b]);
    if (h->ev_comp[b]) cudaEventDestroy(h->ev_comp[b]);
    if (h->ev_out[b]) cudaEventDestroy(h->ev_out[b]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
}

const char *vr_last_error(const vr_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vr_set_veglib(vr_handle *h, const vr_veglib *lib)
{
  if (!h || !lib || lib->nclass < 0) return VR_ERR_ARG;
  const int n = lib->nclass;
  h->lib_overstory.assign(lib->overstory, lib->overstory + n);
  h->lib_rarc.assign(lib->rarc, lib->rarc + n); h->lib_rmin.assign(lib->rmin, lib->rmin + n);
  h->lib_wind_h.assign(lib->wind_h, lib->wind_h + n); h->lib_RGL.assign(lib->RGL, lib->RGL + n);
  h->lib_rad_atten.assign(lib->rad_atten, lib->rad_atten + n);
  h->lib_wind_atten.assign(lib->wind_atten, lib->wind_atten + n);
  h->lib_trunk.assign(lib->trunk_ratio, lib->trunk_ratio + n);
  h->lib_rough.assign(lib->roughness, lib->roughness + (size_t)n * 12);
  h->lib_disp.assign(lib->displacement, lib->displacement + (size_t)n * 12);
  h->lib_LAI.clear(); h->lib_albedo.clear();
  if (lib->LAI && lib->albedo) {
    h->lib_LAI.assign(lib->LAI, lib->LAI + (size_t)n * 12);
    h->lib_albedo.assign(lib->albedo, lib->albedo + (size_t)n * 12);
  }
  h->have_lib = true;
  h->have_cells = false;
  h->initialised = false;
  return VR_OK;
}

int vr_set_cells(vr_handle *h, const vr_cells *cells)
{
  if (!h || !cells) return VR_ERR_ARG;
  if (!h->have_lib) { h->err = "vr_set_veglib must be called before vr_set_cells"; return VR_ERR_STATE; }
  if (cells->ncell <= 0) { h->err = "ncell must be positive"; return VR_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  vr_veglib lib;
  lib.nclass = (int)h->lib_overstory.size();
  lib.overstory = h->lib_overstory.data(); lib.rarc = h->lib_rarc.data(); lib.rmin = h->lib_rmin.data();
  lib.wind_h = h->lib_wind_h.data(); lib.RGL = h->lib_RGL.data(); lib.rad_atten = h->lib_rad_atten.data();
  lib.wind_atten = h->lib_wind_atten.data(); lib.trunk_ratio = h->lib_trunk.data();
  lib.roughness = h->lib_rough.data(); lib.displacement = h->lib_disp.data();
  lib.LAI = h->lib_LAI.empty() ? nullptr : h->lib_LAI.data();
  lib.albedo = h->lib_albedo.empty() ? nullptr : h->lib_albedo.data();
  h->P = Prepared();
  if (!prepare(h->opt, lib, *cells, VR_CTA, h->P)) { h->err = h->P.error; return VR_ERR_ARG; }
  Prepared &P = h->P;
  const int64_t C = P.C, U = P.U;
  {
    auto &k = h->keep;
    const int NL = h->opt.nlayer, NB = h->opt.nbands;
    const int64_t T = cells->tile_ptr[C];
    k.nclass = lib.nclass;
    k.tile_ptr.assign(cells->tile_ptr, cells->tile_ptr + C + 1);
    k.tile_class.assign(cells->tile_class, cells->tile_class + T);
    k.tile_Cv.assign(cells->tile_Cv, cells->tile_Cv + T);
    k.AreaFract.assign(cells->AreaFract, cells->AreaFract + (size_t)NB * C);
    k.Tfactor.assign(cells->Tfactor, cells->Tfactor + (size_t)NB * C);
    k.init_moist.assign(cells->init_moist, cells->init_moist + (size_t)NL * C);
    k.max_moist.assign(cells->max_moist, cells->max_moist + (size_t)NL * C);
    k.depth0.assign(cells->depth, cells->depth + C);
    k.dp.assign(cells->dp, cells->dp + C);
    k.avg_temp.assign(cells->avg_temp, cells->avg_temp + C);
  }
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, h->device);
  h->cta = VR_CTA;
  h->ncta = U > 0 ? (U + VR_CTA - 1) / VR_CTA : 0;
  std::vector<double> im((size_t)P.NL * C);
  for (size_t i = 0; i < im.size(); i++) im[i] = cells->init_moist[i];
  CK(h->d_cc.upload(P.cc)); CK(h->d_cell_uptr.upload(P.cell_uptr));
  CK(h->d_u_cell.upload(P.u_cell)); CK(h->d_u_pos.upload(P.u_pos)); CK(h->d_u_tile.upload(P.u_tile)); CK(h->d_u_flags.upload(P.u_flags));
  CK(h->d_u_aero.upload(P.u_aero)); CK(h->d_cell_fail.upload(P.cell_fail)); CK(h->d_u_cv.upload(P.u_cv));
  CK(h->d_u_af.upload(P.u_af)); CK(h->d_u_Tf.upload(P.u_Tfactor)); CK(h->d_u_Pf.upload(P.u_Pfactor));
  CK(h->d_u_rarc.upload(P.u_rarc)); CK(h->d_u_rmin.upload(P.u_rmin)); CK(h->d_u_RGL.upload(P.u_RGL));
  CK(h->d_u_root.upload(P.u_root)); CK(h->d_tm.upload(P.tm)); CK(h->d_ae.upload(P.ae));
  CK(h->d_tl.upload(P.tl)); CK(h->d_ap.upload(P.ap));
  CK(h->d_init_moist.upload(im));
  CK(h->d_order.upload(P.order)); CK(h->d_u_slot.upload(P.u_slot)); CK(h->d_u_logical.upload(P.u_logical));
  CK(h->d_state.reserve((size_t)(ST_N + FO_N) * (U ? U : 1)));     // state rows, then the FrontOut rows (one L2 window)
  CK(h->d_sv.reserve((size_t)2 * SV_N * C));
  CK(h->d_cell_status.reserve(C));
  CK(cudaMemset(h->d_state.p, 0, (size_t)ST_N * (U ? U : 1) * sizeof(double)));
  CK(cudaMemset(h->d_cell_status.p, 0, C * sizeof(int32_t)));
  DevModel &M = h->M;
  memset(&M, 0, sizeof M);
  M.NL = P.NL; M.NB = P.NB; M.dt_hours = h->opt.dt_hours; M.n_out = h->opt.n_out;
  M.snow_step_hours = h->opt.snow_step_hours; M.NF = h->opt.dt_hours / h->opt.snow_step_hours;
  M.moistfract = h->opt.moistfract != 0;
  M.max_snow_temp = h->opt.max_snow_temp; M.min_rain_temp = h->opt.min_rain_temp;
  M.C = C; M.U = U; M.FC = C;
  M.cc = h->d_cc.p; M.cell_uptr = h->d_cell_uptr.p; M.fcol = nullptr;
  M.u_cell = h->d_u_cell.p; M.u_pos = h->d_u_pos.p; M.u_tile = h->d_u_tile.p; M.u_flags = h->d_u_flags.p; M.u_aero = h->d_u_aero.p;
  M.cell_fail = h->d_cell_fail.p; M.u_cv = h->d_u_cv.p; M.u_af = h->d_u_af.p; M.u_Tfactor = h->d_u_Tf.p;
  M.u_Pfactor = h->d_u_Pf.p; M.u_rarc = h->d_u_rarc.p; M.u_rmin = h->d_u_rmin.p; M.u_RGL = h->d_u_RGL.p;
  M.tl = h->d_tl.p; M.ap = h->d_ap.p;
  M.tm = h->d_tm.p; M.ae = h->d_ae.p; M.init_moist = h->d_init_moist.p; M.u_root = h->d_u_root.p;
  M.state = h->d_state.p; M.sv = h->d_sv.p; M.cell_status = h->d_cell_status.p;
  M.order = h->d_order.p; M.u_slot = h->d_u_slot.p; M.u_logical = h->d_u_logical.p;
  M.ccn = P.ccn;
  M.US = ((U ? U : 1) + 255) / 256 * 256;
  CK(h->d_ccu.reserve((size_t)P.ccn * M.US));
  CK(h->d_tmu.reserve((size_t)12 * TM_N * M.US));
  CK(cudaMemset(h->d_ccu.p, 0, (size_t)P.ccn * M.US * sizeof(double)));
  CK(cudaMemset(h->d_tmu.p, 0, (size_t)12 * TM_N * M.US * sizeof(double)));
  M.ccu = h->d_ccu.p; M.tmu = h->d_tmu.p;
  CK(h->d_uflag.reserve(U ? U : 1));
  CK(h->d_snow_list.reserve(U ? U : 1));
  M.fo = h->d_state.p + (size_t)ST_N * (U ? U : 1); M.uflag = h->d_uflag.p; M.snow_list = h->d_snow_list.p;
  h->flags_dirty = true;
  if (U > 0) {
    k_unit_consts<<<(unsigned)((U + 255) / 256), 256, 0, h->stream>>>(M, h->d_ccu.p, h->d_tmu.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  for (int v = 0; v < h->opt.n_out; v++) M.out_vars[v] = h->opt.out_vars[v];
  build_term_map(h->opt.out_vars, h->opt.n_out, M.tmap);
  h->extras = term_extras(M.tmap);
  {   // records per chunk: the caller's hint or 8 (the snow class of a chunk is every unit that needs the snow model in ANY
      // of its records, so short chunks keep that class small; measured 2 .. 8 alike, 16 3 % slower: profiles/r2_tuning.md),
      // capped so that the two term buffers stay below 24 GB (a 1M-cell, 5-band grid has 1.2e7 units)
    h->chunk = h->opt.cells_per_block_hint > 0 ? h->opt.cells_per_block_hint : 8;
    const double per_rec = 2.0 * M.tmap.n * (double)(U ? U : 1) * sizeof(double);
    const int cap = (int)(24e9 / per_rec);
    if (h->chunk > cap) h->chunk = cap < 1 ? 1 : cap;
  }
  h->have_cells = true;
  h->initialised = false;
  return VR_OK;
}

int vr_set_forcing_map(vr_handle *h, const int64_t *col, int64_t ncol)
{
  if (!h) return VR_ERR_ARG;
  if (!h->have_cells) { h->err = "vr_set_cells must be called before vr_set_forcing_map"; return VR_ERR_STATE; }
  CK(cudaSetDevice(h->device));
  if (!col) { h->M.fcol = nullptr; h->M.FC = h->M.C; return VR_OK; }
  if (ncol <= 0) return VR_ERR_ARG;
  std::vector<int64_t> v(col, col + h->M.C);
  for (int64_t x : v)
    if (x < 0 || x >= ncol) { h->err = "forcing map entry out of range"; return VR_ERR_ARG; }
  CK(h->d_fcol.upload(v));
  h->M.fcol = h->d_fcol.p;
  h->M.FC = ncol;
  return VR_OK;
}

int vr_init_state(vr_handle *h, const double *tair0)
{
  if (!h || !tair0) return VR_ERR_ARG;
  if (!h->have_cells) { h->err = "vr_set_cells must be called before vr_init_state"; return VR_ERR_STATE; }
  CK(cudaSetDevice(h->device));
  CK(h->d_tair0.reserve(h->M.FC));
  CK(cudaMemcpyAsync(h->d_tair0.p, tair0, h->M.FC * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->tair0_host.assign(tair0, tair0 + h->M.FC);
  const size_t U1 = (size_t)(h->M.U ? h->M.U : 1);
  CK(h->d_terms.reserve((size_t)h->M.tmap.n * U1 * h->chunk));
  CK(h->d_sums.reserve((size_t)h->M.tmap.n * h->M.C * h->chunk));
  if (h->M.U > 0) k_init_units<<<(unsigned)h->ncta, h->cta, 0, h->stream>>>(h->M, h->d_tair0.p, h->d_terms.p);
  k_init_cells<<<(unsigned)((h->M.C + 127) / 128), 128, 0, h->stream>>>(h->M, h->d_terms.p, h->d_sums.p, h->svb);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->s_cells));
  h->initialised = true;
  h->flags_dirty = true;
  h->rec_done = 0;
  return VR_OK;
}

// One block of records = chunks of h->chunk records.  Per record of a chunk the step stream st runs k_front,
// k_snow and k_back for all units; the chunk ends with k_cells on h->s_cells, so that the aggregation of chunk k
// overlaps the step of chunk k+1 (two term buffers).  With join the step stream waits for the last aggregation, so
// the caller sees ordinary stream semantics; the host pipeline of vr_step_block orders its copies on the events instead.
static int launch_step(vr_handle *h, const DevForcing &F0, double *out_dev, cudaStream_t st, bool join)
{
  const size_t U1 = (size_t)(h->M.U ? h->M.U : 1);
  const int CH = h->chunk;
  const size_t tsize = (size_t)h->M.tmap.n * U1 * CH;
  CK(h->d_terms.reserve(2 * tsize));
  CK(h->d_sums.reserve((size_t)h->M.tmap.n * h->M.C * CH));
  const int NS = h->M.NF > 1 ? h->M.NF + 1 : 1;          // forcing sub-records per record (sub-steps + the record itself)
  CK(h->d_fslots.reserve((size_t)CH * NS * VR_NFORCING * U1));
  CK(h->d_snow_cnt.reserve(1));
  h->M.snow_cnt = h->d_snow_cnt.p;
  if (!h->d_snow_total.p) {
    CK(h->d_snow_total.reserve(1));
    CK(cudaMemset(h->d_snow_total.p, 0, sizeof(unsigned long long)));
  }
  h->M.snow_total = h->d_snow_total.p;
  cudaEventRecord(h->ev0, st);
  cudaEventRecord(h->ev_fork, st);                       // what the caller queued on st (forcing, free output buffer)
  cudaStreamWaitEvent(h->s_cells, h->ev_fork, 0);
  if (h->flags_dirty && h->M.U > 0) {
    k_unit_flags<<<(unsigned)((h->M.U + 255) / 256), 256, 0, st>>>(h->M);
    h->launches += 1;
    h->flags_dirty = false;
  }
  int b = 0;
  for (int r0 = 0; r0 < F0.nrec; r0 += CH) {
    DevForcing F = F0;
    F.nrec = (F0.nrec - r0 < CH) ? F0.nrec - r0 : CH;
    F.rec0 = F0.rec0 + r0;
    F.month = F0.month + r0; F.doy = F0.doy + r0;
    for (int i = 0; i < VR_NFORCING; i++) F.field[i] = F0.field[i] + (size_t)r0 * NS * h->M.FC;
    F.snowflag = F0.snowflag ? F0.snowflag + (size_t)r0 * h->M.FC : nullptr;
    F.rec_abs = F0.rec_abs + r0;
    b = (int)(h->pair_seq & 1);
    double *terms = h->d_terms.p + (size_t)b * tsize;
    vr_handle::EvTriple *t = nullptr;
    if (h->evs_used < 8192) {                      // beyond that the oldest launches simply go untimed
      if (h->evs_used == h->evs.size()) {
        vr_handle::EvTriple n;
        cudaEventCreate(&n.u0); cudaEventCreate(&n.u1); cudaEventCreate(&n.c0); cudaEventCreate(&n.c1);
        n.nrec = 0;
        h->evs.push_back(n);
      }
      t = &h->evs[h->evs_used++];
      t->nrec = F.nrec;
    }
    if (h->pair_seq >= 2) cudaStreamWaitEvent(st, h->ev_cells[b], 0);     // chunk i-2 has been aggregated out of terms[b]
    if (t) cudaEventRecord(t->u0, st);
    if (h->M.U > 0) {
      cudaStream_t ss = h->snow_stream ? h->s_snow : st;
      cudaMemsetAsync(h->d_snow_cnt.p, 0, sizeof(int32_t), st);
      k_forcing_slots<<<dim3((unsigned)((h->M.U + 255) / 256), (unsigned)(F.nrec * NS)), 256, 0, st>>>(h->M, F, h->d_fslots.p);
      F.slots = h->d_fslots.p;
      k_classify<<<(unsigned)((h->M.U + 255) / 256), 256, 0, st>>>(h->M, F);
      cudaEventRecord(h->ev_cls, st);
      cudaStreamWaitEvent(ss, h->ev_cls, 0);           // the snow stream starts when the chunk's classes are known
      for (int p = 1; p < h->snow_parts; p++) cudaStreamWaitEvent(h->s_sp[p - 1], h->ev_cls, 0);
      for (int a = 1; a < h->na; a++) cudaStreamWaitEvent(h->s_a[a - 1], h->ev_cls, 0);
      h->launches += 2;
      for (int rec = 0; rec < F.nrec; rec++) {
        switch (h->extras) {      // one set of kernel instances per set of optional results
          case 0: launch_record<0>(h, F, rec, terms, st, ss); break;
          case EX_PET: launch_record<EX_PET>(h, F, rec, terms, st, ss); break;
          case EX_ZWT: launch_record<EX_ZWT>(h, F, rec, terms, st, ss); break;
          default: launch_record<EX_PET | EX_ZWT>(h, F, rec, terms, st, ss); break;
        }
        h->launches += 2 + 2 * h->snow_parts;
      }
      cudaEventRecord(h->ev_snow, ss);
      cudaStreamWaitEvent(st, h->ev_snow, 0);          // both classes have finished the chunk
      for (int p = 1; p < h->snow_parts; p++) {
        cudaEventRecord(h->ev_sp[p - 1], h->s_sp[p - 1]);
        cudaStreamWaitEvent(st, h->ev_sp[p - 1], 0);
      }
      for (int a = 1; a < h->na; a++) {
        cudaEventRecord(h->ev_a[a - 1], h->s_a[a - 1]);
        cudaStreamWaitEvent(st, h->ev_a[a - 1], 0);
      }
    }
    if (t) cudaEventRecord(t->u1, st);
    cudaEventRecord(h->ev_units[b], st);
    cudaStreamWaitEvent(h->s_cells, h->ev_units[b], 0);
    if (t) cudaEventRecord(t->c0, h->s_cells);
    k_cells<<<dim3((unsigned)((h->M.C + 255) / 256), (unsigned)F.nrec), 256, 0, h->s_cells>>>(h->M, F, terms, out_dev, h->svb);
    h->svb ^= 1;
    if (t) cudaEventRecord(t->c1, h->s_cells);
    cudaEventRecord(h->ev_cells[b], h->s_cells);
    h->pair_seq++;
    h->launches += 1;
  }
  if (join) cudaStreamWaitEvent(st, h->ev_cells[b], 0);
  cudaEventRecord(h->ev1, st);
  h->timed = true;
  CK(cudaGetLastError());
  return VR_OK;
}

int vr_step_block_device(vr_handle *h, const vr_forcing *f, double *out_dev, void *cuda_stream)
{
  if (!h || !f || !out_dev || f->nrec <= 0) return VR_ERR_ARG;
  if (!h->initialised) { h->err = "vr_init_state (or vr_set_state) must precede vr_step_block"; return VR_ERR_STATE; }
  CK(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
  CK(h->d_month.reserve(f->nrec)); CK(h->d_doy.reserve(f->nrec));
  CK(cudaMemcpyAsync(h->d_month.p, f->month, f->nrec * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->d_doy.p, f->day_in_year, f->nrec * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  DevForcing F;
  F.slots = nullptr;
  F.snowflag = f->snowflag; F.rec_abs = (int)h->rec_done;
  F.nrec = f->nrec; F.rec0 = 0; F.nrec_total = f->nrec; F.month = h->d_month.p; F.doy = h->d_doy.p;
  for (int i = 0; i < VR_NFORCING; i++) {
    if (!f->field[i]) { h->err = "null forcing field"; return VR_ERR_ARG; }
    F.field[i] = f->field[i];
  }
  h->rec_done += f->nrec;
  return launch_step(h, F, out_dev, st, true);
}

// Host-buffer entry point.  The block of records is cut into chunks of h->chunk records that flow
// through four streams (copies in, k_units, k_cells, copies out): H2D of later chunks, the kernels of chunk k
// and D2H of earlier chunks overlap (PCIe is full duplex), with a ring of device buffers per direction.  Forcing
// [field][rec][col] and outputs
// [var][rec][cell] are record-major, so a chunk is one contiguous copy per field / variable.
int vr_step_block(vr_handle *h, const vr_forcing *f, double *out_host, int32_t *cell_status_host)
{
  if (!h || !f || !out_host || f->nrec <= 0) return VR_ERR_ARG;
  if (!h->initialised) { h->err = "vr_init_state (or vr_set_state) must precede vr_step_block"; return VR_ERR_STATE; }
  for (int i = 0; i < VR_NFORCING; i++)
    if (!f->field[i]) { h->err = "null forcing field"; return VR_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  const int nrec = f->nrec, n_out = h->opt.n_out, CH = h->chunk < nrec ? h->chunk : nrec;
  const size_t FC = (size_t)h->M.FC, C = (size_t)h->M.C;
  const size_t NS = h->M.NF > 1 ? h->M.NF + 1 : 1;               // forcing sub-records per record
  const size_t in_chunk = (size_t)CH * NS * FC * VR_NFORCING, out_chunk = (size_t)CH * C * n_out;
  // Depth of the buffer rings.  The kernels of a chunk take 4 .. 9 ms with the season (share of units under snow), the
  // D2H of its outputs is constant, so a shallow ring makes every chunk cost max(kernels, copy); a deep one lets the
  // kernels run ahead through the summer and the copies catch up in winter.  As deep as the block has chunks, NBUF, and
  // a third of the free device memory allow (VICRES_PIPE_DEPTH overrides), at least 2.
  int NB = (nrec + CH - 1) / CH;
  if (NB > vr_handle::NBUF) NB = vr_handle::NBUF;
  {
    size_t free_b = 0, total_b = 0;
    const size_t held = (h->d_forc.n + h->d_out.n) * sizeof(double), per = (in_chunk + out_chunk) * sizeof(double);
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const size_t fit = ((free_b + held) / 3) / (per ? per : 1);
      if ((size_t)NB > fit) NB = (int)fit;
    }
    if (const char *e = getenv("VICRES_PIPE_DEPTH")) NB = atoi(e) < vr_handle::NBUF ? atoi(e) : vr_handle::NBUF;
    if (NB < 2) NB = 2;
  }
  const bool flags = NS > 1 && f->snowflag;
  if (flags) {
    CK(h->d_snowflag.reserve((size_t)nrec * FC));
    CK(cudaMemcpyAsync(h->d_snowflag.p, f->snowflag, (size_t)nrec * FC * sizeof(int32_t), cudaMemcpyHostToDevice, h->s_in));
  }
  CK(h->d_forc.reserve(NB * in_chunk));
  CK(h->d_out.reserve(NB * out_chunk));
  CK(h->d_month.reserve(nrec)); CK(h->d_doy.reserve(nrec));
  CK(cudaMemcpyAsync(h->d_month.p, f->month, nrec * sizeof(int32_t), cudaMemcpyHostToDevice, h->s_in));
  CK(cudaMemcpyAsync(h->d_doy.p, f->day_in_year, nrec * sizeof(int32_t), cudaMemcpyHostToDevice, h->s_in));
  // VICRES_PIPE_TRACE=1: when each chunk's copy in, kernels and copy out ended, in ms after the call began (stderr)
  const bool trace = getenv("VICRES_PIPE_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  if (trace) {
    tev.resize(1 + 3 * (size_t)((nrec + CH - 1) / CH));
    for (cudaEvent_t &e : tev) cudaEventCreate(&e);
    cudaEventRecord(tev[0], h->s_in);
  }
  int k = 0;
  for (int r0 = 0; r0 < nrec; r0 += CH, k++) {
    const int nr = (nrec - r0 < CH) ? nrec - r0 : CH, b = k % NB;
    double *din = h->d_forc.p + (size_t)b * in_chunk, *dout = h->d_out.p + (size_t)b * out_chunk;
    if (k >= NB) CK(cudaStreamWaitEvent(h->s_in, h->ev_comp[b], 0));      // the kernels of chunk k-NB have consumed din
    DevForcing F;
    F.slots = nullptr;
    F.snowflag = flags ? h->d_snowflag.p + (size_t)r0 * FC : nullptr;
    F.rec_abs = (int)h->rec_done + r0;
    F.nrec = nr; F.rec0 = 0; F.nrec_total = nr; F.month = h->d_month.p + r0; F.doy = h->d_doy.p + r0;
    // the nine fields of the chunk: one strided copy when the caller's fields are equally spaced in memory (one array
    // [field][rec][col], the usual case), else one copy per field
    const size_t row_in = (size_t)nr * NS * FC * sizeof(double);
    const ptrdiff_t gap = (const char *)f->field[1] - (const char *)f->field[0];
    const size_t MAX_PITCH = 2147483647;                 // cudaDeviceProp::memPitch: the widest pitch a strided copy takes
    bool even = gap >= (ptrdiff_t)row_in && (size_t)gap <= MAX_PITCH;
    for (int i = 2; even && i < VR_NFORCING; i++) even = ((const char *)f->field[i] - (const char *)f->field[i - 1]) == gap;
    if (even)
      CK(cudaMemcpy2DAsync(din, row_in, f->field[0] + (size_t)r0 * NS * FC, (size_t)gap, row_in, VR_NFORCING, cudaMemcpyHostToDevice,
                           h->s_in));
    for (int i = 0; i < VR_NFORCING; i++) {
      if (!even)
        CK(cudaMemcpyAsync(din + (size_t)i * nr * NS * FC, f->field[i] + (size_t)r0 * NS * FC, row_in, cudaMemcpyHostToDevice, h->s_in));
      F.field[i] = din + (size_t)i * nr * NS * FC;
    }
    CK(cudaEventRecord(h->ev_in[b], h->s_in));
    if (trace) cudaEventRecord(tev[1 + 3 * k], h->s_in);
    CK(cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
    if (k >= NB) CK(cudaStreamWaitEvent(h->stream, h->ev_out[b], 0));     // D2H of chunk k-NB has drained dout
    // not joined: k_cells(k) trails behind k_units(k+1) on its own stream; the rings of copy buffers
    // keep H2D / kernels / D2H flowing although the outputs of a chunk leave one kernel later
    int rc = launch_step(h, F, dout, h->stream, false);
    if (rc != VR_OK) return rc;
    CK(cudaEventRecord(h->ev_comp[b], h->s_cells));                       // forcing consumed and outputs written
    if (trace) cudaEventRecord(tev[2 + 3 * k], h->s_cells);
    CK(cudaStreamWaitEvent(h->s_out, h->ev_comp[b], 0));
    // the chunk's records of every output variable: one strided copy (rows of nr * C values, nrec * C apart on the host)
    if ((size_t)nrec * C * sizeof(double) <= MAX_PITCH)
      CK(cudaMemcpy2DAsync(out_host + (size_t)r0 * C, (size_t)nrec * C * sizeof(double), dout, (size_t)nr * C * sizeof(double),
                           (size_t)nr * C * sizeof(double), n_out, cudaMemcpyDeviceToHost, h->s_out));
    else
      for (int v = 0; v < n_out; v++)
        CK(cudaMemcpyAsync(out_host + ((size_t)v * nrec + r0) * C, dout + (size_t)v * nr * C, (size_t)nr * C * sizeof(double),
                           cudaMemcpyDeviceToHost, h->s_out));
    CK(cudaEventRecord(h->ev_out[b], h->s_out));
    if (trace) cudaEventRecord(tev[3 + 3 * k], h->s_out);
  }
  if (cell_status_host) {
    CK(cudaStreamWaitEvent(h->s_out, h->ev_comp[(k - 1) % NB], 0));
    CK(cudaMemcpyAsync(cell_status_host, h->d_cell_status.p, C * sizeof(int32_t), cudaMemcpyDeviceToHost, h->s_out));
  }
  h->rec_done += nrec;
  CK(cudaStreamSynchronize(h->s_in));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->s_cells));
  CK(cudaStreamSynchronize(h->s_out));
  if (trace) {
    fprintf(stderr, "vr_step_block pipeline (ring depth %d): chunk, copy in / kernels / copy out ended at (ms)\n", NB);
    for (int i = 0; i < k; i++) {
      float a = 0, b2 = 0, c2 = 0;
      cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * i]); cudaEventElapsedTime(&b2, tev[0], tev[2 + 3 * i]);
      cudaEventElapsedTime(&c2, tev[0], tev[3 + 3 * i]);
      fprintf(stderr, "  %3d %9.3f %9.3f %9.3f\n", i, a, b2, c2);
    }
    for (cudaEvent_t &e : tev) cudaEventDestroy(e);
  }
  if (cell_status_host)
    for (size_t c = 0; c < C; c++)
      if (cell_status_host[c]) return VR_ERR_CELL;
  return VR_OK;
}

int vr_full_energy_compat(vr_handle *h, int64_t rec, int32_t month, int32_t day_in_year, const double *atmos, const int32_t *snowflag,
                          double *out, int32_t *cell_status)
{
  if (!h || !atmos || !out) return VR_ERR_ARG;
  if (!h->initialised) { h->err = "vr_init_state (or vr_set_state) must precede vr_full_energy_compat"; return VR_ERR_STATE; }
  if (rec != (int64_t)h->rec_done) {
    h->err = "vr_full_energy_compat: record " + std::to_string(rec) + " is not the next record of the run (" +
             std::to_string(h->rec_done) + ")";
    return VR_ERR_STATE;
  }
  const size_t NS = h->M.NF > 1 ? h->M.NF + 1 : 1;
  vr_forcing f;
  memset(&f, 0, sizeof f);
  f.nrec = 1; f.month = &month; f.day_in_year = &day_in_year;
  for (int i = 0; i < VR_NFORCING; i++) f.field[i] = atmos + (size_t)i * NS * (size_t)h->M.FC;
  f.snowflag = snowflag;
  return vr_step_block(h, &f, out, cell_status);
}

int vr_synchronize(vr_handle *h)
{
  if (!h) return VR_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->s_cells));
  return VR_OK;
}

int64_t vr_state_size(const vr_handle *h)
{
  if (!h || !h->have_cells) return 0;
  return (int64_t)(((size_t)ST_N * h->M.U + (size_t)SV_N * h->M.C) * sizeof(double));
}

int vr_get_state(vr_handle *h, void *blob)
{
  if (!h || !blob) return VR_ERR_ARG;
  if (!h->initialised) { h->err = "no state yet"; return VR_ERR_STATE; }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->s_cells));
  char *b = (char *)blob;
  const size_t ns = (size_t)ST_N * h->M.U * sizeof(double), nv = (size_t)SV_N * h->M.C * sizeof(double);
  CK(cudaMemcpy(b, h->d_state.p, ns, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b + ns, h->d_sv.p + (size_t)h->svb * SV_N * h->M.C, nv, cudaMemcpyDeviceToHost));
  return VR_OK;
}

int vr_set_state(vr_handle *h, const void *blob)
{
  if (!h || !blob) return VR_ERR_ARG;
  if (!h->have_cells) { h->err = "vr_set_cells must be called before vr_set_state"; return VR_ERR_STATE; }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->s_cells));
  const char *b = (const char *)blob;
  const size_t ns = (size_t)ST_N * h->M.U * sizeof(double), nv = (size_t)SV_N * h->M.C * sizeof(double);
  CK(cudaMemcpy(h->d_state.p, b, ns, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_sv.p + (size_t)h->svb * SV_N * h->M.C, b + ns, nv, cudaMemcpyHostToDevice));
  h->initialised = true;
  h->flags_dirty = true;
  return VR_OK;
}

// ---- the reference's state file (write_model_state.c:100-300, open_state_file.c, read_initial_model_state.c,
// initialize_model_state.c:233-330, 590-633) ------------------------------------------------------------------
namespace {
// slot of unit (cell c, tile t, band b), or -1: the construction order of prepare() (sorted cells, tiles with Cv > 0,
// bands with AreaFract > 0) gives the unit's index in reference order, P.u_slot its thread slot
std::vector<int64_t> unit_slots(const vr_handle *h)
{
  const Prepared &P = h->P;
  const auto &k = h->keep;
  const int64_t C = P.C;
  const int NB = h->opt.nbands;
  const int64_t T = k.tile_ptr[C];
  std::vector<int64_t> slot((size_t)T * NB, -1);
  int64_t u = 0;
  for (int64_t cs = 0; cs < C; cs++) {
    const int64_t c = P.order[cs];
    for (int64_t t = k.tile_ptr[c]; t < k.tile_ptr[c + 1]; t++) {
      if (!(k.tile_Cv[t] > 0.0)) continue;
      for (int b = 0; b < NB; b++) {
        if (!(k.AreaFract[(size_t)b * C + c] > 0)) continue;
        slot[(size_t)t * NB + b] = P.u_slot[u++];
      }
    }
  }
  return slot;
}
}  // namespace

int vr_write_state_file(vr_handle *h, const char *path, int32_t year, int32_t month, int32_t day, int32_t binary,
                        const int32_t *gridcel)
{
  if (!h || !path || !gridcel) return VR_ERR_ARG;
  if (!h->initialised) { h->err = "no state yet"; return VR_ERR_STATE; }
  const int64_t C = h->M.C, U = h->M.U;
  const int NL = h->opt.nlayer, NB = h->opt.nbands, NN = 3, ncls = h->keep.nclass;
  std::vector<double> st((size_t)ST_N * (U ? U : 1) + (size_t)SV_N * C);
  int rc = vr_get_state(h, st.data());
  if (rc != VR_OK) return rc;
  FILE *f = fopen(path, binary ? "wb" : "w");
  if (!f) { h->err = std::string("cannot open ") + path; return VR_ERR_ARG; }
  const std::vector<int64_t> slot = unit_slots(h);
  const auto &k = h->keep;
  auto S = [&](int w, int64_t sl) { return st[(size_t)w * U + sl]; };
  auto wi = [&](int v) { fwrite(&v, sizeof(int), 1, f); };
  auto wd = [&](double v) { fwrite(&v, sizeof(double), 1, f); };
  if (binary) { wi(year); wi(month); wi(day); wi(NL); wi(NN); }
  else fprintf(f, "%i %i %i\n%i %i\n", year, month, day, NL, NN);
  for (int64_t c = 0; c < C; c++) {
    const int64_t t0 = k.tile_ptr[c], t1 = k.tile_ptr[c + 1];
    const bool has_bare = t1 > t0 && k.tile_class[t1 - 1] == ncls;
    const int Nveg = (int)(t1 - t0) - (has_bare ? 1 : 0);
    // QUICK_FLUX node layout, initialize_model_state.c:311-317
    const double dz[3] = {k.depth0[c], k.depth0[c], 2. * (k.dp[c] - 1.5 * k.depth0[c])}, zs[3] = {0., k.depth0[c], k.dp[c]};
    if (binary) {
      wi(gridcel[c]); wi(Nveg); wi(NB);
      const int nbytes = NN * 8 * 2 + (Nveg + 1) * NB * 2 * 4 + (Nveg + 1) * NB * NL * 8 * 2 + Nveg * NB * 8 +
                         (Nveg + 1) * NB * (4 + 1 + 9 * 8) + (Nveg + 1) * NB * NN * 8;       // write_model_state.c:120-135
      wi(nbytes);
      for (int i = 0; i < NN; i++) wd(dz[i]);
      for (int i = 0; i < NN; i++) wd(zs[i]);
    }
    else {
      fprintf(f, "%i %i %i", gridcel[c], Nveg, NB);
      for (int i = 0; i < NN; i++) fprintf(f, " %f ", dz[i]);
      for (int i = 0; i < NN; i++) fprintf(f, " %f ", zs[i]);
      fprintf(f, "\n");
    }
    double surf = h->tair0_host.empty() ? 0. : h->tair0_host[h->P.C == (int64_t)h->tair0_host.size() ? c : 0];
    if (surf < -1.) surf = -1.;
    for (int veg = 0; veg <= Nveg; veg++) {
      const int64_t t = veg < Nveg ? t0 + veg : (has_bare ? t1 - 1 : -1);
      for (int b = 0; b < NB; b++) {
        const int64_t sl = t >= 0 ? slot[(size_t)t * NB + b] : -1;
        // tiles / bands without area keep what initialize_model_state gave them: the initial moisture everywhere,
        // node temperatures only for tiles with Cv > 0 (:336-445)
        double moist[MAXL], wdew = 0., sn[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, T[3] = {surf, surf, k.avg_temp[c]};
        if (t < 0) T[0] = T[1] = T[2] = 0.;
        int last_snow = (int)MISSINGV, melting = 0;
        for (int l = 0; l < NL; l++) {
          moist[l] = k.init_moist[(size_t)l * C + c];
          if (moist[l] > k.max_moist[(size_t)l * C + c]) moist[l] = k.max_moist[(size_t)l * C + c];
        }
        if (sl >= 0) {
          for (int l = 0; l < NL; l++) moist[l] = S(ST_MOIST0 + l, sl);
          wdew = S(ST_WDEW, sl);
          last_snow = (int)S(ST_LAST_SNOW, sl); melting = (int)S(ST_MELTING, sl);
          const int w[9] = {ST_COVERAGE, ST_SWQ, ST_SURF_TEMP, ST_SURF_WATER, ST_PACK_TEMP, ST_PACK_WATER, ST_DENSITY, ST_COLDCONTENT,
                            ST_SNOW_CANOPY};
          for (int i = 0; i < 9; i++) sn[i] = S(w[i], sl);
          T[0] = S(ST_T0, sl); T[1] = S(ST_T1, sl);
          // the third node is diagnosed from T1 each step (calc_surf_energy_bal.c:588, Zsum_node[2] = dp)
          T[2] = k.avg_temp[c] + (T[1] - k.avg_temp[c]) * exp(-(k.dp[c] - k.depth0[c]) / k.dp[c]);
        }
        if (binary) {
          wi(veg); wi(b);
          for (int l = 0; l < NL; l++) wd(moist[l]);
          for (int l = 0; l < NL; l++) wd(0.);
          if (veg < Nveg) wd(wdew);
          wi(last_snow);
          char m = (char)melting; fwrite(&m, 1, 1, f);
          for (int i = 0; i < 9; i++) wd(sn[i]);
          for (int i = 0; i < NN; i++) wd(T[i]);
        }
        else {
          fprintf(f, "%i %i", veg, b);
          for (int l = 0; l < NL; l++) fprintf(f, " %f", moist[l]);
          for (int l = 0; l < NL; l++) fprintf(f, " %f", 0.);
          if (veg < Nveg) fprintf(f, " %f", wdew);
          fprintf(f, " %i %i %f %f %f %f %f %f %f %f %f", last_snow, melting, sn[0], sn[1], sn[2], sn[3], sn[4], sn[5], sn[6], sn[7], sn[8]);
          for (int i = 0; i < NN; i++) fprintf(f, " %f", T[i]);
          fprintf(f, "\n");
        }
      }
    }
  }
  fclose(f);
  return VR_OK;
}

int vr_read_state_file(vr_handle *h, const char *path, int32_t binary, const double *tair0, const int32_t *gridcel)
{
  if (!h || !path || !tair0 || !gridcel) return VR_ERR_ARG;
  if (!h->have_cells) { h->err = "vr_set_cells must be called before vr_read_state_file"; return VR_ERR_STATE; }
  if (h->M.fcol) { h->err = "state files and a forcing map (parameter-set axis) are not combined"; return VR_ERR_UNSUPPORTED; }
  FILE *f = fopen(path, binary ? "rb" : "r");
  if (!f) { h->err = std::string("cannot open ") + path; return VR_ERR_ARG; }
  const int64_t C = h->M.C, U = h->M.U;
  const int NL = h->opt.nlayer, NB = h->opt.nbands, ncls = h->keep.nclass;
  const auto &k = h->keep;
  const std::vector<int64_t> slot = unit_slots(h);
  std::vector<double> st((size_t)ST_N * (U ? U : 1) + (size_t)SV_N * C, 0.0);
  auto S = [&](int w, int64_t sl) -> double & { return st[(size_t)w * U + sl]; };
  bool ok = true;
  auto ri = [&]() { int v = 0; if (binary) ok = ok && fread(&v, sizeof(int), 1, f) == 1; else ok = ok && fscanf(f, "%d", &v) == 1; return v; };
  auto rd = [&]() { double v = 0; if (binary) ok = ok && fread(&v, sizeof(double), 1, f) == 1; else ok = ok && fscanf(f, "%lf", &v) == 1; return v; };
  ri(); ri(); ri();                                        // date of the state (check_state_file.c:60-70)
  const int fl = ri(), fn = ri();
  if (!ok || fl != NL) { fclose(f); h->err = "state file: number of soil layers does not match"; return VR_ERR_ARG; }
  if (fn != 3) { fclose(f); h->err = "state file: number of soil thermal nodes is not 3 (QUICK_FLUX)"; return VR_ERR_UNSUPPORTED; }
  std::vector<char> seen(C, 0);
  std::vector<int64_t> of_cell;                            // gridcel -> cell index (cells come in any order)
  {
    int32_t mx = 0;
    for (int64_t c = 0; c < C; c++) if (gridcel[c] > mx) mx = gridcel[c];
    of_cell.assign((size_t)mx + 1, -1);
    for (int64_t c = 0; c < C; c++) if (gridcel[c] >= 0) of_cell[gridcel[c]] = c;
  }
  for (;;) {
    const int cellnum = ri();
    if (!ok) break;                                        // end of file
    const int fveg = ri(), fband = ri();
    if (binary) ri();                                      // Nbytes
    for (int i = 0; i < 6; i++) rd();                      // dz_node, Zsum_node: recomputed from depth and dp
    const int64_t c = (cellnum >= 0 && cellnum < (int)of_cell.size()) ? of_cell[cellnum] : -1;
    int Nveg = -1;
    int64_t t0 = 0, t1 = 0;
    bool has_bare = false;
    if (c >= 0) {
      t0 = k.tile_ptr[c]; t1 = k.tile_ptr[c + 1];
      has_bare = t1 > t0 && k.tile_class[t1 - 1] == ncls;
      Nveg = (int)(t1 - t0) - (has_bare ? 1 : 0);
      if (fveg != Nveg || fband != NB) {
        fclose(f);
        h->err = "state file: vegetation tiles or snow bands of cell " + std::to_string(cellnum) + " do not match the parameter files";
        return VR_ERR_ARG;
      }
      seen[c] = 1;
    }
    for (int veg = 0; veg <= fveg && ok; veg++)
      for (int b = 0; b < fband && ok; b++) {
        ri(); ri();
        double moist[MAXL] = {0, 0, 0}, wdew = 0., sn[9], T[3];
        for (int l = 0; l < fl; l++) moist[l] = rd();
        for (int l = 0; l < fl; l++) rd();                 // ice (Nfrost == 1)
        if (veg < fveg) wdew = rd();
        const int last_snow = ri();
        int melting;
        if (binary) { char m = 0; ok = ok && fread(&m, 1, 1, f) == 1; melting = m; } else melting = ri();
        for (int i = 0; i < 9; i++) sn[i] = rd();
        for (int i = 0; i < 3; i++) T[i] = rd();
        if (c < 0 || !ok) continue;
        const int64_t t = veg < Nveg ? t0 + veg : (has_bare ? t1 - 1 : -1);
        const int64_t sl = t >= 0 ? slot[(size_t)t * NB + b] : -1;
        if (sl < 0) continue;
        for (int l = 0; l < NL; l++) {                     // initialize_model_state.c:240-250
          const double mm = k.max_moist[(size_t)l * C + c];
          S(ST_MOIST0 + l, sl) = moist[l] > mm ? mm : moist[l];
        }
        S(ST_WDEW, sl) = wdew;
        double coverage = sn[0], swq = sn[1], surf_temp = sn[2], surf_water = sn[3], pack_temp = sn[4], pack_water = sn[5],
               density = sn[6], coldcontent = sn[7], snow_canopy = sn[8];
        const double depth = density > 0. ? 1000. * swq / density : 0.;     // read_initial_model_state.c:341-343
        double pack_swq, surf_swq;                          // initialize_model_state.c:288-305
        if (swq > 0.125) { pack_swq = swq - 0.125; surf_swq = 0.125; }
        else { pack_swq = 0; surf_swq = swq; pack_temp = 0; }
        if (surf_water > 0.035 * surf_swq) { pack_water += surf_water - (0.035 * surf_swq); surf_water = 0.035 * surf_swq; }
        if (pack_water > 0.035 * pack_swq) pack_water = 0.035 * pack_swq;
        S(ST_SWQ, sl) = swq; S(ST_SURF_TEMP, sl) = surf_temp; S(ST_PACK_TEMP, sl) = pack_temp; S(ST_SURF_WATER, sl) = surf_water;
        S(ST_PACK_WATER, sl) = pack_water; S(ST_DENSITY, sl) = density; S(ST_DEPTH, sl) = depth; S(ST_ALBEDO, sl) = 0.;
        S(ST_COLDCONTENT, sl) = coldcontent; S(ST_COVERAGE, sl) = coverage; S(ST_SNOW_CANOPY, sl) = snow_canopy;
        S(ST_TMP_INT, sl) = 0.; S(ST_LAST_SNOW, sl) = (double)last_snow; S(ST_MELTING, sl) = (double)melting;
        S(ST_T0, sl) = T[0]; S(ST_T1, sl) = T[1];
        const double tk = T[0] + KELVIN;                    // initialize_model_state.c:623-626
        S(ST_LONGUNDEROUT, sl) = STEFAN_B * tk * tk * tk * tk;
        S(ST_TFOLIAGE, sl) = tair0[c] + k.Tfactor[(size_t)b * C + c];
        S(ST_FB_SNOW, sl) = 0.; S(ST_FB_FOL, sl) = 0.;
      }
  }
  fclose(f);
  for (int64_t c = 0; c < C; c++)
    if (!seen[c]) { h->err = "state file: cell " + std::to_string(gridcel[c]) + " not found"; return VR_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  int rc = vr_set_state(h, st.data());
  if (rc != VR_OK) return rc;
  h->tair0_host.assign(tair0, tair0 + C);
  // put_data(rec = -nrecs): the storage terms of the restored state (vicNl.c:287)
  const size_t U1 = (size_t)(U ? U : 1);
  CK(h->d_terms.reserve((size_t)h->M.tmap.n * U1 * h->chunk));
  CK(h->d_sums.reserve((size_t)h->M.tmap.n * C * h->chunk));
  if (U > 0) k_restored_units<<<(unsigned)h->ncta, h->cta, 0, h->stream>>>(h->M, h->d_terms.p);
  k_init_cells<<<(unsigned)((C + 127) / 128), 128, 0, h->stream>>>(h->M, h->d_terms.p, h->d_sums.p, h->svb);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  h->initialised = true;
  h->flags_dirty = true;
  return VR_OK;
}

int64_t vr_num_units(const vr_handle *h) { return (h && h->have_cells) ? h->M.U : 0; }
int64_t vr_num_launches(const vr_handle *h) { return h ? h->launches : 0; }

int vr_last_kernel_ms(vr_handle *h, float *ms)
{
  if (!h || !ms) return VR_ERR_ARG;
  if (!h->timed) { h->err = "no step launched yet"; return VR_ERR_STATE; }
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return VR_OK;
}

int vr_kernel_times(vr_handle *h, double *units_ms, double *cells_ms, int64_t *launch_pairs, int64_t *records)
{
  if (!h || !units_ms || !cells_ms || !launch_pairs || !records) return VR_ERR_ARG;
  CK(cudaSetDevice(h->device));
  *units_ms = 0; *cells_ms = 0; *launch_pairs = 0; *records = 0;
  for (size_t i = 0; i < h->evs_used; i++) {
    vr_handle::EvTriple &t = h->evs[i];
    float a = 0, b = 0;
    CK(cudaEventSynchronize(t.c1));
    CK(cudaEventElapsedTime(&a, t.u0, t.u1));
    CK(cudaEventElapsedTime(&b, t.c0, t.c1));
    *units_ms += a; *cells_ms += b; *launch_pairs += 1; *records += t.nrec;
  }
  h->evs_used = 0;
  return VR_OK;
}

// diagnostic: since the last call -- out[0..2] = summed durations (ms) of k_front, k_snow, k_back (only with
// VICRES_TIME_KERNELS=1 in the environment at vr_create, else 0), out[3] = records timed, out[4] = unit-records that ran
// the snow instance (k_snow), out[5] = unit-records in total, out[6] = summed durations of k_back_list
int vr_step_diag(vr_handle *h, double *out)
{
  if (!h || !out) return VR_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 7; i++) out[i] = 0;
  for (size_t i = 0; i + 6 <= h->dev_used; i += 6) {
    float a = 0, b = 0, c = 0, d = 0;
    cudaEventElapsedTime(&a, h->dev[i], h->dev[i + 1]);
    cudaEventElapsedTime(&b, h->dev[i + 3], h->dev[i + 5]);
    cudaEventElapsedTime(&c, h->dev[i + 1], h->dev[i + 2]);
    cudaEventElapsedTime(&d, h->dev[i + 5], h->dev[i + 4]);
    out[0] += a; out[1] += b; out[2] += c; out[3] += 1; out[6] += d;
  }
  h->dev_used = 0;
  if (h->d_snow_total.p) {
    unsigned long long n = 0;
    CK(cudaMemcpy(&n, h->d_snow_total.p, sizeof n, cudaMemcpyDeviceToHost));
    CK(cudaMemset(h->d_snow_total.p, 0, sizeof n));
    out[4] = (double)n;
  }
  out[5] = (double)h->diag_unit_records;
  h->diag_unit_records = 0;
  return VR_OK;
}

// aggregation type of a VR_OUT_* variable (output_list_utils.c:296-385): 0 END, 1 SUM, 2 AVG
static int agg_kind(int var)
{
  switch (var) {
    case VR_OUT_ASAT: case VR_OUT_ROOTMOIST: case VR_OUT_SNOW_CANOPY: case VR_OUT_SNOW_COVER: case VR_OUT_SNOW_DEPTH:
    case VR_OUT_SOIL_LIQ0: case VR_OUT_SOIL_LIQ1: case VR_OUT_SOIL_LIQ2: case VR_OUT_SOIL_WET: case VR_OUT_SWE: case VR_OUT_WDEW:
    case VR_OUT_ZWT: case VR_OUT_ZWT_LUMPED:
      return 0;
    case VR_OUT_BASEFLOW: case VR_OUT_DELINTERCEPT: case VR_OUT_DELSOILMOIST: case VR_OUT_DELSWE: case VR_OUT_EVAP:
    case VR_OUT_EVAP_BARE: case VR_OUT_EVAP_CANOP: case VR_OUT_INFLOW: case VR_OUT_PET_SATSOIL: case VR_OUT_PET_H2OSURF:
    case VR_OUT_PET_SHORT: case VR_OUT_PET_TALL: case VR_OUT_PET_NATVEG: case VR_OUT_PET_VEGNOCR: case VR_OUT_PREC: case VR_OUT_RAINF:
    case VR_OUT_RUNOFF: case VR_OUT_SNOW_MELT: case VR_OUT_SNOWF: case VR_OUT_SUB_CANOP: case VR_OUT_SUB_SNOW: case VR_OUT_TRANSP_VEG:
      return 1;
    default:
      return 2;
  }
}

int vr_aggregate_device(int device, int32_t n, const int32_t *vars, int32_t cond_for_resist, int32_t nrec, int32_t ratio, int64_t ncell,
                        const double *in_dev, double *out_dev, void *cuda_stream)
{
  if (!vars || !in_dev || !out_dev || n < 1 || n > VR_MAX_OUT || ratio < 1 || nrec < ratio || nrec % ratio || ncell < 1) return VR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  int kind[VR_MAX_OUT];
  for (int v = 0; v < n; v++) kind[v] = (vars[v] == VR_OUT_AERO_RESIST && cond_for_resist) ? 3 : agg_kind(vars[v]);
  int *d_kind = nullptr;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (cudaMalloc((void **)&d_kind, n * sizeof(int)) != cudaSuccess) return VR_ERR_CUDA;
  cudaMemcpyAsync(d_kind, kind, n * sizeof(int), cudaMemcpyHostToDevice, st);
  k_aggregate<<<dim3((unsigned)((ncell + 255) / 256), (unsigned)(nrec / ratio < 65535 ? nrec / ratio : 65535), (unsigned)n), 256, 0, st>>>(n, d_kind, nrec, ratio, ncell,
                                                                                                     in_dev, out_dev);
  const cudaError_t e = cudaGetLastError();
  cudaStreamSynchronize(st);
  cudaFree(d_kind);
  return e == cudaSuccess ? VR_OK : VR_ERR_CUDA;
}

int vr_alma_output_device(int device, int32_t n, const int32_t *vars, int32_t nrec, int64_t ncell, int32_t out_dt_hours, double *data_dev,
                          void *cuda_stream)
{
  if (!vars || !data_dev || n < 1 || n > VR_MAX_OUT || nrec < 1 || ncell < 1 || out_dt_hours < 1) return VR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  int op[VR_MAX_OUT], aux[VR_MAX_OUT], sub_snow = -1, sub_canop = -1;
  for (int v = 0; v < n; v++) {
    aux[v] = -1;
    switch (vars[v]) {
      case VR_OUT_BASEFLOW: case VR_OUT_EVAP: case VR_OUT_EVAP_BARE: case VR_OUT_EVAP_CANOP: case VR_OUT_INFLOW: case VR_OUT_PREC:
      case VR_OUT_RAINF: case VR_OUT_RUNOFF: case VR_OUT_SNOW_MELT: case VR_OUT_SNOWF: case VR_OUT_SUB_CANOP: case VR_OUT_TRANSP_VEG:
        op[v] = 1; break;
      case VR_OUT_SNOW_PACK_TEMP: case VR_OUT_SNOW_SURF_TEMP: case VR_OUT_SURF_TEMP: case VR_OUT_AIR_TEMP: op[v] = 2; break;
      case VR_OUT_DELTAH: op[v] = 3; break;
      case VR_OUT_PRESSURE: case VR_OUT_VP: case VR_OUT_VPD: op[v] = 5; break;
      case VR_OUT_SUB_SNOW: op[v] = 4; sub_snow = v; break;
      default: op[v] = 0;
    }
    if (vars[v] == VR_OUT_SUB_CANOP) sub_canop = v;
  }
  if (sub_snow >= 0 && sub_canop < 0) return VR_ERR_ARG;       // vr_outsel_step_vars(mode & 2) puts it there
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int *d_op = nullptr;
  if (cudaMalloc((void **)&d_op, 2 * n * sizeof(int)) != cudaSuccess) return VR_ERR_CUDA;
  cudaMemcpyAsync(d_op, op, n * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_op + n, aux, n * sizeof(int), cudaMemcpyHostToDevice, st);
  const double sec = (double)(out_dt_hours * 3600);
  const dim3 grid((unsigned)((ncell + 255) / 256), (unsigned)(nrec < 65535 ? nrec : 65535), (unsigned)n);
  k_alma<<<grid, 256, 0, st>>>(n, d_op, d_op + n, nrec, ncell, sec, data_dev);
  if (sub_snow >= 0)
    k_alma_add<<<dim3(grid.x, grid.y), 256, 0, st>>>(sub_snow, sub_canop, nrec, ncell, sec, data_dev);
  const cudaError_t e = cudaGetLastError();
  cudaStreamSynchronize(st);
  cudaFree(d_op);
  return e == cudaSuccess ? VR_OK : VR_ERR_CUDA;
}

int vr_forcing_outputs_device(int device, int32_t n, const int32_t *vars, int32_t nrecs, int32_t nf, int64_t ncell, double max_snow_temp,
                              double min_rain_temp, const double *const fields_dev[9], double *out_dev, void *cuda_stream)
{
  if (!vars || !fields_dev || !out_dev || n < 1 || n > VR_MAX_OUT || nrecs < 1 || nf < 1 || ncell < 1) return VR_ERR_ARG;
  for (int v = 0; v < n; v++)
    switch (vars[v]) {
      case VR_OUT_AIR_TEMP: case VR_OUT_DENSITY: case VR_OUT_LONGWAVE: case VR_OUT_PREC: case VR_OUT_PRESSURE: case VR_OUT_QAIR:
      case VR_OUT_REL_HUMID: case VR_OUT_SHORTWAVE: case VR_OUT_VP: case VR_OUT_VPD: case VR_OUT_WIND: case VR_OUT_RAINF:
      case VR_OUT_SNOWF: break;
      default: return VR_ERR_UNSUPPORTED;       // not a forcing variable (or OUT_CATM / COSZEN / FDIR / PAR / TSKC: not kept)
    }
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  ForceFields F;
  for (int i = 0; i < VR_NFORCING; i++) { if (!fields_dev[i]) return VR_ERR_ARG; F.f[i] = fields_dev[i]; }
  int *d_vars = nullptr;
  if (cudaMalloc((void **)&d_vars, n * sizeof(int)) != cudaSuccess) return VR_ERR_CUDA;
  cudaMemcpyAsync(d_vars, vars, n * sizeof(int), cudaMemcpyHostToDevice, st);
  const int nrows = nrecs * nf;
  k_forcing_outputs<<<dim3((unsigned)((ncell + 255) / 256), (unsigned)(nrows < 65535 ? nrows : 65535), (unsigned)n), 256, 0, st>>>(n, d_vars, nrows, nf, ncell, max_snow_temp,
                                                                                                 min_rain_temp, F, out_dev);
  const cudaError_t e = cudaGetLastError();
  cudaStreamSynchronize(st);
  cudaFree(d_vars);
  return e == cudaSuccess ? VR_OK : VR_ERR_CUDA;
}

int vr_forcing_outputs(int device, int32_t n, const int32_t *vars, int32_t nrecs, int32_t nf, int64_t ncell, double max_snow_temp,
                       double min_rain_temp, const double *const fields[9], double *out)
{
  if (!fields || !out || n < 1 || nrecs < 1 || nf < 1 || ncell < 1) return VR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  const size_t nin = (size_t)nrecs * (nf > 1 ? nf + 1 : 1) * ncell, nout = (size_t)n * nrecs * nf * ncell;
  double *d = nullptr;
  if (cudaMalloc((void **)&d, (VR_NFORCING * nin + nout) * sizeof(double)) != cudaSuccess) return VR_ERR_CUDA;
  const double *fd[VR_NFORCING];
  for (int i = 0; i < VR_NFORCING; i++) {
    if (!fields[i]) { cudaFree(d); return VR_ERR_ARG; }
    cudaMemcpy(d + (size_t)i * nin, fields[i], nin * sizeof(double), cudaMemcpyHostToDevice);
    fd[i] = d + (size_t)i * nin;
  }
  double *dout = d + (size_t)VR_NFORCING * nin;
  const int rc = vr_forcing_outputs_device(device, n, vars, nrecs, nf, ncell, max_snow_temp, min_rain_temp, fd, dout, nullptr);
  if (rc == VR_OK) cudaMemcpy(out, dout, nout * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return rc;
}

int vr_alma_output(int device, int32_t n, const int32_t *vars, int32_t nrec, int64_t ncell, int32_t out_dt_hours, double *data)
{
  if (!data || n < 1 || nrec < 1 || ncell < 1) return VR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  double *d = nullptr;
  const size_t bytes = (size_t)n * nrec * ncell * sizeof(double);
  if (cudaMalloc((void **)&d, bytes) != cudaSuccess) return VR_ERR_CUDA;
  cudaMemcpy(d, data, bytes, cudaMemcpyHostToDevice);
  const int rc = vr_alma_output_device(device, n, vars, nrec, ncell, out_dt_hours, d, nullptr);
  if (rc == VR_OK) cudaMemcpy(data, d, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return rc;
}

int vr_aggregate(int device, int32_t n, const int32_t *vars, int32_t cond_for_resist, int32_t nrec, int32_t ratio, int64_t ncell,
                 const double *in, double *out)
{
  if (!in || !out || n < 1 || ratio < 1 || nrec < ratio || nrec % ratio || ncell < 1) return VR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  double *d_in = nullptr, *d_out = nullptr;
  const size_t ni = (size_t)n * nrec * ncell, no = (size_t)n * (nrec / ratio) * ncell;
  if (cudaMalloc((void **)&d_in, ni * sizeof(double)) != cudaSuccess) return VR_ERR_CUDA;
  if (cudaMalloc((void **)&d_out, no * sizeof(double)) != cudaSuccess) { cudaFree(d_in); return VR_ERR_CUDA; }
  cudaMemcpy(d_in, in, ni * sizeof(double), cudaMemcpyHostToDevice);
  const int rc = vr_aggregate_device(device, n, vars, cond_for_resist, nrec, ratio, ncell, d_in, d_out, nullptr);
  if (rc == VR_OK) cudaMemcpy(out, d_out, no * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_out);
  return rc;
}

int vr_diag_fp64_peak(int device, double *dfma_per_s)
{
  if (!dfma_per_s) return VR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double *sink = nullptr;
  if (cudaMalloc((void **)&sink, 8) != cudaSuccess) return VR_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * 8, iters = 200000;
  k_diag_dfma<<<blocks, 256>>>(1000, 1.0, sink);               // warm-up
  cudaEventRecord(e0);
  k_diag_dfma<<<blocks, 256>>>(iters, 1.0, sink);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  if (e != cudaSuccess || ms <= 0) return VR_ERR_CUDA;
  *dfma_per_s = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
  return VR_OK;
}

int vr_diag_pow(int device, int64_t n, const double *x, const double *y, double *z)
{
  if (!x || !y || !z || n < 1) return VR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VR_ERR_CUDA;
  double *d = nullptr;
  if (cudaMalloc((void **)&d, 3 * n * sizeof(double)) != cudaSuccess) return VR_ERR_CUDA;
  cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(d + n, y, n * sizeof(double), cudaMemcpyHostToDevice);
  k_diag_pow<<<(unsigned)((n + 255) / 256), 256>>>(n, d, d + n, d + 2 * n);
  cudaError_t e = cudaMemcpy(z, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? VR_OK : VR_ERR_CUDA;
}

int vr_fallback_counts(vr_handle *h, int64_t counts[2])
{
  if (!h || !counts) return VR_ERR_ARG;
  if (!h->initialised) { h->err = "no state yet"; return VR_ERR_STATE; }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->s_cells));
  const int64_t U = h->M.U;
  std::vector<double> a(U), b(U);
  CK(cudaMemcpy(a.data(), h->d_state.p + (size_t)ST_FB_SNOW * U, U * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), h->d_state.p + (size_t)ST_FB_FOL * U, U * sizeof(double), cudaMemcpyDeviceToHost));
  counts[0] = counts[1] = 0;
  for (int64_t u = 0; u < U; u++) { counts[0] += (int64_t)a[u]; counts[1] += (int64_t)b[u]; }
  return VR_OK;
}

}  // extern "C"
