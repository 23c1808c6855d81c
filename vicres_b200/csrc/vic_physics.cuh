// vic_physics.cuh -- device physics of the B200-native VIC-Res water-balance cell step.
//
// One call of unit_step() advances ONE tile x band pair ("unit") of one grid cell through ONE
// model record: the work the reference does in surface_fluxes() + runoff() for that pair
// (runoff/model/surface_fluxes.c:11-1168, runoff.c:7-576) in frozen-soil-free water-balance mode
// with SNOW_STEP == TIME_STEP.  It is a re-design, not a transcription:
//   * state is a flat register struct (26 live words), not the reference's ~2 kB AoS structs;
//     the reference's iter_/step_/snow_/soil_ struct copies collapse to scalars;
//   * everything that depends only on parameters is hoisted to the host at vr_set_cells():
//     aerodynamic-resistance coefficients per (class, soil roughness, month) -- the reference runs
//     7 CalcAerodynamic per tile-step (full_energy.c:329-369) --, soil thermal constants, the
//     QUICK_FLUX exponentials, ARNO exponents, baseflow coefficients, monthly canopy terms;
//   * the Penman-Monteith air-state terms (2 exp) are computed once per unit-step and shared by
//     the up-to-5 penman() evaluations of canopy_evap/transpiration/arno_evap;
//   * the 6 PET reference evaporations (compute_pot_evap.c) and compute_zwt are compile-time options of
//     unit_step (template argument `extras`): they feed only OUT_PET_* / OUT_ZWT*.
// Operation order inside each formula follows the reference so results stay within 1e-9 relative.
//
// The functions are __host__ __device__ so that tests/hostsim can compile this header with g++ and
// check the arithmetic against the oracle without a GPU; the product library (vicres_capi.cu)
// only ever launches them from CUDA kernels.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VR_HD __host__ __device__ __forceinline__
#define VR_HDN __host__ __device__ __noinline__
#else
#define VR_HD inline
#define VR_HDN inline
#endif

// One out-of-line copy of each transcendental: CUDA inlines pow/exp/log at every call site (60-250
// SASS instructions each, ~35 sites), which made the step kernel 30k instructions / 480 KB and stalled the
// warps on instruction fetch (profiles/r1_k_step_ncu.md, v1).  Out of line the whole step fits the
// instruction cache.
#ifndef VR_POW
#define VR_POW(x, y) ::vr::m_pow((x), (y))
#endif
// Phase barrier: keeps the warps of a thread block at the same place in the (large) step code so
// that one instruction-cache fill serves all of them (the step kernel is instruction-supply bound, see
// profiles/).  Every thread of the block must run unit_step() for this to be legal.
#ifndef VR_BAR_MASK
#define VR_BAR_MASK 15      // which of the four phase barriers of unit_step are compiled in (tuning knob)
#endif
#if defined(__CUDA_ARCH__) && defined(VR_PHASE_BARRIERS)
#define VR_PHASE_SYNC(i, ok) do { if (((VR_BAR_MASK >> (i)) & 1) && (ok)) __syncthreads(); } while (0)
#else
#define VR_PHASE_SYNC(i, ok)
#endif
#if defined(__CUDA_ARCH__) && defined(VR_LOOP_BARRIERS)
#define VR_LOOP_SYNC() __syncthreads()
#else
#define VR_LOOP_SYNC()
#endif
#define VR_EXP(x) ::vr::m_exp(x)
#define VR_LOG(x) ::vr::m_log(x)
#define VR_LOG10(x) ::vr::m_log10(x)

namespace vr {

// x^y for the step's power laws (Brooks-Corey drainage, ARNO curves, albedo decay ...): 72 of the
// ~85 pow() per tile-day sit in the hourly drainage loop (runoff.c:320-479) and CUDA's pow() costs
// ~250 instructions (double-double log), which made it 80 % of the first step kernel.  On the device
// x^y = 2^(y*log2 x) is evaluated directly with two small tables in shared memory (tools/make_pow_tables.py):
//   log2 x = e + l2c[i] + log2(1 + r),  i = round((m - 1) * 128),  r = m * inv_c[i] - 1 (one exact fma, |r| <= 2^-8),
//            log2(1 + r) = r * (degree-6 Taylor polynomial); entries with c >= sqrt(2) belong to the binade above so
//            that x -> 1 from either side has no cancellation; the exponent part y*e is kept separate (added exactly);
//   2^t    = 2^n * 2^(j/64) * 2^g,  t = n + j/64 + g,  |g| <= 1/128,  2^g - 1 = g * (degree-4 Taylor polynomial).
// Relative error <= ~2e-16 * (2 + |y*log2 x|), i.e. < 1e-13 for every argument range of the model -- four orders
// below the 1e-9 parity bar (tests: test_device_pow_error_bound on the GPU, test_pow_table_error_bound on the host).
// About 60 SASS instructions (the atanh-series version of round 1: 179).  Zero, negative, subnormal, non-finite or
// extreme arguments take CUDA's pow().  With -DVR_EXACT_LIBM (and always in host builds of the step) plain pow() is
// used, which is what makes the host build bit-identical to the reference.
#include "vic_pow_tables.h"
struct PowTables { double logtab[2 * 129]; double exptab[64]; };
VR_HD double pow_table_eval(double x, double y, const double *logtab, const double *exptab, bool *handled)
{
  constexpr double L[7] = VR_POW_L_INIT;
  constexpr double E[5] = VR_POW_E_INIT;
  *handled = false;
  if (!((x >= 2.2250738585072014e-308) && (x <= 1e300) && (fabs(y) <= 1e5))) return 0.;
#if defined(__CUDA_ARCH__)
  const int hi = __double2hiint(x);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
#else
  long long ix;
  memcpy(&ix, &x, 8);
  const int hi = (int)(ix >> 32);
  const long long im = (ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
  double m;
  memcpy(&m, &im, 8);
#endif
  const int i = ((hi & 0x000fffff) + 0x1000) >> 13;                    // 0..128
  const int e = (hi >> 20) - 1023 + (i >= VR_POW_SPLIT ? 1 : 0);
  const double r = fma(m, logtab[2 * i], -1.0);
  const double r2 = r * r;
  const double p01 = fma(L[1], r, L[0]), p23 = fma(L[3], r, L[2]), p45 = fma(L[5], r, L[4]);
  const double q = fma(p23, r2, p01), u = fma(L[6], r2, p45);
  const double pl = fma(u, r2 * r2, q);
  const double l2m = fma(r, pl, logtab[2 * i + 1]);                     // log2(m) relative to the entry's binade
  const double t = fma(y, l2m, y * (double)e);                         // y * log2(x)
  if (t < -1100.0) { *handled = true; return 0.; }                     // x^y underflows to +0 (ARNO curves with 1/b up to 1000)
  if (!(fabs(t) < 1000.0)) return 0.;
  const double kd0 = fma(t, 64.0, 6755399441055744.0);                 // 1.5 * 2^52: the integer lands in the low word
#if defined(__CUDA_ARCH__)
  const int k = __double2loint(kd0);
#else
  long long ik;
  memcpy(&ik, &kd0, 8);
  const int k = (int)(ik & 0xffffffffLL);
#endif
  const double kd = kd0 - 6755399441055744.0;
  const double g = fma(kd, -0.015625, t);                              // exact
  double pe = fma(E[4], g, E[3]);
  pe = fma(pe, g, E[2]); pe = fma(pe, g, E[1]); pe = fma(pe, g, E[0]);
  const double T = exptab[k & 63];
  const double res = fma(T * g, pe, T);
  const int n = k >> 6;
  *handled = true;
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(res) + (n << 20), __double2loint(res));
#else
  long long ir;
  memcpy(&ir, &res, 8);
  ir += (long long)n << 52;
  double out;
  memcpy(&out, &ir, 8);
  return out;
#endif
}

#if defined(__CUDACC__)
static __device__ __align__(16) const double VR_POW_LOGTAB_G[2 * 129] = VR_POW_LOGTAB_INIT;
static __device__ __align__(16) const double VR_POW_EXPTAB_G[64] = VR_POW_EXPTAB_INIT;
__shared__ __align__(16) PowTables vr_pow_tables;
// every kernel that reaches m_pow() calls this first (all threads of the block)
__device__ __forceinline__ void pow_tables_init()
{
#if !defined(VR_EXACT_LIBM)
  for (int i = threadIdx.x; i < 2 * 129; i += blockDim.x) vr_pow_tables.logtab[i] = VR_POW_LOGTAB_G[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) vr_pow_tables.exptab[i] = VR_POW_EXPTAB_G[i];
  __syncthreads();
#endif
}
#endif
VR_HDN double m_pow(double x, double y)
{
#if defined(__CUDA_ARCH__) && !defined(VR_EXACT_LIBM)
  if (x == 1.0 || y == 0.0) return 1.0;
  if (x == 0.0 && y > 0.0) return 0.0;            // dry layer: Q12 = Ksat * 0^expt
  bool handled;
  const double v = pow_table_eval(x, y, vr_pow_tables.logtab, vr_pow_tables.exptab, &handled);
  if (handled) return v;
  return pow(x, y);
#else
  return pow(x, y);
#endif
}
// the same, inlined: for the two Brooks-Corey powers of the hourly drainage loop, where the call (argument moves and the
// caller's live constants kept out of the callee's registers) costs more than the second copy of the code
VR_HD double m_pow_inl(double x, double y)
{
#if defined(__CUDA_ARCH__) && !defined(VR_EXACT_LIBM) && !defined(VR_POW_CALL_IN_LOOP)
  if (x == 1.0 || y == 0.0) return 1.0;
  if (x == 0.0 && y > 0.0) return 0.0;
  bool handled;
  const double v = pow_table_eval(x, y, vr_pow_tables.logtab, vr_pow_tables.exptab, &handled);
  if (handled) return v;
  return m_pow(x, y);
#else
  return m_pow(x, y);
#endif
}
#if defined(__CUDACC__)
// polynomial coefficients in constant memory: the FP64 pipe takes them as c[bank][offset] operands, so they cost neither
// registers nor the two moves per coefficient that an immediate needs (10 % of the drainage kernel's instructions were UMOV)
#if !defined(VR_POW_IMMEDIATES)
__constant__ double VR_POW_LC[7] = VR_POW_L_INIT;
__constant__ double VR_POW_EC[5] = VR_POW_E_INIT;
#endif
// x^y for the hourly Brooks-Corey drainage (runoff.c:336): x = relative saturation, 0 <= x <= ~1 and finite, y = the
// layer's exponent, y >= 1 (the caller checks).  Same arithmetic as pow_table_eval() -- identical results wherever that
// one handles the argument -- without its range checks and special cases: x == 1 gives exactly 1 through the table
// (r == 0, t == 0), x == 0 or subnormal has the biased exponent 0, hence t <= -1022 y, and everything with t <= -1021
// returns 0 (x^y < 2^-1021: the result pow() gives there is below 2.3e-308).  No branch.  tb = shared-memory address of
// vr_pow_tables.
__device__ __forceinline__ double drain_pow(double x, double y, uint32_t tb)
{
#if defined(VR_POW_IMMEDIATES)       // A/B: coefficients as immediates (two moves each) instead of constant-bank loads
  constexpr double VR_POW_LC[7] = VR_POW_L_INIT;
  constexpr double VR_POW_EC[5] = VR_POW_E_INIT;
#endif
  const int hi = __double2hiint(x);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const int i = ((hi & 0x000fffff) + 0x1000) >> 13;                    // 0..128
  const int e = (hi >> 20) - 1023 + (i >= VR_POW_SPLIT ? 1 : 0);
  double inv_c, l2c;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(inv_c), "=d"(l2c) : "r"(tb + 16u * (uint32_t)i));
  const double r = fma(m, inv_c, -1.0);
  const double r2 = r * r;
  const double p01 = fma(VR_POW_LC[1], r, VR_POW_LC[0]), p23 = fma(VR_POW_LC[3], r, VR_POW_LC[2]), p45 = fma(VR_POW_LC[5], r, VR_POW_LC[4]);
  const double q = fma(p23, r2, p01), u = fma(VR_POW_LC[6], r2, p45);
  const double pl = fma(u, r2 * r2, q);
  const double l2m = fma(r, pl, l2c);
  const double t = fma(y, l2m, y * (double)e);                         // y * log2(x)
  const double kd0 = fma(t, 64.0, 6755399441055744.0);
  const int k = __double2loint(kd0);
  const double kd = kd0 - 6755399441055744.0;
  const double g = fma(kd, -0.015625, t);
  double pe = fma(VR_POW_EC[4], g, VR_POW_EC[3]);
  pe = fma(pe, g, VR_POW_EC[2]); pe = fma(pe, g, VR_POW_EC[1]); pe = fma(pe, g, VR_POW_EC[0]);
  double T;
  asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tb + (uint32_t)(2 * 129 * 8) + 8u * (uint32_t)(k & 63)));
  const double res = fma(T * g, pe, T);
  const double out = __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
  return (t > -1021.0) ? out : 0.0;
}
#endif
VR_HDN double m_exp(double x) { return exp(x); }
VR_HDN double m_log(double x) { return log(x); }
VR_HDN double m_log10(double x) { return log10(x); }

// ---- constants (runoff/model/vicNl_def.h:160-440, snow.h:40-90) ----
constexpr double HUGE_RESIST = 1.e20;
constexpr double SMALLV = 1.e-12;
constexpr double MISSINGV = -99999.;
constexpr double VERR = -999.;
constexpr double BARE_SOIL_ALBEDO = 0.2;
constexpr double SEC_PER_DAY = 86400.;
constexpr double ICE_DENSITY = 917.;
constexpr double KELVIN = 273.15;
constexpr double STEFAN_B = 5.6696e-8;
constexpr double Lf = 3.337e5;
constexpr double RHO_W = 999.842594;
constexpr double Cp = 1013.0;
constexpr double CH_ICE = 2100.0e3;
constexpr double CH_WATER = 4186.8e3;
constexpr double K_SNOW = 2.9302e-6;
constexpr double EPS = 0.62196351;
constexpr double G_GRAV = 9.81;
constexpr double JOULESPCAL = 4.1868;
constexpr double GRAMSPKG = 1000.;
constexpr double A_SVP = 0.61078, B_SVP = 17.269, C_SVP = 237.3;
constexpr double CLOSURE = 4000, RSMAX = 5000, VPDMINFACTOR = 0.1, RARC_SOIL = 100.0;
constexpr double CP_PM = 1013, PS_PM = 101300;
constexpr double SNOW_DT = 5.0;
constexpr double LIQUID_WATER_CAPACITY = 0.035;
constexpr double LAI_SNOW_MULTIPLIER = 0.0005;
constexpr double MIN_INTERCEPTION_STORAGE = 0.005;
constexpr double MAX_SURFACE_SWE = 0.125;
constexpr double NEW_SNOW_DENSITY = 50.;
constexpr double SNDENS_ETA0 = 3.6e6, SNDENS_C5 = 0.08, SNDENS_C6 = 0.021;
constexpr double MIN_SWQ_EB_THRES = 0.0010;
constexpr double NEW_SNOW_ALB = 0.85;
constexpr double SNOW_ALB_ACCUM_A = 0.94, SNOW_ALB_ACCUM_B = 0.58, SNOW_ALB_THAW_A = 0.82, SNOW_ALB_THAW_B = 0.46;
constexpr double TraceSnow = 0.03;

// ---- per-cell constants, precomputed on the host (vic_host_prep.h) ----
// Rows of the constant block cc[k*C + c].  The rows the kernels before runoff() read (k_front, k_front_snow) come first:
// those kernels stage exactly that prefix into shared memory (bulk copies), k_back reads its rows from global memory.
// Per-layer fields are laid out for MAXL layers whatever NLAYER is.
constexpr int MAXL = 3;
enum {
  // -- front block --
  CC_LAT = 0, CC_ELEV, CC_HALF_ELEV_LAPSE, CC_B, CC_INV_B, CC_INV_B1, CC_ARNO_MAXM, CC_ARNO_MAX_INFIL, CC_AVG_T, CC_DP,
  CC_D1, CC_EXP_M, CC_EXP_P, CC_KB_BARE,
  CC_FL0,      // then per layer l: CC_FL0 + l*CLF_N + field
  CLF_WCR = 0, CLF_WPWP, CLF_WCR_M_WPWP, CLF_RESID, CLF_DEPTH, CLF_KDRY, CLF_KSAT_M_KDRY, CLF_POROS, CLF_SOILFR,
  CLF_CS_SOLID, CLF_N,
  CC_FRONT_N = CC_FL0 + MAXL * CLF_N,
  // -- back block --
  CC_ONE_PLUS_B = CC_FRONT_N, CC_EX, CC_TOP_MAX, CC_TOP_MAX_INFIL, CC_DSMAX24, CC_BF_FRAC, CC_BF_NL, CC_WS, CC_ONE_M_WS,
  CC_C, CC_INV_ONE_M_WS,
  CC_BL0,      // then per layer l: CC_BL0 + l*CLB_N + field
  CLB_EXPT = 0, CLB_KSAT24, CLB_MAX_MOIST, CLB_QDEN, CLB_INV_QDEN, CLB_WET_DEN, CLB_N,
  CC_N = CC_BL0 + MAXL * CLB_N
};
#define VR_CC_N(nlayer) (::vr::CC_N)
// optional tail of the block: the zwtvmoist curves, (nlayer+2) curves x ZWT_NP points, zwt then moist
constexpr int ZWT_NP = 11;      // MAX_ZWTVMOIST
#define VR_CC_ZWT_N(nlayer) (2 * ((nlayer) + 2) * ::vr::ZWT_NP)

struct CellC {   // read-only view of one cell's constants: row k at p[k*C] (global block, or the shared-memory stage)
  const double *p;
  int64_t C;
  VR_HD double g(int k) const { return p[(size_t)k * C]; }
  VR_HD double lat() const { return g(CC_LAT); }
  VR_HD double elev() const { return g(CC_ELEV); }
  VR_HD double half_elev_lapse() const { return g(CC_HALF_ELEV_LAPSE); }
  VR_HD double b() const { return g(CC_B); }
  VR_HD double inv_b() const { return g(CC_INV_B); }
  VR_HD double inv_b1() const { return g(CC_INV_B1); }
  VR_HD double one_plus_b() const { return g(CC_ONE_PLUS_B); }
  VR_HD double ex() const { return g(CC_EX); }
  VR_HD double top_max() const { return g(CC_TOP_MAX); }
  VR_HD double top_max_infil() const { return g(CC_TOP_MAX_INFIL); }
  VR_HD double arno_maxm() const { return g(CC_ARNO_MAXM); }
  VR_HD double arno_max_infil() const { return g(CC_ARNO_MAX_INFIL); }
  VR_HD double dsmax24() const { return g(CC_DSMAX24); }
  VR_HD double bf_frac() const { return g(CC_BF_FRAC); }
  VR_HD double bf_nl() const { return g(CC_BF_NL); }
  VR_HD double Ws() const { return g(CC_WS); }
  VR_HD double one_m_Ws() const { return g(CC_ONE_M_WS); }
  VR_HD double c() const { return g(CC_C); }
  VR_HD double avg_T() const { return g(CC_AVG_T); }
  VR_HD double dp() const { return g(CC_DP); }
  VR_HD double D1() const { return g(CC_D1); }
  VR_HD double exp_m() const { return g(CC_EXP_M); }
  VR_HD double exp_p() const { return g(CC_EXP_P); }
  VR_HD double kb_bare() const { return g(CC_KB_BARE); }
  VR_HD double inv_one_m_Ws() const { return g(CC_INV_ONE_M_WS); }
  VR_HD double expt(int l) const { return g(CC_BL0 + l * CLB_N + CLB_EXPT); }
  VR_HD double ksat24(int l) const { return g(CC_BL0 + l * CLB_N + CLB_KSAT24); }
  VR_HD double max_moist(int l) const { return g(CC_BL0 + l * CLB_N + CLB_MAX_MOIST); }
  VR_HD double qden(int l) const { return g(CC_BL0 + l * CLB_N + CLB_QDEN); }
  VR_HD double inv_qden(int l) const { return g(CC_BL0 + l * CLB_N + CLB_INV_QDEN); }
  VR_HD double wet_den(int l) const { return g(CC_BL0 + l * CLB_N + CLB_WET_DEN); }
  VR_HD double resid(int l) const { return g(CC_FL0 + l * CLF_N + CLF_RESID); }
  VR_HD double Wcr(int l) const { return g(CC_FL0 + l * CLF_N + CLF_WCR); }
  VR_HD double Wpwp(int l) const { return g(CC_FL0 + l * CLF_N + CLF_WPWP); }
  VR_HD double wcr_m_wpwp(int l) const { return g(CC_FL0 + l * CLF_N + CLF_WCR_M_WPWP); }
  VR_HD double depth(int l) const { return g(CC_FL0 + l * CLF_N + CLF_DEPTH); }
  VR_HD double kdry(int l) const { return g(CC_FL0 + l * CLF_N + CLF_KDRY); }
  VR_HD double ksat_m_kdry(int l) const { return g(CC_FL0 + l * CLF_N + CLF_KSAT_M_KDRY); }
  VR_HD double poros(int l) const { return g(CC_FL0 + l * CLF_N + CLF_POROS); }
  VR_HD double soilfr(int l) const { return g(CC_FL0 + l * CLF_N + CLF_SOILFR); }
  VR_HD double cs_solid(int l) const { return g(CC_FL0 + l * CLF_N + CLF_CS_SOLID); }
  // zwtvmoist curve k, point i (only valid when the block carries the optional tail)
  VR_HD double zwt_z(int NL, int k, int i) const { return g(VR_CC_N(NL) + k * ZWT_NP + i); }
  VR_HD double zwt_m(int NL, int k, int i) const { return g(VR_CC_N(NL) + (NL + 2 + k) * ZWT_NP + i); }
};

// per-unit constants (tile x band)
struct UnitC {
  int is_bare, overstory, root0_mask;
  double area, Tfactor, Pfactor, rarc, rmin, RGL;
  float root[MAXL];
};

// per (tile, month) canopy terms (full_energy.c:210-227,306-308) -- tm[(t*12+mon)*TM_N + k]
enum { TM_VEGCOVER = 0, TM_ALBEDO, TM_LAI, TM_WDMAX, TM_SURF_ATTEN, TM_N };
// separate per (tile, month) table read only by the potential evaporation: the LIBRARY's LAI / albedo of the
// tile's class (the SATSOIL reference entry for the bare tile), compute_pot_evap.c:62-66
enum { TL_LAI = 0, TL_ALB, TL_N };
// per (aero key, month) CalcAerodynamic coefficients: Ra = RA/wind, U = UC*wind
enum { AE_RA0 = 0, AE_RA1, AE_RA2, AE_U0, AE_U1, AE_U2, AE_REF0, AE_REF1, AE_REF2, AE_ROUGH0, AE_ROUGH1,
       AE_ROUGH2, AE_DISP0, AE_DISP1, AE_DISP2, AE_N };
// separate per (aero key, month) table read only by the potential evaporation: entry 3*p + i = RA coefficient i
// of PET reference cover p (SATSOIL, H2OSURF, SHORT, TALL; full_energy.c:329-369)
enum { AP_N = 12 };

// atmos fields of one record or sub-step; snowflag: atmos->snowflag[NR] of the record (initialize_atmos.c:1390-1400),
// read only when the snow model is sub-stepped (SNOW_STEP < TIME_STEP)
struct Forc { double air_temp, prec, wind, vp, vpd, pressure, density, shortwave, longwave; int mon, doy, snowflag; };

// carried state of one unit (what write_model_state.c keeps, plus LongUnderOut/Tfoliage)
struct UnitS {
  double moist[MAXL], Wdew, T0, T1, LongUnderOut, Tfoliage;
  double swq, surf_temp, pack_temp, surf_water, pack_water, density, depth, albedo, coldcontent, coverage,
         snow_canopy, tmp_int_storage;
  int last_snow, MELTING, fb_snow, fb_fol;
};
enum { ST_MOIST0 = 0, ST_MOIST1, ST_MOIST2, ST_WDEW, ST_T0, ST_T1, ST_LONGUNDEROUT, ST_TFOLIAGE, ST_SWQ,
       ST_SURF_TEMP, ST_PACK_TEMP, ST_SURF_WATER, ST_PACK_WATER, ST_DENSITY, ST_DEPTH, ST_ALBEDO,
       ST_COLDCONTENT, ST_COVERAGE, ST_SNOW_CANOPY, ST_TMP_INT, ST_LAST_SNOW, ST_MELTING, ST_FB_SNOW,
       ST_FB_FOL, ST_N };

// what one unit hands to the per-cell aggregation (collect_wb_terms / collect_eb_terms inputs)
struct UnitOut {
  double evap[MAXL], bef[MAXL], vapor_flux, canopy_vapor_flux, canopyevap, asat, runoff, baseflow, inflow,
         aero_resist[2], wetness, rootmoist, melt, Tsurf, NetShortAtmos, NetLongAtmos, in_long, albedo,
         grnd_flux, deltaH, out_prec, out_rain, out_snow;
  int snow_flag;
};

// optional results (unit_step's `extras`); a separate struct so that the default kernel instance carries none of it
struct UnitExtra { double pot_evap[6], zwt, zwt_lumped; };

// view of the constants of the cell at sorted position pos
VR_HD void load_cell_consts(const double *cc, int64_t C, int64_t pos, int NL, CellC &x)
{
  (void)NL;
  x.p = cc + pos;
  x.C = C;
}

// initialize_model_state.c:321-350,733-760 + initialize_snow/soil/veg: the state before record 0
VR_HD void unit_init(const double *init_moist, const double *max_moist, int NL, double tair0, double Tfactor, UnitS &s)
{
  double surf_temp = tair0;
  if (surf_temp < -1.) surf_temp = -1.;
#pragma unroll
  for (int l = 0; l < MAXL; l++) {
    s.moist[l] = 0;
    if (l < NL) {
      s.moist[l] = init_moist[l];
      if (s.moist[l] > max_moist[l]) s.moist[l] = max_moist[l];
    }
  }
  s.Wdew = 0; s.T0 = surf_temp; s.T1 = surf_temp;
  double tmp = s.T0 + KELVIN;
  s.LongUnderOut = STEFAN_B * tmp * tmp * tmp * tmp;
  s.Tfoliage = tair0 + Tfactor;
  s.swq = 0; s.surf_temp = 0; s.pack_temp = 0; s.surf_water = 0; s.pack_water = 0; s.density = 0; s.depth = 0;
  s.albedo = 0; s.coldcontent = 0; s.coverage = 0; s.snow_canopy = 0; s.tmp_int_storage = 0;
  s.last_snow = (int)MISSINGV; s.MELTING = 0; s.fb_snow = 0; s.fb_fol = 0;
}

// ---------------------------------------------------------------------------------------
// L0 math
// ---------------------------------------------------------------------------------------
VR_HDN double svp(double temp)   // svp.c:7-24
{
  double SVP = A_SVP * VR_EXP((B_SVP * temp) / (C_SVP + temp));
  if (temp < 0) SVP *= 1.0 + .00972 * temp + .000042 * temp * temp;
  return SVP * 1000.;
}

// Penman-Monteith (penman.c:201-249) split into its air-state part (once per unit-step) and the
// per-surface evaluation.
struct PenmanAir { double slope, lv, gamma, r_air_cp; };
VR_HDN PenmanAir penman_air(double tair, double half_elev_lapse, double elevation)
{
  PenmanAir p;
  p.slope = (B_SVP * C_SVP) / ((C_SVP + tair) * (C_SVP + tair)) * svp(tair);
  double h = 287 / 9.81 * ((tair + 273.15) + half_elev_lapse);
  double pz = PS_PM * VR_EXP(-elevation / h);
  p.lv = 2501000 - 2361 * tair;
  p.gamma = 1628.6 * pz / p.lv;
  double r_air = 0.003486 * pz / (275 + tair);
  p.r_air_cp = r_air * CP_PM;
  return p;
}
VR_HDN double penman(PenmanAir p, double rad, double vpd, double ra, double rc, double rarc)
{
  double evap = (p.slope * rad + p.r_air_cp * vpd / ra) / (p.lv * (p.slope + p.gamma * (1 + (rc + rarc) / ra))) *
                SEC_PER_DAY;
  if (vpd >= 0.0 && evap < 0.0) evap = 0.0;
  return evap;
}

VR_HDN double calc_rc(double rs, double net_short, double RGL, double tair, double vpd, double lai, double gsm_inv)
{   // penman.c:44-91, Jarvis branch
  double rc;
  if (rs == 0) rc = 0;
  else if (lai == 0) rc = RSMAX;
  else {
    double DAYfactor;
    if (rs > 0.) {
      double f = net_short / RGL;
      DAYfactor = (1. + f) / (f + rs / RSMAX);
    }
    else DAYfactor = 1.;
    double Tfactor = .08 * tair - 0.0016 * tair * tair;
    Tfactor = (Tfactor <= 0.0) ? 1e-10 : Tfactor;
    double vpdfactor = 1 - vpd / CLOSURE;
    vpdfactor = (vpdfactor < VPDMINFACTOR) ? VPDMINFACTOR : vpdfactor;
    rc = rs / (lai * gsm_inv * Tfactor * vpdfactor) * DAYfactor;
    rc = (rc > RSMAX) ? RSMAX : rc;
  }
  return rc;
}

// StabilityCorrection.c:44-81.  lnz5 = log((Z - d) / Z0) + 5, which does not depend on the temperatures: the caller
// evaluates it once per energy-balance solve instead of once per Brent evaluation (same value, same result).
VR_HD double stability_lnz5(double Z, double d, double Z0) { return VR_LOG((Z - d) / Z0) + 5; }
VR_HDN double stability_correction(double Zmd, double lnz5, double TSurf, double Tair, double Wind)
{
  double Correction = 1.0;
  if (TSurf != Tair) {
    double Ri = G_GRAV * (Tair - TSurf) * (Zmd) / (((Tair + 273.15) + (TSurf + 273.15)) / 2.0 * Wind * Wind);
    double RiLimit = (Tair + 273.15) / (((Tair + 273.15) + (TSurf + 273.15)) / 2.0 * (lnz5));
    if (Ri > RiLimit) Ri = RiLimit;
    if (Ri > 0.0) Correction = (1 - Ri / 0.2) * (1 - Ri / 0.2);
    else {
      if (Ri < -0.5) Ri = -0.5;
      Correction = sqrt(1 - 16 * Ri);
    }
  }
  return Correction;
}

// small by-value aggregates: the CUDA ABI passes and returns these in registers, so the out-of-line leaves below have
// no object in local memory (round 1 passed UnitS / VegV / PenmanAir / layer arrays by reference: a 2.4 kB frame whose
// misses in L1 were the largest stall class of the step kernel, profiles/r2_k_units_ncu.md)
struct M3 { double v[MAXL]; };

// root_brent.c:105-367 as a resumable state machine, so that a caller has ONE evaluation site of its energy balance
// (the reference calls the function at 5 places inside root_brent and at 2 more around it); bracket expansion and
// Brent iteration are kept evaluation for evaluation.  start(a, b), then feed(f(x)) until it returns true: result is the
// root, or VERR when there is no bracket / no convergence.
struct Brent {
  double a, b, c, d, e, fa, fb, fc, x, result;
  int stage, j, i;
  VR_HD void start(double a_, double b_)
  {
    a = a_; b = b_; c = 0; d = 0; e = 0; fa = 0; fb = 0; fc = 0; x = a_; result = VERR; stage = 0; j = 0; i = 0;
  }
  VR_HD bool feed(double fx)
  {
    if (stage == 0) { fa = fx; x = b; stage = 1; return false; }
    if (stage == 1) {
      fb = fx;
      if ((fa * fb) >= 0) {
        if (j < 5) { a -= 10; b += 10; j++; x = a; stage = 0; return false; }
        result = VERR;
        return true;
      }
      fc = fb;
      stage = 2;
    }
    else {
      fb = fx;
      i++;
      if (i >= 1000) { result = VERR; return true; }
    }
    double m, p, q, r, s, tol;
    if (fb * fc > 0) { c = a; fc = fa; d = b - a; e = d; }
    if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    tol = 2 * 3e-8 * fabs(b) + 1e-7;
    m = 0.5 * (c - b);
    if (fabs(m) <= tol || fb == 0) { result = b; return true; }
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
    x = b;
    return false;
  }
};

// ---------------------------------------------------------------------------------------
// snow pack (snow_melt.c:119-579, SnowPackEnergyBalance.c:88-328, latent_heat_from_snow.c)
// ---------------------------------------------------------------------------------------
struct SnowPackEB {
  // inputs
  double Dt, Ra, Z, lnz5, AirDens, EactAir, LongSnowIn, Lv, Press, Rain, NetShortUnder, Vpd, Wind, OldTSurf,
         SnowDepth, SnowDensity, SurfaceLiquidWater, SweSurfaceLayer, Tair, TGrnd;
  // outputs of the last evaluation
  double Ra_used0, AdvectedEnergy, DeltaColdContent, GroundFlux, LatentHeat, LatentHeatSub, NetLongUnder,
         RefreezeEnergy, SensibleHeat, vapor_flux, surface_flux;
  VR_HD double operator()(double TSurf)
  {
    double TMean = TSurf;
    if (Wind > 0.0) Ra_used0 = Ra / stability_correction(Z - 0., lnz5, TMean, Tair, Wind);
    else Ra_used0 = HUGE_RESIST;
    double Tmp = TMean + KELVIN;
    NetLongUnder = LongSnowIn - STEFAN_B * Tmp * Tmp * Tmp * Tmp;
    double NetRad = NetShortUnder + NetLongUnder;
    SensibleHeat = AirDens * Cp * (Tair - TMean) / Ra_used0;
    // latent_heat_from_snow (BlowingMassFlux == 0: BLOWING FALSE)
    double EsSnow = svp(TMean);
    double SurfaceMassFlux = AirDens * (EPS / Press) * (EactAir - EsSnow) / Ra_used0;
    if (Vpd == 0.0 && SurfaceMassFlux < 0.0) SurfaceMassFlux = 0.0;
    double VaporMassFlux = SurfaceMassFlux + 0.;
    if (TMean >= 0.0) {
      LatentHeat = Lv * VaporMassFlux;
      LatentHeatSub = 0;
    }
    else {
      double Ls = (677. - 0.07 * TMean) * JOULESPCAL * GRAMSPKG;
      LatentHeatSub = Ls * VaporMassFlux;
      LatentHeat = 0;
    }
    vapor_flux = VaporMassFlux * Dt / RHO_W;
    surface_flux = SurfaceMassFlux * Dt / RHO_W;
    if (TMean == 0) AdvectedEnergy = (CH_WATER * (Tair)*Rain) / (Dt);
    else AdvectedEnergy = 0.;
    DeltaColdContent = CH_ICE * SweSurfaceLayer * (TSurf - OldTSurf) / (Dt);
    if (SnowDepth > 0.) GroundFlux = 2.9302e-6 * SnowDensity * SnowDensity * (TGrnd - TMean) / SnowDepth / (Dt);
    else GroundFlux = 0;
    DeltaColdContent -= GroundFlux;
    double RestTerm = NetRad + SensibleHeat + LatentHeat + LatentHeatSub + AdvectedEnergy + GroundFlux -
                      DeltaColdContent + 0.;
    RefreezeEnergy = (SurfaceLiquidWater * Lf * RHO_W) / (Dt);
    if (TSurf == 0.0 && RestTerm > -(RefreezeEnergy)) {
      RefreezeEnergy = -RestTerm;
      RestTerm = 0.0;
    }
    else RestTerm += RefreezeEnergy;
    return RestTerm;
  }
};

// the pack words snow_melt reads and writes
struct PackS { double swq, surf_temp, pack_temp, surf_water, pack_water, depth, density, coldcontent; };
struct SnowMeltOut {
  PackS s;
  double melt, NetLongSnow, OldTSurf, advection, deltaCC, latent, latent_sub, refreeze_energy, sensible,
         vapor_flux, aero_used0;
  int fallback;      // 1 when TFALLBACK kept the old surface temperature (snow_melt.c:361-366)
};

VR_HDN SnowMeltOut snow_melt(PackS s, double Le, double NetShortSnow, double Tcanopy, double Tgrnd, double Z0_2,
                             double aero_resist, double air_temp, double delta_t, double density, double LongSnowIn,
                             double pressure, double rainfall, double snowfall, double vp, double vpd, double wind,
                             double z2)
{
  SnowMeltOut o;
  o.fallback = 0;
  double SnowFall = snowfall / 1000., RainFall = rainfall / 1000.;
  o.OldTSurf = s.surf_temp;
  double Ice = s.swq - s.pack_water - s.surf_water;
  double SurfaceSwq = (Ice > MAX_SURFACE_SWE) ? MAX_SURFACE_SWE : Ice;
  double PackSwq = Ice - SurfaceSwq;
  double SurfaceCC = CH_ICE * SurfaceSwq * s.surf_temp;
  double PackCC = CH_ICE * PackSwq * s.pack_temp;
  double SnowFallCC = (air_temp > 0.0) ? 0.0 : CH_ICE * SnowFall * air_temp;
  if (SnowFall > (MAX_SURFACE_SWE - SurfaceSwq) && (MAX_SURFACE_SWE - SurfaceSwq) > SMALLV) {
    double DeltaPackSwq = SurfaceSwq + SnowFall - MAX_SURFACE_SWE, DeltaPackCC;
    if (DeltaPackSwq > SurfaceSwq) DeltaPackCC = SurfaceCC + (SnowFall - MAX_SURFACE_SWE) / SnowFall * SnowFallCC;
    else DeltaPackCC = DeltaPackSwq / SurfaceSwq * SurfaceCC;
    SurfaceSwq = MAX_SURFACE_SWE;
    SurfaceCC += SnowFallCC - DeltaPackCC;
    PackSwq += DeltaPackSwq;
    PackCC += DeltaPackCC;
  }
  else {
    SurfaceSwq += SnowFall;
    SurfaceCC += SnowFallCC;
  }
  s.surf_temp = (SurfaceSwq > 0.0) ? SurfaceCC / (CH_ICE * SurfaceSwq) : 0.0;
  s.pack_temp = (PackSwq > 0.0) ? PackCC / (CH_ICE * PackSwq) : 0.0;
  Ice += SnowFall;
  s.surf_water += RainFall;

  SnowPackEB x;
  x.Dt = delta_t; x.Ra = aero_resist; x.Z = z2; x.lnz5 = (wind > 0.0) ? stability_lnz5(z2, 0., Z0_2) : 0.; x.AirDens = density; x.EactAir = vp;
  x.LongSnowIn = LongSnowIn; x.Lv = Le; x.Press = pressure; x.Rain = RainFall; x.NetShortUnder = NetShortSnow;
  x.Vpd = vpd; x.Wind = wind; x.OldTSurf = o.OldTSurf; x.SnowDepth = s.depth; x.SnowDensity = s.density;
  x.SurfaceLiquidWater = s.surf_water; x.SweSurfaceLayer = SurfaceSwq; x.Tair = Tcanopy; x.TGrnd = Tgrnd;

  // ONE evaluation site of the pack's energy balance: stage 0 = at 0 C (snow_melt.c:316), stage 1 = the evaluations
  // of root_brent (:349-357), stage 2 = at the root (:381)
  double Qnet, vapor_flux = 0, RefreezeEnergy = 0, T = 0.0;
  int stage = 0;
  bool melting = false, solved = false;
  Brent br;
  for (;;) {
    Qnet = x(T);
    if (stage == 0) {
      vapor_flux = x.vapor_flux;
      RefreezeEnergy = x.RefreezeEnergy;
      if (Qnet == 0.0) { melting = true; break; }
      if (SurfaceSwq > MIN_SWQ_EB_THRES) {
        br.start(s.surf_temp - SNOW_DT, s.surf_temp + SNOW_DT);
        T = br.x;
        stage = 1;
        continue;
      }
      s.surf_temp = 999;
      break;
    }
    if (stage == 1) {
      if (!br.feed(Qnet)) { T = br.x; continue; }
      s.surf_temp = br.result;
      if (s.surf_temp <= -998) {    // TFALLBACK (snow_melt.c:361-366)
        s.surf_temp = o.OldTSurf;
        o.fallback = 1;
      }
      if (s.surf_temp > -998 && s.surf_temp < 999) { T = s.surf_temp; stage = 2; continue; }
      vapor_flux = x.vapor_flux;
      RefreezeEnergy = x.RefreezeEnergy;
      break;
    }
    vapor_flux = x.vapor_flux;
    RefreezeEnergy = x.RefreezeEnergy;
    solved = true;
    break;
  }
  if (melting) {
    double SnowMelt;
    s.surf_temp = 0.0;
    if (RefreezeEnergy >= 0.0) {
      double RefrozenWater = RefreezeEnergy / (Lf * RHO_W) * delta_t;
      if (RefrozenWater > s.surf_water) {
        RefrozenWater = s.surf_water;
        RefreezeEnergy = RefrozenWater * Lf * RHO_W / (delta_t);
      }
      SurfaceSwq += RefrozenWater;
      Ice += RefrozenWater;
      s.surf_water -= RefrozenWater;
      if (s.surf_water < 0.0) s.surf_water = 0.0;
      SnowMelt = 0.0;
    }
    else SnowMelt = fabs(RefreezeEnergy) / (Lf * RHO_W) * delta_t;
    if (s.surf_water < -(vapor_flux)) {
      vapor_flux = -(s.surf_water);
      s.surf_water = 0.0;
    }
    else s.surf_water += vapor_flux;
    if (SnowMelt < Ice) {
      if (SnowMelt <= PackSwq) {
        s.surf_water += SnowMelt;
        PackSwq -= SnowMelt;
        Ice -= SnowMelt;
      }
      else {
        s.surf_water += SnowMelt + s.pack_water;
        s.pack_water = 0.0;
        PackSwq = 0.0;
        Ice -= SnowMelt;
        SurfaceSwq = Ice;
      }
    }
    else {
      SnowMelt = Ice;
      s.surf_water += Ice;
      SurfaceSwq = 0.0;
      s.surf_temp = 0.0;
      PackSwq = 0.0;
      s.pack_temp = 0.0;
      Ice = 0.0;
      RefreezeEnergy = RefreezeEnergy / fabs(RefreezeEnergy) * SnowMelt * Lf * RHO_W / (delta_t);
    }
  }
  else if (solved) {
    SurfaceSwq += s.surf_water;
    Ice += s.surf_water;
    s.surf_water = 0.0;
    if (SurfaceSwq < -(vapor_flux)) {
      vapor_flux = -SurfaceSwq;
      SurfaceSwq = 0.0;
      Ice = PackSwq;
    }
    else {
      SurfaceSwq += vapor_flux;
      Ice += vapor_flux;
    }
  }
  double melt;
  double MaxLiquidWater = LIQUID_WATER_CAPACITY * SurfaceSwq;
  if (s.surf_water > MaxLiquidWater) {
    melt = s.surf_water - MaxLiquidWater;
    s.surf_water = MaxLiquidWater;
  }
  else melt = 0.0;
  s.pack_water += melt;
  double PackRefreezeEnergy = s.pack_water * Lf * RHO_W;
  if (PackCC < -PackRefreezeEnergy) {
    PackSwq += s.pack_water;
    Ice += s.pack_water;
    s.pack_water = 0.0;
    if (PackSwq > 0.0) {
      PackCC = PackSwq * CH_ICE * s.pack_temp + PackRefreezeEnergy;
      s.pack_temp = PackCC / (CH_ICE * PackSwq);
      if (s.pack_temp > 0.) s.pack_temp = 0.;
    }
    else s.pack_temp = 0.0;
  }
  else {
    s.pack_temp = 0.0;
    double DeltaPackSwq = -PackCC / (Lf * RHO_W);
    s.pack_water -= DeltaPackSwq;
    PackSwq += DeltaPackSwq;
    Ice += DeltaPackSwq;
  }
  MaxLiquidWater = LIQUID_WATER_CAPACITY * PackSwq;
  if (s.pack_water > MaxLiquidWater) {
    melt = s.pack_water - MaxLiquidWater;
    s.pack_water = MaxLiquidWater;
  }
  else melt = 0.0;
  Ice = PackSwq + SurfaceSwq;
  if (Ice > MAX_SURFACE_SWE) {
    SurfaceCC = CH_ICE * s.surf_temp * SurfaceSwq;
    PackCC = CH_ICE * s.pack_temp * PackSwq;
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
    s.pack_temp = PackCC / (CH_ICE * PackSwq);
    s.surf_temp = SurfaceCC / (CH_ICE * SurfaceSwq);
  }
  else s.pack_temp = 0.0;
  s.swq = Ice + s.pack_water + s.surf_water;
  if (s.swq == 0.0) {
    s.surf_temp = 0.0;
    s.pack_temp = 0.0;
  }
  o.melt = melt * 1000.;
  s.coldcontent = SurfaceCC;
  o.vapor_flux = vapor_flux * -1.;
  o.advection = x.AdvectedEnergy;
  o.deltaCC = x.DeltaColdContent;
  o.latent = x.LatentHeat;
  o.latent_sub = x.LatentHeatSub;
  o.sensible = x.SensibleHeat;
  o.refreeze_energy = RefreezeEnergy;
  o.NetLongSnow = x.NetLongUnder;
  o.aero_used0 = x.Ra_used0;
  o.s = s;
  return o;
}

VR_HD double new_snow_density(double air_temp)   // snow_utility.c, DENS_BRAS
{
  air_temp = air_temp * 9. / 5. + 32.;
  if (air_temp > 0) return NEW_SNOW_DENSITY + 1000. * (air_temp / 100.) * (air_temp / 100.);
  return NEW_SNOW_DENSITY;
}

VR_HDN double snow_density(double surf_temp, double depth_in, double new_snow, double sswq, double Tair, double dt)
{   // snow_utility.c snow_density, DENS_BRAS
  double density_new = (new_snow > 0.) ? new_snow_density(Tair) : 0.0;
  double Tavg = surf_temp + KELVIN, depth = depth_in, swq = sswq, density;
  if (new_snow > 0) {
    if (depth > 0.) {
      double delta_depth = (((new_snow / 25.4) * (depth / 0.0254)) / (swq / 0.0254) *
                            VR_POW((depth / 0.0254) / 10., 0.35)) * 0.0254;
      if (delta_depth > 0.9 * depth) delta_depth = 0.9 * depth;
      double depth_new = new_snow / density_new;
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
  else density = 1000. * swq / depth_in;
  if (depth > 0.) {
    double overburden = 0.5 * G_GRAV * RHO_W * swq;
    double viscosity = SNDENS_ETA0 * VR_EXP(-SNDENS_C5 * (Tavg - KELVIN) + SNDENS_C6 * density);
    double delta_depth = overburden / viscosity * depth * dt * 3600;
    if (delta_depth > 0.9 * depth) delta_depth = 0.9 * depth;
    depth -= delta_depth;
    density = 1000. * swq / depth;
  }
  return density;
}

VR_HDN double snow_albedo(double new_snow, double swq, double albedo, double cold_content, double dt,
                         int last_snow, int MELTING)
{   // snow_utility.c snow_albedo
  if (new_snow > TraceSnow && cold_content < 0.0) albedo = NEW_SNOW_ALB;
  else if (swq > 0.0) {
    if (cold_content < 0.0 && !MELTING)
      albedo = NEW_SNOW_ALB * VR_POW(SNOW_ALB_ACCUM_A, VR_POW((double)last_snow * dt / 24., SNOW_ALB_ACCUM_B));
    else
      albedo = NEW_SNOW_ALB * VR_POW(SNOW_ALB_THAW_A, VR_POW((double)last_snow * dt / 24., SNOW_ALB_THAW_B));
  }
  else albedo = 0;
  return albedo;
}

// ---------------------------------------------------------------------------------------
// evaporation (canopy_evap.c:46-529, arno_evap.c:62-203)
// ---------------------------------------------------------------------------------------
// what the evaporation leaves need of a vegetation tile (veg_lib / veg_con): by value, 9 registers
struct VegP { double rarc, rmin, RGL; float root[MAXL]; };

// canopy_evap.c:332-529 (transpiration); returns the layer evaporation of the step (mm)
VR_HDN M3 transpiration(M3 moist, double LAI, VegP vp, CellC cc, int NL, PenmanAir pa, double rad, double vpd,
                        double net_short, double air_temp, double ra, double dryFrac, double delta_t)
{
  M3 le;
  double avail[MAXL], moist1 = 0.0, Wcr1 = 0.0, moist2;
  const int il = NL - 1;
#pragma unroll
  for (int i = 0; i < MAXL; i++) {
    le.v[i] = 0.;
    avail[i] = 0.;
    if (i < il) {
      if (vp.root[i] > 0.) {
        avail[i] = 0. + ((moist.v[i] - 0.) * 1.);
        moist1 += avail[i];
        Wcr1 += cc.Wcr(i);
      }
    }
  }
  const double moist_il = (il == 2) ? moist.v[2] : moist.v[1];
  const double Wcr_il = (il == 2) ? cc.Wcr(2) : cc.Wcr(1);
  const double root_il = (il == 2) ? (double)vp.root[2] : (double)vp.root[1];
  const float rootf_il = (il == 2) ? vp.root[2] : vp.root[1];
  moist2 = 0. + ((moist_il - 0.) * 1.);
  if (il == 2) avail[2] = moist2; else avail[1] = moist2;
  (void)root_il;
  if ((moist1 >= Wcr1 && moist2 >= Wcr_il && Wcr1 > 0.) || (moist1 >= Wcr1 && (1 - rootf_il) >= 0.5) ||
      (moist2 >= Wcr_il && rootf_il >= 0.5)) {
    double rc = calc_rc(vp.rmin, net_short, vp.RGL, air_temp, vpd, LAI, 1.0);
    double evap = penman(pa, rad, vpd, ra, rc, vp.rarc) * delta_t / SEC_PER_DAY * dryFrac;
    double root_sum = 1.0, spare_evap = 0.0;
#pragma unroll
    for (int i = 0; i < MAXL; i++) {
      if (i < NL) {
        if (avail[i] >= cc.Wcr(i)) le.v[i] = evap * (double)vp.root[i];
        else {
          double gsm_inv = (avail[i] >= cc.Wpwp(i)) ? (avail[i] - cc.Wpwp(i)) / cc.wcr_m_wpwp(i) : 0.0;
          le.v[i] = evap * gsm_inv * (double)vp.root[i];
          root_sum -= vp.root[i];
          spare_evap = evap * (double)vp.root[i] * (1.0 - gsm_inv);
        }
      }
    }
    if (spare_evap > 0.0) {
#pragma unroll
      for (int i = 0; i < MAXL; i++)
        if (i < NL && avail[i] >= cc.Wcr(i)) le.v[i] += (double)vp.root[i] * spare_evap / root_sum;
    }
  }
  else {
#pragma unroll
    for (int i = 0; i < MAXL; i++) {
      if (i < NL) {
        double gsm_inv;
        if (avail[i] >= cc.Wcr(i)) gsm_inv = 1.0;
        else if (avail[i] >= cc.Wpwp(i) && avail[i] < cc.Wcr(i)) gsm_inv = (avail[i] - cc.Wpwp(i)) / cc.wcr_m_wpwp(i);
        else gsm_inv = 0.0;
        if (gsm_inv > 0.0) {
          double rc = calc_rc(vp.rmin, net_short, vp.RGL, air_temp, vpd, LAI, gsm_inv);
          le.v[i] = penman(pa, rad, vpd, ra, rc, vp.rarc) * delta_t / SEC_PER_DAY * dryFrac * (double)vp.root[i];
        }
        else le.v[i] = 0.0;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXL; i++) {
    if (i < NL) {
      if (le.v[i] > moist.v[i] - cc.Wpwp(i)) le.v[i] = moist.v[i] - cc.Wpwp(i);
      if (le.v[i] < 0.0) le.v[i] = 0.0;
    }
  }
  return le;
}

// canopy_evap.c:46-330.  Returns Evap (m/s), the canopy evaporation, throughfall and new canopy storage (mm) and
// the layer evaporation.
struct CanopyOut { double Evap, canopyevap, throughfall, Wdew; M3 le; };
VR_HDN CanopyOut canopy_evap(M3 moist, double LAI, double Wdmax, bool CALC_EVAP, VegP vp, CellC cc, int NL,
                             PenmanAir pa, double Wdew_in, double delta_t, double rad, double vpd, double net_short,
                             double air_temp, double ra, double ppt)
{
  CanopyOut o;
  double throughfall = 0, tmp_Wdew = Wdew_in, canopyevap, f, wet23 = 0;
#pragma unroll
  for (int i = 0; i < MAXL; i++) o.le.v[i] = 0;
  if (tmp_Wdew > Wdmax) {
    throughfall = tmp_Wdew - Wdmax;
    tmp_Wdew = Wdmax;
  }
  if (LAI > 0) {
    wet23 = VR_POW((tmp_Wdew / Wdmax), (2.0 / 3.0));
    // calc_rc(rs = 0) == 0
    canopyevap = 240 * wet23 * penman(pa, rad, vpd, ra, 0., vp.rarc) * delta_t / SEC_PER_DAY;
  }
  else canopyevap = 0;
  if (canopyevap > 0.0 && delta_t == SEC_PER_DAY)
    f = (1.0 < ((tmp_Wdew + ppt) / canopyevap)) ? 1.0 : ((tmp_Wdew + ppt) / canopyevap);
  else if (canopyevap > 0.0) f = (1.0 < ((tmp_Wdew) / canopyevap)) ? 1.0 : ((tmp_Wdew) / canopyevap);
  else f = 1.0;
  canopyevap *= f;
  double dryFrac;
  if (Wdmax > 0) {
    if (!(LAI > 0)) wet23 = VR_POW((tmp_Wdew / Wdmax), (2.0 / 3.0));
    dryFrac = 1.0 - f * wet23;
  }
  else dryFrac = 0;
  tmp_Wdew += ppt - canopyevap;
  if (tmp_Wdew < 0.0) tmp_Wdew = 0.0;
  if (tmp_Wdew > Wdmax) {
    throughfall += tmp_Wdew - Wdmax;
    tmp_Wdew = Wdmax;
  }
  if (CALC_EVAP) o.le = transpiration(moist, LAI, vp, cc, NL, pa, rad, vpd, net_short, air_temp, ra, dryFrac, delta_t);
  o.canopyevap = canopyevap;
  o.throughfall = throughfall;
  o.Wdew = tmp_Wdew;
  double tmp_Evap = canopyevap;
#pragma unroll
  for (int i = 0; i < MAXL; i++)
    if (i < NL) tmp_Evap += o.le.v[i];
  o.Evap = 0. + tmp_Evap / (1000. * delta_t);
  return o;
}

// arno_evap.c:62-203.  Evap (m/s; VERR on the reference's ERROR returns) and layer[0].evap (mm)
struct ArnoOut { double Evap, evap0; };
VR_HDN ArnoOut arno_evap(double moist0, CellC cc, PenmanAir pa, double rad, double vpd, double ra, double delta_t)
{
  ArnoOut o;
  o.evap0 = 0.;
  double moist = 0. + (moist0 - 0.) * 1., evap, tmp, ratio;
  if (moist > cc.arno_maxm()) moist = cc.arno_maxm();
  double Epot = 240 * penman(pa, rad, vpd, ra, 0.0, RARC_SOIL) * delta_t / SEC_PER_DAY;
  if (cc.b() == -1.0) tmp = cc.arno_max_infil();
  else {
    ratio = 1.0 - (moist) / (cc.arno_maxm());
    if (ratio > 1.0 || ratio < 0.0) { o.Evap = VERR; return o; }
    ratio = VR_POW(ratio, cc.inv_b1());
    tmp = cc.arno_max_infil() * (1.0 - ratio);
  }
  if (tmp >= cc.arno_max_infil()) evap = Epot;
  else {
    ratio = tmp / cc.arno_max_infil();
    ratio = 1.0 - ratio;
    if (ratio > 1.0 || ratio < 0.0) { o.Evap = VERR; return o; }
    if (ratio != 0.0) ratio = VR_POW(ratio, cc.b());
    double as = 1 - ratio;
    ratio = VR_POW(ratio, cc.inv_b());
    double dummy = 1.0, tmpsum = 1.0;
    const double b = cc.b();
#pragma unroll 1
    for (int n = 1; n <= 30; n++) {       // arno_evap.c:168-173; the inner product is a running power
      tmpsum = (n == 1) ? ratio : tmpsum * ratio;
      dummy += b * tmpsum / (b + n);
    }
    double beta_asp = as + (1.0 - as) * (1.0 - ratio) * dummy;
    evap = 240 * Epot * beta_asp;
  }
  if (evap > 0.0) {
    if (moist > cc.resid(0)) {
      if (evap > moist - cc.resid(0)) evap = moist - cc.resid(0);
    }
    else evap = 0.0;
  }
  o.evap0 = evap;
  o.Evap = 0. + evap / 1000. / delta_t;
  return o;
}

// ---------------------------------------------------------------------------------------
// canopy snow interception (snow_intercept.c:8-560, func_canopy_energy_bal.c:29-290, massrelease.c)
// ---------------------------------------------------------------------------------------
VR_HD void mass_release(double &IntSnow, double &TempInt, double &Released, double &Drip)
{   // massrelease.c:35-93: the tail recursion written as a loop
  for (;;) {
    if (IntSnow > MIN_INTERCEPTION_STORAGE) {
      double Threshold = 0.10 * IntSnow, MaxRelease = 0.17 * IntSnow;
      if (TempInt >= Threshold) {
        Drip += Threshold;
        IntSnow -= Threshold;
        TempInt -= Threshold;
        double rel;
        if (IntSnow < MIN_INTERCEPTION_STORAGE) rel = 0.0;
        else rel = ((IntSnow - MIN_INTERCEPTION_STORAGE) < MaxRelease) ? (IntSnow - MIN_INTERCEPTION_STORAGE) : MaxRelease;
        Released += rel;
        IntSnow -= rel;
        continue;
      }
      double TempDrip = (TempInt < IntSnow) ? TempInt : IntSnow;
      Drip += TempDrip;
      IntSnow -= TempDrip;
      return;
    }
    double TempDrip = (TempInt < IntSnow) ? TempInt : IntSnow;
    Drip += TempDrip;
    IntSnow -= TempDrip;
    TempInt = 0.0;
    return;
  }
}

struct CanopyEB {
  // inputs
  double delta_t, AirDens, EactAir, Press, Le, Tcanopy, Vpd, Ra0, Ra1, Rainfall, IntRainOrg, IntSnow,
         LongOverIn, LongUnderOut, NetShortOver, LAI, Wdmax;
  M3 moist; VegP vp; CellC cc; PenmanAir pa; int NL;
  // the vegetation words canopy_evap updates through the reference's pointers (veg_var->Wdew, canopyevap, throughfall)
  double v_Wdew, v_canopyevap, v_throughfall;
  // outputs of the last evaluation
  double Ra_used0, Ra_used1, IntRain, AdvectedEnergy, LatentHeat, LatentHeatSub, LongOverOut, NetLongOver,
         RefreezeEnergy, SensibleHeat, VaporMassFlux;
  VR_HD double operator()(double Tfoliage)
  {
    double Tmp = Tfoliage + KELVIN;
    LongOverOut = STEFAN_B * (Tmp * Tmp * Tmp * Tmp);
    double NetRadiation = NetShortOver + LongOverIn + LongUnderOut - 2 * (LongOverOut);
    NetLongOver = LongOverIn - (LongOverOut);
    if (IntSnow > 0) {
      Ra_used0 = Ra0;
      Ra_used1 = Ra1;
      Ra_used1 *= 10.;                      // AR_406_FULL (func_canopy_energy_bal.c:195-196)
      double EsSnow = svp(Tfoliage);
      VaporMassFlux = AirDens * (EPS / Press) * (EactAir - EsSnow) / Ra_used1 / RHO_W;
      if (Vpd == 0.0 && VaporMassFlux < 0.0) VaporMassFlux = 0.0;
      double Ls = (677. - 0.07 * Tfoliage) * JOULESPCAL * GRAMSPKG;
      LatentHeatSub = Ls * VaporMassFlux * RHO_W;
      LatentHeat = 0;
      v_throughfall = 0;
    }
    else {
      Ra_used0 = Ra0;
      Ra_used1 = Ra1;
      const CanopyOut ce = canopy_evap(moist, LAI, Wdmax, false, vp, cc, NL, pa, IntRainOrg * 1000., delta_t, NetRadiation,
                                       Vpd, NetShortOver, Tcanopy, Ra_used1, Rainfall * 1000);
      v_Wdew = ce.Wdew; v_canopyevap = ce.canopyevap; v_throughfall = ce.throughfall;
      IntRain = v_Wdew / 1000.;    // *Wdew aliases *IntRain in the reference call (snow_intercept.c:399-416)
      LatentHeat = Le * ce.Evap * RHO_W;
      LatentHeatSub = 0;
    }
    SensibleHeat = AirDens * Cp * (Tcanopy - Tfoliage) / Ra_used1;
    AdvectedEnergy = (CH_WATER * Tcanopy * Rainfall) / (delta_t);
    double RestTerm = SensibleHeat + LatentHeat + LatentHeatSub + NetRadiation + AdvectedEnergy;
    if (IntSnow > 0) {
      RefreezeEnergy = (IntRainOrg * Lf * RHO_W) / (delta_t);
      if (Tfoliage == 0.0 && RestTerm > -(RefreezeEnergy)) {
        RefreezeEnergy = -RestTerm;
        RestTerm = 0.0;
      }
      else RestTerm += RefreezeEnergy;
    }
    else RefreezeEnergy = 0;
    return RestTerm;
  }
};

struct InterceptIn {
  double Tfoliage, LongUnderOut, snow_canopy, tmp_int_storage, Wdew, Wdmax, LAI, canopyevap, throughfall;
  double density, vp, pressure, vpd;          // atmos fields of the (sub-)step
  double Dt, Le, LongOverIn, ShortOverIn, Tcanopy, bare_albedo, Ra0, Ra1, wind1, RainFall, SnowFall;
};
struct InterceptOut {
  double Tfoliage, snow_canopy, tmp_int_storage, Wdew, canopyevap, throughfall, RainFall, SnowFall;
  double AlbedoOver, canopy_refreeze, NetLongOver, NetShortOver, LongOverOut, canopy_vapor_flux, Ra_used0, Ra_used1;
  int fallback;
};

// Dt seconds; F == 1.  RainFall / SnowFall in mm: in above the canopy, out below it.
VR_HDN InterceptOut snow_intercept(InterceptIn in, M3 moist, VegP vp, CellC cc, int NL, PenmanAir pa)
{
  InterceptOut o;
  o.fallback = 0;
  double Tfoliage = in.Tfoliage, tmp_int_storage = in.tmp_int_storage;
  double IntRain = in.Wdew, IntSnow = in.snow_canopy, MaxInt = in.Wdmax, LAI = in.LAI;
  double RainFall = in.RainFall, SnowFall = in.SnowFall;
  const double Dt = in.Dt, Tcanopy = in.Tcanopy, wind1 = in.wind1;
  RainFall /= 1000.;
  SnowFall /= 1000.;
  IntRain /= 1000.;
  MaxInt /= 1000.;
  double IntRainOrg = IntRain;
  IntSnow /= 1.;
  IntRain /= 1.;
  double InitialSnowInt = IntSnow, Drip = 0.0, ReleasedMass = 0.0, OldTfoliage = Tfoliage;
  double Imax1 = 4.0 * LAI_SNOW_MULTIPLIER * LAI, MaxSnowInt, DeltaSnowInt;
  if (Tfoliage < -1.0 && Tfoliage > -3.0) MaxSnowInt = (Tfoliage * 3.0 / 2.0) + (11.0 / 2.0);
  else if (Tfoliage > -1.0) MaxSnowInt = 4.0;
  else MaxSnowInt = 1.0;
  MaxSnowInt *= LAI_SNOW_MULTIPLIER * LAI;
  if (MaxSnowInt > 0) {
    DeltaSnowInt = (1 - IntSnow / MaxSnowInt) * SnowFall;
    if (DeltaSnowInt + IntSnow > MaxSnowInt) DeltaSnowInt = MaxSnowInt - IntSnow;
    if (DeltaSnowInt < 0.0) DeltaSnowInt = 0.0;
  }
  else DeltaSnowInt = 0.0;
  if (Tfoliage < -3.0 && DeltaSnowInt > 0.0 && wind1 > 1.0) {
    double BlownSnow = (0.2 * wind1 - 0.2) * DeltaSnowInt;
    if (BlownSnow >= DeltaSnowInt) BlownSnow = DeltaSnowInt;
    DeltaSnowInt -= BlownSnow;
  }
  if (IntSnow + DeltaSnowInt > Imax1) DeltaSnowInt = 0.0;
  double SnowThroughFall = (SnowFall - DeltaSnowInt) * 1. + (SnowFall) * (1 - 1.);
  if (SnowFall == 0 && IntSnow < MIN_SWQ_EB_THRES) {
    SnowThroughFall += IntSnow;
    DeltaSnowInt -= IntSnow;
  }
  IntSnow += DeltaSnowInt;
  if (IntSnow < SMALLV) IntSnow = 0.0;
  double MaxWaterInt = LIQUID_WATER_CAPACITY * (IntSnow) + MaxInt, RainThroughFall;
  if ((IntRain + RainFall) <= MaxWaterInt) {
    IntRain += RainFall;
    RainThroughFall = RainFall * (1 - 1.);
  }
  else {
    RainThroughFall = (IntRain + RainFall - MaxWaterInt) * 1. + (RainFall * (1 - 1.));
    IntRain = MaxWaterInt;
  }
  if (RainFall == 0 && IntRain < MIN_SWQ_EB_THRES) {
    RainThroughFall += IntRain;
    IntRain = 0.0;
  }
  if (IntRain + IntSnow > Imax1) {
    double Overload = (IntSnow + IntRain) - Imax1;
    double IntRainFract = IntRain / (IntRain + IntSnow);
    double IntSnowFract = IntSnow / (IntRain + IntSnow);
    IntRain = IntRain - Overload * IntRainFract;
    IntSnow = IntSnow - Overload * IntSnowFract;
    RainThroughFall = RainThroughFall + (Overload * IntRainFract) * 1.;
    SnowThroughFall = SnowThroughFall + (Overload * IntSnowFract) * 1.;
  }
  if (IntRain + IntSnow < SMALLV) Tfoliage = Tcanopy;

  CanopyEB x;
  x.delta_t = Dt; x.AirDens = in.density; x.EactAir = in.vp; x.Press = in.pressure; x.Le = in.Le; x.Tcanopy = Tcanopy;
  x.Vpd = in.vpd; x.Ra0 = in.Ra0; x.Ra1 = in.Ra1; x.Rainfall = RainFall; x.IntRainOrg = IntRainOrg; x.IntSnow = IntSnow;
  x.LongOverIn = in.LongOverIn; x.LongUnderOut = in.LongUnderOut; x.LAI = LAI; x.Wdmax = in.Wdmax;
  x.moist = moist; x.vp = vp; x.cc = cc; x.pa = pa; x.NL = NL;
  x.v_Wdew = in.Wdew; x.v_canopyevap = in.canopyevap; x.v_throughfall = in.throughfall;
  x.IntRain = IntRain; x.RefreezeEnergy = 0; x.VaporMassFlux = 0;
  x.Ra_used0 = in.Ra0; x.Ra_used1 = in.Ra1;
  x.NetLongOver = 0; x.LongOverOut = 0;

  // ONE evaluation site of the canopy energy balance: stage 0 = at 0 C (snow_intercept.c:358-372), stage 1 = the
  // evaluations of root_brent (:399-416), stage 2 = at the root (:425-440)
  double Qnet, T;
  int stage;
  Brent br;
  if (IntSnow > 0 || SnowFall > 0) {
    o.AlbedoOver = NEW_SNOW_ALB;
    o.NetShortOver = (1. - o.AlbedoOver) * in.ShortOverIn;
    x.NetShortOver = o.NetShortOver;
    T = 0.;
    stage = 0;
  }
  else {
    o.AlbedoOver = in.bare_albedo;
    o.NetShortOver = (1. - o.AlbedoOver) * in.ShortOverIn;
    x.NetShortOver = o.NetShortOver;
    br.start(Tfoliage - SNOW_DT, Tfoliage + SNOW_DT);
    T = br.x;
    stage = 1;
  }
  for (;;) {
    Qnet = x(T);
    if (stage == 0) {
      if (Qnet != 0) {
        const double Tlower = (Tfoliage <= 0.) ? Tfoliage - SNOW_DT : -SNOW_DT;
        br.start(Tlower, 0.);
        T = br.x;
        stage = 1;
        continue;
      }
      Tfoliage = 0.;
      break;
    }
    if (stage == 1) {
      if (!br.feed(Qnet)) { T = br.x; continue; }
      Tfoliage = br.result;
      if (Tfoliage <= -998) {     // TFALLBACK (snow_intercept.c:418-423)
        Tfoliage = OldTfoliage;
        o.fallback = 1;
      }
      T = Tfoliage;
      stage = 2;
      continue;
    }
    break;
  }
  IntRain = x.IntRain;
  double RefreezeEnergy = x.RefreezeEnergy, VaporMassFlux = x.VaporMassFlux;
  if (IntSnow <= 0) RainThroughFall = x.v_throughfall / 1000.;
  RefreezeEnergy *= Dt;
  MaxWaterInt = LIQUID_WATER_CAPACITY * (IntSnow) + MaxInt;
  VaporMassFlux *= Dt;
  if (Tfoliage == 0) {
    if (-(VaporMassFlux) > IntRain) {
      VaporMassFlux = -(IntRain);
      IntRain = 0.;
    }
    else IntRain += VaporMassFlux;
    double PotSnowMelt;
    if (RefreezeEnergy < 0)
      PotSnowMelt = ((-RefreezeEnergy / Lf / RHO_W) < IntSnow) ? (-RefreezeEnergy / Lf / RHO_W) : IntSnow;
    else PotSnowMelt = 0;
    if ((IntRain + PotSnowMelt) <= MaxWaterInt) {
      IntSnow -= PotSnowMelt;
      IntRain += PotSnowMelt;
    }
    else {
      double ExcessSnowMelt = PotSnowMelt + IntRain - MaxWaterInt;
      IntSnow -= MaxWaterInt - (IntRain);
      IntRain = MaxWaterInt;
      if (IntSnow < 0.0) IntSnow = 0.0;
      if (SnowThroughFall > 0.0 && InitialSnowInt <= MIN_INTERCEPTION_STORAGE) {
        Drip += ExcessSnowMelt;
        IntSnow -= ExcessSnowMelt;
        if (IntSnow < 0.0) IntSnow = 0.0;
      }
      else tmp_int_storage += ExcessSnowMelt;
      mass_release(IntSnow, tmp_int_storage, ReleasedMass, Drip);
    }
    MaxWaterInt = LIQUID_WATER_CAPACITY * (IntSnow) + MaxInt;
    if (IntRain > MaxWaterInt) {
      Drip += IntRain - MaxWaterInt;
      IntRain = MaxWaterInt;
    }
  }
  else {
    tmp_int_storage = 0.0;
    if (-RefreezeEnergy > -(IntRain)*Lf) {
      IntSnow += fabs(RefreezeEnergy) / Lf;
      IntRain -= fabs(RefreezeEnergy) / Lf;
      RefreezeEnergy = 0.0;
    }
    else {
      IntSnow += IntRain;
      IntRain = 0.0;
    }
    if (-(VaporMassFlux) > IntSnow) {
      VaporMassFlux = -(IntSnow);
      IntSnow = 0.0;
    }
    else IntSnow += VaporMassFlux;
  }
  if (IntSnow == 0 && IntRain > MaxInt) {
    RainThroughFall += IntRain - MaxInt;
    IntRain = MaxInt;
  }
  RainFall = RainThroughFall + Drip;
  SnowFall = SnowThroughFall + ReleasedMass;
  VaporMassFlux *= -1.;
  RainFall *= 1000.;
  SnowFall *= 1000.;
  IntRain *= 1000.;
  o.Tfoliage = Tfoliage; o.snow_canopy = IntSnow; o.tmp_int_storage = tmp_int_storage;
  o.Wdew = IntRain; o.canopyevap = x.v_canopyevap; o.throughfall = x.v_throughfall;
  o.RainFall = RainFall; o.SnowFall = SnowFall;
  o.canopy_refreeze = RefreezeEnergy / Dt;
  o.NetLongOver = x.NetLongOver;
  o.LongOverOut = x.LongOverOut;
  o.canopy_vapor_flux = VaporMassFlux;
  o.Ra_used0 = x.Ra_used0;
  o.Ra_used1 = x.Ra_used1;
  (void)Qnet;
  return o;
}

// ---------------------------------------------------------------------------------------
// runoff.c:7-576 (Nfrost == 1, no ice): infiltration, hourly drainage, ARNO baseflow
// ---------------------------------------------------------------------------------------
struct AsatOut { double A, runoff; };
VR_HDN AsatOut runoff_and_asat(CellC cc, double top_moist, double inflow)
{
  AsatOut o;
  if (top_moist > cc.top_max()) top_moist = cc.top_max();
  o.A = 1.0 - VR_POW((1.0 - top_moist / cc.top_max()), cc.ex());
  if (inflow == 0.0) { o.runoff = 0.0; return o; }
  double i_0 = cc.top_max_infil() * (1.0 - VR_POW((1.0 - o.A), cc.inv_b()));
  if (cc.top_max_infil() == 0.0) o.runoff = inflow;
  else if ((i_0 + inflow) > cc.top_max_infil()) o.runoff = inflow - cc.top_max() + top_moist;
  else {
    double basis = 1.0 - (i_0 + inflow) / cc.top_max_infil();
    o.runoff = (inflow - cc.top_max() + top_moist + cc.top_max() * VR_POW(basis, cc.one_plus_b()));
  }
  if (o.runoff < 0.) o.runoff = 0.;
  return o;
}

// Division by a per-cell constant.  The device multiplies by the reciprocal prepared on the host
// (1 ulp apart from the reference's x/d; the drainage loop alone has 96 such divisions per tile-day);
// the host / VR_EXACT_LIBM build divides, which keeps tests/hostsim bit-identical to the reference.
#if defined(__CUDA_ARCH__) && !defined(VR_EXACT_LIBM)
#define VR_DIVC(x, d, inv_d) ((x) * (inv_d))
#else
#define VR_DIVC(x, d, inv_d) ((x) / (d))
#endif

// moist: in/out (mm); evap: the layer evaporation of the step (mm), the bottom layer's may be reduced.
// Layers are scalars (no dynamically indexed arrays -> no local memory in the hourly loop); the
// reference's while-loops that push excess water upwards (runoff.c:362-391,448-475) are unrolled.
struct RunoffOut { M3 moist; double evap_bottom, asat, runoff, baseflow; };

// The per-cell constants of the hourly loop.  LoopRegs: values (read once from the constant block, live in registers);
// LoopSmem (device): the same 18 words in a shared-memory column that the step kernel staged, re-read every hour, so
// that the loop runs in few registers and the kernel fits all its warps on the SMs at once.
enum { LK_R0 = 0, LK_R1, LK_RB, LK_M0, LK_M1, LK_MB, LK_KS0, LK_KS1, LK_EX0, LK_EX1, LK_IQ0, LK_IQ1, LK_IQB, LK_BF_FRAC, LK_BF_NL,
       LK_WS, LK_INV_ONE_M_WS, LK_C, LK_N };
// row of the constant block that holds loop constant k (NL layers)
VR_HD int loop_const_row(int k, int NL)
{
  const int LB = NL - 1;
  switch (k) {
    case LK_R0: return CC_FL0 + 0 * CLF_N + CLF_RESID;
    case LK_R1: return CC_FL0 + 1 * CLF_N + CLF_RESID;
    case LK_RB: return CC_FL0 + LB * CLF_N + CLF_RESID;
    case LK_M0: return CC_BL0 + 0 * CLB_N + CLB_MAX_MOIST;
    case LK_M1: return CC_BL0 + 1 * CLB_N + CLB_MAX_MOIST;
    case LK_MB: return CC_BL0 + LB * CLB_N + CLB_MAX_MOIST;
    case LK_KS0: return CC_BL0 + 0 * CLB_N + CLB_KSAT24;
    case LK_KS1: return CC_BL0 + 1 * CLB_N + CLB_KSAT24;
    case LK_EX0: return CC_BL0 + 0 * CLB_N + CLB_EXPT;
    case LK_EX1: return CC_BL0 + 1 * CLB_N + CLB_EXPT;
    case LK_IQ0: return CC_BL0 + 0 * CLB_N + CLB_INV_QDEN;
    case LK_IQ1: return CC_BL0 + 1 * CLB_N + CLB_INV_QDEN;
    case LK_IQB: return CC_BL0 + LB * CLB_N + CLB_INV_QDEN;
    case LK_BF_FRAC: return CC_BF_FRAC;
    case LK_BF_NL: return CC_BF_NL;
    case LK_WS: return CC_WS;
    case LK_INV_ONE_M_WS: return CC_INV_ONE_M_WS;
    default: return CC_C;
  }
}
struct LoopRegs {
  double v[LK_N], qd0, qd1, qdB, one_m_Ws;
  VR_HD double k(int i) const { return v[i]; }
};
template <int NL>
VR_HD LoopRegs loop_regs(CellC cc)
{
  constexpr int LB = NL - 1;
  LoopRegs K;
#pragma unroll
  for (int i = 0; i < LK_N; i++) K.v[i] = cc.g(loop_const_row(i, NL));
  K.qd0 = cc.qden(0); K.qd1 = cc.qden(1); K.qdB = cc.qden(LB); K.one_m_Ws = cc.one_m_Ws();
  return K;
}
#if defined(__CUDACC__)
struct LoopSmem {
  uint32_t a;                   // shared-memory address of the thread's column of the stage
  int S;                        // row stride in bytes
  double qd0, qd1, qdB, one_m_Ws;     // unused on the device (divisions by constants are reciprocal multiplications)
  __device__ __forceinline__ double k(int i) const
  {
    double x;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a + (uint32_t)(i * S)));
    return x;
  }
};
#endif

// moist: in/out (mm); evap: the layer evaporation of the step (mm), the bottom layer's may be reduced.
// Layers are scalars (no dynamically indexed arrays -> no local memory in the hourly loop); the
// reference's while-loops that push excess water upwards (runoff.c:362-391,448-475) are unrolled.
template <int NL, class LoopK>
VR_HD RunoffOut runoff_core(CellC cc, const LoopK &K, int dt, double ppt, M3 moist, M3 evap)
{
  constexpr int LB = NL - 1;                      // bottom layer
#define r0 K.k(LK_R0)
#define r1 K.k(LK_R1)
#define rB K.k(LK_RB)
#define m0 K.k(LK_M0)
#define m1 K.k(LK_M1)
#define mB K.k(LK_MB)
#define ks0 K.k(LK_KS0)
#define ks1 K.k(LK_KS1)
#define ex0 K.k(LK_EX0)
#define ex1 K.k(LK_EX1)
#define iq0 K.k(LK_IQ0)
#define iq1 K.k(LK_IQ1)
#define iqB K.k(LK_IQB)
#define bf_frac K.k(LK_BF_FRAC)
#define bf_nl K.k(LK_BF_NL)
#define Ws K.k(LK_WS)
#define inv_one_m_Ws K.k(LK_INV_ONE_M_WS)
#define cexp K.k(LK_C)
#define qd0 K.qd0
#define qd1 K.qd1
#define qdB K.qdB
#define one_m_Ws K.one_m_Ws
  double liq0 = moist.v[0], liq1 = moist.v[1], liqB = moist.v[LB];
  double ev[MAXL];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    double e = evap.v[l] / (double)dt;
    if (e > 0) {
      double avail_liq = (moist.v[l] - cc.resid(l));
      if (avail_liq < 0) avail_liq = 0;
      double evap_fraction = (avail_liq > 0) ? e / avail_liq : 1.0;
      e = avail_liq * evap_fraction;
    }
    ev[l] = e;
  }
  const double ev0 = ev[0], ev1 = ev[1], evB = ev[LB];
  double A, runoff_, baseflow = 0;
  {
    const AsatOut a = runoff_and_asat(cc, (NL == 3) ? (0. + liq0) + liq1 : 0. + liq0, ppt);
    A = a.A; runoff_ = a.runoff;
  }
  const double tmp_dt_runoff = runoff_ / (double)dt, dt_inflow = ppt / (double)dt;
#if defined(__CUDA_ARCH__) && !defined(VR_EXACT_LIBM) && !defined(VR_POW_CALL_IN_LOOP)
  // branch-free power for the two drainage terms of an hour; legal for exponents >= 1 (Brooks-Corey: expt > 3)
  const bool fastpow = ex0 >= 1.0 && (NL != 3 || ex1 >= 1.0);
  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(&vr_pow_tables);
#define VR_DRAIN_POW(x, y) (fastpow ? drain_pow((x), (y), tb) : m_pow((x), (y)))
#else
#define VR_DRAIN_POW(x, y) m_pow_inl((x), (y))
#endif
#pragma unroll 1
  for (int ts = 0; ts < dt; ts++) {
    // the constants used more than once in an hour: one read each (LoopSmem re-reads shared memory on every K.k())
    const double c_r0 = r0, c_m0 = m0, c_rB = rB, c_mB = mB;
    const double c_r1 = (NL == 3) ? r1 : 0., c_m1 = (NL == 3) ? m1 : 0.;
    // Brooks-Corey drainage out of the upper layers (runoff.c:330-340).  A layer at or below its residual moisture has
    // t == r, i.e. the power of zero: the reference's `liq > resid ? ... : 0` needs no branch
    double t0 = liq0 - ev0;
    if (t0 < c_r0) t0 = c_r0;
    double Q0 = ks0 * VR_DRAIN_POW(VR_DIVC(t0 - c_r0, qd0, iq0), ex0);
    if (!(liq0 > c_r0)) Q0 = 0.;
    double Q1 = 0.;
    if (NL == 3) {
      double t1 = liq1 - ev1;
      if (t1 < c_r1) t1 = c_r1;
      Q1 = ks1 * VR_DRAIN_POW(VR_DIVC(t1 - c_r1, qd1, iq1), ex1);
      if (!(liq1 > c_r1)) Q1 = 0.;
    }
    // top layer
    liq0 = liq0 + (dt_inflow - tmp_dt_runoff) - (Q0 + ev0);
    if (liq0 > c_m0) { Q0 += liq0 - c_m0; liq0 = c_m0; }
    if (liq0 < 0) { Q0 += liq0; liq0 = 0; }
    if (liq0 < c_r0) { Q0 += liq0 - c_r0; liq0 = c_r0; }
    double Qin = Q0;                       // what reaches the layer below
    if (NL == 3) {
      liq1 = liq1 + Qin - (Q1 + ev1);
      if (liq1 > c_m1) {
        double ex = liq1 - c_m1;
        liq1 = c_m1;
        liq0 += ex;
        if (liq0 > c_m0) { runoff_ += liq0 - c_m0; liq0 = c_m0; }
      }
      if (liq1 < 0) { Q1 += liq1; liq1 = 0; }
      if (liq1 < c_r1) { Q1 += liq1 - c_r1; liq1 = c_r1; }
      Qin = Q1;
    }
    // bottom layer: ARNO baseflow (runoff.c:420-435)
    const double rel_moist = VR_DIVC(liqB - c_rB, qdB, iqB);
    double dt_baseflow = bf_frac * rel_moist;
    const double c_Ws = Ws;
    if (rel_moist > c_Ws) {
      const double frac = VR_DIVC(rel_moist - c_Ws, one_m_Ws, inv_one_m_Ws);
      const double c_c = cexp;
      dt_baseflow += bf_nl * ((c_c == 2.0) ? frac * frac : VR_POW(frac, c_c));   // pow(x,2) == x*x
    }
    if (dt_baseflow < 0) dt_baseflow = 0;
    liqB += Qin - (evB + dt_baseflow);
    if (liqB < c_rB) { dt_baseflow += liqB - c_rB; liqB = c_rB; }
    if (liqB > c_mB) {
      double ex = liqB - c_mB;
      liqB = c_mB;
      if (NL == 3) {
        liq1 += ex;
        if (liq1 > c_m1) { ex = liq1 - c_m1; liq1 = c_m1; }
        else ex = 0;
      }
      if (ex > 0) {
        liq0 += ex;
        if (liq0 > c_m0) { runoff_ += liq0 - c_m0; liq0 = c_m0; }
      }
    }
    baseflow += dt_baseflow;
    if (NL == 2) liq1 = liqB;
  }
#undef VR_DRAIN_POW
  RunoffOut o;
  o.evap_bottom = evap.v[LB];
  if (baseflow < 0) {
    o.evap_bottom += baseflow;
    baseflow = 0;
  }
  const AsatOut a = runoff_and_asat(cc, (NL == 3) ? (0. + liq0) + liq1 : 0. + liq0, 0.);
  o.moist.v[0] = liq0;
  o.moist.v[1] = (NL == 3) ? liq1 : liqB;
  o.moist.v[2] = (NL == 3) ? liqB : moist.v[2];
  o.asat = a.A;
  o.runoff = runoff_;
  o.baseflow = baseflow;
  return o;
#undef r0
#undef r1
#undef rB
#undef m0
#undef m1
#undef mB
#undef ks0
#undef ks1
#undef ex0
#undef ex1
#undef iq0
#undef iq1
#undef iqB
#undef bf_frac
#undef bf_nl
#undef Ws
#undef inv_one_m_Ws
#undef cexp
#undef qd0
#undef qd1
#undef qdB
#undef one_m_Ws
}

VR_HDN RunoffOut runoff_step(CellC cc, int NL, int dt, double ppt, M3 moist, M3 evap)
{
  if (NL == 3) return runoff_core<3>(cc, loop_regs<3>(cc), dt, ppt, moist, evap);
  return runoff_core<2>(cc, loop_regs<2>(cc), dt, ppt, moist, evap);
}
#if defined(__CUDACC__)
// the same with the loop constants in the staged shared-memory column
__device__ __noinline__ RunoffOut runoff_step_staged(CellC cc, LoopSmem K, int NL, int dt, double ppt, M3 moist, M3 evap)
{
  if (NL == 3) return runoff_core<3>(cc, K, dt, ppt, moist, evap);
  return runoff_core<2>(cc, K, dt, ppt, moist, evap);
}
#endif

// ---------------------------------------------------------------------------------------
// one unit, one record
// ---------------------------------------------------------------------------------------
// bits of the step's `extras`
enum { EX_PET = 1, EX_ZWT = 2 };

// compute_pot_evap.c:8-80 with the reference-cover parameters of global.h:53-67.  ra0/ra1/raU: the unit's raw
// aerodynamic resistances [0], [1], [UnderStory]; used0/used1: the ones the iteration ended with
// (surface_fluxes.c:868-894 rebuilds a stability factor from them and applies it to every reference cover).
// calc_rc() of cover i sees the net_short of cover i-1 (the reference assigns it afterwards); only NATVEG
// depends on it, and it gets TALL's.
struct P6 { double v[6]; };
VR_HDN P6 pot_evap6(VegP vp, int overstory, const double *tl, const double *ap, PenmanAir pa, int UnderStory, double wind,
                    double raU, double ra1, double used0, double used1, double shortwave, double net_longwave,
                    double tair, double vpd, int dt_hours)
{
  P6 pot;
  double sf0, sf1;
  if (used0 == HUGE_RESIST) sf0 = HUGE_RESIST; else sf0 = used0 / raU;
  if (used1 == used0) sf1 = sf0;
  else if (used1 == HUGE_RESIST) sf1 = HUGE_RESIST;
  else sf1 = used1 / ra1;
  double net_short = 0.;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double ref_rmin = (i < 2) ? 0.0 : 100, ref_rarc = (i < 2) ? 0.0 : 25;
    const double ref_lai = (i < 2) ? 1.0 : ((i == 2) ? 2.88 : 4.45);
    const double ref_alb = (i == 0) ? BARE_SOIL_ALBEDO : ((i == 1) ? 0.08 : 0.23);
    double rs, rarc, RGL, lai, albedo, a0, a1;
    if (i < 4) {
      rs = ref_rmin; rarc = ref_rarc; RGL = (i < 2) ? 0.0 : 100; lai = ref_lai; albedo = ref_alb;
      const double *q = ap + 3 * i;
      if (wind > 0.) { a0 = ((UnderStory == 2) ? q[2] : q[0]) / wind; a1 = q[1] / wind; }
      else { a0 = HUGE_RESIST; a1 = HUGE_RESIST; }
    }
    else {
      rs = (i == 5) ? 0 : vp.rmin; rarc = vp.rarc; RGL = vp.RGL; lai = tl[TL_LAI]; albedo = tl[TL_ALB];
      a0 = raU; a1 = ra1;
    }
    double rc;
    if (rs == 0) rc = 0;
    else if (lai == 0) rc = RSMAX;
    else if (i == 2 || i == 3) { rc = rs / (lai * 0.5); rc = (rc > RSMAX) ? RSMAX : rc; }
    else rc = calc_rc(rs, net_short, RGL, tair, vpd, lai, 1.0);
    const double s0 = (sf0 == HUGE_RESIST) ? HUGE_RESIST : a0 * sf0, s1 = (sf1 == HUGE_RESIST) ? HUGE_RESIST : a1 * sf1;
    const double ra = (i < 4 || !overstory) ? s0 : s1;
    net_short = (1.0 - albedo) * shortwave;
    const double net_rad = net_short + net_longwave;
    pot.v[i] = penman(pa, net_rad, vpd, ra, rc, rarc) * dt_hours / 24.0;
  }
  return pot;
}

// compute_zwt.c:7-45
VR_HD double zwt_curve(CellC cc, int NL, int k, double moist)
{
  double zwt = MISSINGV;
  int i = ZWT_NP - 1;
  while (i >= 1 && moist > cc.zwt_m(NL, k, i)) i--;
  if (i == ZWT_NP - 1) {
    if (moist < cc.zwt_m(NL, k, i)) zwt = 999;
    else if (moist == cc.zwt_m(NL, k, i)) zwt = cc.zwt_z(NL, k, i);
  }
  else
    zwt = cc.zwt_z(NL, k, i + 1) + (cc.zwt_z(NL, k, i) - cc.zwt_z(NL, k, i + 1)) * (moist - cc.zwt_m(NL, k, i + 1)) /
                                       (cc.zwt_m(NL, k, i) - cc.zwt_m(NL, k, i + 1));
  return zwt;
}

// wrap_compute_zwt (compute_zwt.c:48-113), called at the end of runoff() (runoff.c:503)
struct ZwtOut { double zwt, lumped; };
VR_HDN ZwtOut water_table(CellC cc, int NL, M3 moist)
{
  double total_depth = 0, lz[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; l++) {
    lz[l] = 0;
    if (l < NL) total_depth += cc.depth(l);
  }
#pragma unroll
  for (int l = 0; l < MAXL; l++)
    if (l < NL) lz[l] = zwt_curve(cc, NL, l, moist.v[l]);
  // bottom layer
  if (NL == 3) { if (lz[2] == 999) lz[2] = -total_depth * 100; }
  else { if (lz[1] == 999) lz[1] = -total_depth * 100; }
  double tmp_depth = total_depth, zwt = 0;
  bool found = false;
#pragma unroll
  for (int l = MAXL - 1; l >= 0; l--) {
    if (l < NL && !found) {
      if (cc.max_moist(l) - moist.v[l] <= SMALLV) tmp_depth -= cc.depth(l);
      else {
        found = true;
        if (l < NL - 1) zwt = (lz[l] != 999) ? lz[l] : -tmp_depth * 100;
        else zwt = lz[l];
      }
    }
  }
  if (!found) zwt = 0;
  double tmp_moist = 0;
#pragma unroll
  for (int k = 0; k < MAXL; k++)
    if (k < NL) tmp_moist += moist.v[k];
  double zl = zwt_curve(cc, NL, NL + 1, tmp_moist);
  if (zl == 999) zl = -total_depth * 100;
  ZwtOut o;
  o.zwt = zwt; o.lumped = zl;
  return o;
}

// carried state split by temperature: the words every record touches, and the snow words that only the snow path reads
struct HotS { M3 moist; double Wdew, T0, T1, LongUnderOut, Tfoliage; };
struct SnowS {
  double swq, surf_temp, pack_temp, surf_water, pack_water, density, depth, albedo, coldcontent, coverage, snow_canopy,
         tmp_int_storage;
  int last_snow, MELTING;
};
// what the reference's else-branch of solve_snow leaves behind when there is no snow anywhere (solve_snow.c:540-577)
VR_HD SnowS snow_none()
{
  SnowS z;
  z.swq = 0; z.surf_temp = 0; z.pack_temp = 0; z.surf_water = 0; z.pack_water = 0; z.density = 0; z.depth = 0;
  z.albedo = NEW_SNOW_ALB; z.coldcontent = 0; z.coverage = 0; z.snow_canopy = 0; z.tmp_int_storage = 0;
  z.last_snow = (int)MISSINGV; z.MELTING = 0;
  return z;
}

// calc_rainonly.c:8-58 + solve_snow.c:185-197: rain / snow partition of the (sub-)step's precipitation
VR_HD void rain_snow_split(double Tair, double step_prec, double max_snow_temp, double min_rain_temp, double &rainfall,
                           double &snowfall)
{
  double rainonly = 0.;
  if (Tair < max_snow_temp && Tair > min_rain_temp) rainonly = (Tair - min_rain_temp) / (max_snow_temp - min_rain_temp) * step_prec;
  else if (Tair >= max_snow_temp) rainonly = step_prec;
  snowfall = 1. * (step_prec - rainonly);
  rainfall = 1. * rainonly;
  if (snowfall < 1e-5) snowfall = 0.;
}

// per (unit, month) canopy terms handed to the step (full_energy.c:210-227,306-316)
struct CanopyM { double vegcover, LAI, Wdmax, albedo, surf_atten; };

// results of ONE pass of surface_fluxes' sub-step loop (surface_fluxes.c:480-1025) that the caller sums / averages
struct SurfOut {
  double ppt, canopyevap, throughfall, vapor_flux, canopy_vapor_flux, melt, out_prec, out_rain, out_snow;
  M3 evap, bef;
  double Tsurf, NetShortAtmos, NetLongAtmos, LongOverIn, LongUnderIn, AlbedoOver, AlbedoUnder, grnd_flux, deltaH,
         ra_used0, ra_used1;
  P6 pot;
  int snow_flag, fb_snow, fb_fol, err;
};

// One pass of the sub-step loop for one unit: solve_snow.c:7-577, then calc_surf_energy_bal.c / func_surf_energy_bal.c in
// water-balance mode (FULL_ENERGY FALSE: Tsurf = Tair), canopy_evap / arno_evap.  SNOW == false is the instance for a
// unit without any snow state and without snowfall in this pass: the snow words are the constants of snow_none() and none
// of the snow code is compiled in; the caller picks the instance per WARP, so a warp never runs both.
// h: hot state (moist is read only; Wdew, T0, T1, LongUnderOut, Tfoliage are advanced); sn: snow state (advanced);
// coverage, surf_atten: surface_fluxes' variables of those names, which live across the passes of a record (a pass with
// snow on a tile without overstory leaves surf_atten = 1 behind for the rest of the record, solve_snow.c:215); Tgrnd:
// energy->T[0] at the start of the record; dt_hours: length of this pass.
template <int extras, bool SNOW>
VR_HD SurfOut surface_pass(CellC cc, VegP vp, bool is_veg, bool overstory, double Tfactor, double Pfactor, CanopyM cm,
                           const double *ae, Forc f, int NL, int dt_hours, double max_snow_temp, double min_rain_temp,
                           double Tgrnd, HotS &h, SnowS &sn, double &coverage, double &surf_atten, const double *tl,
                           const double *ap, bool sync_ok, int rec_dt_hours)
{
  (void)sync_ok;
  SurfOut o;
  o.err = 0; o.fb_snow = 0; o.fb_fol = 0;
  const double delta_t = (double)dt_hours * 3600.;
  double BareAlbedo, vegcover, LAI, Wdmax;
  if (is_veg) { vegcover = cm.vegcover; LAI = cm.LAI; Wdmax = cm.Wdmax; BareAlbedo = cm.albedo; }
  else { vegcover = 0; LAI = 0; Wdmax = 0; BareAlbedo = BARE_SOIL_ALBEDO; }
  double v_Wdew = h.Wdew, v_canopyevap = 0, v_throughfall = 0;

  // ---- prepare_full_energy.c:55-110 -> soil_conduction.c:5-61: kappa, Cs of layers 0 and 1 from the
  //      pre-step moisture; Kdry, Ksat and the solid heat capacity are per-cell constants ----
  double kappa[2], Cs[2];
#pragma unroll
  for (int l = 0; l < 2; l++) {
    double mv = h.moist.v[l] / cc.depth(l) / 1000;
    double K;
    if (mv > 0.) {
      double Sr = mv / cc.poros(l);
      double Ke = 0.7 * VR_LOG10(Sr) + 1.0;
      K = cc.ksat_m_kdry(l) * Ke + cc.kdry(l);
      if (K < cc.kdry(l)) K = cc.kdry(l);
    }
    else K = cc.kdry(l);
    kappa[l] = K;
    double c = cc.cs_solid(l);
    c += 4.2e6 * mv;
    c += 1.3e3 * (1. - (cc.soilfr(l) + mv + 0.));
    Cs[l] = c;
  }

  // ---- aerodynamics of the current cover: Ra = RA/wind, U = UC*wind (CalcAerodynamic.c:232-257) ----
  double Ra0, Ra1, Ra2;
  if (f.wind > 0.) { Ra0 = ae[AE_RA0] / f.wind; Ra1 = ae[AE_RA1] / f.wind; Ra2 = ae[AE_RA2] / f.wind; }
  else { Ra0 = HUGE_RESIST; Ra1 = HUGE_RESIST; Ra2 = HUGE_RESIST; }
  const double U0 = ae[AE_U0] * f.wind, U1 = ae[AE_U1] * f.wind, U2 = ae[AE_U2] * f.wind;
  double ra_used0 = Ra0, ra_used1 = Ra1;
  (void)U1; (void)U2; (void)Ra2;

  // ---- surface_fluxes.c:480-560 ----
  const double Tair = f.air_temp + Tfactor;
  const double step_prec = f.prec * Pfactor;
  const double Tcanopy = Tair;
  const double step_depth_old = sn.depth, step_surf_temp_old = sn.surf_temp;
  const PenmanAir pa = penman_air(Tair, cc.half_elev_lapse(), cc.elev());

  // ---- solve_snow.c:7-577 ----
  double snowfall, rainfall;
  rain_snow_split(Tair, step_prec, max_snow_temp, min_rain_temp, rainfall, snowfall);
  const double out_prec = snowfall + rainfall, out_rain = rainfall, out_snow = snowfall, store_snowfall = snowfall;
  const double Le = (2.501e6 - 0.002361e6 * Tair);
  int UnderStory = 0;
  double ShortUnderIn = f.shortwave, LongUnderIn = f.longwave;
  double melt = 0., ppt = 0., melt_energy = 0., NetLongSnow = 0., NetShortSnow = 0., NetShortGrnd = 0.,
         delta_coverage = 0., AlbedoUnder, OldTSurf = 0., vapor_flux = 0., canopy_vapor_flux = 0.;
  double es_latent = 0., es_latent_sub = 0., es_sensible = 0., es_advection = 0., es_deltaCC = 0., es_refreeze = 0.;
  double AlbedoOver = 0., NetLongOver = 0., NetShortOver = 0., LongOverIn = 0.;
  int snow_flag = 0;
  if (SNOW && (sn.swq > 0 || snowfall > 0. || (sn.snow_canopy > 0. && overstory))) {
    UnderStory = (sn.swq > 0 || snowfall > 0) ? 2 : 0;
    snow_flag = 1;
    if (!overstory) surf_atten = 1.;
    const double old_coverage = sn.coverage;
    if (is_veg) {
      if (overstory) {
        ShortUnderIn *= surf_atten;
        double ShortOverIn = (1. - surf_atten) * f.shortwave;
        ShortOverIn /= vegcover;
        InterceptIn ii;
        ii.Tfoliage = h.Tfoliage; ii.LongUnderOut = h.LongUnderOut; ii.snow_canopy = sn.snow_canopy;
        ii.tmp_int_storage = sn.tmp_int_storage; ii.Wdew = v_Wdew; ii.Wdmax = Wdmax; ii.LAI = LAI;
        ii.canopyevap = v_canopyevap; ii.throughfall = v_throughfall;
        ii.density = f.density; ii.vp = f.vp; ii.pressure = f.pressure; ii.vpd = f.vpd;
        ii.Dt = (double)dt_hours * 3600; ii.Le = Le; ii.LongOverIn = f.longwave; ii.ShortOverIn = ShortOverIn;
        ii.Tcanopy = Tcanopy; ii.bare_albedo = BareAlbedo; ii.Ra0 = Ra0; ii.Ra1 = Ra1; ii.wind1 = U1;
        ii.RainFall = rainfall; ii.SnowFall = snowfall;
        const InterceptOut io = snow_intercept(ii, h.moist, vp, cc, NL, pa);
        h.Tfoliage = io.Tfoliage; sn.snow_canopy = io.snow_canopy; sn.tmp_int_storage = io.tmp_int_storage;
        v_Wdew = io.Wdew; v_canopyevap = io.canopyevap;
        rainfall = io.RainFall; snowfall = io.SnowFall;
        o.fb_fol = io.fallback;
        LongUnderIn = io.LongOverOut;
        AlbedoOver = io.AlbedoOver; NetLongOver = io.NetLongOver; NetShortOver = io.NetShortOver;
        canopy_vapor_flux = io.canopy_vapor_flux;
        ra_used0 = io.Ra_used0; ra_used1 = io.Ra_used1;
        v_throughfall = rainfall + snowfall;
        LongOverIn = f.longwave;
      }
      else if (snowfall > 0. && v_Wdew > 0.) {
        rainfall += v_Wdew;
        v_throughfall = rainfall + snowfall;
        v_Wdew = 0.;
        h.Tfoliage = Tair;
      }
      else {
        v_throughfall = rainfall + snowfall;
        h.Tfoliage = Tair;
      }
      v_throughfall = (1 - vegcover) * (out_prec) + vegcover * v_throughfall;
      rainfall = (1 - vegcover) * (out_rain) + vegcover * (rainfall);
      snowfall = (1 - vegcover) * (out_snow) + vegcover * (snowfall);
      canopy_vapor_flux *= vegcover;
      sn.snow_canopy *= vegcover;
      v_Wdew *= vegcover;
      v_canopyevap *= vegcover;
      NetShortOver *= vegcover;
      NetLongOver *= vegcover;
    }
    if (sn.swq > 0.0 || snowfall > 0) {
      const double old_swq = sn.swq;
      UnderStory = 2;
      if (sn.swq > 0 && store_snowfall == 0) {
        sn.last_snow++;
        sn.albedo = snow_albedo(snowfall, sn.swq, sn.albedo, sn.coldcontent, (double)dt_hours, sn.last_snow, sn.MELTING);
        AlbedoUnder = (coverage * sn.albedo + (1. - coverage) * BareAlbedo);
      }
      else {
        sn.last_snow = 0;
        sn.albedo = NEW_SNOW_ALB;
        AlbedoUnder = sn.albedo;
      }
      NetShortSnow = (1.0 - AlbedoUnder) * (ShortUnderIn);
      PackS pk;
      pk.swq = sn.swq; pk.surf_temp = sn.surf_temp; pk.pack_temp = sn.pack_temp; pk.surf_water = sn.surf_water;
      pk.pack_water = sn.pack_water; pk.depth = sn.depth; pk.density = sn.density; pk.coldcontent = sn.coldcontent;
      const SnowMeltOut mo = snow_melt(pk, Le, NetShortSnow, Tcanopy, Tgrnd, ae[AE_ROUGH2], Ra2, Tair, (double)dt_hours * 3600,
                                       f.density, LongUnderIn, f.pressure, rainfall, snowfall, f.vp, f.vpd, U2, ae[AE_REF2]);
      sn.swq = mo.s.swq; sn.surf_temp = mo.s.surf_temp; sn.pack_temp = mo.s.pack_temp; sn.surf_water = mo.s.surf_water;
      sn.pack_water = mo.s.pack_water; sn.depth = mo.s.depth; sn.density = mo.s.density; sn.coldcontent = mo.s.coldcontent;
      o.fb_snow = mo.fallback;
      ra_used0 = mo.aero_used0;
      vapor_flux = mo.vapor_flux; NetLongSnow = mo.NetLongSnow; OldTSurf = mo.OldTSurf;
      es_advection = mo.advection; es_deltaCC = mo.deltaCC; es_latent = mo.latent; es_latent_sub = mo.latent_sub;
      es_refreeze = mo.refreeze_energy; es_sensible = mo.sensible;
      melt = mo.melt;
      ppt += melt;
      if (sn.swq > 0.) {
        if (sn.surf_temp <= 0) sn.density = snow_density(sn.surf_temp, sn.depth, snowfall, old_swq, Tair, (double)dt_hours);
        else if (sn.last_snow == 0) sn.density = new_snow_density(Tair);
        sn.depth = 1000. * sn.swq / sn.density;
        if (sn.coldcontent >= 0 && ((cc.lat() >= 0 && (f.doy > 60 && f.doy < 273)) || (cc.lat() < 0 && (f.doy < 60 || f.doy > 273))))
          sn.MELTING = 1;
        else if (sn.MELTING && snowfall > TraceSnow) sn.MELTING = 0;
        sn.coverage = 1.;
      }
      else sn.coverage = 0.;
      delta_coverage = old_coverage - sn.coverage;
      if (delta_coverage != 0) {
        if (old_coverage > sn.coverage) {
          coverage = (old_coverage);
          AlbedoUnder = (coverage - sn.coverage) / (1. - sn.coverage) * sn.albedo;
          AlbedoUnder += (1. - coverage) / (1. - sn.coverage) * BareAlbedo;
          melt_energy = (delta_coverage) * (es_advection - es_deltaCC + es_latent + es_latent_sub + es_sensible + es_refreeze + 0.);
        }
        else {
          coverage = sn.coverage;
          delta_coverage = 0;
        }
      }
      else if (old_coverage == 0 && sn.coverage == 0) {
        delta_coverage = 1.;
        coverage = 0.;
        melt_energy = (es_advection - es_deltaCC + es_latent + es_latent_sub + es_sensible + es_refreeze + 0.);
      }
      NetLongSnow *= (sn.coverage);
      NetShortSnow *= (sn.coverage);
      NetShortGrnd *= (sn.coverage);
      es_latent *= (sn.coverage + delta_coverage);
      es_latent_sub *= (sn.coverage + delta_coverage);
      es_sensible *= (sn.coverage + delta_coverage);
      if (sn.swq == 0) {
        sn.density = 0.; sn.depth = 0.; sn.surf_water = 0; sn.pack_water = 0; sn.surf_temp = 0; sn.pack_temp = 0;
        sn.coverage = 0; sn.MELTING = 0;
      }
      snowfall = 0;
      rainfall = 0;
    }
    else {
      ppt += rainfall;
      AlbedoOver = 0.;
      AlbedoUnder = BareAlbedo;
      sn.last_snow = (int)MISSINGV;
      sn.MELTING = 0;
    }
  }
  else {
    UnderStory = 0;
    snow_flag = 0;
    AlbedoUnder = BareAlbedo;
    h.Tfoliage = Tcanopy;
    sn.MELTING = 0;
    sn.last_snow = (int)MISSINGV;
    sn.albedo = NEW_SNOW_ALB;
  }
  (void)AlbedoUnder;

  VR_PHASE_SYNC(1, sync_ok);
  // ---- surface_fluxes.c:680-695: thin pack -> solved together with the ground ----
  bool INCLUDE_SNOW = false;
  double eo_advection = 0.;
  if (SNOW && sn.surf_temp == 999 && sn.swq > 0) {
    INCLUDE_SNOW = true;
    eo_advection = es_advection;
    sn.surf_temp = step_surf_temp_old;
    melt_energy = 0;
  }

  // ---- calc_surf_energy_bal.c (FULL_ENERGY FALSE: Tsurf = Tair) + func_surf_energy_bal.c ----
  const bool VEG = is_veg && vegcover > 0.0;
  const double T2 = cc.avg_T(), Ts_old = h.T0, T1_old = h.T1, D1 = cc.D1(), D2 = cc.D1();
  const double kappa1 = kappa[0], kappa2 = kappa[1], Cs1 = Cs[0], Cs2 = Cs[1];
  double Tsnow_surf = sn.surf_temp;
  const double Wdew_in = v_Wdew;
  const double snow_coverage = sn.coverage, SnowAlbedo = sn.albedo;
  const double snow_depth = (step_depth_old + sn.depth) / 2.;
  (void)snow_depth;
  const double NetShortBare = (ShortUnderIn * (1. - (snow_coverage + delta_coverage)) * (1. - BareAlbedo) +
                               ShortUnderIn * (delta_coverage) * (1. - SnowAlbedo));
  const double LongBareIn = (1. - snow_coverage) * LongUnderIn;
  double TmpNetLongSnow, TmpNetShortSnow, LongSnowIn;
  if (INCLUDE_SNOW || sn.swq == 0) {
    TmpNetLongSnow = NetLongSnow;
    TmpNetShortSnow = NetShortSnow;
    LongSnowIn = snow_coverage * LongUnderIn;
  }
  else { TmpNetShortSnow = 0.; TmpNetLongSnow = 0.; LongSnowIn = 0.; }
  const double Tsurf = Tcanopy, TMean = Tsurf;
  const double Tmp = TMean + KELVIN;
  if (INCLUDE_SNOW) Tsnow_surf = TMean;
  // estimate_T1.c:8-47 with exp(-D1/dp), exp(D1/dp) hoisted
  double T1;
  {
    double C1 = Cs2 * cc.dp() / D2 * (1. - cc.exp_m());
    double C2 = -(1. - cc.exp_p()) * cc.exp_m();
    double C3 = kappa1 / D1 - kappa2 / D1 + kappa2 / D1 * cc.exp_m();
    T1 = (kappa1 / 2. / D1 / D2 * (TMean) + C1 / delta_t * T1_old + (2. * C2 - 1. + cc.exp_m()) * kappa2 / 2. / D1 / D2 * T2) /
         (C1 / delta_t + kappa2 / D1 / D2 * C2 + C3 / 2. / D2);
  }
  const double grnd_flux = (snow_coverage + (1. - snow_coverage) * surf_atten) *
                           (kappa1 / D1 * ((T1)-TMean) + (kappa2 / D2 * (1. - cc.exp_m()) * (T2 - (T1)))) / 2.;
  const double deltaH = (Cs1 * ((Ts_old + T1_old) - (TMean + T1)) * D1 / delta_t / 2.);
  double eo_deltaCC = 0., eo_refreeze = 0.;
  if (INCLUDE_SNOW) {
    if (TMean > 0) eo_deltaCC = CH_ICE * (sn.swq - sn.surf_water) * (0 - OldTSurf) / delta_t;
    else eo_deltaCC = CH_ICE * (sn.swq - sn.surf_water) * (TMean - OldTSurf) / delta_t;
    eo_refreeze = (sn.surf_water * Lf * sn.density) / delta_t;
    eo_deltaCC *= snow_coverage;
    eo_refreeze *= snow_coverage;
  }
  const double LongBareOut = STEFAN_B * Tmp * Tmp * Tmp * Tmp;
  if (INCLUDE_SNOW) TmpNetLongSnow = (LongSnowIn - snow_coverage * LongBareOut);
  const double NetLongBare = (LongBareIn - (1. - snow_coverage) * LongBareOut);
  const double NetBareRad = (NetShortBare + (NetLongBare) + grnd_flux + deltaH + 0.);
  // Ra_used[0]: StabilityCorrection(TSurf == Tair) == 1 (func_surf_energy_bal.c:684-693)
  const double U_us = (UnderStory == 2) ? U2 : U0, Ra_us = (UnderStory == 2) ? Ra2 : Ra0;
  if (U_us > 0.0) ra_used0 = Ra_us / 1.0;
  else ra_used0 = HUGE_RESIST;
  if (vegcover < 1) {      // bare-gap blend, func_surf_energy_bal.c:712-749 (always for the bare tile)
    if (ra_used0 > 0) {
      double ga_veg = 1 / ra_used0;
      double Ra_bare0 = (U0 > 0.) ? cc.kb_bare() / U0 : HUGE_RESIST;
      Ra_bare0 /= 1.0;
      if (Ra_bare0 > 0) {
        double ga_bare = 1 / Ra_bare0;
        double ga_average = vegcover * ga_veg + (1 - vegcover) * ga_bare;
        ra_used0 = 1 / ga_average;
      }
      else ra_used0 = 0;
    }
    else ra_used0 = 0;
  }
  M3 evap, bef;
  double Evap;
#pragma unroll
  for (int l = 0; l < MAXL; l++) { evap.v[l] = 0; bef.v[l] = 0; }
  VR_PHASE_SYNC(2, sync_ok);
  if (VEG && !snow_flag && vegcover > 0) {
    const CanopyOut ce = canopy_evap(h.moist, LAI, Wdmax, true, vp, cc, NL, pa, Wdew_in, delta_t, NetBareRad, f.vpd, NetShortBare,
                                     Tcanopy, ra_used1, rainfall);
    Evap = ce.Evap; evap = ce.le;
    v_canopyevap = ce.canopyevap; v_throughfall = ce.throughfall; v_Wdew = ce.Wdew;
    if (vegcover < 1) {
      M3 transp = evap;
#pragma unroll
      for (int i = 0; i < MAXL; i++) evap.v[i] = 0;
      Evap *= vegcover;
      const ArnoOut ao = arno_evap(h.moist.v[0], cc, pa, surf_atten * NetBareRad, f.vpd, ra_used0, delta_t);
      if (ao.Evap == VERR) o.err = 1;
      evap.v[0] = ao.evap0;
      Evap += (1 - vegcover) * ao.Evap;
#pragma unroll
      for (int i = 0; i < MAXL; i++) {
        if (i < NL) {
          evap.v[i] = vegcover * transp.v[i] + (1 - vegcover) * evap.v[i];
          if (evap.v[i] > 0) bef.v[i] = 1 - (vegcover * transp.v[i]) / evap.v[i];
          else bef.v[i] = 0;
        }
      }
      v_throughfall = (1 - vegcover) * rainfall + vegcover * v_throughfall;
      v_canopyevap *= vegcover;
      v_Wdew *= vegcover;
    }
  }
  else if (!snow_flag) {
    const ArnoOut ao = arno_evap(h.moist.v[0], cc, pa, NetBareRad, f.vpd, ra_used0, delta_t);
    if (ao.Evap == VERR) o.err = 1;
    Evap = ao.Evap; evap.v[0] = ao.evap0;
#pragma unroll
    for (int i = 0; i < MAXL; i++)
      if (i < NL) bef.v[i] = 1;
  }
  else Evap = 0.;
  if (INCLUDE_SNOW) {
    double latent_heat = -RHO_W * Le * Evap, latent_heat_sub;
    // latent_heat_from_snow with the ground-surface resistance
    double EsSnow = svp(TMean);
    double SurfaceMassFlux = f.density * (EPS / f.pressure) * (f.vp - EsSnow) / ra_used0;
    if (f.vpd == 0.0 && SurfaceMassFlux < 0.0) SurfaceMassFlux = 0.0;
    double VaporMassFlux = SurfaceMassFlux + 0.;
    double tlh, tls;
    if (TMean >= 0.0) { tlh = Le * VaporMassFlux; tls = 0; }
    else { tls = ((677. - 0.07 * TMean) * JOULESPCAL * GRAMSPKG) * VaporMassFlux; tlh = 0; }
    latent_heat += tlh * snow_coverage;
    latent_heat_sub = tls * snow_coverage;
    vapor_flux = VaporMassFlux * delta_t / ICE_DENSITY;
    double sensible_heat = f.density * Cp * (Tcanopy - (TMean)) / ra_used0;
    double error = (NetBareRad + NetShortGrnd + TmpNetShortSnow + 1. * (TmpNetLongSnow)) + sensible_heat +
                   (latent_heat + latent_heat_sub) + 0. * snow_coverage + melt_energy + eo_advection - eo_deltaCC;
    if (Tsnow_surf == 0.0 && error > -(eo_refreeze)) eo_refreeze = -error;
  }
  // ---- calc_surf_energy_bal.c:584-731 ----
  h.T0 = Tsurf;
  h.T1 = T1;
  if (!snow_flag && !INCLUDE_SNOW) ppt = is_veg ? v_throughfall : rainfall;
  double NetLongUnder, NetShortUnder;
  if (INCLUDE_SNOW) {
    NetLongUnder = NetLongBare + TmpNetLongSnow;
    NetShortUnder = NetShortBare + TmpNetShortSnow + NetShortGrnd;
  }
  else {
    NetLongUnder = NetLongBare + NetLongSnow;
    NetShortUnder = NetShortBare + NetShortSnow + NetShortGrnd;
  }
  h.LongUnderOut = LongUnderIn - NetLongUnder;
  const double AlbedoUnderOut = ((1. - (snow_coverage + delta_coverage)) * BareAlbedo + (snow_coverage + delta_coverage) * SnowAlbedo);
  o.Tsurf = (sn.coverage * sn.surf_temp + (1. - sn.coverage) * Tsurf);
  if (INCLUDE_SNOW) {
    if (-(vapor_flux) > sn.swq) vapor_flux = -(sn.swq);
    sn.swq += vapor_flux;
    sn.surf_water += vapor_flux;
    sn.surf_water = (sn.surf_water < 0) ? 0. : sn.surf_water;
    if (eo_refreeze >= 0.0) {
      double refrozen_water = eo_refreeze / (Lf * RHO_W) * delta_t;
      if (refrozen_water > sn.surf_water) refrozen_water = sn.surf_water;
      sn.surf_water -= refrozen_water;
      if (sn.surf_water < 0.0) sn.surf_water = 0.0;
      melt = 0.0;
    }
    else {
      melt = fabs(eo_refreeze) / (Lf * RHO_W) * delta_t;
      sn.swq -= melt;
      if (sn.swq < 0) {
        melt += sn.swq;
        sn.swq = 0;
      }
    }
    if (sn.swq > 0) {
      sn.surf_temp = (Tsurf > 0) ? 0 : Tsurf;
      sn.coldcontent = CH_ICE * sn.surf_temp * sn.swq;
      sn.depth = 1000. * sn.swq / sn.density;
      sn.coverage = 1.;
    }
    else {
      sn.density = 0.; sn.depth = 0.; sn.surf_water = 0; sn.pack_water = 0; sn.surf_temp = 0; sn.pack_temp = 0;
      sn.coverage = 0;
    }
    vapor_flux *= -1;
    ppt += melt;
  }
  // ---- calc_atmos_energy_bal.c (CLOSE_ENERGY FALSE) / surface_fluxes.c:726-760 ----
  if (snow_flag && overstory) {
    o.NetLongAtmos = (1. * NetLongOver + (1. - 1.) * NetLongUnder);
    o.NetShortAtmos = (NetShortOver + NetShortUnder);
  }
  else {
    o.NetLongAtmos = NetLongUnder;
    o.NetShortAtmos = NetShortUnder;
  }
  // put_data.c:1094-1104 picks the over- or understory value with the record's final snow flag, after the averages
  o.LongOverIn = LongOverIn; o.LongUnderIn = LongUnderIn; o.AlbedoOver = AlbedoOver; o.AlbedoUnder = AlbedoUnderOut;
  if (extras & EX_PET)
    o.pot = pot_evap6(vp, overstory ? 1 : 0, tl, ap, pa, UnderStory, f.wind, Ra_us, Ra1, ra_used0, ra_used1, f.shortwave,
                      o.NetLongAtmos, Tair, f.vpd, rec_dt_hours);     // gp->dt, not step_dt (surface_fluxes.c:897)
  h.Wdew = v_Wdew;
  o.ppt = ppt; o.canopyevap = v_canopyevap; o.throughfall = v_throughfall;
  o.vapor_flux = vapor_flux; o.canopy_vapor_flux = canopy_vapor_flux; o.melt = melt;
  o.out_prec = out_prec; o.out_rain = out_rain; o.out_snow = out_snow;
  o.evap = evap; o.bef = bef;
  o.grnd_flux = grnd_flux; o.deltaH = deltaH; o.ra_used0 = ra_used0; o.ra_used1 = ra_used1;
  o.snow_flag = snow_flag;
  return o;
}

}  // namespace vr
