// vic_cr_trig.cuh -- correctly rounded acos(x) and cos(y), y in [0, pi], by double-double arithmetic.
//
// Why: the reference decides the hour of sunrise from the SIGN of cos(zenith) evaluated at the sunrise hour angle
// itself, cosegeom*cos(-hss) + sinegeom with hss = acos(-sinegeom/cosegeom) (mtclim_vic.c:803-834) -- zero up to
// the rounding of acos and cos.  glibc's acos/cos are correctly rounded except in rare cases; CUDA's are 1-2 ulp.
// Evaluating just these two calls per (cell, yearday) correctly rounded makes the device take the reference's
// branch.  Constants from tools/make_cr_tables.py (mpmath, 120 digits).
#ifndef VIC_CR_TRIG_CUH
#define VIC_CR_TRIG_CUH
#include <math.h>

#if defined(__CUDACC__)
#define VC_HD __host__ __device__ __forceinline__
#define VC_HDN __host__ __device__ __noinline__
#else
#define VC_HD inline
#define VC_HDN inline
#endif

namespace vcr {

struct dd { double hi, lo; };

VC_HD dd quick_two_sum(double a, double b) { double s = a + b; return {s, b - (s - a)}; }
VC_HD dd two_sum(double a, double b) { double s = a + b, bb = s - a; return {s, (a - (s - bb)) + (b - bb)}; }
VC_HD dd two_prod(double a, double b) { double p = a * b; return {p, fma(a, b, -p)}; }
VC_HD dd add(dd x, dd y) { dd s = two_sum(x.hi, y.hi); return quick_two_sum(s.hi, s.lo + (x.lo + y.lo)); }
VC_HD dd mul(dd x, dd y) { dd p = two_prod(x.hi, y.hi); return quick_two_sum(p.hi, p.lo + (x.hi * y.lo + x.lo * y.hi)); }
VC_HD dd neg(dd x) { return {-x.hi, -x.lo}; }

// cos and sin of y in [0, 6.8] to ~2^-100
VC_HDN void sincos_dd(double y, dd &c, dd &s)
{
  const double PI16_H = 0x1.921fb54442d18p-3, PI16_M = 0x1.1a62633145c07p-57, PI16_L = -0x1.f1976b7ed8fbcp-113;
  const double T[17][4] = {   // {cos hi, cos lo, sin hi, sin lo} of k*pi/16
    {0x1.0000000000000p+0, 0x0.0p+0, 0x0.0p+0, 0x0.0p+0},
    {0x1.f6297cff75cb0p-1, 0x1.562172a361fd3p-56, 0x1.8f8b83c69a60bp-3, -0x1.26d19b9ff8d82p-57},
    {0x1.d906bcf328d46p-1, 0x1.457e610231ac2p-56, 0x1.87de2a6aea963p-2, -0x1.72cedd3d5a610p-57},
    {0x1.a9b66290ea1a3p-1, 0x1.9f630e8b6dac8p-60, 0x1.1c73b39ae68c8p-1, 0x1.b25dd267f6600p-55},
    {0x1.6a09e667f3bcdp-1, -0x1.bdd3413b26456p-55, 0x1.6a09e667f3bcdp-1, -0x1.bdd3413b26456p-55},
    {0x1.1c73b39ae68c8p-1, 0x1.b25dd267f6600p-55, 0x1.a9b66290ea1a3p-1, 0x1.9f630e8b6dac8p-60},
    {0x1.87de2a6aea963p-2, -0x1.72cedd3d5a610p-57, 0x1.d906bcf328d46p-1, 0x1.457e610231ac2p-56},
    {0x1.8f8b83c69a60bp-3, -0x1.26d19b9ff8d82p-57, 0x1.f6297cff75cb0p-1, 0x1.562172a361fd3p-56},
    {0x0.0p+0, 0x0.0p+0, 0x1.0000000000000p+0, 0x0.0p+0},
    {-0x1.8f8b83c69a60bp-3, 0x1.26d19b9ff8d82p-57, 0x1.f6297cff75cb0p-1, 0x1.562172a361fd3p-56},
    {-0x1.87de2a6aea963p-2, 0x1.72cedd3d5a610p-57, 0x1.d906bcf328d46p-1, 0x1.457e610231ac2p-56},
    {-0x1.1c73b39ae68c8p-1, -0x1.b25dd267f6600p-55, 0x1.a9b66290ea1a3p-1, 0x1.9f630e8b6dac8p-60},
    {-0x1.6a09e667f3bcdp-1, 0x1.bdd3413b26456p-55, 0x1.6a09e667f3bcdp-1, -0x1.bdd3413b26456p-55},
    {-0x1.a9b66290ea1a3p-1, -0x1.9f630e8b6dac8p-60, 0x1.1c73b39ae68c8p-1, 0x1.b25dd267f6600p-55},
    {-0x1.d906bcf328d46p-1, -0x1.457e610231ac2p-56, 0x1.87de2a6aea963p-2, -0x1.72cedd3d5a610p-57},
    {-0x1.f6297cff75cb0p-1, -0x1.562172a361fd3p-56, 0x1.8f8b83c69a60bp-3, -0x1.26d19b9ff8d82p-57},
    {-0x1.0000000000000p+0, 0x0.0p+0, 0x0.0p+0, 0x0.0p+0}};
  const double F[20][2] = {   // 1/n!
    {1.0, 0.0}, {1.0, 0.0},
    {0x1.0000000000000p-1, 0x0.0p+0}, {0x1.5555555555555p-3, 0x1.5555555555555p-57},
    {0x1.5555555555555p-5, 0x1.5555555555555p-59}, {0x1.1111111111111p-7, 0x1.1111111111111p-63},
    {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65}, {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73},
    {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
    {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76}, {0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80},
    {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83}, {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
    {0x1.93974a8c07c9dp-37, 0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-41, 0x1.1d8656b0ee8cbp-97},
    {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
    {0x1.6827863b97d97p-53, 0x1.eec01221a8b0bp-107}, {0x1.2f49b46814157p-57, 0x1.2650f61dbdcb4p-112}};
  int k = (int)(y * (16.0 / 3.14159265358979323846) + 0.5);
  if (k < 0) k = 0;
  if (k > 35) k = 35;
  // t = y - k*pi/16
  dd p = two_prod((double)k, PI16_H);
  dd t = two_sum(y - p.hi, -(p.lo + ((double)k * PI16_M + (double)k * PI16_L)));
  dd u = mul(t, t);
  dd ps = {F[19][0], F[19][1]}, pc = {F[18][0], F[18][1]};
  for (int m = 8; m >= 0; m--) {
    ps = add(dd{F[2 * m + 1][0], F[2 * m + 1][1]}, neg(mul(u, ps)));
    pc = add(dd{F[2 * m][0], F[2 * m][1]}, neg(mul(u, pc)));
  }
  dd st = mul(t, ps), ct = pc;
  int kk = k & 31;                                      // period 2 pi; the table holds the first half turn
  const double sgn = kk > 16 ? -1.0 : 1.0;
  if (kk > 16) kk = 32 - kk;
  dd ck = {T[kk][0], T[kk][1]}, sk = {sgn * T[kk][2], sgn * T[kk][3]};
  c = add(mul(ck, ct), neg(mul(sk, st)));
  s = add(mul(sk, ct), mul(ck, st));
}

VC_HD double cos_cr(double y)            // |y| <= 6.8
{
  dd c, s;
  sincos_dd(fabs(y), c, s);
  return c.hi + c.lo;
}
VC_HD double sin_cr(double y)            // |y| <= 6.8
{
  dd c, s;
  sincos_dd(fabs(y), c, s);
  double r = s.hi + s.lo;
  return y < 0 ? -r : r;
}

VC_HD double acos_cr(double x)
{
  double y0 = acos(x);
  if (!(x > -1.0 && x < 1.0) || y0 < 1e-6) return y0;
  dd c, s;
  sincos_dd(y0, c, s);
  dd r = add(c, dd{-x, 0.0});
  return y0 + (r.hi + r.lo) / s.hi;
}

}  // namespace vcr
#endif
