// Test harness (not a product path): vr::pow_table_eval of vicres_b200/csrc/vic_physics.cuh -- the arithmetic the
// device power function m_pow runs -- compiled for the host and compared with powl() on seeded arguments of the
// model's ranges.  Prints: n, max of relerr / (2 + |y log2 x|), max relerr, number of arguments not handled.
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../../vicres_b200/csrc/vic_physics.cuh"

static const double LOGTAB[2 * 129] = VR_POW_LOGTAB_INIT;
static const double EXPTAB[64] = VR_POW_EXPTAB_INIT;

static double urand(unsigned long long *s)
{
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((*s >> 11) & ((1ULL << 53) - 1)) / 9007199254740992.0;
}

int main(int argc, char **argv)
{
  const long n = argc > 1 ? atol(argv[1]) : 2000000;
  unsigned long long s = 20261020ULL;
  double worst_norm = 0, worst_rel = 0;
  long unhandled = 0;
  for (long i = 0; i < n; i++) {
    double x, y;
    switch (i % 6) {
      case 0: x = urand(&s); y = 4 + 26 * urand(&s); break;                          // Brooks-Corey: relative moisture ^ expt
      case 1: x = 1.0 - 1e-6 * urand(&s); y = 0.5 + 30 * urand(&s); break;           // x -> 1 from below
      case 2: x = 1.0 + 1e-6 * urand(&s); y = -30 + 60 * urand(&s); break;           // x -> 1 from above
      case 3: x = exp(-30 * urand(&s)); y = 3 * urand(&s); break;                    // ARNO curves, tiny bases
      case 4: x = exp(40 * (urand(&s) - 0.5)); y = 20 * (urand(&s) - 0.5); break;    // both sides of 1
      default: x = 0.5 + urand(&s); y = 2.0 / 3.0; break;                            // canopy wetness ^ (2/3), sqrt(2) split
    }
    bool handled;
    const double v = vr::pow_table_eval(x, y, LOGTAB, EXPTAB, &handled);
    if (!handled) { unhandled++; continue; }
    const long double w = powl((long double)x, (long double)y);
    const double rel = (double)fabsl(((long double)v - w) / w);
    const double t = fabs(y * log2(x));
    if (rel > worst_rel) worst_rel = rel;
    if (rel / (2 + t) > worst_norm) worst_norm = rel / (2 + t);
  }
  printf("%ld %.6e %.6e %ld\n", n, worst_norm, worst_rel, unhandled);
  return 0;
}
