"""Generate the double-double constants of vicres_b200/csrc/vic_cr_trig.cuh (sin/cos of k*pi/16 and pi/16 itself)
with mpmath at 120 digits.  Prints C++ initialisers."""
import mpmath as mp

mp.mp.dps = 120


def dd(x):
    hi = float(x)
    lo = float(x - mp.mpf(hi))
    return hi, lo


def fmt(v):
    return "%s, %s" % (float.hex(v[0]), float.hex(v[1]))


print("// pi/16 as three doubles")
p = mp.pi / 16
h = float(p); r = p - mp.mpf(h); m = float(r); l = float(r - mp.mpf(m))
print("constexpr double PI16_H = %s, PI16_M = %s, PI16_L = %s;" % (float.hex(h), float.hex(m), float.hex(l)))
print("// {cos hi, cos lo, sin hi, sin lo} of k*pi/16, k = 0..16")
for k in range(17):
    c, s = dd(mp.cos(k * p)), dd(mp.sin(k * p))
    print("  {%s, %s}," % (fmt(c), fmt(s)))
print("// 1/n! as double-double, n = 2..19")
for n in range(2, 20):
    print("  {%s}," % fmt(dd(1 / mp.factorial(n))))
