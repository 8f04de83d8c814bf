// sin() and cos() that agree bit for bit with the host C library the reference runs on.
//
// Why: World::Raytrace takes sin/cos of every heading from libm (world.cc:774-775) and is chaotic in
// their last bit for rays that cross a world axis through empty space, so
// "identical results" needs identical trig. The host libm here is glibc 2.39, whose x86-64 ifunc
// picks the FMA build of sysdeps/ieee754/dbl-64/s_sin.c on every FMA-capable CPU. This is that
// algorithm with the fused multiply-adds exactly where that build has them:
//   |x| < 2^-26 (sin) / 2^-27 (cos)  -> x / 1
//   |x| < 0.855469                   -> Taylor (|x| < 0.126) or table-based do_sin / do_cos
//   |x| < 2.426265                   -> pi/2 - |x| with the two-double pi/2 (hp0 + hp1)
//   |x| < 105414350                  -> Cody-Waite reduction by pi/2 in four parts, then as above
//   larger / non-finite              -> not covered here (callers fall back to the CUDA library)
// The table of sin/cos at k/128 is the host libm's own (rt_sincos_tab.h, tools/gen_sincos_tab.py).
// Every a*b+c below that must be fused is written fma(); everything else must NOT be contracted:
// compile device code with -fmad=false and host code with -ffp-contract=off.
// tests/test_libm_port.py checks the host build against libm on tens of millions of arguments, and
// tests/test_gpu_trig.py checks the device build against libm on the GPU box.
// Derived from glibc (LGPL-2.1-or-later): see NOTICE.md at the repository root.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "rt_sincos_tab.h"

#if defined(__CUDACC__)
#define STG_HD __host__ __device__ __forceinline__
#else
#define STG_HD inline
#endif

namespace stg {
namespace libm {

#if defined(__CUDACC__)
static __device__ const double d_sincos_tab[440] = { STG_SINCOS_TAB_VALUES };
#endif
static const double h_sincos_tab[440] = { STG_SINCOS_TAB_VALUES };

STG_HD double tab(int i)
{
#if defined(__CUDA_ARCH__)
  return d_sincos_tab[i];
#else
  return h_sincos_tab[i];
#endif
}

STG_HD uint64_t bits_of(double x)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  return b;
#endif
}

STG_HD double from_bits(uint64_t b)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}

// s_sin.c / usncs.h constants
#define STG_LIBM_K(name, hex) static constexpr double name = hex
STG_LIBM_K(s1, -0x1.5555555555555p-3);
STG_LIBM_K(s2, 0x1.1111111110ecep-7);
STG_LIBM_K(s3, -0x1.a01a019db08b8p-13);
STG_LIBM_K(s4, 0x1.71de27b9a7ed9p-19);
STG_LIBM_K(s5, -0x1.addffc2fcdf59p-26);
STG_LIBM_K(sn3, -0x1.5555555555515p-3);
STG_LIBM_K(sn5, 0x1.11110e829872fp-7);
STG_LIBM_K(cs2, 0.5);
STG_LIBM_K(cs4, -0x1.5555555555535p-5);
STG_LIBM_K(cs6, 0x1.6c16bedd9e239p-10);
STG_LIBM_K(big, 0x1.8p+45);
STG_LIBM_K(toint, 0x1.8p+52);
STG_LIBM_K(hpinv, 0x1.45f306dc9c883p-1);
STG_LIBM_K(hp0, 0x1.921fb54442d18p+0);
STG_LIBM_K(hp1, 0x1.1a62633145c07p-54);
STG_LIBM_K(mp1, 0x1.921fb58p+0);
STG_LIBM_K(mp2, -0x1.dde973cp-27);
STG_LIBM_K(pp3, -0x1.cb3b398p-55);
STG_LIBM_K(pp4, -0x1.d747f23e32ed7p-83);
#undef STG_LIBM_K

// TAYLOR_SIN(xx, a, da), |a| < 0.126
STG_HD double taylor_sin(double a, double da)
{
  const double xx = a * a;
  double t = fma(xx, s5, s4);
  t = fma(xx, t, s3);
  t = fma(xx, t, s2);
  t = fma(xx, t, s1);
  const double h = da * 0.5;
  t = fma(t, a, -h);
  const double r = fma(xx, t, da);
  return a + r;
}

// do_sin(x, dx), |x| >= 0.126; result carries the sign of x
STG_HD double do_sin_tab(double a, double da)
{
  if (a <= 0)
    da = -da;
  const double ax = fabs(a);
  const double u = big + ax;
  const double xr = ax - (u - big);
  const int k = (int)((uint32_t)bits_of(u) << 2);
  const double xx = xr * xr;
  const double t = fma(xx, sn5, sn3);
  const double p = xr * xx;
  const double e = fma(p, t, da);
  const double s = xr + e;
  double c6 = fma(xx, cs6, cs4);
  c6 = fma(xx, c6, cs2);
  const double q = xx * c6;
  const double c = fma(xr, da, q);
  const double sn = tab(k), ssn = tab(k + 1), cs = tab(k + 2), ccs = tab(k + 3);
  double w = fma(s, ccs, ssn);
  w = fma(-c, sn, w);
  const double cor = fma(s, cs, w);
  return copysign(sn + cor, a);
}

// do_cos(x, dx)
STG_HD double do_cos_tab(double a, double da)
{
  if (a < 0)
    da = -da;
  const double ax = fabs(a);
  const double u = big + ax;
  double xr = ax - (u - big);
  xr = xr + da;
  const int k = (int)((uint32_t)bits_of(u) << 2);
  const double xx = xr * xr;
  const double t = fma(xx, sn5, sn3);
  const double p = xr * xx;
  const double s = fma(p, t, xr);
  double c6 = fma(xx, cs6, cs4);
  c6 = fma(xx, c6, cs2);
  const double c = xx * c6;
  const double sn = tab(k), ssn = tab(k + 1), cs = tab(k + 2), ccs = tab(k + 3);
  double w = fma(-s, ssn, ccs);
  w = fma(-c, cs, w);
  const double cor = fma(-s, sn, w);
  return cs + cor;
}

// reduce_sincos: x = n*pi/2 + (a + da)
STG_HD int reduce_sincos(double x, double &a, double &da)
{
  const double t = fma(x, hpinv, toint);
  const double xn = t - toint;
  const int n = (int)((uint32_t)bits_of(t) & 3u);
  double y = fma(-xn, mp1, x);
  y = fma(-xn, mp2, y);
  const double t2 = fma(-xn, pp3, y);
  double db = y - t2;
  db = fma(-xn, pp3, db);
  const double b = fma(-xn, pp4, t2);
  double e = t2 - b;
  e = fma(-xn, pp4, e);
  a = b;
  da = db + e;
  return n;
}

STG_HD double do_sincos(double a, double da, int n)
{
  double r;
  if (n & 1)
    r = do_cos_tab(a, da);
  else
    r = fabs(a) < 0.126 ? taylor_sin(a, da) : do_sin_tab(a, da);
  return (n & 2) ? -r : r;
}

// true iff |x| is in the range this port covers (everything a heading can be)
STG_HD bool covered(double x)
{
  return (uint32_t)(bits_of(x) >> 32 & 0x7fffffffu) < 0x419921fbu;
}

STG_HD double sin_host_exact(double x)
{
  const uint32_t k = (uint32_t)(bits_of(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u)
    return x;
  if (k < 0x3feb6000u)
    return fabs(x) < 0.126 ? taylor_sin(x, 0.0) : do_sin_tab(x, 0.0);
  if (k < 0x400368fdu) {
    const double t = hp0 - fabs(x);
    return copysign(do_cos_tab(t, hp1), x);
  }
  double a, da;
  const int n = reduce_sincos(x, a, da);
  return do_sincos(a, da, n);
}

STG_HD double cos_host_exact(double x)
{
  const uint32_t k = (uint32_t)(bits_of(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u)
    return 1.0;
  if (k < 0x3feb6000u)
    return do_cos_tab(x, 0.0);
  if (k < 0x400368fdu) {
    const double y = hp0 - fabs(x);
    const double a = y + hp1;
    const double da = (y - a) + hp1;
    return fabs(a) < 0.126 ? taylor_sin(a, da) : do_sin_tab(a, da);
  }
  double a, da;
  const int n = reduce_sincos(x, a, da);
  return do_sincos(a, da, n + 1);
}

} // namespace libm
} // namespace stg
