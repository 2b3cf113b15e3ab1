// kernfn.cuh -- covariance as a function of the (clipped) squared scaled distance, fp64.
//
//   k = K_of_r(r),   g = dK/dr(r) * (1/r)  with 1/r := 0 at r == 0   (Stationary._inv_dist, stationary.py:227-234)
// so that  gradients_X(D, x*, X)[q] = (1/l_q) * sum_b D_b * g_b * (xs*_q - Xs_bq)   in lengthscale-scaled inputs
// (stationary.py:332-342 + stationary_utils.c:1-14).  For SE the same form holds with g = -k (se.py:139-147).
//
// The reference evaluates dK/dr and then multiplies by 1/r; here the quotient is simplified algebraically
// (no division, no extra rounding step):
//   RBF       dK/dr / r = -k                                   (rbf.py:45-46:      dK/dr = -r k)
//   Matern52  dK/dr / r = -(5/3) s^2 (1 + sqrt5 r) e^{-sqrt5 r} (stationary.py:532-533: (10/3 r - 5 r - 5 sqrt5/3 r^2) e)
//   Matern32  dK/dr / r = -3 s^2 e^{-sqrt3 r}                   (stationary.py:443-444: -3 s^2 r e)
// with the reference's convention that the weight is exactly 0 where r == 0.  Differences are at the 1e-16 level.
//
// exp / sqrt: the arguments on this path are exp(x <= 0) and sqrt(r2 >= 0).  The generic libdevice routines spend
// most of their instructions on special cases (ncu: the K* kernel was issue-bound on non-fp64 instructions with the
// fp64 pipe 50 % busy), so branch-free versions restricted to these ranges are used: < 2 ulp from the libdevice
// results, far inside the 1e-6 / 1e-9 parity bars.
#pragma once
#include "model.h"

namespace bocf {

// exp(x) for x <= 0.  Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor/Horner (remainder < 4e-18),
// scaling by 2^k through the exponent field.  x < -700 (result < 1e-304) flushes to 0.
// Coefficients live in constant memory so every DFMA takes them as a constant-bank operand: as immediates each 64-bit
// literal costs two extra move instructions per use (the K* kernel was issue-bound on exactly those moves).
static __constant__ double EXPC[16] = {
    1.6059043836821613e-10,      // 1/13!
    2.08767569878681e-09,        // 1/12!
    2.505210838544172e-08,       // 1/11!
    2.755731922398589e-07,       // 1/10!
    2.7557319223985893e-06,      // 1/9!
    2.48015873015873e-05,        // 1/8!
    1.984126984126984e-04,       // 1/7!
    1.3888888888888889e-03,      // 1/6!
    8.333333333333333e-03,       // 1/5!
    4.1666666666666664e-02,      // 1/4!
    1.6666666666666666e-01,      // 1/3!
    1.4426950408889634074,       // [11] log2(e)
    6755399441055744.0,          // [12] 1.5 * 2^52
    6.93147180369123816490e-01,  // [13] ln2 high
    1.90821492927058770002e-10,  // [14] ln2 low
    -700.0};
__device__ __forceinline__ double exp_nonpos(double x) {
  const double xc = fmax(x, EXPC[15]);
  // round-to-nearest integer of xc * log2(e) through the 1.5 * 2^52 shift: no FRND / F2I conversion instructions,
  // and the integer is available in the low word of the shifted value
  const double sh = fma(xc, EXPC[11], EXPC[12]);
  const double kf = sh - EXPC[12];
  double r = fma(-kf, EXPC[13], xc);
  r = fma(-kf, EXPC[14], r);
  double p = EXPC[0];
#pragma unroll
  for (int c = 1; c <= 10; ++c) p = fma(p, r, EXPC[c]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int k = __double2loint(sh);
  const double scaled = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  return (x < EXPC[15]) ? 0.0 : scaled;
}

// sqrt(a) for a >= 0 (finite): one Newton step on a * rsqrt(a).
__device__ __forceinline__ double sqrt_nonneg(double a) {
  const double y = rsqrt(a);
  double s = a * y;
  const double e = fma(-s, s, a);
  s = fma(e, 0.5 * y, s);
  return (a > 0.0) ? s : 0.0;
}

template <int KIND, bool GRAD>
__device__ __forceinline__ void kern_eval(double r2, double variance, double& k, double& g) {
  if (KIND == BOCF_KERN_SE) {
    // se.py:60  variance * exp(-0.5 * sqdist)
    k = variance * exp_nonpos(-0.5 * r2);
    if (GRAD) g = -k;
  } else if (KIND == BOCF_KERN_RBF) {
    // rbf.py:42-46 (r*r of the rounded sqrt differs from r2 by <= 1 ulp)
    k = variance * exp_nonpos(-0.5 * r2);
    if (GRAD) g = (r2 != 0.0) ? -k : 0.0;
  } else if (KIND == BOCF_KERN_MATERN52) {
    // stationary.py:529-533
    const double r = sqrt_nonneg(r2);
    const double s5 = 2.23606797749978969641;   // sqrt(5)
    const double e = variance * exp_nonpos(-s5 * r);
    const double lin = 1.0 + s5 * r;
    k = (lin + 5.0 / 3.0 * r2) * e;
    if (GRAD) g = (r2 != 0.0) ? (-5.0 / 3.0) * (lin * e) : 0.0;
  } else {
    // Matern32, stationary.py:440-444
    const double r = sqrt_nonneg(r2);
    const double s3 = 1.73205080756887729353;   // sqrt(3)
    const double e = variance * exp_nonpos(-s3 * r);
    k = (1.0 + s3 * r) * e;
    if (GRAD) g = (r2 != 0.0) ? -3.0 * e : 0.0;
  }
}

}  // namespace bocf
