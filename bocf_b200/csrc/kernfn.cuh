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
#pragma once
#include "model.h"

namespace bocf {

template <int KIND, bool GRAD>
__device__ __forceinline__ void kern_eval(double r2, double variance, double& k, double& g) {
  if (KIND == BOCF_KERN_SE) {
    // se.py:60  variance * exp(-0.5 * sqdist)
    k = variance * exp(-0.5 * r2);
    if (GRAD) g = -k;
  } else if (KIND == BOCF_KERN_RBF) {
    // rbf.py:42-46 (r*r of the rounded sqrt differs from r2 by <= 1 ulp)
    k = variance * exp(-0.5 * r2);
    if (GRAD) g = (r2 != 0.0) ? -k : 0.0;
  } else if (KIND == BOCF_KERN_MATERN52) {
    // stationary.py:529-533
    const double r = sqrt(r2);
    const double s5 = 2.23606797749978969641;   // sqrt(5)
    const double e = variance * exp(-s5 * r);
    const double lin = 1.0 + s5 * r;
    k = (lin + 5.0 / 3.0 * r2) * e;
    if (GRAD) g = (r2 != 0.0) ? (-5.0 / 3.0) * (lin * e) : 0.0;
  } else {
    // Matern32, stationary.py:440-444
    const double r = sqrt(r2);
    const double s3 = 1.73205080756887729353;   // sqrt(3)
    const double e = variance * exp(-s3 * r);
    k = (1.0 + s3 * r) * e;
    if (GRAD) g = (r2 != 0.0) ? -3.0 * e : 0.0;
  }
}

}  // namespace bocf
