// kernfn.cuh -- covariance as a function of the (clipped) squared scaled distance, fp64.
//
//   k = K_of_r(r),   g = dK/dr(r) * (1/r)  with 1/r := 0 at r == 0   (Stationary._inv_dist, stationary.py:227-234)
// so that  gradients_X(D, x*, X)[q] = (1/l_q) * sum_b D_b * g_b * (xs*_q - Xs_bq)   in lengthscale-scaled inputs
// (stationary.py:332-342 + stationary_utils.c:1-14).  For SE the same form holds with g = -k (se.py:139-147).
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
    // rbf.py:42-46
    const double r = sqrt(r2);
    k = variance * exp(-0.5 * (r * r));
    if (GRAD) {
      const double dkdr = -r * k;
      const double invr = (r != 0.0) ? 1.0 / r : 0.0;
      g = invr * dkdr;
    }
  } else if (KIND == BOCF_KERN_MATERN52) {
    // stationary.py:529-533
    const double r = sqrt(r2);
    const double s5 = 2.23606797749978969641;   // sqrt(5)
    const double e = exp(-s5 * r);
    k = variance * (1.0 + s5 * r + 5.0 / 3.0 * (r * r)) * e;
    if (GRAD) {
      const double dkdr = variance * (10.0 / 3.0 * r - 5.0 * r - 5.0 * s5 / 3.0 * (r * r)) * e;
      const double invr = (r != 0.0) ? 1.0 / r : 0.0;
      g = invr * dkdr;
    }
  } else {
    // Matern32, stationary.py:440-444
    const double r = sqrt(r2);
    const double s3 = 1.73205080756887729353;   // sqrt(3)
    const double e = exp(-s3 * r);
    k = variance * (1.0 + s3 * r) * e;
    if (GRAD) {
      const double dkdr = -3.0 * variance * r * e;
      const double invr = (r != 0.0) ? 1.0 / r : 0.0;
      g = invr * dkdr;
    }
  }
}

}  // namespace bocf
