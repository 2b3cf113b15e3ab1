// kg.cu -- posterior covariance between candidates and ONE reference point, and its gradient w.r.t. the candidate:
// the building block of the fork's knowledge-gradient helpers (GPy/core/gp.py:493-627):
//
//   cov_j(x, x2)      = k_j(x, x2) - K_j(x, X) W_j^-1 K_j(X, x2)            gp.py:577-599, posterior.py covariance_between_points
//   d cov_j / d x     = gradients_X(None, x, x2) - (K_j(x2, X) W_j^-1) gradients_X(None, x, X)     gp.py:601-627
//
// With beta_j = W_j^-1 K_j(X, x2) (= Linv_j^T Linv_j k2, two triangular matrix-vector products on the resident factor)
// both are the posterior-mean contraction with alpha replaced by beta: cov = k(x, x2) - K*(x) beta and
// d cov = d k(x, x2)/dx - gradients_X(beta^T, x, X).  The conditioned-on-next-point variance of gp.py:518-575 follows on
// the host side from the noiseless variance by the rank-one Schur complement (bocf_b200/model.py).
// Not on the benchmark path (no shipped acquisition calls these): fp64 CUDA cores, thread per candidate.
#include "kernfn.cuh"
#include "model.h"
#include "common.cuh"

namespace bocf {

// squared scaled distance the way the kernel family forms it (se.py:73 exact differences; stationary.py:141-146 expansion)
template <int KIND>
__device__ __forceinline__ double sq_dist(const double* __restrict__ a, const double* __restrict__ b, int d) {
  if (KIND == BOCF_KERN_SE) {
    double r2 = 0.0;
    for (int q = 0; q < d; ++q) {
      const double t = a[q] - b[q];
      r2 = fma(t, t, r2);
    }
    return r2;
  }
  double aa = 0.0, bb = 0.0, ab = 0.0;
  for (int q = 0; q < d; ++q) {
    aa += a[q] * a[q];
    bb += b[q] * b[q];
    ab = fma(a[q], b[q], ab);
  }
  return fmax(-2.0 * ab + (aa + bb), 0.0);
}

// k2[hj][b] = k_j(X_b, x2), zero beyond n
template <int KIND>
__global__ void kvec_kernel(const double* __restrict__ XsAll, const OutHyp* __restrict__ hyp, const double* __restrict__ x2,
                            int n, int n_pad, int d, int m, int h, double* __restrict__ k2, int j0) {
  const int j = j0 + blockIdx.y, hj = h * m + j;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_pad) return;
  const OutHyp& hp = hyp[hj];
  double xs2[MAXD];
  for (int q = 0; q < d; ++q) xs2[q] = x2[q] / hp.ls[q];
  double kv = 0.0, gv;
  if (b < n) kern_eval<KIND, false>(sq_dist<KIND>(XsAll + ((size_t)hj * n_pad + b) * d, xs2, d), hp.variance, kv, gv);
  k2[(size_t)j * n_pad + b] = kv;
}

// t = Linv k2, beta = Linv^T t for the m outputs of hyper-sample h (vectors indexed [j][n_pad])
__global__ void kg_linv_matvec_kernel(const double* __restrict__ Linv, const double* __restrict__ v, int m, int h, int n_pad,
                                      double* __restrict__ out) {
  const int j = blockIdx.y, hj = h * m + j;
  const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n_pad) return;
  const int lane = threadIdx.x & 31;
  const double* row = Linv + (size_t)hj * n_pad * n_pad + (size_t)a * n_pad;
  const double* y = v + (size_t)j * n_pad;
  double s = 0.0;
  for (int b = lane; b <= a; b += 32) s += row[b] * y[b];
  s = warp_sum(s);
  if (lane == 0) out[(size_t)j * n_pad + a] = s;
}
__global__ void kg_linv_t_matvec_kernel(const double* __restrict__ Linv, const double* __restrict__ t, int m, int h, int n_pad,
                                        double* __restrict__ out) {
  const int j = blockIdx.y, hj = h * m + j;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_pad) return;
  const double* Li = Linv + (size_t)hj * n_pad * n_pad;
  const double* tv = t + (size_t)j * n_pad;
  double s = 0.0;
  for (int a = b; a < n_pad; ++a) s += Li[(size_t)a * n_pad + b] * tv[a];
  out[(size_t)j * n_pad + b] = s;
}

// cov[j][i] = k(x_i, x2) - sum_b k(x_i, X_b) beta_b ;  dcov[j][i][q] likewise with the kernel gradients
template <int KIND, bool GRAD>
__global__ void __launch_bounds__(128) cov_point_kernel(const double* __restrict__ Xc, int64_t N, const double* __restrict__ x2,
                                                        const double* __restrict__ XsAll, const OutHyp* __restrict__ hyp,
                                                        const double* __restrict__ beta, int n, int n_pad, int d, int m, int h,
                                                        double* __restrict__ cov, double* __restrict__ dcov, int j0) {
  __shared__ double sX[64][MAXD];
  __shared__ double sb[64];
  const int j = j0 + blockIdx.y, hj = h * m + j;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const OutHyp& hp = hyp[hj];
  double xs[MAXD], xs2[MAXD];
  for (int q = 0; q < d; ++q) {
    xs[q] = (i < N) ? Xc[i * d + q] / hp.ls[q] : 0.0;
    xs2[q] = x2[q] / hp.ls[q];
  }
  double s = 0.0, wsum = 0.0, gm[MAXD];
  for (int q = 0; q < MAXD; ++q) gm[q] = 0.0;
  const double* Xs = XsAll + (size_t)hj * n_pad * d;
  for (int b0 = 0; b0 < n; b0 += 64) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * d; idx += blockDim.x) {
      const int bb = idx / d, q = idx - bb * d;
      sX[bb][q] = (b0 + bb < n) ? Xs[(size_t)(b0 + bb) * d + q] : 0.0;
    }
    if (threadIdx.x < 64) sb[threadIdx.x] = (b0 + threadIdx.x < n) ? beta[(size_t)j * n_pad + b0 + threadIdx.x] : 0.0;
    __syncthreads();
    const int bmax = min(64, n - b0);
    for (int bb = 0; bb < bmax; ++bb) {
      double kv, gv = 0.0;
      kern_eval<KIND, GRAD>(sq_dist<KIND>(xs, sX[bb], d), hp.variance, kv, gv);
      s = fma(kv, sb[bb], s);
      if (GRAD) {
        const double w = gv * sb[bb];
        wsum += w;
        for (int q = 0; q < d; ++q) gm[q] = fma(w, sX[bb][q], gm[q]);
      }
    }
  }
  if (i >= N) return;
  double k12, g12 = 0.0;
  kern_eval<KIND, GRAD>(sq_dist<KIND>(xs, xs2, d), hp.variance, k12, g12);
  cov[(size_t)j * N + i] = k12 - s;
  if (GRAD)
    for (int q = 0; q < d; ++q) {
      // gradients_X(D, x, X)[q] = sum_b D_b g_b (xs_q - Xs_bq) / l_q   (kernfn.cuh)
      const double direct = g12 * (xs[q] - xs2[q]);
      dcov[((size_t)j * N + i) * d + q] = (direct - (xs[q] * wsum - gm[q])) / hp.ls[q];
    }
}

// the two launches that evaluate the kernel family, for one run of outputs
template <int KIND>
static int kvec_t(bocf_model* M, int h, const double* x2, double* k2, OutRun run, cudaStream_t st) {
  kvec_kernel<KIND><<<dim3((unsigned)ceil_div(M->n_pad, 128), (unsigned)run.cnt), 128, 0, st>>>(M->Xs, M->hyp, x2, M->n, M->n_pad,
                                                                                                 M->d, M->m, h, k2, run.j0);
  BOCF_LAUNCH_OK("kvec_kernel");
  return 0;
}
template <int KIND>
static int cov_point_t(bocf_model* M, int h, const double* Xc, int64_t N, const double* x2, const double* beta, double* cov,
                       double* dcov, OutRun run, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)run.cnt);
  if (dcov)
    cov_point_kernel<KIND, true><<<grid, 128, 0, st>>>(Xc, N, x2, M->Xs, M->hyp, beta, M->n, M->n_pad, M->d, M->m, h, cov, dcov, run.j0);
  else
    cov_point_kernel<KIND, false><<<grid, 128, 0, st>>>(Xc, N, x2, M->Xs, M->hyp, beta, M->n, M->n_pad, M->d, M->m, h, cov, dcov, run.j0);
  BOCF_LAUNCH_OK("cov_point_kernel");
  return 0;
}

int launch_cov_point(bocf_model* M, int h, const double* Xc, int64_t N, const double* x2, double* cov, double* dcov,
                     double* work, cudaStream_t st) {
  double* k2 = work;
  double* tv = work + (size_t)M->m * M->n_pad;
  double* beta = work + 2 * (size_t)M->m * M->n_pad;
  int rc = for_each_kind_run(M, [&](int kind, OutRun run) -> int {
    switch (kind) {
      case BOCF_KERN_SE: return kvec_t<BOCF_KERN_SE>(M, h, x2, k2, run, st);
      case BOCF_KERN_RBF: return kvec_t<BOCF_KERN_RBF>(M, h, x2, k2, run, st);
      case BOCF_KERN_MATERN52: return kvec_t<BOCF_KERN_MATERN52>(M, h, x2, k2, run, st);
      case BOCF_KERN_MATERN32: return kvec_t<BOCF_KERN_MATERN32>(M, h, x2, k2, run, st);
    }
    set_error("unknown kernel kind");
    return BOCF_ERR_INVALID;
  });
  if (rc) return rc;
  kg_linv_matvec_kernel<<<dim3((unsigned)ceil_div(M->n_pad, 8), (unsigned)M->m), 256, 0, st>>>(M->Linv, k2, M->m, h, M->n_pad, tv);
  BOCF_LAUNCH_OK("kg_linv_matvec_kernel");
  kg_linv_t_matvec_kernel<<<dim3((unsigned)ceil_div(M->n_pad, 128), (unsigned)M->m), 128, 0, st>>>(M->Linv, tv, M->m, h, M->n_pad, beta);
  BOCF_LAUNCH_OK("kg_linv_t_matvec_kernel");
  return for_each_kind_run(M, [&](int kind, OutRun run) -> int {
    switch (kind) {
      case BOCF_KERN_SE: return cov_point_t<BOCF_KERN_SE>(M, h, Xc, N, x2, beta, cov, dcov, run, st);
      case BOCF_KERN_RBF: return cov_point_t<BOCF_KERN_RBF>(M, h, Xc, N, x2, beta, cov, dcov, run, st);
      case BOCF_KERN_MATERN52: return cov_point_t<BOCF_KERN_MATERN52>(M, h, Xc, N, x2, beta, cov, dcov, run, st);
      case BOCF_KERN_MATERN32: return cov_point_t<BOCF_KERN_MATERN32>(M, h, Xc, N, x2, beta, cov, dcov, run, st);
    }
    set_error("unknown kernel kind");
    return BOCF_ERR_INVALID;
  });
}

}  // namespace bocf
