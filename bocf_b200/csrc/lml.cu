// lml.cu -- log marginal likelihood of every (hyper-sample, output) GP and its gradient with respect to the kernel
// variance, the ARD lengthscales and the Gaussian noise variance, on device in fp64 (the
// objective / gradient pair that ML-II and HMC evaluate thousands of times per BO iteration).
//
//   log p(y)  = -1/2 (n log 2pi + log|Ky| + (y-ybar)^T alpha)          exact_gaussian_inference.py:53
//   dL_dK     =  1/2 (alpha alpha^T - Ky^-1)                            exact_gaussian_inference.py:61
//   d/d var   =  sum K o dL_dK / var                                    stationary.py:197, se.py:181
//   d/d l_q   = -sum dL_dK (dK/dr / r) (x_aq - x_bq)^2 / l_q^3          stationary.py:203-240 + stationary_utils.c:34-48,
//                                                                       se.py:183 (same expression with dK/dr / r = -K)
//   d/d noise =  tr dL_dK                                               gaussian.py:64-71
//
// Ky^-1 = Linv^T Linv is never stored: one CTA per 128 x 128 tile of the lower triangle contracts the two column
// blocks of Linv on the fp64 tensor-core tile engine (k >= the tile's row block only -- Linv is lower triangular) and
// reduces its tile of dL_dK against K, dK/dr/r and the squared coordinate differences in the epilogue.  Tiles are
// summed in a fixed order by a second kernel (no atomics: results are reproducible bit for bit).
#include "gemm_f64.cuh"
#include "kernfn.cuh"
#include "model.h"

namespace bocf {

using LT = gemm::Tile128;

// DP: input dimension rounded up (compile time), so the per-thread lengthscale-gradient sums stay in registers and the
// loops over the dimensions unroll; the staged inputs use an ODD row stride (DP + 1): rows 8 apart in a warp otherwise hit
// one bank (stride 16 doubles: 8-way conflicts on every read of the epilogue -- it cost more than the contraction).
template <int KIND, int DP>
__global__ void __launch_bounds__(LT::NTHREADS, 1) lml_tile_kernel(const double* __restrict__ LinvAll,
                                                                    const double* __restrict__ alphaAll,
                                                                    const double* __restrict__ XsAll,
                                                                    const double* __restrict__ xsqAll,
                                                                    const OutHyp* __restrict__ hyp, int n, int n_pad, int d,
                                                                    int ntiles, double* __restrict__ part, OutRun run,
                                                                    int m) {
  extern __shared__ __align__(16) double smem[];
  const int hj = run_hj(blockIdx.y, run, m);
  int p = blockIdx.x;                                   // lower-triangular tile pair (I >= J)
  int I = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while ((I + 1) * (I + 2) / 2 <= p) ++I;
  while (I * (I + 1) / 2 > p) --I;
  const int J = p - I * (I + 1) / 2;
  const double* Linv = LinvAll + (int64_t)hj * n_pad * n_pad;
  // Wi[a][b] = sum_k Linv[k][a] Linv[k][b]:  A(m = a, k) = Linv[k*ld + a], B(k, n = b) = Linv[k*ld + b]
  double acc[8][4][2];
  gemm::zero_acc(acc);
  gemm::mainloop<LT, true, true>(acc, Linv + I * TILE, n_pad, Linv + J * TILE, n_pad, I * TILE, n_pad, smem);

  // stage the scaled inputs / alpha of both blocks (the pipeline buffers are free now)
  constexpr int SXL = DP + 1;
  double* sxa = smem;                         // TILE x SXL, zero beyond d
  double* sxb = sxa + TILE * SXL;             // TILE x SXL
  double* sal = sxb + TILE * SXL;             // 2 x TILE alpha
  double* ssq = sal + 2 * TILE;               // 2 x TILE |xs|^2
  double* red = ssq + 2 * TILE;               // warps x (MAXD + 2)
  const int tid = threadIdx.x;
  const double* Xs = XsAll + (int64_t)hj * n_pad * d;
  for (int idx = tid; idx < TILE * DP; idx += LT::NTHREADS) {
    const int r = idx / DP, q = idx - r * DP;
    sxa[r * SXL + q] = (q < d) ? Xs[(int64_t)(I * TILE + r) * d + q] : 0.0;
    sxb[r * SXL + q] = (q < d) ? Xs[(int64_t)(J * TILE + r) * d + q] : 0.0;
  }
  for (int r = tid; r < TILE; r += LT::NTHREADS) {
    sal[r] = alphaAll[(int64_t)hj * n_pad + I * TILE + r];
    sal[TILE + r] = alphaAll[(int64_t)hj * n_pad + J * TILE + r];
    ssq[r] = xsqAll[(int64_t)hj * n_pad + I * TILE + r];
    ssq[TILE + r] = xsqAll[(int64_t)hj * n_pad + J * TILE + r];
  }
  __syncthreads();

  const OutHyp& hp = hyp[hj];
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int mbase = (warp >> 2) * 64, nbase = (warp & 3) * 32;
  double gvar = 0.0, gnoise = 0.0, gl[DP];
#pragma unroll
  for (int q = 0; q < DP; ++q) gl[q] = 0.0;
  // the two adjacent columns a thread owns in each 8 x 8 accumulator tile are evaluated together, branch-free inside the
  // pair (two independent dependency chains through the kernel evaluation); pairs are separated by a real branch, which
  // also keeps the scheduler from hoisting all 64 elements' loads at once (that version spilled 3 KB per thread)
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int r = mbase + 8 * i + g, c0 = nbase + 8 * jn + 2 * t;
      const int a = I * TILE + r, b0 = J * TILE + c0;
      if (a >= n || b0 >= n || b0 > a) continue;               // lower triangle of the n x n problem only
      const double* xa = sxa + r * SXL;
      double r2[2], wd[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + e, b = b0 + e;
        const double* xb = sxb + c * SXL;
        const double w = (b >= n || b > a) ? 0.0 : (a == b) ? 1.0 : 2.0;      // symmetric: count (a,b) and (b,a)
        wd[e] = w * 0.5 * (sal[r] * sal[TILE + c] - acc[i][jn][e]);            // w dL/dK
        if (KIND == BOCF_KERN_SE) {
          double x = 0.0;
#pragma unroll
          for (int q = 0; q < DP; ++q) {
            const double df = xa[q] - xb[q];
            x += df * df;
          }
          r2[e] = x;
        } else {
          double dot = 0.0;
#pragma unroll
          for (int q = 0; q < DP; ++q) dot += xa[q] * xb[q];
          r2[e] = fmax(-2.0 * dot + (ssq[r] + ssq[TILE + c]), 0.0);
        }
        if (a == b) r2[e] = 0.0;                                 // stationary.py:136, se.py:58
      }
      double kv[2], gv[2];
      kern_eval<KIND, true>(r2[0], hp.variance, kv[0], gv[0]);
      kern_eval<KIND, true>(r2[1], hp.variance, kv[1], gv[1]);
      gvar += wd[0] * kv[0] + wd[1] * kv[1];
      if (a == b0) gnoise += wd[0];                              // w == 1 on the diagonal
      if (a == b0 + 1) gnoise += wd[1];
      const double wg0 = wd[0] * gv[0], wg1 = wd[1] * gv[1];
      const double* xb0 = sxb + c0 * SXL;
#pragma unroll
      for (int q = 0; q < DP; ++q) {
        const double d0 = xa[q] - xb0[q], d1 = xa[q] - xb0[SXL + q];
        gl[q] += wg0 * d0 * d0 + wg1 * d1 * d1;
      }
    }
  // block reduction of d + 2 values
  gvar = warp_sum(gvar);
  gnoise = warp_sum(gnoise);
#pragma unroll
  for (int q = 0; q < DP; ++q) gl[q] = warp_sum(gl[q]);
  __syncthreads();
  if (lane == 0) {
    red[warp * (MAXD + 2) + 0] = gvar;
    red[warp * (MAXD + 2) + 1] = gnoise;
#pragma unroll
    for (int q = 0; q < DP; ++q) red[warp * (MAXD + 2) + 2 + q] = gl[q];
  }
  __syncthreads();
  if (tid < d + 2) {
    double s = 0.0;
    for (int w8 = 0; w8 < LT::NTHREADS / 32; ++w8) s += red[w8 * (MAXD + 2) + tid];
    part[((int64_t)hj * ntiles + p) * (MAXD + 2) + tid] = s;
  }
}

// sums the tile partials in order, applies the 1/var and -1/l scalings, and forms log p(y)
__global__ void lml_finish_kernel(const double* __restrict__ part, const double* __restrict__ Lmat,
                                  const double* __restrict__ alphaAll, const double* __restrict__ yc,
                                  const OutHyp* __restrict__ hyp, int n, int n_pad, int d, int m, int ntiles,
                                  double* __restrict__ out, OutRun grp) {
  __shared__ double s_ld[32], s_fit[32];
  const int hj = run_hj(blockIdx.x, grp, m), j = hj % m;
  const int tid = threadIdx.x;
  double ld = 0.0, fit = 0.0;
  const double* L = Lmat + (int64_t)hj * n_pad * n_pad;
  for (int a = tid; a < n; a += blockDim.x) {
    ld += log(L[(int64_t)a * n_pad + a]);
    fit += alphaAll[(int64_t)hj * n_pad + a] * yc[(int64_t)j * n_pad + a];
  }
  ld = warp_sum(ld);
  fit = warp_sum(fit);
  if ((tid & 31) == 0) {
    s_ld[tid >> 5] = ld;
    s_fit[tid >> 5] = fit;
  }
  __syncthreads();
  const OutHyp& hp = hyp[hj];
  double* o = out + (int64_t)hj * (MAXD + 3);          // [lml, g_var, g_noise, g_len[0..d)]
  if (tid == 0) {
    double lds = 0.0, fits = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      lds += s_ld[w];
      fits += s_fit[w];
    }
    o[0] = 0.5 * (-(double)n * 1.8378770664093453 - 2.0 * lds - fits);     // log(2 pi)
  }
  if (tid < d + 2) {
    double s = 0.0;
    for (int p = 0; p < ntiles; ++p) s += part[((int64_t)hj * ntiles + p) * (MAXD + 2) + tid];
    if (tid == 0) o[1] = s / hp.variance;
    else if (tid == 1) o[2] = s;
    else o[3 + (tid - 2)] = -s / hp.ls[tid - 2];
  }
}

template <int KIND, int DP>
static int launch_tiles_d(bocf_model* M, int ntiles, double* part, OutRun run, cudaStream_t st) {
  static bool done[64] = {false};              // per device: function attributes belong to the device's context
  if (M->device >= 0 && M->device < 64 && !done[M->device]) {
    BOCF_CUDA_OK(cudaFuncSetAttribute(lml_tile_kernel<KIND, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT::SMEM_BYTES));
    done[M->device] = true;
  }
  lml_tile_kernel<KIND, DP><<<dim3((unsigned)ntiles, (unsigned)(M->H * run.cnt)), LT::NTHREADS, LT::SMEM_BYTES, st>>>(
      M->Linv, M->alpha, M->Xs, M->xsq, M->hyp, M->n, M->n_pad, M->d, ntiles, part, run, M->m);
  BOCF_LAUNCH_OK("lml_tile_kernel");
  return 0;
}
template <int KIND>
static int launch_tiles(bocf_model* M, int ntiles, double* part, OutRun run, cudaStream_t st) {
  const int d = M->d;
  if (d <= 4) return launch_tiles_d<KIND, 4>(M, ntiles, part, run, st);
  if (d <= 6) return launch_tiles_d<KIND, 6>(M, ntiles, part, run, st);
  if (d <= 10) return launch_tiles_d<KIND, 10>(M, ntiles, part, run, st);
  return launch_tiles_d<KIND, MAXD>(M, ntiles, part, run, st);
}

// out_host: H*m x (MAXD + 3) doubles [lml, d/dvariance, d/dnoise, d/dlengthscale[0..d)]
int launch_log_likelihood(bocf_model* M, double* out_host, cudaStream_t st) {
  static_assert(2 * TILE * (MAXD + 1) + 4 * TILE + 8 * (MAXD + 2) <= LT::SMEM_BYTES / (int)sizeof(double), "epilogue staging fits");
  const int Hm = M->H * M->m;
  const int ntiles = M->nb * (M->nb + 1) / 2;
  const size_t n_part = (size_t)Hm * ntiles * (MAXD + 2), n_out = (size_t)Hm * (MAXD + 3);
  if (M->lml_ws_count < n_part + n_out) {
    if (M->lml_ws) cudaFree(M->lml_ws);
    M->lml_ws = nullptr;
    M->lml_ws_count = 0;
    if (cudaMalloc(reinterpret_cast<void**>(&M->lml_ws), sizeof(double) * (n_part + n_out)) != cudaSuccess) {
      set_error("bocf_model_log_likelihood: out of device memory");
      return BOCF_ERR_CUDA;
    }
    M->lml_ws_count = n_part + n_out;
  }
  double* part = M->lml_ws;
  double* out = M->lml_ws + n_part;
  struct Ctx {
    int ntiles;
    double *part, *out;
  } ctx{ntiles, part, out};
  int rc = for_each_output_group(
      M, st,
      [](bocf_model* Mm, OutRun grp, cudaStream_t s, void* vc) -> int {
        const Ctx& c = *static_cast<const Ctx*>(vc);
        int r = for_each_kind_run(Mm, grp, [&](int kind, OutRun run) -> int {
          switch (kind) {
            case BOCF_KERN_SE: return launch_tiles<BOCF_KERN_SE>(Mm, c.ntiles, c.part, run, s);
            case BOCF_KERN_RBF: return launch_tiles<BOCF_KERN_RBF>(Mm, c.ntiles, c.part, run, s);
            case BOCF_KERN_MATERN52: return launch_tiles<BOCF_KERN_MATERN52>(Mm, c.ntiles, c.part, run, s);
            default: return launch_tiles<BOCF_KERN_MATERN32>(Mm, c.ntiles, c.part, run, s);
          }
        });
        if (r) return r;
        lml_finish_kernel<<<Mm->H * grp.cnt, 256, 0, s>>>(c.part, Mm->Lmat, Mm->alpha, Mm->yc, Mm->hyp, Mm->n, Mm->n_pad, Mm->d,
                                                        Mm->m, c.ntiles, c.out, grp);
        BOCF_LAUNCH_OK("lml_finish_kernel");
        return 0;
      },
      &ctx);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(out_host, out, sizeof(double) * Hm * (MAXD + 3), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error(std::string("bocf_model_log_likelihood: ") + cudaGetErrorString(e));
      rc = BOCF_ERR_CUDA;
    }
  }
  return rc;
}

}  // namespace bocf
