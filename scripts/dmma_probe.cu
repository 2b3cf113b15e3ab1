// dmma_probe.cu -- does fp64 tensor-core work (DMMA.8x8x4) run beside the fp64 vector pipe, or instead of it?
// K* (posterior.cu) saturates the fp64 pipe with ~44 DFMA-class instructions per (candidate, training point) pair, 10 of
// them the distance dot product.  On fragments the same dot products cost 24 DMMAs per 512 pairs = 1.5 per warp
// iteration.  Three loops per warp iteration: A = 34 DFMA, B = 44 DFMA (today), C = 34 DFMA + 1.5 DMMA (the proposal).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dmma_probe scripts/dmma_probe.cu && scripts/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int DM2>   // NF DFMAs and DM2 / 2 DMMAs per iteration
__global__ void __launch_bounds__(128) probe(double* out, int iters, double seed) {
  double x[8], c0 = 0.0, c1 = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = seed + threadIdx.x * 1e-3 + k;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; it += 2) {
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
      for (int f = 0; f < NF; ++f) x[f & 7] = fma(x[f & 7], a, b);     // 8 independent chains
    }
#pragma unroll
    for (int q = 0; q < DM2; ++q) dmma884(c0, c1, x[q & 7], x[(q + 3) & 7]);
  }
  double s = c0 + c1;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int DM2>
static float run(const char* name, double* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * 8;                   // 8 CTAs of 4 warps per SM: issue-bound, like K* is not (it runs 1-2 CTAs)
  probe<NF, DM2><<<grid, 128>>>(out, iters, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) probe<NF, DM2><<<grid, 128>>>(out, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5.f;
  const double warp_iters = (double)grid * 4 * iters;
  printf("%-28s %8.3f ms   %.3f ns per warp iteration per SM-slot   (%d DFMA + %.1f DMMA per iteration)\n", name, ms,
         1e6 * ms / (warp_iters / 148.0), NF, DM2 / 2.0);
  return ms;
}

int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 128);
  const int iters = 200000;
  const float a = run<34, 0>("A: 34 DFMA", out, iters);
  const float b = run<44, 0>("B: 44 DFMA (today)", out, iters);
  const float c = run<34, 3>("C: 34 DFMA + 1.5 DMMA", out, iters);
  const float d = run<0, 3>("D: 1.5 DMMA alone", out, iters);
  const float e = run<34, 6>("E: 34 DFMA + 3 DMMA", out, iters);
  printf("B / A = %.3f   C / A = %.3f   C / B = %.3f   D / A = %.3f   E / A = %.3f\n", b / a, c / a, c / b, d / a, e / a);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
  return 0;
}
