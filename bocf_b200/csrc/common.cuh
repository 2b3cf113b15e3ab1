// common.cuh -- shared helpers for the sm_100a kernels of libbocf_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace bocf {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const std::string& msg);
void count_launch(uint64_t n = 1);

#define BOCF_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::bocf::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
      return -2;                                                                             \
    }                                                                                        \
  } while (0)

#define BOCF_LAUNCH_OK(name)                                                                 \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    ::bocf::count_launch();                                                                  \
    if (_e != cudaSuccess) {                                                                 \
      ::bocf::set_error(std::string("launch ") + name + ": " + cudaGetErrorString(_e));      \
      return -2;                                                                             \
    }                                                                                        \
  } while (0)

// Optional per-kernel timing (bocf_profile_enable): CUDA events recorded on the launch stream around a kernel.
struct ProfScope {
  void* rec = nullptr;
  cudaStream_t st;
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
};

inline int64_t round_up(int64_t x, int64_t q) { return (x + q - 1) / q * q; }
inline int64_t ceil_div(int64_t x, int64_t q) { return (x + q - 1) / q; }

// ---- device helpers -------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor-core MMA (SASS: DMMA.8x8x4).
// lane = 4*g + t:  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace bocf
