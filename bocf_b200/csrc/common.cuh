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

// Byte transpose for the digit planes: dg[e] holds the balanced base-256 digits of element e in its bytes 0..S-1;
// out[t] packs digit t of the four elements (element e in byte e).  Two-stage PRMT butterfly: 8 permutes for digits
// 0-3 and 3 more per digit above, instead of 4 shift/mask/or operations per byte.
template <int S>
__device__ __forceinline__ void digits_transpose4(const unsigned long long (&dg)[4], uint32_t (&out)[S]) {
  const uint32_t l0 = (uint32_t)dg[0], l1 = (uint32_t)dg[1], l2 = (uint32_t)dg[2], l3 = (uint32_t)dg[3];
  const uint32_t p0 = __byte_perm(l0, l1, 0x5140), p1 = __byte_perm(l0, l1, 0x7362);
  const uint32_t q0 = __byte_perm(l2, l3, 0x5140), q1 = __byte_perm(l2, l3, 0x7362);
  out[0] = __byte_perm(p0, q0, 0x5410);
  if (S > 1) out[1] = __byte_perm(p0, q0, 0x7632);
  if (S > 2) out[2] = __byte_perm(p1, q1, 0x5410);
  if (S > 3) out[3] = __byte_perm(p1, q1, 0x7632);
  if (S > 4) {
    const uint32_t h0 = (uint32_t)(dg[0] >> 32), h1 = (uint32_t)(dg[1] >> 32), h2 = (uint32_t)(dg[2] >> 32),
                   h3 = (uint32_t)(dg[3] >> 32);
    const uint32_t r0 = __byte_perm(h0, h1, 0x5140), r1 = __byte_perm(h2, h3, 0x5140);
    out[4] = __byte_perm(r0, r1, 0x5410);
    if (S > 5) out[5] = __byte_perm(r0, r1, 0x7632);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace bocf
