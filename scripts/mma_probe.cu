// mma_probe.cu -- standalone probe of the tcgen05.mma kind::i8 instruction stream on B200 (sm_100a).
//
// Measures clocks per "K step" (one instruction per digit plane of the candidate-side operand, stacked factor planes
// along N -- the stream split_gemm.cu issues) for
//   * cta_group::1 (M = 128) and cta_group::2 (M = 256, CTA pairs),
//   * the A operand in shared memory (SS) or in tensor memory (TS),
//   * with and without a concurrent bulk-copy (TMA engine) stream refilling a shared-memory ring from L2.
// Operand contents are irrelevant (timing only).  Output: one line per configuration with the measured clocks per
// K step, the tensor-pipe floor (sum N/2) and the shared-memory operand bytes the stream reads.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bocf_b200/csrc scripts/mma_probe.cu -o scripts/mma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc05.cuh"

using namespace bocf;

constexpr int MAXSEQ = 8;
struct ProbeParams {
  int nseq;            // instructions per K step
  int N[MAXSEQ];       // N of each instruction
  int iters;           // K steps issued
  int ts;              // 1: A operand from tensor memory
  int loads;           // bytes per ring slot refilled by the bulk-copy stream (0 = none)
  int a_distinct;      // 1: every instruction reads its own A plane (the real stream); 0: all read plane 0
  const uint8_t* gsrc; // L2-resident source of the bulk copies
  long long* out;      // [grid][4]: mma clocks, load clocks, bytes loaded
};

constexpr int A_PLANE = 128 * 64;            // one digit plane of a 128-candidate tile, 64-byte K chunk
constexpr int RING_SLOTS = 4;

// compile-time instruction sequences (N per instruction of one K step)
constexpr int NSEQS = 11;
__host__ __device__ constexpr int seq_len(int id) {
  return id < 6 ? 1 : id == 6 ? 5 : id == 7 ? 4 : id == 8 ? 3 : id == 9 ? 3 : 4;
}
__host__ __device__ constexpr int seq_n(int id, int q) {
  switch (id) {
    case 0: return 48;
    case 1: return 96;
    case 2: return 144;
    case 3: return 192;
    case 4: return 240;
    case 5: return 256;
    case 6: return 240 - 48 * q;
    case 7: return 256 - 64 * q;
    case 8: return 192 - 64 * q;
    case 9: return 240 - 80 * q;
    default: return 192 - 48 * q;
  }
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n .reg .pred px;\n elect.sync _|px, 0xffffffff;\n selp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred;
}

template <int CG, int TS, int SEQ, int LEAN>
__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);      // [0] done, [1..4] ring slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 64);
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t off = ((raw + 1024 + 1023u) & ~1023u) - raw;
  uint8_t* sA = smem_raw + off;                  // 6 planes x 8 KB
  uint8_t* sB = sA + 6 * A_PLANE;                // 256 rows x 64 B x ... (16 KB)
  uint8_t* ring = sB + 256 * 64 * 2;             // RING_SLOTS x P.loads
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  for (int i = threadIdx.x; i < (6 * A_PLANE + 256 * 64 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sA)[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) {
    for (int b = 0; b < 1 + RING_SLOTS; ++b) tc::mbar_init(&bars[b], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tc::tmem_alloc2<512>(tmem_slot);
    else tc::tmem_alloc<512>(tmem_slot);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc::fence_before_sync();
  if (CG == 2) tc::cluster_sync_all();
  else __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = (CG == 2) ? tc::cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0 && P.loads > 0) {
    // bulk-copy stream: keep RING_SLOTS copies of P.loads bytes in flight until the MMA side is done
    long long t0 = clock64();
    long long bytes = 0;
    uint32_t phase[RING_SLOTS] = {0, 0, 0, 0};
    const uint8_t* g = P.gsrc + (size_t)blockIdx.x * (size_t)P.loads * RING_SLOTS;
    for (int s = 0; s < RING_SLOTS; ++s) {
      tc::mbar_arrive_expect_tx(&bars[1 + s], (uint32_t)P.loads);
      tc::bulk_g2s(ring + (size_t)s * P.loads, g + (size_t)s * P.loads, (uint32_t)P.loads, &bars[1 + s]);
    }
    int s = 0;
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem_raw + 128);
    while (true) {
      tc::mbar_wait(&bars[1 + s], phase[s]);
      phase[s] ^= 1u;
      bytes += P.loads;
      if (*flag) break;
      tc::mbar_arrive_expect_tx(&bars[1 + s], (uint32_t)P.loads);
      tc::bulk_g2s(ring + (size_t)s * P.loads, g + (size_t)s * P.loads, (uint32_t)P.loads, &bars[1 + s]);
      s = (s + 1) % RING_SLOTS;
    }
    long long t1 = clock64();
    // drain the copies still in flight
    for (int k = 1; k < RING_SLOTS; ++k) {
      const int q = (s + k) % RING_SLOTS;
      tc::mbar_wait(&bars[1 + q], phase[q]);
    }
    P.out[blockIdx.x * 4 + 1] = t1 - t0;
    P.out[blockIdx.x * 4 + 2] = bytes;
  }
  if (warp == 1) {
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem_raw + 128);
    long long t0 = 0, t1 = 0;
    if (crank == 0) {
      const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
      constexpr int L = seq_len(SEQ);
      t0 = clock64();
      if (LEAN) {
        // warp-uniform loop, descriptors advanced by adds on the low word, one elected lane issues
        const uint64_t ad0 = tc::smem_desc_sw64(a0), bd0 = tc::smem_desc_sw64(b0);
        for (int it = 0; it < P.iters; ++it) {
          const uint32_t ks = (uint32_t)(it & 1) * 2u;                 // +32 bytes = 2 sixteen-byte units
          const uint32_t acc = (it == 0) ? 0u : 1u;
          if (elect_one()) {
#pragma unroll
            for (int q = 0; q < L; ++q) {
              constexpr int dummy = 0;
              (void)dummy;
              const int N = seq_n(SEQ, q);
              const uint64_t bdesc = bd0 + (uint64_t)(q * 32 + ks);
              if (TS) {
                const uint32_t a_t = tmem_base + 256u + (uint32_t)(8 * q);
                if (CG == 2) {
                  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_base),
                               "r"(a_t), "l"(bdesc), "r"(tc::idesc_i8_m256(N)), "r"(acc)
                               : "memory");
                } else {
                  tc::mma_i8_ts(tmem_base, a_t, bdesc, tc::idesc_i8(N), acc);
                }
              } else {
                const uint64_t adesc = ad0 + (uint64_t)(q * (A_PLANE / 16) + ks);
                if (CG == 2) tc::mma_i8_pair(tmem_base, adesc, bdesc, tc::idesc_i8_m256(N), acc);
                else tc::mma_i8(tmem_base, adesc, bdesc, tc::idesc_i8(N), acc);
              }
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (CG == 2) tc::mma_commit_pair(&bars[0]);
          else tc::mma_commit(&bars[0]);
        }
      } else if (lane == 0) {
        // the round-1 style: one divergent lane, descriptors rebuilt per instruction
        for (int it = 0; it < P.iters; ++it) {
          const int ks = it & 1;
#pragma unroll
          for (int q = 0; q < L; ++q) {
            const int N = seq_n(SEQ, q);
            const uint64_t bdesc = tc::smem_desc_sw64(b0 + (uint32_t)(q * 512) + ks * 32);
            const uint32_t acc = (it == 0) ? 0u : 1u;
            const uint64_t adesc = tc::smem_desc_sw64(a0 + (uint32_t)(q * A_PLANE) + ks * 32);
            if (CG == 2) tc::mma_i8_pair(tmem_base, adesc, bdesc, tc::idesc_i8_m256(N), acc);
            else tc::mma_i8(tmem_base, adesc, bdesc, tc::idesc_i8(N), acc);
          }
        }
        if (CG == 2) tc::mma_commit_pair(&bars[0]);
        else tc::mma_commit(&bars[0]);
      }
    }
    __syncwarp();
    if (lane == 0) {
      tc::mbar_wait(&bars[0], 0);
      t1 = clock64();
      *flag = 1u;
      if (crank == 0) P.out[blockIdx.x * 4 + 0] = t1 - t0;
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  if (CG == 2) tc::cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::fence_after_sync();
    if (CG == 2) tc::tmem_dealloc2<512>(tmem_base);
    else tc::tmem_dealloc<512>(tmem_base);
  }
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      std::exit(1);                                                                    \
    }                                                                                  \
  } while (0)

template <int CG, int TS, int SEQ, int LEAN>
static void run_t(const char* name, int loads, const uint8_t* gsrc, long long* dout, int sms) {
  ProbeParams P;
  std::memset(&P, 0, sizeof(P));
  P.iters = 4000;
  P.loads = loads;
  P.gsrc = gsrc;
  P.out = dout;
  const int smem = 2048 + 6 * A_PLANE + 256 * 64 * 2 + RING_SLOTS * loads + 1024;
  CK(cudaMemset(dout, 0, sizeof(long long) * 4 * sms));
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(sms / CG * CG));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(probe_kernel<CG, TS, SEQ, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&cfg, probe_kernel<CG, TS, SEQ, LEAN>, P));
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(4 * sms);
  CK(cudaMemcpy(h.data(), dout, sizeof(long long) * 4 * sms, cudaMemcpyDeviceToHost));
  double mma = 0, ld_clk = 0, ld_bytes = 0;
  int nm = 0, nl = 0;
  for (int b = 0; b < sms / CG * CG; ++b) {
    if (h[b * 4 + 0] > 0) {
      mma += (double)h[b * 4 + 0];
      ++nm;
    }
    if (h[b * 4 + 1] > 0) {
      ld_clk += (double)h[b * 4 + 1];
      ld_bytes += (double)h[b * 4 + 2];
      ++nl;
    }
  }
  double floor_clk = 0, opbytes = 0;
  for (int q = 0; q < seq_len(SEQ); ++q) {
    const int N = seq_n(SEQ, q);
    floor_clk += N / 2.0;
    opbytes += (TS ? 0 : 4096) + 32.0 * N / CG;
  }
  const double per_step = nm ? mma / nm / P.iters : 0.0;
  std::printf("%-10s cg=%d %s %s loads=%6d | clk/step %8.1f  floor %6.1f  pipe %5.1f%%  smem-operand B/step/SM %7.0f -> %6.1f B/clk",
              name, CG, TS ? "TS" : "SS", LEAN ? "lean " : "naive", loads, per_step, floor_clk, 100.0 * floor_clk / per_step, opbytes,
              opbytes / per_step);
  if (nl) std::printf("  | bulk copies %6.1f B/clk/SM", ld_bytes / ld_clk);
  std::printf("\n");
  std::fflush(stdout);
}

static const char* SEQ_NAMES[NSEQS] = {"N=48", "N=96", "N=144", "N=192", "N=240", "N=256", "S5 NT48", "S4 NT64", "S3 NT64", "S3 NT80", "S4 NT48"};

template <int CG, int TS, int LEAN, int SEQ = 0>
static void run_all(int loads, const uint8_t* gsrc, long long* dout, int sms) {
  if constexpr (SEQ < NSEQS) {
    run_t<CG, TS, SEQ, LEAN>(SEQ_NAMES[SEQ], loads, gsrc, dout, sms);
    run_all<CG, TS, LEAN, SEQ + 1>(loads, gsrc, dout, sms);
  }
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long* dout = nullptr;
  CK(cudaMalloc(&dout, sizeof(long long) * 4 * sms));
  uint8_t* gsrc = nullptr;
  const size_t gbytes = (size_t)sms * 32768 * RING_SLOTS;
  CK(cudaMalloc(&gsrc, gbytes));
  CK(cudaMemset(gsrc, 1, gbytes));
  run_all<1, 0, 1>(0, gsrc, dout, sms);
  run_all<2, 0, 1>(0, gsrc, dout, sms);
  run_all<1, 1, 1>(0, gsrc, dout, sms);
  run_all<2, 1, 1>(0, gsrc, dout, sms);
  // the round-1 issue style for comparison
  run_t<1, 0, 6, 0>("S5 NT48", 0, gsrc, dout, sms);
  run_t<1, 0, 7, 0>("S4 NT64", 0, gsrc, dout, sms);
  run_t<1, 0, 5, 0>("N=256", 0, gsrc, dout, sms);
  // with the bulk-copy stream refilling a ring next to the operands
  for (int loads : {16384, 32768}) {
    run_t<1, 0, 6, 1>("S5 NT48", loads, gsrc, dout, sms);
    run_t<2, 0, 6, 1>("S5 NT48", loads, gsrc, dout, sms);
    run_t<1, 0, 8, 1>("S3 NT64", loads, gsrc, dout, sms);
    run_t<2, 0, 8, 1>("S3 NT64", loads, gsrc, dout, sms);
    run_t<1, 0, 5, 1>("N=256", loads, gsrc, dout, sms);
    run_t<2, 0, 5, 1>("N=256", loads, gsrc, dout, sms);
  }
  cudaFree(dout);
  cudaFree(gsrc);
  return 0;
}
