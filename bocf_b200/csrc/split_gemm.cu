// split_gemm.cu -- the two candidate-side contractions of the GP posterior on the 5th-generation tensor cores.
//
//   V  = K* Linv^T   (posterior.py:312, the reference's dtrtrs route)      var  = k** - sum_k V^2
//   Wt = V  Linv     (= K* W^-1, gp.py:474)                                 dvar = gradients_X(-2 Wt, x*, X)
//
// Blackwell's tcgen05 has no fp64 kind, so the fp64 operands are split into S signed 8-bit digits (balanced
// base-256 expansion of round(x * 2^(8S-2-e)), e a per-row power-of-two scale) and the product is assembled from
// S(S+1)/2 exact integer GEMMs on `tcgen05.mma.kind::i8` (SASS UTCIMMA): int8 x int8 products accumulated EXACTLY
// in int32 tensor memory, one accumulator per digit weight 256^(ta+tb).  The only error is the operand
// quantisation 2^-(8S-2) (relative to the row scale) -- S = 5 carries 38-bit operands and reproduces the fp64 path
// to ~1e-8 relative on the variance, S = 4 to ~1e-6 (tests/test_gpu_split.py measures both).
//
// Kernel structure (persistent, warp specialised, 320 threads, 1 CTA / SM):
//   warp 0   producer: one lane streams 64-byte-wide K chunks of the packed, pre-swizzled digit planes of A
//            (128 candidates) and B (NT factor rows) into a STAGES-deep shared-memory ring with bulk async copies
//            (TMA engine, cp.async.bulk + mbarrier complete_tx).
//   warp 1   MMA issuer: one lane issues, per 32-byte K step, S instructions  D[128 x (ta+1)NT] += A_ta * [B_tb]^T
//            whose B operand STACKS the digit planes tb = S-1-ta .. S-1 along N, so every A plane is read from
//            shared memory once per step while all S(S+1)/2 digit pairs are covered.  Accumulators live in TMEM,
//            double buffered when 2*S*NT <= 512 columns so the epilogue of tile t overlaps the MMAs of tile t+1.
//   warps 2-9 epilogue (two warps per TMEM lane group, alternating 8-column groups): tcgen05.ld the S int32 levels of
//            a row (lane = candidate), int32 -> fp64 by a magic-number add (no conversion-pipe instructions), Horner
//            them into one fp64 value, apply the column scale, and fuse the reductions of the reference:
//            VAR : sum_k V^2 per candidate (+ re-split V into digit planes = the A operand of the second GEMM)
//            DVAR: T = Wt * G*,  sum_b T and sum_b T Xs_b  (finalize_kernel forms xs_i sum T - sum T Xs_b).
//            The sums live in registers across all column tiles of a work unit (see NP below).
// Triangular structure of Linv is exploited per column tile: K chunks (64) for the loads, K steps (32) for the MMAs.
// A cta_group::2 variant (CG = 2: CTA pairs, M = 256) is kept as a tested option; it measured slower on B200
// (profiles/r1_split_experiments.md).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "model.h"
#include "tc05.cuh"

namespace bocf {
namespace sg {

constexpr int KC = 64;        // bytes (= int8 elements) of K per shared-memory row: one SWIZZLE_64B span
constexpr int TM = 128;       // candidate rows per tile (= TMEM lanes)
constexpr int EPI_WARPS = 8;     // two warps per TMEM lane group: latency hiding on the fp64 epilogue math
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int NTHREADS = 64 + EPI_THREADS;
constexpr int PART_SPLIT = EPI_WARPS / 4;   // partial sums per column tile (one per warp of a lane group)
constexpr int SMEM_MAX = 232448;

__host__ __device__ __forceinline__ uint32_t sw64(int r, int c) {   // byte offset of (row r, byte c) in a packed plane
  return (uint32_t)(r * 64 + ((((c >> 4) ^ ((r >> 1) & 3))) << 4) + (c & 15));
}
__host__ __device__ constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// CG = 1: one CTA per tile (M = 128).  CG = 2: the two CTAs of a cluster form an M = 256 tile (cta_group::2): each CTA
// brings its own 128 candidates (A planes, accumulators, epilogue) and HALF of the rows of every stacked B operand, so
// the shared-memory operand traffic per SM -- the measured limiter of the 1-CTA stream -- drops from 42.5 to 31 KB per
// K step.  Because the stacked operand of instruction ta starts at a different plane for every ta, its halves cannot
// alias one natural plane layout: the pair layout stores, per CTA rank, one region per instruction
// (rows [rank N_ta/2, (rank+1) N_ta/2) of the stack), S(S+1)/2 * NT/2 rows in total.
template <int S, int NT, int DP = 0, int CG = 1>
struct Cfg {
  static constexpr int A_PLANE = TM * KC, B_PLANE = NT * KC;
  static constexpr int B_ROWS = (CG == 1) ? S * NT : (S * (S + 1) / 2) * (NT / 2);
  static constexpr int A_STAGE = S * A_PLANE, B_STAGE = B_ROWS * KC, STAGE = A_STAGE + B_STAGE;
  // byte offset of the B operand of instruction ta (A plane ta against planes S-1-ta .. S-1) inside a stage
  __host__ __device__ static constexpr int b_off(int ta) {
    if (CG == 1) return (S - 1 - ta) * B_PLANE;
    int rows = 0;                                        // regions ordered ta = S-1 (largest) first
    for (int t = S - 1; t > ta; --t) rows += (t + 1) * (NT / 2);
    return rows * KC;
  }
  static constexpr int ACC_COLS = S * NT;
  static constexpr int NBUF = (2 * ACC_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = pow2_cols(NBUF * ACC_COLS);
  static constexpr int XB_BYTES = NT * DP * 8;                       // DVAR: scaled training inputs of the column tile
  static constexpr int HEAD = 1024 + XB_BYTES;                        // barriers, tmem slot, column scales, xb tile
  static constexpr int STAGES_FIT = (SMEM_MAX - HEAD - 512) / STAGE;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int SMEM_BYTES = HEAD + 512 + STAGES * STAGE;
  static_assert(S >= 2 && S <= 6 && NT % 16 == 0 && S * NT <= 256, "stacked B operand must fit one MMA (N <= 256)");
  static_assert(STAGES >= 2, "need at least a double-buffered ring");
  static_assert(B_STAGE % 512 == 0 && A_STAGE % 512 == 0, "planes must keep the 512-byte swizzle period");
  static_assert(NT * 8 <= 512 && (3 * STAGES + 3 * NBUF) * 8 + 16 <= 512, "head area too small");
  static_assert(CG == 1 || (NT / 2) % 8 == 0, "half operands must be whole 8-row groups");
};

enum { EPI_RAW = 0, EPI_VAR = 1, EPI_DVAR = 2 };
enum { TRI_FULL = 0, TRI_K_LE_N = 1, TRI_K_GE_N = 2 };

struct GemmParams {
  const uint8_t* A;        // [m][RT][KCH][S][128][64]   packed digit planes of the candidate-side operand
  const uint8_t* B;        // [Hm][nct][KCH][S][NT][64]  packed digit planes of the factor-side operand
  const double* cs;        // [Hm][nct*NT]               column scale (power of two)
  int m, h, RT, nct, KCH, n, tri;
  int64_t Nc, Nvalid;
  // RAW
  double* raw_out;         // [m*RT*128][ldo]
  const double* raw_rs;    // [m*RT*128] row scale
  int ldo;
  // VAR
  double* part_var;        // [m][nct][Nc]
  uint8_t* A2;             // [m][RT][KCH][S][128][64]   digit planes of V (nullptr: not needed)
  const double* vq;        // [Hm] 2^(8S-2-eV)
  // DVAR
  double* part_dvar;       // [m][nct*PART_SPLIT][Nc][d]   sum_b T_b Xs_bq
  double* part_s0;         // [m][nct*PART_SPLIT][Nc]      sum_b T_b
  const double* GsT;       // [m][n16][Nc]
  const double* Xc;        // [Nvalid][d]
  const double* Xs;        // [Hm][n_pad][d]
  const OutHyp* hyp;
  int d, n16, n_pad;
  int cg;                  // CTAs per tile group (1 or 2)
  int np;                  // parts per candidate tile
  int exp;                 // experiment knob (BOCF_SPLIT_EXP, results invalid): 1 skip the MMAs, 2 skip the bulk loads,
                           // 5 A operand from spare TMEM columns, 6 skip the epilogue work, 8 epilogue = TMEM loads only,
                           // 9 epilogue only (no loads, no MMAs)
};

// Work decomposition.  A UNIT is (output j, candidate tile rt, part p of NP): the column tiles ct = p, p+NP, ... of one
// 128-candidate tile.  A persistent CTA owns whole units and walks their column tiles back to back, so the epilogue
// keeps its per-candidate reductions in registers across the unit and writes ONE partial per (unit, epilogue warp)
// instead of one per column tile; the NP parts of a candidate tile run on neighbouring CTAs at the same time, which
// keeps the digit planes of the tile (re-read by every column tile) L2 resident.
constexpr int NP_DEFAULT = 2;   // parts per candidate tile (BOCF_SPLIT_NP overrides: 1, 2 or 4)
inline int parts_per_tile() {
  static int np = 0;
  if (np == 0) {
    np = NP_DEFAULT;
    if (const char* env = std::getenv("BOCF_SPLIT_NP")) {
      const int v = std::atoi(env);
      if (v == 1 || v == 2 || v == 4) np = v;
    }
  }
  return np;
}

struct TileInfo {
  int j, rt, p, ct, kb, ke;
  int sb, se;                   // first / one-past-last 32-byte K STEP that touches the triangle (MMA issue range)
  bool valid, first, last;      // first / last column tile of the unit
};

__device__ __forceinline__ int tiles_per_unit(const GemmParams& P) { return (P.nct + P.np - 1) / P.np; }
// With CTA pairs (P.cg == 2) the pair is the scheduling entity: both CTAs walk the same (output, part, column tile)
// sequence and CTA rank r of the pair owns candidate tile 2 * (pair's tile index) + r  (RT is even).
__device__ __forceinline__ int local_tile_count(const GemmParams& P) {      // slots this CTA walks (some may be empty)
  const int units = P.m * (P.RT / P.cg) * P.np;
  const int me = (int)blockIdx.x / P.cg, groups = (int)gridDim.x / P.cg;
  const int mine = (units > me) ? (units - 1 - me) / groups + 1 : 0;
  return mine * tiles_per_unit(P);
}

// slot l of this CTA: unit = group + (l / TPU) * groups, position l % TPU inside the unit
__device__ __forceinline__ TileInfo decode_tile(const GemmParams& P, int NT, int l) {
  TileInfo ti;
  const int tpu = tiles_per_unit(P);
  const int k = l / tpu, pos = l - k * tpu;
  const int me = (int)blockIdx.x / P.cg, groups = (int)gridDim.x / P.cg;
  const int u = me + k * groups;
  const int rtg = P.RT / P.cg;
  ti.j = u / (rtg * P.np);
  const int r = u - ti.j * (rtg * P.np);
  ti.rt = (r / P.np) * P.cg + ((int)blockIdx.x % P.cg);
  ti.p = r - (r / P.np) * P.np;
  ti.ct = ti.p + pos * P.np;
  ti.valid = ti.ct < P.nct;
  ti.first = (pos == 0);
  ti.last = (ti.ct + P.np >= P.nct);
  const int kch_used = (P.n + KC - 1) / KC;
  if (P.tri == TRI_K_LE_N) {                              // K index <= column index
    ti.kb = 0;
    ti.ke = min(kch_used, (min((ti.ct + 1) * NT, P.n) + KC - 1) / KC);
    ti.sb = 0;
    ti.se = min(ti.ke * (KC / 32), (min((ti.ct + 1) * NT, P.n) + 31) / 32);
  } else if (P.tri == TRI_K_GE_N) {                       // K index >= column index
    ti.kb = (ti.ct * NT) / KC;
    ti.ke = kch_used;
    ti.sb = (ti.ct * NT) / 32;
    ti.se = ti.ke * (KC / 32);
  } else {
    ti.kb = 0;
    ti.ke = kch_used;
    ti.sb = 0;
    ti.se = ti.ke * (KC / 32);
  }
  return ti;
}

// digits of a 64-bit integer |Y| < 2^(8S-2): byte t of the result is the balanced base-256 digit of weight 256^t
template <int S>
__device__ __forceinline__ unsigned long long balanced_digits(long long Y) {
  constexpr unsigned long long BIAS = (S == 6)   ? 0x808080808080ull
                                      : (S == 5) ? 0x8080808080ull
                                      : (S == 4) ? 0x80808080ull
                                      : (S == 3) ? 0x808080ull
                                                 : 0x8080ull;
  return ((unsigned long long)Y + BIAS) ^ BIAS;
}

// int32 -> double without the (slow) I2F.F64 conversion pipe: (2^52 + 2^31 + c) is exactly representable, one DADD.
__device__ __forceinline__ double i32_to_f64(uint32_t c) {
  return __hiloint2double(0x43300000, (int)(c ^ 0x80000000u)) - 4503601774854144.0;
}
// int64 (|x| < 2^51) -> double the same way: 64-bit integer add of the 1.5 * 2^52 bit pattern, one DADD.
__device__ __forceinline__ double i64_to_f64(long long x) {
  return __longlong_as_double(x + 0x4338000000000000LL) - 6755399441055744.0;
}
// sum_lb 256^lb c[lb]: adjacent levels are merged in 64-bit integer arithmetic first (c[lb] * 256 + c[lb-1] < 2^38), so
// only ceil(S/2) conversions and floor(S/2) fused multiply-adds reach the fp64 pipe -- the epilogue's scarce resource
// (ncu: math_pipe_throttle) -- instead of S and S-1.
template <int S, int W>
__device__ __forceinline__ double levels_to_f64(const uint32_t (&c)[S][W], int e) {
  double v = 0.0;
#pragma unroll
  for (int lb = S - 1; lb >= 0; lb -= 2) {
    if (lb >= 1) {
      const long long t = (long long)(int)c[lb][e] * 256 + (long long)(int)c[lb - 1][e];
      v = (lb == S - 1) ? i64_to_f64(t) : fma(v, 65536.0, i64_to_f64(t));
    } else {
      v = (lb == S - 1) ? i32_to_f64(c[0][e]) : fma(v, 256.0, i32_to_f64(c[0][e]));
    }
  }
  return v;
}

// digits of rint(x) for |x| < 2^46 without F2I: adding 1.5 * 2^52 leaves rint(x) (two's complement) in the low mantissa bits
template <int S>
__device__ __forceinline__ unsigned long long balanced_digits_of(double x) {
  return balanced_digits<S>(__double_as_longlong(x + 6755399441055744.0));
}

template <int S, int NT, int EPI, int DP, int CG>
__global__ void __launch_bounds__(NTHREADS, 1) split_gemm_kernel(const GemmParams P) {
  using C = Cfg<S, NT, DP, CG>;
  extern __shared__ uint8_t smem_raw[];
  // head: [0,512) barriers + tmem slot, [512, 1024) column scales; stages start at the next 512-byte boundary
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + C::NBUF;
  uint64_t* pfull = tempty + C::NBUF;          // CG = 2, leader only: the peer CTA's stage has landed (relayed)
  uint64_t* ptempty = pfull + C::STAGES;       // CG = 2, leader only: the peer CTA's epilogue has drained the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ptempty + C::NBUF);
  double* s_cs = reinterpret_cast<double*>(smem_raw + 512);
  double* s_xb = reinterpret_cast<double*>(smem_raw + 1024);      // [NT][DP]
  const uint32_t raw_addr = tc::smem_u32(smem_raw);
  const uint32_t stage_off = ((raw_addr + C::HEAD + 511u) & ~511u) - raw_addr;
  uint8_t* sA = smem_raw + stage_off;
  uint8_t* sB = sA + C::STAGES * C::A_STAGE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < C::NBUF; ++b) {
      tc::mbar_init(&tfull[b], 1);
      tc::mbar_init(&tempty[b], EPI_THREADS);
      tc::mbar_init(&ptempty[b], 1);
    }
    for (int s = 0; s < C::STAGES; ++s) tc::mbar_init(&pfull[s], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tc::tmem_alloc2<C::TMEM_COLS>(tmem_slot);
    else tc::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  tc::fence_before_sync();
  if (CG == 2) tc::cluster_sync_all();         // peer barriers are initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = (CG == 2) ? tc::cluster_ctarank() : 0u;
  const int num_tiles = local_tile_count(P);   // slots of THIS CTA

  if (warp == 0) {
    // ================================ producer =================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_keep = tc::l2_policy_evict_last();
      for (int t = 0; t < num_tiles; ++t) {
        const TileInfo ti = decode_tile(P, NT, t);
        if (!ti.valid) continue;
        const uint8_t* gA = P.A + ((size_t)(ti.j * P.RT + ti.rt) * P.KCH) * C::A_STAGE;
        // CG = 2: [..][kc][rank][regions]  -- this CTA streams its own half-operand block
        const uint8_t* gB = P.B + ((size_t)((P.h * P.m + ti.j) * P.nct + ti.ct) * P.KCH) * (C::B_STAGE * CG) +
                            (size_t)crank * C::B_STAGE;
        if (EPI == EPI_DVAR) {
          // the epilogue of this tile (one tile later in time) reads 128 G* values of each of its NT columns:
          // pull those 1 KB rows into L2 now so its loads do not pay HBM latency
          const double* g0 = P.GsT + (size_t)ti.j * P.n16 * P.Nc + (size_t)ti.rt * TM;
          const int bend = min(ti.ct * NT + NT, P.n);
          for (int b = ti.ct * NT; b < bend; ++b) tc::prefetch_l2(g0 + (size_t)b * P.Nc, TM * 8);
        }
        for (int kc = ti.kb; kc < ti.ke; ++kc) {
          tc::mbar_wait(&empty[stage], phase ^ 1u);
          if (P.exp == 2 || P.exp == 9) {
            tc::mbar_arrive(&full[stage]);
          } else {
            tc::mbar_arrive_expect_tx(&full[stage], C::STAGE);
            if (P.exp >= 10 && P.exp <= 12) {
              // experiment: the unit's A tile (re-streamed once per column tile) and the factor planes stay in L2
              tc::bulk_g2s_hint(sA + stage * C::A_STAGE, gA + (size_t)kc * C::A_STAGE, C::A_STAGE, &full[stage], pol_keep);
              if (P.exp == 12)
                tc::bulk_g2s_hint(sB + stage * C::B_STAGE, gB + (size_t)kc * (C::B_STAGE * CG), C::B_STAGE, &full[stage], pol_keep);
              else
                tc::bulk_g2s(sB + stage * C::B_STAGE, gB + (size_t)kc * (C::B_STAGE * CG), C::B_STAGE, &full[stage]);
            } else {
              tc::bulk_g2s(sA + stage * C::A_STAGE, gA + (size_t)kc * C::A_STAGE, C::A_STAGE, &full[stage]);
              tc::bulk_g2s(sB + stage * C::B_STAGE, gB + (size_t)kc * (C::B_STAGE * CG), C::B_STAGE, &full[stage]);
            }
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
#define WAITFN(bar, par) tc::mbar_wait(bar, par)
    // ================================ MMA issuer ===============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const TileInfo ti = decode_tile(P, NT, t);
        if (!ti.valid) continue;
        const int buf = it % C::NBUF;
        const uint32_t use = (uint32_t)(it / C::NBUF);
        ++it;
        if (CG == 2 && crank != 0) {
          // follower of the pair: issues nothing; relays "my epilogue drained the buffer" and "my stage landed" to the
          // leader, whose MMAs read this CTA's shared memory and write this CTA's tensor memory
          WAITFN(&tempty[buf], (use & 1u) ^ 1u);
          tc::mbar_arrive_remote(&ptempty[buf], 0);
          for (int kc = ti.kb; kc < ti.ke; ++kc) {
            WAITFN(&full[stage], phase);
            tc::mbar_arrive_remote(&pfull[stage], 0);
            if (++stage == C::STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          continue;
        }
        WAITFN(&tempty[buf], (use & 1u) ^ 1u);          // epilogue has drained this accumulator buffer
        if (CG == 2) WAITFN(&ptempty[buf], use & 1u);   // ... in the peer CTA too
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * C::ACC_COLS);
        bool first = true;
        for (int kc = ti.kb; kc < ti.ke; ++kc) {
          WAITFN(&full[stage], phase);
          if (CG == 2) WAITFN(&pfull[stage], phase);
          tc::fence_after_sync();
          const uint32_t a0 = tc::smem_u32(sA + stage * C::A_STAGE);
          const uint32_t b0 = tc::smem_u32(sB + stage * C::B_STAGE);
#pragma unroll
          for (int ks = 0; ks < KC / 32; ++ks) {
            const int kstep = kc * (KC / 32) + ks;            // K steps wholly outside the triangle multiply zeros: skip
            if (kstep < ti.sb || kstep >= ti.se) continue;
#pragma unroll
            for (int ta = S - 1; ta >= 0; --ta) {
              // A digit plane ta against the stacked B planes tb = S-1-ta .. S-1  ->  levels 0 .. ta
              const uint64_t adesc = tc::smem_desc_sw64(a0 + ta * C::A_PLANE + ks * 32);
              const uint64_t bdesc = tc::smem_desc_sw64(b0 + C::b_off(ta) + ks * 32);
              const uint32_t acc = (first && ta == S - 1) ? 0u : 1u;
              if (P.exp == 5 && CG == 1 && C::NBUF * C::ACC_COLS + 8 * S <= C::TMEM_COLS) {
                // experiment (timing only, results invalid): A operand from spare tensor-memory columns
                tc::mma_i8_ts(d_tmem, tmem_base + (uint32_t)(C::NBUF * C::ACC_COLS + 8 * ta), bdesc, tc::idesc_i8((ta + 1) * NT), acc);
              } else if (P.exp != 1 && P.exp != 9) {
                if (CG == 2) tc::mma_i8_pair(d_tmem, adesc, bdesc, tc::idesc_i8_m256((ta + 1) * NT), acc);
                else tc::mma_i8(d_tmem, adesc, bdesc, tc::idesc_i8((ta + 1) * NT), acc);
              }
            }
            first = false;
          }
          // stage reusable once these MMAs have read it (both CTAs' stages for a pair)
          if (CG == 2) tc::mma_commit_pair(&empty[stage]);
          else tc::mma_commit(&empty[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulators of this tile complete (in both CTAs' tensor memory for a pair)
        if (CG == 2) tc::mma_commit_pair(&tfull[buf]);
        else tc::mma_commit(&tfull[buf]);
      }
    }
#undef WAITFN
  } else {
    // ================================ epilogue (EPI_WARPS warps) ===============================
    const int et = threadIdx.x - 64;
    const int lg = warp & 3;                                    // TMEM lane group this warp may access
    const int hw = (warp - 2) >> 2;                             // which of the lane group's warps: owns column groups hw, hw+PART_SPLIT, ..
    const int row = lg * 32 + lane;
    int it = 0;
    // per-tile constants (column scales, scaled training inputs of the column tile) are fetched one tile AHEAD into
    // registers and parked in shared memory at the top of the tile, so their global-load latency is never exposed
    constexpr int DPA0 = DP > 0 ? DP : 1;
    constexpr int XPT = (EPI == EPI_DVAR) ? (NT * DPA0 + EPI_THREADS - 1) / EPI_THREADS : 1;
    double pre_cs = 0.0, pre_vq = 0.0, pre_xb[XPT];
    auto prefetch_tile_consts = [&](int tnext) {
      TileInfo tn;
      tn.valid = false;
      while (tnext < num_tiles && !(tn = decode_tile(P, NT, tnext)).valid) ++tnext;
      if (!tn.valid) return;
      const int hjn = P.h * P.m + tn.j;
      if (et < NT) pre_cs = __ldg(P.cs + (size_t)hjn * P.nct * NT + tn.ct * NT + et);
      if (EPI == EPI_VAR) pre_vq = __ldg(P.vq + hjn);
      if (EPI == EPI_DVAR) {
        const double* Xb = P.Xs + (size_t)hjn * P.n_pad * P.d;
#pragma unroll
        for (int x = 0; x < XPT; ++x) {
          const int idx = et + x * EPI_THREADS;
          const int cc = idx / DPA0, q = idx - cc * DPA0;
          const int b = tn.ct * NT + cc;
          pre_xb[x] = (idx < NT * DPA0 && q < P.d && b < P.n) ? __ldg(Xb + (size_t)b * P.d + q) : 0.0;
        }
      }
    };
    prefetch_tile_consts(0);
    // reductions over the column tiles of a unit live in registers
    constexpr int DPA = DP > 0 ? DP : 2;
    double sumsq = 0.0, s0 = 0.0, acc[DPA];
    for (int t = 0; t < num_tiles; ++t) {
      const TileInfo ti = decode_tile(P, NT, t);
      if (!ti.valid) continue;
      const int buf = it % C::NBUF;
      const uint32_t use = (uint32_t)(it / C::NBUF);
      ++it;
      const int col0 = ti.ct * NT;
      if (ti.first) {
        sumsq = 0.0;
        s0 = 0.0;
#pragma unroll
        for (int q = 0; q < DPA; ++q) acc[q] = 0.0;
      }
      tc::named_bar_sync(1, EPI_THREADS);                       // previous tile's readers of s_cs / s_xb are done
      if (et < NT) s_cs[et] = pre_cs;
      if (EPI == EPI_DVAR) {
#pragma unroll
        for (int x = 0; x < XPT; ++x) {
          const int idx = et + x * EPI_THREADS;
          if (idx < NT * DPA0) s_xb[idx] = pre_xb[x];
        }
      }
      const double vq = pre_vq;
      tc::named_bar_sync(1, EPI_THREADS);
      prefetch_tile_consts(t + 1);
      const int64_t i = (int64_t)ti.rt * TM + row;              // chunk-local candidate

      // per-tile thread state
      constexpr int CGW = 8;                                    // columns per TMEM load group
      constexpr int NCG = NT / CGW;
      double gv[CGW];
      const double* Gcol = nullptr;
      const bool g_stream = (P.exp == 11 || P.exp == 12);       // experiment: G* is read once -> streaming loads
      if (EPI == EPI_DVAR) {
        Gcol = P.GsT + (size_t)ti.j * P.n16 * P.Nc + i;
#pragma unroll
        for (int e = 0; e < CGW; ++e) {                      // first column group: in flight while the MMAs finish
          const int b = col0 + hw * CGW + e;
          gv[e] = (b < P.n) ? (g_stream ? __ldcs(Gcol + (size_t)b * P.Nc) : __ldg(Gcol + (size_t)b * P.Nc)) : 0.0;
        }
      }

      tc::mbar_wait(&tfull[buf], use & 1u);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * C::ACC_COLS);

#pragma unroll 1
      for (int cg = hw; cg < NCG; cg += PART_SPLIT) {
        if (P.exp == 6) break;                                    // experiment: no epilogue work (timing only)
        uint32_t c[S][CGW];
#pragma unroll
        for (int lb = 0; lb < S; ++lb) tc::tmem_ldw<CGW>(taddr + (uint32_t)(lb * NT + cg * CGW), c[lb]);
        tc::tmem_ld_wait();
        if (P.exp == 8) {                                         // experiment: TMEM loads only, no epilogue math
          uint32_t x = 0;
#pragma unroll
          for (int lb = 0; lb < S; ++lb)
#pragma unroll
            for (int e = 0; e < CGW; ++e) x ^= c[lb][e];
          if (x == 0x9e3779b9u) sumsq += 1.0;                     // keeps the loads alive
          continue;
        }
        uint32_t vec[S][2];
        unsigned long long dgv[CGW];
#pragma unroll
        for (int e = 0; e < CGW; ++e) {
          double v = levels_to_f64<S, CGW>(c, e);
          if (EPI != EPI_DVAR) v *= s_cs[cg * CGW + e];             // DVAR: the column scale is folded into G* by the K* kernel
          if (EPI == EPI_RAW) {
            const int col = col0 + cg * CGW + e;
            const size_t grow = (size_t)(ti.j * P.RT + ti.rt) * TM + row;
            if (col < P.ldo) P.raw_out[grow * P.ldo + col] = v * P.raw_rs[grow];
          } else if (EPI == EPI_VAR) {
            sumsq = fma(v, v, sumsq);
            dgv[e] = balanced_digits_of<S>(v * vq);
          } else {
            const double w = v * gv[e];
            {                                                  // refill the slot with this warp's next column group's G*
              const int bn = col0 + (cg + PART_SPLIT) * CGW + e;
              gv[e] = (cg + PART_SPLIT < NCG && bn < P.n)
                          ? (g_stream ? __ldcs(Gcol + (size_t)bn * P.Nc) : __ldg(Gcol + (size_t)bn * P.Nc)) : 0.0;
            }
            s0 += w;
            const double2* xb2 = reinterpret_cast<const double2*>(s_xb + (cg * CGW + e) * DPA);
#pragma unroll
            for (int q2 = 0; q2 < DPA / 2; ++q2) {
              const double2 x2 = xb2[q2];
              acc[2 * q2] = fma(w, x2.x, acc[2 * q2]);
              acc[2 * q2 + 1] = fma(w, x2.y, acc[2 * q2 + 1]);
            }
          }
        }
        if (EPI == EPI_VAR && P.A2 != nullptr) {
#pragma unroll
          for (int w = 0; w < 2; ++w) {                          // byte transpose: digit t of 4 columns -> one word
            const unsigned long long four[4] = {dgv[4 * w], dgv[4 * w + 1], dgv[4 * w + 2], dgv[4 * w + 3]};
            uint32_t o[S];
            digits_transpose4<S>(four, o);
#pragma unroll
            for (int tt = 0; tt < S; ++tt) vec[tt][w] = o[tt];
          }
          const int k = col0 + cg * CGW;
          if (k < P.KCH * KC) {
            const int kc = k >> 6;
            uint8_t* dst = P.A2 + (((size_t)(ti.j * P.RT + ti.rt) * P.KCH + kc) * S) * C::A_PLANE + sw64(row, k & 63);
#pragma unroll
            for (int tt = 0; tt < S; ++tt)
              if (P.exp == 11 || P.exp == 12) __stcs(reinterpret_cast<uint2*>(dst + (size_t)tt * C::A_PLANE), make_uint2(vec[tt][0], vec[tt][1]));
              else *reinterpret_cast<uint2*>(dst + (size_t)tt * C::A_PLANE) = make_uint2(vec[tt][0], vec[tt][1]);
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&tempty[buf]);                            // accumulator buffer may be overwritten

      // one partial per (unit part, warp of the lane group): summed in fixed order by finalize_kernel
      const size_t pidx = ((size_t)ti.j * P.np + ti.p) * PART_SPLIT + hw;
      if (EPI == EPI_VAR && ti.last) P.part_var[pidx * P.Nc + i] = sumsq;
      if (EPI == EPI_DVAR && ti.last) {
        // xs_iq * S0 - ACC_q is formed by finalize_kernel (one division per candidate instead of one per tile)
        P.part_s0[pidx * P.Nc + i] = s0;
        double* out = P.part_dvar + (pidx * P.Nc + i) * P.d;
#pragma unroll
        for (int q = 0; q < DPA; ++q)
          if (q < P.d) out[q] = acc[q];
      }
    }
  }
  tc::fence_before_sync();
  if (CG == 2) tc::cluster_sync_all();         // no CTA leaves while its peer may still signal it / use its memories
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::fence_after_sync();
    if (CG == 2) tc::tmem_dealloc2<C::TMEM_COLS>(tmem_base);
    else tc::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ---- digit-plane packing of a dense fp64 matrix ------------------------------------------------------------------
// element (r, k) of matrix `mat`: src[mat*mat_stride + r*sr + k*sk], zero outside r < R, k < K.
__global__ void row_exp_kernel(const double* __restrict__ src, int64_t mat_stride, int64_t sr, int64_t sk, int R, int K,
                               int Rpad, int* __restrict__ exps, double* __restrict__ cs,
                               const int* __restrict__ extra, int base) {
  const int mat = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= Rpad) return;
  double amax = 0.0;
  if (row < R)
    for (int k = lane; k < K; k += 32) amax = fmax(amax, fabs(src[mat * mat_stride + row * sr + k * sk]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) {
    int e = 0;
    if (amax > 0.0) frexp(amax * 1.02, &e);
    exps[(size_t)mat * Rpad + row] = e;
    if (cs) cs[(size_t)mat * Rpad + row] = ldexp(1.0, e + (extra ? extra[mat] : 0) + base);
  }
}

template <int S>
__global__ void pack_digits_kernel(const double* __restrict__ src, int64_t mat_stride, int64_t sr, int64_t sk, int R,
                                   int K, const int* __restrict__ exps, int Rpad, int TR, int ntile, int KCH,
                                   uint8_t* __restrict__ out) {
  const int kc = blockIdx.x, tile = blockIdx.y, mat = blockIdx.z;
  uint8_t* obase = out + (((size_t)(mat * ntile + tile) * KCH + kc) * S) * TR * KC;
  for (int idx = threadIdx.x; idx < TR * 4; idx += blockDim.x) {
    int r, piece;
    if (sk == 1) {
      r = idx >> 2;
      piece = idx & 3;
    } else {
      r = idx % TR;
      piece = idx / TR;
    }
    const int row = tile * TR + r;
    const int e = exps[(size_t)mat * Rpad + row];
    const double q = ldexp(1.0, 8 * S - 2 - e);
    uint32_t vec[S][4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      unsigned long long four[4];
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const int k = kc * KC + piece * 16 + 4 * w + e4;
        const double x = (row < R && k < K) ? src[mat * mat_stride + row * sr + k * sk] : 0.0;
        four[e4] = balanced_digits_of<S>(x * q);
      }
      uint32_t o[S];
      digits_transpose4<S>(four, o);
#pragma unroll
      for (int tt = 0; tt < S; ++tt) vec[tt][w] = o[tt];
    }
#pragma unroll
    for (int tt = 0; tt < S; ++tt)
      *reinterpret_cast<uint4*>(obase + (size_t)tt * TR * KC + sw64(r, piece * 16)) =
          make_uint4(vec[tt][0], vec[tt][1], vec[tt][2], vec[tt][3]);
  }
}

// Pair layout of the factor-side operand (cta_group::2, see Cfg): per (tile, K chunk) two blocks (CTA rank 0 / 1), each
// holding for every instruction ta = S-1 .. 0 the rows [rank N_ta/2, (rank+1) N_ta/2) of the stack of planes S-1-ta .. S-1.
template <int S>
__global__ void pack_digits_pair_kernel(const double* __restrict__ src, int64_t mat_stride, int64_t sr, int64_t sk,
                                        int R, int K, const int* __restrict__ exps, int Rpad, int NTr, int ntile,
                                        int KCH, uint8_t* __restrict__ out) {
  const int kc = blockIdx.x, tile = blockIdx.y, mat = blockIdx.z;
  const int half = NTr / 2;
  const int rows_per_rank = (S * (S + 1) / 2) * half;
  uint8_t* obase = out + ((size_t)(mat * ntile + tile) * KCH + kc) * (size_t)(2 * rows_per_rank * KC);
  for (int idx = threadIdx.x; idx < 2 * rows_per_rank * 4; idx += blockDim.x) {
    int piece, rr_all;
    if (sk == 1) {
      rr_all = idx >> 2;
      piece = idx & 3;
    } else {
      rr_all = idx % (2 * rows_per_rank);
      piece = idx / (2 * rows_per_rank);
    }
    const int rank = rr_all / rows_per_rank;
    int rem = rr_all - rank * rows_per_rank;
    int ta = S - 1, off_rows = 0;
    while (rem >= (ta + 1) * half) {                     // regions ordered ta = S-1 first
      rem -= (ta + 1) * half;
      off_rows += (ta + 1) * half;
      --ta;
    }
    const int sidx = rank * (ta + 1) * half + rem;       // row inside the stack of instruction ta
    const int tb = (S - 1 - ta) + sidx / NTr;
    const int r = sidx % NTr;
    const int row = tile * NTr + r;
    const int e = exps[(size_t)mat * Rpad + row];
    const double q = ldexp(1.0, 8 * S - 2 - e);
    uint32_t vec[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int ee = 0; ee < 16; ++ee) {
      const int k = kc * KC + piece * 16 + ee;
      const double x = (row < R && k < K) ? src[mat * mat_stride + row * sr + k * sk] : 0.0;
      const unsigned long long dg = balanced_digits_of<S>(x * q);
      vec[ee >> 2] |= (uint32_t)((dg >> (8 * tb)) & 0xFFull) << (8 * (ee & 3));
    }
    *reinterpret_cast<uint4*>(obase + (size_t)rank * rows_per_rank * KC + (size_t)off_rows * KC + sw64(rem, piece * 16)) =
        make_uint4(vec[0], vec[1], vec[2], vec[3]);
  }
}

__global__ void absmax_kernel(const double* __restrict__ src, size_t count, double* __restrict__ out) {
  double a = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x)
    a = fmax(a, fabs(src[idx]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
  if ((threadIdx.x & 31) == 0 && a > 0.0)
    atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(a));
}

template <int S>
static int pack_t(const double* src, int64_t mat_stride, int64_t sr, int64_t sk, int R, int K, const int* exps, int Rpad,
                  int TR, int ntile, int KCH, int mats, uint8_t* out, cudaStream_t st) {
  dim3 grid((unsigned)KCH, (unsigned)ntile, (unsigned)mats);
  pack_digits_kernel<S><<<grid, 256, 0, st>>>(src, mat_stride, sr, sk, R, K, exps, Rpad, TR, ntile, KCH, out);
  BOCF_LAUNCH_OK("pack_digits_kernel");
  return 0;
}
static int pack_digits_pair(int S, const double* src, int64_t mat_stride, int64_t sr, int64_t sk, int R, int K,
                            const int* exps, int Rpad, int NTr, int ntile, int KCH, int mats, uint8_t* out,
                            cudaStream_t st) {
  dim3 grid((unsigned)KCH, (unsigned)ntile, (unsigned)mats);
  switch (S) {
    case 3: pack_digits_pair_kernel<3><<<grid, 256, 0, st>>>(src, mat_stride, sr, sk, R, K, exps, Rpad, NTr, ntile, KCH, out); break;
    case 4: pack_digits_pair_kernel<4><<<grid, 256, 0, st>>>(src, mat_stride, sr, sk, R, K, exps, Rpad, NTr, ntile, KCH, out); break;
    case 5: pack_digits_pair_kernel<5><<<grid, 256, 0, st>>>(src, mat_stride, sr, sk, R, K, exps, Rpad, NTr, ntile, KCH, out); break;
    case 6: pack_digits_pair_kernel<6><<<grid, 256, 0, st>>>(src, mat_stride, sr, sk, R, K, exps, Rpad, NTr, ntile, KCH, out); break;
    default: set_error("split contraction supports 3..6 digit planes"); return BOCF_ERR_INVALID;
  }
  BOCF_LAUNCH_OK("pack_digits_pair_kernel");
  return 0;
}
static int pack_digits(int S, const double* src, int64_t mat_stride, int64_t sr, int64_t sk, int R, int K,
                       const int* exps, int Rpad, int TR, int ntile, int KCH, int mats, uint8_t* out, cudaStream_t st) {
  switch (S) {
    case 3: return pack_t<3>(src, mat_stride, sr, sk, R, K, exps, Rpad, TR, ntile, KCH, mats, out, st);
    case 4: return pack_t<4>(src, mat_stride, sr, sk, R, K, exps, Rpad, TR, ntile, KCH, mats, out, st);
    case 5: return pack_t<5>(src, mat_stride, sr, sk, R, K, exps, Rpad, TR, ntile, KCH, mats, out, st);
    case 6: return pack_t<6>(src, mat_stride, sr, sk, R, K, exps, Rpad, TR, ntile, KCH, mats, out, st);
  }
  set_error("split contraction supports 3..6 digit planes");
  return BOCF_ERR_INVALID;
}
static int row_exps(const double* src, int64_t mat_stride, int64_t sr, int64_t sk, int R, int K, int Rpad, int mats,
                    int* exps, double* cs, const int* extra, int base, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(Rpad, 8), (unsigned)mats);
  row_exp_kernel<<<grid, 256, 0, st>>>(src, mat_stride, sr, sk, R, K, Rpad, exps, cs, extra, base);
  BOCF_LAUNCH_OK("row_exp_kernel");
  return 0;
}

template <int S, int NT, int EPI, int DP, int CG>
static int launch_cg(const GemmParams& P, cudaStream_t st) {
  using C = Cfg<S, NT, DP, CG>;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static bool attr_done[64] = {false};          // per device: function attributes belong to the device's context
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    BOCF_CUDA_OK(cudaFuncSetAttribute(split_gemm_kernel<S, NT, EPI, DP, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      C::SMEM_BYTES));
    attr_done[dev] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = P.m * (P.RT / CG) * P.np;     // scheduling entities (CTAs, or CTA pairs)
  if (const char* env = std::getenv("BOCF_SPLIT_GRID")) {          // experiment knob: persistent CTAs launched
    const int g = std::atoi(env);
    if (g > 0 && g < sms) sms = g;
  }
  int groups = sms / CG;
  if (units < groups) groups = units;
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, split_gemm_kernel<S, NT, EPI, DP, CG>, P);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("launch split_gemm_kernel: ") + cudaGetErrorString(e));
    return -2;
  }
  return 0;
}
template <int S, int NT, int EPI, int DP = 0>
static int launch_t(const GemmParams& P, cudaStream_t st) {
  if (P.cg == 2) return launch_cg<S, NT, EPI, DP, 2>(P, st);
  return launch_cg<S, NT, EPI, DP, 1>(P, st);
}
template <int EPI, int DP = 0>
static int launch_s(int S, const GemmParams& P, cudaStream_t st) {
  switch (S) {
    case 3: return launch_t<3, 64, EPI, DP>(P, st);
    case 4: return launch_t<4, 64, EPI, DP>(P, st);
    case 5: return launch_t<5, 48, EPI, DP>(P, st);
    case 6: return launch_t<6, 32, EPI, DP>(P, st);
  }
  set_error("split contraction supports 3..6 digit planes");
  return BOCF_ERR_INVALID;
}
static int launch_dvar(int S, int d, const GemmParams& P, cudaStream_t st) {
  if (d <= 4) return launch_s<EPI_DVAR, 4>(S, P, st);
  if (d <= 6) return launch_s<EPI_DVAR, 6>(S, P, st);
  if (d <= 8) return launch_s<EPI_DVAR, 8>(S, P, st);
  if (d <= 10) return launch_s<EPI_DVAR, 10>(S, P, st);
  if (d <= 12) return launch_s<EPI_DVAR, 12>(S, P, st);
  return launch_s<EPI_DVAR, 16>(S, P, st);
}

}  // namespace sg

int split_column_tile(int S) { return S == 5 ? 48 : (S == 6 ? 32 : 64); }
int split_partials_per_tile() { return sg::PART_SPLIT * sg::parts_per_tile(); }   // partial sums per candidate and output

static void free_split(bocf_model* M) {
  auto fr = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  fr(M->B1);
  fr(M->B2);
  fr(M->cs1);
  fr(M->cs2);
  fr(M->aq);
  fr(M->vq);
  M->split_ready = false;
}
void split_release(bocf_model* M) { free_split(M); }

// Largest |Linv| entry over all (h, j): drives the automatic choice of the number of digit planes.
int split_linv_absmax(bocf_model* M, double* out_host, cudaStream_t st) {
  double* d = nullptr;
  BOCF_CUDA_OK(cudaMalloc(&d, sizeof(double)));
  BOCF_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(double), st));
  const size_t count = (size_t)M->H * M->m * M->n_pad * M->n_pad;
  sg::absmax_kernel<<<1024, 256, 0, st>>>(M->Linv, count, d);
  BOCF_LAUNCH_OK("absmax_kernel");
  BOCF_CUDA_OK(cudaMemcpyAsync(out_host, d, sizeof(double), cudaMemcpyDeviceToHost, st));
  BOCF_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

// Digit planes of Linv for both contractions + all power-of-two scales.  Called after the factorisation.
int split_prepare(bocf_model* M, int S, cudaStream_t st) {
  free_split(M);
  const int Hm = M->H * M->m;
  const int NT = split_column_tile(S);
  M->S = S;
  M->NTs = NT;
  M->ncts = (int)ceil_div(M->n, NT);
  M->KCH = (int)ceil_div(M->n, sg::KC);
  const int Rpad = M->ncts * NT;
  // CTA pairs (cta_group::2, BOCF_SPLIT_CG=2) are implemented and bit-exact but measured SLOWER on B200 (first
  // contraction 26.6 vs 15.6 ms per 131072 candidates, profiles/r1_split_experiments.md): single CTAs are the default.
  M->split_cg = 1;
  if (const char* env = std::getenv("BOCF_SPLIT_CG"))
    if (std::atoi(env) == 2) M->split_cg = 2;
  const size_t rows_per_tile = (M->split_cg == 2) ? (size_t)S * (S + 1) / 2 * NT : (size_t)S * NT;   // both ranks
  const size_t plane_bytes = (size_t)Hm * M->ncts * M->KCH * rows_per_tile * sg::KC;
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&M->B1), plane_bytes));
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&M->B2), plane_bytes));
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&M->cs1), sizeof(double) * Hm * Rpad));
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&M->cs2), sizeof(double) * Hm * Rpad));
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&M->aq), sizeof(double) * Hm));
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&M->vq), sizeof(double) * Hm));
  int *exps = nullptr, *extra = nullptr;
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&exps), sizeof(int) * Hm * Rpad));
  BOCF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&extra), sizeof(int) * 2 * Hm));
  // candidate-side scales: K* <= sigma_f^2 (stationary kernels), |V_k| <= sqrt(k**) = sigma_f
  std::vector<int> ex(2 * Hm);
  std::vector<double> aq(Hm), vq(Hm);
  for (int hj = 0; hj < Hm; ++hj) {
    int eA = 0, eV = 0;
    std::frexp(M->hyp_host[hj].variance * 1.02, &eA);
    std::frexp(std::sqrt(M->hyp_host[hj].variance) * 1.02, &eV);
    ex[hj] = eA;
    ex[Hm + hj] = eV;
    aq[hj] = std::ldexp(1.0, 8 * S - 2 - eA);
    vq[hj] = std::ldexp(1.0, 8 * S - 2 - eV);
  }
  BOCF_CUDA_OK(cudaMemcpyAsync(extra, ex.data(), sizeof(int) * 2 * Hm, cudaMemcpyHostToDevice, st));
  BOCF_CUDA_OK(cudaMemcpyAsync(M->aq, aq.data(), sizeof(double) * Hm, cudaMemcpyHostToDevice, st));
  BOCF_CUDA_OK(cudaMemcpyAsync(M->vq, vq.data(), sizeof(double) * Hm, cudaMemcpyHostToDevice, st));
  const int base = -2 * (8 * S - 2) + 8 * (S - 1);
  const int64_t nn = (int64_t)M->n_pad * M->n_pad;
  int rc = 0;
  // first contraction: rows = factor row k, K = b      (element Linv[k][b])
  if (!rc) rc = sg::row_exps(M->Linv, nn, M->n_pad, 1, M->n, M->n, Rpad, Hm, exps, M->cs1, extra, base, st);
  if (!rc)
    rc = (M->split_cg == 2)
             ? sg::pack_digits_pair(S, M->Linv, nn, M->n_pad, 1, M->n, M->n, exps, Rpad, NT, M->ncts, M->KCH, Hm, M->B1, st)
             : sg::pack_digits(S, M->Linv, nn, M->n_pad, 1, M->n, M->n, exps, Rpad, NT, M->ncts, M->KCH, Hm, M->B1, st);
  // second contraction: rows = factor column b, K = k  (element Linv[k][b])
  if (!rc) rc = sg::row_exps(M->Linv, nn, 1, M->n_pad, M->n, M->n, Rpad, Hm, exps, M->cs2, extra + Hm, base, st);
  if (!rc)
    rc = (M->split_cg == 2)
             ? sg::pack_digits_pair(S, M->Linv, nn, 1, M->n_pad, M->n, M->n, exps, Rpad, NT, M->ncts, M->KCH, Hm, M->B2, st)
             : sg::pack_digits(S, M->Linv, nn, 1, M->n_pad, M->n, M->n, exps, Rpad, NT, M->ncts, M->KCH, Hm, M->B2, st);
  cudaError_t e = cudaStreamSynchronize(st);      // ex/aq/vq are host vectors going out of scope
  cudaFree(exps);
  cudaFree(extra);
  if (rc) return rc;
  if (e != cudaSuccess) {
    set_error(std::string("split_prepare: ") + cudaGetErrorString(e));
    return BOCF_ERR_CUDA;
  }
  M->split_ready = true;
  return 0;
}

uint64_t split_chunk_bytes_per_candidate(const bocf_model* M, bool grad) {
  uint64_t per = 0;
  per += (uint64_t)M->m * M->KCH * sg::KC * M->S;       // A1 digit planes of K*
  per += (uint64_t)M->m * sg::parts_per_tile() * sg::PART_SPLIT * 8;  // part_var
  per += 2ull * M->m * 8;                               // mean, var
  if (grad) {
    per += (uint64_t)M->m * M->n16 * 8;                 // GsT
    per += (uint64_t)M->m * M->KCH * sg::KC * M->S;     // A2 digit planes of V
    per += (uint64_t)M->m * sg::parts_per_tile() * sg::PART_SPLIT * (M->d + 1) * 8;   // part_dvar, part_s0
    per += 2ull * M->m * M->d * 8;                      // dmean, dvar
  }
  return per;
}

void split_carve_chunk(const bocf_model* M, void* base, int64_t Nc, bool grad, ChunkBuffers* cb) {
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](uint64_t bytes) {
    uint8_t* r = p;
    p += round_up((int64_t)bytes, 1024);
    return r;
  };
  const uint64_t planes = (uint64_t)M->m * Nc * M->KCH * sg::KC * M->S;
  cb->Nc = Nc;
  cb->KsT = cb->V = nullptr;
  cb->A1 = take(planes);
  cb->part_var = reinterpret_cast<double*>(take((uint64_t)M->m * sg::parts_per_tile() * sg::PART_SPLIT * Nc * 8));
  cb->mean = reinterpret_cast<double*>(take((uint64_t)M->m * Nc * 8));
  cb->var = reinterpret_cast<double*>(take((uint64_t)M->m * Nc * 8));
  if (grad) {
    cb->GsT = reinterpret_cast<double*>(take((uint64_t)M->m * M->n16 * Nc * 8));
    cb->A2 = take(planes);
    cb->part_dvar = reinterpret_cast<double*>(take((uint64_t)M->m * sg::parts_per_tile() * sg::PART_SPLIT * Nc * M->d * 8));
    cb->part_s0 = reinterpret_cast<double*>(take((uint64_t)M->m * sg::parts_per_tile() * sg::PART_SPLIT * Nc * 8));
    cb->dmean = reinterpret_cast<double*>(take((uint64_t)M->m * Nc * M->d * 8));
    cb->dvar = reinterpret_cast<double*>(take((uint64_t)M->m * Nc * M->d * 8));
  } else {
    cb->GsT = cb->part_dvar = cb->dmean = cb->dvar = nullptr;
    cb->A2 = nullptr;
  }
}

static sg::GemmParams base_params(const bocf_model* M, int h, const ChunkBuffers& cb) {
  sg::GemmParams P;
  std::memset(&P, 0, sizeof(P));
  P.m = M->m;
  P.h = h;
  P.RT = (int)(cb.Nc / sg::TM);
  P.nct = M->ncts;
  P.KCH = M->KCH;
  P.n = M->n;
  P.Nc = cb.Nc;
  P.d = M->d;
  P.n16 = M->n16;
  P.n_pad = M->n_pad;
  P.hyp = M->hyp;
  P.cg = M->split_cg;
  P.np = sg::parts_per_tile();
  if (const char* env = std::getenv("BOCF_SPLIT_EXP")) P.exp = std::atoi(env);
  return P;
}

int launch_split_var(bocf_model* M, int h, const ChunkBuffers& cb, bool need_dvar, cudaStream_t st) {
  sg::GemmParams P = base_params(M, h, cb);
  P.A = cb.A1;
  P.B = M->B1;
  P.cs = M->cs1;
  P.tri = sg::TRI_K_LE_N;
  P.part_var = cb.part_var;
  P.A2 = need_dvar ? cb.A2 : nullptr;
  P.vq = M->vq;
  if (M->ncts < sg::parts_per_tile())      // parts without a column tile never write their partial
    BOCF_CUDA_OK(cudaMemsetAsync(cb.part_var, 0, sizeof(double) * M->m * sg::parts_per_tile() * sg::PART_SPLIT * cb.Nc, st));
  ProfScope ps("split_var_kernel", st);
  return sg::launch_s<sg::EPI_VAR>(M->S, P, st);
}

int launch_split_dvar(bocf_model* M, int h, const double* Xc, int64_t Nvalid, const ChunkBuffers& cb, cudaStream_t st) {
  sg::GemmParams P = base_params(M, h, cb);
  P.A = cb.A2;
  P.B = M->B2;
  P.cs = M->cs2;
  P.tri = sg::TRI_K_GE_N;
  P.part_dvar = cb.part_dvar;
  P.part_s0 = cb.part_s0;
  P.GsT = cb.GsT;
  P.Xc = Xc;
  P.Nvalid = Nvalid;
  P.Xs = M->Xs;
  if (M->ncts < sg::parts_per_tile()) {
    BOCF_CUDA_OK(cudaMemsetAsync(cb.part_dvar, 0, sizeof(double) * M->m * sg::parts_per_tile() * sg::PART_SPLIT * cb.Nc * M->d, st));
    BOCF_CUDA_OK(cudaMemsetAsync(cb.part_s0, 0, sizeof(double) * M->m * sg::parts_per_tile() * sg::PART_SPLIT * cb.Nc, st));
  }
  ProfScope ps("split_dvar_kernel", st);
  return sg::launch_dvar(M->S, M->d, P, st);
}

// Test entry: out (R x N) = A (R x K) * B (N x K)^T through the digit-plane machinery.  All pointers [dev] fp64 row-major.
int split_debug_gemm(const double* A, const double* B, int R, int N, int K, int S, int tri, double* out, cudaStream_t st) {
  if (S < 3 || S > 6 || R < 1 || N < 1 || K < 1) {
    set_error("split_debug_gemm: invalid arguments");
    return BOCF_ERR_INVALID;
  }
  const int NT = split_column_tile(S);
  const int RT = (int)round_up(ceil_div(R, sg::TM), 2), nct = (int)ceil_div(N, NT), KCH = (int)ceil_div(K, sg::KC);
  const int RpadA = RT * sg::TM, RpadB = nct * NT;
  uint8_t *pa = nullptr, *pb = nullptr;
  int *ea = nullptr, *eb = nullptr;
  double *rs = nullptr, *cs = nullptr, *tmp = nullptr;
  int rc = 0;
  do {
    if (cudaMalloc(reinterpret_cast<void**>(&pa), (size_t)RT * KCH * S * sg::TM * sg::KC) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&pb), (size_t)nct * KCH * (S * (S + 1) / 2) * NT * sg::KC) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&ea), sizeof(int) * RpadA) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&eb), sizeof(int) * RpadB) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&rs), sizeof(double) * RpadA) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&cs), sizeof(double) * RpadB) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&tmp), sizeof(double) * (size_t)RpadA * N) != cudaSuccess) {
      set_error("split_debug_gemm: out of device memory");
      rc = BOCF_ERR_CUDA;
      break;
    }
    const int base = -2 * (8 * S - 2) + 8 * (S - 1);
    if ((rc = sg::row_exps(A, 0, K, 1, R, K, RpadA, 1, ea, rs, nullptr, 0, st))) break;
    if ((rc = sg::row_exps(B, 0, K, 1, N, K, RpadB, 1, eb, cs, nullptr, base, st))) break;
    if ((rc = sg::pack_digits(S, A, 0, K, 1, R, K, ea, RpadA, sg::TM, RT, KCH, 1, pa, st))) break;
    int cg = 1;
    if (const char* env = std::getenv("BOCF_SPLIT_CG"))
      if (std::atoi(env) == 2) cg = 2;
    if (cg == 2) {
      if ((rc = sg::pack_digits_pair(S, B, 0, K, 1, N, K, eb, RpadB, NT, nct, KCH, 1, pb, st))) break;
    } else if ((rc = sg::pack_digits(S, B, 0, K, 1, N, K, eb, RpadB, NT, nct, KCH, 1, pb, st))) break;
    sg::GemmParams P;
    std::memset(&P, 0, sizeof(P));
    P.A = pa;
    P.B = pb;
    P.cs = cs;
    P.m = 1;
    P.h = 0;
    P.cg = cg;
    P.np = sg::parts_per_tile();
    P.RT = RT;
    P.nct = nct;
    P.KCH = KCH;
    P.n = (tri == sg::TRI_FULL) ? K : (K < N ? K : N);
    if (tri != sg::TRI_FULL) P.n = K;
    P.tri = tri;
    P.Nc = RpadA;
    P.raw_out = tmp;
    P.raw_rs = rs;
    P.ldo = N;
    if (const char* env = std::getenv("BOCF_SPLIT_EXP")) P.exp = std::atoi(env);
    {
      ProfScope ps("split_raw_kernel", st);
      rc = sg::launch_s<sg::EPI_RAW>(S, P, st);
    }
    if (rc) break;
    if (cudaMemcpyAsync(out, tmp, sizeof(double) * (size_t)R * N, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      set_error(std::string("split_debug_gemm: ") + cudaGetErrorString(cudaGetLastError()));
      rc = BOCF_ERR_CUDA;
    }
  } while (0);
  cudaFree(pa);
  cudaFree(pb);
  cudaFree(ea);
  cudaFree(eb);
  cudaFree(rs);
  cudaFree(cs);
  cudaFree(tmp);
  return rc;
}

}  // namespace bocf
