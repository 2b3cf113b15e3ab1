// gemm_f64.cuh -- fp64 tensor-core (DMMA.8x8x4) tile engine.
//
// One CTA of WM x WN warps computes a (8*MI*WM) x (32*WN) fp64 tile
//     acc[m][n] = sum_{k in [k_begin,k_end)} A(m,k) * B(k,n)
// with both operands streamed global -> shared by a 3-stage cp.async ring (BK = 16) and fed to
// mma.sync.m8n8k4.f64.  Each warp owns an (8*MI) x 32 sub-tile = MI x 4 MMA tiles (8*MI fp64 accumulators per
// thread).  Tile shapes in use:
//     Tile<2,4,8> = 128 x 128, 256 threads, 64 acc/thread   Cholesky panel / trailing updates, blocked inverse
//     Tile<2,2,8> = 128 x  64, 128 threads, 64 acc/thread   candidate-side contractions, two CTAs per SM: their
//                   barriers are decoupled (+11 % over one 256-thread CTA) and the narrow 64-column tiles waste
//                   fewer MMAs on the structurally-zero triangle of the factor's diagonal blocks.  Measured
//                   alternatives: 256x64/256 thr (1 CTA) -12 %; 128x64 with 32x32 warp tiles (16 warps/SM) -3 %.
// Operand layouts are template switches so the same loop serves
//     V  = K* . Linv^T      (A k-major, B n-major)         posterior variance (trsm-as-gemm)
//     Wt = V  . Linv        (A m-major, B k-major)         variance-gradient weights
// All matrix dimensions are padded by the caller so no bounds checks are needed inside the loop;
// k_begin / k_end are multiples of BK.  Shared-memory rows are padded by 4 doubles which makes
// every 64-bit fragment load bank-conflict free (row stride == 4 mod 16 doubles).
#pragma once
#include "common.cuh"

namespace bocf {
namespace gemm {

constexpr int BK = 16;
constexpr int STAGES = 3;
constexpr int LD_K = BK + 4;    // operand rows hold BK doubles (k contiguous)

template <int WM_, int WN_, int MI_>
struct Tile {
  static constexpr int WM = WM_, WN = WN_, MI = MI_;
  static constexpr int NTHREADS = 32 * WM * WN;
  static constexpr int WTM = 8 * MI;          // warp tile rows
  static constexpr int BM = WTM * WM;
  static constexpr int BN = 32 * WN;
  static constexpr int LDA_X = BM + 4;     // k-major A rows hold BM doubles
  static constexpr int LDB_X = BN + 4;     // k-major B rows hold BN doubles
  static constexpr int A_DOUBLES = (BM * LD_K > BK * LDA_X) ? BM * LD_K : BK * LDA_X;
  static constexpr int B_DOUBLES = (BN * LD_K > BK * LDB_X) ? BN * LD_K : BK * LDB_X;
  static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);
};
using Tile128 = Tile<2, 4, 8>;     // 128 x 128, 256 threads, 122 880 B, 1 CTA / SM
using TilePost = Tile<2, 2, 8>;    // 128 x  64, 128 threads,  92 160 B, 2 CTAs / SM

// ---- global -> shared stage loads ---------------------------------------------------------------
// KMAJOR == false: element (x,k) at P[x*ld + k]  (k contiguous)   -> smem s[x*LD_K + k]
// KMAJOR == true : element (x,k) at P[k*ld + x]  (x contiguous)   -> smem s[k*LDX + x]
template <bool KMAJOR, int ROWS, int LDX, int NTHR>
__device__ __forceinline__ void load_operand(double* s, const double* __restrict__ P, int64_t ld, int k0,
                                             int tid) {
  constexpr int CHUNKS = ROWS * BK / 2;           // 16-byte chunks per stage
  static_assert(CHUNKS % NTHR == 0, "tile must split evenly over the CTA");
  if (KMAJOR) {
    constexpr int PER_ROW = ROWS / 2;
#pragma unroll
    for (int r = 0; r < CHUNKS / NTHR; ++r) {
      int c = tid + NTHR * r;
      int krow = c / PER_ROW, xc = (c % PER_ROW) * 2;
      cp_async16(s + krow * LDX + xc, P + (int64_t)(k0 + krow) * ld + xc);
    }
  } else {
#pragma unroll
    for (int r = 0; r < CHUNKS / NTHR; ++r) {
      int c = tid + NTHR * r;
      int row = c >> 3, kc = (c & 7) * 2;
      cp_async16(s + row * LD_K + kc, P + (int64_t)row * ld + k0 + kc);
    }
  }
}

// Triangular B operand: which (k, n) entries are structurally non-zero, in GLOBAL indices (n_glob = tri_col0 + n).
//   TRI_NONE        dense
//   TRI_K_LE_N      B(k,n) != 0 only for k <= n_glob   (V = K* Linv^T: Linv[a][b] with b <= a)
//   TRI_K_GE_N      B(k,n) != 0 only for k >= n_glob   (Wt = V Linv:   Linv[a][b] with a >= b)
// Inside the k-tiles that straddle the diagonal the MMAs of all-zero 8-column groups are skipped (warp-uniform).
enum { TRI_NONE = 0, TRI_K_LE_N = 1, TRI_K_GE_N = 2 };

// acc[i][j][e]: row = wm*WTM + 8*i + g, col = wn*32 + 8*j + 2*t + e   (lane = 4*g + t, warp = wm*WN + wn)
template <class T, bool A_KMAJOR, bool B_KMAJOR, int TRI = TRI_NONE>
__device__ __forceinline__ void mainloop(double (&acc)[T::MI][4][2], const double* __restrict__ A, int64_t lda,
                                         const double* __restrict__ B, int64_t ldb, int k_begin, int k_end,
                                         double* smem, int tri_col0 = 0) {
  constexpr int MI = T::MI;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int mbase = (warp / T::WN) * T::WTM, nbase = (warp % T::WN) * 32;
  const int KT = (k_end - k_begin) / BK;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) {
      double* st = smem + s * T::STAGE_DOUBLES;
      load_operand<A_KMAJOR, T::BM, T::LDA_X, T::NTHREADS>(st, A, lda, k_begin + s * BK, tid);
      load_operand<B_KMAJOR, T::BN, T::LDB_X, T::NTHREADS>(st + T::A_DOUBLES, B, ldb, k_begin + s * BK, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nxt = kt + STAGES - 1;
      if (nxt < KT) {
        double* st = smem + (nxt % STAGES) * T::STAGE_DOUBLES;
        load_operand<A_KMAJOR, T::BM, T::LDA_X, T::NTHREADS>(st, A, lda, k_begin + nxt * BK, tid);
        load_operand<B_KMAJOR, T::BN, T::LDB_X, T::NTHREADS>(st + T::A_DOUBLES, B, ldb, k_begin + nxt * BK, tid);
      }
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * T::STAGE_DOUBLES;
    const double* sB = sA + T::A_DOUBLES;
    const int k0 = k_begin + kt * BK;
    // does this k-tile straddle the diagonal of this warp's columns?  (warp-uniform)
    bool diag = false;
    if (TRI == TRI_K_LE_N) diag = (k0 + BK - 1) > (tri_col0 + nbase);            // some k exceeds the first column
    if (TRI == TRI_K_GE_N) diag = k0 < (tri_col0 + nbase + 31);                  // some k precedes the last column
    if (!diag) {
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        double a[MI], b[4];
#pragma unroll
        for (int i = 0; i < MI; ++i)
          a[i] = A_KMAJOR ? sA[(kk + t) * T::LDA_X + mbase + 8 * i + g] : sA[(mbase + 8 * i + g) * LD_K + kk + t];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          b[j] = B_KMAJOR ? sB[(kk + t) * T::LDB_X + nbase + 8 * j + g] : sB[(nbase + 8 * j + g) * LD_K + kk + t];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        double a[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i)
          a[i] = A_KMAJOR ? sA[(kk + t) * T::LDA_X + mbase + 8 * i + g] : sA[(mbase + 8 * i + g) * LD_K + kk + t];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c_lo = tri_col0 + nbase + 8 * j;     // global columns c_lo .. c_lo+7, k-step rows k0+kk .. +3
          const bool need = (TRI == TRI_K_LE_N) ? (k0 + kk <= c_lo + 7) : (k0 + kk + 3 >= c_lo);
          if (need) {
            const double b = B_KMAJOR ? sB[(kk + t) * T::LDB_X + nbase + 8 * j + g] : sB[(nbase + 8 * j + g) * LD_K + kk + t];
#pragma unroll
            for (int i = 0; i < MI; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b);
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // pipeline smem is free for the epilogue after this
}

template <int MI>
__device__ __forceinline__ void zero_acc(double (&acc)[MI][4][2]) {
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

}  // namespace gemm
}  // namespace bocf
