// FP64 tensor-core (DMMA) tile mainloop for sm_100a.
//
// tcgen05.mma has no f64 kind, so binary64 matrix products on Blackwell go through the
// warp-level DMMA path: `mma.sync.aligned.m8n8k4.f64` (SASS: DMMA.8x8x4; the larger
// PTX shapes m16n8k{4,8,16} are split by ptxas into exactly this instruction on
// sm_100a, so the kernels issue it directly and own the fragment layout).
//
// One mainloop serves every product of the hot path.  Operands are described by the
// storage order of their reduction dimension:
//   K-major  : element (row, k) at ptr[row*ld + k]      (k contiguous)
//   MN-major : element (row, k) at ptr[k*ld + row]      (row contiguous)
// so NT = (K,K), TN = (MN,MN), NN = (K,MN).  Tiles are staged with cp.async through a
// multi-stage shared-memory ring; every global access is predicated (zero fill), so
// callers never pad their matrices.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scfgp {

constexpr int BK = 16;  // default reduction depth per pipeline stage (doubles); TileCfg::BKT overrides it

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// 16-byte / 8-byte async copies with zero fill of the bytes beyond `src_bytes`.
__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Shared-memory tile of one operand for one stage.  The +4 padding makes the DMMA
// fragment reads (8 rows x 4 k, one double per lane) hit 16 distinct 8-byte banks per
// half warp in both storage orders (LD = 4 mod 16).
template <int ROWS, bool KMAJOR, int BKT = BK>
struct TileLayout {
  static constexpr int LD = KMAJOR ? (BKT + 4) : (ROWS + 4);
  static constexpr int SIZE = KMAJOR ? ROWS * LD : BKT * LD;
  __device__ __forceinline__ static int idx(int row, int k) { return KMAJOR ? row * LD + k : k * LD + row; }
};

// Maps a tile-local row to a row of the stored matrix (or -1: zero fill).
struct IdentMap {
  int base;   // first row of this tile
  int rows;   // valid rows of the matrix
  __device__ __forceinline__ int operator()(int r) const {
    int g = base + r;
    return g < rows ? g : -1;
  }
};

// Global -> shared copy of one ROWS x BK operand tile at reduction offset k0.
// VEC = 2: 16-byte copies (needs 16B-aligned ptr and even ld); VEC = 1: 8-byte copies.
template <int ROWS, bool KMAJOR, int VEC, int NT, int BKT, class RowMap>
__device__ __forceinline__ void load_tile(double* __restrict__ smem, const double* __restrict__ ptr,
                                          long long ld, const RowMap& map, int k0, int depth, int tid) {
  using L = TileLayout<ROWS, KMAJOR, BKT>;
  if constexpr (KMAJOR) {
    constexpr int CPR = BKT / VEC;
    constexpr int TOTAL = ROWS * CPR;
#pragma unroll
    for (int it = 0; it < (TOTAL + NT - 1) / NT; ++it) {
      int c = it * NT + tid;
      if (TOTAL % NT != 0 && c >= TOTAL) break;
      int r = c / CPR, kc = (c % CPR) * VEC;
      int gr = map(r);
      int k = k0 + kc;
      int nvalid = (gr >= 0) ? min(max(depth - k, 0), VEC) : 0;
      const double* src = nvalid > 0 ? ptr + (long long)gr * ld + k : ptr;
      if constexpr (VEC == 2) cp_async16(smem + r * L::LD + kc, src, nvalid * 8);
      else cp_async8(smem + r * L::LD + kc, src, nvalid * 8);
    }
  } else {
    constexpr int CPR = ROWS / VEC;
    constexpr int TOTAL = BKT * CPR;
#pragma unroll
    for (int it = 0; it < (TOTAL + NT - 1) / NT; ++it) {
      int c = it * NT + tid;
      if (TOTAL % NT != 0 && c >= TOTAL) break;
      int kr = c / CPR, m = (c % CPR) * VEC;
      int k = k0 + kr;
      // MN-major operands are only used with contiguous (identity) row maps
      int g0 = map.base + m;
      int nvalid = (k < depth) ? min(max(map.rows - g0, 0), VEC) : 0;
      const double* src = nvalid > 0 ? ptr + (long long)k * ld + g0 : ptr;
      if constexpr (VEC == 2) cp_async16(smem + kr * L::LD + m, src, nvalid * 8);
      else cp_async8(smem + kr * L::LD + m, src, nvalid * 8);
    }
  }
}

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int BKT_ = BK>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, BKT = BKT_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int NWARPS = WARPS_M * WARPS_N;
  static constexpr int NT = NWARPS * 32;
  static constexpr int MI = WM / 8, NI = WN / 8;
};

template <class Cfg, bool AK, bool BKM>
struct SmemSize {
  static constexpr int STAGE = TileLayout<Cfg::BM, AK, Cfg::BKT>::SIZE + TileLayout<Cfg::BN, BKM, Cfg::BKT>::SIZE;
  static constexpr int DOUBLES = STAGE * Cfg::STAGES;
  static constexpr size_t BYTES = (size_t)DOUBLES * sizeof(double);
};

// acc[i][j][e] = C[wm0 + 8i + g][wn0 + 8j + 2t + e],  g = lane>>2, t = lane&3
// BSUM: bsum[j] += the B fragment values this thread feeds to the DMMAs, i.e. after the loop
//   bsum[j] = sum over k = t (mod 4) of B(wn0 + 8j + g, k): the sum over the four lanes t of a group g is the
//   column sum of the B operand over the whole reduction range (dphi = colsum(Zbar) inside the dW contraction)
template <class Cfg, bool AK, bool BKM, int AVEC, int BVEC, bool BSUM = false, class AMap, class BMap>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[Cfg::MI][Cfg::NI][2],
                                              const double* __restrict__ A, long long lda, const AMap& amap,
                                              const double* __restrict__ B, long long ldb, const BMap& bmap,
                                              int depth, double* smem, double* bsum = nullptr) {
  constexpr int BKT = Cfg::BKT;
  using LA = TileLayout<Cfg::BM, AK, BKT>;
  using LB = TileLayout<Cfg::BN, BKM, BKT>;
  constexpr int STAGE = LA::SIZE + LB::SIZE;
  constexpr int STAGES = Cfg::STAGES;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM;
  const int wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
  const int nk = (depth + BKT - 1) / BKT;

  __syncthreads();  // smem may still be in use by a previous segment / epilogue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      double* st = smem + s * STAGE;
      load_tile<Cfg::BM, AK, AVEC, Cfg::NT, BKT>(st, A, lda, amap, s * BKT, depth, tid);
      load_tile<Cfg::BN, BKM, BVEC, Cfg::NT, BKT>(st + LA::SIZE, B, ldb, bmap, s * BKT, depth, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int kn = kt + STAGES - 1;
      if (kn < nk) {
        double* st = smem + (kn % STAGES) * STAGE;
        load_tile<Cfg::BM, AK, AVEC, Cfg::NT, BKT>(st, A, lda, amap, kn * BKT, depth, tid);
        load_tile<Cfg::BN, BKM, BVEC, Cfg::NT, BKT>(st + LA::SIZE, B, ldb, bmap, kn * BKT, depth, tid);
      }
      cp_async_commit();
    }
    const double* As = smem + (kt % STAGES) * STAGE;
    const double* Bs = As + LA::SIZE;
#pragma unroll
    for (int kk = 0; kk < BKT / 4; ++kk) {
      double a[Cfg::MI], b[Cfg::NI];
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i) a[i] = As[LA::idx(wm0 + 8 * i + g, kk * 4 + t)];
#pragma unroll
      for (int j = 0; j < Cfg::NI; ++j) b[j] = Bs[LB::idx(wn0 + 8 * j + g, kk * 4 + t)];
      if constexpr (BSUM) {
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) bsum[j] += b[j];
      }
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

template <class Cfg>
__device__ __forceinline__ void zero_acc(double (&acc)[Cfg::MI][Cfg::NI][2]) {
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

}  // namespace scfgp
