// TMA-fed versions of the three hot kernels (see gemm_tma.cuh for the pipeline).
// Every kernel is 3 warpgroups: the last one produces (TMA), the first two consume (DMMA)
// and run the fused epilogue.  Tiles whose valid width is <= 32 / <= 64 columns (the
// K mod 128 border: 32 columns at K = 2080) run with an 8x1 / 4x2 warp grid (NW = 1 / 2)
// so they cost a quarter / half of a full tile instead of a whole one.
#pragma once
#include "gemm_tma.cuh"
#include "kernels.cuh"

namespace scfgp {

template <int MI>
__device__ __forceinline__ void zero_accN(double (&acc)[MI][4][2]) {
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// ------------------------------------------------------------------------------------
// Gram accumulation  C[m,n] += sum_r A[r,m] B[r,n]   (A, B: row slabs, reduction over rows)
//   (b) PhiT.Phi and PsiT.Phi.  items: (m block, n block, transpose flag, -).
//   Border row block of the lower triangle: computed with the roles swapped (the narrow
//   side must be the n side) and stored transposed.
// ------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void syrk_consume(const TmaPipe& pipe, int total, int m0, int n0, int m_valid, int n_valid,
                                             bool transpose, double* __restrict__ C, long long ldc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[2 * NW][4][2];
  zero_accN<2 * NW>(acc);
  tma_consumer<false, false, NW>(acc, pipe, total);
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / NW) * (16 * NW), wn0 = (warp % NW) * 32;
#pragma unroll
  for (int i = 0; i < 2 * NW; ++i) {
    const int r = m0 + wm0 + 8 * i + g;
    if (r >= m_valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + wn0 + 8 * j + 2 * t;
      if (!transpose) {
        double* p = C + (long long)r * ldc + c;
        if (c + 1 < n_valid) {
          double2 v = *reinterpret_cast<double2*>(p);
          v.x += acc[i][j][0]; v.y += acc[i][j][1];
          *reinterpret_cast<double2*>(p) = v;
        } else if (c < n_valid) {
          *p += acc[i][j][0];
        }
      } else {
        if (c < n_valid) C[(long long)c * ldc + r] += acc[i][j][0];
        if (c + 1 < n_valid) C[(long long)(c + 1) * ldc + r] += acc[i][j][1];
      }
    }
  }
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
k_syrk_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int depth,
           int m_valid, int n_valid, double* __restrict__ C, long long ldc, const int4* __restrict__ items) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const TmaPipe pipe = tma_pipe_init(smem_raw);
  const int4 it = items[blockIdx.x];
  const int m0 = it.x * 128, n0 = it.y * 128;
  const int nw = tma_nw(n_valid - n0);
  const int dd[1] = {depth};
  if ((threadIdx.x >> 5) >= TMA_CONSUMER_WARPS) {
    BoxRows ra, rb;
#pragma unroll
    for (int b = 0; b < 8; ++b) { ra.row0[b] = m0 + 16 * b; rb.row0[b] = n0 + 16 * b; }
    // with swapped roles the A side reads the n-side slab's columns and vice versa:
    // both slabs hold the same column space, only the maps differ (Psi vs Phi)
    const CUtensorMap* ma[1] = {it.z ? &mapB : &mapA};
    const CUtensorMap* mb[1] = {it.z ? &mapA : &mapB};
    tma_producer<false, false>(pipe, 1, ma, mb, ra, rb, 2 * nw, dd);
    return;
  }
  tma_consumer_regs();
  const int total = tma_total_iters(1, dd);
  if (nw == 1) syrk_consume<1>(pipe, total, m0, n0, m_valid, n_valid, it.z != 0, C, ldc);
  else if (nw == 2) syrk_consume<2>(pipe, total, m0, n0, m_valid, n_valid, it.z != 0, C, ldc);
  else syrk_consume<4>(pipe, total, m0, n0, m_valid, n_valid, it.z != 0, C, ldc);
}

// ------------------------------------------------------------------------------------
// (a6) T = Phi.LiT row sums of squares                                   SCFGP.py:112
// ------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void tgemm_consume(const TmaPipe& pipe, int total, int m0, int rows, int kb,
                                              double* __restrict__ vpart, long long ldv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[2 * NW][4][2];
  zero_accN<2 * NW>(acc);
  tma_consumer<true, true, NW>(acc, pipe, total);

  tma_consumer_sync();  // every consumer is done with the stage buffers: reuse them for the reduction
  double* red = reinterpret_cast<double*>(pipe.stages);  // [NW][128]
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / NW) * (16 * NW), wni = warp % NW;
#pragma unroll
  for (int i = 0; i < 2 * NW; ++i) {
    double sq = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) sq += acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1];
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    if (t == 0) red[wni * 128 + wm0 + 8 * i + pi8(g)] = sq;
  }
  tma_consumer_sync();
  if (threadIdx.x < 128) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += red[w * 128 + threadIdx.x];
    const int r = m0 + threadIdx.x;
    if (r < rows) vpart[(long long)kb * ldv + r] = v;
  }
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
k_tgemm_tma(const __grid_constant__ CUtensorMap mapPhi, const __grid_constant__ CUtensorMap mapLi, int rows, int K,
            double* __restrict__ vpart, long long ldv, const int4* __restrict__ items) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int4 it = items[blockIdx.x];
  const int m0 = it.x * 128, n0 = it.y * 128;
  if (m0 >= rows) return;  // short last chunk: whole CTA leaves before any barrier
  const TmaPipe pipe = tma_pipe_init(smem_raw);
  const int nw = tma_nw(K - n0);
  const int dd[1] = {min(K, n0 + 128)};
  if ((threadIdx.x >> 5) >= TMA_CONSUMER_WARPS) {
    BoxRows ra, rb;
#pragma unroll
    for (int b = 0; b < 8; ++b) { ra.row0[b] = m0 + 16 * b; rb.row0[b] = n0 + 16 * b; }
    const CUtensorMap* ma[1] = {&mapPhi};
    const CUtensorMap* mb[1] = {&mapLi};
    tma_producer<true, true>(pipe, 1, ma, mb, ra, rb, 2 * nw, dd);
    return;
  }
  tma_consumer_regs();
  const int total = tma_total_iters(1, dd);
  if (nw == 1) tgemm_consume<1>(pipe, total, m0, rows, it.y, vpart, ldv);
  else if (nw == 2) tgemm_consume<2>(pipe, total, m0, rows, it.y, vpart, ldv);
  else tgemm_consume<4>(pipe, total, m0, rows, it.y, vpart, ldv);
}

// ------------------------------------------------------------------------------------
// (c) backward, fused with the trig derivative (see k_backward in kernels.cuh)
//   B tile rows: warp column wc owns [16 cos | 16 sin] of features f0 + 16 wc .. +15
// ------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void backward_consume(const TmaPipe& pipe, int total, int m0, int f0, int rows, int J,
                                                 const double* __restrict__ Phi, long long ldp,
                                                 const double* __restrict__ y, const double* __restrict__ gmu,
                                                 const double* __restrict__ pbar, const double* __restrict__ alpha,
                                                 double* __restrict__ Zbar, long long ldz, double* sh) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[2 * NW][4][2];
  zero_accN<2 * NW>(acc);
  tma_consumer<true, true, NW>(acc, pipe, total);
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / NW) * (16 * NW), wni = warp % NW;
  double sacc = 0;
#pragma unroll
  for (int i = 0; i < 2 * NW; ++i) {
    const int r = m0 + wm0 + 8 * i + pi8(g);
    if (r >= rows) continue;
    const double yn = y[r], gm = gmu[r];
    const double* prow = Phi + (long long)r * ldp;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int f = f0 + wni * 16 + 8 * jj + pi8(2 * t + e);
        if (f >= J) continue;
        const double pc = acc[i][jj][e] + yn * pbar[f] + gm * alpha[f];
        const double ps = acc[i][jj + 2][e] + yn * pbar[J + f] + gm * alpha[J + f];
        const double phc = prow[f], phs = prow[J + f];
        Zbar[(long long)r * ldz + f] = -pc * phs + ps * phc;
        sacc += pc * phc + ps * phs;
      }
    }
  }
  sacc = warp_sum(sacc);
  if (lane == 0) sh[warp] = sacc;
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
k_backward_tma(const __grid_constant__ CUtensorMap mapPhi, const __grid_constant__ CUtensorMap mapPsi,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2,
               const double* __restrict__ Phi, long long ldp, int rows, int K, int J,
               const double* __restrict__ y, const double* __restrict__ gmu, const double* __restrict__ pbar,
               const double* __restrict__ alpha, double* __restrict__ Zbar, long long ldz,
               double* __restrict__ slots, const int4* __restrict__ items) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ double sh[TMA_CONSUMER_WARPS];
  const TmaPipe pipe = tma_pipe_init(smem_raw);
  const int4 it = items[blockIdx.x];
  const int m0 = it.x * 128;
  const int f0 = it.y * 64;
  const int vf = J - f0;  // valid features of this block (<= 64 useful)
  const int nw = vf <= 16 ? 1 : (vf <= 32 ? 2 : 4);
  const int dd[2] = {K, K};
  if ((threadIdx.x >> 5) >= TMA_CONSUMER_WARPS) {
    BoxRows ra, rb;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      ra.row0[b] = m0 + 16 * b;
      rb.row0[b] = ((b & 1) ? J : 0) + f0 + (b >> 1) * 16;
    }
    const CUtensorMap* ma[2] = {&mapPhi, &mapPsi};
    const CUtensorMap* mb[2] = {&mapA2, &mapB2};
    tma_producer<true, true>(pipe, 2, ma, mb, ra, rb, 2 * nw, dd);
    return;
  }
  tma_consumer_regs();
  const int total = tma_total_iters(2, dd);
  if (nw == 1) backward_consume<1>(pipe, total, m0, f0, rows, J, Phi, ldp, y, gmu, pbar, alpha, Zbar, ldz, sh);
  else if (nw == 2) backward_consume<2>(pipe, total, m0, f0, rows, J, Phi, ldp, y, gmu, pbar, alpha, Zbar, ldz, sh);
  else backward_consume<4>(pipe, total, m0, f0, rows, J, Phi, ldp, y, gmu, pbar, alpha, Zbar, ldz, sh);
  tma_consumer_sync();
  if (threadIdx.x == 0) {
    double v = 0;
    for (int w = 0; w < TMA_CONSUMER_WARPS; ++w) v += sh[w];
    slots[blockIdx.x] += v;
  }
}

// generic C = A op B through the TMA pipeline (self-test hook): NT (AK=BKM=true) / TN.
// `nw_force` (1, 2, 4) exercises the narrow warp grids on the first 32 / 64 / 128 columns.
template <bool AK, bool BKM, int NW>
__device__ __forceinline__ void dbg_consume(const TmaPipe& pipe, int total, int m0, int n0, int m, int n,
                                            double* __restrict__ C, long long ldc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[2 * NW][4][2];
  zero_accN<2 * NW>(acc);
  tma_consumer<AK, BKM, NW>(acc, pipe, total);
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / NW) * (16 * NW), wn0 = (warp % NW) * 32;
#pragma unroll
  for (int i = 0; i < 2 * NW; ++i) {
    const int r = m0 + tma_idx<AK>(wm0, i, g);
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n0 + tma_idx<BKM>(wn0, j, 2 * t + e);
        if (c < n) C[(long long)r * ldc + c] = acc[i][j][e];
      }
  }
}

template <bool AK, bool BKM>
__global__ void __launch_bounds__(TMA_THREADS, 1)
k_gemm_tma_dbg(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int depth, int m, int n,
               double* __restrict__ C, long long ldc) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const TmaPipe pipe = tma_pipe_init(smem_raw);
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
  const int nw = tma_nw(n - n0);
  const int dd[1] = {depth};
  if ((threadIdx.x >> 5) >= TMA_CONSUMER_WARPS) {
    BoxRows ra, rb;
#pragma unroll
    for (int b = 0; b < 8; ++b) { ra.row0[b] = m0 + 16 * b; rb.row0[b] = n0 + 16 * b; }
    const CUtensorMap* ma[1] = {&mapA};
    const CUtensorMap* mb[1] = {&mapB};
    tma_producer<AK, BKM>(pipe, 1, ma, mb, ra, rb, 2 * nw, dd);
    return;
  }
  tma_consumer_regs();
  const int total = tma_total_iters(1, dd);
  if (nw == 1) dbg_consume<AK, BKM, 1>(pipe, total, m0, n0, m, n, C, ldc);
  else if (nw == 2) dbg_consume<AK, BKM, 2>(pipe, total, m0, n0, m, n, C, ldc);
  else dbg_consume<AK, BKM, 4>(pipe, total, m0, n0, m, n, C, ldc);
}

}  // namespace scfgp
