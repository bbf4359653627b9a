// Kernels of the SCFGP marginal-likelihood hot path (sm_100a, FP64 DMMA).
// Math: SURVEY.md Appendix A, restating SCFGP/SCFGP.py:92-148 of the reference.
#pragma once
#include <type_traits>

#include "gemm_dmma.cuh"

namespace scfgp {

// ------------------------------------------------------------------------------------
// tile shapes
// ------------------------------------------------------------------------------------
using CfgBig = TileCfg<128, 128, 64, 32, 3>;   // 8 warps, 64 accumulators / thread
using CfgTall = TileCfg<128, 64, 32, 32, 3>;   // 8 warps, panels
// feature map and dW contraction: depth-8 stages keep their shared memory (36 KB / 26 KB) small
// enough to share an SM with a resident tcgen05 CTA of the emulated path (186 KB at 5 digits)
using CfgFeat = TileCfg<128, 64, 32, 32, 2, 8>;   // 8 warps, 2 stages
using CfgWide = TileCfg<64, 128, 32, 32, 3>;   // 8 warps, small dense NN products
using CfgDw = TileCfg<64, 128, 32, 32, 3, 16>;  // 8 warps, D x J contraction (dW): 3 stages of depth 16 (77 KB)
using CfgSq = TileCfg<64, 64, 32, 32, 3>;      // 4 warps, 64 x 64 tiles: block-recursive triangular inverse

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------
// sin and cos for the feature map (SCFGP.py:98-102).  Straight-line code for |z| <= kFastTrigMax: k = rint(z 2/pi)
// by the 1.5 * 2^52 magic add, two FMA steps of Cody-Waite reduction on a two-part pi/2 (|k| < 2^20, so k * lo(pi/2)
// is far below the rounding of the first step), the fdlibm kernel polynomials on [-pi/4, pi/4] and the quadrant
// fix-up by selects.  Measured against long-double libm over |z| <= 1e6: max |error| 1.74e-16 for both (< 1 ulp of
// 1; tools/sincos_check.c).  No branch, no call: the compiler interleaves the independent evaluations of an
// unrolled epilogue (the library sincos carries a slow-path branch per call that serialises them).  Larger
// arguments take the library function through one out-of-line copy.
// ------------------------------------------------------------------------------------
constexpr double kFastTrigMax = 1.0e6;
__device__ __forceinline__ void sincos_fast(double z, double& sn, double& cs) {
  const double kd = fma(z, 0.63661977236758138, 6755399441055744.0);
  const int q = __double2loint(kd);
  const double k = kd - 6755399441055744.0;
  double r = fma(-k, 1.5707963267948966, z);
  r = fma(-k, 6.123233995736766e-17, r);
  const double w = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, w, -2.50507602534068634195e-08);
  ps = fma(ps, w, 2.75573137070700676789e-06);
  ps = fma(ps, w, -1.98412698298579493134e-04);
  ps = fma(ps, w, 8.33333333332248946124e-03);
  ps = fma(ps, w, -1.66666666666666324348e-01);
  const double s = fma(r * w, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, w, 2.08757232129817482790e-09);
  pc = fma(pc, w, -2.75573143513906633035e-07);
  pc = fma(pc, w, 2.48015872894767294178e-05);
  pc = fma(pc, w, -1.38888888888741095749e-03);
  pc = fma(pc, w, 4.16666666666666019037e-02);
  const double c = fma(w * w, pc, fma(-0.5, w, 1.0));
  // quadrant q mod 4: (s, c), (c, -s), (-s, -c), (-c, s); the sign flips are integer operations on the high word
  const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
  sn = __hiloint2double(__double2hiint(a) ^ ((q & 2) << 30), __double2loint(a));
  cs = __hiloint2double(__double2hiint(b) ^ (((q + 1) & 2) << 30), __double2loint(b));
}
__device__ __noinline__ void sincos_lib(double z, double* sn, double* cs) { sincos(z, sn, cs); }
template <bool SLOW>
__device__ __forceinline__ void sincos_sel(double z, double& sn, double& cs) {
  if constexpr (SLOW) sincos_lib(z, &sn, &cs);
  else sincos_fast(z, sn, cs);
}

// ------------------------------------------------------------------------------------
// generic GEMM kernel:  C = beta*C + alpha * (A op B)   over a list of tiles
//   item.x = tile row block, item.y = tile col block, item.z = reduction start,
//   item.w = reduction length
// ------------------------------------------------------------------------------------
template <class Cfg, bool AK, bool BKM, int AVEC, int BVEC>
__global__ void __launch_bounds__(Cfg::NT, 1)
k_gemm(const double* __restrict__ A, long long lda, int a_rows,
       const double* __restrict__ B, long long ldb, int b_rows,
       double* C, long long ldc, double alpha, double beta,
       const int4* __restrict__ items, int depth_arg) {
  // C may alias A (in-place panel solve): every read of the tile's operands completes
  // in the mainloop before the epilogue stores.
  extern __shared__ __align__(16) double smem[];
  const int4 it = items[blockIdx.x];
  const int m0 = it.x * Cfg::BM, n0 = it.y * Cfg::BN;
  const int k0 = it.z, depth = it.w >= 0 ? it.w : depth_arg - k0;
  double acc[Cfg::MI][Cfg::NI][2];
  zero_acc<Cfg>(acc);
  IdentMap am{m0, a_rows}, bm{n0, b_rows};
  const double* Ap = AK ? A + k0 : A + (long long)k0 * lda;
  const double* Bp = BKM ? B + k0 : B + (long long)k0 * ldb;
  gemm_mainloop<Cfg, AK, BKM, AVEC, BVEC>(acc, Ap, lda, am, Bp, ldb, bm, depth, smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
    int r = m0 + wm0 + 8 * i + g;
    if (r >= a_rows) continue;
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int c = n0 + wn0 + 8 * j + 2 * t + e;
        if (c < b_rows) {
          double* p = C + (long long)r * ldc + c;
          double v = alpha * acc[i][j][e];
          if (beta != 0.0) v += beta * (*p);
          *p = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// dW partials:  part[split][d][j] += sum over the split's rows of X[r,d] Zbar[r,j]   (A.5)
//   The D x J output has only a handful of tiles, so the row reduction is cut into
//   gridDim.y splits (one partial buffer each; summed in fixed order afterwards).
//   items: (d block of 64, j block of 128)
// ------------------------------------------------------------------------------------
template <int AVEC>
__global__ void __launch_bounds__(CfgDw::NT, 2)
k_dw_splitk(const double* __restrict__ X, long long ldx, int D, const double* __restrict__ Z, long long ldz, int J,
            int rows, int split_len, double* __restrict__ part, long long ldc, long long part_stride,
            const int4* __restrict__ items, double* __restrict__ colpart, long long ldcol) {
  // colpart (optional): colpart[split][j] += sum over the split's rows of Zbar[r, j]  (dphi = colsum(Zbar), A.5) --
  // the B fragments the DMMAs consume anyway are summed on the way (CTAs of the first d block only)
  using Cfg = CfgDw;
  extern __shared__ __align__(16) double smem[];
  const int4 it = items[blockIdx.x];
  const int m0 = it.x * Cfg::BM, n0 = it.y * Cfg::BN;
  const int k0 = blockIdx.y * split_len;
  const int depth = min(split_len, rows - k0);
  if (depth <= 0) return;
  double acc[Cfg::MI][Cfg::NI][2];
  zero_acc<Cfg>(acc);
  double bsum[Cfg::NI];
#pragma unroll
  for (int j = 0; j < Cfg::NI; ++j) bsum[j] = 0.0;
  IdentMap am{m0, D}, bm{n0, J};
  gemm_mainloop<Cfg, false, false, AVEC, 2, true>(acc, X + (long long)k0 * ldx, ldx, am, Z + (long long)k0 * ldz, ldz,
                                                  bm, depth, smem, bsum);
  double* C = part + (long long)blockIdx.y * part_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
    const int r = m0 + wm0 + 8 * i + g;
    if (r >= D) continue;
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n0 + wn0 + 8 * j + 2 * t + e;
        if (c < J) C[(long long)r * ldc + c] += acc[i][j][e];
      }
  }
  if (colpart != nullptr && it.x == 0 && wm0 == 0) {
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) {
      double v = bsum[j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      const int c = n0 + wn0 + 8 * j + g;
      if (t == 0 && c < J) colpart[(long long)blockIdx.y * ldcol + c] += v;
    }
  }
}

// ------------------------------------------------------------------------------------
// (a) feature map  Phi = s*[cos(X.W + phi) | sin(X.W + phi)]        SCFGP.py:98-102
//   X chunk (rows x D, row-major), Wt (J x D row-major, ld = ldw), phi (J)
//   Phi slab (rows x K, ld = ldp);  optional Psi = rowscale[n] * Phi
//   grid = rowblocks(128) * featblocks(64)
// ------------------------------------------------------------------------------------
template <int XVEC>
__global__ void __launch_bounds__(CfgFeat::NT, 2)
k_features(const double* __restrict__ X, long long ldx, int rows, int D,
           const double* __restrict__ Wt, long long ldw, const double* __restrict__ phi,
           int J, double s, double* __restrict__ Phi, long long ldp,
           const double* __restrict__ rowscale, double* __restrict__ Psi, int nfb) {
  using Cfg = CfgFeat;
  extern __shared__ __align__(16) double smem[];
  const int rb = blockIdx.x / nfb, fb = blockIdx.x % nfb;
  const int m0 = rb * Cfg::BM, n0 = fb * Cfg::BN;
  double acc[Cfg::MI][Cfg::NI][2];
  zero_acc<Cfg>(acc);
  IdentMap am{m0, rows}, bm{n0, J};
  gemm_mainloop<Cfg, true, true, XVEC, 2>(acc, X, ldx, am, Wt, ldw, bm, D, smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
  // A thread's two accumulators of a fragment are neighbouring features (2t, 2t+1): they leave
  // as one 16-byte store where the addresses allow it (ldp even; sin half needs J even).
  const bool vec_c = (ldp & 1) == 0 && (reinterpret_cast<uintptr_t>(Phi) & 15) == 0 &&
                     (Psi == nullptr || (reinterpret_cast<uintptr_t>(Psi) & 15) == 0);
  const bool vec_s = vec_c && (J & 1) == 0;
  // phases first: |z| decides (per warp) between the straight-line sincos and the library one
  double ph[Cfg::NI][2];
  bool big = false;
#pragma unroll
  for (int j = 0; j < Cfg::NI; ++j) {
    const int f = n0 + wn0 + 8 * j + 2 * t;   // even
    ph[j][0] = f < J ? phi[f] : 0.0;
    ph[j][1] = f + 1 < J ? phi[f + 1] : 0.0;
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i) {
      acc[i][j][0] += ph[j][0];
      acc[i][j][1] += ph[j][1];
      big |= !(fabs(acc[i][j][0]) <= kFastTrigMax) | !(fabs(acc[i][j][1]) <= kFastTrigMax);
    }
  }
  big = __any_sync(0xffffffffu, big);
  auto epilogue = [&](auto slow_tag) {
    constexpr bool SLOW = decltype(slow_tag)::value;
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) {
      const int f = n0 + wn0 + 8 * j + 2 * t;   // even
      const bool one = f < J, two = f + 1 < J;
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i) {
        const int r = m0 + wm0 + 8 * i + g;
        double sn0, cs0, sn1, cs1;
        sincos_sel<SLOW>(acc[i][j][0], sn0, cs0);
        sincos_sel<SLOW>(acc[i][j][1], sn1, cs1);
        if (r >= rows || !one) continue;
        cs0 *= s; sn0 *= s; cs1 *= s; sn1 *= s;
        double* row = Phi + (long long)r * ldp;
        if (two && vec_c) *reinterpret_cast<double2*>(row + f) = make_double2(cs0, cs1);
        else { row[f] = cs0; if (two) row[f + 1] = cs1; }
        if (two && vec_s) *reinterpret_cast<double2*>(row + J + f) = make_double2(sn0, sn1);
        else { row[J + f] = sn0; if (two) row[J + f + 1] = sn1; }
        if (Psi != nullptr) {
          const double w = rowscale[r];
          double* prow = Psi + (long long)r * ldp;
          if (two && vec_c) *reinterpret_cast<double2*>(prow + f) = make_double2(w * cs0, w * cs1);
          else { prow[f] = w * cs0; if (two) prow[f + 1] = w * cs1; }
          if (two && vec_s) *reinterpret_cast<double2*>(prow + J + f) = make_double2(w * sn0, w * sn1);
          else { prow[J + f] = w * sn0; if (two) prow[J + f + 1] = w * sn1; }
        }
      }
    }
  };
  if (big) epilogue(std::true_type{});
  else epilogue(std::false_type{});
}

// ------------------------------------------------------------------------------------
// weighted column sums of a slab:  part[rs][k] += sum_{n in split rs} w[n]*S[n,k]
//   (p = PhiT.y, q = PhiT.gmu, dphi = colsum(Zbar));  w == nullptr -> 1
//   grid = (ceil(cols/256), nsplit)
// ------------------------------------------------------------------------------------
__global__ void k_colsum(const double* __restrict__ S, long long lds, int rows, int cols,
                         const double* __restrict__ w, double* __restrict__ part, long long ldpart) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cols) return;
  const int nsplit = gridDim.y;
  const int per = (rows + nsplit - 1) / nsplit;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    double w0 = w ? w[r] : 1.0, w1 = w ? w[r + 1] : 1.0, w2 = w ? w[r + 2] : 1.0, w3 = w ? w[r + 3] : 1.0;
    a0 += w0 * S[(long long)r * lds + k];
    a1 += w1 * S[(long long)(r + 1) * lds + k];
    a2 += w2 * S[(long long)(r + 2) * lds + k];
    a3 += w3 * S[(long long)(r + 3) * lds + k];
  }
  for (; r < r1; ++r) a0 += (w ? w[r] : 1.0) * S[(long long)r * lds + k];
  part[(long long)blockIdx.y * ldpart + k] += (a0 + a1) + (a2 + a3);
}

// minibatch gather: Xout[i][:] = X[idx[i]][:], yout[i] = y[idx[i]]   (one warp per selected row)
__global__ void k_gather_rows(const double* __restrict__ X, const double* __restrict__ y, int D,
                              const long long* __restrict__ idx, long long n, double* __restrict__ Xout,
                              double* __restrict__ yout) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const long long r = idx[i];
  for (int d = lane; d < D; d += 32) Xout[i * D + d] = X[r * D + d];
  if (lane == 0) yout[i] = y[r];
}

// out[k] (+)= sum_rs part[rs][k]
__global__ void k_sum_parts(const double* __restrict__ part, long long ldpart, int nsplit, int cols,
                            double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cols) return;
  double a = 0;
  for (int s = 0; s < nsplit; ++s) a += part[(long long)s * ldpart + k];
  out[k] = a;
}

// sum of squares / dot products of short vectors into scalars (single block)
//   out[0] = sum a[i]*b[i]
__global__ void k_dot(const double* __restrict__ a, const double* __restrict__ b, long long n,
                      double* __restrict__ out, int accumulate) {
  __shared__ double sh[32];
  double v = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) v += a[i] * b[i];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + v : v;
  }
}

// yy partials over a long vector: part[blockIdx.x] += sum of squares of its slice
__global__ void k_sumsq_part(const double* __restrict__ y, long long n, double* __restrict__ part) {
  __shared__ double sh[32];
  double v = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    v += y[i] * y[i];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) part[blockIdx.x] += v;
  }
}

// out[o] = sum_i part[i*stride + o]   for o < nout   (single block, deterministic order)
__global__ void k_reduce_slots(const double* __restrict__ part, int nslots, int stride, int nout,
                               double* __restrict__ out) {
  __shared__ double sh[32];
  for (int o = 0; o < nout; ++o) {
    double v = 0;
    for (int i = threadIdx.x; i < nslots; i += blockDim.x) v += part[(long long)i * stride + o];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
      v = warp_sum(v);
      if (threadIdx.x == 0) out[o] = v;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// (a6) T = Phi.LiT, v_n = |T_n|^2 -- only the row sums of squares leave the tile.
//   items: (rb, kb); Li lower triangular => reduction depth min(K, (kb+1)*128)
//   vpart[kb][n] = sum over the tile's 128 columns                    SCFGP.py:112
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CfgBig::NT, 1)
k_tgemm_rowsq(const double* __restrict__ Phi, long long ldp, int rows, int K,
              const double* __restrict__ Li, long long ldl,
              double* __restrict__ vpart, long long ldv, const int4* __restrict__ items) {
  using Cfg = CfgBig;
  extern __shared__ __align__(16) double smem[];
  const int4 it = items[blockIdx.x];
  const int m0 = it.x * Cfg::BM, n0 = it.y * Cfg::BN;
  if (m0 >= rows) return;  // short last chunk: whole CTA leaves before any barrier
  const int depth = min(K, n0 + Cfg::BN);
  double acc[Cfg::MI][Cfg::NI][2];
  zero_acc<Cfg>(acc);
  IdentMap am{m0, rows}, bm{n0, K};
  gemm_mainloop<Cfg, true, true, 2, 2>(acc, Phi, ldp, am, Li, ldl, bm, depth, smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wmi = warp / Cfg::WARPS_N, wni = warp % Cfg::WARPS_N;
  double* red = smem;  // [WARPS_N][BM]
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
    double sq = 0;
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) sq += acc[i][j][0] * acc[i][j][0] + acc[i][j][1] * acc[i][j][1];
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    if (t == 0) red[wni * Cfg::BM + wmi * Cfg::WM + 8 * i + g] = sq;
  }
  __syncthreads();
  if (threadIdx.x < Cfg::BM) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < Cfg::WARPS_N; ++w) v += red[w * Cfg::BM + threadIdx.x];
    int r = m0 + threadIdx.x;
    if (r < rows) vpart[(long long)it.y * ldv + r] = v;
  }
}

// ------------------------------------------------------------------------------------
// per-row moments and adjoint seeds (A.2 pass 2, A.3)               SCFGP.py:111-124
//   mu = Phi.alpha, v = sum_kb vpart, r = mu - y, d = tau (v+1)
//   e = (r^2+v)/d + log(2 pi d)          (closed form of the 30-node Gauss-Hermite sum)
//   k = -(r^2+v)/d^2 + 1/d ; gmu = 2r/d ; gv = 1/d + tau k ; g_tau += k (v+1)
//   writes gmu, gv (global row arrays), Psi = gv*Phi, slot partials of E and g_tau
//   one warp per row; grid-stride; block b owns slot b of `slots` (2 doubles)
// ------------------------------------------------------------------------------------
__global__ void k_seeds(const double* __restrict__ Phi, long long ldp, int rows, int K,
                        const double* __restrict__ alpha, const double* __restrict__ vpart,
                        long long ldv, int nkb, const double* __restrict__ y, double tau,
                        double* __restrict__ gmu, double* __restrict__ gv,
                        double* __restrict__ Psi, double* __restrict__ slots, int want_grad,
                        const double* __restrict__ mu_in) {
  // mu_in != nullptr: mu_n = Phi_n.alpha was produced by the Phi.LiT kernel (emulated mode)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double Eacc = 0, Tacc = 0;
  for (int r = blockIdx.x * nw + warp; r < rows; r += gridDim.x * nw) {
    const double* row = Phi + (long long)r * ldp;
    double mu = 0;
    if (mu_in) mu = mu_in[r];
    else {
      for (int k = lane; k < K; k += 32) mu += row[k] * alpha[k];
      mu = warp_sum(mu);
    }
    double v = 0;
    for (int b = 0; b < nkb; ++b) v += vpart[(long long)b * ldv + r];
    const double res = mu - y[r];
    const double d = tau * (v + 1.0);
    const double num = res * res + v;
    const double kk = -num / (d * d) + 1.0 / d;
    const double gm = 2.0 * res / d;
    const double gvv = 1.0 / d + tau * kk;
    if (lane == 0) {
      Eacc += num / d + log(6.283185307179586476925286766559 * d);
      Tacc += kk * (v + 1.0);
      gmu[r] = gm;
      gv[r] = gvv;
    }
    if (want_grad) {
      double* prow = Psi + (long long)r * ldp;
      for (int k = lane; k < K; k += 32) prow[k] = gvv * row[k];
    }
  }
  __shared__ double sh[2][32];
  if (lane == 0) { sh[0][warp] = Eacc; sh[1][warp] = Tacc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0, tt = 0;
    for (int w = 0; w < nw; ++w) { e += sh[0][w]; tt += sh[1][w]; }
    slots[2 * blockIdx.x] += e;
    slots[2 * blockIdx.x + 1] += tt;
  }
}

// The same per-row moments and seeds when mu_n is already known (emulated mode: mu is a column of the
// Phi.[Li; alpha]T product) and Psi is not wanted: nothing but O(rows) work, one thread per row.
// Block b owns slot b of `slots` (E, g_tau).
__global__ void __launch_bounds__(256)
k_seeds_rows(int rows, const double* __restrict__ mu_in, const double* __restrict__ vpart, long long ldv, int nkb,
             const double* __restrict__ y, double tau, double* __restrict__ gmu, double* __restrict__ gv,
             double* __restrict__ slots) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0, tt = 0;
  if (r < rows) {
    double v = 0;
    for (int b = 0; b < nkb; ++b) v += vpart[(long long)b * ldv + r];
    const double res = mu_in[r] - y[r];
    const double d = tau * (v + 1.0);
    const double num = res * res + v;
    const double kk = -num / (d * d) + 1.0 / d;
    gmu[r] = 2.0 * res / d;
    gv[r] = 1.0 / d + tau * kk;
    e = num / d + log(6.283185307179586476925286766559 * d);
    tt = kk * (v + 1.0);
  }
  __shared__ double sh[2][8];
  e = warp_sum(e); tt = warp_sum(tt);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = e; sh[1][threadIdx.x >> 5] = tt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    e = tt = 0;
    for (int w = 0; w < 8; ++w) { e += sh[0][w]; tt += sh[1][w]; }
    slots[2 * blockIdx.x] += e;
    slots[2 * blockIdx.x + 1] += tt;
  }
}

// ------------------------------------------------------------------------------------
// (c) backward: Phibar = Phi.(2 Abar) + Psi.(2 B) + y pbarT + gmu alphaT, fused with the
// trig derivative:  Zbar = -Phibar_c * Phi_s + Phibar_s * Phi_c ;  sbar += <Phibar, Phi>
//   tile = 128 rows x 64 features; the 128 tile columns are ordered so that each warp's
//   32 columns are [16 cos | 16 sin] of the same 16 features (both halves of a feature
//   land in the same thread).
//   items: (rb, fb).  A2/B2 are symmetric, so their rows serve as K-major B operands.
// ------------------------------------------------------------------------------------
struct TrigPairMap {
  int fbase;  // first feature of this tile
  int J;
  __device__ __forceinline__ int operator()(int c) const {
    const int wc = c >> 5, sub = c & 31;
    const int f = fbase + wc * 16 + (sub & 15);
    if (f >= J) return -1;
    return (sub < 16) ? f : J + f;
  }
};

__global__ void __launch_bounds__(CfgBig::NT, 1)
k_backward(const double* __restrict__ Phi, const double* __restrict__ Psi, long long ldp,
           int rows, int K, int J, const double* __restrict__ A2, const double* __restrict__ B2,
           long long ldm, const double* __restrict__ y, const double* __restrict__ gmu,
           const double* __restrict__ pbar, const double* __restrict__ alpha,
           double* __restrict__ Zbar, long long ldz, double* __restrict__ slots,
           const int4* __restrict__ items) {
  using Cfg = CfgBig;
  extern __shared__ __align__(16) double smem[];
  const int4 it = items[blockIdx.x];
  const int m0 = it.x * Cfg::BM;
  const int f0 = it.y * 64;
  double acc[Cfg::MI][Cfg::NI][2];
  zero_acc<Cfg>(acc);
  IdentMap am{m0, rows};
  TrigPairMap bm{f0, J};
  gemm_mainloop<Cfg, true, true, 2, 2>(acc, Phi, ldp, am, A2, ldm, bm, K, smem);
  gemm_mainloop<Cfg, true, true, 2, 2>(acc, Psi, ldp, am, B2, ldm, bm, K, smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wmi = warp / Cfg::WARPS_N, wni = warp % Cfg::WARPS_N;
  double sacc = 0;
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
    const int r = m0 + wmi * Cfg::WM + 8 * i + g;
    if (r >= rows) continue;
    const double yn = y[r], gm = gmu[r];
    const double* prow = Phi + (long long)r * ldp;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int f = f0 + wni * 16 + 8 * jj + 2 * t + e;
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
  __shared__ double sh[Cfg::NWARPS];
  if (lane == 0) sh[warp] = sacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0;
    for (int w = 0; w < Cfg::NWARPS; ++w) v += sh[w];
    slots[blockIdx.x] += v;
  }
}

// ------------------------------------------------------------------------------------
// predictive finish:  mu = Phi.alpha ; std = sqrt(tau (1 + v))          SCFGP.py:144-145
// ------------------------------------------------------------------------------------
__global__ void k_pred_finish(const double* __restrict__ Phi, long long ldp, int rows, int K,
                              const double* __restrict__ alpha, const double* __restrict__ vpart,
                              long long ldv, int nkb, double tau, double* __restrict__ mu,
                              double* __restrict__ sd, const double* __restrict__ mu_in) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = blockIdx.x * nw + warp; r < rows; r += gridDim.x * nw) {
    const double* row = Phi + (long long)r * ldp;
    double m = 0;
    if (mu_in) m = mu_in[r];
    else {
      for (int k = lane; k < K; k += 32) m += row[k] * alpha[k];
      m = warp_sum(m);
    }
    if (lane == 0) {
      double v = 0;
      for (int b = 0; b < nkb; ++b) v += vpart[(long long)b * ldv + r];
      mu[r] = m;
      sd[r] = sqrt(tau * (1.0 + v));
    }
  }
}

// ------------------------------------------------------------------------------------
// (d) small replicated dense kernels (K x K, fp64)
// ------------------------------------------------------------------------------------
constexpr int NB = 64;  // Cholesky block

__global__ void k_add_diag(double* __restrict__ A, long long ld, int K, double val) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) A[(long long)i * ld + i] += val;
}

// Factor the nb x nb diagonal block at (k0,k0) in place (lower), zero its strict upper
// part, and write inv(L11) (lower) to Dinv (NB x NB, zero padded).        SCFGP.py:106-107
//   Factorisation: right-looking; the 256 threads hold the 64 x 64 block as 4 x 4 register tiles and
//   only column j travels through shared memory (double-buffered: ONE barrier per column).  The
//   trailing update uses the unscaled column:  a[i][c] -= a[i][j] a[c][j] / a[j][j].
//   Inverse: forward substitution, 4 lanes per column split every dot product (quad shuffles).
constexpr size_t POTRF_SMEM = 2 * NB * (NB + 1) * sizeof(double);
__global__ void __launch_bounds__(256, 1)
k_potrf_diag(double* __restrict__ A, long long ld, int k0, int nb, double* __restrict__ Dinv,
             long long* __restrict__ info) {
  static_assert(NB == 64, "4 x 4 tiles of 16 x 16 threads");
  extern __shared__ __align__(16) double smem_pd[];
  double (*L)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem_pd);                  // the factor
  double (*X)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem_pd + NB * (NB + 1));  // its inverse
  __shared__ double colbuf[2][NB];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  double r[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int i = 4 * ty + ii, c = 4 * tx + cc;
      // rows / columns beyond nb: identity padding keeps the arithmetic finite
      r[ii][cc] = (i < nb && c <= i) ? A[(long long)(k0 + i) * ld + k0 + c] : ((i == c && i >= nb) ? 1.0 : 0.0);
    }
  for (int e = tid; e < NB * (NB + 1); e += blockDim.x) { (&L[0][0])[e] = 0.0; (&X[0][0])[e] = 0.0; }
  __syncthreads();
  for (int j = 0; j < NB; ++j) {
    const int jj = j & 3;
    if (tx == (j >> 2)) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
        colbuf[j & 1][4 * ty + ii] = jj == 0 ? r[ii][0] : (jj == 1 ? r[ii][1] : (jj == 2 ? r[ii][2] : r[ii][3]));
    }
    __syncthreads();
    const double* col = colbuf[j & 1];
    const double d = col[j];
    if (tid == 0 && j < nb && !(d > 0.0) && *info == 0) *info = (long long)(k0 + j) + 1;
    const double sq = sqrt(d), rs = 1.0 / sq, rd = 1.0 / d;
    if (tx == (j >> 2)) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const int i = 4 * ty + ii;
        if (i >= j) L[i][j] = (i == j) ? sq : col[i] * rs;
      }
    }
    double ci[4], cc4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { ci[k] = col[4 * ty + k]; cc4[k] = col[4 * tx + k] * rd; }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int i = 4 * ty + ii, c = 4 * tx + cc;
        if (c > j && c <= i) r[ii][cc] -= ci[ii] * cc4[cc];
      }
  }
  __syncthreads();
  // inverse: column c by the quad of lanes 4c .. 4c+3
  {
    const int c = tid >> 2, part = tid & 3;
    for (int i = 0; i < NB; ++i) {
      double sacc = 0.0;
      if (i > c)
        for (int m = c + part; m < i; m += 4) sacc += L[i][m] * X[m][c];
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
      if (part == 0 && i >= c) X[i][c] = ((i == c ? 1.0 : 0.0) - sacc) / L[i][i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    int i = e / NB, j = e % NB;
    if (i < nb && j < nb) A[(long long)(k0 + i) * ld + k0 + j] = (j <= i) ? L[i][j] : 0.0;
    Dinv[e] = (i < nb && j < nb) ? X[i][j] : 0.0;
  }
}

// zero the strict upper triangle / mirror lower to upper
__global__ void k_zero_upper(double* __restrict__ A, long long ld, int K) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)K * K) return;
  int i = (int)(e / K), j = (int)(e % K);
  if (j > i) A[(long long)i * ld + j] = 0.0;
}
__global__ void k_symmetrize(double* __restrict__ A, long long ld, int K) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  int i = bi * 32 + threadIdx.y, j = bj * 32 + threadIdx.x;
  tile[threadIdx.y][threadIdx.x] = (i < K && j < K) ? A[(long long)i * ld + j] : 0.0;
  __syncthreads();
  // write transposed block into (bj, bi)
  int ti = bj * 32 + threadIdx.y, tj = bi * 32 + threadIdx.x;
  if (ti < K && tj < K && tj > ti) A[(long long)ti * ld + tj] = tile[threadIdx.x][threadIdx.y];
}

// all diagonal blocks at once: Li[b NB + i][b NB + j] = Dinv[b][i][j]
__global__ void k_place_diag_blocks(const double* __restrict__ Dinv, double* __restrict__ Li, long long ld, int K) {
  const int b = blockIdx.x;
  const int nb = min(NB, K - b * NB);
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    Li[(long long)(b * NB + i) * ld + b * NB + j] = Dinv[(long long)b * NB * NB + i * NB + j];
  }
}

// copy a block (used to place inv(L_ii) on Li's diagonal)
__global__ void k_copy_block(const double* __restrict__ src, long long lds, double* __restrict__ dst,
                             long long ldd, int rows, int cols) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  int i = e / cols, j = e % cols;
  dst[(long long)i * ldd + j] = src[(long long)i * lds + j];
}

// y = M x  (row-major M, one warp per row);  trans handled by k_colsum
__global__ void k_gemv_n(const double* __restrict__ Mx, long long ld, int rows, int cols,
                         const double* __restrict__ x, double* __restrict__ yout) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = blockIdx.x * nw + warp; r < rows; r += gridDim.x * nw) {
    const double* row = Mx + (long long)r * ld;
    double v = 0;
    for (int k = lane; k < cols; k += 32) v += row[k] * x[k];
    v = warp_sum(v);
    if (lane == 0) yout[r] = v;
  }
}

// scalars of the factor phase: out[0] = sum log L_ii, out[1] = beta.beta, out[2] = p.alpha
__global__ void k_factor_scalars(const double* __restrict__ L, long long ld, int K,
                                 const double* __restrict__ beta, const double* __restrict__ p,
                                 const double* __restrict__ alpha, double* __restrict__ out) {
  __shared__ double sh[3][32];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    a += log(L[(long long)i * ld + i]);
    b += beta[i] * beta[i];
    c += p[i] * alpha[i];
  }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; sh[2][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = c = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
    out[0] = a; out[1] = b; out[2] = c;
  }
}

// A2 = 2*(B - BHB - 0.5 (t aT + a tT) + a aT / sig2n),  B2 = 2 B,  pbar = t - 2 a / sig2n,
// trace slots: tr[blockIdx.x] = partial trace of Abar.                         (A.4)
__global__ void k_abar(const double* __restrict__ Bm, const double* __restrict__ BHB, long long ld, int K,
                       const double* __restrict__ tvec, const double* __restrict__ alpha, double sig2n,
                       double* __restrict__ A2, double* __restrict__ B2, double* __restrict__ pbar,
                       double* __restrict__ trslots) {
  const int i = blockIdx.x;
  if (i >= K) return;
  const double ti = tvec[i], ai = alpha[i];
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    const long long o = (long long)i * ld + j;
    const double bij = Bm[o];
    const double ab = bij - BHB[o] - 0.5 * (ti * alpha[j] + ai * tvec[j]) + ai * alpha[j] / sig2n;
    A2[o] = 2.0 * ab;
    B2[o] = 2.0 * bij;
    if (j == i) trslots[i] = ab;
  }
  if (threadIdx.x == 0) pbar[i] = ti - 2.0 * ai / sig2n;
}

// sbar = sum <Phibar, Phi> (A.5) without touching the rows again:
//   Phibar = Phi (2 Abar) + (2 gv) . (Phi B) + y pbarT + gmu alphaT
//   => sum_n <Phibar_n, Phi_n> = tr(2 Abar G) + tr(2 B H) + pbarT p + alphaT q
// with G = PhiT Phi, H = PhiT diag(gv) Phi, p = PhiT y, q = PhiT gmu of passes 1 and 2 (all-reduced already).
// G and H hold (at least) their lower triangles; A2 = 2 Abar and B2 = 2 B are symmetric.
//   part[i] = row i of the two traces (+ the two dot-product terms), summed by k_reduce_slots
__global__ void k_sbar_rows(const double* __restrict__ A2, const double* __restrict__ B2,
                            const double* __restrict__ G, const double* __restrict__ H, long long ld, int K,
                            const double* __restrict__ pbar, const double* __restrict__ p,
                            const double* __restrict__ alpha, const double* __restrict__ q,
                            double* __restrict__ part) {
  const int i = blockIdx.x;
  if (i >= K) return;
  double v = 0;
  for (int j = threadIdx.x; j <= i; j += blockDim.x) {
    const long long o = (long long)i * ld + j;
    const double w = j == i ? 1.0 : 2.0;
    v = fma(w * A2[o], G[o], v);
    v = fma(w * B2[o], H[o], v);
  }
  __shared__ double sh[32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
    part[i] = t + pbar[i] * p[i] + alpha[i] * q[i];
  }
}

}  // namespace scfgp
