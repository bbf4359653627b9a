// Emulated-FP64 versions of the three hot products on the int8 tensor cores (tcgen05.mma
// kind::i8, see umma_i8.cuh for the number format and the pipeline), plus the digit
// extraction kernels that feed them.  Same math as kernels_tma.cuh:
//   k_gram_i8      G += PhiT.Phi, H += PsiT.Phi          SCFGP.py:103-105, A.3  (MN-major operands: straight
//                                                        from the row-major digits, contraction = rows)
//   k_tgemm_i8     v_n = |Phi_n LiT|^2                   SCFGP.py:112
//   k_backward_i8  Phibar = [Phi | Psi'].[2Abar | 2 sigma B]T, fused trig derivative   A.5
// All scales are powers of two, so scaling is exact; the only approximation is the
// truncation of each operand to 8*KS-2 fractional bits of its row (or global) scale.
#pragma once
#include "umma_i8.cuh"
#include "kernels.cuh"

#ifndef UMMA_B_KEEP
#define UMMA_B_KEEP true   // L2 evict_last hint on the replicated B operand of Phi.LiT and the backward product
#endif

namespace scfgp {
namespace emu {

using umma::TM;
using umma::TN;
using umma::KB;

// digit scale of a quantity bounded by x: the smallest power of two >= 2 x (2 for x == 0), so that
// |value| / scale <= 1/2 and the leading digit stays within +-64
__device__ __forceinline__ double pow2_ceil(double x) {
  if (!(x > 0.0)) return 2.0;
  int e;
  const double m = frexp(x, &e);  // x = m 2^e, m in [0.5, 1)
  return ldexp(1.0, (m == 0.5 ? e - 1 : e) + 1);
}

// row permutation of the backward B operand: cos and sin halves of a feature become
// neighbours so that both land in the same epilogue thread
__device__ __forceinline__ int trig_perm(int j, int J) { return j < J ? 2 * j : 2 * (j - J) + 1; }

// Balanced base-256 digits of q (|q| <= 2^(8 KS - 2)); digit 0 most significant (|.| <= 64),
// the others in [-128, 127].  Adding 128 at every digit position below the top one (the
// constant digit_bias) turns the balanced digits into plain 8-bit fields u_i of q' = q + bias
// with digit_i = u_i - 128, i.e. the int8 byte is u_i ^ 0x80; the top digit is q' >> 8 (KS - 1).
// digit_byte<KS>(q', i) returns u_i (i >= 1) or the top digit's two's-complement byte (i == 0);
// fix_bytes turns four packed u fields into four int8 digits.
constexpr int DB = umma::DIGIT_BITS;
template <int KS>
__device__ __forceinline__ long long digit_bias() {
  long long b = 0;
#pragma unroll
  for (int t = 0; t < KS - 1; ++t) b += 128ll << (DB * t);
  return b;
}
// multiplier that maps |x| <= 1/2 (x already divided by its power-of-two scale) to |q| <= 2^(8 KS - 2)
template <int KS>
__device__ __forceinline__ double digit_mult() { return ldexp(1.0, DB * KS - 1); }
template <int KS>
__device__ __forceinline__ long long biased_q(double x) {
  const long long lim = 1ll << (DB * KS - 2);
  long long q = llrint(x);
  q = q > lim ? lim : (q < -lim ? -lim : q);
  return q + digit_bias<KS>();
}
template <int KS>
__device__ __forceinline__ uint32_t digit_byte(long long qp, int i) {
  const int shift = DB * (KS - 1 - i);
  return (uint32_t)((unsigned long long)(qp >> shift)) & 0xffu;
}
__device__ __forceinline__ uint32_t fix_bytes(uint32_t w) { return w ^ 0x80808080u; }  // per byte: u -> u - 128

// ---- 5-digit fast paths (the default format): q' = q + 0x80808080 is a 40-bit field, top byte : four u fields ----
// quant5: q' of v (= x * 2^39 / scale) as (hi byte, 32-bit word).  Adding 1.5 * 2^52 + bias rounds v to the nearest
// integer (ties to even, as llrint does) and leaves q' in the low mantissa bits (two's complement).
__device__ __forceinline__ void quant5(double v, uint32_t& hi, uint32_t& lo) {
  const double lim = 274877906944.0;  // 2^38
  v = fmin(fmax(v, -lim), lim);
  const double d = v + (6755399441055744.0 + 2155905152.0);
  lo = (uint32_t)__double2loint(d);
  hi = (uint32_t)__double2hiint(d);  // only its low byte is used
}
// the same without the clamp, for values that are inside +-2^38 by construction (the feature map: |Phi| <= s <= s2 / 2)
__device__ __forceinline__ void quant5_inrange(double v, uint32_t& hi, uint32_t& lo) {
  const double d = v + (6755399441055744.0 + 2155905152.0);
  lo = (uint32_t)__double2loint(d);
  hi = (uint32_t)__double2hiint(d);
}
// four q' -> the five plane words (byte e of a word = element e): a 4 x 4 byte transpose with byte permutes
__device__ __forceinline__ void pack5x4(const uint32_t (&hi)[4], const uint32_t (&lo)[4], uint32_t (&w)[5]) {
  const uint32_t a = __byte_perm(lo[0], lo[1], 0x5140), b = __byte_perm(lo[0], lo[1], 0x7362);
  const uint32_t c = __byte_perm(lo[2], lo[3], 0x5140), d = __byte_perm(lo[2], lo[3], 0x7362);
  w[4] = fix_bytes(__byte_perm(a, c, 0x5410));
  w[3] = fix_bytes(__byte_perm(a, c, 0x7632));
  w[2] = fix_bytes(__byte_perm(b, d, 0x5410));
  w[1] = fix_bytes(__byte_perm(b, d, 0x7632));
  w[0] = __byte_perm(__byte_perm(hi[0], hi[1], 0x0040), __byte_perm(hi[2], hi[3], 0x0040), 0x5410);
}
// the five plane words of four elements -> the four integers q as doubles (exact)
__device__ __forceinline__ void unpack5x4(const uint32_t (&w)[5], double (&q)[4]) {
  const uint32_t u1 = fix_bytes(w[1]), u2 = fix_bytes(w[2]), u3 = fix_bytes(w[3]), u4 = fix_bytes(w[4]);
  const uint32_t a = __byte_perm(u4, u3, 0x5140), b = __byte_perm(u4, u3, 0x7362);
  const uint32_t c = __byte_perm(u2, u1, 0x5140), d = __byte_perm(u2, u1, 0x7362);
  const uint32_t lo[4] = {__byte_perm(a, c, 0x5410), __byte_perm(a, c, 0x7632), __byte_perm(b, d, 0x5410),
                          __byte_perm(b, d, 0x7632)};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int hi = (int)(int8_t)(w[0] >> (8 * e));
    q[e] = fma((double)hi, 4294967296.0, (double)lo[e] - 2155905152.0);
  }
}

// the same as exact integers (byte permutes and one 64-bit add per element; the caller converts once)
__device__ __forceinline__ void unpack5x4_q(const uint32_t (&w)[5], long long (&q)[4]) {
  const uint32_t u1 = fix_bytes(w[1]), u2 = fix_bytes(w[2]), u3 = fix_bytes(w[3]), u4 = fix_bytes(w[4]);
  const uint32_t a = __byte_perm(u4, u3, 0x5140), b = __byte_perm(u4, u3, 0x7362);
  const uint32_t c = __byte_perm(u2, u1, 0x5140), d = __byte_perm(u2, u1, 0x7362);
  const uint32_t lo[4] = {__byte_perm(a, c, 0x5410), __byte_perm(a, c, 0x7632), __byte_perm(b, d, 0x5410),
                          __byte_perm(b, d, 0x7632)};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const long long hi = (long long)(int8_t)(w[0] >> (8 * e));
    q[e] = (hi << 32) + (long long)lo[e] - 2155905152ll;   // q' = q + 0x80808080
  }
}

// ------------------------------------------------------------------------------------
// per-row power-of-two scales of [M1 | f2 * M2] (either may be absent)
//   scale[p(r)] = pow2_ceil(max_c |row r|), inv[p(r)] = 1 / scale
// ------------------------------------------------------------------------------------
__global__ void k_rowscale(const double* __restrict__ M1, long long ld1, const double* __restrict__ M2, long long ld2,
                           const double* __restrict__ f2_ptr, int rows, int cols, int perm_J,
                           double* __restrict__ scale, double* __restrict__ inv) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const double f2 = f2_ptr ? fabs(*f2_ptr) : 1.0;
  double m = 0;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    m = fmax(m, fabs(M1[(long long)r * ld1 + c]));
    if (M2) m = fmax(m, f2 * fabs(M2[(long long)r * ld2 + c]));
  }
  __shared__ double sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmax(m, sh[w]);
    const double s = pow2_ceil(m);
    const int o = perm_J > 0 ? trig_perm(r, perm_J) : r;
    scale[o] = s;
    inv[o] = 1.0 / s;
  }
}

// out[i] = in[i] * f
__global__ void k_scale_vector(const double* __restrict__ in, int n, double f, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * f;
}

// power-of-two bound of |v[0..n)|: out[0] = scale, out[1] = 1/scale; running maximum of the
// scale in gmax (bit pattern of a positive double orders like an integer)
__global__ void k_absmax_pow2(const double* __restrict__ v, int n, double* __restrict__ out,
                              unsigned long long* __restrict__ gmax) {
  double m = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(v[i]));
  __shared__ double sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmax(m, sh[w]);
    const double s = pow2_ceil(m);
    out[0] = s;
    out[1] = 1.0 / s;
    if (gmax) atomicMax(gmax, (unsigned long long)__double_as_longlong(s));
  }
}
// out[0] = *gmax (as double), out[1] = 1 / out[0]
__global__ void k_scale_from_bits(const unsigned long long* __restrict__ gmax, double* __restrict__ out) {
  const double s = __longlong_as_double((long long)*gmax);
  out[0] = s > 0 ? s : 1.0;
  out[1] = 1.0 / out[0];
}

// ------------------------------------------------------------------------------------
// digits of a row-major matrix, contraction index = column (K-major operand)
//   x = src[r][c] * rowmul[r] * rowinv[p(r)] * ginv * *ginv_ptr  (absent factors = 1; |x| <= 1)
//   dst[s][p(r)][col_off + c] = digit s of x;   one thread = 4 consecutive columns (32 bytes in,
//   4 bytes per digit out: both fully coalesced); columns [cols, round_up(cols, 16)) are zeroed
// ------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256)
k_digits_rows(const double* __restrict__ src, long long ld, int rows, int cols, const double* __restrict__ rowmul,
              const double* __restrict__ rowinv, double ginv, const double* __restrict__ ginv_ptr, int perm_J,
              int8_t* __restrict__ dst, long long pitch, long long slice_stride, int col_off,
              const double* __restrict__ rowmul2, const double* __restrict__ ginv_ptr2, int col_off2) {
  // optional second output from the same read of src (pass 3: [Phi | (gv / sigma) Phi]):
  //   dst[s][r][col_off2 + c] = digit s of src[r][c] * rowmul2[r] * ginv * *ginv_ptr2
  const int ngroups = (cols + 15) / 16 * 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * ngroups) return;
  const int r = (int)(idx / ngroups), g = (int)(idx % ngroups);
  const int orow = perm_J > 0 ? trig_perm(r, perm_J) : r;
  double mult = digit_mult<KS>() * ginv;
  if (ginv_ptr) mult *= *ginv_ptr;
  if (rowmul) mult *= rowmul[r];
  if (rowinv) mult *= rowinv[orow];
  const double* p = src + (long long)r * ld + g * 4;
  double x[4];
  if (g * 4 + 3 < cols && ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = (g * 4 + e < cols) ? p[e] : 0.0;
  }
  uint32_t packed[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) packed[s] = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const long long qp = biased_q<KS>(x[e] * mult);
#pragma unroll
    for (int s = 0; s < KS; ++s) packed[s] |= digit_byte<KS>(qp, s) << (8 * e);
  }
  int8_t* o = dst + (long long)orow * pitch + col_off + g * 4;
  *reinterpret_cast<uint32_t*>(o) = packed[0];
#pragma unroll
  for (int s = 1; s < KS; ++s) *reinterpret_cast<uint32_t*>(o + (long long)s * slice_stride) = fix_bytes(packed[s]);
  if (rowmul2) {
    const double mult2 = digit_mult<KS>() * ginv * (ginv_ptr2 ? *ginv_ptr2 : 1.0) * rowmul2[r];
#pragma unroll
    for (int s = 0; s < KS; ++s) packed[s] = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long qp = biased_q<KS>(x[e] * mult2);
#pragma unroll
      for (int s = 0; s < KS; ++s) packed[s] |= digit_byte<KS>(qp, s) << (8 * e);
    }
    int8_t* o2 = dst + (long long)orow * pitch + col_off2 + g * 4;
    *reinterpret_cast<uint32_t*>(o2) = packed[0];
#pragma unroll
    for (int s = 1; s < KS; ++s) *reinterpret_cast<uint32_t*>(o2 + (long long)s * slice_stride) = fix_bytes(packed[s]);
  }
}

// digits of four neighbouring elements x[e] * mm -> one 32-bit word per digit plane
template <int KS>
__device__ __forceinline__ void store_digits4(const double (&x)[4], double mm, int8_t* o, long long plane) {
  if constexpr (KS == 5) {
    uint32_t hi[4], lo[4], w[5];
#pragma unroll
    for (int e = 0; e < 4; ++e) quant5(x[e] * mm, hi[e], lo[e]);
    pack5x4(hi, lo, w);
#pragma unroll
    for (int s = 0; s < 5; ++s) *reinterpret_cast<uint32_t*>(o + (long long)s * plane) = w[s];
  } else {
    uint32_t packed[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) packed[s] = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long qp = biased_q<KS>(x[e] * mm);
#pragma unroll
      for (int s = 0; s < KS; ++s) packed[s] |= digit_byte<KS>(qp, s) << (8 * e);
    }
    *reinterpret_cast<uint32_t*>(o) = packed[0];
#pragma unroll
    for (int s = 1; s < KS; ++s) *reinterpret_cast<uint32_t*>(o + (long long)s * plane) = fix_bytes(packed[s]);
  }
}

// ------------------------------------------------------------------------------------
// digits of a streaming row slab (the Phi chunk), all its by-products in ONE pass over the data:
//   source  : the float64 slab (FROM_DIGITS = false) or the digit cache (true: x = recon * sum_i d_i 256^(KS-1-i))
//   out1    : digit s of x * ginv * rowmul1[r] * *gptr1  -> dst1[s][r][4 tx ..]   (absent factors = 1)
//   out2    : the same with rowmul2 / gptr2 -> dst2                                   (optional)
//   cpart   : cpart[blockIdx.y][c] += sum over the block's 128 rows of wsum[r] * x[r][c]   (optional; p, q)
// Rows [rows, next multiple of 128) are written as zeros: they lie inside the Gram reduction.
// grid = (ceil(ceil16(cols) / 256), ceil(rows / 128)); 256 threads = 64 column groups x 4 row lanes.
// ------------------------------------------------------------------------------------
template <int KS, bool FROM_DIGITS>
__global__ void __launch_bounds__(256)
k_digits_slab(const double* __restrict__ src, long long ld, const int8_t* __restrict__ dsrc, long long dpitch,
              long long dplane, double recon, int rows, int cols, double ginv,
              const double* __restrict__ rowmul1, const double* __restrict__ gptr1, int8_t* __restrict__ dst1,
              const double* __restrict__ rowmul2, const double* __restrict__ gptr2, int8_t* __restrict__ dst2,
              long long pitch, long long plane, const double* __restrict__ wsum, double* __restrict__ cpart,
              long long ldpart, double* __restrict__ xout, long long ldx) {
  // xout (optional): the source values themselves as a float64 slab (from the cache: the reconstructed
  // Phi the backward epilogue reads -- a coalesced write here instead of strided byte loads there)
  __shared__ double csum[4][256];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = (blockIdx.x * 64 + tx) * 4;
  const int cpad = (cols + 15) / 16 * 16;
  const int r0 = blockIdx.y * 128;
  const double m1 = digit_mult<KS>() * ginv * (gptr1 ? *gptr1 : 1.0);
  const double m2 = digit_mult<KS>() * ginv * (gptr2 ? *gptr2 : 1.0);
  const bool vec = !FROM_DIGITS && (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < cpad) {
#pragma unroll 4
    for (int it = 0; it < 32; ++it) {
      const int r = r0 + ty + 4 * it;
      double x[4] = {0.0, 0.0, 0.0, 0.0};
      if (r < rows) {
        if (FROM_DIGITS) {
          if constexpr (KS == 5) {
            uint32_t w[5];
#pragma unroll
            for (int s = 0; s < 5; ++s)
              w[s] = *reinterpret_cast<const uint32_t*>(dsrc + (long long)s * dplane + (long long)r * dpitch + c);
            double q[4];
            unpack5x4(w, q);
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (c + e < cols) ? q[e] * recon : 0.0;
          } else {
            long long q[4] = {0, 0, 0, 0};
#pragma unroll
            for (int s = 0; s < KS; ++s) {
              const uint32_t w =
                  *reinterpret_cast<const uint32_t*>(dsrc + (long long)s * dplane + (long long)r * dpitch + c);
#pragma unroll
              for (int e = 0; e < 4; ++e) q[e] = q[e] * 256 + (long long)(int8_t)(w >> (8 * e));
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (c + e < cols) ? (double)q[e] * recon : 0.0;
          }
        } else {
          const double* p = src + (long long)r * ld + c;
          if (vec && c + 3 < cols) {
            const double2 a = *reinterpret_cast<const double2*>(p);
            const double2 b = *reinterpret_cast<const double2*>(p + 2);
            x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (c + e < cols) ? p[e] : 0.0;
          }
        }
        if (wsum) {
          const double w = wsum[r];
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] = fma(w, x[e], acc[e]);
        }
        if (xout) {
          double* o = xout + (long long)r * ldx + c;
          if ((ldx & 1) == 0 && c + 3 < cols) {
            *reinterpret_cast<double2*>(o) = make_double2(x[0], x[1]);
            *reinterpret_cast<double2*>(o + 2) = make_double2(x[2], x[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c + e < cols) o[e] = x[e];
          }
        }
      }
      if (dst1) {
        const double mm = (r < rows && rowmul1) ? m1 * rowmul1[r] : m1;
        store_digits4<KS>(x, mm, dst1 + (long long)r * pitch + c, plane);
      }
      if (dst2) {
        const double mm = (r < rows && rowmul2) ? m2 * rowmul2[r] : m2;
        store_digits4<KS>(x, mm, dst2 + (long long)r * pitch + c, plane);
      }
    }
  }
  if (wsum) {
#pragma unroll
    for (int e = 0; e < 4; ++e) csum[ty][tx * 4 + e] = acc[e];
    __syncthreads();
    const int cc = blockIdx.x * 256 + threadIdx.x;
    if (cc < cols)
      cpart[(long long)blockIdx.y * ldpart + cc] +=
          (csum[0][threadIdx.x] + csum[1][threadIdx.x]) + (csum[2][threadIdx.x] + csum[3][threadIdx.x]);
  }
}

#ifndef DIGITS_CACHE5_MINB
#define DIGITS_CACHE5_MINB 3
#endif
// ------------------------------------------------------------------------------------
// The same for the default case (5 digits, source = the digit cache, one output): one thread = one row x 16
// neighbouring columns -- a 16-byte load per source plane, a 16-byte store per output plane (80 bytes in flight per
// row instead of 20), byte-permute unpacking / packing, the range clamp applied once per row to the row factor
// (|x m1| <= 2^38 holds for every value that comes out of valid digits) instead of once per value.
//   out  : digit s of x * ginv * clamp(rowmul[r] * *gptr, +-1)  -> dst[s][r][c ..]      (x = recon * cached q)
//   cpart: cpart[blockIdx.y][c] += sum over the block's 128 rows of wsum[r] * x[r][c]   (optional; q = PhiT.gmu)
// grid = (ceil(ceil16(cols) / 128), ceil(rows / 128)); 256 threads = 8 column groups x 32 row lanes.
// ------------------------------------------------------------------------------------
template <bool WSUM>
__global__ void __launch_bounds__(256, DIGITS_CACHE5_MINB)
k_digits_cache5(const int8_t* __restrict__ dsrc, long long dpitch, long long dplane, double recon, int rows, int cols,
                double ginv, const double* __restrict__ rowmul, const double* __restrict__ gptr,
                int8_t* __restrict__ dst, long long pitch, long long plane, const double* __restrict__ wsum,
                double* __restrict__ cpart, long long ldpart) {
  __shared__ double csum[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int c = (blockIdx.x * 8 + cg) * 16;
  const int cpad = (cols + 15) / 16 * 16;
  const int r0 = blockIdx.y * 128;
  const double m1 = digit_mult<5>() * ginv;
  const double gs = gptr ? *gptr : 1.0;
  double acc[WSUM ? 16 : 1];      // (WSUM = false: no column sums, 32 registers fewer)
#pragma unroll
  for (int e = 0; e < (WSUM ? 16 : 1); ++e) acc[e] = 0.0;
  if (c < cpad) {
#pragma unroll 2
    for (int it = 0; it < 4; ++it) {
      const int r = r0 + ty + 32 * it;
      const bool rv = r < rows;
      uint4 w[5];
#pragma unroll
      for (int d = 0; d < 5; ++d) w[d] = make_uint4(0u, 0u, 0u, 0u);   // zero digits
      if (rv) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
          w[d] = *reinterpret_cast<const uint4*>(dsrc + (long long)d * dplane + (long long)r * dpitch + c);
      }
      double mm = m1;
      double wr = 0.0;
      if (rv) {
        if (rowmul) mm = m1 * fmin(fmax(rowmul[r] * gs, -1.0), 1.0);
        else mm = m1 * fmin(fmax(gs, -1.0), 1.0);
        if (WSUM) wr = wsum[r];
      }
      uint4 o[5];
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        uint32_t wi[5], hi[4], lo[4], wo[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) wi[d] = g4 == 0 ? w[d].x : (g4 == 1 ? w[d].y : (g4 == 2 ? w[d].z : w[d].w));
        double q[4];
        unpack5x4(wi, q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double x = (c + 4 * g4 + e < cols) ? q[e] * recon : 0.0;
          if constexpr (WSUM) acc[4 * g4 + e] = fma(wr, x, acc[4 * g4 + e]);
          quant5_inrange(x * mm, hi[e], lo[e]);
        }
        pack5x4(hi, lo, wo);
#pragma unroll
        for (int d = 0; d < 5; ++d) {
          if (g4 == 0) o[d].x = wo[d];
          else if (g4 == 1) o[d].y = wo[d];
          else if (g4 == 2) o[d].z = wo[d];
          else o[d].w = wo[d];
        }
      }
#pragma unroll
      for (int d = 0; d < 5; ++d)
        *reinterpret_cast<uint4*>(dst + (long long)d * plane + (long long)r * pitch + c) = o[d];
    }
  }
  if constexpr (WSUM) {
    // the warp's 4 row lanes (lane bits 3, 4) as a reduce-scatter, then the 8 warps through shared memory (fixed order)
    const bool b4 = lane & 16, b3 = lane & 8;
    double a8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      a8[k] = (b4 ? acc[8 + k] : acc[k]) + __shfl_xor_sync(0xffffffffu, b4 ? acc[k] : acc[8 + k], 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double v = (b3 ? a8[4 + k] : a8[k]) + __shfl_xor_sync(0xffffffffu, b3 ? a8[k] : a8[4 + k], 8);
      csum[warp][cg * 16 + (b4 ? 8 : 0) + (b3 ? 4 : 0) + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int cc = blockIdx.x * 128 + threadIdx.x;
      if (cc < cols) {
        double v = 0;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) v += csum[wv][threadIdx.x];
        cpart[(long long)blockIdx.y * ldpart + cc] += v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Pass 1, emulated mode: feature map with the digit extraction fused into its epilogue.
//   Z = X.W + phi (DMMA, as k_features), Phi = s [cos Z | sin Z]                     SCFGP.py:98-102
//   out: dst[d][r][f], dst[d][r][J + f] = digit d of Phi / s2 (5 digits), rows [rows, 64 gridrows) zero;
//        cpart[row block of 64][col] += sum_r wsum[r] Phi[r][col]   (p = PhiT.y, from the float64 values)
// Neither the float64 Phi slab nor a second pass over it exists on this path.  Needs J % 4 == 0 (aligned
// 4-feature words in both halves); the host falls back to k_features + k_digits_slab otherwise.
// Tile = 64 rows x 64 features.  The DMMA accumulator layout (a thread holds feature pairs of rows 8 apart) is
// the wrong shape for the output (a digit-plane word is 4 neighbouring features of one row), so Z goes through
// shared memory once (16-byte chunks, XOR-swizzled: the fragment stores are conflict free, the row reads 2-way)
// and every thread then owns ONE row x 2 x 8 neighbouring features: 16 straight-line sincos, the byte-permute
// digit packing of the digit kernel, and 8-byte stores per digit plane and half (four lanes fill a 32-byte sector;
// 20 stores per thread instead of 80 four-byte ones), with a short code body and 80 registers (3 CTAs per SM).
//   grid = rowblocks(64) * featblocks(64)
// ------------------------------------------------------------------------------------
using CfgFeatD = TileCfg<64, 64, 32, 16, 2, 16>;   // 8 warps (2 x 4), 16 accumulators / thread; 2 stages of depth 16 (41 KB)
constexpr int FEATD_MAIN_BYTES = (int)SmemSize<CfgFeatD, true, true>::BYTES;
constexpr int FEATD_EPI_BYTES = (64 * 64 + 8 * 128) * (int)sizeof(double);   // Z tile + per-warp column sums
constexpr int FEATD_SMEM = FEATD_MAIN_BYTES > FEATD_EPI_BYTES ? FEATD_MAIN_BYTES : FEATD_EPI_BYTES;

#ifndef FEATD_MINB
#define FEATD_MINB 3
#endif
template <int XVEC>
__global__ void __launch_bounds__(CfgFeatD::NT, FEATD_MINB)
k_features_digits5(const double* __restrict__ X, long long ldx, int rows, int D, const double* __restrict__ Wt,
                   long long ldw, const double* __restrict__ phi, int J, double s, double inv_s2,
                   int8_t* __restrict__ dst, long long pitch, long long plane, const double* __restrict__ wsum,
                   double* __restrict__ cpart, long long ldpart, int nfb) {
  using Cfg = CfgFeatD;
  extern __shared__ __align__(16) double smem[];
  const int rb = blockIdx.x / nfb, fb = blockIdx.x % nfb;
  const int m0 = rb * Cfg::BM, n0 = fb * Cfg::BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    double acc[Cfg::MI][Cfg::NI][2];
    zero_acc<Cfg>(acc);
    IdentMap am{m0, rows}, bm{n0, J};
    gemm_mainloop<Cfg, true, true, XVEC, 2>(acc, X, ldx, am, Wt, ldw, bm, D, smem);   // ends with a CTA barrier
    // Z + phase -> shared memory: chunk c (features 2c, 2c+1) of row rl lies at chunk (c & ~7) | ((c & 7) ^ swz),
    // swz = (c >> 3) | ((rl & 1) << 2)
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wni = warp % Cfg::WARPS_N;
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) {
      const int f = n0 + wni * Cfg::WN + 8 * j + 2 * t;
      const double ph0 = f < J ? phi[f] : 0.0, ph1 = f + 1 < J ? phi[f + 1] : 0.0;
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i) {
        const int rl = wm0 + 8 * i + g;
        const int pc = 8 * wni + ((4 * j + t) ^ (wni | ((rl & 1) << 2)));
        *reinterpret_cast<double2*>(smem + rl * Cfg::BN + 2 * pc) = make_double2(acc[i][j][0] + ph0, acc[i][j][1] + ph1);
      }
    }
  }
  __syncthreads();
  // this thread's items: row il of the tile, features 32 hf + 8 gq .. + 7 for hf = 0, 1 (the live set of one item -- 8
  // arguments, 2 x 10 plane words -- keeps the kernel at 80 registers: 3 CTAs per SM; the 4 lanes of a row write 32
  // neighbouring bytes per plane and half: full sectors)
  const int il = threadIdx.x >> 2, gq = threadIdx.x & 3;
  const int r = m0 + il;
  const bool rv = r < rows;
  const double wr = (wsum != nullptr && rv) ? wsum[r] : 0.0;
  const double mq = digit_mult<5>() * inv_s2;          // value -> 2^39 / s2 units
  double* wpart = smem + Cfg::BM * Cfg::BN + warp * 128;   // this warp's column sums: [cos 64 | sin 64]
#pragma unroll 1
  for (int hf = 0; hf < 2; ++hf) {
    double z[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = 16 * hf + 4 * gq + k;            // chunk = feature pair; swizzle as written above
      const int pc = (c & ~7) | ((c & 7) ^ ((c >> 3) | ((il & 1) << 2)));
      const double2 v = *reinterpret_cast<const double2*>(smem + il * Cfg::BN + 2 * pc);
      z[2 * k] = v.x; z[2 * k + 1] = v.y;
    }
    bool big = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) big |= !(fabs(z[e]) <= kFastTrigMax);
    big = __any_sync(0xffffffffu, big);
    const int f8 = n0 + 32 * hf + 8 * gq;              // first feature of this half: the 4 lanes of a row cover 32 neighbours
    uint32_t cw[5][2], sw[5][2];                       // digit-plane words: [plane][4-feature group]
    auto epilogue = [&](auto slow_tag) {
      constexpr bool SLOW = decltype(slow_tag)::value;
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        double cs[4], sn[4];
        uint32_t hi[4], lo[4], w5[5];
        const double se = (rv && f8 + 4 * q2 < J) ? s : 0.0;   // rows / features outside: zero digits (J % 4 == 0)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sincos_sel<SLOW>(z[4 * q2 + e], sn[e], cs[e]);
          cs[e] *= se; sn[e] *= se;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) quant5_inrange(cs[e] * mq, hi[e], lo[e]);
        pack5x4(hi, lo, w5);
#pragma unroll
        for (int d = 0; d < 5; ++d) cw[d][q2] = w5[d];
#pragma unroll
        for (int e = 0; e < 4; ++e) quant5_inrange(sn[e] * mq, hi[e], lo[e]);
        pack5x4(hi, lo, w5);
#pragma unroll
        for (int d = 0; d < 5; ++d) sw[d][q2] = w5[d];
        if (wsum != nullptr) {
          // column sums over the warp's 8 rows (lanes with equal gq) as a reduce-scatter: every step halves the
          // values a lane carries; the lane with bits (b4, b3, b2) ends up with value 4 b4 + 2 b3 + b2 of
          // [cos 0..3 | sin 0..3].  Fixed order: results repeat bit for bit.
          const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
          double a4[4], a2[2];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double vc = wr * cs[k], vs = wr * sn[k];
            a4[k] = (b4 ? vs : vc) + __shfl_xor_sync(0xffffffffu, b4 ? vc : vs, 16);
          }
#pragma unroll
          for (int k = 0; k < 2; ++k)
            a2[k] = (b3 ? a4[2 + k] : a4[k]) + __shfl_xor_sync(0xffffffffu, b3 ? a4[k] : a4[2 + k], 8);
          const double a1 = (b2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? a2[0] : a2[1], 4);
          wpart[(b4 ? 64 : 0) + 32 * hf + 8 * gq + 4 * q2 + (b3 ? 2 : 0) + (b2 ? 1 : 0)] = a1;
        }
      }
    };
    if (big) epilogue(std::true_type{});
    else epilogue(std::false_type{});
    // digit planes: 8 neighbouring features = 8 bytes per plane and half (four lanes fill a 32-byte sector)
    if (f8 < J) {
      int8_t* oc = dst + (long long)r * pitch + f8;
      int8_t* os = oc + J;
      const bool full = f8 + 8 <= J;
      if (full && ((reinterpret_cast<uintptr_t>(oc) | (uintptr_t)plane) & 7) == 0) {
#pragma unroll
        for (int d = 0; d < 5; ++d) *reinterpret_cast<uint2*>(oc + (long long)d * plane) = make_uint2(cw[d][0], cw[d][1]);
      } else {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2)
            if (f8 + 4 * q2 < J) *reinterpret_cast<uint32_t*>(oc + (long long)d * plane + 4 * q2) = cw[d][q2];
      }
      if (full && ((reinterpret_cast<uintptr_t>(os) | (uintptr_t)plane) & 7) == 0) {
#pragma unroll
        for (int d = 0; d < 5; ++d) *reinterpret_cast<uint2*>(os + (long long)d * plane) = make_uint2(sw[d][0], sw[d][1]);
      } else {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2)
            if (f8 + 4 * q2 < J) *reinterpret_cast<uint32_t*>(os + (long long)d * plane + 4 * q2) = sw[d][q2];
      }
    }
  }
  if (wsum != nullptr) {
    __syncthreads();
    if (threadIdx.x < 2 * Cfg::BN) {
      const int half = threadIdx.x / Cfg::BN, fl = threadIdx.x % Cfg::BN;
      const int f = n0 + fl;
      if (f < J) {
        const double* wp = smem + Cfg::BM * Cfg::BN + threadIdx.x;
        double v = 0;
#pragma unroll
        for (int w = 0; w < Cfg::NWARPS; ++w) v += wp[w * 128];   // the 8 row groups of the tile, fixed order
        cpart[(long long)rb * ldpart + half * J + f] += v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Gram accumulation over one row chunk:  C[a][b] += sc * sum_r A[r,a] B[r,b]
//   operands: row-major digit arrays [KS][rows][pitch] read MN-major (their rows are the contraction
//   index); items (a block of 128, b block of 64) cover the lower triangle (blocks that cross the
//   diagonal are computed whole).   sc = scale * *scale_ptr * 2^-14
// ------------------------------------------------------------------------------------
template <int KS>
__global__ void __maxnreg__(umma::MAX_REGS)
k_gram_i8(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
          const __grid_constant__ CUtensorMap mapAt, const __grid_constant__ CUtensorMap mapBt, int row0_a,
          int row0_b, int depth, int split_len, int K, const int4* __restrict__ items, int ngram, int ntiles,
          double* __restrict__ Call, long long ldc, long long split_stride, double scale,
          const double* __restrict__ scale_ptr) {
  // MN-major operands: the maps are row-major digit arrays (feature index contiguous) whose ROWS are
  // the contraction index; the chunk's rows start at row0_a / row0_b of the A-side / B-side array.
  // work item = (row split, Gram tile): split sp reduces rows [sp split_len, (sp+1) split_len) of the
  // chunk into its own accumulator Call + sp * split_stride.
  // item.z = 1: the tile of the K mod 128 border row block, computed with the roles swapped -- the
  // narrow block is the 64-wide N side, the 128-wide M side runs over all feature blocks -- through
  // mapAt (128-byte boxes of the B-side array) / mapBt (64-byte boxes of the A-side array), and
  // stored transposed: C[b][a] += sum_r Bside[r, a] * Aside[r, b].
  extern __shared__ uint8_t smem_raw[];
  umma::Pipe p = umma::pipe_init<KS>(smem_raw);
  const uint32_t tmem_base = *p.tmem_slot;
  const bool front = (threadIdx.x >> 5) < umma::EPI_WARP0;
  int it0 = 0, seq = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++seq) {
    const int sp = tile / ngram;
    const int4 it = items[tile - sp * ngram];
    const bool tr = it.z != 0;
    double* __restrict__ C = Call + (long long)sp * split_stride;
    umma::Tile t;
    t.m0 = it.x * TM;
    t.n0 = it.y * TN;
    const int ra = tr ? row0_b : row0_a, rb = tr ? row0_a : row0_b;   // rows of the M-side / N-side arrays
    t.c_begin = ra + sp * split_len;
    t.c_end = ra + min(depth, (sp + 1) * split_len);
    t.c_off_b = rb - ra;
    if (front) {
      umma::run_mainloop<KS, true>(p, tmem_base, tr ? &mapAt : &mapA, tr ? &mapBt : &mapB, t, it0, seq);
      it0 += umma::tile_stages(t);
      continue;
    }
    const umma::EpiPos e = umma::epilogue_wait(p, tmem_base, seq & 1);
    const int a = t.m0 + e.row;
    const double sc = scale * (scale_ptr ? *scale_ptr : 1.0) * umma::PROD_SCALE;
    double v0[16], v1[16];
    umma::epilogue_load16<KS>(e, e.col0, v0);       // warp-collective: before any row guard
    umma::epilogue_load16<KS>(e, e.col0 + 16, v1);
    umma::epilogue_release(p);                      // TMEM is free: the next tile's MMAs may start
    if (a < K) {
#pragma unroll
      for (int h16 = 0; h16 < 2; ++h16) {
        const double* v = h16 ? v1 : v0;
        const int b0 = t.n0 + e.col0 + 16 * h16;
        const int nvalid = K - b0;
        if (tr) {
          // transposed store: lanes are consecutive a (coalesced along the row of C)
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < nvalid) C[(long long)(b0 + c) * ldc + a] += sc * v[c];
        } else {
          double* crow = C + (long long)a * ldc + b0;
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            if (c + 1 < nvalid) {
              double2 o = *reinterpret_cast<double2*>(crow + c);
              o.x += sc * v[c];
              o.y += sc * v[c + 1];
              *reinterpret_cast<double2*>(crow + c) = o;
            } else if (c < nvalid) {
              crow[c] += sc * v[c];
            }
          }
        }
      }
    }
  }
  umma::pipe_teardown<KS>(p, tmem_base);
}

// ------------------------------------------------------------------------------------
// v_n partials of T = Phi.LiT:  vpart[2 nb + half][n] = sum over the TNW / 2 columns of that half of T[n, .]^2
//   A: row digits of Phi (scale sA), B: row digits of Li with per-row scales liscale
//   tiles = (blocks of TNW Li rows, row blocks of 128); deepest reduction (last nb) first
// ------------------------------------------------------------------------------------
template <int KS, int TNW = TN, bool TWO_PHASE = (KS <= 5)>
__global__ void __maxnreg__(TWO_PHASE ? umma::MAX_REGS_2PHASE : umma::MAX_REGS)
k_tgemm_i8(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapLi, int row0, int rows,
           int K, int ntiles, const double* __restrict__ liscale, double sA, double* __restrict__ vpart,
           long long ldv, double* __restrict__ mu) {
  // row0: first row of this chunk inside the A array (0 for the chunk buffer, r0 for the digit cache)
  // The B operand has K + 1 rows: rows 0..K-1 are Li, row K is alpha, so column K of the product is
  // mu_n = Phi_n . alpha (SCFGP.py:111) and never enters the sum of squares.
  extern __shared__ uint8_t smem_raw[];
  umma::Pipe p = umma::pipe_init<KS, TNW>(smem_raw);
  const uint32_t tmem_base = *p.tmem_slot;
  const bool front = (threadIdx.x >> 5) < umma::EPI_WARP0;
  const int nnb = (K + 1 + TNW - 1) / TNW;
  int it0 = 0, seq = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++seq) {
    // Column-block major, deepest reduction first: tile k goes to CTA k mod gridDim, so every round of
    // tiles has (nearly) one depth and the persistent CTAs stay balanced although Li is triangular;
    // the CTAs of a round share the B operand (the same 64 rows of [Li; alpha]).
    // Row blocks go in super-blocks of 64 (8192 rows): within one, all column blocks are walked before the next
    // begins, so the A operand being re-read by the column blocks (85 MB of digits) stays L2-resident also with
    // 16384-row chunks (whole-chunk order: 1.54 GB of DRAM reads per launch, 7.7 x the operand bytes).
    const int nrb = ntiles / nnb;
    constexpr int SBR = 64;
    const int nsb = (nrb + SBR - 1) / SBR;
    const int sb = min(tile / (SBR * nnb), nsb - 1);
    const int rem = tile - sb * (SBR * nnb);
    const int rbs = sb == nsb - 1 ? nrb - sb * SBR : SBR;   // row blocks of this super-block
    const int nb = nnb - 1 - rem / rbs;
    umma::Tile t;
    const int rb0 = (sb * SBR + rem % rbs) * TM;   // chunk-local first row of the tile
    t.m0 = row0 + rb0;
    t.n0 = nb * TNW;
    t.c_begin = 0;
    t.c_end = min(K, t.n0 + TNW);  // Li is lower triangular
    if (front) {
      umma::run_mainloop<KS, false, TNW, UMMA_B_KEEP>(p, tmem_base, &mapA, &mapLi, t, it0, seq);
      it0 += umma::tile_stages(t);
      continue;
    }
    const umma::EpiPos e = umma::epilogue_wait<TNW>(p, tmem_base, seq & 1);
    const double sc = sA * umma::PROD_SCALE;
    const int r = rb0 + e.row;
    double sq = 0;
    if constexpr (TWO_PHASE) {
      // two-phase: park the recombined accumulators, hand tensor memory back, then do the math
      long long parked[TNW / 32][16];
#pragma unroll
      for (int h16 = 0; h16 < TNW / 32; ++h16) umma::epilogue_drain16<KS, TNW>(e, e.col0 + 16 * h16, parked[h16]);
      umma::epilogue_release(p);
      // (FP64 instructions wait ~300 cycles each while the next tile's MMAs run -- profiles/r2_ncu_backward_stalls.txt --
      //  so every scale is folded into one factor per column: one DMUL + one DFMA per value)
      // TWO_PHASE: `liscale` holds liscale[j] * sA * 2^-14 * 2^-8(KS-1) (host: emu_prepare_li, k_scale_vector)
#pragma unroll
      for (int h16 = 0; h16 < TNW / 32; ++h16) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int j = t.n0 + e.col0 + 16 * h16 + c;
          const double tv = j <= K ? (double)parked[h16][c] * liscale[j] : 0.0;   // = T[n][j] (mu for j == K)
          if (j == K) { if (r < rows) mu[r] = tv; }
          else sq = fma(tv, tv, sq);
        }
      }
      if (r < rows) vpart[(long long)(2 * nb + (e.col0 ? 1 : 0)) * ldv + r] = sq;
      continue;
    } else {
#pragma unroll 1
      for (int h16 = 0; h16 < TNW / 32; ++h16) {
        double v[16];
        umma::epilogue_load16<KS, TNW>(e, e.col0 + 16 * h16, v);
        if (h16 == TNW / 32 - 1) umma::epilogue_release(p);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int j = t.n0 + e.col0 + 16 * h16 + c;
          const double tv = j <= K ? v[c] * liscale[j] : 0.0;
          if (j == K) { if (r < rows) mu[r] = tv * sc; }
          else sq = fma(tv, tv, sq);
        }
      }
    }
    if (r < rows) vpart[(long long)(2 * nb + (e.col0 ? 1 : 0)) * ldv + r] = sq * sc * sc;
  }
  umma::pipe_teardown<KS, TNW>(p, tmem_base);
}

// ------------------------------------------------------------------------------------
// backward tile: 128 rows x TNW / 2 features (TNW permuted columns: cos/sin interleaved)
//   Phibar = sA 2^-14 cscale[j] * sum_c [Phi|Psi'] digits . Cmat digits  + y pbarT + gmu alphaT
//   Zbar = -Phibar_c * Phi_s + Phibar_s * Phi_c ;  slots[cta] += <Phibar, Phi>
//   tiles = (feature blocks of TNW / 2, row blocks of 128)
// ------------------------------------------------------------------------------------
template <int KS, int TNW = TN, bool TWO_PHASE = (KS <= 5)>
__global__ void __maxnreg__(TWO_PHASE ? umma::MAX_REGS_2PHASE : umma::MAX_REGS)
k_backward_i8(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA1,
              const __grid_constant__ CUtensorMap mapC, int seg1_stages, int row0, const int8_t* __restrict__ dphi,
              long long dpitch, long long dplane, double recon, int rows, int J,
              int depth2, int ntiles, const double* __restrict__ cscale, double sA, const double* __restrict__ Phi,
              long long ldp, const double* __restrict__ y, const double* __restrict__ gmu,
              const double* __restrict__ pbar, const double* __restrict__ alpha, double* __restrict__ Zbar,
              long long ldz, double* __restrict__ slots, double* __restrict__ dphi_part, long long ldpart) {
  // dphi_part[row block][f] += column sums of the tile's Zbar (dphi = colsum(Zbar), A.5)
  extern __shared__ uint8_t smem_raw[];
  __shared__ double sh[8];
  __shared__ double shcol[8][TNW / 4];   // per epilogue warp: column sums of its TNW / 4 features
  umma::Pipe p = umma::pipe_init<KS, TNW>(smem_raw);
  const uint32_t tmem_base = *p.tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool front = warp < umma::EPI_WARP0;
  constexpr int FPT = TNW / 2;             // features per tile (cos/sin column pairs)
  const int nfb = (J + FPT - 1) / FPT;
  int it0 = 0, seq = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++seq) {
    umma::Tile t;
    t.m0 = (tile / nfb) * TM;     // row-block major: consecutive tiles share the A operand in L2
    t.n0 = (tile % nfb) * TNW;
    t.c_begin = 0;
    t.c_end = depth2;
    // digit cache: the Phi half of [Phi | Psi'] (the first seg1_stages stages) comes from the cache
    // (rows row0 + ...), the Psi' half from the chunk buffer; seg1_stages == 0: both from mapA
    t.mapA1 = &mapA1;
    t.seg1_stages = seg1_stages;
    t.seg1_row_off = row0;
    if (front) {
      umma::run_mainloop<KS, false, TNW, UMMA_B_KEEP>(p, tmem_base, &mapA, &mapC, t, it0, seq);
      it0 += umma::tile_stages(t);
      continue;
    }
    unsigned long long* const tt = (threadIdx.x == umma::EPI_WARP0 * 32) ? umma::g_tile_timing : nullptr;
    const long long tt0 = tt ? clock64() : 0;
    const umma::EpiPos e = umma::epilogue_wait<TNW>(p, tmem_base, seq & 1);
    const long long tt1 = tt ? clock64() : 0;
    long long tt2 = 0;
    const int r = t.m0 + e.row;
    const bool rvalid = r < rows;
    const double sc = sA * umma::PROD_SCALE;
    const double yn = rvalid ? y[r] : 0.0, gm = rvalid ? gmu[r] : 0.0;
    const double* prow = Phi + (long long)(rvalid ? r : 0) * ldp;
    double sacc = 0;
    // two-phase (KS <= 5): park the recombined accumulators of all TNW / 2 columns, hand tensor memory back to
    // the MMA warp at once, and run the fused epilogue under the next tile's MMAs
    long long parked[TWO_PHASE ? TNW / 32 : 1][16];
    if constexpr (TWO_PHASE) {
#pragma unroll
      for (int h16 = 0; h16 < TNW / 32; ++h16) umma::epilogue_drain16<KS, TNW>(e, e.col0 + 16 * h16, parked[h16]);
      umma::epilogue_release(p);
      if (tt) tt2 = clock64();
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(umma::EPI_THREADS) : "memory");   // sh[], shcol[] of the previous tile consumed
#pragma unroll
    for (int h16 = 0; h16 < TNW / 32; ++h16) {
      double v[16];
      const int pcol = e.col0 + 16 * h16;                 // permuted tile column of v[0]
      if constexpr (TWO_PHASE) {
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = umma::drained_value<KS>(parked[h16][c]);
      } else {
        umma::epilogue_load16<KS, TNW>(e, pcol, v);
        if (h16 == TNW / 32 - 1) {
          umma::epilogue_release(p);
          if (tt) tt2 = clock64();
        }
      }
      double zb[8];
      const int f0 = (t.n0 + pcol) >> 1;
      // Phi of this row's 8 features (cos and sin halves).  With the digit cache they are rebuilt from
      // the cached digits -- the very bytes the TMA producer has just pulled through L2 for this tile --
      // as one 8-byte load per digit plane (needs J % 8 == 0: host passes dphi only then).
      double phc8[8], phs8[8];
      if (dphi != nullptr) {
        long long qc[8], qs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) qc[i] = qs[i] = 0;
        if (rvalid && f0 < J) {
          const int8_t* dp = dphi + (long long)(row0 + r) * dpitch + f0;
          if constexpr (KS == 5) {
            // five plane words per 4 features: byte permutes instead of a 5-step Horner chain per value
            uint32_t wc0[5], wc1[5], ws0[5], ws1[5];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
              const uint2 wc = *reinterpret_cast<const uint2*>(dp + (long long)d * dplane);
              const uint2 ws = *reinterpret_cast<const uint2*>(dp + (long long)d * dplane + J);
              wc0[d] = wc.x; wc1[d] = wc.y; ws0[d] = ws.x; ws1[d] = ws.y;
            }
            long long t4[4];
            unpack5x4_q(wc0, t4);
#pragma unroll
            for (int i = 0; i < 4; ++i) qc[i] = t4[i];
            unpack5x4_q(wc1, t4);
#pragma unroll
            for (int i = 0; i < 4; ++i) qc[4 + i] = t4[i];
            unpack5x4_q(ws0, t4);
#pragma unroll
            for (int i = 0; i < 4; ++i) qs[i] = t4[i];
            unpack5x4_q(ws1, t4);
#pragma unroll
            for (int i = 0; i < 4; ++i) qs[4 + i] = t4[i];
          } else {
#pragma unroll
            for (int d = 0; d < KS; ++d) {
              const uint2 wc = *reinterpret_cast<const uint2*>(dp + (long long)d * dplane);
              const uint2 ws = *reinterpret_cast<const uint2*>(dp + (long long)d * dplane + J);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t bc = (i < 4 ? wc.x : wc.y) >> (8 * (i & 3));
                const uint32_t bs = (i < 4 ? ws.x : ws.y) >> (8 * (i & 3));
                qc[i] = qc[i] * 256 + (long long)(int8_t)bc;
                qs[i] = qs[i] * 256 + (long long)(int8_t)bs;
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { phc8[i] = (double)qc[i] * recon; phs8[i] = (double)qs[i] * recon; }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = rvalid && f0 + i < J;
          phc8[i] = ok ? prow[f0 + i] : 0.0;
          phs8[i] = ok ? prow[J + f0 + i] : 0.0;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = f0 + i;
        zb[i] = 0.0;
        if (rvalid && f < J) {
          const int pj = t.n0 + pcol + 2 * i;
          const double pc = v[2 * i] * (cscale[pj] * sc) + yn * pbar[f] + gm * alpha[f];
          const double ps = v[2 * i + 1] * (cscale[pj + 1] * sc) + yn * pbar[J + f] + gm * alpha[J + f];
          const double phc = phc8[i], phs = phs8[i];
          zb[i] = -pc * phs + ps * phc;
          Zbar[(long long)r * ldz + f] = zb[i];
          sacc += pc * phc + ps * phs;
        }
      }
      // column sums over this warp's 32 rows (features f0 .. f0+7) as a reduce-scatter: every step halves the
      // values a lane carries (9 additions and 9 shuffles per lane instead of 40 for eight butterflies); lanes
      // 4 k .. 4 k + 3 end up with the sum of feature f0 + k.  Fixed order: results repeat bit for bit.
      {
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        double a4[4], a2[2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double keep = b4 ? zb[4 + k] : zb[k], give = b4 ? zb[k] : zb[4 + k];
          a4[k] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const double keep = b3 ? a4[2 + k] : a4[k], give = b3 ? a4[k] : a4[2 + k];
          a2[k] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
        }
        double cs = (b2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? a2[0] : a2[1], 4);
        cs += __shfl_xor_sync(0xffffffffu, cs, 2);
        cs += __shfl_xor_sync(0xffffffffu, cs, 1);
        if ((lane & 3) == 0) shcol[warp - umma::EPI_WARP0][h16 * 8 + (lane >> 2)] = cs;
      }
    }
    sacc = warp_sum(sacc);
    if (lane == 0) sh[warp - umma::EPI_WARP0] = sacc;
    asm volatile("bar.sync 1, %0;\n" ::"n"(umma::EPI_THREADS) : "memory");
    if (warp == umma::EPI_WARP0) {
      if (lane == 0) {
        double sum = 0;
        for (int w = 0; w < 8; ++w) sum += sh[w];
        slots[tile] += sum;
      }
      // warps w and w+4 hold column halves 0 / 1; their 4 lane quarters are the 4 row groups
      for (int fi = lane; fi < FPT; fi += 32) {           // features of the tile: half * (FPT / 2) + k
        const int half = fi / (FPT / 2), k = fi % (FPT / 2);
        const int f = (t.n0 >> 1) + fi;
        if (f < J) {
          double cs = 0;
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) cs += shcol[half * 4 + qd][k];   // the 4 row quarters, fixed order
          dphi_part[(long long)(tile / nfb) * ldpart + f] += cs;
        }
      }
    }
    if (tt) {
      umma::tt_add(tt, 4, tt1 - tt0);
      umma::tt_add(tt, 5, tt2 - tt1);
      umma::tt_add(tt, 6, clock64() - tt2);
    }
  }
  umma::pipe_teardown<KS, TNW>(p, tmem_base);
}

// ------------------------------------------------------------------------------------
// backward tile, two-phase epilogue (KS <= 5; the default).  Same Zbar as k_backward_i8.  What bounds this epilogue
// is the FP64 pipe: while the tcgen05.mma stream of the next tile runs, every DMUL / DFMA / DADD of the epilogue
// warps waits for the pipe for ~300 cycles ("math pipe throttle" is half of all stall samples,
// profiles/r2_ncu_backward_stalls.txt), so the kernel is built around the FP64 instruction count:
//   * the accumulators are recombined exactly in int64 on the integer pipes and parked, tensor memory goes back
//     to the MMA warp at once (phase 1);
//   * all scales are folded into three per-column vectors staged in shared memory (cscale sA 2^-14 2^-32 recon,
//     pbar recon, alpha recon): Phibar costs one DMUL + two DFMA per value, Zbar one DMUL + one DFMA per feature
//     on the integer-valued Phi digits;
//   * the reductions left the kernel: sbar = sum <Phibar, Phi> is a pair of K x K traces in the dense backward
//     phase (k_sbar_rows), dphi = colsum(Zbar) is summed by the dW contraction from its operand fragments (k_dw_splitk).
// ------------------------------------------------------------------------------------
struct PhiWords5 { uint32_t c0[5], c1[5], s0[5], s1[5]; };   // digit planes of 8 cos + 8 sin features of one row

template <int KS, int TNW>
__global__ void __maxnreg__(umma::MAX_REGS_2PHASE)
k_backward_i8_2p(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapC, int seg1_stages, int row0, const int8_t* __restrict__ dphi,
                 long long dpitch, long long dplane, double recon, int rows, int J,
                 int depth2, int ntiles, const double* __restrict__ cscale, double sA, const double* __restrict__ Phi,
                 long long ldp, const double* __restrict__ y, const double* __restrict__ gmu,
                 const double* __restrict__ pbar, const double* __restrict__ alpha, double* __restrict__ Zbar,
                 long long ldz) {
  static_assert(KS <= 5, "two-phase epilogue: exact int64 recombination");
  extern __shared__ uint8_t smem_raw[];
  __shared__ double colv[3][TNW];        // per permuted tile column: cscale * sc * 2^-8(KS-1) * rs, pbar * rs, alpha * rs
  umma::Pipe p = umma::pipe_init<KS, TNW>(smem_raw);
  const uint32_t tmem_base = *p.tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool front = warp < umma::EPI_WARP0;
  constexpr int FPT = TNW / 2;             // features per tile (cos/sin column pairs)
  constexpr int NG = TNW / 32;             // groups of 16 columns (8 features) per thread
  const int nfb = (J + FPT - 1) / FPT;
  // Phi enters Zbar as (integer digits value q) * recon when it comes from the digit cache, as the float64 value
  // itself otherwise: rs is the factor that is folded into the column vectors
  const double rs = dphi != nullptr ? recon : 1.0;
  const double sc = sA * umma::PROD_SCALE * __longlong_as_double((long long)(1023 - umma::DIGIT_BITS * (KS - 1)) << 52) * rs;
  int it0 = 0, seq = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++seq) {
    umma::Tile t;
    t.m0 = (tile / nfb) * TM;     // row-block major: consecutive tiles share the A operand in L2
    t.n0 = (tile % nfb) * TNW;
    t.c_begin = 0;
    t.c_end = depth2;
    t.mapA1 = &mapA1;
    t.seg1_stages = seg1_stages;
    t.seg1_row_off = row0;
    if (front) {
      umma::run_mainloop<KS, false, TNW, UMMA_B_KEEP>(p, tmem_base, &mapA, &mapC, t, it0, seq);
      it0 += umma::tile_stages(t);
      continue;
    }
    unsigned long long* const tt = (threadIdx.x == umma::EPI_WARP0 * 32) ? umma::g_tile_timing : nullptr;
    const long long tt0 = tt ? clock64() : 0;
    // ---- 1. while the tile's MMAs run: row scalars, column vectors, first Phi digits towards L1
    const int quarter = warp & 3, half = (warp - umma::EPI_WARP0) >> 2;
    const int r = t.m0 + quarter * 32 + lane;
    const int col0 = half * (TNW / 2);
    const bool rvalid = r < rows;
    const double yn = rvalid ? y[r] : 0.0, gm = rvalid ? gmu[r] : 0.0;
    const double* prow = Phi + (long long)(rvalid ? r : 0) * ldp;
    const int8_t* drow = dphi ? dphi + (long long)(row0 + (rvalid ? r : 0)) * dpitch : nullptr;
    auto load_words = [&](int g, PhiWords5& w) {
      const int f0 = (t.n0 + col0 + 16 * g) >> 1;
      const bool ok = rvalid && f0 < J;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        uint2 wc = make_uint2(0x80808080u, 0x80808080u), ws = wc;   // (planes 1..4 of q = 0; plane 0 fixed below)
        if (d == 0) wc = ws = make_uint2(0u, 0u);
        if (ok && d < KS) {
          wc = *reinterpret_cast<const uint2*>(drow + (long long)d * dplane + f0);
          ws = *reinterpret_cast<const uint2*>(drow + (long long)d * dplane + J + f0);
        }
        w.c0[d] = wc.x; w.c1[d] = wc.y; w.s0[d] = ws.x; w.s1[d] = ws.y;
      }
    };
    // the next group's digit words are pulled into L1 while the current group is processed (no registers held)
    auto prefetch_words = [&](int g) {
      const int f0 = (t.n0 + col0 + 16 * g) >> 1;
      if (rvalid && f0 < J) {
#pragma unroll
        for (int d = 0; d < KS; ++d) {
          asm volatile("prefetch.global.L1 [%0];\n" ::"l"(drow + (long long)d * dplane + f0));
          asm volatile("prefetch.global.L1 [%0];\n" ::"l"(drow + (long long)d * dplane + J + f0));
        }
      }
    };
    asm volatile("bar.sync 1, %0;\n" ::"n"(umma::EPI_THREADS) : "memory");   // everyone is done with the previous tile
    for (int c = threadIdx.x - umma::EPI_WARP0 * 32; c < 3 * TNW; c += umma::EPI_THREADS) {
      const int which = c / TNW, cc = c - which * TNW;
      const int pj = t.n0 + cc, f = pj >> 1, hf = pj & 1;
      double v = 0.0;
      if (f < J) v = which == 0 ? cscale[pj] * sc : (which == 1 ? pbar[hf * J + f] : alpha[hf * J + f]) * rs;
      colv[which][cc] = v;
    }
    if (dphi != nullptr) prefetch_words(0);
    asm volatile("bar.sync 1, %0;\n" ::"n"(umma::EPI_THREADS) : "memory");   // column vectors visible
    // ---- 2. drain
    const umma::EpiPos e = umma::epilogue_wait<TNW>(p, tmem_base, seq & 1);
    const long long tt1 = tt ? clock64() : 0;
    long long parked[NG][16];
#pragma unroll
    for (int g = 0; g < NG; ++g) umma::epilogue_drain16<KS, TNW>(e, col0 + 16 * g, parked[g]);
    umma::epilogue_release(p);
    const long long tt2 = tt ? clock64() : 0;
    // ---- 3. Zbar from the parked values
    PhiWords5 wcur;
    if (dphi != nullptr) load_words(0, wcur);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int pcol = col0 + 16 * g;                     // permuted tile column of this group's first value
      const int f0 = (t.n0 + pcol) >> 1;
      if (dphi != nullptr && g + 1 < NG) prefetch_words(g + 1);      // in flight during this group's math
      double zb[8];
#pragma unroll
      for (int hq = 0; hq < 2; ++hq) {                    // four features at a time (keeps the live set small)
        double phc4[4], phs4[4];                          // Phi / rs of the row's features
        if (dphi != nullptr) {
          if constexpr (KS == 5) {
            long long q4[4];
            unpack5x4_q(hq ? wcur.c1 : wcur.c0, q4);
#pragma unroll
            for (int i = 0; i < 4; ++i) phc4[i] = (double)q4[i];
            unpack5x4_q(hq ? wcur.s1 : wcur.s0, q4);
#pragma unroll
            for (int i = 0; i < 4; ++i) phs4[i] = (double)q4[i];
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              long long qc = 0, qs = 0;
#pragma unroll
              for (int d = 0; d < KS; ++d) {
                qc = qc * 256 + (long long)(int8_t)((hq ? wcur.c1[d] : wcur.c0[d]) >> (8 * i));
                qs = qs * 256 + (long long)(int8_t)((hq ? wcur.s1[d] : wcur.s0[d]) >> (8 * i));
              }
              phc4[i] = (double)qc; phs4[i] = (double)qs;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = rvalid && f0 + 4 * hq + i < J;
            phc4[i] = ok ? prow[f0 + 4 * hq + i] : 0.0;
            phs4[i] = ok ? prow[J + f0 + 4 * hq + i] : 0.0;
          }
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int i = 4 * hq + i4;
          const int cc = pcol + 2 * i;
          // Phibar * rs of the cos and sin columns (exact int64 -> float64: one rounding)
          const double pc = fma(gm, colv[2][cc], fma(yn, colv[1][cc], (double)parked[g][2 * i] * colv[0][cc]));
          const double ps = fma(gm, colv[2][cc + 1], fma(yn, colv[1][cc + 1], (double)parked[g][2 * i + 1] * colv[0][cc + 1]));
          zb[i] = fma(ps, phc4[i4], -(pc * phs4[i4]));
        }
      }
      if (rvalid && f0 < J) {                             // 8 neighbouring features of one row: 16-byte stores
        double* zrow = Zbar + (long long)r * ldz + f0;
        if (f0 + 7 < J && (ldz & 1) == 0) {
#pragma unroll
          for (int i = 0; i < 8; i += 2) *reinterpret_cast<double2*>(zrow + i) = make_double2(zb[i], zb[i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (f0 + i < J) zrow[i] = zb[i];
        }
      }
      if (dphi != nullptr && g + 1 < NG) load_words(g + 1, wcur);
    }
    if (tt) {
      umma::tt_add(tt, 4, tt1 - tt0);
      umma::tt_add(tt, 5, tt2 - tt1);
      umma::tt_add(tt, 6, clock64() - tt2);
    }
  }
  umma::pipe_teardown<KS, TNW>(p, tmem_base);
}

// generic C[m][n] = sum_c A[m,c] B[n,c] with per-row scales on both sides (self-test hook)
template <int KS>
__global__ void __maxnreg__(umma::MAX_REGS)
k_gemm_i8_dbg(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int M, int Nn,
              int depth, const double* __restrict__ ascale, const double* __restrict__ bscale, double* __restrict__ C,
              long long ldc) {
  extern __shared__ uint8_t smem_raw[];
  umma::Pipe p = umma::pipe_init<KS>(smem_raw);
  const uint32_t tmem_base = *p.tmem_slot;
  umma::Tile t;
  t.m0 = blockIdx.y * TM;
  t.n0 = blockIdx.x * TN;
  t.c_begin = 0;
  t.c_end = depth;
  if ((threadIdx.x >> 5) < umma::EPI_WARP0) {
    umma::run_mainloop<KS>(p, tmem_base, &mapA, &mapB, t, 0);
  } else {
    const umma::EpiPos e = umma::epilogue_wait(p, tmem_base, 0);
    const int r = t.m0 + e.row;
#pragma unroll 1
    for (int h16 = 0; h16 < 2; ++h16) {
      double v[16];
      umma::epilogue_load16<KS>(e, e.col0 + 16 * h16, v);
      if (r < M) {
        const double sa = ascale[r] * umma::PROD_SCALE;
        for (int c = 0; c < 16; ++c) {
          const int j = t.n0 + e.col0 + 16 * h16 + c;
          if (j < Nn) C[(long long)r * ldc + j] = v[c] * sa * bscale[j];
        }
      }
    }
  }
  umma::pipe_teardown<KS>(p, tmem_base);
}

}  // namespace emu
}  // namespace scfgp
