// Column-wise data preprocessing of the reference (SCFGP/Scaler.py:39-135) streamed on the
// device: row f3 of SURVEY.md section 8.  At N = 1e7..1e8 rows the reference's host loop (a
// SciPy SLSQP search per column whose objective re-reads the whole column through NumPy) is the
// wall-clock bottleneck of set_data; here every pass over the data is a bandwidth-bound kernel
// and only the scalar line search stays on the host (same SLSQP driver, so the fitted Box-Cox
// exponents follow the reference's).
//
// All reductions write per-block partials that are summed in a fixed order (deterministic).
#include <cmath>
#include <cstdio>
#include <vector>
#include <algorithm>

#include <cuda_runtime.h>

#include "../../include/scfgp_b200.h"

namespace {

constexpr int kBlocks = 592;   // 4 x 148 SMs
constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Box-Cox with the reference's sign convention (Scaler.py:61,84):  (sign(u) |u|^lm - 1) / lm
__device__ __forceinline__ double boxcox(double u, double lm) {
  const double p = u == 0.0 ? 0.0 : copysign(pow(fabs(u), lm), u);
  return (p - 1.0) / lm;
}

// per-column min / max / "equal to the first row everywhere" (Scaler.py:40-41,44-45)
//   grid = (column tiles of 32, row splits); part[split][3][D]
__global__ void k_colstats(const double* __restrict__ X, long long N, int D, long long ld,
                           double* __restrict__ part, int nsplit) {
  const int d = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rlane = threadIdx.x >> 5, nr = blockDim.x >> 5;
  const long long per = (N + nsplit - 1) / nsplit;
  const long long r0 = (long long)blockIdx.y * per, r1 = min(N, r0 + per);
  double mn = INFINITY, mx = -INFINITY, same = 1.0;
  if (d < D) {
    const double first = X[d];
    for (long long r = r0 + rlane; r < r1; r += nr) {
      const double v = X[r * ld + d];
      mn = fmin(mn, v); mx = fmax(mx, v);
      if (v != first) same = 0.0;
    }
  }
  __shared__ double sh[3][8][32];
  sh[0][rlane][threadIdx.x & 31] = mn; sh[1][rlane][threadIdx.x & 31] = mx; sh[2][rlane][threadIdx.x & 31] = same;
  __syncthreads();
  if (rlane == 0 && d < D) {
    for (int w = 1; w < nr; ++w) {
      mn = fmin(mn, sh[0][w][threadIdx.x]); mx = fmax(mx, sh[1][w][threadIdx.x]); same = fmin(same, sh[2][w][threadIdx.x]);
    }
    part[((long long)blockIdx.y * 3 + 0) * D + d] = mn;
    part[((long long)blockIdx.y * 3 + 1) * D + d] = mx;
    part[((long long)blockIdx.y * 3 + 2) * D + d] = same;
  }
}

// contiguous copy of one column, min-max scaled: col[r] = (X[r][d] - mn) * inv   (inv = 1/(mx-mn), or 1 with mn = 0)
__global__ void k_gather_col(const double* __restrict__ X, long long N, long long ld, int d, double mn, double rng,
                             double* __restrict__ col) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x)
    col[r] = (X[r * ld + d] - mn) / rng;
}

// part[b] = sum over the block's elements of g(col[r])^p,  p = 1 (mode 0) or the central
// powers 2 and 3 about `centre` (mode 1: part[b], part[nb + b]);  g = Box-Cox(lm) or identity (lm <= 0)
__global__ void k_col_moments(const double* __restrict__ col, long long N, double lm, int mode, double centre,
                              double* __restrict__ part) {
  double a = 0, b = 0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
    const double z = lm > 0 ? boxcox(col[r], lm) : col[r];
    if (mode == 0) a += z;
    else { const double t = z - centre; a += t * t; b += t * t * t; }
  }
  __shared__ double sh[2][32];
  a = warp_sum_d(a); b = warp_sum_d(b);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; }
    part[blockIdx.x] = a;
    part[gridDim.x + blockIdx.x] = b;
  }
}

// up to `cap` distinct values met by each thread in its strided slice of the column (NaN-free data);
// out[t][0] = count (cap + 1 = "more than cap"), out[t][1..cap] = the values
__global__ void k_col_distinct(const double* __restrict__ col, long long N, int cap, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  double vals[16];
  int n = 0;
  for (long long r = t; r < N && n <= cap; r += stride) {
    const double v = col[r];
    bool seen = false;
    for (int i = 0; i < n; ++i) seen |= (vals[i] == v);
    if (!seen) { if (n < 16) vals[n] = v; ++n; }
  }
  double* o = out + t * (cap + 1);
  o[0] = (double)min(n, cap + 1);
  for (int i = 0; i < cap; ++i) o[1 + i] = i < n ? vals[i] : 0.0;
}

// forward transforms of Scaler.forward_transform (Scaler.py:96-114) on the selected columns
//   algo: 0 min-max, 1 normal, 2 inv-normal, 3 auto-normal, 4 auto-inv-normal
__global__ void k_forward(const double* __restrict__ X, long long N, long long ld, const int* __restrict__ cols,
                          int ncols, int algo, const double* __restrict__ mn, const double* __restrict__ mx,
                          const double* __restrict__ lm, const double* __restrict__ mu, const double* __restrict__ sd,
                          double* __restrict__ out, long long ldo) {
  const long long total = N * ncols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / ncols;
    const int j = (int)(e % ncols);
    double v = X[r * ld + cols[j]];
    if (algo == 0) v = (v - mn[j]) / (mx[j] - mn[j]);
    else if (algo == 1) v = (v - mu[j]) / sd[j];
    else if (algo == 2) v = 0.5 * erfc(-((v - mu[j]) / sd[j]) * 0.70710678118654752440);
    else {
      v = boxcox((v - mn[j]) / (mx[j] - mn[j]), lm[j]);
      if (algo == 3) v = (v - mu[j]) / sd[j];
      else v = 0.5 * erfc(-((v - mu[j]) / sd[j]) * 0.70710678118654752440);
    }
    out[r * ldo + j] = v;
  }
}

#define SCK(call)                                   \
  do {                                              \
    if ((call) != cudaSuccess) return SCFGP_ERR_CUDA; \
  } while (0)

double sum_fixed(const std::vector<double>& v, int off, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += v[off + i];
  return s;
}

}  // namespace

extern "C" {

int scfgp_scaler_colstats(int device, const double* dX, long long N, int D, long long ld, double* mn, double* mx,
                          int* is_const) {
  if (!dX || N <= 0 || D <= 0 || ld < D || !mn || !mx || !is_const) return SCFGP_ERR_ARG;
  SCK(cudaSetDevice(device));
  const int nsplit = 296;
  double* dpart = nullptr;
  SCK(cudaMalloc((void**)&dpart, (size_t)nsplit * 3 * D * sizeof(double)));
  k_colstats<<<dim3((D + 31) / 32, nsplit), 256>>>(dX, N, D, ld, dpart, nsplit);
  std::vector<double> part((size_t)nsplit * 3 * D);
  cudaError_t e = cudaMemcpy(part.data(), dpart, part.size() * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dpart);
  if (e != cudaSuccess) return SCFGP_ERR_CUDA;
  for (int d = 0; d < D; ++d) {
    double a = INFINITY, b = -INFINITY, same = 1.0;
    for (int s = 0; s < nsplit; ++s) {
      a = std::min(a, part[((size_t)s * 3 + 0) * D + d]);
      b = std::max(b, part[((size_t)s * 3 + 1) * D + d]);
      same = std::min(same, part[((size_t)s * 3 + 2) * D + d]);
    }
    mn[d] = a; mx[d] = b; is_const[d] = same != 0.0;
  }
  return SCFGP_OK;
}

int scfgp_scaler_gather_column(int device, const double* dX, long long N, long long ld, int col, double shift,
                               double range, double* dcol) {
  if (!dX || !dcol || N <= 0 || col < 0 || col >= ld || range == 0.0) return SCFGP_ERR_ARG;
  SCK(cudaSetDevice(device));
  k_gather_col<<<kBlocks, kThreads>>>(dX, N, ld, col, shift, range, dcol);
  SCK(cudaGetLastError());
  return SCFGP_OK;
}

static int column_sums(int device, const double* dcol, long long N, double lambda, int mode, double centre,
                       double* s1, double* s2) {
  SCK(cudaSetDevice(device));
  static double* part_cache[64] = {nullptr};   // one small partial buffer per device, kept for the process
  if (device < 0 || device >= 64) return SCFGP_ERR_ARG;
  if (!part_cache[device]) SCK(cudaMalloc((void**)&part_cache[device], 2 * kBlocks * sizeof(double)));
  double* dpart = part_cache[device];
  std::vector<double> part(2 * kBlocks);
  k_col_moments<<<kBlocks, kThreads>>>(dcol, N, lambda, mode, centre, dpart);
  SCK(cudaMemcpy(part.data(), dpart, 2 * kBlocks * sizeof(double), cudaMemcpyDeviceToHost));
  *s1 = sum_fixed(part, 0, kBlocks);
  *s2 = sum_fixed(part, kBlocks, kBlocks);
  return SCFGP_OK;
}

int scfgp_scaler_column_sums(int device, const double* dcol, long long N, double lambda, int central, double centre,
                             double* s1, double* s2) {
  if (!dcol || N <= 0 || !s1 || !s2) return SCFGP_ERR_ARG;
  return column_sums(device, dcol, N, lambda, central ? 1 : 0, centre, s1, s2);
}

int scfgp_scaler_column_moments(int device, const double* dcol, long long N, double lambda, double* mean,
                                double* var, double* skew_g1) {
  if (!dcol || N <= 0 || !mean || !var || !skew_g1) return SCFGP_ERR_ARG;
  double a = 0, b = 0;
  int rc = column_sums(device, dcol, N, lambda, 0, 0.0, &a, &b);
  if (rc != SCFGP_OK) return rc;
  const double m = a / (double)N;
  rc = column_sums(device, dcol, N, lambda, 1, m, &a, &b);
  if (rc != SCFGP_OK) return rc;
  const double m2 = a / (double)N;
  const double m3 = b / (double)N;
  *mean = m;
  *var = m2;
  *skew_g1 = m2 > 0 ? m3 / (m2 * std::sqrt(m2)) : 0.0;
  return SCFGP_OK;
}

int scfgp_scaler_column_distinct_values(int device, const double* dcol, long long N, int cap, int* ndistinct,
                                        double* values) {
  if (!dcol || N <= 0 || cap < 1 || cap > 15 || !ndistinct) return SCFGP_ERR_ARG;
  SCK(cudaSetDevice(device));
  const int nb = 64, nt = 128;
  double* dout = nullptr;
  SCK(cudaMalloc((void**)&dout, (size_t)nb * nt * (cap + 1) * sizeof(double)));
  k_col_distinct<<<nb, nt>>>(dcol, N, cap, dout);
  std::vector<double> out((size_t)nb * nt * (cap + 1));
  cudaError_t e = cudaMemcpy(out.data(), dout, out.size() * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dout);
  if (e != cudaSuccess) return SCFGP_ERR_CUDA;
  std::vector<double> vals;
  for (long long t = 0; t < (long long)nb * nt; ++t) {
    const double* o = &out[(size_t)t * (cap + 1)];
    const int n = (int)o[0];
    if (n > cap) { *ndistinct = cap + 1; return SCFGP_OK; }
    for (int i = 0; i < n; ++i) vals.push_back(o[1 + i]);
  }
  std::sort(vals.begin(), vals.end());
  const int nd = (int)(std::unique(vals.begin(), vals.end()) - vals.begin());
  *ndistinct = std::min(nd, cap + 1);
  if (values && nd <= cap) std::copy(vals.begin(), vals.begin() + nd, values);
  return SCFGP_OK;
}

int scfgp_scaler_column_distinct(int device, const double* dcol, long long N, int cap, int* ndistinct) {
  return scfgp_scaler_column_distinct_values(device, dcol, N, cap, ndistinct, nullptr);
}

int scfgp_scaler_forward(int device, int algo, const double* dX, long long N, long long ld, const int* cols, int ncols,
                         const double* mn, const double* mx, const double* lambda, const double* mu,
                         const double* sd, double* dOut, long long ldo) {
  if (!dX || !dOut || !cols || N <= 0 || ncols <= 0 || algo < 0 || algo > 4 || ldo < ncols) return SCFGP_ERR_ARG;
  SCK(cudaSetDevice(device));
  int* dcols = nullptr;
  double* dpar = nullptr;
  SCK(cudaMalloc((void**)&dcols, ncols * sizeof(int)));
  SCK(cudaMalloc((void**)&dpar, (size_t)5 * ncols * sizeof(double)));
  std::vector<double> par((size_t)5 * ncols, 1.0);
  const double* src[5] = {mn, mx, lambda, mu, sd};
  for (int k = 0; k < 5; ++k)
    if (src[k]) std::copy(src[k], src[k] + ncols, par.begin() + (size_t)k * ncols);
  cudaError_t e = cudaMemcpy(dcols, cols, ncols * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dpar, par.data(), par.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_forward<<<kBlocks * 2, kThreads>>>(dX, N, ld, dcols, ncols, algo, dpar, dpar + ncols, dpar + 2 * ncols,
                                         dpar + 3 * ncols, dpar + 4 * ncols, dOut, ldo);
    e = cudaDeviceSynchronize();
  }
  cudaFree(dcols);
  cudaFree(dpar);
  return e == cudaSuccess ? SCFGP_OK : SCFGP_ERR_CUDA;
}

}  // extern "C"
