// Bit-for-bit check of the 5-digit fast paths of emu_kernels.cuh (quant5 / pack5x4 / unpack5x4) against the
// generic digit code (biased_q / digit_byte / fix_bytes and the int64 reconstruction) on the GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/digit_pack_check tools/digit_pack_check.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>

#include "../scfgp_b200/csrc/emu_kernels.cuh"

using namespace scfgp;
using namespace scfgp::emu;

__global__ void k_check(const double* __restrict__ v, int n4, unsigned long long* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  double x[4];
  for (int e = 0; e < 4; ++e) x[e] = v[4 * i + e];
  // generic
  uint32_t packed[5] = {0, 0, 0, 0, 0};
  for (int e = 0; e < 4; ++e) {
    const long long qp = biased_q<5>(x[e]);
    for (int s = 0; s < 5; ++s) packed[s] |= digit_byte<5>(qp, s) << (8 * e);
  }
  for (int s = 1; s < 5; ++s) packed[s] = fix_bytes(packed[s]);
  // fast
  uint32_t hi[4], lo[4], w[5];
  for (int e = 0; e < 4; ++e) quant5(x[e], hi[e], lo[e]);
  pack5x4(hi, lo, w);
  bool ok = true;
  for (int s = 0; s < 5; ++s) ok = ok && (w[s] == packed[s]);
  // reconstruction
  long long q[4] = {0, 0, 0, 0};
  for (int s = 0; s < 5; ++s)
    for (int e = 0; e < 4; ++e) q[e] = q[e] * 256 + (long long)(int8_t)(packed[s] >> (8 * e));
  double qd[4];
  unpack5x4(w, qd);
  for (int e = 0; e < 4; ++e) ok = ok && ((double)q[e] == qd[e]);
  if (!ok) atomicAdd(bad, 1ull);
}

int main() {
  const int n = 1 << 24;
  std::vector<double> h(n);
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  const double lim = ldexp(1.0, 38);
  for (int i = 0; i < n; ++i) {
    const int kind = i & 7;
    double v = u(rng) * lim;
    if (kind == 1) v = std::floor(v) + 0.5;                 // ties
    if (kind == 2) v = u(rng) * 300.0;                       // small magnitudes, sign changes of every digit
    if (kind == 3) v = std::floor(u(rng) * 70000.0) * 128.0 + (i % 3 - 1) * 0.5;   // digit boundaries
    if (kind == 4) v = (i & 8) ? lim : -lim;
    if (kind == 5) v = u(rng) * lim * 4.0;                   // beyond the clamp
    h[i] = v;
  }
  h[0] = 0.0; h[1] = -0.0; h[2] = NAN; h[3] = INFINITY; h[4] = -INFINITY; h[5] = 0.5; h[6] = -0.5; h[7] = 1.5;
  double* d; unsigned long long* bad; unsigned long long hb = 0;
  cudaMalloc(&d, n * sizeof(double)); cudaMalloc(&bad, 8);
  cudaMemcpy(d, h.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(bad, &hb, 8, cudaMemcpyHostToDevice);
  k_check<<<(n / 4 + 255) / 256, 256>>>(d, n / 4, bad);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
  printf("digit_pack_check: %d values, %llu mismatching groups, %s\n", n, hb, cudaGetErrorString(e));
  return (hb == 0 && e == cudaSuccess) ? 0 : 1;
}
