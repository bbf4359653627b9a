// FP64 pipe microbenchmark for B200: register-resident DMMA.8x8x4 and DFMA issue rates.
// Gives the in-kernel ceiling the hot-path roofline is judged against (next to the
// cuBLAS DGEMM number measured from Python).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_dmma(double* out, int iters) {
  double c[NACC][2];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters) {
  double c[NACC];
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_sincos(double* out, int iters) {
  double x = threadIdx.x * 0.37 + blockIdx.x, acc = 0;
  for (int it = 0; it < iters; ++it) {
    double s, c;
    sincos(x, &s, &c);
    acc += s * c;
    x += 1.2345;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();  // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int ctas : {1, 2}) {
      if (warps * ctas > 64) continue;
      float ms = time_ms([&] { k_dmma<16><<<sms * ctas, warps * 32>>>(out, iters); });
      double flops = 2.0 * 256 * 16 * (double)iters * warps * ctas * sms;
      printf("dmma884 nacc=16 warps/cta=%d ctas/sm=%d : %.2f ms  %.2f TFLOP/s\n", warps, ctas, ms, flops / ms * 1e-9);
    }
  }
  for (int warps : {8, 16}) {
    float ms = time_ms([&] { k_dmma<32><<<sms, warps * 32>>>(out, iters); });
    double flops = 2.0 * 256 * 32 * (double)iters * warps * sms;
    printf("dmma884 nacc=32 warps/cta=%d : %.2f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
    ms = time_ms([&] { k_dmma<4><<<sms, warps * 32>>>(out, iters); });
    flops = 2.0 * 256 * 4 * (double)iters * warps * sms;
    printf("dmma884 nacc=4  warps/cta=%d : %.2f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  for (int warps : {8, 16, 32}) {
    float ms = time_ms([&] { k_dfma<16><<<sms * 2, warps * 32>>>(out, iters); });
    double flops = 2.0 * 32 * 16 * (double)iters * warps * 2 * sms;
    printf("dfma nacc=16 warps/cta=%d ctas/sm=2 : %.2f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  {
    const int it2 = 2000;
    float ms = time_ms([&] { k_sincos<<<sms * 4, 256>>>(out, it2); });
    double n = (double)it2 * sms * 4 * 256;
    printf("sincos(double): %.2f ms  %.2f G sincos/s\n", ms, n / ms * 1e-6);
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
