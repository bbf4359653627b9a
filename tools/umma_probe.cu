// Stand-alone probe of the int8 tcgen05 mainloop (scfgp_b200/csrc/umma_i8.cuh):
//   * exactness of every class accumulator against a CPU integer reference
//   * throughput of the digit-product pipeline at the cfg4 pass-3 geometry
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/umma_probe tools/umma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>

#include "../scfgp_b200/csrc/umma_i8.cuh"

using namespace scfgp;
using namespace scfgp::umma;

#define CHECK(x)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);         \
      exit(2);                                                                                \
    }                                                                                         \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return (EncodeTiledFn)p;
}
// digits [KS][rows][pitch] int8, contraction contiguous
static CUtensorMap make_map(const int8_t* base, int C, int rows, long long pitch, int KS, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)KS};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
  cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)box_rows, (cuuint32_t)KS};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
  return m;
}

// MN-major digit arrays [KS][depth rows][pitch] (M or N index contiguous): box = (128 or 64 bytes, 64 rows, KS)
static CUtensorMap make_map_mn(const int8_t* base, int width, int depth, long long pitch, int KS, int box_w) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)depth, (cuuint64_t)KS};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * depth};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)KB, (cuuint32_t)KS};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE,
                            box_w == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(mn) failed %d\n", (int)r); exit(2); }
  return m;
}

// one tile per CTA; dumps raw class accumulators and the combined float64 value
template <int KS, bool TR = false>
__global__ void __maxnreg__(MAX_REGS)
k_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int c_begin, int c_end,
        int M, int Nn, int* raw /*[KS][M][Nn] or null*/, double* out /*[M][Nn]*/) {
  extern __shared__ uint8_t smem_raw[];
  Pipe p = pipe_init<KS>(smem_raw);
  const uint32_t tmem_base = *p.tmem_slot;
  Tile t;
  t.m0 = blockIdx.y * TM;
  t.n0 = blockIdx.x * TN;
  t.c_begin = c_begin;
  t.c_end = c_end;
  const int warp = threadIdx.x >> 5;
  if (warp < EPI_WARP0) {
    run_mainloop<KS, TR>(p, tmem_base, &mapA, &mapB, t, 0);
  } else {
    const EpiPos e = epilogue_wait(p, tmem_base, 0);
    const int gr = t.m0 + e.row;
    for (int h16 = 0; h16 < TN / 32; ++h16) {
      double v[16];
      epilogue_load16<KS>(e, e.col0 + 16 * h16, v);
      if (gr < M)
        for (int c = 0; c < 16; ++c) {
          const int gc = t.n0 + e.col0 + 16 * h16 + c;
          if (gc < Nn) out[(size_t)gr * Nn + gc] = v[c];
        }
    }
    if (raw) {
      for (int q = 0; q < KS; ++q)
        for (int h16 = 0; h16 < TN / 32; ++h16) {
          uint32_t r[16];
          tmem_ld16(e.lane_base + (uint32_t)(q * TN + e.col0 + 16 * h16), r);  // warp-collective
          if (gr < M)
            for (int c = 0; c < 16; ++c) {
              const int gc = t.n0 + e.col0 + 16 * h16 + c;
              if (gc < Nn) raw[((size_t)q * M + gr) * Nn + gc] = (int)r[c];
            }
        }
    }
  }
  pipe_teardown<KS>(p, tmem_base);
}

template <int KS>
static int run_case(int M, int Nn, int C, int c_begin, int c_end, bool verify, int reps) {
  const long long pitch = (C + 15) / 16 * 16;
  std::vector<int8_t> hA((size_t)KS * M * pitch, 0), hB((size_t)KS * Nn * pitch, 0);
  std::mt19937 rng(1234 + M + Nn + C);
  std::uniform_int_distribution<int> dist(-128, 127);
  for (int s = 0; s < KS; ++s) {
    for (int r = 0; r < M; ++r)
      for (int c = 0; c < C; ++c) hA[((size_t)s * M + r) * pitch + c] = (int8_t)dist(rng);
    for (int r = 0; r < Nn; ++r)
      for (int c = 0; c < C; ++c) hB[((size_t)s * Nn + r) * pitch + c] = (int8_t)dist(rng);
  }
  int8_t *dA, *dB;
  int* draw = nullptr;
  double* dout;
  CHECK(cudaMalloc(&dA, hA.size()));
  CHECK(cudaMalloc(&dB, hB.size()));
  CHECK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  if (verify) CHECK(cudaMalloc(&draw, (size_t)KS * M * Nn * sizeof(int)));
  CHECK(cudaMalloc(&dout, (size_t)M * Nn * sizeof(double)));
  CUtensorMap mA = make_map(dA, C, M, pitch, KS, TM);
  CUtensorMap mB = make_map(dB, C, Nn, pitch, KS, TN);
  CHECK(cudaFuncSetAttribute(k_probe<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<KS>::SMEM));
  dim3 grid((Nn + TN - 1) / TN, (M + TM - 1) / TM);
  k_probe<KS><<<grid, THREADS, Geo<KS>::SMEM>>>(mA, mB, c_begin, c_end, M, Nn, draw, dout);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  int bad = 0;
  if (verify) {
    std::vector<int> raw((size_t)KS * M * Nn);
    std::vector<double> out((size_t)M * Nn);
    CHECK(cudaMemcpy(raw.data(), draw, raw.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(out.data(), dout, out.size() * sizeof(double), cudaMemcpyDeviceToHost));
    const int ce = c_end < C ? c_end : C;
    for (int r = 0; r < M; ++r)
      for (int n = 0; n < Nn; ++n) {
        double want = 0;
        for (int q = 0; q < KS; ++q) {
          long long acc = 0;
          for (int i = 0; i <= q; ++i) {
            const int8_t* a = &hA[((size_t)i * M + r) * pitch];
            const int8_t* b = &hB[((size_t)(q - i) * Nn + n) * pitch];
            for (int c = c_begin; c < ce; ++c) acc += (int)a[c] * (int)b[c];
          }
          const int got = raw[((size_t)q * M + r) * Nn + n];
          if (got != (int)acc) {
            if (bad < 10) printf("  mismatch class %d row %d col %d: got %d want %lld\n", q, r, n, got, acc);
            ++bad;
          }
          want += (double)acc * ldexp(1.0, -DIGIT_BITS * q);
        }
        const double got = out[(size_t)r * Nn + n];
        if (fabs(got - want) > 1e-9 * (1 + fabs(want))) {
          if (bad < 10) printf("  combined mismatch row %d col %d: got %.17g want %.17g\n", r, n, got, want);
          ++bad;
        }
      }
    printf("KS=%d M=%d N=%d C=%d [%d,%d): %s (%d mismatches)\n", KS, M, Nn, C, c_begin, c_end, bad ? "FAIL" : "ok", bad);
  }
  if (reps > 0) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) k_probe<KS><<<grid, THREADS, Geo<KS>::SMEM>>>(mA, mB, c_begin, c_end, M, Nn, nullptr, dout);
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k_probe<KS><<<grid, THREADS, Geo<KS>::SMEM>>>(mA, mB, c_begin, c_end, M, Nn, nullptr, dout);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double depth = (double)(c_end - c_begin);
    const double prods = KS * (KS + 1) / 2.0;
    const double tiles = (double)grid.x * grid.y;
    const double int8_ops = 2.0 * tiles * TM * TN * depth * prods;
    const double f64_eq = 2.0 * (double)M * Nn * depth;
    printf("KS=%d M=%d N=%d depth=%.0f: %.3f ms  int8 %.1f TOP/s  fp64-equivalent %.1f TFLOP/s  (%d CTAs)\n", KS, M, Nn,
           depth, ms, int8_ops / ms * 1e-9, f64_eq / ms * 1e-9, (int)tiles);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dout);
  if (draw) cudaFree(draw);
  return bad;
}

// MN-major operands: A [KS][depth][M], B [KS][depth][Nn]; C[m][n] = sum_r A[r][m] B[r][n]
template <int KS>
static int run_case_mn(int M, int Nn, int depth, int c_begin, int c_end, bool verify, int reps) {
  const long long pa = (M + 15) / 16 * 16, pb = (Nn + 15) / 16 * 16;
  std::vector<int8_t> hA((size_t)KS * depth * pa, 0), hB((size_t)KS * depth * pb, 0);
  std::mt19937 rng(4321 + M + Nn + depth);
  std::uniform_int_distribution<int> dist(-128, 127);
  for (int s = 0; s < KS; ++s)
    for (int r = 0; r < depth; ++r) {
      for (int c = 0; c < M; ++c) hA[((size_t)s * depth + r) * pa + c] = (int8_t)dist(rng);
      for (int c = 0; c < Nn; ++c) hB[((size_t)s * depth + r) * pb + c] = (int8_t)dist(rng);
    }
  int8_t *dA, *dB;
  int* draw = nullptr;
  double* dout;
  CHECK(cudaMalloc(&dA, hA.size()));
  CHECK(cudaMalloc(&dB, hB.size()));
  CHECK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  if (verify) CHECK(cudaMalloc(&draw, (size_t)KS * M * Nn * sizeof(int)));
  CHECK(cudaMalloc(&dout, (size_t)M * Nn * sizeof(double)));
  CUtensorMap mA = make_map_mn(dA, M, depth, pa, KS, TM);
  CUtensorMap mB = make_map_mn(dB, Nn, depth, pb, KS, TN);
  CHECK(cudaFuncSetAttribute(k_probe<KS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<KS>::SMEM));
  dim3 grid((Nn + TN - 1) / TN, (M + TM - 1) / TM);
  k_probe<KS, true><<<grid, THREADS, Geo<KS>::SMEM>>>(mA, mB, c_begin, c_end, M, Nn, draw, dout);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  int bad = 0;
  if (verify) {
    std::vector<int> raw((size_t)KS * M * Nn);
    CHECK(cudaMemcpy(raw.data(), draw, raw.size() * sizeof(int), cudaMemcpyDeviceToHost));
    const int ce = c_end < depth ? c_end : depth;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < Nn; ++n)
        for (int q = 0; q < KS; ++q) {
          long long acc = 0;
          for (int i = 0; i <= q; ++i)
            for (int r = c_begin; r < ce; ++r)
              acc += (int)hA[((size_t)i * depth + r) * pa + m] * (int)hB[((size_t)(q - i) * depth + r) * pb + n];
          const int got = raw[((size_t)q * M + m) * Nn + n];
          if (got != (int)acc) {
            if (bad < 10) printf("  MN mismatch class %d row %d col %d: got %d want %lld\n", q, m, n, got, acc);
            ++bad;
          }
        }
    printf("MN-major KS=%d M=%d N=%d depth=%d [%d,%d): %s (%d mismatches)\n", KS, M, Nn, depth, c_begin, c_end,
           bad ? "FAIL" : "ok", bad);
  }
  if (reps > 0) {
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) k_probe<KS, true><<<grid, THREADS, Geo<KS>::SMEM>>>(mA, mB, c_begin, c_end, M, Nn, nullptr, dout);
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k_probe<KS, true><<<grid, THREADS, Geo<KS>::SMEM>>>(mA, mB, c_begin, c_end, M, Nn, nullptr, dout);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double prods = KS * (KS + 1) / 2.0;
    const double tiles = (double)grid.x * grid.y;
    printf("MN-major KS=%d M=%d N=%d depth=%d: %.3f ms  int8 %.1f TOP/s  (%d CTAs)\n", KS, M, Nn, c_end - c_begin, ms,
           2.0 * tiles * TM * TN * (c_end - c_begin) * prods / ms * 1e-9, (int)tiles);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dout);
  if (draw) cudaFree(draw);
  return bad;
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IOLBF, 0);
#if UMMA_TN == 64
  if (argc > 1 && !strcmp(argv[1], "sustain")) {
    // seconds-long back-to-back launches: the rate under the power cap (compare with "perf" = burst)
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 20);
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 3000);
    run_case<6>(8192, 2080, 4160, 0, 4160, false, 2000);
    return 0;
  }
#endif
  if (argc > 1 && !strcmp(argv[1], "tn")) {
    // tile-width experiment (build with -DUMMA_TN=96; at most 5 digits): exactness, burst, sustained
    int bad = 0;
    bad += run_case<2>(128, 96, 64, 0, 64, true, 0);
    bad += run_case<4>(256, 100, 512, 0, 512, true, 0);
    bad += run_case<5>(300, 200, 520, 64, 520, true, 0);
    bad += run_case<5>(256, 192, 1000, 128, 960, true, 0);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 5);
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 20);
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 2500);
    return bad ? 1 : 0;
  }
  if (argc > 1 && !strcmp(argv[1], "ts")) {
    // A-in-tensor-memory experiment (build with -DUMMA_A_TMEM=1; at most 5 digits): exactness, burst, sustained
    int bad = 0;
    bad += run_case<2>(128, 64, 64, 0, 64, true, 0);
    bad += run_case<4>(256, 64, 512, 0, 512, true, 0);
    bad += run_case<5>(300, 200, 520, 64, 520, true, 0);
    bad += run_case<5>(256, 128, 1000, 128, 960, true, 0);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 5);
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 20);
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 2500);
    run_case<4>(8192, 2080, 4160, 0, 4160, false, 20);
    return bad ? 1 : 0;
  }
#if UMMA_TN != 64
  printf("built with UMMA_TN=%d: only the \"tn\" mode is available\n", UMMA_TN);
  return 2;
#else
  const bool perf = argc > 1 && !strcmp(argv[1], "perf");
  int bad = 0;
  if (!perf) {
    bad += run_case<2>(128, 64, 64, 0, 64, true, 0);
    bad += run_case<6>(128, 64, 64, 0, 64, true, 0);
    bad += run_case<6>(128, 64, 200, 0, 200, true, 0);
    bad += run_case<6>(256, 128, 1000, 128, 960, true, 0);
    bad += run_case<5>(300, 200, 520, 64, 520, true, 0);
    bad += run_case<7>(128, 128, 333, 0, 333, true, 0);
    bad += run_case<8>(128, 64, 256, 0, 256, true, 0);
    bad += run_case<4>(256, 64, 512, 0, 512, true, 0);
    bad += run_case_mn<2>(128, 64, 64, 0, 64, true, 0);
    bad += run_case_mn<5>(128, 64, 64, 0, 64, true, 0);
    bad += run_case_mn<5>(128, 64, 200, 0, 200, true, 0);
    bad += run_case_mn<5>(256, 192, 1000, 128, 960, true, 0);
    bad += run_case_mn<6>(300, 200, 520, 64, 520, true, 0);
    bad += run_case_mn<8>(128, 128, 256, 0, 256, true, 0);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  } else {
    // chip-level contention check: 33, 66, 132 CTAs (all resident at once) vs many waves
    run_case<6>(128, 2080, 4160, 0, 4160, false, 5);
    run_case<6>(256, 2080, 4160, 0, 4160, false, 5);
    run_case<6>(512, 2080, 4160, 0, 4160, false, 5);
    run_case<6>(8192, 2080, 4160, 0, 4160, false, 5);
    run_case<5>(8192, 2080, 4160, 0, 4160, false, 5);
    run_case<7>(8192, 2080, 4160, 0, 4160, false, 5);
    run_case<6>(16384, 4160, 4160, 0, 4160, false, 3);
    run_case_mn<5>(2080, 2080, 8192, 0, 8192, false, 5);   // one chunk's full (unsymmetrised) Gram
  }
  return bad ? 1 : 0;
#endif
}
