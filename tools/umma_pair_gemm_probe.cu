// CTA-pair version of the digit-product mainloop ("what bounds it") as a stand-alone probe:
// tcgen05.mma.cta_group::2.kind::i8 (M = 256 over the two CTAs of a cluster), 96-column tiles, all classes resident in
// tensor memory, B operand SPLIT over the pair:
//   * an instruction over two B digits (N = 192): CTA 0 holds digit j (96 rows), CTA 1 digit j + 1, at the same offset;
//   * an instruction over one B digit (N = 96): CTA 0 holds its rows 0..47, CTA 1 rows 48..95, at the same offset;
// so each SM stores 336 instead of 480 B rows per stage (5 digits) and the MMAs read half the B bytes per SM.
// Every wait is bounded by a clock budget: a protocol error ends as "TIMED OUT", not as a hang.
// Checks every class accumulator against a CPU integer reference, then times the pass-3 shape.
//
// RESULT (round 1, profiles/r1_umma_stream_and_load_ceilings.txt): every class accumulator bit-exact on the first run -- B rows
// [0, N/2) of a cta_group::2 instruction do come from the leader CTA, the leader's barrier is the own address with bit 24
// cleared, one arrive.expect_tx of 2 x STAGE_BYTES by the leader covers both CTAs' loads -- but only 1 % faster than the
// 1-CTA 96-wide mainloop (0.708 / 0.822 ms against 0.717 / 0.827 ms): operand traffic is not what limits the kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/bin/umma_pair_gemm_probe tools/umma_pair_gemm_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include <cuda.h>
#include <cooperative_groups.h>

#include "../scfgp_b200/csrc/umma_i8.cuh"

namespace cg = cooperative_groups;
using namespace scfgp;
using namespace scfgp::umma;

constexpr int TNW = 96;
constexpr long long BUDGET = 200000000ll;   // about 0.1 s per wait

template <int KS>
struct PairGeo {
  static constexpr int NPAIR = KS / 2;           // B digit pairs (2m, 2m + 1): 96 rows per CTA each
  static constexpr int NSINGLE = (KS + 1) / 2;   // single B digits j = 0, 2, 4, ...: 48 rows per CTA each
  static constexpr int PAIR_BYTES = TNW * KB, SINGLE_BYTES = (TNW / 2) * KB;
  static constexpr int A_BYTES = KS * TM * KB;
  static constexpr int B_BYTES = NPAIR * PAIR_BYTES + NSINGLE * SINGLE_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(KS * TNW <= 512 && STAGE_BYTES % 1024 == 0, "geometry");
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// digits [KS][rows][pitch] int8, contraction contiguous; box = 64 bytes x box_rows x box_digits
static CUtensorMap make_map(const int8_t* base, int C, int rows, long long pitch, int KS, int box_rows, int box_digits) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)KS};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
  cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)box_rows, (cuuint32_t)box_digits};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
  return m;
}

__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (clock64() - t0 < BUDGET) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
// TMA load of one CTA of the pair; the bytes complete on the LEADER's barrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t idesc_i8_pair(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

// grid = 2 x (pairs of row tiles x column tiles), cluster = (2, 1, 1); one 256 x 96 tile per cluster
template <int KS>
__global__ void __maxnreg__(MAX_REGS)
k_pair(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB96,
       const __grid_constant__ CUtensorMap mapB48, int c_begin, int c_end, int M, int Nn, int ntn,
       int* raw /*[KS][M][Nn] or null*/, double* out /*[M][Nn]*/, int* fail) {
  using G = PairGeo<KS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t off = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* stages = smem_raw + off;
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + (size_t)G::STAGES * G::STAGE_BYTES);
  uint64_t* empty = full + G::STAGES;
  uint64_t* acc_full = empty + G::STAGES;
  uint32_t* slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cg::this_cluster().block_rank();
  const int pid = blockIdx.x >> 1;
  const int m0 = (pid / ntn) * 2 * TM + (int)rank * TM;   // this CTA's 128 rows of the pair's 256
  const int n0 = (pid % ntn) * TNW;
  if (threadIdx.x == 0) {
    for (int s = 0; s < G::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cg::this_cluster().sync();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  const int nkb = (c_end - c_begin + KB - 1) / KB;
  bool ok = true;
  if (warp == 0 && lane == 0) {
    // ---- TMA producer (both CTAs): own A rows, own half of every B sub-block
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int st = kb % G::STAGES;
      const uint32_t use = kb / G::STAGES;
      ok = wait_bounded(&empty[st], (use & 1) ^ 1);
      if (!ok) break;
      uint8_t* sa = stages + (size_t)st * G::STAGE_BYTES;
      uint8_t* sb = sa + G::A_BYTES;
      if (rank == 0) mbar_arrive_expect_tx(&full[st], 2 * G::STAGE_BYTES);   // the bytes of both CTAs
      const int c = c_begin + kb * KB;
      tma_load_3d_pair(sa, &mapA, c, m0, 0, &full[st]);
#pragma unroll
      for (int m = 0; m < G::NPAIR; ++m)
        tma_load_3d_pair(sb + m * G::PAIR_BYTES, &mapB96, c, n0, 2 * m + (int)rank, &full[st]);
#pragma unroll
      for (int s = 0; s < G::NSINGLE; ++s)
        tma_load_3d_pair(sb + G::NPAIR * G::PAIR_BYTES + s * G::SINGLE_BYTES, &mapB48, c, n0 + (TNW / 2) * (int)rank, 2 * s,
                         &full[st]);
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ---- MMA issuer (leader only): M = 256 over the pair
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int st = kb % G::STAGES;
      const uint32_t use = kb / G::STAGES;
      ok = wait_bounded(&full[st], use & 1);
      if (!ok) break;
      tc_fence_after();
      const uint32_t sa = smem_u32(stages + (size_t)st * G::STAGE_BYTES);
      const uint32_t sb = sa + G::A_BYTES;
#pragma unroll
      for (int ks = 0; ks < KB / 32; ++ks) {
#pragma unroll
        for (int i = 0; i < KS; ++i) {
          const uint64_t adesc = smem_desc_sw64(sa + i * (TM * KB) + ks * 32);
          const int nd = KS - i;                                   // B digits 0 .. nd - 1
          const uint32_t acc = (kb == 0 && ks == 0 && i == 0) ? 0u : 1u;
#pragma unroll
          for (int m = 0; m < nd / 2; ++m)
            mma_i8_pair(tmem_base + (uint32_t)(TNW * (i + 2 * m)), adesc, smem_desc_sw64(sb + m * G::PAIR_BYTES + ks * 32),
                        idesc_i8_pair(2 * TNW), acc);
          if (nd & 1) {
            const int j = nd - 1;                                  // even
            mma_i8_pair(tmem_base + (uint32_t)(TNW * (i + j)), adesc,
                        smem_desc_sw64(sb + G::NPAIR * G::PAIR_BYTES + (j / 2) * G::SINGLE_BYTES + ks * 32), idesc_i8_pair(TNW), acc);
          }
        }
      }
      tc_commit_pair(&empty[st]);
    }
    tc_commit_pair(acc_full);
  } else if (warp >= EPI_WARP0) {
    // ---- epilogue (both CTAs): own 128 rows
    ok = wait_bounded(acc_full, 0);
    tc_fence_after();
    if (ok) {
      const int quarter = warp & 3, half = (warp - EPI_WARP0) >> 2;
      EpiPos e;
      e.row = quarter * 32 + lane;
      e.col0 = half * (TNW / 2);
      e.lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const int gr = m0 + e.row;
      for (int h16 = 0; h16 < TNW / 32; ++h16) {
        double v[16];
        epilogue_load16<KS, TNW>(e, e.col0 + 16 * h16, v);
        if (gr < M)
          for (int c = 0; c < 16; ++c) {
            const int gc = n0 + e.col0 + 16 * h16 + c;
            if (gc < Nn) out[(size_t)gr * Nn + gc] = v[c];
          }
      }
      if (raw) {
        for (int q = 0; q < KS; ++q)
          for (int h16 = 0; h16 < TNW / 32; ++h16) {
            uint32_t r[16];
            tmem_ld16(e.lane_base + (uint32_t)(q * TNW + e.col0 + 16 * h16), r);
            if (gr < M)
              for (int c = 0; c < 16; ++c) {
                const int gc = n0 + e.col0 + 16 * h16 + c;
                if (gc < Nn) raw[((size_t)q * M + gr) * Nn + gc] = (int)r[c];
              }
          }
      }
    }
  }
  if (!ok) atomicAdd(fail, 1);
  tc_fence_before();
  __syncthreads();
  cg::this_cluster().sync();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int KS>
static int run_case(int M, int Nn, int C, int c_begin, int c_end, bool verify, int reps) {
  using G = PairGeo<KS>;
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
  int *draw = nullptr, *dfail;
  double* dout;
  cudaMalloc(&dA, hA.size());
  cudaMalloc(&dB, hB.size());
  cudaMalloc(&dfail, 4);
  cudaMemset(dfail, 0, 4);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
  if (verify) cudaMalloc(&draw, (size_t)KS * M * Nn * sizeof(int));
  cudaMalloc(&dout, (size_t)M * Nn * sizeof(double));
  CUtensorMap mA = make_map(dA, C, M, pitch, KS, TM, KS);
  CUtensorMap mB96 = make_map(dB, C, Nn, pitch, KS, TNW, 1);
  CUtensorMap mB48 = make_map(dB, C, Nn, pitch, KS, TNW / 2, 1);
  cudaFuncSetAttribute(k_pair<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
  const int ntn = (Nn + TNW - 1) / TNW, ntm2 = (M + 2 * TM - 1) / (2 * TM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ntn * ntm2);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = G::SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_pair<KS>, mA, mB96, mB48, c_begin, c_end, M, Nn, ntn, draw, dout, dfail);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  int hf = 0;
  cudaMemcpy(&hf, dfail, 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess || hf) {
    printf("KS=%d M=%d N=%d: %s, %d timed-out roles\n", KS, M, Nn, cudaGetErrorString(e), hf);
    return 1;
  }
  int bad = 0;
  if (verify) {
    std::vector<int> raw((size_t)KS * M * Nn);
    std::vector<double> out((size_t)M * Nn);
    cudaMemcpy(raw.data(), draw, raw.size() * sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dout, out.size() * sizeof(double), cudaMemcpyDeviceToHost);
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
    printf("pair KS=%d M=%d N=%d C=%d [%d,%d): %s (%d mismatches)\n", KS, M, Nn, C, c_begin, c_end, bad ? "FAIL" : "ok", bad);
  }
  if (reps > 0 && !bad) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) cudaLaunchKernelEx(&cfg, k_pair<KS>, mA, mB96, mB48, c_begin, c_end, M, Nn, ntn, (int*)nullptr, dout, dfail);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&cfg, k_pair<KS>, mA, mB96, mB48, c_begin, c_end, M, Nn, ntn, (int*)nullptr, dout, dfail);
    cudaEventRecord(e1);
    e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    cudaMemcpy(&hf, dfail, 4, cudaMemcpyDeviceToHost);
    const double depth = (double)(c_end - c_begin), prods = KS * (KS + 1) / 2.0;
    const double tiles = (double)ntn * ntm2 * 2;
    printf("pair KS=%d M=%d N=%d depth=%.0f: %s, %d timed-out roles, %.3f ms  int8 %.1f TOP/s  fp64-equivalent %.1f TFLOP/s  (%d CTAs)\n", KS, M,
           Nn, depth, cudaGetErrorString(e), hf, ms, 2.0 * tiles * TM * TNW * depth * prods / ms * 1e-9,
           2.0 * (double)M * Nn * depth / ms * 1e-9, (int)tiles);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dout); cudaFree(dfail);
  if (draw) cudaFree(draw);
  return bad;
}

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  int bad = 0;
  bad += run_case<2>(256, 96, 64, 0, 64, true, 0);
  if (bad) { printf("PAIR PROBE FAILED (first case)\n"); return 1; }
  bad += run_case<4>(256, 100, 512, 0, 512, true, 0);
  bad += run_case<5>(300, 200, 520, 64, 520, true, 0);
  bad += run_case<5>(512, 192, 1000, 128, 960, true, 0);
  printf(bad ? "PAIR PROBE FAILED\n" : "PAIR PROBE OK\n");
  if (bad) return 1;
  run_case<5>(8192, 2080, 4160, 0, 4160, false, 5);     // 1-CTA 96-wide tiles: 0.717 ms; 64-wide: 0.784 ms
  run_case<5>(8192, 2080, 4160, 0, 4160, false, 20);
  run_case<5>(8192, 2080, 4160, 0, 4160, false, 2500);  // sustained: 0.827 / 0.895 ms
  return 0;
}
