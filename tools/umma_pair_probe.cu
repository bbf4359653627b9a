// Instruction timing of tcgen05.mma.kind::i8 issued for ONE CTA (cta_group::1, M = 128) and for a CTA PAIR
// (cta_group::2, M = 256: each CTA holds its 128 rows of A and HALF of the B rows): clocks per instruction as a
// function of N, operands = whatever is in shared memory (timing only, results are not read).
// Every wait is bounded by a clock budget, so the kernel always ends.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_pair_probe tools/umma_pair_probe.cu
#include <cstdio>
#include <cstdint>
#include <cooperative_groups.h>

#include "../scfgp_b200/csrc/umma_i8.cuh"

namespace cg = cooperative_groups;
using namespace scfgp;
using namespace scfgp::umma;

__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity, long long budget) {
  const long long t0 = clock64();
  while (clock64() - t0 < budget) {
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

template <int PAIR>
__global__ void k_time(int n_cols, int per_batch, int batches, int alternate, long long* out /*[2]: clocks, ok*/) {
  extern __shared__ uint8_t raw[];
  const uint32_t off = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
  uint8_t* base = raw + off;                 // 1024-aligned: A at 0 (128 x 64 B), B at 16 KB (256 x 64 B)
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 48 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(base + 48 * 1024 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned rank = 0;
  if (PAIR) rank = cg::this_cluster().block_rank();
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    } else {
      tmem_alloc(slot, 512);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cg::this_cluster().sync();
  tc_fence_after();
  const uint32_t tmem = *slot;
  long long clocks = 0;
  bool ok = true;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = PAIR ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((256u >> 4) << 24))
                                : idesc_i8(n_cols);
    // alternate & 1: successive instructions take A from two different tiles; alternate & 2: B as well
    const uint64_t adesc0 = smem_desc_sw64(smem_u32(base));
    const uint64_t adesc1 = smem_desc_sw64(smem_u32(base) + ((alternate & 1) ? 8 * 1024 : 0));
    const uint64_t bdesc0 = smem_desc_sw64(smem_u32(base) + 16 * 1024);
    const uint64_t bdesc1 = smem_desc_sw64(smem_u32(base) + ((alternate & 2) ? 32 * 1024 : 16 * 1024));
    const long long t0 = clock64();
    for (int b = 0; b < batches && ok; ++b) {
      for (int i = 0; i < per_batch; ++i) {
        const uint64_t adesc = (i & 1) ? adesc1 : adesc0, bdesc = (i & 1) ? bdesc1 : bdesc0;
        if (PAIR)
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
              "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u)
              : "memory");
        else
          mma_i8(tmem, adesc, bdesc, idesc, 1u);
      }
      if (PAIR)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
      else
        tc_commit(bar);
      ok = wait_bounded(bar, (uint32_t)b & 1u, 400000000ll);
    }
    clocks = clock64() - t0;
    out[0] = clocks;
    out[1] = ok ? 1 : 0;
  } else if (PAIR && rank == 1 && warp == 1 && lane == 0) {
    // the peer's barrier receives the multicast arrivals as well: consume them so that phases stay aligned
    for (int b = 0; b < batches; ++b)
      if (!wait_bounded(bar, (uint32_t)b & 1u, 400000000ll)) break;
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cg::this_cluster().sync();
  if (warp == 1) {
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
    else
      tmem_dealloc(tmem, 512);
  }
}

// The product path's instruction stream for 5 digits, without TMA traffic: per 32-byte contraction step six
// instructions (A digit i x B digits j0.., N = 256, 64, 256, 192, 128, 64) into the class columns [64 (i + j0), ...).
// variant 0: as in run_mainloop;  1: all six into the same columns [0, N);  2: same A and B for all six
__global__ void k_mix(int steps, int variant, int commit_every, int random_data, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t off = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
  uint8_t* base = raw + off;                 // A digits at 0 (5 x 8 KB), B digits at 40 KB (5 x 4 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 64 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(base + 64 * 1024 + 64);
  uint64_t* bar2 = reinterpret_cast<uint64_t*>(base + 64 * 1024 + 128);   // extra commits land here (never waited on)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 60 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0x01010101u * (i & 3);
    if (random_data) {                         // full-range int8 digits, as the product's operands
      v = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
    }
    reinterpret_cast<uint32_t*>(base)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1 && lane == 0) {
    const uint32_t sa = smem_u32(base), sb = sa + 40 * 1024;
    constexpr int KS = 5;
    const long long t0 = clock64();
    bool ok = true;
    int done = 0;
    for (int b = 0; done < steps && ok; ++b) {
      for (int kb = 0; kb < 16 && done < steps; ++kb) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks, ++done) {
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            const uint64_t adesc = smem_desc_sw64(sa + (variant == 2 ? 0 : i) * (TM * KB) + ks * 32);
#pragma unroll
            for (int j0 = 0; j0 < KS - i; j0 += 4) {
              const int cnt = (KS - i - j0) < 4 ? (KS - i - j0) : 4;
              const uint64_t bdesc = smem_desc_sw64(sb + (variant == 2 ? 0 : j0) * (TN * KB) + ks * 32);
              mma_i8(tmem + (uint32_t)(variant == 1 ? 0 : TN * (i + j0)), adesc, bdesc, idesc_i8(TN * cnt), 1u);
            }
          }
        }
        // commit_every > 0: a tcgen05.commit after every commit_every-th 64-byte block, as the pipeline does
        // when it hands a stage back to the TMA producer
        if (commit_every > 0 && (kb % commit_every) == commit_every - 1) tc_commit(bar2);
      }
      tc_commit(bar);
      ok = wait_bounded(bar, (uint32_t)b & 1u, 400000000ll);
    }
    out[0] = clock64() - t0;
    out[1] = ok ? 1 : 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

static void run_mix(int variant, int commit_every, int ctas = 1, int steps = 4096, int random_data = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int smem = 66 * 1024 + 1024;
  cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k_mix<<<ctas, 128, smem>>>(steps, variant, commit_every, random_data, d);   // every CTA runs the same stream; the last writer reports
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 1 && ctas > 1)
      printf("  %s operands, %d CTAs (one per SM) x %d steps: %.3f ms wall, %.1f clk per step -> SM clock %.0f MHz, %.0f int8 TOP/s chip-wide\n", random_data ? "random int8" : "low-toggle", ctas, steps,
             ms, (double)h[0] / steps, (double)h[0] / (ms * 1e3), 2.0 * 128 * 960 * 32 * (double)steps * ctas / (ms * 1e9));
    if (rep == 1 || e != cudaSuccess)
      printf("5-digit mix (6 instructions, 960 columns per step), commit every %d blocks, variant %d (%s): %s, %7.1f clk per step (%s); floor 480\n",
             commit_every, variant, variant == 0 ? "product path" : variant == 1 ? "all into columns [0, N)" : "same A and B", cudaGetErrorString(e),
             (double)h[0] / steps, h[1] ? "completed" : "TIMED OUT");
    if (e != cudaSuccess) break;
  }
  cudaFree(d);
}

template <int PAIR>
static void run(int n_cols, int alternate) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int smem = 50 * 1024 + 1024;
  cudaFuncSetAttribute(k_time<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int per_batch = 64, batches = 32;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PAIR ? 2 : 1);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_time<PAIR>, n_cols, per_batch, batches, alternate, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (rep == 1 || e != cudaSuccess)
      printf("%s N=%3d %s: %s, %8.1f clk per instruction (%s)  -> %.1f%% of the 0.5 clk/column floor per SM\n",
             PAIR ? "cta_group::2 M=256" : "cta_group::1 M=128", n_cols,
             alternate == 0 ? "same A, same B    " : alternate == 1 ? "A alternates      " : alternate == 2 ? "B alternates      " : "A and B alternate ",
             cudaGetErrorString(e),
             (double)h[0] / (per_batch * batches), h[1] ? "completed" : "TIMED OUT",
             100.0 * 0.5 * n_cols / ((double)h[0] / (per_batch * batches)));
    if (e != cudaSuccess) break;
  }
  cudaFree(d);
}

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  run_mix(0, 1);
  run_mix(0, 1, 148, 8192);
  run_mix(0, 1, 148, 4000000);
  run_mix(0, 1, 148, 8192, 1);
  run_mix(0, 1, 148, 400000, 1);
  run_mix(0, 1, 148, 4000000, 1);
  for (int alt : {0, 3}) {
    for (int n : {64, 128, 192, 256}) run<0>(n, alt);
    for (int n : {64, 128, 192, 256}) run<1>(n, alt);
  }
  return 0;
}
