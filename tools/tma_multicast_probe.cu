// Load-only probe of the digit-product pipeline's operand stream (no MMA): how long do the TMA loads of one
// backward launch (8192 x 2048 outputs, depth 4160, 5 digits: 60 KB per 64-byte block and CTA) take when the A tile
// (128 rows x 5 digits, shared by the CTAs of one row block) is
//   cluster size 1 : loaded by every CTA on its own (the product path today), or
//   cluster size c : loaded once per cluster -- each CTA fetches 1/c of the rows and multicasts it to all c CTAs.
// Every wait is bounded by a clock budget, so a protocol error ends as "TIMED OUT", not as a hang.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/tma_multicast_probe tools/tma_multicast_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cooperative_groups.h>

#include "../scfgp_b200/csrc/umma_i8.cuh"

namespace cg = cooperative_groups;
using namespace scfgp;
using namespace scfgp::umma;

constexpr int KS = 5;
constexpr int STAGES = 3;
constexpr int A_BYTES = KS * TM * KB, B_BYTES = KS * TN * KB, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr long long BUDGET = 100000000ll;    // about 50 ms

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(const int8_t* base, int C, int rows, long long pitch, int box_rows) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)KS};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
  cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)box_rows, (cuuint32_t)KS};
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

__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}

__device__ __forceinline__ void arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(ra) : "memory");
}

// grid = (column tiles, row tiles), cluster = (csz, 1, 1): the csz CTAs of a cluster share the row tile
__global__ void k_loads(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAslice,
                        const __grid_constant__ CUtensorMap mapB, int depth, int csz, int* fail) {
  extern __shared__ uint8_t raw[];
  const uint32_t off = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
  uint8_t* stages = raw + off;
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = csz > 1 ? cg::this_cluster().block_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], csz); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (csz > 1) cg::this_cluster().sync();
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int nkb = depth / KB;
  const int slice_rows = TM / csz;
  bool ok = true;
  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int st = kb % STAGES;
      const uint32_t use = kb / STAGES;
      ok = wait_bounded(&empty[st], (use & 1) ^ 1);
      if (!ok) break;
      uint8_t* sa = stages + (size_t)st * STAGE_BYTES;
      mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
      if (csz == 1) {
        tma_load_3d(sa, &mapA, kb * KB, m0, 0, &full[st]);
      } else {
        // rows [rank * slice_rows, +slice_rows) of the A tile, all digits, to every CTA of the cluster
        tma_load_3d_mc(sa + (size_t)rank * slice_rows * KB * KS, &mapAslice, kb * KB, m0 + (int)rank * slice_rows, 0, &full[st],
                       (uint16_t)((1u << csz) - 1u));
      }
      tma_load_3d(sa + A_BYTES, &mapB, kb * KB, n0, 0, &full[st]);
    }
  } else if (warp == 1 && lane == 0) {
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int st = kb % STAGES;
      const uint32_t use = kb / STAGES;
      ok = wait_bounded(&full[st], use & 1);
      if (!ok) break;
      if (csz == 1) mbar_arrive(&empty[st]);
      else
        for (int r = 0; r < csz; ++r) arrive_remote(&empty[st], (uint32_t)r);
    }
  }
  if (!ok) atomicAdd(fail, 1);
  __syncthreads();
  if (csz > 1) cg::this_cluster().sync();    // nobody leaves while peers may still write into its shared memory
}

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  const int M = 8192, Nn = 2048, C = 4160;
  const long long pitch = C;
  int8_t *dA, *dB;
  int* dfail;
  cudaMalloc(&dA, (size_t)KS * M * pitch);
  cudaMalloc(&dB, (size_t)KS * Nn * pitch);
  cudaMalloc(&dfail, 4);
  cudaMemset(dA, 1, (size_t)KS * M * pitch);
  cudaMemset(dB, 2, (size_t)KS * Nn * pitch);
  const int smem = STAGES * STAGE_BYTES + 1024 + 256;
  cudaFuncSetAttribute(k_loads, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_loads, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  CUtensorMap mA = make_map(dA, C, M, pitch, TM);
  CUtensorMap mB = make_map(dB, C, Nn, pitch, TN);
  for (int csz : {1, 2, 4, 8}) {
    CUtensorMap mS = make_map(dA, C, M, pitch, TM / csz);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(Nn / TN, M / TM);
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaMemset(dfail, 0, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = cudaSuccess;
    const int reps = 5;
    int hf = 0;
    for (int r = 0; r < 2 + reps && e == cudaSuccess && hf == 0; ++r) {
      if (r == 2) cudaEventRecord(e0);
      e = cudaLaunchKernelEx(&cfg, k_loads, mA, mS, mB, C, csz, dfail);
      if (r < 2) {   // a broken protocol shows up in the warm-up launches: stop before timing
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaMemcpy(&hf, dfail, 4, cudaMemcpyDeviceToHost);
      }
    }
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    cudaMemcpy(&hf, dfail, 4, cudaMemcpyDeviceToHost);
    const double ctas = (double)(Nn / TN) * (M / TM);
    const double delivered = ctas * (C / KB) * STAGE_BYTES;
    const double fetched = ctas * (C / KB) * ((double)A_BYTES / csz + B_BYTES);
    printf("cluster %d: %s, %d timed-out roles, %.3f ms per launch; delivered to shared memory %.1f TB/s, fetched from L2 %.1f TB/s\n",
           csz, cudaGetErrorString(e), hf, ms, delivered / ms * 1e-9, fetched / ms * 1e-9);
    if (e != cudaSuccess) break;
  }
  return 0;
}
