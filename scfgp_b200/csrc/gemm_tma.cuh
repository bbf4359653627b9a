// TMA-fed FP64 DMMA mainloop for the three hot kernels (sm_100a).
//
// One producer warp issues `cp.async.bulk.tensor.2d` (TMA) loads of 16x16-double boxes
// (2 KB, 128-byte swizzled) into a multi-stage shared-memory ring and signals an
// mbarrier per stage (expect_tx / complete_tx); eight consumer warps wait on the "full"
// barrier, feed DMMA.8x8x4 from the swizzled tiles and release the stage through an
// "empty" barrier.  There is no CTA-wide barrier in the k-loop, so the consumer warps
// drift apart and the FP64 pipe of every SM sub-partition always has a warp to issue from.
//
// Tiles are 128 (rows) x 16 (k) doubles per operand per k-step = 8 boxes.
//   K-major operand (element (row,k) at ptr[row*ld+k]): box = 16 k (inner) x 16 rows.
//     Tile row r sits in box r/16, 128-byte line r%16; 16-byte chunk c = k/2 is stored at
//     chunk c ^ (r%8).  A DMMA fragment (8 rows x 4 k) reads tile rows base8 + pi(g),
//     pi(g) = 2*(g%4) + g/4, so the 16 lanes of a half warp hit 16 distinct 8-byte banks.
//   MN-major operand (element (row,k) at ptr[k*ld+row]): box = 16 rows (inner) x 16 k.
//     k-row kr sits at line kr; chunk (row%16)/2 is stored at chunk ^ (kr%8).  The four k
//     of k4-step kk map to lines 8*(kk/2) + 2t + (kk%2) (any fixed permutation of the
//     reduction index is fine as long as both operands use it), again conflict-free.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_dmma.cuh"

namespace scfgp {

constexpr int TMA_BOX_BYTES = 2048;              // 16 x 16 doubles
constexpr int TMA_TILE_BYTES = 8 * TMA_BOX_BYTES;  // 128 x 16 doubles
constexpr int TMA_KSUB = 2;                      // k16 sub-steps per pipeline stage
constexpr int TMA_STAGE_BYTES = 2 * TMA_KSUB * TMA_TILE_BYTES;  // A + B
constexpr int TMA_STAGES = 3;
constexpr int TMA_CONSUMER_WARPS = 8;
// 3 warpgroups: two of consumers, one whose first lane is the TMA producer.  Registers are
// allocated per warpgroup, so the producer group hands most of its registers to the
// consumers with setmaxnreg (40 vs 232 per thread; 128*40 + 256*232 <= 65536).
constexpr int TMA_THREADS = (TMA_CONSUMER_WARPS + 4) * 32;
constexpr int TMA_PRODUCER_REGS = 40;
constexpr int TMA_CONSUMER_REGS = 232;
constexpr size_t TMA_SMEM_BYTES = (size_t)TMA_STAGES * TMA_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA: global (tensor map, coords {c0 inner, c1 outer}) -> shared, completion on mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

__device__ __forceinline__ int pi8(int g) { return 2 * (g & 3) + (g >> 2); }

// Which rows of the stored matrices feed the 8 boxes of one operand tile at k offset k0.
// K-major: box b covers tile rows [16b, 16b+16) -> coords {k0, row0(b)}
// MN-major: box b covers tile rows [16b, 16b+16) -> coords {row0(b), k0}
struct BoxRows {
  int row0[8];
};

// Producer side: fill one stage (TMA_KSUB k16 sub-steps of A and B).  Only the first
// `nbox_b` boxes (16 rows each) of the B tile are fetched: narrow tiles (NW < 4) need fewer.
__device__ __forceinline__ uint32_t tma_stage_tx_bytes(int nbox_b) {
  return (uint32_t)(TMA_KSUB * (TMA_TILE_BYTES + nbox_b * TMA_BOX_BYTES));
}
template <bool AK, bool BKM>
__device__ __forceinline__ void tma_issue_stage(uint8_t* stage, const CUtensorMap* mapA, const BoxRows& ra,
                                                const CUtensorMap* mapB, const BoxRows& rb, int nbox_b, int k0,
                                                uint64_t* full) {
  mbar_arrive_expect_tx(full, tma_stage_tx_bytes(nbox_b));
#pragma unroll
  for (int sub = 0; sub < TMA_KSUB; ++sub) {
    uint8_t* a = stage + sub * 2 * TMA_TILE_BYTES;
    uint8_t* b = a + TMA_TILE_BYTES;
    const int k = k0 + 16 * sub;
#pragma unroll
    for (int bx = 0; bx < 8; ++bx) {
      if (AK) tma_load_2d(a + bx * TMA_BOX_BYTES, mapA, k, ra.row0[bx], full);
      else tma_load_2d(a + bx * TMA_BOX_BYTES, mapA, ra.row0[bx], k, full);
    }
#pragma unroll
    for (int bx = 0; bx < 8; ++bx) {
      if (bx < nbox_b) {
        if (BKM) tma_load_2d(b + bx * TMA_BOX_BYTES, mapB, k, rb.row0[bx], full);
        else tma_load_2d(b + bx * TMA_BOX_BYTES, mapB, rb.row0[bx], k, full);
      }
    }
  }
}

// Consumer side: 4 k4-steps over one 128x16 A tile and one (32 NW)x16 B tile.
//   NW = number of 32-column warp columns that hold valid output: the 8 consumer warps
//   form an (8/NW) x NW grid with warp tiles (16 NW) x 32, so a narrow border tile
//   (NW = 1 or 2) keeps all eight warps busy on a quarter / half of the work.
//   acc[i][j][e]: row = wm0 + 8i + (pi8(g) | g), col = wn0 + 8j + (pi8(2t+e) | 2t+e).
template <bool AK, bool BKM, int NW>
__device__ __forceinline__ void tma_compute_k16(double (&acc)[2 * NW][4][2], const uint8_t* At, const uint8_t* Bt,
                                                int wm0, int wn0, int g, int t) {
  constexpr int MI = 2 * NW;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    double a[MI], b[4];
    if (AK) {
      const int x = ((((2 * kk + (t >> 1)) ^ pi8(g)) << 4) + ((t & 1) << 3));
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = wm0 + 8 * i;  // multiple of 8
        a[i] = *reinterpret_cast<const double*>(At + (r >> 4) * TMA_BOX_BYTES + (((r & 15) + pi8(g)) << 7) + x);
      }
    } else {
      const int kr = 8 * (kk >> 1) + 2 * t + (kk & 1);
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int m = wm0 + 8 * i + g;
        a[i] = *reinterpret_cast<const double*>(At + (m >> 4) * TMA_BOX_BYTES + (kr << 7) +
                                                 ((((m & 15) >> 1) ^ (kr & 7)) << 4) + ((m & 1) << 3));
      }
    }
    if (BKM) {
      const int x = ((((2 * kk + (t >> 1)) ^ pi8(g)) << 4) + ((t & 1) << 3));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = wn0 + 8 * j;
        b[j] = *reinterpret_cast<const double*>(Bt + (r >> 4) * TMA_BOX_BYTES + (((r & 15) + pi8(g)) << 7) + x);
      }
    } else {
      const int kr = 8 * (kk >> 1) + 2 * t + (kk & 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = wn0 + 8 * j + g;
        b[j] = *reinterpret_cast<const double*>(Bt + (n >> 4) * TMA_BOX_BYTES + (kr << 7) +
                                                 ((((n & 15) >> 1) ^ (kr & 7)) << 4) + ((n & 1) << 3));
      }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// Tile-local row of accumulator row index (i, g) / column index (j, 2t+e)
template <bool KMAJOR>
__device__ __forceinline__ int tma_idx(int w0, int frag, int lane_idx) {
  return KMAJOR ? w0 + 8 * frag + pi8(lane_idx) : w0 + 8 * frag + lane_idx;
}

struct TmaPipe {
  uint8_t* stages;
  uint64_t* full;
  uint64_t* empty;
};

// Carve the dynamic shared memory (1024-byte aligned stages, then the barriers) and
// initialise the barriers.  Must be called by all threads; ends with __syncthreads().
__device__ __forceinline__ TmaPipe tma_pipe_init(uint8_t* smem_raw) {
  TmaPipe p;
  // offset arithmetic (not integer casts) so the compiler keeps the shared address space
  const uint32_t off = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  p.stages = smem_raw + off;
  p.full = reinterpret_cast<uint64_t*>(p.stages + (size_t)TMA_STAGES * TMA_STAGE_BYTES);
  p.empty = p.full + TMA_STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) {
      mbar_init(&p.full[s], 1);
      mbar_init(&p.empty[s], TMA_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  return p;
}

// Role split.  Warps >= TMA_CONSUMER_WARPS form the producer warpgroup: they shrink their
// register allocation, lane 0 of the first one streams every stage of every segment, and
// the caller returns right after.  `nseg` segments (1 or 2) share the accumulators; segment
// s reduces over [0, depth[s]) with operand maps mapA[s], mapB[s].
__device__ __forceinline__ int tma_total_iters(int nseg, const int* depth) {
  constexpr int KSTAGE = 16 * TMA_KSUB;
  int total = 0;
  for (int s = 0; s < nseg; ++s) total += (depth[s] + KSTAGE - 1) / KSTAGE;
  return total;
}

template <bool AK, bool BKM>
__device__ __forceinline__ void tma_producer(const TmaPipe& pipe, int nseg, const CUtensorMap* const* mapA,
                                             const CUtensorMap* const* mapB, const BoxRows& ra, const BoxRows& rb,
                                             int nbox_b, const int* depth) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(TMA_PRODUCER_REGS));
  if ((threadIdx.x >> 5) != TMA_CONSUMER_WARPS || (threadIdx.x & 31) != 0) return;
  constexpr int KSTAGE = 16 * TMA_KSUB;
  const int nk0 = (depth[0] + KSTAGE - 1) / KSTAGE;
  const int total = tma_total_iters(nseg, depth);
  for (int it = 0; it < total; ++it) {
    const int st = it % TMA_STAGES;
    const uint32_t use = it / TMA_STAGES;
    mbar_wait(&pipe.empty[st], (use & 1) ^ 1);
    const int seg = it < nk0 ? 0 : 1;
    const int k0 = (seg == 0 ? it : it - nk0) * KSTAGE;
    tma_issue_stage<AK, BKM>(pipe.stages + (size_t)st * TMA_STAGE_BYTES, mapA[seg], ra, mapB[seg], rb, nbox_b, k0,
                             &pipe.full[st]);
  }
}

// Consumers (warps 0..7, 2 x 4 over the 128 x 128 tile): grow the register allocation,
// then run `total` pipeline iterations.
__device__ __forceinline__ void tma_consumer_regs() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(TMA_CONSUMER_REGS));
}
template <bool AK, bool BKM, int NW>
__device__ __forceinline__ void tma_consumer(double (&acc)[2 * NW][4][2], const TmaPipe& pipe, int total) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / NW) * (16 * NW), wn0 = (warp % NW) * 32;
  for (int it = 0; it < total; ++it) {
    const int st = it % TMA_STAGES;
    const uint32_t use = it / TMA_STAGES;
    mbar_wait(&pipe.full[st], use & 1);
    const uint8_t* stage = pipe.stages + (size_t)st * TMA_STAGE_BYTES;
#pragma unroll
    for (int sub = 0; sub < TMA_KSUB; ++sub)
      tma_compute_k16<AK, BKM, NW>(acc, stage + sub * 2 * TMA_TILE_BYTES,
                                   stage + sub * 2 * TMA_TILE_BYTES + TMA_TILE_BYTES, wm0, wn0, g, t);
    __syncwarp();
    if (lane == 0) mbar_arrive(&pipe.empty[st]);
  }
}
// NW for a tile with `valid` useful columns (of 128)
__device__ __forceinline__ int tma_nw(int valid) { return valid <= 32 ? 1 : (valid <= 64 ? 2 : 4); }
// barrier among the 256 consumer threads only (the producer warpgroup has left)
__device__ __forceinline__ void tma_consumer_sync() {
  asm volatile("bar.sync 1, %0;\n" ::"n"(TMA_CONSUMER_WARPS * 32) : "memory");
}

}  // namespace scfgp
