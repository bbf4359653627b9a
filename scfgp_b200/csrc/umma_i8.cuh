// Emulated-FP64 GEMM mainloop on the 5th-generation tensor cores (sm_100a):
// tcgen05.mma.kind::i8 with int32 accumulators in tensor memory, operands staged by TMA.
//
// Number format ("digits").  A real matrix X with |X| <= scale/2 is held as KS signed
// base-256 digits  X ~= scale * 2^-(8 KS - 1) * sum_i d_i 256^(KS-1-i),  d_i in [-128, 127]
// (int8; the leading digit stays within +-64 because of the factor-2 headroom, which is what
// keeps the balanced representation inside int8).  A product  sum_c X[m,c] Y[n,c]  is then
//     sum_{i,j} 2^(-8(i+j)) * ( sum_c dX_i[m,c] dY_j[n,c] )   * scaleX scaleY 2^-14 ,
// truncated to the "classes"  i + j = q < KS  (the dropped terms are below the digit
// quantisation error).  Every class sum is an exact int32: |class q| <= depth * 16384 * (KS - 1)
// (leading digit within +-64), so the host refuses reduction depths beyond
// (2^31 - 1) / (16384 (KS - 1)) = 32767 at 5 digits, 18724 at 8.
//
// Tile = 128 rows (TMEM lanes) x 64 columns.  Class q owns TMEM columns [64 q, 64 q + 64):
// all KS classes are resident, so one pass over the contraction does every digit product.
// The B digits of a stage lie slice-major in shared memory ([slice][64 rows][64 B]), so the
// products of A digit i with B digits j..j+3 are ONE instruction of N = 256 whose output
// columns [64 (i+j), 64 (i+j) + 256) are exactly classes i+j .. i+j+3.
//
// Pipeline (no CTA-wide barrier in the loop):
//   warp 0    : TMA producer, one 3-D box per operand per stage (64 contraction bytes,
//               128 or 64 rows, KS slices), 64-byte swizzle, completion on an mbarrier
//   warp 1    : allocates TMEM, one lane issues the tcgen05.mma stream and commits each
//               stage back to the producer (tcgen05.commit -> mbarrier)
//   warps 2-9 : epilogue; tcgen05.ld the class accumulators, combine them in float64
//               registers (2^-8q weights), 16 columns of one row at a time
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_tma.cuh"

namespace scfgp {
namespace umma {

#ifndef UMMA_TN
#define UMMA_TN 64       // 96 (experimental, tools/umma_probe.cu "tn"): 20 % fewer operand bytes per product, <= 5 digits
#endif
constexpr int TM = 128;  // tile rows  (M of the instruction, TMEM lanes)
constexpr int TN = UMMA_TN;   // tile columns per class (default of the TNW template parameters below)
static_assert(TN % 32 == 0 && TN >= 64 && TN <= 128, "tile columns");
constexpr int KB = 64;   // contraction bytes per pipeline stage (one 64-byte swizzle atom)
constexpr int DIGIT_BITS = 8;
constexpr double PROD_SCALE = 1.0 / 16384.0;  // 2^-14: product of the two leading-digit weights 2^-7
constexpr int THREADS = 320;
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 256;
constexpr int MAX_REGS = 96;  // per thread: 320 x 96 leaves more than half of the register file to co-resident kernels
// two-phase epilogue (KS <= 5): a thread parks its TNW / 2 recombined accumulators as int64 (up to 96 registers at
// TNW = 96) so that tensor memory goes back to the MMA warp before the fused epilogue math starts
constexpr int MAX_REGS_2PHASE = 168;  // 10 warps: one scheduler partition (16 K registers) hosts 3 of them -> at most 170 per thread

// TNW: tile columns per class.  64: four B digits per instruction (N = 256), every digit count; 96: two B digits
// per instruction (N = 192), 20 % fewer operand bytes per digit product, at most 5 digits (480 class columns).
template <int KS, int TNW = TN>
struct Geo {
  static constexpr int MAXCNT = 256 / TNW;  // B digits one instruction may span (N = TNW * cnt <= 256)
  static constexpr int A_BYTES = KS * TM * KB;
  static constexpr int B_BYTES = KS * TNW * KB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (224 * 1024) / STAGE_BYTES;
#ifndef UMMA_A_TMEM
#define UMMA_A_TMEM 0
#endif
#ifndef UMMA_COLLECTOR
#define UMMA_COLLECTOR 0
#endif
#ifndef UMMA_MAX_STAGES
#define UMMA_MAX_STAGES 4
#endif
  static constexpr int STAGES = STAGES_RAW > UMMA_MAX_STAGES ? UMMA_MAX_STAGES : STAGES_RAW;
  static constexpr int CLASS_COLS = KS * TNW;
  static constexpr int TMEM_COLS = (CLASS_COLS <= 256 && !UMMA_A_TMEM) ? 256 : 512;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(KS >= 2 && KS <= 8, "2..8 digits");
  static_assert(CLASS_COLS <= 512, "all classes must fit tensor memory");
  static_assert(TNW % 32 == 0 && TNW >= 64 && TNW <= 128, "tile columns");
  static_assert(STAGES >= 2, "need a double buffer");
};

// ---- PTX wrappers ------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// the same load with an L2 eviction-priority hint (createpolicy): the replicated B operand of the streaming products
// ([2Abar | 2 sigma B], [Li; alpha]: 44 MB / 22 MB at K = 2080) is re-read by every tile and should outlive the
// streaming A operand and the result slab in L2
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                 uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;\n" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// one lane of a converged warp (elect.sync): the compiler knows the guarded region runs on a single thread and
// keeps descriptors / addresses in uniform registers without uniformisation loops
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// all tcgen05.mma issued so far by this thread arrive (once) on the mbarrier when they retire
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T,  int8 x int8 -> int32, K = 32 bytes per instruction
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// collector hints: COLL 1 = keep the fetched A slice for the next instruction (fill), 2 = reuse it (lastuse)
template <int COLL>
__device__ __forceinline__ void mma_i8_coll(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  if (COLL == 1)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with A read from tensor memory (lane = tile row, 8 columns = the 32 contraction bytes)
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared memory (128 rows x 32 bytes, described like an MMA operand) -> 128 lanes x 8 columns of tensor memory
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(taddr), "l"(sdesc) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (quarter*32 + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major operand whose rows are 64 bytes apart and
// 64-byte swizzled (8-row groups 512 bytes apart).  `byte_addr` may point 32 bytes into
// the swizzle atom (second K = 32 step): the swizzle is a function of the absolute address.
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t byte_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((byte_addr & 0x3FFFF) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(512 >> 4) << 32;              // stride byte offset: 8 rows x 64 B
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                       // layout type SWIZZLE_64B
  return d;
}
// MN-major ("transposed") operands: the tile lies in shared memory as [contraction row][M or N bytes],
// i.e. straight out of a row-major digit array whose rows are the contraction index.
//   A: 128 M-bytes per row, 128-byte swizzle; 8-row groups 1024 bytes apart (SBO); one 128-byte
//      chunk covers M = 128, so LBO is unused.
//   B: 64 N-bytes per row per digit, 64-byte swizzle; 8-row groups 512 bytes apart (SBO); the next
//      64 N-bytes (= the next digit's block, N = 64 cnt) are `lbo` bytes further (LBO).
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t byte_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                 uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((byte_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_i8_mn(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
}
// instruction descriptor: D = s32, A = B = s8, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_i8(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// wait that is expected to take tens of microseconds (the epilogue warps wait for the whole
// mainloop): let the hardware suspend the thread instead of spinning on the issue slots
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
        : "memory");
  } while (!ok);
}

// Optional in-kernel tile timing (developer tool, scfgp_dbg_tile_timing): when the pointer is set, one thread of
// each role adds its clock64() intervals per tile to 8 counters of its CTA:
//   0 tiles   1 MMA warp waiting for acc_empty   2 MMA warp waiting for operand stages (full barriers)
//   3 MMA warp: whole tile, first wait to last commit   4 epilogue waiting for acc_full
//   5 epilogue: acc_full -> accumulators released   6 epilogue: release -> end of the tile's epilogue
//   7 producer waiting for free stages
__device__ unsigned long long* g_tile_timing = nullptr;
// L2 evict_last hint on the replicated B operand (run_mainloop<.., B_KEEP = true>); SCFGP_B_KEEP=0 switches it off
__device__ int g_b_keep = 1;
// (A-operand collector hints -- fill ... lastuse on the instructions that share an A digit slice -- were measured
// as a run-time knob in round 2: no gain (profiles/r2_mma_issue_and_epilogue.txt), and the extra branch in the
// single-thread issue loop cost 12 % on the Gram kernel.  They remain as the compile-time option UMMA_COLLECTOR.)
constexpr int TT_SLOTS = 8;
__device__ __forceinline__ void tt_add(unsigned long long* tt, int slot, long long v) {
  if (tt) atomicAdd(tt + (size_t)blockIdx.x * TT_SLOTS + slot, (unsigned long long)v);
}

struct Tile {
  int m0;       // first row of the A tensor map (dimension 1)
  int n0;       // first row of the B tensor map (dimension 1)
  int c_begin;  // contraction range in bytes = elements; c_begin multiple of 64
  int c_end;
  int c_off_b = 0;  // the B map's contraction coordinate is c + c_off_b (operands living in different arrays)
  // optional first segment of the A operand from another array (same column layout): stages
  // kb < seg1_stages read map `mapA1` at rows m0 + seg1_row_off instead of mapA
  const CUtensorMap* mapA1 = nullptr;
  int seg1_stages = 0;
  int seg1_row_off = 0;
};

struct Pipe {
  uint8_t* stages;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* acc_full;   // MMA -> epilogue: the tile's accumulators are complete
  uint64_t* acc_empty;  // epilogue -> MMA: the accumulators have been read, the next tile may overwrite them
  uint32_t* tmem_slot;
};

template <int KS, int TNW = TN>
__device__ __forceinline__ Pipe pipe_init(uint8_t* smem_raw) {
  using G = Geo<KS, TNW>;
  Pipe p;
  const uint32_t off = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  p.stages = smem_raw + off;
  p.full = reinterpret_cast<uint64_t*>(p.stages + (size_t)G::STAGES * G::STAGE_BYTES);
  p.empty = p.full + G::STAGES;
  p.acc_full = p.empty + G::STAGES;
  p.acc_empty = p.acc_full + 1;
  p.tmem_slot = reinterpret_cast<uint32_t*>(p.acc_empty + 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(&p.full[s], 1);
      mbar_init(&p.empty[s], 1);
    }
    mbar_init(p.acc_full, 1);
    mbar_init(p.acc_empty, EPI_THREADS / 32);
    mbar_fence_init();
  }
  if ((threadIdx.x >> 5) == 1) tmem_alloc(p.tmem_slot, G::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return p;
}

template <int KS, int TNW = TN>
__device__ __forceinline__ void pipe_teardown(const Pipe& p, uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Geo<KS, TNW>::TMEM_COLS);
  }
}

// Persistent CTAs: every CTA walks the tiles  blockIdx.x, blockIdx.x + gridDim.x, ...  (`seq` =
// 0, 1, 2, ... counts the tiles of this CTA).  The producer keeps streaming stages across tile
// boundaries (the ring's phases simply continue: `it0` = stages consumed by earlier tiles), so
// the next tile's operands are already in flight while the epilogue drains the accumulators;
// the MMA warp only has to wait for the epilogue to release TMEM (acc_empty).
__device__ __forceinline__ int tile_stages(const Tile& t) { return (t.c_end - t.c_begin + KB - 1) / KB; }

// Bottleneck probe (tools/umma_probe.cu only): UMMA_PROBE_MODE 1 = the loads run but no MMA is
// issued (pure TMA pipeline time), 2 = only the first ring of stages is loaded and the MMAs re-read
// it (pure MMA + shared-memory-read time).  Results are garbage in both modes.
#ifndef UMMA_PROBE_MODE
#define UMMA_PROBE_MODE 0
#endif

// producer + MMA roles of one tile (warps 0 and 1)
// TR = false: K-major operands (tensor maps (contraction, row, digit), tile rows m0 / n0);
// TR = true : MN-major operands (tensor maps (M-or-N index, contraction row, digit)): m0 / n0 are
//             coordinates of dimension 0, the contraction range [c_begin, c_end) of dimension 1.
// B_KEEP: the B operand's loads carry the L2 evict_last hint (K-major operands only)
template <int KS, bool TR = false, int TNW = TN, bool B_KEEP = false>
__device__ __forceinline__ void run_mainloop(const Pipe& p, uint32_t tmem_base, const CUtensorMap* mapA,
                                             const CUtensorMap* mapB, const Tile& t, int it0, int seq = 0) {
  using G = Geo<KS, TNW>;
  constexpr int MAXCNT = G::MAXCNT;
  static_assert(!TR || TNW == 64, "MN-major B boxes are 64 bytes wide");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (t.c_end - t.c_begin + KB - 1) / KB;
  unsigned long long* const tt = g_tile_timing;
  if (warp == 0) {
    if (elect_one()) {
      long long w_empty = 0;
      uint64_t pol_b = 0;
      if constexpr (B_KEEP) pol_b = g_b_keep ? l2_policy_evict_last() : 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int it = it0 + kb;
        const int st = it % G::STAGES;
        const uint32_t use = it / G::STAGES;
        const long long c0 = tt ? clock64() : 0;
        mbar_wait(&p.empty[st], (use & 1) ^ 1);
        if (tt) w_empty += clock64() - c0;
        uint8_t* sa = p.stages + (size_t)st * G::STAGE_BYTES;
#if UMMA_PROBE_MODE == 2
        if (it >= G::STAGES) { mbar_arrive(&p.full[st]); continue; }
#endif
        mbar_arrive_expect_tx(&p.full[st], G::STAGE_BYTES);
        const int c = t.c_begin + kb * KB;
        if (TR) {
          tma_load_3d(sa, mapA, t.m0, c, 0, &p.full[st]);
          tma_load_3d(sa + G::A_BYTES, mapB, t.n0, c + t.c_off_b, 0, &p.full[st]);
        } else {
          if (kb < t.seg1_stages) tma_load_3d(sa, t.mapA1, c, t.m0 + t.seg1_row_off, 0, &p.full[st]);
          else tma_load_3d(sa, mapA, c, t.m0, 0, &p.full[st]);
          if (B_KEEP && pol_b != 0) tma_load_3d_hint(sa + G::A_BYTES, mapB, c + t.c_off_b, t.n0, 0, &p.full[st], pol_b);
          else tma_load_3d(sa + G::A_BYTES, mapB, c + t.c_off_b, t.n0, 0, &p.full[st]);
        }
      }
      tt_add(tt, 7, w_empty);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const long long t_begin = tt ? clock64() : 0;
      long long w_full = 0;
      if (seq > 0) {
        mbar_wait(p.acc_empty, (uint32_t)(seq - 1) & 1u);
        tc_fence_after();
      }
      if (tt) tt_add(tt, 1, clock64() - t_begin);
      for (int kb = 0; kb < nkb; ++kb) {
        const int it = it0 + kb;
        const int st = it % G::STAGES;
        const uint32_t use = it / G::STAGES;
        const long long c0 = tt ? clock64() : 0;
        mbar_wait(&p.full[st], use & 1);
        if (tt) w_full += clock64() - c0;
        tc_fence_after();
        const uint32_t sa = smem_u32(p.stages + (size_t)st * G::STAGE_BYTES);
        const uint32_t sb = sa + G::A_BYTES;
        // descriptors of the stage's first A / B digit block; every instruction adds a compile-time offset to
        // the 14-bit start-address field (16-byte units; the sum is a valid shared-memory address, no carry out)
        const uint64_t adesc0 = TR ? smem_desc_mn(sa, 0, 8 * TM, 2 /*SWIZZLE_128B*/) : smem_desc_sw64(sa);
        const uint64_t bdesc0 = TR ? smem_desc_mn(sb, TNW * KB, 8 * TNW, 4 /*SWIZZLE_64B*/) : smem_desc_sw64(sb);
#if UMMA_A_TMEM
        // A digits staged into tensor memory behind the class accumulators (double-buffered per stage)
        const uint32_t abuf = tmem_base + (uint32_t)(TNW * KS + (it & 1) * (KS * 16));
        if (!TR) {
#pragma unroll
          for (int ks = 0; ks < KB / 32; ++ks)
#pragma unroll
            for (int i = 0; i < KS; ++i)
              tmem_cp_128x256b(abuf + (uint32_t)((ks * KS + i) * 8), smem_desc_sw64(sa + i * (TM * KB) + ks * 32));
        }
#endif
#pragma unroll
        for (int ks = 0; ks < KB / 32; ++ks) {
#if UMMA_PROBE_MODE == 3 || UMMA_PROBE_MODE == 4
          // probe: the same number of output columns, but as uniform N = 256 (mode 3) or N = 64 (mode 4)
          // instructions -- separates the per-instruction overhead from the shared-memory read volume
          {
            constexpr int COLS = TNW * KS * (KS + 1) / 2;
            constexpr int NI = UMMA_PROBE_MODE == 3 ? 256 : 64;
#pragma unroll
            for (int m = 0; m < (COLS + NI - 1) / NI; ++m) {
              const uint64_t adesc = smem_desc_sw64(sa + (m % KS) * (TM * KB) + ks * 32);
              const uint64_t bdesc = smem_desc_sw64(sb + ks * 32);
              mma_i8(tmem_base, adesc, bdesc, idesc_i8(NI), (kb | ks | m) ? 1u : 0u);
            }
          }
          continue;
#endif
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            // one digit block is TM x KB = KB x TM bytes in either layout
            const uint64_t adesc = adesc0 + (uint64_t)((i * (TM * KB) + ks * 32 * (TR ? TM : 1)) >> 4);
#pragma unroll
            for (int j0 = 0; j0 < KS - i; j0 += MAXCNT) {
              const int cnt = (KS - i - j0) < MAXCNT ? (KS - i - j0) : MAXCNT;
              const uint64_t bdesc = bdesc0 + (uint64_t)((j0 * (TNW * KB) + ks * 32 * (TR ? TNW : 1)) >> 4);
              const uint32_t acc = (kb == 0 && ks == 0 && i == 0) ? 0u : 1u;
#if UMMA_PROBE_MODE != 1
#if UMMA_A_TMEM
              if (!TR)
                mma_i8_ts(tmem_base + (uint32_t)(TNW * (i + j0)), abuf + (uint32_t)((ks * KS + i) * 8), bdesc,
                          idesc_i8(TNW * cnt), acc);
              else
#endif
#if UMMA_COLLECTOR
              if (KS - i > MAXCNT) {   // several instructions share this A slice
                if (j0 + MAXCNT < KS - i)            // not the last one: (re)fill the collector
                  mma_i8_coll<1>(tmem_base + (uint32_t)(TNW * (i + j0)), adesc, bdesc,
                                 TR ? idesc_i8_mn(TNW * cnt) : idesc_i8(TNW * cnt), acc);
                else
                  mma_i8_coll<2>(tmem_base + (uint32_t)(TNW * (i + j0)), adesc, bdesc,
                                 TR ? idesc_i8_mn(TNW * cnt) : idesc_i8(TNW * cnt), acc);
              } else
#endif
              mma_i8(tmem_base + (uint32_t)(TNW * (i + j0)), adesc, bdesc,
                     TR ? idesc_i8_mn(TNW * cnt) : idesc_i8(TNW * cnt), acc);
#else
              (void)adesc; (void)bdesc; (void)acc;
#endif
            }
          }
        }
        tc_commit(&p.empty[st]);  // stage reusable once these MMAs have read it
      }
      tc_commit(p.acc_full);  // accumulators complete
      if (tt) {
        tt_add(tt, 0, 1);
        tt_add(tt, 2, w_full);
        tt_add(tt, 3, clock64() - t_begin);
      }
    }
  }
}

// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// epilogue role (warps 2..9).  Each warp owns one TMEM lane quarter (32 tile rows, one per
// lane) and one half (32 columns) of the tile, handled as two groups of 16 columns so that
// the kernels stay within 96 registers per thread (side kernels of other streams can then
// share the SM with a resident tensor-core CTA).
struct EpiPos {
  int row;    // tile row of this thread (0..127)
  int col0;   // first of its TN / 2 tile columns (0 or TN / 2)
  uint32_t lane_base;  // TMEM address of lane quarter, column 0
};
template <int TNW = TN>
__device__ __forceinline__ EpiPos epilogue_wait(const Pipe& p, uint32_t tmem_base, uint32_t acc_parity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;              // TMEM lane quarter this warp may touch
  const int half = (warp - EPI_WARP0) >> 2;  // which half of the TN columns
  EpiPos e;
  e.row = quarter * 32 + lane;
  e.col0 = half * (TNW / 2);
  e.lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
  mbar_wait_long(p.acc_full, acc_parity);
  tc_fence_after();
  return e;
}
// after the last epilogue_load16 of a tile: hand the accumulators back to the MMA warp
__device__ __forceinline__ void epilogue_release(const Pipe& p) {
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(p.acc_empty);
}
// v[c] = sum_q 2^(-8q) acc_q[row][col + c], c < 16   (col = tile column, multiple of 16)
// KS <= 5: exact int64 recombination and one rounding (the same values the two-phase epilogue produces);
// KS >= 6: float64 fma chain, least significant class first
template <int KS, int TNW = TN>
__device__ __forceinline__ void epilogue_load16(const EpiPos& e, int col, double (&v)[16]) {
  if constexpr (KS <= 5) {
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.0;
    long long a[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) a[c] = 0;
#pragma unroll
    for (int q = 0; q < KS; ++q) {
      uint32_t r[16];
      tmem_ld16(e.lane_base + (uint32_t)(q * TNW + col), r);
#pragma unroll
      for (int c = 0; c < 16; ++c) a[c] = a[c] * 256 + (long long)(int)r[c];
    }
    const double w = __longlong_as_double((long long)(1023 - DIGIT_BITS * (KS - 1)) << 52);
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = (double)a[c] * w;
    return;
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = 0.0;
#pragma unroll
  for (int q = KS - 1; q >= 0; --q) {  // least significant class first
    uint32_t r[16];
    tmem_ld16(e.lane_base + (uint32_t)(q * TNW + col), r);
    const double w = __longlong_as_double((long long)(1023 - DIGIT_BITS * q) << 52);  // 2^(-8q)
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = fma((double)(int)r[c], w, v[c]);
  }
}


// ---- two-phase epilogue -------------------------------------------------------------------------------------
// Phase 1 ("drain"): tcgen05.ld all KS class accumulators of 16 columns back to back (one wait), recombine
// them EXACTLY in int64,  S = sum_q acc_q 256^(KS-1-q)  (|S| < 2^62 for KS <= 5: |acc_q| < 2^31), and keep S.
// After the last drain of a tile the accumulators are released (epilogue_release) and the MMA warp starts the
// next tile while phase 2 (scales, fused math, stores) runs on the parked values:  value = S 2^(-8 (KS-1)).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// the loaded registers become valid here: the in/out operands keep their consumers behind the wait
__device__ __forceinline__ void tmem_ld16_landed(uint32_t (&r)[16], bool wait) {
  if (wait)
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
  else
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
template <int KS, int TNW = TN>
__device__ __forceinline__ void epilogue_drain16(const EpiPos& e, int col, long long (&s)[16]) {
  static_assert(KS <= 5, "exact int64 recombination needs 8 (KS - 1) + 31 < 63");
  // two batches of classes (3 + the rest): at most 48 transient registers on top of the parked values
  constexpr int B0 = KS < 3 ? KS : 3;
  {
    uint32_t r[B0][16];
#pragma unroll
    for (int q = 0; q < B0; ++q) tmem_ld16_issue(e.lane_base + (uint32_t)(q * TNW + col), r[q]);
#pragma unroll
    for (int q = 0; q < B0; ++q) tmem_ld16_landed(r[q], q == 0);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      long long a = 0;
#pragma unroll
      for (int q = 0; q < B0; ++q) {
        const int sh = DIGIT_BITS * (KS - 1 - q);
        if (sh < 31) a += (long long)(int)r[q][c] * (long long)(1 << sh);   // mul.wide.s32 + add
        else a += (long long)(int)r[q][c] << sh;
      }
      s[c] = a;
    }
  }
  if constexpr (KS > B0) {
    uint32_t r[KS - B0][16];
#pragma unroll
    for (int q = B0; q < KS; ++q) tmem_ld16_issue(e.lane_base + (uint32_t)(q * TNW + col), r[q - B0]);
#pragma unroll
    for (int q = B0; q < KS; ++q) tmem_ld16_landed(r[q - B0], q == B0);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      long long a = s[c];
#pragma unroll
      for (int q = B0; q < KS; ++q) {
        const int sh = DIGIT_BITS * (KS - 1 - q);
        a += (long long)(int)r[q - B0][c] * (long long)(1 << sh);
      }
      s[c] = a;
    }
  }
}
// phase 2: the parked integer as a float64 (one rounding; exact below 2^53)
template <int KS>
__device__ __forceinline__ double drained_value(long long s) {
  return (double)s * __longlong_as_double((long long)(1023 - DIGIT_BITS * (KS - 1)) << 52);
}

}  // namespace umma
}  // namespace scfgp
