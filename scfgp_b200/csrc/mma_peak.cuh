// Measured ceiling of the int8 tensor pipe for the digit-product instruction mix (developer / bench tool):
// the product path's tcgen05.mma stream of one K = 32 step -- same descriptors, same class columns, a
// tcgen05.commit per 64-byte block -- with NO operand traffic (operands sit in shared memory) and NO
// epilogue, on every SM at once.  bench.py reports roofline fractions against this number as well.
#pragma once
#include "umma_i8.cuh"

namespace scfgp {
namespace umma {

template <int KS, int TNW>
__global__ void k_mma_stream(int steps, int random_data, long long* out /*[2 per CTA]: clocks, ok*/) {
  using G = Geo<KS, TNW>;
  extern __shared__ uint8_t raw[];
  const uint32_t off = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
  uint8_t* base = raw + off;                 // A digits at 0, B digits behind them (one pipeline stage)
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + G::STAGE_BYTES);
  uint32_t* slot = reinterpret_cast<uint32_t*>(base + G::STAGE_BYTES + 64);
  uint64_t* bar2 = reinterpret_cast<uint64_t*>(base + G::STAGE_BYTES + 128);   // per-stage commits land here (never waited on)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < G::STAGE_BYTES / 4; i += blockDim.x) {
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
  if (warp == 1 && elect_one()) {
    const uint32_t sa = smem_u32(base), sb = sa + G::A_BYTES;
    constexpr int MAXCNT = G::MAXCNT;
    const long long t0 = clock64();
    bool ok = true;
    int done = 0;
    for (int b = 0; done < steps && ok; ++b) {
      for (int kb = 0; kb < 16 && done < steps; ++kb) {
#pragma unroll
        for (int ks = 0; ks < KB / 32; ++ks, ++done) {
#pragma unroll
          for (int i = 0; i < KS; ++i) {
            const uint64_t adesc = smem_desc_sw64(sa + i * (TM * KB) + ks * 32);
#pragma unroll
            for (int j0 = 0; j0 < KS - i; j0 += MAXCNT) {
              const int cnt = (KS - i - j0) < MAXCNT ? (KS - i - j0) : MAXCNT;
              const uint64_t bdesc = smem_desc_sw64(sb + j0 * (TNW * KB) + ks * 32);
              mma_i8(tmem + (uint32_t)(TNW * (i + j0)), adesc, bdesc, idesc_i8(TNW * cnt), 1u);
            }
          }
        }
        tc_commit(bar2);
      }
      tc_commit(bar);
      // bounded wait: the kernel always ends
      const long long w0 = clock64();
      ok = false;
      while (clock64() - w0 < 400000000ll) {
        uint32_t got;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(got)
            : "r"(smem_u32(bar)), "r"((uint32_t)b & 1u)
            : "memory");
        if (got) { ok = true; break; }
      }
    }
    out[2 * blockIdx.x] = clock64() - t0;
    out[2 * blockIdx.x + 1] = ok ? 1 : 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace umma
}  // namespace scfgp
