// C ABI + host orchestration of the SCFGP hot path on one B200 (sm_100a).
// Interface: include/scfgp_b200.h.  Math: SURVEY.md Appendix A (SCFGP/SCFGP.py:74-148).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <functional>

#include <cuda_runtime.h>

#include "../../include/scfgp_b200.h"
#include "kernels.cuh"
#include "kernels_tma.cuh"
#include "emu_kernels.cuh"

using namespace scfgp;

namespace {

constexpr double kEps = 1e-6;  // jitter, SCFGP/SCFGP.py:93
constexpr int kColSplit = 64;  // row splits of the column-sum kernels (parallelism: ceil(K/256) x 64 CTAs)
constexpr int kSeedBlocks = 296;
constexpr int kGramSplit = 1;  // row splits of one chunk's Gram tiles in the emulated mode (work items = tiles x splits)
constexpr int kDwSplit = 16;   // row splits of the dW = XT.Zbar contraction (partial buffers per chunk stream)

struct ItemList {
  int4* d = nullptr;
  int count = 0;
};

constexpr int kMaxStreams = 4;
struct Ctx {
  cudaStream_t st = nullptr;   // chunk stream (high priority): features, digits, reductions
  cudaStream_t sg = nullptr;   // low-priority stream of the tcgen05 GEMM kernels of this chunk stream
  cudaEvent_t ev_sg = nullptr, ev_gs = nullptr;  // st -> sg and sg -> st hand-offs
  cudaEvent_t done = nullptr;
  double *Phi = nullptr, *Psi = nullptr, *Zbar = nullptr, *vpart = nullptr;  // slabs
  double *acc = nullptr;      // K x ldk partial of G / H
  double *dW = nullptr;       // D x ldj partial
  double *colpart = nullptr;  // kColSplit x K
  double *slots = nullptr;    // slot_stride scalars
  CUtensorMap mapPhi, mapPsi; // 16x16-box TMA views of the slabs
  // emulated-FP64 path (int8 digits, see emu_kernels.cuh)
  int8_t *digR = nullptr;     // [KS][R][2 Kp]  row digits of [Phi | Psi] (Psi = gv Phi / sigma; feature index contiguous)
  double *sig = nullptr;      // {sigma_chunk, 1/sigma_chunk}
  double *mu = nullptr;       // mu_n = Phi_n.alpha of the chunk (column K of the Phi.[Li; alpha]T product)
  CUtensorMap mapRowsA;                                   // K-major view of digR (contraction = feature), ks2 planes
  // MN-major views of digR (contraction = chunk row) for the Gram products
  CUtensorMap mapPhiMN_A1, mapPhiMN_B1;                   // Phi half, ks planes (pass 1: G)
  CUtensorMap mapPhiMN_B, mapPsiMN_A;                     // Phi / Psi halves, ks2 planes (pass 2: H)
  CUtensorMap mapPhiMN_A, mapPsiMN_B;                     // the same with the box widths swapped (border tiles)
};

}  // namespace

struct scfgp_handle {
  int D = 0, S = 0, M = 0, J = 0, K = 0, fftype = 0, device = 0;
  long long P = 0;
  cudaStream_t stream = nullptr;
  long long N = 0, Nglobal = 0;
  const double* dX = nullptr;
  const double* dy = nullptr;
  double* ownX = nullptr;
  double* owny = nullptr;
  long long own_cap = 0;                 // rows the owned buffers can hold
  // host -> device upload in row blocks on a copy stream, overlapped with the first pass
  cudaStream_t cstream = nullptr;
  std::vector<cudaEvent_t> up_ev;        // one event per uploaded row block
  long long up_block = 0;                // rows per block
  bool upload_pending = false;
  int R = 4096;  // rows per chunk (8192 in the emulated mode unless set by the caller)
  bool R_user = false;
  int xvec = 2;
  long long ldk = 0;  // leading dimension of K x K matrices and slabs
  long long ldj = 0;  // leading dimension of D x J / Zbar
  long long ldw = 0;

  // device buffers
  double *Wt = nullptr, *phi = nullptr;
  // per-stream chunk contexts: consecutive row chunks run on different streams so the tail
  // of one chunk's kernels overlaps the head of the next; each context owns its slabs and
  // partial accumulators (deterministic: chunk c always lands in context c % ns)
  Ctx ctx[kMaxStreams];
  int ns = 4;
  cudaEvent_t ev_fork = nullptr;
  double *acc_all = nullptr, *dW_all = nullptr, *colpart_all = nullptr, *slots_all = nullptr;
  long long slot_stride = 0;
  int use_tma = 1;
  CUtensorMap mapLi, mapA2, mapB2;
  double *red1 = nullptr, *red2 = nullptr, *red3 = nullptr;
  double *L = nullptr, *Li = nullptr, *Bm = nullptr, *BH = nullptr, *BHB = nullptr, *A2 = nullptr,
         *B2 = nullptr, *Dinv = nullptr, *tmpS = nullptr;
  double *alpha = nullptr, *beta = nullptr, *tvec = nullptr, *pbar = nullptr, *scal = nullptr;
  double *gmu = nullptr, *gv = nullptr;
  long long row_cap = 0;
  double *colpart = nullptr, *slots = nullptr, *trslots = nullptr;   // main-stream scratch
  long long* info = nullptr;
  int nrb_cap = 0;

  // emulated-FP64 path: precision mode, digit count and the replicated digit operands
  int prec = SCFGP_PREC_FP64;
  int num_sms = 148;
  int ks = 5;                  // digits of the Phi^T Phi Gram (pass 1: feeds the Cholesky, the sensitive product)
  int ks2 = 5;                 // digits of every other streaming product (Phi.LiT, PsiT.Phi, backward)
  long long Kp = 0;            // K rounded up to 64 (digit row pitch / offset of the Psi half)
  double s2 = 1.0;             // power of two >= s: scale of the Phi digits
  int8_t *digLi = nullptr;     // [KS][K][Kp]
  int8_t *digC = nullptr;      // [KS][K][2 Kp]   rows permuted (cos/sin interleaved): [2Abar | 2 sigma B]
  double *liscale = nullptr, *liinv = nullptr, *cscale = nullptr, *cinv = nullptr, *sigg = nullptr;
  unsigned long long* gvmax = nullptr;
  CUtensorMap mapLiB, mapCB;
  // optional cache of Phi's row digits for ALL rows ([ks][cache_rows][Kc]; 104 GB at cfg4): passes 2 and 3
  // then read Phi from it instead of recomputing features + digits
  int8_t* phi_cache = nullptr;
  long long cache_rows = 0, Kc = 0;
  int cache_ks = 0;
  int want_cache = 1;          // SCFGP_PHI_CACHE=0 disables
  CUtensorMap mapCacheK, mapCacheMN_A, mapCacheMN_B;
  int4* gram_items = nullptr;
  int gram_count = 0;
  int vparts = 0;              // partial rows of vpart per chunk
  int gsplit = 1;              // partial Gram accumulators per chunk stream (kGramSplit in the emulated mode)
  int csplit = kColSplit;      // partial rows of the column-sum buffers per chunk stream (>= R/128)

  // work lists
  ItemList lowerK, syrkK, fullK, tgemm, bwd, dw;
  std::vector<ItemList> potrf_panel, potrf_trail;
  std::vector<ItemList> inv_t, inv_x;   // per level of the block-recursive triangular inverse: T = L21.Li11, Li21 = -Li22.T
  std::vector<int4*> to_free;

  // host state of the evaluation in flight
  std::vector<double> params, lF, rF, F, W, phi_h;
  double a = 0, b = 0, c = 0, sig2n = 0, sigf = 0, tau = 0, s = 0;
  int want_grad = 0;
  bool have_factors = false;

  std::string err;
  long long pivot = -1;
  long long launches = 0;
  cudaEvent_t ev[8] = {};
  float ms_pass[4] = {0, 0, 0, 0};

  // optional launch trace (CUDA event pair around every streaming-kernel launch)
  struct TraceRec { const char* name; int lane; cudaEvent_t e0, e1; };
  bool tracing = false;
  std::vector<TraceRec> trace;

  // optional per-kernel-family timing (CUDA events around every launch of one family)
  int prof_family = SCFGP_FAM_NONE;
  std::vector<cudaEvent_t> prof_ev;   // pairs
  size_t prof_used = 0;
  double prof_ms = 0;
  long long prof_launches = 0;
};

namespace {

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      return SCFGP_ERR_CUDA;                                                            \
    }                                                                                   \
  } while (0)

#define CKL()                                                                           \
  do {                                                                                  \
    h->launches++;                                                                      \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) {                                                            \
      h->err = std::string("kernel launch: ") + cudaGetErrorString(e_) + " at line " +  \
               std::to_string(__LINE__);                                                \
      return SCFGP_ERR_CUDA;                                                            \
    }                                                                                   \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// 2-D row-major float64 view with 16 x 16 boxes, 128-byte swizzle, zero fill out of bounds
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
int make_tmap(scfgp_handle* h, CUtensorMap* out, const double* base, long long cols, long long rows, long long ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { h->err = "cuTensorMapEncodeTiled not available"; return SCFGP_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {16, 16};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return SCFGP_ERR_CUDA; }
  return SCFGP_OK;
}

// 3-D int8 view of a digit array [ks][rows][pitch] (contraction index contiguous): boxes of
// 64 contraction bytes x box_rows rows x ks digits, 64-byte swizzle, zero fill out of bounds
int make_tmap_digits(scfgp_handle* h, CUtensorMap* out, const int8_t* base, long long depth, long long rows,
                     long long pitch, int ks, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { h->err = "cuTensorMapEncodeTiled not available"; return SCFGP_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)depth, (cuuint64_t)rows, (cuuint64_t)ks};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)umma::KB, (cuuint32_t)box_rows, (cuuint32_t)ks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled(digits) failed: " + std::to_string((int)r); return SCFGP_ERR_CUDA; }
  return SCFGP_OK;
}

// MN-major view of a row-major digit array [ks][rows][pitch]: dimension 0 = the (feature) index that is
// contiguous in memory, dimension 1 = the array's rows = the contraction index of the Gram products;
// boxes of box_w (128: A operand, 128-byte swizzle; 64: B operand, 64-byte swizzle) bytes x 64 rows x ks digits
int make_tmap_digits_mn(scfgp_handle* h, CUtensorMap* out, const int8_t* base, long long width, long long rows,
                        long long pitch, int ks, int box_w) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { h->err = "cuTensorMapEncodeTiled not available"; return SCFGP_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)ks};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)umma::KB, (cuuint32_t)ks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_w == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled(digits, MN) failed: " + std::to_string((int)r); return SCFGP_ERR_CUDA; }
  return SCFGP_OK;
}

// run `...` with a compile-time digit count KS matching the handle's runtime one
#define EMU_KS_SWITCH(ks, ...)                                   \
  switch (ks) {                                                  \
    case 4: { constexpr int KS = 4; __VA_ARGS__; } break;        \
    case 5: { constexpr int KS = 5; __VA_ARGS__; } break;        \
    case 6: { constexpr int KS = 6; __VA_ARGS__; } break;        \
    case 7: { constexpr int KS = 7; __VA_ARGS__; } break;        \
    case 8: { constexpr int KS = 8; __VA_ARGS__; } break;        \
    default: h->err = "unsupported digit count"; return SCFGP_ERR_ARG; \
  }

// profiling brackets: record an event pair around launches of the selected family
inline void trace_begin(scfgp_handle* h, const char* name, cudaStream_t st);
inline void trace_end(scfgp_handle* h, cudaStream_t st);
inline const char* fam_name(int fam) {
  switch (fam) {
    case SCFGP_FAM_FEATURES: return "features";
    case SCFGP_FAM_TGEMM: return "tgemm";
    case SCFGP_FAM_BACKWARD: return "backward";
    case SCFGP_FAM_SYRK: return "gram";
    default: return "other";
  }
}
inline void prof_begin(scfgp_handle* h, int fam, cudaStream_t st) {
  trace_begin(h, fam_name(fam), st);
  if (h->prof_family != fam) return;
  if (h->prof_used + 2 > h->prof_ev.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    h->prof_ev.push_back(a); h->prof_ev.push_back(b);
  }
  cudaEventRecord(h->prof_ev[h->prof_used], st);
}
inline void prof_end(scfgp_handle* h, int fam, cudaStream_t st) {
  trace_end(h, st);
  if (h->prof_family != fam) return;
  cudaEventRecord(h->prof_ev[h->prof_used + 1], st);
  h->prof_used += 2;
}
inline void prof_collect(scfgp_handle* h) {
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]) == cudaSuccess) {
      h->prof_ms += ms; h->prof_launches++;
    } else (void)cudaGetLastError();
  }
  h->prof_used = 0;
}

// launch trace: lane = index of the chunk stream (or -1 for the handle's main stream)
inline int trace_lane(const scfgp_handle* h, cudaStream_t st) {
  for (int i = 0; i < kMaxStreams; ++i) {
    if (h->ctx[i].st == st) return i;
    if (h->ctx[i].sg == st) return kMaxStreams + i;
  }
  return -1;
}
inline void trace_begin(scfgp_handle* h, const char* name, cudaStream_t st) {
  if (!h->tracing) return;
  scfgp_handle::TraceRec r;
  r.name = name;
  r.lane = trace_lane(h, st);
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  h->trace.push_back(r);
}
inline void trace_end(scfgp_handle* h, cudaStream_t st) {
  if (!h->tracing || h->trace.empty()) return;
  cudaEventRecord(h->trace.back().e1, st);
}

template <class T>
int dalloc(scfgp_handle* h, T** p, size_t n) {
  CK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  CK(cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
  return SCFGP_OK;
}

int upload_items(scfgp_handle* h, const std::vector<int4>& v, ItemList* out) {
  out->count = (int)v.size();
  out->d = nullptr;
  if (v.empty()) return SCFGP_OK;
  CK(cudaMalloc((void**)&out->d, v.size() * sizeof(int4)));
  CK(cudaMemcpy(out->d, v.data(), v.size() * sizeof(int4), cudaMemcpyHostToDevice));
  h->to_free.push_back(out->d);
  return SCFGP_OK;
}

// lower-triangular tile list of an m x m output with square tiles `bs`
std::vector<int4> lower_tiles(int m, int bs, int k0, int depth) {
  std::vector<int4> v;
  int nb = cdiv(m, bs);
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j <= i; ++j) v.push_back(make_int4(i, j, k0, depth));
  return v;
}

template <class Cfg, bool AK, bool BKM, int AVEC, int BVEC>
int launch_gemm(scfgp_handle* h, const double* A, long long lda, int a_rows, const double* B,
                long long ldb, int b_rows, double* C, long long ldc, double alpha, double beta,
                const ItemList& items, int depth_arg, cudaStream_t st = nullptr, bool use_st = false) {
  if (items.count == 0) return SCFGP_OK;
  if (!use_st) st = h->stream;
  constexpr size_t smem = SmemSize<Cfg, AK, BKM>::BYTES;
  static bool attr_done = false;
  if (!attr_done) {
    CK(cudaFuncSetAttribute(k_gemm<Cfg, AK, BKM, AVEC, BVEC>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  k_gemm<Cfg, AK, BKM, AVEC, BVEC><<<items.count, Cfg::NT, smem, st>>>(
      A, lda, a_rows, B, ldb, b_rows, C, ldc, alpha, beta, items.d, depth_arg);
  CKL();
  return SCFGP_OK;
}

int set_kernel_attrs(scfgp_handle* h) {
  CK(cudaFuncSetAttribute(k_features<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SmemSize<CfgFeat, true, true>::BYTES));
  CK(cudaFuncSetAttribute(k_features<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SmemSize<CfgFeat, true, true>::BYTES));
  CK(cudaFuncSetAttribute(k_tgemm_rowsq, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SmemSize<CfgBig, true, true>::BYTES));
  CK(cudaFuncSetAttribute(k_backward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SmemSize<CfgBig, true, true>::BYTES));
  CK(cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM));
  CK(cudaFuncSetAttribute(k_dw_splitk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SmemSize<CfgDw, false, false>::BYTES));
  CK(cudaFuncSetAttribute(k_dw_splitk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)SmemSize<CfgDw, false, false>::BYTES));
  CK(cudaFuncSetAttribute(k_syrk_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_tgemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_backward_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_BYTES));
  // Kernels of different chunk streams share SMs (a resident tcgen05 CTA leaves registers and
  // ~40 KB of shared memory free).  That only happens if they agree on the L1 / shared-memory
  // split, so every streaming kernel asks for the maximum shared-memory carveout.
  const int mx = (int)cudaSharedmemCarveoutMaxShared;
  CK(cudaFuncSetAttribute(k_features<1>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(k_features<2>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(k_dw_splitk<1>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(k_dw_splitk<2>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(k_colsum, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(k_seeds, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(k_pred_finish, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(emu::k_absmax_pow2, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  return SCFGP_OK;
}

template <int KS>
int set_emu_attrs(scfgp_handle* h) {
  const int smem = (int)umma::Geo<KS>::SMEM;
  CK(cudaFuncSetAttribute(emu::k_gram_i8<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(emu::k_tgemm_i8<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(emu::k_backward_i8<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(emu::k_gemm_i8_dbg<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int mx = (int)cudaSharedmemCarveoutMaxShared;
  CK(cudaFuncSetAttribute(emu::k_gram_i8<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(emu::k_tgemm_i8<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(emu::k_backward_i8<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(emu::k_digits_rows<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute((emu::k_digits_slab<KS, false>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute((emu::k_digits_slab<KS, true>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  return SCFGP_OK;
}

// ---------------------------------------------------------------------------------
// host-side unpack of the flat parameter vector            SCFGP/SCFGP.py:74-90, A.1
// ---------------------------------------------------------------------------------
int unpack_and_upload(scfgp_handle* h, const double* params) {
  const int D = h->D, S = h->S, M = h->M, J = h->J;
  h->params.assign(params, params + h->P);
  const double* p = h->params.data();
  h->a = p[0]; h->b = p[1]; h->c = p[2];
  const double* lf = p + 3;
  const double* rf = lf + (size_t)D * S;
  const double* lp = rf + (size_t)M * S;
  const double* pp = lp + S;
  h->lF.assign(lf, lf + (size_t)D * S);
  h->rF.assign(rf, rf + (size_t)M * S);
  h->F.assign((size_t)D * M, 0.0);
  for (int d = 0; d < D; ++d)
    for (int m = 0; m < M; ++m) {
      double acc = 0;
      for (int s2 = 0; s2 < S; ++s2) acc += lf[(size_t)d * S + s2] * rf[(size_t)m * S + s2];
      h->F[(size_t)d * M + m] = acc;
    }
  h->phi_h.assign(J, 0.0);
  if (h->fftype == SCFGP_FF_PHASE) {
    for (int s2 = 0; s2 < S; ++s2) {
      double mean = 0;
      for (int d = 0; d < D; ++d) mean += lf[(size_t)d * S + s2];
      h->phi_h[s2] = lp[s2] - mean / D;
    }
    for (int m = 0; m < M; ++m) {
      double mean = 0;
      for (int d = 0; d < D; ++d) mean += h->F[(size_t)d * M + m];
      h->phi_h[S + m] = pp[m] - mean / D;
    }
  }
  // Wt (J x ldw): row j = column j of W = [l_F | F]
  std::vector<double> wt((size_t)J * h->ldw, 0.0);
  for (int d = 0; d < D; ++d) {
    for (int s2 = 0; s2 < S; ++s2) wt[(size_t)s2 * h->ldw + d] = lf[(size_t)d * S + s2];
    for (int m = 0; m < M; ++m) wt[(size_t)(S + m) * h->ldw + d] = h->F[(size_t)d * M + m];
  }
  h->sig2n = std::exp(2 * h->a);
  h->sigf = std::exp(h->b);
  h->tau = std::log(1 + std::exp(h->c));
  h->s = h->sigf * std::sqrt(2.0 / M);
  {
    int e;
    const double m = std::frexp(h->s, &e);
    h->s2 = std::ldexp(1.0, (m == 0.5 ? e - 1 : e) + 1);  // digit scale of Phi: power of two >= 2 s = 2 max |Phi|
  }
  CK(cudaMemcpyAsync(h->Wt, wt.data(), wt.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->phi, h->phi_h.data(), (size_t)J * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));  // wt is a stack-lifetime staging buffer
  return SCFGP_OK;
}

int ensure_row_arrays(scfgp_handle* h, long long n) {
  if (n <= h->row_cap) return SCFGP_OK;
  if (h->gmu) cudaFree(h->gmu);
  if (h->gv) cudaFree(h->gv);
  h->gmu = h->gv = nullptr;
  int rc;
  if ((rc = dalloc(h, &h->gmu, (size_t)n))) return rc;
  if ((rc = dalloc(h, &h->gv, (size_t)n))) return rc;
  h->row_cap = n;
  return SCFGP_OK;
}

// Exactness of the emulated path: every class sum must fit int32.  With the leading digit within
// +-64 and the others within +-128 the largest class (q = ks - 1) is bounded by
// depth * 16384 * (ks - 1), so reduction depths (rows per chunk for the Gram products, Kp + K for
// the backward product) are limited to (2^31 - 1) / (16384 (ks - 1)).
inline long long emu_max_depth(int ks) { return 2147483647ll / (16384ll * (ks - 1)); }
int emu_check_depths(scfgp_handle* h) {
  if (h->prec != SCFGP_PREC_EMU_I8) return SCFGP_OK;
  const long long lim = emu_max_depth(std::max(h->ks, h->ks2));
  if (h->R > lim) {
    h->err = "chunk_rows " + std::to_string(h->R) + " exceeds the exact-int32 depth limit " + std::to_string(lim) +
             " of the emulated mode at this digit count";
    return SCFGP_ERR_ARG;
  }
  if (h->Kp + h->K > lim) {
    h->err = "2K = " + std::to_string(2 * h->K) + " features exceed the exact-int32 depth limit " +
             std::to_string(lim) + " of the emulated mode at this digit count";
    return SCFGP_ERR_ARG;
  }
  return SCFGP_OK;
}

// (re)allocate everything that depends on the chunk size R / number of streams
int alloc_chunk_buffers(scfgp_handle* h) {
  for (int i = 0; i < kMaxStreams; ++i) {
    Ctx& c = h->ctx[i];
    for (double** p : {&c.Phi, &c.Psi, &c.Zbar, &c.vpart, &c.sig, &c.mu}) {
      if (*p) cudaFree(*p);
      *p = nullptr;
    }
    if (c.digR) cudaFree(c.digR);
    c.digR = nullptr;
  }
  for (double** p : {&h->acc_all, &h->dW_all, &h->colpart_all, &h->slots_all}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  const int K = h->K, R = h->R;
  const int nkb = cdiv(K, 128);
  int rc;
  // work lists that depend on R
  const int nrb = R / 128;
  std::vector<int4> tg, bw;
  {
    // most expensive first: cost = reduction depth x (fraction of the tile that is valid)
    std::vector<std::pair<double, int>> order;
    for (int kb = 0; kb < nkb; ++kb) {
      const int valid = std::min(128, K - kb * 128);
      const double frac = !h->use_tma ? 1.0 : (valid <= 32 ? 0.25 : (valid <= 64 ? 0.5 : 1.0));
      order.push_back({-frac * std::min(K, (kb + 1) * 128), kb});
    }
    std::sort(order.begin(), order.end());
    for (auto& o : order)
      for (int rb = 0; rb < nrb; ++rb) tg.push_back(make_int4(rb, o.second, 0, 0));
  }
  const int nfb = cdiv(h->J, 64);
  for (int rb = 0; rb < nrb; ++rb)
    for (int fb = 0; fb < nfb; ++fb) bw.push_back(make_int4(rb, fb, 0, 0));
  if ((rc = upload_items(h, tg, &h->tgemm))) return rc;
  if ((rc = upload_items(h, bw, &h->bwd))) return rc;
  const bool emu_on = h->prec == SCFGP_PREC_EMU_I8;
  const int emu_bwd_ctas = emu_on ? nrb * cdiv(h->J, 32) : 0;
  h->vparts = emu_on ? 2 * cdiv(K + 1, umma::TN) : nkb;
  h->slot_stride =
      (std::max<long long>(std::max(std::max(2 * kSeedBlocks, (int)bw.size()), emu_bwd_ctas), 256) + 1) / 2 * 2;
  const size_t KK = (size_t)K * h->ldk, DJ = (size_t)h->D * h->ldj;
  // Emulated mode: the ~300 Gram tiles do not divide evenly over 148 persistent CTAs (2.07 rounds
  // cost 3).  Each tile's row reduction can be cut into gsplit work items with their own partial
  // accumulator (summed in fixed order afterwards); measured at cfg4 (SCFGP_GRAM_SPLIT=1/2/3/4:
  // 0.556 / 0.561 / 0.569 / 0.568 s per 2M rows) the better balance does not pay, because the
  // evaluation is power-limited and idle SMs let the busy ones clock higher -- default 1.
  h->gsplit = emu_on ? kGramSplit : 1;
  if (const char* e = getenv("SCFGP_GRAM_SPLIT")) h->gsplit = emu_on ? std::max(1, std::min(16, atoi(e))) : 1;
  if ((rc = dalloc(h, &h->acc_all, KK * h->ns * h->gsplit))) return rc;
  if ((rc = dalloc(h, &h->dW_all, DJ * kDwSplit * h->ns))) return rc;
  h->csplit = std::max(kColSplit, R / 128);
  if ((rc = dalloc(h, &h->colpart_all, (size_t)h->csplit * K * h->ns))) return rc;
  if ((rc = dalloc(h, &h->slots_all, (size_t)h->slot_stride * h->ns))) return rc;
  for (int i = 0; i < h->ns; ++i) {
    Ctx& c = h->ctx[i];
    if ((rc = dalloc(h, &c.Phi, (size_t)R * h->ldk))) return rc;
    if ((rc = dalloc(h, &c.Psi, (size_t)R * h->ldk))) return rc;
    if ((rc = dalloc(h, &c.Zbar, (size_t)R * h->ldj))) return rc;
    if ((rc = dalloc(h, &c.vpart, (size_t)std::max(nkb, h->vparts) * R))) return rc;
    c.acc = h->acc_all + KK * h->gsplit * i;
    c.dW = h->dW_all + DJ * kDwSplit * i;
    c.colpart = h->colpart_all + (size_t)h->csplit * K * i;
    c.slots = h->slots_all + (size_t)h->slot_stride * i;
    if (h->use_tma) {
      if ((rc = make_tmap(h, &c.mapPhi, c.Phi, K, R, h->ldk))) return rc;
      if ((rc = make_tmap(h, &c.mapPsi, c.Psi, K, R, h->ldk))) return rc;
    }
    if (emu_on) {
      // The leading n digit planes of a digit array are a valid n-digit representation (balanced
      // digits: dropping the tail rounds to nearest), so one array serves both digit counts.
      const long long Kp = h->Kp;
      const int ks = h->ks, ks2 = h->ks2, ksm = std::max(ks, ks2);
      // (dalloc zero-fills: columns [ceil16(K), Kp) and [Kp + ceil16(K), 2 Kp) are never written and stay 0)
      if ((rc = dalloc(h, &c.digR, (size_t)ksm * R * 2 * Kp))) return rc;
      if ((rc = dalloc(h, &c.sig, 2))) return rc;
      if ((rc = dalloc(h, &c.mu, (size_t)R))) return rc;
      if ((rc = make_tmap_digits(h, &c.mapRowsA, c.digR, Kp + K, R, 2 * Kp, ks2, umma::TM))) return rc;
      if ((rc = make_tmap_digits_mn(h, &c.mapPhiMN_A1, c.digR, K, R, 2 * Kp, ks, umma::TM))) return rc;
      if ((rc = make_tmap_digits_mn(h, &c.mapPhiMN_B1, c.digR, K, R, 2 * Kp, ks, umma::TN))) return rc;
      if ((rc = make_tmap_digits_mn(h, &c.mapPhiMN_B, c.digR, K, R, 2 * Kp, ks2, umma::TN))) return rc;
      if ((rc = make_tmap_digits_mn(h, &c.mapPsiMN_A, c.digR + Kp, K, R, 2 * Kp, ks2, umma::TM))) return rc;
      if ((rc = make_tmap_digits_mn(h, &c.mapPhiMN_A, c.digR, K, R, 2 * Kp, ks2, umma::TM))) return rc;
      if ((rc = make_tmap_digits_mn(h, &c.mapPsiMN_B, c.digR + Kp, K, R, 2 * Kp, ks2, umma::TN))) return rc;
    }
  }
  return SCFGP_OK;
}

void free_emu_static(scfgp_handle* h) {
  for (int8_t** p : {&h->digLi, &h->digC}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  for (double** p : {&h->liscale, &h->liinv, &h->cscale, &h->cinv, &h->sigg}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (h->gvmax) cudaFree(h->gvmax);
  h->gvmax = nullptr;
  if (h->gram_items) cudaFree(h->gram_items);
  h->gram_items = nullptr;
  h->gram_count = 0;
}

// replicated digit operands (Li, [2Abar | 2 sigma B]) and the Gram tile list of the emulated path
int alloc_emu_static(scfgp_handle* h) {
  free_emu_static(h);
  if (h->prec != SCFGP_PREC_EMU_I8) return SCFGP_OK;
  const int K = h->K, ks = h->ks2;
  const long long Kp = h->Kp;
  int rc;
  if ((rc = dalloc(h, &h->digLi, (size_t)ks * (K + 1) * Kp))) return rc;   // rows 0..K-1: Li, row K: alpha
  if ((rc = dalloc(h, &h->digC, (size_t)ks * K * 2 * Kp))) return rc;
  for (double** p : {&h->liscale, &h->liinv, &h->cscale, &h->cinv})
    if ((rc = dalloc(h, p, K + 1))) return rc;
  if ((rc = dalloc(h, &h->sigg, 2))) return rc;
  if ((rc = dalloc(h, &h->gvmax, 1))) return rc;
  if ((rc = make_tmap_digits(h, &h->mapLiB, h->digLi, K, K + 1, Kp, ks, umma::TN))) return rc;
  if ((rc = make_tmap_digits(h, &h->mapCB, h->digC, Kp + K, K, 2 * Kp, ks, umma::TN))) return rc;
  // Gram tiles (128 x 64) of the lower triangle.  When the last 128-row block holds at most 64 valid rows
  // (K = 2080: 32), it is computed with the roles swapped -- 17 tiles (all feature blocks x the border's
  // 64-wide block) instead of 33 three-quarters-empty ones -- and stored transposed (item.z = 1); at
  // K = 2080 that also brings the tile count from 306 to 290 <= 2 x 148 persistent CTAs.
  std::vector<int4> v;
  const int nmt = cdiv(K, umma::TM), border = K % umma::TM;
  const bool swap_border = border > 0 && border <= umma::TN && nmt > 1;
  for (int mt = 0; mt < nmt - (swap_border ? 1 : 0); ++mt)
    for (int nt = 0; nt < cdiv(K, umma::TN); ++nt)
      if (nt * umma::TN <= mt * umma::TM + umma::TM - 1) v.push_back(make_int4(mt, nt, 0, 0));
  if (swap_border)
    for (int mt = 0; mt < nmt; ++mt) v.push_back(make_int4(mt, (nmt - 1) * (umma::TM / umma::TN), 1, 0));
  h->gram_count = (int)v.size();
  CK(cudaMalloc((void**)&h->gram_items, v.size() * sizeof(int4)));
  CK(cudaMemcpy(h->gram_items, v.data(), v.size() * sizeof(int4), cudaMemcpyHostToDevice));
  EMU_KS_SWITCH(h->ks, if ((rc = set_emu_attrs<KS>(h))) return rc);
  EMU_KS_SWITCH(h->ks2, if ((rc = set_emu_attrs<KS>(h))) return rc);
  return SCFGP_OK;
}

// fork: the chunk streams start after everything queued on the main stream so far
int fork_streams(scfgp_handle* h) {
  CK(cudaEventRecord(h->ev_fork, h->stream));
  for (int i = 0; i < h->ns; ++i) CK(cudaStreamWaitEvent(h->ctx[i].st, h->ev_fork, 0));
  return SCFGP_OK;
}
// join: the main stream continues after every chunk stream has drained
int join_streams(scfgp_handle* h) {
  for (int i = 0; i < h->ns; ++i) {
    CK(cudaEventRecord(h->ctx[i].done, h->ctx[i].st));
    CK(cudaStreamWaitEvent(h->stream, h->ctx[i].done, 0));
  }
  return SCFGP_OK;
}

// C(lower tiles) += AsT . Bs over the valid rows of two slabs of context c
int run_syrk(scfgp_handle* h, Ctx& c, bool a_is_psi, int rows) {
  const int K = h->K;
  int rc = SCFGP_OK;
  prof_begin(h, SCFGP_FAM_SYRK, c.st);
  if (h->use_tma) {
    // TMA reads whole 32-row stages: rows beyond the valid ones must be zero, not stale
    const int pad = (rows + 31) / 32 * 32;
    if (pad > rows) {
      CK(cudaMemsetAsync(c.Phi + (size_t)rows * h->ldk, 0, (size_t)(pad - rows) * h->ldk * sizeof(double), c.st));
      if (a_is_psi)
        CK(cudaMemsetAsync(c.Psi + (size_t)rows * h->ldk, 0, (size_t)(pad - rows) * h->ldk * sizeof(double), c.st));
    }
    k_syrk_tma<<<h->syrkK.count, TMA_THREADS, TMA_SMEM_BYTES, c.st>>>(a_is_psi ? c.mapPsi : c.mapPhi, c.mapPhi, rows,
                                                                      K, K, c.acc, h->ldk, h->syrkK.d);
    h->launches++;
    cudaError_t e_ = cudaGetLastError();
    if (e_ != cudaSuccess) { h->err = std::string("k_syrk_tma: ") + cudaGetErrorString(e_); rc = SCFGP_ERR_CUDA; }
  } else {
    rc = launch_gemm<CfgBig, false, false, 2, 2>(h, a_is_psi ? c.Psi : c.Phi, h->ldk, K, c.Phi, h->ldk, K, c.acc,
                                                 h->ldk, 1.0, 1.0, h->lowerK, rows, c.st, true);
  }
  prof_end(h, SCFGP_FAM_SYRK, c.st);
  return rc;
}

int run_features(scfgp_handle* h, Ctx& c, const double* Xc, int xvec, int rows, const double* rowscale,
                 double* Psi) {
  const int nfb = cdiv(h->J, 64);
  const int nrb = cdiv(rows, 128);
  const size_t smem = SmemSize<CfgFeat, true, true>::BYTES;
  prof_begin(h, SCFGP_FAM_FEATURES, c.st);
  if (xvec == 2)
    k_features<2><<<nrb * nfb, CfgFeat::NT, smem, c.st>>>(Xc, h->D, rows, h->D, h->Wt, h->ldw, h->phi, h->J,
                                                          h->s, c.Phi, h->ldk, rowscale, Psi, nfb);
  else
    k_features<1><<<nrb * nfb, CfgFeat::NT, smem, c.st>>>(Xc, h->D, rows, h->D, h->Wt, h->ldw, h->phi, h->J,
                                                          h->s, c.Phi, h->ldk, rowscale, Psi, nfb);
  prof_end(h, SCFGP_FAM_FEATURES, c.st);
  CKL();
  return SCFGP_OK;
}

// T = Phi.LiT row sums of squares.  Items are ordered deepest-reduction first; CTAs whose
// row block lies beyond a short last chunk exit immediately.
int run_tgemm(scfgp_handle* h, Ctx& c, int rows) {
  const size_t smem = SmemSize<CfgBig, true, true>::BYTES;
  prof_begin(h, SCFGP_FAM_TGEMM, c.st);
  if (h->use_tma)
    k_tgemm_tma<<<h->tgemm.count, TMA_THREADS, TMA_SMEM_BYTES, c.st>>>(c.mapPhi, h->mapLi, rows, h->K, c.vpart, h->R,
                                                                       h->tgemm.d);
  else
    k_tgemm_rowsq<<<h->tgemm.count, CfgBig::NT, smem, c.st>>>(c.Phi, h->ldk, rows, h->K, h->Li, h->ldk, c.vpart,
                                                              h->R, h->tgemm.d);
  prof_end(h, SCFGP_FAM_TGEMM, c.st);
  CKL();
  return SCFGP_OK;
}

// ---------------------------------------------------------------------------------
// emulated-FP64 path: digit extraction + int8 tensor-core products (emu_kernels.cuh)
// ---------------------------------------------------------------------------------
int emu_digits_rows(scfgp_handle* h, cudaStream_t st, const double* src, long long ld, int rows, int cols,
                    const double* rowmul, const double* rowinv, double ginv, const double* ginv_ptr, int permJ,
                    int8_t* dst, long long pitch, long long slice_stride, int col_off, int nd = 0,
                    const double* rowmul2 = nullptr, const double* ginv_ptr2 = nullptr, int col_off2 = 0) {
  if (nd == 0) nd = h->ks2;
  const long long total = (long long)rows * (cdiv(cols, 16) * 4);
  trace_begin(h, "digits_rows", st);
  EMU_KS_SWITCH(nd, emu::k_digits_rows<KS><<<cdiv(total, 256), 256, 0, st>>>(
                           src, ld, rows, cols, rowmul, rowinv, ginv, ginv_ptr, permJ, dst, pitch, slice_stride, col_off,
                           rowmul2, ginv_ptr2, col_off2));
  trace_end(h, st);
  CKL();
  return SCFGP_OK;
}

// digits of one Phi chunk and their by-products (k_digits_slab): source = the float64 slab of the chunk
// stream, or (from_cache) the digit cache rows r0..; dst1 / dst2 = digit arrays with row pitch `pitch`
int emu_digits_slab(scfgp_handle* h, Ctx& c, int nd, bool from_cache, long long r0, int rows, const double* rowmul1,
                    const double* gptr1, int8_t* dst1, const double* rowmul2, const double* gptr2, int8_t* dst2,
                    long long pitch, long long plane, const double* wsum, double* cpart, bool write_slab = false) {
  const int K = h->K;
  const dim3 grid(cdiv((K + 15) / 16 * 16, 256), cdiv(rows, 128));
  const double recon = h->s2 * std::ldexp(1.0, -(8 * h->cache_ks - 1));
  trace_begin(h, from_cache ? "digits(cache)" : "digits(slab)", c.st);
  if (from_cache) {
    EMU_KS_SWITCH(nd, emu::k_digits_slab<KS, true><<<grid, 256, 0, c.st>>>(
                          nullptr, 0, h->phi_cache + r0 * h->Kc, h->Kc, h->cache_rows * h->Kc, recon, rows, K,
                          1.0 / h->s2, rowmul1, gptr1, dst1, rowmul2, gptr2, dst2, pitch, plane, wsum, cpart,
                          (long long)K, write_slab ? c.Phi : nullptr, h->ldk));
  } else {
    EMU_KS_SWITCH(nd, emu::k_digits_slab<KS, false><<<grid, 256, 0, c.st>>>(
                          c.Phi, h->ldk, nullptr, 0, 0, 0.0, rows, K, 1.0 / h->s2, rowmul1, gptr1, dst1, rowmul2,
                          gptr2, dst2, pitch, plane, wsum, cpart, (long long)K, nullptr, 0));
  }
  trace_end(h, c.st);
  CKL();
  return SCFGP_OK;
}

// The GEMM kernels of a chunk stream run on its low-priority companion stream: enter = the
// companion waits for everything queued on the chunk stream, leave = the chunk stream waits
// for the GEMM.
int gemm_enter(scfgp_handle* h, Ctx& c) {
  CK(cudaEventRecord(c.ev_sg, c.st));
  CK(cudaStreamWaitEvent(c.sg, c.ev_sg, 0));
  return SCFGP_OK;
}
int gemm_leave(scfgp_handle* h, Ctx& c) {
  CK(cudaEventRecord(c.ev_gs, c.sg));
  CK(cudaStreamWaitEvent(c.st, c.ev_gs, 0));
  return SCFGP_OK;
}

// acc(lower tiles) += sc * digits(A)T . digits(Phi) over the rows of one chunk; both operands are
// row-major digit arrays read MN-major (no transposed copies).  Pass 1: A = B = Phi (chunk buffer, or
// the cache rows r0..); pass 2: A = Psi (chunk buffer), B = Phi (chunk buffer or cache).
int emu_gram(scfgp_handle* h, Ctx& c, bool a_is_psi, int rows, long long r0, bool use_cache) {
  const int depth = (rows + umma::KB - 1) / umma::KB * umma::KB;
  // rows per split: a multiple of the stage depth; short chunks use fewer splits
  const int split_len = cdiv(cdiv(depth, h->gsplit), umma::KB) * umma::KB;
  const int nsplit = cdiv(depth, split_len);
  const int nitems = h->gram_count * nsplit;
  const int row0_b = use_cache ? (int)r0 : 0;
  const int row0_a = (use_cache && !a_is_psi) ? (int)r0 : 0;
  int rc;
  if ((rc = gemm_enter(h, c))) return rc;
  prof_begin(h, SCFGP_FAM_SYRK, c.sg);
  // pass 1 (G, feeds the Cholesky): ks digits; pass 2 (H): ks2 leading digits
  EMU_KS_SWITCH(a_is_psi ? h->ks2 : h->ks,
                emu::k_gram_i8<KS><<<std::min(nitems, h->num_sms), umma::THREADS, umma::Geo<KS>::SMEM, c.sg>>>(
                    a_is_psi ? c.mapPsiMN_A : (use_cache ? h->mapCacheMN_A : c.mapPhiMN_A1),
                    use_cache ? h->mapCacheMN_B : (a_is_psi ? c.mapPhiMN_B : c.mapPhiMN_B1),
                    // border tiles (roles swapped): M side = the B-side array in 128-byte boxes, N side = the A-side array
                    use_cache ? h->mapCacheMN_A : (a_is_psi ? c.mapPhiMN_A : c.mapPhiMN_A1),
                    a_is_psi ? c.mapPsiMN_B : (use_cache ? h->mapCacheMN_B : c.mapPhiMN_B1), row0_a, row0_b, depth,
                    split_len, h->K, h->gram_items, h->gram_count, nitems, c.acc, h->ldk, (long long)h->K * h->ldk,
                    h->s2 * h->s2, a_is_psi ? c.sig : nullptr));
  prof_end(h, SCFGP_FAM_SYRK, c.sg);
  CKL();
  return gemm_leave(h, c);
}

int emu_tgemm(scfgp_handle* h, Ctx& c, int rows, long long r0, bool use_cache) {
  const int ntiles = cdiv(h->K + 1, umma::TN) * cdiv(rows, umma::TM);
  int rc;
  if ((rc = gemm_enter(h, c))) return rc;
  prof_begin(h, SCFGP_FAM_TGEMM, c.sg);
  EMU_KS_SWITCH(h->ks2, emu::k_tgemm_i8<KS><<<std::min(ntiles, h->num_sms), umma::THREADS, umma::Geo<KS>::SMEM, c.sg>>>(
                           use_cache ? h->mapCacheK : c.mapRowsA, h->mapLiB, use_cache ? (int)r0 : 0, rows, h->K, ntiles,
                           h->liscale, h->s2, c.vpart, h->R, c.mu));
  prof_end(h, SCFGP_FAM_TGEMM, c.sg);
  CKL();
  return gemm_leave(h, c);
}

int emu_backward(scfgp_handle* h, Ctx& c, int rows, long long r0, bool use_cache) {
  const int ntiles = cdiv(h->J, 32) * cdiv(rows, umma::TM);
  // the epilogue rebuilds Phi from the cached digits (8-byte loads: needs J % 8 == 0), else reads the slab
  const bool phi_from_digits = use_cache && h->J % 8 == 0;
  const double recon = h->s2 * std::ldexp(1.0, -(8 * std::max(h->cache_ks, 1) - 1));
  int rc;
  if ((rc = gemm_enter(h, c))) return rc;
  prof_begin(h, SCFGP_FAM_BACKWARD, c.sg);
  EMU_KS_SWITCH(h->ks2, emu::k_backward_i8<KS><<<std::min(ntiles, h->num_sms), umma::THREADS, umma::Geo<KS>::SMEM, c.sg>>>(
                           c.mapRowsA, use_cache ? h->mapCacheK : c.mapRowsA, h->mapCB,
                           use_cache ? (int)(h->Kp / umma::KB) : 0, use_cache ? (int)r0 : 0,
                           phi_from_digits ? h->phi_cache : nullptr, h->Kc, h->cache_rows * h->Kc, recon, rows, h->J,
                           (int)(h->Kp + h->K), ntiles, h->cscale, h->s2, c.Phi, h->ldk, h->dy + r0, h->gmu + r0,
                           h->pbar, h->alpha, c.Zbar, h->ldj, c.slots, c.colpart, (long long)h->K));
  prof_end(h, SCFGP_FAM_BACKWARD, c.sg);
  CKL();
  return gemm_leave(h, c);
}

// (re)allocate the digit cache for the current data set if it is wanted and fits; otherwise run without
void free_phi_cache(scfgp_handle* h) {
  if (h->phi_cache) cudaFree(h->phi_cache);
  h->phi_cache = nullptr;
  h->cache_rows = 0;
  h->cache_ks = 0;
}
int ensure_phi_cache(scfgp_handle* h) {
  const bool want = h->prec == SCFGP_PREC_EMU_I8 && h->want_cache && h->ks == h->ks2 && h->N > 0;
  const long long rows = (h->N + 127) / 128 * 128;
  if (want && h->phi_cache && h->cache_rows >= rows && h->cache_ks == h->ks) return SCFGP_OK;
  free_phi_cache(h);
  if (!want) return SCFGP_OK;
  h->Kc = (h->K + 15) / 16 * 16;
  const size_t bytes = (size_t)h->ks * rows * h->Kc;
  size_t fr = 0, tot = 0;
  CK(cudaMemGetInfo(&fr, &tot));
  if (bytes + ((size_t)6 << 30) > fr) return SCFGP_OK;   // keep 6 GB of head room for the caller
  if (cudaMalloc((void**)&h->phi_cache, bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    h->phi_cache = nullptr;
    return SCFGP_OK;
  }
  h->cache_rows = rows;
  h->cache_ks = h->ks;
  int rc;
  if ((rc = make_tmap_digits(h, &h->mapCacheK, h->phi_cache, h->K, rows, h->Kc, h->ks, umma::TM))) return rc;
  if ((rc = make_tmap_digits_mn(h, &h->mapCacheMN_A, h->phi_cache, h->K, rows, h->Kc, h->ks, umma::TM))) return rc;
  if ((rc = make_tmap_digits_mn(h, &h->mapCacheMN_B, h->phi_cache, h->K, rows, h->Kc, h->ks, umma::TN))) return rc;
  return SCFGP_OK;
}

// digits of Li with per-row power-of-two scales (B operand of Phi.LiT)
int emu_prepare_li(scfgp_handle* h) {
  const int K = h->K;
  emu::k_rowscale<<<K, 128, 0, h->stream>>>(h->Li, h->ldk, nullptr, 0, nullptr, K, K, 0, h->liscale, h->liinv);
  CKL();
  const long long plane = (long long)(K + 1) * h->Kp;
  int rc;
  if ((rc = emu_digits_rows(h, h->stream, h->Li, h->ldk, K, K, nullptr, h->liinv, 1.0, nullptr, 0, h->digLi, h->Kp,
                            plane, 0)))
    return rc;
  // row K = alpha (its own power-of-two scale): the product's column K is mu = Phi.alpha
  emu::k_rowscale<<<1, 128, 0, h->stream>>>(h->alpha, K, nullptr, 0, nullptr, 1, K, 0, h->liscale + K, h->liinv + K);
  CKL();
  return emu_digits_rows(h, h->stream, h->alpha, K, 1, K, nullptr, h->liinv + K, 1.0, nullptr, 0,
                         h->digLi + (long long)K * h->Kp, h->Kp, plane, 0);
}

// digits of Cmat = [2 Abar | 2 sigma B] (rows permuted cos/sin-interleaved, per-row scales);
// sigma = power-of-two bound of |gv| over this rank's rows (pass 2)
int emu_prepare_c(scfgp_handle* h) {
  const int K = h->K;
  int rc;
  emu::k_scale_from_bits<<<1, 1, 0, h->stream>>>(h->gvmax, h->sigg);
  CKL();
  emu::k_rowscale<<<K, 128, 0, h->stream>>>(h->A2, h->ldk, h->B2, h->ldk, h->sigg, K, K, h->J, h->cscale, h->cinv);
  CKL();
  if ((rc = emu_digits_rows(h, h->stream, h->A2, h->ldk, K, K, nullptr, h->cinv, 1.0, nullptr, h->J, h->digC,
                            2 * h->Kp, (long long)K * 2 * h->Kp, 0)))
    return rc;
  return emu_digits_rows(h, h->stream, h->B2, h->ldk, K, K, nullptr, h->cinv, 1.0, h->sigg, h->J, h->digC, 2 * h->Kp,
                         (long long)K * 2 * h->Kp, (int)h->Kp);
}

// lower Cholesky of h->L in place + Li = inv(L)                SCFGP/SCFGP.py:106-107
int factor_device(scfgp_handle* h, double* L, double* Li, int K, long long ld) {
  int rc;
  const int nblk = cdiv(K, NB);
  for (int bi = 0; bi < nblk; ++bi) {
    const int k0 = bi * NB, nb = std::min(NB, K - k0);
    double* Dk = h->Dinv + (size_t)bi * NB * NB;
    k_potrf_diag<<<1, 256, POTRF_SMEM, h->stream>>>(L, ld, k0, nb, Dk, h->info);
    CKL();
    const int m = K - k0 - nb;
    if (m > 0) {
      double* panel = L + (size_t)(k0 + nb) * ld + k0;
      // L21 = A21 . inv(L11)^T
      if ((rc = launch_gemm<CfgTall, true, true, 2, 2>(h, panel, ld, m, Dk, NB, nb, panel, ld, 1.0, 0.0,
                                                       h->potrf_panel[bi], nb)))
        return rc;
      // A22 -= L21 . L21^T (lower tiles)
      double* trail = L + (size_t)(k0 + nb) * ld + (k0 + nb);
      if ((rc = launch_gemm<CfgBig, true, true, 2, 2>(h, panel, ld, m, panel, ld, m, trail, ld, -1.0, 1.0,
                                                      h->potrf_trail[bi], nb)))
        return rc;
    }
  }
  // Li = inv(L), block-recursively:  inv([L11 0; L21 L22]) = [Li11 0; -Li22 (L21 Li11)  Li22].  The diagonal
  // 64-blocks are the Dinv of the factorisation; every level of the recursion is two launches that do all
  // merges of that level at once (tile lists with per-tile reduction ranges: the triangular zeros are skipped).
  CK(cudaMemsetAsync(Li, 0, (size_t)K * ld * sizeof(double), h->stream));
  k_place_diag_blocks<<<nblk, 256, 0, h->stream>>>(h->Dinv, Li, ld, K);
  CKL();
  for (size_t lev = 0; lev < h->inv_t.size(); ++lev) {
    // T (in the scratch matrix BH, same position as the block it will become) = L21 . Li11
    if ((rc = launch_gemm<CfgSq, true, false, 2, 2>(h, L, ld, K, Li, ld, K, h->BH, ld, 1.0, 0.0, h->inv_t[lev], K)))
      return rc;
    // Li21 = -Li22 . T
    if ((rc = launch_gemm<CfgSq, true, false, 2, 2>(h, Li, ld, K, h->BH, ld, K, Li, ld, -1.0, 0.0, h->inv_x[lev], K)))
      return rc;
  }
  return SCFGP_OK;
}

int build_static_lists(scfgp_handle* h) {
  const int K = h->K;
  int rc;
  if ((rc = upload_items(h, lower_tiles(K, 128, 0, -1), &h->lowerK))) return rc;
  {
    // TMA Gram kernel: full lower tiles first, then the K mod 128 border row block with the
    // operand roles swapped (narrow side = n side) and a transposed store
    std::vector<int4> v;
    const int B = K / 128, border = K % 128;
    for (int i = 0; i < B; ++i)
      for (int j = 0; j <= i; ++j) v.push_back(make_int4(i, j, 0, 0));
    if (border)
      for (int j = 0; j <= B; ++j) v.push_back(make_int4(j, B, 1, 0));
    if ((rc = upload_items(h, v, &h->syrkK))) return rc;
  }
  {
    std::vector<int4> v;
    int nb = cdiv(K, 128);
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j < nb; ++j) v.push_back(make_int4(i, j, 0, -1));
    if ((rc = upload_items(h, v, &h->fullK))) return rc;
  }
  {
    std::vector<int4> v;
    for (int mb = 0; mb < cdiv(h->D, 64); ++mb)
      for (int nb = 0; nb < cdiv(h->J, 128); ++nb) v.push_back(make_int4(mb, nb, 0, -1));
    if ((rc = upload_items(h, v, &h->dw))) return rc;
  }
  const int nblk = cdiv(K, NB);
  h->potrf_panel.resize(nblk);
  h->potrf_trail.resize(nblk);
  for (int bi = 0; bi < nblk; ++bi) {
    const int k0 = bi * NB, nb = std::min(NB, K - k0);
    const int m = K - k0 - nb;
    std::vector<int4> pv;
    for (int i = 0; i < cdiv(m, 128); ++i) pv.push_back(make_int4(i, 0, 0, -1));
    if ((rc = upload_items(h, pv, &h->potrf_panel[bi]))) return rc;
    if ((rc = upload_items(h, lower_tiles(m, 128, 0, -1), &h->potrf_trail[bi]))) return rc;
  }
  // block-recursive inverse: merges grouped by level (height of the recursion tree above the leaves)
  {
    std::vector<std::vector<int4>> lt, lx;
    // returns the level of the merged range [lo, hi) (in 64-blocks); leaves have level -1
    std::function<int(int, int)> rec = [&](int lo, int hi) -> int {
      if (hi - lo <= 1) return -1;
      const int mid = (lo + hi + 1) / 2;
      const int lev = std::max(rec(lo, mid), rec(mid, hi)) + 1;
      if ((int)lt.size() <= lev) { lt.resize(lev + 1); lx.resize(lev + 1); }
      for (int i = mid; i < hi; ++i)
        for (int j = lo; j < mid; ++j) {
          // T[i][j] = sum_{k in [j, mid)} L[i][k] Li[k][j]        (Li11 lower triangular: k >= j)
          lt[lev].push_back(make_int4(i, j, j * NB, std::min(mid * NB, K) - j * NB));
          // Li[i][j] = - sum_{k in [mid, i]} Li[i][k] T[k][j]     (Li22 lower triangular: k <= i)
          lx[lev].push_back(make_int4(i, j, mid * NB, std::min((i + 1) * NB, K) - mid * NB));
        }
      return lev;
    };
    rec(0, nblk);
    h->inv_t.resize(lt.size());
    h->inv_x.resize(lx.size());
    for (size_t l = 0; l < lt.size(); ++l) {
      if ((rc = upload_items(h, lt[l], &h->inv_t[l]))) return rc;
      if ((rc = upload_items(h, lx[l], &h->inv_x[l]))) return rc;
    }
  }
  return SCFGP_OK;
}

void free_all(scfgp_handle* h) {
  for (int i = 0; i < kMaxStreams; ++i) {
    Ctx& c = h->ctx[i];
    for (double* p : {c.Phi, c.Psi, c.Zbar, c.vpart, c.sig, c.mu})
      if (p) cudaFree(p);
    if (c.digR) cudaFree(c.digR);
    if (c.done) cudaEventDestroy(c.done);
    if (c.ev_sg) cudaEventDestroy(c.ev_sg);
    if (c.ev_gs) cudaEventDestroy(c.ev_gs);
    if (c.st) cudaStreamDestroy(c.st);
    if (c.sg) cudaStreamDestroy(c.sg);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  free_phi_cache(h);
  if (h->cstream) cudaStreamDestroy(h->cstream);
  for (auto& e : h->up_ev) cudaEventDestroy(e);
  free_emu_static(h);
  for (double* p : {h->acc_all, h->dW_all, h->colpart_all, h->slots_all, h->Wt, h->phi, h->red1, h->red2, h->red3, h->L,
                    h->Li, h->Bm, h->BH, h->BHB, h->A2, h->B2, h->Dinv, h->tmpS, h->alpha, h->beta, h->tvec,
                    h->pbar, h->scal, h->gmu, h->gv, h->colpart, h->slots, h->trslots, h->ownX, h->owny})
    if (p) cudaFree(p);
  if (h->info) cudaFree(h->info);
  for (int4* p : h->to_free) cudaFree(p);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->prof_ev) cudaEventDestroy(e);
  for (auto& r : h->trace) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
}

// penalty of SCFGP/SCFGP.py:94,114-117,127 and its adjoints (A.6), host
double penalty_and_grad(const scfgp_handle* h, std::vector<double>& dlF, std::vector<double>& dF) {
  const int D = h->D, S = h->S, M = h->M;
  auto stats = [&](const std::vector<double>& Q, int n, std::vector<double>& mean, std::vector<double>& sd,
                   double& mu, double& sig) {
    mean.assign(D, 0.0); sd.assign(D, 0.0); mu = 0; sig = 0;
    for (int d = 0; d < D; ++d) {
      double m = 0;
      for (int j = 0; j < n; ++j) m += Q[(size_t)d * n + j];
      m /= n;
      double v = 0;
      for (int j = 0; j < n; ++j) { double t = Q[(size_t)d * n + j] - m; v += t * t; }
      mean[d] = m; sd[d] = std::sqrt(v / n);
      mu += m; sig += sd[d];
    }
  };
  std::vector<double> ml, sl, mw, sw;
  double mu_l, sig_l, mu_w, sig_w;
  stats(h->lF, S, ml, sl, mu_l, sig_l);
  stats(h->F, M, mw, sw, mu_w, sig_w);
  auto kl = [](double mu, double sig) { return sig + mu * mu - std::log(sig); };
  const double pen = (kl(mu_w, sig_w) * M + kl(mu_l, sig_l) * S) / (S + M);
  auto adj = [&](const std::vector<double>& Q, int n, const std::vector<double>& mean,
                 const std::vector<double>& sd, double mu, double sig, double weight, std::vector<double>& out) {
    out.assign((size_t)D * n, 0.0);
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < n; ++j)
        out[(size_t)d * n + j] =
            weight * (2 * mu / n + (1 - 1 / sig) * (Q[(size_t)d * n + j] - mean[d]) / (n * sd[d]));
  };
  adj(h->lF, S, ml, sl, mu_l, sig_l, (double)S / (S + M), dlF);
  adj(h->F, M, mw, sw, mu_w, sig_w, (double)M / (S + M), dF);
  return pen;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int scfgp_abi_version(void) { return 1; }

int scfgp_create(int D, int S, int M, int fftype, int device, scfgp_handle** out) {
  if (!out) return SCFGP_ERR_ARG;
  *out = nullptr;
  if (D <= 0 || S <= 0 || M <= 0 || (fftype != SCFGP_FF_PHASE && fftype != SCFGP_FF_NOPHASE))
    return SCFGP_ERR_ARG;
  scfgp_handle* h = new scfgp_handle();
  *out = h;
  h->D = D; h->S = S; h->M = M; h->J = S + M; h->K = 2 * (S + M); h->fftype = fftype; h->device = device;
  h->P = 3 + (long long)D * S + (long long)M * S + S + M;
  h->ldk = h->K;                       // K = 2J is even: 16-byte row alignment holds
  h->ldj = (h->J + 1) / 2 * 2;
  h->ldw = (D + 1) / 2 * 2;
  h->Kp = (h->K + 63) / 64 * 64;   // stage-aligned: the Psi half of [Phi | Psi] starts on a pipeline-stage boundary
  if (const char* e = getenv("SCFGP_PHI_CACHE")) h->want_cache = atoi(e) != 0;
  if (const char* e = getenv("SCFGP_PRECISION")) {
    if (!strncmp(e, "emu", 3)) {
      h->prec = SCFGP_PREC_EMU_I8;
      const int d = atoi(e + 3);
      h->ks = (d >= 4 && d <= 8) ? d : 5;
      h->ks2 = h->ks;
      if (const char* sl = strchr(e, '/')) {
        const int d2 = atoi(sl + 1);
        if (d2 >= 4 && d2 <= 8) h->ks2 = d2;
      }
      h->R = 8192;
    }
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { h->err = "no such CUDA device"; return SCFGP_ERR_CUDA; }
  CK(cudaSetDevice(device));
  CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  int rc;
  if ((rc = set_kernel_attrs(h))) return rc;
  const size_t K = h->K, KK = K * h->ldk;
  if ((rc = dalloc(h, &h->Wt, (size_t)h->J * h->ldw))) return rc;
  if ((rc = dalloc(h, &h->phi, h->J))) return rc;
  if ((rc = dalloc(h, &h->red1, KK + K + 1))) return rc;
  if ((rc = dalloc(h, &h->red2, KK + K + 2))) return rc;
  if ((rc = dalloc(h, &h->red3, (size_t)D * h->ldj + h->ldj + 1))) return rc;
  for (double** p : {&h->L, &h->Li, &h->Bm, &h->BH, &h->BHB, &h->A2, &h->B2})
    if ((rc = dalloc(h, p, KK))) return rc;
  if ((rc = dalloc(h, &h->tmpS, (size_t)NB * h->ldk))) return rc;
  if ((rc = dalloc(h, &h->Dinv, (size_t)cdiv(K, NB) * NB * NB))) return rc;
  for (double** p : {&h->alpha, &h->beta, &h->tvec, &h->pbar})
    if ((rc = dalloc(h, p, K))) return rc;
  if ((rc = dalloc(h, &h->scal, 64))) return rc;
  if ((rc = dalloc(h, &h->colpart, (size_t)kColSplit * K))) return rc;
  if ((rc = dalloc(h, &h->slots, 1024))) return rc;
  if (const char* e = getenv("SCFGP_STREAMS")) h->ns = std::min(kMaxStreams, std::max(1, atoi(e)));
  // The block scheduler hands out the CTAs of ready kernels in priority order, so the long
  // tensor-core GEMM kernels go to low-priority streams: the short memory-bound kernels of the
  // other chunk streams are then placed as soon as they are ready, next to the resident GEMM
  // CTAs (which leave registers and shared memory free), instead of queueing behind them.
  int prio_lo = 0, prio_hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  for (int i = 0; i < kMaxStreams; ++i) {
    CK(cudaStreamCreateWithPriority(&h->ctx[i].st, cudaStreamNonBlocking, prio_hi));
    CK(cudaStreamCreateWithPriority(&h->ctx[i].sg, cudaStreamNonBlocking, prio_lo));
    CK(cudaEventCreateWithFlags(&h->ctx[i].done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ctx[i].ev_sg, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ctx[i].ev_gs, cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking));
  if ((rc = dalloc(h, &h->trslots, K))) return rc;
  if ((rc = dalloc(h, &h->info, 1))) return rc;
  if (const char* e = getenv("SCFGP_NO_TMA")) h->use_tma = atoi(e) ? 0 : 1;
  if (h->use_tma) {
    if ((rc = make_tmap(h, &h->mapLi, h->Li, K, K, h->ldk))) return rc;
    if ((rc = make_tmap(h, &h->mapA2, h->A2, K, K, h->ldk))) return rc;
    if ((rc = make_tmap(h, &h->mapB2, h->B2, K, K, h->ldk))) return rc;
  }
  if ((rc = build_static_lists(h))) return rc;
  if ((rc = emu_check_depths(h))) return rc;
  if ((rc = alloc_emu_static(h))) return rc;
  if ((rc = alloc_chunk_buffers(h))) return rc;
  for (auto& e : h->ev) CK(cudaEventCreate(&e));
  CK(cudaStreamSynchronize(h->stream));
  return SCFGP_OK;
}

int scfgp_destroy(scfgp_handle* h) {
  if (!h) return SCFGP_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  free_all(h);
  delete h;
  return SCFGP_OK;
}

const char* scfgp_last_error(const scfgp_handle* h) { return h ? h->err.c_str() : "null handle"; }
long long scfgp_last_pivot(const scfgp_handle* h) { return h ? h->pivot : -1; }
long long scfgp_num_params(const scfgp_handle* h) { return h ? h->P : 0; }
long long scfgp_num_features(const scfgp_handle* h) { return h ? h->K : 0; }
long long scfgp_launch_count(const scfgp_handle* h) { return h ? h->launches : 0; }

int scfgp_set_stream(scfgp_handle* h, void* cuda_stream) {
  if (!h) return SCFGP_ERR_ARG;
  h->stream = (cudaStream_t)cuda_stream;
  return SCFGP_OK;
}

int scfgp_set_chunk_rows(scfgp_handle* h, int rows) {
  if (!h || rows < 128 || rows % 128 != 0) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  h->upload_pending = false;  // (the device-wide sync above completed any upload)
  const int old_R = h->R;
  h->R = rows;
  int rc;
  if ((rc = emu_check_depths(h))) { h->R = old_R; return rc; }
  h->R_user = true;
  return alloc_chunk_buffers(h);
}

int scfgp_set_precision(scfgp_handle* h, int mode, int digits) { return scfgp_set_precision2(h, mode, digits, digits); }

int scfgp_set_precision2(scfgp_handle* h, int mode, int digits, int digits2) {
  if (!h || (mode != SCFGP_PREC_FP64 && mode != SCFGP_PREC_EMU_I8)) return SCFGP_ERR_ARG;
  if (digits == 0) digits = 5;
  if (digits2 == 0) digits2 = digits;
  if (digits < 4 || digits > 8 || digits2 < 4 || digits2 > 8) { h->err = "digits must be 4..8"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  h->upload_pending = false;  // (the device-wide sync above completed any upload)
  const int o_prec = h->prec, o_ks = h->ks, o_ks2 = h->ks2, o_R = h->R;
  h->prec = mode;
  h->ks = digits;
  h->ks2 = digits2;
  if (!h->R_user) h->R = mode == SCFGP_PREC_EMU_I8 ? 8192 : 4096;
  int rc;
  if ((rc = emu_check_depths(h))) { h->prec = o_prec; h->ks = o_ks; h->ks2 = o_ks2; h->R = o_R; return rc; }
  h->have_factors = false;
  free_phi_cache(h);   // rebuilt (if it fits) by the next evaluation
  if ((rc = alloc_emu_static(h))) return rc;
  return alloc_chunk_buffers(h);
}

int scfgp_get_precision(const scfgp_handle* h, int* mode, int* digits) {
  if (!h) return SCFGP_ERR_ARG;
  if (mode) *mode = h->prec;
  if (digits) *digits = h->ks;
  return SCFGP_OK;
}

int scfgp_get_precision2(const scfgp_handle* h, int* mode, int* digits, int* digits2) {
  if (!h) return SCFGP_ERR_ARG;
  if (mode) *mode = h->prec;
  if (digits) *digits = h->ks;
  if (digits2) *digits2 = h->ks2;
  return SCFGP_OK;
}

int scfgp_set_streams(scfgp_handle* h, int n) {
  if (!h || n < 1 || n > kMaxStreams) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  h->upload_pending = false;  // (the device-wide sync above completed any upload)
  h->ns = n;
  return alloc_chunk_buffers(h);
}

static int set_data_common(scfgp_handle* h, long long N, long long n_global) {
  h->N = N;
  h->Nglobal = n_global > 0 ? n_global : N;
  h->have_factors = false;
  h->xvec = (((uintptr_t)h->dX % 16) == 0 && h->D % 2 == 0) ? 2 : 1;
  return ensure_row_arrays(h, N);
}

// wait (on the host) for an upload started by scfgp_set_data_host_async
static int finish_upload(scfgp_handle* h) {
  if (h->upload_pending) {
    CK(cudaStreamSynchronize(h->cstream));
    h->upload_pending = false;
  }
  return SCFGP_OK;
}

int scfgp_set_data_host_async(scfgp_handle* h, const double* X, const double* y, long long N, long long n_global) {
  if (!h || !X || !y || N <= 0) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc;
  if ((rc = finish_upload(h))) return rc;
  if (N > h->own_cap) {
    CK(cudaDeviceSynchronize());
    if (h->ownX) cudaFree(h->ownX);
    if (h->owny) cudaFree(h->owny);
    h->ownX = h->owny = nullptr;
    h->own_cap = 0;
    CK(cudaMalloc((void**)&h->ownX, (size_t)N * h->D * sizeof(double)));
    CK(cudaMalloc((void**)&h->owny, (size_t)N * sizeof(double)));
    h->own_cap = N;
  } else {
    // the previous evaluation may still be reading the buffers
    CK(cudaStreamSynchronize(h->stream));
  }
  // y first, then X in blocks of a few chunks; one event per block
  CK(cudaMemcpyAsync(h->owny, y, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, h->cstream));
  h->up_block = (long long)h->R * 8;
  const long long nblk = (N + h->up_block - 1) / h->up_block;
  while ((long long)h->up_ev.size() < nblk) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->up_ev.push_back(e);
  }
  for (long long b = 0; b < nblk; ++b) {
    const long long r0 = b * h->up_block, rows = std::min(h->up_block, N - r0);
    CK(cudaMemcpyAsync(h->ownX + r0 * h->D, X + r0 * h->D, (size_t)rows * h->D * sizeof(double),
                       cudaMemcpyHostToDevice, h->cstream));
    CK(cudaEventRecord(h->up_ev[b], h->cstream));
  }
  h->upload_pending = true;
  h->dX = h->ownX;
  h->dy = h->owny;
  return set_data_common(h, N, n_global);
}

int scfgp_set_data_host(scfgp_handle* h, const double* X, const double* y, long long N, long long n_global) {
  int rc = scfgp_set_data_host_async(h, X, y, N, n_global);
  if (rc) return rc;
  return finish_upload(h);
}

int scfgp_set_data_device(scfgp_handle* h, const double* dX, const double* dy, long long N, long long n_global) {
  if (!h || !dX || !dy || N <= 0) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rcu;
  if ((rcu = finish_upload(h))) return rcu;
  h->dX = dX;
  h->dy = dy;
  return set_data_common(h, N, n_global);
}

// ---------------------------------------------------------------------------------
int scfgp_eval_begin(scfgp_handle* h, const double* params, int want_grad) {
  if (!h || !params) return SCFGP_ERR_ARG;
  if (!h->dX || h->N <= 0) { h->err = "no training data set"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  h->want_grad = want_grad;
  h->have_factors = false;
  h->pivot = -1;
  int rcc;
  if ((rcc = ensure_phi_cache(h))) return rcc;
  return unpack_and_upload(h, params);
}

int scfgp_eval_pass1(scfgp_handle* h) {
  if (!h) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const int K = h->K;
  const size_t KK = (size_t)K * h->ldk;
  int rc;
  CK(cudaEventRecord(h->ev[0], h->stream));
  double* G = h->red1;
  double* p = h->red1 + KK;
  double* yy = p + K;
  if ((rc = fork_streams(h))) return rc;
  for (int i = 0; i < h->ns; ++i) {
    Ctx& c = h->ctx[i];
    CK(cudaMemsetAsync(c.acc, 0, KK * h->gsplit * sizeof(double), c.st));
    CK(cudaMemsetAsync(c.colpart, 0, (size_t)h->csplit * K * sizeof(double), c.st));
  }
  long long ci = 0;
  for (long long r0 = 0; r0 < h->N; r0 += h->R, ++ci) {
    Ctx& c = h->ctx[ci % h->ns];
    const int rows = (int)std::min<long long>(h->R, h->N - r0);
    if (h->upload_pending) {
      // rows of this chunk arrive with upload block b (blocks complete in order on the copy stream)
      const long long b = (r0 + rows - 1) / h->up_block;
      CK(cudaStreamWaitEvent(c.st, h->up_ev[b], 0));
    }
    if ((rc = run_features(h, c, h->dX + r0 * h->D, h->xvec, rows, nullptr, nullptr))) return rc;
    if (h->prec == SCFGP_PREC_EMU_I8) {
      // row digits of Phi -- into the cache (all rows stay resident for passes 2 and 3) or the chunk
      // buffer; p = PhiT.y comes out of the same pass over the slab; the Gram reads them MN-major
      const bool cached = h->phi_cache != nullptr;
      int8_t* dst = cached ? h->phi_cache + r0 * h->Kc : c.digR;
      const long long pitch = cached ? h->Kc : 2 * h->Kp;
      const long long plane = cached ? h->cache_rows * h->Kc : (long long)h->R * 2 * h->Kp;
      if ((rc = emu_digits_slab(h, c, h->ks, false, r0, rows, nullptr, nullptr, dst, nullptr, nullptr, nullptr,
                                pitch, plane, h->dy + r0, c.colpart)))
        return rc;
      if ((rc = emu_gram(h, c, false, rows, r0, cached))) return rc;
    } else {
      if ((rc = run_syrk(h, c, false, rows))) return rc;
      trace_begin(h, "colsum", c.st);
      k_colsum<<<dim3(cdiv(K, 256), kColSplit), 256, 0, c.st>>>(c.Phi, h->ldk, rows, K, h->dy + r0, c.colpart, K);
      trace_end(h, c.st);
      CKL();
    }
  }
  // yy on the main stream while the chunk streams drain
  if (h->upload_pending) CK(cudaStreamWaitEvent(h->stream, h->up_ev[0], 0));  // y is copied before X block 0
  CK(cudaMemsetAsync(h->slots, 0, 256 * sizeof(double), h->stream));
  k_sumsq_part<<<256, 256, 0, h->stream>>>(h->dy, h->N, h->slots);
  CKL();
  k_reduce_slots<<<1, 256, 0, h->stream>>>(h->slots, 256, 1, 1, yy);
  CKL();
  if ((rc = join_streams(h))) return rc;
  h->upload_pending = false;  // everything queued after this point is ordered behind the whole upload
  k_sum_parts<<<cdiv((long long)KK, 256), 256, 0, h->stream>>>(h->acc_all, (long long)KK, h->ns * h->gsplit, (int)KK, G);
  CKL();
  k_sum_parts<<<cdiv(K, 256), 256, 0, h->stream>>>(h->colpart_all, K, h->csplit * h->ns, K, p);
  CKL();
  CK(cudaEventRecord(h->ev[1], h->stream));
  return SCFGP_OK;
}

int scfgp_eval_factor(scfgp_handle* h) {
  if (!h) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const int K = h->K;
  const size_t KK = (size_t)K * h->ldk;
  int rc;
  CK(cudaEventRecord(h->ev[6], h->stream));
  double* G = h->red1;
  double* p = h->red1 + KK;
  CK(cudaMemsetAsync(h->info, 0, sizeof(long long), h->stream));
  CK(cudaMemcpyAsync(h->L, G, KK * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  k_add_diag<<<cdiv(K, 256), 256, 0, h->stream>>>(h->L, h->ldk, K, h->sig2n + kEps);
  CKL();
  if ((rc = factor_device(h, h->L, h->Li, K, h->ldk))) return rc;
  // beta = Li p ; alpha = LiT beta                           SCFGP/SCFGP.py:108-110
  k_gemv_n<<<cdiv(K, 8), 256, 0, h->stream>>>(h->Li, h->ldk, K, K, p, h->beta);
  CKL();
  CK(cudaMemsetAsync(h->colpart, 0, (size_t)kColSplit * K * sizeof(double), h->stream));
  k_colsum<<<dim3(cdiv(K, 256), kColSplit), 256, 0, h->stream>>>(h->Li, h->ldk, K, K, h->beta, h->colpart, K);
  CKL();
  k_sum_parts<<<cdiv(K, 256), 256, 0, h->stream>>>(h->colpart, K, kColSplit, K, h->alpha);
  CKL();
  k_factor_scalars<<<1, 256, 0, h->stream>>>(h->L, h->ldk, K, h->beta, p, h->alpha, h->scal);
  CKL();
  if (h->prec == SCFGP_PREC_EMU_I8 && (rc = emu_prepare_li(h))) return rc;   // digits of [Li; alpha]
  if (h->want_grad) {
    // B = LiT Li = A^-1 (lower tiles; rows r < m0 of Li[:, m-block] are zero), then mirror
    if ((rc = launch_gemm<CfgBig, false, false, 2, 2>(h, h->Li, h->ldk, K, h->Li, h->ldk, K, h->Bm, h->ldk,
                                                      1.0, 0.0, h->lowerK, K)))
      return rc;
    k_symmetrize<<<dim3(cdiv(K, 32), cdiv(K, 32)), dim3(32, 32), 0, h->stream>>>(h->Bm, h->ldk, K);
    CKL();
  }
  long long info = 0;
  CK(cudaMemcpyAsync(&info, h->info, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(h->ev[7], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (info != 0) {
    h->pivot = info - 1;
    h->err = "matrix is not positive definite (leading minor " + std::to_string(info) + ")";
    return SCFGP_ERR_NOT_PD;
  }
  h->have_factors = true;
  return SCFGP_OK;
}

int scfgp_eval_pass2(scfgp_handle* h) {
  if (!h) return SCFGP_ERR_ARG;
  if (!h->have_factors) { h->err = "pass2 before factor"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  const int K = h->K;
  const size_t KK = (size_t)K * h->ldk;
  int rc;
  CK(cudaEventRecord(h->ev[2], h->stream));
  double* H = h->red2;
  double* q = h->red2 + KK;
  double* Eg = q + K;
  const int nkb = h->vparts;
  const bool emu_on = h->prec == SCFGP_PREC_EMU_I8;
  if (emu_on) CK(cudaMemsetAsync(h->gvmax, 0, sizeof(unsigned long long), h->stream));
  if ((rc = fork_streams(h))) return rc;
  for (int i = 0; i < h->ns; ++i) {
    Ctx& c = h->ctx[i];
    if (h->want_grad) {
      CK(cudaMemsetAsync(c.acc, 0, KK * h->gsplit * sizeof(double), c.st));
      CK(cudaMemsetAsync(c.colpart, 0, (size_t)h->csplit * K * sizeof(double), c.st));
    }
    CK(cudaMemsetAsync(c.slots, 0, (size_t)h->slot_stride * sizeof(double), c.st));
  }
  long long ci = 0;
  for (long long r0 = 0; r0 < h->N; r0 += h->R, ++ci) {
    Ctx& c = h->ctx[ci % h->ns];
    const int rows = (int)std::min<long long>(h->R, h->N - r0);
    const bool cached = emu_on && h->phi_cache != nullptr;   // Phi's digits of every row are resident since pass 1
    const long long rpitch = 2 * h->Kp, rplane = (long long)h->R * 2 * h->Kp;
    if (!cached && (rc = run_features(h, c, h->dX + r0 * h->D, h->xvec, rows, nullptr, nullptr))) return rc;
    if (emu_on) {
      if (!cached && (rc = emu_digits_slab(h, c, h->ks2, false, r0, rows, nullptr, nullptr, c.digR, nullptr, nullptr,
                                           nullptr, rpitch, rplane, nullptr, nullptr)))
        return rc;
      if ((rc = emu_tgemm(h, c, rows, r0, cached))) return rc;
    } else {
      if ((rc = run_tgemm(h, c, rows))) return rc;
    }
    trace_begin(h, "seeds", c.st);
    if (emu_on)
      k_seeds_rows<<<cdiv(rows, 256), 256, 0, c.st>>>(rows, c.mu, c.vpart, h->R, nkb, h->dy + r0, h->tau, h->gmu + r0,
                                                     h->gv + r0, c.slots);
    else
      k_seeds<<<kSeedBlocks, 256, 0, c.st>>>(c.Phi, h->ldk, rows, K, h->alpha, c.vpart, h->R, nkb, h->dy + r0,
                                            h->tau, h->gmu + r0, h->gv + r0, c.Psi, c.slots, h->want_grad, nullptr);
    trace_end(h, c.st);
    CKL();
    if (h->want_grad && emu_on) {
      // H += (gv Phi)T Phi with the chunk's own power-of-two bound of |gv|
      trace_begin(h, "absmax", c.st);
    emu::k_absmax_pow2<<<1, 1024, 0, c.st>>>(h->gv + r0, rows, c.sig, h->gvmax);
      trace_end(h, c.st);
    CKL();
      // row digits of Psi = (gv / sigma_chunk) Phi into the second half of the chunk buffer, and q = PhiT.gmu,
      // from one pass over Phi (slab or cache); H = PsiT.Phi reads both halves MN-major
      if ((rc = emu_digits_slab(h, c, h->ks2, cached, r0, rows, h->gv + r0, c.sig + 1, c.digR + h->Kp, nullptr, nullptr,
                                nullptr, rpitch, rplane, h->gmu + r0, c.colpart)))
        return rc;
      if ((rc = emu_gram(h, c, true, rows, r0, cached))) return rc;
    } else if (h->want_grad) {
      if ((rc = run_syrk(h, c, true, rows))) return rc;
      trace_begin(h, "colsum", c.st);
    k_colsum<<<dim3(cdiv(K, 256), kColSplit), 256, 0, c.st>>>(c.Phi, h->ldk, rows, K, h->gmu + r0, c.colpart, K);
      trace_end(h, c.st);
    CKL();
    }
  }
  if ((rc = join_streams(h))) return rc;
  if (h->want_grad) {
    k_sum_parts<<<cdiv((long long)KK, 256), 256, 0, h->stream>>>(h->acc_all, (long long)KK, h->ns * h->gsplit, (int)KK, H);
    CKL();
    k_sum_parts<<<cdiv(K, 256), 256, 0, h->stream>>>(h->colpart_all, K, h->csplit * h->ns, K, q);
    CKL();
  } else {
    CK(cudaMemsetAsync(h->red2, 0, (KK + K) * sizeof(double), h->stream));
  }
  // slots are [ns][slot_stride] with (E, g_tau) pairs in the first 2*kSeedBlocks entries of each
  k_reduce_slots<<<1, 256, 0, h->stream>>>(h->slots_all, (int)(h->slot_stride / 2) * h->ns, 2, 2, Eg);
  CKL();
  CK(cudaEventRecord(h->ev[3], h->stream));
  return SCFGP_OK;
}

int scfgp_eval_backward(scfgp_handle* h) {
  if (!h) return SCFGP_ERR_ARG;
  if (!h->have_factors || !h->want_grad) { h->err = "backward without factors/grad"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  const int K = h->K;
  const size_t KK = (size_t)K * h->ldk;
  int rc;
  double* H = h->red2;
  double* q = h->red2 + KK;
  k_symmetrize<<<dim3(cdiv(K, 32), cdiv(K, 32)), dim3(32, 32), 0, h->stream>>>(H, h->ldk, K);
  CKL();
  k_gemv_n<<<cdiv(K, 8), 256, 0, h->stream>>>(h->Bm, h->ldk, K, K, q, h->tvec);
  CKL();
  // BH = B.H ; BHB = BH.B  (H, B symmetric -> NT)
  if ((rc = launch_gemm<CfgBig, true, true, 2, 2>(h, h->Bm, h->ldk, K, H, h->ldk, K, h->BH, h->ldk, 1.0, 0.0,
                                                  h->fullK, K)))
    return rc;
  if ((rc = launch_gemm<CfgBig, true, true, 2, 2>(h, h->BH, h->ldk, K, h->Bm, h->ldk, K, h->BHB, h->ldk, 1.0,
                                                  0.0, h->fullK, K)))
    return rc;
  k_abar<<<K, 256, 0, h->stream>>>(h->Bm, h->BHB, h->ldk, K, h->tvec, h->alpha, h->sig2n, h->A2, h->B2,
                                   h->pbar, h->trslots);
  CKL();
  k_reduce_slots<<<1, 256, 0, h->stream>>>(h->trslots, K, 1, 1, h->scal + 3);
  CKL();
  if (h->prec == SCFGP_PREC_EMU_I8 && (rc = emu_prepare_c(h))) return rc;
  return SCFGP_OK;
}

int scfgp_eval_pass3(scfgp_handle* h) {
  if (!h) return SCFGP_ERR_ARG;
  if (!h->have_factors || !h->want_grad) { h->err = "pass3 without factors/grad"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  const int K = h->K, J = h->J, D = h->D;
  int rc;
  CK(cudaEventRecord(h->ev[4], h->stream));
  const size_t DJ = (size_t)D * h->ldj;
  double* dW = h->red3;
  double* dphi = dW + DJ;
  double* sbar = dphi + h->ldj;
  const size_t smem = SmemSize<CfgBig, true, true>::BYTES;
  const int nfb = cdiv(J, 64);
  CK(cudaMemsetAsync(h->red3, 0, (DJ + h->ldj + 1) * sizeof(double), h->stream));
  if ((rc = fork_streams(h))) return rc;
  for (int i = 0; i < h->ns; ++i) {
    Ctx& c = h->ctx[i];
    CK(cudaMemsetAsync(c.dW, 0, DJ * kDwSplit * sizeof(double), c.st));
    CK(cudaMemsetAsync(c.colpart, 0, (size_t)h->csplit * K * sizeof(double), c.st));
    CK(cudaMemsetAsync(c.slots, 0, (size_t)h->slot_stride * sizeof(double), c.st));
  }
  long long ci = 0;
  for (long long r0 = 0; r0 < h->N; r0 += h->R, ++ci) {
    Ctx& c = h->ctx[ci % h->ns];
    const int rows = (int)std::min<long long>(h->R, h->N - r0);
    const double* Xc = h->dX + r0 * D;
    const bool emu_on = h->prec == SCFGP_PREC_EMU_I8;
    if (emu_on) {
      const bool cached = h->phi_cache != nullptr;
      const long long rpitch = 2 * h->Kp, rplane = (long long)h->R * 2 * h->Kp;
      // A operand digits [Phi | (gv / sigma) Phi], both against the scale s2: with the cache only the Psi
      // half is made (from the cached digits), the Phi half is read from the cache itself
      if (cached) {
        // (when J % 8 != 0 it also rewrites the float64 Phi slab from the cached digits for the backward
        // epilogue; otherwise the epilogue rebuilds its Phi values from the cache directly)
        if ((rc = emu_digits_slab(h, c, h->ks2, true, r0, rows, h->gv + r0, h->sigg + 1, c.digR + h->Kp, nullptr,
                                  nullptr, nullptr, rpitch, rplane, nullptr, nullptr, h->J % 8 != 0)))
          return rc;
      } else {
        if ((rc = run_features(h, c, Xc, h->xvec, rows, nullptr, nullptr))) return rc;
        if ((rc = emu_digits_slab(h, c, h->ks2, false, r0, rows, nullptr, nullptr, c.digR, h->gv + r0, h->sigg + 1,
                                  c.digR + h->Kp, rpitch, rplane, nullptr, nullptr)))
          return rc;
      }
      if ((rc = emu_backward(h, c, rows, r0, cached))) return rc;
    } else if ((rc = run_features(h, c, Xc, h->xvec, rows, h->gv + r0, c.Psi))) return rc;
    const int nitems = cdiv(rows, 128) * nfb;  // items are rb-major
    if (!emu_on) {
      prof_begin(h, SCFGP_FAM_BACKWARD, c.st);
      if (h->use_tma)
        k_backward_tma<<<nitems, TMA_THREADS, TMA_SMEM_BYTES, c.st>>>(c.mapPhi, c.mapPsi, h->mapA2, h->mapB2, c.Phi,
                                                                      h->ldk, rows, K, J, h->dy + r0, h->gmu + r0,
                                                                      h->pbar, h->alpha, c.Zbar, h->ldj, c.slots,
                                                                      h->bwd.d);
      else
        k_backward<<<nitems, CfgBig::NT, smem, c.st>>>(c.Phi, c.Psi, h->ldk, rows, K, J, h->A2, h->B2, h->ldk,
                                                       h->dy + r0, h->gmu + r0, h->pbar, h->alpha, c.Zbar, h->ldj,
                                                       c.slots, h->bwd.d);
      prof_end(h, SCFGP_FAM_BACKWARD, c.st);
      CKL();
    }
    // dW partials += XcT . Zbar, the row reduction split over kDwSplit CTAs per tile
    {
      const int split_len = cdiv(cdiv(rows, kDwSplit), 16) * 16;
      const dim3 grid(h->dw.count, kDwSplit);
      const size_t dsm = SmemSize<CfgDw, false, false>::BYTES;
      trace_begin(h, "dw", c.st);
      if (h->xvec == 2)
        k_dw_splitk<2><<<grid, CfgDw::NT, dsm, c.st>>>(Xc, D, D, c.Zbar, h->ldj, J, rows, split_len, c.dW, h->ldj,
                                                         (long long)DJ, h->dw.d);
      else
        k_dw_splitk<1><<<grid, CfgDw::NT, dsm, c.st>>>(Xc, D, D, c.Zbar, h->ldj, J, rows, split_len, c.dW, h->ldj,
                                                         (long long)DJ, h->dw.d);
      trace_end(h, c.st);
      CKL();
    }
    if (!emu_on) {   // (the emulated backward kernel leaves dphi = colsum(Zbar) in c.colpart itself)
      trace_begin(h, "colsum", c.st);
      k_colsum<<<dim3(cdiv(J, 256), kColSplit), 256, 0, c.st>>>(c.Zbar, h->ldj, rows, J, nullptr, c.colpart, K);
      trace_end(h, c.st);
      CKL();
    }
  }
  if ((rc = join_streams(h))) return rc;
  k_sum_parts<<<cdiv((long long)DJ, 256), 256, 0, h->stream>>>(h->dW_all, (long long)DJ, h->ns * kDwSplit, (int)DJ, dW);
  CKL();
  k_sum_parts<<<cdiv(J, 256), 256, 0, h->stream>>>(h->colpart_all, K, h->csplit * h->ns, J, dphi);
  CKL();
  k_reduce_slots<<<1, 256, 0, h->stream>>>(h->slots_all, (int)h->slot_stride * h->ns, 1, 1, sbar);
  CKL();
  CK(cudaEventRecord(h->ev[5], h->stream));
  return SCFGP_OK;
}

int scfgp_eval_finish(scfgp_handle* h, double* cost, double* grad, double* alpha, double* Li) {
  if (!h || !cost) return SCFGP_ERR_ARG;
  if (!h->have_factors) { h->err = "finish before factor"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  const int K = h->K, J = h->J, D = h->D, S = h->S, M = h->M;
  const size_t KK = (size_t)K * h->ldk;
  double sc[8] = {0};       // logdetL, beta.beta, p.alpha, tr(Abar)
  double r1[1], r2[2];
  CK(cudaMemcpyAsync(sc, h->scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(r1, h->red1 + KK + K, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(r2, h->red2 + KK + K, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  std::vector<double> g3;
  if (grad) {
    if (!h->want_grad) { h->err = "gradient requested but eval_begin had want_grad=0"; return SCFGP_ERR_ARG; }
    g3.resize((size_t)D * h->ldj + h->ldj + 1);
    CK(cudaMemcpyAsync(g3.data(), h->red3, g3.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  if (alpha) CK(cudaMemcpyAsync(alpha, h->alpha, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (Li) CK(cudaMemcpy2DAsync(Li, (size_t)K * sizeof(double), h->Li, h->ldk * sizeof(double),
                               (size_t)K * sizeof(double), K, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  prof_collect(h);
  for (int i = 0; i < 3; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[2 * i], h->ev[2 * i + 1]) == cudaSuccess) h->ms_pass[i] = ms;
    else { h->ms_pass[i] = 0; (void)cudaGetLastError(); }
  }
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]) == cudaSuccess) h->ms_pass[3] = ms;
    else { h->ms_pass[3] = 0; (void)cudaGetLastError(); }
  }
  const double Ng = (double)h->Nglobal;
  const double logdetL = sc[0], bb = sc[1], pa = sc[2], trA = sc[3];
  const double yy = r1[0], E = r2[0], gtau = r2[1];
  std::vector<double> pen_dlF, pen_dF;
  const double pen = penalty_and_grad(h, pen_dlF, pen_dF);
  // nlml: SCFGP/SCFGP.py:125-126 ; cost: :128
  const double nlml = 2 * logdetL + E + (yy - bb) / h->sig2n + 2 * (Ng - M) * h->a;
  *cost = (nlml + pen) / Ng;
  if (!std::isfinite(*cost)) { h->err = "cost is not finite"; return SCFGP_ERR_NONFINITE; }
  if (grad) {
    const double* dW = g3.data();
    const double* dphi = dW + (size_t)D * h->ldj;
    const double sbar_s = dphi[h->ldj];  // = sum(Phibar * Phi) = d nlml / d b
    const double d_a = 2 * h->sig2n * (trA - (yy - pa) / (h->sig2n * h->sig2n)) + 2 * (Ng - M);
    const double d_b = sbar_s;
    const double d_c = gtau / (1 + std::exp(-h->c));
    // A.6: centering + penalty + chain through F = l_F r_F^T
    std::vector<double> Wl((size_t)D * S), WF((size_t)D * M);
    const bool phase = h->fftype == SCFGP_FF_PHASE;
    for (int d = 0; d < D; ++d) {
      for (int s2 = 0; s2 < S; ++s2)
        Wl[(size_t)d * S + s2] = dW[(size_t)d * h->ldj + s2] - (phase ? dphi[s2] / D : 0.0) + pen_dlF[(size_t)d * S + s2];
      for (int m = 0; m < M; ++m)
        WF[(size_t)d * M + m] = dW[(size_t)d * h->ldj + S + m] - (phase ? dphi[S + m] / D : 0.0) + pen_dF[(size_t)d * M + m];
    }
    double* g = grad;
    g[0] = d_a / Ng; g[1] = d_b / Ng; g[2] = d_c / Ng;
    double* g_lf = g + 3;
    double* g_rf = g_lf + (size_t)D * S;
    double* g_lp = g_rf + (size_t)M * S;
    double* g_p = g_lp + S;
    for (int d = 0; d < D; ++d)
      for (int s2 = 0; s2 < S; ++s2) {
        double acc = Wl[(size_t)d * S + s2];
        for (int m = 0; m < M; ++m) acc += WF[(size_t)d * M + m] * h->rF[(size_t)m * S + s2];
        g_lf[(size_t)d * S + s2] = acc / Ng;
      }
    for (int m = 0; m < M; ++m)
      for (int s2 = 0; s2 < S; ++s2) {
        double acc = 0;
        for (int d = 0; d < D; ++d) acc += WF[(size_t)d * M + m] * h->lF[(size_t)d * S + s2];
        g_rf[(size_t)m * S + s2] = acc / Ng;
      }
    for (int s2 = 0; s2 < S; ++s2) g_lp[s2] = phase ? dphi[s2] / Ng : 0.0;
    for (int m = 0; m < M; ++m) g_p[m] = phase ? dphi[S + m] / Ng : 0.0;
  }
  return SCFGP_OK;
}

int scfgp_eval(scfgp_handle* h, const double* params, double* cost, double* grad, double* alpha, double* Li) {
  int rc;
  if ((rc = scfgp_eval_begin(h, params, grad != nullptr))) return rc;
  if ((rc = scfgp_eval_pass1(h))) return rc;
  if ((rc = scfgp_eval_factor(h))) return rc;
  if ((rc = scfgp_eval_pass2(h))) return rc;
  if (grad) {
    if ((rc = scfgp_eval_backward(h))) return rc;
    if ((rc = scfgp_eval_pass3(h))) return rc;
  }
  return scfgp_eval_finish(h, cost, grad, alpha, Li);
}

int scfgp_reduce_buffer(scfgp_handle* h, int which, void** dptr, long long* count) {
  if (!h || !dptr || !count) return SCFGP_ERR_ARG;
  const long long K = h->K, KK = K * h->ldk;
  switch (which) {
    case SCFGP_RED1: *dptr = h->red1; *count = KK + K + 1; break;
    case SCFGP_RED2: *dptr = h->red2; *count = KK + K + 2; break;
    case SCFGP_RED3: *dptr = h->red3; *count = (long long)h->D * h->ldj + h->ldj + 1; break;
    default: return SCFGP_ERR_ARG;
  }
  return SCFGP_OK;
}

int scfgp_device_factors(scfgp_handle* h, void** d_alpha, void** d_Li, long long* ld_Li) {
  if (!h || !h->have_factors) return SCFGP_ERR_ARG;
  if (d_alpha) *d_alpha = h->alpha;
  if (d_Li) *d_Li = h->Li;
  if (ld_Li) *ld_Li = h->ldk;
  return SCFGP_OK;
}

int scfgp_set_tracing(scfgp_handle* h, int on) {
  if (!h) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  for (auto& r : h->trace) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  h->trace.clear();
  h->tracing = on != 0;
  return SCFGP_OK;
}

int scfgp_trace_dump(scfgp_handle* h, const char* path) {
  if (!h || !path) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  FILE* f = fopen(path, "w");
  if (!f) { h->err = std::string("cannot open ") + path; return SCFGP_ERR_ARG; }
  fprintf(f, "{\"traceEvents\": [\n");
  bool first = true;
  for (size_t i = 0; i < h->trace.size(); ++i) {
    float t0 = 0, dur = 0;
    if (cudaEventElapsedTime(&t0, h->trace[0].e0, h->trace[i].e0) != cudaSuccess ||
        cudaEventElapsedTime(&dur, h->trace[i].e0, h->trace[i].e1) != cudaSuccess) {
      (void)cudaGetLastError();
      continue;
    }
    fprintf(f, "%s{\"name\": \"%s\", \"ph\": \"X\", \"pid\": 0, \"tid\": %d, \"ts\": %.3f, \"dur\": %.3f}",
            first ? "" : ",\n", h->trace[i].name, h->trace[i].lane, t0 * 1e3, dur * 1e3);
    first = false;
  }
  fprintf(f, "\n]}\n");
  fclose(f);
  return SCFGP_OK;
}

int scfgp_set_profiling(scfgp_handle* h, int family) {
  if (!h || family < SCFGP_FAM_NONE || family > SCFGP_FAM_SYRK) return SCFGP_ERR_ARG;
  h->prof_family = family;
  h->prof_ms = 0; h->prof_launches = 0; h->prof_used = 0;
  return SCFGP_OK;
}

int scfgp_get_profile(const scfgp_handle* h, double* total_ms, long long* launches) {
  if (!h) return SCFGP_ERR_ARG;
  if (total_ms) *total_ms = h->prof_ms;
  if (launches) *launches = h->prof_launches;
  return SCFGP_OK;
}

int scfgp_last_pass_ms(const scfgp_handle* h, float* p1, float* p2, float* p3, float* dense) {
  if (!h) return SCFGP_ERR_ARG;
  if (p1) *p1 = h->ms_pass[0];
  if (p2) *p2 = h->ms_pass[1];
  if (p3) *p3 = h->ms_pass[2];
  if (dense) *dense = h->ms_pass[3];
  return SCFGP_OK;
}

// ---------------------------------------------------------------------------------
// prediction                                                 SCFGP/SCFGP.py:138-148
// ---------------------------------------------------------------------------------
int scfgp_predict_device(scfgp_handle* h, const double* params, const double* dXs, long long T, double* dmu,
                         double* dstd) {
  if (!h || !params || !dXs || !dmu || !dstd || T <= 0) return SCFGP_ERR_ARG;
  if (!h->have_factors) { h->err = "predict needs alpha/Li (evaluate first or pass them)"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  int rc;
  if ((rc = unpack_and_upload(h, params))) return rc;
  const int K = h->K;
  const int nkb = h->vparts;
  const int xvec = (((uintptr_t)dXs % 16) == 0 && h->D % 2 == 0) ? 2 : 1;
  if ((rc = fork_streams(h))) return rc;
  long long ci = 0;
  for (long long r0 = 0; r0 < T; r0 += h->R, ++ci) {
    Ctx& c = h->ctx[ci % h->ns];
    const int rows = (int)std::min<long long>(h->R, T - r0);
    if ((rc = run_features(h, c, dXs + r0 * h->D, xvec, rows, nullptr, nullptr))) return rc;
    if (h->prec == SCFGP_PREC_EMU_I8) {
      if ((rc = emu_digits_slab(h, c, h->ks2, false, r0, rows, nullptr, nullptr, c.digR, nullptr, nullptr, nullptr,
                                2 * h->Kp, (long long)h->R * 2 * h->Kp, nullptr, nullptr)))
        return rc;
      if ((rc = emu_tgemm(h, c, rows, 0, false))) return rc;
    } else if ((rc = run_tgemm(h, c, rows))) return rc;
    trace_begin(h, "pred_finish", c.st);
    k_pred_finish<<<kSeedBlocks, 256, 0, c.st>>>(c.Phi, h->ldk, rows, K, h->alpha, c.vpart, h->R, nkb, h->tau,
                                                dmu + r0, dstd + r0,
                                                h->prec == SCFGP_PREC_EMU_I8 ? c.mu : nullptr);
    trace_end(h, c.st);
    CKL();
  }
  return join_streams(h);
}

int scfgp_predict(scfgp_handle* h, const double* params, const double* Xs, long long T, const double* alpha,
                  const double* Li, double* mu, double* sd) {
  if (!h || !params || !Xs || !mu || !sd || T <= 0) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const int K = h->K;
  if ((alpha == nullptr) != (Li == nullptr)) { h->err = "pass both alpha and Li or neither"; return SCFGP_ERR_ARG; }
  if (alpha) {
    CK(cudaMemcpyAsync(h->alpha, alpha, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpy2DAsync(h->Li, h->ldk * sizeof(double), Li, (size_t)K * sizeof(double),
                         (size_t)K * sizeof(double), K, cudaMemcpyHostToDevice, h->stream));
    h->have_factors = true;
    int rcl;
    if (h->prec == SCFGP_PREC_EMU_I8 && (rcl = emu_prepare_li(h))) return rcl;
  }
  double *dXs = nullptr, *dout = nullptr;
  CK(cudaMalloc((void**)&dXs, (size_t)T * h->D * sizeof(double)));
  if (cudaMalloc((void**)&dout, (size_t)2 * T * sizeof(double)) != cudaSuccess) {
    cudaFree(dXs); h->err = "cudaMalloc(pred out)"; return SCFGP_ERR_CUDA;
  }
  int rc = SCFGP_OK;
  cudaError_t e = cudaMemcpyAsync(dXs, Xs, (size_t)T * h->D * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) rc = scfgp_predict_device(h, params, dXs, T, dout, dout + T);
  if (e == cudaSuccess && rc == SCFGP_OK) e = cudaMemcpyAsync(mu, dout, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && rc == SCFGP_OK) e = cudaMemcpyAsync(sd, dout + T, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(dXs);
  cudaFree(dout);
  if (e != cudaSuccess) { h->err = std::string("predict: ") + cudaGetErrorString(e); return SCFGP_ERR_CUDA; }
  return rc;
}

// ---------------------------------------------------------------------------------
// self-test hooks
// ---------------------------------------------------------------------------------
int scfgp_dbg_gemm(int device, int layout, int m, int n, int k, const double* A, const double* B, double* C) {
  const bool tma = layout >= 10;
  if (tma) layout -= 10;
  if (m <= 0 || n <= 0 || k <= 0 || !A || !B || !C || layout < 0 || layout > 2 || (tma && layout == 2))
    return SCFGP_ERR_ARG;
  scfgp_handle tmp;
  scfgp_handle* h = &tmp;
  h->device = device;
  CK(cudaSetDevice(device));
  // pad leading dimensions to even so the 16-byte path applies
  const long long lda = layout == 1 ? (m + 1) / 2 * 2 : (k + 1) / 2 * 2;
  const long long ldb = layout == 0 ? (k + 1) / 2 * 2 : (n + 1) / 2 * 2;
  const long long ldc = n;
  const int arows_store = layout == 1 ? k : m, acols = layout == 1 ? m : k;
  const int brows_store = layout == 0 ? n : k, bcols = layout == 0 ? k : n;
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  CK(cudaMalloc((void**)&dA, (size_t)arows_store * lda * sizeof(double)));
  CK(cudaMalloc((void**)&dB, (size_t)brows_store * ldb * sizeof(double)));
  CK(cudaMalloc((void**)&dC, (size_t)m * ldc * sizeof(double)));
  CK(cudaMemset(dA, 0, (size_t)arows_store * lda * sizeof(double)));
  CK(cudaMemset(dB, 0, (size_t)brows_store * ldb * sizeof(double)));
  CK(cudaMemcpy2D(dA, lda * sizeof(double), A, (size_t)acols * sizeof(double), (size_t)acols * sizeof(double),
                  arows_store, cudaMemcpyHostToDevice));
  CK(cudaMemcpy2D(dB, ldb * sizeof(double), B, (size_t)bcols * sizeof(double), (size_t)bcols * sizeof(double),
                  brows_store, cudaMemcpyHostToDevice));
  std::vector<int4> v;
  for (int i = 0; i < cdiv(m, 128); ++i)
    for (int j = 0; j < cdiv(n, 128); ++j) v.push_back(make_int4(i, j, 0, -1));
  ItemList il;
  int rc = upload_items(h, v, &il);
  if (!rc && tma) {
    CUtensorMap ta, tb;
    rc = make_tmap(h, &ta, dA, acols, arows_store, lda);
    if (!rc) rc = make_tmap(h, &tb, dB, bcols, brows_store, ldb);
    if (!rc) {
      dim3 grid(cdiv(n, 128), cdiv(m, 128));
      if (layout == 0) {
        cudaFuncSetAttribute(k_gemm_tma_dbg<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_BYTES);
        k_gemm_tma_dbg<true, true><<<grid, TMA_THREADS, TMA_SMEM_BYTES>>>(ta, tb, k, m, n, dC, ldc);
      } else {
        cudaFuncSetAttribute(k_gemm_tma_dbg<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_BYTES);
        k_gemm_tma_dbg<false, false><<<grid, TMA_THREADS, TMA_SMEM_BYTES>>>(ta, tb, k, m, n, dC, ldc);
      }
    }
  } else if (!rc) {
    if (layout == 0) rc = launch_gemm<CfgBig, true, true, 2, 2>(h, dA, lda, m, dB, ldb, n, dC, ldc, 1.0, 0.0, il, k);
    else if (layout == 1) rc = launch_gemm<CfgBig, false, false, 2, 2>(h, dA, lda, m, dB, ldb, n, dC, ldc, 1.0, 0.0, il, k);
    else rc = launch_gemm<CfgBig, true, false, 2, 2>(h, dA, lda, m, dB, ldb, n, dC, ldc, 1.0, 0.0, il, k);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (!rc && e == cudaSuccess) e = cudaMemcpy(C, dC, (size_t)m * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  for (int4* p : h->to_free) cudaFree(p);
  h->to_free.clear();
  if (rc) return rc;
  return e == cudaSuccess ? SCFGP_OK : SCFGP_ERR_CUDA;
}

int scfgp_dbg_emu_gemm(int device, int digits, int m, int n, int k, const double* A, const double* B, double* C) {
  if (m <= 0 || n <= 0 || k <= 0 || !A || !B || !C || digits < 4 || digits > 8) return SCFGP_ERR_ARG;
  scfgp_handle tmp;
  scfgp_handle* h = &tmp;
  h->device = device;
  h->ks = h->ks2 = digits;
  CK(cudaSetDevice(device));
  const long long kp = (k + 15) / 16 * 16;
  double *dA = nullptr, *dB = nullptr, *dC = nullptr, *sc = nullptr;
  int8_t *gA = nullptr, *gB = nullptr;
  CK(cudaMalloc((void**)&dA, (size_t)m * k * sizeof(double)));
  CK(cudaMalloc((void**)&dB, (size_t)n * k * sizeof(double)));
  CK(cudaMalloc((void**)&dC, (size_t)m * n * sizeof(double)));
  CK(cudaMalloc((void**)&sc, (size_t)2 * (m + n) * sizeof(double)));
  CK(cudaMalloc((void**)&gA, (size_t)digits * m * kp));
  CK(cudaMalloc((void**)&gB, (size_t)digits * n * kp));
  CK(cudaMemcpy(dA, A, (size_t)m * k * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice));
  double *as = sc, *ai = sc + m, *bs = sc + 2 * m, *bi = bs + n;
  emu::k_rowscale<<<m, 128>>>(dA, k, nullptr, 0, nullptr, m, k, 0, as, ai);
  emu::k_rowscale<<<n, 128>>>(dB, k, nullptr, 0, nullptr, n, k, 0, bs, bi);
  int rc = emu_digits_rows(h, nullptr, dA, k, m, k, nullptr, ai, 1.0, nullptr, 0, gA, kp, (long long)m * kp, 0);
  if (!rc) rc = emu_digits_rows(h, nullptr, dB, k, n, k, nullptr, bi, 1.0, nullptr, 0, gB, kp, (long long)n * kp, 0);
  CUtensorMap ma, mb;
  if (!rc) rc = make_tmap_digits(h, &ma, gA, k, m, kp, digits, umma::TM);
  if (!rc) rc = make_tmap_digits(h, &mb, gB, k, n, kp, digits, umma::TN);
  if (!rc) {
    const dim3 grid(cdiv(n, umma::TN), cdiv(m, umma::TM));
    EMU_KS_SWITCH(digits, {
      rc = set_emu_attrs<KS>(h);
      if (!rc)
        emu::k_gemm_i8_dbg<KS><<<grid, umma::THREADS, umma::Geo<KS>::SMEM>>>(ma, mb, m, n, k, as, bs, dC, n);
    });
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (!rc && e == cudaSuccess) e = cudaMemcpy(C, dC, (size_t)m * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(sc); cudaFree(gA); cudaFree(gB);
  if (rc) return rc;
  return e == cudaSuccess ? SCFGP_OK : SCFGP_ERR_CUDA;
}

int scfgp_dbg_factor(int device, int K, const double* A, double* L, double* Li, long long* info) {
  if (K <= 0 || K % 2 != 0 || !A || !L || !Li || !info) return SCFGP_ERR_ARG;
  // borrow a handle of matching K (D=1, S=1, M=K/2-1) for its work lists and buffers
  if (K < 4) return SCFGP_ERR_ARG;
  scfgp_handle* h = nullptr;
  int rc = scfgp_create(1, 1, K / 2 - 1, SCFGP_FF_PHASE, device, &h);
  if (rc) { if (h) scfgp_destroy(h); return rc; }
  const size_t KK = (size_t)K * h->ldk;
  cudaError_t e = cudaMemcpy(h->L, A, KK * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->info, 0, sizeof(long long));
  if (e == cudaSuccess) rc = factor_device(h, h->L, h->Li, K, h->ldk);
  if (e == cudaSuccess && !rc) e = cudaDeviceSynchronize();
  if (e == cudaSuccess && !rc) e = cudaMemcpy(L, h->L, KK * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && !rc) e = cudaMemcpy(Li, h->Li, KK * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && !rc) e = cudaMemcpy(info, h->info, sizeof(long long), cudaMemcpyDeviceToHost);
  scfgp_destroy(h);
  if (rc) return rc;
  return e == cudaSuccess ? SCFGP_OK : SCFGP_ERR_CUDA;
}

int scfgp_dbg_features(scfgp_handle* h, const double* params, const double* X, long long N, double* Phi) {
  if (!h || !params || !X || !Phi || N <= 0) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc;
  if ((rc = unpack_and_upload(h, params))) return rc;
  double* dXs = nullptr;
  CK(cudaMalloc((void**)&dXs, (size_t)N * h->D * sizeof(double)));
  CK(cudaMemcpy(dXs, X, (size_t)N * h->D * sizeof(double), cudaMemcpyHostToDevice));
  const int xvec = (h->D % 2 == 0) ? 2 : 1;
  Ctx& c = h->ctx[0];
  for (long long r0 = 0; r0 < N && !rc; r0 += h->R) {
    const int rows = (int)std::min<long long>(h->R, N - r0);
    rc = run_features(h, c, dXs + r0 * h->D, xvec, rows, nullptr, nullptr);
    if (!rc) {
      cudaError_t e = cudaMemcpy2DAsync(Phi + r0 * h->K, (size_t)h->K * sizeof(double), c.Phi,
                                        h->ldk * sizeof(double), (size_t)h->K * sizeof(double), rows,
                                        cudaMemcpyDeviceToHost, c.st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
      if (e != cudaSuccess) { h->err = cudaGetErrorString(e); rc = SCFGP_ERR_CUDA; }
    }
  }
  cudaFree(dXs);
  return rc;
}

}  // extern "C"
