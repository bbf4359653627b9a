// Host utilities: error macros, tensor maps, launch profiling / tracing, device allocation, FP64 GEMM launcher, kernel attributes.
// (part of the single translation unit scfgp_abi.cu; included in this order)
#pragma once

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
  // cudaFuncSetAttribute applies to the current device only: one flag per device ordinal, so a
  // second engine on another GPU of the same process also gets the > 48 KB opt-in
  static std::atomic<unsigned long long> attr_done_mask{0};
  const int dev = h->device;
  if (dev < 0 || dev >= 64 || !((attr_done_mask.load() >> dev) & 1ull)) {
    CK(cudaFuncSetAttribute(k_gemm<Cfg, AK, BKM, AVEC, BVEC>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_done_mask.fetch_or(1ull << dev);
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
  CK(cudaFuncSetAttribute(emu::k_features_digits5<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, emu::FEATD_SMEM));
  CK(cudaFuncSetAttribute(emu::k_features_digits5<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, emu::FEATD_SMEM));
  CK(cudaFuncSetAttribute(emu::k_features_digits5<1>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(emu::k_features_digits5<2>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
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
  CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 64, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute((emu::k_backward_i8<KS, 64, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 64, false>), cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute((emu::k_backward_i8<KS, 64, false>), cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  if constexpr (KS <= 5) {
    CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 64, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute((emu::k_backward_i8_2p<KS, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 64, true>), cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute((emu::k_backward_i8_2p<KS, 64>), cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  }
  CK(cudaFuncSetAttribute(emu::k_gemm_i8_dbg<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int mx = (int)cudaSharedmemCarveoutMaxShared;
  CK(cudaFuncSetAttribute(emu::k_gram_i8<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute(emu::k_digits_rows<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute((emu::k_digits_slab<KS, false>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  CK(cudaFuncSetAttribute((emu::k_digits_slab<KS, true>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  if constexpr (KS == 5) {   // wide (96-column) tiles of the two K-major products
    const int smw = (int)umma::Geo<KS, 96>::SMEM;
    CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 96, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smw));
    CK(cudaFuncSetAttribute((emu::k_backward_i8_2p<KS, 96>), cudaFuncAttributeMaxDynamicSharedMemorySize, smw));
    CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 96, true>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
    CK(cudaFuncSetAttribute((emu::k_backward_i8_2p<KS, 96>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
    CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 96, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smw));
    CK(cudaFuncSetAttribute((emu::k_backward_i8<KS, 96, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smw));
    CK(cudaFuncSetAttribute((emu::k_tgemm_i8<KS, 96, false>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
    CK(cudaFuncSetAttribute((emu::k_backward_i8<KS, 96, false>), cudaFuncAttributePreferredSharedMemoryCarveout, mx));
  }
  return SCFGP_OK;
}

}  // namespace
