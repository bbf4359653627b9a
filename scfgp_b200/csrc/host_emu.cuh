// Launch helpers of the emulated-FP64 path (int8 digit products, emu_kernels.cuh) and the digit cache.
// (part of the single translation unit scfgp_abi.cu; included in this order)
#pragma once

namespace {

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
  // the default case (5 digits from a 5-digit cache, one output, 16-byte aligned planes) has its own kernel
  const bool fast5 = h->fast_digits && from_cache && nd == 5 && h->cache_ks == 5 && dst1 && !dst2 && !write_slab && h->Kc % 16 == 0 &&
                     pitch % 16 == 0 && plane % 16 == 0 && (reinterpret_cast<uintptr_t>(dst1) & 15) == 0;
  if (fast5) {
    const dim3 g5(cdiv((K + 15) / 16 * 16, 128), cdiv(rows, 128));
    if (wsum)
      emu::k_digits_cache5<true><<<g5, 256, 0, c.st>>>(h->phi_cache + r0 * h->Kc, h->Kc, h->cache_rows * h->Kc, recon, rows,
                                                       K, 1.0 / h->s2, rowmul1, gptr1, dst1, pitch, plane, wsum, cpart,
                                                       (long long)K);
    else
      emu::k_digits_cache5<false><<<g5, 256, 0, c.st>>>(h->phi_cache + r0 * h->Kc, h->Kc, h->cache_rows * h->Kc, recon,
                                                        rows, K, 1.0 / h->s2, rowmul1, gptr1, dst1, pitch, plane, nullptr,
                                                        nullptr, (long long)K);
  } else if (from_cache) {
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

// pass 1 of the emulated mode with 5 digits: feature map + digit extraction + p = PhiT.y in ONE kernel
// (k_features_digits5); no float64 Phi slab is written.  Needs J % 4 == 0 (see the kernel).
bool fused_features_ok(const scfgp_handle* h) { return h->fuse_features && h->ks == 5 && h->J % 4 == 0; }
int run_features_digits(scfgp_handle* h, Ctx& c, const double* Xc, int xvec, int rows, int8_t* dst, long long pitch,
                        long long plane, const double* wsum, double* cpart) {
  const int nfb = cdiv(h->J, emu::CfgFeatD::BN);
  // 64-row tiles over the chunk's rows rounded up to 128: the pad rows are written as zero digits
  const int nrb = cdiv(rows, 128) * (128 / emu::CfgFeatD::BM);
  prof_begin(h, SCFGP_FAM_FEATURES, c.st);
  trace_begin(h, "features+digits", c.st);
  if (xvec == 2)
    emu::k_features_digits5<2><<<nrb * nfb, emu::CfgFeatD::NT, emu::FEATD_SMEM, c.st>>>(
        Xc, h->D, rows, h->D, h->Wt, h->ldw, h->phi, h->J, h->s, 1.0 / h->s2, dst, pitch, plane, wsum, cpart,
        (long long)h->K, nfb);
  else
    emu::k_features_digits5<1><<<nrb * nfb, emu::CfgFeatD::NT, emu::FEATD_SMEM, c.st>>>(
        Xc, h->D, rows, h->D, h->Wt, h->ldw, h->phi, h->J, h->s, 1.0 / h->s2, dst, pitch, plane, wsum, cpart,
        (long long)h->K, nfb);
  trace_end(h, c.st);
  prof_end(h, SCFGP_FAM_FEATURES, c.st);
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

// launch `kern<KS, 64>` or, with wide tiles (h->tnw == 96: 5 digits), `kern<5, 96>`
// (h->two_phase: the two-phase epilogue of umma_i8.cuh, available up to 5 digits; SCFGP_EPILOGUE=1 selects the
// one-phase epilogue for A/B measurements)
#define EMU_TILE_LAUNCH(kern, grid, ...)                                                                           \
  do {                                                                                                             \
    if (h->tnw == 96) {                                                                                            \
      if (h->ks2 != 5) { h->err = "wide tiles are built for 5 digits"; return SCFGP_ERR_ARG; }                     \
      if (h->two_phase)                                                                                            \
        emu::kern<5, 96, true><<<grid, umma::THREADS, umma::Geo<5, 96>::SMEM, c.sg>>>(__VA_ARGS__);                \
      else                                                                                                         \
        emu::kern<5, 96, false><<<grid, umma::THREADS, umma::Geo<5, 96>::SMEM, c.sg>>>(__VA_ARGS__);               \
    } else if (h->two_phase && h->ks2 <= 5) {                                                                      \
      switch (h->ks2) {                                                                                            \
        case 4: emu::kern<4, 64, true><<<grid, umma::THREADS, umma::Geo<4>::SMEM, c.sg>>>(__VA_ARGS__); break;     \
        default: emu::kern<5, 64, true><<<grid, umma::THREADS, umma::Geo<5>::SMEM, c.sg>>>(__VA_ARGS__); break;    \
      }                                                                                                            \
    } else {                                                                                                       \
      EMU_KS_SWITCH(h->ks2, emu::kern<KS, 64, false><<<grid, umma::THREADS, umma::Geo<KS>::SMEM, c.sg>>>(__VA_ARGS__)); \
    }                                                                                                              \
  } while (0)

int emu_tgemm(scfgp_handle* h, Ctx& c, int rows, long long r0, bool use_cache) {
  const int ntiles = cdiv(h->K + 1, h->tnw) * cdiv(rows, umma::TM);
  int rc;
  if ((rc = gemm_enter(h, c))) return rc;
  prof_begin(h, SCFGP_FAM_TGEMM, c.sg);
  // (two-phase epilogue: the column scales come pre-multiplied by every constant factor, see emu_prepare_li)
  const bool tg2 = h->two_phase && h->ks2 <= 5;
  EMU_TILE_LAUNCH(k_tgemm_i8, std::min(ntiles, h->num_sms), use_cache ? h->mapCacheK : c.mapRowsA, h->mapLiB,
                  use_cache ? (int)r0 : 0, rows, h->K, ntiles, tg2 ? h->liscale_w : h->liscale, h->s2, c.vpart, h->R,
                  c.mu);
  prof_end(h, SCFGP_FAM_TGEMM, c.sg);
  CKL();
  return gemm_leave(h, c);
}

// the lean two-phase backward kernel (k_backward_i8_2p): up to 5 digits, unless SCFGP_EPILOGUE=1
bool emu_two_phase_backward(const scfgp_handle* h) {
  return h->prec == SCFGP_PREC_EMU_I8 && h->two_phase && h->ks2 <= 5 && (h->tnw != 96 || h->ks2 == 5);
}

int emu_backward(scfgp_handle* h, Ctx& c, int rows, long long r0, bool use_cache) {
  const int ntiles = cdiv(h->J, h->tnw / 2) * cdiv(rows, umma::TM);
  // the epilogue rebuilds Phi from the cached digits (8-byte loads: needs J % 8 == 0), else reads the slab
  const bool phi_from_digits = use_cache && h->J % 8 == 0;
  const double recon = h->s2 * std::ldexp(1.0, -(8 * std::max(h->cache_ks, 1) - 1));
  const int grid = std::min(ntiles, h->num_sms);
  int rc;
  if ((rc = gemm_enter(h, c))) return rc;
  prof_begin(h, SCFGP_FAM_BACKWARD, c.sg);
#define BWD_ARGS_2P                                                                                                  \
  c.mapRowsA, use_cache ? h->mapCacheK : c.mapRowsA, h->mapCB, use_cache ? (int)(h->Kp / umma::KB) : 0,              \
      use_cache ? (int)r0 : 0, phi_from_digits ? h->phi_cache : nullptr, h->Kc, h->cache_rows * h->Kc, recon, rows,  \
      h->J, (int)(h->Kp + h->K), ntiles, h->cscale, h->s2, c.Phi, h->ldk, h->dy + r0, h->gmu + r0, h->pbar,          \
      h->alpha, c.Zbar, h->ldj
#define BWD_ARGS BWD_ARGS_2P, c.slots, c.colpart, (long long)h->K
  const bool two_phase = emu_two_phase_backward(h);
  if (two_phase) {
    if (h->tnw == 96) {
      if (h->ks2 != 5) { h->err = "wide tiles are built for 5 digits"; return SCFGP_ERR_ARG; }
      emu::k_backward_i8_2p<5, 96><<<grid, umma::THREADS, umma::Geo<5, 96>::SMEM, c.sg>>>(BWD_ARGS_2P);
    } else if (h->ks2 == 4) {
      emu::k_backward_i8_2p<4, 64><<<grid, umma::THREADS, umma::Geo<4>::SMEM, c.sg>>>(BWD_ARGS_2P);
    } else {
      emu::k_backward_i8_2p<5, 64><<<grid, umma::THREADS, umma::Geo<5>::SMEM, c.sg>>>(BWD_ARGS_2P);
    }
  } else if (h->tnw == 96) {
    emu::k_backward_i8<5, 96, false><<<grid, umma::THREADS, umma::Geo<5, 96>::SMEM, c.sg>>>(BWD_ARGS);
  } else {
    EMU_KS_SWITCH(h->ks2, emu::k_backward_i8<KS, 64, false><<<grid, umma::THREADS, umma::Geo<KS>::SMEM, c.sg>>>(BWD_ARGS));
  }
#undef BWD_ARGS_2P
#undef BWD_ARGS
  prof_end(h, SCFGP_FAM_BACKWARD, c.sg);
  CKL();
  // (the two-phase kernel leaves the reductions to others: dphi = colsum(Zbar) to the dW contraction that reads the
  // slab next (k_dw_splitk), sbar = sum <Phibar, Phi> to two K x K traces in the dense backward phase, k_sbar_rows)
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
  // The cache is an explicit trade of HBM for time (5 N K bytes: 104 GB at cfg4 on one GPU): it is bounded by
  // SCFGP_PHI_CACHE_MAX_GB (default 120) and by the free memory minus 6 GB of head room, and the handle says on
  // stderr when it runs without it (once per data set / precision), because passes 2 and 3 then recompute the
  // feature map (~5 % slower at cfg4).
  const double gb = (double)bytes / (1024.0 * 1024.0 * 1024.0);
  const char* why = nullptr;
  size_t fr = 0, tot = 0;
  CK(cudaMemGetInfo(&fr, &tot));
  if (gb > h->cache_max_gb) why = "above SCFGP_PHI_CACHE_MAX_GB";
  else if (bytes + ((size_t)6 << 30) > fr) why = "not enough free device memory";
  else if (cudaMalloc((void**)&h->phi_cache, bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    h->phi_cache = nullptr;
    why = "cudaMalloc failed";
  }
  if (why) {
    if (h->cache_log)
      fprintf(stderr, "[scfgp_b200] digit cache of Phi declined (%.1f GB wanted, %.1f GB free, cap %.0f GB: %s); "
                      "passes 2 and 3 recompute the feature map\n", gb, (double)fr / (1 << 30), h->cache_max_gb, why);
    return SCFGP_OK;
  }
  if (h->cache_log > 1)
    fprintf(stderr, "[scfgp_b200] digit cache of Phi: %.1f GB (%d digits x %lld rows x %lld columns)\n", gb, h->ks,
            rows, h->Kc);
  h->cache_rows = rows;
  h->cache_ks = h->ks;
  // (the fused feature kernel never writes the pad columns [K, Kc): they must read as zero digits)
  if (h->Kc != h->K) CK(cudaMemsetAsync(h->phi_cache, 0, bytes, h->stream));
  int rc;
  if ((rc = make_tmap_digits(h, &h->mapCacheK, h->phi_cache, h->K, rows, h->Kc, h->ks, umma::TM))) return rc;
  if ((rc = make_tmap_digits_mn(h, &h->mapCacheMN_A, h->phi_cache, h->K, rows, h->Kc, h->ks, umma::TM))) return rc;
  if ((rc = make_tmap_digits_mn(h, &h->mapCacheMN_B, h->phi_cache, h->K, rows, h->Kc, h->ks, umma::TN))) return rc;
  return SCFGP_OK;
}

// liscale[j] * sA * 2^-14 * 2^-8(ks2-1): the one factor the two-phase Phi.LiT epilogue multiplies by.  sA = s2 is
// the digit scale of Phi, a function of the hyper-parameter b: refreshed whenever the parameters are unpacked
// after the factors were set (scfgp_set_factors followed by scfgp_predict_device).
int emu_refresh_liscale_w(scfgp_handle* h) {
  const int K = h->K;
  emu::k_scale_vector<<<cdiv(K + 1, 256), 256, 0, h->stream>>>(
      h->liscale, K + 1, h->s2 * umma::PROD_SCALE * std::ldexp(1.0, -umma::DIGIT_BITS * (h->ks2 - 1)), h->liscale_w);
  CKL();
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
  if ((rc = emu_digits_rows(h, h->stream, h->alpha, K, 1, K, nullptr, h->liinv + K, 1.0, nullptr, 0,
                            h->digLi + (long long)K * h->Kp, h->Kp, plane, 0)))
    return rc;
  return emu_refresh_liscale_w(h);
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

}  // namespace
