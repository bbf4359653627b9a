// Launch helpers of the FP64 (DMMA) path for one row chunk.
// (part of the single translation unit scfgp_abi.cu; included in this order)
#pragma once

namespace {

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

}  // namespace
