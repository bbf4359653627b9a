// Parameter unpack (SCFGP/SCFGP.py:74-90), device buffers of the chunk streams, stream fork/join.
// (part of the single translation unit scfgp_abi.cu; included in this order)
#pragma once

namespace {

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
// 96-column tiles for the two K-major products (Phi.LiT, backward): 20 % fewer operand bytes per digit product (the
// kernels are bound by operand delivery, profiles/r1_umma_stream_and_load_ceilings.txt); tensor memory holds at most
// 5 digits x 96 columns, and the wide tile is built and tested for exactly 5.  SCFGP_TILE_N=64|96 overrides the default.
inline void select_tile_width(scfgp_handle* h) {
  h->tnw = 64;
  if (h->prec == SCFGP_PREC_EMU_I8 && h->ks2 == 5) {   // the default digit count (6+ digits do not fit tensor memory)
    h->tnw = kWideTile;
    if (const char* e = getenv("SCFGP_TILE_N")) h->tnw = atoi(e) == 96 ? 96 : 64;
  }
}

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
  select_tile_width(h);
  h->vparts = emu_on ? 2 * cdiv(K + 1, h->tnw) : nkb;
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
  h->csplit = std::max(kColSplit, R / 64);   // one partial row per 64-row tile of the fused feature kernel
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
  for (double** p : {&h->liscale, &h->liscale_w, &h->liinv, &h->cscale, &h->cinv, &h->sigg}) {
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
  select_tile_width(h);
  const int K = h->K, ks = h->ks2;
  const long long Kp = h->Kp;
  int rc;
  if ((rc = dalloc(h, &h->digLi, (size_t)ks * (K + 1) * Kp))) return rc;   // rows 0..K-1: Li, row K: alpha
  if ((rc = dalloc(h, &h->digC, (size_t)ks * K * 2 * Kp))) return rc;
  for (double** p : {&h->liscale, &h->liscale_w, &h->liinv, &h->cscale, &h->cinv})
    if ((rc = dalloc(h, p, K + 1))) return rc;
  if ((rc = dalloc(h, &h->sigg, 2))) return rc;
  if ((rc = dalloc(h, &h->gvmax, 1))) return rc;
  if ((rc = make_tmap_digits(h, &h->mapLiB, h->digLi, K, K + 1, Kp, ks, h->tnw))) return rc;
  if ((rc = make_tmap_digits(h, &h->mapCB, h->digC, Kp + K, K, 2 * Kp, ks, h->tnw))) return rc;
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

}  // namespace
