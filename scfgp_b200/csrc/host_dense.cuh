// Replicated dense phase: Cholesky + inverse (SCFGP/SCFGP.py:106-107), static work lists, teardown, penalty (SCFGP.py:114-117).
// (part of the single translation unit scfgp_abi.cu; included in this order)
#pragma once

namespace {

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
    // B = LiT.Li on lower tiles (i, j): rows k < 128 i of Li's column block i are zero, the reduction starts there
    std::vector<int4> v = lower_tiles(K, 128, 0, -1);
    for (int4& it : v) it.z = it.x * 128;
    if ((rc = upload_items(h, v, &h->lowerK_tri))) return rc;
  }
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
                    h->pbar, h->scal, h->gmu, h->gv, h->colpart, h->slots, h->trslots, h->trslots2, h->ownX, h->owny,
                    h->selX, h->sely})
    if (p) cudaFree(p);
  if (h->selidx) cudaFree(h->selidx);
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
