// A-posteriori guard of the emulated path (part of the single translation unit scfgp_abi.cu).
//
// The emulated products round every operand to 8 KS - 2 fractional bits of its scale.  The product that
// matters most is the Gram G = PhiT.Phi of pass 1: its error dG reaches the results amplified by A^-1
// (cond(A) ~ N sigma_f^2 / sigma_n^2).  After the factor phase the guard MEASURES dG on a
// sample of this rank's rows -- the same rows through the FP64 DMMA Gram and through the digit Gram -- and
// pushes it through the first-order perturbation of the results with the factors that were just computed:
//     d(nlml)  ~  tr(B dG) + alphaT dG alpha / sigma_n^2          (log-determinant and quadratic term)
//     d(alpha) ~  -B dG alpha                                      (B = A^-1 = LiT Li)
// The sample's dG is digit rounding noise (zero-mean, independent rows), so it is extrapolated to all rows
// with sqrt(N_global / sample rows) and a safety factor.  scfgp_eval escalates 5 -> 6 digits -> FP64 and
// re-runs the evaluation when an estimate exceeds its threshold (sticky: the handle stays at the higher
// precision); the phase API reports the estimates so that all ranks can take the decision together.
#pragma once

namespace scfgp {

// one block per row a of the symmetric difference dG (both inputs hold the lower triangle):
//   u[a] = sum_b dG[a][b] alpha[b],  part[a] = sum_b B[a][b] dG[a][b],  part[K + a] = sum_b dG[a][b]^2
__global__ void k_guard_rows(const double* __restrict__ Gemu, const double* __restrict__ G64,
                             const double* __restrict__ Bm, const double* __restrict__ alpha, int K, long long ld,
                             double* __restrict__ u, double* __restrict__ part) {
  const int a = blockIdx.x;
  double su = 0, st = 0, sf = 0;
  for (int b = threadIdx.x; b < K; b += blockDim.x) {
    const long long i = b <= a ? (long long)a * ld + b : (long long)b * ld + a;
    const double d = Gemu[i] - G64[i];
    su = fma(d, alpha[b], su);
    st = fma(Bm[(long long)a * ld + b], d, st);
    sf = fma(d, d, sf);
  }
  __shared__ double sh[3][32];
  su = warp_sum(su); st = warp_sum(st); sf = warp_sum(sf);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = su; sh[1][threadIdx.x >> 5] = st; sh[2][threadIdx.x >> 5] = sf; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0, y = 0, z = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { x += sh[0][w]; y += sh[1][w]; z += sh[2][w]; }
    u[a] = x; part[a] = y; part[K + a] = z;
  }
}
// out[0] = tr(B dG), out[1] = alphaT dG alpha, out[2] = max |B dG alpha|, out[3] = max |alpha|, out[4] = |dG|_F
__global__ void k_guard_final(const double* __restrict__ part, const double* __restrict__ u,
                              const double* __restrict__ w, const double* __restrict__ alpha, int K,
                              double* __restrict__ out) {
  __shared__ double sh[5][32];
  double tr = 0, aua = 0, wm = 0, am = 0, fr = 0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    tr += part[i]; fr += part[K + i];
    aua = fma(alpha[i], u[i], aua);
    wm = fmax(wm, fabs(w[i])); am = fmax(am, fabs(alpha[i]));
  }
  tr = warp_sum(tr); aua = warp_sum(aua); fr = warp_sum(fr);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
  }
  if ((threadIdx.x & 31) == 0) {
    const int wi = threadIdx.x >> 5;
    sh[0][wi] = tr; sh[1][wi] = aua; sh[2][wi] = wm; sh[3][wi] = am; sh[4][wi] = fr;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tr = aua = wm = am = fr = 0;
    for (int wi = 0; wi < (blockDim.x >> 5); ++wi) {
      tr += sh[0][wi]; aua += sh[1][wi]; fr += sh[4][wi];
      wm = fmax(wm, sh[2][wi]); am = fmax(am, sh[3][wi]);
    }
    out[0] = tr; out[1] = aua; out[2] = wm; out[3] = am; out[4] = sqrt(fr);
  }
}

}  // namespace scfgp

namespace {

constexpr int kGuardRows = 2048;       // sample rows (a multiple of 128)
constexpr double kGuardSafety = 4.0;   // on top of the sqrt(N / sample) extrapolation of the rounding noise

// Queue the measurement on the handle's streams (after L, Li, alpha and B exist); results land in
// h->scal[8..12] and are read back by the caller together with the pivot flag.
int guard_measure(scfgp_handle* h) {
  h->guard_sample = 0;
  if (h->prec != SCFGP_PREC_EMU_I8 || !h->guard_on || h->N <= 0) return SCFGP_OK;
  const int K = h->K;
  const size_t KK = (size_t)K * h->ldk;
  const long long ns = std::min<long long>(std::min<long long>(h->N, h->R), kGuardRows);
  const long long r0 = ((h->N - ns) / 2) / 128 * 128;
  Ctx& c = h->ctx[0];
  int rc;
  if (!h->want_grad) {   // B = LiT Li is only formed for the gradient otherwise
    if ((rc = launch_gemm<CfgBig, false, false, 2, 2>(h, h->Li, h->ldk, K, h->Li, h->ldk, K, h->Bm, h->ldk, 1.0, 0.0,
                                                      h->lowerK, K)))
      return rc;
    k_symmetrize<<<dim3(cdiv(K, 32), cdiv(K, 32)), dim3(32, 32), 0, h->stream>>>(h->Bm, h->ldk, K);
    CKL();
  }
  CK(cudaEventRecord(h->ev_fork, h->stream));
  CK(cudaStreamWaitEvent(c.st, h->ev_fork, 0));
  // the sample through the FP64 path: float64 Phi slab, DMMA Gram
  CK(cudaMemsetAsync(c.acc, 0, KK * sizeof(double), c.st));
  if ((rc = run_features(h, c, h->dX + r0 * h->D, h->xvec, (int)ns, nullptr, nullptr))) return rc;
  if ((rc = run_syrk(h, c, false, (int)ns))) return rc;
  CK(cudaMemcpyAsync(h->BH, c.acc, KK * sizeof(double), cudaMemcpyDeviceToDevice, c.st));
  CK(cudaMemsetAsync(c.acc, 0, KK * sizeof(double), c.st));
  // the same rows through the digit path (digits of the very same slab, chunk buffer, no cache)
  if ((rc = emu_digits_slab(h, c, h->ks, false, r0, (int)ns, nullptr, nullptr, c.digR, nullptr, nullptr, nullptr,
                            2 * h->Kp, (long long)h->R * 2 * h->Kp, nullptr, nullptr)))
    return rc;
  if ((rc = emu_gram(h, c, false, (int)ns, r0, false))) return rc;
  CK(cudaEventRecord(c.done, c.st));
  CK(cudaStreamWaitEvent(h->stream, c.done, 0));
  scfgp::k_guard_rows<<<K, 128, 0, h->stream>>>(c.acc, h->BH, h->Bm, h->alpha, K, h->ldk, h->tvec, h->trslots2);
  CKL();
  k_gemv_n<<<cdiv(K, 8), 256, 0, h->stream>>>(h->Bm, h->ldk, K, K, h->tvec, h->pbar);
  CKL();
  scfgp::k_guard_final<<<1, 256, 0, h->stream>>>(h->trslots2, h->tvec, h->pbar, h->alpha, K, h->scal + 8);
  CKL();
  h->guard_sample = ns;
  return SCFGP_OK;
}

// host side: turn the five device scalars into the two estimates
void guard_finish(scfgp_handle* h, const double* g5) {
  h->guard_est_cost = h->guard_est_alpha = 0;
  if (h->guard_sample <= 0) return;
  const double f = kGuardSafety * std::sqrt((double)h->Nglobal / (double)h->guard_sample);
  const double d_nlml = std::fabs(g5[0]) + std::fabs(g5[1]) / h->sig2n;
  h->guard_est_cost = f * d_nlml / (double)h->Nglobal;                 // in units of `cost`
  h->guard_est_alpha = f * g5[2] / std::max(1.0, g5[3]);               // relative to max(1, |alpha|_inf)
  h->guard_dg_fro = g5[4];
}

bool guard_exceeded(const scfgp_handle* h) {
  return h->guard_sample > 0 && (h->guard_est_cost > h->guard_cost_tol || h->guard_est_alpha > h->guard_alpha_tol);
}

}  // namespace
