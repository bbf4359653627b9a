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
#include <atomic>

#include <chrono>
#include <cuda_runtime.h>

#include "../../include/scfgp_b200.h"
#include "kernels.cuh"
#include "kernels_tma.cuh"
#include "emu_kernels.cuh"
#include "mma_peak.cuh"

using namespace scfgp;

#include "host_state.cuh"
#include "host_util.cuh"
#include "host_buffers.cuh"
#include "host_fp64.cuh"
#include "host_emu.cuh"
#include "host_dense.cuh"
#include "host_guard.cuh"

template <int KS, int TNW>
static int mma_stream_peak(int ctas, int steps, int random_data, double* tops, double* clk_per_step, double* sm_mhz) {
  using G = umma::Geo<KS, TNW>;
  long long* d = nullptr;
  if (cudaMalloc((void**)&d, (size_t)ctas * 16) != cudaSuccess) return SCFGP_ERR_CUDA;
  const int smem = G::STAGE_BYTES + 1024 + 256;
  cudaFuncSetAttribute(umma::k_mma_stream<KS, TNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = SCFGP_OK;
  float ms = 0;
  std::vector<long long> hst((size_t)ctas * 2, 0);
  for (int rep = 0; rep < 2 && rc == SCFGP_OK; ++rep) {   // rep 0 warms up (clocks, instruction cache)
    cudaMemset(d, 0, (size_t)ctas * 16);
    cudaEventRecord(e0);
    umma::k_mma_stream<KS, TNW><<<ctas, 128, smem>>>(rep == 0 ? std::min(steps, 8192) : steps, random_data, d);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { rc = SCFGP_ERR_CUDA; break; }
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hst.data(), d, (size_t)ctas * 16, cudaMemcpyDeviceToHost);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (rc) return rc;
  double clk = 0;
  for (int i = 0; i < ctas; ++i) {
    if (!hst[2 * i + 1]) return SCFGP_ERR_CUDA;   // a CTA timed out
    clk += (double)hst[2 * i];
  }
  clk /= ctas;
  const double cols = (double)TNW * KS * (KS + 1) / 2;   // output columns per K = 32 step
  if (tops) *tops = 2.0 * 128 * cols * 32 * (double)steps * ctas / (ms * 1e9);
  if (clk_per_step) *clk_per_step = clk / steps;
  if (sm_mhz) *sm_mhz = clk / (ms * 1e3);
  return SCFGP_OK;
}

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

// ---- upload of host training data in row blocks (scfgp_set_data_host_async) ----------------------------------------
// y first, then X in blocks of a few chunks on the copy stream; one event per block (pass 1 waits per chunk)
static int submit_upload(scfgp_handle* h) {
  if (!h->upload_pending || h->up_submitted) return SCFGP_OK;
  const long long N = h->N;
  CK(cudaMemcpyAsync(h->owny, h->up_y, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, h->cstream));
  h->up_block = (long long)h->R * 8;
  const long long nblk = (N + h->up_block - 1) / h->up_block;
  while ((long long)h->up_ev.size() < nblk) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->up_ev.push_back(e);
  }
  for (long long b = 0; b < nblk; ++b) {
    const long long r0 = b * h->up_block, rows = std::min(h->up_block, N - r0);
    CK(cudaMemcpyAsync(h->ownX + r0 * h->D, h->up_X + r0 * h->D, (size_t)rows * h->D * sizeof(double),
                       cudaMemcpyHostToDevice, h->cstream));
    CK(cudaEventRecord(h->up_ev[b], h->cstream));
  }
  h->up_submitted = true;
  return SCFGP_OK;
}
// wait (on the host) for an upload started by scfgp_set_data_host_async
static int finish_upload(scfgp_handle* h) {
  if (h->upload_pending) {
    int rc;
    if ((rc = submit_upload(h))) return rc;
    CK(cudaStreamSynchronize(h->cstream));
    h->upload_pending = false;
    h->up_X = h->up_y = nullptr;
  }
  return SCFGP_OK;
}


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
  if (const char* e = getenv("SCFGP_PHI_CACHE_MAX_GB")) h->cache_max_gb = atof(e);
  if (const char* e = getenv("SCFGP_PHI_CACHE_LOG")) h->cache_log = atoi(e);
  if (const char* e = getenv("SCFGP_EPILOGUE")) h->two_phase = atoi(e) != 1;
  if (const char* e = getenv("SCFGP_FUSED_FEATURES")) h->fuse_features = atoi(e) != 0;
  if (const char* e = getenv("SCFGP_DIGIT_KERNEL")) h->fast_digits = strcmp(e, "generic") != 0;
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
      h->R = kEmuChunkRows;
    }
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { h->err = "no such CUDA device"; return SCFGP_ERR_CUDA; }
  CK(cudaSetDevice(device));
  CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  if (const char* e = getenv("SCFGP_B_KEEP")) {   // A/B knob: L2 evict_last hint on the replicated B operands
    const int v = atoi(e) != 0;
    CK(cudaMemcpyToSymbol(umma::g_b_keep, &v, sizeof(v)));
  }
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
  CK(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_factor, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking));
  if ((rc = dalloc(h, &h->trslots, K))) return rc;
  if ((rc = dalloc(h, &h->trslots2, 2 * K))) return rc;
  if (const char* e = getenv("SCFGP_GUARD")) h->guard_on = atoi(e) != 0;
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

int scfgp_stream_handoff(scfgp_handle* h, void* other_stream, int direction) {
  if (!h || (direction != SCFGP_HANDOFF_TO_OTHER && direction != SCFGP_HANDOFF_FROM_OTHER)) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  cudaStream_t other = (cudaStream_t)other_stream;
  if (other == h->stream) return SCFGP_OK;   // same queue: already ordered
  if (direction == SCFGP_HANDOFF_TO_OTHER) {
    CK(cudaEventRecord(h->ev_out, h->stream));
    CK(cudaStreamWaitEvent(other, h->ev_out, 0));
  } else {
    CK(cudaEventRecord(h->ev_in, other));
    CK(cudaStreamWaitEvent(h->stream, h->ev_in, 0));
  }
  return SCFGP_OK;
}

static bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int scfgp_eval_prefetch_factors(scfgp_handle* h, double* alpha, double* Li) {
  if (!h) return SCFGP_ERR_ARG;
  if (!h->have_factors) { h->err = "prefetch_factors before factor"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  h->pre_alpha = h->pre_Li = nullptr;
  // A device -> pageable-host cudaMemcpyAsync blocks the calling thread until the copy is done, which
  // would keep the host from queueing pass 2; only page-locked destinations are filled early.
  const bool a_ok = alpha && is_pinned_host(alpha), l_ok = Li && is_pinned_host(Li);
  if (!a_ok && !l_ok) return SCFGP_OK;
  const int K = h->K;
  CK(cudaStreamWaitEvent(h->cstream, h->ev_factor, 0));
  if (a_ok) {
    CK(cudaMemcpyAsync(alpha, h->alpha, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, h->cstream));
    h->pre_alpha = alpha;
  }
  if (l_ok) {
    if (h->ldk == K)
      CK(cudaMemcpyAsync(Li, h->Li, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToHost, h->cstream));
    else
      CK(cudaMemcpy2DAsync(Li, (size_t)K * sizeof(double), h->Li, h->ldk * sizeof(double), (size_t)K * sizeof(double),
                           K, cudaMemcpyDeviceToHost, h->cstream));
    h->pre_Li = Li;
  }
  return SCFGP_OK;
}

int scfgp_set_chunk_rows(scfgp_handle* h, int rows) {
  if (!h || rows < 128 || rows % 128 != 0) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  { int rcu = finish_upload(h); if (rcu) return rcu; }
  CK(cudaDeviceSynchronize());
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
  { int rcu = finish_upload(h); if (rcu) return rcu; }
  CK(cudaDeviceSynchronize());
  const int o_prec = h->prec, o_ks = h->ks, o_ks2 = h->ks2, o_R = h->R;
  h->prec = mode;
  h->ks = digits;
  h->ks2 = digits2;
  if (!h->R_user) h->R = mode == SCFGP_PREC_EMU_I8 ? kEmuChunkRows : 4096;
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
  { int rcu = finish_upload(h); if (rcu) return rcu; }
  CK(cudaDeviceSynchronize());
  h->ns = n;
  return alloc_chunk_buffers(h);
}

static int set_data_common(scfgp_handle* h, long long N, long long n_global) {
  h->N = N;
  h->Nglobal = n_global > 0 ? n_global : N;
  h->fullX = h->dX; h->fully = h->dy; h->fullN = N; h->fullNglobal = h->Nglobal;
  h->subset_active = false;
  h->have_factors = false;
  h->xvec = (((uintptr_t)h->dX % 16) == 0 && h->D % 2 == 0) ? 2 : 1;
  return ensure_row_arrays(h, N);
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
  // The copies are queued by the first pass of the next evaluation (submit_upload), i.e. BEHIND that evaluation's
  // parameter upload: the host-to-device copy engine serves its queue in order, and 5 GB of row blocks queued here
  // would hold the (small, pageable) parameter copy -- and the host thread that waits for it before it can launch
  // pass 1 -- until the whole table has landed (measured: the upload was then fully exposed, +80 ms per step).
  h->up_X = X;
  h->up_y = y;
  h->up_submitted = false;
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

int scfgp_select_rows(scfgp_handle* h, const long long* idx, long long n, long long n_global) {
  if (!h || n < 0 || (n > 0 && !idx)) return SCFGP_ERR_ARG;
  if (!h->fullX || h->fullN <= 0) { h->err = "select_rows before set_data"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  int rc;
  if ((rc = finish_upload(h))) return rc;
  h->have_factors = false;
  if (idx == nullptr) {                       // back to the whole resident set
    h->dX = h->fullX; h->dy = h->fully; h->N = h->fullN; h->Nglobal = h->fullNglobal;
    h->subset_active = false;
    return SCFGP_OK;
  }
  for (long long i = 0; i < n; ++i)
    if (idx[i] < 0 || idx[i] >= h->fullN) { h->err = "select_rows: row index out of range"; return SCFGP_ERR_ARG; }
  if (n > h->sel_cap || h->selX == nullptr) {
    CK(cudaStreamSynchronize(h->stream));
    for (void* p : {(void*)h->selX, (void*)h->sely, (void*)h->selidx}) if (p) cudaFree(p);
    h->selX = h->sely = nullptr; h->selidx = nullptr; h->sel_cap = 0;
    const long long cap = std::max<long long>(n, 1024);
    CK(cudaMalloc((void**)&h->selX, (size_t)cap * h->D * sizeof(double)));
    CK(cudaMalloc((void**)&h->sely, (size_t)cap * sizeof(double)));
    CK(cudaMalloc((void**)&h->selidx, (size_t)cap * sizeof(long long)));
    h->sel_cap = cap;
  }
  if (n > 0) {
    // (stream-ordered behind the previous evaluation, which may still read the old subset)
    CK(cudaMemcpyAsync(h->selidx, idx, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    k_gather_rows<<<cdiv(n, 8), 256, 0, h->stream>>>(h->fullX, h->fully, h->D, h->selidx, n, h->selX, h->sely);
    CKL();
    CK(cudaStreamSynchronize(h->stream));     // idx may be a temporary of the caller
  }
  h->dX = h->selX; h->dy = h->sely;
  h->N = n;
  h->Nglobal = n_global > 0 ? n_global : n;
  h->subset_active = true;
  h->xvec = (((uintptr_t)h->dX % 16) == 0 && h->D % 2 == 0) ? 2 : 1;
  return ensure_row_arrays(h, std::max<long long>(n, 1));
}

// ---------------------------------------------------------------------------------
int scfgp_eval_begin(scfgp_handle* h, const double* params, int want_grad) {
  if (!h || !params) return SCFGP_ERR_ARG;
  // (N == 0 is legal for a rank whose shard holds no row of the current minibatch: its partial sums are zero)
  if (!h->dX || h->N < 0 || (h->N == 0 && !h->subset_active)) { h->err = "no training data set"; return SCFGP_ERR_ARG; }
  CK(cudaSetDevice(h->device));
  h->want_grad = want_grad;
  h->have_factors = false;
  h->pivot = -1;
  h->pre_alpha = h->pre_Li = nullptr;
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
  if ((rc = submit_upload(h))) return rc;   // (queued behind this evaluation's parameter upload)
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
    if (h->prec == SCFGP_PREC_EMU_I8) {
      // row digits of Phi -- into the cache (all rows stay resident for passes 2 and 3) or the chunk
      // buffer -- and p = PhiT.y; the Gram reads the digits MN-major
      const bool cached = h->phi_cache != nullptr;
      int8_t* dst = cached ? h->phi_cache + r0 * h->Kc : c.digR;
      const long long pitch = cached ? h->Kc : 2 * h->Kp;
      const long long plane = cached ? h->cache_rows * h->Kc : (long long)h->R * 2 * h->Kp;
      if (fused_features_ok(h)) {
        // one kernel: X.W, sincos, digits, p -- the float64 Phi slab is never written
        if ((rc = run_features_digits(h, c, h->dX + r0 * h->D, h->xvec, rows, dst, pitch, plane, h->dy + r0, c.colpart)))
          return rc;
      } else {
        if ((rc = run_features(h, c, h->dX + r0 * h->D, h->xvec, rows, nullptr, nullptr))) return rc;
        if ((rc = emu_digits_slab(h, c, h->ks, false, r0, rows, nullptr, nullptr, dst, nullptr, nullptr, nullptr,
                                  pitch, plane, h->dy + r0, c.colpart)))
          return rc;
      }
      if ((rc = emu_gram(h, c, false, rows, r0, cached))) return rc;
    } else {
      if ((rc = run_features(h, c, h->dX + r0 * h->D, h->xvec, rows, nullptr, nullptr))) return rc;
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
  h->up_X = h->up_y = nullptr;
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
                                                      1.0, 0.0, h->lowerK_tri, K)))
      return rc;
    k_symmetrize<<<dim3(cdiv(K, 32), cdiv(K, 32)), dim3(32, 32), 0, h->stream>>>(h->Bm, h->ldk, K);
    CKL();
  }
  long long info = 0;
  CK(cudaMemcpyAsync(&info, h->info, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(h->ev[7], h->stream));
  CK(cudaEventRecord(h->ev_factor, h->stream));
  // a-posteriori check of the emulated Gram on a row sample (host_guard.cuh); read back with the pivot flag
  double g5[5] = {0, 0, 0, 0, 0};
  if ((rc = guard_measure(h))) return rc;
  if (h->guard_sample > 0) CK(cudaMemcpyAsync(g5, h->scal + 8, sizeof(g5), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  guard_finish(h, g5);
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
  // BHB is symmetric in exact arithmetic, but it must be computed WHOLE: the rounding error of (BH).B is delta.B, and
  // everything downstream sees it through products like tr(delta.B.G) with B.G ~ I.  Mirroring the lower triangle
  // keeps |delta.B| without that cancellation: at cond(A) ~ 4e7 the a/b/c gradient went from 1e-10 to 1e-4 of the
  // oracle's and l_F from 2e-9 to 3e-6 (tools/envelope.py, round 2) -- for 0.3 ms saved.  (B = LiT.Li, G and H are
  // Gram matrices: their two triangles are the same sums in the same order, so mirroring those is exact.)
  if ((rc = launch_gemm<CfgBig, true, true, 2, 2>(h, h->BH, h->ldk, K, h->Bm, h->ldk, K, h->BHB, h->ldk, 1.0,
                                                  0.0, h->fullK, K)))
    return rc;
  k_abar<<<K, 256, 0, h->stream>>>(h->Bm, h->BHB, h->ldk, K, h->tvec, h->alpha, h->sig2n, h->A2, h->B2,
                                   h->pbar, h->trslots);
  CKL();
  k_reduce_slots<<<1, 256, 0, h->stream>>>(h->trslots, K, 1, 1, h->scal + 3);
  CKL();
  h->sbar_from_traces = emu_two_phase_backward(h);
  if (h->sbar_from_traces) {
    // sbar = sum <Phibar, Phi> = tr(2 Abar G) + tr(2 B H) + pbarT p + alphaT q  (the pass-3 kernel no longer sums it)
    k_sbar_rows<<<K, 128, 0, h->stream>>>(h->A2, h->B2, h->red1, H, h->ldk, K, h->pbar, h->red1 + KK, h->alpha, q,
                                         h->trslots2);
    CKL();
    k_reduce_slots<<<1, 256, 0, h->stream>>>(h->trslots2, K, 1, 1, h->scal + 5);
    CKL();
  }
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
    // dW partials += XcT . Zbar, the row reduction split over kDwSplit CTAs per tile; dphi = colsum(Zbar) is summed
    // from the same operand fragments (except after the one-phase emulated backward kernel, which sums it itself)
    {
      const int split_len = cdiv(cdiv(rows, kDwSplit), 16) * 16;
      const dim3 grid(h->dw.count, kDwSplit);
      const size_t dsm = SmemSize<CfgDw, false, false>::BYTES;
      double* colp = (emu_on && !emu_two_phase_backward(h)) ? nullptr : c.colpart;
      trace_begin(h, "dw", c.st);
      if (h->xvec == 2)
        k_dw_splitk<2><<<grid, CfgDw::NT, dsm, c.st>>>(Xc, D, D, c.Zbar, h->ldj, J, rows, split_len, c.dW, h->ldj,
                                                         (long long)DJ, h->dw.d, colp, (long long)K);
      else
        k_dw_splitk<1><<<grid, CfgDw::NT, dsm, c.st>>>(Xc, D, D, c.Zbar, h->ldj, J, rows, split_len, c.dW, h->ldj,
                                                         (long long)DJ, h->dw.d, colp, (long long)K);
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
  double sc[8] = {0};       // logdetL, beta.beta, p.alpha, tr(Abar), -, sbar (two-phase emulated backward)
  double r1[1], r2[2];
  CK(cudaMemcpyAsync(sc, h->scal, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(r1, h->red1 + KK + K, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(r2, h->red2 + KK + K, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  std::vector<double> g3;
  if (grad) {
    if (!h->want_grad) { h->err = "gradient requested but eval_begin had want_grad=0"; return SCFGP_ERR_ARG; }
    g3.resize((size_t)D * h->ldj + h->ldj + 1);
    CK(cudaMemcpyAsync(g3.data(), h->red3, g3.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  // (alpha / Li may already be on their way on the copy stream: scfgp_eval_prefetch_factors)
  if (alpha && alpha != h->pre_alpha)
    CK(cudaMemcpyAsync(alpha, h->alpha, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (Li && Li != h->pre_Li) {
    if (h->ldk == K)
      CK(cudaMemcpyAsync(Li, h->Li, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    else
      CK(cudaMemcpy2DAsync(Li, (size_t)K * sizeof(double), h->Li, h->ldk * sizeof(double),
                           (size_t)K * sizeof(double), K, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  const auto t_host0 = std::chrono::steady_clock::now();
  if (h->pre_alpha || h->pre_Li) CK(cudaStreamSynchronize(h->cstream));
  const auto t_host1 = std::chrono::steady_clock::now();
  h->pre_alpha = h->pre_Li = nullptr;
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
    // = sum(Phibar * Phi) = d nlml / d b: summed by the pass-3 kernels, or from the two traces (k_sbar_rows)
    const double sbar_s = h->sbar_from_traces ? sc[5] : dphi[h->ldj];
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
    // (contiguous inner loops over the S latent columns: the same summation order as the textbook triple loops --
    //  m ascending / d ascending per output entry -- at a fraction of their 3 ms of host time)
    std::vector<double> accv(S);
    for (int d = 0; d < D; ++d) {
      for (int s2 = 0; s2 < S; ++s2) accv[s2] = Wl[(size_t)d * S + s2];
      const double* wf = &WF[(size_t)d * M];
      for (int m = 0; m < M; ++m) {
        const double w = wf[m];
        const double* rfm = &h->rF[(size_t)m * S];
        for (int s2 = 0; s2 < S; ++s2) accv[s2] += w * rfm[s2];
      }
      for (int s2 = 0; s2 < S; ++s2) g_lf[(size_t)d * S + s2] = accv[s2] / Ng;
    }
    for (size_t i = 0; i < (size_t)M * S; ++i) g_rf[i] = 0.0;
    for (int d = 0; d < D; ++d) {
      const double* wf = &WF[(size_t)d * M];
      const double* lfd = &h->lF[(size_t)d * S];
      for (int m = 0; m < M; ++m) {
        const double w = wf[m];
        double* gr = g_rf + (size_t)m * S;
        for (int s2 = 0; s2 < S; ++s2) gr[s2] += w * lfd[s2];
      }
    }
    for (size_t i = 0; i < (size_t)M * S; ++i) g_rf[i] /= Ng;
    for (int s2 = 0; s2 < S; ++s2) g_lp[s2] = phase ? dphi[s2] / Ng : 0.0;
    for (int m = 0; m < M; ++m) g_p[m] = phase ? dphi[S + m] / Ng : 0.0;
  }
  if (getenv("SCFGP_HOST_TIMING")) {
    const auto t_host2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[scfgp_b200] finish: wait for the factor copy %.3f ms, host math after the last sync %.3f ms\n",
            std::chrono::duration<double, std::milli>(t_host1 - t_host0).count(),
            std::chrono::duration<double, std::milli>(t_host2 - t_host1).count());
  }
  return SCFGP_OK;
}

int scfgp_guard_escalate(scfgp_handle* h) {
  if (!h) return SCFGP_ERR_ARG;
  if (h->prec != SCFGP_PREC_EMU_I8) return SCFGP_OK;
  h->guard_escalations++;
  // 5 (or fewer) digits -> 6 digits -> FP64 DMMA; sticky
  if (std::max(h->ks, h->ks2) < 6 && std::min(h->ks, h->ks2) < 6) return scfgp_set_precision2(h, SCFGP_PREC_EMU_I8, 6, 6);
  return scfgp_set_precision2(h, SCFGP_PREC_FP64, 0, 0);
}

int scfgp_eval(scfgp_handle* h, const double* params, double* cost, double* grad, double* alpha, double* Li) {
  int rc;
  for (;;) {
    if ((rc = scfgp_eval_begin(h, params, grad != nullptr))) return rc;
    if ((rc = scfgp_eval_pass1(h))) return rc;
    if ((rc = scfgp_eval_factor(h))) return rc;
    if (!guard_exceeded(h)) break;
    // the emulated Gram is not accurate enough at these hyper-parameters: more digits, start over
    if ((rc = scfgp_guard_escalate(h))) return rc;
  }
  if ((rc = scfgp_eval_prefetch_factors(h, alpha, Li))) return rc;
  if ((rc = scfgp_eval_pass2(h))) return rc;
  if (grad) {
    if ((rc = scfgp_eval_backward(h))) return rc;
    if ((rc = scfgp_eval_pass3(h))) return rc;
  }
  return scfgp_eval_finish(h, cost, grad, alpha, Li);
}

long long scfgp_phi_cache_bytes(const scfgp_handle* h) {
  return (h && h->phi_cache) ? (long long)h->cache_ks * h->cache_rows * h->Kc : 0;
}

int scfgp_set_phi_cache(scfgp_handle* h, int on, double max_gb) {
  if (!h) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  { int rcu = finish_upload(h); if (rcu) return rcu; }
  CK(cudaDeviceSynchronize());
  h->want_cache = on != 0;
  if (max_gb > 0) h->cache_max_gb = max_gb;
  free_phi_cache(h);   // re-decided by the next evaluation
  return SCFGP_OK;
}

int scfgp_set_guard(scfgp_handle* h, int on, double cost_tol, double alpha_tol) {
  if (!h) return SCFGP_ERR_ARG;
  h->guard_on = on != 0;
  if (cost_tol > 0) h->guard_cost_tol = cost_tol;
  if (alpha_tol > 0) h->guard_alpha_tol = alpha_tol;
  return SCFGP_OK;
}

int scfgp_guard_last(const scfgp_handle* h, double* est_cost, double* est_alpha, double* dg_fro, long long* sample_rows,
                     int* exceeded, int* escalations) {
  if (!h) return SCFGP_ERR_ARG;
  if (est_cost) *est_cost = h->guard_est_cost;
  if (est_alpha) *est_alpha = h->guard_est_alpha;
  if (dg_fro) *dg_fro = h->guard_dg_fro;
  if (sample_rows) *sample_rows = h->guard_sample;
  if (exceeded) *exceeded = guard_exceeded(h) ? 1 : 0;
  if (escalations) *escalations = h->guard_escalations;
  return SCFGP_OK;
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
  if (h->prec == SCFGP_PREC_EMU_I8 && (rc = emu_refresh_liscale_w(h))) return rc;   // (s2 depends on the parameters)
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

int scfgp_set_factors(scfgp_handle* h, const double* alpha, const double* Li) {
  if (!h || !alpha || !Li) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const int K = h->K;
  CK(cudaMemcpyAsync(h->alpha, alpha, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpy2DAsync(h->Li, h->ldk * sizeof(double), Li, (size_t)K * sizeof(double), (size_t)K * sizeof(double), K,
                       cudaMemcpyHostToDevice, h->stream));
  h->have_factors = true;
  int rc;
  if (h->prec == SCFGP_PREC_EMU_I8 && (rc = emu_prepare_li(h))) return rc;
  CK(cudaStreamSynchronize(h->stream));   // the host arrays may go away once this returns
  return SCFGP_OK;
}

int scfgp_predict(scfgp_handle* h, const double* params, const double* Xs, long long T, const double* alpha,
                  const double* Li, double* mu, double* sd) {
  if (!h || !params || !Xs || !mu || !sd || T <= 0) return SCFGP_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const int K = h->K;
  if ((alpha == nullptr) != (Li == nullptr)) { h->err = "pass both alpha and Li or neither"; return SCFGP_ERR_ARG; }
  if (alpha) {
    int rcl = scfgp_set_factors(h, alpha, Li);
    if (rcl) return rcl;
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

int scfgp_dbg_mma_stream_peak(int device, int digits, int tile_n, int steps, int random_data, double* tops,
                              double* clk_per_step, double* sm_mhz) {
  if (steps <= 0) return SCFGP_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return SCFGP_ERR_CUDA;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return SCFGP_ERR_CUDA;
  if (digits == 5 && tile_n == 96) return mma_stream_peak<5, 96>(sms, steps, random_data, tops, clk_per_step, sm_mhz);
  if (digits == 5 && tile_n == 64) return mma_stream_peak<5, 64>(sms, steps, random_data, tops, clk_per_step, sm_mhz);
  if (digits == 6 && tile_n == 64) return mma_stream_peak<6, 64>(sms, steps, random_data, tops, clk_per_step, sm_mhz);
  if (digits == 4 && tile_n == 64) return mma_stream_peak<4, 64>(sms, steps, random_data, tops, clk_per_step, sm_mhz);
  return SCFGP_ERR_ARG;
}

int scfgp_dbg_tile_timing(int device, int enable, unsigned long long* out, int max_ctas) {
  // developer tool: clock64() intervals of the tcgen05 kernels' roles per CTA (umma_i8.cuh, g_tile_timing)
  static unsigned long long* buf[64] = {};
  if (device < 0 || device >= 64 || max_ctas <= 0) return SCFGP_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return SCFGP_ERR_CUDA;
  if (cudaDeviceSynchronize() != cudaSuccess) return SCFGP_ERR_CUDA;
  const size_t bytes = (size_t)max_ctas * umma::TT_SLOTS * sizeof(unsigned long long);
  unsigned long long* nul = nullptr;
  if (buf[device]) {
    if (out && cudaMemcpy(out, buf[device], bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return SCFGP_ERR_CUDA;
    cudaMemcpyToSymbol(umma::g_tile_timing, &nul, sizeof(nul));
    cudaFree(buf[device]);
    buf[device] = nullptr;
  }
  if (enable) {
    if (cudaMalloc((void**)&buf[device], bytes) != cudaSuccess) return SCFGP_ERR_CUDA;
    cudaMemset(buf[device], 0, bytes);
    if (cudaMemcpyToSymbol(umma::g_tile_timing, &buf[device], sizeof(nul)) != cudaSuccess) return SCFGP_ERR_CUDA;
  }
  return SCFGP_OK;
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
