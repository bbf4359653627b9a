// Host-side state of one handle: constants, per-stream chunk context, the handle itself.
// (part of the single translation unit scfgp_abi.cu; included in this order)
#pragma once

namespace {

constexpr double kEps = 1e-6;  // jitter, SCFGP/SCFGP.py:93
constexpr int kColSplit = 64;  // row splits of the column-sum kernels (parallelism: ceil(K/256) x 64 CTAs)
constexpr int kSeedBlocks = 296;
constexpr int kGramSplit = 1;  // row splits of one chunk's Gram tiles in the emulated mode (work items = tiles x splits)
constexpr int kWideTile = 96;  // default tile width of Phi.LiT / backward in the emulated mode with 5 digits (64 otherwise)
// rows per chunk of the emulated mode: the reduction depth of the Gram products (exact in int32 up to 18724 rows at
// 8 digits, 32767 at 5); 16384 halves the Gram tiles' epilogues and pipeline restarts against 8192 (-0.9 % per evaluation)
constexpr int kEmuChunkRows = 16384;
constexpr int kDwSplit = 32;   // row splits of the dW = XT.Zbar contraction (partial buffers per chunk stream)

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
  // minibatch mode (SCFGP.py:172-182,226-235): a row subset of the resident set, gathered on the device
  const double* fullX = nullptr;         // the resident training set the subset is taken from
  const double* fully = nullptr;
  long long fullN = 0, fullNglobal = 0;
  double *selX = nullptr, *sely = nullptr;
  long long* selidx = nullptr;
  long long sel_cap = 0;
  bool subset_active = false;
  // host -> device upload in row blocks on a copy stream, overlapped with the first pass
  cudaStream_t cstream = nullptr;
  std::vector<cudaEvent_t> up_ev;        // one event per uploaded row block
  long long up_block = 0;                // rows per block
  bool upload_pending = false;
  const double *up_X = nullptr, *up_y = nullptr;   // host arrays of a pending upload (scfgp_set_data_host_async)
  bool up_submitted = false;             // its copies have been queued on the copy stream
  int R = 4096;  // rows per chunk (kEmuChunkRows in the emulated mode unless set by the caller)
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
  double *liscale_w = nullptr;  // liscale * (every constant factor of the Phi.LiT epilogue)
  double *liscale = nullptr, *liinv = nullptr, *cscale = nullptr, *cinv = nullptr, *sigg = nullptr;
  unsigned long long* gvmax = nullptr;
  CUtensorMap mapLiB, mapCB;
  // optional cache of Phi's row digits for ALL rows ([ks][cache_rows][Kc]; 104 GB at cfg4): passes 2 and 3
  // then read Phi from it instead of recomputing features + digits
  int8_t* phi_cache = nullptr;
  long long cache_rows = 0, Kc = 0;
  int cache_ks = 0;
  int want_cache = 1;          // SCFGP_PHI_CACHE=0 disables
  double cache_max_gb = 120.0; // SCFGP_PHI_CACHE_MAX_GB
  int cache_log = 1;           // SCFGP_PHI_CACHE_LOG: 0 silent, 1 say when declined (default), 2 also when engaged
  CUtensorMap mapCacheK, mapCacheMN_A, mapCacheMN_B;
  int4* gram_items = nullptr;
  int gram_count = 0;
  int vparts = 0;              // partial rows of vpart per chunk
  int fuse_features = 1;       // pass 1 (5 digits, J % 4 == 0): digits straight from the feature kernel; SCFGP_FUSED_FEATURES=0: two kernels
  int fast_digits = 1;         // 16-byte-wide digit kernel for the cached 5-digit path; SCFGP_DIGIT_KERNEL=generic: k_digits_slab
  int two_phase = 1;           // two-phase epilogue of Phi.LiT / backward (up to 5 digits); SCFGP_EPILOGUE=1: one-phase
  int tnw = 64;                // tile width of the K-major digit products (Phi.LiT, backward): 64, or 96 (5 digits)
  int gsplit = 1;              // partial Gram accumulators per chunk stream (kGramSplit in the emulated mode)
  int csplit = kColSplit;      // partial rows of the column-sum buffers per chunk stream (>= R/128)

  // work lists
  ItemList lowerK, syrkK, fullK, tgemm, bwd, dw;
  ItemList lowerK_tri;   // the same tiles with the reduction starting at the tile's row block (B = LiT.Li: Li is lower triangular)
  std::vector<ItemList> potrf_panel, potrf_trail;
  std::vector<ItemList> inv_t, inv_x;   // per level of the block-recursive triangular inverse: T = L21.Li11, Li21 = -Li22.T
  std::vector<int4*> to_free;

  // host state of the evaluation in flight
  std::vector<double> params, lF, rF, F, W, phi_h;
  double a = 0, b = 0, c = 0, sig2n = 0, sigf = 0, tau = 0, s = 0;
  int want_grad = 0;
  bool have_factors = false;
  bool sbar_from_traces = false;   // this evaluation's d nlml / d b comes from k_sbar_rows (scal[5])

  std::string err;
  long long pivot = -1;
  long long launches = 0;
  cudaEvent_t ev[8] = {};
  float ms_pass[4] = {0, 0, 0, 0};
  // stream hand-off with a caller's stream (NCCL all-reduce between phases) and the early copy-out of
  // alpha / Li on the copy stream while passes 2 and 3 run
  cudaEvent_t ev_out = nullptr, ev_in = nullptr, ev_factor = nullptr;
  double *pre_alpha = nullptr, *pre_Li = nullptr;   // host destinations already being filled
  // a-posteriori guard of the emulated path (host_guard.cuh)
  int guard_on = 1;
  double guard_cost_tol = 1e-7, guard_alpha_tol = 1e-7;
  long long guard_sample = 0;
  double guard_est_cost = 0, guard_est_alpha = 0, guard_dg_fro = 0;
  int guard_escalations = 0;      // since the handle was created
  double* trslots2 = nullptr;     // 2 K scratch

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
