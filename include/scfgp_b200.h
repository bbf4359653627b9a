/*
 * scfgp_b200 -- C ABI of the B200-native SCFGP marginal-likelihood hot path.
 *
 * Drop-in boundary: the reference's hot path is the triple of compiled Theano
 * callables produced by SCFGP.build_theano_models (SCFGP/SCFGP.py:92-148) and handed
 * around by get_compiled_funcs() / optimize(funcs=...) (SCFGP/SCFGP.py:150-151,184,
 * 209-210):
 *     train_func(X, y)            -> [cost, alpha, Li]          SCFGP/SCFGP.py:132-135
 *     train_iter_func(X, y)       -> same + optimizer update    SCFGP/SCFGP.py:136-137
 *     pred_func(Xs, alpha, Li)    -> [mu_pred, std_pred]        SCFGP/SCFGP.py:145-148
 * The entry points below are what a ctypes/cffi binding of those callables binds
 * (see INTEGRATION.md).  Plain pointers and sizes only; all floating point is IEEE
 * binary64 ("double"), matrices are row-major, exactly as the NumPy arrays the
 * reference passes.  One deliberate deviation: the hyper-parameter vector is an
 * explicit argument (the reference captures a Theano shared variable through
 * `givens`, SCFGP/SCFGP.py:135,137,148).
 *
 * Threading: one handle = one caller thread.  No hidden global state.
 * There is no CPU fallback: every call needs a CUDA device (sm_100a).
 */
#ifndef SCFGP_B200_H
#define SCFGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scfgp_handle scfgp_handle;

/* status codes (the reference raises: numpy.linalg.LinAlgError from Theano's
 * Cholesky thunk at SCFGP/SCFGP.py:106, shape errors from theano.function) */
enum {
  SCFGP_OK = 0,
  SCFGP_ERR_ARG = 1,         /* bad shape / null pointer / no data set            */
  SCFGP_ERR_CUDA = 2,        /* CUDA runtime error, text in scfgp_last_error()    */
  SCFGP_ERR_NOT_PD = 3,      /* A = PhiT.Phi + (sig2_n+1e-6) I not positive definite;
                                pivot index in scfgp_last_pivot()                 */
  SCFGP_ERR_NONFINITE = 4    /* cost is NaN/Inf                                   */
};

/* fftype: the README-generation constructor argument (README.md:44-53); the snapshot
 * source implements only the phase variant (SCFGP/SCFGP.py:83-89). */
enum {
  SCFGP_FF_PHASE = 0,        /* Z = X.W + (P - mean_d(F))   (SCFGP/SCFGP.py:88-89,99-100) */
  SCFGP_FF_NOPHASE = 1       /* Z = X.W                                                */
};

/* ABI version of this header */
int scfgp_abi_version(void);

/* Create a model-shaped handle on CUDA device `device`.
 *   D = input columns, S = sparsity/rank, M = nfeats/feature_size
 *   (SCFGP.__init__, SCFGP/SCFGP.py:33-37; D from set_data, SCFGP/SCFGP.py:165). */
int scfgp_create(int D, int S, int M, int fftype, int device, scfgp_handle** out);
int scfgp_destroy(scfgp_handle* h);

/* Last error text / failing Cholesky pivot (0-based row of A), per handle. */
const char* scfgp_last_error(const scfgp_handle* h);
long long scfgp_last_pivot(const scfgp_handle* h);

/* Number of hyper-parameters P = 3 + D*S + M*S + S + M (SCFGP.init_params,
 * SCFGP/SCFGP.py:64-72) and of features K = 2*(M+S) (SCFGP/SCFGP.py:99-102). */
long long scfgp_num_params(const scfgp_handle* h);
long long scfgp_num_features(const scfgp_handle* h);

/* Run all work of this handle on `cuda_stream` (a cudaStream_t; NULL = default). */
int scfgp_set_stream(scfgp_handle* h, void* cuda_stream);

/* Rows per streamed chunk (multiple of 128; default 4096).  Tuning knob only. */
int scfgp_set_chunk_rows(scfgp_handle* h, int rows);

/* Number of CUDA streams consecutive row chunks are spread over (1..4; default 4, or the
 * SCFGP_STREAMS environment variable at create time).  Tuning knob only: results do not
 * depend on timing, chunk c always accumulates into context c % n. */
int scfgp_set_streams(scfgp_handle* h, int n);

/* Arithmetic of the three streaming products (Gram, Phi.LiT, backward).  The reference
 * computes in float64 (TT.dmatrices, SCFGP/SCFGP.py:95-96); both modes return float64 and
 * meet the stated tolerances (cost 1e-6, gradient/alpha/Li/predictions 1e-5).
 *   SCFGP_PREC_FP64   : FP64 tensor path (DMMA.8x8x4), the default.
 *   SCFGP_PREC_EMU_I8 : float64 emulated on the int8 tensor cores (tcgen05.mma kind::i8,
 *                       exact int32 accumulation in tensor memory).  Operands are cut into
 *                       `digits` signed base-256 digits (4..8; 0 = default 5, i.e. 38
 *                       fractional bits of each operand row's power-of-two scale).  Feature map, Cholesky,
 *                       triangular inverse and the small dense backward stay in true FP64.
 * Default at create time: environment SCFGP_PRECISION = "fp64" | "emu" | "emu<digits>". */
enum { SCFGP_PREC_FP64 = 0, SCFGP_PREC_EMU_I8 = 1 };
int scfgp_set_precision(scfgp_handle* h, int mode, int digits);
int scfgp_get_precision(const scfgp_handle* h, int* mode, int* digits);
/* Two digit counts: `digits` for PhiT.Phi (pass 1: it feeds the Cholesky factor, alpha and Li, the
 * results most sensitive to operand rounding) and `digits2` (0 = same) for Phi.LiT, PsiT.Phi and the
 * backward product, whose rounding only reaches the cost through v_n and the gradient.
 * Environment spelling: SCFGP_PRECISION=emu5/4. */
int scfgp_set_precision2(scfgp_handle* h, int mode, int digits, int digits2);
int scfgp_get_precision2(const scfgp_handle* h, int* mode, int* digits, int* digits2);

/* Training data of train_func/train_iter_func (the X, y arguments at
 * SCFGP/SCFGP.py:230,237,265).  X is N x D row-major, y has N entries.
 *   _host  : copies from host memory into device buffers owned by the handle.
 *   _device: borrows device pointers (zero-copy hand-off, e.g. torch data_ptr());
 *            the caller keeps them alive until the next set_data/destroy.
 * `n_global` is the row count of the whole job (sum over ranks); pass N when
 * single-GPU.  It is the N of `cost = (nlml+penalty)/N` and `2(N-M)a`
 * (SCFGP/SCFGP.py:126,128). */
int scfgp_set_data_host(scfgp_handle* h, const double* X, const double* y, long long N,
                        long long n_global);
int scfgp_set_data_device(scfgp_handle* h, const double* dX, const double* dy, long long N,
                          long long n_global);
/* Like scfgp_set_data_host, but returns at once: X is uploaded in row blocks on a copy stream and
 * the first pass of the next evaluation starts on each block as it lands, so the transfer hides
 * behind compute (use page-locked host memory; pageable memory works but is staged by the driver).
 * The copies are queued by that evaluation's first pass, behind its parameter upload (the
 * host-to-device copy engine serves its queue in order), or by the next call that needs the whole
 * table on the device.  X and y must stay valid and unchanged until that evaluation's scfgp_eval /
 * scfgp_eval_pass1 has returned. */
int scfgp_set_data_host_async(scfgp_handle* h, const double* X, const double* y, long long N,
                              long long n_global);

/* train_func(X, y) with explicit hyper-parameters, plus the gradient the reference
 * consumes in-graph (TT.grad, SCFGP/SCFGP.py:129).
 *   params : P doubles (host)
 *   cost   : 1 double out
 *   grad   : P doubles out, or NULL (then only the forward passes run)
 *   alpha  : K doubles out, or NULL          (K x 1 in the reference)
 *   Li     : K*K doubles out row-major, or NULL  (inverse lower Cholesky factor)
 * Single-GPU convenience: runs every phase below in order. */
int scfgp_eval(scfgp_handle* h, const double* params, double* cost, double* grad,
               double* alpha, double* Li);

/* pred_func(Xs, alpha, Li) -> mu_pred (T), std_pred (T)  (SCFGP/SCFGP.py:138-148).
 * Xs is T x D row-major on the host; alpha/Li as returned by scfgp_eval, or NULL to
 * use the factors of the last evaluation still resident on the device. */
int scfgp_predict(scfgp_handle* h, const double* params, const double* Xs, long long T,
                  const double* alpha, const double* Li, double* mu, double* std);
/* alpha (K) and Li (K x K row-major) from the host become the device-resident factors that
 * scfgp_predict(.., NULL, NULL, ..) and scfgp_predict_device use: the (alpha, Li) arguments of the reference's
 * pred_func (SCFGP/SCFGP.py:145-148) for callers whose test rows already live on the device, e.g. a model
 * restored from a checkpoint (SCFGP.py:303-310) predicting in row shards. */
int scfgp_set_factors(scfgp_handle* h, const double* alpha, const double* Li);
/* same with device-resident inputs/outputs (batched predict over shards) */
int scfgp_predict_device(scfgp_handle* h, const double* params, const double* dXs, long long T,
                         double* dmu, double* dstd);

/* Minibatch mode (SCFGP.minibatches / optimize(nbatches > 1), SCFGP/SCFGP.py:172-182,226-235): evaluate on the rows
 * idx[0..n) (host array of LOCAL row indices into the resident training set of this handle) -- gathered on the
 * device, the training set is not uploaded again.  n_global = rows of the batch over all ranks (the N of the
 * reference's cost for this batch; 0 = n).  n = 0 with idx != NULL is legal (a rank without rows of this batch).
 * idx == NULL returns to the whole resident set.  A new scfgp_set_data_* call also ends the selection. */
int scfgp_select_rows(scfgp_handle* h, const long long* idx, long long n, long long n_global);

/* ---- phase API (multi-GPU: rows are sharded over ranks, the host all-reduces the
 * device buffer named by each phase between calls; every rank then runs the small
 * replicated phases on identical inputs).  Order:
 *   begin -> pass1 -> [allreduce RED1] -> factor -> pass2 -> [allreduce RED2]
 *         -> backward -> pass3 -> [allreduce RED3] -> finish
 * (forward only: begin -> pass1 -> factor -> pass2 -> finish with grad==NULL). */
enum { SCFGP_RED1 = 1, SCFGP_RED2 = 2, SCFGP_RED3 = 3 };
int scfgp_eval_begin(scfgp_handle* h, const double* params, int want_grad);
int scfgp_eval_pass1(scfgp_handle* h);     /* G = PhiT.Phi, p = PhiT.y, yy          */
int scfgp_eval_factor(scfgp_handle* h);    /* L, Li, beta, alpha, B = A^-1          */
int scfgp_eval_pass2(scfgp_handle* h);     /* v, mu, E, g_tau, q, H                 */
int scfgp_eval_backward(scfgp_handle* h);  /* Abar, pbar (small, replicated)        */
int scfgp_eval_pass3(scfgp_handle* h);     /* dW, dphi, sbar                        */
int scfgp_eval_finish(scfgp_handle* h, double* cost, double* grad, double* alpha, double* Li);
/* Device pointer + element count (doubles) of the contiguous buffer a phase reduces. */
int scfgp_reduce_buffer(scfgp_handle* h, int which, void** dptr, long long* count);
/* Stream-ordered hand-off between the handle's stream and the stream the caller runs its collective on
 * (no host synchronisation).  SCFGP_HANDOFF_TO_OTHER: `other_stream` waits for everything this handle has
 * queued so far (call it after a pass, before the all-reduce).  SCFGP_HANDOFF_FROM_OTHER: the handle's
 * stream waits for everything queued on `other_stream` so far (call it after the all-reduce). */
enum { SCFGP_HANDOFF_TO_OTHER = 0, SCFGP_HANDOFF_FROM_OTHER = 1 };
int scfgp_stream_handoff(scfgp_handle* h, void* other_stream, int direction);
/* After scfgp_eval_factor: start copying alpha (K) / Li (K x K) to the host destinations on the copy stream
 * while passes 2 and 3 run; scfgp_eval_finish with the same pointers then only waits for that copy.  Only
 * page-locked destinations are filled early (a pageable copy would block the calling thread); others are
 * left to scfgp_eval_finish.  Either pointer may be NULL.  scfgp_eval does this itself. */
int scfgp_eval_prefetch_factors(scfgp_handle* h, double* alpha, double* Li);

/* Digit cache of Phi (emulated mode): pass 1 keeps the int8 digits of Phi for ALL rows of this handle in HBM
 * (digits x ceil128(N) x ceil16(K) bytes: 104 GB at cfg4 on one GPU, 13 GB per GPU on eight) so that passes 2
 * and 3 read them instead of recomputing the feature map.  It is allocated by the first evaluation when it fits
 * under max_gb (default 120, env SCFGP_PHI_CACHE_MAX_GB) and under the free device memory minus 6 GB; when it
 * does not, the handle says so on stderr and runs without it.  on = 0 (env SCFGP_PHI_CACHE=0) never allocates it.
 * scfgp_phi_cache_bytes: bytes currently held (0 = running without the cache). */
int scfgp_set_phi_cache(scfgp_handle* h, int on, double max_gb);
long long scfgp_phi_cache_bytes(const scfgp_handle* h);

/* A-posteriori guard of the emulated arithmetic (on by default in SCFGP_PREC_EMU_I8; SCFGP_GUARD=0 or
 * scfgp_set_guard(h, 0, ..) turns it off).  After every factor phase the handle measures the error of the emulated
 * Gram on a sample of its rows (the same rows through the FP64 DMMA Gram and through the digit Gram) and pushes
 * it through the first-order perturbation of the results with the factors just computed:
 *   est_cost  = estimated absolute error of `cost`      (threshold cost_tol,  default 1e-7)
 *   est_alpha = estimated max error of alpha / max(1, |alpha|_inf)   (threshold alpha_tol, default 1e-7)
 * (north_star's tolerances are 1e-6 relative on the cost and 1e-5 on alpha; the gradient error is up to ~100 x the estimated alpha error at K = 2080 -- profiles/r2_envelope_*.txt -- hence 1e-7).
 * scfgp_eval escalates on its own when an estimate exceeds its threshold: 5 -> 6 digits -> FP64, restarting the
 * evaluation, and the handle stays at the higher precision (scfgp_get_precision2 tells).  With the phase API the
 * caller reads the estimates after scfgp_eval_factor (scfgp_guard_last), combines them over the ranks (max) and
 * calls scfgp_guard_escalate + restarts from scfgp_eval_begin on every rank when they exceed the thresholds.
 * Thresholds <= 0 keep the current value. */
int scfgp_set_guard(scfgp_handle* h, int on, double cost_tol, double alpha_tol);
int scfgp_guard_last(const scfgp_handle* h, double* est_cost, double* est_alpha, double* dg_fro,
                     long long* sample_rows, int* exceeded, int* escalations);
int scfgp_guard_escalate(scfgp_handle* h);

/* Device-resident results of the last evaluation (valid until the next one). */
int scfgp_device_factors(scfgp_handle* h, void** d_alpha, void** d_Li, long long* ld_Li);

/* Counters for bench.py: kernels launched by this handle since creation, and the
 * device time (ms, CUDA events on the handle's stream) of the three streaming passes
 * of the last evaluation. */
long long scfgp_launch_count(const scfgp_handle* h);
int scfgp_last_pass_ms(const scfgp_handle* h, float* pass1, float* pass2, float* pass3,
                       float* dense);

/* Optional launch trace: a CUDA event pair around every streaming-kernel launch (per chunk
 * stream) from scfgp_set_tracing(h, 1) on; scfgp_trace_dump writes them as Chrome-trace JSON
 * (one "thread" per chunk stream; open in chrome://tracing or Perfetto).  Developer tool:
 * the events add a little serialisation, never time a benchmark with tracing on. */
int scfgp_set_tracing(scfgp_handle* h, int on);
int scfgp_trace_dump(scfgp_handle* h, const char* path);

/* Optional kernel-family timing for the roofline report: CUDA events (on the handle's
 * stream) around every launch of ONE family; totals accumulate until the next
 * scfgp_set_profiling call and are collected at scfgp_eval_finish. */
enum {
  SCFGP_FAM_NONE = 0,
  SCFGP_FAM_FEATURES = 1,   /* X.W GEMM + sincos epilogue                      */
  SCFGP_FAM_TGEMM = 2,      /* Phi.LiT with row-sum-of-squares epilogue         */
  SCFGP_FAM_BACKWARD = 3,   /* Phi.(2Abar) + Psi.(2B) with trig-derivative epilogue */
  SCFGP_FAM_SYRK = 4        /* PhiT.Phi / PsiT.Phi Gram accumulation            */
};
int scfgp_set_profiling(scfgp_handle* h, int family);
int scfgp_get_profile(const scfgp_handle* h, double* total_ms, long long* launches);

/* ---- data preprocessing on the device (SCFGP/Scaler.py:39-135).
 * Stateless helpers on device-resident row-major float64 data (N x D, leading dimension ld);
 * the Scaler object itself (fitted fields, algorithm names, SLSQP line search over the
 * Box-Cox exponent) is host code: scfgp_b200/gpu_scaler.py.
 *   colstats       : per-column min, max and "constant column" flags        Scaler.py:40-41,44-45
 *   gather_column  : dcol[r] = (X[r][col] - shift) / range (contiguous copy of one column)
 *   column_moments : mean, population variance and skewness g1 = m3 / m2^1.5 of
 *                    boxcox(dcol, lambda) (lambda <= 0: of dcol itself)      Scaler.py:61-66,71-72
 *   column_sums    : the raw sums behind column_moments over THIS rank's rows, for tables sharded by rows
 *                    (central == 0: s1 = sum g(z); central != 0: s1 = sum (g(z) - centre)^2, s2 = sum (g(z) - centre)^3;
 *                    g = boxcox(., lambda) or the identity for lambda <= 0); the caller all-reduces them
 *   column_distinct: min(number of distinct values, cap + 1), cap <= 15      Scaler.py:63
 *   column_distinct_values: the same, and when there are at most `cap` the sorted values themselves in
 *                    values[0 .. *ndistinct) (host array of cap doubles; the union over ranks decides)
 *   forward        : Scaler.forward_transform on the columns `cols` (host int array); algo
 *                    0 min-max, 1 normal, 2 inv-normal, 3 auto-normal, 4 auto-inv-normal; the
 *                    per-column parameter arrays are host pointers (unused ones may be NULL) */
int scfgp_scaler_colstats(int device, const double* dX, long long N, int D, long long ld, double* mn, double* mx,
                          int* is_const);
int scfgp_scaler_gather_column(int device, const double* dX, long long N, long long ld, int col, double shift,
                               double range, double* dcol);
int scfgp_scaler_column_moments(int device, const double* dcol, long long N, double lambda, double* mean,
                                double* var, double* skew_g1);
int scfgp_scaler_column_sums(int device, const double* dcol, long long N, double lambda, int central, double centre,
                             double* s1, double* s2);
int scfgp_scaler_column_distinct(int device, const double* dcol, long long N, int cap, int* ndistinct);
int scfgp_scaler_column_distinct_values(int device, const double* dcol, long long N, int cap, int* ndistinct,
                                        double* values);
int scfgp_scaler_forward(int device, int algo, const double* dX, long long N, long long ld, const int* cols, int ncols,
                         const double* mn, const double* mx, const double* lambda, const double* mu,
                         const double* sd, double* dOut, long long ldo);

/* ---- self-test hooks used by tests/ (each runs ONE kernel family on caller data so a
 * failing parity test names the kernel, not the pipeline).  Host pointers, row-major. */
/* C (m x n) = op(A).op(B)^T-style products through the DMMA mainloop.
 *   layout 0: NT  C[i,j] = sum_k A[i,k] B[j,k]   (A m x k, B n x k)
 *   layout 1: TN  C[i,j] = sum_r A[r,i] B[r,j]   (A k x m, B k x n)
 *   layout 2: NN  C[i,j] = sum_k A[i,k] B[k,j]   (A m x k, B k x n) */
int scfgp_dbg_gemm(int device, int layout, int m, int n, int k, const double* A,
                   const double* B, double* C);
/* C (m x n) = A (m x k) . B^T (n x k) through digit extraction (per-row power-of-two
 * scales) + the int8 tensor-core mainloop with `digits` digits. */
int scfgp_dbg_emu_gemm(int device, int digits, int m, int n, int k, const double* A, const double* B,
                       double* C);
/* lower Cholesky + inverse factor of a symmetric K x K matrix (info = pivot+1 or 0) */
/* Developer tool: per-CTA clock64() intervals of the tcgen05 kernels' roles (8 counters per CTA, see
 * umma_i8.cuh).  enable != 0 starts recording into a fresh buffer for `max_ctas` CTAs; a later call
 * (enable 0 or 1) first copies the counters recorded so far to `out` (max_ctas * 8 values; may be NULL). */
int scfgp_dbg_tile_timing(int device, int enable, unsigned long long* out, int max_ctas);
/* Measured ceiling of the int8 tensor pipe for the digit-product instruction mix of `digits` digits on
 * 128 x `tile_n` tiles (5/96, 5/64, 6/64, 4/64): the product kernels' tcgen05.mma stream with no operand
 * traffic and no epilogue, `steps` K = 32 steps on every SM at once (random_data != 0: full-range int8
 * operands, the power-relevant case).  Returns chip-wide int8 TOP/s, clocks per step and the SM clock
 * the run averaged.  bench.py runs it in-process for the second roofline denominator. */
int scfgp_dbg_mma_stream_peak(int device, int digits, int tile_n, int steps, int random_data, double* tops,
                              double* clk_per_step, double* sm_mhz);
int scfgp_dbg_factor(int device, int K, const double* A, double* L, double* Li, long long* info);
/* Phi (N x K) for given params through the feature kernel */
int scfgp_dbg_features(scfgp_handle* h, const double* params, const double* X, long long N,
                       double* Phi);

#ifdef __cplusplus
}
#endif
#endif /* SCFGP_B200_H */
