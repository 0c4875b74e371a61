/*
 * ggp_b200.h -- C ABI of the B200-native collapsed sparse-GP hot path.
 *
 * The reference (vr308/Generalised-Gaussian-Processes) has no FFI: its boundary is the Python call its
 * model wrappers make into gpytorch / pymc3.  Each entry point below names the reference call it replaces
 * (paths relative to /root/reference).  A maintainer binds these with ctypes.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers to contiguous float64 / int32 unless marked [host]
 *   - `stream` is a cudaStream_t passed as void*; every call only ENQUEUES work on it (no host sync) unless noted
 *   - return value: 0 ok; <0 argument -k invalid / handle not reserved; >0 CUDA runtime error code
 *   - numerical failure is never an abort: per-batch `info[b]` follows LAPACK potrf (k>0: leading minor k
 *     not positive definite) so the host runs the psd_safe_cholesky jitter ladder
 *   - theta rows are CONSTRAINED values [ell_0..ell_{d-1}, sf2 (outputscale), s2 (noise variance)]
 *   - gradient rows are w.r.t. the constrained values, layout [d_ell[d], d_sf2, d_s2, d_Z[m*d]]
 *   - a handle is bound to one device; one thread / one stream at a time
 */
#ifndef GGP_B200_H
#define GGP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ggp_handle ggp_handle_t;

enum { GGP_KERNEL_RBF = 0, GGP_KERNEL_MATERN32 = 1, GGP_KERNEL_MATERN52 = 2,
       GGP_KERNEL_RQ = 3 /* rational quadratic (1 + d2 / (2 alpha))^(-alpha), alpha = cfg.kernel_param (a constant of the evaluation) */,
       GGP_KERNEL_PERIODIC = 4 /* factor of a composite kernel only: exp(-2 sum_c sin^2(pi (x_c - z_c) / p_c) / ell_c^2) */,
       GGP_KERNEL_COMPOSITE = 5 /* cfg.kernel: the program registered with ggp_set_kernel_program */ };
/* Composite covariance  k(x, z) = sum_t a_t prod_f phi_tf(x, z)  -- the structure of the reference's CO2 model
 * (experiments/co2_bayesian_sgpr_hmc.py:74-83 gpytorch, :107-149 pymc3).  Factor kinds: GGP_KERNEL_RBF / MATERN32 / MATERN52 / RQ /
 * PERIODIC, all with ARD lengthscales.  One parameter row per theta draw, in program order:
 *     for each term t:  a_t (variance),  then for each factor:  ell[d],  then RQ: alpha | PERIODIC: period[d].
 * With GGP_KERNEL_COMPOSITE the `theta` rows of the SGPR entry points keep their [d+2] shape but only carry
 * theta[d] = k(x, x) = sum_t a_t and theta[d+1] = s2 (the ell slots are ignored); the parameter rows and the gradient rows w.r.t. them
 * are the device arrays registered with ggp_set_kernel_params.  The gradient rows of the entry points then hold [0 (d+1 slots), d_s2,
 * d_Z].  Evaluated on the FP64 DMMA plan (GGP_PREC_FP64_I8 falls back to it); the SVGP / SGPMC entry points reject it (-3). */
#define GGP_KPROG_MAX_TERMS 6
#define GGP_KPROG_MAX_FACTORS 3
#define GGP_KPROG_MAX_PARAMS 64
typedef struct {
  int32_t nterms;
  int32_t nfactors[GGP_KPROG_MAX_TERMS];
  int32_t kind[GGP_KPROG_MAX_TERMS][GGP_KPROG_MAX_FACTORS];
} ggp_kprog;
/* GGP_PREC_FP64: FP64 tensor-core DMMA.  GGP_PREC_FP64_I8: the same contractions evaluated to FP64-class accuracy by exact integer
 * slicing (7 balanced radix-256 digits per operand, tcgen05.mma kind::i8, int32 TMEM accumulators; csrc/gemm_i8.cuh); used for the
 * streamed passes when batch == 1, the padded inducing count is in [128, 4096] and d <= 16, the DMMA path otherwise (the host
 * driver evaluates a batch of theta rows on a large problem row by row so that each runs on this path).
 * GGP_PREC_TF32X3 is rejected (-3): split-precision cannot meet the gradient tolerance. */
enum { GGP_PREC_FP64 = 0, GGP_PREC_TF32X3 = 1, GGP_PREC_FP64_I8 = 2 };
enum { GGP_LIK_GAUSSIAN = 0, GGP_LIK_BERNOULLI_PROBIT = 1 };

typedef struct {
  int32_t kernel;      /* GGP_KERNEL_*   (ScaleKernel(RBFKernel(ard)) at models/sgpr.py:36 is GGP_KERNEL_RBF) */
  int32_t precision;   /* GGP_PREC_*     contraction arithmetic; everything else is float64 */
  int32_t chunk_rows;  /* rows of X per streamed chunk; 0 = library default */
  int32_t tile_cache_mib; /* MiB of device memory the handle may use to keep the k(X_local,Z) tiles of pass1 for the pass2 of the
                             same evaluation (saves the second tile build); 0 = never materialise more than chunk_rows x m */
  double kernel_param;    /* GGP_KERNEL_RQ: alpha > 0 (RQKernel / RatQuad, experiments/co2_bayesian_sgpr_hmc.py:77,127); else ignored */
} ggp_cfg;

int ggp_version(void);
const char* ggp_last_error(void);

/* lifetime ------------------------------------------------------------------------------------------------ */
int ggp_create(ggp_handle_t** out, int device);
int ggp_destroy(ggp_handle_t* h);
/* bytes of device workspace ggp_reserve would own for these shapes */
int ggp_workspace_bytes(const ggp_cfg* cfg, int64_t n_local, int m, int d, int batch, size_t* out);
/* (re)allocate the handle's workspace; synchronous (cudaMalloc); call once per shape */
int ggp_reserve(ggp_handle_t* h, const ggp_cfg* cfg, int64_t n_local, int m, int d, int batch);

/* composite kernels: number of parameters of `prog` in d input dimensions (<0: malformed / too many) */
int ggp_kprog_nparams(const ggp_kprog* prog, int d);
/* register the program for cfg.kernel == GGP_KERNEL_COMPOSITE on a RESERVED handle (allocates its row accumulators: synchronous) */
int ggp_set_kernel_program(ggp_handle_t* h, const ggp_kprog* prog, int d);
/* parameter rows kparams[batch, P] and the two gradient outputs the next finish / pass2 write: kgrad_mm[batch, P] (Kzz part + the
 * explicit k(x,x) dependence; complete when grad_mm is) and kgrad_partial[batch, P] (this rank's rows of Kzx; all-reduced with
 * grad_partial when rows are sharded).  Pointers only: nothing is enqueued. */
int ggp_set_kernel_params(ggp_handle_t* h, const double* kparams, double* kgrad_mm, double* kgrad_partial);

/* SGPR collapsed bound + gradient --------------------------------------------------------------------------
 * Replaces, together:  output = self.forward(train_x); loss = -mll(output, train_y); loss.backward()
 *   (models/sgpr.py:123-129, models/bayesian_sgpr_hmc.py:110-114,128-133)  and the pymc3 logp/dlogp of
 *   gp.marginal_likelihood (models/bayesian_sgpr_hmc.py:66-71) that NUTS calls per leapfrog step.
 * Call order per evaluation:  factor -> pass1 -> [allreduce partial over row shards] -> finish
 *                             -> pass2 -> [allreduce grad_partial]  ; total grad = grad_mm + sum(grad_partial)
 */

/* Scheduling hint: the caller is about to enqueue ggp_sgpr_prefetch_tiles[_part] (on another stream) right AFTER the next
 * ggp_sgpr_factor.  The Kzz factorisation then runs as one thread-block-cluster launch that keeps its own 8 SMs (on a high-priority
 * stream of the handle, joined back into `stream`), so that it runs next to the tile build instead of queueing behind its CTAs.
 * (A prefetch enqueued BEFORE the factorisation is detected without this call.)  Same results either way. */
int ggp_sgpr_expect_prefetch(ggp_handle_t* h, int on);

/* Kzz(theta_b) + jitter_b I = L L^T ; L^{-1} kept in the handle.  info[b] device int32.
 * (InducingPointKernel._inducing_mat / _inducing_inv_root + psd_safe_cholesky, reached from models/sgpr.py:41) */
int ggp_sgpr_factor(ggp_handle_t* h, const ggp_cfg* cfg, void* stream,
                    const double* Z /*[m,d]*/, const double* theta /*[batch,d+2]*/,
                    const double* jitter /*[batch]*/, int m, int d, int batch, int32_t* info /*[batch]*/);

/* Optional: build the k(X_local, Z) tiles of ALL local rows into the handle's tile cache (cfg.tile_cache_mib) on `stream` -- which
 * may be a different stream from the one ggp_sgpr_factor runs on: the tiles do not depend on the factorisation, so the two overlap
 * (the host joins the streams before ggp_sgpr_pass1).  The next ggp_sgpr_pass1 with the same X, Z, theta pointers, n_local and batch
 * skips its tile builds; the operands must not change in between.  A no-op (returns 0) when the handle has no tile cache.
 * (the Kxz block of InducingPointKernel.forward, models/sgpr.py:41, evaluated ahead of the Cholesky of Kzz) */
int ggp_sgpr_prefetch_tiles(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X /*[n_local,d]*/, int64_t n_local,
                            const double* Z, const double* theta, int m, int d, int batch);

/* The same for the row range [row0, row0 + nrows) only: parts are issued in ascending order starting at row0 = 0 and together cover
 * [0, n_local) (on the FP64 path each part starts on a chunk boundary); the cache becomes valid with the last part.  Lets a host
 * pipeline the H2D copy of X with the tile build piece by piece. */
int ggp_sgpr_prefetch_tiles_part(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X /*[n_local,d] base*/,
                                 int64_t n_local, int64_t row0, int64_t nrows, const double* Z, const double* theta, int m, int d,
                                 int batch);

/* stream the local rows: partial[b] = [ A A^T (m*m, row-major, symmetric) | A y (m) | y^T y, sum_n k_nn, n_local ]
 * with A = L^{-1} k(Z, X_local).  Never materialises more than chunk_rows x m of k(X,Z). */
int ggp_sgpr_pass1(ggp_handle_t* h, const ggp_cfg* cfg, void* stream,
                   const double* X /*[n_local,d]*/, const double* y /*[n_local]*/, int64_t n_local,
                   const double* Z, const double* theta, int m, int d, int batch,
                   double* partial /*[batch, m*m+m+3]*/);

/* m x m section on the (all-reduced) partial: bound[b] = F (NOT divided by N), grad_mm[b] = the Kzz-, noise- and
 * k_nn-dependent part of dF/d(ell,sf2,s2,Z); P,u kept in the handle for pass2.  info[b]: chol(I + A A^T / s) status.
 * (ExactMarginalLogLikelihood.forward, models/sgpr.py:125; MarginalSparse._build_marginal_likelihood_logp) */
int ggp_sgpr_finish(ggp_handle_t* h, const ggp_cfg* cfg, void* stream,
                    const double* Z, const double* theta, int m, int d, int batch,
                    const double* partial /*[batch, m*m+m+3]*/, int need_grad,
                    double* bound /*[batch]*/, double* grad_mm /*[batch, d+2+m*d] or NULL*/, int32_t* info /*[batch]*/);
/* The state pass 2 needs is ready in stream order when ggp_sgpr_finish returns, and so is `bound` when need_grad = 0.  With
 * need_grad = 1, `bound` and grad_mm are written by a chain that runs on the handle's auxiliary stream next to pass 2 (only beta, u,
 * P_A and Q are on the path to pass 2) and are complete, in `stream` order, once ggp_sgpr_pass2 of the same evaluation has been
 * enqueued -- or after ggp_sgpr_join for a caller that skips pass 2. */
int ggp_sgpr_join(ggp_handle_t* h, void* stream);

/* second streaming pass: grad_partial[b] = sum over local rows of (P Kzx + u y^T) o dKzx/d(ell,sf2,Z)
 * (what loss.backward() propagates into the N x M kernel block; models/sgpr.py:129) */
int ggp_sgpr_pass2(ggp_handle_t* h, const ggp_cfg* cfg, void* stream,
                   const double* X, const double* y, int64_t n_local,
                   const double* Z, const double* theta, int m, int d, int batch,
                   double* grad_partial /*[batch, d+2+m*d]*/);

/* Pass 1 of the eval-mode predictive (replaces ggp_sgpr_pass1 in factor -> pass1 -> [allreduce] -> finish(need_grad = 0) -> predict):
 * ExactGP.__call__ in eval mode evaluates the InducingPointKernel on the TRAINING inputs with the sgpr diagonal correction on
 * (models/sgpr.py:150-160 -> likelihood(self(test_x))), so training row n carries the noise Lambda_n = s2 + max(k_nn - q_nn, 0).
 * partial[b] = [ s2 A W A^T | s2 A W y | 0, n sf2, n ] with W = diag(1 / Lambda): ggp_sgpr_finish then leaves B = I + A W A^T and
 * c = L_B^{-1} A W y for ggp_sgpr_predict (its `bound` output is meaningless for this state).  FP64 DMMA path. */
int ggp_sgpr_predict_pass1(ggp_handle_t* h, const ggp_cfg* cfg, void* stream,
                           const double* X /*[n_local,d]*/, const double* y /*[n_local]*/, int64_t n_local,
                           const double* Z, const double* theta, int m, int d, int batch,
                           double* partial /*[batch, m*m+m+3]*/);

/* Sparse predictive at the state left by factor + (predict_)pass1 + finish  (likelihood(self(test_x)) in eval mode,
 * models/sgpr.py:150-160; models/bayesian_sgpr_hmc.py:198-231).  mean,var: [batch, ns]; cov: [batch, ns, ns] or NULL (any ns: tiled).
 * var/cov include the eval-mode diagonal correction clamp(k** - ||a*||^2, 0) on the test rows and, if add_noise, + s2. */
int ggp_sgpr_predict(ggp_handle_t* h, const ggp_cfg* cfg, void* stream,
                     const double* Xs /*[ns,d]*/, int64_t ns,
                     const double* Z, const double* theta, int m, int d, int batch, int add_noise,
                     double* mean, double* var, double* cov);

/* Whitened sparse-GP marginals + expected log-likelihood, value and gradient.  One routine serves
 *   - the SVGP minibatch ELBO (models/svgp.py:104-110 ; models/bayesian_svgp.py:160-167 with batch = number of theta draws):
 *       data_jitter = 1e-4, lik_scale = 1/nb, kl_scale = 1/num_data
 *   - the SGPMC whitened-conditional log-likelihood tfp-HMC differentiates per leapfrog (models/sgp_hmc.py:63-69):
 *       qLs = NULL (S = 0), qm = v, data_jitter = 0, lik_scale = 1, kl_scale = 0, nb = N (streamed in chunks);
 *       qm_batched = 1: every batch element (HMC chain) carries its own whitened vector v[b] next to its own theta[b], so all the
 *       chains of a rank are evaluated in ONE launch sequence (BASELINE configs[4]: 64 chains sharded over 8 GPUs)
 * out[b] = lik_scale * sum_i E_q[log p(y_i|f_i)] - kl_scale * KL(N(m, Ls Ls^T) || N(0, I)),
 *   mu_i = a_i^T m, var_i = k_ii + data_jitter + ||Ls^T a_i||^2 - ||a_i||^2, a = L^{-1} k(Z, x_i), L L^T = Kzz + jitter_b I.
 * grad[b] layout: [d_ell[d], d_sf2, d_s2, d_Z[m*d], d_m[m], d_Ls[m*m] (lower triangle, row-major, upper = 0)].
 * jitter[b] is the TOTAL diagonal jitter added to Kzz (variational_cholesky_jitter 1e-6 + ladder, or gpflow's 1e-5).
 * likelihood: GGP_LIK_GAUSSIAN (noise s2 from theta) or GGP_LIK_BERNOULLI_PROBIT (20-point Gauss-Hermite). */
int ggp_svgp_elbo(ggp_handle_t* h, const ggp_cfg* cfg, void* stream,
                  const double* xb /*[nb,d]*/, const double* yb /*[nb]*/, int64_t nb,
                  const double* Z, const double* qm /*[m], or [batch,m] when qm_batched*/, int qm_batched,
                  const double* qLs /*[m,m] or NULL*/,
                  const double* theta /*[batch,d+2]*/, const double* jitter /*[batch]*/,
                  int m, int d, int batch, double data_jitter, double lik_scale, double kl_scale, int likelihood,
                  int need_grad, double* elbo /*[batch]*/, double* grad /*[batch, d+2+m*d+m+m*m] or NULL*/, int32_t* info);

/* SVGP predictive marginals at test inputs (posterior_predictive, models/svgp.py:132-141; diagonal only -- the exact
 * expression, see SURVEY A.7 on fast_pred_var):  mean[b][n] = a^T m ; var[b][n] = k** + data_jitter + ||Ls^T a||^2 - ||a||^2 (+ s2) */
int ggp_svgp_predict(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* xs /*[ns,d]*/, int64_t ns,
                     const double* Z, const double* qm /*[m] or [batch,m]*/, int qm_batched, const double* qLs /*or NULL*/,
                     const double* theta, const double* jitter /*[batch]*/, int m, int d, int batch, double data_jitter, int add_noise,
                     double* mean /*[batch,ns]*/, double* var /*[batch,ns]*/, int32_t* info);

/* building blocks, exported for the parity tests and the roofline probes ------------------------------------- */

/* in-place batched Cholesky (lower) of a[batch, m, m] row-major + optional explicit inverse of the factor.
 * (the psd_safe_cholesky call sites; SURVEY 8b ggp_chol_batched) */
int ggp_chol_batched(ggp_handle_t* h, void* stream, double* a /*[batch,m,m]*/, double* linv /*[batch,m,m] or NULL*/,
                     int m, int batch, int32_t* info);
/* C[mm,nn] = alpha * A[mm,kk] * B[nn,kk]^T + beta * C   (row-major, float64, DMMA). ld* in elements. */
int ggp_gemm_nt(ggp_handle_t* h, void* stream, const double* A, int64_t lda, const double* B, int64_t ldb,
                double* C, int64_t ldc, int mm, int nn, int kk, double alpha, double beta);
/* same with the structure switches the streamed passes use: kmode bit0/1 = A lower/upper triangular, bit2/3 = B lower/upper
 * triangular (k-range clipped per tile); sym 1/2 = only upper/lower 128x128 output tiles; splits > 1 = split-K, split s
 * accumulates into C + s*split_stride (beta must be 1, buffers pre-zeroed). */
int ggp_gemm_nt_ex(ggp_handle_t* h, void* stream, const double* A, int64_t lda, const double* B, int64_t ldb,
                   double* C, int64_t ldc, int mm, int nn, int kk, double alpha, double beta, int kmode, int sym,
                   int splits, int64_t split_stride);
/* C[mm,nn] = A[mm,kk] * B[nn,kk]^T evaluated by the sliced-integer tcgen05 path (row-scaled 7 x 8-bit digits, exact int32
 * accumulation, kk <= 16384).  Allocates and frees its digit planes (synchronous): a test / probe entry, not a hot-path one. */
int ggp_gemm_nt_i8(ggp_handle_t* h, void* stream, const double* A, int64_t lda, const double* B, int64_t ldb,
                   double* C, int64_t ldc, int mm, int nn, int kk);
/* k(X1, X2)[n1, n2] dense tile (tests) */
int ggp_kernel_matrix(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X1, int64_t n1,
                      const double* X2, int64_t n2, const double* theta /*[d+2]*/, int d, double* out /*[n1,n2]*/);
/* NUTS tree bookkeeping on the device -----------------------------------------------------------------------
 * Replaces the per-leapfrog host work of  pm.sample(n, tune, chains=1, step=pm.NUTS())  (models/bayesian_sgpr_hmc.py:73-78,
 * models/all_in_HMC.py:60): multinomial NUTS with pymc3's defaults, C chains in lock-step, one warp per chain (csrc/nuts.cuh).
 * The host (hmc.py nuts_sample) owns every buffer, draws the random numbers, runs the doubling loop and the step-size / mass
 * adaptation; these calls do everything between two logp/dlogp evaluations in one launch:
 *   ggp_nuts_begin          p0 = z / sqrt(inv_mass), initial energy, tree = {x}
 *   ggp_nuts_subtree_begin  direction (u row 0), empty subtree, x_eval = first position to evaluate
 *   ggp_nuts_leaf           consumes (lp_eval, g_eval) at x_eval: leaf weight (u row 2 + leaf), divergence, progressive sample,
 *                           U-turn checks of the balanced sub-subtrees ending at this leaf, then the next x_eval; the leaf index is a
 *                           device counter, so ONE captured launch serves every leaf (replayed in the same graph as the evaluation)
 *   ggp_nuts_subtree_end    merge (u row 1), tree U-turn test, *any_active = some chain keeps doubling
 *   ggp_nuts_end            state <- proposal, acc_prob, trace row k (k < 0: tuning step, nothing recorded)
 * All pointers are device pointers; [C,P] arrays are row-major; int arrays are int32 flags / counters. */
typedef struct ggp_nuts_state {
  int32_t C, P, K, pad_;             /* chains, parameters per chain, checkpoint slots (>= max tree depth) */
  double max_energy_error;           /* divergence threshold on |energy change| (pymc3: 1000) */
  /* sampler state, persists across transitions */
  double *x, *lp, *g;                /* [C,P], [C], [C,P] current point, its log density and gradient */
  double *eps, *inv_mass;            /* [C], [C,P] step size, diagonal inverse metric */
  /* exchange with the logp/dlogp evaluation */
  double *x_eval;                    /* [C,P] point to evaluate */
  const double *lp_eval, *g_eval;    /* [C], [C,P] its log density and gradient */
  /* the tree of the current transition */
  double *e0, *xl, *pl, *gl, *xr, *pr, *gr, *x_prop, *lp_prop, *g_prop, *log_w, *p_sum, *sum_acc, *n_leaf;
  int32_t *depth, *diverged, *active;
  /* the subtree of the current doubling */
  double *e, *xe, *pe, *ge, *p_half, *s_log_w, *s_p_sum, *s_x, *s_lp, *s_g;
  double *p_ck, *ps_ck;              /* [K,C,P] checkpoints of the iterative U-turn scheme */
  int32_t *right, *s_turn, *s_div, *building, *leaf;   /* [C] */
  int32_t *any_active;               /* [1] */
  const double *u;                   /* [2 + 2^depth, C] uniforms of the current doubling: row 0 direction, row 1 merge, rows 2.. leaves */
  /* outputs */
  double *acc_prob;                  /* [C] mean over the leaves of min(1, exp(-dE)) */
  double *samples, *lps;             /* [n,C,P], [n,C] */
  int32_t *depths, *nleaps, *divs;   /* [n,C] */
} ggp_nuts_state;
int ggp_nuts_state_size(void);   /* sizeof(ggp_nuts_state), for binding checks */
int ggp_nuts_begin(void* stream, const ggp_nuts_state* s, const double* z /*[C,P] standard normal draws*/);
int ggp_nuts_subtree_begin(void* stream, const ggp_nuts_state* s);
int ggp_nuts_leaf(void* stream, const ggp_nuts_state* s);
int ggp_nuts_subtree_end(void* stream, const ggp_nuts_state* s);
int ggp_nuts_end(void* stream, const ggp_nuts_state* s, int k);

/* pymc3 log-posterior around the collapsed bound (models/bayesian_sgpr_hmc.py:60-71), batched over chains: x[C, d+2] is the
 * unconstrained point (ls_log__[d], sig_f_log__, sig_n_log__).  ggp_vfe_theta writes the constrained theta rows the SGPR entry points
 * take (ell = e^x, sf2 = sig_f^2, s2 = sig_n^2); ggp_vfe_logp turns the bound and its gradient rows (row stride ldg, first d+2
 * entries used) into logp / dlogp with the Gamma(2,1) / HalfCauchy(1) priors and the log-Jacobians (with_prior = 0: the bound and the
 * chain rule only).  info / info_b (nullable, int32[C]): a non-zero entry or a non-finite value gives logp = -inf, dlogp = 0. */
int ggp_vfe_theta(void* stream, const double* x, int C, int d, double* theta /*[C,d+2]*/);
int ggp_vfe_logp(void* stream, const double* x, const double* bound, const double* grad, int64_t ldg, const int* info,
                 const int* info_b, int C, int d, int with_prior, double* lp /*[C]*/, double* dx /*[C,d+2]*/);

/* register-resident mma.sync m8n8k4 f64 loop on every SM: measured FP64 tensor-pipe peak [host out, TFLOP/s]; synchronous */
int ggp_probe_dmma_peak(ggp_handle_t* h, void* stream, int iters, double* tflops_out /*[host]*/);
/* tcgen05.mma kind::i8 issue loop with shared-memory-resident operands on every SM (no loads, no epilogue): the measured tensor-pipe
 * roofline of the sliced-integer GEMMs [host out, int8 TOP/s: best, uniform 128x256x32 MMAs, the MMA mix of the 128x64 tiles, the MMA mix
 * of 128x32 tiles]; synchronous */
int ggp_probe_i8_peak(ggp_handle_t* h, void* stream, int iters, double* tops_out /*[4] host*/);

/* instrumentation: CUDA-event spans around the kernel categories, on the caller's stream (no host sync until read).
 * categories: 0 tile build, 1 triangular multiply, 2 SYRK, 3 backward GEMM+moments, 4 m x m section, 5 other.
 * ggp_profile_read synchronises the device, returns the accumulated milliseconds / span counts per category and the number
 * of kernels launched since the last read, and resets all three. */
int ggp_profile_enable(ggp_handle_t* h, int on);
int ggp_profile_read(ggp_handle_t* h, double* ms_out /*[6] host*/, int64_t* spans_out /*[6] host*/, int64_t* launches_out /*host*/);

#ifdef __cplusplus
}
#endif
#endif /* GGP_B200_H */
