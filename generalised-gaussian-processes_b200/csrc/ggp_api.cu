// C ABI of the B200-native collapsed sparse-GP hot path (see include/ggp_b200.h for the contract and the
// reference call sites each entry point replaces).  Host orchestration only enqueues kernels on the caller's stream.
#include "../../include/ggp_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "dense_mm.cuh"
#include "gemm_dmma.cuh"
#include "gemm_tma.cuh"
#include "gemm_mm64.cuh"
#include "gemm_i8.cuh"
#include "kernel_tiles.cuh"
#include "misc.cuh"
#include "svgp.cuh"
#include "nuts.cuh"
#include "kprog.cuh"

using namespace ggp;

static thread_local char g_err[512] = "";
static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      snprintf(g_err, sizeof(g_err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return (int)e_;                                                                         \
    }                                                                                         \
  } while (0)
#define CKL()            \
  do {                   \
    h->launches++;       \
    CK(cudaGetLastError()); \
  } while (0)

struct ggp_handle {
  int device = 0, sm_count = 148;
  // reserved shapes
  int64_t n_local = 0;
  int m = 0, d = 0, batch = 0, Mp = 0, nc = 0, splits = 1;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  // m x m (per batch stride Mp*Mp)
  double *L = 0, *Linv = 0, *LinvT = 0, *Wk = 0, *Bm = 0, *LBinv = 0, *LBinvT = 0, *Binv = 0, *PA = 0, *Gbar = 0, *T1 = 0,
         *P = 0, *Gzz = 0, *Tblk = 0;
  // vectors (per batch stride Mp)
  double *bvec = 0, *cvec = 0, *beta = 0, *u = 0, *yty = 0, *ds2 = 0, *rowacc = 0, *rk = 0;
  // streamed chunk buffers
  double *Kc = 0, *At = 0, *Spart = 0, *mom_part = 0, *mom_acc = 0;
  // optional HBM cache of the k(X_local, Z) tiles built in pass 1, reused by pass 2 of the same evaluation (cfg.tile_cache_mib)
  double* kc_all = nullptr;
  size_t kc_all_bytes = 0;
  int64_t kc_rows = 0;          // rows per batch element in kc_all
  bool kc_valid = false;
  const void *kc_X = nullptr, *kc_Z = nullptr, *kc_theta = nullptr;
  int64_t kc_n = 0;
  int kc_batch = 0;
  KSpec kc_kind{0, 0.0};
  bool pf_valid = false;        // ggp_sgpr_prefetch_tiles filled the cache for the key below; the next pass 1 with that key skips its builds
  const void *pf_X = nullptr, *pf_Z = nullptr, *pf_theta = nullptr;
  int64_t pf_n = 0;
  int pf_batch = 0;
  KSpec pf_kind{0, 0.0};
  int64_t pf_next_row = 0;      // ggp_sgpr_prefetch_tiles_part: rows [0, pf_next_row) are built
  bool pf_expected = false;     // ggp_sgpr_expect_prefetch: the caller will enqueue a tile prefetch right after the next ggp_sgpr_factor
  double *sv[5] = {0, 0, 0, 0, 0}, *rowout = 0;
  // int8 digit planes of the sliced-integer path (GGP_PREC_FP64_I8; gemm_i8.cuh)
  char* arena_i8 = nullptr;
  size_t arena_i8_bytes = 0;
  int8_t *Lq = 0, *Pq = 0, *Atq = 0, *Kq = 0;
  int *eL = 0, *eP = 0, *eKA = 0;   // row exponents of L^{-1} / Q digits; eKA[0..1] = scalar exponents of the k(X,Z) / A digits (k_i8_exponents)
  bool atq_valid = false;           // atq_all holds the A^T digits of the pass 1 identified by the kc_* key (one-launch triangular multiply)
  int8_t* kq_all = nullptr;
  size_t kq_all_bytes = 0;
  int8_t* atq_all = nullptr;   // digit planes of A^T for ALL local rows: the triangular multiply of pass 1 becomes one launch
  size_t atq_all_bytes = 0;
  long long* i8_dbg = nullptr;   // developer timeline buffer (GGP_I8_TIMELINE=2)
  int i8_dbg_prints = 0;
  int nsv = 0;
  // instrumentation
  long long launches = 0;
  bool profiling = false;
  struct Span { int cat; cudaEvent_t e0, e1; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
  int cur_cat = -1;
  cudaEvent_t cur_e0 = nullptr;
  // CUDA graphs of the blocked Cholesky + inverse
  struct CholGraph { double *A, *Linv, *LinvT; int batch; long long nodes; cudaGraphExec_t exec; bool factored; };
  std::vector<CholGraph> chol_graphs;
  bool use_graphs = true;
  bool chol_fused = true;       // fused panel + trailing-update kernel in the blocked Cholesky (GGP_CHOL_FUSED=0: two library GEMMs)
  bool chol_lookahead = true;   // the trailing-update CTA that owns the next diagonal block factors it in the same launch (GGP_CHOL_LOOKAHEAD=0: own launch)
  bool chol_cluster = true;     // Kzz factorisation next to a tile build: cluster-resident plan on a high-priority stream (GGP_CHOL_CLUSTER=0: the launch chain)
  cudaStream_t hp_stream = nullptr;
  cudaEvent_t ev_hp0 = nullptr, ev_hp1 = nullptr;
  int* chol_ctr = nullptr;
  bool chol_small = true;       // Mp <= 128: the whole factorisation + inverse in one launch (GGP_CHOL_SMALL=0: the multi-launch plan)
  int mm64_max_tiles = 96;      // EPI_STORE products with at most this many 128 x 128 work items run on k_mm64 (GGP_MM64_MAX_TILES; 0 = never)
  cudaStream_t cap_stream = nullptr;
  cudaStream_t aux_stream = nullptr;   // second stream for the independent product chain of the finish section (fork / join by events)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool aux_pending = false;            // work of the last finish() is still running on aux_stream: join before its buffers are touched
  int32_t* info_ws = nullptr;
  long long* ct_dbg = nullptr;         // developer timeline of the Cholesky chain (GGP_CHOL_TIMELINE)
  double* ysc = nullptr;
  double* rk_part = nullptr;
  double* piv_tol = nullptr;   // [batch] pivot threshold of the next Cholesky (k_build_kzz sets it; 0 = LAPACK semantics)
  // composite kernel (GGP_KERNEL_COMPOSITE; kprog.cuh)
  KProgDev prog{};
  bool prog_set = false;
  const double* kth = nullptr;         // [batch, P] parameter rows
  double *kgrad_mm = nullptr, *kgrad_partial = nullptr;
  double* krow = nullptr;              // [batch, m, P + d] row accumulators of k_kprog_grad
  size_t krow_bytes = 0;
};

enum { CAT_BUILD = 0, CAT_TRMM = 1, CAT_SYRK = 2, CAT_BWD = 3, CAT_MM = 4, CAT_OTHER = 5, CAT_COUNT = 6 };

static cudaEvent_t get_event(ggp_handle* h) {
  if (!h->pool.empty()) {
    cudaEvent_t e = h->pool.back();
    h->pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
static void prof_begin(ggp_handle* h, cudaStream_t st, int cat) {
  if (!h->profiling) return;
  h->cur_cat = cat;
  h->cur_e0 = get_event(h);
  cudaEventRecord(h->cur_e0, st);
}
static void prof_end(ggp_handle* h, cudaStream_t st) {
  if (!h->profiling || h->cur_cat < 0) return;
  cudaEvent_t e1 = get_event(h);
  cudaEventRecord(e1, st);
  h->spans.push_back({h->cur_cat, h->cur_e0, e1});
  h->cur_cat = -1;
}
struct ProfScope {
  ggp_handle* h; cudaStream_t st;
  ProfScope(ggp_handle* h_, cudaStream_t st_, int cat) : h(h_), st(st_) { prof_begin(h, st, cat); }
  ~ProfScope() { prof_end(h, st); }
};

static int pad_pow2_blocks(int m) {
  int nb = (m + NB - 1) / NB, p = 1;
  while (p < nb) p <<= 1;
  return p * NB;
}

struct Plan {
  int Mp, nc, splits, nsv;
  size_t bytes;
  size_t off[40];
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Plan make_plan(const ggp_cfg* cfg, int64_t n_local, int m, int d, int batch, int sm_count) {
  Plan p;
  p.Mp = pad_pow2_blocks(m);
  int64_t nmax = std::max<int64_t>(128, (n_local + 127) / 128 * 128);
  int64_t nc = (cfg && cfg->chunk_rows > 0) ? (cfg->chunk_rows + 127) / 128 * 128 : 16384;
  // keep the two chunk buffers under ~4 GiB in total
  const double budget = 4.0 * 1024 * 1024 * 1024;
  while (nc > 128 && 2.0 * batch * (double)nc * p.Mp * 8.0 > budget) nc /= 2;
  nc = std::max<int64_t>(128, nc / 128 * 128);
  if (cfg && cfg->precision == GGP_PREC_FP64_I8) nc = std::min<int64_t>(nc, I8_MAX_K);   // the SYRK's k extent is one chunk
  p.nc = (int)std::min<int64_t>(nc, nmax);
  const int tiles = sym_upper_tiles((m + BM - 1) / BM, (m + BN - 1) / BN);
  const int ctas = sm_count * CTAS_PER_SM;
  int splits = (int)((ctas + tiles * batch / 2) / std::max(1, tiles * batch));
  splits = std::max(1, std::min(16, splits));
  splits = std::min(splits, std::max(1, p.nc / 256));
  p.splits = splits;
  const size_t MM = (size_t)p.Mp * p.Mp * 8, V = (size_t)p.Mp * 8;
  const int nq = 2 * d + 1;
  size_t o = 0;
  int k = 0;
  auto take = [&](size_t bytes) {
    p.off[k++] = o;
    o += align_up(bytes, 256);
  };
  for (int i = 0; i < 14; ++i) take(MM * batch);       // L, Linv, LinvT, Wk, Bm, LBinv, LBinvT, Binv, PA, Gbar, T1, P, Gzz, Tblk
  for (int i = 0; i < 4; ++i) take(V * batch);          // bvec, cvec, beta, u
  take(256);                                            // yty
  take((size_t)batch * 8);                              // ds2
  take((size_t)batch * m * (d + 1) * 8);                // rowacc
  take((size_t)batch * p.nc * p.Mp * 8);                // Kc
  take((size_t)batch * p.nc * p.Mp * 8);                // At
  take((size_t)batch * p.splits * MM);                  // Spart
  // mom_part: one slab per 32 rows of the chunk (tile x warp column) on the per-chunk plans, 2 slabs per CTA on the one-launch plans
  // of the sliced-integer path (register-resident moments / row dots), whichever is larger
  take((size_t)batch * std::max(p.nc / 32, 2 * sm_count) * m * nq * 8);
  take((size_t)batch * m * nq * 8);                     // mom_acc
  take((size_t)batch * 8);                              // rk
  take((size_t)batch * 4 + 256);                        // info_ws
  // SVGP / SGPMC row chunk: 4096 rows, more when the batch is small (about 1 GiB per buffer: 8 chains of M = 512 stream 16384 rows per
  // launch sequence instead of 4096 -- BASELINE configs[4] on 8 GPUs)
  {
    const int64_t by_budget = (int64_t)((1024.0 * 1024 * 1024) / ((double)batch * p.Mp * 8.0)) / 128 * 128;
    p.nsv = (int)std::min<int64_t>(p.nc, std::max<int64_t>(4096, by_budget));
  }
  for (int i = 0; i < 5; ++i) take((size_t)batch * p.nsv * p.Mp * 8);   // SVGP / SGPMC row and transposed buffers
  take((size_t)batch * 4 * p.nsv * 8);                  // rowout
  take((size_t)batch * 8 + 256);                        // piv_tol
  take((size_t)batch * p.nc * 8);                       // ysc (noise-weighted y chunk of the predictive's pass 1)
  take((size_t)batch * RK_BLOCKS * 8);                  // rk_part
  p.bytes = o;
  return p;
}

static bool reserved_for(const ggp_handle* h, int64_t n_local, int m, int d, int batch) {
  return h->arena && h->m == m && h->d == d && batch <= h->batch && n_local <= h->n_local;
}

// ------------------------------------------------------------------------------------------------------------
// ---- TMA tensor maps (driver entry point resolved through the runtime: no link-time dependency on libcuda) ----
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn get_tmap_encode() {
  static tmap_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (tmap_encode_fn)ptr;
  }
  return fn;
}
static int g_use_tma = -1;   // GGP_GEMM_TMA=0 forces the LDGSTS mainloop (developer A/B)

// operand [count_outer][count_inner][rows x K] (row-major, leading dimension ld, element strides s_inner / s_outer) as a 4-D map
// (k, row, inner, outer) with a 16 x 128 box and the 128-byte swizzle.  Returns false when the layout cannot be described.
static bool make_operand_map(CUtensorMap* tm, const double* ptr, int rows, int K, int64_t ld, int cnt_inner, int64_t s_inner,
                             int cnt_outer, int64_t s_outer, int* mul_inner, int* mul_outer) {
  tmap_encode_fn enc = get_tmap_encode();
  if (!enc || !ptr || rows < 1 || K < 1) return false;
  // a batch stride of 0 is a broadcast operand: describe one instance and pin that coordinate to 0
  *mul_inner = (cnt_inner > 1 && s_inner != 0) ? 1 : 0;
  *mul_outer = (cnt_outer > 1 && s_outer != 0) ? 1 : 0;
  if (!*mul_inner) cnt_inner = 1;
  if (!*mul_outer) cnt_outer = 1;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 1) || ld < K) return false;
  if (cnt_inner > 1 && (s_inner <= 0 || (s_inner & 1))) return false;
  if (cnt_outer > 1 && (s_outer <= 0 || (s_outer & 1))) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)std::max(1, cnt_inner), (cuuint64_t)std::max(1, cnt_outer)};
  const cuuint64_t row_b = (cuuint64_t)ld * 8;
  const cuuint64_t strides[3] = {row_b, cnt_inner > 1 ? (cuuint64_t)s_inner * 8 : row_b, cnt_outer > 1 ? (cuuint64_t)s_outer * 8 : row_b};
  for (int i = 0; i < 3; ++i)
    if (strides[i] >= ((cuuint64_t)1 << 40)) return false;
  const cuuint32_t box[4] = {(cuuint32_t)TK, 128u, 1u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// small products (the m x m section) run on 64 x 64 work items, heaviest first, two CTAs per SM -- see gemm_mm64.cuh
static bool mm64_eligible(const ggp_handle* h, int epi, const GemmP& p, int nbatch) {
  auto even16 = [](const double* ptr, int64_t ld, int64_t s1, int64_t s2) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((ld | s1 | s2) & 1) == 0;
  };
  const int64_t t128 = (int64_t)((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * nbatch * p.nz2;
  // (skinny outputs -- the moment products of the SVGP path, N = 2 d + 1 -- always: a 128-wide tile is mostly padding there)
  return epi == EPI_STORE && (t128 <= h->mm64_max_tiles || (p.N <= S_T && h->mm64_max_tiles > 0)) && !p.rowdot && p.splits == 1 && (p.sym == 0 || p.sym == 2) && p.C != p.A && p.C != p.B &&
         even16(p.A, p.lda, p.sA, p.sA2) && even16(p.B, p.ldb, p.sB, p.sB2);
}

static int launch_gemm(ggp_handle* h, cudaStream_t st, int epi, const GemmP& pin, int nbatch) {
  GemmP p = pin;
  p.ntm = (p.M + BM - 1) / BM;
  p.ntn = (p.N + BN - 1) / BN;
  if (p.ntm == 0 || p.ntn == 0 || nbatch == 0) return 0;
  p.tiles_per_z = p.sym == 1 ? sym_upper_tiles(p.ntm, p.ntn) : (p.sym == 2 ? sym_lower_tiles(p.ntm, p.ntn) : p.ntm * p.ntn);
  p.total = p.tiles_per_z * nbatch * p.nz2 * p.splits;
  const bool alias = (p.C == p.A || p.C == p.B);
  if (mm64_eligible(h, epi, p, nbatch)) {
    p.ntm = (p.M + S_T - 1) / S_T;
    p.ntn = (p.N + S_T - 1) / S_T;
    p.tiles_per_z = p.ntm * p.ntn;
    p.total = p.tiles_per_z * nbatch * p.nz2;
    p.fold = (p.kmode != 0 && p.total > h->sm_count && !getenv("GGP_MM64_NO_FOLD")) ? h->sm_count : 0;
    k_mm64<<<p.total, S_THREADS, S_SMEM, st>>>(p);
    CKL();
    return 0;
  }
  if (p.Ct) return fail(-3, "launch_gemm: the transposed second store exists on the small-tile kernel only");
  const int grid = std::min(p.total, h->sm_count * CTAS_PER_SM);   // persistent CTAs, static snake order over the work items
  if (g_use_tma < 0) {
    const char* e = getenv("GGP_GEMM_TMA");
    g_use_tma = (e && e[0] == '0') ? 0 : 1;
  }
  CUtensorMap tmA, tmB;
  // in-place products (C aliases an operand) stay on the LDGSTS kernel: its loads of a tile are complete before its epilogue,
  // whereas the TMA producer prefetches the next tiles while this tile's C is being written
  if (g_use_tma && !alias && make_operand_map(&tmA, p.A, p.M, p.K, p.lda, p.nz2, p.sA2, nbatch, p.sA, &p.tmA_pz, &p.tmA_bz) &&
      make_operand_map(&tmB, p.B, p.N, p.K, p.ldb, p.nz2, p.sB2, nbatch, p.sB, &p.tmB_pz, &p.tmB_bz)) {
    if (epi == EPI_STORE)
      k_gemm_tma<EPI_STORE><<<grid, T_THREADS, T_SMEM, st>>>(tmA, tmB, p);
    else
      k_gemm_tma<EPI_MOMENTS><<<grid, T_THREADS, T_SMEM, st>>>(tmA, tmB, p);
    CKL();
    return 0;
  }
  if (epi == EPI_STORE)
    k_gemm_nt<EPI_STORE><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(p);
  else
    k_gemm_nt<EPI_MOMENTS><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(p);
  CKL();
  return 0;
}

static GemmP gemm_basic(const double* A, int64_t lda, int64_t sA, const double* B, int64_t ldb, int64_t sB, double* C,
                        int64_t ldc, int64_t sC, int M, int N, int K, double alpha, double beta, int kmode = 0) {
  GemmP p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.lda = lda; p.sA = sA;
  p.B = B; p.ldb = ldb; p.sB = sB;
  p.C = C; p.ldc = ldc; p.sC = sC;
  p.M = M; p.N = N; p.K = K;
  p.nz2 = 1; p.splits = 1;
  p.alpha = alpha; p.beta = beta; p.kmode = kmode;
  return p;
}

#define RUN(x)                \
  do {                        \
    int rc_ = (x);            \
    if (rc_ != 0) return rc_; \
  } while (0)

// In-place blocked Cholesky of A[batch][Mp][Mp] (lower), explicit inverse -> Linv, Linv^T -> LinvT.
static int chol_and_inverse_launches(ggp_handle* h, cudaStream_t st, double* A, double* Linv, double* LinvT, int batch,
                                     int32_t* info, bool factored = false) {
  const int Mp = h->Mp, nblk = Mp / NB;
  const int64_t sM = (int64_t)Mp * Mp;
  const dim3 g16(Mp / 16, Mp / 16, batch), b16(16, 16);
  // right-looking: factor the diagonal block, form the panel with the block inverse, update the trailing lower tiles
  const bool ahead = h->chol_fused && h->chol_lookahead;
  if (factored) {   // k_chol_cluster left L in the lower blocks of A (panels in place) and the block inverses in Tblk
    k_tril_merge<<<g16, b16, 0, st>>>(A, A, Mp, sM, sM, h->Tblk, sM, Linv, LinvT);
    CKL();
  } else if (ahead && h->chol_small && nblk <= 2) {   // one or two diagonal blocks: factor + inverse + transposed inverse in one launch
    k_chol_inv_small<<<batch, 256, CT_SMEM, st>>>(A, Mp, sM, h->Tblk, sM, Linv, LinvT, sM, info, h->piv_tol);
    CKL();
    return 0;
  }
  for (int k = 0; k < nblk && !factored; ++k) {
    const int k0 = k * NB;
    if (k == 0 || !ahead) {   // with look-ahead, block k > 0 was factored by the trailing update of step k - 1
      k_potf2_trti2<<<batch, 256, POTF2_SMEM, st>>>(A, Mp, sM, k, h->Tblk, sM, info, h->piv_tol);
      CKL();
    }
    if (k < nblk - 1) {
      const int rem = Mp - k0 - NB;
      if (h->chol_fused) {
        // panel + trailing update of this step in one kernel; the panel goes to the scratch matrix T1 (merged into A below)
        k_chol_trail<<<dim3(rem / NB, rem / NB, batch), 256, CT_SMEM, st>>>(A, Mp, sM, k, h->Tblk, sM, h->T1, sM,
                                                                              ahead ? info : nullptr, h->piv_tol, h->ct_dbg);
        CKL();
        continue;
      }
      double* pan = A + (int64_t)(k0 + NB) * Mp + k0;  // L[k0+NB:, k0:k0+NB] = A[k0+NB:, k0:k0+NB] * T_k^T  (in place, one n-tile)
      GemmP p = gemm_basic(pan, Mp, sM, h->Tblk + (int64_t)k * NB * NB, NB, sM, pan, Mp, sM, rem, NB, NB, 1.0, 0.0);
      RUN(launch_gemm(h, st, EPI_STORE, p, batch));
      // A[k0+NB:, k0+NB:] -= panel * panel^T   (lower tiles only)
      GemmP u = gemm_basic(pan, Mp, sM, pan, Mp, sM, A + (int64_t)(k0 + NB) * (Mp + 1), Mp, sM, rem, rem, NB, -1.0, 1.0);
      u.sym = 2;
      RUN(launch_gemm(h, st, EPI_STORE, u, batch));
    }
  }
  // recursive-doubling triangular inverse.  L^-T is kept in step with L^-1: the block-diagonal start writes both (in the launch that
  // merges the panels into A, on the fused plan), and a level whose second product runs on the small-tile kernel stores its result
  // block transposed as well (otherwise: one transpose per level)
  if (factored) {
  } else if (h->chol_fused) {
    k_tril_merge<<<g16, b16, 0, st>>>(A, h->T1, Mp, sM, sM, h->Tblk, sM, Linv, LinvT);
    CKL();
  } else {
    k_tril<<<g16, b16, 0, st>>>(A, Mp, sM);
    CKL();
    k_init_blockdiag<<<g16, b16, 0, st>>>(Linv, Mp, sM, h->Tblk, sM, LinvT);
    CKL();
  }
  const dim3 gt(Mp / 32, Mp / 32, batch), bt(32, 8);
  bool lt_current = true;
  for (int s = NB; s < Mp; s *= 2) {
    if (!lt_current) {
      k_transpose<<<gt, bt, 0, st>>>(Linv, LinvT, Mp, sM);
      CKL();
      lt_current = true;
    }
    const int npairs = Mp / (2 * s);
    const int64_t sPair = (int64_t)2 * s * (Mp + 1);
    // C1T = Inv11^T-rows x L21-rows :  Wk[pair upper-right block][j,i] = sum_k Inv11T[j,k] L21[i,k]
    GemmP p1 = gemm_basic(LinvT, Mp, sM, A + (int64_t)s * Mp, Mp, sM, h->Wk + s, Mp, sM, s, s, s, 1.0, 0.0, KM_A_UPPER);
    p1.nz2 = npairs; p1.sA2 = sPair; p1.sB2 = sPair; p1.sC2 = sPair;
    RUN(launch_gemm(h, st, EPI_STORE, p1, batch));
    // X = -Inv22 * C1 :  Linv[pair lower-left block][i,j] = -sum_k Inv22[i,k] C1T[j,k]
    GemmP p2 = gemm_basic(Linv + (int64_t)s * (Mp + 1), Mp, sM, h->Wk + s, Mp, sM, Linv + (int64_t)s * Mp, Mp, sM, s, s, s,
                          -1.0, 0.0, KM_A_LOWER);
    p2.nz2 = npairs; p2.sA2 = sPair; p2.sB2 = sPair; p2.sC2 = sPair;
    if (mm64_eligible(h, EPI_STORE, p2, batch)) {
      p2.Ct = LinvT + s;          // the pair's upper-right block of L^-T (pair stride and batch stride as for C)
      p2.ldct = Mp;
    } else {
      lt_current = false;
    }
    RUN(launch_gemm(h, st, EPI_STORE, p2, batch));
  }
  if (!lt_current) {
    k_transpose<<<gt, bt, 0, st>>>(Linv, LinvT, Mp, sM);
    CKL();
  }
  return 0;
}

// The blocked factorisation is ~70 short kernels: replay it as a CUDA graph (captured once per operand set) so the m x m
// section is not launch-latency bound.  info is staged through a handle-owned buffer so the captured pointers stay valid.
// `beside_build`: the caller knows that a tile build may be running next to this factorisation (ggp_sgpr_factor): take the
// cluster-resident plan, which keeps its own SMs, instead of the chain of launches that would queue behind the build's CTAs.
static int chol_and_inverse(ggp_handle* h, cudaStream_t st, double* A, double* Linv, double* LinvT, int batch, int32_t* info,
                            bool beside_build = false) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  CK(cudaStreamIsCapturing(st, &cs));
  bool cluster = beside_build && h->chol_cluster && batch == 1 && h->Mp >= 4 * NB && cs == cudaStreamCaptureStatusNone && h->use_graphs;
  if (cs != cudaStreamCaptureStatusNone || !h->use_graphs) {
    CK(cudaMemsetAsync(info, 0, sizeof(int32_t) * batch, st));
    return chol_and_inverse_launches(h, st, A, Linv, LinvT, batch, info);
  }
  if (cluster) {
    // fork to the handle's high-priority stream: [zero info + the step counters, one cluster launch], join
    if (!h->hp_stream) {
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&h->hp_stream, cudaStreamNonBlocking, hi));
      CK(cudaEventCreateWithFlags(&h->ev_hp0, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_hp1, cudaEventDisableTiming));
      CK(cudaMalloc((void**)&h->chol_ctr, 128 * sizeof(int)));
    }
    CK(cudaEventRecord(h->ev_hp0, st));
    CK(cudaStreamWaitEvent(h->hp_stream, h->ev_hp0, 0));
    CK(cudaMemsetAsync(h->info_ws, 0, sizeof(int32_t), h->hp_stream));
    CK(cudaMemsetAsync(h->chol_ctr, 0, 128 * sizeof(int), h->hp_stream));
    int ccn = CC_N;
    if (const char* e = getenv("GGP_CHOL_CLUSTER_N")) {   // developer A/B: 16 needs the non-portable cluster size
      ccn = atoi(e) == 16 ? 16 : (atoi(e) == 4 ? 4 : CC_N);
      if (ccn == 16) (void)cudaFuncSetAttribute(k_chol_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(ccn, 1, 1);
    lc.blockDim = dim3(256, 1, 1);
    lc.dynamicSmemBytes = CT_SMEM;
    lc.stream = h->hp_stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ccn; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    // the explicit inverse rides in the same launch when the tile build beside it is long enough to cover it (the 8 SMs need ~1.1 ms for
    // factor + inverse at Mp = 1024, the build 5.8 us per 1000 rows); behind a short build the inverse is quicker as launches on all SMs
    // (one rank's share of the 8-GPU run, 125 000 rows: 8.90 ms with the inverse in the cluster launch, 8.50 without)
    // Threshold: validated with whole bench runs at 1e6 and 5e5 local rows.  At 4e5 rows the two plans are within 1 % (24.17 against 24.39 ms
    // per step; one earlier call on another box had shown 30.2 against 24.0, not reproduced since: scripts/r2_anom.sh, scripts/step_times.py),
    // below that the inverse as launches wins -- so the plan is kept to the sizes where the build is at least twice as long as the launch
    const bool inv_in_cluster = !getenv("GGP_CHOL_CLUSTER_NO_INV") && (h->n_local >= 450000 || getenv("GGP_CHOL_CLUSTER_INV"));
    double* nul = nullptr;
    const cudaError_t le = cudaLaunchKernelEx(&lc, k_chol_cluster, A, h->Mp, h->Tblk, h->info_ws, (const double*)h->piv_tol, h->chol_ctr,
                                              inv_in_cluster ? Linv : nul, inv_in_cluster ? LinvT : nul, h->Wk);
    if (le == cudaSuccess) {
      h->launches++;
    } else {
      // a device / partition that cannot co-schedule the cluster (8 SMs of one GPC with 52 KB of shared memory and a full register file
      // each): not an error of the evaluation -- clear it and use the launch chain from now on
      (void)cudaGetLastError();
      h->chol_cluster = false;
      cluster = false;
    }
    CK(cudaEventRecord(h->ev_hp1, h->hp_stream));
    CK(cudaStreamWaitEvent(st, h->ev_hp1, 0));
    if (cluster && inv_in_cluster) {   // factor, inverse and transposed inverse all came out of the one launch
      CK(cudaMemcpyAsync(info, h->info_ws, sizeof(int32_t) * batch, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
  }
  ggp_handle::CholGraph* g = nullptr;
  for (auto& c : h->chol_graphs)
    if (c.A == A && c.Linv == Linv && c.LinvT == LinvT && c.batch == batch && c.factored == cluster) g = &c;
  if (!g) {
    ggp_handle::CholGraph c{};
    c.A = A; c.Linv = Linv; c.LinvT = LinvT; c.batch = batch; c.factored = cluster;
    const long long before = h->launches;
    cudaGraph_t graph = nullptr;
    // capture on a private stream: the caller's stream may be the legacy default stream, which cannot be captured
    if (!h->cap_stream) CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    CK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = chol_and_inverse_launches(h, h->cap_stream, A, Linv, LinvT, batch, h->info_ws, cluster);
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc != 0) return rc;
    CK(e);
    c.nodes = h->launches - before;
    h->launches = before;
    CK(cudaGraphInstantiate(&c.exec, graph, 0));
    cudaGraphDestroy(graph);
    h->chol_graphs.push_back(c);
    g = &h->chol_graphs.back();
  }
  if (!cluster) CK(cudaMemsetAsync(h->info_ws, 0, sizeof(int32_t) * batch, st));
  CK(cudaGraphLaunch(g->exec, st));
  h->launches += g->nodes;
  CK(cudaMemcpyAsync(info, h->info_ws, sizeof(int32_t) * batch, cudaMemcpyDeviceToDevice, st));
  return 0;
}


// ---- sliced-integer (tcgen05 kind::i8) GEMM: host side -----------------------------------------------------------------
// digit planes [I8_NS][rows][ld bytes] (plane stride in bytes) as a 3-D map (k, row, plane), 64-byte swizzle, box 64 x box_rows x 1;
// rows / k beyond the extents given here are zero-filled by the TMA unit
static bool make_i8_map(CUtensorMap* tm, const int8_t* ptr, int64_t rows, int64_t kbytes, int64_t ld, int64_t plane, int box_rows,
                        int nplanes = I8_NS, int box_inner = I8_BKB) {
  tmap_encode_fn enc = get_tmap_encode();
  if (!enc || !ptr || rows < 1 || kbytes < 1) return false;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 15) || (plane & 15)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)kbytes, (cuuint64_t)rows, (cuuint64_t)nplanes};
  const cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)plane};
#ifdef GGP_I8_TMA_PLANES7
  const cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, (cuuint32_t)I8_NS};
#else
  const cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1u};
#endif
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct I8Operand { const int8_t* q; int64_t rows, ld, plane; };

// Tile width of a launch.  Production: 128 x 64 single-buffered tiles for every role.  The 128 x 32 variant with two TMEM accumulator
// sets (the hand-over and the epilogue of tile t overlap the MMAs of tile t + 1) is compiled only with -DGGP_I8_ENABLE_BN32 and selected
// by GGP_I8_BN32=1: MEASURED SLOWER at the headline shape (backward GEMM 34.2 vs 27.8 ms, triangular multiply 15.9 vs 14.8 ms,
// profiles/r2_summary.md) -- a 128 x 32 tile re-reads the A operand for half the work, its 7 narrower MMAs per k-step already need
// ~124 B/clk of shared-memory operand reads (probe: 92 % of the 128 x 64 mix), and the per-thread multiplier loads of the moments
// epilogue (a latency chain per tile, not per MAC) no longer fit in two tile periods.  Kept as a measured experiment, not a product path.
static int i8_tile_width(int epi, const I8P& p) {
#ifdef GGP_I8_ENABLE_BN32
  if (getenv("GGP_I8_BN32") && !p.nchunk && !p.sym) return 32;
#endif
  (void)epi; (void)p;
  return 64;
}

template <int EPI, int BN>
static void launch_i8_kernel(int grid, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const I8P& p) {
#ifndef GGP_I8_ENABLE_BN32
  if (BN == 32) return;   // not compiled in
  k_gemm_i8<EPI, 64><<<grid, I8_THREADS, i8_smem<EPI, 64>(), st>>>(tmA, tmB, p);
#else
  k_gemm_i8<EPI, BN><<<grid, I8_THREADS, i8_smem<EPI, BN>(), st>>>(tmA, tmB, p);
#endif
}

static int launch_i8(ggp_handle* h, cudaStream_t st, int epi, I8P p, const I8Operand& A, const I8Operand& B) {
  if (p.K < 1 || p.M < 1 || p.N < 1) return 0;
  if (p.K > I8_MAX_K) return fail(-4, "launch_i8: k extent exceeds the exact int32 accumulation bound (16384)");
  if (epi == I8_EPI_SLICE && (p.eb || p.alpha != 1.0 || p.K > I8_K_GROUP4))
    return fail(-4, "launch_i8: the digit-plane epilogue takes a scalar column exponent, alpha = 1 and k <= 4096");
  const int bn = i8_tile_width(epi, p);
  p.tiles_m = (p.M + I8_BM - 1) / I8_BM;
  p.tiles_n = (p.N + bn - 1) / bn;
  int tiles = p.tiles_m * p.tiles_n;
  if (p.sym) {
    tiles = 0;
    for (int tm = 0; tm < p.tiles_m; ++tm) tiles += std::max(0, p.tiles_n - 2 * tm);
  }
  if (p.splits < 1) p.splits = 1;
  p.total = tiles * p.splits;
  if (p.nchunk) {   // one launch over all k-chunks (digit planes stacked along the plane axis)
    if (epi != I8_EPI_F64 || p.splits != 1 || tiles > h->sm_count) return fail(-4, "launch_i8: chunked accumulation needs the F64 epilogue, no split-K, tiles <= SMs");
    p.ntile = tiles;
    p.total = tiles * p.nchunk;
  }
  if (getenv("GGP_I8_EXP_SKIPB")) p.exp_skip_b = 1;
  if (getenv("GGP_I8_EXP_SKIPA")) p.exp_skip_a = 1;
  if (getenv("GGP_I8_EXP_NOEPI")) p.exp_no_epi = 1;
  CUtensorMap tmA, tmB;
  const int npl = p.nchunk ? I8_NS * p.nchunk : I8_NS;
  bool okB;
  if (p.b_mn) {
    // MN-major B: planes [plane][k rows = B.rows][columns, contiguous]; the map's inner extent is one column block (chunk array), the
    // box bn column bytes x 64 k-rows (64-byte swizzle for 64 columns, 32-byte for 32); rows / columns beyond the extents are zero-filled
    if (p.sym || p.nchunk || p.splits != 1 || p.lower_a) return fail(-4, "launch_i8: the MN-major B operand serves the plain product only");
    const int64_t cols = p.b_chunk > 0 ? p.b_chunk : (int64_t)(p.N + 63) / 64 * 64;
    okB = make_i8_map(&tmB, B.q, B.rows, cols, B.ld, B.plane, I8_BKB, p.b_planes > 0 ? p.b_planes : I8_NS, bn);
  } else {
    okB = make_i8_map(&tmB, B.q, B.rows, p.K, B.ld, B.plane, bn, npl);
  }
  if (!make_i8_map(&tmA, A.q, A.rows, p.K, A.ld, A.plane, I8_BM, npl) || !okB)
    return fail(-4, "launch_i8: operand planes cannot be described by a tensor map (alignment)");
  int grid = std::min(p.total, h->sm_count);
  if (p.nchunk) grid = p.ntile * std::max(1, std::min(std::min(h->sm_count / p.ntile, p.nchunk), p.nchunk_groups_max > 0 ? p.nchunk_groups_max : p.nchunk));
  if (epi == I8_EPI_MOMENTS && p.mom_accum) {
    if (p.tiles_m > h->sm_count || 2 * p.d + 1 > 24 || p.sym || p.splits != 1)
      return fail(-4, "launch_i8: mom_accum needs tiles_m <= SM count, 2 d + 1 <= 24, no symmetry and no split-K");
    grid = p.tiles_m * std::max(1, std::min(h->sm_count / p.tiles_m, p.tiles_n));
  }
  const char* tl = getenv("GGP_I8_TIMELINE");
  const bool dbg = tl && tl[0] == '2' && h->i8_dbg_prints < 9 && !p.dbg;
  if (dbg) {
    if (!h->i8_dbg) CK(cudaMalloc((void**)&h->i8_dbg, (3 * I8_DBG_ITEMS * 4 + 2 * 256) * sizeof(long long)));
    CK(cudaMemsetAsync(h->i8_dbg, 0, (3 * I8_DBG_ITEMS * 4 + 2 * 256) * sizeof(long long), st));
    p.dbg = h->i8_dbg;
  }
  if (epi == I8_EPI_F64) { if (bn == 64) launch_i8_kernel<I8_EPI_F64, 64>(grid, st, tmA, tmB, p); else launch_i8_kernel<I8_EPI_F64, 32>(grid, st, tmA, tmB, p); }
  else if (epi == I8_EPI_SLICE) { if (bn == 64) launch_i8_kernel<I8_EPI_SLICE, 64>(grid, st, tmA, tmB, p); else launch_i8_kernel<I8_EPI_SLICE, 32>(grid, st, tmA, tmB, p); }
  else { if (bn == 64) launch_i8_kernel<I8_EPI_MOMENTS, 64>(grid, st, tmA, tmB, p); else launch_i8_kernel<I8_EPI_MOMENTS, 32>(grid, st, tmA, tmB, p); }
  CKL();
  if (dbg) {
    CK(cudaStreamSynchronize(st));
    long long hb[3 * I8_DBG_ITEMS * 4 + 2 * 256];
    CK(cudaMemcpy(hb, h->i8_dbg, sizeof(hb), cudaMemcpyDeviceToHost));
    {
      const long long* g = hb + 3 * I8_DBG_ITEMS * 4;
      long long s0 = g[0], s1 = g[0], e0 = g[1], e1 = g[1];
      for (int b = 0; b < grid && b < 256; ++b) {
        s0 = std::min(s0, g[2 * b]); s1 = std::max(s1, g[2 * b]);
        e0 = std::min(e0, g[2 * b + 1]); e1 = std::max(e1, g[2 * b + 1]);
      }
      fprintf(stderr, "== CTA spans (ns): starts within %lld, first end +%lld, last end +%lld; CTA0 %lld..%lld\n", s1 - s0, e0 - s0, e1 - s0,
              g[0] - s0, g[1] - s0);
    }
    const long long t0 = hb[0];
    fprintf(stderr, "== i8 timeline, epilogue %d, M=%d N=%d K=%d lower=%d sym=%d splits=%d\n", epi, p.M, p.N, p.K, p.lower_a, p.sym, p.splits);
    for (int it = 0; it < I8_DBG_ITEMS; ++it) {
      const long long* m = hb + it * 4;
      const long long* ep = hb + (I8_DBG_ITEMS + it) * 4;
      fprintf(stderr, "tile %2d  mma: start %8lld tmem_free %8lld first_kb %8lld all_issued %8lld | epi: wait %8lld full %8lld drained %8lld done %8lld\n",
              it, m[0] - t0, m[1] - t0, m[2] - t0, m[3] - t0, ep[0] - t0, ep[1] - t0, ep[2] - t0, ep[3] - t0);
      const long long* pw = hb + (2 * I8_DBG_ITEMS + it) * 4;
      if (pw[0]) fprintf(stderr, "         pre-wait: bar1 %8lld x-staged %8lld kv-issued %8lld bar2 %8lld\n", pw[0] - t0, pw[1] - t0, pw[2] - t0, pw[3] - t0);
    }
    h->i8_dbg_prints++;
  }
  return 0;
}

static bool use_i8(const ggp_handle* h, const ggp_cfg* cfg, int d, int batch) {
  return cfg && cfg->precision == GGP_PREC_FP64_I8 && cfg->kernel != GGP_KERNEL_COMPOSITE && batch == 1 && h->Mp >= I8_BM && h->Mp <= I8_K_GROUP4 && d <= I8_MAX_D &&
         h->arena_i8 != nullptr;
}

// The Kzz-part chain of ggp_sgpr_finish (G_zz = -1/2 L^{-T} Gbar L^{-1}, its reduction against dKzz/dtheta, the assembly of grad_mm) is
// not on the critical path: pass 2 only needs Q and u.  It runs on the handle's auxiliary stream NEXT TO pass 2 (whose one-launch
// plans leave a few SMs free: 8 x 18 = 144 CTAs on 148 SMs) and is joined when pass 2 has been enqueued -- or by the next entry point
// that touches its buffers.  Legal under stream capture (fork and join both lie inside the captured evaluation).
static int join_aux(ggp_handle* h, cudaStream_t st) {
  if (h->aux_pending) {
    CK(cudaStreamWaitEvent(st, h->ev_join, 0));
    h->aux_pending = false;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
extern "C" {

int ggp_version(void) { return 100; }
const char* ggp_last_error(void) { return g_err; }

int ggp_create(ggp_handle_t** out, int device) {
  if (!out) return fail(-1, "ggp_create: out is NULL");
  CK(cudaSetDevice(device));
  ggp_handle* h = new ggp_handle();
  h->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  CK(cudaFuncSetAttribute(k_gemm_nt<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  CK(cudaFuncSetAttribute(k_gemm_nt<EPI_MOMENTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  CK(cudaFuncSetAttribute(k_gemm_tma<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
  CK(cudaFuncSetAttribute(k_gemm_tma<EPI_MOMENTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
  CK(cudaFuncSetAttribute(k_build_kc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_chol_trail, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
  { const char* e = getenv("GGP_CHOL_FUSED"); h->chol_fused = !(e && e[0] == '0'); }
  { const char* e = getenv("GGP_CHOL_LOOKAHEAD"); h->chol_lookahead = !(e && e[0] == '0'); }
  { const char* e = getenv("GGP_CHOL_SMALL"); h->chol_small = !(e && e[0] == '0'); }
  { const char* e = getenv("GGP_CHOL_CLUSTER"); h->chol_cluster = !(e && e[0] == '0'); }
  CK(cudaFuncSetAttribute(k_chol_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
  CK(cudaFuncSetAttribute(k_chol_inv_small, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
  { const char* e = getenv("GGP_MM64_MAX_TILES"); if (e) h->mm64_max_tiles = atoi(e); }
  CK(cudaFuncSetAttribute(k_mm64, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
  CK(cudaFuncSetAttribute(k_gemm_i8<I8_EPI_F64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, i8_smem<I8_EPI_F64, 64>()));
  CK(cudaFuncSetAttribute(k_gemm_i8<I8_EPI_SLICE, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, i8_smem<I8_EPI_SLICE, 64>()));
  CK(cudaFuncSetAttribute(k_gemm_i8<I8_EPI_MOMENTS, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, i8_smem<I8_EPI_MOMENTS, 64>()));
#ifdef GGP_I8_ENABLE_BN32
  CK(cudaFuncSetAttribute(k_gemm_i8<I8_EPI_F64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, i8_smem<I8_EPI_F64, 32>()));
  CK(cudaFuncSetAttribute(k_gemm_i8<I8_EPI_SLICE, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, i8_smem<I8_EPI_SLICE, 32>()));
  CK(cudaFuncSetAttribute(k_gemm_i8<I8_EPI_MOMENTS, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, i8_smem<I8_EPI_MOMENTS, 32>()));
#endif
  *out = h;
  return 0;
}

int ggp_destroy(ggp_handle_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (auto& c : h->chol_graphs) cudaGraphExecDestroy(c.exec);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->arena) cudaFree(h->arena);
  if (h->kc_all) cudaFree(h->kc_all);
  if (h->kq_all) cudaFree(h->kq_all);
  if (h->atq_all) cudaFree(h->atq_all);
  if (h->arena_i8) cudaFree(h->arena_i8);
  if (h->krow) cudaFree(h->krow);
  if (h->ct_dbg) cudaFree(h->ct_dbg);
  if (h->chol_ctr) cudaFree(h->chol_ctr);
  if (h->hp_stream) cudaStreamDestroy(h->hp_stream);
  if (h->ev_hp0) cudaEventDestroy(h->ev_hp0);
  if (h->ev_hp1) cudaEventDestroy(h->ev_hp1);
  delete h;
  return 0;
}

int ggp_workspace_bytes(const ggp_cfg* cfg, int64_t n_local, int m, int d, int batch, size_t* out) {
  if (!out || m <= 0 || d <= 0 || batch <= 0 || n_local < 0) return fail(-1, "ggp_workspace_bytes: bad argument");
  const Plan p = make_plan(cfg, n_local, m, d, batch, 148);
  size_t total = p.bytes;
  const size_t rows = (size_t)((n_local + 127) / 128 * 128);
  const size_t cache = (size_t)batch * rows * p.Mp * 8;
  const size_t budget = (cfg && cfg->tile_cache_mib > 0) ? (size_t)cfg->tile_cache_mib * 1024 * 1024 : 0;
  const bool have_cache = n_local > 0 && cache <= budget;
  if (have_cache) total += cache;
  if (cfg && cfg->precision == GGP_PREC_FP64_I8 && batch == 1 && p.Mp >= I8_BM) {
    // digit planes of L^{-1}, Q, one chunk of A^T and of k(X,Z), exponents (ggp_reserve: arena_i8) ...
    total += 2 * align_up((size_t)I8_NS * p.Mp * p.Mp, 256) + 2 * align_up((size_t)I8_NS * p.Mp * p.nc, 256) + 3 * align_up((size_t)p.Mp * 4, 256);
    // ... and, with the tile cache, the digit planes of k(X,Z) and of A^T over all local rows
    const size_t needq = rows * p.Mp * I8_NS;
    if (have_cache && 2 * needq <= budget) {
      total += needq;
      if (8 * rows * p.Mp + 2 * needq <= budget) total += (size_t)((n_local + p.nc - 1) / p.nc) * I8_NS * p.Mp * p.nc;
    } else if (have_cache) {
      total -= cache;   // the cache holds both the FP64 tiles and their digit planes, or neither
    }
  }
  *out = total;
  return 0;
}

int ggp_reserve(ggp_handle_t* h, const ggp_cfg* cfg, int64_t n_local, int m, int d, int batch) {
  if (!h) return fail(-1, "ggp_reserve: handle is NULL");
  if (m <= 0 || d <= 0 || batch <= 0 || n_local < 0) return fail(-2, "ggp_reserve: bad shape");
  if ((size_t)(2 * KT_N * d + d) * 8 > 200 * 1024) return fail(-3, "ggp_reserve: input dimension d too large");
  CK(cudaSetDevice(h->device));
  if (h->aux_stream) CK(cudaStreamSynchronize(h->aux_stream));
  h->aux_pending = false;
  Plan p = make_plan(cfg, n_local, m, d, batch, h->sm_count);
  for (auto& c : h->chol_graphs) cudaGraphExecDestroy(c.exec);
  h->chol_graphs.clear();
  if (p.bytes > h->arena_bytes) {
    if (h->arena) CK(cudaFree(h->arena));
    h->arena = nullptr;
    h->arena_bytes = 0;
    CK(cudaMalloc((void**)&h->arena, p.bytes));
    h->arena_bytes = p.bytes;
  }
  h->n_local = n_local; h->m = m; h->d = d; h->batch = batch;
  h->Mp = p.Mp; h->nc = p.nc; h->splits = p.splits;
  h->kc_valid = false;
  h->pf_valid = false;
  h->atq_valid = false;
  {
    const int64_t rows = (n_local + 127) / 128 * 128;
    const size_t need = (size_t)batch * rows * p.Mp * 8;
    const size_t budget = (cfg && cfg->tile_cache_mib > 0) ? (size_t)cfg->tile_cache_mib * 1024 * 1024 : 0;
    if (n_local > 0 && need <= budget) {
      if (need > h->kc_all_bytes) {
        if (h->kc_all) CK(cudaFree(h->kc_all));
        h->kc_all = nullptr;
        h->kc_all_bytes = 0;
        if (cudaMalloc((void**)&h->kc_all, need) == cudaSuccess) {
          h->kc_all_bytes = need;
        } else {   // not enough memory for the cache: fall back to rebuilding the tiles in pass 2
          (void)cudaGetLastError();
          h->kc_all = nullptr;
        }
      }
      h->kc_rows = rows;
    } else if (h->kc_all) {
      CK(cudaFree(h->kc_all));
      h->kc_all = nullptr;
      h->kc_all_bytes = 0;
    }
  }
  // sliced-integer path: digit planes of L^{-1}, P, one chunk of A^T and one chunk of k(X,Z) (+ the whole-N cache next to kc_all)
  if (cfg && cfg->precision == GGP_PREC_FP64_I8 && batch == 1 && p.Mp >= I8_BM) {
    const size_t MMq = align_up((size_t)I8_NS * p.Mp * p.Mp, 256), CHq = align_up((size_t)I8_NS * p.Mp * p.nc, 256),
                 EX = align_up((size_t)p.Mp * 4, 256);
    const size_t need = 2 * MMq + 2 * CHq + 3 * EX;
    if (need > h->arena_i8_bytes) {
      if (h->arena_i8) CK(cudaFree(h->arena_i8));
      h->arena_i8 = nullptr;
      h->arena_i8_bytes = 0;
      CK(cudaMalloc((void**)&h->arena_i8, need));
      h->arena_i8_bytes = need;
    }
    char* q = h->arena_i8;
    h->Lq = (int8_t*)q; q += MMq;
    h->Pq = (int8_t*)q; q += MMq;
    h->Atq = (int8_t*)q; q += CHq;
    h->Kq = (int8_t*)q; q += CHq;
    h->eL = (int*)q; q += EX;
    h->eP = (int*)q; q += EX;
    h->eKA = (int*)q;
    if (h->kc_all) {
      const size_t needq = (size_t)h->kc_rows * p.Mp * I8_NS;
      const size_t budget = (size_t)cfg->tile_cache_mib * 1024 * 1024;
      bool ok = 2 * needq <= budget;
      if (ok && needq > h->kq_all_bytes) {
        if (h->kq_all) CK(cudaFree(h->kq_all));
        h->kq_all = nullptr;
        h->kq_all_bytes = 0;
        if (cudaMalloc((void**)&h->kq_all, needq) == cudaSuccess) h->kq_all_bytes = needq;
        else { (void)cudaGetLastError(); ok = false; }
      }
      if (ok && 8 * (size_t)h->kc_rows * p.Mp + 2 * needq <= budget && !getenv("GGP_I8_NO_ATQ_ALL")) {   // optional third array
        const size_t needa = (size_t)((n_local + p.nc - 1) / p.nc) * I8_NS * p.Mp * p.nc;   // one [plane][m][nc] array per chunk
        if (needa > h->atq_all_bytes) {
          if (h->atq_all) CK(cudaFree(h->atq_all));
          h->atq_all = nullptr;
          h->atq_all_bytes = 0;
          if (cudaMalloc((void**)&h->atq_all, needa) == cudaSuccess) h->atq_all_bytes = needa;
          else (void)cudaGetLastError();
        }
      } else if (h->atq_all) {
        CK(cudaFree(h->atq_all));
        h->atq_all = nullptr;
        h->atq_all_bytes = 0;
      }
      if (!ok) {   // the cache must hold both the FP64 tiles and their digit planes, or neither
        CK(cudaFree(h->kc_all));
        h->kc_all = nullptr;
        h->kc_all_bytes = 0;
        if (h->kq_all) { CK(cudaFree(h->kq_all)); h->kq_all = nullptr; h->kq_all_bytes = 0; }
        if (h->atq_all) { CK(cudaFree(h->atq_all)); h->atq_all = nullptr; h->atq_all_bytes = 0; }
      }
    }
  }
  double** slots[] = {&h->L, &h->Linv, &h->LinvT, &h->Wk, &h->Bm, &h->LBinv, &h->LBinvT, &h->Binv, &h->PA, &h->Gbar, &h->T1,
                      &h->P, &h->Gzz, &h->Tblk, &h->bvec, &h->cvec, &h->beta, &h->u, &h->yty, &h->ds2, &h->rowacc, &h->Kc,
                      &h->At, &h->Spart, &h->mom_part, &h->mom_acc, &h->rk};
  const size_t nslots = sizeof(slots) / sizeof(slots[0]);
  for (size_t i = 0; i < nslots; ++i) *slots[i] = reinterpret_cast<double*>(h->arena + p.off[i]);
  h->info_ws = reinterpret_cast<int32_t*>(h->arena + p.off[nslots]);
  for (int i = 0; i < 5; ++i) h->sv[i] = reinterpret_cast<double*>(h->arena + p.off[nslots + 1 + i]);
  h->rowout = reinterpret_cast<double*>(h->arena + p.off[nslots + 6]);
  h->piv_tol = reinterpret_cast<double*>(h->arena + p.off[nslots + 7]);
  h->ysc = reinterpret_cast<double*>(h->arena + p.off[nslots + 8]);
  h->rk_part = reinterpret_cast<double*>(h->arena + p.off[nslots + 9]);
  h->nsv = p.nsv;
  return 0;
}

// composite kernel: the program is registered for this input dimension and the parameter rows are set
static int prog_ready(const ggp_handle* h, int d, const char* who) {
  if (!h->prog_set || h->prog.d != d || !h->kth) {
    snprintf(g_err, sizeof(g_err), "%s: GGP_KERNEL_COMPOSITE needs ggp_set_kernel_program (same d) and ggp_set_kernel_params first", who);
    return -3;
  }
  return 0;
}
static bool kernel_ok(const ggp_handle* h, KSpec kind) {
  if (kind.kind == GGP_KERNEL_COMPOSITE) return h->prog_set;
  return kind.kind >= GGP_KERNEL_RBF && kind.kind <= GGP_KERNEL_RQ && (kind.kind != GGP_KERNEL_RQ || kind.p > 0.0);
}

int ggp_kprog_nparams(const ggp_kprog* prog, int d) {
  KProgDev pd;
  if (!kprog_compile(prog, d, &pd)) return fail(-1, "ggp_kprog_nparams: malformed program (terms 1..6, factors 1..3, kinds 0..4, d 1..16, <= 64 parameters)");
  return pd.P;
}
int ggp_set_kernel_program(ggp_handle_t* h, const ggp_kprog* prog, int d) {
  if (!h || !prog) return fail(-1, "ggp_set_kernel_program: NULL argument");
  if (h->m <= 0 || h->d != d) return fail(-2, "ggp_set_kernel_program: reserve the handle for this input dimension first");
  KProgDev pd;
  if (!kprog_compile(prog, d, &pd)) return fail(-3, "ggp_set_kernel_program: malformed program (terms 1..6, factors 1..3, kinds 0..4, d 1..16, <= 64 parameters)");
  const size_t need = (size_t)std::max(1, h->batch) * h->m * (pd.P + d) * sizeof(double);
  if (need > h->krow_bytes) {
    CK(cudaDeviceSynchronize());
    if (h->krow) cudaFree(h->krow);
    h->krow = nullptr; h->krow_bytes = 0;
    CK(cudaMalloc(&h->krow, need));
    h->krow_bytes = need;
  }
  h->prog = pd;
  h->prog_set = true;
  h->kth = nullptr; h->kgrad_mm = nullptr; h->kgrad_partial = nullptr;
  h->kc_valid = false; h->pf_valid = false; h->atq_valid = false;
  return 0;
}
int ggp_set_kernel_params(ggp_handle_t* h, const double* kparams, double* kgrad_mm, double* kgrad_partial) {
  if (!h || !kparams) return fail(-1, "ggp_set_kernel_params: NULL argument");
  if (!h->prog_set) return fail(-2, "ggp_set_kernel_params: no program registered");
  h->kth = kparams; h->kgrad_mm = kgrad_mm; h->kgrad_partial = kgrad_partial;
  return 0;
}

int ggp_sgpr_expect_prefetch(ggp_handle_t* h, int on) {
  if (!h) return fail(-1, "ggp_sgpr_expect_prefetch: handle is NULL");
  h->pf_expected = on != 0;
  return 0;
}

int ggp_sgpr_factor(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* Z, const double* theta,
                    const double* jitter, int m, int d, int batch, int32_t* info) {
  if (!h || !Z || !theta || !info) return fail(-1, "ggp_sgpr_factor: NULL argument");
  if (!reserved_for(h, 0, m, d, batch)) return fail(-2, "ggp_sgpr_factor: handle not reserved for this shape");
  cudaStream_t st = (cudaStream_t)stream;
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_sgpr_factor: unknown kernel");
  const int Mp = h->Mp;
  h->kc_valid = false;
  h->atq_valid = false;
  RUN(join_aux(h, st));
  ProfScope ps(h, st, CAT_MM);
  if (kind == GGP_KERNEL_COMPOSITE) {
    if (int r = prog_ready(h, d, "ggp_sgpr_factor")) return r;
    k_build_kzz_prog<<<dim3(Mp / 16, Mp / 16, batch), dim3(16, 16), 0, st>>>(Z, m, Mp, h->prog, h->kth, jitter, h->L, (int64_t)Mp * Mp, h->piv_tol);
  } else {
    k_build_kzz<<<dim3(Mp / 16, Mp / 16, batch), dim3(16, 16), 0, st>>>(Z, m, Mp, d, theta, jitter, kind, h->L, (int64_t)Mp * Mp, h->piv_tol);
  }
  CKL();
  // a tile prefetch has been (ggp_sgpr_prefetch_tiles[_part] on another stream) or is about to be (ggp_sgpr_expect_prefetch) enqueued
  // for this evaluation: its CTAs fill every SM
  const bool beside = h->pf_valid || h->pf_next_row > 0 || h->pf_expected;
  h->pf_expected = false;
  return chol_and_inverse(h, st, h->L, h->Linv, h->LinvT, batch, info, beside);
}

static int build_chunk(ggp_handle* h, cudaStream_t st, const double* Xc, int nv, int d, const double* Z, int m,
                       const double* theta, KSpec kind, int batch, double* dst, int64_t sK) {
  const int Mp = h->Mp;
  if (kind == GGP_KERNEL_COMPOSITE) {
    if (int r = prog_ready(h, d, "build_chunk")) return r;
    k_build_kc_prog<<<dim3(Mp / 32, (nv + 7) / 8, batch), 256, 0, st>>>(Xc, nv, Z, m, Mp, h->prog, h->kth, dst, Mp, sK);
    CKL();
    return 0;
  }
  dim3 grid(Mp / KT_M, (nv + KT_N - 1) / KT_N, batch);
  const size_t smem = (size_t)(KT_N * d + KT_M * d + d) * 8;
  k_build_kc<<<grid, KT_THREADS, smem, st>>>(Xc, nv, nv, d, Z, m, theta, kind, dst, Mp, sK);
  CKL();
  return 0;
}

// sliced-integer path: FP64 tile + digit planes in one kernel
static int build_chunk_i8(ggp_handle* h, cudaStream_t st, const double* Xc, int64_t nv, int d, const double* Z, int m,
                          const double* theta, KSpec kind, double* Kc, int8_t* Kq, int64_t plane) {
  const int Mp = h->Mp;
  const int64_t rows_per_cta = (int64_t)KT_N * KT_RT;
  dim3 grid(Mp / KT_M, (unsigned)((nv + rows_per_cta - 1) / rows_per_cta));
  const size_t smem = (size_t)(2 * KT_N * d + KT_M * d + d) * 8;
  k_build_kc_i8<<<grid, KT_THREADS, smem, st>>>(Xc, nv, d, Z, m, theta, kind, Kc, Mp, Kq, Mp, plane);
  CKL();
  return 0;
}

int ggp_sgpr_prefetch_tiles_part(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X, int64_t n_local, int64_t row0,
                                 int64_t nrows, const double* Z, const double* theta, int m, int d, int batch) {
  if (!h || !Z || !theta || (n_local > 0 && !X)) return fail(-1, "ggp_sgpr_prefetch_tiles: NULL argument");
  if (!reserved_for(h, n_local, m, d, batch)) return fail(-2, "ggp_sgpr_prefetch_tiles: handle not reserved for this shape");
  if (row0 < 0 || nrows < 0 || row0 + nrows > n_local) return fail(-3, "ggp_sgpr_prefetch_tiles_part: row range outside [0, n_local)");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_sgpr_prefetch_tiles_part: unknown kernel");
  if (row0 == 0) { h->pf_valid = false; h->pf_next_row = 0; }
  if (!h->kc_all || n_local <= 0) return 0;   // no tile cache: pass 1 builds chunk by chunk as before
  if (row0 != h->pf_next_row) { h->pf_valid = false; return fail(-3, "ggp_sgpr_prefetch_tiles_part: parts must be issued in ascending row order"); }
  cudaStream_t st = (cudaStream_t)stream;
  const bool i8 = use_i8(h, cfg, d, batch) && h->kq_all;
  const int Mp = h->Mp, nc = h->nc;
  if (!i8 && (row0 % nc) != 0) return fail(-3, "ggp_sgpr_prefetch_tiles_part: parts of the FP64 path start on a chunk boundary");
  ProfScope ps(h, st, CAT_BUILD);
  if (i8) {   // the cache is one row-major [rows][Mp] array (+ digit planes): one launch covers the whole row range
    if (nrows > 0)
      RUN(build_chunk_i8(h, st, X + row0 * d, nrows, d, Z, m, theta, kind, h->kc_all + row0 * Mp, h->kq_all + row0 * Mp, h->kc_rows * Mp));
  } else {
    for (int64_t c0 = row0; c0 < row0 + nrows; c0 += nc) {
      const int nv = (int)std::min<int64_t>(nc, row0 + nrows - c0);
      RUN(build_chunk(h, st, X + c0 * d, nv, d, Z, m, theta, kind, batch, h->kc_all + c0 * Mp, h->kc_rows * Mp));
    }
  }
  h->pf_next_row = row0 + nrows;
  if (h->pf_next_row == n_local) {
    h->pf_valid = true;
    h->pf_X = X; h->pf_Z = Z; h->pf_theta = theta; h->pf_n = n_local; h->pf_batch = batch; h->pf_kind = kind;
  }
  return 0;
}

int ggp_sgpr_prefetch_tiles(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X, int64_t n_local, const double* Z,
                            const double* theta, int m, int d, int batch) {
  return ggp_sgpr_prefetch_tiles_part(h, cfg, stream, X, n_local, 0, n_local, Z, theta, m, d, batch);
}

int ggp_sgpr_pass1(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X, const double* y, int64_t n_local,
                   const double* Z, const double* theta, int m, int d, int batch, double* partial) {
  if (!h || !Z || !theta || !partial || (n_local > 0 && (!X || !y))) return fail(-1, "ggp_sgpr_pass1: NULL argument");
  if (!reserved_for(h, n_local, m, d, batch)) return fail(-2, "ggp_sgpr_pass1: handle not reserved for this shape");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_sgpr_pass1: unknown kernel");
  if (cfg && cfg->precision == GGP_PREC_TF32X3)
    return fail(-3, "ggp_sgpr_pass1: GGP_PREC_TF32X3 is not offered: it cannot meet the gradient tolerance (DESIGN.md 4b); use "
                    "GGP_PREC_FP64 (DMMA) or GGP_PREC_FP64_I8 (exact int8 slicing on tcgen05)");
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = h->Mp, nc = h->nc, splits = h->splits;
  const int64_t sM = (int64_t)Mp * Mp;
  CK(cudaMemsetAsync(h->Spart, 0, (size_t)batch * splits * sM * 8, st));
  CK(cudaMemsetAsync(h->bvec, 0, (size_t)batch * m * 8, st));
  k_sumsq<<<SUMSQ_BLOCKS, 1024, 0, st>>>(y, n_local, h->yty);   // yty: 256 bytes = total + SUMSQ_BLOCKS partials
  CKL();
  k_sumsq_final<<<1, 1, 0, st>>>(h->yty);
  CKL();
  const bool i8 = use_i8(h, cfg, d, batch);
  // tiles already in the cache (ggp_sgpr_prefetch_tiles with the same operands, typically overlapped with the factorisation)
  const bool prefetched = h->kc_all && h->pf_valid && h->pf_X == X && h->pf_Z == Z && h->pf_theta == theta && h->pf_n == n_local &&
                          h->pf_batch == batch && h->pf_kind.kind == kind.kind && h->pf_kind.p == kind.p && (!i8 || h->kq_all);
  h->pf_valid = false;
  h->pf_next_row = 0;
  h->atq_valid = false;
  if (i8) {
    // fixed exponents of the bounded operands: k(x,z) <= sf2 and |A[m,n]| <= sqrt(k_nn) = sqrt(sf2), computed ON THE DEVICE into
    // h->eKA (the kernels read them through I8P::e_dev): the call only enqueues, no host read-back of theta -- see gemm_i8.cuh
    ProfScope ps(h, st, CAT_BUILD);
    k_i8_exponents<<<1, 1, 0, st>>>(theta, d, h->eKA);
    CKL();
    k_slice_rows<<<(Mp + 7) / 8, 256, 0, st>>>(h->Linv, Mp, Mp, Mp, h->Lq, Mp, (int64_t)Mp * Mp, Mp, h->eL);
    CKL();
    CK(cudaMemsetAsync(h->mom_part, 0, (size_t)2 * ((nc + I8_BN - 1) / I8_BN) * m * 8, st));   // b partials, accumulated over the chunks
  }
  // With the tile cache and room for the digits of A^T over all local rows, the triangular multiply is ONE launch (no per-chunk
  // tails: 2048 tiles over 148 CTAs leave the last round 16 % full; no per-launch start-up) and the b-partials stay in registers:
  // in the n-major snake order a CTA only ever works on two row tiles when 2 G is a multiple of the row-tile count.
  const int i8_tiles_m = (m + I8_BM - 1) / I8_BM;
  const int64_t i8_tiles = (int64_t)i8_tiles_m * ((n_local + I8_BN - 1) / I8_BN);
  const bool trmm_once = i8 && h->kq_all && h->atq_all && h->kc_all && i8_tiles >= h->sm_count && (2 * h->sm_count) % i8_tiles_m == 0 &&
                         i8_tiles < ((int64_t)1 << 30) && n_local > 0 && (int64_t)(nc / 32) * (2 * d + 1) >= 2 * h->sm_count /* slab room */;
  if (trmm_once) {
    if (!prefetched) {
      ProfScope ps(h, st, CAT_BUILD);
      RUN(build_chunk_i8(h, st, X, n_local, d, Z, m, theta, kind, h->kc_all, h->kq_all, h->kc_rows * Mp));
    }
    CK(cudaMemsetAsync(h->mom_part, 0, (size_t)2 * h->sm_count * m * 8, st));
    I8P t;
    memset(&t, 0, sizeof(t));
    t.M = m; t.N = (int)n_local; t.K = Mp; t.lower_a = 1; t.splits = 1; t.snake = 1; t.n_major = 1;
    t.ea = h->eL; t.e_dev = h->eKA; t.eb0_sel = 1; t.alpha = 1.0;
    t.Oq = h->atq_all; t.o_ld = nc; t.o_plane = (int64_t)Mp * nc; t.eo_sel = 2;
    t.o_chunk = nc; t.o_chunk_stride = (int64_t)I8_NS * Mp * nc;
    t.yv = y; t.rowdot = h->mom_part; t.rowdot_reg = 1;
    ProfScope ps(h, st, CAT_TRMM);
    RUN(launch_i8(h, st, I8_EPI_SLICE, t, {h->Lq, Mp, Mp, (int64_t)Mp * Mp}, {h->kq_all, n_local, Mp, h->kc_rows * Mp}));
    h->atq_valid = true;
  }
  // ... and the SYRK of the whole pass is one launch too: CTA = (tile, chunk group), the exact int32 sums of each 16384-row chunk are
  // drained into FP64 registers and the tile total is stored once (no read-modify-write, no per-chunk start-up)
  int syrk_tiles = 0;
  for (int tm = 0; tm < i8_tiles_m; ++tm) syrk_tiles += std::max(0, (m + I8_BN - 1) / I8_BN - 2 * tm);
  const bool syrk_once = trmm_once && syrk_tiles <= h->sm_count && !getenv("GGP_I8_NO_SYRK_ONCE");
  if (syrk_once) {
    const int nchunk = (int)((n_local + nc - 1) / nc);
    I8P sy;
    memset(&sy, 0, sizeof(sy));
    sy.M = m; sy.N = m; sy.K = (int)std::min<int64_t>(nc, n_local); sy.sym = 1; sy.splits = 1;
    sy.nchunk = nchunk; sy.k_last = (int)(n_local - (int64_t)(nchunk - 1) * nc); sy.nchunk_groups_max = splits;
    sy.e_dev = h->eKA; sy.ea0_sel = 2; sy.eb0_sel = 2; sy.alpha = 1.0; sy.beta = 0.0;
    sy.C = h->Spart; sy.ldc = Mp; sy.sSplit = sM;
    const I8Operand A{h->atq_all, m, nc, (int64_t)Mp * nc};
    ProfScope ps(h, st, CAT_SYRK);
    RUN(launch_i8(h, st, I8_EPI_F64, sy, A, A));
  }
  for (int64_t c0 = 0; c0 < n_local && !(trmm_once && syrk_once); c0 += nc) {
    const int nv = (int)std::min<int64_t>(nc, n_local - c0);
    double* Kc_c = h->kc_all ? h->kc_all + c0 * Mp : h->Kc;
    const int64_t sK = h->kc_all ? h->kc_rows * Mp : (int64_t)nc * Mp;
    if (i8) {
      int8_t* Kq_c = h->kq_all ? h->kq_all + c0 * Mp : h->Kq;
      const int64_t plK = h->kq_all ? h->kc_rows * Mp : (int64_t)nc * Mp;
      if (!prefetched && !trmm_once) {   // k(X,Z) tile and its digit planes
        ProfScope ps(h, st, CAT_BUILD);
        RUN(build_chunk_i8(h, st, X + c0 * d, nv, d, Z, m, theta, kind, Kc_c, Kq_c, plK));
      }
      if (!trmm_once) {   // A^T digits [m x nv] = L^{-1} (lower) x Kc^T, fused b-partials = A y (accumulated into one slab set over the chunks)
        I8P t;
        memset(&t, 0, sizeof(t));
        t.M = m; t.N = nv; t.K = Mp; t.lower_a = 1; t.splits = 1; t.snake = 1;
        { const char* e = getenv("GGP_I8_TRMM_ORDER"); t.n_major = (e && e[0] == '0') ? 0 : 1; }
        t.ea = h->eL; t.e_dev = h->eKA; t.eb0_sel = 1; t.alpha = 1.0;
        t.Oq = h->Atq; t.o_ld = nc; t.o_plane = (int64_t)Mp * nc; t.eo_sel = 2;
        t.yv = y + c0; t.rowdot = h->mom_part; t.rowdot_acc = 1;
        if (getenv("GGP_I8_TRMM_SERIAL")) t.serial_epi = 1;
        ProfScope ps(h, st, CAT_TRMM);
        RUN(launch_i8(h, st, I8_EPI_SLICE, t, {h->Lq, Mp, Mp, (int64_t)Mp * Mp}, {Kq_c, nv, Mp, plK}));
      }
      if (!syrk_once) {   // S_split += A A^T (tiles touching the upper triangle)
        I8P sy;
        memset(&sy, 0, sizeof(sy));
        sy.M = m; sy.N = m; sy.K = nv; sy.sym = 1;
        int tiles = 0;
        const int tmn = (m + I8_BM - 1) / I8_BM, tnn = (m + I8_BN - 1) / I8_BN;
        for (int tm = 0; tm < tmn; ++tm) tiles += std::max(0, tnn - 2 * tm);
        sy.splits = std::max(1, std::min(std::min(splits, h->sm_count / std::max(1, tiles)), (nv + 4 * I8_BKB - 1) / (4 * I8_BKB)));
        sy.e_dev = h->eKA; sy.ea0_sel = 2; sy.eb0_sel = 2; sy.alpha = 1.0; sy.beta = 1.0;
        sy.C = h->Spart; sy.ldc = Mp; sy.sSplit = sM;
        const I8Operand A = trmm_once ? I8Operand{h->atq_all + (c0 / nc) * (int64_t)I8_NS * Mp * nc, m, nc, (int64_t)Mp * nc}
                                      : I8Operand{h->Atq, m, nc, (int64_t)Mp * nc};
        ProfScope ps(h, st, CAT_SYRK);
        RUN(launch_i8(h, st, I8_EPI_F64, sy, A, A));
      }
      continue;
    }
    if (!prefetched) { ProfScope ps(h, st, CAT_BUILD); RUN(build_chunk(h, st, X + c0 * d, nv, d, Z, m, theta, kind, batch, Kc_c, sK)); }
    // At[m x nv] = Linv[m x m] * Kc[nv x m]^T   (k clipped to the lower triangle)
    GemmP t = gemm_basic(h->Linv, Mp, sM, Kc_c, Mp, sK, h->At, nc, (int64_t)nc * Mp, m, nv, m, 1.0, 0.0,
                         KM_A_LOWER);
    t.heavy_first = 1;
    t.yv = y + c0; t.rowdot = h->mom_part; t.sRowdot = (int64_t)(nc / BN) * m;   // b partials reuse the moment-partial buffer
    { ProfScope ps(h, st, CAT_TRMM); RUN(launch_gemm(h, st, EPI_STORE, t, batch)); }
    // S_split += At * At^T  (upper tiles)
    GemmP s = gemm_basic(h->At, nc, (int64_t)nc * Mp, h->At, nc, (int64_t)nc * Mp, h->Spart, Mp, (int64_t)splits * sM, m, m, nv,
                         1.0, 1.0);
    s.sym = 1; s.splits = splits; s.sSplit = sM;
    { ProfScope ps(h, st, CAT_SYRK); RUN(launch_gemm(h, st, EPI_STORE, s, batch)); }
    ProfScope ps_o(h, st, CAT_OTHER);
    // b += sum over n-tiles of the fused row dots (fixed order)
    k_reduce_moments<<<dim3((unsigned)((m + 31) / 32), batch), 256, 0, st>>>(h->mom_part, m, (int64_t)(nc / BN) * m,
                                                                              (nv + BN - 1) / BN, m, h->bvec);
    CKL();
  }
  if (i8 && n_local > 0) {   // b = sum of the row-dot slabs (fixed order)
    ProfScope ps(h, st, CAT_OTHER);
    const int nslab = trmm_once ? 2 * h->sm_count : 2 * (int)((std::min<int64_t>(nc, n_local) + I8_BN - 1) / I8_BN);
    k_reduce_moments<<<dim3((unsigned)((m + 31) / 32), 1), 256, 0, st>>>(h->mom_part, m, 0, nslab, m, h->bvec);
    CKL();
  }
  if (h->kc_all) {
    h->kc_valid = true;
    h->kc_X = X; h->kc_Z = Z; h->kc_theta = theta; h->kc_n = n_local; h->kc_batch = batch; h->kc_kind = kind;
  }
  k_finalize_partial<<<dim3((m + 15) / 16, (m + 15) / 16, batch), dim3(16, 16), 0, st>>>(
      h->Spart, Mp, sM, (int64_t)splits * sM, splits, h->bvec, m, h->yty, n_local, theta, d, m, partial,
      (int64_t)m * m + m + 3);
  CKL();
  return 0;
}

int ggp_sgpr_predict_pass1(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X, const double* y, int64_t n_local,
                           const double* Z, const double* theta, int m, int d, int batch, double* partial) {
  if (!h || !Z || !theta || !partial || (n_local > 0 && (!X || !y))) return fail(-1, "ggp_sgpr_predict_pass1: NULL argument");
  if (!reserved_for(h, n_local, m, d, batch)) return fail(-2, "ggp_sgpr_predict_pass1: handle not reserved for this shape");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_sgpr_predict_pass1: unknown kernel");
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = h->Mp, nc = h->nc, splits = h->splits;
  const int64_t sM = (int64_t)Mp * Mp, sC = (int64_t)nc * Mp;
  h->kc_valid = false;   // the chunk buffers are reused below
  h->pf_valid = false;
  h->atq_valid = false;
  RUN(join_aux(h, st));
  CK(cudaMemsetAsync(h->Spart, 0, (size_t)batch * splits * sM * 8, st));
  CK(cudaMemsetAsync(h->bvec, 0, (size_t)batch * Mp * 8, st));
  CK(cudaMemsetAsync(h->yty, 0, 256, st));
  for (int64_t c0 = 0; c0 < n_local; c0 += nc) {
    const int nv = (int)std::min<int64_t>(nc, n_local - c0);
    RUN(build_chunk(h, st, X + c0 * d, nv, d, Z, m, theta, kind, batch, h->Kc, sC));
    GemmP t = gemm_basic(h->Linv, Mp, sM, h->Kc, Mp, sC, h->At, nc, sC, m, nv, m, 1.0, 0.0, KM_A_LOWER);
    t.heavy_first = 1;
    RUN(launch_gemm(h, st, EPI_STORE, t, batch));
    k_fitc_scale<<<dim3((nv + 255) / 256, batch), 256, 0, st>>>(h->At, nc, sC, m, nv, y + c0, theta, d, h->ysc, nc);
    CKL();
    GemmP s = gemm_basic(h->At, nc, sC, h->At, nc, sC, h->Spart, Mp, (int64_t)splits * sM, m, m, nv, 1.0, 1.0);
    s.sym = 1; s.splits = splits; s.sSplit = sM;
    RUN(launch_gemm(h, st, EPI_STORE, s, batch));
    k_gemv_acc<<<dim3((m + 7) / 8, batch), 256, 0, st>>>(h->At, nc, sC, h->ysc, nc, h->bvec, Mp, m, nv);
    CKL();
  }
  k_finalize_partial<<<dim3((m + 15) / 16, (m + 15) / 16, batch), dim3(16, 16), 0, st>>>(
      h->Spart, Mp, sM, (int64_t)splits * sM, splits, h->bvec, Mp, h->yty, n_local, theta, d, m, partial, (int64_t)m * m + m + 3);
  CKL();
  return 0;
}

int ggp_sgpr_finish(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* Z, const double* theta, int m, int d,
                    int batch, const double* partial, int need_grad, double* bound, double* grad_mm, int32_t* info) {
  if (!h || !Z || !theta || !partial || !bound || !info) return fail(-1, "ggp_sgpr_finish: NULL argument");
  if (need_grad && !grad_mm) return fail(-1, "ggp_sgpr_finish: grad_mm is NULL");
  if (!reserved_for(h, 0, m, d, batch)) return fail(-2, "ggp_sgpr_finish: handle not reserved for this shape");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_sgpr_finish: unknown kernel");
  if (kind == GGP_KERNEL_COMPOSITE && need_grad) {
    if (int r = prog_ready(h, d, "ggp_sgpr_finish")) return r;
    if (!h->kgrad_mm) return fail(-3, "ggp_sgpr_finish: composite kernel gradient needs kgrad_mm (ggp_set_kernel_params)");
    if ((size_t)batch * m * (h->prog.P + d) * 8 > h->krow_bytes) return fail(-2, "ggp_sgpr_finish: call ggp_set_kernel_program after ggp_reserve");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = h->Mp;
  const int64_t sM = (int64_t)Mp * Mp, sP = (int64_t)m * m + m + 3, sG = (int64_t)d + 2 + (int64_t)m * d;
  const dim3 g16(Mp / 16, Mp / 16, batch), b16(16, 16);
  const dim3 gv((m + 7) / 8, batch);
  RUN(join_aux(h, st));
  ProfScope ps(h, st, CAT_MM);
  k_make_B<<<g16, b16, 0, st>>>(partial, sP, m, Mp, theta, d, h->Bm, sM);
  CKL();
  CK(cudaMemsetAsync(h->piv_tol, 0, sizeof(double) * batch, st));   // B = I + A A^T / s: LAPACK semantics
  RUN(chol_and_inverse(h, st, h->Bm, h->LBinv, h->LBinvT, batch, info));
  // Binv = LBinv^T LBinv.  (The m x m products stay on the FP64 DMMA kernel: routed through the sliced-integer GEMM they were 0.36 ms
  // faster at M = 1024, but its row-scaled fixed point resolves an element only relative to its ROW maximum, and the operands
  // here -- L^{-T}, P_A -- span many orders of magnitude within a row: 8e-8 on the ell-gradient at N = 1777, M = 100, D = 1.)
  RUN(launch_gemm(h, st, EPI_STORE,
                  gemm_basic(h->LBinvT, Mp, sM, h->LBinvT, Mp, sM, h->Binv, Mp, sM, m, m, m, 1.0, 0.0, KM_A_UPPER | KM_B_UPPER),
                  batch));
  const double* bsrc = partial + (int64_t)m * m;
  // c = LBinv b / s ; beta = Binv b ; u = Linv^T beta / s^2
  auto bound_part = [&](cudaStream_t s1) -> int {   // c and the scalars of the bound / dF/ds2: nothing pass 2 waits for
    k_gemv<<<gv, 256, 0, s1>>>(h->LBinv, Mp, sM, bsrc, sP, h->cvec, Mp, m, m, 1.0, theta, d, 1);
    CKL();
    k_bound_scalars<<<batch, 256, 0, s1>>>(partial, sP, m, Mp, theta, d, h->Bm, h->Binv, sM, h->cvec, h->beta, Mp, bound,
                                           need_grad ? h->ds2 : nullptr);
    CKL();
    return 0;
  };
  k_gemv<<<gv, 256, 0, st>>>(h->Binv, Mp, sM, bsrc, sP, h->beta, Mp, m, m, 1.0, theta, d, 0);
  CKL();
  if (!need_grad) return bound_part(st);
  k_gemv<<<gv, 256, 0, st>>>(h->LinvT, Mp, sM, h->beta, Mp, h->u, Mp, m, m, 1.0, theta, d, 2);
  CKL();
  k_make_PA_Gbar<<<g16, b16, 0, st>>>(partial, sP, m, Mp, theta, d, h->Binv, h->beta, Mp, h->PA, h->Gbar, sM);
  CKL();
  // Q = Linv^T PA (kept in h->P; pass 2 forms dF/dKzx = Q A + u y^T from A = L^{-1} Kzx) ;  Gzz = -1/2 Linv^T Gbar Linv
  // (Not P = Linv^T PA Linv applied to Kzx, SURVEY R5 as written: P has entries of size 1 / lambda_min(Kzz) and P Kzx cancels down by
  // cond(Kzz) -- 1e-8 .. 2e-6 of the gradient at the headline Kzz in float64 against an extended-precision evaluation, whereas Q A,
  // which is what autograd through the triangular solve computes, loses cond(L) eps: tests/test_oracle_hp.py)
  // Only beta, u, P_A and Q are on the path to pass 2.  Everything else -- c and the scalars of the bound, rk, the Gzz chain (scratch Wk,
  // free outside the triangular inverse) and what hangs off it -- goes to the auxiliary stream: see join_aux.
  const bool side = !getenv("GGP_MM_ONE_STREAM");
  cudaStream_t st2 = st;
  if (side) {
    if (!h->aux_stream) {
      CK(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    st2 = h->aux_stream;
    CK(cudaEventRecord(h->ev_fork, st));
    CK(cudaStreamWaitEvent(st2, h->ev_fork, 0));
  }
  RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->LinvT, Mp, sM, h->PA, Mp, sM, h->P, Mp, sM, m, m, m, 1.0, 0.0, KM_A_UPPER), batch));
  RUN(bound_part(st2));
  // sum(G o Kzx) = tr(P_A S) + beta^T b / s^2 from the m x m quantities: dF/dsf2 of every kernel kind (see k_grad_from_moments)
  k_rk_partial<<<dim3(RK_BLOCKS, batch), 256, 0, st2>>>(partial, sP, m, Mp, h->PA, sM, h->rk_part);
  CKL();
  k_rk_from_mm<<<batch, 256, 0, st2>>>(partial, sP, m, Mp, theta, d, h->rk_part, h->beta, h->rk);
  CKL();
  double* T2 = side ? h->Wk : h->T1;
  RUN(launch_gemm(h, st2, EPI_STORE, gemm_basic(h->LinvT, Mp, sM, h->Gbar, Mp, sM, T2, Mp, sM, m, m, m, 1.0, 0.0, KM_A_UPPER), batch));
  RUN(launch_gemm(h, st2, EPI_STORE, gemm_basic(T2, Mp, sM, h->LinvT, Mp, sM, h->Gzz, Mp, sM, m, m, m, -0.5, 0.0, KM_B_UPPER), batch));
  if (kind == GGP_KERNEL_COMPOSITE) {
    // sum_ij Gzz_ij dk(z_i, z_j)/d(parameter), and dZ_i = 2 sum_j Gzz_ij dk(z_j, z_i)/d(second argument): z_i moves both as the second
    // argument of row i and, Gzz and k being symmetric, as the first argument of column i.  The amplitudes also get the explicit
    // -N / (2 s2) of sum_n k(x_n, x_n).  krow is shared with pass 2 (which runs on `st` after the join), so the chain stays ordered.
    k_kprog_grad<<<gv, 256, 0, st2>>>(h->Gzz, Mp, sM, nullptr, 0, nullptr, Z, m, Z, m, h->prog, h->kth, h->krow, 0);
    CKL();
    k_kprog_grad_final<<<batch, 256, 0, st2>>>(h->krow, m, h->prog, 2.0, partial + (int64_t)m * m + m + 2, sP, theta, h->ds2,
                                               h->kgrad_mm, grad_mm, sG);
    CKL();
  } else {
    k_grad_kzz_rows<<<gv, 256, 0, st2>>>(h->Gzz, Mp, sM, Z, m, d, theta, kind, h->rowacc, grad_mm + d + 2, sG);
    CKL();
    k_grad_mm_final<<<batch, 256, 0, st2>>>(h->rowacc, m, d, theta, partial, sP, h->ds2, grad_mm, sG, h->rk);
    CKL();
  }
  if (side) {
    CK(cudaEventRecord(h->ev_join, st2));
    h->aux_pending = true;
  }
  return 0;
}

int ggp_sgpr_join(ggp_handle_t* h, void* stream) {
  if (!h) return fail(-1, "ggp_sgpr_join: handle is NULL");
  return join_aux(h, (cudaStream_t)stream);
}

// Composite kernel, pass 2: per chunk  aT = Kc Linv^T,  G = Q aT^T STORED (m x nv, into the chunk buffer the tile no longer needs),
// then the direct contraction  sum_n (G + u y^T)_mn dk(x_n, z_m)/d(parameter, z_m)  (k_kprog_grad) instead of the moment epilogue.
static int pass2_composite(ggp_handle* h, cudaStream_t st, const double* X, const double* y, int64_t n_local, const double* Z,
                           const double* theta, int m, int d, int batch, double* grad_partial) {
  if (int r = prog_ready(h, d, "ggp_sgpr_pass2")) return r;
  if (!h->kgrad_partial) return fail(-3, "ggp_sgpr_pass2: composite kernel gradient needs kgrad_partial (ggp_set_kernel_params)");
  if ((size_t)batch * m * (h->prog.P + d) * 8 > h->krow_bytes) return fail(-2, "ggp_sgpr_pass2: call ggp_set_kernel_program after ggp_reserve");
  const KSpec kind{GGP_KERNEL_COMPOSITE, 0.0};
  const int Mp = h->Mp, nc = h->nc;
  const int64_t sM = (int64_t)Mp * Mp, sG = (int64_t)d + 2 + (int64_t)m * d, sC = (int64_t)nc * Mp;
  const bool cached = h->kc_all && h->kc_valid && h->kc_X == X && h->kc_Z == Z && h->kc_theta == theta && h->kc_n == n_local &&
                      h->kc_batch == batch && h->kc_kind.kind == kind.kind;
  RUN(join_aux(h, st));   // the Kzz-part chain of finish() uses krow on the auxiliary stream
  const dim3 gv((m + 7) / 8, batch);
  CK(cudaMemsetAsync(h->krow, 0, (size_t)batch * m * (h->prog.P + d) * 8, st));
  for (int64_t c0 = 0; c0 < n_local; c0 += nc) {
    const int nv = (int)std::min<int64_t>(nc, n_local - c0);
    double* Kc_c = h->kc_all ? h->kc_all + c0 * Mp : h->Kc;
    const int64_t sK = h->kc_all ? h->kc_rows * Mp : sC;
    if (!cached) { ProfScope ps(h, st, CAT_BUILD); RUN(build_chunk(h, st, X + c0 * d, nv, d, Z, m, theta, kind, batch, Kc_c, sK)); }
    {
      ProfScope ps(h, st, CAT_TRMM);
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(Kc_c, Mp, sK, h->Linv, Mp, sM, h->At, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_LOWER), batch));
    }
    {
      ProfScope ps(h, st, CAT_BWD);
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->P, Mp, sM, h->At, Mp, sC, h->Kc, nc, sC, m, nv, m, 1.0, 0.0), batch));
      k_kprog_grad<<<gv, 256, 0, st>>>(h->Kc, nc, sC, h->u, Mp, y + c0, X + c0 * d, nv, Z, m, h->prog, h->kth, h->krow, 1);
      CKL();
    }
  }
  k_kprog_grad_final<<<batch, 256, 0, st>>>(h->krow, m, h->prog, 1.0, nullptr, 0, theta, nullptr, h->kgrad_partial, grad_partial, sG);
  CKL();
  return 0;
}

int ggp_sgpr_pass2(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X, const double* y, int64_t n_local,
                   const double* Z, const double* theta, int m, int d, int batch, double* grad_partial) {
  if (!h || !Z || !theta || !grad_partial || (n_local > 0 && (!X || !y))) return fail(-1, "ggp_sgpr_pass2: NULL argument");
  if (!reserved_for(h, n_local, m, d, batch)) return fail(-2, "ggp_sgpr_pass2: handle not reserved for this shape");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_sgpr_pass2: unknown kernel");
  if (kind == GGP_KERNEL_COMPOSITE) return pass2_composite(h, (cudaStream_t)stream, X, y, n_local, Z, theta, m, d, batch, grad_partial);
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = h->Mp, nc = h->nc, nq = 2 * d + 1;
  const int64_t sM = (int64_t)Mp * Mp, sG = (int64_t)d + 2 + (int64_t)m * d, sC = (int64_t)nc * Mp;
  const int64_t cnt = (int64_t)m * nq;
  CK(cudaMemsetAsync(h->mom_acc, 0, (size_t)batch * cnt * 8, st));
  // the tiles cached by the pass 1 of this evaluation (same operands, same handle, no factor() since) are reused as they are
  const bool cached = h->kc_all && h->kc_valid && h->kc_X == X && h->kc_Z == Z && h->kc_theta == theta && h->kc_n == n_local &&
                      h->kc_batch == batch && h->kc_kind.kind == kind.kind && h->kc_kind.p == kind.p;
  const bool i8 = use_i8(h, cfg, d, batch);
  // dF/dKzx = Q A + u y^T with Q = L^{-T} P_A (h->P, from finish) and A = L^{-1} Kzx: the streamed operand of the backward product is
  // A^T -- on the sliced-integer path the digit planes the triangular multiply of pass 1 left in atq_all (read MN-major, no second
  // copy), otherwise rebuilt per chunk -- and k(X,Z) (or dk/d(d2)) only enters as the FP64 multiplier of the epilogue.
  const bool a_cached = i8 && cached && h->atq_valid && h->atq_all;
  if (i8) {
    ProfScope ps(h, st, CAT_BUILD);
    k_i8_exponents<<<1, 1, 0, st>>>(theta, d, h->eKA);
    CKL();
    k_slice_rows<<<(Mp + 7) / 8, 256, 0, st>>>(h->P, Mp, Mp, Mp, h->Pq, Mp, (int64_t)Mp * Mp, Mp, h->eP);
    CKL();
    if (!a_cached) {   // the per-chunk triangular multiply below needs the digits of L^{-1} (pass 1 may have run on another plan)
      k_slice_rows<<<(Mp + 7) / 8, 256, 0, st>>>(h->Linv, Mp, Mp, Mp, h->Lq, Mp, (int64_t)Mp * Mp, Mp, h->eL);
      CKL();
    }
  }
  // sliced-integer path: the moments stay in the registers of the CTA that produced them over all its tiles (2 d + 1 <= 24), and
  // with the cached tiles of pass 1 the whole local row range is ONE launch (no per-chunk tails, 2 x 18 slabs to reduce)
#ifdef GGP_I8_MOM_VEC
  const bool i8_accum = i8 && d <= 8 && (Mp / I8_BM) <= h->sm_count && !getenv("GGP_I8_NO_ACCUM");   // vector-pipe moments: 2 d register chains per thread
#else
  const bool i8_accum = i8 && nq <= 24 && (Mp / I8_BM) <= h->sm_count && !getenv("GGP_I8_NO_ACCUM");
#endif
  const int64_t step = (i8_accum && a_cached && kind == GGP_KERNEL_RBF && n_local < (int64_t)1 << 30) ? std::max<int64_t>(n_local, 1) : nc;
  for (int64_t c0 = 0; c0 < n_local; c0 += step) {
    const int nv = (int)std::min<int64_t>(step, n_local - c0);
    double* Kc_c = h->kc_all ? h->kc_all + c0 * Mp : h->Kc;
    const int64_t sK = h->kc_all ? h->kc_rows * Mp : sC;
    if (!cached && !i8) { ProfScope ps(h, st, CAT_BUILD); RUN(build_chunk(h, st, X + c0 * d, nv, d, Z, m, theta, kind, batch, Kc_c, sK)); }
    if (i8) {
      int8_t* Kq_c = h->kq_all ? h->kq_all + c0 * Mp : h->Kq;
      const int64_t plK = h->kq_all ? h->kc_rows * Mp : sC;
      if (!cached) {
        ProfScope ps(h, st, CAT_BUILD);
        RUN(build_chunk_i8(h, st, X + c0 * d, nv, d, Z, m, theta, kind, Kc_c, Kq_c, plK));
      }
      if (!a_cached) {   // A^T digits of this chunk: [plane][m][nc] = L^{-1} (lower) x Kc^T
        I8P t;
        memset(&t, 0, sizeof(t));
        t.M = m; t.N = nv; t.K = Mp; t.lower_a = 1; t.splits = 1; t.snake = 1; t.n_major = 1;
        t.ea = h->eL; t.e_dev = h->eKA; t.eb0_sel = 1; t.alpha = 1.0;
        t.Oq = h->Atq; t.o_ld = nc; t.o_plane = (int64_t)Mp * nc; t.eo_sel = 2;
        ProfScope ps(h, st, CAT_TRMM);
        RUN(launch_i8(h, st, I8_EPI_SLICE, t, {h->Lq, Mp, Mp, (int64_t)Mp * Mp}, {Kq_c, nv, Mp, plK}));
      }
      const double* Kmul = Kc_c;
      if (kind != GGP_KERNEL_RBF) {   // the multiplier is dk/d(d2), built into the FP64 A^T buffer (unused on this path)
        ProfScope ps(h, st, CAT_BUILD);
        dim3 grid(Mp / KT_M, (nv + KT_N - 1) / KT_N, batch);
        const size_t smem = (size_t)(KT_N * d + KT_M * d + d) * 8;
        k_build_kc<<<grid, KT_THREADS, smem, st>>>(X + c0 * d, nv, nv, d, Z, m, theta, kind, h->At, Mp, sC, 1);
        CKL();
        Kmul = h->At;
      }
      I8P g8;
      memset(&g8, 0, sizeof(g8));
      g8.M = m; g8.N = nv; g8.K = Mp; g8.n_major = 1; g8.splits = 1;
      g8.ea = h->eP; g8.e_dev = h->eKA; g8.eb0_sel = 2; g8.alpha = 1.0;
      g8.b_mn = 1; g8.b_chunk = a_cached ? nc : 0;
      g8.u = h->u; g8.yv = y + c0; g8.Kmul = Kmul; g8.ldk = Mp; g8.Xc = X + c0 * d; g8.d = d;
      g8.mom = h->mom_part; g8.sMomTile = cnt;
      g8.mom_accum = i8_accum ? 1 : 0;
      { const char* e = getenv("GGP_I8_MOM_FRAG"); g8.mom_frag_off = (e && e[0] == '0') ? 1 : 0; }
      { const char* e = getenv("GGP_I8_SERIAL_EPI"); g8.serial_epi = (e && e[0] == '0') ? 0 : 1; }
      // B = the A^T digit planes as the triangular multiply stores them: [plane][m (k)][columns], columns contiguous (MN-major);
      // chunk-blocked over all local rows (one [7][Mp][nc] array per chunk) or the single chunk buffer
      const int nchunks = a_cached ? (int)((n_local + nc - 1) / nc) : 1;
      const I8Operand Bop{a_cached ? h->atq_all + (c0 / nc) * (int64_t)I8_NS * Mp * nc : h->Atq, m, nc, (int64_t)Mp * nc};
      g8.b_planes = I8_NS * (a_cached ? nchunks - (int)(c0 / nc) : 1);
      { ProfScope ps(h, st, CAT_BWD); RUN(launch_i8(h, st, I8_EPI_MOMENTS, g8, {h->Pq, Mp, Mp, (int64_t)Mp * Mp}, Bop)); }
      ProfScope ps_o(h, st, CAT_OTHER);
      // slabs the launch wrote: 2 per CTA column group on the register-resident plan, else one per 32 columns
      const int bn2 = i8_tile_width(I8_EPI_MOMENTS, g8);
      const int tiles_m = (m + I8_BM - 1) / I8_BM, tiles_n = (nv + bn2 - 1) / bn2;
      const int nslabs = i8_accum ? 2 * std::max(1, std::min(h->sm_count / tiles_m, tiles_n)) : (bn2 == 64 ? 2 * tiles_n : tiles_n);
      k_reduce_moments<<<dim3((unsigned)((cnt + 31) / 32), batch), 256, 0, st>>>(h->mom_part, cnt, 0, nslabs, cnt, h->mom_acc);
      CKL();
      continue;
    }
    const int ntiles = (nv + BN - 1) / BN;
    // aT[nv x m] = Kc[nv x m] * Linv^T  (rows of A^T of this chunk; k clipped to the triangle)
    {
      ProfScope ps(h, st, CAT_TRMM);
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(Kc_c, Mp, sK, h->Linv, Mp, sM, h->At, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_LOWER), batch));
    }
    GemmP g = gemm_basic(h->P, Mp, sM, h->At, Mp, sC, nullptr, 0, 0, m, nv, m, 1.0, 0.0);
    g.n_major = 1;   // all row tiles of one chunk-row tile back to back: each A^T tile comes from HBM once, then from L2
    g.u = h->u; g.su = Mp;
    g.yv = y + c0;
    g.Kc = Kc_c; g.ldk = Mp; g.sK = sK;
    if (kind != GGP_KERNEL_RBF) {
      // Matern: the epilogue multiplier is dk/d(d2), not k (RBF: -k/2, folded into k_grad_from_moments); built into the chunk buffer
      // (free when the k(X,Z) tiles come from the cache; otherwise it overwrites this chunk's k(X,Z), already consumed above)
      ProfScope ps(h, st, CAT_BUILD);
      dim3 grid(Mp / KT_M, (nv + KT_N - 1) / KT_N, batch);
      const size_t smem = (size_t)(KT_N * d + KT_M * d + d) * 8;
      k_build_kc<<<grid, KT_THREADS, smem, st>>>(X + c0 * d, nv, nv, d, Z, m, theta, kind, h->Kc, Mp, sC, 1);
      CKL();
      g.Kc = h->Kc; g.ldk = Mp; g.sK = sC;
    }
    g.Xc = X + c0 * d; g.d = d;
    g.mom = h->mom_part; g.sMomTile = cnt; g.sMom = (int64_t)(nc / 32) * cnt;
    { ProfScope ps(h, st, CAT_BWD); RUN(launch_gemm(h, st, EPI_MOMENTS, g, batch)); }
    ProfScope ps_o(h, st, CAT_OTHER);
    k_reduce_moments<<<dim3((unsigned)((cnt + 31) / 32), batch), 256, 0, st>>>(h->mom_part, cnt, (int64_t)(nc / 32) * cnt,
                                                                                ntiles * WARPS_N, cnt, h->mom_acc);
    CKL();
  }
  k_grad_from_moments<<<batch, 256, 0, st>>>(h->mom_acc, m, d, Z, theta, grad_partial, sG, h->rk, kind != GGP_KERNEL_RBF ? 1 : 0);
  CKL();
  return join_aux(h, st);   // grad_mm of the finish() of this evaluation is complete from here on
}

int ggp_sgpr_predict(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* Xs, int64_t ns, const double* Z,
                     const double* theta, int m, int d, int batch, int add_noise, double* mean, double* var, double* cov) {
  if (!h || !Xs || !Z || !theta || !mean || !var) return fail(-1, "ggp_sgpr_predict: NULL argument");
  if (!reserved_for(h, 0, m, d, batch)) return fail(-2, "ggp_sgpr_predict: handle not reserved for this shape");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_sgpr_predict: unknown kernel");
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = h->Mp, nc = h->nc;
  const int64_t sM = (int64_t)Mp * Mp, sC = (int64_t)nc * Mp;
  h->kc_valid = false;   // the chunk buffers are scratch here
  h->atq_valid = false;
  RUN(join_aux(h, st));
  // t^T[rows x m] = (k(X*, Z) Linv^T) LBinv^T of the test rows [r0, r0 + nv) into At + off (a^T) and Kc + off (t^T)
  auto rows_t = [&](int64_t r0, int nv, int64_t off) -> int {
    RUN(build_chunk(h, st, Xs + r0 * d, nv, d, Z, m, theta, kind, batch, h->Kc + off, sC));
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->Kc + off, Mp, sC, h->Linv, Mp, sM, h->At + off, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_LOWER), batch));
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->At + off, Mp, sC, h->LBinv, Mp, sM, h->Kc + off, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_LOWER), batch));
    return 0;
  };
  // a full covariance over more test rows than one chunk is tiled over (row chunk, column chunk) pairs of half-chunk size, so that the
  // t^T rows of both chunks are resident in the two halves of the chunk buffers
  const bool tiled = cov && ns > nc;
  const int step = tiled ? std::max(64, nc / 2 / 64 * 64) : nc;
  const int64_t half = (int64_t)step * Mp;
  for (int64_t c0 = 0; c0 < ns; c0 += step) {
    const int nv = (int)std::min<int64_t>(step, ns - c0);
    RUN(rows_t(c0, nv, 0));
    k_predict_rows<<<dim3((nv + 7) / 8, batch), 256, 0, st>>>(h->At, h->Kc, Mp, sC, h->cvec, Mp, theta, d, m, nv, add_noise,
                                                              mean + c0, var + c0, ns);
    CKL();
    if (!cov) continue;
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->Kc, Mp, sC, h->Kc, Mp, sC, cov + c0 * ns + c0, ns, ns * ns, nv, nv, m, 1.0, 0.0), batch));
    for (int64_t c1 = c0 + step; tiled && c1 < ns; c1 += step) {
      const int nw = (int)std::min<int64_t>(step, ns - c1);
      RUN(rows_t(c1, nw, half));
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->Kc, Mp, sC, h->Kc + half, Mp, sC, cov + c0 * ns + c1, ns, ns * ns, nv, nw, m, 1.0, 0.0), batch));
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->Kc + half, Mp, sC, h->Kc, Mp, sC, cov + c1 * ns + c0, ns, ns * ns, nw, nv, m, 1.0, 0.0), batch));
    }
  }
  if (cov) {
    k_cov_diag<<<dim3((unsigned)((ns + 255) / 256), batch), 256, 0, st>>>(cov, ns, var);
    CKL();
  }
  return 0;
}

int ggp_svgp_elbo(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* xb, const double* yb, int64_t nb,
                  const double* Z, const double* qm, int qm_batched, const double* qLs, const double* theta, const double* jitter,
                  int m, int d, int batch, double data_jitter, double lik_scale, double kl_scale, int likelihood, int need_grad,
                  double* elbo, double* grad, int32_t* info) {
  const int64_t sqm = qm_batched ? m : 0;
  if (!h || !xb || !yb || !Z || !qm || !theta || !jitter || !elbo || !info) return fail(-1, "ggp_svgp_elbo: NULL argument");
  if (need_grad && !grad) return fail(-1, "ggp_svgp_elbo: grad is NULL");
  if (!reserved_for(h, 0, m, d, batch)) return fail(-2, "ggp_svgp_elbo: handle not reserved for this shape");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (kind.kind < GGP_KERNEL_RBF || kind.kind > GGP_KERNEL_RQ || (kind.kind == GGP_KERNEL_RQ && !(kind.p > 0.0))) return fail(-3, "ggp_svgp_elbo: unknown kernel");
  const int nq = 2 * d + 1;
  if (nq > h->Mp) return fail(-3, "ggp_svgp_elbo: 2d+1 must not exceed the padded inducing count");
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = h->Mp, nsv = h->nsv;
  const int64_t sM = (int64_t)Mp * Mp, sC = (int64_t)nsv * Mp, sG = (int64_t)d + 2 + (int64_t)m * d + m + (int64_t)m * m;
  const dim3 g16(Mp / 16, Mp / 16, batch), b16(16, 16), g16one(Mp / 16, Mp / 16, 1);
  const bool hasS = qLs != nullptr;
  RUN(join_aux(h, st));
  double *Kc = h->sv[0], *aT = h->sv[1], *wT = h->sv[2], *SL = h->sv[3], *tA = h->sv[4], *tB = h->Kc, *tC = h->At;
  double *LsP = h->Bm, *LsT = h->LBinv, *dLsraw = h->Binv, *Gb = h->PA, *Hm = h->Gbar, *dKzz = h->Gzz, *gk = h->P,
         *dZzz = h->LBinvT, *dm = h->bvec, *scal = h->cvec, *mom = h->mom_acc;
  // note: h->Kc / h->At hold at least nsv x Mp per batch element (nc >= nsv) and are used here as transposed scratch
  {
    ProfScope ps(h, st, CAT_MM);
    k_build_kzz<<<g16, b16, 0, st>>>(Z, m, Mp, d, theta, jitter, kind, h->L, sM, h->piv_tol);
    CKL();
    RUN(chol_and_inverse(h, st, h->L, h->Linv, h->LinvT, batch, info));
  }
  ProfScope ps(h, st, CAT_OTHER);
  if (hasS) {
    k_pad_tril<<<g16one, b16, 0, st>>>(qLs, m, LsP, Mp);
    CKL();
    k_transpose<<<dim3(Mp / 32, Mp / 32, 1), dim3(32, 8), 0, st>>>(LsP, LsT, Mp, sM);
    CKL();
  }
  CK(cudaMemsetAsync(scal, 0, (size_t)batch * 4 * 8, st));
  if (need_grad) {
    CK(cudaMemsetAsync(dm, 0, (size_t)batch * Mp * 8, st));
    CK(cudaMemsetAsync(dLsraw, 0, (size_t)batch * sM * 8, st));
    CK(cudaMemsetAsync(Gb, 0, (size_t)batch * sM * 8, st));
    CK(cudaMemsetAsync(mom, 0, (size_t)batch * m * nq * 8, st));
  }
  for (int64_t c0 = 0; c0 < nb; c0 += nsv) {
    const int nv = (int)std::min<int64_t>(nsv, nb - c0);
    const int nvp = (nv + 15) / 16 * 16;  // padded row count of the transposed buffers (zero filled)
    const double* Xc = xb + c0 * d;
    {
      dim3 grid(Mp / KT_M, (nv + KT_N - 1) / KT_N, batch);
      const size_t smem = (size_t)(KT_N * d + KT_M * d + d) * 8;
      k_build_kc<<<grid, KT_THREADS, smem, st>>>(Xc, nv, nv, d, Z, m, theta, kind, Kc, Mp, sC);
      CKL();
    }
    // aT[nv x m] = Kc * Linv^T ; wT = aT * Ls
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(Kc, Mp, sC, h->Linv, Mp, sM, aT, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_LOWER), batch));
    if (hasS)
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(aT, Mp, sC, LsT, Mp, 0, wT, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_UPPER), batch));
    k_svgp_rows<<<dim3((nv + 7) / 8, batch), 256, 0, st>>>(aT, hasS ? wT : nullptr, Mp, sC, qm, sqm, yb + c0, theta, d, m, nv, likelihood,
                                                          data_jitter, lik_scale, h->rowout, nsv);
    CKL();
    k_svgp_reduce_rows<<<batch, 256, 0, st>>>(h->rowout, nsv, nv, scal);
    CKL();
    if (!need_grad) continue;
    // SL = wT * Ls^T = (S a)^T ;  GAT = gmu m^T + 2 gv (SL - aT)
    if (hasS)
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(wT, Mp, sC, LsP, Mp, 0, SL, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_LOWER), batch));
    k_svgp_dm<<<dim3((m + 31) / 32, batch), 256, 0, st>>>(aT, Mp, sC, h->rowout, nsv, m, nv, dm, Mp);
    CKL();
    const dim3 gT((nvp + 31) / 32, Mp / 32, batch), bT(32, 8);
    if (hasS) {  // dLsraw += (aT o gv)^T wT   (tB / tC are scratch here; the fused kernel below rewrites tB)
      k_transpose_rect<<<gT, bT, 0, st>>>(aT, nullptr, Mp, sC, h->rowout + 2 * nsv, (int64_t)4 * nsv, nv, Mp, tB, nvp, sC);
      CKL();
      k_transpose_rect<<<gT, bT, 0, st>>>(wT, nullptr, Mp, sC, nullptr, 0, nv, Mp, tC, nvp, sC);
      CKL();
      {   // only tril(dLsraw) enters the gradient (k_svgp_final): lower tiles
        GemmP gl = gemm_basic(tB, nvp, sC, tC, nvp, sC, dLsraw, Mp, sM, m, m, nv, 1.0, 1.0);
        gl.sym = 2;
        RUN(launch_gemm(h, st, EPI_STORE, gl, batch));
      }
    }
    // GAT = gmu m^T + 2 gv (SL - aT) in place in SL, and the k-contiguous copies tA = aT^T, tB = GAT^T for  Gbar += GAT^T aT
    k_svgp_gat_t<<<gT, bT, 0, st>>>(SL, aT, Mp, sC, qm, sqm, h->rowout, nsv, m, Mp, nv, hasS ? 1 : 0, tA, tB, nvp);
    CKL();
    {   // only the lower triangle of Gbar is read (k_sym_phi: H = sym(Phi(Gbar))): lower tiles, 10 of 16 at M = 512
      GemmP gg = gemm_basic(tB, nvp, sC, tA, nvp, sC, Gb, Mp, sM, m, m, nv, 1.0, 1.0);
      gg.sym = 2;
      RUN(launch_gemm(h, st, EPI_STORE, gg, batch));
    }
    // dKc = GAT * Linv  (into wT's buffer) ;  mom += (dKc o Kc)^T [1, x, x^2]
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(SL, Mp, sC, h->LinvT, Mp, sM, wT, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_UPPER), batch));
    if (kind != GGP_KERNEL_RBF) {
      // Matern / RQ: the moments are weighted by dk/d(d2) instead of k, so (1) the k-weighted total sum(dKc o Kc) of dF/dsf2 is taken
      // here, (2) the k(X,Z) tile (consumed above) is overwritten by the dk/d(d2) tile
      k_sum_prod_partial<<<dim3(RK_BLOCKS, batch), 256, 0, st>>>(wT, Kc, Mp, sC, nv, m, h->rk_part, RK_BLOCKS);
      CKL();
      k_sum_prod_final<<<batch, 1, 0, st>>>(h->rk_part, RK_BLOCKS, scal);
      CKL();
      dim3 grid(Mp / KT_M, (nv + KT_N - 1) / KT_N, batch);
      const size_t smem = (size_t)(KT_N * d + KT_M * d + d) * 8;
      k_build_kc<<<grid, KT_THREADS, smem, st>>>(Xc, nv, nv, d, Z, m, theta, kind, Kc, Mp, sC, 1);
      CKL();
    }
    k_transpose_rect<<<gT, bT, 0, st>>>(wT, Kc, Mp, sC, nullptr, 0, nv, Mp, tC, nvp, sC);
    CKL();
    k_phiT<<<dim3((nvp + 255) / 256, nq), 256, 0, st>>>(Xc, nv, d, tA, nvp);
    CKL();
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(tC, nvp, sC, tA, nvp, 0, mom, nq, (int64_t)m * nq, m, nq, nv, 1.0, 1.0), batch));
  }
  if (need_grad) {
    // dKzz = -Linv^T sym(Phi(Gbar)) Linv
    k_sym_phi<<<g16, b16, 0, st>>>(Gb, Hm, m, Mp, sM);
    CKL();
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->LinvT, Mp, sM, Hm, Mp, sM, h->T1, Mp, sM, m, m, m, 1.0, 0.0, KM_A_UPPER), batch));
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(h->T1, Mp, sM, h->LinvT, Mp, sM, dKzz, Mp, sM, m, m, m, -1.0, 0.0, KM_B_UPPER), batch));
    k_grad_kzz_rows<<<dim3((m + 7) / 8, batch), 256, 0, st>>>(dKzz, Mp, sM, Z, m, d, theta, kind, h->rowacc, dZzz + d + 2, sM);
    CKL();
    if (kind != GGP_KERNEL_RBF) k_grad_from_moments<<<batch, 256, 0, st>>>(mom, m, d, Z, theta, gk, sM, scal + 3, 1, 4);
    else k_grad_from_moments<<<batch, 256, 0, st>>>(mom, m, d, Z, theta, gk, sM, nullptr, 0);
    CKL();
  }
  k_svgp_final<<<batch, 256, 0, st>>>(scal, gk, sM, h->rowacc, dZzz, dm, Mp, dLsraw, sM, Mp, qm, sqm, qLs, theta, m, d, kl_scale, need_grad,
                                      elbo, grad, sG);
  CKL();
  return 0;
}

int ggp_svgp_predict(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* xs, int64_t ns, const double* Z,
                     const double* qm, int qm_batched, const double* qLs, const double* theta, const double* jitter, int m, int d,
                     int batch, double data_jitter, int add_noise, double* mean, double* var, int32_t* info) {
  const int64_t sqm = qm_batched ? m : 0;
  if (!h || !xs || !Z || !qm || !theta || !jitter || !mean || !var || !info) return fail(-1, "ggp_svgp_predict: NULL argument");
  if (!reserved_for(h, 0, m, d, batch)) return fail(-2, "ggp_svgp_predict: handle not reserved for this shape");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (kind.kind < GGP_KERNEL_RBF || kind.kind > GGP_KERNEL_RQ) return fail(-3, "ggp_svgp_predict: unknown kernel (composite kernels are SGPR-only)");
  cudaStream_t st = (cudaStream_t)stream;
  const int Mp = h->Mp, nsv = h->nsv;
  const int64_t sM = (int64_t)Mp * Mp, sC = (int64_t)nsv * Mp;
  const dim3 g16(Mp / 16, Mp / 16, batch), b16(16, 16);
  const bool hasS = qLs != nullptr;
  RUN(join_aux(h, st));
  double *Kc = h->sv[0], *aT = h->sv[1], *wT = h->sv[2], *LsP = h->Bm, *LsT = h->LBinv;
  k_build_kzz<<<g16, b16, 0, st>>>(Z, m, Mp, d, theta, jitter, kind, h->L, sM, h->piv_tol);
  CKL();
  RUN(chol_and_inverse(h, st, h->L, h->Linv, h->LinvT, batch, info));
  if (hasS) {
    k_pad_tril<<<dim3(Mp / 16, Mp / 16, 1), b16, 0, st>>>(qLs, m, LsP, Mp);
    CKL();
    k_transpose<<<dim3(Mp / 32, Mp / 32, 1), dim3(32, 8), 0, st>>>(LsP, LsT, Mp, sM);
    CKL();
  }
  for (int64_t c0 = 0; c0 < ns; c0 += nsv) {
    const int nv = (int)std::min<int64_t>(nsv, ns - c0);
    dim3 grid(Mp / KT_M, (nv + KT_N - 1) / KT_N, batch);
    const size_t smem = (size_t)(KT_N * d + KT_M * d + d) * 8;
    k_build_kc<<<grid, KT_THREADS, smem, st>>>(xs + c0 * d, nv, nv, d, Z, m, theta, kind, Kc, Mp, sC);
    CKL();
    RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(Kc, Mp, sC, h->Linv, Mp, sM, aT, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_LOWER), batch));
    if (hasS)
      RUN(launch_gemm(h, st, EPI_STORE, gemm_basic(aT, Mp, sC, LsT, Mp, 0, wT, Mp, sC, nv, m, m, 1.0, 0.0, KM_B_UPPER), batch));
    k_svgp_marginals<<<dim3((nv + 7) / 8, batch), 256, 0, st>>>(aT, hasS ? wT : nullptr, Mp, sC, qm, sqm, theta, d, m, nv, data_jitter,
                                                               add_noise, mean + c0, var + c0, ns);
    CKL();
  }
  return 0;
}

int ggp_chol_batched(ggp_handle_t* h, void* stream, double* a, double* linv, int m, int batch, int32_t* info) {
  if (!h || !a || !info) return fail(-1, "ggp_chol_batched: NULL argument");
  if (getenv("GGP_POTF2_TIMELINE") && h->arena && h->m == m) {   // developer switch: clock64 stamps of one diagonal-block factorisation
    long long* dbg = nullptr;
    CK(cudaMalloc((void**)&dbg, 16 * sizeof(long long)));
    CK(cudaMemset(dbg, 0, 16 * sizeof(long long)));
    const int Mp = h->Mp;
    k_pad_copy<<<dim3(Mp / 16, Mp / 16, 1), dim3(16, 16), 0, (cudaStream_t)stream>>>(a, m, h->L, Mp, (int64_t)Mp * Mp, 1);
    CK(cudaMemsetAsync(h->piv_tol, 0, sizeof(double) * batch, (cudaStream_t)stream));
    for (int r = 0; r < 2; ++r)
      k_potf2_trti2<<<1, 256, POTF2_SMEM, (cudaStream_t)stream>>>(h->L, Mp, (int64_t)Mp * Mp, 0, h->Tblk, (int64_t)Mp * Mp, info, h->piv_tol, dbg);
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    long long hb[16];
    CK(cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost));
    fprintf(stderr, "== potf2 timeline (clk): load %lld | micro-block 1: publish %lld, wait %lld... B(8x8 factor) %lld, barrier %lld, C %lld, D %lld | all 8 blocks %lld, store %lld\n",
            hb[1] - hb[0] - (hb[6] - hb[1]) * 0, hb[2] - hb[1], 0ll, hb[3] - hb[2], hb[4] - hb[3], hb[5] - hb[4], hb[6] - hb[5], hb[7] - hb[0], hb[8] - hb[7]);
    cudaFree(dbg);
  }
  if (!(h->arena && h->m == m && batch <= h->batch)) RUN(ggp_reserve(h, nullptr, 0, m, std::max(1, h->d), batch));
  cudaStream_t st = (cudaStream_t)stream;
  RUN(join_aux(h, st));
  const int Mp = h->Mp;
  const int64_t sM = (int64_t)Mp * Mp;
  const dim3 g16(Mp / 16, Mp / 16, batch), b16(16, 16);
  if (getenv("GGP_CHOL_TIMELINE") && !h->ct_dbg) {   // developer switch: stamps of the CTA on the dependent chain of every step
    CK(cudaMalloc((void**)&h->ct_dbg, 64 * 8 * sizeof(long long)));
    CK(cudaMemset(h->ct_dbg, 0, 64 * 8 * sizeof(long long)));
  }
  k_pad_copy<<<g16, b16, 0, st>>>(a, m, h->L, Mp, sM, 1);
  CKL();
  CK(cudaMemsetAsync(h->piv_tol, 0, sizeof(double) * batch, st));
  RUN(chol_and_inverse(h, st, h->L, h->Linv, h->LinvT, batch, info));
  if (h->ct_dbg && getenv("GGP_CHOL_TIMELINE")[0] == '2') {
    CK(cudaStreamSynchronize(st));
    long long hb[64 * 8];
    CK(cudaMemcpy(hb, h->ct_dbg, sizeof(hb), cudaMemcpyDeviceToHost));
    const int nblk = Mp / NB;
    for (int k = 0; k + 1 < nblk && k < 63; ++k) {
      const long long* s = hb + k * 8;
      fprintf(stderr, "== chol step %2d (clk): load %lld  panel products %lld  update product + store %lld  potf2 %lld | kernel %lld ns, gap to the next step %lld ns\n",
              k, s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4], s[6] - s[0], k + 2 < nblk ? hb[(k + 1) * 8] - s[6] : 0ll);
    }
  }
  k_pad_copy<<<g16, b16, 0, st>>>(a, m, h->L, Mp, sM, 0);
  CKL();
  if (linv) {
    k_pad_copy<<<g16, b16, 0, st>>>(linv, m, h->Linv, Mp, sM, 0);
    CKL();
  }
  return 0;
}

int ggp_gemm_nt(ggp_handle_t* h, void* stream, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                int64_t ldc, int mm, int nn, int kk, double alpha, double beta) {
  if (!h || !A || !B || !C) return fail(-1, "ggp_gemm_nt: NULL argument");
  if ((lda & 1) || (ldb & 1) || (((uintptr_t)A) & 15) || (((uintptr_t)B) & 15))
    return fail(-2, "ggp_gemm_nt: operands need 16-byte aligned rows (even leading dimension)");
  return launch_gemm(h, (cudaStream_t)stream, EPI_STORE, gemm_basic(A, lda, 0, B, ldb, 0, C, ldc, 0, mm, nn, kk, alpha, beta), 1);
}

int ggp_gemm_nt_ex(ggp_handle_t* h, void* stream, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                   int64_t ldc, int mm, int nn, int kk, double alpha, double beta, int kmode, int sym, int splits,
                   int64_t split_stride) {
  if (!h || !A || !B || !C) return fail(-1, "ggp_gemm_nt_ex: NULL argument");
  if ((lda & 1) || (ldb & 1) || (((uintptr_t)A) & 15) || (((uintptr_t)B) & 15))
    return fail(-2, "ggp_gemm_nt_ex: operands need 16-byte aligned rows (even leading dimension)");
  GemmP p = gemm_basic(A, lda, 0, B, ldb, 0, C, ldc, 0, mm, nn, kk, alpha, beta, kmode);
  p.sym = sym;
  p.splits = splits < 1 ? 1 : splits;
  p.sSplit = split_stride;
  p.heavy_first = (kmode & KM_A_LOWER) ? 1 : 0;
  return launch_gemm(h, (cudaStream_t)stream, EPI_STORE, p, 1);
}

int ggp_gemm_nt_i8(ggp_handle_t* h, void* stream, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                   int64_t ldc, int mm, int nn, int kk) {
  if (!h || !A || !B || !C) return fail(-1, "ggp_gemm_nt_i8: NULL argument");
  if (mm < 1 || nn < 1 || kk < 1 || kk > I8_MAX_K) return fail(-2, "ggp_gemm_nt_i8: bad shape (1 <= kk <= 16384)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t kp = (kk + 63) / 64 * 64;
  int8_t *qa = nullptr, *qb = nullptr;
  int *ea = nullptr, *eb = nullptr;
  CK(cudaMalloc((void**)&qa, (size_t)I8_NS * mm * kp));
  CK(cudaMalloc((void**)&qb, (size_t)I8_NS * nn * kp));
  CK(cudaMalloc((void**)&ea, (size_t)mm * 4));
  CK(cudaMalloc((void**)&eb, (size_t)nn * 4));
  k_slice_rows<<<(mm + 7) / 8, 256, 0, st>>>(A, mm, kk, lda, qa, kp, (int64_t)mm * kp, (int)kp, ea);
  CKL();
  k_slice_rows<<<(nn + 7) / 8, 256, 0, st>>>(B, nn, kk, ldb, qb, kp, (int64_t)nn * kp, (int)kp, eb);
  CKL();
  I8P p;
  memset(&p, 0, sizeof(p));
  p.M = mm; p.N = nn; p.K = (int)kp; p.splits = 1;
  p.ea = ea; p.eb = eb; p.alpha = 1.0; p.beta = 0.0;
  p.C = C; p.ldc = ldc;
  long long* dbg = nullptr;
  const char* tl = getenv("GGP_I8_TIMELINE");   // developer switch: print CTA 0's clock64 timeline of the first tiles to stderr
  if (tl && tl[0] == '1') {
    CK(cudaMalloc((void**)&dbg, (3 * I8_DBG_ITEMS * 4 + 2 * 256) * sizeof(long long)));   // 3 stamp rows + per-CTA wall-clock spans
    CK(cudaMemsetAsync(dbg, 0, (3 * I8_DBG_ITEMS * 4 + 2 * 256) * sizeof(long long), st));
    p.dbg = dbg;
  }
  cudaEvent_t te0 = nullptr, te1 = nullptr;
  const bool exp_time = getenv("GGP_I8_EXP_TIME") != nullptr;   // developer switch: time the GEMM launch alone (5 repeats) and print it
  if (exp_time) { cudaEventCreate(&te0); cudaEventCreate(&te1); }
  int rc = launch_i8(h, st, I8_EPI_F64, p, {qa, mm, kp, (int64_t)mm * kp}, {qb, nn, kp, (int64_t)nn * kp});
  if (exp_time && rc == 0) {
    cudaEventRecord(te0, st);
    for (int r = 0; r < 5 && rc == 0; ++r) rc = launch_i8(h, st, I8_EPI_F64, p, {qa, mm, kp, (int64_t)mm * kp}, {qb, nn, kp, (int64_t)nn * kp});
    cudaEventRecord(te1, st);
    cudaEventSynchronize(te1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, te0, te1);
    ms /= 5.f;
    const double tiles = (double)((mm + I8_BM - 1) / I8_BM) * ((nn + I8_BN - 1) / I8_BN), kbs = (double)kp / I8_BKB;
    const double per_cta = tiles * kbs / std::min<double>(tiles, h->sm_count);
    fprintf(stderr, "== i8 gemm %d x %d x %d: %.3f ms, %.0f TOP/s (int8, 28 digit products), %.1f ns per k-block and CTA\n", mm, nn, kk, ms,
            28.0 * 2.0 * mm * nn * (double)kp / (ms * 1e-3) / 1e12, ms * 1e6 / per_cta);
    cudaEventDestroy(te0); cudaEventDestroy(te1);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (dbg) {
    long long hb[2 * I8_DBG_ITEMS * 4];
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    const long long t0 = hb[0];
    for (int it = 0; it < I8_DBG_ITEMS; ++it) {
      const long long* m = hb + it * 4;
      const long long* ep = hb + (I8_DBG_ITEMS + it) * 4;
      fprintf(stderr, "tile %2d  mma: start %8lld  tmem_free %8lld  first_kb_issued %8lld  all_issued %8lld | epi: wait %8lld  full %8lld  drained %8lld  done %8lld\n",
              it, m[0] - t0, m[1] - t0, m[2] - t0, m[3] - t0, ep[0] - t0, ep[1] - t0, ep[2] - t0, ep[3] - t0);
    }
    cudaFree(dbg);
  }
  cudaFree(qa); cudaFree(qb); cudaFree(ea); cudaFree(eb);
  if (rc != 0) return rc;
  CK(e);
  return 0;
}

int ggp_kernel_matrix(ggp_handle_t* h, const ggp_cfg* cfg, void* stream, const double* X1, int64_t n1, const double* X2,
                      int64_t n2, const double* theta, int d, double* out) {
  if (!h || !X1 || !X2 || !theta || !out) return fail(-1, "ggp_kernel_matrix: NULL argument");
  if ((size_t)(2 * KT_N * d + d) * 8 > 200 * 1024) return fail(-3, "ggp_kernel_matrix: d too large");
  const KSpec kind{cfg ? cfg->kernel : 0, cfg ? cfg->kernel_param : 0.0};
  if (!kernel_ok(h, kind)) return fail(-3, "ggp_kernel_matrix: unknown kernel");
  if (kind == GGP_KERNEL_COMPOSITE) {   // theta is ignored: the registered parameter row 0
    if (int r = prog_ready(h, d, "ggp_kernel_matrix")) return r;
    k_build_kc_prog<<<dim3((unsigned)((n2 + 31) / 32), (unsigned)((n1 + 7) / 8), 1), 256, 0, (cudaStream_t)stream>>>(
        X1, (int)n1, X2, (int)n2, (int)n2, h->prog, h->kth, out, n2, 0);
    CKL();
    return 0;
  }
  dim3 grid((unsigned)((n2 + KT_M - 1) / KT_M), (unsigned)((n1 + KT_N - 1) / KT_N), 1);
  const size_t smem = (size_t)(KT_N * d + KT_M * d + d) * 8;
  k_build_kc<<<grid, KT_THREADS, smem, (cudaStream_t)stream>>>(X1, (int)n1, (int)n1, d, X2, (int)n2, theta, kind, out, n2, 0);
  CKL();
  return 0;
}

int ggp_profile_enable(ggp_handle_t* h, int on) {
  if (!h) return fail(-1, "ggp_profile_enable: handle is NULL");
  h->profiling = on != 0;
  return 0;
}

int ggp_profile_read(ggp_handle_t* h, double* ms_out, int64_t* spans_out, int64_t* launches_out) {
  if (!h) return fail(-1, "ggp_profile_read: handle is NULL");
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  for (int c = 0; c < CAT_COUNT; ++c) {
    if (ms_out) ms_out[c] = 0.0;
    if (spans_out) spans_out[c] = 0;
  }
  for (auto& sp : h->spans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, sp.e0, sp.e1);
    if (ms_out) ms_out[sp.cat] += ms;
    if (spans_out) spans_out[sp.cat] += 1;
    h->pool.push_back(sp.e0);
    h->pool.push_back(sp.e1);
  }
  h->spans.clear();
  if (launches_out) *launches_out = h->launches;
  h->launches = 0;
  return 0;
}

int ggp_probe_dmma_peak(ggp_handle_t* h, void* stream, int iters, double* tflops_out) {
  if (!h || !tflops_out) return fail(-1, "ggp_probe_dmma_peak: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* sink;
  CK(cudaMalloc((void**)&sink, 8 * 1024 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int warps_per_sm[3] = {8, 16, 32};
  const int frag_rows[3] = {8, 4, 2};
  double best = 0.0;
  for (int v = 0; v < 3; ++v) {
    double bestv = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0, st));
      if (v == 0) k_dmma_probe<8, 256><<<h->sm_count, 256, 0, st>>>(sink, iters);
      if (v == 1) k_dmma_probe<4, 512><<<h->sm_count, 512, 0, st>>>(sink, iters);
      if (v == 2) k_dmma_probe<2, 1024><<<h->sm_count, 1024, 0, st>>>(sink, iters);
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      CKL();
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = (double)h->sm_count * warps_per_sm[v] * (double)iters * frag_rows[v] * 4.0 * 512.0;
      bestv = std::max(bestv, flops / (ms * 1e-3) / 1e12);
    }
    tflops_out[1 + v] = bestv;
    best = std::max(best, bestv);
  }
  tflops_out[0] = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

// ---- NUTS tree bookkeeping (csrc/nuts.cuh) ----
int ggp_nuts_state_size(void) { return (int)sizeof(ggp_nuts_state); }
static int nuts_check(const ggp_nuts_state* s) {
  if (!s || s->C < 1 || s->P < 1 || s->K < 1 || !s->x || !s->x_eval || !s->lp_eval || !s->g_eval || !s->u) return fail(-2, "ggp_nuts: bad state");
  return 0;
}
int ggp_nuts_begin(void* stream, const ggp_nuts_state* s, const double* z) {
  if (int r = nuts_check(s)) return r;
  if (!z) return fail(-3, "ggp_nuts_begin: z is null");
  k_nuts_begin<<<s->C, 32, 0, (cudaStream_t)stream>>>(*s, z);
  CK(cudaGetLastError());
  return 0;
}
int ggp_nuts_subtree_begin(void* stream, const ggp_nuts_state* s) {
  if (int r = nuts_check(s)) return r;
  k_nuts_subtree_begin<<<s->C, 32, 0, (cudaStream_t)stream>>>(*s);
  CK(cudaGetLastError());
  return 0;
}
int ggp_nuts_leaf(void* stream, const ggp_nuts_state* s) {
  if (int r = nuts_check(s)) return r;
  k_nuts_leaf<<<s->C, 32, 0, (cudaStream_t)stream>>>(*s);
  CK(cudaGetLastError());
  return 0;
}
int ggp_nuts_subtree_end(void* stream, const ggp_nuts_state* s) {
  if (int r = nuts_check(s)) return r;
  k_nuts_subtree_end<<<s->C, 32, 0, (cudaStream_t)stream>>>(*s);
  CK(cudaGetLastError());
  return 0;
}
int ggp_nuts_end(void* stream, const ggp_nuts_state* s, int k) {
  if (int r = nuts_check(s)) return r;
  k_nuts_end<<<s->C, 32, 0, (cudaStream_t)stream>>>(*s, k);
  CK(cudaGetLastError());
  return 0;
}

int ggp_vfe_theta(void* stream, const double* x, int C, int d, double* theta) {
  if (!x || !theta || C < 1 || d < 1) return fail(-2, "ggp_vfe_theta: bad argument");
  k_vfe_theta<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, C, d, theta);
  CK(cudaGetLastError());
  return 0;
}
int ggp_vfe_logp(void* stream, const double* x, const double* bound, const double* grad, int64_t ldg, const int* info,
                 const int* info_b, int C, int d, int with_prior, double* lp, double* dx) {
  if (!x || !bound || !grad || !lp || !dx || C < 1 || d < 1 || ldg < d + 2) return fail(-2, "ggp_vfe_logp: bad argument");
  k_vfe_logp<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, bound, grad, ldg, info, info_b, C, d, with_prior, lp, dx);
  CK(cudaGetLastError());
  return 0;
}

int ggp_probe_i8_peak(ggp_handle_t* h, void* stream, int iters, double* tops_out) {
  if (!h || !tops_out || iters < 1) return fail(-1, "ggp_probe_i8_peak: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaFuncSetAttribute(k_i8_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_PROBE_SMEM));
  CK(cudaFuncSetAttribute(k_i8_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_PROBE_SMEM));
  CK(cudaFuncSetAttribute(k_i8_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_PROBE_SMEM));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  // [1] uniform 128 x 256 x 32 MMAs, 4 iterations in flight; [2] production mix (128 x 64 tiles), 2 in flight (the 2-stage ring);
  // [3] the mix of 128 x 32 tiles (7 MMAs per k-step, N = 224 .. 32), 2 in flight
  const int mode[3] = {0, 1, 2}, depth[3] = {4, 2, 2};
  const double macs_per_iter[3] = {16.0 * 128 * 256 * 32, 2.0 * 128 * 1792 * 32, 2.0 * 128 * 1792 * 32};
  double best = 0.0;
  for (int v = 0; v < 3; ++v) {
    double bestv = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0, st));
      if (mode[v] == 0) k_i8_probe<0><<<h->sm_count, 128, I8_PROBE_SMEM, st>>>(iters, depth[v], 17u + rep);
      else if (mode[v] == 1) k_i8_probe<1><<<h->sm_count, 128, I8_PROBE_SMEM, st>>>(iters, depth[v], 17u + rep);
      else k_i8_probe<2><<<h->sm_count, 128, I8_PROBE_SMEM, st>>>(iters, depth[v], 17u + rep);
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      CKL();
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      bestv = std::max(bestv, 2.0 * macs_per_iter[v] * iters * h->sm_count / (ms * 1e-3) / 1e12);
    }
    tops_out[1 + v] = bestv;
    best = std::max(best, bestv);
  }
  tops_out[0] = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

}  // extern "C"
