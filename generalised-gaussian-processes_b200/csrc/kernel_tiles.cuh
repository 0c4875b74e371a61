// Fused ARD kernel-tile construction: k(x_n, z_m) for a streamed chunk of X rows, and the m x m Kzz.
// X rows are staged into shared memory with one 1-D bulk TMA copy per tile (cp.async.bulk -> UBLKCP) when the tile is
// full and 16-byte aligned; Z is pre-scaled by 1/ell into shared memory; each thread produces a 4 x 4 register tile
// (rows 4*ty+i, columns tx+16*j: shared reads are conflict-free and global stores are 128-byte coalesced).
#pragma once
#include "common.cuh"
#include "gemm_i8.cuh"

namespace ggp {

constexpr int KT_N = 64, KT_M = 64, KT_THREADS = 256;

// kernel kind + its shape parameter (alpha of the rational-quadratic kernel; unused otherwise)
struct KSpec {
  int kind;
  double p;
  __host__ __device__ bool operator!=(int k) const { return kind != k; }
  __host__ __device__ bool operator==(int k) const { return kind == k; }
};

// value of the stationary kernel given scaled squared distance d2 (>= 0).  kind: 0 rbf, 1 matern32, 2 matern52,
// 3 rational quadratic (1 + d2 / (2 alpha))^(-alpha)  (gpytorch RQKernel / pymc3 RatQuad, experiments/co2_bayesian_sgpr_hmc.py:77,127)
__device__ __forceinline__ double kval(KSpec ks, double sf2, double d2) {
  const int kind = ks.kind;
  if (kind == 0) return sf2 * exp(-0.5 * d2);
  if (kind == 3) return sf2 * exp(-ks.p * log1p(d2 / (2.0 * ks.p)));
  const double r = sqrt(d2);
  if (kind == 1) {
    const double a = 1.7320508075688772;
    return sf2 * (1.0 + a * r) * exp(-a * r);
  }
  const double a = 2.23606797749979;
  return sf2 * (1.0 + a * r + (5.0 / 3.0) * d2) * exp(-a * r);
}
// dk/d(d2)
__device__ __forceinline__ double kgrad(KSpec ks, double sf2, double d2) {
  const int kind = ks.kind;
  if (kind == 0) return -0.5 * sf2 * exp(-0.5 * d2);
  if (kind == 3) return -0.5 * sf2 * exp(-(ks.p + 1.0) * log1p(d2 / (2.0 * ks.p)));
  const double r = sqrt(d2);
  if (kind == 1) return -1.5 * sf2 * exp(-1.7320508075688772 * r);
  const double a = 2.23606797749979;
  return -(5.0 / 6.0) * sf2 * (1.0 + a * r) * exp(-a * r);
}

// exp(x) for x <= 0 with a 64-entry table of 2^(j/64) (hi + lo) and a degree-6 polynomial on |r| <= ln2/128: the tile build is
// issue-bound (131 instructions per element, ~45 of them CUDA's exp); this form is ~19.  Error <= 0.8 ulp (checked against a
// long-double reference over [-700, 0] on the host: scripts/exp_table_check.py); x < -700 flushes to 0 (the digit planes resolve
// 2^-56 of sf2 anyway).  tab points at 64 (hi, lo) pairs in shared memory.
__device__ const double GGP_EXP2_TAB[128] = {
    1.0, 0.0,
    1.0108892860517005, -1.5234778603368577e-17,
    1.0218971486541166, 5.109225028973444e-17,
    1.0330248790212284, 7.600838874027088e-18,
    1.0442737824274138, 8.551889705537965e-17,
    1.0556451783605572, 1.759325738772092e-18,
    1.0671404006768237, -7.899853966841582e-17,
    1.0787607977571199, -6.656660436056593e-17,
    1.0905077326652577, -3.046782079812471e-17,
    1.102382583307841, 5.2660368715706944e-17,
    1.1143867425958924, 1.0410278456845571e-16,
    1.1265216186082418, 5.165856758795457e-17,
    1.1387886347566916, 8.912812676025408e-17,
    1.1511892299529827, 3.250710218863827e-17,
    1.1637248587775775, 3.8292048369240935e-17,
    1.1763969916502812, 5.554203254218079e-17,
    1.189207115002721, 3.982015231465646e-17,
    1.202156731452703, 6.644981499252301e-17,
    1.215247359980469, -7.712630692681488e-17,
    1.22848053610687, -1.89878163130253e-17,
    1.241857812073484, 4.658027591836937e-17,
    1.255380757024691, -6.7113898212968784e-18,
    1.2690509571917332, 2.667932131342186e-18,
    1.2828700160787783, 1.713594918243561e-17,
    1.2968395546510096, 2.5382502794888315e-17,
    1.3109612115247644, -7.181536135519454e-17,
    1.3252366431597413, -2.8587312100388614e-17,
    1.339667524053303, 8.927282594831732e-17,
    1.3542555469368927, 7.70094837980299e-17,
    1.3690024229745905, 9.593797919118849e-17,
    1.383909881963832, -6.770511658794786e-17,
    1.3989796725383112, -9.614213209051323e-17,
    1.4142135623730951, -9.667293313452913e-17,
    1.42961333839197, -1.2031642489053655e-17,
    1.4451808069770467, -3.0237581349939873e-17,
    1.460917794180647, -5.600377186075216e-17,
    1.4768261459394993, -3.483994556892796e-17,
    1.4929077282912648, 1.4192920154284036e-17,
    1.5091644275934228, -1.016455327754295e-16,
    1.5255981507445384, -1.1024941712342561e-16,
    1.5422108254079407, 7.949834809697621e-17,
    1.559004400237837, 3.7812070533575275e-17,
    1.5759808451078865, -1.0136916471278304e-17,
    1.593142151342267, -1.0094406542311964e-16,
    1.6104903319492543, 2.4707192569797888e-17,
    1.6280274218573478, -6.712955084707084e-17,
    1.645755478153965, -1.0125679913674773e-16,
    1.6636765803267364, 5.8909926967131e-17,
    1.681792830507429, 8.199010020581497e-17,
    1.7001063537185235, -8.0237193703977e-18,
    1.718619298122478, -1.851380418263111e-17,
    1.7373338352737062, 3.164389299292957e-17,
    1.7562521603732995, 2.960140695448873e-17,
    1.7753764925265212, 6.429731796556572e-17,
    1.7947090750031072, 1.8227458427912087e-17,
    1.8142521755003989, -9.969531538920349e-17,
    1.8340080864093424, 3.283107224245627e-17,
    1.8539791250833855, 9.761887490727594e-17,
    1.8741676341103, -6.122763413004143e-17,
    1.8945759815869656, 3.4034035352165297e-17,
    1.9152065613971474, -1.0619946056195963e-16,
    1.9360617934922943, 1.0332385960676326e-16,
    1.9571441241754002, 8.960767791036668e-17,
    1.978456026387951, 4.0388753109278167e-17};
__device__ __forceinline__ double exp_neg(double x, const double2* __restrict__ tab) {
  if (!(x >= -700.0)) return (x != x) ? x : 0.0;
  const double t = fma(x, 92.33248261689366 /* 64 / ln 2 */, 6755399441055744.0 /* 1.5 * 2^52: round to nearest integer */);
  const int ki = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(-kf, 0.010830424696249145 /* ln2/64 hi */, x);
  r = fma(-kf, 3.623510646634843e-19 /* ln2/64 lo */, r);
  double p = 1.0 / 720.0;
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p * r, r, r);   // exp(r) - 1
  const double2 T = tab[ki & 63];
  const double res = T.x + fma(T.x, p, T.y);   // in [1, 2)
  return __hiloint2double(__double2hiint(res) + ((ki >> 6) << 20), __double2loint(res));   // * 2^(ki >> 6), exponent stays > 0
}
// kval with the table exponential (streamed tile build of the sliced-integer path)
__device__ __forceinline__ double kval_tab(KSpec ks, double sf2, double d2, const double2* __restrict__ tab) {
  const int kind = ks.kind;
  if (kind == 0) return sf2 * exp_neg(-0.5 * d2, tab);
  if (kind == 3) return sf2 * exp_neg(-ks.p * log1p(d2 / (2.0 * ks.p)), tab);
  const double r = sqrt(d2);
  if (kind == 1) {
    const double a = 1.7320508075688772;
    return sf2 * (1.0 + a * r) * exp_neg(-a * r, tab);
  }
  const double a = 2.23606797749979;
  return sf2 * (1.0 + a * r + (5.0 / 3.0) * d2) * exp_neg(-a * r, tab);
}

// Kc[b][n][m] = k(x_n, z_m; theta_b) for n in [0, n_fill): rows >= n_valid and columns >= M are written as zero.
// deriv = 1 emits dk/d(d2) instead (the multiplier of the backward epilogue for the non-RBF kernels).
// grid: (ceil(ldk/KT_M), ceil(n_fill/KT_N), batch)
__global__ void __launch_bounds__(KT_THREADS) k_build_kc(const double* __restrict__ X, int n_valid, int n_fill, int d,
                                                        const double* __restrict__ Z, int M,
                                                        const double* __restrict__ theta, KSpec kind,
                                                        double* __restrict__ Kc, int64_t ldk, int64_t sK, int deriv = 0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);   // [KT_N][d]
  double* zs = xs + KT_N * d;                          // [d][KT_M]   (z / ell, transposed: conflict-free)
  double* il = zs + KT_M * d;                          // [d]
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) double2 etab[64];            // 2^(j/64) table of the exponential (exp_neg)

  const int b = blockIdx.z;
  if (threadIdx.x < 64) etab[threadIdx.x] = make_double2(GGP_EXP2_TAB[2 * threadIdx.x], GGP_EXP2_TAB[2 * threadIdx.x + 1]);
  const double* th = theta + (int64_t)b * (d + 2);
  const double sf2 = th[d];
  const int n0 = blockIdx.y * KT_N, m0 = blockIdx.x * KT_M;
  const int tid = threadIdx.x;
  const int rows = min(KT_N, n_valid - n0);  // may be <= 0

  const bool bulk = (rows == KT_N) && ((((uintptr_t)(X + (int64_t)n0 * d)) & 15) == 0) && (((KT_N * d * 8) & 15) == 0);
  if (bulk) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar, KT_N * d * 8);
      tma_bulk_g2s(xs, X + (int64_t)n0 * d, KT_N * d * 8, &bar);
    }
  } else {
    for (int i = tid; i < KT_N * d; i += KT_THREADS) {
      const int r = i / d;
      xs[i] = (r < rows) ? X[(int64_t)n0 * d + i] : 0.0;
    }
  }
  for (int i = tid; i < d; i += KT_THREADS) il[i] = 1.0 / th[i];
  for (int i = tid; i < KT_M * d; i += KT_THREADS) {
    const int r = i / d, c = i % d;
    zs[c * KT_M + r] = (m0 + r < M) ? Z[(int64_t)(m0 + r) * d + c] / th[c] : 0.0;
  }
  if (bulk) mbar_wait(&bar, 0);
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;  // tx -> m, ty -> n
  double d2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d2[i][j] = 0.0;
  for (int c = 0; c < d; ++c) {
    const double ic = il[c];
    double xv[4], zv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = xs[(ty * 4 + i) * d + c] * ic;
#pragma unroll
    for (int j = 0; j < 4; ++j) zv[j] = zs[c * KT_M + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double t = xv[i] - zv[j];
        d2[i][j] = fma(t, t, d2[i][j]);
      }
  }
  double* out = Kc + (int64_t)b * sK;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= n_fill) continue;
    const bool nv = n < n_valid;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + tx + 16 * j;
      if (m >= ldk) continue;
      // values through the table exponential (<= 0.8 ulp, ~19 instructions instead of ~45: this kernel is 6-7 % of a chain-batched
      // SGPMC evaluation); the derivative tile keeps the library exp
      out[(int64_t)n * ldk + m] = (nv && m < M) ? (deriv ? kgrad(kind, sf2, d2[i][j]) : kval_tab(kind, sf2, d2[i][j], etab)) : 0.0;
    }
  }
}

// Tile build of the sliced-integer path (gemm_i8.cuh must be included first): one pass produces
//   * the FP64 tile Kc[n][m] (multiplier of the backward epilogue),
//   * its int8 digit planes Kq[i][n][m] with the fixed exponent e = i8_exp_for(sf2) (k <= sf2).
// Thread (tx, ty) owns rows 4 ty + i and the 4 CONSECUTIVE columns 4 tx + j: 32-byte FP64 stores and 4-byte digit stores, a half-warp
// covers 512 / 64 contiguous bytes of one row.  zs is laid out [d][4][16] so that the shared reads stay conflict-free.
// A CTA stages its 64 scaled z rows ONCE and walks KT_RT row tiles of 64 x rows, the next x tile in flight (bulk TMA, two buffers)
// while the current one is computed: with one row tile per CTA the prologue (z staging, divisions, TMA latency) cost as much as the
// arithmetic (ncu: 36 % warps active, 65 % issue slots, FP64 pipe 37 %).
// grid: (ldk / 64, ceil(n_valid / (64 * KT_RT))); batch = 1.  dynamic shared memory: (2 * 64 * d + 64 * d + d) doubles.
constexpr int KT_RT = 8;
#ifndef GGP_KT_MINB
#define GGP_KT_MINB 1
#endif
__global__ void __launch_bounds__(KT_THREADS, GGP_KT_MINB) k_build_kc_i8(const double* __restrict__ X, int64_t n_valid,
                                                           int d, const double* __restrict__ Z, int M, const double* __restrict__ theta,
                                                           KSpec kind, double* __restrict__ Kc, int64_t ldk, int8_t* __restrict__ Kq,
                                                           int64_t ldq, int64_t plane) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs0 = reinterpret_cast<double*>(smem_raw);   // [2][KT_N][d]
  double* zs = xs0 + 2 * KT_N * d;                      // [d][4][16]   (z / ell)
  double* il = zs + KT_M * d;                           // [d]
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ __align__(16) double2 etab[64];

  const double sf2 = theta[d];
  const int m0 = blockIdx.x * KT_M;
  const int tid = threadIdx.x;
  const int64_t nbase = (int64_t)blockIdx.y * KT_RT * KT_N;
  const int tile_bytes = KT_N * d * 8;
  // a row tile is fetched by one bulk copy when it is full and 16-byte aligned (always, except the ragged last tile / odd d)
  auto is_bulk = [&](int64_t n0) {
    return n0 + KT_N <= n_valid && ((((uintptr_t)(X + n0 * d)) & 15) == 0) && ((tile_bytes & 15) == 0);
  };
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (tid < 64) etab[tid] = make_double2(GGP_EXP2_TAB[2 * tid], GGP_EXP2_TAB[2 * tid + 1]);
  for (int i = tid; i < d; i += KT_THREADS) il[i] = 1.0 / theta[i];
  for (int i = tid; i < KT_M * d; i += KT_THREADS) {
    const int r = i / d, c = i % d;
    zs[c * KT_M + (r & 3) * 16 + (r >> 2)] = (m0 + r < M) ? Z[(int64_t)(m0 + r) * d + c] / theta[c] : 0.0;
  }
  __syncthreads();
  if (tid == 0 && nbase < n_valid && is_bulk(nbase)) {
    mbar_arrive_expect_tx(&bar[0], tile_bytes);
    tma_bulk_g2s(xs0, X + nbase * d, tile_bytes, &bar[0]);
  }
  const int tx = tid & 15, ty = tid >> 4;  // tx -> columns 4 tx + j, ty -> rows 4 ty + i
  const double si = exp2((double)-i8_exp_for(sf2));
  for (int rt = 0; rt < KT_RT; ++rt) {
    const int64_t n0 = nbase + (int64_t)rt * KT_N;
    if (n0 >= n_valid) break;
    double* xs = xs0 + (rt & 1) * KT_N * d;
    // next tile into the other buffer (all threads finished reading it: barrier at the end of the previous iteration)
    if (tid == 0 && rt + 1 < KT_RT && n0 + KT_N < n_valid && is_bulk(n0 + KT_N)) {
      mbar_arrive_expect_tx(&bar[(rt + 1) & 1], tile_bytes);
      tma_bulk_g2s(xs0 + ((rt + 1) & 1) * KT_N * d, X + (n0 + KT_N) * d, tile_bytes, &bar[(rt + 1) & 1]);
    }
    if (is_bulk(n0)) {
      mbar_wait(&bar[rt & 1], (rt >> 1) & 1);
    } else {
      const int rows = (int)min((int64_t)KT_N, n_valid - n0);
      for (int i = tid; i < KT_N * d; i += KT_THREADS) xs[i] = (i / d < rows) ? X[n0 * d + i] : 0.0;
      __syncthreads();
    }
    double d2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) d2[i][j] = 0.0;
    for (int c = 0; c < d; ++c) {
      const double ic = il[c];
      double xv[4], zv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[(ty * 4 + i) * d + c] * ic;
#pragma unroll
      for (int j = 0; j < 4; ++j) zv[j] = zs[c * KT_M + j * 16 + tx];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double t = xv[i] - zv[j];
          d2[i][j] = fma(t, t, d2[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t n = n0 + ty * 4 + i;
      const bool nv = n < n_valid;
      double kv[4];
      uint32_t pk[I8_NS];
      unsigned long long V[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // the FP64 tile keeps the UNQUANTISED value: it is the dk/dtheta factor of the backward epilogue, where the fixed-point grid
        // (relative error 3e-12 at k = 1e-5) moved the gradient by 2.5e-8 at the headline shape
        kv[j] = (nv && m0 + 4 * tx + j < M) ? kval_tab(kind, sf2, d2[i][j], etab) : 0.0;
        V[j] = i8_digit_bytes_of(kv[j] * si);
      }
      i8_pack4(V[0], V[1], V[2], V[3], pk);
      if (nv) {
        double* o = Kc + n * ldk + m0 + 4 * tx;
        *reinterpret_cast<double2*>(o) = make_double2(kv[0], kv[1]);
        *reinterpret_cast<double2*>(o + 2) = make_double2(kv[2], kv[3]);
#pragma unroll
        for (int q = 0; q < I8_NS; ++q) *reinterpret_cast<uint32_t*>(Kq + (int64_t)q * plane + n * ldq + m0 + 4 * tx) = pk[q];
      }
    }
    __syncthreads();   // every thread is done with xs[rt & 1] before it is refilled two iterations later
  }
}

// Kzz[b][i][j] = k(z_i, z_j) + jitter_b * delta_ij for i,j < M ; identity on the padding up to Mp.  grid: (Mp/16, Mp/16, batch)
// piv_tol[b] (optional) = GGP_PIVOT_RTOL * (k(z,z) + jitter_b): the Cholesky reports "not positive definite" for a pivot at or below it.
// LAPACK's potrf (what psd_safe_cholesky calls upstream) only fails for a pivot <= 0, but with exactly duplicated inducing rows
// (np.random.randint draws WITH replacement, experiments/regression.py:83) the true pivot is 0 and the computed one is rounding
// noise of either sign: LAPACK's outcome is a coin flip no other implementation can reproduce, and proceeding on a noise pivot
// means cond(Kzz) ~ 1e16.  The relative threshold makes the ladder deterministic; oracle/linalg.py applies the same rule.
constexpr double GGP_PIVOT_RTOL = 1e-12;
__global__ void k_build_kzz(const double* __restrict__ Z, int M, int Mp, int d, const double* __restrict__ theta,
                            const double* __restrict__ jitter, KSpec kind, double* __restrict__ Kzz, int64_t sK,
                            double* __restrict__ piv_tol = nullptr) {
  const int b = blockIdx.z;
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= Mp || j >= Mp) return;
  const double* th = theta + (int64_t)b * (d + 2);
  if (i == 0 && j == 0 && piv_tol) piv_tol[b] = GGP_PIVOT_RTOL * (kval(kind, th[d], 0.0) + (jitter ? jitter[b] : 0.0));
  double v;
  if (i < M && j < M) {
    double d2 = 0.0;
    for (int c = 0; c < d; ++c) {
      const double t = (Z[(int64_t)i * d + c] - Z[(int64_t)j * d + c]) / th[c];
      d2 = fma(t, t, d2);
    }
    v = kval(kind, th[d], d2);
    if (i == j && jitter) v += jitter[b];
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  Kzz[(int64_t)b * sK + (int64_t)i * Mp + j] = v;
}

}  // namespace ggp
