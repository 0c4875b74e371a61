// Whitened SVGP / SGPMC marginals, expected log-likelihood and its analytic backward (row kernels + small helpers).
//   models/svgp.py:104-106 (VariationalStrategy + VariationalELBO), models/sgp_hmc.py:63 (SGPMC.log_posterior_density)
// The dense products (a = L^{-1} k(Z,x), L_s^T a, S a, G_A L^{-1}, G_A A^T, ...) reuse the DMMA GEMM; these kernels are
// the per-row / element-wise glue between them.
#pragma once
#include "common.cuh"

namespace ggp {

// numpy.polynomial.hermite.hermgauss(20)
__constant__ double c_gh_x[20] = {
    -5.38748089001123276e+00, -4.60368244955074424e+00, -3.94476404011562520e+00, -3.34785456738321630e+00,
    -2.78880605842813045e+00, -2.25497400208927568e+00, -1.73853771211658614e+00, -1.23407621539532308e+00,
    -7.37473728545394391e-01, -2.45340708300901239e-01, 2.45340708300901239e-01, 7.37473728545394391e-01,
    1.23407621539532308e+00, 1.73853771211658614e+00, 2.25497400208927568e+00, 2.78880605842813045e+00,
    3.34785456738321630e+00, 3.94476404011562520e+00, 4.60368244955074424e+00, 5.38748089001123276e+00};
__constant__ double c_gh_w[20] = {
    2.22939364553414471e-13, 4.39934099227317473e-10, 1.08606937076927821e-07, 7.80255647853205987e-06,
    2.28338636016353646e-04, 3.24377334223785669e-03, 2.48105208874636433e-02, 1.09017206020023294e-01,
    2.86675505362834149e-01, 4.62243669600610085e-01, 4.62243669600610085e-01, 2.86675505362834149e-01,
    1.09017206020023294e-01, 2.48105208874636433e-02, 3.24377334223785669e-03, 2.28338636016353646e-04,
    7.80255647853205987e-06, 1.08606937076927821e-07, 4.39934099227317473e-10, 2.22939364553414471e-13};

// log Phi(x) and phi(x)/Phi(x), stable in both tails
__device__ __forceinline__ void log_ndtr_and_ratio(double x, double& lphi, double& ratio) {
  const double RS2 = 0.7071067811865476, ISQ2PI = 0.3989422804014327;
  if (x < 0.0) {
    const double e = erfcx(-x * RS2);  // Phi(x) = exp(-x^2/2) erfcx(-x/sqrt2) / 2
    lphi = log(0.5 * e) - 0.5 * x * x;
    ratio = ISQ2PI / (0.5 * e);
  } else {
    const double P = 0.5 * erfc(-x * RS2);
    lphi = log(P);
    ratio = ISQ2PI * exp(-0.5 * x * x) / P;
  }
}

// lower-masked, zero-padded copy of the raw variational Cholesky factor:  LsP[i][j] = (j <= i < M) ? Ls[i][j] : 0
__global__ void k_pad_tril(const double* __restrict__ Ls, int M, double* __restrict__ LsP, int Mp) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= Mp || j >= Mp) return;
  LsP[(int64_t)i * Mp + j] = (i < M && j <= i) ? Ls[(int64_t)i * M + j] : 0.0;
}

// One warp per data row n.  mu = aT[n,:].m ; var = sf2 + data_jitter + ||wT[n,:]||^2 - ||aT[n,:]||^2
// rowout[b][0][n] = lik_scale * E_q[log p(y_n|f_n)], [1] = d/dmu, [2] = d/dvar, [3] = d/ds2   (row stride ldr)
// grid (ceil(nv/8), batch), block 256.  wT may be NULL (S = 0, SGPMC).
__global__ void __launch_bounds__(256) k_svgp_rows(const double* __restrict__ aT, const double* __restrict__ wT, int64_t ld,
                                                   int64_t sC, const double* __restrict__ qm, int64_t sqm, const double* __restrict__ yv,
                                                   const double* __restrict__ theta, int d, int M, int nv, int likelihood,
                                                   double data_jitter, double lik_scale, double* __restrict__ rowout,
                                                   int64_t ldr) {
  const int b = blockIdx.y, n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= nv) return;
  qm += b * sqm;   // sqm = 0: one variational mean for every batch element; m: one whitened vector v per chain (SGPMC)
  const double* a = aT + b * sC + (int64_t)n * ld;
  double mu = 0.0, v1 = 0.0, v2 = 0.0;
  for (int j = lane; j < M; j += 32) {
    const double aj = a[j];
    mu = fma(aj, qm[j], mu);
    v2 = fma(aj, aj, v2);
  }
  if (wT) {
    const double* w = wT + b * sC + (int64_t)n * ld;
    for (int j = lane; j < M; j += 32) v1 = fma(w[j], w[j], v1);
  }
  mu = warp_sum(mu);
  v1 = warp_sum(v1);
  v2 = warp_sum(v2);
  // (the butterfly sums leave the totals in every lane)
  const double sf2 = theta[(int64_t)b * (d + 2) + d], s2 = theta[(int64_t)b * (d + 2) + d + 1];
  const double var = sf2 + data_jitter + v1 - v2;
  const double y = yv[n];
  double ell, gmu, gv, gs = 0.0;
  if (likelihood == 0) {
    const double r = y - mu;
    ell = -0.5 * ((r * r + var) / s2 + log(s2) + 1.8378770664093453);
    gmu = r / s2;
    gv = -0.5 / s2;
    gs = 0.5 * (r * r + var) / (s2 * s2) - 0.5 / s2;
  } else {
    // Gauss-Hermite, 20 nodes: one node per lane (erfc / erfcx / log / exp in FP64, ~200 instructions each: with lane 0 walking all
    // twenty, this kernel was 10 % of a BASELINE configs[4] evaluation), combined by the fixed butterfly
    const double sgn = 2.0 * y - 1.0, sd = sqrt(2.0 * var);
    double e = 0.0, gm = 0.0, gvv = 0.0;
    if (lane < 20) {
      double lp, ra;
      log_ndtr_and_ratio(sgn * (mu + sd * c_gh_x[lane]), lp, ra);
      e = c_gh_w[lane] * lp;
      gm = c_gh_w[lane] * (sgn * ra);
      gvv = c_gh_w[lane] * (sgn * ra * c_gh_x[lane]);
    }
    e = warp_sum(e);
    gm = warp_sum(gm);
    gvv = warp_sum(gvv);
    const double ISP = 0.5641895835477563;  // 1/sqrt(pi)
    ell = e * ISP;
    gmu = gm * ISP;
    gv = gvv * ISP / sd;  // d sqrt(2 var)/d var = 1/sqrt(2 var)
  }
  if (lane != 0) return;
  double* ro = rowout + (int64_t)b * 4 * ldr;
  ro[n] = lik_scale * ell;
  ro[ldr + n] = lik_scale * gmu;
  ro[2 * ldr + n] = lik_scale * gv;
  ro[3 * ldr + n] = lik_scale * gs;
}

// Predictive marginals only: mean[b][n] = aT[n,:].m ; var[b][n] = sf2 + data_jitter + ||wT[n]||^2 - ||aT[n]||^2 (+ s2)
__global__ void __launch_bounds__(256) k_svgp_marginals(const double* __restrict__ aT, const double* __restrict__ wT, int64_t ld,
                                                        int64_t sC, const double* __restrict__ qm, int64_t sqm,
                                                        const double* __restrict__ theta,
                                                        int d, int M, int nv, double data_jitter, int add_noise,
                                                        double* __restrict__ mean, double* __restrict__ var, int64_t sOut) {
  const int b = blockIdx.y, n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= nv) return;
  qm += b * sqm;
  const double* a = aT + b * sC + (int64_t)n * ld;
  double mu = 0.0, v1 = 0.0, v2 = 0.0;
  for (int j = lane; j < M; j += 32) {
    const double aj = a[j];
    mu = fma(aj, qm[j], mu);
    v2 = fma(aj, aj, v2);
  }
  if (wT) {
    const double* w = wT + b * sC + (int64_t)n * ld;
    for (int j = lane; j < M; j += 32) v1 = fma(w[j], w[j], v1);
  }
  mu = warp_sum(mu);
  v1 = warp_sum(v1);
  v2 = warp_sum(v2);
  if (lane == 0) {
    const double sf2 = theta[(int64_t)b * (d + 2) + d], s2 = theta[(int64_t)b * (d + 2) + d + 1];
    mean[b * sOut + n] = mu;
    var[b * sOut + n] = sf2 + data_jitter + v1 - v2 + (add_noise ? s2 : 0.0);
  }
}

// scal[b][0] += sum_n ell_n ; [1] += sum_n gv_n (direct d/dsf2 through k_nn) ; [2] += sum_n gs_n   (one CTA per batch)
__global__ void __launch_bounds__(256) k_svgp_reduce_rows(const double* __restrict__ rowout, int64_t ldr, int nv,
                                                          double* __restrict__ scal) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* ro = rowout + (int64_t)b * 4 * ldr;
  double e = 0, g = 0, s = 0;
  for (int n = tid; n < nv; n += 256) {
    e += ro[n];
    g += ro[2 * ldr + n];
    s += ro[3 * ldr + n];
  }
  e = block_sum<256>(e, red);
  g = block_sum<256>(g, red);
  s = block_sum<256>(s, red);
  if (tid == 0) {
    scal[b * 4 + 0] += e;
    scal[b * 4 + 1] += g;
    scal[b * 4 + 2] += s;
  }
}

// GAT[n,i] = gmu_n * m_i + 2 gv_n * (SL[n,i] - aT[n,i])     (in place in SL; SL may be NULL => S = 0)
// grid (ceil(Mp/256), nv, batch)
__global__ void k_svgp_gat(double* __restrict__ SL, const double* __restrict__ aT, int64_t ld, int64_t sC,
                           const double* __restrict__ qm, int64_t sqm, const double* __restrict__ rowout, int64_t ldr, int M, int Mp,
                           int has_S) {
  const int i = blockIdx.x * 256 + threadIdx.x, n = blockIdx.y, b = blockIdx.z;
  if (i >= Mp) return;
  qm += b * sqm;
  const double* ro = rowout + (int64_t)b * 4 * ldr;
  const int64_t o = b * sC + (int64_t)n * ld + i;
  double v = 0.0;
  if (i < M) {
    const double s = has_S ? SL[o] : 0.0;
    v = ro[ldr + n] * qm[i] + 2.0 * ro[2 * ldr + n] * (s - aT[o]);
  }
  SL[o] = v;
}

// k_svgp_gat fused with the two transposes that follow it: one pass over aT (and SL) writes
//   SL[n,i]  = GAT[n,i] = gmu_n m_i + 2 gv_n (SL[n,i] - aT[n,i])    (row-major, A operand of dKc = GAT L^-1)
//   tA[i,n]  = aT[n,i]                                             (k-contiguous operands of Gbar += GAT^T aT)
//   tB[i,n]  = GAT[n,i]
// 4 array passes instead of 6 (gat: 2, two k_transpose_rect: 4).  32 x 32 tiles; columns nv <= n < ldo of tA / tB are zero.
// grid (ceil(ldo/32), Mp/32, batch), block (32, 8)
__global__ void __launch_bounds__(256) k_svgp_gat_t(double* __restrict__ SL, const double* __restrict__ aT, int64_t ld, int64_t sC,
                                                    const double* __restrict__ qm, int64_t sqm, const double* __restrict__ rowout,
                                                    int64_t ldr, int M, int Mp, int nv, int has_S, double* __restrict__ tA,
                                                    double* __restrict__ tB, int64_t ldo) {
  __shared__ double ta[32][33], tg[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  qm += b * sqm;
  const double* ro = rowout + (int64_t)b * 4 * ldr;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, i = i0 + threadIdx.x;
    double a = 0.0, v = 0.0;
    if (n < nv && i < Mp) {
      const int64_t o = b * sC + (int64_t)n * ld + i;
      a = aT[o];
      if (i < M) v = ro[ldr + n] * qm[i] + 2.0 * ro[2 * ldr + n] * ((has_S ? SL[o] : 0.0) - a);
      SL[o] = v;
    }
    ta[r][threadIdx.x] = a;
    tg[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, n = n0 + threadIdx.x;
    if (i < Mp && n < ldo) {
      tA[b * sC + (int64_t)i * ldo + n] = ta[threadIdx.x][r];
      tB[b * sC + (int64_t)i * ldo + n] = tg[threadIdx.x][r];
    }
  }
}

// dm[b][i] += sum_n aT[n,i] * gmu_n      grid (ceil(M/32), batch), block 256 = 32 consecutive columns x 8 row lanes
// Row lane r sums the rows n = r, r + 8, ... of its column (coalesced 256-byte row segments, four independent partial sums in
// flight), the 8 lanes are combined in a fixed order: deterministic.  (One thread per column walking all nv rows -- the first
// version -- was a chain of nv dependent strided loads on 2 CTAs per chain: 1 ms per launch, 32 % of a BASELINE configs[4]
// evaluation at 8 chains per GPU.)
__global__ void __launch_bounds__(256) k_svgp_dm(const double* __restrict__ aT, int64_t ld, int64_t sC, const double* __restrict__ rowout,
                                                 int64_t ldr, int M, int nv, double* __restrict__ dm, int64_t sdm) {
  __shared__ double red[8][33];
  const int li = threadIdx.x & 31, r = threadIdx.x >> 5, i = blockIdx.x * 32 + li, b = blockIdx.y;
  const double* ro = rowout + (int64_t)b * 4 * ldr + ldr;
  const double* a = aT + b * sC + i;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (i < M) {
    int n = r;
    for (; n + 24 < nv; n += 32) {
      s0 = fma(a[(int64_t)n * ld], ro[n], s0);
      s1 = fma(a[(int64_t)(n + 8) * ld], ro[n + 8], s1);
      s2 = fma(a[(int64_t)(n + 16) * ld], ro[n + 16], s2);
      s3 = fma(a[(int64_t)(n + 24) * ld], ro[n + 24], s3);
    }
    for (; n < nv; n += 8) s0 = fma(a[(int64_t)n * ld], ro[n], s0);
  }
  red[r][li] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (r == 0 && i < M) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][li];
    dm[b * sdm + i] += s;
  }
}

// out[b][i][n] = in[b][n][i] * (rowscale ? rowscale[b][n] : 1) * (mul ? mul[b][n][i] : 1)   for i < rows_out(Mp), n < nv; zero for nv <= n < ldo
// 32x32 tiles.  grid (ceil(ldo/32), ceil(Mp/32), batch), block (32, 8)
__global__ void k_transpose_rect(const double* __restrict__ in, const double* __restrict__ mul, int64_t ldi, int64_t sIn,
                                 const double* __restrict__ rowscale, int64_t sRS, int nv, int Mp, double* __restrict__ out,
                                 int64_t ldo, int64_t sOut) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, i = i0 + threadIdx.x;
    double v = 0.0;
    if (n < nv && i < Mp) {
      v = in[b * sIn + (int64_t)n * ldi + i];
      if (mul) v *= mul[b * sIn + (int64_t)n * ldi + i];
      if (rowscale) v *= rowscale[b * sRS + n];
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, n = n0 + threadIdx.x;
    if (i < Mp && n < ldo) out[b * sOut + (int64_t)i * ldo + n] = tile[threadIdx.x][r];
  }
}

// sum over (n < nv, i < M) of A[b][n][i] * B[b][n][i] in two deterministic stages (the k-weighted total sum(dKc o Kc) of dF/dsf2 when the
// moments are weighted by dk/d(d2)): part[b][blk] over contiguous row slices, then scal[b][3] += sum_blk part[b][blk] in order
__global__ void __launch_bounds__(256) k_sum_prod_partial(const double* __restrict__ A, const double* __restrict__ B, int64_t ld, int64_t sC,
                                                          int nv, int M, double* __restrict__ part, int nblk) {
  __shared__ double red[8];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int rows = (nv + nblk - 1) / nblk, r0 = blockIdx.x * rows, r1 = min(nv, r0 + rows);
  double acc = 0.0;
  for (int n = r0; n < r1; ++n)
    for (int i = tid; i < M; i += 256) acc = fma(A[b * sC + (int64_t)n * ld + i], B[b * sC + (int64_t)n * ld + i], acc);
  acc = block_sum<256>(acc, red);
  if (tid == 0) part[(int64_t)b * nblk + blockIdx.x] = acc;
}
__global__ void k_sum_prod_final(const double* __restrict__ part, int nblk, double* __restrict__ scal) {
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int k = 0; k < nblk; ++k) acc += part[(int64_t)b * nblk + k];
  scal[b * 4 + 3] += acc;
}

// PhiT[q][n] : q = 0 -> 1 ; 1..d -> x_n[q-1] ; d+1..2d -> x_n[q-1-d]^2 ; zero for n >= nv.   grid (ceil(ldo/256), nq)
__global__ void k_phiT(const double* __restrict__ X, int nv, int d, double* __restrict__ PhiT, int64_t ldo) {
  const int n = blockIdx.x * 256 + threadIdx.x, q = blockIdx.y;
  if (n >= ldo) return;
  double v = 0.0;
  if (n < nv) {
    if (q == 0) v = 1.0;
    else if (q <= d) v = X[(int64_t)n * d + q - 1];
    else { const double x = X[(int64_t)n * d + q - 1 - d]; v = x * x; }
  }
  PhiT[(int64_t)q * ldo + n] = v;
}

// H = sym(Phi(Gbar)):  H[i][j] = Gbar[max(i,j)][min(i,j)] / 2   (zero on the padding)
__global__ void k_sym_phi(const double* __restrict__ G, double* __restrict__ H, int M, int Mp, int64_t sM) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x, b = blockIdx.z;
  if (i >= Mp || j >= Mp) return;
  double v = 0.0;
  if (i < M && j < M) v = 0.5 * G[b * sM + (int64_t)(i > j ? i : j) * Mp + (i > j ? j : i)];
  H[b * sM + (int64_t)i * Mp + j] = v;
}

// final assembly of the SVGP value and gradient (one CTA per batch element).
//   elbo = scal[0] - kl_scale * KL ;  KL = 1/2 [ ||Ls||_F^2 + m.m - M - 2 sum log|Ls_ii| ]
//   grad = [ g_kzx[0..d) + rowacc-sum , g_kzx[d] + rowsum_k/sf2 + scal[1] , scal[2] , dZ_kzx + dZ_kzz , dm - kl_scale m ,
//            2 tril(dLsraw) - kl_scale (Ls - diag(1/Ls_ii)) ]
__global__ void __launch_bounds__(256) k_svgp_final(const double* __restrict__ scal, const double* __restrict__ gkzx, int64_t sGk,
                                                    const double* __restrict__ rowacc, const double* __restrict__ dZzz,
                                                    const double* __restrict__ dm, int64_t sdm,
                                                    const double* __restrict__ dLsraw, int64_t sM, int Mp,
                                                    const double* __restrict__ qm, int64_t sqm, const double* __restrict__ Ls,
                                                    const double* __restrict__ theta, int M, int d, double kl_scale,
                                                    int need_grad, double* __restrict__ elbo, double* __restrict__ grad,
                                                    int64_t sG) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  qm += b * sqm;
  double kl = 0.0;
  if (Ls && kl_scale != 0.0) {
    double fro = 0.0, mm = 0.0, ld = 0.0;
    for (int64_t e = tid; e < (int64_t)M * M; e += 256) {
      const int i = (int)(e / M), j = (int)(e % M);
      if (j <= i) fro = fma(Ls[e], Ls[e], fro);
    }
    for (int i = tid; i < M; i += 256) {
      mm = fma(qm[i], qm[i], mm);
      ld += log(fabs(Ls[(int64_t)i * M + i]));
    }
    fro = block_sum<256>(fro, red);
    mm = block_sum<256>(mm, red);
    ld = block_sum<256>(ld, red);
    kl = 0.5 * (fro + mm - (double)M - 2.0 * ld);
  }
  if (tid == 0) elbo[b] = scal[b * 4 + 0] - kl_scale * kl;
  if (!need_grad) return;
  const double* th = theta + (int64_t)b * (d + 2);
  double* g = grad + b * sG;
  for (int c = 0; c <= d; ++c) {
    double s = 0.0;
    for (int i = tid; i < M; i += 256) s += rowacc[((int64_t)b * M + i) * (d + 1) + c];
    s = block_sum<256>(s, red);
    if (tid == 0) {
      if (c < d) g[c] = gkzx[b * sGk + c] + s;
      else g[d] = gkzx[b * sGk + d] + s / th[d] + scal[b * 4 + 1];
    }
  }
  if (tid == 0) g[d + 1] = scal[b * 4 + 2];
  for (int64_t e = tid; e < (int64_t)M * d; e += 256) g[d + 2 + e] = gkzx[b * sGk + d + 2 + e] + dZzz[b * sGk + d + 2 + e];
  double* gm = g + d + 2 + (int64_t)M * d;
  for (int i = tid; i < M; i += 256) gm[i] = dm[b * sdm + i] - kl_scale * qm[i];
  double* gL = gm + M;
  for (int64_t e = tid; e < (int64_t)M * M; e += 256) {
    const int i = (int)(e / M), j = (int)(e % M);
    double v = 0.0;
    if (Ls && j <= i) {
      v = 2.0 * dLsraw[b * sM + (int64_t)i * Mp + j] - kl_scale * (Ls[e] - ((i == j) ? 1.0 / Ls[e] : 0.0));
    }
    gL[e] = v;
  }
}

}  // namespace ggp
