// Composite covariance functions: sums of scaled products of stationary factors,
//     k(x, z) = sum_t a_t prod_f phi_tf(x, z),
// the structure of the reference's CO2 model (experiments/co2_bayesian_sgpr_hmc.py:74-83 in gpytorch, :107-149 in pymc3):
//     ScaleKernel(Periodic * RBF) + ScaleKernel(RBF) + ScaleKernel(RQ) + ScaleKernel(RBF | Matern32).
// Factors (every one has ARD lengthscales ell[d]; phi(x, x) = 1, so k(x, x) = sum_t a_t):
//     RBF       exp(-d2 / 2),  d2 = sum_c ((x_c - z_c) / ell_c)^2
//     MATERN32  (1 + sqrt(3) r) exp(-sqrt(3) r),  r = sqrt(d2)
//     MATERN52  (1 + sqrt(5) r + 5 d2 / 3) exp(-sqrt(5) r)
//     RQ        (1 + d2 / (2 alpha))^(-alpha)                          extra parameter: alpha                  (RQKernel / RatQuad)
//     PERIODIC  exp(-2 sum_c sin^2(pi (x_c - z_c) / p_c) / ell_c^2)     extra parameters: p[d]  (MacKay's form; gpytorch's
//               PeriodicKernel is this with lengthscale = ell^2, pymc3's Periodic with ls = ell / 2: the host maps)
// Parameter row (one per theta draw), in program order: for each term  a_t, then for each factor  ell[d], then its extra parameters.
//
// The streamed FP64 path evaluates the tile k(X, Z) element by element (k_build_kc_prog) and, instead of the moment trick of the
// single-kernel paths (which needs dk/d(d2) to be ONE weight matrix), contracts the stored dF/dKzx chunk with dk/d(parameter) and
// dk/dz directly (k_kprog_grad): O(N M P) FP64 work next to the O(N M^2) GEMMs.
#pragma once
#include "common.cuh"
#include "../../include/ggp_b200.h"

namespace ggp {

constexpr int KP_MAXT = GGP_KPROG_MAX_TERMS, KP_MAXF = GGP_KPROG_MAX_FACTORS, KP_MAXP = GGP_KPROG_MAX_PARAMS, KP_MAXD = 16;
enum { KF_RBF = GGP_KERNEL_RBF, KF_M32 = GGP_KERNEL_MATERN32, KF_M52 = GGP_KERNEL_MATERN52, KF_RQ = GGP_KERNEL_RQ, KF_PER = GGP_KERNEL_PERIODIC };

struct KProgDev {
  int nterms, P, d;
  int nfac[KP_MAXT];
  int kind[KP_MAXT][KP_MAXF];
  int amp[KP_MAXT];             // index of a_t in the parameter row
  int off[KP_MAXT][KP_MAXF];    // index of the factor's ell[0]; extras follow at off + d
};

// parameter count of one factor
__host__ __device__ inline int kfac_nparams(int kind, int d) { return kind == KF_RQ ? d + 1 : (kind == KF_PER ? 2 * d : d); }

// host: offsets from the public description; returns false if it is malformed
inline bool kprog_compile(const ggp_kprog* pg, int d, KProgDev* out) {
  if (!pg || pg->nterms < 1 || pg->nterms > KP_MAXT || d < 1 || d > KP_MAXD) return false;
  out->nterms = pg->nterms;
  out->d = d;
  int P = 0;
  for (int t = 0; t < KP_MAXT; ++t) {
    out->nfac[t] = 0;
    out->amp[t] = 0;
    for (int f = 0; f < KP_MAXF; ++f) out->kind[t][f] = out->off[t][f] = 0;
  }
  for (int t = 0; t < pg->nterms; ++t) {
    const int nf = pg->nfactors[t];
    if (nf < 1 || nf > KP_MAXF) return false;
    out->nfac[t] = nf;
    out->amp[t] = P++;
    for (int f = 0; f < nf; ++f) {
      const int k = pg->kind[t][f];
      if (k < KF_RBF || k > KF_PER) return false;
      out->kind[t][f] = k;
      out->off[t][f] = P;
      P += kfac_nparams(k, d);
    }
  }
  if (P > KP_MAXP) return false;
  out->P = P;
  return true;
}

// one factor's value at (x, z); p = its parameters (ell[d], extras)
__device__ __forceinline__ double kfac_val(int kind, const double* __restrict__ p, const double* __restrict__ x, const double* __restrict__ z, int d) {
  if (kind == KF_PER) {
    double S = 0.0;
    for (int c = 0; c < d; ++c) {
      const double s = sinpi((x[c] - z[c]) / p[d + c]) / p[c];
      S = fma(s, s, S);
    }
    return exp(-2.0 * S);
  }
  double d2 = 0.0;
  for (int c = 0; c < d; ++c) {
    const double t = (x[c] - z[c]) / p[c];
    d2 = fma(t, t, d2);
  }
  if (kind == KF_RBF) return exp(-0.5 * d2);
  if (kind == KF_RQ) return exp(-p[d] * log1p(d2 / (2.0 * p[d])));
  const double r = sqrt(d2);
  if (kind == KF_M32) return (1.0 + 1.7320508075688772 * r) * exp(-1.7320508075688772 * r);
  return (1.0 + 2.23606797749979 * r + (5.0 / 3.0) * d2) * exp(-2.23606797749979 * r);
}

__device__ __forceinline__ double kprog_eval(const KProgDev& pg, const double* __restrict__ th, const double* __restrict__ x,
                                             const double* __restrict__ z) {
  double k = 0.0;
  for (int t = 0; t < pg.nterms; ++t) {
    double v = th[pg.amp[t]];
    for (int f = 0; f < pg.nfac[t]; ++f) v *= kfac_val(pg.kind[t][f], th + pg.off[t][f], x, z, pg.d);
    k += v;
  }
  return k;
}
__device__ __forceinline__ double kprog_diag(const KProgDev& pg, const double* __restrict__ th) {
  double k = 0.0;
  for (int t = 0; t < pg.nterms; ++t) k += th[pg.amp[t]];
  return k;
}

// accp[...] += coef * d phi / d(parameters),  accz[c] += coef * d phi / d z_c   (z = the SECOND argument)
__device__ __forceinline__ void kfac_grad_acc(int kind, const double* __restrict__ p, const double* __restrict__ x, const double* __restrict__ z,
                                              int d, double coef, double* __restrict__ accp, double* __restrict__ accz) {
  if (kind == KF_PER) {
    double S = 0.0;
    for (int c = 0; c < d; ++c) {
      const double s = sinpi((x[c] - z[c]) / p[d + c]) / p[c];
      S = fma(s, s, S);
    }
    const double cp = coef * exp(-2.0 * S);
    const double PI = 3.141592653589793;
    for (int c = 0; c < d; ++c) {
      const double df = x[c] - z[c], ell = p[c], per = p[d + c];
      double s, co;
      sincospi(df / per, &s, &co);
      const double sc = 4.0 * s * co / (ell * ell);                 // 2 sin(2 pi df / p) / ell^2
      accp[c] += cp * 4.0 * s * s / (ell * ell * ell);
      accp[d + c] += cp * sc * PI * df / (per * per);
      accz[c] += cp * sc * PI / per;
    }
    return;
  }
  double d2 = 0.0;
  for (int c = 0; c < d; ++c) {
    const double t = (x[c] - z[c]) / p[c];
    d2 = fma(t, t, d2);
  }
  double g;   // d phi / d(d2)
  if (kind == KF_RBF) {
    g = -0.5 * exp(-0.5 * d2);
  } else if (kind == KF_RQ) {
    const double al = p[d], q = d2 / (2.0 * al), l1 = log1p(q);
    g = -0.5 * exp(-(al + 1.0) * l1);
    accp[d] += coef * exp(-al * l1) * (q / (1.0 + q) - l1);         // d/d alpha of exp(-alpha log1p(d2 / (2 alpha)))
  } else if (kind == KF_M32) {
    g = -1.5 * exp(-1.7320508075688772 * sqrt(d2));
  } else {
    const double r = sqrt(d2);
    g = -(5.0 / 6.0) * (1.0 + 2.23606797749979 * r) * exp(-2.23606797749979 * r);
  }
  const double cg = coef * g;
  for (int c = 0; c < d; ++c) {
    const double df = x[c] - z[c], ell = p[c];
    accp[c] += cg * (-2.0 * df * df / (ell * ell * ell));
    accz[c] += cg * (-2.0 * df / (ell * ell));
  }
}

// acc[0 .. P) += w dk/d(parameter),  acc[P .. P + d) += w dk/dz
__device__ __forceinline__ void kprog_grad_acc(const KProgDev& pg, const double* __restrict__ th, const double* __restrict__ x,
                                               const double* __restrict__ z, double w, double* __restrict__ acc) {
  const int d = pg.d;
  for (int t = 0; t < pg.nterms; ++t) {
    double phi[KP_MAXF], all = 1.0;
    const int nf = pg.nfac[t];
    for (int f = 0; f < nf; ++f) {
      phi[f] = kfac_val(pg.kind[t][f], th + pg.off[t][f], x, z, d);
      all *= phi[f];
    }
    acc[pg.amp[t]] += w * all;
    const double wa = w * th[pg.amp[t]];
    for (int f = 0; f < nf; ++f) {
      double others = 1.0;
      for (int g = 0; g < nf; ++g)
        if (g != f) others *= phi[g];
      kfac_grad_acc(pg.kind[t][f], th + pg.off[t][f], x, z, d, wa * others, acc + pg.off[t][f], acc + pg.P);
    }
  }
}

// Kzz[b][i][j] = k(z_i, z_j) + jitter_b delta_ij, identity on the padding (k_build_kzz for a program).  grid (Mp/16, Mp/16, batch)
__global__ void k_build_kzz_prog(const double* __restrict__ Z, int M, int Mp, KProgDev pg, const double* __restrict__ kth,
                                 const double* __restrict__ jitter, double* __restrict__ Kzz, int64_t sK, double* __restrict__ piv_tol) {
  const int b = blockIdx.z;
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= Mp || j >= Mp) return;
  const double* th = kth + (int64_t)b * pg.P;
  if (i == 0 && j == 0 && piv_tol) piv_tol[b] = 1e-12 /* GGP_PIVOT_RTOL */ * (kprog_diag(pg, th) + (jitter ? jitter[b] : 0.0));
  double v;
  if (i < M && j < M) {
    v = kprog_eval(pg, th, Z + (int64_t)i * pg.d, Z + (int64_t)j * pg.d);
    if (i == j) v = kprog_diag(pg, th) + (jitter ? jitter[b] : 0.0);
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  Kzz[(int64_t)b * sK + (int64_t)i * Mp + j] = v;
}

// dst[b][n][j] = k(x_n, z_j) for n < nv, j < M; 0 on the padding columns (k_build_kc for a program).  grid (Mp/32, ceil(nv/8), batch)
__global__ void __launch_bounds__(256) k_build_kc_prog(const double* __restrict__ X, int nv, const double* __restrict__ Z, int M, int Mp,
                                                       KProgDev pg, const double* __restrict__ kth, double* __restrict__ dst, int64_t ld,
                                                       int64_t sK) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31), n = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (n >= nv || j >= Mp) return;
  const double* th = kth + (int64_t)b * pg.P;
  dst[(int64_t)b * sK + (int64_t)n * ld + j] = (j < M) ? kprog_eval(pg, th, X + (int64_t)n * pg.d, Z + (int64_t)j * pg.d) : 0.0;
}

// One warp per row i (an inducing input z_i as the SECOND argument of k), lanes over the columns n:
//   w_in = W[b][i][n] + (u ? u[b][i] y[n] : 0)
//   rowacc[b][i][0 .. P)     (+)= sum_n w_in dk(x_n, z_i)/d(parameter)
//   rowacc[b][i][P .. P + d) (+)= sum_n w_in dk(x_n, z_i)/dz_i
// X1 = the rows x_n (the data chunk for dF/dKzx; Z itself for the symmetric dF/dKzz, whose z-gradient the caller scales by 2: z_i is
// also the first argument of column i).  grid (ceil(M/8), batch), block 256.
__global__ void __launch_bounds__(256) k_kprog_grad(const double* __restrict__ W, int64_t ldw, int64_t sW, const double* __restrict__ u,
                                                    int64_t su, const double* __restrict__ y, const double* __restrict__ X1, int nv,
                                                    const double* __restrict__ Z, int M, KProgDev pg, const double* __restrict__ kth,
                                                    double* __restrict__ rowacc, int accumulate) {
  const int b = blockIdx.y, i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= M) return;
  const double* th = kth + (int64_t)b * pg.P;
  const int d = pg.d, na = pg.P + d;
  double acc[KP_MAXP + KP_MAXD];
  for (int k = 0; k < na; ++k) acc[k] = 0.0;
  const double* wrow = W + b * sW + (int64_t)i * ldw;
  const double ui = u ? u[b * su + i] : 0.0;
  const double* zi = Z + (int64_t)i * d;
  for (int n = lane; n < nv; n += 32) {
    const double w = wrow[n] + (u ? ui * y[n] : 0.0);
    kprog_grad_acc(pg, th, X1 + (int64_t)n * d, zi, w, acc);
  }
  double* out = rowacc + ((int64_t)b * M + i) * na;
  for (int k = 0; k < na; ++k) {
    const double s = warp_sum(acc[k]);
    if (lane == 0) out[k] = accumulate ? out[k] + s : s;
  }
}

// kgrad[b][p] = sum_i rowacc[b][i][p]  (+ the explicit k(x_n, x_n) = sum_t a_t dependence of the bound, -N / (2 s2) per amplitude, when
// n_total is given);  grad[b] = [0 (d + 1 unused kernel slots), ds2 ? ds2[b] : 0, dZ = zscale * rowacc[b][i][P + c]]
__global__ void __launch_bounds__(256) k_kprog_grad_final(const double* __restrict__ rowacc, int M, KProgDev pg, double zscale,
                                                          const double* __restrict__ n_total, int64_t sN, const double* __restrict__ theta,
                                                          const double* __restrict__ ds2, double* __restrict__ kgrad,
                                                          double* __restrict__ grad, int64_t sG) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x, d = pg.d, na = pg.P + d;
  for (int p = 0; p < pg.P; ++p) {
    double s = 0.0;
    for (int i = tid; i < M; i += 256) s += rowacc[((int64_t)b * M + i) * na + p];
    s = block_sum<256>(s, red);
    if (tid == 0) {
      bool is_amp = false;
      for (int t = 0; t < pg.nterms; ++t) is_amp = is_amp || pg.amp[t] == p;
      if (is_amp && n_total) s -= 0.5 * n_total[b * sN] / theta[(int64_t)b * (d + 2) + d + 1];
      kgrad[(int64_t)b * pg.P + p] = s;
    }
    __syncthreads();
  }
  for (int c = tid; c < d + 2; c += 256) grad[b * sG + c] = (c == d + 1 && ds2) ? ds2[b] : 0.0;
  for (int e = tid; e < M * d; e += 256) grad[b * sG + d + 2 + e] = zscale * rowacc[((int64_t)b * M + e / d) * na + pg.P + e % d];
}

}  // namespace ggp
