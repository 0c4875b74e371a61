/*
 * sgpr_hp.c -- EXTENDED-PRECISION restatement of the collapsed SGPR bound and its closed-form gradient.
 * TEST INFRASTRUCTURE ONLY (oracle/): the product path never links or calls this.
 *
 * Why it exists: dF/d(ell, sf2, Z) go through Kzz^{-1}; at cond(Kzz) ~ 1e8 (the headline configuration: M = 1024 inducing rows
 * drawn from the data, jitter 0 .. 1e-6) two correct float64 evaluations of the same formula differ by 1e-9 .. 1e-7.  A float64
 * oracle therefore cannot decide whether a GPU path that differs from it by 3e-8 is wrong.  This file evaluates the same algebra
 * in x87 `long double` (64-bit mantissa, eps 1.1e-19; default) or in IEEE binary128 (`-DUSE_QUAD`, libquadmath, eps 1.9e-34), so
 * that the float64 oracle AND the GPU paths can be measured against a reference whose own rounding error is 3 to 18 orders smaller.
 *
 * What it restates (reference call sites; the arithmetic itself lives in gpytorch / pymc3):
 *   models/sgpr.py:123-129        -mll(output, y) over InducingPointKernel(ScaleKernel(RBFKernel(ard)))   (bound F, A.4)
 *   models/sgpr.py:129            loss.backward()                                                         (gradient)
 *   models/bayesian_sgpr_hmc.py:60-71  pm.gp.MarginalSparse(approx="VFE").marginal_likelihood            (same F, jitter 1e-6)
 * Formulas: identical to oracle/sgpr.py::sgpr_grads_closed_form; the kernel derivative sums are taken directly as
 * sum W (z - x)^2 and sum W (z - x) (no moment expansion), which is the mathematically identical, cancellation-free form.
 *
 * Inputs are float64 (exactly what the GPU path and the float64 oracle receive); the jitter is an input (the ladder level the
 * float64 oracle settled on), a constant w.r.t. the gradient as upstream.
 *
 * Build: see oracle/hp/Makefile (gcc -O2 -fopenmp -shared -fPIC).  Outputs go to oracle/_ref/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef USE_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_EXP expq
#define R_SQRT sqrtq
#define R_LOG logq
#define R_PI M_PIq
#else
typedef long double real;
#define R_EXP expl
#define R_SQRT sqrtl
#define R_LOG logl
#define R_PI 3.141592653589793238462643383279502884L
#endif

static real dot_nt(const real* a, const real* b, long k) { /* sum_k a[k] b[k], four independent accumulators */
  real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  long i = 0;
  for (; i + 4 <= k; i += 4) {
    s0 += a[i] * b[i];
    s1 += a[i + 1] * b[i + 1];
    s2 += a[i + 2] * b[i + 2];
    s3 += a[i + 3] * b[i + 3];
  }
  for (; i < k; ++i) s0 += a[i] * b[i];
  return (s0 + s1) + (s2 + s3);
}

/* in-place lower Cholesky of a[m][m] (row-major); returns 0 or the 1-based index of the first non-positive pivot */
static int chol_lower(real* a, int m) {
  for (int j = 0; j < m; ++j) {
    real d = a[(long)j * m + j] - dot_nt(a + (long)j * m, a + (long)j * m, j);
    if (!(d > 0)) return j + 1;
    d = R_SQRT(d);
    a[(long)j * m + j] = d;
#pragma omp parallel for schedule(static)
    for (int i = j + 1; i < m; ++i) a[(long)i * m + j] = (a[(long)i * m + j] - dot_nt(a + (long)i * m, a + (long)j * m, j)) / d;
  }
  for (int i = 0; i < m; ++i)
    for (int j = i + 1; j < m; ++j) a[(long)i * m + j] = 0;
  return 0;
}

/* inverse of a lower-triangular matrix, by columns (forward substitution); linvT receives the transpose */
static void tri_inverse(const real* L, real* Linv, real* LinvT, int m) {
  memset(Linv, 0, sizeof(real) * (size_t)m * m);
#pragma omp parallel for schedule(dynamic, 8)
  for (int c = 0; c < m; ++c) { /* column c of the inverse lives in row c of LinvT (contiguous) */
    real* x = LinvT + (long)c * m;
    for (int i = 0; i < c; ++i) x[i] = 0;
    x[c] = (real)1 / L[(long)c * m + c];
    for (int i = c + 1; i < m; ++i) x[i] = -dot_nt(L + (long)i * m + c, x + c, i - c) / L[(long)i * m + i];
  }
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) Linv[(long)i * m + j] = LinvT[(long)j * m + i];
}

/* C[m][m] = A[m][m] * B[m][m]^T  (all row-major) */
static void mm_nt(const real* A, const real* B, real* C, int m) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) C[(long)i * m + j] = dot_nt(A + (long)i * m, B + (long)j * m, m);
}

static void kernel_rows(const double* X, long n0, int nc, const double* Z, int m, int d, const real* ell, real sf2, real* Kc) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < nc; ++n)
    for (int j = 0; j < m; ++j) {
      real d2 = 0;
      for (int k = 0; k < d; ++k) {
        const real t = ((real)X[(n0 + n) * d + k] - (real)Z[(long)j * d + k]) / ell[k];
        d2 += t * t;
      }
      Kc[(long)n * m + j] = sf2 * R_EXP(-d2 / 2);
    }
}

/* out[0] = F (not divided by N); out[1 .. d] = dF/d ell; out[d+1] = dF/d sf2; out[d+2] = dF/d s2; out[d+3 ..] = dF/dZ [m][d]
 * returns 0; 1000 + k if Kzz + jitter I is not PD at pivot k; 2000 + k if I + A A^T / s is not. */
int sgpr_hp_bound_grad(const double* X, const double* y, const double* Z, const double* theta, double jitter, long n, int m, int d,
                       int nthreads, double* out) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const size_t MM = (size_t)m * m;
  const int NC = 512;
  real* ell = (real*)malloc(sizeof(real) * d);
  for (int k = 0; k < d; ++k) ell[k] = (real)theta[k];
  const real sf2 = (real)theta[d], s2 = (real)theta[d + 1];
  real *Kzz = malloc(sizeof(real) * MM), *L = malloc(sizeof(real) * MM), *Linv = malloc(sizeof(real) * MM),
       *LinvT = malloc(sizeof(real) * MM), *S = calloc(MM, sizeof(real)), *B = malloc(sizeof(real) * MM),
       *LBinv = malloc(sizeof(real) * MM), *LBinvT = malloc(sizeof(real) * MM), *Binv = malloc(sizeof(real) * MM),
       *PA = malloc(sizeof(real) * MM), *T1 = malloc(sizeof(real) * MM), *P = malloc(sizeof(real) * MM),
       *Gzz = malloc(sizeof(real) * MM);
  real *Kc = malloc(sizeof(real) * (size_t)NC * m), *A = malloc(sizeof(real) * (size_t)NC * m);
  real *b = calloc(m, sizeof(real)), *c = malloc(sizeof(real) * m), *beta = malloc(sizeof(real) * m), *u = malloc(sizeof(real) * m);
  real *gl = calloc((size_t)m * d, sizeof(real)), *gz = calloc((size_t)m * d, sizeof(real)), *r = calloc(m, sizeof(real));
  int rc = 0;

  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      real d2 = 0;
      for (int k = 0; k < d; ++k) {
        const real t = ((real)Z[(long)i * d + k] - (real)Z[(long)j * d + k]) / ell[k];
        d2 += t * t;
      }
      Kzz[(long)i * m + j] = sf2 * R_EXP(-d2 / 2);
      L[(long)i * m + j] = Kzz[(long)i * m + j] + (i == j ? (real)jitter : (real)0);
    }
  int info = chol_lower(L, m);
  if (info) { rc = 1000 + info; goto done; }
  tri_inverse(L, Linv, LinvT, m);

  /* pass 1: A = L^{-1} Kzx by row chunks;  S += A A^T, b += A y */
  real yty = 0;
  for (long n0 = 0; n0 < n; n0 += NC) {
    const int nc = (int)((n - n0) < NC ? (n - n0) : NC);
    kernel_rows(X, n0, nc, Z, m, d, ell, sf2, Kc);
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < m; ++i) {
      real bi = 0;
      for (int q = 0; q < nc; ++q) {
        const real a = dot_nt(Linv + (long)i * m, Kc + (long)q * m, i + 1);
        A[(long)i * NC + q] = a;
        bi += a * (real)y[n0 + q];
      }
      b[i] += bi;
    }
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < m; ++i)
      for (int j = 0; j <= i; ++j) S[(long)i * m + j] += dot_nt(A + (long)i * NC, A + (long)j * NC, nc);
    for (int q = 0; q < nc; ++q) yty += (real)y[n0 + q] * (real)y[n0 + q];
  }
  for (int i = 0; i < m; ++i)
    for (int j = i + 1; j < m; ++j) S[(long)i * m + j] = S[(long)j * m + i];

  for (size_t k = 0; k < MM; ++k) B[k] = S[k] / s2;
  for (int i = 0; i < m; ++i) B[(long)i * m + i] += 1;
  memcpy(T1, B, sizeof(real) * MM); /* T1 = L_B */
  info = chol_lower(T1, m);
  if (info) { rc = 2000 + info; goto done; }
  real logdet = 0, trS = 0;
  for (int i = 0; i < m; ++i) { logdet += R_LOG(T1[(long)i * m + i]); trS += S[(long)i * m + i]; }
  tri_inverse(T1, LBinv, LBinvT, m);
  for (int i = 0; i < m; ++i) c[i] = dot_nt(LBinv + (long)i * m, b, i + 1) / s2;
  real cc = 0;
  for (int i = 0; i < m; ++i) cc += c[i] * c[i];
  const real N = (real)n;
  const real F = -N * R_LOG(2 * R_PI) / 2 - N * R_LOG(s2) / 2 - logdet - (yty / s2 - cc) / 2 - (N * sf2 - trS) / (2 * s2);
  out[0] = (double)F;

  mm_nt(LBinvT, LBinvT, Binv, m); /* B^{-1} = L_B^{-T} L_B^{-1} */
  for (int i = 0; i < m; ++i) beta[i] = dot_nt(Binv + (long)i * m, b, m);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j)
      PA[(long)i * m + j] = ((i == j ? (real)1 : (real)0) - Binv[(long)i * m + j]) / s2 - beta[i] * beta[j] / (s2 * s2 * s2);
  mm_nt(LinvT, PA, T1, m);   /* T1 = L^{-T} P_A   (P_A symmetric) */
  mm_nt(T1, LinvT, P, m);    /* P  = T1 L^{-1} */
  for (int i = 0; i < m; ++i) u[i] = dot_nt(LinvT + (long)i * m, beta, m) / (s2 * s2);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j)
      PA[(long)i * m + j] = B[(long)i * m + j] + Binv[(long)i * m + j] - (i == j ? (real)2 : (real)0) + beta[i] * beta[j] / (s2 * s2);
  mm_nt(LinvT, PA, T1, m);
  mm_nt(T1, LinvT, Gzz, m);
  for (size_t k = 0; k < MM; ++k) Gzz[k] = -Gzz[k] / 2;

  /* pass 2: W = (P Kzx + u y^T) o Kzx against (z - x)^2, (z - x), 1 */
  for (long n0 = 0; n0 < n; n0 += NC) {
    const int nc = (int)((n - n0) < NC ? (n - n0) : NC);
    kernel_rows(X, n0, nc, Z, m, d, ell, sf2, Kc);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
      for (int q = 0; q < nc; ++q) {
        const real g = dot_nt(P + (long)i * m, Kc + (long)q * m, m) + u[i] * (real)y[n0 + q];
        const real w = g * Kc[(long)q * m + i];
        r[i] += w;
        for (int k = 0; k < d; ++k) {
          const real dz = (real)Z[(long)i * d + k] - (real)X[(n0 + q) * d + k];
          gl[(long)i * d + k] += w * dz * dz;
          gz[(long)i * d + k] += w * dz;
        }
      }
    }
  }
  /* Kzz part: V = Gzz o Kzz (no jitter: it is a constant) */
  real rsum = 0;
  real* dl = calloc(d, sizeof(real));
  for (int i = 0; i < m; ++i) {
    rsum += r[i];
    for (int j = 0; j < m; ++j) {
      const real v = Gzz[(long)i * m + j] * Kzz[(long)i * m + j];
      rsum += v;
      for (int k = 0; k < d; ++k) {
        const real dz = (real)Z[(long)i * d + k] - (real)Z[(long)j * d + k];
        gl[(long)i * d + k] += v * dz * dz;
        gz[(long)i * d + k] += 2 * v * dz;
      }
    }
  }
  for (int i = 0; i < m; ++i)
    for (int k = 0; k < d; ++k) dl[k] += gl[(long)i * d + k];
  for (int k = 0; k < d; ++k) out[1 + k] = (double)(dl[k] / (ell[k] * ell[k] * ell[k]));
  out[1 + d] = (double)(rsum / sf2 - N / (2 * s2));
  real trBS = 0, bbeta = 0, bSb = 0;
  for (int i = 0; i < m; ++i) {
    trBS += dot_nt(Binv + (long)i * m, S + (long)i * m, m);
    bbeta += b[i] * beta[i];
    bSb += beta[i] * dot_nt(S + (long)i * m, beta, m);
  }
  out[2 + d] = (double)(-N / (2 * s2) + trBS / (2 * s2 * s2) + yty / (2 * s2 * s2) - bbeta / (s2 * s2 * s2) + bSb / (2 * s2 * s2 * s2 * s2) +
                        (N * sf2 - trS) / (2 * s2 * s2));
  for (int i = 0; i < m; ++i)
    for (int k = 0; k < d; ++k) out[3 + d + (long)i * d + k] = (double)(-gz[(long)i * d + k] / (ell[k] * ell[k]));
  free(dl);
done:
  free(ell); free(Kzz); free(L); free(Linv); free(LinvT); free(S); free(B); free(LBinv); free(LBinvT); free(Binv); free(PA);
  free(T1); free(P); free(Gzz); free(Kc); free(A); free(b); free(c); free(beta); free(u); free(gl); free(gz); free(r);
  return rc;
}

int sgpr_hp_mantissa_bits(void) {
#ifdef USE_QUAD
  return 113;
#else
  return 64;
#endif
}
