// Predictive row kernels, padding copies and the DMMA roofline probe.
#pragma once
#include "common.cuh"

namespace ggp {

// One warp per test row n:  mean = tT[n,:].c ; var = ||tT[n,:]||^2 + max(sf2 - ||aT[n,:]||^2, 0) (+ s2)
// grid (ceil(nv/8), batch), block 256
__global__ void __launch_bounds__(256) k_predict_rows(const double* __restrict__ aT, const double* __restrict__ tT, int64_t ld,
                                                      int64_t sC, const double* __restrict__ cvec, int64_t sv,
                                                      const double* __restrict__ theta, int d, int M, int nv, int add_noise,
                                                      double* __restrict__ mean, double* __restrict__ var, int64_t sOut) {
  const int b = blockIdx.y, n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= nv) return;
  const double* a = aT + b * sC + (int64_t)n * ld;
  const double* t = tT + b * sC + (int64_t)n * ld;
  const double* c = cvec + b * sv;
  double mu = 0.0, tt = 0.0, aa = 0.0;
  for (int j = lane; j < M; j += 32) {
    const double tj = t[j], aj = a[j];
    mu = fma(tj, c[j], mu);
    tt = fma(tj, tj, tt);
    aa = fma(aj, aj, aa);
  }
  mu = warp_sum(mu);
  tt = warp_sum(tt);
  aa = warp_sum(aa);
  if (lane == 0) {
    const double sf2 = theta[(int64_t)b * (d + 2) + d], s2 = theta[(int64_t)b * (d + 2) + d + 1];
    mean[b * sOut + n] = mu;
    var[b * sOut + n] = tt + fmax(sf2 - aa, 0.0) + (add_noise ? s2 : 0.0);
  }
}

// cov[b][n][n] = var[b][n]
__global__ void k_cov_diag(double* __restrict__ cov, int64_t ns, const double* __restrict__ var) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n < ns) cov[b * ns * ns + n * ns + n] = var[b * ns + n];
}

// to_padded=1: P[b][i][j] = (i,j<m) ? a[b][i][j] : delta_ij ; to_padded=0: a[b][i][j] = P[b][i][j]
__global__ void k_pad_copy(double* __restrict__ a, int m, double* __restrict__ P, int Mp, int64_t sM, int to_padded) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x, b = blockIdx.z;
  if (i >= Mp || j >= Mp) return;
  if (to_padded) {
    P[b * sM + (int64_t)i * Mp + j] = (i < m && j < m) ? a[(int64_t)b * m * m + (int64_t)i * m + j] : ((i == j) ? 1.0 : 0.0);
  } else if (i < m && j < m) {
    a[(int64_t)b * m * m + (int64_t)i * m + j] = P[b * sM + (int64_t)i * Mp + j];
  }
}

// Register-resident DMMA loop (NI x 4 independent accumulator pairs per warp; NI = 8 is the production fragment shape).
template <int NI, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_dmma_probe(double* __restrict__ sink, int iters) {
  double acc[NI][4][2];
  double a[NI], b[4];
#pragma unroll
  for (int i = 0; i < NI; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 1.0 - 1e-9 * (threadIdx.x + j);
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace ggp
