// Fused ARD kernel-tile construction: k(x_n, z_m) for a streamed chunk of X rows, and the m x m Kzz.
// X rows are staged into shared memory with one 1-D bulk TMA copy per tile (cp.async.bulk -> UBLKCP) when the tile is
// full and 16-byte aligned; Z is pre-scaled by 1/ell into shared memory; each thread produces a 4 x 4 register tile
// (rows 4*ty+i, columns tx+16*j: shared reads are conflict-free and global stores are 128-byte coalesced).
#pragma once
#include "common.cuh"
#include "gemm_i8.cuh"

namespace ggp {

constexpr int KT_N = 64, KT_M = 64, KT_THREADS = 256;

// value of the stationary kernel given scaled squared distance d2 (>= 0).  kind: 0 rbf, 1 matern32, 2 matern52
__device__ __forceinline__ double kval(int kind, double sf2, double d2) {
  if (kind == 0) return sf2 * exp(-0.5 * d2);
  const double r = sqrt(d2);
  if (kind == 1) {
    const double a = 1.7320508075688772;
    return sf2 * (1.0 + a * r) * exp(-a * r);
  }
  const double a = 2.23606797749979;
  return sf2 * (1.0 + a * r + (5.0 / 3.0) * d2) * exp(-a * r);
}
// dk/d(d2)
__device__ __forceinline__ double kgrad(int kind, double sf2, double d2) {
  if (kind == 0) return -0.5 * sf2 * exp(-0.5 * d2);
  const double r = sqrt(d2);
  if (kind == 1) return -1.5 * sf2 * exp(-1.7320508075688772 * r);
  const double a = 2.23606797749979;
  return -(5.0 / 6.0) * sf2 * (1.0 + a * r) * exp(-a * r);
}

// Kc[b][n][m] = k(x_n, z_m; theta_b) for n in [0, n_fill): rows >= n_valid and columns >= M are written as zero.
// deriv = 1 emits dk/d(d2) instead (the multiplier of the backward epilogue for the non-RBF kernels).
// grid: (ceil(ldk/KT_M), ceil(n_fill/KT_N), batch)
__global__ void __launch_bounds__(KT_THREADS) k_build_kc(const double* __restrict__ X, int n_valid, int n_fill, int d,
                                                        const double* __restrict__ Z, int M,
                                                        const double* __restrict__ theta, int kind,
                                                        double* __restrict__ Kc, int64_t ldk, int64_t sK, int deriv = 0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);   // [KT_N][d]
  double* zs = xs + KT_N * d;                          // [d][KT_M]   (z / ell, transposed: conflict-free)
  double* il = zs + KT_M * d;                          // [d]
  __shared__ __align__(8) uint64_t bar;

  const int b = blockIdx.z;
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
      out[(int64_t)n * ldk + m] = (nv && m < M) ? (deriv ? kgrad(kind, sf2, d2[i][j]) : kval(kind, sf2, d2[i][j])) : 0.0;
    }
  }
}

// Tile build of the sliced-integer path (gemm_i8.cuh must be included first): one pass produces
//   * the FP64 tile Kc[n][m] (multiplier of the backward epilogue),
//   * its int8 digit planes Kq[i][n][m] with the fixed exponent e = i8_exp_for(sf2) (k <= sf2).
// Thread (tx, ty) owns rows 4 ty + i and the 4 CONSECUTIVE columns 4 tx + j: 32-byte FP64 stores and 4-byte digit stores, a half-warp
// covers 512 / 64 contiguous bytes of one row.  zs is laid out [d][4][16] so that the shared reads stay conflict-free.
// grid: (ldk / 64, ceil(n_valid / 64)); batch = 1.
__global__ void __launch_bounds__(KT_THREADS) k_build_kc_i8(const double* __restrict__ X, int n_valid,
                                                           int d, const double* __restrict__ Z, int M, const double* __restrict__ theta,
                                                           int kind, double* __restrict__ Kc, int64_t ldk, int8_t* __restrict__ Kq,
                                                           int64_t ldq, int64_t plane) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);   // [KT_N][d]
  double* zs = xs + KT_N * d;                          // [d][4][16]   (z / ell)
  double* il = zs + KT_M * d;                          // [d]
  __shared__ __align__(8) uint64_t bar;

  const double sf2 = theta[d];
  const int n0 = blockIdx.y * KT_N, m0 = blockIdx.x * KT_M;
  const int tid = threadIdx.x;
  const int rows = min(KT_N, n_valid - n0);

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
  for (int i = tid; i < d; i += KT_THREADS) il[i] = 1.0 / theta[i];
  for (int i = tid; i < KT_M * d; i += KT_THREADS) {
    const int r = i / d, c = i % d;
    zs[c * KT_M + (r & 3) * 16 + (r >> 2)] = (m0 + r < M) ? Z[(int64_t)(m0 + r) * d + c] / theta[c] : 0.0;
  }
  if (bulk) mbar_wait(&bar, 0);
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;  // tx -> columns 4 tx + j, ty -> rows 4 ty + i
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
  const double si = exp2((double)-i8_exp_for(sf2));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    const bool nv = n < n_valid;
    double kv[4];
    uint32_t pk[I8_NS];
#pragma unroll
    for (int q = 0; q < I8_NS; ++q) pk[q] = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // the FP64 tile keeps the UNQUANTISED value: it is the dk/dtheta factor of the backward epilogue, where the fixed-point grid
      // (relative error 3e-12 at k = 1e-5) moved the gradient by 2.5e-8 at the headline shape
      kv[j] = (nv && m0 + 4 * tx + j < M) ? kval(kind, sf2, d2[i][j]) : 0.0;
      int8_t dg[I8_NS];
      i8_digits(kv[j] * si, dg);
#pragma unroll
      for (int q = 0; q < I8_NS; ++q) pk[q] |= ((uint32_t)(uint8_t)dg[q]) << (8 * j);
    }
    if (nv) {
      double* o = Kc + (int64_t)n * ldk + m0 + 4 * tx;
      *reinterpret_cast<double2*>(o) = make_double2(kv[0], kv[1]);
      *reinterpret_cast<double2*>(o + 2) = make_double2(kv[2], kv[3]);
#pragma unroll
      for (int q = 0; q < I8_NS; ++q) *reinterpret_cast<uint32_t*>(Kq + (int64_t)q * plane + (int64_t)n * ldq + m0 + 4 * tx) = pk[q];
    }
  }
}

// Kzz[b][i][j] = k(z_i, z_j) + jitter_b * delta_ij for i,j < M ; identity on the padding up to Mp.  grid: (Mp/16, Mp/16, batch)
__global__ void k_build_kzz(const double* __restrict__ Z, int M, int Mp, int d, const double* __restrict__ theta,
                            const double* __restrict__ jitter, int kind, double* __restrict__ Kzz, int64_t sK) {
  const int b = blockIdx.z;
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= Mp || j >= Mp) return;
  const double* th = theta + (int64_t)b * (d + 2);
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
