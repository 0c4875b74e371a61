// FP64 tensor-core (DMMA.8x8x4) "NT" GEMM:  C[i,j] (op)= alpha * sum_k A[i,k] * B[j,k]
//
// One CTA = 128x128 output tile, 8 warps (2 x 4), warp tile 64x32 = 8x4 DMMA fragments (64 accumulator doubles per
// thread).  Operand tiles (128 rows x 16 k) are staged global->shared with a 4-stage cp.async (LDGSTS) ring; rows are
// padded to 20 doubles so the per-half-warp 64-bit fragment loads (row = lane/4, k = lane%4) hit 16 distinct bank pairs.
// The same mainloop serves
//   * the triangular multiply  A^T-chunk = L^{-1} * Kc^T            (kmode: k-range clipped to the triangle)
//   * the symmetric rank-k update  S += A_chunk * A_chunk^T          (sym: upper tiles only, split-K partial buffers)
//   * the backward product  G = P * Kc^T  with the fused moments epilogue  (G o K) * [1, x, x^2]
//   * every m x m product of the "finish" section.
// tcgen05 has no f64 kind, so this legacy warp-level path is the only FP64 tensor route on sm_100a.
#pragma once
#include "common.cuh"

namespace ggp {

#ifndef GGP_BK
#define GGP_BK 32
#endif
#ifndef GGP_STAGES
#define GGP_STAGES 3
#endif
constexpr int BM = 128, BN = 128, BK = GGP_BK, STAGES = GGP_STAGES, LDS = BK + 4, GEMM_THREADS = 256;
constexpr int LD_TPR = BK / 2;                     // loader threads per tile row (16-byte chunks per row)
constexpr int LD_RPP = GEMM_THREADS / LD_TPR;      // rows per loader pass
constexpr int GEMM_SMEM_PIPE = STAGES * (BM + BN) * LDS * 8;  // 163840 B
constexpr int EPI_LDW = BN + 4;                                // W tile row stride (doubles), == 4 mod 16
constexpr int MOM_QB = 24;                                     // moment columns per DMMA block (3 n-fragments)
constexpr int GEMM_SMEM_MOM = (BM * EPI_LDW + MOM_QB * EPI_LDW) * 8;  // 160512 B
constexpr int GEMM_SMEM = GEMM_SMEM_PIPE > GEMM_SMEM_MOM ? GEMM_SMEM_PIPE : GEMM_SMEM_MOM;

enum { KM_A_LOWER = 1, KM_A_UPPER = 2, KM_B_LOWER = 4, KM_B_UPPER = 8 };
enum { EPI_STORE = 0, EPI_MOMENTS = 1 };

struct GemmP {
  const double* A; int64_t lda, sA, sA2;
  const double* B; int64_t ldb, sB, sB2;
  double* C;       int64_t ldc, sC, sC2;
  int M, N, K;
  int nz2;          // inner batch count (blockIdx.z = (b*nz2 + p)*splits + split)
  int splits;       // split-K factor; each split adds into C + split*sSplit (requires beta = 1, pre-zeroed)
  int64_t sSplit;
  double alpha, beta;
  int kmode, sym, heavy_first;
  // EPI_MOMENTS only
  const double* u;  int64_t su;            // [M] per batch
  const double* yv;                        // [N]
  const double* Kc; int64_t ldk, sK;       // [N x ldk] per batch  (k(x_n, z_i) at Kc[n*ldk + i])
  const double* Xc; int d;                 // [N x d]
  double* mom;      int64_t sMomTile, sMom;  // [batch][tile_n][M][2d+1]
  // EPI_STORE, optional: per-tile row dots against yv (b = A y)
  double* rowdot;   int64_t sRowdot;         // [batch][tile_n][M]
};

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_gemm_nt(const GemmP p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + STAGES * BM * LDS;

  const int tn = blockIdx.x;
  const int tm = p.heavy_first ? (gridDim.y - 1 - blockIdx.y) : blockIdx.y;
  if ((p.sym == 1 && tn < tm) || (p.sym == 2 && tn > tm)) return;  // 1: upper tiles only, 2: lower tiles only
  int z = blockIdx.z;
  const int split = z % p.splits; z /= p.splits;
  const int pz = z % p.nz2;
  const int bz = z / p.nz2;

  const double* __restrict__ A = p.A + bz * p.sA + pz * p.sA2;
  const double* __restrict__ B = p.B + bz * p.sB + pz * p.sB2;
  double* __restrict__ C = p.C + bz * p.sC + pz * p.sC2 + split * p.sSplit;

  // k range (multiples of BK by construction of the tile sizes)
  int k_lo = 0, k_hi = p.K;
  if (p.kmode & KM_A_LOWER) k_hi = min(k_hi, (tm + 1) * BM);
  if (p.kmode & KM_B_LOWER) k_hi = min(k_hi, (tn + 1) * BN);
  if (p.kmode & KM_A_UPPER) k_lo = max(k_lo, tm * BM);
  if (p.kmode & KM_B_UPPER) k_lo = max(k_lo, tn * BN);
  int nkt = (k_hi > k_lo) ? (k_hi - k_lo + BK - 1) / BK : 0;
  int it_lo = 0, it_hi = nkt;
  if (p.splits > 1) {
    int per = (nkt + p.splits - 1) / p.splits;
    it_lo = split * per;
    it_hi = min(nkt, it_lo + per);
    if (it_lo >= it_hi) return;
  }
  const int niter = it_hi - it_lo;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
  const int g = lane >> 2, q = lane & 3;

  // loader mapping: 8 threads cover one 128-byte row (16 doubles), 32 rows per pass, 4 passes per operand.
  // Row pointers, validity and shared offsets are computed ONCE; the per-iteration cost is one 64-bit add per row.
  const int ld_row = tid / LD_TPR, ld_chunk = tid % LD_TPR;
  const int row0_m = tm * BM, row0_n = tn * BN;
  // one base pointer per operand; rows of later passes are reached with a constant stride
  const double* gA0 = A + (int64_t)(row0_m + ld_row) * p.lda + ld_chunk * 2;
  const double* gB0 = B + (int64_t)(row0_n + ld_row) * p.ldb + ld_chunk * 2;
  const int64_t strA = (int64_t)LD_RPP * p.lda, strB = (int64_t)LD_RPP * p.ldb;
  unsigned okmask = 0;
#pragma unroll
  for (int r = 0; r < BM / LD_RPP; ++r) okmask |= ((row0_m + ld_row + r * LD_RPP < p.M) ? 1u : 0u) << r;
#pragma unroll
  for (int r = 0; r < BN / LD_RPP; ++r) okmask |= ((row0_n + ld_row + r * LD_RPP < p.N) ? 1u : 0u) << (16 + r);
  const int ld_soff = ld_row * LDS + ld_chunk * 2;
  const int krem0 = p.K - ld_chunk * 2;  // elements left at k = 0 for this thread's 16-byte column

  auto load_stage = [&](int stage, int it) {
    const int k0 = k_lo + it * BK;
    int kb = (krem0 - k0) * 8;
    kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
    double* dA = sA + stage * BM * LDS + ld_soff;
    double* dB = sB + stage * BN * LDS + ld_soff;
#pragma unroll
    for (int r = 0; r < BM / LD_RPP; ++r) {
      const int bytes = ((okmask >> r) & 1u) ? kb : 0;
      cp_async16(dA + r * LD_RPP * LDS, bytes ? (gA0 + r * strA + k0) : A, bytes);
    }
#pragma unroll
    for (int r = 0; r < BN / LD_RPP; ++r) {
      const int bytes = ((okmask >> (16 + r)) & 1u) ? kb : 0;
      cp_async16(dB + r * LD_RPP * LDS, bytes ? (gB0 + r * strB + k0) : B, bytes);
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < niter) load_stage(s, it_lo + s);
    cp_async_commit();
  }

  const int frag_a = (wm * 64 + g) * LDS + q, frag_b = (wn * 32 + g) * LDS + q;
  const bool diag_lower = (p.kmode & KM_A_LOWER) != 0;
  for (int it = 0; it < niter; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const double* cA = sA + (it % STAGES) * BM * LDS + frag_a;
    const double* cB = sB + (it % STAGES) * BN * LDS + frag_b;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = cA[i * 8 * LDS + kk * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = cB[j * 8 * LDS + kk * 4];
      if (diag_lower && k_lo + (it_lo + it) * BK + kk * 4 >= row0_m) {
        // inside the diagonal block of a lower-triangular A: fragment rows [8i, 8i+8) need k <= row only
        const int kfrag = k_lo + (it_lo + it) * BK + kk * 4 - row0_m - wm * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (kfrag < i * 8 + 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      if (kk == 0) {
        // refill the stage consumed in the previous iteration only after this iteration's first MMAs are in flight
        const int nx = it + STAGES - 1;
        if (nx < niter) load_stage(nx % STAGES, it_lo + nx);
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();

  if (EPI == EPI_STORE) {
    if (p.rowdot) {
      // fused  rowdot[tile_n][i] = sum_{n in tile} alpha*acc[i,n] * yv[n]   (b = A y, fixed-order reduction)
      __syncthreads();
      double* sR = reinterpret_cast<double*>(smem_raw);  // [4][BM]
      double yv[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gc = row0_n + wn * 32 + j * 8 + 2 * q;
        yv[j][0] = gc < p.N ? p.yv[gc] : 0.0;
        yv[j][1] = gc + 1 < p.N ? p.yv[gc + 1] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double sdot = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) sdot = fma(acc[i][j][1], yv[j][1], fma(acc[i][j][0], yv[j][0], sdot));
        sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
        sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
        if (q == 0) sR[wn * BM + wm * 64 + i * 8 + g] = sdot;
      }
      __syncthreads();
      if (tid < BM && row0_m + tid < p.M)
        p.rowdot[bz * p.sRowdot + (int64_t)tn * p.M + row0_m + tid] =
            p.alpha * (((sR[tid] + sR[BM + tid]) + sR[2 * BM + tid]) + sR[3 * BM + tid]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gr = row0_m + wm * 64 + i * 8 + g;
      if (gr >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gc = row0_n + wn * 32 + j * 8 + 2 * q;
        double* dst = C + (int64_t)gr * p.ldc + gc;
        double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
        if (gc + 1 < p.N) {
          if (p.beta != 0.0) { v0 += p.beta * dst[0]; v1 += p.beta * dst[1]; }
          dst[0] = v0; dst[1] = v1;
        } else if (gc < p.N) {
          if (p.beta != 0.0) v0 += p.beta * dst[0];
          dst[0] = v0;
        }
      }
    }
  } else {
    // ---- fused backward epilogue:  W = (alpha*acc + u y^T) o K ;  mom[i, :] = sum_n W[i,n] * [1, x_n, x_n^2] ----
    __syncthreads();  // pipeline buffers are dead; alias them
    double* sW = reinterpret_cast<double*>(smem_raw);     // [BM][EPI_LDW]
    double* sPhi = sW + BM * EPI_LDW;                      // [MOM_QB][EPI_LDW]
    const double* __restrict__ Kc = p.Kc + bz * p.sK;
    const double* __restrict__ uu = p.u + bz * p.su;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int lr = wm * 64 + i * 8 + g;
      const int gr = row0_m + lr;
      const double ui = (gr < p.M) ? uu[gr] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int lc = wn * 32 + j * 8 + 2 * q;
        const int gc = row0_n + lc;
        double w0 = 0.0, w1 = 0.0;
        if (gr < p.M) {
          if (gc < p.N) w0 = (p.alpha * acc[i][j][0] + ui * p.yv[gc]) * Kc[(int64_t)gc * p.ldk + gr];
          if (gc + 1 < p.N) w1 = (p.alpha * acc[i][j][1] + ui * p.yv[gc + 1]) * Kc[(int64_t)(gc + 1) * p.ldk + gr];
        }
        sW[lr * EPI_LDW + lc] = w0;
        sW[lr * EPI_LDW + lc + 1] = w1;
      }
    }
    const int nq = 2 * p.d + 1;
    double* mom = p.mom + bz * p.sMom + (int64_t)tn * p.sMomTile;
    for (int q0 = 0; q0 < nq; q0 += MOM_QB) {
      __syncthreads();
      // Phi[qq][n] for this block of moment columns
      for (int idx = tid; idx < MOM_QB * BN; idx += GEMM_THREADS) {
        const int qq = idx / BN, n = idx % BN;
        const int qa = q0 + qq, gn = row0_n + n;
        double v = 0.0;
        if (qa < nq && gn < p.N) {
          if (qa == 0) v = 1.0;
          else if (qa <= p.d) v = p.Xc[(int64_t)gn * p.d + (qa - 1)];
          else { const double x = p.Xc[(int64_t)gn * p.d + (qa - 1 - p.d)]; v = x * x; }
        }
        sPhi[qq * EPI_LDW + n] = v;
      }
      __syncthreads();
      // each warp: 16 rows x 24 cols = 2 x 3 fragments, K = 128
      double m2[2][3][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m2[i][j][0] = m2[i][j][1] = 0.0;
      const double* wA = sW + (warp * 16 + g) * EPI_LDW + q;
      const double* wB = sPhi + g * EPI_LDW + q;
#pragma unroll 4
      for (int kk = 0; kk < BN / 4; ++kk) {
        double a0 = wA[kk * 4], a1 = wA[8 * EPI_LDW + kk * 4];
        double b0 = wB[kk * 4], b1 = wB[8 * EPI_LDW + kk * 4], b2 = wB[16 * EPI_LDW + kk * 4];
        dmma884(m2[0][0][0], m2[0][0][1], a0, b0);
        dmma884(m2[0][1][0], m2[0][1][1], a0, b1);
        dmma884(m2[0][2][0], m2[0][2][1], a0, b2);
        dmma884(m2[1][0][0], m2[1][0][1], a1, b0);
        dmma884(m2[1][1][0], m2[1][1][1], a1, b1);
        dmma884(m2[1][2][0], m2[1][2][1], a1, b2);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int gr = row0_m + warp * 16 + i * 8 + g;
        if (gr >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int qa = q0 + j * 8 + 2 * q;
          if (qa < nq) mom[(int64_t)gr * nq + qa] = m2[i][j][0];
          if (qa + 1 < nq) mom[(int64_t)gr * nq + qa + 1] = m2[i][j][1];
        }
      }
    }
  }
}

}  // namespace ggp
