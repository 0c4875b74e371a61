// Small-tile FP64 tensor-core (DMMA.8x8x4) "NT" GEMM for the m x m section:  C[i,j] = alpha * sum_k A[i,k] * B[j,k] + beta * C[i,j]
// (same GemmP contract as k_gemm_nt / k_gemm_tma, EPI_STORE without row dots, splits or symmetric tile lists).
//
// Why a third mainloop: the products of the m x m section (triangular inverse by recursive doubling, B^-1 = L_B^-T L_B^-1,
// Q = L^-T P_A, G_zz) are 1024^3 or smaller with TRIANGULAR k ranges.  On 128 x 128 tiles a 1024^2 output is 64 work items for
// 148 SMs and its heaviest tile carries the full k range, so the launch takes as long as that one tile: 144 us for a product
// whose flops fit in 29 us (profiles/r2_launches_bench_N1e6_i8.csv); the 512^3 level of the inverse runs on 16 CTAs.  Here the
// work item is a 64 x 64 tile (4 x the items, a quarter of the work in the heaviest one), k ranges are clipped at 64, two CTAs
// share an SM (one CTA's barrier / epilogue under the other's DMMAs), and the grid is enumerated heaviest tiles first so the
// hardware block scheduler does list scheduling.  8 warps as 2 x 4, warp tile 32 x 16 = 4 x 2 DMMA fragments (6 shared loads
// per 8 DMMAs), operands by 16-byte LDGSTS into a 3-stage ring of [64][32 + 4] double tiles (the same conflict-free pitch as
// k_gemm_nt).  Operands must be clean outside their triangle (they are: every kmode caller of the library GEMMs relies on it).
#pragma once
#include "gemm_dmma.cuh"

namespace ggp {

constexpr int S_T = 64, S_BK = 32, S_STAGES = 3, S_LDS = S_BK + 4, S_THREADS = 256;
constexpr int S_SMEM = S_STAGES * 2 * S_T * S_LDS * 8;   // 110592 B: two CTAs per SM

// tile enumeration of one z (batch x pair) slice: y-major, lower-triangular operands reversed so that heavy tiles come first
__device__ __forceinline__ void mm64_tile(const GemmP& p, int t, int& tm, int& tn) {
  const int y = t / p.ntn, x = t - y * p.ntn;
  tm = (p.kmode & KM_A_LOWER) ? p.ntm - 1 - y : y;
  tn = (p.kmode & KM_B_LOWER) ? p.ntn - 1 - x : x;
}

__global__ void __launch_bounds__(S_THREADS, 2) k_mm64(const GemmP p) {
  extern __shared__ __align__(16) unsigned char mm64_raw[];
  double* sA = reinterpret_cast<double*>(mm64_raw);
  double* sB = sA + S_STAGES * S_T * S_LDS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, q = lane & 3;

  // Work item of this CTA.  Items are enumerated heaviest first; the hardware places CTAs 0 .. P - 1 (P = SM count) one per SM and the
  // rest as second CTAs of the same SMs in the same order, so the items beyond the first wave are taken in REVERSE: the lightest tile
  // shares an SM with the heaviest one (two co-resident CTAs split the FP64 pipe; heavy + heavy took 62 us for a triangular 1024^3 product)
  int w = blockIdx.x;
  if (p.fold > 0 && w >= p.fold) w = p.total - 1 - (w - p.fold);
  // slices (batch x pair) are interleaved so that "heaviest first" holds across them
  const int nz = p.total / p.tiles_per_z;
  int z = w % nz;
  const int t = w / nz;
  int tm, tn;
  mm64_tile(p, t, tm, tn);
  const int pz = z % p.nz2, bz = z / p.nz2;
  if (p.sym == 2 && tn > tm) return;   // lower tiles only (the caller reads the lower triangle): the other CTAs leave at once
  int k_lo = 0, k_hi = p.K;
  if (p.kmode & KM_A_LOWER) k_hi = min(k_hi, (tm + 1) * S_T);
  if (p.kmode & KM_B_LOWER) k_hi = min(k_hi, (tn + 1) * S_T);
  if (p.kmode & KM_A_UPPER) k_lo = max(k_lo, tm * S_T);
  if (p.kmode & KM_B_UPPER) k_lo = max(k_lo, tn * S_T);
  const int nkt = k_hi > k_lo ? (k_hi - k_lo + S_BK - 1) / S_BK : 0;
  if (nkt == 0) return;   // like the 128-tile kernels: an empty k range leaves C alone

  const double* __restrict__ Ab = p.A + bz * p.sA + pz * p.sA2;
  const double* __restrict__ Bb = p.B + bz * p.sB + pz * p.sB2;
  // loader: 16 threads per row (16-byte chunks of a 32-double k-step), 16 rows per pass, 4 passes per operand
  const int ld_row = tid >> 4, ld_chunk = tid & 15;
  const int ld_soff = ld_row * S_LDS + ld_chunk * 2;
  unsigned okA = 0, okB = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    okA |= ((tm * S_T + ld_row + 16 * r < p.M) ? 1u : 0u) << r;
    okB |= ((tn * S_T + ld_row + 16 * r < p.N) ? 1u : 0u) << r;
  }
  const double* gA = Ab + (int64_t)(tm * S_T + ld_row) * p.lda + ld_chunk * 2;
  const double* gB = Bb + (int64_t)(tn * S_T + ld_row) * p.ldb + ld_chunk * 2;
  const int64_t strA = (int64_t)16 * p.lda, strB = (int64_t)16 * p.ldb;
  auto issue = [&](int it) {
    if (it < nkt) {
      const int stage = it % S_STAGES, k0 = k_lo + it * S_BK;
      int kb = (k_hi - (k0 + ld_chunk * 2)) * 8;
      kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
      double* dA = sA + stage * S_T * S_LDS + ld_soff;
      double* dB = sB + stage * S_T * S_LDS + ld_soff;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int ba = ((okA >> r) & 1u) ? kb : 0, bb = ((okB >> r) & 1u) ? kb : 0;
        cp_async16(dA + r * 16 * S_LDS, ba ? (gA + r * strA + k0) : Ab, ba);
        cp_async16(dB + r * 16 * S_LDS, bb ? (gB + r * strB + k0) : Bb, bb);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < S_STAGES - 1; ++s) issue(s);

  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int frag_a = (wm * 32 + g) * S_LDS + q, frag_b = (wn * 16 + g) * S_LDS + q;
  for (int it = 0; it < nkt; ++it) {
    cp_async_wait<S_STAGES - 2>();
    __syncthreads();               // stage `it` has landed for everyone; stage it - 1 is free for the refill below
    issue(it + S_STAGES - 1);
    const double* cA = sA + (it % S_STAGES) * S_T * S_LDS + frag_a;
    const double* cB = sB + (it % S_STAGES) * S_T * S_LDS + frag_b;
#pragma unroll
    for (int kk = 0; kk < S_BK / 4; ++kk) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = cA[i * 8 * S_LDS + kk * 4];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = cB[j * 8 * S_LDS + kk * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  double* __restrict__ C = p.C + bz * p.sC + pz * p.sC2;
  double* __restrict__ Ct = p.Ct ? p.Ct + bz * p.sC + pz * p.sC2 : nullptr;
  const bool vec16 = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = tm * S_T + wm * 32 + i * 8 + g;
    if (gr >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gc = tn * S_T + wn * 16 + j * 8 + 2 * q;
      double* dst = C + (int64_t)gr * p.ldc + gc;
      double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
      if (gc + 1 < p.N) {
        if (vec16) {
          double2* d2 = reinterpret_cast<double2*>(dst);
          if (p.beta != 0.0) { const double2 o = *d2; v0 += p.beta * o.x; v1 += p.beta * o.y; }
          *d2 = make_double2(v0, v1);
        } else {
          if (p.beta != 0.0) { v0 += p.beta * dst[0]; v1 += p.beta * dst[1]; }
          dst[0] = v0; dst[1] = v1;
        }
        if (Ct) { Ct[(int64_t)gc * p.ldct + gr] = v0; Ct[(int64_t)(gc + 1) * p.ldct + gr] = v1; }
      } else if (gc < p.N) {
        if (p.beta != 0.0) v0 += p.beta * dst[0];
        dst[0] = v0;
        if (Ct) Ct[(int64_t)gc * p.ldct + gr] = v0;
      }
    }
  }
}

}  // namespace ggp
