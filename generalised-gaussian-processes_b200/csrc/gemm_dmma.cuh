// FP64 tensor-core (DMMA.8x8x4) "NT" GEMM:  C[i,j] (op)= alpha * sum_k A[i,k] * B[j,k]
//
// PERSISTENT kernel: one CTA per SM walks a static round-robin list of work items (output tile x batch x split), and the
// cp.async operand pipeline runs CONTINUOUSLY across tile boundaries (a load cursor runs STAGES-1 k-steps ahead of the compute
// cursor, also into the next tile), so the FP64 pipe never waits for a pipeline fill between tiles.
//
// Work item = 128x128 output tile, 8 warps (2 x 4), warp tile 64x32 = 8x4 DMMA fragments (64 accumulator doubles / thread).
// Operand tiles (128 rows x BK) are staged global->shared with a 3-stage cp.async (LDGSTS) ring; rows are padded to BK+4
// doubles so the per-half-warp 64-bit fragment loads (row = lane/4, k = lane%4) hit 16 distinct bank pairs.
// The same mainloop serves
//   * the triangular multiply  A^T-chunk = L^{-1} * Kc^T     (kmode: k-range clipped to the triangle, staircase inside the
//                                                             diagonal block, heavy tiles first) + fused b = A y row dots
//   * the symmetric rank-k update  S += A_chunk * A_chunk^T   (sym: upper tiles only, split-K partial buffers)
//   * the backward product  G = P * Kc^T  with the fused moments epilogue  ((G + u y^T) o K) * [1, x, x^2]  computed straight
//     from the accumulator registers with a second round of DMMAs (no shared-memory staging)
//   * every m x m product of the "finish" section.
// tcgen05 has no f64 kind, so this warp-level path is the only FP64 tensor route on sm_100a.
#pragma once
#include "common.cuh"

namespace ggp {

#ifndef GGP_BK
#define GGP_BK 32
#endif
#ifndef GGP_STAGES
#define GGP_STAGES 3
#endif
#ifndef GGP_BN
#define GGP_BN 128
#endif
#ifndef GGP_CTAS_PER_SM
#define GGP_CTAS_PER_SM 1
#endif
constexpr int BM = 128, BN = GGP_BN, BK = GGP_BK, STAGES = GGP_STAGES, LDS = BK + 4;
constexpr int WARPS_M = BM / 64, WARPS_N = BN / 32, GEMM_THREADS = WARPS_M * WARPS_N * 32;
constexpr int CTAS_PER_SM = GGP_CTAS_PER_SM;       // co-resident CTAs: one CTA's epilogue / barrier drain overlaps the other's mainloop
constexpr int LD_TPR = BK / 2;                     // loader threads per tile row (16-byte chunks per row)
constexpr int LD_RPP = GEMM_THREADS / LD_TPR;      // rows per loader pass
constexpr int GEMM_SMEM_PIPE = STAGES * (BM + BN) * LDS * 8;
constexpr int GEMM_SMEM = GEMM_SMEM_PIPE + WARPS_N * BM * 8;   // + row-dot exchange [WARPS_N][BM]
static_assert(BM == 128 && (BN == 64 || BN == 128), "tile shapes: 128x128 (8 warps) or 128x64 (4 warps)");
static_assert(BM / LD_RPP <= 16 && BN / LD_RPP <= 16, "okmask layout");
// symmetric outputs with rectangular tiles: tile (tm, tn) touches the upper triangle iff tn >= tm * (BM / BN)
constexpr int TRATIO = BM / BN;
__host__ __device__ inline int sym_upper_tiles(int ntm, int ntn) {
  int t = 0;
  for (int tm = 0; tm < ntm; ++tm) t += max(0, ntn - tm * TRATIO);
  return t;
}
__host__ __device__ inline int sym_lower_tiles(int ntm, int ntn) {
  int t = 0;
  for (int tm = 0; tm < ntm; ++tm) t += min(ntn, (tm + 1) * TRATIO);
  return t;
}

enum { KM_A_LOWER = 1, KM_A_UPPER = 2, KM_B_LOWER = 4, KM_B_UPPER = 8 };
enum { EPI_STORE = 0, EPI_MOMENTS = 1 };

struct GemmP {
  const double* A; int64_t lda, sA, sA2;
  const double* B; int64_t ldb, sB, sB2;
  double* C;       int64_t ldc, sC, sC2;
  int M, N, K;
  int nz2;          // inner batch count (z = (b*nz2 + p)*splits + split)
  int splits;       // split-K factor; each split adds into C + split*sSplit (requires beta = 1, pre-zeroed)
  int64_t sSplit;
  double alpha, beta;
  int kmode, sym, heavy_first;
  int n_major;      // non-symmetric outputs: enumerate all row tiles (tm) of one column tile (tn) consecutively
  // work decomposition (filled by launch_gemm)
  int ntm, ntn, tiles_per_z, total;
  // TMA path: 0/1 multipliers of the (inner, outer) batch coordinates per operand (0 = broadcast operand, stride 0)
  int tmA_pz, tmA_bz, tmB_pz, tmB_bz;
  // EPI_MOMENTS only
  const double* u;  int64_t su;            // [M] per batch
  const double* yv;                        // [N]
  const double* Kc; int64_t ldk, sK;       // [N x ldk] per batch  (k(x_n, z_i) at Kc[n*ldk + i])
  const double* Xc; int d;                 // [N x d]
  double* mom;      int64_t sMomTile, sMom;  // [batch][tile_n*4 + warp_col][M][2d+1]
  // EPI_STORE, optional: per-tile row dots against yv (b = A y)
  double* rowdot;   int64_t sRowdot;         // [batch][tile_n][M]
  // k_mm64 only, optional: the result is ALSO stored transposed, Ct[j * ldct + i] = C[i, j] (same batch / pair strides as C):
  // the recursive triangular inverse keeps L^-1 and L^-T in step without a transpose kernel per level
  double* Ct;       int64_t ldct;
  int fold;         // k_mm64 only: work items beyond the first `fold` (= SM count) are taken in reverse order, 0 = plain order
};

struct WorkItem {
  int tm, tn, bz, pz, split, k_lo, it_lo, niter;
};

template <int KS>
__device__ __forceinline__ void decode_work(const GemmP& p, int w, WorkItem& o) {
  int z = w / p.tiles_per_z, t = w - z * p.tiles_per_z;
  if (p.sym == 1) {  // upper tiles, row tm has ntn - tm*TRATIO of them
    int tm = 0;
    while (t >= p.ntn - tm * TRATIO) { t -= p.ntn - tm * TRATIO; ++tm; }
    o.tm = tm; o.tn = tm * TRATIO + t;
  } else if (p.sym == 2) {  // lower tiles, row tm has min(ntn, (tm+1)*TRATIO)
    int tm = 0;
    while (t >= min(p.ntn, (tm + 1) * TRATIO)) { t -= min(p.ntn, (tm + 1) * TRATIO); ++tm; }
    o.tm = tm; o.tn = t;
  } else {
    if (p.n_major) {
      o.tn = t / p.ntm;
      o.tm = t - o.tn * p.ntm;
    } else {
      const int y = t / p.ntn;
      o.tn = t - y * p.ntn;
      o.tm = p.heavy_first ? (p.ntm - 1 - y) : y;
    }
  }
  o.split = z % p.splits; z /= p.splits;
  o.pz = z % p.nz2;
  o.bz = z / p.nz2;
  int k_lo = 0, k_hi = p.K;
  if (p.kmode & KM_A_LOWER) k_hi = min(k_hi, (o.tm + 1) * BM);
  if (p.kmode & KM_B_LOWER) k_hi = min(k_hi, (o.tn + 1) * BN);
  if (p.kmode & KM_A_UPPER) k_lo = max(k_lo, o.tm * BM);
  if (p.kmode & KM_B_UPPER) k_lo = max(k_lo, o.tn * BN);
  const int nkt = (k_hi > k_lo) ? (k_hi - k_lo + KS - 1) / KS : 0;
  int it_lo = 0, it_hi = nkt;
  if (p.splits > 1) {
    const int per = (nkt + p.splits - 1) / p.splits;
    it_lo = o.split * per;
    it_hi = min(nkt, it_lo + per);
  }
  o.k_lo = k_lo;
  o.it_lo = it_lo;
  o.niter = it_hi > it_lo ? it_hi - it_lo : 0;
}

// Epilogues work on the accumulator registers and global memory only (the operand pipeline keeps streaming the next tile).
// NAMED_BAR: the CTA has a producer warp, so the consumer warps meet on named barrier 1 instead of __syncthreads().
template <bool NAMED_BAR>
__device__ __forceinline__ void epi_sync() {
  if (NAMED_BAR) asm volatile("bar.sync 1, 256;\n" ::: "memory");
  else __syncthreads();
}

template <int EPI, bool NAMED_BAR>
__device__ __forceinline__ void gemm_epilogue(const GemmP& p, const WorkItem& wi, double (&acc)[8][4][2], double* sR, int tid,
                                              int wm, int wn, int g, int q) {
  const int row0_m = wi.tm * BM, row0_n = wi.tn * BN;
  // ---------------- epilogue (registers + global only; the pipeline keeps streaming the next tile) ----------------
  if (EPI == EPI_STORE) {
    double* __restrict__ C = p.C + wi.bz * p.sC + wi.pz * p.sC2 + wi.split * p.sSplit;
    const bool vec16 = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    if (p.rowdot) {
      // fused  rowdot[tile_n][i] = sum_{n in tile} alpha*acc[i,n] * yv[n]   (b = A y, fixed-order reduction)
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
      epi_sync<NAMED_BAR>();
      if (tid < BM && row0_m + tid < p.M)
        p.rowdot[wi.bz * p.sRowdot + (int64_t)wi.tn * p.M + row0_m + tid] =
            p.alpha * (WARPS_N == 4 ? (((sR[tid] + sR[BM + tid]) + sR[2 * BM + tid]) + sR[3 * BM + tid]) : (sR[tid] + sR[BM + tid]));
      epi_sync<NAMED_BAR>();
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
#ifdef GGP_EXP_NOSTORE
        if (v0 == 1.2345e-300) dst[0] = v0;
        continue;
#endif
        if (gc + 1 < p.N) {
          if (vec16) {  // 16-byte accesses: the 4 lanes of a quad cover 2 full 32-B sectors per row
            double2* d2 = reinterpret_cast<double2*>(dst);
            if (p.beta != 0.0) { const double2 o = *d2; v0 += p.beta * o.x; v1 += p.beta * o.y; }
            *d2 = make_double2(v0, v1);
          } else {
            if (p.beta != 0.0) { v0 += p.beta * dst[0]; v1 += p.beta * dst[1]; }
            dst[0] = v0; dst[1] = v1;
          }
        } else if (gc < p.N) {
          if (p.beta != 0.0) v0 += p.beta * dst[0];
          dst[0] = v0;
        }
      }
    }
  } else {
    // ---- fused backward epilogue:  W = (alpha*acc + u y^T) o K ;  mom[i, :] = sum_n W[i,n] * [1, x_n, x_n^2] ----
    // W stays in the accumulator registers and is fed back to the tensor pipe as the A operand: the C fragment holds
    // W[g][2q+e], so taking k' = q with column n = 2q+e (e = 0,1) is a valid k-permutation as long as the B operand uses
    // the same one: lane (g,q) supplies Phi[n = 2q+e][moment column g].
    const double* __restrict__ Kc = p.Kc + wi.bz * p.sK;
    const double* __restrict__ uu = p.u + wi.bz * p.su;
    double yv[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = row0_n + wn * 32 + j * 8 + 2 * q;
      yv[j][0] = gc < p.N ? p.yv[gc] : 0.0;
      yv[j][1] = gc + 1 < p.N ? p.yv[gc + 1] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gr = row0_m + wm * 64 + i * 8 + g;
      const bool rok = gr < p.M;
      const double ui = rok ? uu[gr] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gc = row0_n + wn * 32 + j * 8 + 2 * q;
        const double k0 = (rok && gc < p.N) ? Kc[(int64_t)gc * p.ldk + gr] : 0.0;
        const double k1 = (rok && gc + 1 < p.N) ? Kc[(int64_t)(gc + 1) * p.ldk + gr] : 0.0;
        acc[i][j][0] = fma(ui, yv[j][0], p.alpha * acc[i][j][0]) * k0;
        acc[i][j][1] = fma(ui, yv[j][1], p.alpha * acc[i][j][1]) * k1;
      }
    }
    const int nq = 2 * p.d + 1;
    double* mom = p.mom + wi.bz * p.sMom + (int64_t)(wi.tn * WARPS_N + wn) * p.sMomTile;   // one slab per 32 columns of n
    for (int q0 = 0; q0 < nq; q0 += 8) {
      // B fragments for this block of 8 moment columns: phi[j][e] = Phi[n(j,q,e)][q0 + g]
      const int qa = q0 + g;
      double phi[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gn = row0_n + wn * 32 + j * 8 + 2 * q + e;
          double v = 0.0;
          if (qa < nq && gn < p.N) {
            if (qa == 0) v = 1.0;
            else if (qa <= p.d) v = p.Xc[(int64_t)gn * p.d + (qa - 1)];
            else { const double x = p.Xc[(int64_t)gn * p.d + (qa - 1 - p.d)]; v = x * x; }
          }
          phi[j][e] = v;
        }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double m0 = 0.0, m1 = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma884(m0, m1, acc[i][j][0], phi[j][0]);
          dmma884(m0, m1, acc[i][j][1], phi[j][1]);
        }
        const int gr = row0_m + wm * 64 + i * 8 + g;
        const int qc = q0 + 2 * q;
        if (gr < p.M) {
          if (qc < nq) mom[(int64_t)gr * nq + qc] = m0;
          if (qc + 1 < nq) mom[(int64_t)gr * nq + qc + 1] = m1;
        }
      }
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, CTAS_PER_SM) k_gemm_nt(const GemmP p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + STAGES * BM * LDS;
  double* sR = sB + STAGES * BN * LDS;  // [WARPS_N][BM] row-dot exchange (never aliased with the pipeline)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;  // WARPS_M x WARPS_N warps
  const int g = lane >> 2, q = lane & 3;
  const int ld_row = tid / LD_TPR, ld_chunk = tid % LD_TPR;
  const int ld_soff = ld_row * LDS + ld_chunk * 2;
  const int krem0 = p.K - ld_chunk * 2;  // elements left at k = 0 for this thread's 16-byte column
  const int64_t strA = (int64_t)LD_RPP * p.lda, strB = (int64_t)LD_RPP * p.ldb;
  const int frag_a = (wm * 64 + g) * LDS + q, frag_b = (wn * 32 + g) * LDS + q;
  const bool diag_lower = (p.kmode & KM_A_LOWER) != 0;
  // static snake (boustrophedon) assignment: in round r CTA c takes item r*G + (r odd ? G-1-c : c); with the heavy-first
  // ordering of the triangular multiply this matches dynamic list scheduling
  const int G = gridDim.x, nrounds = (p.total + G - 1) / G;
  auto item_of = [&](int r) { return r * G + ((r & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x); };

  // ---------------- load cursor ----------------
  int rL = 0, itL = 0, nitL = 0, kbaseL = 0, nloads = 0;
  const double *gA0 = p.A, *gB0 = p.B, *baseA = p.A, *baseB = p.B;
  unsigned okmask = 0;
  auto setup_load = [&]() {
    WorkItem wi;
    nitL = 0;
    while (rL < nrounds) {
      const int w = item_of(rL);
      if (w < p.total) {
        decode_work<BK>(p, w, wi);
        if (wi.niter > 0) break;
      }
      ++rL;
    }
    if (rL >= nrounds) return;
    nitL = wi.niter;
    itL = 0;
    kbaseL = wi.k_lo + wi.it_lo * BK;
    baseA = p.A + wi.bz * p.sA + wi.pz * p.sA2;
    baseB = p.B + wi.bz * p.sB + wi.pz * p.sB2;
    const int r0m = wi.tm * BM + ld_row, r0n = wi.tn * BN + ld_row;
    gA0 = baseA + (int64_t)r0m * p.lda + ld_chunk * 2;
    gB0 = baseB + (int64_t)r0n * p.ldb + ld_chunk * 2;
    okmask = 0;
#pragma unroll
    for (int r = 0; r < BM / LD_RPP; ++r) okmask |= ((r0m + r * LD_RPP < p.M) ? 1u : 0u) << r;
#pragma unroll
    for (int r = 0; r < BN / LD_RPP; ++r) okmask |= ((r0n + r * LD_RPP < p.N) ? 1u : 0u) << (16 + r);
  };
  auto issue_load = [&]() {
    if (rL < nrounds) {
      const int stage = nloads % STAGES;
      const int k0 = kbaseL + itL * BK;
      int kb = (krem0 - k0) * 8;
      kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
      double* dA = sA + stage * BM * LDS + ld_soff;
      double* dB = sB + stage * BN * LDS + ld_soff;
#pragma unroll
      for (int r = 0; r < BM / LD_RPP; ++r) {
        const int bytes = ((okmask >> r) & 1u) ? kb : 0;
        cp_async16(dA + r * LD_RPP * LDS, bytes ? (gA0 + r * strA + k0) : baseA, bytes);
      }
#pragma unroll
      for (int r = 0; r < BN / LD_RPP; ++r) {
        const int bytes = ((okmask >> (16 + r)) & 1u) ? kb : 0;
        cp_async16(dB + r * LD_RPP * LDS, bytes ? (gB0 + r * strB + k0) : baseB, bytes);
      }
      if (++itL == nitL) {
        ++rL;
        setup_load();
      }
    }
    ++nloads;
    cp_async_commit();
  };

  setup_load();
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue_load();

  // ---------------- compute cursor ----------------
  int gs = 0;  // global k-step counter (stage = gs % STAGES)
  for (int rC = 0; rC < nrounds; ++rC) {
    const int wC = item_of(rC);
    if (wC >= p.total) continue;
    WorkItem wi;
    decode_work<BK>(p, wC, wi);
    if (wi.niter == 0) continue;
    const int row0_m = wi.tm * BM, row0_n = wi.tn * BN;
    const int kbaseC = wi.k_lo + wi.it_lo * BK;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int it = 0; it < wi.niter; ++it, ++gs) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const double* cA = sA + (gs % STAGES) * BM * LDS + frag_a;
      const double* cB = sB + (gs % STAGES) * BN * LDS + frag_b;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = cA[i * 8 * LDS + kk * 4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = cB[j * 8 * LDS + kk * 4];
        if (diag_lower && kbaseC + it * BK + kk * 4 >= row0_m) {
          // inside the diagonal block of a lower-triangular A: fragment rows [8i, 8i+8) need k <= row only
          // (a predicated-off DMMA still occupies the tensor pipe for its full 16 cycles - measured with ncu - so the staircase
          //  must be real control flow: jump into the fragment-row sequence at the first row that needs this k-step)
          const int kfrag = kbaseC + it * BK + kk * 4 - row0_m - wm * 64;
          const int i_lo = kfrag < 0 ? 0 : (kfrag >> 3);
#define GGP_FRAG_ROW(i)                                                          \
  _Pragma("unroll") for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          switch (i_lo) {
            case 0: GGP_FRAG_ROW(0)  // fallthrough
            case 1: GGP_FRAG_ROW(1)
            case 2: GGP_FRAG_ROW(2)
            case 3: GGP_FRAG_ROW(3)
            case 4: GGP_FRAG_ROW(4)
            case 5: GGP_FRAG_ROW(5)
            case 6: GGP_FRAG_ROW(6)
            case 7: GGP_FRAG_ROW(7)
            default: break;
          }
#undef GGP_FRAG_ROW
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        // refill the stage consumed in the previous k-step only after this step's first MMAs are in flight
        if (kk == 0) issue_load();
      }
    }

    gemm_epilogue<EPI, false>(p, wi, acc, sR, tid, wm, wn, g, q);
  }
  cp_async_wait<0>();
}

}  // namespace ggp
