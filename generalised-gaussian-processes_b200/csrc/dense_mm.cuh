// m x m section kernels: diagonal-block Cholesky + inverse, transposes, element-wise assembly, gemv, reductions.
// All matrices are row-major [Mp x Mp] (Mp = 64 * 2^j >= M, identity on the padding) unless noted.
#pragma once
#include "common.cuh"
#include "kernel_tiles.cuh"

namespace ggp {

constexpr int NB = 64;  // diagonal block size of the blocked Cholesky / recursive triangular inverse
constexpr int POTF2_SMEM = 0;

// Factor diagonal block kb in place (lower) and write its inverse T = L_kk^{-1}.  One CTA (16 x 16 threads) per batch element.
// Thread (tx,ty) owns the 4 x 4 elements (ty+16i, tx+16k) of A and of T (start: I) in registers.  The 64 columns are processed in
// 8 MICRO-BLOCKS of 8: per micro-block
//   A  the owners publish the 8 raw panel columns of A and the 8 current rows of T to shared memory                      (barrier)
//   B  warp 0 factors the 8 x 8 diagonal micro-block (every lane redundantly, in registers: no shuffles on the dependent chain)
//      and lanes 0..7 form its inverse D column by column                                                                (barrier)
//   C  all threads: panel rows L[r, g0:g0+8) = A_raw[r, g0:g0+8) D^T  and the final rows T[g0:g0+8, :] = D T_cur          (barrier)
//   D  all threads: rank-8 updates of their register tiles  A[r,c] -= L[r,:] . L[c,:] ,  T[r,:] -= L[r,:] T[g0:g0+8, :]
// 24 barriers and 8 dependent 8 x 8 factorisations instead of 64 barriers with a square root each on the chain (the per-column
// version took 25.7 us for 64 columns, ~750 clk per column; this one ~7 us).  Same arithmetic otherwise: right-looking, the
// inverse sweep rides in the same loop because the rows of T of a micro-block are final as soon as its columns of L are.
// info[b] = global index (1-based) of the first pivot <= piv_tol[b], LAPACK potrf style; first failure wins.
// piv_tol[b]: pivots at or below it count as "not positive definite" (0 = LAPACK semantics; see k_build_kzz).
constexpr int PF_MB = 8;   // micro-block width
// Device body (256 threads of one CTA, b = batch element): shared by k_potf2_trti2 and by the look-ahead tail of k_chol_trail.
__device__ __forceinline__ void potf2_trti2_block(double* __restrict__ A, int64_t ld, int64_t sA, int kb, double* __restrict__ T,
                                                  int64_t sT, int32_t* info, const double* __restrict__ piv_tol, int b,
                                                  long long* dbg) {
#define PF_STAMP(slot) do { if (dbg && tid == 0 && b == 0) dbg[slot] = clock64(); } while (0)
  __shared__ double Praw[NB][PF_MB + 1];     // raw panel columns (rows >= g0 used)
  __shared__ double Pnew[NB][PF_MB + 1];     // L[:, g0:g0+8) (0 above the diagonal)
  __shared__ double Tcur[PF_MB][NB + 1];     // current rows g0..g0+7 of T
  __shared__ double Tfin[PF_MB][NB + 1];     // final rows g0..g0+7 of T
  __shared__ double Ldd[PF_MB][PF_MB + 1], Dd[PF_MB][PF_MB + 1];
  __shared__ int bad;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
  double* Ab = A + b * sA + (int64_t)kb * NB * (ld + 1);
  const double tol = piv_tol ? piv_tol[b] : 0.0;
  double a[4][4], t[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = ty + 16 * i, c = tx + 16 * k;
      a[i][k] = (c <= r) ? __ldcg(Ab + (int64_t)r * ld + c) : 0.0;   // L2: the block may have been updated by another SM (k_chol_cluster)
      t[i][k] = (r == c) ? 1.0 : 0.0;
    }
  if (tid == 0) bad = 0;
  PF_STAMP(0);
  // owner tests are by VALUE (column / row index against g0) over the unrolled 4 x 4 register tile: no run-time register indexing, so
  // the micro-block loop may stay rolled (measured: 32 k clk per block warm) or be unrolled (26 k clk, 4.6 x the code): unrolled
#pragma unroll
  for (int jb = 0; jb < NB / PF_MB; ++jb) {
    const int g0 = jb * PF_MB;
    if (jb == 1) PF_STAMP(1);
    // ---- A: publish the raw panel columns g0..g0+7 and the current T rows g0..g0+7
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = tx + 16 * k;
      if (c >= g0 && c < g0 + PF_MB) {
#pragma unroll
        for (int i = 0; i < 4; ++i) Praw[ty + 16 * i][c - g0] = a[i][k];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 16 * i;
      if (r >= g0 && r < g0 + PF_MB) {
#pragma unroll
        for (int k = 0; k < 4; ++k) Tcur[r - g0][tx + 16 * k] = t[i][k];
      }
    }
    __syncthreads();
    if (jb == 1) PF_STAMP(2);
    // ---- B: 8 x 8 diagonal micro-block, warp 0
    if (tid < 32) {
      double v[PF_MB][PF_MB], inv[PF_MB];
#pragma unroll
      for (int r = 0; r < PF_MB; ++r)
#pragma unroll
        for (int c = 0; c < PF_MB; ++c) v[r][c] = (c <= r) ? Praw[g0 + r][c] : 0.0;
      int first_bad = 0;
#pragma unroll
      for (int j = 0; j < PF_MB; ++j) {
        const double dj = v[j][j];
        if (!(dj > tol) && first_bad == 0) first_bad = g0 + j + 1;
        inv[j] = rsqrt(dj);            // one reciprocal square root on the chain (<= 1 ulp): L_jj = d * inv, column scaled by inv
        v[j][j] = dj * inv[j];
#pragma unroll
        for (int r = j + 1; r < PF_MB; ++r) v[r][j] *= inv[j];
#pragma unroll
        for (int c = j + 1; c < PF_MB; ++c)
#pragma unroll
          for (int r = c; r < PF_MB; ++r) v[r][c] = fma(-v[r][j], v[c][j], v[r][c]);
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < PF_MB; ++r)
#pragma unroll
          for (int c = 0; c < PF_MB; ++c) Ldd[r][c] = (c <= r) ? v[r][c] : 0.0;
        if (first_bad && bad == 0) bad = first_bad;
      }
      if (lane < PF_MB) {   // column `lane` of D = L_dd^{-1} by forward substitution (1 / L_ii = inv[i])
        double x[PF_MB];
#pragma unroll
        for (int i = 0; i < PF_MB; ++i) {
          double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
          for (int k = 0; k < i; ++k) sacc = fma(-v[i][k], (k >= lane) ? x[k] : 0.0, sacc);
          x[i] = (i >= lane) ? sacc * inv[i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < PF_MB; ++i) Dd[i][lane] = x[i];
      }
    }
    if (jb == 1) PF_STAMP(3);
    __syncthreads();
    if (jb == 1) PF_STAMP(4);
    // ---- C: panel rows and final T rows
    {
      const int r = tid >> 2, q0 = (tid & 3) * 2;   // row r, outputs q0, q0 + 1
      double o0 = 0.0, o1 = 0.0;
      if (r >= g0 + PF_MB) {
#pragma unroll
        for (int pq = 0; pq < PF_MB; ++pq) {
          const double x = Praw[r][pq];
          o0 = fma(x, Dd[q0][pq], o0);         // D is lower triangular: entries with pq > q are zero
          o1 = fma(x, Dd[q0 + 1][pq], o1);
        }
      } else if (r >= g0) {
        o0 = Ldd[r - g0][q0];
        o1 = Ldd[r - g0][q0 + 1];
      }
      Pnew[r][q0] = o0;
      Pnew[r][q0 + 1] = o1;
      const int q = tid >> 5, c0 = (tid & 31) * 2;  // T row g0 + q, columns c0, c0 + 1
      double u0 = 0.0, u1 = 0.0;
#pragma unroll
      for (int pq = 0; pq < PF_MB; ++pq) {
        const double dq = Dd[q][pq];
        u0 = fma(dq, Tcur[pq][c0], u0);
        u1 = fma(dq, Tcur[pq][c0 + 1], u1);
      }
      Tfin[q][c0] = u0;
      Tfin[q][c0 + 1] = u1;
    }
    __syncthreads();
    if (jb == 1) PF_STAMP(5);
    // ---- D: rank-8 updates of the register tiles
    {
      double lr[4][PF_MB];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < PF_MB; ++q) lr[i][q] = Pnew[ty + 16 * i][q];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = tx + 16 * k;
        if (c >= g0 + PF_MB) {                    // trailing columns of A (lower part: r >= c)
          double lc[PF_MB];
#pragma unroll
          for (int q = 0; q < PF_MB; ++q) lc[q] = Pnew[c][q];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i >= k) {
#pragma unroll
              for (int q = 0; q < PF_MB; ++q) a[i][k] = fma(-lr[i][q], lc[q], a[i][k]);
            }
        } else {                                  // T[r, c], c < g0 + 8, for the rows below the micro-block
          double tc[PF_MB];
#pragma unroll
          for (int q = 0; q < PF_MB; ++q) tc[q] = Tfin[q][c];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (ty + 16 * i >= g0 + PF_MB) {
#pragma unroll
              for (int q = 0; q < PF_MB; ++q) t[i][k] = fma(-lr[i][q], tc[q], t[i][k]);
            }
          // final values of the finished columns of L go back into the register tile of their owner
          if (c >= g0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i][k] = Pnew[ty + 16 * i][c - g0];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {   // ... and the final rows of T
        const int r = ty + 16 * i;
        if (r >= g0 && r < g0 + PF_MB) {
#pragma unroll
          for (int k = 0; k < 4; ++k) t[i][k] = Tfin[r - g0][tx + 16 * k];
        }
      }
    }
    // the next publish writes Praw / Tcur, which nobody reads in phase D; Pnew / Tfin are rewritten only after two more barriers
    if (jb == 1) PF_STAMP(6);
  }
  PF_STAMP(7);
  double* Tb = T + b * sT + (int64_t)kb * NB * NB;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = ty + 16 * i, c = tx + 16 * k;
      Ab[(int64_t)r * ld + c] = (c <= r) ? a[i][k] : 0.0;
      Tb[r * NB + c] = (c <= r) ? t[i][k] : 0.0;
    }
  __syncthreads();
  PF_STAMP(8);
  if (tid == 0 && bad && info[b] == 0) info[b] = kb * NB + bad;
#undef PF_STAMP
}
__global__ void __launch_bounds__(256) k_potf2_trti2(double* __restrict__ A, int64_t ld, int64_t sA, int kb,
                                                     double* __restrict__ T, int64_t sT, int32_t* info,
                                                     const double* __restrict__ piv_tol, long long* dbg = nullptr) {
  potf2_trti2_block(A, ld, sA, kb, T, sT, info, piv_tol, blockIdx.x, dbg);
}

// One right-looking step of the blocked Cholesky after the diagonal block kb has been factored (k_potf2_trti2: T = L_kk^{-1}):
//   panel  L21 = A21 T^T                       (rem x NB,  rem = Mp - (kb + 1) NB)
//   update A22 -= L21 L21^T                    (lower NB x NB blocks only)
// fused: CTA (bj, bi), bi >= bj, recomputes the two panel blocks it needs (2 x NB^3 flops: cheaper than a kernel boundary), updates
// its block of A22 in place, and the diagonal CTAs also emit their panel block -- into the SCRATCH matrix S at the same coordinates,
// not into A: other CTAs of this launch still read the unscaled A21 (k_tril_merge moves the panels into A after the last step).
// Two library GEMM launches per step (in-place panel 19 us + K = 64 update 30 us, both latency-bound) become one ~10 us kernel.
// 256 threads = 8 warps, warp w owns 8 rows of each 64 x 64 x 64 product on the FP64 tensor pipe (a plain-FMA version with 4 x 4
// register tiles took 24-35 us: 8 shared loads per 16 FMAs at 8 warps per SM).
constexpr int CT_LD = NB + 4;                   // 68-double pitch: DMMA fragment loads (row g, k q -> bank (4 g + q) mod 16) are conflict-free
constexpr int CT_SMEM = 3 * NB * CT_LD * 8;
// acc[cb][0..1] += sum_k sa[8 w + g][k] * sb[8 cb + g'][k]: warp w owns the 8 rows 8 w .. 8 w + 7 of a 64 x 64 x 64 "NT" product on the
// FP64 tensor pipe (DMMA.8x8x4: lane = 4 g + q holds a = A[g][q], b = B[q][g], c = C[g][2 q .. 2 q + 1]); 9 shared loads per 8 DMMAs.
__device__ __forceinline__ void ct_mm_nt(const double* __restrict__ sa, const double* __restrict__ sb, int w, int lane, double (&acc)[8][2]) {
  const int g = lane >> 2, q = lane & 3;
  const double* pa = sa + (8 * w + g) * CT_LD + q;
  const double* pb = sb + g * CT_LD + q;
#pragma unroll 4
  for (int kk = 0; kk < NB / 4; ++kk) {
    const double a = pa[4 * kk];
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) dmma884(acc[cb][0], acc[cb][1], a, pb[cb * 8 * CT_LD + 4 * kk]);
  }
}
// LOOK-AHEAD (info != NULL): the CTA that owns the next diagonal block (bi = bj = 0) factors and inverts it right after its update,
// in this launch, while the other CTAs are still updating their blocks -- the dependent chain of a step is then ONE kernel
// (load, two products, update, 64-column factor + inverse) instead of two with a launch boundary between them, and the
// factorisation of block k + 1 overlaps the rest of the trailing update of step k.
// the same product when sb is LOWER triangular (sb[c][k] = 0 for k > c: the block inverse T): column block cb needs k-step kk only if
// 4 kk <= 8 cb + 7, i.e. cb >= kk / 2 -- compile-time bounds after unrolling, so the skipped DMMAs are not issued at all (half of them)
__device__ __forceinline__ void ct_mm_nt_lower_b(const double* __restrict__ sa, const double* __restrict__ sb, int w, int lane,
                                                 double (&acc)[8][2]) {
  const int g = lane >> 2, q = lane & 3;
  const double* pa = sa + (8 * w + g) * CT_LD + q;
  const double* pb = sb + g * CT_LD + q;
#pragma unroll
  for (int kk = 0; kk < NB / 4; ++kk) {
    const double a = pa[4 * kk];
#pragma unroll
    for (int cb = 0; cb < 8; ++cb)
      if (cb >= kk / 2) dmma884(acc[cb][0], acc[cb][1], a, pb[cb * 8 * CT_LD + 4 * kk]);
  }
}
// 64 x 64 block (row pitch ld doubles, 16-byte aligned rows) -> shared tile with the CT_LD pitch, by 16-byte LDGSTS: 8 per thread
__device__ __forceinline__ void ct_load_block_async(double* __restrict__ dst, const double* __restrict__ src, int64_t ld, int tid) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int e = tid + 256 * p, r = e >> 5, c2 = (e & 31) * 2;
    cp_async16(dst + r * CT_LD + c2, src + (int64_t)r * ld + c2, 16);
  }
}
__global__ void __launch_bounds__(256) k_chol_trail(double* __restrict__ A, int64_t ld, int64_t sA, int kb, double* __restrict__ T,
                                                    int64_t sT, double* __restrict__ S, int64_t sS, int32_t* info = nullptr,
                                                    const double* __restrict__ piv_tol = nullptr, long long* dbg = nullptr) {
  const int bj = blockIdx.x, bi = blockIdx.y, b = blockIdx.z;
  if (bi < bj) return;
  // developer timeline (GGP_CHOL_TIMELINE): stamps of the CTA on the dependent chain, 8 slots per step
#define CT_STAMP(slot, v) do { if (dbg && threadIdx.x == 0 && bi == 0 && bj == 0 && b == 0) dbg[kb * 8 + slot] = (v); } while (0)
  if (dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); CT_STAMP(0, (long long)gt); CT_STAMP(1, clock64()); }
  extern __shared__ __align__(16) unsigned char ct_raw[];
  double* sAi = reinterpret_cast<double*>(ct_raw);   // A21 block bi, then P_i
  double* sAj = sAi + NB * CT_LD;                    // A21 block bj, then P_j
  double* sTk = sAj + NB * CT_LD;                    // T
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int k0 = kb * NB, r0 = k0 + NB;
  double* Ab = A + b * sA;
  const double* Tb = T + b * sT + (int64_t)kb * NB * NB;
  // (k_potf2_trti2 writes T with explicit zeros above the diagonal, so the block can be copied as it is)
  ct_load_block_async(sAi, Ab + (int64_t)(r0 + bi * NB) * ld + k0, ld, tid);
  if (bi != bj) ct_load_block_async(sAj, Ab + (int64_t)(r0 + bj * NB) * ld + k0, ld, tid);
  ct_load_block_async(sTk, Tb, NB, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  CT_STAMP(2, clock64());
  double pi[8][2], pj[8][2];
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) pi[cb][0] = pi[cb][1] = pj[cb][0] = pj[cb][1] = 0.0;
  ct_mm_nt_lower_b(sAi, sTk, w, lane, pi);      // P_i[r][c] = sum_{k <= c} A21_i[r][k] T[c][k]
  if (bi != bj) ct_mm_nt_lower_b(sAj, sTk, w, lane, pj);
  __syncthreads();                              // everyone is done reading A21_i / A21_j
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    *reinterpret_cast<double2*>(sAi + (8 * w + g) * CT_LD + 8 * cb + 2 * q) = make_double2(pi[cb][0], pi[cb][1]);
    if (bi != bj) *reinterpret_cast<double2*>(sAj + (8 * w + g) * CT_LD + 8 * cb + 2 * q) = make_double2(pj[cb][0], pj[cb][1]);
  }
  if (bi == bj) {   // this CTA owns panel block bi
    double* Sb = S + b * sS + (int64_t)(r0 + bi * NB + 8 * w + g) * ld + k0 + 2 * q;
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) *reinterpret_cast<double2*>(Sb + 8 * cb) = make_double2(pi[cb][0], pi[cb][1]);
  }
  __syncthreads();
  CT_STAMP(3, clock64());
  double c[8][2];
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) c[cb][0] = c[cb][1] = 0.0;
  ct_mm_nt(sAi, bi != bj ? sAj : sAi, w, lane, c);   // P_i P_j^T
  double* dst = Ab + (int64_t)(r0 + bi * NB + 8 * w + g) * ld + r0 + bj * NB + 2 * q;
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    double2 o = *reinterpret_cast<double2*>(dst + 8 * cb);
    o.x -= c[cb][0];
    o.y -= c[cb][1];
    *reinterpret_cast<double2*>(dst + 8 * cb) = o;
  }
  if (info && bi == 0 && bj == 0) {
    __syncthreads();   // the updated diagonal block (global memory, written by this CTA) is visible to all its threads
    CT_STAMP(4, clock64());
    potf2_trti2_block(A, ld, sA, kb + 1, T, sT, info, piv_tol, b, nullptr);
    CT_STAMP(5, clock64());
    if (dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); CT_STAMP(6, (long long)gt); }
  }
#undef CT_STAMP
}
// ---- Cluster-resident blocked Cholesky (batch = 1): the WHOLE factorisation in one launch on CC_N SMs -----------------------------------
// Why: the multi-launch plan is a chain of 16 dependent launches whose CTAs need an SM to themselves (250 registers x 256 threads), so next
// to the tile build of pass 1 -- which keeps every SM filled with its own CTAs -- the chain does not run concurrently but after it: the
// Kzz factorisation sits on the critical path of every evaluation although nothing depends on it until the triangular multiply.  One
// thread-block cluster, launched on a high-priority stream, takes its CC_N SMs once and keeps them; the build runs on the others.
// Schedule per step k (diagonal block k factored, T_k = L_kk^-1 known):
//   panels   L(i,k) = A(i,k) T_k^T, i = k+1+rank, +CC_N, ...  (in place)                                  | cluster barrier
//   updates  A(i,j) -= L(i,k) L(j,k)^T over the trailing lower blocks: rank 0 takes the next diagonal block first and factors it
//            (look-ahead), every CTA then draws blocks from an atomic counter (each block is updated by exactly one CTA per step with
//            the same arithmetic whoever draws it: deterministic)                                         | cluster barrier
// All operand loads are L2 loads (LDGSTS .cg / ld.global.cg): the blocks are written by other SMs between the barriers.
constexpr int CC_N = 8;
__device__ __forceinline__ void cc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// After the factorisation the same launch forms the explicit inverse by recursive doubling on 64 x 64 blocks (what
// chol_and_inverse_launches does with a merge kernel and 8 products): Linv = L^-1, LinvT = L^-T, strict upper triangle of A zeroed.
// Level s (blocks), pair p with first block b0 = 2 s p:   W = L21 Inv11 (stored transposed in Wk),   X = -Inv22 W -> Linv, X^T -> LinvT;
// one work item = one output block with its k loop in registers, items drawn heaviest first from an atomic counter, a cluster
// barrier after each of the two products of a level.
__global__ void __launch_bounds__(256) k_chol_cluster(double* __restrict__ A, int Mp, double* __restrict__ T, int32_t* info,
                                                      const double* __restrict__ piv_tol, int* __restrict__ ctr,
                                                      double* __restrict__ Linv, double* __restrict__ LinvT, double* __restrict__ Wk) {
  extern __shared__ __align__(16) unsigned char ct_raw[];
  double* s0 = reinterpret_cast<double*>(ct_raw);
  double* s1 = s0 + NB * CT_LD;
  double* s2 = s1 + NB * CT_LD;
  __shared__ int s_w;
  const int rank = blockIdx.x, nrank = gridDim.x, tid = threadIdx.x, w8 = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int nblk = Mp / NB;
  const int64_t ld = Mp;
  if (rank == 0) potf2_trti2_block(A, ld, 0, 0, T, 0, info, piv_tol, 0, nullptr);
  cc_cluster_sync();
  // A(i,j) -= L(i,k) L(j,k)^T for one 64 x 64 block
  auto update_block = [&](int i, int j, int k) {
    ct_load_block_async(s0, A + (int64_t)i * NB * ld + k * NB, ld, tid);
    if (i != j) ct_load_block_async(s1, A + (int64_t)j * NB * ld + k * NB, ld, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double c[8][2];
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) c[cb][0] = c[cb][1] = 0.0;
    ct_mm_nt(s0, i != j ? s1 : s0, w8, lane, c);
    double* dst = A + (int64_t)(i * NB + 8 * w8 + g) * ld + j * NB + 2 * q;
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
      double2 o = __ldcg(reinterpret_cast<const double2*>(dst + 8 * cb));
      o.x -= c[cb][0];
      o.y -= c[cb][1];
      *reinterpret_cast<double2*>(dst + 8 * cb) = o;
    }
    __syncthreads();   // s0 / s1 are free again, the block is written
  };
  for (int k = 0; k + 1 < nblk; ++k) {
    // ---- panels of column k
    if (k + 1 + rank < nblk) {
      ct_load_block_async(s2, T + (int64_t)k * NB * NB, NB, tid);
      for (int i = k + 1 + rank; i < nblk; i += nrank) {
        double* blk = A + (int64_t)i * NB * ld + k * NB;
        ct_load_block_async(s0, blk, ld, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        double pi[8][2];
#pragma unroll
        for (int cb = 0; cb < 8; ++cb) pi[cb][0] = pi[cb][1] = 0.0;
        ct_mm_nt_lower_b(s0, s2, w8, lane, pi);
#pragma unroll
        for (int cb = 0; cb < 8; ++cb)
          *reinterpret_cast<double2*>(blk + (int64_t)(8 * w8 + g) * ld + 8 * cb + 2 * q) = make_double2(pi[cb][0], pi[cb][1]);
        __syncthreads();
      }
    }
    cc_cluster_sync();
    // ---- trailing update; look-ahead factor of the next diagonal block by rank 0
    const int nb = nblk - k - 1, ntrail = nb * (nb + 1) / 2;
    if (rank == 0) {
      update_block(k + 1, k + 1, k);
      potf2_trti2_block(A, ld, 0, k + 1, T, 0, info, piv_tol, 0, nullptr);
      __syncthreads();
    }
    for (;;) {
      if (tid == 0) s_w = atomicAdd(ctr + k, 1) + 1;   // block 0 (the next diagonal block) belongs to rank 0
      __syncthreads();
      const int w = s_w;
      __syncthreads();
      if (w >= ntrail) break;
      int bi = 0;
      while ((bi + 1) * (bi + 2) / 2 <= w) ++bi;
      const int bj = w - bi * (bi + 1) / 2;
      update_block(k + 1 + bi, k + 1 + bj, k);
    }
    cc_cluster_sync();
  }
  if (!Linv) return;
  // ---- explicit inverse.  Start: Linv = blockdiag(T_k), LinvT = blockdiag(T_k^T), zero elsewhere; A: strict upper triangle zero
  for (int64_t e = (int64_t)rank * 256 + tid; e < (int64_t)Mp * Mp; e += (int64_t)nrank * 256) {
    const int i = (int)(e / Mp), j = (int)(e % Mp);
    if (j > i) A[e] = 0.0;
    double v = 0.0, vt = 0.0;
    if (i / NB == j / NB) {
      const double* Tb = T + (int64_t)(i / NB) * NB * NB;
      v = __ldcg(Tb + (i % NB) * NB + (j % NB));
      vt = __ldcg(Tb + (j % NB) * NB + (i % NB));
    }
    Linv[e] = v;
    LinvT[e] = vt;
  }
  cc_cluster_sync();
  auto block = [&](double* base, int bi, int bj) { return base + (int64_t)bi * NB * ld + (int64_t)bj * NB; };
  int lvl = 0;
  for (int sb = 1; sb < nblk; sb *= 2, ++lvl) {
    const int npairs = nblk / (2 * sb), items = npairs * sb * sb;
    for (int phase = 0; phase < 2; ++phase) {
      int* c = ctr + 64 + 2 * lvl + phase;   // ctr[0 .. nblk - 2]: the steps of the factorisation (nblk <= 64)
      for (;;) {
        if (tid == 0) s_w = atomicAdd(c, 1);
        __syncthreads();
        const int w = s_w;
        __syncthreads();
        if (w >= items) break;
        // heaviest first: phase 0 has sb - j terms (j ascending), phase 1 has i + 1 terms (i descending)
        const int pr = w % npairs, r = w / npairs, b0 = 2 * sb * pr;
        const int i = phase == 0 ? r % sb : sb - 1 - r / sb, j = phase == 0 ? r / sb : r % sb;
        double acc[8][2];
#pragma unroll
        for (int cb = 0; cb < 8; ++cb) acc[cb][0] = acc[cb][1] = 0.0;
        const int k_lo = phase == 0 ? j : 0, k_hi = phase == 0 ? sb - 1 : i;
        for (int k = k_lo; k <= k_hi; ++k) {
          // phase 0:  W(i,j)  += L21(i,k) Inv11(k,j)      sa = L(b0+sb+i, b0+k),          sb[c][kk] = LinvT(b0+j, b0+k)
          // phase 1:  X(i,j)  += Inv22(i,k) W(k,j)        sa = Linv(b0+sb+i, b0+sb+k),    sb[c][kk] = WkT(b0+j, b0+sb+k)
          ct_load_block_async(s0, phase == 0 ? block(A, b0 + sb + i, b0 + k) : block(Linv, b0 + sb + i, b0 + sb + k), ld, tid);
          ct_load_block_async(s1, phase == 0 ? block(LinvT, b0 + j, b0 + k) : block(Wk, b0 + j, b0 + sb + k), ld, tid);
          cp_async_commit();
          cp_async_wait<0>();
          __syncthreads();
          ct_mm_nt(s0, s1, w8, lane, acc);
          __syncthreads();
        }
        if (phase == 0) {   // W(i,j) transposed into Wk block (b0 + j, b0 + sb + i)
          double* dst = block(Wk, b0 + j, b0 + sb + i);
#pragma unroll
          for (int cb = 0; cb < 8; ++cb) {
            dst[(int64_t)(8 * cb + 2 * q) * ld + 8 * w8 + g] = acc[cb][0];
            dst[(int64_t)(8 * cb + 2 * q + 1) * ld + 8 * w8 + g] = acc[cb][1];
          }
        } else {            // X = -acc -> Linv block (b0 + sb + i, b0 + j), X^T -> LinvT block (b0 + j, b0 + sb + i)
          double* dx = block(Linv, b0 + sb + i, b0 + j) + (int64_t)(8 * w8 + g) * ld + 2 * q;
          double* dt = block(LinvT, b0 + j, b0 + sb + i);
#pragma unroll
          for (int cb = 0; cb < 8; ++cb) {
            *reinterpret_cast<double2*>(dx + 8 * cb) = make_double2(-acc[cb][0], -acc[cb][1]);
            dt[(int64_t)(8 * cb + 2 * q) * ld + 8 * w8 + g] = -acc[cb][0];
            dt[(int64_t)(8 * cb + 2 * q + 1) * ld + 8 * w8 + g] = -acc[cb][1];
          }
        }
      }
      cc_cluster_sync();
    }
  }
}

// Whole factorisation + explicit inverse for Mp <= 128 (one or two diagonal blocks) in ONE launch, one CTA per batch element:
//   L00, T0 = potf2(A00);  L10 = A10 T0^T;  L11, T1 = potf2(A11 - L10 L10^T);  Linv = [[T0, 0], [-T1 L10 T0, T1]];  LinvT = Linv^T
// -- what chol_and_inverse_launches does with 8 dependent launches at this size (2 x potf2, trail, merge, blockdiag, transpose,
// 2 products, transpose), each of them latency-bound.  This is the m x m section of the reference's own problem sizes (co2:
// M = 100, demo: M = 20), evaluated twice per leapfrog step of every HMC chain.  Same arithmetic as the multi-launch plan
// (DMMA 64^3 products, the same diagonal-block kernel body), so the two agree to rounding.
__global__ void __launch_bounds__(256) k_chol_inv_small(double* __restrict__ A, int Mp, int64_t sA, double* __restrict__ T, int64_t sT,
                                                        double* __restrict__ Linv, double* __restrict__ LinvT, int64_t sL,
                                                        int32_t* info, const double* __restrict__ piv_tol) {
  extern __shared__ __align__(16) unsigned char ct_raw[];
  double* s0 = reinterpret_cast<double*>(ct_raw);
  double* s1 = s0 + NB * CT_LD;
  double* s2 = s1 + NB * CT_LD;
  const int b = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  double* Ab = A + b * sA;
  double* Li = Linv + b * sL;
  double* LiT = LinvT + b * sL;
  const double* Tb = T + b * sT;
  potf2_trti2_block(A, Mp, sA, 0, T, sT, info, piv_tol, b, nullptr);
  __syncthreads();
  if (Mp == NB) {
    for (int e = tid; e < NB * NB; e += 256) {
      const int r = e / NB, c = e % NB;
      const double v = (c <= r) ? Tb[r * NB + c] : 0.0;
      Li[(int64_t)r * Mp + c] = v;
      LiT[(int64_t)c * Mp + r] = v;
    }
    return;
  }
  // ---- Mp == 2 NB
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    s0[r * CT_LD + c] = Ab[(int64_t)(NB + r) * Mp + c];          // A10
    s2[r * CT_LD + c] = (c <= r) ? Tb[r * NB + c] : 0.0;         // T0
  }
  __syncthreads();
  double acc[8][2];
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) acc[cb][0] = acc[cb][1] = 0.0;
  ct_mm_nt_lower_b(s0, s2, w, lane, acc);                        // L10 = A10 T0^T
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    const double2 v = make_double2(acc[cb][0], acc[cb][1]);
    *reinterpret_cast<double2*>(s1 + (8 * w + g) * CT_LD + 8 * cb + 2 * q) = v;
    *reinterpret_cast<double2*>(Ab + (int64_t)(NB + 8 * w + g) * Mp + 8 * cb + 2 * q) = v;
    *reinterpret_cast<double2*>(Ab + (int64_t)(8 * w + g) * Mp + NB + 8 * cb + 2 * q) = make_double2(0.0, 0.0);   // A01 = 0
    acc[cb][0] = acc[cb][1] = 0.0;
  }
  __syncthreads();
  ct_mm_nt(s1, s1, w, lane, acc);                                // A11 -= L10 L10^T
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    double2* dst = reinterpret_cast<double2*>(Ab + (int64_t)(NB + 8 * w + g) * Mp + NB + 8 * cb + 2 * q);
    double2 o = *dst;
    o.x -= acc[cb][0];
    o.y -= acc[cb][1];
    *dst = o;
  }
  __syncthreads();
  potf2_trti2_block(A, Mp, sA, 1, T, sT, info, piv_tol, b, nullptr);
  __syncthreads();
  // W = L10 T0 (operand "B" of the NT product = T0^T), then X = T1 W (operand "B" = W^T)
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    s2[c * CT_LD + r] = (c <= r) ? Tb[r * NB + c] : 0.0;                   // s2[c][k] = T0[k][c]
    s0[r * CT_LD + c] = (c <= r) ? Tb[NB * NB + r * NB + c] : 0.0;         // T1
  }
  __syncthreads();
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) acc[cb][0] = acc[cb][1] = 0.0;
  ct_mm_nt(s1, s2, w, lane, acc);
  __syncthreads();                                               // everyone is done reading s2 (T0^T)
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    s2[(8 * cb + 2 * q) * CT_LD + 8 * w + g] = acc[cb][0];       // s2[c][k] = W[k][c]
    s2[(8 * cb + 2 * q + 1) * CT_LD + 8 * w + g] = acc[cb][1];
    acc[cb][0] = acc[cb][1] = 0.0;
  }
  __syncthreads();
  ct_mm_nt(s0, s2, w, lane, acc);
  __syncthreads();
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {                               // s1 = X = -T1 L10 T0 (staged for the coalesced output below)
    s1[(8 * w + g) * CT_LD + 8 * cb + 2 * q] = -acc[cb][0];
    s1[(8 * w + g) * CT_LD + 8 * cb + 2 * q + 1] = -acc[cb][1];
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    const double t0 = (c <= r) ? Tb[r * NB + c] : 0.0, t1 = s0[r * CT_LD + c];
    Li[(int64_t)r * Mp + c] = t0;
    Li[(int64_t)r * Mp + NB + c] = 0.0;
    Li[(int64_t)(NB + r) * Mp + c] = s1[r * CT_LD + c];
    Li[(int64_t)(NB + r) * Mp + NB + c] = t1;
    // transposed copy, written row by row of LinvT: LinvT[r][c] = Linv[c][r]
    LiT[(int64_t)r * Mp + c] = (r <= c) ? Tb[c * NB + r] : 0.0;
    LiT[(int64_t)r * Mp + NB + c] = s1[c * CT_LD + r];
    LiT[(int64_t)(NB + r) * Mp + c] = 0.0;
    LiT[(int64_t)(NB + r) * Mp + NB + c] = s0[c * CT_LD + r];
  }
}

// after the last step: A = [diagonal blocks of A (lower part)] + [panel blocks from S], strict upper triangle zero
// Optionally (T != NULL) the same launch starts the triangular inverse: Linv = blockdiag(T_0, T_1, ...), LinvT = its transpose
// (what k_init_blockdiag does on its own).
__global__ void k_tril_merge(double* __restrict__ A, const double* __restrict__ S, int Mp, int64_t sA, int64_t sS,
                             const double* __restrict__ T = nullptr, int64_t sT = 0, double* __restrict__ Linv = nullptr,
                             double* __restrict__ LinvT = nullptr) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= Mp || j >= Mp) return;
  double* a = A + blockIdx.z * sA + (int64_t)i * Mp + j;
  if (j > i) *a = 0.0;
  else if (i / NB != j / NB) *a = S[blockIdx.z * sS + (int64_t)i * Mp + j];
  if (T) {
    double v = 0.0, vt = 0.0;
    if (i / NB == j / NB) {
      const double* Tb = T + blockIdx.z * sT + (int64_t)(i / NB) * NB * NB;
      v = Tb[(i % NB) * NB + (j % NB)];
      vt = Tb[(j % NB) * NB + (i % NB)];
    }
    Linv[blockIdx.z * sA + (int64_t)i * Mp + j] = v;
    LinvT[blockIdx.z * sA + (int64_t)i * Mp + j] = vt;
  }
}

// zero the strict upper triangle.  grid (Mp/16, Mp/16, batch), block (16,16)
__global__ void k_tril(double* __restrict__ A, int Mp, int64_t sA) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i < Mp && j < Mp && j > i) A[blockIdx.z * sA + (int64_t)i * Mp + j] = 0.0;
}

// Linv = blockdiag(T_0, T_1, ...), zero elsewhere; LinvT (optional) = its transpose
__global__ void k_init_blockdiag(double* __restrict__ Linv, int Mp, int64_t sL, const double* __restrict__ T, int64_t sT,
                                 double* __restrict__ LinvT = nullptr) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  if (i >= Mp || j >= Mp) return;
  const int bi = i / NB, bj = j / NB;
  double v = 0.0, vt = 0.0;
  if (bi == bj) {
    const double* Tb = T + blockIdx.z * sT + (int64_t)bi * NB * NB;
    v = Tb[(i % NB) * NB + (j % NB)];
    vt = Tb[(j % NB) * NB + (i % NB)];
  }
  Linv[blockIdx.z * sL + (int64_t)i * Mp + j] = v;
  if (LinvT) LinvT[blockIdx.z * sL + (int64_t)i * Mp + j] = vt;
}

// out = in^T  (32x32 tiles through shared memory).  grid (Mp/32, Mp/32, batch), block (32, 8)
__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, int Mp, int64_t s) {
  __shared__ double tile[32][33];
  const double* I = in + blockIdx.z * s;
  double* O = out + blockIdx.z * s;
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8)
    if (x < Mp && y0 + r < Mp) tile[r][threadIdx.x] = I[(int64_t)(y0 + r) * Mp + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8)
    if (xo < Mp && yo0 + r < Mp) O[(int64_t)(yo0 + r) * Mp + xo] = tile[threadIdx.x][r];
}

// Bm = I + S/s on the leading M x M (S dense, ld = M, from `partial`), identity on the padding
__global__ void k_make_B(const double* __restrict__ partial, int64_t sP, int M, int Mp, const double* __restrict__ theta,
                         int d, double* __restrict__ Bm, int64_t sB) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x, b = blockIdx.z;
  if (i >= Mp || j >= Mp) return;
  const double s2 = theta[(int64_t)b * (d + 2) + d + 1];
  double v = (i == j) ? 1.0 : 0.0;
  if (i < M && j < M) v += partial[b * sP + (int64_t)i * M + j] / s2;
  Bm[b * sB + (int64_t)i * Mp + j] = v;
}

// y[i] = alpha * s2^(-spow) * sum_j A[i,j] x[j]   (one warp per row; s2 = noise of batch b)  grid (ceil(M/8), batch), block 256
__global__ void __launch_bounds__(256) k_gemv(const double* __restrict__ A, int64_t ld, int64_t sA,
                                              const double* __restrict__ x, int64_t sx, double* __restrict__ y,
                                              int64_t sy, int M, int ncols, double alpha,
                                              const double* __restrict__ theta, int d, int spow) {
  const int b = blockIdx.y, row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const double* a = A + b * sA + (int64_t)row * ld;
  const double* xv = x + b * sx;
  double s = 0.0;
  for (int j = lane; j < ncols; j += 32) s = fma(a[j], xv[j], s);
  s = warp_sum(s);
  if (lane == 0) {
    double sc = alpha;
    if (spow > 0) {
      const double s2 = theta[(int64_t)b * (d + 2) + d + 1];
      sc = (spow == 1) ? alpha / s2 : alpha / (s2 * s2);
    }
    y[b * sy + row] = sc * s;
  }
}
// y[i] += sum_j A[i,j] x[j]  (accumulating variant used for b += A_chunk y_chunk; x has batch stride sx, 0 = shared)
__global__ void __launch_bounds__(256) k_gemv_acc(const double* __restrict__ A, int64_t ld, int64_t sA,
                                                  const double* __restrict__ x, int64_t sx, double* __restrict__ y, int64_t sy,
                                                  int M, int ncols) {
  const int b = blockIdx.y, row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const double* a = A + b * sA + (int64_t)row * ld;
  x += b * sx;
  double s = 0.0;
  for (int j = lane; j < ncols; j += 32) s = fma(a[j], x[j], s);
  s = warp_sum(s);
  if (lane == 0) y[b * sy + row] += s;
}

// Eval-mode training covariance of the sparse predictive (models/sgpr.py:150-160: ExactGP.__call__ evaluates the InducingPointKernel
// on the training inputs in eval mode, so the sgpr diagonal correction clamp(k_nn - q_nn, 0) is added to the TRAINING rows as well):
// Lambda_n = s2 + max(sf2 - ||a_n||^2, 0).  Scales column n of the chunk A^T[m x nv] (row i at At + i * ld) and y_n by sqrt(s2 / Lambda_n),
// so that the ordinary pass-1 sums become s2 * A W A^T and s2 * A W y (W = diag(1 / Lambda)) and ggp_sgpr_finish yields
// B = I + A W A^T, c = L_B^{-1} A W y unchanged.  One thread per column; grid (ceil(nv / 256), batch).
__global__ void __launch_bounds__(256) k_fitc_scale(double* __restrict__ At, int64_t ld, int64_t sA, int M, int nv,
                                                    const double* __restrict__ y, const double* __restrict__ theta, int d,
                                                    double* __restrict__ ysc, int64_t sy) {
  const int n = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (n >= nv) return;
  const double sf2 = theta[(int64_t)b * (d + 2) + d], s2 = theta[(int64_t)b * (d + 2) + d + 1];
  double* a = At + b * sA + n;
  double q = 0.0;
  for (int i = 0; i < M; ++i) { const double v = a[(int64_t)i * ld]; q = fma(v, v, q); }
  const double w = sqrt(s2 / (s2 + fmax(sf2 - q, 0.0)));
  for (int i = 0; i < M; ++i) a[(int64_t)i * ld] *= w;
  ysc[b * sy + n] = y[n] * w;
}

// out[0] = sum_i y_i^2 over n, deterministic: SUMSQ_BLOCKS block partials (out[1 + block], contiguous slices) then a fixed-order sum
constexpr int SUMSQ_BLOCKS = 16;
__global__ void __launch_bounds__(1024) k_sumsq(const double* __restrict__ y, int64_t n, double* __restrict__ out) {
  __shared__ double red[32];
  const int64_t per = (n + SUMSQ_BLOCKS - 1) / SUMSQ_BLOCKS, lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  double s = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 1024) s = fma(y[i], y[i], s);
  s = block_sum<1024>(s, red);
  if (threadIdx.x == 0) out[1 + blockIdx.x] = s;
}
__global__ void k_sumsq_final(double* __restrict__ out) {
  double s = 0.0;
  for (int i = 0; i < SUMSQ_BLOCKS; ++i) s += out[1 + i];
  out[0] = s;
}

// partial[b] = [ S (upper tiles summed over splits, mirrored) | bvec | yty, n*sf2, n ]
// grid (ceil(M/16), ceil(M/16), batch), block (16,16)
__global__ void k_finalize_partial(const double* __restrict__ Spart, int64_t ldS, int64_t sSplit, int64_t sSb, int splits,
                                   const double* __restrict__ bvec, int64_t sbv, const double* __restrict__ yty, int64_t n_local,
                                   const double* __restrict__ theta, int d, int M, double* __restrict__ partial, int64_t sP) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x, b = blockIdx.z;
  double* out = partial + b * sP;
  if (i < M && j < M) {
    // tiles with tile_j >= tile_i were computed: always read the upper element, so S is exactly symmetric
    const int r = i < j ? i : j, c = i < j ? j : i;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += Spart[b * sSb + k * sSplit + (int64_t)r * ldS + c];
    out[(int64_t)i * M + j] = s;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    const int t = threadIdx.y * 16 + threadIdx.x;
    for (int k = t; k < M; k += 256) out[(int64_t)M * M + k] = bvec[b * sbv + k];
    if (t == 0) {
      out[(int64_t)M * M + M + 0] = yty[0];
      out[(int64_t)M * M + M + 1] = (double)n_local * theta[(int64_t)b * (d + 2) + d];
      out[(int64_t)M * M + M + 2] = (double)n_local;
    }
  }
}

// scalars of the bound and of dF/ds2.  One CTA (256 thr) per batch element.
//   c = LBinv b / s ; beta = Binv b     (given)
//   F = -N/2 log2pi - N/2 log s - sum log diag L_B - (yty/s - c.c)/2 - (sum_knn - trS)/(2s)
//   dF/ds = -N/(2s) + (M - tr Binv)/(2s) + yty/(2s^2) - b.beta/s^3 + (b.beta - beta.beta)/(2 s^3) + (sum_knn - trS)/(2s^2)
__global__ void __launch_bounds__(256) k_bound_scalars(const double* __restrict__ partial, int64_t sP, int M, int Mp,
                                                       const double* __restrict__ theta, int d,
                                                       const double* __restrict__ LB, const double* __restrict__ Binv,
                                                       int64_t sMat, const double* __restrict__ cvec,
                                                       const double* __restrict__ beta, int64_t sv,
                                                       double* __restrict__ bound, double* __restrict__ ds2_out) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* P = partial + b * sP;
  const double* bv = P + (int64_t)M * M;
  const double s2 = theta[(int64_t)b * (d + 2) + d + 1];
  double logdiag = 0, cc = 0, trS = 0, trBinv = 0, bbeta = 0, betabeta = 0;
  for (int i = tid; i < M; i += 256) {
    logdiag += log(LB[b * sMat + (int64_t)i * Mp + i]);
    const double ci = cvec[b * sv + i], be = beta[b * sv + i];
    cc = fma(ci, ci, cc);
    trS += P[(int64_t)i * M + i];
    trBinv += Binv[b * sMat + (int64_t)i * Mp + i];
    bbeta = fma(bv[i], be, bbeta);
    betabeta = fma(be, be, betabeta);
  }
  logdiag = block_sum<256>(logdiag, red);
  cc = block_sum<256>(cc, red);
  trS = block_sum<256>(trS, red);
  trBinv = block_sum<256>(trBinv, red);
  bbeta = block_sum<256>(bbeta, red);
  betabeta = block_sum<256>(betabeta, red);
  if (tid == 0) {
    const double yty = bv[M], sumk = bv[M + 1], N = bv[M + 2];
    const double LOG2PI = 1.8378770664093453;
    bound[b] = -0.5 * N * LOG2PI - 0.5 * N * log(s2) - logdiag - 0.5 * (yty / s2 - cc) - 0.5 * (sumk - trS) / s2;
    if (ds2_out) {
      const double s3 = s2 * s2 * s2;
      ds2_out[b] = -0.5 * N / s2 + 0.5 * ((double)M - trBinv) / s2 + 0.5 * yty / (s2 * s2) - bbeta / s3 +
                   0.5 * (bbeta - betabeta) / s3 + 0.5 * (sumk - trS) / (s2 * s2);
    }
  }
}

// PA = (I - Binv)/s - beta beta^T / s^3 ;  Gbar = (I + S/s) + Binv - 2I + beta beta^T / s^2   (zero on the padding)
__global__ void k_make_PA_Gbar(const double* __restrict__ partial, int64_t sP, int M, int Mp,
                               const double* __restrict__ theta, int d, const double* __restrict__ Binv,
                               const double* __restrict__ beta, int64_t sv, double* __restrict__ PA,
                               double* __restrict__ Gbar, int64_t sMat) {
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x, b = blockIdx.z;
  if (i >= Mp || j >= Mp) return;
  double pa = 0.0, gb = 0.0;
  if (i < M && j < M) {
    const double s2 = theta[(int64_t)b * (d + 2) + d + 1];
    const double bi = Binv[b * sMat + (int64_t)i * Mp + j];
    const double bb = beta[b * sv + i] * beta[b * sv + j];
    const double eye = (i == j) ? 1.0 : 0.0;
    pa = (eye - bi) / s2 - bb / (s2 * s2 * s2);
    gb = partial[b * sP + (int64_t)i * M + j] / s2 + bi - eye + bb / (s2 * s2);
  }
  PA[b * sMat + (int64_t)i * Mp + j] = pa;
  Gbar[b * sMat + (int64_t)i * Mp + j] = gb;
}

// Kzz-dependent gradient, one warp per row i:  V = Gzz o Kzz ;
//   rowacc[i][c] = sum_j Gzz_ij g_ij * d(d2)/d ell_c ;  rowacc[i][d] = sum_j Gzz_ij k_ij ;
//   dZ[i][c] = 2 * sum_j Gzz_ij g_ij * d(d2)/d z_ic                     (g = dk/d(d2); Gzz symmetric)
// Dimensions are processed in register blocks of 8 (one kernel evaluation per (j, block)).
// grid (ceil(M/8), batch), block 256
__global__ void __launch_bounds__(256) k_grad_kzz_rows(const double* __restrict__ Gzz, int Mp, int64_t sMat,
                                                       const double* __restrict__ Z, int M, int d,
                                                       const double* __restrict__ theta, KSpec kind,
                                                       double* __restrict__ rowacc /*[batch][M][d+1]*/,
                                                       double* __restrict__ dZ /*[batch][M][d], stride sG*/, int64_t sG) {
  const int b = blockIdx.y, i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= M) return;
  const double* th = theta + (int64_t)b * (d + 2);
  const double sf2 = th[d];
  const double* zi = Z + (int64_t)i * d;
  const double* grow = Gzz + b * sMat + (int64_t)i * Mp;
  for (int c0 = 0; c0 < d; c0 += 8) {
    double al[8], az[8], ksum = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) al[c] = az[c] = 0.0;
    for (int j = lane; j < M; j += 32) {
      const double* zj = Z + (int64_t)j * d;
      double d2 = 0.0;
      for (int cc = 0; cc < d; ++cc) {
        const double t = (zi[cc] - zj[cc]) / th[cc];
        d2 = fma(t, t, d2);
      }
      const double gij = grow[j];
      if (c0 == 0) ksum = fma(gij, kval(kind, sf2, d2), ksum);
      const double gg = gij * kgrad(kind, sf2, d2);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c0 + c < d) {
          const double df = zi[c0 + c] - zj[c0 + c];
          al[c] = fma(gg * df, df, al[c]);
          az[c] = fma(gg, df, az[c]);
        }
      }
    }
    if (c0 == 0) {
      ksum = warp_sum(ksum);
      if (lane == 0) rowacc[((int64_t)b * M + i) * (d + 1) + d] = ksum;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c0 + c < d) {
        const double sl = warp_sum(al[c]), sz = warp_sum(az[c]);
        if (lane == 0) {
          const double e = th[c0 + c];
          // d(d2)/d ell_c = -2 df^2 / ell_c^3 ; d(d2)/d z_ic = 2 df / ell_c^2 (and the symmetric partner doubles it)
          rowacc[((int64_t)b * M + i) * (d + 1) + c0 + c] = -2.0 * sl / (e * e * e);
          dZ[b * sG + (int64_t)i * d + c0 + c] = 4.0 * sz / (e * e);
        }
      }
    }
  }
}

// grad_mm[b] = [ d_ell (sum rows) , d_sf2 = (sum_i rowacc[i][d] + rk) / sf2 - N/(2 s) , d_s2 , dZ already written ]
// rk = sum(G o Kzx) over ALL rows (k_rk_from_mm, from the all-reduced partial): added here once, not in the per-shard gradient partial
__global__ void __launch_bounds__(256) k_grad_mm_final(const double* __restrict__ rowacc, int M, int d,
                                                       const double* __restrict__ theta, const double* __restrict__ partial,
                                                       int64_t sP, const double* __restrict__ ds2, double* __restrict__ grad,
                                                       int64_t sG, const double* __restrict__ rk) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* th = theta + (int64_t)b * (d + 2);
  for (int c = 0; c <= d; ++c) {
    double s = 0.0;
    for (int i = tid; i < M; i += 256) s += rowacc[((int64_t)b * M + i) * (d + 1) + c];
    s = block_sum<256>(s, red);
    if (tid == 0) {
      if (c < d) grad[b * sG + c] = s;
      else {
        const double N = partial[b * sP + (int64_t)M * M + M + 2];
        grad[b * sG + d] = (s + (rk ? rk[b] : 0.0)) / th[d] - 0.5 * N / th[d + 1];
      }
    }
  }
  if (tid == 0) grad[b * sG + d + 1] = ds2[b];
}

// acc[b][idx] += sum_t part[b][t][idx]   (fixed order: 8 interleaved tile groups per index, combined 0..7)
// block 256 = 32 consecutive indices x 8 tile groups; grid (ceil(count/32), batch)
__global__ void __launch_bounds__(256) k_reduce_moments(const double* __restrict__ part, int64_t sTile, int64_t sB, int ntiles,
                                                        int64_t count, double* __restrict__ acc) {
  __shared__ double red[8][33];
  const int li = threadIdx.x & 31, grp = threadIdx.x >> 5, b = blockIdx.y;
  const int64_t idx = (int64_t)blockIdx.x * 32 + li;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (idx < count) {
    const double* src = part + b * sB + idx;
    int t = grp;
    for (; t + 24 < ntiles; t += 32) {
      s0 += src[(int64_t)t * sTile];
      s1 += src[(int64_t)(t + 8) * sTile];
      s2 += src[(int64_t)(t + 16) * sTile];
      s3 += src[(int64_t)(t + 24) * sTile];
    }
    for (; t < ntiles; t += 8) s0 += src[(int64_t)t * sTile];
  }
  red[grp][li] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (grp == 0 && idx < count) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][li];
    acc[(int64_t)b * count + idx] += s;
  }
}

// moments -> gradient partial.  mom[b][i][0]=r_i, [1..d]=Q_ic, [d+1..2d]=T_ic are the moments of W against [1, x, x^2].
//   RBF (rk == NULL): W = G o K and dk/d(d2) = -K/2:
//     d_ell_c = sum_i (z^2 r - 2 z Q + T)_ic / ell_c^3 ; d_sf2 = sum_i r_i / sf2 ; dZ_ic = (Q_ic - z_ic r_i)/ell_c^2
//   other stationary kernels (rk != NULL): W = G o dk/d(d2):
//     d_ell_c = -2 sum_i (z^2 r - 2 z Q + T)_ic / ell_c^3 ; dZ_ic = 2 (z_ic r_i - Q_ic)/ell_c^2 ;
//     d_sf2 = rk / sf2 with rk = sum(G o K) = tr(P_A S) + beta^T b / s^2 from the m x m section (k_rk_from_mm)
//   d_s2 = 0 in both cases.
// rk (optional for RBF): d_sf2 = rk / sf2 instead of sum_i r_i / sf2.  The two are the same number, sum(G o K), but rk comes from the
// m x m quantities (S, b are exact sums) while sum_i r_i is an N-long streamed sum that cancels against the Kzz part by cond(Kzz):
// at the headline shape the streamed form left dF/dsf2 1.5e-8 from the long-double reference, the m x m form within 1e-9.
__global__ void __launch_bounds__(256) k_grad_from_moments(const double* __restrict__ mom, int M, int d,
                                                           const double* __restrict__ Z, const double* __restrict__ theta,
                                                           double* __restrict__ grad, int64_t sG, const double* __restrict__ rk,
                                                           int deriv_weighted, int64_t srk = 0) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nq = 2 * d + 1;
  const double* th = theta + (int64_t)b * (d + 2);
  const double* mb = mom + (int64_t)b * M * nq;
  const double fl = deriv_weighted ? -2.0 : 1.0, fz = deriv_weighted ? -2.0 : 1.0;
  for (int c = 0; c < d; ++c) {
    double s = 0.0;
    for (int i = tid; i < M; i += 256) {
      const double z = Z[(int64_t)i * d + c], r = mb[(int64_t)i * nq], Q = mb[(int64_t)i * nq + 1 + c],
                   T = mb[(int64_t)i * nq + 1 + d + c];
      s += fma(z, fma(z, r, -2.0 * Q), T);
      grad[b * sG + d + 2 + (int64_t)i * d + c] = fz * (Q - z * r) / (th[c] * th[c]);
    }
    s = block_sum<256>(s, red);
    if (tid == 0) grad[b * sG + c] = fl * s / (th[c] * th[c] * th[c]);
  }
  double s = 0.0;
  for (int i = tid; i < M; i += 256) s += mb[(int64_t)i * nq];
  s = block_sum<256>(s, red);
  if (tid == 0) {
    // srk == 0 (SGPR): with rk the sum(G o K) term of dF/dsf2 is added ONCE by the m x m section (k_grad_mm_final), not per row shard;
    // srk > 0 (SVGP, no sharding): rk[b * srk] is the k-weighted total accumulated beside the moments and is used here
    grad[b * sG + d] = rk ? (srk > 0 ? rk[b * srk] / th[d] : 0.0) : s / th[d];
    grad[b * sG + d + 1] = 0.0;
  }
}

// rk[b] = sum_{i,n} G_in K_in = tr(P_A S) + beta^T b / s^2   (G = Q A + u y^T, Kzx = L A, Kzx y = L b)
// -- sum(G o Kzx) from the m x m quantities: the k-weighted total of dF/dsf2 (all kernel kinds).
// Two deterministic stages: RK_BLOCKS CTAs per batch element write block partials (contiguous row slices), one CTA adds them in order.
constexpr int RK_BLOCKS = 64;
__global__ void __launch_bounds__(256) k_rk_partial(const double* __restrict__ partial, int64_t sP, int M, int Mp,
                                                    const double* __restrict__ PA, int64_t sMat, double* __restrict__ part) {
  __shared__ double red[8];
  const int b = blockIdx.y, tid = threadIdx.x;
  const double* S = partial + b * sP;
  const int rows = (M + RK_BLOCKS - 1) / RK_BLOCKS, r0 = blockIdx.x * rows, r1 = min(M, r0 + rows);
  double acc = 0.0;
  for (int i = r0; i < r1; ++i)
    for (int j = tid; j < M; j += 256) acc = fma(PA[b * sMat + (int64_t)i * Mp + j], S[(int64_t)i * M + j], acc);
  acc = block_sum<256>(acc, red);
  if (tid == 0) part[(int64_t)b * RK_BLOCKS + blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) k_rk_from_mm(const double* __restrict__ partial, int64_t sP, int M, int Mp,
                                                    const double* __restrict__ theta, int d, const double* __restrict__ part,
                                                    const double* __restrict__ beta, double* __restrict__ rk) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* bv = partial + b * sP + (int64_t)M * M;
  const double s2 = theta[(int64_t)b * (d + 2) + d + 1];
  double acc2 = 0.0;
  for (int i = tid; i < M; i += 256) acc2 = fma(beta[(int64_t)b * Mp + i], bv[i], acc2);
  acc2 = block_sum<256>(acc2, red);
  if (tid == 0) {
    double acc = 0.0;
    for (int k = 0; k < RK_BLOCKS; ++k) acc += part[(int64_t)b * RK_BLOCKS + k];
    rk[b] = acc + acc2 / (s2 * s2);
  }
}

}  // namespace ggp
