// FP64-class "NT" GEMM on the 5th-generation tensor cores by exact integer slicing (Ozaki scheme):
//     C[i,j] = sum_k A[i,k] * B[j,k]
// tcgen05.mma has no f64 kind, and the DMMA path (gemm_tma.cuh) already keeps the FP64 tensor pipe 88-94 % busy; this is the route
// past that roofline that still meets the 1e-8 parity budget.
//   * operands: I8_NS signed 7-bit digits of a row-scaled fixed-point value, x = 2^e * sum_i d_i 2^(-7 (i+1)), |d_i| <= 64, stored as
//     digit PLANES [I8_NS][rows][K] of int8, K contiguous (k_slice_rows / k_build_kc_i8 / the EPI_SLICE epilogue below);
//   * tcgen05.mma kind::i8 (SASS UTCIMMA): 128 x N x 32 products, int32 accumulators in TMEM.  All digit pairs (i, j) of one
//     significance level l = i + j are accumulated EXACTLY in one accumulator (|d|^2 (l+1) K <= 4096 * 8 * K < 2^31 for K <= 2^16),
//     so a 128 x 64 output tile owns I8_NS accumulators = 512 TMEM columns;
//   * digit i of A is multiplied against digits 0..NS-1-i of B in ONE wide MMA: the B digit tiles are consecutive K-major tiles in
//     shared memory (= one tall tile) and their products belong to consecutive levels (= consecutive accumulator columns), which
//     keeps the shared-memory operand reads below the 128 B/clk limit (12 MMAs per 32-byte k-step instead of 36);
//   * TMA (cp.async.bulk.tensor, 64-byte swizzle) stages 8 + 8 digit tiles per 64-byte k-block into a 2-stage mbarrier ring;
//   * persistent CTAs, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) and TMEM owner, warps 2..5 =
//     epilogue (TMEM lane quarters 2,3,0,1).  The epilogue first drains the accumulators to FP64 registers (levels combined from
//     the least significant up), releases TMEM so the next tile's MMAs start, then runs its role-specific part.
// Measured by scripts/probes/ozaki_probe.cu on a B200: |err| / sum|a||b| = 1.7e-16 (an FP64 FMA chain: ~3e-15), 64 TF/s
// FP64-equivalent at K = 1024 and 88 TF/s at K = 16384 (the DMMA peak is 37.2).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ggp {

constexpr int I8_NS = 8;                       // digits per operand (56 bits)
constexpr int I8_BM = 128, I8_BN = 64;         // output tile; I8_NS * I8_BN = 512 TMEM columns
constexpr int I8_BKB = 64;                     // bytes of k per stage row (one 64-byte swizzle row)
constexpr int I8_STAGES = 2;
constexpr int I8_A_BYTES = I8_BM * I8_BKB, I8_B_BYTES = I8_BN * I8_BKB;
constexpr int I8_STAGE_BYTES = I8_NS * (I8_A_BYTES + I8_B_BYTES);
constexpr int I8_THREADS = 192;
constexpr int I8_MAX_D = 32;                   // input dimension limit of the moments epilogue (shared-memory staging of X)
constexpr int I8_EPI_SMEM = I8_BN * (I8_MAX_D + 1) * 8;
constexpr int I8_SMEM = I8_STAGES * I8_STAGE_BYTES + 1024 + 256 + I8_EPI_SMEM;
constexpr int I8_TMEM_COLS = 512;
constexpr int I8_MAX_K = 65536;                // exact int32 accumulation bound
static_assert(I8_NS * I8_BN <= 512, "level accumulators must fit TMEM");

enum { I8_EPI_F64 = 0, I8_EPI_SLICE = 1, I8_EPI_MOMENTS = 2 };

struct I8P {
  int tiles_m, tiles_n, splits, total;   // work list: tile (tm, tn) x split
  int K;                                 // k extent in bytes (multiple of I8_BKB after padding; TMA zero-fills beyond the tensor)
  int lower_a;                           // A lower triangular: k range of row tile tm clipped to (tm + 1) * 128
  int sym;                               // only tiles that touch the upper triangle (tn * 64 + 63 >= tm * 128)
  int n_major;                           // enumerate the row tiles of one column tile consecutively
  int M, N;                              // valid rows of A / rows of B (stores are clipped to them)
  const int* ea; int ea0;                // exponents of the A rows (array or scalar)
  const int* eb; int eb0;                // exponents of the B rows
  double alpha, beta;
  // I8_EPI_F64:  C[(split)] = alpha * acc + beta * C
  double* C; int64_t ldc, sSplit;
  // I8_EPI_SLICE: digits of alpha * acc with the fixed exponent eo -> Oq[plane][row][col]; optional fused row dots against yv
  int8_t* Oq; int64_t o_ld, o_plane; int eo;
  const double* yv; double* rowdot;      // rowdot[tn][M]
  // I8_EPI_MOMENTS: W = (alpha * acc + u[row] * yv[col]) * Kmul[col][row];  mom[tn][row][:] = sum_col W * [1, x_col, x_col^2]
  const double* u; const double* Kmul; int64_t ldk; const double* Xc; int d; double* mom; int64_t sMomTile;
};

__device__ __forceinline__ void i8_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "I8_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra I8_DONE;\n"
      "bra I8_WAIT;\n"
      "I8_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void i8_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void i8_tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// K-major operand tile, 64-byte swizzle: 64-byte rows, 8-row groups 512 bytes apart (SBO), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t i8_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void i8_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void i8_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void i8_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

struct I8Item {
  int tm, tn, split, kb_lo, kb_hi;
};
__device__ __forceinline__ void i8_decode(const I8P& p, int w, I8Item& o) {
  o.split = w % p.splits;
  int t = w / p.splits;
  if (p.sym) {   // row tile tm owns the column tiles tn >= 2 tm
    int tm = 0;
    while (t >= p.tiles_n - 2 * tm) { t -= p.tiles_n - 2 * tm; ++tm; }
    o.tm = tm; o.tn = 2 * tm + t;
  } else if (p.n_major) {
    o.tn = t / p.tiles_m; o.tm = t - o.tn * p.tiles_m;
  } else {
    const int r = t / p.tiles_n;
    o.tn = t - r * p.tiles_n;
    o.tm = p.lower_a ? (p.tiles_m - 1 - r) : r;   // heavy (long-k) row tiles first
  }
  int k_hi = p.K;
  if (p.lower_a) k_hi = min(k_hi, (o.tm + 1) * I8_BM);
  const int nkb = (k_hi + I8_BKB - 1) / I8_BKB;
  const int per = (nkb + p.splits - 1) / p.splits;
  o.kb_lo = o.split * per;
  o.kb_hi = min(nkb, o.kb_lo + per);
}

// signed 7-bit digits of v (|v| < 1/2), most significant first
__device__ __forceinline__ void i8_digits(double v, int8_t (&dg)[I8_NS]) {
#pragma unroll
  for (int i = 0; i < I8_NS; ++i) {
    v *= 128.0;
    const double r = rint(v);
    v -= r;
    dg[i] = (int8_t)(int)r;
  }
}

template <int EPI>
__global__ void __launch_bounds__(I8_THREADS, 1)
k_gemm_i8(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const I8P p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + I8_STAGES * I8_STAGE_BYTES);
  uint64_t* empty = full + I8_STAGES;
  uint64_t* tmem_full = empty + I8_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  double* xs = reinterpret_cast<double*>(base + I8_STAGES * I8_STAGE_BYTES + 256);   // [I8_BN][d + 1]: x rows and y of the tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < I8_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(I8_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int n = 0;
      for (int w = blockIdx.x; w < p.total; w += G) {
        I8Item it;
        i8_decode(p, w, it);
        for (int kb = it.kb_lo; kb < it.kb_hi; ++kb, ++n) {
          const int s = n % I8_STAGES;
          if (n >= I8_STAGES) i8_mbar_wait(&empty[s], ((n / I8_STAGES) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], I8_STAGE_BYTES);
          unsigned char* st = base + s * I8_STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < I8_NS; ++i) i8_tma_load_3d(st + i * I8_A_BYTES, &tmA, kb * I8_BKB, it.tm * I8_BM, i, &full[s]);
#pragma unroll
          for (int j = 0; j < I8_NS; ++j)
            i8_tma_load_3d(st + I8_NS * I8_A_BYTES + j * I8_B_BYTES, &tmB, kb * I8_BKB, it.tn * I8_BN, j, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int n = 0, item = 0;
    for (int w = blockIdx.x; w < p.total; w += G) {
      I8Item it;
      i8_decode(p, w, it);
      if (it.kb_hi <= it.kb_lo) continue;
      if (item > 0) i8_mbar_wait(tmem_empty, (item - 1) & 1);   // the epilogue has drained the previous tile's accumulators
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      for (int kb = it.kb_lo; kb < it.kb_hi; ++kb, ++n) {
        const int s = n % I8_STAGES;
        i8_mbar_wait(&full[s], (n / I8_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        if (lane == 0) {
          const uint32_t sa = smem_u32(base + s * I8_STAGE_BYTES), sb = sa + I8_NS * I8_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < I8_BKB / 32; ++kk) {
#pragma unroll
            for (int i = 0; i < I8_NS; ++i) {
              const uint64_t ad = i8_desc_sw64(sa + i * I8_A_BYTES + kk * 32);
              const int ncols = I8_BN * (I8_NS - i);
#pragma unroll
              for (int off = 0; off < ncols; off += 256) {
                const int nn = (ncols - off) < 256 ? (ncols - off) : 256;
                // D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
                const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
                const uint64_t bd = i8_desc_sw64(sb + (off / I8_BN) * I8_B_BYTES + kk * 32);
                i8_umma(tmem_base + (uint32_t)(i * I8_BN + off), ad, bd, idesc, (kb > it.kb_lo || kk > 0 || i > 0) ? 1u : 0u);
              }
            }
          }
          i8_umma_commit(&empty[s]);                            // frees the stage once the MMAs that read it are done
          if (kb == it.kb_hi - 1) i8_umma_commit(tmem_full);    // accumulators of this tile complete
        }
        __syncwarp();
      }
      ++item;
    }
  } else {
    // ================= epilogue (4 warps; warp w owns TMEM lanes [32 (w % 4), +32)) =================
    const int quarter = warp & 3;
    const int et = threadIdx.x - 64;   // 0..127
    int item = 0;
    for (int w = blockIdx.x; w < p.total; w += G) {
      I8Item it;
      i8_decode(p, w, it);
      if (it.kb_hi <= it.kb_lo) continue;
      const int row = it.tm * I8_BM + quarter * 32 + lane;
      const int col0 = it.tn * I8_BN;
      if (EPI == I8_EPI_MOMENTS) {
        // stage the tile's x rows and y into shared memory (previous tile's readers are past their last read: barrier below)
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        const int d = p.d;
        for (int i = et; i < I8_BN * (d + 1); i += 128) {
          const int c = i / (d + 1), q = i - c * (d + 1);
          const int gc = col0 + c;
          double v = 0.0;
          if (gc < p.N) v = (q < d) ? p.Xc[(int64_t)gc * d + q] : p.yv[gc];
          xs[i] = v;
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
      }
      i8_mbar_wait(tmem_full, item & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      double acc[I8_BN];
#pragma unroll
      for (int c = 0; c < I8_BN; ++c) acc[c] = 0.0;
#pragma unroll
      for (int c0 = 0; c0 < I8_BN; c0 += 16) {
#pragma unroll
        for (int l = I8_NS - 1; l >= 0; --l) {
          uint32_t v[16];
          i8_tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(l * I8_BN + c0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          const double wl = exp2(-7.0 * (l + 2));
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c0 + c] = fma((double)(int)v[c], wl, acc[c0 + c]);
        }
      }
      // accumulators are in registers: hand TMEM back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
      __syncwarp();
      if (lane == 0) i8_mbar_arrive(tmem_empty);
      ++item;

      const int e_r = p.ea ? p.ea[min(row, p.M - 1)] : p.ea0;
      const bool rok = row < p.M;
      if (p.eb) {
#pragma unroll
        for (int c = 0; c < I8_BN; ++c) acc[c] = p.alpha * ldexp(acc[c], e_r + p.eb[min(col0 + c, p.N - 1)]);
      } else {
        const double sc = p.alpha * exp2((double)(e_r + p.eb0));
#pragma unroll
        for (int c = 0; c < I8_BN; ++c) acc[c] *= sc;
      }

      if (EPI == I8_EPI_F64) {
        if (rok) {
          double* dst = p.C + (int64_t)it.split * p.sSplit + (int64_t)row * p.ldc + col0;
          if (col0 + I8_BN <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int c = 0; c < I8_BN; c += 2) {
              double2* d2 = reinterpret_cast<double2*>(dst + c);
              double v0 = acc[c], v1 = acc[c + 1];
              if (p.beta != 0.0) { const double2 o = *d2; v0 += p.beta * o.x; v1 += p.beta * o.y; }
              *d2 = make_double2(v0, v1);
            }
          } else {
#pragma unroll
            for (int c = 0; c < I8_BN; ++c)
              if (col0 + c < p.N) dst[c] = acc[c] + (p.beta != 0.0 ? p.beta * dst[c] : 0.0);
          }
        }
      } else if (EPI == I8_EPI_SLICE) {
        if (p.rowdot) {
          double sdot = 0.0;
#pragma unroll
          for (int c = 0; c < I8_BN; ++c) sdot = fma(acc[c], (col0 + c < p.N) ? p.yv[col0 + c] : 0.0, sdot);
          if (rok) p.rowdot[(int64_t)it.tn * p.M + row] = sdot;
        }
        // digit planes of the tile row: 64 consecutive bytes per plane (columns beyond N are zero because their B rows are zero)
        const double si = exp2((double)-p.eo);
#pragma unroll
        for (int c0 = 0; c0 < I8_BN; c0 += 16) {
          uint32_t pk[I8_NS][4];
#pragma unroll
          for (int i = 0; i < I8_NS; ++i) pk[i][0] = pk[i][1] = pk[i][2] = pk[i][3] = 0u;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            int8_t dg[I8_NS];
            i8_digits(acc[c0 + c] * si, dg);
#pragma unroll
            for (int i = 0; i < I8_NS; ++i) pk[i][c >> 2] |= ((uint32_t)(uint8_t)dg[i]) << (8 * (c & 3));
          }
          if (rok) {
#pragma unroll
            for (int i = 0; i < I8_NS; ++i)
              *reinterpret_cast<uint4*>(p.Oq + (int64_t)i * p.o_plane + (int64_t)row * p.o_ld + col0 + c0) =
                  make_uint4(pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
          }
        }
      } else {
        // W = (G + u y^T) o Kmul, then the moments against [1, x, x^2] of the tile's 64 columns
        const int d = p.d, nq = 2 * d + 1;
        const double ui = rok ? p.u[row] : 0.0;
        double r0 = 0.0;
#pragma unroll
        for (int c = 0; c < I8_BN; ++c) {
          const int gc = col0 + c;
          const double kv = (rok && gc < p.N) ? p.Kmul[(int64_t)gc * p.ldk + row] : 0.0;
          acc[c] = fma(ui, xs[c * (d + 1) + d], acc[c]) * kv;
          r0 += acc[c];
        }
        double* mo = p.mom + (int64_t)it.tn * p.sMomTile + (int64_t)row * nq;
        if (rok) mo[0] = r0;
        for (int q = 0; q < d; ++q) {
          double m1 = 0.0, m2 = 0.0;
#pragma unroll
          for (int c = 0; c < I8_BN; ++c) {
            const double x = xs[c * (d + 1) + q];
            const double wx = acc[c] * x;
            m1 += wx;
            m2 = fma(wx, x, m2);
          }
          if (rok) {
            mo[1 + q] = m1;
            mo[1 + d + q] = m2;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(I8_TMEM_COLS));
}

// row-scaled signed-digit slicing of X[R x K] (leading dimension ld): planes Xq[i][row][k] (leading dimension ldq, plane stride
// plane), per-row exponent ex[row] with |x| / 2^e < 1/2.  One warp per row.  Columns [K, Kpad) are written as zero.
__global__ void __launch_bounds__(256) k_slice_rows(const double* __restrict__ X, int R, int K, int64_t ld, int8_t* __restrict__ Xq,
                                                    int64_t ldq, int64_t plane, int Kpad, int* __restrict__ ex) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  double mx = 0.0;
  for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(X[(int64_t)row * ld + k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int e = mx > 0.0 ? ilogb(mx) + 2 : 0;
  if (lane == 0) ex[row] = e;
  for (int k = lane; k < Kpad; k += 32) {
    int8_t dg[I8_NS];
    i8_digits(k < K ? ldexp(X[(int64_t)row * ld + k], -e) : 0.0, dg);
#pragma unroll
    for (int i = 0; i < I8_NS; ++i) Xq[(int64_t)i * plane + (int64_t)row * ldq + k] = dg[i];
  }
}

// digit planes of an existing k(X,Z) tile Kc[n][m] (FP64, leading dimension ldk) with the fixed exponent e (k <= sf2 < 2^(e-1)):
// Kq[i][n][m].  Each thread converts 16 consecutive m of one row: 16-byte stores per plane.
__global__ void __launch_bounds__(256) k_slice_fixed(const double* __restrict__ Kc, int64_t rows, int cols, int64_t ldk,
                                                     int8_t* __restrict__ Kq, int64_t ldq, int64_t plane, const double* __restrict__ theta,
                                                     int d) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int cpr = cols / 16;
  const int64_t row = t / cpr;
  const int c0 = (int)(t - row * cpr) * 16;
  if (row >= rows) return;
  const int e = ilogb(theta[d]) + 2;
  const double si = exp2((double)-e);
  uint32_t pk[I8_NS][4];
#pragma unroll
  for (int i = 0; i < I8_NS; ++i) pk[i][0] = pk[i][1] = pk[i][2] = pk[i][3] = 0u;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    int8_t dg[I8_NS];
    i8_digits(Kc[row * ldk + c0 + c] * si, dg);
#pragma unroll
    for (int i = 0; i < I8_NS; ++i) pk[i][c >> 2] |= ((uint32_t)(uint8_t)dg[i]) << (8 * (c & 3));
  }
#pragma unroll
  for (int i = 0; i < I8_NS; ++i)
    *reinterpret_cast<uint4*>(Kq + (int64_t)i * plane + row * ldq + c0) = make_uint4(pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
}

}  // namespace ggp
