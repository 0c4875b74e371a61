// TMA-fed FP64 tensor-core (DMMA.8x8x4) "NT" GEMM:  C[i,j] (op)= alpha * sum_k A[i,k] * B[j,k]     (same contract as k_gemm_nt)
//
// Why a second mainloop: micro-probes on the B200 (scripts/probes/dmma_probe.cu) show that 8 warps feeding DMMA from shared memory
// reach 99.7 % of the tensor-pipe peak when the operand tiles arrive by bulk async copies signalled on mbarriers, but only ~95 %
// with per-thread LDGSTS + a CTA barrier per k-step (and ~91-93 % in the full LDGSTS kernel with its address arithmetic).  So:
//   * warps 8-11 form the PRODUCER warpgroup (setmaxnreg hands its registers to the consumers): one elected lane walks the same static work list as the consumers and issues two
//     cp.async.bulk.tensor (UTMALDG) loads per k-step (A box 128 rows x 16 k, B box 128 rows x 16 k, 128-byte swizzle) into a
//     6-stage ring; out-of-range rows / k are zero-filled by the TMA unit, so there is no edge-case code in the loop;
//   * warps 0-7 are CONSUMERS (2 x 4, warp tile 64 x 32, accumulators in registers): they wait on the stage's "full" mbarrier,
//     issue 128 DMMAs, and release the stage with one arrive per warp on its "empty" mbarrier.  There is no CTA-wide barrier in
//     the mainloop, and the producer runs ahead across tile boundaries, so epilogues never wait for a pipeline refill.
// Bank conflicts: with the 128-byte swizzle the 16-byte chunk c of row r lives at chunk c ^ (r & 7).  A DMMA k-step may use ANY
// four k indices as long as A and B agree, so lane q of k-step kk reads k = 2*kk + (q & 1) + 8*(q >> 1): the half-warp's 16
// (row, k) pairs then fall into 16 distinct 8-byte bank pairs (conflict-free LDS.64), see DESIGN.md.
#pragma once
#include <cuda.h>
#include "gemm_dmma.cuh"

namespace ggp {

constexpr int TK = 16;                         // k-depth of one stage (one 128-byte swizzled row per tile row)
constexpr int TSTAGES = 6;
constexpr int T_CONSUMER_WARPS = 8;
constexpr int T_THREADS = (T_CONSUMER_WARPS + 4) * 32;   // 2 consumer warpgroups + 1 producer warpgroup (register re-split below)
constexpr int T_TILE_BYTES = 128 * TK * 8;     // 16 KB per operand tile
constexpr int T_STAGE_BYTES = 2 * T_TILE_BYTES;
constexpr int T_SMEM = TSTAGES * T_STAGE_BYTES + 4 * BM * 8 + 2 * TSTAGES * 8 + 1024;   // + row-dot exchange + mbarriers + alignment slack
static_assert(BM == 128 && BN == 128, "the TMA mainloop is written for 128 x 128 tiles");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 4-D tiled TMA load (coordinates: k, row, inner batch, outer batch), completion on an mbarrier
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(T_THREADS, 1)
k_gemm_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmP p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  double* sR = reinterpret_cast<double*>(base + TSTAGES * T_STAGE_BYTES);            // [4][BM]
  uint64_t* full = reinterpret_cast<uint64_t*>(base + TSTAGES * T_STAGE_BYTES + 4 * BM * 8);
  uint64_t* empty = full + TSTAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, nrounds = (p.total + G - 1) / G;
  auto item_of = [&](int r) { return r * G + ((r & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x); };

  if (tid == 0) {
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], T_CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp >= T_CONSUMER_WARPS) {
    // ======================= producer warpgroup =======================
    // the kernel is compiled for 384 threads x 168 registers; the producer warpgroup hands most of its share to the consumers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n" ::);
    if (warp == T_CONSUMER_WARPS && lane == 0) {
      int n = 0;
      for (int r = 0; r < nrounds; ++r) {
        const int w = item_of(r);
        if (w >= p.total) continue;
        WorkItem wi;
        decode_work<TK>(p, w, wi);
        const int k_base = wi.k_lo + wi.it_lo * TK;
        for (int it = 0; it < wi.niter; ++it, ++n) {
          const int stage = n % TSTAGES;
          if (n >= TSTAGES) mbar_wait_parity(&empty[stage], ((n / TSTAGES) - 1) & 1);
          mbar_arrive_expect_tx(&full[stage], T_STAGE_BYTES);
          unsigned char* dst = base + stage * T_STAGE_BYTES;
          tma_load_4d(dst, &tmA, k_base + it * TK, wi.tm * BM, wi.pz * p.tmA_pz, wi.bz * p.tmA_bz, &full[stage]);
          tma_load_4d(dst + T_TILE_BYTES, &tmB, k_base + it * TK, wi.tn * BN, wi.pz * p.tmB_pz, wi.bz * p.tmB_bz, &full[stage]);
        }
      }
    }
    return;
  }

  // ======================= consumers =======================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n" ::);
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, q = lane & 3;
  // byte offset of this lane's element inside a tile for k-step kk: row g of the fragment, k = 2*kk + (q&1) + 8*(q>>1)
  int koff[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) koff[kk] = g * 128 + (((kk + 4 * (q >> 1)) ^ g) << 4) + ((q & 1) << 3);
  const int a_row_off = wm * 64 * 128, b_row_off = T_TILE_BYTES + wn * 32 * 128;
  const bool diag_lower = (p.kmode & KM_A_LOWER) != 0;

  int n = 0;
  for (int r = 0; r < nrounds; ++r) {
    const int w = item_of(r);
    if (w >= p.total) continue;
    WorkItem wi;
    decode_work<TK>(p, w, wi);
    if (wi.niter == 0) continue;
    const int row0_m = wi.tm * BM;
    const int k_base = wi.k_lo + wi.it_lo * TK;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int it = 0; it < wi.niter; ++it, ++n) {
      const int stage = n % TSTAGES;
      mbar_wait_parity(&full[stage], (n / TSTAGES) & 1);
      const unsigned char* sa = base + stage * T_STAGE_BYTES + a_row_off;
      const unsigned char* sb = base + stage * T_STAGE_BYTES + b_row_off;
      const int kstep0 = k_base + it * TK - row0_m - wm * 64;   // k of this stage relative to the warp's first row
      const bool stair = diag_lower && kstep0 + TK > 0;          // some k of the stage lies right of the warp's first row
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const double*>(sa + i * 1024 + koff[kk]);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double*>(sb + j * 1024 + koff[kk]);
        if (stair) {
          // lower-triangular A: fragment rows [8i, 8i+8) need this k-step only if its smallest k (kstep0 + 2kk) is <= 8i+7.
          // A predicated-off DMMA still occupies the tensor pipe, so this is real control flow (jump into the row sequence).
          const int kmin = kstep0 + 2 * kk;
          const int i_lo = kmin < 0 ? 0 : (kmin >> 3);
#define GGP_FRAG_ROW(i) \
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
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    gemm_epilogue<EPI, true>(p, wi, acc, sR, tid, wm, wn, g, q);
  }
}

}  // namespace ggp
