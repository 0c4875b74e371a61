// FP64-class "NT" GEMM on the 5th-generation tensor cores by exact integer slicing (Ozaki scheme):
//     C[i,j] = sum_k A[i,k] * B[j,k]
// tcgen05.mma has no f64 kind, and the DMMA path (gemm_tma.cuh) already keeps the FP64 tensor pipe 88-94 % busy; this is the route
// past that roofline that still meets the 1e-8 parity budget.
//   * operands: I8_NS = 7 balanced radix-256 digits (d_i in [-128, 127]) of a row-scaled fixed-point value,
//     x = 2^e * sum_i d_i 2^(-8 (i+1)), |x| / 2^e < 0.498 (i8_exp_for), stored as
//     digit PLANES [I8_NS][rows][K] of int8, K contiguous (k_slice_rows / k_build_kc_i8 / the EPI_SLICE epilogue below).  The digit
//     bytes are the bytes of the 56-bit integer rn(x 2^(56 - e)) + C with the top bit of the six low bytes flipped (i8_digits_fixed);
//   * tcgen05.mma kind::i8 (SASS UTCIMMA): 128 x N x 32 products, int32 accumulators in TMEM.  All digit pairs (i, j) of one
//     significance level l = i + j <= 6 are accumulated EXACTLY in one accumulator (|d_i d_j| (l+1) K <= 2^14 * 7 * K < 2^31 for
//     K <= 16384), so a 128 x 64 output tile owns I8_NS accumulators = 448 TMEM columns; 28 digit products per FP64 product
//     (8 x 7-bit digits needed 36 for the same 56 bits);
//   * digit i of A is multiplied against digits 0..NS-1-i of B in ONE or TWO wide MMAs: the B digit tiles are consecutive K-major
//     tiles in shared memory (= one tall tile) and their products belong to consecutive levels (= consecutive accumulator columns):
//     10 MMAs per 32-byte k-step instead of 28, which keeps the shared-memory operand reads below the 128 B/clk limit;
//   * TMA (cp.async.bulk.tensor, 64-byte swizzle) stages 7 + 7 digit tiles per 64-byte k-block into a 2-stage mbarrier ring;
//   * persistent CTAs, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) and TMEM owner, warps 2..9 =
//     epilogue (TMEM lane quarter w % 4, column half (w - 2) / 4).  The epilogue first drains the accumulators to registers (levels
//     combined in exact 64-bit integers), releases TMEM so the next tile's MMAs start, then runs its role-specific part.
// Measured by scripts/probes/ozaki_probe.cu on a B200 (8 x 7-bit variant): |err| / sum|a||b| = 1.7e-16 (an FP64 FMA chain: ~3e-15).
#pragma once
#include <cuda.h>
#include <type_traits>
#include "common.cuh"

namespace ggp {

constexpr int I8_NS = 7;                       // radix-256 digits per operand (56 bits)
constexpr int I8_BM = 128, I8_BN = 64;         // output tile; I8_NS * I8_BN = 448 TMEM columns
#ifndef GGP_I8_BKB
#define GGP_I8_BKB 64
#endif
#ifndef GGP_I8_STAGES
#define GGP_I8_STAGES 2
#endif
constexpr int I8_BKB = GGP_I8_BKB;             // bytes of k per stage row = one swizzle row (64: SWIZZLE_64B; 32: SWIZZLE_32B, half-size
                                               // stages and a ring twice as deep for the same shared memory)
constexpr int I8_STAGES = GGP_I8_STAGES;
static_assert(I8_BKB == 64 || I8_BKB == 32, "k-block = one 64- or 32-byte swizzle row");
constexpr int I8_A_BYTES = I8_BM * I8_BKB, I8_B_BYTES = I8_BN * I8_BKB;
constexpr int I8_STAGE_BYTES = I8_NS * (I8_A_BYTES + I8_B_BYTES);
constexpr int I8_EPI_WARPS = 8;                // two per TMEM lane quarter
constexpr int I8_EC = 32;                      // tile columns per epilogue warp (both tile widths)
constexpr int I8_THREADS = 64 + 32 * I8_EPI_WARPS;
constexpr int I8_MAX_D = 16;                   // input dimension limit of the moments epilogue (shared-memory staging of X)
constexpr int I8_EPI_SMEM = I8_BN * (I8_MAX_D + 1) * 8;
constexpr int I8_WSTAGE_LD = 12;               // doubles per staged W row (8 used): 24-word stride = conflict-free DMMA fragment loads
constexpr int I8_WSTAGE_BYTES = I8_EPI_WARPS * 32 * I8_WSTAGE_LD * 8;

// Two tile widths (template parameter BN of k_gemm_i8):
//   BN = 64: one 128 x 64 tile owns all 7 x 64 = 448 TMEM columns (single-buffered); the 8 epilogue warps split its columns in two
//            halves.  Right for long k (the SYRK: 256 k-blocks per chunk, the hand-over is amortised) -- its MMAs are wider.
//   BN = 32: 128 x 32 tiles, TWO accumulator sets of 7 x 32 = 224 columns: the MMAs of tile t + 1 run while tile t is drained and
//            its epilogue runs.  The 8 epilogue warps form two teams of 4 (one warp per TMEM lane quarter); team = tile parity, so
//            each team has two tile periods for its drain + epilogue.  Right for the k = m products (triangular multiply, backward
//            GEMM: 2 .. 16 k-blocks per tile), where the single-buffered hand-over left the tensor pipe idle 25-30 % of the time
//            (measured: 1318 ns per k-block with the hand-over bubble alone, 1747 with the FP64 epilogue, against 1197 for the MMAs).
template <int BN> struct I8Tile {
  static_assert(BN == 64 || BN == 32, "tile width");
  static constexpr int NBUF = BN == 32 ? 2 : 1;
  static constexpr int B_BYTES = BN * I8_BKB;
  static constexpr int STAGE_BYTES = I8_NS * (I8_A_BYTES + B_BYTES);
  static constexpr int BUF_COLS = I8_NS * BN;          // TMEM columns of one accumulator set
  static constexpr int TEAM_WARPS = I8_EPI_WARPS / NBUF;   // epilogue warps per tile
};
template <int EPI, int BN> __host__ __device__ constexpr int i8_stages() { return BN == 64 ? I8_STAGES : (EPI == 2 /* moments: staging buffers */ ? 2 : 3); }
template <int EPI, int BN> __host__ __device__ constexpr int i8_smem() {
  return i8_stages<EPI, BN>() * I8Tile<BN>::STAGE_BYTES + 1024 + 256 + I8_EPI_SMEM + I8_WSTAGE_BYTES * (EPI == 2 || BN == 64 ? 1 : 0);
}
constexpr int I8_SMEM = I8_STAGES * I8_STAGE_BYTES + 1024 + 256 + I8_EPI_SMEM + I8_WSTAGE_BYTES;
static_assert(I8_SMEM <= 232448, "shared memory budget (227 KB)");
constexpr int I8_TMEM_COLS = 512;
constexpr int I8_MAX_K = 16384;                // exact int32 accumulation bound: 7 pairs x 2^14 x K < 2^31 (see i8_digits)
constexpr int I8_K_GROUP4 = 4096;              // k extent up to which four level sums combine exactly in one 64-bit integer < 2^53
static_assert(I8_NS * I8_BN <= 512, "level accumulators must fit TMEM");

enum { I8_EPI_F64 = 0, I8_EPI_SLICE = 1, I8_EPI_MOMENTS = 2 };

struct I8P {
  int tiles_m, tiles_n, splits, total;   // work list: tile (tm, tn) x split
  int K;                                 // k extent in bytes (multiple of I8_BKB after padding; TMA zero-fills beyond the tensor)
  int lower_a;                           // A lower triangular: k range of row tile tm clipped to (tm + 1) * 128
  int sym;                               // only tiles that touch the upper triangle (tn * 64 + 63 >= tm * 128)
  int n_major;                           // enumerate the row tiles of one column tile consecutively
  int snake;                             // CTA b takes item b of even rounds and item G-1-b of odd rounds (work list sorted by weight:
                                         // the triangular multiply's k range grows with tm, plain round-robin is 6 % off balance)
  int M, N;                              // valid rows of A / rows of B (stores are clipped to them)
  const int* ea; int ea0;                // exponents of the A rows (array or scalar)
  const int* eb; int eb0;                // exponents of the B rows
  double alpha, beta;
  // I8_EPI_F64:  C[(split)] = alpha * acc + beta * C
  double* C; int64_t ldc, sSplit;
  // I8_EPI_SLICE: digits of alpha * acc with the fixed exponent eo -> Oq[plane][row][col]; optional fused row dots against yv
  int8_t* Oq; int64_t o_ld, o_plane; int eo;
  int o_chunk; int64_t o_chunk_stride;   // o_chunk > 0: column block c = col / o_chunk lives at Oq + c * o_chunk_stride, columns relative to it
                                         // (one [plane][row][o_chunk] array per chunk: the SYRK reads a chunk with a 16 KB row pitch, not n_local)
  const double* yv; double* rowdot;      // rowdot[2 tn + half][M] (+= when rowdot_acc: one buffer over the chunks of a pass)
  int rowdot_acc;
  int rowdot_reg;                        // row dots stay in registers over all tiles of the CTA (n-major snake order with 2 G % tiles_m == 0: a
                                         // CTA only ever sees two row tiles, one per round parity) and are written once: rowdot[2 b + half][M]
  // I8_EPI_MOMENTS: W = (alpha * acc + u[row] * yv[col]) * Kmul[col][row];  mom[tn][row][:] = sum_col W * [1, x_col, x_col^2]
  const double* u; const double* Kmul; int64_t ldk; const double* Xc; int d; double* mom; int64_t sMomTile;
  int nchunk_groups_max;                 // host: upper bound on the chunk groups (= partial buffers available)
  int nchunk, ntile, k_last;             // I8_EPI_F64 over nchunk k-chunks of extent K (the last one k_last) whose digit planes are stacked along the
                                         // plane axis (plane = chunk * NS + digit): CTA b owns tile b % ntile and the chunks b / ntile + groups r,
                                         // drains each chunk's exact int32 sums into FP64 registers and stores the total once
                                         // (C + (b / ntile) * sSplit): the SYRK of a whole pass is one launch, no read-modify-write
  int mom_accum;                         // I8_EPI_MOMENTS, 2 d + 1 <= 24: grid = tiles_m x ng, CTA b owns row tile b % tiles_m and the column tiles
                                         // b / tiles_m + ng k; its moments stay in registers over all its tiles and are written once:
                                         // mom[(b / tiles_m) * 2 + half][row][:]  (2 ng slabs instead of one per 32 columns)
  int serial_epi;                        // 1: hand TMEM back only after the whole epilogue (FP64 epilogue math and the running
                                         // UTCIMMA stream throttle each other on the tensor / FP64 pipe: measured 10x slower when overlapped)
  int b_mn;                              // B operand is MN-major: planes [plane][k rows][n contiguous] (the digit planes of A^T as the triangular
                                         // multiply wrote them, k = inducing index): TMA box = 64 k-rows x 64 n-bytes, UMMA descriptor MN-major SW64
                                         // with the digit tiles 4096 bytes apart (LBO); wide MMAs are split at multiples of 64 columns
  int b_planes;                          // b_mn, host: number of planes the B tensor map spans (NS x column blocks)
  int b_chunk;                           // b_mn: columns per chunk array (column block c = col / b_chunk lives at planes c * NS + digit), 0 = one array
  const int* e_dev;                      // device-side scalar exponents (no host read-back of theta): e_dev[sel - 1] replaces ea0 / eb0 / eo
  int ea0_sel, eb0_sel, eo_sel;          // 0 = use the immediate value
  int exp_skip_b;                        // developer experiment: do not load the B operand (wrong results; measures the L2 -> SM bound)
  int mom_frag_off;                      // developer A/B (GGP_I8_MOM_FRAG=0): moments through the shared-memory W patch instead of the fragment-layout drain
  int exp_skip_a, exp_no_epi;            // developer experiments: no A loads / epilogue warps hand TMEM straight back (mainloop only)
  long long* dbg;                        // developer timeline (CTA 0): [role][item][4] clock64 stamps, or NULL
};
constexpr int I8_DBG_ITEMS = 16;
#define I8_STAMP(role, item, slot)                                                                              \
  do {                                                                                                          \
    if (p.dbg && blockIdx.x == 0 && (item) < I8_DBG_ITEMS) p.dbg[((role) * I8_DBG_ITEMS + (item)) * 4 + (slot)] = clock64(); \
  } while (0)

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
// K-major operand tile, rows of I8_BKB bytes = one swizzle row (64-byte: layout type 4, 32-byte: type 6), 8-row groups 8 * I8_BKB bytes
// apart (SBO), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t i8_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((8 * I8_BKB) >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(I8_BKB == 64 ? 4 : 6) << 61);
}
// MN-major operand tile, 64-byte swizzle: 64 k-rows of 64 contiguous MN bytes (8-row groups 512 bytes apart: SBO), further 64-byte
// MN blocks `lbo` bytes apart (LBO) -- canonical layout ((4,n),(8,k)):((1,LBO),(4,SBO)) in 16-byte units
__device__ __forceinline__ uint64_t i8_desc_mn_sw64(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
// the same for 32-byte MN blocks (BN = 32 tiles): 32-byte swizzle, k-rows 32 bytes apart, 8-row groups 256 bytes apart
// -- canonical ((2,n),(8,k)):((1,LBO),(2,SBO))
__device__ __forceinline__ uint64_t i8_desc_mn_sw32(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)6 << 61);
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

// 16 TMEM lanes x 32 columns in the mma fragment layout: thread 4 g + q gets, for column group j = 0..3 (8 columns each),
// v[4 j + 0..1] = lane g, columns 8 j + 2 q, + 1 and v[4 j + 2..3] = lane g + 8, same columns  (PTX tcgen05.ld shape .16x256b)
__device__ __forceinline__ void i8_tmem_ld_frag(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

struct I8Item {
  int tm, tn, split, kb_lo, kb_hi, chunk;
};
__device__ __forceinline__ void i8_decode(const I8P& p, int w, I8Item& o) {
  o.chunk = 0;
  if (p.nchunk) {   // w = chunk * ntile + tile
    o.chunk = w / p.ntile;
    w -= o.chunk * p.ntile;
  }
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
  int k_hi = (p.nchunk && o.chunk == p.nchunk - 1) ? p.k_last : p.K;
  if (p.lower_a) k_hi = min(k_hi, (o.tm + 1) * I8_BM);
  const int nkb = (k_hi + I8_BKB - 1) / I8_BKB;
  const int per = (nkb + p.splits - 1) / p.splits;
  o.kb_lo = o.split * per;
  o.kb_hi = min(nkb, o.kb_lo + per);
}

// work item of CTA b in round k (all three warp roles enumerate the same list)
__device__ __forceinline__ int i8_item(const I8P& p, int k, int b, int G) {
  if (p.nchunk) {
    const int c = b / p.ntile + (G / p.ntile) * k;
    return c < p.nchunk ? c * p.ntile + b % p.ntile : p.total;
  }
  if (p.mom_accum) {   // n-major encoding of (tm = b % tiles_m, tn = b / tiles_m + ng k); beyond the last column tile -> >= total
    const int ng = G / p.tiles_m, tn = b / p.tiles_m + ng * k;
    return tn < p.tiles_n ? tn * p.tiles_m + b % p.tiles_m : p.total;
  }
  return k * G + ((p.snake && (k & 1)) ? (G - 1 - b) : b);
}
__device__ __forceinline__ int i8_rounds(const I8P& p, int G) {
  if (p.nchunk) { const int ng = G / p.ntile; return (p.nchunk + ng - 1) / ng; }
  if (p.mom_accum) { const int ng = G / p.tiles_m; return (p.tiles_n + ng - 1) / ng; }
  return (p.total + G - 1) / G;
}

// Exponent of a row (or of a bounded operand) with largest magnitude mx: the smallest e with mx / 2^e < 0.498, so that the top digit
// d_0 = floor(x 2^(8 - e) + 0.502) stays in [-128, 127] (a plain |x| / 2^e < 1/2 would let the top 0.4 % of the range reach 128; a
// whole extra bit of headroom costs 4x in the product's relative accuracy).  Host and device must agree: one function.
__host__ __device__ inline int i8_exp_for(double mx) {
  if (!(mx > 0.0)) return 0;
  int e = ilogb(mx) + 2;
  if (ldexp(mx, -e) >= 0.498) ++e;
  return e;
}

// Balanced radix-256 digits of v (-1/2 < v < 0.498), most significant first: T = rn(v 2^56) = sum_i d_i 256^(6-i) with all d_i in
// [-128, 127].  Adding C = sum_{i>=1} 128 * 256^(6-i) turns the balanced recoding (carries) into plain byte fields:
// d_i = byte_{6-i}(T + C) - 128 = byte_{6-i}(T + C) ^ 0x80 as int8, d_0 = (T + C) >> 48: one 64-bit add, one 64-bit xor, and the
// digit bytes are the seven low bytes of the result.
// Balanced digits matter: the digit pairs with i + j >= NS are dropped, and with zero-mean digits what is dropped is zero-mean
// (with unsigned fields it is a one-sided bias that grows linearly in K: measured 50x worse).
// Exactness of the level sums: |d_i d_j| <= 2^14, at most 7 pairs per level -> K < 2^31 / (7 * 2^14) = 18724 (I8_MAX_K = 16384).
__device__ __forceinline__ unsigned long long i8_digit_bytes(long long T) {   // byte 6 - i = digit i
  static_assert(I8_NS == 7, "56-bit fixed point in 7 bytes");
  return (unsigned long long)(T + 0x0000808080808080ll) ^ 0x0000808080808080ull;
}
// T = rn(v 2^56) already in hand (the all-integer epilogue of the triangular multiply)
__device__ __forceinline__ void i8_digits_fixed(long long T, int8_t (&dg)[I8_NS]) {
  const unsigned long long V = i8_digit_bytes(T);
#pragma unroll
  for (int i = 0; i < I8_NS; ++i) dg[i] = (int8_t)(uint8_t)(V >> (8 * (I8_NS - 1 - i)));
}
// digit-plane words of four consecutive columns: pk[i] = digit i of columns 0..3 in bytes 0..3.  Two 4 x 4 byte transposes with
// PRMT (16 instructions for 28 digit bytes) instead of a shift / mask / or per byte.
__device__ __forceinline__ void i8_pack4(unsigned long long V0, unsigned long long V1, unsigned long long V2, unsigned long long V3,
                                         uint32_t (&pk)[I8_NS]) {
  const uint32_t l0 = (uint32_t)V0, l1 = (uint32_t)V1, l2 = (uint32_t)V2, l3 = (uint32_t)V3;
  const uint32_t h0 = (uint32_t)(V0 >> 32), h1 = (uint32_t)(V1 >> 32), h2 = (uint32_t)(V2 >> 32), h3 = (uint32_t)(V3 >> 32);
  const uint32_t a01 = __byte_perm(l0, l1, 0x5140), a23 = __byte_perm(l2, l3, 0x5140);   // bytes 0, 1 of each
  const uint32_t b01 = __byte_perm(l0, l1, 0x7362), b23 = __byte_perm(l2, l3, 0x7362);   // bytes 2, 3 of each
  pk[6] = __byte_perm(a01, a23, 0x5410);   // byte 0 = digit 6
  pk[5] = __byte_perm(a01, a23, 0x7632);   // byte 1 = digit 5
  pk[4] = __byte_perm(b01, b23, 0x5410);
  pk[3] = __byte_perm(b01, b23, 0x7632);
  const uint32_t c01 = __byte_perm(h0, h1, 0x5140), c23 = __byte_perm(h2, h3, 0x5140);   // bytes 4, 5
  const uint32_t d01 = __byte_perm(h0, h1, 0x7362), d23 = __byte_perm(h2, h3, 0x7362);   // bytes 6, (7)
  pk[2] = __byte_perm(c01, c23, 0x5410);   // byte 4 = digit 2
  pk[1] = __byte_perm(c01, c23, 0x7632);
  pk[0] = __byte_perm(d01, d23, 0x5410);   // byte 6 = digit 0
}
__device__ __forceinline__ unsigned long long i8_digit_bytes_of(double v) {
  return i8_digit_bytes(__double2ll_rn(v * 72057594037927936.0));
}
__device__ __forceinline__ void i8_digits(double v, int8_t (&dg)[I8_NS]) { i8_digits_fixed(__double2ll_rn(v * 72057594037927936.0), dg); }

// I8_EPI_SLICE builds its output digits in integer arithmetic (no FP64 instructions next to the UTCIMMA stream).  With k <=
// I8_K_GROUP4 the product value is (t_hi 2^24 + t_lo) 2^(-64 + e_row + e_col), t_hi / t_lo the combined level sums 0..3 / 4..6; the
// output fixed point is rn(value 2^(56 - eo)) = rn((t_hi 2^24 + t_lo) / 2^sh), sh = 8 + eo - e_row - e_col.
__device__ __forceinline__ long long i8_fx_lo(long long t, int sh) {
  if (sh <= 0) return (long long)((unsigned long long)t << min(-sh, 63));
  if (sh <= 24) return (t + (1ll << (sh - 1))) >> sh;
  return (t + (1ll << 23)) >> 24;   // in units of 2^24: joins t_hi before the final shift
}
__device__ __forceinline__ long long i8_fx_hi(long long t, long long lo, int sh) {
  if (sh <= 24) return (long long)((unsigned long long)t << min(24 - sh, 63)) + lo;
  const int s = min(sh - 24, 62);
  return (t + lo + (1ll << (s - 1))) >> s;
}
// the common case -32 < sh <= 24 (whole warp), branch-free: ((t_lo << l1) + rnd) >> r1 and (t_hi << (24 - sh)) + lo
__device__ __forceinline__ long long i8_fx_lo_fast(long long t, int l1, int r1, long long rnd) {
  return ((long long)((unsigned long long)t << l1) + rnd) >> r1;
}
__device__ __forceinline__ long long i8_fx_hi_fast(long long t, long long lo, int s2) {
  return (long long)((unsigned long long)t << s2) + lo;
}
// level sums -> one exact 64-bit integer (IMAD.WIDE chains)
__device__ __forceinline__ long long i8_mad_wide(int a, int b, long long c) {
  long long d;
  asm("mad.wide.s32 %0, %1, %2, %3;\n" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}
__device__ __forceinline__ long long i8_comb4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {   // a0 2^24 + a1 2^16 + a2 2^8 + a3
  long long t = (long long)(int)a3;
  t = i8_mad_wide((int)a2, 256, t);
  t = i8_mad_wide((int)a1, 65536, t);
  return i8_mad_wide((int)a0, 16777216, t);
}
__device__ __forceinline__ long long i8_comb3(uint32_t a0, uint32_t a1, uint32_t a2) {   // a0 2^16 + a1 2^8 + a2
  long long t = (long long)(int)a2;
  t = i8_mad_wide((int)a1, 256, t);
  return i8_mad_wide((int)a0, 65536, t);
}

template <int EPI, int BN>
__global__ void __launch_bounds__(I8_THREADS, 1)
k_gemm_i8(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const I8P p) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (an integer round-trip would turn every later access into
  // a generic load that the compiler must order against the global stores of the epilogue)
  using T = I8Tile<BN>;
  constexpr int STAGES = i8_stages<EPI, BN>(), STAGE_BYTES = T::STAGE_BYTES, B_BYTES = T::B_BYTES, NBUF = T::NBUF;
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;            // [NBUF]
  uint64_t* tmem_empty = tmem_full + 2;            // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  double* xs_all = reinterpret_cast<double*>(base + STAGES * STAGE_BYTES + 256);   // per team: [BN][d] x rows of the tile, then [BN] y
  double* wstage = xs_all + I8_EPI_SMEM / 8;                                        // [8 warps][32 rows][I8_WSTAGE_LD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  if (p.dbg && threadIdx.x == 0) {   // developer timeline: wall-clock span of every CTA
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg[3 * I8_DBG_ITEMS * 4 + 2 * blockIdx.x] = (long long)gt;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], T::TEAM_WARPS);
    }
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
      for (int rnd = 0, nrnd = i8_rounds(p, G); rnd < nrnd; ++rnd) {
        const int w = i8_item(p, rnd, blockIdx.x, G);
        if (w >= p.total) continue;
        I8Item it;
        i8_decode(p, w, it);
        for (int kb = it.kb_lo; kb < it.kb_hi; ++kb, ++n) {
          const int s = n % STAGES;
          if (n >= STAGES) i8_mbar_wait(&empty[s], ((n / STAGES) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], (p.exp_skip_a ? 0 : I8_NS * I8_A_BYTES) + (p.exp_skip_b ? 0 : I8_NS * B_BYTES));
          unsigned char* st = base + s * STAGE_BYTES;
#ifdef GGP_I8_TMA_PLANES7
          // one box per operand and stage covering all NS digit planes (k x rows x NS): 2 TMA instructions instead of 14
          if (!p.exp_skip_a) i8_tma_load_3d(st, &tmA, kb * I8_BKB, it.tm * I8_BM, it.chunk * I8_NS, &full[s]);
          if (p.b_mn) {
            const int cb = p.b_chunk > 0 ? (it.tn * BN) / p.b_chunk : 0;
            i8_tma_load_3d(st + I8_NS * I8_A_BYTES, &tmB, it.tn * BN - cb * p.b_chunk, kb * I8_BKB, cb * I8_NS, &full[s]);
          } else if (!p.exp_skip_b) {
            i8_tma_load_3d(st + I8_NS * I8_A_BYTES, &tmB, kb * I8_BKB, it.tn * BN, it.chunk * I8_NS, &full[s]);
          }
#else
          if (!p.exp_skip_a) {
#pragma unroll
            for (int i = 0; i < I8_NS; ++i) i8_tma_load_3d(st + i * I8_A_BYTES, &tmA, kb * I8_BKB, it.tm * I8_BM, it.chunk * I8_NS + i, &full[s]);
          }
          if (p.b_mn) {   // box = 64 n-bytes x 64 k-rows of plane (column block, digit)
            const int cb = p.b_chunk > 0 ? (it.tn * BN) / p.b_chunk : 0;
            const int cn = it.tn * BN - cb * p.b_chunk;
#pragma unroll
            for (int j = 0; j < I8_NS; ++j)
              i8_tma_load_3d(st + I8_NS * I8_A_BYTES + j * B_BYTES, &tmB, cn, kb * I8_BKB, cb * I8_NS + j, &full[s]);
          } else if (!p.exp_skip_b) {
#pragma unroll
            for (int j = 0; j < I8_NS; ++j)
              i8_tma_load_3d(st + I8_NS * I8_A_BYTES + j * B_BYTES, &tmB, kb * I8_BKB, it.tn * BN, it.chunk * I8_NS + j, &full[s]);
          }
#endif
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int n = 0, item = 0;
    for (int rnd = 0, nrnd = i8_rounds(p, G); rnd < nrnd; ++rnd) {
      const int w = i8_item(p, rnd, blockIdx.x, G);
      if (w >= p.total) continue;
      I8Item it;
      i8_decode(p, w, it);
      if (it.kb_hi <= it.kb_lo) continue;
      if (lane == 0) I8_STAMP(0, item, 0);
      const int buf = item & (NBUF - 1);
      const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * T::BUF_COLS);
      // the epilogue team of this accumulator set has drained the tile that used it last (NBUF tiles ago)
      if (item >= NBUF) i8_mbar_wait(&tmem_empty[buf], ((item / NBUF) - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      if (lane == 0) I8_STAMP(0, item, 1);
      for (int kb = it.kb_lo; kb < it.kb_hi; ++kb, ++n) {
        const int s = n % STAGES;
        i8_mbar_wait(&full[s], (n / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        if (lane == 0) {
          const uint32_t sa = smem_u32(base + s * STAGE_BYTES), sb = sa + I8_NS * I8_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < I8_BKB / 32; ++kk) {
#pragma unroll
            for (int i = 0; i < I8_NS; ++i) {
              const uint64_t ad = i8_desc_sw64(sa + i * I8_A_BYTES + kk * 32);
              // digit i of A x digits 0 .. NS-1-i of B: one tall B tile of ncols rows -> accumulator columns [i * BN, i * BN + ncols);
              // more than 256 columns are issued as two equal halves (multiples of 16 columns and of the 8-row swizzle group)
              const int ncols = BN * (I8_NS - i);
              if (BN == 32) {
                // 128 x 32 tiles: ncols <= 224, one MMA per A digit; B is the tall K-major tile (7 x 32 rows) or, MN-major, 32-byte
                // MN blocks (one per digit) LBO = B_BYTES apart with the k-step advancing 32 rows of 32 bytes
                const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (p.b_mn ? (1u << 16) : 0u) | ((uint32_t)(ncols >> 3) << 17) |
                                       ((uint32_t)(I8_BM >> 4) << 24);
                const uint64_t bd = p.b_mn ? i8_desc_mn_sw32(sb + kk * 32 * BN, B_BYTES) : i8_desc_sw64(sb + kk * 32);
                i8_umma(tmem_acc + (uint32_t)(i * BN), ad, bd, idesc, (kb > it.kb_lo || kk > 0 || i > 0) ? 1u : 0u);
                continue;
              }
              const int nsplit = ncols > 256 ? 2 : 1, nn = ncols / nsplit;
              if (p.b_mn) {
                // MN-major B: the digit tiles are 64-byte MN blocks LBO = 4096 bytes apart, so a wide MMA must start on a tile
                // boundary: 448 = 256 + 192, 384 = 192 + 192, 320 = 192 + 128 columns; the k-step advances 32 rows = 2048 bytes
                const int n1 = ncols > 256 ? ((ncols / 2 + 63) / 64) * 64 : ncols;
#pragma unroll
                for (int hs = 0; hs < nsplit; ++hs) {
                  const int off = hs * n1, nw = hs ? ncols - n1 : n1;
                  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(nw >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
                  const uint64_t bd = i8_desc_mn_sw64(sb + (off / BN) * B_BYTES + kk * 32 * BN, B_BYTES);
                  i8_umma(tmem_acc + (uint32_t)(i * BN + off), ad, bd, idesc, (kb > it.kb_lo || kk > 0 || i > 0) ? 1u : 0u);
                }
                continue;
              }
#pragma unroll
              for (int hs = 0; hs < nsplit; ++hs) {
                const int off = hs * nn;
                // D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
                const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
                const uint64_t bd = i8_desc_sw64(sb + off * I8_BKB + kk * 32);   // row `off` of the tall tile: 8-row groups of 512 bytes
                i8_umma(tmem_acc + (uint32_t)(i * BN + off), ad, bd, idesc, (kb > it.kb_lo || kk > 0 || i > 0) ? 1u : 0u);
              }
            }
          }
          i8_umma_commit(&empty[s]);                            // frees the stage once the MMAs that read it are done
          if (kb == it.kb_hi - 1) i8_umma_commit(&tmem_full[buf]);    // accumulators of this tile complete
          if (kb == it.kb_lo) I8_STAMP(0, item, 2);
        }
        __syncwarp();
      }
      if (lane == 0) I8_STAMP(0, item, 3);
      ++item;
    }
  } else {
    // ================= epilogue (8 warps; warp w owns TMEM lanes [32 (w % 4), +32) and 32 tile columns) =================
    // BN = 64: one team of 8 warps, `half` = column half of the tile.  BN = 32: two teams of 4 warps, `half` = team = parity of the
    // tiles (accumulator sets) the warp serves; every warp covers all 32 columns of its tiles.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;   // 0..255
    constexpr int TT = 256 / NBUF;     // threads of a team
    const int tt = et & (TT - 1), team_bar = 1 + (NBUF == 2 ? half : 0);
    const int chalf = (BN == 64) ? half * I8_EC : 0;   // first tile column of this warp
    double* xs = xs_all + (NBUF == 2 ? half * (32 * (I8_MAX_D + 1)) : 0);   // [BN][d]: x rows of the team's tile
    double* ys = xs + BN * I8_MAX_D;                                         // [BN]
    const int p_ea0 = p.ea0_sel ? p.e_dev[p.ea0_sel - 1] : p.ea0, p_eb0 = p.eb0_sel ? p.e_dev[p.eb0_sel - 1] : p.eb0,
              p_eo = p.eo_sel ? p.e_dev[p.eo_sel - 1] : p.eo;
    double accT[I8_EC];                // I8_EPI_F64 with nchunk: this thread's 32 tile elements summed over the CTA's chunks
#pragma unroll
    for (int c = 0; c < I8_EC; ++c) accT[c] = 0.0;
    double rdA[2] = {0.0, 0.0};        // I8_EPI_SLICE with rowdot_reg: b-partials of this thread's row in the CTA's two row tiles
    double rsA = 0.0;                  // I8_EPI_MOMENTS with mom_accum: row sum of W (the constant moment) of this thread's row
#ifndef GGP_I8_NO_MOM_FRAG
    // FRAGMENT mode of the one-launch backward plan (mom_accum, k <= I8_K_GROUP4): the accumulators are drained with the 16x256b shape,
    // so W = (G + u y^T) o K sits in the registers in DMMA A-fragment layout and goes straight back to the FP64 tensor pipe -- no
    // shared-memory patch, no warp barriers on the serial part of the tile (it was 7 k of the 8.5 k clk between "drained" and "done")
    const bool mom_frag = (EPI == I8_EPI_MOMENTS) && p.mom_accum && p.K <= I8_K_GROUP4 && !p.eb && !p.mom_frag_off;
    double rs4[4] = {0.0, 0.0, 0.0, 0.0};   // partial row sums of W (this thread's 8 columns) for the rows 8 G + g
#else
    constexpr bool mom_frag = false;
#endif
#ifdef GGP_I8_MOM_VEC
    // ... and, d <= 8, this thread's row of the x / x^2 moments on the VECTOR FP64 pipe (thread = TMEM lane = tile row: no staging of W)
    double mvx[8], mvx2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) mvx[q] = mvx2[q] = 0.0;
    const bool mom_vec = (EPI == I8_EPI_MOMENTS) && p.mom_accum;   // the host sets mom_accum only for d <= 8 in this build
#else
    constexpr bool mom_vec = false;
#endif
    double cmA[4][3][2];               // ... and the x / x^2 moments of this warp's rows, over all tiles of the CTA
#pragma unroll
    for (int G4 = 0; G4 < 4; ++G4)
#pragma unroll
      for (int B = 0; B < 3; ++B) cmA[G4][B][0] = cmA[G4][B][1] = 0.0;
    int item = 0;
    for (int rnd = 0, nrnd = i8_rounds(p, G); rnd < nrnd; ++rnd) {
      const int w = i8_item(p, rnd, blockIdx.x, G);
      if (w >= p.total) continue;
      I8Item it;
      i8_decode(p, w, it);
      if (it.kb_hi <= it.kb_lo) continue;
      const int my_item = item++;
      const int buf = my_item & (NBUF - 1);
      if (NBUF == 2 && buf != half) continue;       // the other team's tile
      const int row = it.tm * I8_BM + quarter * 32 + lane;
      const int col0 = it.tn * BN + chalf;          // first column of this warp's 32
      const int colt = it.tn * BN;                  // first column of the tile
      double kv0[16], kv1[16];   // I8_EPI_MOMENTS: Kmul values of column groups, loaded ahead of their use (HBM latency)
      auto load_kv = [&](double (&kv)[16], int g) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int gc = col0 + g * 16 + c;
          kv[c] = (row < p.M && gc < p.N) ? __ldg(p.Kmul + (int64_t)gc * p.ldk + row) : 0.0;
        }
      };
      if (EPI == I8_EPI_MOMENTS) {
        // stage the tile's x rows and y into shared memory (previous tile's readers are past their last read: barrier below)
        asm volatile("bar.sync %0, %1;\n" ::"r"(team_bar), "r"(TT) : "memory");
        if (et == 0) I8_STAMP(2, my_item, 0);
        const int d = p.d;
        {   // the tile's x rows are one contiguous block of 64 d doubles: coalesced copy, 4 independent loads in flight per thread
          const double* src = p.Xc + (int64_t)colt * d;
          const int cnt = BN * d, lim = max(0, min(cnt, (p.N - colt) * d));
          for (int i0 = tt; i0 < cnt; i0 += 4 * TT) {
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int i = i0 + j * TT; v[j] = (i < lim) ? __ldg(src + i) : 0.0; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = i0 + j * TT;
              if (i < cnt) {
                xs[i] = v[j];
                if (mom_vec || (mom_frag && 2 * d <= I8_MAX_D)) xs[cnt + i] = v[j] * v[j];   // squares behind the rows: 2 BN d <= BN I8_MAX_D
              }
            }
          }
          if (tt < BN) ys[tt] = (colt + tt < p.N) ? __ldg(p.yv + colt + tt) : 0.0;
        }
        if (et == 0) I8_STAMP(2, my_item, 1);
        // both groups of this warp's 32 columns are in flight while the MMAs of this tile run.  (With the register file full these 32
        // loads issue one memory latency at a time -- 16 k clk of this 25 k clk phase -- which is hidden behind the 37 k clk mainloop in
        // the serial-epilogue mode.  An overlapped variant with an L2 prefetch here and the loads after the drain cut the phase to
        // 6.5 k clk, but the FP64 work next to the UTCIMMA stream then took 39 k clk: no gain, removed.)
        if (!mom_frag) {
          load_kv(kv0, 0);
          load_kv(kv1, 1);
        }
        if (et == 0) I8_STAMP(2, my_item, 2);
        asm volatile("bar.sync %0, %1;\n" ::"r"(team_bar), "r"(TT) : "memory");
        if (et == 0) I8_STAMP(2, my_item, 3);
      }
#ifndef GGP_I8_NO_MOM_FRAG
      if (EPI == I8_EPI_MOMENTS && mom_frag) {
        const int g = lane >> 2, q4 = lane & 3, d = p.d;
        // this thread's four rows (8 G + g of the warp's 32) and their multipliers, fetched while the MMAs of the tile run
        double sc4[4], ui4[4], kvf[32];
#pragma unroll
        for (int G4 = 0; G4 < 4; ++G4) {
          const int rr = it.tm * I8_BM + quarter * 32 + 8 * G4 + g;
          const int er = p.ea ? p.ea[min(rr, p.M - 1)] : p_ea0;
          sc4[G4] = p.alpha * exp2((double)(er + p_eb0));
          ui4[G4] = rr < p.M ? __ldg(p.u + rr) : 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int gc = col0 + 8 * j + 2 * q4 + e;
              // element index r = 4 j + 2 (G4 & 1) + e of lane half h = G4 >> 1  ->  kvf[16 h + r]
              kvf[16 * (G4 >> 1) + 4 * j + 2 * (G4 & 1) + e] = (rr < p.M && gc < p.N) ? __ldg(p.Kmul + (int64_t)gc * p.ldk + rr) : 0.0;
            }
        }
        if (et == 0) I8_STAMP(1, my_item, 0);
        i8_mbar_wait(&tmem_full[buf], (my_item / NBUF) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        if (et == 0) I8_STAMP(1, my_item, 1);
        double wf[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32 + 16 * h) << 16) + (uint32_t)(buf * T::BUF_COLS + chalf);
          {   // levels 4..6
            uint32_t v0[16], v1[16], v2[16];
            i8_tmem_ld_frag(ta + 4 * BN, v0);
            i8_tmem_ld_frag(ta + 5 * BN, v1);
            i8_tmem_ld_frag(ta + 6 * BN, v2);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int r = 0; r < 16; ++r)   // the row scale rides on the conversion constants: sc 2^-64, sc 2^-40
              wf[16 * h + r] = (double)i8_comb3(v0[r], v1[r], v2[r]) * (sc4[2 * h + ((r >> 1) & 1)] * 5.421010862427522e-20);
          }
          {   // levels 0..3
            uint32_t v0[16], v1[16], v2[16], v3[16];
            i8_tmem_ld_frag(ta, v0);
            i8_tmem_ld_frag(ta + BN, v1);
            i8_tmem_ld_frag(ta + 2 * BN, v2);
            i8_tmem_ld_frag(ta + 3 * BN, v3);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int r = 0; r < 16; ++r)
              wf[16 * h + r] = fma((double)i8_comb4(v0[r], v1[r], v2[r], v3[r]), sc4[2 * h + ((r >> 1) & 1)] * 9.094947017729282e-13, wf[16 * h + r]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
        __syncwarp();
        if (lane == 0 && !p.serial_epi) i8_mbar_arrive(&tmem_empty[buf]);
        if (et == 0) I8_STAMP(1, my_item, 2);
        // W = (alpha acc 2^(e_row + e_col) + u[row] y[col]) * Kmul[col][row]
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int idx = 16 * h + 4 * j + 2 * r2 + e, G4 = 2 * h + r2;
                const double w = fma(ui4[G4], ys[chalf + 8 * j + 2 * q4 + e], wf[idx]) * kvf[idx];
                wf[idx] = w;
                rs4[G4] += w;
              }
        // moments 1 .. 2 d: column 8 j + 2 q + e plays k = q of the DMMA (A and B use the same permutation of k), moment 8 B + g + 1 is n
        const bool sq_staged = 2 * d <= I8_MAX_D;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double* xr = xs + (chalf + 8 * j + 2 * q4 + e) * d;
#pragma unroll
            for (int B = 0; B < 3; ++B) {
              if (B * 8 + 1 > 2 * d) continue;   // warp-uniform
              const int m = B * 8 + g + 1;
              double b = 0.0;
              if (m <= d) b = xr[m - 1];
              else if (m <= 2 * d) { if (sq_staged) b = xr[BN * d + m - 1 - d]; else { const double x = xr[m - 1 - d]; b = x * x; } }
#pragma unroll
              for (int G4 = 0; G4 < 4; ++G4)
                dmma884(cmA[G4][B][0], cmA[G4][B][1], wf[16 * (G4 >> 1) + 4 * j + 2 * (G4 & 1) + e], b);
            }
          }
        if (p.serial_epi) {
          __syncwarp();
          if (lane == 0) i8_mbar_arrive(&tmem_empty[buf]);
        }
        if (et == 0) I8_STAMP(1, my_item, 3);
        continue;
      }
#endif
      const int e_r = p.ea ? p.ea[min(row, p.M - 1)] : p_ea0;
      const int sh = 8 + p_eo - e_r - p_eb0;   // I8_EPI_SLICE (scalar column exponent, alpha = 1, k <= I8_K_GROUP4: checked on the host)
      const bool fx_fast = (EPI == I8_EPI_SLICE) && __all_sync(0xffffffffu, sh > -32 && sh <= 24);
      const int fx_l1 = sh < 0 ? -sh : 0, fx_r1 = sh > 0 ? sh : 0, fx_s2 = (24 - sh) & 63;
      const long long fx_rnd = sh > 0 ? 1ll << ((sh - 1) & 63) : 0ll;
      // everything the serial part after the drain needs and that does not depend on the accumulators is fetched / computed BEFORE the
      // wait (the software exp2 and the dependent global load of u[row] were ~1.5 k clk on the tile's critical path)
      const double sc_pre = (EPI != I8_EPI_SLICE && !p.eb) ? p.alpha * exp2((double)(e_r + p_eb0)) : 0.0;
      const double ui_pre = (EPI == I8_EPI_MOMENTS && row < p.M) ? __ldg(p.u + row) : 0.0;
      if (et == 0) I8_STAMP(1, my_item, 0);
      i8_mbar_wait(&tmem_full[buf], (my_item / NBUF) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      if (et == 0) I8_STAMP(1, my_item, 1);
      if (p.exp_no_epi) {   // developer experiment: mainloop only
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
        __syncwarp();
        if (lane == 0) i8_mbar_arrive(&tmem_empty[buf]);
        continue;
      }
      double acc[I8_EC];
      long long fx[I8_EC];   // I8_EPI_SLICE: rn(value 2^(56 - eo)), integer arithmetic only
#pragma unroll
      for (int c = 0; c < I8_EC; ++c) { acc[c] = 0.0; fx[c] = 0; }
      // The seven level sums of a column are combined exactly in 64-bit integers before ONE int->double conversion per group (or, in
      // the all-integer epilogue, two shifts).  value = U 2^(-64 + e_row + e_col), U = sum_l a_l 2^(8 (6 - l)).
      //   k <= I8_K_GROUP4 (MODE 0 / 1 = integer epilogue, fast / general shifts): U = t_hi 2^24 + t_lo, t_hi = levels 0..3, t_lo = levels 4..6
      //   longer k          (MODE 2):  U = g0 2^32 + g1 2^8 + g2, g0 = levels 0..2, g1 = levels 3..5, g2 = level 6   (each < 2^53)
      static_assert(I8_NS == 7, "level grouping");
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * T::BUF_COLS + chalf);
      auto drain = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
        for (int c0 = 0; c0 < I8_EC; c0 += 16) {
          if (MODE != 2) {
            {   // levels 4..6
              uint32_t v0[16], v1[16], v2[16];
              const uint32_t ta = tbase + (uint32_t)(4 * BN + c0);
              i8_tmem_ld16(ta, v0);
              i8_tmem_ld16(ta + BN, v1);
              i8_tmem_ld16(ta + 2 * BN, v2);
              asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const long long t = i8_comb3(v0[c], v1[c], v2[c]);
                if (EPI == I8_EPI_SLICE) fx[c0 + c] = (MODE == 0) ? i8_fx_lo_fast(t, fx_l1, fx_r1, fx_rnd) : i8_fx_lo(t, sh);
                else acc[c0 + c] = (double)t * 5.421010862427522e-20 /* 2^-64 */;
              }
            }
            {   // levels 0..3
              uint32_t v0[16], v1[16], v2[16], v3[16];
              const uint32_t ta = tbase + (uint32_t)c0;
              i8_tmem_ld16(ta, v0);
              i8_tmem_ld16(ta + BN, v1);
              i8_tmem_ld16(ta + 2 * BN, v2);
              i8_tmem_ld16(ta + 3 * BN, v3);
              asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const long long t = i8_comb4(v0[c], v1[c], v2[c], v3[c]);
                if (EPI == I8_EPI_SLICE) fx[c0 + c] = (MODE == 0) ? i8_fx_hi_fast(t, fx[c0 + c], fx_s2) : i8_fx_hi(t, fx[c0 + c], sh);
                else acc[c0 + c] = fma((double)t, 9.094947017729282e-13 /* 2^-40 */, acc[c0 + c]);
              }
            }
          } else {
            {   // levels 3..5 and 6
              uint32_t v0[16], v1[16], v2[16], v3[16];
              const uint32_t ta = tbase + (uint32_t)(3 * BN + c0);
              i8_tmem_ld16(ta, v0);
              i8_tmem_ld16(ta + BN, v1);
              i8_tmem_ld16(ta + 2 * BN, v2);
              i8_tmem_ld16(ta + 3 * BN, v3);
              asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const long long t = i8_comb3(v0[c], v1[c], v2[c]);
                acc[c0 + c] = fma((double)t, 1.3877787807814457e-17 /* 2^-56 */, (double)(int)v3[c] * 5.421010862427522e-20 /* 2^-64 */);
              }
            }
            {   // levels 0..2
              uint32_t v0[16], v1[16], v2[16];
              const uint32_t ta = tbase + (uint32_t)c0;
              i8_tmem_ld16(ta, v0);
              i8_tmem_ld16(ta + BN, v1);
              i8_tmem_ld16(ta + 2 * BN, v2);
              asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const long long t = i8_comb3(v0[c], v1[c], v2[c]);
                acc[c0 + c] = fma((double)t, 2.3283064365386963e-10 /* 2^-32 */, acc[c0 + c]);
              }
            }
          }
        }
      };
      if (EPI == I8_EPI_SLICE) {
        if (fx_fast) drain(std::integral_constant<int, 0>{});
        else drain(std::integral_constant<int, 1>{});
      } else if (p.K <= I8_K_GROUP4) {
        drain(std::integral_constant<int, 0>{});
      } else {
        drain(std::integral_constant<int, 2>{});
      }
      // accumulators are in registers: hand TMEM back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
      __syncwarp();
      if (lane == 0 && !p.serial_epi) i8_mbar_arrive(&tmem_empty[buf]);
      if (et == 0) I8_STAMP(1, my_item, 2);
      const int item_done = my_item;

      const bool rok = row < p.M;
      if (EPI == I8_EPI_SLICE) {
      } else if (p.eb) {
#pragma unroll
        for (int c = 0; c < I8_EC; ++c) acc[c] = p.alpha * ldexp(acc[c], e_r + p.eb[min(col0 + c, p.N - 1)]);
      } else {
        const double sc = sc_pre;
#pragma unroll
        for (int c = 0; c < I8_EC; ++c) acc[c] *= sc;
      }

      if (EPI == I8_EPI_F64 && p.nchunk) {
#pragma unroll
        for (int c = 0; c < I8_EC; ++c) accT[c] += acc[c];
      } else if (EPI == I8_EPI_F64) {
        if (rok) {
          double* dst = p.C + (int64_t)it.split * p.sSplit + (int64_t)row * p.ldc + col0;
          if (col0 + I8_EC <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            if (p.beta != 0.0) {   // read-modify-write in groups of 8 independent 16-byte loads (not one latency per element pair)
#pragma unroll
              for (int c0 = 0; c0 < I8_EC; c0 += 16) {
                double2 o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = *reinterpret_cast<const double2*>(dst + c0 + 2 * j);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<double2*>(dst + c0 + 2 * j) =
                      make_double2(fma(p.beta, o[j].x, acc[c0 + 2 * j]), fma(p.beta, o[j].y, acc[c0 + 2 * j + 1]));
              }
            } else {
#pragma unroll
              for (int c = 0; c < I8_EC; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(acc[c], acc[c + 1]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < I8_EC; ++c)
              if (col0 + c < p.N) dst[c] = acc[c] + (p.beta != 0.0 ? p.beta * dst[c] : 0.0);
          }
        }
      } else if (EPI == I8_EPI_SLICE) {
        if (p.rowdot) {
          // fused b-partials A y from the QUANTISED A (the digits just built): S = A A^T is formed from those digits, and dF/dKzx =
          // P Kzx + u y^T cancels by cond(Kzz), so b must belong to the same A (b = L^{-1} (Kzx y) computed on the side, even in
          // double-double, moved the gradient by 8e-8 at the headline shape; consistent b: 5e-9 against the FP64 path)
          const double so = exp2((double)(p_eo - 56));
#pragma unroll
          for (int c = 0; c < I8_EC; ++c) acc[c] = (double)fx[c] * so;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
          for (int c = 0; c < I8_EC; c += 4) {
            s0 = fma(acc[c], (col0 + c < p.N) ? __ldg(p.yv + col0 + c) : 0.0, s0);
            s1 = fma(acc[c + 1], (col0 + c + 1 < p.N) ? __ldg(p.yv + col0 + c + 1) : 0.0, s1);
            s2 = fma(acc[c + 2], (col0 + c + 2 < p.N) ? __ldg(p.yv + col0 + c + 2) : 0.0, s2);
            s3 = fma(acc[c + 3], (col0 + c + 3 < p.N) ? __ldg(p.yv + col0 + c + 3) : 0.0, s3);
          }
          const double sv = (s0 + s1) + (s2 + s3);
          if (p.rowdot_reg) {
            rdA[rnd & 1] += sv;
          } else if (rok) {   // one slab per 32 columns
            double* rd = p.rowdot + (int64_t)(BN == 64 ? it.tn * 2 + half : it.tn) * p.M + row;
            *rd = p.rowdot_acc ? *rd + sv : sv;
          }
        }
        // digit planes of the tile row: this thread's 32 columns are 32 consecutive bytes per plane = one full sector, written with one
        // 256-bit store (two 16-byte stores made every sector a pair of partial writes; the kernel runs at the L2 throughput cap).
        // Columns beyond N are zero because their B rows are zero.
        {
          uint32_t pk[I8_NS][I8_EC / 4];
#pragma unroll
          for (int c = 0; c < I8_EC; c += 4) {
            uint32_t w[I8_NS];
            i8_pack4(i8_digit_bytes(fx[c]), i8_digit_bytes(fx[c + 1]), i8_digit_bytes(fx[c + 2]), i8_digit_bytes(fx[c + 3]), w);
#pragma unroll
            for (int i = 0; i < I8_NS; ++i) pk[i][c >> 2] = w[i];
          }
          static_assert(I8_EC == 32, "one 32-byte sector per thread and plane");
          if (rok) {
            int8_t* dst = p.Oq + (int64_t)row * p.o_ld + col0;
            if (p.o_chunk > 0) {
              const int cb = col0 / p.o_chunk;
              dst = p.Oq + (int64_t)cb * p.o_chunk_stride + (int64_t)row * p.o_ld + (col0 - cb * p.o_chunk);
            }
            if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0 && (p.o_plane & 31) == 0) {
#pragma unroll
              for (int i = 0; i < I8_NS; ++i)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dst + (int64_t)i * p.o_plane), "r"(pk[i][0]),
                             "r"(pk[i][1]), "r"(pk[i][2]), "r"(pk[i][3]), "r"(pk[i][4]), "r"(pk[i][5]), "r"(pk[i][6]), "r"(pk[i][7])
                             : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < I8_NS; ++i) {
                *reinterpret_cast<uint4*>(dst + (int64_t)i * p.o_plane) = make_uint4(pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
                *reinterpret_cast<uint4*>(dst + (int64_t)i * p.o_plane + 16) = make_uint4(pk[i][4], pk[i][5], pk[i][6], pk[i][7]);
              }
            }
          }
        }
      } else {
        // W = (G + u y^T) o Kmul, then the moments against [1, x, x^2] of the tile's 64 columns
        const int d = p.d, nq = 2 * d + 1;
        const double ui = ui_pre;
        auto apply_kv = [&](const double (&kv)[16], int g) {
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[g * 16 + c] = fma(ui, ys[chalf + g * 16 + c], acc[g * 16 + c]) * kv[c];
        };
        apply_kv(kv0, 0);
        apply_kv(kv1, 1);
        // mom[row][:] = sum_c W[row][c] * Phi[c][:], Phi = [1, x, x^2]: a 128 x 64 x (2d+1) product.  The vector FP64 pipe is far too
        // slow for it (ncu: math-pipe throttle, the epilogue took longer than the MMAs of the next tile), so it goes to the FP64
        // tensor pipe, idle in this kernel: W is staged 8 columns at a time through a private 32 x 8 shared-memory patch per warp
        // (row stride 12 doubles: conflict-free 64-bit fragment loads) and fed to DMMA.8x8x4 as the A operand.
        // lane = 4g + q:  a = W[8G + g][4kq + q],  b = Phi[4kq + q][8B + g],  (c0, c1) = mom[8G + g][8B + 2q, + 1]
        const int g = lane >> 2, q4 = lane & 3;
        double* wsm = wstage + (warp - 2) * 32 * I8_WSTAGE_LD;   // private patch of this warp
        // off = 1 (mom_accum): the constant moment (row sums of W) is summed on the vector pipe instead, the DMMA blocks start at x_0
        // and a block with no moment left is skipped (d = 8: 2 blocks of 8 instead of 3, a third fewer DMMAs on the serial path)
        auto sweep = [&](int b0, double (&cm)[4][3][2], int off) {   // cm += W (this warp's 32 x 32 patch) x Phi[:, off + 8 b0 .. + 24)
#pragma unroll
          for (int pc = 0; pc < I8_EC / 8; ++pc) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j += 2)
              *reinterpret_cast<double2*>(wsm + lane * I8_WSTAGE_LD + j) = make_double2(acc[pc * 8 + j], acc[pc * 8 + j + 1]);
            __syncwarp();
#pragma unroll
            for (int kq = 0; kq < 2; ++kq) {
              double a[4];
#pragma unroll
              for (int G = 0; G < 4; ++G) a[G] = wsm[(8 * G + g) * I8_WSTAGE_LD + 4 * kq + q4];
              const double* xr = xs + (chalf + pc * 8 + 4 * kq + q4) * d;
#pragma unroll
              for (int B = 0; B < 3; ++B) {
                if ((b0 + B) * 8 + off > 2 * d) continue;   // warp-uniform
                const int m = (b0 + B) * 8 + g + off;
                double b = 0.0;
                if (m == 0) b = 1.0;
                else if (m <= d) b = xr[m - 1];
                else if (m <= 2 * d) { const double x = xr[m - 1 - d]; b = x * x; }
#pragma unroll
                for (int G = 0; G < 4; ++G) dmma884(cm[G][B][0], cm[G][B][1], a[G], b);
              }
            }
          }
        };
#ifdef GGP_I8_MOM_VEC
        if (mom_vec) {
          // mom[row][1 + q] += W[row][c] x_c[q], mom[row][1 + d + q] += W[row][c] x_c[q]^2: 2 d independent FMA chains per thread, operands
          // broadcast from shared memory; same FP64 rate as the DMMA route (64 FMA / clk / SM) without the W patch round trip
          const double* xr = xs + chalf * d;
          const double* xr2 = xr + BN * d;
#pragma unroll
          for (int c = 0; c < I8_EC; ++c) {
            const double w = acc[c];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < d) {
                mvx[q] = fma(w, xr[c * d + q], mvx[q]);
                mvx2[q] = fma(w, xr2[c * d + q], mvx2[q]);
              }
          }
          double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
          for (int c = 0; c < I8_EC; c += 4) { r0 += acc[c]; r1 += acc[c + 1]; r2 += acc[c + 2]; r3 += acc[c + 3]; }
          rsA += (r0 + r1) + (r2 + r3);
        } else
#else
        if (p.mom_accum) {
          sweep(0, cmA, 1);   // one sweep covers the 2 d <= 23 non-constant moments; written after the CTA's last tile
          double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
          for (int c = 0; c < I8_EC; c += 4) { r0 += acc[c]; r1 += acc[c + 1]; r2 += acc[c + 2]; r3 += acc[c + 3]; }
          rsA += (r0 + r1) + (r2 + r3);
        } else
#endif
        {
          for (int b0 = 0; b0 * 8 < nq; b0 += 3) {   // three blocks of 8 moments per sweep (one sweep for d <= 11)
            double cm[4][3][2];
#pragma unroll
            for (int G = 0; G < 4; ++G)
#pragma unroll
              for (int B = 0; B < 3; ++B) cm[G][B][0] = cm[G][B][1] = 0.0;
            sweep(b0, cm, 0);
#pragma unroll
            for (int G = 0; G < 4; ++G) {
              const int rr = it.tm * I8_BM + quarter * 32 + 8 * G + g;
              if (rr >= p.M) continue;
              double* mo = p.mom + (int64_t)(BN == 64 ? it.tn * 2 + half : it.tn) * p.sMomTile + (int64_t)rr * nq;   // one slab per 32 columns
#pragma unroll
              for (int B = 0; B < 3; ++B) {
                const int m = (b0 + B) * 8 + 2 * q4;
                if (m < nq) mo[m] = cm[G][B][0];
                if (m + 1 < nq) mo[m + 1] = cm[G][B][1];
              }
            }
          }
        }
      }
      if (p.serial_epi) {
        __syncwarp();
        if (lane == 0) i8_mbar_arrive(&tmem_empty[buf]);
      }
      if (et == 0) I8_STAMP(1, item_done, 3);
    }
    if (EPI == I8_EPI_F64 && BN == 64 && p.nchunk && blockIdx.x / p.ntile < p.nchunk) {   // the tile total of this CTA's chunks, stored once
      I8Item it;
      i8_decode(p, blockIdx.x % p.ntile, it);
      const int rown = it.tm * I8_BM + quarter * 32 + lane, col0 = it.tn * BN + half * I8_EC;
      if (rown < p.M) {
        double* dst = p.C + (int64_t)(blockIdx.x / p.ntile) * p.sSplit + (int64_t)rown * p.ldc + col0;
#pragma unroll
        for (int c = 0; c < I8_EC; ++c)
          if (col0 + c < p.N) dst[c] = accT[c];
      }
    }
    if (EPI == I8_EPI_SLICE && p.rowdot && p.rowdot_reg) {   // slab 2 b + half (zeroed by the host): rows of the CTA's two row tiles
      int tmv[2] = {-1, -1};
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int w = i8_item(p, par, blockIdx.x, G);
        if (par < i8_rounds(p, G) && w < p.total) { I8Item it; i8_decode(p, w, it); tmv[par] = it.tm; }
      }
      double* slab = p.rowdot + (int64_t)(blockIdx.x * 2 + half) * p.M;
      if (tmv[0] >= 0 && tmv[0] == tmv[1]) { rdA[0] += rdA[1]; tmv[1] = -1; }
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int rown = tmv[par] * I8_BM + quarter * 32 + lane;
        if (tmv[par] >= 0 && rown < p.M) slab[rown] = rdA[par];
      }
    }
    if (EPI == I8_EPI_MOMENTS && p.mom_accum) {
      const int nq = 2 * p.d + 1, g = lane >> 2, q4 = lane & 3;
      const int tm = blockIdx.x % p.tiles_m, slab = (blockIdx.x / p.tiles_m) * 2 + half;
#ifdef GGP_I8_MOM_VEC
      if (mom_vec) {
        const int rr = tm * I8_BM + quarter * 32 + lane;
        if (rr < p.M) {
          double* mo = p.mom + (int64_t)slab * p.sMomTile + (int64_t)rr * nq;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < p.d) { mo[1 + q] = mvx[q]; mo[1 + p.d + q] = mvx2[q]; }
        }
      }
#else
#pragma unroll
      for (int G4 = 0; G4 < 4; ++G4) {
        const int rr = tm * I8_BM + quarter * 32 + 8 * G4 + g;
        if (rr >= p.M) continue;
        double* mo = p.mom + (int64_t)slab * p.sMomTile + (int64_t)rr * nq;
#pragma unroll
        for (int B = 0; B < 3; ++B) {
          const int m = B * 8 + 2 * q4 + 1;
          if (m < nq) mo[m] = cmA[G4][B][0];
          if (m + 1 < nq) mo[m + 1] = cmA[G4][B][1];
        }
      }
#endif
      if (mom_frag) {   // row sums: this thread holds the partial over its 8 columns per tile; the 4 lanes of a quad cover the 32 columns
#ifndef GGP_I8_NO_MOM_FRAG
#pragma unroll
        for (int G4 = 0; G4 < 4; ++G4) {
          double sv = rs4[G4];
          sv += __shfl_xor_sync(0xffffffffu, sv, 1);
          sv += __shfl_xor_sync(0xffffffffu, sv, 2);
          const int rr = tm * I8_BM + quarter * 32 + 8 * G4 + g;
          if (q4 == 0 && rr < p.M) p.mom[(int64_t)slab * p.sMomTile + (int64_t)rr * nq] = sv;
        }
#endif
      } else {
        const int rown = tm * I8_BM + quarter * 32 + lane;
        if (rown < p.M) p.mom[(int64_t)slab * p.sMomTile + (int64_t)rown * nq] = rsA;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg[3 * I8_DBG_ITEMS * 4 + 2 * blockIdx.x + 1] = (long long)gt;
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(I8_TMEM_COLS));
}

// e[0] = exponent of the k(X,Z) digits (k <= sf2), e[1] = exponent of the A digits (|A[m,n]| <= sqrt(k_nn) = sqrt(sf2), one bit of margin:
// the computed |A| may exceed the bound by rounding) -- computed on the device so that no entry point reads theta back to the host
__global__ void k_i8_exponents(const double* __restrict__ theta, int d, int* __restrict__ e) {
  const double sf2 = theta[d];
  e[0] = i8_exp_for(sf2);
  e[1] = i8_exp_for(sqrt(sf2)) + 1;
}

// row-scaled signed-digit slicing of X[R x K] (leading dimension ld): planes Xq[i][row][k] (leading dimension ldq, plane stride
// plane), per-row exponent ex[row] = i8_exp_for(row maximum).  One warp per row.  Columns [K, Kpad) are written as zero.
__global__ void __launch_bounds__(256) k_slice_rows(const double* __restrict__ X, int R, int K, int64_t ld, int8_t* __restrict__ Xq,
                                                    int64_t ldq, int64_t plane, int Kpad, int* __restrict__ ex) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  double mx = 0.0;
  for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(X[(int64_t)row * ld + k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int e = i8_exp_for(mx);
  if (lane == 0) ex[row] = e;
  for (int k = lane; k < Kpad; k += 32) {
    int8_t dg[I8_NS];
    i8_digits(k < K ? ldexp(X[(int64_t)row * ld + k], -e) : 0.0, dg);
#pragma unroll
    for (int i = 0; i < I8_NS; ++i) Xq[(int64_t)i * plane + (int64_t)row * ldq + k] = dg[i];
  }
}

// Tensor-pipe peak probe (roofline denominator of the sliced-integer GEMMs): the MMA issue loop of k_gemm_i8 with the operand tiles
// resident in shared memory -- no TMA, no epilogue, accumulators never read.  One CTA per SM, one issuing thread.
//   MODE 0: 16 x (128 x 256 x 32) per iteration: the shape-independent tcgen05 kind::i8 rate
//   MODE 1: the production mix, two 32-byte k-steps per iteration, 10 MMAs each (N = 224+224, 192+192, 160+160, 256, 192, 128, 64):
//           what the mainloop could reach if its operands were always ready
// `depth` iterations may be in flight (depth = 2 mimics the 2-stage ring: iteration it waits for the commit of iteration it - 2).
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_i8_probe(int iters, int depth, unsigned seed) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + I8_STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  for (int i = threadIdx.x; i < I8_STAGE_BYTES / 4; i += blockDim.x) {   // pseudo-random digit bytes (the data sets the power draw)
    unsigned x = (unsigned)i * 2654435761u + seed + blockIdx.x * 40503u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    reinterpret_cast<uint32_t*>(base)[i] = x;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(I8_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 32) {
    const uint32_t sa = smem_u32(base), sb = sa + I8_NS * I8_A_BYTES;
    for (int it = 0; it < iters; ++it) {
      const int s = it % depth;
      if (it >= depth) i8_mbar_wait(&bars[s], ((it / depth) - 1) & 1);
      if (MODE == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          i8_umma(tmem_base + (uint32_t)((j & 1) * 256), i8_desc_sw64(sa + (j % I8_NS) * I8_A_BYTES + ((j >> 3) & 1) * 32),
                  i8_desc_sw64(sb + ((j >> 1) & 1) * 32), idesc, (it > 0 || j > 1) ? 1u : 0u);
      } else if (MODE == 2) {
        // the mix of a 128 x 32 output tile (double-buffered TMEM: 2 x 7 x 32 columns): 7 MMAs per k-step, N = 224, 192, ..., 32;
        // two tiles' worth per iteration so that the work per iteration equals MODE 1
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
          for (int kk = 0; kk < I8_BKB / 32; ++kk)
#pragma unroll
            for (int i = 0; i < I8_NS; ++i) {
              const int ncols = 32 * (I8_NS - i);
              const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
              i8_umma(tmem_base + (uint32_t)(t2 * 224 + i * 32), i8_desc_sw64(sa + i * I8_A_BYTES + kk * 32),
                      i8_desc_sw64(sb + t2 * 224 * I8_BKB + kk * 32), idesc, (it > 0 || kk > 0 || i > 0) ? 1u : 0u);
            }
      } else {
#pragma unroll
        for (int kk = 0; kk < I8_BKB / 32; ++kk) {
#pragma unroll
          for (int i = 0; i < I8_NS; ++i) {
            const uint64_t ad = i8_desc_sw64(sa + i * I8_A_BYTES + kk * 32);
            const int ncols = I8_BN * (I8_NS - i);
            const int nsplit = ncols > 256 ? 2 : 1, nn = ncols / nsplit;
#pragma unroll
            for (int hs = 0; hs < nsplit; ++hs) {
              const int off = hs * nn;
              const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
              i8_umma(tmem_base + (uint32_t)(i * I8_BN + off), ad, i8_desc_sw64(sb + off * I8_BKB + kk * 32), idesc,
                      (it > 0 || kk > 0 || i > 0) ? 1u : 0u);
            }
          }
        }
      }
      i8_umma_commit(&bars[s]);
    }
    for (int it = max(0, iters - depth); it < iters; ++it) i8_mbar_wait(&bars[it % depth], (it / depth) & 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(I8_TMEM_COLS));
}
constexpr int I8_PROBE_SMEM = I8_STAGE_BYTES + 1024 + 256;

}  // namespace ggp
