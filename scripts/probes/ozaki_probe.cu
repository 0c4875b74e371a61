// Feasibility probe (developer tool, standalone): FP64-class "NT" GEMM on the 5th-gen tensor cores by exact integer slicing
// (Ozaki scheme) --  C[M x N] (f64) = A[M x K] * B[N x K]^T  with
//   * operands cut into NS signed 7-bit digits of a row-scaled fixed-point value (k_slice_rows),
//   * tcgen05.mma kind::i8 (SASS UTCIMMA), operands staged by TMA (cp.async.bulk.tensor, 64-byte swizzle) in a 2-stage mbarrier ring,
//   * int32 accumulators in TMEM, ONE accumulator per significance level l = i + j (all digit pairs of a level are summed exactly
//     in the same accumulator: |d|^2 * (l+1) * K <= 4096 * 8 * K < 2^31 for K <= 2^16),
//   * epilogue: tcgen05.ld the NS level accumulators, combine them in FP64 from the least significant level up, apply the row scales.
// One CTA = one 128 x BN output tile; NS * BN = 512 TMEM columns.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ozaki_probe ozaki_probe.cu
// Run on the B200 box: ./ozaki_probe [M N K]   (prints accuracy vs a long-double host reference on sampled entries, and TFLOP/s)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifndef NS
#define NS 8            // digits per operand (7 bits each)
#endif
#ifndef BN
#define BN 64           // tile columns; NS * BN <= 512 TMEM columns
#endif
constexpr int BM = 128;
#ifndef BKB
#define BKB 64                    // bytes (= int8 elements) of K per stage row: one swizzle row (64 -> SWIZZLE_64B, 32 -> SWIZZLE_32B)
#endif
#ifndef STAGES
#define STAGES 2
#endif
constexpr int A_SLICE_BYTES = BM * BKB, B_SLICE_BYTES = BN * BKB;
constexpr int STAGE_BYTES = NS * (A_SLICE_BYTES + B_SLICE_BYTES);
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 512;
static_assert(NS * BN <= 512, "level accumulators must fit TMEM");
constexpr int THREADS = 192;      // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue (lane quarters 2,3,0,1)

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAITL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONEL;\nbra WAITL;\nDONEL:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major operand tile, 64-byte swizzle: rows of 64 bytes, 8-row groups 512 bytes apart (SBO), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), canonical value 1
  d |= (uint64_t)((8 * BKB) >> 4) << 32;     // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // version
  d |= (uint64_t)(BKB == 64 ? 4 : 6) << 61;  // SWIZZLE_64B / SWIZZLE_32B
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}

// C[m][n] = 2^(ea[m] + eb[n]) * sum_l 2^(-7 (l + 2)) * acc_l[m][n]
__global__ void __launch_bounds__(THREADS, 1)
k_ozaki_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const int* __restrict__ ea,
             const int* __restrict__ eb, double* __restrict__ C, int M, int N, int K, int64_t ldc, int mode) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = N / BN;
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  const int nkb = K / BKB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < (mode == 1 ? (nkb < STAGES ? nkb : STAGES) : nkb); ++kb) {   // mode 1 (diagnostic): fill the ring once, MMAs re-read it
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        unsigned char* st = base + s * STAGE_BYTES;
        // one box per operand covers all NS digit planes (k x rows x planes): 2 TMA instructions per stage
        tma_load_3d(st, &tmA, kb * BKB, tm * BM, 0, &full[s]);
        tma_load_3d(st + NS * A_SLICE_BYTES, &tmB, kb * BKB, tn * BN, 0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      if (mode != 1 || kb < STAGES) mbar_wait(&full[s], (kb / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      if (lane == 0) {
        const uint32_t sa = smem_u32(base + s * STAGE_BYTES), sb = sa + NS * A_SLICE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BKB / 32; ++kk) {
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            // digit i of A against digits 0..NS-1-i of B in ONE wide MMA: the B digit tiles are consecutive K-major tiles, i.e. one
            // tall tile of BN (NS - i) rows, and their products belong to the consecutive levels i..NS-1 = consecutive accumulator
            // columns.  (A is read from shared memory once per <= 256 columns instead of once per digit pair: SS-mode MMAs with
            // N = 64 are shared-memory-read bound.)
            const uint64_t ad = make_desc_sw64(sa + i * A_SLICE_BYTES + kk * 32);
            const int ncols = BN * (NS - i);
#pragma unroll
            for (int off = 0; off < ncols; off += 256) {
              const int nn = (ncols - off) < 256 ? (ncols - off) : 256;
              const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
              const uint64_t bd = make_desc_sw64(sb + (off / BN) * B_SLICE_BYTES + kk * 32);
              umma_i8(tmem_base + (uint32_t)(i * BN + off), ad, bd, idesc, (kb > 0 || kk > 0 || i > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty[s]);                       // frees the stage when the MMAs that read it have completed
        if (kb == nkb - 1) umma_commit(tmem_full);    // accumulators complete
      }
      __syncwarp();
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32 (w % 4), 32 (w % 4) + 32)
    const int quarter = warp & 3;
    const int row = tm * BM + quarter * 32 + lane;
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const int ea_r = ea[row];
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      double sum[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) sum[c] = 0.0;
#pragma unroll
      for (int l = NS - 1; l >= 0; --l) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(l * BN + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const double w = exp2(-7.0 * (l + 2));
#pragma unroll
        for (int c = 0; c < 16; ++c) sum[c] = fma((double)(int)v[c], w, sum[c]);
      }
      double* dst = C + (int64_t)row * ldc + tn * BN + c0;
#pragma unroll
      for (int c = 0; c < 16; ++c) dst[c] = ldexp(sum[c], ea_r + eb[tn * BN + c0 + c]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// row-scaled signed-digit slicing: x = 2^e * sum_i d_i 2^(-7 (i + 1)), |d_i| <= 64; one warp per row
__global__ void k_slice_rows(const double* __restrict__ X, int R, int K, int64_t ld, int8_t* __restrict__ Xq, int* __restrict__ ex) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  double mx = 0.0;
  for (int k = lane; k < K; k += 32) mx = fmax(mx, fabs(X[(int64_t)row * ld + k]));
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int e = mx > 0.0 ? ilogb(mx) + 2 : 0;          // |x| / 2^e < 1/2
  if (lane == 0) ex[row] = e;
  for (int k = lane; k < K; k += 32) {
    double t = ldexp(X[(int64_t)row * ld + k], -e);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      t *= 128.0;
      const double d = rint(t);
      t -= d;
      Xq[((int64_t)i * R + row) * K + k] = (int8_t)(int)d;
    }
  }
}

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* tm, int8_t* ptr, int R, int K, int box_rows) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)R, (cuuint64_t)NS};
  const cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)R * K};
  const cuuint32_t box[3] = {(cuuint32_t)BKB, (cuuint32_t)box_rows, (cuuint32_t)NS};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  return ((tmap_encode_fn)fp)(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              BKB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char** argv) {
  int M = argc > 3 ? atoi(argv[1]) : 1024, N = argc > 3 ? atoi(argv[2]) : 16384, K = argc > 3 ? atoi(argv[3]) : 1024;
  const int mode = argc > 4 ? atoi(argv[4]) : 0;
  if (mode) printf("DIAGNOSTIC mode %d: results are NOT a GEMM\n", mode);
  if (M % BM || N % BN || K % BKB || K > 65536) { printf("shape must be a multiple of the tile (128, %d, %d), K <= 65536\n", BN, BKB); return 1; }
  printf("ozaki probe: M=%d N=%d K=%d  NS=%d digits (%d bits), tile 128x%d, %d products per k-step\n", M, N, K, NS, 7 * NS, BN, NS * (NS + 1) / 2);
  std::vector<double> hA((size_t)M * K), hB((size_t)N * K);
  srand(1);
  auto rnd = []() { return (rand() + 0.5) / (RAND_MAX + 1.0); };
  for (int m = 0; m < M; ++m) {
    const double sc = exp2(20.0 * rnd() - 10.0);       // rows of very different magnitude (like L^{-1} / P)
    for (int k = 0; k < K; ++k) hA[(size_t)m * K + k] = sc * (2.0 * rnd() - 1.0) * exp2(-8.0 * rnd());
  }
  for (size_t i = 0; i < hB.size(); ++i) hB[i] = exp(-6.0 * rnd());   // kernel-tile like values in (0, 1]
  double *dA, *dB, *dC;
  int8_t *qA, *qB;
  int *eA, *eB;
  CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8)); CK(cudaMalloc(&dC, (size_t)M * N * 8));
  CK(cudaMalloc(&qA, (size_t)NS * M * K)); CK(cudaMalloc(&qB, (size_t)NS * N * K));
  CK(cudaMalloc(&eA, M * 4)); CK(cudaMalloc(&eB, N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 8));
  CUtensorMap tmA, tmB;
  if (!make_map(&tmA, qA, M, K, BM) || !make_map(&tmB, qB, N, K, BN)) { printf("tensor map encode failed\n"); return 1; }
  CK(cudaFuncSetAttribute(k_ozaki_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms_slice = 0, ms = 0;
  CK(cudaEventRecord(e0));
  k_slice_rows<<<(M + 7) / 8, 256>>>(dA, M, K, K, qA, eA);
  k_slice_rows<<<(N + 7) / 8, 256>>>(dB, N, K, K, qB, eB);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&ms_slice, e0, e1));
  const int grid = (M / BM) * (N / BN);
  k_ozaki_gemm<<<grid, THREADS, SMEM_BYTES>>>(tmA, tmB, eA, eB, dC, M, N, K, N, mode);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  const int reps = 5;
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) k_ozaki_gemm<<<grid, THREADS, SMEM_BYTES>>>(tmA, tmB, eA, eB, dC, M, N, K, N, mode);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  std::vector<double> hC((size_t)M * N);
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 8, cudaMemcpyDeviceToHost));
  // accuracy on sampled entries against a long-double host reference; error relative to sum |a||b| (the FP64 GEMM error scale)
  double worst = 0.0, worst_rel = 0.0;
  int nbad = 0;
  for (int t = 0; t < 4000; ++t) {
    const int m = rand() % M, n = rand() % N;
    long double ref = 0.0L, mag = 0.0L;
    for (int k = 0; k < K; ++k) {
      const long double p = (long double)hA[(size_t)m * K + k] * (long double)hB[(size_t)n * K + k];
      ref += p; mag += fabsl(p);
    }
    const double got = hC[(size_t)m * N + n];
    const double err = (double)(fabsl((long double)got - ref) / mag);
    if (!(err < 1e-6)) ++nbad;
    if (err > worst || err != err) worst = err;
    const double rel = (double)(fabsl((long double)got - ref) / fabsl(ref));
    if (rel > worst_rel) worst_rel = rel;
  }
  const double flops = 2.0 * M * N * K;
  printf("slicing: %.3f ms for both operands\n", ms_slice);
  printf("gemm: %.3f ms  -> %.2f TFLOP/s FP64-equivalent, %.1f TOP/s int8 (%d products)\n", ms, flops / ms * 1e-9,
         flops / ms * 1e-9 * (NS * (NS + 1) / 2), NS * (NS + 1) / 2);
  printf("accuracy on 4000 sampled entries: max |err| / sum|a||b| = %.3e   max |err| / |ref| = %.3e   entries with err >= 1e-6: %d\n", worst,
         worst_rel, nbad);
  printf("(an FP64 FMA chain of this K has max |err| / sum|a||b| ~ %.1e)\n", sqrt((double)K) * 1.1e-16);
  return nbad ? 2 : 0;
}
