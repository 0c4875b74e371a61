// Developer micro-probe: what limits DMMA.8x8x4 issue on sm_100a?  (standalone; nvcc -arch=sm_100a -o dmma_probe dmma_probe.cu)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* d, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(d)), "l"(g)); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity));
}
// MODE 4: MODE 2 + per-thread LDGSTS traffic (64 KB per CTA per 256-DMMA step, 3-stage ring, wait_group 1)
// MODE 5: MODE 1 + bulk-copy (UBLKCP) traffic issued by one thread, completion on an mbarrier that every warp polls (no CTA barrier)
// MODE 0: register resident.  MODE 1: 12 LDS.64 per 32 DMMA (fragments from smem).  MODE 2: MODE 1 + __syncthreads every 8 k4-steps
// MODE 3: MODE 1 with fragment double buffering (loads for step k+1 issued before the DMMAs of step k)
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_probe(double* sink, int iters, int lds, const double* gsrc) {
  extern __shared__ double sm[];
  __shared__ unsigned long long bars[3];
  double* ring = sm + 256 * lds;   // 3 stages x 64 KB (MODE 4/5)
  const double* gs = gsrc + (size_t)blockIdx.x * 2 * 8192;
  if (MODE == 5 && threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
  }
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, q = lane & 3;
  for (int i = tid; i < 256 * lds; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
  __syncthreads();
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double a[8], b[4], a2[8], b2[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (lane + i);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 1.0 - 1e-9 * (lane + j);
  const double* pa = sm + g * lds + q;
  const double* pb = sm + (128 + g) * lds + q;
  if (MODE == 3) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pa[i * 8 * lds];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = pb[j * 8 * lds];
  }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 3) {
#pragma unroll
      for (int kk = 0; kk < 8; kk += 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a2[i] = pa[i * 8 * lds + (kk + 1) * 4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b2[j] = pb[j * 8 * lds + (kk + 1) * 4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = pa[i * 8 * lds + ((kk + 2) & 7) * 4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = pb[j * 8 * lds + ((kk + 2) & 7) * 4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a2[i], b2[j]);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (MODE == 1 || MODE == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = pa[i * 8 * lds + kk * 4];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = pb[j * 8 * lds + kk * 4];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    if (MODE == 2) __syncthreads();
    if (MODE == 4) {
      double* dst = ring + (it % 2) * 8192;
      const double* src = gs + (it % 2) * 8192;
      for (int c = threadIdx.x; c < 4096; c += WARPS * 32) cp_async16(dst + c * 2, src + c * 2);
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 1;\n" ::);
      __syncthreads();
    }
    if (MODE == 5) {
      if (threadIdx.x == 0) {
        // previous use of this stage was consumed 3 iterations ago (all warps are within one iteration of each other here)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bars[it % 2])), "r"(65536));
        for (int c = 0; c < 4; ++c)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                           smem_u32(ring + (it % 2) * 8192 + c * 2048)), "l"(gs + (it % 2) * 8192 + c * 2048), "r"(16384), "r"(smem_u32(&bars[it % 2])) : "memory");
      }
      if (it >= 1) mbar_wait(&bars[(it - 1) % 2], ((it - 1) / 2) & 1);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 1.2345) sink[tid] = s;
}
template <int MODE, int WARPS>
void run(const char* name, int sms, double* sink) {
  const int warps = WARPS;
  const int iters = 2000, lds = 36;
  const size_t smem = 256 * lds * 8 + (MODE >= 4 ? 2 * 65536 : 0);
  static double* gsrc = nullptr;
  if (!gsrc) { cudaMalloc(&gsrc, (size_t)sms * 2 * 65536); cudaMemset(gsrc, 0, (size_t)sms * 2 * 65536); }
  cudaFuncSetAttribute(k_probe<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    k_probe<MODE, WARPS><<<sms, warps * 32, smem>>>(sink, iters, lds, gsrc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  const double flops = (double)sms * warps * iters * 8.0 * 32.0 * 512.0;
  printf("%-28s warps/SM %2d : %7.3f ms  %6.2f TF/s  %s\n", name, warps, best, flops / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* sink; cudaMalloc(&sink, 1 << 20);
  const int sms = p.multiProcessorCount;
  run<0, 4>("registers", sms, sink); run<0, 8>("registers", sms, sink);
  run<1, 4>("12 LDS.64 / 32 DMMA", sms, sink); run<1, 8>("12 LDS.64 / 32 DMMA", sms, sink);
  run<2, 4>("  + barrier / 256 DMMA", sms, sink); run<2, 8>("  + barrier / 256 DMMA", sms, sink);
  run<4, 4>("LDS+bar+LDGSTS 64KB/step", sms, sink); run<4, 8>("LDS+bar+LDGSTS 64KB/step", sms, sink);
  run<5, 4>("LDS+UBLKCP 64KB/step mbar", sms, sink); run<5, 8>("LDS+UBLKCP 64KB/step mbar", sms, sink);
  return 0;
}
