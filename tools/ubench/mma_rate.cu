// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (M=128, K=8) for several N, A from shared memory (SS) or from
// tensor memory (TS), on every SM at once.  Operands are never refreshed: this measures the tensor pipe only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFF) >> 4); d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
template <int N, bool TS>
__global__ void __launch_bounds__(128) k(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) ((float*)sm)[i] = 1.0f;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = desc(smem_u32(sm)), bd = desc(smem_u32(sm + 16384));
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (TS) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(tm), "r"(tm + 256 + kk * 8), "l"(bd + kk * 2), "r"(idesc), "r"(1u));
        else asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tm), "l"(ad + kk * 2), "l"(bd + kk * 2), "r"(idesc), "r"(1u));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)));
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
template <int N, bool TS> void run(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8); long long h[148];
  const int smem = 16384 + N * 128 + 2048, reps = 2000;
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) { k<N, TS><<<148, 128, smem>>>(d, reps); cudaDeviceSynchronize(); }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  const double cyc = s / 148 / (reps * 4.0);
  printf("%s N=%3d: %.1f cycles per MMA (128xNx8)  -> %.0f FLOP/clk/SM   err=%s\n", name, N, cyc, 2.0 * 128 * N * 8 / cyc, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<32, false>("SS"); run<64, false>("SS"); run<96, false>("SS"); run<128, false>("SS"); run<256, false>("SS");
  run<32, true>("TS"); run<96, true>("TS"); run<128, true>("TS"); run<256, true>("TS");
  return 0;
}
