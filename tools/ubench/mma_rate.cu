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
// the exact issue pattern of the 3xTF32 main loop: per k-step  acc0 += a_hi*b_hi ; acc2 += a_lo*b_hi ; acc2 += a_hi*b_lo
// (A from TMEM ring, B hi/lo tiles of 4 smem stages), 12 MMAs per k-block
template <int N, int MODE, int NOISE = 0>     // MODE 0: as in the kernel; 1: single accumulator; 2: same B tile for all three
__global__ void __launch_bounds__(384) kx3(long long* out, int reps) {
  __shared__ volatile int stop;
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  constexpr int kStage = 2 * N * 128;
  for (int i = threadIdx.x; i < 4 * kStage / 4; i += 128) ((float*)sm)[i] = 1.0f;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (threadIdx.x >= 128) {       // noise warps (8): emulate the splitter's side traffic while thread 0 issues MMAs
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v[32];
    for (int i = 0; i < 32; ++i) v[i] = (float)(i + lane);
    int it = 0;
    while (!stop) {
      if (NOISE == 1) {
        const uint32_t ta = tm + ((uint32_t)((w & 3) * 32) << 16) + 256 + (uint32_t)((it & 3) * 64);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(ta),
          "r"(__float_as_uint(v[0])),"r"(__float_as_uint(v[1])),"r"(__float_as_uint(v[2])),"r"(__float_as_uint(v[3])),"r"(__float_as_uint(v[4])),"r"(__float_as_uint(v[5])),"r"(__float_as_uint(v[6])),"r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])),"r"(__float_as_uint(v[9])),"r"(__float_as_uint(v[10])),"r"(__float_as_uint(v[11])),"r"(__float_as_uint(v[12])),"r"(__float_as_uint(v[13])),"r"(__float_as_uint(v[14])),"r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])),"r"(__float_as_uint(v[17])),"r"(__float_as_uint(v[18])),"r"(__float_as_uint(v[19])),"r"(__float_as_uint(v[20])),"r"(__float_as_uint(v[21])),"r"(__float_as_uint(v[22])),"r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])),"r"(__float_as_uint(v[25])),"r"(__float_as_uint(v[26])),"r"(__float_as_uint(v[27])),"r"(__float_as_uint(v[28])),"r"(__float_as_uint(v[29])),"r"(__float_as_uint(v[30])),"r"(__float_as_uint(v[31])) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int d = 0; d < 40; ++d) v[d & 31] += 1.0f;      // pace: roughly two stores of 32 columns per ~600 cycles and warp
      } else if (NOISE == 2) {
        const float4* row = reinterpret_cast<const float4*>(sm + ((threadIdx.x & 127) * 128));
        for (int c = 0; c < 8; ++c) { const float4 q = row[c ^ (lane & 7)]; v[4*c] += q.x; v[4*c+1] += q.y; v[4*c+2] += q.z; v[4*c+3] += q.w; }
      } else if (NOISE == 3) {
        uint32_t ok; asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)));
        v[0] += ok;
      }
      ++it;
    }
    if (v[3] == 123.456f) out[0] = 0;
  }
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t acc2 = MODE == 1 ? 0u : 128u;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int s = r & 3;
      const uint32_t sb = smem_u32(sm + s * kStage);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t bd = desc(sb + kk * 32), bld = MODE == 2 ? bd : desc(sb + N * 128 + kk * 32);
        const uint32_t ah = tm + 256 + s * 64 + kk * 8, al = ah + 32;
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(tm), "r"(ah), "l"(bd), "r"(idesc), "r"(1u));
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(tm + acc2), "r"(al), "l"(bd), "r"(idesc), "r"(1u));
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(tm + acc2), "r"(ah), "l"(bld), "r"(idesc), "r"(1u));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)));
    out[blockIdx.x] = clock64() - t0;
    stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
template <int N, int MODE, int NOISE = 0> void runx3(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8); long long h[148];
  const int smem = 4 * 2 * N * 128 + 2048, reps = 2000;
  cudaFuncSetAttribute(kx3<N, MODE, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) { kx3<N, MODE, NOISE><<<148, NOISE ? 384 : 128, smem>>>(d, reps); cudaDeviceSynchronize(); }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  printf("x3 pattern %-28s N=%3d: %.0f cycles per k-block (12 MMAs)   err=%s\n", name, N, s / 148 / reps, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
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
  runx3<96, 0>("(as in the kernel)"); runx3<96, 1>("(single accumulator)"); runx3<96, 2>("(one B tile)");
  runx3<32, 0>("(as in the kernel)"); runx3<128, 0>("(as in the kernel)");
  runx3<96, 0, 1>("+ 8 warps tcgen05.st"); runx3<96, 0, 2>("+ 8 warps LDS/ALU"); runx3<96, 0, 3>("+ 8 warps mbarrier polling");
  return 0;
}
