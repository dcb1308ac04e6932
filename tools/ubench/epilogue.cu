// Probe: cost of the tcgen05 GEMM epilogue per 32-column chunk (128 rows x 32 fp32 columns, 4 warps, TMEM lane = row) for different ways of
// getting the accumulator to global memory.  clock64 of thread 0 around the loop over `chunks` chunks.
//   mode 0: tcgen05.ld only          1: + row-per-thread float4 stores (each lane its own row: 32 lines per instruction)
//   mode 2: + transposed scalar stores (coalesced across lanes)     3: + staged through padded shared memory, coalesced float4 row stores
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } } while (0)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <int MODE> __global__ void __launch_bounds__(128) k(float* out, int ld, int chunks, long long* t, int x3) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) float stage[4][32][36];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  const int r = blockIdx.x * 128 + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int c = 0; c < chunks * 32; c += 32) {
    float v[32];
    tmem_ld32(base + (uint32_t)(c & 127), v);
    if (x3) {
      float v2[32];
      tmem_ld32(base + 128u + (uint32_t)(c & 127), v2);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += v2[i];
    }
    if (MODE == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) s += v[i];
      if (s == 12345.678f) out[r] = s;
    }
    if (MODE == 1) {
      float* dst = out + (size_t)r * ld + c;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 32; ++j) out[(size_t)(c + j) * ld + r] = v[j];
    }
    if (MODE == 3) {     // warp-private staging: row = lane, 36-float pitch (conflict-free 16-byte accesses), then 8 lanes per 128-byte row
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(&stage[warp][lane][j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), q4 = (lane & 7) * 4;
        const float4 w = *reinterpret_cast<const float4*>(&stage[warp][row][q4]);
        *reinterpret_cast<float4*>(out + (size_t)(blockIdx.x * 128 + warp * 32 + row) * ld + c + q4) = w;
      }
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { t[0] = t1 - t0; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
template <int MODE> void run(const char* name, int grid, int chunks, int x3) {
  float* out; long long* t; const int ld = MODE == 2 ? grid * 128 : chunks * 32;
  CK(cudaMalloc(&out, (size_t)grid * 128 * chunks * 32 * 4 + 4096)); CK(cudaMalloc(&t, 64));
  for (int i = 0; i < 3; ++i) k<MODE><<<grid, 128>>>(out, ld, chunks, t, x3);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<grid, 128>>>(out, ld, chunks, t, x3); cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-40s grid %3d chunks %d x3 %d: %7.0f cycles per chunk (%.2f us @1.965 GHz), kernel %.2f us\n", name, grid, chunks, x3, (double)h / chunks, (double)h / chunks / 1965.0, ms * 1e3);
  cudaFree(out); cudaFree(t);
}
int main() {
  for (int grid : {1, 148}) for (int x3 : {0, 1}) {
    run<0>("tcgen05.ld only", grid, 4, x3);
    run<1>("row-per-thread float4 stores", grid, 4, x3);
    run<2>("transposed scalar stores", grid, 4, x3);
    run<3>("smem-staged coalesced float4 stores", grid, 4, x3);
  }
  return 0;
}
