// Probe: dependent-launch cost of a kernel as a function of its resource footprint (dynamic shared memory, TMEM allocation, grid size), inside a
// replayed CUDA graph with and without programmatic dependent launch.  Explains the fixed cost of the small tcgen05 GEMM launches of the scans.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } } while (0)
__device__ __forceinline__ void pdl_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <int TMEM> __global__ void __launch_bounds__(192) k(float* p, int spin) {
  pdl_sync();
  extern __shared__ uint8_t s[];
  __shared__ uint32_t slot;
  if (TMEM) {
    if (threadIdx.x < 32) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(TMEM) : "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (p != nullptr && threadIdx.x == 0) p[blockIdx.x] = (float)spin;
  if (TMEM) {
    __syncthreads();
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(TMEM) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
}
template <int TMEM> float run(int grid, size_t smem, bool pdl, int spin, float* buf) {
  CK(cudaFuncSetAttribute(k<TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  const int reps = 100;
  auto launch = [&]() {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k<TMEM>, buf, spin));
  };
  cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < reps; ++i) launch();
  cudaError_t ec = cudaStreamEndCapture(st, &g);
  if (ec != cudaSuccess || g == nullptr) { printf("capture failed: %s\n", cudaGetErrorString(ec)); cudaGetLastError(); cudaStreamDestroy(st); return -1.f; }
  ec = cudaGraphInstantiate(&ge, g, 0);
  if (ec != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(ec)); return -1.f; }
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
  return ms * 1e3f / reps;
}
int main() {
  float* buf; CK(cudaMalloc(&buf, 4096 * 4));
  const int spin = 2000;   // ~1 us of "work"
  printf("us per dependent launch (kernel body spins ~1 us); columns: no-PDL / PDL\n");
  for (int grid : {1, 16, 128, 148, 296}) for (int kb : {0, 48, 100, 200, 220}) {
    printf("grid %3d smem %3d KB  tmem   0: %6.2f %6.2f   tmem 512: %6.2f %6.2f   tmem 128: %6.2f %6.2f\n", grid, kb,
           run<0>(grid, (size_t)kb * 1024, false, spin, buf), run<0>(grid, (size_t)kb * 1024, true, spin, buf),
           run<512>(grid, (size_t)kb * 1024, false, spin, buf), run<512>(grid, (size_t)kb * 1024, true, spin, buf),
           run<128>(grid, (size_t)kb * 1024, false, spin, buf), run<128>(grid, (size_t)kb * 1024, true, spin, buf));
  }
  return 0;
}
