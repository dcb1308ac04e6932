// Common device/host helpers for the WAP sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <utility>

namespace wap {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);          // abi.cu (thread-local message)
#define WAP_CHECK(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::wap::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                     \
    }                                                                                \
  } while (0)
extern unsigned long long g_launches;          // abi.cu: kernels launched by this library (cumulative)
#define WAP_LAUNCH_CHECK()                                                           \
  do { ++::wap::g_launches; WAP_CHECK(cudaGetLastError()); } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A training step is ~1700 short dependent kernels; with the programmatic-stream-serialization attribute the next kernel's
// grid is scheduled as soon as every CTA of the current one has executed pdl_sync(), so its launch latency and CTA
// start-up overlap the current kernel's execution.  Every kernel of this library calls pdl_sync() first:
// griddepcontrol.wait returns when the preceding grids have completed and their writes are visible (immediately when the
// kernel was launched without the attribute), and launch_dependents then lets the successor start queueing.
bool pdl_enabled();                            // abi.cu: wap_set_pdl() of the current context
// (Measured and rejected: triggering before the wait - so that barrier/TMEM set-up or weight prefetch overlap the predecessor -
// gained nothing in the skinny GEMMs, cost 0.2 ms/step in the tcgen05 kernels, and both together ended in launch failures.)
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// optional per-launch-class device timing (cudaEvents on the launching stream); see wap_profile_*()
enum ProfClass {
  kProfConv1Fwd = 0, kProfConv3Fwd, kProfTransFwd, kProfConv3Wgrad, kProfConv3Dgrad, kProfConv1Wgrad, kProfConv1Dgrad,
  kProfTransBwd, kProfCovGemm, kProfOtherGemm, kProfNumClasses
};
void prof_begin(cudaStream_t st, int cls, double flops);
void prof_set_bytes(double bytes);             // algorithmic HBM bytes of the NEXT prof_begin (default: 4(MK+KN+MN) unknown -> 0)
void prof_end(cudaStream_t st, int cls);
#define WAP_REQUIRE(cond, ...)                                                       \
  do {                                                                               \
    if (!(cond)) { ::wap::set_error(__VA_ARGS__); return -1; }                       \
  } while (0)
#define WAP_TRY(call)                                                                \
  do { int _r = (call); if (_r != 0) return _r; } while (0)

// ---------------------------------------------------------------- model constants
// (reference literals: model-single-GPU.py:574, :584-588, :50-55, :81-82, :598)
constexpr int kBlocks = 3, kLevels = 16, kGrowth = 24, kStemC = 48, kBottle = 96;
constexpr int kD = 684;        // annotation channels
constexpr int kN = 256;        // decoder hidden
constexpr int kMw = 256;       // word embedding
constexpr int kNA = 512;       // attention dim
constexpr int kVocab = 111;
constexpr int kCovK = 11, kCovTaps = 121;
constexpr int kNumConv = 99;
constexpr int kMaxBnC = 304;   // >= max conv Cout (300), padded
constexpr int kDy2Ld = 32;     // row stride of the 3x3-conv output gradient (24 channels zero-padded to one 128 B TMA row)
constexpr float kBnEps = 1e-4f, kBnMomentum = 0.9f;
constexpr float kDropRate = 0.2f;
constexpr uint32_t kDropThresh = 13421772u;   // floor(0.8 * 2^24): keep iff (hash>>8) < thresh

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t cdivz(size_t a, size_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- counter-based dropout RNG
// keep-mask for element `idx` of dropout site `site` (conv index k, or kNumConv for the logit
// dropout) under `seed`.  murmur3 finaliser; reproduced bit-exactly in numpy by the tests.
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t site, uint32_t idx) {
  uint32_t h = idx ^ (seed + site * 0x632BE5ABu);
  h *= 0x9E3779B1u; h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t seed, uint32_t site, uint32_t idx) {
  return (drop_hash(seed, site, idx) >> 8) < kDropThresh;
}

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 make4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in all threads; `red` = shared float[33]
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) { r = warp_sum(r); if (lane == 0) red[32] = r; }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) { r = warp_max(r); if (lane == 0) red[32] = r; }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace wap
