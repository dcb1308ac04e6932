// Generic fp32 (FFMA) tiled GEMM / implicit-GEMM convolution core.
//
//   C(m,n) = sum_k A(m,k) * B(k,n)        m<M, n<N, k<K
//
// A and B are *functors* ("loaders") that return four consecutive elements as a float4 either
// along k (kVecK = true) or along the other index (m for A, n for B).  That one abstraction
// covers: plain row-major GEMM, strided slices of the NHWC dense-block buffer (1x1 convs read
// channels [0,C) of a [M,Ctot] buffer in place), 3x3 SAME im2col gathers with zero padding,
// transposed operands for dgrad/wgrad, and the coverage-patch gather of the attention backward.
// Epilogues: plain/accumulating store with optional bias, per-column sum / sum-of-squares
// (batch-norm statistics, fp64 global accumulation), or atomic add (split-K).
//
// This is the fp32-faithful path (parity mode, <=1e-3 vs the fp64 oracle).  The bf16 tcgen05
// path for the dense-block convolutions lives in conv_tc.cu.
#pragma once
#include "common.cuh"

namespace wap {

constexpr int BK = 16;

template <int BM_, int BN_, int RM_, int RN_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, RM = RM_, RN = RN_;
  static constexpr int NTX = BN / (4 * RN), NTY = BM / (4 * RM), NT = NTX * NTY;
  static_assert(BN % (4 * RN) == 0 && BM % (4 * RM) == 0, "tile/fragment mismatch");
};
using Cfg128x96 = TileCfg<128, 96, 1, 3>;    // 256 thr, 4x12 per thread : 1x1 convs (N=96)
using Cfg256x24 = TileCfg<256, 24, 2, 1>;    // 192 thr, 8x4  per thread : 3x3 convs (N=24)
using Cfg128x128 = TileCfg<128, 128, 2, 2>;  // 256 thr, 8x8  per thread : general
using Cfg64x64 = TileCfg<64, 64, 1, 1>;      // 256 thr, 4x4  per thread : small problems

// ----------------------------------------------------------------------------- A loaders
struct ARowMajor {            // A(m,k) = p[m*ld + k]
  static constexpr bool kVecK = true;
  const float* p; int ld, M, K;
  __device__ __forceinline__ float4 load(int m, int k) const {
    if (m < M && k < K) return ldg4(p + (size_t)m * ld + k);
    return make4(0.f);
  }
};
struct AColMajor {            // A(m,k) = p[k*ld + m]      (vector along m)
  static constexpr bool kVecK = false;
  const float* p; int ld, M, K;
  __device__ __forceinline__ float4 load(int m, int k) const {
    if (m < M && k < K) return ldg4(p + (size_t)k * ld + m);
    return make4(0.f);
  }
};
// 3x3 SAME (pad 1/1) im2col over an NHWC tensor x[B,H,W,C] (pixel stride ldx floats):
//   A(m = pixel, k = tap*C + c) = x[pixel + off(tap)][c]  or 0 outside the image.
// `flip` mirrors the tap (used by dgrad).
struct AIm2col3 {
  static constexpr bool kVecK = true;
  const float* p; int ldx, B, H, W, C; bool flip;
  __device__ __forceinline__ float4 load(int m, int k) const {
    if (m >= B * H * W || k >= 9 * C) return make4(0.f);
    int tap = k / C, c = k - tap * C;
    if (flip) tap = 8 - tap;
    const int x = m % W, t = m / W, y = t % H;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if ((unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) return make4(0.f);
    return ldg4(p + (size_t)(m + (tap / 3 - 1) * W + (tap % 3 - 1)) * ldx + c);
  }
};
// transpose of the above for wgrad:  A(m = tap*C + c, k = pixel)   (vector along m, i.e. along c)
struct AIm2col3T {
  static constexpr bool kVecK = false;
  const float* p; int ldx, B, H, W, C;
  __device__ __forceinline__ float4 load(int m, int k) const {
    if (k >= B * H * W || m >= 9 * C) return make4(0.f);
    const int tap = m / C, c = m - tap * C;
    const int x = k % W, t = k / W, y = t % H;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if ((unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) return make4(0.f);
    return ldg4(p + (size_t)(k + (tap / 3 - 1) * W + (tap % 3 - 1)) * ldx + c);
  }
};
// coverage patches:  A(m = pixel of [NB,H,W], k = tap (0..120)) = ap[pixel + off(tap)]   (11x11 SAME, pad 5/5)
struct ACovPatch {
  static constexpr bool kVecK = true;
  const float* p; int NB, H, W;
  __device__ __forceinline__ float4 load(int m, int k) const {
    if (m >= NB * H * W) return make4(0.f);
    const int x = m % W, t = m / W, y = t % H;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tap = k + j;
      const int dy = tap / kCovK - kCovK / 2, dx = tap % kCovK - kCovK / 2;
      v[j] = (tap < kCovTaps && (unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W)
                 ? __ldg(p + m + dy * W + dx) : 0.f;
    }
    return make_float4(v[0], v[1], v[2], v[3]);
  }
};
// coverage patches, transposed:  A(m = tap (0..120), k = pixel of [NB,H,W]) = ap[pixel + off(tap)]
struct ACovPatchT {
  static constexpr bool kVecK = true;
  const float* p; int NB, H, W;
  __device__ __forceinline__ float get(int tap, int pix) const {
    if (pix >= NB * H * W) return 0.f;
    const int x = pix % W, t = pix / W, y = t % H;
    const int dy = tap / kCovK - kCovK / 2, dx = tap % kCovK - kCovK / 2;
    if ((unsigned)(y + dy) >= (unsigned)H || (unsigned)(x + dx) >= (unsigned)W) return 0.f;
    return __ldg(p + pix + dy * W + dx);
  }
  __device__ __forceinline__ float4 load(int m, int k) const {
    if (m >= kCovTaps) return make4(0.f);
    return make_float4(get(m, k), get(m, k + 1), get(m, k + 2), get(m, k + 3));
  }
};

// ----------------------------------------------------------------------------- B loaders
struct BRowMajor {            // B(k,n) = p[k*ld + n]      (vector along n); N%4 handled
  static constexpr bool kVecK = false;
  const float* p; int ld, K, N;
  __device__ __forceinline__ float4 load(int k, int n) const {
    if (k >= K || n >= N) return make4(0.f);
    const float* q = p + (size_t)k * ld + n;
    if (n + 3 < N && ((ld & 3) == 0)) return ldg4(q);
    float4 v = make4(0.f);
    v.x = __ldg(q);
    if (n + 1 < N) v.y = __ldg(q + 1);
    if (n + 2 < N) v.z = __ldg(q + 2);
    if (n + 3 < N) v.w = __ldg(q + 3);
    return v;
  }
};
struct BColMajor {            // B(k,n) = p[n*ld + k]      (vector along k)
  static constexpr bool kVecK = true;
  const float* p; int ld, K, N;
  __device__ __forceinline__ float4 load(int k, int n) const {
    if (k < K && n < N) return ldg4(p + (size_t)n * ld + k);
    return make4(0.f);
  }
};
// dgrad of the 3x3 conv: B(k = tap'*Co + co, n = ci) = w[(8-tap')][ci][co]   w is HWIO [9][Ci][Co]
struct BDgrad3 {
  static constexpr bool kVecK = true;
  const float* w; int Ci, Co;
  __device__ __forceinline__ float4 load(int k, int n) const {
    if (k >= 9 * Co || n >= Ci) return make4(0.f);
    const int tap = k / Co, co = k - tap * Co;
    return ldg4(w + ((size_t)(8 - tap) * Ci + n) * Co + co);
  }
};

// ----------------------------------------------------------------------------- epilogues
struct EpiStore {             // C[m*ld+n] (=|+=) acc (+ bias[n]);  optional column statistics
  float* c; int ld; const float* bias; bool accumulate; double* stats;   // stats: [2][N] or null
};
struct EpiAtomic {            // atomicAdd(C[m*ld+n], acc)        (split-K partial sums)
  float* c; int ld;
};

template <class Epi> struct EpiTraits;
template <> struct EpiTraits<EpiStore> { static constexpr bool kAtomic = false; };
template <> struct EpiTraits<EpiAtomic> { static constexpr bool kAtomic = true; };

// ----------------------------------------------------------------------------- the kernel
template <class Cfg, class AL, class BL, class Epi>
__global__ void __launch_bounds__(Cfg::NT)
gemm_kernel(const AL A, const BL Bm, const Epi epi, const int M, const int N, const int K, const int k_per_split) {
  pdl_sync();
  constexpr int BM = Cfg::BM, BN = Cfg::BN, RM = Cfg::RM, RN = Cfg::RN, NT = Cfg::NT;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  constexpr int NA = (BM * BK / 4 + NT - 1) / NT, NB = (BN * BK / 4 + NT - 1) / NT;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const int tid = threadIdx.x;
  const int tx = tid % Cfg::NTX, ty = tid / Cfg::NTX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int ntiles = (kend - kbeg + BK - 1) / BK;

  float4 ra[NA], rb[NB];
  float acc[RM * 4][RN * 4];
#pragma unroll
  for (int i = 0; i < RM * 4; ++i)
#pragma unroll
    for (int j = 0; j < RN * 4; ++j) acc[i][j] = 0.f;

  auto gload = [&](int kt) {
    const int k0 = kbeg + kt * BK;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + i * NT;
      if (idx < BM * BK / 4) {
        int mm, kk;
        if (AL::kVecK) { kk = (idx % (BK / 4)) * 4; mm = idx / (BK / 4); }
        else           { mm = (idx % (BM / 4)) * 4; kk = idx / (BM / 4); }
        const int k = k0 + kk;
        ra[i] = (k < kend) ? A.load(m0 + mm, k) : make4(0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + i * NT;
      if (idx < BN * BK / 4) {
        int nn, kk;
        if (BL::kVecK) { kk = (idx % (BK / 4)) * 4; nn = idx / (BK / 4); }
        else           { nn = (idx % (BN / 4)) * 4; kk = idx / (BN / 4); }
        const int k = k0 + kk;
        rb[i] = (k < kend) ? Bm.load(k, n0 + nn) : make4(0.f);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + i * NT;
      if (idx < BM * BK / 4) {
        if (AL::kVecK) {
          const int kk = (idx % (BK / 4)) * 4, mm = idx / (BK / 4);
          As[buf][kk + 0][mm] = ra[i].x; As[buf][kk + 1][mm] = ra[i].y;
          As[buf][kk + 2][mm] = ra[i].z; As[buf][kk + 3][mm] = ra[i].w;
        } else {
          const int mm = (idx % (BM / 4)) * 4, kk = idx / (BM / 4);
          *reinterpret_cast<float4*>(&As[buf][kk][mm]) = ra[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + i * NT;
      if (idx < BN * BK / 4) {
        if (BL::kVecK) {
          const int kk = (idx % (BK / 4)) * 4, nn = idx / (BK / 4);
          Bs[buf][kk + 0][nn] = rb[i].x; Bs[buf][kk + 1][nn] = rb[i].y;
          Bs[buf][kk + 2][nn] = rb[i].z; Bs[buf][kk + 3][nn] = rb[i].w;
        } else {
          const int nn = (idx % (BN / 4)) * 4, kk = idx / (BN / 4);
          *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = rb[i];
        }
      }
    }
  };

  if (ntiles > 0) { gload(0); sstore(0); }
  __syncthreads();
  for (int kt = 0; kt < ntiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) gload(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[RM * 4], b[RN * 4];
#pragma unroll
      for (int f = 0; f < RM; ++f) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][f * (BM / RM) + ty * 4]);
        a[f * 4 + 0] = v.x; a[f * 4 + 1] = v.y; a[f * 4 + 2] = v.z; a[f * 4 + 3] = v.w;
      }
#pragma unroll
      for (int f = 0; f < RN; ++f) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][f * (BN / RN) + tx * 4]);
        b[f * 4 + 0] = v.x; b[f * 4 + 1] = v.y; b[f * 4 + 2] = v.z; b[f * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < RM * 4; ++i)
#pragma unroll
        for (int j = 0; j < RN * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ntiles) sstore(buf ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  if constexpr (EpiTraits<Epi>::kAtomic) {
#pragma unroll
    for (int fi = 0; fi < RM; ++fi)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + fi * (BM / RM) + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int fj = 0; fj < RN; ++fj)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + fj * (BN / RN) + tx * 4 + j;
            if (n < N) atomicAdd(epi.c + (size_t)m * epi.ld + n, acc[fi * 4 + i][fj * 4 + j]);
          }
      }
  } else {
    const bool vec_ok = ((epi.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(epi.c) & 15) == 0);
#pragma unroll
    for (int fi = 0; fi < RM; ++fi)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + fi * (BM / RM) + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int fj = 0; fj < RN; ++fj) {
          const int n = n0 + fj * (BN / RN) + tx * 4;
          if (n >= N) continue;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = acc[fi * 4 + i][fj * 4 + j];
            if (epi.bias != nullptr && n + j < N) v[j] += __ldg(epi.bias + n + j);
          }
          float* dst = epi.c + (size_t)m * epi.ld + n;
          if (vec_ok && n + 3 < N) {
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if (epi.accumulate) { const float4 p = *reinterpret_cast<const float4*>(dst); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
            *reinterpret_cast<float4*>(dst) = o;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < N) dst[j] = epi.accumulate ? dst[j] + v[j] : v[j];
          }
        }
      }
    if (epi.stats != nullptr) {       // per-column sum / sum of squares of the raw accumulators
      float* s_sum = &As[0][0][0];    // reuse tile memory (everyone is past the main loop)
      float* s_sq = s_sum + BN;
      __syncthreads();
      for (int i = tid; i < 2 * BN; i += NT) s_sum[i] = 0.f;
      __syncthreads();
#pragma unroll
      for (int fj = 0; fj < RN; ++fj)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int i = 0; i < RM * 4; ++i) { const float x = acc[i][fj * 4 + j]; s += x; q = fmaf(x, x, q); }
          const int nl = fj * (BN / RN) + tx * 4 + j;
          atomicAdd(&s_sum[nl], s);
          atomicAdd(&s_sq[nl], q);
        }
      __syncthreads();
      for (int i = tid; i < BN; i += NT) {
        const int n = n0 + i;
        if (n < N) {
          atomicAdd(epi.stats + n, (double)s_sum[i]);
          atomicAdd(epi.stats + N + n, (double)s_sq[i]);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------- launcher
template <class Cfg, class AL, class BL, class Epi>
inline int launch_gemm(cudaStream_t st, const AL& A, const BL& Bm, const Epi& epi, int M, int N, int K, int splits = 1,
                       int prof_cls = kProfOtherGemm) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  prof_set_bytes(4.0 * ((double)M * K + (double)K * N + (double)M * N));
  prof_begin(st, prof_cls, 2.0 * M * N * K);
  int kps = cdiv(cdiv(K, splits), BK) * BK;
  splits = cdiv(K, kps);
  dim3 grid(cdiv(M, Cfg::BM), cdiv(N, Cfg::BN), splits);
  WAP_CHECK(launch_pdl(gemm_kernel<Cfg, AL, BL, Epi>, grid, Cfg::NT, 0, st, A, Bm, epi, M, N, K, kps));
  prof_end(st, prof_cls);
  WAP_LAUNCH_CHECK();
  return 0;
}

// C[M,N] (ldc) = A[M,K] (lda) * B[K,N] (ldb) (+bias) ; picks a tile shape from N and M
int gemm_nn(cudaStream_t st, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            int M, int N, int K, const float* bias = nullptr, bool accumulate = false);
// C[M,N] = A[M,K] * W[N,K]^T
int gemm_nt(cudaStream_t st, const float* A, int lda, const float* W, int ldw, float* C, int ldc,
            int M, int N, int K, const float* bias = nullptr, bool accumulate = false);
// C[M,N] += A[K,M]^T * B[K,N]   (reduction over the long dimension K, split-K + atomics)
int gemm_tn_atomic(cudaStream_t st, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                   int M, int N, int K);

}  // namespace wap
