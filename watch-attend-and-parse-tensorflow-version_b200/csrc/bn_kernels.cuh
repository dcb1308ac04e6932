// Batch-norm (+ReLU +dropout +optional 2x2 SAME pool) forward and backward kernels of the DenseNet encoder
// (tf.layers.batch_normalization -> relu -> dropout -> pooling, model-single-GPU.py:81-84, 98-101, 107-110, 124-128), templated on the
// storage type of the raw conv output / activation / gradient: fp32 (modes fp32, tf32, tf32x3) or bf16 (mode bf16).  Arithmetic and
// statistics are always fp32 / fp64.  Shared by encoder.cu and encoder_bf16.cu.
#pragma once
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace wap {
namespace {

// =====================================================================================
// batch-norm (+ReLU +dropout +optional 2x2 SAME pool) forward
// =====================================================================================
struct BnActArgs {
  const void* raw; int B, Hi, Wi, C;         // raw conv output [B,Hi,Wi,C] (fp32 or bf16: template parameter TRaw)
  void* out; int Ho, Wo, ld_out;             // destination pixels [B,Ho,Wo] x channel slice (row stride ld_out elements; fp32 or bf16: TOut)
  const double* stats;                       // [2][C] sum / sum of squares (training), else null
  float* mmean; float* mvar;                 // moving statistics (read when !training; updated when training)
  const float* gamma; const float* beta;
  float* saved;                              // [2][kMaxBnC]: mean, invstd (training) or null
  int training, dropout; uint32_t seed, site;
  const uint32_t* seed_src;                  // non-null: the seed is read from device memory (CUDA-graph replay with a fresh seed)
  int round_tf32;                            // round the stored activation to TF32 (RNA): operand of a tcgen05 conv
};

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void bn_prologue(const BnActArgs& a, float* s_scale, float* s_shift) {
  const double cnt = (double)a.B * a.Hi * a.Wi;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    double mean, var;
    if (a.training) {
      mean = a.stats[c] / cnt;
      var = a.stats[a.C + c] / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
    } else {
      mean = a.mmean[c]; var = a.mvar[c];
    }
    const float invstd = (float)(1.0 / sqrt(var + (double)kBnEps));
    const float sc = __ldg(a.gamma + c) * invstd;
    s_scale[c] = sc;
    s_shift[c] = __ldg(a.beta + c) - (float)mean * sc;
    if (a.training && blockIdx.x == 0) {
      if (a.saved != nullptr) { a.saved[c] = (float)mean; a.saved[kMaxBnC + c] = invstd; }
      const double unb = var * (cnt / fmax(cnt - 1.0, 1.0));      // FusedBatchNorm: Bessel-corrected moving variance
      a.mmean[c] = a.mmean[c] * kBnMomentum + (float)mean * (1.f - kBnMomentum);
      a.mvar[c] = a.mvar[c] * kBnMomentum + (float)unb * (1.f - kBnMomentum);
    }
  }
}

__device__ __forceinline__ float bn_relu_drop(float v, float sc, float sh, bool dropout, uint32_t seed, uint32_t site, uint32_t idx) {
  float y = fmaxf(fmaf(v, sc, sh), 0.f);
  if (dropout) y = drop_keep(seed, site, idx) ? y * (1.f / (1.f - kDropRate)) : 0.f;
  return y;
}

// 4 consecutive elements as floats / back (fp32: one 16-byte access, bf16: one 8-byte access)
template <typename T> __device__ __forceinline__ void ld4(const void* base, size_t idx, float* v);
template <> __device__ __forceinline__ void ld4<float>(const void* base, size_t idx, float* v) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const void* base, size_t idx, float* v) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xFFFF0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xFFFF0000u);
}
template <typename T> __device__ __forceinline__ void st4(void* base, size_t idx, const float* v);
template <> __device__ __forceinline__ void st4<float>(void* base, size_t idx, const float* v) {
  *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(void* base, size_t idx, const float* v) {
  uint2 t;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t.x) : "f"(v[1]), "f"(v[0]));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t.y) : "f"(v[3]), "f"(v[2]));
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = t;
}

// POOL: 0 none, 1 max 2x2 SAME, 2 avg 2x2 SAME (valid-count divisor)
template <int POOL, typename TRaw, typename TOut>
__global__ void __launch_bounds__(256) bn_act_kernel(const BnActArgs a) {
  pdl_sync();
  const uint32_t seed = (a.dropout && a.seed_src != nullptr) ? __ldg(a.seed_src) : a.seed;
  __shared__ float s_scale[kMaxBnC], s_shift[kMaxBnC];
  bn_prologue(a, s_scale, s_shift);
  __syncthreads();
  const int C4 = a.C / 4;
  const size_t total = (size_t)a.B * a.Ho * a.Wo * C4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t po = i / C4;
    float r[4];
    if (POOL == 0) {
      float vv[4];
      ld4<TRaw>(a.raw, po * a.C + c, vv);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        r[j] = bn_relu_drop(vv[j], s_scale[c + j], s_shift[c + j], a.dropout, seed, a.site, (uint32_t)(po * a.C + c + j));
    } else {
      const int xo = (int)(po % a.Wo); const size_t t = po / a.Wo; const int yo = (int)(t % a.Ho); const int b = (int)(t / a.Ho);
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = (POOL == 1) ? -INFINITY : 0.f;
      int cnt = 0;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int yi = yo * 2 + dy, xi = xo * 2 + dx;
          if (yi < a.Hi && xi < a.Wi) {
            const size_t pi = ((size_t)b * a.Hi + yi) * a.Wi + xi;
            float vv[4];
            ld4<TRaw>(a.raw, pi * a.C + c, vv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float y = bn_relu_drop(vv[j], s_scale[c + j], s_shift[c + j], a.dropout, seed, a.site, (uint32_t)(pi * a.C + c + j));
              r[j] = (POOL == 1) ? fmaxf(r[j], y) : r[j] + y;
            }
            ++cnt;
          }
        }
      if (POOL == 2) {
        const float inv = 1.f / (float)cnt;
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] *= inv;
      }
    }
    if (a.round_tf32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = tf32_rna(r[j]);
    }
    st4<TOut>(a.out, po * a.ld_out + c, r);
  }
}

// =====================================================================================
// batch-norm (+ReLU +dropout +pool) backward: pass 1 = upstream gradient at the BN output (g) and
// the two per-channel reductions; pass 2 = d raw in place, d gamma / d beta.
// =====================================================================================
struct BnBwdArgs {
  const void* raw; int B, Hi, Wi, C;         // fp32 or bf16 (TRaw)
  const float* dout; int Ho, Wo, ld_dout;    // gradient wrt the (pooled) activated output, channel slice (always fp32)
  const float* saved; const float* gamma; const float* beta;
  void* g; int ld_g;                         // [B*Hi*Wi, ld_g] fp32 or bf16 (TG)
  double* red;                               // [2][C]: sum g, sum g*xhat
  double* red_next;                          // the apply kernel zeroes this buffer (2*kMaxBnC doubles) for the next BN backward
  float* dgamma; float* dbeta;               // += (pass 2)
  int dropout; uint32_t seed, site;
  const uint32_t* seed_src;
  int round_tf32;                            // round d raw to TF32 (operand of the tcgen05 dgrad)
  unsigned* bar;                             // zeroed counter of the single-kernel form (bn_bwd_fused_kernel), or null: two kernels
};

template <int POOL, typename TRaw, typename TG>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdArgs a, const int pix_per_cta) {
  pdl_sync();
  const uint32_t seed = (a.dropout && a.seed_src != nullptr) ? __ldg(a.seed_src) : a.seed;
  __shared__ float s_scale[kMaxBnC], s_shift[kMaxBnC], s_mean[kMaxBnC], s_inv[kMaxBnC];
  __shared__ float s_sg[kMaxBnC], s_sgx[kMaxBnC];
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const float mean = a.saved[c], inv = a.saved[kMaxBnC + c];
    const float sc = __ldg(a.gamma + c) * inv;
    s_scale[c] = sc; s_shift[c] = __ldg(a.beta + c) - mean * sc; s_mean[c] = mean; s_inv[c] = inv;
    s_sg[c] = 0.f; s_sgx[c] = 0.f;
  }
  __syncthreads();
  const int C4 = a.C / 4, rows = blockDim.x / C4;
  const int c = (threadIdx.x % C4) * 4, prow = threadIdx.x / C4;
  const size_t npix = (size_t)a.B * a.Hi * a.Wi;
  const size_t p0 = (size_t)blockIdx.x * pix_per_cta;
  const size_t p1 = min(npix, p0 + (size_t)pix_per_cta);
  float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
  if (prow < rows) {
    size_t pstart = p0 + prow;
    if (POOL == 0) {
      // four pixels per round with all eight loads issued before the first use: in blocks 1-2 a CTA's slice is a handful of pixels per
      // thread and the loop was a chain of dependent load latencies (8 us per launch, 64 launches per step on the backward's critical path)
      for (; pstart < p1; pstart += 4 * (size_t)rows) {
        float vv[4][4], up[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t p = pstart + (size_t)u * rows;
          if (p < p1) {
            ld4<TRaw>(a.raw, p * a.C + c, vv[u]);
            const float4 d = ldg4(a.dout + p * a.ld_dout + c);
            up[u][0] = d.x; up[u][1] = d.y; up[u][2] = d.z; up[u][3] = d.w;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t p = pstart + (size_t)u * rows;
          if (p >= p1) break;
          float gg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float pre = fmaf(vv[u][j], s_scale[c + j], s_shift[c + j]);
            float gj = (pre > 0.f) ? up[u][j] : 0.f;
            if (a.dropout) gj = drop_keep(seed, a.site, (uint32_t)(p * a.C + c + j)) ? gj * (1.f / (1.f - kDropRate)) : 0.f;
            gg[j] = gj;
            sg[j] += gj;
            sgx[j] = fmaf(gj, (vv[u][j] - s_mean[c + j]) * s_inv[c + j], sgx[j]);
          }
          st4<TG>(a.g, p * a.ld_g + c, gg);
        }
      }
    }
    for (size_t p = pstart; p < p1; p += rows) {
      float vv[4];
      ld4<TRaw>(a.raw, p * a.C + c, vv);
      float up[4];
      if (POOL == 0) {
        const float4 d = ldg4(a.dout + p * a.ld_dout + c);
        up[0] = d.x; up[1] = d.y; up[2] = d.z; up[3] = d.w;
      } else {
        const int xi = (int)(p % a.Wi); const size_t t = p / a.Wi; const int yi = (int)(t % a.Hi); const int b = (int)(t / a.Hi);
        const int yo = yi >> 1, xo = xi >> 1;
        const float4 d = ldg4(a.dout + (((size_t)b * a.Ho + yo) * a.Wo + xo) * a.ld_dout + c);
        const float dd[4] = {d.x, d.y, d.z, d.w};
        if (POOL == 2) {
          const int cnt = min(2, a.Hi - yo * 2) * min(2, a.Wi - xo * 2);
          const float inv = 1.f / (float)cnt;
#pragma unroll
          for (int j = 0; j < 4; ++j) up[j] = dd[j] * inv;
        } else {
          // max pool: the gradient goes to the first window cell (row-major) attaining the maximum
#pragma unroll
          for (int j = 0; j < 4; ++j) up[j] = 0.f;
          float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          int arg[4] = {0, 0, 0, 0};
#pragma unroll
          for (int wy = 0; wy < 2; ++wy)
#pragma unroll
            for (int wx = 0; wx < 2; ++wx) {
              const int y2 = yo * 2 + wy, x2 = xo * 2 + wx;
              if (y2 < a.Hi && x2 < a.Wi) {
                float qq[4];
                ld4<TRaw>(a.raw, (((size_t)b * a.Hi + y2) * a.Wi + x2) * a.C + c, qq);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float y = fmaxf(fmaf(qq[j], s_scale[c + j], s_shift[c + j]), 0.f);
                  if (y > best[j]) { best[j] = y; arg[j] = wy * 2 + wx; }
                }
              }
            }
          const int me = (yi & 1) * 2 + (xi & 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) if (arg[j] == me) up[j] = dd[j];
        }
      }
      float gg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pre = fmaf(vv[j], s_scale[c + j], s_shift[c + j]);
        float gj = (pre > 0.f) ? up[j] : 0.f;
        if (a.dropout) gj = drop_keep(seed, a.site, (uint32_t)(p * a.C + c + j)) ? gj * (1.f / (1.f - kDropRate)) : 0.f;
        gg[j] = gj;
        sg[j] += gj;
        sgx[j] = fmaf(gj, (vv[j] - s_mean[c + j]) * s_inv[c + j], sgx[j]);
      }
      st4<TG>(a.g, p * a.ld_g + c, gg);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&s_sg[c + j], sg[j]); atomicAdd(&s_sgx[c + j], sgx[j]); }
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < a.C; cc += blockDim.x) {
    atomicAdd(a.red + cc, (double)s_sg[cc]);
    atomicAdd(a.red + a.C + cc, (double)s_sgx[cc]);
  }
}

template <typename TRaw, typename TG>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs a) {
  pdl_sync();
  if (blockIdx.x == 0 && a.red_next != nullptr)
    for (int i = threadIdx.x; i < 2 * kMaxBnC; i += blockDim.x) a.red_next[i] = 0.0;
  __shared__ float s_k[kMaxBnC], s_mean[kMaxBnC], s_inv[kMaxBnC], s_mg[kMaxBnC], s_mgx[kMaxBnC];
  const double cnt = (double)a.B * a.Hi * a.Wi;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const float inv = a.saved[kMaxBnC + c];
    s_mean[c] = a.saved[c]; s_inv[c] = inv; s_k[c] = __ldg(a.gamma + c) * inv;
    s_mg[c] = (float)(a.red[c] / cnt); s_mgx[c] = (float)(a.red[a.C + c] / cnt);
    if (blockIdx.x == 0) { a.dbeta[c] += (float)a.red[c]; a.dgamma[c] += (float)a.red[a.C + c]; }
  }
  __syncthreads();
  const int C4 = a.C / 4;
  const size_t total = (size_t)a.B * a.Hi * a.Wi * C4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t p = i / C4;
    float vv[4], gg[4];
    ld4<TRaw>(a.raw, p * a.C + c, vv);
    ld4<TG>(a.g, p * a.ld_g + c, gg);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (vv[j] - s_mean[c + j]) * s_inv[c + j];
      gg[j] = s_k[c + j] * (gg[j] - s_mg[c + j] - xh * s_mgx[c + j]);
      if (a.round_tf32) gg[j] = tf32_rna(gg[j]);
    }
    st4<TG>(a.g, p * a.ld_g + c, gg);
  }
}

// Both passes in ONE kernel for feature maps small enough that every thread keeps its elements in registers between them (blocks 1-2, and
// the 24-channel maps of block 0): pass 1 as above, then a grid-wide barrier (a counter in global memory; the grid is at most 2 CTAs per SM,
// so all of it is resident - every kernel of this library releases its dependents at its first instruction, when all its CTAs are running), then
// pass 2 on the registers.  g is written once instead of written, re-read and rewritten, and the backward chain loses one launch per BN.
constexpr int kBnFusedE = 4;        // pixels per thread row
template <typename TRaw, typename TG>
__global__ void __launch_bounds__(256, 2) bn_bwd_fused_kernel(const BnBwdArgs a, const int pix_per_cta) {
  pdl_sync();
  const uint32_t seed = (a.dropout && a.seed_src != nullptr) ? __ldg(a.seed_src) : a.seed;
  __shared__ float s_scale[kMaxBnC], s_shift[kMaxBnC], s_mean[kMaxBnC], s_inv[kMaxBnC];
  __shared__ float s_sg[kMaxBnC], s_sgx[kMaxBnC];
  const int C4 = a.C / 4, rows = blockDim.x / C4;
  const int c = (threadIdx.x % C4) * 4, prow = threadIdx.x / C4;
  const size_t npix = (size_t)a.B * a.Hi * a.Wi;
  const size_t p0 = (size_t)blockIdx.x * pix_per_cta;
  const size_t p1 = min(npix, p0 + (size_t)pix_per_cta);
  float rv[kBnFusedE][4], gg[kBnFusedE][4];
  const bool act = prow < rows;
#pragma unroll
  for (int u = 0; u < kBnFusedE; ++u) {          // every load of the CTA in flight before the per-channel constants are staged
    const size_t p = p0 + prow + (size_t)u * rows;
    if (act && p < p1) {
      ld4<TRaw>(a.raw, p * a.C + c, rv[u]);
      const float4 d = ldg4(a.dout + p * a.ld_dout + c);
      gg[u][0] = d.x; gg[u][1] = d.y; gg[u][2] = d.z; gg[u][3] = d.w;
    }
  }
  for (int cc = threadIdx.x; cc < a.C; cc += blockDim.x) {
    const float mean = a.saved[cc], inv = a.saved[kMaxBnC + cc];
    const float sc = __ldg(a.gamma + cc) * inv;
    s_scale[cc] = sc; s_shift[cc] = __ldg(a.beta + cc) - mean * sc; s_mean[cc] = mean; s_inv[cc] = inv;
    s_sg[cc] = 0.f; s_sgx[cc] = 0.f;
  }
  __syncthreads();
  float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int u = 0; u < kBnFusedE; ++u) {
    const size_t p = p0 + prow + (size_t)u * rows;
    if (act && p < p1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pre = fmaf(rv[u][j], s_scale[c + j], s_shift[c + j]);
        float gj = (pre > 0.f) ? gg[u][j] : 0.f;
        if (a.dropout) gj = drop_keep(seed, a.site, (uint32_t)(p * a.C + c + j)) ? gj * (1.f / (1.f - kDropRate)) : 0.f;
        const float xh = (rv[u][j] - s_mean[c + j]) * s_inv[c + j];
        gg[u][j] = gj; rv[u][j] = xh;                  // keep g and xhat
        sg[j] += gj;
        sgx[j] = fmaf(gj, xh, sgx[j]);
      }
    }
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&s_sg[c + j], sg[j]); atomicAdd(&s_sgx[c + j], sgx[j]); }
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < a.C; cc += blockDim.x) {
    atomicAdd(a.red + cc, (double)s_sg[cc]);
    atomicAdd(a.red + a.C + cc, (double)s_sgx[cc]);
  }
  // ---- grid barrier: every CTA's partial sums are in a.red
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(a.bar, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory"); } while (v < gridDim.x);
  }
  __syncthreads();
  // ---- pass 2 on the registers
  const double cnt = (double)npix;
  float* s_k = s_scale; float* s_mg = s_sg; float* s_mgx = s_sgx;      // reuse
  for (int cc = threadIdx.x; cc < a.C; cc += blockDim.x) {
    const double tg = __ldcg(a.red + cc), tgx = __ldcg(a.red + a.C + cc);
    s_k[cc] = __ldg(a.gamma + cc) * s_inv[cc];
    s_mg[cc] = (float)(tg / cnt); s_mgx[cc] = (float)(tgx / cnt);
    if (blockIdx.x == 0) { a.dbeta[cc] += (float)tg; a.dgamma[cc] += (float)tgx; }
  }
  if (blockIdx.x == 0 && a.red_next != nullptr)
    for (int i = threadIdx.x; i < 2 * kMaxBnC; i += blockDim.x) a.red_next[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kBnFusedE; ++u) {
    const size_t p = p0 + prow + (size_t)u * rows;
    if (act && p < p1) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = s_k[c + j] * (gg[u][j] - s_mg[c + j] - rv[u][j] * s_mgx[c + j]);
        if (a.round_tf32) o[j] = tf32_rna(o[j]);
      }
      st4<TG>(a.g, p * a.ld_g + c, o);
    }
  }
}

static inline int ew_grid(size_t total_vec) {
  return (int)std::min<size_t>(cdivz(total_vec, 256), 148 * 8);
}

template <typename TRaw = float, typename TOut = float>
static int run_bn_act(cudaStream_t st, int pool, const BnActArgs& a) {
  const size_t total = (size_t)a.B * a.Ho * a.Wo * (a.C / 4);
  const int grid = ew_grid(total);
  if (pool == 0) WAP_CHECK(launch_pdl(bn_act_kernel<0, TRaw, TOut>, grid, 256, 0, st, a));
  else if (pool == 1) WAP_CHECK(launch_pdl(bn_act_kernel<1, TRaw, TOut>, grid, 256, 0, st, a));
  else WAP_CHECK(launch_pdl(bn_act_kernel<2, TRaw, TOut>, grid, 256, 0, st, a));
  WAP_LAUNCH_CHECK();
  return 0;
}

// second pass only (the first one ran inside the producing GEMM's epilogue, TcEpi::bnb_*)
template <typename TRaw = float, typename TG = float>
static int run_bn_bwd_apply(cudaStream_t st, const BnBwdArgs& a) {
  const size_t npix = (size_t)a.B * a.Hi * a.Wi;
  WAP_CHECK(launch_pdl(bn_bwd_apply_kernel<TRaw, TG>, ew_grid(npix * (a.C / 4)), 256, 0, st, a));
  WAP_LAUNCH_CHECK();
  return 0;
}
template <typename TRaw = float, typename TG = float>
static int run_bn_bwd(cudaStream_t st, int pool, const BnBwdArgs& a) {
  // a.red was zeroed by the previous BN backward's apply kernel (or by the memset at the start of encoder_backward): a memset
  // node here would sit between two kernels 110 times per step and break their programmatic dependency
  const size_t npix = (size_t)a.B * a.Hi * a.Wi;
  if (pool == 0 && a.bar != nullptr) {      // single-kernel form when every thread can hold its share in registers (<= kBnFusedE pixels per thread row)
    const int rows = 256 / (a.C / 4);
    const int grid = (int)std::min<size_t>(296, cdivz(npix, (size_t)rows));
    const int ppc = (int)cdivz(npix, (size_t)grid);
    if (ppc <= kBnFusedE * rows) {
      WAP_CHECK(launch_pdl(bn_bwd_fused_kernel<TRaw, TG>, (int)cdivz(npix, (size_t)ppc), 256, 0, st, a, ppc));
      WAP_LAUNCH_CHECK();
      return 0;
    }
  }
  // CTAs per SM of the reduce pass: every CTA ends with 2*C fp64 atomics onto the same 2*C addresses, so more CTAs = more contention
  // (tunable for experiments: WAP_BN_BWD_CTAS_PER_SM).  2 = what is resident with the kernel's 94 registers: one wave; with 4 the 32768-pixel maps of
  // block 0 ran 1.4-2 waves of CTAs whose fixed cost (constant staging, shared + global atomics) is half their time: 9.85 vs 9.89 ms per cfg2 step
  static const int per_sm = [] { const char* e = getenv("WAP_BN_BWD_CTAS_PER_SM"); const int v = e ? atoi(e) : 2; return v > 0 ? v : 2; }();
  int ppc = (int)std::max<size_t>(64, cdivz(npix, (size_t)148 * per_sm));
  const int round = 4 * (256 / (a.C / 4));            // one unrolled round of the reduce loop = 4 pixels per thread row: avoid a nearly empty last round
  if (round <= ppc) ppc = cdiv(ppc, round) * round;
  const int grid = (int)cdivz(npix, ppc);
  if (pool == 0) WAP_CHECK(launch_pdl(bn_bwd_reduce_kernel<0, TRaw, TG>, grid, 256, 0, st, a, ppc));
  else if (pool == 1) WAP_CHECK(launch_pdl(bn_bwd_reduce_kernel<1, TRaw, TG>, grid, 256, 0, st, a, ppc));
  else WAP_CHECK(launch_pdl(bn_bwd_reduce_kernel<2, TRaw, TG>, grid, 256, 0, st, a, ppc));
  WAP_LAUNCH_CHECK();
  WAP_CHECK(launch_pdl(bn_bwd_apply_kernel<TRaw, TG>, ew_grid(npix * (a.C / 4)), 256, 0, st, a));
  WAP_LAUNCH_CHECK();
  return 0;
}


}  // namespace
}  // namespace wap
