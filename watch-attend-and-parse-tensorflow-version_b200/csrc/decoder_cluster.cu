// Greedy decoder, one THREAD-BLOCK CLUSTER per image: the whole token loop of WAP.get_word / get_sample(argmax)
// (model-single-GPU.py:379-439, :461-483) with NO grid-wide barrier.
//
// The persistent kernel of decoder_persist.cu slices the weights over all 148 SMs and pays six grid barriers (atomic counter + L2
// round trip, ~1.4 us each, plus an L2 round trip for the exchanged vector) per token.  Here the 16 CTAs of a cluster own one image:
// every dependency of a step (GRU-1 -> speller -> attention energies -> softmax -> context -> GRU-2 -> logits -> arg-max) is an
// exchange INSIDE the cluster - remote shared-memory stores (DSMEM) followed by barrier.cluster (~0.3 us) - and images never
// wait for each other.  Per CTA (rank c of 16):
//   resident in shared memory : GRU-1 weights of its 16 hidden units (6 columns x 256), the pre_h-side GRU-2 columns (3 x 256), its
//                               [121 x 32] slice of the fused coverage filter Q' = Q*U_f, its 8 rows of Wo
//   streamed from L2 per step : its 32 speller columns of W_a, its context-side GRU-2 / logit columns of W_c_z_r, W_c_h_nl, Wc and its
//                               logit columns of Wy, Wh (240 KB per step: coalesced 64..128-byte row segments, many loads in flight)
//   attention                 : its 32 attention channels for ALL positions - coverage stencil as a register-tiled FFMA loop (8 positions
//                               x 4 channels per thread, the 11 x 18 alpha_past window of a tile row read once per filter row), tanh, V_a dot;
//                               the 16 partial energies of a position are reduce-scattered to the CTA that owns the position
//   softmax                   : two-level (local max / sum per owner, 16 pairs all-gathered), alpha and alpha_past kept in every CTA
//   context                   : its 40..44 annotation channels
// Eight exchanges = seven cluster barriers per token.  Token, stop rule and the next embedding are evaluated redundantly by every CTA.
// Limits: h*w <= 1024, (h+10)*(roundup8(w)+10) <= 4096; larger grids use decoder_persist.cu.
#include <cooperative_groups.h>

#include <algorithm>

#include "decoder.cuh"

namespace cg = cooperative_groups;

namespace wap {

int attn_precompute(cudaStream_t st, const float* params, const float* ann, const float* a_mask, int B, int h, int w,
                    float* watch, float* h0, float* qf, float* mean_scratch);

int g_cluster_prof = 0;      // debug: per-stage cycle counters of cluster 0 into the workspace tail
namespace {

constexpr int kCL = 16;                    // CTAs per cluster (non-portable size: opt-in attribute)
constexpr int kT = 256;                    // threads per CTA
constexpr int kUPC = kN / kCL;             // 16 hidden units per CTA
constexpr int kSPC = kNA / kCL;            // 32 speller / attention channels per CTA
constexpr int kLPC = kMw / kCL;            // 16 pre-maxout logit columns per CTA (8 maxout units)
constexpr int kMaxLc = 1024;               // annotation positions per image
constexpr int kPadC = 4096;                // floats of the zero-bordered alpha_past image
constexpr int kOwnMax = kMaxLc / kCL;      // 64 positions owned per CTA for the softmax

struct COff {   // offsets (floats) into the flat parameter buffer
  int U_f_b, W_a, W_a_b, V_a, V_a_b, W_yz_yr, b_yz_yr, U_hz_hr, W_yh, b_yh, U_rh, U_hz_hr_nl, b_hz_hr_nl, W_c_z_r, U_rh_nl,
      b_rh_nl, W_c_h_nl, embed, Wc, bc, Wh, bh, Wy, by, Wo, bo;
};
struct CArgs {
  const float* params; const float* ann; const float* watch; const float* qf; const float* h0;
  COff o;
  int H, W, L, maxlen, ignore_eol;
  int* tokens; int* lengths; int* steps_run;     // [rows][maxlen], [rows], max steps over rows (atomicMax)
  long long* prof;                               // optional [12]: SM cycles of cluster 0 / rank 0 per stage (setup, 1..8, barriers)
};

// shared-memory carve-up (floats)
constexpr int kWA = kUPC * 6 * kN;         // GRU-1 columns          24576
constexpr int kWB = kUPC * 3 * kN;         // GRU-2 pre_h-side       12288
constexpr int kQS = kCovTaps * kSPC;       // Q' slice                3872
constexpr int kWO = (kLPC / 2) * 112;      // Wo rows                  896
constexpr int kVec = 256 + 256 + 256 + 512 + 688 + 256;   // emb | h | pre_h | sp | ctx | (spare)
constexpr int kSmemFloatsC = kWA + kWB + kQS + kWO + kVec + kPadC + kMaxLc /*alpha*/ + kCL * kOwnMax /*epart*/ + 2 * kCL /*ms*/ +
                             kCL * 112 /*lpart*/ + 8 * 64 /*red*/ + 64 /*own gate values*/ + 256 /*misc*/ + 320 /*biases*/;

__device__ __forceinline__ void lcol(float* dst, const float* W, int K, int N, int col) {
  for (int k = threadIdx.x; k < K; k += kT) dst[k] = __ldg(W + (size_t)k * N + col);
}

__global__ void __launch_bounds__(kT, 1) greedy_cluster_kernel(const CArgs a) {
  pdl_sync();
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int c = (int)cluster.block_rank();             // rank inside the cluster
  const int row = blockIdx.x / kCL;                     // image handled by this cluster
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = a.params; const COff& o = a.o;
  const int L = a.L, Wd = a.W, Hd = a.H;
  const int W8 = (Wd + 7) & ~7, PW = W8 + 10;
  const int LPC = (L + kCL - 1) / kCL;                 // positions owned per CTA (softmax)
  // ---- carve-up
  float* WA = sm;
  float* WB = WA + kWA;
  float* Qs = WB + kWB;
  float* WoR = Qs + kQS;
  float* v_emb = WoR + kWO;
  float* v_h = v_emb + 256;
  float* v_pre = v_h + 256;
  float* v_sp = v_pre + 256;
  float* v_ctx = v_sp + 512;                            // 688
  float* v_sp2 = v_ctx + 688;                           // spare 256
  float* apad = v_sp2 + 256;                            // [PH][PW] zero-bordered alpha_past
  float* alpha = apad + kPadC;                          // [L]
  float* epart = alpha + kMaxLc;                        // [kCL][kOwnMax] partial energies of the positions this CTA owns
  float* ms = epart + kCL * kOwnMax;                    // [kCL][2] local max / sum of every owner
  float* lpart = ms + 2 * kCL;                          // [kCL][112] partial output logits
  float* red = lpart + kCL * 112;                       // [8][64]
  float* own = red + 8 * 64;                            // [16][4]: pre_h, g2r, g2z, ph2 of the owned units
  float* misc = own + 64;
  float* bia = misc + 256;                              // [0,48) GRU-1 r|z|h  [48,96) GRU-2 r|z|h  [96,128) speller  [128,144) logit columns  [144,256) bo  [256] V_a_b
  const int j0 = c * kUPC;                              // first owned hidden unit
  const int n0 = c * kLPC;                              // first owned pre-maxout logit column

  // ---- one-off: resident weights
  for (int u = 0; u < kUPC; ++u) {
    const int j = j0 + u;
    float* wa = WA + u * 6 * kN;
    lcol(wa + 0 * kN, P + o.W_yz_yr, kMw, 2 * kN, j);
    lcol(wa + 1 * kN, P + o.W_yz_yr, kMw, 2 * kN, kN + j);
    lcol(wa + 2 * kN, P + o.W_yh, kMw, kN, j);
    lcol(wa + 3 * kN, P + o.U_hz_hr, kN, 2 * kN, j);
    lcol(wa + 4 * kN, P + o.U_hz_hr, kN, 2 * kN, kN + j);
    lcol(wa + 5 * kN, P + o.U_rh, kN, kN, j);
    float* wb = WB + u * 3 * kN;
    lcol(wb + 0 * kN, P + o.U_hz_hr_nl, kN, 2 * kN, j);
    lcol(wb + 1 * kN, P + o.U_hz_hr_nl, kN, 2 * kN, kN + j);
    lcol(wb + 2 * kN, P + o.U_rh_nl, kN, kN, j);
  }
  for (int i = tid; i < kQS; i += kT) Qs[i] = __ldg(a.qf + (size_t)(i / kSPC) * kNA + c * kSPC + (i % kSPC));
  for (int i = tid; i < (kLPC / 2) * 112; i += kT) {
    const int u = i / 112, j = i - u * 112;
    WoR[i] = j < kVocab ? __ldg(P + o.Wo + (size_t)(c * (kLPC / 2) + u) * kVocab + j) : 0.f;
  }
  for (int i = tid; i < kPadC; i += kT) apad[i] = 0.f;
  for (int i = tid; i < kN; i += kT) v_h[i] = __ldg(a.h0 + (size_t)row * kN + i);
  if (tid < kUPC) {
    const int j = j0 + tid;
    bia[tid] = __ldg(P + o.b_yz_yr + j); bia[16 + tid] = __ldg(P + o.b_yz_yr + kN + j); bia[32 + tid] = __ldg(P + o.b_yh + j);
    bia[48 + tid] = __ldg(P + o.b_hz_hr_nl + j); bia[64 + tid] = __ldg(P + o.b_hz_hr_nl + kN + j); bia[80 + tid] = __ldg(P + o.b_rh_nl + j);
    bia[128 + tid] = __ldg(P + o.by + n0 + tid) + __ldg(P + o.bc + n0 + tid) + __ldg(P + o.bh + n0 + tid);
  }
  if (tid < kSPC) bia[96 + tid] = __ldg(P + o.W_a_b + c * kSPC + tid);
  if (tid < 112) bia[144 + tid] = tid < kVocab ? __ldg(P + o.bo + tid) : 0.f;
  if (tid == 0) bia[256] = __ldg(P + o.V_a_b);
  // per-thread constants of the attention stage (channel quad q of the 32 owned channels)
  const float4 att_ub = ldg4(P + o.U_f_b + c * kSPC + (lane & 7) * 4), att_va = ldg4(P + o.V_a + c * kSPC + (lane & 7) * 4);
  int y = -1;                                           // current token (uniform over the cluster)
  int t = 0;
  long long pacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
  auto mark = [&](int ph) { const long long now = clock64(); pacc[ph] += now - tprev; tprev = now; };
  auto csync = [&]() { const long long t0 = clock64(); cluster.sync(); const long long t1 = clock64(); pacc[9] += t1 - t0; tprev = t1; };
  cluster.sync();                                       // every CTA of the cluster is running before the first remote store
  mark(0);

  for (; t < a.maxlen; ++t) {
    // ================================================= 1: GRU-1 for the 16 owned units -> pre_h (all-gather)
    // streamed weights of stage 2 (the 32 speller columns: rows warp + 8 i) go into registers now; they do not depend on the data
    float wsp[32];
    {
      const float* wcol = P + o.W_a + c * kSPC + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) wsp[i] = __ldg(wcol + (size_t)(warp + 8 * i) * kNA);
    }
    for (int i = tid; i < kMw; i += kT) v_emb[i] = y >= 0 ? __ldg(P + o.embed + (size_t)y * kMw + i) : 0.f;
    __syncthreads();
    {   // warp w owns units 2w, 2w+1: twelve 256-long dot products with independent accumulators, one interleaved butterfly
      float acc[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[i] = 0.f;
      const float* w0 = WA + (2 * warp) * 6 * kN;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const int k = lane + 32 * k8;
        const float xe = v_emb[k], xh = v_h[k];
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[i] = fmaf((i % 6) < 3 ? xe : xh, w0[i * kN + k], acc[i]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
      float ph[2];
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int u = 2 * warp + s2; const float* r6 = acc + 6 * s2;
        const float r = sigmoidf_(r6[0] + r6[3] + bia[u]), z = sigmoidf_(r6[1] + r6[4] + bia[16 + u]);
        const float prop = tanhf(fmaf(r6[5], r, r6[2] + bia[32 + u]));
        ph[s2] = z * v_h[j0 + u] + (1.f - z) * prop;
      }
      if (lane == 0) { own[(2 * warp) * 4] = ph[0]; own[(2 * warp + 1) * 4] = ph[1]; }
      if (lane < kCL) { float* d = cluster.map_shared_rank(v_pre, lane); d[j0 + 2 * warp] = ph[0]; d[j0 + 2 * warp + 1] = ph[1]; }
    }
    mark(1);
    csync();
    // ================================================= 2: pre_h-side GRU-2 products (resident) + speller columns (registers) -> sp (all-gather)
    {
      float acc[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) acc[i] = 0.f;
      const float* w0 = WB + (2 * warp) * 3 * kN;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const int k = lane + 32 * k8;
        const float x = v_pre[k];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[i] = fmaf(x, w0[i * kN + k], acc[i]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
      if (lane == 0) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int u = 2 * warp + s2;
          own[u * 4 + 1] = acc[3 * s2 + 0] + bia[48 + u]; own[u * 4 + 2] = acc[3 * s2 + 1] + bia[64 + u]; own[u * 4 + 3] = acc[3 * s2 + 2] + bia[80 + u];
        }
      }
      float sacc = 0.f;                                    // sp[32c + lane]: rows warp + 8 i of W_a held in wsp
#pragma unroll
      for (int i = 0; i < 32; ++i) sacc = fmaf(v_pre[warp + 8 * i], wsp[i], sacc);
      red[warp * 64 + lane] = sacc;
    }
    __syncthreads();
    if (tid < kSPC) {
      float s2 = bia[96 + tid];
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s2 += red[w8 * 64 + tid];
      misc[tid] = s2;
    }
    __syncthreads();
    for (int i = tid; i < kSPC * kCL; i += kT) cluster.map_shared_rank(v_sp, i >> 5)[c * kSPC + (i & 31)] = misc[i & 31];
    mark(2);
    csync();
    // ================================================= 3: attention energies of ALL positions for the 32 owned channels (reduce-scatter)
    {
      const int q = lane & 7, slot = warp * 4 + (lane >> 3);          // channel quad (4 channels), tile slot (32 per pass)
      const int tiles_x = W8 >> 3, ntiles = Hd * tiles_x;
      const int ch = c * kSPC + q * 4;
      const float4 ub = att_ub, va = att_va;
      const float4 sp4 = *reinterpret_cast<const float4*>(v_sp + ch);
      for (int tb = 0; tb < ntiles; tb += 32) {
        const int tile = tb + slot;
        const bool act = tile < ntiles;
        const int ty0 = act ? tile / tiles_x : 0, x0 = act ? (tile - ty0 * tiles_x) * 8 : 0;
        float4 w4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {                                  // watch rows in flight before the stencil
          const int x = min(x0 + i, Wd - 1);
          w4[i] = ldg4(a.watch + ((size_t)row * L + ty0 * Wd + x) * kNA + ch);
        }
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
        if (act) {
#pragma unroll 1
          for (int fy = 0; fy < kCovK; ++fy) {
            const float* prow = apad + (ty0 + fy) * PW + x0;            // padded image: pixel (y, x) sits at (y + 5, x + 5)
            float pv[18];
#pragma unroll
            for (int i = 0; i < 18; ++i) pv[i] = prow[i];
#pragma unroll
            for (int fx = 0; fx < kCovK; ++fx) {
              const float4 qv = *reinterpret_cast<const float4*>(Qs + (fy * kCovK + fx) * kSPC + q * 4);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                acc[i][0] = fmaf(pv[i + fx], qv.x, acc[i][0]); acc[i][1] = fmaf(pv[i + fx], qv.y, acc[i][1]);
                acc[i][2] = fmaf(pv[i + fx], qv.z, acc[i][2]); acc[i][3] = fmaf(pv[i + fx], qv.w, acc[i][3]);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float e = va.x * tanhf(acc[i][0] + ub.x + w4[i].x + sp4.x) + va.y * tanhf(acc[i][1] + ub.y + w4[i].y + sp4.y) +
                    va.z * tanhf(acc[i][2] + ub.z + w4[i].z + sp4.z) + va.w * tanhf(acc[i][3] + ub.w + w4[i].w + sp4.w);
          e += __shfl_xor_sync(0xffffffffu, e, 1); e += __shfl_xor_sync(0xffffffffu, e, 2); e += __shfl_xor_sync(0xffffffffu, e, 4);
          const int x = x0 + i;
          if (act && q == 0 && x < Wd) {
            const int l = ty0 * Wd + x, owner = l / LPC;
            cluster.map_shared_rank(epart, owner)[c * kOwnMax + (l - owner * LPC)] = e;
          }
        }
      }
    }
    mark(3);
    csync();
    // ================================================= 4: owner-local softmax pieces (all-gather of u = exp(e - m_local) and of (m, s))
    // streamed weights of stage 6: per context row k four 64-byte segments - W_c_z_r[k][j0..] (r), W_c_z_r[k][256 + j0..] (z), W_c_h_nl[k][j0..],
    // Wc[k][n0..]; 16 lanes (4 segments x 4 float4) cover a row, a warp-load two rows; thread rows k = 2 warp + sub + 16 i, i < 43.  The first
    // 16 of the 43 float4 of a thread are requested here (stages 4-5 are short and light on registers), the rest inside stage 6.
    const int s6_sub = lane >> 4, s6_seg = (lane >> 2) & 3, s6_qd = lane & 3;
    const float* s6_base = (s6_seg == 0 ? P + o.W_c_z_r + j0 : s6_seg == 1 ? P + o.W_c_z_r + kN + j0 : s6_seg == 2 ? P + o.W_c_h_nl + j0 : P + o.Wc + n0) + s6_qd * 4;
    const int s6_ld = s6_seg < 2 ? 2 * kN : s6_seg == 2 ? kN : kMw;
    const int s6_k0 = warp * 2 + s6_sub;
    float4 wq_a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) wq_a[i] = ldg4(s6_base + (size_t)(s6_k0 + 16 * i) * s6_ld);
    {
      const int l0 = c * LPC, nown = max(0, min(L, l0 + LPC) - l0);
      float e = -INFINITY;
      if (tid < nown) {
        e = bia[256];
#pragma unroll
        for (int r = 0; r < kCL; ++r) e += epart[r * kOwnMax + tid];
      }
      const float m = block_max(e, red);
      const float u = tid < nown ? expf(e - m) : 0.f;
      const float s = block_sum(u, red);
      if (tid < nown) misc[tid] = u;
      __syncthreads();
      for (int i = tid; i < nown * kCL; i += kT) cluster.map_shared_rank(alpha, i / nown)[l0 + (i % nown)] = misc[i % nown];
      if (tid < kCL) { float* d = cluster.map_shared_rank(ms, tid); d[2 * c] = nown > 0 ? m : -INFINITY; d[2 * c + 1] = s; }
    }
    mark(4);
    csync();
    // ================================================= 5: alpha, alpha_past (every CTA), context slice (all-gather)
    {
      float M = -INFINITY;
#pragma unroll
      for (int r = 0; r < kCL; ++r) M = fmaxf(M, ms[2 * r]);
      float S = 0.f;
#pragma unroll
      for (int r = 0; r < kCL; ++r) S += ms[2 * r + 1] * expf(ms[2 * r] - M);
      const float inv = 1.f / S;
      for (int l = tid; l < L; l += kT) {
        const float al = alpha[l] * expf(ms[2 * (l / LPC)] - M) * inv;
        alpha[l] = al;
        const int yy = l / Wd, xx = l - yy * Wd;
        apad[(yy + 5) * PW + xx + 5] += al;
      }
      __syncthreads();
      // context: float4 columns [f0, f0 + nf) of the 171; 16 position groups x 16 float4 lanes
      const int f0 = c < 11 ? 11 * c : 121 + 10 * (c - 11), nf = c < 11 ? 11 : 10;
      const int cq = tid & 15, lg = tid >> 4;
      float4 acc = make4(0.f);
      if (cq < nf) {
        const float* base = a.ann + (size_t)row * L * kD + (f0 + cq) * 4;
        for (int l = lg; l < L; l += 16) {
          const float4 v = ldg4(base + (size_t)l * kD);
          const float al = alpha[l];
          acc.x = fmaf(al, v.x, acc.x); acc.y = fmaf(al, v.y, acc.y); acc.z = fmaf(al, v.z, acc.z); acc.w = fmaf(al, v.w, acc.w);
        }
      }
      float* part = epart;                                   // [16][64] scratch (epart is free until step t+1's stage 3)
      *reinterpret_cast<float4*>(&part[lg * 64 + cq * 4]) = acc;
      __syncthreads();
      if (tid < nf * 4) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) s += part[g * 64 + tid];
        misc[tid] = s;
      }
      __syncthreads();
      for (int i = tid; i < nf * 4 * kCL; i += kT) cluster.map_shared_rank(v_ctx, i / (nf * 4))[f0 * 4 + (i % (nf * 4))] = misc[i % (nf * 4)];
    }
    mark(5);
    csync();
    // ================================================= 6: GRU-2 for the owned units (context-side columns streamed) + ctx / emb part of the logits
    // stage 7's streamed weights (rows k = 4 warp + sub + 32 i of Wy | Wh, two 64-byte segments per row) are requested first
    const int s7_sub = lane >> 3, s7_seg = (lane >> 2) & 1, s7_qd = lane & 3;
    float4 wq_l[8];
    {
      const float* b7 = (s7_seg == 0 ? P + o.Wy : P + o.Wh) + n0 + s7_qd * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) wq_l[i] = ldg4(b7 + (size_t)(warp * 4 + s7_sub + 32 * i) * kMw);
    }
    {
      float4 acc = make4(0.f);
      auto fma4 = [&](const float4& wv, int k) {
        const float xv = v_ctx[k];
        acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
      };
      float4 wq_b[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) wq_b[i] = ldg4(s6_base + (size_t)(s6_k0 + 16 * (16 + i)) * s6_ld);
#pragma unroll
      for (int i = 0; i < 16; ++i) fma4(wq_a[i], s6_k0 + 16 * i);
      float4 wq_c[11];
#pragma unroll
      for (int i = 0; i < 11; ++i) {
        const int k = s6_k0 + 16 * (32 + i);
        wq_c[i] = k < kD ? ldg4(s6_base + (size_t)k * s6_ld) : make4(0.f);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) fma4(wq_b[i], s6_k0 + 16 * (16 + i));
#pragma unroll
      for (int i = 0; i < 11; ++i) { const int k = s6_k0 + 16 * (32 + i); if (k < kD) fma4(wq_c[i], k); }
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
      if (s6_sub == 0) *reinterpret_cast<float4*>(&red[warp * 64 + s6_seg * 16 + s6_qd * 4]) = acc;
    }
    __syncthreads();
    if (tid < 64) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * 64 + tid];
      misc[tid] = s;                                          // [0,16) cr  [16,32) cz  [32,48) ch  [48,64) ctx part of the 16 logit columns
    }
    __syncthreads();
    if (tid < kUPC) {
      const float* ow = own + tid * 4;
      const float r = sigmoidf_(ow[1] + misc[tid]), z = sigmoidf_(ow[2] + misc[16 + tid]);
      const float prop = tanhf(fmaf(ow[3], r, misc[32 + tid]));
      misc[64 + tid] = z * ow[0] + (1.f - z) * prop;          // h_t of the owned unit
    }
    __syncthreads();
    {
      const int dst = tid >> 4, u = tid & 15;
      cluster.map_shared_rank(v_h, dst)[j0 + u] = misc[64 + u];
    }
    mark(6);
    csync();
    // ================================================= 7: emb / h part of the 16 owned logit columns, maxout, partial output logits (all-gather)
    {   // rows k of Wy (x = emb) and Wh (x = h_t), weights already in registers (wq_l)
      const float* xv_ = s7_seg == 0 ? v_emb : v_h;
      float4 acc = make4(0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xv = xv_[warp * 4 + s7_sub + 32 * i];
        acc.x = fmaf(xv, wq_l[i].x, acc.x); acc.y = fmaf(xv, wq_l[i].y, acc.y); acc.z = fmaf(xv, wq_l[i].z, acc.z); acc.w = fmaf(xv, wq_l[i].w, acc.w);
      }
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 8); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 8);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 8); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 8);
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
      if (s7_sub == 0) *reinterpret_cast<float4*>(&red[warp * 64 + s7_seg * 16 + s7_qd * 4]) = acc;
    }
    __syncthreads();
    if (tid < kLPC) {
      float s = misc[48 + tid] + bia[128 + tid];
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * 64 + tid] + red[w8 * 64 + 16 + tid];
      v_sp2[tid] = s;                                          // pre-maxout logit n0 + tid
    }
    __syncthreads();
    if (tid < 112) {                                           // partial output logits of the 8 owned maxout units
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < kLPC / 2; ++u) s = fmaf(fmaxf(v_sp2[2 * u], v_sp2[2 * u + 1]), WoR[u * 112 + tid], s);
      v_sp2[32 + tid] = s;
    }
    __syncthreads();
    for (int i = tid; i < 112 * kCL; i += kT) cluster.map_shared_rank(lpart, i / 112)[c * 112 + (i % 112)] = v_sp2[32 + (i % 112)];
    mark(7);
    csync();
    // ================================================= 8: output logits, arg-max, stop rule (every CTA, redundantly)
    if (warp == 0) {
      float mx = -INFINITY; int am = 0x7fffffff;
      for (int j = lane; j < kVocab; j += 32) {
        float v = bia[144 + j];
#pragma unroll
        for (int r = 0; r < kCL; ++r) v += lpart[r * 112 + j];
        if (v > mx) { mx = v; am = j; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, off);
        const int oa = __shfl_xor_sync(0xffffffffu, am, off);
        if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
      }
      if (lane == 0) reinterpret_cast<int*>(misc)[255] = am;
    }
    __syncthreads();
    y = reinterpret_cast<int*>(misc)[255];
    mark(8);
    if (c == 0 && tid == 0) a.tokens[(size_t)row * a.maxlen + t] = y;
    if ((y == 0 && !a.ignore_eol) || t == a.maxlen - 1) { ++t; break; }
    __syncthreads();
  }
  if (c == 0 && tid == 0 && row == 0 && a.prof != nullptr)
    for (int i = 0; i < 12; ++i) a.prof[i] = pacc[i];
  if (c == 0 && tid == 0) {
    a.lengths[row] = t;
    if (a.steps_run != nullptr) atomicMax(a.steps_run, t);
  }
  cluster.sync();                                             // no CTA exits while a sibling may still address its shared memory
}

}  // namespace

bool cluster_decoder_supports(int h, int w) {
  const int L = h * w, W8 = (w + 7) & ~7;
  return L <= kMaxLc && (h + 10) * (W8 + 10) <= kPadC;
}

// watch [Bn,L,512] | qf [121,512] | h0 [Bn,256] | mean [Bn,684]
size_t greedy_cluster_ws_floats(int Bn, int L) {
  size_t o = 0;
  auto take = [&](size_t n) { o += (n + 3) & ~(size_t)3; };
  take((size_t)Bn * L * kNA); take((size_t)kCovTaps * kNA); take((size_t)Bn * kN); take((size_t)Bn * kD); take(32);
  return o;
}

// 0: ok, 1: this device / driver cannot launch the 16-CTA cluster kernel (caller uses decoder_persist.cu), < 0: error
int greedy_decode_cluster(cudaStream_t st, const float* params, const float* ann, int Bn, int h, int w, int maxlen, int ignore_eol,
                          int* tokens, int* lengths, float* wsp, int* steps_run_dev) {
  const int L = h * w;
  static int state = 0;                                     // 0: unknown, 1: usable, -1: not usable
  const size_t smem = sizeof(float) * (kSmemFloatsC + 32);
  if (state == 0) {
    state = -1;
    if (cudaFuncSetAttribute(greedy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
        cudaFuncSetAttribute(greedy_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, greedy_cluster_kernel, &cfg) == cudaSuccess && nclusters >= 1) state = 1;
    }
    cudaGetLastError();
  }
  if (state != 1 || !cluster_decoder_supports(h, w)) return 1;
  size_t o = 0;
  auto take = [&](size_t n) { float* p = wsp + o; o += (n + 3) & ~(size_t)3; return p; };
  float* watch = take((size_t)Bn * L * kNA); float* qf = take((size_t)kCovTaps * kNA); float* h0 = take((size_t)Bn * kN);
  float* mean = take((size_t)Bn * kD);
  long long* prof = reinterpret_cast<long long*>(take(32));
  CArgs a;
  a.prof = g_cluster_prof ? prof : nullptr;
  a.params = params; a.ann = ann; a.watch = watch; a.qf = cur()->qf_cached != nullptr ? cur()->qf_cached : qf; a.h0 = h0; a.H = h; a.W = w; a.L = L; a.maxlen = maxlen; a.ignore_eol = ignore_eol;
  a.tokens = tokens; a.lengths = lengths; a.steps_run = steps_run_dev;
  const ParamLayout& P = layout();
  COff& d = a.o;
  d.U_f_b = (int)P.U_f_b; d.W_a = (int)P.W_a; d.W_a_b = (int)P.W_a_b; d.V_a = (int)P.V_a; d.V_a_b = (int)P.V_a_b;
  d.W_yz_yr = (int)P.W_yz_yr; d.b_yz_yr = (int)P.b_yz_yr; d.U_hz_hr = (int)P.U_hz_hr; d.W_yh = (int)P.W_yh; d.b_yh = (int)P.b_yh;
  d.U_rh = (int)P.U_rh; d.U_hz_hr_nl = (int)P.U_hz_hr_nl; d.b_hz_hr_nl = (int)P.b_hz_hr_nl; d.W_c_z_r = (int)P.W_c_z_r;
  d.U_rh_nl = (int)P.U_rh_nl; d.b_rh_nl = (int)P.b_rh_nl; d.W_c_h_nl = (int)P.W_c_h_nl; d.embed = (int)P.embed; d.Wc = (int)P.Wc;
  d.bc = (int)P.bc; d.Wh = (int)P.Wh; d.bh = (int)P.bh; d.Wy = (int)P.Wy; d.by = (int)P.by; d.Wo = (int)P.Wo; d.bo = (int)P.bo;
  WAP_TRY(attn_precompute(st, params, ann, nullptr, Bn, h, w, watch, h0, cur()->qf_cached != nullptr ? nullptr : qf, mean));     // hidden_state_0 is unmasked (S:590)
  WAP_CHECK(cudaMemsetAsync(lengths, 0, sizeof(int) * Bn, st));
  if (steps_run_dev != nullptr) WAP_CHECK(cudaMemsetAsync(steps_run_dev, 0, sizeof(int), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(Bn * kCL); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  WAP_CHECK(cudaLaunchKernelEx(&cfg, greedy_cluster_kernel, a));
  WAP_LAUNCH_CHECK();
  return 0;
}

}  // namespace wap
