// Host-side description of the WAP model: flat parameter layout (TF variable names of the
// reference graph) and per-input-shape encoder geometry.  Pure C++ (no CUDA).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "common.cuh"

namespace wap {

enum ConvKind { kStem = 0, kC1 = 1, kC3 = 2, kTr = 3 };

struct ConvLayer {
  int k;                 // TF index: conv2d_k / batch_normalization_k
  int kind, cin, cout, ks, block;
  size_t w, gamma, beta; // offsets (floats) into the flat trainable buffer
  size_t mmean, mvar;    // offsets (floats) into the moving-statistics buffer
};

struct TensorInfo {
  std::string name;
  size_t offset;         // floats
  int ndim;
  int64_t dims[4];
  int flags;             // bit0: receives L2 (model-single-GPU.py:596-598); bit1: moving statistic (separate buffer)
  size_t numel() const { size_t n = 1; for (int i = 0; i < ndim; ++i) n *= (size_t)dims[i]; return n; }
};

struct ParamLayout {
  ConvLayer conv[kNumConv];
  // Attender (model-single-GPU.py:149-163)
  size_t U_f, U_f_b, U_a, U_a_b, W_a, W_a_b, V_a, V_a_b, Q;
  // Parser (:218-242)
  size_t W_yz_yr, b_yz_yr, U_hz_hr, W_yh, b_yh, U_rh, U_hz_hr_nl, b_hz_hr_nl, W_c_z_r, U_rh_nl, b_rh_nl, W_c_h_nl;
  // WAP (:319-333)
  size_t embed, Wa2h, ba2h, Wc, bc, Wh, bh, Wy, by, Wo, bo;
  size_t n_decay;        // flat elements [0, n_decay) get the L2 term; BN gamma/beta follow
  size_t n_total;        // flat trainable floats (each tensor padded to a multiple of 4)
  size_t n_moving;       // floats in the moving-statistics buffer
  std::vector<TensorInfo> tensors;
};

const ParamLayout& layout();      // singleton, built on first use

// SAME-padding geometry for one (B,H,W) input batch
struct EncGeom {
  int B, H, W;
  int Hs, Ws;            // after the 7x7 s2 stem conv
  int h[kBlocks], w[kBlocks];   // spatial size of dense block b (after maxpool / avgpools)
  int M[kBlocks];        // B*h*w
  int Ctot[kBlocks];     // channels of the block buffer: 432, 600, 684
  int Cin0[kBlocks];     // channels entering the block: 48, 216, 300
};
EncGeom make_geom(int B, int H, int W);

// Encoder workspace (floats unless noted) for one geometry.
struct EncWorkspace {
  size_t stem_raw;                       // [B,Hs,Ws,48]
  size_t buf[kBlocks];                   // block buffers; buf[2] unused when the caller passes `annotation`
  size_t y1[kNumConv], a1[kNumConv], y2[kNumConv];   // indexed by the TF conv index of the 1x1 layer (y1,a1) / 3x3 (y2)
  size_t ytr[kBlocks - 1];               // raw transition outputs
  size_t stats;                          // doubles: [99][2][kMaxBnC]   (byte offset is stats*4, 8-aligned)
  size_t saved;                          // floats:  [99][2][kMaxBnC]   mean, invstd
  // backward scratch
  size_t gbuf[kBlocks - 1];              // d block buffer 0,1
  size_t dy1, da1, dy2, dytr, gstem, dy1b, dy2b;   // dy1b/dy2b: ping-pong partners (weight gradients run on a side stream)
  //   largest-block sized scratch: dy1 is [M,192] (odd level | even level), dy2 has row stride kDy2Ld = 32
  size_t wt_fwd[kNumConv], wt_bwd[kNumConv];   // TF32-rounded K-major weight copies for the tcgen05 path (hi parts)
  size_t wt_fwd_lo[kNumConv], wt_bwd_lo[kNumConv];   // residuals rna(w - hi) for the 3xTF32 mode
  int wt_bwd_ld[kNumConv];               // row stride of the dgrad copy: 1x1 layers of a level pair (2j+1, 2j) share one [Cin][192] buffer
                                         // (odd level in columns 0..95, even level in 96..191) so that ONE K=192 GEMM applies both
  size_t red;                            // doubles [2][kMaxBnC]
  size_t bnbar;                          // 256 unsigned counters: grid barriers of the single-kernel BN backward (one per BN layer, zeroed per backward)
  size_t total;                          // floats
};
EncWorkspace make_enc_workspace(const EncGeom& g, bool training);

// Side streams for work that is off the critical path of a training step (weight gradients): forked from / joined to
// the caller's stream with events, so the pattern is capturable into a CUDA graph.  overlap_enabled() is false while
// per-launch profiling is on (wap_profile_enable) or after wap_set_overlap(0).
struct SideStreams {
  cudaStream_t s[3];          // 0: encoder weight gradients   1: decoder weight gradients   2: second chain of the decoder scans
  cudaEvent_t fork[4];        // recorded on the main stream, waited by a side stream
  cudaEvent_t done[8];        // recorded on a side stream, waited by the main stream
};
// Per-handle state (wap_create / wap_destroy / wap_make_current, include/wap_abi.h): everything that used to be process-global
// in round 1.  Two engines in one process (say a validation engine in fp32 next to a training engine in 3xTF32) each own one.
struct Context {
  int precision = 2;                 // conv arithmetic: 0 fp32 FFMA | 1 TF32 tcgen05 single pass | 2 3xTF32 tcgen05 (default: the tensor-core parity mode) | 3 bf16 tcgen05
  int pdl = 1;                       // programmatic dependent launch attribute on every launch
  int overlap = 1;                   // side streams for work off the critical path
  const uint32_t* seed_src = nullptr;   // device word overriding the by-value dropout seed (CUDA-graph replay draws fresh masks)
  SideStreams ss; int ss_state = 0;  // 0: not created, 1: ok, -1: failed
  bool dec_join_pending = false;     // decoder weight gradients still running on ss.s[1] (wap_train_backward joins them)
  const float* qf_cached = nullptr;  // wap_set_fused_coverage_filter(): precomputed Q' = Q*U_f used by the decode entry points instead of recomputing it per call
};
Context* cur();               // the calling host thread's current context (the process default one unless wap_make_current was called)
SideStreams* side_streams();  // of the current context; lazily created; nullptr on failure
bool overlap_enabled();
inline const uint32_t* seed_src() { return cur()->seed_src; }
}  // namespace wap
