// tcgen05 (5th-gen tensor core) TF32 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = A[M,K] * Bt[N,K]^T       fp32 storage, TF32 multiply, fp32 accumulate in TMEM
//
// Both operands are K-major in shared memory with the 128-byte swizzle: one pipeline stage is a
// [128 rows x 32 floats] A tile and a [BN rows x 32 floats] B tile, each row exactly one 128 B
// swizzle span, written by TMA (cp.async.bulk.tensor, SWIZZLE_128B) and consumed by
// tcgen05.mma.kind::tf32 (M=128, N=BN, K=8: four instructions per stage).  The A operand is either
// a plain 2-D row-major matrix (row stride lda: 1x1 convolutions read a channel slice of the
// dense-block buffer in place) or a 3x3 SAME im2col view of an NHWC tensor fetched as 4-D TMA
// boxes [32 ch, TW, TH, 1] whose out-of-bounds halo is zero-filled by the TMA unit itself.
// K tails (C not a multiple of 32) are zero-filled the same way, so no padded copies exist.
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2..5 = epilogue (tcgen05.ld 32x32b, one TMEM lane = one
// output row per thread): store / accumulate, and optional per-column sum and sum of squares
// (batch-norm statistics) via a 31-shuffle transposing reduction per 32 columns.
// Single-pass TF32: two CTAs are resident per SM so one tile's epilogue overlaps the other's main loop.
// 3xTF32 (X3, the parity mode): the epilogue warps double as operand splitters in the main loop - they read the raw fp32 A
// tile from shared memory, write a_hi / a_lo into a ring in TENSOR MEMORY (tcgen05.st) and the MMAs take A from there
// (TS form), B hi/lo tiles come pre-split from global memory; a_hi*b_hi accumulates in one TMEM accumulator, the two
// correction products in a second one; one CTA per SM (TMEM: 2 accumulators + 4 x 64 ring columns = 512).
// IM2COL == 2 is the "row tile" 3x3 conv: vertical taps in K, horizontal taps as 3 x 24 accumulator columns summed with a
// pixel shift in the epilogue (tc_conv3x3_rowtile).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace wap {

struct TcEpi {
  float* out; int ld;        // D[m*ld + n]
  int accumulate;            // D += acc
  double* stats;             // [2][N] column sum / sum of squares of the raw accumulators, or null
  const float* bias;         // optional [N] added to the stored value (not to the statistics)
  int transpose_out;         // store D[m,n] at out[n*ld + m] (coalesced: one TMEM lane = one row m per thread)
  int out_bf16;              // `out` points to bf16 elements (ld in elements); store only (no accumulate / transpose / bias)
  const float* bn_scale;     // inference-mode batch norm folded into the epilogue: D = relu(acc * bn_scale[n] + bn_shift[n]) (both [N]); with
  const float* bn_shift;     //   statistics == nullptr, no accumulate / transpose / bias
  // Fused first pass of the batch-norm backward of the layer that PRODUCED this GEMM's "input gradient" (the 3x3 dgrad writes d a1, the
  // gradient wrt the activated output of the 1x1 conv's BN): with bnb_raw != nullptr the epilogue turns the accumulator u = d a1 into
  // g = u * [BN output > 0] * dropout, stores g (instead of u) and adds the column sums of g and g * xhat into `stats` ([2][N] doubles) -
  // exactly what bn_bwd_reduce_kernel<0> does, without the separate pass over d a1 and the raw output.  N must be <= 96, one column tile.
  const float* bnb_raw; int bnb_ld;      // raw conv output [M, bnb_ld] of that layer (fp32)
  const float* bnb_saved;                // its saved batch mean [kMaxBnC] | inverse std [kMaxBnC]
  const float* bnb_gamma; const float* bnb_beta;
  int bnb_dropout; uint32_t bnb_seed, bnb_site; const uint32_t* bnb_seed_src;
};

// Bt_hi/lo[N][K] = split(transpose(W[K][N])) (transpose != 0) or split(W[N][K]) : operand preparation for tc_gemm's 3xTF32 mode
int tc_prep_bt(cudaStream_t st, const float* W, int K, int N, int transpose, float* hi, float* lo);
// the same for up to 8 matrices in one launch (the decoder head's weights)
struct TcPrepEntry { const float* W; int K, N, transpose; float* hi; float* lo; };
struct TcPrepBatch { TcPrepEntry e[8]; int n; };
int tc_prep_multi(cudaStream_t st, const TcPrepBatch& b);

// D[M,N] = A[M,K](lda) * Bt[N,K](ldb)^T ; N <= 256 per launch column tile (grid.y tiles of BN)
// Bt_lo != null selects the error-compensated 3xTF32 variant: Bt = rna_tf32(W), Bt_lo = rna_tf32(W - Bt); A is split on the fly.
// splitk > 1: the K reduction is spread over grid.z and the partial sums are ADDED into `out` with fp32 atomics (the caller zero-fills it)
int tc_gemm(cudaStream_t st, const float* A, int lda, const float* Bt, const float* Bt_lo, int ldb, int M, int N, int K, const TcEpi& epi, int prof_cls,
            int splitk = 1);
// 3x3 SAME conv: x[B,H,W,C] (pixel stride ldx) * Wt[Nrows, 9*C]^T -> D[B*H*W, N]  (Nrows >= N rows allocated, extra rows zero)
int tc_conv3x3(cudaStream_t st, const float* x, int ldx, int B, int H, int W, int C, const float* Wt, const float* Wt_lo, int Nrows, int N, const TcEpi& epi, int prof_cls);
// weight gradients (reduction over pixels, both operands MN-major): dW += X^T dY ; x3 selects 3xTF32 (both operands split on the fly)
int tc_conv3x3_rowtile(cudaStream_t st, const float* x, int ldx, int B, int H, int W, const float* Wt3, const float* Wt3_lo, const TcEpi& epi, int prof_cls);
int tc_wgrad_1x1(cudaStream_t st, const float* X, int ldx, const float* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo, bool x3, int prof_cls);
int tc_wgrad_3x3(cudaStream_t st, const float* X, int ldx, const float* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo, bool x3,
                 int prof_cls);
// bf16 variants (precision mode 3): bf16 operands in global memory (strides in elements), fp32 accumulation; see tc_gemm.cu
int tc_gemm_bf16(cudaStream_t st, const void* A, int lda, const void* Bt, int ldb, int M, int N, int K, const TcEpi& epi, int prof_cls);
int tc_conv3x3_bf16(cudaStream_t st, const void* x, int ldx, int B, int H, int W, int C, const void* Wt, int Nrows, int N, const TcEpi& epi, int prof_cls);
int tc_conv3x3_rowtile_bf16(cudaStream_t st, const void* x, int ldx, int B, int H, int W, const void* Wt3, const TcEpi& epi, int prof_cls);
int tc_wgrad_1x1_bf16(cudaStream_t st, const void* X, int ldx, const void* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo, int prof_cls);
int tc_wgrad_3x3_bf16(cudaStream_t st, const void* X, int ldx, const void* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo, int prof_cls);
bool tc_available();

}  // namespace wap
