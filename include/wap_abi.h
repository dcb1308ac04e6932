/* C ABI of the B200-native Watch-Attend-and-Parse hot path (libwap_b200.so).
 *
 * The reference (wwjwhen/Watch-Attend-and-Parse-tensorflow-version) has no FFI: its boundary is
 * the TensorFlow session.run feed/fetch contract of model-single-GPU.py.  Each entry point below
 * names the reference call site it replaces (S = model-single-GPU.py, D = model-distributed.py).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; wap_last_error() gives a thread-local message;
 *   - all tensor arguments are CALLER-OWNED DEVICE pointers (fp32, row-major / NHWC, contiguous,
 *     16-byte aligned); the library allocates no device memory;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); the only calls
 *     that synchronise the host are documented as such (wap_greedy_decode with check_every > 0);
 *   - `params` / `grads` / optimiser state are FLAT fp32 buffers of wap_param_floats() elements whose
 *     layout is described by wap_param_info(); moving batch-norm statistics live in a second flat
 *     buffer of wap_moving_floats() elements;
 *   - dropout uses a counter-based hash RNG keyed by (seed, site, element index); tests reproduce
 *     the masks bit-exactly on the host.
 */
#ifndef WAP_ABI_H_
#define WAP_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAP_ABI_VERSION 1

int wap_abi_version(void);
const char* wap_last_error(void);

/* ---- handles (wap_create / wap_destroy).  A handle owns what the reference keeps per tf.Session (S:609-611):
 * the arithmetic mode of the convolutions, the launch switches (PDL, side-stream overlap), the side streams / events and
 * the device source of the dropout seed.  wap_make_current binds a handle to the CALLING HOST THREAD (like
 * cudaSetDevice): every wap_* call made by that thread afterwards runs under it; NULL rebinds the process-wide default
 * handle, which is also what a thread uses before its first wap_make_current.  One handle per engine; two engines with
 * different precision modes may coexist in one process.  A handle is used by one host thread at a time.
 * There is no wap_repack: weight repacking (K-major hi/lo copies, Q' = Q*U_f) is part of every forward call, from the live
 * weights, so there is no stale packed state to refresh after an optimiser step. */
typedef struct wap_handle wap_handle;
int wap_create(wap_handle** out);
int wap_destroy(wap_handle* h);
int wap_make_current(wap_handle* h);
wap_handle* wap_current(void);   /* NULL = the default handle */

/* ---- parameter layout (variables of S:562-607 under their TF names: conv2d[_k]/kernel,
 *      batch_normalization[_k]/{gamma,beta,moving_mean,moving_variance}, U_f ... bo) */
int wap_param_count(void);
/* flags: bit0 = receives the L2 term (S:596-598); bit1 = moving statistic (offset into the moving buffer) */
int wap_param_info(int index, char* name, int name_cap, int64_t* offset, int* ndim, int64_t dims[4], int* flags);
int64_t wap_param_floats(void);   /* flat trainable buffer (each tensor padded to a multiple of 4 floats) */
int64_t wap_decay_floats(void);   /* elements [0, decay) of the flat buffer get the L2 term */
int64_t wap_moving_floats(void);

/* ---- geometry helpers: 'SAME' sizes of the annotation grid (S:78-86,128-130) */
int wap_annotation_hw(int H, int W, int* h, int* w);
/* a_mask[b,i,j] = x_mask[b,16i,16j]   (S:80,86,130: four [:,0::2,0::2] slices) */
int wap_mask_subsample(void* stream, const float* x_mask, int B, int H, int W, float* a_mask);

/* ---- data.py prepare_data (DATA:100-122) on the device.  `pixels`: the B uint8 images of a batch packed back to back
 * (image b = hw[2b] x hw[2b+1] bytes at byte offset offs[b]); all three arrays in DEVICE memory (one small H2D copy of the
 * raw bytes instead of the padded fp32 tensors).  x [B,Hmax,Wmax,1] = pixel/255. (computed in double, rounded to fp32: the
 * value numpy stores) inside the image and 0 in the padding; x_mask [B,Hmax,Wmax] = 1 inside, 0 in the padding. */
int wap_prepare_batch(void* stream, const uint8_t* pixels, const int64_t* offs, const int32_t* hw, int B, int Hmax, int Wmax,
                      float* x, float* x_mask);

/* ---- Watcher_train.dense_net (S:73-132)
 * x [B,H,W,1]; annotation [B,h,w,684].  training != 0: batch statistics (+ moving-stat update,
 * S:604-607) and activations saved in `workspace` for wap_encoder_backward; dropout != 0 enables
 * the rate-0.2 dropout sites with `seed`. */
int64_t wap_encoder_workspace_floats(int B, int H, int W, int training);
int wap_encoder_forward(void* stream, const float* params, float* moving, const float* x, int B, int H, int W,
                        int training, int dropout, uint32_t seed, float* annotation, float* workspace);
/* gradient of dense_net: grads += d params ; d_annotation is clobbered. */
int wap_encoder_backward(void* stream, const float* params, float* grads, const float* x, int B, int H, int W,
                         int dropout, uint32_t seed, const float* annotation, float* d_annotation, float* workspace);

/* ---- Attender precompute: watch = annotation*U_a + U_a_b (S:177, hoisted out of the scan),
 *      qfused = alpha_past_filter*U_f (S:171-173 fused), h0 = tanh(mean(annotation)*Wa2h + ba2h)
 *      (S:345-347 with a_mask, S:590 with a_mask = NULL).  Any of watch/h0/qfused may be NULL.
 *      scratch: B*684 floats. */
int wap_attn_precompute(void* stream, const float* params, const float* annotation, const float* a_mask,
                        int B, int h, int w, float* watch, float* h0, float* qfused, float* scratch);

/* ---- WAP.get_word (S:379-439): one decode step for Bn hypotheses.
 * y[Bn] int32 (y<0 => zero embedding, S:381-384); shared != 0: all rows share annotation/watch
 * of ONE image (the reference tiles ctx0, S:463).  probs [Bn,111], word [Bn] (argmax), either may be NULL. */
int64_t wap_decoder_step_workspace_floats(int Bn, int h, int w);
int wap_decoder_step(void* stream, const float* params, const float* annotation, const float* watch,
                     const float* qfused, int shared, int Bn, int h, int w, const int32_t* y,
                     const float* h_prev, const float* alpha_past, float* probs, int32_t* word,
                     float* h_new, float* alpha_past_new, float* workspace);

/* ---- Parser.one_time_step (S:260-308), the body of the tf.scan in get_ht_ctx (S:246-255), for a GIVEN embedding
 * emb_y [Bn,256]: carry (h, alpha_past) + masks -> h_new [Bn,256], context [Bn,684], alpha [Bn,h,w], alpha_past'.
 * a_mask [Bn,h,w] / y_mask [Bn] may be NULL.  workspace: wap_decoder_step_workspace_floats(). */
int wap_one_time_step(void* stream, const float* params, const float* annotation, const float* watch, const float* qfused,
                      const float* a_mask, const float* y_mask, int Bn, int h, int w, const float* emb_y, const float* h_prev,
                      const float* alpha_past, float* h_new, float* context, float* alpha, float* alpha_past_new, float* workspace);
/* ---- Attender.get_context (S:166-206): annotation, h_{t-1}, alpha_past, a_mask -> context, alpha, alpha_past' */
int wap_get_context(void* stream, const float* params, const float* annotation, const float* watch, const float* qfused,
                    const float* a_mask, int Bn, int h, int w, const float* h_t_1, const float* alpha_past, float* context,
                    float* alpha, float* alpha_past_new, float* workspace);

/* ---- greedy decode (intended semantics of the dead branch S:475-483) of Bn images at once.
 * tokens [Bn,maxlen] int32, lengths [Bn] int32 (tokens written incl. the final <eol>=0 if emitted).
 * check_every > 0: synchronises the host every that many steps to stop early. */
int64_t wap_greedy_workspace_floats(int Bn, int h, int w);
int wap_greedy_decode(void* stream, const float* params, const float* annotation, int Bn, int h, int w,
                      int maxlen, int check_every, int32_t* tokens, int32_t* lengths, float* workspace,
                      int* steps_run);

/* ---- the same greedy decode with the whole token loop on the device and the stop decided there.  Two kernels behind it:
 * (default, h*w <= 1024) one thread-block cluster of 16 CTAs per image, every per-step exchange through distributed shared memory +
 * cluster barriers, no grid-wide synchronisation, images independent; (larger grids) ONE persistent cooperative kernel with all decoder
 * weights resident in shared memory across the grid and grid-wide barriers between the phases of a step.
 * Bn <= 16, h*w <= 2048.  Never synchronises the host; steps_run_dev (device int, may be NULL) receives
 * the number of steps executed.  ignore_eol != 0 (benchmarking): never stop early, always run maxlen steps. */
int64_t wap_greedy_persistent_workspace_floats(int Bn, int h, int w);
int wap_greedy_decode_persistent(void* stream, const float* params, const float* annotation, int Bn, int h, int w,
                                 int maxlen, int ignore_eol, int32_t* tokens, int32_t* lengths, float* workspace,
                                 int32_t* steps_run_dev);

/* ---- WAP.get_sample(k, maxlen, stochastic=False) (S:442-542) + the length-normalised pick of the validation loop
 * (S:671-672) for ONE image, annotation [1,h,w,684], as one persistent cooperative kernel: the k live hypotheses are the rows of
 * the step (shared annotation, S:463), the candidate ranking / hypothesis bookkeeping (S:485-533) runs on the device.  No host
 * synchronisation and no device->host traffic inside the token loop.  k <= 16, maxlen <= 128, h*w <= 2048.
 * tokens [k,maxlen], lengths [k], scores [k] (accumulated -log p): the samples - finished hypotheses in order of completion,
 * then the rows still alive; meta [4] = {number of samples, index of argmin(score/len), steps executed, live rows at the end}.
 * workspace: wap_greedy_persistent_workspace_floats(k, h, w). */
int wap_beam_decode_persistent(void* stream, const float* params, const float* annotation, int h, int w, int k, int maxlen,
                               int32_t* tokens, int32_t* lengths, float* scores, int32_t* meta, float* workspace);

/* ---- WAP.get_cost (S:337-375) forward over the whole caption (tf.scan of Parser.one_time_step,
 *      S:246-308) and its gradient.  y int32 [T,B], y_mask [T,B], a_mask [B,h,w].
 *      cost (device double) = mean_B sum_T mask*CE  (no L2 term; see wap_adadelta_l2_step).
 *      backward: grads += d params (decoder tensors), d_annotation [B,h,w,684] is written. */
int64_t wap_decoder_train_workspace_floats(int B, int h, int w, int T);
int wap_decoder_train_forward(void* stream, const float* params, const float* annotation, const float* a_mask,
                              const int32_t* y, const float* y_mask, int B, int h, int w, int T,
                              int dropout, uint32_t seed, double* cost, float* workspace);
int wap_decoder_train_backward(void* stream, const float* params, float* grads, const float* annotation,
                               const float* a_mask, const int32_t* y, const float* y_mask, int B, int h, int w,
                               int T, int dropout, uint32_t seed, float* d_annotation, float* workspace);

/* ---- decoder backward + encoder backward in one call (same arguments and results as wap_decoder_train_backward followed
 * by wap_encoder_backward).  The decoder's batched weight gradients run on a side stream next to the encoder backward and
 * rejoin `stream` before the call returns its last launch, so the caller sees ordinary stream semantics. */
int wap_train_backward(void* stream, const float* params, float* grads, const float* x, int B, int H, int W, const float* annotation,
                       const float* a_mask, const int* y, const float* y_mask, int T, int dropout, uint32_t seed, float* d_annotation,
                       float* encoder_workspace, float* decoder_workspace);

/* ---- L2 (S:594-598) + tf.train.AdadeltaOptimizer(lr) (S:602-607; rho .95, eps 1e-8) on the flat
 *      buffers, with the data-parallel 1/world_size mean of the summed replica gradients
 *      (SyncReplicasOptimizer, D:704-717) folded in.  g = grad_scale*grads + 2*l2*p (decayed part).
 *      l2_cost (device double, may be NULL) receives l2 * sum(p^2) of the pre-update parameters. */
int wap_adadelta_l2_step(void* stream, float* params, const float* grads, float* accum, float* accum_update,
                         int64_t n, int64_t n_decay, float lr, float rho, float eps, float l2, float grad_scale,
                         double* l2_cost);

/* ---- instrumentation for bench.py
 * wap_launch_count: kernels launched by this library since load (cumulative).
 * wap_profile_*: when enabled, every GEMM / implicit-GEMM launch is bracketed by cudaEvents on its
 * stream; wap_profile_read sums device time (ms), launches and algorithmic FLOPs (2*M*N*K) of class
 * `cls`: 0 conv1x1 fwd, 1 conv3x3 fwd, 2 transition fwd, 3 conv3x3 wgrad, 4 conv3x3 dgrad,
 * 5 conv1x1 wgrad, 6 conv1x1 dgrad, 7 transition bwd, 8 coverage GEMM, 9 other GEMMs. */
uint64_t wap_launch_count(void);
int wap_profile_enable(int on);
int wap_profile_read(int cls, double* ms, int64_t* launches, double* flops);
/* algorithmic HBM bytes (operands read once + result written/accumulated once) summed over the recorded launches of a class */
int wap_profile_read_bytes(int cls, double* bytes);
int wap_profile_reset(void);

/* ---- fp32 GEMM primitives (exported for the kernel-level parity tests) */
int wap_gemm_nn(void* stream, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                const float* bias, int accumulate);
int wap_gemm_nt(void* stream, const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
                const float* bias, int accumulate);
int wap_gemm_tn_atomic(void* stream, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K);
/* ---- tcgen05 TF32 tensor-core primitives (exported for the kernel-level parity tests)
 * D[M,N] (=|+=) A[M,K](lda) * Bt[N,K](ldb)^T ; optional fp64 column sum / sum-of-squares [2][N].
 * Bt_lo != NULL selects 3xTF32: Bt = rna_tf32(W), Bt_lo = rna_tf32(W - Bt); A is split inside the kernel. */
int wap_tc_available(void);
/* precision of the dense-block / transition convolutions (forward and dgrad) of wap_encoder_*:
 * 0 = fp32 FFMA (CUDA cores), 1 = TF32 tcgen05 tensor cores, single pass with RNA-rounded operands
 * (10-bit mantissa products), 2 = error-compensated 3xTF32 tcgen05 (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,
 * fp32-grade accuracy: the tensor-core parity mode and the DEFAULT of every handle), 3 = bf16 speed mode (bf16 storage of
 * activations and weights, tcgen05 kind::f16, fp32 accumulation and statistics; BASELINE configs[2]).  Per handle. */
int wap_set_conv_precision(int mode);
int wap_get_conv_precision(void);
/* test support: gates[pixel, c] = (batch-norm output of conv layer k > 0) for the training-mode forward whose raw conv outputs and
 * batch statistics are still in `encoder_workspace` (B, H, W as passed to wap_encoder_forward), evaluated with the very expression
 * of the forward kernels.  [B*h_k*w_k, Cout_k] bytes (stem: [B*Hs*Ws, 48] before the max-pool).  The gradient tests hand the gates to
 * the fp64 oracle so that elements within fp32 rounding of the ReLU kink are gated identically on both sides. */
int wap_debug_bn_gates(void* stream, const float* params, const float* encoder_workspace, int B, int H, int W, int k, uint8_t* gates);
/* debugging aid: bit mask of the conv classes allowed on the tensor-core path (default 0x3FF); bit 8 = row-tile form of the 3x3 convs,
 * bit 9 = inference-mode batch norm + ReLU folded into the conv epilogues (off: conv -> bn_act kernels as in training), bit 10 (NOT in the default: it measured
 * 0.7 ms slower per cfg2 step - four epilogue warps per CTA do what a full-occupancy kernel did) = first pass of the 1x1 output's batch-norm
 * backward fused into the epilogue of the 3x3 dgrad, bit 11 (NOT in the default: 0.09 ms slower) = batch-norm backward of the small feature maps as
 * one kernel (both passes around a grid barrier, elements kept in registers) */
int wap_debug_set_tc_mask(int mask);
/* debugging aid: during wap_encoder_backward copy [d a1 | d y2 (ld 32) | d y1] of the dense level whose 1x1 conv index is `layer` into buf */
int wap_debug_capture(int layer, float* buf);
/* debugging aid for the tensor-core wgrad kernel: dump stage-0 shared memory of CTA 0 to buf; xor into the instruction descriptor */
int wap_debug_wgrad(float* buf, int idesc_xor);
int wap_tc_gemm(void* stream, const float* A, int lda, const float* Bt, const float* Bt_lo, int ldb, float* D, int ldd, int M, int N, int K,
                int accumulate, double* stats);
/* the same with the transposed store (D[n*ldd + m]), a bias row and split-K (K spread over grid.z, partial sums ADDED with fp32 atomics into a
 * zero-filled D) - the form the BPTT's coverage-backward GEMM uses */
int wap_tc_gemm_ex(void* stream, const float* A, int lda, const float* Bt, const float* Bt_lo, int ldb, float* D, int ldd, int M, int N, int K,
                   int transpose_out, int splitk, const float* bias);
/* 3x3 SAME conv: x[B,H,W,C] (pixel stride ldx) with Wt[Nrows, 9*C] (K-major, row n = output channel,
 * k = tap*C + ci; rows >= N zero) -> D[B*H*W, N] */
int wap_tc_conv3x3(void* stream, const float* x, int ldx, int B, int H, int W, int C, const float* Wt, const float* Wt_lo, int Nrows, int N,
                   float* D, int ldd, double* stats);
/* 3x3 SAME conv 96 -> 24 channels in "row tile" form (W <= 128): Wt3[96][288] with Wt3[24*j + co][96*dy + ci] = w[dy][j][ci][co]
 * (rows 72..95 zero), Wt3_lo its 3xTF32 residual or NULL; D [B*H*W, 24] (ldd), optional statistics [2][24] as in wap_tc_gemm */
int wap_tc_conv3x3_rowtile(void* stream, const float* x, int ldx, int B, int H, int W, const float* Wt3, const float* Wt3_lo, float* D, int ldd,
                           double* stats);

/* weight gradients on tensor cores (MN-major operands, split over pixels, RED.ADD into dW):
 * 1x1: dW[C,N] (ldo) += X[Mpix,C](ldx)^T dY[Mpix,N](ldy), N <= 96;
 * 3x3: dW[9*C,N] += im2col3x3(X[B,H,W,C])^T dY[B,H,W,32-padded], N <= 32, ldy = 32.  x3 != 0: 3xTF32. */
int wap_tc_wgrad_1x1(void* stream, const float* X, int ldx, const float* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo, int x3);
int wap_tc_wgrad_3x3(void* stream, const float* X, int ldx, const float* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo,
                     int x3);
/* ---- bf16 tensor-core primitives of precision mode 3 (exported for the kernel-level parity tests).  Operands are bf16 in device
 * memory, strides in ELEMENTS (byte strides multiples of 16), fp32 accumulation; out_bf16 != 0 stores D as bf16 (store only).
 *   gemm          : D[M,N] (=|+=) A[M,K](lda) * Bt[N,K](ldb)^T ; optional fp64 column statistics [2][N] of the fp32 accumulators
 *   conv3x3       : x[B,H,W,C] * Wt[Nrows][9*Cpad] (k = tap*Cpad + ci, Cpad = C rounded up to 64, padding zero) -> D[B*H*W, N]
 *   conv3x3_rowtile: 96 -> 24 channels, W <= 128; Wt3[96][384], Wt3[24*j + co][128*dy + ci] = w[dy][j][ci][co]
 *   wgrad_1x1     : dW[C,N] (fp32, ldo) += X[Mpix,C]^T dY[Mpix,N], N <= 128
 *   wgrad_3x3     : dW[9*C,N] (fp32) += im2col3x3(X[B,H,W,C])^T dY[B,H,W,:N] (ldy elements per pixel), C <= 128, N <= 32 */
int wap_tc_gemm_bf16(void* stream, const void* A, int lda, const void* Bt, int ldb, void* D, int ldd, int M, int N, int K, int accumulate, int out_bf16,
                     double* stats);
int wap_tc_conv3x3_bf16(void* stream, const void* x, int ldx, int B, int H, int W, int C, const void* Wt, int Nrows, int N, void* D, int ldd, int out_bf16,
                        double* stats);
int wap_tc_conv3x3_rowtile_bf16(void* stream, const void* x, int ldx, int B, int H, int W, const void* Wt3, void* D, int ldd, int out_bf16, double* stats);
int wap_tc_wgrad_1x1_bf16(void* stream, const void* X, int ldx, const void* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo);
int wap_tc_wgrad_3x3_bf16(void* stream, const void* X, int ldx, const void* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo);

/* ---- dropout seed from device memory: when seed_dev != NULL every dropout site (S:89-93, S:361) reads its seed from
 * *seed_dev at kernel run time instead of the by-value `seed` argument, so that a training step captured once in a CUDA
 * graph can be replayed with a fresh seed per step (tf.layers.dropout draws new masks on every sess.run).  NULL restores
 * the by-value behaviour.  Per handle. */
int wap_set_dropout_seed_source(const uint32_t* seed_dev);

/* ---- decoding with frozen weights: qfused != NULL (a [121,512] buffer filled by wap_attn_precompute) makes wap_greedy_decode_persistent /
 * wap_beam_decode_persistent use it instead of recomputing Q' = alpha_past_filter * U_f (63 MFLOP) at every call - the reference decodes
 * ~1000 validation images one by one with the same weights (S:661-683).  The caller must pass NULL again (or refresh the buffer) before the
 * weights change.  Per handle. */
int wap_set_fused_coverage_filter(const float* qfused);

/* ---- programmatic dependent launch (default on): kernels are launched with the programmatic-stream-serialization
 * attribute so that launch latency overlaps the predecessor's execution; 0 restores plain stream ordering. */
int wap_set_pdl(int on);
int wap_get_pdl(void);
/* ---- debug: CTA (0,0) of every following tcgen05 GEMM launch writes clock64() stamps [k-block][TMA issue, data landed, split done,
 * MMA issued] into trace (device int64, >= 1028 entries), and at [1024..1027] the stamps kernel entry (after the PDL wait), set-up done,
 * accumulator ready (epilogue start) and epilogue done; NULL turns it off */
int wap_debug_tc_trace(long long* trace);
/* ---- debug: the bounded mbarrier waits of the tcgen05 kernels trap instead of hanging; the first call arms a note in host-mapped
 * memory, later calls return them: out[1] = count, then notes of 8 ints from out[8]: {site + 100*k-block, blockIdx.x, blockIdx.y, warp,
 * parity, gridDim.x, gridDim.y, blockDim.x} */
int wap_debug_trap_info(int* out256);
/* ---- debug / A-B: fusion of the training scans (default 3).  bit 0: the forward scan evaluates the GRU gates inside the consuming skinny
 * GEMMs; bit 1: the BPTT does the same for the GRU backward passes and runs dalpha / softmax backward / dz as one kernel (four dependent
 * kernels per time step instead of eight).  0 = every elementwise stage is its own launch. */
int wap_debug_set_fuse_gates(int mask);
/* ---- debug / A-B: which kernel wap_greedy_decode_persistent uses: 0 (default) = one 16-CTA cluster per image, no grid barrier
 * (csrc/decoder_cluster.cu) whenever the annotation grid fits it, 1 = always the grid-wide weight-stationary kernel (csrc/decoder_persist.cu) */
int wap_debug_set_greedy_impl(int impl);
/* ---- weight-gradient kernels on side streams (forked/joined with events inside one call; default on) */
int wap_set_overlap(int on);
int wap_get_overlap(void);

/* keep-mask of the dropout RNG for elements [0,n) of `site` under `seed` (1.0 / 0.0), for tests */
int wap_dropout_mask(void* stream, uint32_t seed, uint32_t site, int64_t n, float* mask);
/* byte offsets (in floats) of selected encoder-workspace regions, for tests: what = 0 stem_raw,
 * 1 block buffer b (arg), 2 y1 of conv k (arg), 3 a1 of conv k, 4 y2 of conv k, 5 raw transition b */
int64_t wap_encoder_workspace_offset(int B, int H, int W, int training, int what, int arg);

#ifdef __cplusplus
}
#endif
#endif /* WAP_ABI_H_ */
