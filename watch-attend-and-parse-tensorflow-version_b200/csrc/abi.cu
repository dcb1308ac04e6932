// extern "C" surface of libwap_b200.so (declared in include/wap_abi.h).
#include <stdarg.h>

#include <new>
#include <vector>

#include "../../include/wap_abi.h"
#include "gemm.cuh"
#include "model.h"
#include "tc_gemm.cuh"

namespace wap {

static thread_local char g_err[1024] = "";
unsigned long long g_launches = 0;
static Context g_default_ctx;
static thread_local Context* t_ctx = nullptr;
Context* cur() { return t_ctx != nullptr ? t_ctx : &g_default_ctx; }
bool pdl_enabled() { return cur()->pdl != 0; }
int upload_wprep_table();
extern long long* g_tc_trace;
extern int g_fuse_gates;
extern int g_greedy_impl;
extern int g_cluster_prof;
const int* trap_note_host();
SideStreams* side_streams() {
  Context* c = cur();
  if (c->ss_state == 0) {
    c->ss_state = 1;
    for (int i = 0; i < 3 && c->ss_state == 1; ++i) if (cudaStreamCreateWithFlags(&c->ss.s[i], cudaStreamNonBlocking) != cudaSuccess) c->ss_state = -1;
    for (int i = 0; i < 4 && c->ss_state == 1; ++i) if (cudaEventCreateWithFlags(&c->ss.fork[i], cudaEventDisableTiming) != cudaSuccess) c->ss_state = -1;
    for (int i = 0; i < 8 && c->ss_state == 1; ++i) if (cudaEventCreateWithFlags(&c->ss.done[i], cudaEventDisableTiming) != cudaSuccess) c->ss_state = -1;
  }
  return c->ss_state == 1 ? &c->ss : nullptr;
}

// ---- per-class device timing
struct ProfRec { cudaEvent_t a, b; int cls; double flops, bytes; };
static double g_prof_next_bytes = 0.0;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_ev_pool;
static cudaEvent_t ev_get() {
  if (!g_ev_pool.empty()) { cudaEvent_t e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
bool overlap_enabled() { return cur()->overlap != 0 && !g_prof_on; }
void prof_set_bytes(double bytes) { g_prof_next_bytes = bytes; }
void prof_begin(cudaStream_t st, int cls, double flops) {
  if (!g_prof_on) { g_prof_next_bytes = 0.0; return; }
  ProfRec r; r.a = ev_get(); r.b = ev_get(); r.cls = cls; r.flops = flops; r.bytes = g_prof_next_bytes; g_prof_next_bytes = 0.0;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st, int cls) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().b, st);
}
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int encoder_forward(cudaStream_t st, const float* params, float* moving, const float* x, int B, int H, int W,
                    int training, int dropout, uint32_t seed, float* annotation, float* wsp);
int encoder_backward(cudaStream_t st, const float* params, float* grads, const float* x, int B, int H, int W,
                     int dropout, uint32_t seed, const float* annotation, float* d_annotation, float* wsp);
int attn_precompute(cudaStream_t st, const float* params, const float* ann, const float* a_mask, int B, int h, int w,
                    float* watch, float* h0, float* qf, float* mean_scratch);
size_t step_ws_floats(int Bn, int L);
size_t greedy_ws_floats(int Bn, int L);
size_t decoder_train_ws_floats(int B, int L, int T);
int decoder_word_step(cudaStream_t st, const float* params, const float* ann, const float* watch, const float* qf,
                      const float* a_mask, int shared, int Bn, int h, int w, const int* y, const float* h_prev,
                      const float* ap_in, float* probs, int* word, float* h_new, float* ap_out, float* wsp);
int decoder_one_time_step(cudaStream_t st, const float* params, const float* ann, const float* watch, const float* qf,
                          const float* a_mask, const float* y_mask, int Bn, int h, int w, const float* emb_y, const float* h_prev,
                          const float* ap_in, float* h_new, float* ctx, float* alpha, float* ap_out, float* wsp);
int attender_get_context(cudaStream_t st, const float* params, const float* ann, const float* watch, const float* qf,
                         const float* a_mask, int Bn, int h, int w, const float* h_t_1, const float* ap_in,
                         float* ctx, float* alpha, float* ap_out, float* wsp);
int greedy_decode(cudaStream_t st, const float* params, const float* ann, int Bn, int h, int w, int maxlen,
                  int check_every, int* tokens, int* lengths, float* wsp, int* steps_run);
int decoder_train_forward(cudaStream_t st, const float* params, const float* ann, const float* a_mask, const int* y,
                          const float* y_mask, int B, int h, int w, int T, int dropout, uint32_t seed,
                          double* cost_out, float* wsp);
int decoder_train_backward(cudaStream_t st, const float* params, float* grads, const float* ann, const float* a_mask,
                           const int* y, const float* y_mask, int B, int h, int w, int T, int dropout, uint32_t seed,
                           float* d_ann, float* wsp, bool defer_join);
int decoder_join(cudaStream_t st);
size_t greedy_persist_ws_floats(int Bn, int L);
int greedy_decode_persistent(cudaStream_t st, const float* params, const float* ann, int Bn, int h, int w, int maxlen, int ignore_eol,
                             int* tokens, int* lengths, float* wsp, int* steps_run_dev);
int beam_decode_persistent(cudaStream_t st, const float* params, const float* ann, int h, int w, int k, int maxlen,
                           int* tokens, int* lengths, float* scores, int* meta, float* wsp);
int adadelta_l2_step(cudaStream_t st, float* p, const float* g, float* acc, float* accu, size_t n, size_t n_decay,
                     float lr, float rho, float eps, float l2, float gscale, double* l2_cost);
void set_conv_precision(int mode);
int get_conv_precision();
extern int g_tc_debug_mask;
extern int g_dbg_layer; extern float* g_dbg_buf;
int debug_bn_gates(cudaStream_t st, const float* params, const float* wsp, int B, int H, int W, int k, uint8_t* out);
int mask_subsample(cudaStream_t st, const float* xm, int B, int H, int W, int h, int w, float* am);
int dropout_mask(cudaStream_t st, uint32_t seed, uint32_t site, size_t n, float* mask);
int prepare_batch(cudaStream_t st, const uint8_t* pix, const long long* offs, const int* hw, int B, int H, int W, float* x, float* xm);

}  // namespace wap

namespace wap { float* g_wg_dbg = nullptr; int g_wg_xor = 0; }
using namespace wap;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int wap_abi_version(void) { return WAP_ABI_VERSION; }
uint64_t wap_launch_count(void) { return g_launches; }
int wap_profile_enable(int on) { g_prof_on = on != 0; return 0; }
int wap_profile_read(int cls, double* ms, int64_t* launches, double* flops) {
  double t = 0.0, f = 0.0; int64_t n = 0;
  for (const ProfRec& r : g_prof) {
    if (r.cls != cls) continue;
    float e = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&e, r.a, r.b) != cudaSuccess) {
      set_error("wap_profile_read: event query failed"); return -2;
    }
    t += e; f += r.flops; ++n;
  }
  if (ms) *ms = t; if (launches) *launches = n; if (flops) *flops = f;
  return 0;
}
int wap_profile_read_bytes(int cls, double* bytes) {
  double b = 0.0;
  for (const ProfRec& r : g_prof) if (r.cls == cls) b += r.bytes;
  if (bytes) *bytes = b;
  return 0;
}
int wap_profile_reset(void) {
  for (const ProfRec& r : g_prof) { g_ev_pool.push_back(r.a); g_ev_pool.push_back(r.b); }
  g_prof.clear();
  return 0;
}
const char* wap_last_error(void) { return g_err; }

int wap_param_count(void) { return (int)layout().tensors.size(); }
int wap_param_info(int index, char* name, int name_cap, int64_t* offset, int* ndim, int64_t dims[4], int* flags) {
  const ParamLayout& L = layout();
  WAP_REQUIRE(index >= 0 && index < (int)L.tensors.size(), "wap_param_info: index %d out of range", index);
  const TensorInfo& t = L.tensors[index];
  if (name != nullptr && name_cap > 0) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (offset) *offset = (int64_t)t.offset;
  if (ndim) *ndim = t.ndim;
  if (dims) for (int i = 0; i < 4; ++i) dims[i] = t.dims[i];
  if (flags) *flags = t.flags;
  return 0;
}
int64_t wap_param_floats(void) { return (int64_t)layout().n_total; }
int64_t wap_decay_floats(void) { return (int64_t)layout().n_decay; }
int64_t wap_moving_floats(void) { return (int64_t)layout().n_moving; }

int wap_annotation_hw(int H, int W, int* h, int* w) {
  WAP_REQUIRE(H > 0 && W > 0, "wap_annotation_hw: bad size %dx%d", H, W);
  const EncGeom g = make_geom(1, H, W);
  if (h) *h = g.h[kBlocks - 1];
  if (w) *w = g.w[kBlocks - 1];
  return 0;
}
int wap_mask_subsample(void* stream, const float* x_mask, int B, int H, int W, float* a_mask) {
  const EncGeom g = make_geom(B, H, W);
  return mask_subsample(ST(stream), x_mask, B, H, W, g.h[kBlocks - 1], g.w[kBlocks - 1], a_mask);
}

int wap_prepare_batch(void* stream, const uint8_t* pixels, const int64_t* offs, const int32_t* hw, int B, int Hmax, int Wmax, float* x, float* x_mask) {
  WAP_REQUIRE(pixels && offs && hw && x && x_mask && B > 0 && Hmax > 0 && Wmax > 0, "wap_prepare_batch: bad arguments");
  return prepare_batch(ST(stream), pixels, reinterpret_cast<const long long*>(offs), hw, B, Hmax, Wmax, x, x_mask);
}

int64_t wap_encoder_workspace_floats(int B, int H, int W, int training) {
  return (int64_t)make_enc_workspace(make_geom(B, H, W), training != 0).total;
}
int wap_encoder_forward(void* stream, const float* params, float* moving, const float* x, int B, int H, int W,
                        int training, int dropout, uint32_t seed, float* annotation, float* workspace) {
  WAP_REQUIRE(params && moving && x && workspace && B > 0 && H > 0 && W > 0, "wap_encoder_forward: bad arguments");
  return encoder_forward(ST(stream), params, moving, x, B, H, W, training, dropout, seed, annotation, workspace);
}
int wap_encoder_backward(void* stream, const float* params, float* grads, const float* x, int B, int H, int W,
                         int dropout, uint32_t seed, const float* annotation, float* d_annotation, float* workspace) {
  WAP_REQUIRE(params && grads && x && d_annotation && workspace, "wap_encoder_backward: bad arguments");
  return encoder_backward(ST(stream), params, grads, x, B, H, W, dropout, seed, annotation, d_annotation, workspace);
}

int wap_attn_precompute(void* stream, const float* params, const float* annotation, const float* a_mask,
                        int B, int h, int w, float* watch, float* h0, float* qfused, float* scratch) {
  WAP_REQUIRE(params && annotation && (h0 == nullptr || scratch != nullptr), "wap_attn_precompute: bad arguments");
  return attn_precompute(ST(stream), params, annotation, a_mask, B, h, w, watch, h0, qfused, scratch);
}

int64_t wap_decoder_step_workspace_floats(int Bn, int h, int w) { return (int64_t)step_ws_floats(Bn, h * w); }
int wap_decoder_step(void* stream, const float* params, const float* annotation, const float* watch,
                     const float* qfused, int shared, int Bn, int h, int w, const int32_t* y,
                     const float* h_prev, const float* alpha_past, float* probs, int32_t* word,
                     float* h_new, float* alpha_past_new, float* workspace) {
  WAP_REQUIRE(params && annotation && watch && qfused && y && h_prev && alpha_past && h_new && alpha_past_new && workspace,
              "wap_decoder_step: bad arguments");
  return decoder_word_step(ST(stream), params, annotation, watch, qfused, nullptr, shared, Bn, h, w, y, h_prev, alpha_past,
                           probs, word, h_new, alpha_past_new, workspace);
}

int wap_one_time_step(void* stream, const float* params, const float* annotation, const float* watch, const float* qfused,
                      const float* a_mask, const float* y_mask, int Bn, int h, int w, const float* emb_y, const float* h_prev,
                      const float* alpha_past, float* h_new, float* context, float* alpha, float* alpha_past_new, float* workspace) {
  WAP_REQUIRE(params && annotation && watch && qfused && emb_y && h_prev && alpha_past && h_new && context && alpha && alpha_past_new && workspace,
              "wap_one_time_step: bad arguments");
  return decoder_one_time_step(ST(stream), params, annotation, watch, qfused, a_mask, y_mask, Bn, h, w, emb_y, h_prev, alpha_past, h_new,
                               context, alpha, alpha_past_new, workspace);
}
int wap_get_context(void* stream, const float* params, const float* annotation, const float* watch, const float* qfused,
                    const float* a_mask, int Bn, int h, int w, const float* h_t_1, const float* alpha_past, float* context,
                    float* alpha, float* alpha_past_new, float* workspace) {
  WAP_REQUIRE(params && annotation && watch && qfused && h_t_1 && alpha_past && context && alpha && alpha_past_new && workspace,
              "wap_get_context: bad arguments");
  return attender_get_context(ST(stream), params, annotation, watch, qfused, a_mask, Bn, h, w, h_t_1, alpha_past, context, alpha,
                              alpha_past_new, workspace);
}

int64_t wap_greedy_workspace_floats(int Bn, int h, int w) { return (int64_t)greedy_ws_floats(Bn, h * w); }
int wap_greedy_decode(void* stream, const float* params, const float* annotation, int Bn, int h, int w,
                      int maxlen, int check_every, int32_t* tokens, int32_t* lengths, float* workspace, int* steps_run) {
  WAP_REQUIRE(params && annotation && tokens && lengths && workspace && maxlen > 0, "wap_greedy_decode: bad arguments");
  return greedy_decode(ST(stream), params, annotation, Bn, h, w, maxlen, check_every, tokens, lengths, workspace, steps_run);
}

int64_t wap_greedy_persistent_workspace_floats(int Bn, int h, int w) { return (int64_t)greedy_persist_ws_floats(Bn, h * w); }
int wap_greedy_decode_persistent(void* stream, const float* params, const float* annotation, int Bn, int h, int w, int maxlen,
                                 int ignore_eol, int32_t* tokens, int32_t* lengths, float* workspace, int32_t* steps_run_dev) {
  WAP_REQUIRE(params && annotation && tokens && lengths && workspace && maxlen > 0, "wap_greedy_decode_persistent: bad arguments");
  return greedy_decode_persistent(ST(stream), params, annotation, Bn, h, w, maxlen, ignore_eol, tokens, lengths, workspace, steps_run_dev);
}

int wap_beam_decode_persistent(void* stream, const float* params, const float* annotation, int h, int w, int k, int maxlen,
                               int32_t* tokens, int32_t* lengths, float* scores, int32_t* meta, float* workspace) {
  WAP_REQUIRE(params && annotation && tokens && lengths && scores && meta && workspace, "wap_beam_decode_persistent: bad arguments");
  return beam_decode_persistent(ST(stream), params, annotation, h, w, k, maxlen, tokens, lengths, scores, meta, workspace);
}

int64_t wap_decoder_train_workspace_floats(int B, int h, int w, int T) { return (int64_t)decoder_train_ws_floats(B, h * w, T); }
int wap_decoder_train_forward(void* stream, const float* params, const float* annotation, const float* a_mask,
                              const int32_t* y, const float* y_mask, int B, int h, int w, int T,
                              int dropout, uint32_t seed, double* cost, float* workspace) {
  WAP_REQUIRE(params && annotation && y && y_mask && workspace && T > 0, "wap_decoder_train_forward: bad arguments");
  return decoder_train_forward(ST(stream), params, annotation, a_mask, y, y_mask, B, h, w, T, dropout, seed, cost, workspace);
}
int wap_decoder_train_backward(void* stream, const float* params, float* grads, const float* annotation,
                               const float* a_mask, const int32_t* y, const float* y_mask, int B, int h, int w,
                               int T, int dropout, uint32_t seed, float* d_annotation, float* workspace) {
  WAP_REQUIRE(params && grads && annotation && y && y_mask && d_annotation && workspace, "wap_decoder_train_backward: bad arguments");
  return decoder_train_backward(ST(stream), params, grads, annotation, a_mask, y, y_mask, B, h, w, T, dropout, seed, d_annotation, workspace, false);
}

int wap_train_backward(void* stream, const float* params, float* grads, const float* x, int B, int H, int W, const float* annotation,
                       const float* a_mask, const int* y, const float* y_mask, int T, int dropout, uint32_t seed, float* d_annotation,
                       float* encoder_workspace, float* decoder_workspace) {
  WAP_REQUIRE(params && grads && x && annotation && y && y_mask && d_annotation && encoder_workspace && decoder_workspace,
              "wap_train_backward: bad arguments");
  int h, w;
  if (wap_annotation_hw(H, W, &h, &w) != 0) return -1;
  int r = decoder_train_backward(ST(stream), params, grads, annotation, a_mask, y, y_mask, B, h, w, T, dropout, seed, d_annotation,
                                 decoder_workspace, true);
  if (r == 0) r = encoder_backward(ST(stream), params, grads, x, B, H, W, dropout, seed, annotation, d_annotation, encoder_workspace);
  const int rj = decoder_join(ST(stream));     // always rejoin the side stream, also on error
  return r != 0 ? r : rj;
}

int wap_adadelta_l2_step(void* stream, float* params, const float* grads, float* accum, float* accum_update,
                         int64_t n, int64_t n_decay, float lr, float rho, float eps, float l2, float grad_scale,
                         double* l2_cost) {
  WAP_REQUIRE(params && grads && accum && accum_update && n > 0, "wap_adadelta_l2_step: bad arguments");
  return adadelta_l2_step(ST(stream), params, grads, accum, accum_update, (size_t)n, (size_t)n_decay, lr, rho, eps, l2, grad_scale, l2_cost);
}

int wap_gemm_nn(void* stream, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                const float* bias, int accumulate) {
  return gemm_nn(ST(stream), A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate != 0);
}
int wap_gemm_nt(void* stream, const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
                const float* bias, int accumulate) {
  return gemm_nt(ST(stream), A, lda, W, ldw, C, ldc, M, N, K, bias, accumulate != 0);
}
int wap_gemm_tn_atomic(void* stream, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K) {
  return gemm_tn_atomic(ST(stream), A, lda, B, ldb, C, ldc, M, N, K);
}
int wap_tc_available(void) { return tc_available() ? 1 : 0; }
int wap_set_conv_precision(int mode) {
  WAP_REQUIRE(mode >= 0 && mode <= 3, "wap_set_conv_precision: mode must be 0 (fp32), 1 (tf32), 2 (3xtf32) or 3 (bf16)");
  WAP_REQUIRE(mode == 0 || tc_available(), "TF32 tensor-core mode needs cuTensorMapEncodeTiled from the driver");
  set_conv_precision(mode);
  return 0;
}
int wap_get_conv_precision(void) { return get_conv_precision(); }
int wap_debug_bn_gates(void* stream, const float* params, const float* encoder_workspace, int B, int H, int W, int k, uint8_t* gates) {
  WAP_REQUIRE(params && encoder_workspace && gates, "wap_debug_bn_gates: bad arguments");
  return debug_bn_gates(ST(stream), params, encoder_workspace, B, H, W, k, gates);
}
int wap_debug_set_tc_mask(int mask) { g_tc_debug_mask = mask; return 0; }
int wap_debug_capture(int layer, float* buf) { g_dbg_layer = layer; g_dbg_buf = buf; return 0; }
int wap_tc_gemm(void* stream, const float* A, int lda, const float* Bt, const float* Bt_lo, int ldb, float* D, int ldd, int M, int N, int K,
                int accumulate, double* stats) {
  TcEpi e{D, ldd, accumulate, stats, nullptr, 0};
  return tc_gemm(ST(stream), A, lda, Bt, Bt_lo, ldb, M, N, K, e, kProfOtherGemm);
}
int wap_tc_gemm_ex(void* stream, const float* A, int lda, const float* Bt, const float* Bt_lo, int ldb, float* D, int ldd, int M, int N, int K,
                   int transpose_out, int splitk, const float* bias) {
  TcEpi e{D, ldd, 0, nullptr, bias, transpose_out};
  return tc_gemm(ST(stream), A, lda, Bt, Bt_lo, ldb, M, N, K, e, kProfOtherGemm, splitk);
}
int wap_tc_conv3x3(void* stream, const float* x, int ldx, int B, int H, int W, int C, const float* Wt, const float* Wt_lo, int Nrows, int N,
                   float* D, int ldd, double* stats) {
  TcEpi e{D, ldd, 0, stats, nullptr, 0};
  return tc_conv3x3(ST(stream), x, ldx, B, H, W, C, Wt, Wt_lo, Nrows, N, e, kProfOtherGemm);
}
int wap_debug_wgrad(float* buf, int idesc_xor) { g_wg_dbg = buf; g_wg_xor = idesc_xor; return 0; }
int wap_tc_conv3x3_rowtile(void* stream, const float* x, int ldx, int B, int H, int W, const float* Wt3, const float* Wt3_lo, float* D, int ldd,
                           double* stats) {
  WAP_REQUIRE(x && Wt3 && D, "wap_tc_conv3x3_rowtile: bad arguments");
  return tc_conv3x3_rowtile(ST(stream), x, ldx, B, H, W, Wt3, Wt3_lo, TcEpi{D, ldd, 0, stats, nullptr, 0}, kProfConv3Fwd);
}
int wap_tc_wgrad_1x1(void* stream, const float* X, int ldx, const float* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo, int x3) {
  return tc_wgrad_1x1(ST(stream), X, ldx, dY, ldy, Mpix, C, N, dW, ldo, x3 != 0, kProfOtherGemm);
}
int wap_tc_wgrad_3x3(void* stream, const float* X, int ldx, const float* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo, int x3) {
  return tc_wgrad_3x3(ST(stream), X, ldx, dY, ldy, B, H, W, C, N, dW, ldo, x3 != 0, kProfOtherGemm);
}
/* ---- bf16 tensor-core primitives (exported for the kernel-level parity tests) */
int wap_tc_gemm_bf16(void* stream, const void* A, int lda, const void* Bt, int ldb, void* D, int ldd, int M, int N, int K, int accumulate, int out_bf16,
                     double* stats) {
  TcEpi e{reinterpret_cast<float*>(D), ldd, accumulate, stats, nullptr, 0, out_bf16};
  return tc_gemm_bf16(ST(stream), A, lda, Bt, ldb, M, N, K, e, kProfOtherGemm);
}
int wap_tc_conv3x3_bf16(void* stream, const void* x, int ldx, int B, int H, int W, int C, const void* Wt, int Nrows, int N, void* D, int ldd, int out_bf16,
                        double* stats) {
  TcEpi e{reinterpret_cast<float*>(D), ldd, 0, stats, nullptr, 0, out_bf16};
  return tc_conv3x3_bf16(ST(stream), x, ldx, B, H, W, C, Wt, Nrows, N, e, kProfOtherGemm);
}
int wap_tc_conv3x3_rowtile_bf16(void* stream, const void* x, int ldx, int B, int H, int W, const void* Wt3, void* D, int ldd, int out_bf16, double* stats) {
  TcEpi e{reinterpret_cast<float*>(D), ldd, 0, stats, nullptr, 0, out_bf16};
  return tc_conv3x3_rowtile_bf16(ST(stream), x, ldx, B, H, W, Wt3, e, kProfOtherGemm);
}
int wap_tc_wgrad_1x1_bf16(void* stream, const void* X, int ldx, const void* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo) {
  return tc_wgrad_1x1_bf16(ST(stream), X, ldx, dY, ldy, Mpix, C, N, dW, ldo, kProfOtherGemm);
}
int wap_tc_wgrad_3x3_bf16(void* stream, const void* X, int ldx, const void* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo) {
  return tc_wgrad_3x3_bf16(ST(stream), X, ldx, dY, ldy, B, H, W, C, N, dW, ldo, kProfOtherGemm);
}
int wap_debug_trap_info(int* out256) {   /* arms the notes on first call; copies them: [1] = count, notes of 8 ints from [8] */
  const int* h = wap::trap_note_host();
  if (h == nullptr) return -1;
  if (out256) for (int i = 0; i < 256; ++i) out256[i] = h[i];
  return 0;
}
int wap_debug_tc_trace(long long* trace) { wap::g_tc_trace = trace; return 0; }
int wap_debug_set_fuse_gates(int mask) { wap::g_fuse_gates = mask & 3; return 0; }
int wap_debug_set_greedy_impl(int impl) { wap::g_greedy_impl = impl & 1; wap::g_cluster_prof = (impl >> 1) & 1; return 0; }
int wap_set_overlap(int on) { cur()->overlap = on ? 1 : 0; return 0; }
int wap_get_overlap(void) { return cur()->overlap; }
int wap_set_pdl(int on) { cur()->pdl = on ? 1 : 0; return 0; }
int wap_get_pdl(void) { return cur()->pdl; }
int wap_set_dropout_seed_source(const uint32_t* seed_dev) { cur()->seed_src = seed_dev; return 0; }
int wap_set_fused_coverage_filter(const float* qfused) { cur()->qf_cached = qfused; return 0; }

/* ---- handles: per-engine state (precision mode, PDL / overlap switches, side streams, dropout-seed source) */
int wap_create(wap_handle** out) {
  WAP_REQUIRE(out != nullptr, "wap_create: out is NULL");
  Context* c = new (std::nothrow) Context();
  WAP_REQUIRE(c != nullptr, "wap_create: out of host memory");
  *out = reinterpret_cast<wap_handle*>(c);
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0) { int n = 0; if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) return upload_wprep_table(); }
  cudaGetLastError();            // no device (host-only use of the geometry / layout functions): nothing to upload
  return 0;
}
int wap_destroy(wap_handle* h) {
  Context* c = reinterpret_cast<Context*>(h);
  if (c == nullptr) return 0;
  WAP_REQUIRE(c != &g_default_ctx, "wap_destroy: not a handle from wap_create");
  if (t_ctx == c) t_ctx = nullptr;
  if (c->ss_state == 1) {
    for (int i = 0; i < 3; ++i) cudaStreamDestroy(c->ss.s[i]);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(c->ss.fork[i]);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(c->ss.done[i]);
  }
  delete c;
  return 0;
}
int wap_make_current(wap_handle* h) { t_ctx = reinterpret_cast<Context*>(h); return 0; }
wap_handle* wap_current(void) { return reinterpret_cast<wap_handle*>(t_ctx); }
int wap_dropout_mask(void* stream, uint32_t seed, uint32_t site, int64_t n, float* mask) {
  return dropout_mask(ST(stream), seed, site, (size_t)n, mask);
}
int64_t wap_encoder_workspace_offset(int B, int H, int W, int training, int what, int arg) {
  const EncWorkspace ws = make_enc_workspace(make_geom(B, H, W), training != 0);
  switch (what) {
    case 0: return (int64_t)ws.stem_raw;
    case 1: return (arg >= 0 && arg < kBlocks) ? (int64_t)ws.buf[arg] : -1;
    case 2: return (arg >= 0 && arg < kNumConv) ? (int64_t)ws.y1[arg] : -1;
    case 3: return (arg >= 0 && arg < kNumConv) ? (int64_t)ws.a1[arg] : -1;
    case 4: return (arg >= 0 && arg < kNumConv) ? (int64_t)ws.y2[arg] : -1;
    case 5: return (arg >= 0 && arg < kBlocks - 1) ? (int64_t)ws.ytr[arg] : -1;
    case 6: return (arg >= 0 && arg < kNumConv) ? (int64_t)ws.wt_fwd[arg] : -1;
    case 7: return (arg >= 0 && arg < kNumConv) ? (int64_t)ws.wt_bwd[arg] : -1;
    case 8: return (arg >= 0 && arg < kNumConv) ? (int64_t)ws.wt_fwd_lo[arg] : -1;
    case 9: return (arg >= 0 && arg < kNumConv) ? (int64_t)ws.wt_bwd_lo[arg] : -1;
    case 10: return (int64_t)ws.total;
  }
  return -1;
}

}  // extern "C"
