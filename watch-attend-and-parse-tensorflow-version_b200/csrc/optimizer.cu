// L2 regulariser (model-single-GPU.py:594-598) + Adadelta (tf.train.AdadeltaOptimizer, :602-607)
// as ONE multi-tensor kernel over the flat parameter buffer; also the mask subsample and the
// dropout-mask export used by the tests.
#include <algorithm>
#include "common.cuh"

namespace wap {

// TF Adadelta: acc = rho*acc + (1-rho)*g^2 ; upd = sqrt(acc_upd+eps)/sqrt(acc+eps)*g ;
//              acc_upd = rho*acc_upd + (1-rho)*upd^2 ; p -= lr*upd
__global__ void __launch_bounds__(256)
adadelta_l2_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ acc, float* __restrict__ accu,
                   size_t n4, size_t n_decay, float lr, float rho, float eps, float l2, float gscale, double* l2_cost) {
  pdl_sync();
  __shared__ float red[33];
  float sq = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 av = reinterpret_cast<float4*>(acc)[i];
    float4 uv = reinterpret_cast<float4*>(accu)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}; const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float aa[4] = {av.x, av.y, av.z, av.w}; float uu[4] = {uv.x, uv.y, uv.z, uv.w};
    const bool decay = (i * 4) < n_decay;      // tensors are 4-aligned, so a float4 never straddles the boundary
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = gg[j] * gscale;
      if (decay) { gr = fmaf(2.f * l2, pp[j], gr); sq = fmaf(pp[j], pp[j], sq); }
      aa[j] = rho * aa[j] + (1.f - rho) * gr * gr;
      const float upd = sqrtf(uu[j] + eps) * rsqrtf(aa[j] + eps) * gr;
      uu[j] = rho * uu[j] + (1.f - rho) * upd * upd;
      pp[j] -= lr * upd;
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(acc)[i] = make_float4(aa[0], aa[1], aa[2], aa[3]);
    reinterpret_cast<float4*>(accu)[i] = make_float4(uu[0], uu[1], uu[2], uu[3]);
  }
  if (l2_cost != nullptr) {
    sq = block_sum(sq, red);
    if (threadIdx.x == 0) atomicAdd(l2_cost, (double)sq * (double)l2);
  }
}

int adadelta_l2_step(cudaStream_t st, float* p, const float* g, float* acc, float* accu, size_t n, size_t n_decay,
                     float lr, float rho, float eps, float l2, float gscale, double* l2_cost) {
  WAP_REQUIRE((n & 3) == 0 && (n_decay & 3) == 0, "adadelta: n=%zu n_decay=%zu must be multiples of 4", n, n_decay);
  if (l2_cost != nullptr) WAP_CHECK(cudaMemsetAsync(l2_cost, 0, sizeof(double), st));
  WAP_CHECK(launch_pdl(adadelta_l2_kernel, 148 * 4, 256, 0, st, p, g, acc, accu, n / 4, n_decay, lr, rho, eps, l2, gscale, l2_cost));
  WAP_LAUNCH_CHECK();
  return 0;
}

__global__ void mask_subsample_kernel(const float* __restrict__ xm, float* __restrict__ am, int B, int H, int W, int h, int w) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * h * w) return;
  const int x = i % w, t = i / w, y = t % h, b = t / h;
  am[i] = xm[((size_t)b * H + y * 16) * W + x * 16];
}
int mask_subsample(cudaStream_t st, const float* xm, int B, int H, int W, int h, int w, float* am) {
  WAP_CHECK(launch_pdl(mask_subsample_kernel, cdiv(B * h * w, 256), 256, 0, st, xm, am, B, H, W, h, w));
  WAP_LAUNCH_CHECK();
  return 0;
}

// data.py prepare_data (DATA:100-122) on the device: pad to the batch maximum, /255, masks
__global__ void __launch_bounds__(256) prepare_batch_kernel(const uint8_t* __restrict__ pix, const long long* __restrict__ offs, const int* __restrict__ hw,
                                                            int B, int H, int W, float* __restrict__ x, float* __restrict__ xm) {
  pdl_sync();
  const size_t n = (size_t)B * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % W); const size_t t = i / W; const int r = (int)(t % H), b = (int)(t / H);
    const int hb = __ldg(hw + 2 * b), wb = __ldg(hw + 2 * b + 1);
    const bool in = r < hb && c < wb;
    float v = 0.f;
    if (in) v = (float)((double)__ldg(pix + __ldg(offs + b) + (size_t)r * wb + c) / 255.0);
    x[i] = v; xm[i] = in ? 1.f : 0.f;
  }
}
int prepare_batch(cudaStream_t st, const uint8_t* pix, const long long* offs, const int* hw, int B, int H, int W, float* x, float* xm) {
  const size_t n = (size_t)B * H * W;
  WAP_CHECK(launch_pdl(prepare_batch_kernel, (unsigned)std::min<size_t>(cdivz(n, 256), 148 * 8), 256, 0, st, pix, offs, hw, B, H, W, x, xm));
  WAP_LAUNCH_CHECK();
  return 0;
}

__global__ void dropout_mask_kernel(uint32_t seed, uint32_t site, size_t n, float* mask) {
  pdl_sync();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = drop_keep(seed, site, (uint32_t)i) ? 1.f : 0.f;
}
int dropout_mask(cudaStream_t st, uint32_t seed, uint32_t site, size_t n, float* mask) {
  WAP_CHECK(launch_pdl(dropout_mask_kernel, (unsigned)cdivz(n, 256), 256, 0, st, seed, site, n, mask));
  WAP_LAUNCH_CHECK();
  return 0;
}

}  // namespace wap
