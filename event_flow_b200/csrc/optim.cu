// Fused gradient-norm clipping + Adam over one flat fp32 buffer.
// Reference: torch.nn.utils.clip_grad_norm_ + torch.optim.Adam as driven by train_flow.py:157-163.
#include "common.cuh"

namespace ef {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  __shared__ float s_red[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc = fmaf(g[i], g[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < 8 ? s_red[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) atomicAdd(out, r);
  }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ param, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, const float* __restrict__ sqnorm, float max_norm,
                                                        float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float coef = 1.0f;
  if (max_norm > 0.f) coef = fminf(max_norm / (sqrtf(sqnorm[0]) + 1e-6f), 1.0f);  // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
  const float gi = g[i] * coef;
  const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  param[i] -= (lr / bc1) * (mi / denom);
}

}  // namespace ef

extern "C" int ef_grad_sqnorm(const float* g, int64_t n, float* sqnorm, void* stream) {
  using namespace ef;
  EF_REQUIRE(g && sqnorm && n >= 0, EF_ENULL, "ef_grad_sqnorm: NULL tensor");
  if (n == 0) return EF_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 592) blocks = 592;
  sqnorm_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(g, n, sqnorm);
  return check_launch("sqnorm_kernel");
}

extern "C" int ef_clip_adam(float* param, const float* g, float* m, float* v, int64_t n, const float* sqnorm, float max_norm, float lr,
                            float beta1, float beta2, float eps, int32_t step, void* stream) {
  using namespace ef;
  EF_REQUIRE(param && g && m && v && n >= 0, EF_ENULL, "ef_clip_adam: NULL tensor");
  EF_REQUIRE(max_norm <= 0.f || sqnorm, EF_ENULL, "ef_clip_adam: clipping needs sqnorm");
  EF_REQUIRE(step >= 1, EF_EINVAL, "ef_clip_adam: step counts from 1");
  if (n == 0) return EF_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  clip_adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(param, g, m, v, n, sqnorm, max_norm, lr, beta1, beta2, eps, bc1,
                                                                                bc2_sqrt);
  return check_launch("clip_adam_kernel");
}
