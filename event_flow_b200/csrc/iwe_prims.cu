// Stand-alone versions of the reference's small public primitives, for callers that use them directly (tools, notebooks):
//   utils/iwe.py:4-17   purge_unfeasible        utils/iwe.py:20-74  get_interpolation        utils/iwe.py:77-92  interpolate
//   models/spiking_util.py:13-109  the spike functions (Heaviside forward, surrogate backward)
// Inside the training / evaluation paths these are fused into ef_iwe_loss_fwd / ef_iwe_image / ef_lif_conv_fwd; the kernels
// here reproduce the reference's intermediate tensors (same shapes, same corner order, same fp32 op order) one launch each.
#include "common.cuh"

namespace ef {

// x [B*M][2] (y, x) -> x * mask, mask [B*M][1]; mask = 0 when either coordinate is outside [0,H) x [0,W)
__global__ void __launch_bounds__(256) purge_kernel(const float2* __restrict__ x, size_t n, float H, float W, float2* __restrict__ x_out,
                                                    float* __restrict__ mask) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float2 v = x[i];
  const bool oob = (v.x < 0.f) || (v.x >= H) || (v.y < 0.f) || (v.y >= W);
  const float m = oob ? 0.f : 1.f;
  x_out[i] = make_float2(__fmul_rn(v.x, m), __fmul_rn(v.y, m));
  mask[i] = m;
}

// events [B][N][4] (ts,y,x,p), flow [B][N][2] (fy,fx) -> idx, weights [B][4N][1] (corner order TL,TR,BL,BR along N) or [B][N][1]
__global__ void __launch_bounds__(256) interpolation_kernel(const ef_iwe_interp_params p) {
  const int i = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (i >= p.N) return;
  const float4 e = reinterpret_cast<const float4*>(p.events)[(size_t)b * p.N + i];
  const float2 f = reinterpret_cast<const float2*>(p.flow)[(size_t)b * p.N + i];
  const float dt = __fsub_rn(p.tref, e.x);
  const float yw = __fadd_rn(e.y, __fmul_rn(__fmul_rn(dt, f.x), p.flow_scaling));
  const float xw = __fadd_rn(e.z, __fmul_rn(__fmul_rn(dt, f.y), p.flow_scaling));
  const float H = (float)p.H, W = (float)p.W;
  if (p.round_idx) {
    const float iy = rintf(yw), ix = rintf(xw);  // torch.round: half to even
    const bool oob = (iy < 0.f) || (iy >= H) || (ix < 0.f) || (ix >= W);
    const float m = oob ? 0.f : 1.f;
    p.idx[(size_t)b * p.N + i] = __fadd_rn(__fmul_rn(__fmul_rn(iy, m), W), __fmul_rn(ix, m));
    p.weights[(size_t)b * p.N + i] = m;
    return;
  }
  const float top = floorf(yw), bot = floorf(__fadd_rn(yw, 1.0f)), left = floorf(xw), right = floorf(__fadd_rn(xw, 1.0f));
  const float iy[4] = {top, top, bot, bot}, ix[4] = {left, right, left, right};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float wy = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(yw, iy[k]))));
    const float wx = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(xw, ix[k]))));
    const bool oob = (iy[k] < 0.f) || (iy[k] >= H) || (ix[k] < 0.f) || (ix[k] >= W);
    const float m = oob ? 0.f : 1.f;
    const size_t o = (size_t)b * 4 * p.N + (size_t)k * p.N + i;
    p.idx[o] = __fadd_rn(__fmul_rn(__fmul_rn(iy[k], m), W), __fmul_rn(ix[k], m));
    p.weights[o] = __fmul_rn(__fmul_rn(wy, wx), m);
  }
}

// iwe [B][HW] += weights (* polarity_mask) at idx
__global__ void __launch_bounds__(256) interpolate_kernel(const float* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ pm,
                                                          int M, int HW, float* __restrict__ iwe) {
  const int i = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (i >= M) return;
  const size_t o = (size_t)b * M + i;
  float v = w[o];
  if (pm) v = __fmul_rn(v, pm[o]);
  const long long k = (long long)idx[o];  // .long(): truncation
  if (k >= 0 && k < HW) atomicAdd(iwe + (size_t)b * HW + k, v);
}

// spike functions: z = (x - thresh > 0); g_x = g * surrogate'(x - thresh)
//   thresh_mode 0: one scalar, 1: per channel ([C], x is [B,C,HW]), 2: same shape as x
__global__ void __launch_bounds__(256) spike_fwd_kernel(const float* __restrict__ x, const float* __restrict__ th, int mode, int C, size_t hw,
                                                        size_t n, float* __restrict__ z) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float t = mode == 0 ? th[0] : (mode == 1 ? th[(i / hw) % C] : th[i]);
  z[i] = (__fsub_rn(x[i], t) > 0.f) ? 1.0f : 0.f;
}

__global__ void __launch_bounds__(256) spike_bwd_kernel(const float* __restrict__ x, const float* __restrict__ th, int mode, int C, size_t hw,
                                                        size_t n, const float* __restrict__ g, int kind, float width, float* __restrict__ g_x) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float t = mode == 0 ? th[0] : (mode == 1 ? th[(i / hw) % C] : th[i]);
  g_x[i] = g[i] * surrogate_grad(kind, __fsub_rn(x[i], t), width);
}

}  // namespace ef

extern "C" int ef_iwe_purge_unfeasible(const float* x, int64_t n, int32_t H, int32_t W, float* x_out, float* mask, void* stream) {
  using namespace ef;
  EF_REQUIRE(n >= 0 && H > 0 && W > 0, EF_EINVAL, "ef_iwe_purge_unfeasible: bad dimensions");
  if (n == 0) return EF_OK;
  EF_REQUIRE(x && x_out && mask, EF_ENULL, "ef_iwe_purge_unfeasible: NULL tensor");
  purge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(x), (size_t)n, (float)H, (float)W,
                                                                            reinterpret_cast<float2*>(x_out), mask);
  return check_launch("purge_kernel");
}

extern "C" int ef_iwe_get_interpolation(const ef_iwe_interp_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_get_interpolation: params is NULL");
  const ef_iwe_interp_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.N >= 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_iwe_get_interpolation: bad dimensions");
  if (p.N == 0) return EF_OK;
  EF_REQUIRE(p.events && p.flow && p.idx && p.weights, EF_ENULL, "ef_iwe_get_interpolation: NULL tensor");
  interpolation_kernel<<<dim3(cdiv(p.N, 256), p.B), 256, 0, as_stream(stream)>>>(p);
  return check_launch("interpolation_kernel");
}

extern "C" int ef_iwe_interpolate(const float* idx, const float* weights, const float* polarity_mask, int32_t B, int32_t M, int32_t H, int32_t W,
                                  float* iwe, void* stream) {
  using namespace ef;
  EF_REQUIRE(B > 0 && M >= 0 && H > 0 && W > 0, EF_EINVAL, "ef_iwe_interpolate: bad dimensions");
  EF_REQUIRE(iwe && (M == 0 || (idx && weights)), EF_ENULL, "ef_iwe_interpolate: NULL tensor");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(iwe, 0, (size_t)B * H * W * sizeof(float), st);
  if (M == 0) return EF_OK;
  interpolate_kernel<<<dim3(cdiv(M, 256), B), 256, 0, st>>>(idx, weights, polarity_mask, M, H * W, iwe);
  return check_launch("interpolate_kernel");
}

extern "C" int ef_spike_fwd(const float* x, const float* thresh, int32_t thresh_mode, int32_t C, int64_t hw, int64_t n, float* z, void* stream) {
  using namespace ef;
  EF_REQUIRE(n >= 0 && thresh_mode >= 0 && thresh_mode <= 2 && (thresh_mode != 1 || (C > 0 && hw > 0)), EF_EINVAL, "ef_spike_fwd: bad arguments");
  if (n == 0) return EF_OK;
  EF_REQUIRE(x && thresh && z, EF_ENULL, "ef_spike_fwd: NULL tensor");
  spike_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, thresh, thresh_mode, C, (size_t)hw, (size_t)n, z);
  return check_launch("spike_fwd_kernel");
}

extern "C" int ef_spike_bwd(const float* x, const float* thresh, int32_t thresh_mode, int32_t C, int64_t hw, int64_t n, const float* g,
                            int32_t surrogate, float width, float* g_x, void* stream) {
  using namespace ef;
  EF_REQUIRE(n >= 0 && thresh_mode >= 0 && thresh_mode <= 2 && (thresh_mode != 1 || (C > 0 && hw > 0)) && surrogate >= 0 && surrogate <= 3, EF_EINVAL,
             "ef_spike_bwd: bad arguments");
  if (n == 0) return EF_OK;
  EF_REQUIRE(x && thresh && g && g_x, EF_ENULL, "ef_spike_bwd: NULL tensor");
  spike_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, thresh, thresh_mode, C, (size_t)hw, (size_t)n, g, surrogate, width, g_x);
  return check_launch("spike_bwd_kernel");
}
