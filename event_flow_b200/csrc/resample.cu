// Resampling glue of the U-Net family: 2x bilinear upsampling in front of every decoder cell
// (models/spiking_submodules.py:1010, F.interpolate(scale_factor=2, mode="bilinear", align_corners=False)) and the
// nearest-neighbour upsampling of the multi-resolution flow maps to the input resolution (models/model.py:528-539).
#include "common.cuh"

namespace ef {

// torch's area_pixel_compute_source_index for align_corners=False, scale 1/2: src = (dst + 0.5) / 2 - 0.5, clamped at 0
__device__ __forceinline__ void bilinear2x_src(int d, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float s = ((float)d + 0.5f) * 0.5f - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.0f - l1;
}

__global__ void __launch_bounds__(256) upsample_bilinear2x_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n_planes, int H, int W) {
  const int Wo = 2 * W, Ho = 2 * H;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_planes * Ho * Wo) return;
  const int x = i % Wo, y = (i / Wo) % Ho;
  const size_t pl = i / ((size_t)Wo * Ho);
  int y0, y1, x0, x1;
  float hy0, hy1, hx0, hx1;
  bilinear2x_src(y, H, y0, y1, hy0, hy1);
  bilinear2x_src(x, W, x0, x1, hx0, hx1);
  const float* s = src + pl * H * W;
  // same expression tree as ATen's upsample_bilinear2d: h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d)
  const float top = __fadd_rn(__fmul_rn(hx0, s[(size_t)y0 * W + x0]), __fmul_rn(hx1, s[(size_t)y0 * W + x1]));
  const float bot = __fadd_rn(__fmul_rn(hx0, s[(size_t)y1 * W + x0]), __fmul_rn(hx1, s[(size_t)y1 * W + x1]));
  dst[i] = __fadd_rn(__fmul_rn(hy0, top), __fmul_rn(hy1, bot));
}

// The same upsampling on the internal spike format: bf16 channels-last [B,H,W,C] -> [B,2H,2W,C], one thread = one output pixel x 8
// channels (16-byte loads / stores).  Inputs are spikes / residual sums (small integers): the bilinear weights are multiples of 1/16,
// every result is a multiple of 1/16 below 2^8 and therefore EXACT in bf16 -- the fp32 expression tree above, then a lossless rounding.
__global__ void __launch_bounds__(256) upsample_bilinear2x_cl_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int B, int H, int W,
                                                                     int C) {
  const int Wo = 2 * W, Ho = 2 * H, G = C >> 3;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)B * Ho * Wo * G) return;
  const int g = i % G;
  const size_t pix = i / G;
  const int x = pix % Wo, y = (pix / Wo) % Ho, b = pix / ((size_t)Wo * Ho);
  int y0, y1, x0, x1;
  float hy0, hy1, hx0, hx1;
  bilinear2x_src(y, H, y0, y1, hy0, hy1);
  bilinear2x_src(x, W, x0, x1, hx0, hx1);
  const uint16_t* s = src + (size_t)b * H * W * C + g * 8;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(s + ((size_t)y0 * W + x0) * C)), bq = __ldg(reinterpret_cast<const uint4*>(s + ((size_t)y0 * W + x1) * C));
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(s + ((size_t)y1 * W + x0) * C)), d = __ldg(reinterpret_cast<const uint4*>(s + ((size_t)y1 * W + x1) * C));
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w}, cw[4] = {c.x, c.y, c.z, c.w}, dw[4] = {d.x, d.y, d.z, d.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float r[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float va = h ? bf16_hi(aw[k]) : bf16_lo(aw[k]), vb = h ? bf16_hi(bw[k]) : bf16_lo(bw[k]);
      const float vc = h ? bf16_hi(cw[k]) : bf16_lo(cw[k]), vd = h ? bf16_hi(dw[k]) : bf16_lo(dw[k]);
      const float top = __fadd_rn(__fmul_rn(hx0, va), __fmul_rn(hx1, vb));
      const float bot = __fadd_rn(__fmul_rn(hx0, vc), __fmul_rn(hx1, vd));
      r[h] = __fadd_rn(__fmul_rn(hy0, top), __fmul_rn(hy1, bot));
    }
    o[k] = pack_bf16x2(r[0], r[1]);
  }
  *reinterpret_cast<uint4*>(dst + (((size_t)b * Ho + y) * Wo + x) * C + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256) upsample_nearest_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n_planes, int H, int W,
                                                               int fy, int fx) {
  const int Wo = W * fx, Ho = H * fy;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_planes * Ho * Wo) return;
  const int x = i % Wo, y = (i / Wo) % Ho;
  const size_t pl = i / ((size_t)Wo * Ho);
  dst[i] = src[pl * H * W + (size_t)(y / fy) * W + x / fx];
}

// adjoint of upsample_bilinear2x_kernel in gather form: every source pixel collects from the (at most) 4 x 4 destination pixels
// whose interpolation stencil contains it, with exactly the forward weights (recomputed from the forward index map)
__device__ __forceinline__ float bilinear2x_weight(int d, int n_in, int i) {
  int i0, i1;
  float l0, l1;
  bilinear2x_src(d, n_in, i0, i1, l0, l1);
  return (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
}

__global__ void __launch_bounds__(256) upsample_bilinear2x_bwd_kernel(const float* __restrict__ g_dst, float* __restrict__ g_src, size_t n_planes, int H,
                                                                   int W) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_planes * H * W) return;
  const int x = i % W, y = (i / W) % H;
  const size_t pl = i / ((size_t)W * H);
  const float* g = g_dst + pl * 4 * H * W;
  float acc = 0.f;
  for (int dy = 2 * y - 1; dy <= 2 * y + 2; ++dy) {
    if (dy < 0 || dy >= 2 * H) continue;
    const float wy = bilinear2x_weight(dy, H, y);
    if (wy == 0.f) continue;
    for (int dx = 2 * x - 1; dx <= 2 * x + 2; ++dx) {
      if (dx < 0 || dx >= 2 * W) continue;
      const float wx = bilinear2x_weight(dx, W, x);
      acc += wy * wx * g[(size_t)dy * 2 * W + dx];
    }
  }
  g_src[i] = acc;
}

__global__ void __launch_bounds__(256) upsample_nearest_bwd_kernel(const float* __restrict__ g_dst, float* __restrict__ g_src, size_t n_planes, int H, int W,
                                                                int fy, int fx) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_planes * H * W) return;
  const int x = i % W, y = (i / W) % H;
  const size_t pl = i / ((size_t)W * H);
  const float* g = g_dst + pl * H * W * fy * fx;
  float acc = 0.f;
  for (int a = 0; a < fy; ++a)
    for (int b = 0; b < fx; ++b) acc += g[(size_t)(y * fy + a) * W * fx + x * fx + b];
  g_src[i] = acc;
}

}  // namespace ef

extern "C" int ef_upsample_bilinear2x_bwd(const float* g_dst, float* g_src, int64_t n_planes, int32_t H, int32_t W, void* stream) {
  using namespace ef;
  EF_REQUIRE(g_dst && g_src, EF_ENULL, "ef_upsample_bilinear2x_bwd: NULL tensor");
  EF_REQUIRE(n_planes > 0 && H > 0 && W > 0, EF_EINVAL, "ef_upsample_bilinear2x_bwd: non-positive dimension");
  const size_t n = (size_t)n_planes * H * W;
  upsample_bilinear2x_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(g_dst, g_src, (size_t)n_planes, H, W);
  return check_launch("upsample_bilinear2x_bwd_kernel");
}

extern "C" int ef_upsample_nearest_bwd(const float* g_dst, float* g_src, int64_t n_planes, int32_t H, int32_t W, int32_t fy, int32_t fx, void* stream) {
  using namespace ef;
  EF_REQUIRE(g_dst && g_src, EF_ENULL, "ef_upsample_nearest_bwd: NULL tensor");
  EF_REQUIRE(n_planes > 0 && H > 0 && W > 0 && fy > 0 && fx > 0, EF_EINVAL, "ef_upsample_nearest_bwd: non-positive dimension or factor");
  const size_t n = (size_t)n_planes * H * W;
  upsample_nearest_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(g_dst, g_src, (size_t)n_planes, H, W, fy, fx);
  return check_launch("upsample_nearest_bwd_kernel");
}

extern "C" int ef_upsample_bilinear2x(const float* src, float* dst, int64_t n_planes, int32_t H, int32_t W, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_upsample_bilinear2x: NULL tensor");
  EF_REQUIRE(n_planes > 0 && H > 0 && W > 0, EF_EINVAL, "ef_upsample_bilinear2x: non-positive dimension");
  const size_t n = (size_t)n_planes * H * W * 4;
  upsample_bilinear2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, (size_t)n_planes, H, W);
  return check_launch("upsample_bilinear2x_kernel");
}

extern "C" int ef_upsample_nearest(const float* src, float* dst, int64_t n_planes, int32_t H, int32_t W, int32_t fy, int32_t fx, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_upsample_nearest: NULL tensor");
  EF_REQUIRE(n_planes > 0 && H > 0 && W > 0 && fy > 0 && fx > 0, EF_EINVAL, "ef_upsample_nearest: non-positive dimension or factor");
  const size_t n = (size_t)n_planes * H * W * fy * fx;
  upsample_nearest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, (size_t)n_planes, H, W, fy, fx);
  return check_launch("upsample_nearest_kernel");
}

extern "C" int ef_upsample_bilinear2x_cl(const uint16_t* src, uint16_t* dst, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_upsample_bilinear2x_cl: NULL tensor");
  EF_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, EF_EINVAL, "ef_upsample_bilinear2x_cl: C must be a positive multiple of 8");
  const size_t n = (size_t)B * 4 * H * W * (C / 8);
  upsample_bilinear2x_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, B, H, W, C);
  return check_launch("upsample_bilinear2x_cl_kernel");
}
