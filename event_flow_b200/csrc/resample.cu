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
