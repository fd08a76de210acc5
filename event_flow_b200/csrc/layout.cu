// fp32 NCHW <-> cl (bf16, channels-last [B, H, W, C]) conversion at the API boundary.
#include "common.cuh"

namespace ef {

__global__ void __launch_bounds__(256) pack_cl_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int B, int C, size_t hw) {
  // one thread = 8 channels of one pixel; consecutive threads take consecutive pixels (coalesced fp32 reads per channel)
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // over B * C/8 * hw, pixel fastest
  if (i >= (size_t)B * (C >> 3) * hw) return;
  const size_t pix = i % hw, bg = i / hw;
  const int g = bg % (C >> 3), b = bg / (C >> 3);
  const float* s = src + ((size_t)b * C + g * 8) * hw + pix;
  uint32_t u[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) u[k] = pack_bf16x2(s[(2 * k) * hw], s[(2 * k + 1) * hw]);
  *reinterpret_cast<uint4*>(dst + ((size_t)b * hw + pix) * C + g * 8) = make_uint4(u[0], u[1], u[2], u[3]);
}

__global__ void __launch_bounds__(256) unpack_cl_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, int B, int C, size_t hw) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)B * (C >> 3) * hw) return;
  const size_t pix = i % hw, bg = i / hw;
  const int g = bg % (C >> 3), b = bg / (C >> 3);
  const uint4 u = *reinterpret_cast<const uint4*>(src + ((size_t)b * hw + pix) * C + g * 8);
  float* d = dst + ((size_t)b * C + g * 8) * hw + pix;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    d[(2 * k) * hw] = bf16_lo(w[k]);
    d[(2 * k + 1) * hw] = bf16_hi(w[k]);
  }
}

}  // namespace ef

extern "C" int ef_pack_cl(const float* src, uint16_t* dst, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_pack_cl: NULL tensor");
  EF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, EF_EINVAL, "ef_pack_cl: C must be a positive multiple of 8");
  const size_t n = (size_t)B * (C / 8) * H * W;
  pack_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, B, C, (size_t)H * W);
  return check_launch("pack_cl_kernel");
}

extern "C" int ef_unpack_cl(const uint16_t* src, float* dst, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_unpack_cl: NULL tensor");
  EF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, EF_EINVAL, "ef_unpack_cl: C must be a positive multiple of 8");
  const size_t n = (size_t)B * (C / 8) * H * W;
  unpack_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, B, C, (size_t)H * W);
  return check_launch("unpack_cl_kernel");
}
