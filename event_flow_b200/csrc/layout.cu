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

// fp32 NCHW [B,Cin,H,W] with Cin <= 10 -> bf16 channels-last [B,H,W,32] holding the EXACT three-way split of every value:
// slot s (0 = hi, 1 = mid, 2 = lo) of channel c sits at channel s*SL + c (SL = 8 for Cin <= 8, else 10), the rest is zero;
// hi + mid + lo == x bit for bit.  This lets the fractional network input (voxel grids) run through the tensor-core cell kernel
// with fp32-exact products: the head layer's weight image repeats w[.,c] in the three slots of c (ef_split_weights_head).
__global__ void __launch_bounds__(256) pack_split_cl_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int B, int Cin, int SL, size_t hw) {
  // a block owns 256 consecutive pixels = 16 KB of contiguous output: the rows are assembled in shared memory (run-time channel
  // positions are plain shared-memory addresses, not register indices) and leave as fully coalesced 16-byte stores
  __shared__ __align__(16) uint16_t s_row[256 * 32];
  const int tid = threadIdx.x;
  const size_t n = (size_t)B * hw, i0 = (size_t)blockIdx.x * 256, i = i0 + tid;
  uint4* mine = reinterpret_cast<uint4*>(s_row + tid * 32);
#pragma unroll
  for (int k = 0; k < 4; ++k) mine[k] = make_uint4(0, 0, 0, 0);
  if (i < n) {
    const size_t pix = i % hw, b = i / hw;
    for (int c = 0; c < Cin; ++c) {
      const float x = __ldg(src + ((size_t)b * Cin + c) * hw + pix);
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
      uint16_t* o = s_row + tid * 32;
      o[c] = __bfloat16_as_ushort(hi), o[SL + c] = __bfloat16_as_ushort(mid), o[2 * SL + c] = __bfloat16_as_ushort(lo);
    }
  }
  __syncthreads();
  const uint4* s4 = reinterpret_cast<const uint4*>(s_row);
  uint4* d4 = reinterpret_cast<uint4*>(dst) + i0 * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = j * 256 + tid;
    if (i0 + (idx >> 2) < n) d4[idx] = s4[idx];
  }
}

// space-to-depth variants for the stride-2 cells: virtual channel vc = (py*2 + px)*Cin + c of output pixel (Y, X) = input (2Y+py, 2X+px, c)
__global__ void __launch_bounds__(256) pack_split_s2d_cl_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int B, int Cin, int SL, int H, int W) {
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // over B * Ho * Wo
  if (i >= (size_t)B * Ho * Wo) return;
  const int X = i % Wo, Y = (i / Wo) % Ho, b = i / ((size_t)Wo * Ho);
  uint16_t o[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) o[k] = 0;
  for (int par = 0; par < 4; ++par) {
    for (int c = 0; c < Cin; ++c) {
      const float x = __ldg(src + (((size_t)b * Cin + c) * H + 2 * Y + (par >> 1)) * W + 2 * X + (par & 1));
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
      const int vc = par * Cin + c;
      o[vc] = __bfloat16_as_ushort(hi), o[SL + vc] = __bfloat16_as_ushort(mid), o[2 * SL + vc] = __bfloat16_as_ushort(lo);
    }
  }
  uint4* d = reinterpret_cast<uint4*>(dst + i * 32);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    d[k] = make_uint4(o[8 * k] | ((uint32_t)o[8 * k + 1] << 16), o[8 * k + 2] | ((uint32_t)o[8 * k + 3] << 16), o[8 * k + 4] | ((uint32_t)o[8 * k + 5] << 16),
                      o[8 * k + 6] | ((uint32_t)o[8 * k + 7] << 16));
}

__global__ void __launch_bounds__(256) space_to_depth_cl_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int B, int H, int W, int C) {
  const int Ho = H >> 1, Wo = W >> 1, G = C >> 3;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // over B * Ho * Wo * 4 * G (16-byte pieces of the output)
  if (i >= (size_t)B * Ho * Wo * 4 * G) return;
  const int g = i % G, par = (i / G) % 4;
  const size_t pix = i / (4 * G);
  const int X = pix % Wo, Y = (pix / Wo) % Ho, b = pix / ((size_t)Wo * Ho);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (((size_t)b * H + 2 * Y + (par >> 1)) * W + 2 * X + (par & 1)) * C + g * 8));
  *reinterpret_cast<uint4*>(dst + pix * 4 * C + (size_t)par * C + g * 8) = v;
}

}  // namespace ef

extern "C" int ef_pack_split_s2d_cl(const float* src, uint16_t* dst, int32_t B, int32_t Cin, int32_t H, int32_t W, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_pack_split_s2d_cl: NULL tensor");
  EF_REQUIRE(B > 0 && Cin > 0 && 4 * Cin <= EF_HEAD_MAX_CIN && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, EF_EINVAL,
             "ef_pack_split_s2d_cl: 4*Cin <= %d, even H and W", EF_HEAD_MAX_CIN);
  const size_t n = (size_t)B * (H / 2) * (W / 2);
  pack_split_s2d_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, B, Cin, EF_HEAD_SLOT(4 * Cin), H, W);
  return check_launch("pack_split_s2d_cl_kernel");
}

extern "C" int ef_space_to_depth_cl(const uint16_t* src, uint16_t* dst, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_space_to_depth_cl: NULL tensor");
  EF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, EF_EINVAL, "ef_space_to_depth_cl: C %% 8 == 0, even H and W");
  const size_t n = (size_t)B * (H / 2) * (W / 2) * 4 * (C / 8);
  space_to_depth_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, B, H, W, C);
  return check_launch("space_to_depth_cl_kernel");
}

extern "C" int ef_pack_split_cl(const float* src, uint16_t* dst, int32_t B, int32_t Cin, int32_t H, int32_t W, void* stream) {
  using namespace ef;
  EF_REQUIRE(src && dst, EF_ENULL, "ef_pack_split_cl: NULL tensor");
  EF_REQUIRE(B > 0 && Cin > 0 && Cin <= EF_HEAD_MAX_CIN && H > 0 && W > 0, EF_EINVAL, "ef_pack_split_cl: 1 <= Cin <= %d", EF_HEAD_MAX_CIN);
  const size_t n = (size_t)B * H * W;
  pack_split_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, dst, B, Cin, EF_HEAD_SLOT(Cin), (size_t)H * W);
  return check_launch("pack_split_cl_kernel");
}

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
