// 3x3 convolution + bias + activation for the ANN cells (fp32 CUDA cores): ConvLayer / ConvLayer_ (models/submodules.py:12-83)
// and the three gate convolutions of ConvGRU (models/submodules.py:377-418) including the cat([x, h]) and cat([x, h*reset])
// inputs and the gated blend, so that a ConvGRU step is two launches.
#include "common.cuh"

namespace ef {

constexpr int CA_THREADS = 128, CA_CK = 8, CA_COB = 32, CA_WP = 36;

__device__ __forceinline__ float act_apply(int act, float v) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return 1.0f / (1.0f + expf(-v));
    case 3: return tanhf(v);
    default: return v;
  }
}

// out[b,co,y,x] = blend( act( sum_{ci,tap} in[b,ci,y+dy-1,x+dx-1] * w[co,ci,tap] + bias[co] + residual ) )
// in = cat(x1 [C1], x2 [C2] (* x2_scale)) along channels; blend(o) = h*(1-u) + o*u when blend_h is given.
__global__ void __launch_bounds__(CA_THREADS) conv_ann_fwd_kernel(const ef_conv_ann_params p) {
  __shared__ __align__(16) float s_x[CA_CK * 18 * 18];
  __shared__ __align__(16) float s_w[CA_CK * 9 * CA_WP];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int cblocks = (p.Cout + CA_COB - 1) / CA_COB;
  const int b = blockIdx.z / cblocks, co0 = (blockIdx.z % cblocks) * CA_COB;
  const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
  const int Cin = p.C1 + p.C2, H = p.H, W = p.W;
  const size_t hw = (size_t)H * W;
  float acc[2][CA_COB];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < CA_COB; ++j) acc[i][j] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CA_CK) {
    __syncthreads();
    for (int i = tid; i < CA_CK * 324; i += CA_THREADS) {
      const int ci = ci0 + i / 324, r = i % 324, y = y0 - 1 + r / 18, x = x0 - 1 + r % 18;
      float v = 0.f;
      if (ci < Cin && y >= 0 && y < H && x >= 0 && x < W) {
        const size_t pix = (size_t)y * W + x;
        if (ci < p.C1) {
          v = p.x1[(size_t)b * p.x1_bstride + (size_t)ci * hw + pix];
        } else {
          const int c2 = ci - p.C1;
          v = p.x2[(size_t)b * p.x2_bstride + (size_t)c2 * hw + pix];
          if (p.x2_scale) v *= p.x2_scale[(size_t)b * p.x2_scale_bstride + (size_t)c2 * hw + pix];
        }
      }
      s_x[i] = v;
    }
    for (int i = tid; i < CA_COB * CA_CK * 9; i += CA_THREADS) {
      const int co = i / (CA_CK * 9), r = i % (CA_CK * 9), ci = r / 9;
      float v = 0.f;
      if (co0 + co < p.Cout && ci0 + ci < Cin) v = p.w[((size_t)(co0 + co) * Cin + ci0) * 9 + r];
      s_w[r * CA_WP + co] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < CA_CK; ++ci) {
      const float* sx = s_x + ci * 324;
      float xa[9], xb[9];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          xa[dy * 3 + dx] = sx[(ty + dy) * 18 + tx + dx];
          xb[dy * 3 + dx] = sx[(ty + 8 + dy) * 18 + tx + dx];
        }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + (ci * 9 + tap) * CA_WP);
#pragma unroll
        for (int q = 0; q < CA_COB / 4; ++q) {
          const float4 w4 = wr[q];
          acc[0][4 * q + 0] = fmaf(xa[tap], w4.x, acc[0][4 * q + 0]);
          acc[0][4 * q + 1] = fmaf(xa[tap], w4.y, acc[0][4 * q + 1]);
          acc[0][4 * q + 2] = fmaf(xa[tap], w4.z, acc[0][4 * q + 2]);
          acc[0][4 * q + 3] = fmaf(xa[tap], w4.w, acc[0][4 * q + 3]);
          acc[1][4 * q + 0] = fmaf(xb[tap], w4.x, acc[1][4 * q + 0]);
          acc[1][4 * q + 1] = fmaf(xb[tap], w4.y, acc[1][4 * q + 1]);
          acc[1][4 * q + 2] = fmaf(xb[tap], w4.z, acc[1][4 * q + 2]);
          acc[1][4 * q + 3] = fmaf(xb[tap], w4.w, acc[1][4 * q + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int y = y0 + ty + half * 8, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const size_t pix = (size_t)y * W + x;
#pragma unroll
    for (int co = 0; co < CA_COB; ++co) {
      const int c = co0 + co;
      if (c >= p.Cout) break;
      float v = acc[half][co] + (p.bias ? p.bias[c] : 0.f);
      if (p.residual) v += p.residual[(size_t)b * p.Cout * hw + (size_t)c * hw + pix];
      v = act_apply(p.act, v);
      if (p.blend_h) {
        const float h = p.blend_h[(size_t)b * p.blend_h_bstride + (size_t)c * hw + pix];
        const float u = p.blend_u[(size_t)b * p.blend_u_bstride + (size_t)c * hw + pix];
        v = h * (1.0f - u) + v * u;
      }
      p.out[(size_t)b * p.Cout * hw + (size_t)c * hw + pix] = v;
    }
  }
}

}  // namespace ef

extern "C" int ef_conv_ann_fwd(const ef_conv_ann_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_conv_ann_fwd: params is NULL");
  const ef_conv_ann_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.C1 > 0 && p.C2 >= 0 && p.Cout > 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_conv_ann_fwd: bad dimensions");
  EF_REQUIRE(p.act >= 0 && p.act <= 3, EF_EINVAL, "ef_conv_ann_fwd: bad activation code %d", p.act);
  EF_REQUIRE(p.x1 && p.w && p.out && (p.C2 == 0 || p.x2), EF_ENULL, "ef_conv_ann_fwd: NULL tensor");
  EF_REQUIRE(!p.blend_h == !p.blend_u, EF_ENULL, "ef_conv_ann_fwd: blend needs both h and u");
  dim3 grid(cdiv(p.W, 16), cdiv(p.H, 16), p.B * cdiv(p.Cout, CA_COB));
  conv_ann_fwd_kernel<<<grid, CA_THREADS, 0, as_stream(stream)>>>(p);
  return check_launch("conv_ann_fwd_kernel");
}
