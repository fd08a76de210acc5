// Prediction head: flow = tanh(conv1x1(x) + b), forward and backward.
// Reference: models/submodules.py:52-61 (ConvLayer.forward) as instantiated at models/model.py:197-199.
#include "common.cuh"

namespace ef {

constexpr int PRED_MAX_CIN = 512, PRED_MAX_COUT = 4;  // U-Net prediction layers read up to 8 x base channels
constexpr int PRED_VEC_CIN = 64;                        // widest channels-last row held in registers by the vector path

__device__ __forceinline__ float ld_x(const ef_pred_params& p, int b, int c, size_t pix, size_t hw) {
  if (p.x) return p.x[((size_t)b * p.Cin + c) * hw + pix];
  const uint16_t u = p.x_cl[((size_t)b * hw + pix) * p.Cin + c];
  return __uint_as_float(((uint32_t)u) << 16);
}

__global__ void __launch_bounds__(256) pred_fwd_kernel(const ef_pred_params p) {
  __shared__ float s_w[PRED_MAX_COUT * PRED_MAX_CIN], s_b[PRED_MAX_COUT];
  pdl_launch_dependents();
  pdl_wait();  // the spikes of the last cell
  for (int i = threadIdx.x; i < p.Cout * p.Cin; i += 256) s_w[i] = p.w[i];
  if (threadIdx.x < p.Cout) s_b[threadIdx.x] = p.b[threadIdx.x];
  __syncthreads();
  const size_t hw = (size_t)p.H * p.W;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)p.B * hw) return;
  const int b = i / hw;
  const size_t pix = i % hw;
  float acc[PRED_MAX_COUT];
#pragma unroll
  for (int o = 0; o < PRED_MAX_COUT; ++o) acc[o] = 0.f;
  if (p.x_cl && (p.Cin & 7) == 0 && p.Cin <= PRED_VEC_CIN) {
    // channels-last spikes: the pixel's channels are contiguous -- 16-byte loads, all issued before the first use; the
    // accumulation order over channels is the same as in the scalar loop below
    const uint4* row = reinterpret_cast<const uint4*>(p.x_cl + ((size_t)b * hw + pix) * p.Cin);
    const int nq = p.Cin >> 3;
    uint4 q[PRED_VEC_CIN / 8];
#pragma unroll
    for (int k = 0; k < PRED_VEC_CIN / 8; ++k)
      if (k < nq) q[k] = __ldg(row + k);
#pragma unroll
    for (int k = 0; k < PRED_VEC_CIN / 8; ++k) {
      if (k < nq) {
        const uint32_t u[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xv = (e & 1) ? bf16_hi(u[e >> 1]) : bf16_lo(u[e >> 1]);
#pragma unroll
          for (int o = 0; o < PRED_MAX_COUT; ++o)
            if (o < p.Cout) acc[o] = fmaf(xv, s_w[o * p.Cin + k * 8 + e], acc[o]);
        }
      }
    }
  } else {
    for (int c = 0; c < p.Cin; ++c) {
      const float xv = ld_x(p, b, c, pix, hw);
#pragma unroll
      for (int o = 0; o < PRED_MAX_COUT; ++o)
        if (o < p.Cout) acc[o] = fmaf(xv, s_w[o * p.Cin + c], acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < PRED_MAX_COUT; ++o)
    if (o < p.Cout) p.y[((size_t)b * p.Cout + o) * hw + pix] = tanhf(acc[o] + s_b[o]);
}

// g_pre = g_y (1 - y^2); g_x = W^T g_pre; g_w += sum g_pre x; g_b += sum g_pre
// grid = (pixel blocks of 256 * PB_PIX, groups of 8 input channels): wide prediction layers (256 channels at 32 x 32 in the U-Net's coarse
// scale) spread over the SMs by channel group, every block reduces its partial sums once.
constexpr int PB_PIX = 4;
__global__ void __launch_bounds__(256) pred_bwd_kernel(const ef_pred_params p) {
  __shared__ float s_w[PRED_MAX_COUT * 8];
  __shared__ float s_acc[PRED_MAX_COUT * 8 + PRED_MAX_COUT];
  const int c0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < p.Cout * 8; i += 256) s_w[i] = (c0 + (i & 7) < p.Cin) ? p.w[(i >> 3) * p.Cin + c0 + (i & 7)] : 0.f;
  for (int i = threadIdx.x; i < PRED_MAX_COUT * 8 + PRED_MAX_COUT; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const size_t hw = (size_t)p.H * p.W, n = (size_t)p.B * hw;
  const int lane = threadIdx.x & 31;
  float gw[PRED_MAX_COUT][8], gb[PRED_MAX_COUT];
#pragma unroll
  for (int o = 0; o < PRED_MAX_COUT; ++o) {
    gb[o] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) gw[o][c] = 0.f;
  }
  for (int k = 0; k < PB_PIX; ++k) {
    const size_t i = ((size_t)blockIdx.x * PB_PIX + k) * 256 + threadIdx.x;
    if (i >= n) break;
    const int b = i / hw;
    const size_t pix = i % hw;
    float gp[PRED_MAX_COUT];
#pragma unroll
    for (int o = 0; o < PRED_MAX_COUT; ++o) {
      gp[o] = 0.f;
      if (o < p.Cout) {
        const size_t oo = ((size_t)b * p.Cout + o) * hw + pix;
        const float y = p.y[oo];
        gp[o] = p.g_y[oo] * (1.0f - y * y);
        gb[o] += gp[o];
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c0 + c >= p.Cin) break;
      const float xv = ld_x(p, b, c0 + c, pix, hw);
      float gx = 0.f;
#pragma unroll
      for (int o = 0; o < PRED_MAX_COUT; ++o)
        if (o < p.Cout) {
          gw[o][c] = fmaf(gp[o], xv, gw[o][c]);
          gx = fmaf(gp[o], s_w[o * 8 + c], gx);
        }
      if (p.g_x) p.g_x[((size_t)b * p.Cin + c0 + c) * hw + pix] = gx;
    }
  }
#pragma unroll
  for (int o = 0; o < PRED_MAX_COUT; ++o) {
    if (o >= p.Cout) break;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float r = warp_sum(gw[o][c]);
      if (lane == 0) atomicAdd(&s_acc[o * 8 + c], r);
    }
    if (blockIdx.y == 0) {
      const float r = warp_sum(gb[o]);
      if (lane == 0) atomicAdd(&s_acc[PRED_MAX_COUT * 8 + o], r);
    }
  }
  __syncthreads();
  if (p.g_w)
    for (int i = threadIdx.x; i < p.Cout * 8; i += 256)
      if (c0 + (i & 7) < p.Cin) atomicAdd(p.g_w + (i >> 3) * p.Cin + c0 + (i & 7), s_acc[i]);
  if (blockIdx.y == 0 && threadIdx.x < p.Cout && p.g_b) atomicAdd(p.g_b + threadIdx.x, s_acc[PRED_MAX_COUT * 8 + threadIdx.x]);
}

// Sums of v[0..31] over the 32 lanes of a warp, all 32 entries at once (31 shuffles instead of 32 x 5): step by step the lanes
// trade halves of their arrays (lane bit set: keep the upper half, send the lower one); lane l ends up with the total of entry
// `entry` (a bijection lane -> entry).
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane, int& entry) {
  entry = 0;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const float keep = up ? v[k + half] : v[k];
      const float send = up ? v[k] : v[k + half];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
    entry += up ? half : 0;
  }
  return v[0];
}

// Backward of the FireNet prediction head on the fast formats: 32 channels-last bf16 input channels, 2 outputs.  One pixel per
// thread and iteration (4 x 16-byte loads of the spike row), g_x written fp32 NCHW (one 128-byte line per channel and warp),
// weight-gradient terms kept in registers over the thread's pixels and reduced once with the transposing butterfly.
constexpr int PBC_THREADS = 128, PBC_PIX = 4;
__global__ void __launch_bounds__(PBC_THREADS) pred_bwd_cl32_kernel(const ef_pred_params p) {
  __shared__ float s_w[64], s_acc[66];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 64) s_w[tid] = p.w[tid];
  if (tid < 66) s_acc[tid] = 0.f;
  __syncthreads();
  const size_t hw = (size_t)p.H * p.W, n = (size_t)p.B * hw;
  float gw0[32], gw1[32], gb0 = 0.f, gb1 = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) gw0[c] = gw1[c] = 0.f;
#pragma unroll 1
  for (int k = 0; k < PBC_PIX; ++k) {
    const size_t i = ((size_t)blockIdx.x * PBC_PIX + k) * PBC_THREADS + tid;
    if (i >= n) break;
    const int b = i / hw;
    const size_t pix = i % hw;
    const uint4* row = reinterpret_cast<const uint4*>(p.x_cl + i * 32);
    uint4 q[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) q[g] = __ldg(row + g);
    const size_t o0 = ((size_t)b * 2) * hw + pix;
    const float y0 = p.y[o0], y1 = p.y[o0 + hw];
    const float gp0 = p.g_y[o0] * (1.0f - y0 * y0), gp1 = p.g_y[o0 + hw] * (1.0f - y1 * y1);
    gb0 += gp0, gb1 += gp1;
    float* gx = p.g_x ? p.g_x + (size_t)b * 32 * hw + pix : nullptr;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t u[4] = {q[g].x, q[g].y, q[g].z, q[g].w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = g * 8 + e;
        const float xv = (e & 1) ? bf16_hi(u[e >> 1]) : bf16_lo(u[e >> 1]);
        gw0[c] = fmaf(gp0, xv, gw0[c]);
        gw1[c] = fmaf(gp1, xv, gw1[c]);
        if (gx) gx[(size_t)c * hw] = fmaf(gp1, s_w[32 + c], gp0 * s_w[c]);
      }
    }
  }
  int e0, e1;
  const float t0 = warp_transpose_sum32(gw0, lane, e0);
  const float t1 = warp_transpose_sum32(gw1, lane, e1);
  atomicAdd(&s_acc[e0], t0);
  atomicAdd(&s_acc[32 + e1], t1);
  const float b0 = warp_sum(gb0), b1 = warp_sum(gb1);
  if (lane == 0) atomicAdd(&s_acc[64], b0), atomicAdd(&s_acc[65], b1);
  __syncthreads();
  if (tid < 64 && p.g_w) atomicAdd(p.g_w + tid, s_acc[tid]);
  if (tid < 2 && p.g_b) atomicAdd(p.g_b + tid, s_acc[64 + tid]);
}

static int validate_pred(const ef_pred_params& p, const char* who) {
  EF_REQUIRE(p.B > 0 && p.Cin > 0 && p.Cout > 0 && p.H > 0 && p.W > 0, EF_EINVAL, "%s: bad dimensions", who);
  EF_REQUIRE(p.Cin <= PRED_MAX_CIN && p.Cout <= PRED_MAX_COUT, EF_EUNSUPPORTED, "%s: Cin <= %d, Cout <= %d", who, PRED_MAX_CIN, PRED_MAX_COUT);
  EF_REQUIRE((p.x || p.x_cl) && p.w && p.b && p.y, EF_ENULL, "%s: NULL tensor", who);
  EF_REQUIRE(!p.x_cl || p.Cin % 8 == 0, EF_EINVAL, "%s: cl input needs Cin %% 8 == 0", who);
  return EF_OK;
}

}  // namespace ef

extern "C" int ef_pred_fwd(const ef_pred_params* p, void* stream) {
  using namespace ef;
  EF_REQUIRE(p, EF_ENULL, "ef_pred_fwd: params is NULL");
  if (int rc = validate_pred(*p, "ef_pred_fwd")) return rc;
  const size_t n = (size_t)p->B * p->H * p->W;
  launch_pdl(pred_fwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, as_stream(stream), *p);
  return check_launch("pred_fwd_kernel");
}

extern "C" int ef_pred_bwd(const ef_pred_params* p, void* stream) {
  using namespace ef;
  EF_REQUIRE(p, EF_ENULL, "ef_pred_bwd: params is NULL");
  if (int rc = validate_pred(*p, "ef_pred_bwd")) return rc;
  EF_REQUIRE(p->g_y, EF_ENULL, "ef_pred_bwd: g_y is NULL");
  const size_t n = (size_t)p->B * p->H * p->W;
  if (p->x_cl && p->Cin == 32 && p->Cout == 2) {
    pred_bwd_cl32_kernel<<<(unsigned)((n + PBC_THREADS * PBC_PIX - 1) / (PBC_THREADS * PBC_PIX)), PBC_THREADS, 0, as_stream(stream)>>>(*p);
    return check_launch("pred_bwd_cl32_kernel");
  }
  pred_bwd_kernel<<<dim3((unsigned)((n + 256 * PB_PIX - 1) / (256 * PB_PIX)), (unsigned)((p->Cin + 7) / 8)), 256, 0, as_stream(stream)>>>(*p);
  return check_launch("pred_bwd_kernel");
}
