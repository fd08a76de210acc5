// Backward of one fused conv + spiking-neuron cell-step (BPTT building block), fp32 CUDA-core kernels.
// Implements the recurrences of SURVEY.md 8a, i.e. what torch.autograd derives from
// models/spiking_submodules.py:96-126 (+PLIF/ALIF/XLIF and recurrent siblings) with the surrogate derivatives of
// models/spiking_util.py:39-93.  Three launches: (1) pointwise neuron backward + per-channel parameter gradients,
// (2) data gradient (transposed 3x3 conv of g_I), (3) weight gradient (correlation of inputs with g_I).
#include "common.cuh"

namespace ef {

// ---------------------------------------------------------------------------------------------------------------
// (1) pointwise.  One block = 1024 consecutive pixels of one (b, c) plane; per-channel sums by block reduction.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PW_THREADS = 256, PW_PER_THREAD = 4;

// activation read in either layout: fp32 NCHW or cl (bf16 channels-last [B, H, W, C])
__device__ __forceinline__ float ld_act2(const float* f32, const uint16_t* cl, int b, int c, size_t pix, int C, size_t hw) {
  if (f32) return f32[((size_t)b * C + c) * hw + pix];
  const uint16_t u = cl[((size_t)b * hw + pix) * C + c];
  return __uint_as_float(((uint32_t)u) << 16);
}

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    r = lane < (PW_THREADS / 32) ? s_red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in thread 0
}

template <int NEURON, bool HARD>
__global__ void __launch_bounds__(PW_THREADS) lif_bwd_pointwise_kernel(const ef_lif_conv_bwd_params q, int Ho, int Wo,
                                                                       float* __restrict__ gP_sum) {
  __shared__ float s_red[PW_THREADS / 32];
  const ef_lif_conv_params& p = q.f;
  const int c = blockIdx.y, b = blockIdx.z;
  const size_t plane = (size_t)Ho * Wo;
  const size_t base = ((size_t)b * p.C + c) * plane;
  const ChanConst k = load_chan_const(p, c);
  const float oml = 1.0f - k.lam;

  float s_lam = 0.f, s_thr = 0.f, s_rho = 0.f, s_alpha = 0.f, s_t0 = 0.f, s_t1 = 0.f;
  // a thread owns PW_PER_THREAD consecutive pixels: every tensor is read with ONE 16-byte load per thread (planes whose size is a
  // multiple of 4 -- else element by element), all loads issued before the arithmetic, results leave as 16-byte stores
  const size_t pix0 = ((size_t)blockIdx.x * PW_THREADS + threadIdx.x) * PW_PER_THREAD;
  static_assert(PW_PER_THREAD == 4, "one float4 per tensor and thread");
  const uintptr_t all_ptrs = (uintptr_t)p.v_in | (uintptr_t)p.z_in | (uintptr_t)p.aux_in | (uintptr_t)p.v_out | (uintptr_t)p.aux_out | (uintptr_t)q.g_out |
                             (uintptr_t)q.g_z_out | (uintptr_t)q.g_v_out | (uintptr_t)q.g_aux_out | (uintptr_t)q.scratch_gI | (uintptr_t)q.g_v_in |
                             (uintptr_t)q.g_z_in | (uintptr_t)q.g_aux_in;
  const bool vec = (plane & 3) == 0 && (all_ptrs & 15) == 0;
  const size_t o0 = base + pix0;
  auto ld4 = [&](const float* __restrict__ ptr, float (&d)[4]) {
    if (!ptr || pix0 >= plane) {
      d[0] = d[1] = d[2] = d[3] = 0.f;
    } else if (vec) {
      const float4 t = *reinterpret_cast<const float4*>(ptr + o0);
      d[0] = t.x, d[1] = t.y, d[2] = t.z, d[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = (pix0 + j < plane) ? ptr[o0 + j] : 0.f;
    }
  };
  auto st4 = [&](float* __restrict__ ptr, const float (&d)[4]) {
    if (!ptr || pix0 >= plane) return;
    if (vec) {
      *reinterpret_cast<float4*>(ptr + o0) = make_float4(d[0], d[1], d[2], d[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (pix0 + j < plane) ptr[o0 + j] = d[j];
    }
  };
  float v_p4[4], z_p4[4], a_p4[4], v_n4[4], a_n4[4], g_o4[4], g_zo4[4], g_vo4[4], g_ao4[4];
  ld4(p.v_in, v_p4);
  if (p.z_in) {
    ld4(p.z_in, z_p4);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) z_p4[j] = (p.z_in_cl && pix0 + j < plane) ? ld_act2(p.z_in, p.z_in_cl, b, c, pix0 + j, p.C, plane) : 0.f;
  }
  ld4(NEURON != EF_LIF ? p.aux_in : nullptr, a_p4);
  ld4(p.v_out, v_n4);
  ld4(NEURON != EF_LIF ? p.aux_out : nullptr, a_n4);
  ld4(q.g_out, g_o4);
  ld4(q.g_z_out, g_zo4);
  ld4(q.g_v_out, g_vo4);
  ld4(NEURON != EF_LIF ? q.g_aux_out : nullptr, g_ao4);
  float gI4[4], g_vi4[4], g_zi4[4], g_ai4[4];
  const float inv_oml = 1.0f / oml, inv_omr = (NEURON == EF_PLIF || NEURON == EF_XLIF) ? 1.0f / (1.0f - k.rho) : 0.f;
#pragma unroll
  for (int i = 0; i < PW_PER_THREAD; ++i) {
    const size_t pix = pix0 + i;
    const bool live = pix < plane;
    const float v_p = v_p4[i], z_p = z_p4[i], a_p = a_p4[i], v_n = v_n4[i], a_n = a_n4[i];
    const float thr_t = (NEURON == EF_LIF || NEURON == EF_PLIF) ? k.thr : (k.t0 + k.t1 * a_n);
    const float g_z = g_o4[i] + g_zo4[i];
    const float sg = live ? surrogate_grad(p.surrogate, v_n - thr_t, p.act_width) : 0.f;
    const float g_thr = -g_z * sg;  // dL/d thresh_t
    const float g_v = g_vo4[i] + g_z * sg;
    gI4[i] = oml * g_v;
    const float g_aux_n = g_ao4[i];

    // what the (1-lam) factor multiplied in the forward ("drive"), recovered from v_out
    const float reset_thr = (NEURON == EF_LIF || NEURON == EF_PLIF) ? k.thr : (k.t0 + k.t1 * a_p);
    const float keep = HARD ? v_p * (1.0f - z_p) : v_p;
    const float drive = HARD ? (v_n - k.lam * keep) * inv_oml : (v_n - k.lam * v_p + z_p * reset_thr) * inv_oml;
    s_lam += g_v * (keep - drive);
    g_vi4[i] = HARD ? g_v * k.lam * (1.0f - z_p) : g_v * k.lam;

    float g_z_direct = 0.f, g_aux_p = 0.f;
    if (NEURON == EF_LIF) {
      s_thr += g_thr - (HARD ? 0.f : z_p * g_v);
    } else if (NEURON == EF_PLIF) {
      s_thr += g_thr - (HARD ? 0.f : z_p * g_v);
      const float g_pt = g_aux_n - oml * k.alpha * g_v;
      s_alpha += -oml * a_n * g_v;
      const float P = (a_n - k.rho * a_p) * inv_omr;
      s_rho += g_pt * (a_p - P);
      g_aux_p = k.rho * g_pt;
      if (gP_sum && live) atomicAdd(gP_sum + (size_t)b * plane + pix, (1.0f - k.rho) * g_pt);
    } else if (NEURON == EF_ALIF) {
      const float g_a = g_aux_n + k.t1 * g_thr;
      s_t0 += g_thr - (HARD ? 0.f : z_p * g_v);
      s_t1 += g_thr * a_n - (HARD ? 0.f : z_p * a_p * g_v);
      s_rho += g_a * (a_p - z_p);
      g_z_direct = (1.0f - k.rho) * g_a;
      g_aux_p = k.rho * g_a - (HARD ? 0.f : z_p * k.t1 * g_v);
    } else {  // XLIF
      const float g_pt = g_aux_n + k.t1 * g_thr;
      s_t0 += g_thr - (HARD ? 0.f : z_p * g_v);
      s_t1 += g_thr * a_n - (HARD ? 0.f : z_p * a_p * g_v);
      const float P = (a_n - k.rho * a_p) * inv_omr;
      s_rho += g_pt * (a_p - P);
      g_aux_p = k.rho * g_pt - (HARD ? 0.f : z_p * k.t1 * g_v);
      if (gP_sum && live) atomicAdd(gP_sum + (size_t)b * plane + pix, (1.0f - k.rho) * g_pt);
    }
    if (q.reset_grad)  // differentiable reset (detach=False): the previous spikes also act through the reset term of v_out
      g_z_direct -= HARD ? g_v * k.lam * v_p : g_v * reset_thr;
    g_zi4[i] = g_z_direct;  // the recurrent dgrad (launch 2) accumulates on top
    g_ai4[i] = g_aux_p;
  }
  st4(q.scratch_gI, gI4);
  st4(q.g_v_in, g_vi4);
  st4(q.g_z_in, g_zi4);
  if (NEURON != EF_LIF) st4(q.g_aux_in, g_ai4);

  // per-channel raw-parameter gradients: chain through sigmoid / clamp_min
  float r;
  r = block_sum(s_lam, s_red);
  if (threadIdx.x == 0 && q.g_leak) atomicAdd(q.g_leak + c, r * k.lam * (1.0f - k.lam));
  if (NEURON == EF_LIF || NEURON == EF_PLIF) {
    r = block_sum(s_thr, s_red);
    if (threadIdx.x == 0 && q.g_thresh && p.thresh[c] >= 0.01f) atomicAdd(q.g_thresh + c, r);
  }
  if (NEURON != EF_LIF) {
    r = block_sum(s_rho, s_red);
    if (threadIdx.x == 0 && q.g_leak_aux) atomicAdd(q.g_leak_aux + c, r * k.rho * (1.0f - k.rho));
  }
  if (NEURON == EF_PLIF) {
    r = block_sum(s_alpha, s_red);
    if (threadIdx.x == 0 && q.g_add_pt) atomicAdd(q.g_add_pt + c, r * k.alpha * (1.0f - k.alpha));
  }
  if (NEURON == EF_ALIF || NEURON == EF_XLIF) {
    r = block_sum(s_t0, s_red);
    if (threadIdx.x == 0 && q.g_t0 && p.t0[c] >= 0.01f) atomicAdd(q.g_t0 + c, r);
    r = block_sum(s_t1, s_red);
    if (threadIdx.x == 0 && q.g_t1 && p.t1[c] >= 0.f) atomicAdd(q.g_t1 + c, r);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (2) data gradient, stride 1:  g_in[b,ci,y,x] (+)= sum_{co,dy,dx} w[co,ci,dy,dx] * g_I[b,co,y+1-dy,x+1-dx]
//     (+ PLIF/XLIF trace term: sign(x)/Cin * avgpool^T(gP_sum)).  16x16 tile, 2 rows / thread, 32 ci per block.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DG_THREADS = 128, DG_CK = 8, DG_CIB = 32, DG_WP = 36;

__global__ void __launch_bounds__(DG_THREADS) conv_dgrad_kernel(const float* __restrict__ g_I, const float* __restrict__ w,
                                                                float* __restrict__ g_in, int accumulate, int B, int Cin, int C,
                                                                int H, int W, const float* __restrict__ gP_sum,
                                                                const float* __restrict__ x_for_sign) {
  __shared__ __align__(16) float s_g[DG_CK * 18 * 18];
  __shared__ __align__(16) float s_w[DG_CK * 9 * DG_WP];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ciblocks = (Cin + DG_CIB - 1) / DG_CIB;
  const int b = blockIdx.z / ciblocks, ci0 = (blockIdx.z % ciblocks) * DG_CIB;
  const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
  float acc[2][DG_CIB];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < DG_CIB; ++j) acc[i][j] = 0.f;

  for (int co0 = 0; co0 < C; co0 += DG_CK) {
    __syncthreads();
    for (int i = tid; i < DG_CK * 18 * 18; i += DG_THREADS) {
      const int co = i / 324, r = i % 324, hy = r / 18, hx = r % 18;
      const int y = y0 - 1 + hy, x = x0 - 1 + hx;
      float v = 0.f;
      if (co0 + co < C && y >= 0 && y < H && x >= 0 && x < W) v = g_I[(((size_t)b * C + co0 + co) * H + y) * W + x];
      s_g[i] = v;
    }
    // s_w[(co*9 + tap')][ci] with tap' = flipped tap, so the main loop is a plain correlation over the halo
    for (int i = tid; i < DG_CK * DG_CIB * 9; i += DG_THREADS) {
      const int co = i / (DG_CIB * 9), r = i % (DG_CIB * 9), ci = r / 9, tap = r % 9;
      float v = 0.f;
      if (co0 + co < C && ci0 + ci < Cin) v = w[((size_t)(co0 + co) * Cin + ci0 + ci) * 9 + tap];
      s_w[(co * 9 + (8 - tap)) * DG_WP + ci] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int co = 0; co < DG_CK; ++co) {
      const float* sg = s_g + co * 324;
      float ga[9], gb[9];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          ga[dy * 3 + dx] = sg[(ty + dy) * 18 + tx + dx];
          gb[dy * 3 + dx] = sg[(ty + 8 + dy) * 18 + tx + dx];
        }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + (co * 9 + tap) * DG_WP);
#pragma unroll
        for (int qd = 0; qd < DG_CIB / 4; ++qd) {
          const float4 w4 = wr[qd];
          acc[0][4 * qd + 0] = fmaf(ga[tap], w4.x, acc[0][4 * qd + 0]);
          acc[0][4 * qd + 1] = fmaf(ga[tap], w4.y, acc[0][4 * qd + 1]);
          acc[0][4 * qd + 2] = fmaf(ga[tap], w4.z, acc[0][4 * qd + 2]);
          acc[0][4 * qd + 3] = fmaf(ga[tap], w4.w, acc[0][4 * qd + 3]);
          acc[1][4 * qd + 0] = fmaf(gb[tap], w4.x, acc[1][4 * qd + 0]);
          acc[1][4 * qd + 1] = fmaf(gb[tap], w4.y, acc[1][4 * qd + 1]);
          acc[1][4 * qd + 2] = fmaf(gb[tap], w4.z, acc[1][4 * qd + 2]);
          acc[1][4 * qd + 3] = fmaf(gb[tap], w4.w, acc[1][4 * qd + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int y = y0 + ty + half * 8, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float tr = 0.f;
    if (gP_sum) {  // adjoint of the stride-1 3x3 average pool (count_include_pad): sum of the 3x3 neighbourhood / 9
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) tr += gP_sum[((size_t)b * H + yy) * W + xx];
        }
      tr = tr / 9.0f / (float)Cin;
    }
#pragma unroll
    for (int ci = 0; ci < DG_CIB; ++ci) {
      if (ci0 + ci >= Cin) break;
      const size_t o = (((size_t)b * Cin + ci0 + ci) * H + y) * W + x;
      float v = acc[half][ci];
      if (gP_sum) {
        const float xv = x_for_sign[o];
        v += (xv > 0.f ? tr : (xv < 0.f ? -tr : 0.f));
      }
      g_in[o] = accumulate ? g_in[o] + v : v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (3) weight gradient: g_w[co,ci,dy,dx] += sum_{b,y,x} in[b,ci,y*S+dy-1,x*S+dx-1] * g_I[b,co,y,x]
//     One CTA per 16x16 output tile; thread = (co, pair of ci within an 8-channel chunk); sliding 3-wide window along x.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 128, WG_GP = 257;

__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_kernel(const float* __restrict__ in, const uint16_t* __restrict__ in_cl,
                                                                const float* __restrict__ g_I,
                                                                float* __restrict__ g_w, int B, int Cin, int C, int H, int W,
                                                                int Ho, int Wo) {
  constexpr int STRIDE = 1, HH = 18, HW = 18;
  __shared__ float s_g[32 * WG_GP];
  __shared__ float s_x[8 * HH * HW];
  const int tid = threadIdx.x, co_l = tid & 31, cp = tid >> 5;  // cp: which pair of the chunk's 8 input channels
  const int coblocks = (C + 31) / 32;
  const int b = blockIdx.z / coblocks, co0 = (blockIdx.z % coblocks) * 32;
  const int ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16;
  // g_I tile, zero outside the image
  for (int i = tid; i < 32 * 256; i += WG_THREADS) {
    const int co = i >> 8, r = i & 255, yy = oy0 + (r >> 4), xx = ox0 + (r & 15);
    float v = 0.f;
    if (co0 + co < C && yy < Ho && xx < Wo) v = g_I[(((size_t)b * C + co0 + co) * Ho + yy) * Wo + xx];
    s_g[co * WG_GP + r] = v;
  }
  const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
  for (int ci0 = 0; ci0 < Cin; ci0 += 8) {
    __syncthreads();
    for (int i = tid; i < 8 * HH * HW; i += WG_THREADS) {
      const int ci = i / (HH * HW), r = i % (HH * HW), y = iy0 + r / HW, x = ix0 + r % HW;
      float v = 0.f;
      if (ci0 + ci < Cin && y >= 0 && y < H && x >= 0 && x < W) v = ld_act2(in, in_cl, b, ci0 + ci, (size_t)y * W + x, Cin, (size_t)H * W);
      s_x[i] = v;
    }
    __syncthreads();
    float acc[2][9];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 9; ++j) acc[i][j] = 0.f;
    const float* sx0 = s_x + (cp * 2) * HH * HW;
    const float* sx1 = sx0 + HH * HW;
    const float* sg = s_g + co_l * WG_GP;
    {
      for (int y = 0; y < 16; ++y) {
        float w0[3][3], w1[3][3];  // [dy][window position]
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          w0[dy][1] = sx0[(y + dy) * HW + 0];
          w0[dy][2] = sx0[(y + dy) * HW + 1];
          w1[dy][1] = sx1[(y + dy) * HW + 0];
          w1[dy][2] = sx1[(y + dy) * HW + 1];
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const float g = sg[y * 16 + x];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            w0[dy][0] = w0[dy][1];
            w0[dy][1] = w0[dy][2];
            w0[dy][2] = sx0[(y + dy) * HW + x + 2];
            w1[dy][0] = w1[dy][1];
            w1[dy][1] = w1[dy][2];
            w1[dy][2] = sx1[(y + dy) * HW + x + 2];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              acc[0][dy * 3 + dx] = fmaf(w0[dy][dx], g, acc[0][dy * 3 + dx]);
              acc[1][dy * 3 + dx] = fmaf(w1[dy][dx], g, acc[1][dy * 3 + dx]);
            }
          }
        }
      }
    }
    if (co0 + co_l < C) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ci = ci0 + cp * 2 + i;
        if (ci < Cin) {
          float* dst = g_w + ((size_t)(co0 + co_l) * Cin + ci) * 9;
#pragma unroll
          for (int t = 0; t < 9; ++t) atomicAdd(dst + t, acc[i][t]);
        }
      }
    }
  }
}

// Stride 2: the transposed convolution / the weight correlation of a stride-2 conv equal the stride-1 ones applied to the
// output gradient with zeros inserted between its pixels (up[2oy][2ox] = g[oy][ox]), so the stride-1 kernels above are reused
// on an input-resolution copy (4x their minimal FLOPs; the four encoder convs of the U-Net are the only stride-2 layers).
__global__ void __launch_bounds__(256) zero_insert2_kernel(const float* __restrict__ g, float* __restrict__ up, size_t n_planes, int H, int W, int Ho,
                                                        int Wo) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_planes * H * W) return;
  const int x = i % W, y = (i / W) % H;
  const size_t pl = i / ((size_t)W * H);
  up[i] = ((x | y) & 1) ? 0.f : g[pl * Ho * Wo + (size_t)(y >> 1) * Wo + (x >> 1)];
}

template <int NEURON, bool HARD>
static int launch_pointwise(const ef_lif_conv_bwd_params& q, int Ho, int Wo, float* gP_sum, cudaStream_t st) {
  dim3 grid(cdiv(Ho * Wo, PW_THREADS * PW_PER_THREAD), q.f.C, q.f.B);
  lif_bwd_pointwise_kernel<NEURON, HARD><<<grid, PW_THREADS, 0, st>>>(q, Ho, Wo, gP_sum);
  return check_launch("lif_bwd_pointwise_kernel");
}

}  // namespace ef

extern "C" int ef_lif_conv_bwd(const ef_lif_conv_bwd_params* q, void* stream) {
  using namespace ef;
  EF_REQUIRE(q, EF_ENULL, "ef_lif_conv_bwd: params is NULL");
  const ef_lif_conv_params& p = q->f;
  EF_REQUIRE(p.B > 0 && p.Cin > 0 && p.C > 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_lif_conv_bwd: non-positive dimension");
  EF_REQUIRE(p.ksize == 3 && (p.stride == 1 || p.stride == 2), EF_EUNSUPPORTED, "ef_lif_conv_bwd: kernel_size 3, stride 1 or 2");
  EF_REQUIRE(p.stride == 1 || q->scratch_gI_up, EF_ENULL, "ef_lif_conv_bwd: stride 2 needs scratch_gI_up");
  EF_REQUIRE((p.x || p.x_cl) && p.w_ff && p.leak && p.v_out && q->scratch_gI, EF_ENULL, "ef_lif_conv_bwd: x / w_ff / leak / v_out / scratch is NULL");
  EF_REQUIRE(p.x || !(p.neuron == EF_PLIF || p.neuron == EF_XLIF) || !q->g_x, EF_EUNSUPPORTED, "ef_lif_conv_bwd: PLIF / XLIF data gradient needs the fp32 input");
  EF_REQUIRE(p.neuron == EF_LIF || p.aux_out, EF_ENULL, "ef_lif_conv_bwd: aux_out is NULL");
  EF_REQUIRE(!q->reset_grad || !(p.z_in || p.z_in_cl) || q->g_z_in, EF_ENULL, "ef_lif_conv_bwd: reset_grad needs g_z_in");
  cudaStream_t st = as_stream(stream);
  const int Ho = (p.H - 1) / p.stride + 1, Wo = (p.W - 1) / p.stride + 1;

  float* gP_sum = nullptr;
  if ((p.neuron == EF_PLIF || p.neuron == EF_XLIF) && (q->g_x || (q->neuron_only && q->scratch_gP))) {
    EF_REQUIRE(q->scratch_gP, EF_ENULL, "ef_lif_conv_bwd: PLIF / XLIF data gradient needs scratch_gP");
    gP_sum = q->scratch_gP;
    cudaMemsetAsync(gP_sum, 0, (size_t)p.B * Ho * Wo * sizeof(float), st);
  }

  int rc;
  switch (p.neuron * 2 + (p.hard_reset ? 1 : 0)) {
    case EF_LIF * 2 + 0: rc = launch_pointwise<EF_LIF, false>(*q, Ho, Wo, gP_sum, st); break;
    case EF_LIF * 2 + 1: rc = launch_pointwise<EF_LIF, true>(*q, Ho, Wo, gP_sum, st); break;
    case EF_PLIF * 2 + 0: rc = launch_pointwise<EF_PLIF, false>(*q, Ho, Wo, gP_sum, st); break;
    case EF_PLIF * 2 + 1: rc = launch_pointwise<EF_PLIF, true>(*q, Ho, Wo, gP_sum, st); break;
    case EF_ALIF * 2 + 0: rc = launch_pointwise<EF_ALIF, false>(*q, Ho, Wo, gP_sum, st); break;
    case EF_ALIF * 2 + 1: rc = launch_pointwise<EF_ALIF, true>(*q, Ho, Wo, gP_sum, st); break;
    case EF_XLIF * 2 + 0: rc = launch_pointwise<EF_XLIF, false>(*q, Ho, Wo, gP_sum, st); break;
    case EF_XLIF * 2 + 1: rc = launch_pointwise<EF_XLIF, true>(*q, Ho, Wo, gP_sum, st); break;
    default: return fail(EF_EINVAL, "ef_lif_conv_bwd: bad neuron kind %d", p.neuron);
  }
  if (rc) return rc;
  if (q->neuron_only) return EF_OK;

  // gradient of the feed-forward conv output at the resolution the stride-1 kernels expect
  const float* gI_ff = q->scratch_gI;
  const float* gP_ff = gP_sum;
  if (p.stride == 2) {
    const size_t n = (size_t)p.B * p.C * p.H * p.W;
    zero_insert2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q->scratch_gI, q->scratch_gI_up, (size_t)p.B * p.C, p.H, p.W, Ho, Wo);
    if ((rc = check_launch("zero_insert2_kernel"))) return rc;
    gI_ff = q->scratch_gI_up;
    if (gP_sum) {
      EF_REQUIRE(q->scratch_gP_up, EF_ENULL, "ef_lif_conv_bwd: stride-2 PLIF / XLIF data gradient needs scratch_gP_up");
      const size_t m = (size_t)p.B * p.H * p.W;
      zero_insert2_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(gP_sum, q->scratch_gP_up, (size_t)p.B, p.H, p.W, Ho, Wo);
      if ((rc = check_launch("zero_insert2_kernel(gP)"))) return rc;
      gP_ff = q->scratch_gP_up;
    }
  }
  if (q->g_x) {
    dim3 grid(cdiv(p.W, 16), cdiv(p.H, 16), p.B * cdiv(p.Cin, DG_CIB));
    conv_dgrad_kernel<<<grid, DG_THREADS, 0, st>>>(gI_ff, p.w_ff, q->g_x, 0, p.B, p.Cin, p.C, p.H, p.W, gP_ff, p.x);
    if ((rc = check_launch("conv_dgrad_kernel(ff)"))) return rc;
  }
  if (p.w_rec && q->g_z_in && (p.z_in || p.z_in_cl)) {
    dim3 grid(cdiv(Wo, 16), cdiv(Ho, 16), p.B * cdiv(p.C, DG_CIB));
    conv_dgrad_kernel<<<grid, DG_THREADS, 0, st>>>(q->scratch_gI, p.w_rec, q->g_z_in, 1, p.B, p.C, p.C, Ho, Wo, nullptr, nullptr);
    if ((rc = check_launch("conv_dgrad_kernel(rec)"))) return rc;
  }
  if (q->g_w_ff) {
    const int Hg = p.stride == 2 ? p.H : Ho, Wg = p.stride == 2 ? p.W : Wo;
    dim3 grid(cdiv(Wg, 16), cdiv(Hg, 16), p.B * cdiv(p.C, 32));
    conv_wgrad_kernel<<<grid, WG_THREADS, 0, st>>>(p.x, p.x_cl, gI_ff, q->g_w_ff, p.B, p.Cin, p.C, p.H, p.W, Hg, Wg);
    if ((rc = check_launch("conv_wgrad_kernel(ff)"))) return rc;
  }
  if (p.w_rec && q->g_w_rec && (p.z_in || p.z_in_cl)) {
    dim3 grid(cdiv(Wo, 16), cdiv(Ho, 16), p.B * cdiv(p.C, 32));
    conv_wgrad_kernel<<<grid, WG_THREADS, 0, st>>>(p.z_in, p.z_in_cl, q->scratch_gI, q->g_w_rec, p.B, p.C, p.C, Ho, Wo, Ho, Wo);
    if ((rc = check_launch("conv_wgrad_kernel(rec)"))) return rc;
  }
  return EF_OK;
}

/* Gradients of a plain 3x3, stride-1, padding-1 convolution on fp32 NCHW tensors (the ANN cells, models/submodules.py): the
 * data gradient g_x = conv^T(g_pre, w) (overwritten, may be NULL) and the weight gradient g_w += x * g_pre (may be NULL). */
extern "C" int ef_conv3x3_bwd(const float* g_pre, const float* x, const float* w, float* g_x, float* g_w, int32_t B, int32_t Cin, int32_t C,
                              int32_t H, int32_t W, void* stream) {
  using namespace ef;
  EF_REQUIRE(g_pre && w, EF_ENULL, "ef_conv3x3_bwd: g_pre / w is NULL");
  EF_REQUIRE(B > 0 && Cin > 0 && C > 0 && H > 0 && W > 0, EF_EINVAL, "ef_conv3x3_bwd: non-positive dimension");
  EF_REQUIRE(!g_w || x, EF_ENULL, "ef_conv3x3_bwd: the weight gradient needs x");
  cudaStream_t st = as_stream(stream);
  int rc;
  if (g_x) {
    dim3 grid(cdiv(W, 16), cdiv(H, 16), B * cdiv(Cin, DG_CIB));
    conv_dgrad_kernel<<<grid, DG_THREADS, 0, st>>>(g_pre, w, g_x, 0, B, Cin, C, H, W, nullptr, nullptr);
    if ((rc = check_launch("conv_dgrad_kernel(ann)"))) return rc;
  }
  if (g_w) {
    dim3 grid(cdiv(W, 16), cdiv(H, 16), B * cdiv(C, 32));
    conv_wgrad_kernel<<<grid, WG_THREADS, 0, st>>>(x, nullptr, g_pre, g_w, B, Cin, C, H, W, H, W);
    if ((rc = check_launch("conv_wgrad_kernel(ann)"))) return rc;
  }
  return EF_OK;
}

extern "C" int ef_conv3x3_bwd_s(const float* g_pre, const float* x, const float* w, float* g_x, float* g_w, float* scratch_up, int32_t B, int32_t Cin,
                                int32_t C, int32_t H, int32_t W, int32_t stride, void* stream) {
  using namespace ef;
  if (stride == 1 || stride == 0) return ef_conv3x3_bwd(g_pre, x, w, g_x, g_w, B, Cin, C, H, W, stream);
  EF_REQUIRE(stride == 2, EF_EUNSUPPORTED, "ef_conv3x3_bwd_s: stride %d not supported", stride);
  EF_REQUIRE(g_pre && scratch_up, EF_ENULL, "ef_conv3x3_bwd_s: g_pre / scratch_up is NULL");
  EF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, EF_EINVAL, "ef_conv3x3_bwd_s: non-positive dimension");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const size_t n = (size_t)B * C * H * W;
  zero_insert2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(g_pre, scratch_up, (size_t)B * C, H, W, Ho, Wo);
  if (int rc = check_launch("zero_insert2_kernel(ann)")) return rc;
  return ef_conv3x3_bwd(scratch_up, x, w, g_x, g_w, B, Cin, C, H, W, stream);
}
