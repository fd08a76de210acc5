// Fused conv3x3 + spiking neuron update, fp32 CUDA-core kernel (general shapes: any Cin, C, stride 1|2, all neuron kinds).
// This is the kernel for the head layer (Cin = 2..5, fractional voxel inputs) and for every shape the tcgen05 kernel in
// lif_conv_fwd_tc.cu does not cover.  Reference: models/spiking_submodules.py:96-126 and siblings (see eventflow.h).
#include "common.cuh"

namespace ef {

constexpr int TW = 16, TH = 16;      // output tile
constexpr int NTHREADS = 128;        // 16 x 8 threads, two output rows per thread (ty and ty + 8)
constexpr int COB = 32;              // output channels per block
constexpr int WPITCH = 36;           // padded row of the weight tile (floats), keeps float4 alignment

template <int STRIDE>
struct Geo {
  static constexpr int CK = STRIDE == 1 ? 8 : 4;              // input channels per smem stage
  static constexpr int HH = (TH - 1) * STRIDE + 3;            // halo rows
  static constexpr int HW = (TW - 1) * STRIDE + 3;            // halo cols
};

__device__ __forceinline__ float ld_act(const float* f32, const uint16_t* cl, int b, int c, int y, int x, int C, int H, int W) {
  if (f32) return f32[(((size_t)b * C + c) * H + y) * W + x];
  const uint16_t u = cl[(((size_t)b * H + y) * W + x) * C + c];
  return __uint_as_float(((uint32_t)u) << 16);
}

// Accumulate one convolution (input `src`, weights `w`) into acc[2][COB].
template <int STRIDE, bool TRACE>
__device__ __forceinline__ void conv_accumulate(float (&acc)[2][COB], const float* __restrict__ src_f32,
                                                const uint16_t* __restrict__ src_cl, const float* __restrict__ w, int b, int Cin,
                                                int H, int W, int co0, int C, int oy0, int ox0, float* s_x, float* s_w,
                                                float* s_abs) {
  using G = Geo<STRIDE>;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
  for (int ci0 = 0; ci0 < Cin; ci0 += G::CK) {
    __syncthreads();
    // input halo tile, zero-filled outside the image / beyond Cin
    for (int i = tid; i < G::CK * G::HH * G::HW; i += NTHREADS) {
      const int ci = i / (G::HH * G::HW), r = i % (G::HH * G::HW), hy = r / G::HW, hx = r % G::HW;
      const int y = iy0 + hy, x = ix0 + hx, c = ci0 + ci;
      float v = 0.f;
      if (c < Cin && y >= 0 && y < H && x >= 0 && x < W) v = ld_act(src_f32, src_cl, b, c, y, x, Cin, H, W);
      s_x[i] = v;
    }
    // weight tile [ci*9+tap][co]
    for (int i = tid; i < COB * G::CK * 9; i += NTHREADS) {
      const int co = i / (G::CK * 9), r = i % (G::CK * 9), ci = r / 9;
      float v = 0.f;
      if (co0 + co < C && ci0 + ci < Cin) v = w[((size_t)(co0 + co) * Cin + ci0) * 9 + r];
      s_w[r * WPITCH + co] = v;
    }
    __syncthreads();
    if (TRACE) {  // sum_c |x| per halo position (PLIF / XLIF pre-synaptic trace input)
      for (int i = tid; i < G::HH * G::HW; i += NTHREADS) {
        float a = 0.f;
#pragma unroll
        for (int ci = 0; ci < G::CK; ++ci) a += fabsf(s_x[ci * G::HH * G::HW + i]);
        s_abs[i] += a;
      }
    }
#pragma unroll 1
    for (int ci = 0; ci < G::CK; ++ci) {
      const float* sx = s_x + ci * G::HH * G::HW;
      float xa[9], xb[9];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          xa[dy * 3 + dx] = sx[(ty * STRIDE + dy) * G::HW + tx * STRIDE + dx];
          xb[dy * 3 + dx] = sx[((ty + 8) * STRIDE + dy) * G::HW + tx * STRIDE + dx];
        }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + (ci * 9 + tap) * WPITCH);
#pragma unroll
        for (int q = 0; q < COB / 4; ++q) {
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
}

template <int NEURON, bool HARD, int STRIDE>
__global__ void __launch_bounds__(NTHREADS) lif_conv_fwd_kernel(const ef_lif_conv_params p, int Ho, int Wo) {
  using G = Geo<STRIDE>;
  constexpr bool TRACE = (NEURON == EF_PLIF || NEURON == EF_XLIF);
  constexpr int SX = (G::CK * G::HH * G::HW > 8 * 18 * 18) ? G::CK * G::HH * G::HW : 8 * 18 * 18;
  __shared__ __align__(16) float s_x[SX];
  __shared__ __align__(16) float s_w[8 * 9 * WPITCH];
  __shared__ float s_abs[TRACE ? G::HH * G::HW : 1];
  __shared__ ChanConst s_k[COB];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int cblocks = (p.C + COB - 1) / COB;
  const int b = blockIdx.z / cblocks, co0 = (blockIdx.z % cblocks) * COB;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;

  if (tid < COB && co0 + tid < p.C) s_k[tid] = load_chan_const(p, co0 + tid);
  if (TRACE)
    for (int i = tid; i < G::HH * G::HW; i += NTHREADS) s_abs[i] = 0.f;

  float acc[2][COB];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < COB; ++j) acc[i][j] = 0.f;

  conv_accumulate<STRIDE, TRACE>(acc, p.x, p.x_cl, p.w_ff, b, p.Cin, p.H, p.W, co0, p.C, oy0, ox0, s_x, s_w, s_abs);
  if (p.w_rec && (p.z_in || p.z_in_cl))  // recurrent current: stride-1 conv of the previous spikes at output resolution
    conv_accumulate<1, false>(acc, p.z_in, p.z_in_cl, p.w_rec, b, p.C, Ho, Wo, co0, p.C, oy0, ox0, s_x, s_w, s_abs);
  __syncthreads();

  const size_t plane = (size_t)Ho * Wo;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int oy = oy0 + ty + half * 8, ox = ox0 + tx;
    if (oy >= Ho || ox >= Wo) continue;
    float P = 0.f;
    if (TRACE) {
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) s += s_abs[((ty + half * 8) * STRIDE + dy) * G::HW + tx * STRIDE + dx] / (float)p.Cin;
      P = s / 9.0f;
    }
    const size_t pix = (size_t)oy * Wo + ox;
    uint32_t zpk[COB / 2], opk[COB / 2];
#pragma unroll
    for (int co = 0; co < COB; ++co) {
      const int c = co0 + co;
      float zo = 0.f, oo = 0.f;
      if (c < p.C) {
        const size_t o = ((size_t)b * p.C + c) * plane + pix;
        const float v = p.v_in ? p.v_in[o] : 0.f;
        float z = 0.f;
        if (p.z_in) z = p.z_in[o];
        else if (p.z_in_cl) z = ld_act(nullptr, p.z_in_cl, b, c, oy, ox, p.C, Ho, Wo);
        const float aux = p.aux_in ? p.aux_in[o] : 0.f;
        float vo, ao, thr;
        neuron_update<NEURON, HARD>(acc[half][co], v, z, aux, P, s_k[co], vo, zo, ao, thr);
        p.v_out[o] = vo;
        if (p.z_out) p.z_out[o] = zo;
        if (p.aux_out) p.aux_out[o] = ao;
        oo = p.residual ? __fadd_rn(zo, p.residual[o]) : zo;
        if (p.out) p.out[o] = oo;
      }
      if (co & 1) {
        zpk[co >> 1] = (zpk[co >> 1] & 0xffffu) | (pack_bf16x2(0.f, zo) & 0xffff0000u);
        opk[co >> 1] = (opk[co >> 1] & 0xffffu) | (pack_bf16x2(0.f, oo) & 0xffff0000u);
      } else {
        zpk[co >> 1] = pack_bf16x2(zo, 0.f) & 0xffffu;
        opk[co >> 1] = pack_bf16x2(oo, 0.f) & 0xffffu;
      }
    }
    // channels-last bf16 outputs: 16 B per (pixel, 8-channel group)
#pragma unroll
    for (int g = 0; g < COB / 8; ++g) {
      const int cg = (co0 >> 3) + g;
      if (cg * 8 >= p.C) break;
      const size_t o8 = (((size_t)b * Ho + oy) * Wo + ox) * p.C + cg * 8;
      if (p.z_out_cl) *reinterpret_cast<uint4*>(p.z_out_cl + o8) = make_uint4(zpk[4 * g], zpk[4 * g + 1], zpk[4 * g + 2], zpk[4 * g + 3]);
      if (p.out_cl) *reinterpret_cast<uint4*>(p.out_cl + o8) = make_uint4(opk[4 * g], opk[4 * g + 1], opk[4 * g + 2], opk[4 * g + 3]);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Head layer of the fast path: few input channels (event counts / voxel bins, Cin <= 8, fractional values allowed),
// 32 output channels, LIF.  One thread = one pixel x 32 channels; 32 x 8 pixel tile so that every fp32 NCHW access of a
// warp is one 128-byte line.  Writes the membrane fp32 NCHW and the spikes in cl for the tensor-core layers.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HD_THREADS = 128, HD_MAXC = 8, HD_CTAS_PER_SM = 3;

// packed fp32 pairs (sm_100: two IEEE fmas per instruction; each lane is an ordinary fma.rn, so the result is bit-identical
// to the scalar loop)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// One warp = one strip of 32 x 2 pixels; one thread = two vertically adjacent pixels x 32 channels, so every weight read
// from shared memory (one 16-byte broadcast load = 4 channels) feeds 8 fmas, issued as 4 packed-pair instructions.  Inputs
// are read straight from global memory (L1-resident 3x4 neighbourhood per input channel, prefetched one channel ahead);
// every fp32 NCHW access of a warp is one 128-byte line.  Warps loop over strips independently (no CTA barrier after the
// weight staging).
template <bool HARD>
__global__ void __launch_bounds__(HD_THREADS, HD_CTAS_PER_SM) lif_head_fwd_kernel(const ef_lif_conv_params p, int strips_x, int strips_y, int n_strips) {
  __shared__ __align__(16) float s_w[HD_MAXC * 9 * 32];
  __shared__ ChanConst s_k[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Cin = p.Cin, H = p.H, W = p.W;
  pdl_launch_dependents();  // first kernel of a model step: the next one (a fused cell) may start its prologue early
  pdl_wait();               // everything this kernel reads may come from the previous kernel in the stream
  if (tid < 32) s_k[tid] = load_chan_const(p, tid);
  for (int i = tid; i < Cin * 9 * 32; i += HD_THREADS) {  // s_w[(ci*9 + tap)*32 + co] = w[co][ci][tap]
    const int co = i & 31, r = i >> 5;
    s_w[i] = p.w_ff[(size_t)co * Cin * 9 + r];
  }
  __syncthreads();
  const size_t plane = (size_t)H * W;
  const int per_img = strips_x * strips_y;
  for (int strip = blockIdx.x * (HD_THREADS / 32) + warp; strip < n_strips; strip += gridDim.x * (HD_THREADS / 32)) {
    const int b = strip / per_img, r = strip - b * per_img;
    const int sy = r / strips_x;
    const int y0 = sy * 2, x = (r - sy * strips_x) * 32 + lane;
    const bool in_x = x < W;
    const float* xb = p.x + (size_t)b * Cin * plane;
    // 4 rows x 3 columns around the pixel pair, zero outside the image
    auto load_nb = [&](int ci, float (&nb)[12]) {
      const float* xc = xb + (size_t)ci * plane;
#pragma unroll
      for (int ry = 0; ry < 4; ++ry) {
        const int y = y0 - 1 + ry;
        const bool in_y = y >= 0 && y < H;
#pragma unroll
        for (int rx = 0; rx < 3; ++rx) {
          const int xx = x - 1 + rx;
          nb[ry * 3 + rx] = (in_y && xx >= 0 && xx < W) ? __ldg(xc + (size_t)y * W + xx) : 0.f;
        }
      }
    };
    unsigned long long accA[16], accB[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) accA[j] = accB[j] = 0ull;
    float nb[12], nb_next[12];
    load_nb(0, nb);
#pragma unroll 1
    for (int ci = 0; ci < Cin; ++ci) {
      if (ci + 1 < Cin) load_nb(ci + 1, nb_next);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap % 3;
        const unsigned long long xa = pack2(nb[dy * 3 + dx], nb[dy * 3 + dx]);
        const unsigned long long xbp = pack2(nb[(dy + 1) * 3 + dx], nb[(dy + 1) * 3 + dx]);
        const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(s_w + (ci * 9 + tap) * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const ulonglong2 w4 = wr[q];  // channels 4q .. 4q+3
          accA[2 * q] = fma2(xa, w4.x, accA[2 * q]);
          accA[2 * q + 1] = fma2(xa, w4.y, accA[2 * q + 1]);
          accB[2 * q] = fma2(xbp, w4.x, accB[2 * q]);
          accB[2 * q + 1] = fma2(xbp, w4.y, accB[2 * q + 1]);
        }
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) nb[k] = nb_next[k];
    }
    // neuron update of the two pixels
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int y = y0 + half;
      if (!(in_x && y < H)) continue;
      const size_t pix = (size_t)y * W + x;
      float vin[32];
      uint4 zq[4];
#pragma unroll
      for (int c = 0; c < 32; ++c) vin[c] = p.v_in ? __ldg(p.v_in + ((size_t)b * 32 + c) * plane + pix) : 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        zq[g] = p.z_in_cl ? __ldg(reinterpret_cast<const uint4*>(p.z_in_cl + ((size_t)b * plane + pix) * 32 + g * 8)) : make_uint4(0, 0, 0, 0);
      uint32_t zpk[16];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float a0, a1;
        unpack2(half == 0 ? accA[c >> 1] : accB[c >> 1], a0, a1);
        const float I = (c & 1) ? a1 : a0;
        const uint32_t zw = (&zq[c >> 3].x)[(c & 7) >> 1];
        const float z = (c & 1) ? bf16_hi(zw) : bf16_lo(zw);
        float vo, zo, ao, thr;
        neuron_update<EF_LIF, HARD>(I, vin[c], z, 0.f, 0.f, s_k[c], vo, zo, ao, thr);
        p.v_out[((size_t)b * 32 + c) * plane + pix] = vo;
        const uint32_t zb = zo > 0.f ? 0x3F80u : 0u;
        if (c & 1) zpk[c >> 1] |= zb << 16;
        else zpk[c >> 1] = zb;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(p.z_out_cl + ((size_t)b * plane + pix) * 32 + g * 8) = make_uint4(zpk[4 * g], zpk[4 * g + 1], zpk[4 * g + 2], zpk[4 * g + 3]);
    }
  }
}

static bool head_eligible(const ef_lif_conv_params& p) {
  return p.neuron == EF_LIF && p.C == 32 && p.Cin <= HD_MAXC && p.stride == 1 && p.x && !p.x_cl && !p.w_rec && !p.z_in && !p.residual && !p.out &&
         !p.z_out && !p.out_cl && p.z_out_cl && (!p.v_in == !p.z_in_cl);
}

static int launch_head(const ef_lif_conv_params& p, cudaStream_t st) {
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int strips_x = cdiv(p.W, 32), strips_y = cdiv(p.H, 2), n_strips = strips_x * strips_y * p.B;
  const int slots = n_sms * HD_CTAS_PER_SM;  // persistent grid = one resident wave of CTAs, warps loop over strips
  const int want = cdiv(n_strips, HD_THREADS / 32);
  const int grid = want < slots ? want : slots;
  if (p.hard_reset)
    launch_pdl(lif_head_fwd_kernel<true>, dim3(grid), dim3(HD_THREADS), 0, st, p, strips_x, strips_y, n_strips);
  else
    launch_pdl(lif_head_fwd_kernel<false>, dim3(grid), dim3(HD_THREADS), 0, st, p, strips_x, strips_y, n_strips);
  return check_launch("lif_head_fwd_kernel");
}

template <int NEURON, bool HARD>
static int launch_generic(const ef_lif_conv_params& p, int Ho, int Wo, cudaStream_t st) {
  dim3 grid(cdiv(Wo, TW), cdiv(Ho, TH), p.B * cdiv(p.C, COB));
  if (p.stride == 1)
    lif_conv_fwd_kernel<NEURON, HARD, 1><<<grid, NTHREADS, 0, st>>>(p, Ho, Wo);
  else
    lif_conv_fwd_kernel<NEURON, HARD, 2><<<grid, NTHREADS, 0, st>>>(p, Ho, Wo);
  return check_launch("lif_conv_fwd_kernel");
}

int lif_conv_fwd_generic(const ef_lif_conv_params& p, cudaStream_t st) {
  const int Ho = (p.H - 1) / p.stride + 1, Wo = (p.W - 1) / p.stride + 1;
  switch (p.neuron * 2 + (p.hard_reset ? 1 : 0)) {
    case EF_LIF * 2 + 0: return launch_generic<EF_LIF, false>(p, Ho, Wo, st);
    case EF_LIF * 2 + 1: return launch_generic<EF_LIF, true>(p, Ho, Wo, st);
    case EF_PLIF * 2 + 0: return launch_generic<EF_PLIF, false>(p, Ho, Wo, st);
    case EF_PLIF * 2 + 1: return launch_generic<EF_PLIF, true>(p, Ho, Wo, st);
    case EF_ALIF * 2 + 0: return launch_generic<EF_ALIF, false>(p, Ho, Wo, st);
    case EF_ALIF * 2 + 1: return launch_generic<EF_ALIF, true>(p, Ho, Wo, st);
    case EF_XLIF * 2 + 0: return launch_generic<EF_XLIF, false>(p, Ho, Wo, st);
    case EF_XLIF * 2 + 1: return launch_generic<EF_XLIF, true>(p, Ho, Wo, st);
  }
  return fail(EF_EINVAL, "ef_lif_conv_fwd: bad neuron kind %d", p.neuron);
}

int validate_lif_conv(const ef_lif_conv_params& p, const char* who) {
  EF_REQUIRE(p.B > 0 && p.Cin > 0 && p.C > 0 && p.H > 0 && p.W > 0, EF_EINVAL, "%s: non-positive dimension", who);
  EF_REQUIRE(p.ksize == 3, EF_EUNSUPPORTED, "%s: kernel_size %d not supported (3 only)", who, p.ksize);
  EF_REQUIRE(p.stride == 1 || p.stride == 2, EF_EUNSUPPORTED, "%s: stride %d not supported", who, p.stride);
  EF_REQUIRE(p.neuron >= EF_LIF && p.neuron <= EF_XLIF, EF_EINVAL, "%s: bad neuron kind %d", who, p.neuron);
  EF_REQUIRE(p.x || p.x_cl, EF_ENULL, "%s: x is NULL", who);
  EF_REQUIRE(!p.x_cl || p.Cin % 8 == 0, EF_EINVAL, "%s: cl input needs Cin %% 8 == 0", who);
  EF_REQUIRE(!(p.z_in_cl || p.z_out_cl || p.out_cl) || p.C % 8 == 0, EF_EINVAL, "%s: cl spikes need C %% 8 == 0", who);
  EF_REQUIRE(p.w_ff && p.leak && p.v_out, EF_ENULL, "%s: w_ff / leak / v_out is NULL", who);
  if (p.neuron == EF_LIF || p.neuron == EF_PLIF) EF_REQUIRE(p.thresh, EF_ENULL, "%s: thresh is NULL", who);
  if (p.neuron != EF_LIF) EF_REQUIRE(p.leak_aux, EF_ENULL, "%s: leak_pt / leak_t is NULL", who);
  if (p.neuron == EF_PLIF) EF_REQUIRE(p.add_pt, EF_ENULL, "%s: add_pt is NULL", who);
  if (p.neuron == EF_ALIF || p.neuron == EF_XLIF) EF_REQUIRE(p.t0 && p.t1, EF_ENULL, "%s: t0 / t1 is NULL", who);
  if (p.neuron != EF_LIF) EF_REQUIRE(p.aux_out, EF_ENULL, "%s: aux_out is NULL", who);
  return EF_OK;
}

// ---- neuron update on a GIVEN synaptic current ----------------------------------------------------------------------------------
// The second half of a cell step whose convolution ran elsewhere (ef_lif_neuron_fwd): the tensor-core kernel computes
// cur = conv(x, w_ff) (+ conv(z, w_rec)) for 32-channel cells of any neuron kind, this kernel applies the PLIF / ALIF / XLIF / LIF update
// (models/spiking_submodules.py:164-227, 265-334, 372-435 and the recurrent twins) on the reference's fp32 NCHW state tensors.
// Block = 32 x 8 pixels of one sample, thread = pixel: per channel all accesses of a warp are one 128-byte row segment.
constexpr int NU_TW = 32, NU_TH = 8;
template <int NEURON, bool HARD>
__global__ void __launch_bounds__(NU_TW * NU_TH) neuron_fwd_kernel(const ef_lif_conv_params p, const float* __restrict__ cur) {
  constexpr bool TRACE = (NEURON == EF_PLIF || NEURON == EF_XLIF);
  __shared__ float s_abs[TRACE ? (NU_TH + 2) * (NU_TW + 2) : 1];
  __shared__ ChanConst s_k[64];
  const int tid = threadIdx.x, tx = tid % NU_TW, ty = tid / NU_TW;
  const int b = blockIdx.z, x0 = blockIdx.x * NU_TW, y0 = blockIdx.y * NU_TH;
  const int H = p.H, W = p.W, C = p.C;
  const size_t plane = (size_t)H * W;
  if (TRACE) {  // sum_c |x| on the tile and its 1-pixel halo (zero outside the image: count_include_pad)
    for (int i = tid; i < (NU_TH + 2) * (NU_TW + 2); i += NU_TW * NU_TH) {
      const int hy = i / (NU_TW + 2), hx = i - hy * (NU_TW + 2);
      const int y = y0 + hy - 1, x = x0 + hx - 1;
      float a = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float* xp = p.x + (size_t)b * p.Cin * plane + (size_t)y * W + x;
        int c = 0;
        for (; c + 8 <= p.Cin; c += 8) {  // 8 independent loads in flight
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] = __ldg(xp + (size_t)(c + j) * plane);
#pragma unroll
          for (int j = 0; j < 8; ++j) a += fabsf(t[j]);
        }
        for (; c < p.Cin; ++c) a += fabsf(__ldg(xp + (size_t)c * plane));
      }
      s_abs[i] = a;
    }
  }
  const int y = y0 + ty, x = x0 + tx;
  const bool inside = y < H && x < W;
  float P = 0.f;
  for (int c0 = 0; c0 < C; c0 += 64) {  // channel constants in blocks of 64
    __syncthreads();
    if (tid < 64 && c0 + tid < C) s_k[tid] = load_chan_const(p, c0 + tid);
    __syncthreads();
    if (TRACE && c0 == 0) {
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) s += s_abs[(ty + dy) * (NU_TW + 2) + tx + dx] / (float)p.Cin;
      P = s / 9.0f;
    }
    if (!inside) continue;
    const int cn = min(64, C - c0);
    for (int g0 = 0; g0 < cn; g0 += 8) {  // 8 channels = one 16-byte piece of the channels-last outputs
      // all loads of the group first, then the updates and stores (the tensors come without restrict: a store between the loads of two
      // channels would serialise them)
      float in_c[8], in_v[8], in_z[8], in_a[8], in_r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + g0 + j;
        const size_t o = ((size_t)b * C + min(c, C - 1)) * plane + (size_t)y * W + x;
        in_c[j] = __ldg(cur + o);
        in_v[j] = p.v_in ? __ldg(p.v_in + o) : 0.f;
        in_z[j] = p.z_in ? __ldg(p.z_in + o) : 0.f;
        in_a[j] = p.aux_in ? __ldg(p.aux_in + o) : 0.f;
        in_r[j] = p.residual ? __ldg(p.residual + o) : 0.f;
      }
      uint32_t zpk[4] = {0, 0, 0, 0}, opk[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + g0 + j;
        float zo = 0.f, oo = 0.f;
        if (c < C) {
          const size_t o = ((size_t)b * C + c) * plane + (size_t)y * W + x;
          float vo, ao, thr;
          neuron_update<NEURON, HARD>(in_c[j], in_v[j], in_z[j], in_a[j], P, s_k[g0 + j], vo, zo, ao, thr);
          p.v_out[o] = vo;
          if (p.z_out) p.z_out[o] = zo;
          if (p.aux_out) p.aux_out[o] = ao;
          oo = p.residual ? __fadd_rn(zo, in_r[j]) : zo;
          if (p.out) p.out[o] = oo;
        }
        const uint32_t zb = pack_bf16x2(zo, 0.f) & 0xffffu, ob = pack_bf16x2(oo, 0.f) & 0xffffu;
        zpk[j >> 1] |= (j & 1) ? (zb << 16) : zb;
        opk[j >> 1] |= (j & 1) ? (ob << 16) : ob;
      }
      if (C % 8 == 0) {
        const size_t o8 = (((size_t)b * H + y) * W + x) * C + c0 + g0;
        if (p.z_out_cl) *reinterpret_cast<uint4*>(p.z_out_cl + o8) = make_uint4(zpk[0], zpk[1], zpk[2], zpk[3]);
        if (p.out_cl) *reinterpret_cast<uint4*>(p.out_cl + o8) = make_uint4(opk[0], opk[1], opk[2], opk[3]);
      }
    }
  }
}

template <int NEURON, bool HARD>
static int launch_neuron(const ef_lif_conv_params& p, const float* cur, cudaStream_t st) {
  const dim3 grid(cdiv(p.W, NU_TW), cdiv(p.H, NU_TH), p.B);
  neuron_fwd_kernel<NEURON, HARD><<<grid, NU_TW * NU_TH, 0, st>>>(p, cur);
  return check_launch("neuron_fwd_kernel");
}

int neuron_fwd(const ef_lif_conv_params& p, const float* cur, cudaStream_t st) {
  switch (p.neuron * 2 + (p.hard_reset ? 1 : 0)) {
    case EF_LIF * 2 + 0: return launch_neuron<EF_LIF, false>(p, cur, st);
    case EF_LIF * 2 + 1: return launch_neuron<EF_LIF, true>(p, cur, st);
    case EF_PLIF * 2 + 0: return launch_neuron<EF_PLIF, false>(p, cur, st);
    case EF_PLIF * 2 + 1: return launch_neuron<EF_PLIF, true>(p, cur, st);
    case EF_ALIF * 2 + 0: return launch_neuron<EF_ALIF, false>(p, cur, st);
    case EF_ALIF * 2 + 1: return launch_neuron<EF_ALIF, true>(p, cur, st);
    case EF_XLIF * 2 + 0: return launch_neuron<EF_XLIF, false>(p, cur, st);
    case EF_XLIF * 2 + 1: return launch_neuron<EF_XLIF, true>(p, cur, st);
  }
  return fail(EF_EINVAL, "ef_lif_neuron_fwd: bad neuron kind %d", p.neuron);
}

int lif_conv_fwd_tc(const ef_lif_conv_params& p, cudaStream_t st);  // lif_conv_fwd_tc.cu
bool lif_conv_tc_eligible(const ef_lif_conv_params& p);

}  // namespace ef

extern "C" int ef_lif_conv_fwd(const ef_lif_conv_params* p, void* stream) {
  EF_REQUIRE(p, EF_ENULL, "ef_lif_conv_fwd: params is NULL");
  if (int rc = ef::validate_lif_conv(*p, "ef_lif_conv_fwd")) return rc;
  if (ef::lif_conv_tc_eligible(*p)) return ef::lif_conv_fwd_tc(*p, ef::as_stream(stream));
  if (ef::head_eligible(*p)) return ef::launch_head(*p, ef::as_stream(stream));
  return ef::lif_conv_fwd_generic(*p, ef::as_stream(stream));
}

// The neuron update of a cell step on a synaptic current computed elsewhere (the tensor-core convolution): same parameter block as
// ef_lif_conv_fwd; stride 1, fp32 NCHW x (read for the pre-synaptic trace of PLIF / XLIF only) and states; w_ff / w_rec are not read.
extern "C" int ef_lif_neuron_fwd(const ef_lif_conv_params* p, const float* cur, void* stream) {
  EF_REQUIRE(p && cur, EF_ENULL, "ef_lif_neuron_fwd: params / current is NULL");
  if (int rc = ef::validate_lif_conv(*p, "ef_lif_neuron_fwd")) return rc;
  EF_REQUIRE(p->stride == 1, EF_EUNSUPPORTED, "ef_lif_neuron_fwd: stride 1 only");
  EF_REQUIRE(!p->z_in_cl, EF_EUNSUPPORTED, "ef_lif_neuron_fwd: previous spikes as fp32 NCHW (z_in)");
  EF_REQUIRE((p->neuron != EF_PLIF && p->neuron != EF_XLIF) || p->x, EF_ENULL, "ef_lif_neuron_fwd: PLIF / XLIF need the fp32 input x");
  return ef::neuron_fwd(*p, cur, ef::as_stream(stream));
}
