// Backward of a 32->32 LIF cell-step on the fast-path formats (spikes channels-last bf16, membrane fp32 NCHW), sm_100a.
// Recurrences: SURVEY.md 8a (what autograd derives from models/spiking_submodules.py:96-126, 516-551 with the surrogates of
// models/spiking_util.py:39-93).  Three kernels:
//   (1) lif_bwd_pointwise_cl_kernel  neuron backward: g_I = (1-leak) g_v, g_v_in, per-channel parameter gradients; g_I is
//       written channels-last as two bf16 terms hi + mid (16 significant bits: gradients are checked to 1e-3, not bit-exact)
//   (2) lif_dgrad_tc_kernel          data gradient on tcgen05: g_x (and g_z_in of a recurrent cell) = transposed 3x3 conv of
//       g_I with the flipped weights; same implicit-GEMM structure as the forward kernel (lif_conv_fwd_tc.cu): one padded,
//       64B-swizzled halo tile per operand, taps = descriptor start addresses, weight terms stacked along N
//   (3) conv_wgrad_cl_kernel         weight gradient (CUDA cores, fp32 accumulate) reading the channels-last operands
#include "tc_common.cuh"

namespace ef {

// ---------------------------------------------------------------------------------------------------------------------
// (1) pointwise
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PWC_THREADS = 128, PWC_CB = 8, PWC_GROUPS = 32 / PWC_CB;

// The two neuron-backward kernels below are instruction-issue bound, not memory bound (ncu: 2.4 IPC per SM at 54 % of the DRAM peak,
// 68 instructions per pixel-channel-step): IEEE divisions become reciprocal-multiplies (gradients are checked to 1e-3; these are 2-ulp
// operations), the division by (1 - leak) of the leak-gradient term is factored out of the sums, and g_I is split into its two bf16
// terms two channels at a time.
__device__ __forceinline__ float surrogate_grad_fast(int kind, float x, float w) {
  if (kind == EF_ARCTAN) return __fdividef(1.0f, 1.0f + w * x * x);
  if (kind == EF_SUPERSPIKE) {
    const float d = 1.0f + w * fabsf(x);
    return __fdividef(1.0f, d * d);
  }
  return surrogate_grad(kind, x, w);
}
// g_I of two neighbouring channels -> packed hi and mid bf16 terms (hi + mid = g_I to 16 significant bits)
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& mid) {
  hi = pack_bf16x2(a, b);
  mid = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

// Sums of v[0..15] over the 32 lanes of a warp, all 16 entries at once: 16 shuffles instead of 16 x 5.  Step by step the
// lanes trade halves of their arrays (lane bit set: keep the upper half, send the lower one); after the four halving steps
// lanes l and l ^ 16 hold partial sums of the same entry, which the last shuffle combines.  Returns the total of entry
// `entry` (valid on every lane; lanes l and l ^ 16 return the same entry).
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane, int& entry) {
  entry = 0;
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const float keep = up ? v[k + half] : v[k];
      const float send = up ? v[k] : v[k + half];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
    entry += up ? half : 0;
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// One thread = one pixel x PWC_CB channels (its 16-byte piece of the channels-last rows): 40 independent loads in flight per
// thread at ~60 registers, so that enough warps are resident to cover the DRAM latency (the one-pixel-x-32-channels version
// ran at 12 warps per SM and 50 % of the HBM roofline).  Per-channel parameter-gradient terms are reduced over the warp with
// the transposing butterfly above, then over the block in shared memory.
template <int SURR, bool HARD>
__global__ void __launch_bounds__(PWC_THREADS) lif_bwd_pointwise_cl_kernel(const ef_lif_bwd_tc_params p) {
  __shared__ float s_sum[2 * PWC_CB];
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t hw = (size_t)p.H * p.W;
  const int b = blockIdx.y;
  const int c0 = (blockIdx.x % PWC_GROUPS) * PWC_CB;
  if (tid < 2 * PWC_CB) s_sum[tid] = 0.f;
  __syncthreads();
  const bool has_gout = p.g_out != nullptr, has_gv = p.g_v_out != nullptr, has_gz = p.g_z_out != nullptr, has_v = p.v_in != nullptr;
  const size_t pix = (size_t)(blockIdx.x / PWC_GROUPS) * PWC_THREADS + tid;
  const bool live = pix < hw;
  const size_t pc = live ? pix : hw - 1;  // out-of-range threads compute on a valid pixel and contribute nothing
  const uint4 zq = p.z_in_cl ? __ldg(reinterpret_cast<const uint4*>(p.z_in_cl + ((size_t)b * hw + pc) * 32 + c0)) : make_uint4(0, 0, 0, 0);
  const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
  float v_p[PWC_CB], v_n[PWC_CB], g_o[PWC_CB], g_s[PWC_CB], g_vo[PWC_CB], lam[PWC_CB], thr[PWC_CB];
#pragma unroll
  for (int k = 0; k < PWC_CB; ++k) {
    const size_t o = ((size_t)b * 32 + c0 + k) * hw + pc;
    v_p[k] = has_v ? __ldg(p.v_in + o) : 0.f;
    v_n[k] = __ldg(p.v_out + o);
    g_o[k] = has_gout ? __ldg(p.g_out + o) : 0.f;
    g_s[k] = has_gz ? __ldg(p.g_z_out + o) : 0.f;
    g_vo[k] = has_gv ? __ldg(p.g_v_out + o) : 0.f;
    lam[k] = sigmoidf_acc(__ldg(p.leak + c0 + k));
    thr[k] = fmaxf(__ldg(p.thresh + c0 + k), 0.01f);
  }
  float red[2 * PWC_CB], gI[PWC_CB];
  uint32_t hi[PWC_CB / 2], mid[PWC_CB / 2];
#pragma unroll
  for (int k = 0; k < PWC_CB; ++k) {
    const float z_p = (k & 1) ? bf16_hi(zw[k >> 1]) : bf16_lo(zw[k >> 1]);
    const float g_z = g_o[k] + g_s[k];
    const float sg = surrogate_grad_fast(SURR, v_n[k] - thr[k], p.act_width);
    const float g_v = g_vo[k] + g_z * sg;
    const float oml = 1.0f - lam[k];
    gI[k] = oml * g_v;
    const float keep = HARD ? v_p[k] * (1.0f - z_p) : v_p[k];
    // d v_out / d lam = keep - current, current = (v_out - lam keep [+ z thr]) / (1 - lam)  =>  (keep - v_out [- z thr]) / (1 - lam)
    const float dlam = (HARD ? keep - v_n[k] : v_p[k] - v_n[k] - z_p * thr[k]) * __fdividef(1.0f, oml);
    red[k] = live ? g_v * dlam : 0.f;
    red[PWC_CB + k] = live ? -g_z * sg - (HARD ? 0.f : z_p * g_v) : 0.f;
    const float g_vin = HARD ? g_v * lam[k] * (1.0f - z_p) : g_v * lam[k];
    if (p.g_v_in && live) p.g_v_in[((size_t)b * 32 + c0 + k) * hw + pix] = g_vin;
    if (p.gI_f32 && live) p.gI_f32[((size_t)b * 32 + c0 + k) * hw + pix] = gI[k];  // head mode: fp32 NCHW for the CUDA-core weight gradient
  }
#pragma unroll
  for (int k = 0; k < PWC_CB; k += 2) split2_bf16(gI[k], gI[k + 1], hi[k >> 1], mid[k >> 1]);
  if (live && !p.gI_f32) {
    *reinterpret_cast<uint4*>(p.gI_hi + ((size_t)b * hw + pix) * 32 + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p.gI_mid + ((size_t)b * hw + pix) * 32 + c0) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
  }
  int entry;
  const float tot = warp_transpose_sum16(red, lane, entry);
  if (lane < 16) atomicAdd(&s_sum[entry], tot);
  __syncthreads();
  if (tid < PWC_CB) {
    const float l = sigmoidf_acc(__ldg(p.leak + c0 + tid));
    if (p.g_leak) atomicAdd(p.g_leak + c0 + tid, s_sum[tid] * l * (1.0f - l));
    if (p.g_thresh && __ldg(p.thresh + c0 + tid) >= 0.01f) atomicAdd(p.g_thresh + c0 + tid, s_sum[PWC_CB + tid]);
  }
}

// (1w) the same neuron backward for a WHOLE BPTT window of a feed-forward cell, time loop inside the kernel: a thread owns one pixel x
//      PWC_CB channels and walks t = T-1 ... 0 with dL/dv carried in registers (it never touches memory), every membrane tensor is
//      read once (v[t-1] is "previous potential" at step t and "new potential" at step t-1) and the per-channel parameter-gradient
//      terms are reduced once per window instead of once per step.  Replaces T launches of the kernel above and the
//      [B,32,H,W] dL/dv round trip between them (north-star: state kept in registers across the inner time loop).
#ifndef EF_PWW_OCC
#define EF_PWW_OCC 4
#endif
template <int SURR, bool HARD>
__global__ void __launch_bounds__(PWC_THREADS, EF_PWW_OCC) lif_bwd_pointwise_window_kernel(const ef_lif_bwd_window_params p) {
  __shared__ float s_sum[2 * PWC_CB];
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t hw = (size_t)p.H * p.W;
  const int b = blockIdx.y;
  const int c0 = (blockIdx.x % PWC_GROUPS) * PWC_CB;
  if (tid < 2 * PWC_CB) s_sum[tid] = 0.f;
  __syncthreads();
  const size_t pix = (size_t)(blockIdx.x / PWC_GROUPS) * PWC_THREADS + tid;
  const bool live = pix < hw;
  const size_t pc = live ? pix : hw - 1;  // out-of-range threads compute on a valid pixel and contribute nothing
  const size_t sv = (size_t)p.B * 32 * hw, sz = (size_t)p.B * hw * 32;  // step strides of the fp32 NCHW / bf16 channels-last tensors
  const size_t ov = ((size_t)b * 32 + c0) * hw + pc, oz = ((size_t)b * hw + pc) * 32 + c0;
  float lam[PWC_CB], oml[PWC_CB], thr[PWC_CB], g_v[PWC_CB], v_n[PWC_CB], red[2 * PWC_CB];
#pragma unroll
  for (int k = 0; k < PWC_CB; ++k) {
    lam[k] = sigmoidf_acc(__ldg(p.leak + c0 + k));
    oml[k] = 1.0f - lam[k];
    thr[k] = fmaxf(__ldg(p.thresh + c0 + k), 0.01f);
    g_v[k] = 0.f;
    red[k] = red[PWC_CB + k] = 0.f;
    v_n[k] = __ldg(p.v + (size_t)(p.T - 1) * sv + ov + k * hw);
  }
  // software pipeline: the loads of step t-1 are issued before step t is computed (one more memory round trip in flight per thread)
  float v_p[PWC_CB], g_o[PWC_CB], v_q[PWC_CB], g_q[PWC_CB];
  uint4 zq = make_uint4(0, 0, 0, 0), zq_q = make_uint4(0, 0, 0, 0);
  auto load_step = [&](int t, float (&vp_)[PWC_CB], float (&go_)[PWC_CB], uint4& z_) {
    z_ = make_uint4(0, 0, 0, 0);
    if (t > 0) z_ = __ldg(reinterpret_cast<const uint4*>(p.z_cl + (size_t)(t - 1) * sz + oz));
    else if (p.z_prev_cl) z_ = __ldg(reinterpret_cast<const uint4*>(p.z_prev_cl + oz));
    const float* vp = t > 0 ? p.v + (size_t)(t - 1) * sv : p.v_prev;
#pragma unroll
    for (int k = 0; k < PWC_CB; ++k) {
      vp_[k] = vp ? __ldg(vp + ov + k * hw) : 0.f;
      go_[k] = __ldg(p.g_out + (size_t)t * sv + ov + k * hw);
    }
  };
  load_step(p.T - 1, v_p, g_o, zq);
  for (int t = p.T - 1; t >= 0; --t) {
    if (t > 0) load_step(t - 1, v_q, g_q, zq_q);
    const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
    uint32_t hi[PWC_CB / 2], mid[PWC_CB / 2];
    float gI[PWC_CB];
#pragma unroll
    for (int k = 0; k < PWC_CB; ++k) {
      const float z_p = (k & 1) ? bf16_hi(zw[k >> 1]) : bf16_lo(zw[k >> 1]);
      const float g_z = g_o[k];
      const float sg = surrogate_grad_fast(SURR, v_n[k] - thr[k], p.act_width);
      const float gv = g_v[k] + g_z * sg;
      gI[k] = oml[k] * gv;
      const float keep = HARD ? v_p[k] * (1.0f - z_p) : v_p[k];
      // d v_out / d lam = (keep - v_out [- z thr]) / (1 - lam): the per-channel constant 1 / (1 - lam) multiplies the SUM at the end
      red[k] += gv * (HARD ? keep - v_n[k] : v_p[k] - v_n[k] - z_p * thr[k]);
      red[PWC_CB + k] += -g_z * sg - (HARD ? 0.f : z_p * gv);
      g_v[k] = HARD ? gv * lam[k] * (1.0f - z_p) : gv * lam[k];  // dL/dv of step t-1, stays in the register
      if (p.gI_f32 && live) p.gI_f32[(size_t)t * sv + ov + k * hw] = gI[k];  // head mode: fp32 NCHW for the CUDA-core weight gradient
      v_n[k] = v_p[k];
    }
#pragma unroll
    for (int k = 0; k < PWC_CB; k += 2) split2_bf16(gI[k], gI[k + 1], hi[k >> 1], mid[k >> 1]);
    if (live && !p.gI_f32) {
      *reinterpret_cast<uint4*>(p.gI_hi + (size_t)t * sz + oz) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(p.gI_mid + (size_t)t * sz + oz) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
    }
#pragma unroll
    for (int k = 0; k < PWC_CB; ++k) v_p[k] = v_q[k], g_o[k] = g_q[k];
    zq = zq_q;
  }
  if (p.g_v_prev && live) {
#pragma unroll
    for (int k = 0; k < PWC_CB; ++k) p.g_v_prev[ov + k * hw] = g_v[k];
  }
#pragma unroll
  for (int k = 0; k < PWC_CB; ++k) red[k] = red[k] / oml[k];
  if (!live) {
#pragma unroll
    for (int k = 0; k < 2 * PWC_CB; ++k) red[k] = 0.f;
  }
  int entry;
  const float tot = warp_transpose_sum16(red, lane, entry);
  if (lane < 16) atomicAdd(&s_sum[entry], tot);
  __syncthreads();
  if (tid < PWC_CB) {
    const float l = sigmoidf_acc(__ldg(p.leak + c0 + tid));
    if (p.g_leak) atomicAdd(p.g_leak + c0 + tid, s_sum[tid] * l * (1.0f - l));
    if (p.g_thresh && __ldg(p.thresh + c0 + tid) >= 0.01f) atomicAdd(p.g_thresh + c0 + tid, s_sum[PWC_CB + tid]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// (1b) weight gradient of the head layer: g_w[co][ci][tap] += sum_{b,y,x} x[b,ci,y+dy-1,x+dx-1] * g_I[b,co,y,x], Cin <= 8 fractional
//      fp32 inputs (not a tensor-core shape).  A block owns 32 x 8-pixel tiles; slice s (64 threads) owns tile row s; thread j of
//      a slice owns one (ci, tap) pair and all 32 output channels in registers: per pixel ONE shifted input value and eight
//      16-byte broadcast loads of g_I feed 32 FMAs.  Accumulators live across the block's tiles; one reduction at the end.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HW_THREADS = 512, HW_TW = 32, HW_TH = 8, HW_MAXC = 8;

__global__ void __launch_bounds__(HW_THREADS, 2) lif_head_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g_I, float* __restrict__ g_w,
                                                                     int B, int Cin, int H, int W, int tiles_x, int tiles_y, int n_tiles) {
  __shared__ __align__(16) float s_g[HW_TW * HW_TH * 32];                  // [pixel][32 co], 16-byte groups XOR-swizzled by (pixel & 7)
  __shared__ float s_x[HW_MAXC * (HW_TH + 2) * (HW_TW + 2)];
  const int tid = threadIdx.x, slice = tid >> 6, j = tid & 63;
  const int n_pairs = Cin * 9;
  const bool active = j < n_pairs;
  const int ci = active ? j / 9 : 0, tap = active ? j % 9 : 0, dy = tap / 3, dx = tap % 3;
  const size_t plane = (size_t)H * W;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  constexpr int XW = HW_TW + 2, XH = HW_TH + 2;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / (tiles_x * tiles_y), r = tile % (tiles_x * tiles_y);
    const int x0 = (r % tiles_x) * HW_TW, y0 = (r / tiles_x) * HW_TH;
    __syncthreads();
    for (int i = tid; i < 32 * HW_TH * HW_TW; i += HW_THREADS) {  // coalesced along x; zero outside the image
      const int px = i % HW_TW, py = (i / HW_TW) % HW_TH, co = i / (HW_TW * HW_TH);
      const int y = y0 + py, xx = x0 + px, pix = py * HW_TW + px;
      const float v = (y < H && xx < W) ? __ldg(g_I + ((size_t)b * 32 + co) * plane + (size_t)y * W + xx) : 0.f;
      s_g[pix * 32 + (((co >> 2) ^ (pix & 7)) << 2) + (co & 3)] = v;
    }
    for (int i = tid; i < Cin * XH * XW; i += HW_THREADS) {
      const int c = i / (XH * XW), q = i % (XH * XW), y = y0 - 1 + q / XW, xx = x0 - 1 + q % XW;
      s_x[i] = (y >= 0 && y < H && xx >= 0 && xx < W) ? __ldg(x + ((size_t)b * Cin + c) * plane + (size_t)y * W + xx) : 0.f;
    }
    __syncthreads();
    if (active) {
      const float* xr = s_x + (ci * XH + slice + dy) * XW + dx;
#pragma unroll 4
      for (int px = 0; px < HW_TW; ++px) {
        const float xv = xr[px];
        const int pix = slice * HW_TW + px;
        const float4* gr = reinterpret_cast<const float4*>(s_g + pix * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 g4 = gr[q ^ (pix & 7)];
          acc[4 * q + 0] = fmaf(xv, g4.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(xv, g4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, g4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, g4.w, acc[4 * q + 3]);
        }
      }
    }
  }
  // reduce the 8 slices in shared memory (reusing s_g as [32 co][64 pairs]: conflict-free, one slice at a time, no shared
  // atomics -- float atomicAdd on shared memory is a CAS loop and serialises badly), then one global atomic per weight
  for (int sl = 0; sl < HW_THREADS / 64; ++sl) {
    __syncthreads();
    if (slice == sl && active) {
#pragma unroll
      for (int c = 0; c < 32; ++c) s_g[c * 64 + j] = (sl == 0 ? 0.f : s_g[c * 64 + j]) + acc[c];
    }
  }
  __syncthreads();
  for (int i = tid; i < 32 * 64; i += HW_THREADS) {
    const int pair = i & 63, co = i >> 6;
    if (pair < n_pairs) atomicAdd(g_w + ((size_t)co * Cin + pair / 9) * 9 + pair % 9, s_g[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// (2) data gradient on the tensor cores
// ---------------------------------------------------------------------------------------------------------------------
constexpr int DG_TH = 16, DG_TW = 8;
constexpr int DG_ROW_BYTES = (DG_TW + 8) * PIX_BYTES;       // 1024
constexpr int DG_TILE_BYTES = (DG_TH + 2) * DG_ROW_BYTES;   // 18432: one operand tile (hi or mid)
constexpr int DG_EPI_WARPS = 8, DG_THREADS = 32 * (2 + DG_EPI_WARPS);
constexpr int DG_TMEM_COLS = 256;

struct DgLayout {
  int nb, wtap_bytes, w_bytes, stage_off, stage_bytes, bar_off, total, nstage, acc_cols;
};
__host__ __device__ inline DgLayout dg_layout(bool rec) {
  DgLayout l;
  l.nb = rec ? 4 : 2;                       // 32-row groups stacked along N: [ff_hi, (rec_hi), ff_mid, (rec_mid)]
  l.acc_cols = l.nb * 32;
  l.wtap_bytes = l.nb * 32 * PIX_BYTES;     // one tap: [nb*32 n][32 k] bf16
  l.w_bytes = 9 * l.wtap_bytes;
  l.stage_off = l.w_bytes;
  l.stage_bytes = 2 * DG_TILE_BYTES;        // g_I hi tile + g_I mid tile
  l.nstage = (227 * 1024 - 1280 - l.w_bytes) / l.stage_bytes;
  if (l.nstage > 4) l.nstage = 4;
  l.bar_off = l.stage_off + l.nstage * l.stage_bytes;
  l.total = l.bar_off + 256 + 1024;
  return l;
}

struct DgParams {
  int B, H, W, tiles_x, tiles_y, n_tiles, has_rec;
  const uint16_t* w_bwd;   // prepared by split_weights_bwd_kernel
  float* g_x;              // [B,32,H,W]
  float* g_z_in;           // [B,32,H,W] or NULL
};

template <bool REC>
__global__ void __launch_bounds__(DG_THREADS, 1)
lif_dgrad_tc_kernel(const DgParams p, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_mid) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const DgLayout L = dg_layout(REC);
  const int NST = L.nstage;
  constexpr int ACC = REC ? 128 : 64;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar_w = s_base + L.bar_off;
  auto bar_full = [&](int s) { return bar_w + 8u * (1 + s); };
  auto bar_empty = [&](int s) { return bar_w + 8u * (1 + NST + s); };
  auto bar_accf = [&](int a) { return bar_w + 8u * (1 + 2 * NST + a); };
  auto bar_acce = [&](int a) { return bar_w + 8u * (3 + 2 * NST + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * (5 + 2 * NST));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);  // only the MMA warp consumes a stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);
      mbar_init(bar_acce(a), DG_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(DG_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      mbar_expect_tx(bar_w, L.w_bytes);
      for (int off = 0; off < L.w_bytes; off += 9216) bulk_load_1d(s_base + off, reinterpret_cast<const uint8_t*>(p.w_bwd) + off, 9216, bar_w);
      for (int it = 0; it < n_my; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img, ty = r / p.tiles_x;
        const int y0 = ty * DG_TH, x0 = (r - ty * p.tiles_x) * DG_TW;
        const int s = it % NST;
        mbar_wait(bar_empty(s), ((it / NST) & 1) ^ 1);
        const uint32_t st = s_base + L.stage_off + s * L.stage_bytes;
        mbar_expect_tx(bar_full(s), 2 * DG_TILE_BYTES);
        tma_load_4d(st, &map_hi, bar_full(s), 0, x0 - 1, y0 - 1, b);
        tma_load_4d(st + DG_TILE_BYTES, &map_mid, bar_full(s), 0, x0 - 1, y0 - 1, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer
      mbar_wait(bar_w, 0);
      const uint64_t bw = umma_desc_sw64(s_base, ATOM_BYTES);
      for (int it = 0; it < n_my; ++it) {
        const int s = it % NST, a = it & 1;
        mbar_wait(bar_acce(a), ((it >> 1) & 1) ^ 1);
        mbar_wait(bar_full(s), (it / NST) & 1);
        tc_fence_after();
        const uint32_t st = s_base + L.stage_off + s * L.stage_bytes;
        const uint32_t d_tmem = tmem_base + a * ACC;
        const uint64_t a_hi = umma_desc_sw64(st, DG_ROW_BYTES), a_mid = umma_desc_sw64(st + DG_TILE_BYTES, DG_ROW_BYTES);
        constexpr int WTAP16 = (REC ? 4 : 2) * 32 * PIX_BYTES / 16;  // tap stride of the weight image in 16-byte units
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t aoff = (uint64_t)((tap / 3) * (DG_ROW_BYTES / 16) + (tap % 3) * (PIX_BYTES / 16) + ks * 2);
            const uint64_t boff = (uint64_t)(tap * WTAP16 + ks * 2);
            // hi term of g_I against [w_hi | w_mid]; mid term of g_I against w_hi only (first half of the stacked rows)
            umma_bf16<umma_idesc(ACC)>(d_tmem, a_hi + aoff, bw + boff, (tap | ks) != 0);
            umma_bf16<umma_idesc(ACC / 2)>(d_tmem, a_mid + aoff, bw + boff, 1u);
          }
        }
        umma_commit(bar_empty(s));
        umma_commit(bar_accf(a));
      }
    }
  } else {
    // ---- epilogue: 8 warps = 4 lane quadrants x 2 channel halves; g = D[hi cols] + D[mid cols] -> fp32 NCHW
    const int q = warp & 3, hsel = (warp - 2) >> 2, m = q * 32 + lane;
    const int ph_ = m >> 3, pw_ = m & 7, c0 = 16 * hsel;
    const size_t plane = (size_t)p.H * p.W;
    for (int it = 0; it < n_my; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / tiles_per_img, r = tile - b * tiles_per_img, ty = r / p.tiles_x;
      const int gy = ty * DG_TH + ph_, gx = (r - ty * p.tiles_x) * DG_TW + pw_;
      const int a = it & 1;
      mbar_wait(bar_accf(a), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + a * ACC + c0 + ((uint32_t)(q * 32) << 16);
      uint32_t x_hi[16], x_mid[16], z_hi[16], z_mid[16];
      tmem_ld16(tacc, x_hi);
      tmem_ld16(tacc + ACC / 2, x_mid);
      if (REC) {
        tmem_ld16(tacc + 32, z_hi);
        tmem_ld16(tacc + ACC / 2 + 32, z_mid);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce(a));
      if (gy < p.H && gx < p.W) {
        const size_t o = ((size_t)b * 32 + c0) * plane + (size_t)gy * p.W + gx;
#pragma unroll
        for (int j = 0; j < 16; ++j) p.g_x[o + j * plane] = __uint_as_float(x_hi[j]) + __uint_as_float(x_mid[j]);
        if (REC && p.g_z_in) {
#pragma unroll
          for (int j = 0; j < 16; ++j) p.g_z_in[o + j * plane] = __uint_as_float(z_hi[j]) + __uint_as_float(z_mid[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(DG_TMEM_COLS) : "memory");
  }
}

// Weight image for the data gradient: per halo-window tap tapH = (2-dy)*3 + (2-dx) a block [nb*32 rows n'][32 k = co],
// row groups [ff_hi | rec_hi | ff_mid | rec_mid] (without the rec groups for a feed-forward cell), n' % 32 = ci,
// value = split(w[co][ci][8 - tapH]); rows of 64 B, 8-row atoms, 64-byte swizzle (chunk ^= (row >> 1) & 3).
__global__ void split_weights_bwd_kernel(const float* __restrict__ w_ff, const float* __restrict__ w_rec, uint16_t* __restrict__ out) {
  const int nconv = w_rec ? 2 : 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over conv, co, ci, tap
  if (i >= nconv * 32 * 32 * 9) return;
  const int tap = i % 9, ci = (i / 9) % 32, co = (i / 288) % 32, cv = i / 9216;
  const float w = (cv == 0 ? w_ff : w_rec)[(co * 32 + ci) * 9 + tap];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 mid = __float2bfloat16_rn(w - __bfloat162float(hi));
  const __nv_bfloat16 parts[2] = {hi, mid};
  const int tapH = 8 - tap, nb = nconv * 2;
  for (int sp = 0; sp < 2; ++sp) {
    const int nn = (sp * nconv + cv) * 32 + ci, r = nn & 7;
    const int chunk = (co >> 3) ^ ((r >> 1) & 3);
    out[(size_t)tapH * nb * 32 * 32 + (nn >> 3) * 256 + r * 32 + chunk * 8 + (co & 7)] = *reinterpret_cast<const uint16_t*>(&parts[sp]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// (3) weight gradient, CUDA cores, channels-last operands:
//     g_w[co,ci,dy,dx] += sum_{b,y,x} in[b,y+dy-1,x+dx-1,ci] * (gI_hi + gI_mid)[b,y,x,co]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WGC_THREADS = 128, WGC_GP = 257;

__global__ void __launch_bounds__(WGC_THREADS) conv_wgrad_cl_kernel(const uint16_t* __restrict__ in_cl, const uint16_t* __restrict__ gI_hi,
                                                                    const uint16_t* __restrict__ gI_mid, float* __restrict__ g_w, int B, int H,
                                                                    int W) {
  __shared__ float s_g[32 * WGC_GP];
  __shared__ float s_x[8 * 18 * 18];
  const int tid = threadIdx.x, co_l = tid & 31, cp = tid >> 5;
  const int b = blockIdx.z, ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16;
  for (int i = tid; i < 256 * 4; i += WGC_THREADS) {  // (pixel, 8-channel group) -> 16-byte loads of both terms
    const int r = i >> 2, g = i & 3, yy = oy0 + (r >> 4), xx = ox0 + (r & 15);
    uint4 h = make_uint4(0, 0, 0, 0), m = h;
    if (yy < H && xx < W) {
      const size_t o = (((size_t)b * H + yy) * W + xx) * 32 + g * 8;
      h = __ldg(reinterpret_cast<const uint4*>(gI_hi + o));
      m = __ldg(reinterpret_cast<const uint4*>(gI_mid + o));
    }
    const uint32_t hw_[4] = {h.x, h.y, h.z, h.w}, mw_[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s_g[(g * 8 + 2 * k) * WGC_GP + r] = bf16_lo(hw_[k]) + bf16_lo(mw_[k]);
      s_g[(g * 8 + 2 * k + 1) * WGC_GP + r] = bf16_hi(hw_[k]) + bf16_hi(mw_[k]);
    }
  }
  for (int ci0 = 0; ci0 < 32; ci0 += 8) {
    __syncthreads();
    for (int i = tid; i < 324; i += WGC_THREADS) {
      const int y = oy0 - 1 + i / 18, x = ox0 - 1 + i % 18;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(reinterpret_cast<const uint4*>(in_cl + (((size_t)b * H + y) * W + x) * 32 + ci0));
      const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s_x[(2 * k) * 324 + i] = bf16_lo(vw[k]);
        s_x[(2 * k + 1) * 324 + i] = bf16_hi(vw[k]);
      }
    }
    __syncthreads();
    float acc[2][9];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 9; ++j) acc[i][j] = 0.f;
    const float* sx0 = s_x + (cp * 2) * 324;
    const float* sx1 = sx0 + 324;
    const float* sg = s_g + co_l * WGC_GP;
    for (int y = 0; y < 16; ++y) {
      float w0[3][3], w1[3][3];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        w0[dy][1] = sx0[(y + dy) * 18 + 0];
        w0[dy][2] = sx0[(y + dy) * 18 + 1];
        w1[dy][1] = sx1[(y + dy) * 18 + 0];
        w1[dy][2] = sx1[(y + dy) * 18 + 1];
      }
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const float g = sg[y * 16 + x];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          w0[dy][0] = w0[dy][1];
          w0[dy][1] = w0[dy][2];
          w0[dy][2] = sx0[(y + dy) * 18 + x + 2];
          w1[dy][0] = w1[dy][1];
          w1[dy][1] = w1[dy][2];
          w1[dy][2] = sx1[(y + dy) * 18 + x + 2];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            acc[0][dy * 3 + dx] = fmaf(w0[dy][dx], g, acc[0][dy * 3 + dx]);
            acc[1][dy * 3 + dx] = fmaf(w1[dy][dx], g, acc[1][dy * 3 + dx]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float* dst = g_w + ((size_t)co_l * 32 + ci0 + cp * 2 + i) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(dst + t, acc[i][t]);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// (4) weight gradient on the tensor cores:  g_w[co,ci,dy,dx] = sum_{b,y,x} in[b,y+dy-1,x+dx-1,ci] * (gI_hi + gI_mid)[b,y,x,co]
//     The reduction runs over pixels, so both operands are MN-major (channels are the contiguous dimension of the
//     channels-last tensors).  Per 16x8-pixel tile the CTA holds the padded, 64B-swizzled halo tile of the input
//     [18 rows][16 px][64 B] and the plain tiles of g_I (hi, mid) [16 rows][8 px][64 B].  One MMA has
//       A: M = 128 = 4 pixel shifts x 32 ci  (LBO = 64 B = one pixel to the right; shift 3 is unused),
//          K = 16 = 2 tile rows x 8 pixels  (SBO = 1024 B = one halo-tile row), start address = row 2*ks + dy
//       B: N = 64 = [gI_hi | gI_mid] x 32 co (LBO = distance between the two tiles), K as for A (SBO = 512 B)
//     and accumulates into D[conv][dy] (64 columns each), which stays in tensor memory for all tiles of the CTA.  The epilogue
//     folds hi + mid and writes / adds the CTA's slice of wg_partial [conv][cta][dy][dx][ci][co].
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WG_TH = 16, WG_TW = 8;
constexpr int WG_XROW = (WG_TW + 8) * PIX_BYTES;        // 1024
constexpr int WG_XTILE = (WG_TH + 2) * WG_XROW;         // 18432
constexpr int WG_GTILE = WG_TH * WG_TW * PIX_BYTES;     // 8192
constexpr int WG_THREADS = 32 * 6;                      // TMA, MMA, 4 epilogue warps (one per TMEM lane quadrant)
constexpr int WG_NST = 4;
constexpr int WG_SLICE = 9 * 32 * 32;                   // floats per (conv, cta)

struct WgParams {
  int B, H, W, tiles_x, tiles_y, n_tiles, accumulate;
  float* partial;
};

template <bool REC>
__global__ void __launch_bounds__(WG_THREADS, 1)
lif_wgrad_tc_kernel(const WgParams p, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z,
                    const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_mid) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NCONV = REC ? 2 : 1;
  constexpr int G_OFF = NCONV * WG_XTILE;
  constexpr int STAGE = G_OFF + 2 * WG_GTILE;
  constexpr int BAR_OFF = WG_NST * STAGE;
  constexpr uint32_t TMEM_COLS = REC ? 512 : 256;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar_w = s_base + BAR_OFF;
  auto bar_full = [&](int s) { return bar_w + 8u * s; };
  auto bar_empty = [&](int s) { return bar_w + 8u * (WG_NST + s); };
  const uint32_t bar_done = bar_w + 8u * (2 * WG_NST);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 8 * (2 * WG_NST + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_NST; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      for (int it = 0; it < n_my; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img, ty = r / p.tiles_x;
        const int y0 = ty * WG_TH, x0 = (r - ty * p.tiles_x) * WG_TW;
        const int s = it % WG_NST;
        mbar_wait(bar_empty(s), ((it / WG_NST) & 1) ^ 1);
        const uint32_t st = s_base + s * STAGE;
        mbar_expect_tx(bar_full(s), STAGE);
        tma_load_4d(st, &map_x, bar_full(s), 0, x0 - 1, y0 - 1, b);
        if (REC) tma_load_4d(st + WG_XTILE, &map_z, bar_full(s), 0, x0 - 1, y0 - 1, b);
        tma_load_4d(st + G_OFF, &map_hi, bar_full(s), 0, x0, y0, b);
        tma_load_4d(st + G_OFF + WG_GTILE, &map_mid, bar_full(s), 0, x0, y0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer
      constexpr uint32_t IDESC = umma_idesc(64, true, true);
      for (int it = 0; it < n_my; ++it) {
        const int s = it % WG_NST;
        mbar_wait(bar_full(s), (it / WG_NST) & 1);
        tc_fence_after();
        const uint32_t st = s_base + s * STAGE;
        const uint64_t bdesc = umma_desc_sw64_mn(st + G_OFF, WG_GTILE, ATOM_BYTES);
#pragma unroll
        for (int cv = 0; cv < NCONV; ++cv) {
          const uint64_t adesc = umma_desc_sw64_mn(st + cv * WG_XTILE, PIX_BYTES, WG_XROW);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int ks = 0; ks < WG_TH / 2; ++ks)
              umma_bf16<IDESC>(tmem_base + (cv * 3 + dy) * 64, adesc + (uint64_t)((2 * ks + dy) * (WG_XROW / 16)),
                               bdesc + (uint64_t)(ks * (2 * ATOM_BYTES / 16)), (it | ks) != 0);
          }
        }
        umma_commit(bar_empty(s));
      }
      umma_commit(bar_done);
    }
  } else {
    // ---- epilogue (once): thread = accumulator row m = (dx shift, ci); 64 columns = [hi | mid] x co
    const int q = warp & 3;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    if (q < 3) {
#pragma unroll 1
      for (int a = 0; a < NCONV * 3; ++a) {
        const int cv = a / 3, dy = a - cv * 3;
        const uint32_t tacc = tmem_base + a * 64 + ((uint32_t)(q * 32) << 16);
        uint32_t h0[16], h1[16], m0[16], m1[16];
        tmem_ld16(tacc, h0);
        tmem_ld16(tacc + 16, h1);
        tmem_ld16(tacc + 32, m0);
        tmem_ld16(tacc + 48, m1);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)cv * gridDim.x + blockIdx.x) * WG_SLICE + ((dy * 3 + q) * 32 + lane) * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 lo = make_float4(__uint_as_float(h0[4 * j]) + __uint_as_float(m0[4 * j]), __uint_as_float(h0[4 * j + 1]) + __uint_as_float(m0[4 * j + 1]),
                                  __uint_as_float(h0[4 * j + 2]) + __uint_as_float(m0[4 * j + 2]), __uint_as_float(h0[4 * j + 3]) + __uint_as_float(m0[4 * j + 3]));
          float4 hi = make_float4(__uint_as_float(h1[4 * j]) + __uint_as_float(m1[4 * j]), __uint_as_float(h1[4 * j + 1]) + __uint_as_float(m1[4 * j + 1]),
                                  __uint_as_float(h1[4 * j + 2]) + __uint_as_float(m1[4 * j + 2]), __uint_as_float(h1[4 * j + 3]) + __uint_as_float(m1[4 * j + 3]));
          if (p.accumulate) {
            const float4 a0 = dst[j], a1 = dst[4 + j];
            lo.x += a0.x, lo.y += a0.y, lo.z += a0.z, lo.w += a0.w;
            hi.x += a1.x, hi.y += a1.y, hi.z += a1.z, hi.w += a1.w;
          }
          dst[j] = lo;
          dst[4 + j] = hi;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// g_w[co,ci,dy,dx] += sum over CTAs of partial[conv][cta][dy][dx][ci][co], CTAs in a fixed order (bit-reproducible).
// Block = 32 co x 8 CTA groups for one (conv, dy, dx, ci).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ g_w_ff,
                                                           float* __restrict__ g_w_rec) {
  __shared__ float s[8][32];
  const int co = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int row = blockIdx.x % 288, cv = blockIdx.x / 288;  // row = (dy*3 + dx)*32 + ci
  const float* src = partial + (size_t)cv * n_cta * WG_SLICE + row * 32 + co;
  float acc = 0.f;
  for (int c = grp; c < n_cta; c += 8) acc += __ldg(src + (size_t)c * WG_SLICE);
  s[grp][co] = acc;
  __syncthreads();
  if (grp == 0) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += s[g][co];
    float* g_w = cv == 0 ? g_w_ff : g_w_rec;
    if (g_w) g_w[(co * 32 + (row & 31)) * 9 + (row >> 5)] += t;
  }
}

// Head layer on split inputs (ef_pack_split_cl): the three slots of input channel c all multiply w[.,c], so the weight gradient of
// w[co][c][tap] is the sum of the rows s*SL + c, s = 0..2, of the tensor-core result.  One block per (tap, slot row).
__global__ void __launch_bounds__(256) wgrad_reduce_head_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ g_w, int Cin, int SL) {
  __shared__ float s[8][32];
  const int co = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int row = blockIdx.x;  // (dy*3 + dx)*32 + k
  const int k = row & 31, c = k % SL, slot = k / SL;
  if (slot >= 3 || c >= Cin) return;
  const float* src = partial + row * 32 + co;
  float acc = 0.f;
  for (int i = grp; i < n_cta; i += 8) acc += __ldg(src + (size_t)i * WG_SLICE);
  s[grp][co] = acc;
  __syncthreads();
  if (grp == 0) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += s[g][co];
    atomicAdd(g_w + ((size_t)co * Cin + c) * 9 + (row >> 5), t);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// (4g) the same weight gradient for GENERAL channel counts (multiples of 32; the U-Net family): the channels-last tensors have
//      Cin / Cout channels and a CTA works on one PAIR of 32-channel blocks (ci block, co block) -- the block offsets are the channel
//      coordinates of its TMA boxes -- over its share ("chunk") of the pixel tiles: blockIdx.x = pair * chunks + chunk.  Layers with many
//      block pairs and few pixels (512 x 512 channels at 16 x 16) spread over the SMs by pair, wide layers with few channels by chunk.
//      partial[pair][chunk][dy][dx][ci][co]; wgrad_reduce_g_kernel sums the chunks in order (bit-reproducible) into g_w.
// ---------------------------------------------------------------------------------------------------------------------
struct WgGParams {
  int B, H, W, tiles_x, tiles_y, n_tiles, chunks, n_co_blk;
  float* partial;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
lif_wgrad_tcg_kernel(const WgGParams p, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_hi,
                     const __grid_constant__ CUtensorMap map_mid) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int G_OFF = WG_XTILE;
  constexpr int STAGE = G_OFF + 2 * WG_GTILE;
  constexpr int BAR_OFF = WG_NST * STAGE;
  constexpr uint32_t TMEM_COLS = 256;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar_w = s_base + BAR_OFF;
  auto bar_full = [&](int s) { return bar_w + 8u * s; };
  auto bar_empty = [&](int s) { return bar_w + 8u * (WG_NST + s); };
  const uint32_t bar_done = bar_w + 8u * (2 * WG_NST);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 8 * (2 * WG_NST + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_NST; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int pair = (int)blockIdx.x / p.chunks, chunk = (int)blockIdx.x - pair * p.chunks;
  const int ci0 = (pair / p.n_co_blk) * 32, co0 = (pair % p.n_co_blk) * 32;
  const int n_my = (p.n_tiles - chunk + p.chunks - 1) / p.chunks;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      for (int it = 0; it < n_my; ++it) {
        const int tile = chunk + it * p.chunks;
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img, ty = r / p.tiles_x;
        const int y0 = ty * WG_TH, x0 = (r - ty * p.tiles_x) * WG_TW;
        const int s = it % WG_NST;
        mbar_wait(bar_empty(s), ((it / WG_NST) & 1) ^ 1);
        const uint32_t st = s_base + s * STAGE;
        mbar_expect_tx(bar_full(s), STAGE);
        tma_load_4d(st, &map_x, bar_full(s), ci0, x0 - 1, y0 - 1, b);
        tma_load_4d(st + G_OFF, &map_hi, bar_full(s), co0, x0, y0, b);
        tma_load_4d(st + G_OFF + WG_GTILE, &map_mid, bar_full(s), co0, x0, y0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer (operand layouts as in lif_wgrad_tc_kernel)
      constexpr uint32_t IDESC = umma_idesc(64, true, true);
      for (int it = 0; it < n_my; ++it) {
        const int s = it % WG_NST;
        mbar_wait(bar_full(s), (it / WG_NST) & 1);
        tc_fence_after();
        const uint32_t st = s_base + s * STAGE;
        const uint64_t bdesc = umma_desc_sw64_mn(st + G_OFF, WG_GTILE, ATOM_BYTES);
        const uint64_t adesc = umma_desc_sw64_mn(st, PIX_BYTES, WG_XROW);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int ks = 0; ks < WG_TH / 2; ++ks)
            umma_bf16<IDESC>(tmem_base + dy * 64, adesc + (uint64_t)((2 * ks + dy) * (WG_XROW / 16)), bdesc + (uint64_t)(ks * (2 * ATOM_BYTES / 16)),
                             (it | ks) != 0);
        }
        umma_commit(bar_empty(s));
      }
      umma_commit(bar_done);
    }
  } else {
    // ---- epilogue (once): thread = accumulator row m = (dx shift, ci); 64 columns = [hi | mid] x co
    const int q = warp & 3;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    if (q < 3) {
#pragma unroll 1
      for (int dy = 0; dy < 3; ++dy) {
        const uint32_t tacc = tmem_base + dy * 64 + ((uint32_t)(q * 32) << 16);
        uint32_t h0[16], h1[16], m0[16], m1[16];
        tmem_ld16(tacc, h0);
        tmem_ld16(tacc + 16, h1);
        tmem_ld16(tacc + 32, m0);
        tmem_ld16(tacc + 48, m1);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(p.partial + (size_t)blockIdx.x * WG_SLICE + ((dy * 3 + q) * 32 + lane) * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dst[j] = make_float4(__uint_as_float(h0[4 * j]) + __uint_as_float(m0[4 * j]), __uint_as_float(h0[4 * j + 1]) + __uint_as_float(m0[4 * j + 1]),
                               __uint_as_float(h0[4 * j + 2]) + __uint_as_float(m0[4 * j + 2]), __uint_as_float(h0[4 * j + 3]) + __uint_as_float(m0[4 * j + 3]));
          dst[4 + j] = make_float4(__uint_as_float(h1[4 * j]) + __uint_as_float(m1[4 * j]), __uint_as_float(h1[4 * j + 1]) + __uint_as_float(m1[4 * j + 1]),
                                   __uint_as_float(h1[4 * j + 2]) + __uint_as_float(m1[4 * j + 2]), __uint_as_float(h1[4 * j + 3]) + __uint_as_float(m1[4 * j + 3]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// g_w[co0 + co][ci_off + ci0 + ci][dy][dx] += sum over the chunks (in order) of partial[pair][chunk][dy][dx][ci][co].
// Block = one pair x 8 output channels: the sums are transposed through shared memory so that g_w is updated in runs of 288 contiguous
// floats (the 32 ci x 9 taps of one output channel) instead of one 4-byte access per 32-byte sector.
__global__ void __launch_bounds__(256) wgrad_reduce_g_kernel(const float* __restrict__ partial, int chunks, int n_co_blk, float* __restrict__ g_w,
                                                             int cin_total, int ci_off) {
  __shared__ float s[8][289];
  const int pair = blockIdx.x >> 2, cg = (blockIdx.x & 3) * 8;
  const int ci0 = (pair / n_co_blk) * 32, co0 = (pair % n_co_blk) * 32;
  const int co_l = threadIdx.x & 7;
  const float* src = partial + (size_t)pair * chunks * WG_SLICE + cg + co_l;
#pragma unroll 3
  for (int k = 0; k < 9; ++k) {
    const int row = k * 32 + (threadIdx.x >> 3);  // (dy*3 + dx)*32 + ci
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) acc += __ldg(src + (size_t)c * WG_SLICE + row * 32);
    s[co_l][(row & 31) * 9 + (row >> 5)] = acc;
  }
  __syncthreads();
  const int co = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dst = g_w + ((size_t)(co0 + cg + co) * cin_total + ci_off + ci0) * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) dst[k * 32 + lane] += s[co][k * 32 + lane];
}

inline int wgg_chunks(int B, int H, int W, int pairs) {
  const int n_tiles = B * cdiv(W, WG_TW) * cdiv(H, WG_TH);
  int chunks = cdiv(2 * 148, pairs);
  return chunks < n_tiles ? chunks : n_tiles;
}

inline int wg_n_sms() {
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return n_sms;
}
inline int wg_grid(int B, int H, int W) {
  const int n_tiles = B * cdiv(W, WG_TW) * cdiv(H, WG_TH), n = wg_n_sms();
  return n_tiles < n ? n_tiles : n;
}

}  // namespace ef

extern "C" int64_t ef_lif_wgrad_partial_elems(int32_t B, int32_t H, int32_t W, int32_t has_rec) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)(has_rec ? 2 : 1) * ef::wg_grid(B, H, W) * ef::WG_SLICE;
}

extern "C" int64_t ef_split_weights_bwd_elems(int32_t has_rec) { return (int64_t)ef::dg_layout(has_rec != 0).w_bytes / 2; }

extern "C" int ef_split_weights_bwd(const float* w_ff, const float* w_rec, uint16_t* out, void* stream) {
  using namespace ef;
  EF_REQUIRE(w_ff && out, EF_ENULL, "ef_split_weights_bwd: NULL tensor");
  const int nconv = w_rec ? 2 : 1;
  split_weights_bwd_kernel<<<cdiv(nconv * 9216, 256), 256, 0, as_stream(stream)>>>(w_ff, w_rec, out);
  return check_launch("split_weights_bwd_kernel");
}

namespace ef {

// ---- launch helpers shared by the per-step and the per-window entry points --------------------------------------------------------
template <int DUMMY = 0>
static int run_pointwise_step(const ef_lif_bwd_tc_params& p, cudaStream_t st) {
  const int hw = p.H * p.W;
  const dim3 pgrid(cdiv(hw, PWC_THREADS) * PWC_GROUPS, p.B);
#define EF_PW(S_, H_) lif_bwd_pointwise_cl_kernel<S_, H_><<<pgrid, PWC_THREADS, 0, st>>>(p)
  switch (p.surrogate * 2 + (p.hard_reset ? 1 : 0)) {
    case 0: EF_PW(EF_ARCTAN, false); break;
    case 1: EF_PW(EF_ARCTAN, true); break;
    case 2: EF_PW(EF_SUPERSPIKE, false); break;
    case 3: EF_PW(EF_SUPERSPIKE, true); break;
    case 4: EF_PW(EF_TRIANGLE, false); break;
    case 5: EF_PW(EF_TRIANGLE, true); break;
    case 6: EF_PW(EF_MULTIGAUSS, false); break;
    case 7: EF_PW(EF_MULTIGAUSS, true); break;
    default: return fail(EF_EINVAL, "ef_lif_bwd_tc: bad surrogate %d", p.surrogate);
  }
#undef EF_PW
  return check_launch("lif_bwd_pointwise_cl_kernel");
}

static int run_pointwise_window(const ef_lif_bwd_window_params& p, cudaStream_t st) {
  const int hw = p.H * p.W;
  const dim3 pgrid(cdiv(hw, PWC_THREADS) * PWC_GROUPS, p.B);
#define EF_PW(S_, H_) lif_bwd_pointwise_window_kernel<S_, H_><<<pgrid, PWC_THREADS, 0, st>>>(p)
  switch (p.surrogate * 2 + (p.hard_reset ? 1 : 0)) {
    case 0: EF_PW(EF_ARCTAN, false); break;
    case 1: EF_PW(EF_ARCTAN, true); break;
    case 2: EF_PW(EF_SUPERSPIKE, false); break;
    case 3: EF_PW(EF_SUPERSPIKE, true); break;
    case 4: EF_PW(EF_TRIANGLE, false); break;
    case 5: EF_PW(EF_TRIANGLE, true); break;
    case 6: EF_PW(EF_MULTIGAUSS, false); break;
    case 7: EF_PW(EF_MULTIGAUSS, true); break;
    default: return fail(EF_EINVAL, "ef_lif_bwd_window: bad surrogate %d", p.surrogate);
  }
#undef EF_PW
  return check_launch("lif_bwd_pointwise_window_kernel");
}

static int run_head_wgrad(const float* x_f32, const float* gI_f32, float* g_w_ff, int B, int Cin, int H, int W, cudaStream_t st) {
  const int n_sms = wg_n_sms();
  const int tiles_x = cdiv(W, HW_TW), tiles_y = cdiv(H, HW_TH), n_tiles = tiles_x * tiles_y * B;
  const int grid = n_tiles < 2 * n_sms ? n_tiles : 2 * n_sms;
  lif_head_wgrad_kernel<<<grid, HW_THREADS, 0, st>>>(x_f32, gI_f32, g_w_ff, B, Cin, H, W, tiles_x, tiles_y, n_tiles);
  return check_launch("lif_head_wgrad_kernel");
}

// data gradient of B images: g_x (and g_z_in of a recurrent cell) = transposed conv of g_I = gI_hi + gI_mid with the weights
static int run_dgrad(const uint16_t* gI_hi, const uint16_t* gI_mid, const uint16_t* w_bwd, bool rec, float* g_x, float* g_z_in, int B, int H, int W,
                     cudaStream_t st) {
  const int n_sms = wg_n_sms();
  DgParams q;
  q.B = B, q.H = H, q.W = W, q.has_rec = rec;
  q.tiles_x = cdiv(W, DG_TW), q.tiles_y = cdiv(H, DG_TH), q.n_tiles = B * q.tiles_x * q.tiles_y;
  q.w_bwd = w_bwd, q.g_x = g_x, q.g_z_in = rec ? g_z_in : nullptr;
  CUtensorMap mh, mm;
  int rc;
  if ((rc = get_map(gI_hi, B, H, W, DG_TH + 2, DG_TW + 8, true, &mh))) return rc;
  if ((rc = get_map(gI_mid, B, H, W, DG_TH + 2, DG_TW + 8, true, &mm))) return rc;
  const DgLayout L = dg_layout(rec);
  const int grid = q.n_tiles < n_sms ? q.n_tiles : n_sms;
  static bool attr_set[2] = {false, false};
  if (!attr_set[rec ? 1 : 0]) {
    const cudaError_t e = rec ? cudaFuncSetAttribute(lif_dgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                              : cudaFuncSetAttribute(lif_dgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return check_launch("cudaFuncSetAttribute(lif_dgrad_tc_kernel)");
    attr_set[rec ? 1 : 0] = true;
  }
  if (rec) lif_dgrad_tc_kernel<true><<<grid, DG_THREADS, L.total, st>>>(q, mh, mm);
  else lif_dgrad_tc_kernel<false><<<grid, DG_THREADS, L.total, st>>>(q, mh, mm);
  return check_launch("lif_dgrad_tc_kernel");
}

// tensor-core weight gradient of B images into per-CTA partial sums (+ the fixed-order reduction with EF_WG_FINALIZE).  `rec`: the
// cell has a recurrent convolution (two slices per CTA in wg_partial); z_in_cl may still be NULL (no previous state at this step).
static int run_wgrad(const uint16_t* x_cl, const uint16_t* z_in_cl, const uint16_t* gI_hi, const uint16_t* gI_mid, bool rec, int B, int H, int W,
                     float* wg_partial, int flags, float* g_w_ff, float* g_w_rec, cudaStream_t st) {
  const bool wrec = rec && z_in_cl;
  const int wgrid_n = wg_grid(B, H, W);
  WgParams w;
  w.B = B, w.H = H, w.W = W, w.tiles_x = cdiv(W, WG_TW), w.tiles_y = cdiv(H, WG_TH), w.n_tiles = B * w.tiles_x * w.tiles_y;
  w.accumulate = (flags & EF_WG_ACCUMULATE) ? 1 : 0;
  w.partial = wg_partial;
  int rc;
  if (rec && !wrec && !w.accumulate) {  // recurrent cell without a previous state: its recurrent slices start at zero
    if (cudaMemsetAsync(wg_partial + (size_t)wgrid_n * WG_SLICE, 0, (size_t)wgrid_n * WG_SLICE * sizeof(float), st) != cudaSuccess)
      return check_launch("cudaMemsetAsync(wg_partial)");
  }
  CUtensorMap mx, mz, gh, gm;
  if ((rc = get_map(x_cl, B, H, W, WG_TH + 2, WG_TW + 8, true, &mx))) return rc;
  mz = mx;
  if (wrec && (rc = get_map(z_in_cl, B, H, W, WG_TH + 2, WG_TW + 8, true, &mz))) return rc;
  if ((rc = get_map(gI_hi, B, H, W, WG_TH, WG_TW, true, &gh))) return rc;
  if ((rc = get_map(gI_mid, B, H, W, WG_TH, WG_TW, true, &gm))) return rc;
  static bool wattr[2] = {false, false};
  if (!wattr[wrec ? 1 : 0]) {
    const cudaError_t e = wrec ? cudaFuncSetAttribute(lif_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                               : cudaFuncSetAttribute(lif_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return check_launch("cudaFuncSetAttribute(lif_wgrad_tc_kernel)");
    wattr[wrec ? 1 : 0] = true;
  }
  const int wsmem = WG_NST * ((wrec ? 2 : 1) * WG_XTILE + 2 * WG_GTILE) + 256 + 1024;
  if (wrec) lif_wgrad_tc_kernel<true><<<wgrid_n, WG_THREADS, wsmem, st>>>(w, mx, mz, gh, gm);
  else lif_wgrad_tc_kernel<false><<<wgrid_n, WG_THREADS, wsmem, st>>>(w, mx, mz, gh, gm);
  if ((rc = check_launch("lif_wgrad_tc_kernel"))) return rc;
  if (flags & EF_WG_FINALIZE) {
    wgrad_reduce_kernel<<<(rec ? 2 : 1) * 288, 256, 0, st>>>(wg_partial, wgrid_n, g_w_ff, rec ? g_w_rec : nullptr);
    if ((rc = check_launch("wgrad_reduce_kernel"))) return rc;
  }
  return EF_OK;
}

}  // namespace ef

extern "C" int ef_lif_bwd_tc(const ef_lif_bwd_tc_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_lif_bwd_tc: params is NULL");
  const ef_lif_bwd_tc_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_lif_bwd_tc: bad dimensions");
  const bool head = p.x_f32 != nullptr;
  if (head) {
    EF_REQUIRE(p.v_out && p.leak && p.thresh && p.gI_f32, EF_ENULL, "ef_lif_bwd_tc (head mode): NULL tensor");
    EF_REQUIRE(p.Cin > 0 && p.Cin <= HW_MAXC, EF_EUNSUPPORTED, "ef_lif_bwd_tc (head mode): Cin <= %d (got %d)", HW_MAXC, p.Cin);
  } else {
    EF_REQUIRE(p.x_cl && p.v_out && p.leak && p.thresh && p.w_bwd && p.gI_hi && p.gI_mid && p.g_x, EF_ENULL, "ef_lif_bwd_tc: NULL tensor");
  }
  cudaStream_t st = as_stream(stream);
  int rc;
  if ((rc = run_pointwise_step(p, st))) return rc;
  if (head) return p.g_w_ff ? run_head_wgrad(p.x_f32, p.gI_f32, p.g_w_ff, p.B, p.Cin, p.H, p.W, st) : EF_OK;
  const bool rec = p.has_rec != 0;
  if ((rc = run_dgrad(p.gI_hi, p.gI_mid, p.w_bwd, rec, p.g_x, (rec && p.z_in_cl) ? p.g_z_in : nullptr, p.B, p.H, p.W, st))) return rc;
  if (p.wg_partial) return run_wgrad(p.x_cl, p.z_in_cl, p.gI_hi, p.gI_mid, rec, p.B, p.H, p.W, p.wg_partial, p.wg_flags, p.g_w_ff, p.g_w_rec, st);
  const dim3 wgrid(cdiv(p.W, 16), cdiv(p.H, 16), p.B);
  if (p.g_w_ff) {
    conv_wgrad_cl_kernel<<<wgrid, WGC_THREADS, 0, st>>>(p.x_cl, p.gI_hi, p.gI_mid, p.g_w_ff, p.B, p.H, p.W);
    if ((rc = check_launch("conv_wgrad_cl_kernel(ff)"))) return rc;
  }
  if (rec && p.g_w_rec && p.z_in_cl) {
    conv_wgrad_cl_kernel<<<wgrid, WGC_THREADS, 0, st>>>(p.z_in_cl, p.gI_hi, p.gI_mid, p.g_w_rec, p.B, p.H, p.W);
    if ((rc = check_launch("conv_wgrad_cl_kernel(rec)"))) return rc;
  }
  return EF_OK;
}

// Backward of a feed-forward 32 -> 32 LIF cell (or, head mode, of the Cin <= 8 input cell) over a whole BPTT window: ONE time-fused
// pointwise launch, ONE data-gradient launch and ONE weight-gradient launch (+ its reduction) for all T steps -- the step index is
// folded into the batch dimension of the tensor-core kernels.
extern "C" int ef_lif_bwd_window(const ef_lif_bwd_window_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_lif_bwd_window: params is NULL");
  const ef_lif_bwd_window_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.T > 0 && p.H > 0 && p.W > 0 && (int64_t)p.B * p.T < (1 << 20), EF_EINVAL, "ef_lif_bwd_window: bad dimensions");
  const bool head = p.x_f32 != nullptr;
  const bool head_tc = !head && p.Cin > 0 && p.Cin < 32;  // head layer on split inputs: x_cl from ef_pack_split_cl, no data gradient
  EF_REQUIRE(p.v && p.g_out && p.leak && p.thresh && (p.T == 1 || p.z_cl), EF_ENULL, "ef_lif_bwd_window: NULL tensor");
  if (head_tc) {
    EF_REQUIRE(p.Cin <= EF_HEAD_MAX_CIN && p.x_cl && p.gI_hi && p.gI_mid && !p.g_x, EF_EINVAL, "ef_lif_bwd_window (split-input head): bad arguments");
    EF_REQUIRE(!p.g_w_ff || p.wg_partial, EF_ENULL, "ef_lif_bwd_window: g_w_ff wanted but wg_partial is NULL");
  } else if (head) {
    EF_REQUIRE(p.gI_f32, EF_ENULL, "ef_lif_bwd_window (head mode): gI_f32 is NULL");
    EF_REQUIRE(p.Cin > 0 && p.Cin <= HW_MAXC, EF_EUNSUPPORTED, "ef_lif_bwd_window (head mode): Cin <= %d (got %d)", HW_MAXC, p.Cin);
  } else if (!head_tc) {
    EF_REQUIRE(p.x_cl && p.w_bwd && p.gI_hi && p.gI_mid, EF_ENULL, "ef_lif_bwd_window: NULL tensor");
    EF_REQUIRE(!p.g_w_ff || p.wg_partial, EF_ENULL, "ef_lif_bwd_window: g_w_ff wanted but wg_partial is NULL");
  }
  cudaStream_t st = as_stream(stream);
  int rc;
  if ((rc = run_pointwise_window(p, st))) return rc;
  const int BT = p.B * p.T;
  if (head) return p.g_w_ff ? run_head_wgrad(p.x_f32, p.gI_f32, p.g_w_ff, BT, p.Cin, p.H, p.W, st) : EF_OK;
  if (head_tc) {
    if (!p.g_w_ff) return EF_OK;
    if ((rc = run_wgrad(p.x_cl, nullptr, p.gI_hi, p.gI_mid, false, BT, p.H, p.W, p.wg_partial, 0, nullptr, nullptr, st))) return rc;
    wgrad_reduce_head_kernel<<<288, 256, 0, st>>>(p.wg_partial, wg_grid(BT, p.H, p.W), p.g_w_ff, p.Cin, EF_HEAD_SLOT(p.Cin));
    return check_launch("wgrad_reduce_head_kernel");
  }
  if (p.g_x && (rc = run_dgrad(p.gI_hi, p.gI_mid, p.w_bwd, false, p.g_x, nullptr, BT, p.H, p.W, st))) return rc;
  if (p.g_w_ff) return run_wgrad(p.x_cl, nullptr, p.gI_hi, p.gI_mid, false, BT, p.H, p.W, p.wg_partial, EF_WG_FINALIZE, p.g_w_ff, nullptr, st);
  return EF_OK;
}

// Tensor-core weight gradient alone (the last stage of ef_lif_bwd_tc) for B images whose g_I = gI_hi + gI_mid already exists: lets a
// recurrent cell run pointwise + data gradient step by step (the recurrent path needs them in order) and the weight gradient of the
// whole window in one or two batched calls.  Protocol of wg_partial / flags as for ef_lif_bwd_tc.
extern "C" int ef_lif_wgrad_tc(const uint16_t* x_cl, const uint16_t* z_in_cl, const uint16_t* gI_hi, const uint16_t* gI_mid, int32_t has_rec, int32_t B,
                               int32_t H, int32_t W, float* wg_partial, int32_t wg_flags, float* g_w_ff, float* g_w_rec, void* stream) {
  using namespace ef;
  EF_REQUIRE(B > 0 && H > 0 && W > 0, EF_EINVAL, "ef_lif_wgrad_tc: bad dimensions");
  EF_REQUIRE(x_cl && gI_hi && gI_mid && wg_partial, EF_ENULL, "ef_lif_wgrad_tc: NULL tensor");
  return run_wgrad(x_cl, z_in_cl, gI_hi, gI_mid, has_rec != 0, B, H, W, wg_partial, wg_flags, g_w_ff, g_w_rec, as_stream(stream));
}

namespace ef {

// g_I fp32 NCHW [B,C,Hs,Ws] -> two bf16 channels-last terms [B,H,W,C] (hi + mid = g_I to 16 significant bits): the operand format of
// the tensor-core gradient kernels, for a g_I that some other neuron backward (lif_bwd_pointwise_kernel) produced.  up = 2: the output is
// the zero-inserted form at twice the resolution (H = 2 Hs or 2 Hs - 1), what the data gradient of a stride-2 convolution convolves.
__global__ void __launch_bounds__(256) split2_pack_cl_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ mid, int B,
                                                            int C, int H, int W, int Hs, int Ws, int up) {
  const size_t hw = (size_t)H * W, hws = (size_t)Hs * Ws;
  const int G = C >> 3;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // over B * C/8 channel groups * hw, pixel fastest
  if (i >= (size_t)B * G * hw) return;
  const size_t pix = i % hw, bg = i / hw;
  const int g = (int)(bg % G);
  const size_t b = bg / G;
  uint32_t h[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0};
  const int y = (int)(pix / W), x = (int)(pix % W);
  if (up == 1 || ((x | y) & 1) == 0) {
    const float* s = src + (b * C + g * 8) * hws + (up == 1 ? pix : (size_t)(y >> 1) * Ws + (x >> 1));
#pragma unroll
    for (int k = 0; k < 4; ++k) split2_bf16(s[(2 * k) * hws], s[(2 * k + 1) * hws], h[k], m[k]);
  }
  const size_t o = (b * hw + pix) * C + g * 8;
  *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(mid + o) = make_uint4(m[0], m[1], m[2], m[3]);
}

// after the tensor-core data gradient: g_z_in += conv^T part (kept apart because g_z_in already holds the neuron's direct terms) and the
// PLIF / XLIF trace term of g_x, sign(x) / 32 * adjoint of the 3x3 average pool of gP_sum (as conv_dgrad_kernel adds it)
__global__ void __launch_bounds__(256) conv32_bwd_fixup_kernel(float* __restrict__ g_x, const float* __restrict__ x, const float* __restrict__ gP_sum,
                                                              float* __restrict__ g_z_in, const float* __restrict__ g_z_tmp, int B, int H, int W) {
  const size_t hw = (size_t)H * W, n = (size_t)B * hw;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const size_t b = i / hw, pix = i % hw;
  const int y = (int)(pix / W), xq = (int)(pix % W);
  float tr = 0.f;
  if (gP_sum) {
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = xq + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) tr += gP_sum[(b * H + yy) * W + xx];
      }
    tr = tr / 9.0f / 32.0f;
  }
#pragma unroll 4
  for (int c = 0; c < 32; ++c) {
    const size_t o = (b * 32 + c) * hw + pix;
    if (gP_sum) {
      const float xv = x[o];
      g_x[o] += (xv > 0.f ? tr : (xv < 0.f ? -tr : 0.f));
    }
    if (g_z_in) g_z_in[o] += g_z_tmp[o];
  }
}

}  // namespace ef

extern "C" int ef_conv32_bwd_tc(const ef_conv32_bwd_tc_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_conv32_bwd_tc: params is NULL");
  const ef_conv32_bwd_tc_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0 && p.W % 4 == 0, EF_EINVAL, "ef_conv32_bwd_tc: bad dimensions (W must be a multiple of 4)");
  EF_REQUIRE(p.gI && p.x_cl && p.w_bwd && p.gI_hi && p.gI_mid && p.g_x, EF_ENULL, "ef_conv32_bwd_tc: NULL tensor");
  EF_REQUIRE(!p.g_z_in || p.g_z_tmp, EF_ENULL, "ef_conv32_bwd_tc: g_z_in needs g_z_tmp");
  EF_REQUIRE(!p.gP_sum || p.x_f32, EF_ENULL, "ef_conv32_bwd_tc: the trace term needs the fp32 input");
  EF_REQUIRE(!(p.g_w_ff || p.g_w_rec) || p.wg_partial, EF_ENULL, "ef_conv32_bwd_tc: weight gradients need wg_partial");
  cudaStream_t st = as_stream(stream);
  const bool rec = p.has_rec != 0;
  const size_t hw = (size_t)p.H * p.W, n = (size_t)p.B * 4 * hw;
  split2_pack_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.gI, p.gI_hi, p.gI_mid, p.B, 32, p.H, p.W, p.H, p.W, 1);
  int rc;
  if ((rc = check_launch("split2_pack_cl_kernel"))) return rc;
  float* gz = (rec && p.z_in_cl && p.g_z_in) ? p.g_z_tmp : nullptr;
  if ((rc = run_dgrad(p.gI_hi, p.gI_mid, p.w_bwd, rec, p.g_x, gz, p.B, p.H, p.W, st))) return rc;
  if (gz || p.gP_sum) {
    const size_t m = (size_t)p.B * hw;
    conv32_bwd_fixup_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(p.g_x, p.x_f32, p.gP_sum, gz ? p.g_z_in : nullptr, gz, p.B, p.H, p.W);
    if ((rc = check_launch("conv32_bwd_fixup_kernel"))) return rc;
  }
  if (p.g_w_ff || p.g_w_rec) return run_wgrad(p.x_cl, p.z_in_cl, p.gI_hi, p.gI_mid, rec, p.B, p.H, p.W, p.wg_partial, p.wg_flags, p.g_w_ff, p.g_w_rec, st);
  return EF_OK;
}

// g fp32 NCHW [B,C,Hs,Ws] (C a multiple of 8) -> hi / mid bf16 channels-last [B,H,W,C]; (H, W) = (Hs, Ws), or the zero-inserted form at
// the resolution (H, W) of a stride-2 convolution's input (Hs = (H-1)/2 + 1): the sources of a data gradient that runs as a plain
// convolution on the general tensor-core kernel (ef_lif_conv_fwd_g with flipped / transposed weights).
extern "C" int ef_split2_pack_cl(const float* src, uint16_t* hi, uint16_t* mid, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Hs, int32_t Ws,
                                 void* stream) {
  using namespace ef;
  EF_REQUIRE(src && hi && mid, EF_ENULL, "ef_split2_pack_cl: NULL tensor");
  EF_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, EF_EINVAL, "ef_split2_pack_cl: C must be a positive multiple of 8");
  const bool same = Hs == H && Ws == W, up2 = Hs == (H - 1) / 2 + 1 && Ws == (W - 1) / 2 + 1;
  EF_REQUIRE(same || up2, EF_EINVAL, "ef_split2_pack_cl: source %dx%d is neither the output size %dx%d nor its stride-2 size", Hs, Ws, H, W);
  const size_t n = (size_t)B * (C / 8) * H * W;
  split2_pack_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(src, hi, mid, B, C, H, W, Hs, Ws, same ? 1 : 2);
  return check_launch("split2_pack_cl_kernel");
}

// Tensor-core weight gradient for general channel counts: g_w[co][ci_off + ci][dy][dx] += sum_{b,y,x} x_cl[b,y+dy-1,x+dx-1,ci] g[b,y,x,co] with
// x_cl [B,H,W,cin] (values exact in bf16) and g = g_hi + g_mid [B,H,W,cout] (ef_split2_pack_cl; for a stride-2 convolution its zero-inserted
// form at the input resolution); cin, cout multiples of 32.  g_w is [cout][cin_total][3][3]; ci_off places the cin channels inside it.
extern "C" int64_t ef_wgrad_tcg_partial_elems(int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout) {
  using namespace ef;
  if (B <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0 || cin % 32 || cout % 32) return 0;
  const int pairs = (cin / 32) * (cout / 32);
  return (int64_t)pairs * wgg_chunks(B, H, W, pairs) * WG_SLICE;
}
extern "C" int ef_wgrad_tcg(const uint16_t* x_cl, const uint16_t* g_hi, const uint16_t* g_mid, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout,
                            float* partial, float* g_w, int32_t cin_total, int32_t ci_off, void* stream) {
  using namespace ef;
  EF_REQUIRE(x_cl && g_hi && g_mid && partial && g_w, EF_ENULL, "ef_wgrad_tcg: NULL tensor");
  EF_REQUIRE(B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0 && cin % 32 == 0 && cout % 32 == 0, EF_EINVAL, "ef_wgrad_tcg: channel counts must be multiples of 32");
  EF_REQUIRE(ci_off >= 0 && ci_off + cin <= cin_total, EF_EINVAL, "ef_wgrad_tcg: channel slice outside the weight tensor");
  cudaStream_t st = as_stream(stream);
  WgGParams w;
  w.B = B, w.H = H, w.W = W, w.tiles_x = cdiv(W, WG_TW), w.tiles_y = cdiv(H, WG_TH), w.n_tiles = B * w.tiles_x * w.tiles_y;
  w.n_co_blk = cout / 32;
  const int pairs = (cin / 32) * w.n_co_blk;
  w.chunks = wgg_chunks(B, H, W, pairs);
  w.partial = partial;
  CUtensorMap mx, gh, gm;
  int rc;
  if ((rc = get_map_c(x_cl, B, H, W, cin, WG_TH + 2, WG_TW + 8, true, &mx))) return rc;
  if ((rc = get_map_c(g_hi, B, H, W, cout, WG_TH, WG_TW, true, &gh))) return rc;
  if ((rc = get_map_c(g_mid, B, H, W, cout, WG_TH, WG_TW, true, &gm))) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(lif_wgrad_tcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(lif_wgrad_tcg_kernel)");
    attr = true;
  }
  const int wsmem = WG_NST * (WG_XTILE + 2 * WG_GTILE) + 256 + 1024;
  lif_wgrad_tcg_kernel<<<pairs * w.chunks, WG_THREADS, wsmem, st>>>(w, mx, gh, gm);
  if ((rc = check_launch("lif_wgrad_tcg_kernel"))) return rc;
  wgrad_reduce_g_kernel<<<pairs * 4, 256, 0, st>>>(partial, w.chunks, w.n_co_blk, g_w, cin_total, ci_off);
  return check_launch("wgrad_reduce_g_kernel");
}
