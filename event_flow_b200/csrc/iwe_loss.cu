// Contrast-maximisation event-warping loss (hot path B): forward, analytic backward, and the per-polarity IWE image.
// Reference: loss/flow.py:176-301 (EventWarping.forward), utils/iwe.py:4-92 (purge_unfeasible, get_interpolation,
// interpolate), loss/flow.py:65-79 (per-event flow gather), utils/iwe.py:95-153 (deblur_events, compute_pol_iwe).
// Sub-gradient conventions at ties (integer warped coordinates, empty pixels) follow SURVEY.md 7.4, i.e. what
// torch.autograd produces for the reference expressions.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ef {

// ---- window descriptor (device side) ------------------------------------------------------------------------------
// Both public forms of a window -- concatenated ("map form", ef_iwe_loss_params) and per-pass pointer tables
// (ef_iwe_loss_pass_params: nothing is concatenated or copied) -- are lowered to this by-value table: pass t has its own
// base pointers and its own sample stride.
constexpr int MAXP = EF_IWE_MAX_PASSES, MAXS = EF_IWE_MAX_SCALES;
struct IweWin {
  int S, B, T, Tm, H, W;
  int loss_scaling, use_mask, use_dt;
  float flow_scaling, smooth_coef;  // smooth_coef = weight / components / T_maps
  int vec;                          // W % 4 == 0 and every flow / mask / gradient plane 16-byte aligned: the pixel passes use 16-byte accesses
  int debug_skip;                   // timing experiments only (EF_IWE_SKIP): 1 = no smoothness, 2 = no event pass, 4 = no reduction / adjoint pass
  int n_items;                      // chunks of IWE_CHUNK events of one (scale, sample): sum_t ceil(n_t / IWE_CHUNK)
  int chunk_off[MAXP + 1];          // first chunk of pass t
  int n_pass[MAXP];                 // events of pass t (per sample)
  const float* ev[MAXP];            // pass t: [B][n_t][4], sample stride ev_bs[t] floats
  const float* pm[MAXP];            // pass t: [B][n_t][2], sample stride pm_bs[t] floats
  long long ev_bs[MAXP], pm_bs[MAXP];
  const float* flow[MAXS * MAXP];   // (s, map m): [B][2][H][W], sample stride flow_bs floats
  long long flow_bs;
  const float* mask[MAXP];          // map m: [B][H][W], sample stride mask_bs floats (NULL without smoothing mask)
  long long mask_bs;
  float* g_flow[MAXS * MAXP];       // backward: (s, map m): [B][2][H][W], sample stride g_bs
  long long g_bs;
};
#ifndef EF_IWE_EPT
#define EF_IWE_EPT 1
#endif
#ifndef EF_IWE_OCC_F
#define EF_IWE_OCC_F 4
#endif
#ifndef EF_IWE_OCC_B
#define EF_IWE_OCC_B 4
#endif
#ifndef EF_IWE_NX
#define EF_IWE_NX 4
#endif
constexpr int IWE_THREADS = 256, IWE_EPT = EF_IWE_EPT, IWE_CHUNK = IWE_THREADS * IWE_EPT;  // events per thread and per work item of the event passes
constexpr int RED_PIX = 1024;  // pixels per work item of the per-pixel passes
constexpr int SM_ROWS = 8;      // image rows per work item of the smoothness passes
constexpr int SM_NX = EF_IWE_NX;        // pixels per thread of the smoothness passes (vector path)

// workspace layout (floats):
//   ctr  [16]  (as uint32) grid-barrier counters.  Must be ZERO when a buffer is first used; every call leaves them zero.
//   img  [S][B][2 dir][2 pol][plane]   forward accumulators.  Polarity-planar, (I, Th) interleaved per pixel.  An event adds
//        (w, w*tau) of the two corners of an image row -- neighbouring pixels l, l+1 -- with ONE 16-byte vector reduction
//        (red.global.add.v4.f32), whatever the parity of l: a plane holds the image TWICE, copy A for pairs that start at an even
//        pixel (entry = pixel) and copy B, shifted by one pixel (entry = pixel + 1), for pairs that start at an odd pixel; the
//        per-pixel passes read A[i] + B[i + 1].  4 vector reductions per event instead of 16 scalar atomics (6 with one copy).
//        plane = [HW (A) + HW + 2 (B)][2] floats
//   sums [S][B][2 dir][2 = sum A^2, n]
//   smooth_part [S][MAX_GRID]                  per-CTA partial sums of the smoothness term (fixed-order final reduction)
//   adj  [S][B][2 dir][2 pol][HW][2 = dL/dI, dL/dTh]   adjoint images, written and read by the backward only
struct WsLayout {
  size_t ctr, img, sums, smooth, adj, total;
};
constexpr int IWE_MAX_GRID = 148 * 8;
__host__ __device__ inline size_t plane_px(size_t hw) { return 2 * hw + 2; }  // entries (pixels) of one accumulator plane: A then B
__host__ __device__ inline WsLayout ws_layout(int S, int B, int H, int W) {
  WsLayout l;
  const size_t hw = (size_t)H * W;
  l.ctr = 0;
  l.img = 16;
  l.sums = l.img + (size_t)S * B * 4 * plane_px(hw) * 2;
  l.smooth = l.sums + (size_t)S * B * 4;
  l.adj = l.smooth + (size_t)S * IWE_MAX_GRID;  // backward only: adjoint images [sbd][pol][HW][2]
  l.adj = (l.adj + 3) & ~(size_t)3;
  l.total = l.adj + (size_t)S * B * 8 * hw;
  return l;
}

struct Corner {
  int idx;      // flat pixel, -1 if out of bounds
  float wy, wx; // bilinear factors
  float dy, dx; // y' - iy, x' - ix
};

// utils/iwe.py:37-72 for one event and one reference time.  All arithmetic in the reference's op order, no contraction.
__device__ __forceinline__ void warp_event(float ts, float y, float x, float fy, float fx, float tref, float scale, int H, int W,
                                           float& yw, float& xw, Corner (&c)[4]) {
  const float dt = __fsub_rn(tref, ts);
  yw = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, fy), scale));
  xw = __fadd_rn(x, __fmul_rn(__fmul_rn(dt, fx), scale));
  const float top = floorf(yw), bot = floorf(__fadd_rn(yw, 1.0f));
  const float left = floorf(xw), right = floorf(__fadd_rn(xw, 1.0f));
  const float iy[4] = {top, top, bot, bot};
  const float ix[4] = {left, right, left, right};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c[k].dy = __fsub_rn(yw, iy[k]);
    c[k].dx = __fsub_rn(xw, ix[k]);
    c[k].wy = fmaxf(0.f, __fsub_rn(1.0f, fabsf(c[k].dy)));
    c[k].wx = fmaxf(0.f, __fsub_rn(1.0f, fabsf(c[k].dx)));
    const bool oob = (iy[k] < 0.f) || (iy[k] >= (float)H) || (ix[k] < 0.f) || (ix[k] >= (float)W);
    c[k].idx = oob ? -1 : (int)(iy[k] * (float)W + ix[k]);
  }
}

__device__ __forceinline__ float block_sum256(float v, float* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    r = lane < 8 ? s_red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// ---- grid-wide barrier of a co-resident (cooperative) grid ------------------------------------------------------------
// ctr is zero on entry; the caller resets it once every CTA is known to have left the spin (i.e. after a LATER barrier).
// Bounded spin: a protocol bug traps (an error the host sees) instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
      if (seen < n && clock64() - t0 > 4000000000ll) __trap();
    } while (seen < n);
    __threadfence();
  }
  __syncthreads();
}


// ---- smoothness: Charbonnier terms of the flow maps (loss/flow.py:262-294) ---------------------------------------------
// A thread owns NX consecutive pixels of one image row and loads their 3 x (NX + 2) neighbourhood of (mask, fx, fy) -- 16-byte loads
// for the NX = 4 centre columns -- plus the same pixels of the next / previous map, ALL in one round of independent loads (no
// mask-then-flow dependency: the pass is latency-bound at training sizes, bandwidth-bound on large windows).  Out-of-image
// neighbours carry mask 0, which also encodes that the pair does not exist.
__device__ __forceinline__ float sqrt_fast(float x) {  // sqrt.approx (1 ulp class): far inside the 1e-5 tolerance of the loss value
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
struct SmoothMaps {  // one (scale, sample, map): this map and its temporal neighbours
  const float *fx, *mk;          // flow x plane (y plane = fx + hw), event mask (NULL: no smoothing mask)
  const float *fx_n, *mk_n;      // next map (NULL: none)
  const float *fx_p, *mk_p;      // previous map (gradient only)
};
template <int NX>
__device__ __forceinline__ void load_row(float (&dst)[NX + 2], const float* __restrict__ src, int y, int x, int H, int W, float oob, bool have) {
  // columns x-1 .. x+NX of row y; `have` false (no such array): the constant `oob`... of an absent mask is 1 inside the image
  const bool row_ok = y >= 0 && y < H;
  if (!row_ok || src == nullptr) {
#pragma unroll
    for (int k = 0; k < NX + 2; ++k) dst[k] = (row_ok && !have && x - 1 + k >= 0 && x - 1 + k < W) ? 1.f : oob;
    return;
  }
  const float* r = src + (size_t)y * W;
  dst[0] = x >= 1 ? __ldg(r + x - 1) : oob;
  if (NX == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(r + x));
    dst[1] = v.x, dst[2] = v.y, dst[3] = v.z, dst[4] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < NX; ++k) dst[1 + k] = __ldg(r + x + k);
  }
  dst[NX + 1] = x + NX < W ? __ldg(r + x + NX) : oob;
}
template <int NX>
__device__ __forceinline__ void load_ctr(float (&dst)[NX], const float* __restrict__ src, int y, int x, int W, bool have) {
  if (src == nullptr) {
#pragma unroll
    for (int k = 0; k < NX; ++k) dst[k] = have ? 0.f : 1.f;
    return;
  }
  const float* r = src + (size_t)y * W + x;
  if (NX == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(r));
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < NX; ++k) dst[k] = __ldg(r + k);
  }
}
// one pair (a first, b second): value m sqrt(d^2 + eps), derivative wrt f[a] (both channels) m d / sqrt(d^2 + eps); minus for f[b]
__device__ __forceinline__ void pair_term(float ma, float fxa, float fya, float mb, float fxb, float fyb, float& val, float& der) {
  const float m = ma * mb;
  val = der = 0.f;
  if (m == 0.f) return;  // event masks are sparse: most pairs vanish
  const float d = (fxa - fxb) + (fya - fyb);
  const float q = d * d + 1e-6f;
  const float r = rsqrtf(q);
  val = m * q * r;
  der = m * d * r;
}
// NX pixels (y, x .. x+NX-1) of one map.  Returns the sum of the pairs these pixels are the FIRST element of (right, down, down-right,
// the up-right pair ((y+1,x),(y,x+1)), temporal next); with GRAD also g[k] = d(sum of ALL pairs) / d f(y, x+k) (gather form, unscaled).
template <bool GRAD, int NX>
__device__ __forceinline__ float smooth_strip(const SmoothMaps& sm, int y, int x, int H, int W, size_t hw, float (&g)[NX]) {
  constexpr int R0 = GRAD ? 0 : 1;  // rows y-1 .. y+1 (gradient) or y .. y+1 (value)
  float m[3][NX + 2], fx[3][NX + 2], fy[3][NX + 2];
  float mn[NX], fxn[NX], fyn[NX], mp[NX], fxp[NX], fyp[NX];
#pragma unroll
  for (int r = R0; r < 3; ++r) {
    load_row<NX>(m[r], sm.mk, y - 1 + r, x, H, W, 0.f, sm.mk != nullptr);
    load_row<NX>(fx[r], sm.fx, y - 1 + r, x, H, W, 0.f, true);
    load_row<NX>(fy[r], sm.fx + hw, y - 1 + r, x, H, W, 0.f, true);
  }
  const bool nxt = sm.fx_n != nullptr, prv = GRAD && sm.fx_p != nullptr;
  if (nxt) {
    load_ctr<NX>(mn, sm.mk_n, y, x, W, false);
    load_ctr<NX>(fxn, sm.fx_n, y, x, W, true);
    load_ctr<NX>(fyn, sm.fx_n + hw, y, x, W, true);
  }
  if (prv) {
    load_ctr<NX>(mp, sm.mk_p, y, x, W, false);
    load_ctr<NX>(fxp, sm.fx_p, y, x, W, true);
    load_ctr<NX>(fyp, sm.fx_p + hw, y, x, W, true);
  }
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < NX; ++k) {
    const int c = k + 1;
    float v, d, gk = 0.f;
    const float mo = m[1][c], fo = fx[1][c], go = fy[1][c];
    pair_term(mo, fo, go, m[1][c + 1], fx[1][c + 1], fy[1][c + 1], v, d), acc += v, gk += d;            // right
    pair_term(mo, fo, go, m[2][c], fx[2][c], fy[2][c], v, d), acc += v, gk += d;                        // down
    pair_term(mo, fo, go, m[2][c + 1], fx[2][c + 1], fy[2][c + 1], v, d), acc += v, gk += d;            // down-right
    pair_term(m[2][c], fx[2][c], fy[2][c], m[1][c + 1], fx[1][c + 1], fy[1][c + 1], v, d), acc += v;    // ((y+1,x),(y,x+1)): value only
    if (nxt) pair_term(mo, fo, go, mn[k], fxn[k], fyn[k], v, d), acc += v, gk += d;                     // temporal
    if (GRAD) {
      pair_term(m[1][c - 1], fx[1][c - 1], fy[1][c - 1], mo, fo, go, v, d), gk -= d;                    // left neighbour's right pair
      pair_term(m[0][c], fx[0][c], fy[0][c], mo, fo, go, v, d), gk -= d;                                // upper neighbour's down pair
      pair_term(m[0][c - 1], fx[0][c - 1], fy[0][c - 1], mo, fo, go, v, d), gk -= d;                    // upper-left neighbour's down-right pair
      pair_term(mo, fo, go, m[0][c + 1], fx[0][c + 1], fy[0][c + 1], v, d), gk += d;                    // ((y,x),(y-1,x+1)): first element
      pair_term(m[2][c - 1], fx[2][c - 1], fy[2][c - 1], mo, fo, go, v, d), gk -= d;                    // ((y+1,x-1),(y,x)): second element
      if (prv) pair_term(mp[k], fxp[k], fyp[k], mo, fo, go, v, d), gk -= d;                             // previous map's temporal pair
      g[k] = gk;
    }
  }
  return acc;
}
__device__ __forceinline__ SmoothMaps smooth_maps(const IweWin& w, int s, int b, int t, size_t hw, bool grad) {
  SmoothMaps sm;
  sm.fx = w.flow[s * w.Tm + t] + (size_t)b * w.flow_bs;
  sm.mk = w.use_mask ? w.mask[t] + (size_t)b * w.mask_bs : nullptr;
  const bool dtn = w.use_dt && t + 1 < w.Tm, dtp = grad && w.use_dt && t >= 1;
  sm.fx_n = dtn ? w.flow[s * w.Tm + t + 1] + (size_t)b * w.flow_bs : nullptr;
  sm.mk_n = (dtn && w.use_mask) ? w.mask[t + 1] + (size_t)b * w.mask_bs : nullptr;
  sm.fx_p = dtp ? w.flow[s * w.Tm + t - 1] + (size_t)b * w.flow_bs : nullptr;
  sm.mk_p = (dtp && w.use_mask) ? w.mask[t - 1] + (size_t)b * w.mask_bs : nullptr;
  return sm;
}

// item -> (scale*B + sample, pass, first event of the chunk)
__device__ __forceinline__ void decode_item(const IweWin& w, int item, int& sb, int& t, int& i0) {
  sb = item / w.n_items;
  const int c = item - sb * w.n_items;
  t = 0;
  while (t + 1 < w.T && c >= w.chunk_off[t + 1]) ++t;
  i0 = (c - w.chunk_off[t]) * IWE_CHUNK;
}

// vector reductions without a return value (RED, not ATOM: nothing waits for the round trip)
#ifndef EF_IWE_HINT
#define EF_IWE_HINT 0  // measured on B200 (tools/iwe_sweep.sh): the policies change nothing (689 vs 681 us on the 16 M-event window)
#endif
// L2 policies: the accumulator images are hit again and again by later events (keep: evict_last), the flow values gathered per event
// and the events themselves are used once (evict_first)
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void red2(float* q, float a, float b, uint64_t pol) {
#if EF_IWE_HINT
  asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(q), "f"(a), "f"(b), "l"(pol) : "memory");
#else
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(q), "f"(a), "f"(b) : "memory");
#endif
}
__device__ __forceinline__ void red4(float* q, float a, float b, float c, float d, uint64_t pol) {
#if EF_IWE_HINT
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(q), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
#else
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
#endif
}
__device__ __forceinline__ float ldg_stream(const float* q, uint64_t pol) {
#if EF_IWE_HINT
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(q), "l"(pol));
  return v;
#else
  return __ldg(q);
#endif
}

// events of one work item: thread tid owns events i0 + tid + k * IWE_THREADS, k < IWE_EPT (all loads of a thread are issued before any is used)
struct Ev {
  float4 e;   // ts, y, x, p
  float2 pm;
  float fx, fy;
  int pix;
  bool on;
};
__device__ __forceinline__ void load_events(const IweWin& w, int s, int b, int t, int i0, size_t hw, Ev (&ev)[IWE_EPT], uint64_t pol_stream) {
  const float4* ep = reinterpret_cast<const float4*>(w.ev[t] + (size_t)b * w.ev_bs[t]);
  const float2* pp = reinterpret_cast<const float2*>(w.pm[t] + (size_t)b * w.pm_bs[t]);
  const float* fm = w.flow[s * w.Tm + (w.Tm > 1 ? t : 0)] + (size_t)b * w.flow_bs;
#pragma unroll
  for (int k = 0; k < IWE_EPT; ++k) {
    const int i = i0 + threadIdx.x + k * IWE_THREADS;
    ev[k].on = i < w.n_pass[t];
    if (ev[k].on) {
      ev[k].e = __ldcs(ep + i);  // streamed once: evict-first, the accumulator images stay in L2
      ev[k].pm = __ldcs(pp + i);
    }
  }
#pragma unroll
  for (int k = 0; k < IWE_EPT; ++k) {
    ev[k].on = ev[k].on && !(ev[k].pm.x == 0.f && ev[k].pm.y == 0.f);
    if (ev[k].on) {
      ev[k].pix = (int)(ev[k].e.y * (float)w.W + ev[k].e.z);
      ev[k].fx = ldg_stream(fm + ev[k].pix, pol_stream), ev[k].fy = ldg_stream(fm + hw + ev[k].pix, pol_stream);
    }
  }
}

// accumulated (I, Th) of pixel i of one plane: copy A + copy B (shifted by one pixel)
__device__ __forceinline__ float2 acc_px(const float2* pl, int i, size_t hw) {
  const float2 a = __ldcg(pl + i), b = __ldcg(pl + hw + i + 1);
  return make_float2(a.x + b.x, a.y + b.y);
}

// ---- forward: ONE launch ----------------------------------------------------------------------------------------------
//   phase 0  zero the accumulators; Charbonnier smoothness of the flow maps (loss/flow.py:262-294) into per-CTA partials
//   phase 1  per event: gather flow, warp forward (tref = T) and backward (tref = 0), bilinear scatter (loss/flow.py:193-259,
//            utils/iwe.py:20-92)
//   phase 2  per (scale, sample, direction): sum of squared average timestamps and number of pixels with events (:212-226)
//   phase 3  last CTA: the scalar
__global__ void __launch_bounds__(IWE_THREADS, EF_IWE_OCC_F) iwe_loss_fwd_kernel(const __grid_constant__ IweWin w, float* __restrict__ ws, float* __restrict__ loss) {
  __shared__ float s_red[8];
  __shared__ bool s_last;
  const WsLayout l = ws_layout(w.S, w.B, w.H, w.W);
  unsigned int* ctr = reinterpret_cast<unsigned int*>(ws + l.ctr);
  float* img = ws + l.img;
  float* sums = ws + l.sums;
  const size_t hw = (size_t)w.H * w.W;
  const size_t plane = plane_px(hw) * 2;  // floats per (scale, sample, direction, polarity)
  const int tid = threadIdx.x;
  const size_t gtid = (size_t)blockIdx.x * IWE_THREADS + tid, gsz = (size_t)gridDim.x * IWE_THREADS;

  // ---- phase 0
  {
    float4* z = reinterpret_cast<float4*>(img);
    const size_t n4 = (l.smooth - l.img) / 4;  // images + sums (contiguous, multiple of 4 floats)
    for (size_t i = gtid; i < n4; i += gsz) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // smoothness: work item = SM_ROWS rows of one (scale, sample, map) plane
    const int rb = (w.H + SM_ROWS - 1) / SM_ROWS, planes = w.B * w.Tm;
    const int nx = w.vec ? SM_NX : 1, groups = w.W / nx;
    for (int s = 0; s < w.S; ++s) {
      float acc = 0.f;
      if (w.smooth_coef != 0.f && !(w.debug_skip & 1)) {
        for (int item = blockIdx.x; item < planes * rb; item += gridDim.x) {
          const int bt = item / rb, y0 = (item - bt * rb) * SM_ROWS;
          const int b = bt / w.Tm, t = bt - b * w.Tm;
          const SmoothMaps sm = smooth_maps(w, s, b, t, hw, false);
          for (int q = tid; q < groups * SM_ROWS; q += IWE_THREADS) {
            const int r = q / groups, y = y0 + r, x = (q - r * groups) * nx;
            if (y >= w.H) continue;
            if (w.vec) {
              float g[SM_NX];
              acc += smooth_strip<false, SM_NX>(sm, y, x, w.H, w.W, hw, g);
            } else {
              float g[1];
              acc += smooth_strip<false, 1>(sm, y, x, w.H, w.W, hw, g);
            }
          }
        }
      }
      const float r = block_sum256(acc, s_red);
      if (tid == 0) ws[l.smooth + (size_t)s * IWE_MAX_GRID + blockIdx.x] = r;
    }
  }
  grid_barrier(ctr + 0, gridDim.x);

  // ---- phase 1
  const float Tf = (float)w.T;
  const uint64_t pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
  const int total_items = (w.debug_skip & 2) ? 0 : w.S * w.B * w.n_items;
  for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    int sb, t, i0;
    decode_item(w, item, sb, t, i0);
    const int s = sb / w.B, b = sb - s * w.B;
    Ev ev[IWE_EPT];
    load_events(w, s, b, t, i0, hw, ev, pol_stream);
    float* base = img + (size_t)sb * 4 * plane;
#pragma unroll
    for (int k = 0; k < IWE_EPT; ++k) {
      if (!ev[k].on) continue;
      const float4 e = ev[k].e;
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const float tref = dir == 0 ? Tf : 0.f;
        const float tau = dir == 0 ? e.x : __fsub_rn(Tf, e.x);
        float yw, xw;
        Corner c[4];
        warp_event(e.x, e.y, e.z, ev[k].fy, ev[k].fx, tref, w.flow_scaling, w.H, w.W, yw, xw, c);
        float* d = base + (size_t)dir * 2 * plane;
#pragma unroll
        for (int r = 0; r < 2; ++r) {  // top row (corners 0,1), bottom row (2,3): left and right pixels are neighbours in memory
          const Corner &cl = c[2 * r], &cr = c[2 * r + 1];
          const float wl = cl.idx >= 0 ? __fmul_rn(cl.wy, cl.wx) : 0.f, wr = cr.idx >= 0 ? __fmul_rn(cr.wy, cr.wx) : 0.f;
          if (wl == 0.f && wr == 0.f) continue;
          const float tl = __fmul_rn(wl, tau), tr = __fmul_rn(wr, tau);
#pragma unroll
          for (int pol = 0; pol < 2; ++pol) {
            const float m = pol == 0 ? ev[k].pm.x : ev[k].pm.y;
            if (m == 0.f) continue;
            float* pl = d + (size_t)pol * plane;
            if (cl.idx >= 0 && cr.idx == cl.idx + 1) {
              // both pixels inside the image and neighbours in memory: one 16-byte reduction, into the copy whose pairs start at this parity
              float* q = (cl.idx & 1) ? pl + 2 * hw + (size_t)(cl.idx + 1) * 2 : pl + (size_t)cl.idx * 2;
              red4(q, __fmul_rn(wl, m), __fmul_rn(tl, m), __fmul_rn(wr, m), __fmul_rn(tr, m), pol_keep);
            } else {  // image border: one of the two pixels only
              if (wl != 0.f) red2(pl + (size_t)cl.idx * 2, __fmul_rn(wl, m), __fmul_rn(tl, m), pol_keep);
              if (wr != 0.f) red2(pl + (size_t)cr.idx * 2, __fmul_rn(wr, m), __fmul_rn(tr, m), pol_keep);
            }
          }
        }
      }
    }
  }
  grid_barrier(ctr + 1, gridDim.x);

  // ---- phase 2: chunks of RED_PIX pixels of one (scale, sample, direction)
  {
    const int chunks = (int)((hw + RED_PIX - 1) / RED_PIX), n_sbd = (w.debug_skip & 4) ? 0 : w.S * w.B * 2;
    for (int item = blockIdx.x; item < n_sbd * chunks; item += gridDim.x) {
      const int sbd = item / chunks, p0 = (item - sbd * chunks) * RED_PIX;
      const float2* pos = reinterpret_cast<const float2*>(img + (size_t)sbd * 2 * plane);
      const float2* neg = reinterpret_cast<const float2*>(img + (size_t)sbd * 2 * plane + plane);
      float ssq = 0.f, n = 0.f;
      const int p1 = min(p0 + RED_PIX, (int)hw);
      for (int i = p0 + tid; i < p1; i += IWE_THREADS) {
        const float2 qp = acc_px(pos, i, hw), qn = acc_px(neg, i, hw);
        const float ap = qp.y / (qp.x + 1e-9f) / Tf, an = qn.y / (qn.x + 1e-9f) / Tf;
        ssq += ap * ap + an * an;
        n += (qp.x + qn.x > 0.f) ? 1.f : 0.f;
      }
      const float r0 = block_sum256(ssq, s_red);
      const float r1 = block_sum256(n, s_red);
      if (tid == 0) {
        atomicAdd(sums + sbd * 2 + 0, r0);
        atomicAdd(sums + sbd * 2 + 1, r1);
      }
    }
  }

  // ---- phase 3: the last CTA to get here computes the scalar and restores the counters
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicInc(ctr + 2, gridDim.x - 1) == gridDim.x - 1;  // wraps to 0 with the last arrival
    __threadfence();
  }
  __syncthreads();
  if (!s_last) return;
  float total = 0.f;
  for (int s = 0; s < w.S; ++s) {
    float sm = 0.f;
    for (int c = tid; c < (int)gridDim.x; c += IWE_THREADS) sm += __ldcg(ws + l.smooth + (size_t)s * IWE_MAX_GRID + c);
    sm = block_sum256(sm, s_red);
    float fb = 0.f;
    for (int k = tid; k < w.B * 2; k += IWE_THREADS) {
      const float ssq = __ldcg(sums + ((size_t)s * w.B * 2 + k) * 2), n = __ldcg(sums + ((size_t)s * w.B * 2 + k) * 2 + 1);
      fb += w.loss_scaling ? ssq / n : ssq;
    }
    fb = block_sum256(fb, s_red);
    total += fb + w.smooth_coef * sm;
  }
  if (tid == 0) {
    loss[0] = total / (float)w.S;
    ctr[0] = 0u;
    ctr[1] = 0u;
  }
}

// d max(0, 1-|d|) / d d with torch's tie conventions: abs'(0) = 0; maximum splits the gradient at equality.
__device__ __forceinline__ float dweight(float d) {
  const float u = 1.0f - fabsf(d);
  const float s = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
  return u > 0.f ? s : (u == 0.f ? 0.5f * s : 0.f);
}

// ---- backward: ONE launch ---------------------------------------------------------------------------------------------
//   phase 0  gradient of the smoothness term wrt both flow channels of every pixel (gather form) -> g_flow (overwrites)
//   phase 1  per event and corner: the adjoint of the contrast term is computed on the fly from the accumulator images and
//            the per-(scale, sample, direction) sums, scattered onto the event's own pixel of g_flow
__global__ void __launch_bounds__(IWE_THREADS, EF_IWE_OCC_B) iwe_loss_bwd_kernel(const __grid_constant__ IweWin w, float* __restrict__ ws,
                                                                   const float* __restrict__ g_loss) {
  __shared__ bool s_last;
  const WsLayout l = ws_layout(w.S, w.B, w.H, w.W);
  unsigned int* ctr = reinterpret_cast<unsigned int*>(ws + l.ctr);
  const float* img = ws + l.img;
  float* adj = ws + l.adj;
  const float* sums = ws + l.sums;
  const size_t hw = (size_t)w.H * w.W;
  const size_t plane = plane_px(hw) * 2;
  const int tid = threadIdx.x;
  const float gl = __ldg(g_loss);
  const int W = w.W, H = w.H;

  // ---- phase 0a: smoothness gradient of every pixel (gather form) -> g_flow (both channels carry the same value)
  {
    const float coef = w.smooth_coef / (float)w.S * gl;
    const bool on = w.smooth_coef != 0.f && !(w.debug_skip & 1);
    const int rb = (H + SM_ROWS - 1) / SM_ROWS, planes = w.B * w.Tm;
    const int nx = w.vec ? SM_NX : 1, groups = W / nx;
    for (int item = blockIdx.x; item < w.S * planes * rb; item += gridDim.x) {
      const int s = item / (planes * rb), r0 = item - s * planes * rb;
      const int bt = r0 / rb, y0 = (r0 - bt * rb) * SM_ROWS;
      const int b = bt / w.Tm, t = bt - b * w.Tm;
      const SmoothMaps sm = smooth_maps(w, s, b, t, hw, true);
      float* g = w.g_flow[s * w.Tm + t] + (size_t)b * w.g_bs;
      for (int q = tid; q < groups * SM_ROWS; q += IWE_THREADS) {
        const int r = q / groups, y = y0 + r, x = (q - r * groups) * nx;
        if (y >= H) continue;
        const size_t o = (size_t)y * W + x;
        if (w.vec) {
          float gk[SM_NX] = {0.f, 0.f, 0.f, 0.f};
          if (on) smooth_strip<true, SM_NX>(sm, y, x, H, W, hw, gk);
          const float4 v = make_float4(gk[0] * coef, gk[1] * coef, gk[2] * coef, gk[3] * coef);
          *reinterpret_cast<float4*>(g + o) = v;
          *reinterpret_cast<float4*>(g + hw + o) = v;
        } else {
          float gk[1] = {0.f};
          if (on) smooth_strip<true, 1>(sm, y, x, H, W, hw, gk);
          g[o] = g[hw + o] = gk[0] * coef;
        }
      }
    }
  }
  // ---- phase 0b: adjoint images (one pixel pass): adj [sbd][pol][px][2 = dL/dI, dL/dTh]
  const float Tf = (float)w.T, inv_S = 1.0f / (float)w.S, inv_T = 1.0f / Tf;
  {
    const int chunks = (int)((hw + RED_PIX - 1) / RED_PIX), n_sbd = (w.debug_skip & 4) ? 0 : w.S * w.B * 2;
    for (int item = blockIdx.x; item < n_sbd * chunks; item += gridDim.x) {
      const int sbd = item / chunks, p0 = (item - sbd * chunks) * RED_PIX;
      const float2* pos = reinterpret_cast<const float2*>(img + (size_t)sbd * 2 * plane);
      const float2* neg = reinterpret_cast<const float2*>(img + (size_t)sbd * 2 * plane + plane);
      float2* apos = reinterpret_cast<float2*>(adj + (size_t)sbd * 4 * hw);
      float2* aneg = apos + hw;
      const float ssq = __ldcg(sums + sbd * 2), n = __ldcg(sums + sbd * 2 + 1);
      const float g = gl * inv_S / (w.loss_scaling ? n : 1.f);
      const float gn = -gl * inv_S * ssq / (n * n);  // through the divisor: empty pixels keep gradient 1 into it (in-place masked assignment)
      const int p1 = min(p0 + RED_PIX, (int)hw);
      for (int i = p0 + tid; i < p1; i += IWE_THREADS) {
        const float2 qp = acc_px(pos, i, hw), qn = acc_px(neg, i, hw);
        const float rp = __frcp_rn(qp.x + 1e-9f), rn = __frcp_rn(qn.x + 1e-9f);  // gradients are checked to 1e-3: one reciprocal per image
        const float ap = qp.y * rp * inv_T, an = qn.y * rn * inv_T;
        const float gtp = g * 2.f * ap * rp * inv_T, gtn = g * 2.f * an * rn * inv_T;
        float gip = -gtp * qp.y * rp, gin = -gtn * qn.y * rn;
        if (w.loss_scaling && !(qp.x + qn.x > 0.f)) {
          gip += gn;
          gin += gn;
        }
        apos[i] = make_float2(gip, gtp);
        aneg[i] = make_float2(gin, gtn);
      }
    }
  }
  grid_barrier(ctr + 3, gridDim.x);

  // ---- phase 1: per event and corner delta = dL/dI + tau dL/dTh of the event's polarity, chained through the bilinear weights
  const uint64_t pol_stream = l2_policy_stream();
  const int total_items = (w.debug_skip & 2) ? 0 : w.S * w.B * w.n_items;
  for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    int sb, t, i0;
    decode_item(w, item, sb, t, i0);
    const int s = sb / w.B, b = sb - s * w.B;
    Ev ev[IWE_EPT];
    load_events(w, s, b, t, i0, hw, ev, pol_stream);
    const int m_idx = s * w.Tm + (w.Tm > 1 ? t : 0);
    const float* abase = adj + (size_t)sb * 8 * hw;
    float* g_out = w.g_flow[m_idx] + (size_t)b * w.g_bs;
#pragma unroll
    for (int k = 0; k < IWE_EPT; ++k) {
      if (!ev[k].on) continue;
      const float4 e = ev[k].e;
      const float2 pm = ev[k].pm;
      float gfy = 0.f, gfx = 0.f;
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const float tref = dir == 0 ? Tf : 0.f;
        const float tau = dir == 0 ? e.x : (Tf - e.x);
        const float kk = (tref - e.x) * w.flow_scaling;
        float yw, xw;
        Corner c[4];
        warp_event(e.x, e.y, e.z, ev[k].fy, ev[k].fx, tref, w.flow_scaling, H, W, yw, xw, c);
        const float2* apos = reinterpret_cast<const float2*>(abase + (size_t)dir * 4 * hw);
        const float2* aneg = apos + hw;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (c[q].idx < 0) continue;
          const float dwy = dweight(c[q].dy) * c[q].wx, dwx = c[q].wy * dweight(c[q].dx);
          if (dwy == 0.f && dwx == 0.f) continue;
          float delta = 0.f;
          if (pm.x != 0.f) {
            const float2 a2 = __ldcg(apos + c[q].idx);
            delta += pm.x * (a2.x + tau * a2.y);
          }
          if (pm.y != 0.f) {
            const float2 a2 = __ldcg(aneg + c[q].idx);
            delta += pm.y * (a2.x + tau * a2.y);
          }
          gfy += delta * dwy * kk;
          gfx += delta * dwx * kk;
        }
      }
      if (gfx != 0.f) atomicAdd(g_out + ev[k].pix, gfx);
      if (gfy != 0.f) atomicAdd(g_out + hw + ev[k].pix, gfy);
    }
  }

  // ---- restore the counters: the last CTA to finish knows everyone has left the barrier
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicInc(ctr + 4, gridDim.x - 1) == gridDim.x - 1;
    if (s_last) ctr[3] = 0u;
  }
}

// ---- per-polarity IWE image (utils/iwe.py:95-153) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) iwe_image_kernel(const ef_iwe_image_params p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= p.N) return;
  const size_t hw = (size_t)p.H * p.W;
  const float4 e = reinterpret_cast<const float4*>(p.events)[(size_t)b * p.N + i];
  const float2 pm = reinterpret_cast<const float2*>(p.pol_mask)[(size_t)b * p.N + i];
  float fy, fx;
  if (p.event_flow) {
    const float2 f = reinterpret_cast<const float2*>(p.event_flow)[(size_t)b * p.N + i];
    fy = f.x, fx = f.y;
  } else {
    const int pix = (int)(e.y * (float)p.W + e.z);
    fx = p.flow[(size_t)b * 2 * hw + pix];
    fy = p.flow[(size_t)b * 2 * hw + hw + pix];
  }
  float* out = p.iwe + (size_t)b * 2 * hw;
  if (p.round_idx) {
    const float dt = __fsub_rn(p.tref, e.x);
    const float yw = rintf(__fadd_rn(e.y, __fmul_rn(__fmul_rn(dt, fy), p.flow_scaling)));  // torch.round: half to even
    const float xw = rintf(__fadd_rn(e.z, __fmul_rn(__fmul_rn(dt, fx), p.flow_scaling)));
    const bool oob = yw < 0.f || yw >= (float)p.H || xw < 0.f || xw >= (float)p.W;
    // purge_unfeasible zeroes the weight AND redirects the index to pixel 0: adds +0 there, i.e. nothing
    if (!oob) {
      const int idx = (int)(yw * (float)p.W + xw);
      if (pm.x != 0.f) atomicAdd(out + idx, pm.x);
      if (pm.y != 0.f) atomicAdd(out + hw + idx, pm.y);
    }
  } else {
    float yw, xw;
    Corner c[4];
    warp_event(e.x, e.y, e.z, fy, fx, p.tref, p.flow_scaling, p.H, p.W, yw, xw, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c[k].idx < 0) continue;
      const float w = __fmul_rn(c[k].wy, c[k].wx);
      if (w == 0.f) continue;
      if (pm.x != 0.f) atomicAdd(out + c[k].idx, __fmul_rn(w, pm.x));
      if (pm.y != 0.f) atomicAdd(out + hw + c[k].idx, __fmul_rn(w, pm.y));
    }
  }
}

// ---- validation metrics: FWL / RSAT (loss/flow.py:468-579) and AEE (:582-628) -----------------------------------------
// workspace: img [B][2 = warped, unwarped][HW][4 = I+,Th+,I-,Th-] (pixel-interleaved), then extra [B][2][HW] (what FWL counts
// beyond I+ + I-), then sums [B][2][4 = sum A^2, n, sum I, sum I^2]
__global__ void __launch_bounds__(256) iwe_metric_scatter_kernel(const ef_iwe_metrics_params p, float* __restrict__ img, float* __restrict__ extra) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= p.n_total) return;
  const size_t hw = (size_t)p.H * p.W;
  const float4 e = reinterpret_cast<const float4*>(p.events)[(size_t)b * p.n_total + i];
  const float2 pm = reinterpret_cast<const float2*>(p.pol_mask)[(size_t)b * p.n_total + i];
  int t_e = 0;
  if (p.T_maps > 1) {
    if (p.pass_offsets) {
      while (t_e + 1 < p.T && i >= __ldg(p.pass_offsets + t_e + 1)) ++t_e;
    } else {
      t_e = min(i / p.n_per_pass, p.T_maps - 1);
    }
  }
  const int pix = (int)(e.y * (float)p.W + e.z);
  const float* fm = p.flow_maps + ((size_t)b * p.T_maps + t_e) * 2 * hw;
  const float fx = __ldg(fm + pix), fy = __ldg(fm + hw + pix);
  float* base = img + (size_t)b * 8 * hw;
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // k = 0: warped to tref = T with the flow; k = 1: flow * 0 (image of events)
    const float dt = __fsub_rn((float)p.T, e.x);
    const float yw = rintf(__fadd_rn(e.y, __fmul_rn(__fmul_rn(dt, k == 0 ? fy : __fmul_rn(fy, 0.f)), p.flow_scaling)));
    const float xw = rintf(__fadd_rn(e.z, __fmul_rn(__fmul_rn(dt, k == 0 ? fx : __fmul_rn(fx, 0.f)), p.flow_scaling)));
    if (yw < 0.f || yw >= (float)p.H || xw < 0.f || xw >= (float)p.W) continue;
    const int idx = (int)(yw * (float)p.W + xw);
    float* d = base + (size_t)k * 4 * hw;
    // RSAT scatters per polarity (weights = the polarity mask); FWL scatters weight 1 per event regardless of polarity
    // (loss/flow.py:488,494: interpolate without a mask).  The FWL image is I+ + I- + extra with extra = 1 - pm.x - pm.y: zero for
    // the loaders' one-hot masks, 1 for padded / p = 0 events, so that every event counts exactly once.
    float2* q = reinterpret_cast<float2*>(d + (size_t)idx * 4);
    if (pm.x != 0.f) atomicAdd(q, make_float2(pm.x, __fmul_rn(e.x, pm.x)));
    if (pm.y != 0.f) atomicAdd(q + 1, make_float2(pm.y, __fmul_rn(e.x, pm.y)));
    const float ex = 1.0f - pm.x - pm.y;
    if (ex != 0.f) atomicAdd(extra + ((size_t)b * 2 + k) * hw + idx, ex);
  }
}

__global__ void __launch_bounds__(256) iwe_metric_reduce_kernel(const float* __restrict__ img, const float* __restrict__ extra,
                                                                float* __restrict__ sums, int HW, float T) {
  __shared__ float s_red[8];
  const int bk = blockIdx.y;  // b*2 + k
  const float* d = img + (size_t)bk * 4 * HW;
  const float* ex = extra + (size_t)bk * HW;
  float ssq = 0.f, n = 0.f, s1 = 0.f, s2 = 0.f;
  const int p0 = blockIdx.x * RED_PIX;
  for (int i = p0 + threadIdx.x; i < min(p0 + RED_PIX, HW); i += 256) {
    const float4 q = reinterpret_cast<const float4*>(d)[i];
    const float ip = q.x, tp = q.y, in = q.z, tn = q.w;
    const float ap = tp / (ip + 1e-9f) / T, an = tn / (in + 1e-9f) / T;
    ssq += ap * ap + an * an;
    n += (ip + in > 0.f) ? 1.f : 0.f;
    const float cnt = ip + in + ex[i];
    s1 += cnt;
    s2 += cnt * cnt;
  }
  const float r0 = block_sum256(ssq, s_red), r1 = block_sum256(n, s_red), r2 = block_sum256(s1, s_red), r3 = block_sum256(s2, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(sums + bk * 4 + 0, r0);
    atomicAdd(sums + bk * 4 + 1, r1);
    atomicAdd(sums + bk * 4 + 2, r2);
    atomicAdd(sums + bk * 4 + 3, r3);
  }
}

__global__ void iwe_metric_finalize_kernel(const float* __restrict__ sums, int B, int HW, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* w = sums + (size_t)(b * 2 + 0) * 4;
  const float* z = sums + (size_t)(b * 2 + 1) * 4;
  const float n = (float)HW;
  const float var_w = (w[3] - w[2] * w[2] / n) / (n - 1.f), var_z = (z[3] - z[2] * z[2] / n) / (n - 1.f);  // torch.var: unbiased
  const float rs_w = w[0] / w[1], rs_z = z[0] / z[1];
  out[b * 4 + 0] = var_w / var_z;
  out[b * 4 + 1] = rs_w / rs_z;
  out[b * 4 + 2] = rs_w;
  out[b * 4 + 3] = rs_z;
}

__global__ void __launch_bounds__(256) aee_kernel(const ef_aee_params p, float* __restrict__ ws) {
  __shared__ float s_red[8];
  const int b = blockIdx.y;
  const size_t hw = (size_t)p.H * p.W;
  const float k = p.flow_scaling * p.dt_ratio[b];
  float se = 0.f, sn = 0.f, so = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < hw; i += (size_t)gridDim.x * 256) {
    const float fx = p.flow[((size_t)b * 2) * hw + i] * k, fy = p.flow[((size_t)b * 2 + 1) * hw + i] * k;
    const float gx = p.gtflow[((size_t)b * 2) * hw + i], gy = p.gtflow[((size_t)b * 2 + 1) * hw + i];
    const bool valid = (p.event_mask[(size_t)b * hw + i] != 0.f) && !(gx == 0.f && gy == 0.f);
    const float m = valid ? 1.f : 0.f;
    const float err = sqrtf((fx - gx) * (fx - gx) + (fy - gy) * (fy - gy)) * m;
    const float mag = sqrtf(fx * fx + fy * fy) * m;
    se += err;
    sn += m;
    so += (err > 3.0f && err > 0.05f * mag) ? 1.f : 0.f;
  }
  const float r0 = block_sum256(se, s_red), r1 = block_sum256(sn, s_red), r2 = block_sum256(so, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(ws + b, r0);
    atomicAdd(ws + p.B + b, r1);
    atomicAdd(ws + 2 * p.B, r2);  // outliers are summed over the whole batch (loss/flow.py:625)
  }
}

__global__ void aee_finalize_kernel(const float* __restrict__ ws, int B, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  out[b] = ws[b] / (ws[B + b] + 1e-9f);
  out[B + b] = ws[2 * B] / (ws[B + b] + 1e-9f);
}

// ---- host side: lowering of the two public window forms, launch ------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static int finish_window(IweWin& w, const char* who) {
  EF_REQUIRE(w.S > 0 && w.S <= MAXS && w.B > 0 && w.T > 0 && w.T <= MAXP && w.H > 0 && w.W > 0, EF_EINVAL,
             "%s: bad dimensions (S <= %d, T <= %d)", who, MAXS, MAXP);
  EF_REQUIRE(w.Tm == 1 || w.Tm == w.T, EF_EINVAL, "%s: T_maps must be 1 (overwrite_intermediate) or T", who);
  EF_REQUIRE(((size_t)w.H * w.W) % 2 == 0, EF_EUNSUPPORTED, "%s: H*W must be even (16-byte vector accumulators)", who);
  w.chunk_off[0] = 0;
  for (int t = 0; t < w.T; ++t) {
    EF_REQUIRE(w.n_pass[t] >= 0, EF_EINVAL, "%s: negative event count in pass %d", who, t);
    EF_REQUIRE(w.n_pass[t] == 0 || (w.ev[t] && w.pm[t]), EF_ENULL, "%s: NULL events / pol_mask of pass %d", who, t);
    w.chunk_off[t + 1] = w.chunk_off[t] + cdiv(w.n_pass[t], IWE_CHUNK);
  }
  w.n_items = w.chunk_off[w.T];
  // 16-byte accesses of the pixel passes: rows of 4-pixel groups, every plane 16-byte aligned
  auto al = [](const void* q, long long stride_floats) { return ((uintptr_t)q % 16 == 0) && (stride_floats % 4 == 0); };
  w.vec = w.W % 4 == 0;
  for (int i = 0; i < w.S * w.Tm && w.vec; ++i) w.vec = al(w.flow[i], w.flow_bs) && (!w.g_flow[i] || al(w.g_flow[i], w.g_bs));
  for (int t = 0; t < w.Tm && w.vec && w.use_mask; ++t) w.vec = al(w.mask[t], w.mask_bs);
  static const int dbg = env_int("EF_IWE_SKIP", 0);
  w.debug_skip = dbg;
  for (int i = 0; i < w.S * w.Tm; ++i) EF_REQUIRE(w.flow[i], EF_ENULL, "%s: NULL flow map", who);
  if (w.use_mask)
    for (int t = 0; t < w.Tm; ++t) EF_REQUIRE(w.mask[t], EF_ENULL, "%s: smoothing_mask without event_mask", who);
  return EF_OK;
}

static int lower(const ef_iwe_loss_params& p, IweWin& w, const char* who) {
  memset(&w, 0, sizeof(w));
  EF_REQUIRE(p.T > 0 && p.T <= MAXP, EF_EINVAL, "%s: T must be in [1, %d]", who, MAXP);
  EF_REQUIRE(p.T_maps == (p.overwrite_intermediate ? 1 : p.T), EF_EINVAL, "%s: T_maps must be 1 with overwrite_intermediate, else T", who);
  EF_REQUIRE(p.n_total >= 0 && (p.T_maps == 1 || p.pass_offsets || (p.n_per_pass > 0 && (int64_t)p.n_per_pass * p.T >= p.n_total)), EF_EINVAL,
             "%s: n_per_pass * T must cover n_total", who);
  EF_REQUIRE((p.n_total == 0 || (p.events && p.pol_mask)) && p.flow_maps && p.workspace, EF_ENULL, "%s: NULL tensor", who);
  w.S = p.S, w.B = p.B, w.T = p.T, w.Tm = p.T_maps, w.H = p.H, w.W = p.W;
  w.loss_scaling = p.loss_scaling, w.use_mask = p.smoothing_mask != 0, w.use_dt = !p.overwrite_intermediate;
  w.flow_scaling = p.flow_scaling;
  w.smooth_coef = p.weight / (p.overwrite_intermediate ? 4.f : 5.f) / (float)p.T_maps;
  const size_t hw = (size_t)p.H * p.W;
  for (int t = 0; t < p.T; ++t) {
    int o0, o1;
    if (p.pass_offsets) o0 = p.pass_offsets[t], o1 = p.pass_offsets[t + 1];
    else if (p.T_maps == 1 && p.n_per_pass <= 0) o0 = t == 0 ? 0 : p.n_total, o1 = p.n_total;  // one map: the pass structure is irrelevant
    else o0 = min(t * p.n_per_pass, p.n_total), o1 = t + 1 == p.T ? p.n_total : min((t + 1) * p.n_per_pass, p.n_total);
    EF_REQUIRE(o0 >= 0 && o1 >= o0 && o1 <= p.n_total, EF_EINVAL, "%s: bad pass_offsets", who);
    w.n_pass[t] = o1 - o0;
    w.ev[t] = p.events ? p.events + (size_t)o0 * 4 : nullptr, w.pm[t] = p.pol_mask ? p.pol_mask + (size_t)o0 * 2 : nullptr;
    w.ev_bs[t] = (long long)p.n_total * 4, w.pm_bs[t] = (long long)p.n_total * 2;
  }
  for (int s = 0; s < p.S && s < MAXS; ++s)
    for (int m = 0; m < p.T_maps; ++m) {
      w.flow[s * p.T_maps + m] = p.flow_maps + ((size_t)s * p.B * p.T_maps + m) * 2 * hw;
      if (p.g_flow_maps) w.g_flow[s * p.T_maps + m] = p.g_flow_maps + ((size_t)s * p.B * p.T_maps + m) * 2 * hw;
    }
  w.flow_bs = w.g_bs = (long long)p.T_maps * 2 * hw;
  for (int m = 0; m < p.T_maps; ++m) w.mask[m] = p.event_mask ? p.event_mask + (size_t)m * hw : nullptr;
  w.mask_bs = (long long)p.T_maps * hw;
  return finish_window(w, who);
}

static int lower(const ef_iwe_loss_pass_params& p, IweWin& w, const char* who) {
  memset(&w, 0, sizeof(w));
  EF_REQUIRE(p.T > 0 && p.T <= MAXP && p.S > 0 && p.S <= MAXS, EF_EINVAL, "%s: T must be in [1, %d], S in [1, %d]", who, MAXP, MAXS);
  EF_REQUIRE(p.T_maps == (p.overwrite_intermediate ? 1 : p.T), EF_EINVAL, "%s: T_maps must be 1 with overwrite_intermediate, else T", who);
  EF_REQUIRE(p.workspace, EF_ENULL, "%s: NULL workspace", who);
  w.S = p.S, w.B = p.B, w.T = p.T, w.Tm = p.T_maps, w.H = p.H, w.W = p.W;
  w.loss_scaling = p.loss_scaling, w.use_mask = p.smoothing_mask != 0, w.use_dt = !p.overwrite_intermediate;
  w.flow_scaling = p.flow_scaling;
  w.smooth_coef = p.weight / (p.overwrite_intermediate ? 4.f : 5.f) / (float)p.T_maps;
  const size_t hw = (size_t)p.H * p.W;
  for (int t = 0; t < p.T; ++t) {
    w.n_pass[t] = p.n_pass[t];
    w.ev[t] = p.events[t], w.pm[t] = p.pol_mask[t];
    w.ev_bs[t] = (long long)p.n_pass[t] * 4, w.pm_bs[t] = (long long)p.n_pass[t] * 2;
  }
  for (int i = 0; i < p.S * p.T_maps; ++i) w.flow[i] = p.flow[i], w.g_flow[i] = p.g_flow[i];
  w.flow_bs = w.g_bs = (long long)2 * hw;
  for (int m = 0; m < p.T_maps; ++m) w.mask[m] = p.event_mask[m];
  w.mask_bs = (long long)hw;
  return finish_window(w, who);
}

// Co-resident grid: the kernels synchronise all their CTAs, so the launch carries the cooperative attribute (it fails instead
// of dead-locking should the grid not fit) and the grid is capped by the occupancy of the device.
template <typename K>
static int coop_grid(K kernel, long long work_ctas, int* grid) {
  static int cap = 0;
  if (cap == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, IWE_THREADS, 0) != cudaSuccess || per_sm < 1) return check_launch("occupancy query");
    cap = sms * (per_sm < 6 ? per_sm : 6);
    if (cap > IWE_MAX_GRID) cap = IWE_MAX_GRID;
  }
  long long g = work_ctas < cap ? work_ctas : cap;
  *grid = (int)(g < 1 ? 1 : g);
  return EF_OK;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_coop(void (*kernel)(KArgs...), int grid, cudaStream_t st, Args... args) {
  static const int coop = env_int("EF_IWE_COOP", 1);  // 0: plain launch (the grid is capped to the resident capacity either way)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(IWE_THREADS), cfg.dynamicSmemBytes = 0, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr, cfg.numAttrs = coop ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static long long loss_work_ctas(const IweWin& w) {
  static const int force = env_int("EF_IWE_GRID", 0);
  if (force > 0) return force;
  // enough CTAs for the event chunks and for one pass over the pixels, whichever is larger; small windows get small grids
  // (the cost of a grid barrier grows with the number of CTAs)
  const long long ev = (long long)w.S * w.B * w.n_items;
  const long long px = ((long long)w.S * w.B * 2 * w.H * w.W + 8191) / 8192;
  return ev > px ? ev : px;
}

static int run_fwd(const IweWin& w, float* ws, float* loss, cudaStream_t st) {
  int grid, rc;
  if ((rc = coop_grid(iwe_loss_fwd_kernel, loss_work_ctas(w), &grid))) return rc;
  launch_coop(iwe_loss_fwd_kernel, grid, st, w, ws, loss);
  return check_launch("iwe_loss_fwd_kernel");
}

static int run_bwd(const IweWin& w, float* ws, const float* g_loss, cudaStream_t st) {
  for (int i = 0; i < w.S * w.Tm; ++i) EF_REQUIRE(w.g_flow[i], EF_ENULL, "ef_iwe_loss_bwd: NULL g_flow");
  int grid, rc;
  if ((rc = coop_grid(iwe_loss_bwd_kernel, loss_work_ctas(w), &grid))) return rc;
  launch_coop(iwe_loss_bwd_kernel, grid, st, w, ws, g_loss);
  return check_launch("iwe_loss_bwd_kernel");
}

}  // namespace ef

extern "C" int64_t ef_iwe_loss_workspace_elems(int32_t S, int32_t B, int32_t H, int32_t W) {
  return (int64_t)ef::ws_layout(S, B, H, W).total;
}

extern "C" int ef_iwe_loss_fwd(const ef_iwe_loss_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_loss_fwd: params is NULL");
  IweWin w;
  if (int rc = lower(*pp, w, "ef_iwe_loss_fwd")) return rc;
  EF_REQUIRE(pp->loss, EF_ENULL, "ef_iwe_loss_fwd: loss is NULL");
  return run_fwd(w, pp->workspace, pp->loss, as_stream(stream));
}

extern "C" int ef_iwe_loss_bwd(const ef_iwe_loss_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_loss_bwd: params is NULL");
  EF_REQUIRE(pp->g_loss && pp->g_flow_maps, EF_ENULL, "ef_iwe_loss_bwd: g_loss / g_flow_maps is NULL");
  IweWin w;
  if (int rc = lower(*pp, w, "ef_iwe_loss_bwd")) return rc;
  return run_bwd(w, pp->workspace, pp->g_loss, as_stream(stream));
}

extern "C" int ef_iwe_loss_fwd_passes(const ef_iwe_loss_pass_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_loss_fwd_passes: params is NULL");
  IweWin w;
  if (int rc = lower(*pp, w, "ef_iwe_loss_fwd_passes")) return rc;
  EF_REQUIRE(pp->loss, EF_ENULL, "ef_iwe_loss_fwd_passes: loss is NULL");
  return run_fwd(w, pp->workspace, pp->loss, as_stream(stream));
}

extern "C" int ef_iwe_loss_bwd_passes(const ef_iwe_loss_pass_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_loss_bwd_passes: params is NULL");
  EF_REQUIRE(pp->g_loss, EF_ENULL, "ef_iwe_loss_bwd_passes: g_loss is NULL");
  IweWin w;
  if (int rc = lower(*pp, w, "ef_iwe_loss_bwd_passes")) return rc;
  return run_bwd(w, pp->workspace, pp->g_loss, as_stream(stream));
}

extern "C" int ef_iwe_image(const ef_iwe_image_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_image: params is NULL");
  const ef_iwe_image_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.N >= 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_iwe_image: bad dimensions");
  EF_REQUIRE(p.iwe && (p.N == 0 || (p.events && p.pol_mask && (p.flow || p.event_flow))), EF_ENULL, "ef_iwe_image: NULL tensor");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(p.iwe, 0, (size_t)p.B * 2 * p.H * p.W * sizeof(float), st);
  if (p.N == 0) return EF_OK;
  iwe_image_kernel<<<dim3(cdiv(p.N, 256), p.B), 256, 0, st>>>(p);
  return check_launch("iwe_image_kernel");
}

extern "C" int64_t ef_iwe_metrics_workspace_elems(int32_t B, int32_t H, int32_t W) { return (int64_t)B * 10 * H * W + (int64_t)B * 8; }

extern "C" int ef_iwe_metrics(const ef_iwe_metrics_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_metrics: params is NULL");
  const ef_iwe_metrics_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.T > 0 && p.H > 0 && p.W > 0 && p.n_total >= 0, EF_EINVAL, "ef_iwe_metrics: bad dimensions");
  EF_REQUIRE(p.T_maps == p.T || p.T_maps == 1, EF_EINVAL, "ef_iwe_metrics: T_maps must be T or 1");
  EF_REQUIRE(p.T_maps == 1 || p.pass_offsets || p.n_per_pass > 0, EF_EINVAL, "ef_iwe_metrics: n_per_pass or pass_offsets needed");
  EF_REQUIRE(p.workspace && p.out && (p.n_total == 0 || (p.events && p.pol_mask && p.flow_maps)), EF_ENULL, "ef_iwe_metrics: NULL tensor");
  cudaStream_t st = as_stream(stream);
  const int HW = p.H * p.W;
  float* img = p.workspace;
  float* extra = p.workspace + (size_t)p.B * 8 * HW;
  float* sums = extra + (size_t)p.B * 2 * HW;
  cudaMemsetAsync(p.workspace, 0, ((size_t)p.B * 10 * HW + (size_t)p.B * 8) * sizeof(float), st);
  int rc;
  if (p.n_total > 0) {
    iwe_metric_scatter_kernel<<<dim3(cdiv(p.n_total, 256), p.B), 256, 0, st>>>(p, img, extra);
    if ((rc = check_launch("iwe_metric_scatter_kernel"))) return rc;
  }
  iwe_metric_reduce_kernel<<<dim3(cdiv(HW, RED_PIX), p.B * 2), 256, 0, st>>>(img, extra, sums, HW, (float)p.T);
  if ((rc = check_launch("iwe_metric_reduce_kernel"))) return rc;
  iwe_metric_finalize_kernel<<<cdiv(p.B, 64), 64, 0, st>>>(sums, p.B, HW, p.out);
  return check_launch("iwe_metric_finalize_kernel");
}

extern "C" int ef_aee(const ef_aee_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_aee: params is NULL");
  const ef_aee_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_aee: bad dimensions");
  EF_REQUIRE(p.flow && p.gtflow && p.event_mask && p.dt_ratio && p.workspace && p.out, EF_ENULL, "ef_aee: NULL tensor");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(p.workspace, 0, (2 * (size_t)p.B + 1) * sizeof(float), st);
  const int blocks = cdiv(p.H * p.W, 256 * 8);
  aee_kernel<<<dim3(blocks, p.B), 256, 0, st>>>(p, p.workspace);
  int rc;
  if ((rc = check_launch("aee_kernel"))) return rc;
  aee_finalize_kernel<<<cdiv(p.B, 64), 64, 0, st>>>(p.workspace, p.B, p.out);
  return check_launch("aee_finalize_kernel");
}
