// Contrast-maximisation event-warping loss (hot path B): forward, analytic backward, and the per-polarity IWE image.
// Reference: loss/flow.py:176-301 (EventWarping.forward), utils/iwe.py:4-92 (purge_unfeasible, get_interpolation,
// interpolate), loss/flow.py:65-79 (per-event flow gather), utils/iwe.py:95-153 (deblur_events, compute_pol_iwe).
// Sub-gradient conventions at ties (integer warped coordinates, empty pixels) follow SURVEY.md 7.4, i.e. what
// torch.autograd produces for the reference expressions.
#include "common.cuh"

namespace ef {

// workspace layout (floats):
//   img  [S][B][2 dir][HW][4 = I+,Th+,I-,Th-]      forward accumulators, pixel-interleaved: an event adds (w, w*tau) of its
//                                                  polarity with ONE 8-byte vector atomic per corner (red.global.add.v2.f32),
//                                                  the reductions read one 16-byte vector per pixel
//   adj  [S][B][2 dir][HW][4]                      adjoint images (backward), same layout
//   sums [S][B][2 dir][2 = sum A^2, n]             per-sample reductions
//   smooth [S]
struct WsLayout {
  size_t img, adj, sums, smooth, total;
};
__host__ __device__ inline WsLayout ws_layout(int S, int B, int H, int W) {
  WsLayout l;
  const size_t hw = (size_t)H * W;
  l.img = 0;
  l.adj = l.img + (size_t)S * B * 8 * hw;
  l.sums = l.adj + (size_t)S * B * 8 * hw;
  l.smooth = l.sums + (size_t)S * B * 4;
  l.total = l.smooth + S;
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

__device__ __forceinline__ int pass_of_event(const ef_iwe_loss_params& p, int i) {
  if (p.T_maps <= 1) return 0;
  if (p.pass_offsets) {
    int t = 0;
    while (t + 1 < p.T && i >= __ldg(p.pass_offsets + t + 1)) ++t;
    return t;
  }
  return min(i / p.n_per_pass, p.T_maps - 1);
}

// ---- forward: scatter ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iwe_scatter_kernel(const ef_iwe_loss_params p, float* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y, s = blockIdx.z;
  if (i >= p.n_total) return;
  const size_t hw = (size_t)p.H * p.W;
  const float4 e = reinterpret_cast<const float4*>(p.events)[(size_t)b * p.n_total + i];  // ts, y, x, p
  const float2 pm = reinterpret_cast<const float2*>(p.pol_mask)[(size_t)b * p.n_total + i];
  if (pm.x == 0.f && pm.y == 0.f) return;
  const int t_e = pass_of_event(p, i);
  const int pix = (int)(e.y * (float)p.W + e.z);
  const float* fm = p.flow_maps + (((size_t)s * p.B + b) * p.T_maps + t_e) * 2 * hw;
  const float fx = __ldg(fm + pix), fy = __ldg(fm + hw + pix);
  float* base = img + ((size_t)s * p.B + b) * 8 * hw;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    const float tref = dir == 0 ? (float)p.T : 0.f;
    const float tau = dir == 0 ? e.x : __fsub_rn((float)p.T, e.x);
    float yw, xw;
    Corner c[4];
    warp_event(e.x, e.y, e.z, fy, fx, tref, p.flow_scaling, p.H, p.W, yw, xw, c);
    float* d = base + (size_t)dir * 4 * hw;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c[k].idx < 0) continue;
      const float w = __fmul_rn(c[k].wy, c[k].wx);
      if (w == 0.f) continue;
      const float wt = __fmul_rn(w, tau);
      float2* q = reinterpret_cast<float2*>(d + (size_t)c[k].idx * 4);  // [I+, Th+], [I-, Th-]
      if (pm.x != 0.f) atomicAdd(q, make_float2(__fmul_rn(w, pm.x), __fmul_rn(wt, pm.x)));
      if (pm.y != 0.f) atomicAdd(q + 1, make_float2(__fmul_rn(w, pm.y), __fmul_rn(wt, pm.y)));
    }
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

// ---- forward: per-(scale,sample,direction) contrast sums --------------------------------------------------------
constexpr int RED_PIX = 2048;  // pixels per block
__global__ void __launch_bounds__(256) iwe_reduce_kernel(const float* __restrict__ img, float* __restrict__ sums, int HW, float T) {
  __shared__ float s_red[8];
  const int sbd = blockIdx.y;  // (s*B + b)*2 + dir
  const float* d = img + (size_t)sbd * 4 * HW;
  float ssq = 0.f, n = 0.f;
  const int p0 = blockIdx.x * RED_PIX;
  for (int i = p0 + threadIdx.x; i < min(p0 + RED_PIX, HW); i += 256) {
    const float4 q = reinterpret_cast<const float4*>(d)[i];
    const float ip = q.x, tp = q.y, in = q.z, tn = q.w;
    const float ap = tp / (ip + 1e-9f) / T, an = tn / (in + 1e-9f) / T;
    ssq += ap * ap + an * an;
    n += (ip + in > 0.f) ? 1.f : 0.f;
  }
  const float r0 = block_sum256(ssq, s_red);
  const float r1 = block_sum256(n, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(sums + sbd * 2 + 0, r0);
    atomicAdd(sums + sbd * 2 + 1, r1);
  }
}

// ---- smoothness (loss/flow.py:262-294), forward value and gradient ----------------------------------------------
struct SmoothGeom {
  int B, T, H, W;
  bool use_mask, use_dt;
};

__device__ __forceinline__ float charb_pair(const float* fxm, const float* fym, const float* mk, size_t a, size_t b, bool use_mask,
                                            float& dcoef) {
  const float d = (fxm[a] - fxm[b]) + (fym[a] - fym[b]);
  const float c = sqrtf(d * d + 1e-6f);
  const float m = use_mask ? mk[a] * mk[b] : 1.f;
  dcoef = m * d / c;  // d(term)/d f[a] (both channels); minus for f[b]
  return m * c;
}

// flow maps of one scale: [B][T][2][H][W]; mask [B][T][H][W]
__global__ void __launch_bounds__(256) smooth_fwd_kernel(const float* __restrict__ fm, const float* __restrict__ mask, SmoothGeom g,
                                                         float* __restrict__ out) {
  __shared__ float s_red[8];
  const size_t hw = (size_t)g.H * g.W, n = (size_t)g.B * g.T * hw;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int x = i % g.W, y = (i / g.W) % g.H;
    const size_t bt = i / hw;
    const int t = bt % g.T;
    const float* fxm = fm + bt * 2 * hw;
    const float* fym = fxm + hw;
    const float* mk = mask ? mask + bt * hw : nullptr;
    const size_t o = (size_t)y * g.W + x;
    float dc;
    if (x + 1 < g.W) acc += charb_pair(fxm, fym, mk, o, o + 1, g.use_mask, dc);
    if (y + 1 < g.H) acc += charb_pair(fxm, fym, mk, o, o + g.W, g.use_mask, dc);
    if (x + 1 < g.W && y + 1 < g.H) {
      acc += charb_pair(fxm, fym, mk, o, o + g.W + 1, g.use_mask, dc);      // down-right
      acc += charb_pair(fxm, fym, mk, o + g.W, o + 1, g.use_mask, dc);      // up-right: (y+1,x) - (y,x+1)
    }
    if (g.use_dt && t + 1 < g.T) {  // temporal: same pixel, next pass.  Masks of both passes.
      const float d = (fxm[o] - fxm[o + 2 * hw]) + (fym[o] - fym[o + 2 * hw]);
      const float m = g.use_mask ? mk[o] * mk[o + hw] : 1.f;
      acc += m * sqrtf(d * d + 1e-6f);
    }
  }
  const float r = block_sum256(acc, s_red);
  if (threadIdx.x == 0) atomicAdd(out, r);
}

// gradient of the smoothness term wrt both flow channels of every pixel (gather form, no atomics); writes g_fm.
__global__ void __launch_bounds__(256) smooth_bwd_kernel(const float* __restrict__ fm, const float* __restrict__ mask, SmoothGeom g,
                                                         const float* __restrict__ g_loss, float coef, float* __restrict__ g_fm) {
  const size_t hw = (size_t)g.H * g.W, n = (size_t)g.B * g.T * hw;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int x = i % g.W, y = (i / g.W) % g.H;
  const size_t bt = i / hw;
  const int t = bt % g.T;
  const float* fxm = fm + bt * 2 * hw;
  const float* fym = fxm + hw;
  const float* mk = mask ? mask + bt * hw : nullptr;
  const size_t o = (size_t)y * g.W + x;
  const int W = g.W, H = g.H;
  float acc = 0.f, dc;
  // as first element of a pair: +, as second: -
  if (x + 1 < W) { charb_pair(fxm, fym, mk, o, o + 1, g.use_mask, dc); acc += dc; }
  if (x >= 1) { charb_pair(fxm, fym, mk, o - 1, o, g.use_mask, dc); acc -= dc; }
  if (y + 1 < H) { charb_pair(fxm, fym, mk, o, o + W, g.use_mask, dc); acc += dc; }
  if (y >= 1) { charb_pair(fxm, fym, mk, o - W, o, g.use_mask, dc); acc -= dc; }
  if (x + 1 < W && y + 1 < H) { charb_pair(fxm, fym, mk, o, o + W + 1, g.use_mask, dc); acc += dc; }
  if (x >= 1 && y >= 1) { charb_pair(fxm, fym, mk, o - W - 1, o, g.use_mask, dc); acc -= dc; }
  if (y >= 1 && x + 1 < W) { charb_pair(fxm, fym, mk, o, o - W + 1, g.use_mask, dc); acc += dc; }   // first of up-right pair (y-1,x)
  if (y + 1 < H && x >= 1) { charb_pair(fxm, fym, mk, o + W - 1, o, g.use_mask, dc); acc -= dc; }   // second of pair (y,x-1)
  if (g.use_dt) {
    if (t + 1 < g.T) {
      const float d = (fxm[o] - fxm[o + 2 * hw]) + (fym[o] - fym[o + 2 * hw]);
      const float m = g.use_mask ? mk[o] * mk[o + hw] : 1.f;
      acc += m * d / sqrtf(d * d + 1e-6f);
    }
    if (t >= 1) {
      const float d = (fxm[o - 2 * hw] - fxm[o]) + (fym[o - 2 * hw] - fym[o]);
      const float m = g.use_mask ? mk[o - hw] * mk[o] : 1.f;
      acc -= m * d / sqrtf(d * d + 1e-6f);
    }
  }
  const float v = acc * coef * g_loss[0];
  g_fm[bt * 2 * hw + o] = v;
  g_fm[bt * 2 * hw + hw + o] = v;
}

// ---- forward: final scalar --------------------------------------------------------------------------------------
__global__ void iwe_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ smooth, int S, int B, int loss_scaling,
                                    float smooth_coef, float* __restrict__ loss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float total = 0.f;
  for (int s = 0; s < S; ++s) {
    float fw = 0.f, bw = 0.f;
    for (int b = 0; b < B; ++b) {
      const float* q = sums + ((size_t)(s * B + b) * 2) * 2;
      fw += loss_scaling ? q[0] / q[1] : q[0];
      bw += loss_scaling ? q[2] / q[3] : q[2];
    }
    total += fw + bw + smooth_coef * smooth[s];
  }
  loss[0] = total / (float)S;
}

// ---- backward: adjoint images -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iwe_adjoint_kernel(const float* __restrict__ img, const float* __restrict__ sums,
                                                          float* __restrict__ adj, int HW, float T, int loss_scaling, float inv_S,
                                                          const float* __restrict__ g_loss) {
  const int sbd = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= HW) return;
  const float* d = img + (size_t)sbd * 4 * HW;
  float* a = adj + (size_t)sbd * 4 * HW;
  const float ssq = sums[sbd * 2], n = sums[sbd * 2 + 1];
  const float g = g_loss[0] * inv_S / (loss_scaling ? n : 1.f);
  const float4 q = reinterpret_cast<const float4*>(d)[i];
  const float ip = q.x, tp = q.y, in = q.z, tn = q.w;
  const float ap = tp / (ip + 1e-9f) / T, an = tn / (in + 1e-9f) / T;
  const float gtp = g * 2.f * ap / ((ip + 1e-9f) * T), gtn = g * 2.f * an / ((in + 1e-9f) * T);
  float gip = -gtp * tp / (ip + 1e-9f), gin = -gtn * tn / (in + 1e-9f);
  if (loss_scaling && !(ip + in > 0.f)) {  // empty pixel keeps gradient 1 into the divisor (in-place masked assignment)
    const float gn = -g_loss[0] * inv_S * ssq / (n * n);
    gip += gn;
    gin += gn;
  }
  reinterpret_cast<float4*>(a)[i] = make_float4(gip, gtp, gin, gtn);
}

// d max(0, 1-|d|) / d d with torch's tie conventions: abs'(0) = 0; maximum splits the gradient at equality.
__device__ __forceinline__ float dweight(float d) {
  const float u = 1.0f - fabsf(d);
  const float s = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
  return u > 0.f ? s : (u == 0.f ? 0.5f * s : 0.f);
}

// ---- backward: per-event gradient, scattered to the flow maps ----------------------------------------------------
__global__ void __launch_bounds__(256) iwe_event_grad_kernel(const ef_iwe_loss_params p, const float* __restrict__ adj) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y, s = blockIdx.z;
  if (i >= p.n_total) return;
  const size_t hw = (size_t)p.H * p.W;
  const float4 e = reinterpret_cast<const float4*>(p.events)[(size_t)b * p.n_total + i];
  const float2 pm = reinterpret_cast<const float2*>(p.pol_mask)[(size_t)b * p.n_total + i];
  if (pm.x == 0.f && pm.y == 0.f) return;
  const int t_e = pass_of_event(p, i);
  const int pix = (int)(e.y * (float)p.W + e.z);
  const size_t mo = (((size_t)s * p.B + b) * p.T_maps + t_e) * 2 * hw;
  const float fx = __ldg(p.flow_maps + mo + pix), fy = __ldg(p.flow_maps + mo + hw + pix);
  const float* abase = adj + ((size_t)s * p.B + b) * 8 * hw;
  float gfy = 0.f, gfx = 0.f;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    const float tref = dir == 0 ? (float)p.T : 0.f;
    const float tau = dir == 0 ? e.x : ((float)p.T - e.x);
    const float kk = (tref - e.x) * p.flow_scaling;
    float yw, xw;
    Corner c[4];
    warp_event(e.x, e.y, e.z, fy, fx, tref, p.flow_scaling, p.H, p.W, yw, xw, c);
    const float* a = abase + (size_t)dir * 4 * hw;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c[k].idx < 0) continue;
      float delta = 0.f;
      const float2* q = reinterpret_cast<const float2*>(a + (size_t)c[k].idx * 4);  // adjoints of [I+, Th+], [I-, Th-]
      if (pm.x != 0.f) {
        const float2 g2 = __ldg(q);
        delta += pm.x * (g2.x + tau * g2.y);
      }
      if (pm.y != 0.f) {
        const float2 g2 = __ldg(q + 1);
        delta += pm.y * (g2.x + tau * g2.y);
      }
      gfy += delta * dweight(c[k].dy) * c[k].wx * kk;
      gfx += delta * c[k].wy * dweight(c[k].dx) * kk;
    }
  }
  if (gfx != 0.f) atomicAdd(p.g_flow_maps + mo + pix, gfx);
  if (gfy != 0.f) atomicAdd(p.g_flow_maps + mo + hw + pix, gfy);
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
// workspace: img [B][2 = warped, unwarped][HW][4 = I+,Th+,I-,Th-] (pixel-interleaved like the loss workspace), then sums
// [B][2][4 = sum A^2, n, sum I, sum I^2]
__global__ void __launch_bounds__(256) iwe_metric_scatter_kernel(const ef_iwe_metrics_params p, float* __restrict__ img) {
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
    // FWL scatters weight 1 per event regardless of polarity (no mask); RSAT scatters per polarity.  Channels 0/1 serve
    // both when every event has exactly one polarity bit; events with neither are still counted for FWL in channel 0.
    float2* q = reinterpret_cast<float2*>(d + (size_t)idx * 4);
    if (pm.x != 0.f) atomicAdd(q, make_float2(pm.x, __fmul_rn(e.x, pm.x)));
    if (pm.y != 0.f) atomicAdd(q + 1, make_float2(pm.y, __fmul_rn(e.x, pm.y)));
  }
}

__global__ void __launch_bounds__(256) iwe_metric_reduce_kernel(const float* __restrict__ img, float* __restrict__ sums, int HW, float T) {
  __shared__ float s_red[8];
  const int bk = blockIdx.y;  // b*2 + k
  const float* d = img + (size_t)bk * 4 * HW;
  float ssq = 0.f, n = 0.f, s1 = 0.f, s2 = 0.f;
  const int p0 = blockIdx.x * RED_PIX;
  for (int i = p0 + threadIdx.x; i < min(p0 + RED_PIX, HW); i += 256) {
    const float4 q = reinterpret_cast<const float4*>(d)[i];
    const float ip = q.x, tp = q.y, in = q.z, tn = q.w;
    const float ap = tp / (ip + 1e-9f) / T, an = tn / (in + 1e-9f) / T;
    ssq += ap * ap + an * an;
    n += (ip + in > 0.f) ? 1.f : 0.f;
    s1 += ip + in;
    s2 += (ip + in) * (ip + in);
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

static int validate_loss(const ef_iwe_loss_params& p, const char* who) {
  EF_REQUIRE(p.S > 0 && p.B > 0 && p.T > 0 && p.H > 0 && p.W > 0 && p.n_total >= 0, EF_EINVAL, "%s: bad dimensions", who);
  EF_REQUIRE(p.T_maps == (p.overwrite_intermediate ? 1 : p.T), EF_EINVAL, "%s: T_maps must be 1 with overwrite_intermediate, else T", who);
  EF_REQUIRE(p.T_maps == 1 || p.pass_offsets || (p.n_per_pass > 0 && (int64_t)p.n_per_pass * p.T >= p.n_total), EF_EINVAL,
             "%s: n_per_pass * T must cover n_total", who);
  EF_REQUIRE(p.events && p.pol_mask && p.flow_maps && p.workspace, EF_ENULL, "%s: NULL tensor", who);
  EF_REQUIRE(!p.smoothing_mask || p.event_mask, EF_ENULL, "%s: smoothing_mask without event_mask", who);
  return EF_OK;
}

}  // namespace ef

extern "C" int64_t ef_iwe_loss_workspace_elems(int32_t S, int32_t B, int32_t H, int32_t W) {
  return (int64_t)ef::ws_layout(S, B, H, W).total;
}

extern "C" int ef_iwe_loss_fwd(const ef_iwe_loss_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_loss_fwd: params is NULL");
  const ef_iwe_loss_params& p = *pp;
  if (int rc = validate_loss(p, "ef_iwe_loss_fwd")) return rc;
  EF_REQUIRE(p.loss, EF_ENULL, "ef_iwe_loss_fwd: loss is NULL");
  cudaStream_t st = as_stream(stream);
  const WsLayout l = ws_layout(p.S, p.B, p.H, p.W);
  const int HW = p.H * p.W;
  float* ws = p.workspace;
  cudaMemsetAsync(ws + l.img, 0, (l.adj - l.img) * sizeof(float), st);
  cudaMemsetAsync(ws + l.sums, 0, (l.total - l.sums) * sizeof(float), st);
  int rc;
  if (p.n_total > 0) {
    iwe_scatter_kernel<<<dim3(cdiv(p.n_total, 256), p.B, p.S), 256, 0, st>>>(p, ws + l.img);
    if ((rc = check_launch("iwe_scatter_kernel"))) return rc;
  }
  iwe_reduce_kernel<<<dim3(cdiv(HW, RED_PIX), p.S * p.B * 2), 256, 0, st>>>(ws + l.img, ws + l.sums, HW, (float)p.T);
  if ((rc = check_launch("iwe_reduce_kernel"))) return rc;
  const SmoothGeom g{p.B, p.T_maps, p.H, p.W, p.smoothing_mask != 0, !p.overwrite_intermediate};
  const size_t n = (size_t)p.B * p.T_maps * HW;
  const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  for (int s = 0; s < p.S; ++s) {
    smooth_fwd_kernel<<<blocks, 256, 0, st>>>(p.flow_maps + (size_t)s * p.B * p.T_maps * 2 * HW, p.smoothing_mask ? p.event_mask : nullptr, g,
                                              ws + l.smooth + s);
    if ((rc = check_launch("smooth_fwd_kernel"))) return rc;
  }
  const float components = p.overwrite_intermediate ? 4.f : 5.f;
  iwe_finalize_kernel<<<1, 32, 0, st>>>(ws + l.sums, ws + l.smooth, p.S, p.B, p.loss_scaling, p.weight / components / (float)p.T_maps, p.loss);
  return check_launch("iwe_finalize_kernel");
}

extern "C" int ef_iwe_loss_bwd(const ef_iwe_loss_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_iwe_loss_bwd: params is NULL");
  const ef_iwe_loss_params& p = *pp;
  if (int rc = validate_loss(p, "ef_iwe_loss_bwd")) return rc;
  EF_REQUIRE(p.g_loss && p.g_flow_maps, EF_ENULL, "ef_iwe_loss_bwd: g_loss / g_flow_maps is NULL");
  cudaStream_t st = as_stream(stream);
  const WsLayout l = ws_layout(p.S, p.B, p.H, p.W);
  const int HW = p.H * p.W;
  float* ws = p.workspace;
  int rc;
  const SmoothGeom g{p.B, p.T_maps, p.H, p.W, p.smoothing_mask != 0, !p.overwrite_intermediate};
  const size_t n = (size_t)p.B * p.T_maps * HW;
  const float components = p.overwrite_intermediate ? 4.f : 5.f;
  const float coef = p.weight / components / (float)p.T_maps / (float)p.S;
  for (int s = 0; s < p.S; ++s) {
    const size_t off = (size_t)s * p.B * p.T_maps * 2 * HW;
    smooth_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.flow_maps + off, p.smoothing_mask ? p.event_mask : nullptr, g, p.g_loss, coef,
                                                                   p.g_flow_maps + off);
    if ((rc = check_launch("smooth_bwd_kernel"))) return rc;
  }
  iwe_adjoint_kernel<<<dim3(cdiv(HW, 256), p.S * p.B * 2), 256, 0, st>>>(ws + l.img, ws + l.sums, ws + l.adj, HW, (float)p.T, p.loss_scaling,
                                                                         1.0f / (float)p.S, p.g_loss);
  if ((rc = check_launch("iwe_adjoint_kernel"))) return rc;
  if (p.n_total > 0) {
    iwe_event_grad_kernel<<<dim3(cdiv(p.n_total, 256), p.B, p.S), 256, 0, st>>>(p, ws + l.adj);
    if ((rc = check_launch("iwe_event_grad_kernel"))) return rc;
  }
  return EF_OK;
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

extern "C" int64_t ef_iwe_metrics_workspace_elems(int32_t B, int32_t H, int32_t W) { return (int64_t)B * 8 * H * W + (int64_t)B * 8; }

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
  float* sums = p.workspace + (size_t)p.B * 8 * HW;
  cudaMemsetAsync(p.workspace, 0, ((size_t)p.B * 8 * HW + (size_t)p.B * 8) * sizeof(float), st);
  int rc;
  if (p.n_total > 0) {
    iwe_metric_scatter_kernel<<<dim3(cdiv(p.n_total, 256), p.B), 256, 0, st>>>(p, img);
    if ((rc = check_launch("iwe_metric_scatter_kernel"))) return rc;
  }
  iwe_metric_reduce_kernel<<<dim3(cdiv(HW, RED_PIX), p.B * 2), 256, 0, st>>>(img, sums, HW, (float)p.T);
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
