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
constexpr int PWC_THREADS = 256, PWC_PPT = 2;

__global__ void __launch_bounds__(PWC_THREADS) lif_bwd_pointwise_cl_kernel(const ef_lif_bwd_tc_params p) {
  __shared__ float s_sum[64], lam[32], thr[32];
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t hw = (size_t)p.H * p.W;
  const int b = blockIdx.y;
  if (tid < 64) s_sum[tid] = 0.f;
  if (tid < 32) {
    lam[tid] = sigmoidf_acc(__ldg(p.leak + tid));
    thr[tid] = fmaxf(__ldg(p.thresh + tid), 0.01f);
  }
  __syncthreads();
  float s_lam[32], s_thr[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) s_lam[c] = s_thr[c] = 0.f;
  const bool hard = p.hard_reset != 0;
#pragma unroll 1
  for (int k = 0; k < PWC_PPT; ++k) {
    const size_t pix = ((size_t)blockIdx.x * PWC_PPT + k) * PWC_THREADS + tid;
    if (pix >= hw) break;
    uint4 zq[4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
      zq[g] = p.z_in_cl ? __ldg(reinterpret_cast<const uint4*>(p.z_in_cl + ((size_t)b * hw + pix) * 32 + g * 8)) : make_uint4(0, 0, 0, 0);
    const uint32_t zw[16] = {zq[0].x, zq[0].y, zq[0].z, zq[0].w, zq[1].x, zq[1].y, zq[1].z, zq[1].w,
                             zq[2].x, zq[2].y, zq[2].z, zq[2].w, zq[3].x, zq[3].y, zq[3].z, zq[3].w};
    uint32_t hi[16], mid[16];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const size_t o = ((size_t)b * 32 + c) * hw + pix;
      const float v_p = p.v_in ? __ldg(p.v_in + o) : 0.f;
      const float z_p = (c & 1) ? bf16_hi(zw[c >> 1]) : bf16_lo(zw[c >> 1]);
      const float v_n = __ldg(p.v_out + o);
      const float g_z = (p.g_out ? __ldg(p.g_out + o) : 0.f) + (p.g_z_out ? __ldg(p.g_z_out + o) : 0.f);
      const float sg = surrogate_grad(p.surrogate, v_n - thr[c], p.act_width);
      const float g_v = (p.g_v_out ? __ldg(p.g_v_out + o) : 0.f) + g_z * sg;
      const float oml = 1.0f - lam[c];
      const float g_I = oml * g_v;
      const float keep = hard ? v_p * (1.0f - z_p) : v_p;
      const float drive = hard ? (v_n - lam[c] * keep) / oml : (v_n - lam[c] * v_p + z_p * thr[c]) / oml;
      s_lam[c] += g_v * (keep - drive);
      s_thr[c] += -g_z * sg - (hard ? 0.f : z_p * g_v);
      if (p.g_v_in) p.g_v_in[o] = hard ? g_v * lam[c] * (1.0f - z_p) : g_v * lam[c];
      const __nv_bfloat16 h = __float2bfloat16_rn(g_I);
      const __nv_bfloat16 m = __float2bfloat16_rn(g_I - __bfloat162float(h));
      const uint32_t hb = *reinterpret_cast<const uint16_t*>(&h), mb = *reinterpret_cast<const uint16_t*>(&m);
      if (c & 1) hi[c >> 1] |= hb << 16, mid[c >> 1] |= mb << 16;
      else hi[c >> 1] = hb, mid[c >> 1] = mb;
    }
    uint4* dh = reinterpret_cast<uint4*>(p.gI_hi + ((size_t)b * hw + pix) * 32);
    uint4* dm = reinterpret_cast<uint4*>(p.gI_mid + ((size_t)b * hw + pix) * 32);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      dh[g] = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
      dm[g] = make_uint4(mid[4 * g], mid[4 * g + 1], mid[4 * g + 2], mid[4 * g + 3]);
    }
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const float a = warp_sum(s_lam[c]), t = warp_sum(s_thr[c]);
    if (lane == 0) {
      atomicAdd(&s_sum[c], a);
      atomicAdd(&s_sum[32 + c], t);
    }
  }
  __syncthreads();
  if (tid < 32) {
    const float l = sigmoidf_acc(p.leak[tid]);
    if (p.g_leak) atomicAdd(p.g_leak + tid, s_sum[tid] * l * (1.0f - l));
    if (p.g_thresh && p.thresh[tid] >= 0.01f) atomicAdd(p.g_thresh + tid, s_sum[32 + tid]);
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

}  // namespace ef

extern "C" int64_t ef_split_weights_bwd_elems(int32_t has_rec) { return (int64_t)ef::dg_layout(has_rec != 0).w_bytes / 2; }

extern "C" int ef_split_weights_bwd(const float* w_ff, const float* w_rec, uint16_t* out, void* stream) {
  using namespace ef;
  EF_REQUIRE(w_ff && out, EF_ENULL, "ef_split_weights_bwd: NULL tensor");
  const int nconv = w_rec ? 2 : 1;
  split_weights_bwd_kernel<<<cdiv(nconv * 9216, 256), 256, 0, as_stream(stream)>>>(w_ff, w_rec, out);
  return check_launch("split_weights_bwd_kernel");
}

extern "C" int ef_lif_bwd_tc(const ef_lif_bwd_tc_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_lif_bwd_tc: params is NULL");
  const ef_lif_bwd_tc_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_lif_bwd_tc: bad dimensions");
  EF_REQUIRE(p.x_cl && p.v_out && p.leak && p.thresh && p.w_bwd && p.gI_hi && p.gI_mid && p.g_x, EF_ENULL, "ef_lif_bwd_tc: NULL tensor");
  EF_REQUIRE(!p.has_rec || !p.z_in_cl || p.g_z_in || true, EF_ENULL, "ef_lif_bwd_tc");
  cudaStream_t st = as_stream(stream);
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int rc;
  const int hw = p.H * p.W;
  lif_bwd_pointwise_cl_kernel<<<dim3(cdiv(hw, PWC_THREADS * PWC_PPT), p.B), PWC_THREADS, 0, st>>>(p);
  if ((rc = check_launch("lif_bwd_pointwise_cl_kernel"))) return rc;

  const bool rec = p.has_rec != 0;
  DgParams q;
  q.B = p.B, q.H = p.H, q.W = p.W, q.has_rec = rec;
  q.tiles_x = cdiv(p.W, DG_TW), q.tiles_y = cdiv(p.H, DG_TH), q.n_tiles = p.B * q.tiles_x * q.tiles_y;
  q.w_bwd = p.w_bwd, q.g_x = p.g_x, q.g_z_in = (rec && p.z_in_cl) ? p.g_z_in : nullptr;
  CUtensorMap mh, mm;
  if ((rc = get_map(p.gI_hi, p.B, p.H, p.W, DG_TH + 2, DG_TW + 8, true, &mh))) return rc;
  if ((rc = get_map(p.gI_mid, p.B, p.H, p.W, DG_TH + 2, DG_TW + 8, true, &mm))) return rc;
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
  if ((rc = check_launch("lif_dgrad_tc_kernel"))) return rc;

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
