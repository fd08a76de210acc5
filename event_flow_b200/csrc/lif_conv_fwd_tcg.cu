// Fused conv3x3 + LIF step on tcgen05 for GENERAL channel counts (multiples of 32) and several concatenated input sources, sm_100a only.
// This is the kernel of the EV-FlowNet family (SpikingRecEVFlowNet, models/unet.py:418-465): recurrent encoder cells 64..512 channels
// (the recurrent convolution is one more input source), residual blocks 512 -> 512 (second cell adds the block input to its spikes,
// spiking_submodules.py:933-975), decoders whose input is cat[prediction, x, skip] (unet.py:451-462) -- the concat is never
// materialised: every source keeps its own tensor and contributes its own K blocks.
//
// Same implicit GEMM as lif_conv_fwd_tc.cu (16x8-pixel tile, halo tile landed by ONE TMA box per 32-channel block, taps = descriptor
// start addresses, weights = three exact bf16 terms stacked along N = 96, fp32 accumulator in tensor memory), blocked over channels:
//   work item = (pixel tile, block of 32 output channels); its K loop runs over the 32-channel blocks of all sources.  Per K step the
//   producer thread lands one halo tile (11.5 KB) in the operand ring and one weight block [9 taps][96][32] (54 KB, streamed from the
//   L2-resident weight image) in the weight ring; the MMA thread issues 18 MMAs (N = 96, K = 16) into the item's accumulator.
//   The epilogue is the LIF update of the 32 output channels on the membrane tile (TMA in, in-place update, TMA out), exactly as in the
//   32 -> 32 kernel, plus the optional residual output.
// Inputs must be exactly representable in bf16: spikes {0,1}, residual sums {0,1,2}, their bilinear x2 upsampling (multiples of 1/16),
// event counts; fractional fp32 sources (the upsampled flow prediction in the decoders) come as exact hi/mid/lo splits
// (ef_pack_split_cl) with the weight rows repeated per slot.  Every product is then exact in fp32; only the summation order differs
// from the CPU path.
#include "tc_common.cuh"

namespace ef {

constexpr int G_TH = 16, G_TW = 8;
constexpr int G_HALO_W = G_TW + 2, G_HALO_H = G_TH + 2;
constexpr int G_HALO_PITCH = G_HALO_W * PIX_BYTES;        // 640 B between tile rows of the operand tile
constexpr int G_HALO_BYTES = G_HALO_H * G_HALO_PITCH;     // 11520 B landed by TMA
constexpr int G_HALO_STAGE = 12288;
constexpr int G_V_TILE = 32 * 128 * 4;                    // [32 ch][16][8] fp32
constexpr int G_Z_TILE = 128 * PIX_BYTES;                 // [16][8][32 ch] bf16, 64B-swizzled
constexpr int G_WTAP = 96 * PIX_BYTES;                    // one tap of a weight block: [3 x 32 rows][32 k] bf16
constexpr int G_WBLOCK = 9 * G_WTAP;                      // 55296 B
constexpr int G_NOP = 3, G_NW = 2, G_NV = 2;
constexpr int G_V_STAGE = G_V_TILE + 2 * G_Z_TILE;        // membrane tile + previous spikes (-> new spikes) + residual (-> output)
constexpr int G_W_OFF = 0;
constexpr int G_OP_OFF = G_W_OFF + G_NW * G_WBLOCK;
constexpr int G_V_OFF = G_OP_OFF + G_NOP * G_HALO_STAGE;
constexpr int G_BAR_OFF = G_V_OFF + G_NV * G_V_STAGE;
constexpr int G_SMEM = G_BAR_OFF + 256 + 1024;
static_assert(G_SMEM <= 227 * 1024, "shared memory budget");
constexpr int G_ACC_COLS = 96, G_TMEM_COLS = 256;
constexpr uint32_t G_IDESC = umma_idesc(G_ACC_COLS, false, false, false);
constexpr int G_EPI_WARPS = 16, G_CPT = 8, G_THREADS = 128 + 32 * G_EPI_WARPS;

struct TcgParams {
  int B, H, W, C, tiles_x, tiles_y, n_tiles, nnb, nkb, n_items;
  int has_v, has_z, has_res;
  int n_src, src_blocks[EF_TCG_MAX_SRC];
  int tap_mask;  // bit t set: tap t (= ky*3 + kx) has non-zero weights.  Stride-2 cells on space-to-depth inputs use 4 of the 9 taps
  const uint16_t* w_image;
  const float* leak;
  const float* thresh;
};

__device__ __forceinline__ uint32_t g_sw64(uint32_t row, int c) { return row + ((uint32_t)(c ^ ((row >> 7) & 3)) << 4); }
__device__ __forceinline__ float g_lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void g_sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint4 g_lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void g_sts_u4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <bool HARD>
__global__ void __launch_bounds__(G_THREADS, 1)
lif_conv_fwd_tcg_kernel(const TcgParams p, const __grid_constant__ CUtensorMap m_s0, const __grid_constant__ CUtensorMap m_s1,
                        const __grid_constant__ CUtensorMap m_s2, const __grid_constant__ CUtensorMap m_s3,
                        const __grid_constant__ CUtensorMap map_vin, const __grid_constant__ CUtensorMap map_vout,
                        const __grid_constant__ CUtensorMap map_zc, const __grid_constant__ CUtensorMap map_zout,
                        const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t s_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (s_base - smem_u32(smem_raw));
  // barriers: operand full[NOP] empty[NOP] | weight full[NW] empty[NW] | membrane full[NV] empty[NV] | acc full[2] empty[2] | TMEM word
  const uint32_t bar0 = s_base + G_BAR_OFF;
  auto bar_opf = [&](int s) { return bar0 + 8u * s; };
  auto bar_ope = [&](int s) { return bar0 + 8u * (G_NOP + s); };
  auto bar_wf = [&](int s) { return bar0 + 8u * (2 * G_NOP + s); };
  auto bar_we = [&](int s) { return bar0 + 8u * (2 * G_NOP + G_NW + s); };
  auto bar_vf = [&](int s) { return bar0 + 8u * (2 * G_NOP + 2 * G_NW + s); };
  auto bar_ve = [&](int s) { return bar0 + 8u * (2 * G_NOP + 2 * G_NW + G_NV + s); };
  auto bar_accf = [&](int a) { return bar0 + 8u * (2 * G_NOP + 2 * G_NW + 2 * G_NV + a); };
  auto bar_acce = [&](int a) { return bar0 + 8u * (2 * G_NOP + 2 * G_NW + 2 * G_NV + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G_BAR_OFF + 8 * (2 * G_NOP + 2 * G_NW + 2 * G_NV + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // item -> (output-channel block, image, tile origin): channel blocks of one tile are neighbours in the item order (their operand
  // tiles are shared through L2)
  auto item_origin = [&](int it, int& nb, int& b, int& y0, int& x0) {
    const int g = blockIdx.x + it * gridDim.x;
    nb = g % p.nnb;
    const int tile = g / p.nnb;
    b = tile / tiles_per_img;
    const int r = tile - b * tiles_per_img, ty = r / p.tiles_x;
    y0 = ty * G_TH, x0 = (r - ty * p.tiles_x) * G_TW;
  };
  const uint32_t v_tx = (p.has_v ? G_V_TILE : 0) + (p.has_z ? G_Z_TILE : 0) + (p.has_res ? G_Z_TILE : 0);

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    prefetch_tensormap(&m_s0);
    for (int s = 0; s < G_NOP; ++s) {
      mbar_init(bar_opf(s), 1);
      mbar_init(bar_ope(s), 1);
    }
    for (int s = 0; s < G_NW; ++s) {
      mbar_init(bar_wf(s), 1);
      mbar_init(bar_we(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);
      mbar_init(bar_acce(a), G_EPI_WARPS);
    }
    fence_barrier_init();
  } else if (threadIdx.x == 64) {
    if (p.has_v) prefetch_tensormap(&map_vin);
    for (int s = 0; s < G_NV; ++s) {
      mbar_init(bar_vf(s), 1);
      mbar_init(bar_ve(s), 1);
    }
    fence_barrier_init();
  } else if (threadIdx.x == 128) {
    prefetch_tensormap(&map_vout);
    prefetch_tensormap(&map_zout);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(G_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== producer: operand halo tiles + weight blocks, one pair per K step ===============================
    if (lane == 0) {
      const CUtensorMap* maps[EF_TCG_MAX_SRC] = {&m_s0, &m_s1, &m_s2, &m_s3};
      pdl_wait();
      int ksg = 0;
      for (int it = 0; it < n_my; ++it) {
        int nb, b, y0, x0;
        item_origin(it, nb, b, y0, x0);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w_image) + (size_t)nb * p.nkb * G_WBLOCK;
        for (int s = 0; s < p.n_src; ++s) {
          for (int j = 0; j < p.src_blocks[s]; ++j, ++ksg) {
            const int so = ksg % G_NOP, sw = ksg % G_NW;
            mbar_wait(bar_ope(so), ((ksg / G_NOP) & 1) ^ 1);
            mbar_expect_tx(bar_opf(so), G_HALO_BYTES);
            tma_load_4d(s_base + G_OP_OFF + so * G_HALO_STAGE, maps[s], bar_opf(so), 32 * j, x0 - 1, y0 - 1, b);
            mbar_wait(bar_we(sw), ((ksg / G_NW) & 1) ^ 1);
            mbar_expect_tx(bar_wf(sw), G_WBLOCK);
#pragma unroll
            for (int c = 0; c < 6; ++c)
              bulk_load_1d(s_base + G_W_OFF + sw * G_WBLOCK + c * (G_WBLOCK / 6), wsrc + c * (G_WBLOCK / 6), G_WBLOCK / 6, bar_wf(sw));
            wsrc += G_WBLOCK;
          }
        }
      }
    }
  } else if (warp == 2) {
    // =============================== membrane / previous-spike / residual tile producer ===============================
    if (lane == 0) {
      pdl_wait();
      for (int it = 0; it < n_my; ++it) {
        int nb, b, y0, x0;
        item_origin(it, nb, b, y0, x0);
        const int s = it % G_NV;
        mbar_wait(bar_ve(s), ((it / G_NV) & 1) ^ 1);
        const uint32_t st = s_base + G_V_OFF + s * G_V_STAGE;
        if (v_tx == 0) {
          mbar_arrive(bar_vf(s));
          continue;
        }
        mbar_expect_tx(bar_vf(s), v_tx);
        if (p.has_v) tma_load_4d(st, &map_vin, bar_vf(s), x0, y0, 32 * nb, b);
        if (p.has_z) tma_load_4d(st + G_V_TILE, &map_zc, bar_vf(s), 32 * nb, x0, y0, b);
        if (p.has_res) tma_load_4d(st + G_V_TILE + G_Z_TILE, &map_res, bar_vf(s), 32 * nb, x0, y0, b);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      int ksg = 0;
      for (int it = 0; it < n_my; ++it) {
        const int a = it & 1;
        mbar_wait(bar_acce(a), ((it >> 1) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + a * G_ACC_COLS;
        uint32_t acc_started = 0u;
        for (int kb = 0; kb < p.nkb; ++kb, ++ksg) {
          const int so = ksg % G_NOP, sw = ksg % G_NW;
          mbar_wait(bar_opf(so), (ksg / G_NOP) & 1);
          mbar_wait(bar_wf(sw), (ksg / G_NW) & 1);
          tc_fence_after();
          const uint64_t ax = umma_desc_sw64(s_base + G_OP_OFF + so * G_HALO_STAGE, G_HALO_PITCH);
          const uint64_t bw = umma_desc_sw64(s_base + G_W_OFF + sw * G_WBLOCK, ATOM_BYTES);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (!((p.tap_mask >> tap) & 1)) continue;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              umma_bf16<G_IDESC>(d_tmem, ax + (uint64_t)((tap / 3) * (G_HALO_PITCH / 16) + (tap % 3) * (PIX_BYTES / 16) + ks * 2),
                                 bw + (uint64_t)(tap * (G_WTAP / 16) + ks * 2), acc_started);
              acc_started = 1u;
            }
          }
          umma_commit(bar_ope(so));  // operand tile and weight block may be overwritten once these MMAs have read them
          umma_commit(bar_we(sw));
        }
        umma_commit(bar_accf(a));
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue: 4 lane quadrants x 4 channel groups of 8 ===============================
    const int e = warp - 4;
    const int q = e & 3;
    const int c0 = G_CPT * (e >> 2);
    const int m = q * 32 + lane;
    const bool store_thread = (threadIdx.x == 128);
    if (store_thread) pdl_wait();
    for (int it = 0; it < n_my; ++it) {
      int nb, b, y0, x0;
      item_origin(it, nb, b, y0, x0);
      float lam[G_CPT], thr[G_CPT];
#pragma unroll
      for (int j = 0; j < G_CPT; ++j) {
        lam[j] = sigmoidf_acc(__ldg(p.leak + 32 * nb + c0 + j));
        thr[j] = fmaxf(__ldg(p.thresh + 32 * nb + c0 + j), 0.01f);
      }
      const int sv = it % G_NV, a = it & 1;
      const uint32_t vst = s_base + G_V_OFF + sv * G_V_STAGE;
      const uint32_t v_addr = vst + (uint32_t)(c0 * 512 + m * 4);
      const uint32_t z_row = vst + G_V_TILE + (uint32_t)(m * PIX_BYTES), r_row = z_row + G_Z_TILE;
      float vc[G_CPT];
      mbar_wait(bar_vf(sv), (it / G_NV) & 1);
#pragma unroll
      for (int j = 0; j < G_CPT; ++j) vc[j] = p.has_v ? g_lds_f32(v_addr + j * 512) : 0.f;
      const uint4 zc = p.has_z ? g_lds_u4(g_sw64(z_row, c0 / 8)) : make_uint4(0, 0, 0, 0);
      const uint4 rc = p.has_res ? g_lds_u4(g_sw64(r_row, c0 / 8)) : make_uint4(0, 0, 0, 0);
      mbar_wait(bar_accf(a), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + a * G_ACC_COLS + c0 + ((uint32_t)(q * 32) << 16);
      uint32_t a_hi[8], a_mid[8], a_lo[8];
      tmem_ld8(tacc, a_hi);
      tmem_ld8(tacc + 32, a_mid);
      tmem_ld8(tacc + 64, a_lo);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce(a));
      const uint32_t zw[4] = {zc.x, zc.y, zc.z, zc.w}, rw[4] = {rc.x, rc.y, rc.z, rc.w};
      float vn[G_CPT];
      uint32_t zpk[G_CPT / 2], opk[G_CPT / 2];
#pragma unroll
      for (int j = 0; j < G_CPT; ++j) {
        const float I = __fadd_rn(__fadd_rn(__uint_as_float(a_lo[j]), __uint_as_float(a_mid[j])), __uint_as_float(a_hi[j]));
        const float z = (j & 1) ? bf16_hi(zw[j >> 1]) : bf16_lo(zw[j >> 1]);
        if (HARD) vn[j] = __fadd_rn(__fmul_rn(__fmul_rn(vc[j], lam[j]), __fsub_rn(1.0f, z)), __fmul_rn(__fsub_rn(1.0f, lam[j]), I));
        else vn[j] = __fsub_rn(__fadd_rn(__fmul_rn(vc[j], lam[j]), __fmul_rn(__fsub_rn(1.0f, lam[j]), I)), __fmul_rn(z, thr[j]));
        const bool fire = __fsub_rn(vn[j], thr[j]) > 0.f;
        const uint32_t zb = fire ? 0x3F80u : 0u;  // bf16(1.0)
        const float res = (j & 1) ? bf16_hi(rw[j >> 1]) : bf16_lo(rw[j >> 1]);
        const uint32_t ob = pack_bf16x2(__fadd_rn(fire ? 1.0f : 0.f, res), 0.f) & 0xffffu;  // spikes + residual: small integers, exact
        if (j & 1) zpk[j >> 1] |= zb << 16, opk[j >> 1] |= ob << 16;
        else zpk[j >> 1] = zb, opk[j >> 1] = ob;
      }
      // the TMA stores of the previous item must have read their shared-memory source before the stage goes back to its producer
      if (store_thread && it > 0) {
        bulk_wait_read0();
        mbar_arrive(bar_ve((it - 1) % G_NV));
      }
#pragma unroll
      for (int j = 0; j < G_CPT; ++j) g_sts_f32(v_addr + j * 512, vn[j]);
      g_sts_u4(g_sw64(z_row, c0 / 8), make_uint4(zpk[0], zpk[1], zpk[2], zpk[3]));
      if (p.has_res) g_sts_u4(g_sw64(r_row, c0 / 8), make_uint4(opk[0], opk[1], opk[2], opk[3]));
      fence_proxy_async();
      named_bar_sync(2, 32 * G_EPI_WARPS);
      if (store_thread) {
        tma_store_4d(&map_vout, vst, x0, y0, 32 * nb, b);
        tma_store_4d(&map_zout, vst + G_V_TILE, 32 * nb, x0, y0, b);
        if (p.has_res) tma_store_4d(&map_out, vst + G_V_TILE + G_Z_TILE, 32 * nb, x0, y0, b);
        bulk_commit();
      }
    }
    if (store_thread) {
      bulk_wait0();
      if (n_my > 0) mbar_arrive(bar_ve((n_my - 1) % G_NV));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G_TMEM_COLS) : "memory");
  }
}

// ---- weight image: [C/32 output blocks][K blocks][9 taps][96 = {hi, mid, lo} x 32 n][32 k] bf16, 64B-swizzled per tap (as ef_split_weights) -----
struct WSrc {
  const float* w;   // [C][c_total][3][3]
  int c_total, ch0, n, split, s2d, blocks;
};
struct WSrcs {
  WSrc s[EF_TCG_MAX_SRC];
  int n_src, nkb, C;
};

__global__ void __launch_bounds__(256) split_weights_g_kernel(const WSrcs ws, uint16_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // over C * nkb * 32 (k) * 9 (tap)
  const long long total = (long long)ws.C * ws.nkb * 32 * 9;
  if (i >= total) return;
  const int tap = (int)(i % 9), k = (int)((i / 9) % 32), kb = (int)((i / (9 * 32)) % ws.nkb), co = (int)(i / ((long long)9 * 32 * ws.nkb));
  int s = 0, j = kb;
  while (s + 1 < ws.n_src && j >= ws.s[s].blocks) j -= ws.s[s].blocks, ++s;
  const WSrc& src = ws.s[s];
  float w = 0.f;
  // virtual channel of the source tensor this K row holds (-1: padding)
  const int nv = src.s2d ? 4 * src.n : src.n;  // space-to-depth sources carry the four pixel parities of every channel
  int vc = -1;
  if (src.split) {  // exact hi/mid/lo slots of a fractional source: every slot multiplies the same weight
    const int SL = EF_HEAD_SLOT(nv), c = k % SL, slot = k / SL;
    if (slot < 3 && c < nv) vc = c;
  } else {
    const int c = j * 32 + k;
    if (c < nv) vc = c;
  }
  if (vc >= 0) {
    if (!src.s2d) {
      w = src.w[((size_t)co * src.c_total + src.ch0 + vc) * 9 + tap];
    } else {
      // stride-2 convolution as a stride-1 convolution over the space-to-depth input: s2d[Y, X, (py*2 + px)*n + c] = in[2Y + py, 2X + px, c].
      // Output (y, x) reads input row 2y + dy - 1: dy = 0 -> (Y = y-1, py = 1), dy = 1 -> (Y = y, py = 0), dy = 2 -> (Y = y, py = 1); same in x.
      // In the 3x3 stride-1 kernel over s2d, tap ky = 0 is Y = y-1 and ky = 1 is Y = y; ky = 2 is never used (tap mask 0b000011011).
      const int par = vc / src.n, c = vc - par * src.n, py = par >> 1, px = par & 1, ky = tap / 3, kx = tap % 3;
      const int dy = ky == 0 ? (py == 1 ? 0 : -1) : (ky == 1 ? (py == 0 ? 1 : 2) : -1);
      const int dx = kx == 0 ? (px == 1 ? 0 : -1) : (kx == 1 ? (px == 0 ? 1 : 2) : -1);
      if (dy >= 0 && dx >= 0) w = src.w[((size_t)co * src.c_total + src.ch0 + c) * 9 + dy * 3 + dx];
    }
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const float r1 = w - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  const uint16_t parts[3] = {__bfloat16_as_ushort(hi), __bfloat16_as_ushort(mid), __bfloat16_as_ushort(lo)};
  const int nb = co >> 5, n = co & 31;
  uint16_t* blk = out + ((size_t)nb * ws.nkb + kb) * (G_WBLOCK / 2) + (size_t)tap * (G_WTAP / 2);
#pragma unroll
  for (int sp = 0; sp < 3; ++sp) {
    const int nn = sp * 32 + n, r = nn & 7;
    const int chunk = (k >> 3) ^ ((r >> 1) & 3);
    blk[(nn >> 3) * 256 + r * 32 + chunk * 8 + (k & 7)] = parts[sp];
  }
}

static int src_blocks_of(int32_t n, int32_t split, int32_t s2d) { return split ? 1 : ((s2d ? 4 * n : n) + 31) / 32; }

}  // namespace ef

extern "C" int64_t ef_split_weights_g_elems(int32_t C, int32_t n_src, const ef_wsrc* srcs) {
  if (C <= 0 || C % 32 || n_src < 1 || n_src > EF_TCG_MAX_SRC || !srcs) return 0;
  int nkb = 0;
  for (int i = 0; i < n_src; ++i) nkb += ef::src_blocks_of(srcs[i].n, srcs[i].split, srcs[i].s2d);
  return (int64_t)(C / 32) * nkb * (ef::G_WBLOCK / 2);
}

extern "C" int ef_split_weights_g(const ef_wsrc* srcs, int32_t n_src, int32_t C, uint16_t* out, void* stream) {
  using namespace ef;
  EF_REQUIRE(srcs && out, EF_ENULL, "ef_split_weights_g: NULL argument");
  EF_REQUIRE(C > 0 && C % 32 == 0 && n_src >= 1 && n_src <= EF_TCG_MAX_SRC, EF_EINVAL, "ef_split_weights_g: C must be a multiple of 32, 1..%d sources",
             EF_TCG_MAX_SRC);
  WSrcs ws;
  ws.n_src = n_src, ws.C = C, ws.nkb = 0;
  for (int i = 0; i < n_src; ++i) {
    EF_REQUIRE(srcs[i].w && srcs[i].n > 0 && srcs[i].ch0 >= 0 && srcs[i].ch0 + srcs[i].n <= srcs[i].c_total, EF_EINVAL, "ef_split_weights_g: bad source %d", i);
    EF_REQUIRE(!srcs[i].split || (srcs[i].s2d ? 4 : 1) * srcs[i].n <= EF_HEAD_MAX_CIN, EF_EUNSUPPORTED,
               "ef_split_weights_g: a split source has at most %d (virtual) channels", EF_HEAD_MAX_CIN);
    ws.s[i].w = srcs[i].w, ws.s[i].c_total = srcs[i].c_total, ws.s[i].ch0 = srcs[i].ch0, ws.s[i].n = srcs[i].n, ws.s[i].split = srcs[i].split;
    ws.s[i].s2d = srcs[i].s2d;
    ws.s[i].blocks = src_blocks_of(srcs[i].n, srcs[i].split, srcs[i].s2d);
    ws.nkb += ws.s[i].blocks;
  }
  const long long total = (long long)C * ws.nkb * 32 * 9;
  split_weights_g_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(ws, out);
  return check_launch("split_weights_g_kernel");
}

extern "C" int ef_lif_conv_fwd_g(const ef_lif_conv_g_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_lif_conv_fwd_g: params is NULL");
  const ef_lif_conv_g_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0 && p.C > 0 && p.C % 32 == 0 && p.W % 4 == 0, EF_EINVAL,
             "ef_lif_conv_fwd_g: C must be a multiple of 32 and W a multiple of 4");
  EF_REQUIRE(p.n_src >= 1 && p.n_src <= EF_TCG_MAX_SRC, EF_EINVAL, "ef_lif_conv_fwd_g: 1..%d input sources", EF_TCG_MAX_SRC);
  EF_REQUIRE(p.w_image && p.leak && p.thresh && p.v_out && p.z_out_cl, EF_ENULL, "ef_lif_conv_fwd_g: NULL tensor");
  EF_REQUIRE(!p.v_in == !p.z_in_cl, EF_EINVAL, "ef_lif_conv_fwd_g: v_in and z_in_cl come together");
  EF_REQUIRE(!p.residual_cl == !p.out_cl, EF_EINVAL, "ef_lif_conv_fwd_g: residual_cl and out_cl come together");
  EF_REQUIRE(p.v_in != p.v_out, EF_EINVAL, "ef_lif_conv_fwd_g: v_out must not alias v_in");
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  TcgParams q;
  q.B = p.B, q.H = p.H, q.W = p.W, q.C = p.C;
  q.tiles_x = cdiv(p.W, G_TW), q.tiles_y = cdiv(p.H, G_TH), q.n_tiles = p.B * q.tiles_x * q.tiles_y;
  q.nnb = p.C / 32, q.n_items = q.n_tiles * q.nnb;
  q.has_v = p.v_in != nullptr, q.has_z = p.z_in_cl != nullptr, q.has_res = p.residual_cl != nullptr;
  q.n_src = p.n_src, q.nkb = 0;
  CUtensorMap ms[EF_TCG_MAX_SRC], mv_in, mv_out, mzc, mzo, mres, mout;
  int rc;
  for (int i = 0; i < EF_TCG_MAX_SRC; ++i) {
    q.src_blocks[i] = 0;
    if (i >= p.n_src) {
      ms[i] = ms[0];
      continue;
    }
    EF_REQUIRE(p.src[i] && p.src_c[i] > 0 && p.src_c[i] % 32 == 0 && ((uintptr_t)p.src[i] % 16) == 0, EF_EINVAL,
               "ef_lif_conv_fwd_g: source %d must be a 16-byte aligned cl tensor with a multiple of 32 channels", i);
    q.src_blocks[i] = p.src_c[i] / 32;
    q.nkb += q.src_blocks[i];
    if ((rc = get_map_c(p.src[i], p.B, p.H, p.W, p.src_c[i], G_HALO_H, G_HALO_W, true, &ms[i]))) return rc;
  }
  if ((rc = get_map_vc(p.v_out, p.B, p.H, p.W, p.C, G_TH, G_TW, &mv_out))) return rc;
  if ((rc = get_map_c(p.z_out_cl, p.B, p.H, p.W, p.C, G_TH, G_TW, true, &mzo))) return rc;
  mv_in = mv_out, mzc = mzo, mres = mzo, mout = mzo;
  if (q.has_v && (rc = get_map_vc(p.v_in, p.B, p.H, p.W, p.C, G_TH, G_TW, &mv_in))) return rc;
  if (q.has_z && (rc = get_map_c(p.z_in_cl, p.B, p.H, p.W, p.C, G_TH, G_TW, true, &mzc))) return rc;
  if (q.has_res) {
    if ((rc = get_map_c(p.residual_cl, p.B, p.H, p.W, p.C, G_TH, G_TW, true, &mres))) return rc;
    if ((rc = get_map_c(p.out_cl, p.B, p.H, p.W, p.C, G_TH, G_TW, true, &mout))) return rc;
  }
  q.w_image = p.w_image, q.leak = p.leak, q.thresh = p.thresh;
  q.tap_mask = p.s2d ? 0x1B : 0x1FF;  // stride-2 cells on space-to-depth inputs: taps (ky, kx) in {0,1}^2 only
  cudaStream_t st = as_stream(stream);
  const int grid = q.n_items < n_sms ? q.n_items : n_sms;
  static bool attr_set[2] = {false, false};
  const int h = p.hard_reset ? 1 : 0;
  if (!attr_set[h]) {
    const cudaError_t e = h ? cudaFuncSetAttribute(lif_conv_fwd_tcg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM)
                            : cudaFuncSetAttribute(lif_conv_fwd_tcg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    if (e != cudaSuccess) return check_launch("cudaFuncSetAttribute(lif_conv_fwd_tcg_kernel)");
    attr_set[h] = true;
  }
  if (h) launch_pdl(lif_conv_fwd_tcg_kernel<true>, dim3(grid), dim3(G_THREADS), G_SMEM, st, q, ms[0], ms[1], ms[2], ms[3], mv_in, mv_out, mzc, mzo, mres, mout);
  else launch_pdl(lif_conv_fwd_tcg_kernel<false>, dim3(grid), dim3(G_THREADS), G_SMEM, st, q, ms[0], ms[1], ms[2], ms[3], mv_in, mv_out, mzc, mzo, mres, mout);
  return check_launch("lif_conv_fwd_tcg_kernel");
}
