// Fused conv3x3 + LIF step on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
//
// Implicit GEMM per 16x8-pixel tile:  D[128 px, 96] = sum over 9 taps, 32 (or 64 with the recurrent conv) input channels
//   A = bf16 spikes, channels-last.  ONE TMA box per tile (18 rows x 16 px x 32 ch: the 16x8 tile, its 1-pixel halo, and
//       padding up to whole 8-pixel atoms; hardware zero-fill = the conv padding) lands in shared memory in the canonical
//       64-byte-swizzled K-major UMMA layout: one pixel = one 64-byte row, 8 pixels = one 512-byte swizzle atom, one tile
//       row = 1024 bytes.  A tap (dy, dx) is then only a descriptor start address, tile + dy*1024 + dx*64, with
//       SBO = 1024: no im2col copy.  The start is NOT atom-aligned for dx != 0; this works because the swizzle XOR is a
//       function of the absolute shared-memory address bits [7:8] on both the TMA write and the UMMA read (verified
//       bit-exactly against the CUDA-core kernel, tests/test_gpu_tc.py);
//   B = weights, split into three bf16 terms hi+mid+lo == w (exact), resident in shared memory for the whole kernel and
//       stacked along N: one MMA per (tap, k-step) with N = 96 = {hi, mid, lo} x 32 channels, so the A tile is read from
//       shared memory once instead of three times; same swizzled K-major layout (written by ef_split_weights);
//   D = fp32 accumulator in tensor memory, 96 columns (three partial sums, added in the epilogue), double buffered.
// History (profiles/r01_tc_kernel_notes.md): an un-swizzled halo tile made every MMA ~3.6x slower; three x-shifted
// swizzled copies tripled the L2->SM traffic; the single padded swizzled tile is both the least traffic and full MMA rate.
// Spikes are {0,1,2}: exactly representable in bf16, so every product is exact and only the fp32 summation order differs
// from the CPU path (no TF32/BF16 rounding may enter a spiking conv).
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner (warps 2-3 idle: setmaxnreg moves the
// registers of this warpgroup to the others), warps 4-11 = epilogue: each warp owns a
// TMEM lane quadrant (32 pixels) and one half of the channels; membrane potentials, previous spikes and new spikes go
// straight global <-> registers (the next tile's loads are in flight while the current tile is computed; nothing in the
// epilogue touches shared memory, so no proxy fence or CTA barrier sits on the per-tile path).  Persistent over tiles;
// mbarrier pipelines between the roles.
// Reference semantics: models/spiking_submodules.py:96-126 (ConvLIF), :516-551 (ConvLIFRecurrent).
#include "tc_common.cuh"

namespace ef {

// Output tile = 128 pixels = 16 rows x 8 cols: one 8-pixel atom per tile row is the only shape whose tap-shifted windows
// are expressible as ONE descriptor (constant stride between consecutive 8-row groups).
constexpr int W_BLOCK_BYTES = 96 * PIX_BYTES;         // one tap: [96 n = 3 splits x 32 ch][32 k] bf16, 64B-swizzled: 6144 B
constexpr int W_CONV_BYTES = 9 * W_BLOCK_BYTES;       // 55296 B per convolution
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 32 * (4 + TC_EPI_WARPS);   // warpgroup 0: TMA warp, MMA warp, two idle warps; warpgroups 1-2: epilogue
constexpr int ACC_COLS = 96;                          // fp32 accumulator columns per tile
constexpr int TMEM_COLS = 256;                        // 2 accumulator buffers x 96 columns, rounded up to a power of two

struct TcSmemLayout {
  int w_off, stage_off, stage_bytes, x_off, z_off, bar_off, total, nstage;
  int row_bytes, copy_bytes, a_tile_bytes;
};

__host__ __device__ inline TcSmemLayout tc_smem_layout(bool rec, int th, int tw) {
  TcSmemLayout l;
  l.row_bytes = (tw + 8) * PIX_BYTES;         // one row of the operand tile: tw + 2 pixels needed, padded to whole 8-pixel atoms
  l.copy_bytes = (th + 2) * l.row_bytes;      // the operand tile (halo row above and below, halo pixel left and right)
  l.a_tile_bytes = l.copy_bytes;
  l.w_off = 0;
  const int wbytes = rec ? 2 * W_CONV_BYTES : W_CONV_BYTES;
  l.x_off = 0;
  l.z_off = l.a_tile_bytes;                                 // rec: operand tile of the previous spikes
  l.stage_bytes = l.z_off + (rec ? l.a_tile_bytes : 0);
  l.nstage = (227 * 1024 - 1280 - wbytes) / l.stage_bytes;
  if (l.nstage > 4) l.nstage = 4;
  l.stage_off = wbytes;
  l.bar_off = l.stage_off + l.nstage * l.stage_bytes;
  l.total = l.bar_off + 256 + 1024;  // + slack to align the carve-up to 1024 B at run time
  return l;
}

struct TcParams {
  int B, H, W, tiles_x, tiles_y, n_tiles, th, tw;
  int has_rec, has_v, has_z, hard_reset;
  const uint16_t* w_split;
  const float* leak;
  const float* thresh;
  const float* v_in;
  const uint16_t* z_in;  // previous spikes, channels-last (read directly by the epilogue)
  float* v_out;
  uint16_t* z_out;       // new spikes, channels-last (written directly by the epilogue)
  long long* trace;  // debug: per-CTA timeline (clock64), NULL in production
  int skip;          // debug: ablation mask (1 = no v_out stores, 2 = no v_in loads, 4 = no MMAs, 8 = no spike store, 16 = no tmem loads)
};

constexpr int TRACE_SLOTS = 8, TRACE_MAX_TILES = 32;  // [cta][tile][slot]
#define EF_TRACE(it_, slot_)                                                                              \
  do {                                                                                                    \
    if (DEBUG && p.trace && (it_) < TRACE_MAX_TILES)                                                      \
      p.trace[((size_t)blockIdx.x * TRACE_MAX_TILES + (it_)) * TRACE_SLOTS + (slot_)] = clock64() - t_cta; \
  } while (0)


// ---- the kernel ----------------------------------------------------------------------------------------------------
// DEBUG = true compiles in the timeline trace and the ablation switches (ef_debug_tc_trace / ef_debug_tc_skip); the production
// instantiation carries none of that code in its loops.
template <bool HARD, bool DEBUG>
__global__ void __launch_bounds__(TC_THREADS, 1)
lif_conv_fwd_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_zh) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms need an aligned carve-up
  const bool rec = p.has_rec != 0;
  const TcSmemLayout L = tc_smem_layout(rec, p.th, p.tw);
  const int NST = L.nstage;
  const int TH = p.th, TW = p.tw;
  const int skip = DEBUG ? p.skip : 0;
  const uint32_t s_base = smem_u32(smem);
  // barriers: [0] weights, [1..NST] full, [1+NST..2NST] empty, then acc_full[2], acc_empty[2]; then the TMEM address word
  const uint32_t bar_w = s_base + L.bar_off;
  auto bar_full = [&](int s) { return bar_w + 8u * (1 + s); };
  auto bar_empty = [&](int s) { return bar_w + 8u * (1 + NST + s); };
  auto bar_accf = [&](int a) { return bar_w + 8u * (1 + 2 * NST + a); };
  auto bar_acce = [&](int a) { return bar_w + 8u * (3 + 2 * NST + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * (5 + 2 * NST));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_cta = DEBUG ? clock64() : 0;
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);  // MMA commit (the epilogue reads nothing from the operand stages)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);
      mbar_init(bar_acce(a), TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (DEBUG) {
    if (skip & 32) n_my = 0;                        // prologue + teardown only
    else if ((skip & 128) && n_my > 1) n_my = 1;    // one tile per CTA
  }
  const uint32_t stage_tx = L.a_tile_bytes + ((p.has_z && rec) ? L.a_tile_bytes : 0);
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp < 4) {
  // Register re-allocation between the warpgroups: the launch allocates 384 x 168 = 64512 registers; afterwards
  // 128 x 96 + 256 x 200 = 63488 <= 64512 are in use, so the setmaxnreg.inc below can always be satisfied.
  asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");    // the producer warpgroup hands its registers ...
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const uint32_t wbytes = rec ? 2 * W_CONV_BYTES : W_CONV_BYTES;
      if (DEBUG && (skip & 64)) {
        mbar_arrive(bar_w);
      } else {
        mbar_expect_tx(bar_w, wbytes);
        for (uint32_t off = 0; off < wbytes; off += 13824)  // 55296 = 4 x 13824
          bulk_load_1d(s_base + L.w_off + off, reinterpret_cast<const uint8_t*>(p.w_split) + off, 13824, bar_w);
      }
      for (int it = 0; it < n_my; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
        const int ty = r / p.tiles_x;
        const int y0 = ty * TH, x0 = (r - ty * p.tiles_x) * TW;
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        mbar_wait(bar_empty(s), ph ^ 1);
        const uint32_t st = s_base + L.stage_off + s * L.stage_bytes;
        if (DEBUG && (skip & 512)) {  // no tile loads at all (pure barrier ring)
          mbar_arrive(bar_full(s));
          continue;
        }
        mbar_expect_tx(bar_full(s), stage_tx);
        tma_load_4d(st + L.x_off, &map_x, bar_full(s), 0, x0 - 1, y0 - 1, b);
        if (p.has_z && rec) tma_load_4d(st + L.z_off, &map_zh, bar_full(s), 0, x0 - 1, y0 - 1, b);
        EF_TRACE(it, 0);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      const uint64_t b_ff = umma_desc_sw64(s_base + L.w_off, ATOM_BYTES);
      const uint64_t b_rec = umma_desc_sw64(s_base + L.w_off + W_CONV_BYTES, ATOM_BYTES);
      const bool do_rec = rec && p.has_z;
      for (int it = 0; it < n_my; ++it) {
        const int s = it % NST, a = it & 1;
        const uint32_t ph = (it / NST) & 1, aph = (it >> 1) & 1;
        mbar_wait(bar_acce(a), aph ^ 1);
        mbar_wait(bar_full(s), ph);
        tc_fence_after();
        EF_TRACE(it, 1);
        const uint32_t st = s_base + L.stage_off + s * L.stage_bytes;
        const uint32_t d_tmem = tmem_base + a * ACC_COLS;
        // One elected thread issues all MMAs of the tile: this instruction stream is serial, so everything per MMA is
        // reduced to two 64-bit adds on precomputed descriptors (offsets in 16-byte units are compile-time constants).
        if (DEBUG && (skip & (2048 | 4096 | 8192))) {  // timing experiments only (results are wrong): what paces the MMAs?
          const uint64_t ax = umma_desc_sw64(st + L.x_off, L.row_bytes);
          for (int tap = 0; tap < 9; ++tap)
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ad = ax + (uint64_t)((tap / 3) * (1024 / 16) + ((skip & 4096) ? 0 : (tap % 3) * (PIX_BYTES / 16)) + ((skip & 8192) ? 0 : ks * 2));
              const uint64_t bd = b_ff + (uint64_t)(tap * (W_BLOCK_BYTES / 16) + ks * 2);
              if (skip & 2048) umma_bf16<umma_idesc(32)>(d_tmem, ad, bd, (tap | ks) != 0);
              else umma_bf16<umma_idesc(96)>(d_tmem, ad, bd, (tap | ks) != 0);
            }
        } else if (!(DEBUG && (skip & 4))) {
          const uint64_t ax = umma_desc_sw64(st + L.x_off, L.row_bytes);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16<umma_idesc(96)>(d_tmem, ax + (uint64_t)((tap / 3) * (1024 / 16) + (tap % 3) * (PIX_BYTES / 16) + ks * 2),
                        b_ff + (uint64_t)(tap * (W_BLOCK_BYTES / 16) + ks * 2), (tap | ks) != 0);
          }
          if (do_rec) {
            const uint64_t az = umma_desc_sw64(st + L.z_off, L.row_bytes);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                umma_bf16<umma_idesc(96)>(d_tmem, az + (uint64_t)((tap / 3) * (1024 / 16) + (tap % 3) * (PIX_BYTES / 16) + ks * 2),
                          b_rec + (uint64_t)(tap * (W_BLOCK_BYTES / 16) + ks * 2), 1u);
            }
          }
        }
        umma_commit(bar_empty(s));  // the stage's operand tiles may be overwritten once these MMAs have read them
        umma_commit(bar_accf(a));   // accumulator complete
        EF_TRACE(it, 2);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");  // ... to the epilogue warps (three prefetch register sets)
    // =============================== epilogue (8 warps: 4 lane quadrants x 2 channel halves) ===============================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access (warp id % 4)
    const int hsel = (warp - 4) >> 2;       // channel half: channels [16*hsel, 16*hsel + 16)
    const int m = q * 32 + lane;            // GEMM row = pixel within the tile
    const int ph_ = m / TW, pw_ = m % TW;   // (row, col) inside the tile
    const bool store_thread = (threadIdx.x == 128);
    const int c0 = 16 * hsel;
    float lam[16], thr[16];  // 1 - lambda is recomputed per use (one FADD) instead of held in 16 more registers
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      lam[j] = sigmoidf_acc(__ldg(p.leak + c0 + j));
      thr[j] = fmaxf(__ldg(p.thresh + c0 + j), 0.01f);
    }
    const size_t plane = (size_t)p.H * p.W;
    // The membrane potential of the NEXT tile is prefetched into registers while the current tile is processed, so the DRAM
    // latency of these loads (the only operand not staged by TMA) stays off the per-tile critical path.
    struct TileAt {
      int ov;  // element offset of [b][c0][gy][gx] in the membrane tensors, or -1 when the pixel is outside the image / past the last tile
      int oz;  // element offset of [b][gy][gx][c0] in the channels-last spike tensors
    };
    // tile coordinates advance incrementally by gridDim.x tiles per iteration (no integer division inside the loop)
    const int G = gridDim.x;
    const int g_b = G / tiles_per_img, g_r = G - g_b * tiles_per_img, g_ty = g_r / p.tiles_x, g_tx = g_r - g_ty * p.tiles_x;
    int nb = blockIdx.x / tiles_per_img, nty, ntx;
    {
      const int r0 = blockIdx.x - nb * tiles_per_img;
      nty = r0 / p.tiles_x, ntx = r0 - nty * p.tiles_x;
    }
    int n_it = 0;  // iteration index the (nb, nty, ntx) cursor points at
    const int iplane = p.H * p.W;
    auto locate_next = [&]() {
      TileAt t;
      const int gy_ = nty * TH + ph_, gx_ = ntx * TW + pw_;
      const bool in_ = n_it < n_my && gy_ < p.H && gx_ < p.W;
      const int pix_ = (nb * p.H + gy_) * p.W + gx_;
      t.ov = in_ ? (nb * 32 + c0) * iplane + gy_ * p.W + gx_ : -1;
      t.oz = pix_ * 32 + c0;
      // advance the cursor by G tiles
      ntx += g_tx;
      if (ntx >= p.tiles_x) ntx -= p.tiles_x, ++nty;
      nty += g_ty;
      if (nty >= p.tiles_y) nty -= p.tiles_y, ++nb;
      nb += g_b;
      ++n_it;
      return t;
    };
    const bool ld_v = p.has_v && !(DEBUG && (skip & 2));
    auto load_v = [&](const TileAt& t, float (&dst)[16]) {
      if (t.ov >= 0 && ld_v) {
        const float* src = p.v_in + t.ov;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = __ldg(src + j * plane);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = 0.f;
      }
    };
    // The previous spikes of the pixel are read from global memory as well (32 bytes per thread), NOT from the TMA-written
    // operand tile: an ordinary shared-memory load of TMA-written data right after the mbarrier wait occasionally returned
    // a few stale 16-byte pieces (tools/tc_determinism.py), so no generic-proxy read of async-proxy data is left in this kernel.
    auto load_z = [&](const TileAt& t, uint4 (&dst)[2]) {
      const bool ld = t.ov >= 0 && p.has_z && !(DEBUG && (skip & 1024));
      const uint4* src = reinterpret_cast<const uint4*>(p.z_in + t.oz);
      dst[0] = ld ? __ldg(src) : make_uint4(0, 0, 0, 0);
      dst[1] = ld ? __ldg(src + 1) : make_uint4(0, 0, 0, 0);
    };
    // One tile of the epilogue.  (cur, vc, zc) describe the tile processed now (its membrane potential and previous spikes
    // are already in registers or in flight), (nxt, vnx, znx) receive the prefetch of the tile TWO iterations ahead: one
    // tile time (~0.5 us) is shorter than the DRAM latency under load, two are not.  The loop below calls this three times
    // per trip with the three register sets rotated, so no register of a pending load is ever copied (a MOV from an
    // in-flight load would stall for the DRAM latency and serialise the tiles).
    auto tile_body = [&](const int it, const TileAt& cur, const float (&vc)[16], const uint4 (&zc)[2], TileAt& nxt, float (&vnx)[16],
                         uint4 (&znx)[2]) {
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      nxt = locate_next();
      load_v(nxt, vnx);
      load_z(nxt, znx);

      mbar_wait(bar_accf(a), aph);
      tc_fence_after();
      if (store_thread) EF_TRACE(it, 3);
      const uint32_t tacc = tmem_base + a * ACC_COLS + c0 + ((uint32_t)(q * 32) << 16);
      const uint32_t zw[8] = {zc[0].x, zc[0].y, zc[0].z, zc[0].w, zc[1].x, zc[1].y, zc[1].z, zc[1].w};
      float vn[16];
      uint32_t zpk[8];
      // the accumulator is read in two halves of 8 channels (3 x 8 live registers instead of 3 x 16: the three prefetch
      // register sets need the room)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t a_hi[8], a_mid[8], a_lo[8];
        if (!(DEBUG && (skip & 16))) {
          tmem_ld8(tacc + 8 * h, a_hi);
          tmem_ld8(tacc + 32 + 8 * h, a_mid);
          tmem_ld8(tacc + 64 + 8 * h, a_lo);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) a_hi[j] = a_mid[j] = a_lo[j] = 0;
        }
        if (h == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acce(a));  // accumulator buffer may be overwritten by the MMA of tile it+2
          if (store_thread) EF_TRACE(it, 4);
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = 8 * h + jj;
          const float I = __fadd_rn(__fadd_rn(__uint_as_float(a_lo[jj]), __uint_as_float(a_mid[jj])), __uint_as_float(a_hi[jj]));
          const float z = (j & 1) ? bf16_hi(zw[j >> 1]) : bf16_lo(zw[j >> 1]);
          if (HARD) vn[j] = __fadd_rn(__fmul_rn(__fmul_rn(vc[j], lam[j]), __fsub_rn(1.0f, z)), __fmul_rn(__fsub_rn(1.0f, lam[j]), I));
          else vn[j] = __fsub_rn(__fadd_rn(__fmul_rn(vc[j], lam[j]), __fmul_rn(__fsub_rn(1.0f, lam[j]), I)), __fmul_rn(z, thr[j]));
          const uint32_t zb = (__fsub_rn(vn[j], thr[j]) > 0.f) ? 0x3F80u : 0u;  // bf16(1.0) = 0x3F80
          if (j & 1) zpk[j >> 1] |= zb << 16;
          else zpk[j >> 1] = zb;
        }
      }
      if (cur.ov >= 0) {
        if (!(DEBUG && (skip & 1))) {
          float* vout = p.v_out + cur.ov;
#pragma unroll
          for (int j = 0; j < 16; ++j) vout[j * plane] = vn[j];
        }
        // spikes: this thread's 16 channels are 32 contiguous bytes of the channels-last pixel row -- stored straight from
        // registers (fire and forget).  A staged TMA store needs fence.proxy.async, which waits for every outstanding
        // global load of the thread, i.e. it would turn the prefetch above into a synchronous load (ncu: long-scoreboard
        // stalls on the fence were the top stall of the previous version).
        if (!(DEBUG && (skip & 8))) {
          uint4* zout = reinterpret_cast<uint4*>(p.z_out + cur.oz);
          zout[0] = make_uint4(zpk[0], zpk[1], zpk[2], zpk[3]);
          zout[1] = make_uint4(zpk[4], zpk[5], zpk[6], zpk[7]);
        }
      }
      if (store_thread) EF_TRACE(it, 5);
    };
    TileAt tA = locate_next(), tB, tC;
    float vA[16], vB[16], vC[16];
    uint4 zA[2], zB[2], zC[2];
    load_v(tA, vA);
    load_z(tA, zA);
    tB = locate_next();
    load_v(tB, vB);
    load_z(tB, zB);
    for (int it = 0; it < n_my; it += 3) {
      tile_body(it, tA, vA, zA, tC, vC, zC);
      if (it + 1 < n_my) tile_body(it + 1, tB, vB, zB, tA, vA, zA);
      if (it + 2 < n_my) tile_body(it + 2, tC, vC, zC, tB, vB, zB);
    }
    if (store_thread) EF_TRACE(n_my > 0 ? n_my - 1 : 0, 7);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- weight split kernel --------------------------------------------------------------------------------------------
// out block index conv*9 + tap; inside a block row n' = split*32 + n holds K = 32 input channels (64 B); 8-row atoms of
// 512 B; the 16-byte chunk k/8 of row r = n'%8 is stored at chunk (k/8) ^ ((r >> 1) & 3)  (64-byte swizzle)
__global__ void split_weights_kernel(const float* __restrict__ w_ff, const float* __restrict__ w_rec, uint16_t* __restrict__ out, int nconv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over nconv * 32 (n) * 32 (ci) * 9 (tap)
  if (i >= nconv * 32 * 32 * 9) return;
  const int tap = i % 9, ci = (i / 9) % 32, n = (i / (9 * 32)) % 32, cv = i / (9 * 32 * 32);
  const float w = (cv == 0 ? w_ff : w_rec)[(n * 32 + ci) * 9 + tap];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const float r1 = w - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
  const __nv_bfloat16 parts[3] = {hi, mid, lo};
  const size_t blk = (size_t)cv * 9 + tap;
  for (int sp = 0; sp < 3; ++sp) {
    const int nn = sp * 32 + n, r = nn & 7;
    const int chunk = (ci >> 3) ^ ((r >> 1) & 3);
    out[blk * (W_BLOCK_BYTES / 2) + (nn >> 3) * 256 + r * 32 + chunk * 8 + (ci & 7)] = *reinterpret_cast<const uint16_t*>(&parts[sp]);
  }
}

static long long* g_tc_trace = nullptr;  // set through ef_debug_tc_trace (tools/tc_timeline.py)
static int g_tc_skip = 0;                // set through ef_debug_tc_skip (tools/tc_ablation.py)

bool lif_conv_tc_eligible(const ef_lif_conv_params& p) {
  return p.w_split && p.x_cl && p.z_out_cl && p.Cin == 32 && p.C == 32 && p.ksize == 3 && p.stride == 1 && p.neuron == EF_LIF &&
         !p.residual && !p.out && !p.z_out && !p.out_cl && (!p.v_in == !p.z_in_cl) && !p.z_in && !p.x && ((uintptr_t)p.x_cl % 16 == 0) && ((long long)p.B * p.H * p.W * 32 < (1ll << 31)) &&
         ((uintptr_t)p.z_out_cl % 16 == 0) && p.v_in != p.v_out;
}

int lif_conv_fwd_tc(const ef_lif_conv_params& p, cudaStream_t st) {
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const bool rec = p.w_rec != nullptr;
  TcParams q;
  q.B = p.B, q.H = p.H, q.W = p.W;
  q.th = 16, q.tw = 8;  // one 8-pixel atom per tile row: the only shape whose tap-shifted windows are a single descriptor
  static_assert((8 + 8) * PIX_BYTES == 1024, "the MMA issue loop hard-codes a 1024-byte operand row");
  q.tiles_x = cdiv(p.W, q.tw), q.tiles_y = cdiv(p.H, q.th), q.n_tiles = p.B * q.tiles_x * q.tiles_y;
  q.has_rec = rec, q.has_v = p.v_in != nullptr, q.has_z = p.z_in_cl != nullptr, q.hard_reset = p.hard_reset;
  q.w_split = p.w_split, q.leak = p.leak, q.thresh = p.thresh, q.v_in = p.v_in, q.z_in = p.z_in_cl, q.v_out = p.v_out, q.z_out = p.z_out_cl;
  q.trace = g_tc_trace;
  q.skip = g_tc_skip;
  CUtensorMap mx, mzh;
  int rc;
  if ((rc = get_map(p.x_cl, p.B, p.H, p.W, q.th + 2, q.tw + 8, true, &mx))) return rc;
  mzh = mx;  // placeholder when there is no previous state
  if (q.has_z) {
    if (rec && (rc = get_map(p.z_in_cl, p.B, p.H, p.W, q.th + 2, q.tw + 8, true, &mzh))) return rc;
  }
  const TcSmemLayout L = tc_smem_layout(rec, q.th, q.tw);
  const int grid = q.n_tiles < n_sms ? q.n_tiles : n_sms;
  const bool dbg = q.trace != nullptr || q.skip != 0;
  auto kern = p.hard_reset ? (dbg ? lif_conv_fwd_tc_kernel<true, true> : lif_conv_fwd_tc_kernel<true, false>)
                           : (dbg ? lif_conv_fwd_tc_kernel<false, true> : lif_conv_fwd_tc_kernel<false, false>);
  static bool attr_set[4] = {false, false, false, false};
  const int ki = (p.hard_reset ? 2 : 0) + (dbg ? 1 : 0);
  if (!attr_set[ki]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(lif_conv_fwd_tc_kernel)");
    attr_set[ki] = true;
  }
  kern<<<grid, TC_THREADS, L.total, st>>>(q, mx, mzh);
  return check_launch("lif_conv_fwd_tc_kernel");
}

}  // namespace ef

extern "C" int ef_debug_tc_skip(int mask) {  // ablation switches for tools/tc_ablation.py; 0 = production behaviour
  ef::g_tc_skip = mask;
  return EF_OK;
}

extern "C" int ef_debug_tc_trace(long long* buf) {  // buf: device int64 [n_ctas][32 tiles][8 slots] or NULL to switch tracing off
  ef::g_tc_trace = buf;
  return EF_OK;
}

extern "C" int64_t ef_split_weights_elems(int32_t Cin, int32_t C, int32_t has_rec) {
  if (Cin != 32 || C != 32) return 0;
  return (int64_t)(has_rec ? 2 : 1) * ef::W_CONV_BYTES / 2;
}

extern "C" int ef_split_weights(const float* w_ff, const float* w_rec, int32_t Cin, int32_t C, uint16_t* out, void* stream) {
  using namespace ef;
  EF_REQUIRE(w_ff && out, EF_ENULL, "ef_split_weights: NULL tensor");
  EF_REQUIRE(Cin == 32 && C == 32, EF_EUNSUPPORTED, "ef_split_weights: the tensor-core path covers 32 -> 32 channels (got %d -> %d)", Cin, C);
  const int nconv = w_rec ? 2 : 1;
  split_weights_kernel<<<cdiv(nconv * 32 * 32 * 9, 256), 256, 0, as_stream(stream)>>>(w_ff, w_rec, out, nconv);
  return check_launch("split_weights_kernel");
}
