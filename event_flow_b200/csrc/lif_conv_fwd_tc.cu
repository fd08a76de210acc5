// Fused conv3x3 + LIF step on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
//
// Implicit GEMM per 16x8-pixel tile:  D[128 px, 96] = sum over 9 taps, 32 (or 64 with the recurrent conv) input channels
//   A = bf16 spikes, channels-last.  ONE TMA box per tile (18 rows x 10 px x 32 ch: the 16x8 tile and its 1-pixel halo;
//       hardware zero-fill = the conv padding) lands in shared memory densely, 64-byte-swizzled: one pixel = one 64-byte
//       K-major UMMA row, one halo row = 640 bytes.  A tap (dy, dx) is then only a descriptor start address,
//       tile + dy*640 + dx*64, with SBO = 640 (8 consecutive pixels of a tile row = one 8-row group): no im2col copy.
//       Neither the start nor the group stride is swizzle-atom aligned; this works because the swizzle XOR is a function of
//       the absolute shared-memory address bits [7:8] on both the TMA write and the UMMA read (verified bit-exactly against
//       the CUDA-core kernel, tests/test_gpu_tc.py);
//   B = weights, split into three bf16 terms hi+mid+lo == w (exact), resident in shared memory for the whole kernel and
//       stacked along N: one MMA per (tap, k-step) with N = 96 = {hi, mid, lo} x 32 channels, so the A tile is read from
//       shared memory once instead of three times; same swizzled K-major layout (written by ef_split_weights);
//   D = fp32 accumulator in tensor memory, 96 columns (three partial sums, added in the epilogue), double buffered.
// ALL global traffic goes through the TMA unit.  Earlier versions moved the membrane potential and the spikes of the
// epilogue with per-thread LDG/STG: a 16x8-pixel tile of an NCHW fp32 tensor is 32-byte row segments, i.e. 4 L1 wavefronts
// per warp instruction (16 for the 16-byte spike accesses at 64-byte stride) -- ~1500 wavefronts per tile, which made the
// LSU pipe the pacing resource (ncu: profiles/r01_tc_kernel_notes.md).  Now the membrane tile [32 ch][16][8] fp32 and the
// previous spikes are TMA-loaded into a second ring, updated IN PLACE by the epilogue (conflict-free LDS/STS) and
// TMA-stored from there; the stage is handed back to the producer when the store has read it.
// Spikes are {0,1,2}: exactly representable in bf16, so every product is exact and only the fp32 summation order differs
// from the CPU path (no TF32/BF16 rounding may enter a spiking conv).
// Warp roles: warp 0 = operand TMA producer, warp 1 = MMA issuer + TMEM owner, warp 2 = membrane/spike TMA producer,
// warp 3 idle, warps 4.. = epilogue (each owns a TMEM lane quadrant = 32 pixels and CPT channels).  Persistent over tiles;
// mbarrier pipelines between the roles.
// Reference semantics: models/spiking_submodules.py:96-126 (ConvLIF), :516-551 (ConvLIFRecurrent).
#include <stdlib.h>

#include "tc_common.cuh"

namespace ef {

constexpr int TC_TH = 16, TC_TW = 8;                   // output tile: 8 pixels per tile row = one 8-row UMMA group
constexpr int HALO_W = TC_TW + 2, HALO_H = TC_TH + 2;
constexpr int HALO_PITCH = HALO_W * PIX_BYTES;         // 640 B between tile rows of the operand tile
constexpr int HALO_BYTES = HALO_H * HALO_PITCH;        // 11520 B landed by TMA
constexpr int HALO_STAGE = 12288;                      // padded to a multiple of 1024 B
constexpr int V_TILE_BYTES = 32 * 128 * 4;             // [32 ch][16][8] fp32
constexpr int ZC_TILE_BYTES = 128 * PIX_BYTES;         // [16][8][32 ch] bf16, 64B-swizzled
// Weight terms stacked along N.  W_NSPLIT = 3: three bf16 terms, hi + mid + lo == w bit for bit (N = 96).  W_NSPLIT = 2: two
// fp16 terms of the weight scaled by a per-layer power of two s, w*s = hi + mid * 2^-11 + e with |e| <= half an ulp of the
// fp32 weight (N = 64: one third fewer tensor-core cycles, weight image and accumulator columns); the spikes stay bf16
// (kind::f16 encodes the A and B formats independently) -- but the hardware rejects the mixed descriptor (measured: "illegal
// instruction" on B200 for A = BF16, B = F16), and fp16 spikes would need fp16 gradients in the backward MMAs, so the
// two-term variant is kept as a compile-time option only.  The epilogue undoes the scalings exactly (powers of two).
constexpr int W_NSPLIT = 3;
constexpr int W_BLOCK_BYTES = W_NSPLIT * 32 * PIX_BYTES;  // one tap: [W_NSPLIT x 32 ch][32 k] 16-bit, 64B-swizzled
constexpr int W_CONV_BYTES = 9 * W_BLOCK_BYTES;           // 36864 B (55296 B with three terms) per convolution
constexpr int W_CHUNK = W_CONV_BYTES / 4;                 // bulk-copy granule of the weight image
constexpr int ACC_COLS = 32 * W_NSPLIT;                   // fp32 accumulator columns per tile
constexpr int TMEM_COLS = W_NSPLIT == 2 ? 128 : 256;      // 2 accumulator buffers, rounded up to a power of two
constexpr uint32_t W_IDESC = umma_idesc(ACC_COLS, false, false, W_NSPLIT == 2);

template <bool REC>
struct TcCfg {
  static constexpr int NOP = REC ? (W_NSPLIT == 2 ? 3 : 2) : 4;                      // operand stages
  static constexpr int NV = REC ? 3 : 4;                                             // membrane stages
  static constexpr int W_BYTES = (REC ? 2 : 1) * W_CONV_BYTES;
  static constexpr int OP_STAGE = (REC ? 2 : 1) * HALO_STAGE;                        // x halo tile (+ previous-spike halo tile)
  static constexpr int V_STAGE = V_TILE_BYTES + (REC ? 0 : ZC_TILE_BYTES);           // membrane tile (+ centre spikes when not in the halo tile)
  static constexpr int OP_OFF = W_BYTES;
  static constexpr int V_OFF = OP_OFF + NOP * OP_STAGE;
  static constexpr int ZOUT_OFF = V_OFF + NV * V_STAGE;                              // REC: staging of the new spikes
  static constexpr int BAR_OFF = ZOUT_OFF + (REC ? ZC_TILE_BYTES : 0);
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;                                 // + slack to align the carve-up to 1024 B
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

struct TcParams {
  int B, H, W, tiles_x, tiles_y, n_tiles;
  int T;           // feed-forward cells only: steps of a window fused into this launch (x / z_out hold T*B images, step-major); 1 = one step
  int save_all_v;  // T > 1: store the membrane potential of every step (training) or only of the last one (inference)
  int has_v, has_z;
  const uint16_t* w_split;
  const float* leak;
  const float* thresh;
  long long* trace;  // debug: per-CTA timeline (clock64), NULL in production
  int skip;          // debug: 4 = no MMAs, 32 = prologue + teardown only, 128 = one tile per CTA
};

constexpr int TRACE_SLOTS = 8, TRACE_MAX_TILES = 32;  // [cta][tile][slot]
#define EF_TRACE(it_, slot_)                                                                              \
  do {                                                                                                    \
    if (DEBUG && p.trace && (it_) < TRACE_MAX_TILES)                                                      \
      p.trace[((size_t)blockIdx.x * TRACE_MAX_TILES + (it_)) * TRACE_SLOTS + (slot_)] = clock64() - t_cta; \
  } while (0)

// byte address of the 16-byte chunk `c` of the 64-byte row starting at shared address `row` under the 64B swizzle
__device__ __forceinline__ uint32_t sw64(uint32_t row, int c) { return row + ((uint32_t)(c ^ ((row >> 7) & 3)) << 4); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// ---- the kernel ----------------------------------------------------------------------------------------------------
// CPT = channels per epilogue thread (16: 8 epilogue warps, 8: 16 epilogue warps).  DEBUG = true compiles in the timeline
// trace and the ablation switches (ef_debug_tc_trace / ef_debug_tc_skip).
template <bool HARD, bool REC, int CPT, bool DEBUG>
__global__ void __launch_bounds__(128 + 32 * 4 * (32 / CPT), 1)
lif_conv_fwd_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_zh,
                       const __grid_constant__ CUtensorMap map_vin, const __grid_constant__ CUtensorMap map_vout,
                       const __grid_constant__ CUtensorMap map_zc, const __grid_constant__ CUtensorMap map_zout) {
  using C = TcCfg<REC>;
  constexpr int NOP = C::NOP, NV = C::NV;
  constexpr int EPI_WARPS = 4 * (32 / CPT);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t s_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzled tiles want an aligned carve-up
  uint8_t* smem = smem_raw + (s_base - smem_u32(smem_raw));
  const int skip = DEBUG ? p.skip : 0;
  // barriers: weights | operand full[NOP] empty[NOP] | membrane full[NV] empty[NV] | acc full[2] empty[2] | TMEM address word
  const uint32_t bar_w = s_base + C::BAR_OFF;
  auto bar_opf = [&](int s) { return bar_w + 8u * (1 + s); };
  auto bar_ope = [&](int s) { return bar_w + 8u * (1 + NOP + s); };
  auto bar_vf = [&](int s) { return bar_w + 8u * (1 + 2 * NOP + s); };
  auto bar_ve = [&](int s) { return bar_w + 8u * (1 + 2 * NOP + NV + s); };
  auto bar_accf = [&](int a) { return bar_w + 8u * (1 + 2 * NOP + 2 * NV + a); };
  auto bar_acce = [&](int a) { return bar_w + 8u * (3 + 2 * NOP + 2 * NV + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::BAR_OFF + 8 * (5 + 2 * NOP + 2 * NV));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_cta = DEBUG ? clock64() : 0;
  const bool z_from_halo = REC && p.has_z;  // the epilogue reads the previous spikes from the operand stage
  // items of this CTA: its tiles, each for T consecutive steps (time is the INNER loop: the state of a tile stays on chip over the window)
  const int T = REC ? 1 : p.T;
  int n_my = ((p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * T;
  if (DEBUG) {
    if (skip & 32) n_my = 0;                        // prologue + teardown only
    else if ((skip & 128) && n_my > 1) n_my = 1;    // one tile per CTA
  }
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // Item cursor of the single-thread roles (the two producers, the store thread): items are visited in order, so the decode
  // tile -> (image, tile origin) -- two integer divisions -- runs once per TILE, off the per-step path of a fused window
  // (b = image of the state tensors, bt = t*B + b = image of the per-step tensors)
  struct Cursor {
    int tile, t, b, bt, y0, x0;
  };
  auto cur_decode = [&](Cursor& c) {
    c.b = c.tile / tiles_per_img;
    const int r = c.tile - c.b * tiles_per_img, ty = r / p.tiles_x;
    c.y0 = ty * TC_TH, c.x0 = (r - ty * p.tiles_x) * TC_TW;
    c.bt = c.b;
  };
  auto cur_init = [&](Cursor& c) {
    c.tile = blockIdx.x, c.t = 0;
    cur_decode(c);
  };
  auto cur_next = [&](Cursor& c) {
    if (++c.t == T) {
      c.t = 0, c.tile += gridDim.x;
      cur_decode(c);
    } else {
      c.bt += p.B;
    }
  };
  const uint32_t op_tx = HALO_BYTES * (z_from_halo ? 2 : 1);
  auto op_issue = [&](const Cursor& c, int s) {  // operand tiles of the cursor's item into stage s (the caller has waited for the stage)
    const uint32_t st = s_base + C::OP_OFF + s * C::OP_STAGE;
    mbar_expect_tx(bar_opf(s), op_tx);
    tma_load_4d(st, &map_x, bar_opf(s), 0, c.x0 - 1, c.y0 - 1, c.bt);
    if (z_from_halo) tma_load_4d(st + HALO_STAGE, &map_zh, bar_opf(s), 0, c.x0 - 1, c.y0 - 1, c.bt);
  };
  const bool ld_zc = !REC && p.has_z && !(DEBUG && (skip & 1024));
  const bool ld_v = p.has_v && !(DEBUG && (skip & 2));
  const uint32_t v_tx = (ld_v ? V_TILE_BYTES : 0) + (ld_zc ? ZC_TILE_BYTES : 0);
  auto v_issue = [&](const Cursor& c, int s) {  // membrane tile (+ centre spikes) of the cursor's item into stage s
    const uint32_t st = s_base + C::V_OFF + s * C::V_STAGE;
    if (v_tx == 0 || c.t > 0) {  // steps t > 0 of a fused window carry their state in registers: the stage is only the store staging buffer
      mbar_arrive(bar_vf(s));
      return;
    }
    mbar_expect_tx(bar_vf(s), v_tx);
    if (ld_v) tma_load_4d(st, &map_vin, bar_vf(s), c.x0, c.y0, 0, c.b);
    if (ld_zc) tma_load_4d(st + V_TILE_BYTES, &map_zc, bar_vf(s), 0, c.x0, c.y0, c.b);
  };
  // The two producer threads initialise their own barriers and start the first loads right away, BEFORE the CTA-wide sync:
  // barrier setup by the other thread, the TMEM allocation and the descriptor fetches overlap the first DRAM round trip.
  const int n_op0 = n_my < NOP ? n_my : NOP, n_v0 = n_my < NV ? n_my : NV;
  // Programmatic dependent launch: the next kernel of the step may be scheduled as soon as every CTA has passed this point (its
  // CTAs take over SMs as ours exit and run their own prologue); everything of OURS that does not depend on the previous kernel
  // (barrier setup, TMEM allocation, descriptor prefetch, the weight image) happens before pdl_wait().
  pdl_launch_dependents();
  Cursor cur;  // used by threads 0 (operands), 64 (membrane) and 128 (stores), each walking the items on its own
  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_x);
    if (z_from_halo) prefetch_tensormap(&map_zh);
    mbar_init(bar_w, 1);
    for (int s = 0; s < NOP; ++s) {
      mbar_init(bar_opf(s), 1);
      mbar_init(bar_ope(s), 1 + (z_from_halo ? EPI_WARPS : 0));  // MMA commit (+ the epilogue warps that read z from the stage)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);
      mbar_init(bar_acce(a), EPI_WARPS);
    }
    fence_barrier_init();
    mbar_expect_tx(bar_w, C::W_BYTES);  // the weight image is not written by the previous kernel of the step: load it right away
    for (uint32_t off = 0; off < (uint32_t)C::W_BYTES; off += W_CHUNK)
      bulk_load_1d(s_base + off, reinterpret_cast<const uint8_t*>(p.w_split) + off, W_CHUNK, bar_w);
    pdl_wait();  // the spikes of the previous layer
    cur_init(cur);
    for (int it = 0; it < n_op0; ++it) {
      op_issue(cur, it);
      cur_next(cur);
      EF_TRACE(it, 0);
    }
  } else if (threadIdx.x == 64) {
    if (p.has_v) prefetch_tensormap(&map_vin);
    if (ld_zc) prefetch_tensormap(&map_zc);
    for (int s = 0; s < NV; ++s) {
      mbar_init(bar_vf(s), 1);
      mbar_init(bar_ve(s), 1);  // the store thread, once the TMA store has read the stage
    }
    fence_barrier_init();
    pdl_wait();
    cur_init(cur);
    for (int it = 0; it < n_v0; ++it) {
      v_issue(cur, it);
      cur_next(cur);
    }
  } else if (threadIdx.x == 128) {
    prefetch_tensormap(&map_vout);
    prefetch_tensormap(&map_zout);
  }
  if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== operand TMA producer ===============================
    if (lane == 0) {
      for (int it = n_op0; it < n_my; ++it) {
        mbar_wait(bar_ope(it % NOP), ((it / NOP) & 1) ^ 1);
        op_issue(cur, it % NOP);
        cur_next(cur);
        EF_TRACE(it, 0);
      }
    }
  } else if (warp == 2) {
    // =============================== membrane / centre-spike TMA producer ===============================
    if (lane == 0) {
      for (int it = n_v0; it < n_my; ++it) {
        mbar_wait(bar_ve(it % NV), ((it / NV) & 1) ^ 1);
        v_issue(cur, it % NV);
        cur_next(cur);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      const uint64_t b_ff = umma_desc_sw64(s_base, ATOM_BYTES);
      const uint64_t b_rec = umma_desc_sw64(s_base + W_CONV_BYTES, ATOM_BYTES);
      for (int it = 0; it < n_my; ++it) {
        const int s = it % NOP, a = it & 1;
        mbar_wait(bar_acce(a), ((it >> 1) & 1) ^ 1);
        mbar_wait(bar_opf(s), (it / NOP) & 1);
        tc_fence_after();
        EF_TRACE(it, 1);
        const uint32_t st = s_base + C::OP_OFF + s * C::OP_STAGE;
        const uint32_t d_tmem = tmem_base + a * ACC_COLS;
        // One elected thread issues all MMAs of the tile: everything per MMA is two 64-bit adds on precomputed
        // descriptors (offsets in 16-byte units are compile-time constants).
        if (!(DEBUG && (skip & 4))) {
          const uint64_t ax = umma_desc_sw64(st, HALO_PITCH);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16<W_IDESC>(d_tmem, ax + (uint64_t)((tap / 3) * (HALO_PITCH / 16) + (tap % 3) * (PIX_BYTES / 16) + ks * 2),
                                        b_ff + (uint64_t)(tap * (W_BLOCK_BYTES / 16) + ks * 2), (tap | ks) != 0);
          }
          if (z_from_halo) {
            const uint64_t az = umma_desc_sw64(st + HALO_STAGE, HALO_PITCH);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                umma_bf16<W_IDESC>(d_tmem, az + (uint64_t)((tap / 3) * (HALO_PITCH / 16) + (tap % 3) * (PIX_BYTES / 16) + ks * 2),
                                          b_rec + (uint64_t)(tap * (W_BLOCK_BYTES / 16) + ks * 2), 1u);
            }
          }
        }
        umma_commit(bar_ope(s));   // the stage's operand tiles may be overwritten once these MMAs have read them
        umma_commit(bar_accf(a));  // accumulator complete
        EF_TRACE(it, 2);
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue: 4 lane quadrants x (32 / CPT) channel groups ===============================
    const int e = warp - 4;
    const int q = e & 3;                    // TMEM lane quadrant this warp may access (= warp id % 4)
    const int c0 = CPT * (e >> 2);          // this thread's channels [c0, c0 + CPT)
    const int m = q * 32 + lane;            // GEMM row = pixel within the tile
    const int ty = m >> 3, tx = m & 7;      // (row, col) inside the tile
    const bool store_thread = (threadIdx.x == 128);
    if (store_thread) {
      pdl_wait();  // the thread that issues the global stores
      cur_init(cur);
    }
    float lam[CPT], thr[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      lam[j] = sigmoidf_acc(__ldg(p.leak + c0 + j));
      thr[j] = fmaxf(__ldg(p.thresh + c0 + j), 0.01f);
    }
    // power-of-two scales the weight image was built with (appended to it by ef_split_weights); 1 and unused with three terms
    const float inv_s = W_NSPLIT == 2 ? __ldg(reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.w_split) + C::W_BYTES)) : 1.0f;
    const float mid_s = W_NSPLIT == 2 ? __ldg(reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.w_split) + C::W_BYTES) + 1) : 1.0f;
    constexpr int NCH = CPT / 8;  // 16-byte spike chunks per thread
    // state of this thread's pixel x CPT channels.  In a fused window (T > 1) it lives HERE, in registers, from step to step of a
    // tile: only step 0 reads it from the membrane stage, no later step reads any state from memory
    float vc[CPT];
    uint4 zc[NCH];
    int sv = 0, so = 0, t_step = 0;  // stage indices / step inside the window as counters (no per-item modulo by run-time values)
    uint32_t ph_v = 0, ph_o = 0;     // phase parities of the two rings
    for (int it = 0; it < n_my; ++it) {
      const int a = it & 1;
      const uint32_t vst = s_base + C::V_OFF + sv * C::V_STAGE;
      const uint32_t v_addr = vst + (uint32_t)(c0 * 512 + m * 4);  // [ch][16][8] fp32
      mbar_wait(bar_vf(sv), ph_v);  // stage landed (step 0) / free to be used as store staging buffer (later steps)
      if (t_step == 0) {
        if (p.has_v) {
#pragma unroll
          for (int j = 0; j < CPT; ++j) vc[j] = lds_f32(v_addr + j * 512);
        } else {
#pragma unroll
          for (int j = 0; j < CPT; ++j) vc[j] = 0.f;
        }
      }
      uint32_t z_row;  // shared address of this pixel's 64-byte spike row (input in REC = halo tile, else the in-place centre tile)
      if (REC) {
        z_row = s_base + C::OP_OFF + so * C::OP_STAGE + HALO_STAGE + (uint32_t)((ty + 1) * HALO_PITCH + (tx + 1) * PIX_BYTES);
        if (p.has_z) mbar_wait(bar_opf(so), ph_o);  // TMA-written data: observe the barrier before the generic-proxy read
      } else {
        z_row = vst + V_TILE_BYTES + (uint32_t)(m * PIX_BYTES);
      }
      if (t_step == 0) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) zc[k] = p.has_z ? lds_u4(sw64(z_row, c0 / 8 + k)) : make_uint4(0, 0, 0, 0);
      }

      mbar_wait(bar_accf(a), (it >> 1) & 1);
      tc_fence_after();
      if (store_thread) EF_TRACE(it, 3);
      const uint32_t tacc = tmem_base + a * ACC_COLS + c0 + ((uint32_t)(q * 32) << 16);
      float vn[CPT];
      uint32_t zpk[CPT / 2];
#pragma unroll
      for (int h = 0; h < NCH; ++h) {  // the accumulator is read 8 channels at a time (3 x 8 live registers)
        uint32_t a_hi[8], a_mid[8], a_lo[8];
        tmem_ld8(tacc + 8 * h, a_hi);
        tmem_ld8(tacc + 32 + 8 * h, a_mid);
        if (W_NSPLIT == 3) tmem_ld8(tacc + 64 + 8 * h, a_lo);
        tmem_ld_wait();
        if (h == NCH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acce(a));  // accumulator buffer may be overwritten by the MMA of tile it+2
          if (store_thread) EF_TRACE(it, 4);
        }
        const uint32_t zw[4] = {zc[h].x, zc[h].y, zc[h].z, zc[h].w};
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = 8 * h + jj;
          // three terms: lo + mid + hi; two terms: (hi + mid * 2^-11) / s -- the products with powers of two are exact
          const float I = W_NSPLIT == 3
                              ? __fadd_rn(__fadd_rn(__uint_as_float(a_lo[jj]), __uint_as_float(a_mid[jj])), __uint_as_float(a_hi[jj]))
                              : __fmul_rn(__fmaf_rn(__uint_as_float(a_mid[jj]), mid_s, __uint_as_float(a_hi[jj])), inv_s);
          const float z = (jj & 1) ? bf16_hi(zw[jj >> 1]) : bf16_lo(zw[jj >> 1]);
          if (HARD) vn[j] = __fadd_rn(__fmul_rn(__fmul_rn(vc[j], lam[j]), __fsub_rn(1.0f, z)), __fmul_rn(__fsub_rn(1.0f, lam[j]), I));
          else vn[j] = __fsub_rn(__fadd_rn(__fmul_rn(vc[j], lam[j]), __fmul_rn(__fsub_rn(1.0f, lam[j]), I)), __fmul_rn(z, thr[j]));
          const uint32_t zb = (__fsub_rn(vn[j], thr[j]) > 0.f) ? 0x3F80u : 0u;  // bf16(1.0) = 0x3F80
          if (j & 1) zpk[j >> 1] |= zb << 16;
          else zpk[j >> 1] = zb;
        }
      }
      if (z_from_halo) {  // the previous spikes are in registers (they were consumed above): hand the operand stage back
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ope(so));
      }
      // the TMA stores of the previous tile must have read their shared-memory source before it is reused: the membrane
      // stage goes back to its producer, and (REC) the spike staging buffer may be overwritten after the barrier
      if (store_thread && it > 0) {
        bulk_wait_read0();
        mbar_arrive(bar_ve(sv == 0 ? NV - 1 : sv - 1));
      }
      if (REC) named_bar_sync(1, 32 * EPI_WARPS);
      // new state, written in place (same addresses this thread read) / into the spike staging tile.  In a fused window the membrane
      // potential only goes to memory when it is wanted there: every step for training, the last step otherwise
      const bool store_v = T == 1 || p.save_all_v || t_step == T - 1;
      if (store_v) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) sts_f32(v_addr + j * 512, vn[j]);
      }
      const uint32_t zo_row = REC ? (s_base + C::ZOUT_OFF + (uint32_t)(m * PIX_BYTES)) : z_row;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const uint4 zn = make_uint4(zpk[4 * k], zpk[4 * k + 1], zpk[4 * k + 2], zpk[4 * k + 3]);
        sts_u4(sw64(zo_row, c0 / 8 + k), zn);
        zc[k] = zn;  // next step's previous spikes
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j) vc[j] = vn[j];  // next step's previous potential
      fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
      if (store_thread) EF_TRACE(it, 5);
      named_bar_sync(2, 32 * EPI_WARPS);
      if (store_thread) {
        if (store_v && !(DEBUG && (skip & 1))) tma_store_4d(&map_vout, vst, cur.x0, cur.y0, 0, (T == 1 || p.save_all_v) ? cur.bt : cur.b);
        if (!(DEBUG && (skip & 8))) tma_store_4d(&map_zout, REC ? (s_base + C::ZOUT_OFF) : (vst + V_TILE_BYTES), 0, cur.x0, cur.y0, cur.bt);
        bulk_commit();
        cur_next(cur);
        EF_TRACE(it, 6);
      }
      if (++sv == NV) sv = 0, ph_v ^= 1;
      if (++so == NOP) so = 0, ph_o ^= 1;
      if (!REC && ++t_step == T) t_step = 0;
    }
    if (store_thread) {
      bulk_wait0();
      EF_TRACE(n_my > 0 ? n_my - 1 : 0, 7);
    }
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
// 512 B; the 16-byte chunk k/8 of row r = n'%8 is stored at chunk (k/8) ^ ((r >> 1) & 3)  (64-byte swizzle).
// One block: the two-term image needs the largest |w| of the layer (both convolutions share the accumulator, hence the scale).
__global__ void split_weights_kernel(const float* __restrict__ w_ff, const float* __restrict__ w_rec, uint16_t* __restrict__ out,
                                                             int nconv) {
  __shared__ float s_max[32];
  __shared__ float s_scale;
  const int n_el = nconv * 32 * 32 * 9;
  float scale = 1.0f;
  if (W_NSPLIT == 2) {
    float m = 0.f;
    for (int i = threadIdx.x; i < n_el; i += blockDim.x) m = fmaxf(m, fabsf(i < 9216 ? w_ff[i] : w_rec[i - 9216]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < (int)blockDim.x / 32; ++k) m = fmaxf(m, s_max[k]);
      int e = 0;
      if (m > 0.f && m < 3.0e38f) frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
      int k = 14 - e;                               // max|w| * 2^k in [2^13, 2^14): inside the fp16 range with headroom
      k = k > 100 ? 100 : (k < -100 ? -100 : k);
      s_scale = ldexpf(1.0f, k);
      float* tail = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(out) + (size_t)nconv * W_CONV_BYTES);
      tail[0] = ldexpf(1.0f, -k);  // 1 / s
      tail[1] = ldexpf(1.0f, -11); // scale of the second term
      tail[2] = tail[3] = 0.f;
    }
    __syncthreads();
    scale = s_scale;
  }
  // over nconv * 32 (n) * 32 (ci) * 9 (tap); the three-term image needs no layer-wide scale and is built by many blocks
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, ci = (i / 9) % 32, n = (i / (9 * 32)) % 32, cv = i / (9 * 32 * 32);
    const float w = (cv == 0 ? w_ff : w_rec)[(n * 32 + ci) * 9 + tap];
    uint16_t parts[3];
    if (W_NSPLIT == 2) {
      const float ws = w * scale;                                   // exact (power of two)
      const __half hi = __float2half_rn(ws);
      const float r1 = ws - __half2float(hi);                       // exact
      const __half mid = __float2half_rn(r1 * 2048.0f);             // |r1| <= 2^-11 |ws|: the scaled residual is in range again
      parts[0] = __half_as_ushort(hi), parts[1] = __half_as_ushort(mid), parts[2] = 0;
    } else {
      const __nv_bfloat16 hi = __float2bfloat16_rn(w);
      const float r1 = w - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(mid);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
      parts[0] = __bfloat16_as_ushort(hi), parts[1] = __bfloat16_as_ushort(mid), parts[2] = __bfloat16_as_ushort(lo);
    }
    const size_t blk = (size_t)cv * 9 + tap;
    for (int sp = 0; sp < W_NSPLIT; ++sp) {
      const int nn = sp * 32 + n, r = nn & 7;
      const int chunk = (ci >> 3) ^ ((r >> 1) & 3);
      out[blk * (W_BLOCK_BYTES / 2) + (nn >> 3) * 256 + r * 32 + chunk * 8 + (ci & 7)] = parts[sp];
    }
  }
}

// Weight image of the HEAD layer for inputs packed by ef_pack_split_cl: input slot k = s*SL + c (s = 0..2) carries w[n][c][tap], every
// other k is zero -- the tensor core then forms (x_hi + x_mid + x_lo) * (w_hi + w_mid + w_lo) from nine exact partial products.
__global__ void split_weights_head_kernel(const float* __restrict__ w_ff, uint16_t* __restrict__ out, int Cin, int SL) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 32 (n) * 32 (k slot) * 9 (tap)
  if (i >= 32 * 32 * 9) return;
  const int tap = i % 9, k = (i / 9) % 32, n = i / (9 * 32);
  const int c = k % SL, s = k / SL;
  const float w = (s < 3 && c < Cin) ? w_ff[(n * Cin + c) * 9 + tap] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const float r1 = w - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  const uint16_t parts[3] = {__bfloat16_as_ushort(hi), __bfloat16_as_ushort(mid), __bfloat16_as_ushort(lo)};
  for (int sp = 0; sp < 3; ++sp) {
    const int nn = sp * 32 + n, r = nn & 7;
    const int chunk = (k >> 3) ^ ((r >> 1) & 3);
    out[(size_t)tap * (W_BLOCK_BYTES / 2) + (nn >> 3) * 256 + r * 32 + chunk * 8 + (k & 7)] = parts[sp];
  }
}

static long long* g_tc_trace = nullptr;  // set through ef_debug_tc_trace (tools/tc_timeline.py)
static int g_tc_skip = 0;                // set through ef_debug_tc_skip (tools/tc_ablation.py)
static int g_tc_cpt = 0;                 // channels per epilogue thread: 16 (8 epilogue warps), 8 (16 warps), 0 = per kernel kind; ef_debug_tc_cpt

bool lif_conv_tc_eligible(const ef_lif_conv_params& p) {
  return p.w_split && p.x_cl && p.z_out_cl && p.Cin == 32 && p.C == 32 && p.ksize == 3 && p.stride == 1 && p.neuron == EF_LIF &&
         !p.residual && !p.out && !p.z_out && !p.out_cl && (!p.v_in == !p.z_in_cl) && !p.z_in && !p.x && ((uintptr_t)p.x_cl % 16 == 0) &&
         ((uintptr_t)p.z_out_cl % 16 == 0) && ((uintptr_t)p.v_out % 16 == 0) && ((uintptr_t)p.v_in % 16 == 0) && (p.W % 4 == 0) && p.v_in != p.v_out;
}

template <bool HARD, bool REC, int CPT, bool DEBUG>
static int launch_tc(const TcParams& q, int grid, const CUtensorMap* m, cudaStream_t st) {
  auto kern = lif_conv_fwd_tc_kernel<HARD, REC, CPT, DEBUG>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<REC>::TOTAL) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(lif_conv_fwd_tc_kernel)");
    attr_set = true;
  }
  const cudaError_t le = launch_pdl(kern, dim3(grid), dim3(128 + 32 * 4 * (32 / CPT)), TcCfg<REC>::TOTAL, st, q, m[0], m[1], m[2], m[3], m[4], m[5]);
  static const bool debug_capture = getenv("EF_DEBUG_CAPTURE") != nullptr;  // diagnostics: launch result and capture status of the stream
  if (debug_capture) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const cudaError_t ce = cudaStreamIsCapturing(st, &cs);
    fprintf(stderr, "[ef] tc launch REC=%d HARD=%d grid=%d T=%d has_v=%d has_z=%d -> launch rc=%d (%s), capture status=%d (query rc=%d)\n", (int)REC, (int)HARD, grid,
            q.T, q.has_v, q.has_z, (int)le, cudaGetErrorName(le), (int)cs, (int)ce);
  }
  return check_launch("lif_conv_fwd_tc_kernel");
}

template <bool HARD, bool REC>
static int launch_tc2(const TcParams& q, int grid, const CUtensorMap* m, cudaStream_t st, bool dbg, int cpt) {
  if (dbg) return cpt == 8 ? launch_tc<HARD, REC, 8, true>(q, grid, m, st) : launch_tc<HARD, REC, 16, true>(q, grid, m, st);
  return cpt == 8 ? launch_tc<HARD, REC, 8, false>(q, grid, m, st) : launch_tc<HARD, REC, 16, false>(q, grid, m, st);
}

static int lif_conv_fwd_tc_window(const ef_lif_conv_params& p, int T, int save_all_v, cudaStream_t st);

int lif_conv_fwd_tc(const ef_lif_conv_params& p, cudaStream_t st) { return lif_conv_fwd_tc_window(p, 1, 1, st); }

// T > 1 (feed-forward cells only): x_cl / z_out_cl hold T*B images step-major, v_out T*B images (save_all_v) or the B images of the last
// step; v_in / z_in_cl are the B images of the state before step 0.
static int lif_conv_fwd_tc_window(const ef_lif_conv_params& p, int T, int save_all_v, cudaStream_t st) {
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const bool rec = p.w_rec != nullptr;
  TcParams q;
  q.B = p.B, q.H = p.H, q.W = p.W;
  q.tiles_x = cdiv(p.W, TC_TW), q.tiles_y = cdiv(p.H, TC_TH), q.n_tiles = p.B * q.tiles_x * q.tiles_y;
  q.has_v = p.v_in != nullptr, q.has_z = p.z_in_cl != nullptr;
  q.T = T, q.save_all_v = save_all_v;
  q.w_split = p.w_split, q.leak = p.leak, q.thresh = p.thresh;
  q.trace = g_tc_trace;
  q.skip = g_tc_skip;
  CUtensorMap m[6];  // x halo, z halo, v_in, v_out, z centre in, z out
  int rc;
  const int BT = p.B * T;
  if ((rc = get_map(p.x_cl, BT, p.H, p.W, HALO_H, HALO_W, true, &m[0]))) return rc;
  if ((rc = get_map_v(p.v_out, (T == 1 || save_all_v) ? BT : p.B, p.H, p.W, TC_TH, TC_TW, &m[3]))) return rc;
  if ((rc = get_map(p.z_out_cl, BT, p.H, p.W, TC_TH, TC_TW, true, &m[5]))) return rc;
  m[1] = m[0], m[2] = m[3], m[4] = m[5];  // placeholders when there is no previous state
  if (q.has_v && (rc = get_map_v(p.v_in, p.B, p.H, p.W, TC_TH, TC_TW, &m[2]))) return rc;
  if (q.has_z) {
    if (rec && (rc = get_map(p.z_in_cl, p.B, p.H, p.W, HALO_H, HALO_W, true, &m[1]))) return rc;
    if (!rec && (rc = get_map(p.z_in_cl, p.B, p.H, p.W, TC_TH, TC_TW, true, &m[4]))) return rc;
  }
  const int grid = q.n_tiles < n_sms ? q.n_tiles : n_sms;
  const bool dbg = q.trace != nullptr || q.skip != 0;
  // measured (tools/kbench.py, with programmatic dependent launch): 16 epilogue warps (8 channels per thread) win for both kinds
  // (feed-forward 14.0 vs 14.8 us, recurrent 17.0 vs 18.5 us)
  const int cpt = g_tc_cpt ? g_tc_cpt : 8;
  (void)rec;
  if (p.hard_reset) return rec ? launch_tc2<true, true>(q, grid, m, st, dbg, cpt) : launch_tc2<true, false>(q, grid, m, st, dbg, cpt);
  return rec ? launch_tc2<false, true>(q, grid, m, st, dbg, cpt) : launch_tc2<false, false>(q, grid, m, st, dbg, cpt);
}

}  // namespace ef

// Feed-forward 32 -> 32 LIF cell (or the head layer on split inputs) over a whole window of T steps in ONE launch: time is the inner
// loop of every tile, the membrane potential and the previous spikes stay in registers from step to step.
extern "C" int ef_lif_conv_fwd_window(const ef_lif_conv_window_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_lif_conv_fwd_window: params is NULL");
  const ef_lif_conv_window_params& w = *pp;
  EF_REQUIRE(w.B > 0 && w.T > 0 && w.H > 0 && w.W > 0 && w.W % 4 == 0 && (int64_t)w.B * w.T < (1 << 20), EF_EINVAL,
             "ef_lif_conv_fwd_window: bad dimensions (W must be a multiple of 4)");
  EF_REQUIRE(w.x_cl && w.w_split && w.leak && w.thresh && w.v_out && w.z_out_cl, EF_ENULL, "ef_lif_conv_fwd_window: NULL tensor");
  EF_REQUIRE(!w.v_in == !w.z_in_cl, EF_EINVAL, "ef_lif_conv_fwd_window: v_in and z_in_cl come together");
  ef_lif_conv_params p = {};
  p.B = w.B, p.Cin = 32, p.C = 32, p.H = w.H, p.W = w.W, p.ksize = 3, p.stride = 1, p.neuron = EF_LIF, p.hard_reset = w.hard_reset;
  p.x_cl = w.x_cl, p.v_in = w.v_in, p.z_in_cl = w.z_in_cl, p.leak = w.leak, p.thresh = w.thresh, p.w_split = w.w_split;
  p.v_out = w.v_out, p.z_out_cl = w.z_out_cl;
  EF_REQUIRE(((uintptr_t)p.x_cl % 16 == 0) && ((uintptr_t)p.z_out_cl % 16 == 0) && ((uintptr_t)p.v_out % 16 == 0) && ((uintptr_t)p.v_in % 16 == 0) &&
                 p.v_in != p.v_out,
             EF_EINVAL, "ef_lif_conv_fwd_window: tensors must be 16-byte aligned and v_out must not alias v_in");
  return lif_conv_fwd_tc_window(p, w.T, w.save_all_v ? 1 : 0, as_stream(stream));
}

extern "C" int ef_debug_tc_skip(int mask) {  // ablation switches for tools/tc_ablation.py; 0 = production behaviour
  ef::g_tc_skip = mask;
  return EF_OK;
}

extern "C" int ef_debug_tc_cpt(int cpt) {  // 16 = 8 epilogue warps (default), 8 = 16 epilogue warps
  if (cpt != 0 && cpt != 8 && cpt != 16) return EF_EINVAL;
  ef::g_tc_cpt = cpt;
  return EF_OK;
}

extern "C" int ef_debug_tc_trace(long long* buf) {  // buf: device int64 [n_ctas][32 tiles][8 slots] or NULL to switch tracing off
  ef::g_tc_trace = buf;
  return EF_OK;
}

extern "C" int64_t ef_split_weights_elems(int32_t Cin, int32_t C, int32_t has_rec) {
  if (Cin != 32 || C != 32) return 0;
  return (int64_t)(has_rec ? 2 : 1) * ef::W_CONV_BYTES / 2 + 8;  // + 16 bytes: the power-of-two scales of the two-term image
}

extern "C" int ef_split_weights_head(const float* w_ff, int32_t Cin, uint16_t* out, void* stream) {
  using namespace ef;
  static_assert(W_NSPLIT == 3, "the head image assumes the three-term weight split");
  EF_REQUIRE(w_ff && out, EF_ENULL, "ef_split_weights_head: NULL tensor");
  EF_REQUIRE(Cin > 0 && Cin <= EF_HEAD_MAX_CIN, EF_EUNSUPPORTED, "ef_split_weights_head: 1 <= Cin <= %d (got %d)", EF_HEAD_MAX_CIN, Cin);
  split_weights_head_kernel<<<cdiv(32 * 32 * 9, 256), 256, 0, as_stream(stream)>>>(w_ff, out, Cin, EF_HEAD_SLOT(Cin));
  return check_launch("split_weights_head_kernel");
}

extern "C" int ef_split_weights(const float* w_ff, const float* w_rec, int32_t Cin, int32_t C, uint16_t* out, void* stream) {
  using namespace ef;
  EF_REQUIRE(w_ff && out, EF_ENULL, "ef_split_weights: NULL tensor");
  EF_REQUIRE(Cin == 32 && C == 32, EF_EUNSUPPORTED, "ef_split_weights: the tensor-core path covers 32 -> 32 channels (got %d -> %d)", Cin, C);
  const int nconv = w_rec ? 2 : 1;
  if (W_NSPLIT == 2) split_weights_kernel<<<1, 1024, 0, as_stream(stream)>>>(w_ff, w_rec, out, nconv);  // one block: layer-wide maximum
  else split_weights_kernel<<<cdiv(nconv * 32 * 32 * 9, 256), 256, 0, as_stream(stream)>>>(w_ff, w_rec, out, nconv);
  return check_launch("split_weights_kernel");
}
