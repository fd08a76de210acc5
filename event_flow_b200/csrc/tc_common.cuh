// tcgen05 / TMA / mbarrier plumbing shared by the tensor-core kernels (sm_100a only): PTX wrappers, UMMA descriptors and the
// host-side tensor-map cache.
#pragma once
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace ef {

constexpr int PIX_BYTES = 64;               // 32 channels bf16 = one operand row
constexpr int ATOM_BYTES = 8 * PIX_BYTES;   // 8 pixels = one 64B-swizzle atom (8 rows x 64 B)

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
#ifdef EF_MBAR_POLL
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (kernel error the host sees) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (uint32_t n = 1; !mbar_try_wait(bar, parity); ++n) {
    if ((n & 1023u) == 0 && clock64() - t0 > 4000000000ll) __trap();  // ~2 s at 1.9 GHz
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) { asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// K-major, 64-byte-swizzled shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp): rows of 64 B, 8-row atoms of
// 512 B (SBO between atoms), 16-byte chunks XOR-swizzled by address bits [7:8]; version 1 at bit 46, layout type 4
// (SWIZZLE_64B) at bits 61-63; LBO is not used by swizzled K-major layouts (canonical value 1).
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (4ull << 61);
}
// MN-major, 64-byte-swizzled descriptor (cute/atom/mma_traits_sm100.hpp, "((T,4,m),(8,k)):((1,T,LBO),(4T,SBO))"): 32
// contiguous MN elements per 64-byte row, 8 K-rows per 512-byte atom; LBO = byte stride between 32-element groups along
// MN, SBO = byte stride between 8-row groups along K.
__device__ __forceinline__ uint64_t umma_desc_sw64_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | (4ull << 61);
}
// kind::f16 instruction descriptor (cute/arch/mma_sm100_desc.hpp): D = F32, A = B = BF16, M = 128; a_mn / b_mn select
// MN-major (transposed) operands instead of K-major.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t n, bool a_mn = false, bool b_mn = false, bool b_f16 = false) {
  // bits 4-5 D format (1 = F32), 7-9 A format, 10-12 B format (0 = F16, 1 = BF16; A and B are independent 16-bit types)
  return (1u << 4) | (1u << 7) | ((b_f16 ? 0u : 1u) << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

template <uint32_t IDESC>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {  // 32 lanes x 16 columns, no wait
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {  // 32 lanes x 8 columns, no wait
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- host side: tensor maps -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int B, H, W, kind;  // kind = box rows * 256 + box cols * 2 + swizzled
  bool operator==(const MapKey& o) const { return ptr == o.ptr && B == o.B && H == o.H && W == o.W && kind == o.kind; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)(k.B * 1000003u) ^ ((size_t)k.H << 20) ^ ((size_t)k.W << 8) ^ (size_t)k.kind;
  }
};

// operand copy box: 32 ch x tw px x (th + 2) rows, 64B swizzle; centre box: 32 ch x tw px x th rows, no swizzle
inline int get_map(const void* ptr, int B, int H, int W, int rows, int cols, bool swizzled, CUtensorMap* out) {
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const int kind = rows * 256 + cols * 2 + (swizzled ? 1 : 0);
  const MapKey key{ptr, B, H, W, kind};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return EF_OK;
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(EF_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  CUresult r;
  {
    const cuuint64_t dims[4] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {PIX_BYTES, (cuuint64_t)W * PIX_BYTES, (cuuint64_t)H * W * PIX_BYTES};
    const cuuint32_t box[4] = {32, (cuuint32_t)cols, (cuuint32_t)rows, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzled ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return fail(EF_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d) for kind %d, B=%d H=%d W=%d ptr=%p", (int)r, kind, B, H, W, ptr);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return EF_OK;
}

// membrane tensor fp32 NCHW [B][32][H][W]: box = cols px x rows x 32 ch (one tile of all channels), not swizzled; loads
// zero-fill outside the image, stores clip
inline int get_map_v(const void* ptr, int B, int H, int W, int rows, int cols, CUtensorMap* out) {
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, B, H, W, rows * 256 + cols};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return EF_OK;
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(EF_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 32, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)32 * H * W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)cols, (cuuint32_t)rows, 32, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EF_EINVAL, "cuTensorMapEncodeTiled (membrane) failed (CUresult %d), B=%d H=%d W=%d ptr=%p", (int)r, B, H, W, ptr);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return EF_OK;
}

// General channel counts (multiples of 32): channels-last bf16 [B,H,W,C], box = 32 channels x cols px x rows; and the membrane
// tensor fp32 NCHW [B,C,H,W], box = cols px x rows x 32 channels.  The channel offset of a box is a TMA coordinate.
struct MapKeyC {
  const void* ptr;
  int B, H, W, C, kind;
  bool operator==(const MapKeyC& o) const { return ptr == o.ptr && B == o.B && H == o.H && W == o.W && C == o.C && kind == o.kind; }
};
struct MapKeyCHash {
  size_t operator()(const MapKeyC& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)(k.B * 1000003u) ^ ((size_t)k.H << 20) ^ ((size_t)k.W << 8) ^ ((size_t)k.C << 32) ^ (size_t)k.kind;
  }
};

inline int get_map_c(const void* ptr, int B, int H, int W, int C, int rows, int cols, bool swizzled, CUtensorMap* out) {
  static thread_local std::unordered_map<MapKeyC, CUtensorMap, MapKeyCHash> cache;
  const MapKeyC key{ptr, B, H, W, C, rows * 256 + cols * 2 + (swizzled ? 1 : 0)};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return EF_OK;
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(EF_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {32, (cuuint32_t)cols, (cuuint32_t)rows, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzled ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EF_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d), B=%d H=%d W=%d C=%d ptr=%p", (int)r, B, H, W, C, ptr);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return EF_OK;
}

inline int get_map_vc(const void* ptr, int B, int H, int W, int C, int rows, int cols, CUtensorMap* out) {
  static thread_local std::unordered_map<MapKeyC, CUtensorMap, MapKeyCHash> cache;
  const MapKeyC key{ptr, B, H, W, C, rows * 256 + cols};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return EF_OK;
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(EF_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)cols, (cuuint32_t)rows, 32, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EF_EINVAL, "cuTensorMapEncodeTiled (membrane) failed (CUresult %d), B=%d H=%d W=%d C=%d ptr=%p", (int)r, B, H, W, C, ptr);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return EF_OK;
}

}  // namespace ef
