// Micro-benchmark: cycles per tcgen05.mma (M = 128, N = 96, K = 16, bf16 -> fp32) as a function of the shared-memory layout of the operands.
// Question (profiles/r01_tc_kernel_notes.md, round-2 addendum): the fused cell kernel gets one MMA per ~105 cycles although the math floor
// is 48 -- do K = 16 slices (32 bytes) of 64-byte-swizzled rows cost a whole 64-byte row fetch?  Layouts compared, for A (128 rows) and B
// (96 rows) independently: SWIZZLE_64B rows of 64 bytes (two K slices per row, production) vs SWIZZLE_32B rows of 32 bytes (one K slice).
// Operand contents are irrelevant (timing only).  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_layout_probe mma_layout_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo, uint64_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
__host__ __device__ constexpr uint32_t idesc(uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

template <uint32_t N>
__global__ void __launch_bounds__(128, 1) probe(int a_sw32, int b_sw32, int n_mma, int a_sbo64, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 64 * 1024;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;  // defined operand bits (zeros)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // production A: 64-byte rows, 8-row groups every `a_sbo64` bytes (640 = halo-tile rows of 10 pixels; 512 = dense atoms); SW32: 32-byte rows
    const uint64_t a0 = a_sw32 ? desc(a_base, a_sbo64 / 2, 6) : desc(a_base, a_sbo64, 4);
    const uint64_t b0 = b_sw32 ? desc(b_base, 256, 6) : desc(b_base, 512, 4);
    const long long t0 = clock64();
    // 18 (tap, k-slice) positions per round, fully unrolled with compile-time offsets like the kernel (per MMA: two 64-bit adds + the issue):
    // taps shift the A start, k-slices pick the half row (SW64) / the second plane (SW32)
    const uint32_t a_tap = a_sw32 ? 20u : (uint32_t)(a_sbo64 / 16), a_dx = a_sw32 ? 2u : 4u, a_ks = a_sw32 ? (24u * 1024 / 16) : 2u;
    const uint32_t b_tap = b_sw32 ? (96u * 32 / 16) : (96u * 64 / 16), b_ks = b_sw32 ? (9u * 96 * 32 / 16) : 2u;
    for (int r = 0; r < n_mma / 18; ++r) {
#pragma unroll
      for (int pos = 0; pos < 18; ++pos) {
        const int tap = pos >> 1, ks = pos & 1;
        const uint32_t a_off = (tap / 3) * a_tap + (tap % 3) * a_dx + ks * a_ks;
        const uint32_t b_off = tap * b_tap + ks * b_ks;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(a0 + a_off), "l"(b0 + b_off), "r"(idesc(N)), "r"(1u)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

template <uint32_t N>
static void run(const char* what, int a32, int b32, int sbo) {
  const int n_mma = 18 * 400, n_cta = 148;
  long long* d;
  cudaMalloc(&d, n_cta * sizeof(long long));
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) probe<N><<<n_cta, 128, 200 * 1024>>>(a32, b32, n_mma, sbo, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < n_cta; ++i) s += (double)h[i];
  printf("%-58s N=%3u  %7.1f cycles / MMA   (%s)\n", what, N, s / n_cta / n_mma, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<96>("A SW64 (SBO 640, halo tile)  B SW64   [production]", 0, 0, 640);
  run<96>("A SW64 (SBO 512, dense atoms) B SW64", 0, 0, 512);
  run<96>("A SW32 (SBO 320)             B SW64", 1, 0, 640);
  run<96>("A SW64 (SBO 640)             B SW32", 0, 1, 640);
  run<96>("A SW32 (SBO 320)             B SW32", 1, 1, 640);
  run<32>("A SW64 (SBO 640)             B SW64", 0, 0, 640);
  run<32>("A SW32 (SBO 320)             B SW32", 1, 1, 640);
  run<64>("A SW64 (SBO 640)             B SW64", 0, 0, 640);
  run<64>("A SW32 (SBO 320)             B SW32", 1, 1, 640);
  return 0;
}
