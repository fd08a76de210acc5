// Data-parallel optimiser step as ONE kernel per rank: one-shot all-reduce(SUM) of the flat gradient over NVLink / NVSwitch PEER MEMORY,
// global-norm clip and Adam (train_flow.py:157-163 under data parallelism, SURVEY 8 e / K18).  Replaces ncclAllReduce + ef_grad_sqnorm +
// ef_clip_adam + the zero fill of the gradient buffer: the whole gradient of a LIFFireNet is 299 KB, so the step is pure latency and a
// single launch that loads the peers' buffers directly is the cheapest form (every rank reads all buffers and reduces them in rank order:
// the same sums on every rank, bit for bit, so the replicas stay identical without a broadcast).
//
// Protocol (signal words live in every rank's own memory and are mapped into the peers through CUDA IPC; `epoch` = the step number):
//   A  signal[0] = epoch with system-scope release  ("my gradient is complete");  wait until every rank's signal[0] >= epoch
//   B  g = sum over ranks (fixed order) of their gradient slices, kept in registers;  per-CTA partial of sum g^2  -> grid barrier ->
//      total in CTA order;  clip coefficient;  Adam on this thread's elements
//   C  signal[1] = epoch ("I have read everybody's gradient");  wait until every rank's signal[1] >= epoch;  zero the own gradient
// Bounded waits (timeout_ms, default ten minutes like NCCL's watchdog): a rank that never arrives ends the kernel with status = 1
// (graceful, used by the set-up self test with a short limit) or traps.
#include <string.h>

#include "common.cuh"

namespace ef {

constexpr int DP_THREADS = 256, DP_MAX_PER_THREAD = 8;

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ float ld_peer(const float* p) {  // peer memory: never from a stale local cache line
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long wall_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// every rank's word `which` has reached `epoch`?  threads 0 .. world-1 of the CTA poll one rank each.  Returns false on time-out
// (p.timeout_ms of wall-clock time: a peer may be late for good reasons -- data loading, a checkpoint, validation on one rank).
__device__ __forceinline__ bool wait_all(const ef_dp_step_params& p, int which, int* s_fail) {
  if (threadIdx.x < p.world) {
    const uint32_t* w = p.signals[threadIdx.x] + which;
    const unsigned long long t0 = wall_ns(), limit = (unsigned long long)(p.timeout_ms > 0 ? p.timeout_ms : 600000) * 1000000ull;
    unsigned polls = 0;
    while ((int32_t)(ld_acquire_sys(w) - p.epoch) < 0) {
      if ((++polls & 1023u) == 0 && wall_ns() - t0 > limit) {
        *s_fail = 1;
        break;
      }
      if (polls > 4096u) __nanosleep(200);  // a long wait: leave the memory system alone
    }
  }
  __syncthreads();
  return *s_fail == 0;
}

__global__ void __launch_bounds__(DP_THREADS) dp_step_kernel(const ef_dp_step_params p) {
  __shared__ float s_red[DP_THREADS / 32];
  __shared__ float s_total;
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  uint32_t* mine = p.signals[p.rank];
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // ---- A
  if (blockIdx.x == 0 && tid == 0) {
    __threadfence_system();  // the gradient kernels of this stream are complete (kernel boundary); make their writes visible system-wide
    st_release_sys(mine + 0, p.epoch);
  }
  if (!wait_all(p, 0, &s_fail)) {
    if (tid == 0 && blockIdx.x == 0) *p.status = 1;
    if (!p.graceful) __trap();
    return;
  }
  // ---- B
  const int total_threads = gridDim.x * DP_THREADS, gtid = blockIdx.x * DP_THREADS + tid;
  float g[DP_MAX_PER_THREAD];
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < DP_MAX_PER_THREAD; ++k) {
    const int i = gtid + k * total_threads;
    g[k] = 0.f;
    if (i < p.n) {
      for (int r = 0; r < p.world; ++r) g[k] += ld_peer(p.grads[r] + i);  // rank order: identical sums on every rank
      sq = fmaf(g[k], g[k], sq);
    }
  }
  sq = warp_sum(sq);
  if ((tid & 31) == 0) s_red[tid >> 5] = sq;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < DP_THREADS / 32; ++w) t += s_red[w];
    p.scratch[blockIdx.x] = t;
    // grid barrier on the own counters (word 8: arrivals of this epoch)
    __threadfence();
    atomicAdd(mine + 8, 1u);
    const uint32_t target = p.epoch_launches * gridDim.x;
    const long long t0 = clock64();
    uint32_t seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine + 8) : "memory");
      if (clock64() - t0 > 6000000000ll) {
        s_fail = 1;
        break;
      }
    } while ((int32_t)(seen - target) < 0);
    float total = 0.f;
    for (unsigned c = 0; c < gridDim.x; ++c) total += __ldcg(p.scratch + c);  // CTA order: the same total in every CTA and on every rank
    s_total = total;
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0 && blockIdx.x == 0) *p.status = 2;
    if (!p.graceful) __trap();
    return;
  }
  const float total = s_total;
  float coef = 1.0f;
  if (p.clip > 0.f) coef = fminf(p.clip / (sqrtf(total) + 1e-6f), 1.0f);  // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
#pragma unroll
  for (int k = 0; k < DP_MAX_PER_THREAD; ++k) {
    const int i = gtid + k * total_threads;
    if (i < p.n) {
      const float gi = g[k] * coef;
      const float mi = p.beta1 * p.m[i] + (1.0f - p.beta1) * gi;
      const float vi = p.beta2 * p.v[i] + (1.0f - p.beta2) * gi * gi;
      p.m[i] = mi;
      p.v[i] = vi;
      const float denom = sqrtf(vi) / p.bc2_sqrt + p.eps;
      p.param[i] -= (p.lr / p.bc1) * (mi / denom);
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    p.sqnorm[0] = total;
    st_release_sys(mine + 1, p.epoch);  // (after the grid barrier: every CTA of this rank has read the peers' gradients)
  }
  // ---- C
  if (!wait_all(p, 1, &s_fail)) {
    if (tid == 0 && blockIdx.x == 0) *p.status = 3;
    if (!p.graceful) __trap();
    return;
  }
  float* own = const_cast<float*>(p.grads[p.rank]);
#pragma unroll
  for (int k = 0; k < DP_MAX_PER_THREAD; ++k) {
    const int i = gtid + k * total_threads;
    if (i < p.n) own[i] = 0.f;
  }
}

}  // namespace ef

extern "C" int ef_dp_step(const ef_dp_step_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_dp_step: params is NULL");
  ef_dp_step_params p = *pp;
  EF_REQUIRE(p.world >= 1 && p.world <= EF_DP_MAX_RANKS && p.rank >= 0 && p.rank < p.world && p.n > 0, EF_EINVAL, "ef_dp_step: bad world / rank / n");
  EF_REQUIRE(p.param && p.m && p.v && p.sqnorm && p.scratch && p.status, EF_ENULL, "ef_dp_step: NULL tensor");
  for (int r = 0; r < p.world; ++r) EF_REQUIRE(p.grads[r] && p.signals[r], EF_ENULL, "ef_dp_step: NULL peer pointer of rank %d", r);
  EF_REQUIRE(p.step >= 1 && p.epoch_launches >= 1, EF_EINVAL, "ef_dp_step: step / epoch_launches count from 1");
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int grid = cdiv(p.n, DP_THREADS * 4);  // ~4 elements per thread; all CTAs co-resident (grid barrier)
  if (grid > n_sms) grid = n_sms;
  EF_REQUIRE((long long)grid * DP_THREADS * DP_MAX_PER_THREAD >= p.n, EF_EUNSUPPORTED, "ef_dp_step: at most %d parameters on this device",
             n_sms * DP_THREADS * DP_MAX_PER_THREAD);
  EF_REQUIRE(p.grid_expected == 0 || p.grid_expected == grid, EF_EINVAL, "ef_dp_step: grid %d differs from the one the barrier counters assume (%d)", grid,
             p.grid_expected);
  p.bc1 = 1.0f - powf(p.beta1, (float)p.step);
  p.bc2_sqrt = sqrtf(1.0f - powf(p.beta2, (float)p.step));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(DP_THREADS), cfg.dynamicSmemBytes = 0, cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, dp_step_kernel, p);
  return check_launch("dp_step_kernel");
}

extern "C" int32_t ef_dp_step_grid(int32_t n) {  // CTAs ef_dp_step launches for n parameters (size of `scratch`)
  int dev = 0, n_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = ef::cdiv(n, ef::DP_THREADS * 4);
  return grid > n_sms ? n_sms : grid;
}

// ---- buffers the ranks share through CUDA IPC -------------------------------------------------------------------------------------------
// The exporting rank allocates with cudaMalloc (an allocation of its own: the IPC handle then names exactly this buffer, offset 0) and
// hands the 64-byte handle to its peers; a peer opens it with its OWN device current and cudaIpcMemLazyEnablePeerAccess, which maps the
// buffer into the peer's context with peer access over NVLink.
extern "C" int ef_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
  using namespace ef;
  EF_REQUIRE(bytes > 0 && ptr && handle64, EF_ENULL, "ef_ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* q = nullptr;
  EF_REQUIRE(cudaMalloc(&q, (size_t)bytes) == cudaSuccess, EF_EUNSUPPORTED, "ef_ipc_alloc: cudaMalloc(%lld) failed", (long long)bytes);
  cudaMemset(q, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, q) != cudaSuccess) {
    const cudaError_t e = cudaGetLastError();
    cudaFree(q);
    return fail(EF_EUNSUPPORTED, "ef_ipc_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = q;
  return EF_OK;
}
extern "C" int ef_ipc_open(const unsigned char* handle64, void** ptr) {
  using namespace ef;
  EF_REQUIRE(handle64 && ptr, EF_ENULL, "ef_ipc_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* q = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(EF_EUNSUPPORTED, "ef_ipc_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
  }
  *ptr = q;
  return EF_OK;
}
extern "C" int ef_ipc_close(void* ptr) {
  using namespace ef;
  if (!ptr) return EF_OK;
  const cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e == cudaSuccess) return EF_OK;
  cudaGetLastError();
  return fail(EF_EUNSUPPORTED, "ef_ipc_close: cudaIpcCloseMemHandle failed: %s", cudaGetErrorString(e));
}
extern "C" int ef_ipc_free(void* ptr) {
  using namespace ef;
  if (!ptr) return EF_OK;
  const cudaError_t e = cudaFree(ptr);
  if (e == cudaSuccess) return EF_OK;
  cudaGetLastError();
  return fail(EF_EUNSUPPORTED, "ef_ipc_free: cudaFree failed: %s", cudaGetErrorString(e));
}
