// Shared helpers for libeventflow.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/eventflow.h"

namespace ef {

// ---- error reporting (thread-local message behind ef_last_error) ------------------------------------------------
char* last_error_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define EF_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) return ef::fail(code, __VA_ARGS__); \
  } while (0)

// ---- neuron maths, shared by the CUDA-core and the tensor-core kernels ------------------------------------------
// Per-channel constants derived from the raw module parameters (spiking_submodules.py:108-112, 201-208, 309-315).
struct ChanConst {
  float lam;    // sigmoid(leak | leak_v)
  float thr;    // clamp_min(thresh, 0.01)            (LIF, PLIF)
  float rho;    // sigmoid(leak_pt | leak_t)          (PLIF, ALIF, XLIF)
  float alpha;  // sigmoid(add_pt)                    (PLIF)
  float t0;     // clamp_min(t0, 0.01)                (ALIF, XLIF)
  float t1;     // clamp_min(t1, 0)                   (ALIF, XLIF)
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ ChanConst load_chan_const(const ef_lif_conv_params& p, int c) {
  ChanConst k;
  k.lam = sigmoidf_acc(p.leak[c]);
  k.thr = p.thresh ? fmaxf(p.thresh[c], 0.01f) : 0.f;
  k.rho = p.leak_aux ? sigmoidf_acc(p.leak_aux[c]) : 0.f;
  k.alpha = p.add_pt ? sigmoidf_acc(p.add_pt[c]) : 0.f;
  k.t0 = p.t0 ? fmaxf(p.t0[c], 0.01f) : 0.f;
  k.t1 = p.t1 ? fmaxf(p.t1[c], 0.f) : 0.f;
  return k;
}

// One neuron update.  The op order (and the absence of FMA contraction) follows the reference expressions so that the
// only difference to the CPU path is the summation order inside the convolution.
//   I: input current (ff [+ rec]); v,z,aux: previous state; P: avgpool(mean|x|) (PLIF/XLIF only)
template <int NEURON, bool HARD>
__device__ __forceinline__ void neuron_update(float I, float v, float z, float aux, float P, const ChanConst& k, float& v_out,
                                              float& z_out, float& aux_out, float& thr_eff) {
  const float oml = __fsub_rn(1.0f, k.lam);
  if (NEURON == EF_LIF) {
    thr_eff = k.thr;
    aux_out = 0.f;
    if (HARD)
      v_out = __fadd_rn(__fmul_rn(__fmul_rn(v, k.lam), __fsub_rn(1.0f, z)), __fmul_rn(oml, I));
    else
      v_out = __fsub_rn(__fadd_rn(__fmul_rn(v, k.lam), __fmul_rn(oml, I)), __fmul_rn(z, k.thr));
  } else if (NEURON == EF_PLIF) {
    thr_eff = k.thr;
    aux_out = __fadd_rn(__fmul_rn(aux, k.rho), __fmul_rn(__fsub_rn(1.0f, k.rho), P));
    const float cur = __fsub_rn(I, __fmul_rn(k.alpha, aux_out));
    if (HARD)
      v_out = __fadd_rn(__fmul_rn(__fmul_rn(v, k.lam), __fsub_rn(1.0f, z)), __fmul_rn(oml, cur));
    else
      v_out = __fsub_rn(__fadd_rn(__fmul_rn(v, k.lam), __fmul_rn(oml, cur)), __fmul_rn(z, k.thr));
  } else {  // ALIF: trace of own spikes; XLIF: pre-synaptic trace.  Both adapt the threshold.
    const float drive = (NEURON == EF_ALIF) ? z : P;
    aux_out = __fadd_rn(__fmul_rn(aux, k.rho), __fmul_rn(__fsub_rn(1.0f, k.rho), drive));
    thr_eff = __fadd_rn(k.t0, __fmul_rn(k.t1, aux_out));
    if (HARD)
      v_out = __fadd_rn(__fmul_rn(__fmul_rn(v, k.lam), __fsub_rn(1.0f, z)), __fmul_rn(oml, I));
    else
      v_out = __fsub_rn(__fadd_rn(__fmul_rn(v, k.lam), __fmul_rn(oml, I)), __fmul_rn(z, __fadd_rn(k.t0, __fmul_rn(k.t1, aux))));
  }
  z_out = (__fsub_rn(v_out, thr_eff) > 0.f) ? 1.0f : 0.f;
}

// Surrogate derivative d spike / d (v - thresh), models/spiking_util.py:39-43,56-65,75-79,89-93.
__device__ __forceinline__ float surrogate_grad(int kind, float x, float w) {
  switch (kind) {
    case EF_ARCTAN:
      return 1.0f / (1.0f + w * x * x);
    case EF_SUPERSPIKE: {
      const float d = 1.0f + w * fabsf(x);
      return 1.0f / (d * d);
    }
    case EF_TRIANGLE:
      return fmaxf(1.0f - w * fabsf(x), 0.f);
    default: {  // multi-gaussian
      const float inv_s2pi = 0.3989422804014327f;
      auto gs = [&](float mu, float s) { return expf(-((x - mu) * (x - mu)) / (2.f * s * s)) / s * inv_s2pi; };
      return 1.15f * gs(0.f, w) - 0.15f * gs(w, 6.f * w) - 0.15f * gs(-w, 6.f * w);
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may be scheduled while its predecessor in the stream
// is still draining; it must execute pdl_wait() before touching anything the predecessor wrote, and lets ITS successor start
// early with pdl_launch_dependents().  Inside a captured CUDA graph this becomes a programmatic dependency edge.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();  // api.cu: ef_debug_pdl / EF_PDL=0 switch it off

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace ef
