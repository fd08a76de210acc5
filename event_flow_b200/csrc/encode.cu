// Event encodings on device: per-polarity counts, temporal-bilinear voxel grid, binary event mask, polarity mask.
// Reference: dataloader/encodings.py:30-85 and dataloader/base.py:148-222.
#include "common.cuh"

namespace ef {

__global__ void __launch_bounds__(256) encode_kernel(const ef_encode_params p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= p.N) return;
  const size_t hw = (size_t)p.H * p.W;
  const float4 e = reinterpret_cast<const float4*>(p.events)[(size_t)b * p.N + i];  // ts, y, x, p
  const int yi = (int)e.y, xi = (int)e.z;
  const float pol = e.w;
  if (p.pol_mask) {
    // base.py:207-222: [p>0 ? p : 0, p<0 ? -p : 0]
    reinterpret_cast<float2*>(p.pol_mask)[(size_t)b * p.N + i] = make_float2(pol > 0.f ? pol : 0.f, pol < 0.f ? -pol : 0.f);
  }
  if (yi < 0 || yi >= p.H || xi < 0 || xi >= p.W) return;
  const size_t pix = (size_t)yi * p.W + xi;
  if (p.cnt) {  // encodings.py:70-85: ps * mask_pos = p*p for p>0 (integer-valued, order independent)
    if (pol > 0.f) atomicAdd(p.cnt + ((size_t)b * 2 + 0) * hw + pix, pol * pol);
    if (pol < 0.f) atomicAdd(p.cnt + ((size_t)b * 2 + 1) * hw + pix, pol * pol);
  }
  // index_put_ without accumulate: any writer wins; real events all write |p| = 1.  A padded (p = 0) event must not race a
  // real one at the same pixel and clear the mask, so it does not write at all (the image is zero-filled anyway).
  if (p.mask && pol != 0.f) p.mask[(size_t)b * hw + pix] = fabsf(pol);
  if (p.voxel) {  // encodings.py:48-67
    float ts = __fmul_rn(e.x, (float)(p.num_bins - 1));
    if (p.round_ts) ts = rintf(ts);
    for (int k = 0; k < p.num_bins; ++k) {
      const float w = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(ts, (float)k))));
      if (w != 0.f) atomicAdd(p.voxel + ((size_t)b * p.num_bins + k) * hw + pix, __fmul_rn(pol, w));
    }
  }
}

}  // namespace ef

extern "C" int ef_encode_events(const ef_encode_params* pp, void* stream) {
  using namespace ef;
  EF_REQUIRE(pp, EF_ENULL, "ef_encode_events: params is NULL");
  const ef_encode_params& p = *pp;
  EF_REQUIRE(p.B > 0 && p.N >= 0 && p.H > 0 && p.W > 0, EF_EINVAL, "ef_encode_events: bad dimensions");
  EF_REQUIRE(p.events, EF_ENULL, "ef_encode_events: events is NULL");
  EF_REQUIRE(!p.voxel || p.num_bins >= 1, EF_EINVAL, "ef_encode_events: num_bins must be >= 1");
  cudaStream_t st = as_stream(stream);
  const size_t hw = (size_t)p.H * p.W;
  const size_t n_cnt = (size_t)p.B * 2 * hw, n_vox = (size_t)p.B * p.num_bins * hw, n_mask = (size_t)p.B * hw;
  if (p.cnt && p.voxel && p.mask && p.voxel == p.cnt + n_cnt && p.mask == p.voxel + n_vox) {
    cudaMemsetAsync(p.cnt, 0, (n_cnt + n_vox + n_mask) * sizeof(float), st);  // the three images share one allocation: one fill
  } else {
    if (p.cnt) cudaMemsetAsync(p.cnt, 0, n_cnt * sizeof(float), st);
    if (p.voxel) cudaMemsetAsync(p.voxel, 0, n_vox * sizeof(float), st);
    if (p.mask) cudaMemsetAsync(p.mask, 0, n_mask * sizeof(float), st);
  }
  if (p.N == 0) return EF_OK;
  encode_kernel<<<dim3(cdiv(p.N, 256), p.B), 256, 0, st>>>(p);
  return check_launch("encode_kernel");
}
