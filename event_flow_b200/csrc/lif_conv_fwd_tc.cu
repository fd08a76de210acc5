// tcgen05 tensor-core path of the fused conv + LIF step (placeholder until the kernel lands: never eligible).
#include "common.cuh"

namespace ef {
bool lif_conv_tc_eligible(const ef_lif_conv_params&) { return false; }
int lif_conv_fwd_tc(const ef_lif_conv_params&, cudaStream_t) { return fail(EF_EUNSUPPORTED, "tensor-core path not built"); }
}  // namespace ef

extern "C" int64_t ef_split_weights_elems(int32_t, int32_t, int32_t) { return 0; }
extern "C" int ef_split_weights(const float*, const float*, int32_t, int32_t, uint16_t*, void*) {
  return ef::fail(EF_EUNSUPPORTED, "tensor-core path not built");
}
