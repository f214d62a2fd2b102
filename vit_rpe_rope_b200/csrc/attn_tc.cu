#include "common.cuh"
#include "kernels.h"
namespace vrr {
bool attn_fwd_tc_supported(int, int, int, int, const vrr_bias_desc*) { return false; }
int attn_fwd_tc(const void*, const vrr_bias_desc*, void*, float*, int, int, int, int, float, cudaStream_t) {
  set_error("attn_fwd_tc: not built");
  return VRR_ERR_UNSUPPORTED;
}
}  // namespace vrr
