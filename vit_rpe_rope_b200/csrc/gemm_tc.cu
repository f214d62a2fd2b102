#include "common.cuh"
#include "kernels.h"
namespace vrr {
bool qkv_rope_fwd_tc_supported(int, int, int, int) { return false; }
int qkv_rope_fwd_tc(const void*, const void*, const float*, const float*, void*, int, int, int, int, int, cudaStream_t) {
  set_error("qkv_rope_fwd_tc: not built");
  return VRR_ERR_UNSUPPORTED;
}
}  // namespace vrr
