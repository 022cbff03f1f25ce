// wgrad_tc.cu -- tcgen05 weight-gradient kernel (placeholder until the MN-major path lands).
#include "common.cuh"

bool b2s_wgrad_tc_supported(int32_t, int32_t, int32_t, int64_t) { return false; }
int b2s_conv_wgrad_tc(const float*, const float*, const int32_t*, int64_t, int32_t, int32_t, int32_t, float*,
                      cudaStream_t) {
  b2s_set_error("b2s_conv_wgrad_tc: not built");
  return -1;
}
