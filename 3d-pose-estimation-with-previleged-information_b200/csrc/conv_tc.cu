// tcgen05 implicit-GEMM convolution (placeholder until the tensor-core kernels land).
#include "b2_common.cuh"

bool conv_tc_supported(const B2ConvDesc*, int) { return false; }
size_t conv_tc_workspace_bytes(const B2ConvDesc*, int) { return 0; }
int conv_tc_fprop(const B2ConvDesc*, const void*, const float*, const void*, const float*, void*, float*, float*,
                  double*, void*, cudaStream_t) {
  b2_set_error("conv_tc_fprop: not built");
  return B2_E_UNSUPPORTED;
}
int conv_tc_dgrad(const B2ConvDesc*, const void*, const float*, const void*, const float*, void*, void*,
                  cudaStream_t) {
  b2_set_error("conv_tc_dgrad: not built");
  return B2_E_UNSUPPORTED;
}
int conv_tc_wgrad(const B2ConvDesc*, const void*, const float*, const void*, const float*, float*, void*,
                  cudaStream_t) {
  b2_set_error("conv_tc_wgrad: not built");
  return B2_E_UNSUPPORTED;
}
extern "C" int b2_tc_selftest(const void*, const void*, float*, int32_t, int32_t, int32_t, int32_t, void*) {
  b2_set_error("tc_selftest: not built");
  return B2_E_UNSUPPORTED;
}
