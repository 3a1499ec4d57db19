// C-ABI entry points of libb2pose: argument validation, error reporting and the dispatch
// between the tensor-core (tcgen05) and CUDA-core (FFMA) convolution kernels.
#include <cstdarg>
#include <cstdio>

#include "b2_common.cuh"

static thread_local char g_err[512] = "";

void b2_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int g_b2_sm_reserve = 0;

extern "C" int b2_abi_version(void) { return B2_ABI_VERSION; }
extern "C" int b2_set_sm_reserve(int32_t sms) {
  B2_REQUIRE(sms >= 0 && sms < 128, B2_E_BADARG, "set_sm_reserve: %d SMs", sms);
  g_b2_sm_reserve = sms;
  return B2_OK;
}
extern "C" const char* b2_last_error(void) { return g_err; }

static int check_desc(const B2ConvDesc* d) {
  B2_REQUIRE(d != nullptr, B2_E_BADARG, "conv: null descriptor");
  B2_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->K > 0 && d->R > 0 && d->S > 0, B2_E_BADARG,
             "conv: non-positive dimension (N=%d H=%d W=%d C=%d K=%d R=%d S=%d)", d->N, d->H, d->W, d->C, d->K,
             d->R, d->S);
  B2_REQUIRE(d->stride > 0 && d->dil > 0 && d->pad >= 0, B2_E_BADARG, "conv: bad stride/pad/dilation");
  int ho = (d->H + 2 * d->pad - d->dil * (d->R - 1) - 1) / d->stride + 1;
  int wo = (d->W + 2 * d->pad - d->dil * (d->S - 1) - 1) / d->stride + 1;
  B2_REQUIRE(ho == d->Ho && wo == d->Wo && ho > 0 && wo > 0, B2_E_BADARG,
             "conv: output size %dx%d does not match the geometry (%dx%d expected)", d->Ho, d->Wo, ho, wo);
  B2_REQUIRE(d->dtype == B2_F32 || d->dtype == B2_BF16, B2_E_UNSUPPORTED, "conv: dtype %d", d->dtype);
  B2_REQUIRE((long long)d->N * d->H * d->W < (1LL << 31) && (long long)d->N * d->Ho * d->Wo < (1LL << 31),
             B2_E_UNSUPPORTED, "conv: more than 2^31 pixels");
  return B2_OK;
}

static bool use_tc(const B2ConvDesc* d, int op) {
  if (d->flags & B2_CONV_FORCE_FFMA) return false;
  return conv_tc_supported(d, op);
}

extern "C" int b2_conv_uses_tensor_cores(const B2ConvDesc* d, int op) {
  if (check_desc(d) != B2_OK) return 0;
  return use_tc(d, op) ? 1 : 0;
}

extern "C" size_t b2_conv_workspace_bytes(const B2ConvDesc* d, int op) {
  if (check_desc(d) != B2_OK) return 0;
  return use_tc(d, op) ? conv_tc_workspace_bytes(d, op) : 0;
}

extern "C" int b2_pconv_fprop(const B2ConvDesc* d, const void* x, const float* mask_in, const void* w,
                              const float* bias, void* y, float* mask_out, float* ratio_out, float* bn_sums,
                              void* workspace, size_t ws_bytes, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  B2_REQUIRE(x && w && y, B2_E_BADARG, "pconv_fprop: null tensor");
  const bool partial = d->flags & B2_CONV_PARTIAL;
  B2_REQUIRE(!partial || mask_in, B2_E_BADARG, "pconv_fprop: partial convolution needs mask_in");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(d, 0)) {
    B2_REQUIRE(ws_bytes >= conv_tc_workspace_bytes(d, 0), B2_E_WORKSPACE, "pconv_fprop: workspace too small");
    return conv_tc_fprop(d, x, mask_in, w, bias, y, mask_out, ratio_out, bn_sums, workspace, st);
  }
  B2_REQUIRE(!(d->flags & B2_CONV_X_CONCAT), B2_E_UNSUPPORTED,
             "pconv_fprop: B2_CONV_X_CONCAT needs the bf16 tensor-core path (plain stride-1 layer, C %% 128 == 0)");
  rc = conv_ffma_fprop(d, x, mask_in, w, bias, y, mask_out, ratio_out, st);
  if (rc) return rc;
  if (bn_sums && (d->flags & B2_CONV_BN_TOTALS))
    return b2_bn_stats_totals(y, (int64_t)d->N * d->Ho * d->Wo, d->K, d->dtype, bn_sums, stream);
  if (bn_sums) return b2_bn_stats(y, (int64_t)d->N * d->Ho * d->Wo, d->K, d->dtype, bn_sums, stream);
  return B2_OK;
}

extern "C" int b2_pconv_dgrad(const B2ConvDesc* d, const void* dy, const float* ratio, const void* w,
                              const float* mask_in, void* dx, void* workspace, size_t ws_bytes, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  B2_REQUIRE(dy && w && dx, B2_E_BADARG, "pconv_dgrad: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(d, 1)) {
    B2_REQUIRE(ws_bytes >= conv_tc_workspace_bytes(d, 1), B2_E_WORKSPACE, "pconv_dgrad: workspace too small");
    return conv_tc_dgrad(d, dy, ratio, w, mask_in, dx, workspace, st);
  }
  B2_REQUIRE(!(d->flags & B2_CONV_W_PREPARED), B2_E_UNSUPPORTED,
             "pconv_dgrad: B2_CONV_W_PREPARED needs the tensor-core path");
  B2_REQUIRE(!(d->flags & B2_CONV_DX_ACCUMULATE), B2_E_UNSUPPORTED,
             "pconv_dgrad: B2_CONV_DX_ACCUMULATE needs the bf16 tensor-core path (stride 1, C %% 64 == 0)");
  B2_REQUIRE(!(d->flags & B2_CONV_X_CONCAT), B2_E_UNSUPPORTED,
             "pconv_dgrad: B2_CONV_X_CONCAT needs the bf16 tensor-core path (plain stride-1 layer, C %% 128 == 0)");
  return conv_ffma_dgrad(d, dy, ratio, w, mask_in, dx, st);
}

extern "C" size_t b2_pconv_dgrad_filter_bytes(const B2ConvDesc* d) {
  if (check_desc(d) != B2_OK || !use_tc(d, 1) || !conv_tc_dgrad_needs_filter(d)) return 0;
  return (size_t)d->K * d->R * d->S * d->C * 2;
}

extern "C" int b2_pconv_dgrad_filter(const B2ConvDesc* d, const void* w, void* wt, size_t wt_bytes, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  B2_REQUIRE(w && wt, B2_E_BADARG, "pconv_dgrad_filter: null tensor");
  B2_REQUIRE(use_tc(d, 1), B2_E_UNSUPPORTED, "pconv_dgrad_filter: this dgrad does not run on the tensor-core path");
  B2_REQUIRE(wt_bytes >= (size_t)d->K * d->R * d->S * d->C * 2, B2_E_WORKSPACE, "pconv_dgrad_filter: buffer too small");
  return conv_tc_dgrad_filter(d, w, wt, (cudaStream_t)stream);
}

extern "C" int b2_pconv_wgrad(const B2ConvDesc* d, const void* x, const float* mask_in, const void* dy,
                              const float* ratio, float* dw, void* workspace, size_t ws_bytes, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  B2_REQUIRE(x && dy && dw, B2_E_BADARG, "pconv_wgrad: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(d, 2)) {
    B2_REQUIRE(ws_bytes >= conv_tc_workspace_bytes(d, 2), B2_E_WORKSPACE, "pconv_wgrad: workspace too small");
    return conv_tc_wgrad(d, x, mask_in, dy, ratio, dw, workspace, st);
  }
  B2_REQUIRE(!(d->flags & B2_CONV_X_CONCAT), B2_E_UNSUPPORTED,
             "pconv_wgrad: B2_CONV_X_CONCAT needs the bf16 tensor-core path (plain stride-1 layer, C %% 128 == 0)");
  return conv_ffma_wgrad(d, x, mask_in, dy, ratio, dw, st);
}
