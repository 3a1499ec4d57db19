// Shared helpers for the b2pose kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "b2pose.h"

void b2_set_error(const char* fmt, ...);

#define B2_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      b2_set_error(__VA_ARGS__);         \
      return (code);                     \
    }                                    \
  } while (0)

#define B2_LAUNCH_CHECK(what)                                                     \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      b2_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e__));  \
      return B2_E_LAUNCH;                                                         \
    }                                                                             \
  } while (0)

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float4 (16 B for float, 8 B for bf16); pointers must be aligned.
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// 8 consecutive bf16 <-> 8 floats (16 B)
__device__ __forceinline__ void load8(const bf16* p, float* f) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 issue two fp32 operations per lane and
// instruction, which matters in the issue-bound convolution epilogue)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// two bf16 in one 32-bit word -> packed fp32x2 (exact)
__device__ __forceinline__ uint64_t bf16x2_to_f32x2(uint32_t w) {
  return pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
// two fp32 -> two bf16 (round to nearest even) in one 32-bit word, `lo` in the low half
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// fp32 value of (window / (count + 1e-6)) * clamp(count, 0, 1) exactly as partial_conv.py:41-44
// evaluates it (fp32 add, fp32 divide, fp32 multiply; count is a small integer in fp32).
__device__ __forceinline__ float pconv_ratio(float window, float count) {
  float r = __fdiv_rn(window, __fadd_rn(count, 1e-6f));
  float mo = fminf(fmaxf(count, 0.f), 1.f);
  return __fmul_rn(r, mo);
}

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl may start (block
// scheduling, shared-memory / barrier / TMEM set-up) while its predecessor in the stream is still
// draining; pdl_wait() must precede the first access to global memory, pdl_trigger() lets the
// successor start early.  Both are no-ops for a normal launch.  Measured on the bench workload the
// attribute made the step 3.8 % SLOWER (20.58 vs 19.83 ms: early-resident successor blocks take
// occupancy from the predecessor's tail), so it is OFF unless B2POSE_PDL=1.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static inline bool b2_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2POSE_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = b2_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// SMs the grids of this library may fill: the device's count minus the SMs the caller reserved for kernels running
// beside ours (b2_set_sm_reserve: the NCCL kernels of an overlapped gradient all-reduce -- a persistent one-CTA-per-SM
// grid whose last CTAs have to wait for an SM that a collective holds takes twice as long).
extern int g_b2_sm_reserve;
static inline int b2_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int n = sms - g_b2_sm_reserve;
  return n < 1 ? 1 : n;
}

// ---- internal entry points (defined in the .cu files, dispatched from api.cu) ----
int conv_ffma_fprop(const B2ConvDesc* d, const void* x, const float* mask_in, const void* w, const float* bias,
                    void* y, float* mask_out, float* ratio_out, cudaStream_t st);
int conv_ffma_dgrad(const B2ConvDesc* d, const void* dy, const float* ratio, const void* w, const float* mask_in,
                    void* dx, cudaStream_t st);
int conv_ffma_wgrad(const B2ConvDesc* d, const void* x, const float* mask_in, const void* dy, const float* ratio,
                    float* dw, cudaStream_t st);

bool conv_tc_supported(const B2ConvDesc* d, int op);
size_t conv_tc_workspace_bytes(const B2ConvDesc* d, int op);
int conv_tc_fprop(const B2ConvDesc* d, const void* x, const float* mask_in, const void* w, const float* bias,
                  void* y, float* mask_out, float* ratio_out, float* bn_sums, void* workspace, cudaStream_t st);
int conv_tc_dgrad(const B2ConvDesc* d, const void* dy, const float* ratio, const void* w, const float* mask_in,
                  void* dx, void* workspace, cudaStream_t st);
int conv_tc_dgrad_filter(const B2ConvDesc* d, const void* w, void* wt, cudaStream_t st);
bool conv_tc_dgrad_needs_filter(const B2ConvDesc* d);
int conv_tc_wgrad(const B2ConvDesc* d, const void* x, const float* mask_in, const void* dy, const float* ratio,
                  float* dw, void* workspace, cudaStream_t st);

// bulk-async (TMA 1-D) staged BatchNorm stream kernels for bf16, C a power of two in [64, 2048] (bn_stream.cu)
bool bn_stream_eligible(int C, int dtype);
int bn_stream_stats(const void* y, int64_t rows, int C, float* partials, int totals, cudaStream_t st);
int bn_stream_apply_fin(const void* y, const void* residual, void* z, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, const float* row_mask, int relu, int64_t rows, int C,
                        int mode, const float* totals, float* running_mean, float* running_var, float momentum,
                        float eps, float* mean_out, float* invstd_out, uint8_t* gate_out, cudaStream_t st);
int bn_stream_apply(const void* y, const void* residual, void* z, const float* mean, const float* invstd,
                    const float* gamma, const float* beta, const float* row_mask, int relu, int64_t rows, int C,
                    cudaStream_t st);
int bn_stream_bwd_reduce(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const float* row_mask, int relu, float* partials,
                         int totals, const uint8_t* gate, int64_t rows, int C, cudaStream_t st);
int bn_stream_bwd_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, const float* gsum, const float* row_mask,
                        const float* row_scale, int relu, int training, void* dy, void* d_residual, float* dgamma,
                        float* dbeta, const uint8_t* gate, int64_t rows, int C, cudaStream_t st);
