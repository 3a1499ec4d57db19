// MaxPool2d(kernel 3, stride 2, padding 1) over NHWC activations, pooling the veil (validity
// mask) in the same launch: partial_depthnet.py:219-220 (x = maxpool(relu(x)); veil = maxpool(veil)).
// Forward stores the window position of the maximum (first maximum in scan order, like ATen)
// as one byte per element so the backward pass is a deterministic gather without atomics.
#include "b2_common.cuh"

namespace {

// ---- bf16, 8 channels (16 bytes) per thread -------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_fwd8_kernel(const bf16* __restrict__ x, const float* __restrict__ veil_in,
                                                          bf16* __restrict__ y, uint8_t* __restrict__ argmax,
                                                          float* __restrict__ veil_out, int N, int H, int W, int C,
                                                          int Ho, int Wo) {
  const int C8 = C >> 3;
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C8);
    const long long pix = i / C8;
    const int ow = (int)(pix % Wo);
    const long long t = pix / Wo;
    const int oh = (int)(t % Ho), n = (int)(t / Ho);
    float v[9][8];
    bool ok[9];
    // issue all (up to 9) 16-byte loads first
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ih = oh * 2 - 1 + r, iw = ow * 2 - 1 + s;
        ok[r * 3 + s] = ih >= 0 && ih < H && iw >= 0 && iw < W;
        if (ok[r * 3 + s]) load8(x + (((long long)n * H + ih) * W + iw) * C + cg * 8, v[r * 3 + s]);
      }
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
    float vmax = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (!ok[k]) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[k][j] > best[j] || v[k][j] != v[k][j]) { best[j] = v[k][j]; arg[j] = k; }
      if (veil_in && cg == 0) {
        const int ih = oh * 2 - 1 + k / 3, iw = ow * 2 - 1 + k % 3;
        vmax = fmaxf(vmax, veil_in[((long long)n * H + ih) * W + iw]);
      }
    }
    store8(y + pix * C + cg * 8, best);
    if (argmax) {
      uint2 a;
      a.x = (uint32_t)arg[0] | ((uint32_t)arg[1] << 8) | ((uint32_t)arg[2] << 16) | ((uint32_t)arg[3] << 24);
      a.y = (uint32_t)arg[4] | ((uint32_t)arg[5] << 8) | ((uint32_t)arg[6] << 16) | ((uint32_t)arg[7] << 24);
      *reinterpret_cast<uint2*>(argmax + pix * C + cg * 8) = a;
    }
    if (veil_out && cg == 0) veil_out[pix] = vmax;
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd8_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ argmax,
                                                          bf16* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                                          int Wo) {
  const int C8 = C >> 3;
  const long long total = (long long)N * H * W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C8);
    const long long pix = i / C8;
    const int iw = (int)(pix % W);
    const long long t = pix / W;
    const int ih = (int)(t % H), n = (int)(t / H);
    // the (at most 2 x 2) output windows containing (ih, iw): oh in {(ih+1)/2, and (ih+1)/2 - 1 for odd ih}
    const int oh_hi = (ih + 1) >> 1, ow_hi = (iw + 1) >> 1;
    const int nh = (ih & 1) ? 2 : 1, nw = (iw & 1) ? 2 : 1;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    float g[4][8];
    uint2 a[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int w2 = 0; w2 < 2; ++w2) {
        const int oh = oh_hi - u, ow = ow_hi - w2;
        const int k = u * 2 + w2;
        ok[k] = u < nh && w2 < nw && oh >= 0 && oh < Ho && ow >= 0 && ow < Wo;
        if (ok[k]) {
          const long long op = (((long long)n * Ho + oh) * Wo + ow) * C + cg * 8;
          load8(dy + op, g[k]);
          a[k] = *reinterpret_cast<const uint2*>(argmax + op);
        }
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int w2 = 0; w2 < 2; ++w2) {
        const int k = u * 2 + w2;
        if (!ok[k]) continue;
        const int oh = oh_hi - u, ow = ow_hi - w2;
        const uint32_t code = (uint32_t)((ih - (oh * 2 - 1)) * 3 + (iw - (ow * 2 - 1)));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t aj = ((j < 4 ? a[k].x : a[k].y) >> ((j & 3) * 8)) & 0xffu;
          if (aj == code) acc[j] += g[k][j];
        }
      }
    store8(dx + pix * C + cg * 8, acc);
  }
}

template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, const float* __restrict__ veil_in, T* __restrict__ y,
                                   uint8_t* __restrict__ argmax, float* __restrict__ veil_out, int N, int H, int W,
                                   int C, int Ho, int Wo) {
  const int C4 = C >> 2;
  const long long total = (long long)N * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cg = (int)(i % C4);
    long long pix = i / C4;
    int ow = (int)(pix % Wo);
    long long t = pix / Wo;
    int oh = (int)(t % Ho), n = (int)(t / Ho);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {0, 0, 0, 0};
    float vmax = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        long long ip = ((long long)n * H + ih) * W + iw;
        float4 f = load4(x + ip * C + cg * 4);
        float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (v[j] > best[j] || v[j] != v[j]) { best[j] = v[j]; arg[j] = r * 3 + s; }
        if (veil_in && cg == 0) vmax = fmaxf(vmax, veil_in[ip]);
      }
    }
    store4(y + pix * C + cg * 4, make_float4(best[0], best[1], best[2], best[3]));
    if (argmax) {
      uchar4 a = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
      *reinterpret_cast<uchar4*>(argmax + pix * C + cg * 4) = a;
    }
    if (veil_out && cg == 0) veil_out[pix] = vmax;
  }
}

template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ argmax, T* __restrict__ dx,
                                   int N, int H, int W, int C, int Ho, int Wo) {
  const int C4 = C >> 2;
  const long long total = (long long)N * H * W * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cg = (int)(i % C4);
    long long pix = i / C4;
    int iw = (int)(pix % W);
    long long t = pix / W;
    int ih = (int)(t % H), n = (int)(t / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // output windows that contain (ih, iw): oh*2-1+r == ih, r in 0..2
    for (int r = 0; r < 3; ++r) {
      int th = ih + 1 - r;
      if (th < 0 || (th & 1)) continue;
      int oh = th >> 1;
      if (oh >= Ho) continue;
      for (int s = 0; s < 3; ++s) {
        int tw = iw + 1 - s;
        if (tw < 0 || (tw & 1)) continue;
        int ow = tw >> 1;
        if (ow >= Wo) continue;
        long long op = ((long long)n * Ho + oh) * Wo + ow;
        uchar4 a = *reinterpret_cast<const uchar4*>(argmax + op * C + cg * 4);
        float4 g = load4(dy + op * C + cg * 4);
        int code = r * 3 + s;
        if (a.x == code) acc[0] += g.x;
        if (a.y == code) acc[1] += g.y;
        if (a.z == code) acc[2] += g.z;
        if (a.w == code) acc[3] += g.w;
      }
    }
    store4(dx + pix * C + cg * 4, make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
}

inline int pool_grid(long long total) {
  long long want = (total + 255) / 256, cap = (long long)b2_num_sms() * 16;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" int b2_maxpool3x3s2_fwd(const void* x, const float* veil_in, void* y, uint8_t* argmax, float* veil_out,
                                   int32_t N, int32_t H, int32_t W, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, B2_E_BADARG, "maxpool_fwd: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "maxpool_fwd: C=%d is not a multiple of 4", C);
  B2_REQUIRE((veil_in == nullptr) == (veil_out == nullptr), B2_E_BADARG, "maxpool_fwd: veil_in/veil_out mismatch");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C >> 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2_F32)
    maxpool_fwd_kernel<float><<<pool_grid(total), 256, 0, st>>>((const float*)x, veil_in, (float*)y, argmax, veil_out,
                                                              N, H, W, C, Ho, Wo);
  else if ((C & 7) == 0)
    maxpool_fwd8_kernel<<<pool_grid(total / 2), 256, 0, st>>>((const bf16*)x, veil_in, (bf16*)y, argmax, veil_out, N, H,
                                                             W, C, Ho, Wo);
  else
    maxpool_fwd_kernel<bf16><<<pool_grid(total), 256, 0, st>>>((const bf16*)x, veil_in, (bf16*)y, argmax, veil_out, N,
                                                             H, W, C, Ho, Wo);
  B2_LAUNCH_CHECK("maxpool_fwd");
  return B2_OK;
}

extern "C" int b2_maxpool3x3s2_bwd(const void* dy, const uint8_t* argmax, void* dx, int32_t N, int32_t H, int32_t W,
                                   int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(dy && argmax && dx && N > 0 && H > 0 && W > 0 && C > 0, B2_E_BADARG, "maxpool_bwd: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "maxpool_bwd: C=%d is not a multiple of 4", C);
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * H * W * (C >> 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2_F32)
    maxpool_bwd_kernel<float><<<pool_grid(total), 256, 0, st>>>((const float*)dy, argmax, (float*)dx, N, H, W, C, Ho, Wo);
  else if ((C & 7) == 0)
    maxpool_bwd8_kernel<<<pool_grid(total / 2), 256, 0, st>>>((const bf16*)dy, argmax, (bf16*)dx, N, H, W, C, Ho, Wo);
  else
    maxpool_bwd_kernel<bf16><<<pool_grid(total), 256, 0, st>>>((const bf16*)dy, argmax, (bf16*)dx, N, H, W, C, Ho, Wo);
  B2_LAUNCH_CHECK("maxpool_bwd");
  return B2_OK;
}
