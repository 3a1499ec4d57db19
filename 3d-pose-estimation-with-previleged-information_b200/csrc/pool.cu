// MaxPool2d(kernel 3, stride 2, padding 1) over NHWC activations, pooling the veil (validity
// mask) in the same launch: partial_depthnet.py:219-220 (x = maxpool(relu(x)); veil = maxpool(veil)).
// Forward stores the window position of the maximum (first maximum in scan order, like ATen)
// as one byte per element so the backward pass is a deterministic gather without atomics.
#include "b2_common.cuh"

namespace {

// ---- bf16, 8 channels (16 bytes) per thread -------------------------------------------------------
// One block per output (input) row, 32-bit index arithmetic, packed bf16x2 max / compare: the first version spent
// ~850 (fwd) / ~390 (bwd) instructions per 16-byte item on 64-bit divisions and scalar compare-select chains and
// ran at 1 TB/s (ncu: IPC 0.36, 25 % occupancy at 124 registers).
__device__ __forceinline__ __nv_bfloat162 as_bf162(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

__global__ void __launch_bounds__(256) maxpool_fwd8_kernel(const bf16* __restrict__ x, const float* __restrict__ veil_in,
                                                          bf16* __restrict__ y, uint8_t* __restrict__ argmax,
                                                          float* __restrict__ veil_out, int N, int H, int W, int C,
                                                          int Ho, int Wo) {
  const int C8 = C >> 3;
  const int per_row = Wo * C8;
  const uint32_t ninf2 = 0xff80ff80u;                       // (-inf, -inf) in bf16
  for (int row = blockIdx.x; row < N * Ho; row += gridDim.x) {
    const int n = row / Ho, oh = row - n * Ho;
    const bf16* xin = x + (size_t)n * H * W * C;
    for (int j = threadIdx.x; j < per_row; j += blockDim.x) {
      const int ow = j / C8, cg = j - ow * C8;
      uint4 v[9];
      bool ok[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const int ih = oh * 2 - 1 + r, iw = ow * 2 - 1 + s2, k = r * 3 + s2;
          ok[k] = ih >= 0 && ih < H && iw >= 0 && iw < W;
          if (ok[k]) v[k] = *reinterpret_cast<const uint4*>(xin + ((size_t)ih * W + iw) * C + cg * 8);
        }
      uint32_t best[4] = {ninf2, ninf2, ninf2, ninf2};
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (!ok[k]) continue;
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) best[q] = as_u32(__hmax2_nan(as_bf162(best[q]), as_bf162(w[q])));
      }
      // argmax = the first tap (in window order) that equals the maximum (unordered-equal, so a NaN maximum
      // also resolves): walk the taps backwards and let earlier ones overwrite
      uint32_t arg[4] = {0u, 0u, 0u, 0u};                   // two 16-bit indices per word
#pragma unroll
      for (int k = 8; k >= 0; --k) {
        if (!ok[k]) continue;
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        const uint32_t kk = (uint32_t)k * 0x00010001u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t m = __hequ2_mask(as_bf162(w[q]), as_bf162(best[q]));
          arg[q] = (arg[q] & ~m) | (kk & m);
        }
      }
      const size_t o = ((size_t)row * Wo + ow) * C + cg * 8;
      *reinterpret_cast<uint4*>(y + o) = make_uint4(best[0], best[1], best[2], best[3]);
      if (argmax) {
        uint2 a;       // bytes in channel order: word q holds channels 2q (low half) and 2q+1 (high half)
        a.x = (arg[0] & 0xffu) | ((arg[0] >> 16) << 8) | ((arg[1] & 0xffu) << 16) | ((arg[1] >> 16) << 24);
        a.y = (arg[2] & 0xffu) | ((arg[2] >> 16) << 8) | ((arg[3] & 0xffu) << 16) | ((arg[3] >> 16) << 24);
        *reinterpret_cast<uint2*>(argmax + o) = a;
      }
      if (veil_in && cg == 0) {
        float vmax = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k)
          if (ok[k]) vmax = fmaxf(vmax, veil_in[((size_t)n * H + (oh * 2 - 1 + k / 3)) * W + (ow * 2 - 1 + k % 3)]);
        veil_out[(size_t)row * Wo + ow] = vmax;
      }
    }
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd8_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ argmax,
                                                          bf16* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                                          int Wo) {
  const int C8 = C >> 3;
  const int per_row = W * C8;
  for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
    const int n = row / H, ih = row - n * H;
    // the (at most 2 x 2) output windows containing (ih, iw): oh in {(ih+1)/2, and (ih+1)/2 - 1 for odd ih}
    const int oh_hi = (ih + 1) >> 1, nh = (ih & 1) ? 2 : 1;
    for (int j = threadIdx.x; j < per_row; j += blockDim.x) {
      const int iw = j / C8, cg = j - iw * C8;
      const int ow_hi = (iw + 1) >> 1, nw = (iw & 1) ? 2 : 1;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int w2 = 0; w2 < 2; ++w2) {
          const int oh = oh_hi - u, ow = ow_hi - w2;
          if (!(u < nh && w2 < nw && oh >= 0 && oh < Ho && ow >= 0 && ow < Wo)) continue;
          const size_t op = (((size_t)n * Ho + oh) * Wo + ow) * C + cg * 8;
          const uint4 g = *reinterpret_cast<const uint4*>(dy + op);
          const uint2 a = *reinterpret_cast<const uint2*>(argmax + op);
          const uint32_t code = (uint32_t)((ih - (oh * 2 - 1)) * 3 + (iw - (ow * 2 - 1))) * 0x01010101u;
          const uint32_t m0 = __vcmpeq4(a.x, code), m1 = __vcmpeq4(a.y, code);      // 0xff per matching byte
          const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t mb = ((q < 4 ? m0 : m1) >> ((q & 3) * 8)) & 1u;
            const float gv = __uint_as_float((q & 1) ? (gw[q >> 1] & 0xffff0000u) : (gw[q >> 1] << 16));
            acc[q] += mb ? gv : 0.f;
          }
        }
      store8(dx + ((size_t)row * W + iw) * C + cg * 8, acc);
    }
  }
}

// The same gather with a 2 x 2 input block per thread: the four inputs (2a + di, 2b + dj) only ever look at the four output
// windows (a + u, b + v), so one thread loads those four (gradient, argmax) vectors once and serves all four inputs --
// 4 window loads per 4 inputs instead of 9 (1 / 2 / 2 / 4 by parity).
__global__ void __launch_bounds__(256) maxpool_bwd8_block_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ argmax,
                                                                bf16* __restrict__ dx, int N, int H, int W, int C, int Ho,
                                                                int Wo) {
  const int C8 = C >> 3, H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
  const int per_row = W2 * C8;
  for (int rp = blockIdx.x; rp < N * H2; rp += gridDim.x) {
    const int n = rp / H2, a = rp - n * H2;
    for (int j = threadIdx.x; j < per_row; j += blockDim.x) {
      const int b = j / C8, cg = j - b * C8;
      uint4 g[2][2];
      uint2 am[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int oh = a + u, ow = b + v;
          g[u][v] = make_uint4(0u, 0u, 0u, 0u);
          am[u][v] = make_uint2(0xffffffffu, 0xffffffffu);              // matches no window position
          if (oh < Ho && ow < Wo) {
            const size_t op = (((size_t)n * Ho + oh) * Wo + ow) * C + cg * 8;
            g[u][v] = *reinterpret_cast<const uint4*>(dy + op);
            am[u][v] = *reinterpret_cast<const uint2*>(argmax + op);
          }
        }
#pragma unroll
      for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj) {
          const int ih = 2 * a + di, iw = 2 * b + dj;
          if (ih >= H || iw >= W) continue;
          float acc[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll
          for (int u = 0; u <= di; ++u)
#pragma unroll
            for (int v = 0; v <= dj; ++v) {
              // position of (ih, iw) inside window (a + u, b + v): row di - 2u + 1, column dj - 2v + 1
              const uint32_t code = (uint32_t)((di - 2 * u + 1) * 3 + (dj - 2 * v + 1)) * 0x01010101u;
              const uint32_t m0 = __vcmpeq4(am[u][v].x, code), m1 = __vcmpeq4(am[u][v].y, code);
              const uint32_t gw[4] = {g[u][v].x, g[u][v].y, g[u][v].z, g[u][v].w};
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const uint32_t mb = ((q < 4 ? m0 : m1) >> ((q & 3) * 8)) & 1u;
                const float gv = __uint_as_float((q & 1) ? (gw[q >> 1] & 0xffff0000u) : (gw[q >> 1] << 16));
                acc[q] += mb ? gv : 0.f;
              }
            }
          store8(dx + (((size_t)n * H + ih) * W + iw) * C + cg * 8, acc);
        }
    }
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

inline int row_grid(long long rows) {
  long long cap = (long long)b2_num_sms() * 32;
  return (int)(rows < 1 ? 1 : (rows > cap ? cap : rows));
}

}  // namespace

extern "C" int b2_maxpool3x3s2_fwd(const void* x, const float* veil_in, void* y, uint8_t* argmax, float* veil_out,
                                   int32_t N, int32_t H, int32_t W, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, B2_E_BADARG, "maxpool_fwd: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "maxpool_fwd: C=%d is not a multiple of 4", C);
  B2_REQUIRE((veil_in == nullptr) == (veil_out == nullptr), B2_E_BADARG, "maxpool_fwd: veil_in/veil_out mismatch");
  B2_REQUIRE((long long)N * H < (1LL << 31) && (long long)W * C < (1LL << 31), B2_E_UNSUPPORTED, "maxpool_fwd: tensor too large");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C >> 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B2_F32)
    maxpool_fwd_kernel<float><<<pool_grid(total), 256, 0, st>>>((const float*)x, veil_in, (float*)y, argmax, veil_out,
                                                              N, H, W, C, Ho, Wo);
  else if ((C & 7) == 0)
    maxpool_fwd8_kernel<<<row_grid((long long)N * Ho), 256, 0, st>>>((const bf16*)x, veil_in, (bf16*)y, argmax, veil_out,
                                                                    N, H, W, C, Ho, Wo);
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
    {
      static const bool per_pixel = getenv("B2POSE_POOL_BWD_PIXEL") && atoi(getenv("B2POSE_POOL_BWD_PIXEL")) != 0;   // A/B switch
      if (per_pixel)
        maxpool_bwd8_kernel<<<row_grid((long long)N * H), 256, 0, st>>>((const bf16*)dy, argmax, (bf16*)dx, N, H, W, C, Ho, Wo);
      else
        maxpool_bwd8_block_kernel<<<row_grid((long long)N * ((H + 1) / 2)), 256, 0, st>>>((const bf16*)dy, argmax, (bf16*)dx,
                                                                                        N, H, W, C, Ho, Wo);
    }
  else
    maxpool_bwd_kernel<bf16><<<pool_grid(total), 256, 0, st>>>((const bf16*)dy, argmax, (bf16*)dx, N, H, W, C, Ho, Wo);
  B2_LAUNCH_CHECK("maxpool_bwd");
  return B2_OK;
}
