// HBM-bound helper kernels: row scaling, column sums, veil, layout changes, casts, the
// PartialConv mask algebra on its own, the depth unprojection and the fused clip+Adam step.
// All are grid-stride, 16-byte vectorised where the shape allows; grids are sized as a
// multiple of the SM count.
#include "b2_common.cuh"

namespace {

inline int grid_for(long long work_items, int block, int per_sm = 8) {
  long long want = (work_items + block - 1) / block;
  long long cap = (long long)b2_num_sms() * per_sm;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// ---------------------------------------------------------------- scale_rows
template <typename T>
__global__ void scale_rows_kernel(const T* __restrict__ in, const float* __restrict__ scale, T* __restrict__ out,
                                  long long rows, int C4) {
  long long total = rows * C4;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    long long r = v / C4;
    float s = scale[r];
    float4 f = load4(in + v * 4);
    f.x *= s; f.y *= s; f.z *= s; f.w *= s;
    store4(out + v * 4, f);
  }
}
template <typename T>
__global__ void scale_rows_scalar_kernel(const T* __restrict__ in, const float* __restrict__ scale,
                                         T* __restrict__ out, long long rows, int C) {
  long long total = rows * C;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x)
    out[v] = from_f<T>(to_f(in[v]) * scale[v / C]);
}

// ---------------------------------------------------------------- col_sum
// block (TX, TY): tx = group of 4 channels, ty strides over rows; fp32 per-thread partials over a
// short run of rows, shared-memory reduce over ty, one fp32 atomic per channel per block.
template <typename T>
__global__ void col_sum_kernel(const T* __restrict__ in, const float* __restrict__ roww, float* __restrict__ sums,
                               long long rows, int C) {
  extern __shared__ float4 red[];
  const int C4 = C >> 2;
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < C4) {
    for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < rows;
         r += (long long)gridDim.x * blockDim.y) {
      float4 f = load4(in + r * C + cg * 4);
      float wgt = roww ? roww[r] : 1.f;
      acc.x += f.x * wgt; acc.y += f.y * wgt; acc.z += f.z * wgt; acc.w += f.w * wgt;
    }
  }
  red[threadIdx.y * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && cg < C4) {
    for (int i = 1; i < blockDim.y; ++i) {
      float4 o = red[i * blockDim.x + threadIdx.x];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    atomicAdd(sums + cg * 4 + 0, acc.x);
    atomicAdd(sums + cg * 4 + 1, acc.y);
    atomicAdd(sums + cg * 4 + 2, acc.z);
    atomicAdd(sums + cg * 4 + 3, acc.w);
  }
}
template <typename T>
__global__ void col_sum_scalar_kernel(const T* __restrict__ in, const float* __restrict__ roww,
                                      float* __restrict__ sums, long long rows, int C) {
  // one block per channel (only used for C % 4 != 0, i.e. tiny tensors)
  __shared__ float part[32];
  int c = blockIdx.x;
  float acc = 0.f;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) acc += to_f(in[r * C + c]) * (roww ? roww[r] : 1.f);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(sums + c, v);
  }
}

// ---------------------------------------------------------------- veil / casts / layouts
template <typename T>
__global__ void veil_kernel(const T* __restrict__ depth, float* __restrict__ veil, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    veil[i] = (to_f(depth[i]) != 0.f) ? 1.f : 0.f;
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    store4(dst + i * 4, load4(src + i * 4));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = __float2bfloat16_rn(src[n4 * 4 + threadIdx.x]);
}

__global__ void cast_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    store4(dst + i * 4, load4(src + i * 4));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = __bfloat162float(src[n4 * 4 + threadIdx.x]);
}

// NCHW fp32 -> NHWC T through a shared-memory transpose of [C-chunk x 32 pixels]
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int C, long long HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? in[((long long)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i;
    int c = c0 + threadIdx.x;
    if (c < C && p < HW) out[((long long)n * HW + p) * C + c] = from_f<T>(tile[threadIdx.x][i]);
  }
}
// few channels (network inputs, C <= 4): one thread per pixel, plane reads and pixel writes both coalesced
template <typename T, int C>
__global__ void nchw_to_nhwc_small_kernel(const float* __restrict__ in, T* __restrict__ out, long long HW, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
#pragma unroll
    for (int c = 0; c < C; ++c) out[i * C + c] = from_f<T>(in[(n * C + c) * HW + p]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int C, long long HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i;
    int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? to_f(in[((long long)n * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long p = p0 + threadIdx.x;
    if (c < C && p < HW) out[((long long)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------- mask algebra only
__global__ void mask_update_kernel(const float* __restrict__ mask_in, float* __restrict__ mask_out,
                                   float* __restrict__ ratio_out, int N, int H, int W, int R, int S, int stride,
                                   int pad, int dil, int Ho, int Wo) {
  long long total = (long long)N * Ho * Wo;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    int ow = (int)(m % Wo);
    long long t = m / Wo;
    int oh = (int)(t % Ho), n = (int)(t / Ho);
    float cnt = 0.f;
    for (int r = 0; r < R; ++r) {
      int ih = oh * stride - pad + r * dil;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < S; ++s) {
        int iw = ow * stride - pad + s * dil;
        if (iw < 0 || iw >= W) continue;
        cnt += mask_in[((long long)n * H + ih) * W + iw];
      }
    }
    if (mask_out) mask_out[m] = fminf(fmaxf(cnt, 0.f), 1.f);
    if (ratio_out) ratio_out[m] = pconv_ratio((float)(R * S), cnt);
  }
}

// Box-window version for the dense window of the stems (dil == 1): one block = kMuRows output rows of one image.  The
// input rows are staged in shared memory, summed horizontally once per (input row, output column) and then vertically,
// instead of R*S global loads per output.  Sums of {0, 1} values are exact in any order, so the results are bit-identical.
constexpr int kMuRows = 8;
__global__ void __launch_bounds__(256) mask_update_box_kernel(const float* __restrict__ mask_in, float* __restrict__ mask_out,
                                                              float* __restrict__ ratio_out, int H, int W, int R, int S,
                                                              int stride, int pad, int Ho, int Wo) {
  extern __shared__ float mu_sm[];
  const int in_rows = (kMuRows - 1) * stride + R;
  float* rows = mu_sm;                       // [in_rows][W]
  float* hsum = mu_sm + (size_t)in_rows * W; // [in_rows][Wo]
  const int n = blockIdx.y, oh0 = blockIdx.x * kMuRows, ih0 = oh0 * stride - pad;
  const float* src = mask_in + (long long)n * H * W;
  for (int i = threadIdx.x; i < in_rows * W; i += blockDim.x) {
    const int r = i / W, w = i - r * W, ih = ih0 + r;
    rows[i] = (ih >= 0 && ih < H) ? src[(long long)ih * W + w] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < in_rows * Wo; i += blockDim.x) {
    const int r = i / Wo, ow = i - r * Wo, iw0 = ow * stride - pad;
    float a = 0.f;
    for (int s2 = 0; s2 < S; ++s2) {
      const int iw = iw0 + s2;
      if (iw >= 0 && iw < W) a += rows[r * W + iw];
    }
    hsum[i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMuRows * Wo; i += blockDim.x) {
    const int orow = i / Wo, ow = i - orow * Wo, oh = oh0 + orow;
    if (oh >= Ho) continue;
    float cnt = 0.f;
    for (int r = 0; r < R; ++r) cnt += hsum[(orow * stride + r) * Wo + ow];
    const long long m = ((long long)n * Ho + oh) * Wo + ow;
    if (mask_out) mask_out[m] = fminf(fmaxf(cnt, 0.f), 1.f);
    if (ratio_out) ratio_out[m] = pconv_ratio((float)(R * S), cnt);
  }
}

// ---------------------------------------------------------------- unprojection
// Vector path (W % 4 == 0, fewer than 2^30 vectors): 32-bit index arithmetic (the 64-bit divisions of the generic kernel below cap it at 58 % of the HBM
// rate) and two independent 16-byte loads in flight per thread.
__global__ void __launch_bounds__(256) unproject_vec_kernel(const float* __restrict__ img, float* __restrict__ out, int total,
                                                            int H, int W4, float k00, float k01, float k10, float k11,
                                                            float cx, float cy) {
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
    const int idx[2] = {i0, i0 + stride};
    float4 f[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (idx[u] < total) f[u] = __ldcs(reinterpret_cast<const float4*>(img) + idx[u]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (idx[u] >= total) continue;
      const int row = idx[u] / W4, w0 = (idx[u] - row * W4) * 4;
      const float dv = (float)(row % H) - cy;
      float q[4] = {f[u].x, f[u].y, f[u].z, f[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float du = (float)(w0 + j) - cx;
        const float xn = __fadd_rn(__fmul_rn(du, k00), __fmul_rn(dv, k01));
        const float yn = __fadd_rn(__fmul_rn(du, k10), __fmul_rn(dv, k11));
        const float ss = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xn, xn), __fmul_rn(yn, yn)), 1.f), 1.f);
        q[j] = __fdiv_rn(q[j], __fsqrt_rn(ss));
      }
      __stcs(reinterpret_cast<float4*>(out) + idx[u], make_float4(q[0], q[1], q[2], q[3]));
    }
  }
}

// out = img / sqrt(xn^2 + yn^2 + 1 + 1); 4 pixels per thread when W % 4 == 0.
__global__ void unproject_kernel(const float* __restrict__ img, float* __restrict__ out, long long n_img, int H,
                                 int W, float k00, float k01, float k10, float k11, float cx, float cy) {
  const int W4 = (W + 3) >> 2;
  const bool vec = (W & 3) == 0;
  long long total = n_img * H * W4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int w0 = (int)(i % W4) * 4;
    long long row = i / W4;
    int v = (int)(row % H);
    float dv = (float)v - cy;
    float q[4];
    if (vec) {
      float4 f = *reinterpret_cast<const float4*>(img + row * W + w0);
      q[0] = f.x; q[1] = f.y; q[2] = f.z; q[3] = f.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!vec) {
        if (w0 + j >= W) break;
        q[j] = img[row * W + w0 + j];
      }
      float du = (float)(w0 + j) - cx;
      float xn = __fadd_rn(__fmul_rn(du, k00), __fmul_rn(dv, k01));
      float yn = __fadd_rn(__fmul_rn(du, k10), __fmul_rn(dv, k11));
      float ss = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xn, xn), __fmul_rn(yn, yn)), 1.f), 1.f);
      q[j] = __fdiv_rn(q[j], __fsqrt_rn(ss));
      if (!vec) out[row * W + w0 + j] = q[j];
    }
    if (vec) *reinterpret_cast<float4*>(out + row * W + w0) = make_float4(q[0], q[1], q[2], q[3]);
  }
}

// ---------------------------------------------------------------- optimizer
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  __shared__ float part[32];
  float acc = 0.f;
  long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 f = load4(g + i * 4);
    acc += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    float t = g[n4 * 4 + threadIdx.x];
    acc += t * t;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, (double)v);
  }
}

struct AdamP {
  float lr, b1, b2, eps, wd, max_norm, inv_scale;
  float bc1, bc2_sqrt;  // 1-b1^t, sqrt(1-b2^t)
};

__device__ __forceinline__ void adam_one(float& w, float g, float& m, float& v, const AdamP& p, float coef) {
  g = g * coef + p.wd * w;                       // clip_grad_norm_ scaling, then L2 weight decay
  m = p.b1 * m + (1.f - p.b1) * g;
  v = p.b2 * v + (1.f - p.b2) * g * g;
  float denom = sqrtf(v) / p.bc2_sqrt + p.eps;
  w -= (p.lr / p.bc1) * (m / denom);
}

__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, bf16* __restrict__ w16, long long n, AdamP p,
                            const double* __restrict__ sumsq, const float* __restrict__ dev_hyper) {
  if (dev_hyper) { p.lr = dev_hyper[0]; p.bc1 = dev_hyper[1]; p.bc2_sqrt = dev_hyper[2]; }
  float coef = p.inv_scale;
  if (sumsq) {
    double ss = *sumsq;
    if (!isfinite(ss)) return;                   // the reference's inf-skip (depth_train.py:435-438)
    float total = (float)sqrt(ss) * p.inv_scale;
    if (p.max_norm > 0.f) {
      float c = p.max_norm / (total + 1e-6f);    // torch.nn.utils.clip_grad_norm_
      coef *= fminf(c, 1.f);
    }
  }
  long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 wv = load4(w + i * 4), gv = load4(g + i * 4), mv = load4(m + i * 4), vv = load4(v + i * 4);
    adam_one(wv.x, gv.x, mv.x, vv.x, p, coef);
    adam_one(wv.y, gv.y, mv.y, vv.y, p, coef);
    adam_one(wv.z, gv.z, mv.z, vv.z, p, coef);
    adam_one(wv.w, gv.w, mv.w, vv.w, p, coef);
    store4(w + i * 4, wv); store4(m + i * 4, mv); store4(v + i * 4, vv);
    if (w16) store4(w16 + i * 4, wv);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    long long i = n4 * 4 + threadIdx.x;
    float wv = w[i], mv = m[i], vv = v[i];
    adam_one(wv, g[i], mv, vv, p, coef);
    w[i] = wv; m[i] = mv; v[i] = vv;
    if (w16) w16[i] = __float2bfloat16_rn(wv);
  }
}

}  // namespace

extern "C" int b2_scale_rows(const void* in, const float* scale, void* out, int64_t rows, int32_t C,
                             int32_t dtype, void* stream) {
  B2_REQUIRE(in && scale && out && rows >= 0 && C > 0, B2_E_BADARG, "scale_rows: bad argument");
  if (rows == 0) return B2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if ((C & 3) == 0) {
    int grid = grid_for(rows * (C >> 2), 256);
    if (dtype == B2_F32) scale_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)in, scale, (float*)out, rows, C >> 2);
    else scale_rows_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)in, scale, (bf16*)out, rows, C >> 2);
  } else {
    int grid = grid_for(rows * C, 256);
    if (dtype == B2_F32) scale_rows_scalar_kernel<float><<<grid, 256, 0, st>>>((const float*)in, scale, (float*)out, rows, C);
    else scale_rows_scalar_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)in, scale, (bf16*)out, rows, C);
  }
  B2_LAUNCH_CHECK("scale_rows");
  return B2_OK;
}

extern "C" int b2_col_sum(const void* in, const float* row_weight, float* sums, int64_t rows, int32_t C,
                          int32_t dtype, void* stream) {
  B2_REQUIRE(in && sums && rows >= 0 && C > 0, B2_E_BADARG, "col_sum: bad argument");
  if (rows == 0) return B2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if ((C & 3) == 0) {
    int C4 = C >> 2;
    int tx = C4 < 64 ? C4 : 64;
    int ty = 256 / tx;
    if (ty < 1) ty = 1;
    dim3 block(tx, ty);
    int gy = (C4 + tx - 1) / tx;
    long long gx = (rows + ty * 16 - 1) / (ty * 16);
    long long cap = (long long)b2_num_sms() * 8 / gy;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, gy);
    size_t sh = sizeof(float4) * tx * ty;
    if (dtype == B2_F32) col_sum_kernel<float><<<grid, block, sh, st>>>((const float*)in, row_weight, sums, rows, C);
    else col_sum_kernel<bf16><<<grid, block, sh, st>>>((const bf16*)in, row_weight, sums, rows, C);
  } else {
    if (dtype == B2_F32) col_sum_scalar_kernel<float><<<C, 256, 0, st>>>((const float*)in, row_weight, sums, rows, C);
    else col_sum_scalar_kernel<bf16><<<C, 256, 0, st>>>((const bf16*)in, row_weight, sums, rows, C);
  }
  B2_LAUNCH_CHECK("col_sum");
  return B2_OK;
}

extern "C" int b2_veil_from_depth(const void* depth, float* veil, int64_t n, int32_t dtype, void* stream) {
  B2_REQUIRE(depth && veil && n >= 0, B2_E_BADARG, "veil_from_depth: bad argument");
  if (n == 0) return B2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n, 256);
  if (dtype == B2_F32) veil_kernel<float><<<grid, 256, 0, st>>>((const float*)depth, veil, n);
  else veil_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)depth, veil, n);
  B2_LAUNCH_CHECK("veil_from_depth");
  return B2_OK;
}

extern "C" int b2_nchw_to_nhwc(const float* in, void* out, int32_t N, int32_t C, int32_t H, int32_t W,
                               int32_t dtype, void* stream) {
  B2_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0, B2_E_BADARG, "nchw_to_nhwc: bad argument");
  B2_REQUIRE(N <= 65535, B2_E_UNSUPPORTED, "nchw_to_nhwc: N > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  long long HW = (long long)H * W;
  if (C <= 4) {
    const long long total = HW * N;
    const int g = grid_for(total, 256);
#define SMALL_C(CC)                                                                                            \
  if (dtype == B2_F32) nchw_to_nhwc_small_kernel<float, CC><<<g, 256, 0, st>>>(in, (float*)out, HW, total);  \
  else nchw_to_nhwc_small_kernel<bf16, CC><<<g, 256, 0, st>>>(in, (bf16*)out, HW, total);
    if (C == 1) { SMALL_C(1) } else if (C == 2) { SMALL_C(2) } else if (C == 3) { SMALL_C(3) } else { SMALL_C(4) }
#undef SMALL_C
    B2_LAUNCH_CHECK("nchw_to_nhwc");
    return B2_OK;
  }
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  if (dtype == B2_F32) nchw_to_nhwc_kernel<float><<<grid, block, 0, st>>>(in, (float*)out, C, HW);
  else nchw_to_nhwc_kernel<bf16><<<grid, block, 0, st>>>(in, (bf16*)out, C, HW);
  B2_LAUNCH_CHECK("nchw_to_nhwc");
  return B2_OK;
}

extern "C" int b2_nhwc_to_nchw(const void* in, float* out, int32_t N, int32_t C, int32_t H, int32_t W,
                               int32_t dtype, void* stream) {
  B2_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0, B2_E_BADARG, "nhwc_to_nchw: bad argument");
  B2_REQUIRE(N <= 65535, B2_E_UNSUPPORTED, "nhwc_to_nchw: N > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  if (dtype == B2_F32) nhwc_to_nchw_kernel<float><<<grid, block, 0, st>>>((const float*)in, out, C, HW);
  else nhwc_to_nchw_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)in, out, C, HW);
  B2_LAUNCH_CHECK("nhwc_to_nchw");
  return B2_OK;
}

extern "C" int b2_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  B2_REQUIRE(src && dst && n >= 0, B2_E_BADARG, "cast: bad argument");
  if (n == 0) return B2_OK;
  B2_REQUIRE((((uintptr_t)src) & 15) == 0 && (((uintptr_t)dst) & 7) == 0, B2_E_BADARG, "cast: unaligned pointer");
  cast_bf16_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  B2_LAUNCH_CHECK("cast_f32_to_bf16");
  return B2_OK;
}

extern "C" int b2_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  B2_REQUIRE(src && dst && n >= 0, B2_E_BADARG, "cast: bad argument");
  if (n == 0) return B2_OK;
  B2_REQUIRE((((uintptr_t)dst) & 15) == 0 && (((uintptr_t)src) & 7) == 0, B2_E_BADARG, "cast: unaligned pointer");
  cast_f32_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  B2_LAUNCH_CHECK("cast_bf16_to_f32");
  return B2_OK;
}

extern "C" int b2_pconv_mask_update(const B2ConvDesc* d, const float* mask_in, float* mask_out, float* ratio_out,
                                    void* stream) {
  B2_REQUIRE(d && mask_in && (mask_out || ratio_out), B2_E_BADARG, "mask_update: bad argument");
  long long total = (long long)d->N * d->Ho * d->Wo;
  const size_t box_sh = ((size_t)((kMuRows - 1) * d->stride + d->R) * (d->W + d->Wo)) * sizeof(float);
  if (d->dil == 1 && d->R * d->S >= 25 && box_sh <= 160 * 1024) {
    if (box_sh > 48 * 1024)
      cudaFuncSetAttribute(mask_update_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)box_sh);
    mask_update_box_kernel<<<dim3((d->Ho + kMuRows - 1) / kMuRows, d->N), 256, box_sh, (cudaStream_t)stream>>>(
        mask_in, mask_out, ratio_out, d->H, d->W, d->R, d->S, d->stride, d->pad, d->Ho, d->Wo);
    B2_LAUNCH_CHECK("mask_update");
    return B2_OK;
  }
  mask_update_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      mask_in, mask_out, ratio_out, d->N, d->H, d->W, d->R, d->S, d->stride, d->pad, d->dil, d->Ho, d->Wo);
  B2_LAUNCH_CHECK("mask_update");
  return B2_OK;
}

extern "C" int b2_unproject_depth(const float* img, float* out, int32_t n_img, int32_t H, int32_t W,
                                  const float kinv[4], const float c[2], void* stream) {
  B2_REQUIRE(img && out && kinv && c && n_img > 0 && H > 0 && W > 0, B2_E_BADARG, "unproject_depth: bad argument");
  long long total = (long long)n_img * H * ((W + 3) / 4);
  if ((W & 3) == 0 && total < (1LL << 30) && (((uintptr_t)img | (uintptr_t)out) & 15) == 0) {
    unproject_vec_kernel<<<grid_for((total + 1) / 2, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        img, out, (int)total, H, W / 4, kinv[0], kinv[1], kinv[2], kinv[3], c[0], c[1]);
    B2_LAUNCH_CHECK("unproject_depth");
    return B2_OK;
  }
  unproject_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, out, n_img, H, W, kinv[0], kinv[1],
                                                                         kinv[2], kinv[3], c[0], c[1]);
  B2_LAUNCH_CHECK("unproject_depth");
  return B2_OK;
}

extern "C" int b2_grad_sumsq(const float* g, int64_t n, double* sumsq, void* stream) {
  B2_REQUIRE(g && sumsq && n >= 0, B2_E_BADARG, "grad_sumsq: bad argument");
  if (n == 0) return B2_OK;
  sumsq_kernel<<<grid_for((n + 3) / 4, 256, 4), 256, 0, (cudaStream_t)stream>>>(g, n, sumsq);
  B2_LAUNCH_CHECK("grad_sumsq");
  return B2_OK;
}

extern "C" int b2_adam_step(float* w, const float* g, float* m, float* v, void* w16, int64_t n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int32_t step,
                            const double* sumsq, float max_norm, float inv_scale, const float* dev_hyper,
                            void* stream) {
  B2_REQUIRE(w && g && m && v && n >= 0 && (step >= 1 || dev_hyper), B2_E_BADARG, "adam_step: bad argument");
  if (step < 1) step = 1;
  if (n == 0) return B2_OK;
  AdamP p;
  p.lr = lr; p.b1 = beta1; p.b2 = beta2; p.eps = eps; p.wd = weight_decay; p.max_norm = max_norm;
  p.inv_scale = inv_scale;
  p.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  p.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(w, g, m, v, (bf16*)w16, n, p, sumsq, dev_hyper);
  B2_LAUNCH_CHECK("adam_step");
  return B2_OK;
}
