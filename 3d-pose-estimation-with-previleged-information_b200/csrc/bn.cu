// BatchNorm2d over NHWC activations viewed as [rows, C], fused with ReLU / residual add / veil
// (row mask) in the forward pass and with the PartialConv ratio (row scale) in the backward pass.
// Replaces nn.BatchNorm2d + F.relu + "out += residual" of the reference blocks
// (partial_depthnet.py:143-157, fusionnet.py:107-127).  All four kernels are HBM-bound streams:
// thread (tx, ty) owns 4 consecutive channels (one 16 B / 8 B vector) and strides over rows, so
// per-channel constants live in registers and every warp reads whole contiguous row segments.
// Statistics are accumulated in fp32 over short per-thread runs and combined in fp64.
#include "b2_common.cuh"

namespace {

struct Geo {
  dim3 grid, block;
};

Geo geometry(long long rows, int C, int rows_per_thread) {
  int C4 = C >> 2;
  int tx = C4 < 64 ? C4 : 64;
  int ty = 256 / tx;
  if (ty < 1) ty = 1;
  int gy = (C4 + tx - 1) / tx;
  long long gx = (rows + (long long)ty * rows_per_thread - 1) / ((long long)ty * rows_per_thread);
  long long cap = ((long long)b2_num_sms() * 8 + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  Geo g;
  g.grid = dim3((unsigned)gx, gy);
  g.block = dim3(tx, ty);
  return g;
}

__device__ __forceinline__ void block_reduce_to_double(float4 a, float4 b, double* out_a, double* out_b, int cg,
                                                       int C4, float4* red) {
  // red: [2][blockDim.y][blockDim.x] float4
  const int n = blockDim.x * blockDim.y;
  red[threadIdx.y * blockDim.x + threadIdx.x] = a;
  red[n + threadIdx.y * blockDim.x + threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && cg < C4) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < blockDim.y; ++i) {
      float4 u = red[i * blockDim.x + threadIdx.x], v = red[n + i * blockDim.x + threadIdx.x];
      s[0] += u.x; s[1] += u.y; s[2] += u.z; s[3] += u.w;
      s[4] += v.x; s[5] += v.y; s[6] += v.z; s[7] += v.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(out_a + cg * 4 + j, s[j]);
      atomicAdd(out_b + cg * 4 + j, s[4 + j]);
    }
  }
}

template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ y, long long rows, int C, double* __restrict__ sums) {
  extern __shared__ float4 red[];
  const int C4 = C >> 2, cg = blockIdx.y * blockDim.x + threadIdx.x;
  float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  if (cg < C4) {
    for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < rows;
         r += (long long)gridDim.x * blockDim.y) {
      float4 f = load4(y + r * C + cg * 4);
      s.x += f.x; s.y += f.y; s.z += f.z; s.w += f.w;
      q.x += f.x * f.x; q.y += f.y * f.y; q.z += f.z * f.z; q.w += f.w * f.w;
    }
  }
  block_reduce_to_double(s, q, sums, sums + C, cg, C4, red);
}

struct ApplyP {
  const double* sums;
  const float *gamma, *beta;
  float *running_mean, *running_var, *save_mean, *save_invstd;
  const float* row_mask;
  float momentum, eps;
  int training, relu;
  long long rows;
  int C;
};

template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ y, const T* __restrict__ residual, T* __restrict__ z,
                                ApplyP p) {
  const int C4 = p.C >> 2, cg = blockIdx.y * blockDim.x + threadIdx.x;
  if (cg >= C4) return;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = cg * 4 + j;
    float mean, invstd;
    if (p.training) {
      double n = (double)p.rows;
      double m = p.sums[c] / n;
      double var = p.sums[p.C + c] / n - m * m;
      if (var < 0) var = 0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)p.eps));
      if (blockIdx.x == 0 && threadIdx.y == 0 && p.running_mean) {
        double unbiased = n > 1 ? var * n / (n - 1) : var;
        p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
        p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unbiased;
      }
    } else {
      mean = p.running_mean[c];
      invstd = 1.f / sqrtf(p.running_var[c] + p.eps);
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) {
      if (p.save_mean) p.save_mean[c] = mean;
      if (p.save_invstd) p.save_invstd[c] = invstd;
    }
    sc[j] = invstd * p.gamma[c];
    sh[j] = p.beta[c] - mean * sc[j];
  }
  for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < p.rows;
       r += (long long)gridDim.x * blockDim.y) {
    long long off = r * p.C + cg * 4;
    float4 f = load4(y + off);
    f.x = fmaf(f.x, sc[0], sh[0]); f.y = fmaf(f.y, sc[1], sh[1]);
    f.z = fmaf(f.z, sc[2], sh[2]); f.w = fmaf(f.w, sc[3], sh[3]);
    if (residual) {
      float4 g = load4(residual + off);
      f.x += g.x; f.y += g.y; f.z += g.z; f.w += g.w;
    }
    if (p.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
    if (p.row_mask) {
      float mk = p.row_mask[r];
      f.x *= mk; f.y *= mk; f.z *= mk; f.w *= mk;
    }
    store4(z + off, f);
  }
}

// g = dz * relu'(z) * row_mask ; sums = [sum g | sum g*xhat]
template <typename T>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ row_mask, int relu, double* __restrict__ sums,
                                     long long rows, int C) {
  extern __shared__ float4 red[];
  const int C4 = C >> 2, cg = blockIdx.y * blockDim.x + threadIdx.x;
  float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  if (cg < C4) {
    float4 mu = load4(mean + cg * 4), is = load4(invstd + cg * 4);
    for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < rows;
         r += (long long)gridDim.x * blockDim.y) {
      long long off = r * C + cg * 4;
      float4 g = load4(dz + off);
      if (relu) {
        float4 zz = load4(z + off);
        g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f;
        g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
      }
      if (row_mask) {
        float mk = row_mask[r];
        g.x *= mk; g.y *= mk; g.z *= mk; g.w *= mk;
      }
      float4 f = load4(y + off);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      q.x += g.x * (f.x - mu.x) * is.x; q.y += g.y * (f.y - mu.y) * is.y;
      q.z += g.z * (f.z - mu.z) * is.z; q.w += g.w * (f.w - mu.w) * is.w;
    }
  }
  block_reduce_to_double(s, q, sums, sums + C, cg, C4, red);
}

template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const double* __restrict__ sums,
                                    const float* __restrict__ row_mask, const float* __restrict__ row_scale,
                                    int relu, int training, T* __restrict__ dy, T* __restrict__ d_residual,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int C) {
  const int C4 = C >> 2, cg = blockIdx.y * blockDim.x + threadIdx.x;
  if (cg >= C4) return;
  float mu[4], is[4], a[4], mg[4], mgx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = cg * 4 + j;
    mu[j] = mean[c];
    is[j] = invstd[c];
    a[j] = gamma[c] * is[j];
    double sg = sums[c], sgx = sums[C + c];
    if (training) {
      mg[j] = (float)(sg / (double)rows);
      mgx[j] = (float)(sgx / (double)rows);
    } else {
      mg[j] = 0.f;
      mgx[j] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) {
      if (dgamma) dgamma[c] += (float)sgx;
      if (dbeta) dbeta[c] += (float)sg;
    }
  }
  for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < rows;
       r += (long long)gridDim.x * blockDim.y) {
    long long off = r * C + cg * 4;
    float4 g = load4(dz + off);
    if (relu) {
      float4 zz = load4(z + off);
      g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f;
      g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
    }
    if (row_mask) {
      float mk = row_mask[r];
      g.x *= mk; g.y *= mk; g.z *= mk; g.w *= mk;
    }
    if (d_residual) store4(d_residual + off, g);
    float4 f = load4(y + off);
    float rs = row_scale ? row_scale[r] : 1.f;
    float4 o;
    o.x = a[0] * (g.x - mg[0] - (f.x - mu[0]) * is[0] * mgx[0]) * rs;
    o.y = a[1] * (g.y - mg[1] - (f.y - mu[1]) * is[1] * mgx[1]) * rs;
    o.z = a[2] * (g.z - mg[2] - (f.z - mu[2]) * is[2] * mgx[2]) * rs;
    o.w = a[3] * (g.w - mg[3] - (f.w - mu[3]) * is[3] * mgx[3]) * rs;
    store4(dy + off, o);
  }
}

}  // namespace

#define BN_DISPATCH(dtype, CALL_F32, CALL_BF16) \
  do {                                          \
    if ((dtype) == B2_F32) { CALL_F32; }        \
    else { CALL_BF16; }                         \
  } while (0)

extern "C" int b2_bn_stats(const void* y, int64_t rows, int32_t C, int32_t dtype, double* sums, void* stream) {
  B2_REQUIRE(y && sums && rows > 0 && C > 0, B2_E_BADARG, "bn_stats: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_stats: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  Geo g = geometry(rows, C, 16);
  size_t sh = 2 * sizeof(float4) * g.block.x * g.block.y;
  BN_DISPATCH(dtype, (bn_stats_kernel<float><<<g.grid, g.block, sh, st>>>((const float*)y, rows, C, sums)),
              (bn_stats_kernel<bf16><<<g.grid, g.block, sh, st>>>((const bf16*)y, rows, C, sums)));
  B2_LAUNCH_CHECK("bn_stats");
  return B2_OK;
}

extern "C" int b2_bn_apply(const void* y, const double* sums, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float momentum, float eps, int32_t training,
                           const void* residual, const float* row_mask, int32_t relu, void* z, float* save_mean,
                           float* save_invstd, int64_t rows, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(y && z && gamma && beta && rows > 0 && C > 0, B2_E_BADARG, "bn_apply: bad argument");
  B2_REQUIRE(training ? (sums != nullptr) : (running_mean && running_var), B2_E_BADARG,
             "bn_apply: statistics source missing");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_apply: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  ApplyP p;
  p.sums = sums; p.gamma = gamma; p.beta = beta; p.running_mean = running_mean; p.running_var = running_var;
  p.save_mean = save_mean; p.save_invstd = save_invstd; p.row_mask = row_mask; p.momentum = momentum; p.eps = eps;
  p.training = training; p.relu = relu; p.rows = rows; p.C = C;
  Geo g = geometry(rows, C, 8);
  BN_DISPATCH(dtype,
              (bn_apply_kernel<float><<<g.grid, g.block, 0, st>>>((const float*)y, (const float*)residual, (float*)z, p)),
              (bn_apply_kernel<bf16><<<g.grid, g.block, 0, st>>>((const bf16*)y, (const bf16*)residual, (bf16*)z, p)));
  B2_LAUNCH_CHECK("bn_apply");
  return B2_OK;
}

extern "C" int b2_bn_bwd_reduce(const void* dz, const void* z, const void* y, const float* mean,
                                const float* invstd, const float* row_mask, int32_t relu, double* sums,
                                int64_t rows, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(dz && y && mean && invstd && sums && rows > 0 && C > 0 && (!relu || z), B2_E_BADARG,
             "bn_bwd_reduce: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_bwd_reduce: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  Geo g = geometry(rows, C, 16);
  size_t sh = 2 * sizeof(float4) * g.block.x * g.block.y;
  BN_DISPATCH(dtype,
              (bn_bwd_reduce_kernel<float><<<g.grid, g.block, sh, st>>>((const float*)dz, (const float*)z, (const float*)y,
                                                                        mean, invstd, row_mask, relu, sums, rows, C)),
              (bn_bwd_reduce_kernel<bf16><<<g.grid, g.block, sh, st>>>((const bf16*)dz, (const bf16*)z, (const bf16*)y,
                                                                       mean, invstd, row_mask, relu, sums, rows, C)));
  B2_LAUNCH_CHECK("bn_bwd_reduce");
  return B2_OK;
}

extern "C" int b2_bn_bwd_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                               const float* gamma, const double* sums, const float* row_mask,
                               const float* row_scale, int32_t relu, int32_t training, void* dy, void* d_residual,
                               float* dgamma, float* dbeta, int64_t rows, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(dz && y && mean && invstd && gamma && sums && dy && rows > 0 && C > 0 && (!relu || z), B2_E_BADARG,
             "bn_bwd_apply: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_bwd_apply: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  Geo g = geometry(rows, C, 8);
  BN_DISPATCH(dtype,
              (bn_bwd_apply_kernel<float><<<g.grid, g.block, 0, st>>>(
                  (const float*)dz, (const float*)z, (const float*)y, mean, invstd, gamma, sums, row_mask, row_scale,
                  relu, training, (float*)dy, (float*)d_residual, dgamma, dbeta, rows, C)),
              (bn_bwd_apply_kernel<bf16><<<g.grid, g.block, 0, st>>>(
                  (const bf16*)dz, (const bf16*)z, (const bf16*)y, mean, invstd, gamma, sums, row_mask, row_scale,
                  relu, training, (bf16*)dy, (bf16*)d_residual, dgamma, dbeta, rows, C)));
  B2_LAUNCH_CHECK("bn_bwd_apply");
  return B2_OK;
}
