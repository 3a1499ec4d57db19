// BatchNorm2d over NHWC activations viewed as [rows, C], fused with ReLU / residual add / veil
// (row mask) in the forward pass and with the PartialConv ratio (row scale) in the backward pass.
// Replaces nn.BatchNorm2d + F.relu + "out += residual" of the reference blocks
// (partial_depthnet.py:143-157, fusionnet.py:107-127).  All four kernels are HBM-bound streams:
// thread (tx, ty) owns one 16-byte channel vector (4 fp32 / 8 bf16 channels) and strides over rows
// two at a time with all loads issued before use, so per-channel constants live in registers and
// every warp reads whole contiguous row segments.  Per-channel statistics are produced WITHOUT
// atomics: every block writes its fp32 partial sums to its own slot of a [B2_BN_PARTS][2C] buffer
// (same-address fp64 atomics cost ~40 ns each at L2 and dominated the reduction kernels), and a
// tiny finalize kernel combines the slots in fp64 -- which also makes the statistics deterministic.
#include <cstdlib>

#include "b2_common.cuh"

namespace {

template <typename T, int N> struct VecIO;
template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float* f) {
    float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct VecIO<bf16, 8> {
  static __device__ __forceinline__ void ld(const bf16* p, float* f) { load8(p, f); }
  static __device__ __forceinline__ void st(bf16* p, const float* f) { store8(p, f); }
};
template <> struct VecIO<bf16, 4> {
  static __device__ __forceinline__ void ld(const bf16* p, float* f) {
    float4 v = load4(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void st(bf16* p, const float* f) { store4(p, make_float4(f[0], f[1], f[2], f[3])); }
};

struct Geo {
  dim3 grid, block;
};

// CV = channel vectors per row (C / N)
Geo geometry(long long rows, int CV, int rows_per_thread, int blocks_per_sm, bool slot_writer = false) {
  static const int env_bps = getenv("B2POSE_BN_BPS") ? atoi(getenv("B2POSE_BN_BPS")) : 0;     // tuning overrides
  static const int env_rpt = getenv("B2POSE_BN_RPT") ? atoi(getenv("B2POSE_BN_RPT")) : 0;
  if (env_bps > 0) blocks_per_sm = env_bps;
  if (env_rpt > 0) rows_per_thread = env_rpt;
  int tx = CV < 64 ? CV : 64;
  int ty = 256 / tx;
  if (ty < 1) ty = 1;
  int gy = (CV + tx - 1) / tx;
  long long gx = (rows + (long long)ty * rows_per_thread - 1) / ((long long)ty * rows_per_thread);
  long long cap = ((long long)b2_num_sms() * blocks_per_sm + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (slot_writer && gx > B2_BN_PARTS) gx = B2_BN_PARTS;
  if (gx < 1) gx = 1;
  Geo g;
  g.grid = dim3((unsigned)gx, gy);
  g.block = dim3(tx, ty);
  return g;
}

// block reduce of N-wide partials a[], b[] over threadIdx.y into this block's slot of
// partials[B2_BN_PARTS][2C]; slots beyond gridDim.x are zero-filled (spread over the blocks).
template <int N>
__device__ __forceinline__ void block_reduce_to_slot(const float* a, const float* b, float* __restrict__ partials,
                                                     int C, int cv, int CV, float* red) {
  // red: [2][blockDim.y][blockDim.x][N]
  const int nthr = blockDim.x * blockDim.y;
  float* ra = red + (size_t)(threadIdx.y * blockDim.x + threadIdx.x) * N;
  float* rb = ra + (size_t)nthr * N;
#pragma unroll
  for (int j = 0; j < N; ++j) { ra[j] = a[j]; rb[j] = b[j]; }
  __syncthreads();
  if (threadIdx.y == 0 && cv < CV) {
    float sa[N], sb[N];
#pragma unroll
    for (int j = 0; j < N; ++j) sa[j] = sb[j] = 0.f;
    for (int i = 0; i < blockDim.y; ++i) {
      const float* pa = red + (size_t)(i * blockDim.x + threadIdx.x) * N;
      const float* pb = pa + (size_t)nthr * N;
#pragma unroll
      for (int j = 0; j < N; ++j) { sa[j] += pa[j]; sb[j] += pb[j]; }
    }
    float* slot = partials + (size_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      slot[cv * N + j] = sa[j];
      slot[C + cv * N + j] = sb[j];
    }
    for (int sl = gridDim.x + blockIdx.x; sl < B2_BN_PARTS; sl += gridDim.x) {
      float* zs = partials + (size_t)sl * 2 * C;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        zs[cv * N + j] = 0.f;
        zs[C + cv * N + j] = 0.f;
      }
    }
  }
}

template <typename T, int N>
__global__ void __launch_bounds__(256, 4) bn_stats_kernel(const T* __restrict__ y, long long rows, int C,
                                                           float* __restrict__ partials) {
  extern __shared__ float red[];
  const int CV = C / N, cv = blockIdx.y * blockDim.x + threadIdx.x;
  float s[N], q[N];
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = q[j] = 0.f;
  if (cv < CV) {
    const long long step = (long long)gridDim.x * blockDim.y;
    long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y;
    for (; r + step < rows; r += 2 * step) {
      float f0[N], f1[N];
      VecIO<T, N>::ld(y + r * C + cv * N, f0);
      VecIO<T, N>::ld(y + (r + step) * C + cv * N, f1);
#pragma unroll
      for (int j = 0; j < N; ++j) { s[j] += f0[j] + f1[j]; q[j] = fmaf(f0[j], f0[j], fmaf(f1[j], f1[j], q[j])); }
    }
    if (r < rows) {
      float f0[N];
      VecIO<T, N>::ld(y + r * C + cv * N, f0);
#pragma unroll
      for (int j = 0; j < N; ++j) { s[j] += f0[j]; q[j] = fmaf(f0[j], f0[j], q[j]); }
    }
  }
  block_reduce_to_slot<N>(s, q, partials, C, cv, CV, red);
}

// Sum of the B2_BN_PARTS slots for 32 channels per block: thread (c, l) adds slots l, l+8, ...
// (independent loads), then the 8 lanes of a channel are combined in shared memory in fp64.
__device__ __forceinline__ void slot_sums(const float* __restrict__ partials, int C, int c, double* s_out,
                                          double* q_out, double (*sm)[8][33]) {
  const int l = threadIdx.y;
  double s = 0.0, q = 0.0;
  if (c < C) {
    float ps[B2_BN_PARTS / 8], pq[B2_BN_PARTS / 8];
#pragma unroll
    for (int i = 0; i < B2_BN_PARTS / 8; ++i) {            // 40 independent loads each, then the fp64 sums
      ps[i] = partials[(size_t)(l + 8 * i) * 2 * C + c];
      pq[i] = partials[(size_t)(l + 8 * i) * 2 * C + C + c];
    }
#pragma unroll
    for (int i = 0; i < B2_BN_PARTS / 8; ++i) { s += (double)ps[i]; q += (double)pq[i]; }
  }
  sm[0][l][threadIdx.x] = s;
  sm[1][l][threadIdx.x] = q;
  __syncthreads();
  s = q = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += sm[0][i][threadIdx.x]; q += sm[1][i][threadIdx.x]; }
  *s_out = s;
  *q_out = q;
}

// mean / invstd from the partial sums (training) or the running statistics (eval); block (32, 8)
__global__ void bn_finalize_kernel(const float* __restrict__ partials, long long rows, int C,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, int training, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  __shared__ double sm[2][8][33];
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  float mean = 0.f, invstd = 0.f;
  if (training) {
    double s, q;
    slot_sums(partials, C, c, &s, &q, sm);
    const double n = (double)rows, m = s / n;
    double var = q / n - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean && threadIdx.y == 0 && c < C) {
      const double unbiased = n > 1 ? var * n / (n - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else if (c < C) {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + eps);
  }
  if (threadIdx.y == 0 && c < C) {
    mean_out[c] = mean;
    invstd_out[c] = invstd;
  }
}

// gsum = [sum g | sum g*xhat] from the partial sums; dgamma += sum g*xhat, dbeta += sum g
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int C, float* __restrict__ gsum,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double sm[2][8][33];
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  slot_sums(partials, C, c, &s, &q, sm);
  if (threadIdx.y == 0 && c < C) {
    gsum[c] = (float)s;
    gsum[C + c] = (float)q;
    if (dgamma) dgamma[c] += (float)q;
    if (dbeta) dbeta[c] += (float)s;
  }
}

struct ApplyP {
  const float *mean, *invstd, *gamma, *beta;
  const float* row_mask;
  int relu;
  long long rows;
  int C;
};

template <typename T, int N>
__global__ void __launch_bounds__(256, 4) bn_apply_kernel(const T* __restrict__ y, const T* __restrict__ residual,
                                                           T* __restrict__ z, ApplyP p) {
  const int CV = p.C / N, cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= CV) return;
  float sc[N], sh[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const int c = cv * N + j;
    sc[j] = p.invstd[c] * p.gamma[c];
    sh[j] = p.beta[c] - p.mean[c] * sc[j];
  }
  const long long step = (long long)gridDim.x * blockDim.y;
  for (long long r0 = blockIdx.x * (long long)blockDim.y + threadIdx.y; r0 < p.rows; r0 += 2 * step) {
    const long long r1 = r0 + step;
    const bool two = r1 < p.rows;
    const long long o0 = r0 * p.C + cv * N, o1 = r1 * p.C + cv * N;
    float f0[N], f1[N], g0[N], g1[N];
    VecIO<T, N>::ld(y + o0, f0);
    if (two) VecIO<T, N>::ld(y + o1, f1);
    if (residual) {
      VecIO<T, N>::ld(residual + o0, g0);
      if (two) VecIO<T, N>::ld(residual + o1, g1);
    }
    const float m0 = p.row_mask ? p.row_mask[r0] : 1.f;
    const float m1 = (p.row_mask && two) ? p.row_mask[r1] : 1.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      float a = fmaf(f0[j], sc[j], sh[j]), b = fmaf(f1[j], sc[j], sh[j]);
      if (residual) { a += g0[j]; b += g1[j]; }
      if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      f0[j] = a * m0;
      f1[j] = b * m1;
    }
    VecIO<T, N>::st(z + o0, f0);
    if (two) VecIO<T, N>::st(z + o1, f1);
  }
}

// ReLU gate of the backward pass: from the stored output z when given, else recomputed from y with
// exactly the forward's fp32 expression fma(y, invstd*gamma, beta - mean*invstd*gamma) > 0 (layers
// without a residual), which saves one full read of z in each backward kernel.
template <typename T, int N>
__device__ __forceinline__ void gated_grad(const T* __restrict__ dz, const T* __restrict__ z, long long off, bool regate,
                                           int relu, const float* f, const float* sc, const float* sh, float mk,
                                           float* g) {
  VecIO<T, N>::ld(dz + off, g);
  if (regate) {
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = fmaf(f[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
  } else if (relu) {
    float zz[N];
    VecIO<T, N>::ld(z + off, zz);
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = zz[j] > 0.f ? g[j] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < N; ++j) g[j] *= mk;
}

// g = dz * relu'(z) * row_mask ; sums = [sum g | sum g*xhat]
template <typename T, int N>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ row_mask, int relu,
                     float* __restrict__ partials, long long rows, int C) {
  extern __shared__ float red[];
  const int CV = C / N, cv = blockIdx.y * blockDim.x + threadIdx.x;
  float s[N], q[N];
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = q[j] = 0.f;
  if (cv < CV) {
    const bool regate = relu && z == nullptr;
    float mu[N], sc[N], sh[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int c = cv * N + j;
      mu[j] = mean[c];
      sc[j] = regate ? invstd[c] * gamma[c] : 0.f;
      sh[j] = regate ? beta[c] - mu[j] * sc[j] : 0.f;
    }
    const long long step = (long long)gridDim.x * blockDim.y;
    for (long long r0 = blockIdx.x * (long long)blockDim.y + threadIdx.y; r0 < rows; r0 += 2 * step) {
      const long long r1 = r0 + step;
      const bool two = r1 < rows;
      const long long o0 = r0 * C + cv * N, o1 = r1 * C + cv * N;
      float f0[N], f1[N], g0[N], g1[N];
      VecIO<T, N>::ld(y + o0, f0);
      if (two) VecIO<T, N>::ld(y + o1, f1);
      gated_grad<T, N>(dz, z, o0, regate, relu, f0, sc, sh, row_mask ? row_mask[r0] : 1.f, g0);
      if (two) gated_grad<T, N>(dz, z, o1, regate, relu, f1, sc, sh, row_mask ? row_mask[r1] : 1.f, g1);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        s[j] += g0[j];
        q[j] = fmaf(g0[j], f0[j] - mu[j], q[j]);
        if (two) {
          s[j] += g1[j];
          q[j] = fmaf(g1[j], f1[j] - mu[j], q[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) q[j] *= invstd[cv * N + j];
  }
  block_reduce_to_slot<N>(s, q, partials, C, cv, CV, red);
}

// dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) * row_scale  ==  (g*A + y*B + D) * row_scale
template <typename T, int N>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ gsum, const float* __restrict__ row_mask,
                    const float* __restrict__ row_scale, int relu, int training, T* __restrict__ dy,
                    T* __restrict__ d_residual, long long rows, int C) {
  const int CV = C / N, cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= CV) return;
  const bool regate = relu && z == nullptr;
  float A[N], B[N], D[N], sc[N], sh[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const int c = cv * N + j;
    const float mu = mean[c], is = invstd[c], a = gamma[c] * is;
    const float mg = training ? gsum[c] / (float)rows : 0.f;
    const float mgx = training ? gsum[C + c] / (float)rows : 0.f;
    A[j] = a;
    B[j] = -a * is * mgx;
    D[j] = -a * mg - mu * B[j];
    sc[j] = regate ? a : 0.f;
    sh[j] = regate ? beta[c] - mu * a : 0.f;
  }
  const long long step = (long long)gridDim.x * blockDim.y;
  for (long long r0 = blockIdx.x * (long long)blockDim.y + threadIdx.y; r0 < rows; r0 += 2 * step) {
    const long long r1 = r0 + step;
    const bool two = r1 < rows;
    const long long o0 = r0 * C + cv * N, o1 = r1 * C + cv * N;
    float f0[N], f1[N], g0[N], g1[N];
    VecIO<T, N>::ld(y + o0, f0);
    if (two) VecIO<T, N>::ld(y + o1, f1);
    gated_grad<T, N>(dz, z, o0, regate, relu, f0, sc, sh, row_mask ? row_mask[r0] : 1.f, g0);
    if (two) gated_grad<T, N>(dz, z, o1, regate, relu, f1, sc, sh, row_mask ? row_mask[r1] : 1.f, g1);
    if (d_residual) {
      VecIO<T, N>::st(d_residual + o0, g0);
      if (two) VecIO<T, N>::st(d_residual + o1, g1);
    }
    const float s0 = row_scale ? row_scale[r0] : 1.f;
    const float s1 = (row_scale && two) ? row_scale[r1] : 1.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      f0[j] = fmaf(g0[j], A[j], fmaf(f0[j], B[j], D[j])) * s0;
      f1[j] = fmaf(g1[j], A[j], fmaf(f1[j], B[j], D[j])) * s1;
    }
    VecIO<T, N>::st(dy + o0, f0);
    if (two) VecIO<T, N>::st(dy + o1, f1);
  }
}

// vector width: 4 fp32 / 8 bf16 channels (16 bytes); bf16 tensors with C % 8 != 0 use 4 (8 bytes)
inline int vec_width(int C, int dtype) { return (dtype == B2_BF16 && (C & 7) == 0) ? 8 : 4; }

}  // namespace

extern "C" int b2_bn_stats(const void* y, int64_t rows, int32_t C, int32_t dtype, float* partials, void* stream) {
  B2_REQUIRE(y && partials && rows > 0 && C > 0, B2_E_BADARG, "bn_stats: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_stats: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_stream_eligible(C, dtype)) return bn_stream_stats(y, rows, C, partials, 0, st);
  const int nv = vec_width(C, dtype);
  Geo g = geometry(rows, C / nv, 16, 2, true);
  size_t sh = 2 * sizeof(float) * nv * g.block.x * g.block.y;
  if (dtype == B2_F32) bn_stats_kernel<float, 4><<<g.grid, g.block, sh, st>>>((const float*)y, rows, C, partials);
  else if (nv == 8) bn_stats_kernel<bf16, 8><<<g.grid, g.block, sh, st>>>((const bf16*)y, rows, C, partials);
  else bn_stats_kernel<bf16, 4><<<g.grid, g.block, sh, st>>>((const bf16*)y, rows, C, partials);
  B2_LAUNCH_CHECK("bn_stats");
  return B2_OK;
}

extern "C" int b2_bn_finalize(const float* partials, int64_t rows, int32_t C, float* running_mean, float* running_var,
                              float momentum, float eps, int32_t training, float* mean, float* invstd, void* stream) {
  B2_REQUIRE(mean && invstd && rows > 0 && C > 0, B2_E_BADARG, "bn_finalize: bad argument");
  B2_REQUIRE(training ? (partials != nullptr) : (running_mean && running_var), B2_E_BADARG,
             "bn_finalize: statistics source missing");
  launch_pdl(bn_finalize_kernel, dim3((C + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream, partials,
             (long long)rows, (int)C, running_mean, running_var, momentum, eps, (int)training, mean, invstd);
  B2_LAUNCH_CHECK("bn_finalize");
  return B2_OK;
}

extern "C" int b2_bn_apply(const void* y, const float* mean, const float* invstd, const float* gamma,
                           const float* beta, const void* residual, const float* row_mask, int32_t relu, void* z,
                           int64_t rows, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(y && z && mean && invstd && gamma && beta && rows > 0 && C > 0, B2_E_BADARG, "bn_apply: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_apply: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_stream_eligible(C, dtype))
    return bn_stream_apply(y, residual, z, mean, invstd, gamma, beta, row_mask, relu, rows, C, st);
  ApplyP p;
  p.mean = mean; p.invstd = invstd; p.gamma = gamma; p.beta = beta; p.row_mask = row_mask; p.relu = relu;
  p.rows = rows; p.C = C;
  const int nv = vec_width(C, dtype);
  Geo g = geometry(rows, C / nv, 8, 4);
  if (dtype == B2_F32)
    bn_apply_kernel<float, 4><<<g.grid, g.block, 0, st>>>((const float*)y, (const float*)residual, (float*)z, p);
  else if (nv == 8)
    bn_apply_kernel<bf16, 8><<<g.grid, g.block, 0, st>>>((const bf16*)y, (const bf16*)residual, (bf16*)z, p);
  else
    bn_apply_kernel<bf16, 4><<<g.grid, g.block, 0, st>>>((const bf16*)y, (const bf16*)residual, (bf16*)z, p);
  B2_LAUNCH_CHECK("bn_apply");
  return B2_OK;
}

extern "C" int b2_bn_bwd_reduce(const void* dz, const void* z, const void* y, const float* mean,
                                const float* invstd, const float* gamma, const float* beta,
                                const float* row_mask, int32_t relu, float* partials, int64_t rows, int32_t C,
                                int32_t dtype, void* stream) {
  B2_REQUIRE(dz && y && mean && invstd && partials && rows > 0 && C > 0 && (!relu || z || (gamma && beta)),
             B2_E_BADARG, "bn_bwd_reduce: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_bwd_reduce: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_stream_eligible(C, dtype))
    return bn_stream_bwd_reduce(dz, z, y, mean, invstd, gamma, beta, row_mask, relu, partials, 0, nullptr, rows, C, st);
  const int nv = vec_width(C, dtype);
  Geo g = geometry(rows, C / nv, 16, 2, true);
  size_t sh = 2 * sizeof(float) * nv * g.block.x * g.block.y;
  if (dtype == B2_F32)
    bn_bwd_reduce_kernel<float, 4><<<g.grid, g.block, sh, st>>>((const float*)dz, (const float*)z, (const float*)y, mean,
                                                                 invstd, gamma, beta, row_mask, relu, partials, rows, C);
  else if (nv == 8)
    bn_bwd_reduce_kernel<bf16, 8><<<g.grid, g.block, sh, st>>>((const bf16*)dz, (const bf16*)z, (const bf16*)y, mean,
                                                                invstd, gamma, beta, row_mask, relu, partials, rows, C);
  else
    bn_bwd_reduce_kernel<bf16, 4><<<g.grid, g.block, sh, st>>>((const bf16*)dz, (const bf16*)z, (const bf16*)y, mean,
                                                                invstd, gamma, beta, row_mask, relu, partials, rows, C);
  B2_LAUNCH_CHECK("bn_bwd_reduce");
  return B2_OK;
}

extern "C" int b2_bn_bwd_finalize(const float* partials, int32_t C, float* gsum, float* dgamma, float* dbeta,
                                  void* stream) {
  B2_REQUIRE(partials && gsum && C > 0, B2_E_BADARG, "bn_bwd_finalize: bad argument");
  launch_pdl(bn_bwd_finalize_kernel, dim3((C + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream, partials, (int)C, gsum,
             dgamma, dbeta);
  B2_LAUNCH_CHECK("bn_bwd_finalize");
  return B2_OK;
}

extern "C" int b2_bn_bwd_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                               const float* gamma, const float* beta, const float* gsum, const float* row_mask,
                               const float* row_scale, int32_t relu, int32_t training, void* dy, void* d_residual,
                               int64_t rows, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(dz && y && mean && invstd && gamma && gsum && dy && rows > 0 && C > 0 && (!relu || z || beta), B2_E_BADARG,
             "bn_bwd_apply: bad argument");
  B2_REQUIRE((C & 3) == 0, B2_E_UNSUPPORTED, "bn_bwd_apply: C=%d is not a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_stream_eligible(C, dtype))
    return bn_stream_bwd_apply(dz, z, y, mean, invstd, gamma, beta, gsum, row_mask, row_scale, relu, training, dy,
                               d_residual, nullptr, nullptr, nullptr, rows, C, st);
  const int nv = vec_width(C, dtype);
  Geo g = geometry(rows, C / nv, 8, 4);
  if (dtype == B2_F32)
    bn_bwd_apply_kernel<float, 4><<<g.grid, g.block, 0, st>>>(
        (const float*)dz, (const float*)z, (const float*)y, mean, invstd, gamma, beta, gsum, row_mask, row_scale, relu,
        training, (float*)dy, (float*)d_residual, rows, C);
  else if (nv == 8)
    bn_bwd_apply_kernel<bf16, 8><<<g.grid, g.block, 0, st>>>(
        (const bf16*)dz, (const bf16*)z, (const bf16*)y, mean, invstd, gamma, beta, gsum, row_mask, row_scale, relu,
        training, (bf16*)dy, (bf16*)d_residual, rows, C);
  else
    bn_bwd_apply_kernel<bf16, 4><<<g.grid, g.block, 0, st>>>(
        (const bf16*)dz, (const bf16*)z, (const bf16*)y, mean, invstd, gamma, beta, gsum, row_mask, row_scale, relu,
        training, (bf16*)dy, (bf16*)d_residual, rows, C);
  B2_LAUNCH_CHECK("bn_bwd_apply");
  return B2_OK;
}

// ---------------------------------------------------------------- totals path (no finalize kernels)
extern "C" int b2_bn_totals_supported(int32_t C, int32_t dtype) { return bn_stream_eligible(C, dtype) ? 1 : 0; }

extern "C" int b2_bn_stats_totals(const void* y, int64_t rows, int32_t C, int32_t dtype, float* totals, void* stream) {
  B2_REQUIRE(y && totals && rows > 0 && C > 0, B2_E_BADARG, "bn_stats_totals: bad argument");
  B2_REQUIRE(bn_stream_eligible(C, dtype), B2_E_UNSUPPORTED, "bn_stats_totals: C=%d dtype=%d has no totals path", C, dtype);
  return bn_stream_stats(y, rows, C, totals, 1, (cudaStream_t)stream);
}

extern "C" int b2_bn_apply_totals(const void* y, const float* totals, int64_t rows, float* running_mean,
                                  float* running_var, float momentum, float eps, int32_t training,
                                  const float* gamma, const float* beta, const void* residual,
                                  const float* row_mask, int32_t relu, void* z, float* mean, float* invstd,
                                  uint8_t* gate_out, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(y && z && gamma && beta && mean && invstd && rows > 0 && C > 0, B2_E_BADARG, "bn_apply_totals: bad argument");
  B2_REQUIRE(training ? (totals != nullptr) : (running_mean && running_var), B2_E_BADARG,
             "bn_apply_totals: statistics source missing");
  B2_REQUIRE(bn_stream_eligible(C, dtype), B2_E_UNSUPPORTED, "bn_apply_totals: C=%d dtype=%d has no totals path", C, dtype);
  return bn_stream_apply_fin(y, residual, z, nullptr, nullptr, gamma, beta, row_mask, relu, rows, C, training ? 1 : 2,
                             totals, running_mean, running_var, momentum, eps, mean, invstd, gate_out,
                             (cudaStream_t)stream);
}

extern "C" int b2_bn_bwd_reduce_totals(const void* dz, const void* z, const void* y, const float* mean,
                                       const float* invstd, const float* gamma, const float* beta,
                                       const float* row_mask, int32_t relu, float* gsum, const uint8_t* gate,
                                       int64_t rows, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(dz && y && mean && invstd && gsum && rows > 0 && C > 0 && (!relu || z || gate || (gamma && beta)), B2_E_BADARG,
             "bn_bwd_reduce_totals: bad argument");
  B2_REQUIRE(bn_stream_eligible(C, dtype), B2_E_UNSUPPORTED, "bn_bwd_reduce_totals: C=%d dtype=%d has no totals path", C,
             dtype);
  return bn_stream_bwd_reduce(dz, z, y, mean, invstd, gamma, beta, row_mask, relu, gsum, 1, gate, rows, C,
                              (cudaStream_t)stream);
}

extern "C" int b2_bn_bwd_apply_totals(const void* dz, const void* z, const void* y, const float* mean,
                                      const float* invstd, const float* gamma, const float* beta, const float* gsum,
                                      const float* row_mask, const float* row_scale, int32_t relu, int32_t training,
                                      void* dy, void* d_residual, float* dgamma, float* dbeta, const uint8_t* gate,
                                      int64_t rows, int32_t C, int32_t dtype, void* stream) {
  B2_REQUIRE(dz && y && mean && invstd && gamma && gsum && dy && rows > 0 && C > 0 && (!relu || z || gate || beta), B2_E_BADARG,
             "bn_bwd_apply_totals: bad argument");
  B2_REQUIRE(bn_stream_eligible(C, dtype), B2_E_UNSUPPORTED, "bn_bwd_apply_totals: C=%d dtype=%d has no totals path", C,
             dtype);
  return bn_stream_bwd_apply(dz, z, y, mean, invstd, gamma, beta, gsum, row_mask, row_scale, relu, training, dy,
                             d_residual, dgamma, dbeta, gate, rows, C, (cudaStream_t)stream);
}
