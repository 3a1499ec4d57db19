// Feature-mimic (distillation) losses of the "privileged information" step, depth_train.py:115-129:
//
//   mode 0 (default)  : mean_n || (t - s) * a ||_2
//   mode 1 (sigmoid)  : mean_n || (sigmoid(t) - sigmoid(s)) * a ||_2
//   mode 2 (bin_dist) : mean_all( BCEWithLogits(s, sigmoid(t)) ) * sum(a) / N
//                       (the reference multiplies the already averaged BCE scalar by the attention map
//                        and sums that per sample, depth_train.py:117-121 -- reproduced as written)
//
// t = teacher feature, s = student feature, both [N, C, H, W] logical (NHWC or NCHW memory, fp32 or
// bf16), a = attention map [N, H*W] fp32 broadcast over channels.  Forward = one streaming pass over
// both features (deterministic: per-block partials, combined in a fixed order by a one-block finish
// kernel that also emits the per-sample gradient scale); backward = one pass writing ds.
#include "b2_common.cuh"

namespace {

constexpr int kParts = B2_MIMIC_PARTS;
constexpr int kThreadsM = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int MODE>
__device__ __forceinline__ float mimic_term(float t, float s, float a) {
  if (MODE == 0) {
    const float d = (t - s) * a;
    return d * d;
  } else if (MODE == 1) {
    const float d = (sigmoidf_(t) - sigmoidf_(s)) * a;
    return d * d;
  } else {
    const float y = sigmoidf_(t);
    return fmaxf(s, 0.f) - s * y + log1pf(expf(-fabsf(s)));
  }
}

// d(loss)/ds up to the per-sample factor `scale`
template <int MODE>
__device__ __forceinline__ float mimic_grad(float t, float s, float a) {
  if (MODE == 0) return -(t - s) * a * a;
  if (MODE == 1) {
    const float ss = sigmoidf_(s);
    return -(sigmoidf_(t) - ss) * a * a * ss * (1.f - ss);
  }
  return sigmoidf_(s) - sigmoidf_(t);
}

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void ld(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const bf16* p, float* f) { load8(p, f); }
  static __device__ __forceinline__ void st(bf16* p, const float* f) { store8(p, f); }
};

// element e of sample n: NHWC -> pixel = e / C ; NCHW -> pixel = e % HW
template <typename T, int MODE, bool VEC>
__global__ void __launch_bounds__(kThreadsM)
mimic_reduce_kernel(const T* __restrict__ t, const T* __restrict__ s, const float* __restrict__ atten, int C, int HW,
                    int layout, float* __restrict__ partials) {
  __shared__ float red[kThreadsM / 32];
  const int n = blockIdx.y;
  const long long per = (long long)C * HW;
  const T* tn = t + (long long)n * per;
  const T* sn = s + (long long)n * per;
  const float* an = atten + (long long)n * HW;
  float acc = 0.f;
  if (VEC) {            // NHWC, C % Vec::N == 0: a vector never straddles a pixel
    constexpr int V = Vec<T>::N;
    const long long nv = per / V;
    for (long long v = blockIdx.x * (long long)kThreadsM + threadIdx.x; v < nv; v += (long long)gridDim.x * kThreadsM) {
      const long long e = v * V;
      float ft[V], fs[V];
      Vec<T>::ld(tn + e, ft);
      Vec<T>::ld(sn + e, fs);
      const float a = an[e / C];
#pragma unroll
      for (int j = 0; j < V; ++j) acc += mimic_term<MODE>(ft[j], fs[j], a);
    }
  } else {
    for (long long e = blockIdx.x * (long long)kThreadsM + threadIdx.x; e < per; e += (long long)gridDim.x * kThreadsM) {
      const float a = an[layout == 0 ? e / C : e % HW];
      acc += mimic_term<MODE>(to_f<T>(tn[e]), to_f<T>(sn[e]), a);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < kThreadsM / 32; ++w) tot += red[w];
    partials[(long long)n * kParts + blockIdx.x] = tot;
    for (int b = gridDim.x + blockIdx.x; b < kParts; b += gridDim.x) partials[(long long)n * kParts + b] = 0.f;
  }
}

// one block: combine the partials in a fixed order (fp64), emit the loss and the per-sample gradient scale
__global__ void mimic_finish_kernel(const float* __restrict__ partials, const float* __restrict__ atten, int N, int C,
                                    int HW, int mode, float* __restrict__ scale, float* __restrict__ loss) {
  __shared__ double sh[256];
  const int tid = threadIdx.x;
  if (mode < 2) {
    double mine = 0.0;
    for (int n = tid; n < N; n += blockDim.x) {
      double ss = 0.0;
      for (int b = 0; b < kParts; ++b) ss += (double)partials[(long long)n * kParts + b];
      const double nrm = sqrt(ss);
      mine += nrm;
      scale[n] = nrm > 0.0 ? (float)(1.0 / (nrm * (double)N)) : 0.f;
    }
    sh[tid] = mine;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int i = 0; i < (int)blockDim.x; ++i) tot += sh[i];
      *loss = (float)(tot / (double)N);
    }
  } else {
    double bsum = 0.0, asum = 0.0;
    for (long long i = tid; i < (long long)N * kParts; i += blockDim.x) bsum += (double)partials[i];
    for (long long i = tid; i < (long long)N * HW; i += blockDim.x) asum += (double)atten[i];
    sh[tid] = bsum;
    __syncthreads();
    double btot = 0.0;
    if (tid == 0) for (int i = 0; i < (int)blockDim.x; ++i) btot += sh[i];
    __syncthreads();
    sh[tid] = asum;
    __syncthreads();
    if (tid == 0) {
      double atot = 0.0;
      for (int i = 0; i < (int)blockDim.x; ++i) atot += sh[i];
      const double numel = (double)N * C * HW;
      *loss = (float)(btot / numel * (atot / (double)N));
      sh[0] = atot / (double)N / numel;
    }
    __syncthreads();
    const float sc = (float)sh[0];
    for (int n = tid; n < N; n += blockDim.x) scale[n] = sc;
  }
}

template <typename T, int MODE, bool VEC>
__global__ void __launch_bounds__(kThreadsM)
mimic_bwd_kernel(const T* __restrict__ t, const T* __restrict__ s, const float* __restrict__ atten,
                 const float* __restrict__ scale, const float* __restrict__ dloss, int C, int HW, int layout,
                 T* __restrict__ ds) {
  const int n = blockIdx.y;
  const long long per = (long long)C * HW;
  const T* tn = t + (long long)n * per;
  const T* sn = s + (long long)n * per;
  T* dn = ds + (long long)n * per;
  const float* an = atten + (long long)n * HW;
  const float k = scale[n] * (dloss ? dloss[0] : 1.f);
  if (VEC) {
    constexpr int V = Vec<T>::N;
    const long long nv = per / V;
    for (long long v = blockIdx.x * (long long)kThreadsM + threadIdx.x; v < nv; v += (long long)gridDim.x * kThreadsM) {
      const long long e = v * V;
      float ft[V], fs[V];
      Vec<T>::ld(tn + e, ft);
      Vec<T>::ld(sn + e, fs);
      const float a = an[e / C];
#pragma unroll
      for (int j = 0; j < V; ++j) fs[j] = k * mimic_grad<MODE>(ft[j], fs[j], a);
      Vec<T>::st(dn + e, fs);
    }
  } else {
    for (long long e = blockIdx.x * (long long)kThreadsM + threadIdx.x; e < per; e += (long long)gridDim.x * kThreadsM) {
      const float a = an[layout == 0 ? e / C : e % HW];
      dn[e] = from_f<T>(k * mimic_grad<MODE>(to_f<T>(tn[e]), to_f<T>(sn[e]), a));
    }
  }
}

int blocks_per_sample(long long per, int N) {
  long long want = (per / 8 + kThreadsM - 1) / kThreadsM;               // ~one vector per thread and pass
  long long cap = ((long long)b2_num_sms() * 8 + N - 1) / N;            // ~8 blocks per SM over all samples
  if (want > cap) want = cap;
  if (want > kParts) want = kParts;
  return (int)(want < 1 ? 1 : want);
}

template <typename T, bool VEC>
void launch_reduce(int mode, dim3 grid, cudaStream_t st, const void* t, const void* s, const float* atten, int C, int HW,
                   int layout, float* partials) {
  if (mode == 0)
    mimic_reduce_kernel<T, 0, VEC><<<grid, kThreadsM, 0, st>>>((const T*)t, (const T*)s, atten, C, HW, layout, partials);
  else if (mode == 1)
    mimic_reduce_kernel<T, 1, VEC><<<grid, kThreadsM, 0, st>>>((const T*)t, (const T*)s, atten, C, HW, layout, partials);
  else
    mimic_reduce_kernel<T, 2, VEC><<<grid, kThreadsM, 0, st>>>((const T*)t, (const T*)s, atten, C, HW, layout, partials);
}

template <typename T, bool VEC>
void launch_bwd(int mode, dim3 grid, cudaStream_t st, const void* t, const void* s, const float* atten,
                const float* scale, const float* dloss, int C, int HW, int layout, void* ds) {
  if (mode == 0)
    mimic_bwd_kernel<T, 0, VEC><<<grid, kThreadsM, 0, st>>>((const T*)t, (const T*)s, atten, scale, dloss, C, HW, layout, (T*)ds);
  else if (mode == 1)
    mimic_bwd_kernel<T, 1, VEC><<<grid, kThreadsM, 0, st>>>((const T*)t, (const T*)s, atten, scale, dloss, C, HW, layout, (T*)ds);
  else
    mimic_bwd_kernel<T, 2, VEC><<<grid, kThreadsM, 0, st>>>((const T*)t, (const T*)s, atten, scale, dloss, C, HW, layout, (T*)ds);
}

// radial attention map of utils.get_attention (utils.py:14-42): one block per sample
__global__ void attention_kernel(const float* __restrict__ coords, int J, int side_in, int side_out,
                                 float* __restrict__ out) {
  __shared__ float red[32];
  const int n = blockIdx.x, HW = side_out * side_out;
  const float* cn = coords + (long long)n * J * 2;
  float* on = out + (long long)n * HW;
  const float ratio = (float)side_in / (float)side_out;
  float vmax = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float cy = (float)(i / side_out), cx = (float)(i % side_out);
    float acc = 0.f;
    for (int j = 0; j < J; ++j) {
      const float dx = cx - cn[2 * j] / ratio, dy = cy - cn[2 * j + 1] / ratio;
      acc += expf(-(dx * dx + dy * dy) / 5.0f);
    }
    on[i] = acc;
    vmax = fmaxf(vmax, acc);
  }
  vmax = warp_max(vmax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = vmax;
  __syncthreads();
  float m = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  for (int i = threadIdx.x; i < HW; i += blockDim.x) on[i] = on[i] / m;
}

bool vec_ok(int C, int layout, int dtype) { return layout == 0 && C % (dtype == B2_BF16 ? 8 : 4) == 0; }

}  // namespace

extern "C" int b2_mimic_loss_fwd(const void* teach, const void* student, const float* atten, int32_t N, int32_t C,
                                 int32_t HW, int32_t layout, int32_t dtype, int32_t mode, float* partials,
                                 float* scale, float* loss, void* stream) {
  B2_REQUIRE(teach && student && atten && partials && scale && loss, B2_E_BADARG, "mimic_loss_fwd: null tensor");
  B2_REQUIRE(N > 0 && C > 0 && HW > 0 && (layout == 0 || layout == 1) && mode >= 0 && mode <= 2, B2_E_BADARG,
             "mimic_loss_fwd: bad argument (N=%d C=%d HW=%d layout=%d mode=%d)", N, C, HW, layout, mode);
  B2_REQUIRE(dtype == B2_F32 || dtype == B2_BF16, B2_E_UNSUPPORTED, "mimic_loss_fwd: dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(blocks_per_sample((long long)C * HW, N), N);
  const bool vec = vec_ok(C, layout, dtype);
  if (dtype == B2_F32) {
    if (vec) launch_reduce<float, true>(mode, grid, st, teach, student, atten, C, HW, layout, partials);
    else launch_reduce<float, false>(mode, grid, st, teach, student, atten, C, HW, layout, partials);
  } else {
    if (vec) launch_reduce<bf16, true>(mode, grid, st, teach, student, atten, C, HW, layout, partials);
    else launch_reduce<bf16, false>(mode, grid, st, teach, student, atten, C, HW, layout, partials);
  }
  B2_LAUNCH_CHECK("mimic_reduce");
  mimic_finish_kernel<<<1, 256, 0, st>>>(partials, atten, N, C, HW, mode, scale, loss);
  B2_LAUNCH_CHECK("mimic_finish");
  return B2_OK;
}

extern "C" int b2_mimic_loss_bwd(const void* teach, const void* student, const float* atten, const float* scale,
                                 const float* dloss, int32_t N, int32_t C, int32_t HW, int32_t layout,
                                 int32_t dtype, int32_t mode, void* dstudent, void* stream) {
  B2_REQUIRE(teach && student && atten && scale && dstudent, B2_E_BADARG, "mimic_loss_bwd: null tensor");
  B2_REQUIRE(N > 0 && C > 0 && HW > 0 && (layout == 0 || layout == 1) && mode >= 0 && mode <= 2, B2_E_BADARG,
             "mimic_loss_bwd: bad argument (N=%d C=%d HW=%d layout=%d mode=%d)", N, C, HW, layout, mode);
  B2_REQUIRE(dtype == B2_F32 || dtype == B2_BF16, B2_E_UNSUPPORTED, "mimic_loss_bwd: dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(blocks_per_sample((long long)C * HW, N), N);
  const bool vec = vec_ok(C, layout, dtype);
  if (dtype == B2_F32) {
    if (vec) launch_bwd<float, true>(mode, grid, st, teach, student, atten, scale, dloss, C, HW, layout, dstudent);
    else launch_bwd<float, false>(mode, grid, st, teach, student, atten, scale, dloss, C, HW, layout, dstudent);
  } else {
    if (vec) launch_bwd<bf16, true>(mode, grid, st, teach, student, atten, scale, dloss, C, HW, layout, dstudent);
    else launch_bwd<bf16, false>(mode, grid, st, teach, student, atten, scale, dloss, C, HW, layout, dstudent);
  }
  B2_LAUNCH_CHECK("mimic_bwd");
  return B2_OK;
}

extern "C" int b2_attention_map(const float* image_coords, int32_t N, int32_t J, int32_t side_in, int32_t side_out,
                                float* out, void* stream) {
  B2_REQUIRE(image_coords && out && N > 0 && J > 0 && side_in > 0 && side_out > 0, B2_E_BADARG,
             "attention_map: bad argument");
  attention_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(image_coords, J, side_in, side_out, out);
  B2_LAUNCH_CHECK("attention_map");
  return B2_OK;
}
