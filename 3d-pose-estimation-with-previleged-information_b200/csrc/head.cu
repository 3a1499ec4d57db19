// Volumetric heat-map head: softmax over the H*W*D voxels of every (sample, joint) followed by
// the soft-argmax ("integral regression") expectation, fused into ONE pass over the logits
// (utils.to_heatmap + utils.decode, utils.py:154-194 of the reference), its analytic backward,
// the unfused pieces kept for API compatibility, and the root-relative masked loss
// (depth_train.py:397-405).
//
// Forward: one thread-block CLUSTER of kHeadSplit CTAs per sample (64 samples alone leave 84 of the 148 SMs idle),
// every CTA owns a contiguous slice of the pixels.  Every thread owns one logit channel (d, j) over a subset of its
// CTA's pixels and keeps an online-softmax state (max, sum e, sum e*gx, sum e*gy); states are merged per channel
// inside the CTA, then across the cluster through distributed shared memory by the cluster's first CTA, then over d
// for every joint.  The logits are read exactly once, coalesced in either layout (NHWC: consecutive threads =
// consecutive channels of a pixel; NCHW: consecutive lanes = consecutive pixels of a channel plane).
#include <cooperative_groups.h>

#include "b2_common.cuh"

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ float grid_coord(int i, int n) {
  // torch.linspace(0, 2, n)[i]  (utils.py:182-184): symmetric evaluation like ATen
  if (n <= 1) return 0.f;
  float step = 2.f / (float)(n - 1);
  return (i < n / 2) ? (float)i * step : 2.f - (float)(n - 1 - i) * step;
}

constexpr int kHeadSplit = 8;     // CTAs per sample (portable cluster size)

struct St {
  float m, s, a, b;   // running max, sum e, sum e*gx(w), sum e*gy(h)
};
__device__ __forceinline__ void st_push(St& t, float x, float gx, float gy) {
  float mn = fmaxf(t.m, x);
  float sc = (t.m == -INFINITY) ? 0.f : __expf(t.m - mn);
  float e = __expf(x - mn);
  t.s = t.s * sc + e;
  t.a = t.a * sc + e * gx;
  t.b = t.b * sc + e * gy;
  t.m = mn;
}
__device__ __forceinline__ St st_merge(const St& p, const St& q) {
  St r;
  r.m = fmaxf(p.m, q.m);
  float sp = (p.m == -INFINITY) ? 0.f : __expf(p.m - r.m);
  float sq = (q.m == -INFINITY) ? 0.f : __expf(q.m - r.m);
  r.s = p.s * sp + q.s * sq;
  r.a = p.a * sp + q.a * sq;
  r.b = p.b * sp + q.b * sq;
  return r;
}

// dynamic smem: St ch[CH]; float jm[J], js[J]; St part[blockDim] (NHWC only)
template <typename T, int LAYOUT, bool WRITE_HEAT>
__global__ void __launch_bounds__(1024) head_fwd_kernel(const T* __restrict__ logits, int J, int D, int H, int W,
                                                        float range, float* __restrict__ coords,
                                                        float* __restrict__ vmax, float* __restrict__ vsum,
                                                        float* __restrict__ heat) {
  extern __shared__ float4 smem4[];
  const int CH = D * J, HW = H * W;
  St* ch = reinterpret_cast<St*>(smem4);
  float* jm = reinterpret_cast<float*>(ch + CH);
  float* js = jm + J;
  const int n = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const T* base = logits + (long long)n * CH * HW;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), nsplit = (int)cluster.num_blocks();
  const int per_cta = (HW + nsplit - 1) / nsplit;
  const int p_lo = min(HW, rank * per_cta), p_hi = min(HW, p_lo + per_cta);      // this CTA's pixels

  if (LAYOUT == 0) {
    St* part = reinterpret_cast<St*>(js + J + ((4 - ((2 * J) & 3)) & 3));
    const int PL = nt / CH > 0 ? nt / CH : 1;      // pixel lanes per channel
    St t = {-INFINITY, 0.f, 0.f, 0.f};
    // when CH > blockDim every thread loops over several channels (c += nt)
    for (int c0 = 0; c0 < CH; c0 += nt) {
      int c = c0 + tid % (CH < nt ? CH : nt), pl = (CH < nt) ? tid / CH : 0;
      bool act = c < CH && pl < PL;
      t.m = -INFINITY; t.s = t.a = t.b = 0.f;
      if (act) {
        for (int p = p_lo + pl; p < p_hi; p += PL) {
          float x = to_f(base[(long long)p * CH + c]);
          st_push(t, x, grid_coord(p % W, W), grid_coord(p / W, H));
        }
      }
      if (CH < nt) {
        part[tid] = t;
        __syncthreads();
        if (act && pl == 0) {
          for (int q = 1; q < PL; ++q) t = st_merge(t, part[q * CH + c]);
          ch[c] = t;
        }
      } else if (act) {
        ch[c] = t;
      }
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int c = warp; c < CH; c += nw) {
      St t = {-INFINITY, 0.f, 0.f, 0.f};
      const T* plane = base + (long long)c * HW;
      for (int p = p_lo + lane; p < p_hi; p += 32) st_push(t, to_f(plane[p]), grid_coord(p % W, W), grid_coord(p / W, H));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        St q;
        q.m = __shfl_xor_sync(0xffffffffu, t.m, o);
        q.s = __shfl_xor_sync(0xffffffffu, t.s, o);
        q.a = __shfl_xor_sync(0xffffffffu, t.a, o);
        q.b = __shfl_xor_sync(0xffffffffu, t.b, o);
        t = st_merge(t, q);
      }
      if (lane == 0) ch[c] = t;
    }
  }
  // merge the per-CTA channel states across the cluster (rank 0 reads its peers' shared memory)
  cluster.sync();
  if (rank == 0) {
    for (int c = tid; c < CH; c += nt) {
      St t = ch[c];
      for (int r = 1; r < nsplit; ++r) t = st_merge(t, cluster.map_shared_rank(ch, r)[c]);
      ch[c] = t;
    }
  }
  __syncthreads();
  for (int j = tid; rank == 0 && j < J; j += nt) {
    St t = ch[j];
    float m = t.m;
    for (int d = 1; d < D; ++d) m = fmaxf(m, ch[d * J + j].m);
    float S = 0.f, X = 0.f, Y = 0.f, Z = 0.f;
    for (int d = 0; d < D; ++d) {
      St q = ch[d * J + j];
      float sc = __expf(q.m - m);
      S += q.s * sc;
      X += q.a * sc;
      Y += q.b * sc;
      Z += q.s * sc * grid_coord(d, D);
    }
    float inv = 1.f / S;
    float* o = coords + ((long long)n * J + j) * 3;
    o[0] = X * inv * range;
    o[1] = Y * inv * range;
    o[2] = Z * inv * range;
    if (vmax) vmax[(long long)n * J + j] = m;
    if (vsum) vsum[(long long)n * J + j] = S;
    jm[j] = m;
    js[j] = inv;
  }
  cluster.sync();            // rank 0 is done with its peers' shared memory; jm / js are final
  if (WRITE_HEAT) {
    const float* rjm = cluster.map_shared_rank(jm, 0);
    const float* rjs = cluster.map_shared_rank(js, 0);
    const long long total = (long long)CH * (p_hi - p_lo);
    for (long long i = tid; i < total; i += nt) {
      int c, p;
      if (LAYOUT == 0) { p = p_lo + (int)(i / CH); c = (int)(i % CH); }
      else { c = (int)(i / (p_hi - p_lo)); p = p_lo + (int)(i % (p_hi - p_lo)); }
      int j = c % J, d = c / J;
      const long long src = (LAYOUT == 0) ? (long long)p * CH + c : (long long)c * HW + p;
      float e = __expf(to_f(base[src]) - rjm[j]) * rjs[j];
      heat[(((long long)n * J + j) * HW + p) * D + d] = e;
    }
    cluster.sync();          // peers read rank 0's jm / js until here
  }
}

// dlogit = p * (u.g - u.c)
template <typename T, int LAYOUT>
__global__ void head_bwd_kernel(const T* __restrict__ logits, const float* __restrict__ dcoords,
                                const float* __restrict__ coords, const float* __restrict__ vmax,
                                const float* __restrict__ vsum, int N, int J, int D, int H, int W, float range,
                                T* __restrict__ dlogits) {
  const int CH = D * J, HW = H * W;
  const long long per = (long long)CH * HW, total = per * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i / per);
    long long r = i - (long long)n * per;
    int c, p;
    if (LAYOUT == 0) { p = (int)(r / CH); c = (int)(r - (long long)p * CH); }
    else { c = (int)(r / HW); p = (int)(r - (long long)c * HW); }
    int j = c % J, d = c / J, h = p / W, w = p - h * W;
    long long nj = (long long)n * J + j;
    const float* u = dcoords + nj * 3;
    const float* cc = coords + nj * 3;
    float prob = __expf(to_f(logits[i]) - vmax[nj]) / vsum[nj];
    float ug = (u[0] * grid_coord(w, W) + u[1] * grid_coord(h, H) + u[2] * grid_coord(d, D)) * range;
    float uc = u[0] * cc[0] + u[1] * cc[1] + u[2] * cc[2];
    dlogits[i] = from_f<T>(prob * (ug - uc));
  }
}

// NHWC with CH % 8 == 0: one thread = 8 consecutive channels of one pixel (16-byte bf16 / 2 x 16-byte fp32 accesses,
// 32-bit index arithmetic)
template <typename T>
__global__ void head_bwd_vec_kernel(const T* __restrict__ logits, const float* __restrict__ dcoords,
                                    const float* __restrict__ coords, const float* __restrict__ vmax,
                                    const float* __restrict__ vsum, int N, int J, int D, int H, int W, float range,
                                    T* __restrict__ dlogits) {
  const int CH = D * J, HW = H * W, CV = CH >> 3;
  const int total = N * HW * CV;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % CV, np = i / CV, p = np % HW, n = np / HW;
    const int h = p / W, w = p - h * W;
    const float gx = grid_coord(w, W), gy = grid_coord(h, H);
    const long long off = (long long)np * CH + v * 8;
    float x[8], o[8];
    if (sizeof(T) == 2) {
      load8(reinterpret_cast<const bf16*>(logits) + off, x);
    } else {
      const float4 a = load4(reinterpret_cast<const float*>(logits) + off), b = load4(reinterpret_cast<const float*>(logits) + off + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e, d = c / J, j = c - d * J;
      const int nj = n * J + j;
      const float* u = dcoords + (long long)nj * 3;
      const float* cc = coords + (long long)nj * 3;
      const float prob = __expf(x[e] - vmax[nj]) / vsum[nj];
      const float ug = (u[0] * gx + u[1] * gy + u[2] * grid_coord(d, D)) * range;
      const float uc = u[0] * cc[0] + u[1] * cc[1] + u[2] * cc[2];
      o[e] = prob * (ug - uc);
    }
    if (sizeof(T) == 2) {
      store8(reinterpret_cast<bf16*>(dlogits) + off, o);
    } else {
      store4(reinterpret_cast<float*>(dlogits) + off, make_float4(o[0], o[1], o[2], o[3]));
      store4(reinterpret_cast<float*>(dlogits) + off + 4, make_float4(o[4], o[5], o[6], o[7]));
    }
  }
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
  return r;
}

// utils.decode on a materialised heat-map [N,J,H,W,D]; one CTA per (n, j)
__global__ void decode_kernel(const float* __restrict__ heat, int D, int H, int W, float range,
                              float* __restrict__ coords) {
  __shared__ float sh[32];
  const long long nj = blockIdx.x;
  const int V = H * W * D;
  const float* hp = heat + nj * V;
  float x = 0.f, y = 0.f, z = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    int d = i % D, p = i / D, w = p % W, h = p / W;
    float v = hp[i];
    x += v * grid_coord(w, W);
    y += v * grid_coord(h, H);
    z += v * grid_coord(d, D);
  }
  x = block_sum(x, sh);
  y = block_sum(y, sh);
  z = block_sum(z, sh);
  if (threadIdx.x == 0) {
    coords[nj * 3 + 0] = x * range;
    coords[nj * 3 + 1] = y * range;
    coords[nj * 3 + 2] = z * range;
  }
}

__global__ void decode_bwd_kernel(const float* __restrict__ dcoords, long long NJ, int D, int H, int W, float range,
                                  float* __restrict__ dheat) {
  const int V = H * W * D;
  const long long total = NJ * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long nj = i / V;
    int r = (int)(i - nj * V);
    int d = r % D, p = r / D, w = p % W, h = p / W;
    const float* u = dcoords + nj * 3;
    dheat[i] = (u[0] * grid_coord(w, W) + u[1] * grid_coord(h, H) + u[2] * grid_coord(d, D)) * range;
  }
}

// dlogits = p * (dp - sum p*dp), written back in the logits layout; one CTA per (n, j)
template <typename T, int LAYOUT>
__global__ void softmax_bwd_kernel(const float* __restrict__ heat, const float* __restrict__ dheat, int J, int D,
                                   int H, int W, T* __restrict__ dlogits) {
  __shared__ float sh[32];
  const long long nj = blockIdx.x;
  const int n = (int)(nj / J), j = (int)(nj % J);
  const int HW = H * W, V = HW * D, CH = D * J;
  const float* hp = heat + nj * V;
  const float* gp = dheat + nj * V;
  float dot = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) dot += hp[i] * gp[i];
  dot = block_sum(dot, sh);
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    int d = i % D, p = i / D;
    int c = d * J + j;
    long long off = (LAYOUT == 0) ? ((long long)n * HW + p) * CH + c : ((long long)n * CH + c) * HW + p;
    dlogits[off] = from_f<T>(hp[i] * (gp[i] - dot));
  }
}

// spec = coords - coords[key] + true[key]; masked mean loss of (spec - true)/div; analytic gradient.
__global__ void pose_loss_kernel(const float* __restrict__ coords, const float* __restrict__ true_cam,
                                 const uint8_t* __restrict__ valid, int N, int J, int key, float div, int crit,
                                 float* __restrict__ loss, float* __restrict__ spec_cam,
                                 float* __restrict__ dcoords) {
  __shared__ float sh[32];
  const int NJ = N * J;
  float cnt = 0.f;
  for (int i = threadIdx.x; i < NJ; i += blockDim.x) cnt += valid[i] ? 3.f : 0.f;
  cnt = block_sum(cnt, sh);
  const float invc = cnt > 0.f ? 1.f / cnt : 0.f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < NJ * 3; i += blockDim.x) {
    int a = i % 3, nj = i / 3, n = nj / J;
    long long kb = ((long long)n * J + key) * 3 + a;
    float spec = coords[i] - coords[kb] + true_cam[kb];
    if (spec_cam) spec_cam[i] = spec;
    float g = 0.f;
    if (valid[nj]) {
      float df = spec / div - true_cam[i] / div;
      float ad = fabsf(df);
      if (crit == 0) { acc += ad < 1.f ? 0.5f * df * df : ad - 0.5f; g = ad < 1.f ? df : (df > 0.f ? 1.f : -1.f); }
      else if (crit == 1) { acc += ad; g = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f); }
      else { acc += df * df; g = 2.f * df; }
      g = g / div * invc;
    }
    if (dcoords) dcoords[i] = g;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) loss[0] = cnt > 0.f ? acc * invc : NAN;
  if (!dcoords) return;
  __syncthreads();
  // root joint receives minus the sum over joints (spec depends on coords[key] with weight -1)
  for (int i = threadIdx.x; i < N * 3; i += blockDim.x) {
    int n = i / 3, a = i % 3;
    float s = 0.f;
    for (int j = 0; j < J; ++j) s += dcoords[((long long)n * J + j) * 3 + a];
    dcoords[((long long)n * J + key) * 3 + a] -= s;
  }
}

template <typename T, int LAYOUT>
int launch_head_fwd(const void* logits, int N, int J, int D, int H, int W, float range, float* coords, float* vmax,
                    float* vsum, float* heat, cudaStream_t st) {
  const int CH = D * J;
  int nt = 1024;
  if (LAYOUT == 0 && CH < 1024) {
    // keep whole pixel lanes: nt = CH * PL, rounded down to a warp multiple is not required
    int PL = 1024 / CH;
    nt = ((CH * PL + 31) / 32) * 32;
    if (nt > 1024) nt = 1024;
  }
  size_t sh = sizeof(float4) * CH + sizeof(float) * (2 * J + 4) + (LAYOUT == 0 ? sizeof(float4) * nt : 0);
  B2_REQUIRE(sh <= 200 * 1024, B2_E_UNSUPPORTED, "head_fwd: D*J=%d too large", CH);
  auto kern = heat ? head_fwd_kernel<T, LAYOUT, true> : head_fwd_kernel<T, LAYOUT, false>;
  if (sh > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kHeadSplit, N, 1);
  cfg.blockDim = dim3(nt, 1, 1);
  cfg.dynamicSmemBytes = sh;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kHeadSplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, (const T*)logits, J, D, H, W, range, coords, vmax, vsum, heat);
  B2_REQUIRE(le == cudaSuccess, B2_E_LAUNCH, "head_fwd: launch failed: %s", cudaGetErrorString(le));
  B2_LAUNCH_CHECK("head_fwd");
  return B2_OK;
}

inline int ew_grid(long long total) {
  long long want = (total + 255) / 256, cap = (long long)b2_num_sms() * 8;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

int check_head(int N, int J, int D, int H, int W, int layout, int dtype) {
  B2_REQUIRE(N > 0 && J > 0 && D > 0 && H > 0 && W > 0, B2_E_BADARG, "head: non-positive dimension");
  B2_REQUIRE(layout == 0 || layout == 1, B2_E_BADARG, "head: layout %d", layout);
  B2_REQUIRE(dtype == B2_F32 || dtype == B2_BF16, B2_E_UNSUPPORTED, "head: dtype %d", dtype);
  return B2_OK;
}

}  // namespace

#define HEAD_DISPATCH(FN, ...)                                                     \
  (dtype == B2_F32 ? (layout == 0 ? FN<float, 0>(__VA_ARGS__) : FN<float, 1>(__VA_ARGS__)) \
                   : (layout == 0 ? FN<bf16, 0>(__VA_ARGS__) : FN<bf16, 1>(__VA_ARGS__)))

extern "C" int b2_head_fwd(const void* logits, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                           int32_t layout, int32_t dtype, float depth_range, float* coords, float* vmax,
                           float* vsum, void* stream) {
  int rc = check_head(N, J, D, H, W, layout, dtype);
  if (rc) return rc;
  B2_REQUIRE(logits && coords, B2_E_BADARG, "head_fwd: null tensor");
  return HEAD_DISPATCH(launch_head_fwd, logits, N, J, D, H, W, depth_range, coords, vmax, vsum, nullptr,
                       (cudaStream_t)stream);
}

extern "C" int b2_heatmap_softmax(const void* logits, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                                  int32_t layout, int32_t dtype, float* heat, float* coords_scratch, void* stream) {
  int rc = check_head(N, J, D, H, W, layout, dtype);
  if (rc) return rc;
  B2_REQUIRE(logits && heat && coords_scratch, B2_E_BADARG, "heatmap_softmax: null tensor");
  return HEAD_DISPATCH(launch_head_fwd, logits, N, J, D, H, W, 1.f, coords_scratch, nullptr, nullptr, heat,
                       (cudaStream_t)stream);
}

template <typename T, int LAYOUT>
static int launch_head_bwd(const void* logits, const float* dcoords, const float* coords, const float* vmax,
                           const float* vsum, int N, int J, int D, int H, int W, float range, void* dlogits,
                           cudaStream_t st) {
  long long total = (long long)N * D * J * H * W;
  if (LAYOUT == 0 && (D * J) % 8 == 0 && total < (1LL << 31) &&
      (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) {
    head_bwd_vec_kernel<T><<<ew_grid(total / 8), 256, 0, st>>>((const T*)logits, dcoords, coords, vmax, vsum, N, J, D, H, W,
                                                              range, (T*)dlogits);
    B2_LAUNCH_CHECK("head_bwd");
    return B2_OK;
  }
  head_bwd_kernel<T, LAYOUT><<<ew_grid(total), 256, 0, st>>>((const T*)logits, dcoords, coords, vmax, vsum, N, J, D,
                                                           H, W, range, (T*)dlogits);
  B2_LAUNCH_CHECK("head_bwd");
  return B2_OK;
}

extern "C" int b2_head_bwd(const void* logits, const float* dcoords, const float* coords, const float* vmax,
                           const float* vsum, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                           int32_t layout, int32_t dtype, float depth_range, void* dlogits, void* stream) {
  int rc = check_head(N, J, D, H, W, layout, dtype);
  if (rc) return rc;
  B2_REQUIRE(logits && dcoords && coords && vmax && vsum && dlogits, B2_E_BADARG, "head_bwd: null tensor");
  return HEAD_DISPATCH(launch_head_bwd, logits, dcoords, coords, vmax, vsum, N, J, D, H, W, depth_range, dlogits,
                       (cudaStream_t)stream);
}

extern "C" int b2_heatmap_decode(const float* heat, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                                 float depth_range, float* coords, void* stream) {
  int rc = check_head(N, J, D, H, W, 0, B2_F32);
  if (rc) return rc;
  B2_REQUIRE(heat && coords, B2_E_BADARG, "heatmap_decode: null tensor");
  decode_kernel<<<N * J, 256, 0, (cudaStream_t)stream>>>(heat, D, H, W, depth_range, coords);
  B2_LAUNCH_CHECK("heatmap_decode");
  return B2_OK;
}

extern "C" int b2_heatmap_decode_bwd(const float* dcoords, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                                     float depth_range, float* dheat, void* stream) {
  int rc = check_head(N, J, D, H, W, 0, B2_F32);
  if (rc) return rc;
  B2_REQUIRE(dcoords && dheat, B2_E_BADARG, "heatmap_decode_bwd: null tensor");
  long long total = (long long)N * J * D * H * W;
  decode_bwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(dcoords, (long long)N * J, D, H, W, depth_range,
                                                                    dheat);
  B2_LAUNCH_CHECK("heatmap_decode_bwd");
  return B2_OK;
}

template <typename T, int LAYOUT>
static int launch_softmax_bwd(const float* heat, const float* dheat, int N, int J, int D, int H, int W, void* dlogits,
                              cudaStream_t st) {
  softmax_bwd_kernel<T, LAYOUT><<<N * J, 256, 0, st>>>(heat, dheat, J, D, H, W, (T*)dlogits);
  B2_LAUNCH_CHECK("heatmap_softmax_bwd");
  return B2_OK;
}

extern "C" int b2_heatmap_softmax_bwd(const float* heat, const float* dheat, int32_t N, int32_t J, int32_t D,
                                      int32_t H, int32_t W, int32_t layout, int32_t dtype, void* dlogits,
                                      void* stream) {
  int rc = check_head(N, J, D, H, W, layout, dtype);
  if (rc) return rc;
  B2_REQUIRE(heat && dheat && dlogits, B2_E_BADARG, "heatmap_softmax_bwd: null tensor");
  return HEAD_DISPATCH(launch_softmax_bwd, heat, dheat, N, J, D, H, W, dlogits, (cudaStream_t)stream);
}

extern "C" int b2_pose_loss(const float* coords, const float* true_cam, const uint8_t* valid, int32_t N, int32_t J,
                            int32_t key_index, float loss_div, int32_t criterion, float* loss, float* spec_cam,
                            float* dcoords, void* stream) {
  B2_REQUIRE(coords && true_cam && valid && loss && N > 0 && J > 0, B2_E_BADARG, "pose_loss: bad argument");
  B2_REQUIRE(key_index >= 0 && key_index < J, B2_E_BADARG, "pose_loss: key_index %d out of range", key_index);
  B2_REQUIRE(criterion >= 0 && criterion <= 2, B2_E_BADARG, "pose_loss: criterion %d", criterion);
  pose_loss_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(coords, true_cam, valid, N, J, key_index, loss_div, criterion,
                                                       loss, spec_cam, dcoords);
  B2_LAUNCH_CHECK("pose_loss");
  return B2_OK;
}

// ---------------------------------------------------------------- evaluation metrics
// utils.analyze + utils.statistics (utils.py:197-262) and the back-rotation einsum of the test loops
// (depth_train.py:522-523) for one batch, accumulated into acc[10] (double):
//   [0] valid joints  [1] sum dist  [2] #(dist <= rough)  [3] sum max(0, 1 - dist / rough)
//   [4] solid [5] close [6] depth [7] jitter [8] switch [9] fail      (the sequential elimination of :210-221)
// Epoch totals of these give parse_epoch's batch-size-weighted means exactly (utils.py:224-231).
namespace {
__global__ void pose_metrics_kernel(const float* __restrict__ spec, const float* __restrict__ truth,
                                    const uint8_t* __restrict__ valid, const float* __restrict__ rot,
                                    const int* __restrict__ mirror, int N, int J, float t_solid, float t_close,
                                    float t_rough, double* __restrict__ acc) {
  __shared__ double red[10][8];
  double v[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) v[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * J; i += gridDim.x * blockDim.x) {
    if (!valid[i]) continue;
    const int n = i / J, j = i - n * J;
    const int jm = mirror ? mirror[j] : j;
    const float* s = spec + (long long)i * 3;
    const float* t = truth + (long long)i * 3;
    const float* tm = truth + ((long long)n * J + jm) * 3;
    float ds[3] = {s[0] - t[0], s[1] - t[1], s[2] - t[2]};
    float df[3] = {s[0] - tm[0], s[1] - tm[1], s[2] - tm[2]};
    if (rot) {                                   // R (s - t): distances are rotation invariant, the tangent part is not
      const float* r = rot + (long long)n * 9;
      float a[3], b[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        a[q] = r[q * 3] * ds[0] + r[q * 3 + 1] * ds[1] + r[q * 3 + 2] * ds[2];
        b[q] = r[q * 3] * df[0] + r[q * 3 + 1] * df[1] + r[q * 3 + 2] * df[2];
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) { ds[q] = a[q]; df[q] = b[q]; }
    }
    const float basic = sqrtf(ds[0] * ds[0] + ds[1] * ds[1] + ds[2] * ds[2]);
    const float flip = sqrtf(df[0] * df[0] + df[1] * df[1] + df[2] * df[2]);
    const float tangent = sqrtf(ds[0] * ds[0] + ds[1] * ds[1]);
    v[0] += 1.0;
    v[1] += (double)basic;
    v[2] += (basic / t_rough <= 1.0f) ? 1.0 : 0.0;
    v[3] += (double)fmaxf(0.f, 1.f - basic / t_rough);
    int cat;
    if (basic <= t_solid) cat = 4;
    else if (basic <= t_close) cat = 5;
    else if (tangent <= t_close) cat = 6;
    else if (basic <= t_rough) cat = 7;
    else if (flip <= t_rough) cat = 8;
    else cat = 9;
#pragma unroll
    for (int k = 4; k < 10; ++k) v[k] += (cat == k) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double x = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += red[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, x);
  }
}
}  // namespace

extern "C" int b2_pose_metrics(const float* spec_cam, const float* true_cam, const uint8_t* valid,
                               const float* back_rotate, const int32_t* mirror, int32_t N, int32_t J, float t_solid,
                               float t_close, float t_rough, double* acc, void* stream) {
  B2_REQUIRE(spec_cam && true_cam && valid && acc && N > 0 && J > 0, B2_E_BADARG, "pose_metrics: bad argument");
  B2_REQUIRE(t_rough > 0.f, B2_E_BADARG, "pose_metrics: the rough threshold must be positive");
  int blocks = (N * J + 255) / 256;
  if (blocks > 64) blocks = 64;
  pose_metrics_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(spec_cam, true_cam, valid, back_rotate, mirror, N, J,
                                                             t_solid, t_close, t_rough, acc);
  B2_LAUNCH_CHECK("pose_metrics");
  return B2_OK;
}
