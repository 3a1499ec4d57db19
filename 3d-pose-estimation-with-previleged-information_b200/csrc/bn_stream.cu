// BatchNorm stream kernels with bulk-async (TMA 1-D) staging for bf16 activations.
//
// The register-resident versions in bn.cu keep only ~2 x 16 B per input stream in flight per thread
// and top out at 2.3-3.8 TB/s on the backward kernels (few resident blocks because of the
// per-channel constants).  Here one elected thread streams contiguous 8 KB chunks of every input
// tensor into a 3-stage shared-memory ring with cp.async.bulk + mbarrier, so ~48-72 KB per block
// are in flight independent of register pressure; all threads then read 16-byte vectors from shared
// memory, compute, and write the outputs straight to global memory (coalesced 16-byte stores).
// Same math, same slot-partial reductions as bn.cu; used when C is a power of two in [64, 2048].
#include <cstdlib>

#include "b2_common.cuh"

namespace {

constexpr int kChunkBytes = 8192;            // per input stream and stage
constexpr int kChunkVecs = kChunkBytes / 16; // 16-byte vectors (8 bf16) per chunk
constexpr int kStagesS = 3;
constexpr int kThreadsS = 256;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void s_mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = s_u32(bar);
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if ((it & 1023) == 1023) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void s_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

__device__ __forceinline__ void s_bulk_load_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)), "l"(pol)
               : "memory");
}
// L2 eviction priority of the streamed inputs (B2POSE_BN_L2HINT=1).  The same tensors are read twice back to back --
// statistics / gradient sums first, then the apply pass -- and most of them are smaller than the 126 MB L2: the first
// pass asks the L2 to keep its lines (evict_last), the second pass marks them dead (evict_first).
__device__ __forceinline__ uint64_t l2_policy(int hint) {
  uint64_t pol = 0;
  if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// Chunk order of the streaming passes (B2POSE_BN_ORDER, bit mask: 1 forward apply, 2 backward reduce, 4 backward
// apply run from the END of the tensor to its start).  A pass that starts where its producer just stopped finds the
// most recently written (or read) lines still in the 126 MB L2: the convolution wrote the end of y last, the dgrad
// wrote the end of dz last, and the backward apply pass re-reads what the reduce pass read last.
// Measured in the step (partial_fusionnet ResNet-50, batch 64, one box, 30 steps each): 0: 14.24 ms, 1: 14.19, 2: 14.24,
// 4: 14.26, 3: 14.15, 6: 14.29, 7: 14.18 -- a small effect (the big activations exceed the L2 several times over);
// default 3.
inline int bn_order_mode() {
  static const int v = getenv("B2POSE_BN_ORDER") ? atoi(getenv("B2POSE_BN_ORDER")) : 3;
  return v;
}
inline int l2_hint_mode() {
  static const int v = getenv("B2POSE_BN_L2HINT") ? atoi(getenv("B2POSE_BN_L2HINT")) : 0;
  return v;
}

// Ring of kStagesS stages, NIN input streams each.  `issue(c)` is called by thread 0.
template <int NIN>
struct Ring {
  uint8_t* buf;      // [kStagesS][NIN][kChunkBytes]
  uint64_t* full;    // [kStagesS]
  const uint8_t* src[NIN];
  long long total_bytes;
  uint64_t policy = 0;             // L2 cache-hint policy of the loads (0: none)
  const uint8_t* bits = nullptr;   // optional 1-bit-per-element side stream (ReLU gate): kChunkBytes / 16 bytes per chunk
  uint8_t* bits_buf = nullptr;     // [kStagesS][kChunkBytes / 16]
  long long last_chunk = -1;       // >= 0: the pass runs backwards, logical chunk c is physical chunk last_chunk - c

  __device__ __forceinline__ long long phys(long long chunk) const { return last_chunk >= 0 ? last_chunk - chunk : chunk; }
  __device__ __forceinline__ void issue(int stage, long long chunk) {
    const long long off = phys(chunk) * kChunkBytes;
    long long rem = total_bytes - off;
    const uint32_t bytes = (uint32_t)(rem < kChunkBytes ? rem : kChunkBytes);
    const uint32_t gb = bits ? ((bytes >> 4) + 15u) & ~15u : 0u;       // bulk copies move multiples of 16 bytes
    s_mbar_expect(&full[stage], bytes * NIN + gb);
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
      if (policy) s_bulk_load_hint(buf + ((size_t)stage * NIN + i) * kChunkBytes, src[i] + off, bytes, &full[stage], policy);
      else s_bulk_load(buf + ((size_t)stage * NIN + i) * kChunkBytes, src[i] + off, bytes, &full[stage]);
    }
    if (bits) s_bulk_load(bits_buf + (size_t)stage * (kChunkBytes / 16), bits + (off >> 4), gb, &full[stage]);
  }
  __device__ __forceinline__ const bf16* data(int stage, int i) const {
    return reinterpret_cast<const bf16*>(buf + ((size_t)stage * NIN + i) * kChunkBytes);
  }
};

template <int NIN>
__device__ __forceinline__ void ring_setup(Ring<NIN>& r, uint8_t* smem, long long total_bytes) {
  r.buf = smem;
  r.full = reinterpret_cast<uint64_t*>(smem + (size_t)kStagesS * NIN * kChunkBytes);
  r.total_bytes = total_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesS; ++s) s_mbar_init(&r.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();            // only shared memory was touched so far
}

// block-level reduction of per-thread partials (8 channels x 2 quantities) that share a channel
// vector across the thread groups of the block; writes this block's slot (see bn.cu)
//
// totals != 0: `partials` is ONE pre-zeroed float[2C] vector and every block adds its sums with
// red.global.add.f32 (2C reductions per block); the consumers derive mean / invstd (or use the gradient
// sums) in their prologue, so no finalize kernel sits between producer and consumer.
__device__ __forceinline__ void slot_reduce(const float* s, const float* q, float* __restrict__ partials, int C,
                                            int cv, int CV, float* red, int totals) {
  // red: [2][kThreadsS][8]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 8 + j] = s[j];
    red[(kThreadsS + threadIdx.x) * 8 + j] = q[j];
  }
  __syncthreads();
  float* slot = partials + (totals ? (size_t)0 : (size_t)blockIdx.x * 2 * C);
  // thread t < 2*C handles one output value: quantity = t / C, channel = t % C
  for (int o = threadIdx.x; o < 2 * C; o += kThreadsS) {
    const int qn = o / C, c = o - qn * C, v = c >> 3, j = c & 7;
    float acc = 0.f;
    for (int t = v; t < kThreadsS; t += CV) acc += red[(qn * kThreadsS + t) * 8 + j];
    if (totals) atomicAdd(slot + o, acc);
    else slot[o] = acc;
  }
  if (!totals)
    for (int sl = gridDim.x + blockIdx.x; sl < B2_BN_PARTS; sl += gridDim.x)
      for (int o = threadIdx.x; o < 2 * C; o += kThreadsS) partials[(size_t)sl * 2 * C + o] = 0.f;
  (void)cv;
}

// ---------------------------------------------------------------- stats
__global__ void __launch_bounds__(kThreadsS, 3)
stats_stream_kernel(const bf16* __restrict__ y, long long total_elems, int C, float* __restrict__ partials,
                    int totals) {
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  Ring<1> ring;
  ring.src[0] = reinterpret_cast<const uint8_t*>(y);
  ring_setup(ring, smem, total_elems * 2);
  float* red = reinterpret_cast<float*>(smem + (size_t)kStagesS * kChunkBytes + 64);
  const int CV = C >> 3;                           // CV <= 256 and 256 % CV == 0
  const int cv = threadIdx.x % CV;
  const long long chunks = (total_elems * 2 + kChunkBytes - 1) / kChunkBytes;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (threadIdx.x == 0)
    for (int st = 0; st < kStagesS; ++st)
      if (blockIdx.x + (long long)st * gridDim.x < chunks) ring.issue(st, blockIdx.x + (long long)st * gridDim.x);
  int stage = 0;
  uint32_t phase = 0;
  for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
    s_mbar_wait(&ring.full[stage], phase);
    const long long base = c * (kChunkBytes / 2);
    const bf16* d = ring.data(stage, 0);
#pragma unroll
    for (int i = 0; i < kChunkVecs / kThreadsS; ++i) {
      const int v = threadIdx.x + i * kThreadsS;
      if (base + v * 8 < total_elems) {
        float f[8];
        load8(d + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
      }
    }
    __syncthreads();
    const long long nxt = c + (long long)kStagesS * gridDim.x;
    if (threadIdx.x == 0 && nxt < chunks) ring.issue(stage, nxt);
    if (++stage == kStagesS) { stage = 0; phase ^= 1; }
  }
  slot_reduce(s, q, partials, C, cv, CV, red, totals);
}

// ---------------------------------------------------------------- apply
// mode 0: mean / invstd given; 1: derive them from totals[2C] (block 0 also writes mean_out / invstd_out and
// updates the running statistics, BatchNorm2d training semantics); 2: frozen running statistics
struct ApplyFin {
  int mode;
  const float* totals;
  long long rows;
  float* running_mean;
  float* running_var;
  float momentum, eps;
  float* mean_out;
  float* invstd_out;
};

__global__ void __launch_bounds__(kThreadsS, 3)
apply_stream_kernel(const bf16* __restrict__ y, const bf16* __restrict__ residual, bf16* __restrict__ z,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ row_mask, int relu,
                    long long total_elems, int C, const ApplyFin fin, uint8_t* __restrict__ gate_out, int reverse) {
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  const int CV = C >> 3, cv = threadIdx.x % CV, logC = 31 - __clz(C);
  pdl_wait();            // the per-channel constants below come from the preceding finalize kernel
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    float mu, is;
    if (fin.mode == 0) {                       // finalized by a separate kernel
      mu = mean[c]; is = invstd[c];
    } else {
      if (fin.mode == 1) {                     // batch statistics from the producers' totals
        const double n = (double)fin.rows, m = (double)fin.totals[c] / n;
        double var = (double)fin.totals[C + c] / n - m * m;
        if (var < 0) var = 0;
        mu = (float)m;
        is = (float)(1.0 / sqrt(var + (double)fin.eps));
        if (blockIdx.x == 0 && threadIdx.x < CV && fin.running_mean) {
          const double unbiased = n > 1 ? var * n / (n - 1) : var;
          fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mu;
          fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unbiased;
        }
      } else {                                 // frozen statistics (eval)
        mu = fin.running_mean[c];
        is = 1.f / sqrtf(fin.running_var[c] + fin.eps);
      }
      if (blockIdx.x == 0 && threadIdx.x < CV) { fin.mean_out[c] = mu; fin.invstd_out[c] = is; }
    }
    sc[j] = is * gamma[c];
    sh[j] = beta[c] - mu * sc[j];
  }
  const long long chunks = (total_elems * 2 + kChunkBytes - 1) / kChunkBytes;
  auto body = [&](auto& ring, bool has_res) {
    if (reverse) ring.last_chunk = chunks - 1;
    if (threadIdx.x == 0)
      for (int st = 0; st < kStagesS; ++st)
        if (blockIdx.x + (long long)st * gridDim.x < chunks) ring.issue(st, blockIdx.x + (long long)st * gridDim.x);
    int stage = 0;
    uint32_t phase = 0;
    for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
      s_mbar_wait(&ring.full[stage], phase);
      const long long base = ring.phys(c) * (kChunkBytes / 2);
#pragma unroll
      for (int i = 0; i < kChunkVecs / kThreadsS; ++i) {
        const int v = threadIdx.x + i * kThreadsS;
        const long long e = base + v * 8;
        if (e < total_elems) {
          float f[8], g[8];
          load8(ring.data(stage, 0) + v * 8, f);
          if (has_res) load8(ring.data(stage, 1) + v * 8, g);
          const float mk = row_mask ? row_mask[e >> logC] : 1.f;
          uint32_t bits = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = fmaf(f[j], sc[j], sh[j]);
            if (has_res) a += g[j];
            bits |= (a > 0.f ? 1u : 0u) << j;
            if (relu) a = fmaxf(a, 0.f);
            f[j] = a * mk;
          }
          store8(z + e, f);
          if (gate_out) gate_out[e >> 3] = (uint8_t)bits;      // ReLU gate of residual layers: 1 bit per element
        }
      }
      __syncthreads();
      const long long nxt = c + (long long)kStagesS * gridDim.x;
      if (threadIdx.x == 0 && nxt < chunks) ring.issue(stage, nxt);
      if (++stage == kStagesS) { stage = 0; phase ^= 1; }
    }
  };
  if (residual) {
    Ring<2> ring;
    ring.src[0] = reinterpret_cast<const uint8_t*>(y);
    ring.src[1] = reinterpret_cast<const uint8_t*>(residual);
    ring_setup(ring, smem, total_elems * 2);
    body(ring, true);
  } else {
    Ring<1> ring;
    ring.src[0] = reinterpret_cast<const uint8_t*>(y);
    ring_setup(ring, smem, total_elems * 2);
    body(ring, false);
  }
}

// ---------------------------------------------------------------- backward: shared pieces
struct BwdArgs {
  const bf16 *dz, *z, *y;
  const uint8_t* gate;         // GSRC 2: ReLU gate bitmask written by the forward apply (1 bit per element)
  const float *mean, *invstd, *gamma, *beta, *row_mask, *row_scale, *gsum;
  int relu, training;
  bf16 *dy, *d_residual;
  float* partials;
  int totals;                  // reduce: partials is a pre-zeroed float[2C] (atomic adds)
  int l2_hint;                 // 0 none, 1 keep the inputs in L2 (reduce pass), 2 inputs are dead after this pass
  int reverse;                 // stream the chunks from the end of the tensors to their start (bn_order_mode)
  float *dgamma, *dbeta;       // apply: block 0 accumulates the affine gradients from gsum
  long long total_elems, rows;
  int C;
};

// MODE 0: reduce (partials of g and g*xhat);  MODE 1: apply (dy, d_residual)
// GSRC: where the ReLU gate comes from -- 0: recomputed from y (or no ReLU), 1: the saved output z (third
// input stream), 2: the forward's gate bitmask (one byte per 16-byte vector, read straight from global memory)
template <int MODE, int GSRC>
__global__ void __launch_bounds__(kThreadsS, 3) bwd_stream_kernel(const BwdArgs a) {
  constexpr bool HASZ = GSRC == 1;
  pdl_trigger();
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NIN = HASZ ? 3 : 2;
  Ring<NIN> ring;
  ring.src[0] = reinterpret_cast<const uint8_t*>(a.dz);
  ring.src[1] = reinterpret_cast<const uint8_t*>(a.y);
  if (HASZ) ring.src[NIN - 1] = reinterpret_cast<const uint8_t*>(a.z);
  if (GSRC == 2) {       // the gate bytes ride the same ring (behind the barriers and the reduction scratch)
    ring.bits = a.gate;
    ring.bits_buf = smem + (size_t)kStagesS * NIN * kChunkBytes + 64 + (MODE == 0 ? 2 * kThreadsS * 8 * sizeof(float) : 0);
  }
  ring.policy = l2_policy(a.l2_hint);
  ring_setup(ring, smem, a.total_elems * 2);
  float* red = reinterpret_cast<float*>(smem + (size_t)kStagesS * NIN * kChunkBytes + 64);
  const int C = a.C, CV = C >> 3, cv = threadIdx.x % CV, logC = 31 - __clz(C);
  const bool regate = a.relu && GSRC == 0;
  // forward scale/shift (gate recompute), and either (mu) for the reduce or (A, B, D) for the apply
  float sc[8], sh[8], k0[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const float mu = a.mean[c], is = a.invstd[c], ga = a.gamma[c] * is;
    sc[j] = ga;
    sh[j] = regate ? a.beta[c] - mu * ga : 0.f;
    if (MODE == 0) {
      k0[j] = mu; k1[j] = is; k2[j] = 0.f;
    } else {
      const float mg = a.training ? a.gsum[c] / (float)a.rows : 0.f;
      const float mgx = a.training ? a.gsum[C + c] / (float)a.rows : 0.f;
      k0[j] = ga;                       // A
      k1[j] = -ga * is * mgx;           // B
      k2[j] = -ga * mg - mu * k1[j];    // D
      if (blockIdx.x == 0 && threadIdx.x < CV) {
        if (a.dgamma) a.dgamma[c] += a.gsum[C + c];
        if (a.dbeta) a.dbeta[c] += a.gsum[c];
      }
    }
  }
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const long long chunks = (a.total_elems * 2 + kChunkBytes - 1) / kChunkBytes;
  if (a.reverse) ring.last_chunk = chunks - 1;
  if (threadIdx.x == 0)
    for (int st = 0; st < kStagesS; ++st)
      if (blockIdx.x + (long long)st * gridDim.x < chunks) ring.issue(st, blockIdx.x + (long long)st * gridDim.x);
  int stage = 0;
  uint32_t phase = 0;
  for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
    s_mbar_wait(&ring.full[stage], phase);
    const long long base = ring.phys(c) * (kChunkBytes / 2);
#pragma unroll
    for (int i = 0; i < kChunkVecs / kThreadsS; ++i) {
      const int v = threadIdx.x + i * kThreadsS;
      const long long e = base + v * 8;
      if (e < a.total_elems) {
        float g[8], f[8];
        load8(ring.data(stage, 0) + v * 8, g);
        load8(ring.data(stage, 1) + v * 8, f);
        if (HASZ) {
          float zz[8];
          load8(ring.data(stage, NIN - 1) + v * 8, zz);
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = zz[j] > 0.f ? g[j] : 0.f;
          }
        } else if (GSRC == 2) {
          const uint32_t bits = ring.bits_buf[(size_t)stage * (kChunkBytes / 16) + v];
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = ((bits >> j) & 1u) ? g[j] : 0.f;
        } else if (regate) {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = fmaf(f[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
        }
        const long long r = e >> logC;
        if (a.row_mask) {
          const float mk = a.row_mask[r];
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] *= mk;
        }
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s[j] += g[j]; q[j] = fmaf(g[j], f[j] - k0[j], q[j]); }
        } else {
          if (a.d_residual) store8(a.d_residual + e, g);
          const float rs = a.row_scale ? a.row_scale[r] : 1.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaf(g[j], k0[j], fmaf(f[j], k1[j], k2[j])) * rs;
          store8(a.dy + e, f);
        }
      }
    }
    __syncthreads();
    const long long nxt = c + (long long)kStagesS * gridDim.x;
    if (threadIdx.x == 0 && nxt < chunks) ring.issue(stage, nxt);
    if (++stage == kStagesS) { stage = 0; phase ^= 1; }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] *= k1[j];
    slot_reduce(s, q, a.partials, C, cv, CV, red, a.totals);
  }
}

inline bool enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2POSE_BN_TMA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

inline size_t smem_bytes(int nin, bool reduce, bool bits = false) {
  return (size_t)kStagesS * nin * kChunkBytes + 64 + (reduce ? 2 * kThreadsS * 8 * sizeof(float) : 0) +
         (bits ? (size_t)kStagesS * (kChunkBytes / 16) : 0);
}

inline int stream_grid(long long total_elems, int blocks_per_sm, bool slots) {
  static const int env_bps = getenv("B2POSE_BNS_BPS") ? atoi(getenv("B2POSE_BNS_BPS")) : 0;      // tuning override
  if (env_bps > 0) blocks_per_sm = env_bps;
  long long chunks = (total_elems * 2 + kChunkBytes - 1) / kChunkBytes;
  long long g = (long long)b2_num_sms() * blocks_per_sm;
  if (slots && g > B2_BN_PARTS) g = B2_BN_PARTS;
  if (g > chunks) g = chunks;
  return (int)(g < 1 ? 1 : g);
}

template <typename K>
int opt_in(K kernel, size_t sh) {
  if (sh > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    B2_REQUIRE(e == cudaSuccess, B2_E_LAUNCH, "bn_stream: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  return B2_OK;
}

}  // namespace

bool bn_stream_eligible(int C, int dtype) {
  return enabled() && dtype == B2_BF16 && C >= 64 && C <= 2048 && (C & (C - 1)) == 0;
}

int bn_stream_stats(const void* y, int64_t rows, int C, float* partials, int totals, cudaStream_t st) {
  const size_t sh = smem_bytes(1, true);
  int rc = opt_in(stats_stream_kernel, sh);
  if (rc) return rc;
  launch_pdl(stats_stream_kernel, dim3(stream_grid(rows * C, 2, true)), dim3(kThreadsS), sh, st, (const bf16*)y,
             (long long)(rows * C), C, partials, totals);
  B2_LAUNCH_CHECK("bn_stats(stream)");
  return B2_OK;
}

int bn_stream_apply(const void* y, const void* residual, void* z, const float* mean, const float* invstd,
                    const float* gamma, const float* beta, const float* row_mask, int relu, int64_t rows, int C,
                    cudaStream_t st) {
  return bn_stream_apply_fin(y, residual, z, mean, invstd, gamma, beta, row_mask, relu, rows, C, 0, nullptr, nullptr,
                             nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, st);
}

int bn_stream_apply_fin(const void* y, const void* residual, void* z, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, const float* row_mask, int relu, int64_t rows, int C,
                        int mode, const float* totals, float* running_mean, float* running_var, float momentum,
                        float eps, float* mean_out, float* invstd_out, uint8_t* gate_out, cudaStream_t st) {
  ApplyFin fin{};
  fin.mode = mode; fin.totals = totals; fin.rows = rows; fin.running_mean = running_mean;
  fin.running_var = running_var; fin.momentum = momentum; fin.eps = eps; fin.mean_out = mean_out;
  fin.invstd_out = invstd_out;
  const size_t sh = smem_bytes(residual ? 2 : 1, false);
  int rc = opt_in(apply_stream_kernel, sh);
  if (rc) return rc;
  launch_pdl(apply_stream_kernel, dim3(stream_grid(rows * C, 3, false)), dim3(kThreadsS), sh, st, (const bf16*)y,
             (const bf16*)residual, (bf16*)z, mean, invstd, gamma, beta, row_mask, relu, (long long)(rows * C), C, fin,
             gate_out, (bn_order_mode() & 1) ? 1 : 0);
  B2_LAUNCH_CHECK("bn_apply(stream)");
  return B2_OK;
}

static BwdArgs make_bwd(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, const float* row_mask, int relu, int64_t rows, int C) {
  BwdArgs a{};
  a.dz = (const bf16*)dz; a.z = (const bf16*)z; a.y = (const bf16*)y;
  a.mean = mean; a.invstd = invstd; a.gamma = gamma; a.beta = beta; a.row_mask = row_mask; a.relu = relu;
  a.total_elems = rows * C; a.rows = rows; a.C = C;
  return a;
}

int bn_stream_bwd_reduce(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const float* row_mask, int relu, float* partials,
                         int totals, const uint8_t* gate, int64_t rows, int C, cudaStream_t st) {
  BwdArgs a = make_bwd(dz, z, y, mean, invstd, gamma, beta, row_mask, relu, rows, C);
  a.partials = partials; a.totals = totals; a.gate = gate;
  // keep dz / y in L2 for the apply pass when both fit beside the rest of the working set
  a.l2_hint = (l2_hint_mode() && (long long)rows * C * 4 <= 96LL << 20) ? 1 : 0;
  a.reverse = (bn_order_mode() & 2) ? 1 : 0;
  const int gsrc = !relu ? 0 : (gate ? 2 : (z ? 1 : 0));
  const size_t sh = smem_bytes(gsrc == 1 ? 3 : 2, true, gsrc == 2);
  const int grid = stream_grid(rows * C, 2, true);
  int rc;
  if (gsrc == 1) {
    if ((rc = opt_in(bwd_stream_kernel<0, 1>, sh))) return rc;
    launch_pdl(bwd_stream_kernel<0, 1>, dim3(grid), dim3(kThreadsS), sh, st, a);
  } else if (gsrc == 2) {
    if ((rc = opt_in(bwd_stream_kernel<0, 2>, sh))) return rc;
    launch_pdl(bwd_stream_kernel<0, 2>, dim3(grid), dim3(kThreadsS), sh, st, a);
  } else {
    if ((rc = opt_in(bwd_stream_kernel<0, 0>, sh))) return rc;
    launch_pdl(bwd_stream_kernel<0, 0>, dim3(grid), dim3(kThreadsS), sh, st, a);
  }
  B2_LAUNCH_CHECK("bn_bwd_reduce(stream)");
  return B2_OK;
}

int bn_stream_bwd_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, const float* gsum, const float* row_mask,
                        const float* row_scale, int relu, int training, void* dy, void* d_residual, float* dgamma,
                        float* dbeta, const uint8_t* gate, int64_t rows, int C, cudaStream_t st) {
  BwdArgs a = make_bwd(dz, z, y, mean, invstd, gamma, beta, row_mask, relu, rows, C);
  a.dgamma = dgamma; a.dbeta = dbeta; a.gate = gate;
  a.gsum = gsum; a.row_scale = row_scale; a.training = training; a.dy = (bf16*)dy; a.d_residual = (bf16*)d_residual;
  a.l2_hint = l2_hint_mode() ? 2 : 0;            // dz and y are dead after this pass
  a.reverse = (bn_order_mode() & 4) ? 1 : 0;
  const int gsrc = !relu ? 0 : (gate ? 2 : (z ? 1 : 0));
  const size_t sh = smem_bytes(gsrc == 1 ? 3 : 2, false, gsrc == 2);
  const int grid = stream_grid(rows * C, 3, false);
  int rc;
  if (gsrc == 1) {
    if ((rc = opt_in(bwd_stream_kernel<1, 1>, sh))) return rc;
    launch_pdl(bwd_stream_kernel<1, 1>, dim3(grid), dim3(kThreadsS), sh, st, a);
  } else if (gsrc == 2) {
    if ((rc = opt_in(bwd_stream_kernel<1, 2>, sh))) return rc;
    launch_pdl(bwd_stream_kernel<1, 2>, dim3(grid), dim3(kThreadsS), sh, st, a);
  } else {
    if ((rc = opt_in(bwd_stream_kernel<1, 0>, sh))) return rc;
    launch_pdl(bwd_stream_kernel<1, 0>, dim3(grid), dim3(kThreadsS), sh, st, a);
  }
  B2_LAUNCH_CHECK("bn_bwd_apply(stream)");
  return B2_OK;
}
