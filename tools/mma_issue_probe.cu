// Micro-benchmark: how fast can ONE warp issue tcgen05.mma (M128 x N x K16, bf16, SS operands) on sm_100a?
// Round-1's study (profiles/r01_conv_issue_study.md) saw one MMA leave every ~175-190 cycles whatever N.
// This probe separates the candidate causes: the issue-loop style, the accumulator dependency, commits,
// mbarrier waits in the loop.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_issue_probe tools/mma_issue_probe.cu
//   tools/bin/mma_issue_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (it > 40000000u) __trap();
  }
}
__device__ __forceinline__ uint32_t mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// One asm block per ring stage: poll the NEXT stage's barrier (result consumed at the end, so that its latency overlaps
// the issue), then NK MMAs and the commit under an elect.sync predicate.
template <int NK>
__device__ __forceinline__ uint32_t stage_issue(uint32_t next_bar, uint32_t next_parity, uint32_t tmem_d, uint64_t adesc,
                                                uint64_t bdesc, uint32_t idesc, uint32_t acc0, uint32_t empty_bar) {
  uint32_t ready;
  if (NK == 4) {
    asm volatile(
        "{\n\t.reg .pred pw, pe, pa;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 pw, [%1], %2;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pa, %7, 0;\n\t"
        "add.s64 a1, %4, 2;\n\tadd.s64 a2, %4, 4;\n\tadd.s64 a3, %4, 6;\n\t"
        "add.s64 b1, %5, 2;\n\tadd.s64 b2, %5, 4;\n\tadd.s64 b3, %5, 6;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], %4, %5, %6, pa;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a1, b1, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a2, b2, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a3, b3, %6, 1;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, pw;\n\t}"
        : "=r"(ready)
        : "r"(next_bar), "r"(next_parity), "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc0), "r"(empty_bar)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pw, pe, pa;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 pw, [%1], %2;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pa, %7, 0;\n\t"
        "add.s64 a1, %4, 2;\n\tadd.s64 a2, %4, 4;\n\tadd.s64 a3, %4, 6;\n\t"
        "add.s64 b1, %5, 2;\n\tadd.s64 b2, %5, 4;\n\tadd.s64 b3, %5, 6;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], %4, %5, %6, pa;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a1, b1, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a2, b2, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a3, b3, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], %4, %5, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a1, b1, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a2, b2, %6, 1;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%3], a3, b3, %6, 1;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, pw;\n\t}"
        : "=r"(ready)
        : "r"(next_bar), "r"(next_parity), "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc0), "r"(empty_bar)
        : "memory");
  }
  return ready;
}

struct Bars {
  uint64_t done, full[8], empty[8];
  uint32_t tmem_base;
  volatile int scout;
};

// variant:
//  0  lane 0 inside `if (lane == 0)`: 4 MMAs (one 64-wide k-block, descriptors advance by 32 B) back to back, loop
//  1  same + tcgen05.commit to a barrier nobody waits on after every 4 MMAs (the production loop's `empty` commit)
//  2  same as 1 + an mbarrier wait (already complete: a helper warp keeps `full` ahead) before every 4 MMAs, whole
//     warp waits, __syncwarp afterwards: the production loop
//  3  like 0 but the 4 MMAs rotate over 2 accumulators (independent chains)
//  4  like 0 but issued through elect.sync (warp-uniform control flow)
//  5  like 2 but only lane 0 polls the barrier (the other lanes park at __syncwarp)
//  6  like 0 with the A/B tiles of 4 different ring stages (descriptor rebuilt per k-block from a stage index)
template <int V>
__global__ void __launch_bounds__(320, 1) probe(int N, int groups, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // 4 stages x (A 16 KB + B 32 KB)
  const uint32_t stage_bytes = 16384 + 32768;
  Bars* bars = reinterpret_cast<Bars*>(smem + 4 * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < 4 * stage_bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + ((i * 2654435761u) >> 28) * 0x00010001u;   // small bf16 values
  if (threadIdx.x == 0) {
    bars->scout = 0;
    mbar_init(&bars->done, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t idesc = instr_desc(128, N);

  if (warp >= 2 && (V == 2 || V == 5 || V == 8 || V == 9 || V == 10 || V == 11 || V == 12 || V == 14 || V == 15 || V == 16)) {
    // helpers: warp 2+s keeps the `full` barrier of ring stage s complete ahead of the consumer (one warp per stage so
    // that the helpers are never the bottleneck)
    if (lane == 0) {
      const int stage = warp - 2;
      uint32_t phase = 0;
      for (int g = stage; g < groups; g += 8) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        mbar_arrive(&bars->full[stage]);
        phase ^= 1;
      }
    }
  } else if (warp == 0 && V == 15) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int g = 0; g < groups; ++g) {
        mbar_wait(&bars->full[stage], phase);
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        bars->scout = g + 1;
        if (++stage == 8) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    __syncwarp();
    long long t0 = clock64(), t1 = 0;
    int stage = 0;
    uint32_t phase = 0;
    if (V == 7 || V == 8 || V == 9 || V == 10 || V == 12 || V == 13) {
      // elect.sync issue + commit (+ whole-warp wait): the CUTLASS pattern.  V9 polls the NEXT stage's barrier before
      // issuing this stage's MMAs so that the try_wait latency overlaps the issue; V10 has 8 MMAs per stage.
      uint32_t ready = 0;
      if (V == 9) { mbar_wait(&bars->full[0], 0); ready = 1; }
      for (int g = 0; g < groups; ++g) {
        if (V == 8 || V == 10 || V == 12) {
          mbar_wait(&bars->full[stage], phase);
        } else if (V == 9) {
          if (!ready) mbar_wait(&bars->full[stage], phase);
        }
        if (V != 7 && V != 12) tc_fence_after();
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == 8) { nstage = 0; nphase ^= 1; }
        if (V == 9) ready = (g + 1 < groups) ? mbar_try(&bars->full[nstage], nphase) : 0;
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)(g & 3) * stage_bytes);
          const uint64_t adesc = smem_desc(sa, 0, 1024), bdesc = smem_desc(sa + 16384, 0, 1024);
#pragma unroll
          for (int k = 0; k < (V == 10 ? 8 : 4); ++k)
            umma_bf16(tmem_base, adesc + (uint64_t)((k & 3) * 2), bdesc + (uint64_t)((k & 3) * 2), idesc, (uint32_t)(g | k));
          umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        stage = nstage; phase = nphase;
      }
    } else if (V == 14 || V == 16) {
      mbar_wait(&bars->full[0], 0);
      uint32_t ready = 1;
      for (int g = 0; g < groups; ++g) {
        if (!ready) mbar_wait(&bars->full[stage], phase);
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == 8) { nstage = 0; nphase ^= 1; }
        const uint32_t sa = smem_u32(smem + (size_t)(g & 3) * stage_bytes);
        const uint64_t adesc = smem_desc(sa, 0, 1024), bdesc = smem_desc(sa + 16384, 0, 1024);
        // (the last iteration polls a barrier that never completes: harmless, result unused)
        ready = stage_issue<(V == 16 ? 8 : 4)>(smem_u32(&bars->full[nstage]), nphase, tmem_base, adesc, bdesc, idesc,
                                                (uint32_t)g, smem_u32(&bars->empty[stage]));
        if (g + 1 == groups) ready = 1;
        stage = nstage; phase = nphase;
      }
    } else if (V == 15) {
      // a scout (warp 0) waits on the `full` barriers and publishes the number of ready stages in shared memory;
      // the issuing warp polls that counter with plain shared loads
      for (int g = 0; g < groups; ++g) {
        while (bars->scout <= g) { }
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)(g & 3) * stage_bytes);
          const uint64_t adesc = smem_desc(sa, 0, 1024), bdesc = smem_desc(sa + 16384, 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(g | k));
          umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        if (++stage == 8) { stage = 0; phase ^= 1; }
      }
    } else if (V == 11) {
      // ONE elected thread runs the whole loop (wait, MMAs, commit)
      if (elect_one()) {
        for (int g = 0; g < groups; ++g) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)(g & 3) * stage_bytes);
          const uint64_t adesc = smem_desc(sa, 0, 1024), bdesc = smem_desc(sa + 16384, 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(g | k));
          umma_commit(&bars->empty[stage]);
          if (++stage == 8) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    } else if (V == 4) {
      for (int g = 0; g < groups; ++g) {
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)(g & 3) * stage_bytes);
          const uint64_t adesc = smem_desc(sa, 0, 1024), bdesc = smem_desc(sa + 16384, 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(g | k));
        }
        __syncwarp();
      }
    } else {
      for (int g = 0; g < groups; ++g) {
        if (V == 2) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
        }
        if (lane == 0) {
          if (V == 5) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
          }
          const uint32_t sa = smem_u32(smem + (size_t)((V == 0 || V == 3) ? 0 : (g & 3)) * stage_bytes);
          const uint64_t adesc = smem_desc(sa, 0, 1024), bdesc = smem_desc(sa + 16384, 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t d = (V == 3) ? tmem_base + (uint32_t)((k & 1) * 256) : tmem_base;
            umma_bf16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(g | (V == 3 ? (k >> 1) : k)));
          }
          if (V == 1 || V == 2 || V == 5) umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        if (++stage == 8) { stage = 0; phase ^= 1; }
      }
    }
    t1 = clock64();
    if (lane == 0) umma_commit(&bars->done);
    __syncwarp();
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    long long t2 = clock64();
    if (lane == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int V>
void run(const char* what, int grid) {
  const int groups = 2048;
  const size_t smem = 4 * (16384 + 32768) + 1024 + 256;
  cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* out;
  cudaMalloc(&out, 2 * 148 * sizeof(long long));
  for (int N : {64, 128, 256}) {
    if (V == 3 && N > 256) continue;
    float best_ms = 1e9f;
    long long h[2 * 148];
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      probe<V><<<grid, 320, smem>>>(N, groups, out);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s N=%d: CUDA error %s\n", what, N, cudaGetErrorString(e)); exit(1); }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best_ms) best_ms = ms;
      cudaMemcpy(h, out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    }
    double issue = 0, total = 0;
    for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
    issue /= grid; total /= grid;
    const double mmas = groups * ((V == 10 || V == 16) ? 8.0 : 4.0);
    printf("V%d %-58s grid %3d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d), kernel %.3f ms\n", V, what,
           grid, N, issue / mmas, total / mmas, N / 2, best_ms);
  }
  cudaFree(out);
}

int main() {
  for (int grid : {148}) {
    run<0>("lane0, 4 MMAs back to back", grid);
    run<1>("+ commit after every 4", grid);
    run<2>("+ whole-warp mbarrier wait before every 4 (production)", grid);
    run<3>("2 accumulators alternating", grid);
    run<4>("elect.sync issue", grid);
    run<5>("lane-0-only mbarrier wait + commit", grid);
    run<6>("4 ring stages, descriptors rebuilt", grid);
    run<7>("elect + commit after every 4", grid);
    run<8>("elect + commit + whole-warp wait (CUTLASS pattern)", grid);
    run<9>("elect + commit + wait polled one stage ahead", grid);
    run<10>("elect + commit + wait, 8 MMAs per stage", grid);
    run<11>("one elected thread runs the whole loop", grid);
    run<14>("fused asm: poll next + 4 MMAs + commit, elect", grid);
    run<16>("fused asm: poll next + 8 MMAs + commit, elect", grid);
    run<15>("scout warp + shared-memory counter, elect", grid);
    run<12>("elect + commit + whole-warp wait, NO tcgen05.fence", grid);
    run<13>("elect + commit + tcgen05.fence, NO wait", grid);
  }
  return 0;
}
