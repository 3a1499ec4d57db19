// Micro-benchmark: TMEM -> register throughput of tcgen05.ld (32x32b) with 4 / 8 / 16 warps per SM, x16 and x32 shapes.
// Decides whether the convolution epilogue (128 x N fp32 accumulators per tile) can be sped up with more warps or is at
// the TMEM read rate.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_ld_probe tools/tmem_ld_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

template <int X>
__global__ void probe(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[X];
#pragma unroll
    for (int c = 0; c < 512; c += 4 * X) {          // 4 loads in flight, then one wait
      uint32_t w0[X], w1[X], w2[X], w3[X];
      tmem_ld<X>(base + ((c + (warp >> 2) * X) & 511), w0);
      tmem_ld<X>(base + ((c + X + (warp >> 2) * X) & 511), w1);
      tmem_ld<X>(base + ((c + 2 * X + (warp >> 2) * X) & 511), w2);
      tmem_ld<X>(base + ((c + 3 * X + (warp >> 2) * X) & 511), w3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < X; ++j) acc ^= w0[j] ^ w1[j] ^ w2[j] ^ w3[j];
    }
    (void)v;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}

template <int X>
void run(int warps) {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int iters = 200;
  probe<X><<<148, warps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("x%d warps %d: %s\n", X, warps, cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (int i = 0; i < 148; ++i) cyc += h[i];
  cyc /= 148;
  // every warp reads 32 lanes x 512 columns x 4 B per iteration
  const double bytes = (double)warps * 32 * 512 * 4 * iters;
  printf("tcgen05.ld 32x32b.x%d, %2d warps/SM: %.1f B/clk/SM  (a 128 x 256 fp32 tile = %.0f cycles)\n", X, warps, bytes / cyc,
         128.0 * 256 * 4 / (bytes / cyc));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) { run<16>(w); run<32>(w); }
  return 0;
}
