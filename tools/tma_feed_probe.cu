// Micro-benchmark: how fast can ONE SM pull im2col-style activation tiles out of the L2 with TMA?  The 3x3 layers of the
// shallow stages run at ~22 B/clk/SM of operand feed (9 shifted 16 KB boxes per 128-pixel tile); this probe replays that
// access pattern with no MMA and no epilogue -- a producer thread issuing 4-D tiled boxes into a ring, a consumer thread
// releasing the stages -- and reports bytes per clock and SM for several ring depths, box shapes and channel counts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/tma_feed_probe tools/tma_feed_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct Params {
  int N, H, W, C;         // NHWC bf16 tensor
  int BW, BH;             // pixel brick of a tile (BW * BH = 128)
  int cblocks;            // C / 64
  int stages;
  int taps;               // 9 (3x3, shifted boxes) or 1
  int boxes_per_stage;    // 1: one {64, BW, BH, 1} box per stage; 2: two {64, BW, BH/2, 1} boxes
  long long* cycles;
};

constexpr int kStageBytes = 16384;

__global__ void __launch_bounds__(64, 1) feed_kernel(const __grid_constant__ CUtensorMap map, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * kStageBytes);
  uint64_t* empty = full + p.stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tiles_w = p.W / p.BW, tiles_h = p.H / p.BH;
  const int total = p.N * tiles_h * tiles_w;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int wi = tile % tiles_w, hi = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
      for (int t = 0; t < p.taps; ++t) {
        const int r = p.taps == 9 ? t / 3 - 1 : 0, s = p.taps == 9 ? t % 3 - 1 : 0;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect(&full[stage], kStageBytes);
          const uint32_t dst = smem_u32(smem + (size_t)stage * kStageBytes);
          if (p.boxes_per_stage == 1) {
            tma_load_4d(dst, &map, &full[stage], cb * 64, wi * p.BW + s, hi * p.BH + r, n);
          } else {
            tma_load_4d(dst, &map, &full[stage], cb * 64, wi * p.BW + s, hi * p.BH + r, n);
            tma_load_4d(dst + kStageBytes / 2, &map, &full[stage], cb * 64, wi * p.BW + s, hi * p.BH + r + p.BH / 2, n);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x)
      for (int t = 0; t < p.taps * p.cblocks; ++t) {
        mbar_wait(&full[stage], phase);
        mbar_arrive(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fnp;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* cycles;
  CK(cudaMalloc(&cycles, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  struct Case { const char* name; int N, H, W, C, BW, BH, taps, boxes; CUtensorMapL2promotion promo; };
  const Case cases[] = {
      {"3x3 C=64  64x64 box 64x2       ", 64, 64, 64, 64, 64, 2, 9, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"3x3 C=64  64x64 two boxes 64x1 ", 64, 64, 64, 64, 64, 2, 9, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"3x3 C=64  64x64 box 64x2 promo128", 64, 64, 64, 64, 64, 2, 9, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"3x3 C=64  64x64 box 64x2 no promo", 64, 64, 64, 64, 64, 2, 9, 1, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"1x1 C=64  64x64 box 64x2       ", 64, 64, 64, 64, 64, 2, 1, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"3x3 C=128 32x32 box 32x4       ", 64, 32, 32, 128, 32, 4, 9, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"3x3 C=256 16x16 box 16x8       ", 64, 16, 16, 256, 16, 8, 9, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"1x1 C=256 64x64 box 64x2       ", 64, 64, 64, 256, 64, 2, 1, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
  };
  for (const Case& c : cases) {
    const size_t bytes = (size_t)c.N * c.H * c.W * c.C * 2;
    void* x;
    CK(cudaMalloc(&x, bytes));
    CK(cudaMemset(x, 0, bytes));
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)c.BW, (cuuint32_t)(c.boxes == 2 ? c.BH / 2 : c.BH), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int stages : {3, 7, 12}) {
      Params p{c.N, c.H, c.W, c.C, c.BW, c.BH, c.C / 64, stages, c.taps, c.boxes, cycles};
      const size_t smem = (size_t)stages * kStageBytes + 2 * stages * 8 + 1024 + 64;
      for (int rep = 0; rep < 3; ++rep) {      // rep 0 warms the L2
        feed_kernel<<<sms, 64, smem>>>(map, p);
        CK(cudaDeviceSynchronize());
      }
      long long h[256];
      CK(cudaMemcpy(h, cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const int tiles = c.N * (c.H / c.BH) * (c.W / c.BW);
      const double per_sm_bytes = (double)((tiles + sms - 1) / sms) * c.taps * (c.C / 64) * kStageBytes;
      printf("%s stages %2d: %9lld cycles  %6.1f B/clk/SM  (%5.1f cycles per 128-byte row; tensor %zu MB)\n", c.name, stages, mx,
             per_sm_bytes / (double)mx, (double)mx / (per_sm_bytes / 128.0), bytes >> 20);
    }
    CK(cudaFree(x));
  }
  return 0;
}
