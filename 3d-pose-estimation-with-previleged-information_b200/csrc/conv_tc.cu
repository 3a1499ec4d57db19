// tcgen05 implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
//   fprop : D[pixels, K] = sum over (tap, c) of X[pixel shifted by tap, c] * W[k, tap, c]
//   dgrad : the same kernel run on dy with the filter flipped/transposed (stride-1 layers)
//   wgrad : D[k, (tap, c)] = sum over pixels of dY[pixel, k] * X[pixel shifted by tap, c]  (MN-major operands)
//
// One persistent CTA per SM (320 threads; 192 in the wgrad kernel): warp 0 is the TMA producer, warp 1 issues tcgen05.mma
// (one elected lane), warps 2-9 are the epilogue (two warps share the 32 TMEM lanes of a warp-id
// quarter).  Activation tiles are fetched with 4-D *tiled* TMA boxes (C, W, H, N): a tile of
// output pixels is a BNIxBHxBW brick, and for filter tap (r, s) the A operand is that brick
// shifted by (r*dil - pad, s*dil - pad); out-of-bounds rows/columns (padding, ragged edges) are
// zero-filled by the TMA unit, strided layers use the tensor map's element strides.  Tiles land
// in shared memory in the 128-byte-swizzled K-major layout the UMMA descriptors expect, so no
// thread ever touches the operands.  Accumulators are double buffered in TMEM so the epilogue of
// tile i (ratio / bias / bf16 pack / store) overlaps the MMAs of tile i+1.
//
// PartialConv semantics (partial_conv.py:32-58): 1x1 layers need no input masking at all
// (conv(x*m) == conv(x)*m row-wise, and m is folded into the ratio); 3x3 layers read input that
// the producing BN epilogue already multiplied by the veil (B2_CONV_X_PREMASKED); the epilogue
// computes the mask-window count for its pixel, emits mask_out / ratio and scales the row.
#include <cuda.h>

#include "b2_common.cuh"

namespace {

constexpr int kThreads = 192;          // wgrad kernel: producer, MMA, 4 epilogue warps
constexpr int kConvThreads = 320;      // fprop/dgrad kernel: producer, MMA, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kMaxStages = 8;
constexpr int kTileM = 128;          // UMMA M (output pixels per tile / TMEM lanes)
constexpr int kBlockK = 64;          // bf16 elements per 128-byte swizzle row
constexpr uint32_t kABytes = kTileM * kBlockK * 2;
constexpr int kStaging = 4;          // 16 KB output staging buffers of the TMA-store epilogue

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier_group(int grp) {          // four-warp epilogue groups: barriers 2 and 3
  asm volatile("bar.sync %0, 128;" ::"r"(grp + 2) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
// ---- MMA issue.  Measured with tools/mma_issue_probe.cu (profiles/r02_mma_issue_probe.md): one `tcgen05.mma` issued by
// `if (lane == 0)` code costs ~73 cycles (ptxas wraps every instruction in an ELECT / BRA.U.ANY loop), and the issue
// queue is shallow, so every mbarrier wait between two groups of MMAs (~170 cycles even when the barrier is already
// complete) is a bubble in the tensor pipe: round 1's loop ran at 131-190 cycles per MMA whatever N.  One asm block
// that (1) polls the barriers the NEXT group will need with the non-blocking `test_wait` (a `try_wait` on a phase that
// is not complete yet suspends the warp in front of its own MMAs) -- the predicates are consumed only at the end of
// the block, so the poll latency overlaps the issue -- and (2) issues four MMAs and the commits under ONE elect.sync predicate
// runs at the floor: 55 / 64 / 128 cycles per MMA for N = 64 / 128 / 256.
constexpr uint32_t kPoll1 = 1, kPoll2 = 2, kCommit1 = 4, kCommit2 = 8, kTwoMma = 16;    // kTwoMma: only the first two MMAs
// Four MMAs D[tmem_d] (+)= A_k * B_k, k = 0..3, descriptors advancing by `dstep` (encoded >> 4 units) per k-step.
// acc0: accumulate flag of the first MMA (the others always accumulate).  Must be executed by a converged warp.
__device__ __forceinline__ void mma4_fused(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint64_t dstep, uint32_t idesc,
                                           uint32_t acc0, uint32_t flags, uint32_t poll1_bar, uint32_t poll1_parity,
                                           uint32_t poll2_bar, uint32_t poll2_parity, uint32_t commit1_bar,
                                           uint32_t commit2_bar, uint32_t& ready1, uint32_t& ready2, uint64_t bstep) {
  asm volatile(
      "{\n\t"
      ".reg .pred pw1, pw2, pe, pa, q1, q2, c1, c2, p34;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      ".reg .b32 t;\n\t"
      "and.b32 t, %8, 1;\n\tsetp.ne.b32 q1, t, 0;\n\t"
      "and.b32 t, %8, 2;\n\tsetp.ne.b32 q2, t, 0;\n\t"
      "setp.ne.b32 pw1, 0, 0;\n\tsetp.ne.b32 pw2, 0, 0;\n\t"
      "@q1 mbarrier.test_wait.parity.shared::cta.b64 pw1, [%9], %10;\n\t"
      "@q2 mbarrier.test_wait.parity.shared::cta.b64 pw2, [%11], %12;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "and.b32 t, %8, 4;\n\tsetp.ne.b32 c1, t, 0;\n\tand.pred c1, c1, pe;\n\t"
      "and.b32 t, %8, 8;\n\tsetp.ne.b32 c2, t, 0;\n\tand.pred c2, c2, pe;\n\t"
      "and.b32 t, %8, 16;\n\tsetp.eq.b32 p34, t, 0;\n\tand.pred p34, p34, pe;\n\t"
      "setp.ne.b32 pa, %7, 0;\n\t"
      "add.s64 a1, %3, %5;\n\tadd.s64 a2, a1, %5;\n\tadd.s64 a3, a2, %5;\n\t"
      "add.s64 b1, %4, %15;\n\tadd.s64 b2, b1, %15;\n\tadd.s64 b3, b2, %15;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], %3, %4, %6, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1, %6, 1;\n\t"
      "@p34 tcgen05.mma.cta_group::1.kind::f16 [%2], a2, b2, %6, 1;\n\t"
      "@p34 tcgen05.mma.cta_group::1.kind::f16 [%2], a3, b3, %6, 1;\n\t"
      "@c1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t"
      "@c2 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%14];\n\t"
      "selp.u32 %0, 1, 0, pw1;\n\tselp.u32 %1, 1, 0, pw2;\n\t"
      "}"
      : "=r"(ready1), "=r"(ready2)
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "l"(dstep), "r"(idesc), "r"(acc0), "r"(flags), "r"(poll1_bar),
        "r"(poll1_parity), "r"(poll2_bar), "r"(poll2_parity), "r"(commit1_bar), "r"(commit2_bar), "l"(bstep)
      : "memory");
}
// The common case of mma4_fused with every switch fixed at compile time -- four MMAs, poll the next stage's `full`
// barrier, commit to this stage's `empty` barrier: the flag decoding (ten uniform instructions) disappears from the issue
// loop of all but a tile's last k-block.
__device__ __forceinline__ void mma4_steady(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc0,
                                            uint32_t poll_bar, uint32_t poll_parity, uint32_t commit_bar, uint32_t& ready,
                                            uint64_t bstep) {
  asm volatile(
      "{\n\t"
      ".reg .pred pw, pe, pa;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 pw, [%6], %7;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %5, 0;\n\t"
      "add.s64 a1, %2, 2;\n\tadd.s64 a2, %2, 4;\n\tadd.s64 a3, %2, 6;\n\t"
      "add.s64 b1, %3, %9;\n\tadd.s64 b2, b1, %9;\n\tadd.s64 b3, b2, %9;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], a1, b1, %4, 1;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], a2, b2, %4, 1;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], a3, b3, %4, 1;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
      "selp.u32 %0, 1, 0, pw;\n\t"
      "}"
      : "=r"(ready)
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc0), "r"(poll_bar), "r"(poll_parity), "r"(commit_bar),
        "l"(bstep)
      : "memory");
}
// Row stems: the fourteen MMAs of one tile (seven tap rows x two K = 16 steps) in ONE block -- the cost of a block
// (elect, predicate set-up, the barrier polls) is paid per block, not per MMA.  Descriptors in 16-byte units: the A
// operand of tap row r starts `arow` units after row r - 1, the filter slice `brow` units; the second K step is 2 units
// (32 bytes) further in both.  Polls / commits as in mma4_fused (kPoll1 | kPoll2 in `flags`; both commits always).
__device__ __forceinline__ void mma14_rows(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint64_t arow, uint64_t brow,
                                           uint32_t idesc, uint32_t flags, uint32_t poll1_bar, uint32_t poll1_parity,
                                           uint32_t poll2_bar, uint32_t poll2_parity, uint32_t commit1_bar,
                                           uint32_t commit2_bar, uint32_t& ready1, uint32_t& ready2) {
#define B2_ROW_MMAS(ACC)                                                                   \
  "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], ad, bd, %7, " ACC ";\n\t"                  \
  "add.s64 ad2, ad, 2;\n\tadd.s64 bd2, bd, 2;\n\t"                                        \
  "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], ad2, bd2, %7, 1;\n\t"                      \
  "add.s64 ad, ad, %5;\n\tadd.s64 bd, bd, %6;\n\t"
  asm volatile(
      "{\n\t"
      ".reg .pred pw1, pw2, pe, pz, q1, q2;\n\t"
      ".reg .b64 ad, bd, ad2, bd2;\n\t"
      ".reg .b32 t;\n\t"
      "and.b32 t, %8, 1;\n\tsetp.ne.b32 q1, t, 0;\n\t"
      "and.b32 t, %8, 2;\n\tsetp.ne.b32 q2, t, 0;\n\t"
      "setp.ne.b32 pw1, 0, 0;\n\tsetp.ne.b32 pw2, 0, 0;\n\tsetp.ne.b32 pz, 0, 0;\n\t"
      "@q1 mbarrier.test_wait.parity.shared::cta.b64 pw1, [%9], %10;\n\t"
      "@q2 mbarrier.test_wait.parity.shared::cta.b64 pw2, [%11], %12;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "mov.b64 ad, %3;\n\tmov.b64 bd, %4;\n\t"
      B2_ROW_MMAS("pz") B2_ROW_MMAS("1") B2_ROW_MMAS("1") B2_ROW_MMAS("1") B2_ROW_MMAS("1") B2_ROW_MMAS("1") B2_ROW_MMAS("1")
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%14];\n\t"
      "selp.u32 %0, 1, 0, pw1;\n\tselp.u32 %1, 1, 0, pw2;\n\t"
      "}"
      : "=r"(ready1), "=r"(ready2)
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "l"(arow), "l"(brow), "r"(idesc), "r"(flags), "r"(poll1_bar),
        "r"(poll1_parity), "r"(poll2_bar), "r"(poll2_parity), "r"(commit1_bar), "r"(commit2_bar)
      : "memory");
#undef B2_ROW_MMAS
}
// Producer counterpart of mma4_fused: the single producing thread's serial latency per ring stage (a blocking wait on
// `empty`, two integer divisions, two TMA issues wrapped in ELECT loops: ~400-600 cycles) bounded the shallow layers
// once the MMA issue was fixed.  One asm block polls the NEXT stage's `empty` barrier (non-blocking, consumed at the
// end) and issues expect_tx + the activation box (+ the filter slice when kLoadB) under one elect.sync predicate.
constexpr uint32_t kLoadB = 2;
__device__ __forceinline__ uint32_t produce_fused(uint32_t flags, uint32_t poll_bar, uint32_t poll_parity, uint32_t full_bar,
                                                  uint32_t tx_bytes, uint32_t dst_a, const CUtensorMap* map_a, int a0, int a1,
                                                  int a2, int a3, uint32_t dst_b, const CUtensorMap* map_b, int b0, int b1) {
  uint32_t ready;
  asm volatile(
      "{\n\t"
      ".reg .pred pw, pe, q, lb;\n\t"
      ".reg .b32 t;\n\t"
      "and.b32 t, %1, 1;\n\tsetp.ne.b32 q, t, 0;\n\t"
      "setp.ne.b32 pw, 0, 0;\n\t"
      "@q mbarrier.test_wait.parity.shared::cta.b64 pw, [%2], %3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "and.b32 t, %1, 2;\n\tsetp.ne.b32 lb, t, 0;\n\tand.pred lb, lb, pe;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%4], %5;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%6], [%7, {%8, %9, %10, %11}], [%4];\n\t"
      "@lb cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%12], [%13, {%14, %15}], [%4];\n\t"
      "selp.u32 %0, 1, 0, pw;\n\t"
      "}"
      : "=r"(ready)
      : "r"(flags), "r"(poll_bar), "r"(poll_parity), "r"(full_bar), "r"(tx_bytes), "r"(dst_a), "l"(map_a), "r"(a0), "r"(a1),
        "r"(a2), "r"(a3), "r"(dst_b), "l"(map_b), "r"(b0), "r"(b1)
      : "memory");
  return ready;
}

// wgrad producer stage: poll + expect_tx + the two dY atoms (k0, k0 + 64; 8 KB apart) + the first X atom.
// flags bit 1 (value 2): the layer has at most 64 output channels -- the second atom is all zero, it is cleared once
// at kernel start instead of being zero-filled by TMA for every stage.
__device__ __forceinline__ uint32_t produce_wgrad_fused(uint32_t flags, uint32_t poll_bar, uint32_t poll_parity,
                                                        uint32_t full_bar, uint32_t tx_bytes, uint32_t dst_dy,
                                                        const CUtensorMap* map_dy, int k0, int ow0, int oh0, int n0,
                                                        uint32_t dst_x, const CUtensorMap* map_x, int xc, int xw, int xh) {
  uint32_t ready;
  asm volatile(
      "{\n\t"
      ".reg .pred pw, pe, q, p2;\n\t"
      ".reg .b32 t, d1, k1;\n\t"
      "and.b32 t, %1, 1;\n\tsetp.ne.b32 q, t, 0;\n\t"
      "setp.ne.b32 pw, 0, 0;\n\t"
      "@q mbarrier.test_wait.parity.shared::cta.b64 pw, [%2], %3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "add.u32 d1, %6, 8192;\n\tadd.s32 k1, %8, 64;\n\t"
      "and.b32 t, %1, 2;\n\tsetp.eq.b32 p2, t, 0;\n\tand.pred p2, p2, pe;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%4], %5;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%6], [%7, {%8, %9, %10, %11}], [%4];\n\t"
      "@p2 cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [d1], [%7, {k1, %9, %10, %11}], [%4];\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%12], [%13, {%14, %15, %16, %11}], [%4];\n\t"
      "selp.u32 %0, 1, 0, pw;\n\t"
      "}"
      : "=r"(ready)
      : "r"(flags), "r"(poll_bar), "r"(poll_parity), "r"(full_bar), "r"(tx_bytes), "r"(dst_dy), "l"(map_dy), "r"(k0),
        "r"(ow0), "r"(oh0), "r"(n0), "r"(dst_x), "l"(map_x), "r"(xc), "r"(xw), "r"(xh)
      : "memory");
  return ready;
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(pred));
  return pred;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, 128-byte swizzle (layout type 2), descriptor version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand in a 64-byte-swizzled tile (layout type 4): rows of 64 bytes, `sbo_bytes` between 8-row groups
__device__ __forceinline__ uint64_t smem_desc64(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (4ull << 61);
}
// K-major operand WITHOUT swizzle (layout type 0): 8-row x 16-byte core matrices; `lbo_bytes` between core matrices that
// are neighbours along K, `sbo_bytes` between neighbours along M.  The strides need not tile the buffer: the row stems
// address overlapping windows of one raw input row this way (lbo 16, sbo 128, see FpropParams::vw_rows)
__device__ __forceinline__ uint64_t smem_desc_plain(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: fp32 accumulate, bf16 A/B
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------ fprop / dgrad kernel
struct FpropParams {
  int N, H, W, C, K, R, S, stride, pad, dil, Ho, Wo;   // H,W,C: tensor read by TMA; Ho,Wo,K: tensor written
  int BW, BH, BNI;                                     // pixel brick of one tile
  int tiles_w, tiles_h, tiles_n, tiles_k;              // tile grid (k = output-channel tiles)
  int BN;                                              // output channels per tile (multiple of 16, <= 256)
  int cblocks, kblocks;                                // C/64, taps*C/64
  int stages, tmem_cols;
  int b_resident;                                      // the whole filter (kblocks slices of BN rows) stays in shared memory
  int k32;                                             // operand k-blocks of 32 elements in 64-byte-swizzled tiles (window-map
                                                       // stems: 8-pixel windows) instead of 64 elements / 128-byte swizzle
  int vw_rows;                                         // row stems (7x7, stride 2, input padded to 4 channels): a tile is 128
                                                       // pixels of ONE output row; its stage holds the seven raw input rows
                                                       // [7][kVwRowPitch] once, and the A operand of tap row r is a no-swizzle
                                                       // descriptor over row r whose 16-byte chunks OVERLAP between pixels
                                                       // (pixel m, chunk c -> raw chunk m + c): no im2col copy anywhere
  int b_mn;                                            // dgrad straight from the untransposed filter W[k][tap][c]: B tiles are
                                                       // MN-major atoms [64 k][64 c], taps read in flipped order
  int n_staging;                                       // 16 KB output staging buffers of the TMA-store epilogue (<= kStaging)
  int split_cb;                                        // B2_CONV_X_CONCAT fprop: channel blocks >= split_cb come from the
                                                       // second input tensor (map_a2); 0 = one input tensor
  int split_k;                                         // B2_CONV_X_CONCAT dgrad: output channels >= split_k go to the second
                                                       // output tensor (map_out2); 0 = one output tensor
  int epi_split;                                       // the eight epilogue warps work as two groups of four on ALTERNATE
                                                       // tiles (group = TMEM accumulator buffer): two epilogue latency
                                                       // chains in flight instead of one (narrow tiles, BN <= 128)
  int scale_mode;                                      // 0 none, 1 partial-conv ratio from mask_in, 2 row_scale[]
  int mask_R, mask_S, mask_stride, mask_pad, mask_dil, mask_H, mask_W;   // window geometry for mode 1
  const float* mask_in;
  const float* row_scale;
  const float* bias;
  float* mask_out;
  float* ratio_out;
  bf16* out;
  int pad_w;                                           // horizontal padding (== pad except in strided dgrad classes)
  int stride_w;                                        // horizontal stride (== stride except for the window-map stems)
  int out_stride_sp, out_off_h, out_off_w;             // strided dgrad: row (oh, ow) is stored at (oh*sp+off_h, ow*sp+off_w)
  int out_H, out_W;                                    // spatial size of the tensor written
  int tma_store;                                       // epilogue: smem-staged TMA store (+ fused BN statistics)
  int accumulate;                                      // TMA reduce-add into the output instead of a plain store
  int debug;                                           // timing experiments (B2POSE_TC_DEBUG): 1 skip the epilogue body, 2 skip the TMA store, 4 skip the fused statistics, 8 no operand loads, 16 no filter loads
  float* bn_sums;                                      // partials[B2_BN_PARTS][2*K]: sum / sum of squares of the stored output
  int bn_totals;                                       // bn_sums is one pre-zeroed float[2*K]: add with fp32 reductions
};

constexpr uint32_t kVwRowChunks = 9, kVwRowPitch = kVwRowChunks * 256;   // raw stem row in a stage: 9 x 256 B >= 131 x 16 B

struct __align__(8) PipeBars {
  uint64_t full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2], bfull;
  uint32_t tmem_base;
};

template <bool kSplit>              // kSplit: FpropParams::epi_split, compiled in (a run-time switch cost the wide tiles 10-50 %)
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_out2, const FpropParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t kbe = (p.k32 || p.vw_rows) ? 32u : (uint32_t)kBlockK;              // elements per operand k-block
  const uint32_t a_bytes = p.vw_rows ? (uint32_t)kABytes : kTileM * kbe * 2, b_bytes = (uint32_t)p.BN * kbe * 2;
  const uint32_t stage_bytes = a_bytes + (p.b_resident ? 0u : b_bytes);
  // layout: [stages][A|B] | resident filter (b_resident: kblocks x B, the ring then holds A only) | kStaging x 16 KB
  //         output staging (TMA-store epilogue) | barriers | fp32 statistics [2*K]
  uint8_t* bres = smem + (size_t)p.stages * stage_bytes;
  uint8_t* staging = bres + (p.b_resident ? (size_t)p.kblocks * b_bytes : 0);
  PipeBars* bars = reinterpret_cast<PipeBars*>(staging + (p.tma_store ? p.n_staging * kABytes : 0));
  float* s_stats = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bars + 1) + 15) & ~(uintptr_t)15);
  if (p.bn_sums)      // [8 epilogue warps][2*BN]: warp-private partials, no shared atomics
    for (int i = threadIdx.x; i < 16 * p.BN; i += kConvThreads) s_stats[i] = 0.f;

  // the warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role branches and everything
  // computed inside them (ring counters, shared-memory addresses, MMA descriptors) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
  const int total_tiles = m_tiles * p.tiles_k;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], kSplit ? 4 : 8); }
    mbar_init(&bars->bfull, 1);
    fence_barrier_init();
    prefetch_map(&map_a);
    prefetch_map(&map_b);
    if (p.tma_store) prefetch_map(&map_out);
    if (p.split_cb) prefetch_map(&map_a2);
    if (p.split_k) prefetch_map(&map_out2);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);      // (warp-uniform for the compiler, see `warp`)
  pdl_wait();            // everything above touched only shared memory / TMEM / kernel parameters

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Converged warp; elect.sync inside produce_fused picks the issuing lane.  No divisions in the loop: (tap row,
    // tap column, channel block) advance as counters.
    const uint32_t a_box_bytes = p.vw_rows ? 7u * kVwRowPitch : (uint32_t)(p.BW * p.BH * p.BNI) * kbe * 2;
    const int ring_kblocks = p.vw_rows ? 1 : p.kblocks;               // ring stages per tile
    if (p.b_resident && lane == 0 && !(p.debug & 24)) {
      // the filter is the same for every tile of this CTA (tiles_k == 1): fetch it once
      mbar_expect_tx(&bars->bfull, (uint32_t)p.kblocks * b_bytes);
      int bcol = 0, cb = 0, tap = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        if (p.b_mn) {
          for (int a = 0; a < (p.BN >> 6); ++a)
            tma_load_2d(smem_u32(bres + (size_t)kb * b_bytes + a * 8192), &map_b, &bars->bfull,
                        (p.R * p.S - 1 - tap) * p.K + a * 64, cb * (int)kbe);
        } else {
          tma_load_2d(smem_u32(bres + (size_t)kb * b_bytes), &map_b, &bars->bfull, bcol, 0);
        }
        bcol += (int)kbe;
        if (++cb == p.cblocks) { cb = 0; ++tap; bcol += p.C - p.cblocks * (int)kbe; }
      }
    }
    __syncwarp();
    const uint32_t stage_flags = p.b_resident ? 0u : kLoadB;
    const uint32_t stage_tx = a_box_bytes + (p.b_resident ? 0u : b_bytes);
    int stage = 0;
    uint32_t phase = 0, empty_ready = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int kt = tile % p.tiles_k, mt = tile / p.tiles_k;
      const int wi = mt % p.tiles_w, hi = (mt / p.tiles_w) % p.tiles_h, ni = mt / (p.tiles_w * p.tiles_h);
      // (row stems: 256-byte units of the padded row -- 128 output pixels are 8 of them)
      const int iw0 = p.vw_rows ? wi * 8 : wi * p.BW * p.stride_w - p.pad_w, ih0 = hi * p.BH * p.stride - p.pad, n0 = ni * p.BNI;
      const bool more_tiles = tile + (int)gridDim.x < total_tiles;
      int r = 0, s = 0, cb = 0, bcol = 0;                 // bcol = tap * C + cb * 64
      for (int kb = 0; kb < ring_kblocks; ++kb) {
        if (!empty_ready) mbar_wait(&bars->empty[stage], phase ^ 1);
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        // channel concatenation of two inputs: the block comes from the tensor that holds it
        const bool second = p.split_cb != 0 && cb >= p.split_cb;
        const CUtensorMap* ma = second ? &map_a2 : &map_a;
        const int ac0 = (second ? cb - p.split_cb : cb) * (int)kbe;
        if (p.debug & 24) {
          empty_ready = 0;
          if (lane == 0) {
            if (p.debug & 8) {                     // timing experiment: no operand loads at all
              mbar_arrive(&bars->full[stage]);
            } else {                               // timing experiment: activation tile only (filter "resident")
              mbar_expect_tx(&bars->full[stage], a_box_bytes);
              tma_load_4d(sa, ma, &bars->full[stage], ac0, iw0 + s * p.dil, ih0 + r * p.dil, n0);
            }
          }
          __syncwarp();
        } else {
          const bool has_next = kb + 1 < ring_kblocks || more_tiles;
          const int mn_col = (p.R * p.S - 1 - (r * p.S + s)) * p.K + kt * p.BN;     // b_mn: flipped tap, first atom
          empty_ready = produce_fused((has_next ? 1u : 0u) | stage_flags, smem_u32(&bars->empty[nstage]), nphase ^ 1,
                                      smem_u32(&bars->full[stage]), stage_tx, sa, ma, ac0,
                                      iw0 + s * p.dil, ih0 + r * p.dil, n0, sa + a_bytes, &map_b,
                                      p.b_mn ? mn_col : bcol, p.b_mn ? cb * (int)kbe : kt * p.BN);
          if (p.b_mn && !p.b_resident && p.BN > 64) {
            if (elect_one())
              for (int a = 1; a < (p.BN >> 6); ++a)
                tma_load_2d(sa + a_bytes + a * 8192, &map_b, &bars->full[stage], mn_col + a * 64, cb * (int)kbe);
            __syncwarp();
          }
        }
        bcol += (int)kbe;
        if (++cb == p.cblocks) {
          cb = 0;
          bcol += p.C - p.cblocks * (int)kbe;      // ragged last channel block: the next tap starts at tap * C
          if (++s == p.S) { s = 0; ++r; }
        }
        stage = nstage;
        phase = nphase;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp stays converged (elect.sync inside mma4_fused picks the issuing lane).  Every group of four MMAs
    // polls the barrier the next group needs (`full` of the next ring stage, or the `tempty` of the next tile's
    // accumulator when this is a tile's last k-block), so the blocking waits below normally fall through.
    const uint32_t idesc = instr_desc(kTileM, p.BN, 0, p.b_mn);
    // descriptor = launch constant (swizzle mode, LBO / SBO) + the tile's shared-memory address in 16-byte units.
    // B: K-major filter slice [BN][64], or (b_mn) BN / 64 MN-major atoms [64 k][64 c] straight from the untransposed
    // filter: LBO = one 8 KB atom, 16 k-rows per MMA = 2048 B; k32: 64-byte rows in 64-byte-swizzled tiles (8-row groups
    // 512 B apart), two MMAs per stage
    const uint64_t a_const = p.k32 ? smem_desc64(0, 512) : smem_desc(0, 0, 1024);
    const uint64_t b_const = p.k32 ? smem_desc64(0, 512) : (p.b_mn ? smem_desc(0, 8192, 1024) : smem_desc(0, 0, 1024));
    const uint64_t b_kstep = p.b_mn ? 128ull : 2ull;
    const uint32_t mode_flags = p.k32 ? kTwoMma : 0u;
    const uint32_t ring_lo = smem_u32(smem) >> 4, stage_step = stage_bytes >> 4, a_off = a_bytes >> 4;
    const uint32_t bres_lo = smem_u32(bres) >> 4, b_slice = b_bytes >> 4;
    const uint32_t bar_full = smem_u32(&bars->full[0]), bar_empty = smem_u32(&bars->empty[0]);
    const uint32_t bar_tfull = smem_u32(&bars->tfull[0]), bar_tempty = smem_u32(&bars->tempty[0]);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    uint32_t full_ready = 0, acc_ready = 0;
    if (p.b_resident && !(p.debug & 24) && blockIdx.x < total_tiles) mbar_wait(&bars->bfull, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      if (!acc_ready) mbar_wait(&bars->tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
      const bool more_tiles = tile + (int)gridDim.x < total_tiles;
      const int nacc = (local + 1) & 1;
      const uint32_t nacc_parity = (((local + 1) >> 1) & 1) ^ 1;
      if (p.vw_rows) {
        // one stage = the tile's seven raw input rows; tap row r: two K=16 MMAs over the overlapping 16-byte chunks of
        // raw row r (pixel m, chunk c at (m + c) * 16: K-neighbours 16 B apart, 8-pixel groups 128 B apart)
        if (!full_ready) mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        mma14_rows(tmem_d, smem_desc_plain(sa, 16, 128), smem_desc64(smem_u32(bres), 512), (uint64_t)(kVwRowPitch >> 4),
                   (uint64_t)(b_bytes >> 4), idesc, more_tiles ? (kPoll1 | kPoll2) : 0u, smem_u32(&bars->full[nstage]), nphase,
                   smem_u32(&bars->tempty[nacc]), nacc_parity, smem_u32(&bars->empty[stage]), smem_u32(&bars->tfull[acc]),
                   full_ready, acc_ready);
        stage = nstage;
        phase = nphase;
        continue;
      }
      // (everything that depends only on the launch -- descriptor constants, strides, the k32 / b_mn / resident-filter
      //  modes -- is computed once above the tile loop: the issue loop is ~100 uniform-datapath instructions per stage
      //  and bounds the narrow layers, ncu source view of the 3x3 64-channel fprop)
      uint32_t b_res = bres_lo;                        // resident filter: slice kb
      int kb = 0;
      if (mode_flags == 0u) {
        // all but the last k-block: switches fixed at compile time (mma4_steady)
        for (; kb < p.kblocks - 1; ++kb) {
          if (!full_ready) mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
          const uint32_t a_lo = ring_lo + (uint32_t)stage * stage_step;
          const uint32_t b_lo = p.b_resident ? b_res : a_lo + a_off;
          b_res += b_slice;
          mma4_steady(tmem_d, a_const + a_lo, b_const + b_lo, idesc, (uint32_t)kb, bar_full + 8u * (uint32_t)nstage, nphase,
                      bar_empty + 8u * (uint32_t)stage, full_ready, b_kstep);
          stage = nstage;
          phase = nphase;
        }
      }
      for (; kb < p.kblocks; ++kb) {
        if (!full_ready) mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
        const bool last = kb == p.kblocks - 1;
        const uint32_t flags = mode_flags | ((!last || more_tiles) ? kPoll1 : 0u) | ((last && more_tiles) ? kPoll2 : 0u) |
                               kCommit1 | (last ? kCommit2 : 0u);
        const uint32_t a_lo = ring_lo + (uint32_t)stage * stage_step;
        const uint32_t b_lo = p.b_resident ? b_res : a_lo + a_off;
        b_res += b_slice;
        mma4_fused(tmem_d, a_const + a_lo, b_const + b_lo, 2ull, idesc, (uint32_t)kb, flags,
                   bar_full + 8u * (uint32_t)nstage, nphase, bar_tempty + 8u * (uint32_t)nacc, nacc_parity,
                   bar_empty + 8u * (uint32_t)stage, bar_tfull + 8u * (uint32_t)acc, full_ready, acc_ready, b_kstep);
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quarter: warp (q, half) owns rows 32q..32q+31 and the `half` 32-column
    // part of every 64-column group.
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    constexpr bool split = kSplit;                // two four-warp groups on alternate tiles (FpropParams::epi_split)
    const int grp = split ? (warp - 2) >> 2 : 0;
    const int half = split ? 0 : (warp - 2) >> 2; // (split: a warp converts both 32-column halves of a group)
    const int ew = warp - 2;                      // epilogue warp index 0..7
    const int ep_tid = split ? ((threadIdx.x - 64) & 127) : (threadIdx.x - 64);   // index within the (group's) epilogue threads
    const int lstep = split ? 2 : 1;              // tiles between two iterations of this warp
    const int row = q * 32 + lane;
    const int brick = p.BW * p.BH;
    const int groups = p.BN >> 6;
    const int nsets = p.tma_store ? p.n_staging / groups : 1;   // staging sets of `groups` 16 KB buffers
    uint64_t st_sum[2][4], st_sq[2][4];           // fused BatchNorm statistics (packed fp32x2), see below
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) st_sum[k][e] = st_sq[k][e] = 0ull;
    // Row geometry of this thread in a tile, and the mask values its renormalisation ratio needs.  The mask loads of
    // tile i + 1 are ISSUED while tile i is processed and consumed at the top of the next iteration (ncu: the
    // dependent add right behind the load was 9 % of all stall samples of the partial 1x1 layers).
    struct RowGeo { int kt, wi, hi, ni, n, oh, ow; bool valid; long long pix, opix; };
    const int row_bn = row / brick, row_rem = row - row_bn * brick;
    const int row_bh = row_rem / p.BW, row_bw = row_rem - row_bh * p.BW;
    // Tile coordinates (kt, wi, hi, ni) advance by the mixed-radix digits of gridDim.x: no divisions per tile (the
    // epilogue of the 64-channel layers is bound by its per-tile instruction count: ncu counted ~580 warp instructions
    // per tile and warp, five integer divisions among them, for 32 values per thread)
    int step_k, step_w, step_h, step_n;
    {
      int t = (int)gridDim.x;
      step_k = t % p.tiles_k; t /= p.tiles_k;
      step_w = t % p.tiles_w; t /= p.tiles_w;
      step_h = t % p.tiles_h; step_n = t / p.tiles_h;
    }
    struct TilePos { int kt, wi, hi, ni; };
    auto advance = [&](TilePos& t) {
      t.kt += step_k;
      int c = t.kt >= p.tiles_k ? 1 : 0;
      t.kt -= c ? p.tiles_k : 0;
      t.wi += step_w + c;
      c = t.wi >= p.tiles_w ? 1 : 0;
      t.wi -= c ? p.tiles_w : 0;
      t.hi += step_h + c;
      c = t.hi >= p.tiles_h ? 1 : 0;
      t.hi -= c ? p.tiles_h : 0;
      t.ni += step_n + c;
    };
    auto geometry = [&](const TilePos& t) {
      RowGeo g;
      g.kt = t.kt; g.wi = t.wi; g.hi = t.hi; g.ni = t.ni;
      g.n = g.ni * p.BNI + row_bn; g.oh = g.hi * p.BH + row_bh; g.ow = g.wi * p.BW + row_bw;
      const int sh = g.oh * p.out_stride_sp + p.out_off_h, sw = g.ow * p.out_stride_sp + p.out_off_w;
      g.valid = (row_bn < p.BNI) && (g.n < p.N) && (g.oh < p.Ho) && (g.ow < p.Wo) && (sh < p.out_H) && (sw < p.out_W);
      g.pix = ((long long)g.n * p.Ho + g.oh) * p.Wo + g.ow;
      g.opix = ((long long)g.n * p.out_H + sh) * p.out_W + sw;
      return g;
    };
    const int mask_taps = p.mask_R * p.mask_S;
    const bool prefetch_mask = p.scale_mode == 2 || (p.scale_mode == 1 && mask_taps <= 9);
    auto load_mask = [&](const RowGeo& g, float (&mv)[9]) {       // issue only: nothing here depends on the values
#pragma unroll
      for (int t = 0; t < 9; ++t) mv[t] = 0.f;
      if (!g.valid || !prefetch_mask) return;
      if (p.scale_mode == 2) { mv[0] = __ldg(p.row_scale + g.opix); return; }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (t < mask_taps) {
          const int r = t / p.mask_S, s2 = t - r * p.mask_S;
          const int ih = g.oh * p.mask_stride - p.mask_pad + r * p.mask_dil;
          const int iw = g.ow * p.mask_stride - p.mask_pad + s2 * p.mask_dil;
          if (ih >= 0 && ih < p.mask_H && iw >= 0 && iw < p.mask_W)
            mv[t] = __ldg(p.mask_in + ((long long)g.n * p.mask_H + ih) * p.mask_W + iw);
        }
      }
    };
    TilePos pos;
    const long long first_tile = (long long)blockIdx.x + (long long)grp * gridDim.x;
    {
      int t = first_tile < total_tiles ? (int)first_tile : 0;
      pos.kt = t % p.tiles_k; t /= p.tiles_k;
      pos.wi = t % p.tiles_w; t /= p.tiles_w;
      pos.hi = t % p.tiles_h; pos.ni = t / p.tiles_h;
    }
    RowGeo geo = geometry(pos);
    float mv[9];
    load_mask(geo, mv);
    int local = grp, sset_idx = grp;              // staging set: local % nsets (a group keeps to its own sets)
    for (long long tile = first_tile; tile < total_tiles; tile += (long long)lstep * gridDim.x, local += lstep) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int kt = geo.kt, wi = geo.wi, hi = geo.hi, ni = geo.ni;
      const bool valid = geo.valid;
      const long long opix = geo.opix;
      float scale = 1.f, mo = 1.f;
      if (valid) {
        if (p.scale_mode == 1) {
          float cnt = 0.f;
          if (prefetch_mask) {
#pragma unroll
            for (int t = 0; t < 9; ++t) cnt += mv[t];            // same order as the loop below (row-major taps)
          } else {
            for (int r = 0; r < p.mask_R; ++r) {
              const int ih = geo.oh * p.mask_stride - p.mask_pad + r * p.mask_dil;
              if (ih < 0 || ih >= p.mask_H) continue;
              for (int s2 = 0; s2 < p.mask_S; ++s2) {
                const int iw = geo.ow * p.mask_stride - p.mask_pad + s2 * p.mask_dil;
                if (iw < 0 || iw >= p.mask_W) continue;
                cnt += __ldg(p.mask_in + ((long long)geo.n * p.mask_H + ih) * p.mask_W + iw);
              }
            }
          }
          scale = pconv_ratio((float)mask_taps, cnt);
          mo = fminf(fmaxf(cnt, 0.f), 1.f);
          if (kt == 0 && half == 0) {
            if (p.mask_out) p.mask_out[geo.pix] = mo;
            if (p.ratio_out) p.ratio_out[geo.pix] = scale;
          }
        } else if (p.scale_mode == 2) {
          scale = mv[0];
        }
      }
      {                                                           // next tile: geometry now, mask loads in flight
        const long long nt = tile + (long long)lstep * gridDim.x;
        if (nt < total_tiles) {
          advance(pos);
          if (split) advance(pos);
          geo = geometry(pos);
          load_mask(geo, mv);
        }
      }
      mbar_wait(&bars->tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
      const int kbase = kt * p.BN;
      if (p.debug & 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tempty[acc]);
        continue;
      }
      if (p.tma_store) {
        // ---- smem-staged epilogue: the whole tile goes to `groups` swizzled 16 KB staging buffers
        //      (one per 64 output channels), then one thread issues the TMA stores; the staged
        //      (rounded) values also feed the per-channel BatchNorm partial sums.
        uint8_t* sset = staging + (size_t)(sset_idx * groups) * kABytes;
        sset_idx += lstep;
        if (sset_idx >= nsets) sset_idx = grp;
        if (ep_tid == 0) {                       // the stores this thread issued (nsets / lstep) tiles ago have drained this set
          const int mine = nsets / lstep;        // staging sets this issuing thread cycles through
          if (mine >= 4) bulk_wait_read<3>();
          else if (mine == 2) bulk_wait_read<1>();
          else bulk_wait_read<0>();
        }
        if constexpr (split) epi_barrier_group(grp); else epi_barrier();
        // software-pipelined TMEM reads: the loads of group g+1 are in flight while group g is converted.
        // The epilogue is instruction-issue bound on the shallow layers (8 warps on 4 schedulers), so the
        // conversion uses packed fp32x2 multiplies and skips the multiply when there is no row scale.
        // Out-of-range rows of a ragged tile are clipped by the TMA store but NOT zero in TMEM for R, S > 1 (output
        // row Ho reads the valid input row Ho - 1 through tap r = 0), and the fused BatchNorm statistics below read the
        // staged tile: such rows take the multiply path with scale 0 (round 1 skipped it and polluted the statistics
        // of plain 3x3 layers whose output is not a multiple of the brick, e.g. every 257x257 configuration).
        const float sc = valid ? scale : 0.f;
        const uint64_t sc2 = pack2(sc, sc);
        const uint32_t sset_a = smem_u32(sset);
        const uint32_t row_off = (uint32_t)row * 128u;
        const uint32_t swz = (uint32_t)(row & 7);
        // convert one 32-column half-group held in registers and store it to its swizzled staging rows (explicit
        // shared-space stores; the register arrays are indexed statically -- round 1's `(g & 1) ? vb : va` selection
        // cost one SEL per value and group, 20 % of the kernel's instructions in ncu)
        auto convert_store = [&](const uint32_t (&v)[32], int step) {
          const int g = split ? step >> 1 : step, half = split ? step & 1 : (warp - 2) >> 2;
          const uint32_t sbuf = sset_a + (uint32_t)g * kABytes + row_off;
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // always multiply (scale 1 when the layer has no row scale): a conditional in-place multiply costs two
              // register moves per pair, more than the FMUL2 it saves
              float a, b;
              unpack2(mul2(pack2(__uint_as_float(v[j2 * 8 + 2 * e]), __uint_as_float(v[j2 * 8 + 2 * e + 1])), sc2), a, b);
              w[e] = f32x2_to_bf16x2(a, b);
            }
            sts128(sbuf + ((((uint32_t)(half * 4 + j2)) ^ swz) << 4), w[0], w[1], w[2], w[3]);
          }
        };
        auto load_group = [&](uint32_t (&v)[32], int step) {       // split: step = 2 * group + half = 32-column block
          const uint32_t col = split ? (uint32_t)step * 32u : (uint32_t)(step * 64 + half * 32);
          tmem_ld16(taddr + col, v);
          tmem_ld16(taddr + col + 16, v + 16);
        };
        const int nsteps = split ? 2 * groups : groups;
        // software-pipelined TMEM reads, fully unrolled over the (at most four) 64-column groups: the loads of
        // group g + 1 are in flight while group g is converted
        uint32_t va[32], vb[32];
        load_group(va, 0);
#pragma unroll
        for (int g2 = 0; g2 < 4; g2 += 2) {
          if (g2 < nsteps) {
            tmem_ld_wait();
            if (g2 + 1 < nsteps) load_group(vb, g2 + 1);
            convert_store(va, g2);
          }
          if (g2 + 1 < nsteps) {
            tmem_ld_wait();
            if (g2 + 2 < nsteps) load_group(va, g2 + 2);
            convert_store(vb, g2 + 1);
          }
        }
        tc_fence_before();                        // accumulator fully read: hand TMEM back to the MMA warp
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tempty[acc]);
        fence_async_smem();
        if constexpr (split) epi_barrier_group(grp); else epi_barrier();
        if (ep_tid == 0 && !(p.debug & 2)) {
          for (int g = 0; g < groups; ++g) {
            int ch = kbase + g * 64;
            const CUtensorMap* mo = &map_out;
            if (p.split_k != 0 && ch >= p.split_k) { mo = &map_out2; ch -= p.split_k; }      // second half of a concatenation
            if (p.accumulate)
              tma_reduce_add_4d(mo, smem_u32(sset + (size_t)g * kABytes), ch, wi * p.BW, hi * p.BH, ni * p.BNI);
            else
              tma_store_4d(mo, smem_u32(sset + (size_t)g * kABytes), ch, wi * p.BW, hi * p.BH, ni * p.BNI);
          }
          bulk_commit();
        }
        if (p.bn_sums && !(p.debug & 4)) {
          // per-channel sum / sum of squares of the staged (rounded) tile.  Thread (chunk, ghalf, rsub)
          // owns 8 channels of the 64-channel groups ghalf and ghalf + 2 and rows rsub, rsub + 16, ...:
          // 16-byte shared loads (a quarter-warp reads one whole 128-byte row), packed fp32x2 adds / fmas
          // into registers that live across ALL tiles of this CTA (tiles_k == 1) -- no shuffles, no
          // shared-memory read-modify-writes per tile.
          const int chunk = ep_tid & 7, ghalf = (ep_tid >> 3) & 1, rsub = ep_tid >> 4;        // rsub 0..15
          const int nrows = min(brick * p.BNI, kTileM);
          const uint32_t off = (uint32_t)rsub * 128u + (uint32_t)((chunk ^ (rsub & 7)) << 4);   // (rsub + 16 i) & 7 == rsub & 7
          if (split) {
            // four-warp group: thread (chunk, rsub) owns 8 channels of every 64-channel group (at most two here) and
            // rows rsub, rsub + 16, ...
            const int rs = ep_tid >> 3;                                                       // 0..15
            const uint32_t off4 = (uint32_t)rs * 128u + (uint32_t)((chunk ^ (rs & 7)) << 4);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (k < groups) {
                const uint32_t sbuf = sset_a + (uint32_t)k * kABytes + off4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  uint4 u = make_uint4(0u, 0u, 0u, 0u);
                  if (rs + 16 * i < nrows) u = lds128(sbuf + i * 2048);
                  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const uint64_t t = bf16x2_to_f32x2(w[e]);
                    st_sum[k][e] = add2(st_sum[k][e], t);
                    st_sq[k][e] = fma2(t, t, st_sq[k][e]);
                  }
                }
              }
            }
          } else if (groups == 1) {
            // one 64-channel group: both half-warps work on it, rows rsub + 16 i with i in [4 ghalf, 4 ghalf + 4)
            // (combined by the extra shuffle at the end)
            const uint32_t sbuf = sset_a + off + (uint32_t)ghalf * 4u * 2048u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u = make_uint4(0u, 0u, 0u, 0u);
              if (rsub + 16 * (i + 4 * ghalf) < nrows) u = lds128(sbuf + i * 2048);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint64_t t = bf16x2_to_f32x2(w[e]);
                st_sum[0][e] = add2(st_sum[0][e], t);
                st_sq[0][e] = fma2(t, t, st_sq[0][e]);
              }
            }
          } else
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int g = ghalf + 2 * k;
            if (g < groups) {
              const uint32_t sbuf = sset_a + (uint32_t)g * kABytes + off;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // rows past the tile contribute zeros: the accumulation itself is unconditional (a guarded update
                // compiled to a branch per row plus sixteen register moves)
                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                if (rsub + 16 * i < nrows) u = lds128(sbuf + i * 2048);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint64_t t = bf16x2_to_f32x2(w[e]);
                  st_sum[k][e] = add2(st_sum[k][e], t);
                  st_sq[k][e] = fma2(t, t, st_sq[k][e]);
                }
              }
            }
          }
        }
        continue;
      }
      // ---- direct-store epilogue (ragged channel tiles, strided scatter of dgrad)
      for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
        uint32_t v[32];
        const bool two = (c0 + 16) < p.BN;
        tmem_ld16(taddr + c0, v);
        if (two) tmem_ld16(taddr + c0 + 16, v + 16);
        tmem_ld_wait();
        if (valid) {
          const int ncol = two ? 32 : 16;
          bf16* orow = p.out + opix * p.K + kbase + c0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g * 8 >= ncol) break;
            const int col = kbase + c0 + g * 8;
            if (col >= p.K) break;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float t = __uint_as_float(v[g * 8 + j]) * scale;
              if (p.bias) t = (t + __ldg(p.bias + col + j)) * mo;
              f[j] = t;
            }
            store8(orow + g * 8, f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[acc]);
    }
    if (p.tma_store && ep_tid == 0) bulk_wait_read<0>();           // staging must outlive the last store(s)
    if (p.bn_sums) {
      // lanes l and l ^ 16 hold the same channels (rows rsub and rsub ^ 1): combine, then the low half-warp
      // writes this warp's partial; the 8 warp partials are summed in a fixed order below (deterministic)
      const int chunk = ep_tid & 7, ghalf = (ep_tid >> 3) & 1;
      float* mine = s_stats + (size_t)ew * 2 * p.BN;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float su[8], sq[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          unpack2(st_sum[k][e], su[2 * e], su[2 * e + 1]);
          unpack2(st_sq[k][e], sq[2 * e], sq[2 * e + 1]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          su[e] += __shfl_xor_sync(0xffffffffu, su[e], 16);
          sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 16);
        }
        if (groups == 1 || split) {               // lanes l and l ^ 8 shared a channel group as well
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            su[e] += __shfl_xor_sync(0xffffffffu, su[e], 8);
            sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 8);
          }
        }
        const int g = split ? k : ghalf + 2 * k;
        if ((split ? lane < 8 : lane < 16) && g < groups) {
          const int ch0 = g * 64 + chunk * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) { mine[ch0 + e] = su[e]; mine[p.BN + ch0 + e] = sq[e]; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (p.bn_sums) {      // this CTA's slot of partials[B2_BN_PARTS][2K]; the unused slots are zero-filled
    // With several channel tiles the grid is a multiple of tiles_k, so every tile of this CTA has the same kt =
    // blockIdx.x % tiles_k and the register accumulators above belong to channels [kt * BN, (kt + 1) * BN).
    const int kt0 = blockIdx.x % p.tiles_k, c0 = kt0 * p.BN;
    float* slot = p.bn_sums + (p.bn_totals ? (size_t)0 : (size_t)blockIdx.x * 2 * p.K);
    if (!p.bn_totals && p.tiles_k > 1)
      for (int i = threadIdx.x; i < 2 * p.K; i += kConvThreads) slot[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * p.BN; i += kConvThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_stats[(size_t)w * 2 * p.BN + i];
      const int q = i >= p.BN ? 1 : 0, c = c0 + (i - q * p.BN);
      if (c < p.K) {
        if (p.bn_totals) atomicAdd(slot + q * p.K + c, t);
        else slot[q * p.K + c] = t;
      }
    }
    if (!p.bn_totals)
      for (int sl = gridDim.x + blockIdx.x; sl < B2_BN_PARTS; sl += gridDim.x)
        for (int i = threadIdx.x; i < 2 * p.K; i += kConvThreads) p.bn_sums[(size_t)sl * 2 * p.K + i] = 0.f;
  }
}

// ------------------------------------------------------------------ wgrad kernel
// Work item = (pixel split, tap, c-tile of BNc channels, k-tile of 128 output channels):
// D[128 k, BNc c] accumulated in TMEM over the item's pixel bricks, then red.add into dw (fp32).
struct WgradParams {
  int N, H, W, C, K, R, S, stride, pad, dil, Ho, Wo;
  int stride_w, pad_w;             // horizontal stride / padding (== stride / pad except for the window-map stems)
  int BW, BH, BNI;                 // pixel brick per pipeline stage (BW*BH*BNI = 64 pixels incl. ragged rows)
  int tiles_w, tiles_h, tiles_n;   // bricks over the OUTPUT pixel space
  int BNc, ctiles, ktiles, splits; // columns per tap, ceil(C/BNc), ceil(K/128), pixel splits
  int T, tgroups;                  // filter taps per work item (they share the dY tile), ceil(taps/T)
  int bricks_per_split;
  int stages, tmem_cols;
  int split_c;                     // B2_CONV_X_CONCAT: input channels >= split_c come from the second tensor (map_x2)
  int debug;                       // timing experiments (B2POSE_TC_DEBUG): 1 no reduction into dW, 8 no operand loads, 32 no MMAs
  int halo;                        // 3x3 stride-1 pad-1 layers with 64 channels whose brick is 64 pixels of ONE image row:
                                   // the three taps of a filter row read ONE 72-pixel X tile (pixels w0 - 1 .. w0 + 70).
                                   // Tap s is the same tile shifted by s pixels = s x 128 B; the 128-byte swizzle follows the
                                   // absolute shared-memory address, so the three taps are the three 64-column atoms of ONE
                                   // N = 192 MN-major operand with LBO = 128 B (matrix base offset 0)
  int vw_rows;                     // row stems: a stage holds the brick's seven raw input rows [7][kVwgRowPitch]; the X
                                   // operand of tap row r is an MN-major no-swizzle descriptor over row r (pixel p,
                                   // window chunk c -> raw chunk p + c), 32 window elements per tap row; 1 / 2 selects
                                   // which descriptor field carries which stride
  float* dw;                       // [K][R*S][C] fp32
};
constexpr int kWgPix = 64;         // pixels per stage
constexpr uint32_t kVwgRowUnits = 5, kVwgRowPitch = kVwgRowUnits * 256;   // raw stem row of a 64-pixel brick: 67 x 16 B
constexpr uint32_t kVwgRowsBytes = 9 * 1024;                              // seven of them, rounded to the swizzle atom
constexpr uint32_t kHaloPix = 72, kHaloBytes = kHaloPix * 128;            // halo X tile of a 64-pixel brick (9 swizzle atoms)

// Work item = (pixel split, tap group, c-tile, k-tile).  All T taps of a group read the same dY tile
// (one TMA fetch) and their own shifted X tile; their accumulators sit side by side in TMEM.
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_x2, const WgradParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // stage: A = dy [2 atoms of 64 k][64 pix][128 B]  (16 KB), B = T x ( x [BNc/64 atoms][64 pix][128 B] )
  const uint32_t atom_bytes = kWgPix * 128;
  const uint32_t a_bytes = 2 * atom_bytes, b_bytes = (uint32_t)(p.BNc / 64) * atom_bytes;
  const uint32_t stage_bytes = a_bytes + (p.vw_rows ? kVwgRowsBytes : p.halo ? kHaloBytes : (uint32_t)p.T * b_bytes);
  PipeBars* bars = reinterpret_cast<PipeBars*>(smem + (size_t)p.stages * stage_bytes);
  // the warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role branches and everything
  // computed inside them (ring counters, shared-memory addresses, MMA descriptors) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int taps = p.R * p.S;
  const int total_items = p.splits * p.tgroups * p.ctiles * p.ktiles;
  const int total_bricks = p.tiles_n * p.tiles_h * p.tiles_w;
  const int acc_cols = p.T * p.BNc;          // TMEM columns of one accumulator set

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 4); }
    fence_barrier_init();
    prefetch_map(&map_dy);
    prefetch_map(&map_x);
    if (p.split_c) prefetch_map(&map_x2);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, (uint32_t)p.tmem_cols);
  // K <= 64: output channels 64..127 of the dY^T operand do not exist.  Their atom is cleared here once and never
  // loaded (8 KB less operand feed per stage; the MMA still reads M = 128 rows)
  const bool half_dy = p.K <= 64;
  if (half_dy) {
    for (int st = 0; st < p.stages; ++st)
      for (uint32_t i = threadIdx.x * 16u; i < atom_bytes; i += kThreads * 16u)
        sts128(smem_u32(smem + (size_t)st * stage_bytes + atom_bytes) + i, 0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);      // (warp-uniform for the compiler, see `warp`)
  pdl_wait();            // everything above touched only shared memory / TMEM / kernel parameters
  const uint32_t dy_flag = half_dy ? 2u : 0u, dy_tx = half_dy ? atom_bytes : a_bytes;

  // item decode: split fastest so CTAs running together share the same filter tile / spread pixels
  auto decode = [&](int item, int& sp, int& tg, int& ct, int& kt) {
    sp = item % p.splits; item /= p.splits;
    ct = item % p.ctiles; item /= p.ctiles;
    tg = item % p.tgroups; kt = item / p.tgroups;
  };

  if (warp == 0) {
    // TMA producer: converged warp (elect.sync picks the issuing lane), brick coordinates advance as counters, the
    // next stage's `empty` barrier is polled while this stage's loads are issued (see produce_fused).
    static_assert(kWgPix * 128 == 8192, "produce_wgrad_fused assumes 8 KB atoms");
    int stage = 0;
    uint32_t phase = 0, empty_ready = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int sp, tg, ct, kt;
      decode(item, sp, tg, ct, kt);
      const int tap0 = tg * p.T, nt = min(p.T, taps - tap0);
      const int b0 = sp * p.bricks_per_split;
      const int b1 = min(total_bricks, b0 + p.bricks_per_split);
      const bool more_items = item + (int)gridDim.x < total_items;
      int wi = b0 % p.tiles_w, hi = (b0 / p.tiles_w) % p.tiles_h, ni = b0 / (p.tiles_w * p.tiles_h);
      const int r0 = tap0 / p.S, s0 = tap0 - r0 * p.S;
      const int atoms = p.BNc / 64;
      for (int b = b0; b < b1; ++b) {
        const int ow0 = wi * p.BW, oh0 = hi * p.BH, n0 = ni * p.BNI;
        if (!empty_ready) mbar_wait(&bars->empty[stage], phase ^ 1);
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t fb = smem_u32(&bars->full[stage]);
        const int xw0 = ow0 * p.stride_w - p.pad_w, xh0 = oh0 * p.stride - p.pad;
        const bool has_next = b + 1 < b1 || more_items;
        if (p.debug & 8) {                           // timing experiment: no operand loads
          empty_ready = 0;
          if (lane == 0) mbar_arrive(&bars->full[stage]);
          __syncwarp();
          if (++wi == p.tiles_w) { wi = 0; if (++hi == p.tiles_h) { hi = 0; ++ni; } }
          stage = nstage;
          phase = nphase;
          continue;
        }
        // dy atoms: k channels [kt*128, +64) and [+64, +128); first x atom of the first tap
        // (channel concatenation of two inputs: a channel tile lies in one of them, split_c % BNc == 0)
        const bool second = p.split_c != 0 && ct * p.BNc >= p.split_c;
        const CUtensorMap* mx = second ? &map_x2 : &map_x;
        const int xc0 = ct * p.BNc - (second ? p.split_c : 0);
        if (p.halo)        // one box: 72 pixels of input row oh0 - 1 + r0, starting one pixel left of the brick
          empty_ready = produce_wgrad_fused((has_next ? 1u : 0u) | dy_flag, smem_u32(&bars->empty[nstage]), nphase ^ 1, fb,
                                            dy_tx + kHaloBytes, sa, &map_dy, kt * 128, ow0, oh0, n0,
                                            sa + a_bytes, mx, xc0, xw0, xh0 + r0 * p.dil);
        else if (p.vw_rows)     // one box: the seven raw rows (256-byte units of the padded row, 64 pixels = 4 units)
          empty_ready = produce_wgrad_fused((has_next ? 1u : 0u) | dy_flag, smem_u32(&bars->empty[nstage]), nphase ^ 1, fb,
                                            dy_tx + 7u * kVwgRowPitch, sa, &map_dy, kt * 128, ow0, oh0, n0,
                                            sa + a_bytes, &map_x, 0, ow0 >> 4, oh0 * 2);
        else
        empty_ready = produce_wgrad_fused((has_next ? 1u : 0u) | dy_flag, smem_u32(&bars->empty[nstage]), nphase ^ 1, fb,
                                          dy_tx + (uint32_t)nt * b_bytes, sa, &map_dy, kt * 128, ow0, oh0, n0,
                                          sa + a_bytes, mx, xc0, xw0 + s0 * p.dil, xh0 + r0 * p.dil);
        if (nt * atoms > 1 && !p.halo) {
          if (elect_one()) {
            int r = r0, s = s0;
            for (int j = 0; j < nt; ++j) {
              for (int a = (j == 0 ? 1 : 0); a < atoms; ++a)
                tma_load_4d(sa + a_bytes + j * b_bytes + a * atom_bytes, mx, &bars->full[stage], xc0 + a * 64,
                            xw0 + s * p.dil, xh0 + r * p.dil, n0);
              if (++s == p.S) { s = 0; ++r; }
            }
          }
          __syncwarp();
        }
        if (++wi == p.tiles_w) { wi = 0; if (++hi == p.tiles_h) { hi = 0; ++ni; } }
        stage = nstage;
        phase = nphase;
      }
    }
  } else if (warp == 1) {
    // MMA issuer: converged warp, one fused asm block per tap (see mma4_fused); the first tap's block polls the next
    // stage's `full` barrier (and the next item's accumulator when this is the item's last brick), the last tap's
    // block commits.
    const uint32_t idesc = instr_desc(kTileM, p.BNc, 1, 1);
    // dY: MN-major, 128B swizzle: LBO = bytes between 64-element atoms along M/N, SBO = 1024 (8 pixel rows); 16 pixels per
    // MMA = 2 swizzle row groups = 2048 B = descriptor step 128.  X: the same, or (row stems) MN-major without swizzle --
    // window chunks 16 B apart along N, 8-pixel groups 128 B apart along K, 16 pixels (256 B) per MMA
    const uint64_t a_const = smem_desc(0, atom_bytes, 1024);
    const uint64_t b_const = p.vw_rows ? (p.vw_rows == 1 ? smem_desc_plain(0, 128, 16) : smem_desc_plain(0, 16, 128))
                             : p.halo  ? smem_desc(0, 128, 1024)        // the three taps = three atoms one pixel row apart
                                       : smem_desc(0, atom_bytes, 1024);
    const uint64_t b_kstep = p.vw_rows ? 16ull : 128ull;
    const uint32_t ring_lo = smem_u32(smem) >> 4, stage_step = stage_bytes >> 4, a_off = a_bytes >> 4;
    const uint32_t b_slice = (p.vw_rows ? kVwgRowPitch : b_bytes) >> 4;
    const uint32_t bar_full = smem_u32(&bars->full[0]), bar_empty = smem_u32(&bars->empty[0]);
    const uint32_t bar_tfull = smem_u32(&bars->tfull[0]), bar_tempty = smem_u32(&bars->tempty[0]);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    uint32_t full_ready = 0, acc_ready = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++local) {
      int sp, tg, ct, kt;
      decode(item, sp, tg, ct, kt);
      const int nt = min(p.T, taps - tg * p.T);
      const int b0 = sp * p.bricks_per_split;
      const int b1 = min(total_bricks, b0 + p.bricks_per_split);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      if (!acc_ready) mbar_wait(&bars->tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
      const bool more_items = item + (int)gridDim.x < total_items;
      const int nacc = (local + 1) & 1;
      const uint32_t nacc_parity = (((local + 1) >> 1) & 1) ^ 1;
      for (int b = b0; b < b1; ++b) {
        if (!full_ready) mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
        const bool last = b == b1 - 1;
        // descriptor = launch constant + shared-memory address in 16-byte units (constants hoisted above the item loop)
        const uint32_t a_lo = ring_lo + (uint32_t)stage * stage_step;
        const uint64_t adesc = a_const + a_lo;
        uint32_t b_lo = a_lo + a_off;
        uint32_t r1 = 0, r2 = 0;
        if (p.debug & 32) {                          // timing experiment: no MMAs, only the barrier traffic
          if (lane == 0) {
            mbar_arrive(&bars->empty[stage]);
            if (last) mbar_arrive(&bars->tfull[acc]);
          }
          __syncwarp();
          r1 = r2 = 0;
        } else if (!p.vw_rows) {
          // The X tiles of the item's taps lie side by side (one atom stride apart, like the 64-channel atoms of a wide
          // tile) and so do their accumulators: ONE MMA of N = nt * BNc columns per 16 pixels serves all taps of the
          // group -- a third of the MMA instructions for the 64-channel 3x3 layers (N = 192), half for 128 channels
          const uint32_t flags = ((!last || more_items) ? kPoll1 : 0u) | ((last && more_items) ? kPoll2 : 0u) | kCommit1 |
                                 (last ? kCommit2 : 0u);
          mma4_fused(tmem_d, adesc, b_const + b_lo, 128ull, instr_desc(kTileM, nt * p.BNc, 1, 1), (uint32_t)(b - b0), flags,
                     bar_full + 8u * (uint32_t)nstage, nphase, bar_tempty + 8u * (uint32_t)nacc, nacc_parity,
                     bar_empty + 8u * (uint32_t)stage, bar_tfull + 8u * (uint32_t)acc, r1, r2, b_kstep);
        } else
        for (int j = 0; j < nt; ++j) {
          uint32_t flags = 0, q1, q2;
          if (j == 0) flags |= ((!last || more_items) ? kPoll1 : 0u) | ((last && more_items) ? kPoll2 : 0u);
          if (j == nt - 1) flags |= kCommit1 | (last ? kCommit2 : 0u);
          mma4_fused(tmem_d + (uint32_t)(j * p.BNc), adesc, b_const + b_lo, 128ull,
                     idesc, (uint32_t)(b - b0), flags, bar_full + 8u * (uint32_t)nstage, nphase,
                     bar_tempty + 8u * (uint32_t)nacc, nacc_parity, bar_empty + 8u * (uint32_t)stage,
                     bar_tfull + 8u * (uint32_t)acc, q1, q2, b_kstep);
          b_lo += b_slice;
          if (j == 0) { r1 = q1; r2 = q2; }
        }
        full_ready = r1;
        acc_ready = r2;
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    const int q = warp & 3;
    int local = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++local) {
      int sp, tg, ct, kt;
      decode(item, sp, tg, ct, kt);
      const int tap0 = tg * p.T, nt = min(p.T, taps - tap0);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&bars->tfull[acc], acc_phase);
      tc_fence_after();
      const int k = kt * 128 + q * 32 + lane;
      for (int j = 0; j < nt; ++j) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_cols + j * p.BNc);
        float* drow = p.dw + ((long long)k * taps + tap0 + j) * p.C + ct * p.BNc;
        for (int c0 = 0; c0 < p.BNc; c0 += 32) {
          uint32_t v[32];
          tmem_ld16(taddr + c0, v);
          tmem_ld16(taddr + c0 + 16, v + 16);
          tmem_ld_wait();
          if (k < p.K && !(p.debug & 1)) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)       // 16-byte vector reductions (C % 8 == 0 keeps groups whole)
              if (ct * p.BNc + c0 + jj < p.C)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + jj),
                             "f"(__uint_as_float(v[jj])), "f"(__uint_as_float(v[jj + 1])),
                             "f"(__uint_as_float(v[jj + 2])), "f"(__uint_as_float(v[jj + 3]))
                             : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------ filter transform for dgrad
// Wt[c][t'][k] = W[k][tapmap[t']][c]   (bf16), 32x32 shared-memory transpose per output tap
struct TapMap {
  int n;
  int src[49];
};
__global__ void tap_transpose_kernel(const bf16* __restrict__ w, bf16* __restrict__ wt, int K, int C, int taps_src,
                                     TapMap tm) {
  __shared__ bf16 tile[32][33];
  pdl_trigger();
  pdl_wait();
  const int tdst = blockIdx.z, tsrc = tm.src[tdst];
  const int k0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int k = k0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && c < C) ? w[((long long)k * taps_src + tsrc) * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, k = k0 + threadIdx.x;
    if (k < K && c < C) wt[((long long)c * tm.n + tdst) * K + k] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ stem: im2col + padded filter
// col[pix][(r*S+s)*C + c] = x[n, oh*st-p+r*dil, ow*st-p+s*dil, c] (* mask) ; columns >= R*S*C are zero.
// One CTA per output row: the R input rows it needs are staged in shared memory as bf16 (coalesced
// reads, zero borders, mask applied) together with a column -> patch-offset table, then every thread
// assembles 16-byte column chunks from shared memory (no divisions in the inner loop) and writes them
// coalesced.  The kernel is bound by the write of the matrix.
template <int C>
__global__ void __launch_bounds__(256)
im2col_kernel(const bf16* __restrict__ x, const float* __restrict__ mask, bf16* __restrict__ col, int N, int H, int W,
              int R, int S, int stride, int pad, int dil, int Ho, int Wo, int Kpad) {
  extern __shared__ __align__(16) uint8_t im_smem[];
  const int Wp = W + 2 * pad, rowlen = Wp * C;
  bf16* rows_sm = reinterpret_cast<bf16*>(im_smem);                 // [R][rowlen]
  int* off_sm = reinterpret_cast<int*>(im_smem + (((size_t)R * rowlen * 2 + 15) & ~(size_t)15));   // [Kpad]
  const int oh = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int RSC = R * S * C, SC = S * C;
  for (int cidx = threadIdx.x; cidx < Kpad; cidx += blockDim.x) {
    int o = -1;                                                      // -1: zero column (padding of the matrix)
    if (cidx < RSC) {
      const int r = cidx / SC, rem = cidx - r * SC, s2 = rem / C, c = rem - s2 * C;
      o = r * rowlen + s2 * dil * C + c;
    }
    off_sm[cidx] = o;
  }
  for (int r = 0; r < R; ++r) {
    const int ih = oh * stride - pad + r * dil;
    const bool row_ok = ih >= 0 && ih < H;
    const long long rowbase = ((long long)n * H + ih) * W;
    for (int j = threadIdx.x; j < rowlen; j += blockDim.x) {
      const int pw = j / C, c = j - pw * C, iw = pw - pad;           // C is a compile-time constant
      float v = 0.f;
      if (row_ok && iw >= 0 && iw < W) {
        v = __bfloat162float(x[(rowbase + iw) * C + c]);
        if (mask) v *= mask[rowbase + iw];
      }
      rows_sm[r * rowlen + j] = __float2bfloat16_rn(v);
    }
  }
  __syncthreads();
  const int chunks = Kpad >> 3;
  const long long pix0 = ((long long)n * Ho + oh) * Wo;
  const unsigned short* rs16 = reinterpret_cast<const unsigned short*>(rows_sm);
  for (int i = threadIdx.x; i < Wo * chunks; i += blockDim.x) {
    const int ow = i / chunks, ch = i - ow * chunks;
    const int base = ow * stride * C;
    uint32_t w4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o0 = off_sm[ch * 8 + 2 * j], o1 = off_sm[ch * 8 + 2 * j + 1];
      const uint32_t lo = o0 >= 0 ? rs16[o0 + base] : 0u, hi = o1 >= 0 ? rs16[o1 + base] : 0u;
      w4[j] = lo | (hi << 16);
    }
    *reinterpret_cast<uint4*>(col + (pix0 + ow) * Kpad + ch * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

// ---- 7x7 stride-2 stems: TMA does the im2col -------------------------------------------------------------
// The input is copied once into a zero-bordered 4-channel image xp[N][H+6][Wp][4] (8 bytes per pixel, 3 rows of border
// above / below, 4 pixels left, >= 12 right).  For output pixel (oh, ow) and filter ROW r the seven taps (s, c) are the
// contiguous pixels 2ow-3 .. 2ow+3 of padded row 2oh + r, i.e. bytes [16 ow + 8, 16 ow + 64) of that row: a tensor map
// whose second dimension is the OUTPUT column with a 16-byte stride and whose innermost dimension is a 64-element
// (128-byte, 16-pixel) window hands the tensor core an im2col row per output pixel without the matrix ever existing
// (round 1 wrote N*Ho*Wo*152 bf16 = 319 MB to HBM and read it back, twice per step).  The layer then IS a convolution
// with R = 7, S = 1, C = 64 for the generic kernels: window element j = 4 p + c holds pixel 2ow - 4 + p, channel c, so
// the filter is laid out as Wk[k][r][j] = W[k][r][s = p - 1][c] (zero for p = 0, p > 7, c >= C).
__global__ void __launch_bounds__(256) vw_pad_kernel(const bf16* __restrict__ x, const float* __restrict__ mask,
                                                     bf16* __restrict__ xp, int H, int W, int C, int Hp, int Wp) {
  // one block = one padded image row (blockIdx.x = n * Hp + hp): no divisions per pixel
  const int hp = blockIdx.x % Hp, n = blockIdx.x / Hp;
  const int h = hp - 3;
  const bool row_ok = h >= 0 && h < H;
  const long long src_row = ((long long)n * H + h) * W;
  bf16* dst = xp + (long long)blockIdx.x * Wp * 4;
  for (int wp = threadIdx.x; wp < Wp; wp += blockDim.x) {
    const int w = wp - 4;
    uint32_t lo = 0, hi = 0;
    if (row_ok && w >= 0 && w < W) {
      const long long ip = src_row + w;
      const float m = mask ? mask[ip] : 1.f;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < C; ++c) v[c] = __bfloat162float(x[ip * C + c]) * m;
      lo = f32x2_to_bf16x2(v[0], v[1]);
      hi = f32x2_to_bf16x2(v[2], v[3]);
    }
    *reinterpret_cast<uint2*>(dst + (long long)wp * 4) = make_uint2(lo, hi);
  }
}
// Wk[k][r][4 p + c] = W[k][r][p - 1][c]   (win = 64 or 32 window elements per filter row)
__global__ void vw_filter_kernel(const bf16* __restrict__ w, bf16* __restrict__ wk, int K, int C, int win) {
  const int total = K * 7 * win;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % win, r = (i / win) % 7, k = i / (7 * win);
    const int pp = j >> 2, c = j & 3, s2 = pp - 1;
    bf16 v = __float2bfloat16(0.f);
    if (s2 >= 0 && s2 < 7 && c < C) v = w[((long long)k * 49 + r * 7 + s2) * C + c];
    wk[i] = v;
  }
}
// dw[k][r][s][c] += dWk[k][r][4 (s + 1) + c]
__global__ void vw_unfilter_add_kernel(const float* __restrict__ dwk, float* __restrict__ dw, int K, int C, int win) {
  const int total = K * 49 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, rs = (i / C) % 49, k = i / (C * 49);
    const int r = rs / 7, s2 = rs - r * 7;
    dw[i] += dwk[((long long)k * 7 + r) * win + 4 * (s2 + 1) + c];
  }
}

__global__ void pad_filter_kernel(const bf16* __restrict__ w, bf16* __restrict__ wp, int K, int RSC, int Kpad) {
  const int total = K * Kpad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / Kpad, c = i - k * Kpad;
    wp[i] = c < RSC ? w[(long long)k * RSC + c] : __float2bfloat16(0.f);
  }
}
__global__ void unpad_add_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int K, int RSC, int Kpad) {
  const int total = K * RSC;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / RSC, c = i - k * RSC;
    dw[i] += dwp[(long long)k * Kpad + c];
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// 4-D map over an NHWC bf16 tensor: dims (C, W, H, N), box (64, bw*es, bh*es, bni), element strides es
int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bni, int es) {
  EncodeTiledFn fn = encode_fn();
  B2_REQUIRE(fn != nullptr, B2_E_LAUNCH, "conv_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), (cuuint32_t)bni};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2_REQUIRE(r == CUDA_SUCCESS, B2_E_LAUNCH,
             "conv_tc: cuTensorMapEncodeTiled(activation N=%d H=%d W=%d C=%d box %dx%dx%d es=%d) failed: %d", N, H, W,
             C, bni, bh, bw, es, (int)r);
  return B2_OK;
}

// 4-D window map over the zero-bordered stem input xp[N][Hp][Wp][4]: dims (64-element window, output column, padded
// row, image), strides (16 B per output column, one padded row, one padded image); box (64, bw, 2 bh, bni) with element
// stride 2 along the rows (vertical stride of the stem)
// (win = 64: 16-pixel windows, 128-byte swizzle; win = 32: 8-pixel windows -- all a 7-tap row needs -- 64-byte swizzle)
int make_vw_map(CUtensorMap* m, const void* base, int N, int Hp, int Wp, int Wo, int bw, int bh, int bni, int win = 64) {
  EncodeTiledFn fn = encode_fn();
  B2_REQUIRE(fn != nullptr, B2_E_LAUNCH, "conv_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)win, (cuuint64_t)Wo, (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t strides[3] = {16, (cuuint64_t)Wp * 8, (cuuint64_t)Hp * Wp * 8};
  cuuint32_t box[4] = {(cuuint32_t)win, (cuuint32_t)bw, (cuuint32_t)(bh * 2), (cuuint32_t)bni};
  cuuint32_t estr[4] = {1, 1, 2, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, win == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2_REQUIRE(r == CUDA_SUCCESS, B2_E_LAUNCH,
             "conv_tc: cuTensorMapEncodeTiled(stem window map N=%d Hp=%d Wp=%d Wo=%d box %dx%dx%d) failed: %d", N, Hp, Wp,
             Wo, bni, bh, bw, (int)r);
  return B2_OK;
}

// Row-stem map over the same zero-bordered input: plain (unswizzled) rows in 256-byte units -- dims (128 elements,
// units per padded row, padded row, image), box (128, units, 7, 1) = the seven input rows of one output-row tile (fprop:
// 128 pixels, 9 units; wgrad: 64 pixels, 5 units).  Wp is a multiple of 32 pixels (vw_wp), so the units tile a row
// exactly and the part of a box behind the end of its row is zero-filled
int make_vw_rows_map(CUtensorMap* m, const void* base, int N, int Hp, int Wp, uint32_t units = kVwRowChunks) {
  EncodeTiledFn fn = encode_fn();
  B2_REQUIRE(fn != nullptr, B2_E_LAUNCH, "conv_tc: cuTensorMapEncodeTiled is not available from the driver");
  B2_REQUIRE(Wp % 32 == 0, B2_E_BADARG, "conv_tc: the stem row map needs a padded width that is a multiple of 32");
  cuuint64_t dims[4] = {128, (cuuint64_t)(Wp / 32), (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t strides[3] = {256, (cuuint64_t)Wp * 8, (cuuint64_t)Hp * Wp * 8};
  cuuint32_t box[4] = {128, units, 7, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2_REQUIRE(r == CUDA_SUCCESS, B2_E_LAUNCH, "conv_tc: cuTensorMapEncodeTiled(stem row map N=%d Hp=%d Wp=%d) failed: %d", N,
             Hp, Wp, (int)r);
  return B2_OK;
}

// 2-D map over a [rows, cols] bf16 matrix, box (64, box_rows)
int make_mat_map(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows, int box_cols = 64) {
  EncodeTiledFn fn = encode_fn();
  B2_REQUIRE(fn != nullptr, B2_E_LAUNCH, "conv_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2_REQUIRE(r == CUDA_SUCCESS, B2_E_LAUNCH, "conv_tc: cuTensorMapEncodeTiled(matrix %lldx%lld box %d) failed: %d", rows,
             cols, box_rows, (int)r);
  return B2_OK;
}

// brick of <= limit output pixels maximising tile occupancy
void choose_brick(int N, int Ho, int Wo, int limit, int need_multiple, int* bw_o, int* bh_o, int* bni_o) {
  double best = -1;
  int bbw = 1, bbh = 1, bbn = 1;
  for (int bw = 1; bw <= Wo && bw <= limit; ++bw) {
    int bh = limit / bw;
    if (bh > Ho) bh = Ho;
    if (bh < 1) continue;
    int bni = 1;
    if (bw == Wo && bh == Ho) {
      bni = limit / (bw * bh);
      if (bni > N) bni = N;
      if (bni < 1) bni = 1;
    }
    long long tiles = (long long)((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh) * ((N + bni - 1) / bni);
    double eff = (double)N * Ho * Wo / ((double)tiles * limit);
    (void)need_multiple;
    if (eff > best + 1e-9) { best = eff; bbw = bw; bbh = bh; bbn = bni; }
  }
  *bw_o = bbw; *bh_o = bbh; *bni_o = bbn;
}

int smem_limit() {
  static int lim = 0;
  if (lim == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (lim <= 0) lim = 227 * 1024;
  }
  return lim;
}

int pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

// Generic launch of the fprop-style kernel: act [N,H,W,C] -> out [N,out_H,out_W,K]
struct RunArgs {
  const void* act; int N, H, W, C;
  const void* filt; int K, R, S, stride, pad, dil, Ho, Wo;    // filt: [K][R*S*C] bf16
  void* out; int out_H, out_W, out_stride_sp, out_off_h, out_off_w;
  int pad_w, use_pad_w;                                        // pad_w is read only when use_pad_w != 0
  int vw_hp, vw_wp;                                            // != 0: `act` is a window-map stem input (make_vw_map)
  int b_mn;                                                    // filt is the UNtransposed filter [C][R*S*K] (dgrad, see FpropParams)
  int k32;                                                     // 32-element k-blocks, 64-byte swizzle (window-map stems)
  int vw_rows;                                                 // row stems (FpropParams::vw_rows); filt as for k32
  const void* act2; int split_c;                               // B2_CONV_X_CONCAT fprop: channels >= split_c of `act` live in act2
  void* out2; int split_k;                                     // B2_CONV_X_CONCAT dgrad: output channels >= split_k go to out2
  int accumulate;                                              // dgrad: reduce-add into `out`
  float* bn_sums; bool* stats_fused; int bn_totals;           // optional fused BatchNorm statistics
  int scale_mode; const float* mask_in; const float* row_scale; const float* bias;
  float* mask_out; float* ratio_out;
  int mask_R, mask_S, mask_stride, mask_pad, mask_dil, mask_H, mask_W;
};

// output-channel tile: up to 256 wide (measured: narrower tiles only add per-tile overhead, also for the 1x1 layers)
int choose_bn(int K) {
  static const int env_cap = getenv("B2POSE_TC_BN_CAP") ? atoi(getenv("B2POSE_TC_BN_CAP")) : 0;
  const int bn_cap = env_cap > 0 ? env_cap : 256;
  const int nk = (K + bn_cap - 1) / bn_cap;
  return (((K + nk - 1) / nk) + 15) / 16 * 16;
}

int run_conv_tc(const RunArgs& a, cudaStream_t st) {
  FpropParams p;
  p.N = a.N; p.H = a.H; p.W = a.W; p.C = a.C; p.K = a.K; p.R = a.R; p.S = a.S;
  p.stride = a.stride; p.pad = a.pad; p.dil = a.dil; p.Ho = a.Ho; p.Wo = a.Wo;
  choose_brick(a.N, a.Ho, a.Wo, kTileM, 1, &p.BW, &p.BH, &p.BNI);
  p.vw_rows = a.vw_rows;
  if (a.vw_rows) { p.BW = kTileM; p.BH = 1; p.BNI = 1; }          // 128 pixels of one output row
  p.tiles_w = (a.Wo + p.BW - 1) / p.BW; p.tiles_h = (a.Ho + p.BH - 1) / p.BH; p.tiles_n = (a.N + p.BNI - 1) / p.BNI;
  p.BN = choose_bn(a.K);
  p.tiles_k = (a.K + p.BN - 1) / p.BN;
  const int kbe = (a.k32 || a.vw_rows) ? 32 : kBlockK;
  p.k32 = a.k32;
  p.cblocks = (a.C + kbe - 1) / kbe;               // a ragged last block is zero-filled by TMA on the A side
  p.kblocks = a.R * a.S * p.cblocks;
  const int a_stage_bytes = a.vw_rows ? (int)kABytes : kTileM * kbe * 2;
  static_assert(7 * kVwRowPitch <= kABytes, "the seven raw stem rows must fit one ring stage");
  int stage_bytes = a_stage_bytes + p.BN * kbe * 2;
  static const bool no_tma_store = getenv("B2POSE_TC_NO_TMA_STORE") != nullptr;      // tuning switches
  static const bool no_fused_stats = getenv("B2POSE_TC_NO_FUSED_STATS") != nullptr;
  p.tma_store = (p.BN % 64 == 0 && a.out_stride_sp == 1 && !a.bias && !no_tma_store) ? 1 : 0;   // bias: direct-store path
  static const int env_debug = getenv("B2POSE_TC_DEBUG") ? atoi(getenv("B2POSE_TC_DEBUG")) : 0;
  p.debug = env_debug;
  p.accumulate = a.accumulate;
  B2_REQUIRE(!a.accumulate || p.tma_store, B2_E_UNSUPPORTED,
             "conv_tc: accumulate needs the TMA-store epilogue (output channels %% 64 == 0, stride 1)");
  // fused statistics for the wide-spatial layers (K <= 256); deeper layers are small and keep the separate pass
  // fused statistics: the per-thread accumulators live across all tiles of a CTA, so every tile of a CTA must cover
  // the same channels -- one channel tile, or a grid that is a multiple of tiles_k (below); the deep K > 256 layers
  // ran a separate pass over their (largest) outputs in round 1
  const long long total_tiles_h = (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_k;
  const bool multi_ok = p.tiles_k > 1 && a.K % p.BN == 0 && b2_num_sms() >= 2 * p.tiles_k && total_tiles_h >= b2_num_sms();
  p.bn_sums = (p.tma_store && a.bn_sums && (p.tiles_k == 1 ? a.K <= 256 : multi_ok) && !no_fused_stats) ? a.bn_sums : nullptr;
  p.bn_totals = a.bn_totals;
  if (a.stats_fused) *a.stats_fused = p.bn_sums != nullptr;
  p.b_mn = a.b_mn;
  B2_REQUIRE(!p.b_mn || p.BN % 64 == 0, B2_E_UNSUPPORTED, "conv_tc: MN-major filter tiles need 64-channel groups");
  p.n_staging = kStaging;
  int extra = (p.bn_sums ? 64 * p.BN : 0);
  // Filter resident in shared memory when one channel tile covers K and the whole filter is small (layer1-type layers,
  // the stems): the ring then carries activation tiles only -- a third to two thirds less L2 -> SM traffic per tile and
  // a deeper ring, which is what bounds these layers (B2POSE_TC_B_RESIDENT=0 disables)
  static const bool allow_resident = !(getenv("B2POSE_TC_B_RESIDENT") && atoi(getenv("B2POSE_TC_B_RESIDENT")) == 0);
  const int filt_bytes = p.kblocks * p.BN * kbe * 2;
  p.b_resident = (allow_resident && p.tiles_k == 1 && filt_bytes <= 96 * 1024 &&
                  (long long)p.tiles_w * p.tiles_h * p.tiles_n >= 2LL * b2_num_sms()) ? 1 : 0;
  if (a.vw_rows) {
    B2_REQUIRE(p.tiles_k == 1 && filt_bytes <= 96 * 1024 && a.R == 7 && a.S == 1 && a.C == 32 && !a.b_mn, B2_E_UNSUPPORTED,
               "conv_tc: the row-stem mode needs a resident window filter [K <= 256][7][32]");
    p.b_resident = 1;
  }
  if (p.b_resident) {
    stage_bytes = a_stage_bytes;
    extra += filt_bytes;
    if (p.BN <= 64) p.n_staging = 2;           // two sets of one buffer: room for two more activation stages
  }
  // narrow tiles: two four-warp epilogue groups on alternate tiles (needs one staging set per group)
  static const bool allow_split = !(getenv("B2POSE_TC_EPI_SPLIT") && atoi(getenv("B2POSE_TC_EPI_SPLIT")) == 0);
  p.epi_split = (allow_split && p.tma_store && p.BN <= 128 && p.n_staging / (p.BN / 64) >= 2 &&
                 (p.n_staging / (p.BN / 64)) % 2 == 0) ? 1 : 0;
  extra += p.tma_store ? p.n_staging * (int)kABytes : 0;
  int stages = (smem_limit() - 2048 - (int)sizeof(PipeBars) - extra) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  B2_REQUIRE(stages >= 2, B2_E_UNSUPPORTED, "conv_tc: not enough shared memory for two stages");
  p.stages = stages;
  p.tmem_cols = pow2_cols(2 * p.BN);
  p.scale_mode = a.scale_mode; p.mask_in = a.mask_in; p.row_scale = a.row_scale; p.bias = a.bias;
  p.mask_out = a.mask_out; p.ratio_out = a.ratio_out; p.out = (bf16*)a.out;
  p.mask_R = a.mask_R; p.mask_S = a.mask_S; p.mask_stride = a.mask_stride; p.mask_pad = a.mask_pad;
  p.mask_dil = a.mask_dil; p.mask_H = a.mask_H; p.mask_W = a.mask_W;
  p.out_stride_sp = a.out_stride_sp; p.out_H = a.out_H; p.out_W = a.out_W;
  p.out_off_h = a.out_off_h; p.out_off_w = a.out_off_w;
  p.pad_w = a.use_pad_w ? a.pad_w : a.pad;
  p.stride_w = a.vw_hp ? 1 : a.stride;
  CUtensorMap ma, mb, mo, ma2, mo2;
  p.split_cb = a.act2 ? a.split_c / kbe : 0;
  p.split_k = a.out2 ? a.split_k : 0;
  B2_REQUIRE(!a.act2 || (!a.vw_hp && a.split_c % kbe == 0 && a.split_c > 0 && a.split_c < a.C), B2_E_UNSUPPORTED,
             "conv_tc: a concatenated input must split at a multiple of %d channels", kbe);
  B2_REQUIRE(!a.out2 || (p.tma_store && a.split_k % 64 == 0 && a.split_k > 0 && a.split_k < a.K), B2_E_UNSUPPORTED,
             "conv_tc: a concatenated output needs the TMA-store epilogue and a split at a multiple of 64 channels");
  B2_REQUIRE(!a.k32 || (a.vw_hp && !a.b_mn), B2_E_UNSUPPORTED, "conv_tc: 32-element k-blocks are a window-map stem mode");
  int rc = a.vw_rows ? make_vw_rows_map(&ma, a.act, a.N, a.vw_hp, a.vw_wp)
           : a.vw_hp ? make_vw_map(&ma, a.act, a.N, a.vw_hp, a.vw_wp, a.Wo, p.BW, p.BH, p.BNI, kbe)
                     : make_act_map(&ma, a.act, a.N, a.H, a.W, a.act2 ? a.split_c : a.C, p.BW, p.BH, p.BNI, a.stride);
  if (rc) return rc;
  ma2 = ma;
  if (a.act2) {
    rc = make_act_map(&ma2, a.act2, a.N, a.H, a.W, a.C - a.split_c, p.BW, p.BH, p.BNI, a.stride);
    if (rc) return rc;
  }
  rc = p.b_mn ? make_mat_map(&mb, a.filt, a.C, (long long)a.R * a.S * a.K, 64)
              : make_mat_map(&mb, a.filt, a.K, (long long)a.R * a.S * a.C, p.BN, kbe);
  if (rc) return rc;
  if (p.tma_store) {
    rc = make_act_map(&mo, a.out, a.N, a.out_H, a.out_W, a.out2 ? a.split_k : a.K, p.BW, p.BH, p.BNI, 1);
    if (rc) return rc;
  } else {
    mo = ma;
  }
  mo2 = mo;
  if (a.out2) {
    rc = make_act_map(&mo2, a.out2, a.N, a.out_H, a.out_W, a.K - a.split_k, p.BW, p.BH, p.BNI, 1);
    if (rc) return rc;
  }
  const size_t smem = (size_t)stages * stage_bytes + sizeof(PipeBars) + 1024 + 16 + extra;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit());
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit());
    B2_REQUIRE(e == cudaSuccess, B2_E_LAUNCH, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = smem_limit();
  }
  const long long total = (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_k;
  int grid = (int)(total < b2_num_sms() ? total : b2_num_sms());
  if (p.bn_sums && p.tiles_k > 1) grid = grid / p.tiles_k * p.tiles_k;      // fixed channel tile per CTA
  cudaError_t le = p.epi_split
      ? launch_pdl(conv_tc_kernel<true>, dim3(grid), dim3(kConvThreads), smem, st, ma, mb, mo, ma2, mo2, p)
      : launch_pdl(conv_tc_kernel<false>, dim3(grid), dim3(kConvThreads), smem, st, ma, mb, mo, ma2, mo2, p);
  B2_REQUIRE(le == cudaSuccess, B2_E_LAUNCH, "conv_tc_kernel: launch failed: %s", cudaGetErrorString(le));
  B2_LAUNCH_CHECK("conv_tc_kernel");
  return B2_OK;
}

bool tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2POSE_DISABLE_TC");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

// ------------------------------------------------------------------ dispatch interface
namespace {

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
inline bool is_stem(const B2ConvDesc* d) { return d->C <= 4 && d->R * d->S * d->C >= 32; }
// the networks' stems (7x7, stride 2, pad 3): window-map route, see vw_pad_kernel.  B2POSE_STEM_IM2COL=1 selects round
// 1's materialised im2col matrix + GEMM (kept for the other thin-input geometries) for A/B measurements.
inline bool is_vw_stem(const B2ConvDesc* d) {
  static const int im2col = getenv("B2POSE_STEM_IM2COL") ? atoi(getenv("B2POSE_STEM_IM2COL")) : 0;
  return !im2col && is_stem(d) && d->R == 7 && d->S == 7 && d->stride == 2 && d->pad == 3 && d->dil == 1;
}
// fprop reads 8-pixel windows (32 elements, 64-byte swizzle): half the operand traffic and MMAs of the 16-pixel
// window; B2POSE_STEM_WIN64=1 selects the 128-byte-swizzle variant (the wgrad kernel uses it)
inline bool vw_fprop_k32() {
  static const bool win64 = getenv("B2POSE_STEM_WIN64") && atoi(getenv("B2POSE_STEM_WIN64")) != 0;
  return !win64;
}
inline int vw_hp(const B2ConvDesc* d) { return d->H + 6; }
// >= W + 16 and a multiple of 32 pixels: a padded row is a whole number of the 256-byte units the row maps count in, so
// whatever a box reads past the end of a row is zero-filled by TMA (never the next row, never memory behind the buffer)
inline int vw_wp(const B2ConvDesc* d) { return (d->W + 16 + 31) / 32 * 32; }
inline size_t vw_view_bytes(const B2ConvDesc* d) { return (size_t)d->N * vw_hp(d) * vw_wp(d) * 8; }
// fprop reads whole raw rows and windows them with overlapping no-swizzle descriptors (FpropParams::vw_rows);
// B2POSE_STEM_ROWS=0 selects the window tensor map (TMA gathers one 64-byte window per output pixel)
// wgrad of the stems the same way (WgradParams::vw_rows); B2POSE_STEM_WGRAD_ROWS=0 -> window tensor map, =2 -> the
// other assignment of the two descriptor strides (bring-up switch)
inline int vw_wgrad_rows() {
  static const int v = getenv("B2POSE_STEM_WGRAD_ROWS") ? atoi(getenv("B2POSE_STEM_WGRAD_ROWS")) : 1;
  return v;
}
inline bool vw_fprop_rows() {
  static const bool off = getenv("B2POSE_STEM_ROWS") && atoi(getenv("B2POSE_STEM_ROWS")) == 0;
  return !off;
}
inline int stem_kpad(const B2ConvDesc* d) { return (d->R * d->S * d->C + 7) / 8 * 8; }

int launch_vw_pad(const B2ConvDesc* d, const void* x, const float* mask, bf16* xp, cudaStream_t st) {
  vw_pad_kernel<<<d->N * vw_hp(d), 256, 0, st>>>((const bf16*)x, mask, xp, d->H, d->W, d->C, vw_hp(d), vw_wp(d));
  B2_LAUNCH_CHECK("vw_pad");
  return B2_OK;
}

int launch_im2col(const B2ConvDesc* d, const void* x, const float* mask, bf16* col, cudaStream_t st) {
  const int kpad = stem_kpad(d);
  const size_t sh = (((size_t)d->R * (d->W + 2 * d->pad) * d->C * 2 + 15) & ~(size_t)15) + (size_t)kpad * sizeof(int);
  B2_REQUIRE(sh <= 200 * 1024, B2_E_UNSUPPORTED, "im2col: input row too wide for shared memory");
#define IM2COL(CC)                                                                                              \
  do {                                                                                                          \
    if (sh > 48 * 1024) cudaFuncSetAttribute(im2col_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); \
    im2col_kernel<CC><<<d->N * d->Ho, 256, sh, st>>>((const bf16*)x, mask, col, d->N, d->H, d->W, d->R, d->S,    \
                                                    d->stride, d->pad, d->dil, d->Ho, d->Wo, kpad);            \
  } while (0)
  if (d->C == 1) IM2COL(1); else if (d->C == 2) IM2COL(2); else if (d->C == 3) IM2COL(3); else IM2COL(4);
#undef IM2COL
  B2_LAUNCH_CHECK("im2col");
  return B2_OK;
}

}  // namespace

static bool dgrad_direct(const B2ConvDesc* d);
static int wgrad_bnc(int C) { return C % 256 == 0 ? 256 : (C % 128 == 0 ? 128 : 64); }

bool conv_tc_supported(const B2ConvDesc* d, int op) {
  if (!tc_enabled() || d->dtype != B2_BF16) return false;
  if (d->K % 8 != 0) return false;
  const bool partial = d->flags & B2_CONV_PARTIAL, premasked = d->flags & B2_CONV_X_PREMASKED;
  if (d->flags & B2_CONV_X_CONCAT) {
    // x (dx) = two tensors of C / 2 channels each: plain stride-1 layers whose halves are whole channel blocks
    if (partial || is_stem(d) || d->stride != 1 || d->C % 128 != 0 || (d->flags & B2_CONV_DX_ACCUMULATE)) return false;
    if (op == 1) return dgrad_direct(d) && (d->C / 2) % 64 == 0;
    if (op == 2) return (d->C / 2) % wgrad_bnc(d->C) == 0;
    return op == 0;
  }
  if (is_stem(d)) return op == 0 || op == 2;                // im2col + GEMM; network inputs need no dgrad
  if (d->C % 8 != 0) return false;
  const bool one = (d->R == 1 && d->S == 1);
  if (partial && !one && !premasked && op != 1) return false;   // fprop / wgrad need x*mask in the loader (callers pre-mask);
                                                                 // dgrad only scales its OUTPUT rows by the mask
  if (partial && one && d->stride != 1) return false;
  if (op == 0) return true;
  if (op == 1 && (d->flags & B2_CONV_DX_ACCUMULATE)) return d->stride == 1 && d->C % 64 == 0 && d->C <= 256 * 8;
  if (op == 1) return d->stride == 1 || d->dil == 1;        // strided: one launch per output parity class
  return op == 2;
}

size_t conv_tc_workspace_bytes(const B2ConvDesc* d, int op) {
  size_t ws = 0;
  const bool partial = d->flags & B2_CONV_PARTIAL;
  const size_t dy_bytes = align256((size_t)d->N * d->Ho * d->Wo * d->K * 2);
  if (is_vw_stem(d)) {
    ws += align256(vw_view_bytes(d));                                     // zero-bordered 4-channel input
    ws += align256((size_t)d->K * 7 * 64 * (op == 2 ? 4 : 2));            // window-layout filter / its gradient
    if (op == 0 && partial) ws += align256((size_t)d->N * d->Ho * d->Wo * 4);   // renormalisation ratio
    if (op == 2 && partial && !(d->flags & B2_CONV_DY_PRESCALED)) ws += dy_bytes;
    return ws;
  }
  if (is_stem(d)) {
    const size_t kpad = stem_kpad(d);
    ws += align256((size_t)d->N * d->Ho * d->Wo * kpad * 2);            // im2col matrix
    ws += align256((size_t)d->K * kpad * (op == 2 ? 4 : 2));            // padded filter / padded dw
    if (op == 0 && partial) ws += align256((size_t)d->N * d->Ho * d->Wo * 4);   // renormalisation ratio
    if (op == 2 && partial && !(d->flags & B2_CONV_DY_PRESCALED)) ws += dy_bytes;
    return ws;
  }
  if (op == 1) {
    ws += align256((size_t)d->K * d->R * d->S * d->C * 2);               // transposed (per-class) filters
    if (partial && !(d->flags & B2_CONV_DY_PRESCALED)) ws += dy_bytes;
  }
  if (op == 2 && partial && !(d->flags & B2_CONV_DY_PRESCALED)) ws += dy_bytes;
  return ws;
}

int conv_tc_fprop(const B2ConvDesc* d, const void* x, const float* mask_in, const void* w, const float* bias,
                  void* y, float* mask_out, float* ratio_out, float* bn_sums, void* workspace, cudaStream_t st) {
  RunArgs a{};
  const bool partial = d->flags & B2_CONV_PARTIAL, premasked = d->flags & B2_CONV_X_PREMASKED;
  const float* stem_ratio = nullptr;
  a.act = x; a.N = d->N; a.H = d->H; a.W = d->W; a.C = d->C;
  if (d->flags & B2_CONV_X_CONCAT) {               // x = {first half, second half} of the channel concatenation
    const void* const* xs = (const void* const*)x;
    B2_REQUIRE(xs[0] && xs[1], B2_E_BADARG, "conv_tc_fprop: B2_CONV_X_CONCAT needs two input tensors");
    a.act = xs[0]; a.act2 = xs[1]; a.split_c = d->C / 2;
  }
  a.filt = w; a.K = d->K; a.R = d->R; a.S = d->S; a.stride = d->stride; a.pad = d->pad; a.dil = d->dil;
  if (is_vw_stem(d)) {
    bf16* xp = (bf16*)workspace;
    bf16* wk = (bf16*)((uint8_t*)workspace + align256(vw_view_bytes(d)));
    // B2_CONV_X_PREMASKED: the caller's x is zero wherever the mask is (network stems: veil = depth != 0)
    int rc = launch_vw_pad(d, x, (partial && !premasked) ? mask_in : nullptr, xp, st);
    if (rc) return rc;
    const bool rows = vw_fprop_rows() && d->K <= 256;
    const int win = (rows || vw_fprop_k32()) ? 32 : 64;
    vw_filter_kernel<<<(d->K * 7 * win + 255) / 256, 256, 0, st>>>((const bf16*)w, wk, d->K, d->C, win);
    B2_LAUNCH_CHECK("vw_filter");
    a.act = xp; a.vw_hp = vw_hp(d); a.vw_wp = vw_wp(d); a.k32 = !rows && win == 32; a.vw_rows = rows;
    a.H = vw_hp(d); a.W = d->Wo; a.C = win; a.filt = wk; a.R = 7; a.S = 1; a.stride = 2; a.pad = 0; a.dil = 1;
    a.pad_w = 0; a.use_pad_w = 1;
    if (partial) {
      // a 7x7 window is 49 mask loads per output pixel: do the mask algebra once in its own small
      // kernel and let the convolution epilogue read the ratio as a per-row scale
      float* rbuf = ratio_out ? ratio_out : (float*)((uint8_t*)wk + align256((size_t)d->K * 7 * 64 * 2));
      rc = b2_pconv_mask_update(d, mask_in, mask_out, rbuf, (void*)st);
      if (rc) return rc;
      stem_ratio = rbuf;
    }
  } else if (is_stem(d)) {
    const int kpad = stem_kpad(d);
    bf16* col = (bf16*)workspace;
    bf16* wp = (bf16*)((uint8_t*)workspace + align256((size_t)d->N * d->Ho * d->Wo * kpad * 2));
    int rc = launch_im2col(d, x, (partial && !premasked) ? mask_in : nullptr, col, st);
    if (rc) return rc;
    pad_filter_kernel<<<(d->K * kpad + 255) / 256, 256, 0, st>>>((const bf16*)w, wp, d->K, d->R * d->S * d->C, kpad);
    B2_LAUNCH_CHECK("pad_filter");
    a.act = col; a.H = d->Ho; a.W = d->Wo; a.C = kpad; a.filt = wp; a.R = 1; a.S = 1; a.stride = 1; a.pad = 0; a.dil = 1;
    if (partial) {
      // a 7x7 window is 49 mask loads per output pixel: do the mask algebra once in its own small
      // kernel and let the GEMM epilogue read the ratio as a per-row scale
      float* rbuf = ratio_out ? ratio_out
                              : (float*)((uint8_t*)wp + align256((size_t)d->K * kpad * 2));
      rc = b2_pconv_mask_update(d, mask_in, mask_out, rbuf, (void*)st);
      if (rc) return rc;
      stem_ratio = rbuf;
    }
  }
  a.Ho = d->Ho; a.Wo = d->Wo; a.out = y; a.out_H = d->Ho; a.out_W = d->Wo; a.out_stride_sp = 1;
  a.scale_mode = partial ? 1 : 0;
  a.mask_in = mask_in; a.bias = bias; a.mask_out = mask_out; a.ratio_out = ratio_out;
  if (stem_ratio) {
    a.scale_mode = 2; a.row_scale = stem_ratio; a.mask_out = nullptr; a.ratio_out = nullptr;
  }
  a.mask_R = d->R; a.mask_S = d->S; a.mask_stride = d->stride; a.mask_pad = d->pad; a.mask_dil = d->dil;
  a.mask_H = d->H; a.mask_W = d->W;
  bool fused = false;
  a.bn_sums = bn_sums; a.stats_fused = &fused; a.bn_totals = (d->flags & B2_CONV_BN_TOTALS) ? 1 : 0;
  int rc = run_conv_tc(a, st);
  if (rc) return rc;
  if (bn_sums && !fused) {
    if (d->flags & B2_CONV_BN_TOTALS)
      return b2_bn_stats_totals(y, (int64_t)d->N * d->Ho * d->Wo, d->K, d->dtype, bn_sums, (void*)st);
    return b2_bn_stats(y, (int64_t)d->N * d->Ho * d->Wo, d->K, d->dtype, bn_sums, (void*)st);
  }
  return B2_OK;
}

// stride-1 dgrad straight from the untransposed filter (MN-major B tiles of 64 output channels)
static bool dgrad_direct(const B2ConvDesc* d) {
  static const bool off = getenv("B2POSE_DGRAD_TRANSPOSE") && atoi(getenv("B2POSE_DGRAD_TRANSPOSE")) != 0;
  return !off && d->stride == 1 && choose_bn(d->C) % 64 == 0 && !(d->flags & B2_CONV_W_PREPARED);
}
bool conv_tc_dgrad_needs_filter(const B2ConvDesc* d) { return !dgrad_direct(d); }

// mode 0: transpose the filter into the workspace, then run; 1: ONLY the filter transposes, written to `workspace`
// (b2_pconv_dgrad_filter: the weights are constant during a step, so callers hoist this off the backward critical
// path); 2: `w` already is the buffer mode 1 produced (B2_CONV_W_PREPARED).
static int conv_tc_dgrad_impl(const B2ConvDesc* d, const void* dy, const float* ratio, const void* w,
                              const float* mask_in, void* dx, void* workspace, cudaStream_t st, int mode) {
  const bool partial = d->flags & B2_CONV_PARTIAL, premasked = d->flags & B2_CONV_X_PREMASKED;
  uint8_t* ws = (uint8_t*)workspace;
  bf16* wt = mode == 2 ? (bf16*)const_cast<void*>(w) : (bf16*)ws;
  ws += align256((size_t)d->K * d->R * d->S * d->C * 2);
  const void* dys = dy;
  if (mode != 1 && partial && !(d->flags & B2_CONV_DY_PRESCALED) && ratio) {
    int rc = b2_scale_rows(dy, ratio, ws, (int64_t)d->N * d->Ho * d->Wo, d->K, B2_BF16, (void*)st);
    if (rc) return rc;
    dys = ws;
  }
  RunArgs a{};
  a.act = dys; a.N = d->N; a.H = d->Ho; a.W = d->Wo; a.C = d->K;
  a.K = d->C; a.out = dx; a.out_H = d->H; a.out_W = d->W;
  a.scale_mode = (partial && !premasked && mask_in) ? 2 : 0;
  a.row_scale = mask_in;
  a.stride = 1;
  const int taps = d->R * d->S;
  dim3 tb(32, 8);
  if (d->stride == 1 && dgrad_direct(d)) {
    // dx = conv(dy, flipped/transposed filter), pad' = dil*(R-1) - pad -- the flip is a tap index and the transpose an
    // MN-major B operand: the kernel reads W[k][tap][c] as it is (round 1 ran 86 transpose kernels per step)
    if (mode == 1) return B2_OK;
    a.filt = w; a.b_mn = 1;
    a.R = d->R; a.S = d->S; a.dil = d->dil; a.pad = d->dil * (d->R - 1) - d->pad;
    a.Ho = d->H; a.Wo = d->W; a.out_stride_sp = 1;
    a.accumulate = (d->flags & B2_CONV_DX_ACCUMULATE) ? 1 : 0;
    if (d->flags & B2_CONV_X_CONCAT) {             // dx = {gradient of the first half, of the second half}
      void* const* dxs = (void* const*)dx;
      B2_REQUIRE(dxs[0] && dxs[1], B2_E_BADARG, "conv_tc_dgrad: B2_CONV_X_CONCAT needs two output tensors");
      a.out = dxs[0]; a.out2 = dxs[1]; a.split_k = d->C / 2;
    }
    return run_conv_tc(a, st);
  }
  B2_REQUIRE(!(d->flags & B2_CONV_X_CONCAT), B2_E_UNSUPPORTED, "conv_tc_dgrad: B2_CONV_X_CONCAT needs the direct dgrad");
  if (d->stride == 1) {
    // dx = conv(dy, flipped/transposed filter), pad' = dil*(R-1) - pad
    TapMap tm;
    tm.n = taps;
    for (int r = 0; r < d->R; ++r)
      for (int s2 = 0; s2 < d->S; ++s2) tm.src[r * d->S + s2] = (d->R - 1 - r) * d->S + (d->S - 1 - s2);
    dim3 tg((d->C + 31) / 32, (d->K + 31) / 32, taps);
    if (mode != 2) {
      launch_pdl(tap_transpose_kernel, tg, tb, 0, st, (const bf16*)w, wt, d->K, d->C, taps, tm);
      B2_LAUNCH_CHECK("tap_transpose");
    }
    if (mode == 1) return B2_OK;
    a.filt = wt; a.R = d->R; a.S = d->S; a.dil = d->dil; a.pad = d->dil * (d->R - 1) - d->pad;
    a.Ho = d->H; a.Wo = d->W; a.out_stride_sp = 1;
    a.accumulate = (d->flags & B2_CONV_DX_ACCUMULATE) ? 1 : 0;
    return run_conv_tc(a, st);
  }
  // strided (dil == 1): output pixels of parity class (ph, pw) = (ih % stride, iw % stride) only see the taps
  // r == (ph + pad) mod stride; each class is a dense stride-1 correlation over dy written to a strided
  // sub-lattice of dx.  Classes that see no tap stay zero.
  const int sd = d->stride;
  bool need_zero = false;
  for (int ph = 0; ph < sd; ++ph) {
    int cr = 0, cs = 0;
    for (int r = 0; r < d->R; ++r) cr += (((ph + d->pad - r) % sd + sd) % sd == 0);
    for (int s2 = 0; s2 < d->S; ++s2) cs += (((ph + d->pad - s2) % sd + sd) % sd == 0);
    if (cr == 0 || cs == 0) need_zero = true;
  }
  if (need_zero && mode != 1) {
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)d->N * d->H * d->W * d->C * 2, st);
    B2_REQUIRE(e == cudaSuccess, B2_E_LAUNCH, "conv_tc_dgrad: memset failed: %s", cudaGetErrorString(e));
  }
  bf16* wcls = wt;
  for (int ph = 0; ph < sd; ++ph) {
    int rl[8], nr = 0;
    for (int r = d->R - 1; r >= 0; --r)            // decreasing r -> increasing dy offset
      if (((ph + d->pad - r) % sd + sd) % sd == 0) rl[nr++] = r;
    if (nr == 0 || ph >= d->H) continue;
    for (int pw = 0; pw < sd; ++pw) {
      int sl[8], ns = 0;
      for (int s2 = d->S - 1; s2 >= 0; --s2)
        if (((pw + d->pad - s2) % sd + sd) % sd == 0) sl[ns++] = s2;
      if (ns == 0 || pw >= d->W) continue;
      TapMap tm;
      tm.n = nr * ns;
      for (int i = 0; i < nr; ++i)
        for (int j = 0; j < ns; ++j) tm.src[i * ns + j] = rl[i] * d->S + sl[j];
      dim3 tg((d->C + 31) / 32, (d->K + 31) / 32, tm.n);
      if (mode != 2) {
        launch_pdl(tap_transpose_kernel, tg, tb, 0, st, (const bf16*)w, wcls, d->K, d->C, taps, tm);
        B2_LAUNCH_CHECK("tap_transpose");
      }
      if (mode == 1) { wcls += (size_t)tm.n * d->K * d->C; continue; }
      // floor division of (ph + pad - r_max) by the stride (exact by construction)
      auto fdiv = [](int a_, int b_) { return (a_ >= 0) ? a_ / b_ : -((-a_ + b_ - 1) / b_); };
      const int o0h = fdiv(ph + d->pad - rl[0], sd), o0w = fdiv(pw + d->pad - sl[0], sd);
      a.filt = wcls; a.R = nr; a.S = ns; a.dil = 1; a.pad = 0;
      RunArgs b = a;
      b.Ho = (d->H - ph + sd - 1) / sd; b.Wo = (d->W - pw + sd - 1) / sd;
      b.out_stride_sp = sd; b.out_off_h = ph; b.out_off_w = pw;
      // the kernel addresses dy at (a - pad' + t'), so pad' = -o0 (may differ per dim: use the larger
      // and shift the tap origin through separate pads)
      b.pad = -o0h;
      b.pad_w = -o0w;
      b.use_pad_w = 1;
      int rc = run_conv_tc(b, st);
      if (rc) return rc;
      wcls += (size_t)tm.n * d->K * d->C;
    }
  }
  return B2_OK;
}

int conv_tc_dgrad(const B2ConvDesc* d, const void* dy, const float* ratio, const void* w, const float* mask_in,
                  void* dx, void* workspace, cudaStream_t st) {
  return conv_tc_dgrad_impl(d, dy, ratio, w, mask_in, dx, workspace, st, (d->flags & B2_CONV_W_PREPARED) ? 2 : 0);
}

int conv_tc_dgrad_filter(const B2ConvDesc* d, const void* w, void* wt, cudaStream_t st) {
  return conv_tc_dgrad_impl(d, nullptr, nullptr, w, nullptr, nullptr, wt, st, 1);
}

int conv_tc_wgrad(const B2ConvDesc* d, const void* x, const float* mask_in, const void* dy, const float* ratio,
                  float* dw, void* workspace, cudaStream_t st) {
  // 1x1: the mask is folded into the (pre)scaled dy rows; 3x3: x is pre-masked (see conv_tc_supported);
  // stem: the im2col kernel applies the mask.
  const bool partial = d->flags & B2_CONV_PARTIAL, premasked = d->flags & B2_CONV_X_PREMASKED;
  uint8_t* ws = (uint8_t*)workspace;
  B2ConvDesc g = *d;                 // geometry seen by the GEMM
  float* dw_out = dw;
  float* dwp = nullptr;
  int kpad = 0;
  bool vw = false;
  int vw_rows = 0;
  if (is_vw_stem(d)) {
    vw = true;
    bf16* xp = (bf16*)ws;
    ws += align256(vw_view_bytes(d));
    dwp = (float*)ws;
    ws += align256((size_t)d->K * 7 * 64 * 4);
    if (!(d->flags & B2_CONV_WS_HAS_COL)) {       // (else: the caller kept the fprop workspace, xp is in it)
      int rc = launch_vw_pad(d, x, (partial && !premasked) ? mask_in : nullptr, xp, st);
      if (rc) return rc;
    }
    cudaError_t e = cudaMemsetAsync(dwp, 0, (size_t)d->K * 7 * 64 * 4, st);
    B2_REQUIRE(e == cudaSuccess, B2_E_LAUNCH, "conv_tc_wgrad: memset failed: %s", cudaGetErrorString(e));
    x = xp;
    vw_rows = vw_wgrad_rows();
    g.H = vw_hp(d); g.W = d->Wo; g.C = vw_rows ? 32 : 64; g.R = 7; g.S = 1; g.stride = 2; g.pad = 0; g.dil = 1;
    dw_out = dwp;
  } else if (is_stem(d)) {
    kpad = stem_kpad(d);
    bf16* col = (bf16*)ws;
    ws += align256((size_t)d->N * d->Ho * d->Wo * kpad * 2);
    dwp = (float*)ws;
    ws += align256((size_t)d->K * kpad * 4);
    if (!(d->flags & B2_CONV_WS_HAS_COL)) {       // (else: the caller kept the fprop workspace, the matrix is in it)
      int rc = launch_im2col(d, x, (partial && !premasked) ? mask_in : nullptr, col, st);
      if (rc) return rc;
    }
    cudaError_t e = cudaMemsetAsync(dwp, 0, (size_t)d->K * kpad * 4, st);
    B2_REQUIRE(e == cudaSuccess, B2_E_LAUNCH, "conv_tc_wgrad: memset failed: %s", cudaGetErrorString(e));
    x = col;
    g.H = d->Ho; g.W = d->Wo; g.C = kpad; g.R = 1; g.S = 1; g.stride = 1; g.pad = 0; g.dil = 1;
    dw_out = dwp;
  }
  const void* dys = dy;
  if (partial && !(d->flags & B2_CONV_DY_PRESCALED) && ratio) {
    int rc = b2_scale_rows(dy, ratio, ws, (int64_t)d->N * d->Ho * d->Wo, d->K, B2_BF16, (void*)st);
    if (rc) return rc;
    dys = ws;
  }
  const B2ConvDesc* d0 = d;
  d = &g;
  WgradParams p;
  p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.K = d->K; p.R = d->R; p.S = d->S;
  p.stride = d->stride; p.pad = d->pad; p.dil = d->dil; p.Ho = d->Ho; p.Wo = d->Wo;
  p.stride_w = vw ? 1 : d->stride; p.pad_w = vw ? 0 : d->pad;
  // bricks of exactly 64 pixel slots over the output space (rows of 128 B each; ragged parts zero-filled)
  {
    double best = -1;
    p.BW = 1; p.BH = 1; p.BNI = 1;
    for (int bw = 1; bw <= kWgPix; bw <<= 1) {           // bw*bh*bni must equal 64 exactly
      for (int bh = 1; bw * bh <= kWgPix; bh <<= 1) {
        int bni = kWgPix / (bw * bh);
        long long tiles = (long long)((d->Wo + bw - 1) / bw) * ((d->Ho + bh - 1) / bh) * ((d->N + bni - 1) / bni);
        double eff = (double)d->N * d->Ho * d->Wo / ((double)tiles * kWgPix);
        if (eff > best + 1e-9) { best = eff; p.BW = bw; p.BH = bh; p.BNI = bni; }
      }
    }
  }
  p.vw_rows = vw_rows;
  if (vw_rows) { p.BW = kWgPix; p.BH = 1; p.BNI = 1; }          // 64 pixels of one output row
  // halo mode (WgradParams::halo): tap group = filter row, bricks of 64 pixels of one image row; B2POSE_WGRAD_HALO=0 disables
  static const int env_halo = getenv("B2POSE_WGRAD_HALO") ? atoi(getenv("B2POSE_WGRAD_HALO")) : 1;
  p.halo = (env_halo && !vw && !(d0->flags & B2_CONV_X_CONCAT) && d->R == 3 && d->S == 3 && d->stride == 1 &&
            d->dil == 1 && d->pad == 1 && d->C == 64 && d->Wo % kWgPix == 0) ? 1 : 0;
  if (p.halo) { p.BW = kWgPix; p.BH = 1; p.BNI = 1; }
  p.tiles_w = (d->Wo + p.BW - 1) / p.BW; p.tiles_h = (d->Ho + p.BH - 1) / p.BH; p.tiles_n = (d->N + p.BNI - 1) / p.BNI;
  p.BNc = vw_rows ? 32 : wgrad_bnc(d->C);
  p.ctiles = (d->C + p.BNc - 1) / p.BNc;
  p.ktiles = (d->K + 127) / 128;
  const int taps = d->R * d->S;
  p.T = 256 / p.BNc;                       // 2 accumulator sets x T x BNc <= 512 TMEM columns
  if (p.halo) p.T = 3;                     // (one filter row per work item)
  if (vw_rows) p.T = 7;                    // (row stems: all seven tap rows share the dY tile, 2 x 7 x 32 columns)
  if (p.T > taps) p.T = taps;
  if (p.T < 1) p.T = 1;
  p.tgroups = (taps + p.T - 1) / p.T;
  p.T = (taps + p.tgroups - 1) / p.tgroups;   // balance the groups (9 taps: 3+3+3, not 4+4+1)
  const long long bricks = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
  const long long base_items = (long long)p.tgroups * p.ctiles * p.ktiles;
  // enough pixel splits to fill the machine ~2x, but at least 8 bricks (512 pixels) per item
  // pixel splits: every split ends in 128 x T x BNc fp32 reductions into dW, which cost as much as ~30 pipeline
  // stages, and the kernel runs on the side stream beside the BatchNorm / dgrad chain.  Measured in the step
  // (B2POSE_WGRAD_WAVES = 1 / 2 / 3 / 12): 14.74 / 14.85 / 15.1 / 14.65 ms -- one wave of work items for the 1x1 layers,
  // two for the multi-tap layers (whose few (tap group, tile) items need the splits for parallelism) = 12, the default
  // End of round 2 (faster issue loops, fewer other atomics): 1 / 2-for-multi-tap / 3 waves = 13.35 / 13.48 / 13.92 ms --
  // one wave for every layer (half the reduction traffic of the multi-tap layers: ncu-free timing with
  // B2POSE_TC_DEBUG=1 shows the reduction at 30-36 % of the 128- and 256-channel 3x3 wgrads).  The split count is rounded
  // DOWN so that the items fill whole waves (99 splits x 3 tap groups = 297 items had run as three rounds on 148 SMs).
  static const int env_waves = getenv("B2POSE_WGRAD_WAVES") ? atoi(getenv("B2POSE_WGRAD_WAVES")) : 1;
  const int waves = env_waves == 12 ? (taps == 1 ? 1 : 2) : env_waves;
  long long want = ((long long)waves * b2_num_sms()) / base_items;
  long long max_splits = (bricks + 7) / 8;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  p.bricks_per_split = (int)((bricks + want - 1) / want);
  p.splits = (int)((bricks + p.bricks_per_split - 1) / p.bricks_per_split);
  const int stage_bytes = 2 * kWgPix * 128 + (vw_rows ? (int)kVwgRowsBytes : p.halo ? (int)kHaloBytes
                                                                             : p.T * (p.BNc / 64) * kWgPix * 128);
  // B2POSE_WGRAD_SMEM_KB caps the ring so that blocks of other kernels (the BatchNorm streams running on the
  // main stream while wgrad runs on the side stream) can share the SM
  static const int env_kb = getenv("B2POSE_WGRAD_SMEM_KB") ? atoi(getenv("B2POSE_WGRAD_SMEM_KB")) : 0;
  int budget = smem_limit();
  if (env_kb > 0 && env_kb * 1024 < budget) budget = env_kb * 1024;
  int stages = (budget - 2048 - (int)sizeof(PipeBars)) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.tmem_cols = pow2_cols(2 * p.T * p.BNc);
  p.dw = dw_out;
  static const int env_debug_w = getenv("B2POSE_TC_DEBUG") ? atoi(getenv("B2POSE_TC_DEBUG")) : 0;
  p.debug = env_debug_w;
  CUtensorMap mdy, mx, mx2;
  int rc = make_act_map(&mdy, dys, d->N, d->Ho, d->Wo, d->K, p.BW, p.BH, p.BNI, 1);
  if (rc) return rc;
  p.split_c = 0;
  if (d0->flags & B2_CONV_X_CONCAT) {              // x = {first half, second half} of the channel concatenation
    const void* const* xs = (const void* const*)x;
    B2_REQUIRE(xs[0] && xs[1] && !vw && (d->C / 2) % p.BNc == 0, B2_E_UNSUPPORTED,
               "conv_tc_wgrad: B2_CONV_X_CONCAT needs two inputs of a whole number of channel tiles");
    p.split_c = d->C / 2;
    rc = make_act_map(&mx, xs[0], d->N, d->H, d->W, d->C / 2, p.BW, p.BH, p.BNI, d->stride);
    if (rc) return rc;
    rc = make_act_map(&mx2, xs[1], d->N, d->H, d->W, d->C / 2, p.BW, p.BH, p.BNI, d->stride);
    if (rc) return rc;
  } else {
    rc = vw_rows ? make_vw_rows_map(&mx, x, d->N, vw_hp(d0), vw_wp(d0), kVwgRowUnits)
         : vw    ? make_vw_map(&mx, x, d->N, vw_hp(d0), vw_wp(d0), d->Wo, p.BW, p.BH, p.BNI)
         : p.halo ? make_act_map(&mx, x, d->N, d->H, d->W, d->C, (int)kHaloPix, 1, 1, 1)
                 : make_act_map(&mx, x, d->N, d->H, d->W, d->C, p.BW, p.BH, p.BNI, d->stride);
    if (rc) return rc;
    mx2 = mx;
  }
  const size_t smem = (size_t)stages * stage_bytes + sizeof(PipeBars) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit());
    B2_REQUIRE(e == cudaSuccess, B2_E_LAUNCH, "wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long total = base_items * p.splits;
  int grid = (int)(total < b2_num_sms() ? total : b2_num_sms());
  cudaError_t le = launch_pdl(wgrad_tc_kernel, dim3(grid), dim3(kThreads), smem, st, mdy, mx, mx2, p);
  B2_REQUIRE(le == cudaSuccess, B2_E_LAUNCH, "wgrad_tc_kernel: launch failed: %s", cudaGetErrorString(le));
  B2_LAUNCH_CHECK("wgrad_tc_kernel");
  if (dwp && vw) {
    vw_unfilter_add_kernel<<<(d0->K * 49 * d0->C + 255) / 256, 256, 0, st>>>(dwp, dw, d0->K, d0->C, vw_rows ? 32 : 64);
    B2_LAUNCH_CHECK("vw_unfilter_add");
  } else if (dwp) {
    const int rsc = d0->R * d0->S * d0->C;
    unpad_add_kernel<<<(d0->K * rsc + 255) / 256, 256, 0, st>>>(dwp, dw, d0->K, rsc, kpad);
    B2_LAUNCH_CHECK("unpad_add");
  }
  return B2_OK;
}
