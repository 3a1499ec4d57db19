/*
 * b2pose.h -- C ABI of the B200-native depth-stream hot path.
 *
 * The reference (Hunger-Prevails/3D-Pose-Estimation-with-Previleged-Information) has no
 * FFI / plugin / operator registry: its boundary is a Python module surface built on ATen
 * calls.  Each entry point below therefore names the reference
 * Python interface whose device work it replaces (file:line under /root/reference).  The
 * Python host layer (package dir "3d-pose-estimation-with-previleged-information_b200")
 * binds these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller; the library never allocates,
 *    frees or synchronises; all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *  - activations are NHWC ("channels_last"), filters KRSC, masks [N,H,W] fp32 in {0,1}.
 *  - `dtype` selects the activation/filter element type: B2_F32 (CUDA-core FFMA path, used
 *    for the fp32 parity mode) or B2_BF16 (tcgen05 tensor-core path, fp32 accumulate).
 *  - return value: 0 on success, negative B2_E* otherwise; b2_last_error() gives a
 *    thread-local message.  There is no CPU fallback anywhere.
 */
#ifndef B2POSE_H_
#define B2POSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 5

/* slots of a BatchNorm partial-sum buffer: float[B2_BN_PARTS][2*C] (see the BatchNorm section) */
#define B2_BN_PARTS 320

enum { B2_F32 = 0, B2_BF16 = 1 };

enum {
  B2_OK = 0,
  B2_E_BADARG = -1,      /* null pointer / inconsistent shape */
  B2_E_UNSUPPORTED = -2, /* configuration this build has no kernel for */
  B2_E_LAUNCH = -3,      /* CUDA launch / runtime failure */
  B2_E_WORKSPACE = -4    /* workspace too small */
};

/* conv flags */
enum {
  B2_CONV_PARTIAL = 1,       /* mask-renormalised (PartialConv) semantics                      */
  B2_CONV_X_PREMASKED = 2,   /* caller guarantees x == x*mask_in (skip the multiply)           */
  B2_CONV_DY_PRESCALED = 4,  /* dgrad/wgrad: dy already multiplied by `ratio`                  */
  B2_CONV_FORCE_FFMA = 8,    /* debugging / fp32-accurate path even for bf16 tensors           */
  B2_CONV_DX_ACCUMULATE = 16,/* dgrad: dx += result (bf16 TMA reduce-add) -- folds the gradient of */
                             /* a residual branch into the block input's gradient; tensor-core     */
                             /* path with stride 1 and C % 64 == 0 only, B2_E_UNSUPPORTED otherwise */
  B2_CONV_BN_TOTALS = 32,    /* fprop: bn_partials is ONE pre-zeroed float[2*K] the producer ADDS   */
                             /* to (totals BatchNorm path, needs b2_bn_totals_supported(K, dtype))  */
  B2_CONV_W_PREPARED = 64,   /* dgrad: `w` is the buffer b2_pconv_dgrad_filter produced for this     */
                             /* descriptor (flipped / transposed filter), not the KRSC filter        */
  B2_CONV_WS_HAS_COL = 128,  /* wgrad of a stem layer (C <= 4): `workspace` is the very buffer the   */
                             /* fprop of the same descriptor ran with, its im2col matrix is reused    */
  B2_CONV_X_CONCAT = 256     /* the input is the channel concatenation of TWO NHWC tensors of C / 2   */
                             /* channels each (the fusion unit, fusionnet.py:137) that is never      */
                             /* materialised: `x` (fprop, wgrad) is a const void* const[2], `dx`     */
                             /* (dgrad) a void* const[2] of the two halves.  bf16 tensor-core path,  */
                             /* plain stride-1 layers with C % 128 == 0 (b2_conv_uses_tensor_cores   */
                             /* answers for the flagged descriptor); B2_E_UNSUPPORTED otherwise      */
};

typedef struct B2ConvDesc {
  int32_t N, H, W, C;        /* input  x : [N,H,W,C]                                            */
  int32_t K, R, S;           /* filter w : [K,R,S,C]                                            */
  int32_t stride, pad, dil;  /* same in both spatial dims (all reference layers are square)     */
  int32_t Ho, Wo;            /* output y : [N,Ho,Wo,K]                                          */
  int32_t dtype;             /* B2_F32 | B2_BF16 for x, w, y, dy, dx                            */
  int32_t flags;             /* B2_CONV_*                                                       */
} B2ConvDesc;

int b2_abi_version(void);
/* Leave `sms` streaming multiprocessors free in every grid launched from now on (0 = use the whole device).  For
 * callers that run other kernels beside this library's persistent one-CTA-per-SM grids -- the NCCL kernels of a
 * gradient all-reduce overlapped with backward (the reference exchanges gradients inside nn.DataParallel,
 * depth_main.py:72).  Process-wide; takes effect at the next launch (captured CUDA graphs keep the grids they were
 * captured with). */
int b2_set_sm_reserve(int32_t sms);
const char* b2_last_error(void);
/* 1 if the bf16 tensor-core (tcgen05) kernel will be used for this descriptor and op
 * (0 = fprop, 1 = dgrad, 2 = wgrad), 0 if the FFMA kernel will. */
int b2_conv_uses_tensor_cores(const B2ConvDesc* d, int op);
/* bytes of scratch the call may need (0 for most configurations). */
size_t b2_conv_workspace_bytes(const B2ConvDesc* d, int op);

/* ---- partial convolution --------------------------------------------------------------
 * replaces PartialConv.forward, partial_conv.py:32-58 (and plain nn.Conv2d.forward when
 * B2_CONV_PARTIAL is clear: partial_depthnet.py:201, fusionnet.py:136, ...).
 *   mask_out[n,oh,ow] = clamp(sum_window mask_in, 0, 1)                      (:39,:43)
 *   ratio            = (R*S)/(sum + 1e-6) * mask_out   (fp32)                (:41,:44)
 *   y = conv(x*mask_in, w) * ratio                      (no bias, :53)
 *   y = (conv(x*mask_in, w) * ratio + bias) * mask_out  (bias, :49-51)
 * mask_in/mask_out/ratio_out are fp32; ratio_out (optional) saves the per-pixel factor for
 * the backward pass.  bias is fp32 [K] or NULL.  bn_partials (optional, float[B2_BN_PARTS][2*K])
 * receives the per-channel partial sums / sums of squares of the stored y for training BatchNorm
 * (fused into the tensor-core epilogue where possible, else a b2_bn_stats pass). */
int b2_pconv_fprop(const B2ConvDesc* d, const void* x, const float* mask_in, const void* w,
                   const float* bias, void* y, float* mask_out, float* ratio_out, float* bn_partials,
                   void* workspace, size_t ws_bytes, void* stream);

/* backward of the above w.r.t. x (autograd of partial_conv.py:46-53):
 *   dx = dgrad(w, dy * ratio) * mask_in
 * ratio / mask_in may be NULL (plain convolution). */
int b2_pconv_dgrad(const B2ConvDesc* d, const void* dy, const float* ratio, const void* w,
                   const float* mask_in, void* dx, void* workspace, size_t ws_bytes, void* stream);

/* The tensor-core dgrad reads the filter flipped and transposed ([C][taps][K], per output-parity class for strided
 * layers).  The weights are constant during a training step, so the transform can be hoisted off the backward
 * critical path: b2_pconv_dgrad_filter writes it to `wt` (b2_pconv_dgrad_filter_bytes(d) bytes, 0 = this dgrad does
 * not run on the tensor-core path), and b2_pconv_dgrad with B2_CONV_W_PREPARED takes `wt` in place of `w`. */
size_t b2_pconv_dgrad_filter_bytes(const B2ConvDesc* d);
int b2_pconv_dgrad_filter(const B2ConvDesc* d, const void* w, void* wt, size_t wt_bytes, void* stream);

/* backward w.r.t. w:  dw[K,R,S,C] (fp32, ACCUMULATED into) += wgrad(x*mask_in, dy*ratio). */
int b2_pconv_wgrad(const B2ConvDesc* d, const void* x, const float* mask_in, const void* dy,
                   const float* ratio, float* dw, void* workspace, size_t ws_bytes, void* stream);

/* only the mask algebra (partial_conv.py:39-44), for callers that need the veil without a conv. */
int b2_pconv_mask_update(const B2ConvDesc* d, const float* mask_in, float* mask_out, float* ratio_out,
                         void* stream);

/* ---- row / column helpers (NHWC activations viewed as [rows, C]) ---------------------- */
/* out[r,c] = in[r,c] * scale[r]   (scale fp32) */
int b2_scale_rows(const void* in, const float* scale, void* out, int64_t rows, int32_t C, int32_t dtype,
                  void* stream);
/* sums[c] += sum_r in[r,c] * (w ? w[r] : 1)   (fp32 accumulate into sums[C]) -- bias gradients */
int b2_col_sum(const void* in, const float* row_weight, float* sums, int64_t rows, int32_t C, int32_t dtype,
               void* stream);
/* veil = (depth != 0)  : partial_depthnet.py:215, partial_fusionnet.py:255 */
int b2_veil_from_depth(const void* depth, float* veil, int64_t n, int32_t dtype, void* stream);
/* NCHW fp32 -> NHWC dtype (host batches arrive NCHW fp32: depth_train.py:386) and back */
int b2_nchw_to_nhwc(const float* in, void* out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t dtype,
                    void* stream);
int b2_nhwc_to_nchw(const void* in, float* out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t dtype,
                    void* stream);
/* dst = (dtype) src, elementwise fp32 -> bf16 (weight shadow copies) */
int b2_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* dst = (float) src, elementwise bf16 -> fp32 (gradients exchanged between ranks as bf16: the all-reduce behind
 * nn.DataParallel's gather, depth_main.py:72, moves half the bytes) */
int b2_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);

/* ---- BatchNorm2d (+ReLU, +residual, +veil) --------------------------------------------
 * replaces nn.BatchNorm2d / F.relu / residual add in the blocks, partial_depthnet.py:143-157,
 * fusionnet.py:107-127,138-140.  Training statistics are per call (per rank), momentum
 * update of running stats as torch (unbiased variance).
 *
 * Per-channel reductions are atomics-free and deterministic: a producer (b2_bn_stats, the fused
 * epilogue of b2_pconv_fprop, b2_bn_bwd_reduce) writes fp32 partial sums to its own slot of
 * `partials` = float[B2_BN_PARTS][2*C] (every slot is written, unused ones with zeros, so the
 * buffer needs no initialisation) and a finalize call combines the slots in fp64. */
/* partials[slot][0:C] = sum_r y, partials[slot][C:2C] = sum_r y^2 over the slot's rows */
int b2_bn_stats(const void* y, int64_t rows, int32_t C, int32_t dtype, float* partials, void* stream);
/* training != 0: mean / biased var from `partials` (count = rows), running stats updated;
 * training == 0: from the running stats.  Writes mean[C], invstd[C]. */
int b2_bn_finalize(const float* partials, int64_t rows, int32_t C, float* running_mean, float* running_var,
                   float momentum, float eps, int32_t training, float* mean, float* invstd, void* stream);
/* z = relu?( (y-mean)*invstd*gamma + beta (+ residual) ) (* row_mask[r]) */
int b2_bn_apply(const void* y, const float* mean, const float* invstd, const float* gamma, const float* beta,
                const void* residual, const float* row_mask, int32_t relu, void* z, int64_t rows, int32_t C,
                int32_t dtype, void* stream);
/* backward, pass 1: g = dz * (relu ? z>0 : 1) (* row_mask);  partial sums of g and g * xhat.
 * z may be NULL for a layer WITHOUT residual: the ReLU gate is then recomputed from y, gamma, beta
 * with the forward's own fp32 expression (saves reading z). */
int b2_bn_bwd_reduce(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                     const float* gamma, const float* beta, const float* row_mask, int32_t relu, float* partials,
                     int64_t rows, int32_t C, int32_t dtype, void* stream);
/* gsum[0:C] = sum g, gsum[C:2C] = sum g*xhat;  dgamma += sum g*xhat, dbeta += sum g (fp32, accumulated) */
int b2_bn_bwd_finalize(const float* partials, int32_t C, float* gsum, float* dgamma, float* dbeta, void* stream);
/* backward, pass 2: dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) (* row_scale[r]);
 * d_residual = g (optional).  training == 0 (frozen statistics): dy = gamma*invstd*g. */
int b2_bn_bwd_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                    const float* gamma, const float* beta, const float* gsum, const float* row_mask,
                    const float* row_scale, int32_t relu, int32_t training, void* dy, void* d_residual,
                    int64_t rows, int32_t C, int32_t dtype, void* stream);

/* ---- BatchNorm2d, totals path (bf16, C a power of two in [64, 2048]) --------------------
 * Same math as above with the finalize kernels folded away: producers add their per-block sums
 * into ONE pre-zeroed float[2*C] vector (`totals`: sum | sum of squares; `gsum`: sum g | sum g*xhat)
 * with fp32 reductions, and the consumers derive what they need in their prologue.  Two launches
 * per direction instead of three; the price is a run-to-run rounding difference in the last bit
 * of the statistics (fp32 atomic order).  The caller zeroes totals / gsum (one memset per step). */
int b2_bn_totals_supported(int32_t C, int32_t dtype);
int b2_bn_stats_totals(const void* y, int64_t rows, int32_t C, int32_t dtype, float* totals, void* stream);
/* training != 0: batch statistics from `totals`, running stats updated (momentum, unbiased var);
 * training == 0: running stats (totals may be NULL).  Writes z and mean[C], invstd[C].
 * gate_out (optional, uint8[rows*C/8 rounded up to a multiple of 16 bytes]): bit j of byte v = (pre-ReLU value of element 8v+j > 0) -- the ReLU
 * gate of a residual layer, 1/16 of the bytes of z; pass it as `gate` to the two backward calls (z = NULL). */
int b2_bn_apply_totals(const void* y, const float* totals, int64_t rows, float* running_mean,
                       float* running_var, float momentum, float eps, int32_t training, const float* gamma,
                       const float* beta, const void* residual, const float* row_mask, int32_t relu, void* z,
                       float* mean, float* invstd, uint8_t* gate_out, int32_t C, int32_t dtype, void* stream);
int b2_bn_bwd_reduce_totals(const void* dz, const void* z, const void* y, const float* mean,
                            const float* invstd, const float* gamma, const float* beta, const float* row_mask,
                            int32_t relu, float* gsum, const uint8_t* gate, int64_t rows, int32_t C, int32_t dtype,
                            void* stream);
/* as b2_bn_bwd_apply; additionally dgamma += gsum[C:2C], dbeta += gsum[0:C] (either may be NULL) */
int b2_bn_bwd_apply_totals(const void* dz, const void* z, const void* y, const float* mean,
                           const float* invstd, const float* gamma, const float* beta, const float* gsum,
                           const float* row_mask, const float* row_scale, int32_t relu, int32_t training,
                           void* dy, void* d_residual, float* dgamma, float* dbeta, const uint8_t* gate, int64_t rows,
                           int32_t C, int32_t dtype, void* stream);

/* ---- feature-mimic (distillation) loss -------------------------------------------------
 * replaces Trainer.distill, depth_train.py:115-129 (the "privileged information" term):
 *   mode 0: mean_n || (t - s) * a ||_2     mode 1: mean_n || (sigmoid(t) - sigmoid(s)) * a ||_2
 *   mode 2: mean_all(BCEWithLogits(s, sigmoid(t))) * sum(a) / N          (bin_dist, as written there)
 * teach / student: [N, C, H, W] logical, layout 0 = NHWC memory, 1 = NCHW memory, dtype fp32 | bf16;
 * atten: [N, H*W] fp32.  partials: float[N * B2_MIMIC_PARTS] scratch, scale: float[N] (saved for the
 * backward), loss: one float.  Backward: dstudent = dloss[0] * d(loss)/d(student)  (dloss device
 * pointer or NULL = 1). */
#define B2_MIMIC_PARTS 64
int b2_mimic_loss_fwd(const void* teach, const void* student, const float* atten, int32_t N, int32_t C,
                      int32_t HW, int32_t layout, int32_t dtype, int32_t mode, float* partials, float* scale,
                      float* loss, void* stream);
int b2_mimic_loss_bwd(const void* teach, const void* student, const float* atten, const float* scale,
                      const float* dloss, int32_t N, int32_t C, int32_t HW, int32_t layout, int32_t dtype,
                      int32_t mode, void* dstudent, void* stream);
/* utils.get_attention, utils.py:14-42: out[n, y, x] = sum_j exp(-|(x, y) - coords[n, j] * side_out / side_in|^2 / 5),
 * divided by its maximum.  image_coords: [N, J, 2] fp32 (x, y) pixels of the side_in image; out: [N, side_out^2]. */
int b2_attention_map(const float* image_coords, int32_t N, int32_t J, int32_t side_in, int32_t side_out,
                     float* out, void* stream);

/* ---- evaluation metrics: utils.analyze / statistics (utils.py:197-262) with the back-rotation of the
 * test loops (depth_train.py:522-523), accumulated over batches.  spec_cam / true_cam: [N,J,3] fp32 mm,
 * valid: [N,J] uint8, back_rotate: [N,3,3] fp32 or NULL, mirror: int32[J] or NULL (identity).
 * acc (double[10], ADDED to): valid joints, sum dist, #(dist <= rough), sum max(0, 1 - dist/rough),
 * then the error taxonomy counts solid / close / depth / jitter / switch / fail (utils.py:210-221). */
int b2_pose_metrics(const float* spec_cam, const float* true_cam, const uint8_t* valid, const float* back_rotate,
                    const int32_t* mirror, int32_t N, int32_t J, float t_solid, float t_close, float t_rough,
                    double* acc, void* stream);

/* ---- on-device input pipeline (the per-sample CPU work of depth_datasets.py:153-217) -----
 * Homography crop = cameralib.reproject_image_fast (cameralib.py:667-711): cv2.remap semantics (INTER_LINEAR,
 * coordinates rounded to 1/32 pixel, constant 0 border).  homography: DEVICE float[N][9] (row major,
 * destination pixel -> source pixel, i.e. old_K old_R inv(new_K new_R) in float32).
 * rgb  : src uint8 [N,Hs,Ws,3] -> dst fp32 [N,3,S,S] = ((remap / 255) - mean) / std   (ToTensor + Normalize,
 *        depth_datasets.py:92-94); mean3 / std3 are HOST float[3].
 * depth: src fp32 [N,Hs,Ws] (homography NULL: already cropped [N,S,S]) -> remap -> optional utils.to_depth with
 *        the sample's camera (cam: DEVICE float[N][6] = inv(K[:2,:2]) row major | K[0,2], K[1,2]; NULL = skip)
 *        -> optional enhance_ntu / enhance_pku (depth_datasets.py:39-56: x = v / (10/255); nexponent ?
 *        exp(-x) * (veil_threshold <= x) : x / 3; threshold 0.1 for NTU, 0.5 for PKU) -> dst fp32 [N,1,S,S]. */
int b2_remap_normalize_rgb(const uint8_t* src, int32_t N, int32_t Hs, int32_t Ws, const float* homography,
                           int32_t side_out, const float* mean3, const float* std3, float* dst, void* stream);
int b2_remap_enhance_depth(const float* src, int32_t N, int32_t Hs, int32_t Ws, const float* homography,
                           int32_t side_out, const float* cam, float veil_threshold, int32_t nexponent,
                           int32_t do_enhance, float* dst, void* stream);

/* back_project.projectPoints, back_project.py:12-36: pinhole projection with OpenCV radial / tangential distortion.
 * X: DEVICE float[3][n] world points; R9, t3, K9, Kd5 = [k1, k2, p1, p2, k3]: HOST arrays (row major);
 * out: DEVICE float[3][n] = (u, v, camera z). */
int b2_project_points(const float* X, int32_t n, const float* R9, const float* t3, const float* K9,
                      const float* Kd5, float* out, void* stream);

/* ---- MaxPool2d(3, stride 2, pad 1) on x and veil together: partial_depthnet.py:219-220 --- */
int b2_maxpool3x3s2_fwd(const void* x, const float* veil_in, void* y, uint8_t* argmax, float* veil_out,
                        int32_t N, int32_t H, int32_t W, int32_t C, int32_t dtype, void* stream);
int b2_maxpool3x3s2_bwd(const void* dy, const uint8_t* argmax, void* dx, int32_t N, int32_t H, int32_t W,
                        int32_t C, int32_t dtype, void* stream);

/* ---- volumetric heat-map head ----------------------------------------------------------
 * replaces utils.to_heatmap + utils.decode, utils.py:154-194, in one pass.
 * logits: channel index = d*J + j;  layout 0 = NHWC [N,H,W,D*J], 1 = NCHW [N,D*J,H,W].
 * coords [N,J,3] = (x<-W, y<-H, z<-D) * depth_range; vmax/vsum [N,J] saved for backward. */
int b2_head_fwd(const void* logits, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W, int32_t layout,
                int32_t dtype, float depth_range, float* coords, float* vmax, float* vsum, void* stream);
/* dlogit_i = p_i * (u . g_i - u . c),  u = dcoords, g_i = grid coords of voxel i (x depth_range) */
int b2_head_bwd(const void* logits, const float* dcoords, const float* coords, const float* vmax,
                const float* vsum, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W, int32_t layout,
                int32_t dtype, float depth_range, void* dlogits, void* stream);
/* materialised heat-map for API compatibility: heat [N,J,H,W,D] fp32 (utils.py:154-175);
 * coords_scratch: [N,J,3] floats of scratch. */
int b2_heatmap_softmax(const void* logits, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                       int32_t layout, int32_t dtype, float* heat, float* coords_scratch, void* stream);
/* utils.decode on a materialised heat-map (utils.py:178-194) and its backward */
int b2_heatmap_decode(const float* heat, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                      float depth_range, float* coords, void* stream);
int b2_heatmap_decode_bwd(const float* dcoords, int32_t N, int32_t J, int32_t D, int32_t H, int32_t W,
                          float depth_range, float* dheat, void* stream);
/* softmax backward on the materialised heat-map: dlogits = p*(dp - sum p*dp), written in `layout` */
int b2_heatmap_softmax_bwd(const float* heat, const float* dheat, int32_t N, int32_t J, int32_t D, int32_t H,
                           int32_t W, int32_t layout, int32_t dtype, void* dlogits, void* stream);

/* root-relative shift + masked SmoothL1/L1/MSE (mean) and its gradient: depth_train.py:397-405.
 * criterion 0 = SmoothL1(beta 1), 1 = L1, 2 = MSE.  valid: uint8 [N,J].
 * out: loss[1], spec_cam [N,J,3], dcoords [N,J,3] = d loss / d coords. */
int b2_pose_loss(const float* coords, const float* true_cam, const uint8_t* valid, int32_t N, int32_t J,
                 int32_t key_index, float loss_div, int32_t criterion, float* loss, float* spec_cam,
                 float* dcoords, void* stream);

/* ---- depth unprojection: utils.to_depth (utils.py:68-75) + Camera.image_to_camera
 * (cameralib.py:188-200, no-distortion branch).  kinv = inv(K[:2,:2]) row-major, c = K[:2,2].
 * out = img / sqrt(xn^2 + yn^2 + 2), (xn,yn) = ((u,v) - c) @ kinv^T, fp32.  n_img images. */
int b2_unproject_depth(const float* img, float* out, int32_t n_img, int32_t H, int32_t W, const float kinv[4],
                       const float c[2], void* stream);

/* ---- optimizer: clip_grad_norm_ + Adam (L2 weight decay), depth_train.py:451-456 ---------
 * sumsq (double[1], caller zeroes) += sum g^2 */
int b2_grad_sumsq(const float* g, int64_t n, double* sumsq, void* stream);
/* one fused step over flat fp32 buffers.  coef = min(1, max_norm/(sqrt(sumsq)*inv_scale + 1e-6));
 * g' = g*inv_scale*coef + wd*w; Adam(m, v) with bias correction from `step` (1-based);
 * the whole update is skipped when sumsq is not finite (the reference's inf-skip, :435-438).
 * w16 (optional) receives the bf16 shadow of the new weights.
 * dev_hyper (optional, DEVICE float[3] = {lr, 1-beta1^step, sqrt(1-beta2^step)}) overrides the
 * by-value lr/step so that a captured CUDA graph of the step can be replayed with a moving step
 * count and learning-rate schedule (depth_train.py:621-638). */
int b2_adam_step(float* w, const float* g, float* m, float* v, void* w16, int64_t n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int32_t step, const double* sumsq,
                 float max_norm, float inv_scale, const float* dev_hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2POSE_H_ */
