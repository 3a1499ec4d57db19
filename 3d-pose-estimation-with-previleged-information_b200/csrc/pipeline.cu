// On-device input pipeline of the depth stream -- the per-sample CPU work of
// depth_datasets.Dataset.parse_sample (depth_datasets.py:153-217) as two batched kernels:
//
//   colour : homography crop (cameralib.reproject_image_fast, cameralib.py:667-711: cv2.remap, INTER_LINEAR,
//            constant 0 border) of a uint8 HWC frame, fused with ToTensor + Normalize (depth_datasets.py:92-94)
//            -> fp32 CHW
//   depth  : the same crop of a float frame, fused with utils.to_depth (utils.py:68-75, optional) and
//            enhance_ntu / enhance_pku (depth_datasets.py:39-56) -> fp32 [N,1,S,S]
//
// cv2.remap semantics reproduced here: source coordinates are rounded to 1/32 pixel
// (cvRound(x * INTER_TAB_SIZE)), the four bilinear weights are the products of the 1-D weights
// {1 - f/32, f/32}; uint8 images use 15-bit fixed-point weights and (sum + 2^14) >> 15, float images
// S00*w00 + S01*w01 + S10*w10 + S11*w11 in float; neighbours outside the frame read the border value 0.
// The kernels are gather-bound: one thread per output pixel, destination writes coalesced.
#include "b2_common.cuh"

namespace {

struct SrcCoord {
  int ix, iy, fx, fy;
};

// dst pixel (u, v) -> source coordinate through the float32 homography (cameralib.py:690-692)
__device__ __forceinline__ SrcCoord src_coord(const float* __restrict__ h, int u, int v) {
  const float fu = (float)u, fv = (float)v;
  const float xs = h[0] * fu + h[1] * fv + h[2];
  const float ys = h[3] * fu + h[4] * fv + h[5];
  const float ws = h[6] * fu + h[7] * fv + h[8];
  const float x = __fdiv_rn(xs, ws), y = __fdiv_rn(ys, ws);
  // cvRound(x * 32): round half to even; saturate like the int conversion of a huge / NaN coordinate
  const float x32 = fminf(fmaxf(x * 32.f, -1.0e9f), 1.0e9f), y32 = fminf(fmaxf(y * 32.f, -1.0e9f), 1.0e9f);
  const int sx = (x32 == x32) ? __float2int_rn(x32) : 0x40000000, sy = (y32 == y32) ? __float2int_rn(y32) : 0x40000000;
  SrcCoord c;
  c.ix = sx >> 5; c.iy = sy >> 5; c.fx = sx & 31; c.fy = sy & 31;
  return c;
}

__global__ void __launch_bounds__(256)
remap_rgb_kernel(const uint8_t* __restrict__ src, int Hs, int Ws, const float* __restrict__ hom, int So,
                 float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ dst) {
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= So * So) return;
  const int v = p / So, u = p - v * So;
  const SrcCoord c = src_coord(hom + n * 9, u, v);
  const uint8_t* img = src + (long long)n * Hs * Ws * 3;
  int out[3] = {0, 0, 0};
  if (!(c.ix >= Ws || c.ix + 1 < 0 || c.iy >= Hs || c.iy + 1 < 0)) {
    const int w00 = (32 - c.fx) * (32 - c.fy) * 32, w01 = c.fx * (32 - c.fy) * 32;
    const int w10 = (32 - c.fx) * c.fy * 32, w11 = c.fx * c.fy * 32;
    const bool x0 = c.ix >= 0 && c.ix < Ws, x1 = c.ix + 1 >= 0 && c.ix + 1 < Ws;
    const bool y0 = c.iy >= 0 && c.iy < Hs, y1 = c.iy + 1 >= 0 && c.iy + 1 < Hs;
    const uint8_t* r0 = img + ((long long)c.iy * Ws + c.ix) * 3;
    const uint8_t* r1 = r0 + (long long)Ws * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int p00 = (x0 && y0) ? r0[k] : 0, p01 = (x1 && y0) ? r0[3 + k] : 0;
      const int p10 = (x0 && y1) ? r1[k] : 0, p11 = (x1 && y1) ? r1[3 + k] : 0;
      out[k] = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;
    }
  }
  // ToTensor (/255) + Normalize ((x - mean) / std), fp32 like torchvision
  float* d = dst + (long long)n * 3 * So * So + p;
  d[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)out[0], 255.f), m0), s0);
  d[(long long)So * So] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)out[1], 255.f), m1), s1);
  d[2LL * So * So] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)out[2], 255.f), m2), s2);
}

// enhance_ntu / enhance_pku (depth_datasets.py:39-56) on one value
__device__ __forceinline__ float enhance(float v, float thresh, int nexponent) {
  const float x = __fdiv_rn(v, 0.039215688f);             // image / (10.0 / 255.0), float32 like numpy
  if (!nexponent) return __fdiv_rn(x, 3.0f);
  return thresh <= x ? expf(-x) : 0.f;
}

__global__ void __launch_bounds__(256)
remap_depth_kernel(const float* __restrict__ src, int Hs, int Ws, const float* __restrict__ hom, int So,
                   const float* __restrict__ cam, float thresh, int nexponent, int do_enhance,
                   float* __restrict__ dst) {
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= So * So) return;
  const int v = p / So, u = p - v * So;
  float val;
  if (hom) {
    const SrcCoord c = src_coord(hom + n * 9, u, v);
    const float* img = src + (long long)n * Hs * Ws;
    val = 0.f;
    if (!(c.ix >= Ws || c.ix + 1 < 0 || c.iy >= Hs || c.iy + 1 < 0)) {
      const float ax = (float)c.fx * (1.f / 32.f), ay = (float)c.fy * (1.f / 32.f);
      const float w00 = (1.f - ay) * (1.f - ax), w01 = (1.f - ay) * ax, w10 = ay * (1.f - ax), w11 = ay * ax;
      const bool x0 = c.ix >= 0 && c.ix < Ws, x1 = c.ix + 1 >= 0 && c.ix + 1 < Ws;
      const bool y0 = c.iy >= 0 && c.iy < Hs, y1 = c.iy + 1 >= 0 && c.iy + 1 < Hs;
      const float* r0 = img + (long long)c.iy * Ws + c.ix;
      const float* r1 = r0 + Ws;
      const float p00 = (x0 && y0) ? r0[0] : 0.f, p01 = (x1 && y0) ? r0[1] : 0.f;
      const float p10 = (x0 && y1) ? r1[0] : 0.f, p11 = (x1 && y1) ? r1[1] : 0.f;
      val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p00, w00), __fmul_rn(p01, w01)), __fmul_rn(p10, w10)),
                      __fmul_rn(p11, w11));
    }
  } else {                                    // already cropped: src is [N, So, So]
    val = src[(long long)n * So * So + p];
  }
  if (cam) {                                  // utils.to_depth with this sample's camera: kinv (2x2) | c
    const float* k = cam + n * 6;
    const float du = (float)u - k[4], dv = (float)v - k[5];
    const float xn = __fadd_rn(__fmul_rn(du, k[0]), __fmul_rn(dv, k[1]));        // same arithmetic as unproject_kernel
    const float yn = __fadd_rn(__fmul_rn(du, k[2]), __fmul_rn(dv, k[3]));
    const float ss = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xn, xn), __fmul_rn(yn, yn)), 1.f), 1.f);
    val = __fdiv_rn(val, __fsqrt_rn(ss));
  }
  dst[(long long)n * So * So + p] = do_enhance ? enhance(val, thresh, nexponent) : val;
}

// back_project.projectPoints (back_project.py:12-36): x = K * distort((R X + t) / z), Kd = [k1, k2, p1, p2, k3]
struct ProjP {
  float R[9], t[3], K[9], Kd[5];
};
__global__ void project_points_kernel(const float* __restrict__ X, int n, ProjP c, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float X0 = X[i], X1 = X[n + i], X2 = X[2 * n + i];                   // X is 3 x N like the reference
  float x = c.R[0] * X0 + c.R[1] * X1 + c.R[2] * X2 + c.t[0];
  float y = c.R[3] * X0 + c.R[4] * X1 + c.R[5] * X2 + c.t[1];
  const float z = c.R[6] * X0 + c.R[7] * X1 + c.R[8] * X2 + c.t[2];
  x = x / z;
  y = y / z;
  const float r = x * x + y * y;
  const float radial = 1.f + c.Kd[0] * r + c.Kd[1] * r * r + c.Kd[4] * r * r * r;
  const float xd = x * radial + 2.f * c.Kd[2] * x * y + c.Kd[3] * (r + 2.f * x * x);
  // the reference overwrites row 0 before computing row 1 (back_project.py:29-30): the cross term uses the DISTORTED x
  const float yd = y * radial + 2.f * c.Kd[3] * xd * y + c.Kd[2] * (r + 2.f * y * y);
  out[i] = c.K[0] * xd + c.K[1] * yd + c.K[2];
  out[n + i] = c.K[3] * xd + c.K[4] * yd + c.K[5];
  out[2 * n + i] = z;
}

}  // namespace

extern "C" int b2_project_points(const float* X, int32_t n, const float* R9, const float* t3, const float* K9,
                                 const float* Kd5, float* out, void* stream) {
  B2_REQUIRE(X && R9 && t3 && K9 && Kd5 && out && n > 0, B2_E_BADARG, "project_points: bad argument");
  ProjP c;
  for (int i = 0; i < 9; ++i) { c.R[i] = R9[i]; c.K[i] = K9[i]; }
  for (int i = 0; i < 3; ++i) c.t[i] = t3[i];
  for (int i = 0; i < 5; ++i) c.Kd[i] = Kd5[i];
  project_points_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(X, n, c, out);
  B2_LAUNCH_CHECK("project_points");
  return B2_OK;
}

extern "C" int b2_remap_normalize_rgb(const uint8_t* src, int32_t N, int32_t Hs, int32_t Ws, const float* homography,
                                      int32_t side_out, const float* mean3, const float* std3, float* dst,
                                      void* stream) {
  B2_REQUIRE(src && homography && mean3 && std3 && dst, B2_E_BADARG, "remap_normalize_rgb: null argument");
  B2_REQUIRE(N > 0 && N <= 65535 && Hs > 0 && Ws > 0 && side_out > 0, B2_E_BADARG, "remap_normalize_rgb: bad size");
  const dim3 grid((side_out * side_out + 255) / 256, N);
  remap_rgb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, Hs, Ws, homography, side_out, mean3[0], mean3[1],
                                                        mean3[2], std3[0], std3[1], std3[2], dst);
  B2_LAUNCH_CHECK("remap_normalize_rgb");
  return B2_OK;
}

extern "C" int b2_remap_enhance_depth(const float* src, int32_t N, int32_t Hs, int32_t Ws, const float* homography,
                                      int32_t side_out, const float* cam, float veil_threshold, int32_t nexponent,
                                      int32_t do_enhance, float* dst, void* stream) {
  B2_REQUIRE(src && dst, B2_E_BADARG, "remap_enhance_depth: null argument");
  B2_REQUIRE(N > 0 && N <= 65535 && side_out > 0 && (homography == nullptr || (Hs > 0 && Ws > 0)), B2_E_BADARG,
             "remap_enhance_depth: bad size");
  const dim3 grid((side_out * side_out + 255) / 256, N);
  remap_depth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, Hs, Ws, homography, side_out, cam, veil_threshold,
                                                          nexponent, do_enhance, dst);
  B2_LAUNCH_CHECK("remap_enhance_depth");
  return B2_OK;
}
