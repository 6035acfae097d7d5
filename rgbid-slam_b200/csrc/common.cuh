// common.cuh -- shared device helpers: batched image views, projection, samplers, reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/rgbid_b200.h"
#include "se3.cuh"

namespace rgbid {

constexpr int kNumSMsB200 = 148;

// ------------------------------------------------------------------------------------------------
// Batched pitched image: stream b lives at base + b * sstride (bytes).  This is the reference's
// PtrStepSz<float> (ThirdParty/pcl_gpu_containers/include/kernel_containers.h:54-102) plus a batch
// stride so that one launch covers all independent frame pairs.
// ------------------------------------------------------------------------------------------------
struct ImgB {
  float* p;
  size_t pitch;    // bytes
  size_t sstride;  // bytes between streams
  int rows, cols;
  __host__ __device__ __forceinline__ float* row(int b, int y) const
  {
    return (float*)((char*)p + (size_t)b * sstride + (size_t)y * pitch);
  }
};

inline ImgB make_img(const float* p, size_t pitch, int rows, int cols, size_t sstride = 0)
{
  ImgB i;
  i.p = const_cast<float*>(p); i.pitch = pitch; i.sstride = sstride; i.rows = rows; i.cols = cols;
  return i;
}

// Pixel-space rigid transform Rp = K R K^-1, tp = K t (what the reference passes as Mat33 / float3).
struct Proj {
  float r[9];
  float t[3];
};

__device__ __forceinline__ float qnanf() { return __int_as_float(0x7fffffff); }

// registerPixel (src/cuda/warping_registration.cu:129-146): back-project pixel (x, y) with inverse
// depth w, move it with P, re-project.  Returns the inverse depth in the target camera.
__device__ __forceinline__ float project_pixel(const Proj& P, int x, int y, float w, float& xs, float& ys)
{
  float z = 1.f / w;
  float X = __int2float_rn(x) * z, Y = __int2float_rn(y) * z;
  float Xc = (P.r[0] * X + P.r[1] * Y + P.r[2] * z) + P.t[0];
  float Yc = (P.r[3] * X + P.r[4] * Y + P.r[5] * z) + P.t[1];
  float Zc = (P.r[6] * X + P.r[7] * Y + P.r[8] * z) + P.t[2];
  float wc = 1.f / Zc;
  xs = Xc * wc;
  ys = Yc * wc;
  return wc;
}

// In-image test of the warp kernels (warping_registration.cu:490-491, 528-529) on the +0.5-shifted
// coordinates; true if the sample may be fetched.
__device__ __forceinline__ bool in_image(float xt, float yt, int cols, int rows)
{
  int fx = __float2int_rd(xt), fy = __float2int_rd(yt);
  return !(fx < 0 || fy < 0 || fx >= cols || fy >= rows);
}

// Point-filtered fetch at unnormalised texture coordinate (xt, yt) (cudaFilterModePoint,
// warping_registration.cu:994, fetch :531): texel floor(xt), floor(yt).  Caller guarantees in_image.
__device__ __forceinline__ float sample_nearest(const float* __restrict__ base, size_t pitch, float xt, float yt)
{
  int ix = __float2int_rd(xt), iy = __float2int_rd(yt);
  return __ldg((const float*)((const char*)base + (size_t)iy * pitch) + ix);
}

// Bilinear fetch reproducing cudaFilterModeLinear with clamp addressing
// (warping_registration.cu:943, fetch :493): xB = xt - 0.5, i = floor(xB), alpha = frac(xB) in 1.8 fixed point.
// The texture unit was characterised on B200 with tests/cuda/tex_probe2.cu: ka = round(alpha * 256),
// kb = round(beta * 256), and the four weights are themselves 8-bit:
//   w11 = (ka*kb + 128) >> 8, w10 = ka - w11, w01 = kb - w11, w00 = 256 - ka - kb + w11   (all / 256)
// (zero mismatches over all 257 x 257 weight pairs).
__device__ __forceinline__ float sample_bilinear_q8(const float* __restrict__ base, size_t pitch, int cols,
                                                    int rows, float xt, float yt)
{
  float xB = xt - 0.5f, yB = yt - 0.5f;
  float fxf = floorf(xB), fyf = floorf(yB);
  int ka = __float2int_rd((xB - fxf) * 256.f + 0.5f);
  int kb = __float2int_rd((yB - fyf) * 256.f + 0.5f);
  int w11 = (ka * kb + 128) >> 8;
  int w10 = ka - w11, w01 = kb - w11, w00 = 256 - ka - kb + w11;
  int i0 = (int)fxf, j0 = (int)fyf;
  int i1 = min(max(i0 + 1, 0), cols - 1), j1 = min(max(j0 + 1, 0), rows - 1);
  i0 = min(max(i0, 0), cols - 1);
  j0 = min(max(j0, 0), rows - 1);
  const float* r0 = (const float*)((const char*)base + (size_t)j0 * pitch);
  const float* r1 = (const float*)((const char*)base + (size_t)j1 * pitch);
  float t00 = __ldg(r0 + i0), t10 = __ldg(r0 + i1), t01 = __ldg(r1 + i0), t11 = __ldg(r1 + i1);
  // a tap with zero weight is not blended by the hardware: a NaN texel (the corner pixels of every pyramid
  // level >= 1 are NaN, pyrdown.cu:124-127) only poisons the result when its weight is non-zero
  float acc = 0.f;
  if (w00) acc += __int2float_rn(w00) * t00;
  if (w10) acc += __int2float_rn(w10) * t10;
  if (w01) acc += __int2float_rn(w01) * t01;
  if (w11) acc += __int2float_rn(w11) * t11;
  return acc * (1.f / 256.f);
}

// Current-frame maps of one stream at one level: raw pointers for the software sampler and (optionally)
// texture objects over the same memory: texW = point filter, texI = linear filter, both clamp / unnormalised,
// i.e. exactly the objects the reference creates per call (warping_registration.cu:926-947, 977-998) but
// created once per buffer.
struct CurFrame {
  const float* Wc;
  const float* Ic;
  size_t wpitch, ipitch;
  cudaTextureObject_t texW, texI;
};

// Warp of one keyframe pixel: the fused equivalent of trafo3DKernelInvDepthGridStride
// (warping_registration.cu:505-546) followed by trafo3DKernelIntensityWithInvDepthGridStride
// (:465-501).  geom_is_warped selects the tracker's behaviour (intensity is warped with the
// just-warped inverse depth as geometry, src/visodo.cpp:1121-1126) or KeyframeAlign's (keyframe inverse
// depth as geometry, src/keyframe_align.cpp:239).  TEX: gather through the texture unit (1 instruction per
// fetch, hardware bilinear) instead of the software sampler (same results, see tests).
template <bool TEX>
__device__ __forceinline__ void warp_pixel(const Proj& P, int x, int y, float w0, const CurFrame& C, int cols,
                                           int rows, bool geom_is_warped, float& w1, float& i1)
{
  w1 = qnanf();
  i1 = qnanf();
  if (isnan(w0)) return;
  float xs, ys;
  float w3 = project_pixel(P, x, y, w0, xs, ys);
  float xt = xs + 0.5f, yt = ys + 0.5f;
  bool inside = in_image(xt, yt, cols, rows);
  if (inside) {
    float w2 = TEX ? tex2D<float>(C.texW, xt, yt) : sample_nearest(C.Wc, C.wpitch, xt, yt);
    float tz = P.t[2];
    float v1z = (1.f / w3 - tz) * w0;
    float res = (v1z / (1.f - w2 * tz)) * w2;
    if (res > 0.f) w1 = res;
  }
  if (geom_is_warped) {
    if (isnan(w1)) return;
    project_pixel(P, x, y, w1, xs, ys);
    xt = xs + 0.5f; yt = ys + 0.5f;
    inside = in_image(xt, yt, cols, rows);
  }
  if (inside) {
    float r = TEX ? tex2D<float>(C.texI, xt, yt) : sample_bilinear_q8(C.Ic, C.ipitch, cols, rows, xt, yt);
    i1 = fmaxf(0.f, fminf(r, 255.f));
  }
}

// Branch-free two-stage form of warp_pixel for the texture path: clamp addressing makes every fetch safe, so
// a thread can issue the fetches of all its pixels back to back (memory-level parallelism) and discard the
// invalid ones afterwards.  Stage 1 -> coordinates of the inverse-depth fetch; stage 2 (after the fetch) ->
// warped inverse depth and coordinates of the intensity fetch; stage 3 -> warped intensity.
struct WarpCoord {
  float xt, yt, w3;
  bool inside;
};

__device__ __forceinline__ WarpCoord warp_stage1(const Proj& P, int x, int y, float w0, int cols, int rows)
{
  WarpCoord c;
  float xs, ys;
  c.w3 = project_pixel(P, x, y, w0, xs, ys);
  c.xt = xs + 0.5f; c.yt = ys + 0.5f;
  c.inside = !isnan(w0) && in_image(c.xt, c.yt, cols, rows);
  return c;
}

__device__ __forceinline__ float warp_stage2(const Proj& P, int x, int y, float w0, float w2, WarpCoord& c, int cols,
                                             int rows, bool geom_is_warped)
{
  float tz = P.t[2];
  float v1z = (1.f / c.w3 - tz) * w0;
  float res = (v1z / (1.f - w2 * tz)) * w2;
  const bool ok = c.inside && (res > 0.f);
  const float w1 = ok ? res : qnanf();
  if (geom_is_warped) {  // uniform
    float xs, ys;
    project_pixel(P, x, y, w1, xs, ys);
    c.xt = xs + 0.5f; c.yt = ys + 0.5f;
    c.inside = ok && in_image(c.xt, c.yt, cols, rows);
  }
  return w1;
}

__device__ __forceinline__ float warp_stage3(float r, const WarpCoord& c)
{
  return c.inside ? fmaxf(0.f, fminf(r, 255.f)) : qnanf();
}

// ------------------------------------------------------------------------------------------------
// Reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Robust weights (computeWeight, src/cuda/estimate_VO.cu:141-167; thresholds src/internal.h:74-76)
__device__ __forceinline__ float mest_weight(float e, int mest)
{
  float w = 1.f;
  if (mest == RGBID_HUBER) {
    if (fabsf(e) > 1.345f) w = 1.345f / fabsf(e);
  } else if (mest == RGBID_TUKEY) {
    if (fabsf(e) < 4.685f) {
      float a = (e / 4.685f) * (e / 4.685f);
      w = (1.f - a) * (1.f - a);
    } else w = 0.f;
  } else if (mest == RGBID_STUDENT) {
    w = (5.f + 1.f) / (5.f + e * e);
  }
  return w;
}

}  // namespace rgbid
