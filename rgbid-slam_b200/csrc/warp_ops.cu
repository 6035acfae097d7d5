// warp_ops.cu -- SE(3) inverse warps, covisibility and keyframe inverse-depth fusion.
//
// Drop-in (un-fused) versions of the reference's warp kernels plus the batched/fused variants used by
// the tracker.  The reference creates and destroys a texture object around every warp launch
// (src/cuda/warping_registration.cu:926-964); here the gathers go through the read-only path with the
// texture unit's addressing and 1/256 weight quantisation reproduced in software (common.cuh).
#include "kernels.cuh"

namespace rgbid {

namespace {

constexpr int BX = 32, BY = 8;
inline dim3 grid2d(int cols, int rows, int z) { return dim3((cols + BX - 1) / BX, (rows + BY - 1) / BY, z); }

// K4: trafo3DKernelInvDepthGridStride (warping_registration.cu:505-546)
__global__ void __launch_bounds__(BX* BY) warp_invdepth_kernel(ImgB src, ImgB prev, ImgB dst, Proj P)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  float out = qnanf();
  float w = prev.row(0, y)[x];
  if (!isnan(w)) {
    float xs, ys;
    float w3 = project_pixel(P, x, y, w, xs, ys);
    float xt = xs + 0.5f, yt = ys + 0.5f;
    if (in_image(xt, yt, src.cols, src.rows)) {
      float w2 = sample_nearest(src.p, src.pitch, xt, yt);
      float tz = P.t[2];
      float v1z = (1.f / w3 - tz) * w;
      float res = (v1z / (1.f - w2 * tz)) * w2;
      if (res > 0.f) out = res;
    }
  }
  dst.row(0, y)[x] = out;
}

// K5: trafo3DKernelIntensityWithInvDepthGridStride (warping_registration.cu:465-501)
__global__ void __launch_bounds__(BX* BY) warp_intensity_kernel(ImgB src, ImgB prev, ImgB dst, Proj P)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  float out = qnanf();
  float w = prev.row(0, y)[x];
  if (!isnan(w)) {
    float xs, ys;
    project_pixel(P, x, y, w, xs, ys);
    float xt = xs + 0.5f, yt = ys + 0.5f;
    if (in_image(xt, yt, src.cols, src.rows)) {
      float r = sample_bilinear_q8(src.p, src.pitch, src.cols, src.rows, xt, yt);
      out = fmaxf(0.f, fminf(r, 255.f));
    }
  }
  dst.row(0, y)[x] = out;
}

// K4 + K5 of the tracker for a batch of streams in one pass, at the maps' own level, with the projection of the
// device-resident Gauss-Newton state: the per-iteration level-0 warp of WARP_ORDER = warpFirst
// (src/visodo.cpp:1087-1098).  Writes NaN where the reference's two kernels do.
template <bool TEX>
__global__ void __launch_bounds__(BX* BY)
    warp_pair_kernel(ImgB src_w, ImgB src_i, const cudaTextureObject_t* __restrict__ texW,
                     const cudaTextureObject_t* __restrict__ texI, ImgB kf_w, const GnState* __restrict__ states,
                     ImgB dst_w, ImgB dst_i, int first)
{
  const int b = blockIdx.z + first;
  const GnState& st = states[b];
  if (st.status != RGBID_OK) return;  // uniform over the CTA: lost pairs are skipped by every kernel of the schedule
  __shared__ Proj s_proj;
  const int tid = threadIdx.y * BX + threadIdx.x;
  if (tid < 12) ((float*)&s_proj)[tid] = ((const float*)&st.proj[0])[tid];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst_w.cols || y >= dst_w.rows) return;
  CurFrame cur;
  cur.Wc = src_w.row(b, 0); cur.Ic = src_i.row(b, 0);
  cur.wpitch = src_w.pitch; cur.ipitch = src_i.pitch;
  cur.texW = TEX ? texW[b] : 0; cur.texI = TEX ? texI[b] : 0;
  float w1, i1;
  warp_pixel<TEX>(s_proj, x, y, kf_w.row(b, y)[x], cur, src_w.cols, src_w.rows, true, w1, i1);
  dst_w.row(b, y)[x] = w1;
  dst_i.row(b, y)[x] = i1;
}

// Four pixels per thread through the texture unit: one float4 load of the keyframe inverse depth, the four inverse-depth
// fetches in flight together, then the four intensity fetches (warp_stage1..3 of common.cuh: the arithmetic of
// warp_pixel without its branches), float4 stores.  The scalar kernel above has one dependent fetch chain per thread.
__global__ void __launch_bounds__(256)
    warp_pair_vec_kernel(ImgB src_w, const cudaTextureObject_t* __restrict__ texW,
                         const cudaTextureObject_t* __restrict__ texI, ImgB kf_w, const GnState* __restrict__ states,
                         ImgB dst_w, ImgB dst_i, int first)
{
  const int b = blockIdx.y + first;
  const GnState& st = states[b];
  if (st.status != RGBID_OK) return;  // uniform over the CTA
  __shared__ Proj s_proj;
  const int tid = threadIdx.x;
  if (tid < 12) ((float*)&s_proj)[tid] = ((const float*)&st.proj[0])[tid];
  __syncthreads();
  const Proj proj = s_proj;
  const cudaTextureObject_t tw = texW[b], ti = texI[b];
  const int cols = src_w.cols, rows = src_w.rows;
  const int qpr = cols >> 2, total = qpr * rows;
  for (int q = blockIdx.x * blockDim.x + tid; q < total; q += gridDim.x * blockDim.x) {
    const int y = q / qpr, x0 = (q - y * qpr) * 4;
    float w0[4], w1[4], i1[4], fetched[4];
    *(float4*)w0 = __ldg((const float4*)(kf_w.row(b, y) + x0));
    WarpCoord wc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wc[k] = warp_stage1(proj, x0 + k, y, w0[k], cols, rows);
#pragma unroll
    for (int k = 0; k < 4; ++k) fetched[k] = tex2D<float>(tw, wc[k].xt, wc[k].yt);
#pragma unroll
    for (int k = 0; k < 4; ++k) w1[k] = warp_stage2(proj, x0 + k, y, w0[k], fetched[k], wc[k], cols, rows, true);
#pragma unroll
    for (int k = 0; k < 4; ++k) fetched[k] = tex2D<float>(ti, wc[k].xt, wc[k].yt);
#pragma unroll
    for (int k = 0; k < 4; ++k) i1[k] = warp_stage3(fetched[k], wc[k]);
    *(float4*)(dst_w.row(b, y) + x0) = *(float4*)w1;
    *(float4*)(dst_i.row(b, y) + x0) = *(float4*)i1;
  }
}

// Shared body of K6 (trafo3DKernelInvDepthWeightedGridStride, warping_registration.cu:549-594):
// returns the warped inverse depth (NaN if rejected) and, through weight / has_weight, the fusion weight
// (1 - w2 tz)^4 / v1z^2 when it is positive.
__device__ __forceinline__ float warp_weighted_pixel(const Proj& P, int x, int y, float w, const float* srcb,
                                                     size_t spitch, int cols, int rows, float& weight,
                                                     bool& has_weight)
{
  float out = qnanf();
  has_weight = false;
  if (!isnan(w)) {
    float xs, ys;
    float w3 = project_pixel(P, x, y, w, xs, ys);
    float xt = xs + 0.5f, yt = ys + 0.5f;
    if (in_image(xt, yt, cols, rows)) {
      float w2 = sample_nearest(srcb, spitch, xt, yt);
      float tz = P.t[2];
      float v1z = (1.f / w3 - tz) * w;
      float wf = 1.f - w2 * tz;
      float wf2 = wf * wf;
      float weight_res = (wf2 * wf2) / (v1z * v1z);
      float res = (v1z / wf) * w2;
      if (res > 0.f) out = res;
      if (weight_res > 0.f) { weight = weight_res; has_weight = true; }
    }
  }
  return out;
}

__global__ void __launch_bounds__(BX* BY) warp_invdepth_weighted_kernel(ImgB src, ImgB prev, ImgB dst, ImgB weight,
                                                                         const Proj* __restrict__ P_dev, Proj P_host,
                                                                         const int* __restrict__ active)
{
  const int b = blockIdx.z;
  if (active != nullptr && active[b] == 0) return;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  const Proj P = P_dev ? P_dev[b] : P_host;
  float wgt;
  bool has;
  float out = warp_weighted_pixel(P, x, y, prev.row(b, y)[x], src.row(b, 0), src.pitch, src.cols, src.rows, wgt, has);
  dst.row(b, y)[x] = out;
  if (has) weight.row(b, y)[x] = wgt;  // stale weights survive elsewhere, as in the reference
}

// K7: integrateWarpedFrameKernel (warping_registration.cu:637-669); gate 3 * DEPTHINV_INTEGR_TH (:80,660)
__device__ __forceinline__ void integrate_pixel(float w_sum, float w_src_weight, float& w_kf, float& kf_weight)
{
  if (isnan(w_sum)) return;
  float dw = fabsf(w_sum - w_kf);
  if (isnan(w_kf)) {
    w_kf = w_sum;
    kf_weight = w_src_weight;
  } else if (dw < 3 * 0.0075f) {
    float nw = kf_weight + w_src_weight;
    w_kf = (w_kf * kf_weight + w_sum * w_src_weight) / nw;
    kf_weight = nw;
  }
}

__global__ void __launch_bounds__(BX* BY) integrate_kernel(ImgB wsrc, ImgB wweight, ImgB dst, ImgB dweight,
                                                            const int* __restrict__ active)
{
  const int b = blockIdx.z;
  if (active != nullptr && active[b] == 0) return;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  float ws = wsrc.row(b, y)[x];
  if (isnan(ws)) return;
  float wk = dst.row(b, y)[x], kw = dweight.row(b, y)[x];
  float wk0 = wk, kw0 = kw;
  integrate_pixel(ws, wweight.row(b, y)[x], wk, kw);
  if (!(wk == wk0) || !(kw == kw0)) {
    dst.row(b, y)[x] = wk;
    dweight.row(b, y)[x] = kw;
  }
}

// K6 + K7 fused: the warped map never goes to memory.  `wstate` is the reference's warped_weight_curr_
// buffer, which is NOT cleared between frames (src/visodo.cpp:1708-1717): a pixel whose weight is not
// positive keeps the weight of an earlier frame, so that state has to be carried.
__global__ void __launch_bounds__(BX* BY) warp_integrate_kernel(ImgB cur, ImgB kf, ImgB kf_weight, ImgB wstate,
                                                                 const Proj* __restrict__ P_dev,
                                                                 const int* __restrict__ active)
{
  const int b = blockIdx.z;
  if (active != nullptr && active[b] == 0) return;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= kf.cols || y >= kf.rows) return;
  const Proj P = P_dev[b];
  float wk = kf.row(b, y)[x];
  float wgt;
  bool has;
  float ws = warp_weighted_pixel(P, x, y, wk, cur.row(b, 0), cur.pitch, cur.cols, cur.rows, wgt, has);
  if (has) wstate.row(b, y)[x] = wgt;
  if (isnan(ws)) return;
  if (!has) wgt = wstate.row(b, y)[x];
  float kw = kf_weight.row(b, y)[x];
  float wk0 = wk, kw0 = kw;
  integrate_pixel(ws, wgt, wk, kw);
  if (!(wk == wk0) || !(kw == kw0)) {
    kf.row(b, y)[x] = wk;
    kf_weight.row(b, y)[x] = kw;
  }
}

// 4 pixels per thread: float4 loads of the three keyframe maps, the transform staged in shared memory once per CTA,
// float4 write-back of whatever changed (a thread owns its four pixels, so rewriting unchanged lanes is harmless).
__global__ void __launch_bounds__(256) warp_integrate_vec_kernel(ImgB cur, ImgB kf, ImgB kf_weight, ImgB wstate,
                                                                 const Proj* __restrict__ P_dev,
                                                                 const int* __restrict__ active)
{
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  __shared__ Proj sP;
  if (threadIdx.x < 12) ((float*)&sP)[threadIdx.x] = ((const float*)&P_dev[b])[threadIdx.x];
  __syncthreads();
  const Proj P = sP;
  const int qpr = kf.cols >> 2, total = qpr * kf.rows;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    const int y = q / qpr, x0 = (q - y * qpr) * 4;
    float wk[4], kw[4], wsv[4];
    *(float4*)wk = *(const float4*)(kf.row(b, y) + x0);
    *(float4*)kw = *(const float4*)(kf_weight.row(b, y) + x0);
    *(float4*)wsv = *(const float4*)(wstate.row(b, y) + x0);
    bool ch_state = false, ch_kf = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float wgt;
      bool has;
      const float ws = warp_weighted_pixel(P, x0 + k, y, wk[k], cur.row(b, 0), cur.pitch, cur.cols, cur.rows, wgt, has);
      if (has) { wsv[k] = wgt; ch_state = true; }
      if (isnan(ws)) continue;
      if (!has) wgt = wsv[k];
      const float wk0 = wk[k], kw0 = kw[k];
      integrate_pixel(ws, wgt, wk[k], kw[k]);
      if (!(wk[k] == wk0) || !(kw[k] == kw0)) ch_kf = true;
    }
    if (ch_state) *(float4*)(wstate.row(b, y) + x0) = *(float4*)wsv;
    if (ch_kf) {
      *(float4*)(kf.row(b, y) + x0) = *(float4*)wk;
      *(float4*)(kf_weight.row(b, y) + x0) = *(float4*)kw;
    }
  }
}

// K9 + K10: partialVisibility(WithOverlapMask)Kernel + finalVisibilityReductionKernel
// (warping_registration.cu:297-461).  Counts are exact integers (the reference sums 1.f in float, which
// is exact below 2^24), reduced with a warp ballot and one integer atomic per warp.
__global__ void __launch_bounds__(BX* BY) visibility_kernel(ImgB src, ImgB dst, const Proj* __restrict__ P_dev,
                                                             Proj P_host, unsigned int* __restrict__ counts,
                                                             int count_offset, int count_stride, uint8_t* mask,
                                                             size_t mpitch, size_t mstride,
                                                             const int* __restrict__ active)
{
  const int b = blockIdx.z;
  if (active != nullptr && active[b] == 0) return;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  bool valid = false, visible = false;
  if (x < src.cols && y < src.rows) {
    const Proj P = P_dev ? P_dev[b] : P_host;
    float w = src.row(b, y)[x];
    if (!isnan(w)) {
      valid = true;
      float xd, yd;
      float wd = project_pixel(P, x, y, w, xd, yd);
      if (xd > 0.f && xd < __int2float_rn(src.cols - 1) && yd > 0.f && yd < __int2float_rn(src.rows - 1)) {
        int xi = __float2int_rn(xd), yi = __float2int_rn(yd);
        // geom_tol is ignored by the reference: 0.020 is hard-coded (:332, :405)
        if (fabsf(wd - __ldg(dst.row(b, yi) + xi)) < 0.020f) visible = true;
      }
      if (mask != nullptr) mask[(size_t)b * mstride + (size_t)y * mpitch + x] = visible ? 1 : 0;
    }
  }
  unsigned nval = __popc(__ballot_sync(0xffffffffu, valid));
  unsigned nvis = __popc(__ballot_sync(0xffffffffu, visible));
  __shared__ unsigned s_vis, s_val;
  int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid == 0) { s_vis = 0; s_val = 0; }
  __syncthreads();
  if ((tid & 31) == 0 && nval) { atomicAdd(&s_val, nval); atomicAdd(&s_vis, nvis); }
  __syncthreads();
  if (tid == 0 && s_val) {
    atomicAdd(&counts[b * count_stride + count_offset + 0], s_vis);
    atomicAdd(&counts[b * count_stride + count_offset + 1], s_val);
  }
}

// Batched variant: 4 pixels per thread (one float4 load), the stream's transform staged in shared memory once per
// CTA, grid-stride over the image with per-thread counters -> one pair of atomics per CTA.  Same per-pixel test.
__global__ void __launch_bounds__(256) visibility_vec_kernel(ImgB src, ImgB dst, const Proj* __restrict__ P_dev,
                                                             Proj P_host, unsigned int* __restrict__ counts,
                                                             int count_offset, int count_stride, uint8_t* mask,
                                                             size_t mpitch, size_t mstride,
                                                             const int* __restrict__ active)
{
  const int b = blockIdx.y;
  if (active != nullptr && active[b] == 0) return;
  __shared__ Proj sP;
  __shared__ unsigned s_vis, s_val;
  const int tid = threadIdx.x;
  if (tid < 12) ((float*)&sP)[tid] = P_dev ? ((const float*)&P_dev[b])[tid] : ((const float*)&P_host)[tid];
  if (tid == 0) { s_vis = 0; s_val = 0; }
  __syncthreads();
  const Proj P = sP;
  const int qpr = src.cols >> 2, total = qpr * src.rows;
  const float xmax = __int2float_rn(src.cols - 1), ymax = __int2float_rn(src.rows - 1);
  unsigned nval = 0, nvis = 0;
  for (int q = blockIdx.x * blockDim.x + tid; q < total; q += gridDim.x * blockDim.x) {
    const int y = q / qpr, x0 = (q - y * qpr) * 4;
    float w4[4];
    *(float4*)w4 = __ldg((const float4*)(src.row(b, y) + x0));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float w = w4[k];
      if (isnan(w)) continue;
      ++nval;
      bool visible = false;
      float xd, yd;
      const float wd = project_pixel(P, x0 + k, y, w, xd, yd);
      if (xd > 0.f && xd < xmax && yd > 0.f && yd < ymax) {
        const int xi = __float2int_rn(xd), yi = __float2int_rn(yd);
        // geom_tol is ignored by the reference: 0.020 is hard-coded (:332, :405)
        if (fabsf(wd - __ldg(dst.row(b, yi) + xi)) < 0.020f) visible = true;
      }
      if (visible) ++nvis;
      if (mask != nullptr) mask[(size_t)b * mstride + (size_t)y * mpitch + x0 + k] = visible ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nval += __shfl_xor_sync(0xffffffffu, nval, o);
    nvis += __shfl_xor_sync(0xffffffffu, nvis, o);
  }
  if ((tid & 31) == 0 && nval) { atomicAdd(&s_val, nval); atomicAdd(&s_vis, nvis); }
  __syncthreads();
  if (tid == 0 && s_val) {
    atomicAdd(&counts[b * count_stride + count_offset + 0], s_vis);
    atomicAdd(&counts[b * count_stride + count_offset + 1], s_val);
  }
}

// All four covisibility passes of one tracked frame in ONE launch (computeCovisibility is called twice per frame, each
// a pair of getVisibilityRatio calls in both directions, src/visodo.cpp:1481-1514, 2172-2186):
//   pass 0  current -> odometry keyframe      pass 1  odometry keyframe -> current
//   pass 2  current -> integration keyframe   pass 3  integration keyframe -> current
// Each of the three maps is read once (float4), the 16 projections of a thread's 4 pixels are computed first and their
// 16 gathers issued together (clamped addresses: memory-level parallelism instead of load-compare-load chains).  The
// per-pixel test is visibility_vec_kernel's, the counts are integers: bit-identical results.
// counts[b * 8 + 2 * pass + {0: visible, 1: valid}], transforms P[pass * batch + b].
#ifndef RGBID_VIS4_MINB
#define RGBID_VIS4_MINB 4
#endif
#ifndef RGBID_VIS4_PX
#define RGBID_VIS4_PX 4
#endif
constexpr int kVisPx = RGBID_VIS4_PX;  // pixels per thread and trip (4: float4 loads, 2: float2)
struct __align__(4 * kVisPx) VisVec { float v[kVisPx]; };
__global__ void __launch_bounds__(256, RGBID_VIS4_MINB) visibility4_vec_kernel(ImgB cur, ImgB kf, ImgB ikf, const Proj* __restrict__ P_dev,
                                                                 unsigned int* __restrict__ counts, int batch)
{
  const int b = blockIdx.y;
  __shared__ Proj sP[4];
  __shared__ unsigned s_cnt[8];
  const int tid = threadIdx.x;
  if (tid < 48) ((float*)&sP[tid / 12])[tid % 12] = ((const float*)&P_dev[(tid / 12) * batch + b])[tid % 12];
  if (tid < 8) s_cnt[tid] = 0;
  __syncthreads();
  // per-stream base pointers and pitches in floats: a gather is base + (yi * pitch + xi), one 32-bit multiply-add and one
  // widening add instead of the 64-bit byte arithmetic of ImgB::row (the first version spent 43 % of its issue slots on
  // the integer pipe and ran at 119 registers, 16 warps per SM: profiles/README.md)
  const float* __restrict__ bc = cur.row(b, 0);
  const float* __restrict__ bk = kf.row(b, 0);
  const float* __restrict__ bi = ikf.row(b, 0);
  const int pc = (int)(cur.pitch >> 2), pk = (int)(kf.pitch >> 2), pi = (int)(ikf.pitch >> 2);
  const int qpr = cur.cols / kVisPx, total = qpr * cur.rows;
  const float xmax = __int2float_rn(cur.cols - 1), ymax = __int2float_rn(cur.rows - 1);
  unsigned cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int q = blockIdx.x * blockDim.x + tid; q < total; q += gridDim.x * blockDim.x) {
    const int y = q / qpr, x0 = (q - y * qpr) * kVisPx;
    float wc[kVisPx], wk[kVisPx], wi[kVisPx];
    *(VisVec*)wc = *(const VisVec*)(bc + y * pc + x0);
    *(VisVec*)wk = *(const VisVec*)(bk + y * pk + x0);
    *(VisVec*)wi = *(const VisVec*)(bi + y * pi + x0);
    // valid source pixels: pass 0 and pass 2 read the same map
    {
      unsigned nc = 0, nk = 0, ni = 0;
#pragma unroll
      for (int k = 0; k < kVisPx; ++k) { nc += (wc[k] == wc[k]); nk += (wk[k] == wk[k]); ni += (wi[k] == wi[k]); }
      cnt[1] += nc; cnt[5] += nc; cnt[3] += nk; cnt[7] += ni;
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const float* src = (pass == 0 || pass == 2) ? wc : (pass == 1 ? wk : wi);
      const float* __restrict__ dbase = (pass == 0) ? bk : (pass == 2 ? bi : bc);
      const int dpitch = (pass == 0) ? pk : (pass == 2 ? pi : pc);
      // the pass's transform is re-read from shared memory (volatile: 12 broadcast loads) instead of keeping all four
      // transforms -- 48 registers -- live across the loop
      Proj Pp;
#pragma unroll
      for (int i = 0; i < 12; ++i) ((float*)&Pp)[i] = ((const volatile float*)&sP[pass])[i];
      float wd[kVisPx], got[kVisPx];
      bool ok[kVisPx];
#pragma unroll
      for (int k = 0; k < kVisPx; ++k) {
        float xd, yd;
        wd[k] = project_pixel(Pp, x0 + k, y, src[k], xd, yd);
        // a NaN source gives NaN coordinates: all four comparisons are false
        ok[k] = xd > 0.f && xd < xmax && yd > 0.f && yd < ymax;
        const int off = ok[k] ? __float2int_rn(yd) * dpitch + __float2int_rn(xd) : 0;
        got[k] = __ldg(dbase + off);
      }
#pragma unroll
      for (int k = 0; k < kVisPx; ++k)
        // geom_tol is ignored by the reference: 0.020 is hard-coded (warping_registration.cu:332, :405)
        cnt[2 * pass] += (ok[k] && fabsf(wd[k] - got[k]) < 0.020f) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    unsigned v = cnt[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0 && v) atomicAdd(&s_cnt[c], v);
  }
  __syncthreads();
  if (tid < 8 && s_cnt[tid]) atomicAdd(&counts[b * 8 + tid], s_cnt[tid]);
}

}  // namespace

void launch_visibility4(const LaunchCtx& L, ImgB cur, ImgB kf, ImgB ikf, const Proj* P_dev, unsigned int* counts, int batch)
{
  const int total = (cur.cols / kVisPx) * cur.rows;
  int gx = (total + 255) / 256;
  const int cap = (L.num_sms * 8 + batch - 1) / batch;  // ~8 CTAs of 256 threads per SM over the whole batch
  if (gx > cap) gx = cap;
  visibility4_vec_kernel<<<dim3(gx, batch), 256, 0, L.stream>>>(cur, kf, ikf, P_dev, counts, batch);
  ++*L.launches;
}

bool visibility4_applicable(const ImgB& cur, const ImgB& kf, const ImgB& ikf)
{
  auto ok = [](const ImgB& m) { return (m.cols % 4 == 0) && ((uintptr_t)m.p % 16 == 0) && (m.pitch % 16 == 0) && (m.sstride % 16 == 0); };
  return ok(cur) && ok(kf) && ok(ikf);
}

void launch_warp_invdepth(const LaunchCtx& L, ImgB src, ImgB prev, ImgB dst, const Proj& P)
{
  warp_invdepth_kernel<<<grid2d(dst.cols, dst.rows, 1), dim3(BX, BY), 0, L.stream>>>(src, prev, dst, P);
  ++*L.launches;
}

void launch_warp_intensity(const LaunchCtx& L, ImgB src, ImgB prev, ImgB dst, const Proj& P)
{
  warp_intensity_kernel<<<grid2d(dst.cols, dst.rows, 1), dim3(BX, BY), 0, L.stream>>>(src, prev, dst, P);
  ++*L.launches;
}

void launch_warp_invdepth_weighted(const LaunchCtx& L, ImgB src, ImgB prev, ImgB dst, ImgB weight, const Proj* P_dev,
                                   Proj P_host, int batch, const int* active)
{
  warp_invdepth_weighted_kernel<<<grid2d(dst.cols, dst.rows, batch), dim3(BX, BY), 0, L.stream>>>(
      src, prev, dst, weight, P_dev, P_host, active);
  ++*L.launches;
}

void launch_warp_pair(const LaunchCtx& L, ImgB src_w, ImgB src_i, const cudaTextureObject_t* texW,
                      const cudaTextureObject_t* texI, ImgB kf_w, const GnState* states, ImgB dst_w, ImgB dst_i,
                      int first, int batch)
{
  const dim3 grid = grid2d(dst_w.cols, dst_w.rows, batch), block(BX, BY);
  auto v16 = [](const ImgB& m) { return ((uintptr_t)m.p % 16 == 0) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  if (texW != nullptr && texI != nullptr && dst_w.cols % 4 == 0 && v16(kf_w) && v16(dst_w) && v16(dst_i)) {
    const int total = (dst_w.cols / 4) * dst_w.rows;
    int gx = (total + 255) / 256;
    const int cap = (L.num_sms * 8 + batch - 1) / batch;  // ~8 CTAs of 256 threads per SM over the whole batch
    if (gx > cap) gx = cap;
    warp_pair_vec_kernel<<<dim3(gx, batch), 256, 0, L.stream>>>(src_w, texW, texI, kf_w, states, dst_w, dst_i, first);
  } else if (texW != nullptr && texI != nullptr)
    warp_pair_kernel<true><<<grid, block, 0, L.stream>>>(src_w, src_i, texW, texI, kf_w, states, dst_w, dst_i, first);
  else
    warp_pair_kernel<false><<<grid, block, 0, L.stream>>>(src_w, src_i, texW, texI, kf_w, states, dst_w, dst_i, first);
  ++*L.launches;
}

void launch_integrate(const LaunchCtx& L, ImgB wsrc, ImgB wweight, ImgB dst, ImgB dweight, int batch,
                      const int* active)
{
  integrate_kernel<<<grid2d(dst.cols, dst.rows, batch), dim3(BX, BY), 0, L.stream>>>(wsrc, wweight, dst, dweight,
                                                                                    active);
  ++*L.launches;
}

void launch_warp_integrate(const LaunchCtx& L, ImgB cur, ImgB kf, ImgB kf_weight, ImgB wstate, const Proj* P_dev,
                           int batch, const int* active)
{
  auto v16 = [](const ImgB& m) { return ((uintptr_t)m.p % 16 == 0) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  if (kf.cols % 4 == 0 && v16(kf) && v16(kf_weight) && v16(wstate) && P_dev != nullptr) {
    const int total = (kf.cols / 4) * kf.rows;
    int gx = (total + 255) / 256;
    const int cap = (L.num_sms * 8 + batch - 1) / batch;
    if (gx > cap) gx = cap;
    warp_integrate_vec_kernel<<<dim3(gx, batch), 256, 0, L.stream>>>(cur, kf, kf_weight, wstate, P_dev, active);
  } else {
    warp_integrate_kernel<<<grid2d(kf.cols, kf.rows, batch), dim3(BX, BY), 0, L.stream>>>(cur, kf, kf_weight, wstate,
                                                                                         P_dev, active);
  }
  ++*L.launches;
}

void launch_visibility(const LaunchCtx& L, ImgB src, ImgB dst, const Proj* P_dev, Proj P_host, unsigned int* counts,
                       int count_offset, int count_stride, uint8_t* mask, size_t mpitch, size_t mstride, int batch,
                       const int* active)
{
  const bool vec = src.cols % 4 == 0 && ((uintptr_t)src.p % 16 == 0) && src.pitch % 16 == 0 && src.sstride % 16 == 0;
  if (vec) {
    const int total = (src.cols / 4) * src.rows;
    int gx = (total + 255) / 256;
    const int cap = (L.num_sms * 8 + batch - 1) / batch;  // ~8 CTAs of 256 threads per SM over the whole batch
    if (gx > cap) gx = cap;
    visibility_vec_kernel<<<dim3(gx, batch), 256, 0, L.stream>>>(src, dst, P_dev, P_host, counts, count_offset,
                                                                 count_stride, mask, mpitch, mstride, active);
  } else {
    visibility_kernel<<<grid2d(src.cols, src.rows, batch), dim3(BX, BY), 0, L.stream>>>(
        src, dst, P_dev, P_host, counts, count_offset, count_stride, mask, mpitch, mstride, active);
  }
  ++*L.launches;
}

}  // namespace rgbid
