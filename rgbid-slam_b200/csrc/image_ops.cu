// image_ops.cu -- frame ingest, pyramid, gradients, bilateral filter, vertex / normal maps.
//
// B200 notes: all of these are streaming / small-stencil kernels bound by HBM (or L2 once a frame pair is
// resident).  Each launch covers every stream of the batch (blockIdx.z) and, where the reference issues
// one launch per map (intensity and inverse depth separately), both maps at once.  Loads are coalesced
// along rows; the ingest kernel moves 4 pixels per thread with 32/64/128-bit accesses.
#include "kernels.cuh"

namespace rgbid {

namespace {

constexpr int BX = 32, BY = 8;

inline dim3 grid2d(int cols, int rows, int z) { return dim3((cols + BX - 1) / BX, (rows + BY - 1) / BY, z); }

#define RGBID_ACTIVE_GUARD(b) \
  if (active != nullptr && active[b] == 0) return

// ---- ingest: K18 + K19 fused (src/cuda/misc.cu:105-147) ------------------------------------------
__device__ __forceinline__ float depth_to_invdepth(int value, float inv_factor_1000)
{
  // (1/factor_depth)*1000 / clamp(d_mm, 0, 10000); 0 -> NaN   (misc.cu:116-121)
  return value > 0 ? inv_factor_1000 / __int2float_rn(min(value, 10000)) : qnanf();
}

__device__ __forceinline__ float luma(unsigned r, unsigned g, unsigned b)
{
  float v = 0.2126f * __uint2float_rn(r) + 0.7152f * __uint2float_rn(g) + 0.0722f * __uint2float_rn(b);
  return fmaxf(0.f, fminf(v, 255.f));
}

// 4 pixels per thread: 12 B of RGB (3 x 32-bit), 8 B of depth (1 x 64-bit), two float4 stores.
__global__ void __launch_bounds__(256) ingest_vec4_kernel(const uint16_t* __restrict__ depth, size_t dpitch,
                                                          size_t dstride, const uint8_t* __restrict__ rgb,
                                                          size_t cpitch, size_t cstride, ImgB W, ImgB I,
                                                          float inv_factor_1000)
{
  const int b = blockIdx.y;
  const int qpr = W.cols >> 2;
  const int total = qpr * W.rows;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    int y = q / qpr, xq = q - y * qpr;
    if (depth != nullptr) {
      const uint2 d = *(const uint2*)((const char*)depth + (size_t)b * dstride + (size_t)y * dpitch + (size_t)xq * 8);
      float4 w;
      w.x = depth_to_invdepth(d.x & 0xffff, inv_factor_1000);
      w.y = depth_to_invdepth(d.x >> 16, inv_factor_1000);
      w.z = depth_to_invdepth(d.y & 0xffff, inv_factor_1000);
      w.w = depth_to_invdepth(d.y >> 16, inv_factor_1000);
      *(float4*)(W.row(b, y) + 4 * xq) = w;
    }
    if (rgb != nullptr) {
      const uint32_t* c = (const uint32_t*)((const char*)rgb + (size_t)b * cstride + (size_t)y * cpitch + (size_t)xq * 12);
      uint32_t c0 = c[0], c1 = c[1], c2 = c[2];
      float4 i;
      i.x = luma(c0 & 0xff, (c0 >> 8) & 0xff, (c0 >> 16) & 0xff);
      i.y = luma(c0 >> 24, c1 & 0xff, (c1 >> 8) & 0xff);
      i.z = luma((c1 >> 16) & 0xff, c1 >> 24, c2 & 0xff);
      i.w = luma((c2 >> 8) & 0xff, (c2 >> 16) & 0xff, c2 >> 24);
      *(float4*)(I.row(b, y) + 4 * xq) = i;
    }
  }
}

__global__ void ingest_scalar_kernel(const uint16_t* __restrict__ depth, size_t dpitch, size_t dstride,
                                     const uint8_t* __restrict__ rgb, size_t cpitch, size_t cstride, ImgB W,
                                     ImgB I, float inv_factor_1000)
{
  const int b = blockIdx.z;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  int cols = depth ? W.cols : I.cols, rows = depth ? W.rows : I.rows;
  if (x >= cols || y >= rows) return;
  if (depth != nullptr) {
    const uint16_t* d = (const uint16_t*)((const char*)depth + (size_t)b * dstride + (size_t)y * dpitch);
    W.row(b, y)[x] = depth_to_invdepth(d[x], inv_factor_1000);
  }
  if (rgb != nullptr) {
    const uint8_t* c = rgb + (size_t)b * cstride + (size_t)y * cpitch + (size_t)x * 3;
    I.row(b, y)[x] = luma(c[0], c[1], c[2]);
  }
}

__global__ void decompose_rgb_kernel(const uint8_t* __restrict__ rgb, size_t spitch, ImgB r, ImgB g, ImgB bch)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= r.cols || y >= r.rows) return;
  const uint8_t* c = rgb + (size_t)y * spitch + (size_t)x * 3;
  r.row(0, y)[x] = __uint2float_rn(c[0]);
  g.row(0, y)[x] = __uint2float_rn(c[1]);
  bch.row(0, y)[x] = __uint2float_rn(c[2]);
}

// ---- pyramid: K16 (src/cuda/pyrdown.cu:84-132) --------------------------------------------------
__global__ void __launch_bounds__(BX* BY) pyr_down2_kernel(ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch,
                                                            const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& dst = (z < batch) ? dstA : dstB;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  const int R = 2;
  int tx = min(2 * x + R + 1, src.cols), ty = min(2 * y + R + 1, src.rows);
  float sum1 = 0.f, sum2 = 0.f;
  int count = 0;
  for (int cy = max(0, 2 * y - R); cy < ty; ++cy) {
    const float* srow = src.row(b, cy);
    for (int cx = max(0, 2 * x - R); cx < tx; ++cx) {
      float val = __ldg(srow + cx);
      if (!isnan(val)) {
        float space2 = __int2float_rn((2 * x - cx) * (2 * x - cx) + (2 * y - cy) * (2 * y - cy));
        float weight = __expf(-(space2 * 0.5f));
        sum1 += val * weight;
        sum2 += weight;
        ++count;
      }
    }
  }
  dst.row(b, y)[x] = (count > 12) ? sum1 / sum2 : qnanf();
}

// Two horizontally adjacent outputs per thread from three vector loads per source row (the scalar kernel above issues
// 25 loads per output).  Same taps in the same order (rows, then columns, ascending), the six distinct Gaussian
// weights evaluated once with the reference's expression, taps outside the image treated as NaN = skipped.
__global__ void __launch_bounds__(BX* BY) pyr_down2_vec_kernel(ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch,
                                                                const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& dst = (z < batch) ? dstA : dstB;
  const int xp = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (2 * xp >= dst.cols || y >= dst.rows) return;
  // weight(space2) = __expf(-(space2 * 0.5f)), space2 = dx^2 + dy^2 in {0, 1, 2, 4, 5, 8}
  const float wt[3][3] = {{__expf(-(0.f * 0.5f)), __expf(-(1.f * 0.5f)), __expf(-(4.f * 0.5f))},
                          {__expf(-(1.f * 0.5f)), __expf(-(2.f * 0.5f)), __expf(-(5.f * 0.5f))},
                          {__expf(-(4.f * 0.5f)), __expf(-(5.f * 0.5f)), __expf(-(8.f * 0.5f))}};
  const float nan = qnanf();
  const bool has_left = xp > 0, has_right = 4 * xp + 4 < src.cols;
  float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
  int ca = 0, cb = 0;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const int cy = 2 * y + dy;
    if (cy < 0 || cy >= src.rows) continue;
    const float* srow = src.row(b, cy) + 4 * xp;
    const float4 M = __ldg((const float4*)srow);
    float4 L = make_float4(nan, nan, nan, nan);
    if (has_left) L = __ldg((const float4*)(srow - 4));
    const float R = has_right ? __ldg(srow + 4) : nan;
    const int ady = dy < 0 ? -dy : dy;
    const float va[5] = {L.z, L.w, M.x, M.y, M.z};  // output 2 xp    : source columns 4 xp - 2 .. 4 xp + 2
    const float vb[5] = {M.x, M.y, M.z, M.w, R};    // output 2 xp + 1: source columns 4 xp     .. 4 xp + 4
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const float w = wt[ady][dx < 0 ? -dx : dx];
      const float a = va[dx + 2], c = vb[dx + 2];
      if (!isnan(a)) { s1a += a * w; s2a += w; ++ca; }
      if (!isnan(c)) { s1b += c * w; s2b += w; ++cb; }
    }
  }
  float2 out;
  out.x = (ca > 12) ? s1a / s2a : nan;
  out.y = (cb > 12) ? s1b / s2b : nan;
  *(float2*)(dst.row(b, y) + 2 * xp) = out;
}

// ---- gradients: K21 (src/cuda/misc.cu:176-220) ----------------------------------------------------
__global__ void __launch_bounds__(BX* BY) gradient2_kernel(ImgB srcA, ImgB gxA, ImgB gyA, ImgB srcB, ImgB gxB,
                                                            ImgB gyB, int batch, const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& gx = (z < batch) ? gxA : gxB;
  const ImgB& gy = (z < batch) ? gyA : gyB;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.cols || y >= src.rows) return;
  float rh = 0.f, rv = 0.f;
#pragma unroll
  for (int dx = -1; dx < 2; ++dx) {
#pragma unroll
    for (int dy = -1; dy < 2; ++dy) {
      int cx = min(max(0, x + dx), src.cols - 1);
      int cy = min(max(0, y + dy), src.rows - 1);
      float t = __ldg(src.row(b, cy) + cx);
      // zero weights are kept on purpose: 0 * NaN = NaN invalidates the pixel as in the reference
      rh += t * (float)(dx * (2 - dy * dy));
      rv += t * (float)(dy * (2 - dx * dx));
    }
  }
  gx.row(b, y)[x] = rh / 8.f;
  gy.row(b, y)[x] = rv / 8.f;
}

// Four pixels per thread: per source row one float4 plus the two clamped neighbours instead of 9 scalar loads per
// pixel; same taps, same order, same zero weights as above, float4 stores.
__global__ void __launch_bounds__(BX* BY) gradient2_vec_kernel(ImgB srcA, ImgB gxA, ImgB gyA, ImgB srcB, ImgB gxB,
                                                                ImgB gyB, int batch, const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& gx = (z < batch) ? gxA : gxB;
  const ImgB& gy = (z < batch) ? gyA : gyB;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= src.cols || y >= src.rows) return;
  float v[3][6];
#pragma unroll
  for (int dy = -1; dy < 2; ++dy) {
    const float* srow = src.row(b, min(max(0, y + dy), src.rows - 1));
    const float4 c = __ldg((const float4*)(srow + x0));
    v[dy + 1][0] = __ldg(srow + max(x0 - 1, 0));
    v[dy + 1][1] = c.x; v[dy + 1][2] = c.y; v[dy + 1][3] = c.z; v[dy + 1][4] = c.w;
    v[dy + 1][5] = __ldg(srow + min(x0 + 4, src.cols - 1));
  }
  float rh[4], rv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    rh[k] = 0.f; rv[k] = 0.f;
#pragma unroll
    for (int dx = -1; dx < 2; ++dx) {
#pragma unroll
      for (int dy = -1; dy < 2; ++dy) {
        const float t = v[dy + 1][k + dx + 1];
        rh[k] += t * (float)(dx * (2 - dy * dy));
        rv[k] += t * (float)(dy * (2 - dx * dx));
      }
    }
  }
  *(float4*)(gx.row(b, y) + x0) = make_float4(rh[0] / 8.f, rh[1] / 8.f, rh[2] / 8.f, rh[3] / 8.f);
  *(float4*)(gy.row(b, y) + x0) = make_float4(rv[0] / 8.f, rv[1] / 8.f, rv[2] / 8.f, rv[3] / 8.f);
}

// Four outputs per thread: per source row two or three float4 loads instead of 20 scalar ones (the scalar kernel issues
// ~250 instructions per pixel and is bound by the issue slots, not by its 25 ex2).  Columns outside the image enter the
// window as NaN, which the tap loop skips exactly like a NaN texel; taps, order and arithmetic are the scalar kernel's
// (bit-identical output).
__global__ void __launch_bounds__(BX* BY) bilateral2_vec_kernel(ImgB srcA, ImgB dstA, float sigmaA, ImgB srcB, ImgB dstB,
                                                                float sigmaB, int batch, const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& dst = (z < batch) ? dstA : dstB;
  const float sigma_floatmap = (z < batch) ? sigmaA : sigmaB;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= src.cols || y >= src.rows) return;
  const float nan = qnanf();
  const bool has_left = x0 > 0, has_right = x0 + 4 < src.cols;
  float value[4];
  *(float4*)value = __ldg((const float4*)(src.row(b, y) + x0));
  const float s2ih = 0.5f / (5.f * 5.f);  // sigma_space = 5 (filters.cu:83)
  const float rsigma = __fdividef(1.f, sigma_floatmap);
  float sum1[4] = {0.f, 0.f, 0.f, 0.f}, sum2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const int cy = y + dy;
    if (cy < 0 || cy >= src.rows) continue;
    const float* srow = src.row(b, cy) + x0;
    const float4 M = __ldg((const float4*)srow);
    float4 Lf = make_float4(nan, nan, nan, nan), Rf = make_float4(nan, nan, nan, nan);
    if (has_left) Lf = __ldg((const float4*)(srow - 4));
    if (has_right) Rf = __ldg((const float4*)(srow + 4));
    const float w[8] = {Lf.z, Lf.w, M.x, M.y, M.z, M.w, Rf.x, Rf.y};  // columns x0 - 2 .. x0 + 5
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const float tmp = w[k + dx + 2];
        if (!isnan(tmp)) {
          const float space2 = (float)(dx * dx + dy * dy);
          const float fn = (value[k] - tmp) * rsigma;
          const float weight = __expf(-(s2ih * space2 + 0.5f * fn * fn));
          sum1[k] += tmp * weight;
          sum2[k] += weight;
        }
      }
    }
  }
  float out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = isnan(value[k]) ? nan : sum1[k] / sum2[k];
  *(float4*)(dst.row(b, y) + x0) = *(float4*)out;
}

// One launch for a LIST of maps of different sizes (all pyramid levels of the raw and of the filtered keyframe: 16
// Sobel passes, or the 8 copies of saveCurrentImagesAsOdoKeyframes): the coarse levels are a few microseconds of work
// each and paid a launch apiece.  blockIdx.x walks the tiles of all list entries; per entry the body is the per-map
// kernel's (same taps, same order: bit-identical).
constexpr int kMaxListMaps = 16;
struct MapList {
  ImgB a[kMaxListMaps], b[kMaxListMaps], c[kMaxListMaps];  // gradient: src, gx, gy; copy: src, dst, -
  int tile_begin[kMaxListMaps + 1];
  int tiles_x[kMaxListMaps];
  int n;
};

__device__ __forceinline__ bool list_locate(const MapList& Lm, int& slot, int& x0, int& y)
{
  const int t = blockIdx.x;
  slot = 0;
#pragma unroll 1
  while (slot + 1 < Lm.n && t >= Lm.tile_begin[slot + 1]) ++slot;
  const int tile = t - Lm.tile_begin[slot];
  const int tx = tile % Lm.tiles_x[slot], ty = tile / Lm.tiles_x[slot];
  x0 = 4 * (tx * BX + threadIdx.x);
  y = ty * BY + threadIdx.y;
  return x0 < Lm.a[slot].cols && y < Lm.a[slot].rows;
}

__global__ void __launch_bounds__(BX* BY) gradient_list_kernel(const __grid_constant__ MapList Lm, const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  int slot, x0, y;
  if (!list_locate(Lm, slot, x0, y)) return;
  const ImgB& src = Lm.a[slot];
  const ImgB& gx = Lm.b[slot];
  const ImgB& gy = Lm.c[slot];
  float v[3][6];
#pragma unroll
  for (int dy = -1; dy < 2; ++dy) {
    const float* srow = src.row(b, min(max(0, y + dy), src.rows - 1));
    const float4 c = __ldg((const float4*)(srow + x0));
    v[dy + 1][0] = __ldg(srow + max(x0 - 1, 0));
    v[dy + 1][1] = c.x; v[dy + 1][2] = c.y; v[dy + 1][3] = c.z; v[dy + 1][4] = c.w;
    v[dy + 1][5] = __ldg(srow + min(x0 + 4, src.cols - 1));
  }
  float rh[4], rv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    rh[k] = 0.f; rv[k] = 0.f;
#pragma unroll
    for (int dx = -1; dx < 2; ++dx) {
#pragma unroll
      for (int dy = -1; dy < 2; ++dy) {
        const float t = v[dy + 1][k + dx + 1];
        rh[k] += t * (float)(dx * (2 - dy * dy));
        rv[k] += t * (float)(dy * (2 - dx * dx));
      }
    }
  }
  *(float4*)(gx.row(b, y) + x0) = make_float4(rh[0] / 8.f, rh[1] / 8.f, rh[2] / 8.f, rh[3] / 8.f);
  *(float4*)(gy.row(b, y) + x0) = make_float4(rv[0] / 8.f, rv[1] / 8.f, rv[2] / 8.f, rv[3] / 8.f);
}

__global__ void __launch_bounds__(BX* BY) copy_list_kernel(const __grid_constant__ MapList Lm, const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  int slot, x0, y;
  if (!list_locate(Lm, slot, x0, y)) return;
  *(float4*)(Lm.b[slot].row(b, y) + x0) = *(const float4*)(Lm.a[slot].row(b, y) + x0);
}

// ---- bilateral: K23 (src/cuda/filters.cu:86-135) ----------------------------------------------------
__global__ void __launch_bounds__(BX* BY) bilateral2_kernel(ImgB srcA, ImgB dstA, float sigmaA, ImgB srcB,
                                                             ImgB dstB, float sigmaB, int batch,
                                                             const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& dst = (z < batch) ? dstA : dstB;
  const float sigma_floatmap = (z < batch) ? sigmaA : sigmaB;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.cols || y >= src.rows) return;
  float value = __ldg(src.row(b, y) + x);
  if (isnan(value)) { dst.row(b, y)[x] = qnanf(); return; }
  const float s2ih = 0.5f / (5.f * 5.f);  // sigma_space = 5 (filters.cu:83)
  // (value - tmp) / sigma under --prec-div=false is (value - tmp) * rcp.approx(sigma): the reciprocal is the same for
  // all 25 taps, so it is taken once (25 MUFU per pixel instead of 50 -- the kernel is bound by the SFU pipe); the taps
  // are unrolled over fixed offsets so that the spatial term s2ih * (dx^2 + dy^2) is a compile-time constant -- the same
  // single-precision product the loop computed.  Same taps, same order, same arithmetic: bit-identical output.
  const float rsigma = __fdividef(1.f, sigma_floatmap);  // rcp.approx.ftz, the reciprocal div.approx multiplies by
  float sum1 = 0.f, sum2 = 0.f;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const int cy = y + dy;
    if (cy < 0 || cy >= src.rows) continue;
    const float* srow = src.row(b, cy);
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int cx = x + dx;
      if (cx < 0 || cx >= src.cols) continue;
      const float tmp = __ldg(srow + cx);
      if (!isnan(tmp)) {
        const float space2 = (float)(dx * dx + dy * dy);
        const float fn = (value - tmp) * rsigma;
        const float weight = __expf(-(s2ih * space2 + 0.5f * fn * fn));
        sum1 += tmp * weight;
        sum2 += weight;
      }
    }
  }
  dst.row(b, y)[x] = sum1 / sum2;
}

// ---- copies / fills: K22 --------------------------------------------------------------------------
__global__ void copy2_kernel(ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch, const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& dst = (z < batch) ? dstA : dstB;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.cols || y >= src.rows) return;
  dst.row(b, y)[x] = src.row(b, y)[x];
}

__global__ void copy2_vec_kernel(ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch, const int* __restrict__ active)
{
  const int z = blockIdx.z;
  const int b = z % batch;
  RGBID_ACTIVE_GUARD(b);
  const ImgB& src = (z < batch) ? srcA : srcB;
  const ImgB& dst = (z < batch) ? dstA : dstB;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= src.cols || y >= src.rows) return;
  *(float4*)(dst.row(b, y) + x0) = *(const float4*)(src.row(b, y) + x0);
}

__global__ void fill_kernel(ImgB dst, float value, const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  dst.row(b, y)[x] = value;
}

__global__ void fill_u8_kernel(uint8_t* dst, size_t pitch, size_t sstride, int rows, int cols, uint8_t value,
                               const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  dst[(size_t)b * sstride + (size_t)y * pitch + x] = value;
}

// ---- vertex / normal maps: K24 (src/cuda/maps.cu:63-90, 134-179) ------------------------------------
__global__ void vmap_kernel(ImgB depth_inv, ImgB vmap, float fx_inv, float fy_inv, float cx, float cy,
                            const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  const int rows = depth_inv.rows;
  if (u >= depth_inv.cols || v >= rows) return;
  float z = 1.f / depth_inv.row(b, v)[u];
  if (!isnan(z)) {
    vmap.row(b, v)[u] = z * (__int2float_rn(u) - cx) * fx_inv;
    vmap.row(b, v + rows)[u] = z * (__int2float_rn(v) - cy) * fy_inv;
    vmap.row(b, v + 2 * rows)[u] = z;
  } else {
    vmap.row(b, v)[u] = qnanf();  // only the x plane is invalidated, as in the reference
  }
}

__global__ void nmap_gradients_kernel(ImgB depth_inv, ImgB gx_, ImgB gy_, ImgB nmap, float fx, float fy, float cx,
                                      float cy, const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  const int rows = depth_inv.rows;
  if (u >= depth_inv.cols || v >= rows) return;
  float nx_out = qnanf();
  float w = depth_inv.row(b, v)[u], gx = gx_.row(b, v)[u], gy = gy_.row(b, v)[u];
  if (!(isnan(w) || isnan(gx) || isnan(gy))) {
    float nx = gx * fx, ny = gy * fy;
    float nz = gx * (cx - __int2float_rn(u)) + gy * (cy - __int2float_rn(v)) + w;
    float rn = rsqrtf(nx * nx + ny * ny + nz * nz);
    nx *= rn; ny *= rn; nz *= rn;
    float z = 1.f / w;
    float vx = z * (__int2float_rn(u) - cx) * (1.f / fx);
    float vy = z * (__int2float_rn(v) - cy) * (1.f / fy);
    float rv = rsqrtf(vx * vx + vy * vy + z * z);
    float d = (vx * rv) * nx + (vy * rv) * ny + (z * rv) * nz;
    if (d > 0.1f) {  // grazing-angle cut (maps.cu:170)
      nx_out = nx;
      nmap.row(b, v + rows)[u] = ny;
      nmap.row(b, v + 2 * rows)[u] = nz;
    }
  }
  nmap.row(b, v)[u] = nx_out;
}

// 4 pixels per thread.  The x plane is always written (value or NaN); the y / z planes only where the reference
// writes them, with one float4 store when all four pixels qualify.
__global__ void __launch_bounds__(BX* BY) vmap_vec_kernel(ImgB depth_inv, ImgB vmap, float fx_inv, float fy_inv, float cx,
                                                           float cy, const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  const int u0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), v = blockIdx.y * blockDim.y + threadIdx.y;
  const int rows = depth_inv.rows;
  if (u0 >= depth_inv.cols || v >= rows) return;
  float w[4], vx[4], vy[4], vz[4];
  *(float4*)w = *(const float4*)(depth_inv.row(b, v) + u0);
  bool ok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float z = 1.f / w[k];
    ok[k] = !isnan(z);
    vx[k] = ok[k] ? z * (__int2float_rn(u0 + k) - cx) * fx_inv : qnanf();
    vy[k] = z * (__int2float_rn(v) - cy) * fy_inv;
    vz[k] = z;
  }
  *(float4*)(vmap.row(b, v) + u0) = *(float4*)vx;
  if (ok[0] && ok[1] && ok[2] && ok[3]) {
    *(float4*)(vmap.row(b, v + rows) + u0) = *(float4*)vy;
    *(float4*)(vmap.row(b, v + 2 * rows) + u0) = *(float4*)vz;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) { vmap.row(b, v + rows)[u0 + k] = vy[k]; vmap.row(b, v + 2 * rows)[u0 + k] = vz[k]; }
  }
}

__global__ void __launch_bounds__(BX* BY) nmap_gradients_vec_kernel(ImgB depth_inv, ImgB gx_, ImgB gy_, ImgB nmap, float fx,
                                                                     float fy, float cx, float cy,
                                                                     const int* __restrict__ active)
{
  const int b = blockIdx.z;
  RGBID_ACTIVE_GUARD(b);
  const int u0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), v = blockIdx.y * blockDim.y + threadIdx.y;
  const int rows = depth_inv.rows;
  if (u0 >= depth_inv.cols || v >= rows) return;
  float w4[4], gx4[4], gy4[4], ox[4], oy[4], oz[4];
  *(float4*)w4 = *(const float4*)(depth_inv.row(b, v) + u0);
  *(float4*)gx4 = *(const float4*)(gx_.row(b, v) + u0);
  *(float4*)gy4 = *(const float4*)(gy_.row(b, v) + u0);
  bool ok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int u = u0 + k;
    const float w = w4[k], gx = gx4[k], gy = gy4[k];
    ox[k] = qnanf(); oy[k] = 0.f; oz[k] = 0.f; ok[k] = false;
    if (!(isnan(w) || isnan(gx) || isnan(gy))) {
      float nx = gx * fx, ny = gy * fy;
      float nz = gx * (cx - __int2float_rn(u)) + gy * (cy - __int2float_rn(v)) + w;
      float rn = rsqrtf(nx * nx + ny * ny + nz * nz);
      nx *= rn; ny *= rn; nz *= rn;
      float z = 1.f / w;
      float vx = z * (__int2float_rn(u) - cx) * (1.f / fx);
      float vy = z * (__int2float_rn(v) - cy) * (1.f / fy);
      float rv = rsqrtf(vx * vx + vy * vy + z * z);
      float d = (vx * rv) * nx + (vy * rv) * ny + (z * rv) * nz;
      if (d > 0.1f) { ox[k] = nx; oy[k] = ny; oz[k] = nz; ok[k] = true; }  // grazing-angle cut (maps.cu:170)
    }
  }
  *(float4*)(nmap.row(b, v) + u0) = *(float4*)ox;
  if (ok[0] && ok[1] && ok[2] && ok[3]) {
    *(float4*)(nmap.row(b, v + rows) + u0) = *(float4*)oy;
    *(float4*)(nmap.row(b, v + 2 * rows) + u0) = *(float4*)oz;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) { nmap.row(b, v + rows)[u0 + k] = oy[k]; nmap.row(b, v + 2 * rows)[u0 + k] = oz[k]; }
  }
}

inline bool aligned(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }

}  // namespace

// --------------------------------------------------------------------------------------------------
void launch_ingest(const LaunchCtx& L, const uint16_t* depth, size_t dpitch, size_t dstride, const uint8_t* rgb,
                   size_t cpitch, size_t cstride, ImgB W, ImgB I, int batch, float factor_depth)
{
  const ImgB& ref = depth ? W : I;
  float inv_factor_1000 = (1.f / factor_depth) * 1000.f;
  bool vec = (ref.cols % 4 == 0);
  if (depth) vec = vec && aligned(depth, 8) && dpitch % 8 == 0 && dstride % 8 == 0 && aligned(W.p, 16) && W.pitch % 16 == 0 && W.sstride % 16 == 0;
  if (rgb) vec = vec && aligned(rgb, 4) && cpitch % 4 == 0 && cstride % 4 == 0 && aligned(I.p, 16) && I.pitch % 16 == 0 && I.sstride % 16 == 0;
  if (vec) {
    int total = (ref.cols / 4) * ref.rows;
    int gx = (total + 255) / 256;
    ingest_vec4_kernel<<<dim3(gx, batch), 256, 0, L.stream>>>(depth, dpitch, dstride, rgb, cpitch, cstride, W, I,
                                                             inv_factor_1000);
  } else {
    ingest_scalar_kernel<<<grid2d(ref.cols, ref.rows, batch), dim3(BX, BY), 0, L.stream>>>(
        depth, dpitch, dstride, rgb, cpitch, cstride, W, I, inv_factor_1000);
  }
  ++*L.launches;
}

void launch_depth_to_invdepth(const LaunchCtx& L, const uint16_t* src, size_t spitch, size_t sstride, ImgB dst,
                              int batch, float factor_depth)
{
  launch_ingest(L, src, spitch, sstride, nullptr, 0, 0, dst, dst, batch, factor_depth);
}

void launch_intensity(const LaunchCtx& L, const uint8_t* rgb, size_t spitch, size_t sstride, ImgB dst, int batch)
{
  launch_ingest(L, nullptr, 0, 0, rgb, spitch, sstride, dst, dst, batch, 1.f);
}

void launch_decompose_rgb(const LaunchCtx& L, const uint8_t* rgb, size_t spitch, ImgB r, ImgB g, ImgB b)
{
  decompose_rgb_kernel<<<grid2d(r.cols, r.rows, 1), dim3(BX, BY), 0, L.stream>>>(rgb, spitch, r, g, b);
  ++*L.launches;
}

void launch_pyr_down2(const LaunchCtx& L, ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch, const int* active)
{
  int nm = srcB.p ? 2 : 1;
  auto vec_ok = [](const ImgB& s, const ImgB& d) {
    return s.cols % 4 == 0 && d.cols * 2 == s.cols && aligned(s.p, 16) && s.pitch % 16 == 0 && s.sstride % 16 == 0 &&
           aligned(d.p, 8) && d.pitch % 8 == 0 && d.sstride % 8 == 0;
  };
  if (vec_ok(srcA, dstA) && (nm == 1 || vec_ok(srcB, dstB)))
    pyr_down2_vec_kernel<<<grid2d(dstA.cols / 2, dstA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(srcA, dstA, srcB,
                                                                                                 dstB, batch, active);
  else
    pyr_down2_kernel<<<grid2d(dstA.cols, dstA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(srcA, dstA, srcB, dstB,
                                                                                              batch, active);
  ++*L.launches;
}

void launch_gradient2(const LaunchCtx& L, ImgB srcA, ImgB gxA, ImgB gyA, ImgB srcB, ImgB gxB, ImgB gyB, int batch,
                      const int* active)
{
  int nm = srcB.p ? 2 : 1;
  auto v16 = [](const ImgB& m) { return aligned(m.p, 16) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  bool vec = srcA.cols % 4 == 0 && v16(srcA) && v16(gxA) && v16(gyA);
  if (nm == 2) vec = vec && v16(srcB) && v16(gxB) && v16(gyB);
  if (vec)
    gradient2_vec_kernel<<<grid2d(srcA.cols / 4, srcA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(
        srcA, gxA, gyA, srcB, gxB, gyB, batch, active);
  else
    gradient2_kernel<<<grid2d(srcA.cols, srcA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(srcA, gxA, gyA, srcB,
                                                                                              gxB, gyB, batch, active);
  ++*L.launches;
}

void launch_bilateral2(const LaunchCtx& L, ImgB srcA, ImgB dstA, float sigmaA, ImgB srcB, ImgB dstB, float sigmaB,
                       int batch, const int* active)
{
  int nm = srcB.p ? 2 : 1;
  auto v16 = [](const ImgB& m) { return aligned(m.p, 16) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  bool vec = srcA.cols % 4 == 0 && v16(srcA) && v16(dstA);
  if (nm == 2) vec = vec && v16(srcB) && v16(dstB);
  if (vec)
    bilateral2_vec_kernel<<<grid2d(srcA.cols / 4, srcA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(
        srcA, dstA, sigmaA, srcB, dstB, sigmaB, batch, active);
  else
    bilateral2_kernel<<<grid2d(srcA.cols, srcA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(
        srcA, dstA, sigmaA, srcB, dstB, sigmaB, batch, active);
  ++*L.launches;
}

void launch_copy2(const LaunchCtx& L, ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch, const int* active)
{
  int nm = srcB.p ? 2 : 1;
  auto v16 = [](const ImgB& m) { return aligned(m.p, 16) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  bool vec = srcA.cols % 4 == 0 && v16(srcA) && v16(dstA);
  if (nm == 2) vec = vec && v16(srcB) && v16(dstB);
  if (vec)
    copy2_vec_kernel<<<grid2d(srcA.cols / 4, srcA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(srcA, dstA, srcB, dstB,
                                                                                             batch, active);
  else
    copy2_kernel<<<grid2d(srcA.cols, srcA.rows, batch * nm), dim3(BX, BY), 0, L.stream>>>(srcA, dstA, srcB, dstB, batch,
                                                                                          active);
  ++*L.launches;
}

static bool list_ok(const ImgB& m) { return m.cols % 4 == 0 && aligned(m.p, 16) && m.pitch % 16 == 0 && m.sstride % 16 == 0; }

static void list_finish(MapList& Lm)
{
  int t = 0;
  for (int i = 0; i < Lm.n; ++i) {
    Lm.tile_begin[i] = t;
    Lm.tiles_x[i] = (Lm.a[i].cols / 4 + BX - 1) / BX;
    t += Lm.tiles_x[i] * ((Lm.a[i].rows + BY - 1) / BY);
  }
  Lm.tile_begin[Lm.n] = t;
}

bool launch_gradient_list(const LaunchCtx& L, const ImgB* src, const ImgB* gx, const ImgB* gy, int n, int batch, const int* active)
{
  if (n < 1 || n > kMaxListMaps) return false;
  MapList Lm;
  memset(&Lm, 0, sizeof(Lm));
  for (int i = 0; i < n; ++i) {
    if (!list_ok(src[i]) || !list_ok(gx[i]) || !list_ok(gy[i])) return false;
    Lm.a[i] = src[i]; Lm.b[i] = gx[i]; Lm.c[i] = gy[i];
  }
  Lm.n = n;
  list_finish(Lm);
  gradient_list_kernel<<<dim3(Lm.tile_begin[n], 1, batch), dim3(BX, BY), 0, L.stream>>>(Lm, active);
  ++*L.launches;
  return true;
}

bool launch_copy_list(const LaunchCtx& L, const ImgB* src, const ImgB* dst, int n, int batch, const int* active)
{
  if (n < 1 || n > kMaxListMaps) return false;
  MapList Lm;
  memset(&Lm, 0, sizeof(Lm));
  for (int i = 0; i < n; ++i) {
    if (!list_ok(src[i]) || !list_ok(dst[i])) return false;
    Lm.a[i] = src[i]; Lm.b[i] = dst[i];
  }
  Lm.n = n;
  list_finish(Lm);
  copy_list_kernel<<<dim3(Lm.tile_begin[n], 1, batch), dim3(BX, BY), 0, L.stream>>>(Lm, active);
  ++*L.launches;
  return true;
}

void launch_fill(const LaunchCtx& L, ImgB dst, float value, int batch, const int* active)
{
  fill_kernel<<<grid2d(dst.cols, dst.rows, batch), dim3(BX, BY), 0, L.stream>>>(dst, value, active);
  ++*L.launches;
}

void launch_fill_u8(const LaunchCtx& L, uint8_t* dst, size_t pitch, size_t sstride, int rows, int cols,
                    uint8_t value, int batch, const int* active)
{
  fill_u8_kernel<<<grid2d(cols, rows, batch), dim3(BX, BY), 0, L.stream>>>(dst, pitch, sstride, rows, cols, value,
                                                                          active);
  ++*L.launches;
}

// createVMap + computeGradientDepth + createNMapGradients of the fused keyframe in one pass (src/visodo.cpp:889-892,
// 1753-1758): the three source rows are loaded once, the Sobel pair is written out and fed to the normals from
// registers.  Per output the arithmetic is that of gradient2_vec_kernel / vmap_vec_kernel / nmap_gradients_vec_kernel
// (bit-identical results); 39 MB read + 8 planes written instead of three launches re-reading their inputs.
__global__ void __launch_bounds__(BX* BY) keyframe_maps_vec_kernel(ImgB depth_inv, ImgB gx_, ImgB gy_, ImgB vmap, ImgB nmap,
                                                                   float fx, float fy, float cx, float cy)
{
  const int b = blockIdx.z;
  const int u0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), v = blockIdx.y * blockDim.y + threadIdx.y;
  const int rows = depth_inv.rows, cols = depth_inv.cols;
  if (u0 >= cols || v >= rows) return;
  float t[3][6];
#pragma unroll
  for (int dy = -1; dy < 2; ++dy) {
    const float* srow = depth_inv.row(b, min(max(0, v + dy), rows - 1));
    const float4 c = __ldg((const float4*)(srow + u0));
    t[dy + 1][0] = __ldg(srow + max(u0 - 1, 0));
    t[dy + 1][1] = c.x; t[dy + 1][2] = c.y; t[dy + 1][3] = c.z; t[dy + 1][4] = c.w;
    t[dy + 1][5] = __ldg(srow + min(u0 + 4, cols - 1));
  }
  float gx4[4], gy4[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float rh = 0.f, rv = 0.f;
#pragma unroll
    for (int dx = -1; dx < 2; ++dx) {
#pragma unroll
      for (int dy = -1; dy < 2; ++dy) {
        const float s = t[dy + 1][k + dx + 1];
        rh += s * (float)(dx * (2 - dy * dy));
        rv += s * (float)(dy * (2 - dx * dx));
      }
    }
    gx4[k] = rh / 8.f; gy4[k] = rv / 8.f;
  }
  *(float4*)(gx_.row(b, v) + u0) = *(float4*)gx4;
  *(float4*)(gy_.row(b, v) + u0) = *(float4*)gy4;
  const float fx_inv = 1.f / fx, fy_inv = 1.f / fy;
  float vx[4], vy[4], vz[4], ox[4], oy[4], oz[4];
  bool okv[4], okn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int u = u0 + k;
    const float w = t[1][k + 1];
    {  // vmap_vec_kernel
      const float z = 1.f / w;
      okv[k] = !isnan(z);
      vx[k] = okv[k] ? z * (__int2float_rn(u) - cx) * fx_inv : qnanf();
      vy[k] = z * (__int2float_rn(v) - cy) * fy_inv;
      vz[k] = z;
    }
    {  // nmap_gradients_vec_kernel
      const float gx = gx4[k], gy = gy4[k];
      ox[k] = qnanf(); oy[k] = 0.f; oz[k] = 0.f; okn[k] = false;
      if (!(isnan(w) || isnan(gx) || isnan(gy))) {
        float nx = gx * fx, ny = gy * fy;
        float nz = gx * (cx - __int2float_rn(u)) + gy * (cy - __int2float_rn(v)) + w;
        float rn = rsqrtf(nx * nx + ny * ny + nz * nz);
        nx *= rn; ny *= rn; nz *= rn;
        float z = 1.f / w;
        float px = z * (__int2float_rn(u) - cx) * (1.f / fx);
        float py = z * (__int2float_rn(v) - cy) * (1.f / fy);
        float rv = rsqrtf(px * px + py * py + z * z);
        float d = (px * rv) * nx + (py * rv) * ny + (z * rv) * nz;
        if (d > 0.1f) { ox[k] = nx; oy[k] = ny; oz[k] = nz; okn[k] = true; }  // grazing-angle cut (maps.cu:170)
      }
    }
  }
  *(float4*)(vmap.row(b, v) + u0) = *(float4*)vx;
  if (okv[0] && okv[1] && okv[2] && okv[3]) {
    *(float4*)(vmap.row(b, v + rows) + u0) = *(float4*)vy;
    *(float4*)(vmap.row(b, v + 2 * rows) + u0) = *(float4*)vz;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (okv[k]) { vmap.row(b, v + rows)[u0 + k] = vy[k]; vmap.row(b, v + 2 * rows)[u0 + k] = vz[k]; }
  }
  *(float4*)(nmap.row(b, v) + u0) = *(float4*)ox;
  if (okn[0] && okn[1] && okn[2] && okn[3]) {
    *(float4*)(nmap.row(b, v + rows) + u0) = *(float4*)oy;
    *(float4*)(nmap.row(b, v + 2 * rows) + u0) = *(float4*)oz;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (okn[k]) { nmap.row(b, v + rows)[u0 + k] = oy[k]; nmap.row(b, v + 2 * rows)[u0 + k] = oz[k]; }
  }
}

bool launch_keyframe_maps(const LaunchCtx& L, ImgB depth_inv, ImgB gx, ImgB gy, ImgB vmap, ImgB nmap, float fx, float fy,
                          float cx, float cy, int batch)
{
  auto v16 = [](const ImgB& m) { return aligned(m.p, 16) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  if (!(depth_inv.cols % 4 == 0 && v16(depth_inv) && v16(gx) && v16(gy) && v16(vmap) && v16(nmap))) return false;
  keyframe_maps_vec_kernel<<<grid2d(depth_inv.cols / 4, depth_inv.rows, batch), dim3(BX, BY), 0, L.stream>>>(
      depth_inv, gx, gy, vmap, nmap, fx, fy, cx, cy);
  ++*L.launches;
  return true;
}

void launch_vmap(const LaunchCtx& L, ImgB depth_inv, ImgB vmap, float fx, float fy, float cx, float cy, int batch,
                 const int* active)
{
  auto v16 = [](const ImgB& m) { return aligned(m.p, 16) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  if (depth_inv.cols % 4 == 0 && v16(depth_inv) && v16(vmap))
    vmap_vec_kernel<<<grid2d(depth_inv.cols / 4, depth_inv.rows, batch), dim3(BX, BY), 0, L.stream>>>(
        depth_inv, vmap, 1.f / fx, 1.f / fy, cx, cy, active);
  else
    vmap_kernel<<<grid2d(depth_inv.cols, depth_inv.rows, batch), dim3(BX, BY), 0, L.stream>>>(
        depth_inv, vmap, 1.f / fx, 1.f / fy, cx, cy, active);
  ++*L.launches;
}

void launch_nmap_gradients(const LaunchCtx& L, ImgB depth_inv, ImgB gx, ImgB gy, ImgB nmap, float fx, float fy,
                           float cx, float cy, int batch, const int* active)
{
  auto v16 = [](const ImgB& m) { return aligned(m.p, 16) && m.pitch % 16 == 0 && m.sstride % 16 == 0; };
  if (depth_inv.cols % 4 == 0 && v16(depth_inv) && v16(gx) && v16(gy) && v16(nmap))
    nmap_gradients_vec_kernel<<<grid2d(depth_inv.cols / 4, depth_inv.rows, batch), dim3(BX, BY), 0, L.stream>>>(
        depth_inv, gx, gy, nmap, fx, fy, cx, cy, active);
  else
    nmap_gradients_kernel<<<grid2d(depth_inv.cols, depth_inv.rows, batch), dim3(BX, BY), 0, L.stream>>>(
        depth_inv, gx, gy, nmap, fx, fy, cx, cy, active);
  ++*L.launches;
}

}  // namespace rgbid
