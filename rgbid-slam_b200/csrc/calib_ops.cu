// calib_ops.cu -- the steps either side of the hot path that SURVEY section 8 (f) ranks next:
//   f3  custom-calibration ingest: undistortIntensity / undistortDepthInv (src/cuda/undistortion.cu:94-310) and
//       registerDepthinv (src/cuda/warping_registration.cu:148-281, 597-635, 720-800), used by
//       prepareImagesCustomCalibration (src/visodo.cpp:775-823) when custom_registration = 1;
//   f4  the dormant colour fusion integrateWarpedRGB (warping_registration.cu:672-712, caller commented out at
//       src/visodo.cpp:2214) and the shaded previews generateImage / generateImageRGB (src/cuda/image_generator.cu).
//
// B200 notes.  All of these are one-pass streaming kernels bound by HBM.  Against the reference:
//   * undistortDepthInv is ONE kernel: the reference writes the corrected map, wraps it in a point-filtered texture and
//     gathers from it; here the gather position is computed first and the correction is evaluated on the gathered texel,
//     so the intermediate map (1.2 MB written + read per frame) never exists;
//   * registerDepthinv is two kernels instead of four plus a texture object: the z-buffer splat (atomicMax on the float
//     bit pattern, order-independent and therefore deterministic) and the homography gather, which reads the integer
//     canvas directly (0 = empty) instead of a converted float copy; the canvas is cleared with a memset node.
// Per-pixel arithmetic follows the reference expression by expression (same nvcc numeric flags).
#include "kernels.cuh"

namespace rgbid {

namespace {

constexpr int BX = 32, BY = 8;
inline dim3 grid2d(int cols, int rows, int z = 1) { return dim3((cols + BX - 1) / BX, (rows + BY - 1) / BY, z); }

// distortPixel, undistortion.cu:94-111
__device__ __forceinline__ void distort_pixel(float uu, float vu, float& ud, float& vd, const rgbid_intr& intr)
{
  float r2 = uu * uu + vu * vu;
  float r4 = r2 * r2;
  float r6 = r2 * r4;
  float factor_r = 1.f + intr.k1 * r2 + intr.k2 * r4 + intr.k5 * r6;
  ud = factor_r * uu;
  ud += 2.f * intr.k3 * uu * vu + intr.k4 * (r2 + 2.f * uu * uu);
  vd = factor_r * vu;
  vd += 2.f * intr.k4 * uu * vu + intr.k3 * (r2 + 2.f * vu * vu);
}

// source position of undistorted pixel (xu, yu), undistortion.cu:156-166; false if it falls outside the image
__device__ __forceinline__ bool undistort_source(int xu, int yu, const rgbid_intr& intr, int cols, int rows, float& xd,
                                                 float& yd)
{
  float uu = (__int2float_rn(xu) - intr.cx) * (1.f / intr.fx);
  float vu = (__int2float_rn(yu) - intr.cy) * (1.f / intr.fy);
  float ud, vd;
  distort_pixel(uu, vu, ud, vd, intr);
  xd = intr.fx * ud + intr.cx + 0.5f;
  yd = intr.fy * vd + intr.cy + 0.5f;
  return !((xd <= 0) || (yd <= 0) || (xd >= cols) || (yd >= rows));
}

// K25a: undistortKernel on a linear-filtered texture (undistortIntensity, :212-257)
__global__ void __launch_bounds__(BX* BY) undistort_intensity_kernel(ImgB src, ImgB dst, rgbid_intr intr)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  float res = qnanf(), xd, yd;
  if (undistort_source(x, y, intr, dst.cols, dst.rows, xd, yd))
    res = sample_bilinear_q8(src.p, src.pitch, src.cols, src.rows, xd, yd);
  dst.row(0, y)[x] = res;
}

// correctDepthinv, undistortion.cu:113-137
__device__ __forceinline__ float correct_depthinv(float u, float v, float wm, const rgbid_depth_dist& dp)
{
  float wd = dp.c1 * wm + dp.c0;
  float r2 = u * u + v * v;
  float r4 = r2 * r2;
  float r6 = r2 * r4;
  float uv = u * v;
  float u2v = u * u * v;
  float uv2 = u * v * v;
  float D0 = dp.q0[0] + dp.q0[1] * r2 + dp.q0[2] * r4 + dp.q0[3] * r6 + dp.q0[4] * u + dp.q0[5] * v + dp.q0[6] * uv +
             dp.q0[7] * u2v + dp.q0[8] * uv2;
  float D1 = dp.q1[0] + dp.q1[1] * r2 + dp.q1[2] * r4 + dp.q1[3] * r6 + dp.q1[4] * u + dp.q1[5] * v + dp.q1[6] * uv +
             dp.q1[7] * u2v + dp.q1[8] * uv2;
  return (1.f + D1) * wd + D0;
}

// K25b: depthinvCorrectionKernel (:173-206) fused into undistortKernel on a point-filtered texture (:260-310): the texel
// under (xd, yd) is corrected on the fly
__global__ void __launch_bounds__(BX* BY) undistort_depthinv_kernel(ImgB src, ImgB dst, rgbid_intr intr, rgbid_depth_dist dp)
{
  int xu = blockIdx.x * blockDim.x + threadIdx.x, yu = blockIdx.y * blockDim.y + threadIdx.y;
  if (xu >= dst.cols || yu >= dst.rows) return;
  float res = qnanf(), xd, yd;
  if (undistort_source(xu, yu, intr, dst.cols, dst.rows, xd, yd)) {
    const int x = __float2int_rd(xd), y = __float2int_rd(yd);  // point filter: texel floor(coordinate)
    const int xs = x - dp.xshift, ys = y - dp.yshift;
    if ((xs > 0) && (ys > 0)) {
      float u = (__int2float_rn(x) - intr.cx) * (1.f / intr.fx);
      float v = (__int2float_rn(y) - intr.cy) * (1.f / intr.fy);
      res = correct_depthinv(u, v, __ldg(src.row(0, ys) + xs), dp);
    }
  }
  dst.row(0, yu)[xu] = res;
}

// K26a: depthinvRegistrationTranslationWithDilationKernel (warping_registration.cu:232-281) on a canvas cleared to 0
__global__ void __launch_bounds__(BX* BY) register_splat_kernel(ImgB src, int* __restrict__ canvas, size_t cpitch, int crows,
                                                                 int ccols, float3 t_dc_proj, int offset_x, int offset_y)
{
  int xd = blockIdx.x * blockDim.x + threadIdx.x, yd = blockIdx.y * blockDim.y + threadIdx.y;
  if (xd >= src.cols || yd >= src.rows) return;
  float wd = src.row(0, yd)[xd];
  if (isnan(wd)) return;
  // registerPixelTranslationOnly, :148-165
  float zd = 1.f / wd;
  float Xx = __int2float_rn(xd) * zd - t_dc_proj.x, Xy = __int2float_rn(yd) * zd - t_dc_proj.y, Xz = zd - t_dc_proj.z;
  float w_inter = 1.f / Xz;
  float xc = Xx * w_inter, yc = Xy * w_inter;
  if (w_inter > 0.01f) {
    float dilation = w_inter / wd;
    int iw = __float_as_int(w_inter);
    int xmin = __float2int_rn(xc - 0.5f * dilation) + offset_x, xmax = __float2int_rn(xc + 0.5f * dilation) + offset_x;
    int ymin = __float2int_rn(yc - 0.5f * dilation) + offset_y, ymax = __float2int_rn(yc + 0.5f * dilation) + offset_y;
    for (int x = max(0, xmin); x < min(xmax + 1, ccols); x++)
      for (int y = max(0, ymin); y < min(ymax + 1, crows); y++)
        atomicMax((int*)((char*)canvas + (size_t)y * cpitch) + x, iw);
  }
}

// K26b: conversionRegistrationKernel (:185-203) + homographyKernelInvDepthGridStride (:597-635)
__global__ void __launch_bounds__(BX* BY) register_homography_kernel(const int* __restrict__ canvas, size_t cpitch, int crows,
                                                                      int ccols, ImgB dst, Proj srcHdst, Proj dstHsrc,
                                                                      float offset_x, float offset_y)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.cols || y >= dst.rows) return;
  float out = qnanf();
  float px = __int2float_rn(x), py = __int2float_rn(y);
  float sx = srcHdst.r[0] * px + srcHdst.r[1] * py + srcHdst.r[2] * 1.f;
  float sy = srcHdst.r[3] * px + srcHdst.r[4] * py + srcHdst.r[5] * 1.f;
  float sz = srcHdst.r[6] * px + srcHdst.r[7] * py + srcHdst.r[8] * 1.f;
  float inv = 1.f / sz;
  sx *= inv; sy *= inv; sz *= inv;
  float x_src = sx + 0.5f + offset_x, y_src = sy + 0.5f + offset_y;
  int ix = __float2int_rd(x_src), iy = __float2int_rd(y_src);
  if (!(ix < 0 || iy < 0 || ix >= ccols || iy >= crows)) {
    int bits = __ldg((const int*)((const char*)canvas + (size_t)iy * cpitch) + ix);
    float w_src = bits != 0 ? __int_as_float(bits) : qnanf();
    float dz = dstHsrc.r[6] * sx + dstHsrc.r[7] * sy + dstHsrc.r[8] * sz;
    float res = w_src / dz;
    if (res > 0.f) out = res;
  }
  dst.row(0, y)[x] = out;
}

// K8: integrateWarpedRGBKernel (warping_registration.cu:672-712); gate DEPTHINV_INTEGR_TH = 0.0075 (:80)
__global__ void __launch_bounds__(BX* BY) integrate_rgb_kernel(ImgB dw, ImgB rw, ImgB gw, ImgB bw, ImgB ww, ImgB dd,
                                                                uint8_t* __restrict__ colors, size_t cpitch, ImgB wd)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dd.cols || y >= dd.rows) return;
  const float d_src = dw.row(0, y)[x], r = rw.row(0, y)[x], g = gw.row(0, y)[x], b = bw.row(0, y)[x];
  if (isnan(d_src) || isnan(r) || isnan(g) || isnan(b)) return;
  uint8_t* c = colors + (size_t)y * cpitch + 3 * (size_t)x;
  const float d_dst = dd.row(0, y)[x];
  if (isnan(d_dst)) {
    dd.row(0, y)[x] = d_src;
    c[0] = (uint8_t)__float2int_rn(r); c[1] = (uint8_t)__float2int_rn(g); c[2] = (uint8_t)__float2int_rn(b);
    wd.row(0, y)[x] = ww.row(0, y)[x];
  } else if (((d_dst - d_src) < 0.0075f) && ((d_src - d_dst) < 0.0075f)) {
    const float w_dst = wd.row(0, y)[x], w_src = ww.row(0, y)[x];
    const float new_weight = w_dst + w_src;
    dd.row(0, y)[x] = (d_dst * w_dst + d_src * w_src) / new_weight;
    c[0] = (uint8_t)__float2int_rn((__int2float_rn(c[0]) * w_dst + r * w_src) / new_weight);
    c[1] = (uint8_t)__float2int_rn((__int2float_rn(c[1]) * w_dst + g * w_src) / new_weight);
    c[2] = (uint8_t)__float2int_rn((__int2float_rn(c[2]) * w_dst + b * w_src) / new_weight);
    wd.row(0, y)[x] = new_weight;
  }
}

// K27: generateImageKernel / generateImageRGBKernel (image_generator.cu:60-181), one light source
__global__ void __launch_bounds__(BX* BY) generate_image_kernel(ImgB vmap, ImgB nmap, const uint8_t* __restrict__ rgb,
                                                                 size_t rgb_pitch, float3 light, uint8_t* __restrict__ out,
                                                                 size_t out_pitch, int rows, int cols)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  float vx = vmap.row(0, y)[x], nx = nmap.row(0, y)[x];
  uint8_t cr = 0, cg = 0, cb = 0;
  if (!isnan(vx) && !isnan(nx)) {
    float vy = vmap.row(0, y + rows)[x], vz = vmap.row(0, y + 2 * rows)[x];
    float ny = nmap.row(0, y + rows)[x], nz = nmap.row(0, y + 2 * rows)[x];
    float dx = light.x - vx, dy = light.y - vy, dz = light.z - vz;
    float rn = rsqrtf(dx * dx + dy * dy + dz * dz);  // normalized(): v * rsqrt(dot(v, v))
    float weight = 1.f * fabsf((dx * rn) * nx + (dy * rn) * ny + (dz * rn) * nz);
    int br = (int)(205 * weight) + 50;
    br = max(0, min(255, br));
    if (rgb == nullptr) {
      cr = cg = cb = (uint8_t)br;
    } else {
      const uint8_t* c = rgb + (size_t)y * rgb_pitch + 3 * (size_t)x;
      float br_f = __int2float_rn(br) / 255.f;
      cr = (uint8_t)__float2int_rn(__int2float_rn(c[0]) * br_f);
      cg = (uint8_t)__float2int_rn(__int2float_rn(c[1]) * br_f);
      cb = (uint8_t)__float2int_rn(__int2float_rn(c[2]) * br_f);
    }
  }
  uint8_t* o = out + (size_t)y * out_pitch + 3 * (size_t)x;
  o[0] = cr; o[1] = cg; o[2] = cb;
}

}  // namespace

void launch_undistort_intensity(const LaunchCtx& L, ImgB src, ImgB dst, const rgbid_intr& intr)
{
  undistort_intensity_kernel<<<grid2d(dst.cols, dst.rows), dim3(BX, BY), 0, L.stream>>>(src, dst, intr);
  ++*L.launches;
}

void launch_undistort_depthinv(const LaunchCtx& L, ImgB src, ImgB dst, const rgbid_intr& intr, const rgbid_depth_dist& dp)
{
  undistort_depthinv_kernel<<<grid2d(dst.cols, dst.rows), dim3(BX, BY), 0, L.stream>>>(src, dst, intr, dp);
  ++*L.launches;
}

void launch_register_depthinv(const LaunchCtx& L, ImgB src, ImgB dst, int* canvas, size_t cpitch, int crows, int ccols,
                              const float* dRc_proj, const float* t_dc_proj, const float* cRd_proj)
{
  const int offset_x = (ccols - src.cols) / 2, offset_y = (crows - src.rows) / 2;
  cudaMemsetAsync(canvas, 0, cpitch * crows, L.stream);  // initialiseRegistrationKernel: 0 = empty
  register_splat_kernel<<<grid2d(src.cols, src.rows), dim3(BX, BY), 0, L.stream>>>(
      src, canvas, cpitch, crows, ccols, make_float3(t_dc_proj[0], t_dc_proj[1], t_dc_proj[2]), offset_x, offset_y);
  Proj a, b;
  for (int i = 0; i < 9; ++i) { a.r[i] = dRc_proj[i]; b.r[i] = cRd_proj[i]; }
  a.t[0] = a.t[1] = a.t[2] = b.t[0] = b.t[1] = b.t[2] = 0.f;
  register_homography_kernel<<<grid2d(dst.cols, dst.rows), dim3(BX, BY), 0, L.stream>>>(canvas, cpitch, crows, ccols, dst, a, b,
                                                                                       (float)offset_x, (float)offset_y);
  *L.launches += 2;
}

void launch_integrate_rgb(const LaunchCtx& L, ImgB dw, ImgB rw, ImgB gw, ImgB bw, ImgB ww, ImgB dd, uint8_t* colors,
                          size_t cpitch, ImgB wd)
{
  integrate_rgb_kernel<<<grid2d(dd.cols, dd.rows), dim3(BX, BY), 0, L.stream>>>(dw, rw, gw, bw, ww, dd, colors, cpitch, wd);
  ++*L.launches;
}

void launch_generate_image(const LaunchCtx& L, ImgB vmap, ImgB nmap, const uint8_t* rgb, size_t rgb_pitch, const float* light,
                           uint8_t* out, size_t out_pitch, int rows, int cols)
{
  generate_image_kernel<<<grid2d(cols, rows), dim3(BX, BY), 0, L.stream>>>(vmap, nmap, rgb, rgb_pitch,
                                                                          make_float3(light[0], light[1], light[2]), out,
                                                                          out_pitch, rows, cols);
  ++*L.launches;
}

}  // namespace rgbid
