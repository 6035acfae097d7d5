// internal.hpp -- the reference's bridge-function API (RGBID_SLAM::device::*, src/internal.h:187-453) re-exposed
// header-only on top of the C ABI (include/rgbid_b200.h).  Same namespace, names, argument order and meaning, so
// src/visodo.cpp / src/keyframe_align.cpp compile against this header instead of src/internal.h and link
// librgbid_b200.so instead of the reference's CUDA objects.  See INTEGRATION.md.
//
// Differences from the reference, all deliberate:
//  * a CUDA failure throws std::runtime_error instead of exit(0);
//  * each host thread lazily owns one rgbid_ctx (device `dev_id`, its own stream); the reference uses the
//    per-thread default stream and creates / frees scratch inside every call;
//  * the trailing `numSMs` throttle is accepted and ignored (it tunes a 5-SM laptop GPU, src/keyframe_align.cpp:39);
//  * `gbuf` / `mbuf` of buildSystem* are accepted for source compatibility but not used (the reduction is one
//    launch with context-owned scratch).
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>

#include "../rgbid_b200.h"
#include "device_array.hpp"

using namespace pcl::gpu;

namespace RGBID_SLAM {
namespace device {

// The application defines and fills these (tools/RGBID_SLAMapp.cpp:68-69, 383-387); only dev_id is read here.
extern cudaDeviceProp dev_prop;
extern int dev_id;

typedef unsigned short ushort;
typedef unsigned char uchar;
typedef DeviceArray2D<float> MapArr;
typedef DeviceArray2D<ushort> DepthMap;
typedef DeviceArray2D<uchar> IntensityMap;
typedef DeviceArray2D<float> DepthMapf;
typedef DeviceArray2D<float> IntensityMapf;
typedef DeviceArray2D<float> GradientMap;
typedef DeviceArray2D<uchar> BinaryMap;
typedef double float_type;

enum { B_SIZE = 6, A_SIZE = (B_SIZE * B_SIZE - B_SIZE) / 2 + B_SIZE, TOTAL_SIZE = A_SIZE + B_SIZE };
enum { LSQ, HUBER, TUKEY, STUDENT };
enum { NO_MM, CONSTANT_VELOCITY };
enum { SIGMA_MAD, SIGMA_PDF, SIGMA_CONS };
enum { INDEPENDENT, MIN_WEIGHT, GEOM_ONLY, PHOT_ONLY };
enum { WARP_FIRST, PYR_FIRST };
enum { CHI_SQUARED, ALL_ITERS };
enum { NO_FILTERS, FILTER_GRADS };

const float THRESHOLD_HUBER = 1.345f;
const float THRESHOLD_TUKEY = 4.685f;
const float STUDENT_DOF = 5.f;
const float FOCAL_LENGTH = 543.78f;
const float CENTER_X = 313.45f;
const float CENTER_Y = 235.00f;
const int DEFAULT_MOTION_MODEL = CONSTANT_VELOCITY;
const int DEFAULT_MESTIMATOR = STUDENT;
const int DEFAULT_FINEST_LEVEL = 0;
const int DEFAULT_SIGMA = SIGMA_PDF;
const int DEFAULT_WEIGHTING = INDEPENDENT;
const int DEFAULT_WARPING = WARP_FIRST;
const int DEFAULT_ODO_KF_COUNT = 9999999;
const int DEFAULT_INTEGR_KF_COUNT = 9999999;
const float DEFAULT_VISRATIO_ODO = 0.9f;
const float DEFAULT_VISRATIO_INTEGR = 0.7f;
const int DEFAULT_TERMINATION = ALL_ITERS;
const int DEFAULT_IMAGE_FILTERING = NO_FILTERS;
const int DEFAULT_NSAMPLES = 10000;

/** Camera intrinsics (src/internal.h:119-140) */
struct Intr {
  float fx, fy, cx, cy, k1, k2, k3, k4, k5;
  Intr() {}
  Intr(float fx_, float fy_, float cx_, float cy_, float k1_ = 0.f, float k2_ = 0.f, float k3_ = 0.f, float k4_ = 0.f,
       float k5_ = 0.f)
      : fx(fx_), fy(fy_), cx(cx_), cy(cy_), k1(k1_), k2(k2_), k3(k3_), k4(k4_), k5(k5_) {}
  Intr operator()(int level_index) const
  {
    int div = 1 << level_index;
    return Intr(fx / div, fy / div, cx / div, cy / div, k1, k2, k3, k4, k5);
  }
};

/** 3x3 matrix for device code (src/internal.h:166-169): three float3 rows */
struct Mat33 {
  float3 data[3];
};

/** Depth distortion model (src/internal.h:142-161) */
struct DepthDist {
  float c1, c0;
  float q00, q01, q02, q03, q04, q05, q06, q07, q08;
  float q10, q11, q12, q13, q14, q15, q16, q17, q18;
  int xshift, yshift;
  DepthDist() {}
  DepthDist(float c1_, float c0_, float q00_ = 0.f, float q01_ = 0.f, float q02_ = 0.f, float q03_ = 0.f, float q04_ = 0.f,
            float q05_ = 0.f, float q06_ = 0.f, float q07_ = 0.f, float q08_ = 0.f, float q10_ = 1.f, float q11_ = 0.f,
            float q12_ = 0.f, float q13_ = 0.f, float q14_ = 0.f, float q15_ = 0.f, float q16_ = 0.f, float q17_ = 0.f,
            float q18_ = 0.f, int xshift_ = 4, int yshift_ = 4)
      : c1(c1_), c0(c0_), q00(q00_), q01(q01_), q02(q02_), q03(q03_), q04(q04_), q05(q05_), q06(q06_), q07(q07_), q08(q08_),
        q10(q10_), q11(q11_), q12(q12_), q13(q13_), q14(q14_), q15(q15_), q16(q16_), q17(q17_), q18(q18_), xshift(xshift_),
        yshift(yshift_) {}
};

/** Light source of the shaded previews (src/internal.h:173-177) */
struct LightSource {
  float3 pos[1];
  int number;
};

// ---- per-thread context ---------------------------------------------------------------------------------
struct ThreadContext {
  rgbid_ctx* ctx;
  cudaEvent_t e0, e1;
  ThreadContext() : ctx(nullptr)
  {
    int rc = rgbid_ctx_create(&ctx, dev_id, nullptr);
    if (rc != RGBID_OK) throw std::runtime_error(std::string("rgbid_ctx_create: ") + rgbid_status_string(rc));
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
  }
  ~ThreadContext()
  {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    rgbid_ctx_destroy(ctx);
  }
};

inline ThreadContext& thread_context()
{
  static thread_local ThreadContext tc;
  return tc;
}

inline void check(int rc, const char* what)
{
  if (rc != RGBID_OK) throw std::runtime_error(std::string(what) + ": " + rgbid_status_string(rc));
}

// Every bridge function of the reference returns its elapsed milliseconds (cudaTimer, src/cuda/device.hpp:83-106)
// and synchronises before returning.
struct CallTimer {
  ThreadContext& tc;
  CallTimer() : tc(thread_context()) { cudaEventRecord(tc.e0, (cudaStream_t)rgbid_ctx_stream(tc.ctx)); }
  float done()
  {
    cudaStream_t s = (cudaStream_t)rgbid_ctx_stream(tc.ctx);
    cudaEventRecord(tc.e1, s);
    check(rgbid_ctx_sync(tc.ctx), "sync");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, tc.e0, tc.e1);
    return ms;
  }
};

inline void to_arrays(const Mat33& R, const float3& t, float* Rp, float* tp)
{
  for (int r = 0; r < 3; ++r) { Rp[3 * r] = R.data[r].x; Rp[3 * r + 1] = R.data[r].y; Rp[3 * r + 2] = R.data[r].z; }
  tp[0] = t.x; tp[1] = t.y; tp[2] = t.z;
}

/** synchronizes CUDA execution (src/internal.h:456-457) */
inline void sync() { check(rgbid_ctx_sync(thread_context().ctx), "sync"); }

template <class D, class Matx> D& device_cast(Matx& matx) { return (*reinterpret_cast<D*>(matx.data())); }

// ---- image preparation -------------------------------------------------------------------------------------
inline float pyrDownDepth(DepthMapf& src, DepthMapf& dst, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  dst.create(src.rows() / 2, src.cols() / 2);
  check(rgbid_pyr_down(t.tc.ctx, src.ptr(), src.step(), src.rows(), src.cols(), dst.ptr(), dst.step()), "pyrDownDepth");
  return t.done();
}

inline float pyrDownIntensity(IntensityMapf& src, IntensityMapf& dst, int numSMs = -1) { return pyrDownDepth(src, dst, numSMs); }

inline void convertDepth2InvDepth(const DepthMap& src, DepthMapf& dst, float factor_depth)
{
  CallTimer t;
  dst.create(src.rows(), src.cols());
  check(rgbid_convert_depth_to_invdepth(t.tc.ctx, src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols(),
                                        factor_depth), "convertDepth2InvDepth");
  t.done();
}

inline void computeIntensity(const PtrStepSz<uchar3>& src, IntensityMapf& dst)
{
  CallTimer t;
  dst.create(src.rows, src.cols);
  check(rgbid_compute_intensity(t.tc.ctx, (const uint8_t*)src.data, src.step, dst.ptr(), dst.step(), src.rows, src.cols),
        "computeIntensity");
  t.done();
}

inline void decomposeRGBInChannels(const PtrStepSz<uchar3>& src, IntensityMapf& r, IntensityMapf& g, IntensityMapf& b)
{
  CallTimer t;
  r.create(src.rows, src.cols); g.create(src.rows, src.cols); b.create(src.rows, src.cols);
  if (r.step() != g.step() || r.step() != b.step()) throw std::runtime_error("decomposeRGBInChannels: pitch mismatch");
  check(rgbid_decompose_rgb(t.tc.ctx, (const uint8_t*)src.data, src.step, r.ptr(), g.ptr(), b.ptr(), r.step(), src.rows,
                            src.cols), "decomposeRGBInChannels");
  t.done();
}

inline float computeGradientDepth(const DepthMapf& src, GradientMap& dst_hor, GradientMap& dst_vert, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  dst_hor.create(src.rows(), src.cols()); dst_vert.create(src.rows(), src.cols());
  if (dst_hor.step() != dst_vert.step()) throw std::runtime_error("computeGradient: pitch mismatch");
  check(rgbid_compute_gradient(t.tc.ctx, src.ptr(), src.step(), src.rows(), src.cols(), dst_hor.ptr(), dst_vert.ptr(),
                               dst_hor.step()), "computeGradient");
  return t.done();
}

inline float computeGradientIntensity(const IntensityMapf& src, GradientMap& h, GradientMap& v, int numSMs = -1)
{
  return computeGradientDepth(src, h, v, numSMs);
}

inline float bilateralFilter(const DeviceArray2D<float>& src, DeviceArray2D<float>& dst, const float sigma_floatmap, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  dst.create(src.rows(), src.cols());
  check(rgbid_bilateral_filter(t.tc.ctx, src.ptr(), src.step(), src.rows(), src.cols(), dst.ptr(), dst.step(),
                               sigma_floatmap), "bilateralFilter");
  return t.done();
}

inline void copyImage(const DeviceArray2D<float>& src, DeviceArray2D<float>& dst)
{
  CallTimer t;
  dst.create(src.rows(), src.cols());
  check(rgbid_copy_image(t.tc.ctx, src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols()), "copyImage");
  t.done();
}

inline void copyImages(const DepthMapf& src_depth, const IntensityMapf& src_int, DepthMapf& dst_depth, IntensityMapf& dst_int)
{
  copyImage(src_depth, dst_depth);
  copyImage(src_int, dst_int);
}

inline void initialiseWeightKeyframe(const DepthMapf& src_depth, DeviceArray2D<float>& dst_weight)
{
  CallTimer t;
  dst_weight.create(src_depth.rows(), src_depth.cols());
  check(rgbid_fill_image(t.tc.ctx, dst_weight.ptr(), dst_weight.step(), dst_weight.rows(), dst_weight.cols(), 1.f),
        "initialiseWeightKeyframe");
  t.done();
}

/** showGPUMemoryUsage (src/internal.h:181, src/cuda/misc.cu:526-540): one diagnostic line on stdout */
inline void showGPUMemoryUsage()
{
  size_t free_bytes = 0, total_bytes = 0;
  cudaError_t e = cudaMemGetInfo(&free_bytes, &total_bytes);
  if (e != cudaSuccess) throw std::runtime_error(std::string("showGPUMemoryUsage: ") + cudaGetErrorString(e));
  const double mb = 1024.0 * 1024.0;
  std::cout << "GPU memory usage: used =  " << (double)(total_bytes - free_bytes) / mb << " MB, free = "
            << (double)free_bytes / mb << " MB, total = " << (double)total_bytes / mb << std::endl;
}

/** initialiseDeviceMemory2D<T> (src/internal.h:273-274, src/cuda/misc.cu:327-341, 491-512: instantiated for
 *  unsigned char, unsigned int, char, int and float; the live call clears the overlap mask, src/visodo.cpp:2027).
 *  A constant fill needs no kernel of its own: bytes go through cudaMemset2DAsync, 32-bit values through the float
 *  fill with the value's bit pattern. */
template <typename T>
inline void initialiseDeviceMemory2D(DeviceArray2D<T>& src, T val, int numSMs = -1)
{
  (void)numSMs;
  static_assert(sizeof(T) == 1 || sizeof(T) == 4, "initialiseDeviceMemory2D: the reference instantiates 8- and 32-bit types");
  CallTimer t;
  if (sizeof(T) == 1) {
    unsigned char byte;
    std::memcpy(&byte, &val, 1);
    cudaError_t e = cudaMemset2DAsync(src.ptr(), src.step(), byte, (size_t)src.cols(), (size_t)src.rows(),
                                      (cudaStream_t)rgbid_ctx_stream(t.tc.ctx));
    if (e != cudaSuccess) throw std::runtime_error(std::string("initialiseDeviceMemory2D: ") + cudaGetErrorString(e));
  } else {
    float bits;
    std::memcpy(&bits, &val, 4);
    check(rgbid_fill_image(t.tc.ctx, reinterpret_cast<float*>(src.ptr()), src.step(), src.rows(), src.cols(), bits),
          "initialiseDeviceMemory2D");
  }
  t.done();
}

inline void createVMap(const Intr& intr, const DepthMapf& depth, MapArr& vmap, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  vmap.create(depth.rows() * 3, depth.cols());
  check(rgbid_create_vmap(t.tc.ctx, depth.ptr(), depth.step(), depth.rows(), depth.cols(), intr.fx, intr.fy, intr.cx,
                          intr.cy, vmap.ptr(), vmap.step()), "createVMap");
  t.done();
}

inline void createNMapGradients(const Intr& intr, const DepthMapf& depth_inv, const GradientMap& grad_x,
                                const GradientMap& grad_y, MapArr& nmap, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  nmap.create(depth_inv.rows() * 3, depth_inv.cols());
  if (depth_inv.step() != grad_x.step() || depth_inv.step() != grad_y.step()) throw std::runtime_error("createNMapGradients: pitch mismatch");
  check(rgbid_create_nmap_gradients(t.tc.ctx, depth_inv.ptr(), grad_x.ptr(), grad_y.ptr(), depth_inv.step(),
                                    depth_inv.rows(), depth_inv.cols(), intr.fx, intr.fy, intr.cx, intr.cy, nmap.ptr(),
                                    nmap.step()), "createNMapGradients");
  t.done();
}

// ---- warping / visibility / fusion -----------------------------------------------------------------------------
inline float warpInvDepthWithTrafo3D(DepthMapf& src, DepthMapf& dst, const DepthMapf& depth_prev, Mat33 inv_rotation,
                                     float3 inv_translation, const Intr& intr, int numSMs = -1)
{
  (void)intr; (void)numSMs;
  CallTimer t;
  float Rp[9], tp[3];
  to_arrays(inv_rotation, inv_translation, Rp, tp);
  check(rgbid_warp_invdepth(t.tc.ctx, src.ptr(), src.step(), depth_prev.ptr(), depth_prev.step(), dst.ptr(), dst.step(),
                            dst.rows(), dst.cols(), Rp, tp), "warpInvDepthWithTrafo3D");
  return t.done();
}

inline float warpIntensityWithTrafo3DInvDepth(IntensityMapf& src, IntensityMapf& dst, const DepthMapf& depthinv_prev,
                                              Mat33 inv_rotation, float3 inv_translation, const Intr& intr, int numSMs = -1)
{
  (void)intr; (void)numSMs;
  CallTimer t;
  float Rp[9], tp[3];
  to_arrays(inv_rotation, inv_translation, Rp, tp);
  check(rgbid_warp_intensity(t.tc.ctx, src.ptr(), src.step(), depthinv_prev.ptr(), depthinv_prev.step(), dst.ptr(),
                             dst.step(), dst.rows(), dst.cols(), Rp, tp), "warpIntensityWithTrafo3DInvDepth");
  return t.done();
}

inline float warpInvDepthWithTrafo3DWeighted(DepthMapf& src, DepthMapf& dst, const DepthMapf& depth_prev,
                                             DeviceArray2D<float>& weight_warped, Mat33 inv_rotation_proj,
                                             float3 inv_translation_proj, const Intr& intr, int numSMs = -1)
{
  (void)intr; (void)numSMs;
  CallTimer t;
  float Rp[9], tp[3];
  to_arrays(inv_rotation_proj, inv_translation_proj, Rp, tp);
  check(rgbid_warp_invdepth_weighted(t.tc.ctx, src.ptr(), src.step(), depth_prev.ptr(), depth_prev.step(), dst.ptr(),
                                     dst.step(), weight_warped.ptr(), weight_warped.step(), dst.rows(), dst.cols(), Rp, tp),
        "warpInvDepthWithTrafo3DWeighted");
  return t.done();
}

inline float integrateWarpedFrame(const DepthMapf& warped_depth_src, const DeviceArray2D<float>& warped_weight_src,
                                  DepthMapf& depth_dst, DeviceArray2D<float>& weight_dst, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  check(rgbid_integrate_warped_frame(t.tc.ctx, warped_depth_src.ptr(), warped_depth_src.step(), warped_weight_src.ptr(),
                                     warped_weight_src.step(), depth_dst.ptr(), depth_dst.step(), weight_dst.ptr(),
                                     weight_dst.step(), depth_dst.rows(), depth_dst.cols()), "integrateWarpedFrame");
  return t.done();
}

// ---- custom-calibration ingest (src/internal.h:354, 437, 440), colour fusion, previews (:416-421) ---------------------
inline rgbid_intr to_c(const Intr& i) { rgbid_intr r = {i.fx, i.fy, i.cx, i.cy, i.k1, i.k2, i.k3, i.k4, i.k5}; return r; }
inline rgbid_depth_dist to_c(const DepthDist& d)
{
  rgbid_depth_dist r = {d.c1, d.c0, {d.q00, d.q01, d.q02, d.q03, d.q04, d.q05, d.q06, d.q07, d.q08},
                        {d.q10, d.q11, d.q12, d.q13, d.q14, d.q15, d.q16, d.q17, d.q18}, d.xshift, d.yshift};
  return r;
}

inline float undistortIntensity(IntensityMapf& src, IntensityMapf& dst, const Intr& intr_int, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  rgbid_intr i = to_c(intr_int);
  check(rgbid_undistort_intensity(t.tc.ctx, src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols(), &i),
        "undistortIntensity");
  return t.done();
}

inline float undistortDepthInv(const DepthMapf& src, DepthMapf& src_corr, DepthMapf& dst, const Intr& intr_depth,
                               const DepthDist& dp, int numSMs = -1)
{
  (void)numSMs; (void)src_corr;  // the corrected intermediate map is never materialised (csrc/calib_ops.cu)
  CallTimer t;
  rgbid_intr i = to_c(intr_depth);
  rgbid_depth_dist d = to_c(dp);
  check(rgbid_undistort_depthinv(t.tc.ctx, src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols(), &i, &d),
        "undistortDepthInv");
  return t.done();
}

inline float registerDepthinv(const DepthMapf& src, DepthMapf& intermediate, DeviceArray2D<int>& intermediate_as_int,
                              DepthMapf& dst, const Mat33 dRc_proj, float3 t_dc_proj, const Mat33 cRd_proj, int numSMs = -1)
{
  (void)numSMs; (void)intermediate; (void)intermediate_as_int;  // the canvas lives in the context's scratch
  CallTimer t;
  float a[9], b[9], tt[3], zero[3];
  to_arrays(dRc_proj, t_dc_proj, a, tt);
  to_arrays(cRd_proj, t_dc_proj, b, zero);
  check(rgbid_register_depthinv(t.tc.ctx, src.ptr(), src.step(), dst.ptr(), dst.step(), src.rows(), src.cols(), a, tt, b),
        "registerDepthinv");
  return t.done();
}

inline float integrateWarpedRGB(const DepthMapf& depth_warped_src, const IntensityMapf& r_warped_src,
                                const IntensityMapf& g_warped_src, const IntensityMapf& b_warped_src,
                                const DeviceArray2D<float>& weight_warped_src, DepthMapf& depth_dst, PtrStepSz<uchar3> colors_dst,
                                DeviceArray2D<float>& weight_dst, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  const size_t p = depth_warped_src.step();
  if (r_warped_src.step() != p || g_warped_src.step() != p || b_warped_src.step() != p || weight_warped_src.step() != p ||
      depth_dst.step() != p || weight_dst.step() != p)
    throw std::runtime_error("integrateWarpedRGB: pitch mismatch");
  check(rgbid_integrate_warped_rgb(t.tc.ctx, depth_warped_src.ptr(), r_warped_src.ptr(), g_warped_src.ptr(), b_warped_src.ptr(),
                                   weight_warped_src.ptr(), depth_dst.ptr(), (uint8_t*)colors_dst.data, colors_dst.step,
                                   weight_dst.ptr(), p, depth_dst.rows(), depth_dst.cols()), "integrateWarpedRGB");
  return t.done();
}

inline void generateImage(const MapArr& vmap, const MapArr& nmap, const LightSource& light, PtrStepSz<uchar3> dst)
{
  CallTimer t;
  const float l[3] = {light.pos[0].x, light.pos[0].y, light.pos[0].z};
  check(rgbid_generate_image(t.tc.ctx, vmap.ptr(), nmap.ptr(), vmap.step(), nullptr, 0, l, (uint8_t*)dst.data, dst.step,
                             dst.rows, dst.cols), "generateImage");
  t.done();
}

inline void generateImageRGB(const MapArr& vmap, const MapArr& nmap, const PtrStepSz<uchar3>& rgb, const LightSource& light,
                             PtrStepSz<uchar3> dst)
{
  CallTimer t;
  const float l[3] = {light.pos[0].x, light.pos[0].y, light.pos[0].z};
  check(rgbid_generate_image(t.tc.ctx, vmap.ptr(), nmap.ptr(), vmap.step(), (const uint8_t*)rgb.data, rgb.step, l,
                             (uint8_t*)dst.data, dst.step, dst.rows, dst.cols), "generateImageRGB");
  t.done();
}

inline float getVisibilityRatio(const DepthMapf& depth_src, const DepthMapf& depth_dst, Mat33 rotation, float3 translation,
                                const Intr& intr, float& visibility_ratio, float geom_tol, int numSMs = -1)
{
  (void)intr; (void)geom_tol; (void)numSMs;  // geom_tol is ignored by the reference as well (0.020 hard-coded)
  CallTimer t;
  float Rp[9], tp[3];
  to_arrays(rotation, translation, Rp, tp);
  check(rgbid_visibility_ratio(t.tc.ctx, depth_src.ptr(), depth_src.step(), depth_dst.ptr(), depth_dst.step(),
                               depth_src.rows(), depth_src.cols(), Rp, tp, nullptr, 0, &visibility_ratio), "getVisibilityRatio");
  return t.done();
}

inline float getVisibilityRatioWithOverlapMask(const DepthMapf& depth_src, const DepthMapf& depth_dst, Mat33 rotation,
                                               float3 translation, const Intr& intr, float& visibility_ratio,
                                               float geom_tol, BinaryMap& overlap_mask, int numSMs = -1)
{
  (void)intr; (void)geom_tol; (void)numSMs;
  CallTimer t;
  float Rp[9], tp[3];
  to_arrays(rotation, translation, Rp, tp);
  check(rgbid_visibility_ratio(t.tc.ctx, depth_src.ptr(), depth_src.step(), depth_dst.ptr(), depth_dst.step(),
                               depth_src.rows(), depth_src.cols(), Rp, tp, overlap_mask.ptr(), overlap_mask.step(),
                               &visibility_ratio), "getVisibilityRatioWithOverlapMask");
  return t.done();
}

// ---- residual sampling / scale / chi-square ----------------------------------------------------------------------
inline float computeErrorGridStride(const DeviceArray2D<float>& im1, const DeviceArray2D<float>& im0, DeviceArray<float>& error,
                                    int Nsamples = 9999999, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  int kr, kc, s;
  rgbid_error_geometry(im0.rows(), im0.cols(), Nsamples, &kr, &kc, &s);
  error.create((size_t)kr * kc);
  check(rgbid_compute_error(t.tc.ctx, im1.ptr(), im1.step(), im0.ptr(), im0.step(), im0.rows(), im0.cols(), Nsamples,
                            error.ptr(), nullptr), "computeErrorGridStride");
  return t.done();
}

inline float computeSigmaAndNuStudent(DeviceArray<float>& error, float& bias, float& sigma, float& nu, int Mestimator, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  check(rgbid_sigma_nu_student(t.tc.ctx, error.ptr(), (int)error.size(), &bias, &sigma, &nu, Mestimator), "computeSigmaAndNuStudent");
  return t.done();
}

inline float computeNuStudent(DeviceArray<float>& error, float& bias, float& sigma, float& nu, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  check(rgbid_nu_student(t.tc.ctx, error.ptr(), (int)error.size(), bias, sigma, &nu), "computeNuStudent");
  return t.done();
}

inline float computeSigmaPdf(DeviceArray<float>& error, float& bias, float& sigma, int Mestimator, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  check(rgbid_sigma_pdf(t.tc.ctx, error.ptr(), (int)error.size(), &bias, &sigma, Mestimator), "computeSigmaPdf");
  return t.done();
}

inline float computeChiSquare(DeviceArray<float>& error_int, DeviceArray<float>& error_depth, float sigma_int, float sigma_depth,
                              int Mestimator, float& chi_square, float& chi_test, float& Ndof, int numSMs = -1)
{
  (void)numSMs;
  CallTimer t;
  check(rgbid_chi_square(t.tc.ctx, error_int.ptr(), error_depth.ptr(), (int)error_int.size(), sigma_int, sigma_depth,
                         Mestimator, &chi_square, &chi_test, &Ndof), "computeChiSquare");
  return t.done();
}

// ---- normal equations -----------------------------------------------------------------------------------------
inline float build_system_common(const DepthMapf& W0, const IntensityMapf& I0, const GradientMap& gradW0_x,
                                 const GradientMap& gradW0_y, const GradientMap& gradI0_x, const GradientMap& gradI0_y,
                                 const DepthMapf& W1, const IntensityMapf& I1, const rgbid_system_params& p,
                                 float_type* matrixA_host, float_type* vectorB_host)
{
  CallTimer t;
  const size_t pitch8[8] = {W0.step(), I0.step(), gradW0_x.step(), gradW0_y.step(), gradI0_x.step(), gradI0_y.step(),
                            W1.step(), I1.step()};  // every PtrStep carries its own step
  check(rgbid_build_system_pitched(t.tc.ctx, W0.ptr(), I0.ptr(), gradW0_x.ptr(), gradW0_y.ptr(), gradI0_x.ptr(), gradI0_y.ptr(),
                                   W1.ptr(), I1.ptr(), pitch8, W0.rows(), W0.cols(), &p, matrixA_host, vectorB_host), "buildSystem");
  return t.done();
}

inline float buildSystemGridStride(const float3 delta_trans, const float3 delta_rot, const DepthMapf& W0, const IntensityMapf& I0,
                                   const GradientMap& gradW0_x, const GradientMap& gradW0_y, const GradientMap& gradI0_x,
                                   const GradientMap& gradI0_y, const DepthMapf& W1, const IntensityMapf& I1, int Mestimator,
                                   int weighting, float sigma_depth, float sigma_int, float bias_depth, float bias_int,
                                   const Intr& intr, const int size_A, DeviceArray2D<float_type>& gbuf,
                                   DeviceArray<float_type>& mbuf, float_type* matrixA_host, float_type* vectorB_host,
                                   int numSMs = -1)
{
  (void)delta_trans; (void)delta_rot; (void)size_A; (void)gbuf; (void)mbuf; (void)numSMs;  // dead in the reference too
  rgbid_system_params p = {intr.fx, intr.fy, intr.cx, intr.cy, Mestimator, weighting, 0,
                           sigma_depth, sigma_int, bias_depth, bias_int, 5.f, 5.f};
  return build_system_common(W0, I0, gradW0_x, gradW0_y, gradI0_x, gradI0_y, W1, I1, p, matrixA_host, vectorB_host);
}

inline float buildSystemStudentNuGridStride(const float3 delta_trans, const float3 delta_rot, const DepthMapf& W0,
                                            const IntensityMapf& I0, const GradientMap& gradW0_x, const GradientMap& gradW0_y,
                                            const GradientMap& gradI0_x, const GradientMap& gradI0_y, const DepthMapf& W1,
                                            const IntensityMapf& I1, int Mestimator, int weighting, float sigma_depth,
                                            float sigma_int, float bias_depth, float bias_int, float nu_depth, float nu_int,
                                            const Intr& intr, const int size_A, DeviceArray2D<float_type>& gbuf,
                                            DeviceArray<float_type>& mbuf, float_type* matrixA_host, float_type* vectorB_host,
                                            int numSMs = -1)
{
  (void)delta_trans; (void)delta_rot; (void)size_A; (void)gbuf; (void)mbuf; (void)numSMs;
  rgbid_system_params p = {intr.fx, intr.fy, intr.cx, intr.cy, Mestimator, weighting, 1,
                           sigma_depth, sigma_int, bias_depth, bias_int, nu_depth, nu_int};
  return build_system_common(W0, I0, gradW0_x, gradW0_y, gradI0_x, gradI0_y, W1, I1, p, matrixA_host, vectorB_host);
}

}  // namespace device
}  // namespace RGBID_SLAM
