/*
 * ref_shim.cpp -- C-ABI driver around the UNMODIFIED reference device layer.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled by oracle/Makefile together with the
 * reference's own sources *where they lie* (/root/reference/src/cuda/*.cu and
 * ThirdParty/pcl_gpu_containers/src/*.cpp) into oracle/_ref/libref_oracle.so.  It contains no
 * reference code: it includes the reference's public header (src/internal.h) and calls the
 * RGBID_SLAM::device::* bridge functions exactly as src/visodo.cpp / src/keyframe_align.cpp do,
 * so that (a) tests can compare the new kernels with the reference's own kernels on a B200 and
 * (b) bench.py --impl reference can time the reference's own path.
 *
 * The host Gauss-Newton algebra (Eigen LLT / expMapRot, un-vendored) comes from oracle.c.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>

#include "internal.h" /* /root/reference/src/internal.h via -I */
#include "oracle.h"

namespace RGBID_SLAM { namespace device {
/* the application defines these globals (tools/RGBID_SLAMapp.cpp:68-69) */
cudaDeviceProp dev_prop;
int dev_id;
} }

using namespace RGBID_SLAM::device;
typedef DeviceArray2D<float> Map;

static inline Map wrap(const float* p, size_t pitch, int rows, int cols)
{
  return Map(rows, cols, (void*)p, pitch);
}
static inline Mat33 mat33(const float* R)
{
  Mat33 m;
  m.data[0] = make_float3(R[0], R[1], R[2]);
  m.data[1] = make_float3(R[3], R[4], R[5]);
  m.data[2] = make_float3(R[6], R[7], R[8]);
  return m;
}

extern "C" {

int ref_init(int device)
{
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  dev_id = device;
  if (cudaGetDeviceProperties(&dev_prop, device) != cudaSuccess) return 2;
  return 0;
}

void ref_convert_depth_to_invdepth(const unsigned short* src, size_t spitch, float* dst, size_t dpitch,
                                   int rows, int cols, float factor_depth)
{
  DeviceArray2D<unsigned short> s(rows, cols, (void*)src, spitch);
  Map d = wrap(dst, dpitch, rows, cols);
  convertDepth2InvDepth(s, d, factor_depth);
  sync();
}

void ref_compute_intensity(const unsigned char* rgb, size_t spitch, float* dst, size_t dpitch, int rows,
                           int cols)
{
  PtrStepSz<uchar3> s(rows, cols, (uchar3*)rgb, spitch);
  Map d = wrap(dst, dpitch, rows, cols);
  computeIntensity(s, d);
  sync();
}

float ref_pyr_down(const float* src, size_t spitch, int srows, int scols, float* dst, size_t dpitch)
{
  Map s = wrap(src, spitch, srows, scols);
  Map d = wrap(dst, dpitch, srows / 2, scols / 2);
  return pyrDownDepth(s, d);
}

float ref_gradient(const float* src, size_t spitch, int rows, int cols, float* gx, float* gy, size_t gpitch)
{
  Map s = wrap(src, spitch, rows, cols);
  Map x = wrap(gx, gpitch, rows, cols), y = wrap(gy, gpitch, rows, cols);
  return computeGradientDepth(s, x, y);
}

float ref_bilateral(const float* src, size_t spitch, int rows, int cols, float* dst, size_t dpitch,
                    float sigma_floatmap)
{
  Map s = wrap(src, spitch, rows, cols);
  Map d = wrap(dst, dpitch, rows, cols);
  return bilateralFilter(s, d, sigma_floatmap);
}

float ref_warp_invdepth(const float* src, const float* depth_prev, float* dst, size_t pitch, int rows,
                        int cols, const float* Rp, const float* tp)
{
  Map s = wrap(src, pitch, rows, cols), p = wrap(depth_prev, pitch, rows, cols), d = wrap(dst, pitch, rows, cols);
  return warpInvDepthWithTrafo3D(s, d, p, mat33(Rp), make_float3(tp[0], tp[1], tp[2]), Intr(1, 1, 0, 0));
}

float ref_warp_intensity(const float* src, const float* depth_prev, float* dst, size_t pitch, int rows,
                         int cols, const float* Rp, const float* tp)
{
  Map s = wrap(src, pitch, rows, cols), p = wrap(depth_prev, pitch, rows, cols), d = wrap(dst, pitch, rows, cols);
  return warpIntensityWithTrafo3DInvDepth(s, d, p, mat33(Rp), make_float3(tp[0], tp[1], tp[2]), Intr(1, 1, 0, 0));
}

float ref_warp_invdepth_weighted(const float* src, const float* depth_prev, float* dst, float* weight,
                                 size_t pitch, int rows, int cols, const float* Rp, const float* tp)
{
  Map s = wrap(src, pitch, rows, cols), p = wrap(depth_prev, pitch, rows, cols), d = wrap(dst, pitch, rows, cols);
  Map w = wrap(weight, pitch, rows, cols);
  return warpInvDepthWithTrafo3DWeighted(s, d, p, w, mat33(Rp), make_float3(tp[0], tp[1], tp[2]), Intr(1, 1, 0, 0));
}

/* ---- custom-calibration ingest, colour fusion, shaded previews ---------------------------- */
static inline Intr intr9(const float* k) { return Intr(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8]); }

float ref_undistort_intensity(const float* src, float* dst, size_t pitch, int rows, int cols, const float* intr)
{
  Map a = wrap(src, pitch, rows, cols), b = wrap(dst, pitch, rows, cols);
  return undistortIntensity(a, b, intr9(intr));
}

/* dp: c1 c0 q0[9] q1[9] as floats, then xshift yshift as ints */
float ref_undistort_depthinv(const float* src, float* scratch, float* dst, size_t pitch, int rows, int cols,
                             const float* intr, const float* dp, int xshift, int yshift)
{
  Map a = wrap(src, pitch, rows, cols), c = wrap(scratch, pitch, rows, cols), b = wrap(dst, pitch, rows, cols);
  DepthDist d(dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], dp[7], dp[8], dp[9], dp[10], dp[11], dp[12], dp[13], dp[14],
              dp[15], dp[16], dp[17], dp[18], dp[19], xshift, yshift);
  return undistortDepthInv(a, c, b, intr9(intr), d);
}

float ref_register_depthinv(const float* src, float* inter, int* inter_int, size_t ipitch, float* dst, size_t pitch,
                            int rows, int cols, const float* dRc_proj, const float* t_dc_proj, const float* cRd_proj)
{
  Map a = wrap(src, pitch, rows, cols), b = wrap(dst, pitch, rows, cols);
  Map im = wrap(inter, ipitch, 3 * rows, 3 * cols);
  DeviceArray2D<int> ii(3 * rows, 3 * cols, (void*)inter_int, ipitch);
  return registerDepthinv(a, im, ii, b, mat33(dRc_proj), make_float3(t_dc_proj[0], t_dc_proj[1], t_dc_proj[2]),
                          mat33(cRd_proj));
}

float ref_integrate_warped_rgb(const float* dw, const float* rw, const float* gw, const float* bw, const float* ww,
                               float* dd, unsigned char* colors, size_t cpitch, float* wd, size_t pitch, int rows, int cols)
{
  Map a = wrap(dw, pitch, rows, cols), r = wrap(rw, pitch, rows, cols), g = wrap(gw, pitch, rows, cols);
  Map b = wrap(bw, pitch, rows, cols), w = wrap(ww, pitch, rows, cols), d = wrap(dd, pitch, rows, cols);
  Map e = wrap(wd, pitch, rows, cols);
  return integrateWarpedRGB(a, r, g, b, w, d, PtrStepSz<uchar3>(rows, cols, (uchar3*)colors, cpitch), e);
}

void ref_generate_image(const float* vmap, const float* nmap, size_t pitch, const unsigned char* rgb, size_t rgb_pitch,
                        const float* light_pos, unsigned char* out, size_t out_pitch, int rows, int cols)
{
  Map v = wrap(vmap, pitch, 3 * rows, cols), n = wrap(nmap, pitch, 3 * rows, cols);
  LightSource light;
  light.number = 1;
  light.pos[0] = make_float3(light_pos[0], light_pos[1], light_pos[2]);
  PtrStepSz<uchar3> o(rows, cols, (uchar3*)out, out_pitch);
  if (rgb) generateImageRGB(v, n, PtrStepSz<uchar3>(rows, cols, (uchar3*)rgb, rgb_pitch), light, o);
  else generateImage(v, n, light, o);
  sync();
}

float ref_integrate_warped_frame(const float* wsrc, const float* wweight, float* dst, float* dweight,
                                 size_t pitch, int rows, int cols)
{
  Map a = wrap(wsrc, pitch, rows, cols), b = wrap(wweight, pitch, rows, cols);
  Map c = wrap(dst, pitch, rows, cols), d = wrap(dweight, pitch, rows, cols);
  return integrateWarpedFrame(a, b, c, d);
}

float ref_visibility_ratio(const float* depth_src, const float* depth_dst, size_t pitch, int rows, int cols,
                           const float* Rp, const float* tp, unsigned char* mask, size_t mpitch)
{
  Map a = wrap(depth_src, pitch, rows, cols), b = wrap(depth_dst, pitch, rows, cols);
  float ratio = -1.f;
  if (mask) {
    DeviceArray2D<unsigned char> m(rows, cols, (void*)mask, mpitch);
    getVisibilityRatioWithOverlapMask(a, b, mat33(Rp), make_float3(tp[0], tp[1], tp[2]), Intr(1, 1, 0, 0), ratio,
                                      0.0125f, m);
  } else
    getVisibilityRatio(a, b, mat33(Rp), make_float3(tp[0], tp[1], tp[2]), Intr(1, 1, 0, 0), ratio, 0.0125f);
  return ratio;
}

/* error must hold the sampled size given by orc_error_geometry */
int ref_compute_error(const float* im1, const float* im0, size_t pitch, int rows, int cols, int min_nsamples,
                      float* error)
{
  int kr, kc, s;
  orc_error_geometry(rows, cols, min_nsamples, &kr, &kc, &s);
  Map a = wrap(im1, pitch, rows, cols), b = wrap(im0, pitch, rows, cols);
  DeviceArray<float> e(error, (size_t)kr * kc);
  computeErrorGridStride(a, b, e, min_nsamples);
  return kr * kc;
}

void ref_sigma_nu_student(float* error, int n, float* bias, float* sigma, float* nu, int mest)
{
  DeviceArray<float> e(error, (size_t)n);
  computeSigmaAndNuStudent(e, *bias, *sigma, *nu, mest);
}

void ref_nu_student(float* error, int n, float bias, float sigma, float* nu)
{
  DeviceArray<float> e(error, (size_t)n);
  computeNuStudent(e, bias, sigma, *nu);
}

void ref_sigma_pdf(float* error, int n, float* bias, float* sigma, int mest)
{
  DeviceArray<float> e(error, (size_t)n);
  computeSigmaPdf(e, *bias, *sigma, mest);
}

void ref_chi_square(float* err_int, float* err_depth, int n, float sigma_int, float sigma_depth, int mest,
                    float* chi_squared, float* chi_test, float* ndof)
{
  DeviceArray<float> a(err_int, (size_t)n), b(err_depth, (size_t)n);
  computeChiSquare(a, b, sigma_int, sigma_depth, mest, *chi_squared, *chi_test, *ndof);
}

struct RefSysBuffers { DeviceArray2D<float_type> gbuf; DeviceArray<float_type> mbuf; };
static RefSysBuffers g_sys;

float ref_build_system(const float* W0, const float* I0, const float* gWx, const float* gWy, const float* gIx,
                       const float* gIy, const float* W1, const float* I1, size_t pitch, int rows, int cols,
                       const orc_system_params* P, double* A36, double* b6)
{
  Map w0 = wrap(W0, pitch, rows, cols), i0 = wrap(I0, pitch, rows, cols);
  Map gwx = wrap(gWx, pitch, rows, cols), gwy = wrap(gWy, pitch, rows, cols);
  Map gix = wrap(gIx, pitch, rows, cols), giy = wrap(gIy, pitch, rows, cols);
  Map w1 = wrap(W1, pitch, rows, cols), i1 = wrap(I1, pitch, rows, cols);
  float3 z = make_float3(0, 0, 0);
  Intr intr(P->fx, P->fy, P->cx, P->cy);
  if (P->student_nu)
    return buildSystemStudentNuGridStride(z, z, w0, i0, gwx, gwy, gix, giy, w1, i1, P->mestimator, P->weighting,
                                          P->sigma_depthinv, P->sigma_int, P->bias_depthinv, P->bias_int,
                                          P->nu_depthinv, P->nu_int, intr, 6, g_sys.gbuf, g_sys.mbuf, A36, b6);
  return buildSystemGridStride(z, z, w0, i0, gwx, gwy, gix, giy, w1, i1, P->mestimator, P->weighting,
                               P->sigma_depthinv, P->sigma_int, P->bias_depthinv, P->bias_int, intr, 6, g_sys.gbuf,
                               g_sys.mbuf, A36, b6);
}

void ref_vmap(const float* depth_inv, size_t pitch, int rows, int cols, float fx, float fy, float cx, float cy,
              float* vmap, size_t vpitch)
{
  Map d = wrap(depth_inv, pitch, rows, cols);
  Map v = wrap(vmap, vpitch, 3 * rows, cols);
  createVMap(Intr(fx, fy, cx, cy), d, v);
}

void ref_nmap_gradients(const float* depth_inv, const float* gx, const float* gy, size_t pitch, int rows, int cols,
                        float fx, float fy, float cx, float cy, float* nmap, size_t npitch)
{
  Map d = wrap(depth_inv, pitch, rows, cols), x = wrap(gx, pitch, rows, cols), y = wrap(gy, pitch, rows, cols);
  Map n = wrap(nmap, npitch, 3 * rows, cols);
  createNMapGradients(Intr(fx, fy, cx, cy), d, x, y, n);
}

/*
 * Coarse-to-fine loop of VisodoTracker::estimateVisualOdometry (src/visodo.cpp:1041-1415) /
 * KeyframeAlign::alignKeyframes (src/keyframe_align.cpp:178-350) driving the reference's own
 * bridge functions.  All pyramid pointers are DEVICE pointers with dense rows (pitch = cols*4).
 */
int ref_align(const orc_align_config* C, const orc_pyramids* P, double* R, double* t, double* cov36,
              orc_iter_trace* trace, int trace_cap, int* n_trace, orc_frame_stats* stats)
{
  int nt = 0, status = 0;
  double A[36], b[6];
  memset(A, 0, sizeof(A));
  int numSMs = (C->mode == ORC_MODE_ALIGN) ? 5 : -1; /* keyframe_align.cpp:39 */
  std::vector<Map> W1(C->levels), I1(C->levels);
  std::vector<DeviceArray<float> > eI(C->levels), eW(C->levels);
  for (int l = 0; l < C->levels; ++l) {
    W1[l].create(C->rows >> l, C->cols >> l);
    I1[l].create(C->rows >> l, C->cols >> l);
  }
  DeviceArray2D<float_type> gbuf; DeviceArray<float_type> mbuf;
  float3 z = make_float3(0, 0, 0);
  /* CHI_SQUARED termination, src/visodo.cpp:1134-1164 (see oracle.c for the pyrFirst remark) */
  const bool chi_term = (C->termination == ORC_TERM_CHI_SQUARED) && C->mode == ORC_MODE_TRACKER;
  float rmse_prev = 9999.f;
  double R_before[9], t_before[3];
  memcpy(R_before, R, sizeof(R_before)); memcpy(t_before, t, sizeof(t_before));
  DeviceArray<float> eI0, eW0;
  auto chi_test = [&](int iter) -> bool {
    Map Wkf0 = wrap(P->W_kf[0], (size_t)C->cols * sizeof(float), C->rows, C->cols);
    Map Ikf0 = wrap(P->I_kf[0], (size_t)C->cols * sizeof(float), C->rows, C->cols);
    computeErrorGridStride(I1[0], Ikf0, eI0);
    computeErrorGridStride(W1[0], Wkf0, eW0);
    float chi2 = 1.f, chit = 1.f, ndof = 1.f;
    computeChiSquare(eI0, eW0, 5.f, 0.0025f, C->mestimator, chi2, chit, ndof);
    float rmse = sqrt(chi2) / sqrt(ndof);
    if (iter != 1 && rmse > rmse_prev) return true;
    rmse_prev = rmse;
    return false;
  };

  for (int level = C->levels - 1; level >= C->finest_level && !status; --level) {
    int rows = C->rows >> level, cols = C->cols >> level;
    size_t pitch = (size_t)cols * sizeof(float);
    int div = 1 << level;
    Intr intr(C->fx / div, C->fy / div, C->cx / div, C->cy / div);
    Map Wkf = wrap(P->W_kf[level], pitch, rows, cols), Ikf = wrap(P->I_kf[level], pitch, rows, cols);
    Map gWx = wrap(P->gWx_kf[level], pitch, rows, cols), gWy = wrap(P->gWy_kf[level], pitch, rows, cols);
    Map gIx = wrap(P->gIx_kf[level], pitch, rows, cols), gIy = wrap(P->gIy_kf[level], pitch, rows, cols);
    Map Wc = wrap(P->W_cur[level], pitch, rows, cols), Ic = wrap(P->I_cur[level], pitch, rows, cols);
    for (int iter = 0; iter < C->iterations[level]; ++iter) {
      float Rp[9], tp[3];
      bool end_level = false;
      if (C->warp_first && C->mode == ORC_MODE_TRACKER && level > 0) {
        /* WARP_ORDER = warpFirst, src/visodo.cpp:1078-1105: warp at level 0, pyramid of the warped maps */
        size_t pitch0 = (size_t)C->cols * sizeof(float);
        Intr intr0(C->fx, C->fy, C->cx, C->cy);
        Map Wkf0 = wrap(P->W_kf[0], pitch0, C->rows, C->cols);
        Map Wc0 = wrap(P->W_cur[0], pitch0, C->rows, C->cols), Ic0 = wrap(P->I_cur[0], pitch0, C->rows, C->cols);
        orc_projective_inverse_pose(R, t, intr0.fx, intr0.fy, intr0.cx, intr0.cy, Rp, tp);
        Mat33 dR0 = mat33(Rp); float3 dt0 = make_float3(tp[0], tp[1], tp[2]);
        warpInvDepthWithTrafo3D(Wc0, W1[0], Wkf0, dR0, dt0, intr0);
        warpIntensityWithTrafo3DInvDepth(Ic0, I1[0], W1[0], dR0, dt0, intr0);
        if (chi_term && iter != 0) end_level = chi_test(iter);
        for (int i = 1; i <= level; ++i) {
          pyrDownIntensity(I1[i - 1], I1[i]);
          pyrDownDepth(W1[i - 1], W1[i]);
        }
      } else {
      orc_projective_inverse_pose(R, t, intr.fx, intr.fy, intr.cx, intr.cy, Rp, tp);
      Mat33 dR = mat33(Rp); float3 dt = make_float3(tp[0], tp[1], tp[2]);
      warpInvDepthWithTrafo3D(Wc, W1[level], Wkf, dR, dt, intr, numSMs);
      warpIntensityWithTrafo3DInvDepth(Ic, I1[level], C->mode == ORC_MODE_TRACKER ? W1[level] : Wkf, dR, dt, intr, numSMs);
      if (chi_term && iter != 0 && level == 0) end_level = chi_test(iter);
      }
      if (end_level) { memcpy(R, R_before, sizeof(R_before)); memcpy(t, t_before, sizeof(t_before)); break; }
      float sigma_int = 5.f, sigma_w = 0.0025f, bias_int = 0.f, bias_w = 0.f, nu_int = 5.f, nu_w = 5.f;
      if (C->mode == ORC_MODE_TRACKER) {
        if (C->sigma_estimator == ORC_SIGMA_PDF) {
          computeErrorGridStride(I1[level], Ikf, eI[level], C->nsamples);
          computeErrorGridStride(W1[level], Wkf, eW[level], C->nsamples);
          computeSigmaAndNuStudent(eI[level], bias_int, sigma_int, nu_int, C->mestimator);
          computeSigmaAndNuStudent(eW[level], bias_w, sigma_w, nu_w, C->mestimator);
          nu_int = std::max(nu_int, nu_w);
        }
      } else {
        computeErrorGridStride(W1[level], Wkf, eW[level], C->nsamples, numSMs);
        computeErrorGridStride(I1[level], Ikf, eI[level], C->nsamples, numSMs);
        computeNuStudent(eW[level], bias_w, sigma_w, nu_w, numSMs);
        computeNuStudent(eI[level], bias_int, sigma_int, nu_int, numSMs);
        nu_int = nu_w; /* keyframe_align.cpp:308 passes nu_depthinv twice */
      }
      buildSystemStudentNuGridStride(z, z, Wkf, Ikf, gWx, gWy, gIx, gIy, W1[level], I1[level],
                                     C->mode == ORC_MODE_TRACKER ? C->mestimator : (int)STUDENT,
                                     C->mode == ORC_MODE_TRACKER ? C->weighting : (int)INDEPENDENT,
                                     sigma_w, sigma_int, bias_w, bias_int, nu_w, nu_int, intr, 6, gbuf, mbuf, A, b, numSMs);
      double x[6];
      memcpy(R_before, R, sizeof(R_before)); memcpy(t_before, t, sizeof(t_before));
      int bad = orc_gn_update(A, b, R, t, x);
      if (trace && nt < trace_cap) {
        orc_iter_trace* T = &trace[nt];
        T->level = level; T->iter = iter;
        int shift = 0;
        for (int i = 0; i < 6; ++i) { for (int j = i; j < 6; ++j) T->sums27[shift++] = A[i * 6 + j]; T->sums27[shift++] = b[i]; }
        T->sigma_int = sigma_int; T->sigma_depthinv = sigma_w; T->bias_int = bias_int; T->bias_depthinv = bias_w;
        T->nu_int = nu_int; T->nu_depthinv = nu_w; T->irls_iters_int = -1; T->irls_iters_depthinv = -1;
        memcpy(T->x, x, sizeof(x)); memcpy(T->R, R, sizeof(double) * 9); memcpy(T->t, t, sizeof(double) * 3);
      }
      ++nt;
      if (bad) { status = 1; break; }
      if (C->termination == ORC_TERM_CONVERGENCE) {
        double n2 = 0.0;
        for (int k = 0; k < 6; ++k) n2 += x[k] * x[k];
        if (n2 < (double)C->conv_eps * (double)C->conv_eps) break;
      }
    }
  }
  if (status) {
    if (cov36) for (int i = 0; i < 36; ++i) cov36[i] = (i % 7 == 0) ? 100.0 : 0.0;
  } else if (C->mode == ORC_MODE_TRACKER) {
    int level = C->finest_level;
    int rows = C->rows >> level, cols = C->cols >> level;
    size_t pitch = (size_t)cols * sizeof(float);
    int div = 1 << level;
    Intr intr(C->fx / div, C->fy / div, C->cx / div, C->cy / div);
    Map Wkf = wrap(P->W_kf[level], pitch, rows, cols), Ikf = wrap(P->I_kf[level], pitch, rows, cols);
    Map gWx = wrap(P->gWx_cov[level], pitch, rows, cols), gWy = wrap(P->gWy_cov[level], pitch, rows, cols);
    Map gIx = wrap(P->gIx_cov[level], pitch, rows, cols), gIy = wrap(P->gIy_cov[level], pitch, rows, cols);
    Map Wc = wrap(P->W_cur[level], pitch, rows, cols), Ic = wrap(P->I_cur[level], pitch, rows, cols);
    float Rp[9], tp[3];
    orc_projective_inverse_pose(R, t, intr.fx, intr.fy, intr.cx, intr.cy, Rp, tp);
    Mat33 dR = mat33(Rp); float3 dt = make_float3(tp[0], tp[1], tp[2]);
    warpInvDepthWithTrafo3D(Wc, W1[level], Wkf, dR, dt, intr);
    warpIntensityWithTrafo3DInvDepth(Ic, I1[level], W1[level], dR, dt, intr);
    float sigma_int = expf(logf(5.f) - 0.f * logf(2.f)), sigma_w = expf(logf(0.0025f) - 0.f * logf(2.f));
    buildSystemGridStride(z, z, Wkf, Ikf, gWx, gWy, gIx, gIy, W1[level], I1[level], (int)STUDENT, C->weighting,
                          sigma_w, sigma_int, 0.f, 0.f, intr, 6, gbuf, mbuf, A, b);
    if (cov36) orc_inverse6(A, cov36);
    computeErrorGridStride(I1[level], Ikf, eI[level]);
    computeErrorGridStride(W1[level], Wkf, eW[level]);
    float chi2 = 0, chit = 0, ndof = 0;
    computeChiSquare(eI[level], eW[level], 5.f, 0.0025f, C->mestimator, chi2, chit, ndof);
    if (stats) {
      int shift = 0;
      for (int i = 0; i < 6; ++i) { for (int j = i; j < 6; ++j) stats->cov_sums27[shift++] = A[i * 6 + j]; stats->cov_sums27[shift++] = b[i]; }
      stats->chi_square = chi2; stats->chi_test = chit; stats->ndof = ndof;
    }
  } else {
    if (cov36) orc_inverse6(A, cov36);
  }
  if (n_trace) *n_trace = nt;
  return status;
}

} /* extern "C" */
