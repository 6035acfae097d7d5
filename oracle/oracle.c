/*
 * oracle.c -- CPU restatement of RGBiD-SLAM's dense frame-to-keyframe alignment path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rgbid-slam_b200/ may link, import or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs use it, and
 * only as the checker or the reported CPU baseline -- never as the product path.
 *
 * Every function restates one reference kernel / host routine (cited file:line, paths relative
 * to /root/reference).  Per-pixel arithmetic is float32 in the reference's operation order
 * (IEEE division/sqrt here; the reference's nvcc flags use approximate div/sqrt, a <= 2 ulp
 * difference per operation); reductions are accumulated in double, which is at least as
 * accurate as the reference's float block trees.
 *
 * Parity pinning: the reference has NO tests or golden vectors.  This
 * restatement is pinned against the reference's own CUDA kernels compiled verbatim
 * (oracle/_ref/libref_oracle.so, built by oracle/Makefile from /root/reference) and run on a
 * B200; the committed fixtures under tests/golden/ were produced by that library with
 * tests/golden/make_golden.py.
 *
 * Images are dense row-major float32 [rows][cols]; NaN marks an invalid pixel.
 */
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* src/internal.h:66-76 */
#define THRESHOLD_HUBER 1.345f
#define THRESHOLD_TUKEY 4.685f
#define STUDENT_DOF 5.f

static int g_tex_frac_mode = ORC_TEX_FRAC_ROUND;

void orc_set_tex_frac_mode(int mode) { g_tex_frac_mode = mode; }

/* Thread count of the row-parallel loops (bench.py reports the baseline on all host cores and on one) */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline float qnan(void) { return nanf(""); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------- */
/* Image preparation                                                                            */
/* ------------------------------------------------------------------------------------------- */

/* depth2invDepthKernel, src/cuda/misc.cu:105-124 */
void orc_depth_to_invdepth(const uint16_t* src, float* dst, int rows, int cols, float factor_depth)
{
  for (long i = 0; i < (long)rows * cols; ++i) {
    int value = src[i];
    float r = qnan();
    if (value > 0) {
      int c = imax(0, imin(value, 10000));
      r = (1.f / factor_depth) * 1000.f / (float)c;
    }
    dst[i] = r;
  }
}

/* intensityKernel, src/cuda/misc.cu:128-147 */
void orc_intensity(const uint8_t* rgb, float* dst, int rows, int cols)
{
  for (long i = 0; i < (long)rows * cols; ++i) {
    float r = (float)rgb[3 * i + 0], g = (float)rgb[3 * i + 1], b = (float)rgb[3 * i + 2];
    float v = 0.2126f * r + 0.7152f * g + 0.0722f * b;
    dst[i] = fmaxf(0.f, fminf(v, 255.f));
  }
}

/* decomposeRGBKernel, src/cuda/misc.cu:151-172 */
void orc_decompose_rgb(const uint8_t* rgb, float* r, float* g, float* b, int rows, int cols)
{
  for (long i = 0; i < (long)rows * cols; ++i) {
    r[i] = (float)rgb[3 * i + 0];
    g[i] = (float)rgb[3 * i + 1];
    b[i] = (float)rgb[3 * i + 2];
  }
}

/* pyrDownKernelGridStridef, src/cuda/pyrdown.cu:84-132 (radius 2, sigma 1, valid iff count > 12) */
void orc_pyr_down(const float* src, int srows, int scols, float* dst)
{
  const int R = 2;
  int drows = srows / 2, dcols = scols / 2;
  const float sigma = 1.f;
  const float s2ih = 0.5f / (sigma * sigma);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < drows; ++y)
    for (int x = 0; x < dcols; ++x) {
      int tx = imin(2 * x + R + 1, scols);
      int ty = imin(2 * y + R + 1, srows);
      float sum1 = 0.f, sum2 = 0.f;
      int count = 0;
      for (int cy = imax(0, 2 * y - R); cy < ty; ++cy)
        for (int cx = imax(0, 2 * x - R); cx < tx; ++cx) {
          float val = src[(long)cy * scols + cx];
          if (!isnan(val)) {
            float space2 = (float)((2 * x - cx) * (2 * x - cx) + (2 * y - cy) * (2 * y - cy));
            float weight = expf(-(space2 * s2ih));
            sum1 += val * weight;
            sum2 += weight;
            ++count;
          }
        }
      int d = 2 * R + 1;
      float res = qnan();
      if (count > (d * d) / 2) res = sum1 / sum2;
      dst[(long)y * dcols + x] = res;
    }
}

/* gradientKernel, src/cuda/misc.cu:176-220 (Sobel / 8, clamp to edge, NaN propagates) */
void orc_gradient(const float* src, int rows, int cols, float* gx, float* gy)
{
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float rh = 0.f, rv = 0.f;
      for (int dx = -1; dx < 2; ++dx)
        for (int dy = -1; dy < 2; ++dy) {
          int cx = imin(imax(0, x + dx), cols - 1);
          int cy = imin(imax(0, y + dy), rows - 1);
          int wh = dx * (2 - dy * dy);
          int wv = dy * (2 - dx * dx);
          float t = src[(long)cy * cols + cx];
          rh += t * (float)wh;
          rv += t * (float)wv;
        }
      gx[(long)y * cols + x] = rh / 8.f;
      gy[(long)y * cols + x] = rv / 8.f;
    }
}

/* bilateralKernel, src/cuda/filters.cu:86-135 (radius 2, sigma_space 5; the range term is
 * computed in double in the reference because of the 0.5 literals, then cast to float) */
void orc_bilateral(const float* src, int rows, int cols, float* dst, float sigma_floatmap)
{
  const int R = 2;
  const float sigma_space = 5.f;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float value = src[(long)y * cols + x];
      if (isnan(value)) { dst[(long)y * cols + x] = qnan(); continue; }
      int tx = imin(x + R + 1, cols);
      int ty = imin(y + R + 1, rows);
      float sum1 = 0.f, sum2 = 0.f;
      for (int cy = imax(y - R, 0); cy < ty; ++cy)
        for (int cx = imax(x - R, 0); cx < tx; ++cx) {
          float tmp = src[(long)cy * cols + cx];
          if (!isnan(tmp)) {
            float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
            float fn = (value - tmp) / sigma_floatmap;
            float s2ih = (float)(0.5 / (double)(sigma_space * sigma_space));
            float arg = (float)(-((double)(s2ih * space2) + 0.5 * (double)fn * (double)fn));
            float weight = expf(arg);
            sum1 += tmp * weight;
            sum2 += weight;
          }
        }
      dst[(long)y * cols + x] = sum1 / sum2;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Warping (src/cuda/warping_registration.cu)                                                   */
/* ------------------------------------------------------------------------------------------- */

/* registerPixel, warping_registration.cu:129-146.  Rp row-major 3x3, tp 3 (pixel-space, i.e.
 * K R K^-1 and K t).  Returns inverse depth in the source camera. */
static inline float register_pixel(float* xc, float* yc, int xd, int yd, float wd, const float* Rp,
                                   const float* tp)
{
  float zd = 1.f / wd;
  float X = (float)xd * zd, Y = (float)yd * zd, Z = zd;
  float Xc = (Rp[0] * X + Rp[1] * Y + Rp[2] * Z) + tp[0];
  float Yc = (Rp[3] * X + Rp[4] * Y + Rp[5] * Z) + tp[1];
  float Zc = (Rp[6] * X + Rp[7] * Y + Rp[8] * Z) + tp[2];
  float wc = 1.f / Zc;
  *xc = Xc * wc;
  *yc = Yc * wc;
  return wc;
}

/* cudaFilterModePoint fetch at unnormalised (x, y), clamp addressing: texel floor(x). */
static inline float tex_point(const float* img, int rows, int cols, float x, float y)
{
  int ix = imin(imax((int)floorf(x), 0), cols - 1);
  int iy = imin(imax((int)floorf(y), 0), rows - 1);
  return img[(long)iy * cols + ix];
}

/* cudaFilterModeLinear fetch at unnormalised (x, y), clamp addressing
 * (warping_registration.cu:943, fetch :493).  CUDA: xB = x - 0.5, i = floor(xB), alpha = frac(xB) in 1.8
 * fixed point.  Measured on B200 with tests/cuda/tex_probe2.cu (all 257^2 weight pairs, zero mismatches):
 * ka = round(alpha*256), kb = round(beta*256) and the FOUR weights are 8-bit too:
 *   w11 = (ka*kb + 128) >> 8, w10 = ka - w11, w01 = kb - w11, w00 = 256 - ka - kb + w11   (all / 256). */
static inline float tex_linear(const float* img, int rows, int cols, float x, float y)
{
  float xB = x - 0.5f, yB = y - 0.5f;
  float fx = floorf(xB), fy = floorf(yB);
  float a = xB - fx, b = yB - fy;
  int i0 = (int)fx, j0 = (int)fy;
  int i1 = i0 + 1, j1 = j0 + 1;
  i0 = imin(imax(i0, 0), cols - 1); i1 = imin(imax(i1, 0), cols - 1);
  j0 = imin(imax(j0, 0), rows - 1); j1 = imin(imax(j1, 0), rows - 1);
  float t00 = img[(long)j0 * cols + i0], t10 = img[(long)j0 * cols + i1];
  float t01 = img[(long)j1 * cols + i0], t11 = img[(long)j1 * cols + i1];
  if (g_tex_frac_mode == ORC_TEX_FRAC_ROUND) {
    int ka = (int)floorf(a * 256.f + 0.5f), kb = (int)floorf(b * 256.f + 0.5f);
    int w11 = (ka * kb + 128) >> 8;
    int w10 = ka - w11, w01 = kb - w11, w00 = 256 - ka - kb + w11;
    /* a tap with zero weight is not blended by the hardware: a NaN texel (the corner pixels of every
     * pyramid level >= 1 are NaN, pyrdown.cu:124-127) only poisons the result if its weight is non-zero */
    float acc = 0.f;
    if (w00) acc += (float)w00 * t00;
    if (w10) acc += (float)w10 * t10;
    if (w01) acc += (float)w01 * t01;
    if (w11) acc += (float)w11 * t11;
    return acc * (1.f / 256.f);
  }
  if (g_tex_frac_mode == ORC_TEX_FRAC_TRUNC) {
    a = floorf(a * 256.f) * (1.f / 256.f);
    b = floorf(b * 256.f) * (1.f / 256.f);
  }
  return (1.f - a) * (1.f - b) * t00 + a * (1.f - b) * t10 + (1.f - a) * b * t01 + a * b * t11;
}

/* trafo3DKernelInvDepthGridStride, warping_registration.cu:505-546 */
void orc_warp_invdepth(const float* src, const float* depth_prev, float* dst, int rows, int cols,
                       const float* Rp, const float* tp)
{
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float out = qnan();
      float w = depth_prev[(long)y * cols + x];
      if (!isnan(w)) {
        float xs, ys;
        float w3 = register_pixel(&xs, &ys, x, y, w, Rp, tp);
        xs += 0.5f; ys += 0.5f;
        int fx = (int)floorf(xs), fy = (int)floorf(ys);
        if (!(fx < 0 || fy < 0 || fx >= cols || fy >= rows)) {
          float w2 = tex_point(src, rows, cols, xs, ys);
          float tz = tp[2];
          float v1z = (1.f / w3 - tz) * w;
          float res = (v1z / (1.f - w2 * tz)) * w2;
          if (res > 0.f) out = res;
        }
      }
      dst[(long)y * cols + x] = out;
    }
}

/* trafo3DKernelIntensityWithInvDepthGridStride, warping_registration.cu:465-501 */
void orc_warp_intensity(const float* src, const float* depth_prev, float* dst, int rows, int cols,
                        const float* Rp, const float* tp)
{
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float out = qnan();
      float w = depth_prev[(long)y * cols + x];
      if (!isnan(w)) {
        float xs, ys;
        register_pixel(&xs, &ys, x, y, w, Rp, tp);
        xs += 0.5f; ys += 0.5f;
        int fx = (int)floorf(xs), fy = (int)floorf(ys);
        if (!(fx < 0 || fy < 0 || fx >= cols || fy >= rows)) {
          float r = tex_linear(src, rows, cols, xs, ys);
          out = fmaxf(0.f, fminf(r, 255.f));
        }
      }
      dst[(long)y * cols + x] = out;
    }
}

/* trafo3DKernelInvDepthWeightedGridStride, warping_registration.cu:549-594.
 * weight_warped is only written where the in-image test passes and weight_res > 0
 * (stale values survive elsewhere, as in the reference). */
void orc_warp_invdepth_weighted(const float* src, const float* depth_prev, float* dst,
                                float* weight_warped, int rows, int cols, const float* Rp,
                                const float* tp)
{
#pragma omp parallel for schedule(static)
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      long idx = (long)y * cols + x;
      dst[idx] = qnan();
      float w = depth_prev[idx];
      if (!isnan(w)) {
        float xs, ys;
        float w3 = register_pixel(&xs, &ys, x, y, w, Rp, tp);
        xs += 0.5f; ys += 0.5f;
        int fx = (int)floorf(xs), fy = (int)floorf(ys);
        if (!(fx < 0 || fy < 0 || fx >= cols || fy >= rows)) {
          float w2 = tex_point(src, rows, cols, xs, ys);
          float tz = tp[2];
          float v1z = (1.f / w3 - tz) * w;
          float wf = 1.f - w2 * tz;
          float wf2 = wf * wf;
          float weight_res = (wf2 * wf2) / (v1z * v1z);
          float res = (v1z / wf) * w2;
          if (res > 0.f) dst[idx] = res;
          if (weight_res > 0.f) weight_warped[idx] = weight_res;
        }
      }
    }
}

/* integrateWarpedFrameKernel, warping_registration.cu:637-669 (gate 3 * 0.0075, :80,660) */
void orc_integrate_warped_frame(const float* wsrc, const float* wweight, float* dst, float* dweight,
                                int rows, int cols)
{
  const float TH = 0.0075f;
  for (long i = 0; i < (long)rows * cols; ++i) {
    if (!isnan(wsrc[i])) {
      float w_sum = wsrc[i];
      float w_kf = dst[i];
      float dw = fabsf(w_sum - w_kf);
      if (isnan(w_kf)) {
        dst[i] = w_sum;
        dweight[i] = wweight[i];
      } else if (dw < 3 * TH) {
        float nw = dweight[i] + wweight[i];
        dst[i] = (w_kf * dweight[i] + w_sum * wweight[i]) / nw;
        dweight[i] = nw;
      }
    }
  }
}

/* partialVisibilityKernel / ...WithOverlapMaskKernel + finalVisibilityReductionKernel,
 * warping_registration.cu:297-461; ratio logic :863-866.  geom_tol is ignored (0.020). */
float orc_visibility_ratio(const float* depth_src, const float* depth_dst, int rows, int cols,
                           const float* Rp, const float* tp, uint8_t* overlap_mask,
                           double* n_visible, double* n_valid)
{
  double vis = 0.0, val = 0.0;
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float w = depth_src[(long)y * cols + x];
      if (!isnan(w)) {
        float xd, yd;
        float wd = register_pixel(&xd, &yd, x, y, w, Rp, tp);
        uint8_t flag = 0;
        val += 1.0;
        if (xd > 0 && xd < (float)(cols - 1) && yd > 0 && yd < (float)(rows - 1)) {
          int xi = (int)lrintf(xd), yi = (int)lrintf(yd);
          if (fabsf(wd - depth_dst[(long)yi * cols + xi]) < 0.020f) { vis += 1.0; flag = 1; }
        }
        if (overlap_mask) overlap_mask[(long)y * cols + x] = flag;
      }
    }
  if (n_visible) *n_visible = vis;
  if (n_valid) *n_valid = val;
  if ((float)val < 1.f) return 0.f;
  return (float)vis / (float)val;
}

/* ------------------------------------------------------------------------------------------- */
/* Residual sampling and scale estimation (src/cuda/sigmaFuncs.cu)                              */
/* ------------------------------------------------------------------------------------------- */

/* Sampling geometry of computeErrorGridStride, sigmaFuncs.cu:711-747 */
void orc_error_geometry(int rows, int cols, int min_nsamples, int* kept_rows, int* kept_cols,
                        int* stride)
{
  int error_size = cols * rows;
  int cp = cols, rp = rows;
  if (min_nsamples < error_size) {
    for (;;) {
      int cc = cp / 2, rc = rp / 2;
      if ((2 * cc - cp) != 0 || (2 * rc - rp) != 0 || min_nsamples > cc * rc) {
        error_size = cp * rp;
        break;
      }
      cp = cc; rp = rc;
    }
  }
  *kept_rows = rp; *kept_cols = cp;
  *stride = (int)sqrt((double)((rows * cols) / error_size));
}

/* errorGridStrideKernel, sigmaFuncs.cu:116-134: err[y*cols_kept+x] = im1(s y, s x) - im0(s y, s x) */
int orc_compute_error(const float* im1, const float* im0, int rows, int cols, int min_nsamples,
                      float* error)
{
  int kr, kc, s;
  orc_error_geometry(rows, cols, min_nsamples, &kr, &kc, &s);
  for (int y = 0; y < kr; ++y)
    for (int x = 0; x < kc; ++x)
      error[(long)y * kc + x] = im1[(long)(s * y) * cols + s * x] - im0[(long)(s * y) * cols + s * x];
  return kr * kc;
}

/* digamma: the reference calls boost::math::digamma(float) (src/cuda/device.hpp:76-80), an
 * un-vendored dependency (system Boost).  Boost's default policy evaluates float arguments in
 * double and rounds the result to float; restated here as the textbook recurrence +
 * asymptotic series in double, rounded to float.  Checked against scipy.special.digamma. */
double orc_digamma(double x)
{
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
             f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

static inline float digamma_f(float x) { return (float)orc_digamma((double)x); }

/* C(nu), sigmaFuncs.cu:966 (float arithmetic, left to right) */
static inline float c_nu(float nu, float fw)
{
  return -digamma_f(nu / 2.f) + logf(nu / 2.f) + fw + 1.f + digamma_f((nu + 1.f) / 2.f) -
         logf((nu + 1.f) / 2.f);
}

/* partialBiasAndSigmaStudent + finalReductionBiasAndSigma, sigmaFuncs.cu:281-409 */
static void moments_student(const float* err, int n, int lsq, float bias, float sigma, float nu,
                            float* bias_out, float* sigma_out)
{
  double s_wr2 = 0, s_wr = 0, s_w = 0, s_n = 0;
  for (int i = 0; i < n; ++i) {
    float e = err[i];
    if (!isinf(e) && !isnan(e)) {
      float weight;
      if (lsq) weight = 1.f;
      else {
        float en = (e - bias) / sigma;
        weight = (nu + 1.f) / (nu + en * en);
      }
      float wr = e * weight;
      float wr2 = wr * e;
      s_wr2 += wr2; s_wr += wr; s_w += weight; s_n += 1.0;
    }
  }
  float fwr2 = (float)s_wr2, fwr = (float)s_wr, fw = (float)s_w, fn = (float)s_n;
  float m0 = fwr / fw;
  float m1 = sqrtf((fwr2 - 2.f * m0 * fwr + m0 * m0 * fw) / fn);
  *bias_out = m0; *sigma_out = m1;
}

/* partialBiasAndSigma (generic M-estimator), sigmaFuncs.cu:179-278 */
static void moments_mest(const float* err, int n, int mest, float bias, float sigma,
                         float* bias_out, float* sigma_out)
{
  double s_wr2 = 0, s_wr = 0, s_w = 0, s_n = 0;
  for (int i = 0; i < n; ++i) {
    float e = err[i];
    if (!isinf(e) && !isnan(e)) {
      float weight = 1.f, valid = 1.f;
      float en = (e - bias) / sigma;
      if (mest == ORC_HUBER && fabsf(en) > THRESHOLD_HUBER) weight = THRESHOLD_HUBER / fabsf(en);
      else if (mest == ORC_TUKEY) {
        if (fabsf(en) < THRESHOLD_TUKEY) {
          float a = (en / THRESHOLD_TUKEY) * (en / THRESHOLD_TUKEY);
          weight = (1.f - a) * (1.f - a);
        } else { weight = 0.f; valid = 0.f; }
      } else if (mest == ORC_STUDENT) weight = (STUDENT_DOF + 1.f) / (STUDENT_DOF + en * en);
      float wr = e * weight;
      float wr2 = wr * e;
      s_wr2 += wr2; s_wr += wr; s_w += weight; s_n += valid;
    }
  }
  float fwr2 = (float)s_wr2, fwr = (float)s_wr, fw = (float)s_w, fn = (float)s_n;
  float m0 = fwr / fw;
  float m1 = sqrtf((fwr2 - 2.f * m0 * fwr + m0 * m0 * fw) / fn);
  *bias_out = m0; *sigma_out = m1;
}

/* partialFuncWeightsNu + finalReductionFuncWeightsNu, sigmaFuncs.cu:412-516 */
static float func_weights_nu(const float* err, int n, float bias, float sigma, float nu)
{
  double s_ln = 0, s_w = 0, s_n = 0;
  for (int i = 0; i < n; ++i) {
    float e = err[i];
    if (!isinf(e) && !isnan(e)) {
      float en = (e - bias) / sigma;
      float weight = (nu + 1.f) / (nu + en * en);
      s_ln += logf(weight); s_w += weight; s_n += 1.0;
    }
  }
  return ((float)s_ln - (float)s_w) / (float)s_n;
}

/* nu bisection, sigmaFuncs.cu:946-1047 (identical in computeNuStudent :1104-1207) */
static float estimate_nu(const float* err, int n, float bias, float sigma)
{
  float nu_up = 10.f, nu_down = 2.f, nu_new = 0.f, nu;
  float C_down = c_nu(nu_down, func_weights_nu(err, n, bias, sigma, nu_down));
  float C_up = c_nu(nu_up, func_weights_nu(err, n, bias, sigma, nu_up));
  if (C_up * C_down > 0) {
    nu = (C_down <= 0.f) ? nu_down : nu_up;
  } else {
    for (int j = 0; j < 5; ++j) {
      nu_new = (nu_up + nu_down) / 2;
      if ((nu_up - nu_down) < 1.f) break;
      float C_new = c_nu(nu_new, func_weights_nu(err, n, bias, sigma, nu_new));
      if (C_new * C_up > 0) { C_up = C_new; nu_up = nu_new; }
      else { C_down = C_new; nu_down = nu_new; }
    }
    nu = nu_new;
  }
  return nu;
}

/* computeSigmaAndNuStudent, sigmaFuncs.cu:858-1066 */
int orc_sigma_nu_student(const float* err, int n, float* bias, float* sigma, float* nu, int mest)
{
  float sh_sigma = *sigma, sh_bias = *bias, sh_nu = 5.f;
  int lsq = 1;
  float sigma_prev;
  int iters = 0;
  for (int i = 0; i < 10; ++i) {
    float b, s;
    moments_student(err, n, lsq, sh_bias, sh_sigma, sh_nu, &b, &s);
    *bias = b; *sigma = s;
    sigma_prev = sh_sigma;
    sh_bias = b; sh_sigma = s;
    lsq = (mest == ORC_LSQ);
    ++iters;
    if (i > 0 && (fabsf(s - sigma_prev) / sigma_prev) < 0.1f) break;
  }
  *nu = estimate_nu(err, n, sh_bias, sh_sigma);
  return iters;
}

/* computeNuStudent, sigmaFuncs.cu:1068-1222 */
void orc_nu_student(const float* err, int n, float bias, float sigma, float* nu)
{
  *nu = estimate_nu(err, n, bias, sigma);
}

/* computeSigmaPdf, sigmaFuncs.cu:773-854 */
int orc_sigma_pdf(const float* err, int n, float* bias, float* sigma, int mest)
{
  float sh_sigma = *sigma, sh_bias = *bias;
  int cur = ORC_LSQ, iters = 0;
  for (int i = 0; i < 10; ++i) {
    float b, s;
    moments_mest(err, n, cur, sh_bias, sh_sigma, &b, &s);
    *bias = b; *sigma = s;
    ++iters;
    if (i > 0 && (fabsf(s - sh_sigma) / sh_sigma) < 0.1f) break;
    sh_bias = b; sh_sigma = s; cur = mest;
  }
  return iters;
}

/* computeChiSquare, sigmaFuncs.cu:1225-1297 (kernels :137-150, :541-647) */
void orc_chi_square(const float* err_int, const float* err_depth, int n, float sigma_int,
                    float sigma_depth, int mest, float* chi_squared, float* chi_test, float* ndof)
{
  double s_rho = 0, s_n = 0;
  for (int k = 0; k < 2 * n; ++k) {
    float e = (k < n) ? err_int[k] / sigma_int : err_depth[k - n] / sigma_depth;
    if (!isinf(e) && !isnan(e)) {
      float rho = (e * e) / 2.f;
      if (mest == ORC_HUBER && fabsf(e) > THRESHOLD_HUBER)
        rho = THRESHOLD_HUBER * (fabsf(e) - THRESHOLD_HUBER / 2.f);
      else if (mest == ORC_TUKEY) {
        if (fabsf(e) < THRESHOLD_TUKEY) {
          float a1 = (e / THRESHOLD_TUKEY) * (e / THRESHOLD_TUKEY);
          float a2 = (1.f - a1) * (1.f - a1) * (1.f - a1);
          rho = ((THRESHOLD_TUKEY * THRESHOLD_TUKEY) / 6.f) * (1.f - a2);
        } else rho = (THRESHOLD_TUKEY * THRESHOLD_TUKEY) / 6.f;
      } else if (mest == ORC_STUDENT)
        rho = ((STUDENT_DOF + 1.f) / 2.f) * logf(1.f + (e * e) / STUDENT_DOF);
      s_n += 1.0; s_rho += rho;
    }
  }
  float fn = (float)s_n;
  *chi_squared = (float)s_rho / fn;
  *ndof = fn;
  float z = (*chi_squared - fn) / sqrtf(2.f * fn);
  *chi_test = 0.5f * (1.f + erff(z / sqrtf(2.f)));
}

/* ------------------------------------------------------------------------------------------- */
/* Normal equations (src/cuda/estimate_VO.cu)                                                   */
/* ------------------------------------------------------------------------------------------- */

static inline void cross3(const float* a, const float* b, float* c)
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* computeWeight, estimate_VO.cu:141-167 */
static inline float weight_mest(float e, int mest)
{
  float w = 1.f;
  if (mest == ORC_HUBER) { if (fabsf(e) > THRESHOLD_HUBER) w = THRESHOLD_HUBER / fabsf(e); }
  else if (mest == ORC_TUKEY) {
    if (fabsf(e) < THRESHOLD_TUKEY) {
      float a = (e / THRESHOLD_TUKEY) * (e / THRESHOLD_TUKEY);
      w = (1.f - a) * (1.f - a);
    } else w = 0.f;
  } else if (mest == ORC_STUDENT) w = (STUDENT_DOF + 1.f) / (STUDENT_DOF + e * e);
  return w;
}

/* One pixel of computeSystemGridStride / computeStudentNuSystemGridStride,
 * estimate_VO.cu:176-262 (constraints) and :295-329 / :384-418 (weights, accumulation).
 * Adds this pixel's 27 contributions (float, reference ordering) into acc (double). */
static inline void system_pixel(int x, int y, float w0, float i0, float gwx, float gwy, float gix,
                                float giy, float w1, float i1, const orc_system_params* P,
                                double* acc, double* chi)
{
  float row_i[6] = {0, 0, 0, 0, 0, 0}, row_d[6] = {0, 0, 0, 0, 0, 0};
  float err_i = 0.f, err_d = 0.f, wgt_i = 0.f, wgt_d = 0.f, n_factor = 1.f;
  float px = ((float)x - P->cx) / P->fx;
  float py = ((float)y - P->cy) / P->fy;
  float p[3] = {px, py, 1.f};

  /* invDepthConstraint :214-262 */
  if (!(isnan(w0) || isnan(w1) || isnan(gwx) || isnan(gwy))) {
    float g[3];
    g[0] = gwx * P->fx; g[1] = gwy * P->fy; g[2] = -(g[0] * p[0] + g[1] * p[1]);
    float inv_w0 = 1.f / w0;
    float n[3] = {g[0] * inv_w0, g[1] * inv_w0, g[2] * inv_w0};
    n[2] += 1.f;
    float rn = 1.f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] *= rn; n[1] *= rn; n[2] *= rn;
    float rp = 1.f / sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    float pu[3] = {p[0] * rp, p[1] * rp, p[2] * rp};
    n_factor = fabsf(n[0] * pu[0] + n[1] * pu[1] + n[2] * pu[2]);
    float weight = 1.f / P->sigma_depthinv;
    float rt[3] = {g[0] * w0, g[1] * w0, g[2] * w0};
    rt[2] = rt[2] + w0 * w1;
    g[2] = g[2] + w1;
    float rr[3];
    cross3(g, p, rr);
    rr[0] = -rr[0]; rr[1] = -rr[1]; rr[2] = -rr[2];
    float b = (w1 - w0);
    for (int k = 0; k < 3; ++k) { row_d[k] = rt[k] * weight; row_d[3 + k] = rr[k] * weight; }
    err_d = -b * weight;
    float eu = err_d - (P->bias_depthinv / P->sigma_depthinv);
    float wv = P->student_nu ? (P->nu_depthinv + 1.f) / (P->nu_depthinv + eu * eu)
                             : weight_mest(eu, P->mestimator);
    wgt_d = wv * (float)(1 - (P->weighting == ORC_PHOT_ONLY));
    if (chi) { chi[0] += 1.0; chi[1] += (double)(b * b); }
  }
  /* intensityConstraint :176-212 */
  if (!(isnan(w0) || isnan(i0) || isnan(i1) || isnan(gix) || isnan(giy))) {
    float g[3];
    g[0] = gix * P->fx; g[1] = giy * P->fy; g[2] = -(g[0] * p[0] + g[1] * p[1]);
    float weight = 1.f / P->sigma_int;
    float rr[3];
    cross3(g, p, rr);
    rr[0] = -rr[0]; rr[1] = -rr[1]; rr[2] = -rr[2];
    float rt[3] = {g[0] * w0, g[1] * w0, g[2] * w0};
    float b = (i1 - i0);
    for (int k = 0; k < 3; ++k) { row_i[k] = rt[k] * weight; row_i[3 + k] = rr[k] * weight; }
    err_i = -b * weight;
    float eu = err_i - (P->bias_int / P->sigma_int);
    float wv = P->student_nu ? (P->nu_int + 1.f) / (P->nu_int + eu * eu)
                             : weight_mest(eu, P->mestimator);
    wgt_i = wv * (float)(1 - (P->weighting == ORC_GEOM_ONLY));
    if (chi) { chi[2] += 1.0; chi[3] += (double)(b * b); }
  }
  if (P->weighting == ORC_MIN_WEIGHT) wgt_i = fminf(wgt_d, wgt_i);
  if (wgt_i == 0.f && wgt_d == 0.f) return;

  int shift = 0;
  for (int i = 0; i < 6; ++i) {
    for (int j = i; j < 6; ++j)
      acc[shift++] += (double)(wgt_i * (row_i[i] * row_i[j]) + n_factor * wgt_d * (row_d[i] * row_d[j]));
    acc[shift++] += (double)(wgt_i * (row_i[i] * err_i) + n_factor * wgt_d * (row_d[i] * err_d));
  }
}

/* Host unpack, estimate_VO.cu:627-642 / :771-786 */
void orc_unpack_system(const double* sums27, double* A36, double* b6)
{
  int shift = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 7; ++j) {
      double v = sums27[shift++];
      if (j == 6) b6[i] = v;
      else A36[j * 6 + i] = A36[i * 6 + j] = v;
    }
}

/* buildSystemGridStride / buildSystemStudentNuGridStride, estimate_VO.cu:505-789, on
 * pre-warped W1 / I1 maps (the reference's unfused interface). */
void orc_build_system(const float* W0, const float* I0, const float* gWx, const float* gWy,
                      const float* gIx, const float* gIy, const float* W1, const float* I1, int rows,
                      int cols, const orc_system_params* P, double* sums27, double* A36, double* b6)
{
  double acc[27];
  memset(acc, 0, sizeof(acc));
#pragma omp parallel
  {
    double loc[27];
    memset(loc, 0, sizeof(loc));
#pragma omp for schedule(static) nowait
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) {
        long i = (long)y * cols + x;
        system_pixel(x, y, W0[i], I0[i], gWx[i], gWy[i], gIx[i], gIy[i], W1[i], I1[i], P, loc, NULL);
      }
#pragma omp critical
    for (int k = 0; k < 27; ++k) acc[k] += loc[k];
  }
  if (sums27) memcpy(sums27, acc, sizeof(acc));
  if (A36 && b6) orc_unpack_system(acc, A36, b6);
}

/* ------------------------------------------------------------------------------------------- */
/* Vertex / normal maps (src/cuda/maps.cu)                                                      */
/* ------------------------------------------------------------------------------------------- */

/* computeVmapKernel, maps.cu:63-90: 3 stacked planes [3*rows][cols]; for an invalid pixel only
 * the x plane is set to NaN (y and z planes are left untouched, as in the reference). */
void orc_vmap(const float* depth_inv, int rows, int cols, float fx, float fy, float cx, float cy,
              float* vmap)
{
  float fxi = 1.f / fx, fyi = 1.f / fy;
  for (int v = 0; v < rows; ++v)
    for (int u = 0; u < cols; ++u) {
      float z = 1.f / depth_inv[(long)v * cols + u];
      if (!isnan(z)) {
        vmap[(long)v * cols + u] = z * ((float)u - cx) * fxi;
        vmap[(long)(v + rows) * cols + u] = z * ((float)v - cy) * fyi;
        vmap[(long)(v + 2 * rows) * cols + u] = z;
      } else vmap[(long)v * cols + u] = qnan();
    }
}

/* computeNmapGradientsKernel, maps.cu:134-179: only the x plane is pre-set to NaN. */
void orc_nmap_gradients(const float* depth_inv, const float* gx_, const float* gy_, int rows,
                        int cols, float fx, float fy, float cx, float cy, float* nmap)
{
  for (int v = 0; v < rows; ++v)
    for (int u = 0; u < cols; ++u) {
      long i = (long)v * cols + u;
      nmap[i] = qnan();
      float w = depth_inv[i], gx = gx_[i], gy = gy_[i];
      if (!(isnan(w) || isnan(gx) || isnan(gy))) {
        float nx = gx * fx, ny = gy * fy;
        float nz = gx * (cx - (float)u) + gy * (cy - (float)v) + w;
        float rn = 1.f / sqrtf(nx * nx + ny * ny + nz * nz);
        nx *= rn; ny *= rn; nz *= rn;
        float z = 1.f / w;
        float vx = z * ((float)u - cx) * (1.f / fx);
        float vy = z * ((float)v - cy) * (1.f / fy);
        float vz = z;
        float rv = 1.f / sqrtf(vx * vx + vy * vy + vz * vz);
        float d = (vx * rv) * nx + (vy * rv) * ny + (vz * rv) * nz;
        if ((double)d > 0.1) {
          nmap[i] = nx;
          nmap[(long)(v + rows) * cols + u] = ny;
          nmap[(long)(v + 2 * rows) * cols + u] = nz;
        }
      }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Host-side Gauss-Newton algebra (double), src/util_funcs.cpp:31-155, src/visodo.cpp:1242-1263 */
/* ------------------------------------------------------------------------------------------- */

static void mat3_mul(const double* A, const double* B, double* C)
{
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof(T));
}

static void mat3_vec(const double* A, const double* v, double* r)
{
  double t[3];
  for (int i = 0; i < 3; ++i) t[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
  memcpy(r, t, sizeof(t));
}

void orc_mat3_inverse(const double* M, double* Mi)
{
  double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  double id = 1.0 / det;
  double T[9];
  T[0] = c00 * id; T[1] = (M[2] * M[7] - M[1] * M[8]) * id; T[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  T[3] = c01 * id; T[4] = (M[0] * M[8] - M[2] * M[6]) * id; T[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  T[6] = c02 * id; T[7] = (M[1] * M[6] - M[0] * M[7]) * id; T[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  memcpy(Mi, T, sizeof(T));
}

/* forceOrthogonalisation, util_funcs.cpp:150-155: U V^T of the SVD = orthogonal polar factor.
 * Computed with the Newton iteration X <- (X + X^-T)/2, which converges quadratically to the
 * same factor for any non-singular input. */
void orc_force_orthogonal(const double* M, double* R)
{
  double X[9];
  memcpy(X, M, sizeof(X));
  for (int it = 0; it < 50; ++it) {
    double Xi[9], Y[9];
    orc_mat3_inverse(X, Xi);
    double diff = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Y[3 * i + j] = 0.5 * (X[3 * i + j] + Xi[3 * j + i]);
        diff += fabs(Y[3 * i + j] - X[3 * i + j]);
      }
    memcpy(X, Y, sizeof(X));
    if (diff < 1e-16) break;
  }
  memcpy(R, X, sizeof(X));
}

static void skew3(const double* w, double* S)
{
  S[0] = 0; S[1] = -w[2]; S[2] = w[1];
  S[3] = w[2]; S[4] = 0; S[5] = -w[0];
  S[6] = -w[1]; S[7] = w[0]; S[8] = 0;
}

/* expMapRot, util_funcs.cpp:125-148 */
void orc_exp_map_rot(const double* omega, double* R)
{
  double theta = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double O[9], O2[9], M[9];
  skew3(omega, O);
  mat3_mul(O, O, O2);
  double a, b;
  if (theta < 0.00001) { a = 1.0; b = 0.5; }
  else { a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta); }
  for (int i = 0; i < 9; ++i) M[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
  orc_force_orthogonal(M, R);
}

/* expMap, util_funcs.cpp:85-123: T = [R | Q v] */
void orc_exp_map(const double* omega, const double* v, double* R, double* t)
{
  double theta = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double O[9], O2[9], M[9], Q[9];
  skew3(omega, O);
  mat3_mul(O, O, O2);
  double a, b, c;
  if (theta < 0.00001) { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
  else {
    a = sin(theta) / theta;
    b = (1 - cos(theta)) / (theta * theta);
    c = (1 - (sin(theta) / theta)) / (theta * theta);
  }
  for (int i = 0; i < 9; ++i) {
    double I = (i % 4 == 0) ? 1.0 : 0.0;
    M[i] = I + a * O[i] + b * O2[i];
    Q[i] = I + b * O[i] + c * O2[i];
  }
  orc_force_orthogonal(M, R);
  mat3_vec(Q, v, t);
}

/* logMap, util_funcs.cpp:31-82: twist = [v; omega] */
void orc_log_map(const double* Rin, const double* trans, double* twist)
{
  double R[9];
  orc_force_orthogonal(Rin, R);
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = acos(c), theta2 = theta * theta, th_by_sinth;
  if (s < 1e-5) th_by_sinth = 1.0 + (1.0 / 6.0) * theta2 + (7.0 / 360.0) * theta2 * theta2;
  else th_by_sinth = theta / s;
  double vth = th_by_sinth / 2.0;
  double om[3] = {rx * vth, ry * vth, rz * vth};
  double O[9], O2[9], Q[9], Qi[9];
  skew3(om, O);
  mat3_mul(O, O, O2);
  double th = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  double b, cc;
  if (th < 0.00001) { b = 0.5; cc = 1.0 / 6.0; }
  else { b = (1 - cos(theta)) / (theta * theta); cc = (1 - (sin(theta) / theta)) / (theta * theta); }
  for (int i = 0; i < 9; ++i) Q[i] = ((i % 4 == 0) ? 1.0 : 0.0) + b * O[i] + cc * O2[i];
  orc_mat3_inverse(Q, Qi);
  mat3_vec(Qi, trans, twist);
  twist[3] = om[0]; twist[4] = om[1]; twist[5] = om[2];
}

/* A.llt().solve(b), visodo.cpp:1249 (Eigen LLT = Cholesky, un-vendored; textbook restatement).
 * Returns 0 on success, 1 if A is not positive definite (x is then filled with NaN, which is
 * what propagates to the reference's NaN-pose guard, visodo.cpp:1265). */
int orc_llt_solve6(const double* A, const double* b, double* x)
{
  double L[36];
  memset(L, 0, sizeof(L));
  int bad = 0;
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
    if (!(d > 0)) bad = 1;
    double ljj = sqrt(d);
    L[j * 6 + j] = ljj;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = s / ljj;
    }
  }
  double yv[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * yv[k];
    yv[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = yv[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s / L[i * 6 + i];
  }
  return bad;
}

/* A.inverse() for the 6x6 covariance (visodo.cpp:1409, keyframe_align.cpp:350); Gauss-Jordan
 * with partial pivoting (Eigen uses PartialPivLU for sizes > 4). */
int orc_inverse6(const double* A, double* Ai)
{
  double M[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { M[i][j] = A[i * 6 + j]; M[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
    if (M[p][c] == 0.0) return 1;
    if (p != c) for (int j = 0; j < 12; ++j) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
    double ip = 1.0 / M[c][c];
    for (int j = 0; j < 12; ++j) M[c][j] *= ip;
    for (int r = 0; r < 6; ++r) if (r != c) {
      double f = M[r][c];
      if (f != 0.0) for (int j = 0; j < 12; ++j) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Ai[i * 6 + j] = M[i][6 + j];
  return 0;
}

/* K R^-1 K^-1 and K t^-1 in float, visodo.cpp:1066-1067,1108-1114 (Eigen float products). */
void orc_projective_inverse_pose(const double* R, const double* t, float fx, float fy, float cx,
                                 float cy, float* Rp, float* tp)
{
  double Ri[9], ti[3];
  orc_mat3_inverse(R, Ri);
  mat3_vec(Ri, t, ti);
  for (int k = 0; k < 3; ++k) ti[k] = -ti[k];
  orc_projective_pose(Ri, ti, fx, fy, cx, cy, Rp, tp);
}

/* K R K^-1 and K t in float for an already-chosen direction (visodo.cpp:1496-1500). */
void orc_projective_pose(const double* R, const double* t, float fx, float fy, float cx, float cy,
                         float* Rp, float* tp)
{
  float Rf[9], tf[3];
  for (int k = 0; k < 9; ++k) Rf[k] = (float)R[k];
  for (int k = 0; k < 3; ++k) tf[k] = (float)t[k];
  float K[9] = {fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f};
  float Ki[9] = {1.f / fx, 0.f, -cx / fx, 0.f, 1.f / fy, -cy / fy, 0.f, 0.f, 1.f};
  float T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      T[3 * i + j] = K[3 * i] * Rf[j] + K[3 * i + 1] * Rf[3 + j] + K[3 * i + 2] * Rf[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Rp[3 * i + j] = T[3 * i] * Ki[j] + T[3 * i + 1] * Ki[3 + j] + T[3 * i + 2] * Ki[6 + j];
  for (int i = 0; i < 3; ++i) tp[i] = K[3 * i] * tf[0] + K[3 * i + 1] * tf[1] + K[3 * i + 2] * tf[2];
}

/* One Gauss-Newton pose update, visodo.cpp:1242-1263 / keyframe_align.cpp:312-335.
 * x = [trans; rot]; R_inc^-1 = expMapRot(rot); T <- T_inc T.  Returns 1 if the pose went NaN. */
int orc_gn_update(const double* A36, const double* b6, double* R, double* t, double* x_out)
{
  double x[6];
  orc_llt_solve6(A36, b6, x);
  if (x_out) memcpy(x_out, x, sizeof(x));
  double Rinc_inv[9], Rinc[9], tinc[3];
  orc_exp_map_rot(x + 3, Rinc_inv);
  orc_mat3_inverse(Rinc_inv, Rinc);
  mat3_vec(Rinc, x, tinc);
  for (int k = 0; k < 3; ++k) tinc[k] = -tinc[k];
  double tn[3];
  mat3_vec(Rinc, t, tn);
  for (int k = 0; k < 3; ++k) t[k] = tn[k] + tinc[k];
  mat3_mul(Rinc, R, R);
  double nr = 0, nt = 0;
  for (int k = 0; k < 9; ++k) nr += R[k] * R[k];
  for (int k = 0; k < 3; ++k) nt += t[k] * t[k];
  return (isnan(nr) || isnan(nt)) ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Coarse-to-fine drivers                                                                       */
/* ------------------------------------------------------------------------------------------- */

static inline void level_intr(const orc_align_config* C, int level, float* fx, float* fy, float* cx,
                              float* cy)
{
  /* Intr::operator()(level), src/internal.h:128-132 */
  int div = 1 << level;
  *fx = C->fx / div; *fy = C->fy / div; *cx = C->cx / div; *cy = C->cy / div;
}

/* Restates the Gauss-Newton loops of VisodoTracker::estimateVisualOdometry
 * (src/visodo.cpp:1041-1281 + covariance pass :1283-1415; mode ORC_MODE_TRACKER) and
 * KeyframeAlign::alignKeyframes (src/keyframe_align.cpp:178-350; mode ORC_MODE_ALIGN) on
 * prepared pyramids.  R, t hold the initial guess on entry and the estimate on return. */
int orc_align(const orc_align_config* C, const orc_pyramids* P, double* R, double* t, double* cov36,
              orc_iter_trace* trace, int trace_cap, int* n_trace, orc_frame_stats* stats)
{
  int nt = 0;
  double A[36], b[6], sums[27];
  memset(A, 0, sizeof(A));
  int rows0 = C->rows, cols0 = C->cols;
  long maxpix = (long)rows0 * cols0;
  float* W1buf = (float*)malloc(sizeof(float) * maxpix);
  float* I1buf = (float*)malloc(sizeof(float) * maxpix);
  float* eI = (float*)malloc(sizeof(float) * maxpix);
  float* eW = (float*)malloc(sizeof(float) * maxpix);
  /* WARP_ORDER = warpFirst (tracker only, src/visodo.cpp:1078-1105): second pair of buffers for the per-iteration
   * pyramid of the warped maps */
  const int warp_first = C->warp_first && C->mode == ORC_MODE_TRACKER;
  float* W2 = warp_first ? (float*)malloc(sizeof(float) * maxpix) : NULL;
  float* I2 = warp_first ? (float*)malloc(sizeof(float) * maxpix) : NULL;
  int status = 0;
  /* CHI_SQUARED termination, src/visodo.cpp:1134-1164: RMSE and RMSE_prev live across levels (:987-988), the pose
   * before the last increment is what the reference's undo (:1147-1148) restores */
  const int chi_term = (C->termination == ORC_TERM_CHI_SQUARED) && C->mode == ORC_MODE_TRACKER;
  float rmse_prev = 9999.f;
  double R_before[9], t_before[3];
  memcpy(R_before, R, sizeof(R_before)); memcpy(t_before, t, sizeof(t_before));

  for (int level = C->levels - 1; level >= C->finest_level && !status; --level) {
    int rows = rows0 >> level, cols = cols0 >> level;
    float fx, fy, cx, cy;
    level_intr(C, level, &fx, &fy, &cx, &cy);
    for (int iter = 0; iter < C->iterations[level]; ++iter) {
      float Rp[9], tp[3];
      float *W1 = W1buf, *I1 = I1buf;
      int end_level = 0;
      if (warp_first && level > 0) {
        /* "expensive warping": warp at level 0 with the level-0 calibration, then build the pyramid of the
         * warped maps down to this level, every iteration (visodo.cpp:1078-1105) */
        float fx0, fy0, cx0, cy0;
        level_intr(C, 0, &fx0, &fy0, &cx0, &cy0);
        orc_projective_inverse_pose(R, t, fx0, fy0, cx0, cy0, Rp, tp);
        orc_warp_invdepth(P->W_cur[0], P->W_kf[0], W1, rows0, cols0, Rp, tp);
        orc_warp_intensity(P->I_cur[0], W1, I1, rows0, cols0, Rp, tp);
        if (chi_term && iter != 0) { /* the test reads the level-0 warped maps, :1136-1139 */
          float chi2, chit, ndof;
          int n = orc_compute_error(I1, P->I_kf[0], rows0, cols0, 9999999, eI);
          orc_compute_error(W1, P->W_kf[0], rows0, cols0, 9999999, eW);
          orc_chi_square(eI, eW, n, 5.f, 0.0025f, C->mestimator, &chi2, &chit, &ndof);
          float rmse = sqrtf(chi2) / sqrtf(ndof);
          if (iter != 1 && rmse > rmse_prev) end_level = 1; else rmse_prev = rmse;
        }
        float *Wn = W2, *In = I2;
        for (int i = 1; i <= level; ++i) {
          orc_pyr_down(I1, rows0 >> (i - 1), cols0 >> (i - 1), In);
          orc_pyr_down(W1, rows0 >> (i - 1), cols0 >> (i - 1), Wn);
          float* sw = W1; W1 = Wn; Wn = sw;
          sw = I1; I1 = In; In = sw;
        }
      } else {
        orc_projective_inverse_pose(R, t, fx, fy, cx, cy, Rp, tp);
        orc_warp_invdepth(P->W_cur[level], P->W_kf[level], W1, rows, cols, Rp, tp);
        /* tracker warps intensity with the just-warped iD as geometry (visodo.cpp:1121-1126);
         * KeyframeAlign uses the keyframe iD (keyframe_align.cpp:239) */
        orc_warp_intensity(P->I_cur[level], C->mode == ORC_MODE_TRACKER ? W1 : P->W_kf[level], I1, rows,
                           cols, Rp, tp);
        /* pyrFirst: warped_*_curr_[0] is only refreshed while level 0 iterates; above it the reference tests stale
         * maps, the RMSE repeats and `RMSE > RMSE_prev` never holds -- the test is restated at level 0 only */
        if (chi_term && iter != 0 && level == 0) {
          float chi2, chit, ndof;
          int n = orc_compute_error(I1, P->I_kf[0], rows0, cols0, 9999999, eI);
          orc_compute_error(W1, P->W_kf[0], rows0, cols0, 9999999, eW);
          orc_chi_square(eI, eW, n, 5.f, 0.0025f, C->mestimator, &chi2, &chit, &ndof);
          float rmse = sqrtf(chi2) / sqrtf(ndof);
          if (iter != 1 && rmse > rmse_prev) end_level = 1; else rmse_prev = rmse;
        }
      }
      if (end_level) { /* undo the previous increment and end the iterations at this level, :1145-1151 */
        memcpy(R, R_before, sizeof(R_before)); memcpy(t, t_before, sizeof(t_before));
        break;
      }
      orc_system_params S;
      memset(&S, 0, sizeof(S));
      S.fx = fx; S.fy = fy; S.cx = cx; S.cy = cy;
      S.mestimator = C->mestimator; S.weighting = C->weighting; S.student_nu = 1;
      S.sigma_int = 5.f; S.sigma_depthinv = 0.0025f; S.bias_int = 0.f; S.bias_depthinv = 0.f;
      S.nu_int = 5.f; S.nu_depthinv = 5.f;
      int iters_i = 0, iters_w = 0;
      if (C->mode == ORC_MODE_TRACKER) {
        if (C->sigma_estimator == ORC_SIGMA_PDF) {
          int n = orc_compute_error(I1, P->I_kf[level], rows, cols, C->nsamples, eI);
          orc_compute_error(W1, P->W_kf[level], rows, cols, C->nsamples, eW);
          iters_i = orc_sigma_nu_student(eI, n, &S.bias_int, &S.sigma_int, &S.nu_int, C->mestimator);
          iters_w = orc_sigma_nu_student(eW, n, &S.bias_depthinv, &S.sigma_depthinv, &S.nu_depthinv,
                                         C->mestimator);
          S.nu_int = fmaxf(S.nu_int, S.nu_depthinv); /* visodo.cpp:1186 */
        }
      } else {
        int n = orc_compute_error(W1, P->W_kf[level], rows, cols, C->nsamples, eW);
        orc_compute_error(I1, P->I_kf[level], rows, cols, C->nsamples, eI);
        orc_nu_student(eW, n, S.bias_depthinv, S.sigma_depthinv, &S.nu_depthinv);
        orc_nu_student(eI, n, S.bias_int, S.sigma_int, &S.nu_int);
        /* keyframe_align.cpp:297,308: nu_depthinv is passed for BOTH residuals */
        S.nu_int = S.nu_depthinv;
      }
      orc_build_system(P->W_kf[level], P->I_kf[level], P->gWx_kf[level], P->gWy_kf[level],
                       P->gIx_kf[level], P->gIy_kf[level], W1, I1, rows, cols, &S, sums, A, b);
      double x[6];
      memcpy(R_before, R, sizeof(R_before)); memcpy(t_before, t, sizeof(t_before));
      int bad = orc_gn_update(A, b, R, t, x);
      if (trace && nt < trace_cap) {
        orc_iter_trace* T = &trace[nt];
        T->level = level; T->iter = iter;
        memcpy(T->sums27, sums, sizeof(sums));
        T->sigma_int = S.sigma_int; T->sigma_depthinv = S.sigma_depthinv;
        T->bias_int = S.bias_int; T->bias_depthinv = S.bias_depthinv;
        T->nu_int = S.nu_int; T->nu_depthinv = S.nu_depthinv;
        T->irls_iters_int = iters_i; T->irls_iters_depthinv = iters_w;
        memcpy(T->x, x, sizeof(x));
        memcpy(T->R, R, sizeof(double) * 9); memcpy(T->t, t, sizeof(double) * 3);
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
    /* lost: cov = 100 I (visodo.cpp:1269) */
    if (cov36) for (int i = 0; i < 36; ++i) cov36[i] = (i % 7 == 0) ? 100.0 : 0.0;
  } else if (C->mode == ORC_MODE_TRACKER) {
    /* covariance pass, visodo.cpp:1283-1415 (warps at the finest level in both warp orders) */
    float *W1 = W1buf, *I1 = I1buf;
    int level = C->finest_level;
    int rows = rows0 >> level, cols = cols0 >> level;
    float fx, fy, cx, cy, Rp[9], tp[3];
    level_intr(C, level, &fx, &fy, &cx, &cy);
    orc_projective_inverse_pose(R, t, fx, fy, cx, cy, Rp, tp);
    orc_warp_invdepth(P->W_cur[level], P->W_kf[level], W1, rows, cols, Rp, tp);
    orc_warp_intensity(P->I_cur[level], W1, I1, rows, cols, Rp, tp);
    orc_system_params S;
    memset(&S, 0, sizeof(S));
    S.fx = fx; S.fy = fy; S.cx = cx; S.cy = cy;
    S.mestimator = ORC_STUDENT; S.weighting = C->weighting; S.student_nu = 0;
    S.sigma_int = expf(logf(5.f) - 0.f * logf(2.f));
    S.sigma_depthinv = expf(logf(0.0025f) - 0.f * logf(2.f));
    orc_build_system(P->W_kf[level], P->I_kf[level], P->gWx_cov[level], P->gWy_cov[level],
                     P->gIx_cov[level], P->gIy_cov[level], W1, I1, rows, cols, &S, sums, A, b);
    if (cov36) orc_inverse6(A, cov36);
    if (stats) {
      memcpy(stats->cov_sums27, sums, sizeof(sums));
      /* end-of-frame chi^2 on FULL-resolution residuals (visodo.cpp:1411-1415) */
      int n = orc_compute_error(I1, P->I_kf[level], rows, cols, 9999999, eI);
      orc_compute_error(W1, P->W_kf[level], rows, cols, 9999999, eW);
      orc_chi_square(eI, eW, n, 5.f, 0.0025f, C->mestimator, &stats->chi_square, &stats->chi_test,
                     &stats->ndof);
    }
  } else {
    if (cov36) orc_inverse6(A, cov36); /* keyframe_align.cpp:339-350: last iteration's A */
  }
  if (n_trace) *n_trace = nt;
  free(W1buf); free(I1buf); free(eI); free(eW); free(W2); free(I2);
  return status;
}
