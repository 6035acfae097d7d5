// api.cu -- context management and the drop-in (one call per reference bridge function) entry points.
#include <cstdio>
#include <cstring>
#include <cmath>

#include "ctx.hpp"

using namespace rgbid;

namespace rgbid {

int ctx_reserve_host_stage(rgbid_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->h_stage_bytes) return RGBID_OK;
  if (ctx->h_stage) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; ctx->h_stage_bytes = 0; }
  RGBID_CUDA_TRY(cudaMallocHost(&ctx->h_stage, bytes));
  ctx->h_stage_bytes = bytes;
  return RGBID_OK;
}

int ctx_reserve_device_stage(rgbid_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->d_stage_bytes) return RGBID_OK;
  if (ctx->d_stage) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_stage); ctx->d_stage = nullptr; ctx->d_stage_bytes = 0; }
  RGBID_CUDA_TRY(cudaMalloc(&ctx->d_stage, bytes));
  ctx->d_stage_bytes = bytes;
  return RGBID_OK;
}

int check_last(rgbid_ctx* ctx)
{
  (void)ctx;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? RGBID_OK : RGBID_ERR_CUDA_BASE + (int)e;
}

}  // namespace rgbid

static inline Proj make_proj(const float* Rp, const float* tp)
{
  Proj P;
  for (int i = 0; i < 9; ++i) P.r[i] = Rp[i];
  for (int i = 0; i < 3; ++i) P.t[i] = tp[i];
  return P;
}

extern "C" {

int rgbid_version(void) { return RGBID_B200_VERSION; }

const char* rgbid_status_string(int status)
{
  switch (status) {
    case RGBID_OK: return "ok";
    case RGBID_ERR_NAN: return "numerical failure (NaN pose)";
    case RGBID_ERR_ARG: return "bad argument";
    case RGBID_ERR_NOMEM: return "out of memory";
    case RGBID_ERR_STATE: return "bad call sequence";
    case RGBID_ERR_TIMEOUT: return "device wait timed out";
    default: break;
  }
  if (status >= RGBID_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(status - RGBID_ERR_CUDA_BASE));
  return "unknown status";
}

int rgbid_ctx_create(rgbid_ctx** out, int device, void* stream)
{
  if (!out) return RGBID_ERR_ARG;
  *out = nullptr;
  RGBID_CUDA_TRY(cudaSetDevice(device));
  rgbid_ctx* ctx = new rgbid_ctx();
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  ctx->num_sms = sms > 0 ? sms : kNumSMsB200;
  if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
  else {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return RGBID_ERR_CUDA_BASE + (int)e; }
    ctx->own_stream = true;
  }
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_partials, sizeof(double) * 32 * (size_t)(ctx->num_sms + 8));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_counter, sizeof(unsigned int) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_counts, sizeof(unsigned int) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_out, sizeof(double) * 32);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_scale, sizeof(ScaleState));
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_small, 4096);
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int) * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { rgbid_ctx_destroy(ctx); return RGBID_ERR_CUDA_BASE + (int)e; }
  *out = ctx;
  return RGBID_OK;
}

int rgbid_ctx_destroy(rgbid_ctx* ctx)
{
  if (!ctx) return RGBID_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_partials); cudaFree(ctx->d_counter); cudaFree(ctx->d_counts); cudaFree(ctx->d_out);
  cudaFree(ctx->d_scale); cudaFree(ctx->d_stage);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return RGBID_OK;
}

int rgbid_ctx_sync(rgbid_ctx* ctx)
{
  if (!ctx) return RGBID_ERR_ARG;
  RGBID_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return check_last(ctx);
}

void* rgbid_ctx_stream(rgbid_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long long rgbid_ctx_launch_count(rgbid_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- image preparation ----------------------------------------------------------------------------
int rgbid_convert_depth_to_invdepth(rgbid_ctx* ctx, const uint16_t* src, size_t spitch, float* dst, size_t dpitch,
                                    int rows, int cols, float factor_depth)
{
  if (!ctx || !src || !dst || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_depth_to_invdepth(ctx->L(), src, spitch, 0, make_img(dst, dpitch, rows, cols), 1, factor_depth);
  return check_last(ctx);
}

int rgbid_compute_intensity(rgbid_ctx* ctx, const uint8_t* rgb, size_t spitch, float* dst, size_t dpitch, int rows,
                            int cols)
{
  if (!ctx || !rgb || !dst || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_intensity(ctx->L(), rgb, spitch, 0, make_img(dst, dpitch, rows, cols), 1);
  return check_last(ctx);
}

int rgbid_decompose_rgb(rgbid_ctx* ctx, const uint8_t* rgb, size_t spitch, float* r, float* g, float* b, size_t dpitch,
                        int rows, int cols)
{
  if (!ctx || !rgb || !r || !g || !b || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_decompose_rgb(ctx->L(), rgb, spitch, make_img(r, dpitch, rows, cols), make_img(g, dpitch, rows, cols),
                       make_img(b, dpitch, rows, cols));
  return check_last(ctx);
}

int rgbid_pyr_down(rgbid_ctx* ctx, const float* src, size_t spitch, int srows, int scols, float* dst, size_t dpitch)
{
  if (!ctx || !src || !dst || srows < 2 || scols < 2) return RGBID_ERR_ARG;
  ImgB none = make_img(nullptr, 0, 0, 0);
  launch_pyr_down2(ctx->L(), make_img(src, spitch, srows, scols), make_img(dst, dpitch, srows / 2, scols / 2), none,
                   none, 1);
  return check_last(ctx);
}

int rgbid_compute_gradient(rgbid_ctx* ctx, const float* src, size_t spitch, int rows, int cols, float* gx, float* gy,
                           size_t gpitch)
{
  if (!ctx || !src || !gx || !gy || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  ImgB none = make_img(nullptr, 0, 0, 0);
  launch_gradient2(ctx->L(), make_img(src, spitch, rows, cols), make_img(gx, gpitch, rows, cols),
                   make_img(gy, gpitch, rows, cols), none, none, none, 1);
  return check_last(ctx);
}

int rgbid_bilateral_filter(rgbid_ctx* ctx, const float* src, size_t spitch, int rows, int cols, float* dst,
                           size_t dpitch, float sigma_floatmap)
{
  if (!ctx || !src || !dst || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  ImgB none = make_img(nullptr, 0, 0, 0);
  launch_bilateral2(ctx->L(), make_img(src, spitch, rows, cols), make_img(dst, dpitch, rows, cols), sigma_floatmap,
                    none, none, 0.f, 1);
  return check_last(ctx);
}

int rgbid_copy_image(rgbid_ctx* ctx, const float* src, size_t spitch, float* dst, size_t dpitch, int rows, int cols)
{
  if (!ctx || !src || !dst || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  RGBID_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, (size_t)cols * sizeof(float), rows,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
  return RGBID_OK;
}

int rgbid_fill_image(rgbid_ctx* ctx, float* dst, size_t dpitch, int rows, int cols, float value)
{
  if (!ctx || !dst || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_fill(ctx->L(), make_img(dst, dpitch, rows, cols), value, 1);
  return check_last(ctx);
}

int rgbid_create_vmap(rgbid_ctx* ctx, const float* depth_inv, size_t pitch, int rows, int cols, float fx, float fy,
                      float cx, float cy, float* vmap, size_t vpitch)
{
  if (!ctx || !depth_inv || !vmap || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_vmap(ctx->L(), make_img(depth_inv, pitch, rows, cols), make_img(vmap, vpitch, 3 * rows, cols), fx, fy, cx, cy, 1);
  return check_last(ctx);
}

int rgbid_create_nmap_gradients(rgbid_ctx* ctx, const float* depth_inv, const float* gx, const float* gy, size_t pitch,
                                int rows, int cols, float fx, float fy, float cx, float cy, float* nmap, size_t npitch)
{
  if (!ctx || !depth_inv || !gx || !gy || !nmap || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_nmap_gradients(ctx->L(), make_img(depth_inv, pitch, rows, cols), make_img(gx, pitch, rows, cols),
                        make_img(gy, pitch, rows, cols), make_img(nmap, npitch, 3 * rows, cols), fx, fy, cx, cy, 1);
  return check_last(ctx);
}

// ---- warping / visibility / fusion ------------------------------------------------------------------
int rgbid_warp_invdepth(rgbid_ctx* ctx, const float* src, size_t spitch, const float* prev, size_t ppitch, float* dst,
                        size_t dpitch, int rows, int cols, const float* Rp, const float* tp)
{
  if (!ctx || !src || !prev || !dst || !Rp || !tp || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_warp_invdepth(ctx->L(), make_img(src, spitch, rows, cols), make_img(prev, ppitch, rows, cols),
                       make_img(dst, dpitch, rows, cols), make_proj(Rp, tp));
  return check_last(ctx);
}

int rgbid_warp_intensity(rgbid_ctx* ctx, const float* src, size_t spitch, const float* prev, size_t ppitch, float* dst,
                         size_t dpitch, int rows, int cols, const float* Rp, const float* tp)
{
  if (!ctx || !src || !prev || !dst || !Rp || !tp || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_warp_intensity(ctx->L(), make_img(src, spitch, rows, cols), make_img(prev, ppitch, rows, cols),
                        make_img(dst, dpitch, rows, cols), make_proj(Rp, tp));
  return check_last(ctx);
}

int rgbid_warp_invdepth_weighted(rgbid_ctx* ctx, const float* src, size_t spitch, const float* prev, size_t ppitch,
                                 float* dst, size_t dpitch, float* weight, size_t wpitch, int rows, int cols,
                                 const float* Rp, const float* tp)
{
  if (!ctx || !src || !prev || !dst || !weight || !Rp || !tp || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_warp_invdepth_weighted(ctx->L(), make_img(src, spitch, rows, cols), make_img(prev, ppitch, rows, cols),
                                make_img(dst, dpitch, rows, cols), make_img(weight, wpitch, rows, cols), nullptr,
                                make_proj(Rp, tp), 1);
  return check_last(ctx);
}

int rgbid_integrate_warped_frame(rgbid_ctx* ctx, const float* wd, size_t wd_pitch, const float* ww, size_t ww_pitch,
                                 float* dd, size_t dd_pitch, float* dw, size_t dw_pitch, int rows, int cols)
{
  if (!ctx || !wd || !ww || !dd || !dw || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_integrate(ctx->L(), make_img(wd, wd_pitch, rows, cols), make_img(ww, ww_pitch, rows, cols),
                   make_img(dd, dd_pitch, rows, cols), make_img(dw, dw_pitch, rows, cols), 1);
  return check_last(ctx);
}

int rgbid_visibility_ratio(rgbid_ctx* ctx, const float* dsrc, size_t spitch, const float* ddst, size_t dpitch,
                           int rows, int cols, const float* Rp, const float* tp, uint8_t* mask, size_t mpitch,
                           float* ratio_host)
{
  if (!ctx || !dsrc || !ddst || !Rp || !tp || !ratio_host || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  RGBID_CUDA_TRY(cudaMemsetAsync(ctx->d_counts, 0, sizeof(unsigned int) * 2, ctx->stream));
  launch_visibility(ctx->L(), make_img(dsrc, spitch, rows, cols), make_img(ddst, dpitch, rows, cols), nullptr,
                    make_proj(Rp, tp), ctx->d_counts, 0, 2, mask, mpitch, 0, 1);
  unsigned int* h = (unsigned int*)ctx->h_small;
  RGBID_CUDA_TRY(cudaMemcpyAsync(h, ctx->d_counts, sizeof(unsigned int) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  RGBID_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  // ratio logic of warping_registration.cu:863-866
  float vis = (float)h[0], val = (float)h[1];
  *ratio_host = (val < 1.f) ? 0.f : vis / val;
  return check_last(ctx);
}

// ---- custom-calibration ingest, colour fusion, previews (calib_ops.cu) ------------------------------------
int rgbid_undistort_intensity(rgbid_ctx* ctx, const float* src, size_t spitch, float* dst, size_t dpitch, int rows, int cols,
                              const rgbid_intr* intr)
{
  if (!ctx || !src || !dst || !intr || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_undistort_intensity(ctx->L(), make_img(src, spitch, rows, cols), make_img(dst, dpitch, rows, cols), *intr);
  return check_last(ctx);
}

int rgbid_undistort_depthinv(rgbid_ctx* ctx, const float* src, size_t spitch, float* dst, size_t dpitch, int rows, int cols,
                             const rgbid_intr* intr, const rgbid_depth_dist* dp)
{
  if (!ctx || !src || !dst || !intr || !dp || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_undistort_depthinv(ctx->L(), make_img(src, spitch, rows, cols), make_img(dst, dpitch, rows, cols), *intr, *dp);
  return check_last(ctx);
}

int rgbid_register_depthinv(rgbid_ctx* ctx, const float* src, size_t spitch, float* dst, size_t dpitch, int rows, int cols,
                            const float* dRc_proj, const float* t_dc_proj, const float* cRd_proj)
{
  if (!ctx || !src || !dst || !dRc_proj || !t_dc_proj || !cRd_proj || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  const int crows = 3 * rows, ccols = 3 * cols;  // src/visodo.cpp:623-624
  const size_t cpitch = align_up((size_t)ccols * sizeof(int), 128);
  int rc = ctx_reserve_device_stage(ctx, cpitch * crows);
  if (rc != RGBID_OK) return rc;
  launch_register_depthinv(ctx->L(), make_img(src, spitch, rows, cols), make_img(dst, dpitch, rows, cols), (int*)ctx->d_stage,
                           cpitch, crows, ccols, dRc_proj, t_dc_proj, cRd_proj);
  return check_last(ctx);
}

int rgbid_integrate_warped_rgb(rgbid_ctx* ctx, const float* depth_warped, const float* r_warped, const float* g_warped,
                               const float* b_warped, const float* weight_warped, float* depth_dst, uint8_t* colors_dst,
                               size_t colors_pitch, float* weight_dst, size_t pitch, int rows, int cols)
{
  if (!ctx || !depth_warped || !r_warped || !g_warped || !b_warped || !weight_warped || !depth_dst || !colors_dst ||
      !weight_dst || rows <= 0 || cols <= 0)
    return RGBID_ERR_ARG;
  auto im = [&](const float* p) { return make_img(p, pitch, rows, cols); };
  launch_integrate_rgb(ctx->L(), im(depth_warped), im(r_warped), im(g_warped), im(b_warped), im(weight_warped), im(depth_dst),
                       colors_dst, colors_pitch, im(weight_dst));
  return check_last(ctx);
}

int rgbid_generate_image(rgbid_ctx* ctx, const float* vmap, const float* nmap, size_t map_pitch, const uint8_t* rgb,
                         size_t rgb_pitch, const float* light_pos, uint8_t* out, size_t out_pitch, int rows, int cols)
{
  if (!ctx || !vmap || !nmap || !light_pos || !out || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  launch_generate_image(ctx->L(), make_img(vmap, map_pitch, 3 * rows, cols), make_img(nmap, map_pitch, 3 * rows, cols), rgb,
                        rgb_pitch, light_pos, out, out_pitch, rows, cols);
  return check_last(ctx);
}

// ---- residual sampling / scale / chi-square -----------------------------------------------------------
int rgbid_error_geometry(int rows, int cols, int min_nsamples, int* kept_rows, int* kept_cols, int* stride)
{
  if (rows <= 0 || cols <= 0 || !kept_rows || !kept_cols || !stride) return RGBID_ERR_ARG;
  // computeErrorGridStride, sigmaFuncs.cu:711-747
  int error_size = cols * rows, cp = cols, rp = rows;
  if (min_nsamples < error_size) {
    for (;;) {
      int cc = cp / 2, rc = rp / 2;
      if ((2 * cc - cp) != 0 || (2 * rc - rp) != 0 || min_nsamples > cc * rc) { error_size = cp * rp; break; }
      cp = cc; rp = rc;
    }
  }
  *kept_rows = rp; *kept_cols = cp;
  *stride = (int)std::sqrt((double)((rows * cols) / error_size));
  return RGBID_OK;
}

int rgbid_compute_error(rgbid_ctx* ctx, const float* im1, size_t pitch1, const float* im0, size_t pitch0, int rows,
                        int cols, int min_nsamples, float* error, int* n_out)
{
  if (!ctx || !im1 || !im0 || !error || rows <= 0 || cols <= 0) return RGBID_ERR_ARG;
  int kr, kc, s;
  rgbid_error_geometry(rows, cols, min_nsamples, &kr, &kc, &s);
  launch_compute_error(ctx->L(), make_img(im1, pitch1, rows, cols), make_img(im0, pitch0, rows, cols), error, kr, kc, s);
  if (n_out) *n_out = kr * kc;
  return check_last(ctx);
}

static int run_scale(rgbid_ctx* ctx, const float* error, int n, int op, int mest, float bias, float sigma,
                     ScaleState* host_out)
{
  launch_scale_from_errors(ctx->L(), error, nullptr, n, op, mest, bias, sigma, 0.f, 1.f, ctx->d_scale);
  RGBID_CUDA_TRY(cudaMemcpyAsync(ctx->h_small, ctx->d_scale, sizeof(ScaleState), cudaMemcpyDeviceToHost, ctx->stream));
  RGBID_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *host_out = *(ScaleState*)ctx->h_small;
  return check_last(ctx);
}

int rgbid_sigma_nu_student(rgbid_ctx* ctx, const float* error, int n, float* bias, float* sigma, float* nu,
                           int mestimator)
{
  if (!ctx || !error || n <= 0 || !bias || !sigma || !nu) return RGBID_ERR_ARG;
  ScaleState st;
  int rc = run_scale(ctx, error, n, SCALE_SIGMA_NU, mestimator, *bias, *sigma, &st);
  if (rc != RGBID_OK) return rc;
  *bias = st.bias_int; *sigma = st.sigma_int; *nu = st.nu_int;
  return RGBID_OK;
}

int rgbid_nu_student(rgbid_ctx* ctx, const float* error, int n, float bias, float sigma, float* nu)
{
  if (!ctx || !error || n <= 0 || !nu) return RGBID_ERR_ARG;
  ScaleState st;
  int rc = run_scale(ctx, error, n, SCALE_NU_ONLY, RGBID_STUDENT, bias, sigma, &st);
  if (rc != RGBID_OK) return rc;
  *nu = st.nu_int;
  return RGBID_OK;
}

int rgbid_sigma_pdf(rgbid_ctx* ctx, const float* error, int n, float* bias, float* sigma, int mestimator)
{
  if (!ctx || !error || n <= 0 || !bias || !sigma) return RGBID_ERR_ARG;
  ScaleState st;
  int rc = run_scale(ctx, error, n, SCALE_SIGMA_PDF, mestimator, *bias, *sigma, &st);
  if (rc != RGBID_OK) return rc;
  *bias = st.bias_int; *sigma = st.sigma_int;
  return RGBID_OK;
}

int rgbid_chi_square(rgbid_ctx* ctx, const float* err_int, const float* err_depth, int n, float sigma_int,
                     float sigma_depth, int mestimator, float* chi_square, float* chi_test, float* ndof)
{
  if (!ctx || !err_int || !err_depth || n <= 0 || !chi_square || !chi_test || !ndof) return RGBID_ERR_ARG;
  launch_chi_square(ctx->L(), err_int, err_depth, n, sigma_int, sigma_depth, mestimator, ctx->d_out);
  double* h = (double*)ctx->h_small;
  RGBID_CUDA_TRY(cudaMemcpyAsync(h, ctx->d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RGBID_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  // host part of computeChiSquare, sigmaFuncs.cu:1279-1287
  float fn = (float)h[1];
  *chi_square = (float)h[0] / fn;
  *ndof = fn;
  float z = (*chi_square - fn) / std::sqrt(2.f * fn);
  *chi_test = 0.5f * (1.f + std::erf(z / std::sqrt(2.f)));
  return check_last(ctx);
}

// ---- normal equations -----------------------------------------------------------------------------
int rgbid_build_system(rgbid_ctx* ctx, const float* W0, const float* I0, const float* gWx, const float* gWy,
                       const float* gIx, const float* gIy, const float* W1, const float* I1, size_t pitch, int rows,
                       int cols, const rgbid_system_params* sp, double* A36, double* b6)
{
  if (!ctx || !W0 || !I0 || !gWx || !gWy || !gIx || !gIy || !W1 || !I1 || !sp || !A36 || !b6 || rows <= 0 || cols <= 0)
    return RGBID_ERR_ARG;
  launch_build_system(ctx->L(), make_img(W0, pitch, rows, cols), make_img(I0, pitch, rows, cols),
                      make_img(gWx, pitch, rows, cols), make_img(gWy, pitch, rows, cols),
                      make_img(gIx, pitch, rows, cols), make_img(gIy, pitch, rows, cols),
                      make_img(W1, pitch, rows, cols), make_img(I1, pitch, rows, cols), *sp, ctx->d_partials,
                      ctx->d_counter, ctx->d_out);
  double* h = (double*)ctx->h_small;
  RGBID_CUDA_TRY(cudaMemcpyAsync(h, ctx->d_out, 27 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RGBID_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  unpack_system(h, A36, b6);
  return check_last(ctx);
}

/* Same with one row pitch per map, in the order of the pointer arguments (a PtrStep carries its own step:
 * cudaMallocPitch gives the maps a caller allocates a different pitch from the ones it wraps) */
int rgbid_build_system_pitched(rgbid_ctx* ctx, const float* W0, const float* I0, const float* gWx, const float* gWy,
                               const float* gIx, const float* gIy, const float* W1, const float* I1, const size_t* pitch8,
                               int rows, int cols, const rgbid_system_params* sp, double* A36, double* b6)
{
  if (!ctx || !W0 || !I0 || !gWx || !gWy || !gIx || !gIy || !W1 || !I1 || !pitch8 || !sp || !A36 || !b6 || rows <= 0 || cols <= 0)
    return RGBID_ERR_ARG;
  for (int i = 0; i < 8; ++i)
    if (pitch8[i] < (size_t)cols * sizeof(float)) return RGBID_ERR_ARG;
  launch_build_system(ctx->L(), make_img(W0, pitch8[0], rows, cols), make_img(I0, pitch8[1], rows, cols),
                      make_img(gWx, pitch8[2], rows, cols), make_img(gWy, pitch8[3], rows, cols),
                      make_img(gIx, pitch8[4], rows, cols), make_img(gIy, pitch8[5], rows, cols),
                      make_img(W1, pitch8[6], rows, cols), make_img(I1, pitch8[7], rows, cols), *sp, ctx->d_partials,
                      ctx->d_counter, ctx->d_out);
  double* h = (double*)ctx->h_small;
  RGBID_CUDA_TRY(cudaMemcpyAsync(h, ctx->d_out, 27 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RGBID_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  unpack_system(h, A36, b6);
  return check_last(ctx);
}

}  // extern "C"
