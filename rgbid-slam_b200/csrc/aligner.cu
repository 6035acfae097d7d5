// aligner.cu -- host orchestration of the fused, device-resident coarse-to-fine alignment.
//
// The reference's drivers (VisodoTracker::estimateVisualOdometry, src/visodo.cpp:944-1479;
// KeyframeAlign::alignKeyframes, src/keyframe_align.cpp:115-357) cross the host/device boundary ~40 times
// per Gauss-Newton iteration.  Here the complete schedule -- every (level, iteration), scale estimation,
// normal equations, solve, pose update and the covariance pass -- is recorded once as a CUDA graph and
// replayed per call; the host only writes the initial guess and reads the final state.
#include <cstdlib>
#include <cstring>
#include <new>

#include "aligner.hpp"

using namespace rgbid;

namespace rgbid {

static ImgB sub(const ImgB& m, int first)
{
  ImgB v = m;
  v.p = (float*)((char*)m.p + (size_t)first * m.sstride);
  return v;
}

void aligner_current_pyramid(rgbid_aligner* al, int first, int batch)
{
  LaunchCtx L = al->ctx->L();
  for (int l = 1; l < al->cfg.levels; ++l)  // prepareImages, src/visodo.cpp:768-772
    launch_pyr_down2(L, sub(al->maps[MAP_I_CUR][l - 1], first), sub(al->maps[MAP_I_CUR][l], first),
                     sub(al->maps[MAP_W_CUR][l - 1], first), sub(al->maps[MAP_W_CUR][l], first), batch);
}

void aligner_copy_current_to_keyframe(rgbid_aligner* al, int first, int batch, const int* active)
{
  LaunchCtx L = al->ctx->L();
  {
    // all levels of both maps in one launch
    ImgB src[2 * RGBID_MAX_LEVELS], dst[2 * RGBID_MAX_LEVELS];
    int n = 0;
    for (int l = 0; l < al->cfg.levels; ++l) {
      src[n] = sub(al->maps[MAP_W_CUR][l], first); dst[n++] = sub(al->maps[MAP_W_KF][l], first);
      src[n] = sub(al->maps[MAP_I_CUR][l], first); dst[n++] = sub(al->maps[MAP_I_KF][l], first);
    }
    if (launch_copy_list(L, src, dst, n, batch, active)) return;
  }
  for (int l = 0; l < al->cfg.levels; ++l)  // copyImages per level, src/visodo.cpp:837-840
    launch_copy2(L, sub(al->maps[MAP_W_CUR][l], first), sub(al->maps[MAP_W_KF][l], first),
                 sub(al->maps[MAP_I_CUR][l], first), sub(al->maps[MAP_I_KF][l], first), batch, active);
}

void aligner_keyframe_derivatives(rgbid_aligner* al, int first, int batch, const int* active, bool pyramid_from_l0)
{
  LaunchCtx L = al->ctx->L();
  const int levels = al->cfg.levels;
  auto M = [&](int which, int l) { return sub(al->maps[which][l], first); };
  if (pyramid_from_l0)  // keyframe_align.cpp:157-164
    for (int l = 1; l < levels; ++l)
      launch_pyr_down2(L, M(MAP_W_KF, l - 1), M(MAP_W_KF, l), M(MAP_I_KF, l - 1), M(MAP_I_KF, l), batch, active);
  const bool tracker = (al->cfg.mode == RGBID_MODE_TRACKER);
  if (tracker) {
    // saveCurrentImagesAsOdoKeyframes, src/visodo.cpp:842-857: bilateral filter (range sigma 2*0.0025 for the
    // inverse depth, 3 for the intensity), pyramid of the filtered maps, Sobel -> covariance-only gradients
    launch_bilateral2(L, M(MAP_W_KF, 0), M(MAP_WF, 0), 2.f * 0.0025f, M(MAP_I_KF, 0), M(MAP_IF, 0), 3.f, batch, active);
    for (int l = 1; l < levels; ++l)
      launch_pyr_down2(L, M(MAP_WF, l - 1), M(MAP_WF, l), M(MAP_IF, l - 1), M(MAP_IF, l), batch, active);
    if (al->image_filtering != RGBID_FILTER_GRADS && 4 * levels <= 16) {
      // the Sobel passes of every level of the filtered AND of the raw pyramid in one launch
      ImgB src[16], gx[16], gy[16];
      int n = 0;
      for (int l = 0; l < levels; ++l) {
        src[n] = M(MAP_IF, l); gx[n] = M(MAP_CGIX, l); gy[n++] = M(MAP_CGIY, l);
        src[n] = M(MAP_WF, l); gx[n] = M(MAP_CGWX, l); gy[n++] = M(MAP_CGWY, l);
        src[n] = M(MAP_I_KF, l); gx[n] = M(MAP_GIX, l); gy[n++] = M(MAP_GIY, l);
        src[n] = M(MAP_W_KF, l); gx[n] = M(MAP_GWX, l); gy[n++] = M(MAP_GWY, l);
      }
      if (launch_gradient_list(L, src, gx, gy, n, batch, active)) return;
    }
    for (int l = 0; l < levels; ++l)
      launch_gradient2(L, M(MAP_IF, l), M(MAP_CGIX, l), M(MAP_CGIY, l), M(MAP_WF, l), M(MAP_CGWX, l), M(MAP_CGWY, l),
                       batch, active);
  }
  if (tracker && al->image_filtering == RGBID_FILTER_GRADS) {
    for (int l = 0; l < levels; ++l) {  // src/visodo.cpp:859-866
      launch_copy2(L, M(MAP_CGIX, l), M(MAP_GIX, l), M(MAP_CGIY, l), M(MAP_GIY, l), batch, active);
      launch_copy2(L, M(MAP_CGWX, l), M(MAP_GWX, l), M(MAP_CGWY, l), M(MAP_GWY, l), batch, active);
    }
  } else {
    if (2 * levels <= 16) {
      ImgB src[16], gx[16], gy[16];
      int n = 0;
      for (int l = 0; l < levels; ++l) {
        src[n] = M(MAP_I_KF, l); gx[n] = M(MAP_GIX, l); gy[n++] = M(MAP_GIY, l);
        src[n] = M(MAP_W_KF, l); gx[n] = M(MAP_GWX, l); gy[n++] = M(MAP_GWY, l);
      }
      if (launch_gradient_list(L, src, gx, gy, n, batch, active)) return;
    }
    for (int l = 0; l < levels; ++l)  // src/visodo.cpp:869-877, keyframe_align.cpp:168-176
      launch_gradient2(L, M(MAP_I_KF, l), M(MAP_GIX, l), M(MAP_GIY, l), M(MAP_W_KF, l), M(MAP_GWX, l), M(MAP_GWY, l),
                       batch, active);
  }
}

static GnParams base_params(const rgbid_aligner* al, int level)
{
  const rgbid_align_config& c = al->cfg;
  GnParams P;
  memset(&P, 0, sizeof(P));
  P.batch = c.batch; P.level = level;
  P.rows = al->geom[level].rows; P.cols = al->geom[level].cols;
  float div = (float)(1 << level);  // Intr::operator()(level), src/internal.h:128-132
  P.fx = c.fx / div; P.fy = c.fy / div; P.cx = c.cx / div; P.cy = c.cy / div;
  P.fx0 = c.fx; P.fy0 = c.fy; P.cx0 = c.cx; P.cy0 = c.cy;
  P.levels = c.levels; P.mode = c.mode;
  P.mestimator = c.mestimator; P.weighting = c.weighting;
  P.student_nu = 1; P.use_scale = 1; P.update_pose = 1; P.compute_cov = 0; P.chi_mestimator = -1;
  P.trace_stride = al->trace_stride;
  P.kept_rows = al->geom[level].kept_rows; P.kept_cols = al->geom[level].kept_cols;
  P.sample_stride = al->geom[level].sample_stride;
  P.sigma_op = (c.mode == RGBID_MODE_TRACKER) ? SCALE_SIGMA_NU : SCALE_NU_ONLY;
  P.next_level = -1;
  P.sched_level = level; P.sched_iter = 0;
  P.termination = c.termination;
  P.conv_eps = (c.termination == RGBID_TERM_CONVERGENCE) ? c.conv_eps : 0.f;
  return P;
}

static GnLevelMaps level_maps(const rgbid_aligner* al, int level, bool cov_gradients)
{
  GnLevelMaps M;
  M.W0 = al->maps[MAP_W_KF][level]; M.I0 = al->maps[MAP_I_KF][level];
  M.gWx = al->maps[cov_gradients ? MAP_CGWX : MAP_GWX][level];
  M.gWy = al->maps[cov_gradients ? MAP_CGWY : MAP_GWY][level];
  M.gIx = al->maps[cov_gradients ? MAP_CGIX : MAP_GIX][level];
  M.gIy = al->maps[cov_gradients ? MAP_CGIY : MAP_GIY][level];
  M.Wc = al->maps[MAP_W_CUR][level]; M.Ic = al->maps[MAP_I_CUR][level];
  const int B = al->cfg.batch;
  M.texW = al->use_tex ? al->d_tex + (size_t)(level * 2 + 0) * B : nullptr;
  M.texI = al->use_tex ? al->d_tex + (size_t)(level * 2 + 1) * B : nullptr;
  M.tex_border = al->use_tex ? 1 : 0;
  return M;
}

// Issues the launches of one complete alignment on the context's stream (captured into a graph when enabled).
//
// Optionally (RGBID_CHAINS=2) the batch is issued as two independent groups of frame pairs on two streams (two parallel
// branches of the graph), the second staggered by one scale estimation, so that one group's FMA-bound system kernel
// could run under the other group's latency-bound scale estimation and 6x6 solve.  Measured on B200 (32 streams):
// 2.98 ms per step against 2.82 ms for the single chain -- the two kernels do not share an SM (different shared-memory
// carve-outs), so the chains only interleave at kernel granularity and pay twice the launches.  Off by default.
static void record_chain(rgbid_aligner* al, LaunchCtx L, int first, int count, cudaEvent_t after_first_launch = nullptr,
                         int part = ALIGN_PART_ALL)
{
  bool signalled = (after_first_launch == nullptr);
  const rgbid_align_config& c = al->cfg;
  L.pdl = al->use_pdl;  // scale and system kernels overlap their prologues with the predecessor's tail (kernels.cuh)
  const bool tracker = (c.mode == RGBID_MODE_TRACKER);
  const bool estimate_scale = tracker ? (c.sigma_estimator == RGBID_SIGMA_PDF) : true;
  const bool warp_first = tracker && c.warp_first;  // KeyframeAlign has no such option (src/keyframe_align.cpp:178-350)
  int done = 0;
  for (int level = c.levels - 1; level >= c.finest_level && part != ALIGN_PART_COV; --level) {
    for (int it = 0; it < c.iterations[level]; ++it) {
      GnParams P = base_params(al, level);
      P.first = first; P.batch = count; P.batch_total = c.batch;
      P.iter_index = done;
      P.sched_iter = it;
      if (tracker && c.termination == RGBID_TERM_CHI_SQUARED && it > 0 && (warp_first || level == 0)) {
        // cost-function test that ends the iterations of this level (src/visodo.cpp:1134-1164): robust chi^2 of ALL
        // level-0 residuals at the current pose -- the fused kernel at level 0 with the chi^2 sums switched on and
        // nothing else to do; its tail compares the RMSE with the previous test's and undoes the last increment
        GnParams T = base_params(al, 0);
        T.first = first; T.batch = count; T.batch_total = c.batch;
        T.iter_index = -1; T.sched_level = level; T.sched_iter = it;
        T.use_scale = 0; T.student_nu = 0; T.mestimator = RGBID_STUDENT;
        T.update_pose = 0; T.compute_cov = 0; T.chi_mestimator = c.mestimator;
        T.chi_test = (it == 1) ? 1 : 2;
        launch_gn_build(L, level_maps(al, 0, false), T, al->d_states, nullptr, al->d_partials, 32, al->d_counters, nullptr);
      }
      // the updated pose is consumed at this level again, at the next coarser-to-finer level that has iterations,
      // or by the covariance pass at the finest level
      P.next_level = level;
      if (it + 1 == c.iterations[level]) {
        P.next_level = c.finest_level;
        for (int l = level - 1; l >= c.finest_level; --l)
          if (c.iterations[l] > 0) { P.next_level = l; break; }
      }
      P.use_scale = estimate_scale ? 1 : 0;
      ++done;
      // KeyframeAlign: covariance = inverse of the LAST iteration's A (keyframe_align.cpp:339-350)
      P.compute_cov = (!tracker && done == al->niters) ? 1 : 0;
      GnLevelMaps M = level_maps(al, level, false);
      if (warp_first && level > 0) {
        // WARP_ORDER = warpFirst (src/visodo.cpp:1078-1105): every iteration warps the current frame at level 0 with the
        // current pose and rebuilds the pyramid of the warped maps down to this level; the scale and system kernels
        // then read those maps pixel for pixel.  At level 0 the two warp orders are the same computation.
        const int B = c.batch;
        launch_warp_pair(L, al->maps[MAP_W_CUR][0], al->maps[MAP_I_CUR][0], al->use_tex ? al->d_tex : nullptr,
                         al->use_tex ? al->d_tex + B : nullptr, al->maps[MAP_W_KF][0], al->d_states,
                         al->maps[MAP_W_WARP][0], al->maps[MAP_I_WARP][0], first, count);
        for (int l = 1; l <= level; ++l)
          launch_pyr_down2(L, sub(al->maps[MAP_I_WARP][l - 1], first), sub(al->maps[MAP_I_WARP][l], first),
                           sub(al->maps[MAP_W_WARP][l - 1], first), sub(al->maps[MAP_W_WARP][l], first), count);
        M.Wc = al->maps[MAP_W_WARP][level]; M.Ic = al->maps[MAP_I_WARP][level];
        M.texW = nullptr; M.texI = nullptr; M.tex_border = 0;
        P.prewarped = 1;
      }
      // the level-0 projection is needed whatever level iterates next (warpFirst: level-0 warp every iteration;
      // CHI_SQUARED: the test runs on level 0)
      if (warp_first || (tracker && c.termination == RGBID_TERM_CHI_SQUARED)) P.next_level = -1;
      if (estimate_scale) launch_gn_scale(L, M, P, al->d_states, al->d_scales);
      if (!signalled) { cudaEventRecord(after_first_launch, L.stream); signalled = true; }
      launch_gn_build(L, M, P, al->d_states, al->d_scales, al->d_partials, 32, al->d_counters, al->d_trace);
    }
  }
  if (tracker && part != ALIGN_PART_ITERATIONS) {
    // covariance pass at the finest level on the bilateral-filtered gradients with fixed scales and
    // Student(5) weights, no pose update (src/visodo.cpp:1283-1409) + end-of-frame chi^2 (:1411-1415)
    GnParams P = base_params(al, c.finest_level);
    P.first = first; P.batch = count; P.batch_total = c.batch;
    P.iter_index = al->niters;
    P.use_scale = 0; P.student_nu = 0; P.mestimator = RGBID_STUDENT;
    P.update_pose = 0; P.compute_cov = 1; P.chi_mestimator = c.mestimator;
    GnLevelMaps M = level_maps(al, c.finest_level, true);
    launch_gn_build(L, M, P, al->d_states, nullptr, al->d_partials, 32, al->d_counters, al->d_trace);
  }
}

void aligner_record_schedule(rgbid_aligner* al, int part)
{
  const rgbid_align_config& c = al->cfg;
  LaunchCtx L = al->ctx->L();
  if (part != ALIGN_PART_COV)
    launch_gn_init(L, al->d_states, al->d_init, al->d_init + 9 * c.batch, c.batch, c.levels, c.fx, c.fy, c.cx, c.cy,
                   al->d_trace_flag);
  if (al->side_stream == nullptr || c.batch < 2) {
    record_chain(al, L, 0, c.batch, nullptr, part);
    return;
  }
  // (the two-chain experiment is only ever recorded whole: aligner_can_split() is false with a side stream)
  const int half = c.batch / 2;
  LaunchCtx L2 = L;
  L2.stream = al->side_stream;
  // the second chain starts when the first has finished its first scale estimation: the two chains then alternate
  // (one in its latency-bound phase, the other in its FMA-bound phase) instead of marching in lock step
  record_chain(al, L, 0, half, al->ev_fork);
  cudaStreamWaitEvent(al->side_stream, al->ev_fork, 0);
  record_chain(al, L2, half, c.batch - half);
  cudaEventRecord(al->ev_join, al->side_stream);
  cudaStreamWaitEvent(L.stream, al->ev_join, 0);
}

}  // namespace rgbid

static int copy_in(rgbid_aligner* al, const void* src, size_t spitch, ImgB dst, size_t width_bytes, int from_host)
{
  RGBID_CUDA_TRY(cudaMemcpy2DAsync(dst.p, dst.pitch, src, spitch, width_bytes, dst.rows,
                                   from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, al->ctx->stream));
  return RGBID_OK;
}

extern "C" {

int rgbid_aligner_create(rgbid_ctx* ctx, const rgbid_align_config* cfg, rgbid_aligner** out)
{
  if (!ctx || !cfg || !out) return RGBID_ERR_ARG;
  *out = nullptr;
  if (cfg->levels < 1 || cfg->levels > RGBID_MAX_LEVELS || cfg->batch < 1 || cfg->rows <= 0 || cfg->cols <= 0)
    return RGBID_ERR_ARG;
  if (cfg->finest_level < 0 || cfg->finest_level >= cfg->levels) return RGBID_ERR_ARG;
  if ((cfg->rows % (1 << (cfg->levels - 1))) || (cfg->cols % (1 << (cfg->levels - 1)))) return RGBID_ERR_ARG;
  if (!(cfg->fx > 0.f) || !(cfg->fy > 0.f)) return RGBID_ERR_ARG;
  if (cfg->termination < RGBID_TERM_ALL_ITERS || cfg->termination > RGBID_TERM_CONVERGENCE) return RGBID_ERR_ARG;
  if (cfg->termination == RGBID_TERM_CONVERGENCE && !(cfg->conv_eps > 0.f)) return RGBID_ERR_ARG;
  RGBID_CUDA_TRY(cudaSetDevice(ctx->device));
  // constant tables and kernel attributes now, outside any stream capture (the upload is a synchronous legacy-stream copy)
  if (int e = gn_prepare_device()) return RGBID_ERR_CUDA_BASE + e;
  rgbid_aligner* al = new (std::nothrow) rgbid_aligner();
  if (!al) return RGBID_ERR_NOMEM;
  memset(al, 0, sizeof(*al));
  al->ctx = ctx;
  al->cfg = *cfg;
  if (al->cfg.factor_depth <= 0.f) al->cfg.factor_depth = 1.f;
  if (al->cfg.nsamples <= 0) al->cfg.nsamples = (cfg->mode == RGBID_MODE_TRACKER) ? 10000 : 19200;
  al->niters = 0;
  for (int l = cfg->finest_level; l < cfg->levels; ++l) al->niters += cfg->iterations[l] > 0 ? cfg->iterations[l] : 0;
  al->trace_stride = al->niters + 1;
  const bool tracker = (cfg->mode == RGBID_MODE_TRACKER);
  const int B = cfg->batch;

  size_t total = 0;
  for (int l = 0; l < cfg->levels; ++l) {
    LevelGeom& g = al->geom[l];
    g.rows = cfg->rows >> l; g.cols = cfg->cols >> l;
    g.pitch = align_up((size_t)g.cols * sizeof(float), 128);
    g.sstride = align_up(g.pitch * g.rows, 512);  // texture base alignment
    rgbid_error_geometry(g.rows, g.cols, al->cfg.nsamples, &g.kept_rows, &g.kept_cols, &g.sample_stride);
    int nmaps = tracker ? (al->cfg.warp_first ? (int)MAP_COUNT : (int)MAP_W_WARP) : 8;
    total += (size_t)nmaps * g.sstride * B;
  }
  size_t raw_depth = align_up((size_t)cfg->rows * cfg->cols * 2, 256), raw_rgb = align_up((size_t)cfg->rows * cfg->cols * 3, 256);
  size_t off_depth = total; total += raw_depth * B;
  size_t off_rgb = total; total += raw_rgb * B;
  al->arena_bytes = total;
  cudaError_t e = cudaMalloc(&al->d_arena, total);
  if (e != cudaSuccess) { delete al; return e == cudaErrorMemoryAllocation ? RGBID_ERR_NOMEM : RGBID_ERR_CUDA_BASE + (int)e; }
  size_t off = 0;
  for (int l = 0; l < cfg->levels; ++l) {
    const LevelGeom& g = al->geom[l];
    int nmaps = tracker ? (al->cfg.warp_first ? (int)MAP_COUNT : (int)MAP_W_WARP) : 8;
    for (int m = 0; m < nmaps; ++m) {
      al->maps[m][l] = make_img((float*)(al->d_arena + off), g.pitch, g.rows, g.cols, g.sstride);
      off += g.sstride * B;
    }
  }
  al->d_depth_raw = (uint16_t*)(al->d_arena + off_depth);
  al->d_rgb_raw = (uint8_t*)(al->d_arena + off_rgb);

  al->partial_blocks = ctx->num_sms;
  e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc(&al->d_states, sizeof(GnState) * B);
  if (e == cudaSuccess) e = cudaMalloc(&al->d_scales, sizeof(ScaleState) * B);
  if (e == cudaSuccess) e = cudaMalloc(&al->d_partials, sizeof(double) * 32 * (size_t)al->partial_blocks * B);
  if (e == cudaSuccess) e = cudaMalloc(&al->d_counters, sizeof(unsigned int) * B);
  if (e == cudaSuccess) e = cudaMalloc(&al->d_trace, sizeof(rgbid_iter_trace) * (size_t)al->trace_stride * B);
  if (e == cudaSuccess) e = cudaMalloc(&al->d_init, sizeof(double) * 12 * B);
  if (e == cudaSuccess) e = cudaMalloc(&al->d_trace_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(al->d_trace_flag, 1, sizeof(int), ctx->stream);  // traces on by default
  al->trace_enabled = 1;
  if (e == cudaSuccess) e = cudaMalloc(&al->d_active, sizeof(int) * 4 * B);
  if (e == cudaSuccess) e = cudaMallocHost(&al->h_states, sizeof(GnState) * B);
  if (e == cudaSuccess) e = cudaMallocHost(&al->h_trace, sizeof(rgbid_iter_trace) * (size_t)al->trace_stride * B);
  if (e == cudaSuccess) e = cudaMallocHost(&al->h_init, sizeof(double) * 12 * B);
  if (e == cudaSuccess) e = cudaMallocHost(&al->h_active, sizeof(int) * 4 * B);
  if (e == cudaSuccess) e = cudaMemsetAsync(al->d_counters, 0, sizeof(unsigned int) * B, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(al->d_states, 0, sizeof(GnState) * B, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(al->d_trace, 0, sizeof(rgbid_iter_trace) * (size_t)al->trace_stride * B, ctx->stream);
  // NaN-fill every map: an unset map then behaves as "all pixels invalid" instead of reading garbage
  if (e == cudaSuccess) e = cudaMemsetAsync(al->d_arena, 0xff, off_depth, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { rgbid_aligner_destroy(al); return RGBID_ERR_CUDA_BASE + (int)e; }
  // Texture objects over the current-frame pyramid (created once; the reference creates and destroys one per
  // warp call, warping_registration.cu:926-964).  RGBID_SAMPLER=soft selects the software sampler instead.
  const char* sampler = getenv("RGBID_SAMPLER");
  al->use_tex = !(sampler && sampler[0] == 's');
  if (al->use_tex) {
    const size_t ntex = (size_t)cfg->levels * 2 * B;
    al->h_tex = new (std::nothrow) cudaTextureObject_t[ntex];
    if (!al->h_tex) { rgbid_aligner_destroy(al); return RGBID_ERR_NOMEM; }
    memset(al->h_tex, 0, sizeof(cudaTextureObject_t) * ntex);
    for (int l = 0; l < cfg->levels && e == cudaSuccess; ++l)
      for (int which = 0; which < 2 && e == cudaSuccess; ++which)
        for (int b = 0; b < B && e == cudaSuccess; ++b) {
          ImgB v = al->view(which == 0 ? MAP_W_CUR : MAP_I_CUR, l, b);
          cudaResourceDesc rd;
          memset(&rd, 0, sizeof(rd));
          rd.resType = cudaResourceTypePitch2D;
          rd.res.pitch2D.devPtr = v.p; rd.res.pitch2D.pitchInBytes = v.pitch;
          rd.res.pitch2D.width = v.cols; rd.res.pitch2D.height = v.rows;
          rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
          cudaTextureDesc td;
          memset(&td, 0, sizeof(td));
          td.readMode = cudaReadModeElementType;
          // inverse depth: border addressing (0 outside the image; every user either tests in-image first, as
          // the reference does, or relies on the reference's own `res > 0` test); intensity: clamp as in the
          // reference (the bilinear footprint may touch the clamped edge)
          td.addressMode[0] = td.addressMode[1] = (which == 0) ? cudaAddressModeBorder : cudaAddressModeClamp;
          td.filterMode = which == 0 ? cudaFilterModePoint : cudaFilterModeLinear;
          td.normalizedCoords = 0;
          e = cudaCreateTextureObject(&al->h_tex[(size_t)(l * 2 + which) * B + b], &rd, &td, nullptr);
        }
    if (e == cudaSuccess) e = cudaMalloc(&al->d_tex, sizeof(cudaTextureObject_t) * ntex);
    if (e == cudaSuccess) e = cudaMemcpy(al->d_tex, al->h_tex, sizeof(cudaTextureObject_t) * ntex, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { rgbid_aligner_destroy(al); return RGBID_ERR_CUDA_BASE + (int)e; }
  }
  // second chain (see aligner_record_schedule), off by default: RGBID_CHAINS=2 enables it
  const char* chains = getenv("RGBID_CHAINS");
  if ((chains && chains[0] == '2') && B >= 2 && ctx->stream != (cudaStream_t)0) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&al->side_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&al->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&al->ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) { rgbid_aligner_destroy(al); return RGBID_ERR_CUDA_BASE + (int)e; }
  }
  const char* no_pdl = getenv("RGBID_NO_PDL");
  al->use_pdl = !(no_pdl && no_pdl[0] == '1');
  const char* no_graph = getenv("RGBID_NO_GRAPH");
  al->use_graph = !(no_graph && no_graph[0] == '1') && (ctx->stream != (cudaStream_t)0);
  al->image_filtering = RGBID_NO_FILTERS;
  *out = al;
  return RGBID_OK;
}

int rgbid_aligner_destroy(rgbid_aligner* al)
{
  if (!al) return RGBID_OK;
  cudaStreamSynchronize(al->ctx->stream);
  for (int i = 0; i < 3; ++i)
    if (al->gn_exec[i]) cudaGraphExecDestroy(al->gn_exec[i]);
  if (al->side_stream) { cudaStreamSynchronize(al->side_stream); cudaStreamDestroy(al->side_stream); }
  if (al->ev_fork) cudaEventDestroy(al->ev_fork);
  if (al->ev_join) cudaEventDestroy(al->ev_join);
  if (al->h_tex) {
    const size_t ntex = (size_t)al->cfg.levels * 2 * al->cfg.batch;
    for (size_t i = 0; i < ntex; ++i) if (al->h_tex[i]) cudaDestroyTextureObject(al->h_tex[i]);
    delete[] al->h_tex;
  }
  cudaFree(al->d_tex);
  cudaFree(al->d_arena); cudaFree(al->d_states); cudaFree(al->d_scales); cudaFree(al->d_partials);
  cudaFree(al->d_counters); cudaFree(al->d_trace); cudaFree(al->d_trace_flag); cudaFree(al->d_init); cudaFree(al->d_active);
  if (al->h_states) cudaFreeHost(al->h_states);
  if (al->h_trace) cudaFreeHost(al->h_trace);
  if (al->h_init) cudaFreeHost(al->h_init);
  if (al->h_active) cudaFreeHost(al->h_active);
  delete al;
  return RGBID_OK;
}

int rgbid_aligner_num_iterations(const rgbid_aligner* al) { return al ? al->niters : 0; }

int rgbid_aligner_iterations_done(const rgbid_aligner* al, int* out)
{
  if (!al || !out) return RGBID_ERR_ARG;
  for (int b = 0; b < al->cfg.batch; ++b)
    for (int l = 0; l < RGBID_MAX_LEVELS; ++l) out[b * RGBID_MAX_LEVELS + l] = al->h_states[b].iters_done[l];
  return RGBID_OK;
}

int rgbid_aligner_set_keyframe(rgbid_aligner* al, int index, const float* depthinv, size_t dpitch,
                               const float* intensity, size_t ipitch, int from_host)
{
  if (!al || index < 0 || index >= al->cfg.batch || !depthinv || !intensity) return RGBID_ERR_ARG;
  int rc;
  size_t wb = (size_t)al->cfg.cols * sizeof(float);
  if ((rc = copy_in(al, depthinv, dpitch, al->view(MAP_W_KF, 0, index), wb, from_host)) != RGBID_OK) return rc;
  if ((rc = copy_in(al, intensity, ipitch, al->view(MAP_I_KF, 0, index), wb, from_host)) != RGBID_OK) return rc;
  aligner_keyframe_derivatives(al, index, 1, nullptr, true);
  return check_last(al->ctx);
}

int rgbid_aligner_set_current(rgbid_aligner* al, int index, const float* depthinv, size_t dpitch,
                              const float* intensity, size_t ipitch, int from_host)
{
  if (!al || index < 0 || index >= al->cfg.batch || !depthinv || !intensity) return RGBID_ERR_ARG;
  int rc;
  size_t wb = (size_t)al->cfg.cols * sizeof(float);
  if ((rc = copy_in(al, depthinv, dpitch, al->view(MAP_W_CUR, 0, index), wb, from_host)) != RGBID_OK) return rc;
  if ((rc = copy_in(al, intensity, ipitch, al->view(MAP_I_CUR, 0, index), wb, from_host)) != RGBID_OK) return rc;
  aligner_current_pyramid(al, index, 1);
  return check_last(al->ctx);
}

int rgbid_aligner_set_current_rgbd(rgbid_aligner* al, int index, const uint16_t* depth, size_t dpitch,
                                   const uint8_t* rgb, size_t cpitch, int from_host)
{
  if (!al || index < 0 || index >= al->cfg.batch || !depth || !rgb) return RGBID_ERR_ARG;
  const int rows = al->cfg.rows, cols = al->cfg.cols;
  const uint16_t* d = depth;
  const uint8_t* c = rgb;
  size_t dp = dpitch, cp = cpitch;
  if (from_host) {
    size_t raw_depth = align_up((size_t)rows * cols * 2, 256), raw_rgb = align_up((size_t)rows * cols * 3, 256);
    uint16_t* dd = (uint16_t*)((char*)al->d_depth_raw + raw_depth * index);
    uint8_t* dc = al->d_rgb_raw + raw_rgb * index;
    RGBID_CUDA_TRY(cudaMemcpy2DAsync(dd, (size_t)cols * 2, depth, dpitch, (size_t)cols * 2, rows, cudaMemcpyHostToDevice, al->ctx->stream));
    RGBID_CUDA_TRY(cudaMemcpy2DAsync(dc, (size_t)cols * 3, rgb, cpitch, (size_t)cols * 3, rows, cudaMemcpyHostToDevice, al->ctx->stream));
    d = dd; c = dc; dp = (size_t)cols * 2; cp = (size_t)cols * 3;
  }
  launch_ingest(al->ctx->L(), d, dp, 0, c, cp, 0, al->view(MAP_W_CUR, 0, index), al->view(MAP_I_CUR, 0, index), 1,
                al->cfg.factor_depth);
  aligner_current_pyramid(al, index, 1);
  return check_last(al->ctx);
}

int rgbid_aligner_current_to_keyframe(rgbid_aligner* al, int index)
{
  if (!al || index < 0 || index >= al->cfg.batch) return RGBID_ERR_ARG;
  aligner_copy_current_to_keyframe(al, index, 1, nullptr);
  aligner_keyframe_derivatives(al, index, 1, nullptr, false);
  return check_last(al->ctx);
}

int rgbid_aligner_enqueue(rgbid_aligner* al, const double* R_init, const double* t_init)
{
  if (!al || !R_init || !t_init) return RGBID_ERR_ARG;
  const int B = al->cfg.batch;
  cudaStream_t s = al->ctx->stream;
  memcpy(al->h_init, R_init, sizeof(double) * 9 * B);
  memcpy(al->h_init + 9 * B, t_init, sizeof(double) * 3 * B);
  RGBID_CUDA_TRY(cudaMemcpyAsync(al->d_init, al->h_init, sizeof(double) * 12 * B, cudaMemcpyHostToDevice, s));
  return aligner_enqueue_device_init(al);
}

int rgbid_aligner_set_trace(rgbid_aligner* al, int enable)
{
  if (!al) return RGBID_ERR_ARG;
  enable = enable ? 1 : 0;
  if (enable == al->trace_enabled) return RGBID_OK;
  // read by gn_init_kernel at the start of every run (the recorded schedule itself does not change)
  RGBID_CUDA_TRY(cudaMemsetAsync(al->d_trace_flag, enable, sizeof(int), al->ctx->stream));
  al->trace_enabled = enable;
  return RGBID_OK;
}

int rgbid_aligner_fetch(rgbid_aligner* al, double* R_out, double* t_out, double* cov_out, int* status_out,
                        rgbid_iter_trace* trace_out)
{
  if (!al) return RGBID_ERR_ARG;
  const int B = al->cfg.batch;
  cudaStream_t s = al->ctx->stream;
  RGBID_CUDA_TRY(cudaMemcpyAsync(al->h_states, al->d_states, sizeof(GnState) * B, cudaMemcpyDeviceToHost, s));
  if (trace_out)
    RGBID_CUDA_TRY(cudaMemcpyAsync(al->h_trace, al->d_trace, sizeof(rgbid_iter_trace) * (size_t)al->trace_stride * B,
                                   cudaMemcpyDeviceToHost, s));
  RGBID_CUDA_TRY(cudaStreamSynchronize(s));
  int rc = check_last(al->ctx);
  if (rc != RGBID_OK) return rc;
  for (int b = 0; b < B; ++b) {
    const GnState& st = al->h_states[b];
    if (R_out) memcpy(R_out + 9 * b, st.R, sizeof(double) * 9);
    if (t_out) memcpy(t_out + 3 * b, st.t, sizeof(double) * 3);
    if (cov_out) memcpy(cov_out + 36 * b, st.cov, sizeof(double) * 36);
    if (status_out) status_out[b] = st.status;
  }
  if (trace_out) memcpy(trace_out, al->h_trace, sizeof(rgbid_iter_trace) * (size_t)al->trace_stride * B);
  return RGBID_OK;
}

int rgbid_aligner_run(rgbid_aligner* al, double* R_inout, double* t_inout, double* cov_out, int* status_out,
                      rgbid_iter_trace* trace_out)
{
  if (!al) return RGBID_ERR_ARG;
  int rc = rgbid_aligner_set_trace(al, trace_out != nullptr);  // this call knows whether the trace is wanted
  if (rc != RGBID_OK) return rc;
  rc = rgbid_aligner_enqueue(al, R_inout, t_inout);
  if (rc != RGBID_OK) return rc;
  return rgbid_aligner_fetch(al, R_inout, t_inout, cov_out, status_out, trace_out);
}

int rgbid_aligner_frame_stats(rgbid_aligner* al, float* stats_out)
{
  if (!al || !stats_out) return RGBID_ERR_ARG;
  for (int b = 0; b < al->cfg.batch; ++b) {
    stats_out[3 * b + 0] = al->h_states[b].chi_square;
    stats_out[3 * b + 1] = al->h_states[b].chi_test;
    stats_out[3 * b + 2] = al->h_states[b].ndof;
  }
  return RGBID_OK;
}

size_t rgbid_aligner_state_bytes(void) { return sizeof(GnState); }

int rgbid_aligner_export_systems(rgbid_aligner* al, double* d_out)
{
  if (!al || !d_out) return RGBID_ERR_ARG;
  launch_export_systems(al->ctx->L(), al->d_states, d_out, al->cfg.batch);
  return check_last(al->ctx);
}

static int time_kernel(rgbid_aligner* al, int level, int reps, float* ms_per_launch, bool scale);

int rgbid_aligner_time_build(rgbid_aligner* al, int level, int reps, float* ms_per_launch)
{
  return time_kernel(al, level, reps, ms_per_launch, false);
}

int rgbid_aligner_time_scale(rgbid_aligner* al, int level, int reps, float* ms_per_launch)
{
  return time_kernel(al, level, reps, ms_per_launch, true);
}

static int time_kernel(rgbid_aligner* al, int level, int reps, float* ms_per_launch, bool scale)
{
  if (!al || level < 0 || level >= al->cfg.levels || reps < 1 || !ms_per_launch) return RGBID_ERR_ARG;
  cudaStream_t s = al->ctx->stream;
  cudaEvent_t e0, e1;
  RGBID_CUDA_TRY(cudaEventCreate(&e0));
  RGBID_CUDA_TRY(cudaEventCreate(&e1));
  GnParams P = base_params(al, level);
  // the shipped launch, tail included (final sum, 6x6 solve, pose update, projection refresh), repeatable: see dry_tail
  P.iter_index = -1; P.update_pose = 1; P.dry_tail = 1; P.compute_cov = 0; P.next_level = level;
  const bool tracker = (al->cfg.mode == RGBID_MODE_TRACKER);
  P.use_scale = (tracker && al->cfg.sigma_estimator != RGBID_SIGMA_PDF) ? 0 : 1;
  GnLevelMaps M = level_maps(al, level, false);
  LaunchCtx L = al->ctx->L();
  L.pdl = al->use_pdl;
  auto go = [&]() {
    if (scale) launch_gn_scale(L, M, P, al->d_states, al->d_scales);
    else launch_gn_build(L, M, P, al->d_states, al->d_scales, al->d_partials, 32, al->d_counters, nullptr);
  };
  go();  // warm
  RGBID_CUDA_TRY(cudaEventRecord(e0, s));
  for (int r = 0; r < reps; ++r) go();
  RGBID_CUDA_TRY(cudaEventRecord(e1, s));
  RGBID_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  RGBID_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_launch = ms / reps;
  return check_last(al->ctx);
}

int rgbid_aligner_map(rgbid_aligner* al, int which, int level, int index, float** ptr, size_t* pitch)
{
  if (!al || which < 0 || which >= MAP_COUNT || level < 0 || level >= al->cfg.levels || index < 0 ||
      index >= al->cfg.batch || !ptr || !pitch)
    return RGBID_ERR_ARG;
  if (al->maps[which][level].p == nullptr) return RGBID_ERR_STATE;
  ImgB v = al->view(which, level, index);
  *ptr = v.p; *pitch = v.pitch;
  return RGBID_OK;
}

}  // extern "C"

namespace rgbid {

// Launch (or replay) the schedule; d_init must already hold the initial guesses.
bool aligner_can_split(const rgbid_aligner* al)
{
  return al->cfg.mode == RGBID_MODE_TRACKER && al->side_stream == nullptr;
}

// part: the whole schedule, or -- tracker mode -- the iterations and the covariance pass as two graphs, so that the caller
// can read the pose back and go on with its host work while the covariance pass runs (tracker.cu)
int aligner_enqueue_part(rgbid_aligner* al, int part)
{
  cudaStream_t s = al->ctx->stream;
  if (part != ALIGN_PART_ALL && !aligner_can_split(al)) return RGBID_ERR_STATE;
  if (!al->use_graph) {
    aligner_record_schedule(al, part);
    return check_last(al->ctx);
  }
  if (!al->gn_exec[part]) {
    long long before = al->ctx->launches;
    cudaGraph_t graph = nullptr;
    RGBID_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    aligner_record_schedule(al, part);
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (e != cudaSuccess) return RGBID_ERR_CUDA_BASE + (int)e;
    al->gn_graph_launches[part] = al->ctx->launches - before;
    al->ctx->launches = before;
    e = cudaGraphInstantiate(&al->gn_exec[part], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { al->gn_exec[part] = nullptr; return RGBID_ERR_CUDA_BASE + (int)e; }
  }
  RGBID_CUDA_TRY(cudaGraphLaunch(al->gn_exec[part], s));
  al->ctx->launches += al->gn_graph_launches[part];
  return RGBID_OK;
}

int aligner_enqueue_device_init(rgbid_aligner* al) { return aligner_enqueue_part(al, ALIGN_PART_ALL); }

}  // namespace rgbid
