// tracker.cu -- per-frame state machine of VisodoTracker::trackNewFrame (src/visodo.cpp:1967-2247) for
// `batch` independent RGB-D streams advancing in lock step.
//
// Per frame and per stream the reference performs ~700 kernel launches / stream synchronisations; here one
// frame of ALL streams is: 1 ingest launch + (levels-1) pyramid launches, one graph replay for the whole
// Gauss-Newton schedule + covariance pass, 4 covisibility launches, and 4-5 launches of fusion / map
// maintenance (keyframe refresh launches are predicated per stream on the device).  The host synchronises
// twice per frame (pose read-back, covisibility read-back) to take the keyframe decisions in double
// precision exactly like the reference's host code.
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>
#include <vector>

#include "aligner.hpp"

using namespace rgbid;

namespace {

struct StreamState {
  double R_odoKF[9], t_odoKF[3];  // last_odoKF_global_{rotation,translation}_
  double R_est[9], t_est[3];      // last_estimated_{rotation,translation}_
  double dR[9], dt[3], dcov[36];  // delta_{rotation,translation,covariance}_ : _{odoKF}T^{cur}
  double R_intKF[9], t_intKF[3];  // last_integrKF_global_*
  double vel[3], omega[3];        // constant-velocity model state
  bool lost;
  int global_time, odoKF_count, integrKF_count;
  // odometry-keyframe -> integration-keyframe chain (delta_*_odo2integr_{last,next}_, src/visodo.cpp:1553-1566, 1592-1670)
  double o2i_next_R[9], o2i_next_t[3], o2i_next_cov[36];
  double o2i_last_R[9], o2i_last_t[3], o2i_last_cov[36];
  int last_integrKF_index;
};

void set_identity(double* R, double* t)
{
  for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  t[0] = t[1] = t[2] = 0.0;
}

// ---- 6x6 covariance propagation of the constraint chain (host, double) ---------------------------------------------
// out += J C J^T
void add_JCJt(const double* J, const double* C, double* out)
{
  double T[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double v = 0.0;
      for (int k = 0; k < 6; ++k) v += J[6 * i + k] * C[6 * k + j];
      T[6 * i + j] = v;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double v = 0.0;
      for (int k = 0; k < 6; ++k) v += T[6 * i + k] * J[6 * j + k];
      out[6 * i + j] += v;
    }
}

void set_block3(double* J, int r0, int c0, const double* M, double sign = 1.0)
{
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[6 * (r0 + i) + c0 + j] = sign * M[3 * i + j];
}

// T_new = T_old * T_upd on the odo->integration chain: the part resetOdometryKeyframe (src/visodo.cpp:1553-1566) and
// resetIntegrationKeyframe (:1592-1605) have in common
void chain_compose(StreamState& S)
{
  double J[36] = {0}, tnew[3], K[9], Rn[9];
  set_block3(J, 0, 0, S.o2i_next_R);
  set_block3(J, 3, 3, S.o2i_next_R);
  mat3_vec(S.o2i_next_R, S.dt, tnew);
  skew3(tnew, K);
  set_block3(J, 0, 3, K);
  add_JCJt(J, S.dcov, S.o2i_next_cov);
  for (int k = 0; k < 3; ++k) S.o2i_next_t[k] = tnew[k] + S.o2i_next_t[k];
  mat3_mul(S.o2i_next_R, S.dR, Rn);
  memcpy(S.o2i_next_R, Rn, sizeof(Rn));
}

// Relative constraint between two poses expressed in the same frame, with first-order covariance
// (src/visodo.cpp:2126-2146 for SEQ_ODO, :1610-1629 for SEQ_KF):
//   R = R_last^T R_new, t = R_last^T (t_new - t_last),
//   cov = dLast cov_last dLast^T + dNew cov_new dNew^T
void relative_constraint(const double* R_last, const double* t_last, const double* cov_last, const double* R_new,
                         const double* t_new, const double* cov_new, double* R, double* t, double* cov)
{
  double Rt[9], d[3], K[9], KR[9], Jn[36] = {0}, Jl[36] = {0};
  mat3_transpose(R_last, Rt);
  mat3_mul(Rt, R_new, R);
  for (int k = 0; k < 3; ++k) d[k] = t_new[k] - t_last[k];
  mat3_vec(Rt, d, t);
  set_block3(Jn, 0, 0, Rt); set_block3(Jn, 3, 3, Rt);
  set_block3(Jl, 0, 0, Rt, -1.0); set_block3(Jl, 3, 3, Rt, -1.0);
  skew3(t, K);
  mat3_mul(K, Rt, KR);
  set_block3(Jl, 0, 3, KR);
  for (int i = 0; i < 36; ++i) cov[i] = 0.0;
  add_JCJt(Jl, cov_last, cov);
  add_JCJt(Jn, cov_new, cov);
}

}  // namespace

struct rgbid_tracker {
  rgbid_ctx* ctx;
  rgbid_tracker_config cfg;
  rgbid_aligner* al;
  std::vector<StreamState> st;
  int frame_index;
  // integration keyframe (level 0), batched
  char* d_arena;
  ImgB intW, intWraw, intWeight, wstate, vmap, nmap, intGx, intGy;
  uint8_t* d_mask; size_t mask_pitch, mask_sstride;
  uint8_t* d_colors; size_t colors_sstride;
  Proj* d_proj; Proj* h_proj;              // [4][batch]: odo cur->KF, odo KF->cur, integr cur->KF, integr KF->cur
  unsigned int* d_counts; unsigned int* h_counts;  // [batch][8]
  int* d_flags; int* h_flags;              // [4][batch]: new odo KF, new integration KF, fuse, overlap mask refresh
  // custom calibration (rgbid_tracker_set_custom_calibration): parameters, projective matrices, one stream of scratch
  bool custom_on;
  rgbid_custom_calibration custom;
  float dRc_proj[9], t_dc_proj[3], cRd_proj[9];
  char* d_custom; ImgB cW, cI, cPre; int* d_canvas; size_t canvas_pitch;
  // keyframe hand-off (rgbid_tracker_set_keyframe_sink): pinned host staging for one outgoing keyframe
  rgbid_keyframe_sink sink; void* sink_user;
  char* h_handoff; size_t handoff_bytes;
  // rgbid_tracker_prefetch: second raw-frame staging buffer, filled on a copy stream while the previous frame is tracked
  // (two buffers used alternately: frame k may still be read by its ingest kernel when frame k + 1 starts to arrive)
  cudaStream_t copy_stream; cudaEvent_t ev_copy[2];
  char* d_prefetch[2];                     // [batch] depth, then [batch] rgb (same layout as the aligner's raw staging)
  const void* pf_depth[2]; const void* pf_rgb[2]; bool pf_valid[2]; int pf_next;
  long long pf_issued_at[2], track_calls;  // a prefetched frame is good for the current or the next track call only
  cudaEvent_t ev_track_done; bool ev_track_done_valid;  // end of the device work queued by the last track call
  // split read-back: the pose is read when the iterations are done, the covariance / chi^2 (second copy of the solver
  // state) with the covisibility counts, so that the host's pose bookkeeping runs under the covariance pass
  GnState* h_states_cov; cudaEvent_t ev_pose, ev_iter_done; cudaStream_t rb_stream; bool split_cov;
};

namespace {

void save_integration_keyframes(rgbid_tracker* t, const int* active)
{
  // saveCurrentImagesAsIntegrationKeyframes, src/visodo.cpp:880-893
  rgbid_aligner* al = t->al;
  LaunchCtx L = t->ctx->L();
  const rgbid_align_config& c = al->cfg;
  const int B = c.batch;
  launch_copy2(L, al->maps[MAP_W_CUR][0], t->intW, al->maps[MAP_W_CUR][0], t->intWraw, B, active);
  launch_fill(L, t->intWeight, 1.f, B, active);  // initialiseWeightKernel: 1 everywhere (misc.cu:272-287)
}

void refresh_integration_maps(rgbid_tracker* t)
{
  // createVMap + computeGradientDepth + createNMapGradients (src/visodo.cpp:889-892, 1753-1758)
  rgbid_aligner* al = t->al;
  LaunchCtx L = t->ctx->L();
  const rgbid_align_config& c = al->cfg;
  ImgB none = make_img(nullptr, 0, 0, 0);
  if (launch_keyframe_maps(L, t->intW, t->intGx, t->intGy, t->vmap, t->nmap, c.fx, c.fy, c.cx, c.cy, c.batch)) return;
  launch_vmap(L, t->intW, t->vmap, c.fx, c.fy, c.cx, c.cy, c.batch);
  launch_gradient2(L, t->intW, t->intGx, t->intGy, none, none, none, c.batch);
  launch_nmap_gradients(L, t->intW, t->intGx, t->intGy, t->nmap, c.fx, c.fy, c.cx, c.cy, c.batch);
}

void to_proj(const double* R, const double* tt, const rgbid_align_config& c, Proj* P)
{
  projective_pose(R, tt, c.fx, c.fy, c.cx, c.cy, P->r, P->t);
}

void to_proj_inverse(const double* R, const double* tt, const rgbid_align_config& c, Proj* P)
{
  projective_inverse_pose(R, tt, c.fx, c.fy, c.cx, c.cy, P->r, P->t);
}

// Upload of the next frame WITHOUT the copy engine: a few CTAs read the pinned host buffer directly (unified
// addressing) and write the device staging buffer.  The host->device copy engine is shared with the tracker's own
// small control copies (initial guesses, transforms, keyframe flags), which sit on the critical path of the frame
// being tracked; a 49 MB cudaMemcpyAsync queued ahead of them stalls that frame for the whole transfer (measured:
// 7.2 ms per step instead of 3.9), and splitting it into per-image copies only restores the serial time.
__global__ void __launch_bounds__(256) prefetch_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread: PCIe latency is microseconds
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = src[i];
}

// The tracker's small per-frame control uploads (initial guesses, transforms, keyframe flags: a few KB from pinned host
// memory) go through this one-CTA zero-copy kernel instead of cudaMemcpyAsync, so they never queue behind the bulk
// upload of the next frame on the host->device copy engine.
__global__ void __launch_bounds__(256) control_upload_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n32)
{
  for (int i = threadIdx.x; i < n32; i += blockDim.x) dst[i] = src[i];
}

void upload_control(const LaunchCtx& L, void* dst, const void* src_pinned, size_t bytes)
{
  control_upload_kernel<<<1, 256, 0, L.stream>>>((uint32_t*)dst, (const uint32_t*)src_pinned, (int)(bytes / 4));
  ++*L.launches;
}

bool host_pointer_is_device_readable(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer != nullptr;
}

}  // namespace

extern "C" {

int rgbid_tracker_create(rgbid_ctx* ctx, const rgbid_tracker_config* cfg, rgbid_tracker** out)
{
  if (!ctx || !cfg || !out) return RGBID_ERR_ARG;
  *out = nullptr;
  rgbid_tracker* t = new (std::nothrow) rgbid_tracker();
  if (!t) return RGBID_ERR_NOMEM;
  t->ctx = ctx;
  t->cfg = *cfg;
  t->cfg.align.mode = RGBID_MODE_TRACKER;
  if (t->cfg.delta_t <= 0.f) t->cfg.delta_t = 0.03333f;
  if (t->cfg.visratio_odo <= 0.f) t->cfg.visratio_odo = 0.9f;
  if (t->cfg.visratio_integr <= 0.f) t->cfg.visratio_integr = 0.7f;
  if (t->cfg.max_odo_kf_count <= 0) t->cfg.max_odo_kf_count = 9999999;
  if (t->cfg.max_integr_kf_count <= 0) t->cfg.max_integr_kf_count = 9999999;
  t->al = nullptr; t->d_arena = nullptr; t->d_proj = nullptr; t->h_proj = nullptr; t->d_counts = nullptr;
  t->h_counts = nullptr; t->d_flags = nullptr; t->h_flags = nullptr;
  t->custom_on = false; t->d_custom = nullptr; t->d_canvas = nullptr;
  t->sink = nullptr; t->sink_user = nullptr; t->h_handoff = nullptr; t->handoff_bytes = 0;
  t->ev_track_done = nullptr; t->ev_track_done_valid = false;
  t->h_states_cov = nullptr; t->ev_pose = nullptr; t->ev_iter_done = nullptr; t->rb_stream = nullptr; t->split_cov = false;
  t->copy_stream = nullptr; t->pf_next = 0; t->track_calls = 0; t->pf_issued_at[0] = t->pf_issued_at[1] = 0;
  for (int i = 0; i < 2; ++i) { t->ev_copy[i] = nullptr; t->d_prefetch[i] = nullptr; t->pf_depth[i] = t->pf_rgb[i] = nullptr; t->pf_valid[i] = false; }
  int rc = rgbid_aligner_create(ctx, &t->cfg.align, &t->al);
  if (rc == RGBID_OK) rc = rgbid_aligner_set_trace(t->al, 0);  // the tracker never reads the per-iteration trace
  if (rc != RGBID_OK) { delete t; return rc; }
  t->al->image_filtering = t->cfg.image_filtering;
  const rgbid_align_config& c = t->al->cfg;
  const int B = c.batch, rows = c.rows, cols = c.cols;
  const LevelGeom& g = t->al->geom[0];
  size_t map1 = g.sstride, map3 = align_up(g.pitch * rows * 3, 256);
  t->mask_pitch = align_up((size_t)cols, 128);
  t->mask_sstride = align_up(t->mask_pitch * rows, 256);
  t->colors_sstride = align_up((size_t)rows * cols * 3, 256);
  size_t total = (6 * map1 + 2 * map3 + t->mask_sstride + t->colors_sstride) * B;
  cudaError_t e = cudaMalloc(&t->d_arena, total);
  if (e != cudaSuccess) { rgbid_tracker_destroy(t); return e == cudaErrorMemoryAllocation ? RGBID_ERR_NOMEM : RGBID_ERR_CUDA_BASE + (int)e; }
  size_t off = 0;
  auto carve1 = [&](ImgB& m) { m = make_img((float*)(t->d_arena + off), g.pitch, rows, cols, map1); off += map1 * B; };
  auto carve3 = [&](ImgB& m) { m = make_img((float*)(t->d_arena + off), g.pitch, 3 * rows, cols, map3); off += map3 * B; };
  carve1(t->intW); carve1(t->intWraw); carve1(t->intWeight); carve1(t->wstate); carve1(t->intGx); carve1(t->intGy);
  carve3(t->vmap); carve3(t->nmap);
  t->d_mask = (uint8_t*)(t->d_arena + off); off += t->mask_sstride * B;
  t->d_colors = (uint8_t*)(t->d_arena + off); off += t->colors_sstride * B;
  if (e == cudaSuccess) e = cudaMalloc(&t->d_proj, sizeof(Proj) * 4 * B);
  if (e == cudaSuccess) e = cudaMallocHost(&t->h_proj, sizeof(Proj) * 4 * B);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_counts, sizeof(unsigned int) * 8 * B);
  if (e == cudaSuccess) e = cudaMallocHost(&t->h_counts, sizeof(unsigned int) * 8 * B);
  if (e == cudaSuccess) e = cudaMallocHost(&t->h_states_cov, sizeof(GnState) * B);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_pose, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_iter_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->rb_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_flags, sizeof(int) * 4 * B);
  if (e == cudaSuccess) e = cudaMallocHost(&t->h_flags, sizeof(int) * 4 * B);
  if (e != cudaSuccess) { rgbid_tracker_destroy(t); return RGBID_ERR_CUDA_BASE + (int)e; }
  t->st.resize(B);
  rc = rgbid_tracker_reset(t);
  if (rc != RGBID_OK) { rgbid_tracker_destroy(t); return rc; }
  *out = t;
  return RGBID_OK;
}

int rgbid_tracker_destroy(rgbid_tracker* t)
{
  if (!t) return RGBID_OK;
  cudaStreamSynchronize(t->ctx->stream);
  if (t->al) rgbid_aligner_destroy(t->al);
  if (t->copy_stream) { cudaStreamSynchronize(t->copy_stream); cudaStreamDestroy(t->copy_stream); }
  for (int i = 0; i < 2; ++i) { if (t->ev_copy[i]) cudaEventDestroy(t->ev_copy[i]); cudaFree(t->d_prefetch[i]); }
  if (t->ev_track_done_valid) cudaEventDestroy(t->ev_track_done);
  cudaFree(t->d_arena); cudaFree(t->d_proj); cudaFree(t->d_counts); cudaFree(t->d_flags);
  cudaFree(t->d_custom);
  if (t->h_handoff) cudaFreeHost(t->h_handoff);
  if (t->h_proj) cudaFreeHost(t->h_proj);
  if (t->h_counts) cudaFreeHost(t->h_counts);
  if (t->h_states_cov) cudaFreeHost(t->h_states_cov);
  if (t->ev_pose) cudaEventDestroy(t->ev_pose);
  if (t->ev_iter_done) cudaEventDestroy(t->ev_iter_done);
  if (t->rb_stream) cudaStreamDestroy(t->rb_stream);
  if (t->h_flags) cudaFreeHost(t->h_flags);
  delete t;
  return RGBID_OK;
}

int rgbid_tracker_reset(rgbid_tracker* t)
{
  if (!t) return RGBID_ERR_ARG;
  t->frame_index = 0;
  for (auto& s : t->st) {
    memset(&s, 0, sizeof(s));
    set_identity(s.R_odoKF, s.t_odoKF); set_identity(s.R_est, s.t_est); set_identity(s.dR, s.dt);
    set_identity(s.R_intKF, s.t_intKF);
    set_identity(s.o2i_next_R, s.o2i_next_t); set_identity(s.o2i_last_R, s.o2i_last_t);
    s.last_integrKF_index = 0;
    s.lost = false; s.global_time = 0;
  }
  // a frame prefetched before the reset must not be matched (by pointer) by the first track call after it
  if (t->copy_stream) RGBID_CUDA_TRY(cudaStreamSynchronize(t->copy_stream));
  for (int i = 0; i < 2; ++i) { t->pf_valid[i] = false; t->pf_depth[i] = nullptr; t->pf_rgb[i] = nullptr; t->pf_issued_at[i] = 0; }
  t->track_calls = 0;
  const rgbid_align_config& c = t->al->cfg;
  cudaStream_t s = t->ctx->stream;
  size_t fl = (size_t)(t->d_mask - (uint8_t*)t->d_arena);
  RGBID_CUDA_TRY(cudaMemsetAsync(t->d_arena, 0xff, fl, s));  // NaN-fill float maps
  // warped_weight_curr_ is never cleared by the reference (fresh cudaMalloc memory): start from zero weights
  RGBID_CUDA_TRY(cudaMemsetAsync(t->wstate.p, 0, t->wstate.sstride * c.batch, s));
  RGBID_CUDA_TRY(cudaMemsetAsync(t->d_mask, 0, (t->mask_sstride + t->colors_sstride) * c.batch, s));
  return RGBID_OK;
}

rgbid_aligner* rgbid_tracker_aligner(rgbid_tracker* t) { return t ? t->al : nullptr; }

int rgbid_tracker_keyframe_map(rgbid_tracker* t, int which, int index, float** ptr, size_t* pitch)
{
  if (!t || !ptr || !pitch || index < 0 || index >= t->al->cfg.batch) return RGBID_ERR_ARG;
  const ImgB* m = nullptr;
  switch (which) {
    case 0: m = &t->intW; break;
    case 1: m = &t->intWeight; break;
    case 2: m = &t->intWraw; break;
    case 3: m = &t->vmap; break;
    case 4: m = &t->nmap; break;
    case 5: m = &t->intGx; break;
    case 6: m = &t->intGy; break;
    default: return RGBID_ERR_ARG;
  }
  *ptr = (float*)((char*)m->p + (size_t)index * m->sstride);
  *pitch = m->pitch;
  return RGBID_OK;
}

int rgbid_tracker_overlap_mask(rgbid_tracker* t, int index, uint8_t** ptr, size_t* pitch)
{
  if (!t || !ptr || !pitch || index < 0 || index >= t->al->cfg.batch) return RGBID_ERR_ARG;
  *ptr = t->d_mask + (size_t)index * t->mask_sstride;
  *pitch = t->mask_pitch;
  return RGBID_OK;
}

static int track_core(rgbid_tracker* t, const uint16_t* depth, const uint8_t* rgb, int from_host, size_t in_dpitch,
                      size_t in_dstride, size_t in_cpitch, size_t in_cstride, rgbid_frame_result* results);

int rgbid_tracker_set_custom_calibration(rgbid_tracker* t, const rgbid_custom_calibration* cal)
{
  if (!t) return RGBID_ERR_ARG;
  if (!cal) { t->custom_on = false; return RGBID_OK; }
  const rgbid_align_config& c = t->al->cfg;
  if (!t->d_custom) {
    const LevelGeom& g = t->al->geom[0];
    t->canvas_pitch = align_up((size_t)3 * c.cols * sizeof(int), 128);
    const size_t maps = 3 * g.sstride, canvas = t->canvas_pitch * 3 * c.rows;
    RGBID_CUDA_TRY(cudaMalloc(&t->d_custom, maps + canvas));
    t->cW = make_img((float*)t->d_custom, g.pitch, c.rows, c.cols, 0);
    t->cI = make_img((float*)(t->d_custom + g.sstride), g.pitch, c.rows, c.cols, 0);
    t->cPre = make_img((float*)(t->d_custom + 2 * g.sstride), g.pitch, c.rows, c.cols, 0);
    t->d_canvas = (int*)(t->d_custom + maps);
  }
  t->custom = *cal;
  // dRc_proj = Kd dRc Kc^-1, t_dc_proj = Kd t_dc, cRd_proj = dRc_proj^-1 in float (src/visodo.cpp:792-801)
  const rgbid_intr& kc = cal->rgb;
  const rgbid_intr& kd = cal->depth;
  const float Kd[9] = {kd.fx, 0.f, kd.cx, 0.f, kd.fy, kd.cy, 0.f, 0.f, 1.f};
  const float Kci[9] = {1.f / kc.fx, 0.f, -kc.cx / kc.fx, 0.f, 1.f / kc.fy, -kc.cy / kc.fy, 0.f, 0.f, 1.f};
  float T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[3 * i + j] = Kd[3 * i] * cal->dRc[j] + Kd[3 * i + 1] * cal->dRc[3 + j] + Kd[3 * i + 2] * cal->dRc[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t->dRc_proj[3 * i + j] = T[3 * i] * Kci[j] + T[3 * i + 1] * Kci[3 + j] + T[3 * i + 2] * Kci[6 + j];
  for (int i = 0; i < 3; ++i) t->t_dc_proj[i] = Kd[3 * i] * cal->t_dc[0] + Kd[3 * i + 1] * cal->t_dc[1] + Kd[3 * i + 2] * cal->t_dc[2];
  const float* M = t->dRc_proj;
  const float c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  const float id = 1.f / (M[0] * c00 + M[1] * c01 + M[2] * c02);
  float* I = t->cRd_proj;
  I[0] = c00 * id; I[1] = (M[2] * M[7] - M[1] * M[8]) * id; I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  I[3] = c01 * id; I[4] = (M[0] * M[8] - M[2] * M[6]) * id; I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  I[6] = c02 * id; I[7] = (M[1] * M[6] - M[0] * M[7]) * id; I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  t->custom_on = true;
  return RGBID_OK;
}

int rgbid_tracker_set_keyframe_sink(rgbid_tracker* t, rgbid_keyframe_sink cb, void* user)
{
  if (!t) return RGBID_ERR_ARG;
  t->sink = cb; t->sink_user = user;
  return RGBID_OK;
}

// Downloads the outgoing integration keyframe of stream b (before it is overwritten) and calls the sink.
static int hand_off_keyframe(rgbid_tracker* t, int b, const double* rel_R, const double* rel_t, const double* rel_cov,
                             int frame_index)
{
  const rgbid_align_config& c = t->al->cfg;
  const StreamState& S = t->st[b];
  cudaStream_t s = t->ctx->stream;
  const size_t rows = c.rows, cols = c.cols;
  const size_t mask_b = cols * rows, col_b = cols * rows * 3, map_b = cols * rows * sizeof(float);
  const size_t off_col = align_up(mask_b, 256), off_w = off_col + align_up(col_b, 256), off_n = off_w + align_up(map_b, 256);
  const size_t need = off_n + 3 * map_b;
  if (t->handoff_bytes < need) {
    if (t->h_handoff) cudaFreeHost(t->h_handoff);
    t->h_handoff = nullptr; t->handoff_bytes = 0;
    RGBID_CUDA_TRY(cudaMallocHost(&t->h_handoff, need));
    t->handoff_bytes = need;
  }
  char* h = t->h_handoff;
  RGBID_CUDA_TRY(cudaMemcpy2DAsync(h, cols, t->d_mask + t->mask_sstride * b, t->mask_pitch, cols, rows, cudaMemcpyDeviceToHost, s));
  RGBID_CUDA_TRY(cudaMemcpyAsync(h + off_col, t->d_colors + t->colors_sstride * b, col_b, cudaMemcpyDeviceToHost, s));
  RGBID_CUDA_TRY(cudaMemcpy2DAsync(h + off_w, cols * sizeof(float), (char*)t->intW.p + t->intW.sstride * b, t->intW.pitch,
                                   cols * sizeof(float), rows, cudaMemcpyDeviceToHost, s));
  RGBID_CUDA_TRY(cudaMemcpy2DAsync(h + off_n, cols * sizeof(float), (char*)t->nmap.p + t->nmap.sstride * b, t->nmap.pitch,
                                   cols * sizeof(float), 3 * rows, cudaMemcpyDeviceToHost, s));
  RGBID_CUDA_TRY(cudaStreamSynchronize(s));
  rgbid_keyframe_handoff k;
  memset(&k, 0, sizeof(k));
  k.stream = b; k.kf_index = S.last_integrKF_index; k.frame_index = frame_index;
  k.rows = c.rows; k.cols = c.cols; k.fx = c.fx; k.fy = c.fy; k.cx = c.cx; k.cy = c.cy;
  memcpy(k.R, S.R_intKF, sizeof(k.R)); memcpy(k.t, S.t_intKF, sizeof(k.t));
  memcpy(k.rel_R, rel_R, sizeof(k.rel_R)); memcpy(k.rel_t, rel_t, sizeof(k.rel_t)); memcpy(k.rel_cov, rel_cov, sizeof(k.rel_cov));
  k.overlap_mask = (const uint8_t*)h; k.overlap_mask_pitch = cols;
  k.colors = (const uint8_t*)(h + off_col);
  k.depthinv = (const float*)(h + off_w); k.depthinv_pitch = cols * sizeof(float);
  k.normals = (const float*)(h + off_n); k.normals_pitch = cols * sizeof(float);
  t->sink(t->sink_user, &k);
  return RGBID_OK;
}

int rgbid_tracker_prefetch(rgbid_tracker* t, const uint16_t* depth, const uint8_t* rgb)
{
  if (!t || !depth || !rgb) return RGBID_ERR_ARG;
  const rgbid_align_config& c = t->al->cfg;
  const size_t dsz = (size_t)c.rows * c.cols * 2, csz = (size_t)c.rows * c.cols * 3;
  const size_t raw_depth = align_up(dsz, 256), raw_rgb = align_up(csz, 256);
  const int B = c.batch;
  if (!t->copy_stream) {
    RGBID_CUDA_TRY(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      RGBID_CUDA_TRY(cudaEventCreateWithFlags(&t->ev_copy[i], cudaEventDisableTiming));
      RGBID_CUDA_TRY(cudaMalloc(&t->d_prefetch[i], (raw_depth + raw_rgb) * B));
    }
  }
  // This staging buffer was last read by the frame before the current one: by its ingest kernel and, if that frame
  // became an integration keyframe, by the colour copy queued at the very end of track_core -- after the last stream
  // synchronisation of that call.  The upload therefore waits for the event track_core records when it returns.
  if (t->ev_track_done_valid) RGBID_CUDA_TRY(cudaStreamWaitEvent(t->copy_stream, t->ev_track_done, 0));
  const int slot = t->pf_next;
  t->pf_next ^= 1;
  char* dd = t->d_prefetch[slot];
  char* dc = t->d_prefetch[slot] + raw_depth * B;
  // Bulk upload on the copy engine: it overlaps the tracking of the current frame without touching an SM (the system
  // kernel is sized to fill every register file, so any co-resident copy kernel CTA pushes one of its CTAs into a second
  // wave; measured with a zero-copy upload kernel: +0.5 .. 0.8 ms per step).  The tracker's own control uploads do not
  // use the copy engine (upload_control), so they are not stuck behind these 49 MB.
  static const bool zero_copy = [] { const char* e = getenv("RGBID_PREFETCH_KERNEL"); return e && e[0] == '1'; }();
  const bool dense = (raw_depth == dsz && raw_rgb == csz) && (dsz * B) % 16 == 0 && (csz * B) % 16 == 0 &&
                     ((uintptr_t)depth % 16 == 0) && ((uintptr_t)rgb % 16 == 0);
  if (zero_copy && dense && host_pointer_is_device_readable(depth) && host_pointer_is_device_readable(rgb)) {
    static const int ctas = [] { const char* e = getenv("RGBID_PREFETCH_CTAS"); int v = e ? atoi(e) : 0; return v > 0 ? v : 4; }();
    prefetch_copy_kernel<<<ctas, 256, 0, t->copy_stream>>>((uint4*)dd, (const uint4*)depth, dsz * B / 16);
    prefetch_copy_kernel<<<ctas, 256, 0, t->copy_stream>>>((uint4*)dc, (const uint4*)rgb, csz * B / 16);
    t->ctx->launches += 2;
  } else if (raw_depth == dsz && raw_rgb == csz) {
    RGBID_CUDA_TRY(cudaMemcpyAsync(dd, depth, dsz * B, cudaMemcpyHostToDevice, t->copy_stream));
    RGBID_CUDA_TRY(cudaMemcpyAsync(dc, rgb, csz * B, cudaMemcpyHostToDevice, t->copy_stream));
  } else {
    for (int b = 0; b < B; ++b) {
      RGBID_CUDA_TRY(cudaMemcpyAsync(dd + raw_depth * b, (const char*)depth + dsz * b, dsz, cudaMemcpyHostToDevice, t->copy_stream));
      RGBID_CUDA_TRY(cudaMemcpyAsync(dc + raw_rgb * b, rgb + csz * b, csz, cudaMemcpyHostToDevice, t->copy_stream));
    }
  }
  RGBID_CUDA_TRY(cudaEventRecord(t->ev_copy[slot], t->copy_stream));
  t->pf_depth[slot] = depth; t->pf_rgb[slot] = rgb; t->pf_valid[slot] = true; t->pf_issued_at[slot] = t->track_calls;
  return RGBID_OK;
}

int rgbid_tracker_track(rgbid_tracker* t, const uint16_t* depth, const uint8_t* rgb, int from_host,
                        rgbid_frame_result* results)
{
  if (!t || !depth || !rgb || !results) return RGBID_ERR_ARG;
  const size_t rows = t->al->cfg.rows, cols = t->al->cfg.cols;
  return track_core(t, depth, rgb, from_host, cols * 2, rows * cols * 2, cols * 3, rows * cols * 3, results);
}

int rgbid_tracker_track_device(rgbid_tracker* t, const uint16_t* depth, size_t depth_pitch, size_t depth_stride,
                               const uint8_t* rgb, size_t rgb_pitch, size_t rgb_stride, rgbid_frame_result* results)
{
  if (!t || !depth || !rgb || !results) return RGBID_ERR_ARG;
  return track_core(t, depth, rgb, 0, depth_pitch, depth_stride, rgb_pitch, rgb_stride, results);
}

static int track_core(rgbid_tracker* t, const uint16_t* depth, const uint8_t* rgb, int from_host, size_t in_dpitch,
                      size_t in_dstride, size_t in_cpitch, size_t in_cstride, rgbid_frame_result* results)
{
  rgbid_aligner* al = t->al;
  rgbid_ctx* ctx = t->ctx;
  const rgbid_align_config& c = al->cfg;
  const int B = c.batch, rows = c.rows, cols = c.cols;
  cudaStream_t s = ctx->stream;
  LaunchCtx L = ctx->L();
  const float dt_frame = t->cfg.delta_t;
  const size_t dsz = (size_t)rows * cols * 2, csz = (size_t)rows * cols * 3;

  // ---- prepareImages (src/visodo.cpp:760-773): ingest + pyramid for all streams --------------------------
  const uint16_t* d_depth = depth;
  const uint8_t* d_rgb = rgb;
  size_t dstride = in_dstride, cstride = in_cstride, dpitch = in_dpitch, cpitch = in_cpitch;
  int pf = -1;
  for (int i = 0; i < 2; ++i) {
    // expired: prefetched more than one track call ago (the caller may have refilled that host buffer since)
    if (t->pf_valid[i] && t->track_calls - t->pf_issued_at[i] > 1) t->pf_valid[i] = false;
    if (from_host && t->pf_valid[i] && t->pf_depth[i] == (const void*)depth && t->pf_rgb[i] == (const void*)rgb) pf = i;
  }
  ++t->track_calls;
  if (pf >= 0) {
    // this frame was uploaded by rgbid_tracker_prefetch while the previous one was being tracked
    dpitch = (size_t)cols * 2; cpitch = (size_t)cols * 3;
    size_t raw_depth = align_up(dsz, 256), raw_rgb = align_up(csz, 256);
    RGBID_CUDA_TRY(cudaStreamWaitEvent(s, t->ev_copy[pf], 0));
    d_depth = (const uint16_t*)t->d_prefetch[pf]; d_rgb = (const uint8_t*)(t->d_prefetch[pf] + raw_depth * B);
    dstride = raw_depth; cstride = raw_rgb;
    t->pf_valid[pf] = false;
  } else if (from_host) {
    dpitch = (size_t)cols * 2; cpitch = (size_t)cols * 3;
    size_t raw_depth = align_up(dsz, 256), raw_rgb = align_up(csz, 256);
    if (raw_depth == dsz && raw_rgb == csz) {
      RGBID_CUDA_TRY(cudaMemcpyAsync(al->d_depth_raw, depth, dsz * B, cudaMemcpyHostToDevice, s));
      RGBID_CUDA_TRY(cudaMemcpyAsync(al->d_rgb_raw, rgb, csz * B, cudaMemcpyHostToDevice, s));
    } else {
      for (int b = 0; b < B; ++b) {
        RGBID_CUDA_TRY(cudaMemcpyAsync((char*)al->d_depth_raw + raw_depth * b, (const char*)depth + dsz * b, dsz, cudaMemcpyHostToDevice, s));
        RGBID_CUDA_TRY(cudaMemcpyAsync(al->d_rgb_raw + raw_rgb * b, rgb + csz * b, csz, cudaMemcpyHostToDevice, s));
      }
    }
    d_depth = al->d_depth_raw; d_rgb = al->d_rgb_raw; dstride = raw_depth; cstride = raw_rgb;
  }
  if (!t->custom_on) {
    launch_ingest(L, d_depth, dpitch, dstride, d_rgb, cpitch, cstride, al->maps[MAP_W_CUR][0],
                  al->maps[MAP_I_CUR][0], B, c.factor_depth);
  } else {
    // prepareImagesCustomCalibration (src/visodo.cpp:775-823), stream by stream through one set of scratch maps
    const int crows = 3 * rows, ccols = 3 * cols;
    for (int b = 0; b < B; ++b) {
      launch_ingest(L, (const uint16_t*)((const char*)d_depth + dstride * b), dpitch, 0, d_rgb + cstride * b, cpitch, 0, t->cW,
                    t->cI, 1, c.factor_depth);
      launch_undistort_intensity(L, t->cI, al->view(MAP_I_CUR, 0, b), t->custom.rgb);
      launch_undistort_depthinv(L, t->cW, t->cPre, t->custom.depth, t->custom.dist);
      launch_register_depthinv(L, t->cPre, al->view(MAP_W_CUR, 0, b), t->d_canvas, t->canvas_pitch, crows, ccols, t->dRc_proj,
                               t->t_dc_proj, t->cRd_proj);
    }
  }
  aligner_current_pyramid(al, 0, B);

  // ---- first frame: everything becomes a keyframe (src/visodo.cpp:1994-2045) ---------------------------------
  if (t->frame_index == 0) {
    aligner_copy_current_to_keyframe(al, 0, B, nullptr);
    aligner_keyframe_derivatives(al, 0, B, nullptr, false);
    save_integration_keyframes(t, nullptr);
    for (int b = 0; b < B; ++b)
      RGBID_CUDA_TRY(cudaMemcpy2DAsync(t->d_colors + t->colors_sstride * b, (size_t)cols * 3, (const char*)d_rgb + cstride * b,
                                       cpitch, (size_t)cols * 3, rows, cudaMemcpyDeviceToDevice, s));
    refresh_integration_maps(t);
    launch_fill_u8(L, t->d_mask, t->mask_pitch, t->mask_sstride, rows, cols, 0, B);
    RGBID_CUDA_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < B; ++b) {
      StreamState& S = t->st[b];
      S.global_time = 1;
      rgbid_frame_result& r = results[b];
      memset(&r, 0, sizeof(r));
      set_identity(r.R, r.t); set_identity(r.dR, r.dt);
      r.visibility_odo = 1.f; r.visibility_integr = 1.f;
      r.new_odo_keyframe = 1; r.new_integr_keyframe = 1; r.frame_index = 0; r.status = RGBID_OK;
    }
    t->frame_index = 1;
    return check_last(ctx);
  }

  // ---- estimateVisualOdometry (src/visodo.cpp:944-1479) -------------------------------------------------------
  std::vector<double> prevR(9 * B), prevt(3 * B), prevcov(36 * B);
  std::vector<char> was_lost(B);
  for (int b = 0; b < B; ++b) {
    StreamState& S = t->st[b];
    was_lost[b] = S.lost ? 1 : 0;
    memcpy(&prevR[9 * b], S.dR, sizeof(double) * 9);
    memcpy(&prevt[3 * b], S.dt, sizeof(double) * 3);
    memcpy(&prevcov[36 * b], S.dcov, sizeof(double) * 36);
    double* Ri = al->h_init + 9 * b;
    double* ti = al->h_init + 9 * B + 3 * b;
    if (S.global_time > 1 && t->cfg.motion_model == RGBID_CONSTANT_VELOCITY && !S.lost) {
      // constant-velocity prediction (:1016-1027)
      double vt[3] = {S.vel[0] * dt_frame, S.vel[1] * dt_frame, S.vel[2] * dt_frame};
      double wt[3] = {S.omega[0] * dt_frame, S.omega[1] * dt_frame, S.omega[2] * dt_frame};
      double dRp[9], dtp[3], tmp[3];
      exp_map(wt, vt, dRp, dtp);
      mat3_vec(S.dR, dtp, tmp);
      for (int k = 0; k < 3; ++k) ti[k] = tmp[k] + S.dt[k];
      mat3_mul(S.dR, dRp, Ri);
    } else {
      memcpy(Ri, S.dR, sizeof(double) * 9);
      memcpy(ti, S.dt, sizeof(double) * 3);
    }
  }
  upload_control(L, al->d_init, al->h_init, sizeof(double) * 12 * B);
  // The pose is final when the iterations are: it is read back there, and the covariance pass (a whole level-0 launch
  // with the chi^2 sums) runs while the host does the pose bookkeeping below and queues the covisibility kernel behind
  // it; its results (covariance, chi^2) come back with the covisibility counts.  RGBID_NO_SPLIT_COV=1: one read-back.
  static const bool no_split = [] { const char* e = getenv("RGBID_NO_SPLIT_COV"); return e && e[0] == '1'; }();
  t->split_cov = aligner_can_split(al) && !no_split;
  int rc;
  if (t->split_cov) {
    if ((rc = aligner_enqueue_part(al, ALIGN_PART_ITERATIONS)) != RGBID_OK) return rc;
    // the pose copy goes through a stream of its own, so that the covariance pass does not queue behind it (the pass
    // only writes the covariance / chi^2 fields of the state, which this copy is not read for)
    RGBID_CUDA_TRY(cudaEventRecord(t->ev_iter_done, s));
    RGBID_CUDA_TRY(cudaStreamWaitEvent(t->rb_stream, t->ev_iter_done, 0));
    RGBID_CUDA_TRY(cudaMemcpyAsync(al->h_states, al->d_states, sizeof(GnState) * B, cudaMemcpyDeviceToHost, t->rb_stream));
    RGBID_CUDA_TRY(cudaEventRecord(t->ev_pose, t->rb_stream));
    if ((rc = aligner_enqueue_part(al, ALIGN_PART_COV)) != RGBID_OK) return rc;
    RGBID_CUDA_TRY(cudaMemcpyAsync(t->h_states_cov, al->d_states, sizeof(GnState) * B, cudaMemcpyDeviceToHost, s));
    RGBID_CUDA_TRY(cudaEventSynchronize(t->ev_pose));
  } else {
    if ((rc = aligner_enqueue_device_init(al)) != RGBID_OK) return rc;
    RGBID_CUDA_TRY(cudaMemcpyAsync(al->h_states, al->d_states, sizeof(GnState) * B, cudaMemcpyDeviceToHost, s));
    RGBID_CUDA_TRY(cudaStreamSynchronize(s));
  }
  if ((rc = check_last(ctx)) != RGBID_OK) return rc;

  // ---- pose bookkeeping (:1463-1468, :2059-2117) + covisibility transforms (:1481-1514, :2172-2186) ----------
  for (int b = 0; b < B; ++b) {
    StreamState& S = t->st[b];
    const GnState& g = al->h_states[b];
    rgbid_frame_result& r = results[b];
    memset(&r, 0, sizeof(r));
    // the reference's global_time_: frames of this stream that entered the trajectory (a frame that fails while the
    // stream is already lost does not, src/visodo.cpp:2111-2116)
    r.frame_index = S.global_time;
    r.status = g.status;
    r.chi_square = g.chi_square; r.chi_test = g.chi_test; r.ndof = g.ndof;  // (split read-back: filled in below)
    const bool ok = (g.status == RGBID_OK);
    if (ok) {
      memcpy(S.dR, g.R, sizeof(double) * 9);
      memcpy(S.dt, g.t, sizeof(double) * 3);
      memcpy(S.dcov, g.cov, sizeof(double) * 36);                            // (split read-back: filled in below)
      double Rt[9], dRc[9], dtc[3], diff[3], twist[6];
      mat3_transpose(&prevR[9 * b], Rt);
      mat3_mul(Rt, S.dR, dRc);
      for (int k = 0; k < 3; ++k) diff[k] = S.dt[k] - prevt[3 * b + k];
      mat3_vec(Rt, diff, dtc);
      log_map(dRc, dtc, twist);
      const double inv_dt = (double)(1.f / dt_frame);  // velocity_ = twist * (1.f / delta_t_), :1467-1468
      for (int k = 0; k < 3; ++k) { S.vel[k] = twist[k] * inv_dt; S.omega[k] = twist[3 + k] * inv_dt; }
      S.lost = false;
    } else {
      for (int i = 0; i < 36; ++i) S.dcov[i] = (i % 7 == 0) ? 100.0 : 0.0;
    }
    // last_estimated = last_odoKF_global * delta (:2061-2062)
    double tmp[3];
    mat3_vec(S.R_odoKF, S.dt, tmp);
    for (int k = 0; k < 3; ++k) S.t_est[k] = S.t_odoKF[k] + tmp[k];
    mat3_mul(S.R_odoKF, S.dR, S.R_est);
    // covisibility transforms: [0] cur->odoKF (K R K^-1), [1] odoKF->cur, [2] cur->integrKF, [3] integrKF->cur
    to_proj(S.dR, S.dt, c, &t->h_proj[0 * B + b]);
    to_proj_inverse(S.dR, S.dt, c, &t->h_proj[1 * B + b]);
    double Rki[9], dRi[9], dti[3], d2[3];
    mat3_inverse(S.R_intKF, Rki);
    mat3_mul(Rki, S.R_est, dRi);
    for (int k = 0; k < 3; ++k) d2[k] = S.t_est[k] - S.t_intKF[k];
    mat3_vec(Rki, d2, dti);
    to_proj(dRi, dti, c, &t->h_proj[2 * B + b]);
    to_proj_inverse(dRi, dti, c, &t->h_proj[3 * B + b]);
  }
  upload_control(L, t->d_proj, t->h_proj, sizeof(Proj) * 4 * B);
  RGBID_CUDA_TRY(cudaMemsetAsync(t->d_counts, 0, sizeof(unsigned int) * 8 * B, s));
  Proj dummy;
  memset(&dummy, 0, sizeof(dummy));
  const ImgB& Wcur = al->maps[MAP_W_CUR][0];
  const ImgB& Wkf = al->maps[MAP_W_KF][0];
  if (visibility4_applicable(Wcur, Wkf, t->intWraw)) {
    launch_visibility4(L, Wcur, Wkf, t->intWraw, t->d_proj, t->d_counts, B);  // both tests, both directions, one launch
  } else {
    launch_visibility(L, Wcur, Wkf, t->d_proj + 0 * B, dummy, t->d_counts, 0, 8, nullptr, 0, 0, B);
    launch_visibility(L, Wkf, Wcur, t->d_proj + 1 * B, dummy, t->d_counts, 2, 8, nullptr, 0, 0, B);
    launch_visibility(L, Wcur, t->intWraw, t->d_proj + 2 * B, dummy, t->d_counts, 4, 8, nullptr, 0, 0, B);
    launch_visibility(L, t->intWraw, Wcur, t->d_proj + 3 * B, dummy, t->d_counts, 6, 8, nullptr, 0, 0, B);
  }
  RGBID_CUDA_TRY(cudaMemcpyAsync(t->h_counts, t->d_counts, sizeof(unsigned int) * 8 * B, cudaMemcpyDeviceToHost, s));
  RGBID_CUDA_TRY(cudaStreamSynchronize(s));
  if (t->split_cov) {
    // the covariance pass has long finished: its part of the solver state
    for (int b = 0; b < B; ++b) {
      const GnState& g = t->h_states_cov[b];
      rgbid_frame_result& r = results[b];
      r.chi_square = g.chi_square; r.chi_test = g.chi_test; r.ndof = g.ndof;
      if (g.status == RGBID_OK) memcpy(t->st[b].dcov, g.cov, sizeof(double) * 36);
      memcpy(&al->h_states[b], &g, sizeof(GnState));  // rgbid_aligner_* accessors see the complete state
    }
  }

  // ---- keyframe decisions (:2175-2215) ----------------------------------------------------------------------------
  bool any_odo = false, any_int = false, any_fuse = false, any_mask = false;
  for (int b = 0; b < B; ++b) {
    StreamState& S = t->st[b];
    rgbid_frame_result& r = results[b];
    const unsigned int* cnt = t->h_counts + 8 * b;
    auto ratio = [](unsigned vis, unsigned val) { return ((float)val < 1.f) ? 0.f : (float)vis / (float)val; };
    float vis_odo = fminf(ratio(cnt[2], cnt[3]), ratio(cnt[0], cnt[1]));
    float vis_int = fminf(ratio(cnt[6], cnt[7]), ratio(cnt[4], cnt[5]));
    r.visibility_odo = vis_odo; r.visibility_integr = vis_int;
    int new_odo = 0, new_int = 0, fuse = 0, mask = 0;
    // `again`: the alignment failed while the stream was already lost (src/visodo.cpp:2099-2116) -- both keyframes are
    // re-saved from the current frame and nothing else happens: no keyframe reset, no keyframe or constraint for the
    // back end, global_time_ stands still
    const bool again = (r.status != RGBID_OK) && was_lost[b];
    if (r.status != RGBID_OK) {
      // lost: both keyframes are re-initialised from the current frame (:2066-2097, :2111-2116)
      S.lost = true;
      new_odo = 1; new_int = 1;
    } else {
      S.odoKF_count++; S.integrKF_count++;
      new_odo = (S.odoKF_count >= t->cfg.max_odo_kf_count) || (vis_odo < t->cfg.visratio_odo);
      new_int = (S.integrKF_count >= t->cfg.max_integr_kf_count) || (vis_int < t->cfg.visratio_integr);
      fuse = !new_int;
      mask = new_int;  // computeOverlapping runs on the covisibility path only (:2199), never for a lost frame
    }
    memcpy(r.R, S.R_est, sizeof(double) * 9); memcpy(r.t, S.t_est, sizeof(double) * 3);
    memcpy(r.dR, S.dR, sizeof(double) * 9); memcpy(r.dt, S.dt, sizeof(double) * 3);
    memcpy(r.cov, S.dcov, sizeof(double) * 36);
    r.new_odo_keyframe = new_odo; r.new_integr_keyframe = new_int;
    r.lost_again = again ? 1 : 0;
    if (r.status == RGBID_OK) {
      // odometry-keyframe constraint (KF, i) -> sequential constraint (i - 1, i) (:2126-2156)
      relative_constraint(&prevR[9 * b], &prevt[3 * b], &prevcov[36 * b], S.dR, S.dt, S.dcov, r.seq_R, r.seq_t, r.seq_cov);
    } else {
      // dummy constraint: zero motion, very high covariance (:2068-2071); not pushed when lost_again
      set_identity(r.seq_R, r.seq_t);
      for (int i = 0; i < 36; ++i) r.seq_cov[i] = (i % 7 == 0) ? 100.0 : 0.0;
    }
    if (new_odo && !again) {
      // resetOdometryKeyframe (:1541-1575)
      chain_compose(S);
      S.odoKF_count = 0;
      memcpy(S.R_odoKF, S.R_est, sizeof(double) * 9); memcpy(S.t_odoKF, S.t_est, sizeof(double) * 3);
      set_identity(S.dR, S.dt);
      memset(S.dcov, 0, sizeof(S.dcov));
    }
    if (new_int && !again) {
      // resetIntegrationKeyframe (:1577-1672): close the chain, hand the outgoing keyframe over with its SEQ_KF
      // constraint, switch to the new keyframe
      chain_compose(S);
      double kR[9], kt[3], kcov[36];
      relative_constraint(S.o2i_last_R, S.o2i_last_t, S.o2i_last_cov, S.o2i_next_R, S.o2i_next_t, S.o2i_next_cov, kR, kt, kcov);
      if (t->sink) {
        int hrc = hand_off_keyframe(t, b, kR, kt, kcov, r.frame_index);
        if (hrc != RGBID_OK) return hrc;
      }
      S.integrKF_count = 0;
      S.last_integrKF_index = r.frame_index;
      memcpy(S.R_intKF, S.R_est, sizeof(double) * 9); memcpy(S.t_intKF, S.t_est, sizeof(double) * 3);
      memcpy(S.o2i_last_R, S.dR, sizeof(double) * 9); memcpy(S.o2i_last_t, S.dt, sizeof(double) * 3);
      memcpy(S.o2i_last_cov, S.dcov, sizeof(double) * 36);
      set_identity(S.o2i_next_R, S.o2i_next_t);
      memset(S.o2i_next_cov, 0, sizeof(S.o2i_next_cov));
    }
    t->h_flags[3 * B + b] = mask;
    any_mask |= (mask != 0);
    if (!again) S.global_time++;
    t->h_flags[0 * B + b] = new_odo; t->h_flags[1 * B + b] = new_int; t->h_flags[2 * B + b] = fuse;
    any_odo |= (new_odo != 0); any_int |= (new_int != 0); any_fuse |= (fuse != 0);
  }
  upload_control(L, t->d_flags, t->h_flags, sizeof(int) * 4 * B);
  if (any_odo) {
    // saveCurrentImagesAsOdoKeyframes (:826-878), predicated per stream
    aligner_copy_current_to_keyframe(al, 0, B, t->d_flags + 0 * B);
    aligner_keyframe_derivatives(al, 0, B, t->d_flags + 0 * B, false);
  }
  if (any_mask)
    // computeOverlapping (:1517-1539): mask of the new keyframe (current frame) against the old raw keyframe
    launch_visibility(L, Wcur, t->intWraw, t->d_proj + 2 * B, dummy, t->d_counts, 4, 8, t->d_mask, t->mask_pitch,
                      t->mask_sstride, B, t->d_flags + 3 * B);
  if (any_int) {
    save_integration_keyframes(t, t->d_flags + 1 * B);
    for (int b = 0; b < B; ++b)
      if (t->h_flags[1 * B + b])
        RGBID_CUDA_TRY(cudaMemcpy2DAsync(t->d_colors + t->colors_sstride * b, (size_t)cols * 3, (const char*)d_rgb + cstride * b,
                                         cpitch, (size_t)cols * 3, rows, cudaMemcpyDeviceToDevice, s));
  }
  if (any_fuse) {
    // integrateImagesIntoKeyframes (:1674-1764): K6 + K7 fused; transform = integrKF->cur (h_proj[3])
    launch_warp_integrate(L, Wcur, t->intW, t->intWeight, t->wstate, t->d_proj + 3 * B, B, t->d_flags + 2 * B);
  }
  refresh_integration_maps(t);
  t->frame_index++;
  // rgbid_tracker_prefetch orders its next upload into the staging buffers after everything queued above
  if (!t->ev_track_done_valid) {
    RGBID_CUDA_TRY(cudaEventCreateWithFlags(&t->ev_track_done, cudaEventDisableTiming));
    t->ev_track_done_valid = true;
  }
  RGBID_CUDA_TRY(cudaEventRecord(t->ev_track_done, s));
  return check_last(ctx);
}

}  // extern "C"
