// kernels.cuh -- internal launchers (C++ linkage) shared by the C-ABI layer.
#pragma once
#include <mutex>
#include "common.cuh"

namespace rgbid {

// Launch bookkeeping: every launcher bumps *launches (bench.py's gpu_launches).
struct LaunchCtx {
  cudaStream_t stream;
  long long* launches;
  int num_sms;
  // programmatic dependent launch: the Gauss-Newton kernels are launched with the stream-serialisation attribute, so
  // the next kernel's CTAs become resident (and run their pose-independent prologue) while the previous one is in
  // its serial tail; every such kernel executes griddepcontrol.wait before it reads what a predecessor wrote
  bool pdl = false;
};

// Launch with (or without) cudaLaunchAttributeProgrammaticStreamSerialization; works under stream capture (the graph
// gets programmatic dependency edges).
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     bool pdl, Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// One-time set-up that lives in a device's context (constant tables, kernel attributes) has to happen once per DEVICE
// the process uses, not once per process, and two host threads may get here together (the tracker and the keyframe
// aligner call concurrently).  `once(f)` runs f the first time it is called with a given device current; f
// returns whether it succeeded, and a failed set-up is retried by the next call instead of being marked done.
class PerDevice {
 public:
  template <class F>
  void once(F&& f)
  {
    std::lock_guard<std::mutex> guard(mu_);
    const unsigned long long bit = 1ull << (current() & 63);
    if (!(done_ & bit) && f()) done_ |= bit;
  }
  // grow-only variant: runs f(value) whenever `value` exceeds what this device has been configured for
  template <class F>
  void at_least(size_t value, F&& f)
  {
    std::lock_guard<std::mutex> guard(mu_);
    size_t& have = level_[current() & 63];
    if (value > have) { f(value); have = value; }
  }

 private:
  static int current() { int dev = 0; cudaGetDevice(&dev); return dev; }
  std::mutex mu_;
  unsigned long long done_ = 0;
  size_t level_[64] = {};
};

// ---- image_ops.cu ---------------------------------------------------------------------------------
void launch_depth_to_invdepth(const LaunchCtx& L, const uint16_t* src, size_t spitch, size_t sstride, ImgB dst,
                              int batch, float factor_depth);
void launch_intensity(const LaunchCtx& L, const uint8_t* rgb, size_t spitch, size_t sstride, ImgB dst, int batch);
void launch_ingest(const LaunchCtx& L, const uint16_t* depth, size_t dpitch, size_t dstride, const uint8_t* rgb,
                   size_t cpitch, size_t cstride, ImgB W, ImgB I, int batch, float factor_depth);
void launch_decompose_rgb(const LaunchCtx& L, const uint8_t* rgb, size_t spitch, ImgB r, ImgB g, ImgB b);
// two maps (A, B) down-sampled in one launch; B may have p == nullptr
void launch_pyr_down2(const LaunchCtx& L, ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch,
                      const int* active = nullptr);
void launch_gradient2(const LaunchCtx& L, ImgB srcA, ImgB gxA, ImgB gyA, ImgB srcB, ImgB gxB, ImgB gyB, int batch,
                      const int* active = nullptr);
// one launch for up to 16 maps of different sizes (all levels of both keyframe pyramids); false = not applicable
// (unaligned maps): the caller falls back to the per-level launches
bool launch_gradient_list(const LaunchCtx& L, const ImgB* src, const ImgB* gx, const ImgB* gy, int n, int batch,
                          const int* active = nullptr);
// vertex map + Sobel gradients + normal map of a (fused) inverse-depth map in one pass; false = unaligned maps
bool launch_keyframe_maps(const LaunchCtx& L, ImgB depth_inv, ImgB gx, ImgB gy, ImgB vmap, ImgB nmap, float fx, float fy,
                          float cx, float cy, int batch);
bool launch_copy_list(const LaunchCtx& L, const ImgB* src, const ImgB* dst, int n, int batch, const int* active = nullptr);
void launch_bilateral2(const LaunchCtx& L, ImgB srcA, ImgB dstA, float sigmaA, ImgB srcB, ImgB dstB, float sigmaB,
                       int batch, const int* active = nullptr);
void launch_copy2(const LaunchCtx& L, ImgB srcA, ImgB dstA, ImgB srcB, ImgB dstB, int batch,
                  const int* active = nullptr);
void launch_fill(const LaunchCtx& L, ImgB dst, float value, int batch, const int* active = nullptr);
void launch_fill_u8(const LaunchCtx& L, uint8_t* dst, size_t pitch, size_t sstride, int rows, int cols,
                    uint8_t value, int batch, const int* active = nullptr);
void launch_vmap(const LaunchCtx& L, ImgB depth_inv, ImgB vmap, float fx, float fy, float cx, float cy, int batch,
                 const int* active = nullptr);
void launch_nmap_gradients(const LaunchCtx& L, ImgB depth_inv, ImgB gx, ImgB gy, ImgB nmap, float fx, float fy,
                           float cx, float cy, int batch, const int* active = nullptr);

// ---- warp_ops.cu ----------------------------------------------------------------------------------
void launch_warp_invdepth(const LaunchCtx& L, ImgB src, ImgB prev, ImgB dst, const Proj& P);
void launch_warp_intensity(const LaunchCtx& L, ImgB src, ImgB prev, ImgB dst, const Proj& P);
// batched weighted warp (+ optional fused integration): per-stream Proj read from device memory
void launch_warp_invdepth_weighted(const LaunchCtx& L, ImgB src, ImgB prev, ImgB dst, ImgB weight, const Proj* P_dev,
                                   Proj P_host, int batch, const int* active = nullptr);
// batched K4 + K5 of tracker mode in one pass (warped inverse depth, then intensity warped with it as geometry), per-
// stream projection read from the device-resident Gauss-Newton state (states[first + b].proj[0]); texW / texI: one
// texture object per stream over src_w (point, border) / src_i (linear, clamp), or null for the software sampler
struct GnState;
void launch_warp_pair(const LaunchCtx& L, ImgB src_w, ImgB src_i, const cudaTextureObject_t* texW,
                      const cudaTextureObject_t* texI, ImgB kf_w, const GnState* states, ImgB dst_w, ImgB dst_i,
                      int first, int batch);
void launch_integrate(const LaunchCtx& L, ImgB wsrc, ImgB wweight, ImgB dst, ImgB dweight, int batch,
                      const int* active = nullptr);
// fused K6 + K7: warp current inverse depth into the keyframe and fuse it in the same pass
void launch_warp_integrate(const LaunchCtx& L, ImgB cur, ImgB kf, ImgB kf_weight, ImgB warped_weight_state,
                           const Proj* P_dev, int batch, const int* active);
// visibility: counts[b*4 + {0 visible,1 valid}] (+2,+3 for the second direction) are accumulated with
// integer atomics; mask may be null.  P_dev holds one Proj per stream (per direction).
void launch_visibility(const LaunchCtx& L, ImgB src, ImgB dst, const Proj* P_dev, Proj P_host,
                       unsigned int* counts, int count_offset, int count_stride, uint8_t* mask, size_t mpitch,
                       size_t mstride, int batch, const int* active = nullptr);
// the four covisibility passes of a tracked frame in one launch (counts[b * 8 + 2 * pass + {visible, valid}], P[pass * batch + b])
void launch_visibility4(const LaunchCtx& L, ImgB cur, ImgB kf, ImgB ikf, const Proj* P_dev, unsigned int* counts, int batch);
bool visibility4_applicable(const ImgB& cur, const ImgB& kf, const ImgB& ikf);

// ---- calib_ops.cu ---------------------------------------------------------------------------------
void launch_undistort_intensity(const LaunchCtx& L, ImgB src, ImgB dst, const rgbid_intr& intr);
void launch_undistort_depthinv(const LaunchCtx& L, ImgB src, ImgB dst, const rgbid_intr& intr, const rgbid_depth_dist& dp);
// canvas: int map of crows x ccols (the reference uses 3 rows x 3 cols) with pitch cpitch bytes, cleared inside
void launch_register_depthinv(const LaunchCtx& L, ImgB src, ImgB dst, int* canvas, size_t cpitch, int crows, int ccols,
                              const float* dRc_proj, const float* t_dc_proj, const float* cRd_proj);
void launch_integrate_rgb(const LaunchCtx& L, ImgB dw, ImgB rw, ImgB gw, ImgB bw, ImgB ww, ImgB dd, uint8_t* colors,
                          size_t cpitch, ImgB wd);
void launch_generate_image(const LaunchCtx& L, ImgB vmap, ImgB nmap, const uint8_t* rgb, size_t rgb_pitch, const float* light,
                           uint8_t* out, size_t out_pitch, int rows, int cols);

// ---- scale_est.cu ---------------------------------------------------------------------------------
void launch_compute_error(const LaunchCtx& L, ImgB im1, ImgB im0, float* error, int kept_rows, int kept_cols,
                          int stride);

enum ScaleOp { SCALE_SIGMA_NU = 0, SCALE_NU_ONLY = 1, SCALE_SIGMA_PDF = 2 };

// Result of the scale estimation for one frame pair (device resident)
struct ScaleState {
  float bias_int, sigma_int, nu_int;
  float bias_depthinv, sigma_depthinv, nu_depthinv;
  int irls_iters_int, irls_iters_depthinv;
};

// Scale estimation on explicit error vectors (bridge API); slot 1 may be disabled with err1 == nullptr.
void launch_scale_from_errors(const LaunchCtx& L, const float* err0, const float* err1, int n, int op, int mest,
                              float bias0, float sigma0, float bias1, float sigma1, ScaleState* out);
void launch_chi_square(const LaunchCtx& L, const float* err_int, const float* err_depth, int n, float sigma_int,
                       float sigma_depth, int mest, double* out2 /* [sum_rho, n_valid] */);

// ---- gn_system.cu ---------------------------------------------------------------------------------
struct GnLevelMaps {
  ImgB W0, I0, gWx, gWy, gIx, gIy;  // keyframe maps of this level
  ImgB Wc, Ic;                      // current-frame maps of this level
  // optional texture objects over Wc (point) / Ic (linear), one per stream; null -> software sampler
  const cudaTextureObject_t* texW;
  const cudaTextureObject_t* texI;
  int tex_border;  // 1: texW was created with border addressing (0 outside), which the fast build kernel relies on
};

// Device-resident state of one frame pair's Gauss-Newton problem
struct GnState {
  double R[9], t[3];         // current estimate _{KF}T^{cur}
  double R0[9], t0[3];       // initial guess (restored if the pose goes NaN)
  double cov[36];
  double lastA[36];
  Proj proj[RGBID_MAX_LEVELS];  // K_l R^-1 K_l^-1, -K_l R^-1 t for every level, refreshed after each update
  int status;                // RGBID_OK | RGBID_ERR_NAN
  int iter_count;            // trace cursor
  float chi_square, chi_test, ndof;
  // cfg.termination: level whose remaining launches are no-ops for this pair (-1: none), iterations executed per
  // level, pose before the last update (CHI_SQUARED undo) and the RMSE of the previous test
  int skip_level;
  int iters_done[RGBID_MAX_LEVELS];
  double Rprev[9], tprev[3];
  float rmse_prev;
  int trace_on;              // 0: the solver tail skips the per-iteration trace record (set per run by gn_init_kernel)
};

struct GnParams {
  int batch, level, rows, cols;
  float fx, fy, cx, cy;         // level intrinsics
  float fx0, fy0, cx0, cy0;     // level-0 intrinsics (to refresh proj[] for all levels)
  int levels;
  int mode;                     // RGBID_MODE_*
  int mestimator, weighting;
  int student_nu;               // 1: Student weights with estimated nu; 0: computeWeight(mestimator)
  int use_scale;                // 1: read ScaleState; 0: sigma 5 / 0.0025, nu 5, bias 0
  int update_pose;              // 1: solve + update pose in the last block; 0: covariance pass
  int compute_cov;              // 1: cov = A^-1 in the last block
  int chi_mestimator;           // covariance pass: M-estimator of the end-of-frame chi^2 (-1: skip)
  int iter_index;               // for the trace
  int trace_stride;             // entries per pair (0: no trace)
  int kept_rows, kept_cols, sample_stride;  // residual sampling geometry of this level
  int sigma_op;                 // ScaleOp used by the sampling kernel
  int next_level;               // level of the launch that consumes the updated pose (-1: refresh proj[] of all levels)
  int first;                    // first frame pair of this launch (the batch may be launched in groups, see aligner.cu)
  int batch_total;              // pairs of the whole batch (grid sizing); 0: same as batch
  int sched_level;              // level of the schedule step this launch belongs to (the termination test of a coarse
                                // level runs on level-0 maps); a pair with skip_level == sched_level is skipped
  int sched_iter;               // iteration of sched_level this launch belongs to
  int termination;              // cfg.termination
  int chi_test;                 // CHI_SQUARED termination test launch: 1 = record the RMSE, 2 = compare, undo, end level
  float conv_eps;               // > 0: CONVERGENCE termination (|x| < conv_eps ends the level)
  int dry_tail;                 // measurement hook (rgbid_aligner_time_build): run the whole tail -- final sum, solve, pose
                                // update, projection refresh -- but do not commit it, so that the launch can be repeated
  int prewarped;                // 1: M.Wc / M.Ic already are the current frame warped into the keyframe view at this
                                //    level (WARP_ORDER = warpFirst, see aligner.cu): read them pixel for pixel
};

// fused warp + sample + residual -> IRLS sigma / nu (8-CTA cluster per pair)
void launch_gn_scale(const LaunchCtx& L, const GnLevelMaps& M, const GnParams& P, GnState* states,
                     ScaleState* scales);
// fused warp + bilinear sample + residual + Jacobian + weights + 27-sum reduction (+ solve / covariance in
// the last block of each pair)
void launch_gn_build(const LaunchCtx& L, const GnLevelMaps& M, const GnParams& P, GnState* states,
                     const ScaleState* scales, double* partials, int partial_stride, unsigned int* counters,
                     rgbid_iter_trace* trace);
void launch_gn_init(const LaunchCtx& L, GnState* states, const double* R_init, const double* t_init, int batch,
                    int levels, float fx0, float fy0, float cx0, float cy0, const int* trace_flag);
int gn_build_grid_x(int rows, int cols, int batch, int num_sms);
// One-time per-device set-up of the Gauss-Newton kernels (nu table in constant memory, dynamic shared-memory limits).
// rgbid_aligner_create calls it, so that nothing of it can fall inside a stream capture; returns a cudaError_t value.
int gn_prepare_device();
// [batch][48] doubles (cov 36, R 9, t 3) from the solver state; NaN for lost pairs
void launch_export_systems(const LaunchCtx& L, const GnState* states, double* out, int batch);

// un-fused drop-in (pre-warped W1 / I1), one pair
void launch_build_system(const LaunchCtx& L, ImgB W0, ImgB I0, ImgB gWx, ImgB gWy, ImgB gIx, ImgB gIy, ImgB W1,
                         ImgB I1, const rgbid_system_params& sp, double* partials, unsigned int* counter,
                         double* out27);

}  // namespace rgbid
