/*
 * rgbid_b200.h -- C ABI of the B200-native dense frame-to-keyframe alignment path.
 *
 * This is the drop-in boundary.  The reference (dangut/RGBiD-SLAM) has no FFI layer: its seam is the
 * C++ free-function API RGBID_SLAM::device::* declared in src/internal.h:187-453 and consumed only by
 * src/visodo.cpp and src/keyframe_align.cpp.  Every entry point below replaces one of those bridge
 * functions (cited per function) or fuses a sequence of them; include/rgbid_b200/internal.hpp
 * re-exposes the reference's exact C++ signatures on top of this ABI.
 *
 * Conventions
 *  - plain pointers and sizes only; all image pointers are DEVICE pointers unless the name ends in
 *    _host; `pitch` is the row stride in BYTES (the reference's DeviceArray2D::step()); images are
 *    float32 with NaN = invalid, RGB is 3 x uint8 interleaved, depth is uint16 millimetres;
 *  - transforms are passed in "pixel space" exactly as the reference passes them: Rp = K R K^-1
 *    (row-major 3x3 float), tp = K t (3 floats);
 *  - every call takes a context (device, stream, pre-allocated scratch: no hidden cudaMalloc on the
 *    hot path) and returns an int status: 0 ok, RGBID_ERR_* (< 0), or 1000 + cudaError_t;
 *    nothing calls exit() (the reference's cudaSafeCall does, ThirdParty/pcl_gpu_containers/src/error.cpp:42-46);
 *  - functions that return scalars through host pointers synchronise the context's stream before
 *    returning (the reference synchronises in every bridge function); all others are asynchronous
 *    on the context's stream;
 *  - one context per host thread; contexts on different streams may be used concurrently.
 */
#ifndef RGBID_B200_H_
#define RGBID_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGBID_B200_VERSION 100

#define RGBID_OK 0
#define RGBID_ERR_NAN (-1)      /* numerical failure: pose went NaN (reference: visodo.cpp:1265-1274) */
#define RGBID_ERR_ARG (-2)      /* bad argument (null pointer, size mismatch, unsupported size) */
#define RGBID_ERR_NOMEM (-3)
#define RGBID_ERR_STATE (-4)    /* call sequence error (e.g. track before keyframe) */
#define RGBID_ERR_TIMEOUT (-5)  /* device-side wait exceeded its budget */
#define RGBID_ERR_CUDA_BASE 1000

/* enums of src/internal.h:66-72 (same numeric values) */
enum { RGBID_LSQ = 0, RGBID_HUBER = 1, RGBID_TUKEY = 2, RGBID_STUDENT = 3 };
enum { RGBID_NO_MM = 0, RGBID_CONSTANT_VELOCITY = 1 };
enum { RGBID_SIGMA_MAD = 0, RGBID_SIGMA_PDF = 1, RGBID_SIGMA_CONS = 2 };
enum { RGBID_INDEPENDENT = 0, RGBID_MIN_WEIGHT = 1, RGBID_GEOM_ONLY = 2, RGBID_PHOT_ONLY = 3 };
enum { RGBID_WARP_FIRST = 0, RGBID_PYR_FIRST = 1 };
enum { RGBID_CHI_SQUARED = 0, RGBID_ALL_ITERS = 1 };
enum { RGBID_NO_FILTERS = 0, RGBID_FILTER_GRADS = 1 };

enum { RGBID_MODE_TRACKER = 0, RGBID_MODE_ALIGN = 1 };

#define RGBID_MAX_LEVELS 8
#define RGBID_SYSTEM_SIZE 27 /* TOTAL_SIZE, src/internal.h:59-64 */

typedef struct rgbid_ctx rgbid_ctx;
typedef struct rgbid_aligner rgbid_aligner;
typedef struct rgbid_tracker rgbid_tracker;

/* ---------------------------------------------------------------------------------------------- */
/* Context                                                                                          */
/* ---------------------------------------------------------------------------------------------- */

/* stream: a cudaStream_t owned by the caller, or NULL to let the context create its own. */
int rgbid_ctx_create(rgbid_ctx** ctx, int device, void* stream);
int rgbid_ctx_destroy(rgbid_ctx* ctx);
int rgbid_ctx_sync(rgbid_ctx* ctx);          /* replaces device::sync(), src/internal.h:456-457 */
void* rgbid_ctx_stream(rgbid_ctx* ctx);
int rgbid_version(void);
const char* rgbid_status_string(int status);
/* Number of kernels launched through this context since creation (bench.py's gpu_launches). */
long long rgbid_ctx_launch_count(rgbid_ctx* ctx);

/* ---------------------------------------------------------------------------------------------- */
/* Image preparation                                                                                */
/* ---------------------------------------------------------------------------------------------- */

/* convertDepth2InvDepth, src/internal.h:224 (src/cuda/misc.cu:365-374) */
int rgbid_convert_depth_to_invdepth(rgbid_ctx* ctx, const uint16_t* src, size_t src_pitch, float* dst,
                                    size_t dst_pitch, int rows, int cols, float factor_depth);
/* computeIntensity, src/internal.h:231 (src/cuda/misc.cu:377-386) */
int rgbid_compute_intensity(rgbid_ctx* ctx, const uint8_t* rgb, size_t src_pitch, float* dst, size_t dst_pitch,
                            int rows, int cols);
/* decomposeRGBInChannels, src/internal.h:234 (src/cuda/misc.cu:388-397) */
int rgbid_decompose_rgb(rgbid_ctx* ctx, const uint8_t* rgb, size_t src_pitch, float* r, float* g, float* b,
                        size_t dst_pitch, int rows, int cols);
/* pyrDownIntensity / pyrDownDepth, src/internal.h:195,209 (src/cuda/pyrdown.cu:194-242); dst is
 * (src_rows/2) x (src_cols/2) */
int rgbid_pyr_down(rgbid_ctx* ctx, const float* src, size_t src_pitch, int src_rows, int src_cols, float* dst,
                   size_t dst_pitch);
/* computeGradientIntensity / computeGradientDepth, src/internal.h:242,250 (src/cuda/misc.cu:400-441) */
int rgbid_compute_gradient(rgbid_ctx* ctx, const float* src, size_t src_pitch, int rows, int cols, float* grad_x,
                           float* grad_y, size_t grad_pitch);
/* bilateralFilter, src/internal.h:434 (src/cuda/filters.cu:139-162) */
int rgbid_bilateral_filter(rgbid_ctx* ctx, const float* src, size_t src_pitch, int rows, int cols, float* dst,
                           size_t dst_pitch, float sigma_floatmap);
/* copyImage / copyImages, src/internal.h:260-264 (src/cuda/misc.cu:446-480) */
int rgbid_copy_image(rgbid_ctx* ctx, const float* src, size_t src_pitch, float* dst, size_t dst_pitch, int rows,
                     int cols);
/* initialiseDeviceMemory2D<float>, initialiseWeightKeyframe, src/internal.h:271-274 */
int rgbid_fill_image(rgbid_ctx* ctx, float* dst, size_t dst_pitch, int rows, int cols, float value);
/* createVMap, src/internal.h:382 (src/cuda/maps.cu:300-344); vmap is (3*rows) x cols */
int rgbid_create_vmap(rgbid_ctx* ctx, const float* depth_inv, size_t pitch, int rows, int cols, float fx, float fy,
                      float cx, float cy, float* vmap, size_t vmap_pitch);
/* createNMapGradients, src/internal.h:388 (src/cuda/maps.cu:396-443); nmap is (3*rows) x cols */
int rgbid_create_nmap_gradients(rgbid_ctx* ctx, const float* depth_inv, const float* grad_x, const float* grad_y,
                                size_t pitch, int rows, int cols, float fx, float fy, float cx, float cy,
                                float* nmap, size_t nmap_pitch);

/* ---------------------------------------------------------------------------------------------- */
/* Warping, visibility, fusion                                                                      */
/* ---------------------------------------------------------------------------------------------- */

/* warpInvDepthWithTrafo3D, src/internal.h:345-347 (src/cuda/warping_registration.cu:971-1019) */
int rgbid_warp_invdepth(rgbid_ctx* ctx, const float* src, size_t src_pitch, const float* depth_prev,
                        size_t prev_pitch, float* dst, size_t dst_pitch, int rows, int cols, const float* Rp,
                        const float* tp);
/* warpIntensityWithTrafo3DInvDepth, src/internal.h:333-334 (warping_registration.cu:920-967);
 * bilinear sampling with the texture unit's 1/256 weight quantisation reproduced in software */
int rgbid_warp_intensity(rgbid_ctx* ctx, const float* src, size_t src_pitch, const float* depth_prev,
                         size_t prev_pitch, float* dst, size_t dst_pitch, int rows, int cols, const float* Rp,
                         const float* tp);
/* warpInvDepthWithTrafo3DWeighted, src/internal.h:349-351 (warping_registration.cu:1021-1069) */
int rgbid_warp_invdepth_weighted(rgbid_ctx* ctx, const float* src, size_t src_pitch, const float* depth_prev,
                                 size_t prev_pitch, float* dst, size_t dst_pitch, float* weight_warped,
                                 size_t weight_pitch, int rows, int cols, const float* Rp, const float* tp);
/* integrateWarpedFrame, src/internal.h:357-359 (warping_registration.cu:1072-1095) */
int rgbid_integrate_warped_frame(rgbid_ctx* ctx, const float* warped_depthinv, size_t wd_pitch,
                                 const float* warped_weight, size_t ww_pitch, float* depthinv_dst, size_t dd_pitch,
                                 float* weight_dst, size_t dw_pitch, int rows, int cols);
/* getVisibilityRatio / getVisibilityRatioWithOverlapMask, src/internal.h:370-377
 * (warping_registration.cu:825-913).  overlap_mask may be NULL.  Synchronous. */
int rgbid_visibility_ratio(rgbid_ctx* ctx, const float* depth_src, size_t src_pitch, const float* depth_dst,
                           size_t dst_pitch, int rows, int cols, const float* Rp, const float* tp,
                           uint8_t* overlap_mask, size_t mask_pitch, float* ratio_host);

/* ---------------------------------------------------------------------------------------------- */
/* Residual sampling, scale estimation, chi-square                                                  */
/* ---------------------------------------------------------------------------------------------- */

/* Sampling geometry of computeErrorGridStride (src/cuda/sigmaFuncs.cu:711-747). Host only. */
int rgbid_error_geometry(int rows, int cols, int min_nsamples, int* kept_rows, int* kept_cols, int* stride);
/* computeErrorGridStride, src/internal.h:281 (sigmaFuncs.cu:701-765); error must hold
 * kept_rows*kept_cols floats; *n_out (host, may be NULL) receives that count */
int rgbid_compute_error(rgbid_ctx* ctx, const float* im1, size_t pitch1, const float* im0, size_t pitch0, int rows,
                        int cols, int min_nsamples, float* error, int* n_out);
/* computeSigmaAndNuStudent, src/internal.h:291-292 (sigmaFuncs.cu:858-1066). bias/sigma in-out. Synchronous. */
int rgbid_sigma_nu_student(rgbid_ctx* ctx, const float* error, int n, float* bias_host, float* sigma_host,
                           float* nu_host, int mestimator);
/* computeNuStudent, src/internal.h:294-295 (sigmaFuncs.cu:1068-1222). Synchronous. */
int rgbid_nu_student(rgbid_ctx* ctx, const float* error, int n, float bias, float sigma, float* nu_host);
/* computeSigmaPdf, src/internal.h:288-289 (sigmaFuncs.cu:773-854). Synchronous. */
int rgbid_sigma_pdf(rgbid_ctx* ctx, const float* error, int n, float* bias_host, float* sigma_host,
                    int mestimator);
/* computeChiSquare, src/internal.h:285-286 (sigmaFuncs.cu:1225-1297). Synchronous. */
int rgbid_chi_square(rgbid_ctx* ctx, const float* error_int, const float* error_depth, int n, float sigma_int,
                     float sigma_depth, int mestimator, float* chi_square_host, float* chi_test_host,
                     float* ndof_host);

/* ---------------------------------------------------------------------------------------------- */
/* Normal equations                                                                                 */
/* ---------------------------------------------------------------------------------------------- */

typedef struct {
  float fx, fy, cx, cy; /* Intr of the level being processed (src/internal.h:119-140) */
  int mestimator;       /* used when student_nu == 0 (buildSystemGridStride) */
  int weighting;
  int student_nu;       /* 1: buildSystemStudentNuGridStride, 0: buildSystemGridStride */
  float sigma_depthinv, sigma_int, bias_depthinv, bias_int, nu_depthinv, nu_int;
} rgbid_system_params;

/* buildSystemGridStride / buildSystemStudentNuGridStride, src/internal.h:299-322
 * (src/cuda/estimate_VO.cu:505-789): one launch (per-thread FP32 accumulation, FP64 from the warp
 * level up, deterministic last-block final sum) instead of two kernels + two syncs.
 * A36: 36 doubles row-major with both triangles filled, b6: 6 doubles (host). Synchronous. */
int rgbid_build_system(rgbid_ctx* ctx, const float* W0, const float* I0, const float* gradW0_x,
                       const float* gradW0_y, const float* gradI0_x, const float* gradI0_y, const float* W1,
                       const float* I1, size_t pitch, int rows, int cols, const rgbid_system_params* params,
                       double* A36_host, double* b6_host);
/* Same with one row pitch per map, pitch8[i] for the i-th pointer argument (PtrStep<float>::step of each map,
 * src/internal.h:360-384: maps from cudaMallocPitch and wrapped dense maps may be mixed). */
int rgbid_build_system_pitched(rgbid_ctx* ctx, const float* W0, const float* I0, const float* gradW0_x,
                               const float* gradW0_y, const float* gradI0_x, const float* gradI0_y, const float* W1,
                               const float* I1, const size_t* pitch8, int rows, int cols,
                               const rgbid_system_params* params, double* A36_host, double* b6_host);

/* ---------------------------------------------------------------------------------------------- */
/* Fused, device-resident coarse-to-fine alignment of `batch` independent frame pairs               */
/* (the Gauss-Newton loops of VisodoTracker::estimateVisualOdometry, src/visodo.cpp:944-1479, and   */
/*  KeyframeAlign::alignKeyframes, src/keyframe_align.cpp:115-357)                                  */
/* ---------------------------------------------------------------------------------------------- */

typedef struct {
  int rows, cols;   /* level-0 size; must be divisible by 2^(levels-1) */
  int levels;       /* VisodoTracker::LEVELS = 3 (include/visodo.h:52), KeyframeAlign::LEVELS = 4 */
  int finest_level;
  int iterations[RGBID_MAX_LEVELS]; /* per level; {10,5,3} tracker, {5,5,3,0} align */
  int batch;        /* independent frame pairs processed per call */
  int mode;         /* RGBID_MODE_TRACKER | RGBID_MODE_ALIGN */
  int mestimator;   /* tracker: Mestimator_ */
  int weighting;
  int sigma_estimator; /* tracker only: RGBID_SIGMA_PDF | RGBID_SIGMA_CONS */
  int nsamples;     /* residual sub-sampling target: 10000 tracker, 19200 align */
  float fx, fy, cx, cy; /* level-0 intrinsics */
  float factor_depth;
  int with_fusion;  /* tracker: allocate integration-keyframe buffers */
  int warp_first;   /* tracker: 1 = WARP_ORDER warpFirst (src/visodo.cpp:1078-1105: above level 0 every iteration warps
                     * at level 0 and rebuilds the pyramid of the warped maps), 0 = pyrFirst (the shipped
                     * config_data/visodoRGBDconfig.ini).  NOT the reference's enum value: a zeroed config is pyrFirst. */
  int termination;  /* RGBID_TERM_*: how a pyramid level ends before its iterations[] budget is spent (0: never) */
  float conv_eps;   /* RGBID_TERM_CONVERGENCE: a level ends after the update whose |x| (6-vector, m and rad) < conv_eps */
} rgbid_align_config;

/* TERMINATION_CRITERIA of the tracker (src/internal.h:112; a zeroed config is ALL_ITERS, the shipped default).
 * CHI_SQUARED restates src/visodo.cpp:1134-1164: from the second iteration of a level on, the robust chi^2 of ALL
 * level-0 residuals at the current pose (computeErrorGridStride + computeChiSquare with the reference scales 5 /
 * 0.0025) gives RMSE = sqrt(chi2) / sqrt(Ndof); if it grew since the previous iteration the last increment is undone
 * and the level ends.  The reference evaluates the test on the level-0 WARPED maps, which pyrFirst only refreshes while
 * level 0 iterates (above it the maps are stale, the RMSE repeats and the test never fires), so here the test runs at
 * level 0 under pyrFirst and at every level under warpFirst.
 * CONVERGENCE is not in the reference: BASELINE config 2's "full GN convergence" schedule (iterate a level until
 * |x| < conv_eps or iterations[level] is reached). */
enum { RGBID_TERM_ALL_ITERS = 0, RGBID_TERM_CHI_SQUARED = 1, RGBID_TERM_CONVERGENCE = 2 };

typedef struct {
  int level, iter;
  double sums27[27];
  float sigma_int, sigma_depthinv, bias_int, bias_depthinv, nu_int, nu_depthinv;
  int irls_iters_int, irls_iters_depthinv;
  double x[6];
  double R[9], t[3];
} rgbid_iter_trace;

int rgbid_aligner_create(rgbid_ctx* ctx, const rgbid_align_config* cfg, rgbid_aligner** out);
int rgbid_aligner_destroy(rgbid_aligner* al);
/* Total Gauss-Newton iterations per pair (sum of cfg.iterations over the active levels). */
int rgbid_aligner_num_iterations(const rgbid_aligner* al);
/* After rgbid_aligner_fetch / _run: iterations actually executed per pair and level, out[batch][RGBID_MAX_LEVELS]
 * (equal to cfg.iterations unless cfg.termination ended a level early). */
int rgbid_aligner_iterations_done(const rgbid_aligner* al, int* out);

/* Load level-0 float maps (inverse depth, intensity) of the keyframe ("ini") of pair `index` and build
 * its pyramid, Sobel gradients and -- in tracker mode -- the bilateral-filtered covariance-only
 * gradients (saveCurrentImagesAsOdoKeyframes, src/visodo.cpp:826-878; keyframe_align.cpp:157-176).
 * from_host != 0: pointers are host memory (uploaded through the context's pinned staging). */
int rgbid_aligner_set_keyframe(rgbid_aligner* al, int index, const float* depthinv, size_t dpitch,
                               const float* intensity, size_t ipitch, int from_host);
/* Same for the current frame ("end" keyframe): pyramid only (prepareImages, src/visodo.cpp:760-773). */
int rgbid_aligner_set_current(rgbid_aligner* al, int index, const float* depthinv, size_t dpitch,
                              const float* intensity, size_t ipitch, int from_host);
/* Ingest raw sensor data (uint16 depth in mm, RGB8) as the current frame of pair `index`:
 * convertDepth2InvDepth + computeIntensity + pyramid fused (prepareImages). */
int rgbid_aligner_set_current_rgbd(rgbid_aligner* al, int index, const uint16_t* depth, size_t dpitch,
                                   const uint8_t* rgb, size_t cpitch, int from_host);
/* Promote the current frame of pair `index` to keyframe (copy + gradients (+ filtered gradients)). */
int rgbid_aligner_current_to_keyframe(rgbid_aligner* al, int index);

/* The solver records one rgbid_iter_trace per pair and iteration on the device.  rgbid_aligner_run switches the recording
 * on exactly when trace_out != NULL; callers of rgbid_aligner_enqueue / rgbid_aligner_fetch who do not read traces can
 * switch it off (it is on after rgbid_aligner_create; the tracker's own aligner runs with it off). */
int rgbid_aligner_set_trace(rgbid_aligner* aligner, int enable);
/* Run the whole coarse-to-fine schedule for all `batch` pairs on the device.
 * R_inout: batch x 9 doubles (row-major rotation _{KF}R^{cur}), t_inout: batch x 3 doubles: initial guess
 * in, estimate out.  cov_out: batch x 36 doubles (may be NULL).  status_out: batch ints (RGBID_OK or
 * RGBID_ERR_NAN).  trace_out (may be NULL): batch x num_iterations entries.  Synchronous. */
int rgbid_aligner_run(rgbid_aligner* al, double* R_inout, double* t_inout, double* cov_out, int* status_out,
                      rgbid_iter_trace* trace_out);
/* Asynchronous variant: enqueue only; results are fetched with rgbid_aligner_fetch (which synchronises). */
int rgbid_aligner_enqueue(rgbid_aligner* al, const double* R_init, const double* t_init);
int rgbid_aligner_fetch(rgbid_aligner* al, double* R_out, double* t_out, double* cov_out, int* status_out,
                        rgbid_iter_trace* trace_out);
/* Frame statistics of the last tracker-mode run: chi_square / chi_test / ndof of the end-of-frame test
 * (src/visodo.cpp:1411-1415), 3 floats per pair. */
int rgbid_aligner_frame_stats(rgbid_aligner* al, float* stats_out);
/* Device-side export of the alignment results for a multi-GPU exchange: writes, on the context's
 * stream, [batch][48] doubles into DEVICE memory `d_out` -- per pair the 6x6 covariance (row-major, the inverse of the
 * last normal-equation matrix), R (9, row-major) and t (3) of _{KF}T^{cur} -- straight from the solver state, so that an
 * NCCL all-gather can follow on a side stream without any host copy.  Pairs whose alignment failed export NaN. */
int rgbid_aligner_export_systems(rgbid_aligner* al, double* d_out);
/* Bytes of device->host traffic rgbid_aligner_fetch / rgbid_tracker_track read back per pair (the solver state). */
size_t rgbid_aligner_state_bytes(void);
/* Measurement hook for bench.py's roofline: launches the fused warp+residual+J^T J kernel of `level` `reps`
 * times back to back on the aligner's current maps / poses / scales (no pose update), bracketed by CUDA
 * events on the context's stream; returns the average launch duration in milliseconds.  Synchronous. */
int rgbid_aligner_time_build(rgbid_aligner* al, int level, int reps, float* ms_per_launch_host);
/* Same for the fused sampling + scale-estimation kernel of `level`. */
int rgbid_aligner_time_scale(rgbid_aligner* al, int level, int reps, float* ms_per_launch_host);
/* Device pointer + pitch of an internal pyramid map, for tests and for callers that fill maps in place.
 * which: 0 W_kf, 1 I_kf, 2 gWx, 3 gWy, 4 gIx, 5 gIy, 6 W_cur, 7 I_cur, 8..11 covariance-only gradients */
int rgbid_aligner_map(rgbid_aligner* al, int which, int level, int index, float** ptr, size_t* pitch);

/* ---------------------------------------------------------------------------------------------- */
/* Tracker: per-frame state machine of VisodoTracker::trackNewFrame (src/visodo.cpp:1967-2247) for  */
/* `batch` independent RGB-D streams: ingest, pyramid, Gauss-Newton, covariance, covisibility,      */
/* keyframe switching, inverse-depth fusion.                                                        */
/* ---------------------------------------------------------------------------------------------- */

typedef struct {
  rgbid_align_config align;  /* mode is forced to RGBID_MODE_TRACKER */
  int motion_model;          /* RGBID_NO_MM | RGBID_CONSTANT_VELOCITY (src/visodo.cpp:1016-1032) */
  float visratio_odo;        /* 0.9, src/internal.h:110 */
  float visratio_integr;     /* 0.7, src/internal.h:111 */
  int max_odo_kf_count;      /* 9999999 */
  int max_integr_kf_count;   /* 9999999 */
  int image_filtering;       /* RGBID_NO_FILTERS | RGBID_FILTER_GRADS */
  float delta_t;             /* 0.03333 s in eval mode (src/visodo.cpp:1932) */
} rgbid_tracker_config;

typedef struct {
  double R[9], t[3];          /* global pose (camera-to-world) after this frame */
  double dR[9], dt[3];        /* pose relative to the odometry keyframe, _{KF}T^{cur} */
  double cov[36];             /* covariance of the keyframe-relative estimate */
  float visibility_odo, visibility_integr;
  float chi_square, chi_test, ndof;
  int status;                 /* RGBID_OK or RGBID_ERR_NAN (lost) */
  int new_odo_keyframe, new_integr_keyframe;
  int frame_index;            /* per stream: the reference's global_time_ when the frame was tracked */
  /* sequential odometry constraint (frame_index - 1 -> frame_index) the reference pushes to the keyframe manager as
   * PoseConstraint::SEQ_ODO (src/visodo.cpp:2126-2156): relative transform and its propagated 6x6 covariance; the
   * dummy constraint (identity, 100 I) when tracking was lost (:2068-2071) */
  double seq_R[9], seq_t[3], seq_cov[36];
  /* 1: the alignment failed while the stream was already lost (src/visodo.cpp:2099-2116): both keyframes were re-saved
   * from this frame and nothing else happened -- no constraint, no keyframe hand-off, frame_index (the reference's
   * global_time_) does not advance.  The caller must not push the dummy constraint for such a frame. */
  int lost_again;
} rgbid_frame_result;

/* ---- custom-calibration ingest and colour fusion / previews (f4) ------------------------------------
 * One entry per reference bridge function; same argument meaning. */
typedef struct rgbid_intr { float fx, fy, cx, cy, k1, k2, k3, k4, k5; } rgbid_intr;          /* Intr, src/internal.h:119-140 */
typedef struct rgbid_depth_dist {                                                             /* DepthDist, :142-161 */
  float c1, c0;
  float q0[9], q1[9];
  int xshift, yshift;
} rgbid_depth_dist;
/* undistortIntensity (src/internal.h:437, src/cuda/undistortion.cu:212-257) */
int rgbid_undistort_intensity(rgbid_ctx* ctx, const float* src, size_t spitch, float* dst, size_t dpitch, int rows, int cols,
                              const rgbid_intr* intr);
/* undistortDepthInv (src/internal.h:440, undistortion.cu:260-310); the reference's src_corr scratch map is not needed */
int rgbid_undistort_depthinv(rgbid_ctx* ctx, const float* src, size_t spitch, float* dst, size_t dpitch, int rows, int cols,
                             const rgbid_intr* intr_depth, const rgbid_depth_dist* dp);
/* registerDepthinv (src/internal.h:354, src/cuda/warping_registration.cu:720-800).  The 3 rows x 3 cols canvas
 * (`intermediate` / `intermediate_as_int`) lives in the context's scratch.  dRc_proj = Kd dRc Kc^-1, t_dc_proj = Kd t_dc,
 * cRd_proj = dRc_proj^-1 (row-major 3x3 / 3), as built in src/visodo.cpp:789-807. */
int rgbid_register_depthinv(rgbid_ctx* ctx, const float* src, size_t spitch, float* dst, size_t dpitch, int rows, int cols,
                            const float* dRc_proj, const float* t_dc_proj, const float* cRd_proj);
/* prepareImagesCustomCalibration (src/visodo.cpp:775-823) inside the tracker: when set, every frame is ingested as
 * intensity -> undistortIntensity(rgb), inverse depth -> undistortDepthInv(depth, dist) -> registerDepthinv with
 * dRc_proj = Kd dRc Kc^-1, t_dc_proj = Kd t_dc (float, :789-807) instead of the plain conversion.  cal == NULL switches back.
 * (config_data/calibration_custom.ini: [RGB_CALIBRATION], [DEPTH_CALIBRATION] custom_registration=1, [STEREO_DEPTH2RGB]) */
typedef struct rgbid_custom_calibration {
  rgbid_intr rgb, depth;
  rgbid_depth_dist dist;
  float dRc[9], t_dc[3];
} rgbid_custom_calibration;
int rgbid_tracker_set_custom_calibration(rgbid_tracker* trk, const rgbid_custom_calibration* cal);
/* integrateWarpedRGB (src/internal.h, warping_registration.cu:672-712, 1103-1129): depth_dst, colors_dst (rows x cols x 3
 * uint8, pitch colors_pitch) and weight_dst are updated in place; all float maps share `pitch`. */
int rgbid_integrate_warped_rgb(rgbid_ctx* ctx, const float* depth_warped, const float* r_warped, const float* g_warped,
                               const float* b_warped, const float* weight_warped, float* depth_dst, uint8_t* colors_dst,
                               size_t colors_pitch, float* weight_dst, size_t pitch, int rows, int cols);
/* generateImage / generateImageRGB (src/internal.h:416-421, src/cuda/image_generator.cu): vmap / nmap are 3 rows x cols
 * (x, y, z planes); rgb may be NULL (grey shading); light: position of the single light source; out: rows x cols x 3. */
int rgbid_generate_image(rgbid_ctx* ctx, const float* vmap, const float* nmap, size_t map_pitch, const uint8_t* rgb,
                         size_t rgb_pitch, const float* light_pos, uint8_t* out, size_t out_pitch, int rows, int cols);

/* ---- keyframe hand-off to the back end (resetIntegrationKeyframe, src/visodo.cpp:1577-1672) ------------------------
 * When a stream switches its integration keyframe, the OUTGOING keyframe is handed to the sink: its creation index and
 * global pose, the SEQ_KF constraint from the previous keyframe with the covariance propagated through the
 * odometry-keyframe chain, and host copies of its overlap mask, colours, fused inverse depth and normals (what the
 * reference downloads into a Keyframe, :1639-1642).  The pointers are valid only during the callback, which runs inside
 * rgbid_tracker_track on the calling thread. */
enum { RGBID_SEQ_ODO = 0, RGBID_SEQ_KF = 1 };
typedef struct rgbid_keyframe_handoff {
  int stream;                  /* index in the batch */
  int kf_index;                /* frame at which this keyframe was created (last_integrKF_index_) */
  int frame_index;             /* frame that replaces it (global_time_) */
  int rows, cols;
  float fx, fy, cx, cy;
  double R[9], t[3];           /* global pose of the keyframe */
  double rel_R[9], rel_t[3];   /* SEQ_KF constraint kf_index -> frame_index ... */
  double rel_cov[36];          /* ... and its covariance */
  const uint8_t* overlap_mask; size_t overlap_mask_pitch;   /* rows x cols */
  const uint8_t* colors;       /* rows x cols x 3, dense */
  const float* depthinv; size_t depthinv_pitch;             /* rows x cols, fused */
  const float* normals; size_t normals_pitch;               /* 3 rows x cols (x, y, z planes) */
} rgbid_keyframe_handoff;
typedef void (*rgbid_keyframe_sink)(void* user, const rgbid_keyframe_handoff* keyframe);
/* cb == NULL removes the sink (then nothing is downloaded at a keyframe switch) */
int rgbid_tracker_set_keyframe_sink(rgbid_tracker* trk, rgbid_keyframe_sink cb, void* user);

int rgbid_tracker_create(rgbid_ctx* ctx, const rgbid_tracker_config* cfg, rgbid_tracker** out);
int rgbid_tracker_destroy(rgbid_tracker* trk);
int rgbid_tracker_reset(rgbid_tracker* trk);
/* Track one frame per stream.  depth: batch x rows x cols uint16 (dense), rgb: batch x rows x cols x 3
 * uint8 (dense).  from_host != 0: host pointers (pinned or pageable; copied inside the call).
 * results_host: batch entries.  Synchronous (the reference's trackNewFrame is). */
int rgbid_tracker_track(rgbid_tracker* trk, const uint16_t* depth, const uint8_t* rgb, int from_host,
                        rgbid_frame_result* results_host);
/* Optional: start the host->device upload of the NEXT frame (same layout as rgbid_tracker_track with from_host != 0)
 * on a copy stream and return immediately.  A following rgbid_tracker_track(trk, depth, rgb, 1, ...) with the same two
 * pointers uses the uploaded copy instead of copying again, so the upload overlaps the tracking of the current frame
 * (the reference uploads and tracks strictly one after the other, tools/RGBID_SLAMapp.cpp:163-214).  One frame may be
 * in flight; the host buffers must stay valid and unchanged until that track call returns.  A prefetched frame is used by
 * the current or the next track call only; after that it is dropped and the frame is copied inside the call as usual. */
int rgbid_tracker_prefetch(rgbid_tracker* trk, const uint16_t* depth, const uint8_t* rgb);
/* Same with pitched DEVICE buffers (what VisodoTracker::depth_ / rgb24_ are: DeviceArray2D, include/visodo.h:118-119):
 * row pitch and per-stream stride in bytes (stride is ignored when batch == 1).  The pose read-backs are complete on
 * return, but the keyframe colour copy and the depth fusion of this frame may still be reading the two buffers on the
 * context's stream: call rgbid_ctx_sync (VisodoTracker::trackNewFrame does, like the reference's device::sync()) or
 * order the next write after that stream before refilling them. */
int rgbid_tracker_track_device(rgbid_tracker* trk, const uint16_t* depth, size_t depth_pitch, size_t depth_stride,
                               const uint8_t* rgb, size_t rgb_pitch, size_t rgb_stride, rgbid_frame_result* results_host);
/* Integration-keyframe maps of stream `index` (device pointers; pitch in bytes):
 * which: 0 fused inverse depth, 1 fusion weight, 2 raw inverse depth, 3 vertex map (3*rows), 4 normal map
 * (3*rows) */
int rgbid_tracker_keyframe_map(rgbid_tracker* trk, int which, int index, float** ptr, size_t* pitch);
int rgbid_tracker_overlap_mask(rgbid_tracker* trk, int index, uint8_t** ptr, size_t* pitch);
rgbid_aligner* rgbid_tracker_aligner(rgbid_tracker* trk);

#ifdef __cplusplus
}
#endif
#endif /* RGBID_B200_H_ */
