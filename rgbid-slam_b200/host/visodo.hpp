// visodo.hpp -- RGBID_SLAM::VisodoTracker and RGBID_SLAM::KeyframeAlign with the reference's entry points
// (include/visodo.h:47-135, include/keyframe_align.h:40-54) on top of the C ABI.  The per-frame state machine,
// the Gauss-Newton loop and all kernels live in librgbid_b200.so (csrc/tracker.cu, aligner.cu); these classes own
// the configuration surface the application touches: constructor arguments, the [VISODO] INI keys
// (src/visodo.cpp:321-435), the calibration file ([CALIBRATION] | [RGB_CALIBRATION], src/visodo.cpp:99-180), the
// public device input buffers rgb24_ / depth_, and the pose accessors.
//
// Eigen is not available in this image, so poses are returned as plain row-major structs (Affine3) instead of
// Eigen::Affine3f / Affine3d; everything else keeps the reference's names.  Header-only.
#pragma once
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rgbid_b200/internal.hpp"
#include "settings.hpp"
#include "keyframe.hpp"

namespace RGBID_SLAM {

typedef DeviceArray2D<PixelRGB> View;  // PixelRGB: keyframe.hpp
typedef DeviceArray2D<unsigned short> DepthMap;

/** Rigid transform, row-major: X_world = R X_cam + t */
struct Affine3 {
  double R[9];
  double t[3];
  Affine3() { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0; t[0] = t[1] = t[2] = 0.0; }
};

class VisodoTracker {
 public:
  enum { LEVELS = 3 };  // include/visodo.h:52; `levels` below generalises it (4 for the 4-level benchmark)

  VisodoTracker(int optim_dim = 6, int Mestimator = device::DEFAULT_MESTIMATOR,
                int motion_model = device::DEFAULT_MOTION_MODEL, int sigma_estimator = device::SIGMA_PDF,
                int weighting = device::DEFAULT_WEIGHTING, int warping = device::WARP_FIRST,  // include/visodo.h:59
                int max_odoKF_count = device::DEFAULT_ODO_KF_COUNT, int finest_level = 0,
                int termination = device::DEFAULT_TERMINATION, float visratio_odo = device::DEFAULT_VISRATIO_ODO,
                int image_filtering = device::DEFAULT_IMAGE_FILTERING, float visratio_integr = device::DEFAULT_VISRATIO_INTEGR,
                int max_integrKF_count = device::DEFAULT_INTEGR_KF_COUNT, int Nsamples = 10000, int rows = 480,
                int cols = 640, int levels = LEVELS)
      : rows_(rows), cols_(cols), levels_(levels), optim_dim_(optim_dim), Mestimator_(Mestimator),
        motion_model_(motion_model), sigma_estimator_(sigma_estimator), weighting_(weighting), warping_(warping),
        max_odoKF_count_(max_odoKF_count), finest_level_(finest_level), termination_(termination),
        visibility_ratio_odo_threshold_(visratio_odo), image_filtering_(image_filtering),
        visibility_ratio_integr_threshold_(visratio_integr), max_integrKF_count_(max_integrKF_count),
        Nsamples_(Nsamples), trk_(nullptr), lost_(false), global_time_(0)
  {
    if (optim_dim != 6) throw std::invalid_argument("only the 6-DoF optimisation of the reference is implemented");
    setRGBIntrinsics(device::FOCAL_LENGTH, device::FOCAL_LENGTH, device::CENTER_X, device::CENTER_Y);
    factor_depth_ = 1.f;
    compute_deltat_flag_ = false;
    real_time_flag_ = false;
    const int iters[] = {10, 5, 3, 0, 0, 0, 0, 0};  // src/visodo.cpp:65
    for (int i = 0; i < RGBID_MAX_LEVELS; ++i) visodo_iterations_[i] = i < levels_ ? iters[i] : 0;
    rgb24_.create(rows_, cols_);
    depth_.create(rows_, cols_);
  }

  ~VisodoTracker()
  {
    stop();
    if (trk_) rgbid_tracker_destroy(trk_);
    if (ctx_) rgbid_ctx_destroy(ctx_);
  }

  /** Starts the tracking thread and returns once it is waiting for frames (src/visodo.cpp:437-445).  The grabber
      then hands frames over exactly like tools/RGBID_SLAMapp.cpp:144-214: try-lock `mutex_`, upload into `depth_` /
      `rgb24_`, `new_frame_cond_.notify_one()`.  std:: primitives stand in for the reference's boost:: ones. */
  void start()
  {
    std::unique_lock<std::mutex> lock(created_aux_mutex_);
    created_ = false;
    visodo_thread_.reset(new std::thread([this] { (*this)(); }));
    created_cond_.wait(lock, [this] { return created_; });
  }

  /** Body of the tracking thread (src/visodo.cpp:2249-2267): wait for a frame, track it, until exit_ is set. */
  bool operator()()
  {
    std::unique_lock<std::mutex> lock(mutex_);
    exit_ = false;
    {
      std::lock_guard<std::mutex> aux(created_aux_mutex_);
      created_ = true;
    }
    created_cond_.notify_one();
    while (!exit_) {
      new_frame_cond_.wait(lock);
      if (exit_) break;
      trackNewFrame();
      ++frames_tracked_by_thread_;
    }
    return true;
  }

  /** Ends the tracking thread (the reference sets exit_ from the application and never joins). */
  void stop()
  {
    if (!visodo_thread_) return;
    {
      std::lock_guard<std::mutex> lock(mutex_);
      exit_ = true;
    }
    new_frame_cond_.notify_all();
    if (visodo_thread_->joinable()) visodo_thread_->join();
    visodo_thread_.reset();
  }

  // hand-shake with the grabber thread (include/visodo.h:121-127)
  std::mutex mutex_;
  std::condition_variable new_frame_cond_;
  std::mutex created_aux_mutex_;
  std::condition_variable created_cond_;
  bool exit_ = false;
  int frames_tracked_by_thread_ = 0;

  /** [VISODO] keys of the reference (src/visodo.cpp:321-435) */
  void loadSettings(Settings& settings)
  {
    Section s;
    if (!settings.getSection("VISODO", s)) return;
    Entry e;
    if (s.getEntry("M_ESTIMATOR", e)) {
      const std::string v = e.getValue();
      if (v == "Student") Mestimator_ = device::STUDENT;
      if (v == "LeastSquares") Mestimator_ = device::LSQ;
      if (v == "Tukey") Mestimator_ = device::TUKEY;
      if (v == "Huber") Mestimator_ = device::HUBER;
    }
    if (s.getEntry("MOTION_MODEL", e)) {
      if (e.getValue() == "none") motion_model_ = device::NO_MM;
      if (e.getValue() == "constVelocity") motion_model_ = device::CONSTANT_VELOCITY;
    }
    if (s.getEntry("WARP_ORDER", e)) {
      if (e.getValue() == "warpFirst") warping_ = device::WARP_FIRST;
      if (e.getValue() == "pyrFirst") warping_ = device::PYR_FIRST;
    }
    if (s.getEntry("IMAGE_FILTERING", e)) {
      if (e.getValue() == "none") image_filtering_ = device::NO_FILTERS;
      if (e.getValue() == "gradients") image_filtering_ = device::FILTER_GRADS;
    }
    if (s.getEntry("SIGMA_ESTIMATOR", e)) {
      if (e.getValue() == "sigmaMAD") sigma_estimator_ = device::SIGMA_MAD;
      if (e.getValue() == "sigmaML") sigma_estimator_ = device::SIGMA_PDF;
      if (e.getValue() == "sigmaConst") sigma_estimator_ = device::SIGMA_CONS;
    }
    if (s.getEntry("INTEGRATION_VISRATIO_THRESHOLD", e)) visibility_ratio_integr_threshold_ = (float)atof(e.getValue().c_str());
    if (s.getEntry("ODOMETRY_VISRATIO_THRESHOLD", e)) visibility_ratio_odo_threshold_ = (float)atof(e.getValue().c_str());
    if (s.getEntry("FINEST_PYR_LEVEL", e)) finest_level_ = (int)atof(e.getValue().c_str());
  }

  /** [CALIBRATION] | [RGB_CALIBRATION]: fx fy cx cy kd factor_depth (src/visodo.cpp:99-180) */
  void loadCalibration(const std::string& calib_file)
  {
    std::ifstream f(calib_file.c_str());
    if (!f.is_open()) { std::cout << "Could not open configuration file " << calib_file << std::endl; return; }
    Settings settings(f);
    Section c;
    if (!(settings.getSection("CALIBRATION", c) || settings.getSection("RGB_CALIBRATION", c))) return;
    Entry e;
    if (c.getEntry("fx", e)) fx_ = (float)atof(e.getValue().c_str());
    if (c.getEntry("fy", e)) fy_ = (float)atof(e.getValue().c_str());
    if (c.getEntry("cx", e)) cx_ = (float)atof(e.getValue().c_str());
    if (c.getEntry("cy", e)) cy_ = (float)atof(e.getValue().c_str());
    if (c.getEntry("factor_depth", e)) factor_depth_ = (float)atof(e.getValue().c_str());
    if (c.getEntry("kd", e)) {  // RGB distortion k1..k5 (src/visodo.cpp:160-168)
      std::stringstream ss(e.getValue());
      ss >> custom_.rgb.k1 >> custom_.rgb.k2 >> custom_.rgb.k3 >> custom_.rgb.k4 >> custom_.rgb.k5;
    }
    // [DEPTH_CALIBRATION] + [STEREO_DEPTH2RGB]: the custom registration path (src/visodo.cpp:183-318)
    Section d;
    if (settings.getSection("DEPTH_CALIBRATION", d)) {
      if (d.getEntry("custom_registration", e)) { std::stringstream ss(e.getValue()); ss >> custom_registration_; }
      if (d.getEntry("fx", e)) custom_.depth.fx = (float)atof(e.getValue().c_str());
      if (d.getEntry("fy", e)) custom_.depth.fy = (float)atof(e.getValue().c_str());
      if (d.getEntry("cx", e)) custom_.depth.cx = (float)atof(e.getValue().c_str());
      if (d.getEntry("cy", e)) custom_.depth.cy = (float)atof(e.getValue().c_str());
      if (d.getEntry("kd", e)) {
        std::stringstream ss(e.getValue());
        ss >> custom_.depth.k1 >> custom_.depth.k2 >> custom_.depth.k3 >> custom_.depth.k4 >> custom_.depth.k5;
      }
      if (d.getEntry("c0", e)) { std::stringstream ss(e.getValue()); ss >> custom_.dist.c0; }
      if (d.getEntry("c1", e)) { std::stringstream ss(e.getValue()); ss >> custom_.dist.c1; }
      if (d.getEntry("q0", e)) { std::stringstream ss(e.getValue()); for (int i = 0; i < 9; ++i) ss >> custom_.dist.q0[i]; }
      if (d.getEntry("q1", e)) { std::stringstream ss(e.getValue()); for (int i = 0; i < 9; ++i) ss >> custom_.dist.q1[i]; }
    }
    if (settings.getSection("STEREO_DEPTH2RGB", d)) {
      if (d.getEntry("dRc", e)) { std::stringstream ss(e.getValue()); for (int i = 0; i < 9; ++i) ss >> custom_.dRc[i]; }
      if (d.getEntry("t_dc", e)) { std::stringstream ss(e.getValue()); for (int i = 0; i < 3; ++i) ss >> custom_.t_dc[i]; }
    }
    custom_.rgb.fx = fx_; custom_.rgb.fy = fy_; custom_.rgb.cx = cx_; custom_.rgb.cy = cy_;
    if (trk_) apply_custom_calibration();
  }

  void setRGBIntrinsics(float fx, float fy, float cx = -1, float cy = -1, float = 0.f, float = 0.f, float = 0.f,
                        float = 0.f, float = 0.f)
  {
    fx_ = fx; fy_ = fy;
    cx_ = (cx == -1) ? cols_ / 2 - 0.5f : cx;
    cy_ = (cy == -1) ? rows_ / 2 - 0.5f : cy;
  }

  int cols() { return cols_; }
  int rows() { return rows_; }
  bool visOdoIsLost() { return lost_; }
  size_t getNumberOfPoses() const { return poses_.size(); }

  /** Process the frame held in rgb24_ / depth_ (device buffers, as filled by the application's grabber thread,
      tools/RGBID_SLAMapp.cpp:144-214).  Returns false for the first frame and when tracking is lost, like the
      reference. */
  bool trackNewFrame()
  {
    ensure_created();
    const auto t_begin = std::chrono::steady_clock::now();
    computeInterframeTime();
    rgbid_frame_result r;
    int rc = rgbid_tracker_track_device(trk_, depth_.ptr(), depth_.step(), 0, (const uint8_t*)rgb24_.ptr(), rgb24_.step(), 0, &r);
    if (rc != RGBID_OK) throw std::runtime_error(std::string("rgbid_tracker_track: ") + rgbid_status_string(rc));
    // device::sync() at the end of trackNewFrame (src/visodo.cpp:2230): the keyframe colour copy and the fusion queued
    // behind the last read-back still read rgb24_ / depth_, which the grabber refills (legacy stream) as soon as we return
    rc = rgbid_ctx_sync(ctx_);
    if (rc != RGBID_OK) throw std::runtime_error(std::string("rgbid_ctx_sync: ") + rgbid_status_string(rc));
    last_ = r;
    lost_ = (r.status != RGBID_OK);
    const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    kf_time_accum_ += 1e-3f * ms;
    if (r.lost_again) {
      // failed again while lost (src/visodo.cpp:2111-2116): the keyframes were re-saved from this frame; no pose, no
      // constraint, global_time_ stands still
      vis_odo_times_.push_back(ms);
      return false;
    }
    if (global_time_ > 0)  // SEQ_ODO constraint (or the dummy one when lost), src/visodo.cpp:2068-2071, 2147-2154
      keyframe_buffers_.constraints_.push_back(PoseConstraint(global_time_ - 1, global_time_, PoseConstraint::SEQ_ODO, r.seq_R,
                                                              r.seq_t, 1.f, r.seq_cov));
    Affine3 p;
    std::memcpy(p.R, r.R, sizeof(p.R));
    std::memcpy(p.t, r.t, sizeof(p.t));
    poses_.push_back(p);
    setSharedCameraPose(p);
    // wall time of the frame in milliseconds (pcl::ScopeTime t1 of the reference, src/visodo.cpp:2231; 0 for frame 0, :542)
    vis_odo_times_.push_back(global_time_ == 0 ? 0.f : ms);
    const bool first = (global_time_ == 0);
    ++global_time_;
    return !first && !lost_;
  }

  /** Camera-to-world pose of frame `time` (-1: latest) */
  Affine3 getCameraPose(int time = -1) const
  {
    if (poses_.empty()) return Affine3();
    if (time > (int)poses_.size() || time < 0) time = (int)poses_.size() - 1;
    return poses_[time];
  }

  /** Milliseconds spent on frame `time` (-1: latest), src/visodo.cpp:1853-1860 */
  float getVisOdoTime(int time = -1) const
  {
    if (vis_odo_times_.empty()) return 0.f;
    if (time > (int)vis_odo_times_.size() || time < 0) time = (int)vis_odo_times_.size() - 1;
    return vis_odo_times_[time];
  }

  /** RGB time stamp of frame `time` relative to the first one (0 unless compute_deltat_flag_), src/visodo.cpp:1863-1870 */
  int64_t getTimestamp(int time = -1) const
  {
    if (timestamps_.empty()) return 0;
    if (time > (int)timestamps_.size() || time < 0) time = (int)timestamps_.size() - 1;
    return timestamps_[time];
  }

  /** Latest pose for a consumer thread (the reference's visualisation), src/visodo.cpp:505-512, 1873-1880 */
  void setSharedCameraPose(const Affine3& pose)
  {
    std::lock_guard<std::mutex> lock(mutex_shared_camera_pose_);
    shared_camera_pose_ = pose;
    camera_pose_has_changed_ = true;
  }
  Affine3 getSharedCameraPose()
  {
    std::lock_guard<std::mutex> lock(mutex_shared_camera_pose_);
    camera_pose_has_changed_ = false;
    return shared_camera_pose_;
  }
  bool camera_pose_has_changed_ = false;

  /** Tracking milliseconds accumulated between consecutive integration keyframes (src/visodo.cpp:1656) */
  std::vector<float> kf_times_;

  const rgbid_frame_result& lastResult() const { return last_; }

  /** What the reference pushes to keyframe_manager_ptr_ (buffer_keyframes_, constraints_): every outgoing integration
      keyframe with its SEQ_KF constraint (resetIntegrationKeyframe, src/visodo.cpp:1577-1672) and one SEQ_ODO constraint
      per frame.  The consumer (KeyframeManager) is out of scope; it would drain these. */
  KeyframeBuffers keyframe_buffers_;
  rgbid_tracker* handle() { ensure_created(); return trk_; }

  void reset()
  {
    poses_.clear();
    vis_odo_times_.clear();
    timestamps_.clear();
    kf_times_.clear();
    kf_time_accum_ = 0.f;
    global_time_ = 0;
    lost_ = false;
    if (trk_) rgbid_tracker_reset(trk_);
  }

  // public inputs of the reference (include/visodo.h:118-119)
  View rgb24_;
  DepthMap depth_;
  uint64_t timestamp_rgb_curr_ = 0, timestamp_depth_curr_ = 0, timestamp_ini_ = 0;
  bool compute_deltat_flag_, real_time_flag_;
  int visodo_iterations_[RGBID_MAX_LEVELS];

 private:
  /** computeInterframeTime, src/visodo.cpp:1929-1964: records the zeroed RGB time stamp of the frame.  The interval
      itself stays the evaluation-mode constant (1 / 30 s) that the tracker was created with: compute_deltat_flag_ is
      only set by the live-camera grabber (tools/RGBID_SLAMapp.cpp), which is out of scope. */
  void computeInterframeTime()
  {
    if (!compute_deltat_flag_) { timestamp_rgb_curr_ = 0; timestamp_depth_curr_ = 0; return; }
    if (global_time_ == 0) timestamp_ini_ = std::min(timestamp_rgb_curr_, timestamp_depth_curr_);
    timestamps_.push_back((int64_t)(timestamp_rgb_curr_ - timestamp_ini_));
  }

  void apply_custom_calibration()
  {
    int rc = rgbid_tracker_set_custom_calibration(trk_, custom_registration_ ? &custom_ : nullptr);
    if (rc != RGBID_OK) throw std::runtime_error(std::string("rgbid_tracker_set_custom_calibration: ") + rgbid_status_string(rc));
  }

  static void keyframe_sink(void* user, const rgbid_keyframe_handoff* k)
  {
    VisodoTracker* self = (VisodoTracker*)user;
    self->kf_times_.push_back(1000.f * self->kf_time_accum_);  // src/visodo.cpp:1656
    self->kf_time_accum_ = 0.f;
    self->keyframe_buffers_.buffer_keyframes_.push_back(KeyframePtr(new Keyframe(*k)));
    self->keyframe_buffers_.constraints_.push_back(PoseConstraint(k->kf_index, k->frame_index, PoseConstraint::SEQ_KF, k->rel_R,
                                                                  k->rel_t, 1.f, k->rel_cov));
  }

  void ensure_created()
  {
    if (trk_) return;
    rgbid_tracker_config c;
    std::memset(&c, 0, sizeof(c));
    c.align.rows = rows_; c.align.cols = cols_; c.align.levels = levels_; c.align.finest_level = finest_level_;
    for (int i = 0; i < RGBID_MAX_LEVELS; ++i) c.align.iterations[i] = visodo_iterations_[i];
    if (real_time_flag_) c.align.iterations[0] = 5;  // src/visodo.cpp:961-964
    c.align.batch = 1; c.align.mode = RGBID_MODE_TRACKER;
    c.align.mestimator = Mestimator_; c.align.weighting = weighting_;
    // SIGMA_MAD parses but has no implementation in the reference: it falls through to the constants
    c.align.sigma_estimator = (sigma_estimator_ == device::SIGMA_PDF) ? RGBID_SIGMA_PDF : RGBID_SIGMA_CONS;
    c.align.nsamples = Nsamples_;
    c.align.fx = fx_; c.align.fy = fy_; c.align.cx = cx_; c.align.cy = cy_;
    c.align.factor_depth = factor_depth_;
    c.align.warp_first = (warping_ == device::WARP_FIRST) ? 1 : 0;  // src/visodo.cpp:1078
    // TERMINATION_CRITERIA, src/internal.h:112 / src/visodo.cpp:1134-1164
    c.align.termination = (termination_ == device::CHI_SQUARED) ? RGBID_TERM_CHI_SQUARED : RGBID_TERM_ALL_ITERS;
    c.motion_model = motion_model_;
    c.visratio_odo = visibility_ratio_odo_threshold_; c.visratio_integr = visibility_ratio_integr_threshold_;
    c.max_odo_kf_count = max_odoKF_count_; c.max_integr_kf_count = max_integrKF_count_;
    c.image_filtering = image_filtering_;
    c.delta_t = 0.03333f;  // computeInterframeTime in evaluation mode, src/visodo.cpp:1932
    // The tracker owns its context: a thread-local one would be destroyed with the thread that first tracked a frame
    // (the tracking thread of start()), while the tracker itself is destroyed by the thread that owns the object.
    int rc = ctx_ ? RGBID_OK : rgbid_ctx_create(&ctx_, device::dev_id, nullptr);
    if (rc != RGBID_OK) throw std::runtime_error(std::string("rgbid_ctx_create: ") + rgbid_status_string(rc));
    rc = rgbid_tracker_create(ctx_, &c, &trk_);
    if (rc != RGBID_OK) throw std::runtime_error(std::string("rgbid_tracker_create: ") + rgbid_status_string(rc));
    rgbid_tracker_set_keyframe_sink(trk_, &VisodoTracker::keyframe_sink, this);
    apply_custom_calibration();
  }

  int custom_registration_ = 0;  // src/visodo.cpp:63
  rgbid_custom_calibration custom_ = default_custom_calibration();
  static rgbid_custom_calibration default_custom_calibration()
  {
    rgbid_custom_calibration c;
    std::memset(&c, 0, sizeof(c));
    c.dist.c1 = 1.f; c.dist.q1[0] = 0.f; c.dist.xshift = 4; c.dist.yshift = 4;  // DepthDist defaults, src/internal.h:150-155
    c.dRc[0] = c.dRc[4] = c.dRc[8] = 1.f;                                       // dRc_ = I, t_dc_ = 0 (src/visodo.cpp:59-60)
    return c;
  }
  int rows_, cols_, levels_, optim_dim_, Mestimator_, motion_model_, sigma_estimator_, weighting_, warping_;
  int max_odoKF_count_, finest_level_, termination_;
  float visibility_ratio_odo_threshold_;
  int image_filtering_;
  float visibility_ratio_integr_threshold_;
  int max_integrKF_count_, Nsamples_;
  float fx_, fy_, cx_, cy_, factor_depth_;
  rgbid_tracker* trk_;
  rgbid_ctx* ctx_ = nullptr;
  bool lost_;
  bool created_ = false;
  std::unique_ptr<std::thread> visodo_thread_;
  int global_time_;
  rgbid_frame_result last_;
  std::vector<Affine3> poses_;
  std::vector<float> vis_odo_times_;
  std::vector<int64_t> timestamps_;
  float kf_time_accum_ = 0.f;
  std::mutex mutex_shared_camera_pose_;
  Affine3 shared_camera_pose_;
};

/** The three fields of the reference's Keyframe that KeyframeAlign reads (src/keyframe_align.cpp:119-129):
    K_, depthinv_image_ (CV_32F) and grey_image_ (CV_8U), as plain host pointers (OpenCV is not available). */
struct KeyframeView {
  int rows, cols;
  double K[9];              // row-major 3x3
  const float* depthinv;    // rows x cols, dense
  const unsigned char* grey;  // rows x cols, dense
};

class KeyframeAlign {
 public:
  enum { LEVELS = 4 };  // include/keyframe_align.h:50

  KeyframeAlign() : al_(nullptr), ctx_(nullptr), rows_(0), cols_(0) {}
  ~KeyframeAlign()
  {
    if (al_) rgbid_aligner_destroy(al_);
    if (ctx_) rgbid_ctx_destroy(ctx_);
  }

  /** alignKeyframes(kf_ini, kf_end, rotation_ini2end, translation_ini2end, covariance_ini2end)
      (include/keyframe_align.h:52-54; src/keyframe_align.cpp:115-357).  R (row-major 3x3) and t hold the initial
      guess on entry (from the loop closer's RANSAC) and the refined pose on return; cov is 6x6 row-major. */
  bool alignKeyframes(const KeyframeView& kf_ini, const KeyframeView& kf_end, double* R, double* t, double* cov)
  {
    if (!al_ || kf_ini.rows != rows_ || kf_ini.cols != cols_ || kf_ini.K[0] != K_[0] || kf_ini.K[2] != K_[2]) create(kf_ini);
    grey_f_.resize((size_t)rows_ * cols_);
    const size_t pitch = (size_t)cols_ * sizeof(float);
    for (size_t i = 0; i < grey_f_.size(); ++i) grey_f_[i] = (float)kf_ini.grey[i];  // cv::Mat::convertTo(CV_32F)
    check(rgbid_aligner_set_keyframe(al_, 0, kf_ini.depthinv, pitch, grey_f_.data(), pitch, 1));
    check(rgbid_ctx_sync(ctx_));
    for (size_t i = 0; i < grey_f_.size(); ++i) grey_f_[i] = (float)kf_end.grey[i];
    check(rgbid_aligner_set_current(al_, 0, kf_end.depthinv, pitch, grey_f_.data(), pitch, 1));
    int status = 0;
    check(rgbid_aligner_run(al_, R, t, cov, &status, nullptr));
    return true;  // the reference returns true unconditionally (src/keyframe_align.cpp:356)
  }

 private:
  static void check(int rc)
  {
    if (rc != RGBID_OK) throw std::runtime_error(std::string("KeyframeAlign: ") + rgbid_status_string(rc));
  }
  void create(const KeyframeView& kf)
  {
    if (al_) { rgbid_aligner_destroy(al_); al_ = nullptr; }
    rows_ = kf.rows; cols_ = kf.cols;
    std::memcpy(K_, kf.K, sizeof(K_));
    rgbid_align_config c;
    std::memset(&c, 0, sizeof(c));
    c.rows = rows_; c.cols = cols_; c.levels = LEVELS; c.finest_level = 0;
    const int iters[] = {5, 5, 3, 0};  // src/keyframe_align.cpp:44
    for (int i = 0; i < LEVELS; ++i) c.iterations[i] = iters[i];
    c.batch = 1; c.mode = RGBID_MODE_ALIGN; c.mestimator = RGBID_STUDENT; c.weighting = RGBID_INDEPENDENT;
    c.nsamples = 19200;  // src/keyframe_align.cpp:247-248
    c.fx = (float)kf.K[0]; c.fy = (float)kf.K[4]; c.cx = (float)kf.K[2]; c.cy = (float)kf.K[5];
    c.factor_depth = 1.f;
    if (!ctx_) check(rgbid_ctx_create(&ctx_, device::dev_id, nullptr));  // owned: see VisodoTracker::ensure_created
    check(rgbid_aligner_create(ctx_, &c, &al_));
  }
  rgbid_aligner* al_;
  rgbid_ctx* ctx_;
  int rows_, cols_;
  double K_[9];
  std::vector<float> grey_f_;
};

}  // namespace RGBID_SLAM
