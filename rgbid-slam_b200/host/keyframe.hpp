// keyframe.hpp -- what the tracker hands to the back end: the tracking-side subset of the reference's
// Keyframe (include/keyframe.h:44-75: colors_, depthinv_, normals_, overlap_mask_, K_, id_, cols_, rows_, pose and
// relative pose) and PoseConstraint (include/types.h, src/visodo.cpp:1651-1653, 2147-2149), filled from the C ABI's
// rgbid_keyframe_handoff / rgbid_frame_result.  Eigen is not available here, so matrices are plain row-major arrays.
// The back end itself (KeyframeManager: segmentation, BoW, loop closing, pose graph) is out of scope.
#pragma once
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/rgbid_b200.h"

namespace RGBID_SLAM {

struct PixelRGB { unsigned char r, g, b; };  // include/types.h:88-91

struct PoseConstraint {
  enum Type { SEQ_ODO = RGBID_SEQ_ODO, SEQ_KF = RGBID_SEQ_KF };
  PoseConstraint() : ini_id_(0), end_id_(0), type_(SEQ_ODO), scale_(1.f)
  {
    std::memset(rotation_, 0, sizeof(rotation_)); std::memset(translation_, 0, sizeof(translation_));
    std::memset(covariance_, 0, sizeof(covariance_));
  }
  PoseConstraint(int ini_id, int end_id, Type type, const double* R, const double* t, float scale, const double* cov)
      : ini_id_(ini_id), end_id_(end_id), type_(type), scale_(scale)
  {
    std::memcpy(rotation_, R, sizeof(rotation_)); std::memcpy(translation_, t, sizeof(translation_));
    std::memcpy(covariance_, cov, sizeof(covariance_));
  }
  int ini_id_, end_id_;
  Type type_;
  double rotation_[9], translation_[3];
  float scale_;
  double covariance_[36];
};

struct Keyframe {
  Keyframe() : id_(0), cols_(0), rows_(0) {}
  explicit Keyframe(const rgbid_keyframe_handoff& k) : id_(k.kf_index), cols_(k.cols), rows_(k.rows)
  {
    const double K[9] = {k.fx, 0, k.cx, 0, k.fy, k.cy, 0, 0, 1};
    std::memcpy(K_, K, sizeof(K_));
    std::memcpy(rotation_, k.R, sizeof(rotation_)); std::memcpy(translation_, k.t, sizeof(translation_));
    std::memcpy(rotation_rel_, k.rel_R, sizeof(rotation_rel_)); std::memcpy(translation_rel_, k.rel_t, sizeof(translation_rel_));
    const size_t n = (size_t)k.rows * k.cols;
    colors_.resize(n); depthinv_.resize(n); normals_.resize(3 * n); overlap_mask_.resize(n);
    std::memcpy(colors_.data(), k.colors, n * 3);
    for (int y = 0; y < k.rows; ++y) {
      std::memcpy(&depthinv_[(size_t)y * k.cols], (const char*)k.depthinv + (size_t)y * k.depthinv_pitch, k.cols * sizeof(float));
      std::memcpy(&overlap_mask_[(size_t)y * k.cols], k.overlap_mask + (size_t)y * k.overlap_mask_pitch, k.cols);
    }
    for (int y = 0; y < 3 * k.rows; ++y)  // x, y, z planes stacked (DeviceArray2D<float> of 3 * rows, src/cuda/maps.cu)
      std::memcpy(&normals_[(size_t)y * k.cols], (const char*)k.normals + (size_t)y * k.normals_pitch, k.cols * sizeof(float));
  }
  std::vector<PixelRGB> colors_;
  std::vector<float> depthinv_;
  std::vector<float> normals_;
  std::vector<unsigned char> overlap_mask_;
  double K_[9];
  int id_, cols_, rows_;
  double rotation_[9], translation_[3];          // global pose of the keyframe
  double rotation_rel_[9], translation_rel_[3];  // constraint to the keyframe that replaced it
};
typedef std::shared_ptr<Keyframe> KeyframePtr;

// The two containers of KeyframeManager the tracker writes to (buffer_keyframes_, constraints_;
// src/visodo.cpp:1645-1657, 2147-2154)
struct KeyframeBuffers {
  std::vector<KeyframePtr> buffer_keyframes_;
  std::vector<PoseConstraint> constraints_;
};

}  // namespace RGBID_SLAM
