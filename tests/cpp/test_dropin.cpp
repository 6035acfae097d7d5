// test_dropin.cpp -- uses the drop-in C++ surface exactly the way the reference's own callers do
// (src/visodo.cpp / src/keyframe_align.cpp): DeviceArray2D buffers, RGBID_SLAM::device::* bridge functions,
// VisodoTracker::trackNewFrame, KeyframeAlign::alignKeyframes.  Input: a binary sequence written by
// tests/test_cpp_dropin_gpu.py; output: a text file of results that the Python test compares with the oracle.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

#include "../../rgbid-slam_b200/host/visodo.hpp"

namespace RGBID_SLAM { namespace device {
cudaDeviceProp dev_prop;  // the application defines these (tools/RGBID_SLAMapp.cpp:68-69)
int dev_id = 0;
} }

using namespace RGBID_SLAM;
using namespace RGBID_SLAM::device;

struct Seq {
  int n, rows, cols;
  float fx, fy, cx, cy;
  std::vector<std::vector<unsigned short> > depth;
  std::vector<std::vector<unsigned char> > rgb;
};

static Seq read_seq(const char* path)
{
  Seq s;
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  int hdr[3];
  float intr[4];
  if (fread(hdr, sizeof(int), 3, f) != 3 || fread(intr, sizeof(float), 4, f) != 4) exit(2);
  s.n = hdr[0]; s.rows = hdr[1]; s.cols = hdr[2];
  s.fx = intr[0]; s.fy = intr[1]; s.cx = intr[2]; s.cy = intr[3];
  s.depth.resize(s.n); s.rgb.resize(s.n);
  for (int k = 0; k < s.n; ++k) {
    s.depth[k].resize((size_t)s.rows * s.cols);
    s.rgb[k].resize((size_t)s.rows * s.cols * 3);
    if (fread(s.depth[k].data(), 2, s.depth[k].size(), f) != s.depth[k].size()) exit(2);
    if (fread(s.rgb[k].data(), 1, s.rgb[k].size(), f) != s.rgb[k].size()) exit(2);
  }
  fclose(f);
  return s;
}

int main(int argc, char** argv)
{
  if (argc < 4) { fprintf(stderr, "usage: test_dropin <sequence.bin> <calibration.ini> <out.txt>\n"); return 2; }
  cudaSetDevice(dev_id);
  cudaGetDeviceProperties(&dev_prop, dev_id);
  Seq S = read_seq(argv[1]);
  FILE* out = fopen(argv[3], "w");

  // ---- 1. one Gauss-Newton iteration through the bridge functions, as in src/visodo.cpp:1108-1226 -----------------
  {
    DepthMap depth0, depth1;
    View rgb0, rgb1;
    depth0.upload(S.depth[0].data(), S.cols * 2, S.rows, S.cols);
    depth1.upload(S.depth[1].data(), S.cols * 2, S.rows, S.cols);
    rgb0.upload(S.rgb[0].data(), S.cols * 3, S.rows, S.cols);
    rgb1.upload(S.rgb[1].data(), S.cols * 3, S.rows, S.cols);
    DepthMapf W0, W1, Wwarp;
    IntensityMapf I0, I1, Iwarp;
    GradientMap gWx, gWy, gIx, gIy;
    Intr intr(S.fx, S.fy, S.cx, S.cy);
    convertDepth2InvDepth(depth0, W0, 1.f);
    convertDepth2InvDepth(depth1, W1, 1.f);
    computeIntensity(PtrStepSz<uchar3>(S.rows, S.cols, (uchar3*)rgb0.ptr(), rgb0.step()), I0);
    computeIntensity(PtrStepSz<uchar3>(S.rows, S.cols, (uchar3*)rgb1.ptr(), rgb1.step()), I1);
    computeGradientDepth(W0, gWx, gWy);
    computeGradientIntensity(I0, gIx, gIy);
    Wwarp.create(S.rows, S.cols); Iwarp.create(S.rows, S.cols);
    Mat33 Rid; Rid.data[0] = make_float3(1, 0, 0); Rid.data[1] = make_float3(0, 1, 0); Rid.data[2] = make_float3(0, 0, 1);
    float3 tz = make_float3(0, 0, 0);
    float ms = 0.f;
    ms += warpInvDepthWithTrafo3D(W1, Wwarp, W0, Rid, tz, intr);
    ms += warpIntensityWithTrafo3DInvDepth(I1, Iwarp, Wwarp, Rid, tz, intr);
    DeviceArray<float> resI, resW;
    ms += computeErrorGridStride(Iwarp, I0, resI, 10000);
    ms += computeErrorGridStride(Wwarp, W0, resW, 10000);
    float bias_i = 0.f, sigma_i = 5.f, nu_i = 5.f, bias_w = 0.f, sigma_w = 0.0025f, nu_w = 5.f;
    ms += computeSigmaAndNuStudent(resI, bias_i, sigma_i, nu_i, STUDENT);
    ms += computeSigmaAndNuStudent(resW, bias_w, sigma_w, nu_w, STUDENT);
    nu_i = std::max(nu_i, nu_w);
    double A[36], b[6];
    DeviceArray2D<float_type> gbuf;
    DeviceArray<float_type> sumbuf;
    ms += buildSystemStudentNuGridStride(tz, tz, W0, I0, gWx, gWy, gIx, gIy, Wwarp, Iwarp, STUDENT, INDEPENDENT, sigma_w,
                                         sigma_i, bias_w, bias_i, nu_w, nu_i, intr, B_SIZE, gbuf, sumbuf, A, b);
    float vis = -1.f;
    getVisibilityRatio(W1, W0, Rid, tz, intr, vis, 0.0125f);
    fprintf(out, "scale %.9g %.9g %.9g %.9g %.9g %.9g\n", sigma_i, sigma_w, bias_i, bias_w, nu_i, nu_w);
    fprintf(out, "A");
    for (int i = 0; i < 36; ++i) fprintf(out, " %.17g", A[i]);
    fprintf(out, "\nb");
    for (int i = 0; i < 6; ++i) fprintf(out, " %.17g", b[i]);
    fprintf(out, "\nvis %.9g\nelapsed_ms %.4f\n", vis, ms);
  }

  // ---- 2. VisodoTracker over the sequence (tools/RGBID_SLAMapp.cpp:163-214 feeds it the same way) ---------------
  {
    std::ifstream cfg(argv[2]);
    VisodoTracker tracker(6, STUDENT, CONSTANT_VELOCITY, SIGMA_PDF, INDEPENDENT, PYR_FIRST, DEFAULT_ODO_KF_COUNT, 0,
                          ALL_ITERS, DEFAULT_VISRATIO_ODO, NO_FILTERS, DEFAULT_VISRATIO_INTEGR, DEFAULT_INTEGR_KF_COUNT,
                          10000, S.rows, S.cols);
    tracker.loadCalibration(argv[2]);
    Settings settings(cfg);
    tracker.loadSettings(settings);
    for (int k = 0; k < S.n; ++k) {
      tracker.depth_.upload(S.depth[k].data(), S.cols * 2, S.rows, S.cols);
      tracker.rgb24_.upload(S.rgb[k].data(), S.cols * 3, S.rows, S.cols);
      bool ok = tracker.trackNewFrame();
      Affine3 p = tracker.getCameraPose();
      fprintf(out, "pose %d %d", k, ok ? 1 : 0);
      for (int i = 0; i < 9; ++i) fprintf(out, " %.17g", p.R[i]);
      for (int i = 0; i < 3; ++i) fprintf(out, " %.17g", p.t[i]);
      fprintf(out, " %d %d\n", tracker.lastResult().new_odo_keyframe, tracker.lastResult().new_integr_keyframe);
    }
    // hand-off containers (keyframe_manager_ptr_->buffer_keyframes_ / constraints_ in the reference)
    const KeyframeBuffers& kb = tracker.keyframe_buffers_;
    int n_seq_odo = 0, n_seq_kf = 0;
    for (const PoseConstraint& pc : kb.constraints_) (pc.type_ == PoseConstraint::SEQ_ODO ? n_seq_odo : n_seq_kf)++;
    fprintf(out, "handoff %d %d %d", (int)kb.buffer_keyframes_.size(), n_seq_odo, n_seq_kf);
    for (const KeyframePtr& k : kb.buffer_keyframes_) {
      double s = 0; int valid = 0;
      for (float v : k->depthinv_) if (v == v) { s += v; ++valid; }
      fprintf(out, " %d %d %.9g %d", k->id_, valid, s, (int)k->colors_[k->colors_.size() / 2].g);
    }
    fprintf(out, "\n");
  }

  // ---- 3. KeyframeAlign::alignKeyframes between frames 0 and 2 (src/loop_closer.cpp:319 calls it like this) ------
  {
    std::vector<float> W[2];
    std::vector<unsigned char> G[2];
    const int idx[2] = {0, 2};
    for (int j = 0; j < 2; ++j) {
      const auto& d = S.depth[idx[j]];
      const auto& c = S.rgb[idx[j]];
      W[j].resize(d.size()); G[j].resize(d.size());
      for (size_t i = 0; i < d.size(); ++i) {
        W[j][i] = d[i] > 0 ? 1000.f / (float)std::min<int>(d[i], 10000) : nanf("");
        float v = 0.2126f * c[3 * i] + 0.7152f * c[3 * i + 1] + 0.0722f * c[3 * i + 2];
        G[j][i] = (unsigned char)(v + 0.5f);
      }
    }
    KeyframeView a, e;
    a.rows = e.rows = S.rows; a.cols = e.cols = S.cols;
    const double K[9] = {S.fx, 0, S.cx, 0, S.fy, S.cy, 0, 0, 1};
    std::memcpy(a.K, K, sizeof(K)); std::memcpy(e.K, K, sizeof(K));
    a.depthinv = W[0].data(); a.grey = G[0].data(); e.depthinv = W[1].data(); e.grey = G[1].data();
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0}, cov[36];
    KeyframeAlign aligner;
    aligner.alignKeyframes(a, e, R, t, cov);
    fprintf(out, "align");
    for (int i = 0; i < 9; ++i) fprintf(out, " %.17g", R[i]);
    for (int i = 0; i < 3; ++i) fprintf(out, " %.17g", t[i]);
    fprintf(out, " %.17g\n", cov[0]);
  }
  fclose(out);
  printf("test_dropin: ok\n");
  return 0;
}
