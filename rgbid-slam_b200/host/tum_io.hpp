// tum_io.hpp -- TUM RGB-D sequence input and pose-log output for the drop-in driver.
//
// Restates what tools/evaluation.cpp does around the hot path, without OpenCV / Eigen (neither is installed here):
//   * association lists: `depth_associated.txt` + `rgb_associated.txt` (three header lines each, then
//     "<stamp> <file>" per line, tools/evaluation.cpp:153-181) or one match file with
//     "<t_depth> <depth file> <t_rgb> <rgb file>" per line (:183-199).  Both readers loop on `!eof()`, so a trailing
//     newline yields one extra empty association; grab() then fails on it, exactly as the reference's does;
//   * depth: 16-bit PNG scaled by 0.2 to millimetres with cv::Mat::convertTo semantics -- saturate_cast<ushort> of
//     the value rounded half-to-even (:285, :333; the TUM files store 5000 units per metre);
//   * colour: 8-bit PNG as RGB (the reference reads BGR and swaps, :232-238);
//   * pose log: "<stamp> tx ty tz qx qy qz qw" in fixed notation, float precision 6, the quaternion taken from the
//     float rotation matrix the way Eigen::Quaternionf(Matrix3f) does (:424-436).
// The PNG decoder covers what the TUM sequences (and this repository's synthetic writer) contain: non-interlaced,
// 8-bit grey / RGB / RGBA and 16-bit grey, all five filter types; zlib does the inflation.
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace RGBID_SLAM {
namespace tum {

struct Association {
  double time1 = 0, time2 = 0;  // depth stamp, rgb stamp
  std::string name1, name2;     // depth file, rgb file (relative to the sequence folder)
};

struct Image {
  int rows = 0, cols = 0, channels = 0, bit_depth = 0;
  std::vector<uint8_t> data;  // row-major; 16-bit samples in host byte order
};

// ---- PNG -----------------------------------------------------------------------------------------------------
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool read_png(const std::string& path, Image& img)
{
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) return false;
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (buf.size() < 8 || memcmp(buf.data(), sig, 8) != 0) return false;
  size_t pos = 8;
  int width = 0, height = 0, bit_depth = 0, color_type = 0, interlace = 0;
  bool have_ihdr = false;
  std::vector<uint8_t> idat;
  // The files come from user-supplied dataset folders: every length is checked before it is used.  (Chunk CRCs are
  // not verified; the zlib stream carries its own Adler-32, which uncompress() checks.)
  while (pos + 12 <= buf.size()) {
    const size_t len = be32(&buf[pos]);
    const char* type = (const char*)&buf[pos + 4];
    if (len > buf.size() - pos - 12) return false;  // truncated chunk
    const uint8_t* d = &buf[pos + 8];
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13 || have_ihdr) return false;
      const uint32_t w = be32(d), h = be32(d + 4);
      if (w == 0 || h == 0 || w > 16384 || h > 16384) return false;  // bounds the allocations below
      width = (int)w; height = (int)h;
      bit_depth = d[8]; color_type = d[9]; interlace = d[12];
      have_ihdr = true;
    } else if (!memcmp(type, "IDAT", 4)) {
      if (!have_ihdr) return false;
      idat.insert(idat.end(), d, d + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  int channels = color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 6 ? 4 : color_type == 4 ? 2 : 0;
  if (!have_ihdr || idat.empty() || channels == 0 || interlace != 0 || (bit_depth != 8 && bit_depth != 16)) return false;
  const size_t bpp = (size_t)channels * bit_depth / 8, stride = bpp * width;
  std::vector<uint8_t> raw((stride + 1) * height);
  uLongf out_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) return false;
  img.rows = height; img.cols = width; img.channels = channels; img.bit_depth = bit_depth;
  img.data.assign(stride * height, 0);
  std::vector<uint8_t> zero(stride, 0);
  for (int y = 0; y < height; ++y) {
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t* in = &raw[(stride + 1) * y + 1];
    uint8_t* out = &img.data[stride * y];
    const uint8_t* up = y ? &img.data[stride * (y - 1)] : zero.data();
    for (size_t i = 0; i < stride; ++i) {
      int a = i >= bpp ? out[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0, v = in[i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) / 2; break;
        case 4: {
          int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: return false;
      }
      out[i] = (uint8_t)v;
    }
  }
  if (bit_depth == 16) {  // big endian in the file
    uint16_t* p = (uint16_t*)img.data.data();
    for (size_t i = 0; i < img.data.size() / 2; ++i) p[i] = (uint16_t)((img.data[2 * i] << 8) | img.data[2 * i + 1]);
  }
  return true;
}

// 16-bit grey PNG -> millimetres: convertTo(dst, type, 0.2) = saturate_cast<ushort>(cvRound(v * 0.2)), cvRound rounding
// half to even (tools/evaluation.cpp:285)
inline bool read_depth_mm(const std::string& path, std::vector<uint16_t>& depth, int& rows, int& cols)
{
  Image img;
  if (!read_png(path, img)) return false;
  if (img.bit_depth != 16 || img.channels != 1) {
    fprintf(stderr, "Image was not opened in 16-bit format: %s\n", path.c_str());
    return false;
  }
  rows = img.rows; cols = img.cols;
  depth.resize((size_t)rows * cols);
  const uint16_t* p = (const uint16_t*)img.data.data();
  for (size_t i = 0; i < depth.size(); ++i) {
    long r = std::lrint((double)p[i] * 0.2);  // default rounding mode: to nearest even
    depth[i] = (uint16_t)(r < 0 ? 0 : r > 65535 ? 65535 : r);
  }
  return true;
}

inline bool read_rgb(const std::string& path, std::vector<uint8_t>& rgb, int& rows, int& cols)
{
  Image img;
  if (!read_png(path, img) || img.bit_depth != 8) return false;
  rows = img.rows; cols = img.cols;
  rgb.resize((size_t)rows * cols * 3);
  for (size_t i = 0; i < (size_t)rows * cols; ++i) {
    const uint8_t* s = &img.data[i * img.channels];
    if (img.channels >= 3) { rgb[3 * i] = s[0]; rgb[3 * i + 1] = s[1]; rgb[3 * i + 2] = s[2]; }
    else { rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = s[0]; }  // imread without flags converts grey to 3 channels
  }
  return true;
}

// ---- associations ----------------------------------------------------------------------------------------------
class Sequence {
 public:
  // match_file empty: <folder>/depth_associated.txt + rgb_associated.txt (Evaluation::Evaluation, :118-141)
  Sequence(const std::string& folder, const std::string& match_file) : folder_(folder)
  {
    if (!folder_.empty() && folder_.back() != '/' && folder_.back() != '\\') folder_.push_back('/');
    if (!match_file.empty()) {
      std::ifstream iff((folder_ + match_file).c_str());
      if (!iff) throw std::runtime_error("Can't read " + match_file);
      while (!iff.eof()) {
        Association a;
        iff >> a.time1 >> a.name1 >> a.time2 >> a.name2;
        associations_.push_back(a);
      }
    } else {
      std::ifstream d((folder_ + "depth_associated.txt").c_str()), c((folder_ + "rgb_associated.txt").c_str());
      if (!d || !c) throw std::runtime_error("Can't read rgbd " + folder_ + "depth_associated.txt");
      std::string line;
      for (int i = 0; i < 3; ++i) { std::getline(d, line); std::getline(c, line); }  // three header lines each
      while (!d.eof() || !c.eof()) {
        Association a;
        d >> a.time1 >> a.name1;
        c >> a.time2 >> a.name2;
        associations_.push_back(a);
        if (d.fail() && c.fail()) break;  // both exhausted (the reference's loop ends the same way)
      }
    }
  }

  size_t size() const { return associations_.size(); }
  const Association& operator[](size_t i) const { return associations_[i]; }

  // Evaluation::grab(int, depth, rgb24), :309-356: false past the end or when a file cannot be decoded
  bool grab(size_t i, std::vector<uint16_t>& depth, std::vector<uint8_t>& rgb, int& rows, int& cols) const
  {
    if (i >= associations_.size() || associations_[i].name1.empty() || associations_[i].name2.empty()) return false;
    int r2 = 0, c2 = 0;
    if (!read_depth_mm(folder_ + associations_[i].name1, depth, rows, cols)) return false;
    if (!read_rgb(folder_ + associations_[i].name2, rgb, r2, c2)) return false;
    return r2 == rows && c2 == cols;
  }

 private:
  std::string folder_;
  std::vector<Association> associations_;
};

// ---- pose log --------------------------------------------------------------------------------------------------
// Eigen::Quaternionf(Matrix3f) (quaternionbase_assign_impl<Other,3,3>): returns x, y, z, w
inline void quaternion_from_rotation(const float* R, float* q)
{
  float t = R[0] + R[4] + R[8];
  if (t > 0.f) {
    t = std::sqrt(t + 1.f);
    q[3] = 0.5f * t;
    t = 0.5f / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.f);
    q[i] = 0.5f * t;
    t = 0.5f / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}

// Evaluation::saveAllPoses, :399-440: poses as double R (row-major) / t per frame, cast to float like Affine3f
struct PoseRt { double R[9]; double t[3]; };

inline void save_all_poses(const std::string& path, const Sequence& seq, const std::vector<PoseRt>& poses)
{
  std::ofstream out(path.c_str());
  out.setf(std::ios::fixed, std::ios::floatfield);
  for (size_t i = 0; i < poses.size() && i < seq.size(); ++i) {
    float Rf[9], q[4];
    for (int k = 0; k < 9; ++k) Rf[k] = (float)poses[i].R[k];
    quaternion_from_rotation(Rf, q);
    out << seq[i].time1 << " " << (float)poses[i].t[0] << " " << (float)poses[i].t[1] << " " << (float)poses[i].t[2] << " "
        << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
  }
}

// "<dataset>_misc.txt": mean / std / max tracking time per frame in ms, the header lines of
// Evaluation::saveAllPoses (tools/evaluation.cpp:397-420; the reference writes them to its chi-tests / misc log)
inline void save_misc_log(const std::string& path, const std::vector<float>& vis_odo_times_ms)
{
  const int n = (int)vis_odo_times_ms.size();
  float mean = 0.f, sd = 0.f, mx = 0.f;
  for (int i = 0; i < n; ++i) { mean += vis_odo_times_ms[i] / n; if (vis_odo_times_ms[i] > mx) mx = vis_odo_times_ms[i]; }
  for (int i = 0; i < n; ++i) sd += (vis_odo_times_ms[i] - mean) * (vis_odo_times_ms[i] - mean) / n;
  sd = std::sqrt(sd);
  std::ofstream out(path.c_str());
  out.setf(std::ios::fixed, std::ios::floatfield);
  out << "Mean time per frame: " << mean << std::endl << "Std time per frame: " << sd << std::endl
      << "Max time per frame: " << mx << std::endl;
}

// "<dataset>_kf_times.txt", Evaluation::saveTimeLogFiles (tools/evaluation.cpp:353-377): one line per keyframe handed to
// the back end.  The five KeyframeManager columns (total, segmentation, BoW, loop detection, pose graph) belong to the
// back end, which is out of scope: they are written as 0 so that the file keeps the reference's columns.
inline void save_kf_times_log(const std::string& path, const std::vector<float>& kf_times_ms)
{
  std::ofstream out(path.c_str());
  out.setf(std::ios::fixed, std::ios::floatfield);
  out << "ObtainKeyframe " << "ProcessKeyframeTotal " << "Segmentation " << "DescriptionBoW " << "LoopDetection "
      << "PoseGraphOptim" << std::endl;
  for (size_t i = 0; i < kf_times_ms.size(); ++i)
    out << kf_times_ms[i] << " " << 0.f << " " << 0.f << " " << 0.f << " " << 0.f << " " << 0.f << std::endl;
}

}  // namespace tum
}  // namespace RGBID_SLAM
