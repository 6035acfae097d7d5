// rgbid_slam_app.cpp -- RGBID_SLAMapp-compatible command-line driver for the tracking hot path.
//
// Mirrors the evaluation mode of tools/RGBID_SLAMapp.cpp (:376-500): the same flags, the same sequence reader
// (tools/evaluation.cpp, restated in rgbid-slam_b200/host/tum_io.hpp), the same upload -> trackNewFrame loop
// (simulateLoopCallback, :163-214, without the 30 ms pacing sleep and the visualisation thread) and the same pose log
// ("<dataset>_poses.txt", one "stamp tx ty tz qx qy qz qw" line per frame, evaluation.cpp:424-436), so the reference's
// own evaluation procedure (TUM's evaluate_ate / evaluate_rpe on that log) runs unchanged on this implementation.
// Out of scope, as in DESIGN.md: live capture (-dev / ROS), keyframe manager, loop closing, point-cloud export.
//
//   rgbid_slam_app -eval <folder/> [-match_file <file>] [-config <visodoRGBDconfig.ini>] [-calib <calibration.ini>]
//                  [-gpu <id>] [-n <frames>] [-o <poses log>]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "../rgbid-slam_b200/host/tum_io.hpp"
#include "../rgbid-slam_b200/host/visodo.hpp"

// the application defines the device globals the library layer reads, as tools/RGBID_SLAMapp.cpp:68-69 does
cudaDeviceProp RGBID_SLAM::device::dev_prop;
int RGBID_SLAM::device::dev_id;

using namespace RGBID_SLAM;

static bool find_switch(int argc, char** argv, const char* name)
{
  for (int i = 1; i < argc; ++i)
    if (std::string(argv[i]) == name) return true;
  return false;
}

// pcl::console::parse_argument: value following the LAST occurrence of the flag
template <class T>
static int parse_argument(int argc, char** argv, const char* name, T& out, T (*conv)(const char*))
{
  int idx = -1;
  for (int i = 1; i + 1 < argc; ++i)
    if (std::string(argv[i]) == name) idx = i + 1;
  if (idx > 0) out = conv(argv[idx]);
  return idx;
}
static std::string to_str(const char* s) { return std::string(s); }
static int to_int(const char* s) { return atoi(s); }

static int print_cli_help()
{
  std::cout << "\nRGBID-SLAM (B200 tracking path) parameters:\n"
            << "    --help, -h                 : print this message\n"
            << "    -gpu <id>                  : CUDA device\n"
            << "    -config <file>             : [VISODO] settings (config_data/visodoRGBDconfig.ini dialect)\n"
            << "    -calib <file>              : [CALIBRATION] fx fy cx cy factor_depth (config_data/calibration_*.ini)\n"
            << "    -eval <folder/>            : TUM RGB-D sequence folder (rgb/, depth/, *_associated.txt)\n"
            << "    -match_file <file>         : association file 't_d depth/.. t_rgb rgb/..' inside the folder\n"
            << "    -n <frames>                : stop after this many frames\n"
            << "    -o <file>                  : pose log (default <dataset>_poses.txt)\n";
  return 0;
}

int main(int argc, char** argv)
{
  if (find_switch(argc, argv, "--help") || find_switch(argc, argv, "-h")) return print_cli_help();
  if (find_switch(argc, argv, "-check_io")) {
    // host-only self check of the sequence reader (no CUDA call): one line per frame with sizes and checksums, then
    // the quaternion of a fixed rotation; tests/test_tum_io_cpu.py compares it with OpenCV / numpy
    std::string folder, match;
    parse_argument<std::string>(argc, argv, "-eval", folder, to_str);
    parse_argument<std::string>(argc, argv, "-match_file", match, to_str);
    try {
      tum::Sequence seq(folder, match);
      std::cout << "associations " << seq.size() << std::endl;
      std::vector<uint16_t> depth;
      std::vector<uint8_t> rgb;
      int rows = 0, cols = 0;
      for (size_t i = 0; i < seq.size(); ++i) {
        if (!seq.grab(i, depth, rgb, rows, cols)) { std::cout << "frame " << i << " grab failed" << std::endl; continue; }
        unsigned long long sd = 0, sc = 0;
        for (size_t k = 0; k < depth.size(); ++k) sd += (unsigned long long)depth[k] * (k % 251 + 1);
        for (size_t k = 0; k < rgb.size(); ++k) sc += (unsigned long long)rgb[k] * (k % 251 + 1);
        std::cout.setf(std::ios::fixed, std::ios::floatfield);
        std::cout << "frame " << i << " " << seq[i].time1 << " " << seq[i].time2 << " " << rows << " " << cols << " " << sd
                  << " " << sc << std::endl;
      }
      const float R[9] = {0.36f, 0.48f, -0.8f, -0.8f, 0.6f, 0.f, 0.48f, 0.64f, 0.6f};
      float q[4];
      tum::quaternion_from_rotation(R, q);
      std::cout << "quat " << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
    } catch (const std::exception& e) {
      std::cout << "rgbid_slam_app: " << e.what() << std::endl;
      return 1;
    }
    return 0;
  }
  device::dev_id = 0;
  parse_argument<int>(argc, argv, "-gpu", device::dev_id, to_int);
  cudaSafeCall(cudaSetDevice(device::dev_id));
  cudaSafeCall(cudaGetDeviceProperties(&device::dev_prop, device::dev_id));

  std::string poses_logfile("poses"), misc_logfile("misc"), kf_times_logfile("kf_times"), config_file, calib_file, eval_folder,
      match_file, out_override;
  int max_frames = -1;
  parse_argument<std::string>(argc, argv, "-config", config_file, to_str);
  parse_argument<std::string>(argc, argv, "-calib", calib_file, to_str);
  parse_argument<std::string>(argc, argv, "-match_file", match_file, to_str);
  parse_argument<std::string>(argc, argv, "-o", out_override, to_str);
  parse_argument<int>(argc, argv, "-n", max_frames, to_int);
  if (parse_argument<std::string>(argc, argv, "-eval", eval_folder, to_str) <= 0) {
    std::cout << "This build runs the evaluation mode only: give -eval <folder/> (live capture is out of scope)" << std::endl;
    return print_cli_help(), 1;
  }
  {  // "<dataset>_poses.txt" from the folder name, tools/RGBID_SLAMapp.cpp:414-428
    std::size_t found_last = eval_folder.find_last_of("/\\");
    std::string eval_folder2 = eval_folder.substr(0, found_last);
    std::size_t found_prelast = eval_folder2.find_last_of("/\\");
    std::string dataset_name = eval_folder2.substr(found_prelast + 1);
    poses_logfile = dataset_name + "_" + poses_logfile;
    misc_logfile = dataset_name + "_" + misc_logfile;
    kf_times_logfile = dataset_name + "_" + kf_times_logfile;
  }
  poses_logfile.append(".txt");
  misc_logfile.append(".txt");
  kf_times_logfile.append(".txt");
  if (!out_override.empty()) {  // -o <file>: the two time logs go next to it
    poses_logfile = out_override;
    const std::string stem = out_override.substr(0, out_override.find_last_of('.'));
    misc_logfile = stem + "_misc.txt";
    kf_times_logfile = stem + "_kf_times.txt";
  }
  const bool inline_tracking = find_switch(argc, argv, "-inline");

  try {
    tum::Sequence seq(eval_folder, match_file);
    std::vector<uint16_t> depth;
    std::vector<uint8_t> rgb;
    int rows = 0, cols = 0;
    if (!seq.grab(0, depth, rgb, rows, cols)) {
      std::cout << "Can't read the first frame of " << eval_folder << std::endl;
      return 1;
    }
    // the reference's tracker defaults (`new VisodoTracker`, tools/RGBID_SLAMapp.cpp:82, include/visodo.h:54-70) at the
    // size of the sequence; note WARP_ORDER: the code default is warpFirst, config_data/visodoRGBDconfig.ini says pyrFirst
    VisodoTracker visodo(6, device::STUDENT, device::CONSTANT_VELOCITY, device::SIGMA_PDF, device::INDEPENDENT,
                         device::DEFAULT_WARPING, device::DEFAULT_ODO_KF_COUNT, 0, device::ALL_ITERS, device::DEFAULT_VISRATIO_ODO,
                         device::NO_FILTERS, device::DEFAULT_VISRATIO_INTEGR, device::DEFAULT_INTEGR_KF_COUNT, 10000,
                         rows, cols);
    visodo.setRGBIntrinsics(525.f, 525.f, 319.5f, 239.5f);  // Evaluation::fx .. cy, tools/evaluation.cpp:61-64
    if (!config_file.empty()) {
      std::ifstream f(config_file.c_str());
      if (!f.is_open()) { std::cout << "Could not open configuration file " << config_file << std::endl; return 1; }
      Settings settings(f);
      visodo.loadSettings(settings);
    }
    if (!calib_file.empty()) visodo.loadCalibration(calib_file);

    std::vector<tum::PoseRt> poses;
    int num_failures = 0, lost = 0, odo_kf = 0;
    const auto t0 = std::chrono::steady_clock::now();
    // The reference's structure (tools/RGBID_SLAMapp.cpp:163-214, 459-460): the tracker runs in its own thread and waits
    // on new_frame_cond_; the grabber try-locks visodo.mutex_, uploads the frame and notifies.  The reference then sleeps
    // 30 ms (playback pacing, :209); an evaluation run must neither drop nor repeat a frame, so this grabber waits for
    // the tracker's frame counter instead.  -inline calls trackNewFrame() from this thread.
    if (!inline_tracking) visodo.start();
    for (size_t i = 0; num_failures < 10 && (max_frames < 0 || (int)poses.size() < max_frames); ++i) {
      if (i > 0 && !seq.grab(i, depth, rgb, rows, cols)) { ++num_failures; continue; }
      num_failures = 0;
      if (inline_tracking) {
        visodo.depth_.upload(depth.data(), (size_t)cols * 2, rows, cols);
        visodo.rgb24_.upload(rgb.data(), (size_t)cols * 3, rows, cols);
        visodo.trackNewFrame();
      } else {
        int before;
        for (;;) {
          std::unique_lock<std::mutex> lock(visodo.mutex_, std::try_to_lock);
          if (!lock) { std::this_thread::yield(); continue; }
          before = visodo.frames_tracked_by_thread_;
          visodo.depth_.upload(depth.data(), (size_t)cols * 2, rows, cols);
          visodo.rgb24_.upload(rgb.data(), (size_t)cols * 3, rows, cols);
          visodo.new_frame_cond_.notify_one();
          break;
        }
        for (;;) {  // the tracker holds mutex_ while it works on the frame
          std::unique_lock<std::mutex> lock(visodo.mutex_);
          if (visodo.frames_tracked_by_thread_ != before) break;
          lock.unlock();
          std::this_thread::yield();
        }
      }
      const Affine3 p = visodo.getCameraPose();
      tum::PoseRt pr;
      for (int k = 0; k < 9; ++k) pr.R[k] = p.R[k];
      for (int k = 0; k < 3; ++k) pr.t[k] = p.t[k];
      poses.push_back(pr);
      lost += visodo.visOdoIsLost() ? 1 : 0;
      odo_kf += visodo.lastResult().new_odo_keyframe;
    }
    visodo.stop();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "Writing " << poses.size() << " poses to " << poses_logfile << std::endl;
    tum::save_all_poses(poses_logfile, seq, poses);
    {  // Evaluation::saveAllPoses' time statistics and saveTimeLogFiles (tools/evaluation.cpp:353-420)
      std::vector<float> times;
      for (size_t i = 0; i < visodo.getNumberOfPoses(); ++i) times.push_back(visodo.getVisOdoTime((int)i));
      tum::save_misc_log(misc_logfile, times);
      tum::save_kf_times_log(kf_times_logfile, visodo.kf_times_);
    }
    std::cout << "frames " << poses.size() << "  lost " << lost << "  odometry keyframes " << odo_kf << "  "
              << (poses.size() / secs) << " frames/s including PNG decoding" << std::endl;
  } catch (const std::exception& e) {
    std::cout << "rgbid_slam_app: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
