// ROS-free harness that drives MatcherB200SGM exactly the way the reference node does
// (init_matcher + updateMatcher + stereo_match of /root/reference/src/generate_disparity.cpp:241-368 and the warm-up of
// src/init_stereo_matchers.cpp:39-66).  Reads two raw 8-bit images, writes the CV_32F disparity the node would receive.
//
//   harness <left.raw> <right.raw> <width> <height> <out.f32> minD D window uniq speckleRange speckleSize cap p1 p2 [fullDP]
//           [algorithm: 0 = B200 SGM (default), 1 = B200 block matcher] [textureThreshold] [--bench N]
// --bench N: after the first match, times N more stereo_match() calls on the same pair (the per-frame cost the node pays:
// setImages copy + match + getDisparity copy) and prints "bench_ms_per_frame <ms>" on stderr.
// --back out.f32: also backwardMatch() + getBackDisparity() (the right-view disparity of matcherOpenCVSGBM.cpp:46-51).
// --cloud prefix fx cx cxr cy p14 depth_min depth_max [color.raw channels]: MatcherB200SGM::matchToCloud with the CameraInfo
//   K_l = [fx 0 cx; 0 fx cy; 0 0 1], P_l = [K_l | 0], P_r = P_l with cx -> cxr and P_r[0][3] = p14; writes prefix.dmat.f32,
//   prefix.depth.f32 and prefix.cloud.bin (uint32 count, then count x {x, y, z, rgb}); --bench then times matchToCloud too
//   ("bench_cloud_ms_per_frame").
// --bad-prefilter-size N: pushes setPreFilterSize(N) before matching (cv::StereoBM rejects even or out-of-range sizes).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "matcherB200SGM.h"
#include "matcherB200BM.h"

static AbstractStereoMatcher *matcher = nullptr;
static MatcherB200SGM *b200sgm_matcher = nullptr;
static MatcherB200BM *b200bm_matcher = nullptr;
static int _stereo_algorithm = 0;   // 0: B200 SGM, 1: B200 block matcher
static bool isFirstImagesRecevied = false;

static int _min_disparity = 9, _disparity_range = 64, _correlation_window_size = 15, _uniqueness_ratio = 15;
static int _texture_threshold = 10, _speckle_size = 100, _speckle_range = 4, _preFilterCap = 31, _preFilterSize = 9;
static float _p1 = 200, _p2 = 400;
static bool _interp = false, _fullDP = false;

static void updateMatcher()   // generate_disparity.cpp:241-261, same order
{
  matcher->setDisparityRange(_disparity_range);
  matcher->setWindowSize(_correlation_window_size);
  matcher->setMinDisparity(_min_disparity);
  matcher->setUniquenessRatio(_uniqueness_ratio);
  matcher->setSpeckleFilterRange(_speckle_range);
  matcher->setSpeckleFilterWindow(_speckle_size);
  matcher->setPreFilterCap(_preFilterCap);
  matcher->setP1(_p1);
  matcher->setP2(_p2);
  matcher->setTextureThreshold(_texture_threshold);
  matcher->setPreFilterSize(_preFilterSize);
  matcher->setInterpolation(_interp);
  if (_stereo_algorithm == 0) b200sgm_matcher->setFullDP(_fullDP);   // the hook the reference never wired
}

static void init_matcher(cv::Size image_size)   // generate_disparity.cpp:263-331
{
  std::string empty_str = " ";
  b200sgm_matcher = new MatcherB200SGM(empty_str, image_size);
  b200bm_matcher = new MatcherB200BM(empty_str, image_size);
  matcher = _stereo_algorithm == 1 ? static_cast<AbstractStereoMatcher *>(b200bm_matcher) : b200sgm_matcher;
  updateMatcher();
}

static double g_match_ms = 0;     // time inside matcher->match() (the adapter + engine), of the whole stereo_match() call
static cv::Mat stereo_match(cv::Mat left_image, cv::Mat right_image)   // generate_disparity.cpp:334-368
{
  cv::Mat disp;
  cv::Size image_size(left_image.size().width, left_image.size().height);
  cv::Mat(image_size, CV_32F).copyTo(disp);
  if (!isFirstImagesRecevied) {
    init_matcher(image_size);
    isFirstImagesRecevied = true;
  }
  matcher->setDownsampleScale(1);
  matcher->setImages(&left_image, &right_image);
  const auto tm0 = std::chrono::steady_clock::now();
  int exitCode = matcher->match();
  g_match_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tm0).count();
  if (exitCode == 0) {
    matcher->getDisparity(disp);
  } else {
    std::cerr << "Exit code:" << exitCode << " Failed to compute stereo match" << std::endl;
    return cv::Mat();
  }
  return disp;
}

static bool read_raw(const char *path, cv::Mat &m)
{
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  f.read(reinterpret_cast<char *>(m.data), size_t(m.rows) * m.cols);
  return bool(f);
}

int main(int argc, char **argv)
{
  if (argc < 15) {
    std::cerr << "usage: harness left.raw right.raw W H out.f32 minD D window uniq speckleRange speckleSize cap p1 p2 [fullDP]" << std::endl;
    return 2;
  }
  const int W = atoi(argv[3]), H = atoi(argv[4]);
  _min_disparity = atoi(argv[6]); _disparity_range = atoi(argv[7]); _correlation_window_size = atoi(argv[8]);
  _uniqueness_ratio = atoi(argv[9]); _speckle_range = atoi(argv[10]); _speckle_size = atoi(argv[11]);
  _preFilterCap = atoi(argv[12]); _p1 = float(atof(argv[13])); _p2 = float(atof(argv[14]));
  _fullDP = argc > 15 && atoi(argv[15]) != 0;
  _stereo_algorithm = argc > 16 ? atoi(argv[16]) : 0;
  if (argc > 17) _texture_threshold = atoi(argv[17]);
  int bench_frames = 0;
  const char *back_path = nullptr, *cloud_prefix = nullptr, *color_path = nullptr;
  int color_channels = 0;
  double cam[7] = {0, 0, 0, 0, 0, 0, 0};   // fx cx cxr cy p14 depth_min depth_max
  for (int i = 15; i < argc; i++) {
    if (!strcmp(argv[i], "--bench") && i + 1 < argc) bench_frames = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--back") && i + 1 < argc) back_path = argv[++i];
    else if (!strcmp(argv[i], "--bad-prefilter-size") && i + 1 < argc) _preFilterSize = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--cloud") && i + 8 < argc) {
      cloud_prefix = argv[++i];
      for (int k = 0; k < 7; k++) cam[k] = atof(argv[++i]);
      if (i + 2 < argc && argv[i + 1][0] != '-') { color_path = argv[++i]; color_channels = atoi(argv[++i]); }
    }
  }

  // warm-up exactly like init_stereo_matchers.cpp:41-56: a 10x10 zero pair through setImages/match/getDisparity
  {
    cv::Mat l0 = cv::Mat::zeros(cv::Size(10, 10), CV_8UC1), r0 = cv::Mat::zeros(cv::Size(10, 10), CV_8UC1);
    cv::Mat d0 = stereo_match(l0, r0);
    // the block matcher refuses a 10x10 pair when the window does not fit, exactly like cv::StereoBM throws in the reference
    if (d0.empty() && _stereo_algorithm == 0) { std::cerr << "warm-up failed" << std::endl; return 1; }
  }
  cv::Mat left(cv::Size(W, H), CV_8UC1), right(cv::Size(W, H), CV_8UC1);
  if (!read_raw(argv[1], left) || !read_raw(argv[2], right)) { std::cerr << "cannot read input" << std::endl; return 2; }
  cv::Mat disp = stereo_match(left, right);
  if (disp.empty()) return 1;
  if (disp.type() != CV_32F) { std::cerr << "unexpected disparity type" << std::endl; return 1; }
  if (bench_frames > 0) {
    g_match_ms = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < bench_frames; i++) {
      cv::Mat d = stereo_match(left, right);
      if (d.empty()) return 1;
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / bench_frames;
    std::cerr << "bench_ms_per_frame " << ms << std::endl;
    std::cerr << "bench_match_ms_per_frame " << g_match_ms / bench_frames << std::endl;
  }
  std::ofstream o(argv[5], std::ios::binary);
  for (int y = 0; y < disp.rows; y++) o.write(reinterpret_cast<const char *>(disp.data + y * disp.step), size_t(disp.cols) * 4);
  if (back_path && _stereo_algorithm == 0) {
    if (matcher->backwardMatch() != 0) return 1;
    cv::Mat back;
    matcher->getBackDisparity(back);
    std::ofstream ob(back_path, std::ios::binary);
    for (int y = 0; y < back.rows; y++) ob.write(reinterpret_cast<const char *>(back.data + y * back.step), size_t(back.cols) * 4);
  }
  if (cloud_prefix && _stereo_algorithm == 0) {
    B200StereoCamera c;
    const double K[9] = {cam[0], 0, cam[1], 0, cam[0], cam[3], 0, 0, 1};
    const double Pl[12] = {cam[0], 0, cam[1], 0, 0, cam[0], cam[3], 0, 0, 0, 1, 0};
    const double Pr[12] = {cam[0], 0, cam[2], cam[4], 0, cam[0], cam[3], 0, 0, 0, 1, 0};
    memcpy(c.K_l, K, sizeof K); memcpy(c.P_l, Pl, sizeof Pl); memcpy(c.P_r, Pr, sizeof Pr);
    std::vector<unsigned char> color;
    if (color_path) {
      color.resize(size_t(W) * H * color_channels);
      std::ifstream fc(color_path, std::ios::binary);
      fc.read(reinterpret_cast<char *>(color.data()), color.size());
      if (!fc) { std::cerr << "cannot read the colour image" << std::endl; return 2; }
    }
    cv::Mat dmat, depth;
    std::vector<b200sgm_point> cloud;
    matcher->setImages(&left, &right);
    auto once = [&] {
      return b200sgm_matcher->matchToCloud(c, cam[5], cam[6], color_path ? color.data() : nullptr, size_t(W) * color_channels, color_channels,
                                           dmat, depth, cloud);
    };
    if (once() != 0) return 1;
    const std::string pre(cloud_prefix);
    std::ofstream o1(pre + ".dmat.f32", std::ios::binary), o2(pre + ".depth.f32", std::ios::binary), o3(pre + ".cloud.bin", std::ios::binary);
    for (int y = 0; y < H; y++) {
      o1.write(reinterpret_cast<const char *>(dmat.data + y * dmat.step), size_t(W) * 4);
      o2.write(reinterpret_cast<const char *>(depth.data + y * depth.step), size_t(W) * 4);
    }
    const uint32_t n = uint32_t(cloud.size());
    o3.write(reinterpret_cast<const char *>(&n), 4);
    o3.write(reinterpret_cast<const char *>(cloud.data()), size_t(n) * sizeof(b200sgm_point));
    if (bench_frames > 0) {
      const auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < bench_frames; i++) {
        matcher->setImages(&left, &right);
        if (once() != 0) return 1;
      }
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / bench_frames;
      std::cerr << "bench_cloud_ms_per_frame " << ms << std::endl;
    }
  }
  // mismatched sizes must leave the previous images untouched and still succeed (abstractStereoMatcher.cpp:21-24)
  cv::Mat small(cv::Size(W / 2, H), CV_8UC1);
  matcher->setImages(&left, &small);
  return matcher->match() == 0 ? 0 : 1;
}
