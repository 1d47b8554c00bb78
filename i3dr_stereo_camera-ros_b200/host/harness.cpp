// ROS-free harness that drives MatcherB200SGM exactly the way the reference node does
// (init_matcher + updateMatcher + stereo_match of /root/reference/src/generate_disparity.cpp:241-368 and the warm-up of
// src/init_stereo_matchers.cpp:39-66).  Reads two raw 8-bit images, writes the CV_32F disparity the node would receive.
//
//   harness <left.raw> <right.raw> <width> <height> <out.f32> minD D window uniq speckleRange speckleSize cap p1 p2 [fullDP]
//           [algorithm: 0 = B200 SGM (default), 1 = B200 block matcher] [textureThreshold] [--bench N]
// --bench N: after the first match, times N more stereo_match() calls on the same pair (the per-frame cost the node pays:
// setImages copy + match + getDisparity copy) and prints "bench_ms_per_frame <ms>" on stderr.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iostream>
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
  int exitCode = matcher->match();
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
  for (int i = 15; i + 1 < argc; i++)
    if (!strcmp(argv[i], "--bench")) bench_frames = atoi(argv[i + 1]);

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
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < bench_frames; i++) {
      cv::Mat d = stereo_match(left, right);
      if (d.empty()) return 1;
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / bench_frames;
    std::cerr << "bench_ms_per_frame " << ms << std::endl;
  }
  std::ofstream o(argv[5], std::ios::binary);
  for (int y = 0; y < disp.rows; y++) o.write(reinterpret_cast<const char *>(disp.data + y * disp.step), size_t(disp.cols) * 4);
  // mismatched sizes must leave the previous images untouched and still succeed (abstractStereoMatcher.cpp:21-24)
  cv::Mat small(cv::Size(W / 2, H), CV_8UC1);
  matcher->setImages(&left, &small);
  return matcher->match() == 0 ? 0 : 1;
}
