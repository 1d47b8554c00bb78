// MatcherB200SGM: adapter between the reference's AbstractStereoMatcher contract and the C ABI of the engine.
// Mirrors the behaviour of MatcherOpenCVSGBM (/root/reference/src/stereoMatcher/matcherOpenCVSGBM.cpp):
//   init()          -> default parameters of cv::StereoSGBM::create(64, 9, 5) i.e. (minDisparity 64, numDisparities 9,
//                      blockSize 5) -- harmless there and here because updateMatcher() overwrites all three
//   forwardMatch()  -> disparity_lr = CV_32FC1 holding disparity x16 (matcherOpenCVSGBM.cpp:21,34); 0 ok / -1 error
//                      with the message on std::cerr (matcherOpenCVSGBM.cpp:37-43); no exception leaves the matcher
//   backwardMatch() -> only reachable from the WLS `interp` branch (needs ximgproc): unsupported, returns -1
#include "matcherB200SGM.h"

#include <algorithm>
#include <iostream>

void MatcherB200SGM::init(void)
{
  // what cv::StereoSGBM::create(64, 9, 5) leaves behind (positional: minDisparity, numDisparities, blockSize);
  // every other field is OpenCV's create() default
  params_.minDisparity = 64;
  params_.numDisparities = 9;
  params_.blockSize = 5;
  params_.P1 = 0;
  params_.P2 = 0;
  params_.disp12MaxDiff = 0;
  params_.preFilterCap = 0;
  params_.uniquenessRatio = 0;
  params_.speckleWindowSize = 0;
  params_.speckleRange = 0;
  params_.mode = B200SGM_MODE_SGBM;
}

MatcherB200SGM::~MatcherB200SGM()
{
  if (engine_) b200sgm_destroy(engine_);
}

const char *MatcherB200SGM::lastError() const
{
  return error_.c_str();
}

// (Re)creates the engine when the frame or the disparity range outgrows what was allocated.  The reference builds its
// matchers lazily with the first frame's size (generate_disparity.cpp:342-346) but re-checks nothing; we re-check per call.
int MatcherB200SGM::ensureEngine(int width, int height)
{
  const int d = std::max(params_.numDisparities, 1);
  if (engine_ && width <= cap_w_ && height <= cap_h_ && d <= cap_d_) return 0;
  if (engine_) { b200sgm_destroy(engine_); engine_ = nullptr; }
  cap_w_ = std::max(width, cap_w_);
  cap_h_ = std::max(height, cap_h_);
  cap_d_ = std::max((d + 63) / 64 * 64, cap_d_);
  const int rc = b200sgm_create(device_, cap_w_, cap_h_, cap_d_, 1, &engine_);
  if (rc != B200SGM_OK) {
    engine_ = nullptr;
    cap_w_ = cap_h_ = cap_d_ = 0;
    error_ = "b200sgm_create failed (no usable CUDA device or out of device memory)";
    return rc;
  }
  return 0;
}

int MatcherB200SGM::forwardMatch()
{
  if (left == nullptr || right == nullptr || left->empty() || right->empty()) {
    std::cerr << "Error in B200 SGM matcher: no images set" << std::endl;
    return -1;
  }
  if (interpolate) {
    // WLS filtering (matcherOpenCVSGBM.cpp:22-33) needs cv::ximgproc and a right-view matcher: not provided
    std::cerr << "Error in B200 SGM matcher: interpolation (WLS) is not supported" << std::endl;
    return -1;
  }
  const int w = left->cols, h = left->rows;
  int rc = ensureEngine(w, h);
  if (rc == 0) rc = b200sgm_set_params(engine_, &params_);
  if (rc == 0) {
    if (disparity_lr.rows != h || disparity_lr.cols != w || disparity_lr.type() != CV_32FC1)
      disparity_lr = cv::Mat(cv::Size(w, h), CV_32FC1);
    rc = b200sgm_compute_f32(engine_, left->data, left->step, right->data, right->step, w, h,
                             reinterpret_cast<float *>(disparity_lr.data), disparity_lr.step);
    if (rc != 0) error_ = b200sgm_last_error(engine_);
  }
  if (rc != 0) {
    std::cerr << "Error in B200 SGM parameters" << std::endl;
    std::cerr << error_ << std::endl;
    return -1;
  }
  return 0;
}

int MatcherB200SGM::backwardMatch()
{
  return -1;
}

void MatcherB200SGM::setMinDisparity(int min_disparity)
{
  params_.minDisparity = min_disparity;
  this->min_disparity = min_disparity;
}

void MatcherB200SGM::setDisparityRange(int disparity_range)
{
  // same defaulting as matcherOpenCVSGBM.cpp:59-64
  disparity_range = disparity_range > 0 ? disparity_range : ((image_size.width / 8) + 15) & -16;
  this->disparity_range = disparity_range;
  params_.numDisparities = disparity_range;
}

void MatcherB200SGM::setWindowSize(int window_size)
{
  this->window_size = window_size;
  params_.blockSize = window_size;
}

void MatcherB200SGM::setUniquenessRatio(int ratio) { params_.uniquenessRatio = ratio; }
void MatcherB200SGM::setSpeckleFilterWindow(int window) { params_.speckleWindowSize = window; }
void MatcherB200SGM::setSpeckleFilterRange(int range) { params_.speckleRange = range; }
void MatcherB200SGM::setDisp12MaxDiff(int diff) { params_.disp12MaxDiff = diff; }
void MatcherB200SGM::setInterpolation(bool enable) { this->interpolate = enable; }
// float -> implicit int, like matcher->setP1(p1) on cv::StereoSGBM (matcherOpenCVSGBM.cpp:97-105)
void MatcherB200SGM::setP1(float p1) { params_.P1 = int(p1); }
void MatcherB200SGM::setP2(float p2) { params_.P2 = int(p2); }
void MatcherB200SGM::setPreFilterCap(int cap) { params_.preFilterCap = cap; }
void MatcherB200SGM::setFullDP(bool enable) { params_.mode = enable ? B200SGM_MODE_HH : B200SGM_MODE_SGBM; }
