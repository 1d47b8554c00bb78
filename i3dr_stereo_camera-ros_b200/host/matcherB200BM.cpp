// MatcherB200BM: adapter between the reference's AbstractStereoMatcher contract and b200sgm_bm_compute.
// Mirrors MatcherOpenCVBlock (/root/reference/src/stereoMatcher/matcherOpenCVBlock.cpp):
//   init()         -> the parameters cv::StereoBM::create(64, 9) leaves behind (:13-16): numDisparities 64, blockSize 9 and
//                     OpenCV's defaults (minDisparity 0, preFilterCap 31, textureThreshold 10, uniquenessRatio 15, no speckle
//                     filter, disp12MaxDiff -1)
//   forwardMatch() -> disparity_lr = CV_32FC1 holding disparity x16 (:20,34); 0 ok / -1 with the message on std::cerr when
//                     OpenCV would have thrown (:37-43)
#include "matcherB200BM.h"

#include <algorithm>
#include <iostream>
#include <vector>

void MatcherB200BM::init(void)
{
  params_.minDisparity = 0;
  params_.numDisparities = 64;
  params_.blockSize = 9;
  params_.preFilterCap = 31;
  params_.textureThreshold = 10;
  params_.uniquenessRatio = 15;
  params_.speckleWindowSize = 0;
  params_.speckleRange = 0;
  params_.disp12MaxDiff = -1;
  params_.preFilterSize = 0;   // StereoBM::create's default (9)
}

MatcherB200BM::~MatcherB200BM()
{
  disparity_lr = cv::Mat();
  b200sgm_host_free(lr_buf_);
  if (engine_) b200sgm_destroy(engine_);
}

int MatcherB200BM::ensureEngine(int width, int height)
{
  const int d = std::max(params_.numDisparities, 16);
  if (engine_ && width <= cap_w_ && height <= cap_h_ && d <= cap_d_) return 0;
  if (engine_) { b200sgm_destroy(engine_); engine_ = nullptr; }
  cap_w_ = std::max(width, cap_w_);
  cap_h_ = std::max(height, cap_h_);
  cap_d_ = std::max((d + 63) / 64 * 64, cap_d_);
  const int rc = b200sgm_create_bm(device_, cap_w_, cap_h_, cap_d_, 1, &engine_);
  if (rc != B200SGM_OK) {
    engine_ = nullptr;
    cap_w_ = cap_h_ = cap_d_ = 0;
    error_ = "b200sgm_create failed (no usable CUDA device or out of device memory)";
    return rc;
  }
  return 0;
}

int MatcherB200BM::forwardMatch()
{
  if (left == nullptr || right == nullptr || left->empty() || right->empty()) {
    std::cerr << "Error in B200 block matcher: no images set" << std::endl;
    return -1;
  }
  if (interpolate) {
    std::cerr << "Error in B200 block matcher: interpolation (WLS) is not supported" << std::endl;
    return -1;
  }
  const int w = left->cols, h = left->rows;
  int rc = ensureEngine(w, h);
  if (rc == 0 && lr_cap_ < size_t(w) * h) {
    b200sgm_host_free(lr_buf_);
    lr_buf_ = nullptr; lr_cap_ = 0;
    void *p = nullptr;
    if (b200sgm_host_alloc(size_t(w) * h * sizeof(float), &p) == B200SGM_OK) { lr_buf_ = static_cast<float *>(p); lr_cap_ = size_t(w) * h; }
    else { error_ = "page-locked host allocation failed"; rc = B200SGM_ECUDA; }
  }
  if (rc == 0) {
    // disparity_lr.convertTo(disparity_lr, CV_32FC1): same numeric value, still x16 (matcherOpenCVBlock.cpp:34) -- done on the device
    if (disparity_lr.data != reinterpret_cast<unsigned char *>(lr_buf_) || disparity_lr.rows != h || disparity_lr.cols != w)
      disparity_lr = cv::Mat(h, w, CV_32FC1, lr_buf_);
    rc = b200sgm_bm_compute_f32(engine_, &params_, left->data, left->step, right->data, right->step, w, h, lr_buf_, size_t(w) * sizeof(float));
    if (rc != 0) error_ = b200sgm_last_error(engine_);
  }
  if (rc != 0) {
    std::cerr << "Error in OpenCV StereoBM parameters" << std::endl;   // the reference's wording (matcherOpenCVBlock.cpp:40)
    std::cerr << error_ << std::endl;
    return -1;
  }
  return 0;
}

int MatcherB200BM::backwardMatch()
{
  return -1;
}

void MatcherB200BM::setMinDisparity(int min_disparity)
{
  params_.minDisparity = min_disparity;
  this->min_disparity = min_disparity;
}

void MatcherB200BM::setDisparityRange(int disparity_range)
{
  // same defaulting as matcherOpenCVBlock.cpp:58-63
  disparity_range = disparity_range > 0 ? disparity_range : ((image_size.width / 8) + 15) & -16;
  params_.numDisparities = disparity_range;
  this->disparity_range = disparity_range;
}

void MatcherB200BM::setWindowSize(int window_size)
{
  this->window_size = window_size;
  params_.blockSize = window_size;
}

void MatcherB200BM::setTextureThreshold(int threshold) { params_.textureThreshold = threshold; }
void MatcherB200BM::setUniquenessRatio(int ratio) { params_.uniquenessRatio = ratio; }
void MatcherB200BM::setSpeckleFilterWindow(int window) { params_.speckleWindowSize = window; }
void MatcherB200BM::setSpeckleFilterRange(int range) { params_.speckleRange = range; }
void MatcherB200BM::setDisp12MaxDiff(int diff) { params_.disp12MaxDiff = diff; }
void MatcherB200BM::setInterpolation(bool enable) { this->interpolate = enable; }
void MatcherB200BM::setPreFilterCap(int cap) { params_.preFilterCap = cap; }
void MatcherB200BM::setPreFilterSize(int size) { params_.preFilterSize = size; }
