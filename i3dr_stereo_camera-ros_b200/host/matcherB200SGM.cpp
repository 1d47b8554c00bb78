// MatcherB200SGM: adapter between the reference's AbstractStereoMatcher contract and the C ABI of the engine.
// Mirrors the behaviour of MatcherOpenCVSGBM (/root/reference/src/stereoMatcher/matcherOpenCVSGBM.cpp):
//   init()          -> default parameters of cv::StereoSGBM::create(64, 9, 5) i.e. (minDisparity 64, numDisparities 9,
//                      blockSize 5) -- harmless there and here because updateMatcher() overwrites all three
//   forwardMatch()  -> disparity_lr = CV_32FC1 holding disparity x16 (matcherOpenCVSGBM.cpp:21,34); 0 ok / -1 error
//                      with the message on std::cerr (matcherOpenCVSGBM.cpp:37-43); no exception leaves the matcher
//   backwardMatch() -> the right-view disparity of cv::ximgproc::createRightMatcher(matcher) (matcherOpenCVSGBM.cpp:46-51): the same
//                      matcher with minDisparity -(minD + D) + 1, uniqueness 0, disp12MaxDiff 1000000, no speckle filter, run on
//                      (right, left).  The WLS filter itself (ximgproc, `interp`) is not provided: forwardMatch() with
//                      interpolation enabled returns -1.
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
  disparity_lr = cv::Mat();
  disparity_rl = cv::Mat();
  b200sgm_host_free(lr_buf_); b200sgm_host_free(rl_buf_); b200sgm_host_free(dmat_buf_); b200sgm_host_free(depth_buf_);
  b200sgm_host_free(points_);
  if (engine_) b200sgm_destroy(engine_);
}

// Page-locked buffer of at least n floats (kept and reused; the copies back from the device then run at full PCIe rate).
float *MatcherB200SGM::pinned(float *&buf, size_t &cap, size_t n)
{
  if (cap < n) {
    b200sgm_host_free(buf);
    buf = nullptr; cap = 0;
    void *p = nullptr;
    if (b200sgm_host_alloc(n * sizeof(float), &p) == B200SGM_OK) { buf = static_cast<float *>(p); cap = n; }
  }
  return buf;
}

const char *MatcherB200SGM::lastError() const
{
  return error_.c_str();
}

// -1 with the message on std::cerr, the reference's error shape (matcherOpenCVSGBM.cpp:37-43)
int MatcherB200SGM::report(const std::string &msg)
{
  error_ = msg;
  std::cerr << "Error in B200 SGM parameters" << std::endl;
  std::cerr << error_ << std::endl;
  return -1;
}

// 0 for success and for warnings (printed once: the result was delivered), -1 with the engine's message otherwise.
int MatcherB200SGM::checkResult(int rc, const char *what)
{
  if (rc == 0) return 0;
  if (rc > 0) {
    error_ = b200sgm_last_error(engine_);
    if (!warned_) std::cerr << "Warning in B200 SGM matcher (" << what << "): " << error_ << std::endl;
    warned_ = true;
    return 0;
  }
  return report(b200sgm_last_error(engine_));
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

// One match of (l, r) with parameters p into the page-locked buffer behind `out` (CV_32FC1 holding disparity x16).
int MatcherB200SGM::run(const b200sgm_params &p, const cv::Mat &l, const cv::Mat &r, float *&buf, size_t &cap, cv::Mat &out, const char *what)
{
  const int w = l.cols, h = l.rows;
  if (ensureEngine(w, h) != 0) return report(error_);
  if (b200sgm_set_params(engine_, &p) != 0) return report(b200sgm_last_error(engine_));
  if (!pinned(buf, cap, size_t(w) * h)) return report("page-locked host allocation failed");
  if (out.data != reinterpret_cast<unsigned char *>(buf) || out.rows != h || out.cols != w) out = cv::Mat(h, w, CV_32FC1, buf);
  return checkResult(b200sgm_compute_f32(engine_, l.data, l.step, r.data, r.step, w, h, buf, size_t(w) * sizeof(float)), what);
}

int MatcherB200SGM::forwardMatch()
{
  if (left == nullptr || right == nullptr || left->empty() || right->empty()) {
    std::cerr << "Error in B200 SGM matcher: no images set" << std::endl;
    return -1;
  }
  if (interpolate) {
    // WLS filtering (matcherOpenCVSGBM.cpp:22-33) needs cv::ximgproc's disparity filter: not provided
    std::cerr << "Error in B200 SGM matcher: interpolation (WLS) is not supported" << std::endl;
    return -1;
  }
  return run(params_, *left, *right, lr_buf_, lr_cap_, disparity_lr, "forwardMatch");
}

// cv::ximgproc::createRightMatcher(matcher)->compute(*right, *left, disparity_rl) of matcherOpenCVSGBM.cpp:46-51.  ximgproc builds
// the right-view matcher as StereoSGBM::create(-(minDisparity + numDisparities) + 1, numDisparities, blockSize) with
// uniquenessRatio 0, the left matcher's P1 / P2 / mode / preFilterCap, disp12MaxDiff 1000000 and speckleWindowSize 0.
int MatcherB200SGM::backwardMatch()
{
  if (left == nullptr || right == nullptr || left->empty() || right->empty()) {
    std::cerr << "Error in B200 SGM matcher: no images set" << std::endl;
    return -1;
  }
  b200sgm_params rp = params_;
  rp.minDisparity = -(params_.minDisparity + params_.numDisparities) + 1;
  rp.uniquenessRatio = 0;
  rp.disp12MaxDiff = 1000000;
  rp.speckleWindowSize = 0;
  rp.speckleRange = 0;
  return run(rp, *right, *left, rl_buf_, rl_cap_, disparity_rl, "backwardMatch");
}

int MatcherB200SGM::matchToCloud(const B200StereoCamera &cam, double depth_min, double depth_max, const unsigned char *color,
                                 size_t color_step, int color_channels, cv::Mat &dmat, cv::Mat &depth, std::vector<b200sgm_point> &cloud)
{
  if (left == nullptr || right == nullptr || left->empty() || right->empty()) {
    std::cerr << "Error in B200 SGM matcher: no images set" << std::endl;
    return -1;
  }
  const int w = left->cols, h = left->rows;
  const size_t n = size_t(w) * h;
  if (ensureEngine(w, h) != 0) return report(error_);
  if (b200sgm_set_params(engine_, &params_) != 0) return report(b200sgm_last_error(engine_));
  if (!pinned(dmat_buf_, dmat_cap_, n) || !pinned(depth_buf_, depth_cap_, n)) return report("page-locked host allocation failed");
  if (points_cap_ < n) {
    b200sgm_host_free(points_);
    points_ = nullptr; points_cap_ = 0;
    void *p = nullptr;
    if (b200sgm_host_alloc(n * sizeof(b200sgm_point), &p) != B200SGM_OK) return report("page-locked host allocation failed");
    points_ = static_cast<b200sgm_point *>(p); points_cap_ = n;
  }
  uint32_t count = 0;
  b200sgm_reproject rp;
  b200sgm_reproject_from_camera(&rp, cam.K_l, cam.P_l, cam.P_r, depth_min, depth_max);
  rp.color = color; rp.color_stride = color_step; rp.color_channels = color ? color_channels : 0;
  const int rc = b200sgm_compute_xyz(engine_, left->data, left->step, right->data, right->step, w, h, &rp, nullptr, 0, dmat_buf_,
                                     depth_buf_, points_, &count);
  if (checkResult(rc, "matchToCloud") != 0) return -1;
  dmat = cv::Mat(h, w, CV_32FC1, dmat_buf_);
  depth = cv::Mat(h, w, CV_32FC1, depth_buf_);
  cloud.assign(points_, points_ + count);
  return 0;
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
