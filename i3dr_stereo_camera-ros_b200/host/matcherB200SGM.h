// MatcherB200SGM -- the B200-native SGM matcher as a sibling of MatcherOpenCVSGBM in the reference's plugin layer.
// It implements the same abstract-matcher contract (abstractStereoMatcher.h:12-92) on top of the C ABI in
// include/b200sgm.h; the setters only record RAW values (cv::StereoSGBM's defaulting rules are applied by the
// engine), so the 12 setters that updateMatcher() pushes per reconfigure (generate_disparity.cpp:241-261) are free.
#ifndef MATCHERB200SGM_H
#define MATCHERB200SGM_H

#ifdef B200SGM_STANDALONE
#include "matcher_interface.h"
#else
#include "stereoMatcher/abstractStereoMatcher.h"
#endif
#include "b200sgm.h"

#include <vector>

// What the two nodes either side of the matcher read from the CameraInfo messages (generate_disparity.cpp:441-450,
// disparity_to_depth.cpp:62-93): K of the left camera and the projection matrices of both, row-major.
struct B200StereoCamera
{
  double K_l[9];
  double P_l[12];
  double P_r[12];
};

class MatcherB200SGM : public AbstractStereoMatcher
{
public:
  explicit MatcherB200SGM(std::string &param_file, cv::Size _image_size, int cuda_device = 0)
      : AbstractStereoMatcher(param_file, _image_size), device_(cuda_device)
  {
    image_size = _image_size;
    init();
  }
  ~MatcherB200SGM();

  int forwardMatch(void);
  int backwardMatch(void);

  void setMinDisparity(int min_disparity);
  void setDisparityRange(int disparity_range);
  void setWindowSize(int window_size);
  void setUniquenessRatio(int ratio);
  void setSpeckleFilterWindow(int window);
  void setSpeckleFilterRange(int range);
  void setP1(float p1);
  void setP2(float p2);
  void setDisp12MaxDiff(int diff);
  void setInterpolation(bool enable);
  void setPreFilterCap(int cap);
  // extension (the cfg's `fullDP` flag is never read by the reference; this is its natural hook)
  void setFullDP(bool enable);

  // Row N1 of SURVEY.md section 8f: match + processDisparity (generate_disparity.cpp:426-452) + the reprojection of
  // disparity_to_depth.cpp:136-205 in one device pass on the images handed to setImages().  dmat = the CV_32FC1 image the node
  // publishes in stereo_msgs/DisparityImage (disparity in pixels, outside the depth window -> 10000), depth = the CV_32FC1 depth
  // image, cloud = the XYZRGB points in row-major scan order.  color: NULL (left image, MONO8) or a CV_8UC1 / BGR8 image of the
  // frame's size given as data pointer + step + channels.  Returns 0 / -1 like forwardMatch().
  int matchToCloud(const B200StereoCamera &cam, double depth_min, double depth_max, const unsigned char *color, size_t color_step,
                   int color_channels, cv::Mat &dmat, cv::Mat &depth, std::vector<b200sgm_point> &cloud);

  // Not used by SGBM (BM / I3DRSGM only), same as matcherOpenCVSGBM.h:32-35
  void setTextureThreshold(int threshold) {}
  void setPreFilterSize(int size) {}
  void setOcclusionDetection(bool enable) {}

  const char *lastError() const;

private:
  void init(void);
  int ensureEngine(int width, int height);

  int checkResult(int rc, const char *what);
  int report(const std::string &msg);
  int run(const b200sgm_params &p, const cv::Mat &l, const cv::Mat &r, float *&buf, size_t &cap, cv::Mat &out, const char *what);
  float *pinned(float *&buf, size_t &cap, size_t n);

  b200sgm_handle engine_ = nullptr;
  b200sgm_params params_;
  // page-locked result buffers (disparity_lr / disparity_rl / dmat / depth are headers over them; the cloud is staged in points_)
  float *lr_buf_ = nullptr, *rl_buf_ = nullptr, *dmat_buf_ = nullptr, *depth_buf_ = nullptr;
  size_t lr_cap_ = 0, rl_cap_ = 0, dmat_cap_ = 0, depth_cap_ = 0;
  b200sgm_point *points_ = nullptr;
  size_t points_cap_ = 0;
  bool warned_ = false;
  int device_ = 0;
  int cap_w_ = 0, cap_h_ = 0, cap_d_ = 0;
  std::string error_;
};

#endif // MATCHERB200SGM_H
