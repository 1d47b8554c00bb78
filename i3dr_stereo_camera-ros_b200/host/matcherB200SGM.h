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

  // Not used by SGBM (BM / I3DRSGM only), same as matcherOpenCVSGBM.h:32-35
  void setTextureThreshold(int threshold) {}
  void setPreFilterSize(int size) {}
  void setOcclusionDetection(bool enable) {}

  const char *lastError() const;

private:
  void init(void);
  int ensureEngine(int width, int height);

  b200sgm_handle engine_ = nullptr;
  b200sgm_params params_;
  int device_ = 0;
  int cap_w_ = 0, cap_h_ = 0, cap_d_ = 0;
  std::string error_;
};

#endif // MATCHERB200SGM_H
