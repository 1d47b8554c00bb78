// MatcherB200BM -- the B200 block matcher as a sibling of MatcherOpenCVBlock (the package's default algorithm,
// launch/stereo_matcher.launch:20) in the reference's plugin layer: same abstract-matcher contract
// (abstractStereoMatcher.h:12-92) on top of b200sgm_bm_compute (include/b200sgm.h).
#ifndef MATCHERB200BM_H
#define MATCHERB200BM_H

#ifdef B200SGM_STANDALONE
#include "matcher_interface.h"
#else
#include "stereoMatcher/abstractStereoMatcher.h"
#endif
#include "b200sgm.h"

class MatcherB200BM : public AbstractStereoMatcher
{
public:
  explicit MatcherB200BM(std::string &param_file, cv::Size _image_size, int cuda_device = 0)
      : AbstractStereoMatcher(param_file, _image_size), device_(cuda_device)
  {
    image_size = _image_size;
    init();
  }
  ~MatcherB200BM();

  int forwardMatch(void);
  int backwardMatch(void);

  // the setters of matcherOpenCVBlock.cpp:52-110
  void setMinDisparity(int min_disparity);
  void setDisparityRange(int disparity_range);
  void setWindowSize(int window_size);
  void setTextureThreshold(int threshold);
  void setUniquenessRatio(int ratio);
  void setSpeckleFilterWindow(int window);
  void setSpeckleFilterRange(int range);
  void setDisp12MaxDiff(int diff);
  void setInterpolation(bool enable);
  void setPreFilterCap(int cap);
  void setPreFilterSize(int size);     // PREFILTER_XSOBEL never reads it, but cv::StereoBM::compute validates it all the same
  // not used by the block matcher, same as matcherOpenCVBlock.h
  void setP1(float p1) {}
  void setP2(float p2) {}
  void setOcclusionDetection(bool enable) {}

private:
  void init(void);
  int ensureEngine(int width, int height);

  b200sgm_handle engine_ = nullptr;
  b200sgm_bm_params params_;
  int device_ = 0;
  int cap_w_ = 0, cap_h_ = 0, cap_d_ = 0;
  float *lr_buf_ = nullptr;     // page-locked; disparity_lr is a header over it
  size_t lr_cap_ = 0;
  std::string error_;
};

#endif // MATCHERB200BM_H
