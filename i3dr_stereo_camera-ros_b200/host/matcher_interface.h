// Stand-alone restatement of the plugin contract that a stereo matcher must satisfy in the reference package
// (class AbstractStereoMatcher, /root/reference/include/stereoMatcher/abstractStereoMatcher.h:12-92 and
// src/stereoMatcher/abstractStereoMatcher.cpp).  Used only when the adapter is built outside the reference
// tree (B200SGM_STANDALONE); inside the tree the adapter includes the reference's own header instead.
// The names and call semantics are the contract; the text is ours.
#pragma once
#include <iostream>
#include <string>
#include "cv_stub.h"

class AbstractStereoMatcher {
public:
    // (param_file, image_size): both ignored by the base class, exactly like the reference (.cpp:3-7)
    explicit AbstractStereoMatcher(std::string&, cv::Size) {}
    ~AbstractStereoMatcher() {}

    // --- the 17 hooks every matcher must provide (abstractStereoMatcher.h:27-57) ---
    virtual void setDisparityRange(int) = 0;
    virtual void setWindowSize(int) = 0;
    virtual void setInterpolation(bool) = 0;
    virtual void setMinDisparity(int) = 0;
    virtual void setUniquenessRatio(int) = 0;
    virtual void setSpeckleFilterWindow(int) = 0;
    virtual void setSpeckleFilterRange(int) = 0;
    virtual void setDisp12MaxDiff(int) = 0;
    virtual void setPreFilterCap(int) = 0;
    virtual void setTextureThreshold(int) = 0;
    virtual void setPreFilterSize(int) = 0;
    virtual void setP1(float) = 0;
    virtual void setP2(float) = 0;
    virtual void setOcclusionDetection(bool) = 0;
    virtual int forwardMatch() = 0;
    virtual int backwardMatch() = 0;
    virtual void init() = 0;

    // --- behaviour inherited by every matcher ---
    // same-size check, fresh copies scaled by downsample_scale (scale 1 == exact copy), .cpp:9-25
    void setImages(cv::Mat* l, cv::Mat* r)
    {
        if (!(l->size() == r->size())) {
            std::cerr << "Images MUST be the same resolution" << std::endl;   // no update, like the reference
            return;
        }
        left = new cv::Mat();
        right = new cv::Mat();
        cv::resize(*l, *left, cv::Size(), downsample_scale, downsample_scale, cv::INTER_CUBIC);
        cv::resize(*r, *right, cv::Size(), downsample_scale, downsample_scale, cv::INTER_CUBIC);
        image_size = left->size();
    }
    virtual void setDownsampleScale(double s) { downsample_scale = s; }
    void getDisparity(cv::Mat& dst) { disparity_lr.copyTo(dst); }        // deep copy, .cpp:32-36
    void getBackDisparity(cv::Mat& dst) { disparity_rl.copyTo(dst); }
    cv::Mat* getLeftImage() { return left; }
    cv::Mat* getRighttImage() { return right; }
    // forwardMatch then "convert to CV_32F" (a no-op on an already-float result), .cpp:44-53
    virtual int match()
    {
        const int rc = forwardMatch();
        if (rc == 0) disparity_lr.convertTo(disparity_lr, CV_32F);
        return rc;
    }

protected:
    cv::Mat* left = nullptr;
    cv::Mat* right = nullptr;
    cv::Mat disparity_rl, disparity_lr, disparity_scale;
    cv::Size image_size;
    double downsample_scale = 1;
    int min_disparity = 0, disparity_range = 64, window_size = 9;
    bool interpolate = false;
};
