// Stage launchers that live in their own translation units (so that a kernel change recompiles one file).
#pragma once
#include "sgm_types.h"

namespace b200sgm {

// A.2: Sobel-x + raw prefilter of both images (one launch).  stage_cost.cu
void launch_prefilter(const uint8_t* dL, size_t lp, const uint8_t* dR, size_t rp, int W, int H, int ftzero, Feat* featL, Feat* featR,
                      cudaStream_t st);

// A.3 + A.4: Birchfield-Tomasi cost + block sum into the paired-layout volume C[H][W1][Dp].  `generic_only` forces the
// validation kernel.  *launches is incremented per kernel launched; *errmsg is set for unsupported geometry.  stage_cost.cu
cudaError_t launch_cost(const Feat* fl, const Feat* fr, uint16_t* C, const Eff& e, bool generic_only, int num_sms, cudaStream_t st,
                        int* launches, const char** errmsg);

// A.7 left-right check (in place on disp), A.8 3x3 median, A.8 speckle filter (in place on img).  stage_post.cu
void launch_lrcheck(int16_t* disp, const uint32_t* disp2key, const Eff& e, cudaStream_t st);
void launch_median3(const int16_t* src, int16_t* dst, int W, int H, cudaStream_t st);
cudaError_t launch_speckle(int16_t* img, int* label, int* parent, int* runlen, int* csize, int W, int H, int newVal, int maxSize, int maxDiff,
                           cudaStream_t st, int* launches);
void launch_fill16(int16_t* p, int n, int16_t v, cudaStream_t st);

// Rectification (row N2): cv::initUndistortRectifyMap + cv::remap(INTER_CUBIC, BORDER_CONSTANT).  rectify.cu
void build_cubic_table(int16_t* tab /* [1024][16] */);
bool make_rectify_cam(const double* K, const double* D, int nD, const double* R, const double* P, RectifyCam& c);
void launch_rectify_maps(const RectifyCam& c, int W, int H, RemapEntry* ent, float* map1, float* map2, cudaStream_t st);
void launch_remap_cubic(const uint8_t* src, size_t spitch, int SW, int SH, const RemapEntry* ent, const int16_t* wtab, uint8_t* dst,
                        size_t dpitch, int W, int H, cudaStream_t st);

// StereoBM (row N4): cv::StereoBM::compute of matcherOpenCVBlock.cpp:13-20.  stage_bm.cu
struct BmParams {
    int minDisparity, numDisparities, blockSize, preFilterCap, textureThreshold, uniquenessRatio, speckleWindowSize, speckleRange;
};
cudaError_t launch_bm(const uint8_t* dL, size_t lp, const uint8_t* dR, size_t rp, int W, int H, const BmParams& p, uint8_t* pre /* 2*W*H */,
                      uint16_t* HS, int* HT, int16_t* disp, int* label, int* parent, int* runlen, int* csize, int num_sms,
                      cudaStream_t st, int* launches);

}  // namespace b200sgm
