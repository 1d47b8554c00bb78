// Shared host/device types of the B200 SGBM engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200sgm {

constexpr int kMaxCost = 32767;        // OpenCV's MAX_COST (SHRT_MAX); also the "infinity" of padded cells
constexpr uint32_t kMaxCostX2 = 0x7FFF7FFFu;
constexpr uint32_t kFullMask = 0xFFFFFFFFu;

// Parameters after OpenCV's defaulting rules plus derived geometry.
struct Eff {
    int W, H;        // image size
    int minD, D;     // minDisparity, numDisparities
    int Dp;          // padded disparity count of the cost volumes (multiple of 2*nreg)
    int nreg;        // packed uint32 (2 disparities each) held per lane by the aggregation kernels
    int SW2;         // blockSize/2
    int P1, P2;
    int d12;         // effective disp12MaxDiff (>= 1)
    int uniq;        // effective uniquenessRatio
    int ftzero;      // max(preFilterCap,15)|1
    int speckleWin, speckleRange;
    int mode;        // 0 SGBM (5 paths), 1 HH (8 paths)
    int INVALID;     // (minD-1)*16
    int minX1, W1;   // first valid column, number of valid columns
};

// One prefiltered pixel (A.2): Sobel-x value with its half-pixel interval and the raw value (x64, so that the
// ">> 2" of the raw cost is a byte extraction) with its interval, as signed 16-bit fields already in the form the
// packed Birchfield-Tomasi arithmetic consumes them:
//   x = sob | sob_lo << 16        y = (-sob_hi) | (-sob) << 16
//   z = raw64 | raw64_lo << 16    w = (-raw64_hi) | (-raw64) << 16
typedef uint4 Feat;

// Inverse of (P[:, :3] * R) and the distortion model of cv::initUndistortRectifyMap (no tilt: tauX = tauY = 0).
struct RectifyCam {
    double ir[9];
    double fx, fy, u0, v0;
    double k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4;
};

// One map entry: source position in 1/32 pixel units split like cv::remap does: integer part (minus the one-pixel
// offset of the 4x4 footprint, saturated to int16 first) and the index of the weight set.
struct RemapEntry {
    int16_t x, y;      // top-left tap of the 4x4 footprint
    uint16_t frac;     // (fy << 5) | fx, index into the 1024 x 16 weight table
    uint16_t pad;
};

}  // namespace b200sgm
