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

}  // namespace b200sgm
