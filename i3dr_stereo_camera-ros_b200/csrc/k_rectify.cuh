// Row N2 of SURVEY.md section 8(f): the step before the matcher.  The reference rectifies both images of every frame with
//   cv::initUndistortRectifyMap(K, D, R, P, size, CV_32FC1, map1, map2);
//   cv::remap(image, image_rect, map1, map2, cv::INTER_CUBIC, cv::BORDER_CONSTANT);
// (/root/reference/src/generate_disparity.cpp:370-386, src/rectify.cpp:111-127), recomputing the maps per frame.
// Here the maps are built once per camera (k_rectify_maps) and kept in HBM in the fixed-point form remap() converts
// them to anyway; k_remap_cubic then is one gather of 16 taps per pixel.
#pragma once
#include "sgm_types.h"

namespace b200sgm {

__device__ __forceinline__ int sat_s16(int v) { return min(max(v, -32768), 32767); }

// Double-precision map computation, operation for operation as the scalar loop of cv::initUndistortRectifyMap (explicit
// _rn intrinsics: no FMA contraction), except that the homogeneous coordinates are formed directly (j*ir0 + (i*ir1 + ir2))
// instead of by repeated addition along the row; the float32 maps agree with OpenCV's to within 1 ulp in a handful of
// pixels per 5 Mpixel and the fixed-point coordinates remap() derives from them are identical (tests).
static __global__ void k_rectify_maps(RectifyCam c, int W, int H, RemapEntry* __restrict__ ent, float* __restrict__ map1,
                                      float* __restrict__ map2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= W || i >= H) return;
    const double dj = double(j), di = double(i);
    const double _x = __dadd_rn(__dmul_rn(dj, c.ir[0]), __dadd_rn(__dmul_rn(di, c.ir[1]), c.ir[2]));
    const double _y = __dadd_rn(__dmul_rn(dj, c.ir[3]), __dadd_rn(__dmul_rn(di, c.ir[4]), c.ir[5]));
    const double _w = __dadd_rn(__dmul_rn(dj, c.ir[6]), __dadd_rn(__dmul_rn(di, c.ir[7]), c.ir[8]));
    const double w = __ddiv_rn(1.0, _w), x = __dmul_rn(_x, w), y = __dmul_rn(_y, w);
    const double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y);
    const double r2 = __dadd_rn(x2, y2), _2xy = __dmul_rn(__dmul_rn(2.0, x), y);
    auto poly = [&](double a3, double a2, double a1) {   // 1 + ((a3*r2 + a2)*r2 + a1)*r2
        return __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(a3, r2), a2), r2), a1), r2));
    };
    const double kr = __ddiv_rn(poly(c.k3, c.k2, c.k1), poly(c.k6, c.k5, c.k4));
    double xd = __dmul_rn(x, kr);
    xd = __dadd_rn(xd, __dmul_rn(c.p1, _2xy));
    xd = __dadd_rn(xd, __dmul_rn(c.p2, __dadd_rn(r2, __dmul_rn(2.0, x2))));
    xd = __dadd_rn(xd, __dmul_rn(c.s1, r2));
    xd = __dadd_rn(xd, __dmul_rn(__dmul_rn(c.s2, r2), r2));
    double yd = __dmul_rn(y, kr);
    yd = __dadd_rn(yd, __dmul_rn(c.p1, __dadd_rn(r2, __dmul_rn(2.0, y2))));
    yd = __dadd_rn(yd, __dmul_rn(c.p2, _2xy));
    yd = __dadd_rn(yd, __dmul_rn(c.s3, r2));
    yd = __dadd_rn(yd, __dmul_rn(__dmul_rn(c.s4, r2), r2));
    const float u = float(__dadd_rn(__dmul_rn(c.fx, xd), c.u0));
    const float v = float(__dadd_rn(__dmul_rn(c.fy, yd), c.v0));
    const size_t p = size_t(i) * W + j;
    if (map1) { map1[p] = u; map2[p] = v; }
    // cv::remap: sx = cvRound(map * INTER_TAB_SIZE) (round half to even), tap origin = (sx >> 5) - 1
    const int sx = __float2int_rn(__fmul_rn(u, 32.0f)), sy = __float2int_rn(__fmul_rn(v, 32.0f));
    RemapEntry e;
    e.x = int16_t(sat_s16(sx >> 5)); e.y = int16_t(sat_s16(sy >> 5));
    e.frac = uint16_t(((sy & 31) << 5) | (sx & 31)); e.pad = 0;
    ent[p] = e;
}

// One output pixel (device helper): entry -> weights -> 16 taps.
__device__ __forceinline__ uint8_t remap_cubic_pixel(const uint8_t* __restrict__ src, size_t spitch, int SW, int SH, RemapEntry e,
                                                     const int16_t* __restrict__ wtab)
{
    const int sx = int(e.x) - 1, sy = int(e.y) - 1;
    const uint4* wp = reinterpret_cast<const uint4*>(wtab + size_t(e.frac) * 16);
    const uint4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
    const uint32_t wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    int sum = 0;
    if (sx >= 4 && sx + 8 <= SW && unsigned(sy) < unsigned(max(SH - 3, 0))) {
        // interior fast path: the 4 taps of a row come from the two aligned 32-bit words that cover them (both inside the
        // row), one PRMT to align, two DP2A (int16 weights x uint8 pixels) to accumulate
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const uint8_t* s = src + size_t(sy + a) * spitch + sx;
            const uintptr_t ad = reinterpret_cast<uintptr_t>(s);
            const uint32_t* wa = reinterpret_cast<const uint32_t*>(ad & ~uintptr_t(3));
            const uint32_t px = __byte_perm(__ldg(wa), __ldg(wa + 1), 0x3210u + 0x1111u * uint32_t(ad & 3));
            asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(sum) : "r"(wv[a * 2]), "r"(px));
            asm("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(sum) : "r"(wv[a * 2 + 1]), "r"(px));
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int yy = sy + a;
            if (yy < 0 || yy >= SH) continue;
            const uint8_t* s = src + size_t(yy) * spitch;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int xx = sx + b;
                const int wt = int(int16_t((b & 1) ? (wv[a * 2 + b / 2] >> 16) : (wv[a * 2 + b / 2] & 0xFFFFu)));
                if (xx >= 0 && xx < SW) sum += int(s[xx]) * wt;
            }
        }
    }
    const int r = (sum + (1 << 14)) >> 15;
    return uint8_t(min(max(r, 0), 255));
}

// cv::remap(..., INTER_CUBIC, BORDER_CONSTANT, 0) on CV_8UC1: 4x4 taps, int16 weights summing to 2^15, taps outside the
// source contribute 0, result = saturate_u8((sum + 2^14) >> 15).  wtab: [1024][16] int16 (host-built, rectify.cu).
// A thread produces kRmPPT pixels 256 columns apart (coalesced entry loads and stores per instruction, kRmPPT independent
// entry -> gather chains in flight).
constexpr int kRmPPT = 4;
static __global__ void __launch_bounds__(256) k_remap_cubic(const uint8_t* __restrict__ src, size_t spitch, int SW, int SH,
                                                           const RemapEntry* __restrict__ ent, const int16_t* __restrict__ wtab,
                                                           uint8_t* __restrict__ dst, size_t dpitch, int W, int H)
{
    const int j0 = blockIdx.x * (256 * kRmPPT) + threadIdx.x, i = blockIdx.y;
    if (i >= H) return;
    RemapEntry e[kRmPPT];
#pragma unroll
    for (int k = 0; k < kRmPPT; k++) {
        const int j = j0 + 256 * k;
        const uint2 raw = j < W ? __ldg(reinterpret_cast<const uint2*>(ent + size_t(i) * W + j)) : make_uint2(0u, 0u);
        e[k].x = int16_t(raw.x & 0xFFFFu); e[k].y = int16_t(raw.x >> 16); e[k].frac = uint16_t(raw.y & 0xFFFFu); e[k].pad = 0;
    }
    uint8_t out[kRmPPT];
#pragma unroll
    for (int k = 0; k < kRmPPT; k++) out[k] = remap_cubic_pixel(src, spitch, SW, SH, e[k], wtab);
#pragma unroll
    for (int k = 0; k < kRmPPT; k++) {
        const int j = j0 + 256 * k;
        if (j < W) dst[size_t(i) * dpitch + j] = out[k];
    }
}

}  // namespace b200sgm
