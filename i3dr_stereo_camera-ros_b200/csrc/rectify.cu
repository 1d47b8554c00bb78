// Host side of the rectification stage (k_rectify.cuh): the fixed-point bicubic weight table of cv::remap, the camera
// model of cv::initUndistortRectifyMap, and the launchers.  Replaces rectify() of
// /root/reference/src/generate_disparity.cpp:370-386 (and src/rectify.cpp:111-127).
#include <cmath>
#include <cstring>
#include <vector>
#include "stages.h"
#include "k_rectify.cuh"

namespace b200sgm {

// cv::remap's INTER_CUBIC table for 8-bit images: 32 x 32 sub-pixel positions x 16 int16 weights.  float32 arithmetic
// throughout (A = -0.75), weights = round-half-even(v * 2^15), and the sum is forced to 2^15 by correcting the largest
// (deficit) or smallest (excess) weight found in rows/columns 2..3 of the 4x4 set -- OpenCV's rule, kept as is.
void build_cubic_table(int16_t* tab /* [1024][16] */)
{
    float c1[32][4];
    const float A = -0.75f, scale = 1.f / 32;
    for (int i = 0; i < 32; i++) {
        const float x = i * scale;
        c1[i][0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c1[i][1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c1[i][2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c1[i][3] = 1.f - c1[i][0] - c1[i][1] - c1[i][2];
    }
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            int16_t* it = tab + (i * 32 + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < 4; k1++)
                for (int k2 = 0; k2 < 4; k2++) {
                    const float v = c1[i][k1] * c1[j][k2];
                    long q = std::lrintf(v * 32768.f);
                    q = q < -32768 ? -32768 : (q > 32767 ? 32767 : q);
                    it[k1 * 4 + k2] = int16_t(q);
                    isum += int(q);
                }
            if (isum != 32768) {
                const int diff = isum - 32768;
                int Mk = 2 * 4 + 2, mk = 2 * 4 + 2;
                for (int k1 = 2; k1 < 4; k1++)
                    for (int k2 = 2; k2 < 4; k2++) {
                        if (it[k1 * 4 + k2] < it[mk]) mk = k1 * 4 + k2;
                        else if (it[k1 * 4 + k2] > it[Mk]) Mk = k1 * 4 + k2;
                    }
                if (diff < 0) it[Mk] = int16_t(it[Mk] - diff);
                else it[mk] = int16_t(it[mk] - diff);
            }
        }
}

// K, R: 3x3 row-major; D: nD distortion coefficients (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4 [tauX tauY]]]]); P: 3x4 row-major.
// Returns false for a singular P*R or a tilted sensor model (unsupported).
bool make_rectify_cam(const double* K, const double* D, int nD, const double* R, const double* P, RectifyCam& c)
{
    double k[14] = {0};
    for (int i = 0; i < nD && i < 14; i++) k[i] = D[i];
    if (k[12] != 0.0 || k[13] != 0.0) return false;
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!R) R = I3;
    double m[9];
    for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) m[r * 3 + q] = P[r * 4 + 0] * R[0 * 3 + q] + P[r * 4 + 1] * R[1 * 3 + q] + P[r * 4 + 2] * R[2 * 3 + q];
    // cv::invert on a 3x3: cofactors / determinant
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (det == 0.0) return false;
    const double d = 1. / det;
    c.ir[0] = (m[4] * m[8] - m[5] * m[7]) * d; c.ir[1] = (m[2] * m[7] - m[1] * m[8]) * d; c.ir[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    c.ir[3] = (m[5] * m[6] - m[3] * m[8]) * d; c.ir[4] = (m[0] * m[8] - m[2] * m[6]) * d; c.ir[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    c.ir[6] = (m[3] * m[7] - m[4] * m[6]) * d; c.ir[7] = (m[1] * m[6] - m[0] * m[7]) * d; c.ir[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    c.fx = K[0]; c.fy = K[4]; c.u0 = K[2]; c.v0 = K[5];
    c.k1 = k[0]; c.k2 = k[1]; c.p1 = k[2]; c.p2 = k[3]; c.k3 = k[4]; c.k4 = k[5]; c.k5 = k[6]; c.k6 = k[7];
    c.s1 = k[8]; c.s2 = k[9]; c.s3 = k[10]; c.s4 = k[11];
    return true;
}

void launch_rectify_maps(const RectifyCam& c, int W, int H, RemapEntry* ent, float* map1, float* map2, cudaStream_t st)
{
    dim3 block(128), grid((W + 127) / 128, H);
    k_rectify_maps<<<grid, block, 0, st>>>(c, W, H, ent, map1, map2);
}

void launch_remap_cubic(const uint8_t* src, size_t spitch, int SW, int SH, const RemapEntry* ent, const int16_t* wtab, uint8_t* dst,
                        size_t dpitch, int W, int H, cudaStream_t st)
{
    dim3 block(256), grid((W + 256 * kRmPPT - 1) / (256 * kRmPPT), H);
    k_remap_cubic<<<grid, block, 0, st>>>(src, spitch, SW, SH, ent, wtab, dst, dpitch, W, H);
}

}  // namespace b200sgm
