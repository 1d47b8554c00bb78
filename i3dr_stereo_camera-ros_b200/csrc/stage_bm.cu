// Launcher of the StereoBM path (row N4; kernels in k_bm.cuh).
#include <algorithm>
#include "stages.h"
#include "k_bm.cuh"

namespace b200sgm {

#define BM_LAUNCH_CHECK() do { ++*launches; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return e__; } while (0)

cudaError_t launch_bm(const uint8_t* dL, size_t lp, const uint8_t* dR, size_t rp, int W, int H, const BmParams& p, uint8_t* pre /* 2*W*H */,
                      uint16_t* HS, int* HT, int16_t* disp, int* label, int* parent, int* runlen, int* csize, int num_sms,
                      cudaStream_t st, int* launches)
{
    BmGeom g;
    g.W = W; g.H = H; g.ndisp = p.numDisparities; g.mindisp = p.minDisparity; g.wsz = p.blockSize; g.cap = p.preFilterCap;
    g.tex = p.textureThreshold; g.uniq = p.uniquenessRatio;
    g.lofs = std::max(g.ndisp - 1 + g.mindisp, 0);
    g.rofs = -std::min(g.ndisp - 1 + g.mindisp, 0);
    g.width1 = W - g.rofs - g.ndisp + 1;
    g.FILTERED = (g.mindisp - 1) * 16;
    // valid-disparity ROI (getValidDisparityROI with full-image ROIs) and the row whose overflow OpenCV leaves behind
    const int w2r = g.wsz / 2, maxDr = g.mindisp + g.ndisp - 1;
    const int xmin = std::max(0, maxDr) + w2r, xmax = W - w2r, ymin = w2r, ymax = H - w2r;
    const bool roi_ok = xmax > xmin && ymax > ymin;
    const int nspill = (g.mindisp > 0 && roi_ok && w2r >= 1 && ymax >= 1 && ymax < H) ? std::min(g.width1 - (W - g.lofs), W) : 0;
    g.yspill = nspill > 0 ? ymax - 1 : -1;
    const int npix = W * H;
    // everything starts FILTERED (left / right borders, rows the matcher does not reach)
    launch_fill16(disp, npix, int16_t(g.FILTERED), st);
    BM_LAUNCH_CHECK();
    if (!(g.lofs >= W || g.rofs >= W || g.width1 < 1)) {
        uint8_t *PL = pre, *PR = pre + size_t(npix);
        {
            dim3 block(256), grid((W + 255) / 256, H, 2);
            k_bm_prefilter<<<grid, block, 0, st>>>(dL, lp, dR, rp, W, H, g.cap, PL, PR);
            BM_LAUNCH_CHECK();
        }
        {
            const int tx = g.ndisp / 4;                                   // threads along d (ndisp is a multiple of 16)
            const int ty = std::max(1, 128 / tx);                         // rows per block
            const int nby = (H + ty - 1) / ty;
            int nseg = std::max(1, std::min(g.width1 / 64, (8 * num_sms * 16) / std::max(1, nby)));
            nseg = std::min(nseg, 64);
            const int seg_len = (g.width1 + nseg - 1) / nseg;
            dim3 block(tx, ty), grid(1, nby, (g.width1 + seg_len - 1) / seg_len);
            k_bm_hsad<<<grid, block, 0, st>>>(PL, PR, g, seg_len, HS);
            BM_LAUNCH_CHECK();
            dim3 gt((g.width1 + 255) / 256, H);
            k_bm_htext<<<gt, 256, 0, st>>>(PL, g, HT);
            BM_LAUNCH_CHECK();
        }
        {
            const int npl = (g.ndisp + 31) / 32;
            int nseg = std::max(1, std::min(H / 64, (16 * num_sms * 4) / std::max(1, g.width1)));
            const int seg_rows = (H + nseg - 1) / nseg;
            dim3 grid((g.width1 + 3) / 4, (H + seg_rows - 1) / seg_rows);
            const bool pack16 = (long long)(g.wsz + 1) * g.wsz * 2 * g.cap <= 65535 && g.ndisp <= 65535;
            if (pack16 && npl == 2) k_bm_match_p<2><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (pack16 && npl > 2 && npl <= 4) k_bm_match_p<4><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (pack16 && npl > 4 && npl <= 8) k_bm_match_p<8><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl == 2) k_bm_match_v<2><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl > 2 && npl <= 4) k_bm_match_v<4><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl > 4 && npl <= 8) k_bm_match_v<8><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl <= 1) k_bm_match<1><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl <= 2) k_bm_match<2><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl <= 4) k_bm_match<4><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl <= 8) k_bm_match<8><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else if (npl <= 16) k_bm_match<16><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            else k_bm_match<64><<<grid, 128, 0, st>>>(HS, HT, g, seg_rows, disp);
            BM_LAUNCH_CHECK();
        }
        {
            dim3 grid((W + 255) / 256, H);
            k_bm_mask<<<grid, 256, 0, st>>>(disp, W, H, xmin, xmax, ymin, ymax, g.FILTERED, std::max(nspill, 0));
            BM_LAUNCH_CHECK();
        }
    }
    if (p.speckleRange >= 0 && p.speckleWindowSize > 0)
        return launch_speckle(disp, label, parent, runlen, csize, W, H, g.FILTERED, p.speckleWindowSize, p.speckleRange, st, launches);
    return cudaSuccess;
}

}  // namespace b200sgm
