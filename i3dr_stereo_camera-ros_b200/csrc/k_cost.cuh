// Prefilter (A.2) and Birchfield-Tomasi cost + block sum (A.3, A.4) kernels.
// Replaces the first stages of cv::StereoSGBM::compute as called at
// /root/reference/src/stereoMatcher/matcherOpenCVSGBM.cpp:21 (algorithm: SURVEY.md Appendix A).
#pragma once
#include "sgm_types.h"

namespace b200sgm {

// ------------------------------------------------------------------------------------------------
// A.2: per pixel Sobel-x (clipped to [0, 2*ftzero]) and raw intensity (border columns forced to
// ftzero), each with its half-pixel min/max interval.  One thread per pixel; the three Sobel taps a
// pixel needs are recomputed (the kernel is O(W*H) and memory bound on a 5 MB image).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sobel_at(const uint8_t* __restrict__ r, const uint8_t* __restrict__ rn,
                                        const uint8_t* __restrict__ rs, int x, int W, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero;
    int g = (int(r[x + 1]) - int(r[x - 1])) * 2 + (int(rn[x + 1]) - int(rn[x - 1])) + (int(rs[x + 1]) - int(rs[x - 1]));
    return min(max(g, -ftzero), ftzero) + ftzero;
}

__global__ void k_prefilter(const uint8_t* __restrict__ img, size_t pitch, int W, int H, int ftzero,
                            Feat* __restrict__ feat)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W || y >= H) return;
    const uint8_t* r = img + size_t(y) * pitch;
    const uint8_t* rn = img + size_t(y > 0 ? y - 1 : y) * pitch;
    const uint8_t* rs = img + size_t(y < H - 1 ? y + 1 : y) * pitch;
    int s0 = sobel_at(r, rn, rs, x, W, ftzero);
    int sl = x > 0 ? sobel_at(r, rn, rs, x - 1, W, ftzero) : s0;
    int sr = x < W - 1 ? sobel_at(r, rn, rs, x + 1, W, ftzero) : s0;
    auto rawv = [&](int xx) { return (xx <= 0 || xx >= W - 1) ? ftzero : int(r[xx]); };
    int r0 = rawv(x);
    int rl = x > 0 ? rawv(x - 1) : r0;
    int rr = x < W - 1 ? rawv(x + 1) : r0;
    int sa = x > 0 ? (s0 + sl) >> 1 : s0, sb = x < W - 1 ? (s0 + sr) >> 1 : s0;
    int ra = x > 0 ? (r0 + rl) >> 1 : r0, rb = x < W - 1 ? (r0 + rr) >> 1 : r0;
    int slo = min(s0, min(sa, sb)), shi = max(s0, max(sa, sb));
    int rlo = min(r0, min(ra, rb)), rhi = max(r0, max(ra, rb));
    Feat f;
    f.x = uint32_t(s0) | (uint32_t(slo) << 8) | (uint32_t(shi) << 16) | (uint32_t(r0) << 24);
    f.y = uint32_t(rlo) | (uint32_t(rhi) << 8);
    feat[size_t(y) * W + x] = f;
}

// A.3 for one (left pixel, right pixel) pair.
__device__ __forceinline__ int bt_pixel_cost(Feat a, Feat b)
{
    int u = a.x & 0xFF, ulo = (a.x >> 8) & 0xFF, uhi = (a.x >> 16) & 0xFF;
    int v = b.x & 0xFF, vlo = (b.x >> 8) & 0xFF, vhi = (b.x >> 16) & 0xFF;
    int cs = min(max(0, max(u - vhi, vlo - u)), max(0, max(v - uhi, ulo - v)));
    u = a.x >> 24; ulo = a.y & 0xFF; uhi = (a.y >> 8) & 0xFF;
    v = b.x >> 24; vlo = b.y & 0xFF; vhi = (b.y >> 8) & 0xFF;
    int cr = min(max(0, max(u - vhi, vlo - u)), max(0, max(v - uhi, ulo - v)));
    return cs + (cr >> 2);
}

// ------------------------------------------------------------------------------------------------
// A.3 + A.4, generic version: one CTA owns TX valid columns x DCP disparity pairs and slides down a
// segment of rows.  Per entering row: pixel costs for TX+2*SW2 columns (x1-domain replicate clamp) go
// to shared memory, a sliding horizontal window produces hsum, and a ring of `bs` hsum rows in shared
// memory gives the vertical sliding sum C.  Output volume layout: C[y][x1][Dp] uint16, cells with
// disparity index >= D hold kMaxCost.
// dynamic smem: ((TX + 2*SW2) + bs*TX + TX) * DCP * 4 bytes.
// ------------------------------------------------------------------------------------------------
struct CostGeom {
    int W, H, W1, minX1, minD, D, Dp, SW2;
    int TX, DCP, RS;  // tile columns, disparity pairs per CTA, rows per segment
};

__global__ void __launch_bounds__(256) k_cost_generic(const Feat* __restrict__ fl, const Feat* __restrict__ fr,
                                                      uint16_t* __restrict__ Cvol, CostGeom g)
{
    extern __shared__ uint32_t smem[];
    const int bs = 2 * g.SW2 + 1;
    const int TXH = g.TX + 2 * g.SW2;
    uint32_t* pd = smem;                          // [TXH][DCP]
    uint32_t* ring = pd + TXH * g.DCP;            // [bs][TX][DCP]
    uint32_t* crun = ring + bs * g.TX * g.DCP;    // [TX][DCP]
    const int tx0 = blockIdx.x * g.TX;
    const int k0 = blockIdx.y * g.DCP * 2;        // first disparity index of this chunk
    const int ya = blockIdx.z * g.RS;
    const int yb = min(ya + g.RS, g.H);
    const int t = threadIdx.x;
    const int dp = t % g.DCP, grp = t / g.DCP, ngrp = 256 / g.DCP;
    const int cpg = (g.TX + ngrp - 1) / ngrp;     // columns per thread group
    const int c0 = grp * cpg, c1 = min(c0 + cpg, g.TX);

    for (int i = t; i < g.TX * g.DCP; i += 256) crun[i] = 0;

    const int nsteps = (yb - ya) + bs - 1;
    for (int s = 0; s < nsteps; s++) {
        const int e = min(max(ya - g.SW2 + s, 0), g.H - 1);
        const Feat* frow_l = fl + size_t(e) * g.W;
        const Feat* frow_r = fr + size_t(e) * g.W;
        __syncthreads();
        for (int i = t; i < TXH * g.DCP; i += 256) {
            int col = i / g.DCP, p = i - col * g.DCP;
            int x1 = min(max(tx0 - g.SW2 + col, 0), g.W1 - 1);
            int x = x1 + g.minX1;
            int k = k0 + 2 * p;
            uint32_t v = 0;
            if (k < g.D) {
                Feat a = __ldg(frow_l + x);
                int xr = x - (k + g.minD);
                v = uint32_t(bt_pixel_cost(a, __ldg(frow_r + xr)));
                if (k + 1 < g.D) v |= uint32_t(bt_pixel_cost(a, __ldg(frow_r + xr - 1))) << 16;
            }
            pd[i] = v;
        }
        __syncthreads();
        if (c0 < c1) {
            const int slot = s % bs;
            uint32_t hs = 0;
            for (int j = 0; j < bs; j++) hs += pd[(c0 + j) * g.DCP + dp];
            for (int c = c0; c < c1; c++) {
                if (c > c0) hs = hs + pd[(c + bs - 1) * g.DCP + dp] - pd[(c - 1) * g.DCP + dp];
                uint32_t* rp = ring + (slot * g.TX + c) * g.DCP + dp;
                uint32_t old = s >= bs ? *rp : 0u;
                *rp = hs;
                uint32_t cr = crun[c * g.DCP + dp] + hs - old;
                crun[c * g.DCP + dp] = cr;
                int y = ya + s - (bs - 1);
                int k = k0 + 2 * dp;
                if (y >= ya && tx0 + c < g.W1 && k < g.Dp) {
                    if (k >= g.D) cr = kMaxCostX2;
                    else if (k + 1 >= g.D) cr = (cr & 0xFFFFu) | (uint32_t(kMaxCost) << 16);
                    *reinterpret_cast<uint32_t*>(Cvol + (size_t(y) * g.W1 + tx0 + c) * g.Dp + k) = cr;
                }
            }
        }
    }
}

}  // namespace b200sgm
