// Row N4 of SURVEY.md section 8(f): cv::StereoBM::compute as called by the reference's MatcherOpenCVBlock
// (/root/reference/src/stereoMatcher/matcherOpenCVBlock.cpp:13-20; the package's default algorithm,
// launch/stereo_matcher.launch:20).  Algorithm as restated and pinned in oracle/bm_oracle.py: PREFILTER_XSOBEL, SAD over a
// blockSize x blockSize window with replicate-clamped columns and rows, texture threshold, uniqueness, x16 sub-pixel,
// valid-ROI mask, speckle filter.  First version: correct and measured, with the horizontal SAD volume materialised in the
// lane's cost-volume buffer (the SGBM kernels' C), not yet fused.
#pragma once
#include "sgm_types.h"

namespace b200sgm {

struct BmGeom {
    int W, H, ndisp, mindisp, wsz, cap, tex, uniq;
    int lofs, rofs, width1, FILTERED;
    // cv::StereoBM walks width1 columns from lofs, i.e. minDisparity pixels past the end of every row when minDisparity > 0; what
    // the LAST row of the valid ROI spills into the first pixels of the row below it survives OpenCV's ROI masking.  yspill is
    // that row (-1: none): its virtual columns are stored at the same linear address OpenCV writes them to.
    int yspill;
};

// prefilterXSobel; grid.z = 2 (left, right)
static __global__ void k_bm_prefilter(const uint8_t* __restrict__ imgL, size_t pitchL, const uint8_t* __restrict__ imgR, size_t pitchR,
                                      int W, int H, int cap, uint8_t* __restrict__ outL, uint8_t* __restrict__ outR)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* img = blockIdx.z ? imgR : imgL;
    const size_t pitch = blockIdx.z ? pitchR : pitchL;
    uint8_t* out = blockIdx.z ? outR : outL;
    int v = cap;
    const bool last_odd = (H & 1) && y == H - 1;          // rows are produced in pairs: the odd one out is all `cap`
    if (x > 0 && x < W - 1 && !last_odd) {
        const int up = y > 0 ? y - 1 : min(y + 1, H - 1), dn = y < H - 1 ? y + 1 : max(y - 1, 0);
        const uint8_t *r0 = img + size_t(up) * pitch, *r1 = img + size_t(y) * pitch, *r2 = img + size_t(dn) * pitch;
        const int g = (int(r0[x + 1]) - int(r0[x - 1])) + 2 * (int(r1[x + 1]) - int(r1[x - 1])) + (int(r2[x + 1]) - int(r2[x - 1]));
        v = min(max(g, -cap), cap) + cap;
    }
    out[size_t(y) * W + x] = uint8_t(v);
}

__device__ __forceinline__ int bm_lc(const BmGeom& g, int xx) { return min(max(xx, -g.lofs), g.W - g.lofs - 1) + g.lofs; }
__device__ __forceinline__ int bm_rc(const BmGeom& g, int xx) { return min(max(xx, -g.rofs), g.W - g.rofs - g.ndisp) + g.rofs; }

// Horizontal window sums: HS[y][x][d] = sum_{dx} |PL[y][lc(x+dx)] - PR[y][rc(x+dx) + d]| (uint16: <= 255 * 2*cap).  A thread owns FOUR
// consecutive disparity indices of one row and slides along a segment of columns: the four right-image bytes of a window column
// come from two aligned words + PRMT, the four absolute differences from one VABSDIFF4, and the four running sums live in two
// 16x2 registers (no carries between halves: every partial sum is a sum of non-negative terms below 2^16).
// block = (ndisp/4, rows per block); grid = (1, ceil(H / rows per block), column segments).
__device__ __forceinline__ uint32_t bm_absdiff4(const uint8_t* __restrict__ l, const uint8_t* __restrict__ r, const BmGeom& g, int xx)
{
    const uint32_t l4 = uint32_t(l[bm_lc(g, xx)]) * 0x01010101u;
    const uintptr_t ad = reinterpret_cast<uintptr_t>(r + bm_rc(g, xx));
    const uint32_t* wa = reinterpret_cast<const uint32_t*>(ad & ~uintptr_t(3));
    const uint32_t sh = uint32_t(ad & 3);
    const uint32_t w0 = __ldg(wa), w1 = sh ? __ldg(wa + 1) : 0u;
    return __vabsdiffu4(l4, __byte_perm(w0, w1, 0x3210u + 0x1111u * sh));
}

static __global__ void k_bm_hsad(const uint8_t* __restrict__ PL, const uint8_t* __restrict__ PR, BmGeom g, int seg_len,
                                 uint16_t* __restrict__ HS)
{
    const int d0 = threadIdx.x * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x0 = blockIdx.z * seg_len, x1 = min(x0 + seg_len, g.width1);
    if (d0 >= g.ndisp || y >= g.H || x0 >= x1) return;
    const uint8_t* l = PL + size_t(y) * g.W;
    const uint8_t* r = PR + size_t(y) * g.W + d0;
    const int w2 = g.wsz >> 1;
    uint32_t hA = 0, hB = 0;           // (hs[d0], hs[d0+1]), (hs[d0+2], hs[d0+3])
    for (int dx = -w2; dx <= w2; dx++) {
        const uint32_t a4 = bm_absdiff4(l, r, g, x0 + dx);
        hA += __byte_perm(a4, 0, 0x4140); hB += __byte_perm(a4, 0, 0x4342);
    }
    uint2* o = reinterpret_cast<uint2*>(HS + (size_t(y) * g.width1 + x0) * g.ndisp + d0);
    const size_t ostep = size_t(g.ndisp) / 4;
    for (int x = x0; x < x1; x++, o += ostep) {
        *o = make_uint2(hA, hB);
        const uint32_t a4 = bm_absdiff4(l, r, g, x + w2 + 1), r4 = bm_absdiff4(l, r, g, x - w2);
        hA = hA + __byte_perm(a4, 0, 0x4140) - __byte_perm(r4, 0, 0x4140);
        hB = hB + __byte_perm(a4, 0, 0x4342) - __byte_perm(r4, 0, 0x4342);
    }
}

// Horizontal texture sums HT[y][x] = sum_{dx} |PL[y][lc(x+dx)] - cap|
static __global__ void k_bm_htext(const uint8_t* __restrict__ PL, BmGeom g, int* __restrict__ HT)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= g.width1) return;
    const uint8_t* l = PL + size_t(y) * g.W;
    const int w2 = g.wsz >> 1;
    int s = 0;
    for (int dx = -w2; dx <= w2; dx++) s += abs(int(l[bm_lc(g, x + dx)]) - g.cap);
    HT[size_t(y) * g.width1 + x] = s;
}

// Vertical window sums + winner-take-all.  One warp per (column x, row segment); lane l holds disparity indices l, l+32, ...
// (NPL per lane).  d index -> disparity = ndisp - 1 - d + mindisp; the FIRST minimum over d wins (= the largest disparity).
template <int NPL>
static __global__ void __launch_bounds__(128) k_bm_match(const uint16_t* __restrict__ HS, const int* __restrict__ HT, BmGeom g,
                                                         int seg_rows, int16_t* __restrict__ disp)
{
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (x >= g.width1) return;
    const int ya = blockIdx.y * seg_rows, yb = min(ya + seg_rows, g.H);
    if (ya >= yb) return;
    const int w2 = g.wsz >> 1, nd = g.ndisp;
    const size_t rstride = size_t(g.width1) * nd;
    const uint16_t* hs = HS + size_t(x) * nd;
    auto cy = [&](int yy) { return min(max(yy, 0), g.H - 1); };
    int sad[NPL];
#pragma unroll
    for (int k = 0; k < NPL; k++) sad[k] = 0;
    int tsum = 0;
    for (int r = ya - w2; r <= ya + w2; r++) {
        const uint16_t* p = hs + size_t(cy(r)) * rstride;
#pragma unroll
        for (int k = 0; k < NPL; k++) { const int d = lane + 32 * k; if (d < nd) sad[k] += p[d]; }
        tsum += HT[size_t(cy(r)) * g.width1 + x];
    }
    const bool in_row = g.lofs + x < g.W;
    for (int y = ya; y < yb; y++) {
        // first minimum
        unsigned best = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < NPL; k++) { const int d = lane + 32 * k; if (d < nd) best = min(best, unsigned(sad[k])); }
        const unsigned minsad = __reduce_min_sync(0xFFFFFFFFu, best);
        unsigned dc = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < NPL; k++) { const int d = lane + 32 * k; if (d < nd && unsigned(sad[k]) == minsad) dc = min(dc, unsigned(d)); }
        const int mind = int(__reduce_min_sync(0xFFFFFFFFu, dc));
        bool ok = tsum >= g.tex;
        if (g.uniq > 0) {
            const long long thresh = (long long)minsad + ((long long)minsad * g.uniq / 100);
            bool viol = false;
#pragma unroll
            for (int k = 0; k < NPL; k++) {
                const int d = lane + 32 * k;
                if (d < nd && (d < mind - 1 || d > mind + 1) && (long long)sad[k] <= thresh) viol = true;
            }
            ok = ok && !__any_sync(0xFFFFFFFFu, viol);
        }
        // neighbours of the minimum (sad[-1] = sad[1], sad[ndisp] = sad[ndisp-2])
        const int jm = mind == 0 ? 1 : mind - 1, jp = mind == nd - 1 ? nd - 2 : mind + 1;
        unsigned vn = 0, vp = 0;
#pragma unroll
        for (int k = 0; k < NPL; k++) {
            const int d = lane + 32 * k;
            if (d == jm) vn = unsigned(sad[k]);
            if (d == jp) vp = unsigned(sad[k]);
        }
        const int n = int(__reduce_max_sync(0xFFFFFFFFu, vn)), p = int(__reduce_max_sync(0xFFFFFFFFu, vp));
        if (lane == 0 && (in_row || y == g.yspill)) {
            int out = g.FILTERED;
            if (ok) {
                const int den = p + n - 2 * int(minsad) + abs(p - n);
                const int q = den != 0 ? (p - n) * 256 / den : 0;            // C division: toward zero
                out = ((nd - mind - 1 + g.mindisp) * 256 + q + 15) >> 4;
            }
            disp[size_t(y) * g.W + g.lofs + x] = int16_t(out);
        }
        if (y + 1 < yb) {
            const uint16_t* pa = hs + size_t(cy(y + w2 + 1)) * rstride;
            const uint16_t* pr = hs + size_t(cy(y - w2)) * rstride;
#pragma unroll
            for (int k = 0; k < NPL; k++) { const int d = lane + 32 * k; if (d < nd) sad[k] += int(pa[d]) - int(pr[d]); }
            tsum += HT[size_t(cy(y + w2 + 1)) * g.width1 + x] - HT[size_t(cy(y - w2)) * g.width1 + x];
        }
    }
}

// Same as k_bm_match with the lane <-> disparity mapping turned around for vector loads: lane l holds the NPL CONSECUTIVE indices
// l*NPL .. l*NPL+NPL-1 (one 2*NPL-byte load per window row instead of NPL 2-byte loads); ndisp is a multiple of 16, so a lane is
// either entirely inside [0, ndisp) or entirely outside.  NPL in {2, 4, 8}.
template <int NPL>
static __global__ void __launch_bounds__(128) k_bm_match_v(const uint16_t* __restrict__ HS, const int* __restrict__ HT, BmGeom g,
                                                           int seg_rows, int16_t* __restrict__ disp)
{
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (x >= g.width1) return;
    const int ya = blockIdx.y * seg_rows, yb = min(ya + seg_rows, g.H);
    if (ya >= yb) return;
    const int w2 = g.wsz >> 1, nd = g.ndisp;
    const size_t rstride = size_t(g.width1) * nd;
    const int dbase = lane * NPL;
    const bool act = dbase < nd;
    const uint16_t* hs = HS + size_t(x) * nd + (act ? dbase : 0);
    auto cy = [&](int yy) { return min(max(yy, 0), g.H - 1); };
    auto ldrow = [&](int yy, uint32_t (&w)[NPL / 2]) {
        const uint16_t* p = hs + size_t(cy(yy)) * rstride;
        if constexpr (NPL == 8) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
        else if constexpr (NPL == 4) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); w[0] = v.x; w[1] = v.y; }
        else { w[0] = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    };
    int sad[NPL];
#pragma unroll
    for (int k = 0; k < NPL; k++) sad[k] = act ? 0 : 0x7FFFFFFF;
    int tsum = 0;
    for (int r = ya - w2; r <= ya + w2; r++) {
        uint32_t w[NPL / 2];
        ldrow(r, w);
        if (act) {
#pragma unroll
            for (int k = 0; k < NPL / 2; k++) { sad[2 * k] += int(w[k] & 0xFFFFu); sad[2 * k + 1] += int(w[k] >> 16); }
        }
        tsum += HT[size_t(cy(r)) * g.width1 + x];
    }
    const bool in_row = g.lofs + x < g.W;
    for (int y = ya; y < yb; y++) {
        // prefetch the two rows of the slide
        uint32_t wa[NPL / 2], wr[NPL / 2];
        int ta = 0, tr = 0;
        const bool more = y + 1 < yb;
        if (more) {
            ldrow(y + w2 + 1, wa); ldrow(y - w2, wr);
            ta = HT[size_t(cy(y + w2 + 1)) * g.width1 + x]; tr = HT[size_t(cy(y - w2)) * g.width1 + x];
        }
        int m = sad[0];
#pragma unroll
        for (int k = 1; k < NPL; k++) m = min(m, sad[k]);
        const int minsad = __reduce_min_sync(0xFFFFFFFFu, m);
        int dc = 0x7FFFFFFF;
#pragma unroll
        for (int k = NPL - 1; k >= 0; k--) dc = sad[k] == minsad ? dbase + k : dc;      // first index inside the lane
        const int mind = __reduce_min_sync(0xFFFFFFFFu, dc);
        bool ok = tsum >= g.tex;
        if (g.uniq > 0) {
            // minsad <= 255 * 255 * 126 < 2^23: the product fits 32 bits for uniquenessRatio <= 500
            const long long t64 = g.uniq <= 500 ? (long long)(uint32_t(minsad) + uint32_t(minsad) * uint32_t(g.uniq) / 100u)
                                                : (long long)minsad + ((long long)minsad * g.uniq / 100);
            const int thresh = int(min(t64, (long long)0x7FFFFFFE));
            bool viol = false;
#pragma unroll
            for (int k = 0; k < NPL; k++) {
                const int d = dbase + k;
                viol = viol || (sad[k] <= thresh && (d < mind - 1 || d > mind + 1));
            }
            ok = ok && !__any_sync(0xFFFFFFFFu, viol);
        }
        const int jm = mind == 0 ? 1 : mind - 1, jp = mind == nd - 1 ? nd - 2 : mind + 1;
        int vn = 0, vp = 0;
#pragma unroll
        for (int k = 0; k < NPL; k++) {
            if (k == jm % NPL) vn = sad[k];
            if (k == jp % NPL) vp = sad[k];
        }
        const int n = __shfl_sync(0xFFFFFFFFu, vn, jm / NPL), p = __shfl_sync(0xFFFFFFFFu, vp, jp / NPL);
        if (lane == 0 && (in_row || y == g.yspill)) {
            int out = g.FILTERED;
            if (ok) {
                const int den = p + n - 2 * minsad + abs(p - n);
                const int q = den != 0 ? (p - n) * 256 / den : 0;            // C division: toward zero
                out = ((nd - mind - 1 + g.mindisp) * 256 + q + 15) >> 4;
            }
            disp[size_t(y) * g.W + g.lofs + x] = int16_t(out);
        }
        if (more) {
            if (act) {
#pragma unroll
                for (int k = 0; k < NPL / 2; k++) {
                    sad[2 * k] += int(wa[k] & 0xFFFFu) - int(wr[k] & 0xFFFFu);
                    sad[2 * k + 1] += int(wa[k] >> 16) - int(wr[k] >> 16);
                }
            }
            tsum += ta - tr;
        }
    }
}

// Packed variant of k_bm_match_v for window sums that fit 16 bits ((blockSize + 1) * blockSize * 2 * preFilterCap <= 65535, i.e.
// every window the reference's launch files use): the NPL sums of a lane live in NPL/2 16x2 registers (slide = two packed
// adds), and the first minimum comes out of ONE warp reduction over keys (sum << 16 | index).  NPL in {2, 4, 8}.
template <int NPL>
static __global__ void __launch_bounds__(128) k_bm_match_p(const uint16_t* __restrict__ HS, const int* __restrict__ HT, BmGeom g,
                                                           int seg_rows, int16_t* __restrict__ disp)
{
    constexpr int NW = NPL / 2;
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (x >= g.width1) return;
    const int ya = blockIdx.y * seg_rows, yb = min(ya + seg_rows, g.H);
    if (ya >= yb) return;
    const int w2 = g.wsz >> 1, nd = g.ndisp;
    const size_t rstride = size_t(g.width1) * nd;
    const int dbase = lane * NPL;
    const bool act = dbase < nd;
    const uint16_t* hs = HS + size_t(x) * nd + (act ? dbase : 0);
    auto cy = [&](int yy) { return min(max(yy, 0), g.H - 1); };
    auto ldrow = [&](int yy, uint32_t (&w)[NW]) {
        const uint16_t* p = hs + size_t(cy(yy)) * rstride;
        if constexpr (NPL == 8) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
        else if constexpr (NPL == 4) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); w[0] = v.x; w[1] = v.y; }
        else { w[0] = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    };
    uint32_t s2[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) s2[k] = 0;
    int tsum = 0;
    for (int r = ya - w2; r <= ya + w2; r++) {
        uint32_t w[NW];
        ldrow(r, w);
#pragma unroll
        for (int k = 0; k < NW; k++) s2[k] += w[k];
        tsum += HT[size_t(cy(r)) * g.width1 + x];
    }
    if (!act) {
#pragma unroll
        for (int k = 0; k < NW; k++) s2[k] = 0xFFFFFFFFu;
    }
    const bool in_row = g.lofs + x < g.W;
    for (int y = ya; y < yb; y++) {
        uint32_t wa[NW], wr[NW];
        int ta = 0, tr = 0;
        const bool more = y + 1 < yb;
        if (more) {
            ldrow(y + w2 + 1, wa); ldrow(y - w2, wr);
            ta = HT[size_t(cy(y + w2 + 1)) * g.width1 + x]; tr = HT[size_t(cy(y - w2)) * g.width1 + x];
        }
        // key = sum << 16 | index: the minimum key is the minimum sum and, among equals, the first index
        uint32_t key = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const uint32_t klo = (s2[k] << 16) | uint32_t(dbase + 2 * k), khi = (s2[k] & 0xFFFF0000u) | uint32_t(dbase + 2 * k + 1);
            key = min(key, min(klo, khi));
        }
        key = __reduce_min_sync(0xFFFFFFFFu, key);
        const int minsad = int(key >> 16), mind = int(key & 0xFFFFu);
        bool ok = tsum >= g.tex;
        if (g.uniq > 0) {
            // minsad < 2^16: the product fits 32 bits for any uniquenessRatio below 2^16 (unsigned division by a constant)
            const long long t64 = g.uniq < 65536 ? (long long)(uint32_t(minsad) + uint32_t(minsad) * uint32_t(g.uniq) / 100u)
                                                 : (long long)minsad + ((long long)minsad * g.uniq / 100);
            const uint32_t thr2 = uint32_t(min(t64, (long long)0xFFFE)) * 0x10001u;
            uint32_t viol = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                uint32_t le = __vcmpleu2(s2[k], thr2);                      // 0xFFFF per half with sum <= thresh
                const int d0 = dbase + 2 * k;
                if (unsigned(d0 - mind + 1) < 3u) le &= 0xFFFF0000u;        // indices mind-1 .. mind+1 are exempt
                if (unsigned(d0 + 1 - mind + 1) < 3u) le &= 0x0000FFFFu;
                viol |= le;
            }
            ok = ok && !__any_sync(0xFFFFFFFFu, viol != 0);
        }
        const int jm = mind == 0 ? 1 : mind - 1, jp = mind == nd - 1 ? nd - 2 : mind + 1;
        uint32_t wn = 0, wp = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            if (k == (jm % NPL) / 2) wn = s2[k];
            if (k == (jp % NPL) / 2) wp = s2[k];
        }
        wn = __shfl_sync(0xFFFFFFFFu, wn, jm / NPL); wp = __shfl_sync(0xFFFFFFFFu, wp, jp / NPL);
        const int n = int((jm & 1) ? (wn >> 16) : (wn & 0xFFFFu)), p = int((jp & 1) ? (wp >> 16) : (wp & 0xFFFFu));
        if (lane == 0 && (in_row || y == g.yspill)) {
            int out = g.FILTERED;
            if (ok) {
                const int den = p + n - 2 * minsad + abs(p - n);
                const int q = den != 0 ? (p - n) * 256 / den : 0;            // C division: toward zero
                out = ((nd - mind - 1 + g.mindisp) * 256 + q + 15) >> 4;
            }
            disp[size_t(y) * g.W + g.lofs + x] = int16_t(out);
        }
        if (more) {
            if (act) {
#pragma unroll
                for (int k = 0; k < NW; k++) s2[k] = s2[k] + wa[k] - wr[k];
            }
            tsum += ta - tr;
        }
    }
}

// getValidDisparityROI with full-image ROIs: everything outside [xmin, xmax) x [ymin, ymax) is FILTERED
// nspill: pixels at the start of row ymax that hold the spill of row ymax - 1 (BmGeom::yspill) and are left alone
static __global__ void k_bm_mask(int16_t* __restrict__ disp, int W, int H, int xmin, int xmax, int ymin, int ymax, int FILTERED, int nspill)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    if (y == ymax && x < nspill) return;
    if (x < xmin || x >= xmax || y < ymin || y >= ymax) disp[size_t(y) * W + x] = int16_t(FILTERED);
}

}  // namespace b200sgm
