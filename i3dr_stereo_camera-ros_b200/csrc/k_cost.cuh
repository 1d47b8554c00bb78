// Prefilter (A.2) and Birchfield-Tomasi cost + block sum (A.3, A.4) kernels.
// Replaces the first stages of cv::StereoSGBM::compute as called at
// /root/reference/src/stereoMatcher/matcherOpenCVSGBM.cpp:21 (algorithm: SURVEY.md Appendix A).
#pragma once
#include <type_traits>
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

__device__ __forceinline__ uint32_t pk16(int a, int b) { return (uint32_t(a) & 0xFFFFu) | (uint32_t(b) << 16); }

// grid.z = 2: z = 0 left image, z = 1 right image.  One thread produces kPfPPT consecutive pixels: the Sobel values and
// the byte loads they need are shared between neighbours (6 loads per pixel instead of 21).
constexpr int kPfPPT = 4;
__global__ void __launch_bounds__(128) k_prefilter(const uint8_t* __restrict__ imgL, size_t pitchL, const uint8_t* __restrict__ imgR,
                                                   size_t pitchR, int W, int H, int ftzero, Feat* __restrict__ featL,
                                                   Feat* __restrict__ featR)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * kPfPPT;
    const int y = blockIdx.y;
    if (x0 >= W || y >= H) return;
    const uint8_t* img = blockIdx.z ? imgR : imgL;
    const size_t pitch = blockIdx.z ? pitchR : pitchL;
    Feat* feat = (blockIdx.z ? featR : featL) + size_t(y) * W;
    const uint8_t* r = img + size_t(y) * pitch;
    const uint8_t* rn = img + size_t(y > 0 ? y - 1 : y) * pitch;
    const uint8_t* rs = img + size_t(y < H - 1 ? y + 1 : y) * pitch;
    // pixels x0-2 .. x0+kPfPPT+1 of the three rows (clamped addresses; out-of-image values are never used)
    int pc[kPfPPT + 4], pv[kPfPPT + 4];     // centre row; vertical [1 2 1] column sums
#pragma unroll
    for (int i = 0; i < kPfPPT + 4; i++) {
        const int xx = min(max(x0 - 2 + i, 0), W - 1);
        pc[i] = __ldg(r + xx);
        pv[i] = 2 * pc[i] + int(__ldg(rn + xx)) + int(__ldg(rs + xx));
    }
    // Sobel and raw values of pixels x0-1 .. x0+kPfPPT (index j <-> pixel x0-1+j)
    int sv[kPfPPT + 2], rv[kPfPPT + 2];
#pragma unroll
    for (int j = 0; j < kPfPPT + 2; j++) {
        const int xx = x0 - 1 + j;
        const bool border = xx <= 0 || xx >= W - 1;
        sv[j] = border ? ftzero : min(max(pv[j + 2] - pv[j], -ftzero), ftzero) + ftzero;
        rv[j] = border ? ftzero : pc[j + 1];
    }
#pragma unroll
    for (int k = 0; k < kPfPPT; k++) {
        const int x = x0 + k;
        if (x < W) {
            const int s0 = sv[k + 1], r0 = rv[k + 1];
            const int sa = x > 0 ? (s0 + sv[k]) >> 1 : s0, sb = x < W - 1 ? (s0 + sv[k + 2]) >> 1 : s0;
            const int ra = x > 0 ? (r0 + rv[k]) >> 1 : r0, rb = x < W - 1 ? (r0 + rv[k + 2]) >> 1 : r0;
            const int slo = min(s0, min(sa, sb)), shi = max(s0, max(sa, sb));
            const int rlo = min(r0, min(ra, rb)), rhi = max(r0, max(ra, rb));
            feat[x] = make_uint4(pk16(s0, slo), pk16(-shi, -s0), pk16(r0 * 64, rlo * 64), pk16(-rhi * 64, -r0 * 64));
        }
    }
}

// A.3 for one (left pixel, right pixel) pair.
__device__ __forceinline__ int bt_pixel_cost(Feat a, Feat b)
{
    auto lo16 = [](uint32_t w) { return int(int16_t(w & 0xFFFFu)); };
    auto hi16 = [](uint32_t w) { return int(int16_t(w >> 16)); };
    int u = lo16(a.x), ulo = hi16(a.x), uhi = -lo16(a.y);
    int v = lo16(b.x), vlo = hi16(b.x), vhi = -lo16(b.y);
    int cs = min(max(0, max(u - vhi, vlo - u)), max(0, max(v - uhi, ulo - v)));
    u = lo16(a.z) >> 6; ulo = hi16(a.z) >> 6; uhi = (-lo16(a.w)) >> 6;
    v = lo16(b.z) >> 6; vlo = hi16(b.z) >> 6; vhi = (-lo16(b.w)) >> 6;
    int cr = min(max(0, max(u - vhi, vlo - u)), max(0, max(v - uhi, ulo - v)));
    return cs + (cr >> 2);
}

// ------------------------------------------------------------------------------------------------
// A.3 + A.4, generic version: one CTA owns TX valid columns x DCP disparity pairs and slides down a
// segment of rows.  Per entering row: pixel costs for TX+2*SW2 columns (x1-domain replicate clamp) go
// to shared memory, a sliding horizontal window produces hsum, and a ring of `bs` hsum rows in shared
// memory gives the vertical sliding sum C.  Output volume layout: C[y][x1][Dp/2] uint32 words in the paired
// layout of k_path.cuh (word w = cell w | cell Dh+w << 16); cells with disparity index >= D hold kMaxCost.
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
    const int k0 = blockIdx.y * g.DCP;            // first word (= low-half disparity index) of this chunk
    const int Dh = g.Dp >> 1;
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
            int k = k0 + p;
            uint32_t v = 0;
            if (k < g.D && k < Dh) {
                Feat a = __ldg(frow_l + x);
                int xr = x - (k + g.minD);
                v = uint32_t(bt_pixel_cost(a, __ldg(frow_r + xr)));
                if (k + Dh < g.D) v |= uint32_t(bt_pixel_cost(a, __ldg(frow_r + xr - Dh))) << 16;
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
                int k = k0 + dp;
                if (y >= ya && tx0 + c < g.W1 && k < Dh) {
                    if (k >= g.D) cr = kMaxCostX2;
                    else if (k + Dh >= g.D) cr = (cr & 0xFFFFu) | (uint32_t(kMaxCost) << 16);
                    *reinterpret_cast<uint32_t*>(Cvol + (size_t(y) * g.W1 + tx0 + c) * g.Dp + 2 * k) = cr;
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// A.3 + A.4, fast version (blockSize <= 21).  One CTA (256 threads) owns a tile of 64 x1-columns (56 output
// columns + the 2*SW2 halo for blockSize 9) x 32 disparity pairs and slides down a segment of rows.
//
// Per entering row:
//   stage   : the row's prefiltered pixels are repacked into shared memory as signed 16x2 operands:
//             right image: for every xr the pair (f[xr], f[xr-Dh]) that a disparity pair (d, d+Dh) of the
//             paired volume layout (k_path.cuh) needs;
//             left image: every field replicated into both halves.  The raw channel is pre-scaled by 64
//             so that (cost >> 2) is the high byte of each half (one PRMT).
//   phase 1 : lanes <-> columns.  Birchfield-Tomasi cost of a disparity pair = 2 x {VIADD.16x2,
//             VIADDMNMX.S16x2.RELU, VIADD.16x2, VIADDMNMX.S16x2.RELU, VIMNMX} + PRMT + IADD; the vertical
//             sliding sum lives in registers, the ring of the last `bs` rows in shared memory.
//   phase 2 : lanes <-> disparity pairs.  Sliding horizontal sum over the vertical sums, coalesced
//             128-byte stores of C[y][x1][k0 .. k0+64).
// ------------------------------------------------------------------------------------------------
struct CostFastGeom {
    int W, H, W1, minX1, minD, D, Dp, SW2, RS;
};

constexpr int kCfTXH = 64;    // tile columns incl. halo
constexpr int kCfDCP = 32;    // disparity pairs per CTA
constexpr int kCfNRP = 128;   // right-image pair records per row (>= TXH + DCP - 1)
constexpr int kCfPPT = 8;     // disparity pairs per thread in phase 1

inline size_t cost_fast_smem(int SW2, bool ring8)
{
    const int bs = 2 * SW2 + 1;
    return size_t(2) * (2 * kCfNRP + 2 * kCfTXH) * sizeof(uint4) + size_t(kCfDCP) * (kCfTXH + 1) * 4 +
           size_t(bs) * kCfDCP * kCfTXH * (ring8 ? 2 : 4);
}

// SW2T > 0: blockSize/2 known at compile time (loops unrolled); SW2T == 0: taken from the geometry.
// RING8  : pixel costs fit a byte (2*ftzero + 63 <= 255): the ring of the last `bs` rows holds a disparity pair in 16
//          bits, which halves the dominant shared-memory array (3 CTAs per SM instead of 2)
// NOPAD  : Dp == D and every lane owns a word: no padded cells to force to kMaxCost
template <int SW2T, bool RING8, bool NOPAD>
__global__ void __launch_bounds__(256, RING8 ? 3 : 2) k_cost_fast(const Feat* __restrict__ fl, const Feat* __restrict__ fr,
                                                      uint16_t* __restrict__ Cvol, CostFastGeom g)
{
    extern __shared__ uint4 cf_smem[];
    const int SW2 = SW2T > 0 ? SW2T : g.SW2;
    const int bs = 2 * SW2 + 1;
    const int TX = kCfTXH - 2 * SW2;
    uint4* Rs = cf_smem;                       // [2][NRP]  sobel: (v, lo, -hi, -v) pairs
    uint4* Rr = Rs + 2 * kCfNRP;               // [2][NRP]  raw x64
    uint4* Ls = Rr + 2 * kCfNRP;               // [2][TXH]  sobel: (u, -u, lo, -hi) replicated
    uint4* Lr = Ls + 2 * kCfTXH;               // [2][TXH]  raw x64
    uint32_t* vs = reinterpret_cast<uint32_t*>(Lr + 2 * kCfTXH);   // [DCP][TXH+1]
    using ring_t = typename std::conditional<RING8, uint16_t, uint32_t>::type;
    ring_t* ring = reinterpret_cast<ring_t*>(vs + kCfDCP * (kCfTXH + 1));   // [bs][DCP][TXH]

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int tx0 = blockIdx.x * TX;
    const int k0 = blockIdx.y * kCfDCP;          // first word of this chunk: cells k0+p and Dh+k0+p
    const int Dh = g.Dp >> 1;
    const int ya = blockIdx.z * g.RS, yb = min(ya + g.RS, g.H);
    auto col_x = [&](int c) { return min(max(tx0 - SW2 + c, 0), g.W1 - 1) + g.minX1; };
    const int xr_min = col_x(0) - g.minD - k0 - (kCfDCP - 1);
    const int nR = col_x(kCfTXH - 1) - col_x(0) + (kCfDCP - 1) + 1;

    // phase-1 role
    const int col = (w & 1) * 32 + lane, pg = w >> 1;
    const int rbase = col_x(col) - g.minD - k0 - xr_min;   // record index of pair p: rbase - p
    uint32_t V[kCfPPT];
#pragma unroll
    for (int i = 0; i < kCfPPT; i++) V[i] = 0;
    for (int i = t; i < bs * kCfDCP * kCfTXH; i += 256) ring[i] = 0;

    auto stage = [&](int s, int buf) {
        const int e = min(max(ya - SW2 + s, 0), g.H - 1);
        if (t < nR) {
            const int xr = xr_min + t;
            const Feat a = __ldg(fr + size_t(e) * g.W + min(max(xr, 0), g.W - 1));
            const Feat b = __ldg(fr + size_t(e) * g.W + min(max(xr - Dh, 0), g.W - 1));
            // (v, lo, -hi, -v) of the pixel pair (xr, xr - Dh), low half = xr
            Rs[buf * kCfNRP + t] = make_uint4(__byte_perm(a.x, b.x, 0x5410), __byte_perm(a.x, b.x, 0x7632),
                                              __byte_perm(a.y, b.y, 0x5410), __byte_perm(a.y, b.y, 0x7632));
            Rr[buf * kCfNRP + t] = make_uint4(__byte_perm(a.z, b.z, 0x5410), __byte_perm(a.z, b.z, 0x7632),
                                              __byte_perm(a.w, b.w, 0x5410), __byte_perm(a.w, b.w, 0x7632));
        } else if (t >= 128 && t < 128 + kCfTXH) {
            const int c = t - 128;
            const Feat a = __ldg(fl + size_t(e) * g.W + col_x(c));
            // (u, -u, lo, -hi), each replicated in both halves
            Ls[buf * kCfTXH + c] = make_uint4(__byte_perm(a.x, a.x, 0x1010), __byte_perm(a.y, a.y, 0x3232),
                                              __byte_perm(a.x, a.x, 0x3232), __byte_perm(a.y, a.y, 0x1010));
            Lr[buf * kCfTXH + c] = make_uint4(__byte_perm(a.z, a.z, 0x1010), __byte_perm(a.w, a.w, 0x3232),
                                              __byte_perm(a.z, a.z, 0x3232), __byte_perm(a.w, a.w, 0x1010));
        }
    };

    stage(0, 0);
    __syncthreads();
    const int nsteps = (yb - ya) + bs - 1;
    // phase-2 role: lanes <-> pairs, warp -> run of output columns
    const int cpw = (TX + 7) / 8;
    const int c_lo = w * cpw, c_hi = min(c_lo + cpw, TX);
    const int kk = k0 + lane;                                  // this lane's word: cells kk and Dh + kk
    const uint32_t pad_or = kk >= g.D ? kMaxCostX2 : (kk + Dh >= g.D ? (uint32_t(kMaxCost) << 16) : 0u);
    const uint32_t pad_and = kk >= g.D ? 0u : (kk + Dh >= g.D ? 0x0000FFFFu : 0xFFFFFFFFu);

    // phase-2 output pointer of this lane: row y, column tx0 + c_lo, disparity pair kk
    const bool lane_ok = kk < Dh;
    const size_t colBytes = size_t(g.Dp) * 2;
    int slot = 0;
    for (int s = 0; s < nsteps; s++) {
        const int buf = s & 1;
        // ---- phase 1
        {
            const uint4 ls = Ls[buf * kCfTXH + col], lr = Lr[buf * kCfTXH + col];
            ring_t* rrow = ring + size_t(slot) * kCfDCP * kCfTXH + col;
            slot = slot + 1 == bs ? 0 : slot + 1;
            const uint4* rs_p = Rs + buf * kCfNRP + rbase;
            const uint4* rr_p = Rr + buf * kCfNRP + rbase;
#pragma unroll
            for (int i = 0; i < kCfPPT; i++) {
                const int p = pg + 4 * i;
                const uint4 rs = rs_p[-p], rr = rr_p[-p];
                uint32_t X = __vadd2(rs.y, ls.y);
                uint32_t c0 = __viaddmax_s16x2_relu(ls.x, rs.z, X);
                uint32_t Y = __vadd2(ls.z, rs.w);
                uint32_t c1 = __viaddmax_s16x2_relu(rs.x, ls.w, Y);
                const uint32_t cs = __vmins2(c0, c1);
                X = __vadd2(rr.y, lr.y);
                c0 = __viaddmax_s16x2_relu(lr.x, rr.z, X);
                Y = __vadd2(lr.z, rr.w);
                c1 = __viaddmax_s16x2_relu(rr.x, lr.w, Y);
                const uint32_t cr = __vmins2(c0, c1);
                const uint32_t pd = cs + __byte_perm(cr, 0, 0x4341);   // + (cost_raw >> 2)
                uint32_t old;
                if (RING8) {
                    old = __byte_perm(uint32_t(rrow[p * kCfTXH]), 0, 0x4140);
                    rrow[p * kCfTXH] = ring_t(__byte_perm(pd, 0, 0x4420));
                } else {
                    old = rrow[p * kCfTXH];
                    rrow[p * kCfTXH] = ring_t(pd);
                }
                V[i] = V[i] + pd - old;
                vs[p * (kCfTXH + 1) + col] = V[i];
            }
        }
        __syncthreads();
        // ---- phase 2 (once the vertical window is full) + staging of the next row
        if (s >= bs - 1 && c_lo < c_hi && lane_ok) {
            const int y = ya + s - (bs - 1);
            const uint32_t* vrow = vs + lane * (kCfTXH + 1) + c_lo;
            char* out = reinterpret_cast<char*>(Cvol + (size_t(y) * g.W1 + tx0 + c_lo) * g.Dp + 2 * kk);
            const int ncol = min(c_hi, g.W1 - tx0) - c_lo;     // columns of this run that exist in the image
            uint32_t hs = 0;
            if (SW2T > 0) {
                constexpr int BS = 2 * SW2T + 1, CPW = (kCfTXH - 2 * SW2T + 7) / 8;
                uint32_t v[CPW + BS - 1];
#pragma unroll
                for (int jj = 0; jj < CPW + BS - 1; jj++) v[jj] = vrow[jj];   // may read a few words of the next run: harmless
#pragma unroll
                for (int jj = 0; jj < BS; jj++) hs += v[jj];
#pragma unroll
                for (int c = 0; c < CPW; c++) {
                    if (c > 0) hs = hs + v[c + BS - 1] - v[c - 1];
                    if (c < ncol) *reinterpret_cast<uint32_t*>(out + c * colBytes) = NOPAD ? hs : ((hs & pad_and) | pad_or);
                }
            } else {
                for (int jj = 0; jj < bs; jj++) hs += vrow[jj];
                for (int c = 0; c < ncol; c++) {
                    if (c > 0) hs = hs + vrow[c + bs - 1] - vrow[c - 1];
                    *reinterpret_cast<uint32_t*>(out + c * colBytes) = NOPAD ? hs : ((hs & pad_and) | pad_or);
                }
            }
        }
        if (s + 1 < nsteps) stage(s + 1, buf ^ 1);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// A.3 + A.4, register-tiled version (blockSize <= 21, pixel costs that fit a byte).  One CTA (256 threads) owns a
// tile of 128 x1-columns (128 - 2*SW2 output columns + halo) x 32 disparity pairs and slides down a segment of rows.
// Phase 1: a thread owns TWO adjacent columns x EIGHT consecutive disparity pairs.  The right-image record of
// (column c, pair p) is record c + 31 - p of the row, so the 16 pixel costs of a thread need only 9 right records
// (a rolling pair of registers) + 2 left records: 22 LDS.128 instead of 36 for the same work -- the shared-memory
// pipe is what bounds this kernel.  Records are stored de-interleaved by parity so that the stride-2 column
// ownership stays bank-conflict free; the ring of the last `bs` rows holds the byte costs of both columns of a
// disparity pair in one 32-bit word; the vertical sums go to vs[pair][f(column)], f(c) = (c & 1) * 64 + (c >> 1).
// Tile columns are VIRTUAL (x1 = tx0 - SW2 + c, unclamped): the x1-domain replicate clamp of A.4 is applied by
// phase 2 when it picks the columns of a window, so border tiles need no special record indexing.
// Phase 2: lanes <-> disparity pairs, a warp owns a run of 16 output columns (sliding horizontal window held in
// registers), coalesced 128-byte stores of C[y][x1][k0 .. k0+64).
// ------------------------------------------------------------------------------------------------
constexpr int kC2TXH = 128;   // tile columns incl. halo
constexpr int kC2NRH = 84;    // right-image records per parity (records 0 .. 158), padded so that the parity arrays are 16 banks apart
constexpr int kC2LH = 68;     // left-image records per parity (64), same padding
constexpr int kC2VS = 129;    // row stride of vs (odd: phase 2 reads one column of 32 pairs conflict-free)
constexpr int kC2CPW = 16;    // output columns per warp in phase 2 (even, so f() of a window column is an immediate)

inline size_t cost_tile2_smem(int SW2)
{
    const int bs = 2 * SW2 + 1;
    return size_t(2) * 2 * 2 * kC2NRH * sizeof(uint4) + size_t(2) * 2 * 2 * kC2LH * sizeof(uint4) + size_t(kCfDCP) * kC2VS * 4 +
           size_t(bs) * kCfDCP * (kC2TXH / 2) * 4;
}

template <int SW2T, bool NOPAD>
__global__ void __launch_bounds__(256, 2) k_cost_tile2(const Feat* __restrict__ fl, const Feat* __restrict__ fr,
                                                       uint16_t* __restrict__ Cvol, CostFastGeom g)
{
    extern __shared__ uint4 c2_smem[];
    const int SW2 = SW2T > 0 ? SW2T : g.SW2;
    const int bs = 2 * SW2 + 1;
    const int TX = kC2TXH - 2 * SW2;
    uint4* Rrec = c2_smem;                                  // [2 buf][2 chan][2 parity][kC2NRH]
    uint4* Lrec = Rrec + 2 * 2 * 2 * kC2NRH;                // [2 buf][2 chan][2 parity][kC2LH]
    uint32_t* vs = reinterpret_cast<uint32_t*>(Lrec + 2 * 2 * 2 * kC2LH);   // [32][kC2VS]
    uint32_t* ring = vs + kCfDCP * kC2VS;                   // [bs][32][64]

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int tx0 = blockIdx.x * TX;
    const int k0 = blockIdx.y * kCfDCP;
    const int Dh = g.Dp >> 1;
    const int ya = blockIdx.z * g.RS, yb = min(ya + g.RS, g.H);
    const int xv0 = tx0 - SW2 + g.minX1;                    // image x of (virtual) tile column 0
    const int xr_min = xv0 - g.minD - k0 - (kCfDCP - 1);    // right-image x of record 0

    // phase-1 role: columns 2m, 2m+1; pairs p0 .. p0+7
    const int m = (w >> 2) * 32 + lane, p0 = (w & 3) * 8;
    uint32_t VA[8], VB[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { VA[i] = 0; VB[i] = 0; }
    for (int i = t; i < bs * kCfDCP * (kC2TXH / 2); i += 256) ring[i] = 0;

    // Staging of a row is split: the global loads are issued before phase 1 of the previous row and land in registers
    // while it runs; the repacked records are written to the other buffer after phase 2.
    Feat sa, sb, sl;
    auto stage_load = [&](int s) {
        const int e = min(max(ya - SW2 + s, 0), g.H - 1);
        if (t < 160) {
            const int xr = xr_min + t;
            sa = __ldg(fr + size_t(e) * g.W + min(max(xr, 0), g.W - 1));
            sb = __ldg(fr + size_t(e) * g.W + min(max(xr - Dh, 0), g.W - 1));
        }
        if (t >= 128) sl = __ldg(fl + size_t(e) * g.W + min(max(xv0 + t - 128, 0), g.W - 1));
    };
    auto stage_store = [&](int buf) {
        if (t < 160) {
            uint4* d = Rrec + ((buf * 2 + 0) * 2 + (t & 1)) * kC2NRH + (t >> 1);
            // (v, lo, -hi, -v) of the pixel pair (xr, xr - Dh), low half = xr
            d[0] = make_uint4(__byte_perm(sa.x, sb.x, 0x5410), __byte_perm(sa.x, sb.x, 0x7632),
                              __byte_perm(sa.y, sb.y, 0x5410), __byte_perm(sa.y, sb.y, 0x7632));
            d[2 * kC2NRH] = make_uint4(__byte_perm(sa.z, sb.z, 0x5410), __byte_perm(sa.z, sb.z, 0x7632),
                                       __byte_perm(sa.w, sb.w, 0x5410), __byte_perm(sa.w, sb.w, 0x7632));
        }
        if (t >= 128) {
            const int c = t - 128;
            uint4* d = Lrec + ((buf * 2 + 0) * 2 + (c & 1)) * kC2LH + (c >> 1);
            // (u, -u, lo, -hi), each replicated in both halves
            d[0] = make_uint4(__byte_perm(sl.x, sl.x, 0x1010), __byte_perm(sl.y, sl.y, 0x3232),
                              __byte_perm(sl.x, sl.x, 0x3232), __byte_perm(sl.y, sl.y, 0x1010));
            d[2 * kC2LH] = make_uint4(__byte_perm(sl.z, sl.z, 0x1010), __byte_perm(sl.w, sl.w, 0x3232),
                                      __byte_perm(sl.z, sl.z, 0x3232), __byte_perm(sl.w, sl.w, 0x1010));
        }
    };
    // Birchfield-Tomasi cost of one (left pixel, right pixel pair): sobel + (raw >> 2), 16x2
    auto bt = [](const uint4& ls, const uint4& lr, const uint4& rs, const uint4& rr) {
        uint32_t X = __vadd2(rs.y, ls.y);
        uint32_t c0 = __viaddmax_s16x2_relu(ls.x, rs.z, X);
        uint32_t Y = __vadd2(ls.z, rs.w);
        uint32_t c1 = __viaddmax_s16x2_relu(rs.x, ls.w, Y);
        const uint32_t cs = __vmins2(c0, c1);
        X = __vadd2(rr.y, lr.y);
        c0 = __viaddmax_s16x2_relu(lr.x, rr.z, X);
        Y = __vadd2(lr.z, rr.w);
        c1 = __viaddmax_s16x2_relu(rr.x, lr.w, Y);
        const uint32_t cr = __vmins2(c0, c1);
        return cs + __byte_perm(cr, 0, 0x4341);   // + (cost_raw >> 2)
    };

    sa = sb = sl = make_uint4(0, 0, 0, 0);
    stage_load(0);
    stage_store(0);
    __syncthreads();
    const int nsteps = (yb - ya) + bs - 1;
    // phase-2 role
    const int c_lo = w * kC2CPW, c_hi = min(c_lo + kC2CPW, TX);
    const int kk = k0 + lane;
    const uint32_t pad_or = kk >= g.D ? kMaxCostX2 : (kk + Dh >= g.D ? (uint32_t(kMaxCost) << 16) : 0u);
    const uint32_t pad_and = kk >= g.D ? 0u : (kk + Dh >= g.D ? 0x0000FFFFu : 0xFFFFFFFFu);
    const bool lane_ok = kk < Dh;
    const size_t colBytes = size_t(g.Dp) * 2;
    const int ncol = min(c_hi, g.W1 - tx0) - c_lo;           // output columns of this run that exist in the image
    // tile columns that exist in the x1 domain: windows clamp to [cmin, cmax]
    const int cmin = max(0, SW2 - tx0), cmax = min(kC2TXH - 1, g.W1 - 1 - tx0 + SW2);
    const bool interior = cmin == 0 && cmax == kC2TXH - 1;
    auto vsf = [](int c) { return (c & 1) * 64 + (c >> 1); };
    int slot = 0;
    for (int s = 0; s < nsteps; s++) {
        const int buf = s & 1;
        if (s + 1 < nsteps) stage_load(s + 1);
        // ---- phase 1
        {
            const uint4* Lb = Lrec + buf * 4 * kC2LH + m;
            const uint4 lsA = Lb[0], lsB = Lb[kC2LH], lrA = Lb[2 * kC2LH], lrB = Lb[3 * kC2LH];
            uint32_t* rrow = ring + size_t(slot) * kCfDCP * 64 + p0 * 64 + m;
            slot = slot + 1 == bs ? 0 : slot + 1;
            // record of (column 2m, pair p): 2m + 31 - p; p even -> odd record m + 15 - p/2, p odd -> even record m + (31-p)/2
            const uint4* Rb = Rrec + buf * 4 * kC2NRH + m - (p0 >> 1);
            uint4 rsB = Rb[16], rrB = Rb[2 * kC2NRH + 16];          // record 2m + 32 - p0 (even)
            uint32_t* vrowA = vs + p0 * kC2VS + m;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int off = (i & 1) ? (31 - i) / 2 : kC2NRH + 15 - i / 2;
                const uint4 rsA = Rb[off], rrA = Rb[2 * kC2NRH + off];
                const uint32_t pdA = bt(lsA, lrA, rsA, rrA);
                const uint32_t pdB = bt(lsB, lrB, rsB, rrB);
                const uint32_t rw = rrow[i * 64];
                rrow[i * 64] = __byte_perm(pdA, pdB, 0x6420);
                VA[i] = VA[i] + pdA - __byte_perm(rw, 0, 0x4140);
                VB[i] = VB[i] + pdB - __byte_perm(rw, 0, 0x4342);
                vrowA[i * kC2VS] = VA[i];
                vrowA[i * kC2VS + 64] = VB[i];
                rsB = rsA; rrB = rrA;
            }
        }
        __syncthreads();
        // ---- phase 2 (once the vertical window is full) + staging of the next row
        if (s >= bs - 1 && ncol > 0 && lane_ok) {
            const int y = ya + s - (bs - 1);
            const uint32_t* vrow = vs + lane * kC2VS;
            char* out = reinterpret_cast<char*>(Cvol + (size_t(y) * g.W1 + tx0 + c_lo) * g.Dp + 2 * kk);
            uint32_t hs = 0;
            if (SW2T > 0) {
                constexpr int BS = 2 * SW2T + 1;
                uint32_t v[kC2CPW + BS - 1];
                if (interior) {
                    const uint32_t* vb = vrow + (c_lo >> 1);
#pragma unroll
                    for (int jj = 0; jj < kC2CPW + BS - 1; jj++) v[jj] = vb[(jj & 1) * 64 + (jj >> 1)];   // may run past the tile: harmless
                } else {
#pragma unroll
                    for (int jj = 0; jj < kC2CPW + BS - 1; jj++) v[jj] = vrow[vsf(min(max(c_lo + jj, cmin), cmax))];
                }
#pragma unroll
                for (int jj = 0; jj < BS; jj++) hs += v[jj];
#pragma unroll
                for (int c = 0; c < kC2CPW; c++) {
                    if (c > 0) hs = hs + v[c + BS - 1] - v[c - 1];
                    if (c < ncol) *reinterpret_cast<uint32_t*>(out + c * colBytes) = NOPAD ? hs : ((hs & pad_and) | pad_or);
                }
            } else {
                for (int jj = 0; jj < bs; jj++) hs += vrow[vsf(min(max(c_lo + jj, cmin), cmax))];
                for (int c = 0; c < ncol; c++) {
                    if (c > 0) hs = hs + vrow[vsf(min(max(c_lo + c + bs - 1, cmin), cmax))] - vrow[vsf(min(max(c_lo + c - 1, cmin), cmax))];
                    *reinterpret_cast<uint32_t*>(out + c * colBytes) = NOPAD ? hs : ((hs & pad_and) | pad_or);
                }
            }
        }
        if (s + 1 < nsteps) stage_store(buf ^ 1);
        __syncthreads();
    }
}

}  // namespace b200sgm
