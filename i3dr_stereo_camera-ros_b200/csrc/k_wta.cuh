// Winner-take-all with uniqueness, x16 sub-pixel interpolation and the disp2 scatter (A.6), and the
// left-right consistency check (A.7).
#pragma once
#include "sgm_types.h"
#include "k_path.cuh"

namespace b200sgm {

struct WtaGeom {
    int W, H, W1, minX1, minD, D, Dp;
    int uniq, d12, INVALID;
};

// Index of cell k inside a pixel's paired vector (uint16 units): word k (low half) or word k - Dh (high half).
__device__ __forceinline__ int cell_u16_index(int k, int Dh) { return k < Dh ? 2 * k : 2 * (k - Dh) + 1; }

// Warp-level WTA on the lane-distributed aggregated cost S (paired layout, see k_path.cuh).
// `Srow` points at the pixel's Dp uint16 costs in memory (used to fetch the two sub-pixel neighbours).
// Returns the fixed-point disparity (already offset by minD*16) or INVALID; on a unique winner also
// posts (minS, x) to the right-image pixel x - d with the tie rule "larger x wins".
template <int N>
__device__ __forceinline__ int wta_pixel(const uint32_t (&S)[N], const uint16_t* Srow, const WtaGeom& g,
                                         int x1, int lane, uint32_t* __restrict__ disp2key_row)
{
    // key = S<<16 | k : the warp minimum is the smallest cost and, among equals, the FIRST disparity.
    uint32_t key = 0xFFFFFFFFu;
    const int Dh = g.Dp >> 1;
    const int kbase = lane * N;
#pragma unroll
    for (int j = 0; j < N; j++) {
        int k = kbase + j;
        uint32_t klo = (S[j] << 16) | uint32_t(k);
        uint32_t khi = (S[j] & 0xFFFF0000u) | uint32_t(k + Dh);
        if (k < Dh && k < g.D) key = min(key, klo);
        if (k < Dh && k + Dh < g.D) key = min(key, khi);
    }
    key = __reduce_min_sync(kFullMask, key);
    const int minS = int(key >> 16), best = int(key & 0xFFFFu);
    if (minS >= kMaxCost) return g.INVALID;  // every cost saturated: OpenCV ends up at INVALID as well
    // uniqueness: reject if some k with |k-best| > 1 has S[k]*(100-uniq) < minS*100
    const int T = minS * 100, f = 100 - g.uniq;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < N; j++) {
        int k = kbase + j;
        int s0 = int(S[j] & 0xFFFFu), s1 = int(S[j] >> 16);
        if (k < Dh && k < g.D && abs(k - best) > 1 && s0 * f < T) bad = true;
        if (k < Dh && k + Dh < g.D && abs(k + Dh - best) > 1 && s1 * f < T) bad = true;
    }
    if (__any_sync(kFullMask, bad)) return g.INVALID;
    int dfix = best * 16;
    if (lane == 0) {
        const int x = x1 + g.minX1;
        const int x2 = x - best - g.minD;
        if (x2 >= 0 && x2 < g.W) atomicMin(disp2key_row + x2, (uint32_t(minS) << 16) | uint32_t(0xFFFF - x));
        if (best > 0 && best < g.D - 1) {
            int sm = Srow[cell_u16_index(best - 1, Dh)], sp = Srow[cell_u16_index(best + 1, Dh)];
            int den = max(sm + sp - 2 * minS, 1);
            dfix += ((sm - sp) * 16 + den) / (den * 2);  // C division: truncation toward zero
        }
    }
    dfix = __shfl_sync(kFullMask, dfix, 0);
    return dfix + g.minD * 16;
}

// Stand-alone WTA over a materialised S volume: one warp per valid pixel.
template <int N>
static __global__ void __launch_bounds__(256) k_wta(const uint16_t* __restrict__ Svol, WtaGeom g,
                                             int16_t* __restrict__ disp, uint32_t* __restrict__ disp2key)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= (long long)g.W1 * g.H) return;
    const int y = int(pix / g.W1), x1 = int(pix - (long long)y * g.W1);
    const uint16_t* Srow = Svol + size_t(pix) * g.Dp;
    uint32_t S[N];
#pragma unroll
    for (int j = 0; j < N; j++) S[j] = kMaxCostX2;
    if (lane * 2 * N < g.Dp) ld_regs<N>(Srow + lane * 2 * N, S);   // N paired words per lane
    int d = wta_pixel<N>(S, Srow, g, x1, lane, disp2key + size_t(y) * g.W);
    if (lane == 0) disp[size_t(y) * g.W + x1 + g.minX1] = int16_t(d);
}

// A.7: a pixel survives unless BOTH the floor and the ceil candidate disagree with disp2 by more than
// d12.  disp2 is reconstructed from the atomicMin key; an unassigned entry holds the SCALED invalid
// value (minD-1)*16, which counts as a (failing) candidate when it is >= minD (OpenCV quirk, kept).
static __global__ void k_lrcheck(int16_t* __restrict__ disp, const uint32_t* __restrict__ disp2key, WtaGeom g)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= g.W1) return;
    x += g.minX1;
    int16_t* row = disp + size_t(y) * g.W;
    const uint32_t* krow = disp2key + size_t(y) * g.W;
    int d1 = row[x];
    if (d1 == g.INVALID) return;
    int dlo = d1 >> 4, dhi = (d1 + 15) >> 4;
    int xlo = x - dlo, xhi = x - dhi;
    auto disp2_at = [&](int xx) {
        uint32_t k = krow[xx];
        return k == 0xFFFFFFFFu ? g.INVALID : (0xFFFF - int(k & 0xFFFFu)) - xx;
    };
    bool fail_lo = false, fail_hi = false;
    if (xlo >= 0 && xlo < g.W) { int d2 = disp2_at(xlo); fail_lo = d2 >= g.minD && abs(d2 - dlo) > g.d12; }
    if (xhi >= 0 && xhi < g.W) { int d2 = disp2_at(xhi); fail_hi = d2 >= g.minD && abs(d2 - dhi) > g.d12; }
    if (fail_lo && fail_hi) row[x] = int16_t(g.INVALID);
}

}  // namespace b200sgm
