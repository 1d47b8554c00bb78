// Min-plus path aggregation (A.5): warp-per-chain kernels over the materialised cost volume.
// One warp walks one 1-D chain (row, column or diagonal); the D disparities are spread over the 32
// lanes, 2*N per lane, packed two uint16 per register.  All arithmetic is unsigned 16x2 SIMD
// (VIADD / VIMNMX3.U16x2 / VIMNMX.U16x2) with a CREDUX.MIN warp reduction per step.
#pragma once
#include "sgm_types.h"

namespace b200sgm {

template <int N>
__device__ __forceinline__ void ld_regs(const uint16_t* p, uint32_t (&r)[N])
{
    if constexpr (N == 1) {
        r[0] = *reinterpret_cast<const uint32_t*>(p);
    } else if constexpr (N == 2) {
        uint2 v = *reinterpret_cast<const uint2*>(p);
        r[0] = v.x; r[1] = v.y;
    } else {
#pragma unroll
        for (int j = 0; j < N / 4; j++) {
            uint4 v = reinterpret_cast<const uint4*>(p)[j];
            r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
    }
}

template <int N>
__device__ __forceinline__ void ldg_regs(const uint16_t* p, uint32_t (&r)[N])
{
    if constexpr (N == 1) {
        r[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
    } else if constexpr (N == 2) {
        uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        r[0] = v.x; r[1] = v.y;
    } else {
#pragma unroll
        for (int j = 0; j < N / 4; j++) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + j);
            r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
    }
}

template <int N>
__device__ __forceinline__ void st_regs(uint16_t* p, const uint32_t (&r)[N])
{
    if constexpr (N == 1) {
        *reinterpret_cast<uint32_t*>(p) = r[0];
    } else if constexpr (N == 2) {
        *reinterpret_cast<uint2*>(p) = make_uint2(r[0], r[1]);
    } else {
#pragma unroll
        for (int j = 0; j < N / 4; j++)
            reinterpret_cast<uint4*>(p)[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
}

// Warp-wide minimum over all uint16 halves of r[0..N), returned replicated in both halves (m * 0x10001):
// after the in-lane tree both halves of every lane's word are made equal, so the 32-bit CREDUX.MIN of
// those words is the word with the smallest half.
template <int N>
__device__ __forceinline__ uint32_t warp_min16x2(const uint32_t (&r)[N])
{
    uint32_t mm = r[0];
#pragma unroll
    for (int j = 1; j < N; j++) mm = __vminu2(mm, r[j]);
    mm = __vminu2(mm, __byte_perm(mm, 0, 0x1032));
    return __reduce_min_sync(kFullMask, mm);
}

// One step of  L[k] = C[k] + min(Lp[k], Lp[k-1]+P1, Lp[k+1]+P1, m+P2) - m  on the lane's 2N disparities.
// Lp is replaced by L; m2 (the warp-uniform minimum of Lp, replicated in both halves) is replaced by that
// of L.  Out-of-range neighbours and padded cells hold kMaxCost ("infinity"; kMaxCost + P1 + P2 < 65536 is
// validated on the host).
template <int N>
__device__ __forceinline__ void path_step(const uint32_t (&C)[N], uint32_t (&Lp)[N], uint32_t& m2,
                                          uint32_t P1x2, uint32_t P2x2, int lane)
{
    uint32_t up = __shfl_up_sync(kFullMask, Lp[N - 1], 1);
    uint32_t dn = __shfl_down_sync(kFullMask, Lp[0], 1);
    if (lane == 0) up = kMaxCostX2;
    if (lane == 31) dn = kMaxCostX2;
    const uint32_t mP2 = m2 + P2x2;
    uint32_t q_lo = __byte_perm(up, Lp[0], 0x5432) + P1x2;  // (Lp[2j-1], Lp[2j]) + P1
    uint32_t L[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        uint32_t nxt = (j + 1 < N) ? Lp[j + 1] : dn;
        uint32_t q_hi = __byte_perm(Lp[j], nxt, 0x5432) + P1x2;  // (Lp[2j+1], Lp[2j+2]) + P1
        uint32_t t = __vimin3_u16x2(Lp[j], q_lo, q_hi);
        t = __vminu2(t, mP2);
        L[j] = C[j] + t - m2;
        q_lo = q_hi;
    }
#pragma unroll
    for (int j = 0; j < N; j++) Lp[j] = L[j];
    m2 = warp_min16x2<N>(L);
}

struct PathGeom {
    int W1, H, Dp;
    int dx, dy;       // step of the chain (predecessor of (x,y) is (x-dx, y-dy))
    int nchains;
    uint32_t P1x2, P2x2;
};

__device__ __forceinline__ void chain_start(const PathGeom& g, int c, int& x0, int& y0, int& len)
{
    if (g.dy == 0) { y0 = c; x0 = g.dx > 0 ? 0 : g.W1 - 1; len = g.W1; return; }
    if (g.dx == 0) { x0 = c; y0 = g.dy > 0 ? 0 : g.H - 1; len = g.H; return; }
    if (c < g.W1) { x0 = c; y0 = g.dy > 0 ? 0 : g.H - 1; }
    else { int k = c - g.W1 + 1; x0 = g.dx > 0 ? 0 : g.W1 - 1; y0 = g.dy > 0 ? k : g.H - 1 - k; }
    int lx = g.dx > 0 ? g.W1 - x0 : x0 + 1, ly = g.dy > 0 ? g.H - y0 : y0 + 1;
    len = min(lx, ly);
}

inline int chain_count(int W1, int H, int dx, int dy)
{
    if (dy == 0) return H;
    if (dx == 0) return W1;
    return W1 + H - 1;
}

// Generic per-direction kernel: S (+)= L_r.  FIRST writes S, otherwise read-modify-write with the
// int16 saturation of A.5 (min(32767, sum); all terms are non-negative).
template <int N, bool FIRST>
__global__ void __launch_bounds__(128) k_path_generic(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Svol, PathGeom g)
{
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= g.nchains) return;
    int x0, y0, len;
    chain_start(g, c, x0, y0, len);
    const bool active = lane * 2 * N < g.Dp;
    const ptrdiff_t stride = (ptrdiff_t(g.dy) * g.W1 + g.dx) * g.Dp;
    size_t off = (size_t(y0) * g.W1 + x0) * g.Dp + lane * 2 * N;
    uint32_t Lp[N], Cc[N], Cn[N], Sc[N];
#pragma unroll
    for (int j = 0; j < N; j++) { Lp[j] = 0; Cc[j] = kMaxCostX2; Cn[j] = kMaxCostX2; Sc[j] = 0; }
    uint32_t m = 0;
    if (active) ldg_regs<N>(Cvol + off, Cc);
    for (int i = 0; i < len; i++) {
        if (active && i + 1 < len) ldg_regs<N>(Cvol + off + stride, Cn);
        if (!FIRST && active) ld_regs<N>(Svol + off, Sc);
        path_step<N>(Cc, Lp, m, g.P1x2, g.P2x2, lane);
        if (active) {
            if (FIRST) st_regs<N>(Svol + off, Lp);
            else {
#pragma unroll
                for (int j = 0; j < N; j++) Sc[j] = __vminu2(Sc[j] + Lp[j], kMaxCostX2);
                st_regs<N>(Svol + off, Sc);
            }
        }
#pragma unroll
        for (int j = 0; j < N; j++) Cc[j] = Cn[j];
        off += stride;
    }
}

}  // namespace b200sgm
