// Min-plus path aggregation (A.5): the warp-per-chain step and the generic per-direction kernels.
//
// Volume layout ("paired"): a pixel's Dp costs are stored as Dh = Dp/2 uint32 words; word w holds cell w in its
// low half and cell Dh + w in its high half.  One warp owns a pixel's vector; lane l holds words l*N .. l*N+N-1,
// i.e. cells l*N+j (low halves) and Dh+l*N+j (high halves).  With this pairing the d-1 / d+1 neighbours of BOTH
// halves of register j are simply registers j-1 / j+1 (one shuffle per side at the lane boundary, which wraps
// from the last active lane to lane 0 where the low half ends and the high half begins) -- no byte permutes.
//
// The chain state is kept NORMALISED: Lt = L - min_k L.  With it
//     L_new[k] = C[k] + min(Lt[k], Lt[k-1] + P1, Lt[k+1] + P1, P2)
//              = C[k] + min(min3(Lt[k-1], Lt[k+1], P2 - P1) + P1, Lt[k])           (VIMNMX3 + VIADDMNMX)
// which is cv::StereoSGBM's update with the "- min" folded into the state.
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

// Per-thread constants of the paired layout.
struct LaneCtx {
    int lane;
    int src_up, src_dn;      // lanes that hold cells k-1 of my first word / k+1 of my last word
    uint32_t sel_up, sel_dn; // byte-permute selectors fixing the two wrap points (identity elsewhere)
    bool active;             // lane * N < Dh
    uint32_t P1x2, P2mP1x2;  // P1 and P2 - P1 replicated in both halves
};

template <int N>
__device__ __forceinline__ LaneCtx make_lane_ctx(int lane, int Dp, int P1, int P2)
{
    LaneCtx c;
    const int Dh = Dp >> 1;
    const int last = Dh / N - 1;               // last active lane
    c.lane = lane;
    c.active = lane <= last;
    c.src_up = lane == 0 ? last : lane - 1;
    c.src_dn = lane >= last ? 0 : lane + 1;
    // lane 0: low half has no k-1 (-> 0xFFFF), high half's k-1 is the LOW half fetched from the last lane
    c.sel_up = lane == 0 ? 0x1054u : 0x3210u;
    // last lane: low half's k+1 is the HIGH half fetched from lane 0, high half has no k+1 (-> 0xFFFF)
    c.sel_dn = lane == last ? 0x5432u : 0x3210u;
    c.P1x2 = uint32_t(P1) * 0x10001u;
    c.P2mP1x2 = uint32_t(P2 - P1) * 0x10001u;
    return c;
}

// The add that the compiler should place on the FMA pipe (IMAD) when `one` is an opaque 1: the ALU pipe is the
// busy one in these kernels (VIMNMX / PRMT / LOP3 all live there).
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one) { return a * one + b; }

// One chain step on normalised state.  In: C (costs of this pixel), Lt (normalised L of the predecessor).
// Out: Ln = L of this pixel (what cv::StereoSGBM stores / sums), Lt = Ln - min(Ln).  Padded cells and inactive
// lanes must carry C = kMaxCost; out-of-range neighbours are 0xFFFF ("infinity": min3 caps at P2 - P1 first, so
// nothing can wrap).
template <int N>
__device__ __forceinline__ void path_step(const uint32_t (&C)[N], uint32_t (&Lt)[N], uint32_t (&Ln)[N], const LaneCtx& c)
{
    uint32_t up = __shfl_sync(kFullMask, Lt[N - 1], c.src_up);
    uint32_t dn = __shfl_sync(kFullMask, Lt[0], c.src_dn);
    up = __byte_perm(up, 0xFFFFFFFFu, c.sel_up);
    dn = __byte_perm(dn, 0xFFFFFFFFu, c.sel_dn);
#pragma unroll
    for (int j = 0; j < N; j++) {
        const uint32_t a = j > 0 ? Lt[j - 1] : up;
        const uint32_t b = j + 1 < N ? Lt[j + 1] : dn;
        const uint32_t t = __vimin3_u16x2(a, b, c.P2mP1x2);
        const uint32_t u = __viaddmin_u16x2(t, c.P1x2, Lt[j]);
        Ln[j] = C[j] + u;
    }
    const uint32_t m2 = warp_min16x2<N>(Ln);
#pragma unroll
    for (int j = 0; j < N; j++) Lt[j] = Ln[j] - m2;
}

struct PathGeom {
    int W1, H, Dp;
    int dx, dy;       // step of the chain (predecessor of (x,y) is (x-dx, y-dy))
    int nchains;
    int P1, P2;
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

// Generic per-direction kernel (validation path): S (+)= L_r.  FIRST writes S, otherwise read-modify-write with
// the int16 saturation of A.5 (min(32767, sum); all terms are non-negative).
template <int N, bool FIRST>
__global__ void __launch_bounds__(128) k_path_generic(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Svol, PathGeom g)
{
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= g.nchains) return;
    int x0, y0, len;
    chain_start(g, c, x0, y0, len);
    const LaneCtx lc = make_lane_ctx<N>(lane, g.Dp, g.P1, g.P2);
    const ptrdiff_t stride = (ptrdiff_t(g.dy) * g.W1 + g.dx) * g.Dp;
    size_t off = (size_t(y0) * g.W1 + x0) * g.Dp + lane * 2 * N;
    uint32_t Lt[N], Ln[N], Cc[N], Cn[N], Sc[N];
#pragma unroll
    for (int j = 0; j < N; j++) { Lt[j] = 0; Cc[j] = kMaxCostX2; Cn[j] = kMaxCostX2; Sc[j] = 0; }
    if (lc.active) ldg_regs<N>(Cvol + off, Cc);
    for (int i = 0; i < len; i++) {
        if (lc.active && i + 1 < len) ldg_regs<N>(Cvol + off + stride, Cn);
        if (!FIRST && lc.active) ld_regs<N>(Svol + off, Sc);
        path_step<N>(Cc, Lt, Ln, lc);
        if (lc.active) {
            if (FIRST) st_regs<N>(Svol + off, Ln);
            else {
#pragma unroll
                for (int j = 0; j < N; j++) Sc[j] = __vminu2(Sc[j] + Ln[j], kMaxCostX2);
                st_regs<N>(Svol + off, Sc);
            }
        }
#pragma unroll
        for (int j = 0; j < N; j++) Cc[j] = Cn[j];
        off += stride;
    }
}

}  // namespace b200sgm
