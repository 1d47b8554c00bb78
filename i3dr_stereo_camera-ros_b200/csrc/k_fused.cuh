// Fused aggregation kernels (A.5 + A.6): the fast path.
//
//   k_horiz : one warp per image row runs BOTH horizontal chains (-> from the left end, <- from the right
//             end) in lock step; the first visit of a column parks its L in the S_h volume, the second
//             visit adds the other direction.  C is read once per chain, S_h is written once per visit.
//   k_vert  : one co-resident CTA per column strip sweeps the rows top-down (or bottom-up for the second
//             pass of MODE_HH), one warp per column.  The vertical path lives in registers, the two
//             diagonal paths move between neighbouring warps through shared memory and between
//             neighbouring CTAs through a small global exchange buffer guarded by per-strip row flags.
//             A strip publishes its outgoing diagonal FIRST and consumes its incoming diagonal LAST in
//             each row, so the flag latency hides behind the row's own work.  The summed cost feeds the
//             winner-take-all directly from registers (no S volume in MODE_SGBM).
#pragma once
#include "sgm_types.h"
#include "k_path.cuh"
#include "k_wta.cuh"

namespace b200sgm {

// ------------------------------------------------------------------------------------------------
// Horizontal pair.  Launch: one warp per row.
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) k_horiz(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Sh,
                                               int W1, int H, int Dp, uint32_t P1x2, uint32_t P2x2)
{
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= H) return;
    const bool active = lane * 2 * N < Dp;
    const uint16_t* Crow = Cvol + size_t(y) * W1 * Dp + lane * 2 * N;
    uint16_t* Srow = Sh + size_t(y) * W1 * Dp + lane * 2 * N;
    uint32_t La[N], Lb[N], Ca[N], Cb[N], Cna[N], Cnb[N];
#pragma unroll
    for (int j = 0; j < N; j++) { La[j] = 0; Lb[j] = 0; Ca[j] = Cb[j] = Cna[j] = Cnb[j] = kMaxCostX2; }
    uint32_t ma = 0, mb = 0;
    if (active) { ldg_regs<N>(Crow, Ca); ldg_regs<N>(Crow + size_t(W1 - 1) * Dp, Cb); }
    const int half = W1 >> 1;
    // first visits: xa = i < xb = W1-1-i
    for (int i = 0; i < half; i++) {
        const int xa = i, xb = W1 - 1 - i;
        if (active) {   // next columns: xa+1 <= xb-1 unless the chains meet; both stay inside the row
            ldg_regs<N>(Crow + size_t(xa + 1) * Dp, Cna);
            ldg_regs<N>(Crow + size_t(xb - 1) * Dp, Cnb);
        }
        path_step<N>(Ca, La, ma, P1x2, P2x2, lane);
        path_step<N>(Cb, Lb, mb, P1x2, P2x2, lane);
        if (active) { st_regs<N>(Srow + size_t(xa) * Dp, La); st_regs<N>(Srow + size_t(xb) * Dp, Lb); }
#pragma unroll
        for (int j = 0; j < N; j++) { Ca[j] = Cna[j]; Cb[j] = Cnb[j]; }
    }
    int i = half;
    if (W1 & 1) {   // both chains stand on the middle column: Ca == Cb == C[half]
        path_step<N>(Ca, La, ma, P1x2, P2x2, lane);
        path_step<N>(Cb, Lb, mb, P1x2, P2x2, lane);
        if (active) {
            uint32_t S[N];
#pragma unroll
            for (int j = 0; j < N; j++) S[j] = __vminu2(La[j] + Lb[j], kMaxCostX2);
            st_regs<N>(Srow + size_t(half) * Dp, S);
            if (half + 1 < W1) { ldg_regs<N>(Crow + size_t(half + 1) * Dp, Ca); ldg_regs<N>(Crow + size_t(half - 1) * Dp, Cb); }
        }
        i = half + 1;
    } else if (active && half < W1) {
        // even width: after the loop Ca holds C[half] and Cb holds C[half-1], exactly the next columns
    }
    // second visits: xa = i > xb = W1-1-i; S_h[xa] holds L<-, S_h[xb] holds L->
    for (; i < W1; i++) {
        const int xa = i, xb = W1 - 1 - i;
        uint32_t Sa[N], Sb[N];
#pragma unroll
        for (int j = 0; j < N; j++) { Sa[j] = 0; Sb[j] = 0; }
        if (active) {
            ld_regs<N>(Srow + size_t(xa) * Dp, Sa);
            ld_regs<N>(Srow + size_t(xb) * Dp, Sb);
            if (i + 1 < W1) { ldg_regs<N>(Crow + size_t(xa + 1) * Dp, Cna); ldg_regs<N>(Crow + size_t(xb - 1) * Dp, Cnb); }
        }
        path_step<N>(Ca, La, ma, P1x2, P2x2, lane);
        path_step<N>(Cb, Lb, mb, P1x2, P2x2, lane);
        if (active) {
#pragma unroll
            for (int j = 0; j < N; j++) { Sa[j] = __vminu2(Sa[j] + La[j], kMaxCostX2); Sb[j] = __vminu2(Sb[j] + Lb[j], kMaxCostX2); }
            st_regs<N>(Srow + size_t(xa) * Dp, Sa);
            st_regs<N>(Srow + size_t(xb) * Dp, Sb);
        }
#pragma unroll
        for (int j = 0; j < N; j++) { Ca[j] = Cna[j]; Cb[j] = Cnb[j]; }
    }
}

// ------------------------------------------------------------------------------------------------
// Register-resident WTA (A.6) used by the vertical sweep: same results as wta_pixel<> of k_wta.cuh but
// with a packed uniqueness test and the sub-pixel neighbours fetched by shuffle instead of from memory.
// S must already hold 0xFFFF in cells with disparity index >= D.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ uint32_t sel_reg(const uint32_t (&S)[N], int idx)
{
    uint32_t v = S[0];
#pragma unroll
    for (int j = 1; j < N; j++) v = (j == idx) ? S[j] : v;
    return v;
}

template <int N>
__device__ __forceinline__ int cell_value(const uint32_t (&S)[N], int k, int lane_unused)
{
    // value of cell k fetched from the lane that owns it (k is warp-uniform and in [0, 64N))
    const int owner = k / (2 * N), r = k - owner * 2 * N;
    uint32_t v = sel_reg<N>(S, r >> 1);
    v = __shfl_sync(kFullMask, v, owner);
    return int((r & 1) ? (v >> 16) : (v & 0xFFFFu));
}

template <int N>
__device__ __forceinline__ int wta_regs(const uint32_t (&S)[N], const WtaGeom& g, int x1, int lane,
                                        uint32_t* __restrict__ disp2key_row)
{
    const int kbase = lane * 2 * N;
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const uint32_t k = uint32_t(kbase + 2 * j);
        key = min(key, min((S[j] << 16) | k, (S[j] & 0xFFFF0000u) | (k + 1)));
    }
    key = __reduce_min_sync(kFullMask, key);
    const int minS = int(key >> 16), best = int(key & 0xFFFFu);
    if (minS >= kMaxCost) return g.INVALID;
    const int f = 100 - g.uniq;
    bool bad = false;
    if (f > 0) {
        // S[k] * f < minS * 100  <=>  S[k] < ceil(minS*100 / f); cells best-1..best+1 are exempt
        const uint32_t thr = uint32_t(min((minS * 100 + f - 1) / f, 0xFFFF));
        uint32_t mm = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int k = kbase + 2 * j;
            uint32_t v = S[j];
            if (uint32_t(k - best + 1) <= 2u) v |= 0x0000FFFFu;
            if (uint32_t(k + 1 - best + 1) <= 2u) v |= 0xFFFF0000u;
            mm = __vminu2(mm, v);
        }
        bad = min(mm & 0xFFFFu, mm >> 16) < thr;
    } else {
        const int T = minS * 100;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int k = kbase + 2 * j;
            const int s0 = int(S[j] & 0xFFFFu), s1 = int(S[j] >> 16);
            if (k < g.D && abs(k - best) > 1 && s0 * f < T) bad = true;
            if (k + 1 < g.D && abs(k + 1 - best) > 1 && s1 * f < T) bad = true;
        }
    }
    if (__any_sync(kFullMask, bad)) return g.INVALID;
    int dfix = best * 16;
    if (best > 0 && best < g.D - 1) {
        const int sm = cell_value<N>(S, best - 1, lane), sp = cell_value<N>(S, best + 1, lane);
        const int den = max(sm + sp - 2 * minS, 1);
        dfix += ((sm - sp) * 16 + den) / (den * 2);
    }
    if (lane == 0) {
        const int x = x1 + g.minX1;
        const int x2 = x - best - g.minD;
        if (x2 >= 0 && x2 < g.W) atomicMin(disp2key_row + x2, (uint32_t(minS) << 16) | uint32_t(0xFFFF - x));
    }
    return dfix + g.minD * 16;
}

// ------------------------------------------------------------------------------------------------
// Vertical sweep.
// ------------------------------------------------------------------------------------------------
struct VertGeom {
    WtaGeom w;
    int nstrips;
    int twmax;            // warps per CTA = widest strip
    uint32_t P1x2, P2x2;
    long long spin_limit; // clock64 ticks before a flag wait gives up
};

// Inter-strip exchange, "low latency" protocol: a record is Dp/2 entries of {two packed costs, tag}; each
// entry is one 8-byte volatile store, so data and flag arrive together and no fence or separate flag is
// needed.  tag = row + 1 (the buffer is zeroed before the launch).  [side][strip][row & 3][Dp/2] uint2.
// Four generations are live at once: a strip publishes row R at the START of its row R, having only
// waited (end of row R-1) for the neighbour's record R-2 -- and that neighbour still reads our record
// R-3 at the END of its row R-2.
constexpr int kXbufGen = 4;
__device__ __forceinline__ uint2* xrec(uint2* xbuf, int nstrips, int Dp, int side, int strip, int row)
{
    return xbuf + (size_t((side * nstrips + strip) * kXbufGen + (row & (kXbufGen - 1)))) * (Dp / 2);
}
__device__ __forceinline__ void st_volatile_v2(uint2* p, uint32_t a, uint32_t b)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p)
{
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

template <int N, bool UP, bool DO_WTA>
__global__ void __launch_bounds__(512, 1) k_vert(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Svol, VertGeom g,
                                                  int16_t* __restrict__ disp, uint32_t* __restrict__ disp2key,
                                                  uint2* __restrict__ xbuf, int* __restrict__ err)
{
    extern __shared__ __align__(16) uint16_t smem_v[];
    const int W1 = g.w.W1, H = g.w.H, Dp = g.w.Dp;
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int b = blockIdx.x, n = g.nstrips;
    const int x0 = int((long long)b * W1 / n), x1e = int((long long)(b + 1) * W1 / n);
    const int TW = x1e - x0;
    const int slots = g.twmax + 2;
    // smem: Ld[parity][dir][slot][Dp] then mins[parity][dir][slot] (uint32)
    uint16_t* Ld = smem_v;
    uint32_t* Md = reinterpret_cast<uint32_t*>(smem_v + size_t(4) * slots * Dp);
    auto ld_ptr = [&](int par, int dir, int slot) { return Ld + (size_t((par * 2 + dir) * slots + slot)) * Dp + lane * 2 * N; };
    auto md_ptr = [&](int par, int dir, int slot) { return Md + (par * 2 + dir) * slots + slot; };
    const bool col = j < TW;
    const bool active = lane * 2 * N < Dp;
    const int x = x0 + j;
    const bool left_edge = col && j == 0, right_edge = col && j == TW - 1;
    const bool has_left_nb = b > 0, has_right_nb = b < n - 1;

    uint32_t Lv[N], Cc[N], Cn[N], Sc[N], Sn[N];
#pragma unroll
    for (int q = 0; q < N; q++) { Lv[q] = 0; Cc[q] = Cn[q] = kMaxCostX2; Sc[q] = Sn[q] = 0; }
    uint32_t mv = 0;
    const size_t colOff = size_t(x) * Dp + lane * 2 * N;
    const size_t rowStride = size_t(W1) * Dp;
    {
        const int y = UP ? H - 1 : 0;
        if (col && active) { ldg_regs<N>(Cvol + size_t(y) * rowStride + colOff, Cc); ld_regs<N>(Svol + size_t(y) * rowStride + colOff, Sc); }
    }
    bool dead = false;

    for (int r = 0; r < H; r++) {
        const int y = UP ? H - 1 - r : r;
        const int cur = r & 1, prv = cur ^ 1;
        if (col) {
            if (r + 1 < H && active) {
                const int yn = UP ? y - 1 : y + 1;
                ldg_regs<N>(Cvol + size_t(yn) * rowStride + colOff, Cn);
                ld_regs<N>(Svol + size_t(yn) * rowStride + colOff, Sn);
            }
            // direction 0: predecessor column x-1 (slot j); direction 1: predecessor column x+1 (slot j+2)
            const int dirA = left_edge ? 1 : 0, dirB = 1 - dirA;
            uint32_t LA[N], LB[N];
            uint32_t mA = 0, mB = 0;
            // ---- step A: never needs a halo unless the strip is a single column wide (excluded by the host)
            {
                const int slot = dirA == 0 ? j : j + 2;
#pragma unroll
                for (int q = 0; q < N; q++) LA[q] = 0;
                if (r > 0 && !(dirA == 0 && x == 0) && !(dirA == 1 && x == W1 - 1)) {
                    if (active) ld_regs<N>(ld_ptr(prv, dirA, slot), LA);
                    else {
#pragma unroll
                        for (int q = 0; q < N; q++) LA[q] = kMaxCostX2;
                    }
                    mA = *md_ptr(prv, dirA, slot);
                }
                path_step<N>(Cc, LA, mA, g.P1x2, g.P2x2, lane);
                if (active) st_regs<N>(ld_ptr(cur, dirA, j + 1), LA);
                if (lane == 0) *md_ptr(cur, dirA, j + 1) = mA;
                // publish to the neighbouring strip: left edge sends dir 1 to strip b-1, right edge sends dir 0 to b+1
                const bool pub = (left_edge && has_left_nb) || (right_edge && has_right_nb && dirA == 0);
                if (pub && active) {
                    // record side 1 = left-edge (<-down) values, side 0 = right-edge (->down) values
                    uint2* rec = xrec(xbuf, n, Dp, dirA, b, r) + lane * N;
#pragma unroll
                    for (int q = 0; q < N; q++) st_volatile_v2(rec + q, LA[q], uint32_t(r + 1));
                }
            }
            // ---- vertical path: registers only
            path_step<N>(Cc, Lv, mv, g.P1x2, g.P2x2, lane);
            // ---- step B: may need the halo published by the neighbouring strip for row r-1
            {
                const int slot = dirB == 0 ? j : j + 2;
#pragma unroll
                for (int q = 0; q < N; q++) LB[q] = 0;
                const bool outside = (dirB == 0 && x == 0) || (dirB == 1 && x == W1 - 1);
                if (r > 0 && !outside) {
                    const bool halo = (dirB == 0 && j == 0) || (dirB == 1 && j == TW - 1);
                    if (halo) {
                        const int nb = dirB == 0 ? b - 1 : b + 1;
                        // want nb's right-edge record (side 0) for direction 0, its left-edge record (side 1) for direction 1
                        const uint2* rec = xrec(xbuf, n, Dp, dirB, nb, r - 1) + lane * N;
#pragma unroll
                        for (int q = 0; q < N; q++) LB[q] = kMaxCostX2;
                        if (active && !dead) {
                            const long long t0 = clock64();
                            int spins = 0;
                            while (true) {
                                bool ok = true;
#pragma unroll
                                for (int q = 0; q < N; q++) {
                                    uint2 v = ld_volatile_v2(rec + q);
                                    LB[q] = v.x;
                                    ok = ok && v.y == uint32_t(r);
                                }
                                if (ok) break;
                                if ((++spins & 255) == 0 &&
                                    (clock64() - t0 > g.spin_limit || *reinterpret_cast<volatile int*>(err))) {
                                    atomicExch(err, 1);
                                    dead = true;
                                    break;
                                }
                            }
                        }
                        dead = __any_sync(kFullMask, dead);
                        mB = warp_min16<N>(LB);
                    } else {
                        if (active) ld_regs<N>(ld_ptr(prv, dirB, slot), LB);
                        else {
#pragma unroll
                            for (int q = 0; q < N; q++) LB[q] = kMaxCostX2;
                        }
                        mB = *md_ptr(prv, dirB, slot);
                    }
                }
                path_step<N>(Cc, LB, mB, g.P1x2, g.P2x2, lane);
                if (active) st_regs<N>(ld_ptr(cur, dirB, j + 1), LB);
                if (lane == 0) *md_ptr(cur, dirB, j + 1) = mB;
            }
            // ---- sum: S = sat(S_h + L_v + L_A + L_B)
            uint32_t S[N];
#pragma unroll
            for (int q = 0; q < N; q++) {
                uint32_t s = __vminu2(Sc[q] + Lv[q], kMaxCostX2);
                s = __vminu2(s + LA[q], kMaxCostX2);
                S[q] = __vminu2(s + LB[q], kMaxCostX2);
            }
            if (DO_WTA) {
#pragma unroll
                for (int q = 0; q < N; q++) {   // cells beyond D never win and never veto
                    const int k = lane * 2 * N + 2 * q;
                    if (k >= g.w.D) S[q] = 0xFFFFFFFFu;
                    else if (k + 1 >= g.w.D) S[q] |= 0xFFFF0000u;
                }
                int d = wta_regs<N>(S, g.w, x, lane, disp2key + size_t(y) * g.w.W);
                if (lane == 0) disp[size_t(y) * g.w.W + x + g.w.minX1] = int16_t(d);
            } else if (active) {
                st_regs<N>(Svol + size_t(y) * rowStride + colOff, S);
            }
#pragma unroll
            for (int q = 0; q < N; q++) { Cc[q] = Cn[q]; Sc[q] = Sn[q]; }
        }
        __syncthreads();
    }
}

}  // namespace b200sgm
