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
#include <type_traits>
#include "sgm_types.h"
#include "k_path.cuh"
#include "k_wta.cuh"

namespace b200sgm {

constexpr int kVertRing = 4;   // rows of C / S_h in flight per column (cp.async ring in shared memory)

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc)
{
    const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
// copies this lane's 2N costs
template <int N>
__device__ __forceinline__ void cp_async_lane(uint16_t* smem_dst, const uint16_t* gsrc)
{
    if constexpr (N <= 4) cp_async<4 * N>(smem_dst, gsrc);
    else {
#pragma unroll
        for (int q = 0; q < N / 4; q++) cp_async<16>(smem_dst + 8 * q, gsrc + 8 * q);
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }


// ------------------------------------------------------------------------------------------------
// Horizontal pair.  Launch: one warp per row.
// ------------------------------------------------------------------------------------------------
constexpr int kHorizRing = 6;   // steps of C (and S_h) in flight per chain: cp.async ring in shared memory

// dynamic smem: warps * 4 * kHorizRing * Dp * 2 bytes
template <int N>
__global__ void __launch_bounds__(128) k_horiz(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Sh,
                                               int W1, int H, int Dp, uint32_t P1x2, uint32_t P2x2)
{
    extern __shared__ __align__(16) uint16_t smem_h[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int y = blockIdx.x * (blockDim.x >> 5) + wib;
    if (y >= H) return;
    const bool active = lane * 2 * N < Dp;
    const int lo = lane * 2 * N;
    const uint16_t* Crow = Cvol + size_t(y) * W1 * Dp + lo;
    uint16_t* Srow = Sh + size_t(y) * W1 * Dp + lo;
    // per-warp rings: [array: Ca, Cb, Sa, Sb][slot][Dp]
    uint16_t* ring = smem_h + size_t(wib) * 4 * kHorizRing * Dp + lo;
    auto slot = [&](int arr, int s) { return ring + size_t(arr * kHorizRing + (s % kHorizRing)) * Dp; };
    constexpr int PF = kHorizRing - 1;
    const int half = W1 >> 1;
    const int sec = half + (W1 & 1);                 // first second-visit step
    auto issue = [&](int s) {                        // one commit group per step, even past the end
        if (s < W1 && active) {
            const int xa = s, xb = W1 - 1 - s;
            cp_async_lane<N>(slot(0, s), Crow + size_t(xa) * Dp);
            cp_async_lane<N>(slot(1, s), Crow + size_t(xb) * Dp);
            if (s >= sec + PF) {                     // both first-visit stores are >= PF+1 steps old by now
                cp_async_lane<N>(slot(2, s), Srow + size_t(xa) * Dp);
                cp_async_lane<N>(slot(3, s), Srow + size_t(xb) * Dp);
            }
        }
        cp_async_commit();
    };
    for (int s = 0; s < PF; s++) issue(s);
    uint32_t La[N], Lb[N];
#pragma unroll
    for (int j = 0; j < N; j++) { La[j] = 0; Lb[j] = 0; }
    uint32_t ma = 0, mb = 0;
    for (int i = 0; i < W1; i++) {
        issue(i + PF);
        cp_async_wait<PF>();                         // this lane's copies for step i have landed
        const int xa = i, xb = W1 - 1 - i;
        uint32_t Ca[N], Cb[N];
        if (active) { ld_regs<N>(slot(0, i), Ca); ld_regs<N>(slot(1, i), Cb); }
        else {
#pragma unroll
            for (int j = 0; j < N; j++) { Ca[j] = kMaxCostX2; Cb[j] = kMaxCostX2; }
        }
        path_step<N>(Ca, La, ma, P1x2, P2x2, lane);
        path_step<N>(Cb, Lb, mb, P1x2, P2x2, lane);
        if (!active) continue;
        if (i < half) {                              // first visits: park L in S_h
            st_regs<N>(Srow + size_t(xa) * Dp, La);
            st_regs<N>(Srow + size_t(xb) * Dp, Lb);
        } else if (xa == xb) {                       // odd width: both chains on the middle column
            uint32_t S[N];
#pragma unroll
            for (int j = 0; j < N; j++) S[j] = __vminu2(La[j] + Lb[j], kMaxCostX2);
            st_regs<N>(Srow + size_t(xa) * Dp, S);
        } else {                                     // second visits: S_h[xa] holds L<-, S_h[xb] holds L->
            uint32_t Sa[N], Sb[N];
            if (i >= sec + PF) { ld_regs<N>(slot(2, i), Sa); ld_regs<N>(slot(3, i), Sb); }
            else { ld_regs<N>(Srow + size_t(xa) * Dp, Sa); ld_regs<N>(Srow + size_t(xb) * Dp, Sb); }
#pragma unroll
            for (int j = 0; j < N; j++) { Sa[j] = __vminu2(Sa[j] + La[j], kMaxCostX2); Sb[j] = __vminu2(Sb[j] + Lb[j], kMaxCostX2); }
            st_regs<N>(Srow + size_t(xa) * Dp, Sa);
            st_regs<N>(Srow + size_t(xb) * Dp, Sb);
        }
    }
    cp_async_wait<0>();
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
    int debug_no_exchange; // timing experiments only: never wait for / publish to neighbours (wrong results)
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

// Tight WTA on registers.  `S` holds 0xFFFF in cells >= D.  Returns the output value for the pixel.
//   umagic = floor(2^32 / f) + 1 with f = 100 - uniq > 0 (exact quotient for minS*100 + f - 1 < 2^32 / f)
template <int N>
__device__ __forceinline__ int wta_fast(const uint32_t (&S)[N], const WtaGeom& g, uint32_t umagic, int x1, int lane,
                                        uint32_t* __restrict__ disp2key_row)
{
    const int kbase = lane * 2 * N;
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int q = 0; q < N; q++) {
        const uint32_t k = uint32_t(kbase + 2 * q);
        key = min(key, min((S[q] << 16) | k, (S[q] & 0xFFFF0000u) | (k + 1)));
    }
    key = __reduce_min_sync(kFullMask, key);
    const int minS = int(key >> 16), best = int(key & 0xFFFFu);
    int out = g.INVALID;
    if (minS < kMaxCost) {
        const int f = 100 - g.uniq;
        // S[k] * f < minS * 100  <=>  S[k] < ceil(minS*100 / f); cells best-1..best+1 are exempt
        const uint32_t thr = min(__umulhi(uint32_t(minS * 100 + f - 1), umagic), 0xFFFFu);
        const int rel = best - kbase;     // local index of the winner inside this lane (may be out of range)
        uint32_t mm = 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < N; q++) {
            uint32_t v = S[q];
            if (uint32_t(2 * q - rel + 1) <= 2u) v |= 0x0000FFFFu;
            if (uint32_t(2 * q + 1 - rel + 1) <= 2u) v |= 0xFFFF0000u;
            mm = __vminu2(mm, v);
        }
        const bool bad = min(mm & 0xFFFFu, mm >> 16) < thr;
        if (!__any_sync(kFullMask, bad)) {
            int dfix = best * 16;
            if (best > 0 && best < g.D - 1) {
                const int sm = cell_value<N>(S, best - 1, lane), sp = cell_value<N>(S, best + 1, lane);
                const int den = max(sm + sp - 2 * minS, 1);
                // |quotient| <= 8.5 and numerator, denominator < 2^24: IEEE float division then truncation is exact
                dfix += __float2int_rz(__fdiv_rn(float((sm - sp) * 16 + den), float(den * 2)));
            }
            if (lane == 0) {
                const int x = x1 + g.minX1;
                const int x2 = x - best - g.minD;
                if (x2 >= 0 && x2 < g.W) atomicMin(disp2key_row + x2, (uint32_t(minS) << 16) | uint32_t(0xFFFF - x));
            }
            out = dfix + g.minD * 16;
        }
    }
    return out;
}

// FULL      : Dp == D == 64*N (no padded cells, every lane active)
// CLAMP_EACH: saturate after every addition of the sum (needed when kMaxCost + 3*(Cmax+P2) could exceed 65535)
template <int N, bool UP, bool DO_WTA, bool FULL, bool CLAMP_EACH>
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
    // smem: Ld[parity][dir][slot][Dp] costs, then Md[parity][dir][slot] minima.  Everything starts at zero: a zero
    // predecessor vector with minimum 0 is exactly "predecessor outside the image" (first row, border columns).
    uint16_t* Ld = smem_v;
    uint32_t* Md = reinterpret_cast<uint32_t*>(smem_v + size_t(4) * slots * Dp);
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(smem_v);
        const int nz = 2 * slots * Dp + 4 * slots;
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    if (j >= TW) return;                 // the row barrier below only counts the column warps
    const int nbar = 32 * TW;
    const bool active = FULL || lane * 2 * N < Dp;
    const int x = x0 + j;
    const int lo = lane * 2 * N;
    // direction 0: predecessor column x-1 (slot j); direction 1: predecessor column x+1 (slot j+2).
    // The left-edge warp runs direction 1 first (it publishes it), every other warp direction 0 first.
    const int dirA = j == 0 ? 1 : 0, dirB = 1 - dirA;
    const int slotA = dirA == 0 ? j : j + 2, slotB = dirB == 0 ? j : j + 2;
    const bool pubA = (j == 0 && b > 0) || (j == TW - 1 && j != 0 && b < n - 1);
    const bool haloB = (j == 0 && b > 0) || (j == TW - 1 && j != 0 && b < n - 1);   // same warps consume the other direction
    const int dirStride = slots * Dp, parStride = 2 * slots * Dp;
    const uint16_t* rdA[2]; uint16_t* wrA[2]; const uint16_t* rdB[2]; uint16_t* wrB[2];
    const uint32_t* mrA[2]; uint32_t* mwA[2]; const uint32_t* mrB[2]; uint32_t* mwB[2];
#pragma unroll
    for (int pz = 0; pz < 2; pz++) {
        rdA[pz] = Ld + pz * parStride + dirA * dirStride + slotA * Dp + lo;
        wrA[pz] = Ld + pz * parStride + dirA * dirStride + (j + 1) * Dp + lo;
        rdB[pz] = Ld + pz * parStride + dirB * dirStride + slotB * Dp + lo;
        wrB[pz] = Ld + pz * parStride + dirB * dirStride + (j + 1) * Dp + lo;
        mrA[pz] = Md + (pz * 2 + dirA) * slots + slotA;
        mwA[pz] = Md + (pz * 2 + dirA) * slots + j + 1;
        mrB[pz] = Md + (pz * 2 + dirB) * slots + slotB;
        mwB[pz] = Md + (pz * 2 + dirB) * slots + j + 1;
    }
    // exchange records: I publish side dirA of my strip, I consume side dirB of the neighbour
    const int nb = dirB == 0 ? b - 1 : b + 1;
    uint2* pub_base = xrec(xbuf, n, Dp, dirA, b, 0) + lane * N;
    const uint2* con_base = xrec(xbuf, n, Dp, dirB, haloB ? nb : b, 0) + lane * N;
    const int gen_stride = Dp / 2;

    const ptrdiff_t rowStride = (UP ? -1 : 1) * ptrdiff_t(W1) * Dp;
    const int ystart = UP ? H - 1 : 0;
    const uint16_t* gC = Cvol + (size_t(ystart) * W1 + x) * Dp + lo;
    uint16_t* gS = Svol + (size_t(ystart) * W1 + x) * Dp + lo;
    int16_t* dptr = disp + size_t(ystart) * g.w.W + x + g.w.minX1;
    uint32_t* kptr = disp2key + size_t(ystart) * g.w.W;
    const ptrdiff_t dStride = (UP ? -1 : 1) * ptrdiff_t(g.w.W);
    const int f = 100 - g.w.uniq;
    const uint32_t umagic = f > 0 ? uint32_t((1ull << 32) / uint32_t(f)) + 1u : 0u;

    uint32_t Lv[N], C0[N], C1[N], S0[N], S1[N];
#pragma unroll
    for (int q = 0; q < N; q++) { Lv[q] = 0; C0[q] = C1[q] = kMaxCostX2; S0[q] = S1[q] = 0; }
    uint32_t mv = 0;
    if (active) { ldg_regs<N>(gC, C0); ld_regs<N>(gS, S0); }
    bool dead = false;

    // one row; PAR = parity of r (buffer written), reads the other one
    auto row = [&](auto par_tag, int r, uint32_t (&Cc)[N], uint32_t (&Sc)[N], uint32_t (&Cn)[N], uint32_t (&Sn)[N]) {
        constexpr int PAR = decltype(par_tag)::value;
        if (r + 1 < H && active) { ldg_regs<N>(gC + rowStride, Cn); ld_regs<N>(gS + rowStride, Sn); }
        uint32_t LA[N], LB[N];
        uint32_t mA, mB;
        // ---- step A
        if (active) ld_regs<N>(rdA[PAR ^ 1], LA);
        else {
#pragma unroll
            for (int q = 0; q < N; q++) LA[q] = kMaxCostX2;
        }
        mA = *mrA[PAR ^ 1];
        path_step<N>(Cc, LA, mA, g.P1x2, g.P2x2, lane);
        if (active) st_regs<N>(wrA[PAR], LA);
        if (lane == 0) *mwA[PAR] = mA;
        if (pubA && active) {
            uint2* rec = pub_base + (r & (kXbufGen - 1)) * gen_stride;
#pragma unroll
            for (int q = 0; q < N; q++) st_volatile_v2(rec + q, LA[q], uint32_t(r + 1));
        }
        // ---- vertical path: registers only
        path_step<N>(Cc, Lv, mv, g.P1x2, g.P2x2, lane);
        // ---- step B
        if (haloB && r > 0) {
            const uint2* rec = con_base + ((r - 1) & (kXbufGen - 1)) * gen_stride;
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
                    if ((++spins & 255) == 0 && (clock64() - t0 > g.spin_limit || *reinterpret_cast<volatile int*>(err))) {
                        atomicExch(err, 1);
                        dead = true;
                        break;
                    }
                }
            }
            dead = __any_sync(kFullMask, dead);
            mB = warp_min16x2<N>(LB);
        } else {
            if (active) ld_regs<N>(rdB[PAR ^ 1], LB);
            else {
#pragma unroll
                for (int q = 0; q < N; q++) LB[q] = kMaxCostX2;
            }
            mB = *mrB[PAR ^ 1];
        }
        path_step<N>(Cc, LB, mB, g.P1x2, g.P2x2, lane);
        if (active) st_regs<N>(wrB[PAR], LB);
        if (lane == 0) *mwB[PAR] = mB;
        // ---- S = sat(S_h + L_v + L_A + L_B)
        uint32_t S[N];
#pragma unroll
        for (int q = 0; q < N; q++) {
            if (CLAMP_EACH) {
                uint32_t t = __vminu2(Sc[q] + Lv[q], kMaxCostX2);
                t = __vminu2(t + LA[q], kMaxCostX2);
                S[q] = __vminu2(t + LB[q], kMaxCostX2);
            } else {
                S[q] = __vminu2(Sc[q] + Lv[q] + LA[q] + LB[q], kMaxCostX2);
            }
        }
        if (DO_WTA) {
            if (!FULL) {
#pragma unroll
                for (int q = 0; q < N; q++) {   // cells beyond D never win and never veto
                    const int k = lo + 2 * q;
                    if (k >= g.w.D) S[q] = 0xFFFFFFFFu;
                    else if (k + 1 >= g.w.D) S[q] |= 0xFFFF0000u;
                }
            }
            int d;
            if (f > 0) d = wta_fast<N>(S, g.w, umagic, x, lane, kptr);
            else d = wta_regs<N>(S, g.w, x, lane, kptr);
            if (lane == 0) *dptr = int16_t(d);
        } else if (active) {
            st_regs<N>(gS, S);
        }
        gC += rowStride; gS += rowStride; dptr += dStride; kptr += dStride;
        asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");
    };

    int r = 0;
    for (; r + 1 < H; r += 2) {
        row(std::integral_constant<int, 0>{}, r, C0, S0, C1, S1);
        row(std::integral_constant<int, 1>{}, r + 1, C1, S1, C0, S0);
    }
    if (r < H) row(std::integral_constant<int, 0>{}, r, C0, S0, C1, S1);
}

// ------------------------------------------------------------------------------------------------
// Vertical sweep, warp specialised (the default for D <= 256):
//   path warps (0 .. TW-1)   : one per column; the three path updates of a row (the only work on the
//                              row-to-row dependency chain), the sum S, which they park in a 4-deep shared
//                              memory ring; they run in lock step with one 32*TW-thread barrier per row.
//   WTA warps  (TW .. 2TW-1) : one per column; take S from the ring and do the winner-take-all (or store S in
//                              the first pass of MODE_HH).  They trail the path warps by up to 4 rows and fill
//                              the issue slots the path warps leave idle while they wait for each other.
// Ring hand-over uses named barriers: full[q] (path warps arrive, WTA warps sync) and empty[q] (reverse).
// ------------------------------------------------------------------------------------------------
constexpr int kSoutRing = 4;
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int N, bool UP, bool DO_WTA, bool FULL, bool CLAMP_EACH>
__global__ void __launch_bounds__(1024, 1) k_vert3(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Svol, VertGeom g,
                                                   int16_t* __restrict__ disp, uint32_t* __restrict__ disp2key,
                                                   uint2* __restrict__ xbuf, int* __restrict__ err)
{
    extern __shared__ __align__(16) uint16_t smem_v[];
    const int W1 = g.w.W1, H = g.w.H, Dp = g.w.Dp;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x, n = g.nstrips;
    const int x0 = int((long long)b * W1 / n), x1e = int((long long)(b + 1) * W1 / n);
    const int TW = x1e - x0;
    const int slots = g.twmax + 2;
    // smem: Ld[parity][dir][slot][Dp], Md[parity][dir][slot] (zero at start), Cring/Sring[kVertRing][twmax][Dp],
    //       Sout[kSoutRing][twmax][Dp]
    uint16_t* Ld = smem_v;
    uint32_t* Md = reinterpret_cast<uint32_t*>(Ld + size_t(4) * slots * Dp);
    uint16_t* Cring = reinterpret_cast<uint16_t*>(Md + ((4 * slots + 3) & ~3));
    uint16_t* Sring = Cring + size_t(kVertRing) * g.twmax * Dp;
    uint16_t* Sout = Sring + size_t(kVertRing) * g.twmax * Dp;
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(smem_v);
        const int nz = 2 * slots * Dp + 4 * slots;
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    if (w >= 2 * TW) return;
    const bool wta_role = w >= TW;
    const int j = wta_role ? w - TW : w;
    const bool active = FULL || lane * 2 * N < Dp;
    const int x = x0 + j;
    const int lo = lane * 2 * N;
    const int ringStride = g.twmax * Dp;
    const ptrdiff_t rowStride = (UP ? -1 : 1) * ptrdiff_t(W1) * Dp;
    const int ystart = UP ? H - 1 : 0;
    const int nall = 64 * TW;
    uint16_t* sout = Sout + size_t(j) * Dp + lo;
    constexpr int BAR_ROW = 1, BAR_FULL = 2, BAR_EMPTY = 2 + kSoutRing;

    if (wta_role) {
        // ================================ WTA warps ================================
        uint16_t* gSo = Svol + (size_t(ystart) * W1 + x) * Dp + lo;
        int16_t* dptr = disp + size_t(ystart) * g.w.W + x + g.w.minX1;
        uint32_t* kptr = disp2key + size_t(ystart) * g.w.W;
        const ptrdiff_t dStride = (UP ? -1 : 1) * ptrdiff_t(g.w.W);
        const int f = 100 - g.w.uniq;
        const uint32_t umagic = f > 0 ? uint32_t((1ull << 32) / uint32_t(f)) + 1u : 0u;
        for (int r = 0; r < H; r++) {
            const int q = r & (kSoutRing - 1);
            named_bar_sync(BAR_FULL + q, nall);
            uint32_t S[N];
            if (active) ld_regs<N>(sout + q * ringStride, S);
            else {
#pragma unroll
                for (int i = 0; i < N; i++) S[i] = 0xFFFFFFFFu;
            }
            if (r + kSoutRing < H) named_bar_arrive(BAR_EMPTY + q, nall);
            if (DO_WTA) {
                if (!FULL) {
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        const int k = lo + 2 * i;
                        if (k >= g.w.D) S[i] = 0xFFFFFFFFu;
                        else if (k + 1 >= g.w.D) S[i] |= 0xFFFF0000u;
                    }
                }
                int d;
                if (f > 0) d = wta_fast<N>(S, g.w, umagic, x, lane, kptr);
                else d = wta_regs<N>(S, g.w, x, lane, kptr);
                if (lane == 0) *dptr = int16_t(d);
            } else if (active) {
                st_regs<N>(gSo, S);
            }
            gSo += rowStride; dptr += dStride; kptr += dStride;
        }
        return;
    }

    // ================================ path warps ================================
    const int nrow = 32 * TW;
    const int dirA = j == 0 ? 1 : 0, dirB = 1 - dirA;
    const int slotA = dirA == 0 ? j : j + 2, slotB = dirB == 0 ? j : j + 2;
    const bool edge = !g.debug_no_exchange && ((j == 0 && b > 0) || (j == TW - 1 && j != 0 && b < n - 1));
    const int dirStride = slots * Dp, parStride = 2 * slots * Dp;
    const int nb = dirB == 0 ? b - 1 : b + 1;
    const int gen_stride = Dp / 2;
    const uint16_t* rdA[2]; uint16_t* wrA[2]; const uint16_t* rdB[2]; uint16_t* wrB[2];
    const uint32_t* mrA[2]; uint32_t* mwA[2]; const uint32_t* mrB[2]; uint32_t* mwB[2];
#pragma unroll
    for (int pz = 0; pz < 2; pz++) {
        rdA[pz] = Ld + pz * parStride + dirA * dirStride + slotA * Dp + lo;
        wrA[pz] = Ld + pz * parStride + dirA * dirStride + (j + 1) * Dp + lo;
        rdB[pz] = Ld + pz * parStride + dirB * dirStride + slotB * Dp + lo;
        wrB[pz] = Ld + pz * parStride + dirB * dirStride + (j + 1) * Dp + lo;
        mrA[pz] = Md + (pz * 2 + dirA) * slots + slotA;
        mwA[pz] = Md + (pz * 2 + dirA) * slots + j + 1;
        mrB[pz] = Md + (pz * 2 + dirB) * slots + slotB;
        mwB[pz] = Md + (pz * 2 + dirB) * slots + j + 1;
    }
    uint2* pub_base = xrec(xbuf, n, Dp, dirA, b, 0) + lane * N;
    const uint2* con_base = xrec(xbuf, n, Dp, dirB, edge ? nb : b, 0) + lane * N;
    const uint16_t* gC = Cvol + (size_t(ystart) * W1 + x) * Dp + lo;
    const uint16_t* gS = Svol + (size_t(ystart) * W1 + x) * Dp + lo;
    uint16_t* cring = Cring + size_t(j) * Dp + lo;
    uint16_t* sring = Sring + size_t(j) * Dp + lo;
    int issue_row = 0;
    auto issue = [&]() {   // one commit group per row, even past the end (keeps the wait arithmetic uniform)
        if (issue_row < H && active) {
            const int sl = (issue_row & (kVertRing - 1)) * ringStride;
            cp_async_lane<N>(cring + sl, gC);
            cp_async_lane<N>(sring + sl, gS);
        }
        cp_async_commit();
        gC += rowStride; gS += rowStride;
        issue_row++;
    };
#pragma unroll
    for (int i = 0; i < kVertRing - 1; i++) issue();
    uint32_t Lv[N];
#pragma unroll
    for (int i = 0; i < N; i++) Lv[i] = 0;
    uint32_t mv = 0;
    bool dead = false;

    auto row = [&](auto par_tag, int r) {
        constexpr int PAR = decltype(par_tag)::value;
        issue();
        cp_async_wait<kVertRing - 1>();     // this thread's copies of row r have landed (each lane reads only its own bytes)
        const int sl = (r & (kVertRing - 1)) * ringStride;
        uint32_t LA[N], LB[N], Cc[N], Sc[N];
        uint32_t mA, mB;
        if (active) { ld_regs<N>(rdA[PAR ^ 1], LA); ld_regs<N>(cring + sl, Cc); }
        else {
#pragma unroll
            for (int i = 0; i < N; i++) { LA[i] = kMaxCostX2; Cc[i] = kMaxCostX2; }
        }
        mA = *mrA[PAR ^ 1];
        path_step<N>(Cc, LA, mA, g.P1x2, g.P2x2, lane);
        if (active) st_regs<N>(wrA[PAR], LA);
        if (lane == 0) *mwA[PAR] = mA;
        if (edge && active) {
            uint2* rec = pub_base + (r & (kXbufGen - 1)) * gen_stride;
#pragma unroll
            for (int i = 0; i < N; i++) st_volatile_v2(rec + i, LA[i], uint32_t(r + 1));
        }
        path_step<N>(Cc, Lv, mv, g.P1x2, g.P2x2, lane);
        if (edge && r > 0) {
            const uint2* rec = con_base + ((r - 1) & (kXbufGen - 1)) * gen_stride;
#pragma unroll
            for (int i = 0; i < N; i++) LB[i] = kMaxCostX2;
            if (active && !dead) {
                const long long t0 = clock64();
                int spins = 0;
                while (true) {
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        uint2 v = ld_volatile_v2(rec + i);
                        LB[i] = v.x;
                        ok = ok && v.y == uint32_t(r);
                    }
                    if (ok) break;
                    if ((++spins & 255) == 0 && (clock64() - t0 > g.spin_limit || *reinterpret_cast<volatile int*>(err))) {
                        atomicExch(err, 1);
                        dead = true;
                        break;
                    }
                }
            }
            dead = __any_sync(kFullMask, dead);
            mB = warp_min16x2<N>(LB);
        } else {
            if (active) ld_regs<N>(rdB[PAR ^ 1], LB);
            else {
#pragma unroll
                for (int i = 0; i < N; i++) LB[i] = kMaxCostX2;
            }
            mB = *mrB[PAR ^ 1];
        }
        path_step<N>(Cc, LB, mB, g.P1x2, g.P2x2, lane);
        if (active) st_regs<N>(wrB[PAR], LB);
        if (lane == 0) *mwB[PAR] = mB;
        if (active) ld_regs<N>(sring + sl, Sc);
        else {
#pragma unroll
            for (int i = 0; i < N; i++) Sc[i] = 0;
        }
        uint32_t S[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (CLAMP_EACH) {
                uint32_t t = __vminu2(Sc[i] + Lv[i], kMaxCostX2);
                t = __vminu2(t + LA[i], kMaxCostX2);
                S[i] = __vminu2(t + LB[i], kMaxCostX2);
            } else {
                S[i] = __vminu2(Sc[i] + Lv[i] + LA[i] + LB[i], kMaxCostX2);
            }
        }
        const int q = r & (kSoutRing - 1);
        if (r >= kSoutRing) named_bar_sync(BAR_EMPTY + q, nall);    // the WTA warps have taken row r - kSoutRing
        if (active) st_regs<N>(sout + q * ringStride, S);
        named_bar_arrive(BAR_FULL + q, nall);
        named_bar_sync(BAR_ROW, nrow);
    };
    int r = 0;
    for (; r + 1 < H; r += 2) {
        row(std::integral_constant<int, 0>{}, r);
        row(std::integral_constant<int, 1>{}, r + 1);
    }
    if (r < H) row(std::integral_constant<int, 0>{}, r);
    cp_async_wait<0>();
}

}  // namespace b200sgm
