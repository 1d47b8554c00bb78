// Fused aggregation kernels (A.5 + A.6): the fast path.
//
//   k_horiz : one warp per image row computes S_h = L-> + L<- with ONE write of the S_h volume and no
//             read-modify-write.  Phase A walks the <- chain over the row and keeps only a checkpoint of the
//             (normalised) chain state every kHT columns.  Phase B walks the -> chain; one tile ahead of it the
//             <- chain is recomputed from its checkpoint into a shared-memory tile, so that when -> reaches a
//             column both terms are on chip.  Costs one extra chain step per cell, saves the parked-L write, its
//             read-back and the second write of the S_h volume (3 of the 7 volume passes of the two-ended scheme).
//   k_vert  : one co-resident CTA per column strip sweeps the rows top-down (or bottom-up for the second
//             pass of MODE_HH), one warp per column.  The vertical path lives in registers, the two
//             diagonal paths move between neighbouring warps through shared memory and between
//             neighbouring CTAs through a small global exchange buffer of {data, tag} records.
//             A strip publishes its outgoing diagonal FIRST and consumes its incoming diagonal LAST in
//             each row, so the exchange latency hides behind the row's own work.  The summed cost feeds the
//             winner-take-all directly from registers (no S volume in MODE_SGBM); the WTA of row r is issued
//             inside row r+1 so that its dependency chain overlaps the next row's path updates.
#pragma once
#include <type_traits>
#include "sgm_types.h"
#include "k_path.cuh"
#include "k_wta.cuh"

namespace b200sgm {

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc)
{
    const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
// copies this lane's N words (2N costs)
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
constexpr int kHT = 8;        // tile width = checkpoint spacing of the <- chain
constexpr int kHRing = 6;     // steps of C in flight per chain (cp.async ring in shared memory)

inline size_t horiz_ckpt_elems(int W1, int H, int Dp) { return size_t(H) * ((W1 + kHT - 1) / kHT) * Dp; }
inline size_t horiz_smem_per_warp(int Dp) { return size_t(2 * kHRing + 2 * kHT) * Dp * sizeof(uint16_t); }

// dynamic smem per warp: [2 streams][kHRing][Dp] C ring, then [2][kHT][Dp] tiles of L<-
template <int N>
__global__ void __launch_bounds__(64) k_horiz(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Sh,
                                              uint16_t* __restrict__ ckpt, int W1, int H, int Dp, int P1, int P2)
{
    extern __shared__ __align__(16) uint16_t smem_h[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int y = blockIdx.x * (blockDim.x >> 5) + wib;
    if (y >= H) return;
    const LaneCtx lc = make_lane_ctx<N>(lane, Dp, P1, P2);
    const bool active = lc.active;
    const int lo = lane * 2 * N;
    const int nT = (W1 + kHT - 1) / kHT;
    const uint16_t* Crow = Cvol + size_t(y) * W1 * Dp + lo;
    uint16_t* Srow = Sh + size_t(y) * W1 * Dp + lo;
    uint16_t* ck = ckpt + size_t(y) * nT * Dp + lo;
    uint16_t* ring = smem_h + size_t(wib) * (2 * kHRing + 2 * kHT) * Dp + lo;
    uint16_t* tiles = ring + size_t(2 * kHRing) * Dp;
    auto slot = [&](int stream, int s) { return ring + size_t(stream * kHRing + (s % kHRing)) * Dp; };
    constexpr int PF = kHRing - 1;

    uint32_t Lt[N], Ln[N], Cc[N];
    // ---------------- phase A: <- chain from x = W1-1 down to kHT, checkpoints only ----------------
    {
        const int nA = W1 - kHT;                                  // steps; step s visits x = W1-1-s
        auto issue = [&](int s) {
            if (s < nA && active) cp_async_lane<N>(slot(0, s), Crow + size_t(W1 - 1 - s) * Dp);
            cp_async_commit();
        };
        if (nA > 0) {
            for (int s = 0; s < PF; s++) issue(s);
#pragma unroll
            for (int j = 0; j < N; j++) Lt[j] = 0;
            for (int s = 0; s < nA; s++) {
                issue(s + PF);
                cp_async_wait<PF>();
                if (active) ld_regs<N>(slot(0, s), Cc);
                else {
#pragma unroll
                    for (int j = 0; j < N; j++) Cc[j] = kMaxCostX2;
                }
                path_step<N>(Cc, Lt, Ln, lc);
                const int x = W1 - 1 - s;
                if ((x & (kHT - 1)) == 0 && active) st_regs<N>(ck + size_t(x / kHT - 1) * Dp, Lt);   // entry state of tile x/kHT - 1
            }
            cp_async_wait<0>();
        }
    }
    __syncwarp();
    // ---------------- phase B: -> chain; <- recomputed one tile ahead ----------------
    // Step index u = t*kHT + i.  Stream 0 (->) visits x = u.  Stream 1 (<-) runs one tile ahead: at global step
    // u it recomputes column xb(u + kHT) where xb(v) = (v/kHT)*kHT + kHT-1 - (v%kHT); columns >= W1 are skipped.
    auto xb_of = [&](int v) { return (v / kHT) * kHT + (kHT - 1) - (v % kHT); };
    const int nB = nT * kHT;
    auto issue0 = [&](int u) { if (u < W1 && active) cp_async_lane<N>(slot(0, u), Crow + size_t(u) * Dp); };
    auto issue1 = [&](int v) {
        if (v < nB && active) { const int x = xb_of(v); if (x < W1) cp_async_lane<N>(slot(1, v), Crow + size_t(x) * Dp); }
    };
    uint32_t Ltb[N];                  // <- state
    uint32_t Lck[N];                  // prefetched checkpoint of the next tile to recompute
#pragma unroll
    for (int j = 0; j < N; j++) { Lt[j] = 0; Ltb[j] = 0; Lck[j] = 0; }
    // prologue: recompute tile 0 (stream 1 steps v = 0 .. kHT-1)
    for (int s = 0; s < PF; s++) { issue1(s); cp_async_commit(); }
    if (nT > 1 && active) ld_regs<N>(ck, Ltb);                     // entry state of tile 0 (zeros if it is the last tile)
    if (nT > 2 && active) ld_regs<N>(ck + Dp, Lck);                // (plain loads: written by this lane in phase A)
    for (int v = 0; v < kHT; v++) {
        issue1(v + PF); cp_async_commit();
        cp_async_wait<PF>();
        const int x = xb_of(v);
        if (x < W1) {
            if (active) ld_regs<N>(slot(1, v), Cc);
            else {
#pragma unroll
                for (int j = 0; j < N; j++) Cc[j] = kMaxCostX2;
            }
            path_step<N>(Cc, Ltb, Ln, lc);
            if (active) st_regs<N>(tiles + size_t(x & (kHT - 1)) * Dp, Ln);
        }
    }
    // stream 0 joins: its first PF steps are issued now; from here on every iteration commits one group holding
    // one step of each stream
    for (int s = 0; s < PF; s++) { issue0(s); cp_async_commit(); }
    for (int t = 0; t < nT; t++) {
        // entry state of tile t+1: its checkpoint, or zeros when t+1 is the last tile
        const bool more = t + 1 < nT;
#pragma unroll
        for (int j = 0; j < N; j++) Ltb[j] = (t + 2 < nT) ? Lck[j] : 0u;
        if (t + 3 < nT && active) ld_regs<N>(ck + size_t(t + 2) * Dp, Lck);
        uint16_t* tile_cur = tiles + size_t(t & 1) * kHT * Dp;
        uint16_t* tile_nxt = tiles + size_t((t + 1) & 1) * kHT * Dp;
#pragma unroll 2
        for (int i = 0; i < kHT; i++) {
            const int u = t * kHT + i;
            issue0(u + PF); issue1(u + kHT + PF); cp_async_commit();
            // the PF newest groups carry stream-0 steps u+1 .. u+PF and stream-1 steps u+kHT+1 .. u+kHT+PF;
            // everything older (in particular step u of stream 0 and step u+kHT of stream 1) has landed
            cp_async_wait<PF>();
            const int xb = xb_of(u + kHT);
            uint32_t Cb[N];
            const bool do_b = more && xb < W1;
            if (do_b) {
                if (active) ld_regs<N>(slot(1, u + kHT), Cb);
                else {
#pragma unroll
                    for (int j = 0; j < N; j++) Cb[j] = kMaxCostX2;
                }
            }
            if (u < W1) {
                if (active) ld_regs<N>(slot(0, u), Cc);
                else {
#pragma unroll
                    for (int j = 0; j < N; j++) Cc[j] = kMaxCostX2;
                }
                uint32_t Lb[N];
                if (active) ld_regs<N>(tile_cur + size_t(i) * Dp, Lb);
                path_step<N>(Cc, Lt, Ln, lc);
                if (active) {
#pragma unroll
                    for (int j = 0; j < N; j++) Ln[j] = __vminu2(Ln[j] + Lb[j], kMaxCostX2);
                    st_regs<N>(Srow + size_t(u) * Dp, Ln);
                }
            }
            if (do_b) {
                uint32_t Lnb[N];
                path_step<N>(Cb, Ltb, Lnb, lc);
                if (active) st_regs<N>(tile_nxt + size_t(xb & (kHT - 1)) * Dp, Lnb);
            }
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Register/shared-memory WTA (A.6) used by the vertical sweep.  S must already hold 0xFFFF in cells >= D and in
// inactive lanes.  `scratch` is this warp's private Dp-cell staging area (paired layout).
// ------------------------------------------------------------------------------------------------
struct WtaCtx {
    uint32_t kk0;        // (lane*N) | (lane*N + Dh) << 16 : disparity indices of this lane's first word
    uint32_t umagic;     // floor(2^32 / f) + 1 with f = 100 - uniq > 0
    int f;
    int Dh;
};

template <int N>
__device__ __forceinline__ int wta_staged(const uint32_t (&S)[N], uint16_t* __restrict__ scratch, const WtaGeom& g,
                                          const WtaCtx& w, int x1, int lane, bool active, uint32_t* __restrict__ disp2key_row)
{
    // key = S << 16 | k: the warp minimum is the smallest cost and, among equals, the FIRST disparity
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const uint32_t kk = w.kk0 + uint32_t(j) * 0x10001u;
        key = __vimin3_u32(key, __byte_perm(kk, S[j], 0x5410), __byte_perm(kk, S[j], 0x7632));
    }
    if (active) st_regs<N>(scratch + lane * 2 * N, S);
    key = __reduce_min_sync(kFullMask, key);
    const int minS = int(key >> 16), best = int(key & 0xFFFFu);
    // sub-pixel neighbours (uniform addresses -> broadcast loads); garbage when best is at an end (unused then)
    const int km = max(best - 1, 0), kp = min(best + 1, g.Dp - 1);
    const int sm = scratch[cell_u16_index(km, w.Dh)], sp = scratch[cell_u16_index(kp, w.Dh)];
    __syncwarp();
    // uniqueness: S[k] * f < minS * 100  <=>  S[k] < ceil(minS*100 / f), cells best-1 .. best+1 exempt: lanes 0..2
    // overwrite those three cells with 0xFFFF in the staged copy, then every lane re-reads its words
    {
        const int k = best - 1 + lane;
        if (lane < 3 && k >= 0 && k < g.Dp) scratch[cell_u16_index(k, w.Dh)] = 0xFFFFu;
    }
    __syncwarp();
    uint32_t T[N];
    if (active) ld_regs<N>(scratch + lane * 2 * N, T);
    else {
#pragma unroll
        for (int j = 0; j < N; j++) T[j] = 0xFFFFFFFFu;
    }
    uint32_t mm = T[0];
#pragma unroll
    for (int j = 1; j < N; j++) mm = __vminu2(mm, T[j]);
    const uint32_t thr = min(__umulhi(uint32_t(minS * 100 + w.f - 1), w.umagic), 0xFFFFu);
    const bool bad = min(mm & 0xFFFFu, mm >> 16) < thr;
    const bool reject = __any_sync(kFullMask, bad) || minS >= kMaxCost;
    int dfix = best * 16;
    if (best > 0 && best < g.D - 1) {
        const int den = max(sm + sp - 2 * minS, 1);
        // |quotient| <= 8.5 and numerator, denominator < 2^24: IEEE float division then truncation is exact
        dfix += __float2int_rz(__fdiv_rn(float((sm - sp) * 16 + den), float(den * 2)));
    }
    if (lane == 0 && !reject) {
        const int x = x1 + g.minX1;
        const int x2 = x - best - g.minD;
        if (x2 >= 0 && x2 < g.W) atomicMin(disp2key_row + x2, (uint32_t(minS) << 16) | uint32_t(0xFFFF - x));
    }
    return reject ? g.INVALID : dfix + g.minD * 16;
}

// Exact fallback for uniquenessRatio >= 100 (f <= 0: the comparison cannot be turned into a threshold).
template <int N>
__device__ __forceinline__ int wta_regs_slow(const uint32_t (&S)[N], uint16_t* __restrict__ scratch, const WtaGeom& g,
                                             const WtaCtx& w, int x1, int lane, bool active, uint32_t* __restrict__ disp2key_row)
{
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const uint32_t kk = w.kk0 + uint32_t(j) * 0x10001u;
        key = __vimin3_u32(key, __byte_perm(kk, S[j], 0x5410), __byte_perm(kk, S[j], 0x7632));
    }
    if (active) st_regs<N>(scratch + lane * 2 * N, S);
    key = __reduce_min_sync(kFullMask, key);
    __syncwarp();
    const int minS = int(key >> 16), best = int(key & 0xFFFFu);
    int out = g.INVALID;
    if (minS < kMaxCost) {
        const int Tt = minS * 100;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int k = lane * N + j;
            const int s0 = int(S[j] & 0xFFFFu), s1 = int(S[j] >> 16);
            if (active && k < g.D && abs(k - best) > 1 && s0 * w.f < Tt) bad = true;
            if (active && k + w.Dh < g.D && abs(k + w.Dh - best) > 1 && s1 * w.f < Tt) bad = true;
        }
        if (!__any_sync(kFullMask, bad)) {
            int dfix = best * 16;
            if (best > 0 && best < g.D - 1) {
                const int sm = scratch[cell_u16_index(best - 1, w.Dh)], sp = scratch[cell_u16_index(best + 1, w.Dh)];
                const int den = max(sm + sp - 2 * minS, 1);
                dfix += ((sm - sp) * 16 + den) / (den * 2);
            }
            if (lane == 0) {
                const int x = x1 + g.minX1;
                const int x2 = x - best - g.minD;
                if (x2 >= 0 && x2 < g.W) atomicMin(disp2key_row + x2, (uint32_t(minS) << 16) | uint32_t(0xFFFF - x));
            }
            out = dfix + g.minD * 16;
        }
    }
    __syncwarp();
    return out;
}

// ------------------------------------------------------------------------------------------------
// Vertical sweep.
// ------------------------------------------------------------------------------------------------
struct VertGeom {
    WtaGeom w;
    int nstrips;
    int twmax;            // warps per CTA = widest strip
    int P1, P2;
    long long spin_limit; // clock64 ticks before a record wait gives up
};

// Inter-strip exchange, "low latency" protocol: a record is Dp/2 entries of {two packed costs, tag}; each
// entry is one 8-byte volatile store, so data and flag arrive together and no fence or separate flag is
// needed.  tag = row + 1 (the buffer is zeroed before the launch).  [side][strip][row & 3][Dp/2] uint2.
// Four generations are live at once: a strip publishes row R at the START of its row R, having only
// waited (end of row R-1) for the neighbour's record R-2 -- and that neighbour still reads our record
// R-3 at the END of its row R-2.
constexpr int kXbufGen = 4;
constexpr int kVRing = 4;     // rows of C / S_h in flight per column (cp.async ring in shared memory)
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

// dynamic smem: Ld[2 parity][2 dir][twmax+2][Dp] | ring[kVRing][2 (C,S_h)][twmax][Dp] | scratch[twmax][Dp]   (uint16)
inline size_t vert_smem_bytes(int twmax, int Dp)
{
    return (size_t(4) * (twmax + 2) + size_t(2 * kVRing) * twmax + size_t(twmax)) * Dp * sizeof(uint16_t);
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
    // Ld holds the NORMALISED diagonal states; zero = "predecessor outside the image" (first row, border columns)
    uint16_t* Ld = smem_v;
    uint16_t* ringbase = Ld + size_t(4) * slots * Dp;
    uint16_t* scratchbase = ringbase + size_t(2 * kVRing) * g.twmax * Dp;
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(smem_v);
        const int nz = 2 * slots * Dp;
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    if (j >= TW) return;                 // the row barrier below only counts the column warps
    const int nbar = 32 * TW;
    const LaneCtx lc = make_lane_ctx<N>(lane, Dp, g.P1, g.P2);
    const bool active = FULL || lc.active;
    const int x = x0 + j;
    const int lo = lane * 2 * N;
    // direction 0: predecessor column x-1 (slot j); direction 1: predecessor column x+1 (slot j+2).
    // The left-edge warp runs direction 1 first (it publishes it), every other warp direction 0 first.
    const int dirA = j == 0 ? 1 : 0, dirB = 1 - dirA;
    const int slotA = dirA == 0 ? j : j + 2, slotB = dirB == 0 ? j : j + 2;
    const bool edge = (j == 0 && b > 0) || (j == TW - 1 && j != 0 && b < n - 1);   // publishes dirA, consumes dirB
    const int dirStride = slots * Dp, parStride = 2 * slots * Dp;
    const uint16_t* rdA[2]; uint16_t* wrA[2]; const uint16_t* rdB[2]; uint16_t* wrB[2];
#pragma unroll
    for (int pz = 0; pz < 2; pz++) {
        rdA[pz] = Ld + pz * parStride + dirA * dirStride + slotA * Dp + lo;
        wrA[pz] = Ld + pz * parStride + dirA * dirStride + (j + 1) * Dp + lo;
        rdB[pz] = Ld + pz * parStride + dirB * dirStride + slotB * Dp + lo;
        wrB[pz] = Ld + pz * parStride + dirB * dirStride + (j + 1) * Dp + lo;
    }
    // exchange records: I publish side dirA of my strip, I consume side dirB of the neighbour
    const int nb = dirB == 0 ? b - 1 : b + 1;
    uint2* pub_base = xrec(xbuf, n, Dp, dirA, b, 0) + lane * N;
    const uint2* con_base = xrec(xbuf, n, Dp, dirB, edge ? nb : b, 0) + lane * N;
    const int gen_stride = Dp / 2;

    const ptrdiff_t rowStride = (UP ? -1 : 1) * ptrdiff_t(W1) * Dp;
    const int ystart = UP ? H - 1 : 0;
    const uint16_t* gC = Cvol + (size_t(ystart) * W1 + x) * Dp + lo;      // next row to fetch
    const uint16_t* gSin = Svol + (size_t(ystart) * W1 + x) * Dp + lo;
    uint16_t* gSout = Svol + (size_t(ystart) * W1 + x) * Dp + lo;          // row being computed (first pass of MODE_HH)
    int16_t* dptr = disp + size_t(ystart) * g.w.W + x + g.w.minX1;         // row of the pending WTA
    uint32_t* kptr = disp2key + size_t(ystart) * g.w.W;
    const ptrdiff_t dStride = (UP ? -1 : 1) * ptrdiff_t(g.w.W);
    WtaCtx wc;
    wc.Dh = Dp >> 1;
    wc.kk0 = uint32_t(lane * N) | (uint32_t(lane * N + wc.Dh) << 16);
    wc.f = 100 - g.w.uniq;
    wc.umagic = wc.f > 0 ? uint32_t((1ull << 32) / uint32_t(wc.f)) + 1u : 0u;
    uint16_t* scratch = scratchbase + size_t(j) * Dp;
    // cp.async ring: [slot][array][warp][Dp]
    uint16_t* ring = ringbase + size_t(j) * Dp + lo;
    const int ringArr = g.twmax * Dp, ringSlot = 2 * g.twmax * Dp;
    int issue_row = 0;
    auto issue = [&]() {   // one commit group per row, even past the end (keeps the wait arithmetic uniform)
        if (issue_row < H && active) {
            uint16_t* dst = ring + (issue_row & (kVRing - 1)) * ringSlot;
            cp_async_lane<N>(dst, gC);
            cp_async_lane<N>(dst + ringArr, gSin);
        }
        cp_async_commit();
        gC += rowStride; gSin += rowStride;
        issue_row++;
    };
#pragma unroll
    for (int i = 0; i < kVRing - 1; i++) issue();

    uint32_t LtV[N], Sprev[N];
#pragma unroll
    for (int q = 0; q < N; q++) { LtV[q] = 0; Sprev[q] = 0xFFFFFFFFu; }
    bool dead = false;

    auto do_wta = [&](int xcol) {
        int d;
        if (wc.f > 0) d = wta_staged<N>(Sprev, scratch, g.w, wc, xcol, lane, active, kptr);
        else d = wta_regs_slow<N>(Sprev, scratch, g.w, wc, xcol, lane, active, kptr);
        if (lane == 0) *dptr = int16_t(d);
        dptr += dStride; kptr += dStride;
    };

    // one row; PAR = parity of r (buffer written), reads the other one
    auto row = [&](auto par_tag, int r) {
        constexpr int PAR = decltype(par_tag)::value;
        issue();
        cp_async_wait<kVRing - 1>();     // this thread's copies of row r have landed (each lane reads only its own bytes)
        const uint16_t* rs = ring + (r & (kVRing - 1)) * ringSlot;
        uint32_t Cc[N], Sc[N], LtA[N], LtB[N], LnA[N], LnV[N], LnB[N];
        if (active) { ld_regs<N>(rs, Cc); ld_regs<N>(rs + ringArr, Sc); ld_regs<N>(rdA[PAR ^ 1], LtA); }
        else {
#pragma unroll
            for (int q = 0; q < N; q++) { Cc[q] = kMaxCostX2; Sc[q] = 0; LtA[q] = 0; }
        }
        // ---- step A (the direction this warp publishes)
        path_step<N>(Cc, LtA, LnA, lc);
        if (active) st_regs<N>(wrA[PAR], LtA);
        if (edge && active) {
            uint2* rec = pub_base + (r & (kXbufGen - 1)) * gen_stride;
#pragma unroll
            for (int q = 0; q < N; q++) st_volatile_v2(rec + q, LtA[q], uint32_t(r + 1));
        }
        // ---- vertical path: registers only
        path_step<N>(Cc, LtV, LnV, lc);
        // ---- WTA of the previous row: independent work that overlaps the steps around it
        if (DO_WTA && r > 0) do_wta(x);
        // ---- step B
        if (edge && r > 0) {
            const uint2* rec = con_base + ((r - 1) & (kXbufGen - 1)) * gen_stride;
#pragma unroll
            for (int q = 0; q < N; q++) LtB[q] = 0;
            if (active && !dead) {
                const long long t0 = clock64();
                int spins = 0;
                while (true) {
                    bool ok = true;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        uint2 v = ld_volatile_v2(rec + q);
                        LtB[q] = v.x;
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
        } else {
            if (active) ld_regs<N>(rdB[PAR ^ 1], LtB);
            else {
#pragma unroll
                for (int q = 0; q < N; q++) LtB[q] = 0;
            }
        }
        path_step<N>(Cc, LtB, LnB, lc);
        if (active) st_regs<N>(wrB[PAR], LtB);
        // ---- S = sat(S_h + L_v + L_A + L_B)
#pragma unroll
        for (int q = 0; q < N; q++) {
            if (CLAMP_EACH) {
                uint32_t t = __vminu2(Sc[q] + LnV[q], kMaxCostX2);
                t = __vminu2(t + LnA[q], kMaxCostX2);
                Sprev[q] = __vminu2(t + LnB[q], kMaxCostX2);
            } else {
                Sprev[q] = __vminu2(Sc[q] + LnV[q] + LnA[q] + LnB[q], kMaxCostX2);
            }
        }
        if (DO_WTA) {
            if (!FULL) {
#pragma unroll
                for (int q = 0; q < N; q++) {   // cells beyond D never win and never veto
                    const int k = lane * N + q;
                    if (!active || k >= g.w.D) Sprev[q] = 0xFFFFFFFFu;
                    else if (k + wc.Dh >= g.w.D) Sprev[q] |= 0xFFFF0000u;
                }
            }
        } else {
            if (active) st_regs<N>(gSout, Sprev);
            gSout += rowStride;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");
    };

    int r = 0;
    for (; r + 1 < H; r += 2) {
        row(std::integral_constant<int, 0>{}, r);
        row(std::integral_constant<int, 1>{}, r + 1);
    }
    if (r < H) row(std::integral_constant<int, 0>{}, r);
    if (DO_WTA) do_wta(x);
    cp_async_wait<0>();
}

}  // namespace b200sgm
