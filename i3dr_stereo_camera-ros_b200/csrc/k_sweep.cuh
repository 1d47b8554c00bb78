// Vertical sweep (A.5 paths "down", "down-right", "down-left" -- or their mirror images for the second pass of
// MODE_HH -- summed with S_h, fused with the winner-take-all A.6): the diagonal-skewed, one-directional version.
//
// The W1 valid columns are covered by W1 CHAINS on a cylinder: chain c sits at column x = (c + r) mod W1 in sweep
// row r, i.e. it walks down a down-right diagonal and re-enters at column 0 (as a fresh diagonal) after leaving at
// column W1-1.  With that labelling the three predecessors of cell (x, r) are
//     down-right path : (x-1, r-1) = chain c   itself            -> state stays in the warp's registers
//     down path       : (x,   r-1) = chain c+1 at row r-1        -> from the next warp, previous row
//     down-left path  : (x+1, r-1) = chain c+2 at row r-1        -> from the warp after that, previous row
// so every dependency points the SAME way (towards higher chain numbers, cyclically).  A strip of consecutive chains
// = one CTA, one warp per chain; per row the strip needs three records from the strip to its right (its first chain's
// down / down-left states and its second chain's down-left state) and nothing from the strip to its left.  Compared
// with plain column strips (which need both neighbours every row):
//   * a strip may run up to kXK rows BEHIND its right neighbour (one-way coupling with slack instead of a symmetric
//     lock step of all strips), and the records it needs are usually there long before it asks;
//   * one agent warp per CTA (instead of two that also compute) prefetches the records a row ahead into shared memory;
//   * one diagonal path never leaves the register file (no shared-memory hand-over, no second parity buffer for it).
// Inside a strip the path warps advance in lock step: one named barrier per row orders the hand-over of the down /
// down-left states (named barriers carry the memory ordering; shared-memory flags polled by the consumers were tried
// and are both slower and -- without a membar between a multi-wavefront STS.128 and the flag store -- unsafe).
// Image borders are predicates: at x == 0 the down-right predecessor is outside (state 0), at x == W1-1 the
// down-left one is; row 0 starts from all-zero states.
#pragma once
#include <type_traits>
#include "sgm_types.h"
#include "k_path.cuh"
#include "k_wta.cuh"
#include "k_fused.cuh"

namespace b200sgm {

constexpr int kXK = 16;          // generations of inter-strip records in global memory
constexpr int kXR = 2;           // record rows staged in shared memory by the agent warp
constexpr int kSweepMaxTW = 15;  // chains (= path warps) per strip: 15 path + 15 WTA + 1 agent = 31 warps

// Threads per CTA the register file is budgeted for: 64 registers per thread up to 256 disparities, 72 / 128 / 255 beyond.
constexpr int sweep_max_threads(int n) { return n <= 4 ? 1024 : n == 8 ? 896 : n == 16 ? 512 : 256; }

struct SweepGeom {
    WtaGeom w;
    int nstrips;
    int twmax;            // chains of the widest strip
    int P1, P2;
    long long spin_limit; // clock64 ticks before a record wait gives up
    int debug_flags;      // 8: timing experiment, ignore record tags (wrong results)
    long long* trace;     // development: clock64 stamps of strip `trace_strip`, rows trace_row0 .. +31 (nullptr = off)
    int trace_strip, trace_row0;
};
constexpr int kTracePoints = 4;

// xbuf: [0, 4 KB) progress words (one int per strip), then records [3 kinds][nstrips][kXK][Dp/2] of {data, tag}
inline size_t sweep_xbuf_bytes(int nstrips, int Dp) { return 4096 + size_t(3) * nstrips * kXK * (Dp / 2) * sizeof(uint2); }
// dynamic smem (uint16): Ld[2 parity][2 kinds][tw][Dp] | Cring[tw][RING][Dp] | Sring[tw][RING][Dp] | stage[kStage][tw][Dp] | xring[kXR][3][Dp]
inline size_t sweep_smem_bytes(int tw, int Dp, int ring, bool wta)
{
    return (size_t(4) * tw + size_t(2) * ring * tw + (wta ? size_t(kStage) * tw : 0) + size_t(kXR) * 3) * Dp * sizeof(uint16_t);
}

__device__ __forceinline__ uint2* sweep_rec(uint2* xbuf, int nstrips, int Dp, int kind, int strip, int gen)
{
    return xbuf + 512 + (size_t((kind * nstrips + strip) * kXK + gen)) * (Dp / 2);
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Spins until ready() holds.  A wait that exceeds the limit, or that sees the error word set by somebody else, raises the
// error word and marks this warp dead: it then stops waiting altogether (the frame is reported as failed by the host).
template <typename F>
__device__ __forceinline__ void sweep_wait(F ready, long long limit, int* err, bool& dead)
{
    if (dead || ready()) return;
    const long long t0 = clock64();
    int spins = 0;
    while (!ready()) {
        if ((++spins & 63) == 0 && (clock64() - t0 > limit || *reinterpret_cast<volatile int*>(err))) {
            atomicExch(err, 1);
            dead = true;
            break;
        }
    }
}

// Scalar part of the WTA for up to 32 sweep rows of chain c starting at sweep row rb: lane l resolves row rb + l.
template <bool UP>
__device__ __forceinline__ void wta_flush_diag(const WtaAcc& acc, int cnt, const WtaGeom& g, const WtaCtx& w, int c, int rb, int lane,
                                               int16_t* __restrict__ disp, uint32_t* __restrict__ disp2key)
{
    if (lane >= cnt) return;
    const int r = rb + lane;
    const int x1 = (c + r) % g.W1;
    const int y = UP ? g.H - 1 - r : r;
    const int minS = int(acc.key >> 16), best = int(acc.key & 0xFFFFu);
    const uint32_t n100 = uint32_t(minS * 100 + w.f - 1);
    const uint32_t thr = min(w.f == 1 ? n100 : __umulhi(n100, w.umagic), 0xFFFFu);
    const bool reject = (acc.mm & 0xFFFFu) < thr || minS >= kMaxCost;
    int dfix = best * 16;
    if (best > 0 && best < g.D - 1) {
        const int den = max(acc.sm + acc.sp - 2 * minS, 1);
        dfix += __float2int_rz(__fdiv_rn(float((acc.sm - acc.sp) * 16 + den), float(den * 2)));
    }
    const int x = x1 + g.minX1;
    if (!reject) {
        const int x2 = x - best - g.minD;
        if (x2 >= 0 && x2 < g.W) atomicMin(disp2key + size_t(y) * g.W + x2, (uint32_t(minS) << 16) | uint32_t(0xFFFF - x));
    }
    disp[size_t(y) * g.W + x] = int16_t(reject ? g.INVALID : dfix + g.minD * 16);
}

// Launch: nstrips CTAs (all co-resident: cooperative launch), 32 * (2 * twmax + 1) threads (32 * (twmax + 1) when !DO_WTA):
//   path warps  [0, twmax)         one per chain: the three path updates of a row and the sum S; lock step, one barrier per row
//   WTA warps   [twmax, 2 twmax)   one per chain: resolve parked rows of S two at a time (A.6), trailing by <= kStage rows
//   agent warp  (last)             polls the right neighbour strip's three records of each row in global memory, one row ahead of
//                                  their use, and drops them into shared memory for the strip's last two chains
// RING : rows of C / S_h in flight per chain (cp.async rings)
// FULL : Dp == D == 64 N;  CLAMP_EACH : saturate after every addition of the sum
template <int N, int RING, bool UP, bool DO_WTA, bool FULL, bool CLAMP_EACH>
__global__ void __launch_bounds__(sweep_max_threads(N), 1) k_sweep(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Svol, SweepGeom g,
                                                                    int16_t* __restrict__ disp, uint32_t* __restrict__ disp2key,
                                                                    uint2* __restrict__ xbuf, int* __restrict__ err)
{
    static_assert(RING == 4 || RING == 8, "ring depth");
    static_assert(kXR == 2 && kStage == 4, "slot indices are immediates of the 4-row unrolled loop");
    extern __shared__ __align__(16) uint16_t smem_s[];
    const int W1 = g.w.W1, H = g.w.H, Dp = FULL ? 64 * N : g.w.Dp;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x, n = g.nstrips, tw = g.twmax;
    const int c0 = int((long long)b * W1 / n), c1 = int((long long)(b + 1) * W1 / n);
    const int TW = c1 - c0;                                  // 2 <= TW <= tw
    // Ld[parity][kind][slot][Dp]: slot s holds what path warp s READS: kind 0 = down state of chain c0+s+1, kind 1 = down-left
    // state of chain c0+s+2 (normalised).  Warp j therefore writes kind 0 into slot j-1 and kind 1 into slot j-2; what the
    // last two warps read beyond the strip comes from the agent's xring[slot][record][Dp] (records: V0, B0, B1).
    uint16_t* Ld = smem_s;
    const int kindStride = tw * Dp, parStride = 2 * tw * Dp;
    uint16_t* ringbase = Ld + size_t(2) * parStride;
    uint16_t* sringbase = ringbase + size_t(RING) * tw * Dp;
    uint16_t* stagebase = sringbase + size_t(RING) * tw * Dp;
    uint16_t* xring = stagebase + (DO_WTA ? size_t(kStage) * tw * Dp : 0);
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(Ld);
        const int nz = parStride;                      // both parities, in 32-bit words
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    if (g.trace != nullptr && threadIdx.x == 0) g.trace[4096 + 2 * b] = clock64();     // per-strip begin / end (end: path warp 0)
    const LaneCtx lc = make_lane_ctx<N>(lane, Dp, g.P1, g.P2);
    const bool active = FULL || lc.active;
    const int lo = lane * 2 * N;
    const int agent_w = DO_WTA ? 2 * tw : tw;
    bool dead = false;
    // named barriers: the agent and the strip's last two path warps hand over through XFULL / XEMPTY (96 threads)
    constexpr int BAR_ROW = 1, BAR_FULL = 2, BAR_EMPTY = 2 + kStage, BAR_XFULL = 2 + 2 * kStage, BAR_XEMPTY = BAR_XFULL + kXR;
    static_assert(BAR_XEMPTY + kXR <= 16, "16 named barriers per CTA");
    long long seg_acc[kTracePoints] = {0, 0, 0, 0}, seg_last = 0;      // trace_row0 == -2: time spent per segment over the whole sweep
    auto stamp = [&](int r, int k) {
        if (g.trace != nullptr && b == g.trace_strip && lane == 0) {
            if (g.trace_row0 == -2) {
                const long long now = clock64();
                if (seg_last) seg_acc[k] += now - seg_last;
                seg_last = now;
                if (r >= H - 2) g.trace[w * kTracePoints + k] = seg_acc[k];
            } else if (g.trace_row0 < 0) {                      // whole-sweep mode: row starts of path warp 0 and of the agent
                if (k == 0 && r < 2048 && (w == 0 || w == agent_w)) g.trace[(w == 0 ? 0 : 2048) + r] = clock64();
            } else if (unsigned(r - g.trace_row0) < 32u)
                g.trace[(size_t(w) * 32 + (r - g.trace_row0)) * kTracePoints + k] = clock64();
        }
    };

    if (w == agent_w) {
        // ================================ agent warp ================================
        const int nb = b + 1 == n ? 0 : b + 1;
        const uint2* recV0 = sweep_rec(xbuf, n, Dp, 0, nb, 0) + lane * N;    // down state of the neighbour's first chain
        const uint2* recB0 = sweep_rec(xbuf, n, Dp, 1, nb, 0) + lane * N;    // its down-left state
        const uint2* recB1 = sweep_rec(xbuf, n, Dp, 2, nb, 0) + lane * N;    // down-left state of the neighbour's second chain
        uint16_t* dst = xring + lo;
        int* prog = reinterpret_cast<int*>(xbuf) + b;
        const int recGen = Dp / 2;
        for (int rr = 0; rr + 1 < H; rr++) {
            stamp(rr, 0);
            uint32_t dv0[N], db0[N], db1[N];
#pragma unroll
            for (int q = 0; q < N; q++) { dv0[q] = 0; db0[q] = 0; db1[q] = 0; }
            const int go = (rr & (kXK - 1)) * recGen;
            const uint32_t tag = uint32_t(rr + 1);
            auto poll = [&] {
                bool ok = true;
                if (active) {
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        const uint2 a = ld_volatile_v2(recV0 + go + q), bb = ld_volatile_v2(recB0 + go + q), cc = ld_volatile_v2(recB1 + go + q);
                        dv0[q] = a.x; db0[q] = bb.x; db1[q] = cc.x;
                        ok = ok && a.y == tag && bb.y == tag && cc.y == tag;
                    }
                }
                return __all_sync(kFullMask, ok) || (g.debug_flags & 8);
            };
            sweep_wait(poll, g.spin_limit, err, dead);
            stamp(rr, 1);
            if (lane == 0) st_relaxed_gpu(prog, rr + 1);     // the producer may reuse this record generation
            // slot rr & 1 held record rr - 2 until the two consumers took it (their row rr - 1)
            if (rr >= kXR) named_bar_sync(BAR_XEMPTY + (rr & (kXR - 1)), 96);
            stamp(rr, 2);
            if (active) {
                uint16_t* d = dst + (rr & (kXR - 1)) * 3 * Dp;
                st_regs<N>(d, dv0); st_regs<N>(d + Dp, db0); st_regs<N>(d + 2 * Dp, db1);
            }
            named_bar_arrive(BAR_XFULL + (rr & (kXR - 1)), 96);
        }
        return;
    }
    const bool wta_role = DO_WTA && w >= tw;
    const int j = wta_role ? w - tw : w;
    if (j >= TW) return;
    const int c = c0 + j;                              // this warp's chain
    const int ringSlot = tw * Dp;                      // slot stride of the stage ring [slot][warp][Dp]
    uint16_t* stage = stagebase + size_t(j) * Dp;      // + slot * ringSlot
    const int nboth = 64 * TW, nrow = 32 * TW;

    if (wta_role) {
        // ================================ WTA warps ================================
        WtaCtx wc;
        wc.Dh = Dp >> 1;
        wc.kk0 = uint32_t(lane * N) | (uint32_t(lane * N + wc.Dh) << 16);
        wc.f = 100 - g.w.uniq;
        wc.umagic = wc.f > 0 ? uint32_t((1ull << 32) / uint32_t(wc.f)) + 1u : 0u;
        WtaAcc acc{0xFFFFFFFFu, 0u, 0, 0};
        int pending = 0, rb = 0;            // rows parked in acc, first of them
        auto batch = [&](auto q0_tag, int r, int cnt) {
            constexpr int Q0 = decltype(q0_tag)::value;
            named_bar_sync(BAR_FULL + Q0, nboth);
            if (cnt > 1) named_bar_sync(BAR_FULL + Q0 + 1, nboth);
            uint16_t* sc = stage + Q0 * ringSlot;
            if (wc.f > 0) {
                wta_vec<N>(sc, ringSlot, r, g.w, wc, lane, active, acc);   // a row past the end lands in a lane >= cnt of the flush
                pending += cnt;
                if (((r + kWB) & 31) == 0 || r + kWB >= H) {
                    wta_flush_diag<UP>(acc, pending, g.w, wc, c, rb, lane, disp, disp2key);
                    rb += pending;
                    pending = 0;
                }
            } else {
                __syncwarp();
                for (int q = 0; q < cnt; q++) {
                    const int rr = r + q, x1 = (c + rr) % W1, y = UP ? H - 1 - rr : rr;
                    const int d = wta_slow<N>(sc + q * ringSlot, g.w, wc, x1, lane, active, disp2key + size_t(y) * g.w.W);
                    if (lane == 0) disp[size_t(y) * g.w.W + x1 + g.w.minX1] = int16_t(d);
                }
                __syncwarp();
            }
            if (r + kStage < H) named_bar_arrive(BAR_EMPTY + Q0, nboth);
            if (cnt > 1 && r + 1 + kStage < H) named_bar_arrive(BAR_EMPTY + Q0 + 1, nboth);
        };
        static_assert(kWB == 2, "two batches per stage-ring revolution");
        int r = 0;
        for (; r + 3 < H; r += 4) {
            batch(std::integral_constant<int, 0>{}, r, 2);
            batch(std::integral_constant<int, 2>{}, r + 2, 2);
        }
        if (r < H) batch(std::integral_constant<int, 0>{}, r, min(2, H - r));
        if (r + 2 < H) batch(std::integral_constant<int, 2>{}, r + 2, 1);
        return;
    }

    // ================================ path warps ================================
    const bool producer = j <= 1;                              // the strip to the left needs my states
    const bool consumer = j >= TW - 2;                         // my down-left (and for the last chain also the down) state comes from the agent
    const int cb = b == 0 ? n - 1 : b - 1;                     // the strip that consumes my records
    const int* prog_c = reinterpret_cast<const int*>(xbuf) + cb;
    uint2* pubV = sweep_rec(xbuf, n, Dp, 0, b, 0) + lane * N;
    uint2* pubB = sweep_rec(xbuf, n, Dp, j == 0 ? 1 : 2, b, 0) + lane * N;
    const int recGen = Dp / 2;
    // sources of the two hand-over states: + (parity of row r-1) * stride
    const uint16_t* rdV = (j == TW - 1 ? xring : Ld + size_t(j) * Dp) + lo;
    const int rdVs = j == TW - 1 ? 3 * Dp : parStride;
    const uint16_t* rdB = (consumer ? xring + (j == TW - 1 ? 2 : 1) * Dp : Ld + kindStride + size_t(j) * Dp) + lo;
    const int rdBs = consumer ? 3 * Dp : parStride;
    uint16_t* wrV = Ld + size_t(max(j - 1, 0)) * Dp + lo;                 // j == 0: never stored
    uint16_t* wrB = Ld + kindStride + size_t(max(j - 2, 0)) * Dp + lo;    // j <= 1: never stored
    uint16_t* ring = ringbase + size_t(j) * RING * Dp + lo;
    uint16_t* sring = sringbase + size_t(j) * RING * Dp + lo;
    const int Dh = Dp >> 1;
    const ptrdiff_t rowStep = (UP ? -1 : 1) * ptrdiff_t(W1) * Dp + Dp;   // next sweep row, next column
    const ptrdiff_t wrapBack = ptrdiff_t(W1) * Dp;
    const int ystart = UP ? H - 1 : 0;
    ptrdiff_t offPf = (ptrdiff_t(ystart) * W1 + c) * Dp + lo;   // next row to prefetch, its column
    int xPf = c;
    ptrdiff_t offCur = offPf;                                   // row being computed (first pass of MODE_HH stores there)
    int xcur = c;
    auto issue = [&](int row_) {
        if (row_ < H && active) {
            cp_async_lane<N>(ring + (row_ & (RING - 1)) * Dp, Cvol + offPf);
            cp_async_lane<N>(sring + (row_ & (RING - 1)) * Dp, Svol + offPf);
        }
        offPf += rowStep;
        if (++xPf == W1) { xPf = 0; offPf -= wrapBack; }
    };
#pragma unroll
    for (int i = 0; i < RING - 1; i++) { issue(i); cp_async_commit(); }

    uint32_t LtA[N];           // the down-right path: this chain's own state
#pragma unroll
    for (int q = 0; q < N; q++) LtA[q] = 0;
    int prog_known = 0, prog_next = 0;           // producers: record rows the left strip's agent has taken

    // one row; Q = r mod 4 (stage slot; parity = Q & 1); CONS = this warp takes states from the agent (it then runs its own
    // diagonal first and the agent-fed steps last, so that a late record costs as little as possible)
    auto row = [&](auto q_tag, auto cons_tag, int r) {
        constexpr int Q = decltype(q_tag)::value;
        constexpr bool CONS = decltype(cons_tag)::value;
        constexpr int PAR = Q & 1, PREV = PAR ^ 1;
        stamp(r, 0);
        issue(r + RING - 1);
        cp_async_commit();
        if (producer && (!(g.debug_flags & 16) || (r & 7) == 0)) {
            prog_known = max(prog_known, prog_next);
            prog_next = ld_relaxed_gpu(prog_c);            // inspected one row later: the round trip stays off this row's path
        }
        uint32_t LtV[N], LtB[N], Cc[N], S[N], Ln[N];
        cp_async_wait<RING - 1>();     // this thread's copies of row r have landed (each lane reads only its own bytes)
        if (active) {
            ld_regs<N>(ring + (r & (RING - 1)) * Dp, Cc); ld_regs<N>(sring + (r & (RING - 1)) * Dp, S);
        } else {
#pragma unroll
            for (int q = 0; q < N; q++) { Cc[q] = kMaxCostX2; S[q] = 0; }
        }
        auto add = [&](const uint32_t (&L)[N]) {
#pragma unroll
            for (int q = 0; q < N; q++) S[q] = CLAMP_EACH ? __vminu2(S[q] + L[q], kMaxCostX2) : S[q] + L[q];
        };
        auto step_a = [&] {            // down-right: registers only
            if (xcur == 0) {
#pragma unroll
                for (int q = 0; q < N; q++) LtA[q] = 0;
            }
            path_step<N>(Cc, LtA, Ln, lc);
            add(Ln);
        };
        if (CONS) {
            step_a();
            if (r > 0) named_bar_sync(BAR_XFULL + PREV, 96);      // the agent has staged the records of row r-1
        }
        if (active && r > 0) {
            ld_regs<N>(rdV + PREV * rdVs, LtV); ld_regs<N>(rdB + PREV * rdBs, LtB);
        } else {
#pragma unroll
            for (int q = 0; q < N; q++) { LtV[q] = 0; LtB[q] = 0; }
        }
        if (CONS && r > 0 && r - 1 + kXR < H - 1) named_bar_arrive(BAR_XEMPTY + PREV, 96);   // the agent may refill the slot
        stamp(r, 1);
        if (xcur == W1 - 1) {
#pragma unroll
            for (int q = 0; q < N; q++) LtB[q] = 0;
        }
        // ---- down and down-left: the states the previous warps (or the strip to the left) wait for
        path_step<N>(Cc, LtV, Ln, lc);
        add(Ln);
        path_step<N>(Cc, LtB, Ln, lc);
        add(Ln);
        if (active) {
            if (j >= 1) st_regs<N>(wrV + PAR * parStride, LtV);
            if (j >= 2) st_regs<N>(wrB + PAR * parStride, LtB);
        }
        if (producer) {
            const int need = r - kXK + 1;      // generation r % kXK held row r - kXK
            if (need > prog_known) sweep_wait([&] { prog_known = ld_relaxed_gpu(prog_c); return prog_known >= need; }, g.spin_limit, err, dead);
            if (active) {
                const int go = (r & (kXK - 1)) * recGen;
#pragma unroll
                for (int q = 0; q < N; q++) {
                    if (j == 0) st_volatile_v2(pubV + go + q, LtV[q], uint32_t(r + 1));
                    st_volatile_v2(pubB + go + q, LtB[q], uint32_t(r + 1));
                }
            }
        }
        stamp(r, 2);
        if (!CONS) step_a();
        if (!CLAMP_EACH) {
#pragma unroll
            for (int q = 0; q < N; q++) S[q] = __vminu2(S[q], kMaxCostX2);
        }
        if (DO_WTA) {
            if (!FULL) {
#pragma unroll
                for (int q = 0; q < N; q++) {   // cells beyond D never win and never veto
                    const int k = lane * N + q;
                    if (k >= g.w.D) S[q] = 0xFFFFFFFFu;
                    else if (k + Dh >= g.w.D) S[q] |= 0xFFFF0000u;
                }
            }
            if (r >= kStage) named_bar_sync(BAR_EMPTY + Q, nboth);    // the WTA warps have taken row r - kStage
            if (active) st_regs<N>(stage + Q * ringSlot + lo, S);
            named_bar_arrive(BAR_FULL + Q, nboth);
        } else {
            if (active) st_regs<N>(Svol + offCur, S);
        }
        offCur += rowStep;
        if (++xcur == W1) { xcur = 0; offCur -= wrapBack; }
        stamp(r, 3);
        named_bar_sync(BAR_ROW, nrow);      // every warp's states of row r are in Ld[PAR]; row r-1's have been read
    };

    auto sweep = [&](auto cons_tag) {
        int r = 0;
        for (; r + 3 < H; r += 4) {
            row(std::integral_constant<int, 0>{}, cons_tag, r);
            row(std::integral_constant<int, 1>{}, cons_tag, r + 1);
            row(std::integral_constant<int, 2>{}, cons_tag, r + 2);
            row(std::integral_constant<int, 3>{}, cons_tag, r + 3);
        }
        const int rem = H - r;
        if (rem > 0) row(std::integral_constant<int, 0>{}, cons_tag, r);
        if (rem > 1) row(std::integral_constant<int, 1>{}, cons_tag, r + 1);
        if (rem > 2) row(std::integral_constant<int, 2>{}, cons_tag, r + 2);
    };
    if (consumer) sweep(std::true_type{});
    else sweep(std::false_type{});
    cp_async_wait<0>();
    if (g.trace != nullptr && w == 0 && lane == 0) g.trace[4096 + 2 * b + 1] = clock64();
}

}  // namespace b200sgm
