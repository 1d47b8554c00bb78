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

// ---- TMA bulk copies (cp.async.bulk, completion counted in bytes on an mbarrier) ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, P1;\n"
                 "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Waits for the phase with the given parity to complete.  Bounded: a hand-over that never comes (a bug, or a sweep whose
// neighbour died) raises the sweep's error word instead of hanging the GPU; every other wait then gives up within 256 tries.
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, int* err)
{
    if (mbar_try(bar, parity)) return;
    int spins = 0;
    while (!mbar_try(bar, parity)) {
        if ((++spins & 255) == 0 && (spins > (1 << 22) || *reinterpret_cast<volatile int*>(err))) {
            atomicExch(err, 1);
            break;
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Horizontal pair.  Launch: one warp per row.
// ------------------------------------------------------------------------------------------------
constexpr int kHT = 8;        // tile width = checkpoint spacing of the <- chain
constexpr int kHRingB = 8;    // C columns in flight for the stream that reads DRAM (<-), prefetch distance 7
constexpr int kHRingF = 4;    // C columns in flight for the -> stream of phase B (L2 hits: <- read them a tile earlier)

inline size_t horiz_ckpt_elems(int W1, int H, int Dp) { return size_t(H) * ((W1 + kHT - 1) / kHT) * Dp; }
inline size_t horiz_smem_per_warp(int Dp) { return size_t(kHRingB + kHRingF + 2 * kHT) * Dp * sizeof(uint16_t); }

// dynamic smem per warp: ringB[kHRingB][Dp] | ringF[kHRingF][Dp] | tiles[2][kHT][Dp]  (uint16)
// FULL : Dp == 64*N (every lane active, all shared-memory offsets are immediates); padded counts run the same unrolled loops
//        with run-time offsets and their idle lanes predicated off
// CLAMP: saturate S_h (needed when 2*(Cmax+P2) could exceed 65535; padded cells are masked again before the WTA)
template <int N, bool FULL, bool CLAMP>
__global__ void __launch_bounds__(64) k_horiz(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Sh,
                                              uint16_t* __restrict__ ckpt, int W1, int H, int Dp_rt, int D, int P1, int P2,
                                              unsigned* __restrict__ maxc)
{
    extern __shared__ __align__(16) uint16_t smem_h[];
    const int Dp = FULL ? 64 * N : Dp_rt;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int y = blockIdx.x * (blockDim.x >> 5) + wib;
    if (y >= H) return;
    const LaneCtx lc = make_lane_ctx<N>(lane, Dp, P1, P2);
    const bool active = FULL || lc.active;
    const int lo = lane * 2 * N;
    const int nT = (W1 + kHT - 1) / kHT;
    const uint16_t* Crow = Cvol + size_t(y) * W1 * Dp + lo;
    uint16_t* Srow = Sh + size_t(y) * W1 * Dp + lo;
    uint16_t* ck = ckpt + size_t(y) * nT * Dp + lo;
    uint16_t* ringB = smem_h + size_t(wib) * (kHRingB + kHRingF + 2 * kHT) * Dp + lo;
    uint16_t* ringF = ringB + kHRingB * Dp;
    uint16_t* tiles = ringF + kHRingF * Dp;
    constexpr int PFB = kHRingB - 1, PFF = kHRingF - 1;

    auto ldC = [&](const uint16_t* p, uint32_t (&c)[N]) {
        if (active) ld_regs<N>(p, c);
        else {
#pragma unroll
            for (int j = 0; j < N; j++) c[j] = kMaxCostX2;
        }
    };

    uint32_t Lt[N], Ln[N], Cc[N];
    // Overflow guard: the largest cost-volume cell of the frame (phase B's -> stream sees every column once).  This
    // kernel is HBM bound, the extra packed max per register is free.  cv::StereoSGBM keeps C and L in int16: a frame
    // whose largest C + P2 exceeds 32767 is outside its defined behaviour and is reported by the host (b200sgm_wait).
    uint32_t cmx = 0;
    auto track = [&](const uint32_t (&c)[N]) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            if (FULL) cmx = __vmaxu2(cmx, c[j]);
            else {
                const int k = lane * N + j;
                cmx = __vmaxu2(cmx, c[j] & ((k < D ? 0xFFFFu : 0u) | (k + (Dp >> 1) < D ? 0xFFFF0000u : 0u)));
            }
        }
    };
    // ---------------- phase A: <- chain from x = W1-1 down to kHT, checkpoints only ----------------
    // ring slot of column x is x & 7; column x - 7 is requested while column x is processed
    if (W1 > kHT) {
        auto issueA = [&](int x) {
            if (x >= kHT && active) cp_async_lane<N>(ringB + (x & 7) * Dp, Crow + size_t(x) * Dp);
            cp_async_commit();
        };
        auto stepA = [&](int x) {
            issueA(x - PFB);
            cp_async_wait<PFB>();
            ldC(ringB + (x & 7) * Dp, Cc);
            path_step<N>(Cc, Lt, Ln, lc);
            if ((x & 7) == 0 && active) st_regs<N>(ck + size_t(x / kHT - 1) * Dp, Lt);   // entry state of tile x/kHT - 1
        };
#pragma unroll
        for (int j = 0; j < N; j++) Lt[j] = 0;
        int x = W1 - 1;
        for (int i = 0; i < PFB; i++) issueA(x - i);
        for (; (x & 7) != 7 && x >= kHT; x--) stepA(x);
        {
            // blocks of 8 columns x = xb+7 .. xb (ring slots are immediates; with FULL every shared-memory offset is); needs the
            // prefetched columns xb-7 .. xb-1 >= kHT.  Padded disparity counts run the same blocks with their idle lanes
            // predicated off (they carry kMaxCost through the shuffles).
            const uint16_t* gp = Crow + size_t(x - 7 - PFB) * Dp;      // column (xb - 7): first prefetch target of the block
            uint16_t* ckp = ck + size_t((x - 7) / kHT - 1) * Dp;
            for (; x - 7 - PFB >= kHT; x -= 8, gp -= 8 * Dp, ckp -= Dp) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    // processing column xb + 7 - i (slot 7 - i); request column xb - i (slot (8 - i) & 7)
                    if (active) cp_async_lane<N>(ringB + ((8 - i) & 7) * Dp, gp + (7 - i) * Dp);
                    cp_async_commit();
                    cp_async_wait<PFB>();
                    ldC(ringB + (7 - i) * Dp, Cc);
                    path_step<N>(Cc, Lt, Ln, lc);
                }
                if (active) st_regs<N>(ckp, Lt);
            }
        }
        for (; x >= kHT; x--) stepA(x);
        cp_async_wait<0>();
    }
    __syncwarp();
    // ---------------- phase B: -> chain; <- recomputed one tile ahead ----------------
    // Iteration `it` runs -> on column u = it (if 0 <= u < W1) and <- on its step v = it + 8 (if v < W1).  <- visits
    // the columns of tile v/8 from its right end: xb(v); the last tile may be partial.  Ring slots: <- step v uses
    // slot v & 7 of ringB (requested 7 iterations earlier), -> column u uses slot u & 3 of ringF (requested 3 earlier).
    // tiles[(c >> 3) & 1][c & 7] holds L<- of column c.  Every iteration commits exactly one cp.async group.
    auto xb_of = [&](int v) {
        const int base = v & ~7;
        return base + 8 <= W1 ? base + 7 - (v & 7) : W1 - 1 - (v - base);
    };
    auto issueB = [&](int v) { if (v < W1 && active) cp_async_lane<N>(ringB + (v & 7) * Dp, Crow + size_t(xb_of(v)) * Dp); };
    auto issueF = [&](int u) { if (u >= 0 && u < W1 && active) cp_async_lane<N>(ringF + (u & 3) * Dp, Crow + size_t(u) * Dp); };
    uint32_t Ltb[N], Lck[N];          // <- state; prefetched entry state of the next tile
#pragma unroll
    for (int j = 0; j < N; j++) { Lt[j] = 0; Ltb[j] = 0; Lck[j] = 0; }
    if (nT > 1 && active) ld_regs<N>(ck, Lck);       // entry state of tile 0 (plain loads: written by this lane in phase A)
    auto tile_entry = [&](int tv) {                 // called when <- enters tile tv: take its entry state, prefetch the next
#pragma unroll
        for (int j = 0; j < N; j++) Ltb[j] = (tv + 1 < nT) ? Lck[j] : 0u;
        if (tv + 2 < nT && active) ld_regs<N>(ck + size_t(tv + 1) * Dp, Lck);
    };
    auto iterB = [&](int it) {
        const int u = it, v = it + 8;
        issueB(v + PFB); issueF(u + PFF); cp_async_commit();
        cp_async_wait<PFF>();            // groups older than the newest 3 have landed: <- step v (7 back), -> column u (3 back)
        if (v < W1) {
            if ((v & 7) == 0) tile_entry(v >> 3);
            const int c = xb_of(v);
            uint32_t Cb[N], Lnb[N];
            ldC(ringB + (v & 7) * Dp, Cb);
            path_step<N>(Cb, Ltb, Lnb, lc);
            if (active) st_regs<N>(tiles + (((c >> 3) & 1) * kHT + (c & 7)) * Dp, Lnb);
        }
        if (u >= 0 && u < W1) {
            uint32_t Lb[N];
            ldC(ringF + (u & 3) * Dp, Cc);
            if (active) track(Cc);
            if (active) ld_regs<N>(tiles + (((u >> 3) & 1) * kHT + (u & 7)) * Dp, Lb);
            path_step<N>(Cc, Lt, Ln, lc);
            if (active) {
#pragma unroll
                for (int j = 0; j < N; j++) Ln[j] = CLAMP ? __vminu2(Ln[j] + Lb[j], kMaxCostX2) : Ln[j] + Lb[j];
                st_regs<N>(Srow + size_t(u) * Dp, Ln);
            }
        }
    };
    for (int v = 0; v < PFB; v++) { issueB(v); cp_async_commit(); }
    int it = -8;
    for (; it < 0; it++) iterB(it);
    {
        // fast blocks: -> on full tile t = it/8, <- on full tile t+1, prefetches reach into tile t+2 (must be full too)
        for (; it + 24 <= W1; it += 8) {
            const int t = it >> 3;
            tile_entry(t + 1);
            const uint16_t* gF = Crow + size_t(it) * Dp;            // column it
            uint16_t* gS = Srow + size_t(it) * Dp;
            const uint16_t* tcur = tiles + (t & 1) * kHT * Dp;
            uint16_t* tnxt = tiles + ((t + 1) & 1) * kHT * Dp;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                // <- step v = it+8+i on column it+15-i; its prefetch v+7: i == 0 -> column it+8 (same tile), else it+24-i
                if (active) {
                    cp_async_lane<N>(ringB + ((i + 7) & 7) * Dp, gF + (i == 0 ? 8 : 24 - i) * Dp);
                    cp_async_lane<N>(ringF + ((i + 3) & 3) * Dp, gF + (i + 3) * Dp);
                }
                cp_async_commit();
                cp_async_wait<PFF>();
                uint32_t Cb[N], Lnb[N], Lb[N];
                ldC(ringB + i * Dp, Cb);
                ldC(ringF + (i & 3) * Dp, Cc);
                if (active) { track(Cc); ld_regs<N>(tcur + i * Dp, Lb); }
                path_step<N>(Cb, Ltb, Lnb, lc);
                path_step<N>(Cc, Lt, Ln, lc);
                if (active) {
                    st_regs<N>(tnxt + (7 - i) * Dp, Lnb);
#pragma unroll
                    for (int j = 0; j < N; j++) Ln[j] = CLAMP ? __vminu2(Ln[j] + Lb[j], kMaxCostX2) : Ln[j] + Lb[j];
                    st_regs<N>(gS + i * Dp, Ln);
                }
            }
        }
    }
    for (; it < W1; it++) iterB(it);
    cp_async_wait<0>();
    cmx = __reduce_max_sync(kFullMask, max(cmx & 0xFFFFu, cmx >> 16));
    if (lane == 0 && cmx > 0) atomicMax(maxc, cmx);
}

// Per-register-count policies of the sweep, each picked by measurement on a B200 (2448x2048 unless noted; vertical stage, ms):
//   MBAR             : the stage-ring hand-over between a WTA warp and the path warps (and agent) of ITS columns goes through one
//                      mbarrier pair per WTA warp and slot instead of CTA-wide named barriers.  N = 8 (480 disparities): 5.27
//                      against 6.15; N = 4 (256): 1.65 against 1.53 -- the extra try_wait / elect / arrive instructions
//                      (+25 % on the path warps) cost more than the shorter barrier waits give back.
//   AGENT_PREFETCH   : the agent requests the neighbour's record of row r at the END of its row r (instead of polling at the top
//                      of row r + 1), so the L2 round trip overlaps the row barrier.  640x480x64: 0.237 against 0.261;
//                      1280x720x128 MODE_HH: 1.130 against 1.166; N = 4: 1.556 against 1.531 (not taken there).
//   WTA_INTERLEAVE   : both staged rows of a column go through each phase of the WTA together (N <= 4: 1.531 against 1.551);
//                      for N >= 8 the second row's registers spill (6.15 against 5.52), so rows go one after the other.
//   SPLIT_ROW        : split-phase row barrier of a WTA sweep (an mbarrier): a path warp ARRIVES as soon as its two diagonal
//                      states of the row are stored -- all its neighbours wait for -- then runs the down step, the sum and the
//                      stage hand-over, and only WAITS at the top of the next row.  Needs the non-blocking (MBAR) stage hand-over.
#ifndef B200SGM_VERT_SPLIT
#define B200SGM_VERT_SPLIT 0
#endif
//   FOLD_EMPTY       : (named-barrier hand-over) the WTA warps release a batch of two stage slots by ARRIVING at the row barrier
//                      of the row after next instead of at empty[q] barriers the path warps (and agents) sync on: one blocking
//                      CTA-wide barrier per row instead of two.  Odd rows use two alternating barrier ids for it.  The WTA
//                      warps then have two rows instead of almost three to resolve a batch: 640x480x64 0.241 -> 0.216,
//                      1280x1024x128 MODE_HH 1.153 -> 1.096, but N = 4 (busy WTA warps) 1.529 -> 1.606, also with both columns of
//                      a WTA warp resolved together (1.595): N <= 2 only.
#ifndef B200SGM_VERT_FOLD
#define B200SGM_VERT_FOLD 1
#endif
template <int N> struct VertPolicy {
    static constexpr bool SPLIT_ROW = B200SGM_VERT_SPLIT != 0;
    static constexpr bool FOLD_EMPTY = B200SGM_VERT_FOLD != 0 && N <= 2 && !SPLIT_ROW;
    static constexpr bool MBAR = N >= 8 || SPLIT_ROW, AGENT_PREFETCH = N != 4, WTA_INTERLEAVE = N <= 4;
};

// ------------------------------------------------------------------------------------------------
// Batched shared-memory WTA (A.6) used by the vertical sweep.  The column warp parks the summed cost of kWB
// consecutive rows in its private staging area and then resolves the kWB pixels TOGETHER: the kWB dependency
// chains (arg-min reduction -> neighbours -> uniqueness -> sub-pixel) are independent, so they overlap instead of
// each stalling the warp for its full latency.  Staged vectors must hold 0xFFFF in cells >= D.
// ------------------------------------------------------------------------------------------------
constexpr int kWB = 2;
// B200SGM_WTA_TOGETHER=1: the two columns of a WTA warp go through every phase of wta_vec together (four independent chains
// instead of two).  Measured at c3: 1.585 ms against 1.551 ms one column after the other -- the WTA warps are not what the
// path warps wait for (they wait for each other), and the extra live registers cost the N >= 8 instantiations spills.
#ifndef B200SGM_WTA_TOGETHER
#define B200SGM_WTA_TOGETHER 0
#endif
constexpr bool kWtaTogether = B200SGM_WTA_TOGETHER != 0;
constexpr int kWC = 2;        // columns per WTA warp of the vertical sweep (its path warps own one column each)
struct WtaCtx {
    uint32_t kk0;        // (lane*N) | (lane*N + Dh) << 16 : disparity indices of this lane's first word
    uint32_t umagic;     // floor(2^32 / f) + 1 with f = 100 - uniq > 0
    int f;
    int Dh;
};

// uint16 index of cell k in a paired vector, branch-free
__device__ __forceinline__ int cell_idx2(int k, int Dh, int Dp) { return 2 * k - (k >= Dh ? Dp - 1 : 0); }

// Per-lane accumulator of the deferred scalar part: lane l holds the reductions of row (base + l) of this column.
struct WtaAcc {
    uint32_t key;    // minS << 16 | best
    uint32_t mm;     // minimum cost outside best-1 .. best+1 (both halves)
    int sm, sp;      // S[best-1], S[best+1]
};

// Vector part for kWB staged rows of NC neighbouring columns (`qs` uint16 between rows, Dp between columns); row0 = index of
// staged row 0 in the sweep.  Only the warp-wide reductions happen here; what is left per pixel is a handful of scalars, parked
// in lane (row & 31) of `acc` and resolved for 32 rows at once by wta_flush32 (lane-parallel instead of warp-redundant).
// The NC * kWB pixels go through every phase TOGETHER: their dependency chains (arg-min reduction -> neighbour cells ->
// patched vector -> second reduction) are independent and each is latency bound, so the WTA warp -- the slowest stage of the
// sweep's path-warp / WTA-warp pipeline when the columns were resolved one after the other -- overlaps them.
// Columns >= ncol (the last warp of a strip with an odd column count) are computed on whatever the stage ring holds there and
// dropped: their stores and their accumulators are predicated off.
template <int N, int NC>
__device__ __forceinline__ void wta_vec(uint16_t* __restrict__ scratch, int qs, int row0, int ncol, const WtaGeom& g, const WtaCtx& w,
                                        int lane, bool active, WtaAcc (&acc)[NC])
{
    const int Dp = g.Dp;
    constexpr bool IL = VertPolicy<N>::WTA_INTERLEAVE;      // all reductions of a phase back to back (else one pixel at a time: fewer live registers)
    auto load = [&](int c, int q, uint32_t (&S)[N]) {
        if (active) ld_regs<N>(scratch + c * Dp + q * qs + lane * 2 * N, S);
        else {
#pragma unroll
            for (int j = 0; j < N; j++) S[j] = 0xFFFFFFFFu;
        }
    };
    __syncwarp();
    uint32_t key[NC][kWB];
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int q = 0; q < kWB; q++) {
            uint32_t S[N];
            load(c, q, S);
            // key = S << 16 | k: the warp minimum is the smallest cost and, among equals, the FIRST disparity
            uint32_t kq = 0xFFFFFFFFu;
#pragma unroll
            for (int j = 0; j < N; j++) {
                const uint32_t kk = w.kk0 + uint32_t(j) * 0x10001u;
                kq = __vimin3_u32(kq, S[j] * 0x10000u + (kk & 0xFFFFu), __byte_perm(kk, S[j], 0x7632));
            }
            key[c][q] = IL ? kq : __reduce_min_sync(kFullMask, kq);
        }
    }
    if constexpr (IL) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
#pragma unroll
            for (int q = 0; q < kWB; q++) key[c][q] = __reduce_min_sync(kFullMask, key[c][q]);
        }
    }
    // lanes 0,1,2 address cells best-1, best, best+1: lane 0 / 2 fetch the sub-pixel neighbours, then the three
    // lanes overwrite their cell with 0xFFFF (cells exempt from the uniqueness test)
    const int dl = min(lane, 2) - 1;
    int val[NC][kWB], cidx[NC][kWB];
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int q = 0; q < kWB; q++) {
            const int k = int(key[c][q] & 0xFFFFu) + dl;
            cidx[c][q] = (k >= 0 && k < Dp && c < ncol) ? c * Dp + q * qs + cell_idx2(k, w.Dh, Dp) : -1;
            val[c][q] = scratch[max(cidx[c][q], 0)];
        }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int q = 0; q < kWB; q++)
            if (lane < 3 && cidx[c][q] >= 0) scratch[cidx[c][q]] = 0xFFFFu;
    }
    __syncwarp();
    uint32_t mm[NC][kWB];
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int q = 0; q < kWB; q++) {
            uint32_t T[N];
            load(c, q, T);
            uint32_t m = T[0];
#pragma unroll
            for (int j = 1; j < N; j++) m = __vminu2(m, T[j]);
            m = __vminu2(m, __byte_perm(m, 0, 0x1032));
            mm[c][q] = IL ? m : __reduce_min_sync(kFullMask, m);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int q = 0; q < kWB; q++) {
            const uint32_t m = IL ? __reduce_min_sync(kFullMask, mm[c][q]) : mm[c][q];
            const int smv = __shfl_sync(kFullMask, val[c][q], 0), spv = __shfl_sync(kFullMask, val[c][q], 2);
            if (lane == ((row0 + q) & 31) && c < ncol) { acc[c].key = key[c][q]; acc[c].mm = m; acc[c].sm = smv; acc[c].sp = spv; }
        }
    }
}

// Scalar part for up to 32 rows: lane l resolves row l of the block (cnt = valid rows).  dptr/kptr = output pointers
// of the block's first row, advancing by dStride per row.
__device__ __forceinline__ void wta_flush32(const WtaAcc& acc, int cnt, const WtaGeom& g, const WtaCtx& w, int x1, int lane,
                                            int16_t* __restrict__ dptr, uint32_t* __restrict__ kptr, ptrdiff_t dStride)
{
    if (lane >= cnt) return;
    const int minS = int(acc.key >> 16), best = int(acc.key & 0xFFFFu);
    // uniqueness: S[k] * f < minS * 100  <=>  S[k] < ceil(minS*100 / f)
    // magic division by f = 100 - uniq (2 <= f <= 100); for f == 1 the magic constant does not fit 32 bits and the quotient is n itself
    const uint32_t n100 = uint32_t(minS * 100 + w.f - 1);
    const uint32_t thr = min(w.f == 1 ? n100 : __umulhi(n100, w.umagic), 0xFFFFu);
    const bool reject = (acc.mm & 0xFFFFu) < thr || minS >= kMaxCost;
    int dfix = best * 16;
    if (best > 0 && best < g.D - 1) {
        const int den = max(acc.sm + acc.sp - 2 * minS, 1);
        // |quotient| <= 8.5 and numerator, denominator < 2^24: IEEE float division then truncation is exact
        dfix += __float2int_rz(__fdiv_rn(float((acc.sm - acc.sp) * 16 + den), float(den * 2)));
    }
    if (!reject) {
        const int x = x1 + g.minX1;
        const int x2 = x - best - g.minD;
        if (x2 >= 0 && x2 < g.W) atomicMin(kptr + lane * dStride + x2, (uint32_t(minS) << 16) | uint32_t(0xFFFF - x));
    }
    dptr[lane * dStride] = int16_t(reject ? g.INVALID : dfix + g.minD * 16);
}

// Exact fallback for uniquenessRatio >= 100 (f <= 0: the comparison cannot be turned into a threshold): one row.
template <int N>
__device__ __forceinline__ int wta_slow(const uint16_t* __restrict__ scratch, const WtaGeom& g, const WtaCtx& w, int x1,
                                        int lane, bool active, uint32_t* __restrict__ disp2key_row)
{
    uint32_t S[N];
    if (active) ld_regs<N>(scratch + lane * 2 * N, S);
    else {
#pragma unroll
        for (int j = 0; j < N; j++) S[j] = 0xFFFFFFFFu;
    }
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const uint32_t kk = w.kk0 + uint32_t(j) * 0x10001u;
        key = __vimin3_u32(key, __byte_perm(kk, S[j], 0x5410), __byte_perm(kk, S[j], 0x7632));
    }
    key = __reduce_min_sync(kFullMask, key);
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
                const int sm = scratch[cell_idx2(best - 1, w.Dh, g.Dp)], sp = scratch[cell_idx2(best + 1, w.Dh, g.Dp)];
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
    int debug_flags;      // timing experiments only (results are wrong): 1 = no inter-strip exchange, 2 = no WTA
    int agents;           // 1: two extra warps per CTA poll the neighbours' records on behalf of the edge warps
};

// Inter-strip exchange, "low latency" protocol: a record is Dp/2 entries of {two packed costs, tag}; each
// entry is one 8-byte volatile store, so data and flag arrive together and no fence or separate flag is
// needed.  tag = row + 1 (the buffer is zeroed before the launch).  [side][strip][row & 3][Dp/2] uint2.
// Four generations are live at once: a strip publishes row R early in its row R, having only waited (late in row
// R-1) for the neighbour's record R-2 -- and that neighbour still reads our record R-3 late in its row R-2.
// The consumer issues its record load at the TOP of the row and only inspects the tags right before it needs
// the data, so in the common case (the neighbour is not late) the L2 round trip is off the critical path.
constexpr int kXbufGen = 4;
constexpr int kHaloGen = 16;      // record generations of a HALO > 1 sweep: a strip may run HALO rows ahead of a neighbour whose agent still
                                  // needs a record HALO rows behind its own row (the records come from inner warps, which publish before
                                  // their strip's edge warp has met its agent), so 2 * HALO + 1 generations are live at worst
constexpr int kHaloRows = 8;      // rows of the agents' halo C rings
constexpr int kHaloPF = 4;        // ... requested this many rows ahead
constexpr int kHaloMax = 4;
// RING (template parameter of k_vert) = rows of C and of S_h in flight per column (cp.async rings in shared memory): DRAM latency x
// row rate.  The two rings have the same depth -- cp.async groups retire in order, so a shallower S_h ring would make its wait
// drain the younger C requests as well.  8 up to 256 disparities, 4 beyond (a row is then at least twice the bytes).
// 64 registers per thread up to 256 disparities, 72 for the N >= 8 instantiations.  (A sweep never has more than 16 + 8 + 2
// warps; bounding the N <= 4 kernels by those 832 threads lets ptxas use 72 registers and removes the last spills, but the
// row loop gets slower: 1.62 instead of 1.55 ms at c3 -- measured, kept at 64.)
#ifndef B200SGM_VERT_MAXT
#define B200SGM_VERT_MAXT 1024
#endif
// B200SGM_VERT_CPS = strips (CTAs) per SM: 2 halves the strip width, and the two lock-step groups of an SM fill each other's
// barrier and latency gaps
#ifndef B200SGM_VERT_CPS
#define B200SGM_VERT_CPS 1
#endif
constexpr int kVertCps = B200SGM_VERT_CPS;
// threads per CTA bound the registers per thread: beyond 256 disparities shared memory limits a strip to 13 / 6 / 3 columns, so
// the bound follows: 640 threads = 96 registers for N = 8 (the row loop spills 238 bytes at 80, 24 at 96).  A 13-column strip
// (what 480 disparities leave of the shared memory) then has no room for the two agent warps and its edge warps poll the
// records themselves: 4.97 -> 4.70 ms at 2448x2048x480 all the same; narrower strips keep their agents
constexpr int vert_max_threads(int n) { return kVertCps == 2 ? 448 : (n <= 4 ? B200SGM_VERT_MAXT : (n == 8 ? 640 : (n == 16 ? 384 : 256))); }
constexpr int kRowUnroll = 4; // the row loop is unrolled by this: record generation, stage slot, parity are immediates
__device__ __forceinline__ uint2* xrec(uint2* xbuf, int nstrips, int Dp, int side, int strip, int row, int gens = kXbufGen)
{
    return xbuf + (size_t((side * nstrips + strip) * gens + (row & (gens - 1)))) * (Dp / 2);
}
__device__ __forceinline__ void st_volatile_v2(uint2* p, uint32_t a, uint32_t b)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p)
{
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

// B200SGM_VERT_TMA: a row of a strip is one contiguous run of TW * Dp costs in C and in S_h, so one elected thread fetches it
// with two bulk copies per row (rings [slot][column][Dp], one mbarrier per slot) instead of two cp.async per column warp.
#ifndef B200SGM_VERT_TMA
#define B200SGM_VERT_TMA 0
#endif
constexpr bool kVertTma = B200SGM_VERT_TMA != 0;
constexpr int kStage = 4;     // rows of summed cost S parked for the WTA warps (ring, handed over with named barriers)
static_assert(kStage == kRowUnroll && kXbufGen == kRowUnroll && kStage % kWB == 0, "stage slot = record generation = row & 3");
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// dynamic smem: Ld[2 parity][2 dir][twmax+2][Dp] | Cring[twmax][RING][Dp] | Sring[twmax][RING][Dp] | stage[kStage][twmax][Dp]
//               | xring[2 sides][kXbufGen][Dp]   (uint16)
inline size_t vert_smem_bytes(int twmax, int Dp, int ring, int halo = 1)
{
    return (size_t(4) * (twmax + 2) + size_t(2 * ring + kStage) * twmax + 2 * kXbufGen) * Dp * sizeof(uint16_t)   // 190 KB at c3
           + (halo > 1 ? size_t(4) * kHaloRows * (halo - 1) * Dp * sizeof(uint16_t) : 0)                             // + the halo producers' C rings [side][producer][row][column]
           + (halo > 1 ? size_t(2) * kXbufGen * Dp * sizeof(uint16_t) : 0)                                          // + their hand-over ring [side][row & 3]
           + (kVertTma ? size_t(ring) * sizeof(uint64_t) : 0)                                                        // + the ring's mbarriers
           + size_t(2) * kStage * ((twmax + kWC - 1) / kWC) * sizeof(uint64_t)                                       // + full/empty per WTA warp and slot
           + 2 * sizeof(uint64_t);                                                                                   // + the split-phase row barrier (padded to 16 bytes)
}

// Warp-specialised vertical sweep.  Launch: 32 * twmax threads when !DO_WTA, else 64 * twmax:
//   path warps (0 .. twmax-1)        : one per column; the three path updates of a row -- the only work on the
//                                      row-to-row dependency chain -- and the sum S, parked in the stage ring; they run
//                                      in lock step with one 32*TW-thread barrier per row.
//   WTA warps  (twmax .. 2*twmax-1)  : one per column; resolve kWB parked rows at a time (winner-take-all, uniqueness,
//                                      sub-pixel, disp2 scatter).  They trail the path warps by up to kStage rows and fill
//                                      the issue slots the path warps leave idle while they wait on each other.
//   agent warps (last two, optional) : one per neighbouring strip; poll the neighbour's exchange record of each row
//                                      with back-to-back loads and drop it into shared memory, so the L2 round trips
//                                      of the inter-strip exchange never sit on an edge path warp's critical path.
// Ring hand-over: named barriers full[q] (path warps arrive, WTA warps sync) and empty[q] (the reverse); halo[side]
// pairs an agent with its edge warp once per row.
// FULL      : Dp == D == 64*N (no padded cells, every lane active)
// CLAMP_EACH: saturate after every addition of the sum (needed when the 16-bit sum of the terms could wrap)
// HALO      : 1 = a strip's edge column takes its incoming diagonal state from the neighbour's record of the previous row (one L2
//             hand-over per row: the floor of the row time of narrow strips).  HALO = h > 1: the neighbour publishes the state
//             of the column h columns inside its edge, and the agent advances that diagonal chain through the h - 1 columns in
//             between itself (their C read a second time, by the agent): the record is then needed h - 1 rows after it was
//             written, and neighbouring strips may drift h - 1 rows apart.
template <int N, int RING, bool UP, bool DO_WTA, bool FULL, bool CLAMP_EACH, int HALO = 1>
__global__ void __launch_bounds__(vert_max_threads(N), kVertCps) k_vert(const uint16_t* __restrict__ Cvol, uint16_t* __restrict__ Svol, VertGeom g,
                                                  int16_t* __restrict__ disp, uint32_t* __restrict__ disp2key,
                                                  uint2* __restrict__ xbuf, int* __restrict__ err)
{
    extern __shared__ __align__(16) uint16_t smem_v[];
    constexpr bool kVertMbar = VertPolicy<N>::MBAR, kAgentPrefetch = VertPolicy<N>::AGENT_PREFETCH;
    constexpr bool SPLIT = DO_WTA && VertPolicy<N>::SPLIT_ROW;
    constexpr bool FOLD = DO_WTA && VertPolicy<N>::FOLD_EMPTY;      // (implies !kVertMbar)
    const int W1 = g.w.W1, H = g.w.H, Dp = FULL ? 64 * N : g.w.Dp;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x, n = g.nstrips;
    const int x0 = int((long long)b * W1 / n), x1e = int((long long)(b + 1) * W1 / n);
    const int TW = x1e - x0;
    const int slots = g.twmax + 2;
    // Ld holds the NORMALISED diagonal states; zero = "predecessor outside the image" (first row, border columns)
    uint16_t* Ld = smem_v;
    uint16_t* ringbase = Ld + size_t(4) * slots * Dp;
    uint16_t* sringbase = ringbase + size_t(RING) * g.twmax * Dp;
    uint16_t* stagebase = sringbase + size_t(RING) * g.twmax * Dp;
    uint16_t* xringbase = stagebase + size_t(kStage) * g.twmax * Dp;
    uint64_t* ringbar = reinterpret_cast<uint64_t*>(xringbase + size_t(2) * kXbufGen * Dp);     // [RING], kVertTma only
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(smem_v);
        const int nz = 2 * slots * Dp;
        for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0;
    }
    // hand-over barriers of the stage ring: fullbar[wta warp][slot] (arrivals: the path warps of its columns, plus the agent
    // of an edge column), emptybar[wta warp][slot] (one arrival: the WTA warp)
    const int nwta = (g.twmax + kWC - 1) / kWC;
    uint64_t* fullbar = ringbar + (kVertTma ? RING : 0);
    uint64_t* emptybar = fullbar + nwta * kStage;
    uint64_t* rowbar = emptybar + nwta * kStage;          // SPLIT: one arrival per path warp and agent of the strip and row
    uint16_t* haloring = reinterpret_cast<uint16_t*>(rowbar + 2);       // HALO > 1: [side][kHaloRows][HALO - 1][Dp] (16-byte aligned)
    constexpr int XG = HALO > 1 ? kHaloGen : kXbufGen;                  // record generations
    static_assert(HALO >= 1 && HALO <= kHaloMax && kHaloRows >= kHaloPF + HALO, "halo ring depth");
    if (kVertTma && threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < RING; i++) mbar_init(ringbar + i, 1);
    }
    if (kVertMbar && DO_WTA && threadIdx.x < nwta) {
        const int i = threadIdx.x;
        const bool ag = g.agents && !(g.debug_flags & 1);
        const int cols = min(kWC, TW - kWC * i);
        const int prod = cols + ((ag && b > 0 && i == 0) ? 1 : 0) + ((ag && b < n - 1 && TW > 1 && (TW - 1) / kWC == i) ? 1 : 0);
        for (int q = 0; q < kStage; q++) {
            mbar_init(fullbar + i * kStage + q, max(prod, 1));
            mbar_init(emptybar + i * kStage + q, 1);
        }
    }
    if (SPLIT && threadIdx.x == 0) {
        const bool ag = g.agents && !(g.debug_flags & 1);
        mbar_init(rowbar, TW + (ag ? (b > 0 ? 1 : 0) + (b < n - 1 ? 1 : 0) : 0));
    }
    if ((kVertTma || kVertMbar) && threadIdx.x < 32) mbar_init_fence();
    __syncthreads();
    // ring geometry: cp.async [column][slot][Dp], bulk copies [slot][column][Dp]
    const int ringColStride = kVertTma ? Dp : RING * Dp;
    const int ringSlotStride = kVertTma ? g.twmax * Dp : Dp;
    constexpr int kRingShift = RING == 8 ? 3 : 2;
    static_assert(RING == 8 || RING == 4, "ring depth");
    const int wta_warps = DO_WTA ? (g.twmax + kWC - 1) / kWC : 0;      // launched; (TW + kWC - 1) / kWC of them have columns
    const int agent_base = g.twmax + wta_warps;
    const int TWW = DO_WTA ? (TW + kWC - 1) / kWC : 0;
    const bool exchange_on = !(g.debug_flags & 1);
    const LaneCtx lc = make_lane_ctx<N>(lane, Dp, g.P1, g.P2);
    const bool active = FULL || lc.active;
    constexpr int BAR_ROW = 1, BAR_FULL = 2, BAR_EMPTY = 2 + kStage, BAR_HALO = 2 + 2 * kStage;
    // FOLD: row barrier of odd rows = BAR_ROWX + ((r >> 1) & 1) (the ids of empty[0], empty[1]); from row 3 on it also counts the
    // WTA warps, which arrive there once they have resolved rows r - 3 and r - 2 (whose slots are rewritten in rows r + 1, r + 2)
    constexpr int BAR_ROWX = BAR_EMPTY;
    if (w >= agent_base) {
        // ================================ agent warps ================================
        // agent warps 0, 1: one per side (0: left neighbour, feeds warp 0; 1: right neighbour, feeds warp TW-1).  HALO > 1 launches
        // four more, the halo PRODUCERS (two per side, alternating rows): they advance the incoming diagonal chains through the
        // halo on their own clock and hand the result over through a 4-row ring and a 64-thread named barrier, so that the
        // chain's serial latency is off the strip's row time; agents 0, 1 keep the edge column's own step (WTA sweeps).
        const int aw = w - agent_base;
        const bool producer_role = HALO > 1 && aw >= 2;
        const int side = producer_role ? (aw - 2) >> 1 : aw;
        const int prod = producer_role ? (aw - 2) & 1 : 0;
        if (!g.agents || !exchange_on || (side == 0 ? b == 0 : b == n - 1)) return;
        if (HALO > 1 && !DO_WTA && !producer_role) return;      // sweeps without WTA: the edge path warp consumes the producers' ring itself
        constexpr int BAR_HB = 12;                               // + 2 * side + producer: ids 12 .. 15, free in every sweep kind
        uint16_t* vring = (DO_WTA ? haloring + size_t(4) * kHaloRows * (HALO - 1) * Dp + size_t(side) * kXbufGen * Dp
                                  : xringbase + size_t(side) * kXbufGen * Dp) + lane * 2 * N;
        const uint2* rec0 = xrec(xbuf, n, Dp, side, side == 0 ? b - 1 : b + 1, 0, XG) + lane * N;
        uint16_t* dst = xringbase + size_t(side) * kXbufGen * Dp + lane * 2 * N;
        const int gen_stride = Dp / 2;
        bool dead = false;
        // ---- HALO > 1: the diagonal chain that enters the strip at row i + 1 is seeded with the neighbour's record of row
        // i - HALO + 1 (state of the column HALO columns outside the edge column) and advanced here through the HALO - 1 columns
        // in between; halo_V(i) = its state one column outside the edge column at row i = what the HALO = 1 record of row i holds.
        const int xe = side == 0 ? x0 : x1e - 1;                               // edge column (of the volume)
        const int hdir = side == 0 ? 1 : -1;                                   // the chain moves this way, one column per row
        uint16_t* hring = haloring + size_t(2 * side + prod) * kHaloRows * (HALO - 1) * Dp + lane * 2 * N;
        const uint16_t* gHalo = Cvol + (size_t(UP ? H - 1 : 0) * W1 + (xe - hdir * (HALO - 1))) * Dp + lane * 2 * N;   // halo column k = 1, next row to request
        const ptrdiff_t hRowStride = (UP ? -1 : 1) * ptrdiff_t(W1) * Dp;
        auto halo_issue = [&](int row_) {
            if (HALO > 1) {
                if (row_ < H && active) {
#pragma unroll
                    for (int k = 1; k < HALO; k++)
                        cp_async_lane<N>(hring + ((row_ & (kHaloRows - 1)) * (HALO - 1) + (k - 1)) * Dp, gHalo + ptrdiff_t(hdir * (k - 1)) * Dp);
                }
                gHalo += hRowStride;
                cp_async_commit();
            }
        };
        // the seed of halo_V(i + 1) is requested at the end of halo_V(i): it was published two rows ago, its L2 round trip then
        // overlaps the agent's other work instead of heading the next chain
        uint2 hpre[N];
#pragma unroll
        for (int q = 0; q < N; q++) hpre[q] = make_uint2(0u, 0u);
        auto halo_prefetch = [&](int i) {       // seed record of halo_V(i)
            const int rho0 = i - (HALO - 1);
            if (HALO > 1 && rho0 >= 0 && i + 1 < H && active) {
                const uint2* rec = rec0 + (rho0 & (XG - 1)) * gen_stride;
#pragma unroll
                for (int q = 0; q < N; q++) hpre[q] = ld_volatile_v2(rec + q);
            }
        };
        auto halo_V = [&](int i, int inext, uint32_t (&St)[N]) {
#pragma unroll
            for (int q = 0; q < N; q++) St[q] = 0;
            const int rho0 = i - (HALO - 1);
            int k0 = 1;
            if (rho0 >= 0) {
                if (active && !dead) {
                    bool ok = true;
#pragma unroll
                    for (int q = 0; q < N; q++) { St[q] = hpre[q].x; ok = ok && hpre[q].y == uint32_t(rho0 + 1); }
                    if (!ok) {
                        const uint2* rec = rec0 + (rho0 & (XG - 1)) * gen_stride;
                        const long long t0 = clock64();
                        int spins = 0;
                        while (true) {
                            ok = true;
#pragma unroll
                            for (int q = 0; q < N; q++) {
                                uint2 v = ld_volatile_v2(rec + q);
                                St[q] = v.x;
                                ok = ok && v.y == uint32_t(rho0 + 1);
                            }
                            if (ok) break;
                            if ((++spins & 255) == 0 && (clock64() - t0 > g.spin_limit || *reinterpret_cast<volatile int*>(err))) {
                                atomicExch(err, 1);
                                dead = true;
                                break;
                            }
                        }
                    }
                }
                dead = __any_sync(kFullMask, dead);
            } else {
                k0 = -rho0;                   // the chain starts at the first row of the sweep, from "outside the image"
            }
            halo_prefetch(inext);
            cp_async_wait<kHaloPF>();         // the halo rows up to row i have landed (each lane reads only its own bytes)
#pragma unroll
            for (int k = 1; k < HALO; k++) {
                if (k >= k0) {
                    uint32_t Ch[N], Lh[N];
                    if (active) ld_regs<N>(hring + (((rho0 + k) & (kHaloRows - 1)) * (HALO - 1) + (k - 1)) * Dp, Ch);
                    else {
#pragma unroll
                        for (int q = 0; q < N; q++) Ch[q] = kMaxCostX2;
                    }
                    path_step<N>(Ch, St, Lh, lc);
                }
            }
        };
        if (producer_role) {
            int next = 0;
            for (int i = prod; i + 1 < H; i += 2) {              // state for the strip's row i + 1
                while (next <= i + kHaloPF) halo_issue(next++);  // kHaloPF groups newer than halo row i in flight
                uint32_t St[N];
                halo_V(i, i + 2, St);
                if (active) st_regs<N>(vring + (i & (kXbufGen - 1)) * Dp, St);
                named_bar_sync(BAR_HB + 2 * side + prod, 64);    // meets the consumer at the top of row i + 1
            }
            cp_async_wait<0>();
            return;
        }
        if (DO_WTA) {
            // The agent OWNS the edge column's incoming-diagonal path (the one step of the row that depends on the
            // neighbour strip): record of row r-1 -> path update with C(r) from the column's cp.async ring -> new state for
            // the inner neighbour column (Ld) and L for the sum (stageB, added by the column's WTA warp).  The edge path
            // warp is left with two independent steps, so the exchange latency is off every column warp's critical path.
            const int nag = (b > 0 ? 1 : 0) + (b < n - 1 ? 1 : 0);
            const int nrow_a = 32 * (TW + nag), nfe_a = 32 * (TW + TWW + nag);
            const int je = side == 0 ? 0 : TW - 1;
            uint16_t* wr = Ld + side * (slots * Dp) + (je + 1) * Dp + lane * 2 * N;          // + parity * 2*slots*Dp
            const uint16_t* cring = ringbase + size_t(je) * ringColStride + lane * 2 * N;
            uint16_t* sB = dst;                                                              // stageB[side][kStage][Dp] aliases the xring
            if (!kVertTma) asm volatile("bar.sync 1, %0;" ::"r"(nrow_a) : "memory");           // C(0) is visible
            if constexpr (HALO > 1) {
                static_assert(!VertPolicy<N>::SPLIT_ROW && !kVertTma, "halo agents: named row barrier, cp.async rings");
                uint32_t Vp[N];                     // state one column outside the edge column at row r - 1 (zero: outside the image)
#pragma unroll
                for (int q = 0; q < N; q++) Vp[q] = 0;
                uint64_t* fbh = fullbar + (je / kWC) * kStage;
                uint64_t* ebh = emptybar + (je / kWC) * kStage;
                for (int r = 0; r < H; r++) {
                    if (r > 0) {
                        named_bar_sync(BAR_HB + 2 * side + ((r - 1) & 1), 64);        // the producer of row r - 1 has delivered
                        if (active) ld_regs<N>(vring + ((r - 1) & (kXbufGen - 1)) * Dp, Vp);
                    }
                    uint32_t Cc[N], Ln[N];
                    if (active) ld_regs<N>(cring + (r & (RING - 1)) * ringSlotStride, Cc);
                    else {
#pragma unroll
                        for (int q = 0; q < N; q++) Cc[q] = kMaxCostX2;
                    }
                    path_step<N>(Cc, Vp, Ln, lc);                       // the edge column's incoming-diagonal step of row r
                    if (active) st_regs<N>(wr + (r & 1) * (2 * slots * Dp), Vp);
                    const int q4 = r & (kStage - 1);
                    if (kVertMbar) {
                        if (r >= kStage) mbar_wait_b(ebh + q4, ((r >> 2) + 1) & 1, err);
                        if (active) st_regs<N>(sB + q4 * Dp, Ln);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(fbh + q4);
                    } else {
                        if (!FOLD && r >= kStage) named_bar_sync(BAR_EMPTY + q4, nfe_a);
                        if (active) st_regs<N>(sB + q4 * Dp, Ln);
                        named_bar_arrive(BAR_FULL + q4, nfe_a);
                    }
                    if (FOLD && (r & 1)) named_bar_sync(BAR_ROWX + ((r >> 1) & 1), r >= 3 ? nfe_a : nrow_a);
                    else asm volatile("bar.sync 1, %0;" ::"r"(nrow_a) : "memory");             // BAR_ROW of row r
                }
                cp_async_wait<0>();
                return;
            }
            // the record of row r - 1 is requested at the END of row r - 1 (the neighbour published it early in its own row
            // r - 1), so that its L2 round trip overlaps the row barrier instead of heading this warp's row
            uint2 pre[N];
#pragma unroll
            for (int q = 0; q < N; q++) pre[q] = make_uint2(0u, 0u);
            uint64_t* fb = fullbar + (je / kWC) * kStage;
            uint64_t* eb = emptybar + (je / kWC) * kStage;
            for (int r = 0; r < H; r++) {
                uint32_t Cc[N], Lt[N], Ln[N];
#pragma unroll
                for (int q = 0; q < N; q++) Lt[q] = 0;
                if (SPLIT && r > 0) mbar_wait_b(rowbar, (r - 1) & 1, err);      // row r - 1 of the strip is complete
                if (r > 0 && active && !dead) {
                    bool ok = true;
#pragma unroll
                    for (int q = 0; q < N; q++) { Lt[q] = pre[q].x; ok = ok && pre[q].y == uint32_t(r); }
                    if (!ok && !(g.debug_flags & 8)) {            // 8: timing experiment, take whatever is there
                        const uint2* rec = rec0 + ((r - 1) & (kXbufGen - 1)) * gen_stride;
                        const long long t0 = clock64();
                        int spins = 0;
                        while (true) {
                            ok = true;
#pragma unroll
                            for (int q = 0; q < N; q++) {
                                uint2 v = ld_volatile_v2(rec + q);
                                Lt[q] = v.x;
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
                }
                dead = __any_sync(kFullMask, dead);
                if (kVertTma) mbar_wait(ringbar + (r & (RING - 1)), (r >> kRingShift) & 1);
                if (active) ld_regs<N>(cring + (r & (RING - 1)) * ringSlotStride, Cc);
                else {
#pragma unroll
                    for (int q = 0; q < N; q++) Cc[q] = kMaxCostX2;
                }
                path_step<N>(Cc, Lt, Ln, lc);
                if (active) st_regs<N>(wr + (r & 1) * (2 * slots * Dp), Lt);
                if (SPLIT) {               // the state is all the strip's row waits for from this warp
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rowbar);
                }
                const int q4 = r & (kStage - 1);
                if (kVertMbar) {
                    if (r >= kStage) mbar_wait_b(eb + q4, ((r >> 2) + 1) & 1, err);
                    if (active) st_regs<N>(sB + q4 * Dp, Ln);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(fb + q4);
                } else {
                    if (!FOLD && r >= kStage) named_bar_sync(BAR_EMPTY + q4, nfe_a);
                    if (active) st_regs<N>(sB + q4 * Dp, Ln);
                    named_bar_arrive(BAR_FULL + q4, nfe_a);
                }
                if (kAgentPrefetch && r + 1 < H && active) {
                    const uint2* rec = rec0 + (r & (kXbufGen - 1)) * gen_stride;
#pragma unroll
                    for (int q = 0; q < N; q++) pre[q] = ld_volatile_v2(rec + q);
                }
                if (FOLD && (r & 1)) named_bar_sync(BAR_ROWX + ((r >> 1) & 1), r >= 3 ? nfe_a : nrow_a);
                else if (!SPLIT) asm volatile("bar.sync 1, %0;" ::"r"(nrow_a) : "memory");     // BAR_ROW of row r
            }
            return;
        }
        for (int i = 0; i + 1 < H; i++) {                    // record of row i: consumed by the edge warp in its row i + 1
            const uint2* rec = rec0 + (i & (kXbufGen - 1)) * gen_stride;
            uint32_t d[N];
#pragma unroll
            for (int q = 0; q < N; q++) d[q] = 0;
            if (active && !dead) {
                const long long t0 = clock64();
                int spins = 0;
                while (true) {
                    bool ok = true;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        uint2 v = ld_volatile_v2(rec + q);
                        d[q] = v.x;
                        ok = ok && v.y == uint32_t(i + 1);
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
            if (active) st_regs<N>(dst + (i & (kXbufGen - 1)) * Dp, d);
            named_bar_sync(BAR_HALO + side, 64);
        }
        return;
    }
    const bool wta_role = DO_WTA && w >= g.twmax;
    const int j = wta_role ? kWC * (w - g.twmax) : w;      // (first) column of this warp
    if (j >= TW) return;                 // the barriers below only count the warps that own columns
    // agents of a WTA sweep take part in the row barrier and in the stage-ring hand-over
    const int nag3 = (DO_WTA && g.agents && exchange_on) ? (b > 0 ? 1 : 0) + (b < n - 1 ? 1 : 0) : 0;
    const int nrow = 32 * (TW + nag3), nboth = 32 * (TW + TWW + nag3);
    const int x = x0 + j;
    const int lo = lane * 2 * N;
    const int ringSlot = g.twmax * Dp;                 // slot stride of the stage ring [slot][warp][Dp]
    const ptrdiff_t rowStride = (UP ? -1 : 1) * ptrdiff_t(W1) * Dp;
    const int ystart = UP ? H - 1 : 0;
    uint16_t* stage = stagebase + size_t(j) * Dp;      // + slot * ringSlot
    const bool wta_on = !(g.debug_flags & 2);

    if (wta_role) {
        // ================================ WTA warps ================================
        int16_t* dptr = disp + size_t(ystart) * g.w.W + x + g.w.minX1;
        uint32_t* kptr = disp2key + size_t(ystart) * g.w.W;
        const ptrdiff_t dStride = (UP ? -1 : 1) * ptrdiff_t(g.w.W);
        WtaCtx wc;
        wc.Dh = Dp >> 1;
        wc.kk0 = uint32_t(lane * N) | (uint32_t(lane * N + wc.Dh) << 16);
        wc.f = 100 - g.w.uniq;
        wc.umagic = wc.f > 0 ? uint32_t((1ull << 32) / uint32_t(wc.f)) + 1u : 0u;
        // this warp resolves columns j .. j + ncol - 1 (ncol <= kWC) of every parked row
        const int ncol = min(kWC, TW - j);
        WtaAcc acc[kWC];
        int edge_side[kWC];
#pragma unroll
        for (int c = 0; c < kWC; c++) {
            acc[c] = WtaAcc{0xFFFFFFFFu, 0u, 0, 0};
            const int jc = j + c;
            edge_side[c] = nag3 == 0 ? -1 : ((jc == 0 && b > 0) ? 0 : ((jc == TW - 1 && jc != 0 && b < n - 1) ? 1 : -1));
        }
        int pending = 0;                    // rows parked in acc
        uint64_t* wfull = fullbar + (w - g.twmax) * kStage;
        uint64_t* wempty = emptybar + (w - g.twmax) * kStage;
        // one batch of kWB (= 2) parked rows starting at row r; Q0 = r & 3 is an immediate so that the barrier ids are
        auto batch = [&](auto q0_tag, int r, int cnt) {
            constexpr int Q0 = decltype(q0_tag)::value;
            if (kVertMbar) {
                mbar_wait_b(wfull + Q0, (r >> 2) & 1, err);
                if (cnt > 1) mbar_wait_b(wfull + Q0 + 1, (r >> 2) & 1, err);
            } else {
                named_bar_sync(BAR_FULL + Q0, nboth);
                if (cnt > 1) named_bar_sync(BAR_FULL + Q0 + 1, nboth);
            }
#pragma unroll
            for (int c = 0; c < kWC; c++) {
                if (c >= ncol) break;
                uint16_t* sc = stage + c * Dp + Q0 * ringSlot;
                if (edge_side[c] >= 0) {
                    // edge column: the agent computed the incoming-diagonal path; add its L to the parked partial sum
                    const uint16_t* sB = xringbase + size_t(edge_side[c]) * kStage * Dp + lane * 2 * N;
#pragma unroll
                    for (int q = 0; q < kWB; q++) {
                        if (q < cnt && active) {
                            uint32_t P[N], B[N];
                            ld_regs<N>(sc + q * ringSlot + lane * 2 * N, P);
                            ld_regs<N>(sB + (Q0 + q) * Dp, B);
#pragma unroll
                            for (int i = 0; i < N; i++) {
                                P[i] = __vminu2(P[i] + B[i], kMaxCostX2);
                                if (!FULL) {
                                    const int k = lane * N + i;
                                    if (k >= g.w.D) P[i] = 0xFFFFFFFFu;
                                    else if (k + wc.Dh >= g.w.D) P[i] |= 0xFFFF0000u;
                                }
                            }
                            st_regs<N>(sc + q * ringSlot + lane * 2 * N, P);
                        }
                    }
                }
                if (wta_on && wc.f <= 0) {
                    __syncwarp();
                    for (int q = 0; q < cnt; q++) {
                        const int d = wta_slow<N>(sc + q * ringSlot, g.w, wc, x + c, lane, active, kptr + q * dStride);
                        if (lane == 0) dptr[q * dStride + c] = int16_t(d);
                    }
                    __syncwarp();
                }
            }
            // (a row past the end lands in a lane >= cnt of the flush)
            if (wta_on && wc.f > 0) {
                if constexpr (kWtaTogether) wta_vec<N, kWC>(stage + Q0 * ringSlot, ringSlot, r, ncol, g.w, wc, lane, active, acc);
                else {
#pragma unroll
                    for (int c = 0; c < kWC; c++) {
                        if (c >= ncol) break;
                        WtaAcc (&a1)[1] = *reinterpret_cast<WtaAcc (*)[1]>(&acc[c]);
                        wta_vec<N, 1>(stage + c * Dp + Q0 * ringSlot, ringSlot, r, 1, g.w, wc, lane, active, a1);
                    }
                }
            }
            if (wta_on) {
                if (wc.f > 0) {
                    pending += cnt;
                    if (((r + kWB) & 31) == 0 || r + kWB >= H) {
#pragma unroll
                        for (int c = 0; c < kWC; c++)
                            if (c < ncol) wta_flush32(acc[c], pending, g.w, wc, x + c, lane, dptr + c, kptr, dStride);
                        dptr += pending * dStride; kptr += pending * dStride;
                        pending = 0;
                    }
                } else {
                    dptr += cnt * dStride; kptr += cnt * dStride;
                }
            }
            if (kVertMbar) {
                __syncwarp();
                if (lane == 0) {
                    if (r + kStage < H) mbar_arrive(wempty + Q0);
                    if (cnt > 1 && r + 1 + kStage < H) mbar_arrive(wempty + Q0 + 1);
                }
            } else {
                if (FOLD) {
                    if (r + 3 < H) named_bar_arrive(BAR_ROWX + (Q0 == 0 ? 1 : 0), nboth);     // row r + 3 = 3 or 1 (mod 4)
                } else {
                    if (r + kStage < H) named_bar_arrive(BAR_EMPTY + Q0, nboth);
                    if (cnt > 1 && r + 1 + kStage < H) named_bar_arrive(BAR_EMPTY + Q0 + 1, nboth);
                }
            }
        };
        static_assert(kWB == 2 && kStage == 4, "two batches per stage-ring revolution");
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
    // direction 0: predecessor column x-1 (slot j); direction 1: predecessor column x+1 (slot j+2).
    // The left-edge warp runs direction 1 first (it publishes it), every other warp direction 0 first.
    const int dirA = j == 0 ? 1 : 0, dirB = 1 - dirA;
    const int slotA = dirA == 0 ? j : j + 2, slotB = dirB == 0 ? j : j + 2;
    const bool edge = exchange_on && ((j == 0 && b > 0) || (j == TW - 1 && j != 0 && b < n - 1));   // publishes dirA, consumes dirB
    const bool use_agent = g.agents != 0;
    // edge mode: 1 = poll the neighbour's record in this warp, 2 = an agent delivers it (halo barrier), 3 = an agent
    // computes the incoming-diagonal step itself (WTA sweeps)
    const int edge_mode = !edge ? 0 : (!use_agent ? 1 : (DO_WTA ? 3 : 2));
    const int halo_bar = BAR_HALO + (j == 0 ? 0 : 1);
    const uint16_t* xin = xringbase + size_t(j == 0 ? 0 : 1) * kXbufGen * Dp + lo;   // what my agent received
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
    // HALO > 1: the records come from the columns HALO columns inside the strip's edges (inner warps: step A is their
    // from-the-left step, step B their from-the-right step); the host guarantees TW > HALO
    const bool pubA = HALO > 1 && exchange_on && g.agents && b < n - 1 && j == TW - HALO;
    const bool pubB = HALO > 1 && exchange_on && g.agents && b > 0 && j == HALO - 1;
    uint2* pubA_base = xrec(xbuf, n, Dp, 0, b, 0, XG) + lane * N;
    uint2* pubB_base = xrec(xbuf, n, Dp, 1, b, 0, XG) + lane * N;
    const uint2* con_base = xrec(xbuf, n, Dp, dirB, edge ? nb : b, 0) + lane * N;
    const int gen_stride = Dp / 2;

    const uint16_t* gC = Cvol + (size_t(ystart) * W1 + x) * Dp + lo;      // next row to fetch
    const uint16_t* gSin = Svol + (size_t(ystart) * W1 + x) * Dp + lo;
    uint16_t* gSout = Svol + (size_t(ystart) * W1 + x) * Dp + lo;          // row being computed (first pass of MODE_HH)
    const int Dh = Dp >> 1;
    // cp.async rings: [warp][slot][Dp], slot = row & (ring - 1).  Group G_r (committed at the top of row r) carries C
    // of row r+RING-1 and S_h of row r+RING-1; the prologue commits C rows 0..RING-2 and S_h rows 0..RING-2.
    uint16_t* ring = ringbase + size_t(j) * ringColStride + lo;
    uint16_t* sring = sringbase + size_t(j) * ringColStride + lo;
    auto issue_c = [&](int row_) {
        if (row_ < H && active) cp_async_lane<N>(ring + (row_ & (RING - 1)) * Dp, gC);
        gC += rowStride;
    };
    auto issue_s = [&](int row_) {
        if (row_ < H && active) cp_async_lane<N>(sring + (row_ & (RING - 1)) * Dp, gSin);
        gSin += rowStride;
    };
    // bulk-copy producer: one lane of a middle column's warp (the edge warps carry the exchange); it refills the slot of
    // row r - 1 at the top of row r, i.e. after the row barrier every reader of that slot has passed
    const bool producer = kVertTma && j == (TW >> 1) && lane == 0;
    const uint16_t* gCs = Cvol + (size_t(ystart) * W1 + x0) * Dp;      // strip start of the next row to fetch
    const uint16_t* gSs = Svol + (size_t(ystart) * W1 + x0) * Dp;
    const uint32_t stripBytes = uint32_t(TW) * Dp * sizeof(uint16_t);
    auto issue_bulk = [&](int row_) {
        if (row_ < H) {
            const int sl = row_ & (RING - 1);
            mbar_expect_tx(ringbar + sl, 2 * stripBytes);
            bulk_g2s(ringbase + size_t(sl) * ringSlotStride, gCs, stripBytes, ringbar + sl);
            bulk_g2s(sringbase + size_t(sl) * ringSlotStride, gSs, stripBytes, ringbar + sl);
        }
        gCs += rowStride; gSs += rowStride;
    };
    if (kVertTma) {
        if (producer)
            for (int i = 0; i < RING - 1; i++) issue_bulk(i);
    } else {
#pragma unroll
        for (int i = 0; i < RING - 1; i++) issue_s(i);
#pragma unroll
        for (int i = 0; i < RING - 1; i++) { issue_c(i); cp_async_commit(); }
        if (nag3) {            // agents read this column's C ring: make row 0 visible to them
            cp_async_wait<RING - 2>();
            named_bar_sync(BAR_ROW, nrow);
        }
    }

    uint64_t* pfull = fullbar + (j / kWC) * kStage;
    uint64_t* pempty = emptybar + (j / kWC) * kStage;
    uint32_t LtV[N];
    uint2 pre[N];        // edge warps: the neighbour's record of the previous row, requested one row early
#pragma unroll
    for (int q = 0; q < N; q++) { LtV[q] = 0; pre[q] = make_uint2(0u, 0u); }
    bool dead = false;

    // one row; Q = r mod 4 (ring slot, record generation, stage slot; parity = Q & 1); EDGE = this warp exchanges
    auto row = [&](auto q_tag, auto edge_tag, int r) {
        constexpr int Q = decltype(q_tag)::value;
        constexpr int EMODE = decltype(edge_tag)::value;
        constexpr bool EDGE = EMODE != 0;
        constexpr bool NO_B = EMODE == 3;      // the agent owns the incoming-diagonal step
        constexpr int PAR = Q & 1;
        if (SPLIT && r > 0) mbar_wait_b(rowbar, (r - 1) & 1, err);      // every warp of the strip has stored its states of row r - 1
        if (kVertTma) { if (producer) issue_bulk(r + RING - 1); }
        else { issue_c(r + RING - 1); issue_s(r + RING - 1); cp_async_commit(); }
        uint32_t Cc[N], Sc[N], LtA[N], LtB[N], LnA[N], LnV[N], LnB[N];
        // incoming diagonal of the previous row: fire the loads now, inspect the tags right before step B
        const bool consume = EDGE && r > 0;
        if (EMODE == 1) {
            if (consume && active) {
                const uint2* rec = con_base + ((Q + kXbufGen - 1) & (kXbufGen - 1)) * gen_stride;
#pragma unroll
                for (int q = 0; q < N; q++) pre[q] = ld_volatile_v2(rec + q);
            }
        }
        if (kVertTma) mbar_wait(ringbar + (r & (RING - 1)), (r >> kRingShift) & 1);     // both bulk copies of row r have landed
        else cp_async_wait<RING - 1>();     // this thread's copies of row r have landed (each lane reads only its own bytes)
        if (active) {
            ld_regs<N>(ring + (r & (RING - 1)) * ringSlotStride, Cc); ld_regs<N>(rdA[PAR ^ 1], LtA);
            if (!EDGE) ld_regs<N>(rdB[PAR ^ 1], LtB);
        } else {
#pragma unroll
            for (int q = 0; q < N; q++) { Cc[q] = kMaxCostX2; LtA[q] = 0; LtB[q] = 0; }
        }
        // ---- step A (the direction this warp publishes)
        path_step<N>(Cc, LtA, LnA, lc);
        if (active) st_regs<N>(wrA[PAR], LtA);
        if (EDGE && HALO == 1) {
            if (active) {
                uint2* rec = pub_base + Q * gen_stride;
#pragma unroll
                for (int q = 0; q < N; q++) st_volatile_v2(rec + q, LtA[q], uint32_t(r + 1));
            }
        }
        if (HALO > 1 && pubA && active) {
            uint2* rec = pubA_base + (r & (XG - 1)) * gen_stride;
#pragma unroll
            for (int q = 0; q < N; q++) st_volatile_v2(rec + q, LtA[q], uint32_t(r + 1));
        }
        // ---- vertical path: registers only (SPLIT: after the row's arrival, nobody waits for it)
        if (!SPLIT) path_step<N>(Cc, LtV, LnV, lc);
        // ---- step B
        if (EDGE && !NO_B) {
            if (consume && EMODE == 2) {
                // my agent (HALO > 1: the halo producer of row r - 1) has dropped the state of row r - 1 into shared memory
                named_bar_sync(HALO > 1 ? 12 + 2 * (j == 0 ? 0 : 1) + ((Q & 1) ^ 1) : halo_bar, 64);
                if (active) ld_regs<N>(xin + ((Q + kXbufGen - 1) & (kXbufGen - 1)) * Dp, LtB);
                else {
#pragma unroll
                    for (int q = 0; q < N; q++) LtB[q] = 0;
                }
            } else if (consume) {
                bool ok = true;
#pragma unroll
                for (int q = 0; q < N; q++) LtB[q] = 0;
                if (active) {
#pragma unroll
                    for (int q = 0; q < N; q++) { LtB[q] = pre[q].x; ok = ok && pre[q].y == uint32_t(r); }
                }
                if (!__all_sync(kFullMask, ok)) {       // the neighbour is late: poll
                    if ((g.debug_flags & 4) && lane == 0) atomicAdd(err + 1, 1);
                    const uint2* rec = con_base + ((Q + kXbufGen - 1) & (kXbufGen - 1)) * gen_stride;
                    if (active && !dead) {
                        const long long t0 = clock64();
                        int spins = 0;
                        while (true) {
                            ok = true;
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
                }
            } else {
                if (active) ld_regs<N>(rdB[PAR ^ 1], LtB);
                else {
#pragma unroll
                    for (int q = 0; q < N; q++) LtB[q] = 0;
                }
            }
        }
        if (!NO_B) {
            path_step<N>(Cc, LtB, LnB, lc);
            if (active) st_regs<N>(wrB[PAR], LtB);
            if (HALO > 1 && pubB && active) {
                uint2* rec = pubB_base + (r & (XG - 1)) * gen_stride;
#pragma unroll
                for (int q = 0; q < N; q++) st_volatile_v2(rec + q, LtB[q], uint32_t(r + 1));
            }
        } else {
#pragma unroll
            for (int q = 0; q < N; q++) LnB[q] = 0;
        }
        if (SPLIT) {
            // both diagonal states of the row are stored: that is all the neighbours (and, with C of the next row landed, the
            // agents) wait for.  The rest of the row overlaps their waiting.
            cp_async_wait<RING - 2>();
            __syncwarp();
            if (lane == 0) mbar_arrive(rowbar);
            path_step<N>(Cc, LtV, LnV, lc);
        }
        // ---- S = sat(S_h + L_v + L_A + L_B)
        if (active) ld_regs<N>(sring + (r & (RING - 1)) * ringSlotStride, Sc);
        else {
#pragma unroll
            for (int q = 0; q < N; q++) Sc[q] = 0;
        }
        uint32_t S[N];
#pragma unroll
        for (int q = 0; q < N; q++) {
            if (CLAMP_EACH) {
                uint32_t t = __vminu2(Sc[q] + LnV[q], kMaxCostX2);
                t = __vminu2(t + LnA[q], kMaxCostX2);
                S[q] = __vminu2(t + LnB[q], kMaxCostX2);
            } else {
                S[q] = __vminu2(Sc[q] + LnV[q] + LnA[q] + LnB[q], kMaxCostX2);
            }
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
            if (kVertMbar) {
                if (r >= kStage) mbar_wait_b(pempty + Q, ((r >> 2) + 1) & 1, err);    // my WTA warp has taken row r - kStage
                if (active) st_regs<N>(stage + Q * ringSlot + lo, S);
                __syncwarp();
                if (lane == 0) mbar_arrive(pfull + Q);
            } else {
                if (!FOLD && r >= kStage) named_bar_sync(BAR_EMPTY + Q, nboth);    // the WTA warps have taken row r - kStage
                if (active) st_regs<N>(stage + Q * ringSlot + lo, S);
                named_bar_arrive(BAR_FULL + Q, nboth);
            }
        } else {
            if (active) st_regs<N>(gSout, S);
            gSout += rowStride;
        }
        if (!SPLIT) {
            if (DO_WTA && !kVertTma) cp_async_wait<RING - 2>();   // C of the NEXT row is complete before the barrier: the agents read it
            if (FOLD && (Q & 1)) named_bar_sync(BAR_ROWX + (Q >> 1), r >= 3 ? nboth : nrow);
            else named_bar_sync(BAR_ROW, nrow);
        }
    };

    auto sweep = [&](auto edge_tag) {
        int r = 0;
        for (; r + 3 < H; r += 4) {
            row(std::integral_constant<int, 0>{}, edge_tag, r);
            row(std::integral_constant<int, 1>{}, edge_tag, r + 1);
            row(std::integral_constant<int, 2>{}, edge_tag, r + 2);
            row(std::integral_constant<int, 3>{}, edge_tag, r + 3);
        }
        const int rem = H - r;
        if (rem > 0) row(std::integral_constant<int, 0>{}, edge_tag, r);
        if (rem > 1) row(std::integral_constant<int, 1>{}, edge_tag, r + 1);
        if (rem > 2) row(std::integral_constant<int, 2>{}, edge_tag, r + 2);
    };
    if (edge_mode == 0) sweep(std::integral_constant<int, 0>{});
    else if (edge_mode == 1) sweep(std::integral_constant<int, 1>{});
    else if (DO_WTA) sweep(std::integral_constant<int, DO_WTA ? 3 : 1>{});
    else sweep(std::integral_constant<int, DO_WTA ? 1 : 2>{});
    cp_async_wait<0>();
}

}  // namespace b200sgm
