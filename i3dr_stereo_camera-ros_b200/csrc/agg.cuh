// Host launchers of the aggregation stages (k_horiz + cooperative k_vert, or the generic per-direction kernels),
// templated on N; instantiated once per N in agg_n*.cu so that the instantiations compile in parallel.
#pragma once
#include "engine_internal.h"
#include "k_path.cuh"
#include "k_wta.cuh"
#include "k_fused.cuh"

template <int N>
int launch_paths_generic(b200sgm_engine* h, Lane& ln, const Eff& e, cudaStream_t st)
{
    static const int dirs_sgbm[5][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}};
    static const int dirs_hh[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
    const int nd = e.mode == B200SGM_MODE_HH ? 8 : 5;
    for (int r = 0; r < nd; r++) {
        PathGeom g;
        g.W1 = e.W1; g.H = e.H; g.Dp = e.Dp;
        g.dx = e.mode == B200SGM_MODE_HH ? dirs_hh[r][0] : dirs_sgbm[r][0];
        g.dy = e.mode == B200SGM_MODE_HH ? dirs_hh[r][1] : dirs_sgbm[r][1];
        g.nchains = chain_count(e.W1, e.H, g.dx, g.dy);
        g.P1 = e.P1; g.P2 = e.P2;
        const int wpb = 4;
        dim3 grid((g.nchains + wpb - 1) / wpb), block(32 * wpb);
        if (r == 0) k_path_generic<N, true><<<grid, block, 0, st>>>(ln.C, ln.S, g);
        else k_path_generic<N, false><<<grid, block, 0, st>>>(ln.C, ln.S, g);
        LAUNCH_CHECK(h);
    }
    WtaGeom wg{e.W, e.H, e.W1, e.minX1, e.minD, e.D, e.Dp, e.uniq, e.d12, e.INVALID};
    const long long npix = (long long)e.W1 * e.H;
    k_wta<N><<<unsigned((npix + 7) / 8), 256, 0, st>>>(ln.S, wg, ln.disp_wta, ln.disp2key);
    LAUNCH_CHECK(h);
    return B200SGM_OK;
}

// ---- fused path: k_horiz + cooperative k_vert ---------------------------------------------------------
struct VertPlan { bool ok; int nstrips, twmax; size_t smem; int halo; };

// Ring depth and thread budget per register count: up to 256 disparities 8 rows of C / S_h in flight and 64 registers per
// thread; beyond that a row is at least twice the bytes, so half the depth hides the same DRAM latency, and fewer, fatter
// threads keep the N = 8 row loop out of local memory.
template <int N> struct VertCfg { static constexpr int RING = N >= 8 ? 4 : 8, MAXT = vert_max_threads(N); };
// development switch: B200SGM_VERT_RING=4 halves the ring depth of the N <= 4 instantiations
inline int vert_ring(int n_regs)
{
    static const int env = [] { const char* v = getenv("B200SGM_VERT_RING"); return v ? atoi(v) : 0; }();
    return n_regs >= 8 ? 4 : (env == 4 ? 4 : 8);
}
// threads of a sweep with `tw` columns per strip: one path warp per column, one WTA warp per kWC columns
inline int vert_threads(int tw, bool do_wta) { return 32 * (tw + (do_wta ? (tw + kWC - 1) / kWC : 0)); }

template <int N>
inline VertPlan plan_vert(const b200sgm_engine* h, const Eff& e)
{
    VertPlan p{false, 0, 0, 0, 1};
    if (e.W1 < 2) return p;
    if (kVertTma && e.Dp % 8 != 0) return p;      // bulk copies move multiples of 16 bytes: odd paddings take the hybrid path
    // widest strip a CTA can take: threads (a path warp per column, a WTA warp per kWC columns, two agents) and shared memory
    int twcap = kVertMaxWarps;
    const size_t smem_cap = kVertCps == 1 ? size_t(h->max_smem_optin) : (size_t(h->max_smem_sm) / kVertCps - 1024);
    const int slots = h->num_sms * kVertCps;       // co-resident strips of the GPU
    while (twcap > 1 && (vert_threads(twcap, true) + 64 > VertCfg<N>::MAXT || vert_smem_bytes(twcap, e.Dp, vert_ring(N)) > smem_cap)) twcap--;
    // K = sweeps of this geometry that fit the GPU side by side; each gets num_sms / K strips (K = 1: one strip per SM)
    const int nmin = (e.W1 + twcap - 1) / twcap;
    // (an engine with a single lane never has two frames in flight: it spreads every sweep over all SMs, which is 10 % faster
    // for a frame on its own; B200SGM_VERT_ALL_SMS forces that for measurements)
    static const bool wide_env = getenv("B200SGM_VERT_ALL_SMS") != nullptr;
    const bool wide = wide_env || h->lanes.size() < 2;
    const int K = wide ? 1 : std::max(1, slots / std::max(1, nmin));
    int n = std::min(slots / K, e.W1 / 2);
    n = std::max(1, std::min(n, kMaxStrips));
    // strips too narrow for the halo producers (below): somewhat fewer strips make every strip kHaloMax + 1 columns wide -- a
    // narrow sweep is bound by the hand-over between strips, not by the SMs it leaves idle
    static const bool halo_fit_env = [] { const char* v = getenv("B200SGM_VERT_HALO_FIT"); return !v || atoi(v) != 0; }();
    if (N <= 2 && halo_fit_env && n > 1 && e.W1 / n <= kHaloMax && e.W1 / (kHaloMax + 1) * 4 >= n * 3) n = e.W1 / (kHaloMax + 1);
    int tw = (e.W1 + n - 1) / n;
    // wider than one co-resident wave of strips (or than a CTA has warps): use the hybrid path
    if (tw > kVertMaxWarps || vert_threads(tw, true) > VertCfg<N>::MAXT) return p;
    p.nstrips = n; p.twmax = tw;
    p.smem = vert_smem_bytes(tw, e.Dp, vert_ring(N));
    p.ok = p.smem <= smem_cap;
    // Narrow strips (up to 128 disparities) sweep a row faster than a record crosses the L2: their agents advance the incoming
    // diagonals through a halo of kHaloMax - 1 columns, which gives the records kHaloMax - 1 rows to arrive (k_vert's HALO).
    // Needs the agents, strips wider than the halo, and room for the halo rings.  Only for narrow strips (a single frame spread
    // over all SMs): wide strips are bound by their own work, and the halo's extra steps and C reads then cost more than the
    // hand-over (1280x1024x128 MODE_HH on one lane: 1.115 -> 0.834 ms; four lanes of 16-column strips: 1050 -> 970 frames/s;
    // 256 disparities on 7-column strips, 1280x720: 0.452 -> 0.640 ms -- those strips are bound by their own work).
    static const bool halo_env = [] { const char* v = getenv("B200SGM_VERT_HALO"); return !v || atoi(v) != 0; }();
    static const int halo_tw = [] { const char* v = getenv("B200SGM_VERT_HALO_TW"); return v ? atoi(v) : 10; }();   // widest strip that gets halo agents
    static const bool plain_agents = getenv("B200SGM_NO_AGENTS") == nullptr && getenv("B200SGM_DEBUG_VERT") == nullptr;
    if (p.ok && N <= 2 && halo_env && plain_agents && kVertCps == 1 && !kVertTma && vert_ring(N) == VertCfg<N>::RING && n > 1 &&
        e.W1 / n > kHaloMax && tw <= halo_tw && vert_threads(tw, true) + 192 <= VertCfg<N>::MAXT && vert_smem_bytes(tw, e.Dp, vert_ring(N), kHaloMax) <= smem_cap) {
        p.halo = kHaloMax;
        p.smem = vert_smem_bytes(tw, e.Dp, vert_ring(N), kHaloMax);
    }
    return p;
}

template <int N, int RING, bool UP, bool DO_WTA, bool FULL, bool CLAMP_EACH, int HALO = 1>
int launch_vert_r(b200sgm_engine* h, Lane& ln, const Eff& e, const VertPlan& vp, cudaStream_t st)
{
    VertGeom g;
    g.w = WtaGeom{e.W, e.H, e.W1, e.minX1, e.minD, e.D, e.Dp, e.uniq, e.d12, e.INVALID};
    g.nstrips = vp.nstrips; g.twmax = vp.twmax;
    g.P1 = e.P1; g.P2 = e.P2;
    g.spin_limit = (long long)h->clock_khz * 500;   // ~0.5 s of SM clock ticks
    { static const int dbg = [] { const char* v = getenv("B200SGM_DEBUG_VERT"); return v ? atoi(v) : 0; }(); g.debug_flags = dbg; }
    // two agent warps per CTA when they fit next to the column warps (1024 threads per CTA)
    int nthreads = vert_threads(vp.twmax, DO_WTA);
    { static const bool no_agents = getenv("B200SGM_NO_AGENTS") != nullptr; g.agents = (!no_agents && nthreads + 64 <= VertCfg<N>::MAXT) ? 1 : 0; }
    if (g.agents) nthreads += 64;
    if (HALO > 1) nthreads += 128;                 // the four halo producers
    auto kern = k_vert<N, RING, UP, DO_WTA, FULL, CLAMP_EACH, HALO>;
    if (HALO > 1 && !g.agents) return fail(h, B200SGM_ECUDA, "internal: halo sweep planned without agents");
    {
        static std::atomic<unsigned long long> attr_done{0};   // per instantiation and device: raise the dynamic shared-memory limit once
        const unsigned long long bit = 1ull << (h->device & 63);
        if (!(attr_done.load() & bit)) {
            CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
            // largest shared-memory carve-out, so that what this sweep leaves free can hold CTAs of another frame's kernels
            CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            attr_done.fetch_or(bit);
        }
    }
    CUDA_TRY(h, cudaMemsetAsync(ln.xbuf, 0, size_t(2) * vp.nstrips * (HALO > 1 ? kHaloGen : kXbufGen) * (e.Dp / 2) * sizeof(uint2), st));
    const uint16_t* Cp = ln.C; uint16_t* Sp = ln.S; int16_t* dp = ln.disp_wta; uint32_t* kp = ln.disp2key;
    uint2* xb = ln.xbuf; int* er = ln.d_err;
    void* args[] = {(void*)&Cp, (void*)&Sp, (void*)&g, (void*)&dp, (void*)&kp, (void*)&xb, (void*)&er};
    {
        CoopGate& gate = coop_gate(h->device);
        std::lock_guard<std::mutex> lk(gate.mu);
        if (!gate.ev[0])
            for (auto& ev : gate.ev) CUDA_TRY(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        static const int kmax = [] { const char* v = getenv("B200SGM_VERT_CONCURRENT"); return v ? std::max(1, atoi(v)) : CoopGate::kRing - 1; }();
        const int K = std::max(1, std::min(std::min(h->num_sms * kVertCps / vp.nstrips, CoopGate::kRing - 1), kmax));
        const unsigned long long i = gate.count;
        if (vp.nstrips != gate.last_n) {
            for (unsigned long long q = i > CoopGate::kRing ? i - CoopGate::kRing : 0; q < i; q++)
                CUDA_TRY(h, cudaStreamWaitEvent(st, gate.ev[q % CoopGate::kRing], 0));
        } else if (i >= (unsigned long long)K) {
            CUDA_TRY(h, cudaStreamWaitEvent(st, gate.ev[(i - K) % CoopGate::kRing], 0));
        }
        // The sweep only needs its strips co-resident (no grid-wide barrier).  A plain launch gets that as well -- the gate bounds
        // the sweeps in flight and no other kernel of the engine waits on anything.
        static const bool plain = [] { const char* v = getenv("B200SGM_VERT_PLAIN"); return v && atoi(v) != 0; }();
        if (plain) {
            kern<<<dim3(vp.nstrips), dim3(nthreads), vp.smem, st>>>(Cp, Sp, g, dp, kp, xb, er);
            CUDA_TRY(h, cudaGetLastError());
        } else {
            CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)kern, dim3(vp.nstrips), dim3(nthreads), args, vp.smem, st));
        }
        h->launches++;
        CUDA_TRY(h, cudaEventRecord(gate.ev[i % CoopGate::kRing], st));
        gate.count = i + 1;
        gate.last_n = vp.nstrips;
    }
    return B200SGM_OK;
}

template <int N, bool UP, bool DO_WTA, bool FULL, bool CLAMP_EACH>
int launch_vert_t(b200sgm_engine* h, Lane& ln, const Eff& e, const VertPlan& vp, cudaStream_t st)
{
    if constexpr (N <= 4) {
        if (vert_ring(N) == 4) return launch_vert_r<N, 4, UP, DO_WTA, FULL, CLAMP_EACH>(h, ln, e, vp, st);
    }
    if constexpr (N <= 2) {
        if (vp.halo == kHaloMax) return launch_vert_r<N, VertCfg<N>::RING, UP, DO_WTA, FULL, CLAMP_EACH, kHaloMax>(h, ln, e, vp, st);
    }
    return launch_vert_r<N, VertCfg<N>::RING, UP, DO_WTA, FULL, CLAMP_EACH>(h, ln, e, vp, st);
}

template <int N, bool UP, bool DO_WTA>
int launch_vert(b200sgm_engine* h, Lane& ln, const Eff& e, const VertPlan& vp, cudaStream_t st)
{
    const bool full = e.Dp == e.D && e.D == 64 * N;
    // worst-case cost of a cell: bs^2 * (2*ftzero + 63) (A.5 value bounds); one final clamp is enough when
    // kMaxCost + 3 * (Cmax + P2) cannot wrap 16 bits
    const long long bs = 2 * e.SW2 + 1;
    const long long cmax = bs * bs * (2 * e.ftzero + 63) + e.P2;
    // S_h arrives unclamped (<= 2*cmax) in MODE_SGBM / first sweep, clamped (<= kMaxCost) in the second sweep of MODE_HH
    const bool clamp_each = kMaxCost + 3 * cmax > 65535 || 5 * cmax > 65535;
    if (full) {
        if (clamp_each) return launch_vert_t<N, UP, DO_WTA, true, true>(h, ln, e, vp, st);
        return launch_vert_t<N, UP, DO_WTA, true, false>(h, ln, e, vp, st);
    }
    return launch_vert_t<N, UP, DO_WTA, false, true>(h, ln, e, vp, st);
}

// part: 0 = horizontal pair + vertical sweep(s) (a frame), 1 = horizontal pair only, 2 = vertical sweep(s) only
template <int N>
int launch_fused(b200sgm_engine* h, Lane& ln, const Eff& e, cudaStream_t st, bool hybrid, int part)
{
    int wpb = 2;
    while (wpb > 1 && size_t(wpb) * horiz_smem_per_warp(e.Dp) > 200 * 1024) wpb /= 2;
    const size_t hsmem = size_t(wpb) * horiz_smem_per_warp(e.Dp);
    if (hsmem > 200 * 1024) return fail(h, B200SGM_EINVAL, "numDisparities too large for the horizontal kernel");
    if (part != 2) {
        const bool full = e.Dp == 64 * N;
        const long long bs = 2 * e.SW2 + 1;
        const bool clamp = !full || 2 * (bs * bs * (2 * e.ftzero + 63) + e.P2) > 65535;
        auto kern = !full ? k_horiz<N, false, true> : (clamp ? k_horiz<N, true, true> : k_horiz<N, true, false>);
        static std::atomic<unsigned long long> attr_done[3] = {{0}, {0}, {0}};
        const int ki = !full ? 0 : (clamp ? 1 : 2);
        const unsigned long long bit = 1ull << (h->device & 63);
        if (!(attr_done[ki].load() & bit)) {
            CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            attr_done[ki].fetch_or(bit);
        }
        kern<<<(e.H + wpb - 1) / wpb, 32 * wpb, hsmem, st>>>(ln.C, ln.S, ln.ckpt, e.W1, e.H, e.Dp, e.D, e.P1, e.P2,
                                                              reinterpret_cast<unsigned*>(ln.d_err) + 3);
        LAUNCH_CHECK(h);
    }
    if (part == 0) prof_mark(h, ln, 3, st);
    if (part == 1) return B200SGM_OK;
    const VertPlan vp = plan_vert<N>(h, e);
    if (hybrid || !vp.ok) {
        static const int dirs_sgbm[3][2] = {{1, 1}, {0, 1}, {-1, 1}};
        static const int dirs_hh[6][2] = {{1, 1}, {0, 1}, {-1, 1}, {-1, -1}, {0, -1}, {1, -1}};
        const int nd = e.mode == B200SGM_MODE_HH ? 6 : 3;
        for (int r = 0; r < nd; r++) {
            PathGeom g;
            g.W1 = e.W1; g.H = e.H; g.Dp = e.Dp;
            g.dx = e.mode == B200SGM_MODE_HH ? dirs_hh[r][0] : dirs_sgbm[r][0];
            g.dy = e.mode == B200SGM_MODE_HH ? dirs_hh[r][1] : dirs_sgbm[r][1];
            g.nchains = chain_count(e.W1, e.H, g.dx, g.dy);
            g.P1 = e.P1; g.P2 = e.P2;
            k_path_generic<N, false><<<(g.nchains + 3) / 4, 128, 0, st>>>(ln.C, ln.S, g);
            LAUNCH_CHECK(h);
        }
        WtaGeom wg{e.W, e.H, e.W1, e.minX1, e.minD, e.D, e.Dp, e.uniq, e.d12, e.INVALID};
        const long long npix = (long long)e.W1 * e.H;
        k_wta<N><<<unsigned((npix + 7) / 8), 256, 0, st>>>(ln.S, wg, ln.disp_wta, ln.disp2key);
        LAUNCH_CHECK(h);
        return B200SGM_OK;
    }
    int rc;
    if (e.mode == B200SGM_MODE_HH) {
        rc = launch_vert<N, false, false>(h, ln, e, vp, st);
        if (rc) return rc;
        rc = launch_vert<N, true, true>(h, ln, e, vp, st);
    } else {
        rc = launch_vert<N, false, true>(h, ln, e, vp, st);
    }
    return rc;
}

template <int N>
int launch_agg_n(b200sgm_engine* h, Lane& ln, const Eff& e, cudaStream_t st, int part)
{
    if (h->path == 1) { prof_mark(h, ln, 3, st); return launch_paths_generic<N>(h, ln, e, st); }
    return launch_fused<N>(h, ln, e, st, h->path == 2, part);
}

