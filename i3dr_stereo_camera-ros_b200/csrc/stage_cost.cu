// Launchers of the prefilter and cost stages (A.2 - A.4); kernels in k_cost.cuh.
#include <cstdlib>
#include "stages.h"
#include "k_cost.cuh"

namespace b200sgm {

#define CUDA_TRY_(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return e__; } while (0)
#define LAUNCH_CHECK_() do { ++*launches; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return e__; } while (0)

void launch_prefilter(const uint8_t* dL, size_t lp, const uint8_t* dR, size_t rp, int W, int H, int ftzero, Feat* featL, Feat* featR,
                      cudaStream_t st)
{
    dim3 block(128), grid((W + 128 * kPfPPT - 1) / (128 * kPfPPT), H, 2);
    k_prefilter<<<grid, block, 0, st>>>(dL, lp, dR, rp, W, H, ftzero, featL, featR);
}

static size_t cost_smem_bytes(int TX, int DCP, int SW2)
{
    const int bs = 2 * SW2 + 1;
    return size_t((TX + 2 * SW2) + bs * TX + TX) * DCP * 4;
}

cudaError_t launch_cost(const Feat* fl, const Feat* fr, uint16_t* C, const Eff& e, bool generic_only, int num_sms, cudaStream_t st,
                        int* launches, const char** errmsg)
{
    const int W = e.W, H = e.H;
    static const int cost_variant = [] { const char* v = getenv("B200SGM_COST_VARIANT"); return v ? atoi(v) : 2; }();
    if (e.SW2 <= 10 && !generic_only && 2 * e.ftzero + 63 <= 255 && cost_variant == 2) {
        // register-tiled kernel: 128-column tiles, 2 CTAs per SM.  Rows per segment: the split of the image height that
        // minimises (waves of CTAs) x (rows a CTA walks, incl. the blockSize - 1 warm-up rows)
        const bool nopad = e.Dp == e.D && (e.Dp / 2) % kCfDCP == 0;
        const int TX = kC2TXH - 2 * e.SW2;
        const int ntx = (e.W1 + TX - 1) / TX, ndc = (e.Dp / 2 + kCfDCP - 1) / kCfDCP;
        const size_t smem = cost_tile2_smem(e.SW2);
        const int per_sm = smem * 2 + 2048 <= size_t(227) * 1024 ? 2 : 1;
        int best_rs = H;
        long long best_cost = -1;
        for (int nseg = 1; nseg <= 64 && nseg <= H; nseg++) {
            const int rs = (H + nseg - 1) / nseg;
            const long long ctas = (long long)ntx * ndc * ((H + rs - 1) / rs);
            const long long waves = (ctas + (long long)per_sm * num_sms - 1) / ((long long)per_sm * num_sms);
            const long long cost = waves * (rs + 2 * e.SW2);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rs = rs; }
        }
        CostFastGeom fg{W, H, e.W1, e.minX1, e.minD, e.D, e.Dp, e.SW2, best_rs};
        dim3 grid(ntx, ndc, (H + best_rs - 1) / best_rs);
        void (*kern)(const Feat*, const Feat*, uint16_t*, CostFastGeom);
        if (e.SW2 == 4) kern = nopad ? k_cost_tile2<4, true> : k_cost_tile2<4, false>;
        else if (e.SW2 == 2) kern = nopad ? k_cost_tile2<2, true> : k_cost_tile2<2, false>;
        else if (e.SW2 == 10) kern = nopad ? k_cost_tile2<10, true> : k_cost_tile2<10, false>;   // blockSize 21: the reference's launch default
        else kern = k_cost_tile2<0, false>;
        CUDA_TRY_(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kern<<<grid, 256, smem, st>>>(fl, fr, C, fg);
        LAUNCH_CHECK_();
    } else if (e.SW2 <= 10 && !generic_only) {
        CostFastGeom fg{W, H, e.W1, e.minX1, e.minD, e.D, e.Dp, e.SW2, 128};
        const bool ring8 = 2 * e.ftzero + 63 <= 255;
        const bool nopad = e.Dp == e.D && (e.Dp / 2) % kCfDCP == 0;
        const size_t smem = cost_fast_smem(e.SW2, ring8);
        const int TX = kCfTXH - 2 * e.SW2;
        dim3 grid((e.W1 + TX - 1) / TX, (e.Dp / 2 + kCfDCP - 1) / kCfDCP, (H + fg.RS - 1) / fg.RS);
        void (*kern)(const Feat*, const Feat*, uint16_t*, CostFastGeom);
        if (e.SW2 == 4) kern = ring8 ? (nopad ? k_cost_fast<4, true, true> : k_cost_fast<4, true, false>) : k_cost_fast<4, false, false>;
        else if (e.SW2 == 2) kern = ring8 ? (nopad ? k_cost_fast<2, true, true> : k_cost_fast<2, true, false>) : k_cost_fast<2, false, false>;
        else kern = ring8 ? k_cost_fast<0, true, false> : k_cost_fast<0, false, false>;
        CUDA_TRY_(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kern<<<grid, 256, smem, st>>>(fl, fr, C, fg);
        LAUNCH_CHECK_();
    } else {
        CostGeom cg;
        cg.W = W; cg.H = H; cg.W1 = e.W1; cg.minX1 = e.minX1; cg.minD = e.minD; cg.D = e.D; cg.Dp = e.Dp; cg.SW2 = e.SW2;
        int TX = 32, DCP = 32;
        while (DCP * 2 > e.Dp && DCP > 1) DCP /= 2;
        const size_t limit = 200 * 1024;
        while (cost_smem_bytes(TX, DCP, e.SW2) > limit && TX > 4) TX /= 2;
        while (cost_smem_bytes(TX, DCP, e.SW2) > limit && DCP > 4) DCP /= 2;
        while (cost_smem_bytes(TX, DCP, e.SW2) > limit && TX > 1) TX /= 2;
        if (cost_smem_bytes(TX, DCP, e.SW2) > limit) { *errmsg = "blockSize too large for the cost kernel"; return cudaErrorInvalidValue; }
        cg.TX = TX; cg.DCP = DCP; cg.RS = 64;
        const size_t smem = cost_smem_bytes(TX, DCP, e.SW2);
        CUDA_TRY_(cudaFuncSetAttribute(k_cost_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, int(limit)));
        dim3 grid((e.W1 + TX - 1) / TX, (e.Dp / 2 + DCP - 1) / DCP, (H + cg.RS - 1) / cg.RS);
        k_cost_generic<<<grid, 256, smem, st>>>(fl, fr, C, cg);
        LAUNCH_CHECK_();
    }
    return cudaSuccess;
}

}  // namespace b200sgm
