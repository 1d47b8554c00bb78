// Post filters (A.8): 3x3 median on int16 with replicate border, and the speckle filter as an exact
// connected-components labelling (lock-free union-find), followed by the float conversions and the
// reprojection either side of the matcher (rows a10, a11 and R of SURVEY.md section 8).
#pragma once
#include "sgm_types.h"

namespace b200sgm {

__global__ void k_fill16(int16_t* __restrict__ p, int n, int16_t v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__device__ __forceinline__ void cswap(int& a, int& b) { int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

// cv::medianBlur(disp, 3) on CV_16S (replicate border) -- always applied by StereoSGBM::compute.
__global__ void k_median3(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int W, int H)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const int16_t* r0 = src + size_t(max(y - 1, 0)) * W;
    const int16_t* r1 = src + size_t(y) * W;
    const int16_t* r2 = src + size_t(min(y + 1, H - 1)) * W;
    int p0 = r0[xm], p1 = r0[x], p2 = r0[xp], p3 = r1[xm], p4 = r1[x], p5 = r1[xp], p6 = r2[xm], p7 = r2[x], p8 = r2[xp];
    // 19-exchange median-of-9 network
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5); cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
    cswap(p4, p2);
    dst[size_t(y) * W + x] = int16_t(p4);
}

// ---- speckle filter: cv::filterSpeckles(disp, INVALID, maxSize, 16*range) --------------------------
// Components are 4-connected sets of pixels != newVal where neighbours join iff |a-b| <= maxDiff;
// components of size <= maxSize are overwritten with newVal (order independent).
__device__ __forceinline__ int uf_find(const int* L, int i)
{
    int p = __ldcg(L + i);  // L2 loads: other SMs re-parent nodes concurrently with atomics
    while (p != i) { i = p; p = __ldcg(L + i); }
    return i;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b)
{
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // a > b: hang the larger root under the smaller
        int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;                                   // somebody re-parented a meanwhile: keep merging
    }
}

__global__ void k_speckle_init(const int16_t* __restrict__ img, int* __restrict__ label, int* __restrict__ size,
                               int n, int newVal)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    label[i] = img[i] == newVal ? -1 : i;
    size[i] = 0;
}

__global__ void k_speckle_merge(const int16_t* __restrict__ img, int* __restrict__ label, int W, int H,
                                int newVal, int maxDiff)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int i = y * W + x;
    int v = img[i];
    if (v == newVal) return;
    if (x + 1 < W) { int u = img[i + 1]; if (u != newVal && abs(u - v) <= maxDiff) uf_union(label, i, i + 1); }
    if (y + 1 < H) { int u = img[i + W]; if (u != newVal && abs(u - v) <= maxDiff) uf_union(label, i, i + W); }
}

__global__ void k_speckle_count(int* __restrict__ label, int* __restrict__ size, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (label[i] < 0) return;
    int r = uf_find(label, i);
    label[i] = r;  // flatten (roots keep label[r] == r, so concurrent finds stay correct)
    atomicAdd(&size[r], 1);
}

__global__ void k_speckle_apply(int16_t* __restrict__ img, const int* __restrict__ label, const int* __restrict__ size,
                                int n, int newVal, int maxSize)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = label[i];
    if (r >= 0 && size[r] <= maxSize) img[i] = int16_t(newVal);
}

// ---- a10: CV_16S -> CV_32F, value unchanged (matcherOpenCVSGBM.cpp:34, abstractStereoMatcher.cpp:49)
__global__ void k_to_f32(const int16_t* __restrict__ src, float* __restrict__ dst, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = float(src[i]);
}

// ---- a11 + R fused: /16, depth-window -> 10000 (generate_disparity.cpp:436-452), then the float32
// reprojection of disparity_to_depth.cpp:150-205 with separate multiply and add (no FMA contraction),
// and an order-preserving compaction of the kept points (row-major scan order).
struct ReprojGeom {
    int W, H;
    float q03, q13, wz, q32, q33, depth_min, depth_max, min_disp, max_disp;
};

__device__ __forceinline__ bool reproject_pixel(const ReprojGeom& g, int i, int j, int d16, float& dm, float& X, float& Y, float& Z)
{
    dm = __fmul_rn(float(d16), 0.0625f);
    if (dm < g.min_disp) dm = 10000.0f;
    if (dm > g.max_disp) dm = 10000.0f;
    X = Y = Z = 0.0f;
    if (dm == 0.0f || dm == 10000.0f) return false;
    float w = __fadd_rn(__fmul_rn(dm, g.q32), g.q33);
    X = __fdiv_rn(__fadd_rn(float(j), g.q03), w);
    Y = __fdiv_rn(__fadd_rn(float(i), g.q13), w);
    Z = __fdiv_rn(g.wz, w);
    return w > 0.0f && Z > 0.0f && Z <= g.depth_max && Z >= g.depth_min;
}

// pass 1: dmat + depth + per-block kept count
__global__ void __launch_bounds__(256) k_reproject_count(const int16_t* __restrict__ disp, ReprojGeom g,
                                                         float* __restrict__ dmat, float* __restrict__ depth,
                                                         uint32_t* __restrict__ block_count)
{
    const int n = g.W * g.H;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (p < n) {
        int i = p / g.W, j = p - i * g.W;
        float dm, X, Y, Z;
        keep = reproject_pixel(g, i, j, disp[p], dm, X, Y, Z);
        if (dmat) dmat[p] = dm;
        if (depth) depth[p] = keep ? Z : 0.0f;
    }
    int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_count[blockIdx.x] = uint32_t(c);
}

// pass 2: exclusive scan of the block counts (single CTA, sequential chunks of 1024)
__global__ void __launch_bounds__(1024) k_scan_blocks(uint32_t* __restrict__ block_count, int nblocks, uint32_t* __restrict__ total)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_count[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(kFullMask, s, o); if (lane >= o) s += t; }
        if (lane == 31) warp_sum[wid] = s;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sum[lane], ws = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(kFullMask, ws, o); if (lane >= o) ws += t; }
            warp_sum[lane] = ws - w;  // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_sum[wid] + s - v;
        if (i < nblocks) block_count[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// pass 3: write the kept points at block offset + rank within the block
__global__ void __launch_bounds__(256) k_reproject_write(const int16_t* __restrict__ disp, const uint8_t* __restrict__ gray,
                                                         size_t gray_pitch, ReprojGeom g,
                                                         const uint32_t* __restrict__ block_offset, float4* __restrict__ pts)
{
    __shared__ uint32_t warp_cnt[8];
    const int n = g.W * g.H;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    float dm, X = 0, Y = 0, Z = 0;
    int i = 0, j = 0;
    if (p < n) {
        i = p / g.W; j = p - i * g.W;
        keep = reproject_pixel(g, i, j, disp[p], dm, X, Y, Z);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t ballot = __ballot_sync(kFullMask, keep);
    if (lane == 0) warp_cnt[wid] = __popc(ballot);
    __syncthreads();
    uint32_t off = block_offset[blockIdx.x];
    for (int w = 0; w < wid; w++) off += warp_cnt[w];
    off += __popc(ballot & ((1u << lane) - 1u));
    if (keep) {
        uint32_t gv = gray ? gray[size_t(i) * gray_pitch + j] : 0u;
        pts[off] = make_float4(X, Y, Z, __uint_as_float((gv << 16) | (gv << 8) | gv));
    }
}

}  // namespace b200sgm
