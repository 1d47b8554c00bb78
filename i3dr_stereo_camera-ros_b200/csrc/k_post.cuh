// Post filters (A.8): 3x3 median on int16 with replicate border, and the speckle filter as an exact
// connected-components labelling (lock-free union-find), followed by the float conversions and the
// reprojection either side of the matcher (rows a10, a11 and R of SURVEY.md section 8).
#pragma once
#include "sgm_types.h"

namespace b200sgm {

static __global__ void k_fill16(int16_t* __restrict__ p, int n, int16_t v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__device__ __forceinline__ void cswap(int& a, int& b) { int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

// cv::medianBlur(disp, 3) on CV_16S (replicate border) -- always applied by StereoSGBM::compute.
static __global__ void k_median3(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int W, int H)
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
__device__ __forceinline__ int uf_find(int* L, int i)
{
    int p = __ldcg(L + i);  // L2 loads: other SMs re-parent nodes concurrently with atomics
    while (p != i) {
        // path halving: a non-root never becomes a root again and its parent only ever moves to an ancestor of the
        // same set (smaller index), so this plain store cannot disconnect anything it races with
        const int gp = __ldcg(L + p);
        if (gp != p) __stcg(L + i, gp);
        i = p; p = gp;
    }
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

// Run-based labelling: every maximal horizontal run of connected pixels is one union-find node (its first
// pixel).  k_speckle_runs labels the runs of one image row per CTA and records their lengths, k_speckle_vmerge
// unites vertically touching runs (one union per touching run pair segment), k_speckle_size adds each run's
// length to its root, k_speckle_apply overwrites small components.
//   label[p]  : pixel index of the run start of p (or -1 for invalid pixels)
//   parent[p] : union-find parent, meaningful at run starts only
//   runlen[p] : run length, at run starts only;  csize[p]: component size, at roots only
static __global__ void __launch_bounds__(256) k_speckle_runs(const int16_t* __restrict__ img, int* __restrict__ label,
                                                      int* __restrict__ parent, int* __restrict__ runlen,
                                                      int* __restrict__ csize, int W, int newVal, int maxDiff)
{
    __shared__ int warp_last[8];
    const int y = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int16_t* row = img + size_t(y) * W;
    const int base = y * W;
    const int ppt = (W + 255) / 256;
    const int xa = t * ppt, xb = min(xa + ppt, W);
    // pass 1: last run start inside my segment
    int last = -1;
    const int prev = (xa > 0 && xa < W) ? int(row[xa - 1]) : newVal;
    int pv = prev;
    for (int x = xa; x < xb; x++) {
        const int v = row[x];
        const bool valid = v != newVal;
        const bool start = valid && !(pv != newVal && abs(v - pv) <= maxDiff);
        if (start) last = x;
        pv = v;
    }
    // block-wide exclusive max-scan of `last`
    int inc = last;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(kFullMask, inc, o); if (lane >= o) inc = max(inc, n); }
    if (lane == 31) warp_last[wid] = inc;
    __syncthreads();
    int before = __shfl_up_sync(kFullMask, inc, 1);
    if (lane == 0) before = -1;
    for (int w2 = 0; w2 < wid; w2++) before = max(before, warp_last[w2]);
    // pass 2: labels, run lengths (written by the pixel that ends the run)
    int cur = before;
    pv = prev;
    for (int x = xa; x < xb; x++) {
        const int v = row[x];
        const bool valid = v != newVal;
        const bool start = valid && !(pv != newVal && abs(v - pv) <= maxDiff);
        if (start) cur = x;
        const int p = base + x;
        if (valid) {
            label[p] = base + cur;
            if (start) { parent[p] = p; csize[p] = 0; }   // component sizes are accumulated at roots, and roots are run starts
            const int nv = x + 1 < W ? int(row[x + 1]) : newVal;
            const bool ends = !(nv != newVal && abs(nv - v) <= maxDiff);
            if (ends) runlen[base + cur] = x - cur + 1;
        } else {
            label[p] = -1;
        }
        pv = v;
    }
}

static __global__ void k_speckle_vmerge(const int16_t* __restrict__ img, const int* __restrict__ label, int* __restrict__ parent,
                                 int W, int H, int newVal, int maxDiff)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y + 1 >= H) return;
    const int p = y * W + x;
    const int v = img[p], u = img[p + W];
    if (v == newVal || u == newVal || abs(u - v) > maxDiff) return;
    if (x > 0) {   // the pair one pixel to the left already unites the same two runs
        const int vl = img[p - 1], ul = img[p + W - 1];
        if (vl != newVal && ul != newVal && abs(vl - v) <= maxDiff && abs(ul - u) <= maxDiff && abs(ul - vl) <= maxDiff) return;
    }
    uf_union(parent, label[p], label[p + W]);
}

// csize only has to answer "size <= maxSize": once a root's count is past the threshold further runs skip the atomic,
// which keeps the thousands of runs of a large surface from serialising on one address.
static __global__ void k_speckle_size(const int* __restrict__ label, int* __restrict__ parent, const int* __restrict__ runlen,
                               int* __restrict__ csize, int n, int maxSize)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || label[p] != p) return;   // run starts only
    const int r = uf_find(parent, p);
    parent[p] = r;                          // flatten (roots keep parent[r] == r)
    if (__ldcg(csize + r) <= maxSize) atomicAdd(&csize[r], runlen[p]);
}

static __global__ void k_speckle_apply(int16_t* __restrict__ img, const int* __restrict__ label, const int* __restrict__ parent,
                                const int* __restrict__ csize, int n, int newVal, int maxSize)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = label[p];
    if (s < 0) return;
    // k_speckle_size left every run start one or two hops from its root (its own flattening store can be overtaken by a
    // concurrent path-halving store, which still points at an ancestor)
    int r = parent[s], q = parent[r];
    while (q != r) { r = q; q = parent[r]; }
    if (csize[r] <= maxSize) img[p] = int16_t(newVal);
}

// ---- a10: CV_16S -> CV_32F, value unchanged (matcherOpenCVSGBM.cpp:34, abstractStereoMatcher.cpp:49)
static __global__ void k_to_f32(const int16_t* __restrict__ src, float* __restrict__ dst, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = float(src[i]);
}

// ---- a11 + R fused: /16, depth-window -> 10000 (generate_disparity.cpp:436-452), then the float32
// reprojection of disparity_to_depth.cpp:150-205 with separate multiply and add (no FMA contraction),
// and an order-preserving compaction of the kept points (row-major scan order).
struct ReprojGeom {
    int W, H;
    float q03, q13, wz, q32, q33, min_disp, max_disp;
    double depth_min, depth_max;     // the reference compares the float Z with its double parameters (disparity_to_depth.cpp:175)
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
    return w > 0.0f && Z > 0.0f && double(Z) <= g.depth_max && double(Z) >= g.depth_min;
}

// pass 1: dmat + depth + per-block kept count
static __global__ void __launch_bounds__(256) k_reproject_count(const int16_t* __restrict__ disp, ReprojGeom g,
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
static __global__ void __launch_bounds__(1024) k_scan_blocks(uint32_t* __restrict__ block_count, int nblocks, uint32_t* __restrict__ total)
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
// color: MONO8 (channels 1) or BGR8 (channels 3) rows of `color_pitch` bytes (disparity_to_depth.cpp:111-125, :176-188)
static __global__ void __launch_bounds__(256) k_reproject_write(const int16_t* __restrict__ disp, const uint8_t* __restrict__ color,
                                                         size_t color_pitch, int channels, ReprojGeom g,
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
        uint32_t bb = 0, gg = 0, rr = 0;
        if (color && channels == 1) bb = gg = rr = color[size_t(i) * color_pitch + j];
        else if (color && channels == 3) {
            const uint8_t* px = color + size_t(i) * color_pitch + 3 * j;
            bb = px[0]; gg = px[1]; rr = px[2];
        }
        pts[off] = make_float4(X, Y, Z, __uint_as_float((rr << 16) | (gg << 8) | bb));
    }
}

}  // namespace b200sgm
