// Host side of the B200 SGBM engine and its C ABI (include/b200sgm.h).
// Orchestrates the CUDA kernels that replace cv::StereoSGBM::compute as called by the reference's
// MatcherOpenCVSGBM::forwardMatch (/root/reference/src/stereoMatcher/matcherOpenCVSGBM.cpp:17-44).
// No CPU fallback exists: every result is produced by the kernels below.
#include "engine_internal.h"
#include "k_fused.cuh"   // geometry helpers only; the kernels are instantiated in agg_n*.cu
#include "k_post.cuh"

CoopGate& coop_gate(int device)
{
    static CoopGate gates[64];
    return gates[device & 63];
}

namespace {

int nreg_for(int D)
{
    int n = 1;
    while (64 * n < D) n *= 2;
    return n;
}

// A.1 parameter normalisation + geometry.  Returns 0 or B200SGM_EINVAL with h->err set.
int make_eff(b200sgm_engine* h, int W, int H, Eff& e)
{
    if (h->bm_only) return fail(h, B200SGM_ESTATE, "this engine was created with b200sgm_create_bm: block matcher and rectification only");
    b200sgm_params p;
    {
        std::lock_guard<std::mutex> lk(h->mu);     // set_params may run on another thread
        if (!h->have_params) return fail(h, B200SGM_ESTATE, "b200sgm_set_params has not been called");
        p = h->raw;
    }
    if (p.numDisparities <= 0) return fail(h, B200SGM_EINVAL, "numDisparities must be > 0");
    if (p.numDisparities > h->maxD) return fail(h, B200SGM_ESIZE, "numDisparities exceeds the engine's max_disparities");
    if (W <= 0 || H <= 0) return fail(h, B200SGM_EINVAL, "empty image");
    if (W > h->maxW || H > h->maxH) return fail(h, B200SGM_ESIZE, "image exceeds the engine's max size");
    if (W > 65535) return fail(h, B200SGM_EINVAL, "width > 65535 unsupported");
    e.W = W; e.H = H;
    e.minD = p.minDisparity; e.D = p.numDisparities;
    e.nreg = nreg_for(e.D);
    if (e.nreg > 32) return fail(h, B200SGM_EINVAL, "numDisparities > 2048 unsupported");
    e.Dp = (e.D + 2 * e.nreg - 1) / (2 * e.nreg) * (2 * e.nreg);
    e.SW2 = (p.blockSize > 0 ? p.blockSize : 5) / 2;
    e.P1 = p.P1 > 0 ? p.P1 : 2;
    e.P2 = std::max(p.P2 > 0 ? p.P2 : 5, e.P1 + 1);
    e.d12 = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
    e.uniq = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
    e.ftzero = std::max(p.preFilterCap, 15) | 1;
    e.speckleWin = p.speckleWindowSize; e.speckleRange = p.speckleRange;
    e.mode = p.mode;
    e.INVALID = (e.minD - 1) * 16;
    const int maxD = e.minD + e.D;
    e.minX1 = std::max(maxD, 0);
    e.W1 = W + std::min(e.minD, 0) - e.minX1;
    if (p.mode != B200SGM_MODE_SGBM && p.mode != B200SGM_MODE_HH) return fail(h, B200SGM_EINVAL, "mode must be MODE_SGBM (0) or MODE_HH (1)");
    if (e.ftzero > 127) return fail(h, B200SGM_EINVAL, "preFilterCap > 127 unsupported (Sobel channel must fit uint8)");
    // packed uint16 arithmetic contract: "infinity" + penalties must not wrap, output must fit int16
    if (e.P1 + e.P2 > 32000) return fail(h, B200SGM_EINVAL, "P1 + P2 > 32000 unsupported");
    if (maxD * 16 >= 32768 || e.INVALID < -32768) return fail(h, B200SGM_EINVAL, "disparity range does not fit CV_16S x16");
    if (e.SW2 > 127) return fail(h, B200SGM_EINVAL, "blockSize > 255 unsupported");
    return B200SGM_OK;
}

int launch_aggregation(b200sgm_engine* h, Lane& ln, const Eff& e, cudaStream_t st, int part = 0)
{
    switch (e.nreg) {
        case 1: return launch_agg_n<1>(h, ln, e, st, part);
        case 2: return launch_agg_n<2>(h, ln, e, st, part);
        case 4: return launch_agg_n<4>(h, ln, e, st, part);
        case 8: return launch_agg_n<8>(h, ln, e, st, part);
        case 16: return launch_agg_n<16>(h, ln, e, st, part);
        case 32: return launch_agg_n<32>(h, ln, e, st, part);
    }
    return fail(h, B200SGM_EINVAL, "bad nreg");
}

void prof_harvest(Lane& ln, cudaEvent_t ref = nullptr)
{
    // all recorded sets are complete once the stream is synchronised
    int first = (ln.prof_head - ln.prof_count + kProfRing) % kProfRing;
    for (int f = 0; f < ln.prof_count; f++) {
        size_t base = size_t((first + f) % kProfRing) * (kStages + 1);
        for (int s = 0; s < kStages; s++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ln.prof_events[base + s], ln.prof_events[base + s + 1]) == cudaSuccess) ln.stage_ms[s] += ms;
        }
        if (ref && ln.timeline.size() < size_t(4096) * (kStages + 1)) {
            for (int s = 0; s <= kStages; s++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, ref, ln.prof_events[base + s]);
                ln.timeline.push_back(ms);
            }
        }
        ln.stage_frames++;
    }
    ln.prof_count = 0;
}

// Enqueues the whole matcher on `st`: ln.left/right (device, pitch W) -> ln.disp_out (device, pitch W).
int run_pipeline(b200sgm_engine* h, Lane& ln, const Eff& e, const uint8_t* dL, size_t lp, const uint8_t* dR, size_t rp, cudaStream_t st)
{
    const int W = e.W, H = e.H;
    const int npix = W * H;
    if (h->profile) {
        if (ln.prof_events.empty()) {
            ln.prof_events.resize(size_t(kProfRing) * (kStages + 1));
            for (auto& ev : ln.prof_events) CUDA_TRY(h, cudaEventCreate(&ev));
        }
        if (ln.prof_count == kProfRing) { CUDA_TRY(h, cudaStreamSynchronize(st)); prof_harvest(ln, h->prof_ref); }
    }
    CUDA_TRY(h, cudaMemsetAsync(ln.d_err, 0, kStatusWords * sizeof(int), st));   // a failed frame must not poison the next one
    ln.last_P2 = e.P2;
    prof_mark(h, ln, 0, st);
    {
        launch_prefilter(dL, lp, dR, rp, W, H, e.ftzero, ln.feat_l, ln.feat_r, st);
        LAUNCH_CHECK(h);
    }
    CUDA_TRY(h, cudaMemsetAsync(ln.disp2key, 0xFF, size_t(npix) * 4, st));
    launch_fill16(ln.disp_wta, npix, int16_t(e.INVALID), st);
    LAUNCH_CHECK(h);
    prof_mark(h, ln, 1, st);
    if (e.W1 > 0) {
        const char* msg = nullptr;
        int nl = 0;
        const cudaError_t ce = launch_cost(ln.feat_l, ln.feat_r, ln.C, e, h->path == 1, h->num_sms, st, &nl, &msg);
        h->launches += nl;
        if (msg) return fail(h, B200SGM_EINVAL, msg);
        if (ce != cudaSuccess) return fail(h, B200SGM_ECUDA, std::string("cost kernel launch: ") + cudaGetErrorString(ce));
    }
    prof_mark(h, ln, 2, st);
    if (e.W1 > 0) {
        int rc = launch_aggregation(h, ln, e, st);
        if (rc) return rc;
    } else {
        prof_mark(h, ln, 3, st);
    }
    prof_mark(h, ln, 4, st);
    if (e.W1 > 0) {
        launch_lrcheck(ln.disp_wta, ln.disp2key, e, st);
        LAUNCH_CHECK(h);
    }
    prof_mark(h, ln, 5, st);
    launch_median3(ln.disp_wta, ln.disp_med, W, H, st);
    LAUNCH_CHECK(h);
    prof_mark(h, ln, 6, st);
    CUDA_TRY(h, cudaMemcpyAsync(ln.disp_out, ln.disp_med, size_t(npix) * 2, cudaMemcpyDeviceToDevice, st));
    if (e.speckleWin > 0) {
        int nl = 0;
        const cudaError_t ce = launch_speckle(ln.disp_out, ln.label, ln.parent, ln.runlen, ln.csize, W, H, e.INVALID, e.speckleWin, 16 * e.speckleRange, st, &nl);
        h->launches += nl;
        if (ce != cudaSuccess) return fail(h, B200SGM_ECUDA, std::string("speckle filter launch: ") + cudaGetErrorString(ce));
    }
    prof_mark(h, ln, 7, st);
    if (h->profile && !ln.prof_events.empty()) { ln.prof_head = (ln.prof_head + 1) % kProfRing; ln.prof_count++; }
    CUDA_TRY(h, cudaMemcpyAsync(ln.h_err, ln.d_err, kStatusWords * sizeof(int), cudaMemcpyDeviceToHost, st));
    return B200SGM_OK;
}

// Status of the last frame of `ln` once its stream has been synchronised: error, warning (> 0) or OK.
int frame_status(b200sgm_engine* h, Lane& ln)
{
    if (ln.h_err[0]) return fail(h, B200SGM_ECUDA, "fused aggregation kernel: an inter-strip wait timed out; the disparity of this frame is invalid");
    if (ln.h_err[3] + ln.last_P2 > kMaxCost) {
        h->err = "cost volume reached " + std::to_string(ln.h_err[3]) + " (+P2 > 32767): outside cv::StereoSGBM's int16 contract, results may differ from OpenCV";
        return B200SGM_WARN_COST_RANGE;
    }
    return B200SGM_OK;
}

}  // namespace
namespace {
int lane_check(b200sgm_engine* h, int lane)
{
    if (lane < 0 || lane >= int(h->lanes.size())) return fail(h, B200SGM_EINVAL, "lane out of range");
    return B200SGM_OK;
}

void free_lane(Lane& ln)
{
    cudaFree(ln.left); cudaFree(ln.right); cudaFree(ln.feat_l); cudaFree(ln.feat_r); cudaFree(ln.C); cudaFree(ln.S); cudaFree(ln.ckpt);
    cudaFree(ln.disp2key); cudaFree(ln.disp_wta); cudaFree(ln.disp_med); cudaFree(ln.disp_out); cudaFree(ln.label);
    cudaFree(ln.csize); cudaFree(ln.parent); cudaFree(ln.runlen); cudaFree(ln.f32a); cudaFree(ln.f32b); cudaFree(ln.points); cudaFree(ln.block_count); cudaFree(ln.total);
    if (ln.h_total) cudaFreeHost(ln.h_total);
    cudaFree(ln.xbuf); cudaFree(ln.d_err); cudaFree(ln.color);
    if (ln.h_err) cudaFreeHost(ln.h_err);
    for (auto ev : ln.prof_events) cudaEventDestroy(ev);
    if (ln.done) cudaEventDestroy(ln.done);
    if (ln.stream) cudaStreamDestroy(ln.stream);
    ln = Lane();
}

}  // namespace

extern "C" {

const char* b200sgm_version(void) { return "b200sgm 0.1 (sm_100a)"; }

static int create_engine(int device, int max_width, int max_height, int max_disparities, int lanes, bool bm_only, b200sgm_handle* out)
{
    if (!out) return B200SGM_EINVAL;
    *out = nullptr;
    if (max_width <= 0 || max_height <= 0 || max_disparities <= 0 || lanes <= 0 || lanes > 64) return B200SGM_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return B200SGM_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return B200SGM_ECUDA;
    b200sgm_engine* h = new b200sgm_engine();
    h->device = device; h->maxW = max_width; h->maxH = max_height; h->maxD = max_disparities; h->bm_only = bm_only;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, device);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&h->max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    const size_t npix = size_t(max_width) * max_height;
    const int nreg = nreg_for(max_disparities);
    const size_t Dp = size_t((max_disparities + 2 * nreg - 1) / (2 * nreg) * (2 * nreg));
    const size_t vol = npix * Dp * sizeof(uint16_t);
    h->lanes.resize(lanes);
    bool ok = true;
    for (Lane& ln : h->lanes) {
        ok = ok && cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.left, npix) == cudaSuccess && cudaMalloc(&ln.right, npix) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.feat_l, npix * sizeof(Feat)) == cudaSuccess && cudaMalloc(&ln.feat_r, npix * sizeof(Feat)) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.C, vol) == cudaSuccess;
        if (!bm_only) {
            ok = ok && cudaMalloc(&ln.S, vol) == cudaSuccess;
            ok = ok && cudaMalloc(&ln.ckpt, horiz_ckpt_elems(max_width, max_height, int(Dp)) * sizeof(uint16_t)) == cudaSuccess;
            ok = ok && cudaMalloc(&ln.xbuf, size_t(2) * kMaxStrips * kHaloGen * (Dp / 2) * sizeof(uint2)) == cudaSuccess;
        }
        ok = ok && cudaMalloc(&ln.disp2key, npix * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.disp_wta, npix * 2 + 16) == cudaSuccess && cudaMalloc(&ln.disp_med, npix * 2) == cudaSuccess && cudaMalloc(&ln.disp_out, npix * 2) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.label, npix * 4) == cudaSuccess && cudaMalloc(&ln.csize, npix * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.parent, npix * 4) == cudaSuccess && cudaMalloc(&ln.runlen, npix * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.f32a, npix * 4) == cudaSuccess && cudaMalloc(&ln.f32b, npix * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.points, npix * sizeof(float4)) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.block_count, ((npix + 255) / 256 + 1) * 4) == cudaSuccess && cudaMalloc(&ln.total, 4) == cudaSuccess;
        ok = ok && cudaMallocHost(&ln.h_total, 4) == cudaSuccess;
        ok = ok && cudaMalloc(&ln.d_err, kStatusWords * sizeof(int)) == cudaSuccess && cudaMemset(ln.d_err, 0, kStatusWords * sizeof(int)) == cudaSuccess;
        ok = ok && cudaMallocHost(&ln.h_err, kStatusWords * sizeof(int)) == cudaSuccess;
        if (ok) memset(ln.h_err, 0, kStatusWords * sizeof(int));
        if (!ok) break;
    }
    if (!ok) {
        fprintf(stderr, "b200sgm_create: CUDA allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        for (Lane& ln : h->lanes) free_lane(ln);
        delete h;
        return B200SGM_ECUDA;
    }
    *out = h;
    return B200SGM_OK;
}

int b200sgm_create(int device, int max_width, int max_height, int max_disparities, int lanes, b200sgm_handle* out)
{
    return create_engine(device, max_width, max_height, max_disparities, lanes, false, out);
}

int b200sgm_create_bm(int device, int max_width, int max_height, int max_disparities, int lanes, b200sgm_handle* out)
{
    return create_engine(device, max_width, max_height, max_disparities, lanes, true, out);
}

int b200sgm_host_alloc(size_t bytes, void** ptr)
{
    if (!ptr || bytes == 0) return B200SGM_EINVAL;
    *ptr = nullptr;
    return cudaMallocHost(ptr, bytes) == cudaSuccess ? B200SGM_OK : B200SGM_ECUDA;
}

int b200sgm_host_free(void* ptr)
{
    if (!ptr) return B200SGM_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? B200SGM_OK : B200SGM_ECUDA;
}

int b200sgm_destroy(b200sgm_handle h)
{
    if (!h) return B200SGM_EINVAL;
    cudaSetDevice(h->device);
    for (Lane& ln : h->lanes) { if (ln.stream) cudaStreamSynchronize(ln.stream); free_lane(ln); }
    for (auto& r : h->rect) { cudaFree(r.ent); cudaFree(r.map1); cudaFree(r.map2); }
    cudaFree(h->d_wtab); cudaFree(h->rect_src); cudaFree(h->rect_dst);
    delete h;
    return B200SGM_OK;
}

int b200sgm_set_params(b200sgm_handle h, const b200sgm_params* p)
{
    if (!h || !p) return B200SGM_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    h->raw = *p;
    h->have_params = true;
    return B200SGM_OK;
}

int b200sgm_get_effective_params(b200sgm_handle h, b200sgm_params* out)
{
    if (!h || !out) return B200SGM_EINVAL;
    Eff e;
    int rc = make_eff(h, 1, 1, e);
    if (rc) return rc;
    out->minDisparity = e.minD; out->numDisparities = e.D; out->blockSize = 2 * e.SW2 + 1; out->P1 = e.P1; out->P2 = e.P2;
    out->disp12MaxDiff = e.d12; out->preFilterCap = e.ftzero; out->uniquenessRatio = e.uniq;
    out->speckleWindowSize = e.speckleWin; out->speckleRange = e.speckleRange; out->mode = e.mode;
    return B200SGM_OK;
}

int b200sgm_compute_device(b200sgm_handle h, int lane, const uint8_t* d_left, size_t left_stride, const uint8_t* d_right,
                           size_t right_stride, int width, int height, int16_t* d_disp, size_t disp_stride, void* cuda_stream)
{
    if (!h || !d_left || !d_right || !d_disp) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    Eff e;
    rc = make_eff(h, width, height, e);
    if (rc) return rc;
    if (left_stride < size_t(width) || right_stride < size_t(width) || disp_stride < size_t(width) * 2) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[lane];
    cudaStream_t st = cuda_stream ? cudaStream_t(cuda_stream) : ln.stream;
    rc = run_pipeline(h, ln, e, d_left, left_stride, d_right, right_stride, st);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpy2DAsync(d_disp, disp_stride, ln.disp_out, size_t(width) * 2, size_t(width) * 2, height, cudaMemcpyDeviceToDevice, st));
    return B200SGM_OK;
}

int b200sgm_enqueue(b200sgm_handle h, int lane, const uint8_t* left, size_t left_stride, const uint8_t* right, size_t right_stride,
                    int width, int height, int16_t* disp, size_t disp_stride)
{
    if (!h || !left || !right || !disp) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    Eff e;
    rc = make_eff(h, width, height, e);
    if (rc) return rc;
    if (left_stride < size_t(width) || right_stride < size_t(width) || disp_stride < size_t(width) * 2) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[lane];
    if (ln.busy) return fail(h, B200SGM_ESTATE, "lane is busy: call b200sgm_wait first");
    cudaStream_t st = ln.stream;
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.left, width, left, left_stride, width, height, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.right, width, right, right_stride, width, height, cudaMemcpyHostToDevice, st));
    rc = run_pipeline(h, ln, e, ln.left, width, ln.right, width, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }   // the copies from the caller's buffers are already enqueued: let them finish
    CUDA_TRY(h, cudaMemcpy2DAsync(disp, disp_stride, ln.disp_out, size_t(width) * 2, size_t(width) * 2, height, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaEventRecord(ln.done, st));
    ln.busy = true;
    return B200SGM_OK;
}

int b200sgm_wait(b200sgm_handle h, int lane)
{
    if (!h) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    Lane& ln = h->lanes[lane];
    if (!ln.busy) return fail(h, B200SGM_ESTATE, "lane is idle");
    CUDA_TRY(h, cudaSetDevice(h->device));
    ln.busy = false;
    CUDA_TRY(h, cudaEventSynchronize(ln.done));
    CUDA_TRY(h, cudaGetLastError());
    return frame_status(h, ln);
}

int b200sgm_lane_status(b200sgm_handle h, int lane)
{
    if (!h) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    return frame_status(h, h->lanes[lane]);
}

int b200sgm_compute(b200sgm_handle h, const uint8_t* left, size_t left_stride, const uint8_t* right, size_t right_stride,
                    int width, int height, int16_t* disp, size_t disp_stride)
{
    int rc = b200sgm_enqueue(h, 0, left, left_stride, right, right_stride, width, height, disp, disp_stride);
    if (rc) return rc;
    return b200sgm_wait(h, 0);
}

int b200sgm_compute_f32(b200sgm_handle h, const uint8_t* left, size_t left_stride, const uint8_t* right, size_t right_stride,
                        int width, int height, float* disp32, size_t disp_stride)
{
    if (!h || !left || !right || !disp32) return B200SGM_EINVAL;
    Eff e;
    int rc = make_eff(h, width, height, e);
    if (rc) return rc;
    if (left_stride < size_t(width) || right_stride < size_t(width) || disp_stride < size_t(width) * 4) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[0];
    if (ln.busy) return fail(h, B200SGM_ESTATE, "lane 0 is busy");
    cudaStream_t st = ln.stream;
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.left, width, left, left_stride, width, height, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.right, width, right, right_stride, width, height, cudaMemcpyHostToDevice, st));
    rc = run_pipeline(h, ln, e, ln.left, width, ln.right, width, st);
    if (rc) return rc;
    const int npix = width * height;
    k_to_f32<<<(npix + 255) / 256, 256, 0, st>>>(ln.disp_out, ln.f32a, npix);
    LAUNCH_CHECK(h);
    CUDA_TRY(h, cudaMemcpy2DAsync(disp32, disp_stride, ln.f32a, size_t(width) * 4, size_t(width) * 4, height, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return frame_status(h, ln);
}

int b200sgm_compute_xyz(b200sgm_handle h, const uint8_t* left, size_t left_stride, const uint8_t* right, size_t right_stride,
                        int width, int height, const b200sgm_reproject* rp, int16_t* disp, size_t disp_stride, float* dmat,
                        float* depth, b200sgm_point* points, uint32_t* count)
{
    if (!h || !left || !right || !rp) return B200SGM_EINVAL;
    if (points && !count) return B200SGM_EINVAL;
    Eff e;
    int rc = make_eff(h, width, height, e);
    if (rc) return rc;
    if (left_stride < size_t(width) || right_stride < size_t(width) || (disp && disp_stride < size_t(width) * 2)) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[0];
    if (ln.busy) return fail(h, B200SGM_ESTATE, "lane 0 is busy");
    cudaStream_t st = ln.stream;
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.left, width, left, left_stride, width, height, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.right, width, right, right_stride, width, height, cudaMemcpyHostToDevice, st));
    rc = run_pipeline(h, ln, e, ln.left, width, ln.right, width, st);
    if (rc) return rc;
    const int npix = width * height;
    const int nblk = (npix + 255) / 256;
    if (rp->color && (rp->color_channels != 1 && rp->color_channels != 3)) return fail(h, B200SGM_EINVAL, "color_channels must be 1 (MONO8) or 3 (BGR8)");
    if (rp->color && rp->color_stride < size_t(width) * rp->color_channels) return fail(h, B200SGM_EINVAL, "color stride smaller than a row");
    ReprojGeom g{width, height, rp->q03, rp->q13, rp->wz, rp->q32, rp->q33, rp->min_disparity, rp->max_disparity, rp->depth_min, rp->depth_max};
    const uint8_t* d_color = ln.left;
    size_t color_pitch = size_t(width);
    int channels = 1;
    if (rp->color && points) {
        channels = rp->color_channels;
        color_pitch = size_t(width) * channels;
        if (!ln.color) CUDA_TRY(h, cudaMalloc(&ln.color, size_t(h->maxW) * h->maxH * 3));
        CUDA_TRY(h, cudaMemcpy2DAsync(ln.color, color_pitch, rp->color, rp->color_stride, color_pitch, height, cudaMemcpyHostToDevice, st));
        d_color = ln.color;
    }
    k_reproject_count<<<nblk, 256, 0, st>>>(ln.disp_out, g, ln.f32a, ln.f32b, ln.block_count);
    LAUNCH_CHECK(h);
    k_scan_blocks<<<1, 1024, 0, st>>>(ln.block_count, nblk, ln.total);
    LAUNCH_CHECK(h);
    if (points) {
        k_reproject_write<<<nblk, 256, 0, st>>>(ln.disp_out, d_color, color_pitch, channels, g, ln.block_count, ln.points);
        LAUNCH_CHECK(h);
    }
    CUDA_TRY(h, cudaMemcpyAsync(ln.h_total, ln.total, 4, cudaMemcpyDeviceToHost, st));
    if (disp) CUDA_TRY(h, cudaMemcpy2DAsync(disp, disp_stride, ln.disp_out, size_t(width) * 2, size_t(width) * 2, height, cudaMemcpyDeviceToHost, st));
    if (dmat) CUDA_TRY(h, cudaMemcpyAsync(dmat, ln.f32a, size_t(npix) * 4, cudaMemcpyDeviceToHost, st));
    if (depth) CUDA_TRY(h, cudaMemcpyAsync(depth, ln.f32b, size_t(npix) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    const uint32_t n = *ln.h_total;
    if (count) *count = n;
    if (points && n) CUDA_TRY(h, cudaMemcpy(points, ln.points, size_t(n) * sizeof(b200sgm_point), cudaMemcpyDeviceToHost));
    return frame_status(h, ln);
}

void b200sgm_reproject_from_camera(b200sgm_reproject* rp, const double* Kl, const double* Pl, const double* Pr, double depth_min, double depth_max)
{
    if (!rp || !Kl || !Pl || !Pr) return;
    // calc_q (disparity_to_depth.cpp:62-85): note fx comes from K_l, the principal points from P_l / P_r
    const double cx = Pl[2], cxr = Pr[2], cy = Pl[6], fx = Kl[0];
    const double T = -Pr[3] / fx;
    rp->q03 = float(-cx); rp->q13 = float(-cy); rp->wz = float(fx); rp->q32 = float(1.0 / T); rp->q33 = float(-(cx - cxr) / T);
    // processDisparity (generate_disparity.cpp:441-450): DisparityImage.f / .T are float32 fields
    const float f32 = float(Pl[0]), T32 = float(-Pr[3] / Pr[0]);
    const float tf = T32 * f32;
    rp->min_disparity = float(double(tf) / depth_max);
    rp->max_disparity = float(double(tf) / depth_min);     // depth_min == 0 -> +inf, like the reference
    rp->depth_min = depth_min; rp->depth_max = depth_max;
    rp->color = nullptr; rp->color_stride = 0; rp->color_channels = 0;
}

const char* b200sgm_last_error(b200sgm_handle h) { return h ? h->err.c_str() : "null handle"; }

// ---- rectification (row N2): generate_disparity.cpp:370-386 ------------------------------------------------------------
int b200sgm_set_camera(b200sgm_handle h, int cam, const double* K, const double* D, int nD, const double* R, const double* P)
{
    if (!h) return B200SGM_EINVAL;
    if (cam < 0 || cam > 1 || !K || !P || nD < 0 || nD > 14 || (nD > 0 && !D)) return fail(h, B200SGM_EINVAL, "bad camera arguments");
    RectifyCam c;
    if (!make_rectify_cam(K, D, nD, R, P, c)) return fail(h, B200SGM_EINVAL, "singular P*R or tilted sensor model (tauX/tauY) unsupported");
    h->rect[cam].cam = c; h->rect[cam].have = true; h->rect[cam].dirty = true;
    return B200SGM_OK;
}

namespace {
int ensure_maps(b200sgm_engine* h, int cam, int w, int hgt, bool want_float, cudaStream_t st)
{
    if (cam < 0 || cam > 1) return fail(h, B200SGM_EINVAL, "camera index must be 0 or 1");
    auto& r = h->rect[cam];
    if (!r.have) return fail(h, B200SGM_ESTATE, "b200sgm_set_camera has not been called for this camera");
    if (w <= 0 || hgt <= 0) return fail(h, B200SGM_EINVAL, "empty image");
    if (w > h->maxW || hgt > h->maxH) return fail(h, B200SGM_ESIZE, "image exceeds the engine's max size");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!h->d_wtab) {
        std::vector<int16_t> tab(1024 * 16);
        build_cubic_table(tab.data());
        CUDA_TRY(h, cudaMalloc(&h->d_wtab, tab.size() * sizeof(int16_t)));
        CUDA_TRY(h, cudaMemcpy(h->d_wtab, tab.data(), tab.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
    }
    const size_t cap = size_t(h->maxW) * h->maxH;
    if (!r.ent) CUDA_TRY(h, cudaMalloc(&r.ent, cap * sizeof(RemapEntry)));
    if (want_float && !r.map1) {
        CUDA_TRY(h, cudaMalloc(&r.map1, cap * sizeof(float)));
        CUDA_TRY(h, cudaMalloc(&r.map2, cap * sizeof(float)));
        r.dirty = true;
    }
    if (r.dirty || r.W != w || r.H != hgt) {
        // The map buffers are shared by every lane and stream: nobody may still be reading the old maps while they are
        // rebuilt, and nobody may start remapping before the new ones are complete.  Rare (once per camera and size).
        CUDA_TRY(h, cudaDeviceSynchronize());
        launch_rectify_maps(r.cam, w, hgt, r.ent, r.map1, r.map2, st);
        LAUNCH_CHECK(h);
        CUDA_TRY(h, cudaStreamSynchronize(st));
        r.W = w; r.H = hgt; r.dirty = false;
    }
    return B200SGM_OK;
}
}  // namespace

int b200sgm_rectify_device(b200sgm_handle h, int lane, int cam, const uint8_t* d_src, size_t src_stride, int width, int height,
                           uint8_t* d_dst, size_t dst_stride, void* cuda_stream)
{
    if (!h) return B200SGM_EINVAL;
    if (int rc = lane_check(h, lane)) return rc;
    if (!d_src || !d_dst || src_stride < size_t(width) || dst_stride < size_t(width)) return fail(h, B200SGM_EINVAL, "bad image arguments");
    cudaStream_t st = cuda_stream ? cudaStream_t(cuda_stream) : h->lanes[lane].stream;
    if (int rc = ensure_maps(h, cam, width, height, false, st)) return rc;
    launch_remap_cubic(d_src, src_stride, width, height, h->rect[cam].ent, h->d_wtab, d_dst, dst_stride, width, height, st);
    LAUNCH_CHECK(h);
    return B200SGM_OK;
}

int b200sgm_rectify(b200sgm_handle h, int cam, const uint8_t* src, size_t src_stride, int width, int height, uint8_t* dst,
                    size_t dst_stride)
{
    if (!h) return B200SGM_EINVAL;
    if (!src || !dst || src_stride < size_t(width) || dst_stride < size_t(width)) return fail(h, B200SGM_EINVAL, "bad image arguments");
    if (width <= 0 || height <= 0) return fail(h, B200SGM_EINVAL, "empty image");
    if (width > h->maxW || height > h->maxH) return fail(h, B200SGM_ESIZE, "image exceeds the engine's max size");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t cap = size_t(h->maxW) * h->maxH;
    if (!h->rect_src) { CUDA_TRY(h, cudaMalloc(&h->rect_src, cap)); CUDA_TRY(h, cudaMalloc(&h->rect_dst, cap)); }
    cudaStream_t st = h->lanes[0].stream;
    CUDA_TRY(h, cudaMemcpy2DAsync(h->rect_src, width, src, src_stride, width, height, cudaMemcpyHostToDevice, st));
    if (int rc = b200sgm_rectify_device(h, 0, cam, h->rect_src, width, width, height, h->rect_dst, width, st)) return rc;
    CUDA_TRY(h, cudaMemcpy2DAsync(dst, dst_stride, h->rect_dst, width, width, height, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return B200SGM_OK;
}

int b200sgm_rectify_maps(b200sgm_handle h, int cam, int width, int height, float* map1, float* map2)
{
    if (!h || !map1 || !map2) return B200SGM_EINVAL;
    cudaStream_t st = h->lanes[0].stream;
    if (int rc = ensure_maps(h, cam, width, height, true, st)) return rc;
    const size_t bytes = size_t(width) * height * sizeof(float);
    CUDA_TRY(h, cudaMemcpyAsync(map1, h->rect[cam].map1, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaMemcpyAsync(map2, h->rect[cam].map2, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return B200SGM_OK;
}

// ---- StereoBM (row N4): matcherOpenCVBlock.cpp:13-20 ---------------------------------------------------------------------
namespace {
// cv::StereoBM's own argument checks (CV_Error -> forwardMatch() == -1 in the reference), then the kernels into ln.disp_out
int bm_run(b200sgm_engine* h, Lane& ln, const b200sgm_bm_params* bp, const uint8_t* dL, size_t lp, const uint8_t* dR, size_t rp, int width,
           int height, cudaStream_t st)
{
    if (bp->numDisparities <= 0 || bp->numDisparities % 16 != 0) return fail(h, B200SGM_EINVAL, "numDisparities must be positive and divisible by 16");
    if (bp->numDisparities > h->maxD) return fail(h, B200SGM_ESIZE, "numDisparities exceeds the engine's max_disparities");
    if (width <= 0 || height <= 0) return fail(h, B200SGM_EINVAL, "empty image");
    if (width > h->maxW || height > h->maxH) return fail(h, B200SGM_ESIZE, "image exceeds the engine's max size");
    if (bp->blockSize < 5 || bp->blockSize > 255 || bp->blockSize % 2 == 0 || bp->blockSize >= std::min(width, height))
        return fail(h, B200SGM_EINVAL, "blockSize must be odd, within 5..255 and smaller than the image width and height");
    if (bp->preFilterCap < 1 || bp->preFilterCap > 63) return fail(h, B200SGM_EINVAL, "preFilterCap must be within 1..63");
    if (bp->preFilterSize != 0 && (bp->preFilterSize < 5 || bp->preFilterSize > 255 || bp->preFilterSize % 2 == 0))
        return fail(h, B200SGM_EINVAL, "preFilterSize must be odd and be within 5..255");
    if (bp->textureThreshold < 0 || bp->uniquenessRatio < 0) return fail(h, B200SGM_EINVAL, "textureThreshold and uniquenessRatio must be non-negative");
    if (bp->disp12MaxDiff >= 0) return fail(h, B200SGM_EINVAL, "disp12MaxDiff >= 0 is not supported by the block matcher (the reference never sets it)");
    if ((bp->minDisparity + bp->numDisparities) * 16 >= 32768 || (bp->minDisparity - 1) * 16 < -32768) return fail(h, B200SGM_EINVAL, "disparity range does not fit CV_16S x16");
    if (lp < size_t(width) || rp < size_t(width)) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    BmParams p{bp->minDisparity, bp->numDisparities, bp->blockSize, bp->preFilterCap, bp->textureThreshold, bp->uniquenessRatio,
               bp->speckleWindowSize, bp->speckleRange};
    int nl = 0;
    const cudaError_t ce = launch_bm(dL, lp, dR, rp, width, height, p, reinterpret_cast<uint8_t*>(ln.disp_wta), ln.C, ln.csize, ln.disp_out,
                                     ln.label, ln.parent, ln.runlen, ln.csize, h->num_sms, st, &nl);
    h->launches += nl;
    if (ce != cudaSuccess) return fail(h, B200SGM_ECUDA, std::string("block matcher launch: ") + cudaGetErrorString(ce));
    return B200SGM_OK;
}
}  // namespace

int b200sgm_bm_compute(b200sgm_handle h, const b200sgm_bm_params* bp, const uint8_t* left, size_t left_stride, const uint8_t* right,
                       size_t right_stride, int width, int height, int16_t* disp, size_t disp_stride)
{
    if (!h || !bp || !left || !right || !disp) return B200SGM_EINVAL;
    if (width > 0 && disp_stride < size_t(width) * 2) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    if (width <= 0 || height <= 0) return fail(h, B200SGM_EINVAL, "empty image");
    if (width > h->maxW || height > h->maxH) return fail(h, B200SGM_ESIZE, "image exceeds the engine's max size");
    if (left_stride < size_t(width) || right_stride < size_t(width)) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[0];
    if (ln.busy) return fail(h, B200SGM_ESTATE, "lane 0 is busy: call b200sgm_wait first");
    cudaStream_t st = ln.stream;
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.left, width, left, left_stride, width, height, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.right, width, right, right_stride, width, height, cudaMemcpyHostToDevice, st));
    if (int rc = bm_run(h, ln, bp, ln.left, width, ln.right, width, width, height, st)) return rc;
    CUDA_TRY(h, cudaMemcpy2DAsync(disp, disp_stride, ln.disp_out, size_t(width) * 2, size_t(width) * 2, height, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return B200SGM_OK;
}

int b200sgm_bm_compute_f32(b200sgm_handle h, const b200sgm_bm_params* bp, const uint8_t* left, size_t left_stride, const uint8_t* right,
                           size_t right_stride, int width, int height, float* disp32, size_t disp_stride)
{
    if (!h || !bp || !left || !right || !disp32) return B200SGM_EINVAL;
    if (width <= 0 || height <= 0) return fail(h, B200SGM_EINVAL, "empty image");
    if (width > h->maxW || height > h->maxH) return fail(h, B200SGM_ESIZE, "image exceeds the engine's max size");
    if (left_stride < size_t(width) || right_stride < size_t(width) || disp_stride < size_t(width) * 4) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[0];
    if (ln.busy) return fail(h, B200SGM_ESTATE, "lane 0 is busy: call b200sgm_wait first");
    cudaStream_t st = ln.stream;
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.left, width, left, left_stride, width, height, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpy2DAsync(ln.right, width, right, right_stride, width, height, cudaMemcpyHostToDevice, st));
    if (int rc = bm_run(h, ln, bp, ln.left, width, ln.right, width, width, height, st)) return rc;
    const int npix = width * height;
    k_to_f32<<<(npix + 255) / 256, 256, 0, st>>>(ln.disp_out, ln.f32a, npix);
    LAUNCH_CHECK(h);
    CUDA_TRY(h, cudaMemcpy2DAsync(disp32, disp_stride, ln.f32a, size_t(width) * 4, size_t(width) * 4, height, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return B200SGM_OK;
}

int b200sgm_bm_compute_device(b200sgm_handle h, int lane, const b200sgm_bm_params* bp, const uint8_t* d_left, size_t left_stride,
                              const uint8_t* d_right, size_t right_stride, int width, int height, int16_t* d_disp, size_t disp_stride,
                              void* cuda_stream)
{
    if (!h || !bp || !d_left || !d_right || !d_disp) return B200SGM_EINVAL;
    if (int rc = lane_check(h, lane)) return rc;
    if (width > 0 && disp_stride < size_t(width) * 2) return fail(h, B200SGM_EINVAL, "stride smaller than a row");
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[lane];
    cudaStream_t st = cuda_stream ? cudaStream_t(cuda_stream) : ln.stream;
    if (int rc = bm_run(h, ln, bp, d_left, left_stride, d_right, right_stride, width, height, st)) return rc;
    CUDA_TRY(h, cudaMemcpy2DAsync(d_disp, disp_stride, ln.disp_out, size_t(width) * 2, size_t(width) * 2, height, cudaMemcpyDeviceToDevice, st));
    return B200SGM_OK;
}

int b200sgm_launch_count(b200sgm_handle h, uint64_t* count)
{
    if (!h || !count) return B200SGM_EINVAL;
    *count = h->launches.load();
    return B200SGM_OK;
}

int b200sgm_lane_stream(b200sgm_handle h, int lane, void** cuda_stream)
{
    if (!h || !cuda_stream) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    *cuda_stream = (void*)h->lanes[lane].stream;
    return B200SGM_OK;
}

int b200sgm_profile(b200sgm_handle h, int enable)
{
    if (!h) return B200SGM_EINVAL;
    h->profile = enable != 0;
    if (h->profile) {
        cudaSetDevice(h->device);
        if (!h->prof_ref) cudaEventCreate(&h->prof_ref);
        cudaDeviceSynchronize();
        cudaEventRecord(h->prof_ref, h->lanes[0].stream);
        cudaStreamSynchronize(h->lanes[0].stream);
        for (Lane& ln : h->lanes) ln.timeline.clear();
    }
    return B200SGM_OK;
}

int b200sgm_stage_times(b200sgm_handle h, int lane, double* ms, int n, uint64_t* frames)
{
    if (!h || !ms || n < kStages) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[lane];
    CUDA_TRY(h, cudaStreamSynchronize(ln.stream));
    CUDA_TRY(h, cudaDeviceSynchronize());   // frames may have been enqueued on caller-provided streams
    prof_harvest(ln, h->prof_ref);
    for (int s = 0; s < kStages; s++) { ms[s] = ln.stage_ms[s]; ln.stage_ms[s] = 0; }
    if (frames) *frames = ln.stage_frames;
    ln.stage_frames = 0;
    return B200SGM_OK;
}

int b200sgm_stage_timeline(b200sgm_handle h, int lane, float* out, int max_floats, int* n_floats)
{
    if (!h || !n_floats) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    Lane& ln = h->lanes[lane];
    const int n = int(std::min<size_t>(ln.timeline.size(), size_t(std::max(max_floats, 0))));
    if (out) memcpy(out, ln.timeline.data(), size_t(n) * sizeof(float));
    *n_floats = int(ln.timeline.size());
    return B200SGM_OK;
}

int b200sgm_debug_read(b200sgm_handle h, int lane, const char* what, void* host, size_t bytes, int* dp)
{
    if (!h || !what || !host) return B200SGM_EINVAL;
    int rc = lane_check(h, lane);
    if (rc) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    Lane& ln = h->lanes[lane];
    CUDA_TRY(h, cudaStreamSynchronize(ln.stream));
    const void* src = nullptr;
    if (!strcmp(what, "C")) src = ln.C;
    else if (!strcmp(what, "S")) src = ln.S;
    else if (!strcmp(what, "wta")) src = ln.disp_wta;
    else if (!strcmp(what, "median")) src = ln.disp_med;
    else if (!strcmp(what, "stats")) src = ln.d_err;     // {error flag, late exchange records, poll iterations, max cost}
    else return fail(h, B200SGM_EINVAL, "unknown debug buffer");
    if (!src) return fail(h, B200SGM_ESTATE, "debug buffer not allocated");
    if (dp) {
        Eff e;
        if (make_eff(h, 1, 1, e) == 0) *dp = e.Dp;
    }
    CUDA_TRY(h, cudaMemcpy(host, src, bytes, cudaMemcpyDeviceToHost));
    if (src == ln.C || src == ln.S) {
        // device volumes use the paired layout (word w = cell w | cell Dh+w << 16): hand back natural order
        Eff e;
        if (make_eff(h, 1, 1, e) == 0 && e.Dp > 0) {
            const size_t Dp = size_t(e.Dp), Dh = Dp / 2;
            std::vector<uint16_t> tmp(Dp);
            uint16_t* v = static_cast<uint16_t*>(host);
            for (size_t px = 0; px + 1 <= bytes / (Dp * 2); px++, v += Dp) {
                for (size_t w = 0; w < Dh; w++) { tmp[w] = v[2 * w]; tmp[Dh + w] = v[2 * w + 1]; }
                memcpy(v, tmp.data(), Dp * 2);
            }
        }
    }
    return B200SGM_OK;
}

// Development probe: how long do stages of two different frames take when they share the GPU?  Lanes 0 and 1 must each have
// processed a frame of this size (their volumes are reused as they are; results are not meaningful).  `mask_a` runs on
// lane 0's stream, `mask_b` on lane 1's, concurrently; bit 0 = cost, bit 1 = horizontal pair, bit 2 = vertical sweep + WTA.
// Returns the mean milliseconds from the common start to the completion of both.
int b200sgm_debug_overlap(b200sgm_handle h, int width, int height, int mask_a, int mask_b, int iters, float* ms_out)
{
    if (!h || !ms_out || h->lanes.size() < 2 || iters <= 0) return B200SGM_EINVAL;
    Eff e;
    int rc = make_eff(h, width, height, e);
    if (rc) return rc;
    if (e.W1 <= 0) return B200SGM_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaEvent_t e0, e1, e2, ea;
    CUDA_TRY(h, cudaEventCreate(&e0)); CUDA_TRY(h, cudaEventCreate(&e1)); CUDA_TRY(h, cudaEventCreate(&e2)); CUDA_TRY(h, cudaEventCreate(&ea));
    double ta = 0, tb = 0;
    auto run = [&](Lane& ln, int mask) -> int {
        if (mask & 1) {
            const char* msg = nullptr; int nl = 0;
            if (launch_cost(ln.feat_l, ln.feat_r, ln.C, e, false, h->num_sms, ln.stream, &nl, &msg) != cudaSuccess || msg) return B200SGM_ECUDA;
        }
        if (mask & 2) { int r2 = launch_aggregation(h, ln, e, ln.stream, 1); if (r2) return r2; }
        if (mask & 4) { int r2 = launch_aggregation(h, ln, e, ln.stream, 2); if (r2) return r2; }
        return B200SGM_OK;
    };
    Lane &a = h->lanes[0], &b = h->lanes[1];
    double total = 0;
    for (int i = 0; i < iters + 1; i++) {
        CUDA_TRY(h, cudaEventRecord(e0, a.stream));
        CUDA_TRY(h, cudaStreamWaitEvent(b.stream, e0, 0));
        if ((rc = run(a, mask_a))) return rc;
        CUDA_TRY(h, cudaEventRecord(ea, a.stream));
        if ((rc = run(b, mask_b))) return rc;
        CUDA_TRY(h, cudaEventRecord(e1, b.stream));
        CUDA_TRY(h, cudaStreamWaitEvent(a.stream, e1, 0));
        CUDA_TRY(h, cudaEventRecord(e2, a.stream));
        CUDA_TRY(h, cudaStreamSynchronize(a.stream));
        float ms = 0;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, e0, e2));
        if (i > 0) total += ms;     // first round is a warm-up
        float fa = 0, fb = 0;
        cudaEventElapsedTime(&fa, e0, ea); cudaEventElapsedTime(&fb, e0, e1);
        if (i > 0) { ta += fa; tb += fb; }
    }
    if (getenv("B200SGM_DEBUG_OVERLAP")) fprintf(stderr, "overlap: lane 0 done at %.3f ms, lane 1 done at %.3f ms\n", ta / iters, tb / iters);
    cudaEventDestroy(ea);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    *ms_out = float(total / iters);
    return B200SGM_OK;
}

int b200sgm_debug_set_path(b200sgm_handle h, int path)
{
    if (!h) return B200SGM_EINVAL;
    h->path = path;
    return B200SGM_OK;
}

}  // extern "C"
