// Internal host-side types of the engine shared by engine.cu and the per-N aggregation translation units.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>
#include <mutex>
#include <atomic>

#include "../../include/b200sgm.h"
#include "sgm_types.h"
#include "stages.h"


using namespace b200sgm;

struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
    // device buffers
    uint8_t *left = nullptr, *right = nullptr;      // W x H, pitch = W
    Feat *feat_l = nullptr, *feat_r = nullptr;      // W x H
    uint16_t *C = nullptr, *S = nullptr;            // [H][W1][Dp], paired layout (k_path.cuh)
    uint16_t *ckpt = nullptr;                       // k_horiz checkpoints [H][ceil(W1/kHT)][Dp]
    uint32_t* disp2key = nullptr;                   // W x H
    int16_t *disp_wta = nullptr, *disp_med = nullptr, *disp_out = nullptr;  // W x H
    int *label = nullptr, *csize = nullptr, *parent = nullptr, *runlen = nullptr;   // W x H (speckle filter)
    float *f32a = nullptr, *f32b = nullptr;         // W x H (dmat / depth / CV_32F disparity)
    float4* points = nullptr;                       // W x H
    uint32_t *block_count = nullptr, *total = nullptr;
    uint32_t* h_total = nullptr;                    // pinned
    uint8_t* color = nullptr;                       // colour image of the point cloud (BGR8 / MONO8), allocated on first use
    uint2* xbuf = nullptr;                          // k_vert exchange records (LL protocol)
    // device status words of the frame in flight: [0] inter-strip wait timed out, [1] late exchange records (debug),
    // [2] poll iterations (debug), [3] largest cost-volume cell seen by k_horiz (overflow guard)
    int* d_err = nullptr;
    int* h_err = nullptr;                           // pinned copy of the 4 words, refreshed at the end of every frame
    int last_P2 = 0;                                // P2 of the frame whose status h_err holds
    // stage profiling (b200sgm_profile): ring of event sets, harvested by b200sgm_stage_times
    std::vector<cudaEvent_t> prof_events;           // kProfRing * (kStages + 1)
    int prof_head = 0, prof_count = 0;
    double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // >= kStages
    std::vector<float> timeline;                    // (kStages+1) timestamps per harvested frame, ms since prof_ref
    uint64_t stage_frames = 0;
};

constexpr int kStages = 7;    // prefilter, cost, horizontal, vertical+wta, lrcheck, median, speckle
constexpr int kProfRing = 256;
constexpr int kMaxStrips = 320;
constexpr int kStatusWords = 4;

// Vertical sweeps of ALL engines of a process on one device share the SMs under this gate: the CTAs of a sweep spin on
// their neighbours, so every sweep in flight must be fully resident.  A sweep of n strips takes n SMs (one CTA fills an
// SM's register file); sweeps of narrow images use a fraction of the SMs, and floor(num_sms / n) of them -- of different
// frames -- may run side by side.  Sweep i waits for sweep i - K (hence for every older one of its residue class, and by
// induction at most K are ever running); a sweep whose n differs from its predecessor's waits for all of them.
// Process-wide, keyed by device; the events are never destroyed (they outlive any engine).  Hold `mu` across the wait,
// the launch and the record.
struct CoopGate {
    static constexpr int kRing = 16;
    std::mutex mu;
    cudaEvent_t ev[kRing] = {};
    unsigned long long count = 0;     // sweeps launched so far; sweep i records ev[i % kRing]
    int last_n = 0;
};
CoopGate& coop_gate(int device);
constexpr int kVertMaxWarps = 16;


struct b200sgm_engine {
    int device = 0;
    int maxW = 0, maxH = 0, maxD = 0;
    bool bm_only = false;                         // created by b200sgm_create_bm: no S / checkpoint / exchange buffers
    std::vector<Lane> lanes;
    b200sgm_params raw{};
    bool have_params = false;
    std::string err;
    std::atomic<uint64_t> launches{0};
    int path = 0;
    bool profile = false;
    cudaEvent_t prof_ref = nullptr;   // time origin of the stage timeline
    int num_sms = 148;
    int max_smem_optin = 227 * 1024;
    int max_smem_sm = 228 * 1024;                 // shared memory of an SM (all resident CTAs, incl. 1 KB reserved per CTA)
    int clock_khz = 1965000;
    std::mutex mu;
    // rectification (row N2): per camera (0 left, 1 right) the model and the cached fixed-point maps
    struct Rect {
        bool have = false, dirty = true;
        RectifyCam cam;
        int W = 0, H = 0;
        RemapEntry* ent = nullptr;
        float *map1 = nullptr, *map2 = nullptr;
    } rect[2];
    int16_t* d_wtab = nullptr;                    // cv::remap's bicubic weight table [1024][16]
    uint8_t *rect_src = nullptr, *rect_dst = nullptr;   // staging for the host-pointer call (maxW x maxH)
};


#define CUDA_TRY(h, expr)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                    \
            return B200SGM_ECUDA;                                                              \
        }                                                                                      \
    } while (0)

#define LAUNCH_CHECK(h)                                                                        \
    do {                                                                                       \
        (h)->launches++;                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            (h)->err = std::string("kernel launch at line ") + std::to_string(__LINE__) + ": " + cudaGetErrorString(e__); \
            return B200SGM_ECUDA;                                                              \
        }                                                                                      \
    } while (0)

// Stage profiling: records event `idx` of the current frame's event set (no-op unless profiling is on).
inline void prof_mark(b200sgm_engine* h, Lane& ln, int idx, cudaStream_t st)
{
    if (!h->profile || ln.prof_events.empty()) return;
    cudaEventRecord(ln.prof_events[size_t(ln.prof_head) * (kStages + 1) + idx], st);
}

inline int fail(b200sgm_engine* h, int code, const std::string& msg)
{
    h->err = msg;
    return code;
}


// Aggregation + WTA stages for N packed registers per lane (D <= 64 * N); one explicit instantiation per agg_n*.cu.
template <int N>
int launch_agg_n(b200sgm_engine* h, Lane& ln, const Eff& e, cudaStream_t st, int part);
