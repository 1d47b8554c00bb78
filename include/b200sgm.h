/*
 * b200sgm.h -- C ABI of the B200-native semi-global block matching (SGBM) disparity engine.
 *
 * This is the drop-in boundary for ONE path of i3drobotics/i3dr_stereo_camera-ros: the
 * disparity computation that the reference's MatcherOpenCVSGBM plugin delegates to
 * cv::StereoSGBM::compute (reference: src/stereoMatcher/matcherOpenCVSGBM.cpp:14-44), plus the
 * float conversions and the reprojection either side of it.  A reference maintainer binds these
 * entry points from a new AbstractStereoMatcher subclass (see INTEGRATION.md and
 * i3dr_stereo_camera-ros_b200/host/matcherB200SGM.{h,cpp}).
 *
 * Conventions: every function returns 0 on success and a negative B200SGM_E* code on failure;
 * no C++ types and no exceptions cross the boundary; the caller owns every buffer it passes;
 * "stride" arguments are in BYTES.  There is no CPU fallback: if no CUDA device is usable,
 * b200sgm_create fails with B200SGM_ECUDA.
 */
#ifndef B200SGM_H
#define B200SGM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SGM_OK 0
#define B200SGM_EINVAL (-1)   /* bad argument / parameter outside the supported contract      */
#define B200SGM_ECUDA (-2)    /* CUDA runtime error (no device, launch failure, out of memory) */
#define B200SGM_ESIZE (-3)    /* image larger than the engine was created for                  */
#define B200SGM_ESTATE (-4)   /* call sequence error (e.g. wait on an idle lane)               */
/* Positive = warning: the result was produced and delivered, but the frame left cv::StereoSGBM's defined range: some
 * cost-volume cell C satisfied C + P2 > 32767, where OpenCV's int16 arithmetic wraps or saturates depending on its
 * build (possible only when blockSize^2 * (2*max(preFilterCap,15)|1 + 63) + P2 > 32767, e.g. the launch-file default
 * window 21).  Returned by b200sgm_wait / b200sgm_compute* / b200sgm_lane_status; b200sgm_last_error has the figure. */
#define B200SGM_WARN_COST_RANGE 1

#define B200SGM_MODE_SGBM 0   /* 5 aggregation paths, single top-down sweep (cv::StereoSGBM::MODE_SGBM) */
#define B200SGM_MODE_HH 1     /* 8 aggregation paths, two sweeps            (cv::StereoSGBM::MODE_HH)   */

typedef struct b200sgm_engine *b200sgm_handle;

/*
 * RAW parameter values, exactly what the reference's setters hand to cv::StereoSGBM
 * (matcherOpenCVSGBM.cpp:53-110, driven by generate_disparity.cpp:245-256, parameter set
 * cfg/i3DR_Disparity.cfg:21-39).  The engine applies OpenCV's defaulting rules itself:
 * uniquenessRatio<0 -> 10, disp12MaxDiff<=0 -> 1, P1<=0 -> 2, P2<=0 -> 5, P2 = max(P2, P1+1),
 * blockSize<=0 -> 5 (even sizes act as the next odd), ftzero = max(preFilterCap,15)|1.
 */
typedef struct b200sgm_params {
    int minDisparity;      /* setMinDisparity        matcherOpenCVSGBM.cpp:53-57   */
    int numDisparities;    /* setDisparityRange      matcherOpenCVSGBM.cpp:59-64   */
    int blockSize;         /* setWindowSize          matcherOpenCVSGBM.cpp:66-70   */
    int P1;                /* setP1                  matcherOpenCVSGBM.cpp:97-100  */
    int P2;                /* setP2                  matcherOpenCVSGBM.cpp:102-105 */
    int disp12MaxDiff;     /* setDisp12MaxDiff       matcherOpenCVSGBM.cpp:87-90   */
    int preFilterCap;      /* setPreFilterCap        matcherOpenCVSGBM.cpp:107-110 */
    int uniquenessRatio;   /* setUniquenessRatio     matcherOpenCVSGBM.cpp:72-75   */
    int speckleWindowSize; /* setSpeckleFilterWindow matcherOpenCVSGBM.cpp:77-80   */
    int speckleRange;      /* setSpeckleFilterRange  matcherOpenCVSGBM.cpp:82-85   */
    int mode;              /* B200SGM_MODE_*; the reference leaves MODE_SGBM (create(64,9,5), :14) */
} b200sgm_params;

/* One point of the reprojected cloud: pcl::PointXYZRGB-compatible payload (disparity_to_depth.cpp:176-199). */
typedef struct b200sgm_point {
    float x, y, z;
    uint32_t rgb; /* 0x00RRGGBB */
} b200sgm_point;

/* Reprojection inputs: calc_q() of disparity_to_depth.cpp:62-85 cast to float as :136-140 does. */
typedef struct b200sgm_reproject {
    float q03, q13, wz, q32, q33; /* -cx, -cy, fx, 1/T, -(cx-cxr)/T                                  */
    float min_disparity;          /* float(T*f/depth_max)  (generate_disparity.cpp:449)               */
    float max_disparity;          /* float(T*f/depth_min)  (generate_disparity.cpp:450); +inf if depth_min=0 */
    double depth_min, depth_max;  /* cfg/i3DR_pointCloud.cfg as the DOUBLES the reference compares the float Z with
                                     (disparity_to_depth.cpp:50-51, :174-175)                         */
    /* colour source of the cloud (disparity_to_depth.cpp:111-125, :176-188): a host image of the frame's size, MONO8
     * (color_channels 1) or BGR8 (color_channels 3); NULL = the left image handed to the matcher (MONO8) */
    const uint8_t *color;
    size_t color_stride;
    int color_channels;
} b200sgm_reproject;
/* Fills q03..q33 and the disparity window exactly as the reference does: calc_q() on (K_l, P_r, P_l) in double then cast to
 * float (disparity_to_depth.cpp:62-85, :136-140); f = P_l[0][0] and T = -P_r[0][3] / P_r[0][0] (image_geometry baseline) stored
 * as float32 message fields, min/max_disparity = float(double(f32(T * f)) / depth) (generate_disparity.cpp:441-450).
 * K: 3x3, P_l / P_r: 3x4, row-major.  Leaves `color` NULL. */
void b200sgm_reproject_from_camera(b200sgm_reproject *rp, const double *K_l, const double *P_l, const double *P_r,
                                   double depth_min, double depth_max);

/* ---- lifetime ------------------------------------------------------------------------------------------- */

/* Replaces `new MatcherOpenCVSGBM(param_file, size)` + init() (matcherOpenCVSGBM.h:10-14, .cpp:3-15;
 * created lazily with the first frame's size by init_matcher, generate_disparity.cpp:263-279).
 * Allocates all device memory for images up to max_width x max_height and max_disparities up front so
 * that later calls do not allocate.  `lanes` >= 1 is the number of frames that may be in flight. */
int b200sgm_create(int device, int max_width, int max_height, int max_disparities, int lanes, b200sgm_handle *out);
/* Same for an engine that only ever runs the block matcher (b200sgm_bm_*) and rectification: no aggregated-cost volume,
 * checkpoints or exchange records (about half the memory).  The SGBM entry points fail with B200SGM_ESTATE on it. */
int b200sgm_create_bm(int device, int max_width, int max_height, int max_disparities, int lanes, b200sgm_handle *out);
int b200sgm_destroy(b200sgm_handle h);
/* Page-locked host memory for the buffers a caller hands to the host-pointer entry points (what makes b200sgm_enqueue
 * asynchronous and the synchronous calls copy at full PCIe rate); a matcher keeps its result Mat in such a buffer. */
int b200sgm_host_alloc(size_t bytes, void **ptr);
int b200sgm_host_free(void *ptr);

/* Replaces the 12 setters pushed by updateMatcher() (generate_disparity.cpp:241-261). Cheap, lazy: only
 * stores and validates; takes effect on the next compute/enqueue. */
int b200sgm_set_params(b200sgm_handle h, const b200sgm_params *p);
/* Returns the parameters after OpenCV's defaulting rules (Appendix A.1) -- what the kernels will use. */
int b200sgm_get_effective_params(b200sgm_handle h, b200sgm_params *out);

/* ---- the hot path ----------------------------------------------------------------------------------------- */

/* Replaces matcher->compute(*left, *right, disparity_lr) (matcherOpenCVSGBM.cpp:21).
 * left/right: CV_8UC1 host images; disp: CV_16S host image, disparity x16, invalid = (minDisparity-1)*16.
 * Synchronous: host->device copy, kernels, device->host copy. */
int b200sgm_compute(b200sgm_handle h, const uint8_t *left, size_t left_stride, const uint8_t *right,
                    size_t right_stride, int width, int height, int16_t *disp, size_t disp_stride);

/* Same, but the result is what forwardMatch() must leave in disparity_lr: CV_32FC1 holding the SAME
 * numeric value (still x16) -- disparity_lr.convertTo(CV_32FC1), matcherOpenCVSGBM.cpp:34 and
 * AbstractStereoMatcher::match(), abstractStereoMatcher.cpp:44-53. */
int b200sgm_compute_f32(b200sgm_handle h, const uint8_t *left, size_t left_stride, const uint8_t *right,
                        size_t right_stride, int width, int height, float *disp32, size_t disp_stride);

/* Device-resident variant: all pointers are device memory on the engine's device; work is enqueued on
 * `cuda_stream` (a cudaStream_t, may be 0) of lane `lane` and NOT synchronised. */
int b200sgm_compute_device(b200sgm_handle h, int lane, const uint8_t *d_left, size_t left_stride,
                           const uint8_t *d_right, size_t right_stride, int width, int height,
                           int16_t *d_disp, size_t disp_stride, void *cuda_stream);

/* Streaming: enqueue a host frame on lane `lane` (asynchronous when the host buffers are page-locked),
 * later wait for it.  Frames on different lanes overlap copies and kernels. */
int b200sgm_enqueue(b200sgm_handle h, int lane, const uint8_t *left, size_t left_stride, const uint8_t *right,
                    size_t right_stride, int width, int height, int16_t *disp, size_t disp_stride);
int b200sgm_wait(b200sgm_handle h, int lane);
/* Status of the most recent frame of `lane` for callers of b200sgm_compute_device: call after synchronising the
 * stream the frame ran on.  B200SGM_OK, B200SGM_WARN_COST_RANGE, or B200SGM_ECUDA when the frame's aggregation kernel
 * gave up on an inter-strip wait (its disparity is then invalid; the next frame starts from a clean state). */
int b200sgm_lane_status(b200sgm_handle h, int lane);

/* ---- the steps either side of the path (rows a11 and R of SURVEY.md section 8) ----------------------------- */

/* Matching fused with processDisparity() (generate_disparity.cpp:436-452: /16, depth-window -> 10000) and
 * the reprojection of disparity_to_depth.cpp:136-205.  Any of dmat/depth/points may be NULL.
 * dmat, depth: float H x W (tight); points: capacity W*H; *count = number of points, row-major order. */
int b200sgm_compute_xyz(b200sgm_handle h, const uint8_t *left, size_t left_stride, const uint8_t *right,
                        size_t right_stride, int width, int height, const b200sgm_reproject *rp,
                        int16_t *disp, size_t disp_stride, float *dmat, float *depth,
                        b200sgm_point *points, uint32_t *count);

/* ---- the step before the path (row N2 of SURVEY.md section 8f): rectification ---------------------------------- */

/* Replaces rectify() of generate_disparity.cpp:370-386 (and rectify.cpp:111-127):
 *   cv::initUndistortRectifyMap(K, D, R, P, image.size(), CV_32FC1, map1, map2);
 *   cv::remap(image, image_rect, map1, map2, cv::INTER_CUBIC, cv::BORDER_CONSTANT);
 * The reference rebuilds the maps for every frame; here b200sgm_set_camera installs the CameraInfo of camera
 * `cam` (0 = left, 1 = right) and the maps are built on the device once per (camera, image size).
 * K, R: 3x3 row-major (R may be NULL = identity); D: nD <= 12 distortion coefficients in OpenCV order
 * (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4; a non-zero tilt tauX/tauY is rejected); P: 3x4 row-major. */
int b200sgm_set_camera(b200sgm_handle h, int cam, const double *K, const double *D, int nD, const double *R,
                       const double *P);
/* CV_8UC1 host image in, rectified CV_8UC1 host image out (same size), synchronous.  Bit-exact with cv::remap. */
int b200sgm_rectify(b200sgm_handle h, int cam, const uint8_t *src, size_t src_stride, int width, int height,
                    uint8_t *dst, size_t dst_stride);
/* Device-resident variant, enqueued on `cuda_stream` (or the stream of lane `lane` when NULL), not synchronised:
 * rectify straight into the buffers handed to b200sgm_compute_device. */
int b200sgm_rectify_device(b200sgm_handle h, int lane, int cam, const uint8_t *d_src, size_t src_stride, int width,
                           int height, uint8_t *d_dst, size_t dst_stride, void *cuda_stream);
/* The CV_32FC1 maps of cv::initUndistortRectifyMap for `cam` at this size, copied to the host (H x W, tight). */
int b200sgm_rectify_maps(b200sgm_handle h, int cam, int width, int height, float *map1, float *map2);

/* ---- the package's other OpenCV matcher (row N4 of SURVEY.md section 8f): StereoBM ----------------------------- */

/* Raw cv::StereoBM parameters as MatcherOpenCVBlock's setters install them (matcherOpenCVBlock.cpp:52-110 on top of
 * StereoBM::create(64, 9), :13-16); PREFILTER_XSOBEL (OpenCV's default; the reference never changes it).  disp12MaxDiff must
 * stay negative (OpenCV's default; generate_disparity.cpp:241-261 does not forward it). */
typedef struct b200sgm_bm_params {
    int minDisparity, numDisparities, blockSize, preFilterCap, textureThreshold, uniquenessRatio, speckleWindowSize,
        speckleRange, disp12MaxDiff;
    /* setPreFilterSize (matcherOpenCVBlock.cpp; pushed by generate_disparity.cpp:255).  PREFILTER_XSOBEL never reads it, but
     * cv::StereoBM::compute rejects values outside 5..255 or even ones all the same, and so does this engine; 0 = OpenCV's
     * default (9). */
    int preFilterSize;
} b200sgm_bm_params;

/* matcher->compute(*left, *right, disparity_lr) of MatcherOpenCVBlock::forwardMatch (matcherOpenCVBlock.cpp:20): CV_8UC1 host
 * images in, CV_16S disparity x16 out, filtered pixels = (minDisparity-1)*16.  Synchronous, lane 0.  Bit-exact with
 * cv::StereoBM 4.13, including the `minDisparity` pixels at the start of the first row below the valid ROI when
 * minDisparity > 0, where OpenCV leaves the values it wrote past the end of the last ROI row (reproduced: the reference's
 * launch default has minDisparity 147).  An empty valid ROI is outside the contract (OpenCV returns uninitialised memory). */
int b200sgm_bm_compute(b200sgm_handle h, const b200sgm_bm_params *p, const uint8_t *left, size_t left_stride,
                       const uint8_t *right, size_t right_stride, int width, int height, int16_t *disp,
                       size_t disp_stride);

/* Same, but the result is what MatcherOpenCVBlock::forwardMatch leaves in disparity_lr: CV_32FC1 holding the same x16 value
 * (disparity_lr.convertTo(disparity_lr, CV_32FC1), matcherOpenCVBlock.cpp:34), converted on the device. */
int b200sgm_bm_compute_f32(b200sgm_handle h, const b200sgm_bm_params *p, const uint8_t *left, size_t left_stride,
                           const uint8_t *right, size_t right_stride, int width, int height, float *disp32,
                           size_t disp_stride);

/* Device-resident variant, enqueued on `cuda_stream` (or the stream of lane `lane` when NULL), not synchronised. */
int b200sgm_bm_compute_device(b200sgm_handle h, int lane, const b200sgm_bm_params *p, const uint8_t *d_left,
                              size_t left_stride, const uint8_t *d_right, size_t right_stride, int width, int height,
                              int16_t *d_disp, size_t disp_stride, void *cuda_stream);

/* ---- diagnostics ------------------------------------------------------------------------------------------ */

const char *b200sgm_last_error(b200sgm_handle h);
const char *b200sgm_version(void);
/* Number of kernel launches issued by this engine since creation (bench.py's gpu_launches). */
int b200sgm_launch_count(b200sgm_handle h, uint64_t *count);
/* The engine's internal stream for `lane` (a cudaStream_t) -- lets a caller time with CUDA events on the
 * stream the kernels are launched on. */
int b200sgm_lane_stream(b200sgm_handle h, int lane, void **cuda_stream);
/* Stage profiling: when enabled every frame records CUDA events at the stage boundaries on the stream it
 * runs on.  b200sgm_stage_times synchronises, returns the accumulated milliseconds of the 7 stages
 * {prefilter, cost, horizontal, vertical+wta, lrcheck, median, speckle} since the previous call, the number of frames
 * they cover, and resets the accumulators. */
int b200sgm_profile(b200sgm_handle h, int enable);
int b200sgm_stage_times(b200sgm_handle h, int lane, double *ms, int n, uint64_t *frames);
/* Timeline of the frames harvested by b200sgm_stage_times since profiling was enabled: 8 stage-boundary
 * timestamps per frame, in ms since b200sgm_profile(h, 1).  *n_floats receives the number available. */
int b200sgm_stage_timeline(b200sgm_handle h, int lane, float *out, int max_floats, int *n_floats);
/* Measures the packed 16-bit integer issue peak of `device` with a register-resident VIMNMX3.U16x2 /
 * VIADD.16x2 microbenchmark (the ALU roofline denominator).  Results in 1e12
 * elementary 16-bit ops per second: [0] = 3-input min (4 ops per lane-instruction), [1] = 2-input min
 * (2 ops), [2] = the aggregation instruction mix. */
int b200sgm_alu_peak(int device, double tera_ops[3]);
/* Copies an internal stage buffer of `lane` to the host (tests only): what = "C" | "S" (volumes
 * [H][W1][Dp] uint16, *dp receives the padded disparity count), "wta" | "median" (H x W int16). */
int b200sgm_debug_read(b200sgm_handle h, int lane, const char *what, void *host, size_t bytes, int *dp);
/* Selects the kernel path: 0 = default (fastest validated), 1 = generic per-direction chains. Tests only. */
int b200sgm_debug_set_path(b200sgm_handle h, int path);
/* Development probe (tools/overlap_probe.py): mean milliseconds for stage set `mask_a` on lane 0 and `mask_b` on lane 1
 * (bit 0 cost, bit 1 horizontal pair, bit 2 vertical sweep + WTA) launched together on their two streams.  Both lanes
 * must have processed a frame of this size; the volumes are reused as they are and the results are meaningless. */
int b200sgm_debug_overlap(b200sgm_handle h, int width, int height, int mask_a, int mask_b, int iters, float *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* B200SGM_H */
