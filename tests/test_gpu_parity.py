"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle and the OpenCV golden vectors.
Bit-exact (integer path): zero differing pixels is the only accepted result."""
import json
import os

import numpy as np
import pytest

import b200sgm
from b200sgm import SGBMParams, CONFIGS, synth, Engine
from oracle import oracle, cv2_reference as ref

pytestmark = pytest.mark.gpu


def run_gpu(L, R, p, lanes=1):
    H, W = L.shape
    eng = Engine(0, W, H, p.numDisparities, lanes, p)
    try:
        return eng.compute(L, R)
    finally:
        eng.close()


def test_golden_small_bit_exact(golden_small):
    for i, (L, R, p, want) in enumerate(golden_small):
        got = run_gpu(L, R, p)
        assert np.array_equal(got, want), "golden case %d: %d px differ (%s)" % (i, (got != want).sum(), p)


def test_stage_dumps_match_oracle():
    """Cost volume C, aggregated cost S, WTA+LR output and the median stage, one by one."""
    for W, H, p in [(96, 64, SGBMParams(numDisparities=32)), (130, 50, SGBMParams(numDisparities=48, minDisparity=-8, blockSize=5, mode=1)),
                    (200, 40, SGBMParams(numDisparities=128, minDisparity=9))]:
        L, R = synth.make_pair(W, H, p.numDisparities, p.minDisparity, 3)
        want, st = oracle.compute(L, R, p, dumps=True)
        eng = Engine(0, W, H, p.numDisparities, 1, p)
        eng.set_path(1)   # generic per-direction path materialises both volumes
        got = eng.compute(L, R)
        assert np.array_equal(eng.debug_volume("C", W, H), st["C"].view(np.uint16))
        assert np.array_equal(eng.debug_volume("S", W, H), st["S"].view(np.uint16))
        assert np.array_equal(eng.debug_image("wta", W, H), st["disp_wta"])
        assert np.array_equal(eng.debug_image("median", W, H), st["disp_med"])
        assert np.array_equal(got, want)
        eng.close()


def test_cost_volume_fast_path_tile_geometry():
    """Cost stage (A.2-A.4) of the default path against the oracle's C volume at sizes chosen for the tiled kernels:
    W1 on either side of the 120/124-column tile width, a single-column image, several row segments, padded and
    unpadded disparity counts, block sizes with and without unrolled instantiations, and a preFilterCap whose pixel
    costs no longer fit a byte (falls back to the 32-bit ring kernel)."""
    cases = [  # (W1, H, D, minD, blockSize, preFilterCap)
        (120, 70, 64, 0, 9, 31), (121, 70, 64, 0, 9, 31), (119, 33, 64, 0, 9, 31), (241, 300, 64, 0, 9, 31),
        (1, 20, 16, 0, 9, 31), (5, 20, 16, 3, 5, 31), (124, 40, 48, -8, 5, 63), (250, 150, 256, 0, 9, 31),
        (130, 60, 32, 9, 3, 15), (140, 50, 64, 0, 15, 31), (126, 45, 128, 0, 21, 7), (100, 40, 64, 0, 9, 127),
        (100, 40, 32, 0, 25, 31),
    ]
    for W1, H, D, minD, bs, cap in cases:
        W = W1 + max(minD + D, 0) - min(minD, 0)
        p = SGBMParams(minDisparity=minD, numDisparities=D, blockSize=bs, preFilterCap=cap)
        assert p.w1(W) == W1
        L, R = synth.make_pair(W, H, D, minD, 21 + W1)
        want, st = oracle.compute(L, R, p, dumps=True)
        eng = Engine(0, W, H, D, 1, p)
        got = eng.compute(L, R)
        C = eng.debug_volume("C", W, H)
        eng.close()
        assert np.array_equal(C, st["C"].view(np.uint16)), "C differs in %d cells for case %s" % (
            (C != st["C"].view(np.uint16)).sum(), (W1, H, D, minD, bs, cap))
        assert np.array_equal(got, want)


def test_random_parameter_sweep_vs_oracle():
    rng = np.random.default_rng(99)
    n = 0
    for it in range(60):
        W = int(rng.integers(70, 300)); H = int(rng.integers(24, 100))
        D = int(rng.choice([16, 32, 48, 8, 24, 40, 64, 80, 128, 144, 256])); minD = int(rng.choice([-8, 0, 1, 2, 9, -20, 30]))
        if W - (D + abs(minD)) < 8:
            continue
        p = SGBMParams(minDisparity=minD, numDisparities=D, blockSize=int(rng.choice([3, 5, 9, 15, 8, 21])),
                       P1=int(rng.choice([200, 8, 1800, 0])), P2=int(rng.choice([400, 32, 7200, 0])),
                       disp12MaxDiff=int(rng.choice([0, 1, 2, 5, -1])), preFilterCap=int(rng.choice([7, 31, 63, 1])),
                       uniquenessRatio=int(rng.choice([0, 2, 10, 15, -1, 50])), speckleWindowSize=int(rng.choice([0, 100, 20, 400])),
                       speckleRange=int(rng.choice([4, 2, 1, 0])), mode=int(rng.integers(0, 2)))
        L, R = synth.make_pair(W, H, D, minD, seed=int(rng.integers(1 << 30)))
        if it % 3 == 0:
            R = np.clip(R.astype(int) + rng.integers(-25, 26, R.shape), 0, 255).astype(np.uint8)
        want = oracle.compute(L, R, p)
        got = run_gpu(L, R, p)
        assert np.array_equal(got, want), "%d px differ for %s (%dx%d)" % ((got != want).sum(), p, W, H)
        n += 1
    assert n >= 35


def test_edge_cases():
    p = SGBMParams(numDisparities=64)
    # W1 <= 0: everything INVALID
    L = np.full((8, 40), 7, np.uint8)
    assert (run_gpu(L, L, p) == p.invalid()).all()
    # constant images, saturated images, tiny heights
    for img in (np.zeros((5, 100), np.uint8), np.full((3, 120), 255, np.uint8), np.full((1, 90), 9, np.uint8)):
        q = SGBMParams(numDisparities=16)
        assert np.array_equal(run_gpu(img, img, q), oracle.compute(img, img, q))
    # strided (non-tight) host images
    L, R = synth.make_pair(150, 40, 32, 0, 5)
    Lp = np.zeros((40, 200), np.uint8); Lp[:, :150] = L
    q = SGBMParams(numDisparities=32)
    assert np.array_equal(run_gpu(Lp[:, :150], R, q), oracle.compute(L, R, q))


def test_uniqueness_ratio_extremes():
    """uniquenessRatio 99 (f = 1: the magic-division constant of the threshold does not fit 32 bits), 100 and beyond
    (f <= 0: exact per-cell comparison path), on textured, noisy and flat inputs."""
    rng = np.random.default_rng(3)
    for uniq in (99, 100, 150, 98, 1):
        for kind in range(3):
            W, H, D = 260, 70, 64
            L, R = synth.make_pair(W, H, D, 0, 40 + kind)
            if kind == 1:
                L = rng.integers(0, 256, L.shape).astype(np.uint8); R = rng.integers(0, 256, R.shape).astype(np.uint8)
            elif kind == 2:
                L = (L // 32 * 32).astype(np.uint8); R = (R // 32 * 32).astype(np.uint8)
            for mode in (0, 1):
                p = SGBMParams(numDisparities=D, uniquenessRatio=uniq, mode=mode, speckleWindowSize=0)
                got, want = run_gpu(L, R, p), oracle.compute(L, R, p)
                assert np.array_equal(got, want), "uniq %d kind %d mode %d: %d px differ" % (uniq, kind, mode, (got != want).sum())


def test_error_behaviour():
    eng = Engine(0, 64, 64, 32, 1)
    L = np.zeros((64, 64), np.uint8)
    with pytest.raises(b200sgm.B200SGMError):   # no params yet
        eng.compute(L, L)
    eng.set_params(SGBMParams(numDisparities=64))
    with pytest.raises(b200sgm.B200SGMError):   # exceeds max_disparities
        eng.compute(L, L)
    eng.set_params(SGBMParams(numDisparities=32))
    with pytest.raises(b200sgm.B200SGMError):   # exceeds max size
        eng.compute(np.zeros((80, 64), np.uint8), np.zeros((80, 64), np.uint8))
    with pytest.raises(ValueError):             # "Images MUST be the same resolution"
        eng.compute(L, np.zeros((32, 64), np.uint8))
    eng.set_params(SGBMParams(numDisparities=0))
    with pytest.raises(b200sgm.B200SGMError):
        eng.compute(L, L)
    eng.close()


def test_config_c1_vs_oracle_and_golden_crc(golden_crc):
    c = CONFIGS["c1"]
    L, R = synth.make_pair(c.width, c.height, c.params.numDisparities, 0, 1000)
    got = run_gpu(L, R, c.params)
    assert synth.crc32(got) == golden_crc["c1"]["disp"]
    assert np.array_equal(got, oracle.compute(L, R, c.params))
    # the node's compiled-in default min_disparity (generate_disparity.cpp:100) and a negative one
    for minD in (9, -32):
        p = c.params.replace(minDisparity=minD)
        L, R = synth.make_pair(c.width, c.height, 64, minD, 1001)
        assert np.array_equal(run_gpu(L, R, p), oracle.compute(L, R, p))


def test_config_c2_hh_golden_crc(golden_crc):
    c = CONFIGS["c2"]
    L, R = synth.make_pair(c.width, c.height, c.params.numDisparities, 0, 1000)
    got = run_gpu(L, R, c.params)
    assert synth.crc32(got) == golden_crc["c2"]["disp"]
    assert int(got.astype(np.int64).sum()) == golden_crc["c2"]["disp_sum"]


def test_config_c3_full_size_golden_crc_and_properties(golden_crc):
    c = CONFIGS["c3"]
    p = c.params
    L, R = synth.make_pair(c.width, c.height, p.numDisparities, 0, 1000)
    eng = Engine(0, c.width, c.height, p.numDisparities, 2, p)
    got = eng.compute(L, R)
    assert synth.crc32(got) == golden_crc["c3"]["disp"]
    assert int(got.astype(np.int64).sum()) == golden_crc["c3"]["disp_sum"]
    # size-independent properties: left band invalid, idempotent post filters, determinism across lanes
    assert (got[:, :p.min_x1()] == p.invalid()).all()
    out2 = np.empty_like(got)
    eng.enqueue(1, L, R, out2); eng.wait(1)
    assert np.array_equal(out2, got)
    if ref.have_cv2():
        import cv2
        f = got.copy()
        cv2.filterSpeckles(f, p.invalid(), p.speckleWindowSize, 16 * p.speckleRange)
        assert np.array_equal(f, got)      # speckle filter is idempotent on its own output
    eng.close()


def test_config_cL_reference_launch_default(golden_crc):
    """cL = what `roslaunch stereo_matcher.launch stereo_algorithm:=1` installs (/root/reference/launch/stereo_matcher.launch:37-48):
    2448x2048, minD 147, 480 disparities, window 21, uniqueness 2, speckle 1000/4, cap 7.  Pinned by the cv2 golden CRC (and by
    cv2 live when importable); must run on the fused path (480 disparities: 8 packed registers per lane) without warnings."""
    c = CONFIGS["cL"]
    p = c.params
    L, R = synth.make_pair(c.width, c.height, p.numDisparities, p.minDisparity, 1000)
    assert synth.crc32(L) == golden_crc["cL"]["left"] and synth.crc32(R) == golden_crc["cL"]["right"]
    eng = Engine(0, c.width, c.height, p.numDisparities, 1, p)
    got = eng.compute(L, R)
    assert eng.last_warning is None
    assert synth.crc32(got) == golden_crc["cL"]["disp"]
    assert int(got.astype(np.int64).sum()) == golden_crc["cL"]["disp_sum"]
    assert (got[:, :p.min_x1()] == p.invalid()).all()
    if ref.have_cv2():
        L2, R2 = synth.make_pair(c.width, c.height, p.numDisparities, p.minDisparity, 1001)
        assert np.array_equal(eng.compute(L2, R2), ref.compute(L2, R2, p))
    eng.close()


def test_wide_disparity_ranges_on_the_fused_path():
    """Disparity counts above 256 (8 and 16 packed registers per lane, shallower rings) against the oracle, incl. padded counts."""
    for W, H, D, minD, bs in [(700, 60, 480, 147, 21), (640, 48, 512, 0, 9), (600, 40, 320, -16, 5), (1400, 30, 1024, 0, 5)]:
        p = SGBMParams(minDisparity=minD, numDisparities=D, blockSize=bs, preFilterCap=7 if bs == 21 else 31)
        L, R = synth.make_pair(W, H, D, minD, 5 + D)
        for mode in (0, 1):
            q = p.replace(mode=mode)
            assert np.array_equal(run_gpu(L, R, q), oracle.compute(L, R, q)), (W, H, D, minD, bs, mode)


def test_halo_agents_on_narrow_strips():
    """Strips of 5 to 10 columns up to 128 disparities run the vertical sweep with halo agents (the neighbour's record is used three
    rows after it was written, strips drift apart): both modes, few and many rows, widths either side of the eligibility limits,
    a one-lane and a several-lane engine (whose strips are wide: no halo) must agree with the oracle and with each other."""
    for W, H, D, minD in [(900, 37, 64, 0), (1000, 150, 128, 3), (1290, 9, 64, -8), (1560, 64, 32, 0), (820, 5, 64, 0), (1700, 3, 128, 0)]:
        p = SGBMParams(minDisparity=minD, numDisparities=D, P1=24, P2=96, uniquenessRatio=5)
        L, R = synth.make_pair(W, H, D, minD, 900 + W)
        for mode in (0, 1):
            q = p.replace(mode=mode)
            want = oracle.compute(L, R, q)
            assert np.array_equal(run_gpu(L, R, q), want), (W, H, D, minD, mode)
            eng = Engine(0, W, H, D, 3, q)
            try:
                assert np.array_equal(eng.compute(L, R), want), ("3 lanes", W, H, D, minD, mode)
            finally:
                eng.close()


def test_cost_range_warning_and_lane_status():
    """int16 contract guard: a frame whose largest cost-volume cell + P2 exceeds 32767 is delivered with a warning
    (B200SGM_WARN_COST_RANGE) -- uncorrelated noise through a 21x21 window at the launch-default cap; the same images
    inside the contract (window 9) raise nothing.  A failed or flagged frame does not poison the next one."""
    rng = np.random.default_rng(5)
    W, H, D = 400, 120, 64
    L = rng.integers(0, 256, (H, W)).astype(np.uint8)
    R = rng.integers(0, 256, (H, W)).astype(np.uint8)
    eng = Engine(0, W, H, D, 1, SGBMParams(numDisparities=D, blockSize=21, preFilterCap=63, P2=8000))
    eng.compute(L, R)
    assert eng.last_warning is not None and eng.last_warning[0] == 1, eng.last_warning
    assert eng.lane_status(0) == 1
    p = SGBMParams(numDisparities=D)
    eng.set_params(p)
    got = eng.compute(L, R)
    assert eng.last_warning is None and eng.lane_status(0) == 0
    assert np.array_equal(got, oracle.compute(L, R, p))
    eng.close()


@pytest.mark.skipif(not ref.have_cv2(), reason="cv2 not importable")
def test_live_cv2_reference_call_sequence():
    """Same inputs through the reference's own call sequence on OpenCV (oracle/cv2_reference.py)."""
    for name, seed in (("c1", 1234),):
        c = CONFIGS[name]
        L, R = synth.make_pair(c.width, c.height, c.params.numDisparities, 0, seed)
        assert np.array_equal(run_gpu(L, R, c.params), ref.compute(L, R, c.params))


def test_f32_output_and_streaming_lanes():
    """a10: forwardMatch leaves CV_32FC1 holding the x16 value; lanes give identical results in any order."""
    p = SGBMParams(numDisparities=64)
    W, H = 320, 200
    frames = [synth.make_pair(W, H, 64, 0, 50 + i) for i in range(6)]
    want = [oracle.compute(L, R, p) for L, R in frames]
    eng = Engine(0, W, H, 64, 3, p)
    f32 = eng.compute_f32(*frames[0])
    assert f32.dtype == np.float32 and np.array_equal(f32, oracle.to_float(want[0]))
    outs = [np.empty((H, W), np.int16) for _ in frames]
    for i, (L, R) in enumerate(frames):
        if i >= 3:
            eng.wait(i % 3)
        eng.enqueue(i % 3, L, R, outs[i])
    for ln in range(3):
        eng.wait(ln)
    for o, w in zip(outs, want):
        assert np.array_equal(o, w)
    eng.close()


def test_device_resident_path_torch():
    import torch
    p = SGBMParams(numDisparities=64)
    W, H = 320, 200
    L, R = synth.make_pair(W, H, 64, 0, 77)
    eng = Engine(0, W, H, 64, 1, p)
    dL, dR = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    dD = torch.empty((H, W), dtype=torch.int16, device="cuda")
    s = torch.cuda.Stream()
    eng.compute_device(0, dL.data_ptr(), W, dR.data_ptr(), W, W, H, dD.data_ptr(), W * 2, stream=s.cuda_stream)
    s.synchronize()
    assert np.array_equal(dD.cpu().numpy(), oracle.compute(L, R, p))
    eng.close()


def test_reprojection_c5_style():
    """Rows a11 + R: /16 + depth window + float32 reprojection, compacted row-major; exact float32 equality
    (tolerance 0: the kernel uses non-contracted mul/add/div, IEEE round-to-nearest, like the x86 reference)."""
    p = SGBMParams(numDisparities=64)
    W, H = 320, 240
    L, R = synth.make_pair(W, H, 64, 0, 11)
    cam = b200sgm.C5_CAMERA
    q = oracle.calc_q(cam["fx"], W / 2.0, W / 2.0, H / 2.0, cam["p14"])
    fT = np.float32(0.3 * 2400.0)
    depth_max = 60.0   # D=64 only reaches 11 m and beyond at this focal length / baseline
    min_disp = float(fT / np.float32(depth_max))
    eng = Engine(0, W, H, 64, 1, p)
    disp, dmat, depth, pts, n = eng.compute_xyz(L, R, q, cam["depth_min"], depth_max, min_disp, float("inf"))
    want = oracle.compute(L, R, p)
    assert np.array_equal(disp, want)
    wdm = oracle.process_disparity(want, min_disp, float("inf"))
    assert np.array_equal(dmat, wdm)
    wdepth, wpts = oracle.reproject(wdm, L, q, cam["depth_min"], depth_max)
    assert n == wpts.shape[0] and n > 1000
    assert np.array_equal(depth, wdepth)
    assert np.array_equal(pts.view(np.uint32), wpts.view(np.uint32))
    # BGR8 colour source (disparity_to_depth.cpp:117-125, :185-188) and a depth window whose bounds are not float32 values
    bgr = np.random.default_rng(2).integers(0, 256, (H, W, 3)).astype(np.uint8)
    lo, hi = oracle.disparity_window(cam["fx"], 0.3, 0.35, 41.3)
    disp, dmat, depth, pts, n = eng.compute_xyz(L, R, q, 0.35, 41.3, lo, hi, color=bgr)
    wdm = oracle.process_disparity(want, lo, hi)
    wdepth, wpts = oracle.reproject(wdm, bgr, q, 0.35, 41.3)
    assert np.array_equal(dmat, wdm) and np.array_equal(depth, wdepth)
    assert n == wpts.shape[0] and np.array_equal(pts.view(np.uint32), wpts.view(np.uint32))
    # the C helper that forms q and the window like the reference does, against the oracle's statement of the same rules
    Kl = np.array([[cam["fx"], 0, W / 2.0], [0, cam["fx"], H / 2.0], [0, 0, 1.0]])
    Pl = np.array([[cam["fx"], 0, W / 2.0, 0], [0, cam["fx"], H / 2.0, 0], [0, 0, 1.0, 0]])
    Pr = Pl.copy(); Pr[0, 3] = cam["p14"]
    rp = b200sgm.reproject_from_camera(Kl, Pl, Pr, 0.35, 41.3)
    assert np.array_equal(np.array([rp.q03, rp.q13, rp.wz, rp.q32, rp.q33], np.float32), q)
    assert (rp.min_disparity, rp.max_disparity) == (np.float32(lo), np.float32(hi))
    eng.close()


def test_config_c5_full_size_reprojection(golden_crc):
    """c5: c3 + processDisparity + reprojection at full size (2448x2048x256, SURVEY 8d camera).  The disparity is pinned by
    the golden CRC; dmat, depth and the compacted XYZRGB list must equal the oracle's restatement of
    generate_disparity.cpp:436-452 / disparity_to_depth.cpp:136-205 bit for bit (float32, tolerance 0)."""
    import time
    c = CONFIGS["c5"]
    p = c.params
    W, H = c.width, c.height
    L, R = synth.make_pair(W, H, p.numDisparities, 0, 1000)
    cam = b200sgm.C5_CAMERA
    q = oracle.calc_q(cam["fx"], cam["cx"], cam["cx"], cam["cy"], cam["p14"])
    fT = np.float32(0.3 * 2400.0)
    depth_min, depth_max = cam["depth_min"], cam["depth_max"]
    min_disp = float(fT / np.float32(depth_max))
    max_disp = float("inf") if depth_min == 0 else float(fT / np.float32(depth_min))
    eng = Engine(0, W, H, p.numDisparities, 1, p)
    eng.compute_xyz(L, R, q, depth_min, depth_max, min_disp, max_disp)           # warm-up
    t0 = time.perf_counter()
    disp, dmat, depth, pts, n = eng.compute_xyz(L, R, q, depth_min, depth_max, min_disp, max_disp)
    dt = time.perf_counter() - t0
    eng.close()
    assert synth.crc32(disp) == golden_crc["c3"]["disp"]
    wdm = oracle.process_disparity(disp, min_disp, max_disp)
    assert np.array_equal(dmat, wdm)
    wdepth, wpts = oracle.reproject(wdm, L, q, depth_min, depth_max)
    assert n == wpts.shape[0] and n > 100000
    assert np.array_equal(depth, wdepth)
    assert np.array_equal(pts.view(np.uint32), wpts.view(np.uint32))
    print("c5 end to end (pageable host buffers, %d points): %.1f ms" % (n, dt * 1e3))
