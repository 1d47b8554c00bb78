#!/usr/bin/env python
"""Benchmark of the SGBM disparity hot path (BASELINE.json metric: disparity frames/s at 2448x2048x256d).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3]

A "step" is one pass of the hot path over one batch of `--frames` synthetic rectified stereo pairs per GPU.
N > 1 is launched by torchrun (one rank per GPU); frames are independent, so ranks never exchange data
(weak scaling, no collective on the data path); torch.distributed is used only for the barrier and the
max-over-ranks of the device-timed duration.

value      : whole-job frames/s with the input frames resident in HBM (16 distinct pairs, 160 MB > L2).
e2e        : the same metric through the host-buffer C-ABI call (b200sgm_enqueue/wait) with pinned host
             images in and the CV_16S disparity out, copies inside the timed region.
roofline   : dominant stage timed live with CUDA events on the stream it runs on (engine stage profiling).
cpu_baseline / --impl reference : cv::StereoSGBM (cv2, the library the reference calls) with the reference's
             call sequence on the host cores, one matcher per thread, frame-parallel.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import b200sgm  # noqa: E402
from b200sgm import CONFIGS, synth, stream  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=16, help="distinct frames per step per GPU")
    ap.add_argument("--lanes", type=int, default=4, help="frames in flight per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the c1/c2/cL/c5 block and the plugin-call measurement")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU sample (default: one per core)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# CPU arm: cv::StereoSGBM through cv2 (kind "reference"), else the C oracle port (kind "port").
# ---------------------------------------------------------------------------------------------------------
def cpu_sample(cfg, frames, n_threads):
    """Times `len(frames)` frames on `n_threads` host threads. Returns (fps, kind, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cv2_reference as ref
    p = cfg.params
    if ref.have_cv2():
        import cv2
        cv2.setNumThreads(1)
        kind = "reference"
        tl = threading.local()

        def work(fr):
            if not hasattr(tl, "m"):
                tl.m = ref.make_matcher(p)
            return tl.m.compute(fr[0], fr[1])
    else:
        from oracle import oracle
        kind = "port"

        def work(fr):
            return oracle.compute(fr[0], fr[1], p)
    with ThreadPoolExecutor(n_threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(work, frames))
        dt = time.perf_counter() - t0
    return len(frames) / dt, kind, dt


def reference_outputs(cfg, pairs, n_threads):
    """cv::StereoSGBM (cv2, the reference's call sequence) on every pair, frame-parallel; the oracle port when cv2 is absent."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cv2_reference as ref
    p = cfg.params
    if ref.have_cv2():
        import cv2
        cv2.setNumThreads(1)
        tl = threading.local()

        def work(fr):
            if not hasattr(tl, "m"):
                tl.m = ref.make_matcher(p)
            return tl.m.compute(fr[0], fr[1])
        kind = "cv2 %s" % cv2.__version__
    else:
        from oracle import oracle

        def work(fr):
            return oracle.compute(fr[0], fr[1], p)
        kind = "oracle port"
    with ThreadPoolExecutor(max(1, n_threads)) as ex:
        return list(ex.map(work, pairs)), kind


def count_equal(outs, refs):
    return sum(int(np.array_equal(o, r)) for o, r in zip(outs, refs))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 alone
    cores = os.cpu_count() or 1
    p = cfg.params
    n = args.cpu_frames or cores
    frames = [synth.make_pair(cfg.width, cfg.height, p.numDisparities, p.minDisparity, 1000 + i) for i in range(min(n, 4))]
    frames = [frames[i % len(frames)] for i in range(n)]
    for _ in range(args.warmup):
        cpu_sample(cfg, frames[:max(1, min(len(frames), cores))], cores)
    t_total, kind = 0.0, "reference"
    for _ in range(args.steps):
        fps, kind, dt = cpu_sample(cfg, frames, cores)
        t_total += dt
    fps = args.steps * n / t_total
    line = {
        "impl": "reference", "metric": "disparity frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": workload(cfg, n, 0),
        "gpix_disp_per_s": fps * cfg.gpix_disp,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "%d frames per step, one cv::StereoSGBM matcher per thread, %d threads, %s" % (n, cores, cpu_model())},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def launch_summary_path(name):
    """Newest committed ncu launch summary of a config (profiles/rN_launch_summary_<config>.json)."""
    import glob
    key = "c3" if name in ("c4", "c5") else name
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_launch_summary_%s.json" % key)))
    if not cands:
        raise FileNotFoundError(key)
    return cands[-1]


def workload(cfg, frames, lanes):
    p = cfg.params
    return {"workload": "%s: %dx%d pair, %d disparities, blockSize %d, %s, P1/P2 %d/%d, uniq %d, speckle %d/%d" % (
        cfg.name, cfg.width, cfg.height, p.numDisparities, p.blockSize, "MODE_HH 8-path" if p.mode else "MODE_SGBM 5-path",
        p.P1, p.P2, p.uniquenessRatio, p.speckleWindowSize, p.speckleRange),
        "frames_per_step_per_gpu": frames, "lanes_per_gpu": lanes,
        "l2_policy": "inputs larger than L2: %d distinct pairs cycled, cost volumes 2 x %.2f GB per lane" % (
            frames, cfg.width * cfg.height * p.numDisparities * 2 / 1e9)}


# ---------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def run_b200(args, cfg):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0's stdout carries exactly one JSON line: NCCL's version/INFO banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    p = cfg.params
    W, H, D = cfg.width, cfg.height, p.numDisparities
    NF, lanes = args.frames, max(1, min(args.lanes, args.frames))
    eng = b200sgm.Engine(local, W, H, D, lanes, p)

    # ---- synthetic frames: NF distinct pairs per rank; pinned host copies + device-resident copies
    hostL = torch.empty((NF, H, W), dtype=torch.uint8).pin_memory()
    hostR = torch.empty((NF, H, W), dtype=torch.uint8).pin_memory()
    hostD = torch.empty((NF, H, W), dtype=torch.int16).pin_memory()
    n_unique = min(NF, 16)
    for i in range(n_unique):
        L, R = synth.make_pair(W, H, D, p.minDisparity, 1000 + rank * n_unique + i)
        hostL[i].copy_(torch.from_numpy(L)); hostR[i].copy_(torch.from_numpy(R))
    for i in range(n_unique, NF):
        hostL[i].copy_(hostL[i % n_unique]); hostR[i].copy_(hostR[i % n_unique])
    devL, devR = hostL.to(dev), hostR.to(dev)
    devD = torch.empty((NF, H, W), dtype=torch.int16, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    main = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def device_step():
        for i in range(NF):
            ln = i % lanes
            eng.compute_device(ln, devL[i].data_ptr(), W, devR[i].data_ptr(), W, W, H, devD[i].data_ptr(), W * 2,
                               stream=streams[ln].cuda_stream)

    def timed_device(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for _ in range(steps):
            device_step()
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    # ---- parity gate inside the bench: frame 0 of rank 0 (seed 1000) against the committed golden CRC
    checked = None
    eng.compute_device(0, devL[0].data_ptr(), W, devR[0].data_ptr(), W, W, H, devD[0].data_ptr(), W * 2, stream=streams[0].cuda_stream)
    torch.cuda.synchronize()
    if rank == 0:
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_crc.json")))
            key = "c3" if cfg.name in ("c3", "c4", "c5") else cfg.name
            checked = synth.crc32(devD[0].cpu().numpy()) == gold[key]["disp"]
        except Exception as e:  # pragma: no cover
            checked = "unavailable: %s" % e

    for _ in range(args.warmup):
        device_step()
    l0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed_device(args.steps)
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    fps = world * args.steps * NF / (ms_max * 1e-3)

    # ---- e2e: pinned host images in, CV_16S disparity out, through b200sgm_enqueue / b200sgm_wait
    e2e = None
    if not args.no_e2e:
        def host_step():
            stream.run_lanes(lanes, range(NF),
                             lambda ln, i: eng.enqueue_ptr(ln, hostL[i].data_ptr(), W, hostR[i].data_ptr(), W, W, H, hostD[i].data_ptr(), W * 2),
                             eng.wait)
        for _ in range(max(1, args.warmup)):
            host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * args.steps * NF / float(t.item()), "unit": "frames/s",
               "h2d_bytes_per_step": 2 * W * H * NF * world, "d2h_bytes_per_step": 2 * W * H * NF * world,
               "api": "b200sgm_enqueue/b200sgm_wait, pinned host buffers, %d lanes" % lanes}

    # ---- parity of EVERY frame of the timed runs (all lanes; device-resident outputs and the e2e host outputs) against
    # cv::StereoSGBM on the host cores; a mismatch fails the run (exit code 3 after the JSON line)
    cores = os.cpu_count() or 1
    pairs = [(hostL[i].numpy(), hostR[i].numpy()) for i in range(n_unique)]
    refs, ref_kind = reference_outputs(cfg, pairs, max(1, cores // world))
    dev_out = devD.cpu().numpy()
    n_checked = NF + (NF if e2e is not None else 0)
    n_ok = count_equal([dev_out[i] for i in range(NF)], [refs[i % n_unique] for i in range(NF)])
    if e2e is not None:
        n_ok += count_equal([hostD[i].numpy() for i in range(NF)], [refs[i % n_unique] for i in range(NF)])
    pt = torch.tensor([n_checked, n_ok], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.SUM)
    parity_checked, parity_ok = int(pt[0].item()), int(pt[1].item())

    # ---- roofline of the dominant stage: stage events on one lane, frames back to back (rank 0)
    roofline, stages = None, None
    if rank == 0:
        eng.profile(True)
        eng.stage_times(0)
        nprof = min(NF, 8)
        for i in range(nprof):
            eng.compute_device(0, devL[i].data_ptr(), W, devR[i].data_ptr(), W, W, H, devD[i].data_ptr(), W * 2, stream=0)
        st_ms, nfr = eng.stage_times(0)
        eng.profile(False)
        stages = {k: v / max(nfr, 1) for k, v in st_ms.items()}
        stages["aggregate_wta"] = stages.get("horizontal", 0.0) + stages.get("vertical_wta", 0.0)
        total = sum(v for k, v in stages.items() if k != "aggregate_wta")
        dom = max((k for k in stages if k != "aggregate_wta"), key=stages.get)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        W1 = p.w1(W)
        R = 8 if p.mode else 5
        alg_bytes = 4 * W * H if p.mode == 0 else 4 * W * H + 4 * W1 * H * D   # SURVEY 8d, per frame
        alg_ops = W1 * H * D * (30 + 9 * R)                                    # SURVEY 8d, per frame
        frame_s = total * 1e-3
        try:
            alu = b200sgm.alu_peak(local)
        except Exception:
            alu = (None, None, None)
        ach_gbs = alg_bytes / frame_s / 1e9
        ach_tops = alg_ops / frame_s / 1e12
        # measured DRAM traffic per launch from the committed ncu launch list of this config (profiles/), if present
        prof, traffic, kern_traffic = None, None, {}
        try:
            prof = json.load(open(launch_summary_path(cfg.name)))
            traffic = prof.get("frame_dram_bytes")
            for k in prof.get("kernels", []):
                kern_traffic.setdefault(k["kernel"].split("<")[0], []).append(k.get("dram_read_bytes", 0) + k.get("dram_write_bytes", 0))
        except Exception:
            pass
        stage_kernel = {"cost": "k_cost_tile2", "horizontal": "k_horiz", "vertical_wta": "k_vert"}
        # algorithmic int16 ops per cell of each volume stage (BT 21 + block sum 4; 9 per path; WTA 5)
        n_h, n_v = 2, (6 if p.mode else 3)
        stage_ops = {"cost": 25 * W1 * H * D, "horizontal": 9 * n_h * W1 * H * D, "vertical_wta": (9 * n_v + 5) * W1 * H * D}
        per_stage = {}
        for st_name, kname in stage_kernel.items():
            by = sum(kern_traffic.get(kname, [])) or None
            ms_k = stages.get(st_name, 0.0)
            tops_k = stage_ops[st_name] / (ms_k * 1e-3) / 1e12 if ms_k > 0 else None
            per_stage[st_name] = {"kernel": kname, "ms": ms_k, "share_of_frame": ms_k / max(total, 1e-9),
                                  "dram_bytes_per_launch_ncu": by,
                                  "dram_gbs": (by / (ms_k * 1e-3) / 1e9) if by and ms_k > 0 else None,
                                  "frac_of_hbm_peak": (by / (ms_k * 1e-3) / 1e9 / hbm_peak) if by and ms_k > 0 else None,
                                  "algorithmic_tops": tops_k,
                                  "frac_of_alu_peak": (tops_k / alu[0]) if tops_k and alu[0] else None}
        frac_alu = (ach_tops / alu[0]) if alu[0] else None
        frac_hbm = ach_gbs / hbm_peak
        alu_binds = frac_alu is not None and frac_alu >= frac_hbm
        roofline = {
            # two roofs, both reported; the headline fraction is the binding one = max(frac_alu, frac_hbm)
            "bound": "alu" if alu_binds else "hbm",
            "achieved": ach_tops if alu_binds else ach_gbs,
            "peak": alu[0] if alu_binds else hbm_peak,
            "unit": "Tops/s (elementary int16 ops)" if alu_binds else "GB/s",
            "frac": frac_alu if alu_binds else frac_hbm,
            "traffic": traffic,
            "peak_source": ("b200sgm_alu_peak: packed-int16 issue rate measured on this GPU (VIMNMX3.U16x2 x4 ops); min2 %.1f, mix %.1f Tops/s"
                            % (alu[1] or 0, alu[2] or 0)) if alu_binds else "MEASURED_PEAKS.json hbm_gbs",
            "kernel": "whole pipeline of one frame; dominant kernel %s (%s): %.3f ms = %.1f%% of the frame, timed live with CUDA events "
                      "on its stream" % (stage_kernel.get(dom, dom), dom, stages[dom], 100 * stages[dom] / max(total, 1e-9)),
            "algorithmic_ops_per_frame": alg_ops,
            "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": frac_hbm,
                    "algorithmic_bytes_per_frame": alg_bytes,
                    "traffic_gbs": (traffic / frame_s / 1e9) if traffic else None,
                    "traffic_frac_of_peak": (traffic / frame_s / 1e9 / hbm_peak) if traffic else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "note": "achieved = ALGORITHMIC bytes (images in + disparity out, SURVEY 8d) / frame time; traffic = DRAM bytes the "
                            "kernels really move per frame (ncu launch list under profiles/): the two materialised volumes"},
            "kernels": per_stage,
            "stage_ms_per_frame": stages, "single_lane_fps": 1.0 / frame_s,
            # the same ratio at the measured whole-pipeline throughput of this rank (`value` / n_gpus: frames of several lanes overlap)
            "frac_pipelined": (alg_ops * (fps / max(world, 1)) / 1e12 / alu[0]) if alu[0] else None,
        }

    # ---- row N2: rectification of both images in front of the matcher, device-resident (rank 0)
    rect = None
    if rank == 0:
        try:
            for cam in (0, 1):
                eng.set_camera(cam, *synth.sample_camera(W, H, 7 + 2 * cam, 1.0))
            rL, rR = torch.empty_like(devL[0]), torch.empty_like(devR[0])
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            nrep = 20
            for rep in range(nrep + 3):
                if rep == 3:
                    ev[0].record(streams[0])
                i = rep % NF
                eng.rectify_device(0, 0, devL[i].data_ptr(), W, W, H, rL.data_ptr(), W, stream=streams[0].cuda_stream)
                eng.rectify_device(0, 1, devR[i].data_ptr(), W, W, H, rR.data_ptr(), W, stream=streams[0].cuda_stream)
            ev[1].record(streams[0])
            streams[0].synchronize()
            ms_pair = ev[0].elapsed_time(ev[1]) / nrep
            by = 2 * W * H * (1 + 1 + 8)          # source + rectified image + fixed-point map entry per pixel, both cameras
            rect = {"kernel": "k_remap_cubic (x2)", "ms_per_pair": ms_pair, "algorithmic_bytes_per_pair": by,
                    "gbs": by / (ms_pair * 1e-3) / 1e9, "frac_of_hbm_peak": by / (ms_pair * 1e-3) / 1e9 / hbm_peak,
                    "note": "cv::initUndistortRectifyMap + cv::remap(INTER_CUBIC, BORDER_CONSTANT) of generate_disparity.cpp:370-386; "
                            "maps built once per camera; not part of `value` (the matcher's metric)"}
        except Exception as e:  # pragma: no cover
            rect = {"error": str(e)}

    # ---- row N4: the package's other OpenCV matcher, StereoBM, device-resident (rank 0), cv2 beside it at N=1
    bm = None
    if rank == 0:
        try:
            kw = dict(numDisparities=D, blockSize=9, speckleWindowSize=p.speckleWindowSize, speckleRange=p.speckleRange)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            nrep = 8
            for rep in range(nrep + 2):
                if rep == 2:
                    ev[0].record(streams[0])
                i = rep % NF
                eng.bm_compute_device(0, devL[i].data_ptr(), W, devR[i].data_ptr(), W, W, H, devD[i].data_ptr(), W * 2,
                                      stream=streams[0].cuda_stream, **kw)
            ev[1].record(streams[0])
            streams[0].synchronize()
            ms_bm = ev[0].elapsed_time(ev[1]) / nrep
            bm = {"ms_per_frame": ms_bm, "frames_per_s": 1e3 / ms_bm, "params": kw,
                  "note": "cv::StereoBM::compute of matcherOpenCVBlock.cpp:20 (first, unfused version: SAD volume materialised); "
                          "not part of `value`"}
            if world == 1 and not args.no_cpu_baseline:
                try:
                    import cv2
                    m = cv2.StereoBM_create(64, 9)
                    m.setNumDisparities(D); m.setBlockSize(9); m.setSpeckleWindowSize(p.speckleWindowSize); m.setSpeckleRange(p.speckleRange)
                    l0, r0 = hostL[0].numpy(), hostR[0].numpy()
                    ref0 = m.compute(l0, r0)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        m.compute(l0, r0)
                    bm["cpu_cv2_ms_per_frame"] = (time.perf_counter() - t0) / 3 * 1e3
                    bm["cpu_threads"] = cv2.getNumThreads()
                    eng.bm_compute_device(0, devL[0].data_ptr(), W, devR[0].data_ptr(), W, W, H, devD[0].data_ptr(), W * 2,
                                          stream=streams[0].cuda_stream, **kw)
                    streams[0].synchronize()
                    bm["matches_cv2"] = bool(np.array_equal(devD[0].cpu().numpy(), ref0))
                except ImportError:
                    pass
        except Exception as e:  # pragma: no cover
            bm = {"error": str(e)}

    # ---- the other BASELINE.json configurations (+ cL, the reference's launch default), N = 1 only: device-resident fps,
    # stage times, roofline fractions and parity of every timed frame against cv::StereoSGBM
    side, plugin = None, None
    if rank == 0 and world == 1 and not args.no_side:
        side = {}
        for name in ("c1", "c2", "cL", "c5"):
            try:
                # c1's sweeps take a quarter of the SMs each: 8 frames in flight keep 4 of them running side by side
                side[name] = side_config(torch, dev, CONFIGS[name], alu[0] if roofline else None, hbm_peak if roofline else 6457.1,
                                         frames=16 if name == "c1" else 8, lanes=8 if name == "c1" else 4)
            except Exception as e:  # pragma: no cover
                side[name] = {"error": repr(e)}
        try:
            plugin = plugin_call(cfg, hostL, hostR, refs, n_unique)
        except Exception as e:  # pragma: no cover
            plugin = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = args.cpu_frames or cores
        fr = [(hostL[i % n_unique].numpy(), hostR[i % n_unique].numpy()) for i in range(n)]
        v, kind, dt = cpu_sample(cfg, fr, cores)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
               "sample": "%d frames of the same workload in %.1f s, one cv::StereoSGBM matcher per thread, %d threads, %s" % (n, dt, cores, cpu_model())}

    if rank == 0:
        line = {
            "metric": "disparity frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic", "config": workload(cfg, NF, lanes),
            "gpix_disp_per_s": fps * cfg.gpix_disp, "parity_checked_vs_golden_crc": checked,
            "parity_frames_checked": parity_checked, "parity_frames_ok": parity_ok, "parity_reference": ref_kind,
            "configs": side, "e2e_plugin": plugin,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "rectify": rect, "stereobm": bm,
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if parity_ok != parity_checked:
        sys.stderr.write("bench.py: PARITY FAILURE: %d of %d frames differ from %s\n" % (parity_checked - parity_ok, parity_checked, ref_kind))
        raise SystemExit(3)


def side_config(torch, dev, cfg, alu_peak, hbm_peak, frames=8, lanes=4, reps=4):
    """One of the other configurations on the same GPU: `frames` distinct pairs resident in HBM cycled over `lanes` lanes
    (the lanes' cost volumes exceed L2 for every config but c1, whose 4 x 71 MB of volumes still do), CUDA events around
    `reps` passes, stage events on one lane, every output compared with cv::StereoSGBM."""
    p = cfg.params
    W, H, D = cfg.width, cfg.height, p.numDisparities
    eng = b200sgm.Engine(dev.index or 0, W, H, D, lanes, p)
    try:
        pairs = [synth.make_pair(W, H, D, p.minDisparity, 1000 + i) for i in range(frames)]
        hL = torch.stack([torch.from_numpy(a) for a, _ in pairs]).pin_memory()
        hR = torch.stack([torch.from_numpy(b) for _, b in pairs]).pin_memory()
        dL, dR = hL.to(dev), hR.to(dev)
        dD = torch.empty((frames, H, W), dtype=torch.int16, device=dev)
        streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
        out = {"workload": workload(cfg, frames, lanes)["workload"], "lanes": lanes}

        def one_pass():
            for i in range(frames):
                ln = i % lanes
                eng.compute_device(ln, dL[i].data_ptr(), W, dR[i].data_ptr(), W, W, H, dD[i].data_ptr(), W * 2, stream=streams[ln].cuda_stream)

        if cfg.name == "c5":
            # c3 + processDisparity + reprojection through the host-buffer call b200sgm_compute_xyz (pinned buffers);
            # 120 MB per frame cross PCIe (2 x 5 MB in; disparity 10, dmat 20, depth 20, cloud <= 80 MB out)
            cam = b200sgm.C5_CAMERA
            from oracle import oracle
            q = oracle.calc_q(cam["fx"], cam["cx"], cam["cxr"], cam["cy"], cam["p14"])
            fT = np.float32(0.3 * 2400.0)
            min_disp = float(fT / np.float32(cam["depth_max"]))
            # page-locked result buffers, allocated once (what MatcherB200SGM::matchToCloud does through b200sgm_host_alloc)
            pin = [torch.empty((H, W), dtype=torch.int16).pin_memory(), torch.empty((H, W), dtype=torch.float32).pin_memory(),
                   torch.empty((H, W), dtype=torch.float32).pin_memory(), torch.empty((H * W, 4), dtype=torch.float32).pin_memory()]
            outbuf = tuple(t.numpy() for t in pin)
            eng.compute_xyz(pairs[0][0], pairs[0][1], q, cam["depth_min"], cam["depth_max"], min_disp, float("inf"), out=outbuf)
            t0 = time.perf_counter()
            n_pts = 0
            for i in range(frames):
                disp, dmat, depth, pts, n = eng.compute_xyz(hL[i].numpy(), hR[i].numpy(), q, cam["depth_min"], cam["depth_max"], min_disp,
                                                            float("inf"), out=outbuf)
                n_pts += n
            dt = (time.perf_counter() - t0) / frames
            by = 2 * W * H + W * H * (2 + 4 + 4) + 16 * (n_pts // frames)
            # pageable results for comparison (fresh numpy arrays per call: what a naive caller pays)
            t1 = time.perf_counter()
            for i in range(2):
                eng.compute_xyz(hL[i].numpy(), hR[i].numpy(), q, cam["depth_min"], cam["depth_max"], min_disp, float("inf"))
            dt_pageable = (time.perf_counter() - t1) / 2
            out.update({"api": "b200sgm_compute_xyz, page-locked host buffers in and out, synchronous, 1 lane", "ms_per_frame": dt * 1e3,
                        "frames_per_s": 1.0 / dt, "ms_per_frame_pageable_results": dt_pageable * 1e3,
                        "points_per_frame": n_pts // frames, "pcie_bytes_per_frame": by, "pcie_gbs": by / dt / 1e9,
                        "pcie_roof": "host link: ~55 GB/s per direction measured with page-locked buffers on this box class (PCIe 5 x16); "
                                     "the frame's %.0f MB alone take ~%.1f ms of it, the matcher ~4.1 ms, nothing overlaps in this synchronous call"
                                     % (by / 1e6, by / 55e9 * 1e3)})
            refs, kind = reference_outputs(cfg, pairs[-1:], 1)
            out["parity_frames_checked"], out["parity_frames_ok"], out["parity_reference"] = 1, count_equal([disp], refs), kind
            return out
        for _ in range(2):
            one_pass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        for _ in range(reps):
            one_pass()
        for st in streams:
            ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * frames)
        eng.profile(True); eng.stage_times(0)
        for i in range(frames):
            eng.compute_device(0, dL[i].data_ptr(), W, dR[i].data_ptr(), W, W, H, dD[i].data_ptr(), W * 2, stream=0)
        st_ms, nfr = eng.stage_times(0)
        eng.profile(False)
        stages = {k: v / max(nfr, 1) for k, v in st_ms.items()}
        frame_ms = sum(stages.values())
        W1 = p.w1(W)
        R = 8 if p.mode else 5
        alg_ops = W1 * H * D * (30 + 9 * R)
        alg_bytes = 4 * W * H if p.mode == 0 else 4 * W * H + 4 * W1 * H * D
        tops = alg_ops / (frame_ms * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(launch_summary_path(cfg.name))).get("frame_dram_bytes")
        except Exception:
            pass
        one_pass()
        torch.cuda.synchronize()
        refs, kind = reference_outputs(cfg, pairs, os.cpu_count() or 1)
        got = dD.cpu().numpy()
        # one frame on its own on a ONE-lane engine (what the plugin adapter creates): its sweeps take all SMs, narrow strips
        # use the halo producers; checked against the same reference
        eng1 = b200sgm.Engine(dev.index or 0, W, H, D, 1, p)
        try:
            d1 = torch.empty((H, W), dtype=torch.int16, device=dev)
            eng1.compute_device(0, dL[0].data_ptr(), W, dR[0].data_ptr(), W, W, H, d1.data_ptr(), W * 2, stream=0)
            torch.cuda.synchronize()
            eng1.profile(True); eng1.stage_times(0)
            for i in range(frames):
                eng1.compute_device(0, dL[i].data_ptr(), W, dR[i].data_ptr(), W, W, H, d1.data_ptr(), W * 2, stream=0)
            torch.cuda.synchronize()
            st1, n1 = eng1.stage_times(0)
            out["one_lane_engine_ms"] = sum(st1.values()) / max(n1, 1)
            out["one_lane_engine_ok"] = bool(np.array_equal(d1.cpu().numpy(), refs[frames - 1]))
        finally:
            eng1.close()
        out.update({"frames_per_s": 1e3 / ms, "ms_per_frame_pipelined": ms, "stage_ms_per_frame": stages, "single_lane_ms": frame_ms,
                    "roofline": {"bound": "alu", "achieved": tops, "peak": alu_peak, "unit": "Tops/s (elementary int16 ops)",
                                 "frac": tops / alu_peak if alu_peak else None, "traffic": traffic,
                                 "frac_pipelined": (alg_ops * (1e3 / ms) / 1e12 / alu_peak) if alu_peak else None,
                                 "hbm_frac_algorithmic": alg_bytes / (frame_ms * 1e-3) / 1e9 / hbm_peak,
                                 "hbm_frac_traffic": traffic / (frame_ms * 1e-3) / 1e9 / hbm_peak if traffic else None},
                    "parity_frames_checked": frames, "parity_frames_ok": count_equal([got[i] for i in range(frames)], refs),
                    "parity_reference": kind, "warning": eng.last_warning})
        return out
    finally:
        eng.close()


def plugin_call(cfg, hostL, hostR, refs, n_unique, frames=8):
    """The call the ROS node makes: MatcherB200SGM::setImages / match / getDisparity (host/matcherB200SGM.cpp) on pageable
    cv::Mat-like buffers, one frame at a time, one lane -- timed through the C++ harness loop (host/harness --bench), which runs
    init_matcher + updateMatcher + stereo_match of generate_disparity.cpp:241-368 around the adapter."""
    import importlib
    import tempfile
    b = importlib.import_module("i3dr_stereo_camera-ros_b200.build")
    harness = b.build_host()
    p = cfg.params
    W, H = cfg.width, cfg.height
    with tempfile.TemporaryDirectory() as td:
        lp, rp, op = os.path.join(td, "l.raw"), os.path.join(td, "r.raw"), os.path.join(td, "d.f32")
        hostL[0].numpy().tofile(lp); hostR[0].numpy().tofile(rp)
        args = [harness, lp, rp, W, H, op, p.minDisparity, p.numDisparities, p.blockSize, p.uniquenessRatio, p.speckleRange,
                p.speckleWindowSize, p.preFilterCap, p.P1, p.P2, p.mode, 0, 10, "--bench", frames]
        r = subprocess.run([str(a) for a in args], capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            return {"error": r.stderr[-400:]}
        ms = match_ms = None
        for line in r.stderr.splitlines():
            if line.startswith("bench_ms_per_frame"):
                ms = float(line.split()[1])
            if line.startswith("bench_match_ms_per_frame"):
                match_ms = float(line.split()[1])
        got = np.fromfile(op, np.float32).reshape(H, W)
        ok = bool(np.array_equal(got, refs[0].astype(np.float32)))
    return {"api": "MatcherB200SGM::setImages/match/getDisparity via host/harness (b200sgm_compute_f32, pageable buffers, 1 lane)",
            "ms_per_frame": ms, "frames_per_s": 1e3 / ms if ms else None, "match_ms_per_frame": match_ms,
            "frames": frames, "matches_reference": ok,
            "h2d_bytes_per_frame": 2 * W * H, "d2h_bytes_per_frame": 4 * W * H,
            "note": "ms_per_frame = the node's whole stereo_match() (generate_disparity.cpp:334-368: a fresh CV_32F Mat, setImages' two "
                    "cv::resize copies into new Mats, match, getDisparity's deep copy -- reference code around the plugin); "
                    "match_ms_per_frame = MatcherB200SGM::match() alone (H2D of the pageable images, the engine, the float result "
                    "into the adapter's page-locked buffer).  Compare with the reference's stereo_match() on cv::StereoSGBM: "
                    "cpu_baseline single-frame latency = cores / value"}


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    cfg = CONFIGS[args.config]
    # Libraries (NCCL's version banner, for one) write to fd 1: keep the original stdout for the JSON line only and send
    # everything else to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
