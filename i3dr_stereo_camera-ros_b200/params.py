"""Parameter set of the SGBM path and the BASELINE.json configurations.

The fields carry RAW values exactly as the reference forwards them to cv::StereoSGBM
(/root/reference/src/generate_disparity.cpp:245-256 -> src/stereoMatcher/matcherOpenCVSGBM.cpp:53-110);
the engine applies OpenCV's defaulting rules itself.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, asdict, replace

MODE_SGBM = 0
MODE_HH = 1


class CParams(ctypes.Structure):
    """Mirror of `b200sgm_params` in include/b200sgm.h (and of sgbm_oracle_params)."""
    _fields_ = [(n, ctypes.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


@dataclass(frozen=True)
class SGBMParams:
    # defaults = cfg/i3DR_Disparity.cfg:21-39 with the window overridden to 9 (BASELINE config 1)
    minDisparity: int = 0
    numDisparities: int = 64
    blockSize: int = 9
    P1: int = 200
    P2: int = 400
    disp12MaxDiff: int = 0
    preFilterCap: int = 31
    uniquenessRatio: int = 15
    speckleWindowSize: int = 100
    speckleRange: int = 4
    mode: int = MODE_SGBM

    def to_c(self) -> CParams:
        return CParams(**asdict(self))

    def replace(self, **kw) -> "SGBMParams":
        return replace(self, **kw)

    # --- derived geometry (SURVEY.md section 8 notation) ---
    def min_x1(self) -> int:
        return max(self.minDisparity + self.numDisparities, 0)

    def w1(self, width: int) -> int:
        return width + min(self.minDisparity, 0) - self.min_x1()

    def invalid(self) -> int:
        return (self.minDisparity - 1) * 16


@dataclass(frozen=True)
class Config:
    name: str
    width: int
    height: int
    params: SGBMParams
    note: str = ""

    @property
    def gpix_disp(self) -> float:
        return self.width * self.height * self.params.numDisparities / 1e9


# BASELINE.json `configs` (SURVEY.md section 8d "Parameters per config")
CONFIGS = {
    "c1": Config("c1", 640, 480, SGBMParams(), "640x480x64 MODE_SGBM cfg defaults, bs 9"),
    "c2": Config("c2", 1280, 1024, SGBMParams(numDisparities=128, mode=MODE_HH, uniquenessRatio=10, disp12MaxDiff=1),
                 "1280x1024x128 MODE_HH uniq 10 d12 1"),
    "c3": Config("c3", 2448, 2048, SGBMParams(numDisparities=256, speckleWindowSize=100, speckleRange=2),
                 "2448x2048x256 MODE_SGBM speckle 100/2"),
}
CONFIGS["c4"] = Config("c4", 2448, 2048, CONFIGS["c3"].params, "stream of c3 frames sharded over GPUs")
CONFIGS["c5"] = Config("c5", 2448, 2048, CONFIGS["c3"].params, "c3 + processDisparity + reprojection to XYZ")
# cL: what `roslaunch ... stereo_algorithm:=1` installs (/root/reference/launch/stereo_matcher.launch:37-48) on the camera's
# default frame (launch/stereo_capture.launch:14-15).  blockSize 21 with preFilterCap 7 sits in the regime where the int16
# cost bound bs^2 * (2*ftzero + 63) = 41 013 exceeds 32 767 on adversarial data.
CONFIGS["cL"] = Config("cL", 2448, 2048,
                       SGBMParams(minDisparity=147, numDisparities=480, blockSize=21, uniquenessRatio=2, speckleWindowSize=1000,
                                  speckleRange=4, preFilterCap=7, P1=200, P2=400),
                       "2448x2048, minD 147, 480 disparities, window 21: the reference's launch-file default for SGBM")

# c5 camera model: fx=f=2400, cx=cxr=1224, cy=1024, baseline 0.3 m
C5_CAMERA = dict(fx=2400.0, cx=1224.0, cxr=1224.0, cy=1024.0, p14=-2400.0 * 0.3, depth_min=0.0, depth_max=10.0)
