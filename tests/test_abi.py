"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/b200sgm.h declares; parameter plumbing; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

import b200sgm
from b200sgm import SGBMParams, CONFIGS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import importlib
    importlib.import_module("i3dr_stereo_camera-ros_b200.build").build()
    return b200sgm.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "b200sgm.h")).read()
    declared = set(re.findall(r"\b(b200sgm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(b200sgm.SYMBOLS), declared ^ set(b200sgm.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), "missing export " + s


def test_version_and_null_handle(lib):
    assert b"sm_100a" in lib.b200sgm_version()
    assert lib.b200sgm_set_params(None, None) == -1
    assert lib.b200sgm_destroy(None) == -1


def test_params_struct_layout_matches_header():
    hdr = open(os.path.join(ROOT, "include", "b200sgm.h")).read()
    body = re.search(r"typedef struct b200sgm_params \{(.*?)\} b200sgm_params;", hdr, re.S).group(1)
    fields = re.findall(r"int\s+(\w+);", body)
    assert fields == [n for n, _ in b200sgm.CParams._fields_]
    assert ctypes.sizeof(b200sgm.CParams) == 4 * len(fields)


def test_geometry_of_baseline_configs():
    # c1 W1=576, c2 W1=1152, c3 W1=2192
    assert CONFIGS["c1"].params.w1(640) == 576
    assert CONFIGS["c2"].params.w1(1280) == 1152
    assert CONFIGS["c3"].params.w1(2448) == 2192
    assert SGBMParams(minDisparity=-20, numDisparities=16).w1(110) == 90
    assert SGBMParams(minDisparity=9).invalid() == 128


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_engine_fails_loudly_without_gpu(lib):
    with pytest.raises(b200sgm.B200SGMError):
        b200sgm.Engine(0, 64, 64, 16, 1, SGBMParams(numDisparities=16))
