"""Builds the CUDA engine in-tree: csrc/*.cu -> libb200sgm.so (sm_100a only, no torch dependency).

The shared library is the C-ABI drop-in boundary declared in include/b200sgm.h.  It is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development aid: B200SGM_VARIANT=<name> B200SGM_CFLAGS="-D..." builds libb200sgm_<name>.so beside the product library
# (objects under csrc/_obj_<name>); B200SGM_LIB=<path> makes engine.py load it instead
VARIANT = os.environ.get("B200SGM_VARIANT", "")
LIB = os.path.join(HERE, "libb200sgm%s.so" % ("_" + VARIANT if VARIANT else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "--use_fast_math"]
# --use_fast_math only affects float intrinsics the engine does not use on the parity path; the reprojection
# kernel uses explicit __fmul_rn/__fadd_rn/__fdiv_rn.  Keep it off to be safe:
FLAGS.remove("--use_fast_math")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


OBJ = os.path.join(CSRC, "_obj" + ("_" + VARIANT if VARIANT else ""))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps(src: str, seen=None):
    """src plus every header it includes (quoted includes, resolved relative to the including file), transitively."""
    import re
    seen = set() if seen is None else seen
    src = os.path.normpath(src)
    if src in seen or not os.path.exists(src):
        return seen
    seen.add(src)
    for inc in re.findall(r'^\s*#include\s+"([^"]+)"', open(src).read(), flags=re.M):
        _deps(os.path.join(os.path.dirname(src), inc), seen)
    return seen


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every csrc/*.cu to its own object (in parallel, only the stale ones) and links libb200sgm.so."""
    os.makedirs(OBJ, exist_ok=True)
    cflags = [f for f in FLAGS if f != "-shared"] + os.environ.get("B200SGM_CFLAGS", "").split()
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        only = os.environ.get("B200SGM_ONLY")   # development aid: rebuild just these translation units (comma separated)
        if only and os.path.exists(obj) and os.path.basename(src)[:-3] not in only.split(","):
            continue
        if force or _stale(obj, _deps(src)):
            cmd = [NVCC] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            print("[b200sgm] " + " ".join(cmd), file=sys.stderr)
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if procs or force or _stale(LIB, objs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
        print("[b200sgm] " + " ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


HOST = os.path.join(HERE, "host")
HARNESS = os.path.join(HOST, "harness")


def build_host(force: bool = False) -> str:
    """Builds the ROS/OpenCV-free harness around the C++ matcher adapter (host/matcherB200SGM.cpp)."""
    srcs = [os.path.join(HOST, "harness.cpp"), os.path.join(HOST, "matcherB200SGM.cpp"), os.path.join(HOST, "matcherB200BM.cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in ("matcherB200SGM.h", "matcherB200BM.h", "matcher_interface.h", "cv_stub.h")] + [LIB]
    if not force and os.path.exists(HARNESS) and all(os.path.getmtime(d) <= os.path.getmtime(HARNESS) for d in deps):
        return HARNESS
    cmd = ["g++", "-std=c++17", "-O2", "-DB200SGM_STANDALONE", "-I" + os.path.join(HERE, "..", "include"), "-I" + HOST,
           "-o", HARNESS] + srcs + ["-L" + HERE, "-lb200sgm", "-Wl,-rpath,$ORIGIN/.."]
    print("[b200sgm] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return HARNESS


STREAM_BENCH = os.path.join(HOST, "stream_bench")


def build_stream(force: bool = False) -> str:
    """Builds the C++ frame-stream driver (host/stream_driver.cpp) with its single-process multi-GPU bench (host/stream_bench.cpp)."""
    srcs = [os.path.join(HOST, "stream_bench.cpp"), os.path.join(HOST, "stream_driver.cpp")]
    deps = srcs + [os.path.join(HOST, "stream_driver.h"), LIB]
    if not force and os.path.exists(STREAM_BENCH) and all(os.path.getmtime(d) <= os.path.getmtime(STREAM_BENCH) for d in deps):
        return STREAM_BENCH
    cmd = ["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(HERE, "..", "include"), "-I" + HOST, "-o", STREAM_BENCH] + srcs + \
          ["-L" + HERE, "-lb200sgm", "-Wl,-rpath,$ORIGIN/.."]
    print("[b200sgm] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return STREAM_BENCH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
    build_stream(force="--force" in sys.argv)
    print(LIB)
