"""In-tree build of libslamcu.so (hand-written CUDA for sm_100a behind the C ABI in include/slam/cuda/slamcu.h).

nvcc cross-compiles without a GPU.  -fmad=false: the reference arithmetic is never FMA-contracted
(baseline x86-64), explicit fmaf() calls remain where OpenCV uses them.  -lineinfo for ncu source pages.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libslamcu.so")
SOURCES = ["api.cu", "fast_ref.cu", "sortnms.cu", "describe_ref.cu", "match.cu", "match_tc.cu", "prep.cu", "ransac.cu", "microbench.cu", "orb.cu", "essential.cu", "pack.cu"]
# Exactness-critical: -fmad=false and NO --use_fast_math (every bit-exactness guarantee depends on them; CMakeLists.txt
# passes the same pair, tests/test_cabi_and_host.py checks both recipes).
EXACT_FLAGS = ["-fmad=false"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", *EXACT_FLAGS, "-Xcompiler", "-fPIC"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "slam", "cuda", "slamcu.h"))
    return _newest(deps) > os.path.getmtime(OUT)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", OUT, *objs, "-lcudart"])
    return OUT


def build_debug() -> str:
    """libslamcu_dbg.so: the same sources with -DSLAMCU_DEBUG_BOUNDS (every list / tile / gather index is checked and traps);
    used by tests/test_gpu_debug_bounds.py in place of compute-sanitizer."""
    out = os.path.join(HERE, "libslamcu_dbg.so")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "slam", "cuda", "slamcu.h")]
    if os.path.exists(out) and _newest(deps) <= os.path.getmtime(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.join(HERE, "build", "dbg"), exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", "dbg", src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, "-DSLAMCU_DEBUG_BOUNDS", "-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} (debug-bounds build):\n{o}")
    subprocess.check_call([nvcc, "-shared", "-o", out, *objs, "-lcudart"])
    return out


def build_host_tools(verbose: bool = False) -> list:
    """Host C++ programs above the C ABI: tools/cli/slam_bench and test/frontend/test_frontend_cuda (g++, no CUDA
    headers needed; they link libslamcu.so with an rpath to the package directory)."""
    root = os.path.normpath(os.path.join(HERE, ".."))
    outdir = os.path.join(HERE, "build")
    os.makedirs(outdir, exist_ok=True)
    outs = []
    for src, name in ((os.path.join(root, "tools", "cli", "slam_bench.cpp"), "slam_bench"),
                      (os.path.join(root, "test", "frontend", "test_frontend_cuda.cpp"), "test_frontend_cuda")):
        out = os.path.join(outdir, name)
        deps = [src, os.path.join(root, "include", "slam", "cuda", "frontend.hpp"), os.path.join(root, "include", "slam", "cuda", "slamcu.h"),
                os.path.join(root, "include", "slam", "cuda", "yaml_lite.hpp")]
        if not os.path.exists(out) or _newest(deps) > os.path.getmtime(out):
            cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(root, "include"), src, "-o", out, "-L", HERE, "-lslamcu",
                   "-Wl,-rpath," + HERE, "-Wl,-rpath,/usr/local/cuda/lib64"]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        outs.append(out)
    return outs


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_tools(verbose="-v" in sys.argv))
    print(build_debug())
