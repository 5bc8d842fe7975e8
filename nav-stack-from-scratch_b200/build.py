"""In-tree build of the native libraries (no JIT cache: the .so files must travel with the repo snapshot).

  libnsfs_gpu.so       csrc/*.cu                 nvcc, sm_100a only        -> the C-ABI of include/nsfs.h
  libnsfs_planners.so  host/*.cpp                g++ -std=c++17            -> GridPlannerCore children + C shim

Run as ``python nav-stack-from-scratch_b200/build.py`` or through ``__graft_entry__.build()``.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
GPU_SO = os.path.join(PKG, "libnsfs_gpu.so")
HOST_SO = os.path.join(PKG, "libnsfs_planners.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-shared", "-cudart", "static",
]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc_path() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def build_gpu(force: bool = False, verbose: bool = False, extra=()) -> str:
    srcs = sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))
    deps = srcs + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + [os.path.join(ROOT, "include", "nsfs.h")]
    if force or _newer(GPU_SO, deps):
        cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
               "-o", GPU_SO, *srcs]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return GPU_SO


def build_host(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(PKG, "host", "*.cpp")))
    if not srcs:
        return ""
    deps = srcs + glob.glob(os.path.join(PKG, "host", "*.hpp")) + [os.path.join(ROOT, "include", "nsfs.h"), GPU_SO]
    if force or _newer(HOST_SO, deps):
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"),
               "-I", os.path.join(PKG, "host"), "-o", HOST_SO, *srcs, "-L", PKG, "-lnsfs_gpu", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return HOST_SO


def build_all(force: bool = False, verbose: bool = False):
    return build_gpu(force, verbose), build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", GPU_SO)
