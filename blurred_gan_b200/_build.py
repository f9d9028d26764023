"""Builds ``libbgan_b200.so`` in-tree with nvcc for sm_100a (and nothing else).

The shared library is git-ignored but travels to the GPU box with the snapshot; there is no JIT
cache and no pip install.  ``python -m blurred_gan_b200._build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libbgan_b200.so")
SOURCES = ["capi.cu", "blur.cu", "blur_tc.cu", "blur_direct.cu", "blur_axis.cu", "swd.cu", "project_tc.cu", "sort.cu", "sort_sample.cu", "hostrng.cu"]
# host-only sources compiled by g++ (name -> extra flags); selected at run time by a CPU feature check in hostrng.cu
CPP_SOURCES = {"hostrng_avx2.cpp": ["-mavx2"], "hostrng_avx512.cpp": ["-mavx512f"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(REPO_DIR, "include"),
    "-I", CSRC,
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libbgan_b200.so")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(REPO_DIR, "include", "bgan.h"))

    def compile_one(src: str) -> str:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
            if verbose:
                sys.stderr.write(res.stderr)
        return obj

    def compile_cpp(src: str) -> str:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cpp", ".o"))
        if force or _stale(obj, [path]):
            gxx = shutil.which("g++") or "g++"
            cmd = [gxx, "-O3", "-std=c++17", "-fPIC", *CPP_SOURCES[src], "-c", path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("g++ failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    objs += [compile_cpp(src) for src in CPP_SOURCES]
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB_PATH, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
