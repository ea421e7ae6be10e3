"""Builds libhmz.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python muzero-hanoi_b200/build.py [--force] [--verbose]

One object per csrc/*.cu (compiled in parallel, rebuilt only when the source or a header is
newer), linked into muzero-hanoi_b200/libhmz.so.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot; nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO, "include")
# HMZ_VARIANT=name (tooling): an A/B build with HMZ_NVCC_EXTRA's flags into variants/libhmz_<name>.so (own object
# directory), loaded with HMZ_LIB_PATH; the shipped library is always muzero-hanoi_b200/libhmz.so built without either.
VARIANT = os.environ.get("HMZ_VARIANT", "")
OBJ_DIR = os.path.join(PKG_DIR, "build" + ("_" + VARIANT if VARIANT else ""))
LIB_PATH = os.path.join(PKG_DIR, "variants", f"libhmz_{VARIANT}.so") if VARIANT else os.path.join(PKG_DIR, "libhmz.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA = os.environ.get("HMZ_NVCC_EXTRA", "").split() + (["-DHMZ_VARIANT"] if os.environ.get("HMZ_VARIANT") else [])
NVCC_FLAGS = [*EXTRA, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
              "-I", INCLUDE, "-I", CSRC]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _newest_header() -> float:
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(
        os.path.join(INCLUDE, "*.h")) + [os.path.abspath(__file__)]
    return max(os.path.getmtime(h) for h in hs)


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log = p.stdout + p.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = nvcc_path()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if not sources:
        raise RuntimeError("no CUDA sources under " + CSRC)
    hdr_time = _newest_header()
    jobs, objs = [], []
    for src in sources:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs))
    need_link = bool(jobs) or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)
    if need_link:
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
