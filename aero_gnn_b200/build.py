"""Build libaero_sm100.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The library has a plain C ABI (include/aero_gnn.h) and links only against the CUDA runtime, so it
is compiled by invoking nvcc directly; objects are rebuilt when a source or header is newer.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
BUILD = PKG / "build"
LIB = PKG / "libaero_sm100.so"
PROBE_LIB = PKG / "libaero_probe.so"     # hardware probes (include/aero_gnn_debug.h): diagnostics, not in the product library
PROBE_SOURCES = ["umma_probe.cu"]

SOURCES = ["abi.cu", "sort_plan.cu", "segment.cu", "block_simt.cu", "block_umma.cu", "block_umma_bwd.cu", "block_umma_bwd2.cu",
           "bistride.cu", "train_tail.cu", "wgrad.cu", "rowgemm.cu", "tma.cu", "thin_linear.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xptxas", "-v",
] + os.environ.get("AERO_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newer(a: Path, b: Path) -> bool:
    return (not b.exists()) or a.stat().st_mtime > b.stat().st_mtime


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    jobs = []
    for s in SOURCES + PROBE_SOURCES:
        src = CSRC / s
        obj = BUILD / (s + ".o")
        stale = force or _newer(src, obj) or any(_newer(h, obj) for h in headers)
        if stale:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src.name + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(compile_one, jobs))
    objs = [BUILD / (s + ".o") for s in SOURCES]
    if jobs or not LIB.exists():
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "shared"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if jobs or not PROBE_LIB.exists():
        # the probes use the product library's error plumbing: link against it, found next to the probe library
        cmd = [_nvcc(), "-shared", "-o", str(PROBE_LIB), *[str(BUILD / (s + ".o")) for s in PROBE_SOURCES],
               "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", f"-L{PKG}", "-l:libaero_sm100.so",
               "-Xlinker", "-rpath=$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"probe link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
