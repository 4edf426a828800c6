"""Builds libetpgt_b200.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python gat-recommendation_b200/build.py [--force] [--verbose]

One object per .cu (compiled in parallel), linked into
gat-recommendation_b200/etpgt_b200/libetpgt_b200.so.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.  No torch dependency: the library is plain CUDA runtime.
"""

from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT_DIR = HERE / "etpgt_b200"
BUILD = HERE / "build"
LIB = OUT_DIR / "libetpgt_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (Path(cand).exists() or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def source_digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for dep in [src, *sorted(CSRC.glob("*.cuh")), HERE.parent / "include" / "etpgt_b200.h"]:
        h.update(dep.read_bytes())
    return h.hexdigest()


def compile_one(src: Path, force: bool, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".sha")
    digest = source_digest(src)
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return obj
    cmd = [nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    (BUILD / (src.stem + ".ptxas.log")).write_text(proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        sys.stderr.write(proc.stderr)
    stamp.write_text(digest)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        objs = list(pool.map(lambda s: compile_one(s, force, verbose), sources))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-lcudart"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    print(build(args.force, args.verbose))
