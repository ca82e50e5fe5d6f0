"""Build libhippo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hippomm_b200.build [--force]

The shared library lands next to this file so that it travels with the source tree; it is
git-ignored.  Only the CUDA runtime is linked (statically); `cuTensorMapEncodeTiled` is
resolved from the driver at run time, so the library also loads on a machine without
libcuda (the symbol-export test relies on that).
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = PKG_DIR / "csrc" / "_obj"
LIB_PATH = PKG_DIR / "libhippo_b200.so"

SOURCES = [
    "lib.cu",
    "bank_build.cu",
    "topk_single.cu",
    "topk_batched.cu",
    "topk_rows.cu",
    "topk_small.cu",
    "recall.cu",
    "exchange.cu",
    "sim_tc.cu",
    "consolidate.cu",
    "frames.cu",
    "audio.cu",
    "segment.cu",
    "pattern.cu",
]
# per-file extra flags; the boundary state machine must not contract a*b+c into an FMA
EXTRA_FLAGS = {"segment.cu": ["-fmad=false"]}

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; set NVCC or put it on PATH")
    return cand


def _deps(src: Path) -> list[Path]:
    return [src, *CSRC.glob("*.cuh"), PKG_DIR.parent / "include" / "hippo_b200.h"]


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def _compile(src_name: str, force: bool, verbose: bool) -> Path:
    src = CSRC / src_name
    obj = OBJ_DIR / (src.stem + ".o")
    if not force and not _stale(obj, _deps(src)):
        return obj
    cmd = [_nvcc(), *ARCH_FLAGS, *COMMON_FLAGS, *EXTRA_FLAGS.get(src_name, []), "-c", str(src), "-o", str(obj)]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src_name}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link libhippo_b200.so. Returns its path."""
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    if force or _stale(LIB_PATH, objs):
        tmp = LIB_PATH.with_suffix(".so.tmp")
        cmd = [_nvcc(), *ARCH_FLAGS, "-shared", "-o", str(tmp), *map(str, objs)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print(f"built {p} ({p.stat().st_size} bytes)")
