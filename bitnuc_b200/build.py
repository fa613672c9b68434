"""Builds libbitnuc_cuda.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libbitnuc_cuda.so"
SOURCES = ["api.cu", "codec.cu", "kmer.cu", "hamming.cu", "counts.cu", "batch.cu", "split.cu", "gather.cu", "windows.cu", "fastq.cu", "synth.cu", "multi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    *os.environ.get("BN_NVCC_EXTRA", "").split(),   # experiments only (-DNAME=value knobs of a kernel under test)
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libbitnuc_cuda.so cannot be built")
    return exe


def _stale() -> bool:
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*.cu*")) + list(CSRC.glob("*.h")) +
                 [PKG.parent / "include" / "bitnuc_cuda.h"])
    return LIB.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    if not force and not _stale():
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    exe = nvcc()

    def compile_one(src: str) -> Path:
        obj = objdir / (src + ".o")
        cmd = [exe, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    tmp = LIB.with_suffix(".so.tmp")
    r = subprocess.run([exe, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
                        *map(str, objs)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    tmp.replace(LIB)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
