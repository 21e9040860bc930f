"""Build ginfinity_b200/libgfx.so in-tree with nvcc for sm_100a.

    python -m ginfinity_b200.build_native [--force]

One translation unit per .cu file, compiled in parallel, linked into a single
shared library next to this file (git-ignored; it travels to the GPU box with
the working tree).  `-lineinfo` keeps ncu's source page usable.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
INCLUDE = HERE.parent / "include"
BUILD = HERE / "csrc" / "build"
LIB = HERE / "libgfx.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v", "--expt-relaxed-constexpr", f"-I{INCLUDE}", f"-I{CSRC}"]


def _fingerprint() -> str:
    h = hashlib.sha256()
    for path in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh"))
                       + [INCLUDE / "gfx.h", Path(__file__)]):
        h.update(path.name.encode())
        h.update(path.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = BUILD / "fingerprint.txt"
    want = _fingerprint()
    if (not force and LIB.is_file() and stamp.is_file()
            and stamp.read_text().strip() == want):
        return LIB
    BUILD.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))

    def compile_one(src: Path) -> Path:
        obj = BUILD / (src.stem + ".o")
        cmd = [NVCC, *ARCH, *FLAGS, "-c", str(src), "-o", str(obj)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src.stem + ".ptxas.log")).write_text(proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{proc.stderr}")
        if verbose:
            sys.stderr.write(proc.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        objects = list(pool.map(compile_one, sources))
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objects),
           "-lcudart"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed:\n{proc.stderr}")
    stamp.write_text(want + "\n")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
