"""Builds fresnel_b200/csrc/libfresnel_b200.so (the C-ABI library) with nvcc for sm_100a.

The library is built in-tree so that it travels with the repository snapshot; there is no
JIT cache and no CPU fallback.  ``python -m fresnel_b200.build`` rebuilds it.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libfresnel_b200.so")
OBJ_DIR = os.path.join(CSRC, "build")

SOURCES = ["project.cu", "sort.cu", "tile_lists.cu", "composite.cu", "composite_phase.cu", "wave.cu", "asm.cu", "fourier.cu", "io.cu", "head.cu", "loss.cu", "simple.cu", "exchange.cu", "pipeline.cu"]
HEADERS = ["frb_math.h", "frb_head.h", "frb_common.cuh", "composite_common.cuh", os.path.join(ROOT, "include", "fresnel_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("fresnel_b200: nvcc not found; the CUDA library cannot be built")


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    srcs = sources()
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    stamp = os.path.join(OBJ_DIR, "stamp")
    want = _digest([os.path.join(CSRC, s) for s in srcs] + hdrs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    # --no-undefined: a function declared in the header but lost from the sources must fail the build, not the first load
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcufft", "-Xlinker", "-rpath,/usr/local/cuda/lib64",
           "-Xlinker", "--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
