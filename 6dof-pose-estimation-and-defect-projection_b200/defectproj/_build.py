"""Builds libdefectproj.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
SO_PATH = os.path.join(_HERE, "libdefectproj.so")
SOURCES = ["api.cu", "compact.cu", "build.cu", "trace.cu", "depth.cu", "prep.cu", "icp.cu", "normals.cu", "peer.cu"]
HEADERS = ["dp_internal.cuh", os.path.join("..", "..", "include", "defectproj.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math=false"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def is_stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra=(), out=None) -> str:
    out = out or SO_PATH
    if not force and not is_stale() and out == SO_PATH:
        return SO_PATH
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + list(extra)
    if verbose:
        flags = flags + ["-Xptxas", "-v"]
    objdir = os.path.join(CSRC, "build" + ("_" + os.path.basename(out) if out != SO_PATH else ""))
    os.makedirs(objdir, exist_ok=True)

    def one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(one, SOURCES))
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs, "-Xcompiler", "-fPIC",
           "-cudart", "shared"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
