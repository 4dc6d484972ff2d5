"""In-tree build of libmrx.so (the C-ABI library declared in include/mrx.h).

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to
the GPU box with the gpurun snapshot.  `python -m matterport_maskrcnn_with_tensorflow_serving_b200.build`
or `__graft_entry__.build()` runs this.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmrx.so")
STAMP_PATH = os.path.join(LIB_DIR, "libmrx.stamp")

SOURCES = ["capi.cu", "anchors.cu", "unmold.cu", "expand_team.cu", "expand_ws4.cu", "mold.cu", "composite.cu", "pack.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "expand.cuh"),
           os.path.join(os.path.dirname(PKG_DIR), "include", "mrx.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--cudart", "shared",          # share torch's CUDA runtime instance (device/stream state)
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libmrx.so cannot be built")
    return exe


def _fingerprint():
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(path, "rb") as f:
            h.update(f.read())
    h.update((" ".join(NVCC_FLAGS) + os.environ.get("MRX_NVCC_FLAGS", "")).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile csrc/*.cu into lib/libmrx.so (skipped when sources are unchanged)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("MRX_NVCC_FLAGS", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmrx.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    with open(STAMP_PATH, "w") as f:
        f.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
