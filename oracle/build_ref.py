"""Build the reference's own CUDA operator as a second oracle and as the GPU baseline-to-beat.

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

    python -m oracle.build_ref            # -> oracle/_ref/libsimilarity_ref.so

Compiles, with the container's nvcc and nothing else,
    /root/reference/GAN-Based-SR/basicsr/losses/similarity/similarity.cu   (unmodified, read in place)
    oracle/ref_shim.cu                                                     (extern "C" names + error code)
for sm_100 (plus compute_100 PTX), the way the reference's import-time JIT would build it on a B200
box (similaritywrapper.py:15-23 passes no arch or optimisation flags; nvcc's default device
optimisation level is already -O3).
The reference sources are never copied into this repository: only the built library lands in
oracle/_ref/ (git-ignored, but it travels to the GPU box with the snapshot, where /root/reference
does not exist).  If /root/reference is absent the prebuilt library is used as is.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/GAN-Based-SR/basicsr/losses/similarity"
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libsimilarity_ref.so")


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build(force: bool = False) -> str | None:
    """Returns the library path, or None when it can neither be built nor found."""
    src = os.path.join(REF_DIR, "similarity.cu")
    shim = os.path.join(HERE, "ref_shim.cu")
    if not os.path.exists(src):
        return LIB if os.path.exists(LIB) else None          # GPU box: use what travelled
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(src),
                                                                          os.path.getmtime(shim)):
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        return LIB if os.path.exists(LIB) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100,code=sm_100", "-gencode", "arch=compute_100,code=compute_100",
           "-shared", "-Xcompiler", "-fPIC", "-I", REF_DIR, "-o", LIB, src, shim]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building the reference similarity.cu")
    return LIB


_lib = None


def load():
    """ctypes handle with typed entry points, or None if the library is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = build()
    if path is None or not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    p, i = ctypes.c_void_p, ctypes.c_int
    lib.ref_compute_similarity.restype = i
    lib.ref_compute_similarity.argtypes = [p, p, p, i, i, i, i, i, i]
    lib.ref_compute_similarity_backward.restype = i
    lib.ref_compute_similarity_backward.argtypes = [p, p, p, p, i, i, i, i, i, i]
    _lib = lib
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
