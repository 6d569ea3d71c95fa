"""Ahead-of-time build of libssl_b200.so for sm_100a (no JIT, no torch headers).

    python -m ssl_b200.csrc.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/ssl_b200.h); it stays in-tree next to the
sources so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libssl_b200.so")
SOURCES = ["ssl_b200.cu"]
HEADERS = ["common.cuh", "edge_list.cuh", "row_ops.cuh", "ssg_point.cuh", "plane_geom.cuh", "plane_host.cuh",
           "ssg_plane_fwd.cuh", "ssg_plane_bwd.cuh", "row_loss_t.cuh", "pad.cuh", "tma.cuh", "pool_ops.cuh",
           os.path.join(ROOT, "include", "ssl_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libssl_b200.so cannot be built")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return all(not os.path.exists(d) or os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", HERE,
           "-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libssl_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
