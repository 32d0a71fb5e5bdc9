"""Build ``libqe_b200.so`` in-tree for sm_100a (``python -m dist_classicrl_b200.build``)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "qe_engine.cu")
SRC2 = os.path.join(PKG, "csrc", "qe_shard.cu")  # the peer-memory sharded table: its own translation unit (compiled in parallel)
DEPS = [SRC, SRC2, os.path.join(PKG, "csrc", "qe_small.cuh"), os.path.join(PKG, "csrc", "qe_shard.cuh"), os.path.join(PKG, "csrc", "qe_pipe.cuh"), os.path.join(PKG, "csrc", "qe_flow.cuh"), os.path.join(PKG, "csrc", "qe_kernels.cuh"), os.path.join(PKG, "csrc", "qe_common.cuh"), os.path.join(PKG, "csrc", "qe_sorted.cuh"), os.path.join(PKG, "csrc", "qe_radix.cuh"),
        os.path.join(os.path.dirname(PKG), "include", "qe_engine.h")]
OUT = os.path.join(PKG, "_lib", "libqe_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--threads", "2",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libqe_b200.so")
    return exe


def up_to_date() -> bool:
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc(), *NVCC_FLAGS, "-o", OUT, SRC, SRC2]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(PKG, "_lib", "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}); see {log}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
