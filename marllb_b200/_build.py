"""Build recipe for the in-tree CUDA library (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libmarllb_b200.so")
SOURCES = ["mlb_api.cu", "mlb_ops.cu", "mlb_policy.cu", "mlb_linear_tc.cu", "mlb_gemm_tc.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the marllb_b200 CUDA library cannot be built")


HASH_PATH = LIB_PATH + ".hash"


def _source_hash() -> str:
    """sha256 over every source the library is built from + the flags.  Content, not mtimes: the built .so
    travels to GPU boxes in a snapshot whose timestamps mean nothing."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "marllb_b200.h"))
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "marllb_b200_policy.h"))
    for p in deps:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile marllb_b200/libmarllb_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl
    # one builder at a time (torchrun starts N ranks at once); the others wait, then find a fresh library
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            want = _source_hash()
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
            cmd = [_nvcc(), *NVCC_FLAGS, "-o", tmp, *srcs]
            if os.path.exists("/usr/bin/g++"):
                cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.run(cmd, check=True, cwd=CSRC)
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH + ".tmp", "w") as f:
                f.write(want + "\n")
            os.replace(HASH_PATH + ".tmp", HASH_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
