"""In-tree build of libfishing_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ("abi.cu", "step_kernels.cu", "rollout_kernels.cu")
OUT = os.path.join(_HERE, "libfishing_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(_HERE), "include", "fishing_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source into gym_fishing_b200/libfishing_b200.so.  Returns the library path."""
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", OUT + ".tmp", *(os.path.join(CSRC, s) for s in SOURCES)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        failed = res.returncode != 0
        log = res.stdout + res.stderr
    except OSError as e:   # nvcc not on this machine
        failed, log = True, str(e)
    if failed:
        if os.path.exists(OUT) and not force:
            # e.g. a box without nvcc, or file times scrambled by a copy: keep the library that travelled with the tree
            print("gym_fishing_b200.build: rebuild failed, keeping the existing library\n" + log[-400:])
            return OUT
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + log)
    os.replace(OUT + ".tmp", OUT)
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
