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


def _run(cmd):
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        return res.returncode != 0, res.stdout + res.stderr
    except OSError as e:   # nvcc not on this machine
        return True, str(e)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every CUDA source into gym_fishing_b200/libfishing_b200.so (one nvcc per source, in parallel, then one
    link).  `extra_flags` / `out` build a variant of the same library for A/B runs (see _native.FISHING_B200_LIB).
    Returns the library path."""
    out = out or OUT
    if not force and out == OUT and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(os.path.dirname(_HERE), "build", "obj_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + (["-Xptxas=-v"] if verbose else [])
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    cmds = [[_nvcc(), *flags, "-c", os.path.join(CSRC, s), "-o", o] for s, o in zip(SOURCES, objs)]
    with ThreadPoolExecutor(len(cmds)) as ex:
        results = list(ex.map(_run, cmds))
    failed = any(f for f, _ in results)
    log = "".join(l for _, l in results)
    cmd = cmds[0]
    if not failed:
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out + ".tmp", *objs]
        failed, l = _run(cmd)
        log += l
    if failed:
        if os.path.exists(out) and not force:
            # e.g. a box without nvcc, or file times scrambled by a copy: keep the library that travelled with the tree
            print("gym_fishing_b200.build: rebuild failed, keeping the existing library\n" + log[-400:])
            return out
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + log[-4000:])
    os.replace(out + ".tmp", out)
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
