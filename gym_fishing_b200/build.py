"""In-tree build of libfishing_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import hashlib
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


_ID_MARK = b"FISHING_BUILD_ID="


def source_id(extra_flags=()):
    """16 hex digits over everything the library is compiled from (csrc/*, the public header, the nvcc flags).  The same
    string is baked into the binary (fishing_build_id()), so a library can be checked against the tree it ships with."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(os.path.dirname(_HERE), "include", "fishing_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(list(NVCC_FLAGS) + list(extra_flags)).encode())
    return h.hexdigest()[:16]


def library_id(path=None):
    """The build id baked into a built library (read from the file, without loading it); None if absent."""
    path = path or OUT
    try:
        with open(path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    k = blob.find(_ID_MARK)
    return blob[k + len(_ID_MARK):k + len(_ID_MARK) + 16].decode("ascii", "replace") if k >= 0 else None


def needs_build():
    """True when there is no library or it was compiled from other sources than the tree holds (by content, not by file
    time: a copy to another machine scrambles the times)."""
    return library_id(OUT) != source_id()


def _run(cmd):
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        return res.returncode != 0, res.stdout + res.stderr
    except OSError as e:   # nvcc not on this machine
        return True, str(e)


def build(force=False, verbose=False, extra_flags=(), out=None, sources=SOURCES):
    """Compile every CUDA source into gym_fishing_b200/libfishing_b200.so (one nvcc per source, in parallel, then one
    link).  `extra_flags` / `out` build a variant of the same library for A/B runs (see _native.FISHING_B200_LIB).
    Returns the library path."""
    out = out or OUT
    if not force and out == OUT and not needs_build():
        return OUT
    import fcntl
    from concurrent.futures import ThreadPoolExecutor

    # one builder at a time (torchrun starts eight ranks that all find the same stale library); whoever waited re-checks
    lock = open(out + ".lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and out == OUT and not needs_build():
            return OUT
        return _build_locked(verbose, extra_flags, out, sources)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose, extra_flags, out, sources):
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(os.path.dirname(_HERE), "build", "obj_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    flags = ([f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + (["-Xptxas=-v"] if verbose else []) +
             ['-DFISHING_BUILD_ID_STR="' + source_id(extra_flags) + '"'])
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in sources]
    cmds = [[_nvcc(), *flags, "-c", os.path.join(CSRC, s), "-o", o] for s, o in zip(sources, objs)]
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
        # A library that does not match the tree must never be used silently: tests and benchmarks would run yesterday's
        # kernels.  FISHING_ALLOW_STALE_LIB=1 is the explicit way out (e.g. a box without nvcc).
        if os.path.exists(out) and os.environ.get("FISHING_ALLOW_STALE_LIB") == "1":
            print("gym_fishing_b200.build: rebuild failed, FISHING_ALLOW_STALE_LIB=1: keeping the existing library\n" + log[-400:])
            return out
        raise RuntimeError("nvcc failed (set FISHING_ALLOW_STALE_LIB=1 to keep a library built from other sources):\n" +
                           " ".join(cmd) + "\n" + log[-4000:])
    os.replace(out + ".tmp", out)
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
