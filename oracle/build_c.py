"""Builds oracle/_build/liboracle.so from oracle/fishing_oracle.c (gcc, no FMA contraction, OpenMP).  TEST INFRASTRUCTURE."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fishing_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", OUT, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed: " + res.stderr)
    return OUT


def load():
    import ctypes as C

    lib = C.CDLL(build())
    dp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    lib.oracle_step.restype = None
    lib.oracle_step.argtypes = [C.c_int, i64, dp, dp, dbl, dbl, dbl, dbl, dbl, i32, C.c_int, dbl, dp, dp, dp, dp, dp, dp]
    lib.oracle_rollout.restype = None
    lib.oracle_rollout.argtypes = [C.c_int, i64, i64, dbl, dbl, dbl, dbl, dbl, i32, dbl, dp, dp, dp, dp, dp, dp]
    return lib


if __name__ == "__main__":
    print(build(force=True))
