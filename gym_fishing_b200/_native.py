"""ctypes binding of libfishing_b200.so (the C ABI declared in include/fishing_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is called without a CUDA
device, an exception is raised.  `python __graft_entry__.py build` (or gym_fishing_b200.build.build()) compiles the
library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FISHING_B200_LIB selects another build of the same library (A/B runs of kernel variants); never a different backend.
LIB_PATH = os.environ.get("FISHING_B200_LIB") or os.path.join(_HERE, "libfishing_b200.so")

# enums (include/fishing_b200.h)
V0, V1, V2, V4 = 0, 1, 2, 4
V5, V6, V7, V8, V9, V10, V11 = 5, 6, 7, 8, 9, 10, 11
KIND_OF_ID = {"fishing-v0": V0, "fishing-v1": V1, "fishing-v2": V2, "fishing-v4": V4, "fishing-v5": V5, "fishing-v6": V6,
              "fishing-v7": V7, "fishing-v8": V8, "fishing-v9": V9, "fishing-v10": V10, "fishing-v11": V11}
LOGNORMAL_KINDS = (V5, V6, V7, V8, V9, V10, V11)
FN_ALLEN, FN_BEVERTON_HOLT, FN_MYERS, FN_MAY, FN_RICKER = 0, 1, 2, 3, 4
FN_OF_NAME = {"allen": FN_ALLEN, "beverton_holt": FN_BEVERTON_HOLT, "myers": FN_MYERS, "may": FN_MAY, "ricker": FN_RICKER}
MAX_MODELS = 5
ABI_VERSION = 3
ACT_F64, ACT_F32, ACT_I32, ACT_I64 = 0, 1, 2, 3
AUTO_RESET, FREEZE_DONE = 1, 2
POLICY_ACTIONS, POLICY_ESCAPEMENT, POLICY_MSY = 0, 1, 2
STREAM_STEP, STREAM_AUTORESET, STREAM_RESET = 0, 1, 2
WIRE_OBS_F32, WIRE_REWARD_F32, WIRE_DONE_BITS = 1, 2, 4
ARRAY_STATE, ARRAY_YEARS, ARRAY_K, ARRAY_R, ARRAY_MODEL, ARRAY_TERMINAL = 0, 1, 2, 3, 4, 5

EXPORTS = (
    "fishing_abi_version", "fishing_build_id", "fishing_error_string", "fishing_launch_count", "fishing_reset", "fishing_step",
    "fishing_noise_fill", "fishing_rollout", "fishing_reduce_scratch_bytes", "fishing_reduce_stats", "fishing_bmsy",
    "fishing_vec_create", "fishing_vec_destroy", "fishing_vec_reset", "fishing_vec_step", "fishing_host_alloc",
    "fishing_host_free", "fishing_vec_soa", "fishing_population_draw", "fishing_step_many", "fishing_vec_fetch_terminal", "fishing_vec_step_many",
    "fishing_reset_models", "fishing_population_draw_ln", "fishing_vec_step_many_f32", "fishing_probe_fp64", "fishing_step_ex",
    "fishing_vec_step_many_wire", "fishing_vec_fetch_array", "fishing_vec_store_array", "fishing_vec_set_params",
    "fishing_probe_math",
)


class GrowthParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("r", "K", "C", "M", "theta", "q", "b", "a", "sigma")]


class FishingParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_actions", C.c_int32), ("Tmax", C.c_int32), ("flags", C.c_uint32),
                ("r", C.c_double), ("K", C.c_double), ("sigma", C.c_double), ("C", C.c_double), ("x0", C.c_double),
                ("sigma_p", C.c_double), ("seed", C.c_uint64),
                # ABI v2: log-normal growth family
                ("growth_K", C.c_double), ("M", C.c_double), ("theta", C.c_double), ("q", C.c_double), ("b", C.c_double),
                ("a", C.c_double), ("alpha", C.c_double), ("n_models", C.c_int32), ("model_fn", C.c_int32 * MAX_MODELS),
                ("model_params", GrowthParams * MAX_MODELS)]


class FishingSoA(C.Structure):
    _fields_ = [("state", C.c_void_p), ("years", C.c_void_p), ("K_env", C.c_void_p), ("r_env", C.c_void_p),
                ("model", C.c_void_p)]


class FishingStepExtra(C.Structure):
    _fields_ = [("ep_return", C.c_void_p), ("last_ep_return", C.c_void_p), ("last_ep_length", C.c_void_p),
                ("out_obs32", C.c_void_p), ("out_reward32", C.c_void_p), ("out_done_bits", C.c_void_p)]


class FishingRolloutArgs(C.Structure):
    _fields_ = [("policy", C.c_int32), ("action_dtype", C.c_int32), ("T", C.c_int64), ("step0", C.c_uint64),
                ("S", C.c_double), ("S_env", C.c_void_p), ("msy_q", C.c_double), ("actions", C.c_void_p),
                ("noise", C.c_void_p), ("noise_id_div", C.c_int64), ("out_return", C.c_void_p),
                ("out_done_step", C.c_void_p), ("trace_obs", C.c_void_p), ("trace_reward", C.c_void_p),
                ("trace_done", C.c_void_p), ("trace_action", C.c_void_p)]


class FishingError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().fishing_error_string(code)
        super().__init__(f"{where}: error {code}: {msg.decode() if msg else '?'}")


_lib = None


def lib():
    """Load the shared library (once).  Raises ImportError with build instructions if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("FISHING_B200_LIB") and os.environ.get("FISHING_ALLOW_STALE_LIB") != "1":
        # a library compiled from other sources than the tree holds is rebuilt here (nvcc is part of the image); if that
        # fails the error is raised -- the engine never runs yesterday's kernels silently
        from .build import build, needs_build

        if needs_build():
            build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA engine is not built and gym_fishing_b200 has no CPU fallback. "
            "Run `python __graft_entry__.py build` (needs nvcc).")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER(FishingParams)
    L.fishing_abi_version.restype = C.c_int
    L.fishing_abi_version.argtypes = []
    L.fishing_error_string.restype = C.c_char_p
    L.fishing_error_string.argtypes = [C.c_int]
    L.fishing_launch_count.restype = u64
    L.fishing_launch_count.argtypes = []
    L.fishing_probe_fp64.restype = C.c_int
    L.fishing_probe_fp64.argtypes = [vp, i32, i32, vp]
    L.fishing_probe_math.restype = C.c_int
    L.fishing_probe_math.argtypes = [i32, i64, vp, vp, vp]
    L.fishing_reset.restype = C.c_int
    L.fishing_reset.argtypes = [P, i64, i64, u64, vp, vp, FishingSoA, vp, vp]
    L.fishing_reset_models.restype = C.c_int
    L.fishing_reset_models.argtypes = [P, i64, i64, u64, vp, vp, FishingSoA, vp, vp]
    L.fishing_population_draw_ln.restype = C.c_int
    L.fishing_population_draw_ln.argtypes = [P, i32, i64, vp, vp, vp, vp]
    L.fishing_step.restype = C.c_int
    L.fishing_step.argtypes = [P, i64, i64, u64, vp, i32, vp, vp, FishingSoA, vp, vp, vp, vp, vp, vp, vp]
    L.fishing_step_ex.restype = C.c_int
    L.fishing_step_ex.argtypes = [P, i64, i64, u64, vp, i32, vp, vp, FishingSoA, vp, vp, vp, vp, C.POINTER(FishingStepExtra), vp]
    L.fishing_step_many.restype = C.c_int
    L.fishing_step_many.argtypes = [P, i64, i64, u64, i64, vp, i32, vp, FishingSoA, vp, vp, vp, vp, vp]
    L.fishing_vec_step_many.restype = C.c_int
    L.fishing_vec_step_many.argtypes = [vp, i64, vp, i32, vp, vp, vp]
    L.fishing_vec_step_many_f32.restype = C.c_int
    L.fishing_vec_step_many_f32.argtypes = [vp, i64, vp, i32, vp, vp, vp]
    L.fishing_vec_step_many_wire.restype = C.c_int
    L.fishing_vec_step_many_wire.argtypes = [vp, i64, vp, i32, vp, vp, vp, C.c_uint32]
    L.fishing_vec_fetch_array.restype = C.c_int
    L.fishing_vec_fetch_array.argtypes = [vp, i32, vp]
    L.fishing_vec_store_array.restype = C.c_int
    L.fishing_vec_store_array.argtypes = [vp, i32, vp]
    L.fishing_vec_set_params.restype = C.c_int
    L.fishing_vec_set_params.argtypes = [vp, P]
    L.fishing_vec_fetch_terminal.restype = C.c_int
    L.fishing_vec_fetch_terminal.argtypes = [vp, vp]
    L.fishing_noise_fill.restype = C.c_int
    L.fishing_noise_fill.argtypes = [u64, i32, i64, i64, u64, i64, vp, vp]
    L.fishing_rollout.restype = C.c_int
    L.fishing_rollout.argtypes = [P, i64, i64, C.POINTER(FishingRolloutArgs), FishingSoA, vp]
    L.fishing_reduce_scratch_bytes.restype = i64
    L.fishing_reduce_scratch_bytes.argtypes = [i64, i64]
    L.fishing_reduce_stats.restype = C.c_int
    L.fishing_reduce_stats.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    L.fishing_bmsy.restype = C.c_int
    L.fishing_bmsy.argtypes = [i32, dbl, dbl, dbl, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.fishing_population_draw.restype = C.c_int
    L.fishing_population_draw.argtypes = [P, i64, vp, i32, dbl, vp, vp]
    L.fishing_vec_create.restype = C.c_int
    L.fishing_vec_create.argtypes = [P, i64, i64, i32, C.POINTER(vp)]
    L.fishing_vec_destroy.restype = C.c_int
    L.fishing_vec_destroy.argtypes = [vp]
    L.fishing_vec_reset.restype = C.c_int
    L.fishing_vec_reset.argtypes = [vp, vp]
    L.fishing_vec_step.restype = C.c_int
    L.fishing_vec_step.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.fishing_host_alloc.restype = C.c_int
    L.fishing_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.fishing_host_free.restype = C.c_int
    L.fishing_host_free.argtypes = [vp]
    L.fishing_vec_soa.restype = C.c_int
    L.fishing_vec_soa.argtypes = [vp, C.POINTER(FishingSoA)]
    if L.fishing_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {L.fishing_abi_version()} != {ABI_VERSION}; rebuild")
    L.fishing_build_id.restype = C.c_char_p
    L.fishing_build_id.argtypes = []
    if not os.environ.get("FISHING_B200_LIB") and os.environ.get("FISHING_ALLOW_STALE_LIB") != "1":
        # the product library must have been compiled from THIS tree (an A/B build named by FISHING_B200_LIB is exempt)
        from .build import source_id

        have, want = L.fishing_build_id().decode(), source_id()
        if have != want:
            raise ImportError(f"{LIB_PATH} was built from other sources (build id {have}, tree {want}); run "
                              "`python __graft_entry__.py build`, or set FISHING_ALLOW_STALE_LIB=1 to use it anyway")
    _lib = L
    return L


def build_id():
    return lib().fishing_build_id().decode()


def check(code, where):
    if code != 0:
        raise FishingError(code, where)


def make_params(kind, *, r=0.3, K=1.0, sigma=0.0, C_=0.5, x0=0.75, Tmax=100, n_actions=100, sigma_p=0.1, seed=0,
                flags=0, growth_K=None, M=1.0, theta=1.0, q=1.0, b=1.0, a=0.0, alpha=0.0, models=None):
    """models (fishing-v11): list of (growth function name, params dict) in the order of the env's `models` list."""
    p = FishingParams(kind=int(kind), n_actions=int(n_actions), Tmax=int(Tmax), flags=int(flags), r=float(r),
                      K=float(K), sigma=float(sigma), C=float(C_), x0=float(x0), sigma_p=float(sigma_p),
                      seed=int(seed) & 0xFFFFFFFFFFFFFFFF, growth_K=float(K if growth_K is None else growth_K),
                      M=float(M), theta=float(theta), q=float(q), b=float(b), a=float(a), alpha=float(alpha))
    if models:
        if len(models) > MAX_MODELS:
            raise ValueError(f"at most {MAX_MODELS} models")
        p.n_models = len(models)
        for i, (name, mp) in enumerate(models):
            p.model_fn[i] = FN_OF_NAME[name]
            g = p.model_params[i]
            g.r, g.K, g.sigma = float(mp.get("r", 0.0)), float(mp.get("K", 1.0)), float(mp.get("sigma", 0.0))
            g.C, g.M, g.theta = float(mp.get("C", 0.0)), float(mp.get("M", 1.0)), float(mp.get("theta", 1.0))
            g.q, g.b, g.a = float(mp.get("q", 1.0)), float(mp.get("b", 1.0)), float(mp.get("a", 0.0))
    return p


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("gym_fishing_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch
