/*
 * fishing_b200.h -- C ABI of the B200-native batched rollout engine for gym_fishing (fishing-v0/v1/v2/v4).
 *
 * The reference (boettiger-lab/gym_fishing) is pure Python and has no FFI; these entry points are what a
 * ctypes binding on the reference side would call in place of the per-environment Python methods cited next
 * to each function (reference paths are relative to the gym_fishing repository root).  INTEGRATION.md shows
 * that binding.  Plain pointers and sizes only -- no torch / numpy types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FISHING_E_* code for a bad argument, or a positive
 *     cudaError_t; fishing_error_string() describes any of them.  Nothing is allocated or freed behind the
 *     caller's back by the device-pointer entry points; they enqueue work on `stream` (a cudaStream_t passed
 *     as void*, NULL = legacy default stream) and return without synchronising.
 *   - "state" is the reference's normalised observation s = x/K - 1 (float64).  All arrays are
 *     structure-of-arrays over environments.  Vectorised (16-byte) access is used when every pointer is
 *     16-byte aligned and env_id0 and n are even; otherwise a scalar path computes identical results.
 *   - arithmetic is IEEE-754 binary64, one rounding per reference operation, never contracted to FMA, so
 *     results are bit-identical to the reference for v0/v1/v4 and within 1 ulp of libm `exp` for v2.
 *   - noise: if `noise` is non-NULL it supplies the standard-normal draw of every environment for this step
 *     (the reference's np.random.normal(0,1), base_fishing_env.py:130); otherwise the engine generates it from
 *     a counter-based Philox4x32-10 stream keyed by (seed; env_id>>1, step_idx) + Box-Muller, so results do
 *     not depend on how environments are partitioned over GPUs.  fishing_noise_fill() exports exactly those
 *     draws so that they can be injected into the reference for parity checks.
 */
#ifndef FISHING_B200_H
#define FISHING_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FISHING_ABI_VERSION 3

/* environment ids, gym_fishing/envs/__init__.py:17-35 */
enum fishing_kind {
  FISHING_V0 = 0, /* FishingEnv, discrete actions         gym_fishing/envs/fishing_env.py:6-24        */
  FISHING_V1 = 1, /* FishingCtsEnv, Box actions           gym_fishing/envs/fishing_cts_env.py:4-12    */
  FISHING_V2 = 2, /* FishingTippingEnv                    gym_fishing/envs/fishing_tipping_env.py:6-35 */
  FISHING_V4 = 4, /* FishingModelError, per-env r and K   gym_fishing/envs/fishing_model_error.py:6-48 */
  /* "next row" f1 of SURVEY.md section 8: log-normal growth family, gym_fishing/envs/growth_models.py */
  FISHING_V5 = 5,   /* Allen          growth_models.py:6-25,208-217   */
  FISHING_V6 = 6,   /* BevertonHolt   growth_models.py:28-40,220-226  */
  FISHING_V7 = 7,   /* May            growth_models.py:75-106,229-242 */
  FISHING_V8 = 8,   /* Myers          growth_models.py:43-70,247-255  */
  FISHING_V9 = 9,   /* Ricker         growth_models.py:109-121,258-261 */
  FISHING_V10 = 10, /* NonStationary  growth_models.py:124-150 (Beverton-Holt with r += alpha before every draw) */
  FISHING_V11 = 11  /* ModelUncertainty growth_models.py:153-204 (growth model drawn per env at every reset) */
};

/* growth functions of the log-normal family: x' = max(0, exp(mu(x) + sigma z)) */
enum fishing_growth_fn { FISHING_FN_ALLEN = 0, FISHING_FN_BEVERTON_HOLT = 1, FISHING_FN_MYERS = 2, FISHING_FN_MAY = 3,
                         FISHING_FN_RICKER = 4 };
#define FISHING_MAX_MODELS 5

/* the entries of the reference's `params` dict that a growth function reads (growth_models.py:208-261) */
typedef struct fishing_growth_params {
  double r, K, C, M, theta, q, b, a, sigma;
} fishing_growth_params;

enum fishing_action_dtype {
  FISHING_ACT_F64 = 0, /* Box action as float64 (the parity path)                                         */
  FISHING_ACT_F32 = 1, /* Box action as float32: quota computed in float32 like numpy>=2 does (NEP 50)    */
  FISHING_ACT_I32 = 2, /* Discrete action as int32                                                        */
  FISHING_ACT_I64 = 3  /* Discrete action as int64 (what stable-baselines3 hands over)                    */
};

enum fishing_flags {
  FISHING_AUTO_RESET = 1u << 0, /* stable-baselines3 VecEnv semantics: reset an env in the step that ends it */
  FISHING_FREEZE_DONE = 1u << 1, /* rollout only: stop stepping an env after its first done                 */
  FISHING_PDL = 1u << 2          /* step only: launch with programmatic dependent launch.  The caller guarantees
                                    that the action / noise buffers of this step are not written by the kernel
                                    that immediately precedes it on the stream (the state, years and output
                                    buffers may be: the kernel waits for its predecessor before touching them).
                                    fishing_step_many sets this by itself.                                      */
};

enum fishing_policy {
  FISHING_POLICY_ACTIONS = 0,    /* caller-supplied action tensor [T, n]                                   */
  FISHING_POLICY_ESCAPEMENT = 1, /* escapement.predict, gym_fishing/models/policies.py:27-31               */
  FISHING_POLICY_MSY = 2         /* msy.predict (constant quota), gym_fishing/models/policies.py:16-19     */
};

enum fishing_noise_stream {
  FISHING_STREAM_STEP = 0,       /* process noise z of step t: ctr = (env>>1, t), lane = env&1              */
  FISHING_STREAM_AUTORESET = 1,  /* v4 (zK, zr) of the auto-reset fired by step t: ctr = (env, t)           */
  FISHING_STREAM_RESET = 2       /* v4 (zK, zr) of an explicit reset number `epoch`: ctr = (env, epoch)     */
};

enum fishing_error {
  FISHING_E_ARG = -1,      /* NULL / negative size / unknown enum                       */
  FISHING_E_KIND = -2,     /* action dtype or policy not valid for this environment id  */
  FISHING_E_ALIGN = -3,    /* pointer not aligned to its element size                   */
  FISHING_E_NODEVICE = -4, /* no CUDA device / not an sm_100 device                     */
  FISHING_E_STATE = -5     /* handle used before reset, or already destroyed            */
};

/* Constructor keyword arguments of the reference classes (SURVEY.md section 5 "Config"):
 * fishing_env.py:7-16, fishing_cts_env.py:5-7, fishing_tipping_env.py:7-16, fishing_model_error.py:9-19. */
typedef struct fishing_params {
  int32_t kind;      /* enum fishing_kind */
  int32_t n_actions; /* v0 only (default 100) */
  int32_t Tmax;      /* done = years_passed > Tmax (strict), base_fishing_env.py:76 */
  uint32_t flags;    /* enum fishing_flags */
  double r;          /* v4: r_mean */
  double K;          /* v4: K_mean */
  double sigma;
  double C;          /* v2 only */
  double x0;         /* init_state */
  double sigma_p;    /* v4 only */
  uint64_t seed;     /* Philox key */
  /* ---- ABI v2: fishing-v5..v11.  For these ids `r`, `sigma`, `C` above and the fields below are the entries of the
   * env's `params` dict, which is what the growth functions read; `K` stays the env ATTRIBUTE used to scale
   * states/quotas and `growth_K` is params["K"] (the two are equal unless a caller changed one). */
  double growth_K;
  double M, theta, q, b, a; /* May: M q b a; Myers: M theta */
  double alpha;             /* v10: r <- r + alpha before every population draw; `r` is the value BEFORE this step */
  int32_t n_models;         /* v11: length of the `models` list (<= FISHING_MAX_MODELS) */
  int32_t model_fn[FISHING_MAX_MODELS];                      /* v11: enum fishing_growth_fn of each list entry */
  fishing_growth_params model_params[FISHING_MAX_MODELS];    /* v11: params[models[i]] */
} fishing_params;

/* Structure-of-arrays environment state (device pointers for the functions below). */
typedef struct fishing_soa {
  double* state;  /* [n] normalised state s; v4 straight after a reset holds x0 un-normalised (upstream quirk,
                     fishing_model_error.py:44) */
  int32_t* years; /* [n] years_passed */
  double* K_env;  /* [n] v4 only, else NULL */
  double* r_env;  /* [n] v4 only, else NULL */
  int32_t* model; /* [n] v11 only, else NULL: index into the `models` list, redrawn at every reset */
} fishing_soa;

int fishing_abi_version(void);
/* Identifier of the sources the library was compiled from (16 hex digits; gym_fishing_b200/build.py:source_id).  New in
 * round 2: the Python loader refuses a library whose id differs from the tree it sits in.  No reference counterpart. */
const char* fishing_build_id(void);
const char* fishing_error_string(int code);

/* Number of kernel launches issued through this library since it was loaded (all entry points). */
uint64_t fishing_launch_count(void);

/*
 * Measurement aid (no reference counterpart): one launch of `blocks` x 256 threads, each running `iters` rounds of 8
 * independent double-precision FMA chains = blocks * 8 * iters * 8 DFMA warp-instructions.  Timed by the caller with
 * events, it gives the FP64-pipe peak that the fused rollout kernel's roofline is quoted against (bench.py).
 * `scratch` is a device array of >= 256 doubles (never written in practice).
 */
int fishing_probe_fp64(double* scratch, int32_t blocks, int32_t iters, void* stream);
/*
 * Test aid: out[i] = f(x[i]) for the engine's own table-assisted float64 functions (csrc/normal.cuh), which stand where the
 * reference calls numpy: fn = 0 log (np.log in growth_models.py:208-261), fn = 1 exp (np.exp in fishing_tipping_env.py:31,
 * np.random.lognormal in growth_models.py).  x, out: device arrays of n doubles.
 */
int fishing_probe_math(int32_t fn, int64_t n, const double* x, double* out, void* stream);

/*
 * reset -- BaseFishingEnv.reset (base_fishing_env.py:83-91), FishingModelError.reset / __init__ draws
 * (fishing_model_error.py:35-48).  `mask` (uint8[n], NULL = all) selects environments.  v4 draws K then r per
 * env from reset_noise[0..n) and reset_noise[n..2n) when given, else from Philox stream FISHING_STREAM_RESET
 * at counter `epoch`.  v11 draws the model index (growth_models.py:200): from reset_model[n] (int32) when given, else
 * floor(u * n_models) with u the first Philox word of the same stream.  Writes soa.* and, if non-NULL, out_obs[n].
 */
int fishing_reset(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t epoch, const uint8_t* mask,
                  const double* reset_noise, fishing_soa soa, double* out_obs, void* stream);
/* same, with injected fishing-v11 model indices (int32[n]) */
int fishing_reset_models(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t epoch, const uint8_t* mask,
                         const int32_t* reset_model, fishing_soa soa, double* out_obs, void* stream);

/*
 * step -- BaseFishingEnv.step (base_fishing_env.py:60-81) with get_quota (:135-147), get_fish_population
 * (:158-160), harvest_draw (:112-119), population_draw (:121-133; v2: fishing_tipping_env.py:24-35) and
 * get_state (:162-164) fused, for n environments in lock-step.
 *   action       [n] of `action_dtype`
 *   noise        [n] standard normals or NULL (see header comment)
 *   reset_noise  [2n] (zK then zr) for v4 auto-resets or NULL (Philox stream FISHING_STREAM_AUTORESET)
 *   out_obs      [n] next observation (after auto-reset if that fired); may alias soa.state
 *   out_reward   [n] max(harvest, 0)
 *   out_done     [n] uint8, years_passed > Tmax or stock <= 0
 *   out_terminal [n] or NULL: observation that ended the episode, written only where done (auto-reset mode)
 *   out_stock    [n] or NULL: fish_population after growth (the attribute the reference leaves on the env)
 *   out_harvest  [n] or NULL: harvest = min(stock, quota) before the max(., 0) of the reward
 * step_idx is the global lock-step counter t (it keys the Philox stream); env_id0 the global id of element 0.
 */
int fishing_step(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step_idx, const void* action,
                 int32_t action_dtype, const double* noise, const double* reset_noise, fishing_soa soa,
                 double* out_obs, double* out_reward, uint8_t* out_done, double* out_terminal, double* out_stock,
                 double* out_harvest, void* stream);

/*
 * step with the outputs a stable-baselines3 loop wants next to the reference's (ABI v3).  Same kernel family; everything in
 * `x` is optional (NULL):
 *   ep_return / last_ep_return / last_ep_length  [n] each, all three or none: VecMonitor-style episode statistics kept on
 *       the device -- ep_return accumulates the reward of the running episode; in the step that ends an episode its value
 *       and the episode length go to last_ep_* and the accumulator restarts at 0 (stable_baselines3.common.vec_env
 *       VecMonitor, which wraps the VecEnv the reference's tests train on: tests/test-PPO.py:9-21)
 *   out_done_bits  the done flags bit-packed (1/8 of the bytes over PCIe)
 *   out_obs32 / out_reward32  [n] float32 copies of out_obs / out_reward (the observation space is declared float32,
 *       base_fishing_env.py:49-58, and DummyVecEnv keeps rewards in a float32 buffer)
 */
typedef struct fishing_step_extra {
  double* ep_return;
  double* last_ep_return;
  int32_t* last_ep_length;
  float* out_obs32;
  float* out_reward32;
  uint64_t* out_done_bits; /* [ceil(n/64)] or NULL: out_done bit-packed, bit (i & 63) of word (i >> 6) = environment i; needs the
                              vector path (even n, even env_id0, 16-byte aligned arrays), FISHING_E_ALIGN otherwise */
} fishing_step_extra;
int fishing_step_ex(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step_idx, const void* action,
                    int32_t action_dtype, const double* noise, const double* reset_noise, fishing_soa soa, double* out_obs,
                    double* out_reward, uint8_t* out_done, double* out_terminal, const fishing_step_extra* x, void* stream);

/*
 * step_many -- T consecutive lock-steps into a caller-owned rollout buffer (what an on-policy collector fills):
 * step t reads actions[t*n ..], writes out_obs / out_reward / out_done (/ out_terminal) at [t*n ..], and takes the
 * observations of step t-1 as its state, so no separate state array is read or written between steps.  One kernel
 * launch per step, issued back to back on `stream` (capturable into a CUDA graph); soa.state is updated with the
 * final state at the end.  Identical results to T calls of fishing_step with step_idx = step0 + t.
 */
int fishing_step_many(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step0, int64_t T,
                      const void* actions, int32_t action_dtype, const double* noise, fishing_soa soa, double* out_obs,
                      double* out_reward, uint8_t* out_done, double* out_terminal, void* stream);

/*
 * noise_fill -- exports the engine's own draws: out[t*n + i] = z(env_id0 + i, step0 + t) for stream STEP, or
 * out[(t*2 + j)*n + i] = (zK, zr)[j] for the v4 streams.  Same device code as used inside step / rollout.
 */
int fishing_noise_fill(uint64_t seed, int32_t noise_stream, int64_t n, int64_t env_id0, uint64_t step0, int64_t T,
                       double* out, void* stream);

/*
 * rollout -- T fused steps with the state held in registers: the loop of simulate_mdp (shared_env.py:29-54)
 * with the policy applied in-kernel.
 *   policy        enum fishing_policy; ESCAPEMENT uses S (float32-valued BMSY result widened to double, or
 *                 S_env[n] if non-NULL, e.g. an escapement sweep); MSY uses the constant quota msy_q
 *   actions       [T, n] for FISHING_POLICY_ACTIONS (dtype action_dtype), else NULL
 *   noise         [T, n] or NULL (Philox)
 *   noise_id_div  0/1: noise keyed by env id; m > 1: env id / m is the noise id (common random numbers
 *                 across the m policy levels of a sweep, SURVEY.md section 8d cfg 5)
 *   out_return    [n] sum of rewards up to and including the first done step (sequential float64 adds)
 *   out_done_step [n] int32 first t (0-based, relative to this call) with done, -1 if none
 *   trace_*       [T, n] optional per-step outputs (NULL to skip)
 * soa is read at entry and written back at exit (final state, years, K/r).  With FISHING_AUTO_RESET the
 * environments restart inside the kernel and out_return accumulates over all episodes; with
 * FISHING_FREEZE_DONE an environment stops at its first done (what simulate_mdp's `break` does).
 */
typedef struct fishing_rollout_args {
  int32_t policy;
  int32_t action_dtype;
  int64_t T;
  uint64_t step0;
  double S;
  const double* S_env;
  double msy_q;
  const void* actions;
  const double* noise;
  int64_t noise_id_div;
  double* out_return;
  int32_t* out_done_step;
  double* trace_obs;
  double* trace_reward;
  uint8_t* trace_done;
  double* trace_action;
} fishing_rollout_args;

int fishing_rollout(const fishing_params* p, int64_t n, int64_t env_id0, const fishing_rollout_args* a,
                    fishing_soa soa, void* stream);

/*
 * reduce_stats -- fixed-shape (pairwise tree, no atomics) reduction of per-env returns to
 * out[g*3 + {0,1,2}] = {sum ret, sum ret^2, number with done_step >= 0} for `groups` interleaved groups:
 * element i belongs to group i % groups (groups = 1: everything; groups = 1024: the levels of a sweep laid
 * out [replicate, level]).  scratch must hold fishing_reduce_scratch_bytes(n, groups) bytes.
 */
int64_t fishing_reduce_scratch_bytes(int64_t n, int64_t groups);
int fishing_reduce_stats(const double* ret, const int32_t* done_step, int64_t n, int64_t groups, void* scratch,
                         double* out, void* stream);

/*
 * bmsy -- BMSY (gym_fishing/models/policies.py:51-67): argmax of deterministic growth over the 10001-point
 * float32 grid on [0, 2K] (float64 grid when K_is_f64, the fishing-v4 case).  Synchronous; writes S and
 * f(S) - S (the msy quota, policies.py:12) to host memory.
 */
int fishing_bmsy(int32_t kind, double r, double K, double C, int32_t K_is_f64, double* S_out, double* msy_out);

/*
 * population_draw -- BaseFishingEnv.population_draw (base_fishing_env.py:121-133; v2: fishing_tipping_env.py:24-35)
 * on an array of stocks with ONE scalar draw z for the whole array, the way the reference's own callers use it
 * (BMSY / msy, policies.py:9-14,60-65).  x and out are device arrays of float32 (is_f32 != 0: every operation is a
 * float32 operation, numpy's result for a float32 array and Python-number parameters) or float64.
 */
int fishing_population_draw(const fishing_params* p, int64_t n, const void* x, int32_t is_f32, double z, void* out,
                            void* stream);

/* the same for the log-normal growth family (fishing-v5..v11): float64 arrays, one draw z[i] PER ELEMENT (NULL = 0), the way
 * np.random.lognormal(mu_array, sigma) draws (growth_models.py:217-261); model_index selects the fishing-v11 list entry */
int fishing_population_draw_ln(const fishing_params* p, int32_t model_index, int64_t n, const double* x, const double* z,
                               double* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host-buffer front end: the engine owns device state; the caller passes ordinary host arrays (numpy, pinned
 * or pageable).  This is the entry a reference-side VecEnv binding uses and what bench.py's
 * `e2e` figure times.  Copies are chunked and pipelined over three streams.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct fishing_vec fishing_vec;

int fishing_vec_create(const fishing_params* p, int64_t n, int64_t env_id0, int32_t device, fishing_vec** out);
int fishing_vec_destroy(fishing_vec* v);
/* reset all environments; obs_out[n] host */
int fishing_vec_reset(fishing_vec* v, double* obs_out);
/* one lock-step; all pointers host; terminal_out may be NULL */
int fishing_vec_step(fishing_vec* v, const void* action, int32_t action_dtype, const double* noise, double* obs_out,
                     double* reward_out, uint8_t* done_out, double* terminal_out);
/* page-locked host memory helpers so that callers can make their numpy buffers DMA-able */
int fishing_host_alloc(void** ptr, int64_t bytes);
int fishing_host_free(void* ptr);
/* T lock-steps in one call: actions[T, n] in, obs / reward / done [T, n] out (all host).  Copies and kernels of
 * consecutive steps are pipelined over the handle's streams (the host-side counterpart of fishing_step_many). */
int fishing_vec_step_many(fishing_vec* v, int64_t T, const void* actions, int32_t action_dtype, double* obs_out,
                          double* reward_out, uint8_t* done_out);
/* the same with the dtypes a stable-baselines3 VecEnv hands to the learner: observations and rewards as float32 (the
 * declared observation space is float32, base_fishing_env.py:54-58, and DummyVecEnv keeps rewards in a float32 buffer); the
 * state stays float64 on the device, the float32 values are its round-to-nearest casts.  Halves the PCIe traffic. */
int fishing_vec_step_many_f32(fishing_vec* v, int64_t T, const void* actions, int32_t action_dtype, float* obs_out,
                              float* reward_out, uint8_t* done_out);
/* the general form (ABI v3): every output array picks its wire format.  obs_out / reward_out are [T, n] float64, or
 * float32 with FISHING_WIRE_OBS_F32 / FISHING_WIRE_REWARD_F32; done_out is [T, n] uint8, or with FISHING_WIRE_DONE_BITS
 * [T, ceil(n/8)] bytes, bit (i & 7) of byte (i >> 3) = environment i (numpy.unpackbits(..., bitorder="little")).
 * float32 actions in, float32 observations and rewards and packed dones out is 4 + 8.125 bytes per environment-step over
 * PCIe instead of 8 + 17. */
enum fishing_wire { FISHING_WIRE_OBS_F32 = 1, FISHING_WIRE_REWARD_F32 = 2, FISHING_WIRE_DONE_BITS = 4 };
int fishing_vec_step_many_wire(fishing_vec* v, int64_t T, const void* actions, int32_t action_dtype, void* obs_out,
                               void* reward_out, void* done_out, uint32_t wire);
/* terminal observations of the most recent step (valid where that step reported done), fetched on demand so that the
 * per-step device-to-host traffic stays at obs + reward + done */
int fishing_vec_fetch_terminal(fishing_vec* v, double* terminal_out);
/* Per-environment arrays of the handle, copied to / from host memory (ABI v3): what the reference exposes as attributes of
 * each env object (state, years_passed: base_fishing_env.py:35-39; fishing-v4's drawn K, r: fishing_model_error.py:37-38)
 * and a stable-baselines3 VecEnv reads / writes with get_attr / set_attr.  [n] float64 or int32 by array;
 * FISHING_E_KIND when the environment id has no such array. */
enum fishing_array { FISHING_ARRAY_STATE = 0, FISHING_ARRAY_YEARS = 1, FISHING_ARRAY_K = 2, FISHING_ARRAY_R = 3,
                     FISHING_ARRAY_MODEL = 4, FISHING_ARRAY_TERMINAL = 5 };
int fishing_vec_fetch_array(fishing_vec* v, int32_t which, void* out);
int fishing_vec_store_array(fishing_vec* v, int32_t which, const void* in);
/* replace the constructor parameters (r, K, sigma, Tmax, ... ; same environment id) -- VecEnv.set_attr on a shared parameter */
int fishing_vec_set_params(fishing_vec* v, const fishing_params* p);
/* device pointers of the handle's SoA state, for inspection */
int fishing_vec_soa(fishing_vec* v, fishing_soa* out);

#ifdef __cplusplus
}
#endif
#endif /* FISHING_B200_H */
