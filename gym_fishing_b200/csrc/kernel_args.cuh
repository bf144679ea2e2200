// Kernel argument blocks and launcher prototypes shared by step_kernels.cu, rollout_kernels.cu and abi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dynamics.cuh"
#include "philox.cuh"

namespace fishing {

// internal launch flag (upper bits of StepArgs::flags, never part of the public fishing_flags): use programmatic
// dependent launch for this step
constexpr uint32_t STEP_FLAG_PDL = 1u << 16;
// internal: phase A of K1 prefetches this thread's state lines into the L2 (set by the host when the state no longer fits there)
constexpr uint32_t STEP_FLAG_PREFETCH = 1u << 17;

struct StepArgs {
  int64_t n;          // environments in this launch
  int64_t env_id0;    // global id of element 0 (keys the Philox stream)
  uint64_t step_idx;  // global lock-step counter t
  uint64_t seed;
  double r, K, sigma, C;
  double s_reset;     // x0 / K - 1 (x0 itself for v4), precomputed on the host with the same IEEE operations
  double K_mean, r_mean, sigma_p;
  double n_actions;
  double rcp_n_actions;  // RN(1 / n_actions), computed on the host (div_count)
  KDiv kd;               // reciprocal of the uniform K (div_k)
  int32_t Tmax;
  uint32_t flags;
  const void* __restrict__ action;
  const double* __restrict__ noise;
  const double* __restrict__ reset_noise;
  const double* state_in;
  int32_t* years;
  double* K_env;
  double* r_env;
  double* out_obs;
  double* __restrict__ out_reward;
  uint8_t* __restrict__ out_done;
  double* __restrict__ out_terminal;
  double* __restrict__ out_stock;    // optional: fish_population after growth
  double* __restrict__ out_harvest;  // optional: harvest (un-clipped reward)
  double* __restrict__ out_obs2;     // optional second copy of the observations (zero-copy host front end: out_obs is
                                     // the device state, out_obs2 a device-mapped pinned host array)
  // VecMonitor-style episode statistics (optional, all three or none; fishing_step_ex): the running return of the current
  // episode, and the return / length of the last finished one, rolled over in the step that ends it
  double* ep_return;
  double* __restrict__ last_ep_return;
  int32_t* __restrict__ last_ep_length;
  // optional float32 copies of observations / rewards (what a stable-baselines3 VecEnv hands to the learner)
  float* __restrict__ out_obs32;
  float* __restrict__ out_reward32;
  // optional bit-packed copy of out_done: bit (i & 63) of word (i >> 6) = done of environment i (vector path only)
  unsigned long long* __restrict__ out_done_bits;
  // log-normal family (fishing-v5..v11)
  GrowthP gp;
  int32_t derive_on_device;          // fishing-v10: recompute gp's derived fields with the device's log (like K2 does)
  PhiloxKeys keys;                   // round keys of `seed`
  int32_t* model;                    // v11: per-env model index (updated at auto-reset)
  ModelTable tab;                    // v11: the model table (by value in the kernel parameters)
};

struct ResetArgs {
  int64_t n, env_id0;
  uint64_t epoch, seed;
  double s_reset, K_mean, r_mean, sigma_p;
  int32_t per_env;
  int32_t n_models;                  // > 0: fishing-v11, draw / inject the model index
  const int32_t* __restrict__ reset_model;
  int32_t* model;
  const uint8_t* __restrict__ mask;
  const double* __restrict__ reset_noise;
  double* state;
  int32_t* years;
  double* K_env;
  double* r_env;
  double* out_obs;
};

struct RolloutArgs {
  int64_t n, env_id0, T;
  uint64_t step0, seed;
  double r, K, sigma, C, s_reset, K_mean, r_mean, sigma_p, n_actions, rcp_n_actions;
  KDiv kd;               // reciprocal of the uniform K (div_k)
  int32_t Tmax;
  uint32_t flags;
  int32_t discrete;      // fishing-v0: Discrete action space
  int32_t action_dtype;  // FISHING_POLICY_ACTIONS only
  double S, msy_q;
  const double* __restrict__ S_env;
  const void* __restrict__ actions;
  const double* __restrict__ noise;
  int64_t noise_id_div;
  // auto-reset in the fused fast path with K = 1: the reset is folded into the np.maximum select -- the stock becomes x_reset_fold
  // = s_reset + 1 in the step that ends an episode, and s = x - 1 then IS s_reset -- when the host has checked that
  // (s_reset + 1) - 1 == s_reset in binary64 (fold_reset; otherwise the state is overwritten after the step as before)
  double x_reset_fold;
  int32_t fold_reset;
  double* state;
  int32_t* years;
  double* K_env;
  double* r_env;
  double* __restrict__ out_return;
  int32_t* __restrict__ out_done_step;
  double* __restrict__ trace_obs;
  double* __restrict__ trace_reward;
  uint8_t* __restrict__ trace_done;
  double* __restrict__ trace_action;
  // log-normal family (fishing-v5..v11)
  GrowthP gp;
  double alpha;                      // v10: r += alpha before every draw
  int32_t nonstationary;
  int32_t* model;
  PhiloxKeys keys;                   // round keys of `seed` (the step-noise stream reads them from the constant bank)
  ModelTable tab;
};

cudaError_t launch_step(const StepArgs& a, int kind, int act, bool vec, cudaStream_t st);
cudaError_t launch_reset(const ResetArgs& a, cudaStream_t st);
cudaError_t launch_noise_fill(uint64_t seed, int32_t stream_id, int64_t n, int64_t env_id0, uint64_t step0, int64_t T,
                              double* out, cudaStream_t st);
cudaError_t launch_growth(int kind, int64_t n, const void* x, bool f32, double r, double K, double C, double sigma,
                          double z, void* out, cudaStream_t st);
cudaError_t launch_growth_ln(int fn, int64_t n, const double* x, const GrowthP& gp, bool derive_on_device, const double* z,
                             double* out, cudaStream_t st);
cudaError_t launch_fp64_probe(double* scratch, int blocks, int iters, cudaStream_t st);
cudaError_t launch_math_probe(int fn, int64_t n, const double* x, double* out, cudaStream_t st);
cudaError_t launch_bmsy(int kind, double r, double K, double C, bool f64, double* out_dev, cudaStream_t st);
cudaError_t launch_rollout(const RolloutArgs& a, int kind, int policy, cudaStream_t st);
int64_t reduce_scratch_doubles(int64_t n, int64_t groups);
cudaError_t launch_reduce(const double* ret, const int32_t* done_step, int64_t n, int64_t groups, double* scratch,
                          double* out, cudaStream_t st);

}  // namespace fishing
