// K1: the lock-step single-step kernel, K3: reset / fishing-v4 parameter draws, plus the noise exporter and BMSY.
// HBM-bound structure-of-arrays float64 streaming: per env-step the kernel reads state(8) + action(8|4) +
// years(4) and writes obs(8) + reward(8) + done(1) + years(4); see DESIGN.md for the byte accounting.
#include <cuda_runtime.h>

#include "common.cuh"
#include "dynamics.cuh"
#include "kernel_args.cuh"
#include "philox.cuh"

namespace fishing {

// One environment: everything after the loads.  Returns the values to store.
struct EnvResult {
  double obs, reward, K, r;
  int32_t years, model;
  uint8_t done;
  bool reset_fired;
  double terminal, stock, harvest;
};

template <int GROWTH, int PM, int ACT, bool DEFER_V4_DRAW = false, bool LEAN = false>
__device__ __forceinline__ EnvResult advance(const StepArgs& a, int64_t i, double s, typename ActT<ACT>::type act,
                                             double z, bool noisy, int32_t years, double K, double r,
                                             int32_t model = 0) {
  const double quota = quota_of<PM, ACT>(act, a.n_actions, a.rcp_n_actions, K);
  GrowthP gp_local;
  const GrowthP* gp = &a.gp;
  if (GROWTH == G_BEVERTON_HOLT && a.derive_on_device) {
    gp_local = a.gp;
    growth_derive(gp_local, FISHING_FN_BEVERTON_HOLT);
    gp = &gp_local;
  }
  const StepOut o = step_core<GROWTH, PM>(s, quota, z, r, K, a.C, a.sigma, noisy, gp, &a.tab, model, a.kd);
  EnvResult e;
  e.model = model;
  e.years = years + 1;                                   // base_fishing_env.py:75
  e.done = (e.years > a.Tmax) || o.dead;                 // :76-79
  e.obs = o.s2;
  e.reward = o.reward;
  e.K = K;
  e.r = r;
  e.reset_fired = false;
  e.terminal = o.s2;
  e.stock = o.stock;
  e.harvest = o.harvest;
  if ((a.flags & FISHING_AUTO_RESET) && e.done) {        // SB3 VecEnv: reset inside the step that ended the episode
    e.reset_fired = true;
    e.years = 0;
    const bool injected = !LEAN && a.reset_noise != nullptr;     // LEAN: the caller passed no optional array at all
    if (PM == P_PERENV && !(DEFER_V4_DRAW && !injected)) {        // FishingModelError.reset, fishing_model_error.py:41-48
      double zK, zr;
      if (injected) {
        zK = a.reset_noise[i];
        zr = a.reset_noise[a.n + i];
      } else {
        const Normal2 d = normal_pair_keys(a.keys, (uint64_t)(a.env_id0 + i), a.step_idx, FISHING_STREAM_AUTORESET);
        zK = d.z0;
        zr = d.z1;
      }
      e.K = v4_param(a.K_mean, a.sigma_p, zK);
      e.r = v4_param(a.r_mean, a.sigma_p, zr);
    }
    if (GROWTH == G_MODEL)                                // ModelUncertainty.reset, growth_models.py:197-204
      e.model = (int32_t)uniform_index(a.seed, (uint64_t)(a.env_id0 + i), a.step_idx, FISHING_STREAM_AUTORESET,
                                       (uint32_t)a.tab.n_models);
    e.obs = a.s_reset;
  }
  return e;
}

// ------------------------------------------------------------------------------------------------ K1, vector path
// One thread owns ONE PAIR of adjacent environments (2j, 2j+1): 16-byte loads/stores on the float64 arrays, one Philox
// block + one Box-Muller for the pair.  256-thread blocks, one pair per thread -> 2048 blocks at n = 2^20: fine-grained
// enough that the tail of launch t is filled by the head of launch t+1 when programmatic dependent launch is on.
//
// The kernel has two phases.  Phase A touches only this launch's own inputs (actions, injected noise) and does all the
// state-independent work -- the Philox rounds and the fp64 Box-Muller, i.e. most of the instructions.  Phase B reads
// the state written by the previous launch.  With PDL (sm_90+ `griddepcontrol`) every block first releases the next
// launch in the stream, and waits for the previous launch only between the phases, so consecutive lock-steps overlap:
// the next step's noise is generated while this step drains.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

//
// LEAN = the caller passed none of the optional arrays (injected noise, second observation copy, stock / harvest /
// terminal-observation outputs): their null tests and address arithmetic -- about 20 of the kernel's 333 instructions per
// pair, in a kernel that is bound by instruction issue -- are compiled out (cfg 2: 6.83 -> 6.53 us per launch).  The
// rollout-buffer loop of bench.py and FishingVecEnv's torch path run this instantiation.  The lean fishing-v4 kernel is
// held to 48 registers (1280 resident threads per SM instead of 1024): 6.78 -> 6.57 ms per 128 steps of cfg 4.
// EXTRA (never together with LEAN) = the instantiation that also keeps the VecMonitor statistics and writes float32 copies of
// observations / rewards (fishing_step_ex): a kernel of its own, so that neither the lean nor the plain instantiation pays
// registers for it.
#ifndef FISHING_V4_RESIDENT
#define FISHING_V4_RESIDENT 1280      // resident threads per SM the lean fishing-v4 kernel is compiled for (timing build: 1536 = 40 registers)
#endif
template <int GROWTH, int PM, int ACT, int BS = 128, bool LEAN = false, bool EXTRA = false>
__global__ void __launch_bounds__(BS, (LEAN && PM == P_PERENV && BS <= 256) ? FISHING_V4_RESIDENT / BS : 0) step_kernel_vec2(const StepArgs a) {
  using AT2 = typename ActT<ACT>::vec2;
  if (a.flags & STEP_FLAG_PDL) pdl_launch_dependents();
  // fishing-v4 reset jobs of the whole block (see the end of the kernel)
  __shared__ unsigned s_count, s_arrived;
  __shared__ uint16_t s_job[2 * BS];
  if (PM == P_PERENV) {
    if (threadIdx.x == 0) { s_count = 0; s_arrived = 0; }
    __syncthreads();
  }
  const int npairs = (int)(a.n >> 1);
  const int j = (int)(blockIdx.x * (unsigned)BS + threadIdx.x);
  const bool live = j < npairs;
  const bool ext_noise = !LEAN && a.noise != nullptr;
  const bool noisy = ext_noise || a.sigma != 0.0;

  // ---- phase A: independent of the previous launch
  AT2 act = {};
  double2 z = make_double2(0.0, 0.0);
  if (live) {
    act = ld_stream(reinterpret_cast<const AT2*>(a.action) + j);
    // Start pulling this thread's state lines towards the L2 now, so that phase B finds them close by instead of paying a
    // second DRAM trip after the wait.  Only when the per-launch state no longer fits in the L2 (the host sets the flag from
    // n x bytes of state per environment, abi.cu): measured in round 2 (tools/k1_scaling.py, fishing-v1 / fishing-v4, sigma = 0.01)
    // the prefetch costs 2 - 3 % while the state is L2-resident (2^20 .. 2^22 / 2^21 envs) and gains 1 - 4 % beyond.
    if (a.flags & STEP_FLAG_PREFETCH) {
      prefetch_l2(reinterpret_cast<const double2*>(a.state_in) + j);
      prefetch_l2(reinterpret_cast<const int2*>(a.years) + j);
      if (PM == P_PERENV) {
        prefetch_l2(reinterpret_cast<const double2*>(a.K_env) + j);
        prefetch_l2(reinterpret_cast<const double2*>(a.r_env) + j);
      }
    }
    if (ext_noise) z = ld_stream(reinterpret_cast<const double2*>(a.noise) + j);
  }
#ifdef FISHING_K1_SMEM_TABLES
  // timing build: the block stages the Box-Muller tables in shared memory instead of reading them through the L1
  __shared__ double2 s_tab[NORMAL_TABLE_ENTRIES];
  if (noisy && !ext_noise) {     // block-uniform
    const SharedTabs tb = stage_normal_tables<BS>(s_tab);
    if (live) {
      const Normal2 d = normal_pair_keys(a.keys, ((uint64_t)a.env_id0 >> 1) + (uint64_t)j, a.step_idx, FISHING_STREAM_STEP, tb);
      z = make_double2(d.z0, d.z1);
    }
  }
#else
  if (live && noisy && !ext_noise) {
    const Normal2 d = normal_pair_keys(a.keys, ((uint64_t)a.env_id0 >> 1) + (uint64_t)j, a.step_idx, FISHING_STREAM_STEP);
    z = make_double2(d.z0, d.z1);
  }
#endif
  if (a.flags & STEP_FLAG_PDL) pdl_wait();
  // the fishing-v4 kernel keeps idle lanes alive: every warp of the block has to sign off at the end (reset job list); so
  // does the EXTRA instantiation (its done-bit packing is a full-warp ballot)
  if (!live && PM != P_PERENV && !EXTRA) return;

  // ---- phase B: state written by the previous launch
  double2 s = make_double2(0.0, 0.0);
  int2 yr = make_int2(0, 0);
  double2 Kv = make_double2(a.K, a.K), rv = make_double2(a.r, a.r);
  if (live) {
    s = reinterpret_cast<const double2*>(a.state_in)[j];
    yr = reinterpret_cast<const int2*>(a.years)[j];
    if (PM == P_PERENV) {
      Kv = reinterpret_cast<const double2*>(a.K_env)[j];
      rv = reinterpret_cast<const double2*>(a.r_env)[j];
    }
  }
  const int64_t i0 = 2 * (int64_t)j;
  int2 mdl = make_int2(0, 0);
  if (GROWTH == G_MODEL && live) mdl = reinterpret_cast<const int2*>(a.model)[j];
  EnvResult e0 = {}, e1 = {};   // idle lanes (fishing-v4 only, see above) never step and never reset
  if (live) {
    e0 = advance<GROWTH, PM, ACT, true, LEAN>(a, i0, s.x, act.x, z.x, noisy, yr.x, Kv.x, rv.x, mdl.x);
    e1 = advance<GROWTH, PM, ACT, true, LEAN>(a, i0 + 1, s.y, act.y, z.y, noisy, yr.y, Kv.y, rv.y, mdl.y);
  }
  // the (K, r) redraw of environments that reset in this step is done for the whole block at the end of the kernel
  const bool deferred = PM == P_PERENV && (a.flags & FISHING_AUTO_RESET) && (LEAN || !a.reset_noise);
  if (live) {
    reinterpret_cast<double2*>(a.out_obs)[j] = make_double2(e0.obs, e1.obs);
    if (!LEAN && a.out_obs2) st_stream(reinterpret_cast<double2*>(a.out_obs2) + j, make_double2(e0.obs, e1.obs));
    st_stream(reinterpret_cast<double2*>(a.out_reward) + j, make_double2(e0.reward, e1.reward));
    st_stream(reinterpret_cast<uchar2*>(a.out_done) + j, make_uchar2(e0.done, e1.done));
    reinterpret_cast<int2*>(a.years)[j] = make_int2(e0.years, e1.years);
    if (!LEAN && a.out_stock) reinterpret_cast<double2*>(a.out_stock)[j] = make_double2(e0.stock, e1.stock);
    if (!LEAN && a.out_harvest) reinterpret_cast<double2*>(a.out_harvest)[j] = make_double2(e0.harvest, e1.harvest);
    if (EXTRA) {
      if (a.ep_return) {
        double2 acc = reinterpret_cast<const double2*>(a.ep_return)[j];
        acc.x = __dadd_rn(acc.x, e0.reward);
        acc.y = __dadd_rn(acc.y, e1.reward);
        if (e0.done) { a.last_ep_return[i0] = acc.x; a.last_ep_length[i0] = yr.x + 1; acc.x = 0.0; }
        if (e1.done) { a.last_ep_return[i0 + 1] = acc.y; a.last_ep_length[i0 + 1] = yr.y + 1; acc.y = 0.0; }
        reinterpret_cast<double2*>(a.ep_return)[j] = acc;
      }
      if (a.out_obs32) st_stream(reinterpret_cast<float2*>(a.out_obs32) + j, make_float2((float)e0.obs, (float)e1.obs));
      if (a.out_reward32) st_stream(reinterpret_cast<float2*>(a.out_reward32) + j, make_float2((float)e0.reward, (float)e1.reward));
    }
    if (e0.reset_fired | e1.reset_fired) {
      if (!LEAN && a.out_terminal) {
        if (e0.reset_fired) a.out_terminal[i0] = e0.terminal;
        if (e1.reset_fired) a.out_terminal[i0 + 1] = e1.terminal;
      }
      if (PM == P_PERENV && !deferred) {
        reinterpret_cast<double2*>(a.K_env)[j] = make_double2(e0.K, e1.K);
        reinterpret_cast<double2*>(a.r_env)[j] = make_double2(e0.r, e1.r);
      }
      if (GROWTH == G_MODEL) reinterpret_cast<int2*>(a.model)[j] = make_int2(e0.model, e1.model);
    }
  }
  if (EXTRA && a.out_done_bits) {
    // 64 environments per warp -> one 64-bit word: even environments from one ballot, odd ones from the other, interleaved
    const unsigned full = 0xffffffffu;
    const unsigned m0 = __ballot_sync(full, live && e0.done), m1 = __ballot_sync(full, live && e1.done);
    if ((threadIdx.x & 31) == 0 && live) {
      auto spread = [](unsigned m) {
        unsigned long long x = m;
        x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
        x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
        x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
        x = (x | (x << 2)) & 0x3333333333333333ull;
        x = (x | (x << 1)) & 0x5555555555555555ull;
        return x;
      };
      a.out_done_bits[j >> 5] = spread(m0) | (spread(m1) << 1);
    }
  }
  if (deferred) {
    // Block-wide parameter redraw without a barrier.  A reset needs a Philox block + Box-Muller (~200 instructions), and
    // with 64 environments per warp nearly every warp has at least one reset per step, so a per-warp pass runs at ~20 %
    // lane utilisation.  Instead every warp appends its resets to a job list in shared memory and signs off; the warp that
    // signs off LAST draws for the whole block (ceil(jobs / 32) passes instead of one per warp) and writes the new K and r
    // straight to the state arrays -- nobody waits for anybody, the other warps have already retired.  Same draws as the
    // per-lane path: ctr = (global env id, step), stream AUTORESET.
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned m0 = __ballot_sync(full, e0.reset_fired), m1 = __ballot_sync(full, e1.reset_fired);
    const int n0 = __popc(m0), mine = n0 + __popc(m1);
    if (mine > 0) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&s_count, (unsigned)mine);
      base = __shfl_sync(full, base, 0);
      const unsigned below = (1u << lane) - 1u;
      if (e0.reset_fired) s_job[base + __popc(m0 & below)] = (uint16_t)(2 * threadIdx.x);
      if (e1.reset_fired) s_job[base + n0 + __popc(m1 & below)] = (uint16_t)(2 * threadIdx.x + 1);
    }
    __threadfence_block();
    unsigned arrived = 0;
    if (lane == 0) arrived = atomicAdd(&s_arrived, 1u);
    arrived = __shfl_sync(full, arrived, 0);
    if (arrived == BS / 32 - 1) {
      __threadfence_block();
      const int total = (int)*reinterpret_cast<volatile unsigned*>(&s_count);
      const int64_t block0 = 2 * (int64_t)(blockIdx.x * (unsigned)BS);
#ifdef FISHING_AB_V4_ILP2      // timing build: two jobs per lane and pass (two independent Philox + Box-Muller chains)
      for (int k = lane; k < total; k += 64) {
        const bool two = k + 32 < total;
        const int64_t ia = block0 + reinterpret_cast<volatile uint16_t*>(s_job)[k];
        const int64_t ib = two ? block0 + reinterpret_cast<volatile uint16_t*>(s_job)[k + 32] : ia;
        const Normal2 da = normal_pair_keys(a.keys, (uint64_t)(a.env_id0 + ia), a.step_idx, FISHING_STREAM_AUTORESET);
        const Normal2 db = normal_pair_keys(a.keys, (uint64_t)(a.env_id0 + ib), a.step_idx, FISHING_STREAM_AUTORESET);
        a.K_env[ia] = v4_param(a.K_mean, a.sigma_p, da.z0);
        a.r_env[ia] = v4_param(a.r_mean, a.sigma_p, da.z1);
        if (two) {
          a.K_env[ib] = v4_param(a.K_mean, a.sigma_p, db.z0);
          a.r_env[ib] = v4_param(a.r_mean, a.sigma_p, db.z1);
        }
      }
      if (false)
#endif
      for (int k = lane; k < total; k += 32) {
        const int64_t i = block0 + reinterpret_cast<volatile uint16_t*>(s_job)[k];
#ifdef FISHING_AB_V4_NODRAW      // timing build: what do the draws cost (the scattered writes stay)
        Normal2 d;
        d.z0 = d.z1 = (double)k * 1e-3;
#else
        const Normal2 d = normal_pair_keys(a.keys, (uint64_t)(a.env_id0 + i), a.step_idx, FISHING_STREAM_AUTORESET);
#endif
#ifndef FISHING_AB_V4_NOWRITE    // timing build: what do the scattered 8-byte writes cost (the draws stay)
        a.K_env[i] = v4_param(a.K_mean, a.sigma_p, d.z0);
        a.r_env[i] = v4_param(a.r_mean, a.sigma_p, d.z1);
#else
        if (d.z0 == 1.2345e300) a.K_env[i] = d.z1;
#endif
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ K1, scalar path
// Same results for odd n / odd env_id0 / unaligned slices: one environment per thread, which recomputes its pair's
// Philox block and picks its half.
template <int GROWTH, int PM, int ACT>
__global__ void __launch_bounds__(256) step_kernel_scalar(const StepArgs a) {
  using AT = typename ActT<ACT>::type;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool ext_noise = a.noise != nullptr;
  const bool noisy = ext_noise || a.sigma != 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const uint64_t gid = (uint64_t)(a.env_id0 + i);
    double z = 0.0;
    if (ext_noise) {
      z = a.noise[i];
    } else if (noisy) {
      const Normal2 d = normal_pair(a.seed, gid >> 1, a.step_idx, FISHING_STREAM_STEP);
      z = (gid & 1) ? d.z1 : d.z0;
    }
    const double K = PM == P_PERENV ? a.K_env[i] : a.K;
    const double r = PM == P_PERENV ? a.r_env[i] : a.r;
    const EnvResult e = advance<GROWTH, PM, ACT>(a, i, a.state_in[i], reinterpret_cast<const AT*>(a.action)[i], z, noisy,
                                                 a.years[i], K, r, GROWTH == G_MODEL ? a.model[i] : 0);
    a.out_obs[i] = e.obs;
    if (a.out_obs2) a.out_obs2[i] = e.obs;
    a.out_reward[i] = e.reward;
    a.out_done[i] = e.done;
    const int32_t years_in = a.years[i];
    a.years[i] = e.years;
    if (a.out_stock) a.out_stock[i] = e.stock;
    if (a.out_harvest) a.out_harvest[i] = e.harvest;
    if (a.ep_return) {
      double acc = __dadd_rn(a.ep_return[i], e.reward);
      if (e.done) { a.last_ep_return[i] = acc; a.last_ep_length[i] = years_in + 1; acc = 0.0; }
      a.ep_return[i] = acc;
    }
    if (a.out_obs32) a.out_obs32[i] = (float)e.obs;
    if (a.out_reward32) a.out_reward32[i] = (float)e.reward;
    if (e.reset_fired) {
      if (a.out_terminal) a.out_terminal[i] = e.terminal;
      if (PM == P_PERENV) {
        a.K_env[i] = e.K;
        a.r_env[i] = e.r;
      }
      if (GROWTH == G_MODEL) a.model[i] = e.model;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K3: reset
__global__ void __launch_bounds__(256) reset_kernel(const ResetArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    if (a.mask && !a.mask[i]) {
      if (a.out_obs && a.out_obs != a.state) a.out_obs[i] = a.state[i];
      continue;
    }
    if (a.per_env) {
      double zK, zr;
      if (a.reset_noise) {
        zK = a.reset_noise[i];
        zr = a.reset_noise[a.n + i];
      } else {
        const Normal2 d = normal_pair(a.seed, (uint64_t)(a.env_id0 + i), a.epoch, FISHING_STREAM_RESET);
        zK = d.z0;
        zr = d.z1;
      }
      a.K_env[i] = v4_param(a.K_mean, a.sigma_p, zK);
      a.r_env[i] = v4_param(a.r_mean, a.sigma_p, zr);
    }
    if (a.n_models > 0)
      a.model[i] = a.reset_model ? a.reset_model[i]
                                 : (int32_t)uniform_index(a.seed, (uint64_t)(a.env_id0 + i), a.epoch, FISHING_STREAM_RESET,
                                                          (uint32_t)a.n_models);
    a.state[i] = a.s_reset;
    a.years[i] = 0;
    if (a.out_obs) a.out_obs[i] = a.s_reset;
  }
}

// ------------------------------------------------------------------------------------------------ noise exporter
__global__ void __launch_bounds__(256) noise_fill_kernel(uint64_t seed, int32_t stream_id, int64_t n, int64_t env_id0,
                                                          uint64_t step0, int64_t T, double* __restrict__ out) {
  const int64_t total = n * T;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
    const int64_t t = k / n, i = k - t * n;
    const uint64_t gid = (uint64_t)(env_id0 + i);
    if (stream_id == FISHING_STREAM_STEP) {
      const Normal2 d = normal_pair(seed, gid >> 1, step0 + (uint64_t)t, FISHING_STREAM_STEP);
      out[k] = (gid & 1) ? d.z1 : d.z0;
    } else {
      const Normal2 d = normal_pair(seed, gid, step0 + (uint64_t)t, (uint32_t)stream_id);
      out[(t * 2 + 0) * n + i] = d.z0;
      out[(t * 2 + 1) * n + i] = d.z1;
    }
  }
}

// ------------------------------------------------------------------------------------------------ BMSY
// policies.py:51-67.  The reference evaluates population_draw ONCE on the whole float32 grid, so for v0/v1/v2 every
// operation is a float32 operation (K, r, C are Python numbers and take the array's dtype); for v4 K and r are
// np.float64 scalars and the grid is promoted to float64.  One block; argmax keeps the FIRST maximum like np.argmax.
template <typename F>
__device__ __forceinline__ F f_exp(F v);
// numpy's float32 exp (numpy/_core/src/umath/loops_exponent_log.dispatch.c.src, simd_exp_FLOAT; the AVX2+FMA3 and
// AVX512F builds share it): Cody-Waite reduction by ln2 with FMAs, a (5,2) rational minimax P(r)/Q(r), scale by 2^n.
// It is NOT correctly rounded (it differs from the rounded true value on ~40 % of inputs), and fishing-v2's BMSY is
// an argmax over float32 values that tie to the last bit, so the reference's S = 0x1.9374bcp-1 is only reproduced by
// reproducing this algorithm operation for operation.  tests/test_np_expf.py pins the restatement to np.exp itself.
__device__ __forceinline__ float np_expf(float x) {
  if (x != x) return x;
  if (x > 88.72283935546875f) return INFINITY;
  if (x < -103.97208404541015625f) return 0.0f;
  const float magic = 12582912.0f;  // 0x1.8p23: add/subtract rounds to the nearest integer
  float q = __fmul_rn(x, 1.442695040888963407359924681001892137f);
  q = __fsub_rn(__fadd_rn(q, magic), magic);
  float r = __fmaf_rn(q, -6.93145752e-1f, x);
  r = __fmaf_rn(q, -1.42860677e-6f, r);
  float num = __fmaf_rn(5.082762527590693718096e-04f, r, 6.757896990527504603057e-03f);
  num = __fmaf_rn(num, r, 5.114512081637298353406e-02f);
  num = __fmaf_rn(num, r, 2.473615434895520810817e-01f);
  num = __fmaf_rn(num, r, 7.257664613233124478488e-01f);
  num = __fmaf_rn(num, r, 9.999999999980870924916e-01f);
  float den = __fmaf_rn(2.159509375685829852307e-02f, r, -2.742335390411667452936e-01f);
  den = __fmaf_rn(den, r, 1.0f);
  return ldexpf(__fdiv_rn(num, den), (int)q);
}
template <>
__device__ __forceinline__ float f_exp<float>(float v) { return np_expf(v); }
template <>
__device__ __forceinline__ double f_exp<double>(double v) { return exp(v); }

// population_draw on a value of dtype F (float32 when the reference's array is float32 and r, K, C, sigma, z are
// Python numbers, which then take the array's dtype under numpy's promotion rules).
template <typename F>
__device__ __forceinline__ F growth_f(int kind, F x0, F r, F K, F C, F sigma, F z) {
  // float and double arithmetic below is written with the non-contracting intrinsics of the matching width
  auto mul = [](F p, F q) -> F { if constexpr (sizeof(F) == 4) return __fmul_rn(p, q); else return __dmul_rn(p, q); };
  auto add = [](F p, F q) -> F { if constexpr (sizeof(F) == 4) return __fadd_rn(p, q); else return __dadd_rn(p, q); };
  auto sub = [](F p, F q) -> F { if constexpr (sizeof(F) == 4) return __fsub_rn(p, q); else return __dsub_rn(p, q); };
  auto div = [](F p, F q) -> F { if constexpr (sizeof(F) == 4) return __fdiv_rn(p, q); else return __ddiv_rn(p, q); };
  const F room = sub((F)1, div(x0, K));
  const F noise = mul(mul(x0, sigma), z);
  F g;
  if (kind == FISHING_V2) {
    const F e = add(mul(mul(r, room), sub(x0, C)), noise);
    g = mul(x0, f_exp<F>(e));
  } else {
    g = add(add(x0, mul(mul(r, x0), room)), noise);
  }
  return (g > (F)0 || g != g) ? g : (F)0;
}

template <typename F>
__device__ __forceinline__ F growth_minus_x(int kind, F x0, F r, F K, F C) {
  const F g = growth_f<F>(kind, x0, r, K, C, (F)0, (F)0);  // sigma forced to 0, policies.py:61-63
  if constexpr (sizeof(F) == 4) return __fsub_rn(g, x0);
  else return __dsub_rn(g, x0);
}

// Stand-alone BaseFishingEnv.population_draw on an array (one scalar z for the whole array, as the reference draws it).
template <typename F>
__global__ void __launch_bounds__(256) growth_kernel(int kind, int64_t n, const F* __restrict__ x, double r, double K,
                                                      double C, double sigma, double z, F* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = growth_f<F>(kind, x[i], (F)r, (F)K, (F)C, (F)sigma, (F)z);
}

// np.linspace(float32(-1), float32(1), num, dtype=float32) evaluates arange(num, dtype=f32) * (f32(2) / (num-1)) +
// f32(-1) entirely in float32 and then overwrites the last point with the stop value (numpy/_core/function_base.py).
__device__ __forceinline__ float linspace_f32(int i, int num) {
  if (i == num - 1) return 1.0f;
  return __fadd_rn(__fmul_rn((float)i, __fdiv_rn(2.0f, (float)(num - 1))), -1.0f);
}

template <typename F>
__global__ void __launch_bounds__(1024) bmsy_kernel(int kind, double r_, double K_, double C_, double* out) {
  constexpr int NGRID = 10001;
  __shared__ double best_v[1024];
  __shared__ int best_i[1024];
  const F r = (F)r_, K = (F)K_, C = (F)C_;
  double bv = -INFINITY;
  int bi = NGRID;
  bool nan_seen = false;
  for (int i = threadIdx.x; i < NGRID; i += blockDim.x) {
    const float g32 = linspace_f32(i, NGRID);
    F x0;
    if constexpr (sizeof(F) == 4) x0 = __fmul_rn(__fadd_rn(g32, 1.0f), K);        // (np.float32 + 1) * K, float32
    else x0 = __dmul_rn((double)__fadd_rn(g32, 1.0f), K);                         // float32 sum, float64 product
    const double v = (double)growth_minus_x<F>(kind, x0, r, K, C);
    if (v != v) {  // np.argmax returns the first NaN
      if (!nan_seen || i < bi) { bi = nan_seen ? min(bi, i) : i; nan_seen = true; bv = INFINITY; }
    } else if (!nan_seen && (v > bv || (v == bv && i < bi))) {
      bv = v;
      bi = i;
    }
  }
  best_v[threadIdx.x] = bv;
  best_i[threadIdx.x] = bi;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      const double ov = best_v[threadIdx.x + w];
      const int oi = best_i[threadIdx.x + w];
      if (ov > best_v[threadIdx.x] || (ov == best_v[threadIdx.x] && oi < best_i[threadIdx.x])) {
        best_v[threadIdx.x] = ov;
        best_i[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int i = best_i[0];
    const float g32 = linspace_f32(i, NGRID);
    F S;
    if constexpr (sizeof(F) == 4) S = __fmul_rn(__fadd_rn(g32, 1.0f), K);
    else S = __dmul_rn((double)__fadd_rn(g32, 1.0f), K);
    out[0] = (double)S;
    out[1] = (double)growth_minus_x<F>(kind, S, r, K, C);  // msy.__init__, policies.py:9-12
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static int grid_for(int64_t work_items, int threads, int max_blocks_per_sm) {
  const int64_t need = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * max_blocks_per_sm;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

#ifndef FISHING_K1_BS
#define FISHING_K1_BS 256      // threads per block of the lock-step kernel (timing builds: 128 / 512)
#endif
constexpr int K1_BS_DEFAULT = FISHING_K1_BS;
#ifndef FISHING_K1_BS_V4
// per-environment parameters (fishing-v4): 128-thread blocks.  The warp that signs off last redraws (K, r) for the block's resets
// while the block's registers stay allocated; with ~10 % of the environments resetting per step a 256-thread block needs two 32-job
// passes, a 128-thread block one (cfg 4, random actions: 8.41e10 -> 8.63e10 env-steps/s; 64 threads: 8.74e10 but -4 % at ~1 %
// resets; two jobs per lane and pass instead: no gain -- profiles/ab_r02.md).  Also measured and NOT kept (r02o / r02p there): a
// persistent kernel that walks per-warp tiles with the next tile's loads in flight (in registers, or staged with cp.async in
// per-lane shared-memory slots) and redraws from a warp-private queue in full 32-lane passes -- bit-identical, but 7.1 - 7.3e10
// against 8.7e10 env-steps/s, and slower even without resets (1.00e11 against 1.12e11).
#define FISHING_K1_BS_V4 128
#endif

template <int GROWTH, int PM, int ACT>
static cudaError_t launch_step_t(const StepArgs& a, bool vec, cudaStream_t st) {
  constexpr bool CAN_LEAN = true;
  constexpr int K1_BS = PM == P_PERENV ? FISHING_K1_BS_V4 : K1_BS_DEFAULT;
  const bool extra = a.ep_return || a.out_obs32 || a.out_reward32 || a.out_done_bits;
  const bool lean = CAN_LEAN && !extra && !a.noise && !a.reset_noise && !a.out_obs2 && !a.out_stock && !a.out_harvest && !a.out_terminal;
  if (vec) {
    const int64_t npairs = a.n >> 1;
    const int64_t blocks = (npairs + K1_BS - 1) / K1_BS;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (a.flags & STEP_FLAG_PDL) {
      // programmatic dependent launch: this grid may start while the previous kernel in the stream is still draining
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)blocks);
      cfg.blockDim = dim3(K1_BS);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      count_launch();
      if (lean) return cudaLaunchKernelEx(&cfg, step_kernel_vec2<GROWTH, PM, ACT, K1_BS, CAN_LEAN>, a);
      if (extra) return cudaLaunchKernelEx(&cfg, step_kernel_vec2<GROWTH, PM, ACT, K1_BS, false, true>, a);
      return cudaLaunchKernelEx(&cfg, step_kernel_vec2<GROWTH, PM, ACT, K1_BS, false>, a);
    }
    if (lean) step_kernel_vec2<GROWTH, PM, ACT, K1_BS, CAN_LEAN><<<(unsigned)blocks, K1_BS, 0, st>>>(a);
    else if (extra) step_kernel_vec2<GROWTH, PM, ACT, K1_BS, false, true><<<(unsigned)blocks, K1_BS, 0, st>>>(a);
    else step_kernel_vec2<GROWTH, PM, ACT, K1_BS, false><<<(unsigned)blocks, K1_BS, 0, st>>>(a);
  } else {
    const int grid = grid_for(a.n, 256, 8);
    step_kernel_scalar<GROWTH, PM, ACT><<<grid, 256, 0, st>>>(a);
  }
  count_launch();
  return cudaGetLastError();
}

template <int GROWTH, int PM>
static cudaError_t launch_step_a(const StepArgs& a, int act, bool vec, cudaStream_t st) {
  switch (act) {
    case FISHING_ACT_F64: return launch_step_t<GROWTH, PM, FISHING_ACT_F64>(a, vec, st);
    case FISHING_ACT_F32: return launch_step_t<GROWTH, PM, FISHING_ACT_F32>(a, vec, st);
    case FISHING_ACT_I32: return launch_step_t<GROWTH, PM, FISHING_ACT_I32>(a, vec, st);
    default: return launch_step_t<GROWTH, PM, FISHING_ACT_I64>(a, vec, st);
  }
}

// the log-normal family has Box actions only and always uses the general-K state scaling
template <int GROWTH>
static cudaError_t launch_step_ln(const StepArgs& a, int act, bool vec, cudaStream_t st) {
  if (act == FISHING_ACT_F32) return launch_step_t<GROWTH, P_UNIFORM, FISHING_ACT_F32>(a, vec, st);
  return launch_step_t<GROWTH, P_UNIFORM, FISHING_ACT_F64>(a, vec, st);
}

cudaError_t launch_step(const StepArgs& a, int kind, int act, bool vec, cudaStream_t st) {
  switch (kind) {
    case FISHING_V5: return launch_step_ln<G_ALLEN>(a, act, vec, st);
    case FISHING_V6: case FISHING_V10: return launch_step_ln<G_BEVERTON_HOLT>(a, act, vec, st);
    case FISHING_V7: return launch_step_ln<G_MAY>(a, act, vec, st);
    case FISHING_V8: return launch_step_ln<G_MYERS>(a, act, vec, st);
    case FISHING_V9: return launch_step_ln<G_RICKER>(a, act, vec, st);
    case FISHING_V11: return launch_step_ln<G_MODEL>(a, act, vec, st);
    default: break;
  }
  if (kind == FISHING_V4) return launch_step_a<LOGISTIC, P_PERENV>(a, act, vec, st);
  const bool kone = a.K == 1.0;
  if (kind == FISHING_V2)
    return kone ? launch_step_a<TIPPING, P_KONE>(a, act, vec, st) : launch_step_a<TIPPING, P_UNIFORM>(a, act, vec, st);
  return kone ? launch_step_a<LOGISTIC, P_KONE>(a, act, vec, st) : launch_step_a<LOGISTIC, P_UNIFORM>(a, act, vec, st);
}

cudaError_t launch_reset(const ResetArgs& a, cudaStream_t st) {
  reset_kernel<<<grid_for(a.n, 256, 8), 256, 0, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_noise_fill(uint64_t seed, int32_t stream_id, int64_t n, int64_t env_id0, uint64_t step0, int64_t T,
                              double* out, cudaStream_t st) {
  noise_fill_kernel<<<grid_for(n * T, 256, 8), 256, 0, st>>>(seed, stream_id, n, env_id0, step0, T, out);
  count_launch();
  return cudaGetLastError();
}

// stand-alone population_draw of the log-normal family on an array, ONE draw per element like np.random.lognormal(mu)
__global__ void __launch_bounds__(256) growth_ln_kernel(int fn, int64_t n, const double* __restrict__ x, GrowthP gp,
                                                         int derive_on_device, const double* __restrict__ z,
                                                         double* __restrict__ out) {
  // fishing-v10: log(A) and B follow the drifting r; derived with the DEVICE's log, like K1 and K2 do (StepArgs::derive_on_device)
  if (derive_on_device) growth_derive(gp, fn);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double zi = z ? z[i] : 0.0;
    double v;
    switch (fn) {    // kernel-uniform: the same device functions as the single-model K1 / K2 instantiations (bit-identical)
      case FISHING_FN_ALLEN: v = grow_lognormal<FISHING_FN_ALLEN>(x[i], gp, zi, true); break;
      case FISHING_FN_BEVERTON_HOLT: v = grow_lognormal<FISHING_FN_BEVERTON_HOLT>(x[i], gp, zi, true); break;
      case FISHING_FN_MYERS: v = grow_lognormal<FISHING_FN_MYERS>(x[i], gp, zi, true); break;
      case FISHING_FN_MAY: v = grow_lognormal<FISHING_FN_MAY>(x[i], gp, zi, true); break;
      default: v = grow_lognormal<FISHING_FN_RICKER>(x[i], gp, zi, true); break;
    }
    out[i] = v;
  }
}

cudaError_t launch_growth_ln(int fn, int64_t n, const double* x, const GrowthP& gp, bool derive_on_device, const double* z,
                             double* out, cudaStream_t st) {
  growth_ln_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(fn, n, x, gp, derive_on_device ? 1 : 0, z, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_growth(int kind, int64_t n, const void* x, bool f32, double r, double K, double C, double sigma,
                          double z, void* out, cudaStream_t st) {
  const int grid = grid_for(n, 256, 8);
  if (f32) growth_kernel<float><<<grid, 256, 0, st>>>(kind, n, (const float*)x, r, K, C, sigma, z, (float*)out);
  else growth_kernel<double><<<grid, 256, 0, st>>>(kind, n, (const double*)x, r, K, C, sigma, z, (double*)out);
  count_launch();
  return cudaGetLastError();
}

// FP64-pipe peak probe (measurement aid for the K2 roofline, SURVEY.md section 8d "measure with a DFMA loop"): every
// thread runs `iters` rounds of 8 independent DFMA chains; 16 warps per SM sub-partition keep the pipe saturated.
// Total DFMA warp-instructions = blocks * 8 warps * iters * 8; the caller times the launch with events.
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, double a, double b, int iters) {
  double x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = a + (double)(c + (int)threadIdx.x);
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += x[c];
  if (s == 1.2345) out[threadIdx.x] = s;   // never true: keeps the chains alive
}

cudaError_t launch_fp64_probe(double* scratch, int blocks, int iters, cudaStream_t st) {
  fp64_probe_kernel<<<blocks, 256, 0, st>>>(scratch, 1.0000001, 1e-9, iters);
  count_launch();
  return cudaGetLastError();
}

// The table-assisted log / exp of csrc/normal.cuh evaluated over an array (test aid: tests/test_gpu_lognormal.py compares the
// device's results with numpy's and with the host build of the same header).
__global__ void __launch_bounds__(256) math_probe_kernel(int fn, int64_t n, const double* __restrict__ x, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = fn == 0 ? log_tab(x[i], GlobalTabs()) : exp_tab(x[i], GlobalTabs());
}

cudaError_t launch_math_probe(int fn, int64_t n, const double* x, double* out, cudaStream_t st) {
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  math_probe_kernel<<<blocks, 256, 0, st>>>(fn, n, x, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bmsy(int kind, double r, double K, double C, bool f64, double* out_dev, cudaStream_t st) {
  if (f64) bmsy_kernel<double><<<1, 1024, 0, st>>>(kind, r, K, C, out_dev);
  else bmsy_kernel<float><<<1, 1024, 0, st>>>(kind, r, K, C, out_dev);
  count_launch();
  return cudaGetLastError();
}

}  // namespace fishing
