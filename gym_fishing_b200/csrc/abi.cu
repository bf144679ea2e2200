// extern "C" boundary of libfishing_b200.so (declared in include/fishing_b200.h): argument validation, dispatch to the
// kernels, and the host-buffer front end (fishing_vec_*) with its chunked three-stream copy/compute pipeline.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/fishing_b200.h"
#include "common.cuh"

#include "kernel_args.cuh"

namespace fishing {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

static bool kind_ok(int kind) {
  return kind == FISHING_V0 || kind == FISHING_V1 || kind == FISHING_V2 || kind == FISHING_V4 ||
         (kind >= FISHING_V5 && kind <= FISHING_V11);
}
static bool is_lognormal(int kind) { return kind >= FISHING_V5 && kind <= FISHING_V11; }

#ifdef FISHING_AB_LIBLOG
constexpr bool kDeriveOnDevice = true;    // timing build with the library's log: host and device logs differ in the last bit
#else
constexpr bool kDeriveOnDevice = false;
#endif

static int growth_fn_of_kind(int kind) {
  switch (kind) {
    case FISHING_V5: return FISHING_FN_ALLEN;
    case FISHING_V6: case FISHING_V10: return FISHING_FN_BEVERTON_HOLT;
    case FISHING_V7: return FISHING_FN_MAY;
    case FISHING_V8: return FISHING_FN_MYERS;
    case FISHING_V9: return FISHING_FN_RICKER;
    default: return -1;
  }
}

static GrowthP growth_of(const fishing_params* p) {
  GrowthP g;
  g.r = p->r; g.K = p->growth_K; g.C = p->C; g.M = p->M; g.theta = p->theta; g.q = p->q; g.b = p->b; g.a = p->a;
  g.sigma = p->sigma;
  growth_derive(g, growth_fn_of_kind(p->kind));
  return g;
}

static int fill_table(const fishing_params* p, ModelTable* t, double* any_sigma) {
  std::memset(t, 0, sizeof(*t));
  *any_sigma = 0.0;
  if (p->kind != FISHING_V11) return 0;
  if (p->n_models < 1 || p->n_models > FISHING_MAX_MODELS) return FISHING_E_ARG;
  t->n_models = p->n_models;
  for (int i = 0; i < p->n_models; ++i) {
    if (p->model_fn[i] < FISHING_FN_ALLEN || p->model_fn[i] > FISHING_FN_RICKER) return FISHING_E_ARG;
    t->fn[i] = p->model_fn[i];
    const fishing_growth_params& m = p->model_params[i];
    t->p[i].r = m.r; t->p[i].K = m.K; t->p[i].C = m.C; t->p[i].M = m.M; t->p[i].theta = m.theta; t->p[i].q = m.q;
    t->p[i].b = m.b; t->p[i].a = m.a; t->p[i].sigma = m.sigma;
    growth_derive(t->p[i], t->fn[i]);
    if (m.sigma != 0.0) *any_sigma = 1.0;
  }
  return 0;
}

static int act_size(int dt) {
  switch (dt) {
    case FISHING_ACT_F64: return 8;
    case FISHING_ACT_F32: return 4;
    case FISHING_ACT_I32: return 4;
    case FISHING_ACT_I64: return 8;
    default: return 0;
  }
}

static bool act_matches_kind(int kind, int dt) {
  const bool discrete = dt == FISHING_ACT_I32 || dt == FISHING_ACT_I64;
  return (kind == FISHING_V0) == discrete;
}

// x0 / K - 1, or x0 un-normalised for fishing-v4 (fishing_model_error.py:44).  Host IEEE double arithmetic; this
// translation unit is compiled without -ffast-math, and a division followed by a subtraction cannot be contracted.
static double reset_state_value(const fishing_params* p) {
  if (p->kind == FISHING_V4) return p->x0;
  volatile double q = p->x0 / p->K;
  return q - 1.0;
}

// FISHING_NO_PDL=1 in the environment switches programmatic dependent launch off (for A/B timing and debugging)
static bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("FISHING_NO_PDL");
    return !(e && e[0] && e[0] != '0');
  }();
  return on;
}

// K1's phase-A prefetch of the state lines: on when one launch's state (state + years, + K and r for fishing-v4) is larger
// than ~2/3 of the 126 MB L2, i.e. when consecutive lock-steps find it in DRAM.  FISHING_K1_PREFETCH=0 / 1 forces it.
static bool k1_prefetch(int64_t state_bytes) {
  static const int forced = [] {
    const char* e = std::getenv("FISHING_K1_PREFETCH");
    return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
  }();
  if (forced >= 0) return forced == 1;
  return state_bytes >= (int64_t)80 << 20;
}

// device-side address of a page-locked host allocation (NULL if the pointer is ordinary pageable memory)
static void* device_alias(const void* host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return at.devicePointer;
}

// FISHING_VEC_ZEROCOPY=1 opts in to the zero-copy path of the host front end.  Measured on the B200 box it is not faster
// than the staged copy-engine pipeline (2.46e9 vs 2.54e9 env-steps/s): both sit at the PCIe limit of ~45 GB/s per
// direction under bidirectional load, so the staged path stays the default.
static bool zero_copy_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("FISHING_VEC_ZEROCOPY");
    return e && e[0] == '1';
  }();
  return on;
}

// largest batch that fishing_vec_step serves with its one-launch zero-copy route (FISHING_VEC_ZC_MAX overrides)
static int64_t zero_copy_max_envs() {
  static const int64_t n = [] {
    const char* e = std::getenv("FISHING_VEC_ZC_MAX");
    return e ? (int64_t)std::atoll(e) : (int64_t)65536;
  }();
  return n;
}

static bool overlaps(const void* a, size_t na, const void* b, size_t nb) {
  if (!a || !b) return false;
  const uintptr_t x = reinterpret_cast<uintptr_t>(a), y = reinterpret_cast<uintptr_t>(b);
  return x < y + nb && y < x + na;
}

static int check_soa(const fishing_params* p, const fishing_soa& s) {
  if (!s.state || !s.years) return FISHING_E_ARG;
  if (p->kind == FISHING_V4 && (!s.K_env || !s.r_env)) return FISHING_E_ARG;
  if (!aligned(s.state, 8) || !aligned(s.years, 4)) return FISHING_E_ALIGN;
  if (p->kind == FISHING_V4 && (!aligned(s.K_env, 8) || !aligned(s.r_env, 8))) return FISHING_E_ALIGN;
  if (p->kind == FISHING_V11 && !s.model) return FISHING_E_ARG;
  if (p->kind == FISHING_V11 && !aligned(s.model, 4)) return FISHING_E_ALIGN;
  return 0;
}

}  // namespace fishing

using namespace fishing;

extern "C" {

int fishing_abi_version(void) { return FISHING_ABI_VERSION; }

// hash of the sources this library was compiled from (gym_fishing_b200/build.py:source_id); the marker lets the build
// script read it from the file without loading the library
#ifndef FISHING_BUILD_ID_STR
#define FISHING_BUILD_ID_STR "unknown"
#endif
#ifdef FISHING_BUILD_SELFTEST_FAIL   // tests/test_cpu_host.py: a compile error must make the build fail loudly
#error "deliberate compile error (FISHING_BUILD_SELFTEST_FAIL)"
#endif
const char* fishing_build_id(void) {
  static const char id[] = "FISHING_BUILD_ID=" FISHING_BUILD_ID_STR;
  return id + sizeof("FISHING_BUILD_ID=") - 1;
}

uint64_t fishing_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int fishing_probe_fp64(double* scratch, int32_t blocks, int32_t iters, void* stream) {
  if (!scratch || blocks <= 0 || iters <= 0) return FISHING_E_ARG;
  return (int)launch_fp64_probe(scratch, blocks, iters, (cudaStream_t)stream);
}

int fishing_probe_math(int32_t fn, int64_t n, const double* x, double* out, void* stream) {
  if (n < 0 || (fn != 0 && fn != 1) || (n > 0 && (!x || !out))) return FISHING_E_ARG;
  if (n == 0) return 0;
  return (int)launch_math_probe(fn, n, x, out, (cudaStream_t)stream);
}

const char* fishing_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case FISHING_E_ARG: return "fishing: bad argument (NULL pointer, negative size or unknown enum)";
    case FISHING_E_KIND: return "fishing: action dtype / policy not valid for this environment id";
    case FISHING_E_ALIGN: return "fishing: pointer not aligned to its element size";
    case FISHING_E_NODEVICE: return "fishing: no usable CUDA device";
    case FISHING_E_STATE: return "fishing: handle not reset yet or already destroyed";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "fishing: unknown error";
}

static int reset_impl(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t epoch, const uint8_t* mask,
                      const double* reset_noise, const int32_t* reset_model, fishing_soa soa, double* out_obs, void* stream);

int fishing_reset(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t epoch, const uint8_t* mask,
                  const double* reset_noise, fishing_soa soa, double* out_obs, void* stream) {
  return reset_impl(p, n, env_id0, epoch, mask, reset_noise, nullptr, soa, out_obs, stream);
}

int fishing_reset_models(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t epoch, const uint8_t* mask,
                         const int32_t* reset_model, fishing_soa soa, double* out_obs, void* stream) {
  if (reset_model && !aligned(reset_model, 4)) return FISHING_E_ALIGN;
  return reset_impl(p, n, env_id0, epoch, mask, nullptr, reset_model, soa, out_obs, stream);
}

static int reset_impl(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t epoch, const uint8_t* mask,
                      const double* reset_noise, const int32_t* reset_model, fishing_soa soa, double* out_obs, void* stream) {
  if (!p || n < 0 || env_id0 < 0 || !kind_ok(p->kind)) return FISHING_E_ARG;
  if (int e = check_soa(p, soa)) return e;
  if (out_obs && !aligned(out_obs, 8)) return FISHING_E_ALIGN;
  if (n == 0) return 0;
  ResetArgs a;
  a.n = n; a.env_id0 = env_id0; a.epoch = epoch; a.seed = p->seed;
  a.s_reset = reset_state_value(p);
  a.K_mean = p->K; a.r_mean = p->r; a.sigma_p = p->sigma_p;
  a.per_env = p->kind == FISHING_V4;
  a.n_models = 0; a.reset_model = reset_model; a.model = soa.model;
  if (p->kind == FISHING_V11) {
    if (p->n_models < 1 || p->n_models > FISHING_MAX_MODELS) return FISHING_E_ARG;
    a.n_models = p->n_models;
  }
  a.mask = mask; a.reset_noise = reset_noise;
  a.state = soa.state; a.years = soa.years; a.K_env = soa.K_env; a.r_env = soa.r_env; a.out_obs = out_obs;
  return (int)launch_reset(a, static_cast<cudaStream_t>(stream));
}

static int step_impl(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step_idx, const void* action,
                     int32_t action_dtype, const double* noise, const double* reset_noise, fishing_soa soa, double* out_obs,
                     double* out_reward, uint8_t* out_done, double* out_terminal, double* out_stock, double* out_harvest,
                     double* out_obs2, void* stream, const fishing_step_extra* x = nullptr);

int fishing_step(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step_idx, const void* action,
                 int32_t action_dtype, const double* noise, const double* reset_noise, fishing_soa soa, double* out_obs,
                 double* out_reward, uint8_t* out_done, double* out_terminal, double* out_stock, double* out_harvest,
                 void* stream) {
  return step_impl(p, n, env_id0, step_idx, action, action_dtype, noise, reset_noise, soa, out_obs, out_reward, out_done,
                   out_terminal, out_stock, out_harvest, nullptr, stream);
}

int fishing_step_ex(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step_idx, const void* action,
                    int32_t action_dtype, const double* noise, const double* reset_noise, fishing_soa soa, double* out_obs,
                    double* out_reward, uint8_t* out_done, double* out_terminal, const fishing_step_extra* x, void* stream) {
  return step_impl(p, n, env_id0, step_idx, action, action_dtype, noise, reset_noise, soa, out_obs, out_reward, out_done,
                   out_terminal, nullptr, nullptr, nullptr, stream, x);
}

static int step_impl(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step_idx, const void* action,
                     int32_t action_dtype, const double* noise, const double* reset_noise, fishing_soa soa, double* out_obs,
                     double* out_reward, uint8_t* out_done, double* out_terminal, double* out_stock, double* out_harvest,
                     double* out_obs2, void* stream, const fishing_step_extra* x) {
  if (!p || n < 0 || env_id0 < 0 || !kind_ok(p->kind) || !action || !out_obs || !out_reward || !out_done)
    return FISHING_E_ARG;
  if (x) {
    const int have = (x->ep_return != nullptr) + (x->last_ep_return != nullptr) + (x->last_ep_length != nullptr);
    if (have != 0 && have != 3) return FISHING_E_ARG;
    if ((x->ep_return && !aligned(x->ep_return, 8)) || (x->last_ep_return && !aligned(x->last_ep_return, 8)) ||
        (x->last_ep_length && !aligned(x->last_ep_length, 4)) || (x->out_obs32 && !aligned(x->out_obs32, 4)) ||
        (x->out_reward32 && !aligned(x->out_reward32, 4)) || (x->out_done_bits && !aligned(x->out_done_bits, 8)))
      return FISHING_E_ALIGN;
  }
  const int asz = act_size(action_dtype);
  if (!asz) return FISHING_E_ARG;
  if (!act_matches_kind(p->kind, action_dtype)) return FISHING_E_KIND;
  if (int e = check_soa(p, soa)) return e;
  if (!aligned(action, asz) || !aligned(out_obs, 8) || !aligned(out_reward, 8) || (noise && !aligned(noise, 8)) ||
      (reset_noise && !aligned(reset_noise, 8)) || (out_terminal && !aligned(out_terminal, 8)) ||
      (out_stock && !aligned(out_stock, 8)) || (out_harvest && !aligned(out_harvest, 8)))
    return FISHING_E_ALIGN;
  if (n == 0) return 0;
  StepArgs a;
  a.n = n; a.env_id0 = env_id0; a.step_idx = step_idx; a.seed = p->seed;
  a.r = p->r; a.K = p->K; a.sigma = p->sigma; a.C = p->C;
  a.s_reset = reset_state_value(p);
  a.K_mean = p->K; a.r_mean = p->r; a.sigma_p = p->sigma_p;
  a.n_actions = (double)p->n_actions;
  a.rcp_n_actions = 1.0 / (double)p->n_actions;
  a.kd = make_kdiv(p->K, p->r);
  a.Tmax = p->Tmax; a.flags = p->flags & (FISHING_AUTO_RESET | FISHING_FREEZE_DONE);
  a.action = action; a.noise = noise; a.reset_noise = reset_noise;
  a.state_in = soa.state; a.years = soa.years; a.K_env = soa.K_env; a.r_env = soa.r_env;
  a.out_obs = out_obs; a.out_reward = out_reward; a.out_done = out_done; a.out_terminal = out_terminal;
  a.out_stock = out_stock; a.out_harvest = out_harvest;
  a.out_obs2 = out_obs2;
  a.ep_return = x ? x->ep_return : nullptr;
  a.last_ep_return = x ? x->last_ep_return : nullptr;
  a.last_ep_length = x ? x->last_ep_length : nullptr;
  a.out_obs32 = x ? x->out_obs32 : nullptr;
  a.out_reward32 = x ? x->out_reward32 : nullptr;
  a.out_done_bits = x ? reinterpret_cast<unsigned long long*>(x->out_done_bits) : nullptr;
  a.gp = growth_of(p);
  a.keys = philox_round_keys(p->seed);
  a.model = soa.model;
  double any_sigma = 0.0;
  if (int e = fill_table(p, &a.tab, &any_sigma)) return e;
  if (p->kind == FISHING_V11) a.sigma = any_sigma;              // only steers "is any noise needed"
  if (p->kind == FISHING_V10) {                                 // r <- r + alpha BEFORE the draw (growth_models.py:148)
    volatile double r_next = p->r + p->alpha;
    a.gp.r = r_next;
    if (!kDeriveOnDevice) growth_derive(a.gp, FISHING_FN_BEVERTON_HOLT);   // log(r + 1), K / r for the drifted r
  }
  // K1 and the fused K2 (which drifts r in-kernel) must agree bit for bit: growth_derive() gives the same bits on the host and
  // on the device (csrc/dynamics.cuh), so the host derives once per launch; with the library log they both derive on the device
  a.derive_on_device = kDeriveOnDevice && p->kind == FISHING_V10;
  // 16-byte vector path: every array that is accessed two-wide must be aligned to its pair size
  bool vec = (n % 2 == 0) && (env_id0 % 2 == 0) && aligned(action, 2 * asz) && aligned(soa.state, 16) &&
             aligned(soa.years, 8) && aligned(out_obs, 16) && aligned(out_reward, 16) && aligned(out_done, 2) &&
             (!noise || aligned(noise, 16)) && (!out_stock || aligned(out_stock, 16)) &&
             (!out_harvest || aligned(out_harvest, 16)) && (!out_obs2 || aligned(out_obs2, 16)) &&
             (!a.ep_return || aligned(a.ep_return, 16)) && (!a.out_obs32 || aligned(a.out_obs32, 8)) &&
             (!a.out_reward32 || aligned(a.out_reward32, 8));
  if (p->kind == FISHING_V4) vec = vec && aligned(soa.K_env, 16) && aligned(soa.r_env, 16);
  if (p->kind == FISHING_V11) vec = vec && aligned(soa.model, 8);
  if (vec && (n >> 1) > 0x7fffffffLL) vec = false;
  if (a.out_done_bits && !vec) return FISHING_E_ALIGN;   // the bit packing is a warp ballot of the pair-per-thread kernel
  if (vec && (p->flags & FISHING_PDL) && pdl_enabled()) a.flags |= STEP_FLAG_PDL;
  if (vec && k1_prefetch(n * (p->kind == FISHING_V4 ? 28 : 12))) a.flags |= STEP_FLAG_PREFETCH;
  return (int)launch_step(a, p->kind, action_dtype, vec, static_cast<cudaStream_t>(stream));
}

int fishing_step_many(const fishing_params* p, int64_t n, int64_t env_id0, uint64_t step0, int64_t T,
                      const void* actions, int32_t action_dtype, const double* noise, fishing_soa soa, double* out_obs,
                      double* out_reward, uint8_t* out_done, double* out_terminal, void* stream) {
  if (!p || !actions || !out_obs || !out_reward || !out_done || T < 0 || n < 0) return FISHING_E_ARG;
  const int asz = act_size(action_dtype);
  if (!asz) return FISHING_E_ARG;
  // consecutive steps may overlap (PDL) when no step's inputs alias any step's outputs
  const size_t tn = (size_t)T * (size_t)n;
  fishing_params q = *p;
  const void* ins[2] = {actions, noise};
  const size_t in_bytes[2] = {tn * asz, tn * 8};
  const void* outs[5] = {out_obs, out_reward, out_done, out_terminal, soa.state};
  const size_t out_bytes[5] = {tn * 8, tn * 8, tn, tn * 8, (size_t)n * 8};
  bool alias = false;
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 5; ++k) alias = alias || overlaps(ins[i], in_bytes[i], outs[k], out_bytes[k]);
  fishing_soa cur = soa;
  for (int64_t t = 0; t < T; ++t) {
    const int64_t off = t * n;
    // the first launch of a call is fully serialised behind whatever precedes it on the stream (its inputs might have
    // been produced there); only launches that follow one of OUR launches of this call overlap with it
    if (alias || t == 0) q.flags &= ~(uint32_t)FISHING_PDL;
    else q.flags |= FISHING_PDL;
    int rc = fishing_step(&q, n, env_id0, step0 + (uint64_t)t, static_cast<const char*>(actions) + off * asz, action_dtype,
                          noise ? noise + off : nullptr, nullptr, cur, out_obs + off, out_reward + off, out_done + off,
                          out_terminal ? out_terminal + off : nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    cur.state = out_obs + off;
    if (q.kind == FISHING_V10) {
      volatile double r_next = q.r + q.alpha;
      q.r = r_next;
    }
  }
  if (T > 0 && n > 0 && cur.state != soa.state)
    return (int)cudaMemcpyAsync(soa.state, cur.state, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice,
                                static_cast<cudaStream_t>(stream));
  return 0;
}

int fishing_noise_fill(uint64_t seed, int32_t noise_stream, int64_t n, int64_t env_id0, uint64_t step0, int64_t T,
                       double* out, void* stream) {
  if (n < 0 || T < 0 || env_id0 < 0 || !out || noise_stream < 0 || noise_stream > 2) return FISHING_E_ARG;
  if (!aligned(out, 8)) return FISHING_E_ALIGN;
  if (n == 0 || T == 0) return 0;
  return (int)launch_noise_fill(seed, noise_stream, n, env_id0, step0, T, out, static_cast<cudaStream_t>(stream));
}

int fishing_rollout(const fishing_params* p, int64_t n, int64_t env_id0, const fishing_rollout_args* r, fishing_soa soa,
                    void* stream) {
  if (!p || !r || n < 0 || env_id0 < 0 || !kind_ok(p->kind) || r->T < 0) return FISHING_E_ARG;
  if (r->policy < FISHING_POLICY_ACTIONS || r->policy > FISHING_POLICY_MSY) return FISHING_E_ARG;
  if (int e = check_soa(p, soa)) return e;
  int asz = 8;
  if (r->policy == FISHING_POLICY_ACTIONS) {
    if (!r->actions) return FISHING_E_ARG;
    asz = act_size(r->action_dtype);
    if (!asz) return FISHING_E_ARG;
    if (!act_matches_kind(p->kind, r->action_dtype)) return FISHING_E_KIND;
  }
  const int64_t m = r->noise_id_div > 1 ? r->noise_id_div : 1;
  // the fused kernel owns environments in pairs
  if (n % 2 != 0 || env_id0 % 2 != 0) return FISHING_E_ARG;
  if (m > 1 && (n % (2 * m) != 0 || env_id0 % (2 * m) != 0)) return FISHING_E_ARG;
  if (m == 1) {
    const bool ok = aligned(soa.state, 16) && aligned(soa.years, 8) &&
                    (p->kind != FISHING_V4 || (aligned(soa.K_env, 16) && aligned(soa.r_env, 16))) &&
                    (!r->actions || aligned(r->actions, 2 * asz)) && (!r->noise || aligned(r->noise, 16)) &&
                    (!r->S_env || aligned(r->S_env, 16)) && (!r->out_return || aligned(r->out_return, 16)) &&
                    (!r->out_done_step || aligned(r->out_done_step, 8)) && (!r->trace_obs || aligned(r->trace_obs, 16)) &&
                    (!r->trace_reward || aligned(r->trace_reward, 16)) && (!r->trace_done || aligned(r->trace_done, 2)) &&
                    (!r->trace_action || aligned(r->trace_action, 16));
    if (!ok) return FISHING_E_ALIGN;
  }
  if (n == 0) return 0;
  RolloutArgs a;
  a.n = n; a.env_id0 = env_id0; a.T = r->T; a.step0 = r->step0; a.seed = p->seed;
  a.keys = philox_round_keys(p->seed);
  a.r = p->r; a.K = p->K; a.sigma = p->sigma; a.C = p->C;
  a.s_reset = reset_state_value(p);
  {
    volatile double xr = a.s_reset + 1.0;        // volatile: one binary64 rounding each, whatever the host compiler would like to fold
    volatile double back = xr - 1.0;
    a.x_reset_fold = xr;
    const double b = back, s0 = a.s_reset;
    a.fold_reset = (std::memcmp(&b, &s0, sizeof(double)) == 0 && xr >= 0.0) ? 1 : 0;    // same BITS (a signed zero counts)
  }
  a.K_mean = p->K; a.r_mean = p->r; a.sigma_p = p->sigma_p;
  a.n_actions = (double)p->n_actions;
  a.rcp_n_actions = 1.0 / (double)p->n_actions;
  a.kd = make_kdiv(p->K, p->r);
  a.Tmax = p->Tmax; a.flags = p->flags;
  a.discrete = p->kind == FISHING_V0;
  a.action_dtype = r->action_dtype;
  a.S = r->S; a.msy_q = r->msy_q; a.S_env = r->S_env;
  a.actions = r->actions; a.noise = r->noise; a.noise_id_div = m;
  a.state = soa.state; a.years = soa.years; a.K_env = soa.K_env; a.r_env = soa.r_env;
  a.out_return = r->out_return; a.out_done_step = r->out_done_step;
  a.trace_obs = r->trace_obs; a.trace_reward = r->trace_reward; a.trace_done = r->trace_done;
  a.trace_action = r->trace_action;
  a.gp = growth_of(p);
  a.alpha = p->alpha;
  a.nonstationary = p->kind == FISHING_V10;
  a.model = soa.model;
  double any_sigma = 0.0;
  if (int e = fill_table(p, &a.tab, &any_sigma)) return e;
  if (p->kind == FISHING_V11) a.sigma = any_sigma;
  if (m == 1 && p->kind == FISHING_V11 && !aligned(soa.model, 8)) return FISHING_E_ALIGN;
  return (int)launch_rollout(a, p->kind, r->policy, static_cast<cudaStream_t>(stream));
}

int64_t fishing_reduce_scratch_bytes(int64_t n, int64_t groups) {
  if (n <= 0 || groups <= 0 || n % groups != 0) return 0;
  return reduce_scratch_doubles(n, groups) * (int64_t)sizeof(double);
}

int fishing_reduce_stats(const double* ret, const int32_t* done_step, int64_t n, int64_t groups, void* scratch,
                         double* out, void* stream) {
  if (!ret || !scratch || !out || n <= 0 || groups <= 0 || n % groups != 0) return FISHING_E_ARG;
  if (!aligned(ret, 8) || !aligned(scratch, 8) || !aligned(out, 8)) return FISHING_E_ALIGN;
  return (int)launch_reduce(ret, done_step, n, groups, static_cast<double*>(scratch), out,
                            static_cast<cudaStream_t>(stream));
}

int fishing_population_draw(const fishing_params* p, int64_t n, const void* x, int32_t is_f32, double z, void* out,
                            void* stream) {
  if (!p || !kind_ok(p->kind) || n < 0 || !x || !out) return FISHING_E_ARG;
  // the log-normal family (fishing-v5..v11) has its own growth functions: fishing_population_draw_ln.  The kernel behind
  // this entry point only knows the logistic and tipping-point models and would silently return a logistic draw.
  if (is_lognormal(p->kind)) return FISHING_E_KIND;
  if (!aligned(x, is_f32 ? 4 : 8) || !aligned(out, is_f32 ? 4 : 8)) return FISHING_E_ALIGN;
  if (n == 0) return 0;
  return (int)launch_growth(p->kind, n, x, is_f32 != 0, p->r, p->K, p->C, p->sigma, z, out,
                            static_cast<cudaStream_t>(stream));
}

int fishing_population_draw_ln(const fishing_params* p, int32_t model_index, int64_t n, const double* x, const double* z,
                               double* out, void* stream) {
  if (!p || !is_lognormal(p->kind) || n < 0 || !x || !out) return FISHING_E_ARG;
  if (!aligned(x, 8) || !aligned(out, 8) || (z && !aligned(z, 8))) return FISHING_E_ALIGN;
  if (n == 0) return 0;
  GrowthP g = growth_of(p);
  int fn;
  switch (p->kind) {
    case FISHING_V5: fn = FISHING_FN_ALLEN; break;
    case FISHING_V6: case FISHING_V10: fn = FISHING_FN_BEVERTON_HOLT; break;
    case FISHING_V7: fn = FISHING_FN_MAY; break;
    case FISHING_V8: fn = FISHING_FN_MYERS; break;
    case FISHING_V9: fn = FISHING_FN_RICKER; break;
    default: {
      ModelTable t;
      double any;
      if (int e = fill_table(p, &t, &any)) return e;
      if (model_index < 0 || model_index >= t.n_models) return FISHING_E_ARG;
      fn = t.fn[model_index];
      g = t.p[model_index];
    }
  }
  if (p->kind == FISHING_V10) {     // r <- r + alpha before the draw; the derived fields follow (as in K1 / K2)
    volatile double r_next = p->r + p->alpha;
    g.r = r_next;
    if (!kDeriveOnDevice) growth_derive(g, fn);
  }
  return (int)launch_growth_ln(fn, n, x, g, kDeriveOnDevice && p->kind == FISHING_V10, z, out, static_cast<cudaStream_t>(stream));
}

int fishing_bmsy(int32_t kind, double r, double K, double C, int32_t K_is_f64, double* S_out, double* msy_out) {
  if (!kind_ok(kind) || !S_out) return FISHING_E_ARG;
  if (is_lognormal(kind)) return FISHING_E_KIND;   // no logistic BMSY for fishing-v5..v11 (models/policies.py evaluates theirs)
  double* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, 2 * sizeof(double));
  if (e != cudaSuccess) return (int)e;
  e = launch_bmsy(kind, r, K, C, K_is_f64 != 0 || kind == FISHING_V4, dev, nullptr);
  double host[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(host, dev, sizeof(host), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) return (int)e;
  *S_out = host[0];
  if (msy_out) *msy_out = host[1];
  return 0;
}

// ================================================================================================ host-buffer front end
struct fishing_vec {
  fishing_params p;
  int64_t n, env_id0;
  int device;
  uint64_t step_idx, epoch;
  bool was_reset;
  fishing_soa soa;
  // device staging for one step
  void* d_action;      // n * 8
  double* d_noise;     // n
  double* d_obs;       // == soa.state (obs are written in place)
  double* d_reward;
  uint8_t* d_done;
  double* d_terminal;
  // staging of fishing_vec_step_many*: [T, chunk] blocks per pipeline chunk, grown on demand
  char* m_action;      // T * n * 8
  double* m_obs;       // T * n
  double* m_reward;
  uint8_t* m_done;
  float* m_obs32;      // only for the float32 wire formats
  float* m_reward32;
  uint64_t* m_bits;    // only for bit-packed dones: T * words(n)
  int64_t m_T, m_T32, m_Tbits;   // rows the three groups are currently sized for
  cudaStream_t streams[3];
  cudaEvent_t ev[3];
  int64_t chunk;       // environments per pipeline chunk (even)
};

static void vec_free(fishing_vec* v) {
  if (!v) return;
  cudaFree(v->soa.state); cudaFree(v->soa.years); cudaFree(v->soa.K_env); cudaFree(v->soa.r_env); cudaFree(v->soa.model);
  cudaFree(v->d_action); cudaFree(v->d_noise); cudaFree(v->d_reward); cudaFree(v->d_done); cudaFree(v->d_terminal);
  cudaFree(v->m_action); cudaFree(v->m_obs); cudaFree(v->m_reward); cudaFree(v->m_done);
  cudaFree(v->m_obs32); cudaFree(v->m_reward32); cudaFree(v->m_bits);
  for (int i = 0; i < 3; ++i) {
    if (v->streams[i]) cudaStreamDestroy(v->streams[i]);
    if (v->ev[i]) cudaEventDestroy(v->ev[i]);
  }
  delete v;
}

int fishing_vec_create(const fishing_params* p, int64_t n, int64_t env_id0, int32_t device, fishing_vec** out) {
  if (!p || !out || n <= 0 || env_id0 < 0 || !kind_ok(p->kind)) return FISHING_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return FISHING_E_NODEVICE;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  fishing_vec* v = new (std::nothrow) fishing_vec();
  if (!v) return FISHING_E_ARG;
  std::memset(v, 0, sizeof(*v));
  v->p = *p; v->n = n; v->env_id0 = env_id0; v->device = device;
  const size_t nd = (size_t)n * sizeof(double);
  bool ok = cudaMalloc(&v->soa.state, nd) == cudaSuccess && cudaMalloc(&v->soa.years, (size_t)n * 4) == cudaSuccess &&
            cudaMalloc(&v->d_action, nd) == cudaSuccess && cudaMalloc(&v->d_noise, nd) == cudaSuccess &&
            cudaMalloc(&v->d_reward, nd) == cudaSuccess && cudaMalloc(&v->d_done, (size_t)n) == cudaSuccess &&
            cudaMalloc(&v->d_terminal, nd) == cudaSuccess;
  if (ok && p->kind == FISHING_V4)
    ok = cudaMalloc(&v->soa.K_env, nd) == cudaSuccess && cudaMalloc(&v->soa.r_env, nd) == cudaSuccess;
  if (ok && p->kind == FISHING_V11) ok = cudaMalloc(&v->soa.model, (size_t)n * 4) == cudaSuccess;
  for (int i = 0; ok && i < 3; ++i)
    ok = cudaStreamCreateWithFlags(&v->streams[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&v->ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    e = cudaGetLastError();
    vec_free(v);
    return e != cudaSuccess ? (int)e : (int)cudaErrorMemoryAllocation;
  }
  v->d_obs = v->soa.state;
  // NaN until an episode actually ends; on the stream the first step will use (the handle's streams are non-blocking, so
  // a legacy-stream memset would not be ordered against them)
  if (cudaMemsetAsync(v->d_terminal, 0xFF, nd, v->streams[0]) != cudaSuccess || cudaStreamSynchronize(v->streams[0]) != cudaSuccess) {
    e = cudaGetLastError();
    vec_free(v);
    return e != cudaSuccess ? (int)e : (int)cudaErrorUnknown;
  }
  // a few chunks in flight hide the first H2D / last D2H; chunks stay even and 16-byte aligned
  int64_t nchunks = 2;
  if (const char* e = std::getenv("FISHING_VEC_CHUNKS")) {
    const long v = std::atol(e);
    if (v >= 1 && v <= 64) nchunks = v;
  }
  int64_t chunk = (n + nchunks - 1) / nchunks;
  chunk = (chunk + 1023) / 1024 * 1024;
  if (chunk > n) chunk = n;
  v->chunk = chunk;
  *out = v;
  return 0;
}

int fishing_vec_destroy(fishing_vec* v) {
  if (!v) return FISHING_E_ARG;
  cudaSetDevice(v->device);
  for (int i = 0; i < 3; ++i) cudaStreamSynchronize(v->streams[i]);
  vec_free(v);
  return 0;
}

int fishing_vec_soa(fishing_vec* v, fishing_soa* out) {
  if (!v || !out) return FISHING_E_ARG;
  *out = v->soa;
  return 0;
}

int fishing_vec_reset(fishing_vec* v, double* obs_out) {
  if (!v) return FISHING_E_ARG;
  cudaError_t e = cudaSetDevice(v->device);
  if (e != cudaSuccess) return (int)e;
  cudaStream_t st = v->streams[0];
  int rc = fishing_reset(&v->p, v->n, v->env_id0, v->epoch, nullptr, nullptr, v->soa, nullptr, st);
  if (rc) return rc;
  v->epoch += 1;
  v->was_reset = true;
  if (obs_out) {
    e = cudaMemcpyAsync(obs_out, v->soa.state, (size_t)v->n * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
  }
  return (int)cudaStreamSynchronize(st);
}

// bookkeeping after one successful lock-step of the whole batch
static void vec_advance(fishing_vec* v) {
  v->step_idx += 1;
  if (v->p.kind == FISHING_V10) {   // r <- r + alpha, sequential adds like the reference (growth_models.py:147-150)
    volatile double r_next = v->p.r + v->p.alpha;
    v->p.r = r_next;
  }
}

int fishing_vec_step(fishing_vec* v, const void* action, int32_t action_dtype, const double* noise, double* obs_out,
                     double* reward_out, uint8_t* done_out, double* terminal_out) {
  if (!v || !action || !obs_out || !reward_out || !done_out) return FISHING_E_ARG;
  if (!v->was_reset) return FISHING_E_STATE;
  const int asz = act_size(action_dtype);
  if (!asz) return FISHING_E_ARG;
  if (!act_matches_kind(v->p.kind, action_dtype)) return FISHING_E_KIND;
  cudaError_t e = cudaSetDevice(v->device);
  if (e != cudaSuccess) return (int)e;
  int rc = 0;
  // Small batches with page-locked caller buffers: one launch that reads the actions and writes observations / rewards /
  // dones straight over PCIe, then one synchronisation.  Up to ~64k environments a step is latency, not bandwidth: the
  // staged route costs five asynchronous operations, this one a launch (measured per FishingVecEnv.step call, numpy in
  // and out: 36 -> 22 us at n = 1024, 95 -> 64 us at n = 65536; at 262144 the staged pipeline is faster again).
  if (!noise && v->n <= zero_copy_max_envs()) {
    void* d_act = device_alias(action);
    void* d_obs = device_alias(obs_out);
    void* d_rew = device_alias(reward_out);
    void* d_done = device_alias(done_out);
    void* d_term = terminal_out ? device_alias(terminal_out) : static_cast<void*>(v->d_terminal);
    if (d_act && d_obs && d_rew && d_done && d_term) {
      fishing_params q = v->p;
      q.flags &= ~(uint32_t)FISHING_PDL;
      rc = step_impl(&q, v->n, v->env_id0, v->step_idx, d_act, action_dtype, nullptr, nullptr, v->soa, v->soa.state,
                     static_cast<double*>(d_rew), static_cast<uint8_t*>(d_done), static_cast<double*>(d_term), nullptr,
                     nullptr, static_cast<double*>(d_obs), v->streams[0]);
      e = cudaStreamSynchronize(v->streams[0]);
      if (e != cudaSuccess && rc == 0) rc = (int)e;
      if (rc == 0) vec_advance(v);     // a failed step leaves the counters where they were
      return rc;
    }
  }
  int k = 0;
  for (int64_t off = 0; off < v->n && rc == 0; off += v->chunk, ++k) {
    const int64_t len = (v->n - off < v->chunk) ? v->n - off : v->chunk;
    cudaStream_t st = v->streams[k % 3];
    char* d_act = static_cast<char*>(v->d_action) + off * asz;
    e = cudaMemcpyAsync(d_act, static_cast<const char*>(action) + off * asz, (size_t)len * asz, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && noise)
      e = cudaMemcpyAsync(v->d_noise + off, noise + off, (size_t)len * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = (int)e; break; }
    fishing_soa s = v->soa;
    s.state += off; s.years += off;
    if (s.K_env) { s.K_env += off; s.r_env += off; }
    if (s.model) s.model += off;
    rc = fishing_step(&v->p, len, v->env_id0 + off, v->step_idx, d_act, action_dtype, noise ? v->d_noise + off : nullptr,
                      nullptr, s, v->d_obs + off, v->d_reward + off, v->d_done + off,
                      v->d_terminal + off, nullptr, nullptr, st);
    if (rc) break;
    e = cudaMemcpyAsync(obs_out + off, v->d_obs + off, (size_t)len * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(reward_out + off, v->d_reward + off, (size_t)len * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(done_out + off, v->d_done + off, (size_t)len, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && terminal_out)
      e = cudaMemcpyAsync(terminal_out + off, v->d_terminal + off, (size_t)len * 8, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = (int)e;
  }
  for (int i = 0; i < 3; ++i) {
    e = cudaStreamSynchronize(v->streams[i]);
    if (e != cudaSuccess && rc == 0) rc = (int)e;
  }
  if (rc == 0) vec_advance(v);
  return rc;
}

// Grow the [T, n] staging blocks.  Each group is allocated as a whole or not at all (a half-grown group is never left behind).
static int vec_ensure_staging(fishing_vec* v, int64_t T, uint32_t wire) {
  const size_t n = (size_t)v->n;
  if (T > v->m_T) {
    cudaFree(v->m_action); cudaFree(v->m_obs); cudaFree(v->m_reward); cudaFree(v->m_done);
    v->m_action = nullptr; v->m_obs = nullptr; v->m_reward = nullptr; v->m_done = nullptr; v->m_T = 0;
    if (cudaMalloc(&v->m_action, T * n * 8) != cudaSuccess || cudaMalloc(&v->m_obs, T * n * 8) != cudaSuccess ||
        cudaMalloc(&v->m_reward, T * n * 8) != cudaSuccess || cudaMalloc(&v->m_done, T * n) != cudaSuccess) {
      cudaFree(v->m_action); cudaFree(v->m_obs); cudaFree(v->m_reward); cudaFree(v->m_done);
      v->m_action = nullptr; v->m_obs = nullptr; v->m_reward = nullptr; v->m_done = nullptr;
      cudaGetLastError();
      return (int)cudaErrorMemoryAllocation;
    }
    v->m_T = T;
  }
  if ((wire & (FISHING_WIRE_OBS_F32 | FISHING_WIRE_REWARD_F32)) && T > v->m_T32) {
    cudaFree(v->m_obs32); cudaFree(v->m_reward32);
    v->m_obs32 = nullptr; v->m_reward32 = nullptr; v->m_T32 = 0;
    if (cudaMalloc(&v->m_obs32, T * n * 4) != cudaSuccess || cudaMalloc(&v->m_reward32, T * n * 4) != cudaSuccess) {
      cudaFree(v->m_obs32); cudaFree(v->m_reward32);
      v->m_obs32 = nullptr; v->m_reward32 = nullptr;
      cudaGetLastError();
      return (int)cudaErrorMemoryAllocation;
    }
    v->m_T32 = T;
  }
  if ((wire & FISHING_WIRE_DONE_BITS) && T > v->m_Tbits) {
    cudaFree(v->m_bits);
    v->m_bits = nullptr; v->m_Tbits = 0;
    // every pipeline chunk rounds its row up to whole 64-bit words: at most 64 extra words per row
    if (cudaMalloc(&v->m_bits, T * ((n + 63) / 64 + 64) * 8) != cudaSuccess) {
      cudaGetLastError();
      return (int)cudaErrorMemoryAllocation;
    }
    v->m_Tbits = T;
  }
  return 0;
}

// T lock-steps with host arrays.  The environment range is cut into chunks that rotate over three streams; a chunk's whole
// [T, len] action block comes over in ONE strided copy, its T kernels run back to back (programmatic dependent launch between
// them, row t of the observation block is the state of step t + 1), and every output array goes back in ONE strided copy:
// 4 transfers per chunk whatever T is, large enough to run at the link rate, with the chunks on the other two streams
// keeping both copy engines and the SMs busy meanwhile.
static int vec_step_many_core(fishing_vec* v, int64_t T, const void* actions, int32_t action_dtype, void* obs_out,
                              void* reward_out, void* done_out, uint32_t wire) {
  if (!v || !actions || !obs_out || !reward_out || !done_out || T < 0) return FISHING_E_ARG;
  if (wire & ~(uint32_t)(FISHING_WIRE_OBS_F32 | FISHING_WIRE_REWARD_F32 | FISHING_WIRE_DONE_BITS)) return FISHING_E_ARG;
  if (!v->was_reset) return FISHING_E_STATE;
  const int asz = act_size(action_dtype);
  if (!asz) return FISHING_E_ARG;
  if (!act_matches_kind(v->p.kind, action_dtype)) return FISHING_E_KIND;
  const int64_t n = v->n;
  if ((wire & FISHING_WIRE_DONE_BITS) && (n % 2 != 0 || v->env_id0 % 2 != 0)) return FISHING_E_ALIGN;
  if (T == 0) return 0;
  cudaError_t e = cudaSetDevice(v->device);
  if (e != cudaSuccess) return (int)e;
  if (int rc = vec_ensure_staging(v, T, wire)) return rc;
  const bool obs32 = wire & FISHING_WIRE_OBS_F32, rew32 = wire & FISHING_WIRE_REWARD_F32, bits = wire & FISHING_WIRE_DONE_BITS;
  const size_t host_bits_row = (size_t)(n + 7) / 8;
  // chunks rotate over three streams; the first action block (nothing can overlap it) is 1/chunks of the call's H2D traffic.
  // Measured on B200 (tools/e2e_chunks.py, 2^20 envs, float64 wire): T = 8 per call 2.67 / 2.78 / 2.74 / 2.54e9 env-steps/s
  // for 3 / 6 / 12 / 24 chunks, T = 32: 2.75 / 2.90 / 2.97 / 2.94e9.
  static const int64_t forced = [] {
    const char* e = std::getenv("FISHING_VEC_MANY_CHUNKS");
    const long c = e ? std::atol(e) : 0;
    return (int64_t)((c >= 1 && c <= 64) ? c : 0);
  }();
  const int64_t nchunks = forced ? forced : (T >= 16 ? 12 : 6);
  int64_t chunk = ((n + nchunks - 1) / nchunks + 1023) / 1024 * 1024;
  if (chunk > n) chunk = n;
  int rc = 0;
  double r_after = v->p.r;
  int k = 0;
  int64_t bit_words_before = 0;   // words of bit rows of the chunks before this one
  for (int64_t off = 0; off < n && rc == 0; off += chunk, ++k) {
    const int64_t len = (n - off < chunk) ? n - off : chunk;
    const int64_t words = (len + 63) / 64;
    cudaStream_t st = v->streams[k % 3];
    char* d_act = v->m_action + (size_t)T * off * asz;
    double* d_obs = v->m_obs + (size_t)T * off;
    double* d_rew = v->m_reward + (size_t)T * off;
    uint8_t* d_done = v->m_done + (size_t)T * off;
    float* d_obs32 = obs32 ? v->m_obs32 + (size_t)T * off : nullptr;
    float* d_rew32 = rew32 ? v->m_reward32 + (size_t)T * off : nullptr;
    uint64_t* d_bits = bits ? v->m_bits + (size_t)T * bit_words_before : nullptr;
    e = cudaMemcpy2DAsync(d_act, (size_t)len * asz, static_cast<const char*>(actions) + (size_t)off * asz, (size_t)n * asz,
                          (size_t)len * asz, (size_t)T, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = (int)e; break; }
    fishing_soa cur = v->soa;
    cur.state += off; cur.years += off;
    if (cur.K_env) { cur.K_env += off; cur.r_env += off; }
    if (cur.model) cur.model += off;
    double* const state_home = cur.state;
    fishing_params q = v->p;
    for (int64_t t = 0; t < T && rc == 0; ++t) {
      q.flags = (v->p.flags & ~(uint32_t)FISHING_PDL) | (t > 0 ? (uint32_t)FISHING_PDL : 0u);
      fishing_step_extra x = {};
      x.out_obs32 = d_obs32 ? d_obs32 + t * len : nullptr;
      x.out_reward32 = d_rew32 ? d_rew32 + t * len : nullptr;
      x.out_done_bits = d_bits ? d_bits + t * words : nullptr;
      rc = step_impl(&q, len, v->env_id0 + off, v->step_idx + (uint64_t)t, d_act + (size_t)t * len * asz, action_dtype, nullptr,
                     nullptr, cur, d_obs + t * len, d_rew + t * len, d_done + t * len, v->d_terminal + off, nullptr, nullptr,
                     nullptr, st, wire ? &x : nullptr);
      cur.state = d_obs + t * len;
      if (q.kind == FISHING_V10) {
        volatile double r_next = q.r + q.alpha;
        q.r = r_next;
      }
    }
    if (rc) break;
    r_after = q.r;
    e = cudaMemcpyAsync(state_home, cur.state, (size_t)len * 8, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)
      e = obs32 ? cudaMemcpy2DAsync(static_cast<char*>(obs_out) + (size_t)off * 4, (size_t)n * 4, d_obs32, (size_t)len * 4,
                                    (size_t)len * 4, (size_t)T, cudaMemcpyDeviceToHost, st)
                : cudaMemcpy2DAsync(static_cast<char*>(obs_out) + (size_t)off * 8, (size_t)n * 8, d_obs, (size_t)len * 8,
                                    (size_t)len * 8, (size_t)T, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = rew32 ? cudaMemcpy2DAsync(static_cast<char*>(reward_out) + (size_t)off * 4, (size_t)n * 4, d_rew32, (size_t)len * 4,
                                    (size_t)len * 4, (size_t)T, cudaMemcpyDeviceToHost, st)
                : cudaMemcpy2DAsync(static_cast<char*>(reward_out) + (size_t)off * 8, (size_t)n * 8, d_rew, (size_t)len * 8,
                                    (size_t)len * 8, (size_t)T, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = bits ? cudaMemcpy2DAsync(static_cast<char*>(done_out) + (size_t)off / 8, host_bits_row, d_bits, (size_t)words * 8,
                                   (size_t)(len + 7) / 8, (size_t)T, cudaMemcpyDeviceToHost, st)
               : cudaMemcpy2DAsync(static_cast<char*>(done_out) + (size_t)off, (size_t)n, d_done, (size_t)len, (size_t)len, (size_t)T,
                                   cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = (int)e;
    bit_words_before += words;
  }
  for (int i = 0; i < 3; ++i) {
    e = cudaStreamSynchronize(v->streams[i]);
    if (e != cudaSuccess && rc == 0) rc = (int)e;
  }
  if (rc == 0) {   // a failed call leaves the step counter (and fishing-v10's drifting r) where they were
    v->step_idx += (uint64_t)T;
    v->p.r = r_after;
  }
  return rc;
}

int fishing_vec_step_many(fishing_vec* v, int64_t T, const void* actions, int32_t action_dtype, double* obs_out,
                          double* reward_out, uint8_t* done_out) {
  if (!v || !actions || !obs_out || !reward_out || !done_out || T < 0) return FISHING_E_ARG;
  if (!v->was_reset) return FISHING_E_STATE;
  // Zero-copy mode (opt-in, FISHING_VEC_ZEROCOPY=1): when every host array is page-locked the kernel reads the actions and
  // writes observations / rewards / dones straight over PCIe, one launch per lock-step -- no staging, no copy engines.
  if (zero_copy_enabled()) {
    const int asz = act_size(action_dtype);
    if (!asz) return FISHING_E_ARG;
    if (!act_matches_kind(v->p.kind, action_dtype)) return FISHING_E_KIND;
    void* d_act = device_alias(actions);
    void* d_obs = device_alias(obs_out);
    void* d_rew = device_alias(reward_out);
    void* d_done = device_alias(done_out);
    if (d_act && d_obs && d_rew && d_done) {
      cudaError_t e = cudaSetDevice(v->device);
      if (e != cudaSuccess) return (int)e;
      cudaStream_t st = v->streams[0];
      fishing_params q = v->p;
      int rc = 0;
      for (int64_t t = 0; t < T && rc == 0; ++t) {
        q.flags = (v->p.flags & ~(uint32_t)FISHING_PDL) | (t > 0 ? (uint32_t)FISHING_PDL : 0u);
        rc = step_impl(&q, v->n, v->env_id0, v->step_idx + (uint64_t)t, static_cast<const char*>(d_act) + t * v->n * asz,
                       action_dtype, nullptr, nullptr, v->soa, v->soa.state, static_cast<double*>(d_rew) + t * v->n,
                       static_cast<uint8_t*>(d_done) + t * v->n, v->d_terminal, nullptr, nullptr,
                       static_cast<double*>(d_obs) + t * v->n, st);
        if (q.kind == FISHING_V10) {
          volatile double r_next = q.r + q.alpha;
          q.r = r_next;
        }
      }
      e = cudaStreamSynchronize(st);
      if (rc == 0 && e != cudaSuccess) rc = (int)e;
      if (rc == 0) {
        v->step_idx += (uint64_t)T;
        v->p.r = q.r;
      }
      return rc;
    }
  }
  return vec_step_many_core(v, T, actions, action_dtype, obs_out, reward_out, done_out, 0u);
}

int fishing_vec_step_many_f32(fishing_vec* v, int64_t T, const void* actions, int32_t action_dtype, float* obs_out,
                              float* reward_out, uint8_t* done_out) {
  return vec_step_many_core(v, T, actions, action_dtype, obs_out, reward_out, done_out,
                            FISHING_WIRE_OBS_F32 | FISHING_WIRE_REWARD_F32);
}

int fishing_vec_step_many_wire(fishing_vec* v, int64_t T, const void* actions, int32_t action_dtype, void* obs_out,
                               void* reward_out, void* done_out, uint32_t wire) {
  return vec_step_many_core(v, T, actions, action_dtype, obs_out, reward_out, done_out, wire);
}

int fishing_vec_fetch_terminal(fishing_vec* v, double* terminal_out) {
  if (!v || !terminal_out) return FISHING_E_ARG;
  cudaError_t e = cudaSetDevice(v->device);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(terminal_out, v->d_terminal, (size_t)v->n * sizeof(double), cudaMemcpyDeviceToHost, v->streams[0]);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(v->streams[0]);
}

// which: 0 state f64, 1 years i32, 2 K_env f64, 3 r_env f64, 4 model i32, 5 terminal observations f64
static int vec_array(fishing_vec* v, int32_t which, void** ptr, size_t* elem) {
  switch (which) {
    case FISHING_ARRAY_STATE: *ptr = v->soa.state; *elem = 8; break;
    case FISHING_ARRAY_YEARS: *ptr = v->soa.years; *elem = 4; break;
    case FISHING_ARRAY_K: *ptr = v->soa.K_env; *elem = 8; break;
    case FISHING_ARRAY_R: *ptr = v->soa.r_env; *elem = 8; break;
    case FISHING_ARRAY_MODEL: *ptr = v->soa.model; *elem = 4; break;
    case FISHING_ARRAY_TERMINAL: *ptr = v->d_terminal; *elem = 8; break;
    default: return FISHING_E_ARG;
  }
  return *ptr ? 0 : FISHING_E_KIND;   // the environment id has no such array (K / r: fishing-v4, model: fishing-v11)
}

int fishing_vec_fetch_array(fishing_vec* v, int32_t which, void* out) {
  if (!v || !out) return FISHING_E_ARG;
  void* src; size_t elem;
  if (int rc = vec_array(v, which, &src, &elem)) return rc;
  cudaError_t e = cudaSetDevice(v->device);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(out, src, (size_t)v->n * elem, cudaMemcpyDeviceToHost, v->streams[0]);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(v->streams[0]);
}

int fishing_vec_store_array(fishing_vec* v, int32_t which, const void* in) {
  if (!v || !in) return FISHING_E_ARG;
  if (which == FISHING_ARRAY_TERMINAL) return FISHING_E_ARG;
  void* dst; size_t elem;
  if (int rc = vec_array(v, which, &dst, &elem)) return rc;
  cudaError_t e = cudaSetDevice(v->device);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(dst, in, (size_t)v->n * elem, cudaMemcpyHostToDevice, v->streams[0]);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(v->streams[0]);
}

int fishing_vec_set_params(fishing_vec* v, const fishing_params* p) {
  if (!v || !p) return FISHING_E_ARG;
  if (p->kind != v->p.kind) return FISHING_E_KIND;     // the arrays of the handle belong to its environment id
  v->p = *p;
  return 0;
}

int fishing_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return FISHING_E_ARG;
  return (int)cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault);
}

int fishing_host_free(void* ptr) { return ptr ? (int)cudaFreeHost(ptr) : FISHING_E_ARG; }

}  // extern "C"
