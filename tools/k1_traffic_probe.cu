// What does the memory system allow for K1's access pattern?  A compute-free replica of the lock-step kernel's traffic
// (per environment: action 8 B in, state 8 B in, years 4 B in/out, obs 8 B out, reward 8 B out, done 1 B out = 41 B), same
// launch shape as K1 (one pair of environments per thread, 256-thread blocks, T launches replayed from a CUDA graph with
// programmatic dependent launch between them, obs row t = state of launch t + 1), timed with CUDA events.
// Variants drop or shrink individual streams to show which bytes cost what.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/k1_traffic_probe tools/k1_traffic_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Args {
  const double* act; const double* state_in; double* obs; double* rew; uint8_t* done; int32_t* years; int npairs; int flags;
};
enum { F_YEARS_R = 1, F_YEARS_W = 2, F_REWARD = 4, F_DONE = 8, F_PREFETCH = 16, F_ACTION = 32 };

template <int PAIRS>
__global__ void __launch_bounds__(256) traffic(const Args a) {
  asm volatile("griddepcontrol.launch_dependents;");
  const int j0 = (blockIdx.x * 256 + threadIdx.x) * PAIRS;
  double2 act[PAIRS];
#pragma unroll
  for (int p = 0; p < PAIRS; ++p) {
    act[p] = make_double2(0.0, 0.0);
    if (j0 + p < a.npairs) {
      if (a.flags & F_ACTION) act[p] = __ldcs(reinterpret_cast<const double2*>(a.act) + j0 + p);
      if (a.flags & F_PREFETCH) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const double2*>(a.state_in) + j0 + p));
        if (a.flags & F_YEARS_R) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const int2*>(a.years) + j0 + p));
      }
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int p = 0; p < PAIRS; ++p) {
    const int j = j0 + p;
    if (j >= a.npairs) continue;
    const double2 s = reinterpret_cast<const double2*>(a.state_in)[j];
    int2 y = make_int2(0, 0);
    if (a.flags & F_YEARS_R) y = reinterpret_cast<const int2*>(a.years)[j];
    const double2 o = make_double2(s.x * 0.999 + act[p].x * 1e-3, s.y * 0.999 + act[p].y * 1e-3);
    reinterpret_cast<double2*>(a.obs)[j] = o;
    if (a.flags & F_REWARD) __stcs(reinterpret_cast<double2*>(a.rew) + j, make_double2(act[p].x + s.y, act[p].y + s.x));
    if (a.flags & F_DONE) __stcs(reinterpret_cast<uchar2*>(a.done) + j, make_uchar2(o.x > 0.5, o.y > 0.5));
    if (a.flags & F_YEARS_W) reinterpret_cast<int2*>(a.years)[j] = make_int2(y.x + 1, y.y + 1);
  }
}

template <int PAIRS>
static float run(int n, int T, int flags, bool pdl, const double* act, double* obs, double* rew, uint8_t* done, int32_t* years,
                 cudaStream_t st) {
  const int npairs = n / 2;
  const int blocks = (npairs + 256 * PAIRS - 1) / (256 * PAIRS);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int t = 0; t < T; ++t) {
    Args a{act + (size_t)t * n, obs + (size_t)t * n, obs + (size_t)(t + 1) * n, rew + (size_t)t * n, done + (size_t)t * n, years, npairs, flags};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && t > 0) ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, traffic<PAIRS>, a));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0, st));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  return best * 1e3f / T;   // us per launch
}

int main(int argc, char** argv) {
  const int T = 1000;
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (int lg : {20, 21, 23}) {
    const int n = 1 << lg;
    const int Tn = lg == 23 ? 200 : T;
    double *act, *obs, *rew; uint8_t* done; int32_t* years;
    CK(cudaMalloc(&act, (size_t)Tn * n * 8));
    CK(cudaMalloc(&obs, (size_t)(Tn + 1) * n * 8));
    CK(cudaMalloc(&rew, (size_t)Tn * n * 8));
    CK(cudaMalloc(&done, (size_t)Tn * n));
    CK(cudaMalloc(&years, (size_t)n * 4));
    CK(cudaMemset(act, 0, (size_t)Tn * n * 8));
    CK(cudaMemset(obs, 0, (size_t)(Tn + 1) * n * 8));
    CK(cudaMemset(years, 0, (size_t)n * 4));
    struct V { const char* name; int flags; int bytes; };
    const int ALL = F_YEARS_R | F_YEARS_W | F_REWARD | F_DONE | F_PREFETCH | F_ACTION;
    const V vs[] = {
        {"K1 replica: all streams (41 B)", ALL, 41},
        {"no L2 prefetch (41 B)", ALL & ~F_PREFETCH, 41},
        {"years read only (37 B)", ALL & ~F_YEARS_W, 37},
        {"no years (33 B)", ALL & ~(F_YEARS_W | F_YEARS_R), 33},
        {"no years, no done (32 B)", ALL & ~(F_YEARS_W | F_YEARS_R | F_DONE), 32},
        {"state -> obs only (16 B)", F_PREFETCH, 16},
    };
    for (const V& v : vs) {
      const float us1 = run<1>(n, Tn, v.flags, true, act, obs, rew, done, years, st);
      const float us2 = run<2>(n, Tn, v.flags, true, act, obs, rew, done, years, st);
      const float us0 = run<1>(n, Tn, v.flags, false, act, obs, rew, done, years, st);
      printf("n=2^%d  %-34s  1 pair/thread %7.2f us (%5.0f GB/s)   2 pairs/thread %7.2f us   without PDL %7.2f us\n", lg, v.name, us1,
             (double)v.bytes * n / us1 * 1e-3, us2, us0);
    }
    cudaFree(act); cudaFree(obs); cudaFree(rew); cudaFree(done); cudaFree(years);
  }
  return 0;
}
