// Issue-slot probe for the B200 SM sub-partition (sm_100a): how many cycles does the scheduler need per instruction of a
// given mix, measured with clock64 INSIDE the kernel (so the figures do not depend on the SM clock, which drops under a
// saturating FP64 load) on ONE resident wave (one block per SM, w warps per sub-partition).
// Question behind it: are the fused kernels bound by the FP64 pipe or by instruction issue, i.e.
// what does an FP64 / IMAD.WIDE instruction cost next to ALU work?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu ; run: tools/issue_probe
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

constexpr int CH = 8;
constexpr int ITER = 4096;

#define DFMA_() asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b))
#define DADD_() asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(b))
#define DMUL_() asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a))
#define DSETP_() asm volatile("{.reg .pred p; setp.gt.f64 p, %1, %2; selp.u32 %0, %0, %3, p;}" : "+r"(u[c]) : "d"(x[c]), "d"(b), "r"(k))
#define LOP3U_() asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(k))
#define LOP3V_() asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(u[c]), "r"(k))
#define IADD_() asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(k))
#define IMAD_() asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(k), "r"(v[c]))
#define IMADW_() asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t;}" : "+r"(u[c]), "+r"(v[c]) : "r"(k))
#define IMADHI_() asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(k))
#define FSEL_() asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.f32 %0, %0, %2, p;}" : "+f"(f[c]) : "r"(u[c]), "f"(g))
#define SHF_() asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(u[c]) : "r"(v[c]))
#define FFMA_() asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(g), "f"(g))

struct Mode { const char* name; int n_instr; };
static const Mode MODES[] = {
    {"DFMA", 1}, {"DADD", 1}, {"DMUL", 1}, {"LOP3", 1}, {"IADD", 1}, {"IMAD", 1}, {"IMAD.WIDE", 1}, {"IMAD.HI", 1}, {"FFMA", 1},
    {"DFMA+LOP3", 2}, {"DFMA+2LOP3", 3}, {"DFMA+3 (2LOP3+IADD)", 4}, {"DFMA+IMAD", 2}, {"DFMA+IMAD.WIDE", 2}, {"DFMA+FFMA", 2},
    {"LOP3+IMAD.WIDE", 2}, {"2LOP3+IMAD.WIDE", 3}, {"LOP3+IMAD", 2}, {"LOP3+FFMA", 2}, {"DSETP+SEL", 2}, {"DFMA+DSETP+SEL", 3},
    {"2DFMA+2LOP3+IADD+IMAD.WIDE (K2-like)", 6}, {"LOP3+IADD (both ALU)", 2}, {"DFMA+DADD", 2},
    // does the ORDER of FP64 and ALU instructions inside a warp matter?  (same counts, alternating against grouped)
    {"DFMA+LOP3 ordered D D L L", 2}, {"DFMA+LOP3 ordered 4D 4L", 2}, {"DFMA+LOP3 ordered 8D 8L", 2}, {"D D L", 3}, {"D D D L", 4},
    {"D L L L", 4}, {"ISETP+FSEL", 2}, {"DFMA+ISETP+FSEL", 3}, {"2DFMA+ISETP+FSEL", 4}, {"DFMA+SHF", 2},
    // the instruction mix of the fused kernel's loop, scaled to 8 FP64 (per pair-step it is 64 FP64, 60 ALU, 16 IMAD.WIDE, 13 IMAD,
    // 13 IADD), without its dependencies: per group of 8 chains
    {"mix 8D+2W (per D)", 1}, {"mix 8D+8L+2W (per D)", 1}, {"mix 8D+8L+2W+2IMAD+2IADD (per D; K2's mix)", 1}, {"mix 8D+4L+2W (per D)", 1},
};
constexpr int NMODES = sizeof(MODES) / sizeof(MODES[0]);

template <int MODE>
__global__ void __launch_bounds__(1024) probe(double* out, uint32_t* outi, long long* cyc, double a, double b, uint32_t k, int iters) {
  double x[CH];
  uint32_t u[CH], v[CH];
  float f[CH];
  const float g = (float)b;
#pragma unroll
  for (int c = 0; c < CH; ++c) { x[c] = a + c + threadIdx.x; u[c] = k + c * 77u + threadIdx.x; v[c] = k * 3u + c + threadIdx.x * 5u; f[c] = (float)c; }   // v per thread: a uniform v would be re-materialised with a MOV per use
  __syncthreads();
  const long long c0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH && !(MODE >= 24 && MODE <= 26); ++c) {
      if (MODE == 0) DFMA_();
      if (MODE == 1) DADD_();
      if (MODE == 2) DMUL_();
      if (MODE == 3) LOP3U_();
      if (MODE == 4) IADD_();
      if (MODE == 5) IMAD_();
      if (MODE == 6) IMADW_();
      if (MODE == 7) IMADHI_();
      if (MODE == 8) FFMA_();
      if (MODE == 9) { DFMA_(); LOP3U_(); }
      if (MODE == 10) { DFMA_(); LOP3U_(); LOP3V_(); }
      if (MODE == 11) { DFMA_(); LOP3U_(); IADD_(); LOP3U_(); }
      if (MODE == 12) { DFMA_(); IMAD_(); }
      if (MODE == 13) { DFMA_(); IMADW_(); }
      if (MODE == 14) { DFMA_(); FFMA_(); }
      if (MODE == 15) { LOP3U_(); IMADW_(); }
      if (MODE == 16) { LOP3U_(); IMADW_(); LOP3V_(); }
      if (MODE == 17) { LOP3U_(); IMAD_(); }
      if (MODE == 18) { LOP3U_(); FFMA_(); }
      if (MODE == 19) DSETP_();
      if (MODE == 20) { DFMA_(); DSETP_(); }
      if (MODE == 21) { DFMA_(); LOP3U_(); DADD_(); IADD_(); IMADW_(); LOP3V_(); }
      if (MODE == 22) { LOP3U_(); IADD_(); }
      if (MODE == 23) { DFMA_(); DADD_(); }
      if (MODE == 27) { DFMA_(); DADD_(); LOP3U_(); }
      if (MODE == 28) { DFMA_(); DADD_(); DMUL_(); LOP3U_(); }
      if (MODE == 29) { DFMA_(); LOP3U_(); LOP3V_(); LOP3U_(); }
      if (MODE == 30) FSEL_();
      if (MODE == 31) { DFMA_(); FSEL_(); }
      if (MODE == 32) { DFMA_(); DADD_(); FSEL_(); }
      if (MODE == 33) { DFMA_(); SHF_(); }
      if (MODE == 34) { DFMA_(); if ((c & 3) == 0) IMADW_(); }
      if (MODE == 35) { DFMA_(); LOP3U_(); if ((c & 3) == 0) IMADW_(); }
      if (MODE == 36) { DFMA_(); LOP3U_(); if ((c & 3) == 0) IMADW_(); if ((c & 3) == 1) IMAD_(); if ((c & 3) == 2) IADD_(); }
      if (MODE == 37) { DFMA_(); if (c & 1) LOP3U_(); if ((c & 3) == 0) IMADW_(); }
    }
    // the same 8 DFMA + 8 LOP3 per iteration as "DFMA+LOP3", in a different ORDER inside the warp (normalised per D + L pair)
    if (MODE == 24) {
#pragma unroll
      for (int c = 0; c < CH; c += 2) { DFMA_(); c++; DFMA_(); c--; LOP3U_(); c++; LOP3U_(); c--; }
    }
    if (MODE == 25) {
#pragma unroll
      for (int c0 = 0; c0 < CH; c0 += 4) {
#pragma unroll
        for (int c = c0; c < c0 + 4; ++c) DFMA_();
#pragma unroll
        for (int c = c0; c < c0 + 4; ++c) LOP3U_();
      }
    }
    if (MODE == 26) {
#pragma unroll
      for (int c = 0; c < CH; ++c) DFMA_();
#pragma unroll
      for (int c = 0; c < CH; ++c) LOP3U_();
    }
  }
  const long long c1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = c1 - c0;
  double s = 0; uint32_t w = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) { s += x[c]; w ^= u[c] ^ v[c] ^ __float_as_uint(f[c]); }
  if (s == 1.2345 || w == 0x12345u) { out[threadIdx.x] = s; outi[threadIdx.x] = w; }
}

template <int MODE>
static void run(double* out, uint32_t* outi, long long* cyc, int nsm) {
  for (int wps : {1, 2, 4, 8}) {   // warps per sub-partition
    const int threads = wps * 4 * 32;
    cudaMemset(cyc, 0, sizeof(long long) * nsm * 32);
    probe<MODE><<<nsm, threads>>>(out, outi, cyc, 1.0000001, 1e-9, 0x9E3779B9u, 64);
    probe<MODE><<<nsm, threads>>>(out, outi, cyc, 1.0000001, 1e-9, 0x9E3779B9u, ITER);
    cudaDeviceSynchronize();
    std::vector<long long> h(nsm * 32);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm * 32, cudaMemcpyDeviceToHost);
    std::vector<long long> per_block;
    for (int b = 0; b < nsm; ++b) {
      long long m = 0;
      for (int w = 0; w < threads / 32; ++w) m = std::max(m, h[b * 32 + w]);
      per_block.push_back(m);
    }
    std::sort(per_block.begin(), per_block.end());
    const double med = (double)per_block[per_block.size() / 2];
    const double groups = (double)ITER * CH * wps;   // groups issued per sub-partition
    printf("%-40s warps/SMSP=%d  cycles per group per SMSP: %6.3f   per instruction: %5.3f\n", MODES[MODE].name, wps, med / groups,
           med / groups / MODES[MODE].n_instr);
  }
}

template <int M>
struct RunAll {
  static void go(double* o, uint32_t* oi, long long* c, int nsm) {
    RunAll<M - 1>::go(o, oi, c, nsm);
    run<M>(o, oi, c, nsm);
  }
};
template <>
struct RunAll<-1> {
  static void go(double*, uint32_t*, long long*, int) {}
};

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs; cycles are clock64 ticks of the SM (independent of the clock frequency)\n", p.name, p.multiProcessorCount);
  double* out; uint32_t* outi; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&outi, 4 * 1024); cudaMalloc(&cyc, sizeof(long long) * p.multiProcessorCount * 32);
  RunAll<NMODES - 1>::go(out, outi, cyc, p.multiProcessorCount);
  return cudaDeviceSynchronize() != cudaSuccess;
}
