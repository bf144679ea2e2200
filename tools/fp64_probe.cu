// Micro-benchmark of the FP64 pipe and of how it shares the issue port with the integer pipes (sm_100a).
// Answers two questions DESIGN.md's K2 roofline needs:
//   1. the measured FP64 peak (warp instructions per cycle per SM sub-partition) for DFMA / DADD / DMUL,
//   2. whether ALU (LOP3) / FMA-pipe (IMAD) instructions issue "for free" next to a saturated FP64 pipe, or whether an
//      FP64 warp instruction holds the issue port for both of its cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu ; run: tools/fp64_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 8;        // independent chains per thread
constexpr int ITER = 16384;

template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, uint32_t* outi, long long* cyc, double a, double b, uint32_t k, int iters) {
  const long long c0 = clock64();
  double x[CH];
  uint32_t u[CH], v[CH];
  float f[CH];
  const float g = (float)b;
#pragma unroll
  for (int c = 0; c < CH; ++c) { x[c] = a + c + threadIdx.x; u[c] = k + c * 77u + threadIdx.x; v[c] = k * 3u + c; f[c] = (float)c; }
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      // asm volatile keeps the exact instruction mix and order (no folding, no hoisting)
#define DFMA_() asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b))
#define DADD_() asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(b))
#define DMUL_() asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a))
#define LOP3U_() asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(k))
#define LOP3V_() asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(u[c]), "r"(k))
#define IMAD_() asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(k), "r"(v[c]))
#define IMADW_() asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t;}" : "+r"(u[c]), "+r"(v[c]) : "r"(k))
#define IMADHI_() asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(k))
#define IMADLO_() asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(u[c]) : "r"(k))
#define FSEL_() asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.f32 %0, %0, %2, p;}" : "+f"(f[c]) : "r"(k), "f"(g))
#define DSETP_() asm volatile("{.reg .pred p; setp.gt.f64 p, %1, %2; selp.u32 %0, %0, %3, p;}" : "+r"(u[c]) : "d"(x[c]), "d"(b), "r"(k))
      if (MODE == 0) DFMA_();
      if (MODE == 1) DADD_();
      if (MODE == 2) DMUL_();
      if (MODE == 3) { DFMA_(); LOP3U_(); }
      if (MODE == 4) { DFMA_(); LOP3U_(); LOP3V_(); }
      if (MODE == 5) { DFMA_(); IMAD_(); }
      if (MODE == 6) LOP3U_();
      if (MODE == 7) IMAD_();
      if (MODE == 8) { DFMA_(); LOP3U_(); LOP3V_(); IMAD_(); }
      if (MODE == 9) { DFMA_(); LOP3U_(); DADD_(); LOP3V_(); }
      if (MODE == 10) IMADW_();
      if (MODE == 11) { DFMA_(); IMADW_(); }
      if (MODE == 12) FSEL_();
      if (MODE == 13) { DFMA_(); FSEL_(); }
      if (MODE == 14) DSETP_();
      if (MODE == 15) { DFMA_(); LOP3U_(); DADD_(); FSEL_(); IMADW_(); }   // K2-like mix: 2 fp64 : 2 alu : 1 imad.wide
      if (MODE == 16) { DFMA_(); DADD_(); LOP3U_(); }                      // 2 fp64 : 1 alu
      if (MODE == 17) IMADHI_();
      if (MODE == 18) { IMADLO_(); IMADHI_(); }
      if (MODE == 19) { DFMA_(); LOP3U_(); DADD_(); FSEL_(); IMADLO_(); IMADHI_(); }   // the mix of mode 15 with the wide multiply split
    }
  }
  const long long c1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
  double s = 0; uint32_t w = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) { s += x[c]; w ^= u[c] ^ v[c] ^ __float_as_uint(f[c]); }
  if (s == 1.2345 || w == 0x12345u) { out[threadIdx.x] = s; outi[threadIdx.x] = w; }
}

// dependent-issue latency: one warp per SM, one chain
template <int OP>
__global__ void lat(double* out, long long* cyc, double a, double b, uint32_t k, int iters) {
  double x = a + threadIdx.x; uint32_t u = k + threadIdx.x, v = k;
  const long long c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
    if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b));
    if (OP == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(a));
    if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(v), "r"(k));
    if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u) : "r"(k), "r"(v));
    if (OP == 5) asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t;}" : "+r"(u), "+r"(v) : "r"(k));
  }
  const long long c1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
  if (x == 1.2345 || (u ^ v) == 0x12345u) out[threadIdx.x] = x + u;
}

static void latency(double mhz) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 8192 * 8);
  const int iters = 1 << 16;
  const char* names[6] = {"DFMA", "DADD", "DMUL", "LOP3", "IMAD", "IMAD.WIDE"};
#define LAT(OP) lat<OP><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 0x9E3779B9u, iters); cudaDeviceSynchronize(); \
  lat<OP><<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 0x9E3779B9u, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("dependent-issue latency %-10s %.2f cycles\n", names[OP], (double)h / iters);
  LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5)
  cudaFree(out); cudaFree(cyc);
}

template <int MODE>
void run(const char* name, int fp64_per, int other_per, int sms, double mhz) {
  double* out; uint32_t* outi;
  long long* cyc; static long long hcyc[8192];
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&outi, 4096 * 4); cudaMalloc(&cyc, 8192 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps_per_smsp : {1, 2, 4, 8}) {
    const int blocks = sms * warps_per_smsp / 2;   // 256-thread blocks = 8 warps = 2 per SMSP
    const int blk = warps_per_smsp >= 2 ? blocks : sms;
    const int thr = warps_per_smsp >= 2 ? 256 : 128;
    probe<MODE><<<blk, thr>>>(out, outi, cyc, 1.0000001, 1e-9, 0x9E3779B9u, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      probe<MODE><<<blk, thr>>>(out, outi, cyc, 1.0000001, 1e-9, 0x9E3779B9u, ITER);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaMemcpy(hcyc, cyc, blk * 8, cudaMemcpyDeviceToHost);
    double cycles = 0; for (int i = 0; i < blk; ++i) cycles += (double)hcyc[i]; cycles /= blk;   // SM cycles (clock64), mean over blocks
    const double mhz_eff = cycles / (best * 1e-3) * 1e-6;
    const double warp_instr_per_smsp = (double)warps_per_smsp * ITER * CH;   // per instruction kind
    printf("%-28s warps/SMSP=%d  %8.3f ms  cycles per (group of %d fp64 + %d other) per SMSP: %.3f", name, warps_per_smsp,
           best, fp64_per, other_per, cycles / warp_instr_per_smsp);
    if (fp64_per) printf("   fp64 warp-instr/cycle/SMSP: %.3f", fp64_per * warp_instr_per_smsp / cycles);
    printf("   (clock64/event-time = %.0f MHz)", mhz_eff);
    printf("\n");
  }
  cudaFree(out); cudaFree(outi); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  printf("device %s, %d SMs, clock attr %.0f MHz (cycle figures assume this clock)\n", p.name, p.multiProcessorCount, mhz);
  const int sms = p.multiProcessorCount;
  run<0>("DFMA", 1, 0, sms, mhz);
  run<1>("DADD", 1, 0, sms, mhz);
  run<2>("DMUL", 1, 0, sms, mhz);
  run<6>("LOP3", 0, 1, sms, mhz);
  run<7>("IMAD", 0, 1, sms, mhz);
  run<3>("DFMA+LOP3", 1, 1, sms, mhz);
  run<4>("DFMA+2LOP3", 1, 2, sms, mhz);
  run<5>("DFMA+IMAD", 1, 1, sms, mhz);
  run<8>("DFMA+2LOP3+IMAD", 1, 3, sms, mhz);
  run<9>("DFMA+LOP3+DADD+LOP3", 2, 2, sms, mhz);
  run<10>("IMAD.WIDE", 0, 1, sms, mhz);
  run<11>("DFMA+IMAD.WIDE", 1, 1, sms, mhz);
  run<12>("ISETP+FSEL", 0, 2, sms, mhz);
  run<13>("DFMA+ISETP+FSEL", 1, 2, sms, mhz);
  run<14>("DSETP+SEL", 1, 1, sms, mhz);
  run<15>("DFMA+LOP3+DADD+SETP+FSEL+IMAD.WIDE", 2, 4, sms, mhz);
  run<16>("DFMA+DADD+LOP3", 2, 1, sms, mhz);
  run<17>("IMAD.HI", 0, 1, sms, mhz);
  run<18>("IMAD.LO+IMAD.HI", 0, 2, sms, mhz);
  run<19>("DFMA+LOP3+DADD+SETP+FSEL+IMAD.LO+IMAD.HI", 2, 5, sms, mhz);
  latency(mhz);
  return 0;
}
