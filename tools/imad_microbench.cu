// Integer-pipe microbenchmark for sm_100a: issue rate of the instructions a Montgomery
// multiplication can be built from.  Prints warp-instructions per clock per SM for each variant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/imad_microbench tools/imad_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

template <int V>
__global__ void bench(uint32_t* out, uint32_t seed) {
  uint32_t a[ILP], b[ILP], c[ILP];
  uint64_t w[ILP];
  double d[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) { a[k] = seed + threadIdx.x * 7 + k; b[k] = seed * 3 + k * 13 + 1; w[k] = a[k]; d[k] = a[k] * 1.0; c[k] = k; }
  const uint32_t m = seed | 1u;
  const double dm = 1.0000001 + seed;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      if (V == 0) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(b[k]) : "r"(a[(k + 1) % ILP]), "r"(m));
      if (V == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(b[k]));
      if (V == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(b[k]));
      if (V == 3) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(b[k]));
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(m), "r"(a[k])); }
      if (V == 4) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(b[k]) : "r"(a[(k + 1) % ILP]), "r"(m));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c[k]) : "r"(m), "r"(a[k])); }
      if (V == 5) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[k]) : "d"(dm));
      if (V == 6) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(b[k]) : "r"(a[(k + 1) % ILP]), "r"(m));
                    asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[k]) : "d"(dm)); }
      if (V == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k]));
      if (V == 8) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(b[k]) : "r"(a[(k + 1) % ILP]), "r"(m));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(c[k]) : "r"(a[k])); }
      if (V == 9) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(b[k]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(b[k]) : "r"(a[k])); }
      if (V == 10) { asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(a[k]) : "h"((uint16_t)a[k]), "h"((uint16_t)m)); }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) r ^= a[k] ^ b[k] ^ c[k] ^ (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32) ^ (uint32_t)d[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V>
void run(const char* name, int per_iter, uint32_t* d_out, int sms, double clk_ghz) {
  const int blocks = sms * 4, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<V><<<blocks, threads>>>(d_out, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<V><<<blocks, threads>>>(d_out, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)blocks * (threads / 32) * ITERS * ILP * per_iter;
  double cycles = ms * 1e-3 * clk_ghz * 1e9;
  printf("%-34s %8.3f ms  %6.2f warp-instr/clk/SM  (%5.1f lanes/clk/SM)\n", name, ms, warp_instr / cycles / sms, 32 * warp_instr / cycles / sms);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ghz = clk_khz / 1e6;
  printf("%s, %d SMs, clock attr %.3f GHz (rates assume this clock)\n", p.name, p.multiProcessorCount, ghz);
  uint32_t* d_out; cudaMalloc(&d_out, 148 * 8 * 256 * 4 * 4);
  int sms = p.multiProcessorCount;
  run<0>("IMAD.WIDE.U32", 1, d_out, sms, ghz);
  run<1>("IMAD (lo)", 1, d_out, sms, ghz);
  run<2>("IMAD.HI.U32", 1, d_out, sms, ghz);
  run<3>("IMAD lo + IMAD.HI pair", 2, d_out, sms, ghz);
  run<4>("IMAD.WIDE + IMAD lo pair", 2, d_out, sms, ghz);
  run<5>("DFMA", 1, d_out, sms, ghz);
  run<6>("IMAD.WIDE + DFMA pair", 2, d_out, sms, ghz);
  run<7>("IADD", 1, d_out, sms, ghz);
  run<8>("IMAD.WIDE + IADD pair", 2, d_out, sms, ghz);
  run<9>("IMAD.HI + IADD pair", 2, d_out, sms, ghz);
  return 0;
}
