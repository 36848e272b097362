// Throughput of Fq Montgomery multiplication on B200: integer (IMAD.WIDE) path, FP64 (DFMA) path, and both
// at once with the warps of each CTA split between the two pipes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mul_microbench tools/mul_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fq_f64.cuh"
#include "fq_karatsuba.cuh"
#include "fq_f64v2.cuh"
using namespace gm;

#define ITERS 512
#ifndef THREADS
#define THREADS 128
#endif

// mode 0: all warps integer; 1: all warps FP64; 2: warps with (warp % den) < num use FP64
__global__ void __launch_bounds__(THREADS, 512 / THREADS) bench(const Fq* in, Fq* out, int mode, int num, int den) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x = in[tid & 1023], y = in[(tid + 7) & 1023], b = in[(tid + 13) & 1023];
  const int warp = threadIdx.x >> 5;
  const bool fp = mode == 1 || (mode == 2 && (warp % den) < num);
  if (mode == 5 || (mode >= 6 && (warp % den) < num)) {
    // mode >= 6: the FP64 warps run ITERS * (mode - 5) / 8 iterations, so that a sweep finds the work ratio at which both
    // classes of warps finish together (= the aggregate rate of the two pipes running side by side)
    const int iters = mode >= 6 ? ITERS * (mode - 5) / 8 : ITERS;
    for (int it = 0; it < iters; it++) {
      f64v2::fq_mul(x.v, x.v, b.v);
      f64v2::fq_mul(y.v, y.v, b.v);
    }
  } else if (mode >= 6) {
    for (int it = 0; it < ITERS; it++) {
      x = x * b;
      y = y * b;
    }
  } else if (mode == 3) {
    for (int it = 0; it < ITERS; it++) {
      mont_mul_karatsuba<FqParams>(x.v, x.v, b.v);
      mont_mul_karatsuba<FqParams>(y.v, y.v, b.v);
    }
  } else if (mode == 4) {
    for (int it = 0; it < ITERS; it++) {
      mont_mul_karatsuba<FqParams, 2>(x.v, x.v, b.v);
      mont_mul_karatsuba<FqParams, 2>(y.v, y.v, b.v);
    }
  } else if (fp) {
    for (int it = 0; it < ITERS; it++) {
      f64::fq_mul_f64(x.v, x.v, b.v);
      f64::fq_mul_f64(y.v, y.v, b.v);
    }
  } else {
    for (int it = 0; it < ITERS; it++) {
      x = x * b;
      y = y * b;
    }
  }
  out[tid] = x + y;
}

// share of the warps of a CTA in class k (0 = integer, 1 = FP64) for a (num, den) split over THREADS / 32 warps
static double den_frac(int num, int den, int k) {
  int fp = 0, w = THREADS / 32;
  for (int i = 0; i < w; i++) fp += (i % den) < num;
  return k ? (double)fp / w : (double)(w - fp) / w;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8 * 128 / THREADS, threads = THREADS;
  Fq* in; Fq* out;
  cudaMalloc(&in, 1024 * sizeof(Fq)); cudaMalloc(&out, (size_t)blocks * threads * sizeof(Fq));
  Fq h[1024];
  for (int i = 0; i < 1024; i++) for (int j = 0; j < 12; j++) h[i].v[j] = (j == 11) ? (0x0a000000u + i) : (0x9e3779b9u * (i * 12 + j + 1));
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  struct { const char* name; int mode, num, den; } cfg[] = {
      {"integer only (IMAD.WIDE)", 0, 0, 1}, {"FP64 only (DFMA)", 1, 0, 1}, {"split 1/4 FP64", 2, 1, 4},
      {"split 2/4 FP64", 2, 2, 4}, {"split 3/4 FP64", 2, 3, 4}, {"integer, Karatsuba (1 level)", 3, 0, 1},
      {"integer, Karatsuba (2 levels)", 4, 0, 1},
      {"FP64 48-bit limbs only", 5, 0, 1},
      {"1/4 FP64 warps, 2/8 work", 7, 1, 4}, {"1/4 FP64 warps, 4/8 work", 9, 1, 4}, {"1/4 FP64 warps, 6/8 work", 11, 1, 4}, {"1/4 FP64 warps, 8/8 work", 13, 1, 4},
      {"2/4 FP64 warps, 2/8 work", 7, 2, 4}, {"2/4 FP64 warps, 4/8 work", 9, 2, 4}, {"2/4 FP64 warps, 6/8 work", 11, 2, 4}, {"2/4 FP64 warps, 8/8 work", 13, 2, 4},
      {"3/8 FP64 warps, 3/8 work", 8, 3, 8}, {"3/8 FP64 warps, 5/8 work", 10, 3, 8}, {"3/8 FP64 warps, 7/8 work", 12, 3, 8}};
  Fq ref[4];
  for (auto& c : cfg) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<<<blocks, threads>>>(in, out, c.mode, c.num, c.den);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<<<blocks, threads>>>(in, out, c.mode, c.num, c.den);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    Fq got[4]; cudaMemcpy(got, out, sizeof(got), cudaMemcpyDeviceToHost);
    bool same = true;
    if (c.mode == 0) for (int k = 0; k < 4; k++) ref[k] = got[k];
    else for (int k = 0; k < 4; k++) same = same && (got[k] == ref[k]);
    double muls = (double)blocks * threads * ITERS * 2;
    if (c.mode >= 6) muls = muls * (den_frac(c.num, c.den, 0) + den_frac(c.num, c.den, 1) * (c.mode - 5) / 8.0);
    printf("%-28s %8.3f ms  %7.2f G Fq-mul/s  %s (%s)\n", c.name, ms, muls / ms / 1e6, cudaGetErrorString(cudaGetLastError()), same ? "results match integer path" : "MISMATCH");
  }
  return 0;
}
