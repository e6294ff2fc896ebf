// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double* out, double a, double b) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma884(double* out, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(b), "d"(b), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA + DFMA interleaved, to see whether they share a pipe
__global__ void k_mixed(double* out, double a, double b) {
  double c0[8], c1[8], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; f[i] = i * 0.5; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { dmma884(c0[i], c1[i], a, b); f[i] = fma(f[i], a, b); }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM (1 block per SM... use blocks = sms*?)
    int threads = 256; int blocks = sms * (wps * 32 / threads > 0 ? wps * 32 / threads : 1);
    if (wps * 32 < threads) { threads = wps * 32; blocks = sms; }
    double nthreads = (double)threads * blocks;
    float ms;
    ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DFMA        : %8.2f TFLOP/s\n", wps, nthreads * 16.0 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_dmma884<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA884 x4  : %8.2f TFLOP/s\n", wps, nthreads / 32 * 4.0 * ITERS * 512 / ms / 1e9);
    ms = timeit([&] { k_dmma884<16><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA884 x16 : %8.2f TFLOP/s\n", wps, nthreads / 32 * 16.0 * ITERS * 512 / ms / 1e9);
    ms = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA16816   : %8.2f TFLOP/s\n", wps, nthreads / 32 * 8.0 * (ITERS / 8) * 4096 / ms / 1e9);
    ms = timeit([&] { k_mixed<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf("warps/SM %2d  mixed       : %8.2f TFLOP/s (dmma %.2f + dfma %.2f)\n", wps,
           nthreads / 32 * 8.0 * ITERS * (512 + 64) / ms / 1e9, nthreads / 32 * 8.0 * ITERS * 512 / ms / 1e9,
           nthreads / 32 * 8.0 * ITERS * 64 / ms / 1e9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
