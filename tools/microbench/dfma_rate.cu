// FP64 FMA issue rate of one SM sub-partition: independent DFMA chains, all warps busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rate dfma_rate.cu && ./dfma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int CH> __global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FIR-shaped: acc[r] += h[q] * x[(r + q) % 8], every operand a different register (what a register-blocked FP64 FIR issues)
__global__ void k_dfma_fir(double *out, const double *in, int iters) {
  double acc[8], x[8], h[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    acc[i] = 0.0;
    x[i] = in[threadIdx.x + 32 * i];
    h[i] = in[threadIdx.x + 32 * i + 256];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
#pragma unroll
      for (int r = 0; r < 8; r++) acc[r] = fma(h[q], x[(r + q) % 8], acc[r]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH> __global__ void k_ffma(float *out, int iters, float a, float b) {
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double *d;
  cudaMalloc(&d, sizeof(double) * sms * 4 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      k_dfma<8><<<sms, threads>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)sms * threads * 8.0 * iters;
      if (rep) printf("DFMA threads/SM %4d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM at %d MHz nominal\n", threads, ms,
                      2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
    }
  }
  cudaMemset(d, 0, sizeof(double) * 1024);
  for (int threads : {128, 256, 512}) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      k_dfma_fir<<<sms, threads>>>(d + 2048, d, iters / 8);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)sms * threads * 64.0 * (iters / 8);
      if (rep) printf("DFMA (FIR-shaped, 3 register operands) threads/SM %4d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM\n", threads, ms,
                      2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (khz * 1e3));
    }
  }
  for (int threads : {512, 1024}) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      k_ffma<8><<<sms, threads>>>((float *)d, iters, 1.0000001f, 1e-9f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)sms * threads * 8.0 * iters;
      if (rep) printf("FFMA threads/SM %4d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM\n", threads, ms, 2 * fma / ms * 1e-9,
                      fma / (ms * 1e-3) / sms / (khz * 1e3));
    }
  }
  return 0;
}
