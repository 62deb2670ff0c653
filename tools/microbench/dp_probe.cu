// Microbenchmark: FP64 FMA dependent-issue latency and throughput, sincos(double) cost on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_dfma(double *o, int n, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = fma(x, a, b);
  long long t1 = clock64();
  if (threadIdx.x == 0) { o[0] = x; o[1] = (double)(t1 - t0) / n; }
}
__global__ void lat_ffma(double *o, int n, float a, float b) {
  float x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = fmaf(x, a, b);
  long long t1 = clock64();
  if (threadIdx.x == 0) { o[0] = x; o[1] = (double)(t1 - t0) / n; }
}
__global__ void lat_sincos(double *o, int n) {
  double x = 0.3 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { double s, c; sincos(x, &s, &c); x = x + s * 1e-3 + c * 1e-3; }
  long long t1 = clock64();
  if (threadIdx.x == 0) { o[0] = x; o[1] = (double)(t1 - t0) / n; }
}
__global__ void lat_cvt(double *o, int n) {
  float x = 0.3f + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { double d = (double)x; d = d * 1.0000001; x = (float)d; }
  long long t1 = clock64();
  if (threadIdx.x == 0) { o[0] = x; o[1] = (double)(t1 - t0) / n; }
}
__global__ void thr_dfma(double *o, int n, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < n; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  o[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  double *d; cudaMalloc(&d, 1 << 24);
  double h[2];
  lat_dfma<<<1, 32>>>(d, 100000, 1.0000001, 1e-9); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("DFMA dependent latency: %.1f cycles\n", h[1]);
  lat_ffma<<<1, 32>>>(d, 100000, 1.0000001f, 1e-9f); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("FFMA dependent latency: %.1f cycles\n", h[1]);
  lat_sincos<<<1, 32>>>(d, 20000); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("sincos(double)+2 DFMA chain: %.1f cycles\n", h[1]);
  lat_cvt<<<1, 32>>>(d, 100000); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("F2F.64 + DMUL + F2F.32 chain: %.1f cycles\n", h[1]);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int n = 20000, blocks = 148 * 8, threads = 256;
  thr_dfma<<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
  cudaEventRecord(a); thr_dfma<<<blocks, threads>>>(d, n, 1.0000001, 1e-9); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("DFMA throughput: %.2f TDFMA/s (%.2f TFLOP/s)\n", (double)blocks * threads * 8.0 * n / (ms * 1e-3) / 1e12, 2.0 * blocks * threads * 8.0 * n / (ms * 1e-3) / 1e12);
  return 0;
}
