// Dependent-issue latency of the instruction classes on the serial recurrences (one warp).
// nvcc -arch=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
#define TEST(name, decl, body, sink)                                                   \
  __global__ void k_##name(long long *t, double *o) {                                  \
    decl;                                                                              \
    long long t0 = clock64();                                                          \
    _Pragma("unroll 16") for (int i = 0; i < N_IT; i++) { body; }                      \
    long long t1 = clock64();                                                          \
    if (threadIdx.x == 0) t[0] = t1 - t0;                                              \
    o[threadIdx.x] = (double)(sink);                                                   \
  }
__device__ double g_d = 1.000001;
__device__ float g_f = 1.000001f;
TEST(ffma, float a = g_f; float b = g_f * 0.5f, a = fmaf(a, b, 0.25f), a)
TEST(dfma, double a = g_d; double b = g_d * 0.5, a = fma(a, b, 0.25), a)
TEST(dadd, double a = g_d; double b = g_d * 0.5, a = a + b, a)
TEST(dmul, double a = g_d; double b = g_d, a = a * b, a)
TEST(f2f_rt, float a = g_f, a = (float)((double)a + 1e-9), a)            // F2F.F64.F32 + DADD + F2F.F32.F64
TEST(f2f_pair, float a = g_f, a = (float)(double)a * 1.0000001f, a)       // may fold; see SASS
TEST(mufu_rcp, float a = g_f; float r, asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); a = r, a)
TEST(f2i_i2f, float a = g_f * 100.f, a = (float)((int)a) + 0.5f, a)
TEST(dsetp_sel, double a = g_d; double b = g_d * 3, a = (a < b) ? a + 1e-9 : b, a)
TEST(fsel, float a = g_f; float b = g_f * 3, a = (a < b) ? a + 1e-6f : b, a)
TEST(fmnmx, float a = g_f; float b = g_f * 3, a = fminf(a, b) + 1e-6f, a)
__global__ void k_lds(long long *t, double *o) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i + 32) & 1023;
  __syncthreads();
  int a = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) a = s[a];
  long long t1 = clock64();
  if (threadIdx.x == 0) t[0] = t1 - t0;
  o[threadIdx.x] = a;
}
__global__ void k_ldg(long long *t, double *o, const int *g) {
  int a = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) a = g[a];
  long long t1 = clock64();
  if (threadIdx.x == 0) t[0] = t1 - t0;
  o[threadIdx.x] = a;
}
__global__ void k_div(long long *t, double *o) {
  float a = g_f, b = g_f * 3.f;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) a = a / b + 1.0f;
  long long t1 = clock64();
  if (threadIdx.x == 0) t[0] = t1 - t0;
  o[threadIdx.x] = a;
}
int main() {
  long long *t; double *o; int *g;
  cudaMalloc(&t, 8); cudaMalloc(&o, 32 * 8); cudaMalloc(&g, 4096);
  int h[1024]; for (int i = 0; i < 1024; i++) h[i] = (i + 32) & 1023;
  cudaMemcpy(g, h, 4096, cudaMemcpyHostToDevice);
  long long ht;
#define RUN(name, ...)                                                    \
  for (int r = 0; r < 2; r++) { k_##name<<<1, 32>>>(t, o, ##__VA_ARGS__); cudaDeviceSynchronize(); } \
  cudaMemcpy(&ht, t, 8, cudaMemcpyDeviceToHost);                          \
  printf("%-10s %.2f cycles/iter\n", #name, (double)ht / N_IT);
  RUN(ffma) RUN(dfma) RUN(dadd) RUN(dmul) RUN(f2f_rt) RUN(f2f_pair) RUN(mufu_rcp) RUN(f2i_i2f) RUN(dsetp_sel) RUN(fsel) RUN(fmnmx) RUN(lds) RUN(ldg, g) RUN(div)
  return 0;
}
