// tma_probe.cu — does cuTensorMapEncodeTiled accept the tensor map the fused front end wants (rank 5, rows that overlap
// in memory, strides that are not monotonic), and where does SWIZZLE_128B put the bytes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
// View of one channel's cf32 input: element (f, t, q, k, c) = float f of 128-byte chunk q of stream tile t of block k of
// channel c at byte offset 4 f + 3840 t + 128 q + 480000 k + stride c. Box {32, 32, 2, 1, 1}: 2 chunks of 32 tiles.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, float *out, int t0, int q0, int k, int c) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned mbar = base + 8192;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(8192u) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(base),
        "l"(&tm), "r"(0), "r"(t0), "r"(q0), "r"(k), "r"(c), "r"(mbar)
        : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mbar)
      : "memory");
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr) != cudaSuccess || !enc) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  const size_t stride = 1 << 20; // samples per channel
  const int C = 3, nblk = 2;
  std::vector<float> h(2 * stride * C);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 16777216);
  float *d, *dout;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&dout, 8192);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const size_t off = 6 * 2; // base offset in floats (16-byte aligned: even sample index)
  for (int order = 0; order < 2; order++) {
    CUtensorMap tm;
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
    if (order == 0) { // tiles before chunks (strides not monotonic): smem [chunk][tile][128 B]
      const cuuint64_t dd[5] = {32, 125, 44, (cuuint64_t)nblk, (cuuint64_t)C};
      const cuuint64_t ss[4] = {3840, 128, 480000, stride * 8};
      const cuuint32_t bb[5] = {32, 32, 2, 1, 1};
      for (int i = 0; i < 5; i++) dims[i] = dd[i], box[i] = bb[i];
      for (int i = 0; i < 4; i++) strides[i] = ss[i];
    } else { // chunks before tiles: smem [tile][chunk][128 B]
      const cuuint64_t dd[5] = {32, 44, 125, (cuuint64_t)nblk, (cuuint64_t)C};
      const cuuint64_t ss[4] = {128, 3840, 480000, stride * 8};
      const cuuint32_t bb[5] = {32, 2, 32, 1, 1};
      for (int i = 0; i < 5; i++) dims[i] = dd[i], box[i] = bb[i];
      for (int i = 0; i < 4; i++) strides[i] = ss[i];
    }
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d + off, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("order %d: encode -> %d\n", order, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 64);
    const int t0 = 96, q0 = 6, k = 1, c = 2; // tiles 96..127 (125..127 out of bounds -> zero)
    if (order == 0) {
      probe<<<1, 128, 8192 + 64>>>(tm, dout, t0, q0, k, c);
    } else {
      probe<<<1, 128, 8192 + 64>>>(tm, dout, q0, t0, k, c);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("order %d: kernel -> %s\n", order, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(2048);
    cudaMemcpy(o.data(), dout, 8192, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int tt = 0; tt < 32; tt++) {
      for (int q = 0; q < 2; q++) {
        const int line = (order == 0) ? q * 32 + tt : tt * 2 + q; // 128-byte line of the box
        for (int u = 0; u < 8; u++) {
          for (int f = 0; f < 4; f++) {
            const size_t g = off + (size_t)(4 * u + f) + 960 * (size_t)(t0 + tt) + 32 * (size_t)(q0 + q) + 120000 * (size_t)k +
                             2 * stride * (size_t)c;
            const float want = (t0 + tt < 125) ? h[g] : 0.f;
            const float got = o[line * 32 + ((u ^ (line & 7)) << 2) + f];
            if (got != want) {
              if (bad < 5) printf("  mismatch tile %d chunk %d unit %d: got %.0f want %.0f\n", tt, q, u, got, want);
              bad++;
            }
          }
        }
      }
    }
    printf("order %d: %s (%d mismatches)\n", order, bad ? "FAIL" : "swizzle model ok", bad);
  }
  return 0;
}
