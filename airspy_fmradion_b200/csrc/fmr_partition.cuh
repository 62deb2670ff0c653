// fmr_partition.cuh — SM partitioning with CUDA green contexts.
//
// The serial recurrences of the path (AGC, PLL, DC block) are one warp per 32 channels and
// bound by dependent-issue latency. Run concurrently with the throughput kernels on the same SMs
// they lose their issue slots and slow down 2-9x (measured, FMR_TRACE), so plain stream overlap
// buys nothing. A green context gives them a small private set of SMs; the throughput kernels
// get the rest. Driver entry points are resolved through cudaGetDriverEntryPoint, so the
// library does not link libcuda and still loads on a machine without a driver.
#ifndef FMR_PARTITION_CUH
#define FMR_PARTITION_CUH

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

namespace fmr {

struct SmPartition {
  bool ok = false;
  int sms_serial = 0, sms_main = 0;
  CUgreenCtx g_main = nullptr, g_serial = nullptr;
  cudaStream_t s_front = nullptr; // main partition: front end
  cudaStream_t s_post = nullptr;  // main partition: IF filter, multipath, discriminator, statistics
  cudaStream_t s_post2 = nullptr; // main partition: audio resamplers and pilot-cut FIR
  cudaStream_t s_serial = nullptr; // serial partition: AGC
  cudaStream_t s_serial2 = nullptr; // serial partition: PLL
  cudaStream_t s_serial3 = nullptr; // serial partition: DC block / matrix

  template <typename F> static bool entry(const char *name, F *fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
  }

  // Returns true when the partition was created; on any failure the handle simply runs unpartitioned.
  bool init(int device, int want_serial_sms, bool verbose) {
    typedef CUresult (*pfnDeviceGet)(CUdevice *, int);
    typedef CUresult (*pfnGetRes)(CUdevice, CUdevResource *, CUdevResourceType);
    typedef CUresult (*pfnSplit)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int,
                                 unsigned int);
    typedef CUresult (*pfnGenDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
    typedef CUresult (*pfnGreenCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*pfnGreenStream)(CUstream *, CUgreenCtx, unsigned int, int);
    pfnDeviceGet fDeviceGet;
    pfnGetRes fGetRes;
    pfnSplit fSplit;
    pfnGenDesc fGenDesc;
    pfnGreenCreate fGreenCreate;
    pfnGreenStream fGreenStream;
    if (!entry("cuDeviceGet", &fDeviceGet) || !entry("cuDeviceGetDevResource", &fGetRes) ||
        !entry("cuDevSmResourceSplitByCount", &fSplit) || !entry("cuDevResourceGenerateDesc", &fGenDesc) ||
        !entry("cuGreenCtxCreate", &fGreenCreate) || !entry("cuGreenCtxStreamCreate", &fGreenStream)) {
      if (verbose) fprintf(stderr, "[fmr] green-context entry points unavailable\n");
      return false;
    }
    CUdevice dev;
    if (fDeviceGet(&dev, device) != CUDA_SUCCESS) return false;
    CUdevResource full, grp, rest;
    if (fGetRes(dev, &full, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned int n = 1;
    if (fSplit(&grp, &n, &full, &rest, 0, (unsigned int)want_serial_sms) != CUDA_SUCCESS || n != 1) return false;
    CUdevResourceDesc d_serial, d_main;
    if (fGenDesc(&d_serial, &grp, 1) != CUDA_SUCCESS || fGenDesc(&d_main, &rest, 1) != CUDA_SUCCESS) return false;
    if (fGreenCreate(&g_serial, d_serial, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (fGreenCreate(&g_main, d_main, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream a, b, b2, c, c2, c3;
    if (fGreenStream(&a, g_main, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    // the post-processing streams outrank the front end: their kernels sit between two serial
    // stages, so every microsecond they queue behind front-end CTAs stalls the serial pipeline
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (fGreenStream(&b, g_main, CU_STREAM_NON_BLOCKING, prio_greatest) != CUDA_SUCCESS) return false;
    if (fGreenStream(&b2, g_main, CU_STREAM_NON_BLOCKING, prio_greatest) != CUDA_SUCCESS) return false;
    if (fGreenStream(&c, g_serial, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    if (fGreenStream(&c2, g_serial, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    if (fGreenStream(&c3, g_serial, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    s_front = (cudaStream_t)a;
    s_post = (cudaStream_t)b;
    s_post2 = (cudaStream_t)b2;
    s_serial = (cudaStream_t)c;
    s_serial2 = (cudaStream_t)c2;
    s_serial3 = (cudaStream_t)c3;
    sms_serial = (int)grp.sm.smCount;
    sms_main = (int)rest.sm.smCount;
    ok = true;
    if (verbose) fprintf(stderr, "[fmr] SM partition: %d SMs serial, %d SMs main\n", sms_serial, sms_main);
    return true;
  }
};

} // namespace fmr
#endif
