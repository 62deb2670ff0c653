// fmr_mpf.cuh — the multipath equaliser (reference: MultipathFilter, MultipathFilter.cpp:
// 92-197; its use in FmDecoder::process, FmDecode.cpp:107-128).
//
// Sample-serial NLMS/CMA exactly as the reference runs it: FIR output for every sample,
// coefficient update on every 4th sample OF THE CURRENT process() CALL ((i & 3) == 0),
// step size renormalised from the instantaneous state energy, reference tap pinned to
// 1+0j, failure (non-finite output or error) -> coefficients re-initialised and the whole
// call passed through unfiltered. One warp owns one channel: the 4*stages+1 complex
// coefficients live in registers (tap k on lane k%32), the delay line in a shared-memory
// ring, the dot product is reduced with warp shuffles, so a sample costs no block barrier.
// The ring is stored twice back to back so that the N-tap window is always contiguous (tap
// loads with immediate offsets, no per-tap wrap arithmetic); a lane's partial dot product runs
// as two independent accumulator chains (the summation order differs from VOLK's SIMD order
// either way); the window stays in registers for the coefficient update; input samples are
// prefetched one batch of eight ahead.
//
// Batching: the coefficients only change after the samples i = 0, 4, 8, ... of a call, so the four outputs i = 4q+1 ..
// 4q+4 all use the coefficients left by update q and are evaluated TOGETHER (four windows one sample apart, sixteen
// independent accumulator chains, eight interleaved warp reductions), then update q+1 runs on the window and output of
// i = 4q+4. Every output and update is the same expression on the same operands as in the sample-by-sample order
// (MultipathFilter.cpp:176-186); only the latency of one sample's dependent chain is no longer paid four times.
#ifndef FMR_MPF_CUH
#define FMR_MPF_CUH

#include "fmr_host.cuh"

namespace fmr {

constexpr int kMpfRing = 1024;
constexpr int kMpfWarps = 3; // 3 x 16 KB of mirrored ring = the 48 KB static shared-memory limit; 4 CTAs per SM

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// MultipathFilter::update_coeff (MultipathFilter.cpp:108-161) on the window `sv` and the output (yr, yi). A free
// __forceinline__ function (not a lambda called from two places): the coefficient and window arrays must stay in
// registers. The window registers of taps k >= N hold whatever older samples the ring has there: their coefficients are
// kept at zero HERE, so the output loops can run over whole rows without a per-tap test.
template <int J>
__device__ __forceinline__ bool mpf_update(float2 (&cf)[J], float2 (&sv)[J], float yr, float yi, int lane, int N, int jf,
                                           int ref_idx, double &err_keep) {
#pragma unroll
  for (int j = 0; j < J; j++) {
    if (j >= jf && lane + 32 * j >= N) sv[j] = make_float2(0.f, 0.f);
  }
  float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < J; j++) m4[j & 3] = fmaf(sv[j].y, sv[j].y, fmaf(sv[j].x, sv[j].x, m4[j & 3]));
  const float ms = warp_sum((m4[0] + m4[1]) + (m4[2] + m4[3]));
  const double env = (double)(yr * yr + yi * yi);
  const double err = 1.0 - env;
  const float mu = (float)(0.1 / ((double)ms + 1e-10));
  const float factor = (float)(err * (double)mu);
  const float fr = factor * yr, fi = factor * yi;
#pragma unroll
  for (int j = 0; j < J; j++) {
    cf[j].x = fmaf(fi, sv[j].y, fmaf(fr, sv[j].x, cf[j].x));
    cf[j].y = fmaf(-fr, sv[j].y, fmaf(fi, sv[j].x, cf[j].y));
  }
  // the reference tap is pinned to 1+0j (MultipathFilter.cpp:158-160)
  if (lane == (ref_idx & 31)) {
#pragma unroll
    for (int j = 0; j < J; j++) {
      if (j == (ref_idx >> 5)) cf[j] = make_float2(1.f, 0.f);
    }
  }
  err_keep = err;
  return isfinite(err);
}

// One sample: ring write, output, and the update when `upd`; false = the reference's failure path
// (MultipathFilter.cpp:163-195). Free __forceinline__ function for the same reason as mpf_update.
template <int J>
__device__ __forceinline__ bool mpf_step1(float2 (&cf)[J], float2 *ring, uint32_t &cnt, float2 x, bool upd, int lane, int N,
                                          int jf, int ref_idx, double &err_keep, Ring<float2> out, int c, int64_t t_out) {
  constexpr uint32_t mask = kMpfRing - 1;
  cnt++;
  if (lane == 0) {
    ring[cnt & mask] = x;
    ring[(cnt & mask) + kMpfRing] = x;
  }
  __syncwarp();
  // tap k = lane + 32 j reads the sample N-1-k steps behind the newest: contiguous from `w`
  const float2 *__restrict__ w = ring + ((cnt - (uint32_t)(N - 1)) & mask) + lane;
  float2 sv[J];
  // (the same two accumulator chains per output as in the batch: a sample's value does not depend on which of the two
  // paths the call partition sends it through)
  float ar[2] = {0.f, 0.f}, ai[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < J; j++) {
    // (J = 32 only: the last row would read two slots past the mirrored ring)
    sv[j] = (J < 32 || j < J - 1 || lane + 32 * j < N) ? w[32 * j] : make_float2(0.f, 0.f);
    ar[j & 1] = fmaf(-sv[j].y, cf[j].y, fmaf(sv[j].x, cf[j].x, ar[j & 1]));
    ai[j & 1] = fmaf(sv[j].y, cf[j].x, fmaf(sv[j].x, cf[j].y, ai[j & 1]));
  }
  const float yr = warp_sum(ar[0] + ar[1]), yi = warp_sum(ai[0] + ai[1]);
  if (!isfinite(yr) || !isfinite(yi)) return false;
  if (lane == 0) out.st(c, t_out, make_float2(yr, yi));
  return upd ? mpf_update<J>(cf, sv, yr, yi, lane, N, jf, ref_idx, err_keep) : true;
}

template <int J>
__global__ void __launch_bounds__(32 * kMpfWarps, 4)
    k_mpf(Ring<float2> in, Ring<float2> out, FmChanState *__restrict__ st, float2 *__restrict__ g_coeff,
          float2 *__restrict__ g_state, int N, int ref_idx, const uint32_t *__restrict__ call_end, int n_calls,
          int64_t t0, int C) {
  __shared__ float2 ring_all[kMpfWarps][2 * kMpfRing];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kMpfWarps + warp;
  if (c >= C) return;
  float2 *ring = ring_all[warp];
  float2 cf[J];
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int k = lane + 32 * j;
    cf[j] = (k < N) ? g_coeff[(size_t)c * kMpfRing + k] : make_float2(0.f, 0.f);
  }
  for (int k = lane; k < kMpfRing; k += 32) {
    const float2 v = (k < N) ? g_state[(size_t)c * kMpfRing + k] : make_float2(0.f, 0.f);
    ring[k] = v;
    ring[k + kMpfRing] = v;
  }
  __syncwarp();
  uint32_t cnt = (uint32_t)(N - 1); // ring index of the newest sample
  const int jf = N >> 5;            // number of full rows of 32 taps
  uint32_t wait = st[c].mpf_wait;
  double err_keep = st[c].mpf_error;
  uint32_t prev_end = 0;
  for (int b = 0; b < n_calls; b++) {
    const uint32_t end = call_end[b];
    const int n = (int)(end - prev_end);
    if (n == 0) continue;
    const int64_t tb = t0 + prev_end;
    prev_end = end;
    if (wait > 0) {
      // FmDecode.cpp:107-110: bypass, filter state untouched
      wait--;
      for (int q = lane; q < n; q += 32) out.st(c, tb + q, in.ld(c, tb + q));
      continue;
    }
    bool ok = true;
    const uint32_t mask = kMpfRing - 1;
    // sample 0 of the call: output with the carried coefficients, then the first update
    ok = mpf_step1<J>(cf, ring, cnt, in.ld(c, tb), true, lane, N, jf, ref_idx, err_keep, out, c, tb);
    int i = 1;
    // batches of four: outputs i .. i+3 with the same coefficients, update after i+3 (i = 1 mod 4). Lanes 0..3 fetch the
    // batch's samples one batch ahead, so the global-load latency hides behind the previous batch's arithmetic.
    float2 xnext = (lane < 4 && 1 + lane < n) ? in.ld(c, tb + 1 + lane) : make_float2(0.f, 0.f);
    for (; ok && i + 3 < n; i += 4) {
      const float2 xq = xnext;
      xnext = (lane < 4 && i + 4 + lane < n) ? in.ld(c, tb + i + 4 + lane) : make_float2(0.f, 0.f);
      if (lane < 4) {
        const uint32_t pz = (cnt + 1 + lane) & mask;
        ring[pz] = xq;
        ring[pz + kMpfRing] = xq;
      }
      cnt += 4;
      __syncwarp();
      // window of output d (d = 0..3): newest sample cnt - 3 + d; tap k reads slot w[32 j + d]
      const float2 *__restrict__ w = ring + ((cnt - 3 - (uint32_t)(N - 1)) & mask) + lane;
      float2 sv[J]; // window of the fourth output (the one the update uses)
      float ar[4][2], ai[4][2];
#pragma unroll
      for (int d = 0; d < 4; d++) ar[d][0] = ar[d][1] = ai[d][0] = ai[d][1] = 0.f;
#pragma unroll
      for (int j = 0; j < J; j++) {
#pragma unroll
        for (int d = 0; d < 4; d++) {
          const float2 v = (J < 32 || j < J - 1 || lane + 32 * j < N) ? w[32 * j + d] : make_float2(0.f, 0.f);
          if (d == 3) sv[j] = v;
          ar[d][j & 1] = fmaf(-v.y, cf[j].y, fmaf(v.x, cf[j].x, ar[d][j & 1]));
          ai[d][j & 1] = fmaf(v.y, cf[j].x, fmaf(v.x, cf[j].y, ai[d][j & 1]));
        }
      }
      float yr[4], yi[4];
#pragma unroll
      for (int d = 0; d < 4; d++) {
        yr[d] = ar[d][0] + ar[d][1];
        yi[d] = ai[d][0] + ai[d][1];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 4; d++) {
          yr[d] += __shfl_xor_sync(0xffffffffu, yr[d], o);
          yi[d] += __shfl_xor_sync(0xffffffffu, yi[d], o);
        }
      }
#pragma unroll
      for (int d = 0; d < 4; d++) {
        if (ok && (!isfinite(yr[d]) || !isfinite(yi[d]))) ok = false;
      }
      if (!ok) break;
      if (lane < 4) {
        float2 y = make_float2(yr[0], yi[0]);
#pragma unroll
        for (int d = 1; d < 4; d++) {
          if (lane == d) y = make_float2(yr[d], yi[d]);
        }
        out.st(c, tb + i + lane, y);
      }
      ok = mpf_update<J>(cf, sv, yr[3], yi[3], lane, N, jf, ref_idx, err_keep);
    }
    // the last (n - 1) mod 4 samples of the call: outputs only (already fetched: lane d holds sample i + d)
    for (int d = 0; ok && i < n; i++, d++) {
      const float2 x = make_float2(__shfl_sync(0xffffffffu, xnext.x, d), __shfl_sync(0xffffffffu, xnext.y, d));
      ok = mpf_step1<J>(cf, ring, cnt, x, false, lane, N, jf, ref_idx, err_keep, out, c, tb + i);
    }
    if (!ok) {
      // FmDecode.cpp:114-123: reset coefficients, pass the call through unfiltered
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int k = lane + 32 * j;
        cf[j] = (k == ref_idx) ? make_float2(1.f, 0.f) : make_float2(0.f, 0.f);
      }
      __syncwarp();
      for (int q = lane; q < n; q += 32) out.st(c, tb + q, in.ld(c, tb + q));
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int k = lane + 32 * j;
    if (k < N) g_coeff[(size_t)c * kMpfRing + k] = cf[j];
  }
  // linearise the delay line: entry k is the sample N-1-k steps behind the newest
  float2 tmp[J];
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int k = lane + 32 * j;
    tmp[j] = (k < N) ? ring[(cnt - (uint32_t)(N - 1 - k)) & (kMpfRing - 1)] : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int k = lane + 32 * j;
    if (k < N) g_state[(size_t)c * kMpfRing + k] = tmp[j];
  }
  if (lane == 0) {
    st[c].mpf_wait = wait;
    st[c].mpf_error = err_keep;
  }
}

struct MpfDev {
  int N = 0, ref_idx = 0, C = 0, J = 0;
  float2 *d_coeff = nullptr, *d_state = nullptr;

  fmr_status init(uint32_t stages, int channels, DevMem &mem) {
    // MultipathFilter::MultipathFilter (MultipathFilter.cpp:35-75)
    N = (int)stages * 4 + 1;
    ref_idx = (int)stages * 3 + 1;
    C = channels;
    if (N > kMpfRing - 3) return fail(FMR_ERR_UNSUPPORTED, "multipath_stages > 255 is not supported");
    const int need = (N + 31) / 32;
    const int opts[] = {4, 8, 13, 16, 20, 26, 32};
    J = 32;
    for (int o : opts) {
      if (o >= need) {
        J = o;
        break;
      }
    }
    FMR_CUDA(mem.alloc(&d_coeff, (size_t)C * kMpfRing));
    FMR_CUDA(mem.alloc(&d_state, (size_t)C * kMpfRing));
    std::vector<float2> init((size_t)C * kMpfRing, make_float2(0.f, 0.f));
    for (int c = 0; c < C; c++) init[(size_t)c * kMpfRing + ref_idx] = make_float2(1.f, 0.f);
    FMR_CUDA(cudaMemcpy(d_coeff, init.data(), init.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return FMR_OK;
  }

  // channels [c0, c0+cn) of the handle; in/out/st are already offset to c0 by the caller
  void run(Ring<float2> in, Ring<float2> out, FmChanState *st, const uint32_t *call_end, int n_calls, int64_t t0,
           cudaStream_t s, int c0, int cn) {
    float2 *d_coeff = this->d_coeff + (size_t)c0 * kMpfRing;
    float2 *d_state = this->d_state + (size_t)c0 * kMpfRing;
    const int C = cn;
    dim3 grid((C + kMpfWarps - 1) / kMpfWarps);
    dim3 block(32 * kMpfWarps);
#define FMR_MPF_CASE(j)                                                                                          \
  case j:                                                                                                        \
    cudaFuncSetAttribute(k_mpf<j>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    k_mpf<j><<<grid, block, 0, s>>>(in, out, st, d_coeff, d_state, N, ref_idx, call_end, n_calls, t0, C);         \
    break;
    switch (J) {
      FMR_MPF_CASE(4)
      FMR_MPF_CASE(8)
      FMR_MPF_CASE(13)
      FMR_MPF_CASE(16)
      FMR_MPF_CASE(20)
      FMR_MPF_CASE(26)
    default:
      cudaFuncSetAttribute(k_mpf<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      k_mpf<32><<<grid, block, 0, s>>>(in, out, st, d_coeff, d_state, N, ref_idx, call_end, n_calls, t0, C);
      break;
    }
#undef FMR_MPF_CASE
  }

  fmr_status read_coeffs(uint32_t channel, float *re_im, size_t n_complex) {
    if ((int)n_complex < N) return fail(FMR_ERR_CAPACITY, "coefficient buffer smaller than 4*stages+1");
    FMR_CUDA(cudaMemcpy(re_im, d_coeff + (size_t)channel * kMpfRing, (size_t)N * sizeof(float2),
                        cudaMemcpyDeviceToHost));
    return FMR_OK;
  }
};

} // namespace fmr
#endif
