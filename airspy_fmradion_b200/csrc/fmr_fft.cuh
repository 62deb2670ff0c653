// fmr_fft.cuh — the long zero-phase low-pass as overlap-save fast convolution, entirely in
// shared memory (reference: r8b::CDSPBlockConvolver::process, CDSPBlockConvolver.h:252-353,
// which does the same linear convolution with a 16384-point double-precision FFT).
//
// One CTA filters one block of one channel: 16384 complex FP32 points live in 136 KB of
// shared memory; forward FFT, multiply by the filter spectrum H (precomputed on the host
// in double, 1/N folded in), inverse FFT via the conjugation identity, and only the
// 16384-(klen-1) alias-free outputs are written. The complex IQ stream rides through one
// complex FFT (the filter is real, so I and Q need no separate transforms).
//
// FFT: in-place Stockham autosort, radix 16,16,16,4 (DIT twiddles on the inputs), every
// thread keeps its butterfly inputs in registers across the barrier so one buffer
// suffices. The first forward pass reads straight from the global ring, the last forward
// pass applies H and the conjugation, the last inverse pass writes straight to global:
// 7 shared-memory round trips per block. Shared-memory layout is skewed (one pad slot per
// 16) so both the stride-1024 reads and the stride-16 writes are bank-conflict free.
#ifndef FMR_FFT_CUH
#define FMR_FFT_CUH

#include "fmr_kernels.cuh"

namespace fmr {

constexpr int kFftN = 16384;
constexpr int kFftThreads = 512;
constexpr int kFftSmemBytes = (kFftN + kFftN / 16) * 8 + 256 * 8;

__device__ __forceinline__ int fpad(int n) { return n + (n >> 4); }
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y);
  const float2 d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
  const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y);
  const float2 d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
  a0 = make_float2(s02.x + s13.x, s02.y + s13.y);
  a2 = make_float2(s02.x - s13.x, s02.y - s13.y);
  a1 = make_float2(d02.x + d13.y, d02.y - d13.x); // d02 - j d13
  a3 = make_float2(d02.x - d13.y, d02.y + d13.x); // d02 + j d13
}

// 16-point forward DFT in registers; result X[m] is left in v[4*(m&3) + (m>>2)].
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) fft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
  // twiddles W16^(n2*k1), element v[n2 + 4*k1]
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
  // k1 = 1: W^n2
  v[1 + 4] = cmulf(v[1 + 4], make_float2(c1, -s1));
  v[2 + 4] = cmulf(v[2 + 4], make_float2(r2, -r2));
  v[3 + 4] = cmulf(v[3 + 4], make_float2(s1, -c1));
  // k1 = 2: W^(2 n2)
  v[1 + 8] = cmulf(v[1 + 8], make_float2(r2, -r2));
  v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x); // W^4 = -j
  v[3 + 8] = cmulf(v[3 + 8], make_float2(-r2, -r2));
  // k1 = 3: W^(3 n2)
  v[1 + 12] = cmulf(v[1 + 12], make_float2(s1, -c1));
  v[2 + 12] = cmulf(v[2 + 12], make_float2(-r2, -r2));
  v[3 + 12] = cmulf(v[3 + 12], make_float2(-c1, s1)); // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) fft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// W_16384^m from the two-level table: tw[0..127] = W^(128 q), tw[128..255] = W^l.
__device__ __forceinline__ float2 tw_lookup(const float2 *tw, int m) {
  return cmulf(tw[m >> 7], tw[128 + (m & 127)]);
}

// apply w^r (r = 1..15) to v[r]
__device__ __forceinline__ void twiddle16(float2 (&v)[16], float2 w1) {
  const float2 w2 = cmulf(w1, w1);
  const float2 w3 = cmulf(w2, w1);
  const float2 w4 = cmulf(w2, w2);
  const float2 w5 = cmulf(w4, w1);
  const float2 w6 = cmulf(w3, w3);
  const float2 w7 = cmulf(w4, w3);
  const float2 w8 = cmulf(w4, w4);
  v[1] = cmulf(v[1], w1);
  v[2] = cmulf(v[2], w2);
  v[3] = cmulf(v[3], w3);
  v[4] = cmulf(v[4], w4);
  v[5] = cmulf(v[5], w5);
  v[6] = cmulf(v[6], w6);
  v[7] = cmulf(v[7], w7);
  v[8] = cmulf(v[8], w8);
  v[9] = cmulf(v[9], cmulf(w8, w1));
  v[10] = cmulf(v[10], cmulf(w8, w2));
  v[11] = cmulf(v[11], cmulf(w8, w3));
  v[12] = cmulf(v[12], cmulf(w8, w4));
  v[13] = cmulf(v[13], cmulf(w8, w5));
  v[14] = cmulf(v[14], cmulf(w8, w6));
  v[15] = cmulf(v[15], cmulf(w8, w7));
}

// One radix-16 Stockham pass over the whole buffer, in place: p = 16 or 256.
__device__ __forceinline__ void pass16_smem(float2 *buf, const float2 *tw, int p, int scale) {
  float2 va[16], vb[16];
  const int ia = threadIdx.x, ib = threadIdx.x + kFftThreads;
#pragma unroll
  for (int r = 0; r < 16; r++) {
    va[r] = buf[fpad(ia + r * 1024)];
    vb[r] = buf[fpad(ib + r * 1024)];
  }
  __syncthreads();
  {
    const int k = ia & (p - 1);
    twiddle16(va, tw_lookup(tw, k * scale));
    fft16(va);
    const int j = (ia - k) * 16 + k;
#pragma unroll
    for (int r = 0; r < 16; r++) buf[fpad(j + r * p)] = va[4 * (r & 3) + (r >> 2)];
  }
  {
    const int k = ib & (p - 1);
    twiddle16(vb, tw_lookup(tw, k * scale));
    fft16(vb);
    const int j = (ib - k) * 16 + k;
#pragma unroll
    for (int r = 0; r < 16; r++) buf[fpad(j + r * p)] = vb[4 * (r & 3) + (r >> 2)];
  }
  __syncthreads();
}

// k_fir_fft: y[q] = sum_j h[j] x[q*down - fl2 + j] for q in [q0, q0+n_out), block-wise.
//   n_in_avail : number of valid input samples in the ring (indices >= it read as zero)
//   lq         : outputs per block = (16384 - klen + 1) / down
static __global__ void __launch_bounds__(kFftThreads, 1)
    k_fir_fft(Ring<float2> in, Ring<float2> out, const float2 *__restrict__ H, int klen, int down, int64_t q0,
              int n_out, int64_t n_in_avail, int lq) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);
  float2 *tw = buf + (kFftN + kFftN / 16);
  const uint32_t c = blockIdx.y;
  const int blk = blockIdx.x;
  int cnt = n_out - blk * lq;
  if (cnt > lq) cnt = lq;
  if (cnt <= 0) return;
  const int fl2 = (klen - 1) / 2;
  const int64_t t_first = (q0 + (int64_t)blk * lq) * down;
  const int64_t base = t_first - fl2; // input sample index held by buffer slot 0
  // twiddle tables
  if (threadIdx.x < 256) {
    const int q = threadIdx.x;
    const int m = (q < 128) ? (q * 128) : (q - 128);
    float s, co;
    sincospif(-2.0f * (float)m / 16384.0f, &s, &co);
    tw[q] = make_float2(co, s);
  }
  // ---- forward pass 1 (p = 1, no twiddles), inputs straight from the global ring
  {
    float2 va[16], vb[16];
    const int ia = threadIdx.x, ib = threadIdx.x + kFftThreads;
#pragma unroll
    for (int r = 0; r < 16; r++) {
      const int64_t ta = base + ia + r * 1024;
      const int64_t tb = base + ib + r * 1024;
      va[r] = (ta < n_in_avail) ? in.ld(c, ta) : make_float2(0.f, 0.f);
      vb[r] = (tb < n_in_avail) ? in.ld(c, tb) : make_float2(0.f, 0.f);
    }
    fft16(va);
    fft16(vb);
#pragma unroll
    for (int r = 0; r < 16; r++) {
      buf[fpad(ia * 16 + r)] = va[4 * (r & 3) + (r >> 2)];
      buf[fpad(ib * 16 + r)] = vb[4 * (r & 3) + (r >> 2)];
    }
    __syncthreads();
  }
  pass16_smem(buf, tw, 16, 64);
  pass16_smem(buf, tw, 256, 4);
  // ---- forward pass 4 (radix 4, p = 4096) fused with Y = conj(X * H)
#pragma unroll 2
  for (int b = 0; b < 8; b++) {
    const int i = threadIdx.x + b * kFftThreads;
    float2 a0 = buf[fpad(i)], a1 = buf[fpad(i + 4096)], a2 = buf[fpad(i + 8192)], a3 = buf[fpad(i + 12288)];
    const float2 w1 = tw_lookup(tw, i);
    const float2 w2 = cmulf(w1, w1);
    a1 = cmulf(a1, w1);
    a2 = cmulf(a2, w2);
    a3 = cmulf(a3, cmulf(w2, w1));
    fft4(a0, a1, a2, a3);
    buf[fpad(i)] = cconj(cmulf(a0, H[i]));
    buf[fpad(i + 4096)] = cconj(cmulf(a1, H[i + 4096]));
    buf[fpad(i + 8192)] = cconj(cmulf(a2, H[i + 8192]));
    buf[fpad(i + 12288)] = cconj(cmulf(a3, H[i + 12288]));
  }
  __syncthreads();
  // ---- inverse = conj(FFT(conj(.))): pass 1 from shared memory
  {
    float2 va[16], vb[16];
    const int ia = threadIdx.x, ib = threadIdx.x + kFftThreads;
#pragma unroll
    for (int r = 0; r < 16; r++) {
      va[r] = buf[fpad(ia + r * 1024)];
      vb[r] = buf[fpad(ib + r * 1024)];
    }
    __syncthreads();
    fft16(va);
    fft16(vb);
#pragma unroll
    for (int r = 0; r < 16; r++) {
      buf[fpad(ia * 16 + r)] = va[4 * (r & 3) + (r >> 2)];
      buf[fpad(ib * 16 + r)] = vb[4 * (r & 3) + (r >> 2)];
    }
    __syncthreads();
  }
  pass16_smem(buf, tw, 16, 64);
  pass16_smem(buf, tw, 256, 4);
  // ---- inverse pass 4, outputs straight to the global ring (only the alias-free part)
  const int64_t qb = q0 + (int64_t)blk * lq;
#pragma unroll 2
  for (int b = 0; b < 8; b++) {
    const int i = threadIdx.x + b * kFftThreads;
    float2 a0 = buf[fpad(i)], a1 = buf[fpad(i + 4096)], a2 = buf[fpad(i + 8192)], a3 = buf[fpad(i + 12288)];
    const float2 w1 = tw_lookup(tw, i);
    const float2 w2 = cmulf(w1, w1);
    a1 = cmulf(a1, w1);
    a2 = cmulf(a2, w2);
    a3 = cmulf(a3, cmulf(w2, w1));
    fft4(a0, a1, a2, a3);
    const float2 y[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int n = i + r * 4096;        // buffer slot; output u sits at slot u*down + klen - 1
      const int rel = n - (klen - 1);
      if (rel >= 0) {
        const int u = rel / down;
        if (u * down == rel && u < cnt) out.st(c, qb + u, cconj(y[r]));
      }
    }
  }
}

} // namespace fmr
#endif
