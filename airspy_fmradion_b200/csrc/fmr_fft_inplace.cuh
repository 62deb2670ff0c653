// fmr_fft_inplace.cuh — the 16384-point FP32 overlap-save low-pass (+ fused polyphase bank) with TRUE in-place
// passes: decimation in frequency forward (natural order in, digit-reversed out), the filter spectrum stored in
// the same digit-reversed order, decimation in time back (digit-reversed in, natural out). Same linear convolution
// as k_fir_fft (fmr_fft.cuh; reference r8b::CDSPBlockConvolver::process, CDSPBlockConvolver.h:252-353).
//
// Why: the Stockham passes of k_fir_fft write to other slots than they read, so every pass needs a barrier between
// "all loads done" and "first store" as well as one at its end; the SM alternates between a shared-memory phase and a
// math phase (ncu: 41 % shared-memory wavefront utilisation + 47 % issue utilisation, back to back, insensitive to
// -10 % instructions and to 2x warps). Here every butterfly stores exactly where it loaded, so a pass needs ONE
// barrier, the warps of a pass drift apart (some load while others compute), the forward radix-4 pass, the
// multiplication by H and the first inverse pass fuse in registers, and no reordering pass exists at all:
// 8 barriers and 5.5 shared-memory round trips per block instead of 15 and 7.
//
// Index algebra (N = 16 * 16 * 16 * 4, position p = 1024 d1 + 64 d2 + 4 d3 + d4 holds frequency
// k = d1 + 16 d2 + 256 d3 + 4096 d4):
//   DIF pass, stride s in {1024, 64, 4}: butterfly b of a chunk of 16 s takes {b + s a}, a 16-point DFT over a -> d,
//     times W_(16 s)^(b d), stored at {b + s d}; then radix 4 on {4 i + a}.
//   DIT pass, stride s in {4, 64, 1024}: {b + s d} times W_(16 s)^(b d), 16-point DFT over d -> a, stored at {b + s a}.
// Shared-memory layout: slot n at n + (n >> 4) (one pad word per 16): every pass is bank-conflict free per half warp.
//
// The per-thread bodies of the passes are __host__ __device__ so that tests/cpp/fft_inplace_host_test.cu can run the
// identical index algebra on the CPU, thread by thread and pass by pass, against a direct convolution.
#ifndef FMR_FFT_INPLACE_CUH
#define FMR_FFT_INPLACE_CUH

#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define FMR_IP_HD __host__ __device__ __forceinline__
#else
#define FMR_IP_HD inline
#endif

namespace fmr {
namespace ipfft {

constexpr int kN = 16384;
constexpr int kBufLen = kN + kN / 16;          // padded block
constexpr int kTw = 0, kT64 = 256, kT4 = 256 + 1024, kTabLen = 256 + 1024 + 64;
// table layout (float2): [kTw + q] q < 128: W_N^(128 q), q >= 128: W_N^(q - 128);
//                        [kT64 + d * 64 + b] = W_1024^(b d);  [kT4 + d * 4 + b] = W_64^(b d)

FMR_IP_HD int pad(int n) { return n + (n >> 4); }
FMR_IP_HD float2 mk(float a, float b) {
  float2 v;
  v.x = a;
  v.y = b;
  return v;
}
FMR_IP_HD float2 cmul(float2 a, float2 b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FMR_IP_HD float2 cconj(float2 a) { return mk(a.x, -a.y); }
FMR_IP_HD float2 cadd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __fadd2_rn(a, b);
#else
  return mk(a.x + b.x, a.y + b.y);
#endif
}
FMR_IP_HD float2 csub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
#else
  return mk(a.x - b.x, a.y - b.y);
#endif
}
// 4-point forward DFT, natural order in and out
FMR_IP_HD void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  const float2 t = mk(d13.y, -d13.x); // -j d13
  a1 = cadd(d02, t);
  a3 = csub(d02, t);
}
// 16-point forward DFT in registers; X[m] is left in v[4 * (m & 3) + (m >> 2)]
FMR_IP_HD void fft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) fft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, r2 = 0.70710678118654752440f;
  v[1 + 4] = cmul(v[1 + 4], mk(c1, -s1));
  v[2 + 4] = cmul(v[2 + 4], mk(r2, -r2));
  v[3 + 4] = cmul(v[3 + 4], mk(s1, -c1));
  v[1 + 8] = cmul(v[1 + 8], mk(r2, -r2));
  v[2 + 8] = mk(v[2 + 8].y, -v[2 + 8].x);
  v[3 + 8] = cmul(v[3 + 8], mk(-r2, -r2));
  v[1 + 12] = cmul(v[1 + 12], mk(s1, -c1));
  v[2 + 12] = cmul(v[2 + 12], mk(-r2, -r2));
  v[3 + 12] = cmul(v[3 + 12], mk(-c1, s1));
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) fft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}
FMR_IP_HD int nat(int m) { return 4 * (m & 3) + (m >> 2); } // register that holds X[m] after fft16
FMR_IP_HD float2 tw_lookup(const float2 *tab, int m) { return cmul(tab[kTw + (m >> 7)], tab[kTw + 128 + (m & 127)]); }
// powers w^1 .. w^15
FMR_IP_HD void powers16(float2 w1, float2 (&w)[16]) {
  w[0] = mk(1.0f, 0.0f);
  w[1] = w1;
  w[2] = cmul(w1, w1);
  w[3] = cmul(w[2], w1);
  w[4] = cmul(w[2], w[2]);
  w[5] = cmul(w[4], w1);
  w[6] = cmul(w[3], w[3]);
  w[7] = cmul(w[4], w[3]);
  w[8] = cmul(w[4], w[4]);
#pragma unroll
  for (int r = 9; r < 16; r++) w[r] = cmul(w[8], w[r - 8]);
}

// ---- DIF, stride 1024: butterfly b in [0, 1024); inputs x[b + 1024 a] come from `ld` (the global ring)
template <typename LD> FMR_IP_HD void dif_first(int b, LD ld, float2 *buf, const float2 *tab) {
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; r++) v[r] = ld(b + 1024 * r);
  fft16(v);
  float2 w[16];
  powers16(tw_lookup(tab, b), w);
  float2 *dst = buf + pad(b);
  dst[0] = v[nat(0)];
#pragma unroll
  for (int d = 1; d < 16; d++) dst[d * 1088] = cmul(v[nat(d)], w[d]); // pad(b + 1024 d) = pad(b) + 1088 d
}
// ---- DIF, stride 64: i in [0, 1024): chunk j = i >> 6 of 1024, b = i & 63
FMR_IP_HD void dif_64(int i, float2 *buf, const float2 *tab) {
  const int b = i & 63, e0 = 1024 * (i >> 6) + b;
  float2 *p = buf + pad(e0); // pad(e0 + 64 r) = pad(e0) + 68 r
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; r++) v[r] = p[68 * r];
  fft16(v);
  const float2 *t = tab + kT64 + b;
  p[0] = v[nat(0)];
#pragma unroll
  for (int d = 1; d < 16; d++) p[68 * d] = cmul(v[nat(d)], t[64 * d]);
}
// ---- DIF, stride 4: i in [0, 1024): chunk j = i >> 2 of 64, b = i & 3
FMR_IP_HD void dif_4(int i, float2 *buf, const float2 *tab) {
  const int b = i & 3, j = i >> 2;
  float2 *p = buf + 68 * j + b; // pad(64 j + b + 4 r) = 68 j + b + 4 r + (r >> 2)
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; r++) v[r] = p[4 * r + (r >> 2)];
  fft16(v);
  const float2 *t = tab + kT4 + b;
  p[0] = v[nat(0)];
#pragma unroll
  for (int d = 1; d < 16; d++) p[4 * d + (d >> 2)] = cmul(v[nat(d)], t[4 * d]);
}
// ---- radix 4 forward, times H (digit-reversed order, 1/N folded in), conjugate, radix 4 back: i in [0, 4096)
FMR_IP_HD void mid_r4(int i, float2 *buf, const float2 *__restrict__ hrev) {
  float2 *p = buf + 4 * i + (i >> 2); // pad(4 i + r) = 4 i + r + (i >> 2)
  float2 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
  fft4(a0, a1, a2, a3);
  const float4 h01 = reinterpret_cast<const float4 *>(hrev)[2 * i], h23 = reinterpret_cast<const float4 *>(hrev)[2 * i + 1];
  a0 = cconj(cmul(a0, mk(h01.x, h01.y)));
  a1 = cconj(cmul(a1, mk(h01.z, h01.w)));
  a2 = cconj(cmul(a2, mk(h23.x, h23.y)));
  a3 = cconj(cmul(a3, mk(h23.z, h23.w)));
  fft4(a0, a1, a2, a3);
  p[0] = a0;
  p[1] = a1;
  p[2] = a2;
  p[3] = a3;
}
// ---- DIT, stride 4
FMR_IP_HD void dit_4(int i, float2 *buf, const float2 *tab) {
  const int b = i & 3, j = i >> 2;
  float2 *p = buf + 68 * j + b;
  const float2 *t = tab + kT4 + b;
  float2 v[16];
  v[0] = p[0];
#pragma unroll
  for (int d = 1; d < 16; d++) v[d] = cmul(p[4 * d + (d >> 2)], t[4 * d]);
  fft16(v);
#pragma unroll
  for (int a = 0; a < 16; a++) p[4 * a + (a >> 2)] = v[nat(a)];
}
// ---- DIT, stride 64
FMR_IP_HD void dit_64(int i, float2 *buf, const float2 *tab) {
  const int b = i & 63, e0 = 1024 * (i >> 6) + b;
  float2 *p = buf + pad(e0);
  const float2 *t = tab + kT64 + b;
  float2 v[16];
  v[0] = p[0];
#pragma unroll
  for (int d = 1; d < 16; d++) v[d] = cmul(p[68 * d], t[64 * d]);
  fft16(v);
#pragma unroll
  for (int a = 0; a < 16; a++) p[68 * a] = v[nat(a)];
}
// ---- DIT, stride 1024, into registers: y[a] = conj(z[b + 1024 a]) = the filtered sample of buffer slot b + 1024 a
FMR_IP_HD void dit_last(int b, const float2 *buf, const float2 *tab, float2 (&y)[16]) {
  const float2 *p = buf + pad(b);
  float2 w[16];
  powers16(tw_lookup(tab, b), w);
  float2 v[16];
  v[0] = p[0];
#pragma unroll
  for (int d = 1; d < 16; d++) v[d] = cmul(p[1088 * d], w[d]);
  fft16(v);
#pragma unroll
  for (int a = 0; a < 16; a++) y[a] = cconj(v[nat(a)]);
}

// position p of the digit-reversed spectrum holds frequency k
FMR_IP_HD int freq_of_pos(int p) { return (p >> 10) + 16 * ((p >> 6) & 15) + 256 * ((p >> 2) & 15) + 4096 * (p & 3); }

} // namespace ipfft

} // namespace fmr

#if defined(__CUDACC__) && defined(FMR_FFT_CUH)
namespace fmr {

// Fused form only (IF chain): block -> filtered block in shared memory (plain order) -> polyphase bank.
// Parameters as k_fir_fft<float, 16384, true>; H is the digit-reversed spectrum, fz.twtab the ipfft table.
// (Variants measured and dropped in round 2, profiles/sweep_fft_variants_r02.jsonl: radix 32 x 32 x 16 passes, the
// polyphase bank in shared memory, an 8192-point in-place form — none faster than this one.)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
    k_fir_fft_ip(Ring<float2> in, Ring<float2> out, const float2 *__restrict__ Hrev, int klen, int64_t n_in_avail, FftFuse fz) {
  using namespace ipfft;
  constexpr int SETS = 1024 / THREADS, NB4 = 4096 / THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *buf = reinterpret_cast<float2 *>(smem_raw);
  float2 *tab = buf + kBufLen;
  const uint32_t c = blockIdx.y;
  const int blk = blockIdx.x;
  int cnt = fz.n_m - blk * fz.mo;
  if (cnt > fz.mo) cnt = fz.mo;
  if (cnt <= 0) return;
  const int64_t mb = fz.m0 + (int64_t)blk * fz.mo;
  const int64_t qb = (mb * fz.instep) / fz.outstep - (fz.flen / 2 - 1);
  const int fl2 = (klen - 1) / 2;
  const int64_t base = qb - fl2; // input sample index held by buffer slot 0
  {
    const float2 *__restrict__ g = reinterpret_cast<const float2 *>(fz.twtab);
    for (int i = threadIdx.x; i < kTabLen; i += THREADS) tab[i] = __ldg(g + i);
  }
  __syncthreads();
  // ---- forward, decimation in frequency
  {
    const uint32_t pos0 = (uint32_t)base & (in.cap - 1);
    if (base >= 0 && base + kN <= n_in_avail && pos0 + (uint32_t)kN <= in.cap) {
      const float2 *__restrict__ row = in.base + (size_t)c * in.cap + pos0;
#pragma unroll
      for (int s = 0; s < SETS; s++) dif_first(threadIdx.x + s * THREADS, [&](int n) { return row[n]; }, buf, tab);
    } else {
#pragma unroll
      for (int s = 0; s < SETS; s++) {
        dif_first(threadIdx.x + s * THREADS,
                  [&](int n) {
                    const int64_t t = base + n;
                    return (t < n_in_avail) ? in.ld(c, t) : make_float2(0.f, 0.f);
                  },
                  buf, tab);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < SETS; s++) dif_64(threadIdx.x + s * THREADS, buf, tab);
  __syncthreads();
#pragma unroll
  for (int s = 0; s < SETS; s++) dif_4(threadIdx.x + s * THREADS, buf, tab);
  __syncthreads();
#pragma unroll 2
  for (int b = 0; b < NB4; b++) mid_r4(threadIdx.x + b * THREADS, buf, Hrev);
  __syncthreads();
  // ---- back, decimation in time
#pragma unroll
  for (int s = 0; s < SETS; s++) dit_4(threadIdx.x + s * THREADS, buf, tab);
  __syncthreads();
#pragma unroll
  for (int s = 0; s < SETS; s++) dit_64(threadIdx.x + s * THREADS, buf, tab);
  __syncthreads();
  float2 y[SETS][16];
#pragma unroll
  for (int s = 0; s < SETS; s++) dit_last(threadIdx.x + s * THREADS, buf, tab, y[s]);
  __syncthreads();
  // ---- filtered block back to shared memory in PLAIN order (slot n holds filter output qb + n - (klen-1)); the
  // newest samples of the last block also go to the intermediate ring (see FftFuse)
#pragma unroll
  for (int s = 0; s < SETS; s++) {
#pragma unroll
    for (int a = 0; a < 16; a++) {
      const int n = threadIdx.x + s * THREADS + 1024 * a;
      const int64_t t = qb + (n - (klen - 1));
      const float2 v = (t >= 0) ? y[s][a] : make_float2(0.f, 0.f);
      if (t >= fz.tail_lo && t < fz.tail_hi && n >= klen - 1 && blk == (int)gridDim.x - 1) {
        Ring<float2>{reinterpret_cast<float2 *>(fz.tail_base), fz.tail_cap}.st(c, t, v);
      }
      buf[n] = v;
    }
  }
  __syncthreads();
  const float *__restrict__ bank = reinterpret_cast<const float *>(fz.bank);
  const int rem_b = (int)((mb * fz.instep) % fz.outstep);
  if (fz.flen == 18) {
    fi_epilogue<float2, 18, THREADS>(buf, bank, fz.instep, fz.outstep, klen, rem_b, cnt, out, c, mb);
  } else if (fz.flen == 24) {
    fi_epilogue<float2, 24, THREADS>(buf, bank, fz.instep, fz.outstep, klen, rem_b, cnt, out, c, mb);
  } else {
    for (int i = threadIdx.x; i < cnt; i += THREADS) {
      const int prel = i * fz.instep + rem_b;
      const int dip = prel / fz.outstep;
      const int ph = prel - dip * fz.outstep;
      const float *__restrict__ row = bank + (size_t)ph * fz.flen;
      const int n0 = (klen - 1) + dip;
      float2 acc = make_float2(0.f, 0.f);
      for (int k = 0; k < fz.flen; k++) {
        const float2 x = buf[n0 + k];
        const float h = __ldg(row + k);
        acc.x += h * x.x;
        acc.y += h * x.y;
      }
      out.st(c, mb + i, acc);
    }
  }
}

constexpr int kIpSmemBytes = (ipfft::kBufLen + ipfft::kTabLen) * (int)sizeof(float2);

} // namespace fmr
#endif
#endif
