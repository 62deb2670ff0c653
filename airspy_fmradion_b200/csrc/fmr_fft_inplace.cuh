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

// ---------------------------------------------------------------------------------------------------------------
// Radix 32 x 32 x 16 form of the same in-place scheme: three passes each way instead of four, i.e. 9 instead of 13
// sweeps of the block through shared memory (the in-place radix-16 kernel sits on the shared-memory pipe).
//   position p = 512 d1 + 16 d2 + d3 holds frequency k = d1 + 32 d2 + 1024 d3 (d1, d2 < 32, d3 < 16)
//   DIF stride 512: butterfly b < 512 takes {b + 512 a}, 32-point DFT, times W_N^(b d), stored at {b + 512 d}
//   DIF stride 16 : chunk c < 32 of 512, b < 16: {512 c + b + 16 a}, 32-point DFT, times W_512^(b d)
//   middle        : 16 contiguous slots: 16-point DFT, times H (digit-reversed), conjugate, 16-point DFT back
//   DIT stride 16 / stride 512: the mirror images (twiddle first).
// One 32-point butterfly per thread and pass (512 threads), two 16-point ones in the middle.
namespace ipfft32 {
using ipfft::cadd;
using ipfft::cconj;
using ipfft::cmul;
using ipfft::csub;
using ipfft::fft16;
using ipfft::mk;
using ipfft::nat;
using ipfft::pad;
constexpr int kN = 16384;
constexpr int kBufLen = kN + kN / 16;
constexpr int kTabLen = 256; // two-level table of W_N: [q] q < 128: W^(128 q), [128 + l]: W^l

FMR_IP_HD float2 tw_lookup(const float2 *tab, int m) { return cmul(tab[m >> 7], tab[128 + (m & 127)]); }

// 32-point forward DFT. In: e[n] = x[2 n], o[n] = x[2 n + 1]. Out: X[k] in e[nat(k)], X[k + 16] in o[nat(k)], k < 16.
FMR_IP_HD void fft32(float2 (&e)[16], float2 (&o)[16]) {
  fft16(e);
  fft16(o);
  // W_32^k, k = 0 .. 15
  const float c[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                       0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                       0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                       -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
  const float sn[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                        0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                        1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                        0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int r = nat(k);
    float2 t;
    if (k == 0) {
      t = o[r];
    } else if (k == 8) {
      t = mk(o[r].y, -o[r].x); // times -j
    } else {
      t = cmul(o[r], mk(c[k], -sn[k]));
    }
    const float2 a = e[r];
    e[r] = cadd(a, t);
    o[r] = csub(a, t);
  }
}

// powers of a root: w[j] = w1^j (j < 8), w8, w16, w24
struct Pow32 {
  float2 w[8], w8, w16, w24;
};
// m: exponent of the root (W_N^m); the three coarse powers come from the table instead of repeated squaring, which
// would multiply the rounding error of the root by 8, 16 and 24 (24 m < N for every caller)
FMR_IP_HD void powers32(const float2 *tab, int m, Pow32 &P) {
  const float2 w1 = tw_lookup(tab, m);
  P.w[0] = mk(1.0f, 0.0f);
  P.w[1] = w1;
  P.w[2] = cmul(w1, w1);
  P.w[3] = cmul(P.w[2], w1);
  P.w[4] = cmul(P.w[2], P.w[2]);
  P.w[5] = cmul(P.w[4], w1);
  P.w[6] = cmul(P.w[3], P.w[3]);
  P.w[7] = cmul(P.w[4], P.w[3]);
  P.w8 = tw_lookup(tab, 8 * m);
  P.w16 = tw_lookup(tab, 16 * m);
  P.w24 = tw_lookup(tab, 24 * m);
}
FMR_IP_HD float2 pow_of(const Pow32 &P, int d) { // w1^d, d = 1 .. 31 (compile-time d after unrolling)
  const int j = d & 7, g = d >> 3;
  if (g == 0) return P.w[j];
  const float2 wg = (g == 1) ? P.w8 : (g == 2) ? P.w16 : P.w24;
  return (j == 0) ? wg : cmul(wg, P.w[j]);
}

// ---- DIF, stride 512: butterfly b in [0, 512); inputs x[b + 512 a] from `ld`
template <typename LD> FMR_IP_HD void dif_first(int b, LD ld, float2 *buf, const float2 *tab) {
  float2 e[16], o[16];
#pragma unroll
  for (int n = 0; n < 16; n++) {
    e[n] = ld(b + 512 * (2 * n));
    o[n] = ld(b + 512 * (2 * n + 1));
  }
  fft32(e, o);
  Pow32 P;
  powers32(tab, b, P);
  float2 *dst = buf + pad(b); // pad(b + 512 d) = pad(b) + 544 d
  dst[0] = e[nat(0)];
#pragma unroll
  for (int d = 1; d < 32; d++) dst[544 * d] = cmul((d < 16) ? e[nat(d)] : o[nat(d - 16)], pow_of(P, d));
}
// ---- DIF, stride 16: i in [0, 512): chunk c = i >> 4 of 512, b = i & 15
FMR_IP_HD void dif_16(int i, float2 *buf, const float2 *tab) {
  const int b = i & 15, c = i >> 4;
  float2 *p = buf + 544 * c + b; // pad(512 c + b + 16 a) = 544 c + b + 17 a
  float2 e[16], o[16];
#pragma unroll
  for (int n = 0; n < 16; n++) {
    e[n] = p[17 * (2 * n)];
    o[n] = p[17 * (2 * n + 1)];
  }
  fft32(e, o);
  Pow32 P;
  powers32(tab, 32 * b, P); // W_512^b = W_N^(32 b)
  p[0] = e[nat(0)];
#pragma unroll
  for (int d = 1; d < 32; d++) p[17 * d] = cmul((d < 16) ? e[nat(d)] : o[nat(d - 16)], pow_of(P, d));
}
// ---- middle: 16 contiguous slots, i in [0, 1024)
FMR_IP_HD void mid_r16(int i, float2 *buf, const float2 *__restrict__ hrev) {
  float2 *p = buf + 17 * i; // pad(16 i + r) = 17 i + r
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; r++) v[r] = p[r];
  fft16(v);
  const float4 *__restrict__ h4 = reinterpret_cast<const float4 *>(hrev) + 8 * i;
  float2 u[16];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 h = h4[q];
    u[2 * q] = cconj(cmul(v[nat(2 * q)], mk(h.x, h.y)));
    u[2 * q + 1] = cconj(cmul(v[nat(2 * q + 1)], mk(h.z, h.w)));
  }
  fft16(u);
#pragma unroll
  for (int a = 0; a < 16; a++) p[a] = u[nat(a)];
}
// ---- DIT, stride 16
FMR_IP_HD void dit_16(int i, float2 *buf, const float2 *tab) {
  const int b = i & 15, c = i >> 4;
  float2 *p = buf + 544 * c + b;
  Pow32 P;
  powers32(tab, 32 * b, P);
  float2 e[16], o[16];
  e[0] = p[0];
  o[0] = cmul(p[17], pow_of(P, 1));
#pragma unroll
  for (int n = 1; n < 16; n++) {
    e[n] = cmul(p[17 * (2 * n)], pow_of(P, 2 * n));
    o[n] = cmul(p[17 * (2 * n + 1)], pow_of(P, 2 * n + 1));
  }
  fft32(e, o);
#pragma unroll
  for (int a = 0; a < 32; a++) p[17 * a] = (a < 16) ? e[nat(a)] : o[nat(a - 16)];
}
// ---- DIT, stride 512, into registers: y[a] = filtered sample of buffer slot b + 512 a
FMR_IP_HD void dit_last(int b, const float2 *buf, const float2 *tab, float2 (&y)[32]) {
  const float2 *p = buf + pad(b);
  Pow32 P;
  powers32(tab, b, P);
  float2 e[16], o[16];
  e[0] = p[0];
  o[0] = cmul(p[544], pow_of(P, 1));
#pragma unroll
  for (int n = 1; n < 16; n++) {
    e[n] = cmul(p[544 * (2 * n)], pow_of(P, 2 * n));
    o[n] = cmul(p[544 * (2 * n + 1)], pow_of(P, 2 * n + 1));
  }
  fft32(e, o);
#pragma unroll
  for (int a = 0; a < 32; a++) y[a] = cconj((a < 16) ? e[nat(a)] : o[nat(a - 16)]);
}
FMR_IP_HD int freq_of_pos(int p) { return (p >> 9) + 32 * ((p >> 4) & 31) + 1024 * (p & 15); }

} // namespace ipfft32

// ---------------------------------------------------------------------------------------------------------------
// 8192-point form, radix 32 x 16 x 16 (the remainder block of the 10 MHz chain, the 1 MHz-class chains whose filters
// fit an 8192-point block, and the block size the fused persistent front end of DESIGN.md section 10 needs).
//   position p = 256 d1 + 16 d2 + d3 holds frequency k = d1 + 32 d2 + 512 d3 (d1 < 32, d2, d3 < 16)
// 256 threads: one 32-point butterfly per thread in the outer passes, two 16-point ones in the inner passes.
namespace ipfft8k {
using ipfft::cconj;
using ipfft::cmul;
using ipfft::fft16;
using ipfft::mk;
using ipfft::nat;
using ipfft::pad;
using ipfft::powers16;
using ipfft32::fft32;
using ipfft32::mid_r16; // 16 contiguous slots: identical (pad(16 i + r) = 17 i + r), i in [0, 512)
using ipfft32::Pow32;
using ipfft32::pow_of;
constexpr int kN = 8192;
constexpr int kBufLen = kN + kN / 16;
constexpr int kTabLen = 256; // [q] q < 64: W_N^(128 q), [128 + l]: W_N^l

// the two-level table has the same layout as ipfft32's (its contents are W_8192), so its lookup and power helpers serve
using ipfft32::powers32; // 24 m < N for m < 256
using ipfft32::tw_lookup;
// ---- DIF, stride 256, radix 32: butterfly b in [0, 256)
template <typename LD> FMR_IP_HD void dif_first(int b, LD ld, float2 *buf, const float2 *tab) {
  float2 e[16], o[16];
#pragma unroll
  for (int n = 0; n < 16; n++) {
    e[n] = ld(b + 256 * (2 * n));
    o[n] = ld(b + 256 * (2 * n + 1));
  }
  fft32(e, o);
  Pow32 P;
  powers32(tab, b, P);
  float2 *dst = buf + pad(b); // pad(b + 256 d) = pad(b) + 272 d
  dst[0] = e[nat(0)];
#pragma unroll
  for (int d = 1; d < 32; d++) dst[272 * d] = cmul((d < 16) ? e[nat(d)] : o[nat(d - 16)], pow_of(P, d));
}
// ---- DIF, stride 16, radix 16: i in [0, 512): chunk c = i >> 4 of 256, b = i & 15
FMR_IP_HD void dif_16(int i, float2 *buf, const float2 *tab) {
  const int b = i & 15, c = i >> 4;
  float2 *p = buf + 272 * c + b; // pad(256 c + b + 16 a) = 272 c + b + 17 a
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; r++) v[r] = p[17 * r];
  fft16(v);
  float2 w[16];
  powers16(tw_lookup(tab, 32 * b), w); // W_256^b = W_N^(32 b)
  p[0] = v[nat(0)];
#pragma unroll
  for (int d = 1; d < 16; d++) p[17 * d] = cmul(v[nat(d)], w[d]);
}
// ---- DIT, stride 16, radix 16
FMR_IP_HD void dit_16(int i, float2 *buf, const float2 *tab) {
  const int b = i & 15, c = i >> 4;
  float2 *p = buf + 272 * c + b;
  float2 w[16];
  powers16(tw_lookup(tab, 32 * b), w);
  float2 v[16];
  v[0] = p[0];
#pragma unroll
  for (int d = 1; d < 16; d++) v[d] = cmul(p[17 * d], w[d]);
  fft16(v);
#pragma unroll
  for (int a = 0; a < 16; a++) p[17 * a] = v[nat(a)];
}
// ---- DIT, stride 256, radix 32, into registers: y[a] = filtered sample of buffer slot b + 256 a
FMR_IP_HD void dit_last(int b, const float2 *buf, const float2 *tab, float2 (&y)[32]) {
  const float2 *p = buf + pad(b);
  Pow32 P;
  powers32(tab, b, P);
  float2 e[16], o[16];
  e[0] = p[0];
  o[0] = cmul(p[272], pow_of(P, 1));
#pragma unroll
  for (int n = 1; n < 16; n++) {
    e[n] = cmul(p[272 * (2 * n)], pow_of(P, 2 * n));
    o[n] = cmul(p[272 * (2 * n + 1)], pow_of(P, 2 * n + 1));
  }
  fft32(e, o);
#pragma unroll
  for (int a = 0; a < 32; a++) y[a] = cconj((a < 16) ? e[nat(a)] : o[nat(a - 16)]);
}
FMR_IP_HD int freq_of_pos(int p) { return (p >> 8) + 32 * ((p >> 4) & 15) + 512 * (p & 15); }

} // namespace ipfft8k
} // namespace fmr

namespace fmr {
// Polyphase epilogue with the bank and the per-row window offsets in shared memory. The source-level profile of
// k_fir_fft_ip (profiles/README.md) attributes a third of the kernel's stall samples to fi_epilogue: per bank row a
// warp does an integer division (MUFU.RCP + fix-up) and eighteen LDG.CONSTANT loads of the row's coefficients, which
// miss the small L1 that is left beside 150 KB of shared memory, and only then starts its FFMA chain. Here the bank
// (rows padded to kEpiRow floats) is copied to shared memory once per block and (window offset,
// bank row) of every output phase are tabulated once per block, so a row costs one LDS.64 + flen/2 broadcast LDS.64
// before the same eighteen LDS.64 + FFMA pairs in the same order (results are bit-identical to fi_epilogue).
constexpr int kEpiRow = 24;       // floats per bank row in shared memory (covers flen 18 and 24; rows 16-byte aligned)
constexpr int kEpiMaxRows = 192;  // output phases (outstep) the shared-memory tables hold
// `tid` is the thread index and st(i, value) stores interpolator output i of the block, so that the host test can run
// the identical mapping thread by thread.
template <int FLEN, int NT, typename ST>
FMR_IP_HD void fi_epilogue_smem(int tid, const float2 *__restrict__ buf, const float *__restrict__ sbank,
                                const int2 *__restrict__ srow, int instep, int outstep, int klen, int cnt, ST st) {
  static_assert(FLEN <= kEpiRow && (FLEN % 2) == 0, "bank row layout");
  const int lane = tid & 31, warp = tid >> 5, nwarps = NT / 32;
  const int nq = (cnt + outstep - 1) / outstep;
  const int L = nq < 32 ? nq : 32;
  const int G = 32 / L;
  const int sgrp = lane / L, ql = lane - sgrp * L;
  if (sgrp >= G) return;
  // two bank rows per iteration: their loads and FFMA chains are independent, so one hides the other's latency
  const int pstep = nwarps * G;
  for (int pg = warp * G; pg < outstep; pg += 2 * pstep) {
    const int p0 = pg + sgrp, p1 = p0 + pstep;
    const bool v0 = p0 < outstep, v1 = p1 < outstep;
    const int2 r0 = srow[v0 ? p0 : 0], r1 = srow[v1 ? p1 : 0]; // x: window offset dp, y: bank row ph
    const float *__restrict__ row0 = sbank + r0.y * kEpiRow;
    const float *__restrict__ row1 = sbank + r1.y * kEpiRow;
    float h0[FLEN], h1[FLEN];
#pragma unroll
    for (int k = 0; k < FLEN; k += 2) {
      const float2 a = *reinterpret_cast<const float2 *>(row0 + k);
      const float2 b = *reinterpret_cast<const float2 *>(row1 + k);
      h0[k] = a.x;
      h0[k + 1] = a.y;
      h1[k] = b.x;
      h1[k + 1] = b.y;
    }
    for (int q = ql; q < nq; q += L) {
      const int i0 = p0 + outstep * q, i1 = p1 + outstep * q;
      const bool w0 = v0 && i0 < cnt, w1 = v1 && i1 < cnt;
      // windows of outputs that do not exist are read from the first row's window (always inside the block)
      const float2 *__restrict__ x0 = buf + (klen - 1) + r0.x + instep * (w0 ? q : 0);
      const float2 *__restrict__ x1 = buf + (klen - 1) + r1.x + instep * (w1 ? q : 0);
      float2 acc0 = ipfft::mk(0.f, 0.f), acc1 = ipfft::mk(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < FLEN; k++) {
        const float2 a = x0[k], b = x1[k];
        acc0.x += h0[k] * a.x;
        acc0.y += h0[k] * a.y;
        acc1.x += h1[k] * b.x;
        acc1.y += h1[k] * b.y;
      }
      if (w0) st(i0, acc0);
      if (w1) st(i1, acc1);
    }
  }
}

// (window offset, bank row) of output phase p of a block whose first output has remainder rem_b
FMR_IP_HD int2 epi_row(int p, int instep, int outstep, int rem_b) {
  const int pp = p * instep + rem_b;
  const int dp = pp / outstep;
  int2 r;
  r.x = dp;
  r.y = pp - dp * outstep;
  return r;
}

} // namespace fmr

#if defined(__CUDACC__) && defined(FMR_FFT_CUH)
namespace fmr {

// Fused form only (IF chain): block -> filtered block in shared memory (plain order) -> polyphase bank.
// Parameters as k_fir_fft<float, 16384, true>; H is the digit-reversed spectrum, fz.twtab the ipfft table.
// BOUND > THREADS compiles for a nominal larger block, i.e. caps the registers (65536 / BOUND): <512, 896> = 72
// registers, which leaves room for one half-band stream CTA or fused-core CTAs of ANOTHER handle on the same SM
// (bench.py --handles: independent handles on their own streams overlap their HBM-, shared-memory- and latency-bound
// kernels).
// EPI = 1: fi_epilogue_smem (needs kIpEpiSmemBytes of dynamic shared memory; flen 18 or 24 and outstep <= kEpiMaxRows).
template <int THREADS, int BOUND = THREADS, int EPI = 0>
__global__ void __launch_bounds__(BOUND, 1)
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
  float *sbank = reinterpret_cast<float *>(tab + kTabLen);
  int2 *srow = reinterpret_cast<int2 *>(sbank + kEpiMaxRows * kEpiRow);
  if (EPI) {
    const float *__restrict__ gb = reinterpret_cast<const float *>(fz.bank);
    const int flen = fz.flen;
    for (int i = threadIdx.x; i < fz.outstep * flen; i += THREADS) {
      const int rr = i / flen;
      sbank[rr * kEpiRow + (i - rr * flen)] = __ldg(gb + i);
    }
    const int rem0 = (int)((mb * fz.instep) % fz.outstep);
    for (int pz = threadIdx.x; pz < fz.outstep; pz += THREADS) srow[pz] = epi_row(pz, fz.instep, fz.outstep, rem0);
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
  if (EPI && fz.flen == 18) {
    fi_epilogue_smem<18, THREADS>(threadIdx.x, buf, sbank, srow, fz.instep, fz.outstep, klen, cnt,
                                  [&](int i, float2 v) { out.st(c, mb + i, v); });
  } else if (EPI && fz.flen == 24) {
    fi_epilogue_smem<24, THREADS>(threadIdx.x, buf, sbank, srow, fz.instep, fz.outstep, klen, cnt,
                                  [&](int i, float2 v) { out.st(c, mb + i, v); });
  } else if (fz.flen == 18) {
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

// Radix 32 x 32 x 16 form (ipfft32), 512 threads = one 32-point butterfly per thread and pass. Same parameters; Hrev
// in ipfft32's digit-reversed order, fz.twtab = its 256-entry table.
template <int EPI>
__global__ void __launch_bounds__(512, 1)
    k_fir_fft_ip32(Ring<float2> in, Ring<float2> out, const float2 *__restrict__ Hrev, int klen, int64_t n_in_avail, FftFuse fz) {
  using namespace ipfft32;
  constexpr int THREADS = 512;
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
  const int64_t base = qb - fl2;
  if (threadIdx.x < kTabLen) tab[threadIdx.x] = __ldg(reinterpret_cast<const float2 *>(fz.twtab) + threadIdx.x);
  float *sbank = reinterpret_cast<float *>(tab + kTabLen);
  int2 *srow = reinterpret_cast<int2 *>(sbank + kEpiMaxRows * kEpiRow);
  if (EPI) {
    const float *__restrict__ gb = reinterpret_cast<const float *>(fz.bank);
    const int flen = fz.flen;
    for (int i = threadIdx.x; i < fz.outstep * flen; i += THREADS) {
      const int rr = i / flen;
      sbank[rr * kEpiRow + (i - rr * flen)] = __ldg(gb + i);
    }
    const int rem0 = (int)((mb * fz.instep) % fz.outstep);
    for (int pz = threadIdx.x; pz < fz.outstep; pz += THREADS) srow[pz] = epi_row(pz, fz.instep, fz.outstep, rem0);
  }
  __syncthreads();
  {
    const uint32_t pos0 = (uint32_t)base & (in.cap - 1);
    if (base >= 0 && base + kN <= n_in_avail && pos0 + (uint32_t)kN <= in.cap) {
      const float2 *__restrict__ row = in.base + (size_t)c * in.cap + pos0;
      dif_first(threadIdx.x, [&](int n) { return row[n]; }, buf, tab);
    } else {
      dif_first(threadIdx.x,
                [&](int n) {
                  const int64_t t = base + n;
                  return (t < n_in_avail) ? in.ld(c, t) : make_float2(0.f, 0.f);
                },
                buf, tab);
    }
  }
  __syncthreads();
  dif_16(threadIdx.x, buf, tab);
  __syncthreads();
  mid_r16(threadIdx.x, buf, Hrev);
  mid_r16(threadIdx.x + THREADS, buf, Hrev);
  __syncthreads();
  dit_16(threadIdx.x, buf, tab);
  __syncthreads();
  float2 y[32];
  dit_last(threadIdx.x, buf, tab, y);
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 32; a++) {
    const int n = threadIdx.x + 512 * a;
    const int64_t t = qb + (n - (klen - 1));
    const float2 v = (t >= 0) ? y[a] : make_float2(0.f, 0.f);
    if (t >= fz.tail_lo && t < fz.tail_hi && n >= klen - 1 && blk == (int)gridDim.x - 1) {
      Ring<float2>{reinterpret_cast<float2 *>(fz.tail_base), fz.tail_cap}.st(c, t, v);
    }
    buf[n] = v;
  }
  __syncthreads();
  const float *__restrict__ bank = reinterpret_cast<const float *>(fz.bank);
  const int rem_b = (int)((mb * fz.instep) % fz.outstep);
  if (EPI && fz.flen == 18) {
    fi_epilogue_smem<18, THREADS>(threadIdx.x, buf, sbank, srow, fz.instep, fz.outstep, klen, cnt,
                                  [&](int i, float2 v) { out.st(c, mb + i, v); });
  } else if (EPI && fz.flen == 24) {
    fi_epilogue_smem<24, THREADS>(threadIdx.x, buf, sbank, srow, fz.instep, fz.outstep, klen, cnt,
                                  [&](int i, float2 v) { out.st(c, mb + i, v); });
  } else if (fz.flen == 18) {
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

// 8192-point form (ipfft8k), 256 threads, two CTAs per SM. Fused epilogue as above.
static __global__ void __launch_bounds__(256, 2)
    k_fir_fft_ip8k(Ring<float2> in, Ring<float2> out, const float2 *__restrict__ Hrev, int klen, int64_t n_in_avail, FftFuse fz) {
  using namespace ipfft8k;
  constexpr int THREADS = 256;
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
  const int64_t base = qb - fl2;
  tab[threadIdx.x] = __ldg(reinterpret_cast<const float2 *>(fz.twtab) + threadIdx.x); // kTabLen == THREADS
  __syncthreads();
  {
    const uint32_t pos0 = (uint32_t)base & (in.cap - 1);
    if (base >= 0 && base + kN <= n_in_avail && pos0 + (uint32_t)kN <= in.cap) {
      const float2 *__restrict__ row = in.base + (size_t)c * in.cap + pos0;
      dif_first(threadIdx.x, [&](int n) { return row[n]; }, buf, tab);
    } else {
      dif_first(threadIdx.x,
                [&](int n) {
                  const int64_t t = base + n;
                  return (t < n_in_avail) ? in.ld(c, t) : make_float2(0.f, 0.f);
                },
                buf, tab);
    }
  }
  __syncthreads();
  dif_16(threadIdx.x, buf, tab);
  dif_16(threadIdx.x + THREADS, buf, tab);
  __syncthreads();
  mid_r16(threadIdx.x, buf, Hrev);
  mid_r16(threadIdx.x + THREADS, buf, Hrev);
  __syncthreads();
  dit_16(threadIdx.x, buf, tab);
  dit_16(threadIdx.x + THREADS, buf, tab);
  __syncthreads();
  float2 y[32];
  dit_last(threadIdx.x, buf, tab, y);
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 32; a++) {
    const int n = threadIdx.x + 256 * a;
    const int64_t t = qb + (n - (klen - 1));
    const float2 v = (t >= 0) ? y[a] : make_float2(0.f, 0.f);
    if (t >= fz.tail_lo && t < fz.tail_hi && n >= klen - 1 && blk == (int)gridDim.x - 1) {
      Ring<float2>{reinterpret_cast<float2 *>(fz.tail_base), fz.tail_cap}.st(c, t, v);
    }
    buf[n] = v;
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

constexpr int kIpEpiSmemBytes = (ipfft::kBufLen + ipfft::kTabLen) * (int)sizeof(float2) + kEpiMaxRows * kEpiRow * (int)sizeof(float) +
                                kEpiMaxRows * (int)sizeof(int2);
constexpr int kIp8kSmemBytes = (ipfft8k::kBufLen + ipfft8k::kTabLen) * (int)sizeof(float2);
constexpr int kIpSmemBytes = (ipfft::kBufLen + ipfft::kTabLen) * (int)sizeof(float2);
constexpr int kIp32SmemBytes = (ipfft32::kBufLen + ipfft32::kTabLen) * (int)sizeof(float2);
constexpr int kIp32EpiSmemBytes = kIp32SmemBytes + kEpiMaxRows * kEpiRow * (int)sizeof(float) + kEpiMaxRows * (int)sizeof(int2);

} // namespace fmr
#endif
#endif
