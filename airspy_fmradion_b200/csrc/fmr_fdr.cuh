// fmr_fdr.cuh — frequency-domain form of the IF resampler's last two stages: the long zero-phase low-pass
// (r8b::CDSPBlockConvolver::process, CDSPBlockConvolver.h:252-353) and the whole-step polyphase bank behind it
// (r8b::CDSPFracInterpolator::process, CDSPFracInterpolator.h:861-925), for the 625:192 rate pairs (10 MHz and
// 2.5 MHz -> 384 kHz, where the stream in front of the low-pass runs at 1.25 MHz).
//
// What the reference computes:  y = h * x  (2307 taps, zero phase),  z[m] = sum_k bank[(625 m) mod 192][k] y[q_m + k],
// q_m = floor(625 m / 192) - 8, i.e. z[m] = y(625 m / 192) evaluated with an 18-tap fractional-delay row. Two facts
// about r8brain's design (checked on the frozen tables by tests/test_fdr_cpu.py) turn this into ONE forward and one
// much smaller inverse FFT with no time-domain epilogue at all:
//   * |H(f)| < 2e-9 for |f| >= 192/1250 (the output Nyquist): y is band-limited to the output band;
//   * every bank row is an ideal fractional delay to 1e-8 for |f| <= 0.16: on such a y the bank IS band-limited
//     interpolation.
// So for a block of N_in = 16 * 625 = 10000 input samples starting at a multiple of 625:
//     X = FFT_10000(x);  Z[k] = X[k] H0[k] / N_in for |k| < 1536 (the other 6928 bins are below 2e-9 and dropped);
//     z = IFFT_3072(Z): sample i of the block sits at input time base + i * 625/192, exactly the reference's grid.
// Overlap-save: the circular wrap of H0 (+-1153) and the 18-tap support of the implied interpolation kernel corrupt
// less than 1250 input samples at either end, so a block yields 12 * 192 = 2304 outputs from 7500 new input samples.
// Against the 16384-point convolution + bank from shared memory (k_fir_fft_ip): 2.1x fewer shared-memory accesses and
// 1.6x fewer floating-point operations per output, and the difference to the reference stays below FP32 rounding.
//
// Transforms (all passes store where they load, one barrier per pass, no reordering pass):
//   forward 10000 = 16 x 25 x 25, decimation in frequency: position 625 k1 + 25 k2 + k3 ends up holding frequency
//     k1 + 16 k2 + 400 k3; the last pass only evaluates the 8 of 25 outputs (k3 in {0..3, 21..24}) inside the band,
//     multiplies by H0 and scatters them in natural order into the 3072-point buffer (conjugated: the inverse is run
//     as a forward transform);
//   inverse 3072 = 16 x 16 x 12, decimation in frequency, buffer padded one word per 192; the last pass writes its
//     outputs straight to the 384 kHz ring (lanes = consecutive output samples).
// The per-thread bodies are __host__ __device__: tests/cpp/fdr_host_test.cu runs them on the CPU against the
// reference's two stages evaluated directly in double.
#ifndef FMR_FDR_CUH
#define FMR_FDR_CUH

#include "fmr_fft_inplace.cuh"

namespace fmr {
namespace fdr {
using ipfft::cadd;
using ipfft::cconj;
using ipfft::cmul;
using ipfft::csub;
using ipfft::fft16;
using ipfft::fft4;
using ipfft::mk;
using ipfft::nat;
using ipfft::powers16;

constexpr int kNin = 10000;            // forward transform, 16 x 25 x 25, for every rate pair
// Geometry of one rate pair: the inverse transform has kNout = 256 * RL points (RL = its last radix):
//   625:192 (10 MHz, 2.5 MHz -> 384 kHz; 1.25 MHz in front of the low-pass):  RL = 12, 3072 points
//   125:48  (1 MHz -> 384 kHz; 1 MHz -> 48 kHz behind three half-bands):      RL = 15, 3840 points
//   125:32  (6 MHz, 3 MHz -> 384 kHz; 1.5 MHz in front of the low-pass):      RL = 10, 2560 points
template <int RL> struct Geo {
  static constexpr int kNout = 256 * RL;
  static constexpr int kChunk = 16 * RL;                     // slots per chunk of the inverse's first pass
  static constexpr int kZLen = kNout + 16;                   // one pad word per chunk
  static constexpr int kHalf = kNout / 2;                    // bins |k| < kHalf are kept (k = -kHalf too)
  static constexpr int kL = (kHalf - 1) / 400;               // band outputs of a last-pass butterfly: k3 in [0, kL] and [24 - kL, 24]
  static constexpr int kKeep = 2 * (kL + 1);
  static constexpr int kTi1 = 625 + 200;                     // [kChunk] W_kNout^b
  static constexpr int kTi2 = kTi1 + kChunk;                 // [RL]     W_kChunk^k3
  static constexpr int kTabLen = kTi2 + RL;
  static constexpr int kHsLen = kKeep * 400;
};
// twiddle table (float2), built in double by the host (fdr_make_tables):
constexpr int kTw1 = 0;                // [625]    W_10000^b
constexpr int kTw2 = kTw1 + 625;       // [25][8]  W_625^(n3 * {1,2,3,4,5,10,15,20})
// spectrum table (float, H0 is real: the low-pass is symmetric): Hs[e * 400 + r] = H0[freq(r, e)] / N_in,
// r = k1 + 16 k2 in [0, 400), e in [0, kKeep) <-> k3 = e (e <= kL) or 25 - kKeep + e; 0 where the bin is outside the band.

constexpr int kMaxAdvOut = 3072;       // most outputs per block over the pairs (125:48 with its guard of 1000)
// The 625:192 pair (what the fused front end, fmr_frontend.cuh, is built for)
using G12 = Geo<12>;
constexpr int kNout = G12::kNout;
constexpr int kInStep = 625, kOutStep = 192;
constexpr int kGuardIn = 1250;         // input samples dropped at either end of a block (2 * 625 >= 1153 + 9)
constexpr int kAdvIn = kNin - 2 * kGuardIn;                  // 7500 new input samples per block
constexpr int kGuardOut = kGuardIn / kInStep * kOutStep;     // 384
constexpr int kAdvOut = kAdvIn / kInStep * kOutStep;         // 2304 outputs per block
constexpr int kZLen = G12::kZLen;
constexpr int kKeep = G12::kKeep;

template <int RL> FMR_IP_HD int zpos(int k) { return k + k / (16 * RL); }
FMR_IP_HD float2 smul(float s, float2 a) { return mk(s * a.x, s * a.y); }
FMR_IP_HD float2 sfma(float s, float2 a, float2 b) { // s * a + b
#ifdef __CUDA_ARCH__
  return __ffma2_rn(make_float2(s, s), a, b);
#else
  return mk(s * a.x + b.x, s * a.y + b.y);
#endif
}
FMR_IP_HD float2 mulmj(float2 a) { return mk(a.y, -a.x); } // -j a
FMR_IP_HD float2 mulpj(float2 a) { return mk(-a.y, a.x); } // +j a

// ---- 5-point forward DFT, natural order in and out
constexpr float kC51 = 0.30901699437494745126f, kC52 = -0.80901699437494734024f;
constexpr float kS51 = 0.95105651629515353118f, kS52 = 0.58778525229247324813f;
FMR_IP_HD void dft5(float2 &x0, float2 &x1, float2 &x2, float2 &x3, float2 &x4) {
  const float2 a1 = cadd(x1, x4), a2 = cadd(x2, x3), b1 = csub(x1, x4), b2 = csub(x2, x3);
  const float2 c1 = sfma(kC52, a2, sfma(kC51, a1, x0));
  const float2 c2 = sfma(kC51, a2, sfma(kC52, a1, x0));
  const float2 s1 = sfma(kS52, b2, smul(kS51, b1));
  const float2 s2 = sfma(-kS51, b2, smul(kS52, b1));
  x0 = cadd(x0, cadd(a1, a2));
  x1 = cadd(c1, mulmj(s1));
  x4 = cadd(c1, mulpj(s1));
  x2 = cadd(c2, mulmj(s2));
  x3 = cadd(c2, mulpj(s2));
}
// outputs 0 and 4 only (x0 <- X0, x4 <- X4)
FMR_IP_HD void dft5_04(float2 &x0, float2 x1, float2 x2, float2 x3, float2 &x4) {
  const float2 a1 = cadd(x1, x4), a2 = cadd(x2, x3), b1 = csub(x1, x4), b2 = csub(x2, x3);
  const float2 c1 = sfma(kC52, a2, sfma(kC51, a1, x0));
  const float2 s1 = sfma(kS52, b2, smul(kS51, b1));
  x0 = cadd(x0, cadd(a1, a2));
  x4 = cadd(c1, mulpj(s1));
}
// W_25^m for the products n2 * k1 that occur inside the 25-point butterfly
FMR_IP_HD float2 w25(int m) {
  switch (m) {
  case 1: return mk(0.96858316112863107605f, -0.24868988716485479484f);
  case 2: return mk(0.87630668004386358394f, -0.48175367410171532345f);
  case 3: return mk(0.72896862742141155245f, -0.68454710592868861507f);
  case 4: return mk(0.53582679497899654564f, -0.84432792550201507531f);
  case 6: return mk(0.062790519529313526537f, -0.99802672842827155897f);
  case 8: return mk(-0.42577929156507271502f, -0.9048270524660194658f);
  case 9: return mk(-0.63742398974868974548f, -0.77051324277578925326f);
  case 12: return mk(-0.99211470131447776488f, -0.12533323356430453588f);
  default: return mk(-0.63742398974868952344f, 0.77051324277578936428f); // 16
  }
}
// first two steps of the 25-point DFT (n = 5 n1 + n2): five 5-point DFTs over n1 and the inner twiddles;
// leaves T[k1][n2] * W_25^(n2 k1) in v[5 k1 + n2]
FMR_IP_HD void dft25_head(float2 (&v)[25]) {
#pragma unroll
  for (int n2 = 0; n2 < 5; n2++) dft5(v[n2], v[5 + n2], v[10 + n2], v[15 + n2], v[20 + n2]);
#pragma unroll
  for (int k1 = 1; k1 < 5; k1++) {
#pragma unroll
    for (int n2 = 1; n2 < 5; n2++) v[5 * k1 + n2] = cmul(v[5 * k1 + n2], w25(n2 * k1));
  }
}
// 25-point forward DFT in registers; X[k] is left in v[nat25(k)]
FMR_IP_HD void dft25(float2 (&v)[25]) {
  dft25_head(v);
#pragma unroll
  for (int k1 = 0; k1 < 5; k1++) dft5(v[5 * k1], v[5 * k1 + 1], v[5 * k1 + 2], v[5 * k1 + 3], v[5 * k1 + 4]);
}
FMR_IP_HD int nat25(int k) { return 5 * (k % 5) + k / 5; }
// the 2 (L + 1) outputs inside the band: o[e] = X[k3], k3 = e (e <= L), 25 - 2 (L + 1) + e (e > L); L = 3 or 4.
// X[k1 + 5 k2] is output k2 of the k1-th second-step DFT: the band needs k2 = 0 (k3 = k1 <= L) and k2 = 4 (k3 = 20 + k1).
template <int L> FMR_IP_HD void dft25_band(float2 (&v)[25], float2 (&o)[2 * (L + 1)]) {
  static_assert(L == 3 || L == 4, "band width");
  dft25_head(v);
#pragma unroll
  for (int k1 = 0; k1 < 5; k1++) {
    const bool lo = k1 <= L, hi = 20 + k1 >= 24 - L;
    if (lo && !hi) {
      o[k1] = cadd(cadd(v[5 * k1], cadd(v[5 * k1 + 1], v[5 * k1 + 4])), cadd(v[5 * k1 + 2], v[5 * k1 + 3]));
    } else {
      float2 t0 = v[5 * k1];
      dft5_04(t0, v[5 * k1 + 1], v[5 * k1 + 2], v[5 * k1 + 3], v[5 * k1 + 4]);
      if (lo) o[k1] = t0;
      if (hi) o[(L + 1) + (20 + k1) - (24 - L)] = v[5 * k1 + 4];
    }
  }
}

// ---- 12-point forward DFT in registers (n = 3 n1 + n2, k = k1 + 4 k2); X[k] is left in v[3 (k & 3) + (k >> 2)]
FMR_IP_HD void dft12(float2 (&v)[12]) {
#pragma unroll
  for (int n2 = 0; n2 < 3; n2++) fft4(v[n2], v[3 + n2], v[6 + n2], v[9 + n2]);
  const float c = 0.86602540378443864676f;
  v[3 + 1] = cmul(v[3 + 1], mk(c, -0.5f));   // W_12^1
  v[6 + 1] = cmul(v[6 + 1], mk(0.5f, -c));   // W_12^2
  v[9 + 1] = mulmj(v[9 + 1]);                // W_12^3
  v[3 + 2] = cmul(v[3 + 2], mk(0.5f, -c));   // W_12^2
  v[6 + 2] = cmul(v[6 + 2], mk(-0.5f, -c));  // W_12^4
  v[9 + 2] = mk(-v[9 + 2].x, -v[9 + 2].y);   // W_12^6
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    const float2 x0 = v[3 * k1], s = cadd(v[3 * k1 + 1], v[3 * k1 + 2]), d = csub(v[3 * k1 + 1], v[3 * k1 + 2]);
    const float2 m = sfma(-0.5f, s, x0), jd = smul(c, d);
    v[3 * k1] = cadd(x0, s);
    v[3 * k1 + 1] = cadd(m, mulmj(jd));
    v[3 * k1 + 2] = cadd(m, mulpj(jd));
  }
}
FMR_IP_HD int nat12(int k) { return 3 * (k & 3) + (k >> 2); }

// ---- 3-point forward DFT
FMR_IP_HD void dft3(float2 &x0, float2 &x1, float2 &x2) {
  const float c = 0.86602540378443864676f;
  const float2 s = cadd(x1, x2), d = csub(x1, x2);
  const float2 m = sfma(-0.5f, s, x0), jd = smul(c, d);
  x0 = cadd(x0, s);
  x1 = cadd(m, mulmj(jd));
  x2 = cadd(m, mulpj(jd));
}
// ---- 15-point forward DFT in registers (n = 3 n1 + n2, k = k1 + 5 k2); X[k] is left in v[3 (k % 5) + k / 5]
FMR_IP_HD void dft15(float2 (&v)[15]) {
#pragma unroll
  for (int n2 = 0; n2 < 3; n2++) dft5(v[n2], v[3 + n2], v[6 + n2], v[9 + n2], v[12 + n2]);
  // W_15^(n2 k1)
  v[3 + 1] = cmul(v[3 + 1], mk(0.9135454576426008666f, -0.40673664307580015276f));   // 1
  v[6 + 1] = cmul(v[6 + 1], mk(0.66913060635885823757f, -0.7431448254773941331f));   // 2
  v[9 + 1] = cmul(v[9 + 1], mk(0.30901699437494745126f, -0.95105651629515353118f));  // 3
  v[12 + 1] = cmul(v[12 + 1], mk(-0.10452846326765333207f, -0.99452189536827340088f)); // 4
  v[3 + 2] = cmul(v[3 + 2], mk(0.66913060635885823757f, -0.7431448254773941331f));   // 2
  v[6 + 2] = cmul(v[6 + 2], mk(-0.10452846326765333207f, -0.99452189536827340088f)); // 4
  v[9 + 2] = cmul(v[9 + 2], mk(-0.80901699437494734024f, -0.58778525229247324813f)); // 6
  v[12 + 2] = cmul(v[12 + 2], mk(-0.97814760073380568883f, 0.20791169081775906502f)); // 8
#pragma unroll
  for (int k1 = 0; k1 < 5; k1++) dft3(v[3 * k1], v[3 * k1 + 1], v[3 * k1 + 2]);
}
// ---- 10-point forward DFT in registers (n = 2 n1 + n2, k = k1 + 5 k2); X[k] is left in v[2 (k % 5) + k / 5]
FMR_IP_HD void dft10(float2 (&v)[10]) {
#pragma unroll
  for (int n2 = 0; n2 < 2; n2++) dft5(v[n2], v[2 + n2], v[4 + n2], v[6 + n2], v[8 + n2]);
  v[2 + 1] = cmul(v[2 + 1], mk(0.80901699437494745126f, -0.5877852522924731371f));   // W_10^1
  v[4 + 1] = cmul(v[4 + 1], mk(0.30901699437494745126f, -0.95105651629515353118f));  // 2
  v[6 + 1] = cmul(v[6 + 1], mk(-0.30901699437494734024f, -0.9510565162951536422f));  // 3
  v[8 + 1] = cmul(v[8 + 1], mk(-0.80901699437494734024f, -0.58778525229247324813f)); // 4
#pragma unroll
  for (int k1 = 0; k1 < 5; k1++) {
    const float2 a = v[2 * k1], b = v[2 * k1 + 1];
    v[2 * k1] = cadd(a, b);
    v[2 * k1 + 1] = csub(a, b);
  }
}
// last pass of the inverse: RL-point DFT in registers, X[k] is left in v[nat_last<RL>(k)]
template <int RL> FMR_IP_HD void dft_last(float2 (&v)[RL]) {
  static_assert(RL == 12 || RL == 15 || RL == 10, "inverse size");
  if constexpr (RL == 12) dft12(v);
  if constexpr (RL == 15) dft15(v);
  if constexpr (RL == 10) dft10(v);
}
template <int RL> FMR_IP_HD int nat_last(int k) {
  return (RL == 12) ? 3 * (k & 3) + (k >> 2) : (RL == 15) ? 3 * (k % 5) + k / 5 : 2 * (k % 5) + k / 5;
}

// ---- forward pass 1: radix 16, stride 625. b in [0, 625); x[b + 625 a] comes from ld(b, a)
template <typename LD> FMR_IP_HD void fwd1(int b, LD ld, float2 *A, const float2 *__restrict__ tab) {
  float2 v[16];
#pragma unroll
  for (int a = 0; a < 16; a++) v[a] = ld(b, a);
  fft16(v);
  float2 w[16];
  powers16(tab[kTw1 + b], w);
  float2 *dst = A + b;
  dst[0] = v[nat(0)];
#pragma unroll
  for (int d = 1; d < 16; d++) dst[625 * d] = cmul(v[nat(d)], w[d]);
}
// ---- forward pass 2: radix 25 inside chunk k1, stride 25. i in [0, 400): k1 = i & 15, n3 = i >> 4
// (lanes = consecutive k1: 625 = 1 mod 16, conflict free)
FMR_IP_HD void fwd2(int i, float2 *A, const float2 *__restrict__ tab) {
  const int k1 = i & 15, n3 = i >> 4;
  float2 *p = A + 625 * k1 + n3;
  float2 v[25];
#pragma unroll
  for (int a = 0; a < 25; a++) v[a] = p[25 * a];
  dft25(v);
  // W_625^(n3 k2), k2 = 5 g + c: (w^(5 g)) * (w^c), both factors straight from the table (one rounding each)
  const float2 *__restrict__ t = tab + kTw2 + 8 * n3;
  const float2 wc[5] = {mk(1.f, 0.f), t[0], t[1], t[2], t[3]};
  const float2 wg[5] = {mk(1.f, 0.f), t[4], t[5], t[6], t[7]};
  p[0] = v[nat25(0)];
#pragma unroll
  for (int k2 = 1; k2 < 25; k2++) {
    const int g = k2 / 5, c = k2 % 5;
    const float2 w = (g == 0) ? wc[c] : (c == 0) ? wg[g] : cmul(wg[g], wc[c]);
    p[25 * k2] = cmul(v[nat25(k2)], w);
  }
}
// ---- forward pass 3: the band outputs of the radix-25 butterfly over 25 contiguous slots, times H0, conjugated.
// i in [0, 400): k1 = i & 15, k2 = i >> 4, and r = k1 + 16 k2 = i.
template <int RL> FMR_IP_HD void fwd3_compute(int i, const float2 *A, const float *__restrict__ Hs, float2 (&o)[Geo<RL>::kKeep]) {
  const float2 *p = A + 625 * (i & 15) + 25 * (i >> 4);
  float2 v[25];
#pragma unroll
  for (int a = 0; a < 25; a++) v[a] = p[a];
  dft25_band<Geo<RL>::kL>(v, o);
#pragma unroll
  for (int e = 0; e < Geo<RL>::kKeep; e++) {
    const float h = Hs[400 * e + i];
    o[e] = mk(h * o[e].x, -h * o[e].y);
  }
}
// natural-order bin of band output e of butterfly r, or -1 if it lies outside the band
template <int RL> FMR_IP_HD int band_bin(int r, int e) {
  using G = Geo<RL>;
  if (e <= G::kL) {
    const int k = r + 400 * e;
    return (k < G::kHalf) ? k : -1;
  }
  const int k = r + 400 * (25 - G::kKeep + e) - kNin; // negative frequency
  return (k >= -G::kHalf) ? k + G::kNout : -1;
}
template <int RL> FMR_IP_HD void fwd3_store(int i, float2 *Z, const float2 (&o)[Geo<RL>::kKeep]) {
#pragma unroll
  for (int e = 0; e < Geo<RL>::kKeep; e++) {
    const int k = band_bin<RL>(i, e);
    if (k >= 0) Z[zpos<RL>(k)] = o[e];
  }
}
// ---- inverse (run as a forward transform of the conjugated spectrum), pass 1: radix 16, stride kChunk. b in [0, kChunk)
template <int RL> FMR_IP_HD void inv1(int b, float2 *Z, const float2 *__restrict__ tab) {
  constexpr int S = Geo<RL>::kChunk + 1; // zpos(b + kChunk a) = b + (kChunk + 1) a
  float2 *p = Z + b;
  float2 v[16];
#pragma unroll
  for (int a = 0; a < 16; a++) v[a] = p[S * a];
  fft16(v);
  float2 w[16];
  powers16(tab[Geo<RL>::kTi1 + b], w);
  p[0] = v[nat(0)];
#pragma unroll
  for (int d = 1; d < 16; d++) p[S * d] = cmul(v[nat(d)], w[d]);
}
// ---- pass 2: radix 16 inside chunk i1, stride RL. u in [0, kChunk): i1 = u & 15, k3 = u >> 4 (kChunk + 1 = 1 mod 16)
template <int RL> FMR_IP_HD void inv2(int u, float2 *Z, const float2 *__restrict__ tab) {
  const int i1 = u & 15, k3 = u >> 4;
  float2 *p = Z + (Geo<RL>::kChunk + 1) * i1 + k3;
  float2 v[16];
#pragma unroll
  for (int a = 0; a < 16; a++) v[a] = p[RL * a];
  fft16(v);
  float2 w[16];
  powers16(tab[Geo<RL>::kTi2 + k3], w);
  p[0] = v[nat(0)];
#pragma unroll
  for (int d = 1; d < 16; d++) p[RL * d] = cmul(v[nat(d)], w[d]);
}
// ---- pass 3: radix RL over RL contiguous slots; t in [0, 256): i1 = t & 15, i2 = t >> 4; output sample
// i = t + 256 i3 of the block (i3 in [0, RL)) goes to st(i, value). Lanes = consecutive output samples.
template <int RL, typename ST> FMR_IP_HD void inv3(int t, const float2 *Z, ST st) {
  const float2 *p = Z + (Geo<RL>::kChunk + 1) * (t & 15) + RL * (t >> 4);
  float2 v[RL];
#pragma unroll
  for (int a = 0; a < RL; a++) v[a] = p[a];
  dft_last<RL>(v);
#pragma unroll
  for (int i3 = 0; i3 < RL; i3++) st(t + 256 * i3, cconj(v[nat_last<RL>(i3)]));
}

} // namespace fdr
} // namespace fmr

#if defined(__CUDACC__) && defined(FMR_KERNELS_CUH)
namespace fmr {
// k_fdr: one CTA = one block of the absolute block grid of one channel. Block j holds input samples
// [adv_in j - guard_in, adv_in j - guard_in + 10000) of the ring in front of the low-pass and yields outputs
// [adv_out j, adv_out (j + 1)) of the resampled stream (625:192: 7500 / 1250 / 2304); only outputs in [m_lo, m_hi) are
// stored. Input samples at negative indices or at indices >= avail read as zero, as in k_fir_fft.
// The grid is absolute, so the result does not depend on how the stream was cut into calls.
struct FdrParams {
  int64_t j0;         // block index of blockIdx.x == 0
  int64_t m_lo, m_hi; // outputs this launch stores
  int64_t avail;      // valid input samples in the ring
  int adv_in, guard_in, adv_out, guard_out;
};
constexpr int kFdrThreads = 256;
template <int RL> constexpr int fdr_smem_bytes() { return (fdr::kNin + fdr::Geo<RL>::kZLen) * (int)sizeof(float2); }

template <int RL>
__global__ void __launch_bounds__(kFdrThreads, 2)
    k_fdr(Ring<float2> in, Ring<float2> out, const float *__restrict__ Hs, const float2 *__restrict__ tab, FdrParams P) {
  using namespace fdr;
  using G = Geo<RL>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *A = reinterpret_cast<float2 *>(smem_raw);
  float2 *Z = A + kNin;
  const uint32_t c = blockIdx.y;
  const int64_t j = P.j0 + blockIdx.x;
  const int64_t base = j * P.adv_in - P.guard_in;
  const int tid = threadIdx.x;
  {
    const uint32_t pos0 = (uint32_t)base & (in.cap - 1);
    if (base >= 0 && base + kNin <= P.avail && pos0 + (uint32_t)kNin <= in.cap) {
      const float2 *__restrict__ row = in.base + (size_t)c * in.cap + pos0;
      for (int b = tid; b < 625; b += kFdrThreads) fwd1(b, [&](int bb, int a) { return row[bb + 625 * a]; }, A, tab);
    } else {
      for (int b = tid; b < 625; b += kFdrThreads) {
        fwd1(b,
             [&](int bb, int a) {
               const int64_t t = base + bb + 625 * a;
               return (t < P.avail) ? in.ld(c, t) : make_float2(0.f, 0.f);
             },
             A, tab);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 400; i += kFdrThreads) fwd2(i, A, tab);
  __syncthreads();
  for (int i = tid; i < 400; i += kFdrThreads) {
    float2 o[G::kKeep];
    fwd3_compute<RL>(i, A, Hs, o);
    fwd3_store<RL>(i, Z, o);
  }
  __syncthreads();
  if (tid < G::kChunk) inv1<RL>(tid, Z, tab);
  __syncthreads();
  if (tid < G::kChunk) inv2<RL>(tid, Z, tab);
  __syncthreads();
  const int64_t mb = j * P.adv_out - P.guard_out; // output index of block sample 0
  const int64_t lo = max(P.m_lo, j * P.adv_out), hi = min(P.m_hi, (j + 1) * P.adv_out);
  inv3<RL>(tid, Z, [&](int i, float2 v) {
    const int64_t m = mb + i;
    if (m >= lo && m < hi) out.st(c, m, v);
  });
}
} // namespace fmr
#endif

// ---- host side: tables (double precision, rounded once)
#include <cmath>
#include <vector>
namespace fmr {
namespace fdr {
inline float2 fdr_w(double num, double den) { // W_den^num
  const double a = -2.0 * 3.14159265358979323846264338327950288 * std::fmod(num, den) / den;
  return mk((float)std::cos(a), (float)std::sin(a));
}
template <int RL>
inline void fdr_make_tables(const double *taps, int klen, std::vector<float2> &tab, std::vector<float> &Hs) {
  using G = Geo<RL>;
  tab.assign(G::kTabLen, mk(0.f, 0.f));
  for (int b = 0; b < 625; b++) tab[kTw1 + b] = fdr_w(b, kNin);
  const int mult[8] = {1, 2, 3, 4, 5, 10, 15, 20};
  for (int n3 = 0; n3 < 25; n3++) {
    for (int q = 0; q < 8; q++) tab[kTw2 + 8 * n3 + q] = fdr_w((double)n3 * mult[q], 625.0);
  }
  for (int b = 0; b < G::kChunk; b++) tab[G::kTi1 + b] = fdr_w(b, G::kNout);
  for (int k3 = 0; k3 < RL; k3++) tab[G::kTi2 + k3] = fdr_w(k3, G::kChunk);
  // H0[k] = sum_i h[i] cos(2 pi k (i - fl2) / N) (zero phase, real for the symmetric low-pass), with 1/N folded in
  Hs.assign(G::kHsLen, 0.f);
  const int fl2 = (klen - 1) / 2;
  for (int e = 0; e < G::kKeep; e++) {
    for (int r = 0; r < 400; r++) {
      if (band_bin<RL>(r, e) < 0) continue;
      const int k = (e <= G::kL) ? r + 400 * e : r + 400 * (25 - G::kKeep + e) - kNin; // signed frequency
      double acc = 0.0;
      for (int i = 0; i < klen; i++) {
        const long long ph = ((long long)k * (i - fl2)) % kNin;
        acc += taps[i] * std::cos(2.0 * 3.14159265358979323846264338327950288 * (double)ph / kNin);
      }
      Hs[400 * e + r] = (float)(acc / kNin);
    }
  }
}
} // namespace fdr
} // namespace fmr
#endif
