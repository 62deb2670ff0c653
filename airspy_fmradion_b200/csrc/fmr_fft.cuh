// fmr_fft.cuh — the long zero-phase low-pass as overlap-save fast convolution, entirely in
// shared memory (reference: r8b::CDSPBlockConvolver::process, CDSPBlockConvolver.h:252-353,
// which does the same linear convolution with a double-precision FFT of 8192/16384 points).
//
// One CTA filters one block of one channel: N complex points live in shared memory
// (N = 16384 FP32: 136 KB; N = 8192 FP32: 68 KB; N = 8192 FP64: 139 KB); forward FFT, multiply
// by the filter spectrum H (precomputed on the host in double, 1/N folded in), inverse FFT via
// the conjugation identity, and only the N-(klen-1) alias-free outputs are used. A complex
// stream rides through one complex FFT (the filter is real): I/Q for the IF chain, (mono, L-R)
// for the two audio resamplers that the reference runs in lock step (FmDecode.cpp:172-183).
//
// FFT: in-place Stockham autosort, radix 16,16,16,R (R = N/4096 = 4 or 2; DIT twiddles on the
// inputs), every thread keeps its butterfly inputs in registers across the barrier so one buffer
// suffices. The first forward pass reads straight from the global ring, the last forward pass
// applies H and the conjugation: 7 shared-memory round trips per block. Shared-memory layout
// is skewed (one pad slot per 16) so both the strided reads and the stride-16 writes are
// bank-conflict free.
//
// Epilogue, two forms:
//   plain : the last inverse pass writes the alias-free outputs (every `down`-th) to the ring;
//   fused : (IF chain) the filtered block stays in shared memory and the whole-step polyphase
//           interpolator (r8b::CDSPFracInterpolator::convolve0, CDSPFracInterpolator.h:992-1060)
//           is applied from there, so the 1.25 MHz intermediate stream never touches HBM.
//           Blocks are placed so that every interpolator window lies inside one block's
//           alias-free region (consecutive blocks overlap by flen extra samples).
#ifndef FMR_FFT_CUH
#define FMR_FFT_CUH

#include "fmr_kernels.cuh"

namespace fmr {

constexpr int kFftThreads = 512;
// Build-time experiment (-DFMR_FFT_REGCAP_THREADS=608): compiling the 16384-point FP32 kernel for a nominal
// 608-thread block caps it at 96 registers per thread (no spills, +0.8 % run time) instead of 111; at 512 threads
// that leaves room for one CTA of the fused 384 kHz core (128 threads x 96 registers) on the same SM, which the
// time-chunk pipeline (fmr_fm.cu, FMR_TIME_CHUNKS + FMR_FUSED_CHUNKS) needs to run the latency-bound core of chunk
// k underneath the front end of chunk k+1. Measured: the overlap happens but chunking costs more than it hides
// (profiles/README.md), so the default stays 512 = no cap.
#ifndef FMR_FFT_REGCAP_THREADS
#define FMR_FFT_REGCAP_THREADS 512
#endif
constexpr int kFftRegCapThreads = FMR_FFT_REGCAP_THREADS;

template <typename S, int N, int THREADS = kFftThreads> struct FftCfg {
  using V = typename V2<S>::type;
  static constexpr int kN = N;
  static constexpr int kThreads = THREADS;      // threads per CTA (512; 1024 = the 32-warp form of the 16384-point kernel)
  static constexpr int kQ = N / 16;             // stride of the radix-16 passes
  static constexpr int kSets = N / 16 / THREADS; // butterflies of 16 per thread per pass
  static constexpr int kR4 = N / 4096;          // radix of the last pass
  static constexpr int kSmemBytes = (N + N / 16) * (int)sizeof(V) + 256 * (int)sizeof(V);
  static constexpr int kSmemBytesTw = kSmemBytes + (256 + 4096) * (int)sizeof(V); // + the full twiddle tables
  static_assert(kSets >= 1 && (kR4 == 2 || kR4 == 4), "unsupported FFT size");
};

__device__ __forceinline__ int fpad(int n) { return n + (n >> 4); }

template <typename V> __device__ __forceinline__ V cmk(decltype(V().x) a, decltype(V().x) b) {
  V v;
  v.x = a;
  v.y = b;
  return v;
}
template <typename V> __device__ __forceinline__ V cmulv(V a, V b) {
  return cmk<V>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename V> __device__ __forceinline__ V cconjv(V a) { return cmk<V>(a.x, -a.y); }

template <typename V> __device__ __forceinline__ void fft4(V &a0, V &a1, V &a2, V &a3) {
  const V s02 = cmk<V>(a0.x + a2.x, a0.y + a2.y);
  const V d02 = cmk<V>(a0.x - a2.x, a0.y - a2.y);
  const V s13 = cmk<V>(a1.x + a3.x, a1.y + a3.y);
  const V d13 = cmk<V>(a1.x - a3.x, a1.y - a3.y);
  a0 = cmk<V>(s02.x + s13.x, s02.y + s13.y);
  a2 = cmk<V>(s02.x - s13.x, s02.y - s13.y);
  a1 = cmk<V>(d02.x + d13.y, d02.y - d13.x); // d02 - j d13
  a3 = cmk<V>(d02.x - d13.y, d02.y + d13.x); // d02 + j d13
}
// FP32: a complex add is one packed FADD2, a complex subtract one FFMA2 with the scalar -1 (exact,
// the same IEEE results as the scalar form) — Blackwell's packed FP32 pipe issues both lanes of a
// (re, im) pair in one slot. Only the two "times -+j" outputs need a component swap.
__device__ __forceinline__ float2 csub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
template <> __device__ __forceinline__ void fft4<float2>(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  const float2 s02 = __fadd2_rn(a0, a2);
  const float2 d02 = csub2(a0, a2);
  const float2 s13 = __fadd2_rn(a1, a3);
  const float2 d13 = csub2(a1, a3);
  a0 = __fadd2_rn(s02, s13);
  a2 = csub2(s02, s13);
  const float2 t = make_float2(d13.y, -d13.x); // -j d13
  a1 = __fadd2_rn(d02, t);
  a3 = csub2(d02, t);
}

// 16-point forward DFT in registers; result X[m] is left in v[4*(m&3) + (m>>2)].
template <typename V> __device__ __forceinline__ void fft16(V (&v)[16]) {
  using S = decltype(V().x);
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) fft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
  // twiddles W16^(n2*k1), element v[n2 + 4*k1]
  const S c1 = (S)0.92387953251128675613, s1 = (S)0.38268343236508977173, r2 = (S)0.70710678118654752440;
  // k1 = 1: W^n2
  v[1 + 4] = cmulv(v[1 + 4], cmk<V>(c1, -s1));
  v[2 + 4] = cmulv(v[2 + 4], cmk<V>(r2, -r2));
  v[3 + 4] = cmulv(v[3 + 4], cmk<V>(s1, -c1));
  // k1 = 2: W^(2 n2)
  v[1 + 8] = cmulv(v[1 + 8], cmk<V>(r2, -r2));
  v[2 + 8] = cmk<V>(v[2 + 8].y, -v[2 + 8].x); // W^4 = -j
  v[3 + 8] = cmulv(v[3 + 8], cmk<V>(-r2, -r2));
  // k1 = 3: W^(3 n2)
  v[1 + 12] = cmulv(v[1 + 12], cmk<V>(s1, -c1));
  v[2 + 12] = cmulv(v[2 + 12], cmk<V>(-r2, -r2));
  v[3 + 12] = cmulv(v[3 + 12], cmk<V>(-c1, s1)); // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) fft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// W_N^m from the two-level table: tw[0..127] = W^(128 q), tw[128..255] = W^l.
template <typename V> __device__ __forceinline__ V tw_lookup(const V *tw, int m) {
  return cmulv(tw[m >> 7], tw[128 + (m & 127)]);
}

// apply w^r (r = 1..15) to v[r]
template <typename V> __device__ __forceinline__ void twiddle16(V (&v)[16], V w1) {
  const V w2 = cmulv(w1, w1);
  const V w3 = cmulv(w2, w1);
  const V w4 = cmulv(w2, w2);
  const V w5 = cmulv(w4, w1);
  const V w6 = cmulv(w3, w3);
  const V w7 = cmulv(w4, w3);
  const V w8 = cmulv(w4, w4);
  v[1] = cmulv(v[1], w1);
  v[2] = cmulv(v[2], w2);
  v[3] = cmulv(v[3], w3);
  v[4] = cmulv(v[4], w4);
  v[5] = cmulv(v[5], w5);
  v[6] = cmulv(v[6], w6);
  v[7] = cmulv(v[7], w7);
  v[8] = cmulv(v[8], w8);
  v[9] = cmulv(v[9], cmulv(w8, w1));
  v[10] = cmulv(v[10], cmulv(w8, w2));
  v[11] = cmulv(v[11], cmulv(w8, w3));
  v[12] = cmulv(v[12], cmulv(w8, w4));
  v[13] = cmulv(v[13], cmulv(w8, w5));
  v[14] = cmulv(v[14], cmulv(w8, w6));
  v[15] = cmulv(v[15], cmulv(w8, w7));
}

// First pass of a transform whose input already sits in the buffer (p = 1, no twiddles).
template <typename CFG> __device__ __forceinline__ void pass16_first_smem(typename CFG::V *buf) {
  using V = typename CFG::V;
  V v[CFG::kSets][16];
  // fpad(i + r*Q) = fpad(i) + r*(Q + Q/16) and fpad(16 i + r) = 17 i + r: one base address per
  // set, compile-time offsets for the 16 elements
#pragma unroll
  for (int s = 0; s < CFG::kSets; s++) {
    const V *src = buf + fpad(threadIdx.x + s * CFG::kThreads);
#pragma unroll
    for (int r = 0; r < 16; r++) v[s][r] = src[r * (CFG::kQ + CFG::kQ / 16)];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < CFG::kSets; s++) {
    V *dst = buf + 17 * (threadIdx.x + s * CFG::kThreads);
    fft16(v[s]);
#pragma unroll
    for (int r = 0; r < 16; r++) dst[r] = v[s][4 * (r & 3) + (r >> 2)];
  }
  __syncthreads();
}

// Full twiddle tables of the two twiddled radix-16 passes (TW instantiations): entry (r, k) of pass p is
// W_N^(k * scale_p * r) = exp(-2 pi i k r / (16 p)), stored [r][k] so that consecutive lanes (consecutive k) read
// consecutive words. 16 x 16 entries for p = 16 followed by 16 x 256 for p = 256; the values do not depend on N.
constexpr int kTwTabP16 = 0, kTwTabP256 = 256, kTwTabLen = 256 + 4096;

// One radix-16 Stockham pass over the whole buffer, in place: p = 16 or 256. With TW the fifteen twiddles of a
// butterfly come from the table (fifteen independent shared-memory loads) instead of one looked-up root and a
// chain of fourteen complex multiplications for its powers.
template <typename CFG, int p, bool TW = false>
__device__ __forceinline__ void pass16_smem(typename CFG::V *buf, const typename CFG::V *tw,
                                            const typename CFG::V *twtab = nullptr) {
  using V = typename CFG::V;
  constexpr int scale = CFG::kN / (16 * p);
  V v[CFG::kSets][16];
#pragma unroll
  for (int s = 0; s < CFG::kSets; s++) {
    const V *src = buf + fpad(threadIdx.x + s * CFG::kThreads);
#pragma unroll
    for (int r = 0; r < 16; r++) v[s][r] = src[r * (CFG::kQ + CFG::kQ / 16)];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < CFG::kSets; s++) {
    const int i = threadIdx.x + s * CFG::kThreads;
    const int k = i & (p - 1);
    if (TW) {
      const V *tp = twtab + (p == 16 ? kTwTabP16 : kTwTabP256) + k;
#pragma unroll
      for (int r = 1; r < 16; r++) v[s][r] = cmulv(v[s][r], tp[r * p]);
    } else {
      twiddle16(v[s], tw_lookup(tw, k * scale));
    }
    fft16(v[s]);
    const int j = (i - k) * 16 + k;
    // p is 16 or 256: fpad(j + r*p) = fpad(j) + r*(p + p/16)
    V *dst = buf + fpad(j);
    const int step = p + (p >> 4);
#pragma unroll
    for (int r = 0; r < 16; r++) dst[r * step] = v[s][4 * (r & 3) + (r >> 2)];
  }
  __syncthreads();
}

// Last pass (radix 4 or 2, p = 4096) on the elements {i + r*4096}; in place: a[r] <- X[i + r*4096].
template <typename CFG>
__device__ __forceinline__ void last_pass_load(const typename CFG::V *buf, const typename CFG::V *tw, int i,
                                               typename CFG::V (&a)[4]) {
  using V = typename CFG::V;
  const V w1 = tw_lookup(tw, i);
  if (CFG::kR4 == 4) {
    const V *src = buf + fpad(i);
    a[0] = src[0];
    a[1] = src[4352];
    a[2] = src[2 * 4352];
    a[3] = src[3 * 4352];
    const V w2 = cmulv(w1, w1);
    a[1] = cmulv(a[1], w1);
    a[2] = cmulv(a[2], w2);
    a[3] = cmulv(a[3], cmulv(w2, w1));
    fft4(a[0], a[1], a[2], a[3]);
  } else {
    const V x0 = buf[fpad(i)];
    const V x1 = cmulv(buf[fpad(i) + 4352], w1);
    a[0] = cmk<V>(x0.x + x1.x, x0.y + x1.y);
    a[1] = cmk<V>(x0.x - x1.x, x0.y - x1.y);
    a[2] = a[0];
    a[3] = a[0];
  }
}

// Whole-step polyphase interpolator (r8b::CDSPFracInterpolator::convolve0) over one filtered block
// in shared memory (plain order). Output i = p + outstep*q of the block has bank row
// ph(p) = (p*instep + rem) mod outstep and its window starts instep*q samples after that of
// output p. So the warp fixes p — one bank row, loaded once into registers and shared by all
// lanes — and spreads q over the lanes: consecutive lanes then read addresses `instep` elements
// apart, and instep is odd for every shipped chain, i.e. conflict free. When a block holds fewer
// than 32 periods the warp takes several p at once (G groups of L lanes).
template <typename V, int FLEN, int NT>
__device__ __forceinline__ void fi_epilogue(const V *__restrict__ buf, const decltype(V().x) *__restrict__ bank, int instep,
                                            int outstep, int klen, int rem_b, int cnt, Ring<V> out, uint32_t c, int64_t mb) {
  using S = decltype(V().x);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = NT / 32;
  const int nq = (cnt + outstep - 1) / outstep;
  const int L = nq < 32 ? nq : 32;
  const int G = 32 / L;
  const int sgrp = lane / L, ql = lane - sgrp * L;
  if (sgrp >= G) return;
  for (int pg = warp * G; pg < outstep; pg += nwarps * G) {
    const int p = pg + sgrp;
    if (p >= outstep) continue;
    const int pp = p * instep + rem_b;
    const int dp = pp / outstep;
    const int ph = pp - dp * outstep;
    const S *__restrict__ row = bank + (size_t)ph * FLEN;
    S h[FLEN];
#pragma unroll
    for (int k = 0; k < FLEN; k++) h[k] = __ldg(row + k);
    for (int q = ql; q < nq; q += L) {
      const int i = p + outstep * q;
      if (i < cnt) {
        const V *__restrict__ w = buf + (klen - 1) + dp + instep * q;
        V acc = cmk<V>(0, 0);
#pragma unroll
        for (int k = 0; k < FLEN; k++) {
          const V x = w[k];
          acc.x += h[k] * x.x;
          acc.y += h[k] * x.y;
        }
        out.st(c, mb + i, acc);
      }
    }
  }
}

struct FftFuse {
  const void *bank; // [outstep][flen] interpolator bank, scalar type of the chain
  int instep, outstep, flen;
  int64_t m0;       // first interpolator output of this launch
  int n_m;          // interpolator outputs of this launch
  int mo;           // interpolator outputs per block
  // The last block of a call also leaves the newest filtered samples [tail_lo, tail_hi) in the
  // intermediate ring (`tail_base`, capacity `tail_cap`), where the unfused kernels of a later,
  // smaller call expect the interpolator's history. tail_hi <= tail_lo disables it.
  void *tail_base;
  uint32_t tail_cap;
  int64_t tail_lo, tail_hi;
  const void *twtab; // TW instantiations: [kTwTabLen] twiddles in the chain's complex type (global memory)
};

// k_fir_fft: y[q] = sum_j h[j] x[q*down - fl2 + j].
//   plain: q in [q0, q0+n_out), lq = (N - klen + 1) / down outputs per block.
//   fused: block b produces interpolator outputs m in [m0 + b*mo, ...) from the filtered samples
//          it holds (down must be 1); samples with negative index read as zero like the ring does.
//   n_in_avail: number of valid input samples in the ring (indices >= it read as zero)
template <typename S, int N, bool FUSE, bool TW = false, int THREADS = kFftThreads>
__global__ void __launch_bounds__((THREADS != kFftThreads) ? THREADS
                                  : (N == 16384 && sizeof(S) == 4) ? kFftRegCapThreads : kFftThreads,
                                  (N == 8192 && sizeof(S) == 4) ? 2 : 1)
    k_fir_fft(Ring<typename V2<S>::type> in, Ring<typename V2<S>::type> out, const typename V2<S>::type *__restrict__ H,
              int klen, int down, int64_t q0, int n_out, int64_t n_in_avail, int lq, FftFuse fz) {
  using CFG = FftCfg<S, N, THREADS>;
  using V = typename CFG::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V *buf = reinterpret_cast<V *>(smem_raw);
  V *tw = buf + (N + N / 16);
  V *twtab = tw + 256;
  const uint32_t c = blockIdx.y;
  const int blk = blockIdx.x;
  int cnt;        // plain: filter outputs of this block; fused: interpolator outputs of this block
  int64_t qb;     // filter output index held by buffer slot klen-1
  int64_t mb = 0; // fused: first interpolator output of this block
  if (FUSE) {
    cnt = fz.n_m - blk * fz.mo;
    if (cnt > fz.mo) cnt = fz.mo;
    if (cnt <= 0) return;
    mb = fz.m0 + (int64_t)blk * fz.mo;
    qb = (mb * fz.instep) / fz.outstep - (fz.flen / 2 - 1);
  } else {
    cnt = n_out - blk * lq;
    if (cnt > lq) cnt = lq;
    if (cnt <= 0) return;
    qb = q0 + (int64_t)blk * lq;
  }
  const int fl2 = (klen - 1) / 2;
  const int64_t base = qb * down - fl2; // input sample index held by buffer slot 0
  // twiddle tables
  if (threadIdx.x < 256) {
    const int q = threadIdx.x;
    const int m = (q < 128) ? (q * 128) : (q - 128);
    S s, co;
    if (sizeof(S) == 4) {
      float fs, fc;
      sincospif(-2.0f * (float)m / (float)N, &fs, &fc);
      s = (S)fs;
      co = (S)fc;
    } else {
      double ds, dc;
      sincospi(-2.0 * (double)m / (double)N, &ds, &dc);
      s = (S)ds;
      co = (S)dc;
    }
    tw[q] = cmk<V>(co, s);
  }
  if (TW) {
    const V *__restrict__ g = reinterpret_cast<const V *>(fz.twtab);
    for (int i = threadIdx.x; i < kTwTabLen; i += CFG::kThreads) twtab[i] = __ldg(g + i);
  }
  // ---- forward pass 1 (p = 1, no twiddles), inputs straight from the global ring
  {
    V v[CFG::kSets][16];
    // block entirely inside the valid, unwrapped part of the ring (CTA-uniform): one pointer, constant offsets
    const uint32_t pos0 = (uint32_t)base & (in.cap - 1);
    if (TW && base >= 0 && base + N <= n_in_avail && pos0 + (uint32_t)N <= in.cap) {
      const V *__restrict__ row = in.base + (size_t)c * in.cap + pos0 + threadIdx.x;
#pragma unroll
      for (int s = 0; s < CFG::kSets; s++) {
#pragma unroll
        for (int r = 0; r < 16; r++) v[s][r] = row[s * CFG::kThreads + r * CFG::kQ];
      }
    } else {
#pragma unroll
      for (int s = 0; s < CFG::kSets; s++) {
        const int i = threadIdx.x + s * CFG::kThreads;
#pragma unroll
        for (int r = 0; r < 16; r++) {
          const int64_t t = base + i + r * CFG::kQ;
          v[s][r] = (t < n_in_avail) ? in.ld(c, t) : cmk<V>(0, 0);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < CFG::kSets; s++) {
      V *dst = buf + 17 * (threadIdx.x + s * CFG::kThreads);
      fft16(v[s]);
#pragma unroll
      for (int r = 0; r < 16; r++) dst[r] = v[s][4 * (r & 3) + (r >> 2)];
    }
    __syncthreads();
  }
  pass16_smem<CFG, 16, TW>(buf, tw, twtab);
  pass16_smem<CFG, 256, TW>(buf, tw, twtab);
  // ---- forward last pass fused with Y = conj(X * H)
#pragma unroll 2
  for (int b = 0; b < 4096 / CFG::kThreads; b++) {
    const int i = threadIdx.x + b * CFG::kThreads;
    V a[4];
    last_pass_load<CFG>(buf, tw, i, a);
#pragma unroll
    for (int r = 0; r < CFG::kR4; r++) buf[fpad(i) + r * 4352] = cconjv(cmulv(a[r], H[i + r * 4096]));
  }
  __syncthreads();
  // ---- inverse = conj(FFT(conj(.)))
  pass16_first_smem<CFG>(buf);
  pass16_smem<CFG, 16, TW>(buf, tw, twtab);
  pass16_smem<CFG, 256, TW>(buf, tw, twtab);
  if (!FUSE) {
    // ---- inverse last pass, outputs straight to the global ring (only the alias-free part)
#pragma unroll 2
    for (int b = 0; b < 4096 / CFG::kThreads; b++) {
      const int i = threadIdx.x + b * CFG::kThreads;
      V a[4];
      last_pass_load<CFG>(buf, tw, i, a);
#pragma unroll
      for (int r = 0; r < CFG::kR4; r++) {
        const int n = i + r * 4096; // buffer slot; output u sits at slot u*down + klen - 1
        const int rel = n - (klen - 1);
        if (rel >= 0) {
          const int u = rel / down;
          if (u * down == rel && u < cnt) out.st(c, qb + u, cconjv(a[r]));
        }
      }
    }
  } else {
    // ---- inverse last pass into registers, then back to shared memory in PLAIN order (slot n
    // holds filter output qb + n - (klen-1)): the interpolator below reads windows whose start
    // advances by `instep` per lane, which is conflict free only without the FFT's skew
    constexpr int NB = 4096 / CFG::kThreads;
    V y[NB][CFG::kR4];
#pragma unroll
    for (int b = 0; b < NB; b++) {
      const int i = threadIdx.x + b * CFG::kThreads;
      V a[4];
      last_pass_load<CFG>(buf, tw, i, a);
#pragma unroll
      for (int r = 0; r < CFG::kR4; r++) {
        const int n = i + r * 4096;
        const int64_t t = qb + (n - (klen - 1));
        y[b][r] = (t >= 0) ? cconjv(a[r]) : cmk<V>(0, 0);
        if (t >= fz.tail_lo && t < fz.tail_hi && n >= klen - 1 && blk == (int)gridDim.x - 1) {
          Ring<V>{reinterpret_cast<V *>(fz.tail_base), fz.tail_cap}.st(c, t, y[b][r]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; b++) {
#pragma unroll
      for (int r = 0; r < CFG::kR4; r++) buf[threadIdx.x + b * CFG::kThreads + r * 4096] = y[b][r];
    }
    __syncthreads();
    const S *__restrict__ bank = reinterpret_cast<const S *>(fz.bank);
    const int rem_b = (int)((mb * fz.instep) % fz.outstep);
    if (fz.flen == 18) {
      fi_epilogue<V, 18, CFG::kThreads>(buf, bank, fz.instep, fz.outstep, klen, rem_b, cnt, out, c, mb);
    } else if (fz.flen == 24) {
      fi_epilogue<V, 24, CFG::kThreads>(buf, bank, fz.instep, fz.outstep, klen, rem_b, cnt, out, c, mb);
    } else {
      // generic bank length: one output per thread, taps in a loop
      for (int i = threadIdx.x; i < cnt; i += CFG::kThreads) {
        const int prel = i * fz.instep + rem_b;
        const int dip = prel / fz.outstep;
        const int ph = prel - dip * fz.outstep;
        const S *__restrict__ row = bank + (size_t)ph * fz.flen;
        const int n0 = (klen - 1) + dip;
        V acc = cmk<V>(0, 0);
        for (int k = 0; k < fz.flen; k++) {
          const V x = buf[n0 + k];
          const S h = __ldg(row + k);
          acc.x += h * x.x;
          acc.y += h * x.y;
        }
        out.st(c, mb + i, acc);
      }
    }
  }
}

} // namespace fmr
#endif
