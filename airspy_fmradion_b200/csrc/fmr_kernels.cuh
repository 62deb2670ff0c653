// fmr_kernels.cuh — sm_100a device code of the demodulation hot path.
//
// Data layout in HBM: every inter-stage stream is a channel-major ring, element (c, i)
// at base[c*cap + (i & (cap-1))] where i is the ABSOLUTE sample index of that stream since
// the handle was created. Absolute indexing makes every stage a pure function of
// (input stream, output index range); the host's integer schedule (fmr_tables.h) decides
// which index ranges a process call covers. Samples at negative absolute indices are zero
// (the reference's zero-initialised delay lines).
#ifndef FMR_KERNELS_CUH
#define FMR_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace fmr {

template <typename S> struct V2;
template <> struct V2<float> {
  using type = float2;
};
template <> struct V2<double> {
  using type = double2;
};

// Sample type AV of the FM audio chain between the 384 kHz core and the 48 kHz DC block (audio half-bands, 1621-tap
// low-pass, pilot-cut FIR): float2 by default, double2 with FMR_AUDIO_FP64=1 at handle creation. The recurrences on
// either side (PLL, deemphasis, DC block) are FP64 like the reference in both cases; the linear filters between them run
// in FP32 like the IF resampler (audio error against the reference 1.5e-6 instead of 2e-7).
template <typename AV> __host__ __device__ inline AV aud_mk(double a, double b) {
  AV v;
  v.x = (decltype(v.x))a;
  v.y = (decltype(v.y))b;
  return v;
}

template <typename V> struct Ring {
  V *base;
  uint32_t cap; // power of two, per channel
  __device__ __forceinline__ V ld(uint32_t c, int64_t i) const {
    if (i < 0) {
      V z;
      z.x = 0;
      z.y = 0;
      return z;
    }
    return base[(size_t)c * cap + ((uint32_t)i & (cap - 1))];
  }
  __device__ __forceinline__ void st(uint32_t c, int64_t i, V v) const {
    base[(size_t)c * cap + ((uint32_t)i & (cap - 1))] = v;
  }
};

// Input of a resampler: either the caller's linear buffer for this call plus a short
// history of the previous call's tail, or a ring.
template <typename V> struct InSrc {
  const V *lin;      // [C][stride], sample 0 has absolute index `start`
  int fmt;           // 0: lin holds elements of V; 1: lin holds int16 (re,im) pairs, value/32768
                     // (what FileSource's sf_read_float yields for 16-bit PCM, FileSource.cpp:491-531)
  size_t stride;
  const V *hist;     // [C][HIST], absolute indices start-HIST .. start-1
  int64_t start;
  int64_t n_new;
  Ring<V> ring;
};
constexpr int kHist = 256;

template <typename V, bool LINEAR>
__device__ __forceinline__ V src_ld(const InSrc<V> &s, uint32_t c, int64_t i) {
  if (LINEAR) {
    V z;
    z.x = 0;
    z.y = 0;
    if (i < 0) return z;
    int64_t r = i - s.start;
    if (r < 0) {
      return (r >= -kHist) ? s.hist[(size_t)c * kHist + (kHist + r)] : z;
    }
    if (r >= s.n_new) return z;
    if (s.fmt == 1) {
      const short2 q = reinterpret_cast<const short2 *>(s.lin)[(size_t)c * s.stride + r];
      V v;
      v.x = (float)q.x * (1.0f / 32768.0f);
      v.y = (float)q.y * (1.0f / 32768.0f);
      return v;
    }
    return s.lin[(size_t)c * s.stride + r];
  } else {
    return s.ring.ld(c, i);
  }
}

// ---------------------------------------------------------------------------------------
// Half-band decimator cascade (reference: r8b::CDSPHBDownsampler::process,
// CDSPHBDownsampler.h:137-239; kernel form CDSPHBDownsampler.inc:620-629):
//   y[m] = x[2m] + sum_k t[k] * (x[2m+2k+1] + x[2m-2k-1]),   zero-phase, gain 2 per stage.
// NST = 0 degenerates to a copy (used for rates that need no half-band stage and to apply
// the Fs/4 shift of FourthConverterIQ, include/FourthConverterIQ.h:38-82: y[n] = x[n](-j)^n).
// One CTA computes kHbTile final-rate outputs of one channel; all intermediate rates live
// in shared memory, so the input is read from HBM exactly once.
template <typename S> struct HbTaps {
  int n[3];
  int sl[3]; // shared-memory sub-array length per level (host-computed, hb_sub_len)
  S t[3][14];
};
#ifndef FMR_HB_TILE
#define FMR_HB_TILE 256
#endif
#ifndef FMR_HB_THREADS
#define FMR_HB_THREADS 128
#endif
constexpr int kHbTile = FMR_HB_TILE;
constexpr int kHbThreads = FMR_HB_THREADS;
constexpr int kHbR = 4; // consecutive outputs per thread (register blocking)
// Final-rate outputs per CTA. A stage hands groups of kHbR outputs to the kHbThreads threads; the two-stage cascades of
// the 384 kHz -> 48 kHz chains (7 + 13 taps: FM audio resampler; 6 + 11: AM IF resampler) get the tile for which both stages fill whole rounds of threads (255 and 121 groups for 128
// threads; with 256 the rounds were 141 and 64 groups: half of the thread slots idle).
__host__ __device__ constexpr int hb_tile_of(int n1, int n2) {
  return ((n1 == 7 && n2 == 13) || (n1 == 6 && n2 == 11)) ? 484 : kHbTile; // (6, 11): 253 and 121 groups
}

// Shared-memory layout of one level: even- and odd-indexed samples in two separate arrays
// (E[m] = x[2m], O[j] = x[2j+1]) because a half-band output reads x[2m] and only ODD
// neighbours; each array is kHbR-way interleaved (element j lives in sub-array j % kHbR at
// position j / kHbR) so that a thread that owns kHbR consecutive outputs, and whose window
// of odd neighbours therefore advances kHbR elements per thread, still gives the warp
// contiguous (bank-conflict-free) addresses for every window slot.
__host__ __device__ inline int hb_level_len(const int *ntaps, int nst, int s, int tile) {
  int full = tile; // number of samples of level s needed by a full tile
  for (int q = nst; q > s; q--) full = 2 * full + 4 * ntaps[q - 1] - 3;
  return full;
}
// Per-sub-array length, padded so that the kHbR sub-arrays start 32 bytes apart modulo the
// 128-byte bank window: threads that scatter consecutive samples over the sub-arrays (level
// fill, stage write-back) then hit distinct banks. elem_bytes = sizeof(float2 / double2).
__host__ __device__ inline int hb_sub_len(int len, int elem_bytes) {
  int sl = (len / 2 + 2) / kHbR + 2;
  const int mod = (elem_bytes == 8) ? 16 : 8, want = (elem_bytes == 8) ? 4 : 2;
  while (sl % mod != want) sl++;
  return sl;
}
__device__ __forceinline__ int hb_pos(int j, int sl) { return (j % kHbR) * sl + j / kHbR; }

template <typename V> __device__ __forceinline__ V fs4_rot(V v, int64_t ai) {
  const int ph = (int)(ai & 3);
  V w;
  if (ph == 0) {
    w = v;
  } else if (ph == 1) {
    w.x = v.y;
    w.y = -v.x;
  } else if (ph == 2) {
    w.x = -v.x;
    w.y = -v.y;
  } else {
    w.x = -v.y;
    w.y = v.x;
  }
  return w;
}

// y += t * (u + v) on an (re, im) pair. For float2 this is one FADD2 + one FFMA2 (Blackwell's
// packed FP32 pipe: both lanes of the pair in one issue slot, the tap as a broadcast scalar).
__device__ __forceinline__ float2 hb_acc(float2 y, float t, float2 u, float2 v) {
  return __ffma2_rn(make_float2(t, t), __fadd2_rn(u, v), y);
}
__device__ __forceinline__ double2 hb_acc(double2 y, double t, double2 u, double2 v) {
  y.x += t * (u.x + v.x);
  y.y += t * (u.y + v.y);
  return y;
}

// One half-band stage with N taps: outputs [lo_out, lo_out+len_out) from the source level's
// E/O arrays (base index ebp = lo_out - N), written to the next level's arrays or to `out`.
template <typename S, int N, bool LAST>
__device__ __forceinline__ void hb_stage(const typename V2<S>::type *__restrict__ Es,
                                         const typename V2<S>::type *__restrict__ Os, int sl_src,
                                         typename V2<S>::type *En, typename V2<S>::type *On, int sl_dst,
                                         int64_t ebn, const S *__restrict__ t, int64_t lo_out, int len_out,
                                         Ring<typename V2<S>::type> out, uint32_t c) {
  using V = typename V2<S>::type;
  constexpr int W = kHbR + 2 * N - 1; // odd-neighbour window of kHbR consecutive outputs
  const int ngroups = (len_out + kHbR - 1) / kHbR;
  for (int g = threadIdx.x; g < ngroups; g += kHbThreads) {
    // local source index of output r's centre: q = kHbR*g + r + N
    V w[W], e[kHbR];
#pragma unroll
    for (int sidx = 0; sidx < W; sidx++) w[sidx] = Os[(sidx % kHbR) * sl_src + g + sidx / kHbR];
#pragma unroll
    for (int r = 0; r < kHbR; r++) e[r] = Es[((r + N) % kHbR) * sl_src + g + (r + N) / kHbR];
#pragma unroll
    for (int r = 0; r < kHbR; r++) {
      V y = e[r];
#pragma unroll
      for (int k = 0; k < N; k++) {
        y = hb_acc(y, t[k], w[r + N + k], w[r + N - k - 1]);
      }
      const int i = kHbR * g + r;
      if (i < len_out) {
        const int64_t m = lo_out + i;
        if (m < 0) {
          y.x = 0;
          y.y = 0;
        }
        if (LAST) {
          out.st(c, m, y);
        } else {
          const int j = (int)((m >> 1) - ebn);
          if (m & 1) {
            On[hb_pos(j, sl_dst)] = y;
          } else {
            En[hb_pos(j, sl_dst)] = y;
          }
        }
      }
    }
  }
}

// The tap counts are template parameters (N1, N2, N3; 0 = stage absent) so that every stage's
// register window has a compile-time size; the host instantiates the combinations r8brain
// produces for the shipped rates.
template <typename S, int NST, bool LINEAR, int N1, int N2, int N3>
__global__ void __launch_bounds__(kHbThreads)
    k_hb_cascade(InSrc<typename V2<S>::type> in, Ring<typename V2<S>::type> out, HbTaps<S> taps,
                 int64_t o0, int n_out, int fs4) {
  using V = typename V2<S>::type;
  constexpr int kTile = hb_tile_of(N1, N2);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t c = blockIdx.y;
  const int64_t a_fin = o0 + (int64_t)blockIdx.x * kTile;
  int cnt_fin = n_out - (int)(blockIdx.x * kTile);
  if (cnt_fin > kTile) cnt_fin = kTile;
  if (cnt_fin <= 0) return;

  if (NST == 0) { // plain copy (+ Fs/4 shift)
    for (int i = threadIdx.x; i < cnt_fin; i += kHbThreads) {
      const int64_t ai = a_fin + i;
      V v = src_ld<V, LINEAR>(in, c, ai);
      if (fs4) v = fs4_rot(v, ai);
      out.st(c, ai, v);
    }
    return;
  }
  // index ranges per level: level NST = final outputs, level 0 = raw input
  int64_t lo[4];
  int len[4];
  lo[NST] = a_fin;
  len[NST] = cnt_fin;
#pragma unroll
  for (int s = NST; s >= 1; s--) {
    const int h = 2 * taps.n[s - 1] - 1;
    lo[s - 1] = 2 * lo[s] - h;
    len[s - 1] = 2 * len[s] + 2 * h - 1;
  }
  V *E[3], *O[3];
  int sl[3];
  {
    V *p = reinterpret_cast<V *>(smem_raw);
#pragma unroll
    for (int s = 0; s < NST; s++) {
      sl[s] = taps.sl[s];
      // level 2 re-uses level 0's storage: level 0 is dead once stage 1 has run
      V *q = (s == 2) ? reinterpret_cast<V *>(smem_raw) : p;
      E[s] = q;
      O[s] = q + kHbR * sl[s];
      if (s < 2) p += 2 * kHbR * sl[s];
    }
  }
  // ---- level 0: pairs (x[2p], x[2p+1]) -> E0[p - eb], O0[p - eb]
  {
    const int64_t eb = lo[0] >> 1; // floor
    const int npairs = (int)(((lo[0] + len[0] - 1) >> 1) - eb) + 1 + kHbR; // a few extra: window overrun
    constexpr int kN1 = (N1 > 0 ? N1 : 1), kN2 = (N2 > 0 ? N2 : 0), kN3 = (N3 > 0 ? N3 : 0);
    // level-0 samples of a full tile (compile time) -> loads per thread
    constexpr int kLen0 = (NST == 1)   ? (2 * kTile + 4 * kN1 - 3)
                          : (NST == 2) ? (2 * (2 * kTile + 4 * kN2 - 3) + 4 * kN1 - 3)
                                       : (2 * (2 * (2 * kTile + 4 * kN3 - 3) + 4 * kN2 - 3) + 4 * kN1 - 3);
    constexpr int kMaxPairs = kLen0 / 2 + 2 + kHbR;
    constexpr int kBatch = (kMaxPairs + kHbThreads - 1) / kHbThreads;
    bool fast = false;
    if (LINEAR && sizeof(V) == 8) {
      const int64_t r0 = 2 * eb - in.start;
      const float2 *pp = reinterpret_cast<const float2 *>(in.lin) + (size_t)c * in.stride + r0;
      const short2 *ps = reinterpret_cast<const short2 *>(in.lin) + (size_t)c * in.stride + r0;
      const bool inside = (r0 >= 0) && (r0 + 2 * (int64_t)npairs <= in.n_new);
      if (in.fmt == 1 && inside && ((reinterpret_cast<uintptr_t>(ps) & 7) == 0)) {
        // int16 IQ: one 64-bit load per pair of samples, converted on the fly
        fast = true;
        const uint2 *p2 = reinterpret_cast<const uint2 *>(ps);
        uint2 q[kBatch];
#pragma unroll
        for (int bb = 0; bb < kBatch; bb++) {
          const int i = threadIdx.x + bb * kHbThreads;
          q[bb] = (i < npairs) ? p2[i] : make_uint2(0u, 0u);
        }
        const int ph0 = (int)((2 * eb) & 3);
#pragma unroll
        for (int bb = 0; bb < kBatch; bb++) {
          const int i = threadIdx.x + bb * kHbThreads;
          if (i < npairs && i / kHbR < sl[0]) {
            const float k = 1.0f / 32768.0f;
            float2 v0 = make_float2((float)(short)(q[bb].x & 0xffff) * k, (float)(short)(q[bb].x >> 16) * k);
            float2 v1 = make_float2((float)(short)(q[bb].y & 0xffff) * k, (float)(short)(q[bb].y >> 16) * k);
            if (fs4) {
              const int ph = (ph0 + 2 * i) & 3;
              v0 = fs4_rot(v0, ph);
              v1 = fs4_rot(v1, ph + 1);
            }
            const int pos = hb_pos(i, sl[0]);
            reinterpret_cast<float2 *>(E[0])[pos] = v0;
            reinterpret_cast<float2 *>(O[0])[pos] = v1;
          }
        }
      }
      if (!fast && in.fmt == 0) {
        fast = inside && ((reinterpret_cast<uintptr_t>(pp) & 15) == 0);
      }
      if (fast && in.fmt == 0) {
        // whole tile inside this call's buffer and 16-byte aligned: all 128-bit loads of a
        // thread are issued back to back (latency overlapped), then scattered to shared memory
        const float4 *p4 = reinterpret_cast<const float4 *>(pp);
        float4 q[kBatch];
#pragma unroll
        for (int bb = 0; bb < kBatch; bb++) {
          const int i = threadIdx.x + bb * kHbThreads;
          q[bb] = (i < npairs) ? p4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int ph0 = (int)((2 * eb) & 3); // Fs/4 phase of pair 0's even sample (0 or 2)
#pragma unroll
        for (int bb = 0; bb < kBatch; bb++) {
          const int i = threadIdx.x + bb * kHbThreads;
          if (i < npairs && i / kHbR < sl[0]) {
            float2 v0 = make_float2(q[bb].x, q[bb].y), v1 = make_float2(q[bb].z, q[bb].w);
            if (fs4) {
              const int ph = (ph0 + 2 * i) & 3;
              v0 = fs4_rot(v0, ph);
              v1 = fs4_rot(v1, ph + 1);
            }
            const int pos = hb_pos(i, sl[0]);
            reinterpret_cast<float2 *>(E[0])[pos] = v0;
            reinterpret_cast<float2 *>(O[0])[pos] = v1;
          }
        }
      }
    }
    if (!fast) {
      // generic source (ring, history, int16 at an odd offset, FP64): still all loads of a batch before the first
      // store, so a thread pays one memory round trip per kGen pairs instead of one per pair
      constexpr int kGen = 4;
      for (int i0 = threadIdx.x; i0 < npairs; i0 += kGen * kHbThreads) {
        V v0[kGen], v1[kGen];
#pragma unroll
        for (int q = 0; q < kGen; q++) {
          const int i = i0 + q * kHbThreads;
          const int64_t a0 = 2 * (eb + i);
          if (i < npairs) {
            v0[q] = src_ld<V, LINEAR>(in, c, a0);
            v1[q] = src_ld<V, LINEAR>(in, c, a0 + 1);
          }
        }
#pragma unroll
        for (int q = 0; q < kGen; q++) {
          const int i = i0 + q * kHbThreads;
          const int64_t a0 = 2 * (eb + i);
          if (i < npairs && i / kHbR < sl[0]) {
            V w0 = v0[q], w1 = v1[q];
            if (fs4) {
              w0 = fs4_rot(w0, a0);
              w1 = fs4_rot(w1, a0 + 1);
            }
            E[0][hb_pos(i, sl[0])] = w0;
            O[0][hb_pos(i, sl[0])] = w1;
          }
        }
      }
    }
  }
  __syncthreads();
  if (NST == 1) {
    hb_stage<S, (N1 > 0 ? N1 : 1), true>(E[0], O[0], sl[0], nullptr, nullptr, 0, 0, taps.t[0], lo[1], len[1], out, c);
  } else {
    hb_stage<S, (N1 > 0 ? N1 : 1), false>(E[0], O[0], sl[0], E[1], O[1], sl[1], lo[1] >> 1, taps.t[0], lo[1], len[1],
                                          out, c);
    __syncthreads();
    if (NST == 2) {
      hb_stage<S, (N2 > 0 ? N2 : 1), true>(E[1], O[1], sl[1], nullptr, nullptr, 0, 0, taps.t[1], lo[2], len[2], out,
                                           c);
    } else {
      hb_stage<S, (N2 > 0 ? N2 : 1), false>(E[1], O[1], sl[1], E[2], O[2], sl[2], lo[2] >> 1, taps.t[1], lo[2],
                                            len[2], out, c);
      __syncthreads();
      hb_stage<S, (N3 > 0 ? N3 : 1), true>(E[2], O[2], sl[2], nullptr, nullptr, 0, 0, taps.t[2], lo[3], len[3], out,
                                           c);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Long zero-phase low-pass, direct form (reference: r8b::CDSPBlockConvolver::process,
// CDSPBlockConvolver.h:252-353 — overlap-save FFT there; same linear convolution here):
//   y[q] = sum_j h[j] * x[q*down - fl2 + j],  j = 0..klen-1,  fl2 = (klen-1)/2.
constexpr int kFirThreads = 256;
constexpr int kFirR = 4;
constexpr int kFirTile = kFirThreads * kFirR;

template <typename S>
__global__ void __launch_bounds__(kFirThreads)
    k_fir_long(Ring<typename V2<S>::type> in, Ring<typename V2<S>::type> out, const S *__restrict__ taps,
               int klen, int down, int64_t q0, int n_out) {
  using V = typename V2<S>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t c = blockIdx.y;
  const int tile0 = blockIdx.x * kFirTile;
  int cnt = n_out - tile0;
  if (cnt > kFirTile) cnt = kFirTile;
  if (cnt <= 0) return;
  const int span = (kFirTile - 1) * down + klen;
  V *xs = reinterpret_cast<V *>(smem_raw);
  S *hs = reinterpret_cast<S *>(xs + span);
  const int fl2 = (klen - 1) / 2;
  const int64_t x0 = (q0 + tile0) * down - fl2;
  const int need = (cnt - 1) * down + klen;
  for (int i = threadIdx.x; i < span; i += kFirThreads) {
    V v;
    v.x = 0;
    v.y = 0;
    if (i < need) v = in.ld(c, x0 + i);
    xs[i] = v;
  }
  for (int i = threadIdx.x; i < klen; i += kFirThreads) hs[i] = taps[i];
  __syncthreads();
  V acc[kFirR];
#pragma unroll
  for (int r = 0; r < kFirR; r++) {
    acc[r].x = 0;
    acc[r].y = 0;
  }
  const int base = threadIdx.x * down;
  const int rstep = kFirThreads * down;
  if ((int)threadIdx.x >= cnt) return; // a short tile (the few outputs redone at stream start): no output in any round
  for (int k = 0; k < klen; k++) {
    const S h = hs[k];
#pragma unroll
    for (int r = 0; r < kFirR; r++) {
      const V x = xs[base + r * rstep + k];
      acc[r].x += h * x.x;
      acc[r].y += h * x.y;
    }
  }
#pragma unroll
  for (int r = 0; r < kFirR; r++) {
    const int q = threadIdx.x + r * kFirThreads;
    if (q < cnt) out.st(c, q0 + tile0 + q, acc[r]);
  }
}

// ---------------------------------------------------------------------------------------
// Double-precision decimate-by-2 low-pass for the audio resamplers (reference:
// r8b::CDSPBlockConvolver with DownFactor 2 inside AudioResampler, AudioResampler.cpp:37-61;
// 1621 taps at 96 kHz, both audio lanes (mono, L-R) ride as one double2 stream).
// Polyphase form: y[u] = sum_i h0[i] X0[u+i] + sum_i h1[i] X1[u+i] with X0/X1 the even/odd
// input samples, so every output costs klen FMAs per lane instead of 2*klen. Each thread
// owns kDecR consecutive outputs and keeps a sliding window of inputs in registers (one new
// shared-memory load per tap for 2*kDecR FMAs); the input phases are stored kDecR-way
// interleaved so that the window loads of a warp are contiguous (bank-conflict free).
constexpr int kDecR = 8;
constexpr int kDecThreads = 64;
constexpr int kDecTile = kDecR * kDecThreads; // outputs per CTA
constexpr int kDecMaxTaps = 1024;             // per phase, padded to a multiple of kDecR

__host__ __device__ inline int dec2_len(int ntp) { return kDecThreads + ntp / kDecR + 3; }
__host__ inline size_t dec2_smem(int klen) {
  const int ntp = ((klen + 1) / 2 + kDecR - 1) / kDecR * kDecR;
  return (size_t)2 * kDecR * dec2_len(ntp) * sizeof(double2) + (size_t)2 * ntp * sizeof(double);
}

static __global__ void __launch_bounds__(kDecThreads)
    k_fir_dec2_f64(Ring<double2> in, Ring<double2> out, const double *__restrict__ taps, int klen, int64_t q0,
                   int n_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t c = blockIdx.y;
  const int tile0 = blockIdx.x * kDecTile;
  int cnt = n_out - tile0;
  if (cnt > kDecTile) cnt = kDecTile;
  if (cnt <= 0) return;
  const int nt0 = (klen + 1) / 2;                       // taps of the even phase
  const int ntp = (nt0 + kDecR - 1) / kDecR * kDecR;    // padded taps per phase
  const int LEN = dec2_len(ntp);
  double2 *xs = reinterpret_cast<double2 *>(smem_raw);  // [2][kDecR][LEN]
  double *hs = reinterpret_cast<double *>(xs + 2 * kDecR * LEN); // [2][ntp]
  const int fl2 = (klen - 1) / 2;
  const int64_t B = 2 * (q0 + tile0) - fl2;
  const int span = 2 * (kDecTile + ntp + kDecR);        // X[0..span)
  const int need = 2 * (cnt - 1) + klen;
  for (int i = threadIdx.x; i < span; i += kDecThreads) {
    double2 v = make_double2(0.0, 0.0);
    if (i < need) v = in.ld(c, B + i);
    const int rho = i & 1, m = i >> 1;
    if (m / kDecR < LEN) xs[(rho * kDecR + (m % kDecR)) * LEN + m / kDecR] = v;
  }
  for (int i = threadIdx.x; i < 2 * ntp; i += kDecThreads) {
    const int rho = i / ntp, k = i - rho * ntp;
    const int j = 2 * k + rho;
    hs[i] = (j < klen) ? taps[j] : 0.0;
  }
  __syncthreads();
  double2 acc[kDecR];
#pragma unroll
  for (int r = 0; r < kDecR; r++) acc[r] = make_double2(0.0, 0.0);
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int rho = 0; rho < 2; rho++) {
    const double2 *xp = xs + (size_t)rho * kDecR * LEN;
    const double *hp = hs + rho * ntp;
    double2 w[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; r++) w[r] = xp[r * LEN + tid];
    for (int i0 = 0; i0 < ntp; i0 += kDecR) {
#pragma unroll
      for (int ii = 0; ii < kDecR; ii++) {
        const double h = hp[i0 + ii];
#pragma unroll
        for (int r = 0; r < kDecR; r++) {
          const double2 x = w[(r + ii) % kDecR];
          acc[r].x = fma(h, x.x, acc[r].x);
          acc[r].y = fma(h, x.y, acc[r].y);
        }
        w[ii] = xp[ii * LEN + tid + i0 / kDecR + 1];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kDecR; r++) {
    const int u = kDecR * tid + r;
    if (u < cnt) out.st(c, q0 + tile0 + u, acc[r]);
  }
}

// ---------------------------------------------------------------------------------------
// Whole-step polyphase interpolator (reference: r8b::CDSPFracInterpolator::convolve0,
// CDSPFracInterpolator.h:992-1060): output m reads flen inputs starting at
// floor(m*instep/outstep) - (flen/2 - 1) with the bank row (m*instep) mod outstep.
constexpr int kFiThreads = 128;
constexpr int kFiPer = 4; // outputs per thread (strided), so one copy of the bank serves 512 outputs
constexpr int kFiTile = kFiThreads * kFiPer;

// smem: input window of the tile + the whole bank, rows padded to an odd length
__host__ __device__ inline int fi_row(int flen) { return flen | 1; }
__host__ inline size_t fi_smem(int instep, int outstep, int flen, size_t vbytes, size_t sbytes) {
  const int span = (int)(((int64_t)kFiTile * instep) / outstep) + flen + 4;
  return (size_t)span * vbytes + (size_t)outstep * fi_row(flen) * sbytes + 16;
}

template <typename S>
__global__ void __launch_bounds__(kFiThreads)
    k_frac_interp(Ring<typename V2<S>::type> in, Ring<typename V2<S>::type> out, const S *__restrict__ bank, int instep,
                  int outstep, int flen, int64_t m0, int n_out) {
  using V = typename V2<S>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t c = blockIdx.y;
  const int tile0 = blockIdx.x * kFiTile;
  int cnt = n_out - tile0;
  if (cnt > kFiTile) cnt = kFiTile;
  if (cnt <= 0) return;
  const int span = (int)(((int64_t)kFiTile * instep) / outstep) + flen + 4;
  V *xs = reinterpret_cast<V *>(smem_raw);
  S *bs = reinterpret_cast<S *>(xs + span);
  const int rowp = fi_row(flen);
  const int64_t mt = m0 + tile0;
  const int64_t xbase = (mt * instep) / outstep - (flen / 2 - 1); // first input the tile reads
  for (int i = threadIdx.x; i < span; i += kFiThreads) xs[i] = in.ld(c, xbase + i);
  for (int i = threadIdx.x; i < outstep * flen; i += kFiThreads) {
    const int ph = i / flen, k = i - ph * flen;
    bs[ph * rowp + k] = bank[i];
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < kFiPer; rr++) {
    const int i = threadIdx.x + rr * kFiThreads;
    if (i >= cnt) break;
    const int64_t m = mt + i;
    const int64_t pos = m * instep;
    const int64_t ip = pos / outstep;
    const int ph = (int)(pos - ip * outstep);
    const S *row = bs + ph * rowp;
    const int x0 = (int)(ip - (flen / 2 - 1) - xbase);
    V acc;
    acc.x = 0;
    acc.y = 0;
    for (int k = 0; k < flen; k++) {
      const V x = xs[x0 + k];
      const S h = row[k];
      acc.x += h * x.x;
      acc.y += h * x.y;
    }
    out.st(c, m, acc);
  }
}

// ---------------------------------------------------------------------------------------
// Short symmetric FIR with the reference's per-call head-loop quirk (reference:
// LowPassFilterFirIQ::process Filter.cpp:37-96 and LowPassFilterFirAudio::process
// Filter.cpp:108-163): the first min(n, order) outputs of every process() call are
// computed by a loop that starts at coefficient 1, i.e. the coeff[0]*x[p] term is never
// added for them. call_end[] holds the cumulative per-call ends
// (relative to j0) of this launch's range.
__device__ __forceinline__ int find_call(const uint32_t *__restrict__ call_end, int n_calls, uint32_t rel) {
  int lo = 0, hi = n_calls - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (call_end[mid] > rel) {
      hi = mid;
    } else {
      lo = mid + 1;
    }
  }
  return lo;
}

// The head-loop quirk as a correction of a full convolution (used behind the FFT form of the long channel filters,
// fmr_am.cu): y[i] -= coeff[0] * x[i] for the outputs that fall into the first `order` samples of their reference call.
template <typename S>
__global__ void k_fir_head_fix(Ring<typename V2<S>::type> x, Ring<typename V2<S>::type> y, S c0, int order, int64_t j0, int n_out,
                               const uint32_t *__restrict__ call_end, int n_calls) {
  using V = typename V2<S>::type;
  const uint32_t c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int b = find_call(call_end, n_calls, (uint32_t)i);
  const uint32_t cstart = (b == 0) ? 0u : call_end[b - 1];
  if (i - (int)cstart < order) {
    const V xv = x.ld(c, j0 + i);
    V yv = y.ld(c, j0 + i);
    yv.x -= c0 * xv.x;
    yv.y -= c0 * xv.y;
    y.st(c, j0 + i, yv);
  }
}

// acc += h * x on an (re, im) pair: one packed FFMA2 for float2 (the same rounding per lane as two FFMA), two DFMA for double2
__device__ __forceinline__ float2 fir_mac(float2 acc, float h, float2 x) { return __ffma2_rn(make_float2(h, h), x, acc); }
__device__ __forceinline__ double2 fir_mac(double2 acc, double h, double2 x) {
  acc.x += h * x.x;
  acc.y += h * x.y;
  return acc;
}

constexpr int kQR = 4;        // consecutive outputs per thread (8 measured: no gain in FP32 or FP64, profiles/README.md)
constexpr int kQThreads = 64; // tile = 256 outputs
constexpr int kQTile = kQR * kQThreads;

__host__ __device__ inline int fq_len(int ntp) { return kQThreads + ntp / kQR + 3; }
__host__ inline size_t fq_smem(int ntaps, size_t vbytes, size_t sbytes) {
  const int ntp = (ntaps + kQR - 1) / kQR * kQR;
  return (size_t)kQR * fq_len(ntp) * vbytes + (size_t)ntp * sbytes + 16;
}

// Register-blocked form: every thread owns kQR consecutive outputs and slides a window of
// inputs through registers (one shared-memory load per tap for 2*kQR FMAs); inputs are stored
// kQR-way interleaved so the window loads of a warp are contiguous. All taps are accumulated;
// the coeff[0]*x[p] term is then taken out again for the outputs that fall into the head region
// of their reference call (same value as never adding it, up to one float rounding).
template <typename S>
__global__ void __launch_bounds__(kQThreads)
    k_fir_quirk(Ring<typename V2<S>::type> in, Ring<typename V2<S>::type> out, const S *__restrict__ coeff,
                int ntaps, int64_t j0, int n_out, const uint32_t *__restrict__ call_end, int n_calls) {
  using V = typename V2<S>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t c = blockIdx.y;
  const int tile0 = blockIdx.x * kQTile;
  int cnt = n_out - tile0;
  if (cnt > kQTile) cnt = kQTile;
  if (cnt <= 0) return;
  const int order = ntaps - 1;
  const int ntp = (ntaps + kQR - 1) / kQR * kQR;
  const int LEN = fq_len(ntp);
  // The tile is stored REVERSED in time and shifted by one: Z[i] = x[E - 1 - i] with E the last
  // output index of a full tile; thread outputs are v = kQTile-1-u, so y_v = sum_{k>=1} coeff[k] *
  // Z[v + k - 1] walks the taps in increasing k — the order of the reference's head loop
  // (Filter.cpp:59-68). Tap 0 is NOT part of the loop: like the reference it is added only for
  // outputs outside the head region, so a head output never sees x[p] at all (exact zeros at
  // stream start stay exact zeros, which the discriminator's atan2(0,0) depends on).
  V *xs = reinterpret_cast<V *>(smem_raw); // [kQR][LEN]
  S *hs = reinterpret_cast<S *>(xs + kQR * LEN); // hs[k'] = coeff[k' + 1]
  // The reference call of this thread's oldest output (the others follow by stepping through call_end). The
  // bisection is a chain of dependent loads: it runs before the tile's loads are issued, so that each step waits
  // for a cache hit and not for the memory round trips queued behind it.
  const int tid = threadIdx.x;
  int bcall = find_call(call_end, n_calls, (uint32_t)(tile0 + kQTile - kQR * (tid + 1)));
  uint32_t cend = call_end[bcall], cstart = (bcall == 0) ? 0u : call_end[bcall - 1];
  const int64_t E = j0 + tile0 + kQTile - 2;
  const int span = kQTile + ntp + kQR;
  const int64_t last = j0 + tile0 + cnt - 1; // newest sample this tile may read
  // the loads of a thread go out back to back (one memory round trip per kFill * kQThreads samples), then the stores
  constexpr int kFill = 8;
  for (int i0 = threadIdx.x; i0 < span; i0 += kFill * kQThreads) {
    V v[kFill];
#pragma unroll
    for (int q = 0; q < kFill; q++) {
      const int i = i0 + q * kQThreads;
      v[q].x = 0;
      v[q].y = 0;
      const int64_t xi = E - i;
      if (i < span && xi <= last && i < kQTile + order) v[q] = in.ld(c, xi);
    }
#pragma unroll
    for (int q = 0; q < kFill; q++) {
      const int i = i0 + q * kQThreads;
      if (i < span && i / kQR < LEN) xs[(i % kQR) * LEN + i / kQR] = v[q];
    }
  }
  for (int i = threadIdx.x; i < ntp; i += kQThreads) hs[i] = (i < order) ? coeff[i + 1] : (S)0;
  // x[p] of every output for the coeff[0] term, fetched while the tile is in flight
  V xp[kQR];
#pragma unroll
  for (int r = 0; r < kQR; r++) {
    const int u = kQTile - 1 - (kQR * tid + r);
    xp[r].x = 0;
    xp[r].y = 0;
    if (u < cnt) xp[r] = in.ld(c, j0 + tile0 + u);
  }
  __syncthreads();
  V acc[kQR], w[kQR];
#pragma unroll
  for (int r = 0; r < kQR; r++) {
    acc[r].x = 0;
    acc[r].y = 0;
    w[r] = xs[r * LEN + tid];
  }
  // y[v] = sum_k' hs[k'] * Z[v + k'], v = kQR*tid + r
  for (int q0 = 0; q0 < ntp; q0 += kQR) {
#pragma unroll
    for (int qq = 0; qq < kQR; qq++) {
      const S h = hs[q0 + qq];
#pragma unroll
      for (int r = 0; r < kQR; r++) acc[r] = fir_mac(acc[r], h, w[(r + qq) % kQR]);
      w[qq] = xs[qq * LEN + tid + q0 / kQR + 1];
    }
  }
  const S c0 = coeff[0];
#pragma unroll
  for (int r = kQR - 1; r >= 0; r--) { // oldest output first: the call index only moves forward
    const int v = kQR * tid + r;
    const int u = kQTile - 1 - v;
    const int i = tile0 + u; // index within this launch
    if (u < cnt) {
      while ((uint32_t)i >= cend && bcall < n_calls - 1) {
        cstart = cend;
        cend = call_end[++bcall];
      }
      V y = acc[r];
      if (i - (int)cstart >= order) { // outside the head loop of the reference: coeff[0]*x[p] is added
        y.x += c0 * xp[r].x;
        y.y += c0 * xp[r].y;
      }
      out.st(c, j0 + i, y);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Per-channel persistent state of the FM decoder (reference: FmDecoder members,
// include/FmDecode.h:127-163, and the members of the blocks it owns).
struct FmChanState {
  // IfSimpleAgc (IfSimpleAgc.h:55-60)
  float agc_gain;
  // PhaseDiscriminator m_save_value (PhaseDiscriminator.h:44-48)
  float disc_prev;
  // FmDecoder statistics (FmDecode.h:134-137)
  float baseband_mean, baseband_level, if_rms;
  int stereo_detected;
  uint32_t mpf_wait; // m_wait_multipath_blocks
  int lock_cnt;      // PilotPhaseLock m_lock_cnt
  int pilot_periods;
  uint32_t n_pps; // events recorded in the last process call
  // PilotPhaseLock (PilotPhaseLock.h:81-95)
  double pll_phase, pll_freq;
  double bi_x1, bi_x2, bq_x1, bq_x2; // biquad delay lines (I, Q)
  double lf_x1;                      // loop filter delay
  double pilot_level;                // m_pilot_level (not doubled)
  double freq_err;
  unsigned long long pps_cnt, sample_cnt;
  // LowPassFilterRC delay lines (mono, L-R)
  double de_m_x1, de_s_x1;
  // HighPassFilterIir delay lines (mono, L-R)
  double dc_m_x1, dc_m_x2, dc_s_x1, dc_s_x2;
  // MultipathFilter m_error
  double mpf_error;
  unsigned long long decoder_calls;
};

struct PpsEventDev {
  unsigned long long pps_index, sample_index;
  double block_position;
  uint32_t block;
  uint32_t pad;
};
constexpr int kMaxPps = 16;

struct FmCoreParams {
  // constants
  float agc_max, agc_rate;         // IfSimpleAgc(1.0, 100000.0, 0.0001)  FmDecode.cpp:74
  float disc_inv_norm, disc_bound; // PhaseDiscriminator.cpp:27-30
  double pll_minfreq, pll_maxfreq; // PilotPhaseLock.cpp:35-36
  double bq_b0, bq_a1, bq_a2;      // PilotPhaseLock.cpp:48-49
  double lf_b0, lf_b1;             // PilotPhaseLock.cpp:51
  int lock_delay;                  // PilotPhaseLock.cpp:43
  double minsignal;                // PilotPhaseLock.h:37
  double de_a1, de_b0;             // LowPassFilterRC  Filter.cpp:186-188
  int stereo, pilot_shift, deemph_on_stereo;
  int n_channels;
};

// The 384 kHz core (reference: FmDecoder::process FmDecode.cpp:85-183 up to the audio
// resamplers) is split by data dependence into three launches:
//   k_fm_agc2 — IfSimpleAgc, a float recurrence of a handful of instructions per sample,
//               one lane per channel;
//   k_fm_disc — everything between the recurrences that is parallel in time: phase
//               discriminator, and the per-call statistics (IF RMS, baseband mean/RMS);
//   k_fm_pll2 — PilotPhaseLock + L-R mix + both deemphasis filters, one lane per channel.
// With the multipath filter enabled k_mpf runs between k_fm_agc2 and k_fm_disc.
constexpr int kCoreChunk = 8;

// Branch-free atan2f for the phase discriminator of the fused core (volk_32fc_s32f_atan2_32f in the
// reference, whose accuracy depends on the VOLK version and machine: generic = libm atan2f, newer
// SIMD kernels = polynomial). CUDA's atan2f is ~75 instructions with branches and a checked
// division; in a one-warp pipeline stage that was ~380 cycles per sample and made the
// discriminator, not the PLL, the bottleneck. This form is 27 straight-line instructions:
// quotient min/max by reciprocal + Newton + exact-residual correction, atan(z)/z on [0,1] by the
// 8-term polynomial of Abramowitz & Stegun 4.4.49 (|err| <= 2e-8), octant fix-up by selects.
// Measured against double atan2 on 3e7 inputs (tests/cpp/fast_atan2_host_test.cpp): max 2.8e-7 rad
// (2.8 ulp), rms 7e-8 rad — the same class as CUDA's atan2f (3 ulp).
__device__ __forceinline__ float fmr_atan2f(float y, float x) {
  const float ya = fabsf(y), xa = fabsf(x);
  const float mn = fminf(ya, xa), mx = fmaxf(ya, xa);
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(mx));
  const float r1 = fmaf(fmaf(-mx, r0, 1.0f), r0, r0);
  float z = mn * r1;
  z = fmaf(fmaf(-mx, z, mn), r1, z);
  const float s = z * z;
  float p = 0.0028662257f;
  p = fmaf(p, s, -0.0161657367f);
  p = fmaf(p, s, 0.0429096138f);
  p = fmaf(p, s, -0.0752896400f);
  p = fmaf(p, s, 0.1065626393f);
  p = fmaf(p, s, -0.1420889944f);
  p = fmaf(p, s, 0.1999355085f);
  p = fmaf(p, s, -0.3333314528f);
  float r = fmaf(p * s, z, z);
  r = (ya > xa) ? 1.57079632679489661923f - r : r;
  r = (x < 0.0f) ? 3.14159265358979323846f - r : r;
  r = (mx > 0.0f) ? r : 0.0f; // atan2f(+-0, +-0): the discriminator needs exact zeros at stream start
  return copysignf(r, y);
}

// k_fm_agc2 — IfSimpleAgc (IfSimpleAgc.cpp:33-60) with the bookkeeping taken out of the loop (row
// pointers and ring masks hoisted, no sign checks on the absolute index): the kernel is bound by
// dependent-issue latency of ONE warp per SM sub-partition, so every instruction that is not the
// recurrence costs as much as one that is.
static __global__ void __launch_bounds__(32)
    k_fm_agc2(Ring<float2> iq_in, Ring<float2> iq_out, FmChanState *__restrict__ st, int n_total, int64_t t0,
              FmCoreParams P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_channels) return;
  float g = st[c].agc_gain;
  const float2 *__restrict__ irow = iq_in.base + (size_t)c * iq_in.cap;
  float2 *__restrict__ orow = iq_out.base + (size_t)c * iq_out.cap;
  const uint32_t imask = iq_in.cap - 1, omask = iq_out.cap - 1, t0lo = (uint32_t)t0;
  const double rate = (double)P.agc_rate;
  const float gmax = P.agc_max;
  float2 nxt[kCoreChunk];
#pragma unroll
  for (int u = 0; u < kCoreChunk; u++) nxt[u] = (u < n_total) ? irow[(t0lo + (uint32_t)u) & imask] : make_float2(0.f, 0.f);
  for (int i0 = 0; i0 < n_total; i0 += kCoreChunk) {
    float2 xin[kCoreChunk];
#pragma unroll
    for (int u = 0; u < kCoreChunk; u++) xin[u] = nxt[u];
#pragma unroll
    for (int u = 0; u < kCoreChunk; u++) {
      const int i = i0 + kCoreChunk + u;
      nxt[u] = (i < n_total) ? irow[(t0lo + (uint32_t)i) & imask] : make_float2(0.f, 0.f);
    }
    if (i0 + kCoreChunk <= n_total) {
#pragma unroll
      for (int u = 0; u < kCoreChunk; u++) {
        // IfSimpleAgc::process (IfSimpleAgc.cpp:37-57)
        float2 x2;
        x2.x = xin[u].x * g;
        x2.y = xin[u].y * g;
        const float nrm = x2.x * x2.x + x2.y * x2.y;
        const float z = (float)(1.0 + (rate * (1.0 - (double)nrm)));
        g *= z;
        g = isfinite(g) ? ((g > gmax) ? gmax : g) : 1.0f;
        orow[(t0lo + (uint32_t)(i0 + u)) & omask] = x2;
      }
    } else {
      for (int u = 0; u < kCoreChunk && i0 + u < n_total; u++) {
        float2 x2;
        x2.x = xin[u].x * g;
        x2.y = xin[u].y * g;
        const float nrm = x2.x * x2.x + x2.y * x2.y;
        const float z = (float)(1.0 + (rate * (1.0 - (double)nrm)));
        g *= z;
        g = isfinite(g) ? ((g > gmax) ? gmax : g) : 1.0f;
        orow[(t0lo + (uint32_t)(i0 + u)) & omask] = x2;
      }
    }
  }
  st[c].agc_gain = g;
}

// One thread per 384 kHz sample: PhaseDiscriminator::process (PhaseDiscriminator.cpp:33-46).
// The previous sample's phase is recomputed from the ring (bit-identical to the value the
// serial loop would have carried); before the very first sample m_save_value is 0.
static __global__ void k_fm_disc(Ring<float2> iq, Ring<float> mpx, int n_total, int64_t t0, FmCoreParams P) {
  const uint32_t c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  const int64_t t = t0 + i;
  const float2 x = iq.base[(size_t)c * iq.cap + ((uint32_t)t & (iq.cap - 1))];
  const float ph = atan2f(x.y, x.x) * P.disc_inv_norm;
  float prev = 0.0f;
  if (t > 0) {
    const float2 xp = iq.base[(size_t)c * iq.cap + ((uint32_t)(t - 1) & (iq.cap - 1))];
    prev = atan2f(xp.y, xp.x) * P.disc_inv_norm;
  }
  float d = ph - prev;
  if (d > P.disc_bound) d -= 2 * P.disc_bound;
  if (d < -P.disc_bound) d += 2 * P.disc_bound;
  if (isnan(d)) d = 0.0f;
  mpx.base[(size_t)c * mpx.cap + ((uint32_t)t & (mpx.cap - 1))] = d;
}

// One warp per (call, channel): Utility::rms_level_sample on the decoder input
// (FmDecode.cpp:95, Utility.h:118-132) and Utility::samples_mean_rms on the MPX
// (FmDecode.cpp:146, Utility.h:135-152). stats[(c*n_calls + b)*3 + {0,1,2}] = if_rms, mean, rms.
static __global__ void k_fm_call_stats(Ring<float2> if_raw, Ring<float> mpx, float *__restrict__ stats,
                                const uint32_t *__restrict__ call_end, int n_calls, int64_t t0) {
  const int warps_per_block = blockDim.x >> 5;
  const int b = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const uint32_t c = blockIdx.y;
  if (b >= n_calls) return;
  const int lane = threadIdx.x & 31;
  const uint32_t beg = b ? call_end[b - 1] : 0u, end = call_end[b];
  const int n = (int)(end - beg);
  float sq = 0.f, vs = 0.f, vq = 0.f;
  for (int i = lane; i < n; i += 32) {
    const int64_t t = t0 + beg + i;
    const float2 x = if_raw.ld(c, t);
    sq += x.x * x.x + x.y * x.y;
    const float d = mpx.base[(size_t)c * mpx.cap + ((uint32_t)t & (mpx.cap - 1))];
    vs += d;
    vq += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    vs += __shfl_xor_sync(0xffffffffu, vs, o);
    vq += __shfl_xor_sync(0xffffffffu, vq, o);
  }
  if (lane == 0 && n > 0) {
    float *o = stats + ((size_t)c * n_calls + b) * 3;
    o[0] = sqrtf(sq / (float)n);
    o[1] = vs / (float)n;
    o[2] = sqrtf(vq / (float)n);
  }
}

// PilotPhaseLock::process (PilotPhaseLock.cpp:56-171), FmDecoder::demod_stereo
// (FmDecode.cpp:224-239), both LowPassFilterRC deemphasis filters (FmDecode.cpp:168-170,180),
// per-call statistics EMA (FmDecode.cpp:147-150) and lock bookkeeping, one lane per channel.
//
// The reference evaluates sin/cos of the accumulated phase with libm every sample. Here the
// pilot phasor (sin, cos) is advanced by a rotation through m_freq each sample — the small
// correction to the 19 kHz centre frequency is |d| < 5e-4 rad, so sin d, cos d come from
// three-term series with error < 3e-19 — and re-anchored with an exact sincos(m_phase) at
// the start of every reference call, so rounding cannot accumulate beyond one call. m_phase
// itself is accumulated and wrapped exactly as the reference does.
// ---------------------------------------------------------------------------------------
// k_fm_pll2 — PilotPhaseLock::process + demod_stereo + deemphasis (PilotPhaseLock.cpp:60-180, FmDecode.cpp:152-170), with
// the per-sample recurrence written for the shortest dependent chain. The loop
//   phase -> (sin, cos) -> x*sin, x*cos -> biquads -> fast_atan2f -> loop filter -> freq -> phase
// is strictly serial and one warp per SM sub-partition runs it, so its time is the latency of
// the chain times the number of 384 kHz samples, whatever the channel count. Against a literal
// transcription of the reference loop, none of which alters a rounding that the reference performs in float:
//   * fast_atan2f without branches: min/max instead of the if/else quotient, the octant logic
//     folded into one fused multiply-add  angle = K + sg*base  with K in {0, pi/2, pi} and
//     sg = +-1 chosen from the signs while the division is still in flight (K + sg*base rounds
//     exactly like the reference's single add/subtract), table entries stored as
//     (tbl[i], tbl[i+1]-tbl[i]) pairs in shared memory (one LDS.64 instead of two dependent loads
//     and a subtract; the float difference is the same IEEE operation done once on the host);
//   * the phasor is first rotated by the constant centre frequency (off the critical path) and
//     only the small correction through dl = freq - f0 sits behind the loop filter;
//   * biquad and loop-filter feedback terms are formed before the sample's own product arrives;
//   * phase wrap, period counting and the PPS test are selects plus one rarely taken branch.
struct PllTab {
  float2 e[256]; // (tbl[i], tbl[i+1] - tbl[i])
};

// TAB: callable index -> (tbl[i], tbl[i+1]-tbl[i])
template <typename TAB> __device__ __forceinline__ float fast_atan2f_bf_t(float y, float x, TAB tab) {
  const float ya = fabsf(y), xa = fabsf(x);
  const float num = fminf(ya, xa), den = fmaxf(ya, xa);
  const bool xbig = xa > ya, xpos = x >= 0.0f, ypos = y >= 0.0f;
  float K = xbig ? (xpos ? 0.0f : 3.14159265358979323846f) : 1.57079632679489661923f;
  float sg = (xbig == xpos) ? 1.0f : -1.0f;
  K = ypos ? K : -K;
  sg = ypos ? sg : -sg;
  // z = num / den: reciprocal estimate, quotient, one exact-residual correction. For the operand
  // range of the loop (den is a pilot amplitude, never denormal) this is the correctly rounded
  // quotient in all but a vanishing fraction of cases and within 1 ulp otherwise; it drops the
  // denormal check and its branch from the recurrence.
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float z = num * r;
  z = fmaf(fmaf(-den, z, num), r, z);
  float alpha = z * 255.0f;
  // index = (int)alpha and (float)index without the two conversion instructions (~12 cycles each
  // on the critical path): adding 2^23 with round-towards-zero leaves trunc(alpha) in the low
  // mantissa bits, and subtracting 2^23 again gives it back as a float, both exactly (0 <= alpha <= 255)
  const float t = __fadd_rz(alpha, 8388608.0f);
  alpha -= t - 8388608.0f;
  const float2 te = tab(__float_as_int(t) & 0xff);
  float base = fmaf(te.y, alpha, te.x);
  base = (z < __int_as_float(0x3b808082)) ? z : base; // (double)z < 0.003921569, see fast_atan2f_dev
  const float angle = fmaf(sg, base, K);
  return (den > 0.0f) ? angle : 0.0f; // both inputs zero (Utility.h:245-247)
}

__device__ __forceinline__ float fast_atan2f_bf(float y, float x, const float2 *__restrict__ tab) {
  return fast_atan2f_bf_t(y, x, [tab](int i) { return tab[i]; });
}
// table in shared memory addressed by its 32-bit shared-window address (no generic-pointer
// conversion inside the recurrence)
__device__ __forceinline__ float fast_atan2f_bf_s(float y, float x, unsigned tab_s) {
  return fast_atan2f_bf_t(y, x, [tab_s](int i) {
    float2 v;
    asm("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(tab_s + 8u * (unsigned)i));
    return v;
  });
}

template <typename AV>
static __global__ void __launch_bounds__(32)
    k_fm_pll2(Ring<float> mpx, Ring<AV> out384, FmChanState *__restrict__ st, uint8_t *__restrict__ flags,
              PpsEventDev *__restrict__ pps, const float *__restrict__ stats, const uint32_t *__restrict__ call_end,
              int n_calls, int64_t t0, FmCoreParams P, const float *__restrict__ atan_tbl, int block_off,
              int reset_pps) {
  __shared__ float2 tab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = make_float2(atan_tbl[i], atan_tbl[i + 1] - atan_tbl[i]);
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_channels) return;
  FmChanState s = st[c];
  if (reset_pps) s.n_pps = 0;
  const double kTwoPi = 2.0 * 3.14159265358979323846;
  const double f0 = (19000.0 / 384000.0) * kTwoPi;
  double sf0, cf0;
  sincos(f0, &sf0, &cf0);
  const int n_total = n_calls ? (int)call_end[n_calls - 1] : 0;
  const float *__restrict__ mrow = mpx.base + (size_t)c * mpx.cap;
  AV *__restrict__ orow = out384.base + (size_t)c * out384.cap;
  const uint32_t mmask = mpx.cap - 1, omask = out384.cap - 1, t0lo = (uint32_t)t0;
  const bool stereo = P.stereo != 0, shift = P.pilot_shift != 0, de_st = P.deemph_on_stereo != 0;
  // loop constants pinned in registers (an operand fetched from the constant bank on the critical
  // path costs a load latency every sample)
  double minf = P.pll_minfreq, maxf = P.pll_maxfreq, lf_b0 = P.lf_b0, lf_b1 = P.lf_b1, bq_b0 = P.bq_b0;
  double dlmin = minf - f0, dlmax = maxf - f0;
  asm volatile("" : "+d"(minf), "+d"(maxf), "+d"(lf_b0), "+d"(lf_b1), "+d"(bq_b0), "+d"(dlmin), "+d"(dlmax));
  float nxt[kCoreChunk];
#pragma unroll
  for (int u = 0; u < kCoreChunk; u++) nxt[u] = (u < n_total) ? mrow[(t0lo + (uint32_t)u) & mmask] : 0.f;
  // working registers of the recurrences
  double bi1 = s.bi_x1, bi2 = s.bi_x2, bq1 = s.bq_x1, bq2 = s.bq_x2, lf1 = s.lf_x1;
  double freq = s.pll_freq, phase = s.pll_phase, ferr = s.freq_err;
  double dem = s.de_m_x1, des = s.de_s_x1;
  int periods = s.pilot_periods;
  uint32_t prev_end = 0;
  for (int b = 0; b < n_calls; b++) {
    const uint32_t end = call_end[b];
    const int n = (int)(end - prev_end);
    if (n == 0) { // main.cpp:933-936: the decoder is not called
      flags[(size_t)c * n_calls + b] = (uint8_t)s.stereo_detected;
      continue;
    }
    const uint32_t beg = prev_end;
    prev_end = end;
    s.decoder_calls++;
    const bool was_locked = (s.lock_cnt >= P.lock_delay);
    double last_i = 0.0, last_q = 0.0;
    double psin = 0.0, pcos = 1.0;
    if (stereo) sincos(phase, &psin, &pcos); // exact re-anchor once per reference call
    for (int i0 = 0; i0 < n; i0 += kCoreChunk) {
      const int valid = (n - i0 < kCoreChunk) ? (n - i0) : kCoreChunk;
      float din[kCoreChunk];
#pragma unroll
      for (int u = 0; u < kCoreChunk; u++) din[u] = nxt[u];
      {
        const int pos = (int)beg + i0 + valid;
#pragma unroll
        for (int u = 0; u < kCoreChunk; u++) {
          nxt[u] = (pos + u < n_total) ? mrow[(t0lo + (uint32_t)(pos + u)) & mmask] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < kCoreChunk; u++) {
        if (u >= valid) break;
        const int i = i0 + u;
        const double xd = (double)din[u];
        double ster = 0.0;
        if (stereo) {
          // ---- off the critical path: feedback terms, centre-frequency rotation, tone
          const double fb_i = P.bq_a1 * bi1 + P.bq_a2 * bi2;
          const double fb_q = P.bq_a1 * bq1 + P.bq_a2 * bq2;
          const double fb_l = __dmul_rn(lf_b1, lf1);
          const double as = psin * cf0 + pcos * sf0; // sin(phase + f0)
          const double ac = pcos * cf0 - psin * sf0; // cos(phase + f0)
          const double tone = shift ? (2 * pcos * pcos - 1) : (2 * psin * pcos);
          // ---- critical path
          const double i0v = psin * xd - fb_i;
          const double q0v = pcos * xd - fb_q;
          const double new_i = bq_b0 * i0v;
          const double new_q = bq_b0 * q0v;
          bi2 = bi1;
          bi1 = i0v;
          bq2 = bq1;
          bq1 = q0v;
          const double perr = (double)fast_atan2f_bf((float)new_q, (float)new_i, tab);
          last_i = new_i;
          last_q = new_q;
          ferr = fma(lf_b0, perr, fb_l);
          lf1 = perr;
          const double fraw = freq + ferr;
          // std::max(m_minfreq, std::min(m_maxfreq, m_freq)) (PilotPhaseLock.cpp:119) with both
          // comparisons taken on the raw value, and the same clamp applied to dl = freq - f0
          // directly (maxf - f0 and minf - f0 are the values the subtraction would give)
          const bool below_max = fraw < maxf, above_min = minf < fraw;
          const double dlr = fraw - f0;
          freq = below_max ? (above_min ? fraw : minf) : maxf;
          const double dl = below_max ? (above_min ? dlr : dlmin) : dlmax;
          {
            // sin(dl) = dl - dl^3/6, cos(dl) - 1 = -dl^2/2: |dl| < 4.91e-4 (the +-30 Hz clamp), so the
            // dropped terms are below 2.5e-15 per sample and the phasor is re-anchored every call
            const double d2 = dl * dl;
            const double sd = fma(dl * d2, -1.0 / 6.0, dl);
            psin = fma(ac, sd, fma(-0.5 * as, d2, as));
            pcos = fma(-as, sd, fma(-0.5 * ac, d2, ac));
          }
          // ---- phase accumulator (its own short recurrence), wrap and PPS
          phase += freq;
          const bool wrap = phase > kTwoPi;
          phase = wrap ? phase - kTwoPi : phase;
          periods += wrap ? 1 : 0;
          if (wrap && periods == 19000) {
            periods = 0;
            if (was_locked) {
              if (s.n_pps < (uint32_t)kMaxPps) {
                PpsEventDev ev;
                ev.pps_index = s.pps_cnt;
                ev.sample_index = s.sample_cnt + (unsigned long long)i;
                ev.block_position = (double)i / (double)n;
                ev.block = (uint32_t)(b + block_off);
                ev.pad = 0;
                pps[(size_t)c * kMaxPps + s.n_pps] = ev;
              }
              s.n_pps++;
              s.pps_cnt++;
            }
          }
          ster = (tone * xd) * 2.0;
          if (de_st) {
            const double x0 = ster - P.de_a1 * des;
            ster = P.de_b0 * x0;
            des = x0;
          }
        }
        const double m0 = xd - P.de_a1 * dem;
        const double mono = P.de_b0 * m0;
        dem = m0;
        orow[(t0lo + beg + (uint32_t)i) & omask] = aud_mk<AV>(mono, ster);
      }
    }
    // per-call statistics (FmDecode.cpp:95,146-150)
    {
      const float *sv = stats + ((size_t)c * n_calls + b) * 3;
      s.if_rms = sv[0];
      s.baseband_mean = (float)(0.95 * (double)s.baseband_mean + 0.05 * (double)sv[1]);
      s.baseband_level = (float)(0.95 * (double)s.baseband_level + 0.05 * (double)sv[2]);
    }
    if (stereo) {
      s.pilot_level = sqrt(last_i * last_i + last_q * last_q); // PilotPhaseLock.cpp:106, last sample
      if (2 * s.pilot_level > P.minsignal) {                     // PilotPhaseLock.cpp:153-170
        if (s.lock_cnt < P.lock_delay) s.lock_cnt += n;
      } else {
        s.lock_cnt = 0;
      }
      if (s.lock_cnt < P.lock_delay) {
        periods = 0;
        s.pps_cnt = 0;
        while (s.n_pps > 0 && s.n_pps <= (uint32_t)kMaxPps &&
               pps[(size_t)c * kMaxPps + s.n_pps - 1].block == (uint32_t)(b + block_off)) {
          s.n_pps--;
        }
      }
      s.sample_cnt += (unsigned long long)n;
      s.stereo_detected = (s.lock_cnt >= P.lock_delay) ? 1 : 0;
    }
    flags[(size_t)c * n_calls + b] = (uint8_t)s.stereo_detected;
  }
  {
    FmChanState *o = st + c;
    o->baseband_mean = s.baseband_mean;
    o->baseband_level = s.baseband_level;
    o->if_rms = s.if_rms;
    o->stereo_detected = s.stereo_detected;
    o->lock_cnt = s.lock_cnt;
    o->pilot_periods = periods;
    o->n_pps = s.n_pps;
    o->pll_phase = phase;
    o->pll_freq = freq;
    o->bi_x1 = bi1;
    o->bi_x2 = bi2;
    o->bq_x1 = bq1;
    o->bq_x2 = bq2;
    o->lf_x1 = lf1;
    o->pilot_level = s.pilot_level;
    o->freq_err = ferr;
    o->pps_cnt = s.pps_cnt;
    o->sample_cnt = s.sample_cnt;
    o->de_m_x1 = dem;
    o->de_s_x1 = des;
    o->decoder_calls = s.decoder_calls;
  }
}

// ---------------------------------------------------------------------------------------
// 48 kHz tail, one lane per channel: DC block (HighPassFilterIir::process_inplace,
// Filter.cpp:304-311, biquad Filter.cpp:243-250) on mono and L-R, then the per-call
// matrix (FmDecoder::process FmDecode.cpp:194-220, stereo_to_left_right :255-270).
struct FmTailParams {
  double b0, b1, b2, a1, a2; // HighPassFilterIir(0.0001) FmDecode.cpp:62
  int stereo, pilot_shift;
  int n_channels;
};

// One CTA = 32 channels. Warp 0 runs the recurrences (lane = channel) and nothing else: a single warp pays the full
// latency of every instruction it issues, so everything that is not on the dependent chain is done by the three
// helper warps beside it. Warps 2 and 3 move tiles of kTailT samples between HBM and shared memory in rows (lane =
// sample: 512 contiguous bytes of one channel per instruction, transposed through shared memory with a +1 pad), one
// tile ahead of / behind the recurrence warp. Warp 1 prepares, for the next tile, the list of reference calls that
// intersect it: (end, stereo flags of the 32 channels as one word) — lane = channel loads the flag bytes of four calls
// at a time and a ballot packs them. A reference call is only ~10 samples long at 48 kHz: a lookup inside the
// recurrence would stall it every few samples, and behind a batch of row loads every lookup would wait for the batch.
// The recurrence warp works in groups of eight samples: first only the two dependent chains (DC-block state of mono
// and L-R, 2 FP64 operations per sample each), then the eight outputs, which are independent of each other.
constexpr int kTailT = 32;
constexpr int kTailG = 8;
constexpr int kTailThreads = 128;
template <typename AV> struct TailSmem {
  AV tin[2][32][kTailT + 1];
  double2 tout[2][32][kTailT + 1];
  uint32_t ent_end[2][kTailT + 2];  // [tile parity][entry]: end (sample index of the launch) of a call of the tile
  uint32_t ent_mask[2][kTailT + 2]; // bit r = stereo flag of channel c0 + r during that call
};

template <typename AV>
static __global__ void __launch_bounds__(kTailThreads)
k_fm_tail(Ring<AV> in48, double *__restrict__ audio, size_t audio_stride, FmChanState *__restrict__ st,
          const uint8_t *__restrict__ flags, const uint32_t *__restrict__ call_end48, int n_calls, int64_t j0,
          FmTailParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TailSmem<AV> &sm = *reinterpret_cast<TailSmem<AV> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = blockIdx.x * 32, c = c0 + lane;
  const int rows = min(32, P.n_channels - c0);
  const int n_total = n_calls ? (int)call_end48[n_calls - 1] : 0;
  const int n_tiles = (n_total + kTailT - 1) / kTailT;
  // 16-byte row stores need every channel row of the caller's buffer on a 16-byte boundary
  const bool wide = P.stereo && (audio_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(audio) & 15) == 0;

  // warp 1: the calls of tile k in order; empty calls and calls that ended before the tile are left out, the last
  // entry reaches the end of the tile
  int bq = 0; // first call that may still intersect the next tile (warp-uniform)
  auto flag_tile = [&](int k) {
    if (!P.stereo) return;
    const int par = k & 1;
    const uint32_t t1 = (uint32_t)min((k + 1) * kTailT, n_total);
    uint32_t pos = (uint32_t)(k * kTailT);
    int n = 0;
    const uint8_t *frow = flags + (size_t)min(c, P.n_channels - 1) * n_calls;
    for (bool done = false; !done;) {
      uint32_t e[4], m[4];
      uint8_t f[4];
      const int b0 = bq;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int bb = min(b0 + q, n_calls - 1);
        e[q] = call_end48[bb];
        f[q] = frow[bb];
      }
#pragma unroll
      for (int q = 0; q < 4; q++) m[q] = __ballot_sync(0xffffffffu, f[q] != 0 && lane < rows);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (done) break;
        if (e[q] > pos) { // a call with samples in [pos, ..)
          if (lane == 0) {
            sm.ent_end[par][n] = e[q];
            sm.ent_mask[par][n] = m[q];
          }
          n++;
          pos = e[q];
        }
        if (e[q] >= t1) {
          done = true; // this call may go on into the next tile: bq stays on it
        } else {
          bq++;
        }
      }
    }
  };
  // warps 2, 3: the even / odd rows of tile k (all loads of a warp in flight together)
  auto load_tile = [&](int k) {
    const int par = k & 1, i = k * kTailT + lane;
    AV v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int r = (warp - 2) + 2 * q;
      v[q] = (r < rows && i < n_total) ? in48.ld(c0 + r, j0 + i) : aud_mk<AV>(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) sm.tin[par][(warp - 2) + 2 * q][lane] = v[q];
  };
  auto store_tile = [&](int k) { // warps 2, 3: the even / odd rows of tile k
    const int par = k & 1;
    const size_t i = (size_t)k * kTailT + lane;
    if (i >= (size_t)n_total) return;
    for (int r = warp - 2; r < rows; r += 2) {
      double *o = audio + (size_t)(c0 + r) * audio_stride;
      const double2 v = sm.tout[par][r][lane];
      if (!P.stereo) {
        o[i] = v.x;
      } else if (wide) {
        *reinterpret_cast<double2 *>(o + 2 * i) = v;
      } else {
        o[2 * i] = v.x;
        o[2 * i + 1] = v.y;
      }
    }
  };

  const bool live = warp == 0 && c < P.n_channels;
  double m1 = 0.0, m2 = 0.0, s1 = 0.0, s2 = 0.0;
  if (live) {
    m1 = st[c].dc_m_x1;
    m2 = st[c].dc_m_x2;
    s1 = st[c].dc_s_x1;
    s2 = st[c].dc_s_x2;
  }
  // one output pair from the filter states (m0, m1, m2 newest first), exactly the reference's expressions
  auto emit = [&](double m0, double ma, double mb, double s0, double sa, double sb2, bool det) -> double2 {
    const double mono = P.b0 * m0 + P.b1 * ma + P.b2 * mb;
    if (!P.stereo) return make_double2(mono, 0.0);
    const double ster = P.b0 * s0 + P.b1 * sa + P.b2 * sb2;
    double l, rr;
    if (P.pilot_shift) {
      l = det ? ster : 0.0;
      rr = l;
    } else {
      const double sb = 1.017 * ster;
      l = det ? mono + sb : mono;
      rr = det ? mono - sb : mono;
    }
    return make_double2(l, rr);
  };
  if (n_tiles > 0) {
    if (warp == 1) flag_tile(0);
    if (warp >= 2) load_tile(0);
  }
  __syncthreads();
  for (int k = 0; k < n_tiles; k++) {
    const int par = k & 1;
    if (warp == 1) {
      if (k + 1 < n_tiles) flag_tile(k + 1);
    } else if (warp >= 2) {
      if (k + 1 < n_tiles) load_tile(k + 1);
      if (k > 0) store_tile(k - 1);
    } else if (live) {
      const int nu = min(kTailT, n_total - k * kTailT);
      int ej = 0;
      uint32_t eend = P.stereo ? sm.ent_end[par][0] : 0xffffffffu, emask = P.stereo ? sm.ent_mask[par][0] : 0u;
      int u0 = 0;
      for (; u0 + kTailG <= nu; u0 += kTailG) {
        AV x[kTailG];
        double mv[kTailG + 2], sv[kTailG + 2]; // [q + 2] = state after sample q; [1], [0] = the two before the group
        bool det[kTailG];
#pragma unroll
        for (int q = 0; q < kTailG; q++) x[q] = sm.tin[par][lane][u0 + q];
        mv[1] = m1;
        mv[0] = m2;
        sv[1] = s1;
        sv[0] = s2;
#pragma unroll
        for (int q = 0; q < kTailG; q++) { // the dependent chains only
          mv[q + 2] = (double)x[q].x - (P.a1 * mv[q + 1] + P.a2 * mv[q]);
          if (P.stereo) sv[q + 2] = (double)x[q].y - (P.a1 * sv[q + 1] + P.a2 * sv[q]);
        }
#pragma unroll
        for (int q = 0; q < kTailG; q++) {
          const uint32_t i = (uint32_t)(k * kTailT + u0 + q);
          while (i >= eend) {
            ej++;
            eend = sm.ent_end[par][ej];
            emask = sm.ent_mask[par][ej];
          }
          det[q] = (emask >> lane) & 1u;
        }
#pragma unroll
        for (int q = 0; q < kTailG; q++) {
          sm.tout[par][lane][u0 + q] = emit(mv[q + 2], mv[q + 1], mv[q], sv[q + 2], sv[q + 1], sv[q], det[q]);
        }
        m1 = mv[kTailG + 1];
        m2 = mv[kTailG];
        if (P.stereo) {
          s1 = sv[kTailG + 1];
          s2 = sv[kTailG];
        }
      }
      for (int u = u0; u < nu; u++) { // the last, shorter group of the launch
        const AV x = sm.tin[par][lane][u];
        const double m0 = (double)x.x - (P.a1 * m1 + P.a2 * m2);
        double s0 = 0.0;
        if (P.stereo) s0 = (double)x.y - (P.a1 * s1 + P.a2 * s2);
        const uint32_t i = (uint32_t)(k * kTailT + u);
        while (i >= eend) {
          ej++;
          eend = sm.ent_end[par][ej];
          emask = sm.ent_mask[par][ej];
        }
        sm.tout[par][lane][u] = emit(m0, m1, m2, s0, s1, s2, (emask >> lane) & 1u);
        m2 = m1;
        m1 = m0;
        if (P.stereo) {
          s2 = s1;
          s1 = s0;
        }
      }
    }
    __syncthreads();
  }
  if (warp >= 2 && n_tiles > 0) store_tile(n_tiles - 1);
  if (live) {
    st[c].dc_m_x1 = m1;
    st[c].dc_m_x2 = m2;
    st[c].dc_s_x1 = s1;
    st[c].dc_s_x2 = s2;
  }
}

// Keep the last kHist input samples of every channel for the next call's halo.
template <typename V>
__global__ void k_save_hist(const V *__restrict__ lin, size_t stride, int64_t n_new, const V *__restrict__ hist_old,
                            V *__restrict__ hist_new, int fmt) {
  const uint32_t c = blockIdx.x;
  for (int i = threadIdx.x; i < kHist; i += blockDim.x) {
    const int64_t r = n_new - kHist + i; // index into this call's samples
    V v;
    if (r >= 0) {
      if (fmt == 1) {
        const short2 q = reinterpret_cast<const short2 *>(lin)[(size_t)c * stride + r];
        v.x = (float)q.x * (1.0f / 32768.0f);
        v.y = (float)q.y * (1.0f / 32768.0f);
      } else {
        v = lin[(size_t)c * stride + r];
      }
    } else {
      const int64_t h = kHist + r; // = i + n_new, < kHist
      v = hist_old[(size_t)c * kHist + h];
    }
    hist_new[(size_t)c * kHist + i] = v;
  }
}

} // namespace fmr
#endif
