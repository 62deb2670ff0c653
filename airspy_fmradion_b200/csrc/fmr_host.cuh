// fmr_host.cuh — host-side plumbing shared by the FM and AM handles: error reporting,
// device buffers, the generic streaming resampler (half-band cascade -> long low-pass ->
// polyphase bank) that stands in for r8b::CDSPResampler (IfResampler.cpp:25-79,
// AudioResampler.cpp:25-61), and pinned staging for the per-call tables.
#ifndef FMR_HOST_CUH
#define FMR_HOST_CUH

#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fmradion_b200.h"
#include <cmath>
#include <complex>

#include "fmr_fft.cuh"
#include "fmr_fft_inplace.cuh"
#include "fmr_fdr.cuh"
#include "fmr_frontend.cuh"
#include "fmr_hbstream.cuh"
#include "fmr_kernels.cuh"
#include "fmr_tables.h"

namespace fmr {

// cudaFuncAttributeMaxDynamicSharedMemorySize is ONE value per kernel (and device). Kernels that several resamplers or
// handles of a process share (k_fir_long, k_frac_interp, k_fir_quirk) must never have it lowered by the one that is
// created last and needs less: keep the largest request per (kernel, device) and set that.
template <typename F> inline cudaError_t raise_smem_limit(F *func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> cur;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  size_t &v = cur[std::make_pair(reinterpret_cast<const void *>(func), dev)];
  if (bytes > v) v = bytes;
  return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v);
}


extern thread_local std::string g_err;

inline fmr_status fail(fmr_status code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define FMR_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      return fmr::fail(FMR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    }                                                                                          \
  } while (0)

inline uint32_t pow2ceil(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return (uint32_t)p;
}

struct DevMem {
  std::vector<void *> ptrs;
  size_t total = 0;
  template <typename T> cudaError_t alloc(T **p, size_t count, bool zero = true) {
    void *q = nullptr;
    const size_t bytes = count * sizeof(T);
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 1);
    if (e != cudaSuccess) return e;
    ptrs.push_back(q);
    total += bytes;
    if (zero && bytes) {
      e = cudaMemset(q, 0, bytes);
      if (e != cudaSuccess) return e;
    }
    *p = reinterpret_cast<T *>(q);
    return cudaSuccess;
  }
  void release() {
    for (void *q : ptrs) cudaFree(q);
    ptrs.clear();
  }
};

// Pinned staging slots for small per-call tables; a slot is reused only after the copy
// that read it has completed.
struct PinnedSlots {
  static const int kSlots = 8;
  void *host[kSlots] = {nullptr};
  cudaEvent_t ev[kSlots] = {nullptr};
  bool used[kSlots] = {false};
  size_t bytes = 0;
  int next = 0;
  cudaError_t init(size_t b) {
    bytes = b;
    for (int i = 0; i < kSlots; i++) {
      cudaError_t e = cudaMallocHost(&host[i], b ? b : 1);
      if (e != cudaSuccess) return e;
      e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  void *acquire(int *slot) {
    const int s = next;
    next = (next + 1) % kSlots;
    if (used[s]) cudaEventSynchronize(ev[s]);
    used[s] = true;
    *slot = s;
    return host[s];
  }
  void commit(int slot, cudaStream_t st) { cudaEventRecord(ev[slot], st); }
  void release() {
    for (int i = 0; i < kSlots; i++) {
      if (host[i]) cudaFreeHost(host[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
      host[i] = nullptr;
      ev[i] = nullptr;
    }
  }
};

// Optional per-stage timing with CUDA events on the launching stream.
struct Prof {
  static const int kMax = 16;
  bool on = false;
  const char *names[kMax] = {nullptr};
  cudaEvent_t a[kMax] = {nullptr}, b[kMax] = {nullptr};
  bool ran[kMax] = {false};
  int n = 0;
  int add(const char *name) {
    names[n] = name;
    return n++;
  }
  void enable() {
    if (on) return;
    for (int i = 0; i < n; i++) {
      cudaEventCreate(&a[i]);
      cudaEventCreate(&b[i]);
    }
    on = true;
  }
  float acc[kMax] = {0.f}; // time of earlier launches of the stage within the same public call (host calls run in chunks)
  void reset() {
    for (int i = 0; i < n; i++) {
      ran[i] = false;
      acc[i] = 0.f;
    }
  }
  void begin(int i, cudaStream_t st) {
    if (!on || i < 0) return;
    if (ran[i]) { // the stage already ran in this call: bank its time before the event pair is reused
      float t = 0.f;
      if (cudaEventSynchronize(b[i]) == cudaSuccess && cudaEventElapsedTime(&t, a[i], b[i]) == cudaSuccess) acc[i] += t;
    }
    cudaEventRecord(a[i], st);
  }
  void end(int i, cudaStream_t st) {
    if (on && i >= 0) {
      cudaEventRecord(b[i], st);
      ran[i] = true;
    }
  }
  fmr_status read(float *ms, const char **nm, uint32_t cap, uint32_t *cnt) {
    uint32_t k = 0;
    for (int i = 0; i < n && k < cap; i++) {
      if (!on || !ran[i]) continue;
      float t = 0;
      if (cudaEventSynchronize(b[i]) != cudaSuccess || cudaEventElapsedTime(&t, a[i], b[i]) != cudaSuccess) {
        return fail(FMR_ERR_CUDA, "event timing failed");
      }
      ms[k] = acc[i] + t;
      nm[k] = names[i];
      k++;
    }
    *cnt = k;
    return FMR_OK;
  }
  void release() {
    if (!on) return;
    for (int i = 0; i < n; i++) {
      cudaEventDestroy(a[i]);
      cudaEventDestroy(b[i]);
    }
    on = false;
  }
};

// Host-side double-precision radix-2 FFT (only used once per handle to build the filter
// spectrum for k_fir_fft).
inline void host_fft(std::vector<std::complex<double>> &a) {
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = -2.0 * M_PI / (double)len;
    for (size_t i = 0; i < n; i += len) {
      for (size_t k = 0; k < len / 2; k++) {
        const std::complex<double> w(std::cos(ang * (double)k), std::sin(ang * (double)k));
        const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
    }
  }
}

// Filter outputs per launch below which the direct-form kernels are used instead of the FFT one
// (a partial 8192-point block still beats klen multiply-adds per output well below one block).
constexpr int kFftMinOutF32 = 1500;
constexpr int kFftMinOutF64 = 500;

template <typename S> struct Resampler {
  using V = typename V2<S>::type;
  using Scalar = S;
  const ChainDesc *d = nullptr;
  int C = 0;
  bool linear_in = false;
  Ring<V> r_hb{nullptr, 0}, r_bc{nullptr, 0};
  // channel group the next run() works on (the caller offsets `src` and `out` itself)
  int gc0 = 0, gcn = 0;
  Ring<V> sub(Ring<V> r) const { return Ring<V>{r.base + (size_t)gc0 * r.cap, r.cap}; }
  S *d_bc = nullptr, *d_fi = nullptr;
  HbTaps<S> hbt;
  int64_t cum_in = 0;
  size_t smem_hb = 0, smem_fir = 0, smem_fi = 0;
  Prof *prof = nullptr;
  int p_hb = -1, p_bc = -1, p_fi = -1;
  V *d_H16 = nullptr, *d_H8 = nullptr; // filter spectra for k_fir_fft (16384: float chains only)
  V *d_twtab = nullptr;                // full twiddle tables of the 16384-point kernel (FMR_FFT_TW=0: off)
  bool fft_tw = false;
  V *d_H16rev = nullptr, *d_iptab = nullptr; // in-place form (fmr_fft_inplace.cuh)
  bool fft_inplace = true;                   // FMR_FFT_INPLACE=0: the Stockham form (k_fir_fft)
  bool use_fdr = false;                      // frequency-domain low-pass + resampling (fmr_fdr.cuh); FMR_FDR=0: off
  int fdr_rl = 0;                            // last radix of its inverse transform: 12 (625:192), 15 (125:48), 10 (125:32)
  int fdr_adv_in = 0, fdr_guard_in = 0, fdr_adv_out = 0, fdr_guard_out = 0; // block grid of this chain
  float *d_fdr_Hs = nullptr;
  float2 *d_fdr_tab = nullptr;
  bool use_fe = false;                       // fused persistent front end (fmr_frontend.cuh); FMR_FE=0: off
  int fe_variant = 0;                        // FMR_FE_VARIANT: 0 = CfgC4 (default), 1 = CfgS, 2 = CfgA (fmr_frontend.cuh)
  int fe_min_blocks = 2;                     // fewer whole blocks inside the call's buffer: unfused kernels only
  int p_fe = -1;
  typedef CUresult (*TmEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  TmEncodeFn tm_encode = nullptr;
  uint64_t last_plan[5] = {0, 0, 0, 0, 0};   // fmr_fm_last_plan
  int64_t fdr_next = 0;                      // first block of the absolute grid that has not been computed yet
  int64_t fdr_out_done = 0;                  // outputs stored so far (beyond fdr_next's blocks: a partial block's head)
  int64_t fdr_hist = 512;                    // output samples before f0 that readers of the output ring may still need
  bool use_fft = false;
  bool use_dec2 = false; // double chains with a decimate-by-2 low-pass (audio resampler)
  bool fuse_fi = true;   // FMR_FUSE_FI=0: keep the polyphase bank as its own launch
  bool hb_stream = false; // streaming register-resident half-band cascade (10 MHz chain, cf32 input)
  bool hbs_tma = true;    // FMR_HBS_TMA=0: cp.async (LDGSTS) staging instead of TMA bulk copies
  int sm_count = 148;
  int fft_min_out = 0;

  static size_t hb_smem(const HbTaps<S> &t, int nst) {
    size_t total = 0;
    // levels 0 and 1 are live together; level 2 aliases level 0 (see k_hb_cascade)
    for (int s = 0; s < nst && s < 2; s++) {
      total += 2 * (size_t)kHbR * hb_sub_len(hb_level_len(t.n, nst, s, hb_tile_of(t.n[0], t.n[1])), (int)sizeof(V));
    }
    return total * sizeof(V) + 16;
  }

  // Dispatch over the half-band tap-count combinations of the shipped chains. F is called with
  // the kernel's function pointer.
  template <typename F> bool hb_dispatch(F &&f) const {
    const int n1 = hbt.n[0], n2 = hbt.n[1], n3 = hbt.n[2];
#define FMR_HB_COMBO(NST, A, B, Cc)                                                       \
  if (d->n_hb == NST && n1 == A && n2 == B && n3 == Cc) {                                  \
    if (linear_in) {                                                                      \
      f(k_hb_cascade<S, NST, true, A, B, Cc>);                                            \
    } else {                                                                              \
      f(k_hb_cascade<S, NST, false, A, B, Cc>);                                           \
    }                                                                                     \
    return true;                                                                          \
  }
    FMR_HB_COMBO(0, 0, 0, 0)
    FMR_HB_COMBO(3, 4, 5, 8)  // 10 MHz -> 384 kHz
    FMR_HB_COMBO(2, 5, 8, 0)  // 6 MHz
    FMR_HB_COMBO(1, 8, 0, 0)  // 2.5 MHz
    FMR_HB_COMBO(2, 6, 11, 0) // 384 kHz -> 48 kHz (IfResampler spec, AM)
    FMR_HB_COMBO(2, 7, 13, 0) // 384 kHz -> 48 kHz (AudioResampler spec)
    FMR_HB_COMBO(1, 11, 0, 0) // 2.048 MHz -> 384 kHz; 192 kHz, 256 kHz -> 48 kHz
    FMR_HB_COMBO(3, 5, 6, 11) // 768 kHz, 1 MHz -> 48 kHz
#undef FMR_HB_COMBO
    return false;
  }

  fmr_status init(const ChainDesc *desc, int channels, int64_t max_in, bool lin, DevMem &mem) {
    d = desc;
    C = channels;
    linear_in = lin;
    cum_in = 0;
    memset(&hbt, 0, sizeof(hbt));
    for (int s = 0; s < d->n_hb; s++) {
      hbt.n[s] = d->hb[s].ntaps;
      if (hbt.n[s] > 14) return fail(FMR_ERR_UNSUPPORTED, "half-band stage longer than 14 taps");
      for (int k = 0; k < hbt.n[s]; k++) hbt.t[s][k] = (S)d->hb[s].taps[k];
    }
    for (int s = 0; s < d->n_hb; s++) hbt.sl[s] = hb_sub_len(hb_level_len(hbt.n, d->n_hb, s, hb_tile_of(hbt.n[0], hbt.n[1])), (int)sizeof(V));
    smem_hb = hb_smem(hbt, d->n_hb);
    cudaError_t e = cudaSuccess;
    const size_t smem_need = smem_hb;
    if (!hb_dispatch([&](auto kern) {
          e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need);
        })) {
      return fail(FMR_ERR_UNSUPPORTED, "half-band tap combination not instantiated");
    }
    FMR_CUDA(e);
    if constexpr (sizeof(S) == sizeof(float)) {
      if (lin && d->n_hb == 3 && hbt.n[0] == 4 && hbt.n[1] == 5 && hbt.n[2] == 8 && !env_off("FMR_HB_STREAM")) {
        FMR_CUDA((cudaFuncSetAttribute(k_hb_stream<4, 5, 8, kHbsU, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       hbs_smem_bytes(4))));
        hb_stream = true;
        hbs_tma = !env_off("FMR_HBS_TMA");
        FMR_CUDA((cudaFuncSetAttribute(k_hb_stream_tma<4, 5, 8, kHbsU, 2, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       HbsTma<2>::smem_bytes(3, 4))));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      }
    }
    const int64_t max_hb = (max_in >> d->n_hb) + 4;
    if (d->n_hb > 0 || lin) {
      r_hb.cap = pow2ceil((uint64_t)(max_hb + d->bc.latency + d->bc.klen + 64));
      FMR_CUDA(mem.alloc(&r_hb.base, (size_t)C * r_hb.cap));
    }
    {
      std::vector<S> h(d->bc.klen);
      for (int i = 0; i < d->bc.klen; i++) h[i] = (S)d->bc.taps[i];
      FMR_CUDA(mem.alloc(&d_bc, h.size(), false));
      FMR_CUDA(cudaMemcpy(d_bc, h.data(), h.size() * sizeof(S), cudaMemcpyHostToDevice));
    }
    if (d->bc.klen < 8192 / 2 && !env_off("FMR_FFT") && !(sizeof(S) == sizeof(double) && env_off("FMR_FFT_F64"))) {
      // filter spectra, computed in double, 1/N folded in
      auto make_H = [&](int n, V **dst) -> cudaError_t {
        std::vector<std::complex<double>> hc(n, std::complex<double>(0.0, 0.0));
        for (int i = 0; i < d->bc.klen; i++) hc[i] = d->bc.taps[i];
        host_fft(hc);
        std::vector<V> hf(n);
        for (int i = 0; i < n; i++) {
          hf[i].x = (S)(hc[i].real() / n);
          hf[i].y = (S)(hc[i].imag() / n);
        }
        cudaError_t e2 = mem.alloc(dst, (size_t)n, false);
        if (e2 != cudaSuccess) return e2;
        return cudaMemcpy(*dst, hf.data(), sizeof(V) * n, cudaMemcpyHostToDevice);
      };
      FMR_CUDA(make_H(8192, &d_H8));
      FMR_CUDA((cudaFuncSetAttribute(k_fir_fft<S, 8192, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     FftCfg<S, 8192>::kSmemBytes)));
      FMR_CUDA((cudaFuncSetAttribute(k_fir_fft<S, 8192, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     FftCfg<S, 8192>::kSmemBytes)));
      if constexpr (sizeof(S) == sizeof(float)) {
        FMR_CUDA(make_H(16384, &d_H16));
        FMR_CUDA((cudaFuncSetAttribute(k_fir_fft<S, 16384, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FftCfg<S, 16384>::kSmemBytes)));
        FMR_CUDA((cudaFuncSetAttribute(k_fir_fft<S, 16384, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FftCfg<S, 16384>::kSmemBytes)));
        FMR_CUDA((cudaFuncSetAttribute(k_fir_fft<S, 16384, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FftCfg<S, 16384>::kSmemBytesTw)));
        // full twiddle tables of the two twiddled radix-16 passes, computed in double (fmr_fft.cuh, kTwTab*)
        std::vector<V> tt(kTwTabLen);
        for (int r = 0; r < 16; r++) {
          for (int k = 0; k < 16; k++) {
            const double a = -2.0 * M_PI * (double)(k * r) / 256.0;
            tt[kTwTabP16 + r * 16 + k].x = (S)std::cos(a);
            tt[kTwTabP16 + r * 16 + k].y = (S)std::sin(a);
          }
          for (int k = 0; k < 256; k++) {
            const double a = -2.0 * M_PI * (double)(k * r) / 4096.0;
            tt[kTwTabP256 + r * 256 + k].x = (S)std::cos(a);
            tt[kTwTabP256 + r * 256 + k].y = (S)std::sin(a);
          }
        }
        FMR_CUDA(mem.alloc(&d_twtab, tt.size(), false));
        FMR_CUDA(cudaMemcpy(d_twtab, tt.data(), sizeof(V) * tt.size(), cudaMemcpyHostToDevice));
        fft_tw = !env_off("FMR_FFT_TW");
        {
          // in-place (DIF / DIT) form of the same block, fmr_fft_inplace.cuh: spectrum in digit-reversed order,
          // its twiddle tables; FMR_FFT_INPLACE=1 selects it
          using namespace ipfft;
          std::vector<std::complex<double>> hc(kN, std::complex<double>(0.0, 0.0));
          for (int i = 0; i < d->bc.klen; i++) hc[i] = d->bc.taps[i];
          host_fft(hc);
          std::vector<V> hr(kN), tb(kTabLen);
          for (int pz = 0; pz < kN; pz++) {
            const std::complex<double> hv = hc[freq_of_pos(pz)] / (double)kN;
            hr[pz].x = (S)hv.real();
            hr[pz].y = (S)hv.imag();
          }
          auto wv = [](double num, double den) {
            const double a = -2.0 * M_PI * num / den;
            V v;
            v.x = (S)std::cos(a);
            v.y = (S)std::sin(a);
            return v;
          };
          for (int q = 0; q < 128; q++) {
            tb[kTw + q] = wv(128.0 * q, (double)kN);
            tb[kTw + 128 + q] = wv((double)q, (double)kN);
          }
          for (int dd = 0; dd < 16; dd++) {
            for (int b = 0; b < 64; b++) tb[kT64 + dd * 64 + b] = wv((double)(b * dd), 1024.0);
            for (int b = 0; b < 4; b++) tb[kT4 + dd * 4 + b] = wv((double)(b * dd), 64.0);
          }
          FMR_CUDA(mem.alloc(&d_H16rev, hr.size(), false));
          FMR_CUDA(cudaMemcpy(d_H16rev, hr.data(), sizeof(V) * hr.size(), cudaMemcpyHostToDevice));
          FMR_CUDA(mem.alloc(&d_iptab, tb.size(), false));
          FMR_CUDA(cudaMemcpy(d_iptab, tb.data(), sizeof(V) * tb.size(), cudaMemcpyHostToDevice));
          FMR_CUDA((cudaFuncSetAttribute(k_fir_fft_ip<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIpSmemBytes)));
          fft_inplace = !env_off("FMR_FFT_INPLACE");
        }
      }
      use_fft = true;
      // fewer outputs than this: direct form (its cost grows with the outputs, an FFT block costs the same for any count;
      // a decimating convolver sees `down` input samples per output)
      fft_min_out = (sizeof(S) == sizeof(float)) ? kFftMinOutF32 / d->bc.down : kFftMinOutF64;
      fuse_fi = !env_off("FMR_FUSE_FI");
    }
    if constexpr (sizeof(S) == sizeof(float)) {
      // 625:192, 125:48 and 125:32 pairs: the low-pass and the polyphase bank as one forward + one small inverse FFT
      // (fmr_fdr.cuh). Blocks of 10000 input samples; guard = the filter's half length + the bank's, rounded up to a
      // whole number of input steps so that every block starts on an output sample.
      const int fdr_j = d->has_fi && d->fi.instep > 0 ? fdr::kNin / d->fi.instep : 0;
      const int fdr_nout = fdr_j * (d->has_fi ? d->fi.outstep : 0);
      const int rl = (fdr_j * d->fi.instep == fdr::kNin && fdr_nout % 256 == 0) ? fdr_nout / 256 : 0;
      if (use_fft && d->has_fi && d->bc.down == 1 && (rl == 12 || rl == 15 || rl == 10) && !env_off("FMR_FDR")) {
        const int need = (d->bc.klen - 1) / 2 + d->fi.flen / 2 + 1;
        fdr_guard_in = (need + d->fi.instep - 1) / d->fi.instep * d->fi.instep;
        fdr_adv_in = fdr::kNin - 2 * fdr_guard_in;
        fdr_guard_out = fdr_guard_in / d->fi.instep * d->fi.outstep;
        fdr_adv_out = fdr_adv_in / d->fi.instep * d->fi.outstep;
        fdr_rl = rl;
        std::vector<float2> tb;
        std::vector<float> hs;
        if (rl == 12) fdr::fdr_make_tables<12>(d->bc.taps, d->bc.klen, tb, hs);
        if (rl == 15) fdr::fdr_make_tables<15>(d->bc.taps, d->bc.klen, tb, hs);
        if (rl == 10) fdr::fdr_make_tables<10>(d->bc.taps, d->bc.klen, tb, hs);
        FMR_CUDA(mem.alloc(&d_fdr_tab, tb.size(), false));
        FMR_CUDA(cudaMemcpy(d_fdr_tab, tb.data(), sizeof(float2) * tb.size(), cudaMemcpyHostToDevice));
        FMR_CUDA(mem.alloc(&d_fdr_Hs, hs.size(), false));
        FMR_CUDA(cudaMemcpy(d_fdr_Hs, hs.data(), sizeof(float) * hs.size(), cudaMemcpyHostToDevice));
        FMR_CUDA(cudaFuncSetAttribute(k_fdr<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, fdr_smem_bytes<12>()));
        FMR_CUDA(cudaFuncSetAttribute(k_fdr<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, fdr_smem_bytes<15>()));
        FMR_CUDA(cudaFuncSetAttribute(k_fdr<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, fdr_smem_bytes<10>()));
        use_fdr = true;
        // fused persistent front end: 10 MHz cf32 chain only (three half-band stages 4/5/8 in front)
        if (lin && hb_stream && rl == 12 && fdr_guard_in == fdr::kGuardIn && !env_off("FMR_FE")) {
          cudaDriverEntryPointQueryResult qr;
          void *fn = nullptr;
          if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn &&
              qr == cudaDriverEntryPointSuccess) {
            tm_encode = reinterpret_cast<TmEncodeFn>(fn);
            CUtensorMap probe;
            if (const char *ev = getenv("FMR_FE_VARIANT")) fe_variant = std::min(2, std::max(0, atoi(ev)));
            if (fe_encode(&probe, r_hb.base, 1, 1, (size_t)fe::kBlockIn)) {
              FMR_CUDA(fe_dispatch([&](auto cf) {
                using CF = decltype(cf);
                return cudaFuncSetAttribute(fe::k_frontend_fused<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kSmemBytes);
              }));
              use_fe = true;
              if (const char *ev = getenv("FMR_FE_MIN_BLOCKS")) fe_min_blocks = std::max(1, atoi(ev));
            }
          }
        }
      }
    }
    if (sizeof(S) == sizeof(double) && d->bc.down == 2 && (d->bc.klen + 1) / 2 <= kDecMaxTaps) {
      FMR_CUDA(cudaFuncSetAttribute(k_fir_dec2_f64, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)dec2_smem(d->bc.klen)));
      use_dec2 = true;
    }
    smem_fir = ((size_t)(kFirTile - 1) * d->bc.down + d->bc.klen) * sizeof(V) + (size_t)d->bc.klen * sizeof(S);
    FMR_CUDA(raise_smem_limit(k_fir_long<S>, smem_fir));
    if (d->has_fi) {
      int64_t max_bc = max_hb / d->bc.down + 4;
      // With the polyphase bank fused behind the FFT low-pass the intermediate ring only ever holds
      // what the unfused kernels of SMALL calls produce (fewer than fft_min_out filter outputs) plus
      // the history tail the fused kernel leaves for them.
      if (use_fft && fuse_fi && d->bc.down == 1 && max_bc > fft_min_out + 64) max_bc = fft_min_out + 64;
      r_bc.cap = pow2ceil((uint64_t)(max_bc + 4 * d->fi.flen + 64));
      FMR_CUDA(mem.alloc(&r_bc.base, (size_t)C * r_bc.cap));
      const size_t nt = (size_t)d->fi.outstep * d->fi.flen;
      std::vector<S> h(nt);
      for (size_t i = 0; i < nt; i++) h[i] = (S)d->fi.taps[i];
      FMR_CUDA(mem.alloc(&d_fi, nt, false));
      FMR_CUDA(cudaMemcpy(d_fi, h.data(), nt * sizeof(S), cudaMemcpyHostToDevice));
      smem_fi = fi_smem(d->fi.instep, d->fi.outstep, d->fi.flen, sizeof(V), sizeof(S));
      FMR_CUDA(raise_smem_limit(k_frac_interp<S>, smem_fi));
    }
    return FMR_OK;
  }

  int64_t max_out(int64_t max_in) const { return chain_out(d, max_in + (int64_t)1) + 8; }

  static bool env_off(const char *name) {
    const char *e = getenv(name);
    return e && atoi(e) == 0;
  }
  // Block plan of one FFT launch group: `per16` / `per8` results per 16384- / 8192-point block;
  // full 16384 blocks first, the remainder in one more 16384 block or in 8192 blocks, whichever
  // is less work. Double chains only have the 8192-point kernel.
  static void fft_plan(int n, int per16, int per8, bool have16, int *nb16, int *nb8) {
    *nb16 = 0;
    *nb8 = 0;
    if (!have16 || per16 <= 0) {
      *nb8 = (n + per8 - 1) / per8;
      return;
    }
    *nb16 = n / per16;
    const int rem = n - *nb16 * per16;
    if (rem > per8) {
      (*nb16)++;
    } else if (rem > 0) {
      *nb8 = 1;
    }
  }
  // plain form: filter outputs [q0, q0+n) -> o
  int launch_fft(Ring<V> in, Ring<V> o, int64_t q0, int n, int64_t avail, cudaStream_t st) {
    const int klen = d->bc.klen, down = d->bc.down;
    const int lq16 = (16384 - klen + 1) / down, lq8 = (8192 - klen + 1) / down;
    int nb16, nb8;
    fft_plan(n, lq16, lq8, sizeof(S) == sizeof(float), &nb16, &nb8);
    FftFuse fz;
    memset(&fz, 0, sizeof(fz));
    int launched = 0;
    int done = 0;
    if constexpr (sizeof(S) == sizeof(float)) {
      if (nb16 > 0) {
        const int cnt = std::min(n, nb16 * lq16);
        dim3 grid(nb16, gcn);
        k_fir_fft<S, 16384, false><<<grid, kFftThreads, FftCfg<S, 16384>::kSmemBytes, st>>>(in, o, d_H16, klen, down, q0,
                                                                                          cnt, avail, lq16, fz);
        done = cnt;
        launched++;
      }
    }
    if (nb8 > 0 && done < n) {
      dim3 grid(nb8, gcn);
      k_fir_fft<S, 8192, false><<<grid, kFftThreads, FftCfg<S, 8192>::kSmemBytes, st>>>(in, o, d_H8, klen, down, q0 + done,
                                                                                      n - done, avail, lq8, fz);
      launched++;
    }
    return launched;
  }
  // fused form: interpolator outputs [m0, m0+n_m) -> o straight from the filtered blocks; the
  // newest filtered samples before `b1` also go to the intermediate ring (see FftFuse).
  int launch_fft_fused(Ring<V> in, Ring<V> o, int64_t m0, int n_m, int64_t avail, int64_t b1, cudaStream_t st) {
    const int klen = d->bc.klen;
    const int lq16 = 16384 - klen + 1, lq8 = 8192 - klen + 1;
    auto per = [&](int lq) { return (int)(((int64_t)(lq - d->fi.flen - 8) * d->fi.outstep) / d->fi.instep); };
    const int mo16 = per(lq16), mo8 = per(lq8);
    int nb16, nb8;
    fft_plan(n_m, mo16, mo8, sizeof(S) == sizeof(float), &nb16, &nb8);
    FftFuse fz;
    memset(&fz, 0, sizeof(fz));
    fz.bank = d_fi;
    fz.instep = d->fi.instep;
    fz.outstep = d->fi.outstep;
    fz.flen = d->fi.flen;
    const Ring<V> tail = sub(r_bc);
    fz.tail_base = tail.base;
    fz.tail_cap = tail.cap;
    int launched = 0, done = 0;
    if constexpr (sizeof(S) == sizeof(float)) {
      if (nb16 > 0) {
        const int cnt = std::min(n_m, nb16 * mo16);
        fz.m0 = m0;
        fz.n_m = cnt;
        fz.mo = mo16;
        const bool last = (cnt == n_m);
        fz.tail_hi = last ? b1 : 0;
        fz.tail_lo = last ? b1 - (2 * d->fi.flen + 16) : 0;
        dim3 grid(nb16, gcn);
        if (fft_inplace) {
          fz.twtab = d_iptab;
          k_fir_fft_ip<512><<<grid, 512, kIpSmemBytes, st>>>(in, o, reinterpret_cast<const float2 *>(d_H16rev), klen, avail, fz);
        } else if (fft_tw) {
          fz.twtab = d_twtab;
          k_fir_fft<S, 16384, true, true><<<grid, kFftThreads, FftCfg<S, 16384>::kSmemBytesTw, st>>>(in, o, d_H16, klen, 1,
                                                                                                   0, 0, avail, lq16, fz);
        } else {
          k_fir_fft<S, 16384, true><<<grid, kFftThreads, FftCfg<S, 16384>::kSmemBytes, st>>>(in, o, d_H16, klen, 1, 0, 0,
                                                                                           avail, lq16, fz);
        }
        done = cnt;
        launched++;
      }
    }
    if (nb8 > 0 && done < n_m) {
      fz.m0 = m0 + done;
      fz.n_m = n_m - done;
      fz.mo = mo8;
      fz.tail_hi = b1;
      fz.tail_lo = b1 - (2 * d->fi.flen + 16);
      dim3 grid(nb8, gcn);
      k_fir_fft<S, 8192, true><<<grid, kFftThreads, FftCfg<S, 8192>::kSmemBytes, st>>>(in, o, d_H8, klen, 1, 0, 0, avail,
                                                                                     lq8, fz);
      launched++;
    }
    return launched;
  }
  template <typename F> auto fe_dispatch(F &&f) const {
    if (fe_variant == 1) return f(fe::CfgS{});
    if (fe_variant == 2) return f(fe::CfgA{});
    return f(fe::CfgC4{});
  }
  int fe_in_span() const {
    return fe_dispatch([](auto cf) { return decltype(cf)::kInSpan; });
  }
  // Tensor map of the fused front end's input (fmr_frontend.cuh): element (f, t, q, k, c) = float f of 128-byte chunk q
  // of stream tile t of block k of channel c; `first` = input sample (row 0, tile 0, block 0, channel 0). Rows of
  // neighbouring tiles overlap in memory (a tile re-reads the end of its predecessor's range as warm-up).
  bool fe_encode(CUtensorMap *tm, const void *first, int n_blocks, int n_ch, size_t stride_samples) const {
    if (!tm_encode) return false;
    return fe_dispatch([&](auto cf) {
      using CF = decltype(cf);
      const cuuint64_t dims[5] = {32, (cuuint64_t)CF::kTiles, (cuuint64_t)CF::kRowChunks, (cuuint64_t)n_blocks, (cuuint64_t)n_ch};
      const cuuint64_t strides[4] = {(cuuint64_t)CF::kTileIn * 8, 128, (cuuint64_t)fe::kBlockIn * 8, (cuuint64_t)stride_samples * 8};
      const cuuint32_t box[5] = {32, (cuuint32_t)CF::kWarpTiles, (cuuint32_t)CF::kRowSteps, 1, 1}, es[5] = {1, 1, 1, 1, 1};
      return tm_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void *>(first), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    });
  }
  // blocks [ja, jb] of the absolute grid, every channel, in the fused kernel
  bool launch_fe(const InSrc<float2> &src, Ring<float2> hb, Ring<float2> o, int64_t ja, int64_t jb, cudaStream_t st) {
    CUtensorMap tm;
    const float2 *first = src.lin + (ja * fe::kBlockIn + fe::kIn0 - src.start);
    if (!fe_encode(&tm, first, (int)(jb - ja + 1), gcn, src.stride)) return false;
    fe::Params P;
    P.hb_ring = hb.base;
    P.hb_cap = hb.cap;
    P.out = o.base;
    P.out_cap = o.cap;
    P.Hs = d_fdr_Hs;
    P.tab = d_fdr_tab;
    P.j0 = ja;
    P.n_blocks = (int)(jb - ja + 1);
    P.n_channels = gcn;
    P.ring_from = (jb + 1) * fdr::kAdvIn - fdr::kGuardIn;
    for (int k = 0; k < 8; k++) {
      P.t1[k] = hbt.t[0][k];
      P.t2[k] = hbt.t[1][k];
      P.t3[k] = hbt.t[2][k];
    }
    fe_dispatch([&](auto cf) {
      using CF = decltype(cf);
      fe::k_frontend_fused<CF><<<std::min(sm_count, gcn), CF::kThreads, CF::kSmemBytes, st>>>(tm, P);
      return 0;
    });
    return true;
  }
  bool launch_fe(const InSrc<double2> &, Ring<double2>, Ring<double2>, int64_t, int64_t, cudaStream_t) { return false; }
  // frequency-domain form (fmr_fdr.cuh): blocks [j0, j1] of the absolute grid, whole
  void launch_fdr(Ring<float2> in, Ring<float2> o, int64_t j0, int64_t j1, int64_t avail, cudaStream_t st, int64_t m_lo = -1,
                  int64_t m_hi = -1) {
    FdrParams P;
    P.j0 = j0;
    P.m_lo = (m_lo >= 0) ? m_lo : j0 * fdr_adv_out;
    P.m_hi = (m_hi >= 0) ? m_hi : (j1 + 1) * fdr_adv_out;
    P.avail = avail;
    P.adv_in = fdr_adv_in;
    P.guard_in = fdr_guard_in;
    P.adv_out = fdr_adv_out;
    P.guard_out = fdr_guard_out;
    dim3 grid((unsigned)(j1 - j0 + 1), gcn);
    if (fdr_rl == 12) k_fdr<12><<<grid, kFdrThreads, fdr_smem_bytes<12>(), st>>>(in, o, d_fdr_Hs, d_fdr_tab, P);
    if (fdr_rl == 15) k_fdr<15><<<grid, kFdrThreads, fdr_smem_bytes<15>(), st>>>(in, o, d_fdr_Hs, d_fdr_tab, P);
    if (fdr_rl == 10) k_fdr<10><<<grid, kFdrThreads, fdr_smem_bytes<10>(), st>>>(in, o, d_fdr_Hs, d_fdr_tab, P);
  }
  void launch_fdr(Ring<double2>, Ring<double2>, int64_t, int64_t, int64_t, cudaStream_t, int64_t = -1, int64_t = -1) {}
  void launch_dec2(Ring<double2> in, Ring<double2> o, int64_t q0, int n, cudaStream_t st) {
    dim3 grid((n + kDecTile - 1) / kDecTile, gcn);
    k_fir_dec2_f64<<<grid, kDecThreads, dec2_smem(d->bc.klen), st>>>(in, o, d_bc, d->bc.klen, q0, n);
  }
  void launch_dec2(Ring<float2>, Ring<float2>, int64_t, int, cudaStream_t) {}

  // Streaming half-band cascade over the part [*sa, return) of [h0, h1) whose input lies entirely in
  // the caller's buffer of this call (the tiled kernel does the few outputs before and after it,
  // which need the history buffer or are not a whole stream tile). Returns *sa when not applicable.
  int64_t hbs_launch(const InSrc<float2> &src, Ring<float2> o, int64_t h0, int64_t h1, int64_t *sa, cudaStream_t st,
                     int *launches) {
    using D = HbsDelays<4, 5, 8>;
    *sa = h0;
    if ((src.start & 1) || (src.stride & 1) || (reinterpret_cast<uintptr_t>(src.lin) & 15)) return h0;
    // first even output whose warm-up chunks start inside the buffer
    const int64_t c0 = (src.start + 15) / 16; // first whole 16-sample chunk of the buffer
    int64_t a = 2 * (c0 + D::kWarm) - D::A3;
    if (a < h0) a = h0;
    a += (a & 1);
    const int64_t avail = h1 - a;
    if (avail < 64) return h0;
    // largest tile that still fills every SM three CTAs deep; warm-up costs 2*kWarm outputs per tile
    int tile = 512;
    while (tile > 128 && (int64_t)gcn * (avail / tile) < (int64_t)sm_count * 3 * kHbsThreads) tile >>= 1;
    tile = std::max(2 * kHbsU, tile / (2 * kHbsU) * (2 * kHbsU));
    const int nbs = (D::kWarm + tile / 2 + kHbsU - 1) / kHbsU;
    // last chunk (exclusive) a tile starting at output m touches: (m + A3)/2 - kWarm + nbs*U
    const int64_t c_end = (src.start + src.n_new) / 16; // chunks [.., c_end) are complete
    // m_last + A3 <= 2 * (c_end - nbs*U + kWarm)
    const int64_t m_last_max = 2 * (c_end - (int64_t)nbs * kHbsU + D::kWarm) - D::A3;
    if (m_last_max < a) return h0;
    int64_t tiles = (m_last_max - a) / tile + 1;
    if (tiles > avail / tile) tiles = avail / tile;
    if (tiles <= 0) return h0;
    HbsParams P;
    P.lin = src.lin;
    P.stride = src.stride;
    P.start = src.start;
    P.a_out = a;
    P.tile = tile;
    P.tiles_per_ch = (int)tiles;
    P.n_streams = (int)(tiles * gcn);
    P.n_block_steps = nbs;
    for (int k = 0; k < 8; k++) {
      P.t1[k] = hbt.t[0][k];
      P.t2[k] = hbt.t[1][k];
      P.t3[k] = hbt.t[2][k];
    }
    const int grid = (P.n_streams + kHbsThreads - 1) / kHbsThreads;
    P.l2_prefetch = 0;
    if (hbs_tma) {
      k_hb_stream_tma<4, 5, 8, kHbsU, 2, 3, 4><<<(P.n_streams + 127) / 128, 128, HbsTma<2>::smem_bytes(3, 4), st>>>(P, o);
    } else {
      k_hb_stream<4, 5, 8, kHbsU, 4><<<grid, kHbsThreads, hbs_smem_bytes(4), st>>>(P, o);
    }
    (*launches)++;
    *sa = a;
    return a + tiles * tile;
  }

  // Consume n_new more input samples; produce the reference's output index range into `out`.
  fmr_status run(InSrc<V> src, int64_t n_new, Ring<V> out, int fs4, cudaStream_t st, int64_t *o0,
                 int64_t *o1, int *launches, bool advance = true) {
    if (gcn == 0) {
      gc0 = 0;
      gcn = C;
    }
    const int64_t N0 = cum_in, N1 = cum_in + n_new;
    const int64_t h0 = hb_out(d, N0), h1 = hb_out(d, N1);
    const int64_t b0 = bc_out(d, h0), b1 = bc_out(d, h1);
    const int64_t f0 = fi_out(d, b0), f1 = fi_out(d, b1);
    Ring<V> bc_in = src.ring;
    // ---- which blocks of the frequency-domain resampler's grid this call computes, and which of them the fused
    // front end takes (fmr_frontend.cuh): those whose whole 10 MHz input lies in this call's buffer
    int64_t j_last = -1, ja = 0, jb = -1, j_part = -1;
    if (use_fdr) {
      // every block whose 10000 input samples are complete (and whose outputs fit the output ring) is computed as a
      // whole, possibly ahead of the reference's release schedule, which lags the input by more than a block (the
      // block convolver's latency of 15231 samples), so the released range [f0, f1) is always covered
      const int64_t j_in = (h1 >= fdr::kNin - fdr_guard_in) ? (h1 - (fdr::kNin - fdr_guard_in)) / fdr_adv_in : -1;
      const int64_t j_cap = (f0 - fdr_hist + (int64_t)out.cap) / fdr_adv_out - 1;
      j_last = std::min(j_in, j_cap);
      // Chains whose block convolver withholds less than a block (1 MHz: 7269 samples) release outputs of the block that
      // is still filling: that block is then evaluated on what has arrived (zeros behind it, which only its not yet
      // released outputs can see) and only [.., f1) is stored; it is evaluated again, whole, when it is complete.
      if ((j_last + 1) * fdr_adv_out < f1) {
        j_part = j_last + 1;
        if ((j_part + 1) * fdr_adv_out < f1 || j_part > j_cap) {
          return fail(FMR_ERR_INVALID, "internal: block grid behind the release schedule");
        }
      }
      if (use_fe && j_last >= fdr_next && src.fmt == 0 && !fs4 && !(src.start & 1) && !(src.stride & 1) &&
          !(reinterpret_cast<uintptr_t>(src.lin) & 15)) {
        auto cdiv = [](int64_t a, int64_t b) { return (a >= 0) ? (a + b - 1) / b : -((-a) / b); };
        auto fdiv = [](int64_t a, int64_t b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
        ja = std::max(fdr_next, cdiv(src.start - fe::kIn0, fe::kBlockIn));
        jb = std::min(j_last, fdiv(src.start + src.n_new - fe_in_span(), fe::kBlockIn));
        if (jb - ja + 1 < fe_min_blocks) jb = ja - 1;
      }
    }
    const bool fused = jb >= ja;
    if (use_fdr) {
      const int64_t n_all = std::max<int64_t>(0, j_last - fdr_next + 1), n_fe = fused ? jb - ja + 1 : 0;
      last_plan[0] = (uint64_t)n_fe;
      last_plan[1] = (uint64_t)(n_all - n_fe);
      last_plan[2] = 0;
      last_plan[3] = (uint64_t)fdr_adv_in << d->n_hb;
      last_plan[4] = (uint64_t)fdr_adv_out;
    }
    if (d->n_hb > 0 || linear_in) {
      const int n = (int)(h1 - h0);
      if (n > 0) {
        if (prof) prof->begin(p_hb, st);
        const Ring<V> hb_out_ring = sub(r_hb);
        const HbTaps<S> tp = hbt;
        const size_t sm = smem_hb;
        auto tiled = [&](int64_t o0, int cnt) {
          if (cnt <= 0) return;
          const int tile = hb_tile_of(hbt.n[0], hbt.n[1]);
          dim3 grid((cnt + tile - 1) / tile, gcn);
          hb_dispatch([&](auto kern) { kern<<<grid, kHbThreads, sm, st>>>(src, hb_out_ring, tp, o0, cnt, fs4); });
          (*launches)++;
        };
        auto hb_range = [&](int64_t lo, int64_t hi) {
          if (hi <= lo) return;
          last_plan[2] += (uint64_t)(hi - lo);
          int64_t sa = lo, sb = lo; // [sa, sb): outputs the streaming kernel produces
          if constexpr (sizeof(S) == sizeof(float)) {
            if (hb_stream && src.fmt == 0 && !fs4) sb = hbs_launch(src, hb_out_ring, lo, hi, &sa, st, launches);
          }
          tiled(lo, (int)(sa - lo));
          tiled(sb, (int)(hi - sb));
        };
        // With fused blocks [ja, jb] only the ends of the 1.25 MHz stream go to the ring: what the unfused blocks before
        // ja read (which includes the 2500 samples block ja shares with its predecessor), and everything from the first
        // sample of block jb + 1 on (history of the next call).
        const int64_t head_hi = fused ? std::min(h1, ja * fdr::kAdvIn + fdr::kGuardIn) : h1;
        // (the fused kernel itself writes the last block's final 2500 samples to the ring: the tail starts behind them)
        const int64_t tail_lo = fused ? std::max(h0, (jb + 1) * fdr::kAdvIn + fdr::kGuardIn) : h1;
        if (fused && head_hi < tail_lo) {
          hb_range(h0, head_hi);
          hb_range(tail_lo, h1);
        } else {
          hb_range(h0, h1);
        }
        if (prof) prof->end(p_hb, st);
      }
      bc_in = sub(r_hb);
    }
    const int n_bc = (int)(b1 - b0), n_fi = (int)(f1 - f0);
    if (use_fdr) {
      // Frequency-domain low-pass + resampling on the absolute block grid (fmr_fdr.cuh), for calls of any size
      auto unfused = [&](int64_t j0, int64_t j1) {
        if (j1 < j0) return;
        // (outputs below fdr_out_done were stored from the block's partial evaluation and have been consumed: keep them)
        launch_fdr(bc_in, out, j0, j1, h1, st, std::max(j0 * (int64_t)fdr_adv_out, fdr_out_done), -1);
        (*launches)++;
      };
      if (j_last >= fdr_next || j_part >= 0) {
        if (prof) prof->begin(p_bc, st);
        unfused(fdr_next, fused ? ja - 1 : j_last);
        if (fused) unfused(jb + 1, j_last);
        if (j_part >= 0) {
          launch_fdr(bc_in, out, j_part, j_part, h1, st, std::max(j_part * (int64_t)fdr_adv_out, fdr_out_done), f1);
          (*launches)++;
        }
        if (prof) prof->end(p_bc, st);
        if (fused) {
          if (prof) prof->begin(p_fe, st);
          if (!launch_fe(src, bc_in, out, ja, jb, st)) return fail(FMR_ERR_CUDA, "tensor map of the fused front end rejected");
          (*launches)++;
          if (prof) prof->end(p_fe, st);
        }
        if (fdr_next == 0) { // (after the block kernels: it overwrites their first outputs)
          // Stream start: the reference's bank starts on a zero-initialised delay line, i.e. it sees zeros instead of the
          // zero-phase filter's pre-ringing in front of filter output 0 (CDSPFracInterpolator.h:861-925). That only
          // shows in the outputs whose window begins before it: redo those few with the time-domain kernels.
          const int half = d->fi.flen / 2 - 1;
          const int n_m = (half * d->fi.outstep + d->fi.instep - 1) / d->fi.instep; // first m with floor(m I / O) >= half
          const int n_y = (int)(((int64_t)(n_m - 1) * d->fi.instep) / d->fi.outstep) - half + d->fi.flen + 1;
          dim3 g1((n_y + kFirTile - 1) / kFirTile, gcn);
          k_fir_long<S><<<g1, kFirThreads, smem_fir, st>>>(bc_in, sub(r_bc), d_bc, d->bc.klen, d->bc.down, 0, n_y);
          dim3 g2((n_m + kFiTile - 1) / kFiTile, gcn);
          k_frac_interp<S><<<g2, kFiThreads, smem_fi, st>>>(sub(r_bc), out, d_fi, d->fi.instep, d->fi.outstep, d->fi.flen, 0, n_m);
          (*launches) += 2;
        }
      }
      if (advance) {
        const bool first = (fdr_next == 0);
        fdr_next = std::max(fdr_next, j_last + 1);
        fdr_out_done = std::max(fdr_out_done, (j_part >= 0) ? f1 : fdr_next * (int64_t)fdr_adv_out);
        (void)first;
      }
    } else if (d->has_fi && use_fft && fuse_fi && d->bc.down == 1 && n_bc >= fft_min_out && n_fi > 0) {
      if (prof) prof->begin(p_bc, st);
      (*launches) += launch_fft_fused(bc_in, out, f0, n_fi, h1, b1, st);
      if (prof) prof->end(p_bc, st);
    } else {
      if (n_bc > 0) {
        if (prof) prof->begin(p_bc, st);
        if (use_fft && n_bc >= fft_min_out) {
          (*launches) += launch_fft(bc_in, d->has_fi ? sub(r_bc) : out, b0, n_bc, h1, st);
        } else if (use_dec2) {
          launch_dec2(bc_in, d->has_fi ? sub(r_bc) : out, b0, n_bc, st);
          (*launches)++;
        } else {
          dim3 grid((n_bc + kFirTile - 1) / kFirTile, gcn);
          k_fir_long<S><<<grid, kFirThreads, smem_fir, st>>>(bc_in, d->has_fi ? sub(r_bc) : out, d_bc, d->bc.klen,
                                                              d->bc.down, b0, n_bc);
          (*launches)++;
        }
        if (prof) prof->end(p_bc, st);
      }
      if (d->has_fi && n_fi > 0) {
        dim3 grid((n_fi + kFiTile - 1) / kFiTile, gcn);
        if (prof) prof->begin(p_fi, st);
        k_frac_interp<S><<<grid, kFiThreads, smem_fi, st>>>(sub(r_bc), out, d_fi, d->fi.instep, d->fi.outstep,
                                                            d->fi.flen, f0, n_fi);
        if (prof) prof->end(p_fi, st);
        (*launches)++;
      }
    }
    if (advance) cum_in = N1;
    *o0 = f0;
    *o1 = f1;
    FMR_CUDA(cudaGetLastError());
    return FMR_OK;
  }
};

} // namespace fmr
#endif
