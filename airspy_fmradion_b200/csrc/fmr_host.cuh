// fmr_host.cuh — host-side plumbing shared by the FM and AM handles: error reporting,
// device buffers, the generic streaming resampler (half-band cascade -> long low-pass ->
// polyphase bank) that stands in for r8b::CDSPResampler (IfResampler.cpp:25-79,
// AudioResampler.cpp:25-61), and pinned staging for the per-call tables.
#ifndef FMR_HOST_CUH
#define FMR_HOST_CUH

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fmradion_b200.h"
#include <cmath>
#include <complex>

#include "fmr_fft.cuh"
#include "fmr_kernels.cuh"
#include "fmr_tables.h"

namespace fmr {

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
  void reset() {
    for (int i = 0; i < n; i++) ran[i] = false;
  }
  void begin(int i, cudaStream_t st) {
    if (on && i >= 0) cudaEventRecord(a[i], st);
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
      ms[k] = t;
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

// Outputs per launch below which the direct-form kernel is used instead of the FFT one.
constexpr int kFftMinOut = 3000;

template <typename S> struct Resampler {
  using V = typename V2<S>::type;
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
  float2 *d_H = nullptr; // filter spectrum for k_fir_fft (float chains only)
  bool use_fft = false;
  bool use_dec2 = false; // double chains with a decimate-by-2 low-pass (audio resampler)

  static size_t hb_smem(const HbTaps<S> &t, int nst) {
    size_t total = 0;
    // levels 0 and 1 are live together; level 2 aliases level 0 (see k_hb_cascade)
    for (int s = 0; s < nst && s < 2; s++) {
      total += 2 * (size_t)kHbR * hb_sub_len(hb_level_len(t.n, nst, s), (int)sizeof(V));
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
    for (int s = 0; s < d->n_hb; s++) hbt.sl[s] = hb_sub_len(hb_level_len(hbt.n, d->n_hb, s), (int)sizeof(V));
    smem_hb = hb_smem(hbt, d->n_hb);
    cudaError_t e = cudaSuccess;
    const size_t smem_need = smem_hb;
    if (!hb_dispatch([&](auto kern) {
          e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need);
        })) {
      return fail(FMR_ERR_UNSUPPORTED, "half-band tap combination not instantiated");
    }
    FMR_CUDA(e);
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
    if (sizeof(S) == sizeof(float) && d->bc.klen < kFftN / 2) {
      std::vector<std::complex<double>> hc(kFftN, std::complex<double>(0.0, 0.0));
      for (int i = 0; i < d->bc.klen; i++) hc[i] = d->bc.taps[i];
      host_fft(hc);
      std::vector<float2> hf(kFftN);
      for (int i = 0; i < kFftN; i++) {
        hf[i] = make_float2((float)(hc[i].real() / kFftN), (float)(hc[i].imag() / kFftN));
      }
      FMR_CUDA(mem.alloc(&d_H, (size_t)kFftN, false));
      FMR_CUDA(cudaMemcpy(d_H, hf.data(), sizeof(float2) * kFftN, cudaMemcpyHostToDevice));
      FMR_CUDA(cudaFuncSetAttribute(k_fir_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, kFftSmemBytes));
      use_fft = true;
    }
    if (sizeof(S) == sizeof(double) && d->bc.down == 2 && (d->bc.klen + 1) / 2 <= kDecMaxTaps) {
      FMR_CUDA(cudaFuncSetAttribute(k_fir_dec2_f64, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)dec2_smem(d->bc.klen)));
      use_dec2 = true;
    }
    smem_fir = ((size_t)(kFirTile - 1) * d->bc.down + d->bc.klen) * sizeof(V) + (size_t)d->bc.klen * sizeof(S);
    FMR_CUDA(cudaFuncSetAttribute(k_fir_long<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fir));
    if (d->has_fi) {
      const int64_t max_bc = max_hb / d->bc.down + 4;
      r_bc.cap = pow2ceil((uint64_t)(max_bc + 4 * d->fi.flen + 64));
      FMR_CUDA(mem.alloc(&r_bc.base, (size_t)C * r_bc.cap));
      const size_t nt = (size_t)d->fi.outstep * d->fi.flen;
      std::vector<S> h(nt);
      for (size_t i = 0; i < nt; i++) h[i] = (S)d->fi.taps[i];
      FMR_CUDA(mem.alloc(&d_fi, nt, false));
      FMR_CUDA(cudaMemcpy(d_fi, h.data(), nt * sizeof(S), cudaMemcpyHostToDevice));
      smem_fi = fi_smem(d->fi.instep, d->fi.outstep, d->fi.flen, sizeof(V), sizeof(S));
      FMR_CUDA(cudaFuncSetAttribute(k_frac_interp<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fi));
    }
    return FMR_OK;
  }

  int64_t max_out(int64_t max_in) const { return chain_out(d, max_in + (int64_t)1) + 8; }

  void launch_fft(Ring<float2> in, Ring<float2> o, int64_t q0, int n, int64_t avail, cudaStream_t st) {
    const int lq = (kFftN - d->bc.klen + 1) / d->bc.down;
    dim3 grid((n + lq - 1) / lq, gcn);
    k_fir_fft<<<grid, kFftThreads, kFftSmemBytes, st>>>(in, o, d_H, d->bc.klen, d->bc.down, q0, n, avail, lq);
  }
  void launch_fft(Ring<double2>, Ring<double2>, int64_t, int, int64_t, cudaStream_t) {}
  void launch_dec2(Ring<double2> in, Ring<double2> o, int64_t q0, int n, cudaStream_t st) {
    dim3 grid((n + kDecTile - 1) / kDecTile, gcn);
    k_fir_dec2_f64<<<grid, kDecThreads, dec2_smem(d->bc.klen), st>>>(in, o, d_bc, d->bc.klen, q0, n);
  }
  void launch_dec2(Ring<float2>, Ring<float2>, int64_t, int, cudaStream_t) {}

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
    if (d->n_hb > 0 || linear_in) {
      const int n = (int)(h1 - h0);
      if (n > 0) {
        if (prof) prof->begin(p_hb, st);
        {
          dim3 grid((n + kHbTile - 1) / kHbTile, gcn);
          const Ring<V> hb_out_ring = sub(r_hb);
          const HbTaps<S> tp = hbt;
          const size_t sm = smem_hb;
          hb_dispatch([&](auto kern) { kern<<<grid, kHbThreads, sm, st>>>(src, hb_out_ring, tp, h0, n, fs4); });
        }
        (*launches)++;
        if (prof) prof->end(p_hb, st);
      }
      bc_in = sub(r_hb);
    }
    {
      const int n = (int)(b1 - b0);
      if (n > 0) {
        if (prof) prof->begin(p_bc, st);
        if (use_fft && n >= kFftMinOut) {
          launch_fft(bc_in, d->has_fi ? sub(r_bc) : out, b0, n, h1, st);
        } else if (use_dec2) {
          launch_dec2(bc_in, d->has_fi ? sub(r_bc) : out, b0, n, st);
        } else {
          dim3 grid((n + kFirTile - 1) / kFirTile, gcn);
          k_fir_long<S><<<grid, kFirThreads, smem_fir, st>>>(bc_in, d->has_fi ? sub(r_bc) : out, d_bc, d->bc.klen,
                                                              d->bc.down, b0, n);
        }
        if (prof) prof->end(p_bc, st);
        (*launches)++;
      }
    }
    if (d->has_fi) {
      const int n = (int)(f1 - f0);
      if (n > 0) {
        dim3 grid((n + kFiTile - 1) / kFiTile, gcn);
        if (prof) prof->begin(p_fi, st);
        k_frac_interp<S><<<grid, kFiThreads, smem_fi, st>>>(sub(r_bc), out, d_fi, d->fi.instep, d->fi.outstep,
                                                            d->fi.flen, f0, n);
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
