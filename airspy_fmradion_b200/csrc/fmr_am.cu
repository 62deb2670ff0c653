// fmr_am.cu — the 48 kHz decoders: FourthConverterIQ -> IfResampler -> AmDecoder::process for
// ModType::AM (main.cpp:912-971, AmDecode.cpp:96-218) and NbfmDecoder::process for ModType::NBFM
// (main.cpp:959-962, NbfmDecode.cpp:47-96), many channels per launch.
#include <complex>
#include <cmath>

#include "fmr_core.cuh"
#include "fmr_host.cuh"
#include "fmr_io.cuh"

using namespace fmr;

namespace {

struct AmChanState {
  float if_gain;                          // IfSimpleAgc m_current_gain
  float baseband_mean, baseband_level, if_rms;
  float disc_prev;                        // NBFM: PhaseDiscriminator m_save_value
  double af_gain;                         // AfSimpleAgc m_current_gain
  double dc_x1, dc_x2;                    // HighPassFilterIir delay line
  double de_x1;                           // LowPassFilterRC delay line
  unsigned long long decoder_calls;
};

struct AmCoreParams {
  float if_max, if_rate;                  // IfSimpleAgc(1.0, 1000000.0, 0.0003)  AmDecode.cpp:71-77
  double af_max, af_ref, af_rate;         // AfSimpleAgc(1.0, 1.5, 0.6, 0.001)    AmDecode.cpp:54-66
  double b0, b1, b2, a1, a2;              // HighPassFilterIir(60/48000)          AmDecode.cpp:45
  double de_a1, de_b0;                    // LowPassFilterRC(100us * 48 kHz)      AmDecode.cpp:49
  int demod_real;                         // demodulate_dsb (real part) instead of demodulate_am (magnitude)
  int deemph;                             // deemphasis runs for ModType::AM only (AmDecode.cpp:212-214)
  int n_channels;
};

// AmDecoder::process after the channel filter (AmDecode.cpp:153-217): IF RMS, IF AGC,
// envelope detector, DC block, AF AGC, statistics, deemphasis. All of it is a chain of short
// recurrences at 48 kHz, so one lane per channel runs it serially.
__global__ void k_am_core(Ring<float2> in, double *__restrict__ audio, size_t audio_stride,
                          AmChanState *__restrict__ st, const uint32_t *__restrict__ call_end, int n_calls,
                          int64_t t0, AmCoreParams P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_channels) return;
  AmChanState s = st[c];
  double *o = audio + (size_t)c * audio_stride;
  uint32_t prev_end = 0;
  for (int b = 0; b < n_calls; b++) {
    const uint32_t end = call_end[b];
    const int n = (int)(end - prev_end);
    if (n == 0) continue;
    const int64_t tb = t0 + prev_end;
    const uint32_t ob = prev_end;
    prev_end = end;
    s.decoder_calls++;
    float sumsq = 0.f, vsum = 0.f, vsq = 0.f;
    for (int i0 = 0; i0 < n; i0 += kCoreChunk) {
      float2 xin[kCoreChunk];
#pragma unroll
      for (int u = 0; u < kCoreChunk; u++) xin[u] = (i0 + u < n) ? in.ld(c, tb + i0 + u) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kCoreChunk; u++) {
        const int i = i0 + u;
        if (i >= n) break;
        const float2 x = xin[u];
        sumsq += x.x * x.x + x.y * x.y; // Utility::rms_level_sample (AmDecode.cpp:154)
        // IfSimpleAgc::process (IfSimpleAgc.cpp:37-57)
        const float xr = x.x * s.if_gain, xi = x.y * s.if_gain;
        const float nrm = xr * xr + xi * xi;
        const float z = (float)(1.0 + ((double)P.if_rate * (1.0 - (double)nrm)));
        s.if_gain *= z;
        if (!isfinite(s.if_gain)) {
          s.if_gain = 1.0f;
        } else if (s.if_gain > P.if_max) {
          s.if_gain = P.if_max;
        }
        // demodulate_am (AmDecode.cpp:221-226): volk_32fc_magnitude_32f; demodulate_dsb (:229-234): real part
        const float mag = P.demod_real ? xr : sqrtf(xr * xr + xi * xi);
        vsum += mag;
        vsq += mag * mag;
        // DC block (AmDecode.cpp:194; Filter.cpp:243-250)
        const double d0 = (double)mag - (P.a1 * s.dc_x1 + P.a2 * s.dc_x2);
        const double v = P.b0 * d0 + P.b1 * s.dc_x1 + P.b2 * s.dc_x2;
        s.dc_x2 = s.dc_x1;
        s.dc_x1 = d0;
        // AfSimpleAgc::process (AfSimpleAgc.cpp:36-58)
        const double x2 = v * s.af_gain;
        const double y = x2 * P.af_ref;
        const double za = 1.0 + (P.af_rate * (1.0 - (x2 * x2)));
        s.af_gain *= za;
        if (!isfinite(s.af_gain)) {
          s.af_gain = 1.0;
        } else if (s.af_gain > P.af_max) {
          s.af_gain = P.af_max;
        }
        // deemphasis (AmDecode.cpp:212-214), ModType::AM only
        if (P.deemph) {
          const double e0 = y - P.de_a1 * s.de_x1;
          o[ob + i] = P.de_b0 * e0;
          s.de_x1 = e0;
        } else {
          o[ob + i] = y;
        }
      }
    }
    s.if_rms = sqrtf(sumsq / (float)n);
    const float mean = vsum / (float)n, rms = sqrtf(vsq / (float)n);
    s.baseband_mean = (float)(0.95 * (double)s.baseband_mean + 0.05 * (double)mean);
    s.baseband_level = (float)(0.95 * (double)s.baseband_level + 0.05 * (double)rms);
  }
  st[c] = s;
}

// k_am_core_fused — the same block as k_am_core with its three recurrences on three warps of one
// CTA (lane = channel), pipelined over chunks of kCfT samples through shared memory with named
// barriers, exactly like the FM core (fmr_core.cuh):
//   warp 0  IF RMS, IfSimpleAgc, envelope |x|, baseband statistics      (float recurrence)
//   warp 1  DC block biquad, AfSimpleAgc                                (double recurrence)
//   warp 2  deemphasis, store
// k_am_core runs them back to back per sample (~520 cycles of dependent latency); here the step
// time is the longest single recurrence.
struct AmSmem {
  float mag[2][kCfT][32];
  double y[2][kCfT][32];
};

__global__ void __launch_bounds__(96)
    k_am_core_fused(Ring<float2> in, double *__restrict__ audio, size_t audio_stride, AmChanState *__restrict__ st,
                    const uint32_t *__restrict__ call_end, int n_calls, int64_t t0, AmCoreParams P) {
  __shared__ AmSmem S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c_raw = blockIdx.x * 32 + lane;
  const bool act = c_raw < P.n_channels;
  const int c = act ? c_raw : (P.n_channels - 1); // idle lanes shadow the last channel, store nothing
  const int n_total = n_calls ? (int)call_end[n_calls - 1] : 0;
  const int K = (n_total + kCfT - 1) / kCfT;
  const uint32_t t0lo = (uint32_t)t0;
  // link 0: warp 0 -> warp 1 (mag), link 1: warp 1 -> warp 2 (y)
  if (warp == 0) {
    AmChanState *sp = st + c;
    float g = sp->if_gain, if_rms = sp->if_rms, bmean = sp->baseband_mean, blevel = sp->baseband_level;
    unsigned long long calls = sp->decoder_calls;
    float sumsq = 0.f, vsum = 0.f, vsq = 0.f;
    const float2 *__restrict__ irow = in.base + (size_t)c * in.cap;
    const uint32_t imask = in.cap - 1;
    const double rate = (double)P.if_rate;
    CfCalls cl;
    cl.init(call_end, n_calls);
    float2 nx[kCfT];
#pragma unroll
    for (int u = 0; u < kCfT; u++) nx[u] = (u < n_total) ? irow[(t0lo + (uint32_t)u) & imask] : make_float2(0.f, 0.f);
    for (int k = 0; k < K; k++) {
      const int slot = k & 1, p0 = k * kCfT;
      float2 x[kCfT];
#pragma unroll
      for (int u = 0; u < kCfT; u++) x[u] = nx[u];
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        const int p = p0 + kCfT + u;
        nx[u] = (p < n_total) ? irow[(t0lo + (uint32_t)p) & imask] : make_float2(0.f, 0.f);
      }
      cf_sync(cf_empty(0, slot));
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        const int p = p0 + u;
        if (p < n_total) {
          if ((uint32_t)p == cl.end) {
            cl.next();
            calls++;
          }
          sumsq += x[u].x * x[u].x + x[u].y * x[u].y; // Utility::rms_level_sample (AmDecode.cpp:154)
          // IfSimpleAgc::process (IfSimpleAgc.cpp:37-57)
          const float xr = x[u].x * g, xi = x[u].y * g;
          const float nrm = xr * xr + xi * xi;
          const float z = (float)(1.0 + (rate * (1.0 - (double)nrm)));
          g *= z;
          g = isfinite(g) ? ((g > P.if_max) ? P.if_max : g) : 1.0f;
          // demodulate_am (AmDecode.cpp:221-226) / demodulate_dsb (:229-234)
          const float mag = P.demod_real ? xr : sqrtf(xr * xr + xi * xi);
          vsum += mag;
          vsq += mag * mag;
          S.mag[slot][u][lane] = mag;
          if ((uint32_t)(p + 1) == cl.end) {
            const float n = (float)(cl.end - cl.beg);
            if_rms = sqrtf(sumsq / n);
            const float mean = vsum / n, rms = sqrtf(vsq / n);
            bmean = (float)(0.95 * (double)bmean + 0.05 * (double)mean);
            blevel = (float)(0.95 * (double)blevel + 0.05 * (double)rms);
            sumsq = 0.f;
            vsum = 0.f;
            vsq = 0.f;
          }
        }
      }
      cf_arrive(cf_full(0, slot));
    }
    if (act) {
      sp->if_gain = g;
      sp->if_rms = if_rms;
      sp->baseband_mean = bmean;
      sp->baseband_level = blevel;
      sp->decoder_calls = calls;
    }
  } else if (warp == 1) {
    AmChanState *sp = st + c;
    double x1 = sp->dc_x1, x2s = sp->dc_x2, ag = sp->af_gain;
    const double a1 = P.a1, a2 = P.a2, b0 = P.b0, b1 = P.b1, b2 = P.b2;
    const double af_ref = P.af_ref, af_rate = P.af_rate, af_max = P.af_max;
    cf_arrive(cf_empty(0, 0));
    cf_arrive(cf_empty(0, 1));
    for (int k = 0; k < K; k++) {
      const int slot = k & 1, p0 = k * kCfT;
      cf_sync(cf_full(0, slot));
      float m[kCfT];
#pragma unroll
      for (int u = 0; u < kCfT; u++) m[u] = S.mag[slot][u][lane];
      if (k + 2 < K) cf_arrive(cf_empty(0, slot));
      cf_sync(cf_empty(1, slot));
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        if (p0 + u < n_total) {
          // DC block (AmDecode.cpp:194; Filter.cpp:243-250)
          const double d0 = (double)m[u] - (a1 * x1 + a2 * x2s);
          const double v = b0 * d0 + b1 * x1 + b2 * x2s;
          x2s = x1;
          x1 = d0;
          // AfSimpleAgc::process (AfSimpleAgc.cpp:36-58)
          const double xa = v * ag;
          const double y = xa * af_ref;
          const double za = 1.0 + (af_rate * (1.0 - (xa * xa)));
          ag *= za;
          ag = isfinite(ag) ? ((ag > af_max) ? af_max : ag) : 1.0;
          S.y[slot][u][lane] = y;
        }
      }
      cf_arrive(cf_full(1, slot));
    }
    if (act) {
      sp->dc_x1 = x1;
      sp->dc_x2 = x2s;
      sp->af_gain = ag;
    }
  } else {
    AmChanState *sp = st + c;
    double e1 = sp->de_x1;
    const double de_a1 = P.de_a1, de_b0 = P.de_b0;
    double *__restrict__ o = audio + (size_t)c * audio_stride;
    cf_arrive(cf_empty(1, 0));
    cf_arrive(cf_empty(1, 1));
    for (int k = 0; k < K; k++) {
      const int slot = k & 1, p0 = k * kCfT;
      cf_sync(cf_full(1, slot));
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        if (p0 + u < n_total) {
          // deemphasis (AmDecode.cpp:212-214), ModType::AM only
          const double yv = S.y[slot][u][lane];
          if (P.deemph) {
            const double e0 = yv - de_a1 * e1;
            if (act) o[p0 + u] = de_b0 * e0;
            e1 = e0;
          } else if (act) {
            o[p0 + u] = yv;
          }
        }
      }
      if (k + 2 < K) cf_arrive(cf_empty(1, slot));
    }
    if (act) sp->de_x1 = e1;
  }
}

// FineTuner::process (FineTuner.cpp:55-70): out[i] = in[i] * table[(index + i) mod size], complex
// float product with the reference's separate roundings (no FMA contraction). All channels of a
// handle have consumed the same number of samples, so the table index is one scalar per tuner.
__global__ void k_finetune(Ring<float2> in, Ring<float2> out, const float2 *__restrict__ table, int size,
                           uint32_t idx0, int64_t t0, int n) {
  const uint32_t c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 x = in.ld(c, t0 + i);
  const float2 w = table[(idx0 + (uint32_t)i) % (uint32_t)size];
  float2 y;
  y.x = __fsub_rn(__fmul_rn(x.x, w.x), __fmul_rn(x.y, w.y));
  y.y = __fadd_rn(__fmul_rn(x.x, w.y), __fmul_rn(x.y, w.x));
  out.st(c, t0 + i, y);
}

struct NbfmCoreParams {
  float if_max, if_rate;         // IfSimpleAgc(1.0, 100000.0, 0.0001)  NbfmDecode.cpp:43
  float disc_inv_norm, disc_bound; // PhaseDiscriminator(freq_dev / 48000) NbfmDecode.cpp:35, PhaseDiscriminator.cpp:27-30
  int n_channels;
};

// NbfmDecoder::process between the IF filter and the audio filter (NbfmDecode.cpp:54-86): IF RMS of
// the filtered block, IF AGC, phase discriminator, float -> double, baseband mean / RMS with their
// EMA. The AGC is a serial recurrence at 48 kHz, so one lane per channel runs the lot; the MPX goes
// out as the real part of a double2 stream for the audio FIR kernel.
__global__ void k_nbfm_core(Ring<float2> in, Ring<double2> bb, AmChanState *__restrict__ st,
                            const uint32_t *__restrict__ call_end, int n_calls, int64_t t0, NbfmCoreParams P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_channels) return;
  AmChanState s = st[c];
  uint32_t prev_end = 0;
  for (int b = 0; b < n_calls; b++) {
    const uint32_t end = call_end[b];
    const int n = (int)(end - prev_end);
    if (n == 0) continue; // main.cpp:933-936: the decoder is not called
    const int64_t tb = t0 + prev_end;
    prev_end = end;
    s.decoder_calls++;
    float sumsq = 0.f, vsum = 0.f, vsq = 0.f;
    for (int i0 = 0; i0 < n; i0 += kCoreChunk) {
      float2 xin[kCoreChunk];
#pragma unroll
      for (int u = 0; u < kCoreChunk; u++) xin[u] = (i0 + u < n) ? in.ld(c, tb + i0 + u) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kCoreChunk; u++) {
        const int i = i0 + u;
        if (i >= n) break;
        const float2 x = xin[u];
        sumsq += x.x * x.x + x.y * x.y; // Utility::rms_level_sample on the filtered block (NbfmDecode.cpp:54)
        // IfSimpleAgc::process (IfSimpleAgc.cpp:37-57)
        const float xr = x.x * s.if_gain, xi = x.y * s.if_gain;
        const float nrm = xr * xr + xi * xi;
        const float z = (float)(1.0 + ((double)P.if_rate * (1.0 - (double)nrm)));
        s.if_gain *= z;
        if (!isfinite(s.if_gain)) {
          s.if_gain = 1.0f;
        } else if (s.if_gain > P.if_max) {
          s.if_gain = P.if_max;
        }
        // PhaseDiscriminator::process (PhaseDiscriminator.cpp:33-46)
        const float ph = atan2f(xi, xr) * P.disc_inv_norm;
        float d = ph - s.disc_prev;
        s.disc_prev = ph;
        if (d > P.disc_bound) d -= 2 * P.disc_bound;
        if (d < -P.disc_bound) d += 2 * P.disc_bound;
        if (isnan(d)) d = 0.0f;
        vsum += d;
        vsq += d * d;
        bb.st(c, tb + i, make_double2((double)d, 0.0));
      }
    }
    s.if_rms = sqrtf(sumsq / (float)n);
    const float mean = vsum / (float)n, rms = sqrtf(vsq / (float)n); // samples_mean_rms, NbfmDecode.cpp:83-86
    s.baseband_mean = (float)(0.95 * (double)s.baseband_mean + 0.05 * (double)mean);
    s.baseband_level = (float)(0.95 * (double)s.baseband_level + 0.05 * (double)rms);
  }
  st[c] = s;
}

// audio = filtered baseband * 10^(-3/20) (NbfmDecode.cpp:92-96), thread per output sample
__global__ void k_nbfm_out(Ring<double2> in, double *__restrict__ audio, size_t audio_stride, int64_t t0, int n,
                           double gain) {
  const uint32_t c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  audio[(size_t)c * audio_stride + i] = in.ld(c, t0 + i).x * gain;
}

} // namespace

struct fmr_am {
  fmr_am_config cfg;
  int C = 0;
  const ChainDesc *ifc = nullptr; // null when input_rate == 48000
  DevMem mem;
  PinnedSlots slots;
  cudaStream_t own_stream = nullptr;
  Resampler<float> ifres;
  float2 *hist[2] = {nullptr, nullptr};
  int hist_cur = 0;
  Ring<float2> r_if{nullptr, 0};  // 48 kHz decoder input
  Ring<float2> r_flt{nullptr, 0}; // after the channel filter
  AmChanState *d_state = nullptr;
  uint32_t *d_e48 = nullptr;
  float *d_amfilter = nullptr;
  int amfilter_taps = 0;
  Prof prof;
  int p_hist = -1, p_flt = -1, p_core = -1;
  AmCoreParams core;
  bool core_fused = true; // FMR_CORE_FUSED=0: the single-warp k_am_core
  // USB / LSB / CW / WSPR (AmDecode.cpp:103-137): FineTuner -> 2049-tap filter -> FineTuner
  Ring<float2> r_t1{nullptr, 0}, r_t2{nullptr, 0};
  float *d_cwfilter = nullptr, *d_ssbfilter = nullptr;
  // FFT form of the channel filters (AM / DSB / SSB / CW; not NBFM): spectra of the causal taps padded to 2 K - 1
  float2 *d_H_am = nullptr, *d_H_cw = nullptr, *d_H_ssb = nullptr;
  float c0_am = 0.f, c0_cw = 0.f, c0_ssb = 0.f;
  bool fft_filter = false; // FMR_AM_FFT_FILTER=0: direct form
  float2 *d_tab_cw = nullptr, *d_tab_up = nullptr, *d_tab_down = nullptr; // FineTuner tables (480 entries)
  uint32_t idx_cw = 0, idx_up = 0, idx_down = 0;                         // FineTuner m_index
  int p_tune = -1;
  // NBFM (mode 1)
  bool nbfm = false;
  NbfmCoreParams ncore;
  double freq_dev = 8000.0;
  Ring<double2> r_bb{nullptr, 0};  // discriminator output as double
  Ring<double2> r_aud{nullptr, 0}; // after the audio FIR
  double *d_audiofilter = nullptr;
  int p_aud = -1, p_out = -1;
  int64_t cum_in = 0, cum48 = 0;
  uint32_t last_launches = 0;
  float *d_iq = nullptr;
  double *d_audio = nullptr;
  size_t audio_cap = 0;
  // file-format ingest and output stage (fmr_am_process_*_io, fmr_io.cuh)
  uint8_t *d_raw = nullptr;
  size_t raw_cap = 0;
  uint8_t *d_out = nullptr;
  BlockLevelDev *d_levels = nullptr;
  bool have_levels = false;
  uint32_t last_blocks = 0;
};

static int64_t am_out_total(const ChainDesc *ifc, int64_t n) { return ifc ? chain_out(ifc, n) : n; }

extern "C" fmr_status fmr_am_schedule(double input_rate, uint64_t start_sample, const uint32_t *block_len,
                                      uint32_t n_blocks, uint32_t *audio_len) {
  if (!block_len || !audio_len) return fail(FMR_ERR_INVALID, "null argument");
  const ChainDesc *ifc = nullptr;
  if (input_rate != 48000.0) {
    ifc = find_chain(input_rate, 48000.0, 0);
    if (!ifc) return fail(FMR_ERR_UNSUPPORTED, "no resampler tables for this input_rate -> 48000");
  }
  int64_t n = (int64_t)start_sample;
  int64_t prev = am_out_total(ifc, n);
  for (uint32_t b = 0; b < n_blocks; b++) {
    n += block_len[b];
    const int64_t cur = am_out_total(ifc, n);
    audio_len[b] = (uint32_t)(cur - prev);
    prev = cur;
  }
  return FMR_OK;
}

// FFT form of a causal FIR y[i] = sum_j c[j] x[i - j] for k_fir_fft<float, 8192, false>, which evaluates the centred
// convolution y[q] = sum_j h[j] x[q + (klen - 1) / 2 - j]: h = K - 1 zeros followed by the taps, klen = 2 K - 1. Spectrum
// in double, 1 / N folded in. K <= 2049 (the SSB / CW filters: half of every 8192-point block is payload).
static fmr_status am_make_fft_filter(DevMem &mem, const float *c, int K, float2 **d_H) {
  const int N = 8192;
  if (2 * K - 1 > N / 2 + 1) return fail(FMR_ERR_UNSUPPORTED, "channel filter too long for the FFT form");
  std::vector<std::complex<double>> hc(N, std::complex<double>(0.0, 0.0));
  for (int j = 0; j < K; j++) hc[K - 1 + j] = (double)c[j];
  host_fft(hc);
  std::vector<float2> hf(N);
  for (int i = 0; i < N; i++) hf[i] = make_float2((float)(hc[i].real() / N), (float)(hc[i].imag() / N));
  FMR_CUDA(mem.alloc(d_H, (size_t)N, false));
  FMR_CUDA(cudaMemcpy(*d_H, hf.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
  return FMR_OK;
}

static fmr_status am_build(fmr_am *h) {
  const fmr_am_config &cfg = h->cfg;
  FMR_CUDA(cudaSetDevice(cfg.device));
  const int C = h->C = (int)cfg.n_channels;
  const int64_t max_in = cfg.max_samples_per_call;
  const int max_blocks = (int)cfg.max_blocks_per_call;
  if (cfg.mode < 1 || cfg.mode > 7) return fail(FMR_ERR_UNSUPPORTED, "mode must be a ModType value 1 (NBFM) .. 7 (WSPR)");
  h->nbfm = (cfg.mode == 1);
  if (cfg.input_rate != 48000.0) {
    h->ifc = find_chain(cfg.input_rate, 48000.0, 0);
    if (!h->ifc) return fail(FMR_ERR_UNSUPPORTED, "no resampler tables for this input_rate -> 48000");
  }
  FMR_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  FMR_CUDA(h->slots.init(sizeof(uint32_t) * (size_t)max_blocks));
  int64_t max48 = max_in + 8;
  if (h->ifc) {
    fmr_status s = h->ifres.init(h->ifc, C, max_in, true, h->mem);
    if (s != FMR_OK) return s;
    max48 = (int64_t)std::ceil((double)max_in * 48000.0 / cfg.input_rate) + 8;
  } else {
    HbTaps<float> t;
    memset(&t, 0, sizeof(t));
    FMR_CUDA((cudaFuncSetAttribute(k_hb_cascade<float, 0, true, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)Resampler<float>::hb_smem(t, 0))));
  }
  FMR_CUDA(h->mem.alloc(&h->hist[0], (size_t)C * kHist));
  FMR_CUDA(h->mem.alloc(&h->hist[1], (size_t)C * kHist));
  // + what the frequency-domain resampler (1 MHz -> 48 kHz chain) may produce ahead of the release schedule, and the
  // history the longest channel filter (2049 taps) reads behind it
  h->r_if.cap = pow2ceil((uint64_t)max48 + 2560 + fdr::kMaxAdvOut);
  h->ifres.fdr_hist = 2560;
  FMR_CUDA(h->mem.alloc(&h->r_if.base, (size_t)C * h->r_if.cap));
  h->r_flt.cap = h->r_if.cap;
  FMR_CUDA(h->mem.alloc(&h->r_flt.base, (size_t)C * h->r_flt.cap));
  {
    const float *tbl = k_jj1bdx_am_48khz_default; // main.cpp:785-810
    h->amfilter_taps = 255;
    if (cfg.amfilter == 1) tbl = k_jj1bdx_am_48khz_medium;
    if (cfg.amfilter == 2) tbl = k_jj1bdx_am_48khz_narrow;
    if (cfg.amfilter == 3) {
      tbl = k_jj1bdx_am_48khz_wide;
      h->amfilter_taps = 127;
    }
    if (h->nbfm) {
      h->amfilter_taps = 127;
      tbl = k_jj1bdx_nbfm_48khz_default;
      if (cfg.amfilter == 1) tbl = k_jj1bdx_nbfm_48khz_medium;
      if (cfg.amfilter == 2) tbl = k_jj1bdx_nbfm_48khz_narrow;
      if (cfg.amfilter == 3) tbl = k_jj1bdx_nbfm_48khz_wide;
    }
    if (cfg.amfilter == 4) {
      if (!cfg.amfilter_coeff || cfg.amfilter_ntaps < 2 || cfg.amfilter_ntaps > 4096) {
        return fail(FMR_ERR_INVALID, "amfilter == 4 needs amfilter_coeff with 2..4096 taps");
      }
      tbl = cfg.amfilter_coeff;
      h->amfilter_taps = (int)cfg.amfilter_ntaps;
    }
    FMR_CUDA(h->mem.alloc(&h->d_amfilter, (size_t)h->amfilter_taps, false));
    FMR_CUDA(cudaMemcpy(h->d_amfilter, tbl, h->amfilter_taps * sizeof(float), cudaMemcpyHostToDevice));
    // Long channel filters as overlap-save FFT + the head-loop correction (k_fir_head_fix). Not for NBFM: its
    // discriminator needs the exact zeros the direct form keeps at stream start (atan2(0, 0)).
    h->fft_filter = !h->nbfm && h->amfilter_taps >= 96 && h->amfilter_taps <= 2049 && !Resampler<float>::env_off("FMR_AM_FFT_FILTER");
    if (h->fft_filter) {
      fmr_status sf = am_make_fft_filter(h->mem, tbl, h->amfilter_taps, &h->d_H_am);
      if (sf != FMR_OK) return sf;
      h->c0_am = tbl[0];
      FMR_CUDA((cudaFuncSetAttribute(k_fir_fft<float, 8192, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     FftCfg<float, 8192>::kSmemBytes)));
    }
  }
  FMR_CUDA(raise_smem_limit(k_fir_quirk<float>, fq_smem(h->amfilter_taps, sizeof(float2), sizeof(float))));
  FMR_CUDA(h->mem.alloc(&h->d_state, (size_t)C));
  FMR_CUDA(h->mem.alloc(&h->d_e48, (size_t)max_blocks));
  {
    std::vector<AmChanState> st(C);
    memset(st.data(), 0, sizeof(AmChanState) * C);
    for (int c = 0; c < C; c++) {
      st[c].if_gain = 1.0f;
      st[c].af_gain = 1.0;
    }
    FMR_CUDA(cudaMemcpy(h->d_state, st.data(), sizeof(AmChanState) * C, cudaMemcpyHostToDevice));
  }
  AmCoreParams &P = h->core;
  memset(&P, 0, sizeof(P));
  P.if_max = 1000000.0f;
  P.if_rate = 0.0003f;
  P.af_max = 1.5;
  P.af_ref = 0.6;
  P.af_rate = 0.001;
  {
    // HighPassFilterIir::HighPassFilterIir (Filter.cpp:254-290), cutoff 60/48000
    using CD = std::complex<double>;
    const double w = 2 * M_PI * (60.0 / 48000.0);
    const CD p1s = w / std::exp((2 * 1 + 2 - 1) / double(2 * 2) * CD(0, M_PI));
    const CD p1z = std::exp(p1s);
    const double A1 = -2 * std::real(p1z), A2 = std::abs(p1z * p1z);
    const double g = (1.0 + 2.0 + 1.0) / (1 - A1 + A2);
    P.b0 = 1.0 / g;
    P.b1 = -2.0 / g;
    P.b2 = 1.0 / g;
    P.a1 = A1;
    P.a2 = A2;
  }
  {
    const double tc = 100.0 * 48000.0 * 1.0e-6;
    P.de_a1 = -std::exp(-1 / tc);
    P.de_b0 = 1 + P.de_a1;
  }
  P.n_channels = C;
  {
    // AmDecoder::AmDecoder (AmDecode.cpp:54-77): AGC references and rates depend on the mode
    const int m = cfg.mode;
    const bool ssbcw = (m == 4 || m == 5 || m == 6 || m == 7), cw = (m == 6 || m == 7);
    P.af_ref = ssbcw ? 0.24 : 0.6;
    P.af_rate = cw ? 0.00125 : 0.001;
    P.if_rate = cw ? 0.0006f : 0.0003f;
    P.demod_real = (m != 2) ? 1 : 0;
    P.deemph = (m == 2) ? 1 : 0;
    if (ssbcw) {
      h->r_t1.cap = h->r_if.cap;
      FMR_CUDA(h->mem.alloc(&h->r_t1.base, (size_t)C * h->r_t1.cap));
      h->r_t2.cap = h->r_if.cap;
      FMR_CUDA(h->mem.alloc(&h->r_t2.base, (size_t)C * h->r_t2.cap));
      FMR_CUDA(h->mem.alloc(&h->d_cwfilter, 2049, false));
      FMR_CUDA(cudaMemcpy(h->d_cwfilter, k_jj1bdx_cw_48khz_500hz, 2049 * sizeof(float), cudaMemcpyHostToDevice));
      FMR_CUDA(h->mem.alloc(&h->d_ssbfilter, 2049, false));
      FMR_CUDA(cudaMemcpy(h->d_ssbfilter, k_jj1bdx_ssb_48khz_1500hz, 2049 * sizeof(float), cudaMemcpyHostToDevice));
      FMR_CUDA(raise_smem_limit(k_fir_quirk<float>, fq_smem(2049, sizeof(float2), sizeof(float))));
      if (!Resampler<float>::env_off("FMR_AM_FFT_FILTER")) {
        fmr_status sf = am_make_fft_filter(h->mem, k_jj1bdx_cw_48khz_500hz, 2049, &h->d_H_cw);
        if (sf == FMR_OK) sf = am_make_fft_filter(h->mem, k_jj1bdx_ssb_48khz_1500hz, 2049, &h->d_H_ssb);
        if (sf != FMR_OK) return sf;
        h->c0_cw = k_jj1bdx_cw_48khz_500hz[0];
        h->c0_ssb = k_jj1bdx_ssb_48khz_1500hz[0];
        FMR_CUDA((cudaFuncSetAttribute(k_fir_fft<float, 8192, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FftCfg<float, 8192>::kSmemBytes)));
      }
      // FineTuner::set_freq_shift (FineTuner.cpp:32-52): table_size 480 = 48000/100, shifts +5, +15, -15
      auto make_tab = [&](int shift, float2 **dst) -> cudaError_t {
        std::vector<float2> t(480);
        const double step = 2.0 * M_PI / 480.0;
        for (unsigned i = 0; i < 480; i++) {
          const int64_t r = ((int64_t)shift * (int64_t)i) % (int64_t)480;
          const double phi = (double)r * step;
          t[i] = make_float2((float)std::cos(phi), (float)std::sin(phi));
        }
        cudaError_t e = h->mem.alloc(dst, (size_t)480, false);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, t.data(), sizeof(float2) * 480, cudaMemcpyHostToDevice);
      };
      FMR_CUDA(make_tab(5, &h->d_tab_cw));
      FMR_CUDA(make_tab(15, &h->d_tab_up));
      FMR_CUDA(make_tab(-15, &h->d_tab_down));
    }
  }
  if (h->nbfm) {
    h->freq_dev = cfg.nbfm_freq_dev > 0 ? cfg.nbfm_freq_dev : 8000.0; // NbfmDecoder::freq_dev_normal
    NbfmCoreParams &N = h->ncore;
    N.if_max = 100000.0f;
    N.if_rate = 0.0001f;
    const double max_freq_dev = h->freq_dev / 48000.0; // NbfmDecode.cpp:35
    N.disc_inv_norm = 1.0f / (float)(max_freq_dev * 2.0 * M_PI);
    N.disc_bound = (float)(1.0 / (max_freq_dev * 2.0));
    N.n_channels = C;
    h->r_bb.cap = h->r_if.cap;
    FMR_CUDA(h->mem.alloc(&h->r_bb.base, (size_t)C * h->r_bb.cap));
    h->r_aud.cap = h->r_if.cap;
    FMR_CUDA(h->mem.alloc(&h->r_aud.base, (size_t)C * h->r_aud.cap));
    FMR_CUDA(h->mem.alloc(&h->d_audiofilter, 63, false));
    FMR_CUDA(cudaMemcpy(h->d_audiofilter, k_jj1bdx_48khz_nbfmaudio, 63 * sizeof(double), cudaMemcpyHostToDevice));
    FMR_CUDA(raise_smem_limit(k_fir_quirk<double>, fq_smem(63, sizeof(double2), sizeof(double))));
  }
  h->audio_cap = (size_t)max48;
  h->ifres.prof = &h->prof;
  h->ifres.p_hb = h->prof.add("if_halfband_cascade");
  h->ifres.p_bc = h->prof.add("if_lowpass");
  h->ifres.p_fi = h->prof.add("if_polyphase");
  h->p_hist = h->prof.add("save_hist");
  h->p_flt = h->prof.add("am_channel_filter");
  h->p_tune = h->prof.add("am_finetuners");
  h->p_core = h->prof.add(h->nbfm ? "nbfm_core_48k" : "am_core_48k");
  h->p_aud = h->prof.add("nbfm_audio_fir");
  h->p_out = h->prof.add("nbfm_gain_out");
  return FMR_OK;
}

extern "C" fmr_status fmr_am_create(const fmr_am_config *cfg, fmr_am **out) {
  if (!cfg || !out) return fail(FMR_ERR_INVALID, "null argument");
  if (cfg->n_channels == 0 || cfg->max_samples_per_call == 0 || cfg->max_blocks_per_call == 0) {
    return fail(FMR_ERR_INVALID, "n_channels, max_samples_per_call and max_blocks_per_call must be > 0");
  }
  if (cfg->amfilter < 0 || cfg->amfilter > 4) return fail(FMR_ERR_INVALID, "amfilter must be 0..4");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    return fail(FMR_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  fmr_am *h = new fmr_am();
  h->cfg = *cfg;
  if (const char *e = getenv("FMR_CORE_FUSED")) h->core_fused = atoi(e) != 0;
  fmr_status s = am_build(h);
  if (s != FMR_OK) {
    std::string keep = g_err;
    fmr_am_destroy(h);
    g_err = keep;
    return s;
  }
  *out = h;
  return FMR_OK;
}

extern "C" void fmr_am_destroy(fmr_am *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  h->mem.release();
  h->slots.release();
  h->prof.release();
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

extern "C" fmr_status fmr_am_query_output(fmr_am *h, const uint32_t *block_len, uint32_t n_blocks,
                                          uint64_t *audio_doubles_total, uint32_t *audio_len) {
  if (!h || !block_len) return fail(FMR_ERR_INVALID, "null argument");
  std::vector<uint32_t> tmp(n_blocks);
  fmr_status s = fmr_am_schedule(h->cfg.input_rate, (uint64_t)h->cum_in, block_len, n_blocks, tmp.data());
  if (s != FMR_OK) return s;
  uint64_t tot = 0;
  for (uint32_t b = 0; b < n_blocks; b++) {
    tot += tmp[b];
    if (audio_len) audio_len[b] = tmp[b];
  }
  if (audio_doubles_total) *audio_doubles_total = tot;
  return FMR_OK;
}

extern "C" fmr_status fmr_am_process_device(fmr_am *h, const float *d_iq, size_t iq_stride,
                                            const uint32_t *block_len, uint32_t n_blocks, double *d_audio,
                                            size_t audio_stride, uint32_t *audio_len, void *stream) {
  if (!h || !d_iq || !block_len || !d_audio) return fail(FMR_ERR_INVALID, "null argument");
  if (n_blocks == 0) return FMR_OK;
  if (n_blocks > h->cfg.max_blocks_per_call) return fail(FMR_ERR_CAPACITY, "n_blocks > max_blocks_per_call");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int C = h->C;
  int slot = 0;
  uint32_t *e48 = (uint32_t *)h->slots.acquire(&slot);
  int64_t n = 0;
  const int64_t base48 = am_out_total(h->ifc, h->cum_in);
  for (uint32_t b = 0; b < n_blocks; b++) {
    if (h->ifc && block_len[b] > 65536) return fail(FMR_ERR_INVALID, "block_len > 65536 (IfResampler limit)");
    n += block_len[b];
    e48[b] = (uint32_t)(am_out_total(h->ifc, h->cum_in + n) - base48);
  }
  const uint64_t total_in = (uint64_t)n;
  if (total_in > h->cfg.max_samples_per_call) return fail(FMR_ERR_CAPACITY, "sum(block_len) > max_samples_per_call");
  if (total_in > iq_stride) return fail(FMR_ERR_INVALID, "iq_stride < sum(block_len)");
  const uint32_t n48 = e48[n_blocks - 1];
  if ((size_t)n48 > audio_stride) return fail(FMR_ERR_CAPACITY, "audio_stride too small for this call");
  if (audio_len) {
    for (uint32_t b = 0; b < n_blocks; b++) audio_len[b] = e48[b] - (b ? e48[b - 1] : 0);
  }
  FMR_CUDA(cudaMemcpyAsync(h->d_e48, e48, sizeof(uint32_t) * n_blocks, cudaMemcpyHostToDevice, st));
  h->slots.commit(slot, st);
  int launches = 0;
  Prof &pf = h->prof;
  pf.reset();
  InSrc<float2> src;
  src.lin = reinterpret_cast<const float2 *>(d_iq);
  src.stride = iq_stride;
  src.fmt = 0;
  src.hist = h->hist[h->hist_cur];
  src.start = h->cum_in;
  src.n_new = (int64_t)total_in;
  src.ring = Ring<float2>{nullptr, 0};
  const int64_t t0 = h->cum48;
  if (h->ifc) {
    int64_t o0, o1;
    fmr_status s = h->ifres.run(src, (int64_t)total_in, h->r_if, h->cfg.fs4_shift, st, &o0, &o1, &launches);
    if (s != FMR_OK) return s;
    if (o0 != t0 || o1 != t0 + n48) return fail(FMR_ERR_INVALID, "internal: IF schedule mismatch");
    if (total_in > 0) {
      pf.begin(h->p_hist, st);
      k_save_hist<float2><<<C, 128, 0, st>>>(src.lin, iq_stride, (int64_t)total_in, h->hist[h->hist_cur],
                                             h->hist[h->hist_cur ^ 1], 0);
      pf.end(h->p_hist, st);
      h->hist_cur ^= 1;
      launches++;
    }
  } else if (total_in > 0) {
    HbTaps<float> t;
    memset(&t, 0, sizeof(t));
    dim3 grid((unsigned)((total_in + kHbTile - 1) / kHbTile), C);
    pf.begin(h->ifres.p_hb, st);
    k_hb_cascade<float, 0, true, 0, 0, 0><<<grid, kHbThreads, Resampler<float>::hb_smem(t, 0), st>>>(
        src, h->r_if, t, t0, (int)total_in, h->cfg.fs4_shift);
    pf.end(h->ifres.p_hb, st);
    launches++;
  }
  if (n48 > 0) {
    dim3 grid((n48 + kQTile - 1) / kQTile, C);
    // K-tap causal channel filter as overlap-save FFT blocks (k_fir_fft with the padded taps, klen = 2 K - 1), then the
    // head-loop quirk taken out of the outputs it applies to (LowPassFilterFirIQ::process, Filter.cpp:37-96)
    auto fft_filter = [&](Ring<float2> a, Ring<float2> b, const float2 *H, int K, float c0) {
      const int klen = 2 * K - 1, lq = 8192 - klen + 1;
      FftFuse fz;
      memset(&fz, 0, sizeof(fz));
      dim3 fg((n48 + lq - 1) / lq, C);
      k_fir_fft<float, 8192, false><<<fg, kFftThreads, FftCfg<float, 8192>::kSmemBytes, st>>>(a, b, H, klen, 1, t0, (int)n48,
                                                                                            t0 + (int64_t)n48, lq, fz);
      dim3 hg((n48 + 255) / 256, C);
      k_fir_head_fix<float><<<hg, 256, 0, st>>>(a, b, c0, K - 1, t0, (int)n48, h->d_e48, (int)n_blocks);
      launches++;
    };
    const int mode = h->cfg.mode;
    if (mode >= 4 && mode <= 7) {
      // USB: down, ssb filter, up. LSB: up, ssb filter, down. CW: cw filter, cw tuner. WSPR: down, cw filter, up
      // (AmDecode.cpp:103-137)
      dim3 tg((n48 + 255) / 256, C);
      auto tune = [&](Ring<float2> a, Ring<float2> b, const float2 *tab, uint32_t *idx) {
        k_finetune<<<tg, 256, 0, st>>>(a, b, tab, 480, *idx, t0, (int)n48);
        *idx = (uint32_t)((*idx + n48) % 480u);
        launches++;
      };
      auto filt = [&](Ring<float2> a, Ring<float2> b, const float *taps) {
        const float2 *H = (taps == h->d_ssbfilter) ? h->d_H_ssb : h->d_H_cw;
        if (H) {
          fft_filter(a, b, H, 2049, (taps == h->d_ssbfilter) ? h->c0_ssb : h->c0_cw);
        } else {
          k_fir_quirk<float><<<grid, kQThreads, fq_smem(2049, sizeof(float2), sizeof(float)), st>>>(a, b, taps, 2049, t0, (int)n48,
                                                                                               h->d_e48, (int)n_blocks);
        }
      };
      pf.begin(h->p_tune, st);
      if (mode == 4 || mode == 7) tune(h->r_if, h->r_t1, h->d_tab_down, &h->idx_down);
      if (mode == 5) tune(h->r_if, h->r_t1, h->d_tab_up, &h->idx_up);
      pf.end(h->p_tune, st);
      pf.begin(h->p_flt, st);
      filt(mode == 6 ? h->r_if : h->r_t1, h->r_t2, (mode == 4 || mode == 5) ? h->d_ssbfilter : h->d_cwfilter);
      pf.end(h->p_flt, st);
      if (mode == 4 || mode == 7) tune(h->r_t2, h->r_flt, h->d_tab_up, &h->idx_up);
      if (mode == 5) tune(h->r_t2, h->r_flt, h->d_tab_down, &h->idx_down);
      if (mode == 6) tune(h->r_t2, h->r_flt, h->d_tab_cw, &h->idx_cw);
    } else {
      pf.begin(h->p_flt, st);
      if (h->fft_filter) {
        fft_filter(h->r_if, h->r_flt, h->d_H_am, h->amfilter_taps, h->c0_am);
      } else {
        k_fir_quirk<float><<<grid, kQThreads, fq_smem(h->amfilter_taps, sizeof(float2), sizeof(float)), st>>>(
            h->r_if, h->r_flt, h->d_amfilter, h->amfilter_taps, t0, (int)n48, h->d_e48, (int)n_blocks);
      }
      pf.end(h->p_flt, st);
    }
    pf.begin(h->p_core, st);
    if (h->nbfm) {
      k_nbfm_core<<<(C + 31) / 32, 32, 0, st>>>(h->r_flt, h->r_bb, h->d_state, h->d_e48, (int)n_blocks, t0, h->ncore);
      pf.end(h->p_core, st);
      // LowPassFilterFirAudio (jj1bdx_48khz_nbfmaudio, 63 taps, per-call head loop), NbfmDecode.cpp:89
      pf.begin(h->p_aud, st);
      k_fir_quirk<double><<<grid, kQThreads, fq_smem(63, sizeof(double2), sizeof(double)), st>>>(
          h->r_bb, h->r_aud, h->d_audiofilter, 63, t0, (int)n48, h->d_e48, (int)n_blocks);
      pf.end(h->p_aud, st);
      pf.begin(h->p_out, st);
      dim3 og((n48 + 255) / 256, C);
      k_nbfm_out<<<og, 256, 0, st>>>(h->r_aud, d_audio, audio_stride, t0, (int)n48, std::pow(10.0, (-3.0 / 20.0)));
      pf.end(h->p_out, st);
      launches += 4;
    } else {
      if (h->core_fused) {
        k_am_core_fused<<<(C + 31) / 32, 96, 0, st>>>(h->r_flt, d_audio, audio_stride, h->d_state, h->d_e48,
                                                      (int)n_blocks, t0, h->core);
      } else {
        k_am_core<<<(C + 31) / 32, 32, 0, st>>>(h->r_flt, d_audio, audio_stride, h->d_state, h->d_e48, (int)n_blocks, t0,
                                                h->core);
      }
      pf.end(h->p_core, st);
      launches += 2;
    }
  }
  FMR_CUDA(cudaGetLastError());
  h->cum_in += (int64_t)total_in;
  h->cum48 += n48;
  h->last_launches = (uint32_t)launches;
  return FMR_OK;
}

extern "C" fmr_status fmr_am_process_host(fmr_am *h, const float *iq, size_t iq_stride, const uint32_t *block_len,
                                          uint32_t n_blocks, double *audio, size_t audio_stride,
                                          uint32_t *audio_len) {
  if (!h || !iq || !block_len || !audio) return fail(FMR_ERR_INVALID, "null argument");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_blocks; b++) total += block_len[b];
  if (total > h->cfg.max_samples_per_call) return fail(FMR_ERR_CAPACITY, "sum(block_len) > max_samples_per_call");
  if (total > iq_stride) return fail(FMR_ERR_INVALID, "iq_stride < sum(block_len)");
  const int C = h->C;
  if (!h->d_iq) {
    FMR_CUDA(h->mem.alloc(&h->d_iq, (size_t)C * h->cfg.max_samples_per_call * 2, false));
    FMR_CUDA(h->mem.alloc(&h->d_audio, (size_t)C * h->audio_cap, false));
  }
  cudaStream_t st = h->own_stream;
  if (total > 0) {
    FMR_CUDA(cudaMemcpy2DAsync(h->d_iq, (size_t)total * 8, iq, iq_stride * 8, (size_t)total * 8, C,
                               cudaMemcpyHostToDevice, st));
  }
  uint64_t out_total = 0;
  fmr_status s = fmr_am_query_output(h, block_len, n_blocks, &out_total, nullptr);
  if (s != FMR_OK) return s;
  if (out_total > audio_stride) return fail(FMR_ERR_CAPACITY, "audio_stride too small for this call");
  s = fmr_am_process_device(h, h->d_iq, (size_t)total, block_len, n_blocks, h->d_audio, h->audio_cap, audio_len,
                            (void *)st);
  if (s != FMR_OK) return s;
  if (out_total > 0) {
    FMR_CUDA(cudaMemcpy2DAsync(audio, audio_stride * 8, h->d_audio, h->audio_cap * 8, (size_t)out_total * 8, C,
                               cudaMemcpyDeviceToHost, st));
  }
  FMR_CUDA(cudaStreamSynchronize(st));
  return FMR_OK;
}

static fmr_status am_ensure_staging(fmr_am *h, size_t raw_bytes, bool want_sink) {
  const int C = h->C;
  if (!h->d_iq) {
    FMR_CUDA(h->mem.alloc(&h->d_iq, (size_t)C * h->cfg.max_samples_per_call * 2, false));
    FMR_CUDA(h->mem.alloc(&h->d_audio, (size_t)C * h->audio_cap, false));
  }
  if (raw_bytes > 0 && !h->d_raw) {
    // once, for the widest file format (cf32: 8 bytes per IQ sample): a handle that alternates formats never reallocates
    h->raw_cap = (size_t)C * h->cfg.max_samples_per_call * 8;
    FMR_CUDA(h->mem.alloc(&h->d_raw, h->raw_cap, false));
  }
  if (raw_bytes > h->raw_cap) return fail(FMR_ERR_CAPACITY, "raw staging buffer too small for this call");
  if (want_sink && !h->d_levels) {
    FMR_CUDA(h->mem.alloc(&h->d_out, (size_t)C * h->audio_cap * 8, false));
    FMR_CUDA(h->mem.alloc(&h->d_levels, (size_t)C * h->cfg.max_blocks_per_call));
  }
  return FMR_OK;
}

// FileSource sample formats in, the block loop's output stage out (main.cpp:977-1002), around
// fmr_am_process_device; everything is enqueued on `stream`.
extern "C" fmr_status fmr_am_process_device_io(fmr_am *h, const void *d_iq, int iq_format, size_t iq_stride,
                                               const uint32_t *block_len, uint32_t n_blocks,
                                               const fmr_output_config *out_cfg, void *d_audio, size_t audio_stride,
                                               uint32_t *audio_len, void *stream) {
  if (!h || !d_iq || !block_len || !d_audio) return fail(FMR_ERR_INVALID, "null argument");
  if (iq_format_bytes(iq_format) == 0) return fail(FMR_ERR_INVALID, "unknown iq_format");
  if (out_cfg && out_format_bytes(out_cfg->out_format) == 0) return fail(FMR_ERR_INVALID, "unknown out_format");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_blocks; b++) total += block_len[b];
  if (total > h->cfg.max_samples_per_call) return fail(FMR_ERR_CAPACITY, "sum(block_len) > max_samples_per_call");
  if (total > iq_stride) return fail(FMR_ERR_INVALID, "iq_stride < sum(block_len)");
  if (n_blocks > h->cfg.max_blocks_per_call) return fail(FMR_ERR_CAPACITY, "n_blocks > max_blocks_per_call");
  const bool direct = (iq_format == FMR_IQ_CF32);
  fmr_status s = FMR_OK;
  if (!direct || out_cfg) {
    s = am_ensure_staging(h, 0, out_cfg != nullptr);
    if (s != FMR_OK) return s;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float *d_in = reinterpret_cast<const float *>(d_iq);
  size_t stride = iq_stride;
  uint32_t extra = 0;
  if (!direct) {
    FMR_CUDA(launch_ingest_convert(d_iq, iq_format, iq_stride, reinterpret_cast<float2 *>(h->d_iq), (size_t)total, total,
                                   h->C, st));
    d_in = h->d_iq;
    stride = (size_t)total;
    extra++;
  }
  h->have_levels = false;
  h->last_blocks = n_blocks;
  if (!out_cfg) {
    s = fmr_am_process_device(h, d_in, stride, block_len, n_blocks, reinterpret_cast<double *>(d_audio), audio_stride,
                              audio_len, stream);
    if (s == FMR_OK) h->last_launches += extra;
    return s;
  }
  uint64_t out_total = 0;
  s = fmr_am_query_output(h, block_len, n_blocks, &out_total, nullptr);
  if (s != FMR_OK) return s;
  if (out_total > audio_stride) return fail(FMR_ERR_CAPACITY, "audio_stride too small for this call");
  const int64_t t0 = h->cum48;
  s = fmr_am_process_device(h, d_in, stride, block_len, n_blocks, h->d_audio, h->audio_cap, audio_len, stream);
  if (s != FMR_OK) return s;
  if (n_blocks > 0) {
    SinkParams P;
    P.out_fmt = out_cfg->out_format;
    P.w = 1; // AmDecoder / NbfmDecoder produce mono (AmDecode.h:53, NbfmDecode.h:48)
    P.gain = out_cfg->gain;
    P.squelch = out_cfg->squelch_level;
    dim3 sg(n_blocks, h->C);
    // both 48 kHz decoders measure the IF level behind their channel filter / tuner chain (m_buf_filtered2,
    // AmDecode.cpp:153-154; m_buf_filtered, NbfmDecode.cpp:50-54) = r_flt; it shares the call table with the audio
    k_audio_sink<<<sg, kSinkThreads, 0, st>>>(h->r_flt, t0, h->d_e48, h->d_audio, h->audio_cap, h->d_e48, (int)n_blocks,
                                              d_audio, audio_stride, h->d_levels, P);
    FMR_CUDA(cudaGetLastError());
    extra++;
  }
  h->last_launches += extra;
  h->have_levels = true;
  return FMR_OK;
}

extern "C" fmr_status fmr_am_process_host_io(fmr_am *h, const void *iq, int iq_format, size_t iq_stride,
                                             const uint32_t *block_len, uint32_t n_blocks,
                                             const fmr_output_config *out_cfg, void *audio, size_t audio_stride,
                                             uint32_t *audio_len) {
  if (!h || !iq || !block_len || !audio) return fail(FMR_ERR_INVALID, "null argument");
  const size_t esz = (size_t)iq_format_bytes(iq_format);
  if (esz == 0) return fail(FMR_ERR_INVALID, "unknown iq_format");
  if (out_cfg && out_format_bytes(out_cfg->out_format) == 0) return fail(FMR_ERR_INVALID, "unknown out_format");
  const size_t osz = out_cfg ? (size_t)out_format_bytes(out_cfg->out_format) : 8;
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_blocks; b++) total += block_len[b];
  if (total > h->cfg.max_samples_per_call) return fail(FMR_ERR_CAPACITY, "sum(block_len) > max_samples_per_call");
  if (total > iq_stride) return fail(FMR_ERR_INVALID, "iq_stride < sum(block_len)");
  const int C = h->C;
  fmr_status s = am_ensure_staging(h, (size_t)C * h->cfg.max_samples_per_call * esz, true);
  if (s != FMR_OK) return s;
  cudaStream_t st = h->own_stream;
  if (total > 0) {
    FMR_CUDA(cudaMemcpy2DAsync(h->d_raw, (size_t)total * esz, iq, iq_stride * esz, (size_t)total * esz, C,
                               cudaMemcpyHostToDevice, st));
  }
  uint64_t out_total = 0;
  s = fmr_am_query_output(h, block_len, n_blocks, &out_total, nullptr);
  if (s != FMR_OK) return s;
  if (out_total > audio_stride) return fail(FMR_ERR_CAPACITY, "audio_stride too small for this call");
  s = fmr_am_process_device_io(h, h->d_raw, iq_format, (size_t)total, block_len, n_blocks, out_cfg, h->d_out, h->audio_cap,
                               audio_len, (void *)st);
  if (s != FMR_OK) return s;
  if (out_total > 0) {
    FMR_CUDA(cudaMemcpy2DAsync(audio, audio_stride * osz, h->d_out, h->audio_cap * osz, (size_t)out_total * osz, C,
                               cudaMemcpyDeviceToHost, st));
  }
  FMR_CUDA(cudaStreamSynchronize(st));
  return FMR_OK;
}

extern "C" fmr_status fmr_am_block_levels(fmr_am *h, uint32_t channel, fmr_block_level_t *out, uint32_t n_blocks) {
  if (!h || !out || channel >= (uint32_t)h->C || n_blocks != h->last_blocks) {
    return fail(FMR_ERR_INVALID, "bad argument (n_blocks must equal the last call's)");
  }
  if (!h->have_levels) return fail(FMR_ERR_INVALID, "the last call had no output stage");
  static_assert(sizeof(BlockLevelDev) == sizeof(fmr_block_level_t), "level record layout");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  FMR_CUDA(cudaMemcpy(out, h->d_levels + (size_t)channel * n_blocks, n_blocks * sizeof(BlockLevelDev),
                      cudaMemcpyDeviceToHost));
  return FMR_OK;
}

extern "C" fmr_status fmr_am_stats(fmr_am *h, uint32_t channel, fmr_am_stats_t *out) {
  if (!h || !out || channel >= (uint32_t)h->C) return fail(FMR_ERR_INVALID, "bad argument");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  AmChanState s;
  FMR_CUDA(cudaMemcpy(&s, h->d_state + channel, sizeof(s), cudaMemcpyDeviceToHost));
  out->baseband_level = s.baseband_level;
  out->af_agc_gain = (float)s.af_gain;
  out->if_agc_gain = s.if_gain;
  out->if_rms = s.if_rms;
  out->decoder_calls = s.decoder_calls;
  out->tuning_offset = h->nbfm ? (float)(s.baseband_mean * h->freq_dev) : 0.0f; // NbfmDecode.h:59
  return FMR_OK;
}

extern "C" uint32_t fmr_am_last_launches(fmr_am *h) { return h ? h->last_launches : 0; }

extern "C" fmr_status fmr_am_set_profiling(fmr_am *h, int enable) {
  if (!h) return fail(FMR_ERR_INVALID, "null handle");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  if (enable) {
    h->prof.enable();
  } else {
    h->prof.release();
  }
  return FMR_OK;
}

extern "C" fmr_status fmr_am_stage_times(fmr_am *h, float *ms, const char **names, uint32_t cap, uint32_t *n) {
  if (!h || !ms || !names || !n) return fail(FMR_ERR_INVALID, "null argument");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  return h->prof.read(ms, names, cap, n);
}
