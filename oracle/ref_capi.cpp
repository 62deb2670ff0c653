// TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product path.
//
// Thin C wrapper around the UNMODIFIED reference classes, compiled from the sources where
// they lie under /root/reference (see oracle/Makefile). It drives them exactly like the
// reference's block loop does (main.cpp:912-974: FourthConverterIQ when the source is
// zero-IF, IfResampler when ifrate != demodulator rate, `continue` on empty output, then
// FmDecoder::process / AmDecoder::process) but without FileSource's real-time pacing
// (FileSource.cpp:430-454) and without the final adjust_gain(0.5) of main.cpp:1000 (that
// scale belongs to the caller, not to the decoder block API).
//
// It also exposes: per-call intermediate taps (private members are read through the
// `#define private public` trick below, which does not change class layout), the r8brain
// stage tables of a resampler chain (used by tools/gen_r8b_tables.py to generate the
// coefficient tables the product ships), and a multi-threaded throughput timer used as
// the CPU baseline (bench.py --impl reference).

#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "CDSPResampler.h"

#include "AmDecode.h"
#include "AudioResampler.h"
#include "FilterParameters.h"
#include "FmDecode.h"
#include "FourthConverterIQ.h"
#include "IfResampler.h"
#include "NbfmDecode.h"
#include "Utility.h"
#undef private
#undef protected

namespace {

struct RefChain {
  int mode; // 0 = FM, 1 = AM, 2 = NBFM
  double ifrate;
  double demod_rate;
  bool fs4;
  bool downsample;
  IQSampleCoeff fmfilter_coeff;
  IQSampleCoeff amfilter_coeff;
  IQSampleCoeff nbfmfilter_coeff;
  std::unique_ptr<NbfmDecoder> nbfm;
  std::unique_ptr<FourthConverterIQ> fourth;
  std::unique_ptr<IfResampler> ifres;
  std::unique_ptr<FmDecoder> fm;
  std::unique_ptr<AmDecoder> am;
  IQSampleVector last_if; // decoder input of the last call
  uint64_t calls = 0;
  uint64_t decoder_calls = 0;
};

void front_end(RefChain *c, const float *iq, int n, IQSampleVector &if_samples) {
  IQSampleVector iqsamples(n);
  for (int i = 0; i < n; i++) {
    iqsamples[i] = IQSample(iq[2 * i], iq[2 * i + 1]);
  }
  IQSampleVector shifted;
  if (c->fs4) {
    c->fourth->process(iqsamples, shifted);
  } else {
    shifted = std::move(iqsamples);
  }
  if (c->downsample) {
    c->ifres->process(shifted, if_samples);
  } else {
    if_samples = std::move(shifted);
  }
}

} // namespace

extern "C" {

// filter: 0 default (no IF filter), 1 medium, 2 narrow  (main.cpp:785-810)
void *ref_fm_create(double ifrate, int fs4_shift, int filter, int stereo,
                    double deemphasis_us, int pilot_shift,
                    unsigned int mpf_stages) {
  RefChain *c = new RefChain();
  c->mode = 0;
  c->ifrate = ifrate;
  c->demod_rate = FmDecoder::sample_rate_if;
  c->fs4 = fs4_shift != 0;
  c->downsample = (ifrate != c->demod_rate);
  bool enable = false;
  switch (filter) {
  case 1:
    c->fmfilter_coeff = FilterParameters::jj1bdx_fm_384kHz_medium;
    enable = true;
    break;
  case 2:
    c->fmfilter_coeff = FilterParameters::jj1bdx_fm_384kHz_narrow;
    enable = true;
    break;
  default:
    c->fmfilter_coeff = FilterParameters::delay_3taps_only_iq;
    break;
  }
  c->fourth = std::make_unique<FourthConverterIQ>(false);
  c->ifres = std::make_unique<IfResampler>(ifrate, c->demod_rate);
  c->fm = std::make_unique<FmDecoder>(enable, c->fmfilter_coeff, stereo != 0,
                                      deemphasis_us, pilot_shift != 0,
                                      mpf_stages);
  return c;
}

// filter: 0 default, 1 medium, 2 narrow, 3 wide. mode: ModType enum value (2 = AM).
void *ref_am_create(double ifrate, int fs4_shift, int filter, int modtype) {
  RefChain *c = new RefChain();
  c->mode = 1;
  c->ifrate = ifrate;
  c->demod_rate = AmDecoder::internal_rate_pcm;
  c->fs4 = fs4_shift != 0;
  c->downsample = (ifrate != c->demod_rate);
  switch (filter) {
  case 1:
    c->amfilter_coeff = FilterParameters::jj1bdx_am_48khz_medium;
    break;
  case 2:
    c->amfilter_coeff = FilterParameters::jj1bdx_am_48khz_narrow;
    break;
  case 3:
    c->amfilter_coeff = FilterParameters::jj1bdx_am_48khz_wide;
    break;
  default:
    c->amfilter_coeff = FilterParameters::jj1bdx_am_48khz_default;
    break;
  }
  c->fourth = std::make_unique<FourthConverterIQ>(false);
  c->ifres = std::make_unique<IfResampler>(ifrate, c->demod_rate);
  c->am = std::make_unique<AmDecoder>(c->amfilter_coeff,
                                      static_cast<ModType>(modtype));
  return c;
}

// filter: 0 default, 1 medium, 2 narrow, 3 wide (main.cpp:785-810); freq_dev in Hz
// (NbfmDecoder::freq_dev_normal = 8000, main.cpp:828-830).
void *ref_nbfm_create(double ifrate, int fs4_shift, int filter, double freq_dev) {
  RefChain *c = new RefChain();
  c->mode = 2;
  c->ifrate = ifrate;
  c->demod_rate = NbfmDecoder::internal_rate_pcm;
  c->fs4 = fs4_shift != 0;
  c->downsample = (ifrate != c->demod_rate);
  switch (filter) {
  case 1:
    c->nbfmfilter_coeff = FilterParameters::jj1bdx_nbfm_48khz_medium;
    break;
  case 2:
    c->nbfmfilter_coeff = FilterParameters::jj1bdx_nbfm_48khz_narrow;
    break;
  case 3:
    c->nbfmfilter_coeff = FilterParameters::jj1bdx_nbfm_48khz_wide;
    break;
  default:
    c->nbfmfilter_coeff = FilterParameters::jj1bdx_nbfm_48khz_default;
    break;
  }
  c->fourth = std::make_unique<FourthConverterIQ>(false);
  c->ifres = std::make_unique<IfResampler>(ifrate, c->demod_rate);
  c->nbfm = std::make_unique<NbfmDecoder>(c->nbfmfilter_coeff, freq_dev);
  return c;
}

void ref_destroy(void *h) { delete static_cast<RefChain *>(h); }

// One source block. Returns the number of doubles written to `audio`
// (0 when the resampler or decoder produced nothing), or -1 if cap is too small.
int ref_process_block(void *h, const float *iq, int n, double *audio, int cap) {
  RefChain *c = static_cast<RefChain *>(h);
  c->calls++;
  IQSampleVector if_samples;
  front_end(c, iq, n, if_samples);
  c->last_if = if_samples;
  if (if_samples.empty()) {
    return 0; // main.cpp:933-936
  }
  c->decoder_calls++;
  SampleVector out;
  if (c->mode == 0) {
    c->fm->process(std::move(if_samples), out);
  } else if (c->mode == 2) {
    c->nbfm->process(if_samples, out);
  } else {
    c->am->process(std::move(if_samples), out);
  }
  if ((int)out.size() > cap) {
    return -1;
  }
  std::copy(out.begin(), out.end(), audio);
  return (int)out.size();
}

// ---- taps of the last call (complex taps are interleaved re,im) ----
static int copy_c(const IQSampleVector &v, float *out, int cap) {
  int n = std::min<int>(v.size(), cap);
  for (int i = 0; i < n; i++) {
    out[2 * i] = v[i].real();
    out[2 * i + 1] = v[i].imag();
  }
  return (int)v.size();
}
static int copy_d(const SampleVector &v, double *out, int cap) {
  int n = std::min<int>(v.size(), cap);
  std::copy(v.begin(), v.begin() + n, out);
  return (int)v.size();
}

int ref_tap_if(void *h, float *out, int cap) {
  return copy_c(static_cast<RefChain *>(h)->last_if, out, cap);
}
// IQ that went into the discriminator (after AGC and multipath filter).
int ref_tap_fm_preDisc(void *h, float *out, int cap) {
  return copy_c(
      static_cast<RefChain *>(h)->fm->m_samples_in_multipathfiltered, out, cap);
}
int ref_tap_fm_mpx(void *h, float *out, int cap) {
  const auto &v = static_cast<RefChain *>(h)->fm->m_buf_decoded;
  int n = std::min<int>(v.size(), cap);
  std::copy(v.begin(), v.begin() + n, out);
  return (int)v.size();
}
// after deemphasis, 384 kHz
int ref_tap_fm_mono384(void *h, double *out, int cap) {
  return copy_d(static_cast<RefChain *>(h)->fm->m_buf_baseband, out, cap);
}
int ref_tap_fm_stereo384(void *h, double *out, int cap) {
  return copy_d(static_cast<RefChain *>(h)->fm->m_buf_rawstereo, out, cap);
}
int ref_tap_fm_mono48_first(void *h, double *out, int cap) {
  return copy_d(static_cast<RefChain *>(h)->fm->m_buf_mono_firstout, out, cap);
}
int ref_tap_fm_stereo48_first(void *h, double *out, int cap) {
  return copy_d(static_cast<RefChain *>(h)->fm->m_buf_stereo_firstout, out,
                cap);
}

struct RefFmStats {
  int stereo_detected;
  float tuning_offset;
  float baseband_level;
  double pilot_level;
  float if_rms;
  double mpf_error;
  float agc_gain;
  double pll_freq;
  double pll_phase;
  int pll_lock_cnt;
  uint64_t decoder_calls;
  int n_pps;
};

void ref_fm_stats(void *h, RefFmStats *s) {
  RefChain *c = static_cast<RefChain *>(h);
  FmDecoder &fm = *c->fm;
  s->stereo_detected = fm.stereo_detected();
  s->tuning_offset = fm.get_tuning_offset();
  s->baseband_level = fm.get_baseband_level();
  s->pilot_level = fm.get_pilot_level();
  s->if_rms = fm.get_if_rms();
  s->mpf_error = fm.get_multipath_error();
  s->agc_gain = fm.m_ifagc.get_current_gain();
  s->pll_freq = fm.m_pilotpll.m_freq;
  s->pll_phase = fm.m_pilotpll.m_phase;
  s->pll_lock_cnt = fm.m_pilotpll.m_lock_cnt;
  s->decoder_calls = c->decoder_calls;
  s->n_pps = (int)fm.m_pilotpll.m_pps_events.size();
}

// PPS events of the last call: triples (pps_index, sample_index, block_position).
int ref_fm_pps(void *h, double *out, int cap_events) {
  RefChain *c = static_cast<RefChain *>(h);
  auto ev = c->fm->get_pps_events();
  int n = std::min<int>(ev.size(), cap_events);
  for (int i = 0; i < n; i++) {
    out[3 * i] = (double)ev[i].pps_index;
    out[3 * i + 1] = (double)ev[i].sample_index;
    out[3 * i + 2] = ev[i].block_position;
  }
  return (int)ev.size();
}

int ref_fm_mpf_coeffs(void *h, float *out, int cap) {
  RefChain *c = static_cast<RefChain *>(h);
  const MfCoeffVector &v = c->fm->get_multipath_coefficients();
  int n = std::min<int>(v.size(), cap);
  for (int i = 0; i < n; i++) {
    out[2 * i] = v[i].real();
    out[2 * i + 1] = v[i].imag();
  }
  return (int)v.size();
}

struct RefAmStats {
  double baseband_level;
  float af_agc_gain;
  float if_agc_gain;
  float if_rms;
  uint64_t decoder_calls;
};

void ref_am_stats(void *h, RefAmStats *s) {
  RefChain *c = static_cast<RefChain *>(h);
  s->baseband_level = c->am->get_baseband_level();
  s->af_agc_gain = c->am->get_af_agc_current_gain();
  s->if_agc_gain = c->am->get_if_agc_current_gain();
  s->if_rms = c->am->get_if_rms();
  s->decoder_calls = c->decoder_calls;
}

struct RefNbfmStats {
  float tuning_offset;
  float baseband_level;
  float if_rms;
  float if_agc_gain;
  uint64_t decoder_calls;
};

void ref_nbfm_stats(void *h, RefNbfmStats *s) {
  RefChain *c = static_cast<RefChain *>(h);
  s->tuning_offset = c->nbfm->get_tuning_offset();
  s->baseband_level = c->nbfm->get_baseband_level();
  s->if_rms = c->nbfm->get_if_rms();
  s->if_agc_gain = c->nbfm->m_ifagc.get_current_gain();
  s->decoder_calls = c->decoder_calls;
}

// ---- stand-alone resampler access (for the schedule / table checks) ----
// kind 0: IfResampler-style CDSPResampler24 (one real lane); kind 1: AudioResampler-style.
void *ref_r8b_create(double src, double dst, int kind) {
  if (kind == 0) {
    return new r8b::CDSPResampler24(src, dst, 65536);
  }
  return new r8b::CDSPResampler(src, dst, 32768);
}
void ref_r8b_destroy(void *h) { delete static_cast<r8b::CDSPResampler *>(h); }
int ref_r8b_process(void *h, const double *in, int n, double *out, int cap) {
  std::vector<double> tmp(in, in + n);
  double *op;
  int m = static_cast<r8b::CDSPResampler *>(h)->process(tmp.data(), n, op);
  if (m > cap) {
    return -1;
  }
  std::copy(op, op + m, out);
  return m;
}

// Dump the stages of the chain r8brain builds for (src, dst, kind) as text:
//   chain <nsteps>
//   hb <taps> <latency>            followed by <taps> coefficient lines
//   bc <klen> <inputlen> <latency> <up> <down> <outoffset> <inputdelay> <downskipinit>
//                                  followed by <klen> effective time-domain taps
//   fi <instep> <outstep> <flen> <initfracposw> <latency>
//                                  followed by outstep*flen taps, phase-major
// Time-domain taps of the block convolver are measured by pushing an impulse through a
// fresh 1:1 convolver built on the same cached filter object (they include FinGain).
// Returns 0 on success.
int ref_r8b_dump(double src, double dst, int kind, const char *path) {
  std::unique_ptr<r8b::CDSPResampler> rs;
  if (kind == 0) {
    rs.reset(new r8b::CDSPResampler24(src, dst, 65536));
  } else {
    rs.reset(new r8b::CDSPResampler(src, dst, 32768));
  }
  FILE *f = fopen(path, "w");
  if (!f) {
    return 1;
  }
  fprintf(f, "chain %d\n", rs->StepCount);
  for (int i = 0; i < rs->StepCount; i++) {
    r8b::CDSPProcessor *p = rs->Steps[i];
    if (auto *hb = dynamic_cast<r8b::CDSPHBDownsampler *>(p)) {
      fprintf(f, "hb %d %d\n", hb->fll, hb->Latency);
      for (int k = 0; k < hb->fll; k++) {
        fprintf(f, "%.17g\n", hb->fltp[k]);
      }
    } else if (auto *bc = dynamic_cast<r8b::CDSPBlockConvolver *>(p)) {
      r8b::CDSPFIRFilter *flt = bc->Filter;
      const int klen = flt->getKernelLen();
      fprintf(f, "bc %d %d %d %d %d %d %d %d\n", klen, bc->InputLen,
              bc->Latency, bc->UpFactor, bc->DownFactor, bc->OutOffset,
              bc->InputDelay, bc->DownSkipInit);
      r8b::CDSPFIRFilter &f2 = r8b::CDSPFIRFilterCache::getLPFilter(
          flt->ReqNormFreq, flt->ReqTransBand, flt->ReqAtten, flt->ReqPhase,
          flt->ReqGain);
      r8b::CDSPBlockConvolver probe(f2, 1, 1, 0.0);
      const int P = klen;
      const int total = P + klen + probe.InputLen + klen + 16;
      std::vector<double> in(total, 0.0), out(total + 16, 0.0);
      in[P] = 1.0;
      std::vector<double> resp;
      int pos = 0;
      while (pos < total) {
        int chunk = std::min(4096, total - pos);
        double *op = out.data();
        int m = probe.process(in.data() + pos, chunk, op);
        resp.insert(resp.end(), op, op + m);
        pos += chunk;
      }
      const int fl2 = (klen - 1) / 2;
      for (int k = -fl2; k <= fl2; k++) {
        // zero-phase: y[t] = h[t - P]; average the two mirror taps to shed FFT noise
        double a = resp[P + k], b = resp[P - k];
        fprintf(f, "%.17g\n", 0.5 * (a + b));
      }
    } else if (auto *fi = dynamic_cast<r8b::CDSPFracInterpolator *>(p)) {
      if (!fi->IsWhole) {
        fclose(f);
        return 2;
      }
      fprintf(f, "fi %d %d %d %d %d\n", fi->InStep, fi->OutStep, fi->FilterLen,
              fi->InitFracPosW, fi->Latency);
      for (int ph = 0; ph < fi->OutStep; ph++) {
        const double *t = &(*fi->FilterBank)[ph];
        for (int k = 0; k < fi->FilterLen; k++) {
          fprintf(f, "%.17g\n", t[k]);
        }
      }
    } else {
      fclose(f);
      return 3;
    }
  }
  fclose(f);
  return 0;
}

// Dump a FilterParameters table by name: returns length, writes up to cap doubles.
int ref_filter_table(const char *name, double *out, int cap) {
  std::string s(name);
  std::vector<double> v;
  auto cf = [&](const IQSampleCoeff &c) { v.assign(c.begin(), c.end()); };
  auto cd = [&](const SampleCoeff &c) { v.assign(c.begin(), c.end()); };
  if (s == "delay_3taps_only_iq")
    cf(FilterParameters::delay_3taps_only_iq);
  else if (s == "jj1bdx_48khz_fmaudio")
    cd(FilterParameters::jj1bdx_48khz_fmaudio);
  else if (s == "jj1bdx_48khz_nbfmaudio")
    cd(FilterParameters::jj1bdx_48khz_nbfmaudio);
  else if (s == "jj1bdx_nbfm_48khz_default")
    cf(FilterParameters::jj1bdx_nbfm_48khz_default);
  else if (s == "jj1bdx_nbfm_48khz_medium")
    cf(FilterParameters::jj1bdx_nbfm_48khz_medium);
  else if (s == "jj1bdx_nbfm_48khz_narrow")
    cf(FilterParameters::jj1bdx_nbfm_48khz_narrow);
  else if (s == "jj1bdx_nbfm_48khz_wide")
    cf(FilterParameters::jj1bdx_nbfm_48khz_wide);
  else if (s == "jj1bdx_cw_48khz_500hz")
    cf(FilterParameters::jj1bdx_cw_48khz_500hz);
  else if (s == "jj1bdx_ssb_48khz_1500hz")
    cf(FilterParameters::jj1bdx_ssb_48khz_1500hz);
  else if (s == "jj1bdx_am_48khz_narrow")
    cf(FilterParameters::jj1bdx_am_48khz_narrow);
  else if (s == "jj1bdx_am_48khz_medium")
    cf(FilterParameters::jj1bdx_am_48khz_medium);
  else if (s == "jj1bdx_am_48khz_default")
    cf(FilterParameters::jj1bdx_am_48khz_default);
  else if (s == "jj1bdx_am_48khz_wide")
    cf(FilterParameters::jj1bdx_am_48khz_wide);
  else if (s == "jj1bdx_fm_384kHz_narrow")
    cf(FilterParameters::jj1bdx_fm_384kHz_narrow);
  else if (s == "jj1bdx_fm_384kHz_medium")
    cf(FilterParameters::jj1bdx_fm_384kHz_medium);
  else
    return -1;
  int n = std::min<int>(v.size(), cap);
  std::copy(v.begin(), v.begin() + n, out);
  return (int)v.size();
}

float ref_fast_atan2f(float y, float x) { return Utility::fast_atan2f(y, x); }

// The 257-entry table behind fast_atan2f (include/Utility.h:165-217).
int ref_fast_atan_table(float *out, int cap) {
  const int n = (int)(sizeof(Utility::fast_atan_table) / sizeof(float));
  for (int i = 0; i < n && i < cap; i++) {
    out[i] = Utility::fast_atan_table[i];
  }
  return n;
}

// ---- CPU baseline: nthreads independent chains, each consuming `blocks` blocks of
// `blklen` samples read cyclically from iq[0..n_iq). Returns wall seconds (all threads).
// mode 0 = FM (stereo, deemph 50, given mpf stages), 1 = AM default filter.
double ref_bench(int mode, double ifrate, int stereo, unsigned int mpf_stages,
                 int nthreads, long blocks, int blklen, const float *iq,
                 long n_iq) {
  std::vector<void *> chains(nthreads);
  for (int t = 0; t < nthreads; t++) {
    chains[t] = (mode == 0)
                    ? ref_fm_create(ifrate, 0, 0, stereo, 50.0, 0, mpf_stages)
                    : ref_am_create(ifrate, 0, 0, (int)ModType::AM);
  }
  std::atomic<int> ready(0);
  std::atomic<bool> go(false);
  std::vector<std::thread> th;
  std::vector<double> sink(nthreads, 0.0);
  for (int t = 0; t < nthreads; t++) {
    th.emplace_back([&, t]() {
      std::vector<double> audio(1 << 16);
      long nblk_avail = n_iq / blklen;
      ready++;
      while (!go.load()) {
        std::this_thread::yield();
      }
      double acc = 0;
      for (long b = 0; b < blocks; b++) {
        const float *p = iq + 2 * (size_t)((b + 7 * t) % nblk_avail) * blklen;
        int m = ref_process_block(chains[t], p, blklen, audio.data(),
                                  (int)audio.size());
        if (m > 0) {
          acc += audio[0];
        }
      }
      sink[t] = acc;
    });
  }
  while (ready.load() < nthreads) {
    std::this_thread::yield();
  }
  auto t0 = std::chrono::steady_clock::now();
  go.store(true);
  for (auto &x : th) {
    x.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  for (int t = 0; t < nthreads; t++) {
    ref_destroy(chains[t]);
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
