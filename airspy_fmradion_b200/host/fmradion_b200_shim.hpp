// fmradion_b200_shim.hpp — C++ host-side mirror of the reference's decoder classes over the
// C ABI of include/fmradion_b200.h.
//
// The reference has no plugin/FFI layer; its boundary for this path is the C++ class API that
// main.cpp consumes (include/FmDecode.h:34-105, include/AmDecode.h:32-65). This header declares
// classes with THE SAME names, constructor arguments, process() signature and getters, so
// main.cpp can include it instead of FmDecode.h / AmDecode.h and link libfmradion_b200.so
// (see INTEGRATION.md for the exact diff). Each object is one channel; the multi-channel
// super-block interface is the C ABI itself.
//
// Two additive constructor arguments (defaulted) let the GPU absorb the front end that the
// reference runs just before the decoder: `input_rate` (IfResampler, main.cpp:775-778,921-926)
// and `fs4_shift` (FourthConverterIQ, main.cpp:773,912-919). With the defaults the object
// expects 384 kHz (FM) / 48 kHz (AM) input exactly like the reference's decoder.
#ifndef FMRADION_B200_SHIM_HPP
#define FMRADION_B200_SHIM_HPP

#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fmradion_b200.h"

// Same aliases as the reference's include/SoftFM.h:33-45.
using IQSample = std::complex<float>;
using IQSampleVector = std::vector<IQSample>;
using Sample = double;
using SampleVector = std::vector<Sample>;
using IQSampleCoeff = std::vector<IQSample::value_type>;
using SampleCoeff = std::vector<SampleVector::value_type>;
using MfCoeff = std::complex<float>;
using MfCoeffVector = std::vector<MfCoeff>;
#ifndef INCLUDE_SOFTFM_H
enum class ModType { FM, NBFM, AM, DSB, USB, LSB, CW, WSPR };
#endif

namespace fmr_b200_detail {
inline void check(fmr_status s) {
  // The reference's process() never throws; CUDA failures are not a condition it knows.
  // They are fatal here (no CPU fallback), reported with the library's message.
  if (s != FMR_OK) throw std::runtime_error(std::string("fmradion_b200: ") + fmr_last_error());
}
} // namespace fmr_b200_detail

class PilotPhaseLock { // only the event type the FmDecoder API exposes (PilotPhaseLock.h:39-44)
public:
  struct PpsEvent {
    std::uint64_t pps_index;
    std::uint64_t sample_index;
    double block_position;
  };
};

class FmDecoder {
public:
  static constexpr double sample_rate_if = 384000;
  static constexpr double sample_rate_pcm = 48000;
  static constexpr double freq_dev = 75000;
  static constexpr double bandwidth_pcm = 15000;
  static constexpr double pilot_freq = 19000;
  static constexpr double deemphasis_time_eu = 50;
  static constexpr double deemphasis_time_na = 75;

  // Arguments as include/FmDecode.h:49-64. fmfilter_coeff is used only when fmfilter_enable.
  FmDecoder(bool fmfilter_enable, IQSampleCoeff &fmfilter_coeff, bool stereo, double deemphasis, bool pilot_shift,
            unsigned int multipath_stages, double input_rate = sample_rate_if, bool fs4_shift = false,
            int device = 0)
      : m_stereo(stereo), m_stages(multipath_stages) {
    fmr_fm_config cfg{};
    cfg.input_rate = input_rate;
    cfg.fs4_shift = fs4_shift ? 1 : 0;
    cfg.fmfilter = fmfilter_enable ? 3 : 0; // 3 = caller-supplied coefficients
    cfg.fmfilter_coeff = fmfilter_enable ? fmfilter_coeff.data() : nullptr;
    cfg.fmfilter_ntaps = fmfilter_enable ? (uint32_t)fmfilter_coeff.size() : 0;
    cfg.stereo = stereo ? 1 : 0;
    cfg.deemphasis_us = deemphasis;
    cfg.pilot_shift = pilot_shift ? 1 : 0;
    cfg.multipath_stages = multipath_stages;
    cfg.n_channels = 1;
    cfg.max_samples_per_call = 65536; // IfResampler::max_input_length (IfResampler.h:31)
    cfg.max_blocks_per_call = 1;
    cfg.device = device;
    fmr_b200_detail::check(fmr_fm_create(&cfg, &m_h));
  }
  ~FmDecoder() { fmr_fm_destroy(m_h); }
  FmDecoder(const FmDecoder &) = delete;
  FmDecoder &operator=(const FmDecoder &) = delete;

  // include/FmDecode.h:74. An empty block leaves the decoder untouched (FmDecode.cpp:88-92).
  void process(IQSampleVector samples_in, SampleVector &audio) {
    const uint32_t n = (uint32_t)samples_in.size();
    if (n == 0) {
      audio.resize(0);
      return;
    }
    uint64_t total = 0;
    fmr_b200_detail::check(fmr_fm_query_output(m_h, &n, 1, &total, nullptr));
    audio.resize((size_t)total);
    m_scratch.resize(total ? (size_t)total : 1);
    uint32_t len = 0;
    fmr_b200_detail::check(fmr_fm_process_host(m_h, reinterpret_cast<const float *>(samples_in.data()), n, &n, 1,
                                               m_scratch.data(), m_scratch.size(), &len));
    for (size_t i = 0; i < (size_t)total; i++) audio[i] = m_scratch[i];
    fmr_b200_detail::check(fmr_fm_stats(m_h, 0, &m_stats));
    m_pps.clear();
    if (m_stats.n_pps) {
      fmr_pps_event_t ev[16];
      uint32_t k = 0;
      fmr_b200_detail::check(fmr_fm_pps_events(m_h, 0, ev, 16, &k));
      for (uint32_t i = 0; i < k; i++) m_pps.push_back({ev[i].pps_index, ev[i].sample_index, ev[i].block_position});
    }
  }

  bool stereo_detected() const { return m_stats.stereo_detected != 0; }
  float get_tuning_offset() const { return m_stats.tuning_offset; }
  float get_baseband_level() const { return m_stats.baseband_level; }
  double get_pilot_level() const { return m_stats.pilot_level; }
  float get_if_rms() const { return m_stats.if_rms; }
  std::vector<PilotPhaseLock::PpsEvent> get_pps_events() const { return m_pps; }
  void erase_first_pps_event() {
    if (!m_pps.empty()) m_pps.erase(m_pps.begin());
  }
  double get_multipath_error() { return m_stats.multipath_error; }
  const MfCoeffVector &get_multipath_coefficients() {
    const size_t n = 4 * (size_t)(m_stages ? m_stages : 1) + 1;
    m_coeff.assign(n, MfCoeff(0, 0));
    if (m_stages) {
      fmr_b200_detail::check(fmr_fm_coeffs(m_h, 0, reinterpret_cast<float *>(m_coeff.data()), n));
    } else {
      m_coeff[4] = MfCoeff(1, 0); // MultipathFilter(1) untouched: reference tap only (FmDecode.cpp:79)
    }
    return m_coeff;
  }

private:
  fmr_fm *m_h = nullptr;
  bool m_stereo;
  unsigned int m_stages;
  fmr_fm_stats_t m_stats{};
  std::vector<PilotPhaseLock::PpsEvent> m_pps;
  std::vector<double> m_scratch;
  MfCoeffVector m_coeff;
};

class AmDecoder {
public:
  static constexpr double sample_rate_pcm = 48000;
  static constexpr double internal_rate_pcm = 48000;
  static constexpr double bandwidth_pcm = 4500;
  static constexpr double deemphasis_time = 100;

  // Arguments as include/AmDecode.h:42-48; ModType AM, DSB, USB, LSB, CW and WSPR (NBFM: see NbfmDecoder).
  AmDecoder(IQSampleCoeff &amfilter_coeff, const ModType mode, double input_rate = internal_rate_pcm,
            bool fs4_shift = false, int device = 0) {
    fmr_am_config cfg{};
    cfg.input_rate = input_rate;
    cfg.fs4_shift = fs4_shift ? 1 : 0;
    cfg.amfilter = 4; // caller-supplied coefficients
    cfg.amfilter_coeff = amfilter_coeff.data();
    cfg.amfilter_ntaps = (uint32_t)amfilter_coeff.size();
    cfg.mode = static_cast<int>(mode);
    cfg.n_channels = 1;
    cfg.max_samples_per_call = 65536;
    cfg.max_blocks_per_call = 1;
    cfg.device = device;
    fmr_b200_detail::check(fmr_am_create(&cfg, &m_h));
  }
  ~AmDecoder() { fmr_am_destroy(m_h); }
  AmDecoder(const AmDecoder &) = delete;
  AmDecoder &operator=(const AmDecoder &) = delete;

  void process(IQSampleVector samples_in, SampleVector &audio) {
    const uint32_t n = (uint32_t)samples_in.size();
    if (n == 0) {
      audio.resize(0);
      return;
    }
    uint64_t total = 0;
    fmr_b200_detail::check(fmr_am_query_output(m_h, &n, 1, &total, nullptr));
    audio.resize(total ? (size_t)total : 1);
    uint32_t len = 0;
    fmr_b200_detail::check(fmr_am_process_host(m_h, reinterpret_cast<const float *>(samples_in.data()), n, &n, 1,
                                               audio.data(), audio.size(), &len));
    audio.resize((size_t)total);
    fmr_b200_detail::check(fmr_am_stats(m_h, 0, &m_stats));
  }
  double get_baseband_level() const { return m_stats.baseband_level; }
  float get_af_agc_current_gain() const { return m_stats.af_agc_gain; }
  float get_if_agc_current_gain() const { return m_stats.if_agc_gain; }
  float get_if_rms() const { return m_stats.if_rms; }

private:
  fmr_am *m_h = nullptr;
  fmr_am_stats_t m_stats{};
};

// Narrow-band FM decoder with the reference's interface (include/NbfmDecode.h:30-63).
class NbfmDecoder {
public:
  static constexpr double sample_rate_pcm = 48000;
  static constexpr double internal_rate_pcm = 48000;
  static constexpr double freq_dev_normal = 8000;
  static constexpr double freq_dev_wide = 17000;

  NbfmDecoder(IQSampleCoeff &nbfmfilter_coeff, const double freq_dev, double input_rate = internal_rate_pcm,
              bool fs4_shift = false, int device = 0) {
    fmr_am_config cfg{};
    cfg.input_rate = input_rate;
    cfg.fs4_shift = fs4_shift ? 1 : 0;
    cfg.amfilter = 4; // caller-supplied coefficients
    cfg.amfilter_coeff = nbfmfilter_coeff.data();
    cfg.amfilter_ntaps = (uint32_t)nbfmfilter_coeff.size();
    cfg.mode = static_cast<int>(ModType::NBFM);
    cfg.nbfm_freq_dev = freq_dev;
    cfg.n_channels = 1;
    cfg.max_samples_per_call = 65536;
    cfg.max_blocks_per_call = 1;
    cfg.device = device;
    fmr_b200_detail::check(fmr_am_create(&cfg, &m_h));
  }
  ~NbfmDecoder() { fmr_am_destroy(m_h); }
  NbfmDecoder(const NbfmDecoder &) = delete;
  NbfmDecoder &operator=(const NbfmDecoder &) = delete;

  void process(const IQSampleVector &samples_in, SampleVector &audio) {
    const uint32_t n = (uint32_t)samples_in.size();
    if (n == 0) {
      audio.resize(0);
      return;
    }
    uint64_t total = 0;
    fmr_b200_detail::check(fmr_am_query_output(m_h, &n, 1, &total, nullptr));
    audio.resize(total ? (size_t)total : 1);
    uint32_t len = 0;
    fmr_b200_detail::check(fmr_am_process_host(m_h, reinterpret_cast<const float *>(samples_in.data()), n, &n, 1,
                                               audio.data(), audio.size(), &len));
    audio.resize((size_t)total);
    fmr_b200_detail::check(fmr_am_stats(m_h, 0, &m_stats));
  }
  float get_tuning_offset() const { return m_stats.tuning_offset; }
  float get_baseband_level() const { return (float)m_stats.baseband_level; }
  float get_if_rms() const { return m_stats.if_rms; }

private:
  fmr_am *m_h = nullptr;
  fmr_am_stats_t m_stats{};
};

#endif
