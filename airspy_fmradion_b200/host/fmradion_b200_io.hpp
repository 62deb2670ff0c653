// fmradion_b200_io.hpp — host-side C++ for the two steps either side of the decoder path
// (SURVEY.md §8 f1, f4), in the reference's language and with the reference's class names:
//
//   FileSource     sfmbase/FileSource.cpp:55-246,472-531, include/FileSource.h:32-117
//                  "filename=...,srate=...,blklen=...,zero_offset,format=...,raw" -> blocks of IQ samples.
//                  The reference reads through libsndfile (absent here, un-vendored); this class parses the
//                  containers FileSource accepts itself (WAV, WAVEX, W64, RAW; FileSource.cpp:303-307) for the
//                  sub-types it accepts (PCM_S8, PCM_16, PCM_24, PCM_U8, FLOAT; FileSource.cpp:206-216,342-346)
//                  and converts like sf_read_float does with its default normalisation.
//                  Additive: get_raw_block() hands out the file's own bytes + sample_format(), so that the
//                  conversion runs on the GPU (fmr_*_process_host_io) and only those bytes cross PCIe.
//                  Not rebuilt (control plane): the pacing thread (FileSource.cpp:379-470) and DataBuffer.
//   SndfileOutput  sfmbase/AudioOutput.cpp:33-167 — WAV / raw writer for the sink formats main.cpp:592-623 uses
//                  (PCM_16 and FLOAT, little endian; RF64 is written as plain WAV, which is what libsndfile's
//                  SFC_RF64_AUTO_DOWNGRADE leaves on disk below 4 GB). Additive: write_native() for audio that
//                  the GPU output stage already converted to the sink's sample format.
//   BlockLoopLevels  the scalars the block loop keeps per block (main.cpp:872-876,950,976,996,1028): IF level and
//                  audio level EMAs, fed from fmr_block_level_t; format_pps_line (main.cpp:1084-1096).
//
// Header-only, no CUDA types; needs only include/fmradion_b200.h for the format enums.
#ifndef FMRADION_B200_IO_HPP
#define FMRADION_B200_IO_HPP

#include <cerrno>
#include <climits>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/fmradion_b200.h"

#ifndef FMRADION_B200_SHIM_HPP
using IQSample = std::complex<float>;
using IQSampleVector = std::vector<IQSample>;
using Sample = double;
using SampleVector = std::vector<Sample>;
#endif

namespace fmr_b200 {

// ---- "foo=x,bar,baz=10" (include/ConfigParser.h:29-34): only the leftmost '=' splits, a bare key maps to "".
inline std::map<std::string, std::string> parse_config_string(const std::string &text) {
  std::map<std::string, std::string> out;
  size_t pos = 0;
  while (pos <= text.size()) {
    size_t comma = text.find(',', pos);
    if (comma == std::string::npos) comma = text.size();
    const std::string item = text.substr(pos, comma - pos);
    if (!item.empty()) {
      const size_t eq = item.find('=');
      if (eq == std::string::npos) {
        out[item] = "";
      } else {
        out[item.substr(0, eq)] = item.substr(eq + 1);
      }
    }
    pos = comma + 1;
  }
  return out;
}

// Utility::parse_int (include/Utility.h:96-116): decimal integer, optional "k" suffix when allow_unit.
inline bool parse_int(const char *s, int &v, bool allow_unit = false) {
  char *endp = nullptr;
  errno = 0;
  long t = std::strtol(s, &endp, 10);
  if (endp == s || errno == ERANGE) return false;
  if (allow_unit && *endp == 'k' && t > INT_MIN / 1000 && t < INT_MAX / 1000) {
    t *= 1000;
    endp++;
  }
  if (*endp != '\0' || t < INT_MIN || t > INT_MAX) return false;
  v = (int)t;
  return true;
}

inline int iq_format_bytes(int fmt) { // bytes per complex sample
  switch (fmt) {
  case FMR_IQ_CF32: return 8;
  case FMR_IQ_S16: return 4;
  case FMR_IQ_S8: return 2;
  case FMR_IQ_U8: return 2;
  case FMR_IQ_S24: return 6;
  default: return 0;
  }
}

// sf_read_float for the accepted sub-types: value * 2^-(bits-1), U8 re-centred by 128; FLOAT unchanged
// (libsndfile pcm.c sc2f_array / uc2f_array / les2f_array / let2f_array with normalisation on).
inline void convert_iq(const uint8_t *raw, int fmt, size_t n, IQSample *out) {
  for (size_t i = 0; i < n; i++) {
    float re = 0.f, im = 0.f;
    switch (fmt) {
    case FMR_IQ_CF32: {
      float v[2];
      std::memcpy(v, raw + 8 * i, 8);
      re = v[0];
      im = v[1];
      break;
    }
    case FMR_IQ_S16: {
      int16_t v[2];
      std::memcpy(v, raw + 4 * i, 4);
      re = (float)v[0] * (1.0f / 32768.0f);
      im = (float)v[1] * (1.0f / 32768.0f);
      break;
    }
    case FMR_IQ_S8:
      re = (float)(int8_t)raw[2 * i] * (1.0f / 128.0f);
      im = (float)(int8_t)raw[2 * i + 1] * (1.0f / 128.0f);
      break;
    case FMR_IQ_U8:
      re = (float)((int)raw[2 * i] - 128) * (1.0f / 128.0f);
      im = (float)((int)raw[2 * i + 1] - 128) * (1.0f / 128.0f);
      break;
    case FMR_IQ_S24: {
      const uint8_t *p = raw + 6 * i;
      const int32_t a = (int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24));
      const int32_t b = (int32_t)(((uint32_t)p[3] << 8) | ((uint32_t)p[4] << 16) | ((uint32_t)p[5] << 24));
      re = (float)a * (1.0f / 2147483648.0f);
      im = (float)b * (1.0f / 2147483648.0f);
      break;
    }
    default: break;
    }
    out[i] = IQSample(re, im);
  }
}

} // namespace fmr_b200

class FileSource {
public:
  static constexpr int default_block_length = 2048;              // FileSource.h:34
  static constexpr std::uint32_t default_sample_rate = 384000;   // :35
  static constexpr std::int32_t default_frequency = 82500000;    // :36
  static constexpr int max_expected_us = 10000;                  // :39
  enum class FormatType { Unknown = 0, S8_LE = 1, S16_LE = 2, S24_LE = 3, U8_LE = 5, Float = 6 }; // :73-80
  enum class Container { None, Wav, WavEx, W64, Raw };

  explicit FileSource(int dev_index = 0) { (void)dev_index; }
  ~FileSource() { close(); }
  FileSource(const FileSource &) = delete;
  FileSource &operator=(const FileSource &) = delete;

  // FileSource::configure(std::string) (FileSource.cpp:55-162): same keys, same defaults, same failures.
  bool configure(const std::string &configuration) {
    auto m = fmr_b200::parse_config_string(configuration);
    std::string filename;
    bool raw = false, zero_offset = false;
    FormatType format_type = FormatType::S16_LE;
    std::uint32_t sample_rate = default_sample_rate, frequency = (std::uint32_t)default_frequency;
    int block_length = default_block_length;
    if (m.count("filename")) filename = m["filename"];
    if (m.count("srate")) {
      int v = 0;
      if (!fmr_b200::parse_int(m["srate"].c_str(), v, true)) return set_error("FileSource::configure: invalid samplerate");
      sample_rate = (std::uint32_t)v;
    }
    if (m.count("freq")) {
      int v = 0;
      if (!fmr_b200::parse_int(m["freq"].c_str(), v, true)) return set_error("FileSource::configure: invalid frequency");
      frequency = (std::uint32_t)v;
    }
    if (m.count("blklen")) {
      if (!fmr_b200::parse_int(m["blklen"].c_str(), block_length) || block_length <= 0) {
        return set_error("FileSource::configure: invalid blklen");
      }
    }
    if (m.count("zero_offset")) zero_offset = true;
    if (m.count("format")) {
      const std::string &f = m["format"];
      if (f == "S8_LE") format_type = FormatType::S8_LE;
      else if (f == "S16_LE") format_type = FormatType::S16_LE;
      else if (f == "S24_LE") format_type = FormatType::S24_LE;
      else if (f == "U8_LE") format_type = FormatType::U8_LE;
      else if (f == "FLOAT") format_type = FormatType::Float;
      else return set_error("FileSource::configure: format: " + f + " is not supported.");
    }
    if (m.count("raw")) raw = true;
    return configure(filename, raw, format_type, sample_rate, frequency, zero_offset, block_length);
  }

  // FileSource::configure(fname, raw, ...) (FileSource.cpp:164-246).
  bool configure(const std::string &fname, bool raw, FormatType format_type = FormatType::S16_LE,
                 std::uint32_t sample_rate = default_sample_rate, std::uint32_t frequency = (std::uint32_t)default_frequency,
                 bool zero_offset = false, int block_length = default_block_length) {
    close();
    m_error.clear();
    m_devname = fname;
    m_sample_rate = sample_rate;
    m_frequency = frequency;
    m_zero_offset = zero_offset;
    m_block_length = block_length;
    m_fp = std::fopen(fname.c_str(), "rb");
    if (!m_fp) return set_error("Failed to open " + fname + " : " + std::strerror(errno));
    if (raw) {
      m_container = Container::Raw;
      switch (format_type) {
      case FormatType::S8_LE: m_fmt = FMR_IQ_S8; break;
      case FormatType::S16_LE: m_fmt = FMR_IQ_S16; break;
      case FormatType::S24_LE: m_fmt = FMR_IQ_S24; break;
      case FormatType::U8_LE: m_fmt = FMR_IQ_U8; break;
      case FormatType::Float: m_fmt = FMR_IQ_CF32; break;
      default: return set_error("Unsupported sub type " + fname);
      }
      m_data_offset = 0;
      m_data_bytes = file_size();
    } else if (!parse_header()) {
      return false;
    }
    if (m_sample_rate == 0) return set_error("FileSource: sample rate must not be zero");
    // Limit too large block length (FileSource.cpp:236-244).
    const double rate_per_us = (double)m_sample_rate / 1e6;
    if ((double)m_block_length / rate_per_us > (double)max_expected_us) {
      m_block_length = (int)round_power((int)((double)max_expected_us * rate_per_us));
    }
    m_pos = 0;
    if (std::fseek(m_fp, (long)m_data_offset, SEEK_SET) != 0) return set_error("Failed to seek " + fname);
    return true;
  }

  std::uint32_t get_sample_rate() const { return m_sample_rate; }
  std::uint32_t get_frequency() const { return m_frequency; }
  bool is_low_if() const { return !m_zero_offset; }      // FileSource.cpp:275-276
  int get_block_length() const { return m_block_length; }
  Container container() const { return m_container; }
  int sample_format() const { return m_fmt; }            // FMR_IQ_*
  std::uint64_t total_samples() const { return m_data_bytes / (std::uint64_t)fmr_b200::iq_format_bytes(m_fmt); }
  operator bool() const { return m_error.empty(); }
  const std::string &error() const { return m_error; }

  // Up to block_length complex samples in the file's own sample format. Returns the number of complex
  // samples; 0 at end of file (an odd trailing item is dropped like n_read / 2 does, FileSource.cpp:519).
  std::uint32_t get_raw_block(std::vector<std::uint8_t> &raw) {
    if (!m_fp || m_block_length <= 0 || m_block_length > (1 << 24)) return 0; // FileSource.cpp:503-506
    const std::uint64_t esz = (std::uint64_t)fmr_b200::iq_format_bytes(m_fmt);
    std::uint64_t want = (std::uint64_t)m_block_length * esz;
    const std::uint64_t left = m_data_bytes - m_pos;
    if (want > left) want = left - left % esz;
    raw.resize((size_t)want);
    if (want == 0) return 0;
    const size_t got = std::fread(raw.data(), 1, (size_t)want, m_fp);
    const size_t whole = got - got % (size_t)esz;
    raw.resize(whole);
    m_pos += got;
    return (std::uint32_t)(whole / esz);
  }

  // FileSource::get_samples (FileSource.cpp:472-531): false at end of file.
  bool get_samples(IQSampleVector *samples) {
    if (!samples) return false;
    const std::uint32_t n = get_raw_block(m_rawbuf);
    if (n == 0) return false;
    samples->resize(n);
    fmr_b200::convert_iq(m_rawbuf.data(), m_fmt, n, samples->data());
    return true;
  }

  void close() {
    if (m_fp) std::fclose(m_fp);
    m_fp = nullptr;
  }

  static std::uint32_t round_power(int n) { // FileSource.cpp:251-267: largest power of two <= n
    if (n <= 0) return 0;
    std::uint32_t r = 1;
    while (n > 1) {
      r <<= 1;
      n >>= 1;
    }
    return r;
  }

private:
  bool set_error(const std::string &e) {
    m_error = e;
    close();
    return false;
  }
  std::uint64_t file_size() {
    const long cur = std::ftell(m_fp);
    std::fseek(m_fp, 0, SEEK_END);
    const long end = std::ftell(m_fp);
    std::fseek(m_fp, cur, SEEK_SET);
    return end > 0 ? (std::uint64_t)end : 0;
  }
  static std::uint32_t le32(const std::uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((std::uint32_t)p[3] << 24); }
  static std::uint64_t le64(const std::uint8_t *p) { return (std::uint64_t)le32(p) | ((std::uint64_t)le32(p + 4) << 32); }
  static std::uint32_t le16(const std::uint8_t *p) { return p[0] | (p[1] << 8); }

  // fmt chunk body -> sub-type. WAVE_FORMAT_PCM = 1, IEEE_FLOAT = 3, EXTENSIBLE = 0xFFFE (sub-format GUID's first
  // two bytes are the tag). WAV 8-bit PCM is unsigned.
  bool take_fmt_chunk(const std::uint8_t *b, std::uint64_t len) {
    if (len < 16) return set_error("Failed to open " + m_devname + " : truncated fmt chunk");
    std::uint32_t tag = le16(b);
    const std::uint32_t rate = le32(b + 4), bits = le16(b + 14);
    if (tag == 0xFFFE) {
      if (len < 40) return set_error("Failed to open " + m_devname + " : truncated extensible fmt chunk");
      tag = le16(b + 24);
      m_container = Container::WavEx;
    }
    int fmt = -1;
    if (tag == 1 && bits == 8) fmt = FMR_IQ_U8;
    if (tag == 1 && bits == 16) fmt = FMR_IQ_S16;
    if (tag == 1 && bits == 24) fmt = FMR_IQ_S24;
    if (tag == 3 && bits == 32) fmt = FMR_IQ_CF32;
    if (fmt < 0) { // PCM_32, DOUBLE, companded ...: FileSource.cpp:195-201
      char msg[96];
      std::snprintf(msg, sizeof(msg), " : tag %u, %u bits", tag, bits);
      return set_error("Unsupported sub type " + m_devname + msg);
    }
    m_fmt = fmt;
    if (rate != m_sample_rate) m_sample_rate = rate; // "overwrite sample rate" (FileSource.cpp:180-185)
    return true;
  }

  bool parse_header() {
    std::uint8_t h[40];
    const std::uint64_t fsz = file_size();
    if (std::fread(h, 1, 12, m_fp) != 12) return set_error("Failed to open " + m_devname + " : not a sound file");
    bool have_fmt = false;
    if (!std::memcmp(h, "RIFF", 4) && !std::memcmp(h + 8, "WAVE", 4)) {
      m_container = Container::Wav;
      std::uint64_t pos = 12;
      for (;;) {
        if (std::fseek(m_fp, (long)pos, SEEK_SET) != 0 || std::fread(h, 1, 8, m_fp) != 8) break;
        const std::uint64_t len = le32(h + 4);
        if (!std::memcmp(h, "fmt ", 4)) {
          std::uint8_t body[40] = {0};
          const size_t want = len < 40 ? (size_t)len : 40;
          if (std::fread(body, 1, want, m_fp) != want) break;
          if (!take_fmt_chunk(body, len)) return false;
          have_fmt = true;
        } else if (!std::memcmp(h, "data", 4)) {
          if (!have_fmt) break;
          m_data_offset = pos + 8;
          const std::uint64_t avail = fsz - m_data_offset;
          // a streamed file may carry 0 or 0xFFFFFFFF here: read to the end of the file
          m_data_bytes = (len == 0 || len == 0xFFFFFFFFu || len > avail) ? avail : len;
          return true;
        }
        pos += 8 + len + (len & 1);
      }
      return set_error("Failed to open " + m_devname + " : no fmt/data chunk");
    }
    static const std::uint8_t w64_riff[16] = {'r', 'i', 'f', 'f', 0x2E, 0x91, 0xCF, 0x11, 0xA5, 0xD6, 0x28, 0xDB, 0x04, 0xC1, 0x00, 0x00};
    static const std::uint8_t w64_tail[12] = {0xF3, 0xAC, 0xD3, 0x11, 0x8C, 0xD1, 0x00, 0xC0, 0x4F, 0x8E, 0xDB, 0x8A};
    std::fseek(m_fp, 0, SEEK_SET);
    if (std::fread(h, 1, 40, m_fp) == 40 && !std::memcmp(h, w64_riff, 16) && !std::memcmp(h + 24, "wave", 4) &&
        !std::memcmp(h + 28, w64_tail, 12)) {
      m_container = Container::W64;
      std::uint64_t pos = 40;
      for (;;) {
        std::uint8_t ch[24];
        if (std::fseek(m_fp, (long)pos, SEEK_SET) != 0 || std::fread(ch, 1, 24, m_fp) != 24) break;
        const std::uint64_t len = le64(ch + 16); // includes the 24-byte chunk header
        if (len < 24 || std::memcmp(ch + 4, w64_tail, 12)) break;
        if (!std::memcmp(ch, "fmt ", 4)) {
          std::uint8_t body[40] = {0};
          const std::uint64_t blen = len - 24;
          const size_t want = blen < 40 ? (size_t)blen : 40;
          if (std::fread(body, 1, want, m_fp) != want) break;
          if (!take_fmt_chunk(body, blen)) return false;
          m_container = Container::W64;
          have_fmt = true;
        } else if (!std::memcmp(ch, "data", 4)) {
          if (!have_fmt) break;
          m_data_offset = pos + 24;
          const std::uint64_t avail = fsz - m_data_offset, blen = len - 24;
          m_data_bytes = blen > avail ? avail : blen;
          return true;
        }
        pos += (len + 7) & ~(std::uint64_t)7;
      }
      return set_error("Failed to open " + m_devname + " : no fmt/data chunk");
    }
    // RF64, AIFF, FLAC ...: libsndfile would open them, FileSource then refuses the major format (FileSource.cpp:188-193)
    return set_error("Unsupported major format " + m_devname);
  }

  std::uint32_t m_sample_rate = default_sample_rate, m_frequency = (std::uint32_t)default_frequency;
  bool m_zero_offset = false;
  int m_block_length = default_block_length;
  std::string m_devname, m_error;
  std::FILE *m_fp = nullptr;
  Container m_container = Container::None;
  int m_fmt = FMR_IQ_S16;
  std::uint64_t m_data_offset = 0, m_data_bytes = 0, m_pos = 0;
  std::vector<std::uint8_t> m_rawbuf;
};

// libsndfile's format words as main.cpp:592-623 composes them (values of sndfile.h's enum).
#ifndef SNDFILE_H
enum {
  SF_FORMAT_WAV = 0x010000,
  SF_FORMAT_RAW = 0x040000,
  SF_FORMAT_RF64 = 0x220000,
  SF_FORMAT_PCM_16 = 0x0002,
  SF_FORMAT_FLOAT = 0x0006,
  SF_ENDIAN_LITTLE = 0x10000000,
  SF_FORMAT_SUBMASK = 0x0000FFFF,
  SF_FORMAT_TYPEMASK = 0x0FFF0000
};
#endif

class SndfileOutput {
public:
  // SndfileOutput::SndfileOutput (AudioOutput.cpp:33-77). "-" = stdout.
  SndfileOutput(const std::string &filename, unsigned int samplerate, bool stereo, int format)
      : m_channels(stereo ? 2 : 1), m_rate(samplerate) {
    const int major = format & SF_FORMAT_TYPEMASK, sub = format & SF_FORMAT_SUBMASK;
    if ((major != SF_FORMAT_WAV && major != SF_FORMAT_RF64 && major != SF_FORMAT_RAW) ||
        (sub != SF_FORMAT_PCM_16 && sub != SF_FORMAT_FLOAT)) {
      m_error = "SF_INFO for file '" + filename + "' is invalid";
      m_zombie = true;
      return;
    }
    m_header = (major != SF_FORMAT_RAW);
    m_out_format = (sub == SF_FORMAT_PCM_16) ? FMR_OUT_S16 : FMR_OUT_F32;
    if (filename == "-") {
      m_fp = stdout;
      m_seekable = false;
    } else {
      m_fp = std::fopen(filename.c_str(), "wb");
      if (!m_fp) {
        m_error = "can not open '" + filename + "' (" + std::strerror(errno) + ")";
        m_zombie = true;
        return;
      }
    }
    if (m_header) write_header();
  }
  ~SndfileOutput() {
    if (!m_closed) output_close();
  }
  SndfileOutput(const SndfileOutput &) = delete;
  SndfileOutput &operator=(const SndfileOutput &) = delete;

  int out_format() const { return m_out_format; } // FMR_OUT_S16 or FMR_OUT_F32: what write_native expects
  operator bool() const { return !m_zombie && m_error.empty(); }
  const std::string &error() const { return m_error; }

  // SndfileOutput::write (AudioOutput.cpp:153-167) = sf_write_double: PCM_16 lrint(x * 32767) without clipping,
  // FLOAT (float)x (libsndfile defaults: norm_double on, add_clipping off).
  bool write(const SampleVector &samples) {
    if (m_zombie) return false;
    const size_t n = samples.size();
    if (m_out_format == FMR_OUT_S16) {
      m_i16.resize(n);
      for (size_t i = 0; i < n; i++) m_i16[i] = (std::int16_t)std::lrint(samples[i] * 32767.0);
      return write_native(m_i16.data(), n);
    }
    m_f32.resize(n);
    for (size_t i = 0; i < n; i++) m_f32[i] = (float)samples[i];
    return write_native(m_f32.data(), n);
  }

  // Values already in the sink's sample format (out_format()), e.g. from fmr_fm_process_host_io.
  bool write_native(const void *values, size_t n_values) {
    if (m_zombie) return false;
    const size_t bytes = n_values * (m_out_format == FMR_OUT_S16 ? 2 : 4);
    if (bytes && std::fwrite(values, 1, bytes, m_fp) != bytes) {
      m_error = std::string("write failed (") + std::strerror(errno) + ")";
      return false;
    }
    m_data_bytes += bytes;
    if (m_header && m_seekable) { // SFC_SET_UPDATE_HEADER_AUTO (AudioOutput.cpp:91-100)
      write_header();
      std::fseek(m_fp, 0, SEEK_END);
    }
    return true;
  }

  void output_close() {
    if (m_fp && m_fp != stdout) std::fclose(m_fp);
    if (m_fp == stdout) std::fflush(stdout);
    m_fp = nullptr;
    m_closed = true;
  }

private:
  static void put32(std::uint8_t *p, std::uint32_t v) {
    p[0] = (std::uint8_t)v;
    p[1] = (std::uint8_t)(v >> 8);
    p[2] = (std::uint8_t)(v >> 16);
    p[3] = (std::uint8_t)(v >> 24);
  }
  static void put16(std::uint8_t *p, std::uint32_t v) {
    p[0] = (std::uint8_t)v;
    p[1] = (std::uint8_t)(v >> 8);
  }
  void write_header() {
    std::uint8_t h[44];
    const std::uint32_t bits = (m_out_format == FMR_OUT_S16) ? 16 : 32, align = m_channels * bits / 8;
    // a pipe cannot be patched afterwards: announce "until end of stream"
    const std::uint64_t d = m_seekable ? m_data_bytes : 0xFFFFFFFFull;
    const std::uint32_t dlen = d > 0xFFFFFFFFull ? 0xFFFFFFFFu : (std::uint32_t)d;
    std::memcpy(h, "RIFF", 4);
    put32(h + 4, dlen > 0xFFFFFFFFu - 36 ? 0xFFFFFFFFu : dlen + 36);
    std::memcpy(h + 8, "WAVEfmt ", 8);
    put32(h + 16, 16);
    put16(h + 20, m_out_format == FMR_OUT_S16 ? 1 : 3);
    put16(h + 22, m_channels);
    put32(h + 24, m_rate);
    put32(h + 28, m_rate * align);
    put16(h + 32, align);
    put16(h + 34, bits);
    std::memcpy(h + 36, "data", 4);
    put32(h + 40, dlen);
    if (m_seekable) std::fseek(m_fp, 0, SEEK_SET);
    if (m_seekable || m_data_bytes == 0) std::fwrite(h, 1, 44, m_fp);
  }

  unsigned int m_channels, m_rate;
  int m_out_format = FMR_OUT_S16;
  bool m_header = true, m_seekable = true, m_zombie = false, m_closed = false;
  std::FILE *m_fp = nullptr;
  std::uint64_t m_data_bytes = 0;
  std::string m_error;
  std::vector<std::int16_t> m_i16;
  std::vector<float> m_f32;
};

namespace fmr_b200 {

// The block loop's running levels (main.cpp:872-876,950,976,996,1028).
struct BlockLoopLevels {
  float if_level = 0.f;
  float audio_level = 0.f;
  // feed one block; returns false when the block produced no IF samples (the loop `continue`s, main.cpp:933-936)
  bool feed(const fmr_block_level_t &l, bool audio_exists) {
    if (l.if_rms < 0.f) return false;
    if_level = 0.75 * if_level + 0.25 * (double)l.if_rms;                        // main.cpp:976
    if (audio_exists) audio_level = 0.95 * audio_level + 0.05 * l.audio_rms;     // main.cpp:996
    return true;
  }
  float if_level_db() const { return 20 * std::log10(if_level + 1e-9); }            // main.cpp:950
  float audio_level_db() const { return 20 * std::log10(audio_level + 1e-9) + 3.01; } // main.cpp:1028
};

// "{:>8} {:>14} {:18.6f} {:+9.3f}" (main.cpp:1089-1091)
inline std::string format_pps_line(std::uint64_t pps_index, std::uint64_t sample_index, double ts, double if_level_db) {
  char buf[96];
  std::snprintf(buf, sizeof(buf), "%8llu %14llu %18.6f %+9.3f", (unsigned long long)pps_index,
                (unsigned long long)sample_index, ts, if_level_db);
  return buf;
}

// Squelch level from -l dB (main.cpp:484-489): pow(10, -(dB / 20)), 0 when squelch is not requested.
inline double squelch_level_from_db(double squelch_level_db, bool enabled) {
  return enabled ? std::pow(10.0, -(squelch_level_db / 20.0)) : 0.0;
}

} // namespace fmr_b200
#endif
