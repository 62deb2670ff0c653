// End-to-end drop-in check of the file path (SURVEY.md §8 f1 + path + f4), written like the reference's block loop
// (main.cpp:879-1002) but with super-blocks: FileSource -> [GPU: sample-format decode, IfResampler, FmDecoder,
// level metering, squelch gain, sink format] -> SndfileOutput. Only the file's own bytes go to the device and only
// the sink's int16 comes back.
//   usage: file_decode "<FileSource configuration>" <out.wav> <blocks per call> <squelch dB or -1>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../airspy_fmradion_b200/host/fmradion_b200_io.hpp"

#define CHECK(expr)                                                        \
  do {                                                                     \
    fmr_status _s = (expr);                                                \
    if (_s != FMR_OK) {                                                    \
      fprintf(stderr, "%s: %s\n", #expr, fmr_last_error());                \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  const uint32_t per_call = (uint32_t)atoi(argv[3]);
  const double squelch_db = atof(argv[4]);
  FileSource src(0);
  if (!src.configure(std::string(argv[1]))) {
    fprintf(stderr, "FileSource: %s\n", src.error().c_str());
    return 1;
  }
  const uint32_t blklen = (uint32_t)src.get_block_length();
  const int fmt = src.sample_format();
  const size_t esz = (size_t)fmr_b200::iq_format_bytes(fmt);
  fmr_fm_config cfg{};
  cfg.input_rate = (double)src.get_sample_rate();
  cfg.fs4_shift = src.is_low_if() ? 0 : 1; // main.cpp:677,773
  cfg.stereo = 1;
  cfg.deemphasis_us = 50.0;
  cfg.n_channels = 1;
  cfg.max_samples_per_call = blklen * per_call;
  cfg.max_blocks_per_call = per_call;
  fmr_fm *h = nullptr;
  CHECK(fmr_fm_create(&cfg, &h));
  SndfileOutput out(argv[2], 48000, true, SF_FORMAT_RF64 | SF_FORMAT_PCM_16 | SF_ENDIAN_LITTLE); // main.cpp:610-616
  if (!out) {
    fprintf(stderr, "SndfileOutput: %s\n", out.error().c_str());
    return 1;
  }
  fmr_output_config oc{};
  oc.out_format = out.out_format();
  oc.squelch_level = fmr_b200::squelch_level_from_db(squelch_db, squelch_db >= 0);
  oc.gain = 0.5;
  std::vector<uint8_t> super(esz * blklen * per_call), raw;
  std::vector<uint32_t> block_len, audio_len(per_call);
  std::vector<int16_t> pcm(48000);
  std::vector<fmr_block_level_t> lv(per_call);
  fmr_b200::BlockLoopLevels levels;
  size_t blocks = 0, written = 0, muted = 0;
  bool eof = false;
  while (!eof) {
    block_len.clear();
    size_t fill = 0;
    while (block_len.size() < per_call) {
      const uint32_t n = src.get_raw_block(raw);
      if (n == 0) {
        eof = true;
        break;
      }
      memcpy(super.data() + fill, raw.data(), raw.size());
      fill += raw.size();
      block_len.push_back(n);
    }
    const uint32_t nb = (uint32_t)block_len.size();
    if (nb == 0) break;
    uint64_t total = 0;
    CHECK(fmr_fm_query_output(h, block_len.data(), nb, &total, nullptr));
    if (pcm.size() < total) pcm.resize((size_t)total);
    CHECK(fmr_fm_process_host_io(h, super.data(), fmt, fill / esz, block_len.data(), nb, &oc, pcm.data(), pcm.size(),
                                 audio_len.data()));
    CHECK(fmr_fm_block_levels(h, 0, lv.data(), nb));
    size_t off = 0;
    for (uint32_t b = 0; b < nb; b++, blocks++) {
      if (!levels.feed(lv[b], audio_len[b] > 0)) continue; // no IF samples yet (main.cpp:933-936)
      if (audio_len[b] == 0) continue;                     // main.cpp:981-984
      if (lv[b].gain == 0.f) muted++;
      if (!out.write_native(pcm.data() + off, audio_len[b])) { // audio_output->write (main.cpp:1002)
        fprintf(stderr, "write: %s\n", out.error().c_str());
        return 1;
      }
      off += audio_len[b];
      written += audio_len[b];
    }
  }
  fmr_fm_stats_t st;
  CHECK(fmr_fm_stats(h, 0, &st));
  out.output_close();
  printf("blocks=%zu written=%zu muted=%zu stereo=%d if_level_db=%.3f audio_level_db=%.3f\n", blocks, written, muted,
         st.stereo_detected, levels.if_level_db(), levels.audio_level_db());
  fmr_fm_destroy(h);
  return 0;
}
