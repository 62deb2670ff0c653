// Host-only driver for airspy_fmradion_b200/host/fmradion_b200_io.hpp (no CUDA, no library):
//   io_host_test read  "<FileSource configuration string>" <out.cf32>   -> all blocks through get_samples()
//   io_host_test raw   "<FileSource configuration string>" <out.bin>    -> all blocks through get_raw_block()
//   io_host_test write <in.f64> <out file> <rate> <stereo 0|1> <wav16|wavf32|raw16|rawf32> <values per write>
//   io_host_test misc                                                    -> parse / format helpers
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../airspy_fmradion_b200/host/fmradion_b200_io.hpp"

static const char *container_name(FileSource::Container c) {
  switch (c) {
  case FileSource::Container::Wav: return "WAV";
  case FileSource::Container::WavEx: return "WAVEX";
  case FileSource::Container::W64: return "W64";
  case FileSource::Container::Raw: return "RAW";
  default: return "NONE";
  }
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string cmd = argv[1];
  if (cmd == "read" || cmd == "raw") {
    if (argc < 4) return 2;
    FileSource src(0);
    if (!src.configure(std::string(argv[2]))) {
      printf("configure failed: %s\n", src.error().c_str());
      return 1;
    }
    printf("rate=%u freq=%u low_if=%d blklen=%d fmt=%d container=%s total=%llu\n", src.get_sample_rate(),
           src.get_frequency(), (int)src.is_low_if(), src.get_block_length(), src.sample_format(),
           container_name(src.container()), (unsigned long long)src.total_samples());
    FILE *fo = fopen(argv[3], "wb");
    if (!fo) return 3;
    size_t blocks = 0, last = 0;
    if (cmd == "read") {
      IQSampleVector v;
      while (src.get_samples(&v)) {
        fwrite(v.data(), sizeof(IQSample), v.size(), fo);
        blocks++;
        last = v.size();
      }
    } else {
      std::vector<uint8_t> raw;
      uint32_t n;
      while ((n = src.get_raw_block(raw)) > 0) {
        fwrite(raw.data(), 1, raw.size(), fo);
        blocks++;
        last = n;
      }
    }
    fclose(fo);
    printf("blocks=%zu last=%zu\n", blocks, last);
    return 0;
  }
  if (cmd == "write") {
    if (argc < 8) return 2;
    FILE *fi = fopen(argv[2], "rb");
    if (!fi) return 3;
    const unsigned rate = (unsigned)atoi(argv[4]);
    const bool stereo = atoi(argv[5]) != 0;
    const std::string kind = argv[6];
    const size_t per = (size_t)atoi(argv[7]);
    int format = 0; // as main.cpp:592-623 composes them
    if (kind == "wav16") format = SF_FORMAT_RF64 | SF_FORMAT_PCM_16 | SF_ENDIAN_LITTLE;
    if (kind == "wavf32") format = SF_FORMAT_RF64 | SF_FORMAT_FLOAT | SF_ENDIAN_LITTLE;
    if (kind == "raw16") format = SF_FORMAT_RAW | SF_FORMAT_PCM_16 | SF_ENDIAN_LITTLE;
    if (kind == "rawf32") format = SF_FORMAT_RAW | SF_FORMAT_FLOAT | SF_ENDIAN_LITTLE;
    SndfileOutput out(argv[3], rate, stereo, format);
    if (!out) {
      printf("open failed: %s\n", out.error().c_str());
      return 1;
    }
    SampleVector v(per);
    size_t n, total = 0;
    while ((n = fread(v.data(), sizeof(double), per, fi)) > 0) {
      v.resize(n);
      if (!out.write(v)) {
        printf("write failed: %s\n", out.error().c_str());
        return 1;
      }
      total += n;
      v.resize(per);
    }
    fclose(fi);
    out.output_close();
    printf("written=%zu out_format=%d\n", total, out.out_format());
    return 0;
  }
  if (cmd == "misc") {
    auto m = fmr_b200::parse_config_string("alpha=100,beta,gamma=x=yz,,delta=");
    printf("map:");
    for (auto &kv : m) printf(" [%s]=[%s]", kv.first.c_str(), kv.second.c_str());
    printf("\n");
    int v = -1;
    printf("int: %d", (int)fmr_b200::parse_int("10000k", v, true));
    printf(" %d", v);
    printf(" %d", (int)fmr_b200::parse_int("10k", v, false));
    printf(" %d", (int)fmr_b200::parse_int("12x", v, true));
    printf(" %d", (int)fmr_b200::parse_int("", v, true));
    const bool neg_ok = fmr_b200::parse_int("-42", v);
    printf(" %d %d\n", (int)neg_ok, v);
    printf("round_power: %u %u %u %u\n", FileSource::round_power(0), FileSource::round_power(1), FileSource::round_power(480),
           FileSource::round_power(4096));
    printf("pps: [%s]\n", fmr_b200::format_pps_line(3, 1234567, 1700000000.25, -12.3456).c_str());
    printf("squelch: %.9g %.9g\n", fmr_b200::squelch_level_from_db(40.0, true), fmr_b200::squelch_level_from_db(40.0, false));
    fmr_b200::BlockLoopLevels lv;
    fmr_block_level_t a{0.5f, 0.f, 0.25f, 0.5f}, none{-1.f, 0.f, 0.f, 0.f};
    const bool r0 = lv.feed(none, false), r1 = lv.feed(a, true), r2 = lv.feed(a, false);
    printf("levels: %d %d %d %.9g %.9g %.6f %.6f\n", (int)r0, (int)r1, (int)r2, lv.if_level, lv.audio_level, lv.if_level_db(),
           lv.audio_level_db());
    return 0;
  }
  return 2;
}
