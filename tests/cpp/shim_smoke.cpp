// Drop-in check: drives the shim's FmDecoder / AmDecoder exactly like main.cpp:812-830,953-974
// drives the reference's classes (one block per process() call, SampleVector out, getters),
// reads IQ from a raw cf32 file and writes the audio as raw f64.
//   usage: shim_smoke fm|am <input_rate> <iq.cf32> <audio.f64> [blocklen]
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../airspy_fmradion_b200/host/fmradion_b200_shim.hpp"

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  const bool fm = !strcmp(argv[1], "fm");
  const double rate = atof(argv[2]);
  const size_t blk = argc > 5 ? (size_t)atoi(argv[5]) : 2048;
  FILE *fi = fopen(argv[3], "rb"), *fo = fopen(argv[4], "wb");
  if (!fi || !fo) return 3;
  IQSampleCoeff fmfilter_coeff = {0.0f, 1.0f, 0.0f}; // delay_3taps_only_iq, unused (fmfilter_enable=false)
  IQSampleCoeff amfilter_coeff(255, 0.0f);
  amfilter_coeff[127] = 1.0f; // identity channel filter for the AM smoke run
  FmDecoder *fmd = nullptr;
  AmDecoder *amd = nullptr;
  if (fm) {
    fmd = new FmDecoder(false, fmfilter_coeff, true, FmDecoder::deemphasis_time_eu, false, 0, rate, false);
  } else {
    amd = new AmDecoder(amfilter_coeff, ModType::AM, rate, false);
  }
  IQSampleVector block(blk);
  SampleVector audio;
  size_t total = 0, calls = 0;
  while (true) {
    size_t n = fread(block.data(), sizeof(IQSample), blk, fi);
    if (n == 0) break;
    IQSampleVector in(block.begin(), block.begin() + n);
    if (fm) {
      fmd->process(std::move(in), audio);
    } else {
      amd->process(std::move(in), audio);
    }
    if (!audio.empty()) fwrite(audio.data(), sizeof(double), audio.size(), fo);
    total += audio.size();
    calls++;
  }
  if (fm) {
    printf("calls=%zu audio=%zu stereo=%d pilot=%.6f if_rms=%.6f tuning=%.3f\n", calls, total, (int)fmd->stereo_detected(),
           fmd->get_pilot_level(), fmd->get_if_rms(), fmd->get_tuning_offset());
  } else {
    printf("calls=%zu audio=%zu if_rms=%.6f afgain=%.6f\n", calls, total, amd->get_if_rms(),
           amd->get_af_agc_current_gain());
  }
  delete fmd;
  delete amd;
  fclose(fi);
  fclose(fo);
  return 0;
}
