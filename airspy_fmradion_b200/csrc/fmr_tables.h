// fmr_tables.h — coefficient tables and the integer model of the reference's resampler
// bookkeeping (host side).
//
// The r8brain chain the reference builds for a rate pair (CDSPResampler.h:117-394) is, for
// every rate this library ships, of the shape
//     [half-band /2]* -> long zero-phase low-pass (optionally /2) -> [whole-step polyphase bank]
// Each stage is a zero-phase (time-aligned) linear filter; what makes the reference's
// per-call output sizes (0,0,...,34,78,79,78,...) is only WHEN each stage releases its
// outputs. That is pure integer bookkeeping and depends on the cumulative input count
// alone, not on how the input was chunked:
//   half-band, n taps  (CDSPHBDownsampler.h:166-238):  out(N) = max(0, floor(N/2) - (n-1))
//   block convolver    (CDSPBlockConvolver.h:252-353): out(N) = max(0, N - Latency), then
//                      every 2nd sample if it decimates inside the FFT
//   whole-step bank    (CDSPFracInterpolator.h:861-925,992+): output m is released once
//                      floor(m*InStep/OutStep) + FilterLen/2 + 1 inputs have arrived.
// tests/test_schedule.py checks these formulas against the compiled reference.
#ifndef FMR_TABLES_H
#define FMR_TABLES_H

#include <cstdint>
#include <cstdlib>

namespace fmr {

struct HbStage {
  int ntaps;
  const double *taps;
};
struct BcStage {
  int klen;      // kernel length (odd), zero-phase, includes the chain's final gain
  int inputlen;  // r8brain's FFT block payload (only documents where Latency comes from)
  int latency;   // samples of the convolver's output withheld at stream start
  int down;      // 1, 2 or 3
  int outoffset; // (klen-1)/2
  const double *taps;
};
struct FiStage {
  int instep, outstep, flen;
  const double *taps; // [outstep][flen]
};
struct ChainDesc {
  double src, dst;
  int kind; // 0: IfResampler spec (180.15 dB), 1: AudioResampler spec (206.91 dB)
  int n_hb;
  HbStage hb[3];
  BcStage bc;
  int has_fi;
  FiStage fi;
  int verified; // 1: the GPU path for this pair has passed the parity suite on a B200 (tools/gen_tables.py)
};

#include "fmr_tables_generated.inc"

// Every pair in kChains has passed the GPU parity suite on a B200 (profiles/pytest_newrates_r02.log); a pair whose
// `verified` flag is 0 (tables shipped ahead of a GPU run, tools/gen_tables.py) is refused like a pair without tables.
inline const ChainDesc *find_chain(double src, double dst, int kind) {
  for (int i = 0; i < kNumChains; i++) {
    if (kChains[i].src == src && kChains[i].dst == dst && kChains[i].kind == kind) {
      return kChains[i].verified ? &kChains[i] : nullptr;
    }
  }
  return nullptr;
}

// Cumulative stage output counts as a function of cumulative stage input counts.
inline int64_t hb_out(const ChainDesc *d, int64_t n) {
  for (int s = 0; s < d->n_hb; s++) {
    n = n / 2 - (d->hb[s].ntaps - 1);
    if (n < 0) n = 0;
  }
  return n;
}
inline int64_t bc_out(const ChainDesc *d, int64_t n) {
  int64_t c = n - d->bc.latency;
  if (c < 0) c = 0;
  if (d->bc.down > 1) c = (c + d->bc.down - 1) / d->bc.down; // every down-th sample, the first one kept
  return c;
}
inline int64_t fi_out(const ChainDesc *d, int64_t n) {
  if (!d->has_fi) return n;
  const int fl2 = d->fi.flen / 2;
  if (n < fl2 + 1) return 0;
  const int64_t a = (n - fl2) * (int64_t)d->fi.outstep;
  return (a + d->fi.instep - 1) / d->fi.instep;
}
inline int64_t chain_out(const ChainDesc *d, int64_t n) {
  return fi_out(d, bc_out(d, hb_out(d, n)));
}

} // namespace fmr
#endif
