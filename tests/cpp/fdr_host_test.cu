// Host emulation of the frequency-domain resampler (airspy_fmradion_b200/csrc/fmr_fdr.cuh): the per-thread pass bodies
// are __host__ __device__, so the identical index algebra runs here on the CPU - pass by pass, the "threads" of a pass
// in a scrambled order - and the block's outputs are compared with what the reference computes: the 2307-tap zero-phase
// low-pass (CDSPBlockConvolver.h:252-353) followed by the 192 x 18 polyphase bank (CDSPFracInterpolator.h:861-925),
// evaluated directly in double from the frozen r8brain tables. Input is white noise over the whole 1.25 MHz band (the
// half-band cascade leaves everything up to 625 kHz in the stream). Built with nvcc, runs without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../airspy_fmradion_b200/csrc/fmr_fdr.cuh"
#include "../../airspy_fmradion_b200/csrc/fmr_tables.h"

using namespace fmr::fdr;

template <int RL> static int run_pair(double src, double dst, int kind, int instep, int outstep) {
  using G = Geo<RL>;
  const fmr::ChainDesc *d = fmr::find_chain(src, dst, kind);
  if (!d || d->fi.instep != instep || d->fi.outstep != outstep) {
    printf("FAIL: chain %g -> %g\n", src, dst);
    return 1;
  }
  const int klen = d->bc.klen, fl2 = (klen - 1) / 2, flen = d->fi.flen;
  const int need = fl2 + flen / 2 + 1;
  const int guard_in = (need + instep - 1) / instep * instep, adv_in = kNin - 2 * guard_in;
  const int guard_out = guard_in / instep * outstep, adv_out = adv_in / instep * outstep;
  std::vector<float2> tab;
  std::vector<float> Hs;
  fdr_make_tables<RL>(d->bc.taps, klen, tab, Hs);
  std::mt19937 rng(11);
  std::normal_distribution<float> nd(0.f, 0.5f);
  const long long base = 5LL * adv_in - guard_in; // block j = 5
  const int pre = 64, total = kNin + 2 * pre;
  std::vector<float2> x(total);
  for (auto &v : x) v = fmr::ipfft::mk(nd(rng), nd(rng));
  auto X = [&](long long t) -> const float2 & { return x[(size_t)(t - base + pre)]; };
  // ---- emulated kernel
  std::vector<float2> A(kNin), Z(G::kZLen, fmr::ipfft::mk(NAN, NAN)), out(G::kNout);
  std::vector<int> order;
  auto scrambled = [&](int n) {
    order.resize(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::shuffle(order.begin(), order.end(), rng);
  };
  scrambled(625);
  for (int b : order) fwd1(b, [&](int bb, int a) { return X(base + bb + 625 * a); }, A.data(), tab.data());
  scrambled(400);
  for (int i : order) fwd2(i, A.data(), tab.data());
  scrambled(400);
  for (int i : order) {
    float2 o[G::kKeep];
    fwd3_compute<RL>(i, A.data(), Hs.data(), o);
    fwd3_store<RL>(i, Z.data(), o);
  }
  int nan_pads = 0, nan_data = 0;
  for (int p = 0; p < G::kZLen; p++) {
    const bool is_pad = (p % (G::kChunk + 1)) == G::kChunk;
    if (std::isnan(Z[p].x)) (is_pad ? nan_pads : nan_data)++;
  }
  if (nan_data != 0 || nan_pads != 16) {
    printf("FAIL: band scatter left %d bins unwritten, %d pads untouched (want 0 / 16)\n", nan_data, nan_pads);
    return 1;
  }
  scrambled(G::kChunk);
  for (int b : order) inv1<RL>(b, Z.data(), tab.data());
  scrambled(G::kChunk);
  for (int u : order) inv2<RL>(u, Z.data(), tab.data());
  std::vector<int> hits(G::kNout, 0);
  scrambled(256);
  for (int t : order) {
    inv3<RL>(t, Z.data(), [&](int i, float2 v) {
      out[i] = v;
      hits[i]++;
    });
  }
  for (int i = 0; i < G::kNout; i++) {
    if (hits[i] != 1) {
      printf("FAIL: output %d written %d times\n", i, hits[i]);
      return 1;
    }
  }
  // ---- the reference's two stages, directly, in double
  const long long m0 = base / instep * outstep; // output index of block sample 0
  double maxerr = 0, rms = 0, ref_rms = 0;
  int cnt = 0;
  for (int i = guard_out; i < guard_out + adv_out; i += ((i < guard_out + 48 || i >= guard_out + adv_out - 49) ? 1 : 7)) {
    const long long m = m0 + i;
    const long long q = (m * instep) / outstep - (flen / 2 - 1);
    const int ph = (int)((m * instep) % outstep);
    double zr = 0, zi = 0;
    for (int k = 0; k < flen; k++) {
      double yr = 0, yi = 0;
      for (int t = 0; t < klen; t++) {
        const float2 &v = X(q + k + fl2 - t);
        yr += d->bc.taps[t] * v.x;
        yi += d->bc.taps[t] * v.y;
      }
      zr += d->fi.taps[ph * flen + k] * yr;
      zi += d->fi.taps[ph * flen + k] * yi;
    }
    const double er = out[i].x - zr, ei = out[i].y - zi;
    maxerr = std::max(maxerr, std::max(std::fabs(er), std::fabs(ei)));
    rms += er * er + ei * ei;
    ref_rms += zr * zr + zi * zi;
    cnt++;
  }
  rms = std::sqrt(rms / (2 * cnt));
  ref_rms = std::sqrt(ref_rms / (2 * cnt));
  printf("fdr %d:%d (%g -> %g, %d-point inverse, guard %d): %d outputs checked, reference rms %.4g, max |err| %.3e, rms err %.3e\n",
         instep, outstep, src, dst, G::kNout, guard_in, cnt, ref_rms, maxerr, rms);
  if (!(maxerr < 6e-7 && rms < 1.5e-7)) {
    printf("FAIL: error too large\n");
    return 1;
  }
  return 0;
}

int main() {
  int bad = 0;
  bad += run_pair<12>(1.0e7, 384000.0, 0, 625, 192);
  bad += run_pair<15>(1.0e6, 384000.0, 0, 125, 48);
  bad += run_pair<10>(6.0e6, 384000.0, 0, 125, 32);
  bad += run_pair<15>(1.0e6, 48000.0, 0, 125, 48);
  if (bad) return 1;
  printf("fdr host emulation: ok\n");
  return 0;
}
