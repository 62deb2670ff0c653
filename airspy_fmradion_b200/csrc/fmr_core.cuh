// fmr_core.cuh — the 384 kHz core of FmDecoder::process (FmDecode.cpp:85-183, stereo, no
// multipath filter) as ONE warp-specialised kernel.
//
// The core is three strictly serial recurrences per channel (IfSimpleAgc, PilotPhaseLock, the
// deemphasis filters) with time-parallel work in between (PhaseDiscriminator, the L-R mix,
// statistics). As separate launches (k_fm_agc2 -> k_fm_disc/k_fm_call_stats -> k_fm_pll2) their
// latencies add up and every hand-over goes through HBM. Here a CTA owns 32 channels (lane =
// channel) and three warps run as a pipeline over chunks of kCfT samples:
//   warp 0  AGC           global IF ring -> gain recurrence -> shared memory (float2)
//   warp 1  disc + post   atan2/phase difference/statistics -> shared memory (MPX);
//                         and, two chunks behind, 2*x*sin(2 phi), both deemphasis filters -> HBM
//   warp 2  PLL           the pilot phase lock recurrence only (its dependent chain is the
//                         critical path of the whole kernel); hands sin/cos of the pilot phase on
// Chunks are handed over through double-buffered shared memory with named barriers
// (bar.arrive / bar.sync, producer/consumer pairs of 64 threads); nothing of the core except
// the (mono, L-R) result touches HBM. Step time = the PLL chain alone instead of
// AGC + discriminator + PLL.
//
// Arithmetic follows k_fm_agc2 / k_fm_disc / k_fm_pll2 expression for expression; the per-call
// statistics (IF RMS, baseband mean/RMS; getters only) are accumulated sequentially per channel
// instead of by a parallel reduction.
#ifndef FMR_CORE_CUH
#define FMR_CORE_CUH

#include "fmr_kernels.cuh"

namespace fmr {

constexpr int kCfT = 8;        // samples per chunk
constexpr int kCfThreads = 96; // three warps

struct CfSmem {
  float2 iq[2][kCfT][32];
  float mpx[2][kCfT][32];
  double ps[2][kCfT][32];
  double pc[2][kCfT][32];
  float2 tab[256];
};

// named barriers: producer/consumer pairs of two warps
__device__ __forceinline__ void cf_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void cf_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
// link 0: AGC -> disc, link 1: disc -> PLL, link 2: PLL -> post
__device__ __forceinline__ int cf_full(int link, int slot) { return 1 + link * 4 + slot; }
__device__ __forceinline__ int cf_empty(int link, int slot) { return 3 + link * 4 + slot; }

// Walks the cumulative call table over the flat sample index of a launch (empty calls skipped).
struct CfCalls {
  const uint32_t *__restrict__ ce;
  int n_calls, b;
  uint32_t beg, end;
  __device__ __forceinline__ void init(const uint32_t *call_end, int n) {
    ce = call_end;
    n_calls = n;
    b = -1;
    beg = 0;
    end = 0;
  }
  // p == end: move to the next non-empty call; returns false when there is none
  __device__ __forceinline__ bool next() {
    const uint32_t e0 = end;
    do {
      b++;
    } while (b < n_calls && ce[b] == e0);
    if (b >= n_calls) return false;
    beg = e0;
    end = ce[b];
    return true;
  }
};

static __global__ void __launch_bounds__(kCfThreads)
    k_fm_core_fused(Ring<float2> if_raw, Ring<float2> iq_in, Ring<double2> out384, FmChanState *__restrict__ st,
                    uint8_t *__restrict__ flags, PpsEventDev *__restrict__ pps, const uint32_t *__restrict__ call_end,
                    int n_calls, int64_t t0, FmCoreParams P, const float *__restrict__ atan_tbl, int block_off,
                    int reset_pps) {
  __shared__ CfSmem S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 256; i += kCfThreads) S.tab[i] = make_float2(atan_tbl[i], atan_tbl[i + 1] - atan_tbl[i]);
  __syncthreads();
  const int c_raw = blockIdx.x * 32 + lane;
  const bool act = c_raw < P.n_channels;
  const int c = act ? c_raw : (P.n_channels - 1); // idle lanes shadow the last channel, store nothing
  const int n_total = n_calls ? (int)call_end[n_calls - 1] : 0;
  const int K = (n_total + kCfT - 1) / kCfT;
  const uint32_t t0lo = (uint32_t)t0;

  if (warp == 0) {
    // =============================== AGC (IfSimpleAgc.cpp:37-57) + IF RMS (Utility.h:118-132)
    const float2 *__restrict__ irow = iq_in.base + (size_t)c * iq_in.cap;
    const float2 *__restrict__ rrow = if_raw.base + (size_t)c * if_raw.cap;
    const bool two = (if_raw.base != iq_in.base); // IF filter enabled: statistics on the unfiltered input
    const uint32_t imask = iq_in.cap - 1, rmask = if_raw.cap - 1;
    float g = st[c].agc_gain, if_rms = st[c].if_rms, sq = 0.f;
    const double rate = (double)P.agc_rate;
    const float gmax = P.agc_max;
    CfCalls cl;
    cl.init(call_end, n_calls);
    float2 nx[kCfT], nr[kCfT];
#pragma unroll
    for (int u = 0; u < kCfT; u++) {
      nx[u] = (u < n_total) ? irow[(t0lo + (uint32_t)u) & imask] : make_float2(0.f, 0.f);
      nr[u] = (two && u < n_total) ? rrow[(t0lo + (uint32_t)u) & rmask] : nx[u];
    }
    for (int k = 0; k < K; k++) {
      const int slot = k & 1, p0 = k * kCfT;
      float2 x[kCfT], xr[kCfT];
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        x[u] = nx[u];
        xr[u] = nr[u];
      }
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        const int p = p0 + kCfT + u;
        nx[u] = (p < n_total) ? irow[(t0lo + (uint32_t)p) & imask] : make_float2(0.f, 0.f);
        nr[u] = (two && p < n_total) ? rrow[(t0lo + (uint32_t)p) & rmask] : nx[u];
      }
      cf_sync(cf_empty(0, slot));
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        const int p = p0 + u;
        if (p < n_total) {
          if ((uint32_t)p == cl.end) cl.next();
          float2 x2;
          x2.x = x[u].x * g;
          x2.y = x[u].y * g;
          const float nrm = x2.x * x2.x + x2.y * x2.y;
          const float z = (float)(1.0 + (rate * (1.0 - (double)nrm)));
          g *= z;
          g = isfinite(g) ? ((g > gmax) ? gmax : g) : 1.0f;
          S.iq[slot][u][lane] = x2;
          sq += xr[u].x * xr[u].x + xr[u].y * xr[u].y;
          if ((uint32_t)(p + 1) == cl.end) {
            if_rms = sqrtf(sq / (float)(cl.end - cl.beg));
            sq = 0.f;
          }
        }
      }
      cf_arrive(cf_full(0, slot));
    }
    if (act) {
      st[c].agc_gain = g;
      st[c].if_rms = if_rms;
    }
  } else if (warp == 1) {
    // =============================== discriminator + statistics, and the post stage
    FmChanState *sp = st + c;
    float prev = sp->disc_prev;
    float bmean = sp->baseband_mean, blevel = sp->baseband_level;
    double dem = sp->de_m_x1, des = sp->de_s_x1;
    float vs = 0.f, vq = 0.f;
    const bool shift = P.pilot_shift != 0, de_st = P.deemph_on_stereo != 0;
    double2 *__restrict__ orow = out384.base + (size_t)c * out384.cap;
    const uint32_t omask = out384.cap - 1;
    CfCalls cl;
    cl.init(call_end, n_calls);
    cf_arrive(cf_empty(0, 0));
    cf_arrive(cf_empty(0, 1));
    cf_arrive(cf_empty(2, 0));
    cf_arrive(cf_empty(2, 1));
    for (int j = 0; j <= K; j++) {
      if (j < K) {
        // ---- PhaseDiscriminator::process (PhaseDiscriminator.cpp:33-46) on chunk j
        const int slot = j & 1, p0 = j * kCfT;
        cf_sync(cf_full(0, slot));
        float2 x[kCfT];
#pragma unroll
        for (int u = 0; u < kCfT; u++) x[u] = S.iq[slot][u][lane];
        if (j + 2 < K) cf_arrive(cf_empty(0, slot));
        float ph[kCfT];
#pragma unroll
        for (int u = 0; u < kCfT; u++) ph[u] = atan2f(x[u].y, x[u].x) * P.disc_inv_norm;
        cf_sync(cf_empty(1, slot));
#pragma unroll
        for (int u = 0; u < kCfT; u++) {
          const int p = p0 + u;
          if (p < n_total) {
            if ((uint32_t)p == cl.end) cl.next();
            float d = ph[u] - prev;
            prev = ph[u];
            if (d > P.disc_bound) d -= 2 * P.disc_bound;
            if (d < -P.disc_bound) d += 2 * P.disc_bound;
            if (isnan(d)) d = 0.0f;
            S.mpx[slot][u][lane] = d;
            vs += d;
            vq += d * d;
            if ((uint32_t)(p + 1) == cl.end) {
              // samples_mean_rms + the EMA of FmDecode.cpp:146-150
              const float n = (float)(cl.end - cl.beg);
              const float mean = vs / n, rms = sqrtf(vq / n);
              bmean = (float)(0.95 * (double)bmean + 0.05 * (double)mean);
              blevel = (float)(0.95 * (double)blevel + 0.05 * (double)rms);
              vs = 0.f;
              vq = 0.f;
            }
          }
        }
        cf_arrive(cf_full(1, slot));
      }
      if (j >= 1) {
        // ---- post stage on chunk j-1: demod_stereo (FmDecode.cpp:224-239) + LowPassFilterRC x2
        const int kk = j - 1, slot = kk & 1, p0 = kk * kCfT;
        cf_sync(cf_full(2, slot));
#pragma unroll
        for (int u = 0; u < kCfT; u++) {
          const int p = p0 + u;
          if (p < n_total) {
            const double xd = (double)S.mpx[slot][u][lane];
            const double ps = S.ps[slot][u][lane], pc = S.pc[slot][u][lane];
            const double tone = shift ? (2 * pc * pc - 1) : (2 * ps * pc);
            double ster = (tone * xd) * 2.0;
            if (de_st) {
              const double x0 = ster - P.de_a1 * des;
              ster = P.de_b0 * x0;
              des = x0;
            }
            const double m0 = xd - P.de_a1 * dem;
            const double mono = P.de_b0 * m0;
            dem = m0;
            if (act) orow[(t0lo + (uint32_t)p) & omask] = make_double2(mono, ster);
          }
        }
        if (kk + 2 < K) cf_arrive(cf_empty(2, slot));
      }
    }
    if (act) {
      sp->disc_prev = prev;
      sp->baseband_mean = bmean;
      sp->baseband_level = blevel;
      sp->de_m_x1 = dem;
      sp->de_s_x1 = des;
    }
  } else {
    // =============================== PilotPhaseLock::process (PilotPhaseLock.cpp:56-171)
    FmChanState s = st[c];
    if (reset_pps) s.n_pps = 0;
    const double kTwoPi = 2.0 * 3.14159265358979323846;
    const double f0 = (19000.0 / 384000.0) * kTwoPi;
    double sf0, cf0;
    sincos(f0, &sf0, &cf0);
    double minf = P.pll_minfreq, maxf = P.pll_maxfreq, lf_b0 = P.lf_b0, lf_b1 = P.lf_b1, bq_b0 = P.bq_b0;
    double bq_a1 = P.bq_a1, bq_a2 = P.bq_a2;
    double dlmin = minf - f0, dlmax = maxf - f0;
    asm volatile("" : "+d"(minf), "+d"(maxf), "+d"(lf_b0), "+d"(lf_b1), "+d"(bq_b0), "+d"(dlmin), "+d"(dlmax),
                 "+d"(bq_a1), "+d"(bq_a2));
    double bi1 = s.bi_x1, bi2 = s.bi_x2, bq1 = s.bq_x1, bq2 = s.bq_x2, lf1 = s.lf_x1;
    double freq = s.pll_freq, phase = s.pll_phase, ferr = s.freq_err;
    int periods = s.pilot_periods;
    double psin = 0.0, pcos = 1.0, last_i = 0.0, last_q = 0.0;
    bool was_locked = false;
    CfCalls cl;
    cl.init(call_end, n_calls);
    int b_flag = 0; // next call whose flag has not been written
    cf_arrive(cf_empty(1, 0));
    cf_arrive(cf_empty(1, 1));
    for (int k = 0; k < K; k++) {
      const int slot = k & 1, p0 = k * kCfT;
      const int valid = (n_total - p0 < kCfT) ? (n_total - p0) : kCfT;
      cf_sync(cf_full(1, slot));
      cf_sync(cf_empty(2, slot));
      float x_next = S.mpx[slot][0][lane];
#pragma unroll 1
      for (int u = 0; u < valid; u++) {
        const int p = p0 + u;
        const double xd = (double)x_next;
        if (u + 1 < valid) x_next = S.mpx[slot][u + 1][lane];
        if ((uint32_t)p == cl.end) {
          // ---- a reference call begins: flags of skipped empty calls, lock state, exact re-anchor
          cl.next();
          if (act) {
            for (; b_flag < cl.b; b_flag++) flags[(size_t)c * n_calls + b_flag] = (uint8_t)s.stereo_detected;
          }
          s.decoder_calls++;
          was_locked = (s.lock_cnt >= P.lock_delay);
          sincos(phase, &psin, &pcos);
        }
        S.ps[slot][u][lane] = psin;
        S.pc[slot][u][lane] = pcos;
        // ---- off the critical path
        const double fb_i = bq_a1 * bi1 + bq_a2 * bi2;
        const double fb_q = bq_a1 * bq1 + bq_a2 * bq2;
        const double fb_l = __dmul_rn(lf_b1, lf1);
        const double as = psin * cf0 + pcos * sf0;
        const double ac = pcos * cf0 - psin * sf0;
        // ---- critical path (see k_fm_pll2)
        const double i0v = psin * xd - fb_i;
        const double q0v = pcos * xd - fb_q;
        const double new_i = bq_b0 * i0v;
        const double new_q = bq_b0 * q0v;
        bi2 = bi1;
        bi1 = i0v;
        bq2 = bq1;
        bq1 = q0v;
        const double perr = (double)fast_atan2f_bf((float)new_q, (float)new_i, S.tab);
        last_i = new_i;
        last_q = new_q;
        ferr = fma(lf_b0, perr, fb_l);
        lf1 = perr;
        const double fraw = freq + ferr;
        const bool below_max = fraw < maxf, above_min = minf < fraw;
        const double dlr = fraw - f0;
        freq = below_max ? (above_min ? fraw : minf) : maxf;
        const double dl = below_max ? (above_min ? dlr : dlmin) : dlmax;
        {
          const double d2 = dl * dl;
          const double sd = fma(dl * d2, -1.0 / 6.0, dl);
          psin = fma(ac, sd, fma(-0.5 * as, d2, as));
          pcos = fma(-as, sd, fma(-0.5 * ac, d2, ac));
        }
        phase += freq;
        const bool wrap = phase > kTwoPi;
        phase = wrap ? phase - kTwoPi : phase;
        periods += wrap ? 1 : 0;
        if (wrap && periods == 19000) {
          periods = 0;
          if (was_locked) {
            if (s.n_pps < (uint32_t)kMaxPps && act) {
              PpsEventDev ev;
              ev.pps_index = s.pps_cnt;
              ev.sample_index = s.sample_cnt + (unsigned long long)((uint32_t)p - cl.beg);
              ev.block_position = (double)((uint32_t)p - cl.beg) / (double)(cl.end - cl.beg);
              ev.block = (uint32_t)(cl.b + block_off);
              ev.pad = 0;
              pps[(size_t)c * kMaxPps + s.n_pps] = ev;
            }
            s.n_pps++;
            s.pps_cnt++;
          }
        }
        if ((uint32_t)(p + 1) == cl.end) {
          // ---- the reference call ends (PilotPhaseLock.cpp:106,153-170)
          const int n = (int)(cl.end - cl.beg);
          s.pilot_level = sqrt(last_i * last_i + last_q * last_q);
          if (2 * s.pilot_level > P.minsignal) {
            if (s.lock_cnt < P.lock_delay) s.lock_cnt += n;
          } else {
            s.lock_cnt = 0;
          }
          if (s.lock_cnt < P.lock_delay) {
            periods = 0;
            s.pps_cnt = 0;
            while (s.n_pps > 0 && s.n_pps <= (uint32_t)kMaxPps &&
                   pps[(size_t)c * kMaxPps + s.n_pps - 1].block == (uint32_t)(cl.b + block_off)) {
              s.n_pps--;
            }
          }
          s.sample_cnt += (unsigned long long)n;
          s.stereo_detected = (s.lock_cnt >= P.lock_delay) ? 1 : 0;
          if (act) flags[(size_t)c * n_calls + cl.b] = (uint8_t)s.stereo_detected;
          b_flag = cl.b + 1;
        }
      }
      cf_arrive(cf_full(2, slot));
      if (k + 2 < K) cf_arrive(cf_empty(1, slot));
    }
    if (act) {
      for (; b_flag < n_calls; b_flag++) flags[(size_t)c * n_calls + b_flag] = (uint8_t)s.stereo_detected;
      FmChanState *o = st + c;
      o->stereo_detected = s.stereo_detected;
      o->lock_cnt = s.lock_cnt;
      o->pilot_periods = periods;
      o->n_pps = s.n_pps;
      o->pll_phase = phase;
      o->pll_freq = freq;
      o->bi_x1 = bi1;
      o->bi_x2 = bi2;
      o->bq_x1 = bq1;
      o->bq_x2 = bq2;
      o->lf_x1 = lf1;
      o->pilot_level = s.pilot_level;
      o->freq_err = ferr;
      o->pps_cnt = s.pps_cnt;
      o->sample_cnt = s.sample_cnt;
      o->decoder_calls = s.decoder_calls;
    }
  }
}

} // namespace fmr
#endif
