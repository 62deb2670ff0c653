// fmr_core.cuh — the 384 kHz core of FmDecoder::process (FmDecode.cpp:85-183, stereo, no
// multipath filter) as ONE warp-specialised kernel.
//
// The core is three strictly serial recurrences per channel (IfSimpleAgc, PilotPhaseLock, the
// deemphasis filters) with time-parallel work in between (PhaseDiscriminator, the L-R mix,
// statistics). As separate launches (k_fm_agc2 -> k_fm_disc/k_fm_call_stats -> k_fm_pll2) their
// latencies add up and every hand-over goes through HBM. Here a CTA owns 32 channels (lane =
// channel) and four warps run as a pipeline over chunks of kCfT samples:
//   warp 0  AGC    global IF ring -> gain recurrence -> shared memory (float2); IF RMS
//   warp 1  disc   atan2 / phase difference / baseband statistics -> shared memory (MPX)
//   warp 2  PLL    the pilot phase lock recurrence only (its dependent chain is the critical
//                  path of the whole kernel); hands sin/cos of the pilot phase on
//   warp 3  post   2*x*sin(2 phi), both deemphasis filters -> HBM
// Chunks are handed over through double-buffered shared memory with named barriers
// (bar.arrive / bar.sync between producer and consumer warps); nothing of the core except
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
constexpr int kCfThreads = 128; // four warps

struct CfSmem {
  float2 iq[2][kCfT][32];
  float mpx[2][kCfT][32];
  double ps[2][kCfT][32];
  double pc[2][kCfT][32];
  float2 tab[256];
  double kc[16]; // loop constants of the PLL, read back into registers (see warp 2)
};

// named barriers between producer and consumer warps (64 threads; 96 where a slot has two readers)
__device__ __forceinline__ void cf_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void cf_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void cf_sync3(int id) { asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void cf_arrive3(int id) { asm volatile("bar.arrive %0, 96;" ::"r"(id) : "memory"); }
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

template <typename AV>
static __global__ void __launch_bounds__(kCfThreads)
    k_fm_core_fused(Ring<float2> if_raw, Ring<float2> iq_in, Ring<AV> out384, FmChanState *__restrict__ st,
                    uint8_t *__restrict__ flags, PpsEventDev *__restrict__ pps, const uint32_t *__restrict__ call_end,
                    int n_calls, int64_t t0, FmCoreParams P, const float *__restrict__ atan_tbl, int block_off,
                    int reset_pps, int sm_count) {
  __shared__ CfSmem S;
  const int rot = (int)(blockIdx.x / (unsigned)sm_count); // CTAs b and b + #SMs tend to share an SM
  // Role of this warp. Hardware warp w of a CTA issues on SM sub-partition w % 4; CTAs that share
  // an SM rotate the roles so that their PLL warps (the critical recurrence) do not queue on the
  // same scheduler.
  const int lane = threadIdx.x & 31, warp = ((threadIdx.x >> 5) + rot) & 3;
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
    // =============================== PhaseDiscriminator::process (PhaseDiscriminator.cpp:33-46)
    // + samples_mean_rms and the EMA of FmDecode.cpp:146-150
    FmChanState *sp = st + c;
    float prev = sp->disc_prev;
    float bmean = sp->baseband_mean, blevel = sp->baseband_level;
    float vs = 0.f, vq = 0.f;
    const float inv_norm = P.disc_inv_norm, bound = P.disc_bound;
    CfCalls cl;
    cl.init(call_end, n_calls);
    cf_arrive(cf_empty(0, 0));
    cf_arrive(cf_empty(0, 1));
    for (int j = 0; j < K; j++) {
      const int slot = j & 1, p0 = j * kCfT;
      const int valid = (n_total - p0 < kCfT) ? (n_total - p0) : kCfT;
      cf_sync(cf_full(0, slot));
      // the eight arctangents of the chunk are independent: all in flight together
      float ph[kCfT];
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        const float2 x = S.iq[slot][u][lane];
        ph[u] = fmr_atan2f(x.y, x.x) * inv_norm;
      }
      if (j + 2 < K) cf_arrive(cf_empty(0, slot)); // the AGC warp may refill this slot now
      cf_sync3(cf_empty(1, slot));                 // PLL and post are done with this MPX slot
#pragma unroll
      for (int u = 0; u < kCfT; u++) {
        const int p = p0 + u;
        if (u < valid) {
          if ((uint32_t)p == cl.end) cl.next();
          float d = ph[u] - prev;
          prev = ph[u];
          if (d > bound) d -= 2 * bound;
          if (d < -bound) d += 2 * bound;
          if (isnan(d)) d = 0.0f;
          S.mpx[slot][u][lane] = d;
          vs += d;
          vq += d * d;
          if ((uint32_t)(p + 1) == cl.end) {
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
    if (act) {
      sp->disc_prev = prev;
      sp->baseband_mean = bmean;
      sp->baseband_level = blevel;
    }
  } else if (warp == 3) {
    // =============================== demod_stereo (FmDecode.cpp:224-239) + LowPassFilterRC x2
    FmChanState *sp = st + c;
    double dem = sp->de_m_x1, des = sp->de_s_x1;
    const bool shift = P.pilot_shift != 0, de_st = P.deemph_on_stereo != 0;
    const double de_a1 = P.de_a1, de_b0 = P.de_b0;
    AV *__restrict__ orow = out384.base + (size_t)c * out384.cap;
    const uint32_t omask = out384.cap - 1;
    cf_arrive(cf_empty(2, 0));
    cf_arrive(cf_empty(2, 1));
    cf_arrive3(cf_empty(1, 0));
    cf_arrive3(cf_empty(1, 1));
    for (int k = 0; k < K; k++) {
      const int slot = k & 1, p0 = k * kCfT;
      const int valid = (n_total - p0 < kCfT) ? (n_total - p0) : kCfT;
      cf_sync(cf_full(2, slot));
#pragma unroll 2
      for (int u = 0; u < valid; u++) {
        const double xd = (double)S.mpx[slot][u][lane];
        const double ps = S.ps[slot][u][lane], pc = S.pc[slot][u][lane];
        const double tone = shift ? (2 * pc * pc - 1) : (2 * ps * pc);
        double ster = (tone * xd) * 2.0;
        if (de_st) {
          const double x0 = ster - de_a1 * des;
          ster = de_b0 * x0;
          des = x0;
        }
        const double m0 = xd - de_a1 * dem;
        const double mono = de_b0 * m0;
        dem = m0;
        if (act) orow[(t0lo + (uint32_t)(p0 + u)) & omask] = aud_mk<AV>(mono, ster);
      }
      if (k + 2 < K) {
        cf_arrive(cf_empty(2, slot));
        cf_arrive3(cf_empty(1, slot));
      }
    }
    if (act) {
      sp->de_m_x1 = dem;
      sp->de_s_x1 = des;
    }
  } else {
    // =============================== PilotPhaseLock::process (PilotPhaseLock.cpp:56-171)
    FmChanState s = st[c];
    if (reset_pps) s.n_pps = 0;
    // Loop constants take a round trip through shared memory: kernel parameters and immediates
    // get re-materialised from the constant bank / uniform moves INSIDE the recurrence otherwise,
    // and every such instruction is a latency on the critical path of a one-warp kernel.
    {
      const double two_pi = 2.0 * 3.14159265358979323846;
      const double f0c = (19000.0 / 384000.0) * two_pi;
      double sf, cf;
      sincos(f0c, &sf, &cf);
      if (lane == 0) {
        S.kc[0] = two_pi;
        S.kc[1] = f0c;
        S.kc[2] = sf;
        S.kc[3] = cf;
        S.kc[4] = P.pll_minfreq;
        S.kc[5] = P.pll_maxfreq;
        S.kc[6] = P.pll_minfreq - f0c;
        S.kc[7] = P.pll_maxfreq - f0c;
        S.kc[8] = P.lf_b0;
        S.kc[9] = P.lf_b1;
        S.kc[10] = P.bq_b0;
        S.kc[11] = P.bq_a1;
        S.kc[12] = P.bq_a2;
        S.kc[13] = -1.0 / 6.0;
        S.kc[14] = -0.5;
      }
      __syncwarp();
    }
    const volatile double *kc = S.kc;
    const double kTwoPi = kc[0], f0 = kc[1], sf0 = kc[2], cf0 = kc[3], minf = kc[4], maxf = kc[5], dlmin = kc[6],
                 dlmax = kc[7], lf_b0 = kc[8], lf_b1 = kc[9], bq_b0 = kc[10], bq_a1 = kc[11], bq_a2 = kc[12],
                 k_m16 = kc[13], k_mh = kc[14];
    double bi1 = s.bi_x1, bi2 = s.bi_x2, bq1 = s.bq_x1, bq2 = s.bq_x2, lf1 = s.lf_x1;
    double freq = s.pll_freq, phase = s.pll_phase, ferr = s.freq_err;
    int periods = s.pilot_periods;
    double psin = 0.0, pcos = 1.0, last_i = 0.0, last_q = 0.0;
    bool was_locked = false;
    CfCalls cl;
    cl.init(call_end, n_calls);
    int b_flag = 0; // next call whose flag has not been written
    cf_arrive3(cf_empty(1, 0));
    cf_arrive3(cf_empty(1, 1));
    // 32-bit shared-window addresses of this lane's columns (explicit ld/st.shared below: a generic
    // pointer would be converted again for every access, inside the recurrence)
    const unsigned a_mpx = (unsigned)__cvta_generic_to_shared(&S.mpx[0][0][lane]);
    const unsigned a_ps = (unsigned)__cvta_generic_to_shared(&S.ps[0][0][lane]);
    const unsigned a_pc = (unsigned)__cvta_generic_to_shared(&S.pc[0][0][lane]);
    const unsigned a_tab = (unsigned)__cvta_generic_to_shared(&S.tab[0]);
    uint32_t p = 0; // flat sample index
    for (int k = 0; k < K; k++) {
      const int slot = k & 1;
      const uint32_t p_chunk_end = min((uint32_t)n_total, (uint32_t)(k + 1) * kCfT);
      cf_sync(cf_full(1, slot));
      cf_sync(cf_empty(2, slot));
      unsigned am = a_mpx + slot * (kCfT * 32 * 4), aps = a_ps + slot * (kCfT * 32 * 8), apc = a_pc + slot * (kCfT * 32 * 8);
      float x_next;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x_next) : "r"(am));
      while (p < p_chunk_end) {
        if (p == cl.end) {
          // ---- a reference call begins: flags of skipped empty calls, lock state, exact re-anchor
          cl.next();
          if (act) {
            for (; b_flag < cl.b; b_flag++) flags[(size_t)c * n_calls + b_flag] = (uint8_t)s.stereo_detected;
          }
          s.decoder_calls++;
          was_locked = (s.lock_cnt >= P.lock_delay);
          sincos(phase, &psin, &pcos);
        }
        // samples up to the end of the chunk or of the reference call, whichever comes first
        const uint32_t run_end = min(p_chunk_end, cl.end);
#pragma unroll 1
        for (; p < run_end; p++) {
          const double xd = (double)x_next;
          am += 32 * 4;
          if (p + 1 < p_chunk_end) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x_next) : "r"(am));
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(aps), "d"(psin) : "memory");
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(apc), "d"(pcos) : "memory");
          aps += 32 * 8;
          apc += 32 * 8;
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
          const double perr = (double)fast_atan2f_bf_s((float)new_q, (float)new_i, a_tab);
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
            const double sd = fma(dl * d2, k_m16, dl);
            psin = fma(ac, sd, fma(k_mh * as, d2, as));
            pcos = fma(-as, sd, fma(k_mh * ac, d2, ac));
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
                ev.sample_index = s.sample_cnt + (unsigned long long)(p - cl.beg);
                ev.block_position = (double)(p - cl.beg) / (double)(cl.end - cl.beg);
                ev.block = (uint32_t)(cl.b + block_off);
                ev.pad = 0;
                pps[(size_t)c * kMaxPps + s.n_pps] = ev;
              }
              s.n_pps++;
              s.pps_cnt++;
            }
          }
        }
        if (p == cl.end) {
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
      if (k + 2 < K) cf_arrive3(cf_empty(1, slot));
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
