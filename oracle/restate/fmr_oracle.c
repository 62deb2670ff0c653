/*
 * TEST INFRASTRUCTURE ONLY — never linked into or called by the product.
 *
 * Plain-C, scalar, block-by-block restatement of the reference's demodulation hot path
 * (SURVEY.md §8(a)), written from the reference's algorithm description, each function
 * citing the reference file:line it follows. It processes one source block per call, with
 * the reference's own per-call semantics, so that its per-call output sizes, statistics
 * and block-boundary quirks can be compared one to one.
 *
 * Pinning: tests/test_oracle.py checks this restatement against the compiled reference
 * (oracle/_ref/libfmref.so, built from /root/reference) when that library exists, and
 * against the committed golden vectors under tests/golden/ (generated from the compiled
 * reference by tools/gen_golden.py) always.
 *
 * Third-party arithmetic: the r8brain resampler filters (half-band taps, long low-pass
 * taps, polyphase banks) are data read back from the objects the reference constructs
 * (tools/gen_tables.py -> fmr_tables_generated.inc); the convolution structure and the
 * release schedule of each stage are restated here (CDSPHBDownsampler.h:137-239,
 * CDSPBlockConvolver.h:252-353, CDSPFracInterpolator.h:861-925,992-1060). VOLK kernels
 * follow the published generic protokernels (sequential accumulation).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define nullptr 0
typedef struct HbStage { int ntaps; const double *taps; } HbStage;
typedef struct BcStage { int klen, inputlen, latency, down, outoffset; const double *taps; } BcStage;
typedef struct FiStage { int instep, outstep, flen; const double *taps; } FiStage;
typedef struct ChainDesc { double src, dst; int kind; int n_hb; HbStage hb[3]; BcStage bc; int has_fi; FiStage fi; int verified; } ChainDesc;
#include "../../airspy_fmradion_b200/csrc/fmr_tables_generated.inc"

static const ChainDesc *find_chain(double src, double dst, int kind) {
  for (int i = 0; i < kNumChains; i++)
    if (kChains[i].src == src && kChains[i].dst == dst && kChains[i].kind == kind) return &kChains[i];
  return 0;
}

/* ------------------------------------------------------------------ growable stream -- */
/* Samples of one real stream with absolute indexing; old samples are dropped lazily. */
typedef struct { double *v; int64_t base, n, cap; } Stream;
static void st_init(Stream *s) { s->v = 0; s->base = 0; s->n = 0; s->cap = 0; }
static void st_free(Stream *s) { free(s->v); }
static double st_get(const Stream *s, int64_t i) { return (i < 0) ? 0.0 : s->v[i - s->base]; }
static void st_push(Stream *s, double x) {
  if (s->n - s->base == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 4096;
    s->v = (double *)realloc(s->v, sizeof(double) * (size_t)s->cap);
  }
  s->v[s->n - s->base] = x;
  s->n++;
}
static void st_drop_before(Stream *s, int64_t keep_from) {
  if (keep_from - s->base > 65536) {
    memmove(s->v, s->v + (keep_from - s->base), sizeof(double) * (size_t)(s->n - keep_from));
    s->base = keep_from;
  }
}

/* ------------------------------------------------- one real lane of r8b::CDSPResampler -- */
typedef struct {
  const ChainDesc *d;
  Stream in[6]; /* in[0] raw input, then the input stream of every following stage */
  int64_t emitted; /* outputs released so far by the last stage */
} R8Lane;

static void r8_init(R8Lane *r, const ChainDesc *d) {
  r->d = d;
  r->emitted = 0;
  for (int i = 0; i < 6; i++) st_init(&r->in[i]);
}
static void r8_free(R8Lane *r) { for (int i = 0; i < 6; i++) st_free(&r->in[i]); }

/* Feed n samples; append the released outputs to `out` (caller's Stream), return count. */
static int r8_process(R8Lane *r, const double *x, int n, Stream *out) {
  const ChainDesc *d = r->d;
  for (int i = 0; i < n; i++) st_push(&r->in[0], x[i]);
  int si = 0;
  /* half-band stages: CDSPHBDownsampler::process (CDSPHBDownsampler.h:166-238),
     y[m] = x[2m] + sum_k t[k](x[2m+2k+1] + x[2m-2k-1]) (CDSPHBDownsampler.inc:620-629);
     output m is released once floor(N/2) - (taps-1) > m. */
  for (int s = 0; s < d->n_hb; s++, si++) {
    Stream *a = &r->in[si], *b = &r->in[si + 1];
    const int nt = d->hb[s].ntaps;
    int64_t avail = a->n / 2 - (nt - 1);
    for (int64_t m = b->n; m < avail; m++) {
      double y = st_get(a, 2 * m);
      for (int k = 0; k < nt; k++) y += d->hb[s].taps[k] * (st_get(a, 2 * m + 2 * k + 1) + st_get(a, 2 * m - 2 * k - 1));
      st_push(b, y);
    }
    st_drop_before(a, 2 * b->n - 2 * nt - 2);
  }
  /* block convolver: CDSPBlockConvolver::process (CDSPBlockConvolver.h:252-353): zero-phase
     linear convolution; sample t is released once N - Latency > t; with DownFactor 2 only
     even t are emitted (copyToOutput, :486-560). */
  {
    Stream *a = &r->in[si], *b = (d->has_fi ? &r->in[si + 1] : out);
    const int fl2 = (d->bc.klen - 1) / 2;
    int64_t c = a->n - d->bc.latency;
    if (c < 0) c = 0;
    int64_t avail = (d->bc.down > 1) ? (c + d->bc.down - 1) / d->bc.down : c;
    int64_t q0 = d->has_fi ? b->n : r->emitted;
    int produced = 0;
    for (int64_t q = q0; q < avail; q++) {
      const int64_t t = q * d->bc.down;
      double y = 0.0;
      for (int j = 0; j < d->bc.klen; j++) y += d->bc.taps[j] * st_get(a, t - fl2 + j);
      st_push(b, y);
      produced++;
    }
    if (!d->has_fi) {
      r->emitted += produced;
      st_drop_before(a, (q0 + produced) * d->bc.down - fl2 - 2);
      return produced;
    }
    st_drop_before(a, b->n * d->bc.down - fl2 - 2);
    si++;
  }
  /* whole-step polyphase interpolator: CDSPFracInterpolator::convolve0
     (CDSPFracInterpolator.h:992-1060); released while BufLeft - fl2 > 0 (:1010). */
  {
    Stream *a = &r->in[si];
    const FiStage *f = &d->fi;
    const int fl2 = f->flen / 2, fll = fl2 - 1;
    int produced = 0;
    int64_t m = r->emitted;
    for (;; m++) {
      const int64_t pos = m * f->instep;
      const int64_t ip = pos / f->outstep;
      const int ph = (int)(pos - ip * f->outstep);
      if (ip + fl2 + 1 > a->n) break;
      const double *row = f->taps + (size_t)ph * f->flen;
      double y = 0.0;
      for (int k = 0; k < f->flen; k++) y += row[k] * st_get(a, ip - fll + k);
      st_push(out, y);
      produced++;
    }
    r->emitted = m;
    st_drop_before(a, (m * f->instep) / f->outstep - f->flen);
    return produced;
  }
}

/* ------------------------------------------------------------------ small filters ---- */
/* FirstOrderIirFilter::process (Filter.cpp:172-178) */
typedef struct { double b0, b1, a1, x1; } FoIir;
static double fo_process(FoIir *f, double in) {
  double x0 = in - f->a1 * f->x1;
  double y = f->b0 * x0 + f->b1 * f->x1;
  f->x1 = x0;
  return y;
}
/* BiquadIirFilter::process (Filter.cpp:243-250) */
typedef struct { double b0, b1, b2, a1, a2, x1, x2; } Biquad;
static double bq_process(Biquad *f, double in) {
  double x0 = in - (f->a1 * f->x1 + f->a2 * f->x2);
  double y = f->b0 * x0 + f->b1 * f->x1 + f->b2 * f->x2;
  f->x2 = f->x1;
  f->x1 = x0;
  return y;
}
/* HighPassFilterIir::HighPassFilterIir (Filter.cpp:254-290): 2-pole matched-Z high-pass. */
static void hp_init(Biquad *f, double cutoff) {
  const double w = 2 * M_PI * cutoff;
  /* p1s = w / exp(j*3pi/4); p1z = exp(p1s) */
  const double ang = 3.0 / 4.0 * M_PI;
  const double re_s = w * cos(ang), im_s = -w * sin(ang);
  const double mag = exp(re_s);
  const double re_z = mag * cos(im_s), im_z = mag * sin(im_s);
  double b0 = 1, b1 = -2, b2 = 1;
  f->a1 = -2 * re_z;
  f->a2 = sqrt((re_z * re_z - im_z * im_z) * (re_z * re_z - im_z * im_z) + (2 * re_z * im_z) * (2 * re_z * im_z));
  const double g = (b0 - b1 + b2) / (1 - f->a1 + f->a2);
  f->b0 = b0 / g;
  f->b1 = b1 / g;
  f->b2 = b2 / g;
  f->x1 = f->x2 = 0;
}
/* LowPassFilterRC (Filter.cpp:186-188) */
static void rc_init(FoIir *f, double timeconst) {
  f->a1 = -exp(-1 / timeconst);
  f->b0 = 1 + f->a1;
  f->b1 = 0;
  f->x1 = 0;
}

/* LowPassFilterFirAudio::process (Filter.cpp:108-163) — including the head loop that starts
   at coefficient 1 for the first min(n, order) outputs of every call. */
typedef struct { const double *coeff; int order; double *state; } FirAudio;
static void fira_init(FirAudio *f, const double *coeff, int ntaps) {
  f->coeff = coeff;
  f->order = ntaps - 1;
  f->state = (double *)calloc((size_t)f->order, sizeof(double));
}
static void fira_process(FirAudio *f, const double *in, int n, double *out) {
  const int order = f->order;
  int p = 0;
  for (; p < n && p < order; p++) {
    double y = 0;
    for (int j = p + 1; j <= order; j++) y += f->state[order + p - j] * f->coeff[j];
    for (int j = 1; j <= p; j++) y += in[p - j] * f->coeff[j];
    out[p] = y;
  }
  const int half = (order - 1) / 2;
  for (; p < n; p++) {
    double y = 0;
    for (int k = 0; k <= half; k++) y += (in[p - k] + in[p - (order - k)]) * f->coeff[k];
    if ((order % 2) == 0) y += in[p - order / 2] * f->coeff[order / 2];
    out[p] = y;
  }
  if (n < order) {
    memmove(f->state, f->state + n, sizeof(double) * (size_t)(order - n));
    memcpy(f->state + (order - n), in, sizeof(double) * (size_t)n);
  } else {
    memcpy(f->state, in + (n - order), sizeof(double) * (size_t)order);
  }
}
/* LowPassFilterFirIQ::process (Filter.cpp:37-96), downsample 1, float accumulation. */
typedef struct { const float *coeff; int order; float *sre, *sim; } FirIQ;
static void firiq_init(FirIQ *f, const float *coeff, int ntaps) {
  f->coeff = coeff;
  f->order = ntaps - 1;
  f->sre = (float *)calloc((size_t)f->order, sizeof(float));
  f->sim = (float *)calloc((size_t)f->order, sizeof(float));
}
static void firiq_process(FirIQ *f, const float *in /*re,im*/, int n, float *out) {
  const int order = f->order;
  int p = 0;
  for (; p < n && p < order; p++) {
    float yr = 0, yi = 0;
    for (int j = p + 1; j <= order; j++) {
      yr += f->sre[order + p - j] * f->coeff[j];
      yi += f->sim[order + p - j] * f->coeff[j];
    }
    for (int j = 1; j <= p; j++) {
      yr += in[2 * (p - j)] * f->coeff[j];
      yi += in[2 * (p - j) + 1] * f->coeff[j];
    }
    out[2 * p] = yr;
    out[2 * p + 1] = yi;
  }
  const int half = (order - 1) / 2;
  for (; p < n; p++) {
    float yr = 0, yi = 0;
    for (int k = 0; k <= half; k++) {
      yr += (in[2 * (p - k)] + in[2 * (p - (order - k))]) * f->coeff[k];
      yi += (in[2 * (p - k) + 1] + in[2 * (p - (order - k)) + 1]) * f->coeff[k];
    }
    if ((order % 2) == 0) {
      yr += in[2 * (p - order / 2)] * f->coeff[order / 2];
      yi += in[2 * (p - order / 2) + 1] * f->coeff[order / 2];
    }
    out[2 * p] = yr;
    out[2 * p + 1] = yi;
  }
  for (int i = 0; i < order; i++) {
    /* state keeps the last `order` input samples */
    int src = n - order + i;
    if (src >= 0) {
      f->sre[i] = in[2 * src];
      f->sim[i] = in[2 * src + 1];
    } else {
      f->sre[i] = f->sre[i + n];
      f->sim[i] = f->sim[i + n];
    }
  }
}

/* Utility::fast_atan2f (Utility.h:236-304) */
static float fast_atan2f_c(float y, float x) {
  float x_abs, y_abs, z, alpha, angle, base_angle;
  int index;
  y_abs = fabsf(y);
  x_abs = fabsf(x);
  if (!((y_abs > 0.0f) || (x_abs > 0.0f))) return 0.0f;
  z = (y_abs < x_abs) ? (y_abs / x_abs) : (x_abs / y_abs);
  if (z < 0.003921569) {
    base_angle = z;
  } else {
    alpha = z * (float)255;
    index = ((int)alpha) & 0xff;
    alpha -= (float)index;
    base_angle = k_fast_atan_table[index];
    base_angle += (k_fast_atan_table[index + 1] - k_fast_atan_table[index]) * alpha;
  }
  if (x_abs > y_abs) {
    if (x >= 0.0) {
      angle = (y >= 0.0) ? base_angle : -base_angle;
    } else {
      angle = 3.14159265358979323846;
      if (y >= 0.0) angle -= base_angle; else angle = base_angle - angle;
    }
  } else {
    if (y >= 0.0) {
      angle = 1.57079632679489661923;
      if (x >= 0.0) angle -= base_angle; else angle += base_angle;
    } else {
      angle = -1.57079632679489661923;
      if (x >= 0.0) angle += base_angle; else angle -= base_angle;
    }
  }
  return angle;
}

/* ------------------------------------------------------------------ FM chain ---------- */
typedef struct {
  double ifrate;
  int fs4, fmfilter, stereo, pilot_shift;
  unsigned mpf_stages;
  const ChainDesc *ifc;
  R8Lane if_re, if_im, au_m, au_s;
  unsigned fs4_idx; /* FourthConverterIQ m_index */
  FirIQ fmf;
  /* IfSimpleAgc */
  float agc_gain, agc_max, agc_rate;
  /* MultipathFilter */
  int mpf_n, mpf_ref;
  float *mpf_cre, *mpf_cim, *mpf_sre, *mpf_sim;
  unsigned mpf_wait;
  double mpf_error;
  /* PhaseDiscriminator */
  float disc_norm, disc_bound, disc_save;
  /* statistics */
  float baseband_mean, baseband_level, if_rms;
  int stereo_detected;
  /* PilotPhaseLock */
  double minfreq, maxfreq, freq, phase, pilot_level, freq_err;
  int lock_delay, lock_cnt, pilot_periods;
  uint64_t pps_cnt, sample_cnt;
  Biquad bq_i, bq_q;
  FoIir loopf;
  int n_pps;
  double pps[16 * 3];
  FoIir de_m, de_s;
  FirAudio pc_m, pc_s;
  Biquad dc_m, dc_s;
  uint64_t decoder_calls;
  /* taps of the last call */
  float *tap_if; int tap_if_n;
} OrcFm;

static void mpf_init_coeff(OrcFm *h) {
  for (int i = 0; i < h->mpf_n; i++) { h->mpf_cre[i] = 0; h->mpf_cim[i] = 0; }
  h->mpf_cre[h->mpf_ref] = 1;
}

void *orc_fm_create(double ifrate, int fs4, int filter, int stereo, double deemph_us, int pilot_shift, unsigned mpf_stages) {
  OrcFm *h = (OrcFm *)calloc(1, sizeof(OrcFm));
  h->ifrate = ifrate;
  h->fs4 = fs4;
  h->fmfilter = filter;
  h->stereo = stereo;
  h->pilot_shift = pilot_shift;
  h->mpf_stages = mpf_stages;
  if (ifrate != 384000.0) {
    h->ifc = find_chain(ifrate, 384000.0, 0);
    if (!h->ifc) { free(h); return 0; }
    r8_init(&h->if_re, h->ifc);
    r8_init(&h->if_im, h->ifc);
  }
  const ChainDesc *au = find_chain(384000.0, 48000.0, 1);
  r8_init(&h->au_m, au);
  r8_init(&h->au_s, au);
  firiq_init(&h->fmf, filter == 1 ? k_jj1bdx_fm_384kHz_medium : (filter == 2 ? k_jj1bdx_fm_384kHz_narrow : k_delay_3taps_only_iq),
             filter ? 127 : 3);
  h->agc_gain = 1.0f; h->agc_max = 100000.0f; h->agc_rate = 0.0001f;          /* FmDecode.cpp:74 */
  unsigned st = mpf_stages > 0 ? mpf_stages : 1;                               /* FmDecode.cpp:79 */
  h->mpf_n = (int)st * 4 + 1; h->mpf_ref = (int)st * 3 + 1;                    /* MultipathFilter.cpp:41-49 */
  h->mpf_cre = (float *)calloc((size_t)h->mpf_n, 4); h->mpf_cim = (float *)calloc((size_t)h->mpf_n, 4);
  h->mpf_sre = (float *)calloc((size_t)h->mpf_n, 4); h->mpf_sim = (float *)calloc((size_t)h->mpf_n, 4);
  mpf_init_coeff(h);
  h->mpf_wait = 100;                                                           /* FmDecode.cpp:33 */
  const double mfd = 75000.0 / 384000.0;
  h->disc_norm = (float)(mfd * 2.0 * M_PI); h->disc_bound = (float)(1.0 / (mfd * 2.0)); /* PhaseDiscriminator.cpp:27-30 */
  const double freq = 19000.0 / 384000.0, bw = 30.0 / 384000.0;                /* PilotPhaseLock.cpp:35-51 */
  h->minfreq = (freq - bw) * 2.0 * M_PI; h->maxfreq = (freq + bw) * 2.0 * M_PI;
  h->freq = freq * 2.0 * M_PI; h->lock_delay = (int)(15.0 / bw);
  h->bq_i.b0 = h->bq_q.b0 = 1.46974784e-06; h->bq_i.a1 = h->bq_q.a1 = -1.99682419; h->bq_i.a2 = h->bq_q.a2 = 0.996825659;
  h->loopf.b0 = 0.000304341788; h->loopf.b1 = -0.000304324564;
  const double tc = (deemph_us == 0) ? 1.0 : (deemph_us * 384000.0 * 1.0e-6);   /* FmDecode.cpp:67-70 */
  rc_init(&h->de_m, tc); rc_init(&h->de_s, tc);
  fira_init(&h->pc_m, k_jj1bdx_48khz_fmaudio, 127); fira_init(&h->pc_s, k_jj1bdx_48khz_fmaudio, 127);
  hp_init(&h->dc_m, 0.0001); hp_init(&h->dc_s, 0.0001);                          /* FmDecode.cpp:62 */
  return h;
}

void orc_fm_destroy(void *p) {
  OrcFm *h = (OrcFm *)p;
  if (!h) return;
  if (h->ifc) { r8_free(&h->if_re); r8_free(&h->if_im); }
  r8_free(&h->au_m); r8_free(&h->au_s);
  free(h->mpf_cre); free(h->mpf_cim); free(h->mpf_sre); free(h->mpf_sim);
  free(h->fmf.sre); free(h->fmf.sim); free(h->pc_m.state); free(h->pc_s.state); free(h->tap_if);
  free(h);
}

/* Front end shared by FM and AM: FourthConverterIQ::process (FourthConverterIQ.h:38-82) and
   IfResampler::process (IfResampler.cpp:37-79). Returns the number of complex outputs. */
static int front_end(int fs4, unsigned *fs4_idx, const ChainDesc *ifc, R8Lane *lre, R8Lane *lim, const float *iq, int n,
                     float **out) {
  double *re = (double *)malloc(sizeof(double) * (size_t)(n + 1)), *im = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  for (int i = 0; i < n; i++) {
    float a = iq[2 * i], b = iq[2 * i + 1], yr = a, yi = b;
    if (fs4) {
      switch (*fs4_idx) { /* downconvert: +1, -j... table order 0->1->2->3 */
      case 0: yr = a; yi = b; break;
      case 1: yr = b; yi = -a; break;
      case 2: yr = -a; yi = -b; break;
      default: yr = -b; yi = a; break;
      }
      *fs4_idx = (*fs4_idx + 1) & 3;
    }
    re[i] = (double)yr;
    im[i] = (double)yi;
  }
  int m;
  if (ifc) {
    Stream ore, oim;
    st_init(&ore); st_init(&oim);
    m = r8_process(lre, re, n, &ore);
    int m2 = r8_process(lim, im, n, &oim);
    (void)m2;
    *out = (float *)malloc(sizeof(float) * 2 * (size_t)(m + 1));
    for (int i = 0; i < m; i++) { (*out)[2 * i] = (float)ore.v[i]; (*out)[2 * i + 1] = (float)oim.v[i]; }
    st_free(&ore); st_free(&oim);
  } else {
    m = n;
    *out = (float *)malloc(sizeof(float) * 2 * (size_t)(m + 1));
    for (int i = 0; i < m; i++) { (*out)[2 * i] = (float)re[i]; (*out)[2 * i + 1] = (float)im[i]; }
  }
  free(re); free(im);
  return m;
}

/* IfSimpleAgc::process (IfSimpleAgc.cpp:37-57), in place */
static void if_agc(float *gain, float max_gain, float rate, float *x, int n) {
  for (int i = 0; i < n; i++) {
    float xr = x[2 * i] * *gain, xi = x[2 * i + 1] * *gain;
    x[2 * i] = xr; x[2 * i + 1] = xi;
    float nrm = xr * xr + xi * xi;
    float z = (float)(1.0 + ((double)rate * (1.0 - (double)nrm)));
    *gain *= z;
    if (!isfinite(*gain)) *gain = 1.0f; else if (*gain > max_gain) *gain = max_gain;
  }
}

/* MultipathFilter::process (MultipathFilter.cpp:164-197) with single_process (:92-105) and
   update_coeff (:108-161). Returns 0 on failure. out may not alias in. */
static int mpf_process(OrcFm *h, const float *in, int n, float *out) {
  const int N = h->mpf_n;
  for (int i = 0; i < n; i++) {
    memmove(h->mpf_sre, h->mpf_sre + 1, sizeof(float) * (size_t)(N - 1));
    memmove(h->mpf_sim, h->mpf_sim + 1, sizeof(float) * (size_t)(N - 1));
    h->mpf_sre[N - 1] = in[2 * i]; h->mpf_sim[N - 1] = in[2 * i + 1];
    float yr = 0, yi = 0;
    for (int k = 0; k < N; k++) {
      yr += h->mpf_sre[k] * h->mpf_cre[k] - h->mpf_sim[k] * h->mpf_cim[k];
      yi += h->mpf_sre[k] * h->mpf_cim[k] + h->mpf_sim[k] * h->mpf_cre[k];
    }
    if (!isfinite(yr) || !isfinite(yi)) return 0;
    out[2 * i] = yr; out[2 * i + 1] = yi;
    if ((i & 3) == 0) {
      const double env = (double)(yr * yr + yi * yi);
      const double error = 1.0 - env;
      float ms = 0;
      for (int k = 0; k < N; k++) ms += h->mpf_sre[k] * h->mpf_sre[k] + h->mpf_sim[k] * h->mpf_sim[k];
      const float mu = (float)(0.1 / ((double)ms + 1e-10));
      const float factor = (float)(error * (double)mu);
      const float fr = factor * yr, fi = factor * yi;
      for (int k = 0; k < N; k++) {
        h->mpf_cre[k] += fr * h->mpf_sre[k] + fi * h->mpf_sim[k];
        h->mpf_cim[k] += fi * h->mpf_sre[k] - fr * h->mpf_sim[k];
      }
      h->mpf_cre[h->mpf_ref] = 1; h->mpf_cim[h->mpf_ref] = 0;
      h->mpf_error = error;
      if (!isfinite(error)) return 0;
    }
  }
  return 1;
}

/* PilotPhaseLock::process (PilotPhaseLock.cpp:56-171) */
static void pll_process(OrcFm *h, const double *in, int n, double *out) {
  const int was_locked = (h->lock_cnt >= h->lock_delay);
  h->n_pps = 0;
  if (n == 0) return;
  h->pilot_level = 1000.0;
  for (int i = 0; i < n; i++) {
    const double psin = sin(h->phase), pcos = cos(h->phase);
    out[i] = h->pilot_shift ? (2 * pcos * pcos - 1) : (2 * psin * pcos);
    const double x = in[i];
    const double ni = bq_process(&h->bq_i, psin * x), nq = bq_process(&h->bq_q, pcos * x);
    const double perr = (double)fast_atan2f_c((float)nq, (float)ni);
    h->pilot_level = sqrt(ni * ni + nq * nq);
    h->freq_err = fo_process(&h->loopf, perr);
    h->freq += h->freq_err;
    h->freq = fmax(h->minfreq, fmin(h->maxfreq, h->freq));
    h->phase += h->freq;
    if (h->phase > 2.0 * M_PI) {
      h->phase -= 2.0 * M_PI;
      h->pilot_periods++;
      if (h->pilot_periods == 19000) {
        h->pilot_periods = 0;
        if (was_locked) {
          if (h->n_pps < 16) {
            h->pps[3 * h->n_pps] = (double)h->pps_cnt;
            h->pps[3 * h->n_pps + 1] = (double)(h->sample_cnt + (uint64_t)i);
            h->pps[3 * h->n_pps + 2] = (double)i / (double)n;
          }
          h->n_pps++;
          h->pps_cnt++;
        }
      }
    }
  }
  if (2 * h->pilot_level > 0.001) { if (h->lock_cnt < h->lock_delay) h->lock_cnt += n; } else h->lock_cnt = 0;
  if (h->lock_cnt < h->lock_delay) { h->pilot_periods = 0; h->pps_cnt = 0; h->n_pps = 0; }
  h->sample_cnt += (uint64_t)n;
}

/* One source block through main.cpp:912-956 and FmDecoder::process (FmDecode.cpp:85-221).
   Returns the number of audio doubles written (interleaved L,R when stereo). */
int orc_fm_process_block(void *p, const float *iq, int n, double *audio, int cap) {
  OrcFm *h = (OrcFm *)p;
  float *ifs = 0;
  const int m = front_end(h->fs4, &h->fs4_idx, h->ifc, &h->if_re, &h->if_im, iq, n, &ifs);
  free(h->tap_if);
  h->tap_if = (float *)malloc(sizeof(float) * 2 * (size_t)(m + 1));
  memcpy(h->tap_if, ifs, sizeof(float) * 2 * (size_t)m);
  h->tap_if_n = m;
  if (m == 0) { free(ifs); return 0; }       /* main.cpp:933-936 */
  h->decoder_calls++;
  /* Utility::rms_level_sample (Utility.h:118-132) */
  {
    float level = 0;
    for (int i = 0; i < m; i++) level += ifs[2 * i] * ifs[2 * i] + ifs[2 * i + 1] * ifs[2 * i + 1];
    h->if_rms = sqrtf(level / (float)m);
  }
  float *x = (float *)malloc(sizeof(float) * 2 * (size_t)m);
  if (h->fmfilter) firiq_process(&h->fmf, ifs, m, x); else memcpy(x, ifs, sizeof(float) * 2 * (size_t)m);
  if_agc(&h->agc_gain, h->agc_max, h->agc_rate, x, m);
  if (h->mpf_wait > 0) {
    h->mpf_wait--;
  } else if (h->mpf_stages > 0) {
    float *y = (float *)malloc(sizeof(float) * 2 * (size_t)m);
    if (mpf_process(h, x, m, y)) { memcpy(x, y, sizeof(float) * 2 * (size_t)m); } else { mpf_init_coeff(h); }
    free(y);
  }
  /* PhaseDiscriminator::process (PhaseDiscriminator.cpp:33-46): generic atan2 / fm_detect */
  float *dec = (float *)malloc(sizeof(float) * (size_t)m);
  {
    const float inv = 1.0f / h->disc_norm;
    float prev = h->disc_save, last = 0;
    for (int i = 0; i < m; i++) {
      const float ph = atan2f(x[2 * i + 1], x[2 * i]) * inv;
      float d = ph - prev;
      if (d > h->disc_bound) d -= 2 * h->disc_bound;
      if (d < -h->disc_bound) d += 2 * h->disc_bound;
      prev = ph; last = ph;
      dec[i] = isnan(d) ? 0.0f : d;
    }
    h->disc_save = last;
  }
  double *bb = (double *)malloc(sizeof(double) * (size_t)m), *rs = (double *)malloc(sizeof(double) * (size_t)m);
  float vsum = 0, vsq = 0;
  for (int i = 0; i < m; i++) { bb[i] = (double)dec[i]; vsum += dec[i]; vsq += dec[i] * dec[i]; }
  h->baseband_mean = (float)(0.95 * h->baseband_mean + 0.05 * (vsum / (float)m));
  h->baseband_level = (float)(0.95 * h->baseband_level + 0.05 * sqrtf(vsq / (float)m));
  Stream s48m, s48s;
  st_init(&s48m); st_init(&s48s);
  int ns = 0;
  if (h->stereo) {
    pll_process(h, bb, m, rs);
    h->stereo_detected = (h->lock_cnt >= h->lock_delay);
    for (int i = 0; i < m; i++) { rs[i] = rs[i] * bb[i]; rs[i] = rs[i] * 2.0; }   /* demod_stereo :224-239 */
    if (!h->pilot_shift) for (int i = 0; i < m; i++) rs[i] = fo_process(&h->de_s, rs[i]);
    ns = r8_process(&h->au_s, rs, m, &s48s);
  }
  for (int i = 0; i < m; i++) bb[i] = fo_process(&h->de_m, bb[i]);
  const int nm = r8_process(&h->au_m, bb, m, &s48m);
  int produced = 0;
  if (nm > 0) {
    double *mono = (double *)malloc(sizeof(double) * (size_t)nm), *ster = (double *)malloc(sizeof(double) * (size_t)(nm + 1));
    fira_process(&h->pc_m, s48m.v, nm, mono);
    for (int i = 0; i < nm; i++) mono[i] = bq_process(&h->dc_m, mono[i]);
    if (h->stereo) {
      (void)ns;
      fira_process(&h->pc_s, s48s.v, nm, ster);
      for (int i = 0; i < nm; i++) ster[i] = bq_process(&h->dc_s, ster[i]);
      if (2 * nm <= cap) {
        for (int i = 0; i < nm; i++) {
          double l, r;
          if (h->stereo_detected) {
            if (h->pilot_shift) { l = r = ster[i]; } else { const double sb = 1.017 * ster[i]; l = mono[i] + sb; r = mono[i] - sb; }
          } else {
            if (h->pilot_shift) { l = r = 0.0; } else { l = r = mono[i]; }
          }
          audio[2 * i] = l; audio[2 * i + 1] = r;
        }
        produced = 2 * nm;
      } else produced = -1;
    } else {
      if (nm <= cap) { memcpy(audio, mono, sizeof(double) * (size_t)nm); produced = nm; } else produced = -1;
    }
    free(mono); free(ster);
  }
  st_free(&s48m); st_free(&s48s);
  free(bb); free(rs); free(dec); free(x); free(ifs);
  return produced;
}

int orc_fm_tap_if(void *p, float *out, int cap) {
  OrcFm *h = (OrcFm *)p;
  int n = h->tap_if_n < cap ? h->tap_if_n : cap;
  memcpy(out, h->tap_if, sizeof(float) * 2 * (size_t)n);
  return h->tap_if_n;
}

typedef struct {
  int stereo_detected; float tuning_offset, baseband_level; double pilot_level; float if_rms; double mpf_error;
  float agc_gain; double pll_freq, pll_phase; int pll_lock_cnt; uint64_t decoder_calls; int n_pps;
} OrcFmStats;
void orc_fm_stats(void *p, OrcFmStats *s) {
  OrcFm *h = (OrcFm *)p;
  s->stereo_detected = h->stereo_detected; s->tuning_offset = h->baseband_mean * 75000.0f;
  s->baseband_level = h->baseband_level; s->pilot_level = 2 * h->pilot_level; s->if_rms = h->if_rms;
  s->mpf_error = h->mpf_error; s->agc_gain = h->agc_gain; s->pll_freq = h->freq; s->pll_phase = h->phase;
  s->pll_lock_cnt = h->lock_cnt; s->decoder_calls = h->decoder_calls; s->n_pps = h->n_pps;
}
int orc_fm_mpf_coeffs(void *p, float *out, int cap) {
  OrcFm *h = (OrcFm *)p;
  for (int i = 0; i < h->mpf_n && i < cap; i++) { out[2 * i] = h->mpf_cre[i]; out[2 * i + 1] = h->mpf_cim[i]; }
  return h->mpf_n;
}
int orc_fm_pps(void *p, double *out, int cap) {
  OrcFm *h = (OrcFm *)p;
  for (int i = 0; i < h->n_pps && i < cap && i < 16; i++) { out[3 * i] = h->pps[3 * i]; out[3 * i + 1] = h->pps[3 * i + 1]; out[3 * i + 2] = h->pps[3 * i + 2]; }
  return h->n_pps;
}

/* ------------------------------------------------------------------ AM chain ---------- */
/* FineTuner (FineTuner.cpp:25-72): table-driven NCO multiply, index carried across calls. */
typedef struct { float *re, *im; unsigned size, index; } Tuner;
static void tuner_init(Tuner *t, unsigned table_size, int freq_shift) {
  t->size = table_size; t->index = 0;
  t->re = (float *)malloc(sizeof(float) * table_size); t->im = (float *)malloc(sizeof(float) * table_size);
  const double phase_step = 2.0 * M_PI / (double)table_size;                     /* FineTuner.cpp:41 */
  for (unsigned i = 0; i < table_size; i++) {
    /* ((int64_t)freq_shift * i) % m_table_size with m_table_size unsigned int: int64 remainder,
       sign of the dividend (FineTuner.cpp:43) */
    const int64_t r = ((int64_t)freq_shift * (int64_t)i) % (int64_t)table_size;
    const double phi = (double)r * phase_step;
    t->re[i] = (float)cos(phi); t->im[i] = (float)sin(phi);
  }
}
static void tuner_process(Tuner *t, const float *in, int n, float *out) {        /* FineTuner.cpp:55-70 */
  unsigned idx = t->index;
  for (int i = 0; i < n; i++) {
    const float a = in[2 * i], b = in[2 * i + 1], c = t->re[idx], d = t->im[idx];
    out[2 * i] = a * c - b * d;
    out[2 * i + 1] = a * d + b * c;
    if (++idx == t->size) idx = 0;
  }
  t->index = idx;
}
typedef struct {
  double ifrate; int fs4; unsigned fs4_idx; int mode;
  const ChainDesc *ifc; R8Lane if_re, if_im;
  FirIQ amf, cwf, ssbf;
  Tuner cw_up, ws_up, ws_down;
  float if_gain, if_max, if_rate, if_rms, baseband_mean, baseband_level;
  Biquad dcblock; FoIir deemph;
  double af_gain, af_max, af_ref, af_rate;
  uint64_t decoder_calls;
} OrcAm;

/* mode: ModType value (include/SoftFM.h:49): 2 AM, 3 DSB, 4 USB, 5 LSB, 6 CW, 7 WSPR */
void *orc_am_create_mode(double ifrate, int fs4, int filter, int mode) {
  OrcAm *h = (OrcAm *)calloc(1, sizeof(OrcAm));
  h->ifrate = ifrate; h->fs4 = fs4; h->mode = mode;
  if (ifrate != 48000.0) {
    h->ifc = find_chain(ifrate, 48000.0, 0);
    if (!h->ifc) { free(h); return 0; }
    r8_init(&h->if_re, h->ifc); r8_init(&h->if_im, h->ifc);
  }
  const float *c = k_jj1bdx_am_48khz_default; int nt = 255;                     /* main.cpp:785-810 */
  if (filter == 1) c = k_jj1bdx_am_48khz_medium; else if (filter == 2) c = k_jj1bdx_am_48khz_narrow;
  else if (filter == 3) { c = k_jj1bdx_am_48khz_wide; nt = 127; }
  firiq_init(&h->amf, c, nt);
  firiq_init(&h->cwf, k_jj1bdx_cw_48khz_500hz, 2049);                           /* AmDecode.cpp:35 */
  firiq_init(&h->ssbf, k_jj1bdx_ssb_48khz_1500hz, 2049);                        /* AmDecode.cpp:39 */
  hp_init(&h->dcblock, 60.0 / 48000.0);                                         /* AmDecode.cpp:45 */
  rc_init(&h->deemph, 100.0 * 48000.0 * 1.0e-6);                                /* AmDecode.cpp:49 */
  const int ssbcw = (mode == 4 || mode == 5 || mode == 6 || mode == 7), cw = (mode == 6 || mode == 7);
  h->af_gain = 1.0; h->af_max = 1.5; h->af_ref = ssbcw ? 0.24 : 0.6; h->af_rate = cw ? 0.00125 : 0.001; /* :54-66 */
  h->if_gain = 1.0f; h->if_max = 1000000.0f; h->if_rate = cw ? 0.0006f : 0.0003f;                      /* :71-77 */
  tuner_init(&h->cw_up, 480, 5);                                                /* AmDecode.cpp:82 */
  tuner_init(&h->ws_up, 480, 15); tuner_init(&h->ws_down, 480, -15);            /* AmDecode.cpp:88-89 */
  return h;
}
void *orc_am_create(double ifrate, int fs4, int filter) { return orc_am_create_mode(ifrate, fs4, filter, 2); }
void orc_am_destroy(void *p) {
  OrcAm *h = (OrcAm *)p;
  if (!h) return;
  if (h->ifc) { r8_free(&h->if_re); r8_free(&h->if_im); }
  free(h->amf.sre); free(h->amf.sim); free(h->cwf.sre); free(h->cwf.sim); free(h->ssbf.sre); free(h->ssbf.sim);
  free(h->cw_up.re); free(h->cw_up.im); free(h->ws_up.re); free(h->ws_up.im); free(h->ws_down.re); free(h->ws_down.im);
  free(h);
}
/* main.cpp:912-971 + AmDecoder::process (AmDecode.cpp:96-218), all modes */
int orc_am_process_block(void *p, const float *iq, int n, double *audio, int cap) {
  OrcAm *h = (OrcAm *)p;
  float *ifs = 0;
  const int m = front_end(h->fs4, &h->fs4_idx, h->ifc, &h->if_re, &h->if_im, iq, n, &ifs);
  if (m == 0) { free(ifs); return 0; }
  if (m > cap) { free(ifs); return -1; }
  h->decoder_calls++;
  float *x = (float *)malloc(sizeof(float) * 2 * (size_t)m), *t1 = (float *)malloc(sizeof(float) * 2 * (size_t)m);
  switch (h->mode) {                                                            /* :97-152 */
  case 4: tuner_process(&h->ws_down, ifs, m, x); firiq_process(&h->ssbf, x, m, t1); tuner_process(&h->ws_up, t1, m, x); break;
  case 5: tuner_process(&h->ws_up, ifs, m, x); firiq_process(&h->ssbf, x, m, t1); tuner_process(&h->ws_down, t1, m, x); break;
  case 6: firiq_process(&h->cwf, ifs, m, t1); tuner_process(&h->cw_up, t1, m, x); break;
  case 7: tuner_process(&h->ws_down, ifs, m, x); firiq_process(&h->cwf, x, m, t1); tuner_process(&h->ws_up, t1, m, x); break;
  default: firiq_process(&h->amf, ifs, m, x); break;                            /* AM, DSB :101 */
  }
  free(t1);
  {
    float level = 0;
    for (int i = 0; i < m; i++) level += x[2 * i] * x[2 * i] + x[2 * i + 1] * x[2 * i + 1];
    h->if_rms = sqrtf(level / (float)m);                                        /* :154 */
  }
  if_agc(&h->if_gain, h->if_max, h->if_rate, x, m);                             /* :157 */
  float vsum = 0, vsq = 0;
  for (int i = 0; i < m; i++) {
    /* demodulate_am :221-226 (magnitude) / demodulate_dsb :229-234 (real part) */
    const float mag = (h->mode == 2) ? sqrtf(x[2 * i] * x[2 * i] + x[2 * i + 1] * x[2 * i + 1]) : x[2 * i];
    vsum += mag; vsq += mag * mag;
    double v = bq_process(&h->dcblock, (double)mag);                            /* :194 */
    /* AfSimpleAgc::process (AfSimpleAgc.cpp:36-58) */
    const double x2 = v * h->af_gain;
    const double o = x2 * h->af_ref;
    const double z = 1.0 + (h->af_rate * (1.0 - (x2 * x2)));
    h->af_gain *= z;
    if (!isfinite(h->af_gain)) h->af_gain = 1.0; else if (h->af_gain > h->af_max) h->af_gain = h->af_max;
    audio[i] = (h->mode == 2) ? fo_process(&h->deemph, o) : o;                  /* :212-214: deemphasis for AM only */
  }
  h->baseband_mean = (float)(0.95 * h->baseband_mean + 0.05 * (vsum / (float)m));  /* :206-209 */
  h->baseband_level = (float)(0.95 * h->baseband_level + 0.05 * sqrtf(vsq / (float)m));
  free(x); free(ifs);
  return m;
}
typedef struct { double baseband_level; float af_agc_gain, if_agc_gain, if_rms; uint64_t decoder_calls; } OrcAmStats;
void orc_am_stats(void *p, OrcAmStats *s) {
  OrcAm *h = (OrcAm *)p;
  s->baseband_level = h->baseband_level; s->af_agc_gain = (float)h->af_gain; s->if_agc_gain = h->if_gain;
  s->if_rms = h->if_rms; s->decoder_calls = h->decoder_calls;
}

/* ------------------------------------------------------------------ NBFM chain -------- */
typedef struct {
  double ifrate; int fs4; unsigned fs4_idx;
  const ChainDesc *ifc; R8Lane if_re, if_im;
  FirIQ nbf; FirAudio audiof;
  float agc_gain, agc_max, agc_rate, if_rms, baseband_mean, baseband_level;
  float disc_norm, disc_bound, disc_save;
  double freq_dev;
  uint64_t decoder_calls;
} OrcNbfm;

/* NbfmDecoder::NbfmDecoder (NbfmDecode.cpp:24-45), filter choice main.cpp:785-810 */
void *orc_nbfm_create(double ifrate, int fs4, int filter, double freq_dev) {
  OrcNbfm *h = (OrcNbfm *)calloc(1, sizeof(OrcNbfm));
  h->ifrate = ifrate; h->fs4 = fs4; h->freq_dev = freq_dev;
  if (ifrate != 48000.0) {
    h->ifc = find_chain(ifrate, 48000.0, 0);
    if (!h->ifc) { free(h); return 0; }
    r8_init(&h->if_re, h->ifc); r8_init(&h->if_im, h->ifc);
  }
  const float *c = k_jj1bdx_nbfm_48khz_default;
  if (filter == 1) c = k_jj1bdx_nbfm_48khz_medium; else if (filter == 2) c = k_jj1bdx_nbfm_48khz_narrow;
  else if (filter == 3) c = k_jj1bdx_nbfm_48khz_wide;
  firiq_init(&h->nbf, c, 127);
  fira_init(&h->audiof, k_jj1bdx_48khz_nbfmaudio, 63);                            /* NbfmDecode.cpp:39 */
  {
    const double max_freq_dev = freq_dev / 48000.0;                               /* NbfmDecode.cpp:35 */
    h->disc_norm = (float)(max_freq_dev * 2.0 * M_PI);                            /* PhaseDiscriminator.cpp:27-30 */
    h->disc_bound = (float)(1.0 / (max_freq_dev * 2.0));
  }
  h->agc_gain = 1.0f; h->agc_max = 100000.0f; h->agc_rate = 0.0001f;              /* NbfmDecode.cpp:43 */
  return h;
}
void orc_nbfm_destroy(void *p) {
  OrcNbfm *h = (OrcNbfm *)p;
  if (!h) return;
  if (h->ifc) { r8_free(&h->if_re); r8_free(&h->if_im); }
  free(h->nbf.sre); free(h->nbf.sim); free(h->audiof.state); free(h);
}
/* main.cpp:912-961 + NbfmDecoder::process (NbfmDecode.cpp:47-96) */
int orc_nbfm_process_block(void *p, const float *iq, int n, double *audio, int cap) {
  OrcNbfm *h = (OrcNbfm *)p;
  float *ifs = 0;
  const int m = front_end(h->fs4, &h->fs4_idx, h->ifc, &h->if_re, &h->if_im, iq, n, &ifs);
  if (m == 0) { free(ifs); return 0; }
  if (m > cap) { free(ifs); return -1; }
  h->decoder_calls++;
  float *x = (float *)malloc(sizeof(float) * 2 * (size_t)m);
  firiq_process(&h->nbf, ifs, m, x);                                              /* :51 */
  {
    float level = 0;
    for (int i = 0; i < m; i++) level += x[2 * i] * x[2 * i] + x[2 * i + 1] * x[2 * i + 1];
    h->if_rms = sqrtf(level / (float)m);                                          /* :54 */
  }
  if_agc(&h->agc_gain, h->agc_max, h->agc_rate, x, m);                            /* :57 */
  double *bb = (double *)malloc(sizeof(double) * (size_t)m), *flt = (double *)malloc(sizeof(double) * (size_t)m);
  float vsum = 0, vsq = 0;
  {
    const float inv = 1.0f / h->disc_norm;                                        /* :60, PhaseDiscriminator.cpp:33-46 */
    float prev = h->disc_save, last = 0;
    for (int i = 0; i < m; i++) {
      const float ph = atan2f(x[2 * i + 1], x[2 * i]) * inv;
      float d = ph - prev;
      if (d > h->disc_bound) d -= 2 * h->disc_bound;
      if (d < -h->disc_bound) d += 2 * h->disc_bound;
      prev = ph; last = ph;
      if (isnan(d)) d = 0.0f;
      bb[i] = (double)d; vsum += d; vsq += d * d;                                 /* :71-72, :84 */
    }
    h->disc_save = last;
  }
  h->baseband_mean = (float)(0.95 * h->baseband_mean + 0.05 * (vsum / (float)m));  /* :85-86 */
  h->baseband_level = (float)(0.95 * h->baseband_level + 0.05 * sqrtf(vsq / (float)m));
  fira_process(&h->audiof, bb, m, flt);                                           /* :89 */
  const double audio_gain = pow(10.0, (-3.0 / 20.0));                             /* :92-93 */
  for (int i = 0; i < m; i++) audio[i] = flt[i] * audio_gain;
  free(x); free(ifs); free(bb); free(flt);
  return m;
}
typedef struct { float tuning_offset, baseband_level, if_rms, if_agc_gain; uint64_t decoder_calls; } OrcNbfmStats;
void orc_nbfm_stats(void *p, OrcNbfmStats *s) {
  OrcNbfm *h = (OrcNbfm *)p;
  s->tuning_offset = (float)(h->baseband_mean * h->freq_dev);                      /* NbfmDecode.h:59 */
  s->baseband_level = h->baseband_level; s->if_rms = h->if_rms; s->if_agc_gain = h->agc_gain;
  s->decoder_calls = h->decoder_calls;
}

/* Cumulative release schedule of a chain (same integer model the product's host code uses,
   restated independently): outputs released after N inputs. */
int64_t orc_chain_out(double src, double dst, int kind, int64_t n) {
  const ChainDesc *d = find_chain(src, dst, kind);
  if (!d) return -1;
  for (int s = 0; s < d->n_hb; s++) { n = n / 2 - (d->hb[s].ntaps - 1); if (n < 0) n = 0; }
  int64_t c = n - d->bc.latency; if (c < 0) c = 0;
  if (d->bc.down > 1) c = (c + d->bc.down - 1) / d->bc.down;
  if (!d->has_fi) return c;
  const int fl2 = d->fi.flen / 2;
  if (c < fl2 + 1) return 0;
  const int64_t a = (c - fl2) * (int64_t)d->fi.outstep;
  return (a + d->fi.instep - 1) / d->fi.instep;
}

/* A stand-alone resampler lane for stage-level checks. */
void *orc_r8_create(double src, double dst, int kind) {
  const ChainDesc *d = find_chain(src, dst, kind);
  if (!d) return 0;
  R8Lane *r = (R8Lane *)calloc(1, sizeof(R8Lane));
  r8_init(r, d);
  return r;
}
int orc_r8_process(void *p, const double *in, int n, double *out, int cap) {
  Stream o; st_init(&o);
  int m = r8_process((R8Lane *)p, in, n, &o);
  if (m > cap) { st_free(&o); return -1; }
  memcpy(out, o.v, sizeof(double) * (size_t)m);
  st_free(&o);
  return m;
}
void orc_r8_destroy(void *p) { if (p) { r8_free((R8Lane *)p); free(p); } }
