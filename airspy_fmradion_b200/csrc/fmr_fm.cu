// fmr_fm.cu — FM broadcast handle: the C ABI of include/fmradion_b200.h for the path
// FourthConverterIQ -> IfResampler -> FmDecoder::process (main.cpp:912-956,
// FmDecode.cpp:85-221), many channels per launch.
#include <algorithm>
#include <complex>
#include <cmath>
#include <cstdlib>

#include "fmr_core.cuh"
#include "fmr_host.cuh"
#include "fmr_io.cuh"
#include "fmr_mpf.cuh"

namespace fmr {
thread_local std::string g_err;
}

using namespace fmr;

constexpr int kMaxHostChunks = 8;       // time chunks of fmr_fm_process_host's copy/compute pipeline
constexpr int kHostChunkMinBlocks = 16; // a chunk is at least this many source blocks

// The FM audio chain between the 384 kHz core and the 48 kHz DC block, in sample type AS (see aud_mk in fmr_kernels.cuh).
template <typename AS> struct AudioPath {
  using AV = typename V2<AS>::type;
  Resampler<AS> aures;       // 2x AudioResampler (mono, L-R as one complex stream)
  Ring<AV> r_384{nullptr, 0}; // (mono, L-R) after deemphasis
  Ring<AV> r_48a{nullptr, 0}; // audio resampler output
  Ring<AV> r_48b{nullptr, 0}; // after pilot-cut FIR
  AS *d_pilotcut = nullptr;
};

struct fmr_fm {
  fmr_fm_config cfg;
  bool core_fused = true; // FMR_CORE_FUSED=0: AGC / discriminator / PLL as separate launches
  int rot_sms = 148;      // SM count: CTAs that share an SM rotate the warp roles of the fused core
  int C = 0;
  const ChainDesc *ifc = nullptr; // null when input_rate == 384000 (no IfResampler, main.cpp:778)
  const ChainDesc *auc = nullptr;
  DevMem mem;
  PinnedSlots slots;
  cudaStream_t own_stream = nullptr;

  Resampler<float> ifres;
  float2 *hist[2] = {nullptr, nullptr};
  int hist_cur = 0;
  Ring<float2> r_if{nullptr, 0};   // 384 kHz decoder input
  Ring<float2> r_iff{nullptr, 0};  // after the optional IF filter
  Ring<float2> r_agc{nullptr, 0};  // after AGC (multipath path only)
  Ring<float2> r_mpf{nullptr, 0};  // after multipath filter
  Ring<float> r_mpx{nullptr, 0};   // discriminator output
  float *d_stats = nullptr;        // per (channel, block): if_rms, baseband mean, baseband rms
  AudioPath<float> a32;   // the audio chain of this handle: FP32 filters (default) ...
  AudioPath<double> a64;  // ... or the all-FP64 chain (FMR_AUDIO_FP64=1); only one of the two is initialised
  bool audio_f64 = false;
  FmChanState *d_state = nullptr;
  uint8_t *d_flags = nullptr;
  PpsEventDev *d_pps = nullptr;
  uint32_t *d_e384 = nullptr, *d_e48 = nullptr;
  float *d_fmfilter = nullptr;
  int fmfilter_taps = 0;
  float *d_atan = nullptr;
  MpfDev mpf;
  Prof prof;
  int p_hist = -1, p_fmf = -1, p_core = -1, p_agc = -1, p_mpf = -1, p_core2 = -1, p_pcut = -1, p_tail = -1, p_fused = -1;
  FmCoreParams core;
  FmTailParams tail;

  int64_t cum_in = 0, cum384 = 0, cum48 = 0;
  uint32_t last_blocks = 0;
  // chunked host calls: where the current process_device launch stores its per-block flags,
  // and the (first block, count) of every launch of the last call, for fmr_fm_block_flags
  int iq_fmt = 0; // sample format of the d_iq pointer of the call in progress (0 cf32, 1 int16 pairs)
  bool in_host_call = false;
  uint32_t host_b0 = 0;
  std::vector<std::pair<uint32_t, uint32_t>> last_chunks;
  uint32_t last_launches = 0;
  int64_t last_t0 = 0, last_t1 = 0;

  // host staging for the *_host entry point
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_in[kMaxHostChunks] = {nullptr}, ev_done[kMaxHostChunks] = {nullptr};
  float *d_iq = nullptr;
  double *d_audio = nullptr;
  size_t audio_cap = 0; // doubles per channel

  // file-format ingest and output stage (fmr_fm_process_*_io, fmr_io.cuh)
  uint8_t *d_raw = nullptr; // raw file samples of formats that are converted by their own launch
  size_t raw_cap = 0;       // bytes
  uint8_t *d_out = nullptr; // sink-format audio of the host entry point, [C][audio_cap] values
  BlockLevelDev *d_levels = nullptr; // chunk-major like d_flags
  bool sink_on = false;     // set for the duration of an *_io call with an out_cfg
  SinkParams sink{};
  uint8_t *sink_out = nullptr; // where the launch in progress stores value (c, i): sink_out[(c*stride + i) * bytes]
  size_t sink_out_stride = 0;
  bool have_levels = false;
};

static int64_t if_out_total(const fmr_fm *h, int64_t n) { return h->ifc ? chain_out(h->ifc, n) : n; }

static void hp_coeffs(double cutoff, double *b0, double *b1, double *b2, double *a1, double *a2) {
  // HighPassFilterIir::HighPassFilterIir (Filter.cpp:254-290)
  using CD = std::complex<double>;
  const double w = 2 * M_PI * cutoff;
  const CD p1s = w / std::exp((2 * 1 + 2 - 1) / double(2 * 2) * CD(0, M_PI));
  const CD p1z = std::exp(p1s);
  double B0 = 1, B1 = -2, B2 = 1;
  const double A1 = -2 * std::real(p1z);
  const double A2 = std::abs(p1z * p1z);
  const double g = (B0 - B1 + B2) / (1 - A1 + A2);
  *b0 = B0 / g;
  *b1 = B1 / g;
  *b2 = B2 / g;
  *a1 = A1;
  *a2 = A2;
}

extern "C" const char *fmr_last_error(void) { return g_err.c_str(); }
extern "C" const char *fmr_version(void) { return "fmradion_b200 0.1 (sm_100a)"; }
extern "C" int fmr_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return n;
}

static fmr_status fm_build(fmr_fm *h) {
  const fmr_fm_config &cfg = h->cfg;
  FMR_CUDA(cudaSetDevice(cfg.device));
  const int C = h->C = (int)cfg.n_channels;
  const int64_t max_in = cfg.max_samples_per_call;
  const int max_blocks = (int)cfg.max_blocks_per_call;
  if (cfg.input_rate != 384000.0) {
    h->ifc = find_chain(cfg.input_rate, 384000.0, 0);
    if (!h->ifc) return fail(FMR_ERR_UNSUPPORTED, "no resampler tables for this input_rate -> 384000");
  }
  h->auc = find_chain(384000.0, 48000.0, 1);
  if (!h->auc) return fail(FMR_ERR_UNSUPPORTED, "audio resampler tables missing");
  FMR_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  FMR_CUDA(h->slots.init(4 * sizeof(uint32_t) * (size_t)max_blocks));
  int64_t max384 = max_in + 8;
  if (h->ifc) {
    fmr_status s = h->ifres.init(h->ifc, C, max_in, true, h->mem);
    if (s != FMR_OK) return s;
    max384 = chain_out(h->ifc, max_in) - chain_out(h->ifc, 0) + 8;
    // the per-call maximum can exceed the from-zero count by the start-up latency
    max384 = (int64_t)std::ceil((double)max_in * 384000.0 / cfg.input_rate) + 8;
  } else {
    HbTaps<float> t;
    memset(&t, 0, sizeof(t));
    FMR_CUDA((cudaFuncSetAttribute(k_hb_cascade<float, 0, true, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)Resampler<float>::hb_smem(t, 0))));
  }
  FMR_CUDA(h->mem.alloc(&h->hist[0], (size_t)C * kHist));
  FMR_CUDA(h->mem.alloc(&h->hist[1], (size_t)C * kHist));
  // + one block of the frequency-domain resampler's grid: it produces whole blocks, up to 2303 samples ahead of the
  // reference's release schedule (Resampler::run, fmr_fdr.cuh); up to 3071 for the 125:48 pairs
  h->r_if.cap = pow2ceil((uint64_t)max384 + 512 + fdr::kMaxAdvOut);
  FMR_CUDA(h->mem.alloc(&h->r_if.base, (size_t)C * h->r_if.cap));
  h->r_iff = h->r_if;
  if (cfg.fmfilter) {
    const float *tbl = (cfg.fmfilter == 1) ? k_jj1bdx_fm_384kHz_medium : k_jj1bdx_fm_384kHz_narrow;
    h->fmfilter_taps = 127;
    if (cfg.fmfilter == 3) {
      if (!cfg.fmfilter_coeff || cfg.fmfilter_ntaps < 2 || cfg.fmfilter_ntaps > 4096) {
        return fail(FMR_ERR_INVALID, "fmfilter == 3 needs fmfilter_coeff with 2..4096 taps");
      }
      tbl = cfg.fmfilter_coeff;
      h->fmfilter_taps = (int)cfg.fmfilter_ntaps;
    }
    FMR_CUDA(h->mem.alloc(&h->d_fmfilter, (size_t)h->fmfilter_taps, false));
    FMR_CUDA(cudaMemcpy(h->d_fmfilter, tbl, h->fmfilter_taps * sizeof(float), cudaMemcpyHostToDevice));
    h->r_iff.cap = h->r_if.cap;
    FMR_CUDA(h->mem.alloc(&h->r_iff.base, (size_t)C * h->r_iff.cap));
  }
  h->r_agc.cap = h->r_if.cap;
  FMR_CUDA(h->mem.alloc(&h->r_agc.base, (size_t)C * h->r_agc.cap));
  h->r_mpx.cap = h->r_if.cap;
  FMR_CUDA(h->mem.alloc(&h->r_mpx.base, (size_t)C * h->r_mpx.cap));
  FMR_CUDA(h->mem.alloc(&h->d_stats, (size_t)C * max_blocks * 3));
  if (cfg.multipath_stages > 0) {
    h->r_mpf.cap = h->r_if.cap;
    FMR_CUDA(h->mem.alloc(&h->r_mpf.base, (size_t)C * h->r_mpf.cap));
    fmr_status s = h->mpf.init(cfg.multipath_stages, C, h->mem);
    if (s != FMR_OK) return s;
  }
  const int64_t max48 = max384 / 8 + 16;
  if (const char *ev = getenv("FMR_AUDIO_FP64")) h->audio_f64 = atoi(ev) != 0;
  auto init_audio = [&](auto &ap) -> fmr_status {
    using AS = typename std::remove_reference<decltype(ap.aures)>::type::Scalar;
    ap.r_384.cap = pow2ceil((uint64_t)max384 + 512);
    FMR_CUDA(h->mem.alloc(&ap.r_384.base, (size_t)C * ap.r_384.cap));
    fmr_status s = ap.aures.init(h->auc, C, max384, false, h->mem);
    if (s != FMR_OK) return s;
    ap.r_48a.cap = pow2ceil((uint64_t)max48 + 512);
    FMR_CUDA(h->mem.alloc(&ap.r_48a.base, (size_t)C * ap.r_48a.cap));
    ap.r_48b.cap = ap.r_48a.cap;
    FMR_CUDA(h->mem.alloc(&ap.r_48b.base, (size_t)C * ap.r_48b.cap));
    FMR_CUDA(h->mem.alloc(&ap.d_pilotcut, 127, false));
    AS pc[127];
    for (int i = 0; i < 127; i++) pc[i] = (AS)k_jj1bdx_48khz_fmaudio[i];
    FMR_CUDA(cudaMemcpy(ap.d_pilotcut, pc, sizeof(pc), cudaMemcpyHostToDevice));
    FMR_CUDA(cudaFuncSetAttribute(k_fm_tail<typename V2<AS>::type>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(TailSmem<typename V2<AS>::type>)));
    return FMR_OK;
  };
  {
    const fmr_status s = h->audio_f64 ? init_audio(h->a64) : init_audio(h->a32);
    if (s != FMR_OK) return s;
  }
  FMR_CUDA(h->mem.alloc(&h->d_state, (size_t)C));
  FMR_CUDA(h->mem.alloc(&h->d_flags, (size_t)C * max_blocks));
  FMR_CUDA(h->mem.alloc(&h->d_pps, (size_t)C * kMaxPps));
  FMR_CUDA(h->mem.alloc(&h->d_e384, (size_t)max_blocks));
  FMR_CUDA(h->mem.alloc(&h->d_e48, (size_t)max_blocks));
  FMR_CUDA(h->mem.alloc(&h->d_atan, 257, false));
  FMR_CUDA(cudaMemcpy(h->d_atan, k_fast_atan_table, 257 * sizeof(float), cudaMemcpyHostToDevice));

  FMR_CUDA(raise_smem_limit(k_fir_quirk<float>, fq_smem(std::max(h->fmfilter_taps, 127), sizeof(float2), sizeof(float))));
  FMR_CUDA(raise_smem_limit(k_fir_quirk<double>, fq_smem(127, sizeof(double2), sizeof(double))));
  // initial state (constructors: FmDecode.cpp:25-83, PilotPhaseLock.cpp:35-54, IfSimpleAgc.cpp:22-32)
  {
    std::vector<FmChanState> st(C);
    memset(st.data(), 0, sizeof(FmChanState) * C);
    for (int c = 0; c < C; c++) {
      st[c].agc_gain = 1.0f;
      st[c].mpf_wait = 100; // FmDecode.cpp:33
      st[c].pll_freq = (19000.0 / 384000.0) * 2.0 * M_PI;
    }
    FMR_CUDA(cudaMemcpy(h->d_state, st.data(), sizeof(FmChanState) * C, cudaMemcpyHostToDevice));
  }
  FmCoreParams &P = h->core;
  memset(&P, 0, sizeof(P));
  P.agc_max = 100000.0f;
  P.agc_rate = 0.0001f;
  {
    const double max_freq_dev = 75000.0 / 384000.0; // FmDecode.cpp:52
    const double norm = max_freq_dev * 2.0 * M_PI;  // PhaseDiscriminator.cpp:27
    P.disc_inv_norm = 1.0f / (float)norm;
    P.disc_bound = (float)(1.0 / (max_freq_dev * 2.0)); // PhaseDiscriminator.cpp:29
  }
  {
    const double freq = 19000.0 / 384000.0, bw = 30.0 / 384000.0;
    P.pll_minfreq = (freq - bw) * 2.0 * M_PI;
    P.pll_maxfreq = (freq + bw) * 2.0 * M_PI;
    P.lock_delay = int(15.0 / bw);
  }
  P.bq_b0 = 1.46974784e-06;
  P.bq_a1 = -1.99682419;
  P.bq_a2 = 0.996825659;
  P.lf_b0 = 0.000304341788;
  P.lf_b1 = -0.000304324564;
  P.minsignal = 0.001;
  {
    const double tc = (cfg.deemphasis_us == 0) ? 1.0 : (cfg.deemphasis_us * 384000.0 * 1.0e-6); // FmDecode.cpp:67-70
    P.de_a1 = -std::exp(-1 / tc);
    P.de_b0 = 1 + P.de_a1;
  }
  P.stereo = cfg.stereo ? 1 : 0;
  P.pilot_shift = cfg.pilot_shift ? 1 : 0;
  P.deemph_on_stereo = cfg.pilot_shift ? 0 : 1;
  P.n_channels = C;
  FmTailParams &T = h->tail;
  hp_coeffs(0.0001, &T.b0, &T.b1, &T.b2, &T.a1, &T.a2);
  T.stereo = P.stereo;
  T.pilot_shift = P.pilot_shift;
  T.n_channels = C;

  h->audio_cap = (size_t)(max48 * (cfg.stereo ? 2 : 1));
  // profiling stage registry (order = pipeline order)
  h->ifres.prof = &h->prof;
  h->ifres.p_hb = h->prof.add("if_halfband_cascade");
  h->ifres.p_bc = h->prof.add("if_lowpass");
  h->ifres.p_fi = h->prof.add("if_polyphase");
  h->ifres.p_fe = h->prof.add("if_frontend_fused");
  h->p_hist = h->prof.add("save_hist");
  h->p_fmf = h->prof.add("fm_if_filter");
  h->p_fused = h->prof.add("fm_core_fused");
  h->p_agc = h->prof.add("fm_agc");
  h->p_mpf = h->prof.add("fm_multipath");
  h->p_core = h->prof.add("fm_discriminator_stats");
  h->p_core2 = h->prof.add("fm_pll_stereo_deemph");
  {
    const int p_hb = h->prof.add("audio_halfband_cascade"), p_bc = h->prof.add("audio_lowpass");
    h->a32.aures.prof = h->a64.aures.prof = &h->prof;
    h->a32.aures.p_hb = h->a64.aures.p_hb = p_hb;
    h->a32.aures.p_bc = h->a64.aures.p_bc = p_bc;
  }
  h->p_pcut = h->prof.add("pilot_cut_fir");
  h->p_tail = h->prof.add("dcblock_matrix");
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_create(const fmr_fm_config *cfg, fmr_fm **out) {
  if (!cfg || !out) return fail(FMR_ERR_INVALID, "null argument");
  if (cfg->n_channels == 0 || cfg->max_samples_per_call == 0 || cfg->max_blocks_per_call == 0) {
    return fail(FMR_ERR_INVALID, "n_channels, max_samples_per_call and max_blocks_per_call must be > 0");
  }
  if (cfg->fmfilter < 0 || cfg->fmfilter > 3) return fail(FMR_ERR_INVALID, "fmfilter must be 0..3");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    return fail(FMR_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  fmr_fm *h = new fmr_fm();
  h->cfg = *cfg;
  if (const char *e = getenv("FMR_CORE_FUSED")) h->core_fused = atoi(e) != 0;
  {
    int sms = fmr_device_sm_count(cfg->device);
    h->rot_sms = sms > 0 ? sms : 148;
  }
  fmr_status s = fm_build(h);
  if (s != FMR_OK) {
    std::string keep = g_err;
    fmr_fm_destroy(h);
    g_err = keep;
    return s;
  }
  *out = h;
  return FMR_OK;
}

extern "C" void fmr_fm_destroy(fmr_fm *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  h->mem.release();
  h->slots.release();
  h->prof.release();
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  for (int k = 0; k < kMaxHostChunks; k++) {
    if (h->ev_in[k]) cudaEventDestroy(h->ev_in[k]);
    if (h->ev_done[k]) cudaEventDestroy(h->ev_done[k]);
  }
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

static fmr_status fm_schedule(const fmr_fm *h, const uint32_t *block_len, uint32_t n_blocks, uint32_t *e384,
                              uint32_t *e48, uint64_t *total_in) {
  int64_t n = 0;
  const int64_t base384 = if_out_total(h, h->cum_in);
  const int64_t base48 = chain_out(h->auc, h->cum384);
  if (base384 != h->cum384) return fail(FMR_ERR_INVALID, "internal: stream position mismatch");
  for (uint32_t b = 0; b < n_blocks; b++) {
    // IfResampler asserts input <= 65536 per call (IfResampler.cpp:41, IfResampler.h:31)
    if (h->ifc && block_len[b] > 65536) return fail(FMR_ERR_INVALID, "block_len > 65536 (IfResampler limit)");
    n += block_len[b];
    const int64_t c384 = if_out_total(h, h->cum_in + n);
    e384[b] = (uint32_t)(c384 - base384);
    // AudioResampler asserts input <= 32768 per call (AudioResampler.cpp:41)
    const uint32_t prev = b ? e384[b - 1] : 0;
    if (e384[b] - prev > 32768) return fail(FMR_ERR_INVALID, "block yields > 32768 IF samples (AudioResampler limit)");
    e48[b] = (uint32_t)(chain_out(h->auc, c384) - base48);
  }
  *total_in = (uint64_t)n;
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_schedule(double input_rate, int stereo, uint64_t start_sample, const uint32_t *block_len,
                                      uint32_t n_blocks, uint32_t *if_len, uint32_t *audio_len) {
  if (!block_len) return fail(FMR_ERR_INVALID, "null argument");
  const ChainDesc *ifc = nullptr;
  if (input_rate != 384000.0) {
    ifc = find_chain(input_rate, 384000.0, 0);
    if (!ifc) return fail(FMR_ERR_UNSUPPORTED, "no resampler tables for this input_rate -> 384000");
  }
  const ChainDesc *auc = find_chain(384000.0, 48000.0, 1);
  int64_t n = (int64_t)start_sample;
  int64_t p384 = ifc ? chain_out(ifc, n) : n;
  int64_t p48 = chain_out(auc, p384);
  const uint32_t w = stereo ? 2 : 1;
  for (uint32_t b = 0; b < n_blocks; b++) {
    n += block_len[b];
    const int64_t c384 = ifc ? chain_out(ifc, n) : n;
    const int64_t c48 = chain_out(auc, c384);
    if (if_len) if_len[b] = (uint32_t)(c384 - p384);
    if (audio_len) audio_len[b] = (uint32_t)(c48 - p48) * w;
    p384 = c384;
    p48 = c48;
  }
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_query_output(fmr_fm *h, const uint32_t *block_len, uint32_t n_blocks,
                                          uint64_t *audio_doubles_total, uint32_t *audio_len) {
  if (!h || !block_len) return fail(FMR_ERR_INVALID, "null argument");
  std::vector<uint32_t> e384(n_blocks), e48(n_blocks);
  uint64_t tot = 0;
  fmr_status s = fm_schedule(h, block_len, n_blocks, e384.data(), e48.data(), &tot);
  if (s != FMR_OK) return s;
  const uint32_t w = h->cfg.stereo ? 2 : 1;
  if (audio_len) {
    for (uint32_t b = 0; b < n_blocks; b++) audio_len[b] = (e48[b] - (b ? e48[b - 1] : 0)) * w;
  }
  if (audio_doubles_total) *audio_doubles_total = n_blocks ? (uint64_t)e48[n_blocks - 1] * w : 0;
  return FMR_OK;
}

static fmr_status fm_process_device_impl(fmr_fm *h, const float *d_iq, size_t iq_stride, const uint32_t *block_len,
                                         uint32_t n_blocks, double *d_audio, size_t audio_stride,
                                         uint32_t *audio_len, void *stream);

extern "C" fmr_status fmr_fm_process_device(fmr_fm *h, const float *d_iq, size_t iq_stride,
                                            const uint32_t *block_len, uint32_t n_blocks, double *d_audio,
                                            size_t audio_stride, uint32_t *audio_len, void *stream) {
  if (!h) return fail(FMR_ERR_INVALID, "null argument");
  if (!h->in_host_call) h->iq_fmt = 0;
  return fm_process_device_impl(h, d_iq, iq_stride, block_len, n_blocks, d_audio, audio_stride, audio_len, stream);
}

extern "C" fmr_status fmr_fm_process_device_i16(fmr_fm *h, const int16_t *d_iq, size_t iq_stride,
                                                const uint32_t *block_len, uint32_t n_blocks, double *d_audio,
                                                size_t audio_stride, uint32_t *audio_len, void *stream) {
  if (!h) return fail(FMR_ERR_INVALID, "null argument");
  if (h->cfg.input_rate == 384000.0) return fail(FMR_ERR_UNSUPPORTED, "int16 ingest needs an IF resampler stage");
  h->iq_fmt = 1;
  fmr_status s = fm_process_device_impl(h, reinterpret_cast<const float *>(d_iq), iq_stride, block_len, n_blocks,
                                        d_audio, audio_stride, audio_len, stream);
  if (!h->in_host_call) h->iq_fmt = 0;
  return s;
}

template <typename AS>
static fmr_status fm_process_device_t(fmr_fm *h, AudioPath<AS> &ap, const float *d_iq, size_t iq_stride,
                                      const uint32_t *block_len, uint32_t n_blocks, double *d_audio, size_t audio_stride,
                                      uint32_t *audio_len, void *stream);

static fmr_status fm_process_device_impl(fmr_fm *h, const float *d_iq, size_t iq_stride, const uint32_t *block_len,
                                         uint32_t n_blocks, double *d_audio, size_t audio_stride,
                                         uint32_t *audio_len, void *stream) {
  if (!h || !d_iq || !block_len || !d_audio) return fail(FMR_ERR_INVALID, "null argument");
  return h->audio_f64 ? fm_process_device_t(h, h->a64, d_iq, iq_stride, block_len, n_blocks, d_audio, audio_stride, audio_len, stream)
                      : fm_process_device_t(h, h->a32, d_iq, iq_stride, block_len, n_blocks, d_audio, audio_stride, audio_len, stream);
}

template <typename AS>
static fmr_status fm_process_device_t(fmr_fm *h, AudioPath<AS> &ap, const float *d_iq, size_t iq_stride,
                                      const uint32_t *block_len, uint32_t n_blocks, double *d_audio, size_t audio_stride,
                                      uint32_t *audio_len, void *stream) {
  using AV = typename V2<AS>::type;
  if (n_blocks == 0) return FMR_OK;
  if (n_blocks > h->cfg.max_blocks_per_call) return fail(FMR_ERR_CAPACITY, "n_blocks > max_blocks_per_call");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int C = h->C;
  int slot = 0;
  uint32_t *tab = (uint32_t *)h->slots.acquire(&slot);
  uint32_t *e384 = tab, *e48 = tab + h->cfg.max_blocks_per_call;
  uint64_t total_in = 0;
  fmr_status s = fm_schedule(h, block_len, n_blocks, e384, e48, &total_in);
  if (s != FMR_OK) return s;
  if (total_in > h->cfg.max_samples_per_call) return fail(FMR_ERR_CAPACITY, "sum(block_len) > max_samples_per_call");
  if (total_in > iq_stride) return fail(FMR_ERR_INVALID, "iq_stride < sum(block_len)");
  const uint32_t w = h->cfg.stereo ? 2 : 1;
  const uint32_t n384 = e384[n_blocks - 1], n48 = e48[n_blocks - 1];
  if ((size_t)n48 * w > audio_stride) return fail(FMR_ERR_CAPACITY, "audio_stride too small for this call");
  if (audio_len) {
    for (uint32_t b = 0; b < n_blocks; b++) audio_len[b] = (e48[b] - (b ? e48[b - 1] : 0)) * w;
  }
  int launches = 0;
  Prof &pf = h->prof;
  // one profile per public call: a host call's chunks accumulate into it (fmr_fm_stage_times)
  if (!h->in_host_call || h->host_b0 == 0) pf.reset();
  const int64_t t0 = h->cum384, t1 = h->cum384 + n384;
  const int64_t j0 = h->cum48;
  // The call is one straight pipeline on the caller's stream: front end -> [IF filter] -> 384 kHz core -> audio
  // resamplers -> pilot cut -> DC block + matrix. (Two-stream pipelines over time chunks or channel groups were measured in
  // rounds 1 and 2 and are slower: chunking inflates every stage by more than the overlap hides, profiles/README.md.)
  FMR_CUDA(cudaMemcpyAsync(h->d_e384, e384, sizeof(uint32_t) * n_blocks, cudaMemcpyHostToDevice, st));
  FMR_CUDA(cudaMemcpyAsync(h->d_e48, e48, sizeof(uint32_t) * n_blocks, cudaMemcpyHostToDevice, st));
  h->slots.commit(slot, st);
  h->ifres.gc0 = 0;
  h->ifres.gcn = C;
  ap.aures.gc0 = 0;
  ap.aures.gcn = C;
  const uint32_t flag_b0 = h->in_host_call ? h->host_b0 : 0;
  if (!h->in_host_call) h->last_chunks.clear();
  h->last_chunks.push_back(std::make_pair(flag_b0, n_blocks));
  const uint32_t nb = n_blocks;
  const uint32_t *d_e384 = h->d_e384, *d_e48 = h->d_e48;
  uint8_t *d_flags = h->d_flags + (size_t)C * flag_b0;
  float *d_stats = h->d_stats;
  // ---- Fs/4 shift + IF resampler -> r_if[t0, t1)
  InSrc<float2> src;
  src.lin = reinterpret_cast<const float2 *>(d_iq);
  src.stride = iq_stride;
  src.fmt = h->iq_fmt;
  src.hist = h->hist[h->hist_cur];
  src.start = h->cum_in;
  src.n_new = (int64_t)total_in;
  src.ring = Ring<float2>{nullptr, 0};
  if (h->ifc) {
    int64_t o0, o1;
    s = h->ifres.run(src, (int64_t)total_in, h->r_if, h->cfg.fs4_shift, st, &o0, &o1, &launches, true);
    if (s != FMR_OK) return s;
    if (o0 != t0 || o1 != t1) return fail(FMR_ERR_INVALID, "internal: IF schedule mismatch");
  } else if (total_in > 0) {
    HbTaps<float> t;
    memset(&t, 0, sizeof(t));
    dim3 grid((unsigned)((total_in + kHbTile - 1) / kHbTile), C);
    pf.begin(h->ifres.p_hb, st);
    k_hb_cascade<float, 0, true, 0, 0, 0><<<grid, kHbThreads, Resampler<float>::hb_smem(t, 0), st>>>(
        src, h->r_if, t, t0, (int)total_in, h->cfg.fs4_shift);
    pf.end(h->ifres.p_hb, st);
    launches++;
  }
  if (h->ifc && total_in > 0) {
    pf.begin(h->p_hist, st);
    k_save_hist<float2><<<C, 128, 0, st>>>(src.lin, iq_stride, (int64_t)total_in, h->hist[h->hist_cur],
                                           h->hist[h->hist_cur ^ 1], h->iq_fmt);
    pf.end(h->p_hist, st);
    launches++;
  }
  if (n384 > 0) {
    // ---- optional IF filter (FmDecode.cpp:98-102)
    if (h->cfg.fmfilter) {
      dim3 grid((n384 + kQTile - 1) / kQTile, C);
      pf.begin(h->p_fmf, st);
      k_fir_quirk<float><<<grid, kQThreads, fq_smem(h->fmfilter_taps, sizeof(float2), sizeof(float)), st>>>(
          h->r_if, h->r_iff, h->d_fmfilter, h->fmfilter_taps, t0, (int)n384, d_e384, (int)nb);
      pf.end(h->p_fmf, st);
      launches++;
    }
    // ---- 384 kHz core. Stereo without the multipath filter: one warp-specialised kernel (fmr_core.cuh);
    // otherwise AGC (serial) -> [multipath] -> discriminator + statistics (parallel) -> PLL (serial).
    const bool fused = h->core_fused && h->cfg.stereo && h->cfg.multipath_stages == 0;
    dim3 cgrid((C + 31) / 32);
    const int first = (flag_b0 == 0) ? 1 : 0;
    if (fused) {
      pf.begin(h->p_fused, st);
      k_fm_core_fused<AV><<<cgrid, kCfThreads, 0, st>>>(h->r_if, h->r_iff, ap.r_384, h->d_state, d_flags, h->d_pps, d_e384, (int)nb, t0,
                                                   h->core, h->d_atan, (int)flag_b0, first, h->rot_sms);
      pf.end(h->p_fused, st);
      launches++;
    } else {
      pf.begin(h->p_agc, st);
      k_fm_agc2<<<cgrid, 32, 0, st>>>(h->r_iff, h->r_agc, h->d_state, (int)n384, t0, h->core);
      pf.end(h->p_agc, st);
      launches++;
      Ring<float2> disc_in = h->r_agc;
      if (h->cfg.multipath_stages > 0) {
        pf.begin(h->p_mpf, st);
        h->mpf.run(h->r_agc, h->r_mpf, h->d_state, d_e384, (int)nb, t0, st, 0, C);
        pf.end(h->p_mpf, st);
        launches++;
        disc_in = h->r_mpf;
      }
      pf.begin(h->p_core, st);
      {
        dim3 g1((n384 + 255) / 256, C);
        k_fm_disc<<<g1, 256, 0, st>>>(disc_in, h->r_mpx, (int)n384, t0, h->core);
        dim3 g2((nb + 3) / 4, C);
        k_fm_call_stats<<<g2, 128, 0, st>>>(h->r_if, h->r_mpx, d_stats, d_e384, (int)nb, t0);
      }
      pf.end(h->p_core, st);
      pf.begin(h->p_core2, st);
      k_fm_pll2<AV><<<cgrid, 32, 0, st>>>(h->r_mpx, ap.r_384, h->d_state, d_flags, h->d_pps, d_stats, d_e384, (int)nb, t0, h->core,
                                     h->d_atan, (int)flag_b0, first);
      pf.end(h->p_core2, st);
      launches += 3;
    }
    // ---- audio resamplers (mono and L-R in lock step, FmDecode.cpp:172-183)
    InSrc<AV> asrc;
    memset(&asrc, 0, sizeof(asrc));
    asrc.ring = ap.r_384;
    int64_t a0, a1;
    s = ap.aures.run(asrc, (int64_t)n384, ap.r_48a, 0, st, &a0, &a1, &launches, true);
    if (s != FMR_OK) return s;
    if (a0 != j0 || a1 != j0 + n48) return fail(FMR_ERR_INVALID, "internal: audio schedule mismatch");
    if (n48 > 0) {
      // ---- pilot-cut FIR (FmDecode.cpp:190,196) then DC block + matrix
      dim3 grid((n48 + kQTile - 1) / kQTile, C);
      pf.begin(h->p_pcut, st);
      k_fir_quirk<AS><<<grid, kQThreads, fq_smem(127, sizeof(AV), sizeof(AS)), st>>>(ap.r_48a, ap.r_48b, ap.d_pilotcut,
                                                                                                 127, j0, (int)n48, d_e48, (int)nb);
      pf.end(h->p_pcut, st);
      pf.begin(h->p_tail, st);
      k_fm_tail<AV><<<cgrid, kTailThreads, sizeof(TailSmem<AV>), st>>>(ap.r_48b, d_audio, audio_stride, h->d_state, d_flags, d_e48, (int)nb, j0, h->tail);
      pf.end(h->p_tail, st);
      launches += 2;
    }
    if (h->sink_on) {
      // ---- output stage of the block loop (main.cpp:977-1002): levels, squelch gain, sink format
      dim3 sg(nb, C);
      k_audio_sink<<<sg, kSinkThreads, 0, st>>>(h->r_if, t0, d_e384, d_audio, audio_stride, d_e48, (int)nb, h->sink_out,
                                                h->sink_out_stride, h->d_levels + (size_t)C * flag_b0, h->sink);
      launches++;
    }
  }
  if (h->ifc && total_in > 0) h->hist_cur ^= 1;
  FMR_CUDA(cudaGetLastError());
  h->cum_in += (int64_t)total_in;
  h->cum384 += n384;
  h->cum48 += n48;
  if (!h->in_host_call) {
    h->last_blocks = n_blocks;
    h->last_launches = (uint32_t)launches;
    h->last_t0 = t0;
  } else {
    h->last_launches += (uint32_t)launches;
  }
  h->last_t1 = t1;
  return FMR_OK;
}

// Staging buffers of the *_host and *_io entry points, allocated on first use.
static fmr_status fm_ensure_staging(fmr_fm *h, size_t raw_bytes, bool want_sink) {
  const int C = h->C;
  if (!h->d_iq) {
    FMR_CUDA(h->mem.alloc(&h->d_iq, (size_t)C * h->cfg.max_samples_per_call * 2, false));
    FMR_CUDA(h->mem.alloc(&h->d_audio, (size_t)C * h->audio_cap, false));
    FMR_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    FMR_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    for (int k = 0; k < kMaxHostChunks; k++) {
      FMR_CUDA(cudaEventCreateWithFlags(&h->ev_in[k], cudaEventDisableTiming));
      FMR_CUDA(cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
    }
  }
  if (raw_bytes > 0 && !h->d_raw) {
    // once, for the widest file format (S24: 6 bytes per IQ sample): a handle that alternates formats never reallocates
    h->raw_cap = (size_t)C * h->cfg.max_samples_per_call * 6;
    FMR_CUDA(h->mem.alloc(&h->d_raw, h->raw_cap, false));
  }
  if (raw_bytes > h->raw_cap) return fail(FMR_ERR_CAPACITY, "raw staging buffer too small for this call");
  if (want_sink && !h->d_levels) {
    FMR_CUDA(h->mem.alloc(&h->d_out, (size_t)C * h->audio_cap * 8, false));
    FMR_CUDA(h->mem.alloc(&h->d_levels, (size_t)C * h->cfg.max_blocks_per_call));
  }
  return FMR_OK;
}

static fmr_status check_io_args(int iq_format, const fmr_output_config *oc) {
  if (iq_format_bytes(iq_format) == 0) return fail(FMR_ERR_INVALID, "unknown iq_format");
  if (oc && out_format_bytes(oc->out_format) == 0) return fail(FMR_ERR_INVALID, "unknown out_format");
  return FMR_OK;
}

// Host-buffer entry point. The super-block is cut into a few time chunks so that the H2D copy
// of chunk k+1, the kernels of chunk k and the D2H copy of chunk k-1 overlap (three streams,
// events); with pinned host memory the call is PCIe-bound instead of copy + compute + copy.
// `fmt` is the sample format of `iq` (cf32, and int16 when an IF resampler follows, are converted inside the
// first kernel's load; the others by k_ingest_convert); `oc` != null adds the block loop's output stage, and
// then only the sink's format travels back.
static fmr_status fm_process_host_impl(fmr_fm *h, const void *iq_v, size_t iq_stride, const uint32_t *block_len,
                                       uint32_t n_blocks, void *audio_v, size_t audio_stride, uint32_t *audio_len,
                                       int fmt, const fmr_output_config *oc);

extern "C" fmr_status fmr_fm_process_host(fmr_fm *h, const float *iq, size_t iq_stride, const uint32_t *block_len,
                                          uint32_t n_blocks, double *audio, size_t audio_stride,
                                          uint32_t *audio_len) {
  return fm_process_host_impl(h, iq, iq_stride, block_len, n_blocks, audio, audio_stride, audio_len, FMR_IQ_CF32,
                              nullptr);
}

extern "C" fmr_status fmr_fm_process_host_i16(fmr_fm *h, const int16_t *iq, size_t iq_stride,
                                              const uint32_t *block_len, uint32_t n_blocks, double *audio,
                                              size_t audio_stride, uint32_t *audio_len) {
  if (h && h->cfg.input_rate == 384000.0) return fail(FMR_ERR_UNSUPPORTED, "int16 ingest needs an IF resampler stage");
  return fm_process_host_impl(h, iq, iq_stride, block_len, n_blocks, audio, audio_stride, audio_len, FMR_IQ_S16,
                              nullptr);
}

extern "C" fmr_status fmr_fm_process_host_io(fmr_fm *h, const void *iq, int iq_format, size_t iq_stride,
                                             const uint32_t *block_len, uint32_t n_blocks,
                                             const fmr_output_config *out_cfg, void *audio, size_t audio_stride,
                                             uint32_t *audio_len) {
  fmr_status s = check_io_args(iq_format, out_cfg);
  if (s != FMR_OK) return s;
  return fm_process_host_impl(h, iq, iq_stride, block_len, n_blocks, audio, audio_stride, audio_len, iq_format,
                              out_cfg);
}

static fmr_status fm_process_host_impl(fmr_fm *h, const void *iq_v, size_t iq_stride, const uint32_t *block_len,
                                       uint32_t n_blocks, void *audio_v, size_t audio_stride, uint32_t *audio_len,
                                       int fmt, const fmr_output_config *oc) {
  if (!h || !iq_v || !block_len || !audio_v) return fail(FMR_ERR_INVALID, "null argument");
  const uint8_t *iq = reinterpret_cast<const uint8_t *>(iq_v);
  uint8_t *audio = reinterpret_cast<uint8_t *>(audio_v);
  const size_t esz = (size_t)iq_format_bytes(fmt); // bytes per complex sample
  // formats the first kernel reads directly; the rest goes through k_ingest_convert
  const bool direct = (fmt == FMR_IQ_CF32) || (fmt == FMR_IQ_S16 && h->cfg.input_rate != 384000.0);
  const size_t osz = oc ? (size_t)out_format_bytes(oc->out_format) : 8;
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_blocks; b++) total += block_len[b];
  if (total > h->cfg.max_samples_per_call) return fail(FMR_ERR_CAPACITY, "sum(block_len) > max_samples_per_call");
  if (total > iq_stride) return fail(FMR_ERR_INVALID, "iq_stride < sum(block_len)");
  if (n_blocks > h->cfg.max_blocks_per_call) return fail(FMR_ERR_CAPACITY, "n_blocks > max_blocks_per_call");
  const int C = h->C;
  fmr_status s = fm_ensure_staging(h, direct ? 0 : (size_t)C * h->cfg.max_samples_per_call * esz, oc != nullptr);
  if (s != FMR_OK) return s;
  uint64_t out_total = 0;
  s = fmr_fm_query_output(h, block_len, n_blocks, &out_total, nullptr);
  if (s != FMR_OK) return s;
  if (out_total > audio_stride) return fail(FMR_ERR_CAPACITY, "audio_stride too small for this call");
  if (n_blocks == 0) return FMR_OK;
  int n_chunks = (int)(n_blocks / kHostChunkMinBlocks);
  if (n_chunks > kMaxHostChunks) n_chunks = kMaxHostChunks;
  if (n_chunks < 1) n_chunks = 1;
  cudaStream_t st = h->own_stream;
  size_t in_off = 0, out_off = 0;
  h->in_host_call = true;
  h->last_chunks.clear();
  h->last_launches = 0;
  h->last_t0 = h->cum384;
  h->last_blocks = n_blocks;
  h->have_levels = false;
  if (oc) {
    h->sink_on = true;
    h->sink.out_fmt = oc->out_format;
    h->sink.w = h->cfg.stereo ? 2 : 1;
    h->sink.gain = oc->gain;
    h->sink.squelch = oc->squelch_level;
    h->sink_out_stride = h->audio_cap;
  }
  // On any early return the copies already queued still read `iq` and write `audio`: drain them before the caller gets
  // its buffers back. (The stream position has then advanced by the chunks that completed; fmr_last_error says why.)
  struct Guard {
    fmr_fm *h;
    cudaStream_t st;
    bool drain;
    ~Guard() {
      if (drain) {
        cudaStreamSynchronize(h->s_h2d);
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(h->s_d2h);
      }
      h->in_host_call = false;
      h->iq_fmt = 0;
      h->sink_on = false;
    }
  } guard{h, st, true};
  uint8_t *d_in = direct ? reinterpret_cast<uint8_t *>(h->d_iq) : h->d_raw;
  for (int k = 0; k < n_chunks; k++) {
    const uint32_t b0 = (uint32_t)((uint64_t)n_blocks * k / n_chunks), b1 = (uint32_t)((uint64_t)n_blocks * (k + 1) / n_chunks);
    uint64_t n_in = 0;
    for (uint32_t b = b0; b < b1; b++) n_in += block_len[b];
    if (n_in > 0) {
      FMR_CUDA(cudaMemcpy2DAsync(d_in + esz * in_off, (size_t)total * esz, iq + esz * in_off, iq_stride * esz,
                                 (size_t)n_in * esz, C, cudaMemcpyHostToDevice, h->s_h2d));
    }
    FMR_CUDA(cudaEventRecord(h->ev_in[k], h->s_h2d));
    FMR_CUDA(cudaStreamWaitEvent(st, h->ev_in[k], 0));
    uint64_t n_out = 0;
    s = fmr_fm_query_output(h, block_len + b0, b1 - b0, &n_out, nullptr);
    if (s != FMR_OK) return s;
    h->host_b0 = b0;
    const float *d_chunk;
    if (direct) {
      h->iq_fmt = fmt;
      d_chunk = reinterpret_cast<const float *>(d_in + esz * in_off);
    } else {
      FMR_CUDA(launch_ingest_convert(h->d_raw + esz * in_off, fmt, (size_t)total,
                                     reinterpret_cast<float2 *>(h->d_iq) + in_off, (size_t)total, n_in, C, st));
      h->last_launches++;
      h->iq_fmt = FMR_IQ_CF32;
      d_chunk = h->d_iq + 2 * in_off;
    }
    if (oc) h->sink_out = h->d_out + out_off * osz;
    s = fm_process_device_impl(h, d_chunk, (size_t)total, block_len + b0, b1 - b0, h->d_audio + out_off, h->audio_cap,
                               audio_len ? audio_len + b0 : nullptr, (void *)st);
    if (s != FMR_OK) return s;
    FMR_CUDA(cudaEventRecord(h->ev_done[k], st));
    if (n_out > 0) {
      const uint8_t *d_res = oc ? h->d_out : reinterpret_cast<const uint8_t *>(h->d_audio);
      FMR_CUDA(cudaStreamWaitEvent(h->s_d2h, h->ev_done[k], 0));
      FMR_CUDA(cudaMemcpy2DAsync(audio + out_off * osz, audio_stride * osz, d_res + out_off * osz, h->audio_cap * osz,
                                 (size_t)n_out * osz, C, cudaMemcpyDeviceToHost, h->s_d2h));
    }
    in_off += (size_t)n_in;
    out_off += (size_t)n_out;
  }
  FMR_CUDA(cudaStreamSynchronize(st));
  FMR_CUDA(cudaStreamSynchronize(h->s_d2h));
  guard.drain = false;
  h->have_levels = (oc != nullptr);
  return FMR_OK;
}

// Device-pointer form: conversion (if the format needs its own launch) and the output stage run on `stream`.
extern "C" fmr_status fmr_fm_process_device_io(fmr_fm *h, const void *d_iq, int iq_format, size_t iq_stride,
                                               const uint32_t *block_len, uint32_t n_blocks,
                                               const fmr_output_config *out_cfg, void *d_audio, size_t audio_stride,
                                               uint32_t *audio_len, void *stream) {
  if (!h || !d_iq || !block_len || !d_audio) return fail(FMR_ERR_INVALID, "null argument");
  fmr_status s = check_io_args(iq_format, out_cfg);
  if (s != FMR_OK) return s;
  if (h->in_host_call) return fail(FMR_ERR_INVALID, "handle is inside a host call");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  const bool direct = (iq_format == FMR_IQ_CF32) || (iq_format == FMR_IQ_S16 && h->cfg.input_rate != 384000.0);
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_blocks; b++) total += block_len[b];
  if (total > h->cfg.max_samples_per_call) return fail(FMR_ERR_CAPACITY, "sum(block_len) > max_samples_per_call");
  if (total > iq_stride) return fail(FMR_ERR_INVALID, "iq_stride < sum(block_len)");
  if (!direct || out_cfg) {
    s = fm_ensure_staging(h, 0, out_cfg != nullptr);
    if (s != FMR_OK) return s;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float *d_in = reinterpret_cast<const float *>(d_iq);
  size_t stride = iq_stride;
  uint32_t extra = 0;
  h->iq_fmt = direct ? iq_format : FMR_IQ_CF32;
  if (!direct) {
    FMR_CUDA(launch_ingest_convert(d_iq, iq_format, iq_stride, reinterpret_cast<float2 *>(h->d_iq), (size_t)total, total,
                                   h->C, st));
    d_in = h->d_iq;
    stride = (size_t)total;
    extra = 1;
  }
  h->have_levels = false;
  double *d_dec = reinterpret_cast<double *>(d_audio);
  size_t dec_stride = audio_stride;
  if (out_cfg) {
    uint64_t out_total = 0;
    s = fmr_fm_query_output(h, block_len, n_blocks, &out_total, nullptr);
    if (s != FMR_OK) return s;
    if (out_total > audio_stride) return fail(FMR_ERR_CAPACITY, "audio_stride too small for this call");
    h->sink_on = true;
    h->sink.out_fmt = out_cfg->out_format;
    h->sink.w = h->cfg.stereo ? 2 : 1;
    h->sink.gain = out_cfg->gain;
    h->sink.squelch = out_cfg->squelch_level;
    h->sink_out = reinterpret_cast<uint8_t *>(d_audio);
    h->sink_out_stride = audio_stride;
    d_dec = h->d_audio;
    dec_stride = h->audio_cap;
  }
  s = fm_process_device_impl(h, d_in, stride, block_len, n_blocks, d_dec, dec_stride, audio_len, stream);
  h->sink_on = false;
  h->iq_fmt = 0;
  if (s != FMR_OK) return s;
  h->last_launches += extra;
  h->have_levels = (out_cfg != nullptr);
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_block_levels(fmr_fm *h, uint32_t channel, fmr_block_level_t *out, uint32_t n_blocks) {
  if (!h || !out || channel >= (uint32_t)h->C || n_blocks != h->last_blocks) {
    return fail(FMR_ERR_INVALID, "bad argument (n_blocks must equal the last call's)");
  }
  if (!h->have_levels) return fail(FMR_ERR_INVALID, "the last call had no output stage");
  static_assert(sizeof(BlockLevelDev) == sizeof(fmr_block_level_t), "level record layout");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  for (const auto &ch : h->last_chunks) {
    const uint32_t b0 = ch.first, nb = ch.second;
    if (nb == 0) continue;
    FMR_CUDA(cudaMemcpy(out + b0, h->d_levels + (size_t)h->C * b0 + (size_t)channel * nb, nb * sizeof(BlockLevelDev),
                        cudaMemcpyDeviceToHost));
  }
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_stats(fmr_fm *h, uint32_t channel, fmr_fm_stats_t *out) {
  if (!h || !out || channel >= (uint32_t)h->C) return fail(FMR_ERR_INVALID, "bad argument");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  FmChanState s;
  FMR_CUDA(cudaMemcpy(&s, h->d_state + channel, sizeof(s), cudaMemcpyDeviceToHost));
  out->stereo_detected = s.stereo_detected;
  out->tuning_offset = s.baseband_mean * 75000.0f; // FmDecode.h:80
  out->baseband_level = s.baseband_level;
  out->pilot_level = 2 * s.pilot_level; // PilotPhaseLock.h:66
  out->if_rms = s.if_rms;
  out->multipath_error = s.mpf_error;
  out->if_agc_gain = s.agc_gain;
  out->pll_freq = s.pll_freq;
  out->pll_phase = s.pll_phase;
  out->pll_lock_cnt = s.lock_cnt;
  out->decoder_calls = s.decoder_calls;
  out->n_pps = s.n_pps;
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_pps_events(fmr_fm *h, uint32_t channel, fmr_pps_event_t *out, uint32_t cap,
                                        uint32_t *n) {
  if (!h || !n || channel >= (uint32_t)h->C) return fail(FMR_ERR_INVALID, "bad argument");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  FmChanState s;
  FMR_CUDA(cudaMemcpy(&s, h->d_state + channel, sizeof(s), cudaMemcpyDeviceToHost));
  PpsEventDev ev[kMaxPps];
  FMR_CUDA(cudaMemcpy(ev, h->d_pps + (size_t)channel * kMaxPps, sizeof(ev), cudaMemcpyDeviceToHost));
  uint32_t cnt = s.n_pps < (uint32_t)kMaxPps ? s.n_pps : (uint32_t)kMaxPps;
  *n = cnt;
  for (uint32_t i = 0; i < cnt && i < cap && out; i++) {
    out[i].pps_index = ev[i].pps_index;
    out[i].sample_index = ev[i].sample_index;
    out[i].block_position = ev[i].block_position;
    out[i].block = ev[i].block;
  }
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_coeffs(fmr_fm *h, uint32_t channel, float *re_im, size_t n_complex) {
  if (!h || !re_im || channel >= (uint32_t)h->C) return fail(FMR_ERR_INVALID, "bad argument");
  if (h->cfg.multipath_stages == 0) return fail(FMR_ERR_INVALID, "multipath filter is disabled");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  return h->mpf.read_coeffs(channel, re_im, n_complex);
}

extern "C" fmr_status fmr_fm_block_flags(fmr_fm *h, uint32_t channel, uint8_t *stereo, uint32_t n_blocks) {
  if (!h || !stereo || channel >= (uint32_t)h->C || n_blocks != h->last_blocks) {
    return fail(FMR_ERR_INVALID, "bad argument (n_blocks must equal the last call's)");
  }
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  for (const auto &ch : h->last_chunks) {
    const uint32_t b0 = ch.first, nb = ch.second;
    if (nb == 0) continue;
    FMR_CUDA(cudaMemcpy(stereo + b0, h->d_flags + (size_t)h->C * b0 + (size_t)channel * nb, nb,
                        cudaMemcpyDeviceToHost));
  }
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_tap_if(fmr_fm *h, uint32_t channel, float *re_im, size_t cap_complex,
                                    uint64_t *n_complex) {
  if (!h || !n_complex || channel >= (uint32_t)h->C) return fail(FMR_ERR_INVALID, "bad argument");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  FMR_CUDA(cudaDeviceSynchronize());
  const int64_t n = h->last_t1 - h->last_t0;
  *n_complex = (uint64_t)n;
  if (!re_im) return FMR_OK;
  const uint32_t cap = h->r_if.cap;
  for (int64_t i = 0; i < n && (size_t)i < cap_complex;) {
    const uint32_t pos = (uint32_t)((h->last_t0 + i) & (cap - 1));
    int64_t run = cap - pos;
    if (run > n - i) run = n - i;
    if ((size_t)(i + run) > cap_complex) run = (int64_t)cap_complex - i;
    FMR_CUDA(cudaMemcpy(re_im + 2 * i, h->r_if.base + (size_t)channel * cap + pos, (size_t)run * 8,
                        cudaMemcpyDeviceToHost));
    i += run;
  }
  return FMR_OK;
}

extern "C" uint32_t fmr_fm_last_launches(fmr_fm *h) { return h ? h->last_launches : 0; }

extern "C" size_t fmr_fm_describe(fmr_fm *h, char *buf, size_t cap) {
  if (!h || !buf || cap == 0) return 0;
  const Resampler<float> &r = h->ifres;
  const int n = snprintf(buf, cap,
                         "input_rate=%.0f channels=%d if_chain=%s frontend=%s lowpass=%s halfbands=%s core=%s audio_lowpass=%s "
                         "multipath_stages=%u fmfilter=%d",
                         h->cfg.input_rate, h->C, h->ifc ? "r8brain-tables" : "none",
                         r.use_fe ? (r.fe_variant == 1 ? "fused(split-lanes)" : r.fe_variant == 2 ? "fused(8-consumer-warps,setmaxnreg)" : "fused") : "unfused",
                         r.use_fdr ? (r.fdr_rl == 12 ? "fdr(3072)" : r.fdr_rl == 15 ? "fdr(3840)" : "fdr(2560)")
                                   : (r.use_fft ? (r.fft_inplace ? "fft16384-inplace+bank" : "fft16384-stockham+bank") : "direct"),
                         r.hb_stream ? (r.hbs_tma ? "stream(tma)" : "stream(cp.async)") : "tiled",
                         (h->core_fused && h->cfg.stereo && h->cfg.multipath_stages == 0) ? "fused" : "agc|disc|pll",
                         h->audio_f64 ? (h->a64.aures.use_fft ? "fft8192-f64" : "direct-f64") : (h->a32.aures.use_fft ? "fft-f32" : "direct-f32"), h->cfg.multipath_stages, h->cfg.fmfilter);
  return (n < 0) ? 0 : ((size_t)n < cap ? (size_t)n : cap - 1);
}

extern "C" fmr_status fmr_fm_last_plan(fmr_fm *h, uint64_t plan[5]) {
  if (!h || !plan) return fail(FMR_ERR_INVALID, "null argument");
  for (int i = 0; i < 5; i++) plan[i] = h->ifc ? h->ifres.last_plan[i] : 0;
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_set_profiling(fmr_fm *h, int enable) {
  if (!h) return fail(FMR_ERR_INVALID, "null handle");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  if (enable) {
    h->prof.enable();
  } else {
    h->prof.release();
  }
  return FMR_OK;
}

extern "C" fmr_status fmr_fm_stage_times(fmr_fm *h, float *ms, const char **names, uint32_t cap, uint32_t *n) {
  if (!h || !ms || !names || !n) return fail(FMR_ERR_INVALID, "null argument");
  FMR_CUDA(cudaSetDevice(h->cfg.device));
  return h->prof.read(ms, names, cap, n);
}
