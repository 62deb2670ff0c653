/*
 * fmradion_b200.h — C ABI of the B200-native demodulation hot path.
 *
 * This is the drop-in boundary for the block API the reference exposes as C++ classes:
 *
 *   FmDecoder::FmDecoder(...)                     include/FmDecode.h:63-64
 *   void FmDecoder::process(IQSampleVector, SampleVector&)          include/FmDecode.h:74
 *   stereo_detected / get_tuning_offset / get_baseband_level /
 *   get_pilot_level / get_if_rms                  include/FmDecode.h:77-89
 *   get_pps_events / erase_first_pps_event        include/FmDecode.h:92-97
 *   get_multipath_error / get_multipath_coefficients   include/FmDecode.h:100-105
 *   AmDecoder::AmDecoder(...), process, getters   include/AmDecode.h:48-65
 *
 * plus the two front-end objects the reference's block loop runs immediately before the
 * decoder and which the GPU path absorbs because that is where the bytes are:
 *
 *   FourthConverterIQ::process                    include/FourthConverterIQ.h:38-82  (main.cpp:912-919)
 *   IfResampler::process                          sfmbase/IfResampler.cpp:37-79      (main.cpp:921-926)
 *
 * One handle decodes `n_channels` independent IQ streams in lock step (same block lengths
 * for every channel). A call hands over `n_blocks` consecutive source blocks per channel
 * (a "super-block"); `block_len[]` carries the reference's block partition so that every
 * behaviour that depends on where the reference's process() calls begin and end (stereo
 * switch-over, per-call statistics, the FIR head-loop quirk, LMS update phase, multipath
 * warm-up count) is reproduced inside one launch sequence.
 *
 * Conventions: plain pointers and sizes, integer status codes, no exceptions cross the ABI,
 * caller-owned buffers, one caller thread per handle. There is NO CPU fallback: every entry
 * point that computes fails with FMR_ERR_CUDA if the device or kernels are unavailable.
 */
#ifndef FMRADION_B200_H
#define FMRADION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int fmr_status;
enum {
  FMR_OK = 0,
  FMR_ERR_INVALID = 1,     /* bad argument */
  FMR_ERR_UNSUPPORTED = 2, /* rate pair or option outside the shipped tables */
  FMR_ERR_CUDA = 3,        /* CUDA runtime error; see fmr_last_error() */
  FMR_ERR_CAPACITY = 4     /* output buffer or max_samples_per_call too small */
};

/* Human-readable description of the last error on this thread. */
const char *fmr_last_error(void);
/* Library version string and number of multiprocessors of `device` (<=0 on failure). */
const char *fmr_version(void);
int fmr_device_sm_count(int device);

/* ------------------------------------------------------------------ FM broadcast ---- */

typedef struct fmr_fm fmr_fm;

typedef struct fmr_fm_config {
  double input_rate;          /* IQ sample rate in Hz: 384000, 1e6, 2.5e6, 6e6, 1e7      */
  int fs4_shift;              /* 1: FourthConverterIQ(false) first (zero-IF sources)      */
  int fmfilter;               /* 0 none (default/wide), 1 medium, 2 narrow (main.cpp:785),
                                 3 = fmfilter_coeff[fmfilter_ntaps] (FmDecoder ctor `fmfilter_coeff`) */
  int stereo;                 /* FmDecoder ctor `stereo`                                  */
  double deemphasis_us;       /* 50 (EU/JP), 75 (NA), 0 = off                             */
  int pilot_shift;            /* FmDecoder ctor `pilot_shift` (-X)                        */
  uint32_t multipath_stages;  /* FmDecoder ctor `multipath_stages` (-E), 0 = off          */
  uint32_t n_channels;        /* independent streams decoded by this handle               */
  uint32_t max_samples_per_call; /* largest sum(block_len) a process call may carry       */
  uint32_t max_blocks_per_call;  /* largest n_blocks                                       */
  int device;                 /* CUDA device ordinal                                      */
  const float *fmfilter_coeff;   /* used when fmfilter == 3: symmetric FIR taps (Filter.cpp:27-35) */
  uint32_t fmfilter_ntaps;
} fmr_fm_config;

typedef struct fmr_fm_stats_t {
  int stereo_detected;      /* FmDecoder::stereo_detected()      */
  float tuning_offset;      /* FmDecoder::get_tuning_offset()    */
  float baseband_level;     /* FmDecoder::get_baseband_level()   */
  double pilot_level;       /* FmDecoder::get_pilot_level()      */
  float if_rms;             /* FmDecoder::get_if_rms()           */
  double multipath_error;   /* FmDecoder::get_multipath_error()  */
  float if_agc_gain;        /* IfSimpleAgc::get_current_gain()   */
  double pll_freq;          /* PilotPhaseLock m_freq (rad/sample) */
  double pll_phase;         /* PilotPhaseLock m_phase            */
  int pll_lock_cnt;         /* PilotPhaseLock m_lock_cnt         */
  uint64_t decoder_calls;   /* non-empty FmDecoder::process calls so far */
  uint32_t n_pps;           /* PPS events recorded during the last process call */
} fmr_fm_stats_t;

typedef struct fmr_pps_event_t { /* PilotPhaseLock::PpsEvent + the block it fell in */
  uint64_t pps_index;
  uint64_t sample_index;
  double block_position;
  uint32_t block; /* index into block_len[] of the last process call */
} fmr_pps_event_t;

/* Replaces the FmDecoder constructor (include/FmDecode.h:49-64) for cfg->n_channels independent streams.
 * Arithmetic: every recurrence (PLL, deemphasis, DC block) in double like the reference; the linear filters (IF
 * resampler, audio resamplers, pilot cut) in float. With FMR_AUDIO_FP64=1 in the environment at creation the audio
 * resamplers and the pilot cut run in double as well (audio error against the reference 2e-7 instead of 1.5e-6 of
 * full scale, about 7 % lower throughput). */
fmr_status fmr_fm_create(const fmr_fm_config *cfg, fmr_fm **out);
void fmr_fm_destroy(fmr_fm *h);

/*
 * Decode n_blocks source blocks for every channel.
 *   iq            : interleaved (re,im) float pairs, channel-major:
 *                   channel c occupies iq[2*c*iq_stride .. 2*c*iq_stride + 2*sum(block_len));
 *                   iq_stride is in complex samples (>= sum(block_len)).
 *   audio         : channel-major, channel c at audio[c*audio_stride ..]; stereo handles write
 *                   interleaved L,R doubles, mono handles one double per frame (FmDecode.h:66-74).
 *   audio_stride  : capacity per channel in doubles.
 *   audio_len     : [n_blocks] doubles produced by each block (identical for all channels;
 *                   0 while the resamplers fill, exactly as the reference's per-call sizes).
 * The *_host form takes pageable or pinned host pointers and does the copies itself; the
 * *_device form takes device pointers and enqueues on `stream` (a cudaStream_t, 0 = default)
 * without synchronising; audio_len is always a host pointer and is filled before return.
 */
fmr_status fmr_fm_process_host(fmr_fm *h, const float *iq, size_t iq_stride,
                               const uint32_t *block_len, uint32_t n_blocks,
                               double *audio, size_t audio_stride, uint32_t *audio_len);
fmr_status fmr_fm_process_device(fmr_fm *h, const float *d_iq, size_t iq_stride,
                                 const uint32_t *block_len, uint32_t n_blocks,
                                 double *d_audio, size_t audio_stride, uint32_t *audio_len,
                                 void *stream);

/* Same, for IQ delivered as interleaved int16 (re,im) pairs, value/32768 — what FileSource's
 * sf_read_float yields for 16-bit PCM files (sfmbase/FileSource.cpp:491-531). The conversion is
 * fused into the first kernel's load, so only half the bytes cross PCIe/HBM. iq_stride in complex
 * samples. Needs input_rate != 384000 (the conversion lives in the IF resampler's first stage). */
fmr_status fmr_fm_process_host_i16(fmr_fm *h, const int16_t *iq, size_t iq_stride,
                                   const uint32_t *block_len, uint32_t n_blocks,
                                   double *audio, size_t audio_stride, uint32_t *audio_len);
fmr_status fmr_fm_process_device_i16(fmr_fm *h, const int16_t *d_iq, size_t iq_stride,
                                     const uint32_t *block_len, uint32_t n_blocks,
                                     double *d_audio, size_t audio_stride, uint32_t *audio_len,
                                     void *stream);

/* Audio doubles per channel the next process call with these block lengths will produce
 * (lets the caller size `audio`); does not advance the stream. */
fmr_status fmr_fm_query_output(fmr_fm *h, const uint32_t *block_len, uint32_t n_blocks,
                               uint64_t *audio_doubles_total, uint32_t *audio_len);

/* Pure host bookkeeping, no device needed: the IF (384 kHz) and audio (doubles) lengths the
 * reference produces per block for a stream that has already consumed start_sample samples.
 * Same integer model fmr_fm_process_* uses (r8brain's release schedule, see csrc/fmr_tables.h). */
fmr_status fmr_fm_schedule(double input_rate, int stereo, uint64_t start_sample, const uint32_t *block_len,
                           uint32_t n_blocks, uint32_t *if_len, uint32_t *audio_len);

fmr_status fmr_fm_stats(fmr_fm *h, uint32_t channel, fmr_fm_stats_t *out);
/* Copies up to cap events of the last process call; returns the count in *n. */
fmr_status fmr_fm_pps_events(fmr_fm *h, uint32_t channel, fmr_pps_event_t *out, uint32_t cap,
                             uint32_t *n);
/* Multipath filter coefficients as (re,im) float pairs; n_complex = 4*stages+1. */
fmr_status fmr_fm_coeffs(fmr_fm *h, uint32_t channel, float *re_im, size_t n_complex);
/* Per-block stereo flags of the last process call, channel-major [n_blocks]. */
fmr_status fmr_fm_block_flags(fmr_fm *h, uint32_t channel, uint8_t *stereo, uint32_t n_blocks);
/* Debug/parity tap: the 384 kHz IF samples (decoder input) of the last process call. */
fmr_status fmr_fm_tap_if(fmr_fm *h, uint32_t channel, float *re_im, size_t cap_complex,
                         uint64_t *n_complex);
/* Kernel launches issued by the last process call (for bench accounting). */
uint32_t fmr_fm_last_launches(fmr_fm *h);
/* How the last process call's IF front end was executed, per channel (for bench accounting of each kernel's own
 * bytes): plan[0] = blocks of the frequency-domain resampler's grid taken by the fused front-end kernel, plan[1] = blocks
 * taken by the unfused kernel, plan[2] = 1.25 MHz samples the unfused half-band kernels produced, plan[3] = input
 * samples per block (60000 at 10 Msps), plan[4] = output samples per block (2304). All zero for chains without it. */
fmr_status fmr_fm_last_plan(fmr_fm *h, uint64_t plan[5]);
/* Which implementation of every stage this handle selected at creation (environment switches, sample rate), as one line
 * of "key=value" words, so that a run can be audited. Returns the length written (without the terminating 0). */
size_t fmr_fm_describe(fmr_fm *h, char *buf, size_t cap);
/* Per-stage device timing (CUDA events on the launching stream). Enable, run a process
 * call, synchronise, then read: ms[i] is the duration of stage names[i] in the last call. */
fmr_status fmr_fm_set_profiling(fmr_fm *h, int enable);
fmr_status fmr_fm_stage_times(fmr_fm *h, float *ms, const char **names, uint32_t cap, uint32_t *n);

/* ------------------------------------------- file formats in, sink formats out ---- */
/* IQ sample formats FileSource accepts (sfmbase/FileSource.cpp:120-138,206-216), as sf_read_float
 * delivers them (FileSource.cpp:491-531): integer value / 2^(bits-1), U8 as (value-128)/128. The decode
 * runs on the device, so only the file's own bytes cross PCIe (S16: 4, S8/U8: 2, S24: 6 instead of 8). */
enum { FMR_IQ_CF32 = 0, FMR_IQ_S16 = 1, FMR_IQ_S8 = 2, FMR_IQ_U8 = 3, FMR_IQ_S24 = 4 };
/* Audio sample formats of the reference's SndfileOutput (main.cpp:592-623, AudioOutput.cpp:153-167):
 * F64 = the decoder's own doubles, F32 = (float)x, S16 = lrint(x * 32767) (sf_write_double, no clipping). */
enum { FMR_OUT_F64 = 0, FMR_OUT_F32 = 1, FMR_OUT_S16 = 2 };

/* The block loop's output stage, main.cpp:977-1002: per-block audio level (Utility::samples_mean_rms on the
 * float copy), Utility::adjust_gain(audio, if_rms >= squelch_level ? gain : 0) and the sink's sample format. */
typedef struct fmr_output_config {
  int out_format;        /* FMR_OUT_*                                                   */
  double squelch_level;  /* linear IF level, main.cpp:484-489 (0 = squelch always open) */
  double gain;           /* 0.5 = the reference's nominal -6 dB (main.cpp:1000)         */
} fmr_output_config;

typedef struct fmr_block_level_t { /* what the block loop derives per source block */
  float if_rms;      /* decoder's get_if_rms() after this block's process call (main.cpp:956-976);
                        -1 when the block produced no IF samples (the loop `continue`s, main.cpp:933-936) */
  float audio_mean;  /* Utility::samples_mean_rms of the block's audio as float (main.cpp:989-996) */
  float audio_rms;
  float gain;        /* gain applied: cfg.gain or 0 (squelch closed)                */
} fmr_block_level_t;

/* process_host / process_device with the input in `iq_format` and, when out_cfg != NULL, the output stage
 * applied on the device: `audio` then receives values of out_cfg->out_format (audio_stride and audio_len
 * count values, not bytes). out_cfg == NULL: `audio` receives the decoder's doubles unchanged. */
fmr_status fmr_fm_process_host_io(fmr_fm *h, const void *iq, int iq_format, size_t iq_stride,
                                  const uint32_t *block_len, uint32_t n_blocks, const fmr_output_config *out_cfg,
                                  void *audio, size_t audio_stride, uint32_t *audio_len);
fmr_status fmr_fm_process_device_io(fmr_fm *h, const void *d_iq, int iq_format, size_t iq_stride,
                                    const uint32_t *block_len, uint32_t n_blocks, const fmr_output_config *out_cfg,
                                    void *d_audio, size_t audio_stride, uint32_t *audio_len, void *stream);
/* Per-block levels of the last *_io call that had an out_cfg; n_blocks must equal that call's. */
fmr_status fmr_fm_block_levels(fmr_fm *h, uint32_t channel, fmr_block_level_t *out, uint32_t n_blocks);

/* ------------------------------------------------------ AM and narrow-band FM ------- */
/* One handle type serves the two 48 kHz decoders of the reference:
 *   mode 2 (ModType::AM)   AmDecoder::process    include/AmDecode.h:48-65, sfmbase/AmDecode.cpp:96-218
 *   mode 1 (ModType::NBFM) NbfmDecoder::process  include/NbfmDecode.h:43-63, sfmbase/NbfmDecode.cpp:47-96
 * both behind FourthConverterIQ + IfResampler(input_rate -> 48 kHz) exactly as main.cpp:912-971 runs them. */

typedef struct fmr_am fmr_am;

typedef struct fmr_am_config {
  double input_rate;    /* 48000 or 384000 */
  int fs4_shift;
  int amfilter;         /* 0 default, 1 medium, 2 narrow, 3 wide (main.cpp:785-810),
                           4 = amfilter_coeff[amfilter_ntaps] (AmDecoder ctor `amfilter_coeff`) */
  int mode;             /* ModType value (include/SoftFM.h:49): 1 NBFM, 2 AM, 3 DSB, 4 USB, 5 LSB, 6 CW, 7 WSPR
                           (AmDecoder::process, AmDecode.cpp:96-218; FineTuner.cpp:55-70 for the pitch shifts).
                           For NBFM `amfilter` selects jj1bdx_nbfm_48khz_{default,medium,narrow,wide}
                           (main.cpp:785-810) or, with 4, the caller's `amfilter_coeff` (ctor `nbfmfilter_coeff`) */
  uint32_t n_channels;
  uint32_t max_samples_per_call;
  uint32_t max_blocks_per_call;
  int device;
  const float *amfilter_coeff; /* used when amfilter == 4 */
  uint32_t amfilter_ntaps;
  double nbfm_freq_dev; /* NBFM: full-scale deviation in Hz (NbfmDecoder ctor `freq_dev`); 0 = freq_dev_normal = 8000 */
} fmr_am_config;

typedef struct fmr_am_stats_t {
  double baseband_level;  /* AmDecoder::get_baseband_level()        */
  float af_agc_gain;      /* AmDecoder::get_af_agc_current_gain()   */
  float if_agc_gain;      /* AmDecoder::get_if_agc_current_gain()   */
  float if_rms;           /* AmDecoder::get_if_rms() / NbfmDecoder::get_if_rms() */
  uint64_t decoder_calls;
  float tuning_offset;    /* NbfmDecoder::get_tuning_offset() (0 for AM) */
} fmr_am_stats_t;

fmr_status fmr_am_create(const fmr_am_config *cfg, fmr_am **out);
void fmr_am_destroy(fmr_am *h);
fmr_status fmr_am_process_host(fmr_am *h, const float *iq, size_t iq_stride,
                               const uint32_t *block_len, uint32_t n_blocks,
                               double *audio, size_t audio_stride, uint32_t *audio_len);
fmr_status fmr_am_process_device(fmr_am *h, const float *d_iq, size_t iq_stride,
                                 const uint32_t *block_len, uint32_t n_blocks,
                                 double *d_audio, size_t audio_stride, uint32_t *audio_len,
                                 void *stream);
fmr_status fmr_am_query_output(fmr_am *h, const uint32_t *block_len, uint32_t n_blocks,
                               uint64_t *audio_doubles_total, uint32_t *audio_len);
fmr_status fmr_am_schedule(double input_rate, uint64_t start_sample, const uint32_t *block_len,
                           uint32_t n_blocks, uint32_t *audio_len);
fmr_status fmr_am_stats(fmr_am *h, uint32_t channel, fmr_am_stats_t *out);
uint32_t fmr_am_last_launches(fmr_am *h);
fmr_status fmr_am_set_profiling(fmr_am *h, int enable);
fmr_status fmr_am_stage_times(fmr_am *h, float *ms, const char **names, uint32_t cap, uint32_t *n);

/* Same as the fmr_fm_*_io entry points, for the 48 kHz decoders. */
fmr_status fmr_am_process_host_io(fmr_am *h, const void *iq, int iq_format, size_t iq_stride,
                                  const uint32_t *block_len, uint32_t n_blocks, const fmr_output_config *out_cfg,
                                  void *audio, size_t audio_stride, uint32_t *audio_len);
fmr_status fmr_am_process_device_io(fmr_am *h, const void *d_iq, int iq_format, size_t iq_stride,
                                    const uint32_t *block_len, uint32_t n_blocks, const fmr_output_config *out_cfg,
                                    void *d_audio, size_t audio_stride, uint32_t *audio_len, void *stream);
fmr_status fmr_am_block_levels(fmr_am *h, uint32_t channel, fmr_block_level_t *out, uint32_t n_blocks);

#ifdef __cplusplus
}
#endif
#endif /* FMRADION_B200_H */
