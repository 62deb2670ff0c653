// fmr_io.cuh — the two stages either side of the decoder, on the device so
// that only the file's own bytes cross PCIe on the way in and only the sink's format on the way out.
//
//   ingest : what FileSource::get_sf_read_float hands to the block loop (sfmbase/FileSource.cpp:491-531):
//            libsndfile's sf_read_float with its default normalisation, for the sub-types FileSource
//            accepts (FileSource.cpp:206-216): PCM_S8, PCM_16, PCM_24, PCM_U8, FLOAT. libsndfile is an
//            un-vendored dependency; the conversions restated here are its pcm.c readers
//            (sc2f_array, uc2f_array, les2f_array, let2f_array): integer value times a power of two,
//            exact in float.
//   output : main.cpp:989-1002 — audio level (double -> float, Utility::samples_mean_rms,
//            Utility.h:135-152), Utility::adjust_gain with 0.5 or 0 from the IF squelch (Utility.h:307-312),
//            then SndfileOutput::write = sf_write_double into the sink's sample format
//            (AudioOutput.cpp:153-167; PCM_16: lrint(x * 32767) without clipping, FLOAT: (float)x — libsndfile
//            pcm.c d2s_array / float32.c d2f_array with norm_double on, clipping off, its defaults).
//            The per-block IF RMS (FmDecoder::get_if_rms of that call, main.cpp:956-976) is recomputed
//            here from the decoder-input ring with a parallel reduction; like every VOLK reduction on this
//            path its summation order is not pinned by the reference (statistics only).
#ifndef FMR_IO_CUH
#define FMR_IO_CUH

#include "fmr_kernels.cuh"

namespace fmr {

__host__ __device__ inline int iq_format_bytes(int fmt) { // bytes per complex sample
  switch (fmt) {
  case FMR_IQ_CF32: return 8;
  case FMR_IQ_S16: return 4;
  case FMR_IQ_S8: return 2;
  case FMR_IQ_U8: return 2;
  case FMR_IQ_S24: return 6;
  default: return 0;
  }
}
__host__ __device__ inline int out_format_bytes(int fmt) { // bytes per audio value
  switch (fmt) {
  case FMR_OUT_F64: return 8;
  case FMR_OUT_F32: return 4;
  case FMR_OUT_S16: return 2;
  default: return 0;
  }
}

// raw file samples -> cf32, channel-major on both sides. Thread = one complex sample; a warp reads
// 64..192 consecutive bytes and writes 256.
static __global__ void __launch_bounds__(256)
    k_ingest_convert(const uint8_t *__restrict__ raw, int fmt, size_t src_stride, float2 *__restrict__ dst,
                     size_t dst_stride, uint32_t n) {
  const uint32_t c = blockIdx.y;
  const int esz = iq_format_bytes(fmt);
  const uint8_t *row = raw + (size_t)c * src_stride * (size_t)esz;
  float2 *orow = dst + (size_t)c * dst_stride;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float2 v;
    if (fmt == FMR_IQ_S16) {
      const short2 q = reinterpret_cast<const short2 *>(row)[i];
      v.x = (float)q.x * (1.0f / 32768.0f);
      v.y = (float)q.y * (1.0f / 32768.0f);
    } else if (fmt == FMR_IQ_S8) {
      const char2 q = reinterpret_cast<const char2 *>(row)[i];
      v.x = (float)q.x * (1.0f / 128.0f);
      v.y = (float)q.y * (1.0f / 128.0f);
    } else if (fmt == FMR_IQ_U8) {
      const uchar2 q = reinterpret_cast<const uchar2 *>(row)[i];
      v.x = (float)((int)q.x - 128) * (1.0f / 128.0f);
      v.y = (float)((int)q.y - 128) * (1.0f / 128.0f);
    } else if (fmt == FMR_IQ_S24) {
      const uint8_t *p = row + (size_t)i * 6;
      const int re = (int)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24));
      const int im = (int)(((uint32_t)p[3] << 8) | ((uint32_t)p[4] << 16) | ((uint32_t)p[5] << 24));
      v.x = (float)re * (1.0f / 2147483648.0f);
      v.y = (float)im * (1.0f / 2147483648.0f);
    } else {
      v = reinterpret_cast<const float2 *>(row)[i];
    }
    orow[i] = v;
  }
}

inline cudaError_t launch_ingest_convert(const void *d_raw, int fmt, size_t src_stride, float2 *dst, size_t dst_stride,
                                         uint64_t n, int C, cudaStream_t st) {
  if (n == 0 || C == 0) return cudaSuccess;
  unsigned gx = (unsigned)((n + 1023) / 1024);
  if (gx < 1) gx = 1;
  dim3 grid(gx, (unsigned)C);
  k_ingest_convert<<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t *>(d_raw), fmt, src_stride, dst, dst_stride,
                                         (uint32_t)n);
  return cudaGetLastError();
}

struct SinkParams {
  int out_fmt;     // FMR_OUT_*
  int w;           // doubles per audio frame (2 = interleaved stereo)
  double gain;     // 0.5 in main.cpp:1000
  double squelch;  // linear IF level below which the block is muted (main.cpp:484-489)
};

struct BlockLevelDev { // one per (chunk, channel, block)
  float if_rms, audio_mean, audio_rms, gain;
};

constexpr int kSinkThreads = 128;

__device__ __forceinline__ float sink_block_sum(float v, float *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < kSinkThreads / 32; k++) t += sh[k];
  return t;
}

// CTA = one source block of one channel. `ifr` is the decoder-input ring (384 kHz for FM, 48 kHz for
// AM / NBFM), e_if / e_au the call-relative cumulative ends per block (complex samples, audio frames).
static __global__ void __launch_bounds__(kSinkThreads)
    k_audio_sink(Ring<float2> ifr, int64_t t0, const uint32_t *__restrict__ e_if, const double *__restrict__ audio,
                 size_t audio_stride, const uint32_t *__restrict__ e_au, int nb, void *__restrict__ out,
                 size_t out_stride, BlockLevelDev *__restrict__ levels, SinkParams P) {
  __shared__ float sh[kSinkThreads / 32];
  const int b = blockIdx.x;
  const uint32_t c = blockIdx.y;
  const uint32_t i0 = b ? e_if[b - 1] : 0u, i1 = e_if[b];
  const uint32_t a0 = (b ? e_au[b - 1] : 0u) * (uint32_t)P.w, a1 = e_au[b] * (uint32_t)P.w;
  // ---- IF RMS of this decoder call (Utility::rms_level_sample, Utility.h:118-132)
  float sq = 0.f;
  for (uint32_t i = i0 + threadIdx.x; i < i1; i += kSinkThreads) {
    const float2 x = ifr.ld(c, t0 + (int64_t)i);
    sq += x.x * x.x + x.y * x.y;
  }
  sq = sink_block_sum(sq, sh);
  const uint32_t n_if = i1 - i0;
  // no IF samples: the block loop `continue`s before the decoder (main.cpp:933-936); marked with -1
  const float if_rms = n_if ? sqrtf(sq / (float)n_if) : -1.0f;
  // ---- audio level on the float copy (main.cpp:989-996)
  const double *arow = audio + (size_t)c * audio_stride;
  float vs = 0.f, vq = 0.f;
  for (uint32_t i = a0 + threadIdx.x; i < a1; i += kSinkThreads) {
    const float f = (float)arow[i];
    vs += f;
    vq += f * f;
  }
  vs = sink_block_sum(vs, sh);
  vq = sink_block_sum(vq, sh);
  const uint32_t n_au = a1 - a0;
  const float mean = n_au ? vs / (float)n_au : 0.f;
  const float rms = n_au ? sqrtf(vq / (float)n_au) : 0.f;
  // ---- squelch + nominal volume (main.cpp:998-1000), then the sink's sample format
  const double g = ((double)if_rms >= P.squelch) ? P.gain : 0.0;
  if (threadIdx.x == 0) {
    BlockLevelDev l;
    l.if_rms = if_rms;
    l.audio_mean = mean;
    l.audio_rms = rms;
    l.gain = (float)g;
    levels[(size_t)c * nb + b] = l;
  }
  for (uint32_t i = a0 + threadIdx.x; i < a1; i += kSinkThreads) {
    const double y = arow[i] * g; // Utility::adjust_gain (Utility.h:307-312)
    const size_t o = (size_t)c * out_stride + i;
    if (P.out_fmt == FMR_OUT_F64) {
      reinterpret_cast<double *>(out)[o] = y;
    } else if (P.out_fmt == FMR_OUT_F32) {
      reinterpret_cast<float *>(out)[o] = (float)y;
    } else {
      reinterpret_cast<short *>(out)[o] = (short)__double2ll_rn(y * 32767.0);
    }
  }
}

} // namespace fmr
#endif
