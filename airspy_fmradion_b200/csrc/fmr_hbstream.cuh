// fmr_hbstream.cuh — streaming form of the 3-stage half-band decimator cascade (reference:
// r8b::CDSPHBDownsampler::process, CDSPHBDownsampler.h:137-239; the three third-band stages
// CDSPResampler builds in front of the block convolver for 10 MHz -> 1.25 MHz,
// CDSPResampler.h:372-381).
//
// The tiled kernel (k_hb_cascade) keeps every intermediate rate of a tile in shared memory and
// is bound by shared-memory wavefronts and index arithmetic (~60 instructions per input
// sample). This form gives every THREAD its own contiguous time range ("stream tile") of one
// channel and runs the cascade as a streaming filter whose delay lines live entirely in
// registers: per 16 input samples a thread produces 8 + 4 + 2 outputs of the three stages with
// compile-time register indices, no index arithmetic, no barriers.
//
// Staging: a warp's 32 streams are 32 different address ranges, so a per-lane load would touch
// 32 lines per instruction. Instead the warp copies one 128-byte chunk per stream with
// cp.async (LDGSTS.128) in a transposed assignment — 8 lanes cover one stream's line, 8
// instructions cover the 32 streams — into a per-warp ring of shared-memory stages, and every
// lane then reads its own row with conflict-free LDS.128 (row pitch 144 B). The ring keeps
// kHbsStages - 1 chunks per stream in flight; only __syncwarp() is needed.
//
// Arithmetic is the same expression, in the same order, as hb_stage() in fmr_kernels.cuh
// (y = x[2m]; y += t[k]*(x[2m+2k+1] + x[2m-2k-1]), k ascending, FADD2 + FFMA2), so both kernels
// produce bit-identical streams.
#ifndef FMR_HBSTREAM_CUH
#define FMR_HBSTREAM_CUH

#if defined(__CUDACC__)
#define FMR_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define FMR_HD inline
struct float2 {
  float x, y;
};
#endif

namespace fmr {

FMR_HD float2 hbs_acc(float2 y, float t, float2 u, float2 v) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(make_float2(t, t), __fadd2_rn(u, v), y);
#else
  float2 r;
  r.x = fmaf(t, u.x + v.x, y.x);
  r.y = fmaf(t, u.y + v.y, y.y);
  return r;
#endif
}

// real-valued form (one component of the complex stream per thread): the same two roundings per lane as FADD2 / FFMA2
FMR_HD float hbs_acc(float y, float t, float u, float v) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(t, __fadd_rn(u, v), y);
#else
  return fmaf(t, u + v, y);
#endif
}
FMR_HD void hbs_fill(float2 &d, float v) { d.x = d.y = v; }
FMR_HD void hbs_fill(float &d, float v) { d = v; }

constexpr int hbs_even_ceil(int v) { return (v + 1) & ~1; }

// Delay bookkeeping of the streaming cascade. Macro-step i consumes input samples
// [16i, 16i+16) and produces stage-s outputs [ (16>>s) i - A_s, (16>>s) i - A_s + (16>>s) ).
template <int N1, int N2, int N3> struct HbsDelays {
  static constexpr int A1 = hbs_even_ceil(N1 - 1);
  static constexpr int A2 = hbs_even_ceil(N2 - 1 + A1 / 2);
  static constexpr int A3 = hbs_even_ceil(N3 - 1 + A2 / 2);
  static constexpr int CE1 = A1, CE2 = A2 - A1 / 2, CE3 = A3 - A2 / 2; // carried even-phase samples
  // First macro-step a stream has to run so that the stage-3 outputs of macro-step `i_out`
  // are exact: walk the oldest window sample of every stage back to the step that made it.
  static constexpr long long first_step(long long i_out) {
    const long long o2 = 2 * i_out - A3 - N3; // oldest O2 index read by stage 3
    const long long m2 = 2 * o2 + 1;          // = x2 index
    const long long i2 = (m2 + A2) >> 2;      // step that produced x2[m2] (floor)
    const long long o1 = 4 * i2 - A2 - N2;    // oldest O1 index read by stage 2 at i2
    const long long m1 = 2 * o1 + 1;
    const long long i1 = (m1 + A1) >> 3;      // step that produced x1[m1]
    return i1 - 1;                            // stage 1 at i1 also reads the samples of step i1-1
  }
  static constexpr int warm_steps() {
    int w = 0;
    for (long long i = 4096; i < 4096 + 8; i++) {
      const int d = (int)(i - first_step(i));
      if (d > w) w = d;
    }
    return w;
  }
  // macro-steps a stream runs before the one that emits its first wanted output
  static constexpr int kWarm = warm_steps();
};

// One stage: NEW outputs per call. wE[q] = E[m0 + q], wO[q] = O[m0 - N + q] with m0 the first
// output index of the call; CE even-phase and CO = CE + N odd-phase samples are carried.
template <int N, int NEW, int CE, typename V = float2> struct HbsStage {
  static constexpr int CO = CE + N;
  V wO[CO + NEW];
  V wE[CE + NEW];
  FMR_HD void clear(float v = 0.f) {
#pragma unroll
    for (int q = 0; q < CO + NEW; q++) hbs_fill(wO[q], v);
#pragma unroll
    for (int q = 0; q < CE + NEW; q++) hbs_fill(wE[q], v);
  }
  // in[0] has an even absolute index
  FMR_HD void run(const V (&in)[2 * NEW], const float *t, V (&out)[NEW]) {
#pragma unroll
    for (int q = 0; q < NEW; q++) {
      wE[CE + q] = in[2 * q];
      wO[CO + q] = in[2 * q + 1];
    }
#pragma unroll
    for (int r = 0; r < NEW; r++) {
      V y = wE[r];
#pragma unroll
      for (int k = 0; k < N; k++) y = hbs_acc(y, t[k], wO[r + N + k], wO[r + N - k - 1]);
      out[r] = y;
    }
#pragma unroll
    for (int q = 0; q < CO; q++) wO[q] = wO[q + NEW];
#pragma unroll
    for (int q = 0; q < CE; q++) wE[q] = wE[q + NEW];
  }
};

// The cascade over U macro-steps (16*U input samples -> 2*U outputs): stage 1 runs per
// macro-step, stages 2 and 3 once per block step.
template <int N1, int N2, int N3, int U, typename V = float2> struct HbsCascade {
  using D = HbsDelays<N1, N2, N3>;
  HbsStage<N1, 8, D::CE1, V> s1;
  HbsStage<N2, 4 * U, D::CE2, V> s2;
  HbsStage<N3, 2 * U, D::CE3, V> s3;
  V x1[8 * U];
  FMR_HD void clear(float v = 0.f) {
    s1.clear(v);
    s2.clear(v);
    s3.clear(v);
  }
  // u-th macro-step of the block: 16 input samples
  FMR_HD void feed(int u, const V (&x)[16], const float *t1) {
    V o[8];
    s1.run(x, t1, o);
#pragma unroll
    for (int q = 0; q < 8; q++) x1[8 * u + q] = o[q];
  }
  // after U feeds: final outputs y[0..2U) = x3[2*i0 - A3 ...], i0 = first macro-step of the block
  FMR_HD void finish(const float *t2, const float *t3, V (&y)[2 * U]) {
    V x2[4 * U];
    s2.run(x1, t2, x2);
    s3.run(x2, t3, y);
  }
};

#if defined(__CUDACC__)
constexpr int kHbsThreads = 128;
constexpr int kHbsU = 4;                         // macro-steps per block step
constexpr int kHbsPitch = 144;                   // bytes per stream row: 128 + 16 (bank skew)
constexpr int kHbsStageBytes = 32 * kHbsPitch;   // per warp
constexpr int hbs_smem_bytes(int stages) { return (kHbsThreads / 32) * stages * kHbsStageBytes; }

struct HbsParams {
  const float2 *lin; // [C][stride] caller's buffer of this call; lin[c][0] has absolute index `start`
  size_t stride;
  long long start;
  long long a_out;   // first final-rate output (absolute, even)
  int tile;          // outputs per stream tile (multiple of 2*U)
  int tiles_per_ch;
  int n_streams;     // channels * tiles_per_ch
  int n_block_steps; // block steps every stream runs: ceil((kWarm + tile/2) / U)
  int l2_prefetch;   // > 0: every 8 chunks a lane asks the L2 for the 1 KB that lies this many chunks ahead
  float t1[8], t2[8], t3[8];
};

// k_hb_stream: outputs [a_out, a_out + tiles_per_ch*tile) of every channel. All input samples the
// streams touch must lie inside the call's buffer and be 16-byte aligned (the host checks).
template <int N1, int N2, int N3, int U, int kHbsStages>
__global__ void __launch_bounds__(kHbsThreads, (kHbsStages <= 4 ? 3 : 2)) k_hb_stream(HbsParams P, Ring<float2> out) {
  using D = HbsDelays<N1, N2, N3>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *ring = smem_raw + warp * (kHbsStages * kHbsStageBytes);
  int sid = blockIdx.x * kHbsThreads + threadIdx.x;
  const bool live = sid < P.n_streams;
  if (!live) sid = P.n_streams - 1; // duplicate work, no stores: keeps the warp's copies uniform
  const int ch = sid / P.tiles_per_ch, tl = sid - ch * P.tiles_per_ch;
  const long long m_lo = P.a_out + (long long)tl * P.tile;
  const long long i_out = (m_lo + D::A3) >> 1;       // macro-step that emits x3[m_lo]
  const long long i_first = i_out - D::kWarm;
  // this stream's first chunk
  const char *my_src = reinterpret_cast<const char *>(P.lin + (size_t)ch * P.stride) + (16 * i_first - P.start) * 8;
  // transposed copy assignment: instruction k, lane l -> stream 4k + (l >> 3), bytes (l & 7)*16
  const char *src[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const unsigned long long a = (unsigned long long)my_src;
    const int from = 4 * k + (lane >> 3);
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)a, from);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(a >> 32), from);
    src[k] = reinterpret_cast<const char *>(((unsigned long long)hi << 32) | lo) + (lane & 7) * 16;
  }
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
  const unsigned dst0 = ring_s + (lane >> 3) * kHbsPitch + (lane & 7) * 16; // + k*4*pitch + stage*stageBytes
  const unsigned row0 = ring_s + lane * kHbsPitch;
  const int n_steps = P.n_block_steps * U;
  auto issue = [&](int step) {
    if (step < n_steps) {
      const unsigned d = dst0 + (step % kHbsStages) * kHbsStageBytes;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + k * 4 * kHbsPitch),
                     "l"(src[k] + (size_t)step * 128));
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    if (P.l2_prefetch > 0 && (step & 7) == 0 && step + P.l2_prefetch + 8 <= n_steps) {
      // DRAM-friendly read-ahead: one 1 KB request per stream instead of eight 128-byte ones
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], 1024;\n" ::"l"(my_src + (size_t)(step + P.l2_prefetch) * 128));
    }
  };
#pragma unroll
  for (int s = 0; s < kHbsStages - 1; s++) issue(s);

  HbsCascade<N1, N2, N3, U> cas;
  cas.clear();
  float t1[N1], t2[N2], t3[N3];
#pragma unroll
  for (int k = 0; k < N1; k++) t1[k] = P.t1[k];
#pragma unroll
  for (int k = 0; k < N2; k++) t2[k] = P.t2[k];
#pragma unroll
  for (int k = 0; k < N3; k++) t3[k] = P.t3[k];
  float2 *__restrict__ orow = out.base + (size_t)ch * out.cap;
  const unsigned omask = out.cap - 1;
  const long long m_hi = m_lo + P.tile;
  long long m = 2 * i_first - D::A3; // first stage-3 output index of the next block step
  for (int bs = 0; bs < P.n_block_steps; bs++) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int step = bs * U + u;
      asm volatile("cp.async.wait_group %0;\n" ::"n"(kHbsStages - 2));
      __syncwarp();
      const unsigned r = row0 + (step % kHbsStages) * kHbsStageBytes;
      float2 x[16];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(r + q * 16));
        x[2 * q] = make_float2(v.x, v.y);
        x[2 * q + 1] = make_float2(v.z, v.w);
      }
      issue(step + kHbsStages - 1);
      cas.feed(u, x, t1);
    }
    float2 y[2 * U];
    cas.finish(t2, t3, y);
    if (live) {
#pragma unroll
      for (int q = 0; q < 2 * U; q += 2) {
        const long long mq = m + q;
        if (mq >= m_lo && mq < m_hi) {
          // two consecutive outputs, 16-byte aligned (mq even)
          float4 v = make_float4(y[q].x, y[q].y, y[q + 1].x, y[q + 1].y);
          *reinterpret_cast<float4 *>(orow + ((unsigned)mq & omask)) = v;
        }
      }
    }
    m += 2 * U;
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
}

// ---------------------------------------------------------------------------------------
// k_hb_stream_tma — the same streaming cascade with the staging done by the TMA unit: every lane
// issues ONE bulk copy (cp.async.bulk.shared.global, SASS UBLKCP) of ROWSTEPS*128 contiguous
// bytes of its own stream into its own shared-memory row; completion is counted in bytes on an
// mbarrier per (warp, stage). Compared with the LDGSTS version the requests are 2-4x larger
// (friendlier to the DRAM pages), do not occupy L1 miss-tracking resources, and cost one
// instruction per row instead of eight per 128 bytes.
template <int ROWSTEPS> struct HbsTma {
  static constexpr int kRowBytes = ROWSTEPS * 128;
  static constexpr int kPitch = kRowBytes + 16; // odd multiple of 16 bytes: conflict-free LDS.128
  static constexpr int kStageBytes = 32 * kPitch;
  static constexpr int smem_bytes(int stages, int warps) { return 256 + warps * stages * kStageBytes; }
};

__device__ __forceinline__ void mbar_init(unsigned a, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

template <int N1, int N2, int N3, int U, int ROWSTEPS, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_hb_stream_tma(HbsParams P, Ring<float2> out) {
  using D = HbsDelays<N1, N2, N3>;
  using T = HbsTma<ROWSTEPS>;
  static_assert(U % ROWSTEPS == 0, "a block step must be a whole number of rows");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem_raw);
  const unsigned mbar0 = smem_s + warp * (STAGES * 8);
  const unsigned ring_s = smem_s + 256 + warp * (STAGES * T::kStageBytes);
  static_assert(WARPS * STAGES * 8 <= 256, "mbarrier area");
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; st++) mbar_init(mbar0 + 8 * st, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  int sid = blockIdx.x * (WARPS * 32) + threadIdx.x;
  const bool live = sid < P.n_streams;
  if (!live) sid = P.n_streams - 1; // duplicate work, no stores: keeps the byte count per stage fixed
  const int ch = sid / P.tiles_per_ch, tl = sid - ch * P.tiles_per_ch;
  const long long m_lo = P.a_out + (long long)tl * P.tile;
  const long long i_out = (m_lo + D::A3) >> 1;
  const long long i_first = i_out - D::kWarm;
  const char *my_src = reinterpret_cast<const char *>(P.lin + (size_t)ch * P.stride) + (16 * i_first - P.start) * 8;
  const unsigned row_s = ring_s + lane * T::kPitch;
  const int n_rows = P.n_block_steps * (U / ROWSTEPS);
  auto issue = [&](int r) {
    if (r < n_rows) {
      const int st = r % STAGES;
      if (lane == 0) mbar_expect_tx(mbar0 + 8 * st, 32 * T::kRowBytes);
      bulk_g2s(row_s + st * T::kStageBytes, my_src + (size_t)r * T::kRowBytes, T::kRowBytes, mbar0 + 8 * st);
    }
  };
#pragma unroll
  for (int r = 0; r < STAGES; r++) issue(r);

  HbsCascade<N1, N2, N3, U> cas;
  cas.clear();
  float t1[N1], t2[N2], t3[N3];
#pragma unroll
  for (int k = 0; k < N1; k++) t1[k] = P.t1[k];
#pragma unroll
  for (int k = 0; k < N2; k++) t2[k] = P.t2[k];
#pragma unroll
  for (int k = 0; k < N3; k++) t3[k] = P.t3[k];
  float2 *__restrict__ orow = out.base + (size_t)ch * out.cap;
  const unsigned omask = out.cap - 1;
  const long long m_hi = m_lo + P.tile;
  long long m = 2 * i_first - D::A3;
  int r = 0;
  for (int bs = 0; bs < P.n_block_steps; bs++) {
#pragma unroll
    for (int q = 0; q < U / ROWSTEPS; q++, r++) {
      const int st = r % STAGES;
      mbar_wait(mbar0 + 8 * st, (unsigned)((r / STAGES) & 1));
      const unsigned a = row_s + st * T::kStageBytes;
#pragma unroll
      for (int ms = 0; ms < ROWSTEPS; ms++) {
        float2 x[16];
#pragma unroll
        for (int w = 0; w < 8; w++) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(a + ms * 128 + w * 16));
          x[2 * w] = make_float2(v.x, v.y);
          x[2 * w + 1] = make_float2(v.z, v.w);
        }
        cas.feed(q * ROWSTEPS + ms, x, t1);
      }
      // every lane has consumed its row (the loads feed the arithmetic above): hand the stage back
      __syncwarp();
      issue(r + STAGES);
    }
    float2 y[2 * U];
    cas.finish(t2, t3, y);
    if (live) {
#pragma unroll
      for (int q = 0; q < 2 * U; q += 2) {
        const long long mq = m + q;
        if (mq >= m_lo && mq < m_hi) {
          float4 v = make_float4(y[q].x, y[q].y, y[q + 1].x, y[q + 1].y);
          *reinterpret_cast<float4 *>(orow + ((unsigned)mq & omask)) = v;
        }
      }
    }
    m += 2 * U;
  }
}
#endif // __CUDACC__

} // namespace fmr
#endif
