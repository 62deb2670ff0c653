// fmr_frontend.cuh — the fused, persistent front end of the 10 MHz chain: FourthConverterIQ-free cf32 input ->
// three half-band stages (r8b::CDSPHBDownsampler, CDSPHBDownsampler.h:137-239) -> long low-pass + polyphase bank
// (CDSPBlockConvolver.h:252-353, CDSPFracInterpolator.h:861-925, in the frequency-domain form of fmr_fdr.cuh) in ONE
// kernel; the 1.25 MHz stream between them lives in shared memory only. Reference shape: r8brain runs all stages of a
// chain back to back on a cache-resident block (CDSPResampler.h:559-575).
//
// One CTA per SM, persistent over channels; a CTA takes one channel at a time through all blocks of the call.
// Warp-specialised, register budgets moved with setmaxnreg:
//   * 8 producer warps (152 registers): every thread runs the register-resident streaming half-band cascade
//     (HbsCascade, fmr_hbstream.cuh) over its own stream tile = 30 consecutive 1.25 MHz samples of the current block
//     (250 tiles x 30 = the 7500 new samples of a block). Its 10 MHz input rows are staged by the TMA unit: one
//     5-D tensor-map copy per warp and row (box = 32 tiles x 128 bytes, SWIZZLE_128B so that the lanes' LDS.128 of
//     their own rows are conflict free), completion counted on an mbarrier per warp and stage. Outputs go straight
//     into B, a 10000-sample circular buffer in shared memory indexed by the absolute 1.25 MHz sample number.
//     (Two producer warps per scheduler: one alone runs at half the FMA pipe's packed-FP32 rate, measured.)
//   * 8 consumer warps (104 registers): when B holds a whole block (7500 new + 2500 samples it shares with the
//     previous block), pass 1 of the 10000-point forward FFT reads B and writes A; B is handed back to the producers,
//     which meanwhile kept their TMA pipeline filled; passes 2 and 3, the multiplication by H0, the 3072-point inverse
//     and the stores to the 384 kHz ring run out of A while the producers fill B with the next block.
// So the HBM-bound half (input streaming) and the shared-memory/FP32-bound half (FFT) of the front end overlap on
// every SM instead of running as two kernels with a ring in HBM between them.
//
// A block can only be taken by this kernel when all the 10 MHz input its stream tiles touch (including the 192 samples
// of warm-up in front of every tile) lies in the caller's buffer of this call; the one or two blocks that straddle a call
// boundary go through the unfused kernels (k_hb_stream_tma / k_hb_cascade -> ring -> k_fdr), which also leave in the
// 1.25 MHz ring the 2500 samples the first fused block shares with its predecessor. Both paths run the same per-thread
// code on the same absolute block grid, so the result is bit-identical (tests/test_frontend_gpu.py).
#ifndef FMR_FRONTEND_CUH
#define FMR_FRONTEND_CUH

#include <cuda.h>

#include <type_traits>

#include "fmr_fdr.cuh"
#include "fmr_hbstream.cuh"

namespace fmr {
namespace fe {

using D = HbsDelays<4, 5, 8>;
constexpr int kConsWarpsDefault = 8;
constexpr int kStages = 2;
constexpr int kBlockIn = 8 * fdr::kAdvIn;                          // 60000 input samples between blocks
// first input sample of (block j, tile 0, row 0): 16 * (((7500 j + 1250 + A3) >> 1) - kWarm) = 60000 j + kIn0
constexpr int kIn0 = 16 * (((fdr::kGuardIn + D::A3) >> 1) - D::kWarm);
static_assert(fdr::kZLen <= fdr::kNin, "the 3072-point buffer aliases A");

// One shape of the kernel: PW producer warps, U macro-steps (16 input samples) per block step of the cascade,
// stream tiles of TILE outputs, RS macro-steps per staged row, setmaxnreg budgets PREGS / CREGS.
// SPLIT: a stream tile is run by TWO threads (even lane: real parts, odd lane: imaginary parts; the filters are real), i.e.
// twice the warps for the same arithmetic, scalar FADD / FFMA instead of the packed forms, half the registers.
template <int PW, int U_, int TILE, int RS, int PREGS, int CREGS, bool SPLIT = false, int CW = kConsWarpsDefault> struct Cfg {
  static constexpr int kConsWarps = CW, kConsThreads = 32 * CW;
  static constexpr int kProdWarps = PW, kU = U_, kTile = TILE, kRowSteps = RS, kProdRegs = PREGS, kConsRegs = CREGS;
  static constexpr bool kSplit = SPLIT;
  static constexpr int kProdThreads = 32 * PW, kThreads = kProdThreads + kConsThreads;
  static constexpr int kTiles = fdr::kAdvIn / TILE;
  static constexpr int kWarpTiles = SPLIT ? 16 : 32;                // stream tiles per producer warp
  static_assert(kTiles * TILE == fdr::kAdvIn && kTiles <= PW * kWarpTiles && (TILE % 2) == 0, "tiles must cover the block's new samples");
  static_assert(U_ % RS == 0, "a block step is a whole number of rows");
  static constexpr int kBlockSteps = (D::kWarm + TILE / 2 + U_ - 1) / U_;
  static constexpr int kRows = kBlockSteps * U_ / RS;               // staged rows per tile and block
  static constexpr int kStageBytes = RS * kWarpTiles * 128;        // one warp's box: RS chunks x its tiles x 128 bytes
  static constexpr int kTileIn = 8 * TILE;                         // input samples between tiles
  static constexpr int kRowChunks = RS * kRows;                    // chunks of 16 samples per tile and block
  static constexpr int kInSpan = kIn0 + (kTiles - 1) * kTileIn + kRowChunks * 16; // block j's input ends at 60000 j + kInSpan
  // shared memory map (bytes)
  static constexpr int kOffA = 0;
  static constexpr int kOffB = fdr::kNin * 8;
  static constexpr int kOffStage = ((2 * fdr::kNin * 8 + 1023) / 1024) * 1024;
  static constexpr int kOffBar = kOffStage + PW * kStages * kStageBytes;
  static constexpr int kSmemBytes = kOffBar + 8 * (PW * kStages + 2);
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(PW * 32 * PREGS + kConsThreads * CREGS <= 65536, "register file");
};
// (measured and dropped, profiles/sweep_frontend_variants_r02.jsonl: 8 packed producer warps on tiles of 30 outputs with
// block steps of 32 or 16 samples - 87 % warm-up, register spills or twice the register moves - 13 % / 21 % slower)
using CfgA = Cfg<4, 4, 60, 2, 208, 120>;  // 4 producer warps, long tiles (47 % warm-up), cascade block step of 64 samples
using CfgS = Cfg<8, 4, 60, 2, 128, 128, true>; // 8 producer warps on long tiles: real and imaginary part on two lanes
// DEFAULT: CfgA's producers with four consumer warps (one per scheduler) and one register budget for all eight warps
// (220 registers, no setmaxnreg): the FFT takes twice as long per block but still fits the block period, and takes fewer
// issue slots from the producer warp of its scheduler at any one time: 9.37 vs 9.74 ms (8192 channels)
using CfgC4 = Cfg<4, 4, 60, 2, 232, 232, false, 4>;

struct Params {
  float2 *hb_ring;       // 1.25 MHz ring: read for the samples the first block shares with its predecessor (the unfused
                         // kernels or the previous call left them there), written from `ring_from` on (the last block's
                         // final 2500 samples: what the NEXT block, which straddles the call boundary, shares with it)
  uint32_t hb_cap;
  float2 *out;           // 384 kHz ring
  uint32_t out_cap;
  const float *Hs;       // fdr spectrum table
  const float2 *tab;     // fdr twiddle table
  int64_t j0;            // first block of the absolute grid this launch takes
  int n_blocks;          // blocks per channel
  int n_channels;
  int64_t ring_from;     // first of the 2500 samples of the last block that the consumers copy to the ring
  float t1[8], t2[8], t3[8];
};

} // namespace fe

#if defined(__CUDACC__)
namespace fe {

__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, int c4,
                                            unsigned mbar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(mbar)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
template <int CT> __device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }

// The row that a producer warp requests next: rows of a tile follow each other through the blocks of a channel and
// then through the CTA's channels, so the TMA pipeline never drains at a block or channel boundary.
struct RowIter {
  int ch, blk, row;
  __device__ __forceinline__ void next(int n_rows, int n_blocks, int ch_step) {
    if (++row == n_rows) {
      row = 0;
      if (++blk == n_blocks) {
        blk = 0;
        ch += ch_step;
      }
    }
  }
};

// wait with back-off: the consumers wait for a whole block of producer work, spinning would take issue slots from it
__device__ __forceinline__ void mbar_wait_sleep(unsigned a, unsigned parity) {
  unsigned done;
  for (;;) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(256);
  }
}

template <class CF>
__global__ void __launch_bounds__(CF::kThreads, 1) k_frontend_fused(const __grid_constant__ CUtensorMap tm, const Params P) {
  constexpr int kProdWarps = CF::kProdWarps, kU = CF::kU, kTile = CF::kTile, kRowSteps = CF::kRowSteps;
  constexpr int kProdThreads = CF::kProdThreads, kTiles = CF::kTiles, kBlockSteps = CF::kBlockSteps, kRows = CF::kRows;
  constexpr int kStageBytes = CF::kStageBytes, kOffA = CF::kOffA, kOffB = CF::kOffB, kOffStage = CF::kOffStage, kOffBar = CF::kOffBar;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem_raw);
  float2 *A = reinterpret_cast<float2 *>(smem_raw + kOffA);
  float2 *B = reinterpret_cast<float2 *>(smem_raw + kOffB);
  const unsigned bar_full = smem_s + kOffBar;          // [warp][stage]
  const unsigned bar_bfull = bar_full + 8 * kProdWarps * kStages;
  const unsigned bar_bfree = bar_bfull + 8;
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kProdWarps * kStages; i++) mbar_init(bar_full + 8 * i, 1);
    mbar_init(bar_bfull, kProdThreads);
    mbar_init(bar_bfree, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ch_step = gridDim.x;
  if (tid < kProdThreads) {
    // =============================== producers: half-band cascade, HBM -> B ===============================
    if constexpr (CF::kProdRegs != CF::kConsRegs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CF::kProdRegs));
    constexpr bool kSplit = CF::kSplit;
    constexpr int kWarpTiles = CF::kWarpTiles;
    using V = typename std::conditional<kSplit, float, float2>::type;
    const int lane = tid & 31, warp = tid >> 5;
    const int wl = kSplit ? (lane >> 1) : lane;  // this lane's tile within the warp's box
    const int comp = kSplit ? (lane & 1) : 0;    // split form: 0 = real parts, 1 = imaginary parts
    const int tile = warp * kWarpTiles + wl;
    const bool live = tile < kTiles;
    const unsigned mbar0 = bar_full + 8 * kStages * warp;
    const unsigned stage0 = smem_s + kOffStage + warp * (kStages * kStageBytes);
    // this lane's 128-byte line of a row's macro-step ms is line (kWarpTiles ms + wl) of the stage: SWIZZLE_128B xors
    // the 16-byte unit index with (line & 7) = (wl & 7)
    const unsigned line0 = stage0 + wl * 128;
    const unsigned sw = (wl & 7) << 4;
    float t1[4], t2[5], t3[8];
#pragma unroll
    for (int k = 0; k < 4; k++) t1[k] = P.t1[k];
#pragma unroll
    for (int k = 0; k < 5; k++) t2[k] = P.t2[k];
#pragma unroll
    for (int k = 0; k < 8; k++) t3[k] = P.t3[k];
    RowIter nx{(int)blockIdx.x, 0, 0};
    auto issue = [&](int rs) {
      if (nx.ch < P.n_channels) {
        if (lane == 0) {
          const unsigned mb = mbar0 + 8 * (rs % kStages);
          mbar_expect_tx(mb, kStageBytes);
          tma_load_5d(stage0 + (rs % kStages) * kStageBytes, &tm, 0, kWarpTiles * warp, kRowSteps * nx.row, nx.blk, nx.ch, mb);
        }
        nx.next(kRows, P.n_blocks, ch_step);
      }
    };
    int rseq = 0; // rows consumed so far by this warp (all channels)
#pragma unroll
    for (int s = 0; s < kStages; s++) issue(s);
    uint32_t bseq = 0; // blocks finished so far by this CTA
    for (int ch = blockIdx.x; ch < P.n_channels; ch += ch_step) {
      for (int blk = 0; blk < P.n_blocks; blk++, bseq++) {
        const int64_t j = P.j0 + blk;
        const int64_t m_lo = j * fdr::kAdvIn + fdr::kGuardIn + (int64_t)kTile * (live ? tile : kTiles - 1);
        int pos = (int)(m_lo % fdr::kNin); // B is indexed by the absolute sample number modulo 10000
        if (blk == 0) {
          // first block of a channel: the 2500 samples it shares with its predecessor come from the 1.25 MHz ring
          if (bseq > 0) mbar_wait(bar_bfree, (bseq - 1) & 1);
          const int64_t o_lo = j * fdr::kAdvIn - fdr::kGuardIn;
          const float2 *__restrict__ row = P.hb_ring + (size_t)ch * P.hb_cap;
          for (int q = 2 * tid; q < 2 * fdr::kGuardIn; q += 2 * kProdThreads) {
            const int64_t m = o_lo + q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m >= 0) v = *reinterpret_cast<const float4 *>(row + ((uint32_t)m & (P.hb_cap - 1)));
            const int pb = (int)((m + 4 * (int64_t)fdr::kNin) % fdr::kNin);
            *reinterpret_cast<float4 *>(B + pb) = v;
          }
        }
        HbsCascade<4, 5, 8, kU, V> cas;
        cas.clear();
        for (int bs = 0; bs < kBlockSteps; bs++) {
#pragma unroll
          for (int q = 0; q < kU / kRowSteps; q++, rseq++) {
            const int st = rseq % kStages;
            mbar_wait(mbar0 + 8 * st, (unsigned)((rseq / kStages) & 1));
#pragma unroll
            for (int ms = 0; ms < kRowSteps; ms++) {
              V x[16];
              const unsigned a = line0 + st * kStageBytes + ms * (kWarpTiles * 128);
#pragma unroll
              for (int w = 0; w < 8; w++) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(a + ((w << 4) ^ sw)));
                if constexpr (kSplit) {
                  x[2 * w] = comp ? v.y : v.x;
                  x[2 * w + 1] = comp ? v.w : v.z;
                } else {
                  x[2 * w] = make_float2(v.x, v.y);
                  x[2 * w + 1] = make_float2(v.z, v.w);
                }
              }
              cas.feed(kRowSteps * q + ms, x, t1);
            }
            __syncwarp();
            issue(rseq + kStages);
          }
          V y[2 * kU];
          cas.finish(t2, t3, y);
          if (bs >= D::kWarm / kU) {
            // B still holds the previous block until pass 1 of its FFT has read it
            if (bs == D::kWarm / kU && blk > 0) mbar_wait(bar_bfree, (bseq - 1) & 1);
            if (live) {
#pragma unroll
              for (int q = 0; q < 2 * kU; q += 2) {
                if (bs < kBlockSteps - 1 || q < kTile - 2 * kU * (kBlockSteps - 1 - D::kWarm / kU)) {
                  if constexpr (kSplit) {
                    float *bp = reinterpret_cast<float *>(B + pos) + comp;
                    bp[0] = y[q];
                    bp[2] = y[q + 1];
                  } else {
                    *reinterpret_cast<float4 *>(B + pos) = make_float4(y[q].x, y[q].y, y[q + 1].x, y[q + 1].y);
                  }
                  pos += 2;
                  if (pos >= fdr::kNin) pos -= fdr::kNin;
                }
              }
            }
          }
        }
        mbar_arrive(bar_bfull);
      }
    }
  } else {
    // =============================== consumers: B -> FFT -> 384 kHz ring ===============================
    if constexpr (CF::kProdRegs != CF::kConsRegs) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CF::kConsRegs));
    constexpr int CT = CF::kConsThreads;
    constexpr int R3 = (400 + CT - 1) / CT; // rounds of the last forward pass (its outputs are held over a barrier)
    const int ct = tid - kProdThreads;
    uint32_t bseq = 0;
    for (int ch = blockIdx.x; ch < P.n_channels; ch += ch_step) {
      for (int blk = 0; blk < P.n_blocks; blk++, bseq++) {
        const int64_t j = P.j0 + blk;
        const int o625 = (int)((12 * j + 14) & 15); // (7500 j - 1250) mod 10000 = 625 * o625
        mbar_wait_sleep(bar_bfull, bseq & 1);
        for (int b = ct; b < 625; b += CT) {
          fdr::fwd1(b, [&](int bb, int a) { return B[bb + 625 * ((o625 + a) & 15)]; }, A, P.tab);
        }
        if (blk == P.n_blocks - 1) {
          // last block of the channel: its final 2500 samples are what the next block (which straddles the call boundary
          // and goes through the unfused kernels, now or in the next call) shares with it — leave them in the 1.25 MHz ring
          float2 *__restrict__ rrow = P.hb_ring + (size_t)ch * P.hb_cap;
          for (int q = 2 * ct; q < 2 * fdr::kGuardIn; q += 2 * CT) {
            const int64_t m = P.ring_from + q;
            *reinterpret_cast<float4 *>(rrow + ((uint32_t)m & (P.hb_cap - 1))) = *reinterpret_cast<const float4 *>(B + (int)(m % fdr::kNin));
          }
        }
        cons_sync<CT>();
        if (ct == 0) mbar_arrive(bar_bfree);
        for (int i = ct; i < 400; i += CT) fdr::fwd2(i, A, P.tab);
        cons_sync<CT>();
        {
          float2 o[R3][fdr::kKeep];
#pragma unroll
          for (int r = 0; r < R3; r++) {
            if (ct + r * CT < 400) fdr::fwd3_compute<12>(ct + r * CT, A, P.Hs, o[r]);
          }
          cons_sync<CT>(); // the 3072-point buffer aliases A: every butterfly has loaded before anyone stores
#pragma unroll
          for (int r = 0; r < R3; r++) {
            if (ct + r * CT < 400) fdr::fwd3_store<12>(ct + r * CT, A, o[r]);
          }
        }
        cons_sync<CT>();
        for (int b = ct; b < 192; b += CT) fdr::inv1<12>(b, A, P.tab);
        cons_sync<CT>();
        for (int b = ct; b < 192; b += CT) fdr::inv2<12>(b, A, P.tab);
        cons_sync<CT>();
        {
          const int64_t mb = j * fdr::kAdvOut - fdr::kGuardOut;
          float2 *__restrict__ orow = P.out + (size_t)ch * P.out_cap;
          const uint32_t omask = P.out_cap - 1;
          for (int t = ct; t < 256; t += CT) {
            fdr::inv3<12>(t, A, [&](int i, float2 v) {
              if (i >= fdr::kGuardOut && i < fdr::kGuardOut + fdr::kAdvOut) orow[(uint32_t)(mb + i) & omask] = v;
            });
          }
        }
        cons_sync<CT>(); // A is rewritten by pass 1 of the next block
      }
    }
  }
}

} // namespace fe
#endif // __CUDACC__
} // namespace fmr
#endif
