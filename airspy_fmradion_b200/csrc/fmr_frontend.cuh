// fmr_frontend.cuh — the fused, persistent front end of the 10 MHz chain: FourthConverterIQ-free cf32 input ->
// three half-band stages (r8b::CDSPHBDownsampler, CDSPHBDownsampler.h:137-239) -> long low-pass + polyphase bank
// (CDSPBlockConvolver.h:252-353, CDSPFracInterpolator.h:861-925, in the frequency-domain form of fmr_fdr.cuh) in ONE
// kernel; the 1.25 MHz stream between them lives in shared memory only. Reference shape: r8brain runs all stages of a
// chain back to back on a cache-resident block (CDSPResampler.h:559-575).
//
// One CTA per SM, persistent over channels; a CTA takes one channel at a time through all blocks of the call.
// Warp-specialised, register budgets moved with setmaxnreg:
//   * 4 producer warps (208 registers): every thread runs the register-resident streaming half-band cascade
//     (HbsCascade, fmr_hbstream.cuh) over its own stream tile = 60 consecutive 1.25 MHz samples of the current block
//     (125 tiles x 60 = the 7500 new samples of a block). Its 10 MHz input rows are staged by the TMA unit: one
//     5-D tensor-map copy per warp and row (box = 32 tiles x 256 bytes, SWIZZLE_128B so that the lanes' LDS.128 of
//     their own rows are conflict free), completion counted on an mbarrier per warp and stage. Outputs go straight
//     into B, a 10000-sample circular buffer in shared memory indexed by the absolute 1.25 MHz sample number.
//   * 8 consumer warps (120 registers): when B holds a whole block (7500 new + 2500 samples it shares with the
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

#include "fmr_fdr.cuh"
#include "fmr_hbstream.cuh"

namespace fmr {
namespace fe {

using D = HbsDelays<4, 5, 8>;
constexpr int kU = 4;                                              // macro-steps (16 input samples) per block step
constexpr int kProdWarps = 4, kConsWarps = 8;
constexpr int kProdThreads = 32 * kProdWarps, kConsThreads = 32 * kConsWarps, kThreads = kProdThreads + kConsThreads;
constexpr int kTile = 60, kTiles = fdr::kAdvIn / kTile;            // 125 stream tiles of 60 outputs per block
static_assert(kTiles * kTile == fdr::kAdvIn && kTiles <= kProdThreads, "tiles must cover the block's new samples");
constexpr int kBlockSteps = (D::kWarm + kTile / 2 + kU - 1) / kU;  // 11: 3 of warm-up, 7.5 of output
constexpr int kRows = kBlockSteps * kU / 2;                        // 22 rows of 32 input samples per tile and block
constexpr int kStages = 2;
constexpr int kStageBytes = 2 * 32 * 128;                          // one warp's box: 2 chunks x 32 tiles x 128 bytes
constexpr int kTileIn = 8 * kTile;                                 // 480 input samples between tiles
constexpr int kBlockIn = 8 * fdr::kAdvIn;                          // 60000 input samples between blocks
constexpr int kRowChunks = 2 * kRows;                              // 44 chunks of 16 samples per tile and block
// first input sample of (block j, tile 0, row 0): 16 * (((7500 j + 1250 + A3) >> 1) - kWarm) = 60000 j + kIn0
constexpr int kIn0 = 16 * (((fdr::kGuardIn + D::A3) >> 1) - D::kWarm);
constexpr int kInSpan = kIn0 + (kTiles - 1) * kTileIn + kRows * 32; // input of block j ends at 60000 j + kInSpan (exclusive)
// shared memory map (bytes)
constexpr int kOffA = 0;
constexpr int kOffB = fdr::kNin * 8;
constexpr int kOffStage = ((2 * fdr::kNin * 8 + 1023) / 1024) * 1024;
constexpr int kOffBar = kOffStage + kProdWarps * kStages * kStageBytes;
constexpr int kSmemBytes = kOffBar + 128;
static_assert(kSmemBytes <= 232448, "shared memory budget");
static_assert(fdr::kZLen <= fdr::kNin, "the 3072-point buffer aliases A");

struct Params {
  const float2 *hb_ring; // 1.25 MHz ring (unfused path's output): samples shared with the block before the first one
  uint32_t hb_cap;
  float2 *out;           // 384 kHz ring
  uint32_t out_cap;
  const float *Hs;       // fdr spectrum table
  const float2 *tab;     // fdr twiddle table
  int64_t j0;            // first block of the absolute grid this launch takes
  int n_blocks;          // blocks per channel
  int n_channels;
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
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); }

// The row that a producer warp requests next: rows of a tile follow each other through the blocks of a channel and
// then through the CTA's channels, so the TMA pipeline never drains at a block or channel boundary.
struct RowIter {
  int ch, blk, row;
  __device__ __forceinline__ void next(int n_blocks, int ch_step) {
    if (++row == kRows) {
      row = 0;
      if (++blk == n_blocks) {
        blk = 0;
        ch += ch_step;
      }
    }
  }
};

__global__ void __launch_bounds__(kThreads, 1) k_frontend_fused(const __grid_constant__ CUtensorMap tm, const Params P) {
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int lane = tid & 31, warp = tid >> 5;
    const bool live = tid < kTiles;
    const unsigned mbar0 = bar_full + 8 * kStages * warp;
    const unsigned stage0 = smem_s + kOffStage + warp * (kStages * kStageBytes);
    // this lane's two 128-byte lines of a stage (chunk 0 / chunk 1 of its tile) and their swizzle terms
    const unsigned line0 = stage0 + lane * 128, line1 = line0 + 32 * 128;
    const unsigned sw0 = (lane & 7) << 4, sw1 = sw0; // line index = chunk * 32 + lane: (line & 7) = (lane & 7)
    float t1[4], t2[5], t3[8];
#pragma unroll
    for (int k = 0; k < 4; k++) t1[k] = P.t1[k];
#pragma unroll
    for (int k = 0; k < 5; k++) t2[k] = P.t2[k];
#pragma unroll
    for (int k = 0; k < 8; k++) t3[k] = P.t3[k];
    RowIter nx{(int)blockIdx.x, 0, 0};
    auto issue = [&](int rseq) {
      if (nx.ch < P.n_channels) {
        if (lane == 0) {
          const unsigned mb = mbar0 + 8 * (rseq % kStages);
          mbar_expect_tx(mb, kStageBytes);
          tma_load_5d(stage0 + (rseq % kStages) * kStageBytes, &tm, 0, 32 * warp, 2 * nx.row, nx.blk, nx.ch, mb);
        }
        nx.next(P.n_blocks, ch_step);
      }
    };
    int rseq = 0; // rows consumed so far by this warp (all channels)
#pragma unroll
    for (int s = 0; s < kStages; s++) issue(s);
    uint32_t bseq = 0; // blocks finished so far by this CTA
    for (int ch = blockIdx.x; ch < P.n_channels; ch += ch_step) {
      for (int blk = 0; blk < P.n_blocks; blk++, bseq++) {
        const int64_t j = P.j0 + blk;
        const int64_t m_lo = j * fdr::kAdvIn + fdr::kGuardIn + (int64_t)kTile * (live ? tid : kTiles - 1);
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
        HbsCascade<4, 5, 8, kU> cas;
        cas.clear();
        for (int bs = 0; bs < kBlockSteps; bs++) {
#pragma unroll
          for (int q = 0; q < kU / 2; q++, rseq++) {
            const int st = rseq % kStages;
            mbar_wait(mbar0 + 8 * st, (unsigned)((rseq / kStages) & 1));
            const unsigned a0 = line0 + st * kStageBytes, a1 = line1 + st * kStageBytes;
#pragma unroll
            for (int ms = 0; ms < 2; ms++) {
              float2 x[16];
              const unsigned a = ms ? a1 : a0, sw = ms ? sw1 : sw0;
#pragma unroll
              for (int w = 0; w < 8; w++) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(a + ((w << 4) ^ sw)));
                x[2 * w] = make_float2(v.x, v.y);
                x[2 * w + 1] = make_float2(v.z, v.w);
              }
              cas.feed(2 * q + ms, x, t1);
            }
            __syncwarp();
            issue(rseq + kStages);
          }
          float2 y[2 * kU];
          cas.finish(t2, t3, y);
          if (bs >= D::kWarm / kU) {
            // B still holds the previous block until pass 1 of its FFT has read it
            if (bs == D::kWarm / kU && blk > 0) mbar_wait(bar_bfree, (bseq - 1) & 1);
            if (live) {
#pragma unroll
              for (int q = 0; q < 2 * kU; q += 2) {
                if (bs < kBlockSteps - 1 || q < kTile - 2 * kU * (kBlockSteps - 1 - D::kWarm / kU)) {
                  *reinterpret_cast<float4 *>(B + pos) = make_float4(y[q].x, y[q].y, y[q + 1].x, y[q + 1].y);
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    const int ct = tid - kProdThreads;
    uint32_t bseq = 0;
    for (int ch = blockIdx.x; ch < P.n_channels; ch += ch_step) {
      for (int blk = 0; blk < P.n_blocks; blk++, bseq++) {
        const int64_t j = P.j0 + blk;
        const int o625 = (int)((12 * j + 14) & 15); // (7500 j - 1250) mod 10000 = 625 * o625
        mbar_wait(bar_bfull, bseq & 1);
        for (int b = ct; b < 625; b += kConsThreads) {
          fdr::fwd1(b, [&](int bb, int a) { return B[bb + 625 * ((o625 + a) & 15)]; }, A, P.tab);
        }
        cons_sync();
        if (ct == 0) mbar_arrive(bar_bfree);
        for (int i = ct; i < 400; i += kConsThreads) fdr::fwd2(i, A, P.tab);
        cons_sync();
        {
          float2 o0[fdr::kKeep], o1[fdr::kKeep];
          fdr::fwd3_compute(ct, A, P.Hs, o0);
          if (ct + kConsThreads < 400) fdr::fwd3_compute(ct + kConsThreads, A, P.Hs, o1);
          cons_sync(); // the 3072-point buffer aliases A: every butterfly has loaded before anyone stores
          fdr::fwd3_store(ct, A, o0);
          if (ct + kConsThreads < 400) fdr::fwd3_store(ct + kConsThreads, A, o1);
        }
        cons_sync();
        if (ct < 192) fdr::inv1(ct, A, P.tab);
        cons_sync();
        if (ct < 192) fdr::inv2(ct, A, P.tab);
        cons_sync();
        {
          const int64_t mb = j * fdr::kAdvOut - fdr::kGuardOut;
          float2 *__restrict__ orow = P.out + (size_t)ch * P.out_cap;
          const uint32_t omask = P.out_cap - 1;
          fdr::inv3(ct, A, [&](int i, float2 v) {
            if (i >= fdr::kGuardOut && i < fdr::kGuardOut + fdr::kAdvOut) orow[(uint32_t)(mb + i) & omask] = v;
          });
        }
        cons_sync(); // A is rewritten by pass 1 of the next block
      }
    }
  }
}

} // namespace fe
#endif // __CUDACC__
} // namespace fmr
#endif
