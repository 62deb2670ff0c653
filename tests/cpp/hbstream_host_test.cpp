// Host emulation of the streaming half-band cascade (csrc/fmr_hbstream.cuh): the register
// delay-line bookkeeping must reproduce the direct three-stage evaluation bit for bit, from a
// poisoned (NaN) initial state, for every alignment of the stream tile.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../airspy_fmradion_b200/csrc/fmr_hbstream.cuh"
using namespace fmr;

static std::vector<float2> stage_direct(const std::vector<float2> &x, const float *t, int N) {
  // y[m] = x[2m] + sum_k t[k] (x[2m+2k+1] + x[2m-2k-1]); out-of-range -> NaN (must not be used)
  const long long L = (long long)x.size();
  std::vector<float2> y(L / 2);
  for (long long m = 0; m < L / 2; m++) {
    bool ok = (2 * m - 2 * (N - 1) - 1 >= 0) && (2 * m + 2 * (N - 1) + 1 < L);
    float2 v;
    if (!ok) {
      v.x = v.y = NAN;
    } else {
      v = x[2 * m];
      for (int k = 0; k < N; k++) v = hbs_acc(v, t[k], x[2 * m + 2 * k + 1], x[2 * m - 2 * k - 1]);
    }
    y[m] = v;
  }
  return y;
}

template <int N1, int N2, int N3, int U> static int run_case(unsigned seed) {
  using D = HbsDelays<N1, N2, N3>;
  const long long L = 16 * 400;
  std::vector<float2> x0(L);
  srand(seed);
  for (auto &v : x0) {
    v.x = (float)rand() / RAND_MAX - 0.5f;
    v.y = (float)rand() / RAND_MAX - 0.5f;
  }
  float t1[8], t2[8], t3[8];
  for (int k = 0; k < 8; k++) {
    t1[k] = 0.6f / (k + 1) * ((k & 1) ? -1 : 1);
    t2[k] = 0.61f / (k + 1.5f) * ((k & 1) ? -1 : 1);
    t3[k] = 0.63f / (k + 1.2f) * ((k & 1) ? -1 : 1);
  }
  auto x1 = stage_direct(x0, t1, N1);
  auto x2 = stage_direct(x1, t2, N2);
  auto x3 = stage_direct(x2, t3, N3);
  int bad = 0;
  for (long long m_lo = 60; m_lo < 60 + 16; m_lo += 2) {
    const int tile = 8 * U * 6;
    const long long i_out = (m_lo + D::A3) >> 1;
    const long long i_first = i_out - D::kWarm;
    const int nbs = (D::kWarm + tile / 2 + U - 1) / U;
    if (16 * i_first < 0 || 16 * (i_first + (long long)nbs * U) > L) {
      printf("range error\n");
      return 1;
    }
    HbsCascade<N1, N2, N3, U> cas;
    cas.clear(NAN);
    long long m = 2 * i_first - D::A3;
    for (int bs = 0; bs < nbs; bs++) {
      for (int u = 0; u < U; u++) {
        float2 x[16];
        const long long i = i_first + bs * U + u;
        for (int q = 0; q < 16; q++) x[q] = x0[16 * i + q];
        cas.feed(u, x, t1);
      }
      float2 y[2 * U];
      cas.finish(t2, t3, y);
      for (int q = 0; q < 2 * U; q++) {
        const long long mq = m + q;
        if (mq >= m_lo && mq < m_lo + tile) {
          if (memcmp(&y[q], &x3[mq], sizeof(float2)) != 0) {
            if (bad < 5) printf("mismatch N=(%d,%d,%d) U=%d m_lo=%lld m=%lld: %g %g vs %g %g\n", N1, N2, N3, U, m_lo, mq, y[q].x, y[q].y, x3[mq].x, x3[mq].y);
            bad++;
          }
        }
      }
      m += 2 * U;
    }
  }
  printf("N=(%d,%d,%d) U=%d A=(%d,%d,%d) warm=%d : %s\n", N1, N2, N3, U, D::A1, D::A2, D::A3, D::kWarm, bad ? "FAIL" : "ok");
  return bad;
}

int main() {
  int bad = 0;
  bad += run_case<4, 5, 8, 1>(1);
  bad += run_case<4, 5, 8, 2>(2);
  bad += run_case<4, 5, 8, 4>(3);
  bad += run_case<3, 4, 6, 2>(4);
  bad += run_case<5, 7, 8, 2>(5);
  return bad ? 1 : 0;
}
