// Host check of the branch-free fast_atan2f used inside the PLL recurrence (csrc/fmr_kernels.cuh,
// fast_atan2f_bf_t): with an exact quotient it must equal the reference's table method
// (include/Utility.h:236-304) bit for bit on every input; with the reciprocal + residual-correction
// quotient (reciprocal perturbed by +-1 ulp to cover the hardware approximation) mismatches must be
// vanishingly rare and tiny.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
namespace fmr {
struct HbStage{int ntaps; const double* taps;}; struct BcStage{int klen,inputlen,latency,down,outoffset; const double* taps;}; struct FiStage{int instep,outstep,flen; const double* taps;};
struct ChainDesc{double src,dst;int kind;int n_hb;HbStage hb[3];BcStage bc;int has_fi;FiStage fi;int verified;};
#include "../../airspy_fmradion_b200/csrc/fmr_tables_generated.inc"
}
using namespace fmr;
// reference restatement (oracle) of fast_atan2f, table method
static float ref_atan2(float y, float x) {
  const float* tbl = k_fast_atan_table;
  float y_abs = fabsf(y), x_abs = fabsf(x);
  if (!((y_abs > 0.0f) || (x_abs > 0.0f))) return 0.0f;
  float z = (y_abs < x_abs) ? y_abs / x_abs : x_abs / y_abs;
  float base;
  if ((double)z < 0.003921569) base = z;
  else { float alpha = z * 255.0f; int index = ((int)alpha) & 0xff; alpha -= (float)index; base = tbl[index]; base += (tbl[index+1]-tbl[index]) * alpha; }
  float angle;
  if (x_abs > y_abs) { if (x >= 0) angle = (y >= 0) ? base : -base; else { angle = 3.14159265358979323846f; if (y >= 0) angle -= base; else angle = base - angle; } }
  else { if (y >= 0) { angle = 1.57079632679489661923f; if (x >= 0) angle -= base; else angle += base; } else { angle = -1.57079632679489661923f; if (x >= 0) angle += base; else angle -= base; } }
  return angle;
}
static float bf_atan2(float y, float x, int exactdiv) {
  const float* tbl = k_fast_atan_table;
  const float ya = fabsf(y), xa = fabsf(x);
  const float num = fminf(ya, xa), den = fmaxf(ya, xa);
  const bool xbig = xa > ya, xpos = x >= 0.0f, ypos = y >= 0.0f;
  float K = xbig ? (xpos ? 0.0f : 3.14159265358979323846f) : 1.57079632679489661923f;
  float sg = (xbig == xpos) ? 1.0f : -1.0f;
  K = ypos ? K : -K; sg = ypos ? sg : -sg;
  float z;
  if (exactdiv) z = num / den; else { float r = (float)(1.0 / (double)den); uint32_t u; memcpy(&u,&r,4); u += (rand()%3)-1; memcpy(&r,&u,4); z = num * r; z = fmaf(fmaf(-den, z, num), r, z); }
  float alpha = z * 255.0f; const int index = (int)alpha; alpha -= (float)index;
  const float t0 = tbl[index & 0xff], t1 = tbl[(index & 0xff) + 1] - tbl[index & 0xff];
  float base = alpha * t1 + t0;   // reference rounding (no FMA) for the comparison; device uses FFMA like fast_atan2f_dev
  uint32_t T = 0x3b808082; float Tf; memcpy(&Tf,&T,4);
  base = (z < Tf) ? z : base;
  float angle = fmaf(sg, base, K);
  return (den > 0.0f) ? angle : 0.0f;
}
// fmr_atan2f (csrc/fmr_kernels.cuh): the discriminator's branch-free atan2f, same operations;
// the device's rcp.approx is modelled as the rounded reciprocal perturbed by +-1 ulp.
static inline float poly_atan2f(float y, float x) {
  const float ya = fabsf(y), xa = fabsf(x);
  const float mn = fminf(ya, xa), mx = fmaxf(ya, xa);
  float r0 = (float)(1.0 / (double)mx);
  { uint32_t u; memcpy(&u, &r0, 4); u += (rand() % 3) - 1; memcpy(&r0, &u, 4); }
  const float r1 = fmaf(fmaf(-mx, r0, 1.0f), r0, r0);
  float z = mn * r1;
  z = fmaf(fmaf(-mx, z, mn), r1, z);
  const float s = z * z;
  float p = 0.0028662257f;
  p = fmaf(p, s, -0.0161657367f);
  p = fmaf(p, s, 0.0429096138f);
  p = fmaf(p, s, -0.0752896400f);
  p = fmaf(p, s, 0.1065626393f);
  p = fmaf(p, s, -0.1420889944f);
  p = fmaf(p, s, 0.1999355085f);
  p = fmaf(p, s, -0.3333314528f);
  float r = fmaf(p * s, z, z);
  r = (ya > xa) ? 1.57079632679489661923f - r : r;
  r = (x < 0.0f) ? 3.14159265358979323846f - r : r;
  r = (mx > 0.0f) ? r : 0.0f;
  return copysignf(r, y);
}
static int check_poly_atan2() {
  double maxe = 0, sum2 = 0;
  long n = 0;
  for (long it = 0; it < 4000000; it++) {
    float y = ((float)rand() / RAND_MAX - 0.5f) * powf(10.f, (rand() % 6) - 3);
    float x = ((float)rand() / RAND_MAX - 0.5f) * powf(10.f, (rand() % 6) - 3);
    if (it % 1001 == 0) y = 0;
    if (it % 1777 == 0) x = 0;
    if (it % 5003 == 0) y = x;
    const double ref = atan2((double)y, (double)x);
    double e = fabs((double)poly_atan2f(y, x) - ref);
    if (e > 3.2) e = fabs(e - 2 * M_PI); // -pi vs +pi on the cut
    if (e > maxe) maxe = e;
    sum2 += e * e;
    n++;
  }
  const int zero_ok = (poly_atan2f(0.f, 0.f) == 0.0f);
  printf("polynomial atan2f: n=%ld max abs err %.3e rad, rms %.3e rad, atan2(0,0)==0: %d\n", n, maxe, sqrt(sum2 / n), zero_ok);
  return (maxe < 4e-7 && sqrt(sum2 / n) < 1e-7 && zero_ok) ? 0 : 1;
}

int main() {
  if (check_poly_atan2()) return 2;
  srand(1); long bad0 = 0, bad1 = 0, n = 0; double maxd = 0;
  for (int it = 0; it < 4000000; it++) {
    float y = ((float)rand()/RAND_MAX - 0.5f) * powf(10.f, (rand()%8) - 6), x = ((float)rand()/RAND_MAX - 0.5f) * powf(10.f, (rand()%8) - 6);
    if (it % 1000 == 0) y = 0; if (it % 1777 == 0) x = 0; if (it % 5003 == 0) y = x; if (it % 7001 == 0) y = -x;
    float a = ref_atan2(y, x), b = bf_atan2(y, x, 1), c = bf_atan2(y, x, 0);
    if (memcmp(&a,&b,4)) { if (bad0 < 5) printf("exactdiv mismatch y=%g x=%g ref=%.9g bf=%.9g\n", y, x, a, b); bad0++; }
    if (memcmp(&a,&c,4)) { bad1++; double d = fabs((double)a-c); if (d > maxd) maxd = d; }
    n++;
  }
  printf("n=%ld exact-division form mismatches=%ld; approx-rcp(+-1ulp) form mismatches=%ld (max abs diff %.3g)\n", n, bad0, bad1, maxd);
  return (bad0 != 0 || bad1 > n / 100000 || maxd > 1e-6) ? 1 : 0;
}
