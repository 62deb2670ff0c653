// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Generic-C stand-in for the VOLK kernels the reference demodulation path calls
// (SURVEY.md Appendix C). VOLK is an un-vendored dependency of the reference
// (CMakeLists.txt:50, no version pin) and is absent from this image, so the reference
// sources under /root/reference are compiled against these scalar loops, which follow
// the published "generic" protokernel semantics of gnuradio/volk 3.x.
// Call sites in the reference: Utility.h:128-129,147-148; IfResampler.cpp:50;
// FmDecode.cpp:143,236; PhaseDiscriminator.cpp:40,42; MultipathFilter.cpp:102,123,125,153;
// AmDecode.cpp:190,225,233.
#ifndef ORACLE_VOLK_SHIM_H
#define ORACLE_VOLK_SHIM_H

#include <cmath>
#include <complex>

#define VOLK_VERSION_MAJOR 3
#define VOLK_VERSION_MINOR 2
#define VOLK_VERSION_MAINT 0
#define VOLK_VERSION 030200

typedef std::complex<float> lv_32fc_t;

static inline void volk_32fc_magnitude_squared_32f(float *o, const lv_32fc_t *a,
                                                   unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    const float re = a[i].real(), im = a[i].imag();
    o[i] = re * re + im * im;
  }
}

static inline void volk_32fc_magnitude_32f(float *o, const lv_32fc_t *a,
                                           unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    const float re = a[i].real(), im = a[i].imag();
    o[i] = sqrtf(re * re + im * im);
  }
}

static inline void volk_32f_accumulator_s32f(float *r, const float *a,
                                             unsigned int n) {
  float s = 0.0f;
  for (unsigned int i = 0; i < n; i++) {
    s += a[i];
  }
  *r = s;
}

static inline void volk_32f_x2_dot_prod_32f(float *r, const float *a,
                                            const float *b, unsigned int n) {
  float s = 0.0f;
  for (unsigned int i = 0; i < n; i++) {
    s += a[i] * b[i];
  }
  *r = s;
}

static inline void volk_32fc_x2_dot_prod_32fc(lv_32fc_t *r, const lv_32fc_t *a,
                                              const lv_32fc_t *b,
                                              unsigned int n) {
  float sr = 0.0f, si = 0.0f;
  for (unsigned int i = 0; i < n; i++) {
    const float ar = a[i].real(), ai = a[i].imag();
    const float br = b[i].real(), bi = b[i].imag();
    sr += ar * br - ai * bi;
    si += ar * bi + ai * br;
  }
  *r = lv_32fc_t(sr, si);
}

// c = a + conj(b) * s
static inline void volk_32fc_x2_s32fc_multiply_conjugate_add2_32fc(
    lv_32fc_t *c, const lv_32fc_t *a, const lv_32fc_t *b, const lv_32fc_t *s,
    unsigned int n) {
  const float sr = s->real(), si = s->imag();
  for (unsigned int i = 0; i < n; i++) {
    const float br = b[i].real(), bi = -b[i].imag();
    c[i] = lv_32fc_t(a[i].real() + (br * sr - bi * si),
                     a[i].imag() + (br * si + bi * sr));
  }
}

static inline void volk_32fc_s32f_atan2_32f(float *o, const lv_32fc_t *a,
                                            const float normalize_factor,
                                            unsigned int n) {
  const float inv = 1.0f / normalize_factor;
  for (unsigned int i = 0; i < n; i++) {
    o[i] = atan2f(a[i].imag(), a[i].real()) * inv;
  }
}

static inline void volk_32f_s32f_32f_fm_detect_32f(float *o, const float *in,
                                                   const float bound,
                                                   float *save,
                                                   unsigned int n) {
  if (n < 1) {
    return;
  }
  float prev = *save;
  for (unsigned int i = 0; i < n; i++) {
    float d = in[i] - prev;
    if (d > bound) {
      d -= 2 * bound;
    }
    if (d < -bound) {
      d += 2 * bound;
    }
    o[i] = d;
    prev = in[i];
  }
  *save = in[n - 1];
}

static inline void volk_32f_convert_64f(double *o, const float *a,
                                        unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    o[i] = (double)a[i];
  }
}

static inline void volk_64f_convert_32f(float *o, const double *a,
                                        unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    o[i] = (float)a[i];
  }
}

static inline void volk_64f_x2_multiply_64f(double *c, const double *a,
                                            const double *b, unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    c[i] = a[i] * b[i];
  }
}

static inline void volk_32fc_deinterleave_real_32f(float *o, const lv_32fc_t *a,
                                                   unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    o[i] = a[i].real();
  }
}

static inline void volk_32fc_deinterleave_64f_x2(double *re, double *im,
                                                 const lv_32fc_t *a,
                                                 unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    re[i] = (double)a[i].real();
    im[i] = (double)a[i].imag();
  }
}

#endif
