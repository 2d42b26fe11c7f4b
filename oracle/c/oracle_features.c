/*
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU restatement of the reference's
 * feature functions, used as the parity checker for the CUDA path and as the "port" CPU
 * baseline in bench.py.  Never linked into or called by the product library.
 *
 * Restates, from the behaviour documented in SURVEY.md appendix A/B:
 *   get_energy          reference describealign.py:545-555
 *   get_zero_crossings  reference describealign.py:557-566
 *   downsample_blur     reference describealign.py:568-573
 *   get_freq_bands      reference describealign.py:575-593
 * including the summation orders of the numpy / OpenBLAS routines the reference calls
 *  and glibc 2.39 log10f, so that the f32 features are
 * bit-identical to the reference run with numpy's AVX-512 math dispatch disabled
 * ("portable" oracle mode).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- int16 -> float16 (round to nearest even) -> float32: describealign.py:156 ---- */
static inline float s16_as_f16(int16_t s) {
  float f = (float)s;
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&f, &u, 4);
  return f;
}

/* f16((f32 l + f32 r) / 2): np.mean over axis 0 of a float16 array (describealign.py:576).
 * numpy accumulates float16 means in float32 and rounds the quotient back to float16. */
static inline float round_to_f16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&f, &u, 4);
  return f;
}

/* ---- glibc 2.39 log10f for x >= 1 ---- */
static const double LOGF_INVC[16] = {
    0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0,  0x1.3c995b0b80385p+0,
    0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0,  0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
    0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
    0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
static const double LOGF_LOGC[16] = {
    -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
    -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c81p-3,   -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
    -0x1.252f438e10c1ep-5, 0x0p+0,                0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
    0x1.526e57720db08p-3,  0x1.bc2860d22477p-3,   0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2};

static float glibc_logf(float x) {
  uint32_t ix;
  memcpy(&ix, &x, 4);
  if (ix == 0x3f800000u) return 0.0f;
  uint32_t tmp = ix - 0x3f330000u;
  int i = (int)((tmp >> 19) & 15u);
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  float zf;
  memcpy(&zf, &iz, 4);
  double z = (double)zf;
  double r = z * LOGF_INVC[i] - 1.0;
  double y0 = LOGF_LOGC[i] + (double)k * 0x1.62e42fefa39efp-1;
  double r2 = r * r;
  double p = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
  p = -0x1.00ea348b88334p-2 * r2 + p;
  p = p * r2 + (y0 + r);
  return (float)p;
}

/* "native numpy" mode: tests may route log10f through the host's np.log10 to check the
 * device's correction table; NULL = glibc's formula below. */
static float (*g_log10f_hook)(float) = 0;
void oracle_set_log10f_hook(float (*fn)(float)) { g_log10f_hook = fn; }

float oracle_log10f(float x) {
  /* glibc e_log10f.c for positive normal x */
  if (g_log10f_hook) return g_log10f_hook(x);
  uint32_t hx;
  memcpy(&hx, &x, 4);
  int32_t k = (int32_t)(hx >> 23) - 127;
  uint32_t i = ((uint32_t)k & 0x80000000u) >> 31;
  hx = (hx & 0x007fffffu) | ((0x7fu - i) << 23);
  float y = (float)(k + (int32_t)i);
  float m;
  memcpy(&m, &hx, 4);
  float z = y * 7.9034151668e-07f + 4.3429449201e-01f * glibc_logf(m);
  return z + y * 3.0102920532e-01f;
}

/* ---- convolution kernels ------------------------------------------------------- */
/* 'same' convolution with a symmetric-length window, >= 12 taps (SURVEY.md B.2 iii):
 * f32 products, sequential f64 accumulation over ascending signal index, one rounding. */
static void conv_same_f32_long(const float *a, int64_t n, const float *w, int taps, float *out) {
  int half = taps / 2; /* odd taps */
  for (int64_t t = 0; t < n; ++t) {
    double acc = 0.0;
    for (int j = 0; j < taps; ++j) {
      int64_t idx = t - half + j;
      if (idx < 0 || idx >= n) continue;
      float prod = a[idx] * w[taps - 1 - j];
      acc += (double)prod;
    }
    out[t] = (float)acc;
  }
}

/* ---- get_energy ------------------------------------------------------------------ */
/* pcm: interleaved int16, S samples per channel.  out length = ceil((S/105)/2). */
int64_t oracle_energy(const int16_t *pcm, int64_t S, int ch, const float *hann13, float *out) {
  int64_t nb = S / 105;
  float *e = (float *)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
  float *es = (float *)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
  int cnt = 105 * ch;
  float denom = (float)cnt;
  for (int64_t b = 0; b < nb; ++b) {
    const int16_t *p = pcm + b * cnt;
    float l[4] = {0.f, 0.f, 0.f, 0.f};
    int k = 0;
    /* numpy einsum sum-of-products, contiguous two-operand float kernel (SURVEY.md B.2 i) */
    for (; k + 16 <= cnt; k += 16) {
      for (int c4 = 3; c4 >= 0; --c4)
        for (int q = 0; q < 4; ++q) {
          float x = s16_as_f16(p[k + 4 * c4 + q]);
          float pr = x * x;
          l[q] = l[q] + pr;
        }
    }
    for (; k < cnt; k += 4) {
      for (int q = 0; q < 4; ++q) {
        float x = (k + q < cnt) ? s16_as_f16(p[k + q]) : 0.0f;
        float pr = x * x;
        l[q] = l[q] + pr;
      }
    }
    float s = (l[0] + l[1]) + (l[2] + l[3]);
    e[b] = s / denom;
  }
  conv_same_f32_long(e, nb, hann13, 13, es);
  int64_t m = 0;
  for (int64_t b = 0; b < nb; b += 2) {
    float v = 1.0f + es[b];
    out[m++] = oracle_log10f(v) / 2.0f;
  }
  free(e);
  free(es);
  return m;
}

/* ---- get_zero_crossings ---------------------------------------------------------- */
int64_t oracle_zero_crossings(const int16_t *pcm, int64_t S, int ch, const float *hann13, float *out) {
  int64_t L = S / 210;
  float *z = (float *)malloc(sizeof(float) * (size_t)(L > 0 ? L : 1));
  for (int64_t f = 0; f < L; ++f) {
    int count = 0;
    for (int c = 0; c < ch; ++c) {
      for (int k = 0; k < 210; ++k) {
        int64_t n = f * 210 + k;
        int sb = pcm[n * ch + c] < 0;
        int prev = (n == 0) ? 0 : (pcm[(n - 1) * ch + c] < 0);
        count += sb ^ prev;
      }
    }
    float v = (float)count;
    if (ch == 1) v = v * 2.0f;
    z[f] = v;
  }
  conv_same_f32_long(z, L, hann13, 13, out);
  free(z);
  return L;
}

/* ---- downsample_blur, f32, b taps per phase ------------------------------------------ */
/* phases with <= 11 taps: f32 accumulator (SURVEY.md B.2 ii); >= 12: f64 accumulator of f32
 * products (B.2 iii).  Phases summed sequentially in f32 starting from 0 (B.2 v). */
static void ds_blur_f32(const float *a, int64_t n_in, int d, int b, const float *w, float *out) {
  int64_t n = n_in / d;
  int half = b / 2;
  for (int64_t t = 0; t < n; ++t) {
    float total = 0.0f;
    for (int p = 0; p < d; ++p) {
      float ph;
      if (b <= 11) {
        float acc = 0.0f;
        for (int j = 0; j < b; ++j) {
          int64_t idx = t - half + j;
          if (idx < 0 || idx >= n) continue;
          float prod = a[idx * d + p] * w[p + (b - 1 - j) * d];
          acc = acc + prod;
        }
        ph = acc;
      } else {
        double acc = 0.0;
        for (int j = 0; j < b; ++j) {
          int64_t idx = t - half + j;
          if (idx < 0 || idx >= n) continue;
          float prod = a[idx * d + p] * w[p + (b - 1 - j) * d];
          acc += (double)prod;
        }
        ph = (float)acc;
      }
      total = total + ph;
    }
    out[t] = total;
  }
}

/* ---- get_freq_bands ---------------------------------------------------------------- */
/* windows: w15 (15 taps), w21, w630, w90: normalised f32 Hann windows. */
int64_t oracle_freq_bands(const int16_t *pcm, int64_t S, int ch, const float *w15, const float *w21,
                          const float *w630, const float *w90, float *b0, float *b1, double *b2) {
  int64_t L = S / 210;
  int64_t n = L * 210;
  if (L == 0) return 0;
  float *m = (float *)malloc(sizeof(float) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (ch == 1) {
      m[i] = s16_as_f16(pcm[i]);
    } else {
      float l = s16_as_f16(pcm[2 * i]), r = s16_as_f16(pcm[2 * i + 1]);
      float s = l + r;
      m[i] = round_to_f16(s / 2.0f);
    }
  }
  int64_t n1 = n / 5, n2 = n1 / 7;
  float *lp1 = (float *)malloc(sizeof(float) * (size_t)n1);
  float *lp2 = (float *)malloc(sizeof(float) * (size_t)n2);
  float *be = (float *)malloc(sizeof(float) * (size_t)n1);
  float *sm = (float *)malloc(sizeof(float) * (size_t)L);
  ds_blur_f32(m, n, 5, 3, w15, lp1);
  for (int64_t t = 0; t < n1; ++t) {
    float acc = 0.0f;
    for (int i = 0; i < 5; ++i) {
      float dlt = m[5 * t + i] - lp1[t];
      float sq = dlt * dlt;
      acc = (i == 0) ? sq : acc + sq;
    }
    be[t] = acc;
  }
  ds_blur_f32(be, n1, 42, 15, w630, sm);
  for (int64_t t = 0; t < L; ++t) {
    float v = sm[t] / 210.0f;
    b0[t] = oracle_log10f(1.0f + v) / 2.0f;
  }
  ds_blur_f32(lp1, n1, 7, 3, w21, lp2);
  for (int64_t t = 0; t < n2; ++t) {
    float acc = 0.0f;
    for (int i = 0; i < 7; ++i) {
      float dlt = lp1[7 * t + i] - lp2[t];
      float sq = dlt * dlt;
      acc = (i == 0) ? sq : acc + sq;
    }
    be[t] = acc;
  }
  ds_blur_f32(be, n2, 6, 15, w90, sm);
  for (int64_t t = 0; t < L; ++t) {
    float v = sm[t] / 210.0f;
    b1[t] = oracle_log10f(1.0f + v) / 2.0f;
  }
  /* band 2 in f64 (the int64 zero subtrahend promotes, describealign.py:583,588) */
  double *be2 = (double *)malloc(sizeof(double) * (size_t)L);
  for (int64_t t = 0; t < L; ++t) {
    double acc = 0.0;
    for (int i = 0; i < 6; ++i) {
      double x = (double)lp2[6 * t + i];
      double sq = x * x;
      acc = (i == 0) ? sq : acc + sq;
    }
    be2[t] = acc;
  }
  /* f64 convolution, 15 taps: OpenBLAS ddot tail = sequential FMA chain (SURVEY.md B.2 iv) */
  for (int64_t t = 0; t < L; ++t) {
    double acc = 0.0;
    for (int j = 0; j < 15; ++j) {
      int64_t idx = t - 7 + j;
      if (idx < 0 || idx >= L) continue;
      acc = fma((double)w15[14 - j], be2[idx], acc);
    }
    b2[t] = log10(1.0 + acc / 210.0) / 2.0;
  }
  free(m); free(lp1); free(lp2); free(be); free(sm); free(be2);
  return L;
}
