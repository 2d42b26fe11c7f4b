// Feature extraction for one track: the five 210 Hz feature vectors of the reference
// (get_energy describealign.py:545-555, get_zero_crossings :557-566, downsample_blur
// :568-573, get_freq_bands :575-593) in ONE pass over the PCM.
//
// One CTA produces FT consecutive output frames.  It stages the samples of frames
// [t0-7, t0+FT+7) plus 40 samples of filter halo in shared memory as float16 (the
// reference's own sample type, :156), derives every intermediate of the polyphase filter
// bank (lp1 @8820 Hz, lp2 @1260 Hz, the three residual band energies, 105-sample block
// energies, per-frame zero-crossing counts) in shared memory, and writes only the five
// outputs.  HBM traffic: the PCM once (+ halo re-reads, served from L2) and 24 B per frame.
//
// Bit-exactness: every sum is evaluated in the order numpy / OpenBLAS
// use in the reference - parallelism is across outputs, never across the taps of one
// output - with separately rounded multiplies and adds (-fmad=false), f32 products
// accumulated in f64 where numpy does, and glibc's log10f restated in IEEE operations.
#include <cuda_fp16.h>

#include "common.cuh"
#include "hann_tables.h"

namespace {

constexpr int FT = 96;            // output frames per CTA
constexpr int FH = 7;             // halo frames each side (15-tap smoothing at the frame rate)
constexpr int NF = FT + 2 * FH;   // frames staged
constexpr int PADS = 40;          // halo samples each side (lp2 needs lp1 +-7, lp1 needs +-5 samples)
constexpr int NS = NF * 210 + 2 * PADS;
constexpr int N1 = NF * 42 + 14;  // lp1 entries staged (+-7)
constexpr int THREADS = 512;

__constant__ float c_w13[13];
__constant__ float c_w15[15];
__constant__ float c_w21[21];
__constant__ float c_w90[90];
__constant__ float c_w630[630];
__constant__ double c_logf_invc[16];
__constant__ double c_logf_logc[16];

const double h_logf_invc[16] = {
    0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0,  0x1.3c995b0b80385p+0,
    0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0,  0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
    0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
    0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
const double h_logf_logc[16] = {
    -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
    -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c81p-3,   -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
    -0x1.252f438e10c1ep-5, 0x0p+0,                0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
    0x1.526e57720db08p-3,  0x1.bc2860d22477p-3,   0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2};

// glibc 2.39 logf / log10f for x >= 1 in plain IEEE operations.
__device__ __forceinline__ float glibc_logf(float x) {
  uint32_t ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> 19) & 15;
  int k = (int)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  double z = (double)__uint_as_float(iz);
  double r = z * c_logf_invc[i] - 1.0;
  double y0 = c_logf_logc[i] + (double)k * 0x1.62e42fefa39efp-1;
  double r2 = r * r;
  double p = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
  p = -0x1.00ea348b88334p-2 * r2 + p;
  p = p * r2 + (y0 + r);
  return (float)p;
}

__device__ __forceinline__ float glibc_log10f(float x) {
  uint32_t hx = __float_as_uint(x);
  int k = (int)(hx >> 23) - 127;
  uint32_t i = ((uint32_t)k & 0x80000000u) >> 31;
  hx = (hx & 0x007fffffu) | ((0x7fu - i) << 23);
  float y = (float)(k + (int)i);
  float m = __uint_as_float(hx);
  float z = y * 7.9034151668e-07f + 4.3429449201e-01f * glibc_logf(m);
  return z + y * 3.0102920532e-01f;
}

template <int FMT>
__device__ __forceinline__ float sample_f32(const void *pcm, int64_t idx) {
  if (FMT == DAB_PCM_S16) {
    short s = reinterpret_cast<const short *>(pcm)[idx];
    return __half2float(__short2half_rn(s));   // int16 -> float16 (RNE) as describealign.py:156
  } else {
    return __half2float(reinterpret_cast<const __half *>(pcm)[idx]);
  }
}

template <int FMT>
__device__ __forceinline__ bool sample_neg(const void *pcm, int64_t idx) {
  // np.signbit on the float16 array (:558); -0.0 cannot come out of an int16 conversion but an
  // F16 caller may pass it, so test the sign bit, not "< 0".
  return (reinterpret_cast<const unsigned short *>(pcm)[idx] & 0x8000u) != 0;
}

struct FeatArgs {
  const void *pcm;
  int64_t S;      // samples per channel
  int64_t L;      // S / 210
  int64_t nb;     // S / 105
  int64_t Le;     // ceil(nb / 2)
  float *energy, *zc, *b0, *b1;
  double *b2;
};

template <int FMT, int CH>
__global__ void __launch_bounds__(THREADS) features_kernel(FeatArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half *sig = reinterpret_cast<__half *>(smem_raw);                 // NS (padded to even)
  float *lp1 = reinterpret_cast<float *>(sig + ((NS + 7) & ~7));      // N1
  float *be0 = lp1 + N1;                                               // NF*42
  float *ph = be0 + NF * 42;                                           // FT*42 phase sums for band 0
  float *lp2 = ph + FT * 42;                                           // NF*6
  float *be1 = lp2 + NF * 6;                                           // NF*6
  float *eb = be1 + NF * 6;                                            // 2*NF block energies
  float *zf = eb + 2 * NF;                                             // NF zero-crossing counts
  double *be2 = reinterpret_cast<double *>(zf + NF + (NF & 1));        // NF

  const int tid = threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.x * FT;          // first output frame of this tile
  const int64_t f0 = t0 - FH;                           // first staged frame (may be < 0)
  const int64_t s0 = f0 * 210 - PADS;                   // first staged sample (may be < 0)
  const int64_t Sb = a.L * 210;                         // band signal length (:577)

  // ---- stage the mono / mid signal as float16, zero outside [0, Sb) ----------------------
  for (int k = tid; k < NS; k += THREADS) {
    int64_t g = s0 + k;
    __half h = __float2half_rn(0.0f);
    if (g >= 0 && g < Sb) {
      if (CH == 1) {
        h = __float2half_rn(sample_f32<FMT>(a.pcm, g));  // exact: value is already a float16
      } else {
        float l = sample_f32<FMT>(a.pcm, 2 * g), r = sample_f32<FMT>(a.pcm, 2 * g + 1);
        float s = l + r;                                 // np.mean over 2 float16 channels: f32
        h = __float2half_rn(s / 2.0f);                   // accumulate, divide, round to float16
      }
    }
    sig[k] = h;
  }

  // ---- block energies (einsum order, SURVEY.md B.2 i) and zero-crossing counts, from global --
  for (int k = tid; k < 2 * NF; k += THREADS) {
    int64_t b = 2 * f0 + k;
    float e = 0.0f;
    if (b >= 0 && b < a.nb) {
      const int cnt = 105 * CH;
      const int64_t base = b * cnt;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      int q = 0;
      for (; q + 16 <= cnt; q += 16) {
#pragma unroll
        for (int c4 = 3; c4 >= 0; --c4) {
          float x0 = sample_f32<FMT>(a.pcm, base + q + 4 * c4 + 0);
          float x1 = sample_f32<FMT>(a.pcm, base + q + 4 * c4 + 1);
          float x2 = sample_f32<FMT>(a.pcm, base + q + 4 * c4 + 2);
          float x3 = sample_f32<FMT>(a.pcm, base + q + 4 * c4 + 3);
          l0 = l0 + x0 * x0; l1 = l1 + x1 * x1; l2 = l2 + x2 * x2; l3 = l3 + x3 * x3;
        }
      }
      for (; q < cnt; q += 4) {
        float x0 = (q + 0 < cnt) ? sample_f32<FMT>(a.pcm, base + q + 0) : 0.f;
        float x1 = (q + 1 < cnt) ? sample_f32<FMT>(a.pcm, base + q + 1) : 0.f;
        float x2 = (q + 2 < cnt) ? sample_f32<FMT>(a.pcm, base + q + 2) : 0.f;
        float x3 = (q + 3 < cnt) ? sample_f32<FMT>(a.pcm, base + q + 3) : 0.f;
        l0 = l0 + x0 * x0; l1 = l1 + x1 * x1; l2 = l2 + x2 * x2; l3 = l3 + x3 * x3;
      }
      e = ((l0 + l1) + (l2 + l3)) / (float)cnt;
    }
    eb[k] = e;
  }
  for (int k = tid; k < NF; k += THREADS) {
    int64_t f = f0 + k;
    float z = 0.0f;
    if (f >= 0 && f < a.L) {
      int count = 0;
      for (int c = 0; c < CH; ++c) {
        int64_t n = f * 210;
        bool prev = (n == 0) ? false : sample_neg<FMT>(a.pcm, (n - 1) * CH + c);
        for (int q = 0; q < 210; ++q) {
          bool cur = sample_neg<FMT>(a.pcm, (n + q) * CH + c);
          count += (cur != prev);
          prev = cur;
        }
      }
      z = (float)count;
      if (CH == 1) z = z * 2.0f;
    }
    zf[k] = z;
  }
  __syncthreads();

  // ---- lp1 = downsample_blur(m, 5, 3): 5 phases x 3 taps, f32 accumulators (B.2 ii) ---------
  const int64_t n1_first = f0 * 42 - 7;          // global lp1 index of lp1[0]
  const int64_t len1 = a.L * 42;
  for (int k = tid; k < N1; k += THREADS) {
    int64_t n = n1_first + k;
    float total = 0.0f;
    if (n >= 0 && n < len1) {
      // sample index of m[(n-1+j)*5 + p] relative to the staged window
      const int rel = (int)((n - 1) * 5 - s0);
#pragma unroll
      for (int p = 0; p < 5; ++p) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          // taps outside [0, len1) are zero padding of the phase signal
          int64_t nn = n - 1 + j;
          float x = (nn >= 0 && nn < len1) ? __half2float(sig[rel + j * 5 + p]) : 0.0f;
          acc = acc + x * c_w15[p + (2 - j) * 5];
        }
        total = total + acc;
      }
    }
    lp1[k] = total;
  }
  __syncthreads();

  // ---- band-0 residual energy at 8820 Hz and lp2 = downsample_blur(lp1, 7, 3) ----------------
  for (int k = tid; k < NF * 42; k += THREADS) {
    int64_t n = f0 * 42 + k;
    float acc = 0.0f;
    if (n >= 0 && n < len1) {
      const float lo = lp1[k + 7];
      const int rel = (int)(n * 5 - s0);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        float d = __half2float(sig[rel + i]) - lo;
        float sq = d * d;
        acc = (i == 0) ? sq : acc + sq;
      }
    }
    be0[k] = acc;
  }
  const int64_t len2 = a.L * 6;
  for (int k = tid; k < NF * 6; k += THREADS) {
    int64_t n2 = f0 * 6 + k;
    float total = 0.0f;
    if (n2 >= 0 && n2 < len2) {
      const int rel = (int)((n2 - 1) * 7 - n1_first);
#pragma unroll
      for (int p = 0; p < 7; ++p) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          int64_t nn = n2 - 1 + j;
          float x = (nn >= 0 && nn < len2) ? lp1[rel + j * 7 + p] : 0.0f;
          acc = acc + x * c_w21[p + (2 - j) * 7];
        }
        total = total + acc;
      }
    }
    lp2[k] = total;
  }
  __syncthreads();

  // ---- band-1 residual energy at 1260 Hz, band-2 energy per frame (f64, :583/:588) ----------
  for (int k = tid; k < NF * 6; k += THREADS) {
    int64_t n2 = f0 * 6 + k;
    float acc = 0.0f;
    if (n2 >= 0 && n2 < len2) {
      const float lo = lp2[k];
      const int rel = (int)(n2 * 7 - n1_first);
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        float d = lp1[rel + i] - lo;
        float sq = d * d;
        acc = (i == 0) ? sq : acc + sq;
      }
    }
    be1[k] = acc;
  }
  for (int k = tid; k < NF; k += THREADS) {
    int64_t f = f0 + k;
    double acc = 0.0;
    if (f >= 0 && f < a.L) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double x = (double)lp2[k * 6 + i];
        double sq = x * x;
        acc = (i == 0) ? sq : acc + sq;
      }
    }
    be2[k] = acc;
  }
  __syncthreads();

  // ---- band 0: 42 phases x 15 taps; each phase = f32 products accumulated in f64 (B.2 iii) ----
  for (int k = tid; k < FT * 42; k += THREADS) {
    const int t = k / 42, p = k - t * 42;
    // output frame t0 + t uses be0 frames (t0 + t - 7 + j), j = 0..14 -> staged frame index t + j
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 15; ++j) {
      float prod = be0[(t + j) * 42 + p] * c_w630[p + (14 - j) * 42];
      acc += (double)prod;
    }
    ph[k] = (float)acc;
  }
  __syncthreads();

  // ---- outputs --------------------------------------------------------------------------------
  for (int t = tid; t < FT; t += THREADS) {
    const int64_t f = t0 + t;
    if (f < a.L) {
      // band 0: phases added sequentially in f32 (B.2 v), /210, log10(1+x)/2
      float tot = 0.0f;
      for (int p = 0; p < 42; ++p) tot = tot + ph[t * 42 + p];
      a.b0[f] = glibc_log10f(1.0f + tot / 210.0f) / 2.0f;
      // band 1: 6 phases x 15 taps
      tot = 0.0f;
      for (int p = 0; p < 6; ++p) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 15; ++j) {
          float prod = be1[(t + j) * 6 + p] * c_w90[p + (14 - j) * 6];
          acc += (double)prod;
        }
        tot = tot + (float)acc;
      }
      a.b1[f] = glibc_log10f(1.0f + tot / 210.0f) / 2.0f;
      // band 2: one 15-tap f64 filter = OpenBLAS ddot tail, a sequential FMA chain (B.2 iv)
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 15; ++j) acc = fma((double)c_w15[14 - j], be2[t + j], acc);
      a.b2[f] = log10(1.0 + acc / 210.0) / 2.0;
      // zero crossings: 13-tap Hann over frames f-6 .. f+6
      acc = 0.0;
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        float prod = zf[t + FH - 6 + j] * c_w13[12 - j];
        acc += (double)prod;
      }
      a.zc[f] = (float)acc;
    }
    if (f < a.Le) {
      // energy: 13-tap Hann over blocks 2f-6 .. 2f+6, log10(1+x)/2, every second block (:553-555)
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        float prod = eb[2 * (t + FH) - 6 + j] * c_w13[12 - j];
        acc += (double)prod;
      }
      a.energy[f] = glibc_log10f(1.0f + (float)acc) / 2.0f;
    }
  }
}

size_t feat_smem_bytes() {
  size_t b = (size_t)((NS + 7) & ~7) * 2;
  b += sizeof(float) * (size_t)(N1 + NF * 42 + FT * 42 + NF * 6 + NF * 6 + 2 * NF + NF + (NF & 1));
  b += sizeof(double) * NF;
  return b;
}

bool g_const_ready[64] = {};

int upload_constants(dab_ctx *ctx) {
  int dev = 0;
  DAB_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && g_const_ready[dev]) return DAB_OK;
  DAB_CUDA(cudaMemcpyToSymbol(c_w13, DAB_HANN13_F32, sizeof(DAB_HANN13_F32)));
  DAB_CUDA(cudaMemcpyToSymbol(c_w15, DAB_HANN15_F32, sizeof(DAB_HANN15_F32)));
  DAB_CUDA(cudaMemcpyToSymbol(c_w21, DAB_HANN21_F32, sizeof(DAB_HANN21_F32)));
  DAB_CUDA(cudaMemcpyToSymbol(c_w90, DAB_HANN90_F32, sizeof(DAB_HANN90_F32)));
  DAB_CUDA(cudaMemcpyToSymbol(c_w630, DAB_HANN630_F32, sizeof(DAB_HANN630_F32)));
  DAB_CUDA(cudaMemcpyToSymbol(c_logf_invc, h_logf_invc, sizeof(h_logf_invc)));
  DAB_CUDA(cudaMemcpyToSymbol(c_logf_logc, h_logf_logc, sizeof(h_logf_logc)));
  if (dev < 64) g_const_ready[dev] = true;
  return DAB_OK;
}

template <int FMT, int CH>
int launch(dab_pair *pr, const FeatArgs &fa, int64_t tiles) {
  dab_ctx *ctx = pr->ctx;
  size_t smem = feat_smem_bytes();
  DAB_CUDA(cudaFuncSetAttribute(features_kernel<FMT, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  features_kernel<FMT, CH><<<(unsigned)tiles, THREADS, smem, pr->stream>>>(fa);
  DAB_LAUNCHED(pr);
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

}  // namespace

int dab_run_features(dab_pair *pr, int track, const void *d_pcm, int format) {
  dab_ctx *ctx = pr->ctx;
  Track &tk = pr->trk[track];
  DAB_TRY(upload_constants(ctx));
  const int64_t L = tk.S / 210, nb = tk.S / 105, Le = (nb + 1) / 2;
  tk.L = L;
  tk.Le = Le;
  DAB_TRY(dab_ensure(ctx, tk.energy, sizeof(float) * (size_t)(Le + 1)));
  DAB_TRY(dab_ensure(ctx, tk.zc, sizeof(float) * (size_t)(L + 1)));
  DAB_TRY(dab_ensure(ctx, tk.b0, sizeof(float) * (size_t)(L + 1)));
  DAB_TRY(dab_ensure(ctx, tk.b1, sizeof(float) * (size_t)(L + 1)));
  DAB_TRY(dab_ensure(ctx, tk.b2, sizeof(double) * (size_t)(L + 1)));
  FeatArgs fa;
  fa.pcm = d_pcm; fa.S = tk.S; fa.L = L; fa.nb = nb; fa.Le = Le;
  fa.energy = tk.energy.as<float>(); fa.zc = tk.zc.as<float>();
  fa.b0 = tk.b0.as<float>(); fa.b1 = tk.b1.as<float>(); fa.b2 = tk.b2.as<double>();
  const int64_t tiles = cdiv(Le > L ? Le : L, FT);
  if (tiles > 0) {
    if (format == DAB_PCM_S16 && tk.ch == 1) DAB_TRY((launch<DAB_PCM_S16, 1>(pr, fa, tiles)));
    else if (format == DAB_PCM_S16 && tk.ch == 2) DAB_TRY((launch<DAB_PCM_S16, 2>(pr, fa, tiles)));
    else if (format == DAB_PCM_F16 && tk.ch == 1) DAB_TRY((launch<DAB_PCM_F16, 1>(pr, fa, tiles)));
    else if (format == DAB_PCM_F16 && tk.ch == 2) DAB_TRY((launch<DAB_PCM_F16, 2>(pr, fa, tiles)));
    else { ctx->err = "unsupported PCM format / channel count"; return DAB_E_ARG; }
  }
  tk.have_features = true;
  return DAB_OK;
}
