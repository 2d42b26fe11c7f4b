// Corridor planning on the device (reference describealign.py:895-932): per line cluster, the audio-row
// range that gets scored and the optional sub-frame refinement of the line's offset (:912-930).
// Included by stage_b.cu inside its anonymous namespace.
//
//   x_limits (:895-900)      rows where the line stays inside both tracks, optionally widened by 30 s
//   refinement (:916-930)    over the cluster's own rows: err = audio - lerp(video), kept where the mean
//                            over the three features is < 0.1; a one-parameter least-squares fit of err
//                            against the central difference of the interpolated video; the offset moves
//                            by the coefficient when the fit explains enough (sigmas > 8, |coef| < 2)
//
// The reference solves the fit with np.linalg.lstsq (LAPACK gelsd, whose Householder / dnrm2 steps are
// not reproducible bit for bit outside that library); here the three sums sum(dv*err), sum(dv*dv),
// sum(err*err) are accumulated in float64 in a fixed order and the coefficient is their quotient: equal
// to the reference's to ~1e-15 relative, not bit-identical.  The decisions (kept rows, counts, thresholds)
// use the reference's expressions and are identical except on exact ties.
constexpr int RF_THREADS = 256;
constexpr int RF_CHUNK = 2048;         // interior rows per block

struct RefineArgs {
  const float *a_scaled;   // (n_a, 3)
  const float *v_scaled;   // (n_v, 3)
  int64_t n_a, n_v;
  const dab_cluster *cl;
  int32_t n_cl;
  int32_t max_blocks;      // blocks per cluster the launch provides
  double *partial;         // [n_cl][max_blocks][4]: count, sum dv*err, sum dv*dv, sum err*err
  dab_corridor *cor;       // out: one corridor per cluster (empty when the cluster is not scored)
};

// int(np.ceil(x)) / int(np.floor(x)) / int(x) of the reference, clamped into int32
__device__ __forceinline__ long long rf_to_ll(double x) {
  if (!(x > -9.0e15)) return -9000000000000000LL;
  if (!(x < 9.0e15)) return 9000000000000000LL;
  return (long long)x;
}

// describealign.py:895-900.  Half-open on neither side: returns lo, hi as the reference does.
__device__ __forceinline__ void rf_x_limits(double x_first, double x_last, double offset, double slope, int64_t n_a, int64_t n_v,
                                            long long extend, long long &lo, long long &hi) {
  lo = rf_to_ll(x_first) - extend;
  if (lo < 0) lo = 0;
  hi = rf_to_ll(x_last) + extend;
  if (hi > n_a - 1) hi = n_a - 1;
  const long long l2 = rf_to_ll(ceil((4.0 - offset) / slope));
  const long long h2 = rf_to_ll(floor(((double)(n_v - 4) - offset) / slope));
  if (l2 > lo) lo = l2;
  if (h2 < hi) hi = h2;
}

// degree-1 interpolation of the video rows at y, the float64 form of scipy's make_interp_spline(k=1)
__device__ __forceinline__ void rf_lerp(const float *v_scaled, double y, double out[3]) {
  const double fl = floor(y);
  const int64_t f = (int64_t)fl;
  const double t = y - fl, omt = 1.0 - t;
  const float *v0 = v_scaled + f * 3, *v1 = v0 + 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[c] = __dadd_rn(__dmul_rn((double)v0[c], omt), __dmul_rn((double)v1[c], t));
}

__global__ void __launch_bounds__(RF_THREADS) refine_partial_kernel(RefineArgs a) {
  __shared__ double red[4][RF_THREADS / 32];
  const int c = blockIdx.y;
  const dab_cluster cl = a.cl[c];
  long long lo, hi;
  rf_x_limits(cl.x_first, cl.x_last, cl.offset, cl.slope, a.n_a, a.n_v, 0, lo, hi);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (hi > lo + 100) {
    // rows lo .. hi-1 are interpolated; the fit uses the interior rows lo+1 .. hi-2
    const long long first = lo + 1, count = hi - lo - 2;
    const long long begin = (long long)blockIdx.x * RF_CHUNK;
    const long long end = begin + RF_CHUNK < count ? begin + RF_CHUNK : count;
    for (long long m = begin + threadIdx.x; m < end; m += RF_THREADS) {
      const long long r = first + m;
      double vm[3], vp[3], vn[3];
      rf_lerp(a.v_scaled, __dadd_rn(__dmul_rn(cl.slope, (double)r), cl.offset), vm);
      rf_lerp(a.v_scaled, __dadd_rn(__dmul_rn(cl.slope, (double)(r - 1)), cl.offset), vp);
      rf_lerp(a.v_scaled, __dadd_rn(__dmul_rn(cl.slope, (double)(r + 1)), cl.offset), vn);
      const float *ar = a.a_scaled + r * 3;
      const double e0 = (double)ar[0] - vm[0], e1 = (double)ar[1] - vm[1], e2 = (double)ar[2] - vm[2];
      const double mean = __ddiv_rn(__dadd_rn(__dadd_rn(e0, e1), e2), 3.0);     // np.mean over the last axis
      if (mean < 0.1) {
        const double d0 = (vn[0] - vp[0]) / 2.0, d1 = (vn[1] - vp[1]) / 2.0, d2 = (vn[2] - vp[2]) / 2.0;
        acc[0] += 1.0;
        acc[1] += d0 * e0 + d1 * e1 + d2 * e2;
        acc[2] += d0 * d0 + d1 * d1 + d2 * d2;
        acc[3] += e0 * e0 + e1 * e1 + e2 * e2;
      }
    }
  }
  // fixed-order reduction: lanes by xor butterfly, warps in index order
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[q][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int x = 0; x < RF_THREADS / 32; ++x) v += red[threadIdx.x][x];
    a.partial[((int64_t)c * a.max_blocks + blockIdx.x) * 4 + threadIdx.x] = v;
  }
}

// one thread per cluster: the fit, the decision, the final row range
__global__ void refine_plan_kernel(RefineArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_cl) return;
  const dab_cluster cl = a.cl[c];
  dab_corridor out;
  out.cluster = cl.cluster; out.lo = 0; out.hi = 0; out.reserved = 0; out.slope = cl.slope; out.offset = cl.offset;
  long long lo, hi;
  rf_x_limits(cl.x_first, cl.x_last, cl.offset, cl.slope, a.n_a, a.n_v, 0, lo, hi);
  if (!(hi < lo + 5)) {
    double x_first = cl.x_first, x_last = cl.x_last, offset = cl.offset;
    if (hi > lo + 100) {
      const long long count = hi - lo - 2;
      const long long nb = (count + RF_CHUNK - 1) / RF_CHUNK;
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (long long b = 0; b < nb && b < a.max_blocks; ++b)
        for (int q = 0; q < 4; ++q) s[q] += a.partial[((int64_t)c * a.max_blocks + b) * 4 + q];
      if (s[0] > 50.0 && s[2] > 0.0) {
        const double coef = s[1] / s[2];
        const double resid = s[3] - coef * s[1];
        const double explained = 1.0 - resid / s[3];
        const double sigmas = sqrt(explained * (3.0 * s[0])) - 1.0;
        if (sigmas > 8.0 && fabs(coef) < 2.0) offset = offset + coef;
      }
      x_first = (double)lo;            // rows[0], rows[-1] of np.arange(lo, hi)
      x_last = (double)(hi - 1);
    }
    long long lo2, hi2;
    rf_x_limits(x_first, x_last, offset, cl.slope, a.n_a, a.n_v, 6300, lo2, hi2);
    out.offset = offset;
    if (hi2 > lo2) { out.lo = (int32_t)lo2; out.hi = (int32_t)hi2; }
  }
  a.cor[c] = out;
}

// np.max of the energy columns of the scaled arrays (describealign.py:908-909): block maxima combined with
// an atomic max on an order-preserving integer image of the float; decoded by refine_plan_kernel / the
// decode kernel below into out[0] (audio), out[1] (video).
__device__ __forceinline__ unsigned int rf_ordered(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float rf_unordered(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void column_max_kernel(const float *a_scaled, int64_t n_a, const float *v_scaled, int64_t n_v, unsigned int *keys) {
  __shared__ float red[32];
  const float *src = blockIdx.y == 0 ? a_scaled : v_scaled;
  const int64_t n = blockIdx.y == 0 ? n_a : n_v;
  float m = -INFINITY;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, src[3 * k]);
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int x = 1; x < (int)(blockDim.x >> 5); ++x) m = fmaxf(m, red[x]);
    atomicMax(keys + blockIdx.y, rf_ordered(m));
  }
}

__global__ void column_max_decode_kernel(const unsigned int *keys, float *out) {
  if (threadIdx.x < 2) out[threadIdx.x] = rf_unordered(keys[threadIdx.x]);
}
