// Jump search of the reference's pitch-preserving time stretch (describealign.py:252-304 and :330-333; SURVEY.md
// section 8(f) N3): for every 512-sample window of a segment and every candidate jump distance, the position whose
// 512-sample Pearson correlation with the signal `jump` samples away is largest, and that correlation.
//
// The reference evaluates this with float64 running sums (np.cumsum) over PIECES of 51 windows that overlap by two,
// each piece with its own epsilon; a running sum is sequential by definition, so bit-identical results need the same
// adds in the same order.  That order is kept: one thread owns one (piece, jump) and walks the piece once, carrying
// the running sum twice - at the window's end and, by the identical sequence of adds 512 steps behind, at its start -
// so the windowed sum c[i] - c[i-512] needs no stored prefix array.  The parallelism is across pieces x jumps (a
// 22-minute segment: 2 300 pieces x 10 to 482 jumps), which is plenty.
//   stretch_energy_kernel   thread per piece: window energies (float64), their maximum -> epsilon, rms = sqrt(e + eps)
//   stretch_jump_kernel     thread per (piece, jump): correlations in position order, running first-maximum per window
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int SW = 512;                 // window (describealign.py:252)
constexpr int S_CACHED = 50;            // windows a piece yields (:259)
constexpr int S_CUT = S_CACHED * SW;

struct Piece {
  int64_t start;      // first sample
  int32_t len;        // samples
  int32_t lo;         // first yielded local window
  int32_t count;      // yielded windows
  int32_t first_out;  // global index of the first yielded window
};

struct StretchArgs {
  const __half *x;          // [ch][n]
  int64_t n;
  int ch;
  int negative;
  const Piece *pieces;
  int32_t n_pieces;
  const int32_t *jumps;
  int32_t n_jumps;
  double *rms;              // [n_pieces][rms_stride]
  double *eps;              // [n_pieces]
  int64_t rms_stride;
  int16_t *loc;             // [windows][n_jumps]
  double *best;             // [windows][n_jumps]
};

// np.sum(input.astype(np.float32) ** 2, axis=0)[i] and np.sum(input[:, jump:].astype(np.float32) * input[:, :n-jump], axis=0)[i]:
// float32 products (exact for float16 factors), channels added in float32
__device__ __forceinline__ float prod_at(const StretchArgs &a, int64_t i, int64_t k) {
  const float p0 = __half2float(a.x[k]) * __half2float(a.x[i]);
  if (a.ch == 1) return p0;
  const float p1 = __half2float(a.x[a.n + k]) * __half2float(a.x[a.n + i]);
  return p0 + p1;
}

__global__ void stretch_energy_kernel(StretchArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n_pieces) return;
  const Piece pc = a.pieces[p];
  const int n = pc.len;
  double *rms = a.rms + (int64_t)p * a.rms_stride;
  // window energies: c[t + 511] - c[t - 1] with c = sequential float64 running sum (np.cumsum)
  double lead = 0.0, lag = 0.0, mx = 0.0;
  for (int i = 0; i < SW - 1; ++i) lead = lead + (double)prod_at(a, pc.start + i, pc.start + i);
  for (int t = 0; t + SW <= n; ++t) {
    lead = lead + (double)prod_at(a, pc.start + t + SW - 1, pc.start + t + SW - 1);
    double e = lead;
    if (t > 0) {
      lag = lag + (double)prod_at(a, pc.start + t - 1, pc.start + t - 1);
      e = lead - lag;
    }
    rms[t] = e;
    mx = (t == 0 || e > mx) ? e : mx;
  }
  const double eps = 1e-4 * (mx > 1.0 ? mx : 1.0);       // 1e-4 * max(1, np.max(window_rms))
  a.eps[p] = eps;
  for (int t = 0; t + SW <= n; ++t) rms[t] = sqrt(rms[t] + eps);
}

__global__ void stretch_jump_kernel(StretchArgs a) {
  const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (int64_t)a.n_pieces * a.n_jumps) return;
  const int p = (int)(id / a.n_jumps), k = (int)(id - (int64_t)p * a.n_jumps);
  const Piece pc = a.pieces[p];
  const int n = pc.len, jump = a.jumps[k];
  const double *rms = a.rms + (int64_t)p * a.rms_stride;
  const double eps = a.eps[p];
  const int rows = n - SW + 1;                        // window start positions of the piece
  const int w_lo = pc.lo, w_hi = pc.lo + pc.count;    // the local windows this piece yields
  const double ninf = -__longlong_as_double(0x7ff0000000000000ll);
  const int n_u = n - jump - SW + 1;                  // pairs (u, u + jump) with both windows inside the piece
  double lead = 0.0, lag = 0.0;
  if (n_u > 0)
    for (int i = 0; i < SW - 1; ++i) lead = lead + (double)prod_at(a, pc.start + i, pc.start + i + jump);
  // positions in order; a position holds a value when its pair exists (negative jumps: the pair (t - jump, t) is
  // stored at t; positive: (t, t + jump) at t), -inf otherwise - np.argmax then returns the FIRST maximum of a window,
  // 0 for a window without any value
  double bv = ninf;
  int bl = 0;
  for (int t = 0; t < rows; ++t) {
    const int r = t & (SW - 1);
    if (r == 0) { bv = ninf; bl = 0; }
    const int u = a.negative ? t - jump : t;
    if (u >= 0 && u < n_u) {
      lead = lead + (double)prod_at(a, pc.start + u + SW - 1, pc.start + u + SW - 1 + jump);
      double acw = lead;
      if (u > 0) {
        lag = lag + (double)prod_at(a, pc.start + u - 1, pc.start + u - 1 + jump);
        acw = lead - lag;
      }
      const double num = acw + eps;
      // negative: divided by rms[u] first, then by rms[t]; positive: by rms[u + jump] first, then by rms[t]
      const double v = a.negative ? (num / rms[u]) / rms[t] : (num / rms[u + jump]) / rms[t];
      if (v > bv) { bv = v; bl = r; }
    }
    if (r == SW - 1 || t == rows - 1) {
      const int w = t / SW;
      if (w >= w_lo && w < w_hi) {
        const int64_t g = (int64_t)pc.first_out + (w - w_lo);
        a.loc[g * a.n_jumps + k] = (int16_t)bl;
        a.best[g * a.n_jumps + k] = bv;
      }
    }
  }
}

}  // namespace

extern "C" {

int dab_stretch_best_jumps(dab_ctx *ctx, const void *segment_f16, int32_t channels, int64_t n, int32_t negative,
                           const int32_t *jumps, int32_t n_jumps, int16_t *loc, double *best) {
  if (!ctx || !segment_f16 || !jumps || !loc || !best || (channels != 1 && channels != 2) || n_jumps <= 0) return DAB_E_ARG;
  if (n < 3 * SW - 1) { dab_set_err(ctx, "Invalid state in Pearson generator."); return DAB_E_ARG; }
  for (int k = 0; k < n_jumps; ++k)
    if (jumps[k] < 1 || jumps[k] >= SW) { dab_set_err(ctx, "dab_stretch_best_jumps: jump distances must be in [1, 512)"); return DAB_E_ARG; }
  DAB_CUDA(cudaSetDevice(ctx->device));
  // the reference's pieces (describealign.py:257-272): split while longer than 52 * 1.1 windows, advance by 49 windows
  std::vector<Piece> pieces;
  {
    const double limit = (S_CACHED + 2) * 1.1 * SW;
    int64_t start = 0;
    bool first = true;
    int32_t out = 0;
    for (;;) {
      const int64_t left = n - start;
      const bool last = !((double)left > limit);
      Piece pc;
      pc.start = start;
      pc.len = (int32_t)(last ? left : S_CUT + SW);
      pc.lo = first ? 0 : 1;
      const int32_t hi = last ? pc.len / SW : S_CACHED;
      pc.count = hi > pc.lo ? hi - pc.lo : 0;
      pc.first_out = out;
      out += pc.count;
      pieces.push_back(pc);
      if (last) break;
      start += S_CUT - SW;
      first = false;
    }
  }
  const int64_t n_windows = n / SW;
  const int32_t np = (int32_t)pieces.size();
  const int64_t rms_stride = (int64_t)((S_CACHED + 2) * 1.1 * SW) + 8;
  cudaStream_t st = nullptr;
  DAB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  void *d_x = nullptr, *d_pieces = nullptr, *d_jumps = nullptr, *d_rms = nullptr, *d_eps = nullptr, *d_loc = nullptr, *d_best = nullptr;
  const size_t xb = sizeof(__half) * (size_t)n * channels, ob = (size_t)n_windows * n_jumps;
  cudaError_t e = cudaMallocAsync(&d_x, xb, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_pieces, sizeof(Piece) * np, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_jumps, sizeof(int32_t) * n_jumps, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_rms, sizeof(double) * (size_t)np * rms_stride, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_eps, sizeof(double) * np, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_loc, sizeof(int16_t) * (ob + 1), st);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_best, sizeof(double) * (ob + 1), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, segment_f16, xb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_pieces, pieces.data(), sizeof(Piece) * np, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_jumps, jumps, sizeof(int32_t) * n_jumps, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    StretchArgs a;
    a.x = static_cast<const __half *>(d_x); a.n = n; a.ch = channels; a.negative = negative ? 1 : 0;
    a.pieces = static_cast<const Piece *>(d_pieces); a.n_pieces = np;
    a.jumps = static_cast<const int32_t *>(d_jumps); a.n_jumps = n_jumps;
    a.rms = static_cast<double *>(d_rms); a.eps = static_cast<double *>(d_eps); a.rms_stride = rms_stride;
    a.loc = static_cast<int16_t *>(d_loc); a.best = static_cast<double *>(d_best);
    stretch_energy_kernel<<<(unsigned)cdiv(np, 32), 32, 0, st>>>(a);
    const int64_t threads = (int64_t)np * n_jumps;
    stretch_jump_kernel<<<(unsigned)cdiv(threads, 64), 64, 0, st>>>(a);
    ctx->launches += 2;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && ob) e = cudaMemcpyAsync(loc, d_loc, sizeof(int16_t) * ob, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && ob) e = cudaMemcpyAsync(best, d_best, sizeof(double) * ob, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  void *bufs[] = {d_x, d_pieces, d_jumps, d_rms, d_eps, d_loc, d_best};
  for (void *b : bufs)
    if (b) cudaFreeAsync(b, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (e != cudaSuccess) {
    dab_set_err(ctx, std::string("dab_stretch_best_jumps: ") + cudaGetErrorString(e));
    return DAB_E_CUDA;
  }
  return DAB_OK;
}

}  // extern "C"
