// Device stage A of align() (reference describealign.py:596-700):
//   prep      41-tap Hann local-mean subtraction, 41-frame sliding L2 norm   (:599-608)
//   codes     7 quantised samples per feature -> packed digits + flags       (:622-628, 638-644)
//   select    not-quiet frames; video keeps every 4th of that list           (:629-630, 657-658)
//   tables    5 direct-address tables over the 7^7 codes, video frames stored under every
//             code reachable through their flagged digits                     (:615-633)
//   gate      per audio query: (>= 2 of tables 0-2) and (table 3 or 4)         (:649-660)
//   score     3 x 41-tap normalised correlation -> prob -> qual                (:662-673)
//   dp1       frontier DP as a prefix-max over video rank, with back pointers  (:654-656, 674-682)
//   trace     traceback through checkpointed back pointers                     (:690-697)
//
// Layout: every per-frame array is structure-of-arrays in HBM (feature-major) so that a warp
// touching consecutive frames reads consecutive addresses.  The hash "dict of sets" of the
// reference becomes a counting-sort CSR (start[code], items[]) per table: 7^7 = 823 543
// codes x 5 tables x 4 B = 16.5 MB of starts, L2 resident.  The gate is evaluated as an
// integer predicate on packed 3-bit digits, so a bucket entry costs five 4-byte loads.
#include "common.cuh"
#include "hann_tables.h"

namespace {

__constant__ double c_hflip[41];   // flipped 41-tap Hann (np.convolve flips; the window is not bit-symmetric)

// ------------------------------------------------------------------------------------------
// prep: ms = feat - conv_same(feat, hann41)   (f64, OpenBLAS ddot order incl. shorter edge dots)
// ------------------------------------------------------------------------------------------
struct PrepArgs {
  const float *f32[4];
  const double *f64;
  int64_t len[5];     // per-feature length (energy may be one longer)
  int64_t Lp;         // frames kept for ms / nrm storage (min length)
  double *ms;         // [5][Lms]  Lms = max len (stride)
  int64_t stride;
};

__device__ __forceinline__ double feat_at(const PrepArgs &a, int f, int64_t t) {
  return f < 4 ? (double)a.f32[f][t] : a.f64[t];
}

// A block walks MS_TILES consecutive tiles of 256 outputs of one feature; the 296 inputs a tile reads are converted to
// float64 once and staged in shared memory (every input is used by 41 outputs).  The next tile's inputs are fetched
// into registers before the current tile is computed, so the global-load latency (ncu: 6 of 13 stall cycles per
// issue were the long scoreboard with one tile per block) hides behind the 76 FP64 operations per output.
constexpr int MS_TILES = 4;

__global__ void __launch_bounds__(256, 3) meansub_kernel(PrepArgs a) {
  __shared__ double tile[2][256 + 40];
  const int f = blockIdx.y;
  const int64_t n = a.len[f];
  const int64_t first = (int64_t)blockIdx.x * (256 * MS_TILES);
  if (first >= n) return;
  const int tid = threadIdx.x;
  auto fetch = [&](int64_t t0, double &r0, double &r1) {
    const int64_t g0 = t0 - 20 + tid, g1 = g0 + 256;
    r0 = (g0 >= 0 && g0 < n) ? feat_at(a, f, g0) : 0.0;
    r1 = (tid < 40 && g1 >= 0 && g1 < n) ? feat_at(a, f, g1) : 0.0;
  };
  double r0, r1;
  fetch(first, r0, r1);
  tile[0][tid] = r0;
  if (tid < 40) tile[0][256 + tid] = r1;
  __syncthreads();
#pragma unroll 1
  for (int k = 0; k < MS_TILES; ++k) {
    const int64_t t0 = first + 256 * k;
    if (t0 >= n) break;
    const bool has_next = k + 1 < MS_TILES && t0 + 256 < n;
    if (has_next) fetch(t0 + 256, r0, r1);
    const double *cur = tile[k & 1];
    const int64_t base = t0 - 20;
    const int64_t t = t0 + tid;
    if (t < n) {
      int64_t lo = t - 20, hi = t + 21, klo = 0;
      if (lo < 0) { klo = -lo; lo = 0; }
      if (hi > n) hi = n;
      const int m = (int)(hi - lo);
      const double *src = cur + (lo - base);
      auto X = [&](int i) { return src[i]; };
      auto Y = [&](int i) { return c_hflip[klo + i]; };
      // interior outputs (m == 41 implies klo == 0): the window taps are compile-time constant-bank operands
      auto Y0 = [&](int i) { return c_hflip[i]; };
      const double mean = (m == 41) ? ddot41_skx(X, Y0) : ddot_skx(X, Y, m);
      a.ms[(int64_t)f * a.stride + t] = cur[t - base] - mean;
    }
    if (has_next) {
      tile[(k + 1) & 1][tid] = r0;
      if (tid < 40) tile[(k + 1) & 1][256 + tid] = r1;
    }
    __syncthreads();
  }
}

// nrm[t] = max(1e-3, sqrt(sum_{k<41} ms[t+k]^2)) and the digit codes of frame t.
struct CodeArgs {
  const double *ms;     // [5][stride]
  int64_t stride;
  int64_t len[5];
  double *nrm;          // [3][nstride] (features 0..2 only are stored)
  int64_t nstride;
  uint32_t *pack;       // [5][nstride]
  int32_t *code;        // [5][nstride]
  int is_video;
};

__global__ void norm_codes_kernel(CodeArgs a) {
  const int f = blockIdx.y;
  const int64_t ncodes = a.len[f] - 40;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncodes) return;
  const double *m = a.ms + (int64_t)f * a.stride + t;
  auto X = [&](int i) { double v = m[i]; return v * v; };
  auto Y = [&](int) { return 1.0; };
  double nr = sqrt(ddot41_skx(X, Y));
  if (nr < 0.001) nr = 0.001;
  if (f < 3) a.nrm[(int64_t)f * a.nstride + t] = nr;
  uint32_t pk = 0;
  int32_t code = 0, p7 = 1;
  // The digits are floor(8 * (m / nr) + offset) and the flags compare its fraction with .6: integers.  The
  // quotient is first taken as m * (1 / nr) - within 2 ulp of the true quotient - and only a value that
  // lands within 1e-9 of a boundary (an integer, an integer + .6, the clip limits) is divided exactly, so
  // the digits are those of the exact arithmetic with one division per frame instead of seven.
  const double inv = 1.0 / nr;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const double mk = m[2 + 6 * k];
    double d = 8.0 * (mk * inv) + (a.is_video ? 3.3 : 3.5);
    {
      const double fr = d - floor(d);
      const bool risky = fr < 1e-9 || fr > 1.0 - 1e-9 || (a.is_video && fabs(fr - 0.6) < 1e-9) || !(fabs(d) < 1e6);
      if (risky) {
        d = mk / nr;
        d = 8.0 * d;
        d = d + (a.is_video ? 3.3 : 3.5);
      }
    }
    int dig;
    if (a.is_video) {
      d = fmin(fmax(d, 0.0), 6.0);
      double fl = floor(d);
      if (d - fl > 0.6) pk |= 8u << (4 * k);
      dig = (int)fl;
    } else {
      double fl = floor(d);
      fl = fmin(fmax(fl, 0.0), 6.0);
      dig = (int)fl;
    }
    pk |= (uint32_t)dig << (4 * k);
    code += dig * p7;
    p7 *= 7;
  }
  a.pack[(int64_t)f * a.nstride + t] = pk;
  a.code[(int64_t)f * a.nstride + t] = code;
}

// ------------------------------------------------------------------------------------------
// scans / compaction
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *smem /* >= 32 ints */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += y;
    }
    smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  int res = inc - v + smem[warp];
  *total = smem[32];
  __syncthreads();
  return res;
}

// n_dev (optional): the number of valid elements lives in device memory (a count produced earlier in
// the same stream); launches are sized for the capacity n and elements past *n_dev count as zero.
__device__ __forceinline__ int64_t scan_len(int64_t n, const int32_t *n_dev) {
  if (!n_dev) return n;
  const int64_t m = *n_dev;
  return m < n ? (m < 0 ? 0 : m) : n;
}

// Single-pass exclusive scan (decoupled look-back): one launch per scan.  Tiles take their index from an
// atomic ticket, so a tile's predecessors have always started; a tile publishes its aggregate, then walks
// back over its predecessors' published (aggregate | inclusive prefix) words until it meets an inclusive
// prefix.  Status words carry the epoch of the scan call, so the array needs no clearing between calls;
// the last tile to finish resets the ticket counter.
//   status[t] = epoch << 34 | kind << 32 | value      kind 1 = aggregate, 2 = inclusive prefix
struct ScanArgs {
  const int32_t *in;
  int32_t *out;
  int64_t n;
  const int32_t *n_dev;
  unsigned long long *status;      // [tiles]
  unsigned int *ticket;            // [0] next tile, [1] tiles finished
  unsigned int epoch;              // 1 .. 2^30 - 1, different from the previous call's on this array
  int32_t *total_out;
};

__global__ void __launch_bounds__(SCAN_THREADS) scan_onepass_kernel(ScanArgs a) {
  __shared__ int sm[33];
  __shared__ unsigned int s_tile;
  __shared__ int s_prefix;
  const int64_t n = scan_len(a.n, a.n_dev);
  if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  if ((int64_t)tile > n / SCAN_TILE) {
    // past the tile that holds position n (a device-side count far below the capacity): nobody looks back here
    if (threadIdx.x == 0) {
      const unsigned int done = atomicAdd(a.ticket + 1, 1u);
      if (done == gridDim.x - 1) { a.ticket[0] = 0u; a.ticket[1] = 0u; __threadfence(); }
    }
    return;
  }
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? a.in[base + k] : 0;
    s += v[k];
  }
  int total;
  const int ex = block_exclusive_scan(s, &total, sm);
  if (threadIdx.x < 32) {
    // warp 0 publishes the tile's aggregate, then looks back 32 predecessors at a time
    const int lane = threadIdx.x;
    const unsigned long long tag = (unsigned long long)a.epoch << 34;
    volatile unsigned long long *st = a.status;
    int prefix = 0;
    if (tile > 0) {
      if (lane == 0) { st[tile] = tag | (1ull << 32) | (unsigned int)total; __threadfence(); }
      int end = (int)tile;                      // tiles [.., end) are still to be accounted for
      while (end > 0) {
        const int t = end - 1 - lane;           // lane 0 = nearest predecessor
        unsigned long long w = 0ull;
        if (t >= 0) { do { w = st[t]; } while ((w >> 34) != a.epoch); }
        const bool is_prefix = t >= 0 && ((w >> 32) & 3ull) == 2ull;
        const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
        const int stop = pm ? __ffs(pm) - 1 : 32;             // nearest lane holding an inclusive prefix
        int v = (t >= 0 && lane <= stop) ? (int)(unsigned int)w : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        prefix += v;
        if (pm) break;
        end -= 32;
      }
    }
    if (lane == 0) {
      st[tile] = tag | (2ull << 32) | (unsigned int)(prefix + total);
      __threadfence();
      s_prefix = prefix;
    }
  }
  __syncthreads();
  int run = ex + s_prefix;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) a.out[base + k] = run;
    run += v[k];
  }
  // the tile that holds position n (or the last tile) writes the total
  const int64_t last_tile = n / SCAN_TILE;
  if ((int64_t)tile == last_tile && threadIdx.x == 0) {
    // total of all elements = prefix of this tile + its own total (elements past n are zero)
    a.out[n] = s_prefix + total;
    if (a.total_out) *a.total_out = s_prefix + total;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(a.ticket + 1, 1u);
    if (done == gridDim.x - 1) { a.ticket[0] = 0u; a.ticket[1] = 0u; __threadfence(); }
  }
}

__global__ void notquiet_flag_kernel(const float *energy, int64_t n, int32_t *flag) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flag[t] = energy[t] > 0.5f ? 1 : 0;
}

// list[rank / every] = t for flagged t whose rank % every == 0
__global__ void compact_kernel(const int32_t *flag, const int32_t *off, int64_t n, int every, int32_t *list) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && flag[t]) {
    int r = off[t];
    if (r % every == 0) list[r / every] = (int32_t)t;
  }
}

// ------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t pack_to_code(uint32_t pk) {
  int32_t c = 0, p7 = 1;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    c += (int32_t)((pk >> (4 * k)) & 7u) * p7;
    p7 *= 7;
  }
  return c;
}

// One thread per (selected video frame, table).  A frame is stored under every code reachable through its
// flagged digits (2^flags codes, :630-633).  Three steps without a second round of atomics:
//   table_expand_kernel   how many codes each (frame, table) expands to -> scan -> where its entries' bucket
//                         positions are kept
//   table_kernel<false>   atomicAdd on the code's counter; the value it returns IS the entry's position in
//                         its bucket and is kept
//   table_kernel<true>    after the scan of the counters: items[start[code] + position] = frame, plain stores
__global__ void table_expand_kernel(const uint32_t *pack, int64_t nstride, const int32_t *sel, const int32_t *dc, int64_t vsel_cap,
                                    int32_t *ecount) {
  const int f = blockIdx.y;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= vsel_cap) return;
  int32_t c = 0;
  if (s < dc[DC_N_VSEL]) c = 1 << __popc(pack[(int64_t)f * nstride + sel[s]] & 0x8888888u);
  ecount[(int64_t)f * vsel_cap + s] = c;
}

template <bool FILL>
__global__ void table_kernel(const uint32_t *pack, int64_t nstride, const int32_t *sel, const int32_t *dc, int64_t vsel_cap,
                             const int32_t *ebase, int32_t *tpos, int32_t *count, const int32_t *start, int32_t *items,
                             int64_t items_cap) {
  const int f = blockIdx.y;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= dc[DC_N_VSEL]) return;
  if (dc[DC_OVERFLOW] & DAB_OVF_ENTRIES) return;
  const uint32_t pk = pack[(int64_t)f * nstride + sel[s]];
  const uint32_t fl = pk & 0x8888888u;             // flag bits, one per nibble
  const int32_t base = pack_to_code(pk);
  const int32_t P7[7] = {1, 7, 49, 343, 2401, 16807, 117649};
  int64_t e = ebase[(int64_t)f * vsel_cap + s];
  // four codes per step: the four atomics (count pass) / the four position and start loads (fill pass) are
  // issued before the first result is used, so their latencies overlap instead of adding up
  uint32_t sub = fl;
  bool more = true;
  while (more) {
    int64_t slot[4];
    bool live[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      live[u] = more;
      int32_t c = base;
#pragma unroll
      for (int k = 0; k < 7; ++k)
        if (sub & (8u << (4 * k))) c += P7[k];
      slot[u] = (int64_t)f * DAB_NCODE + c;
      if (more) { if (sub == 0) more = false; else sub = (sub - 1) & fl; }
    }
    int32_t v0[4], v1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v0[u] = v1[u] = 0;
      if (live[u] && e + u < items_cap) {
        if (!FILL) v0[u] = atomicAdd(&count[slot[u]], 1);
        else { v0[u] = tpos[e + u]; v1[u] = start[slot[u]]; }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (live[u] && e + u < items_cap) {
        if (!FILL) tpos[e + u] = v0[u];
        else {
          const int64_t at = (int64_t)v1[u] + v0[u];
          if (at < items_cap) items[at] = (int32_t)s;
        }
      }
    }
    e += 4;
  }
}

// ------------------------------------------------------------------------------------------
// gate
// ------------------------------------------------------------------------------------------
// The gate compares the 7 quantised digits of an audio frame with those of a video frame, per feature
// (SURVEY.md A.4): the frames match in a feature when every audio digit equals the video digit, or the
// video digit + 1 where that digit is flagged (fraction > .6, stored under both codes, :622-633).
// Digits live one per nibble (the format norm_codes_kernel writes): bits 0-2 the digit (0..6), bit 3 the
// flag (video) / a guard bit set by the gate (audio).
//   d = (audio | 0x8888888) - (video digits)         per nibble 8 + a - v, no borrow between nibbles
//   d ^ 0x8888888                                    a - v for a >= v (0..6), >= 10 for a < v
//   ... & ~flags                                     0 exactly when a - v is 0, or 1 on a flagged digit
// i.e. five integer instructions per feature instead of a per-digit loop.
__device__ __forceinline__ bool digits_match(uint32_t a_guarded, uint32_t v) {
  const uint32_t d = (a_guarded - (v & 0x7777777u)) ^ 0x8888888u;
  return (d & ~((v >> 3) & 0x1111111u)) == 0u;
}

// The five digit words (one digit per nibble, see digits_match) of every hashed video frame side by side
// (32 bytes per frame, indexed by the
// frame's rank in the hashed list): a bucket entry then costs the gate ONE 32-byte sector instead of a
// rank -> frame look-up plus five gathers from five arrays.
__global__ void video_records_kernel(const uint32_t *pack, int64_t nstride, const int32_t *sel, const int32_t *dc, uint4 *rec) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= dc[DC_N_VSEL]) return;
  const int32_t v = sel[s];
  uint4 r0, r1;
  r0.x = pack[v]; r0.y = pack[nstride + v]; r0.z = pack[2 * nstride + v]; r0.w = pack[3 * nstride + v];
  r1.x = pack[4 * nstride + v]; r1.y = (uint32_t)v; r1.z = 0u; r1.w = 0u;
  rec[2 * s] = r0;
  rec[2 * s + 1] = r1;
}

constexpr int GATE_STASH = 4;   // rows with at most this many candidates (nearly all) are not enumerated a second time

struct GateArgs {
  const int32_t *a_code;    // [5][a_nstride]
  const uint32_t *a_pack;
  int64_t a_nstride;
  const int32_t *a_list;    // not-quiet audio frames
  const int32_t *dc;        // device counts: query range [DC_Q_LO, DC_Q_HI) of the list
  int64_t q_cap;            // launch size (rows)
  const uint4 *v_rec;       // per hashed video frame (rank): its five digit packs, 32 bytes = one sector
  const int32_t *start;     // [5*NCODE + 1]
  const int32_t *items;
  int32_t *row_count;       // count pass output
  const int32_t *row_off;   // fill pass input
  int32_t *cand_tmp, *cand_s, *cand_i;
  int64_t cand_cap;
  int32_t *stash;           // [q_cap][GATE_STASH]: the first candidates of every row, kept by the count pass
  int32_t *big_list;        // rows with more than GATE_STASH candidates (fill pass work list)
  int32_t *big_count;
};

// Fill pass, rows of up to GATE_STASH candidates (nearly all): thread per query; the candidates the count pass
// stashed are ordered by video rank and written out.  Larger rows go to a work list for gate_fill_big_kernel.
__global__ void gate_fill_small_kernel(GateArgs g) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.q_cap || q >= g.dc[DC_N_Q]) return;
  if (g.dc[DC_OVERFLOW] & DAB_OVF_CAND) return;
  const int64_t off0 = g.row_off[q];
  const int n_row = g.row_off[q + 1] - (int)off0;
  if (n_row == 0) return;
  if (n_row > GATE_STASH) {
    g.big_list[atomicAdd(g.big_count, 1)] = (int32_t)q;
    return;
  }
  if (off0 + n_row > g.cand_cap) return;
  const int32_t i = g.a_list[g.dc[DC_Q_LO] + q];
  const int4 st = *reinterpret_cast<const int4 *>(g.stash + q * GATE_STASH);
  static_assert(GATE_STASH == 4, "the stash is read as one int4");
  int32_t x[4] = {st.x, n_row > 1 ? st.y : 0x7fffffff, n_row > 2 ? st.z : 0x7fffffff, n_row > 3 ? st.w : 0x7fffffff};
  // sorting network of four
  auto cswap = [&](int a, int b) { if (x[b] < x[a]) { const int32_t t = x[a]; x[a] = x[b]; x[b] = t; } };
  cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3); cswap(1, 2);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n_row) { g.cand_s[off0 + k] = x[k]; g.cand_i[off0 + k] = i; }
}

// Fill pass, rows with more than GATE_STASH candidates (false-match clusters; a few hundred per pair): a warp per
// row of the work list enumerates the row's two buckets again and orders what it finds by video rank.
__global__ void gate_fill_big_kernel(GateArgs g) {
  const int lane = threadIdx.x & 31;
  const int n_big = *g.big_count;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_big; w += n_warps) {
    const int64_t q = g.big_list[w];
    const int32_t i = g.a_list[g.dc[DC_Q_LO] + q];
    // lanes 0-4 fetch the query's five digit words and bucket ranges, one feature each; then broadcast
    uint32_t my_ap = 0u;
    int32_t my_st = 0, my_en = 0;
    if (lane < 5) {
      my_ap = g.a_pack[(int64_t)lane * g.a_nstride + i] | 0x8888888u;
      const int64_t slot = (int64_t)lane * DAB_NCODE + g.a_code[(int64_t)lane * g.a_nstride + i];
      my_st = g.start[slot];
      my_en = g.start[slot + 1];
    }
    uint32_t ap[5];
    int32_t st[5], nb[5];
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      ap[f] = __shfl_sync(0xffffffffu, my_ap, f);
      st[f] = __shfl_sync(0xffffffffu, my_st, f);
      nb[f] = __shfl_sync(0xffffffffu, my_en, f) - st[f];
    }
    // every candidate lies in at least one bucket of any pair out of {0,1,2}, and in bucket 3 or 4
    int fa = 0, fb = 1, best = nb[0] + nb[1];
    if (nb[0] + nb[2] < best) { best = nb[0] + nb[2]; fa = 0; fb = 2; }
    if (nb[1] + nb[2] < best) { best = nb[1] + nb[2]; fa = 1; fb = 2; }
    if (nb[3] + nb[4] < best) { best = nb[3] + nb[4]; fa = 3; fb = 4; }
    int found = 0;
    const int64_t off = g.row_off[q];
    // the two buckets are walked as one list of `best` entries, 32 per step
    const int sa = st[fa], ca = nb[fa], sb = st[fb];
    for (int base = 0; base < best; base += 32) {
      const int e = base + lane;
      bool ok = false;
      int32_t s = 0;
      if (e < best) {
        const bool second = e >= ca;
        s = g.items[second ? sb + (e - ca) : sa + e];
        const uint4 r0 = __ldg(g.v_rec + 2 * (int64_t)s), r1 = __ldg(g.v_rec + 2 * (int64_t)s + 1);
        const uint32_t vp[5] = {r0.x, r0.y, r0.z, r0.w, r1.x};
        bool m[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) m[f] = digits_match(ap[f], vp[f]);
        ok = ((int)m[0] + (int)m[1] + (int)m[2] >= 2) && (m[3] || m[4]);
        if (second && m[fa]) ok = false;   // already enumerated from the first bucket
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int64_t pos = off + found + __popc(bal & ((1u << lane) - 1u));
        if (pos < g.cand_cap) g.cand_tmp[pos] = s;
      }
      found += __popc(bal);
    }
    if (off + found > g.cand_cap) continue;
    __syncwarp();
    // order the row's candidates by video rank: position = number of smaller entries
    if (found <= 32) {
      // one candidate per lane, ranks by 32 register shuffles instead of found^2 memory reads
      const int32_t x = lane < found ? g.cand_tmp[off + lane] : 0x7fffffff;
      int r = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) r += (__shfl_sync(0xffffffffu, x, j) < x);
      if (lane < found) { g.cand_s[off + r] = x; g.cand_i[off + r] = i; }
      continue;
    }
    for (int k = lane; k < found; k += 32) {
      const int32_t x = g.cand_tmp[off + k];
      int r = 0;
      for (int j = 0; j < found; ++j) r += (g.cand_tmp[off + j] < x);
      g.cand_s[off + r] = x;
      g.cand_i[off + r] = i;
    }
  }
}

// ------------------------------------------------------------------------------------------
// gate, count pass, balanced over bucket ENTRIES instead of queries.  A query visits the entries of its two
// cheapest buckets; their number ranges from none to several hundred, so a warp per query leaves most lanes
// idle (ncu: 380 warp instructions per query for 33 entries on average).  Three steps:
//   gate_plan_kernel      thread per query: digit words, the two buckets, their joint length `best`
//   scan                  P = exclusive scan of best; its total is the number of entries to test
//   gate_entries_kernel   tiles of 2048 consecutive entries of the concatenated bucket lists; the tile finds
//                         the queries it spans by a warp-wide 32-ary search in P, marks where each one starts,
//                         spreads the owners by a max-scan, and every thread tests 8 consecutive entries.
//                         Matches are appended to the query's stash through an atomic row counter.
// The fill pass takes rows of up to GATE_STASH candidates from the stash (gate_fill_small_kernel; in any
// order - it sorts them by video rank), larger rows are enumerated again by a warp.
// ------------------------------------------------------------------------------------------
struct __align__(16) GateQ {
  uint32_t ap[5];     // guarded digit words of the query
  int32_t sa, sb;     // bucket starts in items[]
  uint32_t ca_fa;     // entries in the first bucket | first bucket's feature << 28
};

__global__ void gate_plan_kernel(GateArgs g, GateQ *rec, int32_t *best_out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.q_cap) return;
  g.row_count[q] = 0;
  if (q >= g.dc[DC_N_Q]) { best_out[q] = 0; return; }
  const int32_t i = g.a_list[g.dc[DC_Q_LO] + q];
  uint32_t ap[5];
  int32_t st[5], nb[5];
#pragma unroll
  for (int f = 0; f < 5; ++f) {
    ap[f] = g.a_pack[(int64_t)f * g.a_nstride + i] | 0x8888888u;
    const int64_t slot = (int64_t)f * DAB_NCODE + g.a_code[(int64_t)f * g.a_nstride + i];
    st[f] = g.start[slot];
    nb[f] = g.start[slot + 1] - st[f];
  }
  // every candidate lies in at least one bucket of any pair out of {0,1,2}, and in bucket 3 or 4
  int fa = 0, fb = 1, best = nb[0] + nb[1];
  if (nb[0] + nb[2] < best) { best = nb[0] + nb[2]; fa = 0; fb = 2; }
  if (nb[1] + nb[2] < best) { best = nb[1] + nb[2]; fa = 1; fb = 2; }
  if (nb[3] + nb[4] < best) { best = nb[3] + nb[4]; fa = 3; fb = 4; }
  GateQ r;
#pragma unroll
  for (int f = 0; f < 5; ++f) r.ap[f] = ap[f];
  int32_t sa = 0, sb = 0, ca = 0;
#pragma unroll
  for (int f = 0; f < 5; ++f) {
    if (f == fa) { sa = st[f]; ca = nb[f]; }
    if (f == fb) sb = st[f];
  }
  r.sa = sa; r.sb = sb; r.ca_fa = (uint32_t)ca | ((uint32_t)fa << 28);
  rec[q] = r;
  best_out[q] = best;
}

constexpr int GE_THREADS = 256, GE_PER = 8, GE_TILE = GE_THREADS * GE_PER;

// last index q in [0, n] with P[q] <= x (P non-decreasing, P[0] = 0 <= x), found by one warp: 32 probes per round
__device__ __forceinline__ int warp_last_le(const int32_t *P, int n, int x, int lane) {
  int lo = 0, hi = n + 1;                 // P[lo] <= x; hi == n + 1 or P[hi] > x
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) / 32;
    const int idx = lo + (lane + 1) * step;
    const bool le = idx < hi && P[idx] <= x;
    const int cnt = __popc(__ballot_sync(0xffffffffu, le));     // monotone: the lanes with le form a prefix
    const int nlo = lo + cnt * step;
    const int nhi = lo + (cnt + 1) * step;
    lo = nlo;
    if (nhi < hi) hi = nhi;
  }
  return lo;
}

__global__ void __launch_bounds__(GE_THREADS) gate_entries_kernel(GateArgs g, const GateQ *rec, const int32_t *P) {
  __shared__ __align__(16) int owner[GE_TILE];
  __shared__ int s_q[2];
  __shared__ int s_wmax[GE_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = g.dc[DC_N_Q];
  const int total = g.dc[DC_ENUM_LO];
  if (g.dc[DC_OVERFLOW] & DAB_OVF_ENTRIES) return;
  for (int64_t E0l = (int64_t)blockIdx.x * GE_TILE; E0l < total; E0l += (int64_t)gridDim.x * GE_TILE) {
    const int E0 = (int)E0l;
    const int E1 = E0 + GE_TILE < total ? E0 + GE_TILE : total;
    // the queries this tile spans: q0 owns entry E0, q1 owns entry E1 - 1 (both non-empty rows)
    if (warp < 2) {
      const int r = warp_last_le(P, nq, warp == 0 ? E0 : E1 - 1, lane);
      if (lane == 0) s_q[warp] = r;
    }
#pragma unroll
    for (int j = 0; j < GE_PER; ++j) owner[tid + j * GE_THREADS] = 0;
    __syncthreads();
    const int q0 = s_q[0], q1 = s_q[1];
    // every later query that starts inside the tile marks its first entry (empty rows share a position with the
    // non-empty row that follows them: the largest index wins)
    for (int q = q0 + 1 + tid; q <= q1; q += GE_THREADS) {
      const int pos = P[q] - E0;
      if (pos > 0 && pos < GE_TILE) atomicMax(&owner[pos], q - q0);
    }
    __syncthreads();
    // inclusive max-scan: slot e belongs to the last marked query at or before it
    int m[GE_PER];
    {
      const int4 a = *reinterpret_cast<const int4 *>(&owner[tid * GE_PER]);
      const int4 b = *reinterpret_cast<const int4 *>(&owner[tid * GE_PER + 4]);
      m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
    }
#pragma unroll
    for (int j = 1; j < GE_PER; ++j) m[j] = max(m[j], m[j - 1]);
    int inc = m[GE_PER - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc = max(inc, y);
    }
    if (lane == 31) s_wmax[warp] = inc;
    int before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) before = 0;
    __syncthreads();
    for (int w = 0; w < warp; ++w) before = max(before, s_wmax[w]);
    // owners back to shared memory, then entries interleaved over the threads: consecutive lanes read consecutive
    // bucket items (coalesced) and mostly share a query, whose record is then one broadcast load
    __syncthreads();
#pragma unroll
    for (int j = 0; j < GE_PER; ++j) owner[tid * GE_PER + j] = max(before, m[j]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < GE_PER; ++j) {
      const int e = E0 + j * GE_THREADS + tid;
      if (e < E1) {
        const int q = q0 + owner[j * GE_THREADS + tid];
        const uint4 *rp = reinterpret_cast<const uint4 *>(rec + q);
        const uint4 x0 = __ldg(rp), x1 = __ldg(rp + 1);
        const int el = e - P[q];
        const int ca = (int)(x1.w & 0x0fffffffu), fa = (int)(x1.w >> 28);
        const bool second = el >= ca;
        const int32_t s = g.items[second ? (int32_t)x1.z + (el - ca) : (int32_t)x1.y + el];
        // features 0-3 are the first 16 bytes of the frame's record; feature 4 is only fetched when it decides
        const uint4 v0 = __ldg(g.v_rec + 2 * (int64_t)s);
        const bool m0 = digits_match(x0.x, v0.x), m1 = digits_match(x0.y, v0.y), m2 = digits_match(x0.z, v0.z),
                   m3 = digits_match(x0.w, v0.w);
        bool ok = (int)m0 + (int)m1 + (int)m2 >= 2;
        if (ok && second) {
          // already enumerated from the first bucket
          const bool mfa = fa == 0 ? m0 : (fa == 1 ? m1 : m3);
          if (mfa) ok = false;
        }
        if (ok && !m3) ok = digits_match(x1.x, __ldg(reinterpret_cast<const uint32_t *>(g.v_rec + 2 * (int64_t)s + 1)));
        if (ok) {
          const int idx = atomicAdd(&g.row_count[q], 1);
          if (idx < GATE_STASH) g.stash[(int64_t)q * GATE_STASH + idx] = s;
        }
      }
    }
    __syncthreads();     // owner[] and s_q[] are rewritten by the next tile
  }
}

// ------------------------------------------------------------------------------------------
// scoring: thread per candidate
// ------------------------------------------------------------------------------------------
struct ScoreArgs {
  const int32_t *cand_s, *cand_i;
  const int32_t *dc;             // n_cand = dc[DC_N_CAND]
  const int32_t *v_sel;
  const double *a_ms, *v_ms;     // [5][stride]
  int64_t a_stride, v_stride;
  const double *a_nrm, *v_nrm;   // [3][nstride]
  int64_t a_nstride, v_nstride;
  double *qual;                  // < 0: rejected
  int32_t *keep;
};

__global__ void score_kernel(ScoreArgs s) {
  // the launch is a fixed number of blocks (the candidate count is only known to the device): grid stride
  const int64_t n_cand = s.dc[DC_N_CAND];
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cand; c += (int64_t)gridDim.x * blockDim.x) {
  const int32_t i = s.cand_i[c];
  const int32_t v = s.v_sel[s.cand_s[c]];
  double prob = 1.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double *am = s.a_ms + (int64_t)j * s.a_stride + i;
    const double *vm = s.v_ms + (int64_t)j * s.v_stride + v;
    auto X = [&](int k) { return am[k]; };
    auto Y = [&](int k) { return vm[k]; };
    double corr = ddot41_skx(X, Y);
    corr = corr / (s.a_nrm[(int64_t)j * s.a_nstride + i] * s.v_nrm[(int64_t)j * s.v_nstride + v]);
    prob = prob * fmax(1e-8, 1.0 - corr);
  }
  prob = pow(prob, 2.9);
  double qual = -1.0;
  if (!(prob > 1e-8)) qual = fmin(50.0, pow(prob / 1e-12, -1.0 / 3));
  s.qual[c] = qual;
  s.keep[c] = qual >= 0.0 ? 1 : 0;
  }
}

__global__ void gather_points_kernel(const int32_t *keep, const int32_t *off, const int32_t *dc,
                                     const int32_t *cand_i, const int32_t *cand_s, const double *qual,
                                     int32_t *pt_i, int32_t *pt_s, double *pt_q) {
  const int64_t n_cand = dc[DC_N_CAND];
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cand; c += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[c]) continue;
    const int o = off[c];
    pt_i[o] = cand_i[c];
    pt_s[o] = cand_s[c];
    pt_q[o] = qual[c];
  }
}

// ------------------------------------------------------------------------------------------
// frontier DP #1: cum(p) = q(p) + max{cum(p') : p' earlier, rank' <= rank},
// ties to the smallest rank; the seed has cum 0.  One warp walks the points in (i, v) order;
// the prefix maximum lives in a 32-ary max tree over the video ranks (HBM/L2 resident).
// ------------------------------------------------------------------------------------------
struct __align__(16) Node1 {
  double cum;     // 0 = empty / seed
  int32_t id;     // point id, -1 = seed
  int32_t rank;
};

constexpr int DP_LEVELS = 4;
constexpr int DP_CHECK = 256;   // checkpoint spacing for the parallel traceback

struct Dp1Args {
  const int32_t *pt_s;
  const double *pt_q;
  const int32_t *n_points;       // device count
  Node1 *level[DP_LEVELS];
  int4 *meta;                    // per point (back pointer, chain length, checkpoint id, -)
  int32_t *result;               // [0] end id, [1] path length
};

__device__ __forceinline__ bool key_better(double ca, int ra, double cb, int rb) {
  return ca > cb || (ca == cb && ra < rb);
}

__device__ __forceinline__ Node1 load_node_cg(const Node1 *p) {
  // L2 read (bypasses L1): the node may have been rewritten by another lane's fire-and-forget store
  const int4 v = __ldcg(reinterpret_cast<const int4 *>(p));
  Node1 n;
  n.cum = __hiloint2double(v.y, v.x);
  n.id = v.z;
  n.rank = v.w;
  return n;
}

// The walk keeps the frontier's top entry (largest cum; ties -> smallest rank) in registers.
// A point whose rank is >= the top's rank has the top as its predecessor and becomes the new
// top (quals are > 0): no tree loads, one f64 add on the dependent chain; the leaf, the back
// record and the old top's level-k ancestors (written only when the top leaves that ancestor's
// block - the ancestors of the current top are implied by the registers; every other node is
// exact) are stored lane-parallel for a whole run of such points at a time.  Only points left
// of the top (false matches, backward jumps) run the full prefix-max query: one 32-node row per
// level (512 B, coalesced), lanes left of the path feed the query and the lane on the path keeps
// the old node for the conditional update.
__global__ void __launch_bounds__(32, 1) dp1_kernel(Dp1Args a) {
  __shared__ int s_r[2][32];
  __shared__ double s_q[2][32];
  const int lane = threadIdx.x;
  const int n = *a.n_points;
  double top_cum = 0.0;
  int top_id = -1, top_rank = -1, top_len = 0, top_cp = -1;
  const int role = lane < DP_LEVELS ? lane : 0;
  int4 *const my_level = reinterpret_cast<int4 *>(a.level[role]);
  const int my_shift = 5 * role;
  int rr = lane < n ? a.pt_s[lane] : 0;
  double qq = lane < n ? a.pt_q[lane] : 0.0;
  for (int base = 0; base < n; base += 32) {
    const int buf = (base >> 5) & 1;
    s_r[buf][lane] = rr;
    s_q[buf][lane] = qq;
    __syncwarp();
    if (base + 32 + lane < n) { rr = a.pt_s[base + 32 + lane]; qq = a.pt_q[base + 32 + lane]; }
    const int cnt = n - base < 32 ? n - base : 32;
    // ---- runs of points that each extend the top (rank >= the rank before it) are handled 32 at
    //      a time: the cums are one sequential chain of f64 adds (exactly the reference's order),
    //      everything else - back records, leaves, ancestor flushes - is lane-parallel ----------
    const int my_r = s_r[buf][lane];
    const int prev_r = lane > 0 ? s_r[buf][lane - 1] : 0;
    const int next_r = lane + 1 < cnt ? s_r[buf][lane + 1] : -1;
    const unsigned inc = __ballot_sync(0xffffffffu, lane < cnt && (lane == 0 || my_r >= prev_r));
    int t = 0;
    while (t < cnt) {
      const int p = base + t;
      const int r = s_r[buf][t];
      const double q = s_q[buf][t];
      if (r >= top_rank) {
        int len = 1;
        if (t + 1 < 32) {
          const unsigned stop = (~inc) >> (t + 1);
          len += stop ? __ffs(stop) - 1 : 31 - t;
        }
        if (len > cnt - t) len = cnt - t;
        const bool in_run = lane >= t && lane < t + len;
        // The cums of the run are top_cum + q, + q, ... added one after the other in float64 (the
        // reference's order).  While the sums stay inside the binade of top_cum every such add is exact
        // arithmetic on multiples of that binade's unit once q is rounded to it (fl(x + q) = x + rn_u(q),
        // ties aside), so a warp prefix sum gives all of them in five steps; each lane then checks its
        // value against the one rounded add the reference performs, and only a run that fails the check
        // (binade crossing, rounding tie) is added up serially.
        double my_cum = 0.0;
        {
          const double my_q = in_run ? s_q[buf][lane] : 0.0;
          const unsigned long long tb = (unsigned long long)__double_as_longlong(top_cum);
          const double big = __longlong_as_double((long long)((tb & 0x7ff0000000000000ull) | 0x0008000000000000ull));  // 1.5 * 2^e
          double ps = in_run ? __dadd_rn(__dadd_rn(my_q, big), -big) : 0.0;       // q rounded to the unit of top_cum's binade
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, ps, o);
            if (lane >= t + o) ps = ps + y;
          }
          const double spec = top_cum + ps;
          double prev = __shfl_up_sync(0xffffffffu, spec, 1);
          if (lane == t) prev = top_cum;
          const bool ok = !in_run || (top_cum > 0.0 && __double_as_longlong(prev + my_q) == __double_as_longlong(spec));
          if (__all_sync(0xffffffffu, ok)) {
            my_cum = spec;
          } else {
            double c = top_cum;
            for (int u = 0; u < len; ++u) {
              c = c + s_q[buf][t + u];
              if (lane == t + u) my_cum = c;
            }
          }
        }
        const bool first = lane == t;
        double old_cum = __shfl_up_sync(0xffffffffu, my_cum, 1);
        int old_id = base + lane - 1, old_rank = prev_r;
        if (first) { old_cum = top_cum; old_id = top_id; old_rank = top_rank; }
        const int ln = top_len + 1 + (lane - t);
        const unsigned fm = __ballot_sync(0xffffffffu, in_run && (ln % DP_CHECK == 0 || (first && top_id < 0)));
        const unsigned fml = fm & ((2u << lane) - 1u);
        const int cpv = fml ? base + 31 - __clz(fml) : top_cp;
        if (in_run) {
          int4 m; m.x = old_id; m.y = ln; m.z = cpv; m.w = 0;
          a.meta[base + lane] = m;
          if (lane == t + len - 1 || next_r != my_r) {       // the last point on a rank owns its leaf
            Node1 me; me.cum = my_cum; me.id = base + lane; me.rank = my_r;
            a.level[0][my_r] = me;
          }
          if (old_id >= 0) {                                  // the old top leaves ancestor blocks
#pragma unroll
            for (int k = 1; k < DP_LEVELS; ++k) {
              if ((my_r >> (5 * k)) != (old_rank >> (5 * k))) {
                Node1 me; me.cum = old_cum; me.id = old_id; me.rank = old_rank;
                a.level[k][old_rank >> (5 * k)] = me;
              }
            }
          }
        }
        const int last = t + len - 1;
        top_cum = __shfl_sync(0xffffffffu, my_cum, last);
        top_rank = __shfl_sync(0xffffffffu, my_r, last);
        top_len = __shfl_sync(0xffffffffu, ln, last);
        top_cp = __shfl_sync(0xffffffffu, cpv, last);
        top_id = base + last;
        t += len;
        continue;
      }
      ++t;
      __syncwarp();   // orders the stores above before the loads below
      Node1 nd[DP_LEVELS];
      int pos[DP_LEVELS];
#pragma unroll
      for (int k = 0; k < DP_LEVELS; ++k) {
        const int g = r >> (5 * k);
        pos[k] = g & 31;
        nd[k] = load_node_cg(&a.level[k][(g & ~31) + lane]);
      }
      // lane-local best over eligible nodes; higher levels hold smaller ranks
      double bc = 0.0;
      int bid = -1, brank = 0x7fffffff;
#pragma unroll
      for (int k = DP_LEVELS - 1; k >= 0; --k) {
        const bool elig = (k == 0) ? (lane <= pos[0]) : (lane < pos[k]);
        if (elig && key_better(nd[k].cum, nd[k].rank, bc, brank)) { bc = nd[k].cum; bid = nd[k].id; brank = nd[k].rank; }
      }
      // warp arg-max on (cum desc, rank asc); all cum >= 0 so the bit patterns order like integers
      const unsigned long long bits = (unsigned long long)__double_as_longlong(bc);
      const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
      const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
      bool alive = hi == mhi;
      const unsigned mlo = __reduce_max_sync(0xffffffffu, alive ? lo : 0u);
      alive = alive && lo == mlo;
      const unsigned mr = __reduce_min_sync(0xffffffffu, alive ? (unsigned)brank : 0xffffffffu);
      alive = alive && (unsigned)brank == mr;
      const int src = __ffs(__ballot_sync(0xffffffffu, alive)) - 1;
      const double pc = __shfl_sync(0xffffffffu, bc, src);
      const int pid = __shfl_sync(0xffffffffu, bid, src);
      const double cum = pc + q;
      // updates: leaf always (a later point on the same rank chains from the earlier one), upper
      // levels when the new key wins.  (A node on the current top's path may be stale-low; the
      // top is written over it when it leaves the block, and it is >= everything in that block.)
      if (lane == pos[0]) {
        Node1 me; me.cum = cum; me.id = p; me.rank = r;
        a.level[0][r] = me;
      }
#pragma unroll
      for (int k = 1; k < DP_LEVELS; ++k) {
        if (lane == pos[k] && key_better(cum, r, nd[k].cum, nd[k].rank)) {
          Node1 me; me.cum = cum; me.id = p; me.rank = r;
          a.level[k][r >> (5 * k)] = me;
        }
      }
      int ln = 1, cpv = p;
      if (pid >= 0) {
        const int4 pm = __ldcg(a.meta + pid);
        ln = pm.y + 1;
        cpv = (ln % DP_CHECK == 0) ? p : pm.z;
      }
      if (lane == 0) { int4 m; m.x = pid; m.y = ln; m.z = cpv; m.w = 0; a.meta[p] = m; }
      if (key_better(cum, r, top_cum, top_rank)) {
        // the old top's ancestors are flushed where the new top lies in another block
        if (lane >= 1 && lane < DP_LEVELS && top_id >= 0 && (r >> my_shift) != (top_rank >> my_shift)) {
          int4 val;
          val.x = __double2loint(top_cum); val.y = __double2hiint(top_cum); val.z = top_id; val.w = top_rank;
          my_level[top_rank >> my_shift] = val;
        }
        top_cum = cum; top_id = p; top_rank = r; top_len = ln; top_cp = cpv;
      }
      __syncwarp();
    }
    __syncwarp();
  }
  if (lane == 0) {
    a.result[0] = top_id;
    a.result[1] = top_id < 0 ? 0 : top_len;
  }
}

// Traceback: thread 0 hops from checkpoint to checkpoint (path_len / 256 dependent steps),
// then every segment between checkpoints is walked by its own thread.
struct TraceArgs {
  const int4 *meta;              // (back, len, cp, -)
  const int32_t *result;
  const int32_t *pt_i, *pt_s, *v_sel;
  int32_t *seg;          // scratch: segment start ids
  int32_t *path_x, *path_y;
};

__global__ void trace1_kernel(TraceArgs a) {
  __shared__ int nseg;
  if (threadIdx.x == 0) {
    int k = 0;
    int cur = a.result[0];
    while (cur >= 0) {
      a.seg[k++] = cur;
      const int c = a.meta[cur].z;  // last node of this segment (checkpoint or chain root)
      cur = a.meta[c].x;
    }
    nseg = k;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    int cur = a.seg[s];
    const int stop = a.meta[cur].z;
    while (true) {
      const int4 m = a.meta[cur];
      const int pos = m.y - 1;
      a.path_x[pos] = a.pt_i[cur];
      a.path_y[pos] = a.v_sel[a.pt_s[cur]];
      if (cur == stop) break;
      cur = m.x;
    }
  }
}

__global__ void fill_nodes_kernel(Node1 *p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { Node1 z; z.cum = 0.0; z.id = -1; z.rank = 0x7fffffff; p[i] = z; }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------
int dab_exclusive_scan(dab_pair *pr, const int32_t *in, int32_t *out, int64_t n, const int32_t *n_dev,
                       int32_t *total_out) {
  dab_ctx *ctx = pr->ctx;
  // tiles cover positions 0 .. n inclusive, so that the tile holding position n exists and writes the total
  const int nblocks = (int)(n / SCAN_TILE) + 1;
  const size_t need = sizeof(unsigned long long) * (size_t)(nblocks + 2);
  if (need > pr->scan_tmp.cap || !pr->scan_tmp.p) {
    DAB_TRY(dab_ensure(ctx, pr->scan_tmp, need + 4096));
    DAB_CUDA(cudaMemsetAsync(pr->scan_tmp.p, 0, pr->scan_tmp.cap, pr->stream));   // epoch 0 = never written; ticket = 0
    pr->scan_epoch = 0;
  }
  ScanArgs sa;
  sa.in = in; sa.out = out; sa.n = n; sa.n_dev = n_dev; sa.total_out = total_out;
  sa.ticket = pr->scan_tmp.as<unsigned int>();
  sa.status = pr->scan_tmp.as<unsigned long long>() + 1;
  pr->scan_epoch = pr->scan_epoch >= 0x3ffffffe ? 1 : pr->scan_epoch + 1;
  sa.epoch = pr->scan_epoch;
  scan_onepass_kernel<<<nblocks, SCAN_THREADS, 0, pr->stream>>>(sa);
  pr->ctx->launches += 1;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

// derived counts after both tracks' not-quiet scans: hashed video frames, this shard's query range
static __global__ void counts_lists_kernel(int32_t *dc, const int32_t *v_off, int64_t v_n, const int32_t *a_off, int64_t a_n,
                                    int64_t row_lo, int64_t row_hi) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int32_t n_vnq = v_off[v_n];
  dc[DC_N_VNQ] = n_vnq;
  dc[DC_N_VSEL] = (n_vnq + 3) / 4;
  dc[DC_N_AQ_ALL] = a_off[a_n];
  const int64_t lo = row_lo < 0 ? 0 : (row_lo > a_n ? a_n : row_lo);
  const int64_t hi = row_hi < lo ? lo : (row_hi > a_n ? a_n : row_hi);
  dc[DC_Q_LO] = a_off[lo];
  dc[DC_Q_HI] = a_off[hi];
  dc[DC_N_Q] = a_off[hi] - a_off[lo];
}

// a scan total against the capacity of the buffer it sizes: clamps the count and raises the overflow bit
static __global__ void check_capacity_kernel(int32_t *dc, int word, int64_t cap, int bit) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if ((int64_t)dc[word] > cap) { dc[word] = 0; atomicOr(&dc[DC_OVERFLOW], bit); }
}

static int prep_track(dab_pair *pr, int track) {
  dab_ctx *ctx = pr->ctx;
  Track &tk = pr->trk[track];
  const int64_t Lmax = tk.Le > tk.L ? tk.Le : tk.L;
  const int64_t Lmin = tk.Le < tk.L ? tk.Le : tk.L;
  const int64_t nc = Lmax - 40;     // stride of the code arrays (energy may have one more)
  tk.n_codes = Lmin - 40;
  DAB_TRY(dab_ensure(ctx, tk.ms, sizeof(double) * (size_t)(5 * Lmax)));
  DAB_TRY(dab_ensure(ctx, tk.nrm, sizeof(double) * (size_t)(3 * nc)));
  DAB_TRY(dab_ensure(ctx, tk.pack, sizeof(uint32_t) * (size_t)(5 * nc)));
  DAB_TRY(dab_ensure(ctx, tk.code, sizeof(int32_t) * (size_t)(5 * nc)));
  PrepArgs pa;
  pa.f32[0] = tk.energy.as<float>(); pa.f32[1] = tk.zc.as<float>();
  pa.f32[2] = tk.b0.as<float>(); pa.f32[3] = tk.b1.as<float>();
  pa.f64 = tk.b2.as<double>();
  pa.len[0] = tk.Le; pa.len[1] = pa.len[2] = pa.len[3] = pa.len[4] = tk.L;
  pa.Lp = Lmin; pa.ms = tk.ms.as<double>(); pa.stride = Lmax;
  dim3 g1((unsigned)cdiv(Lmax, 256 * MS_TILES), 5);
  meansub_kernel<<<g1, 256, 0, pr->stream>>>(pa);
  CodeArgs ca;
  ca.ms = tk.ms.as<double>(); ca.stride = Lmax;
  for (int f = 0; f < 5; ++f) ca.len[f] = pa.len[f];
  ca.nrm = tk.nrm.as<double>(); ca.nstride = nc;
  ca.pack = tk.pack.as<uint32_t>(); ca.code = tk.code.as<int32_t>();
  ca.is_video = track == DAB_TRACK_VIDEO;
  dim3 g2((unsigned)cdiv(nc, 256), 5);
  norm_codes_kernel<<<g2, 256, 0, pr->stream>>>(ca);
  ctx->launches += 2;
  // not-quiet selection: t < len(energy) - 41 with energy[t] > .5
  const int64_t nqn = tk.Le - 41;
  DAB_TRY(dab_ensure(ctx, tk.nq_flag, sizeof(int32_t) * (size_t)(2 * (nqn + 2))));
  DAB_TRY(dab_ensure(ctx, tk.nq_list, sizeof(int32_t) * (size_t)(nqn + 2)));
  int32_t *flag = tk.nq_flag.as<int32_t>();
  int32_t *off = flag + (nqn + 1);
  notquiet_flag_kernel<<<(unsigned)cdiv(nqn, 256), 256, 0, pr->stream>>>(tk.have_gate ? tk.gate.as<float>() : tk.energy.as<float>(), nqn, flag);
  ctx->launches += 1;
  DAB_TRY(dab_exclusive_scan(pr, flag, off, nqn));
  compact_kernel<<<(unsigned)cdiv(nqn, 256), 256, 0, pr->stream>>>(flag, off, nqn, track == DAB_TRACK_VIDEO ? 4 : 1,
                                                                  tk.nq_list.as<int32_t>());
  ctx->launches += 1;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

static int ensure_hann(dab_ctx *ctx) {
  static std::atomic<int> ready[64];
  int dev = 0;
  DAB_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && ready[dev].load(std::memory_order_acquire)) return DAB_OK;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (dev < 64 && ready[dev].load(std::memory_order_acquire)) return DAB_OK;
  double hf[41];
  for (int k = 0; k < 41; ++k) hf[k] = DAB_HANN41_F64[40 - k];
  DAB_CUDA(cudaMemcpyToSymbol(c_hflip, hf, sizeof(hf)));
  if (dev < 64) ready[dev].store(1, std::memory_order_release);
  return DAB_OK;
}

// Capacities of the buffers whose fill level only the device knows (table entries, candidates).  They
// start from what pairs of this size normally need and double when a run overflows (the overflow bit
// comes back with the counts; the stage is then run again - rare, and only ever during the first pairs).
static void default_caps(dab_pair *pr) {
  Track &V = pr->trk[DAB_TRACK_VIDEO], &A = pr->trk[DAB_TRACK_AUDIO];
  const int64_t vsel_ub = (V.Le - 41 + 3) / 4 + 1;
  const int64_t q_ub = A.Le - 41;
  const int64_t want_e = 5 * 16 * vsel_ub + 4096;        // ~10 expanded entries per frame and table
  const int64_t want_c = 2 * q_ub + 4096;                // < 1 candidate per query on matching tracks
  if (pr->cap_entries < want_e) pr->cap_entries = want_e;
  if (pr->cap_cand < want_c) pr->cap_cand = want_c;
}

// ---- match stage, enqueued without a host round trip: prep, codes, lists, tables, gate, scoring ----
int dab_enqueue_stage_a_match(dab_pair *pr, int64_t row_lo, int64_t row_hi) {
  dab_ctx *ctx = pr->ctx;
  DAB_TRY(ensure_hann(ctx));
  Track &V = pr->trk[DAB_TRACK_VIDEO], &A = pr->trk[DAB_TRACK_AUDIO];
  cudaStream_t st = pr->stream;
  pr->n_points1 = pr->n_path1 = 0;
  default_caps(pr);
  const int64_t v_nq = V.Le - 41, a_nq = A.Le - 41;
  const int64_t vsel_ub = (v_nq + 3) / 4 + 1, q_ub = a_nq > 0 ? a_nq : 1;
  const int64_t cap_e = pr->cap_entries, cap_c = pr->cap_cand;

  DAB_TRY(dab_ensure(ctx, pr->counters, sizeof(int32_t) * DC_WORDS));
  int32_t *dc = pr->counters.as<int32_t>();
  DAB_CUDA(cudaEventRecord(pr->ev[4], st));
  DAB_CUDA(cudaMemsetAsync(dc, 0, sizeof(int32_t) * DC_WORDS, st));
  DAB_TRY(prep_track(pr, DAB_TRACK_VIDEO));
  DAB_TRY(prep_track(pr, DAB_TRACK_AUDIO));
  counts_lists_kernel<<<1, 32, 0, st>>>(dc, V.nq_flag.as<int32_t>() + (v_nq + 1), v_nq, A.nq_flag.as<int32_t>() + (a_nq + 1), a_nq,
                                        row_lo, row_hi);
  ctx->launches += 1;
  DAB_CUDA(cudaEventRecord(pr->ev[5], st));
  pr->stats.n_video_frames = V.L; pr->stats.n_audio_frames = A.L;

  // ---- tables ----
  const int64_t nslots = 5LL * DAB_NCODE;
  DAB_TRY(dab_ensure(ctx, pr->tbl_count, sizeof(int32_t) * (size_t)(nslots + 1)));
  DAB_TRY(dab_ensure(ctx, pr->tbl_start, sizeof(int32_t) * (size_t)(nslots + 2)));
  DAB_TRY(dab_ensure(ctx, pr->tbl_items, sizeof(int32_t) * (size_t)(cap_e + 1)));
  DAB_CUDA(cudaMemsetAsync(pr->tbl_count.p, 0, sizeof(int32_t) * (size_t)(nslots + 1), st));
  const int64_t v_nstride = (V.Le > V.L ? V.Le : V.L) - 40;
  const int64_t a_nstride = (A.Le > A.L ? A.Le : A.L) - 40;
  DAB_CUDA(cudaEventRecord(pr->ev[6], st));
  dim3 gt((unsigned)cdiv(vsel_ub, 128), 5);
  DAB_TRY(dab_ensure(ctx, pr->tbl_ecount, sizeof(int32_t) * (size_t)(2 * (5 * vsel_ub + 2))));
  DAB_TRY(dab_ensure(ctx, pr->tbl_pos, sizeof(int32_t) * (size_t)(cap_e + 1)));
  int32_t *ecount = pr->tbl_ecount.as<int32_t>(), *ebase = ecount + (5 * vsel_ub + 1);
  table_expand_kernel<<<gt, 128, 0, st>>>(V.pack.as<uint32_t>(), v_nstride, V.nq_list.as<int32_t>(), dc, vsel_ub, ecount);
  DAB_TRY(dab_exclusive_scan(pr, ecount, ebase, 5 * vsel_ub, nullptr, dc + DC_N_ENTRIES));
  check_capacity_kernel<<<1, 32, 0, st>>>(dc, DC_N_ENTRIES, cap_e, DAB_OVF_ENTRIES);
  table_kernel<false><<<gt, 128, 0, st>>>(V.pack.as<uint32_t>(), v_nstride, V.nq_list.as<int32_t>(), dc, vsel_ub, ebase,
                                          pr->tbl_pos.as<int32_t>(), pr->tbl_count.as<int32_t>(), nullptr, nullptr, cap_e);
  DAB_TRY(dab_exclusive_scan(pr, pr->tbl_count.as<int32_t>(), pr->tbl_start.as<int32_t>(), nslots));
  table_kernel<true><<<gt, 128, 0, st>>>(V.pack.as<uint32_t>(), v_nstride, V.nq_list.as<int32_t>(), dc, vsel_ub, ebase,
                                         pr->tbl_pos.as<int32_t>(), pr->tbl_count.as<int32_t>(), pr->tbl_start.as<int32_t>(),
                                         pr->tbl_items.as<int32_t>(), cap_e);
  ctx->launches += 4;
  DAB_CUDA(cudaEventRecord(pr->ev[7], st));

  // ---- gate: count, scan, fill ----
  DAB_TRY(dab_ensure(ctx, pr->row_count, sizeof(int32_t) * (size_t)(q_ub + 2)));
  DAB_TRY(dab_ensure(ctx, pr->row_off, sizeof(int32_t) * (size_t)(q_ub + 2)));
  DAB_TRY(dab_ensure(ctx, pr->cand_tmp, sizeof(int32_t) * (size_t)(cap_c + 1)));
  DAB_TRY(dab_ensure(ctx, pr->cand_s, sizeof(int32_t) * (size_t)(cap_c + 1)));
  DAB_TRY(dab_ensure(ctx, pr->cand_i, sizeof(int32_t) * (size_t)(cap_c + 1)));
  GateArgs ga;
  ga.a_code = A.code.as<int32_t>(); ga.a_pack = A.pack.as<uint32_t>(); ga.a_nstride = a_nstride;
  ga.a_list = A.nq_list.as<int32_t>(); ga.dc = dc; ga.q_cap = q_ub;
  DAB_TRY(dab_ensure(ctx, pr->v_rec, sizeof(uint4) * 2 * (size_t)(vsel_ub + 1)));
  video_records_kernel<<<(unsigned)cdiv(vsel_ub, 256), 256, 0, st>>>(V.pack.as<uint32_t>(), v_nstride, V.nq_list.as<int32_t>(), dc,
                                                                    pr->v_rec.as<uint4>());
  ctx->launches += 1;
  ga.v_rec = pr->v_rec.as<uint4>();
  ga.start = pr->tbl_start.as<int32_t>(); ga.items = pr->tbl_items.as<int32_t>();
  ga.row_count = pr->row_count.as<int32_t>(); ga.row_off = pr->row_off.as<int32_t>();
  ga.cand_tmp = pr->cand_tmp.as<int32_t>(); ga.cand_s = pr->cand_s.as<int32_t>(); ga.cand_i = pr->cand_i.as<int32_t>();
  ga.cand_cap = cap_c;
  DAB_TRY(dab_ensure(ctx, pr->row_stash, sizeof(int32_t) * GATE_STASH * (size_t)(q_ub + 1)));
  ga.stash = pr->row_stash.as<int32_t>();
  DAB_TRY(dab_ensure(ctx, pr->gate_big, sizeof(int32_t) * (size_t)(q_ub + 1)));
  ga.big_list = pr->gate_big.as<int32_t>();
  ga.big_count = dc + DC_N_BIGROWS;
  DAB_CUDA(cudaEventRecord(pr->ev[8], st));
  // count pass, balanced over bucket entries (see gate_entries_kernel)
  DAB_TRY(dab_ensure(ctx, pr->gate_rec, sizeof(GateQ) * (size_t)(q_ub + 1)));
  DAB_TRY(dab_ensure(ctx, pr->gate_best, sizeof(int32_t) * (size_t)(2 * (q_ub + 2))));
  int32_t *g_best = pr->gate_best.as<int32_t>(), *g_P = g_best + (q_ub + 1);
  gate_plan_kernel<<<(unsigned)cdiv(q_ub, 128), 128, 0, st>>>(ga, pr->gate_rec.as<GateQ>(), g_best);
  DAB_TRY(dab_exclusive_scan(pr, g_best, g_P, q_ub, nullptr, dc + DC_ENUM_LO));
  gate_entries_kernel<<<(unsigned)(8 * ctx->sm_count), GE_THREADS, 0, st>>>(ga, pr->gate_rec.as<GateQ>(), g_P);
  ctx->launches += 2;
  DAB_TRY(dab_exclusive_scan(pr, pr->row_count.as<int32_t>(), pr->row_off.as<int32_t>(), q_ub, nullptr, dc + DC_N_CAND));
  check_capacity_kernel<<<1, 32, 0, st>>>(dc, DC_N_CAND, cap_c, DAB_OVF_CAND);
  gate_fill_small_kernel<<<(unsigned)cdiv(q_ub, 256), 256, 0, st>>>(ga);
  gate_fill_big_kernel<<<(unsigned)(2 * ctx->sm_count), 256, 0, st>>>(ga);
  ctx->launches += 4;
  DAB_CUDA(cudaEventRecord(pr->ev[9], st));

  // ---- scoring + compaction ----
  DAB_CUDA(cudaEventRecord(pr->ev[10], st));
  DAB_TRY(dab_ensure(ctx, pr->cand_q, sizeof(double) * (size_t)(cap_c + 1)));
  DAB_TRY(dab_ensure(ctx, pr->keep_flag, sizeof(int32_t) * (size_t)(cap_c + 2)));
  DAB_TRY(dab_ensure(ctx, pr->keep_off, sizeof(int32_t) * (size_t)(cap_c + 2)));
  DAB_TRY(dab_ensure(ctx, pr->pt_i, sizeof(int32_t) * (size_t)(cap_c + 1)));
  DAB_TRY(dab_ensure(ctx, pr->pt_s, sizeof(int32_t) * (size_t)(cap_c + 1)));
  DAB_TRY(dab_ensure(ctx, pr->pt_q, sizeof(double) * (size_t)(cap_c + 1)));
  ScoreArgs sa;
  sa.cand_s = pr->cand_s.as<int32_t>(); sa.cand_i = pr->cand_i.as<int32_t>(); sa.dc = dc;
  sa.v_sel = V.nq_list.as<int32_t>();
  sa.a_ms = A.ms.as<double>(); sa.v_ms = V.ms.as<double>();
  sa.a_stride = (A.Le > A.L ? A.Le : A.L); sa.v_stride = (V.Le > V.L ? V.Le : V.L);
  sa.a_nrm = A.nrm.as<double>(); sa.v_nrm = V.nrm.as<double>();
  sa.a_nstride = a_nstride; sa.v_nstride = v_nstride;
  sa.qual = pr->cand_q.as<double>(); sa.keep = pr->keep_flag.as<int32_t>();
  const unsigned n_fixed = (unsigned)(4 * ctx->sm_count);     // grid-stride kernels over a device-side count
  score_kernel<<<n_fixed, 128, 0, st>>>(sa);
  DAB_TRY(dab_exclusive_scan(pr, pr->keep_flag.as<int32_t>(), pr->keep_off.as<int32_t>(), cap_c, dc + DC_N_CAND, dc + DC_N_PTS1));
  gather_points_kernel<<<n_fixed, 256, 0, st>>>(
      pr->keep_flag.as<int32_t>(), pr->keep_off.as<int32_t>(), dc, pr->cand_i.as<int32_t>(),
      pr->cand_s.as<int32_t>(), pr->cand_q.as<double>(), pr->pt_i.as<int32_t>(), pr->pt_s.as<int32_t>(),
      pr->pt_q.as<double>());
  ctx->launches += 2;
  DAB_CUDA(cudaEventRecord(pr->ev[11], st));
  for (int s = 2; s <= 5; ++s) pr->ev_used[s] = true;
  pr->cap_points1 = cap_c;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

// ---- DP #1 + traceback over pr->pt_* (dc[DC_N_PTS1] points sorted by (audio frame, video rank)) ----
int dab_enqueue_stage_a_dp(dab_pair *pr) {
  dab_ctx *ctx = pr->ctx;
  Track &V = pr->trk[DAB_TRACK_VIDEO];
  cudaStream_t st = pr->stream;
  int32_t *dc = pr->counters.as<int32_t>();
  const int64_t cap_p = pr->cap_points1 > 0 ? pr->cap_points1 : 1;
  const int64_t vsel_ub = (V.Le - 41 + 3) / 4 + 1;
  DAB_CUDA(cudaEventRecord(pr->ev[12], st));
  if (vsel_ub > (1LL << (5 * DP_LEVELS))) { dab_set_err(ctx, "too many hashed video frames for the DP tree"); return DAB_E_CAPACITY; }
  int64_t lv[DP_LEVELS], tot = 0;
  int64_t m = vsel_ub;
  for (int k = 0; k < DP_LEVELS; ++k) { lv[k] = cdiv(m > 0 ? m : 1, 32) * 32; tot += lv[k]; m = cdiv(m, 32); }
  DAB_TRY(dab_ensure(ctx, pr->tree1, sizeof(Node1) * (size_t)tot));
  fill_nodes_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, st>>>(pr->tree1.as<Node1>(), tot);
  DAB_TRY(dab_ensure(ctx, pr->back1, sizeof(int4) * (size_t)(cap_p + 1)));
  DAB_TRY(dab_ensure(ctx, pr->seglist, sizeof(int32_t) * (size_t)(cap_p / DP_CHECK + 16)));
  DAB_TRY(dab_ensure(ctx, pr->path1_x, sizeof(int32_t) * (size_t)(cap_p + 1)));
  DAB_TRY(dab_ensure(ctx, pr->path1_y, sizeof(int32_t) * (size_t)(cap_p + 1)));
  Dp1Args da;
  da.pt_s = pr->pt_s.as<int32_t>(); da.pt_q = pr->pt_q.as<double>();
  da.n_points = dc + DC_N_PTS1;
  Node1 *base = pr->tree1.as<Node1>();
  for (int k = 0; k < DP_LEVELS; ++k) { da.level[k] = base; base += lv[k]; }
  da.meta = pr->back1.as<int4>();
  da.result = dc + DC_DP1_END;
  dp1_kernel<<<1, 32, 0, st>>>(da);
  TraceArgs ta;
  ta.meta = da.meta; ta.result = da.result;
  ta.pt_i = pr->pt_i.as<int32_t>(); ta.pt_s = pr->pt_s.as<int32_t>(); ta.v_sel = V.nq_list.as<int32_t>();
  ta.seg = pr->seglist.as<int32_t>(); ta.path_x = pr->path1_x.as<int32_t>(); ta.path_y = pr->path1_y.as<int32_t>();
  trace1_kernel<<<1, 256, 0, st>>>(ta);
  ctx->launches += 3;
  DAB_CUDA(cudaEventRecord(pr->ev[13], st));
  pr->ev_used[6] = true;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

// The counts of the stage, copied to the pair's mapped host page by the stream (dab_enqueue_counts) and
// read here once the stream has passed that point.  Returns DAB_E_CAPACITY after growing the capacity
// that overflowed: the caller runs the stage again.
int dab_enqueue_counts(dab_pair *pr) {
  dab_ctx *ctx = pr->ctx;
  DAB_CUDA(dab_readback(pr, pr->h_counters, pr->counters.as<int32_t>(), sizeof(int32_t) * DC_WORDS));
  return DAB_OK;
}

int dab_collect_stage_a(dab_pair *pr, bool with_dp) {
  dab_ctx *ctx = pr->ctx;
  const int32_t *hc = reinterpret_cast<const int32_t *>(pr->h_counters);
  Track &V = pr->trk[DAB_TRACK_VIDEO], &A = pr->trk[DAB_TRACK_AUDIO];
  V.n_list = hc[DC_N_VSEL];
  A.n_list = hc[DC_N_AQ_ALL];
  pr->stats.n_video_selected = hc[DC_N_VSEL];
  pr->stats.n_audio_queries = hc[DC_N_Q];
  pr->stats.n_table_entries = hc[DC_N_ENTRIES];
  pr->stats.n_enumerated = (int64_t)(((uint64_t)(uint32_t)hc[DC_ENUM_HI] << 32) | (uint32_t)hc[DC_ENUM_LO]);
  pr->stats.n_candidates = hc[DC_N_CAND];
  const int ovf = hc[DC_OVERFLOW];
  if (ovf & (DAB_OVF_ENTRIES | DAB_OVF_CAND)) {
    if (ovf & DAB_OVF_ENTRIES) pr->cap_entries *= 2;
    if (ovf & DAB_OVF_CAND) pr->cap_cand *= 2;
    dab_set_err(ctx, "stage A: a device buffer overflowed, capacity doubled");
    return DAB_E_CAPACITY;
  }
  pr->n_points1 = hc[DC_N_PTS1];
  pr->stats.n_points1 = pr->n_points1;
  if (with_dp) {
    pr->n_path1 = hc[DC_DP1_END] < 0 ? 0 : hc[DC_N_PATH1];
    pr->stats.n_path1 = pr->n_path1;
  }
  return DAB_OK;
}

int dab_run_stage_a_match(dab_pair *pr, int64_t row_lo, int64_t row_hi) {
  dab_ctx *ctx = pr->ctx;
  for (int attempt = 0; attempt < 8; ++attempt) {
    DAB_TRY(dab_enqueue_stage_a_match(pr, row_lo, row_hi));
    DAB_TRY(dab_enqueue_counts(pr));
    DAB_CUDA(dab_wait_stream(pr->stream));
    const int rc = dab_collect_stage_a(pr, false);
    if (rc != DAB_E_CAPACITY) return rc;
  }
  return DAB_E_CAPACITY;
}

int dab_run_stage_a_dp(dab_pair *pr) {
  dab_ctx *ctx = pr->ctx;
  if (pr->cap_points1 < pr->n_points1) pr->cap_points1 = pr->n_points1;
  DAB_TRY(dab_enqueue_stage_a_dp(pr));
  DAB_TRY(dab_enqueue_counts(pr));
  DAB_CUDA(dab_wait_stream(pr->stream));
  const int32_t *hc = reinterpret_cast<const int32_t *>(pr->h_counters);
  pr->n_path1 = (pr->n_points1 <= 0 || hc[DC_DP1_END] < 0) ? 0 : hc[DC_N_PATH1];
  pr->stats.n_path1 = pr->n_path1;
  return DAB_OK;
}

int dab_run_stage_a(dab_pair *pr) {
  dab_ctx *ctx = pr->ctx;
  for (int attempt = 0; attempt < 8; ++attempt) {
    DAB_TRY(dab_enqueue_stage_a_match(pr, 0, INT64_MAX));
    DAB_TRY(dab_enqueue_stage_a_dp(pr));
    DAB_TRY(dab_enqueue_counts(pr));
    DAB_CUDA(dab_wait_stream(pr->stream));
    const int rc = dab_collect_stage_a(pr, true);
    if (rc != DAB_E_CAPACITY) return rc;
  }
  return DAB_E_CAPACITY;
}

// ---- row-sharded match stage: exchange of match points -----------------------------------
namespace {
// video frame -> rank in the selected-frame list (binary search; every frame of a point is in the list).
// Also checks what dp1_kernel relies on: points sorted by (audio frame, video frame) and quals > 0.
__global__ void frames_to_ranks_kernel(const int32_t *i_audio, const int32_t *v_frame, const double *qual, int64_t n,
                                       const int32_t *v_sel, int32_t n_sel, int32_t *rank, int32_t *bad) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t v = v_frame[k];
  int lo = 0, hi = n_sel;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (v_sel[mid] < v) lo = mid + 1; else hi = mid;
  }
  if (lo >= n_sel || v_sel[lo] != v) { atomicOr(bad, 1); lo = 0; }
  if (!(qual[k] > 0.0)) atomicOr(bad, 2);
  if (k > 0) {
    const int32_t ip = i_audio[k - 1], ic = i_audio[k];
    if (ip > ic || (ip == ic && v_frame[k - 1] >= v)) atomicOr(bad, 4);
  }
  rank[k] = lo;
}
}  // namespace

static __global__ void set_word_kernel(int32_t *dc, int word, int32_t value) {
  if (threadIdx.x == 0 && blockIdx.x == 0) dc[word] = value;
}

int dab_run_import_points1(dab_pair *pr, const int32_t *i_audio, const int32_t *v_video, const double *qual,
                           int64_t n, int src_on_device) {
  dab_ctx *ctx = pr->ctx;
  Track &V = pr->trk[DAB_TRACK_VIDEO];
  cudaStream_t st = pr->stream;
  const cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  DAB_TRY(dab_ensure(ctx, pr->pt_i, sizeof(int32_t) * (size_t)(n + 1)));
  DAB_TRY(dab_ensure(ctx, pr->pt_s, sizeof(int32_t) * (size_t)(n + 1)));
  DAB_TRY(dab_ensure(ctx, pr->pt_q, sizeof(double) * (size_t)(n + 1)));
  DAB_TRY(dab_ensure(ctx, pr->cand_tmp, sizeof(int32_t) * (size_t)(n + 1)));
  DAB_TRY(dab_ensure(ctx, pr->dpres, sizeof(int32_t) * 16));
  if (n > 0) {
    DAB_CUDA(cudaMemcpyAsync(pr->pt_i.p, i_audio, sizeof(int32_t) * (size_t)n, kind, st));
    DAB_CUDA(cudaMemcpyAsync(pr->cand_tmp.p, v_video, sizeof(int32_t) * (size_t)n, kind, st));
    DAB_CUDA(cudaMemcpyAsync(pr->pt_q.p, qual, sizeof(double) * (size_t)n, kind, st));
    DAB_CUDA(cudaMemsetAsync(pr->dpres.as<int32_t>() + 8, 0, sizeof(int32_t), st));
    frames_to_ranks_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(pr->pt_i.as<int32_t>(), pr->cand_tmp.as<int32_t>(), pr->pt_q.as<double>(), n,
                                                                    V.nq_list.as<int32_t>(), (int32_t)V.n_list, pr->pt_s.as<int32_t>(),
                                                                    pr->dpres.as<int32_t>() + 8);
    ctx->launches += 1;
    DAB_CUDA(dab_readback(pr, &pr->h_counters[22], pr->dpres.as<int32_t>() + 8, sizeof(int32_t)));
    DAB_CUDA(dab_wait_stream(st));
    const int32_t badbits = (int32_t)pr->h_counters[22];
    if (badbits != 0) {
      dab_set_err(ctx, badbits & 1 ? "import_points1: a point's video frame is not one of this pair's hashed video frames"
                        : (badbits & 2 ? "import_points1: match qualities must be > 0"
                                       : "import_points1: points must be sorted by (audio frame, video frame), without duplicates"));
      return DAB_E_ARG;
    }
  }
  set_word_kernel<<<1, 32, 0, st>>>(pr->counters.as<int32_t>(), DC_N_PTS1, (int32_t)n);
  ctx->launches += 1;
  DAB_CUDA(cudaGetLastError());
  pr->n_points1 = n;
  pr->stats.n_points1 = n;
  pr->n_path1 = 0;
  return DAB_OK;
}
