// describealign_b200 - shared declarations of the CUDA implementation (sm_100a).
// All arithmetic files are compiled with -fmad=false: the reference's numpy code never
// contracts a*b+c, and where OpenBLAS does (its ddot kernels), fma() is written explicitly.
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/describealign_b200.h"

#define DAB_NCODE 823543  // 7^7 digit codes per hash table (reference describealign.py:610-628)
#define DAB_WIN 41        // frames in the correlation / norm window (:597)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool pooled = false;   // allocated with cudaMallocAsync
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct dab_ctx {
  int device = 0;
  std::string err;                    // last error; written under err_mu (pairs of one context run on several host threads)
  std::mutex err_mu;
  std::atomic<int64_t> launches{0};   // kernels launched through this context (many host threads)
  int sm_count = 148;
  int opt_dp2_generic = 0;   // force the tree DP for pass 2 (testing)
  int opt_dp_reserve_kb = 0; // dynamic shared memory the pass-2 DP kernel asks for without using it (see dab_set_option)
  int opt_dp2_impl = 0;      // 0 scan kernel (dp2_scan.cuh), 2 generic tree DP (both exact; the tests compare them)
};

// Features and prep data of one track, device resident.
struct Track {
  int64_t S = 0;       // samples per channel
  int ch = 1;
  int64_t L = 0;       // frames = S / 210 (length of zc / bands)
  int64_t Le = 0;      // length of the energy feature (L or L + 1)
  bool have_features = false;
  DevBuf pcm;          // staging for host PCM
  DevBuf energy, zc, b0, b1, b2;   // f32 x4, f64
  DevBuf feat_ticket;  // u32: tile ticket of the persistent feature kernel
  DevBuf gate;         // the separate *_energy argument of align() when it is not features[0] (:629, :657)
  bool have_gate = false;
  // stage A prep (features 0..2 keep ms / nrm for scoring; 3, 4 only feed the codes)
  DevBuf ms;           // f64 [3][Lp]   Lp = min feature length
  DevBuf nrm;          // f64 [3][Lp-40]
  DevBuf pack;         // u32 [5][Lp-40] one digit per nibble (bits 0-2), bit 3 of the nibble = flag (video)
  DevBuf code;         // i32 [5][Lp-40] base-7 code (audio: lookup key)
  DevBuf nq_flag;      // i32 [n] not-quiet flags, then reused
  DevBuf nq_list;      // i32 selected frame list (video: every 4th not-quiet; audio: all not-quiet)
  int64_t n_codes = 0; // Lp - 40
  int64_t n_list = 0;
};

struct dab_pair {
  dab_ctx *ctx = nullptr;
  cudaStream_t stream = nullptr;
  Track trk[2];
  // scan scratch
  DevBuf scan_tmp;            // single-pass scan: ticket counter + tile status words
  unsigned int scan_epoch = 0;
  // tables
  DevBuf tbl_count;   // i32 [5*NCODE + 1]
  DevBuf tbl_start;   // i32 [5*NCODE + 1]
  DevBuf tbl_items;   // i32 sel-ranks
  DevBuf tbl_ecount;  // i32 [2][5 * vsel_cap + 1]: codes per (frame, table) and their exclusive scan
  DevBuf tbl_pos;     // i32 [entries]: position of every expanded entry inside its bucket
  DevBuf v_rec;       // uint4[2] per hashed video frame: its five digit packs (gate)
  // gate / scoring
  DevBuf row_count, row_off;   // i32 [n_queries + 1]
  DevBuf row_stash;            // i32 [n_queries][4]: first candidates of each row (gate count pass)
  DevBuf gate_big;             // i32: rows with more than 4 candidates (work list of the fill pass)
  DevBuf gate_rec, gate_best;  // per query: digit words + its two buckets (32 B); bucket entries to visit and their scan
  DevBuf cand_tmp, cand_s, cand_i;  // i32 [n_cand]
  DevBuf cand_q;               // f64 [n_cand]
  DevBuf keep_flag, keep_off;  // i32 [n_cand + 1]
  DevBuf pt_i, pt_s, pt_q;     // match points of pass 1
  DevBuf counters;             // i64 [16] device counters
  // dp1
  DevBuf tree1;                // Node16 levels
  DevBuf back1, len1, cp1;     // i32 [n_points]
  DevBuf dpres;                // i32/i64 results (end id, path len)
  DevBuf seglist;              // i32 checkpoints
  DevBuf path1_x, path1_y;     // i32
  int64_t n_points1 = 0, n_path1 = 0;
  int64_t cap_entries = 0, cap_cand = 0, cap_points1 = 0;   // capacities of device-sized buffers (grown on overflow)
  bool matched = false;        // tables / hashed-frame list of the current features are built
  // stage B
  DevBuf a_scaled, v_scaled;   // f32 (n,3)
  DevBuf corridors;            // dab_corridor[]
  DevBuf row2_count, row2_off; // i32 [n_audio + 1]
  DevBuf p2_i, p2_c, p2_rank, p2_k;  // i32
  DevBuf pm2, pmoff2;          // corridor-state DP: per-corridor running maxima, row offsets
  DevBuf lift_up, lift_dep;    // pointer-jumping traceback
  std::vector<dab_corridor> h_cor;   // host copy of the corridors of the current stage_b call
  DevBuf p2_j, p2_q;           // f64
  DevBuf tree2;                // frontier tree of pass 2
  DevBuf cache2;               // prev_cache rows
  DevBuf back2;                // per point: predecessor tuple
  DevBuf len2, cp2, backid2;   // i32
  DevBuf path2;                // f64 (n,5)
  int64_t n_points2 = 0, n_path2 = 0;
  int64_t cap_points2 = 0;     // upper bound of pass-2 points of the current stage_b call (sum of corridor rows)
  float b_amax = 0.f, b_vmax = 0.f;   // np.max of the scaled energy columns (describealign.py:908-909)
  int32_t b_n_cor = 0, b_n_clusters = 0;   // of the stage-B call in progress (dab_pair_stage_b_score .. dab_pair_dp2)
  bool b_device_planned = false;      // corridors and energy maxima come from the device (dab_pair_stage_b_clusters)
  DevBuf clusters, refine_partial, maxes;
  dab_stats stats = {};
  cudaEvent_t ev[32] = {};
  bool ev_used[16] = {};
  // pinned host mirror of small results
  int64_t *h_counters = nullptr;       // pinned + mapped host block
  int64_t *d_counters_map = nullptr;   // its device-side address
  // host time spent inside the library's calls for this pair since the last set_pcm (microseconds):
  // [0] set_pcm / set_features, [1] stage A, [2] stage B, [3] get_* copies
  int64_t api_us[4] = {0, 0, 0, 0};
};

inline void dab_set_err(dab_ctx *ctx, const std::string &msg) {
  std::lock_guard<std::mutex> g(ctx->err_mu);
  ctx->err = msg;
}

struct ApiTimer {
  int64_t *slot;
  int64_t t0;
  static int64_t now();
  explicit ApiTimer(int64_t *s) : slot(s), t0(now()) {}
  ~ApiTimer() { *slot += now() - t0; }
};

#define DAB_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      char b__[512];                                                                       \
      snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
               __FILE__, __LINE__);                                                        \
      dab_set_err(ctx, b__);                                                                    \
      return DAB_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

#define DAB_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != DAB_OK) return rc__; \
  } while (0)

// Device buffers are allocated stream-ordered (cudaMallocAsync / cudaFreeAsync on the stream of the
// pair whose API call is running): cudaMalloc / cudaFree would wait for every other pair's kernels.
extern thread_local cudaStream_t dab_t_stream;
struct StreamScope {
  cudaStream_t prev;
  explicit StreamScope(cudaStream_t s) : prev(dab_t_stream) { dab_t_stream = s; }
  ~StreamScope() { dab_t_stream = prev; }
};
int dab_ensure(dab_ctx *ctx, DevBuf &b, size_t bytes);
// Host thread waits for everything queued in `st`.  Mode 0/1: cudaStreamSynchronize (spinning or
// blocking, as the device's schedule flags say); mode 2: cudaStreamQuery + short sleeps with back-off
// (see dab_set_host_wait).
extern int dab_g_wait_mode;
cudaError_t dab_wait_stream(cudaStream_t st);
cudaError_t dab_readback(dab_pair *pr, void *host_dst, const void *dev_src, size_t bytes);
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// exclusive scan of n int32 values; out[n] receives the total (out has n + 1 entries).  n_dev (optional):
// device-side count of valid elements (<= n; launches are sized for n); total_out (optional): device
// word that also receives the total.
int dab_exclusive_scan(dab_pair *pr, const int32_t *in, int32_t *out, int64_t n, const int32_t *n_dev = nullptr,
                       int32_t *total_out = nullptr);

// Counts that only the device knows while a stage is in flight (int32 words of dab_pair::counters).  Every
// kernel of a stage reads the counts it needs from here, so a whole stage is enqueued without a host
// round trip; the block is copied to the pair's mapped host page when the stage ends.
enum {
  DC_N_VNQ = 0,        // not-quiet video frames
  DC_N_AQ_ALL = 1,     // not-quiet audio frames
  DC_N_ENTRIES = 2,    // expanded table entries
  DC_N_VSEL = 3,       // hashed video frames (every 4th not-quiet)
  DC_ENUM_LO = 4, DC_ENUM_HI = 5,   // u64: bucket entries visited by the gate
  DC_N_CAND = 6,
  DC_N_PTS1 = 7,
  DC_DP1_END = 8,      // end id, path length (two consecutive words, written by dp1_kernel)
  DC_N_PATH1 = 9,
  DC_Q_LO = 10, DC_Q_HI = 11, DC_N_Q = 12,    // query range of this shard in the not-quiet audio list
  DC_OVERFLOW = 13,    // bit 0 table entries, bit 1 candidates, bit 2 > 32 corridors on a row, bit 3 pass-2 points
  DC_BAD_IMPORT = 14,
  DC_N_BIGROWS = 15,   // gate rows with more than GATE_STASH candidates
  DC_N_PTS2 = 16,
  DC_OVERFLOW_B = 17,  // stage B: bit 2 more than 32 corridors on one audio row
  DC_DP2_END = 18,     // end id, path length, then the frontier value (double at words 20-21)
  DC_N_PATH2 = 19,
  DC_DP2_CNT = 24,     // 4 x u64 DP-2 work counters
  DC_WORDS = 32
};
enum { DAB_OVF_ENTRIES = 1, DAB_OVF_CAND = 2, DAB_OVF_ROWCOR = 4, DAB_OVF_PTS2 = 8 };

// stage entry points implemented in the .cu files
int dab_run_features(dab_pair *pr, int track, const void *d_pcm, int format);
int dab_run_stage_a(dab_pair *pr);
int dab_run_stage_a_match(dab_pair *pr, int64_t row_lo, int64_t row_hi);   // prep .. match points of audio rows [lo, hi)
int dab_run_stage_a_dp(dab_pair *pr);                                       // DP #1 + traceback over the pair's points
// the same stages, enqueued on the pair's stream without waiting (counts stay on the device); the counts
// block is copied to the mapped host page by dab_enqueue_counts and interpreted by dab_collect_*
int dab_enqueue_stage_a_match(dab_pair *pr, int64_t row_lo, int64_t row_hi);
int dab_enqueue_stage_a_dp(dab_pair *pr);
int dab_enqueue_counts(dab_pair *pr);
int dab_collect_stage_a(dab_pair *pr, bool with_dp);
int dab_enqueue_stage_b(dab_pair *pr, int32_t n_corridors, int32_t n_clusters);
int dab_enqueue_stage_b_points(dab_pair *pr, int32_t n_corridors, int32_t n_clusters, int64_t q_lo, int64_t q_hi);
int dab_enqueue_stage_b_dp(dab_pair *pr, int32_t n_corridors, int32_t n_clusters);
int dab_enqueue_set_quals2(dab_pair *pr, const double *d_q_all);
int dab_row_range_points2(dab_pair *pr, int64_t lo, int64_t hi, int64_t *first, int64_t *count);
int dab_enqueue_plan_corridors(dab_pair *pr, const dab_cluster *clusters, int32_t n_clusters, int64_t n_audio, int64_t n_video);
int dab_collect_stage_b(dab_pair *pr);
int dab_run_import_points1(dab_pair *pr, const int32_t *i_audio, const int32_t *v_video, const double *qual,
                           int64_t n, int src_on_device);
int dab_run_stage_b(dab_pair *pr, int32_t n_corridors, int32_t n_clusters);

#define DAB_LAUNCHED(pr) ((pr)->ctx->launches++)

// ------------------------------------------------------------------------------------------
// OpenBLAS 0.3.30 SkylakeX ddot summation order (what np.dot / np.convolve do on f64 in the
// reference, SURVEY.md B.2 iv): blocks of 32 through 4 x 8 FMA lanes folded to 4 x 4, blocks
// of 16 through 4 x 4 FMA lanes, tail as one FMA chain.  X(i), Y(i) are accessor macros.
// ------------------------------------------------------------------------------------------
template <typename FX, typename FY>
__device__ __forceinline__ double ddot_skx(FX X, FY Y, int n) {
  double a[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int l = 0; l < 4; ++l) a[k][l] = 0.0;
  int n1 = n & ~15;
  int n32 = n1 & ~31;
  int i = 0;
  if (n32) {
    double z[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int l = 0; l < 8; ++l) z[k][l] = 0.0;
    for (; i < n32; i += 32) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) z[k][l] = fma(X(i + 8 * k + l), Y(i + 8 * k + l), z[k][l]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int l = 0; l < 4; ++l) a[k][l] = z[k][l] + z[k][l + 4];
  }
  for (; i < n1; i += 16) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int l = 0; l < 4; ++l) a[k][l] = fma(X(i + 4 * k + l), Y(i + 4 * k + l), a[k][l]);
  }
  double s0 = ((a[0][0] + a[1][0]) + a[2][0]) + a[3][0];
  double s1 = ((a[0][1] + a[1][1]) + a[2][1]) + a[3][1];
  double s2 = ((a[0][2] + a[1][2]) + a[2][2]) + a[3][2];
  double s3 = ((a[0][3] + a[1][3]) + a[2][3]) + a[3][3];
  double dot = (s0 + s2) + (s1 + s3);
  for (; i < n; ++i) dot = fma(Y(i), X(i), dot);
  return dot;
}

// Fixed n = 41 specialisation: one 32-block, no 16-block, 9-element tail.
template <typename FX, typename FY>
__device__ __forceinline__ double ddot41_skx(FX X, FY Y) {
  double z[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int l = 0; l < 8; ++l) z[k][l] = fma(X(8 * k + l), Y(8 * k + l), 0.0);
  double s[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    double a0 = z[0][l] + z[0][l + 4];
    double a1 = z[1][l] + z[1][l + 4];
    double a2 = z[2][l] + z[2][l + 4];
    double a3 = z[3][l] + z[3][l + 4];
    s[l] = ((a0 + a1) + a2) + a3;
  }
  double dot = (s[0] + s[2]) + (s[1] + s[3]);
#pragma unroll
  for (int i = 32; i < 41; ++i) dot = fma(Y(i), X(i), dot);
  return dot;
}
