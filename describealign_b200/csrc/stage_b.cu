// Device stage B of align() (reference describealign.py:895-993):
//   corridor scoring  every audio row inside a cluster's +-30 s corridor gets one point on the
//                     cluster's line; qual from the linearly interpolated video features and the
//                     energy gates; first cluster to claim (i, int(j)) wins          (:931-944)
//   dp2               second frontier DP with global / same-cluster / local steps    (:946-983)
//   trace             traceback, rows (j, i, cluster, qual, cum)                     (:985-990)
//
// The reference keys its frontier by the float video coordinate j.  Here every point gets
// rank(j) = 1 + number of corridor rows (over all corridors) whose coordinate is < j, found by
// one binary search per corridor on the same float expression; equal j share a rank and the
// order is preserved, so the frontier becomes a prefix-max tree over ranks with no sort.
#include "common.cuh"

namespace {

constexpr int MAXC = 32;   // corridors that may overlap one audio row

struct __align__(16) P2Rec {     // one pass-2 point, as the DP walks it
  double j, q;
  int32_t i;
  int32_t kf;      // corridor index | P2_* flags
  int32_t cell;    // int(j)
  int32_t ro;      // row offset inside its corridor
};

// static facts about a point, worked out by the (parallel) scoring kernel so that the serial DP
// does not have to:
constexpr int P2_NEAR = 1 << 8;    // a point of another corridor may share its prev_cache cells
constexpr int P2_VIS1 = 1 << 9;    // the corridor's previous point is a prev_cache candidate
constexpr int P2_VIS2 = 1 << 10;   // ... and so is the one before it
constexpr int P2_GAP = 1 << 11;    // the corridor has no point on row i - 1 (its cell was claimed)
constexpr int P2_MAYQ = 1 << 12;   // another corridor has processed rows to the right of j

struct ScoreBArgs {
  const float *a_scaled;   // (n_a, 3)
  const float *v_scaled;   // (n_v, 3)
  int64_t n_a, n_v;
  const dab_corridor *cor;
  int32_t n_cor;
  float a_max, v_max;
  const float *maxes;      // device copy of (a_max, v_max) when they were computed on the device, else null
  int32_t *row_count;
  const int32_t *row_off;
  int32_t *p_i, *p_c, *p_rank;
  P2Rec *rec;
  double *p_j, *p_q;
  int32_t *overflow;
  int32_t want_rank;       // the generic DP needs rank(j); the corridor-state DP does not
  int64_t q_lo, q_hi;      // quals are computed for audio rows q_lo <= i < q_hi only (long-pair sharding); others get 0
};

__device__ __forceinline__ double line_at(const dab_corridor &c, int64_t i) {
  // numpy: slope * x + offset on an int64 arange -> f64 multiply, then add (no fma)
  return __dadd_rn(__dmul_rn(c.slope, (double)i), c.offset);
}

// corridors holding a point on row r, in cluster order, with the cells they claimed (:937-941)
__device__ __forceinline__ int claim_row(const ScoreBArgs &s, int64_t r, int *ks, long long *cells) {
  int n = 0;
  if (r < 0 || r >= s.n_a) return 0;
  for (int k = 0; k < s.n_cor; ++k) {
    const dab_corridor c = s.cor[k];
    if (r < c.lo || r >= c.hi) continue;
    const long long cell = (long long)line_at(c, r);
    bool dup = false;
    for (int m = 0; m < n; ++m) dup = dup || (cells[m] == cell);
    if (dup || n == MAXC) continue;
    ks[n] = k; cells[n] = cell; ++n;
  }
  return n;
}

template <bool FILL>
__global__ void corridor_kernel(ScoreBArgs s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n_a) return;
  double js[MAXC];
  int cs[MAXC], ks[MAXC];
  int n = 0;
  for (int k = 0; k < s.n_cor; ++k) {
    const dab_corridor c = s.cor[k];
    if (i < c.lo || i >= c.hi) continue;
    const double j = line_at(c, i);
    const long long cell = (long long)j;
    bool dup = false;
    for (int m = 0; m < n; ++m) dup = dup || ((long long)js[m] == cell);
    if (dup) continue;     // an earlier (lower-index) cluster already claimed (i, int(j))
    if (n == MAXC) { atomicOr(s.overflow, DAB_OVF_ROWCOR); break; }
    js[n] = j; cs[n] = c.cluster; ks[n] = k; ++n;
  }
  if (!FILL) { s.row_count[i] = n; return; }
  // insertion sort by j (cells are unique within the row, hence so are the j)
  for (int a = 1; a < n; ++a) {
    double j = js[a]; int c = cs[a]; int kk = ks[a]; int b = a - 1;
    while (b >= 0 && js[b] > j) { js[b + 1] = js[b]; cs[b + 1] = cs[b]; ks[b + 1] = ks[b]; --b; }
    js[b + 1] = j; cs[b + 1] = c; ks[b + 1] = kk;
  }
  const int64_t off = s.row_off[i];
  const float a0 = s.a_scaled[i * 3 + 0], a1 = s.a_scaled[i * 3 + 1], a2 = s.a_scaled[i * 3 + 2];
  // which corridors had a point on the two rows before (for the prev_cache visibility flags)
  int ks1[MAXC], ks2[MAXC];
  long long cells1[MAXC], cells2[MAXC];
  const int n1 = claim_row(s, i - 1, ks1, cells1), n2 = claim_row(s, i - 2, ks2, cells2);
  for (int m = 0; m < n; ++m) {
    const double j = js[m];
    const double fl = floor(j);
    const int64_t f = (int64_t)fl;
    const double t = j - fl;
    const double omt = 1.0 - t;
    const float *v0 = s.v_scaled + f * 3, *v1 = v0 + 3;
    const double vl0 = (double)v0[0] * omt + (double)v1[0] * t;
    const double vl1 = (double)v0[1] * omt + (double)v1[1] * t;
    const double vl2 = (double)v0[2] * omt + (double)v1[2] * t;
    double q = 0.0;
    if (i >= s.q_lo && i < s.q_hi) {
      const double t0 = -.5 - log10(1e-4 + fabs((double)a0 - vl0));
      const double t1 = -.5 - log10(1e-4 + fabs((double)a1 - vl1));
      const double t2 = -.5 - log10(1e-4 + fabs((double)a2 - vl2));
      q = (t0 + t1) + t2;
      const float v_max = s.maxes ? s.maxes[1] : s.v_max, a_max = s.maxes ? s.maxes[0] : s.a_max;
      q = q * fmin(fmax((vl0 + 2.5) - (double)v_max, 0.0), 1.0);
      // the audio gate is evaluated in float32 by numpy (f32 array, weak Python scalars)
      float ag = (a0 + 2.5f) - a_max;
      ag = fminf(fmaxf(ag, 0.0f), 1.0f) * 0.1f;
      q = q + (double)ag;
    }
    // rank of j over all corridor rows
    int rank = 1;
    if (s.want_rank) {
      for (int k = 0; k < s.n_cor; ++k) {
        const dab_corridor c = s.cor[k];
        int lo = c.lo, hi = c.hi;        // first row in [lo, hi) whose coordinate is >= j
        while (lo < hi) {
          const int mid = lo + ((hi - lo) >> 1);
          if (line_at(c, mid) < j) lo = mid + 1; else hi = mid;
        }
        rank += lo - c.lo;
      }
    }
    // neighbourhood flag for the corridor-state DP: could a point of another corridor have
    // written one of the prev_cache cells int(j)-2 .. int(j) during rows i-2 .. i?  (A superset
    // test on the lines themselves: it ignores which duplicates were dropped.)
    int near = 0;
    const long long cell = (long long)j;
    for (int k = 0; k < s.n_cor; ++k) {
      if (k == ks[m]) continue;
      const dab_corridor c = s.cor[k];
      for (int64_t r = i - 2; r <= i; ++r) {
        if (r < c.lo || r >= c.hi) continue;
        const long long oc = (long long)line_at(c, r);
        near |= (oc >= cell - 2 && oc <= cell);
      }
    }
    s.p_i[off + m] = (int32_t)i;
    s.p_j[off + m] = j;
    s.p_c[off + m] = cs[m];
    s.p_q[off + m] = q;
    s.p_rank[off + m] = rank;
    // own-corridor prev_cache candidates (describealign.py:966-973): the corridor's points on rows
    // i-1 / i-2, as long as their cell is within two of this one and was not overwritten
    bool has1 = false, has2 = false;
    long long cell_a = 0, cell_b = 0;
    for (int e = 0; e < n1; ++e) if (ks1[e] == ks[m]) { has1 = true; cell_a = cells1[e]; }
    for (int e = 0; e < n2; ++e) if (ks2[e] == ks[m]) { has2 = true; cell_b = cells2[e]; }
    const bool vis1 = has1 ? cell_a >= cell - 2 : (has2 && cell_b >= cell - 2);
    const bool vis2 = has1 && has2 && cell_b >= cell - 2 && cell_b != cell_a;
    const int ro = (int)(i - s.cor[ks[m]].lo);
    const bool gap = !has1 && ro > 0;
    // can the frontier's best entry lie to the right of this point?  Only if another corridor has
    // processed rows (< i) with a larger coordinate.
    bool mayq = false;
    for (int k = 0; k < s.n_cor; ++k) {
      if (k == ks[m]) continue;
      const dab_corridor c = s.cor[k];
      if (c.hi <= c.lo || c.lo > i - 1) continue;
      const int64_t r = i - 1 < c.hi - 1 ? i - 1 : c.hi - 1;
      mayq = mayq || line_at(c, r) > j;
    }
    P2Rec rc;
    rc.j = j; rc.q = q; rc.i = (int32_t)i; rc.cell = (int32_t)cell; rc.ro = ro;
    rc.kf = ks[m] | (near ? P2_NEAR : 0) | (vis1 ? P2_VIS1 : 0) | (vis2 ? P2_VIS2 : 0) | (gap ? P2_GAP : 0) |
            (mayq ? P2_MAYQ : 0);
    s.rec[off + m] = rc;
  }
}

// ------------------------------------------------------------------------------------------
// DP #2.  One warp walks the points in (i, j) order.
// ------------------------------------------------------------------------------------------
struct __align__(16) Node2 {
  double val;     // cum - 1000 of the best point in the subtree, -inf = empty
  int32_t id;     // point id, -1 = the seed (0, 0, -1, 0, 0)
  int32_t rank;
};

struct __align__(16) Cell2 {   // prev_cache row
  double j, q, cum;
  int32_t i, c, id, pad;       // id -2 = never written
};

constexpr int L2N = 5;          // 32^5 ranks
constexpr int CHECK2 = 256;

struct Dp2Args {
  const int32_t *p_i, *p_c, *p_rank;
  const double *p_j, *p_q;
  const int32_t *n_points;   // device count
  Node2 *level[L2N];
  Cell2 *cache;
  double *cb_val;     // per cluster: cum - 50 of its best point
  int32_t *cb_id;
  int32_t *back_id, *len, *cp;
  double *back_cum;
  int32_t *result;    // [0] end id, [1] path length ; double result_val at +2
};

__device__ __forceinline__ unsigned long long order_bits(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ bool key2_better(double va, int ra, double vb, int rb) {
  return va > vb || (va == vb && ra < rb);
}

__global__ void __launch_bounds__(32, 1) dp2_kernel(Dp2Args a) {
  const int lane = threadIdx.x;
  const int n = *a.n_points;
  // the overall best frontier entry: starts as the seed (val 0 at rank 0)
  double top_val = 0.0;
  int top_id = -1, top_rank = 0;
  for (int p = 0; p < n; ++p) {
    const int r = a.p_rank[p];
    const int i = a.p_i[p], c = a.p_c[p];
    const double j = a.p_j[p], q = a.p_q[p];
    Node2 nd[L2N];
    int pos[L2N];
#pragma unroll
    for (int k = 0; k < L2N; ++k) {
      const int g = r >> (5 * k);
      pos[k] = g & 31;
      nd[k] = a.level[k][(g & ~31) + lane];
    }
    const long long ij = (long long)j;
    const long long c_lo = ij - 2 > 0 ? ij - 2 : 0;
    // prev_cache cells ij-2 .. ij and the cluster's best, fetched alongside the tree rows
    Cell2 cell[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const long long idx = ij - 2 + m;
      if (idx >= c_lo) cell[m] = a.cache[idx]; else cell[m].id = -2;
    }
    const double cl_val = a.cb_val[c];
    const int cl_id = a.cb_id[c];

    // (1) global frontier: best val among ranks <= r ; ties -> smaller rank
    double bv = -INFINITY;
    int bid = -2, brank = 0x7fffffff;
#pragma unroll
    for (int k = L2N - 1; k >= 0; --k) {
      const bool elig = (k == 0) ? (lane <= pos[0]) : (lane < pos[k]);
      if (elig && key2_better(nd[k].val, nd[k].rank, bv, brank)) { bv = nd[k].val; bid = nd[k].id; brank = nd[k].rank; }
    }
    const unsigned long long bits = order_bits(bv);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    bool alive = hi == mhi;
    const unsigned mlo = __reduce_max_sync(0xffffffffu, alive ? lo : 0u);
    alive = alive && lo == mlo;
    const unsigned mr = __reduce_min_sync(0xffffffffu, alive ? (unsigned)brank : 0xffffffffu);
    alive = alive && (unsigned)brank == mr;
    const int src = __ffs(__ballot_sync(0xffffffffu, alive)) - 1;
    const double front_val = __shfl_sync(0xffffffffu, bv, src);
    const int front_id = __shfl_sync(0xffffffffu, bid, src);

    double best = front_val;
    int pred = front_id;
    // (2) same-cluster jump
    if (cl_val >= best) { best = cl_val; pred = cl_id; }
    // (3) local steps
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      if (cell[m].id == -2) continue;
      double cum = cell[m].cum;
      if (c != cell[m].c) {
        const double d = (j - cell[m].j) - (double)(i - cell[m].i);
        cum = cum - (100.0 + 100.0 * (d * d));
      }
      if (cell[m].i >= i - 2 && cell[m].j <= j && cum >= best) { best = cum; pred = cell[m].id; }
    }
    const double cum = best + q;
    if (lane == 0) {
      Cell2 me; me.j = j; me.q = q; me.cum = cum; me.i = i; me.c = c; me.id = p; me.pad = 0;
      a.cache[ij] = me;
      a.back_id[p] = pred;
      a.back_cum[p] = best;
      const int ln = pred < 0 ? 1 : a.len[pred] + 1;
      a.len[p] = ln;
      a.cp[p] = (ln % CHECK2 == 0 || pred < 0) ? p : a.cp[pred];
      if (cl_val < cum - 50.0) { a.cb_val[c] = cum - 50.0; a.cb_id[c] = p; }
    }
    const double jump = cum - 1000.0;
    if (front_val < jump) {
      // leaf: keep the earlier entry on equal val (strictly greater replaces)
      if (lane == pos[0] && jump > nd[0].val) {
        Node2 me; me.val = jump; me.id = p; me.rank = r;
        a.level[0][r] = me;
      }
#pragma unroll
      for (int k = 1; k < L2N; ++k) {
        if (lane == pos[k] && key2_better(jump, r, nd[k].val, nd[k].rank)) {
          Node2 me; me.val = jump; me.id = p; me.rank = r;
          a.level[k][r >> (5 * k)] = me;
        }
      }
      if (key2_better(jump, r, top_val, top_rank)) { top_val = jump; top_id = p; top_rank = r; }
    }
    __syncwarp();
  }
  if (lane == 0) {
    a.result[0] = top_id;
    a.result[1] = top_id < 0 ? 0 : a.len[top_id];
    *reinterpret_cast<double *>(a.result + 2) = top_val;
  }
}

struct Trace2Args {
  const int32_t *back_id, *len, *cp, *result;
  const double *back_cum;
  const int32_t *p_i, *p_c;
  const double *p_j, *p_q;
  int32_t *seg;
  double *rows;    // (n_path, 5)
};

__global__ void trace2_kernel(Trace2Args a) {
  __shared__ int nseg;
  if (threadIdx.x == 0) {
    int k = 0;
    int cur = a.result[0];
    while (cur >= 0) {
      a.seg[2 * k] = cur;
      const int c = a.cp[cur];
      const int nxt = a.back_id[c];
      a.seg[2 * k + 1] = c;        // last node of the segment; its predecessor starts the next
      ++k;
      cur = nxt;
    }
    nseg = k;
  }
  __syncthreads();
  const double top_val = *reinterpret_cast<const double *>(a.result + 2);
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    int cur = a.seg[2 * s];
    const int stop = a.seg[2 * s + 1];
    // 5th column of a row = the (penalised) cumulative value its successor started from;
    // for the last row it is the frontier value of the end point
    double col5;
    if (s == 0) col5 = top_val;
    else col5 = a.back_cum[a.seg[2 * (s - 1) + 1]];
    while (true) {
      double *row = a.rows + (int64_t)(a.len[cur] - 1) * 5;
      row[0] = a.p_j[cur]; row[1] = (double)a.p_i[cur]; row[2] = (double)a.p_c[cur]; row[3] = a.p_q[cur]; row[4] = col5;
      if (cur == stop) break;
      col5 = a.back_cum[cur];
      cur = a.back_id[cur];
    }
  }
}

__global__ void init_dp2_kernel(Node2 *nodes, int64_t n_nodes, Cell2 *cache, int64_t n_cells, double *cb_val,
                                int32_t *cb_id, int32_t n_clusters, const int64_t *level_off) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_nodes) {
    Node2 z; z.val = -INFINITY; z.id = -2; z.rank = 0x7fffffff;
    // rank 0 of every level lies on the seed's path: the seed (0, 0, -1, 0, 0) has val 0
    bool seed = false;
#pragma unroll
    for (int k = 0; k < L2N; ++k) seed = seed || (t == level_off[k]);
    if (seed) { z.val = 0.0; z.id = -1; z.rank = 0; }
    nodes[t] = z;
  }
  if (t < n_cells) {
    Cell2 e; e.j = 0; e.q = 0; e.cum = 0; e.i = 0; e.c = -1; e.id = (t == 0) ? -1 : -2; e.pad = 0;
    cache[t] = e;
  }
  if (t < n_clusters) { cb_val[t] = -1000.0; cb_id[t] = -1; }
}

// ------------------------------------------------------------------------------------------
// DP #2, corridor-state formulation (the fast path; the tree kernel above stays as the generic
// fallback for > 32 corridors or lines with non-positive slope).
//
// Every point lies on one of <= 32 corridor lines j = slope * i + offset with slope > 0, so
//   * prev_cache (describealign.py:956, 966-973) only ever yields points of rows i-2 .. i: the
//     last three points of each corridor are enough state ("history");
//   * the frontier (describealign.py:946, 961-962, 976-981) seen from (i, j) is
//         F(j) = max over corridors c' of PM_c'[ #rows of c' processed so far with j' <= j ]
//     where PM_c' is the running maximum of cum - 1000 along corridor c' (ties: earliest row),
//     because j' grows with the row inside a corridor.  PM_c' is append-only, lives in HBM,
//     its last 16 rows in shared memory;
//   * when the frontier's top entry lies at j' <= j it IS F(j): no query at all.  A query is
//     only needed for points below the top's coordinate whose own candidates do not already
//     beat the top value; then lane c' looks PM_c' up (head / recent ring / a small per-lane
//     window cache of older rows, refilled 16 rows at a time) and the warp arg-max-reduces.
// One warp walks the points in (i, j) order; the state of the corridor being extended lives in
// registers (uniform across lanes), the other corridors' state in shared memory.  The dependent
// chain per point is compare + select + one f64 add.  cum values are bit-identical to the
// reference because every cum is still "chosen predecessor value + qual".
// ------------------------------------------------------------------------------------------
struct __align__(16) PmEntry {
  double val;
  int32_t id;
  int32_t pad;
};

struct __align__(16) BackRec {   // per point result: value it started from, predecessor id
  double best;
  int32_t pred;
  int32_t pad;
};

// state of one corridor (shared memory copy; the corridor being extended lives in registers)
struct __align__(16) CorState {
  double h_cum0, h_cum1;                         // history: last three points of this corridor
  int32_t h_row0, h_row1, h_cell0, h_cell1;
  int32_t h_id0, h_id1, h_row2, h_cell2;
  double h_cum2; int32_t h_id2; int32_t filled;  // filled: last PM row written
  double cl_v; int32_t cl_i; int32_t pm_i;       // cluster best (cum - 50), PM head id
  double pm_v; long long pm_base;                // PM head (cum - 1000), first PM row in a.pm
};

constexpr int RING = 16;     // most recent PM rows kept in shared memory
constexpr int RUN_MIN = 8;   // shortest run of chain points worth the lane-parallel commit
constexpr int WAYS = 4;      // window cache: ways per lane
constexpr int WLEN = 16;     // rows per window

struct Dp2LArgs {
  const P2Rec *rec;
  const double *p_j;
  const int32_t *n_points;   // device count
  const dab_corridor *cor;
  int32_t n_cor;
  int64_t pm_off[32];        // first PM row of each corridor
  PmEntry *pm;
  BackRec *back;
  int32_t *result;           // [0] end id ; double top value at +2
  unsigned long long *counters;   // [0] frontier queries, [1] window refills, [2] neighbour points
};

// ------------------------------------------------------------------------------------------
// DP #2, block formulation (the product path).  State and one-point rules are those of
// dp2_lane_kernel (lane l owns corridor l); what changes is that up to 32 consecutive points are
// evaluated together instead of one after the other (tools/dp2_block_model.py is an executable
// statement of the control flow, checked against the sequential rules on full-size inputs):
//   * a corridor whose last cum lies well above the frontier's best value ("leader") cannot take a
//     frontier jump: its owner lane walks ITS points of the block with the local rules only
//     (cluster best, the corridor's previous two points) - all leaders at once, each in its lane;
//   * the frontier's best entry before every point of the block is a prefix arg-max over the
//     leaders' new entries (warp scan, lane = point);
//   * the other corridors ("followers") need the frontier: its best entry when that lies at
//     j' <= j, else F(j) = best entry with j' <= j, computed by the point's lane from the running-max
//     rows written before the block (one L2 read per corridor that can matter) and the leaders'
//     entries of the block; then their owner lanes walk their points with the full rules;
//   * lane = point again: every assumption is checked (leaders: frontier best <= what they chose
//     from; followers: their new entry stays below the frontier's best, and no follower entry of
//     the block could have been a candidate of a queried point).  The points before the first
//     failed check are committed - by induction they are exactly what the sequential rules give -
//     and the failing point, like every point flagged NEAR or GAP, takes the one-point path.
// Every cum is still "chosen predecessor value + qual" in the reference's order, so values and
// decisions are bit-identical; only the evaluation schedule changed.
// ------------------------------------------------------------------------------------------
// order-preserving integer image of a double (no NaNs here): a > b  <=>  dkey(a) > dkey(b), except that
// -0.0 sorts below +0.0 (callers treat "key says greater but the values compare equal" as ambiguous)
__device__ __forceinline__ long long dkey(double x) {
  const long long b = __double_as_longlong(x);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}

constexpr int MIN_BLOCK = 4;            // shortest range worth a block evaluation
constexpr double LEAD_MARGIN = 500.0;   // leader: last cum >= frontier best + margin (a heuristic; checked per point)

__global__ void __launch_bounds__(32, 1) dp2_block_kernel(Dp2LArgs a) {
  __shared__ PmEntry s_ring[RING][32];          // [row & 15][corridor]: last 16 running-max rows
  __shared__ PmEntry s_win[32][WAYS * WLEN];    // per-lane window cache of older rows
  __shared__ P2Rec s_rec[2][32];
  // per corridor: line, extent, running-max rows; and the state at the start of the current block
  __shared__ double s_csl[32], s_cof[32], s_cinv[32], s_headv[32];
  __shared__ int s_clo[32], s_crows[32], s_fill0[32], s_headi[32];
  __shared__ PmEntry *s_pmbase[32];
  __shared__ unsigned s_mask[32];               // points of each corridor in the current block
  // per point of the current block (index = lane of the point)
  __shared__ double s_best[32], s_m[32], s_cum[32], s_pmv[32], s_topv[32], s_topj[32], s_fv[32];
  __shared__ int s_pred[32], s_pmi[32], s_topi[32], s_fi[32];
  __shared__ double s_mb[32];                   // lite walk: largest cum among the corridor's earlier points of the block ...
  __shared__ int s_mbi[32];                     // ... and the lane of that point (-1: none yet)
  __shared__ double s_bmax[32];                 // per corridor: largest new frontier entry of the current block

  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  const int n = *a.n_points;
  const int n_cor = a.n_cor;
  const double NEG = -INFINITY;

  // ---- this lane's corridor -----------------------------------------------------------------
  const bool have = lane < n_cor;
  int lo = 0x7fffffff, rows = 0, cluster = -1;
  double sl = 1.0, of = 0.0, inv = 1.0;
  PmEntry *pm = a.pm;
  if (have) {
    const dab_corridor c = a.cor[lane];
    lo = c.lo; rows = c.hi > c.lo ? c.hi - c.lo : 0; cluster = c.cluster;
    sl = c.slope; of = c.offset; inv = 1.0 / c.slope;
    pm = a.pm + a.pm_off[lane];
  }
  s_csl[lane] = sl; s_cof[lane] = of; s_cinv[lane] = inv;
  s_clo[lane] = lo; s_crows[lane] = rows; s_pmbase[lane] = pm;
  double c0 = NEG, c1 = NEG, c2 = NEG;          // cums of the corridor's last three points
  int id0 = -2, id1 = -2, id2 = -2;
  double cl_v = -1000.0; int cl_i = -1;          // clusters_best_so_far seed (describealign.py:948)
  double pm_v = NEG; int pm_i = -2;              // head of the running maximum of cum - 1000
  int filled = -1;                               // last running-max row written
  int wbase[WAYS], wnext = 0;
#pragma unroll
  for (int w = 0; w < WAYS; ++w) wbase[w] = -0x40000000;
  // frontier top: seed (0, 0, -1, 0, 0) (describealign.py:947); uniform across lanes
  double top_v = 0.0, top_j = 0.0;
  int top_i = -1;
  unsigned n_query = 0, n_refill = 0, n_near = 0, n_blockpts = 0;

  // F(j) seen from a point of corridor k on row i: lane c' contributes PM_c'[rows of c' with j' <= j]
  auto frontier_query = [&](int i, double j, int k, double &fv_out, int &fi_out) {
    double v = NEG;
    int id = -2;
    if (lane == k) { v = 0.0; id = -1; }          // the frontier's seed entry, j' = 0
    else if (have && lo <= i) {
      double est = floor((j - of) * inv) - (double)lo + 1.0;
      int kk = est < 0.0 ? 0 : (est > (double)rows ? rows : (int)est);
      while (kk < rows && __dadd_rn(__dmul_rn(sl, (double)(lo + kk)), of) <= j) ++kk;
      while (kk > 0 && __dadd_rn(__dmul_rn(sl, (double)(lo + kk - 1)), of) > j) --kk;
      const int done = (i + 1 < lo + rows ? i + 1 : lo + rows) - lo;   // rows <= i
      const int idx = kk < done ? kk : done;
      const int f = filled;
      if (idx > 0 && f >= 0) {
        const int x = idx - 1;
        if (x >= f) { v = pm_v; id = pm_i; }
        else if (x > f - RING) { const PmEntry e = s_ring[x & (RING - 1)][lane]; v = e.val; id = e.id; }
        else {
          int hit = -1;
#pragma unroll
          for (int w = 0; w < WAYS; ++w) if (x >= wbase[w] && x < wbase[w] + WLEN) hit = w;
          if (hit < 0) {
            hit = wnext; wnext = (wnext + 1) & (WAYS - 1);
            ++n_refill;
            const PmEntry *src = pm + x;                   // rows x .. x+15 < f are final
#pragma unroll
            for (int e = 0; e < WLEN; ++e) {
              const int4 raw = __ldcg(reinterpret_cast<const int4 *>(src + e));
              *reinterpret_cast<int4 *>(&s_win[lane][hit * WLEN + e]) = raw;
            }
#pragma unroll
            for (int w = 0; w < WAYS; ++w) if (w == hit) wbase[w] = x;
          }
          int wb = 0;
#pragma unroll
          for (int w = 0; w < WAYS; ++w) if (w == hit) wb = wbase[w];
          const PmEntry e = s_win[lane][hit * WLEN + (x - wb)];
          v = e.val; id = e.id;
        }
      }
    }
    // warp arg-max on (val desc, j' asc, id asc)
    const unsigned long long ob = order_bits(v);
    const unsigned hi = (unsigned)(ob >> 32), lo32 = (unsigned)ob;
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    bool alive = hi == mhi;
    const unsigned mlo = __reduce_max_sync(FULL, alive ? lo32 : 0u);
    alive = alive && lo32 == mlo;
    unsigned bal = __ballot_sync(FULL, alive);
    if (__popc(bal) > 1) {
      const double jj = !alive ? INFINITY : (id < 0 ? 0.0 : a.p_j[id]);
      const unsigned long long jb = (unsigned long long)__double_as_longlong(jj);   // jj >= 0
      const unsigned jh = (unsigned)(jb >> 32), jl = (unsigned)jb;
      const unsigned nh = __reduce_min_sync(FULL, alive ? jh : 0xffffffffu);
      alive = alive && jh == nh;
      const unsigned nl = __reduce_min_sync(FULL, alive ? jl : 0xffffffffu);
      alive = alive && jl == nl;
      const unsigned ni = __reduce_min_sync(FULL, alive ? (unsigned)(id + 2) : 0xffffffffu);
      alive = alive && (unsigned)(id + 2) == ni;
      bal = __ballot_sync(FULL, alive);
    }
    const int src = __ffs(bal) - 1;
    fv_out = __shfl_sync(FULL, v, src);
    fi_out = __shfl_sync(FULL, id, src);
  };

  P2Rec rr;
  if (lane < n) rr = a.rec[lane];
  for (int base = 0; base < n; base += 32) {
    const int buf = (base >> 5) & 1;
    s_rec[buf][lane] = rr;
    __syncwarp();
    if (base + 32 + lane < n) rr = a.rec[base + 32 + lane];
    const int cnt = n - base < 32 ? n - base : 32;
    const P2Rec own = s_rec[buf][lane];
    const int own_k = own.kf & 0xff;
    // points a block may not contain: NEAR / GAP points, and the lanes past the last point
    const unsigned hard = __ballot_sync(FULL, lane >= cnt || (own.kf & (P2_NEAR | P2_GAP)) != 0);

    // the owner lane's walk over its points `todo` of the block: local rules, plus the frontier
    // candidate prepared per point (s_topv/j/i, s_fv/s_fi) when with_frontier.  No warp-level
    // operation inside: lanes run it divergently.
    auto own_pass = [&](unsigned todo, const bool with_frontier) {
      double bmax = NEG;
      while (todo) {
        const int u = __ffs(todo) - 1;
        todo &= todo - 1;
        const P2Rec r = s_rec[buf][u];
        const int kf = r.kf;
        double m = cl_v;
        int mi = cl_i;
        const bool t2 = (kf & P2_VIS2) && c1 >= m;
        m = t2 ? c1 : m; mi = t2 ? id1 : mi;
        const bool t1 = (kf & P2_VIS1) && c0 >= m;
        m = t1 ? c0 : m; mi = t1 ? id0 : mi;
        double best = m;
        int pred = mi;
        if (with_frontier) {
          double fv = NEG;
          int fi = -2;
          if (s_topj[u] <= r.j) { fv = s_topv[u]; fi = s_topi[u]; }     // the top entry is F(j) itself
          else if (kf & P2_MAYQ) { fv = s_fv[u]; fi = s_fi[u]; }
          if (fv > m) { best = fv; pred = fi; }
        }
        const double cum = best + r.q;
        const int p = base + u;
        c2 = c1; id2 = id1; c1 = c0; id1 = id0; c0 = cum; id0 = p;
        const double cj = cum - 50.0;
        if (cl_v < cj) { cl_v = cj; cl_i = p; }
        const double jump = cum - 1000.0;
        if (jump > pm_v) { pm_v = jump; pm_i = p; }
        bmax = jump > bmax ? jump : bmax;
        filled = r.ro;
        s_best[u] = best; s_pred[u] = pred; s_m[u] = m; s_cum[u] = cum; s_pmv[u] = pm_v; s_pmi[u] = pm_i;
      }
      s_bmax[lane] = bmax;
    };

    // ---- fast path: a full group whose 32 points all lie on ONE leader corridor and all see the
    //      corridor's previous two points.  If the cluster best is never the choice, cum is the chain
    //      max(c0, c1) + q - the only serial part, 6 instructions per point, every lane runs it - and
    //      predecessor, checks, running maxima and the new state are per-point work plus three warp
    //      scans.  All or nothing: a failed check leaves everything to the general path below.
    bool fast_done = false;
    if (hard == 0u) {
      const int k0 = __shfl_sync(FULL, own_k, 0);
      const int both = P2_VIS1 | P2_VIS2;
      const bool uniform = __all_sync(FULL, own_k == k0 && (own.kf & both) == both);
      const double k_c0 = __shfl_sync(FULL, c0, k0), k_c1 = __shfl_sync(FULL, c1, k0);
      if (uniform && k_c0 >= top_v + LEAD_MARGIN) {
        // max(ca, cb) + q == (ca >= cb ? ca + q : cb + q) bit for bit (rounding is monotone), and in
        // this form the add and the compare both start from ca: one f64 latency per point, not two
        double ca = k_c0, cb = k_c1;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const double q = s_rec[buf][u].q;
          const double x = ca + q, y = cb + q;
          const double cum = ca >= cb ? x : y;
          cb = ca; ca = cum;
          s_cum[u] = cum;                        // same value from every lane
        }
        __syncwarp();
        const double my_cum = s_cum[lane];
        const int k_id0 = __shfl_sync(FULL, id0, k0), k_id1 = __shfl_sync(FULL, id1, k0);
        const double k_clv = __shfl_sync(FULL, cl_v, k0), k_pmv = __shfl_sync(FULL, pm_v, k0);
        const int k_pmi = __shfl_sync(FULL, pm_i, k0);
        double c0u = __shfl_up_sync(FULL, my_cum, 1), c1u = __shfl_up_sync(FULL, my_cum, 2);
        if (lane == 0) { c0u = k_c0; c1u = k_c1; }
        if (lane == 1) c1u = k_c0;
        const int id0u = lane == 0 ? k_id0 : base + lane - 1;
        const int id1u = lane == 0 ? k_id1 : (lane == 1 ? k_id0 : base + lane - 2);
        const bool take0 = c0u >= c1u;             // the later candidate wins ties (describealign.py:966-973)
        const double m_u = take0 ? c0u : c1u;
        const int pred_u = take0 ? id0u : id1u;
        // cluster best before this point: running maximum of cum - 50 over the earlier points
        const double cj = my_cum - 50.0;
        double pcj = cj;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double o = __shfl_up_sync(FULL, pcj, d);
          if (lane >= d && o > pcj) pcj = o;
        }
        const double xcj = __shfl_up_sync(FULL, pcj, 1);
        const double clv_u = (lane > 0 && xcj > k_clv) ? xcj : k_clv;
        // frontier entries of the block: running arg-max of cum - 1000, first occurrence on ties
        // (inside one corridor the earlier point also has the smaller j)
        double sv = my_cum - 1000.0;
        int sl_ = lane;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double ov = __shfl_up_sync(FULL, sv, d);
          const int ol = __shfl_up_sync(FULL, sl_, d);
          if (lane >= d && !(sv > ov)) { sv = ov; sl_ = ol; }
        }
        const double xv = __shfl_up_sync(FULL, sv, 1);
        const double tv_u = (lane > 0 && xv > top_v) ? xv : top_v;     // value of the frontier's best entry before this point
        const bool ok = m_u >= clv_u && tv_u <= m_u;
        if (__all_sync(FULL, ok)) {
          BackRec b; b.best = m_u; b.pred = pred_u; b.pad = 0;
          a.back[base + lane] = b;
          const bool newhead = sv > k_pmv;
          PmEntry en; en.val = newhead ? sv : k_pmv; en.id = newhead ? base + sl_ : k_pmi; en.pad = 0;
          s_pmbase[k0][own.ro] = en;
          if (lane >= 32 - RING) s_ring[own.ro & (RING - 1)][k0] = en;
          // the owner lane's new state
          const double n_c0 = __shfl_sync(FULL, my_cum, 31), n_c1 = __shfl_sync(FULL, my_cum, 30);
          const double n_c2 = __shfl_sync(FULL, my_cum, 29);
          const double maxcj = __shfl_sync(FULL, pcj, 31);
          const int first_cj = __ffs(__ballot_sync(FULL, cj == maxcj)) - 1;
          const double h_v = __shfl_sync(FULL, en.val, 31);
          const int h_i = __shfl_sync(FULL, en.id, 31);
          const int n_ro = __shfl_sync(FULL, own.ro, 31);
          if (lane == k0) {
            c0 = n_c0; c1 = n_c1; c2 = n_c2;
            id0 = base + 31; id1 = base + 30; id2 = base + 29;
            if (maxcj > cl_v) { cl_v = maxcj; cl_i = base + first_cj; }
            pm_v = h_v; pm_i = h_i;
            filled = n_ro;
          }
          // the frontier's best entry after the block
          const double e_v = __shfl_sync(FULL, sv, 31);
          const int e_l = __shfl_sync(FULL, sl_, 31);
          const double e_j = __shfl_sync(FULL, own.j, e_l);
          if (e_v > top_v || (e_v == top_v && e_j < top_j)) { top_v = e_v; top_j = e_j; top_i = base + e_l; }
          n_blockpts += 32;
          fast_done = true;
          __syncwarp();
        }
      }
    }

    int t = fast_done ? cnt : 0;
    while (t < cnt) {
      const unsigned stopbits = hard >> t;
      const int e = stopbits ? t + __ffs(stopbits) - 1 : 32;       // block = points [t, e)
      if (e - t >= MIN_BLOCK) {
        const bool inr = lane >= t && lane < e;
        // ---- who owns which points; leaders and followers --------------------------------------
        const unsigned grp = __match_any_sync(FULL, inr ? own_k : 32 + lane);
        s_mask[lane] = 0u;
        s_fill0[lane] = filled; s_headv[lane] = pm_v; s_headi[lane] = pm_i;
        __syncwarp();
        if (inr) s_mask[own_k] = grp;
        __syncwarp();
        const unsigned mine_mask = s_mask[lane];
        const bool leader = have && c0 >= top_v + LEAD_MARGIN;
        const unsigned leadlanes = __ballot_sync(FULL, leader);
        const bool pt_leader = inr && ((leadlanes >> own_k) & 1u);
        const unsigned lead_pts = __ballot_sync(FULL, pt_leader);
        // the owner's registers at the start of the block (restored if only a prefix commits)
        const double sv_c0 = c0, sv_c1 = c1, sv_c2 = c2, sv_clv = cl_v, sv_pmv = pm_v;
        const int sv_id0 = id0, sv_id1 = id1, sv_id2 = id2, sv_cli = cl_i, sv_pmi = pm_i, sv_filled = filled;

        // The owner lanes' walks come in two forms.  The LITE walk assumes that a point never restarts
        // from its cluster's best (true for 98.6 % of the points of a C2 pair): then cum is just
        // max(c0, c1, frontier candidate) + q, 42-56 instructions per point (ncu), and predecessor ids,
        // running maxima and the assumption itself are worked out afterwards by the points' own lanes.
        // It needs every point of the block to see the corridor's previous two points.  If a point
        // violates the assumption the block is evaluated again with the exact walk (own_pass).
        const int both = P2_VIS1 | P2_VIS2;
        bool use_lite = __all_sync(FULL, !inr || (own.kf & both) == both) != 0;
        auto lite_pass = [&](unsigned todo, const bool with_frontier) {
          // cum = max(la, lb, f) + q, written so that the adds and the compares all start from la (see
          // the fast path): (f > la && f > lb) ? f + q : (la >= lb ? la + q : lb + q); the next point's
          // index and inputs are fetched before the dependent arithmetic of this one.  The warp issues
          // in order and an f64 add or compare takes ~45 cycles, so nothing else in the loop may use
          // the f64 pipe: the only bookkeeping is the running maximum of cum (first occurrence), kept
          // by INTEGER compares of order-preserving keys.  Cluster best and running-max head follow
          // from it per point afterwards (point_head), because rounding is monotone:
          // max fl(cum - c) = fl(max cum - c).  The owner's registers stay untouched.
          double la = c0, lb = c1, mval = NEG;
          long long mkey = dkey(NEG);
          int midx = -1;
          int u = __ffs(todo) - 1;
          todo &= todo - 1;
          double q = s_rec[buf][u].q, f = with_frontier ? s_fv[u] : NEG;
          for (;;) {
            int un = -1;
            double qn = 0.0, fn = NEG;
            if (todo) {
              un = __ffs(todo) - 1;
              todo &= todo - 1;
              qn = s_rec[buf][un].q;
              if (with_frontier) fn = s_fv[un];
            }
            const double x = la + q, y = lb + q, z = f + q;
            const bool pf = f > la && f > lb;
            const double cum = pf ? z : (la >= lb ? x : y);
            lb = la; la = cum;
            s_cum[u] = cum; s_mb[u] = mval; s_mbi[u] = midx;
            const long long k = dkey(cum);
            if (k > mkey) { mkey = k; mval = cum; midx = u; }
            if (un < 0) break;
            u = un; q = qn; f = fn;
          }
          s_bmax[lane] = mval - 1000.0;          // the corridor's largest new frontier entry of the block
        };
        // pre-block state of each point's corridor, for the lite walk's per-point work
        const int ksrc = own_k & 31;
        const double K_c0 = __shfl_sync(FULL, c0, ksrc), K_c1 = __shfl_sync(FULL, c1, ksrc);
        const int K_id0 = __shfl_sync(FULL, id0, ksrc), K_id1 = __shfl_sync(FULL, id1, ksrc);
        const double K_clv = __shfl_sync(FULL, cl_v, ksrc), K_pmv = __shfl_sync(FULL, pm_v, ksrc);
        const int K_pmi = __shfl_sync(FULL, pm_i, ksrc);
        const unsigned present = __ballot_sync(FULL, mine_mask != 0u);    // corridors with points in the block
        // lite: what the exact walk would have recorded as running-max head after this lane's point, and
        // the cluster best before it, from the running maximum of cum the walk left in s_mb / s_mbi.
        // "ambiguous": a new maximum whose rounded cum - 50 or cum - 1000 equals that of the previous
        // maximum - the exact walk keeps the earlier id there (strict compares), so it has to decide.
        double h_clvb = NEG;
        bool h_amb = false;
        auto point_head = [&](double &pmv_out, int &pmi_out) {
          const double mb = s_mb[lane], cum = s_cum[lane];
          const int mbi = s_mbi[lane];
          const bool newmax = dkey(cum) > dkey(mb);
          const double cjm = mb - 50.0, jm = mb - 1000.0;              // -inf stays -inf
          h_clvb = cjm > K_clv ? cjm : K_clv;
          const double ma = newmax ? cum : mb;
          const int mai = newmax ? lane : mbi;
          const double ja = ma - 1000.0;
          const bool hp = ja > K_pmv;
          pmv_out = hp ? ja : K_pmv;
          pmi_out = hp ? base + mai : K_pmi;
          h_amb = newmax && mbi >= 0 && (!(cum > mb) || (cum - 50.0) == cjm || (cum - 1000.0) == jm);
          s_pmv[lane] = pmv_out; s_pmi[lane] = pmi_out;
        };
        double sv, sj, tv_l, tj_l, inc_v, inc_j;
        int si, ti_l, inc_i;
        bool needq;
        double r_m = NEG, r_best = NEG, r_pmv = NEG;   // this lane's point: chosen-from value, best, head after it
        int r_pred = -2, r_pmi = -2;
        for (;;) {
          // ---- A: leaders ----------------------------------------------------------------------
          if (leader && mine_mask) { if (use_lite) lite_pass(mine_mask, false); else own_pass(mine_mask, false); }
          __syncwarp();
          if (use_lite) {
            if (pt_leader) point_head(r_pmv, r_pmi);       // the frontier look-ups below read the leaders' heads
            __syncwarp();
          }

          // ---- B: the frontier's best entry before / after every point (leaders' entries only) ----
          sv = NEG; sj = INFINITY; si = -2;
          if (pt_leader) { sv = s_cum[lane] - 1000.0; sj = own.j; si = base + lane; }
  #pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const double ov = __shfl_up_sync(FULL, sv, d), oj = __shfl_up_sync(FULL, sj, d);
            const int oi = __shfl_up_sync(FULL, si, d);
            // the later entry replaces the earlier one only when strictly better (value, then smaller j)
            const bool keep = lane < d || sv > ov || (sv == ov && sj < oj);
            sv = keep ? sv : ov; sj = keep ? sj : oj; si = keep ? si : oi;
          }
          double xv = __shfl_up_sync(FULL, sv, 1), xj = __shfl_up_sync(FULL, sj, 1);
          int xi = __shfl_up_sync(FULL, si, 1);
          if (lane == 0) { xv = NEG; xj = INFINITY; xi = -2; }
          const bool nb = xv > top_v || (xv == top_v && xj < top_j);
          tv_l = nb ? xv : top_v; tj_l = nb ? xj : top_j;      // top before this lane's point
          ti_l = nb ? xi : top_i;
          const bool nb2 = sv > top_v || (sv == top_v && sj < top_j);
          inc_v = nb2 ? sv : top_v; inc_j = nb2 ? sj : top_j;  // top after this lane's point
          inc_i = nb2 ? si : top_i;
          s_topv[lane] = tv_l; s_topj[lane] = tj_l; s_topi[lane] = ti_l;


          // ---- Q: F(j) for follower points whose top entry lies to their right ----------------------
          needq = inr && !pt_leader && (own.kf & P2_MAYQ) && !(tj_l <= own.j);
          if (__ballot_sync(FULL, needq)) {
            const double clv0 = __shfl_sync(FULL, sv_clv, own_k);      // the point's m is at least this
            if (needq) {
              double bv = 0.0;                 // the frontier's seed entry (j' = 0, id -1)
              int bi = -1;
              auto consider = [&](double v, int id) {
                if (v > bv) { bv = v; bi = id; }
                else if (v == bv) {            // (value desc, j' asc, id asc)
                  const double ja = id < 0 ? 0.0 : a.p_j[id], jb = bi < 0 ? 0.0 : a.p_j[bi];
                  if (ja < jb || (ja == jb && id < bi)) bi = id;
                }
              };
              const double j = own.j;
              const int i = own.i;
              for (int c = 0; c < n_cor; ++c) {
                if (c == own_k) continue;
                // a leader's points of this block: the corridor's coordinate grows with the row, so the
                // qualifying ones (j' <= j) are a prefix; the running-max head recorded at the last of them
                // is the corridor's best entry with j' <= j, rows before the block included
                if ((leadlanes >> c) & 1u) {
                  unsigned mc = s_mask[c] & ((1u << lane) - 1u);
                  int ustar = -1;
                  while (mc) {
                    const int uh = 31 - __clz(mc);
                    if (s_rec[buf][uh].j <= j) { ustar = uh; break; }
                    mc &= ~(1u << uh);
                    if (mc && s_rec[buf][__ffs(mc) - 1].j > j) break;      // none qualifies
                  }
                  if (ustar >= 0) { consider(s_pmv[ustar], s_pmi[ustar]); continue; }
                }
                const int f0 = s_fill0[c], lo2 = s_clo[c], rows2 = s_crows[c];
                const double hv = s_headv[c];
                // rows written before the block; an entry that cannot beat the point's own cluster
                // best can never be chosen (its value would have to exceed m >= clv0)
                if (f0 < 0 || lo2 > i || !(hv > clv0)) continue;
                const double sl2 = s_csl[c], of2 = s_cof[c];
                double est = floor((j - of2) * s_cinv[c]) - (double)lo2 + 1.0;
                int kk = est < 0.0 ? 0 : (est > (double)rows2 ? rows2 : (int)est);
                while (kk < rows2 && __dadd_rn(__dmul_rn(sl2, (double)(lo2 + kk)), of2) <= j) ++kk;
                while (kk > 0 && __dadd_rn(__dmul_rn(sl2, (double)(lo2 + kk - 1)), of2) > j) --kk;
                const int done = (i + 1 < lo2 + rows2 ? i + 1 : lo2 + rows2) - lo2;
                const int idx = kk < done ? kk : done;
                if (idx <= 0) continue;
                const int x = idx - 1;
                if (x >= f0) consider(hv, s_headi[c]);
                else {
                  const int4 raw = __ldcg(reinterpret_cast<const int4 *>(s_pmbase[c] + x));
                  consider(__hiloint2double(raw.y, raw.x), raw.z);
                }
              }
              s_fv[lane] = bv; s_fi[lane] = bi;
            }
          }
          __syncwarp();


          if (use_lite) {
            // the one frontier candidate of every follower point
            if (inr && !pt_leader) {
              double fc = NEG;
              int fci = -2;
              if (tj_l <= own.j) { fc = tv_l; fci = ti_l; }
              else if (own.kf & P2_MAYQ) { fc = s_fv[lane]; fci = s_fi[lane]; }
              s_fv[lane] = fc; s_fi[lane] = fci;
            }
            __syncwarp();
          }

          // ---- C: followers --------------------------------------------------------------------
          if (have && !leader && mine_mask) { if (use_lite) lite_pass(mine_mask, true); else own_pass(mine_mask, true); }
          __syncwarp();
          if (!use_lite) break;

          // ---- lite: each point's lane works out what the exact walk would have recorded -----------
          bool hyp = true;
          if (inr) {
            const unsigned prevm = grp & ((1u << lane) - 1u);     // the corridor's earlier points of the block
            int p1 = -1, p2 = -1;
            if (prevm) {
              p1 = 31 - __clz(prevm);
              const unsigned r = prevm & ~(1u << p1);
              if (r) p2 = 31 - __clz(r);
            }
            const double c0u = p1 >= 0 ? s_cum[p1] : K_c0;
            const int id0u = p1 >= 0 ? base + p1 : K_id0;
            const double c1u = p2 >= 0 ? s_cum[p2] : (p1 >= 0 ? K_c0 : K_c1);
            const int id1u = p2 >= 0 ? base + p2 : (p1 >= 0 ? K_id0 : K_id1);
            const bool take0 = c0u >= c1u;                        // the later candidate wins ties
            const double mm = take0 ? c0u : c1u;
            const int mid = take0 ? id0u : id1u;
            double fc = NEG;
            int fci = -2;
            if (!pt_leader) { fc = s_fv[lane]; fci = s_fi[lane]; }
            const bool tf = fc > mm;
            if (!pt_leader) point_head(r_pmv, r_pmi);
            hyp = (mm >= h_clvb || fc > h_clvb) && !h_amb;         // else the exact walk has to decide
            r_m = mm; r_best = tf ? fc : mm; r_pred = tf ? fci : mid;
          }
          if (__any_sync(FULL, inr && !hyp)) { use_lite = false; __syncwarp(); continue; }
          break;
        }
        if (!use_lite && inr) {
          r_m = s_m[lane]; r_best = s_best[lane]; r_pred = s_pred[lane]; r_pmv = s_pmv[lane]; r_pmi = s_pmi[lane];
        }

        // ---- D: check the assumptions, commit the verified prefix ----------------------------------
        bool ok = true;
        if (inr) {
          if (pt_leader) ok = tv_l <= r_m;
          else {
            ok = (s_cum[lane] - 1000.0) < tv_l;
            if (needq) {
              // follower corridors whose largest new entry could matter at all (rarely any)
              unsigned fcs = present & ~leadlanes & ~(1u << own_k);
              while (fcs) {
                const int c = __ffs(fcs) - 1;
                fcs &= fcs - 1;
                if (s_bmax[c] <= r_m) continue;
                unsigned fm = s_mask[c] & ((1u << lane) - 1u);
                while (fm) {
                  const int u = __ffs(fm) - 1;
                  fm &= fm - 1;
                  if (s_rec[buf][u].j <= own.j && !((s_cum[u] - 1000.0) <= r_m)) ok = false;
                }
              }
            }
          }
        }
        const unsigned badm = __ballot_sync(FULL, inr && !ok);
        const int stop_at = badm ? __ffs(badm) - 1 : e;
        const int glen = stop_at - t;
        if (use_lite) {
          // the lite walk left the owners' registers alone: take the new state from the committed points
          const unsigned mc = mine_mask & (stop_at >= 32 ? FULL : ((1u << stop_at) - 1u));
          const int nmc = __popc(mc);
          if (nmc >= 1) {
            const int l1 = 31 - __clz(mc);
            const unsigned m2 = mc & ~(1u << l1);
            const int l2 = m2 ? 31 - __clz(m2) : l1;
            const unsigned m3 = m2 ? (m2 & ~(1u << l2)) : 0u;
            const int l3 = m3 ? 31 - __clz(m3) : l1;
            const double oc0 = c0, oc1 = c1;
            const int oi0 = id0, oi1 = id1;
            const double x1 = s_cum[l1];
            c0 = x1; id0 = base + l1;
            c1 = nmc >= 2 ? s_cum[l2] : oc0; id1 = nmc >= 2 ? base + l2 : oi0;
            c2 = nmc >= 3 ? s_cum[l3] : (nmc == 2 ? oc0 : oc1);
            id2 = nmc >= 3 ? base + l3 : (nmc == 2 ? oi0 : oi1);
            // running maximum of the committed own points (inclusive of l1), then cluster best and head
            const double mb1 = s_mb[l1];
            const bool nm = dkey(x1) > dkey(mb1);
            const double ma = nm ? x1 : mb1;
            const int mai = nm ? l1 : s_mbi[l1];
            const double cja = ma - 50.0;
            if (cl_v < cja) { cl_v = cja; cl_i = base + mai; }
            pm_v = s_pmv[l1]; pm_i = s_pmi[l1];
            filled = s_rec[buf][l1].ro;
          }
        } else if (stop_at < e) {
          // only a prefix holds: put the owners' registers back and walk the prefix again
          c0 = sv_c0; c1 = sv_c1; c2 = sv_c2; cl_v = sv_clv; pm_v = sv_pmv;
          id0 = sv_id0; id1 = sv_id1; id2 = sv_id2; cl_i = sv_cli; pm_i = sv_pmi; filled = sv_filled;
          const unsigned lim = mine_mask & ((1u << stop_at) - 1u);
          if (lim) own_pass(lim, !leader);
          __syncwarp();
        }
        if (glen > 0) {
          if (lane >= t && lane < stop_at) {
            BackRec b; b.best = r_best; b.pred = r_pred; b.pad = 0;
            a.back[base + lane] = b;
            PmEntry en; en.val = r_pmv; en.id = r_pmi; en.pad = 0;
            s_pmbase[own_k][own.ro] = en;
            // shared ring: the corridor's last RING committed rows
            const unsigned same = s_mask[own_k] & ((stop_at >= 32 ? FULL : ((1u << stop_at) - 1u)));
            const int last_ro = s_rec[buf][31 - __clz(same)].ro;
            if (own.ro > last_ro - RING) s_ring[own.ro & (RING - 1)][own_k] = en;
          }
          top_v = __shfl_sync(FULL, inc_v, stop_at - 1);
          top_j = __shfl_sync(FULL, inc_j, stop_at - 1);
          top_i = __shfl_sync(FULL, inc_i, stop_at - 1);
          n_blockpts += glen;
          __syncwarp();
        }
        t = stop_at;
        if (t >= cnt) break;
      }
      // ---- one point by the sequential rules (dp2_lane_kernel) -----------------------------------
      const int p = base + t;
      const P2Rec pt = s_rec[buf][t];
      ++t;
      const double j = pt.j, q = pt.q;
      const int kf = pt.kf, k = kf & 0xff, ro = pt.ro;
      const bool mine = lane == k;
      double best;
      int pred;
      if (!(kf & P2_NEAR)) {
        double m = cl_v;
        int mi = cl_i;
        const bool t2 = (kf & P2_VIS2) && c1 >= m;
        m = t2 ? c1 : m; mi = t2 ? id1 : mi;
        const bool t1 = (kf & P2_VIS1) && c0 >= m;
        m = t1 ? c0 : m; mi = t1 ? id0 : mi;
        const bool left = top_j <= j;              // the top entry is F(j) itself
        const bool tt = left && top_v > m;
        best = tt ? top_v : m; pred = tt ? top_i : mi;
        if (kf & P2_MAYQ) {
          // the top lies right of the point: F(j) <= top value, needed only if that beats m
          if (__ballot_sync(FULL, mine && !left && m < top_v)) {
            ++n_query;
            double fv; int fi;
            frontier_query(pt.i, j, k, fv, fi);
            if (fv > m) { best = fv; pred = fi; }
          }
        }
      } else {
        // ---- a point of another corridor may sit in this point's prev_cache cells: generic
        //      evaluation over the last three points of every corridor (uniform values)
        ++n_near;
        const int i = pt.i, cell = pt.cell;
        double ub = NEG; int up = -2;
        if (top_j <= j) { ub = top_v; up = top_i; }
        else { ++n_query; frontier_query(i, j, k, ub, up); }
        const double clk = __shfl_sync(FULL, cl_v, k);
        const int cik = __shfl_sync(FULL, cl_i, k);
        const int cluster_k = __shfl_sync(FULL, cluster, k);
        if (clk >= ub) { ub = clk; up = cik; }
        int hr0 = -100, hr1 = -100, hr2 = -100, hc0 = -100, hc1 = -100, hc2 = -100;
        if (id0 >= 0) { const P2Rec r = a.rec[id0]; hr0 = r.i; hc0 = r.cell; }
        if (id1 >= 0) { const P2Rec r = a.rec[id1]; hr1 = r.i; hc1 = r.cell; }
        if (id2 >= 0) { const P2Rec r = a.rec[id2]; hr2 = r.i; hc2 = r.cell; }
#pragma unroll 1
        for (int x = cell - 2; x <= cell; ++x) {
          int brow = -1, bh = 0;
          if (hc0 == x && hr0 > brow) { brow = hr0; bh = 0; }
          if (hc1 == x && hr1 > brow) { brow = hr1; bh = 1; }
          if (hc2 == x && hr2 > brow) { brow = hr2; bh = 2; }
          const int mrow = (int)__reduce_max_sync(FULL, (unsigned)(brow + 1)) - 1;
          if (mrow < 0 || mrow < i - 2) continue;         // nothing written recently enough
          const int src = __ffs(__ballot_sync(FULL, brow == mrow)) - 1;
          const double myc = bh == 0 ? c0 : (bh == 1 ? c1 : c2);
          const int myid = bh == 0 ? id0 : (bh == 1 ? id1 : id2);
          double pc = __shfl_sync(FULL, myc, src);
          const int pid = __shfl_sync(FULL, myid, src);
          const double pj = __shfl_sync(FULL, __dadd_rn(__dmul_rn(sl, (double)mrow), of), src);
          if (__shfl_sync(FULL, cluster, src) != cluster_k) {
            const double d = (j - pj) - (double)(i - mrow);
            pc = pc - (100.0 + 100.0 * (d * d));
          }
          if (pj <= j && pc >= ub) { ub = pc; up = pid; }
        }
        best = ub; pred = up;
      }

      // ---- commit (lane k) ---------------------------------------------------------------------
      const double cum = best + q;
      c2 = mine ? c1 : c2; id2 = mine ? id1 : id2;
      c1 = mine ? c0 : c1; id1 = mine ? id0 : id1;
      c0 = mine ? cum : c0; id0 = mine ? p : id0;
      const double cj = cum - 50.0;
      const bool ucl = mine && cl_v < cj;
      cl_v = ucl ? cj : cl_v; cl_i = ucl ? p : cl_i;
      const double jump = cum - 1000.0;
      if (kf & P2_GAP) {
        // rows without a point (their cell was claimed by an earlier cluster) repeat the head
        if (mine) {
          PmEntry e; e.val = pm_v; e.id = pm_i; e.pad = 0;
          for (int r = filled + 1; r < ro; ++r) { pm[r] = e; s_ring[r & (RING - 1)][lane] = e; }
        }
        __syncwarp();
      }
      const bool upm = mine && jump > pm_v;
      pm_v = upm ? jump : pm_v; pm_i = upm ? p : pm_i;
      filled = mine ? ro : filled;
      if (mine) {
        PmEntry e; e.val = pm_v; e.id = pm_i; e.pad = 0;
        s_ring[ro & (RING - 1)][lane] = e;
        pm[ro] = e;
        BackRec b; b.best = best; b.pred = pred; b.pad = 0;
        a.back[p] = b;
      }
      const double jk = __shfl_sync(FULL, jump, k);
      const bool ut = jk > top_v || (jk == top_v && j < top_j);
      top_v = ut ? jk : top_v; top_j = ut ? j : top_j; top_i = ut ? p : top_i;
    }
    __syncwarp();
  }
  n_refill = __reduce_add_sync(FULL, n_refill);
  if (lane == 0) {
    a.result[0] = top_i;
    *reinterpret_cast<double *>(a.result + 2) = top_v;
    a.counters[0] = n_query; a.counters[1] = n_refill; a.counters[2] = n_near; a.counters[3] = n_blockpts;
  }
}

#include "dp2_scan.cuh"
#include "refine.cuh"

// ------------------------------------------------------------------------------------------
// Traceback by pointer jumping (binary lifting): up[k][p] = 2^k-th predecessor, node n = root.
// depth doubles alongside; the ancestors of the end point are marked level by level from the
// top; each marked point writes its own path row at position depth - 1.
// ------------------------------------------------------------------------------------------
__global__ void lift_init_kernel(const BackRec *back, const int32_t *n_dev, int32_t *up0, int32_t *dep0, int32_t *mark,
                                 const int32_t *result) {
  const int n = *n_dev;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  if (p == n) { up0[p] = n; dep0[p] = 0; mark[p] = 0; return; }
  const int b = back[p].pred;
  up0[p] = b < 0 ? n : b;
  dep0[p] = 1;
  mark[p] = (p == result[0]) ? 1 : 0;
}

__global__ void lift_step_kernel(const int32_t *up_in, const int32_t *dep_in, const int32_t *n_dev, int32_t *up_out,
                                 int32_t *dep_out) {
  const int n = *n_dev;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  const int u = up_in[p];
  up_out[p] = up_in[u];
  dep_out[p] = dep_in[p] + dep_in[u];
}

__global__ void lift_mark_kernel(const int32_t *up_k, const int32_t *n_dev, int32_t *mark) {
  const int n = *n_dev;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n || !mark[p]) return;
  const int u = up_k[p];
  if (u != n) mark[u] = 1;
}

struct EmitArgs {
  const int32_t *mark, *dep, *result;
  const BackRec *back;
  const int32_t *p_i, *p_c;
  const double *p_j, *p_q;
  const int32_t *n_dev;
  double *rows;
  int32_t *n_path;
};

__global__ void lift_emit_kernel(EmitArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= *a.n_dev || !a.mark[p]) return;
  const int pos = a.dep[p] - 1;
  double *row = a.rows + (int64_t)pos * 5;
  row[0] = a.p_j[p]; row[1] = (double)a.p_i[p]; row[2] = (double)a.p_c[p]; row[3] = a.p_q[p];
  // 5th column of a row = the (penalised) value its successor started from (describealign.py:983);
  // the end row carries the frontier value of the end point
  if (pos > 0) a.rows[(int64_t)(pos - 1) * 5 + 4] = a.back[p].best;
  if (p == a.result[0]) {
    row[4] = *reinterpret_cast<const double *>(a.result + 2);
    *a.n_path = pos + 1;
  }
}

}  // namespace

// The corridor-state DP needs: at most 32 corridors, positive slopes, every coordinate >= 3 (so the
// seeded prev_cache cell 0 is never in reach; x_limits keeps lines >= 4, describealign.py:898).
static bool corridor_dp_eligible(const dab_pair *pr, int32_t n_cor) {
  if (pr->ctx->opt_dp2_generic || pr->ctx->opt_dp2_impl == 2 || n_cor > 32) return false;
  for (int k = 0; k < n_cor; ++k) {
    const dab_corridor &c = pr->h_cor[k];
    if (!(c.slope > 0.0)) return false;
    // (device-planned corridors: h_cor holds row ranges that contain the planned ones; the planned lines
    // start at coordinate >= 4 by construction, describealign.py:898)
    if (!pr->b_device_planned && c.hi > c.lo && !(c.slope * (double)c.lo + c.offset >= 3.0)) return false;
  }
  return true;
}

// describealign.py:895-900 on the host, for buffer sizes only (the device computes the ranges it scores)
static void host_x_limits(double x_first, double x_last, double offset, double slope, int64_t n_a, int64_t n_v,
                          long long extend, long long &lo, long long &hi) {
  auto to_ll = [](double x) { return !(x > -9.0e15) ? -9000000000000000LL : (!(x < 9.0e15) ? 9000000000000000LL : (long long)x); };
  lo = to_ll(x_first) - extend;
  if (lo < 0) lo = 0;
  hi = to_ll(x_last) + extend;
  if (hi > n_a - 1) hi = n_a - 1;
  const long long l2 = to_ll(ceil((4.0 - offset) / slope));
  const long long h2 = to_ll(floor(((double)(n_v - 4) - offset) / slope));
  if (l2 > lo) lo = l2;
  if (h2 < hi) hi = h2;
}

// Corridor planning on the device (refine.cuh).  `clusters` is read by an asynchronous copy: it must stay
// valid until the stream has consumed it.  Leaves pr->h_cor with row ranges that CONTAIN the planned ones
// (the refined offset moves the limits by at most 2 / slope rows), for buffer sizes.
int dab_enqueue_plan_corridors(dab_pair *pr, const dab_cluster *clusters, int32_t n_clusters, int64_t n_a, int64_t n_v) {
  dab_ctx *ctx = pr->ctx;
  cudaStream_t st = pr->stream;
  pr->h_cor.clear();
  for (int k = 0; k < n_clusters; ++k) {
    const dab_cluster &c = clusters[k];
    if (c.cluster < 0 || (k > 0 && c.cluster <= clusters[k - 1].cluster) || !(c.slope == c.slope) || c.slope == 0.0) {
      dab_set_err(ctx, "dab_pair_stage_b_clusters: clusters must be in ascending cluster order with a non-zero slope");
      return DAB_E_ARG;
    }
    long long lo, hi;
    host_x_limits(c.x_first, c.x_last, c.offset, c.slope, n_a, n_v, 0, lo, hi);
    dab_corridor out;
    out.cluster = c.cluster; out.lo = 0; out.hi = 0; out.reserved = 0; out.slope = c.slope; out.offset = c.offset;
    if (!(hi < lo + 5)) {
      double xf = c.x_first, xl = c.x_last;
      if (hi > lo + 100) { xf = (double)lo; xl = (double)(hi - 1); }
      // widest range any |coef| < 2 can give: the audio-side limits only
      long long lo2 = (long long)xf - 6300, hi2 = (long long)xl + 6300;
      if (lo2 < 0) lo2 = 0;
      if (hi2 > n_a - 1) hi2 = n_a - 1;
      if (hi2 > lo2) { out.lo = (int32_t)lo2; out.hi = (int32_t)hi2; }
    }
    pr->h_cor.push_back(out);
  }
  const int max_blocks = (int)cdiv(n_a > 0 ? n_a : 1, RF_CHUNK);
  DAB_TRY(dab_ensure(ctx, pr->clusters, sizeof(dab_cluster) * (size_t)(n_clusters + 1)));
  DAB_TRY(dab_ensure(ctx, pr->corridors, sizeof(dab_corridor) * (size_t)(n_clusters + 1)));
  DAB_TRY(dab_ensure(ctx, pr->refine_partial, sizeof(double) * 4 * (size_t)max_blocks * (size_t)(n_clusters + 1)));
  DAB_TRY(dab_ensure(ctx, pr->maxes, sizeof(float) * 4));
  column_max_kernel<<<2, 1024, 0, st>>>(pr->a_scaled.as<float>(), n_a, pr->v_scaled.as<float>(), n_v, pr->maxes.as<float>());
  ctx->launches += 1;
  if (n_clusters > 0) {
    DAB_CUDA(cudaMemcpyAsync(pr->clusters.p, clusters, sizeof(dab_cluster) * (size_t)n_clusters, cudaMemcpyHostToDevice, st));
    RefineArgs ra;
    ra.a_scaled = pr->a_scaled.as<float>(); ra.v_scaled = pr->v_scaled.as<float>();
    ra.n_a = n_a; ra.n_v = n_v;
    ra.cl = pr->clusters.as<dab_cluster>(); ra.n_cl = n_clusters; ra.max_blocks = max_blocks;
    ra.partial = pr->refine_partial.as<double>();
    ra.cor = pr->corridors.as<dab_corridor>();
    refine_partial_kernel<<<dim3((unsigned)max_blocks, (unsigned)n_clusters), RF_THREADS, 0, st>>>(ra);
    refine_plan_kernel<<<(unsigned)cdiv(n_clusters, 64), 64, 0, st>>>(ra);
    ctx->launches += 2;
  }
  DAB_CUDA(cudaGetLastError());
  pr->b_device_planned = true;
  return DAB_OK;
}

// Stage B enqueued on the pair's stream without a host round trip: every (corridor, row) yields at most
// one point, so the sum of the corridor rows (known to the host) bounds all buffers; the actual point
// count stays on the device and every later kernel reads it from there.
int dab_enqueue_stage_b(dab_pair *pr, int32_t n_cor, int32_t n_clusters) {
  DAB_TRY(dab_enqueue_stage_b_points(pr, n_cor, n_clusters, 0, INT64_MAX));
  return dab_enqueue_stage_b_dp(pr, n_cor, n_clusters);
}

// corridor scoring: the sorted point list with flags; quals only for audio rows [q_lo, q_hi)
int dab_enqueue_stage_b_points(dab_pair *pr, int32_t n_cor, int32_t n_clusters, int64_t q_lo, int64_t q_hi) {
  dab_ctx *ctx = pr->ctx;
  cudaStream_t st = pr->stream;
  const int64_t n_a = pr->stats.n_audio_frames, n_v = pr->stats.n_video_frames;  // set by the caller
  (void)n_clusters; (void)n_v;
  pr->n_points2 = pr->n_path2 = 0;
  const bool fast = corridor_dp_eligible(pr, n_cor);
  int64_t pm_off[32], rows = 0;
  for (int k = 0; k < n_cor; ++k) {
    if (k < 32) pm_off[k] = rows;
    rows += pr->h_cor[k].hi > pr->h_cor[k].lo ? pr->h_cor[k].hi - pr->h_cor[k].lo : 0;
  }
  const int64_t cap = rows > 0 ? rows : 1;
  pr->cap_points2 = cap;
  DAB_TRY(dab_ensure(ctx, pr->counters, sizeof(int32_t) * DC_WORDS));
  int32_t *dc = pr->counters.as<int32_t>();
  DAB_CUDA(cudaMemsetAsync(dc + DC_N_PTS2, 0, sizeof(int32_t) * (DC_WORDS - DC_N_PTS2), st));
  ScoreBArgs sb;
  sb.a_scaled = pr->a_scaled.as<float>(); sb.v_scaled = pr->v_scaled.as<float>();
  sb.n_a = n_a; sb.n_v = n_v;
  sb.cor = pr->corridors.as<dab_corridor>(); sb.n_cor = n_cor;
  sb.a_max = pr->b_amax;
  sb.v_max = pr->b_vmax;
  sb.maxes = pr->b_device_planned ? pr->maxes.as<float>() : nullptr;
  sb.want_rank = fast ? 0 : 1;
  sb.q_lo = q_lo; sb.q_hi = q_hi;
  DAB_TRY(dab_ensure(ctx, pr->row2_count, sizeof(int32_t) * (size_t)(n_a + 2)));
  DAB_TRY(dab_ensure(ctx, pr->row2_off, sizeof(int32_t) * (size_t)(n_a + 2)));
  DAB_TRY(dab_ensure(ctx, pr->p2_i, sizeof(int32_t) * (size_t)(cap + 1)));
  DAB_TRY(dab_ensure(ctx, pr->p2_c, sizeof(int32_t) * (size_t)(cap + 1)));
  DAB_TRY(dab_ensure(ctx, pr->p2_rank, sizeof(int32_t) * (size_t)(cap + 1)));
  DAB_TRY(dab_ensure(ctx, pr->p2_k, sizeof(P2Rec) * (size_t)(cap + 1)));
  DAB_TRY(dab_ensure(ctx, pr->p2_j, sizeof(double) * (size_t)(cap + 1)));
  DAB_TRY(dab_ensure(ctx, pr->p2_q, sizeof(double) * (size_t)(cap + 1)));
  DAB_TRY(dab_ensure(ctx, pr->path2, sizeof(double) * 5 * (size_t)(cap + 1)));
  sb.row_count = pr->row2_count.as<int32_t>(); sb.row_off = pr->row2_off.as<int32_t>();
  sb.p_i = pr->p2_i.as<int32_t>(); sb.p_c = pr->p2_c.as<int32_t>(); sb.p_rank = pr->p2_rank.as<int32_t>();
  sb.rec = pr->p2_k.as<P2Rec>();
  sb.p_j = pr->p2_j.as<double>(); sb.p_q = pr->p2_q.as<double>();
  sb.overflow = dc + DC_OVERFLOW_B;
  DAB_CUDA(cudaEventRecord(pr->ev[14], st));
  if (n_a > 0 && n_cor > 0 && rows > 0) {
    const unsigned gb = (unsigned)cdiv(n_a, 128);
    corridor_kernel<false><<<gb, 128, 0, st>>>(sb);
    DAB_TRY(dab_exclusive_scan(pr, sb.row_count, pr->row2_off.as<int32_t>(), n_a, nullptr, dc + DC_N_PTS2));
    corridor_kernel<true><<<gb, 128, 0, st>>>(sb);
    ctx->launches += 2;
  }
  DAB_CUDA(cudaEventRecord(pr->ev[15], st));
  pr->ev_used[7] = true;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

// DP #2 + traceback over the pair's pass-2 points
int dab_enqueue_stage_b_dp(dab_pair *pr, int32_t n_cor, int32_t n_clusters) {
  dab_ctx *ctx = pr->ctx;
  cudaStream_t st = pr->stream;
  const int64_t n_v = pr->stats.n_video_frames;
  const bool fast = corridor_dp_eligible(pr, n_cor);
  int64_t pm_off[32], rows = 0;
  for (int k = 0; k < n_cor; ++k) {
    if (k < 32) pm_off[k] = rows;
    rows += pr->h_cor[k].hi > pr->h_cor[k].lo ? pr->h_cor[k].hi - pr->h_cor[k].lo : 0;
  }
  const int64_t cap = rows > 0 ? rows : 1;
  int32_t *dc = pr->counters.as<int32_t>();
  DAB_CUDA(cudaEventRecord(pr->ev[16], st));
  if (rows > 0 && fast) {
    // ---- scan DP + pointer-jumping traceback ----
    int levels = 1;                      // up[0 .. levels-1], 2^(levels-1) >= number of points
    while ((1LL << (levels - 1)) < cap) ++levels;
    const int64_t np1 = cap + 1;
    DAB_TRY(dab_ensure(ctx, pr->pm2, sizeof(PmEntry) * (size_t)(rows + WLEN + 1)));
    DAB_TRY(dab_ensure(ctx, pr->back2, sizeof(BackRec) * (size_t)(cap + 2)));
    DAB_TRY(dab_ensure(ctx, pr->lift_up, sizeof(int32_t) * (size_t)(levels * np1)));
    DAB_TRY(dab_ensure(ctx, pr->lift_dep, sizeof(int32_t) * (size_t)(3 * np1)));
    Dp2LArgs la;
    la.rec = pr->p2_k.as<P2Rec>(); la.p_j = pr->p2_j.as<double>();
    la.n_points = dc + DC_N_PTS2;
    la.cor = pr->corridors.as<dab_corridor>(); la.n_cor = n_cor;
    for (int k = 0; k < 32; ++k) la.pm_off[k] = k < n_cor ? pm_off[k] : 0;
    la.pm = pr->pm2.as<PmEntry>();
    la.back = pr->back2.as<BackRec>();
    la.result = dc + DC_DP2_END;
    la.counters = reinterpret_cast<unsigned long long *>(dc + DC_DP2_CNT);
    if (pr->ctx->opt_dp2_impl == 1) {
      // the one-warp block kernel of round 1 (kept for cross-checks)
      dp2_block_kernel<<<1, 32, 0, st>>>(la);
    } else {
      static std::atomic<int> attr_set[64];
      if (ctx->device < 64 && !attr_set[ctx->device].load()) {
        DAB_CUDA(cudaFuncSetAttribute(dp2_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanShared)));
        attr_set[ctx->device].store(1);
      }
      dp2_scan_kernel<<<1, SC_T, sizeof(ScanShared), st>>>(la);
    }
    DAB_CUDA(cudaEventRecord(pr->ev[18], st));
    int32_t *up = pr->lift_up.as<int32_t>();
    int32_t *dep0 = pr->lift_dep.as<int32_t>(), *dep1 = dep0 + np1, *mark = dep1 + np1;
    const unsigned gl = (unsigned)cdiv(np1, 256);
    const int32_t *n_dev = dc + DC_N_PTS2;
    lift_init_kernel<<<gl, 256, 0, st>>>(la.back, n_dev, up, dep0, mark, la.result);
    int32_t *din = dep0, *dout = dep1;
    for (int k = 1; k < levels; ++k) {
      lift_step_kernel<<<gl, 256, 0, st>>>(up + (int64_t)(k - 1) * np1, din, n_dev, up + (int64_t)k * np1, dout);
      int32_t *t = din; din = dout; dout = t;
    }
    for (int k = levels - 1; k >= 0; --k) lift_mark_kernel<<<gl, 256, 0, st>>>(up + (int64_t)k * np1, n_dev, mark);
    EmitArgs ea;
    ea.mark = mark; ea.dep = din; ea.back = la.back; ea.result = la.result;
    ea.p_i = pr->p2_i.as<int32_t>(); ea.p_c = pr->p2_c.as<int32_t>(); ea.p_j = la.p_j; ea.p_q = pr->p2_q.as<double>();
    ea.n_dev = n_dev; ea.rows = pr->path2.as<double>(); ea.n_path = dc + DC_N_PATH2;
    lift_emit_kernel<<<gl, 256, 0, st>>>(ea);
    ctx->launches += 3 + 2 * levels;
    DAB_CUDA(cudaEventRecord(pr->ev[17], st));
  } else if (rows > 0) {
    // ---- generic tree DP (more than 32 corridors, or a line with non-positive slope) ----
    const int64_t dom = 1 + rows;        // rank domain: 1 + total corridor rows
    if (dom > (1LL << (5 * L2N))) { dab_set_err(ctx, "pass-2 rank domain too large"); return DAB_E_CAPACITY; }
    int64_t lv[L2N], loff[L2N], tot = 0, m = dom;
    for (int k = 0; k < L2N; ++k) { lv[k] = cdiv(m > 0 ? m : 1, 32) * 32; loff[k] = tot; tot += lv[k]; m = cdiv(m, 32); }
    DAB_TRY(dab_ensure(ctx, pr->tree2, sizeof(Node2) * (size_t)tot + sizeof(int64_t) * L2N));
    DAB_TRY(dab_ensure(ctx, pr->cache2, sizeof(Cell2) * (size_t)(n_v + 1)));
    DAB_TRY(dab_ensure(ctx, pr->back2, (sizeof(double) + sizeof(double)) * (size_t)(n_clusters + cap + 2) + 64));
    DAB_TRY(dab_ensure(ctx, pr->backid2, sizeof(int32_t) * (size_t)(cap + n_clusters + 2)));
    DAB_TRY(dab_ensure(ctx, pr->len2, sizeof(int32_t) * (size_t)(cap + 1)));
    DAB_TRY(dab_ensure(ctx, pr->cp2, sizeof(int32_t) * (size_t)(cap + 1)));
    DAB_TRY(dab_ensure(ctx, pr->seglist, sizeof(int32_t) * (size_t)(2 * (cap / CHECK2 + 16))));
    Node2 *nodes = pr->tree2.as<Node2>();
    int64_t *d_loff = reinterpret_cast<int64_t *>(nodes + tot);
    DAB_CUDA(cudaMemcpyAsync(d_loff, loff, sizeof(int64_t) * L2N, cudaMemcpyHostToDevice, st));
    double *back_cum = pr->back2.as<double>();
    double *cb_val = back_cum + (cap + 1);
    int32_t *back_id = pr->backid2.as<int32_t>();
    int32_t *cb_id = back_id + (cap + 1);
    int64_t span = tot > n_v ? tot : n_v;
    if (n_clusters > span) span = n_clusters;
    init_dp2_kernel<<<(unsigned)cdiv(span, 256), 256, 0, st>>>(nodes, tot, pr->cache2.as<Cell2>(), n_v, cb_val, cb_id,
                                                              n_clusters, d_loff);
    Dp2Args da;
    da.p_i = pr->p2_i.as<int32_t>(); da.p_c = pr->p2_c.as<int32_t>(); da.p_rank = pr->p2_rank.as<int32_t>();
    da.p_j = pr->p2_j.as<double>(); da.p_q = pr->p2_q.as<double>();
    da.n_points = dc + DC_N_PTS2;
    for (int k = 0; k < L2N; ++k) da.level[k] = nodes + loff[k];
    da.cache = pr->cache2.as<Cell2>(); da.cb_val = cb_val; da.cb_id = cb_id;
    da.back_id = back_id; da.len = pr->len2.as<int32_t>(); da.cp = pr->cp2.as<int32_t>();
    da.back_cum = back_cum; da.result = dc + DC_DP2_END;
    dp2_kernel<<<1, 32, 0, st>>>(da);
    DAB_CUDA(cudaEventRecord(pr->ev[18], st));
    Trace2Args ta;
    ta.back_id = back_id; ta.len = da.len; ta.cp = da.cp; ta.result = da.result; ta.back_cum = back_cum;
    ta.p_i = da.p_i; ta.p_c = da.p_c; ta.p_j = da.p_j; ta.p_q = da.p_q;
    ta.seg = pr->seglist.as<int32_t>(); ta.rows = pr->path2.as<double>();
    trace2_kernel<<<1, 256, 0, st>>>(ta);
    ctx->launches += 3;
    DAB_CUDA(cudaEventRecord(pr->ev[17], st));
  } else {
    DAB_CUDA(cudaEventRecord(pr->ev[18], st));
    DAB_CUDA(cudaEventRecord(pr->ev[17], st));
  }
  pr->ev_used[8] = true;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

// quals of all pass-2 points replaced by `q_all` (device memory): the exchange step of a long pair whose
// corridor rows were scored on several GPUs
static __global__ void set_quals_kernel(const double *q_all, const int32_t *n_dev, P2Rec *rec, double *p_q) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= *n_dev) return;
  const double q = q_all[p];
  rec[p].q = q;
  p_q[p] = q;
}

int dab_enqueue_set_quals2(dab_pair *pr, const double *d_q_all) {
  dab_ctx *ctx = pr->ctx;
  const int64_t cap = pr->cap_points2 > 0 ? pr->cap_points2 : 1;
  set_quals_kernel<<<(unsigned)cdiv(cap, 256), 256, 0, pr->stream>>>(d_q_all, pr->counters.as<int32_t>() + DC_N_PTS2,
                                                                    pr->p2_k.as<P2Rec>(), pr->p2_q.as<double>());
  ctx->launches += 1;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}

// first point and number of points of the audio rows [lo, hi) (needs a finished dab_enqueue_stage_b_points)
int dab_row_range_points2(dab_pair *pr, int64_t lo, int64_t hi, int64_t *first, int64_t *count) {
  dab_ctx *ctx = pr->ctx;
  const int64_t n_a = pr->stats.n_audio_frames;
  lo = lo < 0 ? 0 : (lo > n_a ? n_a : lo);
  hi = hi < lo ? lo : (hi > n_a ? n_a : hi);
  int32_t off[2] = {0, 0};
  DAB_CUDA(cudaMemcpyAsync(&off[0], pr->row2_off.as<int32_t>() + lo, sizeof(int32_t), cudaMemcpyDeviceToHost, pr->stream));
  DAB_CUDA(cudaMemcpyAsync(&off[1], pr->row2_off.as<int32_t>() + hi, sizeof(int32_t), cudaMemcpyDeviceToHost, pr->stream));
  DAB_CUDA(dab_wait_stream(pr->stream));
  *first = off[0];
  *count = off[1] - off[0];
  return DAB_OK;
}

int dab_collect_stage_b(dab_pair *pr) {
  dab_ctx *ctx = pr->ctx;
  const int32_t *hc = reinterpret_cast<const int32_t *>(pr->h_counters);
  if (hc[DC_OVERFLOW_B] & DAB_OVF_ROWCOR) { dab_set_err(ctx, "more than 32 corridors overlap one audio row"); return DAB_E_CAPACITY; }
  pr->n_points2 = hc[DC_N_PTS2];
  pr->stats.n_points2 = pr->n_points2;
  pr->n_path2 = (pr->n_points2 > 0 && hc[DC_DP2_END] >= 0) ? hc[DC_N_PATH2] : 0;
  pr->stats.n_path2 = pr->n_path2;
  const int64_t *cnt = reinterpret_cast<const int64_t *>(hc + DC_DP2_CNT);
  pr->stats.n_dp2_queries = cnt[0];
  pr->stats.n_dp2_refills = cnt[1];
  pr->stats.n_dp2_neighbour = cnt[2];
  pr->stats.n_dp2_run_points = cnt[3];
  return DAB_OK;
}

int dab_run_stage_b(dab_pair *pr, int32_t n_cor, int32_t n_clusters) {
  dab_ctx *ctx = pr->ctx;
  DAB_TRY(dab_enqueue_stage_b(pr, n_cor, n_clusters));
  DAB_TRY(dab_enqueue_counts(pr));
  DAB_CUDA(dab_wait_stream(pr->stream));
  return dab_collect_stage_b(pr);
}
