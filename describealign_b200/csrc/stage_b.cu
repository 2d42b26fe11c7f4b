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

struct ScoreBArgs {
  const float *a_scaled;   // (n_a, 3)
  const float *v_scaled;   // (n_v, 3)
  int64_t n_a, n_v;
  const dab_corridor *cor;
  int32_t n_cor;
  float a_max, v_max;
  int32_t *row_count;
  const int32_t *row_off;
  int32_t *p_i, *p_c, *p_rank;
  double *p_j, *p_q;
  int32_t *overflow;
};

__device__ __forceinline__ double line_at(const dab_corridor &c, int64_t i) {
  // numpy: slope * x + offset on an int64 arange -> f64 multiply, then add (no fma)
  return __dadd_rn(__dmul_rn(c.slope, (double)i), c.offset);
}

template <bool FILL>
__global__ void corridor_kernel(ScoreBArgs s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n_a) return;
  double js[MAXC];
  int cs[MAXC];
  int n = 0;
  for (int k = 0; k < s.n_cor; ++k) {
    const dab_corridor c = s.cor[k];
    if (i < c.lo || i >= c.hi) continue;
    const double j = line_at(c, i);
    const long long cell = (long long)j;
    bool dup = false;
    for (int m = 0; m < n; ++m) dup = dup || ((long long)js[m] == cell);
    if (dup) continue;     // an earlier (lower-index) cluster already claimed (i, int(j))
    if (n == MAXC) { atomicExch(s.overflow, 1); break; }
    js[n] = j; cs[n] = c.cluster; ++n;
  }
  if (!FILL) { s.row_count[i] = n; return; }
  // insertion sort by j (cells are unique within the row, hence so are the j)
  for (int a = 1; a < n; ++a) {
    double j = js[a]; int c = cs[a]; int b = a - 1;
    while (b >= 0 && js[b] > j) { js[b + 1] = js[b]; cs[b + 1] = cs[b]; --b; }
    js[b + 1] = j; cs[b + 1] = c;
  }
  const int64_t off = s.row_off[i];
  const float a0 = s.a_scaled[i * 3 + 0], a1 = s.a_scaled[i * 3 + 1], a2 = s.a_scaled[i * 3 + 2];
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
    const double t0 = -.5 - log10(1e-4 + fabs((double)a0 - vl0));
    const double t1 = -.5 - log10(1e-4 + fabs((double)a1 - vl1));
    const double t2 = -.5 - log10(1e-4 + fabs((double)a2 - vl2));
    double q = (t0 + t1) + t2;
    q = q * fmin(fmax((vl0 + 2.5) - (double)s.v_max, 0.0), 1.0);
    // the audio gate is evaluated in float32 by numpy (f32 array, weak Python scalars)
    float ag = (a0 + 2.5f) - s.a_max;
    ag = fminf(fmaxf(ag, 0.0f), 1.0f) * 0.1f;
    q = q + (double)ag;
    // rank of j over all corridor rows
    int rank = 1;
    for (int k = 0; k < s.n_cor; ++k) {
      const dab_corridor c = s.cor[k];
      int lo = c.lo, hi = c.hi;        // first row in [lo, hi) whose coordinate is >= j
      while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (line_at(c, mid) < j) lo = mid + 1; else hi = mid;
      }
      rank += lo - c.lo;
    }
    s.p_i[off + m] = (int32_t)i;
    s.p_j[off + m] = j;
    s.p_c[off + m] = cs[m];
    s.p_q[off + m] = q;
    s.p_rank[off + m] = rank;
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
  int32_t n_points;
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
  const int n = a.n_points;
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

}  // namespace

int dab_run_stage_b(dab_pair *pr, int32_t n_cor, int32_t n_clusters) {
  dab_ctx *ctx = pr->ctx;
  cudaStream_t st = pr->stream;
  const int64_t n_a = pr->stats.n_audio_frames, n_v = pr->stats.n_video_frames;  // set by the caller
  pr->n_points2 = pr->n_path2 = 0;
  ScoreBArgs sb;
  sb.a_scaled = pr->a_scaled.as<float>(); sb.v_scaled = pr->v_scaled.as<float>();
  sb.n_a = n_a; sb.n_v = n_v;
  sb.cor = pr->corridors.as<dab_corridor>(); sb.n_cor = n_cor;
  sb.a_max = reinterpret_cast<float *>(pr->h_counters + 8)[0];
  sb.v_max = reinterpret_cast<float *>(pr->h_counters + 8)[1];
  DAB_TRY(dab_ensure(ctx, pr->row2_count, sizeof(int32_t) * (size_t)(n_a + 2)));
  DAB_TRY(dab_ensure(ctx, pr->row2_off, sizeof(int32_t) * (size_t)(n_a + 2)));
  DAB_TRY(dab_ensure(ctx, pr->dpres, sizeof(int32_t) * 8));
  DAB_CUDA(cudaMemsetAsync(pr->dpres.p, 0, sizeof(int32_t) * 8, st));
  sb.row_count = pr->row2_count.as<int32_t>(); sb.row_off = pr->row2_off.as<int32_t>();
  sb.p_i = sb.p_c = sb.p_rank = nullptr; sb.p_j = sb.p_q = nullptr;
  sb.overflow = pr->dpres.as<int32_t>() + 6;
  DAB_CUDA(cudaEventRecord(pr->ev[14], st));
  int64_t n_pts = 0;
  if (n_a > 0 && n_cor > 0) {
    const unsigned gb = (unsigned)cdiv(n_a, 128);
    corridor_kernel<false><<<gb, 128, 0, st>>>(sb);
    ctx->launches += 1;
    DAB_TRY(dab_exclusive_scan(pr, sb.row_count, pr->row2_off.as<int32_t>(), n_a));
    DAB_CUDA(cudaMemcpyAsync(&pr->h_counters[9], pr->row2_off.as<int32_t>() + n_a, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DAB_CUDA(cudaMemcpyAsync(&pr->h_counters[10], pr->dpres.as<int32_t>() + 6, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DAB_CUDA(cudaStreamSynchronize(st));
    if ((int32_t)pr->h_counters[10] != 0) { ctx->err = "more than 32 corridors overlap one audio row"; return DAB_E_CAPACITY; }
    n_pts = (int32_t)pr->h_counters[9];
    DAB_TRY(dab_ensure(ctx, pr->p2_i, sizeof(int32_t) * (size_t)(n_pts + 1)));
    DAB_TRY(dab_ensure(ctx, pr->p2_c, sizeof(int32_t) * (size_t)(n_pts + 1)));
    DAB_TRY(dab_ensure(ctx, pr->p2_rank, sizeof(int32_t) * (size_t)(n_pts + 1)));
    DAB_TRY(dab_ensure(ctx, pr->p2_j, sizeof(double) * (size_t)(n_pts + 1)));
    DAB_TRY(dab_ensure(ctx, pr->p2_q, sizeof(double) * (size_t)(n_pts + 1)));
    sb.p_i = pr->p2_i.as<int32_t>(); sb.p_c = pr->p2_c.as<int32_t>(); sb.p_rank = pr->p2_rank.as<int32_t>();
    sb.p_j = pr->p2_j.as<double>(); sb.p_q = pr->p2_q.as<double>();
    if (n_pts > 0) {
      corridor_kernel<true><<<gb, 128, 0, st>>>(sb);
      ctx->launches += 1;
    }
  }
  pr->n_points2 = n_pts;
  pr->stats.n_points2 = n_pts;
  DAB_CUDA(cudaEventRecord(pr->ev[15], st));

  DAB_CUDA(cudaEventRecord(pr->ev[16], st));
  int64_t n_path = 0;
  if (n_pts > 0) {
    // rank domain: 1 + total corridor rows
    int64_t dom = 1;
    {
      // corridors were validated by the caller; sizes come from the host copy kept in h_counters[11]
      dom += pr->h_counters[11];
    }
    if (dom > (1LL << (5 * L2N))) { ctx->err = "pass-2 rank domain too large"; return DAB_E_CAPACITY; }
    int64_t lv[L2N], loff[L2N], tot = 0, m = dom;
    for (int k = 0; k < L2N; ++k) { lv[k] = cdiv(m > 0 ? m : 1, 32) * 32; loff[k] = tot; tot += lv[k]; m = cdiv(m, 32); }
    DAB_TRY(dab_ensure(ctx, pr->tree2, sizeof(Node2) * (size_t)tot + sizeof(int64_t) * L2N));
    DAB_TRY(dab_ensure(ctx, pr->cache2, sizeof(Cell2) * (size_t)(n_v + 1)));
    DAB_TRY(dab_ensure(ctx, pr->back2, (sizeof(double) + sizeof(double)) * (size_t)(n_clusters + n_pts + 2) + 64));
    DAB_TRY(dab_ensure(ctx, pr->backid2, sizeof(int32_t) * (size_t)(n_pts + n_clusters + 2)));
    DAB_TRY(dab_ensure(ctx, pr->len2, sizeof(int32_t) * (size_t)(n_pts + 1)));
    DAB_TRY(dab_ensure(ctx, pr->cp2, sizeof(int32_t) * (size_t)(n_pts + 1)));
    DAB_TRY(dab_ensure(ctx, pr->seglist, sizeof(int32_t) * (size_t)(2 * (n_pts / CHECK2 + 16))));
    DAB_TRY(dab_ensure(ctx, pr->path2, sizeof(double) * 5 * (size_t)(n_pts + 1)));
    Node2 *nodes = pr->tree2.as<Node2>();
    int64_t *d_loff = reinterpret_cast<int64_t *>(nodes + tot);
    DAB_CUDA(cudaMemcpyAsync(d_loff, loff, sizeof(int64_t) * L2N, cudaMemcpyHostToDevice, st));
    double *back_cum = pr->back2.as<double>();
    double *cb_val = back_cum + (n_pts + 1);
    int32_t *back_id = pr->backid2.as<int32_t>();
    int32_t *cb_id = back_id + (n_pts + 1);
    int64_t span = tot > n_v ? tot : n_v;
    if (n_clusters > span) span = n_clusters;
    init_dp2_kernel<<<(unsigned)cdiv(span, 256), 256, 0, st>>>(nodes, tot, pr->cache2.as<Cell2>(), n_v, cb_val, cb_id,
                                                              n_clusters, d_loff);
    Dp2Args da;
    da.p_i = pr->p2_i.as<int32_t>(); da.p_c = pr->p2_c.as<int32_t>(); da.p_rank = pr->p2_rank.as<int32_t>();
    da.p_j = pr->p2_j.as<double>(); da.p_q = pr->p2_q.as<double>();
    da.n_points = (int32_t)n_pts;
    for (int k = 0; k < L2N; ++k) da.level[k] = nodes + loff[k];
    da.cache = pr->cache2.as<Cell2>(); da.cb_val = cb_val; da.cb_id = cb_id;
    da.back_id = back_id; da.len = pr->len2.as<int32_t>(); da.cp = pr->cp2.as<int32_t>();
    da.back_cum = back_cum; da.result = pr->dpres.as<int32_t>();
    dp2_kernel<<<1, 32, 0, st>>>(da);
    Trace2Args ta;
    ta.back_id = back_id; ta.len = da.len; ta.cp = da.cp; ta.result = da.result; ta.back_cum = back_cum;
    ta.p_i = da.p_i; ta.p_c = da.p_c; ta.p_j = da.p_j; ta.p_q = da.p_q;
    ta.seg = pr->seglist.as<int32_t>(); ta.rows = pr->path2.as<double>();
    trace2_kernel<<<1, 256, 0, st>>>(ta);
    ctx->launches += 3;
    DAB_CUDA(cudaMemcpyAsync(&pr->h_counters[12], pr->dpres.as<int32_t>(), 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DAB_CUDA(cudaEventRecord(pr->ev[17], st));
    DAB_CUDA(cudaStreamSynchronize(st));
    n_path = reinterpret_cast<int32_t *>(&pr->h_counters[12])[1];
  } else {
    DAB_CUDA(cudaEventRecord(pr->ev[17], st));
  }
  pr->n_path2 = n_path;
  pr->stats.n_path2 = n_path;
  pr->ev_used[7] = pr->ev_used[8] = true;
  DAB_CUDA(cudaGetLastError());
  return DAB_OK;
}
