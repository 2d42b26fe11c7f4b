// Pass-2 frontier DP (reference describealign.py:946-983) as a block-parallel exact scan.
// Included by stage_b.cu inside its anonymous namespace (uses P2Rec, PmEntry, BackRec, Dp2LArgs,
// line_at, order_bits and the P2_* flags defined there).  tools/dp2_scan_model.py is the executable
// statement of this control flow, checked on the CPU against the one-point rules and the oracle.
//
// Along one corridor the recurrence is  cum_k = max(cum_{k-1}, cum_{k-2}, E_k) + q_k  where E_k is
// whatever enters the chain from outside (frontier jump, cluster best).  While the values of a chain
// stay inside one binade [2^e, 2^(e+1)), an IEEE round-to-nearest add is  fl(x + q) = x + rn_u(q)
// with u = 2^(e-52) (ties aside): exact arithmetic on multiples of u, which float64 itself carries
// without rounding.  Max-plus recurrences over exact numbers are associative, so up to 1024 points -
// all corridors of the block at once, segmented by chain - are ONE scan of 2x2 max-plus matrices
// with an affine term (S1) instead of a serial chain of dependent f64 adds.  Whatever S1 produced is
// then verified: every point recomputes, with the reference's own float64 rules and tie-breaks, its
// cluster best (S2: segmented running maxima), its frontier entry (S3: per-corridor running-max
// rows, committed ones from HBM/L2 and the block's own from shared memory), its predecessor and its
// cum (S4) from the block's values.  If every point reproduces its value bit for bit the block is
// the reference's result by induction over the processing order.  Otherwise, if some point's outside
// input changed (followers of a corridor evaluated in the same block, restarts from a cluster best
// reached inside the block) S1 is repeated with the refreshed inputs; else the verified prefix is
// committed and the failing point (binade crossing, rounding tie) evaluated by the one-point rules,
// as are NEAR / GAP points.  Speculation can only cost time, never exactness.
#ifndef DAB_SC_T
#define DAB_SC_T 512
#endif
#ifndef DAB_SC_K
#define DAB_SC_K 2
#endif
constexpr int SC_T = DAB_SC_T;       // threads per pair
constexpr int SC_K = DAB_SC_K;       // chain slots per thread
constexpr int SC_NB = SC_T * SC_K;   // points per block
constexpr int SC_W = SC_T / 32;
constexpr int SC_MAXP = 3;           // scan passes per block before the verified prefix is committed
constexpr int SC_MINB = 8;           // shorter ranges go through the one-point rules

struct Map2 {                        // x -> A (x) x (+) b over (max, +); x = (cum_k, cum_{k-1})
  double a00, a01, a10, a11, b0, b1;
};

struct ScanShared {
  // per slot; slots are ordered by (corridor, row): every chain is contiguous
  double q[SC_NB], j[SC_NB], cum[SC_NB], ev[SC_NB], clv[SC_NB], pmv[SC_NB];
  int id[SC_NB], kf[SC_NB], ei[SC_NB], cli[SC_NB], pmi[SC_NB];
  // per warp: scan aggregates
  Map2 wmap[SC_W];
  double wclv[SC_W], wpmv[SC_W];
  int wcli[SC_W], wpmi[SC_W], wstart[SC_W];
  int red[SC_W][2];
  int bc[4];
  // per corridor: state after the last committed point
  double c0[32], c1[32], c2[32], clv_c[32], pmv_c[32];
  int id0[32], id1[32], id2[32], cli_c[32], pmi_c[32], filled[32];
  // per corridor: line and extent
  double sl[32], of[32], inv[32];
  int lo[32], rows[32], cluster[32];
  PmEntry *pmbase[32];
  // per corridor, per block
  int cnt[32], base[32];
  unsigned rel[32];
  double hmax[32], big[32];
};

__device__ __forceinline__ double sc_max(double x, double y) { return x > y ? x : y; }

// h = g after f
__device__ __forceinline__ Map2 sc_compose(const Map2 &g, const Map2 &f) {
  Map2 h;
  h.a00 = sc_max(g.a00 + f.a00, g.a01 + f.a10);
  h.a01 = sc_max(g.a00 + f.a01, g.a01 + f.a11);
  h.a10 = sc_max(g.a10 + f.a00, g.a11 + f.a10);
  h.a11 = sc_max(g.a10 + f.a01, g.a11 + f.a11);
  h.b0 = sc_max(sc_max(g.a00 + f.b0, g.a01 + f.b1), g.b0);
  h.b1 = sc_max(sc_max(g.a10 + f.b0, g.a11 + f.b1), g.b1);
  return h;
}

__device__ __forceinline__ Map2 sc_shfl_up(const Map2 &m, int d) {
  const unsigned FULL = 0xffffffffu;
  Map2 r;
  r.a00 = __shfl_up_sync(FULL, m.a00, d); r.a01 = __shfl_up_sync(FULL, m.a01, d);
  r.a10 = __shfl_up_sync(FULL, m.a10, d); r.a11 = __shfl_up_sync(FULL, m.a11, d);
  r.b0 = __shfl_up_sync(FULL, m.b0, d); r.b1 = __shfl_up_sync(FULL, m.b1, d);
  return r;
}

// number of rows r' <= i of corridor c (line sl, of; rows lo .. lo+rows-1) whose coordinate is <= j
__device__ __forceinline__ int sc_rows_le(double sl, double of, double inv, int lo, int rows, int i, double j) {
  double est = floor((j - of) * inv) - (double)lo + 1.0;
  int kk = est < 0.0 ? 0 : (est > (double)rows ? rows : (int)est);
  while (kk < rows && __dadd_rn(__dmul_rn(sl, (double)(lo + kk)), of) <= j) ++kk;
  while (kk > 0 && __dadd_rn(__dmul_rn(sl, (double)(lo + kk - 1)), of) > j) --kk;
  const int done = (i + 1 < lo + rows ? i + 1 : lo + rows) - lo;
  return kk < done ? kk : done;
}

// block-wide minimum of two ints (all threads call; result broadcast)
__device__ __forceinline__ void sc_block_min2(ScanShared &S, int &v0, int &v1) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int m0 = __reduce_min_sync(FULL, v0), m1 = __reduce_min_sync(FULL, v1);
  if (lane == 0) { S.red[w][0] = m0; S.red[w][1] = m1; }
  __syncthreads();
  int r0 = S.red[0][0], r1 = S.red[0][1];
#pragma unroll
  for (int x = 1; x < SC_W; ++x) { r0 = min(r0, S.red[x][0]); r1 = min(r1, S.red[x][1]); }
  __syncthreads();
  v0 = r0; v1 = r1;
}

// One point by the sequential rules, warp 0 only (lane = corridor).  Exact for every kind of point:
// frontier query over the committed running-max rows, cluster best, then either the corridor's own
// previous points (VIS flags) or - NEAR points - the generic prev_cache rule over the last three
// points of every corridor.
__device__ void sc_one_point(ScanShared &S, const Dp2LArgs &a, int p) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  const double NEG = -INFINITY;
  const P2Rec pt = a.rec[p];
  const double j = pt.j, q = pt.q;
  const int kf = pt.kf, k = kf & 0xff, ro = pt.ro, i = pt.i;
  const bool have = lane < a.n_cor;
  // ---- frontier: best entry with j' <= j among the other corridors and the seed ----
  double v = NEG;
  int id = -2;
  if (lane == k) { v = 0.0; id = -1; }
  else if (have && S.rows[lane] > 0 && S.lo[lane] <= i && S.filled[lane] >= 0) {
    const int idx = sc_rows_le(S.sl[lane], S.of[lane], S.inv[lane], S.lo[lane], S.rows[lane], i, j);
    if (idx > 0) {
      const int f = S.filled[lane];
      const int x = idx - 1;
      if (x >= f) { v = S.pmv_c[lane]; id = S.pmi_c[lane]; }
      else {
        const int4 raw = __ldcg(reinterpret_cast<const int4 *>(S.pmbase[lane] + x));
        v = __hiloint2double(raw.y, raw.x); id = raw.z;
      }
    }
  }
  {
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
    v = __shfl_sync(FULL, v, src);
    id = __shfl_sync(FULL, id, src);
  }
  double ub = v;
  int up = id;
  // ---- same-cluster jump ----
  const double clk = S.clv_c[k];
  if (clk >= ub) { ub = clk; up = S.cli_c[k]; }
  // ---- local steps ----
  if (!(kf & P2_NEAR)) {
    if ((kf & P2_VIS2) && S.c1[k] >= ub) { ub = S.c1[k]; up = S.id1[k]; }
    if ((kf & P2_VIS1) && S.c0[k] >= ub) { ub = S.c0[k]; up = S.id0[k]; }
  } else {
    const int cell = pt.cell;
    const int cluster_k = S.cluster[k];
    int hr0 = -100, hr1 = -100, hr2 = -100, hc0 = -100, hc1 = -100, hc2 = -100;
    const int l_id0 = have ? S.id0[lane] : -2, l_id1 = have ? S.id1[lane] : -2, l_id2 = have ? S.id2[lane] : -2;
    if (l_id0 >= 0) { const P2Rec r = a.rec[l_id0]; hr0 = r.i; hc0 = r.cell; }
    if (l_id1 >= 0) { const P2Rec r = a.rec[l_id1]; hr1 = r.i; hc1 = r.cell; }
    if (l_id2 >= 0) { const P2Rec r = a.rec[l_id2]; hr2 = r.i; hc2 = r.cell; }
    const double l_c0 = S.c0[lane], l_c1 = S.c1[lane], l_c2 = S.c2[lane];
    const double l_sl = S.sl[lane], l_of = S.of[lane];
    const int l_cluster = S.cluster[lane];
#pragma unroll 1
    for (int x = cell - 2; x <= cell; ++x) {
      int brow = -1, bh = 0;
      if (hc0 == x && hr0 > brow) { brow = hr0; bh = 0; }
      if (hc1 == x && hr1 > brow) { brow = hr1; bh = 1; }
      if (hc2 == x && hr2 > brow) { brow = hr2; bh = 2; }
      const int mrow = (int)__reduce_max_sync(FULL, (unsigned)(brow + 1)) - 1;
      if (mrow < 0 || mrow < i - 2) continue;         // nothing written recently enough
      const int src = __ffs(__ballot_sync(FULL, brow == mrow)) - 1;
      const double myc = bh == 0 ? l_c0 : (bh == 1 ? l_c1 : l_c2);
      const int myid = bh == 0 ? l_id0 : (bh == 1 ? l_id1 : l_id2);
      double pc = __shfl_sync(FULL, myc, src);
      const int pid = __shfl_sync(FULL, myid, src);
      const double pj = __shfl_sync(FULL, __dadd_rn(__dmul_rn(l_sl, (double)mrow), l_of), src);
      if (__shfl_sync(FULL, l_cluster, src) != cluster_k) {
        const double d = (j - pj) - (double)(i - mrow);
        pc = pc - (100.0 + 100.0 * (d * d));
      }
      if (pj <= j && pc >= ub) { ub = pc; up = pid; }
    }
  }
  __syncwarp();
  // ---- commit (lane k writes) ----
  if (lane == k) {
    const double cum = ub + q;
    S.c2[k] = S.c1[k]; S.id2[k] = S.id1[k];
    S.c1[k] = S.c0[k]; S.id1[k] = S.id0[k];
    S.c0[k] = cum; S.id0[k] = p;
    const double cj = cum - 50.0;
    if (S.clv_c[k] < cj) { S.clv_c[k] = cj; S.cli_c[k] = p; }
    PmEntry e; e.val = S.pmv_c[k]; e.id = S.pmi_c[k]; e.pad = 0;
    // rows without a point (their cell was claimed by an earlier cluster) repeat the head
    for (int r = S.filled[k] + 1; r < ro; ++r) S.pmbase[k][r] = e;
    const double jump = cum - 1000.0;
    if (jump > e.val) { e.val = jump; e.id = p; S.pmv_c[k] = jump; S.pmi_c[k] = p; }
    S.filled[k] = ro;
    S.pmbase[k][ro] = e;
    BackRec b; b.best = ub; b.pred = up; b.pad = 0;
    a.back[p] = b;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(SC_T, 1) dp2_scan_kernel(Dp2LArgs a) {
  extern __shared__ __align__(16) unsigned char sc_smem_raw[];
  ScanShared &S = *reinterpret_cast<ScanShared *>(sc_smem_raw);
  const unsigned FULL = 0xffffffffu;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int n = *a.n_points;
  const int n_cor = a.n_cor;
  const double NEG = -INFINITY;
  const int IMAX = 0x7fffffff;

  if (t < 32) {
    const bool have = t < n_cor;
    dab_corridor c;
    c.cluster = -1; c.lo = 0x7fffffff; c.hi = 0; c.slope = 1.0; c.offset = 0.0;
    if (have) c = a.cor[t];
    S.sl[t] = c.slope; S.of[t] = c.offset; S.inv[t] = 1.0 / c.slope;
    S.lo[t] = c.lo; S.rows[t] = (have && c.hi > c.lo) ? c.hi - c.lo : 0; S.cluster[t] = c.cluster;
    S.pmbase[t] = a.pm + (have ? a.pm_off[t] : 0);
    S.c0[t] = S.c1[t] = S.c2[t] = NEG;
    S.id0[t] = S.id1[t] = S.id2[t] = -2;
    S.clv_c[t] = -1000.0; S.cli_c[t] = -1;       // clusters_best_so_far seed (describealign.py:948)
    S.pmv_c[t] = NEG; S.pmi_c[t] = -2;           // head of the running maximum of cum - 1000
    S.filled[t] = -1;
  }
  __syncthreads();

  unsigned n_blocks = 0, n_passes = 0, n_single = 0, n_blockpts = 0;
  int p0 = 0;
  while (p0 < n) {
    // ---- load up to SC_NB points in processing order; the block ends before the first NEAR / GAP point
    P2Rec rc[SC_K];
    int first_hard = IMAX, dummy = IMAX;
#pragma unroll
    for (int m = 0; m < SC_K; ++m) {
      const int p = p0 + t + SC_T * m;
      rc[m].kf = -1;
      if (p < n) {
        rc[m] = a.rec[p];
        if ((rc[m].kf & (P2_NEAR | P2_GAP)) != 0 && p < first_hard) first_hard = p;
      }
    }
    sc_block_min2(S, first_hard, dummy);
    int p1 = p0 + SC_NB < n ? p0 + SC_NB : n;
    if (first_hard < p1) p1 = first_hard;
    const int nblk = p1 - p0;
    if (nblk < SC_MINB) {
      // a short range (or a NEAR / GAP point right away): one point at a time
      const int stop = nblk == 0 ? p0 + 1 : p1;
      if (w == 0)
        for (int p = p0; p < stop; ++p) sc_one_point(S, a, p);
      n_single += stop - p0;
      p0 = stop;
      __syncthreads();
      continue;
    }
    ++n_blocks;

    // ---- chain slots: slot = base[corridor] + (row - first uncommitted row of the corridor) ----
    if (t < 32) S.cnt[t] = 0;
    __syncthreads();
    int pos[SC_K];
#pragma unroll
    for (int m = 0; m < SC_K; ++m) {
      const int p = p0 + t + SC_T * m;
      const bool in = p < p1;
      const int k = in ? (rc[m].kf & 0xff) : (32 + lane);
      pos[m] = in ? rc[m].ro - (S.filled[k & 31] + 1) : -1;
      const unsigned grp = __match_any_sync(FULL, k);
      const int mx = __reduce_max_sync(grp, pos[m]);           // rows grow with the id inside a corridor
      if (in && pos[m] == mx) atomicMax(&S.cnt[k], mx + 1);
    }
    __syncthreads();
    if (w == 0) {
      const int c = S.cnt[lane];
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
      }
      S.base[lane] = inc - c;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < SC_K; ++m) {
      const int p = p0 + t + SC_T * m;
      if (p < p1) {
        const int k = rc[m].kf & 0xff;
        const int s = S.base[k] + pos[m];
        S.q[s] = rc[m].q; S.j[s] = rc[m].j; S.id[s] = p; S.kf[s] = rc[m].kf;
      }
    }
    // per corridor: rounding unit of the chain (S1) and the corridors whose committed running-max
    // head could beat the chain's cluster best (the only entries a frontier look-up can return)
    if (w == 0) {
      const double clk = S.clv_c[lane];
      unsigned rel = 0u;
      for (int c = 0; c < n_cor; ++c) {
        const double h = S.pmv_c[c];          // uniform
        if (c != lane && S.filled[c] >= 0 && h > clk) rel |= 1u << c;
      }
      S.rel[lane] = rel;
    }
    __syncthreads();

    // the frontier entry (value > floor_v only; ties: smaller j', then smaller id) seen by a slot
    auto frontier = [&](int s, int k, int i, double j, double floor_v, bool inblock, double &fv, int &fi) {
      fv = NEG; fi = -2;
      double fj = 0.0;
      if (0.0 > floor_v) { fv = 0.0; fi = -1; }               // the seed (0, 0, -1, 0, 0)
      unsigned mask = S.rel[k];
      while (mask) {
        const int c = __ffs(mask) - 1;
        mask &= mask - 1;
        if (S.lo[c] > i) continue;
        const int idx = sc_rows_le(S.sl[c], S.of[c], S.inv[c], S.lo[c], S.rows[c], i, j);
        if (idx <= 0) continue;
        const int x = idx - 1;
        const int f0 = S.filled[c];
        double v;
        int id;
        if (inblock && S.cnt[c] > 0 && x > f0) {
          const int o = x - f0 - 1 < S.cnt[c] - 1 ? x - f0 - 1 : S.cnt[c] - 1;
          v = S.pmv[S.base[c] + o]; id = S.pmi[S.base[c] + o];
        } else {
          if (f0 < 0) continue;
          if (x >= f0) { v = S.pmv_c[c]; id = S.pmi_c[c]; }
          else {
            const int4 raw = __ldcg(reinterpret_cast<const int4 *>(S.pmbase[c] + x));
            v = __hiloint2double(raw.y, raw.x); id = raw.z;
          }
        }
        if (!(v > floor_v)) continue;
        if (v > fv) { fv = v; fi = id; fj = id < 0 ? 0.0 : a.p_j[id]; }
        else if (v == fv) {
          const double vj = id < 0 ? 0.0 : a.p_j[id];
          if (vj < fj || (vj == fj && id < fi)) { fi = id; fj = vj; }
        }
      }
      (void)s;
    };

    // ---- initial outside inputs: committed state only ----
#pragma unroll
    for (int u = 0; u < SC_K; ++u) {
      const int s = t * SC_K + u;
      if (s < nblk) {
        const int k = S.kf[s] & 0xff;
        const int i = S.lo[k] + S.filled[k] + 1 + (s - S.base[k]);
        const double clk = S.clv_c[k];
        double fv; int fi;
        frontier(s, k, i, S.j[s], clk, false, fv, fi);
        if (fi == -2) { fv = clk; fi = S.cli_c[k]; }      // cl >= F: the cluster best stands
        S.ev[s] = fv; S.ei[s] = fi;
      }
    }
    __syncthreads();
    if (t < 32) {
      // rounding unit per chain: the binade of the corridor's last cum, or of what its first point starts from
      double ref = S.c0[t];
      if (S.cnt[t] > 0) {
        if (!(ref > NEG)) ref = S.ev[S.base[t]];
        const unsigned long long bits = (unsigned long long)__double_as_longlong(ref);
        const unsigned long long ex = (bits >> 52) & 0x7ffull;
        S.big[t] = __longlong_as_double((long long)((ex << 52) | 0x0008000000000000ull));   // 1.5 * 2^e
      }
    }
    __syncthreads();

    int good = p1;
    double r_best[SC_K];
    int r_pred[SC_K];
    for (int pass = 0; pass < SC_MAXP; ++pass) {
      ++n_passes;
      // ================= S1: max-plus scan over all chains of the block =================
      Map2 agg;
      agg.a00 = 0.0; agg.a01 = NEG; agg.a10 = NEG; agg.a11 = 0.0; agg.b0 = NEG; agg.b1 = NEG;
      double e_q[SC_K], e_e[SC_K], e_c0[SC_K];
      int e_fl[SC_K];          // bit0 v1, bit1 v2, bit2 first of chain, bit3 valid
#pragma unroll
      for (int u = 0; u < SC_K; ++u) {
        const int s = t * SC_K + u;
        e_fl[u] = 0; e_q[u] = 0.0; e_e[u] = NEG; e_c0[u] = NEG;
        if (s < nblk) {
          const int kf = S.kf[s], k = kf & 0xff;
          const double big = S.big[k];
          const double qr = __dadd_rn(__dadd_rn(S.q[s], big), -big);     // q rounded to the chain's unit
          const bool v1 = (kf & P2_VIS1) != 0, v2 = (kf & P2_VIS2) != 0;
          const bool first = s == S.base[k];
          const double E = S.ev[s];
          e_fl[u] = (v1 ? 1 : 0) | (v2 ? 2 : 0) | (first ? 4 : 0) | 8;
          e_q[u] = qr; e_e[u] = E;
          if (first) {
            const double pc0 = S.c0[k], pc1 = S.c1[k];
            const double m = sc_max(sc_max(v1 ? pc0 : NEG, v2 ? pc1 : NEG), E);
            e_c0[u] = pc0;
            agg.a00 = NEG; agg.a01 = NEG; agg.a10 = NEG; agg.a11 = NEG;
            agg.b0 = m + qr; agg.b1 = pc0;
            e_e[u] = m;           // for a chain's first point: the value it starts from
          } else {
            const double n00 = sc_max(v1 ? agg.a00 : NEG, v2 ? agg.a10 : NEG) + qr;
            const double n01 = sc_max(v1 ? agg.a01 : NEG, v2 ? agg.a11 : NEG) + qr;
            const double nb0 = sc_max(sc_max(v1 ? agg.b0 : NEG, v2 ? agg.b1 : NEG), E) + qr;
            agg.a10 = agg.a00; agg.a11 = agg.a01; agg.b1 = agg.b0;
            agg.a00 = n00; agg.a01 = n01; agg.b0 = nb0;
          }
        }
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const Map2 L = sc_shfl_up(agg, d);
        if (lane >= d) agg = sc_compose(agg, L);
      }
      if (lane == 31) S.wmap[w] = agg;
      __syncthreads();
      double x0 = NEG, x1 = NEG;          // state entering this thread's slots
      for (int x = 0; x < w; ++x) {
        const Map2 m = S.wmap[x];
        const double y0 = sc_max(sc_max(m.a00 + x0, m.a01 + x1), m.b0);
        const double y1 = sc_max(sc_max(m.a10 + x0, m.a11 + x1), m.b1);
        x0 = y0; x1 = y1;
      }
      {
        const Map2 ex = sc_shfl_up(agg, 1);
        if (lane > 0) {
          const double y0 = sc_max(sc_max(ex.a00 + x0, ex.a01 + x1), ex.b0);
          const double y1 = sc_max(sc_max(ex.a10 + x0, ex.a11 + x1), ex.b1);
          x0 = y0; x1 = y1;
        }
      }
#pragma unroll
      for (int u = 0; u < SC_K; ++u) {
        const int s = t * SC_K + u;
        if (e_fl[u] & 8) {
          double c;
          if (e_fl[u] & 4) { c = e_e[u] + e_q[u]; x1 = e_c0[u]; }
          else {
            const double m = sc_max(sc_max((e_fl[u] & 1) ? x0 : NEG, (e_fl[u] & 2) ? x1 : NEG), e_e[u]);
            c = m + e_q[u];
            x1 = x0;
          }
          x0 = c;
          S.cum[s] = c;
        }
      }
      __syncthreads();

      // ================= S2: cluster best and running-max head after every point =================
      // segmented inclusive scans of (cum - 50, id) and (cum - 1000, id), first occurrence on ties,
      // seeded with the corridor's committed cluster best / head
      double a_clv = NEG, a_pmv = NEG;
      int a_cli = -2, a_pmi = -2, a_start = 0;
      double l_clv[SC_K], l_pmv[SC_K];
      int l_cli[SC_K], l_pmi[SC_K];
#pragma unroll
      for (int u = 0; u < SC_K; ++u) {
        const int s = t * SC_K + u;
        if (s < nblk) {
          const int k = S.kf[s] & 0xff;
          if (s == S.base[k]) { a_clv = S.clv_c[k]; a_cli = S.cli_c[k]; a_pmv = S.pmv_c[k]; a_pmi = S.pmi_c[k]; a_start = 1; }
          const double cum = S.cum[s];
          const double cj = cum - 50.0, jp = cum - 1000.0;
          if (a_clv < cj) { a_clv = cj; a_cli = S.id[s]; }
          if (jp > a_pmv) { a_pmv = jp; a_pmi = S.id[s]; }
        }
        l_clv[u] = a_clv; l_cli[u] = a_cli; l_pmv[u] = a_pmv; l_pmi[u] = a_pmi;
      }
      // warp scan of the thread totals (right operand wins when strictly greater, or when it starts a chain)
      double s_clv = a_clv, s_pmv = a_pmv;
      int s_cli = a_cli, s_pmi = a_pmi, s_start = a_start;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o_clv = __shfl_up_sync(FULL, s_clv, d), o_pmv = __shfl_up_sync(FULL, s_pmv, d);
        const int o_cli = __shfl_up_sync(FULL, s_cli, d), o_pmi = __shfl_up_sync(FULL, s_pmi, d);
        const int o_start = __shfl_up_sync(FULL, s_start, d);
        if (lane >= d && !s_start) {
          if (!(s_clv > o_clv)) { s_clv = o_clv; s_cli = o_cli; }
          if (!(s_pmv > o_pmv)) { s_pmv = o_pmv; s_pmi = o_pmi; }
          s_start = o_start;
        }
      }
      if (lane == 31) { S.wclv[w] = s_clv; S.wcli[w] = s_cli; S.wpmv[w] = s_pmv; S.wpmi[w] = s_pmi; S.wstart[w] = s_start; }
      __syncthreads();
      {
        // carry entering this thread: warps before, then lanes before (stops at the nearest chain start)
        double c_clv = NEG, c_pmv = NEG;
        int c_cli = -2, c_pmi = -2;
        for (int x = 0; x < w; ++x) {
          if (S.wstart[x]) { c_clv = S.wclv[x]; c_cli = S.wcli[x]; c_pmv = S.wpmv[x]; c_pmi = S.wpmi[x]; }
          else {
            if (S.wclv[x] > c_clv) { c_clv = S.wclv[x]; c_cli = S.wcli[x]; }
            if (S.wpmv[x] > c_pmv) { c_pmv = S.wpmv[x]; c_pmi = S.wpmi[x]; }
          }
        }
        double p_clv = __shfl_up_sync(FULL, s_clv, 1), p_pmv = __shfl_up_sync(FULL, s_pmv, 1);
        int p_cli = __shfl_up_sync(FULL, s_cli, 1), p_pmi = __shfl_up_sync(FULL, s_pmi, 1);
        const int p_start = __shfl_up_sync(FULL, s_start, 1);
        if (lane > 0) {
          if (p_start) { c_clv = p_clv; c_cli = p_cli; c_pmv = p_pmv; c_pmi = p_pmi; }
          else {
            if (p_clv > c_clv) { c_clv = p_clv; c_cli = p_cli; }
            if (p_pmv > c_pmv) { c_pmv = p_pmv; c_pmi = p_pmi; }
          }
        }
        // apply to the local running values up to (not including) this thread's first chain start
        bool started = false;
#pragma unroll
        for (int u = 0; u < SC_K; ++u) {
          const int s = t * SC_K + u;
          if (s < nblk) {
            if (s == S.base[S.kf[s] & 0xff]) started = true;
            double v1 = l_clv[u], v2 = l_pmv[u];
            int i1 = l_cli[u], i2 = l_pmi[u];
            if (!started) {
              if (!(v1 > c_clv)) { v1 = c_clv; i1 = c_cli; }
              if (!(v2 > c_pmv)) { v2 = c_pmv; i2 = c_pmi; }
            }
            S.clv[s] = v1; S.cli[s] = i1; S.pmv[s] = v2; S.pmi[s] = i2;
          }
        }
      }
      __syncthreads();
      // corridors whose running-max head (the block's entries included) could beat a chain's cluster best
      if (w == 0) {
        const bool present = S.cnt[lane] > 0;
        const double h = present ? S.pmv[S.base[lane] + S.cnt[lane] - 1] : S.pmv_c[lane];
        S.hmax[lane] = (lane < n_cor && (present || S.filled[lane] >= 0)) ? h : NEG;
        __syncwarp();
        const double clk = S.clv_c[lane];
        unsigned rel = 0u;
        for (int c = 0; c < n_cor; ++c)
          if (c != lane && S.hmax[c] > clk) rel |= 1u << c;
        S.rel[lane] = rel;
      }
      __syncthreads();

      // ================= S3 + S4: every point by the reference's float64 rules =================
      int first_bad = IMAX, first_chg = IMAX;
      double n_ev[SC_K];
      int n_ei[SC_K];
#pragma unroll
      for (int u = 0; u < SC_K; ++u) {
        const int s = t * SC_K + u;
        n_ev[u] = NEG; n_ei[u] = -2;
        if (s < nblk) {
          const int kf = S.kf[s], k = kf & 0xff;
          const int po = s - S.base[k];
          const int i = S.lo[k] + S.filled[k] + 1 + po;
          // cluster best before this point
          double clb = S.clv_c[k];
          int clbi = S.cli_c[k];
          if (po > 0) { clb = S.clv[s - 1]; clbi = S.cli[s - 1]; }
          double fv; int fi;
          frontier(s, k, i, S.j[s], clb, true, fv, fi);
          if (fi == -2) { fv = clb; fi = clbi; }
          n_ev[u] = fv; n_ei[u] = fi;
          double best = fv;
          int pred = fi;
          const double pc0 = po >= 1 ? S.cum[s - 1] : S.c0[k];
          const int pi0 = po >= 1 ? S.id[s - 1] : S.id0[k];
          const double pc1 = po >= 2 ? S.cum[s - 2] : (po == 1 ? S.c0[k] : S.c1[k]);
          const int pi1 = po >= 2 ? S.id[s - 2] : (po == 1 ? S.id0[k] : S.id1[k]);
          if ((kf & P2_VIS2) && pc1 >= best) { best = pc1; pred = pi1; }
          if ((kf & P2_VIS1) && pc0 >= best) { best = pc0; pred = pi0; }
          r_best[u] = best; r_pred[u] = pred;
          const double chk = best + S.q[s];
          if (__double_as_longlong(chk) != __double_as_longlong(S.cum[s])) first_bad = min(first_bad, S.id[s]);
          if (__double_as_longlong(fv) != __double_as_longlong(S.ev[s])) first_chg = min(first_chg, S.id[s]);
        }
      }
      sc_block_min2(S, first_bad, first_chg);
      if (first_bad == IMAX) { good = p1; break; }
      good = first_bad;
      if (pass + 1 == SC_MAXP || first_chg > first_bad) break;
      // refresh the outside inputs from this pass and scan again
#pragma unroll
      for (int u = 0; u < SC_K; ++u) {
        const int s = t * SC_K + u;
        if (s < nblk) { S.ev[s] = n_ev[u]; S.ei[s] = n_ei[u]; }
      }
      __syncthreads();
    }

    // ================= commit the verified prefix (ids < good) =================
#pragma unroll
    for (int u = 0; u < SC_K; ++u) {
      const int s = t * SC_K + u;
      if (s < nblk && S.id[s] < good) {
        const int k = S.kf[s] & 0xff;
        const int po = s - S.base[k];
        const int ro = S.filled[k] + 1 + po;
        BackRec b; b.best = r_best[u]; b.pred = r_pred[u]; b.pad = 0;
        a.back[S.id[s]] = b;
        PmEntry e; e.val = S.pmv[s]; e.id = S.pmi[s]; e.pad = 0;
        S.pmbase[k][ro] = e;
      }
    }
    __syncthreads();
    // new corridor state: the last committed slot of every chain
    {
      double n_c0[SC_K], n_c1[SC_K], n_c2[SC_K];
      int n_i0[SC_K], n_i1[SC_K], n_i2[SC_K], n_k[SC_K], n_ro[SC_K];
#pragma unroll
      for (int u = 0; u < SC_K; ++u) {
        const int s = t * SC_K + u;
        n_k[u] = -1;
        if (s < nblk && S.id[s] < good) {
          const int k = S.kf[s] & 0xff;
          const int po = s - S.base[k];
          const bool last = po == S.cnt[k] - 1 || S.id[s + 1] >= good;
          if (last) {
            n_k[u] = k; n_ro[u] = S.filled[k] + 1 + po;
            n_c0[u] = S.cum[s]; n_i0[u] = S.id[s];
            n_c1[u] = po >= 1 ? S.cum[s - 1] : S.c0[k]; n_i1[u] = po >= 1 ? S.id[s - 1] : S.id0[k];
            n_c2[u] = po >= 2 ? S.cum[s - 2] : (po == 1 ? S.c0[k] : S.c1[k]);
            n_i2[u] = po >= 2 ? S.id[s - 2] : (po == 1 ? S.id0[k] : S.id1[k]);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < SC_K; ++u) {
        if (n_k[u] >= 0) {
          const int s = t * SC_K + u, k = n_k[u];
          S.c0[k] = n_c0[u]; S.c1[k] = n_c1[u]; S.c2[k] = n_c2[u];
          S.id0[k] = n_i0[u]; S.id1[k] = n_i1[u]; S.id2[k] = n_i2[u];
          S.clv_c[k] = S.clv[s]; S.cli_c[k] = S.cli[s];
          S.pmv_c[k] = S.pmv[s]; S.pmi_c[k] = S.pmi[s];
          S.filled[k] = n_ro[u];
        }
      }
    }
    n_blockpts += good - p0;
    __syncthreads();
    p0 = good;
    if (good < p1) {
      if (w == 0) sc_one_point(S, a, p0);
      ++n_single;
      ++p0;
      __syncthreads();
    }
  }

  // ---- the frontier's best entry: seed or a corridor's head (value desc, j' asc, id asc) ----
  if (w == 0) {
    double bv = NEG, bj = INFINITY;
    int bi = -2;
    if (lane < n_cor && S.pmi_c[lane] >= 0) { bv = S.pmv_c[lane]; bi = S.pmi_c[lane]; bj = a.p_j[bi]; }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const double ov = __shfl_xor_sync(FULL, bv, o), oj = __shfl_xor_sync(FULL, bj, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      const bool take = ov > bv || (ov == bv && (oj < bj || (oj == bj && oi < bi)));
      if (take) { bv = ov; bj = oj; bi = oi; }
    }
    // the seed entry (value 0 at j' = 0, id -1) is part of the frontier
    if (bi == -2 || 0.0 > bv || (0.0 == bv && 0.0 <= bj)) { bv = 0.0; bj = 0.0; bi = -1; }
    if (lane == 0) {
      a.result[0] = bi;
      *reinterpret_cast<double *>(a.result + 2) = bv;
      a.counters[0] = n_passes; a.counters[1] = n_blocks; a.counters[2] = n_single; a.counters[3] = n_blockpts;
    }
  }
}
