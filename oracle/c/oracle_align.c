/*
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU restatement of the loops of the
 * reference's align() (describealign.py:595-1027) that are too slow as pure Python.
 * Used as the parity checker and as the "port" CPU baseline; never linked into the product.
 *
 *   oracle_ddot            OpenBLAS 0.3.30 SkylakeX ddot summation order (numpy's np.dot /
 *                          np.convolve on f64; SURVEY.md B.2 iv, re-derived empirically)
 *   oracle_meansub_norm    describealign.py:599-608
 *   oracle_codes           describealign.py:622-628, 638-644
 *   oracle_match           describealign.py:615-633 (tables), 649-673 (lookup + scoring)
 *   oracle_dp1             describealign.py:654-656, 674-700 (frontier DP + traceback)
 *   oracle_dp2             describealign.py:946-989 (second frontier DP + traceback)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- OpenBLAS ddot, SkylakeX kernel order ------------------------------------------ */
double oracle_ddot(const double *x, const double *y, int64_t n) {
  double a[4][4] = {{0}};
  int64_t n1 = n & ~(int64_t)15;
  int64_t n32 = n1 & ~(int64_t)31;
  int64_t i = 0;
  if (n32) {
    double z[4][8] = {{0}};
    for (; i < n32; i += 32)
      for (int k = 0; k < 4; ++k)
        for (int l = 0; l < 8; ++l) z[k][l] = fma(x[i + 8 * k + l], y[i + 8 * k + l], z[k][l]);
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 4; ++l) a[k][l] = z[k][l] + z[k][l + 4];
  }
  for (; i < n1; i += 16)
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 4; ++l) a[k][l] = fma(x[i + 4 * k + l], y[i + 4 * k + l], a[k][l]);
  double s[4];
  for (int l = 0; l < 4; ++l) s[l] = ((a[0][l] + a[1][l]) + a[2][l]) + a[3][l];
  double dot = (s[0] + s[2]) + (s[1] + s[3]);
  for (; i < n; ++i) dot = fma(y[i], x[i], dot);
  return dot;
}

/* ---- mean subtraction + sliding norm (describealign.py:599-608) ----------------------- */
/* feat: n values (already f64).  ms: n values.  nrm: n-40 values. */
void oracle_meansub_norm(const double *feat, int64_t n, const double *hann41, double *ms, double *nrm) {
  /* np.convolve flips the window; scipy's Hann window is not bit-symmetric */
  double hflip[41];
  for (int k = 0; k < 41; ++k) hflip[k] = hann41[40 - k];
  for (int64_t t = 0; t < n; ++t) {
    int64_t lo = t - 20, hi = t + 21; /* window [lo, hi) */
    int64_t klo = 0;
    if (lo < 0) { klo = -lo; lo = 0; }
    if (hi > n) hi = n;
    double mean = oracle_ddot(feat + lo, hflip + klo, hi - lo);
    ms[t] = feat[t] - mean;
  }
  if (n < 41) return;
  double *sq = (double *)malloc(sizeof(double) * (size_t)n);
  double ones[41];
  for (int i = 0; i < 41; ++i) ones[i] = 1.0;
  for (int64_t t = 0; t < n; ++t) sq[t] = ms[t] * ms[t];
  for (int64_t t = 0; t + 41 <= n; ++t) {
    double v = sqrt(oracle_ddot(sq + t, ones, 41));
    nrm[t] = v < 0.001 ? 0.001 : v;
  }
  free(sq);
}

/* ---- digit codes (describealign.py:622-628 video, 638-644 audio) ---------------------- */
/* code[t] = sum digit_k 7^k ; flags[t] bit k = frac(d_k) > .6 (video only). */
void oracle_codes(const double *ms, const double *nrm, int64_t n, int is_video, int32_t *code, uint8_t *flags) {
  for (int64_t t = 0; t + 41 <= n; ++t) {
    int32_t c = 0, p7 = 1;
    uint8_t fl = 0;
    for (int k = 0; k < 7; ++k) {
      double d = ms[t + 2 + 6 * k] / nrm[t];
      int dig;
      if (is_video) {
        d = 8.0 * d;
        d = d + 3.3;
        if (d < 0.0) d = 0.0;
        if (d > 6.0) d = 6.0;
        double fl_d = floor(d);
        if (d - fl_d > 0.6) fl |= (uint8_t)(1u << k);
        dig = (int)fl_d;
      } else {
        d = 8.0 * d;
        d = d + 3.5;
        double fl_d = floor(d);
        /* astype(int) then clip: values are far inside the int64 range */
        if (fl_d < 0.0) fl_d = 0.0;
        if (fl_d > 6.0) fl_d = 6.0;
        dig = (int)fl_d;
      }
      c += dig * p7;
      p7 *= 7;
    }
    code[t] = c;
    if (flags) flags[t] = fl;
  }
}

/* ---- hash tables + candidate lookup + scoring ------------------------------------------ */
#define NCODE 823543 /* 7^7 */

typedef struct {
  int32_t *start; /* NCODE + 1 */
  int32_t *items;
} table_t;

static void build_table(table_t *tb, const int32_t *code, const uint8_t *flags, const int32_t *sel, int64_t nsel) {
  static const int32_t P7[7] = {1, 7, 49, 343, 2401, 16807, 117649};
  tb->start = (int32_t *)calloc(NCODE + 2, sizeof(int32_t));
  int64_t total = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int64_t s = 0; s < nsel; ++s) {
      int32_t v = sel[s];
      uint8_t fl = flags[v];
      /* all subsets of the flagged digits (describealign.py:616-620, 632-633) */
      for (uint32_t sub = fl;; sub = (sub - 1) & fl) {
        int32_t c = code[v];
        for (int k = 0; k < 7; ++k)
          if (sub & (1u << k)) c += P7[k];
        if (pass == 0) {
          tb->start[c + 1]++;
          total++;
        } else {
          tb->items[tb->start[c]++] = v;
        }
        if (sub == 0) break;
      }
    }
    if (pass == 0) {
      for (int64_t c = 0; c < NCODE; ++c) tb->start[c + 1] += tb->start[c];
      tb->items = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
    } else {
      /* start[c] now points at the end of bucket c; shift back */
      for (int64_t c = NCODE; c > 0; --c) tb->start[c] = tb->start[c - 1];
      tb->start[0] = 0;
    }
  }
}

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

/* Returns the number of match points; fills (pi, pv, pq) up to cap entries, sorted by (i, v).
 * If the return value exceeds cap the caller retries with a larger buffer.
 * stats[0] = set elements touched, stats[1] = candidates scored. */
int64_t oracle_match(const double *const *a_ms, const double *const *a_nrm, const int32_t *const *a_code,
                     const int32_t *a_nq, int64_t n_anq,
                     const double *const *v_ms, const double *const *v_nrm, const int32_t *const *v_code,
                     const uint8_t *const *v_flags, const int32_t *v_sel, int64_t n_vsel, int64_t Lv,
                     int32_t *pi, int32_t *pv, double *pq, int64_t cap, int64_t *stats) {
  table_t tb[5];
  for (int f = 0; f < 5; ++f) build_table(&tb[f], v_code[f], v_flags[f], v_sel, n_vsel);
  uint8_t *mark = (uint8_t *)calloc((size_t)(Lv > 0 ? Lv : 1), 1);
  int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Lv > 0 ? Lv : 1));
  int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Lv > 0 ? Lv : 1));
  int64_t np = 0, touched_total = 0, scored = 0;
  for (int64_t q = 0; q < n_anq; ++q) {
    int32_t i = a_nq[q];
    int64_t nt = 0;
    for (int f = 0; f < 5; ++f) {
      int32_t c = a_code[f][i];
      for (int32_t e = tb[f].start[c]; e < tb[f].start[c + 1]; ++e) {
        int32_t v = tb[f].items[e];
        if (mark[v] == 0) touched[nt++] = v;
        /* low 2 bits: count over tables 0-2 ; bit 2: seen in table 3 or 4 */
        if (f < 3) mark[v] = (uint8_t)(mark[v] + 1) | 0x80; else mark[v] |= 0x84;
        touched_total++;
      }
    }
    int64_t nc = 0;
    for (int64_t k = 0; k < nt; ++k) {
      int32_t v = touched[k];
      uint8_t m = mark[v];
      mark[v] = 0;
      if ((m & 3) >= 2 && (m & 4)) cand[nc++] = v;
    }
    qsort(cand, (size_t)nc, sizeof(int32_t), cmp_i32);
    for (int64_t k = 0; k < nc; ++k) {
      int32_t v = cand[k];
      scored++;
      double prob = 1.0;
      for (int j = 0; j < 3; ++j) {
        double corr = oracle_ddot(a_ms[j] + i, v_ms[j] + v, 41);
        corr = corr / (a_nrm[j][i] * v_nrm[j][v]);
        double om = 1.0 - corr;
        prob = prob * (om > 1e-8 ? om : 1e-8);
      }
      prob = pow(prob, 2.9);
      if (prob > 1e-8) continue;
      double qual = pow(prob / 1e-12, -1.0 / 3);
      if (qual > 50.0) qual = 50.0;
      if (np < cap) { pi[np] = i; pv[np] = v; pq[np] = qual; }
      np++;
    }
  }
  if (stats) { stats[0] = touched_total; stats[1] = scored; }
  for (int f = 0; f < 5; ++f) { free(tb[f].start); free(tb[f].items); }
  free(mark); free(touched); free(cand);
  return np;
}

/* ---- frontier DP #1 as a prefix-max --------------------------------------
 * cum(p) = q(p) + max{cum(p') : p' processed before p, v' <= v} (0 from the seed if none is
 * positive... the seed has cum 0 and v = -1, so it is always eligible); ties on cum resolve
 * to the smallest v'.  Points arrive sorted by (i, v).  Segment tree over v. */
typedef struct { double cum; int32_t v; int32_t id; } node1_t;

static inline int better1(const node1_t *a, const node1_t *b) {
  /* is a preferable to b as predecessor? larger cum, then smaller v */
  if (a->cum != b->cum) return a->cum > b->cum;
  return a->v < b->v;
}

/* Returns path length; path_i/path_v filled in ascending order (capacity np).
 * cum_out (np) and back_out (np, -1 = seed) are optional debugging outputs. */
int64_t oracle_dp1(const int32_t *pi, const int32_t *pv, const double *pq, int64_t np, int64_t Lv,
                   int32_t *path_i, int32_t *path_v, double *cum_out, int32_t *back_out) {
  (void)pi;
  int64_t size = 1;
  while (size < Lv + 1) size <<= 1;
  node1_t *tree = (node1_t *)malloc(sizeof(node1_t) * (size_t)(2 * size));
  for (int64_t k = 0; k < 2 * size; ++k) { tree[k].cum = -INFINITY; tree[k].v = INT32_MAX; tree[k].id = -1; }
  int32_t *back = (int32_t *)malloc(sizeof(int32_t) * (size_t)(np > 0 ? np : 1));
  node1_t best_all = {0.0, -1, -1}; /* the seed (-1, -1, 0) */
  for (int64_t p = 0; p < np; ++p) {
    int64_t v = pv[p];
    node1_t best = {0.0, -1, -1};
    /* prefix query [0, v] */
    int64_t lo = size, hi = size + v + 1; /* [lo, hi) */
    while (lo < hi) {
      if (lo & 1) { if (better1(&tree[lo], &best)) best = tree[lo]; lo++; }
      if (hi & 1) { --hi; if (better1(&tree[hi], &best)) best = tree[hi]; }
      lo >>= 1; hi >>= 1;
    }
    double cum = best.cum + pq[p];
    back[p] = best.id;
    if (cum_out) cum_out[p] = cum;
    node1_t me = {cum, (int32_t)v, (int32_t)p};
    int64_t k = size + v;
    /* a later point at the same v always has a larger cum than the earlier one */
    tree[k] = me;
    for (k >>= 1; k >= 1; k >>= 1) {
      const node1_t *l = &tree[2 * k], *r = &tree[2 * k + 1];
      tree[k] = better1(r, l) ? *r : *l;
    }
    if (better1(&me, &best_all)) best_all = me;
  }
  if (back_out) memcpy(back_out, back, sizeof(int32_t) * (size_t)np);
  int64_t len = 0;
  for (int32_t p = best_all.id; p >= 0; p = back[p]) len++;
  int64_t w = len;
  for (int32_t p = best_all.id; p >= 0; p = back[p]) { --w; path_i[w] = pi[p]; path_v[w] = pv[p]; }
  free(tree); free(back);
  return len;
}

/* ---- frontier DP #2 (describealign.py:946-989) ---------------------------
 * Points sorted by (i, j, cluster).  j_rank: rank of j among all distinct j values, >= 1
 * (rank 0 is reserved for the seed at j = 0).  Output rows (j, i, cluster, qual, cum) in
 * ascending order exactly as the reference's `path` before the /210 scaling. */
typedef struct { double val; double j; int32_t order; int32_t id; } node2_t; /* id -1 = seed */

static inline int better2(const node2_t *a, const node2_t *b) {
  if (a->val != b->val) return a->val > b->val;
  if (a->j != b->j) return a->j < b->j;
  return a->order < b->order;
}

typedef struct { double j, i, c, q, cum; } tup5_t;

int64_t oracle_dp2(const int32_t *pi, const double *pj, const int32_t *pc, const double *pq,
                   const int32_t *j_rank, int64_t np, int64_t n_rank, int64_t n_clusters, int64_t Lv,
                   double *path_out /* np x 5 */) {
  int64_t size = 1;
  while (size < n_rank + 1) size <<= 1;
  node2_t *tree = (node2_t *)malloc(sizeof(node2_t) * (size_t)(2 * size));
  for (int64_t k = 0; k < 2 * size; ++k) { tree[k].val = -INFINITY; tree[k].j = INFINITY; tree[k].order = INT32_MAX; tree[k].id = -2; }
  /* seed (0, 0, -1, 0, 0) at rank 0 */
  {
    node2_t seed = {0.0, 0.0, -1, -1};
    int64_t k = size;
    tree[k] = seed;
    for (k >>= 1; k >= 1; k >>= 1) tree[k] = better2(&tree[2 * k + 1], &tree[2 * k]) ? tree[2 * k + 1] : tree[2 * k];
  }
  tup5_t *cache = (tup5_t *)malloc(sizeof(tup5_t) * (size_t)(Lv > 0 ? Lv : 1));
  for (int64_t k = 0; k < Lv; ++k) { cache[k].j = cache[k].i = cache[k].c = cache[k].q = cache[k].cum = -INFINITY; }
  if (Lv > 0) { cache[0].j = 0; cache[0].i = 0; cache[0].c = -1; cache[0].q = 0; cache[0].cum = 0; }
  tup5_t *cbest = (tup5_t *)malloc(sizeof(tup5_t) * (size_t)(n_clusters > 0 ? n_clusters : 1));
  for (int64_t c = 0; c < n_clusters; ++c) { cbest[c].j = 0; cbest[c].i = 0; cbest[c].c = (double)c; cbest[c].q = 0; cbest[c].cum = -1000.0; }
  tup5_t *back = (tup5_t *)malloc(sizeof(tup5_t) * (size_t)(np > 0 ? np : 1));
  /* map from (pred j, pred i) to a point id is needed for the traceback: remember ids */
  int32_t *back_id = (int32_t *)malloc(sizeof(int32_t) * (size_t)(np > 0 ? np : 1));
  int32_t *cache_id = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Lv > 0 ? Lv : 1));
  int32_t *cbest_id = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_clusters > 0 ? n_clusters : 1));
  for (int64_t k = 0; k < Lv; ++k) cache_id[k] = -1;
  for (int64_t c = 0; c < n_clusters; ++c) cbest_id[c] = -1;

  for (int64_t p = 0; p < np; ++p) {
    double j = pj[p], q = pq[p];
    int32_t i = pi[p], c = pc[p];
    /* (1) global frontier: best val among processed points with j' <= j */
    node2_t fb = {-INFINITY, INFINITY, INT32_MAX, -2};
    {
      int64_t lo = size, hi = size + j_rank[p] + 1;
      while (lo < hi) {
        if (lo & 1) { if (better2(&tree[lo], &fb)) fb = tree[lo]; lo++; }
        if (hi & 1) { --hi; if (better2(&tree[hi], &fb)) fb = tree[hi]; }
        lo >>= 1; hi >>= 1;
      }
    }
    tup5_t prev;
    int32_t prev_id = fb.id;
    if (fb.id < 0) { prev.j = 0; prev.i = 0; prev.c = -1; prev.q = 0; prev.cum = 0; }
    else { prev.j = pj[fb.id]; prev.i = pi[fb.id]; prev.c = pc[fb.id]; prev.q = pq[fb.id]; prev.cum = fb.val; }
    double frontier_val = fb.val;
    /* (2) same-cluster jump */
    tup5_t cl = cbest[c];
    if (cl.cum >= prev.cum) { prev = cl; prev.c = (double)c; prev_id = cbest_id[c]; }
    /* (3) local steps through prev_cache */
    int64_t ij = (int64_t)j; /* int(j), j >= 0 */
    for (int64_t pjx = (ij - 2 > 0 ? ij - 2 : 0); pjx <= ij; ++pjx) {
      tup5_t nd = cache[pjx];
      if ((double)c != nd.c) {
        double d = (j - nd.j) - ((double)i - nd.i);
        double pen = 100.0 + 100.0 * (d * d);
        nd.cum = nd.cum - pen;
      }
      if (nd.i >= (double)(i - 2) && nd.j <= j && nd.cum >= prev.cum) { prev = nd; prev_id = cache_id[pjx]; }
    }
    double cum = prev.cum + q;
    cache[ij].j = j; cache[ij].i = i; cache[ij].c = c; cache[ij].q = q; cache[ij].cum = cum;
    cache_id[ij] = (int32_t)p;
    double jump = cum - 1000.0;
    if (frontier_val < jump) {
      node2_t me = {jump, j, (int32_t)p, (int32_t)p};
      int64_t k = size + j_rank[p];
      if (better2(&me, &tree[k])) {
        tree[k] = me;
        for (k >>= 1; k >= 1; k >>= 1) tree[k] = better2(&tree[2 * k + 1], &tree[2 * k]) ? tree[2 * k + 1] : tree[2 * k];
      }
    }
    double cjump = cum - 50.0;
    if (cl.cum < cjump) { cbest[c].j = j; cbest[c].i = i; cbest[c].q = q; cbest[c].cum = cjump; cbest_id[c] = (int32_t)p; }
    back[p] = prev;
    back_id[p] = prev_id;
  }
  /* traceback from the frontier's last entry = overall best (val, then smallest j, earliest) */
  node2_t top = tree[1];
  int64_t len = 0;
  if (top.id >= 0) {
    for (int32_t p = top.id; p >= 0; p = back_id[p]) len++;
    int64_t w = len;
    int32_t p = top.id;
    tup5_t row = {pj[p], (double)pi[p], (double)pc[p], pq[p], top.val};
    while (p >= 0) {
      --w;
      path_out[5 * w + 0] = row.j; path_out[5 * w + 1] = row.i; path_out[5 * w + 2] = row.c;
      path_out[5 * w + 3] = row.q; path_out[5 * w + 4] = row.cum;
      row = back[p];
      p = back_id[p];
    }
  }
  free(tree); free(cache); free(cbest); free(back); free(back_id); free(cache_id); free(cbest_id);
  return len;
}
