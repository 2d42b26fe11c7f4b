// Host stage of align() that stays on the CPU (reference describealign.py:702-767 and the assembly of the
// rate-change LP, :769-836), in C++ instead of numpy.  SURVEY.md section 8(f) N1.
//
//   dab_host_continuity_error   distance of every pass-1 path point from lines through its smoothed future /
//                               past neighbours (:706-724)
//   dab_host_compress_path      70 -> 1 compression of well-behaved runs and merging of equal audio indices
//                               (:743-767, quirks included)
//   dab_host_lp_assemble        objective, equality constraints (CSC, as scipy's coo.tocsc() lays them out) and
//                               right-hand side of the L1 rate-change fit (:769-836)
//   dab_host_line_clusters      grouping of the fit's points into co-linear clusters (:861-884)
//
// scipy.optimize.linprog itself (HiGHS) stays in Python: the LP has degenerate optima and only the same solver
// on the same input reproduces the reference's segments - which is why every number handed to it is formed in
// the order numpy / OpenBLAS form it: np.convolve = one OpenBLAS ddot per output (SkylakeX kernel order, edges as
// shorter dots), np.sum / np.mean of float64 = numpy's pairwise summation, everything else IEEE element-wise.
// Plain host code: no CUDA call in this file, usable without a device.  Compiled with -ffp-contract=off.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/describealign_b200.h"
#include "hann_tables.h"

namespace {

// OpenBLAS 0.3.30 SkylakeX ddot order (SURVEY.md B.2 iv), host copy of common.cuh's ddot_skx
double ddot_skx_host(const double *x, const double *y, int n) {
  double a[4][4] = {};
  const int n1 = n & ~15, n32 = n1 & ~31;
  int i = 0;
  if (n32) {
    double z[4][8] = {};
    for (; i < n32; i += 32)
      for (int k = 0; k < 4; ++k)
        for (int l = 0; l < 8; ++l) z[k][l] = std::fma(x[i + 8 * k + l], y[i + 8 * k + l], z[k][l]);
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 4; ++l) a[k][l] = z[k][l] + z[k][l + 4];
  }
  for (; i < n1; i += 16)
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 4; ++l) a[k][l] = std::fma(x[i + 4 * k + l], y[i + 4 * k + l], a[k][l]);
  const double s0 = ((a[0][0] + a[1][0]) + a[2][0]) + a[3][0];
  const double s1 = ((a[0][1] + a[1][1]) + a[2][1]) + a[3][1];
  const double s2 = ((a[0][2] + a[1][2]) + a[2][2]) + a[3][2];
  const double s3 = ((a[0][3] + a[1][3]) + a[2][3]) + a[3][3];
  double dot = (s0 + s2) + (s1 + s3);
  for (; i < n; ++i) dot = std::fma(y[i], x[i], dot);
  return dot;
}

// numpy's pairwise summation of contiguous float64 (np.sum / np.mean / np.add.reduce)
double pairwise_sum(const double *a, int64_t n) {
  if (n < 8) {
    double res = 0.0;
    for (int64_t i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

// np.convolve(a, kernel, mode="valid") for len(a) >= klen: out[k] = ddot(a[k : k + klen], kernel reversed)
void convolve_valid(const double *a, int64_t n, const double *kernel, int klen, std::vector<double> &out) {
  std::vector<double> kr(klen);
  for (int i = 0; i < klen; ++i) kr[i] = kernel[klen - 1 - i];
  const int64_t m = n - klen + 1;
  out.resize(m > 0 ? m : 0);
  for (int64_t k = 0; k < m; ++k) out[k] = ddot_skx_host(a + k, kr.data(), klen);
}

// np.convolve(window41, a, mode="same")[:n] for n >= 41: output t is the dot of a[t-20 .. t+20] clipped to the array
// with the flipped window, edges as shorter dots
void local_mean41(const double *a, int64_t n, const double *window, std::vector<double> &out) {
  double flip[41];
  for (int k = 0; k < 41; ++k) flip[k] = window[40 - k];
  out.resize(n);
  for (int64_t t = 0; t < n; ++t) {
    int64_t lo = t - 20, hi = t + 21, klo = 0;
    if (lo < 0) { klo = -lo; lo = 0; }
    if (hi > n) hi = n;
    out[t] = ddot_skx_host(a + lo, flip + klo, (int)(hi - lo));
  }
}

void window_and_head(double *window, double *head, double *head_rev) {
  std::memcpy(window, DAB_HANN41_F64, sizeof(double) * 41);
  // head = window[:20] / np.sum(window[:20])
  const double s = pairwise_sum(window, 20);
  for (int i = 0; i < 20; ++i) head[i] = window[i] / s;
  for (int i = 0; i < 20; ++i) head_rev[i] = head[19 - i];
}

constexpr int SPN = 21, HALF = SPN / 2, DELAY = SPN + HALF - 2;   // frames per node, 10, 29

}  // namespace

extern "C" {

// err has n - (deriv ? 1 : 0) entries.  n must be > DELAY + 20 (the reference's own arrays are far longer).
int dab_host_continuity_error(const int64_t *x, const int64_t *y, int64_t n, int deriv, double *err) {
  return dab_host_continuity_error_f64(nullptr, nullptr, x, y, n, deriv, err);
}

int dab_host_continuity_error_f64(const double *xf_in, const double *yf_in, const int64_t *xi, const int64_t *yi, int64_t n,
                                  int deriv, double *err) {
  if (n < 0 || !err || (!(xf_in && yf_in) && !(xi && yi))) return DAB_E_ARG;
  if (n < DELAY + 20 + 2) return DAB_E_ARG;
  double window[41], head[20], head_rev[20];
  window_and_head(window, head, head_rev);
  std::vector<double> xf(n), yf(n);
  for (int64_t k = 0; k < n; ++k) {
    xf[k] = xf_in ? xf_in[k] : (double)xi[k];
    yf[k] = yf_in ? yf_in[k] : (double)yi[k];
  }
  const int shift = deriv ? 1 : 0;
  const int64_t ne = n - shift;
  const double inf = std::numeric_limits<double>::infinity();
  for (int64_t k = 0; k < ne; ++k) err[k] = inf;
  std::vector<double> xs, ys;
  const int64_t m = n - 20 + 1;        // length of the 'valid' convolutions
  const int64_t ml = m - HALF;         // lines per direction
  // future lines (kernel = head)
  convolve_valid(xf.data(), n, head, 20, xs);
  convolve_valid(yf.data(), n, head, 20, ys);
  const int64_t kk = DELAY - shift;
  // err[:-kk] = |slope_f * x[:-DELAY] + off_f - y[:-DELAY]|  (both sides have ml = n - DELAY entries)
  if (ne - kk != ml || n - DELAY != ml) return DAB_E_ARG;
  for (int64_t k = 0; k < ml; ++k) {
    const double slope = (ys[k + HALF] - ys[k]) / (xs[k + HALF] - xs[k]);
    const double off = ys[k] - xs[k] * slope;
    err[k] = std::fabs(slope * xf[k] + off - yf[k]);
  }
  // past lines (kernel = head reversed)
  convolve_valid(xf.data(), n, head_rev, 20, xs);
  convolve_valid(yf.data(), n, head_rev, 20, ys);
  for (int64_t k = 0; k < ml; ++k) {
    const double slope = (ys[k + HALF] - ys[k]) / (xs[k + HALF] - xs[k]);
    const double off = ys[k + HALF] - xs[k + HALF] * slope;
    const double e = std::fabs(slope * xf[k + DELAY] + off - yf[k + DELAY]);
    double &dst = err[kk + k];
    // np.minimum propagates NaN
    dst = (std::isnan(dst) || std::isnan(e)) ? std::numeric_limits<double>::quiet_NaN() : (e < dst ? e : dst);
  }
  return DAB_OK;
}

// Compression of the kept pass-1 path.  out_x / out_y need room for n entries; *n_out receives the count.
// Returns DAB_E_ARG for a path too short to compress (the reference raises "Alignment failed" there).
int dab_host_compress_path(const int64_t *x, const int64_t *y, int64_t n, double *out_x, double *out_y, int64_t *n_out) {
  if (!x || !y || !out_x || !out_y || !n_out || n < 41) return DAB_E_ARG;
  if (n - 80 <= 10) return DAB_E_ARG;      // range(10, n - 80, 70) is empty
  double window[41];
  std::memcpy(window, DAB_HANN41_F64, sizeof(window));
  std::vector<double> xf(n), yf(n), sx, sy;
  for (int64_t k = 0; k < n; ++k) { xf[k] = (double)x[k]; yf[k] = (double)y[k]; }
  local_mean41(xf.data(), n, window, sx);
  local_mean41(yf.data(), n, window, sy);
  std::vector<double> err(n - 1);
  for (int64_t k = 0; k + 1 < n; ++k) {
    const double slope = (sy[k + 1] - sy[k]) / (sx[k + 1] - sx[k]);
    const double offset = sy[k] - sx[k] * slope;
    err[k] = slope * xf[k] + offset - yf[k];
  }
  std::vector<double> cx, cy;
  cx.reserve(n); cy.reserve(n);
  for (int64_t k = 0; k < 10; ++k) { cx.push_back(xf[k]); cy.push_back(yf[k]); }
  int64_t start = 10, last = -1;
  for (; start < n - 80; start += 70) {
    last = start;
    bool ok = true;
    for (int64_t k = start; k < start + 70; ++k)
      if (!(std::fabs(err[k]) < 3)) { ok = false; break; }
    if (ok) {
      // np.mean of 70 integers: exact sum, one division
      int64_t sxi = 0, syi = 0;
      for (int64_t k = start; k < start + 70; ++k) { sxi += x[k]; syi += y[k]; }
      cx.push_back((double)sxi / 70.0);
      cy.push_back((double)syi / 70.0);
    } else {
      for (int64_t k = start; k < start + 70; ++k) { cx.push_back(xf[k]); cy.push_back(yf[k]); }
    }
  }
  // one further fixed slice is kept raw; points after it are dropped (reference quirk)
  for (int64_t k = last + 70; k < last + 140 && k < n; ++k) { cx.push_back(xf[k]); cy.push_back(yf[k]); }
  // merge entries that share an audio index, in order of first appearance; the merged value is np.mean of the list
  std::unordered_map<double, int64_t> slot;
  std::vector<std::vector<double>> groups;
  std::vector<double> keys;
  for (size_t k = 0; k < cx.size(); ++k) {
    auto it = slot.find(cx[k]);
    if (it == slot.end()) {
      slot.emplace(cx[k], (int64_t)groups.size());
      keys.push_back(cx[k]);
      groups.emplace_back(1, cy[k]);
    } else {
      groups[it->second].push_back(cy[k]);
    }
  }
  for (size_t g = 0; g < groups.size(); ++g) {
    out_x[g] = keys[g];
    out_y[g] = pairwise_sum(groups[g].data(), (int64_t)groups[g].size()) / (double)groups[g].size();
  }
  *n_out = (int64_t)groups.size();
  return DAB_OK;
}

// The LP of describealign.py:769-836 over n fit points (x, y float64): variables in the reference's order
//   fit_err+-  [2n] | jump+- [2(n-1)] | shot_noise+- [2n] | shot_noise_jump+- [2(n-1)] | rate_change_jump+- [2(n-1)] |
//   rate_change+- [2(n-2)] | median slope [1]
// cost[12n-9]; equality constraints A (3n-4 rows) in CSC: indptr[12n-8], indices / data [*nnz = 23n - 21 <= 24n], b[3n-4].
int dab_host_lp_assemble(const double *x, const double *y, int64_t n, double *cost, int32_t *indptr, int32_t *indices,
                         double *data, double *b, int64_t *nnz_out) {
  if (!x || !y || !cost || !indptr || !indices || !data || !b || !nnz_out || n < DELAY + 22) return DAB_E_ARG;
  std::vector<double> cerr(n - 1);
  int rc = dab_host_continuity_error_f64(x, y, nullptr, nullptr, n, 1, cerr.data());
  if (rc != DAB_OK) return rc;
  std::vector<double> dx(n - 1), dy(n - 1), inv(n - 1), jump(n - 1);
  for (int64_t k = 0; k + 1 < n; ++k) {
    dx[k] = x[k + 1] - x[k];
    dy[k] = y[k + 1] - y[k];
    inv[k] = 1.0 / dx[k];
    const double s = std::sqrt(cerr[k] / 3.0);
    // np.maximum(1, s) propagates NaN
    const double d = std::isnan(s) ? s : (s > 1.0 ? s : 1.0);
    jump[k] = 10.0 / d;
  }
  // ---- cost ----
  int64_t c = 0;
  for (int64_t k = 0; k < 2 * n; ++k) cost[c++] = 1.0;
  for (int rep = 0; rep < 2; ++rep)
    for (int64_t k = 0; k + 1 < n; ++k) cost[c++] = jump[k];
  for (int64_t k = 0; k < 2 * n; ++k) cost[c++] = .01;
  for (int64_t k = 0; k < 2 * (n - 1); ++k) cost[c++] = 3.0;
  for (int64_t k = 0; k < 2 * (n - 1); ++k) cost[c++] = .001;
  for (int64_t k = 0; k < 2 * (n - 2); ++k) cost[c++] = 10.0 * 4000;
  cost[c++] = 0.0;
  // ---- columns ----
  const int64_t c_fe_p = 0, c_fe_m = n, c_j_p = 2 * n, c_j_m = 3 * n - 1, c_sn_p = 4 * n - 2, c_sn_m = 5 * n - 2,
                c_snj_p = 6 * n - 2, c_snj_m = 7 * n - 3, c_rcj_p = 8 * n - 4, c_rcj_m = 9 * n - 5, c_rc_p = 10 * n - 6,
                c_rc_m = 11 * n - 8, c_med = 12 * n - 10;
  const int64_t R2 = n - 1, R3 = 2 * (n - 1);     // first rows of blocks 2 and 3
  int64_t p = 0;
  auto put = [&](int64_t row, double v) { indices[p] = (int32_t)row; data[p] = v; ++p; };
  for (int64_t col = 0; col <= c_med; ++col) {
    indptr[col] = (int32_t)p;
    if (col < c_j_p) {
      // fit_err+ (sign +1) / fit_err- (sign -1), point k: row k-1 gets +inv[k-1], row k gets -inv[k]
      const double sign = col < c_fe_m ? 1.0 : -1.0;
      const int64_t k = col - (col < c_fe_m ? c_fe_p : c_fe_m);
      if (k >= 1) put(k - 1, sign * inv[k - 1]);
      if (k < n - 1) put(k, sign * -inv[k]);
    } else if (col < c_sn_p) {
      const double sign = col < c_j_m ? 1.0 : -1.0;
      const int64_t k = col - (col < c_j_m ? c_j_p : c_j_m);
      put(k, sign * inv[k]);
    } else if (col < c_snj_p) {
      // shot noise of point k: block 2 row k-1 gets +-1, row k gets -+1
      const bool plus = col < c_sn_m;
      const int64_t k = col - (plus ? c_sn_p : c_sn_m);
      if (k >= 1) put(R2 + k - 1, plus ? 1.0 : -1.0);
      if (k < n - 1) put(R2 + k, plus ? -1.0 : 1.0);
    } else if (col < c_rcj_p) {
      const bool plus = col < c_snj_m;
      const int64_t k = col - (plus ? c_snj_p : c_snj_m);
      put(k, (plus ? 1.0 : -1.0) * inv[k]);
      put(R2 + k, plus ? -1.0 : 1.0);
    } else if (col < c_rc_p) {
      // rate-change jump of interval k: block 1 row k; block 3 row k-1 gets +inv[k], row k gets -inv[k]
      const double sign = col < c_rcj_m ? 1.0 : -1.0;
      const int64_t k = col - (col < c_rcj_m ? c_rcj_p : c_rcj_m);
      put(k, sign * inv[k]);
      if (k >= 1) put(R3 + k - 1, sign * inv[k]);
      if (k < n - 2) put(R3 + k, sign * -inv[k]);
    } else if (col < c_med) {
      const bool plus = col < c_rc_m;
      const int64_t k = col - (plus ? c_rc_p : c_rc_m);
      put(R3 + k, plus ? -1.0 : 1.0);
    } else {
      for (int64_t r = 0; r < n - 1; ++r) put(r, 1.0);
    }
  }
  indptr[c_med + 1] = (int32_t)p;
  *nnz_out = p;
  for (int64_t k = 0; k + 1 < n; ++k) b[k] = dy[k] / dx[k];
  for (int64_t k = n - 1; k < 3 * n - 4; ++k) b[k] = 0.0;
  return DAB_OK;
}

// Grouping of the fitted path's points into co-linear clusters (describealign.py:861-884): every point votes, with
// the slope on either side of it, for the line (slope rounded to 6 decimals, integer offset) through it; the largest
// groups absorb every other group whose first and last point lie within 3 frames of their line; clusters are sorted,
// and those spanning more than 10 frames with more than 5 points are kept.  x, y: the n fit points with the fit error
// removed from y; slopes: the n + 1 slopes (first and last repeated).  Output: the kept clusters' points,
// concatenated (room for 2n each), and cluster_start[n_clusters + 1] (room for 2n + 1).  The per-cluster line fit
// (np.linalg.lstsq) stays with the caller.
int dab_host_line_clusters(const double *x, const double *y, const double *slopes, int64_t n, double *out_x, double *out_y,
                           int64_t *cluster_start, int64_t *n_clusters) {
  if (!x || !y || !slopes || !out_x || !out_y || !cluster_start || !n_clusters || n < 1) return DAB_E_ARG;
  struct Group {
    double slope;
    int64_t offset;
    std::vector<std::pair<double, double>> pts;
    bool alive = true;
  };
  struct Key {
    uint64_t s;
    int64_t o;
    bool operator==(const Key &k) const { return s == k.s && o == k.o; }
  };
  struct KeyHash {
    size_t operator()(const Key &k) const { return (size_t)(k.s * 0x9e3779b97f4a7c15ull) ^ (size_t)(k.o * 0xc2b2ae3d27d4eb4full); }
  };
  std::vector<Group> groups;
  std::unordered_map<Key, size_t, KeyHash> index;
  for (int64_t k = 0; k < n; ++k) {
    for (int side = 0; side < 2; ++side) {
      const double slope = slopes[k + side];
      if (slope < .1 || slope > 10) continue;
      // round(np.float64, 6) = rint(v * 1e6) / 1e6; round(np.float64, 0) = rint(v)
      double ks = std::nearbyint(slope * 1e6) / 1e6;
      if (ks == 0.0) ks = 0.0;                                       // -0.0 and 0.0 are one dict key
      const double off = std::nearbyint(y[k] - slope * x[k]);
      if (!(std::fabs(off) < 9e18)) return DAB_E_ARG;                // int(nan) / int(inf) raise in the reference
      Key key;
      std::memcpy(&key.s, &ks, sizeof(double));
      key.o = (int64_t)off;
      auto it = index.find(key);
      size_t g;
      if (it == index.end()) {
        g = groups.size();
        index.emplace(key, g);
        groups.emplace_back();
        groups[g].slope = ks;
        groups[g].offset = key.o;
      } else {
        g = it->second;
      }
      groups[g].pts.emplace_back(x[k], y[k]);
    }
  }
  // largest first, ties in order of first appearance (sizes as they are now)
  std::vector<size_t> order(groups.size());
  for (size_t g = 0; g < groups.size(); ++g) order[g] = g;
  std::vector<size_t> size0(groups.size());
  for (size_t g = 0; g < groups.size(); ++g) size0[g] = groups[g].pts.size();
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return size0[a] > size0[b]; });
  std::vector<size_t> seeds;
  for (size_t g : order) {
    if (!groups[g].alive) continue;
    groups[g].alive = false;
    const double slope = groups[g].slope;
    const double offset = (double)groups[g].offset;
    for (size_t h = 0; h < groups.size(); ++h) {
      if (!groups[h].alive) continue;
      const auto &first = groups[h].pts.front();
      const auto &last = groups[h].pts.back();
      if (std::fabs(first.second - (first.first * slope + offset)) < 3 && std::fabs(last.second - (last.first * slope + offset)) < 3) {
        groups[g].pts.insert(groups[g].pts.end(), groups[h].pts.begin(), groups[h].pts.end());
        groups[h].alive = false;
      }
    }
    seeds.push_back(g);
  }
  int64_t at = 0, nc = 0;
  cluster_start[0] = 0;
  for (size_t g : seeds) {
    auto &c = groups[g].pts;
    std::sort(c.begin(), c.end());
    if (!(std::fabs(c.front().first - c.back().first) > 10 && c.size() > 5)) continue;
    for (const auto &p : c) { out_x[at] = p.first; out_y[at] = p.second; ++at; }
    cluster_start[++nc] = at;
  }
  *n_clusters = nc;
  return DAB_OK;
}

// Drift dynamic programme of the pitch-preserving time stretch and its traceback (describealign.py:320-371; SURVEY.md
// 8f N3): per 512-sample window and drift (0 .. 3072) the cheapest way to have reached it - no jump, or a jump of one
// of the candidate distances from two windows back at the loss of that window's best position - then the walk back
// from the last window.  loc / best: [windows][n_jumps] as dab_stretch_best_jumps returns them.  out_at / out_dist:
// room for `windows` entries; jumps come out in input order, distances still unsigned (the caller flips the sign
// for a longer output).  Returns DAB_E_ARG where the reference's own index arithmetic would leave the arrays
// (absurd stretch ratios): the caller then runs the numpy statement, which raises what the reference raises.
int dab_host_stretch_plan(int64_t n_in, int64_t n_out, const int32_t *jumps, int32_t n_jumps, const int16_t *loc,
                          const double *best, int64_t *out_at, int64_t *out_dist, int64_t *count) {
  if (!jumps || !loc || !best || !out_at || !out_dist || !count || n_jumps < 1) return DAB_E_ARG;
  constexpr int W = 512, MAXD = 3 * 512, WIDTH = 2 * MAXD + 1;
  const int64_t total = n_out - n_in;
  const int64_t nw = n_in / W;
  if (nw < 2) return DAB_E_ARG;
  auto floordiv = [](int64_t a, int64_t b) { int64_t q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; };
  auto offset_at = [&](int64_t w) {
    const int64_t c = w < 0 ? 0 : (w > nw - 1 ? nw - 1 : w);
    return floordiv(total * c, nw - 1);
  };
  auto offset_step = [&](int64_t w) { const int64_t d = offset_at(w) - offset_at(w - 1); return d < 0 ? -d : d; };
  for (int k = 0; k < n_jumps; ++k)
    if (jumps[k] < 1 || jumps[k] >= WIDTH) return DAB_E_ARG;
  const double inf = std::numeric_limits<double>::infinity();
  std::vector<int16_t> back((size_t)nw * WIDTH);
  std::vector<double> cum(3 * (size_t)WIDTH, inf), next(WIDTH), losses(n_jumps);
  cum[1 * WIDTH + MAXD] = 0.0;
  cum[2 * WIDTH + MAXD] = 0.0;
  int64_t last_step = 0;
  for (int64_t w = 0; w < nw; ++w) {
    const int64_t step = offset_step(w), step2 = step + last_step;
    if (step2 >= WIDTH) return DAB_E_ARG;
    for (int k = 0; k < n_jumps; ++k) losses[k] = 1.0 - best[w * n_jumps + k];
    const double *one = cum.data() + ((w + 2) % 3) * WIDTH;      // (w - 1) mod 3
    const double *two = cum.data() + ((w + 1) % 3) * WIDTH;      // (w - 2) mod 3
    int16_t *bk = back.data() + (size_t)w * WIDTH;
    // row 0: no jump
    for (int c = 0; c < WIDTH; ++c) {
      next[c] = c < WIDTH - step ? one[c + step] : inf;
      bk[c] = 0;
    }
    // rows 1..: a jump from two windows back; the first minimum of a column wins (np.argmin)
    for (int k = 0; k < n_jumps; ++k) {
      const int64_t jump = jumps[k], cut = step2 - jump;
      const int64_t lo = jump, hi = WIDTH - (cut > 0 ? cut : 0);
      const double loss = losses[k];
      const double *src = two + (step2 - jump);
      for (int64_t c = lo; c < hi; ++c) {
        const double v = src[c] + loss;
        if (v < next[c]) { next[c] = v; bk[c] = (int16_t)(k + 1); }
      }
    }
    std::memcpy(cum.data() + (w % 3) * WIDTH, next.data(), sizeof(double) * WIDTH);
    last_step = step;
  }
  int64_t drift = MAXD, m = 0;
  bool skip = false;
  for (int64_t w = nw - 1; w >= 0; --w) {
    drift += offset_step(w + 1);
    if (skip) { skip = false; continue; }
    if (drift < 0 || drift >= WIDTH) return DAB_E_ARG;
    const int k = (int)back[(size_t)w * WIDTH + drift] - 1;
    if (k == -1) continue;
    out_at[m] = w * W + (int64_t)loc[w * n_jumps + k];
    out_dist[m] = jumps[k];
    ++m;
    drift -= jumps[k];
    skip = true;
  }
  for (int64_t a = 0, b = m - 1; a < b; ++a, --b) { std::swap(out_at[a], out_at[b]); std::swap(out_dist[a], out_dist[b]); }
  *count = m;
  return DAB_OK;
}

}  // extern "C"
