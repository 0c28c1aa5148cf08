// R-compatible random draws and k-means on the HOST (no GPU involved): what lets the Python host layer replay an R
// session -- `set.seed(1234); haystack(dat.tsne, dat.expression)` -- and hand the B200 the very grid, permutation ids
// and cross-validation folds R would have produced (host half).
//
// Under R none of this is used: the shim calls R's own kmeans() / sample() (rshim/R/haystack_continuous_highD.R).
// Call sites in the reference: sample() at R/haystack_continuous.R:259, :275, :437, :505 and R/haystack_highD.R:95,
// :110; kmeans() at R/haystack_highD.R:83.  The algorithms are base R's, which the reference does not vendor:
// Mersenne-Twister (Matsumoto & Nishimura 1998) seeded by set.seed()'s LCG scrambling, R_unif_index ("Rejection"
// sampling since R 3.6.0, "Rounding" before), the partial Fisher-Yates of sample.int(), the sorted-cumulative and
// Walker-alias paths of sample(prob =) for a single draw, and Hartigan & Wong's Algorithm AS 136 as stats::kmeans
// runs it (nstart restarts from distinct rows).  Pinned end to end: with these the GPU path reproduces the table
// printed in docs/articles/a01_toy_example.html (tests/test_gpu_r_parity.py).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <numeric>
#include <vector>

#include "hs_common.cuh"

struct hs_rng {
  static constexpr int kN = 624, kM = 397;
  uint32_t mt[kN];
  int pos = kN;
  bool rounding = false;

  void seed(uint32_t s) {
    for (int i = 0; i < 50; ++i) s = 69069u * s + 1u;       // initial scrambling
    s = 69069u * s + 1u;                                     // word 0 of R's seed vector is the position counter
    for (int i = 0; i < kN; ++i) {
      s = 69069u * s + 1u;
      mt[i] = s;
    }
    pos = kN;
  }
  void refill() {
    auto twist = [](uint32_t u, uint32_t v) { return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u); };
    for (int k = 0; k < kN - kM; ++k) mt[k] = mt[k + kM] ^ twist(mt[k], mt[k + 1]);
    for (int k = kN - kM; k < kN - 1; ++k) mt[k] = mt[k + kM - kN] ^ twist(mt[k], mt[k + 1]);
    mt[kN - 1] = mt[kM - 1] ^ twist(mt[kN - 1], mt[0]);
    pos = 0;
  }
  uint32_t next32() {
    if (pos >= kN) refill();
    uint32_t y = mt[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  double unif() {      // unif_rand(): kept inside the open interval
    const double eps = 2.328306437080797e-10;
    const double v = (double)next32() * 2.3283064365386963e-10;
    if (v <= 0.0) return 0.5 * eps;
    if (1.0 - v <= 0.0) return 1.0 - 0.5 * eps;
    return v;
  }
  double index(double dn) {      // R_unif_index
    if (rounding) return floor(dn * unif());
    if (dn <= 0) return 0.0;
    const int bits = (int)ceil(log2(dn));
    for (;;) {
      int64_t v = 0;
      for (int n = 0; n <= bits; n += 16) v = 65536 * v + (int)floor(unif() * 65536.0);
      if (bits < 64) v &= ((int64_t)1 << bits) - 1;
      if ((double)v < dn) return (double)v;
    }
  }
  // sample.int(n, k), no replacement, no weights: 1-based ids
  void sample_int(int n, int k, int32_t* out, std::vector<int32_t>& pool) {
    if (k < 2) {
      for (int i = 0; i < k; ++i) out[i] = (int32_t)(index((double)n) + 1);
      return;
    }
    pool.resize((size_t)n);
    std::iota(pool.begin(), pool.end(), 0);
    for (int i = 0; i < k; ++i) {
      const int j = (int)index((double)n);
      out[i] = pool[(size_t)j] + 1;
      pool[(size_t)j] = pool[(size_t)--n];
    }
  }
};

namespace {

// ---- Hartigan & Wong, AS 136 -------------------------------------------------------------------------------------
// a: m x n column-major data; c: k x n column-major centres (in: initial, out: final).  Step numbers are 1-based as in
// the published algorithm (live sets, ncp).
struct HartiganWong {
  const double* a;
  int m, n, k;
  std::vector<double> c, an1, an2, d, wss;
  std::vector<int> ic1, ic2, nc, ncp, itran, live;
  static constexpr double kBig = 1e30;

  double& C(int l, int j) { return c[(size_t)l + (size_t)j * k]; }
  double A(int i, int j) const { return a[(size_t)i + (size_t)j * m]; }
  double dist2(int i, int l) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) {
      const double t = A(i, j) - C(l, j);
      s += t * t;
    }
    return s;
  }
  // squared distance, abandoned (returns false) as soon as it reaches `limit`
  bool dist2_below(int i, int l, double limit, double& out) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) {
      const double t = A(i, j) - C(l, j);
      s += t * t;
      if (s >= limit) return false;
    }
    out = s;
    return true;
  }
  void move(int i, int l1, int l2) {      // point i leaves l1 for l2: centres, sizes and the two factors
    const double al1 = nc[l1], alw = al1 - 1.0, al2 = nc[l2], alt = al2 + 1.0;
    for (int j = 0; j < n; ++j) {
      C(l1, j) = (C(l1, j) * al1 - A(i, j)) / alw;
      C(l2, j) = (C(l2, j) * al2 + A(i, j)) / alt;
    }
    --nc[l1];
    ++nc[l2];
    an2[l1] = alw / al1;
    an1[l1] = alw > 1.0 ? alw / (alw - 1.0) : kBig;
    an1[l2] = alt / al2;
    an2[l2] = alt / (alt + 1.0);
    ic1[i] = l2;
    ic2[i] = l1;
  }

  void optimal_transfer(int& indx) {
    for (int l = 0; l < k; ++l)
      if (itran[l] == 1) live[l] = m;
    for (int i = 0; i < m; ++i) {
      ++indx;
      const int l1 = ic1[i];
      if (nc[l1] != 1) {
        if (ncp[l1] != 0) d[i] = dist2(i, l1) * an1[l1];
        int l2 = ic2[i];
        const int ll = l2;
        double r2 = dist2(i, l2) * an2[l2];
        for (int l = 0; l < k; ++l) {
          if ((i + 1 >= live[l1] && i + 1 >= live[l]) || l == l1 || l == ll) continue;
          double dc;
          if (!dist2_below(i, l, r2 / an2[l], dc)) continue;
          r2 = dc * an2[l];
          l2 = l;
        }
        if (r2 >= d[i]) {
          ic2[i] = l2;
        } else {
          indx = 0;
          live[l1] = live[l2] = m + i + 1;
          ncp[l1] = ncp[l2] = i + 1;
          move(i, l1, l2);
        }
      }
      if (indx == m) return;
    }
    for (int l = 0; l < k; ++l) {
      itran[l] = 0;
      live[l] -= m;
    }
  }

  // returns false when the step limit is hit
  bool quick_transfer(int& indx, int max_steps) {
    int icoun = 0, istep = 0;
    for (;;) {
      for (int i = 0; i < m; ++i) {
        ++icoun;
        ++istep;
        if (istep >= max_steps) return false;
        const int l1 = ic1[i], l2 = ic2[i];
        if (nc[l1] != 1) {
          if (istep <= ncp[l1]) d[i] = dist2(i, l1) * an1[l1];
          if (istep < ncp[l1] || istep < ncp[l2]) {
            double dd;
            if (dist2_below(i, l2, d[i] / an2[l2], dd)) {
              icoun = 0;
              indx = 0;
              itran[l1] = itran[l2] = 1;
              ncp[l1] = ncp[l2] = istep + m;
              move(i, l1, l2);
            }
          }
        }
        if (icoun == m) return true;
      }
    }
  }

  // ifault: 0 converged, 1 empty cluster, 2 iteration limit, 3 bad k, 4 quick-transfer step limit
  int run(int iter_max, int max_qtran) {
    if (k <= 1 || k >= m) return 3;
    an1.assign(k, 0.0); an2.assign(k, 0.0); d.assign(m, 0.0); wss.assign(k, 0.0);
    ic1.assign(m, 0); ic2.assign(m, 1); nc.assign(k, 0); ncp.assign(k, -1); itran.assign(k, 1); live.assign(k, 0);
    for (int i = 0; i < m; ++i) {      // the two closest centres of every point
      double d0 = dist2(i, 0), d1 = dist2(i, 1);
      ic1[i] = 0;
      ic2[i] = 1;
      if (d0 > d1) {
        std::swap(d0, d1);
        ic1[i] = 1;
        ic2[i] = 0;
      }
      for (int l = 2; l < k; ++l) {
        double db;
        if (!dist2_below(i, l, d1, db)) continue;
        if (db < d0) {
          d1 = d0;
          ic2[i] = ic1[i];
          d0 = db;
          ic1[i] = l;
        } else {
          d1 = db;
          ic2[i] = l;
        }
      }
    }
    std::fill(c.begin(), c.end(), 0.0);
    for (int i = 0; i < m; ++i) {
      ++nc[ic1[i]];
      for (int j = 0; j < n; ++j) C(ic1[i], j) += A(i, j);
    }
    for (int l = 0; l < k; ++l)
      if (nc[l] == 0) return 1;
    for (int l = 0; l < k; ++l) {
      const double aa = nc[l];
      for (int j = 0; j < n; ++j) C(l, j) /= aa;
      an2[l] = aa / (aa + 1.0);
      an1[l] = aa > 1.0 ? aa / (aa - 1.0) : kBig;
    }
    int indx = 0, ifault = 2;
    for (int it = 1; it <= iter_max; ++it) {
      optimal_transfer(indx);
      if (indx == m) { ifault = 0; break; }
      if (!quick_transfer(indx, max_qtran)) { ifault = 4; break; }
      if (k == 2) { ifault = 0; break; }
      std::fill(ncp.begin(), ncp.end(), 0);
    }
    // final centres (cluster means) and within-cluster sums of squares
    std::fill(c.begin(), c.end(), 0.0);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) C(ic1[i], j) += A(i, j);
    for (int j = 0; j < n; ++j) {
      for (int l = 0; l < k; ++l) C(l, j) /= (double)nc[l];
      for (int i = 0; i < m; ++i) {
        const double t = A(i, j) - C(ic1[i], j);
        wss[ic1[i]] += t * t;
      }
    }
    return ifault;
  }
};

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int hs_r_rng_create(uint32_t seed, int sample_kind_rounding, hs_rng** out) {
  if (!out) return HS_ERR_INVALID;
  hs_rng* g = new (std::nothrow) hs_rng();
  if (!g) return HS_ERR_NOMEM;
  g->rounding = sample_kind_rounding != 0;
  g->seed(seed);
  *out = g;
  return HS_OK;
}

void hs_r_rng_destroy(hs_rng* g) { delete g; }

double hs_r_unif_rand(hs_rng* g) { return g ? g->unif() : NAN; }

int hs_r_sample_int(hs_rng* g, int32_t n, int32_t size, int32_t count, int32_t* ids_out) {
  try {
  if (!g || n < 1 || size < 0 || size > n || count < 0 || (!ids_out && (int64_t)size * count > 0))
    return hs_fail(nullptr, HS_ERR_INVALID, "hs_r_sample_int: bad arguments (cannot take a sample larger than the population)");
  if (n > 10000000 && size <= n / 2)
    return hs_fail(nullptr, HS_ERR_INVALID, "hs_r_sample_int: R switches to a hashing algorithm for n > 1e7; not available");
  std::vector<int32_t> pool;
  for (int32_t c = 0; c < count; ++c) g->sample_int(n, size, ids_out + (size_t)c * size, pool);
  return HS_OK;
  } catch (...) { return hs_fail(nullptr, HS_ERR_NOMEM, "hs_r_sample_int: host allocation failed"); }
}

int hs_r_sample_prob_one(hs_rng* g, int32_t n, const double* prob, int32_t* id_out) {
  try {
  if (!g || n < 1 || !prob || !id_out) return hs_fail(nullptr, HS_ERR_INVALID, "hs_r_sample_prob_one: bad arguments");
  std::vector<double> p(prob, prob + n);
  double sum = 0.0;
  for (double v : p) {
    if (!(v >= 0.0) || !std::isfinite(v)) return hs_fail(nullptr, HS_ERR_INVALID, "hs_r_sample_prob_one: NA or negative probability");
    sum += v;
  }
  if (!(sum > 0.0)) return hs_fail(nullptr, HS_ERR_INVALID, "hs_r_sample_prob_one: too few positive probabilities");
  int large = 0;
  for (double& v : p) {
    v /= sum;
    if (n * v > 0.1) ++large;
  }
  if (large > 200) {      // Walker's alias method
    std::vector<double> q((size_t)n);
    std::vector<int> alias((size_t)n, 0), hl((size_t)n);
    int h = -1, l = n;
    for (int i = 0; i < n; ++i) {
      q[i] = p[i] * n;
      if (q[i] < 1.0) hl[++h] = i; else hl[--l] = i;
    }
    if (h >= 0 && l < n) {
      for (int k = 0; k < n - 1; ++k) {
        const int i = hl[k], j = hl[l];
        alias[i] = j;
        q[j] += q[i] - 1.0;
        if (q[j] < 1.0) ++l;
        if (l >= n) break;
      }
    }
    for (int i = 0; i < n; ++i) q[i] += i;
    const double ru = g->unif() * n;
    const int k = (int)ru;
    *id_out = ru < q[k] ? k + 1 : alias[k] + 1;
    return HS_OK;
  }
  // probabilities in descending order (R's revsort: a heap sort, so equal weights land in ITS order), cumulated, searched
  std::vector<int> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), 1);
  {
    double* a = p.data() - 1;
    int* ib = perm.data() - 1;
    if (n > 1) {
      int l = (n >> 1) + 1, ir = n;
      for (;;) {
        double ra;
        int ii;
        if (l > 1) {
          --l;
          ra = a[l];
          ii = ib[l];
        } else {
          ra = a[ir];
          ii = ib[ir];
          a[ir] = a[1];
          ib[ir] = ib[1];
          if (--ir == 1) {
            a[1] = ra;
            ib[1] = ii;
            break;
          }
        }
        int i = l, j = l << 1;
        while (j <= ir) {
          if (j < ir && a[j] > a[j + 1]) ++j;
          if (ra > a[j]) {
            a[i] = a[j];
            ib[i] = ib[j];
            i = j;
            j += j;
          } else {
            j = ir + 1;
          }
        }
        a[i] = ra;
        ib[i] = ii;
      }
    }
  }
  for (int i = 1; i < n; ++i) p[i] += p[i - 1];
  const double ru = g->unif();
  int j = 0;
  while (j < n - 1 && !(ru <= p[j])) ++j;
  *id_out = perm[j];
  return HS_OK;
  } catch (...) { return hs_fail(nullptr, HS_ERR_NOMEM, "hs_r_sample_prob_one: host allocation failed"); }
}

// stats::kmeans(x, centers = k, iter.max, nstart, algorithm = "Hartigan-Wong")$centers on R's random stream.
// x: m x d column-major.  centers_out: k x d column-major.  ifault_out (optional): AS 136 fault of the winning run
// (2 = iteration limit reached, which the reference silences: R/haystack_highD.R:82-84).
int hs_r_kmeans(hs_rng* g, const double* x, int32_t m, int32_t d, int32_t k, int32_t iter_max, int32_t nstart,
                double* centers_out, double* tot_withinss_out, int32_t* ifault_out) {
  try {
  if (!g || !x || !centers_out || m < 1 || d < 1 || k < 1 || nstart < 1 || iter_max < 1)
    return hs_fail(nullptr, HS_ERR_INVALID, "hs_r_kmeans: bad arguments");
  // rows the initial centres are drawn from: all rows for a single start (if they come out distinct), the distinct
  // rows (first occurrences, in order) otherwise
  auto row_less = [&](int a_, int b_) {
    for (int j = 0; j < d; ++j) {
      const double u = x[(size_t)a_ + (size_t)j * m], v = x[(size_t)b_ + (size_t)j * m];
      if (u != v) return u < v;
    }
    return false;
  };
  auto row_eq = [&](int a_, int b_) { return !row_less(a_, b_) && !row_less(b_, a_); };
  std::vector<int> distinct;
  auto build_distinct = [&]() {
    std::vector<int> ord((size_t)m);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), row_less);
    std::vector<char> first((size_t)m, 0);
    for (int i = 0; i < m; ++i)
      if (i == 0 || !row_eq(ord[i - 1], ord[i])) first[(size_t)ord[i]] = 1;      // stable: lowest index of each group
    for (int i = 0; i < m; ++i)
      if (first[(size_t)i]) distinct.push_back(i);
  };
  std::vector<int32_t> pick((size_t)k), pool;
  std::vector<int> init((size_t)k);
  bool use_distinct = nstart >= 2;
  if (nstart == 1) {
    if (k > m) return hs_fail(nullptr, HS_ERR_INVALID, "cannot take a sample larger than the population when 'replace = FALSE'");
    g->sample_int(m, k, pick.data(), pool);
    for (int i = 0; i < k; ++i) init[i] = pick[i] - 1;
    std::vector<int> srt(init);
    std::sort(srt.begin(), srt.end(), row_less);
    for (int i = 1; i < k; ++i)
      if (row_eq(srt[i - 1], srt[i])) use_distinct = true;
  }
  if (use_distinct) {
    build_distinct();
    if ((int)distinct.size() < k) return hs_fail(nullptr, HS_ERR_INVALID, "more cluster centers than distinct data points.");
    g->sample_int((int)distinct.size(), k, pick.data(), pool);
    for (int i = 0; i < k; ++i) init[i] = distinct[(size_t)pick[i] - 1];
  }
  HartiganWong hw;
  hw.a = x; hw.m = m; hw.n = d; hw.k = k;
  const int max_qtran = (int)std::min<int64_t>(2147483647LL, 50LL * m);
  std::vector<double> best_c;
  double best_ss = 0.0;
  int best_fault = 0;
  for (int s = 0; s < nstart; ++s) {
    if (s > 0) {
      g->sample_int((int)distinct.size(), k, pick.data(), pool);
      for (int i = 0; i < k; ++i) init[i] = distinct[(size_t)pick[i] - 1];
    }
    hw.c.assign((size_t)k * d, 0.0);
    for (int l = 0; l < k; ++l)
      for (int j = 0; j < d; ++j) hw.C(l, j) = x[(size_t)init[l] + (size_t)j * m];
    const int fault = hw.run(iter_max, max_qtran);
    if (fault == 1) return hs_fail(nullptr, HS_ERR_INVALID, "empty cluster: try a better set of initial centers");
    if (fault == 3) return hs_fail(nullptr, HS_ERR_INVALID, "number of cluster centres must lie between 1 and nrow(x)");
    double ss = 0.0;
    for (double v : hw.wss) ss += v;
    if (s == 0 || ss < best_ss) {
      best_ss = ss;
      best_c = hw.c;
      best_fault = fault;
    }
  }
  memcpy(centers_out, best_c.data(), sizeof(double) * (size_t)k * d);
  if (tot_withinss_out) *tot_withinss_out = best_ss;
  if (ifault_out) *ifault_out = best_fault;
  return HS_OK;
  } catch (...) { return hs_fail(nullptr, HS_ERR_NOMEM, "hs_r_kmeans: host allocation failed"); }
}

}  // extern "C"
#pragma GCC visibility pop
